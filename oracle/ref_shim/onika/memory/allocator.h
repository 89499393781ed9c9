#pragma once
#include "../../xsref_common.h"
namespace onika { namespace memory { template<class T> using CudaMMVector = std::vector<T>; } }
