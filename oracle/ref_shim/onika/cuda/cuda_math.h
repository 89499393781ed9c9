#pragma once
#include "../../xsref_common.h"
#include <algorithm>
namespace onika { namespace cuda { using std::min; using std::max; } }
