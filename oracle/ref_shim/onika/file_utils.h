#pragma once
#include "../xsref_common.h"
namespace onika { inline std::string data_file_path(const std::string& p) { return p; } }
