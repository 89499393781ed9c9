#pragma once
#include "../xsref_common.h"
namespace onika { struct NullLog { template<class T> NullLog& operator<<(const T&) { return *this; } NullLog& operator<<(std::ostream&(*)(std::ostream&)) { return *this; } };
  static NullLog lout, lerr, ldbg; }
namespace exanb { using onika::lout; using onika::lerr; using onika::ldbg; }
