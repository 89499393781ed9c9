#pragma once
#include "../../xsref_common.h"
namespace onika { namespace cuda {
  template<class T> struct ReadOnlyShallowCopyType { using type = T; };
  template<class T> using ro_shallow_copy_t = typename ReadOnlyShallowCopyType<T>::type;
} }
