#pragma once
#include "../../xsref_common.h"
namespace onika { namespace math {
  struct Vec3d { double x, y, z; };
  inline Vec3d& operator *= (Vec3d& a, double s) { a.x *= s; a.y *= s; a.z *= s; return a; }
} }
namespace exanb { using onika::math::Vec3d; }
namespace exaStamp { using onika::math::Vec3d; }
