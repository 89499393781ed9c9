// mock of onika/math/basic_types.h: Vec3d, Mat3d (row-major m11..m33), IJK
#pragma once
#include <cstddef>
namespace onika { namespace math {
struct Vec3d { double x = 0, y = 0, z = 0; };
struct Mat3d { double m11 = 1, m12 = 0, m13 = 0, m21 = 0, m22 = 1, m23 = 0, m31 = 0, m32 = 0, m33 = 1; };
struct IJK { long i = 0, j = 0, k = 0; };
} }
