#pragma once
#include "../../xsref_common.h"
namespace onika { namespace physics {
  // onika's own values are not in the reference tree: CODATA 2018 (exact e, recommended amu)
  static constexpr double elementaryCharge = 1.602176634e-19;   // C
  static constexpr double atomicMass = 1.66053906660e-27;       // kg
} }
