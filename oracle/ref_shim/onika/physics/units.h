#pragma once
#include "../../xsref_common.h"
namespace onika { namespace physics {
  // value already expressed in internal units by the test harness
  struct Quantity { double v = 0.0; double convert() const { return v; } };
} }
