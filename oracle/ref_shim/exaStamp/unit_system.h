#pragma once
#include "../xsref_common.h"
#include <onika/physics/units.h>
// EXASTAMP_CONST_QUANTITY( 1. * eV * ang ) -> value in internal units
namespace xsref_units { static constexpr double eV = XSREF_EV_INTERNAL; static constexpr double ang = 1.0; }
#define EXASTAMP_CONST_QUANTITY(...) ([]() constexpr { using namespace ::xsref_units; return double(__VA_ARGS__); }())
#define EXASTAMP_QUANTITY(...) EXASTAMP_CONST_QUANTITY(__VA_ARGS__)
