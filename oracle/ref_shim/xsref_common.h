// TEST INFRASTRUCTURE ONLY.  Minimal stand-ins for the un-vendored onika / yaml-cpp headers so that the
// reference's own math headers under /root/reference compile UNMODIFIED into oracle/_ref/libxsref.so.
// Nothing here restates reference code: these are empty shells / trivial typedefs written for this repo.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <iostream>
#include <map>
#include <string>
#include <vector>
#define ONIKA_HOST_DEVICE_FUNC
#define ONIKA_ALWAYS_INLINE inline
#define ONIKA_CU_ABORT() std::abort()
// internal units: angstrom, Da, ps (include/exaStamp/unit_system.h:28-36) -- CODATA 2018
#define XSREF_EV_INTERNAL (1.602176634e-19 / (1.66053906660e-27 * 1.0e4))
