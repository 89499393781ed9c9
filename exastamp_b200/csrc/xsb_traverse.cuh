// xsb_traverse.cuh -- the compute_cell_particle_pairs traversal as a device-side
// pattern: a group of TPA consecutive lanes owns one central atom and strides over its CSR neighbour list;
// partial sums are combined with xor-shuffles inside the group (no atomics: Newton-off, every atom is written
// by exactly one group).
#pragma once
#include "xsb_ctx.h"

namespace xsb
{

template<int TPA>
__device__ __forceinline__ double group_sum(double v)
{
# pragma unroll
  for(int o = TPA / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// `*p += v` by the single writer of an atom as a fire-and-forget reduction (RED.E.ADD.F64 at L2): same IEEE addition as
// load-add-store, but the warp does not wait for the load (ncu: ~9 % of the force pass's stall samples sat on those loads)
__device__ __forceinline__ void red_add(double* p, double v) { asm volatile("red.global.add.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory"); }

struct XForm { double m[9]; };

template<bool XFORM>
__device__ __forceinline__ void apply_xform(const XForm& X, double& dx, double& dy, double& dz)
{
  if( XFORM )
  {
    const double x = X.m[0]*dx + X.m[1]*dy + X.m[2]*dz;
    const double y = X.m[3]*dx + X.m[4]*dy + X.m[5]*dz;
    const double z = X.m[6]*dx + X.m[7]*dy + X.m[8]*dz;
    dx = x; dy = y; dz = z;
  }
}

struct ParticleView
{
  const double* __restrict__ rx; const double* __restrict__ ry; const double* __restrict__ rz;
  const unsigned char* __restrict__ type;
  const unsigned long long* __restrict__ nbh_off; const unsigned* __restrict__ nbh_idx;
  const unsigned* __restrict__ atoms;   // central atoms to process (nullptr = all, identity map)
  unsigned n_atoms;
};

inline XForm make_xform(const xsb_grid_desc& g) { XForm X; for(int i = 0; i < 9; i++) X.m[i] = g.xform[i]; return X; }

template<int TPA> inline unsigned groups_grid(unsigned n_atoms, int block) { return unsigned((uint64_t(n_atoms) * TPA + block - 1) / block); }

} // namespace xsb
