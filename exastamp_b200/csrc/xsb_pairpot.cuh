// xsb_pairpot.cuh -- per type-pair coefficient records and the evaluation of the pair potentials behind <pot>_compute_force /
// <pot>_multi_force (xsb_pair.cu); shared with xsb_eam.cu, whose force pass can evaluate a chained pair operator on the pairs it
// visits anyway (xsb_ctx::pending_eam).
#pragma once
#include "xsb_ctx.h"

namespace xsb
{

// per type-pair coefficients of one pair potential: everything that does not depend on r, precomputed on the host
//   lj         k = { 4 eps, 24 eps, sigma^2 }                                  (lennard_jones.h:40-50)
//   zbl        k = { d1a, d2a, d3a, d4a, zze, sw1, sw2, sw3, sw4, sw5, r1, - }   (zbl/potential.h:180-303)
//   exp6       k = { A, B, C, D }                                               (exp6.h:66-84)
//   buckingham k = { A, Rho, C }                                                (buckingham.h:41-52)
//   yukawa k = { A, kappa } (yukawa.h:39-48); relax k = { r1, rc } (relax/potential.h:43-51); zero: no parameters
struct LJPair { double k[12]; double ecut, rcut2, rc2_pot; int pot, pad_; };   // ecut = e(rcut) ; rc2_pot: zbl's own rc^2

struct LJMulti { LJPair pp[16]; };   // indexed by unique_pair_id (MAX_TYPE_PAIR_IDS = 16, multiparam.h:68)

__host__ __device__ inline unsigned unique_pair_id(unsigned a, unsigned b) { return a > b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }

constexpr double XSB_EV_INTERNAL = 1.602176634e-19 / (1.66053906660e-27 * 1.0e4);   // 1e-4 e / amu (zbl/potential.h:298)

__host__ __device__ __forceinline__ float  xexp(float x)  { return expf(x); }
__host__ __device__ __forceinline__ double xexp(double x) { return exp(x); }
__host__ __device__ __forceinline__ float  xsqrt(float x)  { return sqrtf(x); }
__host__ __device__ __forceinline__ double xsqrt(double x) { return sqrt(x); }

// pair energy (cut-off shift applied) and de/r from d2.  lj needs no square root: e = 4 eps (s^12 - s^6),
// de/r = -24 eps (2 s^12 - s^6) / r^2 (algebraically identical to lj_compute_energy followed by de/r).
template<class real>
__host__ __device__ __forceinline__ void lj_eval(const LJPair& p, real d2, real& e, real& de_r)
{
  const real rinv2 = real(1) / d2;
  if( p.pot == XSB_POT_LJ )
  {
    const real s2 = real(p.k[2]) * rinv2;
    const real s6 = s2 * s2 * s2;
    const real s12 = s6 * s6;
    e = real(p.k[0]) * (s12 - s6) - real(p.ecut);
    de_r = -real(p.k[1]) * (real(2) * s12 - s6) * rinv2;
    return;
  }
  const real r = xsqrt(d2), rinv = real(1) / r;
  real ee = real(0), de = real(0);
  if( p.pot == XSB_POT_ZBL )
  {
    if( d2 < real(p.rc2_pot) )
    {
      const real e1 = xexp(-real(p.k[0]) * r), e2 = xexp(-real(p.k[1]) * r), e3 = xexp(-real(p.k[2]) * r), e4 = xexp(-real(p.k[3]) * r);
      real sum = real(0.02817) * e1; sum += real(0.28022) * e2; sum += real(0.50986) * e3; sum += real(0.18175) * e4;
      real sum_p = -real(0.02817) * real(p.k[0]) * e1; sum_p -= real(0.28022) * real(p.k[1]) * e2; sum_p -= real(0.50986) * real(p.k[2]) * e3; sum_p -= real(0.18175) * real(p.k[3]) * e4;
      const real zze = real(p.k[4]);
      de = zze * (sum_p - sum * rinv) * rinv;
      ee = zze * sum * rinv + real(p.k[9]);
      const real r1 = real(p.k[10]);
      if( d2 > r1 * r1 )
      {
        const real t = r - r1;
        de += t * t * (real(p.k[5]) + real(p.k[6]) * t);
        ee += t * t * t * (real(p.k[7]) + real(p.k[8]) * t);
      }
    }
    ee *= real(XSB_EV_INTERNAL); de *= real(XSB_EV_INTERNAL);
  }
  else if( p.pot == XSB_POT_EXP6 )
  {
    const real one_rB = real(1) / (r * real(p.k[1]));
    const real r6 = d2 * d2 * d2;
    const real Cr6 = real(p.k[2]) / r6;
    const real t12 = real(12) * one_rB, t2 = t12 * t12, t4 = t2 * t2;
    const real D12 = real(p.k[3]) * (t4 * t4 * t4);
    const real Ae = real(p.k[0]) * xexp(-real(p.k[1]) * r);
    ee = Ae - Cr6 + D12;
    de = -real(p.k[1]) * Ae + (real(6) * Cr6 - real(12) * D12) / r;
  }
  else if( p.pot == XSB_POT_BUCKINGHAM )
  {
    const real x6 = d2 * d2 * d2, x7 = x6 * r;
    const real Ae = real(p.k[0]) * xexp(-r / real(p.k[1]));
    ee = Ae - (real(p.k[2]) / x6);
    de = (real(6) * real(p.k[2]) / x7) - (Ae / real(p.k[1]));
  }
  else if( p.pot == XSB_POT_YUKAWA )
  {
    // yukawa.h:39-48, `de` as the reference writes it: e (1/r - kappa)
    ee = (real(p.k[0]) * rinv) * xexp(-real(p.k[1]) * r);
    de = ee * (rinv - real(p.k[1]));
  }
  else if( p.pot == XSB_POT_RELAX )
  {
    // relax/potential.h:43-51: r clamped to [r1, rc], e = rc / r - 1, de = -e
    real rr = r;
    if( rr < real(p.k[0]) ) rr = real(p.k[0]);
    if( rr > real(p.k[1]) ) rr = real(p.k[1]);
    ee = (real(p.k[1]) / rr) - real(1);
    de = -ee;
  }
  // XSB_POT_ZERO (zero/potential.h:49-54): e = de = 0
  e = ee - real(p.ecut);
  de_r = de * rinv;
}

} // namespace xsb

// xsb_eam.cu: the force phase of eam_alloy_force with `chain` (nullable) evaluated on the same pairs
int xsb_internal_eam_force_phase(xsb_ctx* ctx, double rcut, int phases, int flags, const xsb::LJMulti* chain);
