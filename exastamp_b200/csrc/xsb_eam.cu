// xsb_eam.cu -- EAM force operators.
//  * johnson_force & co: single-species analytic EAM, two pair passes (rho -> F'(rho) ; force).
//      operator src/potential/eam_potential_template/eam_potential.cu:69-176, functors
//      eam_force_op_singlemat.h:39-171, math src/potential/eam_potentials/johnson/johnson.h:56-166.
//  * eam_alloy_force: multi-species tabulated (setfl) EAM, phases rho / rho2emb / force.
//      operator eam_potential_multimat.cu:65-259, functors eam_force_op_multimat.h:107-343,
//      tables + evaluation src/potential/eam_potentials/eam_alloy/eam_alloy.h:37-313, reader eam_alloy.cpp:66-278.
#include "xsb_tilepass.cuh"
#include "xsb_pairpot.cuh"
#include <climits>
#include <cmath>
#include <fstream>
#include <limits>
#include <string>

namespace xsb
{

// ------------------------------------------------------------------------------------------------
// Johnson analytic EAM
// ------------------------------------------------------------------------------------------------
// The same struct carries the parameters of the other analytic single-species models of eam_potential_template in its first
// slots (reference struct order): model 1 sutton_chen {c, epsilon, a0, n, m}, model 2 vniitf {rmax, rmin, rt0, Ecoh, E0, beta, A,
// Z, n, alpha, D, eta, mu}; the operators only differ in the three functions rho(r), phi(r), F(rho).
template<class real>
struct JohnsonPT
{
  real re, fe, rhoe, alpha, beta, A, B, kappa, lambda, Fn0, Fn1, Fn2, Fn3, F0, F1, F2, F3, Fo, eta;
  int model, pad_;
  __host__ __device__ __forceinline__ real v(int i) const { return (&re)[i]; }
};
typedef JohnsonPT<double> JohnsonP;
// XSB_FLAG_MIXED: the pair functions rho(r), phi(r) in FP32 (parameters rounded once on the host); F(rho), distances and sums FP64
static JohnsonPT<float> johnson_f32(const JohnsonP& p)
{
  JohnsonPT<float> q; for(int i = 0; i < 19; i++) (&q.re)[i] = float(p.v(i));
  q.model = p.model; q.pad_ = 0; return q;
}
__device__ __forceinline__ float  xpow(float x, float y)   { return powf(x, y); }
__device__ __forceinline__ double xpow(double x, double y) { return pow(x, y); }

// c^20 and c^19 by squaring (johnson.h uses pow(c2,20) and 20*c3/c2)
template<class real>
__device__ __forceinline__ void pow20_19(real c, real& c20, real& c19)
{
  const real c2 = c * c, c4 = c2 * c2, c8 = c4 * c4, c16 = c8 * c8;
  c20 = c16 * c4; c19 = c16 * c2 * c;
}

// one term  s*exp(-k(x-1)) / (1+(x-l)^20)  and its derivative wrt r (ire = 1/re)
template<class real>
__device__ __forceinline__ void johnson_term(real s, real k, real l, real x, real ire, real& f, real& df)
{
  const real num = s * xexp(-k * (x - real(1.0)));
  real c20, c19; pow20_19(x - l, c20, c19);
  const real den = real(1.0) + c20, iden = real(1.0) / den;
  f = num * iden;
  df = ire * ((-k * num) * den - num * (real(20.0) * c19)) * iden * iden;
}

template<class real>
__device__ __forceinline__ void johnson_rho(const JohnsonPT<real>& p, real r, real& rho, real& drho)
{
  const real ire = real(1.0) / p.re;
  johnson_term(p.fe, p.beta, p.lambda, r * ire, ire, rho, drho);
}

template<class real>
__device__ __forceinline__ void johnson_phi(const JohnsonPT<real>& p, real r, real& phi, real& dphi)
{
  const real ire = real(1.0) / p.re, x = r * ire;
  real f1, d1, f2, d2;
  johnson_term(p.A, p.alpha, p.kappa, x, ire, f1, d1);
  johnson_term(-p.B, p.beta, p.lambda, x, ire, f2, d2);
  phi = f1 + f2; dphi = d1 + d2;
}

__device__ __forceinline__ void johnson_fEmbed(const JohnsonP& p, double rho, double& f, double& df)
{
  const double rhon = 0.85 * p.rhoe, rho0 = 1.15 * p.rhoe;
  if( rho < rhon )
  {
    const double q1 = rho / rhon - 1., q2 = q1 * q1, q3 = q1 * q2;
    f = p.Fn0 + p.Fn1 * q1 + p.Fn2 * q2 + p.Fn3 * q3;
    df = (p.Fn1 + 2. * p.Fn2 * q1 + 3. * p.Fn3 * q2) / rhon;
  }
  else if( rho < rho0 )
  {
    const double q1 = rho / p.rhoe - 1., q2 = q1 * q1, q3 = q1 * q2;
    f = p.F0 + p.F1 * q1 + p.F2 * q2 + p.F3 * q3;
    df = (p.F1 + 2. * p.F2 * q1 + 3. * p.F3 * q2) / p.rhoe;
  }
  else
  {
    const double rap = rho / p.rhoe;
    const double rpe = pow(rap, p.eta), l = log(rpe);
    f = p.Fo * (1. - l) * rpe;
    df = (-p.eta * rpe + (1. - l) * p.eta * rpe) * p.Fo / (p.rhoe * rap);
  }
}

// sutton_chen.h:33-62
template<class real>
__device__ __forceinline__ void sutton_chen_rho(const JohnsonPT<real>& p, real r, real& rho, real& drho)
{
  rho = xpow(p.v(2) / r, p.v(4));
  drho = -1 * p.v(4) * rho / r;
}
template<class real>
__device__ __forceinline__ void sutton_chen_phi(const JohnsonPT<real>& p, real r, real& phi, real& dphi)
{
  phi = p.v(1) * xpow(p.v(2) / r, p.v(3));
  dphi = -1 * p.v(3) * phi / r;
}
__device__ __forceinline__ void sutton_chen_fEmbed(const JohnsonP& p, double rho, double& f, double& df)
{
  f = -1. * p.v(0) * p.v(1) * sqrt(rho);
  df = 0.5 * f / (rho > 0 ? rho : 0);
}

// vniitf.h:48-125 ; v(0..12) = rmax rmin rt0 Ecoh E0 beta A Z n alpha D eta mu
template<class real>
__device__ __forceinline__ void vniitf_switch(const JohnsonPT<real>& p, real r, real& S, real& dS)
{
  const real x = (p.v(0) - r) / (p.v(0) - p.v(1));
  const real x2 = x * x, x3 = x2 * x;
  S = x2 * x2 * ( -20 * x2 * x + 70 * x2 - 84 * x + 35 );
  dS = (140 * x3 * ( -1 * x3 + 3 * x2 - 3 * x + 1 )) / (p.v(1) - p.v(0));
  if( x < 0 ) { S = real(0.); dS = real(0.); }
  else if( x > 1 ) { S = real(1.); dS = real(0.); }
}
template<class real>
__device__ __forceinline__ void vniitf_rho(const JohnsonPT<real>& p, real r, real& rho, real& drho)
{
  const real irt0 = 1 / p.v(2);
  const real F = xexp(-p.v(5) * (r * irt0 - real(1.0))) / p.v(7), dF = -p.v(5) * F * irt0;
  real S, dS; vniitf_switch(p, r, S, dS);
  rho = F * S; drho = F * dS + S * dF;
}
template<class real>
__device__ __forceinline__ void vniitf_phi(const JohnsonPT<real>& p, real r, real& phi, real& dphi)
{
  const real ir = 1 / r, irt0 = 1 / p.v(2), dr = r * irt0 - real(1.0), dr2 = dr * dr;
  const real alpha = p.v(9), eta = p.v(11), mu = p.v(12);
  const real a = -2 * p.v(3) / p.v(7), b = alpha * alpha * alpha * p.v(10) * p.v(2);
  const real f1  = a * ( 1 + alpha * dr + eta * dr2 + (mu + b * ir) * dr2 * dr );
  const real df1 = a * ( alpha * irt0 + 2 * eta * irt0 * dr + 3 * mu * irt0 * dr2 + b * (3 * irt0 - dr * ir) * dr2 * ir );
  const real f2 = xexp(-alpha * dr), df2 = -alpha * irt0 * f2;
  real S, dS; vniitf_switch(p, r, S, dS);
  phi = (p.v(4) + f1 * f2) * S;
  dphi = (p.v(4) + f1 * f2) * dS + (f1 * df2 + f2 * df1) * S;
}
__device__ __forceinline__ void vniitf_fEmbed(const JohnsonP& p, double rho, double& f, double& df)
{
  if( rho <= 0. ) { f = 0.; df = 0.; return; }
  const double a = pow(rho, p.v(8)), b = p.v(6) * p.v(3) * a, c = log(a);
  f = b * (c - 1);
  df = p.v(8) * b * c / rho;
}

// the model switch is uniform over the launch
template<class real>
__device__ __forceinline__ void eam1_rho(const JohnsonPT<real>& p, real r, real& f, real& df)
{ if( p.model == 0 ) johnson_rho(p, r, f, df); else if( p.model == 1 ) sutton_chen_rho(p, r, f, df); else vniitf_rho(p, r, f, df); }
template<class real>
__device__ __forceinline__ void eam1_phi(const JohnsonPT<real>& p, real r, real& f, real& df)
{ if( p.model == 0 ) johnson_phi(p, r, f, df); else if( p.model == 1 ) sutton_chen_phi(p, r, f, df); else vniitf_phi(p, r, f, df); }
__device__ __forceinline__ void eam1_fEmbed(const JohnsonP& p, double x, double& f, double& df)
{ if( p.model == 0 ) johnson_fEmbed(p, x, f, df); else if( p.model == 1 ) sutton_chen_fEmbed(p, x, f, df); else vniitf_fEmbed(p, x, f, df); }

// pass 1 (EmbOp): rho_i = sum rho(r_ij) ; ep_i += F(rho_i) ; rho_dEmb_i = F'(rho_i)
template<int TPA, bool XFORM>
__global__ void __launch_bounds__(256) johnson_emb_kernel(ParticleView P, XForm X, JohnsonP p, double rcut2,
                                                           double* __restrict__ ep, double* __restrict__ rho_dEmb)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned g = t / TPA, sub = t % TPA;
  const bool valid = g < P.n_atoms;
  unsigned a = 0; unsigned long long e0 = 0, e1 = 0; double xa = 0, ya = 0, za = 0;
  if( valid ) { a = P.atoms ? P.atoms[g] : g; e0 = P.nbh_off[a]; e1 = P.nbh_off[a+1]; xa = P.rx[a]; ya = P.ry[a]; za = P.rz[a]; }
  double srho = 0.0; unsigned cnt = 0;
  for(unsigned long long e = e0 + sub; e < e1; e += TPA)
  {
    const unsigned b = P.nbh_idx[e];
    double dx = P.rx[b] - xa, dy = P.ry[b] - ya, dz = P.rz[b] - za;
    apply_xform<XFORM>(X, dx, dy, dz);
    const double d2 = dx * dx + dy * dy + dz * dz;
    if( d2 <= rcut2 )
    {
      double rho, drho; eam1_rho(p, sqrt(d2), rho, drho);
      srho += rho; ++cnt;
    }
  }
  srho = group_sum<TPA>(srho);
# pragma unroll
  for(int o = TPA / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if( valid && sub == 0 && cnt > 0 )   // the functor only runs for non-empty pair buffers
  {
    double f, df; eam1_fEmbed(p, srho, f, df);
    ep[a] += f; rho_dEmb[a] = df;
  }
}

// pass 2 (ForceOp): de = (rho'(r)(F'_i+F'_j) + phi'(r))/r
template<int TPA, bool XFORM, bool VIRIAL>
__global__ void __launch_bounds__(256) johnson_force_kernel(ParticleView P, XForm X, JohnsonP p, double rcut2, const double* __restrict__ rho_dEmb,
                                                             double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz,
                                                             double* __restrict__ ep, double* __restrict__ vir)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned g = t / TPA, sub = t % TPA;
  const bool valid = g < P.n_atoms;
  unsigned a = 0; unsigned long long e0 = 0, e1 = 0; double xa = 0, ya = 0, za = 0, fpi = 0;
  if( valid ) { a = P.atoms ? P.atoms[g] : g; e0 = P.nbh_off[a]; e1 = P.nbh_off[a+1]; xa = P.rx[a]; ya = P.ry[a]; za = P.rz[a]; fpi = rho_dEmb[a]; }
  double sfx = 0, sfy = 0, sfz = 0, sep = 0;
  double v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0, v6 = 0, v7 = 0, v8 = 0;
  for(unsigned long long e = e0 + sub; e < e1; e += TPA)
  {
    const unsigned b = P.nbh_idx[e];
    double dx = P.rx[b] - xa, dy = P.ry[b] - ya, dz = P.rz[b] - za;
    apply_xform<XFORM>(X, dx, dy, dz);
    const double d2 = dx * dx + dy * dy + dz * dz;
    if( d2 <= rcut2 )
    {
      const double r = sqrt(d2);
      double rho, drho, phi, dphi;
      eam1_rho(p, r, rho, drho);
      eam1_phi(p, r, phi, dphi);
      const double de = (drho * (fpi + rho_dEmb[b]) + dphi) / r;
      const double fex = de * dx, fey = de * dy, fez = de * dz;
      sfx += fex; sfy += fey; sfz += fez; sep += 0.5 * phi;
      if( VIRIAL )
      {
        v0 -= 0.5 * fex * dx; v1 -= 0.5 * fex * dy; v2 -= 0.5 * fex * dz;
        v3 -= 0.5 * fey * dx; v4 -= 0.5 * fey * dy; v5 -= 0.5 * fey * dz;
        v6 -= 0.5 * fez * dx; v7 -= 0.5 * fez * dy; v8 -= 0.5 * fez * dz;
      }
    }
  }
  sfx = group_sum<TPA>(sfx); sfy = group_sum<TPA>(sfy); sfz = group_sum<TPA>(sfz); sep = group_sum<TPA>(sep);
  if( VIRIAL )
  {
    v0 = group_sum<TPA>(v0); v1 = group_sum<TPA>(v1); v2 = group_sum<TPA>(v2);
    v3 = group_sum<TPA>(v3); v4 = group_sum<TPA>(v4); v5 = group_sum<TPA>(v5);
    v6 = group_sum<TPA>(v6); v7 = group_sum<TPA>(v7); v8 = group_sum<TPA>(v8);
  }
  if( valid && sub == 0 )
  {
    fx[a] += sfx; fy[a] += sfy; fz[a] += sfz; ep[a] += sep;
    if( VIRIAL )
    {
      double* v = vir + 9ull * a;
      v[0] += v0; v[1] += v1; v[2] += v2; v[3] += v3; v[4] += v4; v[5] += v5; v[6] += v6; v[7] += v7; v[8] += v8;
    }
  }
}

// ---- the same two passes as functors for the persistent tile kernel (xsb_tilepass.cuh) ----------------------------
template<bool PWO_, class real = double>
struct JohnsonEmbTileOp
{
  static constexpr bool HAS_W = false, TYPES = false, D2_ONLY = true, PW_OUT = PWO_;
  double rcut2; JohnsonP p; double *ep, *rho_dEmb; JohnsonPT<real> q;      // q: the pair functions' copy of p in `real`
  __host__ __device__ size_t table_bytes() const { return 0; }
  __device__ __forceinline__ void load_tables(unsigned char*, int) const {}
  struct Acc { double rho; unsigned cnt; };
  __device__ __forceinline__ void init(Acc& A) const { A.rho = 0.0; A.cnt = 0; }
  __device__ __forceinline__ void start(Acc&, unsigned, unsigned, const StageBuf<false, false>&, const unsigned char*) const {}
  // returns rho'(r): kept per pair for the force pass of the step (PW_OUT), which then skips one exp + one power
  __device__ __forceinline__ double pair_d2(Acc& A, double d2, unsigned, const StageBuf<false, false>&, const unsigned char*) const
  {
    real rho, drho; eam1_rho<real>(q, xsqrt(real(d2)), rho, drho);
    A.rho += double(rho); ++A.cnt;
    return double(drho);
  }
  __device__ __forceinline__ void pair(Acc& A, double, double, double, double d2, unsigned j, const StageBuf<false, false>& B, const unsigned char* t) const { pair_d2(A, d2, j, B, t); }
  template<int TPA> __device__ __forceinline__ void finish(Acc& A, unsigned a, bool valid, unsigned sub) const
  {
    A.rho = group_sum<TPA>(A.rho);
#   pragma unroll
    for(int o = TPA / 2; o > 0; o >>= 1) A.cnt += __shfl_xor_sync(0xffffffffu, A.cnt, o);
    if( valid && sub == 0 && A.cnt > 0 ) { double f, df; eam1_fEmbed(p, A.rho, f, df); ep[a] += f; rho_dEmb[a] = df; }
  }
};

template<bool VIRIAL, bool PWI_, class real = double>
struct JohnsonForceTileOp
{
  static constexpr bool HAS_W = true, TYPES = false, D2_ONLY = false, PW_IN = PWI_;
  double rcut2; JohnsonPT<real> p; double *fx, *fy, *fz, *ep, *vir;
  __host__ __device__ size_t table_bytes() const { return 0; }
  __device__ __forceinline__ void load_tables(unsigned char*, int) const {}
  struct Acc { double fx, fy, fz, ep, fpi; Vir9 v; };
  __device__ __forceinline__ void init(Acc& A) const { A.fx = A.fy = A.fz = A.ep = A.fpi = 0.0; if( VIRIAL ) A.v.zero(); }
  __device__ __forceinline__ void start(Acc& A, unsigned, unsigned sa, const StageBuf<true, false>& B, const unsigned char*) const { A.fpi = B.w[sa]; }
  template<bool HAVE>
  __device__ __forceinline__ void eval(Acc& A, double dx, double dy, double dz, double d2, unsigned j, const StageBuf<true, false>& B, double drho_in) const
  {
    const real r = xsqrt(real(d2));
    real rho, drho = real(drho_in), phi, dphi;
    if( !HAVE ) eam1_rho<real>(p, r, rho, drho);
    eam1_phi<real>(p, r, phi, dphi);
    const double de = double((drho * (real(A.fpi) + real(B.w[j])) + dphi) / r);
    const double fex = de * dx, fey = de * dy, fez = de * dz;
    A.fx += fex; A.fy += fey; A.fz += fez; A.ep += 0.5 * double(phi);
    if( VIRIAL ) A.v.add(fex, fey, fez, dx, dy, dz);
  }
  __device__ __forceinline__ void pair(Acc& A, double dx, double dy, double dz, double d2, unsigned j, const StageBuf<true, false>& B, const unsigned char*) const
  { eval<false>(A, dx, dy, dz, d2, j, B, 0.0); }
  __device__ __forceinline__ void pair_pw(Acc& A, double dx, double dy, double dz, double d2, unsigned j, const StageBuf<true, false>& B, const unsigned char*, double pv, double) const
  { eval<true>(A, dx, dy, dz, d2, j, B, pv); }
  template<int TPA> __device__ __forceinline__ void finish(Acc& A, unsigned a, bool valid, unsigned sub) const
  {
    A.fx = group_sum<TPA>(A.fx); A.fy = group_sum<TPA>(A.fy); A.fz = group_sum<TPA>(A.fz); A.ep = group_sum<TPA>(A.ep);
    if( VIRIAL ) A.v.template reduce<TPA>();
    if( valid && sub == 0 ) { red_add(fx + a, A.fx); red_add(fy + a, A.fy); red_add(fz + a, A.fz); red_add(ep + a, A.ep); if( VIRIAL ) A.v.store_add(vir, a); }
  }
};

// ------------------------------------------------------------------------------------------------
// eam/alloy tabulated EAM
// ------------------------------------------------------------------------------------------------
struct EamAlloyView
{
  const double* __restrict__ frho; const double* __restrict__ rhor; const double* __restrict__ z2r;
  int nr, nrho; double rdr, rdrho, rhomax, conv_z2r, conv_frho;
};

__device__ __forceinline__ int z2r_index(int a, int b) { return a > b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }

// spline row lookup for an r-indexed table (eam_alloy.h:196-200): no lower clamp on m
__device__ __forceinline__ void r_lookup(const EamAlloyView& T, double r, int& m, double& p)
{
  p = r * T.rdr + 1.0;
  m = static_cast<int>(p);
  m = min(m, T.nr - 1);
  p -= m;
  p = fmin(p, 1.0);
}

template<int TPA, bool XFORM>
__global__ void __launch_bounds__(256) eam_alloy_rho_kernel(ParticleView P, XForm X, EamAlloyView T, double rcut2, double* __restrict__ rho_dEmb)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned g = t / TPA, sub = t % TPA;
  const bool valid = g < P.n_atoms;
  unsigned a = 0; unsigned long long e0 = 0, e1 = 0; double xa = 0, ya = 0, za = 0;
  if( valid ) { a = P.atoms ? P.atoms[g] : g; e0 = P.nbh_off[a]; e1 = P.nbh_off[a+1]; xa = P.rx[a]; ya = P.ry[a]; za = P.rz[a]; }
  double srho = 0.0;
  for(unsigned long long e = e0 + sub; e < e1; e += TPA)
  {
    const unsigned b = P.nbh_idx[e];
    double dx = P.rx[b] - xa, dy = P.ry[b] - ya, dz = P.rz[b] - za;
    apply_xform<XFORM>(X, dx, dy, dz);
    const double d2 = dx * dx + dy * dy + dz * dz;
    if( d2 <= rcut2 )
    {
      int m; double p; r_lookup(T, sqrt(d2), m, p);
      const double* c = T.rhor + (size_t(P.type[b]) * (T.nr + 1) + m) * 8;   // density of the NEIGHBOUR's element
      srho += ((c[3] * p + c[4]) * p + c[5]) * p + c[6];
    }
  }
  srho = group_sum<TPA>(srho);
  if( valid && sub == 0 ) rho_dEmb[a] += srho;
}

__global__ void eam_alloy_rho2emb_kernel(const unsigned* __restrict__ atoms, unsigned n_atoms, EamAlloyView T, const unsigned char* __restrict__ type,
                                         double* __restrict__ rho_dEmb, double* __restrict__ ep)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if( t >= n_atoms ) return;
  const unsigned a = atoms ? atoms[t] : t;
  const double rho = rho_dEmb[a];
  double p = rho * T.rdrho + 1.0;
  int m = static_cast<int>(p);
  m = max(1, min(m, T.nrho - 1));
  p -= m;
  p = fmin(p, 1.0);
  const double* c = T.frho + (size_t(type[a]) * (T.nrho + 1) + m) * 8;
  double fp = (c[0] * p + c[1]) * p + c[2];
  double phi = ((c[3] * p + c[4]) * p + c[5]) * p + c[6];
  if( rho > T.rhomax ) phi += fp * (rho - T.rhomax);
  rho_dEmb[a] = fp * T.conv_frho;
  if( ep ) ep[a] += phi * T.conv_frho;
}

template<int TPA, bool XFORM, bool EFLAG, bool VIRIAL>
__global__ void __launch_bounds__(256) eam_alloy_force_kernel(ParticleView P, XForm X, EamAlloyView T, double rcut2, const double* __restrict__ dEmb,
                                                               double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz,
                                                               double* __restrict__ ep, double* __restrict__ vir)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned g = t / TPA, sub = t % TPA;
  const bool valid = g < P.n_atoms;
  unsigned a = 0; unsigned long long e0 = 0, e1 = 0; double xa = 0, ya = 0, za = 0, fpi = 0; int ta = 0;
  if( valid ) { a = P.atoms ? P.atoms[g] : g; e0 = P.nbh_off[a]; e1 = P.nbh_off[a+1]; xa = P.rx[a]; ya = P.ry[a]; za = P.rz[a]; fpi = dEmb[a]; ta = P.type[a]; }
  double sfx = 0, sfy = 0, sfz = 0, sep = 0;
  double v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0, v6 = 0, v7 = 0, v8 = 0;
  for(unsigned long long e = e0 + sub; e < e1; e += TPA)
  {
    const unsigned b = P.nbh_idx[e];
    double dx = P.rx[b] - xa, dy = P.ry[b] - ya, dz = P.rz[b] - za;
    apply_xform<XFORM>(X, dx, dy, dz);
    const double d2 = dx * dx + dy * dy + dz * dz;
    if( d2 <= rcut2 )
    {
      const double r = sqrt(d2);
      int m; double p; r_lookup(T, r, m, p);
      const int tb = P.type[b];
      const double* ci = T.rhor + (size_t(ta) * (T.nr + 1) + m) * 8;
      const double* cj = T.rhor + (size_t(tb) * (T.nr + 1) + m) * 8;
      const double* cz = T.z2r + (size_t(z2r_index(ta, tb)) * (T.nr + 1) + m) * 8;
      const double rhoip = (ci[0] * p + ci[1]) * p + ci[2];
      const double rhojp = (cj[0] * p + cj[1]) * p + cj[2];
      const double z2p = (cz[0] * p + cz[1]) * p + cz[2];
      const double z2 = ((cz[3] * p + cz[4]) * p + cz[5]) * p + cz[6];
      const double recip = 1.0 / r;
      double phi = z2 * recip;
      const double phip = (z2p * recip - phi * recip) * T.conv_z2r;
      phi *= T.conv_z2r;
      const double fpair = (fpi * rhojp + dEmb[b] * rhoip + phip) * recip;
      const double fex = dx * fpair, fey = dy * fpair, fez = dz * fpair;
      sfx += fex; sfy += fey; sfz += fez;
      if( EFLAG ) sep += 0.5 * phi;
      if( VIRIAL )
      {
        v0 -= 0.5 * fex * dx; v1 -= 0.5 * fex * dy; v2 -= 0.5 * fex * dz;
        v3 -= 0.5 * fey * dx; v4 -= 0.5 * fey * dy; v5 -= 0.5 * fey * dz;
        v6 -= 0.5 * fez * dx; v7 -= 0.5 * fez * dy; v8 -= 0.5 * fez * dz;
      }
    }
  }
  sfx = group_sum<TPA>(sfx); sfy = group_sum<TPA>(sfy); sfz = group_sum<TPA>(sfz);
  if( EFLAG ) sep = group_sum<TPA>(sep);
  if( VIRIAL )
  {
    v0 = group_sum<TPA>(v0); v1 = group_sum<TPA>(v1); v2 = group_sum<TPA>(v2);
    v3 = group_sum<TPA>(v3); v4 = group_sum<TPA>(v4); v5 = group_sum<TPA>(v5);
    v6 = group_sum<TPA>(v6); v7 = group_sum<TPA>(v7); v8 = group_sum<TPA>(v8);
  }
  if( valid && sub == 0 )
  {
    fx[a] += sfx; fy[a] += sfy; fz[a] += sfz;
    if( EFLAG ) ep[a] += sep;
    if( VIRIAL )
    {
      double* v = vir + 9ull * a;
      v[0] += v0; v[1] += v1; v[2] += v2; v[3] += v3; v[4] += v4; v[5] += v5; v[6] += v6; v[7] += v7; v[8] += v8;
    }
  }
}

// ---- tile path: the r-tables as Hermite knots {f, c5} (16 B/row instead of 64 B) ------------------------------------
// The reference rows hold c6=f[m], c5=f'[m]*delta and c4, c3, c2..c0 derived from (f[m], f[m+1], c5[m], c5[m+1])
// (interpolate(), eam_alloy.cpp:29-58).  Re-deriving c4, c3 per pair costs 5 FP64 ops and shrinks a table 4x, so the
// window of rows a simulation can touch (r >= ~0.9 r_min) fits in shared memory next to the stage buffers:
// a 5000-row table is 61 KiB instead of 320 KiB.  Values agree with the 7-coefficient rows to rounding (~1e-16 rel).
struct EamFcView
{
  const double2* __restrict__ g;   // global {f, c5}: [ntab][nr+1], tables ordered rhor[nel] then z2r[npairs]
  int nr, m_lo, rows, ntab_smem;   // smem window = rows [m_lo, m_lo+rows) of the tables [t0, t0+ntab_smem) (rows = 0: none)
  int t0;
  double rdr;
  __host__ __device__ size_t table_bytes() const { return size_t(rows) * size_t(ntab_smem) * sizeof(double2); }
  __device__ __forceinline__ void load(unsigned char* smem, int nt) const
  {
    double2* sm = reinterpret_cast<double2*>(smem);
    const int tot = rows * ntab_smem;
    for(int i = threadIdx.x; i < tot; i += nt) { const int t = i / rows, r = i - t * rows; sm[i] = g[size_t(t0 + t) * (nr + 1) + m_lo + r]; }
  }
  __device__ __forceinline__ void lookup(double r, int& m, double& p) const
  {
    p = r * rdr + 1.0;
    m = __double2int_rz(p);
    m = min(m, nr - 1);
    p -= m;
    p = fmin(p, 1.0);
  }
  // knots m and m+1 of table t
  __device__ __forceinline__ void knots(const unsigned char* smem, int t, int m, double2& k0, double2& k1) const
  {
    if( m >= m_lo && t >= t0 ) { const double2* sm = reinterpret_cast<const double2*>(smem) + (t - t0) * rows + (m - m_lo); k0 = sm[0]; k1 = sm[1]; }
    else { const double2* q = g + size_t(t) * (nr + 1) + m; k0 = q[0]; k1 = q[1]; }
  }
};

__device__ __forceinline__ void hermite_c(const double2 k0, const double2 k1, double& c3, double& c4)
{
  const double df = k1.x - k0.x;
  c4 = 3.0 * df - 2.0 * k0.y - k1.y;
  c3 = k0.y + k1.y - 2.0 * df;
}

template<bool MULTI, bool PWO_>
struct EamRhoTileOp
{
  static constexpr bool HAS_W = false, TYPES = MULTI, D2_ONLY = true, PW_OUT = PWO_;
  static constexpr int PW_N = MULTI ? 2 : 1;
  double rcut2; EamFcView T; double* rho_dEmb;
  __host__ __device__ size_t table_bytes() const { return T.table_bytes(); }
  __device__ __forceinline__ void load_tables(unsigned char* smem, int nt) const { T.load(smem, nt); }
  struct Acc { double rho; int ta; };
  __device__ __forceinline__ void init(Acc& A) const { A.rho = 0.0; A.ta = 0; }
  __device__ __forceinline__ void start(Acc& A, unsigned, unsigned sa, const StageBuf<HAS_W, TYPES>& B, const unsigned char*) const { if( MULTI ) A.ta = B.t[sa]; }
  __device__ __forceinline__ void pair(Acc& A, double, double, double, double d2, unsigned j, const StageBuf<HAS_W, TYPES>& B, const unsigned char* tab) const { pair_d2(A, d2, j, B, tab); }
  // Returns what the force pass of the same step needs again for this pair (PW_OUT: the traversal keeps it next to the
  // sub-list entry, so that pass does not fetch the density knots again): rho'(r) of the neighbour's element (rhojp
  // there) and, for a multi-element system, also rho'(r) of the central atom's element (rhoip).
  __device__ __forceinline__ auto pair_d2(Acc& A, double d2, unsigned j, const StageBuf<HAS_W, TYPES>& B, const unsigned char* tab) const
  {
    // a sub-list with an inner skin also holds pairs just beyond rcut: they contribute nothing and are marked dead (NaN)
    const bool live = d2 <= rcut2;
    const double r = d2 * rsqrt(d2);
    int m; double p; T.lookup(r, m, p);
    const int tb = MULTI ? int(B.t[j]) : 0;
    double2 k0, k1; T.knots(tab, tb, m, k0, k1);    // density table of the NEIGHBOUR's element
    double c3, c4; hermite_c(k0, k1, c3, c4);
    const double val = ((c3 * p + c4) * p + k0.y) * p + k0.x;
    A.rho += live ? val : 0.0;
    const double rhojp = !live ? pw_dead() : (PWO_ ? ((3.0 * c3 * p + 2.0 * c4) * p + k0.y) * T.rdr : 0.0);
    if constexpr ( MULTI )
    {
      double rhoip = rhojp;
      if( PWO_ && live && tb != A.ta ) { T.knots(tab, A.ta, m, k0, k1); hermite_c(k0, k1, c3, c4); rhoip = ((3.0 * c3 * p + 2.0 * c4) * p + k0.y) * T.rdr; }
      return make_double2(rhojp, rhoip);
    }
    else return rhojp;
  }
  template<int TPA> __device__ __forceinline__ void finish(Acc& A, unsigned a, bool valid, unsigned sub) const
  {
    A.rho = group_sum<TPA>(A.rho);
    if( valid && sub == 0 ) red_add(rho_dEmb + a, A.rho);
  }
};

// A pair operator chained behind the EAM operator (compute_force: [eam_alloy_force, lj_multi_force], xsb_ctx::pending_eam)
// rides along in the force pass: CHAIN adds its table and one lj_eval per visited pair inside its own cut-off.
struct NoChain {};
template<bool MULTI, bool CHAIN, class real>
__device__ __forceinline__ void chain_eval(const std::conditional_t<CHAIN, LJMulti, NoChain>& prm, unsigned ta, unsigned tb, double d2, double& fpair, double& e_half)
{
  if constexpr ( CHAIN )
  {
    const LJPair& pp = MULTI ? prm.pp[unique_pair_id(ta, tb)] : prm.pp[0];
    if( d2 <= pp.rcut2 )
    {
      // lennard_jones.h:40-50 as in lj_eval (only Lennard-Jones operators are taken along: the other potentials' exp / sqrt
      // branches do not fit the 64 registers of the 1024-thread passes)
      const real rinv2 = real(1) / real(d2);
      const real s2 = real(pp.k[2]) * rinv2;
      const real s6 = s2 * s2 * s2;
      const real s12 = s6 * s6;
      fpair += double(-real(pp.k[1]) * (real(2) * s12 - s6) * rinv2);
      e_half += 0.5 * double(real(pp.k[0]) * (s12 - s6) - real(pp.ecut));
    }
  }
}

template<bool MULTI, bool EFLAG, bool VIRIAL, bool PWI_, bool CHAIN = false>
struct EamForceTileOp
{
  static constexpr bool HAS_W = true, TYPES = MULTI, D2_ONLY = false, PW_IN = PWI_;
  static constexpr bool NO_AHEAD = !VIRIAL;      // 64 registers at 1024 threads: the prefetch state would spill (measured 1.79 -> 1.88 ms)
  static constexpr int PW_N = MULTI ? 2 : 1;
  double rcut2; EamFcView T; int nel; double conv_z2r;
  double *fx, *fy, *fz, *ep, *vir;
  std::conditional_t<CHAIN, LJMulti, NoChain> chain;
  __host__ __device__ size_t table_bytes() const { return T.table_bytes(); }
  __device__ __forceinline__ void load_tables(unsigned char* smem, int nt) const { T.load(smem, nt); }
  struct Acc { double fx, fy, fz, ep, fpi; int ta; Vir9 v; };
  __device__ __forceinline__ void init(Acc& A) const { A.fx = A.fy = A.fz = A.ep = A.fpi = 0.0; A.ta = 0; if( VIRIAL ) A.v.zero(); }
  __device__ __forceinline__ void start(Acc& A, unsigned, unsigned sa, const StageBuf<HAS_W, TYPES>& B, const unsigned char*) const
  { A.fpi = B.w[sa]; if( MULTI ) A.ta = B.t[sa]; }
  // HAVE: rhojp_in / rhoip_in = rho'(r) of the neighbour's / the central atom's element, cached by the rho pass of this
  // step (same arithmetic): only the pair table z2r is looked up here
  template<bool HAVE>
  __device__ __forceinline__ void eval(Acc& A, double dx, double dy, double dz, double d2, unsigned j, const StageBuf<HAS_W, TYPES>& B, const unsigned char* tab, double rhojp_in, double rhoip_in) const
  {
    const double recip = rsqrt(d2), r = d2 * recip;
    int m; double p; T.lookup(r, m, p);
    const int tb = MULTI ? int(B.t[j]) : 0;
    double2 k0, k1; double c3, c4;
    double rhoip, rhojp;
    if( HAVE )
    {
      rhojp = rhojp_in; rhoip = MULTI ? rhoip_in : rhojp_in;
    }
    else
    {
      T.knots(tab, A.ta, m, k0, k1); hermite_c(k0, k1, c3, c4);
      rhoip = ((3.0 * c3 * p + 2.0 * c4) * p + k0.y) * T.rdr;
      rhojp = rhoip;
      if( MULTI && tb != A.ta ) { T.knots(tab, tb, m, k0, k1); hermite_c(k0, k1, c3, c4); rhojp = ((3.0 * c3 * p + 2.0 * c4) * p + k0.y) * T.rdr; }
    }
    T.knots(tab, nel + z2r_index(A.ta, tb), m, k0, k1); hermite_c(k0, k1, c3, c4);
    const double z2p = ((3.0 * c3 * p + 2.0 * c4) * p + k0.y) * T.rdr;
    const double z2 = ((c3 * p + c4) * p + k0.y) * p + k0.x;
    double phi = z2 * recip;
    const double phip = (z2p * recip - phi * recip) * conv_z2r;
    phi *= conv_z2r;
    double fpair = (A.fpi * rhojp + B.w[j] * rhoip + phip) * recip;
    double eh = 0.5 * phi;
    chain_eval<MULTI, CHAIN, double>(chain, A.ta, tb, d2, fpair, eh);
    const double fex = dx * fpair, fey = dy * fpair, fez = dz * fpair;
    A.fx += fex; A.fy += fey; A.fz += fez;
    if( EFLAG ) A.ep += eh;
    if( VIRIAL ) A.v.add(fex, fey, fez, dx, dy, dz);
  }
  __device__ __forceinline__ void pair(Acc& A, double dx, double dy, double dz, double d2, unsigned j, const StageBuf<HAS_W, TYPES>& B, const unsigned char* tab) const
  { eval<false>(A, dx, dy, dz, d2, j, B, tab, 0.0, 0.0); }
  __device__ __forceinline__ void pair_pw(Acc& A, double dx, double dy, double dz, double d2, unsigned j, const StageBuf<HAS_W, TYPES>& B, const unsigned char* tab, double pv, double pv2) const
  { eval<true>(A, dx, dy, dz, d2, j, B, tab, pv, pv2); }
  template<int TPA> __device__ __forceinline__ void finish(Acc& A, unsigned a, bool valid, unsigned sub) const
  {
    A.fx = group_sum<TPA>(A.fx); A.fy = group_sum<TPA>(A.fy); A.fz = group_sum<TPA>(A.fz);
    if( EFLAG ) A.ep = group_sum<TPA>(A.ep);
    if( VIRIAL ) A.v.template reduce<TPA>();
    if( valid && sub == 0 ) { red_add(fx + a, A.fx); red_add(fy + a, A.fy); red_add(fz + a, A.fz); if( EFLAG ) red_add(ep + a, A.ep); if( VIRIAL ) A.v.store_add(vir, a); }
  }
};

// ---- mixed precision (XSB_FLAG_MIXED, tolerance 1e-5): FP64 positions, distances and accumulation as before, the spline
// lookup and the pair expression in FP32.  A row of the FP32 table holds the four cubic coefficients of its interval
// {c6 = f[m], c5, c4, c3} (16 B; derived in FP64 on the host -- re-deriving c4, c3 from FP32 knots would cancel to ~1e-4),
// so a lookup is one 16-byte shared-memory load instead of two, and the FP64 rsqrt + ~30 FP64 operations per pair
// become FP32.
struct EamFcView32
{
  const float4* __restrict__ g;    // global [ntab][nr+1]
  int nr, m_lo, rows, ntab_smem, t0;
  float rdr;
  __host__ __device__ size_t table_bytes() const { return size_t(rows) * size_t(ntab_smem) * sizeof(float4); }
  __device__ __forceinline__ void load(unsigned char* smem, int nt) const
  {
    float4* sm = reinterpret_cast<float4*>(smem);
    const int tot = rows * ntab_smem;
    for(int i = threadIdx.x; i < tot; i += nt) { const int t = i / rows, r = i - t * rows; sm[i] = g[size_t(t0 + t) * (nr + 1) + m_lo + r]; }
  }
  __device__ __forceinline__ void lookup(float r, int& m, float& p) const
  {
    p = r * rdr + 1.0f;
    m = __float2int_rz(p);
    m = min(m, nr - 1);
    p -= float(m);
    p = fminf(p, 1.0f);
  }
  __device__ __forceinline__ float4 knots(const unsigned char* smem, int t, int m) const
  {
    if( m >= m_lo && t >= t0 ) return reinterpret_cast<const float4*>(smem)[(t - t0) * rows + (m - m_lo)];
    return g[size_t(t) * (nr + 1) + m];
  }
};

__device__ __forceinline__ void hermite_c32(const float4 k, float& c3, float& c4) { c4 = k.z; c3 = k.w; }

template<bool MULTI, bool PWO_>
struct EamRhoTileOp32
{
  static constexpr bool HAS_W = false, TYPES = MULTI, D2_ONLY = true, PW_OUT = PWO_;
  static constexpr int PW_N = MULTI ? 2 : 1;
  double rcut2; EamFcView32 T; double* rho_dEmb;
  __host__ __device__ size_t table_bytes() const { return T.table_bytes(); }
  __device__ __forceinline__ void load_tables(unsigned char* smem, int nt) const { T.load(smem, nt); }
  struct Acc { double rho; int ta; };
  __device__ __forceinline__ void init(Acc& A) const { A.rho = 0.0; A.ta = 0; }
  __device__ __forceinline__ void start(Acc& A, unsigned, unsigned sa, const StageBuf<HAS_W, TYPES>& B, const unsigned char*) const { if( MULTI ) A.ta = B.t[sa]; }
  __device__ __forceinline__ void pair(Acc& A, double, double, double, double d2, unsigned j, const StageBuf<HAS_W, TYPES>& B, const unsigned char* tab) const { pair_d2(A, d2, j, B, tab); }
  __device__ __forceinline__ auto pair_d2(Acc& A, double d2, unsigned j, const StageBuf<HAS_W, TYPES>& B, const unsigned char* tab) const
  {
    const bool live = d2 <= rcut2;
    const float d2f = float(d2), r = d2f * rsqrtf(d2f);
    int m; float p; T.lookup(r, m, p);
    const int tb = MULTI ? int(B.t[j]) : 0;
    float4 k = T.knots(tab, tb, m);
    float c3, c4; hermite_c32(k, c3, c4);
    A.rho += live ? double(((c3 * p + c4) * p + k.y) * p + k.x) : 0.0;
    const double rhojp = !live ? pw_dead() : (PWO_ ? double(((3.0f * c3 * p + 2.0f * c4) * p + k.y) * T.rdr) : 0.0);
    if constexpr ( MULTI )
    {
      double rhoip = rhojp;
      if( PWO_ && live && tb != A.ta ) { k = T.knots(tab, A.ta, m); hermite_c32(k, c3, c4); rhoip = double(((3.0f * c3 * p + 2.0f * c4) * p + k.y) * T.rdr); }
      return make_double2(rhojp, rhoip);
    }
    else return rhojp;
  }
  template<int TPA> __device__ __forceinline__ void finish(Acc& A, unsigned a, bool valid, unsigned sub) const
  {
    A.rho = group_sum<TPA>(A.rho);
    if( valid && sub == 0 ) red_add(rho_dEmb + a, A.rho);
  }
};

template<bool MULTI, bool EFLAG, bool VIRIAL, bool PWI_, bool CHAIN = false>
struct EamForceTileOp32
{
  static constexpr bool HAS_W = true, TYPES = MULTI, D2_ONLY = false, PW_IN = PWI_;
  static constexpr int PW_N = MULTI ? 2 : 1;
  double rcut2; EamFcView32 T; int nel; float conv_z2r;
  double *fx, *fy, *fz, *ep, *vir;
  std::conditional_t<CHAIN, LJMulti, NoChain> chain;
  __host__ __device__ size_t table_bytes() const { return T.table_bytes(); }
  __device__ __forceinline__ void load_tables(unsigned char* smem, int nt) const { T.load(smem, nt); }
  struct Acc { double fx, fy, fz, ep, fpi; int ta; Vir9 v; };
  __device__ __forceinline__ void init(Acc& A) const { A.fx = A.fy = A.fz = A.ep = A.fpi = 0.0; A.ta = 0; if( VIRIAL ) A.v.zero(); }
  __device__ __forceinline__ void start(Acc& A, unsigned, unsigned sa, const StageBuf<HAS_W, TYPES>& B, const unsigned char*) const
  { A.fpi = B.w[sa]; if( MULTI ) A.ta = B.t[sa]; }
  template<bool HAVE>
  __device__ __forceinline__ void eval(Acc& A, double dx, double dy, double dz, double d2, unsigned j, const StageBuf<HAS_W, TYPES>& B, const unsigned char* tab, double rhojp_in, double rhoip_in) const
  {
    const float d2f = float(d2), recip = rsqrtf(d2f), r = d2f * recip;
    int m; float p; T.lookup(r, m, p);
    const int tb = MULTI ? int(B.t[j]) : 0;
    float4 k; float c3, c4, rhoip, rhojp;
    if( HAVE ) { rhojp = float(rhojp_in); rhoip = MULTI ? float(rhoip_in) : rhojp; }
    else
    {
      k = T.knots(tab, A.ta, m); hermite_c32(k, c3, c4);
      rhoip = ((3.0f * c3 * p + 2.0f * c4) * p + k.y) * T.rdr;
      rhojp = rhoip;
      if( MULTI && tb != A.ta ) { k = T.knots(tab, tb, m); hermite_c32(k, c3, c4); rhojp = ((3.0f * c3 * p + 2.0f * c4) * p + k.y) * T.rdr; }
    }
    k = T.knots(tab, nel + z2r_index(A.ta, tb), m); hermite_c32(k, c3, c4);
    const float z2p = ((3.0f * c3 * p + 2.0f * c4) * p + k.y) * T.rdr;
    const float z2 = ((c3 * p + c4) * p + k.y) * p + k.x;
    float phi = z2 * recip;
    const float phip = (z2p * recip - phi * recip) * conv_z2r;
    phi *= conv_z2r;
    double fpair = double((float(A.fpi) * rhojp + float(B.w[j]) * rhoip + phip) * recip);
    double eh = 0.5 * double(phi);
    chain_eval<MULTI, CHAIN, float>(chain, A.ta, tb, d2, fpair, eh);
    const double fex = dx * fpair, fey = dy * fpair, fez = dz * fpair;
    A.fx += fex; A.fy += fey; A.fz += fez;
    if( EFLAG ) A.ep += eh;
    if( VIRIAL ) A.v.add(fex, fey, fez, dx, dy, dz);
  }
  __device__ __forceinline__ void pair(Acc& A, double dx, double dy, double dz, double d2, unsigned j, const StageBuf<HAS_W, TYPES>& B, const unsigned char* tab) const
  { eval<false>(A, dx, dy, dz, d2, j, B, tab, 0.0, 0.0); }
  __device__ __forceinline__ void pair_pw(Acc& A, double dx, double dy, double dz, double d2, unsigned j, const StageBuf<HAS_W, TYPES>& B, const unsigned char* tab, double pv, double pv2) const
  { eval<true>(A, dx, dy, dz, d2, j, B, tab, pv, pv2); }
  template<int TPA> __device__ __forceinline__ void finish(Acc& A, unsigned a, bool valid, unsigned sub) const
  {
    A.fx = group_sum<TPA>(A.fx); A.fy = group_sum<TPA>(A.fy); A.fz = group_sum<TPA>(A.fz);
    if( EFLAG ) A.ep = group_sum<TPA>(A.ep);
    if( VIRIAL ) A.v.template reduce<TPA>();
    if( valid && sub == 0 ) { red_add(fx + a, A.fx); red_add(fy + a, A.fy); red_add(fz + a, A.fz); if( EFLAG ) red_add(ep + a, A.ep); if( VIRIAL ) A.v.store_add(vir, a); }
  }
};

// first EAM pass of a step, device-side choice between re-filtering the neighbour list and re-evaluating the sub-list
__global__ void sub_decide_kernel(SubCtl* ctl, int force_build, double half_skin)
{
  const bool build = force_build || !(ctl->acc <= half_skin);
  ctl->mode = build ? 0 : 1;
  if( build ) { ctl->acc = 0.0; ctl->builds++; } else ctl->reuses++;
}

template<bool HAS_W, bool TYPES>
static EamFcView32 make_fc_view32(const xsb_ctx* ctx, int t0, int ntab, size_t queue_bytes)
{
  const EamAlloyDev& E = ctx->eam;
  EamFcView32 T{ reinterpret_cast<const float4*>(E.fc32.p), E.nr, INT_MAX, 0, ntab, t0, float(E.rdr) };
  const size_t fixed = tile_smem_bytes<HAS_W, TYPES>(ctx->tile_s_cap, 0, queue_bytes, 2) + 256;
  if( fixed >= TILE_SMEM_MAX ) return T;
  const size_t max_rows = (TILE_SMEM_MAX - fixed) / (sizeof(float4) * size_t(ntab));
  const double rmin = 0.9 * std::sqrt(ctx->nbh_d2min > 0.0 ? ctx->nbh_d2min : 0.0);
  int m_lo = std::max(1, int(rmin * E.rdr + 1.0) - 1);
  int rows = E.nr + 1 - m_lo;
  if( rows < 2 ) return T;
  if( size_t(rows) > max_rows ) { rows = int(max_rows); m_lo = E.nr + 1 - rows; }
  if( rows < 64 ) return T;
  T.m_lo = m_lo; T.rows = rows;
  return T;
}

// shared-memory window of the {f,c5} tables for a tile pass: rows [m_lo, nr] of the first ntab tables, as large a
// window as the 227 KiB budget allows next to the 2 stage buffers (pairs below the window read the global copy)
template<bool HAS_W, bool TYPES>
static EamFcView make_fc_view(const xsb_ctx* ctx, int t0, int ntab, size_t queue_bytes)
{
  const EamAlloyDev& E = ctx->eam;
  EamFcView T{ reinterpret_cast<const double2*>(E.fc.p), E.nr, INT_MAX, 0, ntab, t0, E.rdr };
  const size_t fixed = tile_smem_bytes<HAS_W, TYPES>(ctx->tile_s_cap, 0, queue_bytes, 2) + 256;
  if( fixed >= TILE_SMEM_MAX ) return T;
  const size_t max_rows = (TILE_SMEM_MAX - fixed) / (sizeof(double2) * size_t(ntab));
  // rows a pair can reach: m >= floor(r_min * rdr + 1); keep a 10 % margin in r for the drift until the next rebuild
  const double rmin = 0.9 * std::sqrt(ctx->nbh_d2min > 0.0 ? ctx->nbh_d2min : 0.0);
  int m_lo = std::max(1, int(rmin * E.rdr + 1.0) - 1);
  int rows = E.nr + 1 - m_lo;
  if( rows < 2 ) return T;
  if( size_t(rows) > max_rows ) { rows = int(max_rows); m_lo = E.nr + 1 - rows; }   // keep the far end (most pairs are at large r)
  if( rows < 64 ) return T;
  T.m_lo = m_lo; T.rows = rows;
  return T;
}

// LAMMPS-style 7-coefficient spline rows (eam_alloy.cpp:29-58), rows padded to 8 doubles, row 0 unused
static void interpolate(int n, double delta, const double* f, double* s)
{
  auto S = [&](int m, int k) -> double& { return s[size_t(m) * 8 + k]; };
  for(int m = 1; m <= n; m++) S(m,6) = f[m];
  S(1,5) = S(2,6) - S(1,6);
  S(2,5) = 0.5 * (S(3,6) - S(1,6));
  S(n-1,5) = 0.5 * (S(n,6) - S(n-2,6));
  S(n,5) = S(n,6) - S(n-1,6);
  for(int m = 3; m <= n - 2; m++) S(m,5) = ((S(m-2,6) - S(m+2,6)) + 8.0 * (S(m+1,6) - S(m-1,6))) / 12.0;
  for(int m = 1; m <= n - 1; m++)
  {
    S(m,4) = 3.0 * (S(m+1,6) - S(m,6)) - 2.0 * S(m,5) - S(m+1,5);
    S(m,3) = S(m,5) + S(m+1,5) - 2.0 * (S(m+1,6) - S(m,6));
  }
  S(n,4) = 0.0; S(n,3) = 0.0;
  for(int m = 1; m <= n; m++) { S(m,2) = S(m,5) / delta; S(m,1) = 2.0 * S(m,4) / delta; S(m,0) = 3.0 * S(m,3) / delta; }
}

} // namespace xsb

using namespace xsb;
#define COMMA ,

extern "C" {

int xsb_eam_johnson_force(xsb_ctx* ctx, const double* params19, double rcut, int phases, int flags)
{
  return xsb_eam_analytic_force(ctx, XSB_EAM_JOHNSON, params19, 19, rcut, phases, flags);
}

int xsb_eam_analytic_force(xsb_ctx* ctx, int model, const double* params, int nparams, double rcut, int phases, int flags)
{
  XSB_ENTER(ctx);
  const int want = model == XSB_EAM_JOHNSON ? 19 : (model == XSB_EAM_SUTTON_CHEN ? 5 : (model == XSB_EAM_VNIITF ? 13 : -1));
  XSB_REQUIRE(ctx, want > 0, XSB_ERR_UNSUPPORTED, "single-species EAM model not implemented (johnson, sutton_chen, vniitf)");
  XSB_REQUIRE(ctx, params != nullptr && nparams == want && rcut > 0.0, XSB_ERR_INVALID, "eam: null parameters, wrong parameter count (johnson 19, sutton_chen 5, vniitf 13) or rcut <= 0");
  XSB_REQUIRE(ctx, ctx->nbh_built, XSB_ERR_STATE, "chunk_neighbors must be built before a force operator");
  XSB_REQUIRE(ctx, rcut <= ctx->nbh_dist, XSB_ERR_INVALID, "rcut exceeds the neighbour-list distance nbh_dist_lab");
  XSB_CUDA(ctx, cudaSetDevice(ctx->device));
  double params19[20] = {};                  // the per-pair cache key: parameters zero-padded + model
  std::memcpy(params19, params, size_t(want) * sizeof(double)); params19[19] = double(model);
  JohnsonP p; std::memcpy(&p, params19, 19 * sizeof(double)); p.model = model; p.pad_ = 0;
  const bool virial = flags & XSB_FLAG_VIRIAL;
  const bool mixed = (flags & XSB_FLAG_MIXED) && ctx->tile_ok;       // FP32 rho(r), phi(r) on the tile path (tolerance 1e-5)
  const int pw_kind = mixed ? (2 | 16) : 2;                          // a cached rho'(r) computed in FP32 serves FP32 force passes only
  if( virial ) { int rc = xsb_internal_ensure_virial(ctx); if( rc ) return rc; }
  const XForm X = make_xform(ctx->grid); const bool xf = !ctx->grid.xform_is_identity;
  constexpr int TPA = 8; const int block = 256; const double rc2 = rcut * rcut;
  ParticleView P{ ctx->f64[XSB_F_RX].p, ctx->f64[XSB_F_RY].p, ctx->f64[XSB_F_RZ].p, ctx->type.p, ctx->nbh_off.p, ctx->nbh_idx.p, nullptr, 0 };
  double* emb = ctx->f64[XSB_F_RHO_DEMB].p;
  if( ctx->tile_ok )
  {
    if( (phases & 1) && ctx->n )
    {
      XSB_CUDA(ctx, cudaMemsetAsync(emb, 0, ctx->n * sizeof(double), ctx->stream));
      const bool pwo = !ctx->pair_cache_off;
      ctx->sub_pw_kind = 0;
      ctx->prof_begin(XSB_PROF_EAM_RHO);
      int rc;
      if( mixed )    { JohnsonEmbTileOp<true, float> op{ rc2, p, ctx->f64[XSB_F_EP].p, emb, johnson_f32(p) }; rc = launch_tile_pass<32, 1024>(ctx, (phases & 2) != 0, op, nullptr, LIST_FULL_WRITE_SUB); }
      else if( pwo ) { JohnsonEmbTileOp<true>  op{ rc2, p, ctx->f64[XSB_F_EP].p, emb, p }; rc = launch_tile_pass<32, 1024>(ctx, (phases & 2) != 0, op, nullptr, LIST_FULL_WRITE_SUB); }
      else           { JohnsonEmbTileOp<false> op{ rc2, p, ctx->f64[XSB_F_EP].p, emb, p }; rc = launch_tile_pass<32, 1024>(ctx, (phases & 2) != 0, op, nullptr, LIST_FULL_WRITE_SUB); }
      ctx->prof_end(XSB_PROF_EAM_RHO);
      if( rc ) return rc;
      ctx->sub_epoch = ctx->pos_epoch; ctx->sub_rcut = rcut; ctx->sub_ghost = (phases & 2) != 0;
      if( pwo || mixed ) { ctx->sub_pw_kind = pw_kind; std::memcpy(ctx->sub_pw_johnson, params19, sizeof(ctx->sub_pw_johnson)); }
    }
    if( (phases & 4) && ctx->n_own )
    {
      double *fx = ctx->f64[XSB_F_FX].p, *fy = ctx->f64[XSB_F_FY].p, *fz = ctx->f64[XSB_F_FZ].p, *ep = ctx->f64[XSB_F_EP].p;
      int rc;
      const int lmode = ctx->sub_valid(rcut, false) ? LIST_SUB : LIST_FULL;
      // the cached rho'(r) is only good for the parameter set that produced it
      const bool pwi = lmode == LIST_SUB && ctx->sub_pw_kind == pw_kind && std::memcmp(ctx->sub_pw_johnson, params19, sizeof(ctx->sub_pw_johnson)) == 0;
      ctx->prof_begin(XSB_PROF_EAM_FORCE);
      if( mixed )
      {
        const JohnsonPT<float> q = johnson_f32(p);
        if( virial ) { if( pwi ) { JohnsonForceTileOp<true, true, float>   op{ rc2, q, fx, fy, fz, ep, ctx->f64[XSB_F_VIRIAL].p }; rc = launch_tile_pass<8, 512>(ctx, false, op, emb, lmode); }
                       else      { JohnsonForceTileOp<true, false, float>  op{ rc2, q, fx, fy, fz, ep, ctx->f64[XSB_F_VIRIAL].p }; rc = launch_tile_pass<8, 512>(ctx, false, op, emb, lmode); } }
        else         { if( pwi ) { JohnsonForceTileOp<false, true, float>  op{ rc2, q, fx, fy, fz, ep, nullptr }; rc = launch_tile_pass<16, 1024>(ctx, false, op, emb, lmode); }
                       else      { JohnsonForceTileOp<false, false, float> op{ rc2, q, fx, fy, fz, ep, nullptr }; rc = launch_tile_pass<16, 1024>(ctx, false, op, emb, lmode); } }
      }
      else
      if( virial ) { if( pwi ) { JohnsonForceTileOp<true, true>   op{ rc2, p, fx, fy, fz, ep, ctx->f64[XSB_F_VIRIAL].p }; rc = launch_tile_pass<8, 512>(ctx, false, op, emb, lmode); }
                     else      { JohnsonForceTileOp<true, false>  op{ rc2, p, fx, fy, fz, ep, ctx->f64[XSB_F_VIRIAL].p }; rc = launch_tile_pass<8, 512>(ctx, false, op, emb, lmode); } }
      else         { if( pwi ) { JohnsonForceTileOp<false, true>  op{ rc2, p, fx, fy, fz, ep, nullptr }; rc = launch_tile_pass<16, 1024>(ctx, false, op, emb, lmode); }
                     else      { JohnsonForceTileOp<false, false> op{ rc2, p, fx, fy, fz, ep, nullptr }; rc = launch_tile_pass<16, 1024>(ctx, false, op, emb, lmode); } }
      ctx->prof_end(XSB_PROF_EAM_FORCE);
      if( rc ) return rc;
    }
    return XSB_OK;
  }
  if( (phases & 1) && ctx->n )
  {
    const bool ghost = phases & 2;
    P.atoms = ghost ? nullptr : ctx->own_atoms.p; P.n_atoms = unsigned(ghost ? ctx->n : ctx->n_own);
    XSB_CUDA(ctx, cudaMemsetAsync(emb, 0, ctx->n * sizeof(double), ctx->stream));
    if( P.n_atoms )
    {
      const unsigned grid = groups_grid<TPA>(P.n_atoms, block);
      if( xf ) johnson_emb_kernel<TPA, true ><<<grid, block, 0, ctx->stream>>>(P, X, p, rc2, ctx->f64[XSB_F_EP].p, emb);
      else     johnson_emb_kernel<TPA, false><<<grid, block, 0, ctx->stream>>>(P, X, p, rc2, ctx->f64[XSB_F_EP].p, emb);
      XSB_LAUNCH_CHECK(ctx);
    }
  }
  if( (phases & 4) && ctx->n_own )
  {
    P.atoms = ctx->own_atoms.p; P.n_atoms = unsigned(ctx->n_own);
    const unsigned grid = groups_grid<TPA>(P.n_atoms, block);
    double *fx = ctx->f64[XSB_F_FX].p, *fy = ctx->f64[XSB_F_FY].p, *fz = ctx->f64[XSB_F_FZ].p, *ep = ctx->f64[XSB_F_EP].p;
    double* vir = virial ? ctx->f64[XSB_F_VIRIAL].p : nullptr;
    if( xf ) { if( virial ) johnson_force_kernel<TPA, true, true><<<grid, block, 0, ctx->stream>>>(P, X, p, rc2, emb, fx, fy, fz, ep, vir);
               else         johnson_force_kernel<TPA, true, false><<<grid, block, 0, ctx->stream>>>(P, X, p, rc2, emb, fx, fy, fz, ep, vir); }
    else     { if( virial ) johnson_force_kernel<TPA, false, true><<<grid, block, 0, ctx->stream>>>(P, X, p, rc2, emb, fx, fy, fz, ep, vir);
               else         johnson_force_kernel<TPA, false, false><<<grid, block, 0, ctx->stream>>>(P, X, p, rc2, emb, fx, fy, fz, ep, vir); }
    XSB_LAUNCH_CHECK(ctx);
  }
  return XSB_OK;
}

int xsb_eam_alloy_read(const char* path, xsb_eam_alloy_tables* out, char* names, size_t names_len)
{
  if( !path || !out ) return XSB_ERR_INVALID;
  std::memset(out, 0, sizeof(*out));
  std::ifstream file(path);
  if( !file ) return XSB_ERR_IO;
  for(int i = 0; i < 3; i++) file.ignore(std::numeric_limits<std::streamsize>::max(), '\n');   // 3 comment lines
  size_t nel = 0; file >> nel;
  if( !file || nel == 0 || nel > 7 ) return XSB_ERR_IO;                                           // MAX_ELEMENTS = 7 (eam_alloy.h:137)
  std::string all;
  for(size_t i = 0; i < nel; i++) { std::string s; file >> s; all += (i ? " " : "") + s; }
  size_t nrho = 0, nr = 0; double drho = 0, dr = 0, rc = 0;
  file >> nrho >> drho >> nr >> dr >> rc;
  if( !file || nrho < 5 || nr < 5 || drho <= 0 || dr <= 0 ) return XSB_ERR_IO;
  const size_t nz = nel * (nel + 1) / 2;
  std::vector<double> f(std::max(nrho, nr) + 1);
  double* frho = new double[nel * (nrho + 1) * 8](); double* rhor = new double[nel * (nr + 1) * 8](); double* z2r = new double[nz * (nr + 1) * 8]();
  bool ok = true;
  for(size_t i = 0; i < nel && ok; i++)
  {
    int z; double mass, a0; std::string st;
    file >> z >> mass >> a0 >> st;
    f[0] = 0.0; for(size_t k = 0; k < nrho; k++) file >> f[k + 1];
    if( !file ) { ok = false; break; }
    interpolate(int(nrho), drho, f.data(), frho + i * (nrho + 1) * 8);
    for(size_t k = 0; k < nr; k++) file >> f[k + 1];
    if( !file ) { ok = false; break; }
    interpolate(int(nr), dr, f.data(), rhor + i * (nr + 1) * 8);
  }
  for(size_t i = 0; i < nz && ok; i++)
  {
    for(size_t k = 0; k < nr; k++) file >> f[k + 1];
    if( !file ) { ok = false; break; }
    interpolate(int(nr), dr, f.data(), z2r + i * (nr + 1) * 8);
  }
  if( !ok ) { delete[] frho; delete[] rhor; delete[] z2r; return XSB_ERR_IO; }
  out->nelements = int(nel); out->nr = int(nr); out->nrho = int(nrho);
  out->rdr = 1.0 / dr; out->rdrho = 1.0 / drho; out->rc = rc; out->rhomax = (nrho - 1) * drho;
  const double ev = 1.602176634e-19 / (1.66053906660e-27 * 1.0e4);   // 1 eV in (Da, ang, ps) units
  out->conversion_z2r = ev; out->conversion_frho = ev;
  out->frho = frho; out->rhor = rhor; out->z2r = z2r;
  if( names && names_len ) { std::strncpy(names, all.c_str(), names_len - 1); names[names_len - 1] = 0; }
  return XSB_OK;
}

void xsb_eam_alloy_free(xsb_eam_alloy_tables* t)
{
  if( !t ) return;
  delete[] t->frho; delete[] t->rhor; delete[] t->z2r;
  t->frho = t->rhor = t->z2r = nullptr;
}

int xsb_eam_alloy_set(xsb_ctx* ctx, const xsb_eam_alloy_tables* t)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, t && t->frho && t->rhor && t->z2r, XSB_ERR_INVALID, "null eam tables");
  XSB_REQUIRE(ctx, t->nelements >= 1 && t->nelements <= 7 && t->nr >= 5 && t->nrho >= 5, XSB_ERR_INVALID, "bad eam table sizes");
  XSB_CUDA(ctx, cudaSetDevice(ctx->device));
  EamAlloyDev& E = ctx->eam;
  const size_t nel = t->nelements, nz = nel * (nel + 1) / 2;
  const size_t nf = nel * (t->nrho + 1) * 8, nrr = nel * (t->nr + 1) * 8, nzz = nz * (t->nr + 1) * 8;
  XSB_CUDA(ctx, E.frho.reserve(nf));
  XSB_CUDA(ctx, E.rtab.reserve(nrr + nzz));
  XSB_CUDA(ctx, cudaMemcpyAsync(E.frho.p, t->frho, nf * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  XSB_CUDA(ctx, cudaMemcpyAsync(E.rtab.p, t->rhor, nrr * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  XSB_CUDA(ctx, cudaMemcpyAsync(E.rtab.p + nrr, t->z2r, nzz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  {
    // Hermite-knot view {f = c6, c5} of the r-tables for the tile kernels
    std::vector<double> fc(2 * (nrr + nzz) / 8);
    for(size_t row = 0; row < nrr / 8; row++) { fc[2*row] = t->rhor[8*row + 6]; fc[2*row + 1] = t->rhor[8*row + 5]; }
    for(size_t row = 0; row < nzz / 8; row++) { fc[2*(nrr/8 + row)] = t->z2r[8*row + 6]; fc[2*(nrr/8 + row) + 1] = t->z2r[8*row + 5]; }
    XSB_CUDA(ctx, E.fc.reserve(fc.size() + 2));
    XSB_CUDA(ctx, cudaMemcpyAsync(E.fc.p, fc.data(), fc.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    // FP32 rows of the mixed-precision passes: the cubic of interval m of table t as {c6, c5, c4, c3} (interpolate()'s
    // coefficients, rounded once from FP64)
    const size_t nrow = fc.size() / 2;
    std::vector<float> f32(4 * nrow);
    for(size_t row = 0; row < nrr / 8; row++) { const double* c = t->rhor + 8 * row; f32[4*row] = float(c[6]); f32[4*row + 1] = float(c[5]); f32[4*row + 2] = float(c[4]); f32[4*row + 3] = float(c[3]); }
    for(size_t row = 0; row < nzz / 8; row++) { const double* c = t->z2r + 8 * row; const size_t o = 4 * (nrr / 8 + row); f32[o] = float(c[6]); f32[o + 1] = float(c[5]); f32[o + 2] = float(c[4]); f32[o + 3] = float(c[3]); }
    XSB_CUDA(ctx, E.fc32.reserve(f32.size() + 4));
    XSB_CUDA(ctx, cudaMemcpyAsync(E.fc32.p, f32.data(), f32.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  E.nelements = t->nelements; E.nr = t->nr; E.nrho = t->nrho; E.rdr = t->rdr; E.rdrho = t->rdrho; E.rc = t->rc; E.rhomax = t->rhomax;
  E.conv_z2r = t->conversion_z2r; E.conv_frho = t->conversion_frho; E.set = true;
  ctx->sub_pw_kind = 0;     // a cached rho'(r) belongs to the previous tables
  ctx->tables_id++;
  return XSB_OK;
}

// Inner skin of the in-range sub-list of eam_alloy_force (angstrom, 0 = off; default from env XSB_INNER_SKIN): see SubCtl.
int xsb_eam_inner_skin(xsb_ctx* ctx, double skin)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, skin >= 0.0 && skin < 1.0e3, XSB_ERR_INVALID, "inner skin must be >= 0");
  if( skin != ctx->inner_skin ) { ctx->inner_skin = skin; ctx->sub_foreign = 0; }
  return XSB_OK;
}

// how often the first EAM pass re-filtered the neighbour list / only re-evaluated its sub-list since the context was created
int xsb_eam_sublist_stats(xsb_ctx* ctx, uint64_t* refiltered, uint64_t* reused)
{
  XSB_ENTER(ctx);
  SubCtl h{};
  if( ctx->sub_ctl.p )
  {
    XSB_CUDA(ctx, cudaMemcpyAsync(&h, ctx->sub_ctl.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  if( refiltered ) *refiltered = h.builds;
  if( reused ) *reused = h.reuses;
  return XSB_OK;
}

} // extern "C"

// third phase of eam_alloy_force (eam_force_op_multimat.h:243-343), optionally with the pair operator chained behind it
int xsb_internal_eam_force_phase(xsb_ctx* ctx, double rcut, int phases, int flags, const xsb::LJMulti* chain)
{
  const EamAlloyDev& E = ctx->eam;
  EamAlloyView T{ E.frho.p, E.rtab.p, E.rtab.p + size_t(E.nelements) * (E.nr + 1) * 8, E.nr, E.nrho, E.rdr, E.rdrho, E.rhomax, E.conv_z2r, E.conv_frho };
  const bool eflag = phases & XSB_EAM_EFLAG;
  const bool virial = eflag && (flags & XSB_FLAG_VIRIAL);
  const bool mixed = (flags & XSB_FLAG_MIXED) && ctx->tile_ok;
  const XForm X = make_xform(ctx->grid); const bool xf = !ctx->grid.xform_is_identity;
  constexpr int TPA = 8; const int block = 256; const double rc2 = rcut * rcut;
  ParticleView P{ ctx->f64[XSB_F_RX].p, ctx->f64[XSB_F_RY].p, ctx->f64[XSB_F_RZ].p, ctx->type.p, ctx->nbh_off.p, ctx->nbh_idx.p, nullptr, 0 };
  double* emb = ctx->f64[XSB_F_RHO_DEMB].p;
  const bool tile = ctx->tile_ok, multi = E.nelements > 1;
  if( tile && (phases & XSB_EAM_FORCE) && ctx->n_own )
  {
    double *fx = ctx->f64[XSB_F_FX].p, *fy = ctx->f64[XSB_F_FY].p, *fz = ctx->f64[XSB_F_FZ].p, *ep = ctx->f64[XSB_F_EP].p;
    double* vir = virial ? ctx->f64[XSB_F_VIRIAL].p : nullptr;
    const int ntab = E.nelements + E.nelements * (E.nelements + 1) / 2;
    int rc;
    ctx->prof_begin(XSB_PROF_EAM_FORCE);
    const int lmode = ctx->sub_valid(rcut, false) ? LIST_SUB : LIST_FULL;
    const bool pwi = lmode == LIST_SUB && ctx->sub_pw_kind == ((multi ? 3 : 1) | (mixed ? 16 : 0)) && !(multi && ctx->type_external);
    const int npair = E.nelements * (E.nelements + 1) / 2;
    XSB_REQUIRE(ctx, pwi || !chain, XSB_ERR_STATE, "chained pair operator without the per-pair cache of the rho pass");
#   define XSB_EAM_TILE(MU, EF, VIR, TPA_, NT_) { \
      if( pwi && chain ) { EamForceTileOp<MU, EF, VIR, true, true> op{ rc2, make_fc_view<true, MU>(ctx, E.nelements, npair, 0), E.nelements, E.conv_z2r, fx, fy, fz, ep, vir, *chain }; rc = launch_tile_pass<TPA_, NT_>(ctx, false, op, emb, lmode); } \
      else if( pwi ) { EamForceTileOp<MU, EF, VIR, true>  op{ rc2, make_fc_view<true, MU>(ctx, E.nelements, npair, 0), E.nelements, E.conv_z2r, fx, fy, fz, ep, vir }; rc = launch_tile_pass<TPA_, NT_>(ctx, false, op, emb, lmode); } \
      else      { EamForceTileOp<MU, EF, VIR, false> op{ rc2, make_fc_view<true, MU>(ctx, 0, ntab, 0), E.nelements, E.conv_z2r, fx, fy, fz, ep, vir }; rc = launch_tile_pass<TPA_, NT_>(ctx, false, op, emb, lmode); } }
#   define XSB_EAM_TILE32(MU, EF, VIR, TPA_, NT_) { \
      if( pwi && chain ) { EamForceTileOp32<MU, EF, VIR, true, true> op{ rc2, make_fc_view32<true, MU>(ctx, E.nelements, npair, 0), E.nelements, float(E.conv_z2r), fx, fy, fz, ep, vir, *chain }; rc = launch_tile_pass<TPA_, NT_>(ctx, false, op, emb, lmode); } \
      else if( pwi ) { EamForceTileOp32<MU, EF, VIR, true>  op{ rc2, make_fc_view32<true, MU>(ctx, E.nelements, npair, 0), E.nelements, float(E.conv_z2r), fx, fy, fz, ep, vir }; rc = launch_tile_pass<TPA_, NT_>(ctx, false, op, emb, lmode); } \
      else      { EamForceTileOp32<MU, EF, VIR, false> op{ rc2, make_fc_view32<true, MU>(ctx, 0, ntab, 0), E.nelements, float(E.conv_z2r), fx, fy, fz, ep, vir }; rc = launch_tile_pass<TPA_, NT_>(ctx, false, op, emb, lmode); } }
    if( mixed )
    {
      if( multi ) { if( virial ) XSB_EAM_TILE32(true, true, true, 8, 512) else if( eflag ) XSB_EAM_TILE32(true, true, false, 16, 1024) else XSB_EAM_TILE32(true, false, false, 16, 1024) }
      else        { if( virial ) XSB_EAM_TILE32(false, true, true, 8, 512) else if( eflag ) XSB_EAM_TILE32(false, true, false, 16, 1024) else XSB_EAM_TILE32(false, false, false, 16, 1024) }
    }
    else
    if( multi ) { if( virial ) XSB_EAM_TILE(true, true, true, 8, 512) else if( eflag ) XSB_EAM_TILE(true, true, false, 16, 1024) else XSB_EAM_TILE(true, false, false, 16, 1024) }
    else        { if( virial ) XSB_EAM_TILE(false, true, true, 8, 512) else if( eflag ) XSB_EAM_TILE(false, true, false, 16, 1024) else if( ctx->exp_tpa == 8 ) XSB_EAM_TILE(false, false, false, 8, 1024) else XSB_EAM_TILE(false, false, false, 16, 1024) }
#   undef XSB_EAM_TILE32
#   undef XSB_EAM_TILE
    ctx->prof_end(XSB_PROF_EAM_FORCE);
    if( rc ) return rc;
  }
  if( !tile && (phases & XSB_EAM_FORCE) && ctx->n_own )
  {
    P.atoms = ctx->own_atoms.p; P.n_atoms = unsigned(ctx->n_own);
    const unsigned grid = groups_grid<TPA>(P.n_atoms, block);
    double *fx = ctx->f64[XSB_F_FX].p, *fy = ctx->f64[XSB_F_FY].p, *fz = ctx->f64[XSB_F_FZ].p, *ep = ctx->f64[XSB_F_EP].p;
    double* vir = virial ? ctx->f64[XSB_F_VIRIAL].p : nullptr;
#   define XSB_EAM_GO(XF, EF, VIR) eam_alloy_force_kernel<TPA, XF, EF, VIR><<<grid, block, 0, ctx->stream>>>(P, X, T, rc2, emb, fx, fy, fz, ep, vir)
    ctx->prof_begin(XSB_PROF_EAM_FORCE);
    if( xf ) { if( virial ) XSB_EAM_GO(true, true, true); else if( eflag ) XSB_EAM_GO(true, true, false); else XSB_EAM_GO(true, false, false); }
    else     { if( virial ) XSB_EAM_GO(false, true, true); else if( eflag ) XSB_EAM_GO(false, true, false); else XSB_EAM_GO(false, false, false); }
    ctx->prof_end(XSB_PROF_EAM_FORCE);
#   undef XSB_EAM_GO
    XSB_LAUNCH_CHECK(ctx);
  }
  return XSB_OK;
}

int xsb_internal_flush_pending(xsb_ctx* ctx)
{
  const xsb_ctx::PendingEamForce pe = ctx->pending_eam;
  ctx->pending_eam.active = false;
  return xsb_internal_eam_force_phase(ctx, pe.rcut, pe.phases, pe.flags, nullptr);
}

extern "C" {

int xsb_eam_alloy_force(xsb_ctx* ctx, double rcut, int phases, int flags)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, ctx->eam.set, XSB_ERR_STATE, "xsb_eam_alloy_set must be called first");
  XSB_REQUIRE(ctx, rcut > 0.0, XSB_ERR_INVALID, "rcut must be > 0");
  XSB_REQUIRE(ctx, ctx->nbh_built, XSB_ERR_STATE, "chunk_neighbors must be built before a force operator");
  XSB_REQUIRE(ctx, rcut <= ctx->nbh_dist, XSB_ERR_INVALID, "rcut exceeds the neighbour-list distance nbh_dist_lab");
  XSB_CUDA(ctx, cudaSetDevice(ctx->device));
  const EamAlloyDev& E = ctx->eam;
  EamAlloyView T{ E.frho.p, E.rtab.p, E.rtab.p + size_t(E.nelements) * (E.nr + 1) * 8, E.nr, E.nrho, E.rdr, E.rdrho, E.rhomax, E.conv_z2r, E.conv_frho };
  const bool eflag = phases & XSB_EAM_EFLAG, ghost = phases & XSB_EAM_GHOST;
  const bool virial = eflag && (flags & XSB_FLAG_VIRIAL);
  const bool mixed = (flags & XSB_FLAG_MIXED) && ctx->tile_ok;      // FP32 spline + pair math on the tile path (tolerance 1e-5)
  if( virial ) { int rc = xsb_internal_ensure_virial(ctx); if( rc ) return rc; }
  const XForm X = make_xform(ctx->grid); const bool xf = !ctx->grid.xform_is_identity;
  constexpr int TPA = 8; const int block = 256; const double rc2 = rcut * rcut;
  ParticleView P{ ctx->f64[XSB_F_RX].p, ctx->f64[XSB_F_RY].p, ctx->f64[XSB_F_RZ].p, ctx->type.p, ctx->nbh_off.p, ctx->nbh_idx.p, nullptr, 0 };
  double* emb = ctx->f64[XSB_F_RHO_DEMB].p;
  const unsigned* sel = ghost ? nullptr : ctx->own_atoms.p; const unsigned nsel = unsigned(ghost ? ctx->n : ctx->n_own);
  const bool tile = ctx->tile_ok, multi = E.nelements > 1;
  if( tile && (phases & XSB_EAM_RHO) && ctx->n )
  {
    XSB_CUDA(ctx, cudaMemsetAsync(emb, 0, ctx->n * sizeof(double), ctx->stream));
    int rc;
    ctx->prof_begin(XSB_PROF_EAM_RHO);
    // one warp per atom + compaction queue; the in-range sub-list it leaves behind serves the force pass of this step
    const size_t qb = tile_queue_bytes<1024>();
    // the force pass of this step reuses rho'(r) of every in-range pair: cache it only when that pass can follow
    const bool pwo = !ctx->pair_cache_off;
    ctx->sub_pw_kind = 0;
    // Inner skin (XSB_INNER_SKIN > 0, needs the per-pair cache): the sub-list keeps the pairs up to rcut + skin; while no atom
    // has moved further than skin / 2 since it was filtered (SubCtl::acc, fed by the integrator kernels) the pass only
    // re-evaluates its entries -- dense walk, no compaction, 147 instead of 196 candidates per atom at C2 -- and the
    // device itself picks the launch that runs (the other one returns at once): no host read-back.
    double skin = (pwo || mixed) ? std::min(ctx->inner_skin, ctx->nbh_dist - rcut) : 0.0;
    if( !(skin > 0.0) ) skin = 0.0;
    const double lrc2 = (rcut + skin) * (rcut + skin);
    const int* mode = nullptr;
    bool two = false;
    if( skin > 0.0 )
    {
      if( !ctx->sub_ctl.p ) { XSB_CUDA(ctx, ctx->sub_ctl.reserve(1)); XSB_CUDA(ctx, cudaMemsetAsync(ctx->sub_ctl.p, 0, sizeof(SubCtl), ctx->stream)); ctx->sub_foreign = 0; }
      const int kind_now = (multi ? 3 : 1) | (mixed ? 16 : 0);
      const bool reusable = ctx->sub_foreign == ctx->foreign_epoch && ctx->sub_rcut == rcut && ctx->sub_list_rc == rcut + skin && ctx->sub_ghost == ghost &&
                            ctx->sub_tables_id == ctx->tables_id && ctx->sub_pw_kind_built == kind_now && !ctx->pos_external && !(multi && ctx->type_external);
      sub_decide_kernel<<<1, 1, 0, ctx->stream>>>(ctx->sub_ctl.p, reusable ? 0 : 1, 0.5 * skin);
      XSB_LAUNCH_CHECK(ctx);
      mode = &ctx->sub_ctl.p->mode; two = true;
    }
#   define XSB_RHO_PASS(OP, ...) { OP op{ __VA_ARGS__ }; rc = launch_tile_pass<32, 1024>(ctx, ghost, op, nullptr, LIST_FULL_WRITE_SUB, lrc2, mode, 0); \
                                   if( rc == XSB_OK && two ) rc = launch_tile_pass<16, 1024>(ctx, ghost, op, nullptr, LIST_SUB_REWRITE, lrc2, mode, 1); }
    if( mixed )
    {
      if( multi ) XSB_RHO_PASS(EamRhoTileOp32<true COMMA true>, rc2, make_fc_view32<false, true >(ctx, 0, E.nelements, qb), emb)
      else        XSB_RHO_PASS(EamRhoTileOp32<false COMMA true>, rc2, make_fc_view32<false, false>(ctx, 0, E.nelements, qb), emb)
    }
    else
    if( multi ) { if( pwo ) XSB_RHO_PASS(EamRhoTileOp<true COMMA true>, rc2, make_fc_view<false, true >(ctx, 0, E.nelements, qb), emb)
                  else      { EamRhoTileOp<true, false>  op{ rc2, make_fc_view<false, true >(ctx, 0, E.nelements, qb), emb }; rc = launch_tile_pass<32, 1024>(ctx, ghost, op, nullptr, LIST_FULL_WRITE_SUB); } }
    else        { if( pwo ) XSB_RHO_PASS(EamRhoTileOp<false COMMA true>, rc2, make_fc_view<false, false>(ctx, 0, E.nelements, qb), emb)
                  else      { EamRhoTileOp<false, false> op{ rc2, make_fc_view<false, false>(ctx, 0, E.nelements, qb), emb }; rc = launch_tile_pass<32, 1024>(ctx, ghost, op, nullptr, LIST_FULL_WRITE_SUB); } }
#   undef XSB_RHO_PASS
    if( rc == XSB_OK && skin > 0.0 ) { ctx->sub_foreign = ctx->foreign_epoch; ctx->sub_list_rc = rcut + skin; ctx->sub_tables_id = ctx->tables_id; ctx->sub_pw_kind_built = (multi ? 3 : 1) | (mixed ? 16 : 0); }
    else ctx->sub_foreign = 0;
    ctx->prof_end(XSB_PROF_EAM_RHO);
    if( rc ) return rc;
    ctx->sub_epoch = ctx->pos_epoch; ctx->sub_rcut = rcut; ctx->sub_ghost = ghost;
    if( pwo || mixed ) ctx->sub_pw_kind = (multi ? 3 : 1) | (mixed ? 16 : 0);      // 3: two values per pair (rhojp, rhoip); +16: computed in FP32
  }
  if( !tile && (phases & XSB_EAM_RHO) && ctx->n )
  {
    XSB_CUDA(ctx, cudaMemsetAsync(emb, 0, ctx->n * sizeof(double), ctx->stream));
    P.atoms = sel; P.n_atoms = nsel;
    if( nsel )
    {
      const unsigned grid = groups_grid<TPA>(nsel, block);
      ctx->prof_begin(XSB_PROF_EAM_RHO);
      if( xf ) eam_alloy_rho_kernel<TPA, true ><<<grid, block, 0, ctx->stream>>>(P, X, T, rc2, emb);
      else     eam_alloy_rho_kernel<TPA, false><<<grid, block, 0, ctx->stream>>>(P, X, T, rc2, emb);
      ctx->prof_end(XSB_PROF_EAM_RHO);
      XSB_LAUNCH_CHECK(ctx);
    }
  }
  if( (phases & XSB_EAM_RHO2EMB) && nsel )
  {
    ctx->prof_begin(XSB_PROF_EAM_RHO2EMB);
    eam_alloy_rho2emb_kernel<<<(nsel + 255) / 256, 256, 0, ctx->stream>>>(sel, nsel, T, ctx->type.p, emb, eflag ? ctx->f64[XSB_F_EP].p : nullptr);
    ctx->prof_end(XSB_PROF_EAM_RHO2EMB);
    XSB_LAUNCH_CHECK(ctx);
  }
  if( (phases & XSB_EAM_FORCE) && ctx->n_own )
  {
    // the force phase waits for the next entry point: a pair operator chained behind this one joins its pass (xsb_ctx::pending_eam)
    const bool pwi_now = tile && ctx->sub_valid(rcut, false) && ctx->sub_pw_kind == ((multi ? 3 : 1) | (mixed ? 16 : 0)) && !(multi && ctx->type_external);
    if( ctx->chain_fusion && pwi_now ) { ctx->pending_eam.active = true; ctx->pending_eam.rcut = rcut; ctx->pending_eam.phases = phases; ctx->pending_eam.flags = flags; }
    else return xsb_internal_eam_force_phase(ctx, rcut, phases, flags, nullptr);
  }
  return XSB_OK;
}

} // extern "C"
