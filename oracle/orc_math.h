// TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into or called by the product path.
//
// orc_math.h : plain C++ restatement of the per-pair / per-atom arithmetic of the exaStamp
// short-range force operators.  Every function cites the reference file:line it follows
// (paths relative to /root/reference).  The restatement is pinned by tests/test_oracle_math.py
// against oracle/_ref/libxsref.so, which compiles the reference's own unmodified headers.
#pragma once
#include <cmath>
#include <cstring>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <string>
#include <vector>

namespace orc
{

// ---------------------------------------------------------------------------------------------
// internal unit system: angstrom, Dalton, picosecond, e, K (include/exaStamp/unit_system.h:28-36)
// => energy unit Da.ang^2/ps^2.  onika's constant table is external (not vendored); CODATA-2018
// values are used here and in the product library (one place each), so both sides agree.
// ---------------------------------------------------------------------------------------------
static constexpr double ELEMENTARY_CHARGE_C = 1.602176634e-19;
static constexpr double ATOMIC_MASS_KG      = 1.66053906660e-27;
static constexpr double EV_INTERNAL         = ELEMENTARY_CHARGE_C / ( ATOMIC_MASS_KG * 1.0e4 ); // 1 eV
static constexpr double EV_ANG_INTERNAL     = EV_INTERNAL;                                       // 1 eV.ang

// src/potential/pair_potentials/lennard_jones/include/.../lennard_jones.h:40-50
struct LJParams { double epsilon, sigma; };
static inline void lj_compute_energy(const LJParams& p, double r, double& e, double& de)
{
  const double ratio   = p.sigma / r;
  const double ratio2  = ratio * ratio;
  const double ratio6  = ratio2 * ratio2 * ratio2;
  const double ratio12 = ratio6 * ratio6;
  e  = 4. * p.epsilon * (ratio12 - ratio6);
  de = ( -24. * p.epsilon * (2. * ratio12 - ratio6) ) / r;
}

// src/potential/pair_potential_template/pair_potential_impl.hxx:488-498 (energy_cutoff)
static inline double lj_energy_cutoff(const LJParams& p, double rcut)
{
  double e = 0.0, de = 0.0;
  if( rcut > 0.0 ) lj_compute_energy(p, rcut, e, de);
  return e;
}

// ---- further pair potentials behind the same <pot>_compute_force / <pot>_multi_force template ----------------------
// pot ids (include/xsb200.h xsb_pair_pot): 0 lj {epsilon, sigma}; 1 zbl {r1, rc, z_a, z_b}; 2 exp6 {A, B, C, D};
// 3 buckingham {A, Rho, C}.

// src/potential/pair_potentials/zbl/potential.h:36-57 (constants), :89-178 (e_zbl, dzbldr, d2zbldr2), :180-303
// (zbl_compute_energy: LAMMPS pair_zbl with the inner/outer switching polynomial; e and de stay 0 for r >= rc).
// Energies come out in eV and are scaled by 1e-4 e / amu = 1 eV in internal units (:298-301); the onika constants
// behind that factor are not in the reference tree (CODATA 2018 values assumed: parity of this factor is unpinned).
struct ZBLParams { double r1, rc; };
namespace zblc { constexpr double pzbl = 0.23, a0 = 0.46850, c1 = 0.02817, c2 = 0.28022, c3 = 0.50986, c4 = 0.18175,
                                  d1 = 0.20162, d2 = 0.40290, d3 = 0.94229, d4 = 3.19980; }
static inline double zbl_e(double r, double d1a, double d2a, double d3a, double d4a, double zze)
{
  const double rinv = 1.0 / r;
  double sum = zblc::c1 * exp(-d1a * r);
  sum += zblc::c2 * exp(-d2a * r);
  sum += zblc::c3 * exp(-d3a * r);
  sum += zblc::c4 * exp(-d4a * r);
  return zze * sum * rinv;
}
static inline double zbl_dedr(double r, double d1a, double d2a, double d3a, double d4a, double zze)
{
  const double rinv = 1.0 / r;
  const double e1 = exp(-d1a * r), e2 = exp(-d2a * r), e3 = exp(-d3a * r), e4 = exp(-d4a * r);
  double sum = zblc::c1 * e1; sum += zblc::c2 * e2; sum += zblc::c3 * e3; sum += zblc::c4 * e4;
  double sum_p = -zblc::c1 * d1a * e1; sum_p -= zblc::c2 * d2a * e2; sum_p -= zblc::c3 * d3a * e3; sum_p -= zblc::c4 * d4a * e4;
  return zze * (sum_p - sum * rinv) * rinv;
}
static inline double zbl_d2edr2(double r, double d1a, double d2a, double d3a, double d4a, double zze)
{
  const double rinv = 1.0 / r;
  const double e1 = exp(-d1a * r), e2 = exp(-d2a * r), e3 = exp(-d3a * r), e4 = exp(-d4a * r);
  double sum = zblc::c1 * e1; sum += zblc::c2 * e2; sum += zblc::c3 * e3; sum += zblc::c4 * e4;
  double sum_p = zblc::c1 * e1 * d1a; sum_p += zblc::c2 * e2 * d2a; sum_p += zblc::c3 * e3 * d3a; sum_p += zblc::c4 * e4 * d4a;
  double sum_pp = zblc::c1 * e1 * d1a * d1a; sum_pp += zblc::c2 * e2 * d2a * d2a; sum_pp += zblc::c3 * e3 * d3a * d3a; sum_pp += zblc::c4 * e4 * d4a * d4a;
  return zze * (sum_pp + 2.0 * sum_p * rinv + 2.0 * sum * rinv * rinv) * rinv;
}
// per type-pair constants of zbl_compute_energy (everything that does not depend on r)
struct ZBLPair { double d1a, d2a, d3a, d4a, zze, sw1, sw2, sw3, sw4, sw5, r1, rc; };
static inline ZBLPair zbl_pair(const ZBLParams& p, double z_a, double z_b)
{
  ZBLPair q;
  const double qqr2e = 14.399645, qelectron = 1.0;
  const double ainv = (pow(z_a, zblc::pzbl) + pow(z_b, zblc::pzbl)) / zblc::a0;
  q.d1a = zblc::d1 * ainv; q.d2a = zblc::d2 * ainv; q.d3a = zblc::d3 * ainv; q.d4a = zblc::d4 * ainv;
  q.zze = z_a * z_b * qqr2e * qelectron * qelectron;
  const double tc = p.rc - p.r1;
  const double fc = zbl_e(p.rc, q.d1a, q.d2a, q.d3a, q.d4a, q.zze);
  const double fcp = zbl_dedr(p.rc, q.d1a, q.d2a, q.d3a, q.d4a, q.zze);
  const double fcpp = zbl_d2edr2(p.rc, q.d1a, q.d2a, q.d3a, q.d4a, q.zze);
  const double swa = (-3.0 * fcp + tc * fcpp) / (tc * tc);
  const double swb = (2.0 * fcp - tc * fcpp) / (tc * tc * tc);
  const double swc = -fc + (tc / 2.0) * fcp - (tc * tc / 12.0) * fcpp;
  q.sw1 = swa; q.sw2 = swb; q.sw3 = swa / 3.0; q.sw4 = swb / 4.0; q.sw5 = swc; q.r1 = p.r1; q.rc = p.rc;
  return q;
}
static inline void zbl_compute_energy(const ZBLPair& q, double rij, double& e, double& de)
{
  const double rsq = rij * rij, cut_innersq = q.r1 * q.r1, cut_globalsq = q.rc * q.rc;
  if( rsq < cut_globalsq )
  {
    const double r = sqrt(rsq);
    de = zbl_dedr(r, q.d1a, q.d2a, q.d3a, q.d4a, q.zze);
    if( rsq > cut_innersq ) { const double t = r - q.r1; de += t * t * (q.sw1 + q.sw2 * t); }
    e = zbl_e(r, q.d1a, q.d2a, q.d3a, q.d4a, q.zze);
    e += q.sw5;
    if( rsq > cut_innersq ) { const double t = r - q.r1; e += t * t * t * (q.sw3 + q.sw4 * t); }
  }
  e *= EV_INTERNAL; de *= EV_INTERNAL;
}

// src/potential/pair_potentials/exp6/include/exaStamp/potential/pair_potentials/exp6/exp6.h:66-84
static inline void exp6_compute_energy(const double* p /* A B C D */, double r, double& e, double& de)
{
  const double one_rB = 1 / (r * p[1]);
  const double r2 = r * r, r6 = r2 * r2 * r2;
  const double Cr6 = p[2] / r6;
  const double twelve_rB = 12 * one_rB, twelve_rB2 = twelve_rB * twelve_rB, twelve_rB4 = twelve_rB2 * twelve_rB2;
  const double twelve_rB12 = twelve_rB4 * twelve_rB4 * twelve_rB4;
  const double Dtwelve_rB12 = p[3] * twelve_rB12;
  const double AexpmBr = p[0] * exp(-p[1] * r);
  e = AexpmBr - Cr6 + Dtwelve_rB12;
  de = -p[1] * AexpmBr + (6 * Cr6 - 12 * Dtwelve_rB12) / r;
}

// src/potential/pair_potentials/buckingham/buckingham.h:41-52
static inline void buckingham_energy(const double* p /* A Rho C */, double x, double& e, double& de)
{
  const double x2 = x * x, x6 = x2 * x2 * x2, x7 = x6 * x;
  e = p[0] * exp(-x / p[1]) - (p[2] / x6);
  de = (6 * p[2] / x7) - (p[0] * exp(-x / p[1]) / p[1]);
}

// src/potential/pair_potentials/yukawa/include/exaStamp/potential/pair_potentials/yukawa/yukawa.h:39-48 (the derivative is
// restated as the reference writes it: de = e (1/r - kappa))
static inline void yukawa_compute_energy(const double* p /* A kappa */, double r, double& e, double& de)
{
  double ratio = p[0] / r;
  double rinv = 1. / r;
  e  = ratio * exp( - p[1] * r );
  de = e * ( rinv - p[1] );
}

// src/potential/pair_potentials/relax/potential.h:43-51 (overlap-relaxation ramp: r clamped to [r1, rc], de = -e)
static inline void relax_compute_energy(const double* p /* r1 rc */, double rij, double& e, double& de)
{
  double r = rij;
  if( r < p[0] ) r = p[0];
  if( r > p[1] ) r = p[1];
  e = ( p[1] / r ) - 1.0;
  de = - e;
}

// one pair evaluation of potential `pot` with its raw parameter vector (what USTAMP_POTENTIAL_COMPUTE expands to)
// 4 yukawa {A, kappa}; 5 relax {r1, rc}; 6 zero {} (src/potential/pair_potentials/zero/potential.h:49-54: e = de = 0)
struct PairPot
{
  int pot = 0; double prm[4] = {0, 0, 0, 0}; ZBLPair zbl{};
  PairPot() = default;
  PairPot(int pot_, const double* params) : pot(pot_)
  {
    const int n = nparams(pot);
    for(int i = 0; i < n; i++) prm[i] = params[i];
    if( pot == 1 ) zbl = zbl_pair(ZBLParams{ params[0], params[1] }, params[2], params[3]);
  }
  static int nparams(int pot) { return pot == 6 ? 0 : (pot == 0 || pot == 4 || pot == 5 ? 2 : (pot == 3 ? 3 : 4)); }
  inline void compute(double r, double& e, double& de) const
  {
    switch( pot )
    {
      case 0: lj_compute_energy(LJParams{ prm[0], prm[1] }, r, e, de); break;
      case 1: zbl_compute_energy(zbl, r, e, de); break;
      case 2: exp6_compute_energy(prm, r, e, de); break;
      case 3: buckingham_energy(prm, r, e, de); break;
      case 4: yukawa_compute_energy(prm, r, e, de); break;
      case 5: relax_compute_energy(prm, r, e, de); break;
      default: e = 0.0; de = 0.0; break;
    }
  }
  // pair_potential_impl.hxx:488-498 (energy_cutoff)
  inline double energy_cutoff(double rcut) const { double e = 0.0, de = 0.0; if( rcut > 0.0 ) compute(rcut, e, de); return e; }
};

// src/potential/eam_potentials/johnson/johnson.h:29-50 (19 scalars, same order)
struct JohnsonParams
{
  double re, fe, rhoe, alpha, beta, A, B, kappa, lambda, Fn0, Fn1, Fn2, Fn3, F0, F1, F2, F3, Fo, eta;
};

// johnson.h:56-91
static inline void johnson_phi(const JohnsonParams& p, double r, double& phi, double& dphi)
{
  const int n = 20, m = 20;
  const double x = r / p.re, ire = 1. / p.re;
  double c1 = p.A * std::exp(-p.alpha * (x - 1.));
  double c2 = x - p.kappa;
  double c3 = std::pow(c2, m);
  double num = c1, den = 1. + c3;
  double dnum = -p.alpha * c1, dden = m * c3 / c2;
  phi  = num / den;
  dphi = ire * (dnum * den - num * dden) / (den * den);
  c1 = -p.B * std::exp(-p.beta * (x - 1.));
  c2 = x - p.lambda;
  c3 = std::pow(c2, n);
  num = c1; den = 1. + c3;
  dnum = -p.beta * c1; dden = n * c3 / c2;
  phi  += num / den;
  dphi += ire * (dnum * den - num * dden) / (den * den);
}

// johnson.h:97-114
static inline void johnson_rho(const JohnsonParams& p, double r, double& rho, double& drho)
{
  const int n = 20;
  const double x = r / p.re, ire = 1. / p.re;
  const double c1 = p.fe * std::exp(-p.beta * (x - 1.));
  const double c2 = x - p.lambda;
  const double c3 = std::pow(c2, n);
  const double num = c1, den = 1. + c3;
  const double dnum = -p.beta * c1, dden = n * c3 / c2;
  rho  = num / den;
  drho = ire * (dnum * den - num * dden) / (den * den);
}

// johnson.h:120-166
static inline void johnson_fEmbed(const JohnsonParams& p, double rho, double& f, double& df)
{
  const double rhon = 0.85 * p.rhoe, rho0 = 1.15 * p.rhoe;
  const double irhon = 1. / rhon, irhoe = 1. / p.rhoe;
  if( rho < rhon )
  {
    const double q1 = rho / rhon - 1., q2 = q1 * q1, q3 = q1 * q2;
    f  = p.Fn0 + p.Fn1 * q1 + p.Fn2 * q2 + p.Fn3 * q3;
    df = p.Fn1 + 2. * p.Fn2 * q1 + 3. * p.Fn3 * q2;
    df *= irhon;
  }
  else if( rho < rho0 )
  {
    const double q1 = rho * irhoe - 1., q2 = q1 * q1, q3 = q1 * q2;
    f  = p.F0 + p.F1 * q1 + p.F2 * q2 + p.F3 * q3;
    df = p.F1 + 2. * p.F2 * q1 + 3. * p.F3 * q2;
    df *= irhoe;
  }
  else
  {
    const double raprho = rho * irhoe;
    const double rpe = std::pow(raprho, p.eta);
    const double lrpe = std::log(rpe);
    f  = p.Fo * (1. - lrpe) * rpe;
    df = -p.eta * rpe + (1. - lrpe) * p.eta * rpe;
    df *= p.Fo * irhoe / raprho;
  }
}

// src/potential/eam_potentials/sutton_chen/sutton_chen.h:33-62 (5 scalars: c, epsilon, a0, n, m)
struct SuttonChenParams { double c, epsilon, a0, n, m; };
static inline void sutton_chen_phi(const SuttonChenParams& p, double r, double& phiValue, double& dphi)
{
  double ratio = p.a0 / r;
  phiValue = p.epsilon * std::pow(ratio, p.n);
  dphi = -1 * p.n * phiValue / r;
}
static inline void sutton_chen_rho(const SuttonChenParams& p, double r, double& rhoValue, double& drho)
{
  double ratio = p.a0 / r;
  rhoValue = std::pow(ratio, p.m);
  drho = -1 * p.m * rhoValue / r;
}
static inline void sutton_chen_fEmbed(const SuttonChenParams& p, double rhoValue, double& f, double& df)
{
  double sqrtRho = std::sqrt(rhoValue);
  f  = -1. * p.c * p.epsilon * sqrtRho;
  df = 0.5 * f / (rhoValue > 0 ? rhoValue : 0);
}

// src/potential/eam_potentials/vniitf/vniitf.h:31-125 (13 scalars, same order)
struct VniitfParams { double rmax, rmin, rt0, Ecoh, E0, beta, A, Z, n, alpha, D, eta, mu; };
static inline double vniitf_dS3(double x) { double x2 = x * x; double x3 = x2 * x; return 140 * x3 * ( -1 * x3 + 3 * x2 - 3 * x + 1 ); }
static inline double vniitf_S3(double x) { double x2 = x * x; return x2 * x2 * ( -20 * x2 * x + 70 * x2 - 84 * x + 35 ); }
static inline void vniitf_switch(const VniitfParams& p, double r, double& S, double& dS)
{
  double drS = (p.rmax - r) / (p.rmax - p.rmin);
  S  = vniitf_S3(drS);
  dS = vniitf_dS3(drS) / (p.rmin - p.rmax);
  if( drS < 0 ) { S = 0.; dS = 0.; }
  else if( drS > 1 ) { S = 1.; dS = 0.; }
}
static inline void vniitf_phi(const VniitfParams& p, double r, double& phi, double& dphi)
{
  double ir = 1 / r;
  double irt0 = 1 / p.rt0;
  double dr = r * irt0 - 1.0;
  double dr2 = dr * dr;
  double a = -2 * p.Ecoh / p.Z;
  double b = p.alpha * p.alpha * p.alpha * p.D * p.rt0;
  double f1  = a * ( 1 + p.alpha * dr + p.eta * dr2 + (p.mu + b * ir) * dr2 * dr );
  double df1 = a * ( p.alpha * irt0 + 2 * p.eta * irt0 * dr + 3 * p.mu * irt0 * dr2 + b * (3 * irt0 - dr * ir) * dr2 * ir );
  double f2  = std::exp(-p.alpha * dr);
  double df2 = -p.alpha * irt0 * f2;
  double S, dS; vniitf_switch(p, r, S, dS);
  phi  = (p.E0 + f1 * f2) * S;
  dphi = (p.E0 + f1 * f2) * dS + (f1 * df2 + f2 * df1) * S;
}
static inline void vniitf_rho(const VniitfParams& p, double r, double& rho, double& drho)
{
  double irt0 = 1 / p.rt0;
  double F  = exp(-p.beta * (r * irt0 - 1.0)) / p.Z;
  double dF = -p.beta * F * irt0;
  double S, dS; vniitf_switch(p, r, S, dS);
  rho  = F * S;
  drho = F * dS + S * dF;
}
static inline void vniitf_fEmbed(const VniitfParams& p, double rho, double& f, double& df)
{
  if( rho <= 0. ) { f = 0.; df = 0.; }
  else
  {
    double a = pow(rho, p.n);
    double b = p.A * p.Ecoh * a;
    double c = log(a);
    f  = b * (c - 1);
    df = p.n * b * c / rho;
  }
}

// the analytic single-species models behind eam_potential_template (USTAMP_POTENTIAL_EAM_RHO / _PHI / _EMB):
// model ids as include/xsb200.h xsb_eam_model: 0 johnson (19 scalars), 1 sutton_chen (5), 2 vniitf (13)
struct EamAnalytic
{
  int model = 0; JohnsonParams j{}; SuttonChenParams sc{}; VniitfParams vn{};
  static int nparams(int model) { return model == 0 ? 19 : (model == 1 ? 5 : (model == 2 ? 13 : -1)); }
  EamAnalytic(int m, const double* prm) : model(m)
  {
    if( m == 0 ) std::memcpy(&j, prm, sizeof(j)); else if( m == 1 ) std::memcpy(&sc, prm, sizeof(sc)); else std::memcpy(&vn, prm, sizeof(vn));
  }
  inline void rho(double r, double& f, double& df) const { if( model == 0 ) johnson_rho(j, r, f, df); else if( model == 1 ) sutton_chen_rho(sc, r, f, df); else vniitf_rho(vn, r, f, df); }
  inline void phi(double r, double& f, double& df) const { if( model == 0 ) johnson_phi(j, r, f, df); else if( model == 1 ) sutton_chen_phi(sc, r, f, df); else vniitf_phi(vn, r, f, df); }
  inline void fEmbed(double x, double& f, double& df) const { if( model == 0 ) johnson_fEmbed(j, x, f, df); else if( model == 1 ) sutton_chen_fEmbed(sc, x, f, df); else vniitf_fEmbed(vn, x, f, df); }
};

// ---------------------------------------------------------------------------------------------
// eam/alloy (setfl) tables, LAMMPS pair_eam 7-coefficient splines.
// reader: src/potential/eam_potentials/eam_alloy/eam_alloy.cpp:66-278 ; interpolate :29-58 ;
// evaluation eam_alloy.h:185-313 ; type tables eam_alloy.h:70-133.
// Rows are stored 8 doubles wide like SplineCoeffs (eam_alloy.h:37-42); row 0 is unused.
// ---------------------------------------------------------------------------------------------
struct EamAlloy
{
  int nelements = 0, nr = 0, nrho = 0;
  double rdr = 0, rdrho = 0, rc = 0, rhomax = 0, dr = 0, drho = 0;
  std::vector<std::string> names;
  std::vector<double> mass;
  std::vector<double> frho;  // [nelements][nrho+1][8]
  std::vector<double> rhor;  // [nelements][nr+1][8]
  std::vector<double> z2r;   // [nelements(nelements+1)/2][nr+1][8]
  double conv_z2r = EV_ANG_INTERNAL, conv_frho = EV_INTERNAL;

  // type maps of eam_alloy.h:70-133 collapse (map[i]=i-1) to: rhor(i,j) -> table i ;
  // z2r(i,j) -> hi*(hi+1)/2+lo ; frho(i) -> table i  (0-based types)
  static inline int z2r_index(int a, int b) { int hi = a > b ? a : b, lo = a > b ? b : a; return hi * (hi + 1) / 2 + lo; }
};

static inline void eam_interpolate(int n, double delta, const double* f /*[n+1], 1-based*/, double* s /*[n+1][8]*/)
{
  auto S = [&](int m, int k) -> double& { return s[size_t(m) * 8 + k]; };
  for(int m = 1; m <= n; m++) S(m,6) = f[m];
  S(1,5)   = S(2,6) - S(1,6);
  S(2,5)   = 0.5 * (S(3,6) - S(1,6));
  S(n-1,5) = 0.5 * (S(n,6) - S(n-2,6));
  S(n,5)   = S(n,6) - S(n-1,6);
  for(int m = 3; m <= n - 2; m++) S(m,5) = ((S(m-2,6) - S(m+2,6)) + 8.0 * (S(m+1,6) - S(m-1,6))) / 12.0;
  for(int m = 1; m <= n - 1; m++)
  {
    S(m,4) = 3.0 * (S(m+1,6) - S(m,6)) - 2.0 * S(m,5) - S(m+1,5);
    S(m,3) = S(m,5) + S(m+1,5) - 2.0 * (S(m+1,6) - S(m,6));
  }
  S(n,4) = 0.0; S(n,3) = 0.0;
  for(int m = 1; m <= n; m++)
  {
    S(m,2) = S(m,5) / delta;
    S(m,1) = 2.0 * S(m,4) / delta;
    S(m,0) = 3.0 * S(m,3) / delta;
  }
}

static inline bool eam_alloy_read(const char* path, EamAlloy& v)
{
  std::ifstream file(path);
  if( !file ) return false;
  for(int i = 0; i < 3; i++) file.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
  size_t nel = 0; file >> nel;
  v.names.resize(nel); v.mass.resize(nel);
  for(size_t i = 0; i < nel; i++) file >> v.names[i];
  size_t nrho = 0, nr = 0; double drho = 0, dr = 0, rcut = 0;
  file >> nrho >> drho >> nr >> dr >> rcut;
  if( !file || nel == 0 || nr < 5 || nrho < 5 ) return false;
  v.nelements = int(nel); v.nr = int(nr); v.nrho = int(nrho);
  v.dr = dr; v.drho = drho; v.rdr = 1.0 / dr; v.rdrho = 1.0 / drho; v.rc = rcut; v.rhomax = (nrho - 1) * drho;
  const size_t nz2r = nel * (nel + 1) / 2;
  std::vector<std::vector<double>> frho(nel), rhor(nel), z2r(nz2r);
  for(size_t i = 0; i < nel; i++)
  {
    int z; double mass, a0; std::string st;
    file >> z >> mass >> a0 >> st;
    v.mass[i] = mass;
    frho[i].assign(nrho + 1, 0.0); rhor[i].assign(nr + 1, 0.0);
    for(size_t k = 0; k < nrho; k++) file >> frho[i][k + 1];
    for(size_t k = 0; k < nr; k++) file >> rhor[i][k + 1];
  }
  for(size_t i = 0; i < nz2r; i++)
  {
    z2r[i].assign(nr + 1, 0.0);
    for(size_t k = 0; k < nr; k++) file >> z2r[i][k + 1];
  }
  if( !file ) return false;
  v.frho.assign(nel * (nrho + 1) * 8, 0.0);
  v.rhor.assign(nel * (nr + 1) * 8, 0.0);
  v.z2r.assign(nz2r * (nr + 1) * 8, 0.0);
  for(size_t i = 0; i < nel; i++)  eam_interpolate(int(nrho), drho, frho[i].data(), v.frho.data() + i * (nrho + 1) * 8);
  for(size_t i = 0; i < nel; i++)  eam_interpolate(int(nr), dr, rhor[i].data(), v.rhor.data() + i * (nr + 1) * 8);
  for(size_t i = 0; i < nz2r; i++) eam_interpolate(int(nr), dr, z2r[i].data(), v.z2r.data() + i * (nr + 1) * 8);
  return true;
}

// eam_alloy.h:185-207 : rho contribution, table of element `itype`
static inline double eam_alloy_rho_noderiv(const EamAlloy& eam, double r, int itype, int /*jtype*/)
{
  double p = r * eam.rdr + 1.0;
  int m = static_cast<int>(p);
  m = std::min(m, eam.nr - 1);
  p -= m;
  p = std::min(p, 1.0);
  const double* c = eam.rhor.data() + (size_t(itype) * (eam.nr + 1) + m) * 8;
  return ((c[3] * p + c[4]) * p + c[5]) * p + c[6];
}

// eam_alloy.h:211-265 : returns fpair (reference multiplies dr in place) and phi
static inline double eam_alloy_mm_force(const EamAlloy& eam, double& phi, double r, double fpi, double fpj, int itype, int jtype)
{
  double p = r * eam.rdr + 1.0;
  int m = static_cast<int>(p);
  m = std::min(m, eam.nr - 1);
  p -= m;
  p = std::min(p, 1.0);
  const double* ci = eam.rhor.data() + (size_t(itype) * (eam.nr + 1) + m) * 8;
  const double rhoip = (ci[0] * p + ci[1]) * p + ci[2];
  const double* cj = eam.rhor.data() + (size_t(jtype) * (eam.nr + 1) + m) * 8;
  const double rhojp = (cj[0] * p + cj[1]) * p + cj[2];
  const double* c = eam.z2r.data() + (size_t(EamAlloy::z2r_index(itype, jtype)) * (eam.nr + 1) + m) * 8;
  const double z2p = (c[0] * p + c[1]) * p + c[2];
  const double z2  = ((c[3] * p + c[4]) * p + c[5]) * p + c[6];
  const double recip = 1.0 / r;
  phi = z2 * recip;
  const double phip = (z2p * recip - phi * recip) * eam.conv_z2r;
  phi *= eam.conv_z2r;
  const double psip = fpi * rhojp + fpj * rhoip + phip;
  return psip * recip;
}

// eam_alloy.h:289-313
static inline void eam_alloy_fEmbed(const EamAlloy& eam, double rho, double& phi, double& fp, int itype)
{
  double p = rho * eam.rdrho + 1.0;
  int m = static_cast<int>(p);
  m = std::max(1, std::min(m, eam.nrho - 1));
  p -= m;
  p = std::min(p, 1.0);
  const double* c = eam.frho.data() + (size_t(itype) * (eam.nrho + 1) + m) * 8;
  fp  = (c[0] * p + c[1]) * p + c[2];
  phi = ((c[3] * p + c[4]) * p + c[5]) * p + c[6];
  if( rho > eam.rhomax ) phi += fp * (rho - eam.rhomax);
  phi *= eam.conv_frho;
  fp  *= eam.conv_frho;
}

// ext exanb unique_pair_id (symmetric triangular index; used at
// pair_potential_force_op_multiparam.h:94 and eam_force_op_multimat.h:148)
static inline unsigned unique_pair_id(unsigned a, unsigned b) { if( a > b ) { unsigned t = a; a = b; b = t; } return b * (b + 1) / 2 + a; }

} // namespace orc
