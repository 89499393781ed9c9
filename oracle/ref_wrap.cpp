// TEST INFRASTRUCTURE ONLY.  extern "C" entry points around the reference's OWN, UNMODIFIED math headers,
// compiled where they lie under /root/reference (see oracle/Makefile; output oracle/_ref/libxsref.so).
// Used to pin the restatement in orc_math.h and to generate tests/golden/*.json.  No reference source is
// copied into this repository: this file only #includes it and forwards arguments.
#include <exaStamp/potential/pair_potentials/lennard_jones/lennard_jones.h>   // src/potential/pair_potentials/lennard_jones/include
#include "johnson.h"                                                           // src/potential/eam_potentials/johnson
#include "eam_alloy.h"                                                         // src/potential/eam_potentials/eam_alloy
#include <exaStamp/potential/pair_potentials/exp6/exp6.h>                      // src/potential/pair_potentials/exp6/include
#include "buckingham.h"                                                        // src/potential/pair_potentials/buckingham
#include <exaStamp/potential/pair_potentials/yukawa/yukawa.h>                  // src/potential/pair_potentials/yukawa/include
#include "sutton_chen.h"                                                       // src/potential/eam_potentials/sutton_chen
#include "vniitf.h"                                                            // src/potential/eam_potentials/vniitf
#include <cstring>
// zbl/potential.h also defines the USTAMP_* template macros: harmless here, nothing expands them
#include "zbl/potential.h"                                                     // src/potential/pair_potentials/zbl
// relax/potential.h and zero/potential.h define their function next to the same USTAMP_* macros: undefine between includes
#undef USTAMP_POTENTIAL_NAME
#undef USTAMP_POTENTIAL_PARAMS
#undef USTAMP_POTENTIAL_COMPUTE
#undef USTAMP_POTENTIAL_PAIR_PARAMS_EXTRACT
#undef USTAMP_POTENTIAL_ENABLE_CUDA
#undef USTAMP_POTENTIAL_ENABLE_RIGIDMOL
#include "relax/potential.h"                                                   // src/potential/pair_potentials/relax
#undef USTAMP_POTENTIAL_NAME
#undef USTAMP_POTENTIAL_PARAMS
#undef USTAMP_POTENTIAL_COMPUTE
#undef USTAMP_POTENTIAL_PAIR_PARAMS_EXTRACT
#undef USTAMP_POTENTIAL_ENABLE_CUDA
#undef USTAMP_POTENTIAL_ENABLE_RIGIDMOL
#include "zero/potential.h"                                                    // src/potential/pair_potentials/zero

using namespace exaStamp;

extern "C" {

void xsref_lj(double epsilon, double sigma, double r, double* e, double* de)
{
  LennardJonesParms p{ epsilon, sigma };
  PairPotentialMinimalParameters pp{};
  lj_compute_energy(p, pp, r, *e, *de);
}

// pot ids as in include/xsb200.h: 1 zbl {r1, rc, z_a, z_b}, 2 exp6 {A, B, C, D}, 3 buckingham {A, Rho, C}, 4 yukawa {A, kappa},
// 5 relax {r1, rc}, 6 zero {}
void xsref_pair(int pot, const double* prm, double r, double* e, double* de)
{
  PairPotentialMinimalParameters pp{};
  double a = 0.0, b = 0.0;       // the operators initialise e, de to 0 before the call (force_op_impl2.hxx:37)
  if( pot == 0 ) { LennardJonesParms p{ prm[0], prm[1] }; lj_compute_energy(p, pp, r, a, b); }
  else if( pot == 1 ) { ZBLParms p{}; p.r1 = prm[0]; p.rc = prm[1]; pp.m_atom_a.m_z = unsigned(prm[2]); pp.m_atom_b.m_z = unsigned(prm[3]); zbl_compute_energy(p, pp, r, a, b); }
  else if( pot == 2 ) { Exp6Parms p{ prm[0], prm[1], prm[2], prm[3] }; exp6_compute_energy(p, pp, r, a, b); }
  else if( pot == 3 ) { BuckinghamParms p{ prm[0], prm[1], prm[2] }; buckingham_energy(p, pp, r, a, b); }
  else if( pot == 4 ) { YukawaParms p{ prm[0], prm[1] }; yukawa_compute_energy(p, pp, r, a, b); }
  else if( pot == 5 ) { RelaxParms p{ prm[0], prm[1] }; relax_compute_energy(p, pp, r, a, b); }
  else { ZeroPotentialParameters p{}; zero_potential_compute_force(p, pp, r, a, b); }
  *e = a; *de = b;
}

// single-species analytic EAM models: 0 johnson, 1 sutton_chen {c, epsilon, a0, n, m}, 2 vniitf (13 scalars); what: 0 phi, 1 rho, 2 fEmbed
void xsref_eam_analytic(int model, const double* prm, int what, double x, double* f, double* df)
{
  if( model == 0 ) { EamJohnsonParameters p; std::memcpy(&p, prm, sizeof(p)); if( what == 0 ) eam_johnson_phi(p, x, *f, *df); else if( what == 1 ) eam_johnson_rho(p, x, *f, *df); else eam_johnson_fEmbed(p, x, *f, *df); }
  else if( model == 1 )
  {
    EamSuttonChenParameters p; static_assert( sizeof(p) == 5 * sizeof(double) ); std::memcpy(&p, prm, sizeof(p));
    if( what == 0 ) sutton_chen_phi(p, x, *f, *df); else if( what == 1 ) sutton_chen_rho(p, x, *f, *df); else sutton_chen_fEmbed(p, x, *f, *df);
  }
  else
  {
    EamVniitfParameters p; static_assert( sizeof(p) == 13 * sizeof(double) ); std::memcpy(&p, prm, sizeof(p));
    if( what == 0 ) eam_vniitf_phi(p, x, *f, *df); else if( what == 1 ) eam_vniitf_rho(p, x, *f, *df); else eam_vniitf_fEmbed(p, x, *f, *df);
  }
}

void xsref_johnson(const double* params19, int what, double x, double* f, double* df)
{
  EamJohnsonParameters p; static_assert( sizeof(p) == 19 * sizeof(double) ); std::memcpy(&p, params19, sizeof(p));
  if( what == 0 ) eam_johnson_phi(p, x, *f, *df);
  else if( what == 1 ) eam_johnson_rho(p, x, *f, *df);
  else eam_johnson_fEmbed(p, x, *f, *df);
}

void* xsref_eam_alloy_load(const char* path, int n_types)
{
  YAML::Node node; node.scalar = path;
  EamAlloyParameters* v = new EamAlloyParameters;
  if( !YAML::convert<EamAlloyParameters>::decode(node, *v) || v->nr == 0 ) { delete v; return nullptr; }
  v->initialize_types_table( n_types > 0 ? n_types : v->nelements , nullptr );
  return v;
}
void xsref_eam_alloy_free(void* h) { delete static_cast<EamAlloyParameters*>(h); }
void xsref_eam_alloy_info(void* h, int* nelements, int* nr, int* nrho, double* rdr, double* rdrho, double* rc, double* rhomax)
{
  const auto& e = *static_cast<EamAlloyParameters*>(h);
  *nelements = e.nelements; *nr = e.nr; *nrho = e.nrho; *rdr = e.rdr; *rdrho = e.rdrho; *rc = e.rc; *rhomax = e.rhomax;
}
const double* xsref_eam_alloy_table(void* h, int which)
{
  const auto& e = *static_cast<EamAlloyParameters*>(h);
  const auto& v = which == 0 ? e.frho_spline_data : which == 1 ? e.rhor_spline_data : e.z2r_spline_data;
  return v.data()->coeffs;
}
// what: 0 rho_noderiv(r; ti,tj) ; 1 fEmbed(rho; ti) -> phi, *out2 = fp ; 2 mm_force(r, fpi, fpj; ti,tj) -> fpair, *out2 = phi
double xsref_eam_alloy_eval(void* h, int what, double x, int ti, int tj, double fpi, double fpj, double* out2)
{
  EamAlloyParametersRO ro( *static_cast<EamAlloyParameters*>(h) );
  if( what == 0 ) return eam_alloy_rho_noderiv(ro, x, ti, tj);
  if( what == 1 ) { double phi = 0, fp = 0; eam_alloy_fEmbed(ro, x, phi, fp, ti); *out2 = fp; return phi; }
  Vec3d dr{ 1.0, 0.0, 0.0 }; double phi = 0;
  eam_alloy_mm_force(ro, dr, phi, x, fpi, fpj, ti, tj);
  *out2 = phi; return dr.x;
}
double xsref_ev_internal() { return EamAlloyParametersRO::conversion_frho; }

} // extern "C"
