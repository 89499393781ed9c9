// TEST INFRASTRUCTURE ONLY.  extern "C" entry points around the reference's OWN, UNMODIFIED math headers,
// compiled where they lie under /root/reference (see oracle/Makefile; output oracle/_ref/libxsref.so).
// Used to pin the restatement in orc_math.h and to generate tests/golden/*.json.  No reference source is
// copied into this repository: this file only #includes it and forwards arguments.
#include <exaStamp/potential/pair_potentials/lennard_jones/lennard_jones.h>   // src/potential/pair_potentials/lennard_jones/include
#include "johnson.h"                                                           // src/potential/eam_potentials/johnson
#include "eam_alloy.h"                                                         // src/potential/eam_potentials/eam_alloy
#include <cstring>

using namespace exaStamp;

extern "C" {

void xsref_lj(double epsilon, double sigma, double r, double* e, double* de)
{
  LennardJonesParms p{ epsilon, sigma };
  PairPotentialMinimalParameters pp{};
  lj_compute_energy(p, pp, r, *e, *de);
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
