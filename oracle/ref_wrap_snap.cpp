// TEST INFRASTRUCTURE ONLY.  extern "C" entry point around the reference's OWN, UNMODIFIED in-tree bispectrum code
// (src/potential/snaplegacy/lib/SnapLegacy{BS,CG,GSH}.cpp), compiled where it lies under /root/reference with
// -DLAMMPS (LAMMPS triple set j2 <= j1 <= j, theta0 = rfac0*pi*r/rcut with the hard-coded rfac0 = 0.99363,
// SnapLegacyGSH.cpp:38-49) into oracle/_ref/libxsref_snap.so.  Call sequence = tests/snap-compute-bs/snap-compute-bs.cpp:31-82.
// Used to pin oracle/snap_oracle.cpp and to generate tests/golden/snap_legacy.json.
#include <exaStamp/potential/snaplegacy/SnapLegacyBS.h>
#include <vector>

extern "C" int xsref_snap_nidx(double jmax) { return SnapLegacyBS::n_idx_bs(int(2 * jmax)); }

// bs[nidx] (real parts; the class divides by the number of in-range neighbours, SnapLegacyBS.cpp:278), dbs[n][nidx][3]
extern "C" int xsref_snap_bs(double jmax, double rcut, int n, const double* rx, const double* ry, const double* rz, double* bs_out, double* dbs_out, double* bs_imag_max)
{
  SnapLegacyCG cg(jmax, 2);
  cg.compute();
  const int nidx = SnapLegacyBS::n_idx_bs(int(2 * jmax));
  std::vector<double> coefs(size_t(nidx) + 1, 1.0); double factor[1] = { 1.0 };
  std::vector<int> species(n, 0);
  SnapLegacyBS bs(jmax, coefs.data(), factor);
  if( bs.set_neighbours(rx, ry, rz, species.data(), rcut, size_t(n)) ) return 1;
  if( bs.compute_cmm(rcut) ) return 2;
  if( bs.compute_bs(0, rcut, cg) ) return 3;
  double im = 0.0;
  for(int i = 0; i < nidx; i++) { bs_out[i] = bs.bs_val(i).real(); im = std::max(im, std::abs(bs.bs_val(i).imag())); }
  if( dbs_out )
    for(int k = 0; k < n * nidx; k++) { const complex3d d = bs.dbs_val(k); dbs_out[3*k] = d.x.real(); dbs_out[3*k+1] = d.y.real(); dbs_out[3*k+2] = d.z.real(); }
  if( bs_imag_max ) *bs_imag_max = im;
  return 0;
}
