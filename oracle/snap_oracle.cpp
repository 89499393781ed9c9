// TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into or called by the product path.
//
// snap_oracle.cpp : SNAP bispectrum / energy / force in double precision.
//
// The production operator `snap_force` of the reference is a 39-line shell around exaNBody's md::SnapForceGeneric
// (src/potential/snap/snap_force.cu:18-37), which is NOT in /root/reference (un-vendored exaNBody v2.0.1,
// contribs/md/snap).  What the reference does show is the call sequence of the LAMMPS SNA class that exaNBody's
// implementation follows -- src/potential/snaplmp/snap_force_op.h:177-337 (compute_ui, compute_yi, per neighbour
// compute_duidrj + compute_deidrj, f_i += fij, f_j -= fij, virial -fij (x) rij on the centre, energy e0 + beta.B)
// and snap_bispectrum_op.h:123-131 (compute_ui, compute_zi, compute_bi) -- and the constructor arguments
// (snaplmp.cpp:205-216).  This file restates the published algorithm of LAMMPS ML-SNAP sna.cpp (Thompson et al.,
// JCP 285, 316 (2015); index tables idxu/idxz/idxb/idxcg, VMK 4.4(2) inversion symmetry, adjoint Y of Bartok/Wood)
// in that call order.
//
// PARITY PINNING: the bispectrum B and its derivatives dB/dr_j are pinned against the reference's own in-tree
// Bartok-style implementation SnapLegacyBS/CG/GSH (src/potential/snaplegacy/lib, compiled unmodified with -DLAMMPS
// into oracle/_ref/libxsref_snap.so) and against its stored vectors tests/snap-compute-bs/bs.ref2 (through the
// (2j+1) symmetry of B); forces are additionally pinned by finite differences of the energy.  The production
// `snap_force` outputs themselves are "parity unpinned" (its .dat fixtures are not in the reference tree).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace
{

struct Sna
{
  int twojmax = 0, switchflag = 1, bzeroflag = 0, nelements = 1, ncoeff = 0;
  double rfac0 = 0.99363, rmin0 = 0.0, rcutfac = 0.0, wself = 1.0;
  std::vector<double> radelem, wjelem, beta;          // beta: [nelements][ncoeff+1] (beta0 first), already in output energy units
  int idxu_max = 0, idxz_max = 0, idxb_max = 0, idxcg_max = 0, jdim = 0;
  std::vector<int> idxu_block, idxcg_block, idxz_block, idxb_block;   // [jdim] / [jdim^3]
  struct Z { int j1, j2, j, ma1min, ma2max, mb1min, mb2max, na, nb, jju; };
  struct B3 { int j1, j2, j; };
  std::vector<Z> idxz; std::vector<B3> idxb;
  std::vector<double> rootpq, cglist, bzero;

  int b3(int j1, int j2, int j) const { return (j1 * jdim + j2) * jdim + j; }
  double rpq(int p, int q) const { return rootpq[size_t(p) * (jdim + 1) + q]; }

  static double factorial(int n) { double f = 1.0; for(int i = 2; i <= n; i++) f *= i; return f; }
  static double deltacg(int j1, int j2, int j)
  {
    const double sfaccg = factorial((j1 + j2 + j) / 2 + 1);
    return std::sqrt(factorial((j1 + j2 - j) / 2) * factorial((j1 - j2 + j) / 2) * factorial((-j1 + j2 + j) / 2) / sfaccg);
  }

  void init()
  {
    jdim = twojmax + 1;
    idxu_block.assign(jdim, 0); idxcg_block.assign(size_t(jdim) * jdim * jdim, 0); idxz_block = idxcg_block; idxb_block = idxcg_block;
    int c = 0;
    for(int j1 = 0; j1 <= twojmax; j1++) for(int j2 = 0; j2 <= j1; j2++) for(int j = j1 - j2; j <= std::min(twojmax, j1 + j2); j += 2)
    { idxcg_block[b3(j1, j2, j)] = c; c += (j1 + 1) * (j2 + 1); }
    idxcg_max = c;
    c = 0;
    for(int j = 0; j <= twojmax; j++) { idxu_block[j] = c; c += (j + 1) * (j + 1); }
    idxu_max = c;
    idxb.clear();
    for(int j1 = 0; j1 <= twojmax; j1++) for(int j2 = 0; j2 <= j1; j2++) for(int j = j1 - j2; j <= std::min(twojmax, j1 + j2); j += 2)
      if( j >= j1 ) { idxb_block[b3(j1, j2, j)] = int(idxb.size()); idxb.push_back(B3{j1, j2, j}); }
    idxb_max = int(idxb.size()); ncoeff = idxb_max;
    idxz.clear();
    for(int j1 = 0; j1 <= twojmax; j1++) for(int j2 = 0; j2 <= j1; j2++) for(int j = j1 - j2; j <= std::min(twojmax, j1 + j2); j += 2)
    {
      idxz_block[b3(j1, j2, j)] = int(idxz.size());
      for(int mb = 0; 2 * mb <= j; mb++) for(int ma = 0; ma <= j; ma++)
      {
        Z z; z.j1 = j1; z.j2 = j2; z.j = j;
        z.ma1min = std::max(0, (2 * ma - j - j2 + j1) / 2);
        z.ma2max = (2 * ma - j - (2 * z.ma1min - j1) + j2) / 2;
        z.na = std::min(j1, (2 * ma - j + j2 + j1) / 2) - z.ma1min + 1;
        z.mb1min = std::max(0, (2 * mb - j - j2 + j1) / 2);
        z.mb2max = (2 * mb - j - (2 * z.mb1min - j1) + j2) / 2;
        z.nb = std::min(j1, (2 * mb - j + j2 + j1) / 2) - z.mb1min + 1;
        z.jju = idxu_block[j] + (j + 1) * mb + ma;
        idxz.push_back(z);
      }
    }
    idxz_max = int(idxz.size());
    rootpq.assign(size_t(jdim + 1) * (jdim + 1), 0.0);
    for(int p = 1; p <= twojmax; p++) for(int q = 1; q <= twojmax; q++) rootpq[size_t(p) * (jdim + 1) + q] = std::sqrt(double(p) / q);
    // Clebsch-Gordan coefficients (VMK 8.2.1(3)), doubled integer indices
    cglist.assign(idxcg_max, 0.0);
    c = 0;
    for(int j1 = 0; j1 <= twojmax; j1++) for(int j2 = 0; j2 <= j1; j2++) for(int j = j1 - j2; j <= std::min(twojmax, j1 + j2); j += 2)
      for(int m1 = 0; m1 <= j1; m1++)
      {
        const int aa2 = 2 * m1 - j1;
        for(int m2 = 0; m2 <= j2; m2++)
        {
          const int bb2 = 2 * m2 - j2, m = (aa2 + bb2 + j) / 2;
          if( m < 0 || m > j ) { cglist[c++] = 0.0; continue; }
          double sum = 0.0;
          for(int z = std::max(0, std::max(-(j - j2 + aa2) / 2, -(j - j1 - bb2) / 2));
              z <= std::min((j1 + j2 - j) / 2, std::min((j1 - aa2) / 2, (j2 + bb2) / 2)); z++)
          {
            const int ifac = z % 2 ? -1 : 1;
            sum += ifac / (factorial(z) * factorial((j1 + j2 - j) / 2 - z) * factorial((j1 - aa2) / 2 - z) * factorial((j2 + bb2) / 2 - z) *
                           factorial((j - j2 + aa2) / 2 + z) * factorial((j - j1 - bb2) / 2 + z));
          }
          const int cc2 = 2 * m - j;
          const double dcg = deltacg(j1, j2, j);
          const double sfaccg = std::sqrt(factorial((j1 + aa2) / 2) * factorial((j1 - aa2) / 2) * factorial((j2 + bb2) / 2) * factorial((j2 - bb2) / 2) *
                                          factorial((j + cc2) / 2) * factorial((j - cc2) / 2) * (j + 1));
          cglist[c++] = sum * dcg * sfaccg;
        }
      }
    bzero.assign(jdim, 0.0);
    if( bzeroflag ) { const double www = wself * wself * wself; for(int j = 0; j <= twojmax; j++) bzero[j] = www * (j + 1); }
  }

  double sfac(double r, double rcut) const
  {
    if( switchflag == 0 ) return 1.0;
    if( r <= rmin0 ) return 1.0;
    if( r > rcut ) return 0.0;
    return 0.5 * (std::cos((r - rmin0) * M_PI / (rcut - rmin0)) + 1.0);
  }
  double dsfac(double r, double rcut) const
  {
    if( switchflag == 0 ) return 0.0;
    if( r <= rmin0 ) return 0.0;
    if( r > rcut ) return 0.0;
    const double rcutfac_ = M_PI / (rcut - rmin0);
    return -0.5 * std::sin((r - rmin0) * rcutfac_) * rcutfac_;
  }

  // ---- per-atom work arrays ----
  struct Work
  {
    std::vector<double> utot_r, utot_i, u_r, u_i, z_r, z_i, y_r, y_i, du_r, du_i, blist;
  };
  void alloc(Work& w) const
  {
    w.utot_r.assign(idxu_max, 0); w.utot_i = w.utot_r; w.u_r = w.utot_r; w.u_i = w.utot_r; w.y_r = w.utot_r; w.y_i = w.utot_r;
    w.z_r.assign(idxz_max, 0); w.z_i = w.z_r; w.du_r.assign(size_t(idxu_max) * 3, 0); w.du_i = w.du_r; w.blist.assign(idxb_max, 0);
  }

  void uarray(Work& w, double x, double y, double z, double z0, double r) const
  {
    const double r0inv = 1.0 / std::sqrt(r * r + z0 * z0);
    const double a_r = r0inv * z0, a_i = -r0inv * z, b_r = r0inv * y, b_i = -r0inv * x;
    double* ur = w.u_r.data(); double* ui = w.u_i.data();
    ur[0] = 1.0; ui[0] = 0.0;
    for(int j = 1; j <= twojmax; j++)
    {
      int jju = idxu_block[j], jjup = idxu_block[j - 1];
      for(int mb = 0; 2 * mb <= j; mb++)
      {
        ur[jju] = 0.0; ui[jju] = 0.0;
        for(int ma = 0; ma < j; ma++)
        {
          double q = rpq(j - ma, j - mb);
          ur[jju] += q * (a_r * ur[jjup] + a_i * ui[jjup]);
          ui[jju] += q * (a_r * ui[jjup] - a_i * ur[jjup]);
          q = rpq(ma + 1, j - mb);
          ur[jju + 1] = -q * (b_r * ur[jjup] + b_i * ui[jjup]);
          ui[jju + 1] = -q * (b_r * ui[jjup] - b_i * ur[jjup]);
          jju++; jjup++;
        }
        jju++;
      }
      // right half by inversion symmetry u[j-ma][j-mb] = (-1)^(ma-mb) conj(u[ma][mb])
      jju = idxu_block[j]; jjup = jju + (j + 1) * (j + 1) - 1;
      int mbpar = 1;
      for(int mb = 0; 2 * mb <= j; mb++)
      {
        int mapar = mbpar;
        for(int ma = 0; ma <= j; ma++)
        {
          if( mapar == 1 ) { ur[jjup] = ur[jju]; ui[jjup] = -ui[jju]; } else { ur[jjup] = -ur[jju]; ui[jjup] = ui[jju]; }
          mapar = -mapar; jju++; jjup--;
        }
        mbpar = -mbpar;
      }
    }
  }

  void compute_ui(Work& w, int n, const double* dx, const double* dy, const double* dz, const double* wj, const double* rcut) const
  {
    for(int k = 0; k < idxu_max; k++) { w.utot_r[k] = 0.0; w.utot_i[k] = 0.0; }
    for(int j = 0; j <= twojmax; j++) { int jju = idxu_block[j]; for(int ma = 0; ma <= j; ma++) { w.utot_r[jju] = wself; jju += j + 2; } }
    for(int i = 0; i < n; i++)
    {
      const double x = dx[i], y = dy[i], z = dz[i], r = std::sqrt(x * x + y * y + z * z);
      const double theta0 = (r - rmin0) * rfac0 * M_PI / (rcut[i] - rmin0);
      const double z0 = r / std::tan(theta0);
      uarray(w, x, y, z, z0, r);
      const double s = sfac(r, rcut[i]) * wj[i];
      for(int k = 0; k < idxu_max; k++) { w.utot_r[k] += s * w.u_r[k]; w.utot_i[k] += s * w.u_i[k]; }
    }
  }

  void zelem(const Work& w, const Z& q, double& zr, double& zi) const
  {
    const double* cgblock = cglist.data() + idxcg_block[b3(q.j1, q.j2, q.j)];
    zr = 0.0; zi = 0.0;
    int jju1 = idxu_block[q.j1] + (q.j1 + 1) * q.mb1min, jju2 = idxu_block[q.j2] + (q.j2 + 1) * q.mb2max, icgb = q.mb1min * (q.j2 + 1) + q.mb2max;
    for(int ib = 0; ib < q.nb; ib++)
    {
      double sr = 0.0, si = 0.0;
      const double *u1r = &w.utot_r[jju1], *u1i = &w.utot_i[jju1], *u2r = &w.utot_r[jju2], *u2i = &w.utot_i[jju2];
      int ma1 = q.ma1min, ma2 = q.ma2max, icga = q.ma1min * (q.j2 + 1) + q.ma2max;
      for(int ia = 0; ia < q.na; ia++)
      {
        sr += cgblock[icga] * (u1r[ma1] * u2r[ma2] - u1i[ma1] * u2i[ma2]);
        si += cgblock[icga] * (u1r[ma1] * u2i[ma2] + u1i[ma1] * u2r[ma2]);
        ma1++; ma2--; icga += q.j2;
      }
      zr += cgblock[icgb] * sr; zi += cgblock[icgb] * si;
      jju1 += q.j1 + 1; jju2 -= q.j2 + 1; icgb += q.j2;
    }
  }

  void compute_zi(Work& w) const { for(int jjz = 0; jjz < idxz_max; jjz++) zelem(w, idxz[jjz], w.z_r[jjz], w.z_i[jjz]); }

  void compute_bi(Work& w) const
  {
    for(int jjb = 0; jjb < idxb_max; jjb++)
    {
      const B3& t = idxb[jjb];
      int jjz = idxz_block[b3(t.j1, t.j2, t.j)], jju = idxu_block[t.j];
      double sumzu = 0.0;
      for(int mb = 0; 2 * mb < t.j; mb++) for(int ma = 0; ma <= t.j; ma++) { sumzu += w.utot_r[jju] * w.z_r[jjz] + w.utot_i[jju] * w.z_i[jjz]; jjz++; jju++; }
      if( t.j % 2 == 0 )
      {
        const int mb = t.j / 2;
        for(int ma = 0; ma < mb; ma++) { sumzu += w.utot_r[jju] * w.z_r[jjz] + w.utot_i[jju] * w.z_i[jjz]; jjz++; jju++; }
        sumzu += 0.5 * (w.utot_r[jju] * w.z_r[jjz] + w.utot_i[jju] * w.z_i[jjz]);
      }
      w.blist[jjb] = 2.0 * sumzu - (bzeroflag ? bzero[t.j] : 0.0);
    }
  }

  void compute_yi(Work& w, const double* b /* ncoeff, without beta0 */) const
  {
    for(int k = 0; k < idxu_max; k++) { w.y_r[k] = 0.0; w.y_i[k] = 0.0; }
    for(int jjz = 0; jjz < idxz_max; jjz++)
    {
      const Z& q = idxz[jjz];
      double zr, zi; zelem(w, q, zr, zi);
      const int j1 = q.j1, j2 = q.j2, j = q.j;
      double betaj;
      if( j >= j1 ) { const int jjb = idxb_block[b3(j1, j2, j)]; betaj = (j1 == j) ? ((j2 == j) ? 3.0 * b[jjb] : 2.0 * b[jjb]) : b[jjb]; }
      else if( j >= j2 ) { const int jjb = idxb_block[b3(j, j2, j1)]; betaj = ((j2 == j) ? 2.0 * b[jjb] : b[jjb]) * (j1 + 1) / (j + 1.0); }
      else { const int jjb = idxb_block[b3(j2, j, j1)]; betaj = b[jjb] * (j1 + 1) / (j + 1.0); }
      w.y_r[q.jju] += betaj * zr; w.y_i[q.jju] += betaj * zi;
    }
  }

  void compute_duidrj(Work& w, double x, double y, double z, double wj_, double rcut) const
  {
    const double rsq = x * x + y * y + z * z, r = std::sqrt(rsq);
    const double rscale0 = rfac0 * M_PI / (rcut - rmin0), theta0 = (r - rmin0) * rscale0, cs = std::cos(theta0), sn = std::sin(theta0);
    const double z0 = r * cs / sn, dz0dr = z0 / r - (r * rscale0) * (rsq + z0 * z0) / rsq;
    const double rinv = 1.0 / r, u[3] = { x * rinv, y * rinv, z * rinv };
    const double r0inv = 1.0 / std::sqrt(r * r + z0 * z0);
    const double a_r = z0 * r0inv, a_i = -z * r0inv, b_r = y * r0inv, b_i = -x * r0inv;
    const double dr0invdr = -std::pow(r0inv, 3.0) * (r + z0 * dz0dr);
    double dr0inv[3], dz0[3], da_r[3], da_i[3], db_r[3], db_i[3];
    for(int k = 0; k < 3; k++) { dr0inv[k] = dr0invdr * u[k]; dz0[k] = dz0dr * u[k]; da_r[k] = dz0[k] * r0inv + z0 * dr0inv[k]; da_i[k] = -z * dr0inv[k]; }
    da_i[2] += -r0inv;
    for(int k = 0; k < 3; k++) { db_r[k] = y * dr0inv[k]; db_i[k] = -x * dr0inv[k]; }
    db_i[0] += -r0inv; db_r[1] += r0inv;
    uarray(w, x, y, z, z0, r);
    double* ur = w.u_r.data(); double* ui = w.u_i.data(); double* dr_ = w.du_r.data(); double* di_ = w.du_i.data();
    for(int k = 0; k < 3; k++) { dr_[k] = 0.0; di_[k] = 0.0; }
    for(int j = 1; j <= twojmax; j++)
    {
      int jju = idxu_block[j], jjup = idxu_block[j - 1];
      for(int mb = 0; 2 * mb <= j; mb++)
      {
        for(int k = 0; k < 3; k++) { dr_[3 * jju + k] = 0.0; di_[3 * jju + k] = 0.0; }
        for(int ma = 0; ma < j; ma++)
        {
          double q = rpq(j - ma, j - mb);
          for(int k = 0; k < 3; k++)
          {
            dr_[3 * jju + k] += q * (da_r[k] * ur[jjup] + da_i[k] * ui[jjup] + a_r * dr_[3 * jjup + k] + a_i * di_[3 * jjup + k]);
            di_[3 * jju + k] += q * (da_r[k] * ui[jjup] - da_i[k] * ur[jjup] + a_r * di_[3 * jjup + k] - a_i * dr_[3 * jjup + k]);
          }
          q = rpq(ma + 1, j - mb);
          for(int k = 0; k < 3; k++)
          {
            dr_[3 * (jju + 1) + k] = -q * (db_r[k] * ur[jjup] + db_i[k] * ui[jjup] + b_r * dr_[3 * jjup + k] + b_i * di_[3 * jjup + k]);
            di_[3 * (jju + 1) + k] = -q * (db_r[k] * ui[jjup] - db_i[k] * ur[jjup] + b_r * di_[3 * jjup + k] - b_i * dr_[3 * jjup + k]);
          }
          jju++; jjup++;
        }
        jju++;
      }
      jju = idxu_block[j]; jjup = jju + (j + 1) * (j + 1) - 1;
      int mbpar = 1;
      for(int mb = 0; 2 * mb <= j; mb++)
      {
        int mapar = mbpar;
        for(int ma = 0; ma <= j; ma++)
        {
          for(int k = 0; k < 3; k++)
          {
            if( mapar == 1 ) { dr_[3 * jjup + k] = dr_[3 * jju + k]; di_[3 * jjup + k] = -di_[3 * jju + k]; }
            else { dr_[3 * jjup + k] = -dr_[3 * jju + k]; di_[3 * jjup + k] = di_[3 * jju + k]; }
          }
          mapar = -mapar; jju++; jjup--;
        }
        mbpar = -mbpar;
      }
    }
    const double s = sfac(r, rcut) * wj_, ds = dsfac(r, rcut) * wj_;
    for(int jju = 0; jju < idxu_max; jju++) for(int k = 0; k < 3; k++)
    {
      dr_[3 * jju + k] = ds * ur[jju] * u[k] + s * dr_[3 * jju + k];
      di_[3 * jju + k] = ds * ui[jju] * u[k] + s * di_[3 * jju + k];
    }
  }

  void compute_deidrj(const Work& w, double* dedr) const
  {
    dedr[0] = dedr[1] = dedr[2] = 0.0;
    for(int j = 0; j <= twojmax; j++)
    {
      int jju = idxu_block[j];
      for(int mb = 0; 2 * mb < j; mb++) for(int ma = 0; ma <= j; ma++)
      { for(int k = 0; k < 3; k++) dedr[k] += w.du_r[3 * jju + k] * w.y_r[jju] + w.du_i[3 * jju + k] * w.y_i[jju]; jju++; }
      if( j % 2 == 0 )
      {
        const int mb = j / 2;
        for(int ma = 0; ma < mb; ma++) { for(int k = 0; k < 3; k++) dedr[k] += w.du_r[3 * jju + k] * w.y_r[jju] + w.du_i[3 * jju + k] * w.y_i[jju]; jju++; }
        for(int k = 0; k < 3; k++) dedr[k] += (w.du_r[3 * jju + k] * w.y_r[jju] + w.du_i[3 * jju + k] * w.y_i[jju]) * 0.5;
      }
    }
    for(int k = 0; k < 3; k++) dedr[k] *= 2.0;
  }

  // dB_k/dr_j for every component (compute_dbidrj of sna.cpp): used only to pin against SnapLegacyBS::dbs
  void compute_dbidrj(const Work& w, double* dbdr /* [ncoeff][3] */) const
  {
    for(int jjb = 0; jjb < idxb_max; jjb++)
    {
      const B3& t = idxb[jjb];
      const int j1 = t.j1, j2 = t.j2, j = t.j;
      double* out = dbdr + 3 * jjb;
      auto term = [&](int ja, int jb, int jc, double fac, double* acc)
      {
        // sum over half of layer jc of Conj(dudr(jc,ma,mb)) * z(ja,jb,jc,ma,mb)
        int jjz = idxz_block[b3(ja, jb, jc)], jju = idxu_block[jc];
        double s[3] = {0, 0, 0};
        for(int mb = 0; 2 * mb < jc; mb++) for(int ma = 0; ma <= jc; ma++)
        { for(int k = 0; k < 3; k++) s[k] += w.du_r[3 * jju + k] * w.z_r[jjz] + w.du_i[3 * jju + k] * w.z_i[jjz]; jjz++; jju++; }
        if( jc % 2 == 0 )
        {
          const int mb = jc / 2;
          for(int ma = 0; ma < mb; ma++) { for(int k = 0; k < 3; k++) s[k] += w.du_r[3 * jju + k] * w.z_r[jjz] + w.du_i[3 * jju + k] * w.z_i[jjz]; jjz++; jju++; }
          for(int k = 0; k < 3; k++) s[k] += (w.du_r[3 * jju + k] * w.z_r[jjz] + w.du_i[3 * jju + k] * w.z_i[jjz]) * 0.5;
        }
        for(int k = 0; k < 3; k++) acc[k] += 2.0 * s[k] * fac;
      };
      out[0] = out[1] = out[2] = 0.0;
      term(j1, j2, j, 1.0, out);
      // z(j,j2,j1) and z(j,j1,j2) exist in idxz only with the smaller second index first: j2 <= j1 <= j
      term(j, j2, j1, double(j + 1) / (j1 + 1.0), out);
      term(j, j1, j2, double(j + 1) / (j2 + 1.0), out);
    }
  }
};

} // namespace

extern "C" {

void* orc_snap_create(int twojmax, double rfac0, double rmin0, int switchflag, int bzeroflag, int nelements, const double* radelem, const double* wjelem,
                      double rcutfac, const double* beta)
{
  Sna* s = new Sna;
  s->twojmax = twojmax; s->rfac0 = rfac0; s->rmin0 = rmin0; s->switchflag = switchflag; s->bzeroflag = bzeroflag; s->nelements = nelements; s->rcutfac = rcutfac;
  s->radelem.assign(radelem, radelem + nelements); s->wjelem.assign(wjelem, wjelem + nelements);
  s->init();
  if( beta ) s->beta.assign(beta, beta + size_t(nelements) * (s->ncoeff + 1)); else s->beta.assign(size_t(nelements) * (s->ncoeff + 1), 0.0);
  return s;
}
void orc_snap_free(void* h) { delete static_cast<Sna*>(h); }
int orc_snap_ncoeff(void* h) { return static_cast<Sna*>(h)->ncoeff; }
void orc_snap_sizes(void* h, int* idxu_max, int* idxz_max, int* idxb_max, int* idxcg_max)
{ Sna* s = static_cast<Sna*>(h); *idxu_max = s->idxu_max; *idxz_max = s->idxz_max; *idxb_max = s->idxb_max; *idxcg_max = s->idxcg_max; }
void orc_snap_cut(void* h, int elem_i, int elem_j, double* rcut) { const Sna& s = *static_cast<Sna*>(h); *rcut = (s.radelem[elem_i] + s.radelem[elem_j]) * s.rcutfac; }
void orc_snap_idxb(void* h, int* triples /* [ncoeff][3] */) { Sna* s = static_cast<Sna*>(h); for(int k = 0; k < s->idxb_max; k++) { triples[3*k] = s->idxb[k].j1; triples[3*k+1] = s->idxb[k].j2; triples[3*k+2] = s->idxb[k].j; } }

// one neighbourhood: bispectrum B[ncoeff], optional dB/dr_j [n][ncoeff][3] (derivative wrt the NEIGHBOUR position), energy and dE/dr_j [n][3]
void orc_snap_atom(void* h, int n, const double* dx, const double* dy, const double* dz, const int* elem_j, int elem_i,
                   double* B, double* dB, double* energy, double* dedr)
{
  const Sna& s = *static_cast<Sna*>(h);
  thread_local Sna::Work w; thread_local const Sna* w_owner = nullptr;
  if( w_owner != &s || int(w.utot_r.size()) != s.idxu_max || int(w.z_r.size()) != s.idxz_max || int(w.blist.size()) != s.idxb_max ) { s.alloc(w); w_owner = &s; }
  std::vector<double> wj(n), rc(n);
  for(int i = 0; i < n; i++) { const int ej = elem_j ? elem_j[i] : 0; wj[i] = s.wjelem[ej]; rc[i] = (s.radelem[elem_i] + s.radelem[ej]) * s.rcutfac; }
  s.compute_ui(w, n, dx, dy, dz, wj.data(), rc.data());
  s.compute_zi(w); s.compute_bi(w);
  if( B ) for(int k = 0; k < s.ncoeff; k++) B[k] = w.blist[k];
  const double* beta = s.beta.data() + size_t(elem_i) * (s.ncoeff + 1);
  if( energy ) { double e = beta[0]; for(int k = 0; k < s.ncoeff; k++) e += beta[k + 1] * w.blist[k]; *energy = e; }
  if( dB || dedr ) s.compute_yi(w, beta + 1);
  for(int i = 0; i < n && (dB || dedr); i++)
  {
    s.compute_duidrj(w, dx[i], dy[i], dz[i], wj[i], rc[i]);
    if( dB ) s.compute_dbidrj(w, dB + size_t(i) * s.ncoeff * 3);
    if( dedr ) s.compute_deidrj(w, dedr + 3 * i);
  }
}

// debug/identity helper: 2*sum_half Re(conj(Utot).Y) for the neighbourhood (Euler: = 3 * sum_k beta_k B_k when bzeroflag = 0),
// the form in which the CUDA kernel obtains the energy without a separate compute_bi pass
double orc_snap_uy(void* h, int n, const double* dx, const double* dy, const double* dz, const int* elem_j, int elem_i)
{
  const Sna& s = *static_cast<Sna*>(h);
  Sna::Work w; s.alloc(w);
  std::vector<double> wj(n), rc(n);
  for(int i = 0; i < n; i++) { const int ej = elem_j ? elem_j[i] : 0; wj[i] = s.wjelem[ej]; rc[i] = (s.radelem[elem_i] + s.radelem[ej]) * s.rcutfac; }
  s.compute_ui(w, n, dx, dy, dz, wj.data(), rc.data());
  s.compute_yi(w, s.beta.data() + size_t(elem_i) * (s.ncoeff + 1) + 1);
  double sum = 0.0;
  for(int j = 0; j <= s.twojmax; j++)
  {
    int jju = s.idxu_block[j];
    for(int mb = 0; 2 * mb < j; mb++) for(int ma = 0; ma <= j; ma++) { sum += w.utot_r[jju] * w.y_r[jju] + w.utot_i[jju] * w.y_i[jju]; jju++; }
    if( j % 2 == 0 )
    {
      const int mb = j / 2;
      for(int ma = 0; ma < mb; ma++) { sum += w.utot_r[jju] * w.y_r[jju] + w.utot_i[jju] * w.y_i[jju]; jju++; }
      sum += 0.5 * (w.utot_r[jju] * w.y_r[jju] + w.utot_i[jju] * w.y_i[jju]);
    }
  }
  return 2.0 * sum;
}

} // extern "C"
