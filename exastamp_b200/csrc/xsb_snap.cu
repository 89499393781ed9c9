// xsb_snap.cu -- snap_force: SNAP bispectrum energy and forces.
//
// Reference operator: src/potential/snap/snap_force.cu:18-37 registers md::SnapForceGeneric of the un-vendored exaNBody
// (contribs/md/snap); its call sequence is the one visible in src/potential/snaplmp/snap_force_op.h:177-337
// (neighbour filter rsq < cutsq_ij && rsq > 1e-20, compute_ui, compute_yi(beta), per neighbour compute_duidrj +
// compute_deidrj, f_i += fij, f_j -= fij, virial -fij (x) rij on the centre, energy e0 + beta.B) with the LAMMPS SNA
// conventions (constructor arguments snaplmp.cpp:205-216).  Written from the published algorithm (Thompson et al. 2015;
// adjoint Y: Wood & Thompson 2018), FP64.
//
// B200 mapping: one CTA per central atom, 32 x (J+1) threads: lane = neighbour, warp = row mb of the Wigner matrices.
// The U recursion of row mb over the levels j = 2mb..2J is a chain that only needs the same row of level j-1, so a
// thread keeps its row in registers through all levels (fully unrolled); the only cross-thread dependency -- the birth
// of row mb at level 2mb from the mirrored row mb-1 -- goes through a small shared-memory mailbox.  Utot is the
// shuffle-reduction over the lanes (neighbours); Y is accumulated in shared memory over the idxz table; the second
// sweep recomputes the chains together with their three derivatives and contracts them with Y on the fly, so neither
// U_ij nor dU_ij is ever stored.  The energy comes from Euler's theorem for the trilinear B: sum_k beta_k B_k =
// (1/3) 2 sum_half Re(conj(Utot) Y) -- no separate Z/B pass.
// The production path is the split pipeline snap_u_kernel -> snap_y2_kernel -> snap_fr_kernel (reverse-mode force sweep,
// register-blocked compute_yi); the fused kernel below remains for XSB_SNAP_FUSED / A-B runs.
#include "xsb_ctx.h"
#include "xsb_traverse.cuh"
#include "xsb_tile.cuh"
#include <algorithm>
#include <cmath>
#include <type_traits>
#include <vector>

namespace xsb
{

constexpr int SNAP_NN_MAX = 64;      // in-range neighbours of one atom held in shared memory at a time (2 sweeps of 32); atoms with
                                     // more are processed in batches (Utot, dE/dr and the forces are sums over neighbours)
constexpr int SNAP_NN_TAB = 192;     // in-range neighbours per atom the Utot kernel can hand to the force kernel of the split pipeline

struct SnapZ { unsigned char j1, j2, j, ma1min, ma2max, na, mb1min, mb2max, nb, pad; unsigned short jju; int cgoff; };   // 16 B

// snap_y2_kernel tables (see the kernel)
struct SnapYTri { unsigned char j1, j2, s, P; int cgp; };         // table of the triple at cgp: D[ma2][m], row stride P = j + 1 rounded up to even
struct SnapYTask { unsigned char j, mb, ma0, W; unsigned short tri0, ntri; };

// per-launch constants in the arithmetic type of the kernels (double, or float for XSB_FLAG_MIXED)
template<class real>
struct SnapConstT
{
  real rootpq[10][10];
  int idxu_block[10];
  int twojmax, idxu_max, idxz_max, ncoeff, nelements, switchflag, bzeroflag;
  real rfac0, rmin0, rcutfac, wself;
  real radelem[8], wjelem[8];
  double beta0[8], bzero_e[8];    // bzero_e[elem] = sum_k beta_k bzero[j_k]; energies are accumulated in double
};
typedef SnapConstT<double> SnapConst;

template<class real> struct R2;
template<> struct R2<double> { typedef double2 type; };
template<> struct R2<float>  { typedef float2 type; };
template<class real> __device__ __forceinline__ typename R2<real>::type mk2(real a, real b) { typename R2<real>::type v; v.x = a; v.y = b; return v; }
__device__ __forceinline__ void xsincos(double a, double* s, double* c) { sincos(a, s, c); }
__device__ __forceinline__ void xsincos(float a, float* s, float* c) { sincosf(a, s, c); }
__device__ __forceinline__ double xrsqrt(double a) { return rsqrt(a); }
__device__ __forceinline__ float xrsqrt(float a) { return rsqrtf(a); }
__device__ __forceinline__ double xsqrt(double a) { return sqrt(a); }
__device__ __forceinline__ float xsqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ double xcos(double a) { return cos(a); }
__device__ __forceinline__ float xcos(float a) { return cosf(a); }
__device__ __forceinline__ double xsin(double a) { return sin(a); }
__device__ __forceinline__ float xsin(float a) { return sinf(a); }
// inside the kernel templates `real2` is the complex pair of the arithmetic type
#define real2 typename R2<real>::type

struct SnapDev
{
  bool set = false;
  SnapConst K{};
  DevBuf<SnapZ> idxz; DevBuf<double> cglist; DevBuf<double> betaz;   // betaz[elem][jjz]: beta_k with the multiplicity / (j1+1)/(j+1) factors of compute_yi
  DevBuf<int> err;
  DevBuf<unsigned long long> clk; bool clocks = false;
  DevBuf<SnapZ> zsort; DevBuf<double> betaz_sort; DevBuf<int4> ytask; int n_ytask = 0;   // snap_y_kernel work items
  DevBuf<double2> ubuf, ybuf;                                                          // chunk staging (AoSoA)
  DevBuf<double> nbtab; DevBuf<unsigned> nbcnt;                                        // in-range neighbours of the chunk's atoms (Utot kernel -> force kernel)
  SnapConstT<float> K32{};                                                               // the same constants rounded to float (XSB_FLAG_MIXED)
  DevBuf<float> cglist32, betaz32, betaz_sort32;
  DevBuf<SnapYTri> y2tri; DevBuf<SnapYTask> y2task; DevBuf<double> y2cg, y2beta; DevBuf<float> y2cg32, y2beta32; int n_y2tri = 0, n_y2task = 0, n_y2cg = 0;   // snap_y2_kernel
  double rcut_max = 0.0;
  bool overflowed = false;       // a call hit SNAP_NN_MAX since xsb_snap_overflow() was last read
};

// the SNAP state of a context hangs off the context itself (xsb_ctx::snap), like GhostState: no process-global table
static SnapDev* g_snap_of(xsb_ctx* ctx) { return static_cast<SnapDev*>(ctx->snap); }

template<class real> __device__ __forceinline__ real snap_sfac(const SnapConstT<real>& K, real r, real rcut)
{
  if( K.switchflag == 0 || r <= K.rmin0 ) return real(1.0);
  if( r > rcut ) return real(0.0);
  return real(0.5) * (xcos((r - K.rmin0) * real(M_PI) / (rcut - K.rmin0)) + real(1.0));
}
template<class real> __device__ __forceinline__ real snap_dsfac(const SnapConstT<real>& K, real r, real rcut)
{
  if( K.switchflag == 0 || r <= K.rmin0 || r > rcut ) return real(0.0);
  const real f = real(M_PI) / (rcut - K.rmin0);
  return -real(0.5) * xsin((r - K.rmin0) * f) * f;
}

// mailbox layout per neighbour: source row m (published at level 2m+1, 2m+2 elements) starts at m(m+1)
__device__ __forceinline__ int mbox_off(int m) { return m * (m + 1); }
// a lane's mailbox is n real2 words; the 32 mailboxes of a warp are laid out with an odd stride (in 16-byte words) so
// that the lanes of a quarter-warp hit 8 different bank groups: with the natural stride (a multiple of 128 bytes at
// 2J = 8) every mailbox access was an 8-way bank conflict (ncu: 590 M shared wavefronts for 149 M ideal)
#define SNAP_MBOX_STRIDE(n) (((n) | 1))

// Sum N per-lane values over the 32 lanes of a warp by recursive halving: at each step a lane keeps one half of its values
// and receives the partner's copy of that half, so after the five steps value number `off` (if len > 0) is complete in
// v[0] of exactly one lane.  N <= 32.
template<class real, int N, int S> struct WarpHalving
{
  static __device__ __forceinline__ void run(real* v, unsigned lane, int& off, int& len)
  {
    constexpr int H = (N + 1) / 2;
    const bool up = (lane & unsigned(S)) != 0u;
#   pragma unroll
    for(int i = 0; i < H; i++)
    {
      const real hi = (H + i < N) ? v[H + i] : real(0.0);
      const real keep = up ? hi : v[i], send = up ? v[i] : hi;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, S);
    }
    if( up ) { off += H; len = len > H ? len - H : 0; } else len = len < H ? len : H;
    WarpHalving<real, H, S / 2>::run(v, lane, off, len);
  }
};
template<class real, int N> struct WarpHalving<real, N, 0> { static __device__ __forceinline__ void run(real*, unsigned, int&, int&) { static_assert(N == 1, "at most 32 values"); } };

// One sweep over the levels for the thread's (neighbour, row mb).  DERIV = false: accumulate sfac*wj*u into utot (shuffle
// reduction over the lanes).  DERIV = true: carry du/dr_k, contract with Y -> dedr[3].
template<class real, int TJ, bool DERIV>
__device__ __forceinline__ void snap_sweep(const SnapConstT<real>& K, int mb, bool valid, real x, real y, real z, real wj, real rcut,
                                           real2* __restrict__ utot, const real2* __restrict__ ylist,
                                           real2* __restrict__ mbox /* this neighbour's mailbox: [ (TJ/2)(TJ/2+1) ] x (DERIV ? 4 : 1) */,
                                           real dedr[3])
{
  constexpr int NE = TJ + 1;
  constexpr int MB = (TJ / 2) * (TJ / 2 + 1);     // mailbox entries per field
  real ur[NE], ui[NE];
  real dur[DERIV ? NE : 1][3], dui[DERIV ? NE : 1][3];
  const real rsq = x * x + y * y + z * z, r = xsqrt(rsq);
  const real rscale0 = K.rfac0 * real(M_PI) / (rcut - K.rmin0), theta0 = (r - K.rmin0) * rscale0;
  real sn, cs; xsincos(theta0, &sn, &cs);
  const real z0 = r * cs / sn;
  const real r0inv = xrsqrt(rsq + z0 * z0);
  const real a_r = z0 * r0inv, a_i = -z * r0inv, b_r = y * r0inv, b_i = -x * r0inv;
  const real sfac = valid ? snap_sfac(K, r, rcut) * wj : real(0.0);
  real da_r[3], da_i[3], db_r[3], db_i[3], uvec[3], dsfac = real(0.0);
  if( DERIV )
  {
    const real rinv = real(1.0) / r;
    uvec[0] = x * rinv; uvec[1] = y * rinv; uvec[2] = z * rinv;
    const real dz0dr = z0 * rinv - (r * rscale0) * (rsq + z0 * z0) / rsq;
    const real dr0invdr = -r0inv * r0inv * r0inv * (r + z0 * dz0dr);
#   pragma unroll
    for(int k = 0; k < 3; k++)
    {
      const real dr0inv = dr0invdr * uvec[k], dz0 = dz0dr * uvec[k];
      da_r[k] = dz0 * r0inv + z0 * dr0inv; da_i[k] = -z * dr0inv;
      db_r[k] = y * dr0inv; db_i[k] = -x * dr0inv;
    }
    da_i[2] += -r0inv; db_i[0] += -r0inv; db_r[1] += r0inv;
    dsfac = valid ? snap_dsfac(K, r, rcut) * wj : real(0.0);
    dedr[0] = dedr[1] = dedr[2] = real(0.0);
  }
# pragma unroll
  for(int e = 0; e < NE; e++) { ur[e] = real(0.0); ui[e] = real(0.0); if( DERIV ) { for(int k = 0; k < 3; k++) { dur[e][k] = real(0.0); dui[e][k] = real(0.0); } } }
  if( mb == 0 ) ur[0] = real(1.0);        // level 0: u = 1

  // contribution of the thread's row at level j (called after the row has been advanced to level j)
  auto emit = [&](int j, auto jc)
  {
    constexpr int J = decltype(jc)::value;
    const int base = K.idxu_block[J] + (J + 1) * mb;
    if( !DERIV )
    {
      // sum of the row over the 32 neighbours: recursive halving leaves each of the 2(J+1) totals in one lane
      // (2(J+1) - 1 + 4 shuffles per row instead of 10(J+1))
      real v[2 * (J + 1)];
#     pragma unroll
      for(int ma = 0; ma <= J; ma++) { v[2 * ma] = sfac * ur[ma]; v[2 * ma + 1] = sfac * ui[ma]; }
      int off = 0, len = 2 * (J + 1);
      WarpHalving<real, 2 * (J + 1), 16>::run(v, threadIdx.x & 31u, off, len);
      if( len > 0 ) reinterpret_cast<real*>(utot + base)[off] += v[0];
    }
    else
    {
      const bool middle = 2 * mb == J;
#     pragma unroll
      for(int ma = 0; ma <= J; ma++)
      {
        real w = real(1.0);
        if( middle ) w = ma < mb ? real(1.0) : (ma == mb ? real(0.5) : real(0.0));
        const real2 Y = ylist[base + ma];
#       pragma unroll
        for(int k = 0; k < 3; k++)
        {
          const real fr = dsfac * ur[ma] * uvec[k] + sfac * dur[ma][k];
          const real fi = dsfac * ui[ma] * uvec[k] + sfac * dui[ma][k];
          dedr[k] += w * (fr * Y.x + fi * Y.y);
        }
      }
    }
    (void)j;
  };

  if( mb == 0 ) emit(0, std::integral_constant<int, 0>{});

  auto level = [&](auto jc)
  {
    constexpr int J = decltype(jc)::value;        // advance from level J-1 to level J
    if( 2 * mb <= J )
    {
      if( 2 * mb == J )
      {
        // birth of the row: level J-1 row mb is the mirror of row mb-1 (published to the mailbox at level J-1):
        // u[J-1][mb][ma] = (-1)^(mb-1+ma') conj(u[J-1][mb-1][ma']),  ma' = J-1-ma
        const int o = mbox_off(mb - 1);
#       pragma unroll
        for(int ma = 0; ma < J; ma++)
        {
          const int mp = J - 1 - ma;
          const real sgn = ((mb - 1 + mp) & 1) ? -real(1.0) : real(1.0);
          const real2 v = mbox[o + mp];
          ur[ma] = sgn * v.x; ui[ma] = -sgn * v.y;
          if( DERIV )
          {
#           pragma unroll
            for(int k = 0; k < 3; k++) { const real2 d = mbox[MB * (1 + k) + o + mp]; dur[ma][k] = sgn * d.x; dui[ma][k] = -sgn * d.y; }
          }
        }
      }
      real nr[J + 1], ni[J + 1], dnr[DERIV ? J + 1 : 1][3], dni[DERIV ? J + 1 : 1][3];
      nr[0] = real(0.0); ni[0] = real(0.0);
      if( DERIV ) { for(int k = 0; k < 3; k++) { dnr[0][k] = real(0.0); dni[0][k] = real(0.0); } }
#     pragma unroll
      for(int ma = 0; ma < J; ma++)
      {
        real q = K.rootpq[J - ma][J - mb];
        nr[ma] += q * (a_r * ur[ma] + a_i * ui[ma]);
        ni[ma] += q * (a_r * ui[ma] - a_i * ur[ma]);
        if( DERIV )
        {
#         pragma unroll
          for(int k = 0; k < 3; k++)
          {
            dnr[ma][k] += q * (da_r[k] * ur[ma] + da_i[k] * ui[ma] + a_r * dur[ma][k] + a_i * dui[ma][k]);
            dni[ma][k] += q * (da_r[k] * ui[ma] - da_i[k] * ur[ma] + a_r * dui[ma][k] - a_i * dur[ma][k]);
          }
        }
        q = K.rootpq[ma + 1][J - mb];
        nr[ma + 1] = -q * (b_r * ur[ma] + b_i * ui[ma]);
        ni[ma + 1] = -q * (b_r * ui[ma] - b_i * ur[ma]);
        if( DERIV )
        {
#         pragma unroll
          for(int k = 0; k < 3; k++)
          {
            dnr[ma + 1][k] = -q * (db_r[k] * ur[ma] + db_i[k] * ui[ma] + b_r * dur[ma][k] + b_i * dui[ma][k]);
            dni[ma + 1][k] = -q * (db_r[k] * ui[ma] - db_i[k] * ur[ma] + b_r * dui[ma][k] - b_i * dur[ma][k]);
          }
        }
      }
#     pragma unroll
      for(int ma = 0; ma <= J; ma++)
      {
        ur[ma] = nr[ma]; ui[ma] = ni[ma];
        if( DERIV ) { for(int k = 0; k < 3; k++) { dur[ma][k] = dnr[ma][k]; dui[ma][k] = dni[ma][k]; } }
      }
      emit(J, jc);
      if( J == 2 * mb + 1 && J < TJ )
      {
        // publish for the birth of row mb+1 at the next level
        const int o = mbox_off(mb);
#       pragma unroll
        for(int ma = 0; ma <= J; ma++)
        {
          mbox[o + ma] = mk2<real>(ur[ma], ui[ma]);
          if( DERIV ) { for(int k = 0; k < 3; k++) mbox[MB * (1 + k) + o + ma] = mk2<real>(dur[ma][k], dui[ma][k]); }
        }
      }
    }
    __syncthreads();
  };
  // levels 1..TJ, unrolled at compile time
  if constexpr ( TJ >= 1 ) level(std::integral_constant<int, 1>{});
  if constexpr ( TJ >= 2 ) level(std::integral_constant<int, 2>{});
  if constexpr ( TJ >= 3 ) level(std::integral_constant<int, 3>{});
  if constexpr ( TJ >= 4 ) level(std::integral_constant<int, 4>{});
  if constexpr ( TJ >= 5 ) level(std::integral_constant<int, 5>{});
  if constexpr ( TJ >= 6 ) level(std::integral_constant<int, 6>{});
  if constexpr ( TJ >= 7 ) level(std::integral_constant<int, 7>{});
  if constexpr ( TJ >= 8 ) level(std::integral_constant<int, 8>{});
}

// Force sweep for ONE Cartesian direction kd: the thread's (neighbour, row mb) chain of u and du/dr_kd through all levels,
// contracted with Y on the fly (forward mode, kept for A/B against snap_sweep_rev).  The three directions of a neighbour run in
// three different CTAs (snap_fd_kernel), which cuts the per-thread state from 4 rows to 2 (fits ~128 registers) and triples the warps an SM can hold; the price is
// that the u chain itself is carried three times.
template<class real, int TJ>
__device__ __forceinline__ real snap_sweep_dir(const SnapConstT<real>& K, int mb, int kd, bool valid, real x, real y, real z, real wj, real rcut,
                                                 const real2* __restrict__ ylist, real2* __restrict__ mbox /* [2][MB] of this (neighbour, kd) */)
{
  constexpr int NE = TJ + 1;
  constexpr int MB = (TJ / 2) * (TJ / 2 + 1);
  real ur[NE], ui[NE], dur[NE], dui[NE];
  const real rsq = x * x + y * y + z * z, r = xsqrt(rsq);
  const real rscale0 = K.rfac0 * real(M_PI) / (rcut - K.rmin0), theta0 = (r - K.rmin0) * rscale0;
  real sn, cs; xsincos(theta0, &sn, &cs);
  const real z0 = r * cs / sn;
  const real r0inv = xrsqrt(rsq + z0 * z0);
  const real a_r = z0 * r0inv, a_i = -z * r0inv, b_r = y * r0inv, b_i = -x * r0inv;
  const real sfac = valid ? snap_sfac(K, r, rcut) * wj : real(0.0);
  const real rinv = real(1.0) / r;
  const real uv = (kd == 0 ? x : (kd == 1 ? y : z)) * rinv;
  const real dz0dr = z0 * rinv - (r * rscale0) * (rsq + z0 * z0) / rsq;
  const real dr0invdr = -r0inv * r0inv * r0inv * (r + z0 * dz0dr);
  const real dr0inv = dr0invdr * uv, dz0 = dz0dr * uv;
  const real da_r = dz0 * r0inv + z0 * dr0inv, da_i = -z * dr0inv + (kd == 2 ? -r0inv : real(0.0));
  const real db_r = y * dr0inv + (kd == 1 ? r0inv : real(0.0)), db_i = -x * dr0inv + (kd == 0 ? -r0inv : real(0.0));
  const real dsfac = valid ? snap_dsfac(K, r, rcut) * wj : real(0.0);
  const real dsu = dsfac * uv;
  real dedr = real(0.0);
# pragma unroll
  for(int e = 0; e < NE; e++) { ur[e] = real(0.0); ui[e] = real(0.0); dur[e] = real(0.0); dui[e] = real(0.0); }
  if( mb == 0 ) ur[0] = real(1.0);

  auto emit = [&](auto jc)
  {
    constexpr int J = decltype(jc)::value;
    const int base = K.idxu_block[J] + (J + 1) * mb;
    const bool middle = 2 * mb == J;
#   pragma unroll
    for(int ma = 0; ma <= J; ma++)
    {
      real w = real(1.0);
      if( middle ) w = ma < mb ? real(1.0) : (ma == mb ? real(0.5) : real(0.0));
      const real2 Y = ylist[base + ma];
      const real fr = dsu * ur[ma] + sfac * dur[ma];
      const real fi = dsu * ui[ma] + sfac * dui[ma];
      dedr += w * (fr * Y.x + fi * Y.y);
    }
  };
  if( mb == 0 ) emit(std::integral_constant<int, 0>{});

  auto level = [&](auto jc)
  {
    constexpr int J = decltype(jc)::value;
    if( 2 * mb <= J )
    {
      if( 2 * mb == J )
      {
        const int o = mbox_off(mb - 1);
#       pragma unroll
        for(int ma = 0; ma < J; ma++)
        {
          const int mp = J - 1 - ma;
          const real sgn = ((mb - 1 + mp) & 1) ? -real(1.0) : real(1.0);
          const real2 v = mbox[o + mp], d = mbox[MB + o + mp];
          ur[ma] = sgn * v.x; ui[ma] = -sgn * v.y; dur[ma] = sgn * d.x; dui[ma] = -sgn * d.y;
        }
      }
      // in place, descending ma: new[ma] = qa a* old[ma] - qb b* old[ma-1]
#     pragma unroll
      for(int ma = J; ma >= 0; ma--)
      {
        real nr = real(0.0), ni = real(0.0), dnr = real(0.0), dni = real(0.0);
        if( ma < J )
        {
          const real q = K.rootpq[J - ma][J - mb];
          nr = q * (a_r * ur[ma] + a_i * ui[ma]);
          ni = q * (a_r * ui[ma] - a_i * ur[ma]);
          dnr = q * (da_r * ur[ma] + da_i * ui[ma] + a_r * dur[ma] + a_i * dui[ma]);
          dni = q * (da_r * ui[ma] - da_i * ur[ma] + a_r * dui[ma] - a_i * dur[ma]);
        }
        if( ma > 0 )
        {
          const real q = K.rootpq[ma][J - mb];
          nr -= q * (b_r * ur[ma - 1] + b_i * ui[ma - 1]);
          ni -= q * (b_r * ui[ma - 1] - b_i * ur[ma - 1]);
          dnr -= q * (db_r * ur[ma - 1] + db_i * ui[ma - 1] + b_r * dur[ma - 1] + b_i * dui[ma - 1]);
          dni -= q * (db_r * ui[ma - 1] - db_i * ur[ma - 1] + b_r * dui[ma - 1] - b_i * dur[ma - 1]);
        }
        ur[ma] = nr; ui[ma] = ni; dur[ma] = dnr; dui[ma] = dni;
      }
      emit(jc);
      if( J == 2 * mb + 1 && J < TJ )
      {
        const int o = mbox_off(mb);
#       pragma unroll
        for(int ma = 0; ma <= J; ma++) { mbox[o + ma] = mk2<real>(ur[ma], ui[ma]); mbox[MB + o + ma] = mk2<real>(dur[ma], dui[ma]); }
      }
    }
    __syncthreads();
  };
  if constexpr ( TJ >= 1 ) level(std::integral_constant<int, 1>{});
  if constexpr ( TJ >= 2 ) level(std::integral_constant<int, 2>{});
  if constexpr ( TJ >= 3 ) level(std::integral_constant<int, 3>{});
  if constexpr ( TJ >= 4 ) level(std::integral_constant<int, 4>{});
  if constexpr ( TJ >= 5 ) level(std::integral_constant<int, 5>{});
  if constexpr ( TJ >= 6 ) level(std::integral_constant<int, 6>{});
  if constexpr ( TJ >= 7 ) level(std::integral_constant<int, 7>{});
  if constexpr ( TJ >= 8 ) level(std::integral_constant<int, 8>{});
  return dedr;
}

// Reverse-mode force sweep: dE/dr of ALL three Cartesian directions from one forward chain + one backward (adjoint) chain.
// G(a, b) = sum_levels sum_half w Re(conj(Y) u) is a polynomial in the Cayley-Klein parameters a, b through the recursion
// u^J[mb][ma] = qa conj(a) u^{J-1}[mb][ma] - qb conj(b) u^{J-1}[mb][ma-1]; instead of carrying du/dr_k along (one chain per
// direction, snap_sweep_dir), the thread stores its row of every level in shared memory on the way up and walks back down
// with the adjoint row ubar^J = w Y^J + (what level J+1 handed down): ubar^{J-1}[m] = qa a ubar^J[m] - qb b ubar^J[m+1],
// abar += qa (ubar^J[m] . u^{J-1}[m]), bbar -= qb (ubar^J[m+1] . u^{J-1}[m]).  The birth of a row (mirror of row mb-1) is
// real-linear, its adjoint goes back through the same mailbox.  out = { G, dG/da_r, dG/da_i, dG/db_r, dG/db_i } of this row;
// the caller sums the rows and applies da/dr_k, db/dr_k, sfac and dsfac.  ~32 FP64 operations per element and neighbour
// for the three directions together, against 3 x 35 of the per-direction sweeps.
template<class real, int TJ>
__device__ __forceinline__ void snap_sweep_rev(const SnapConstT<real>& K, int mb, real x, real y, real z, real rcut,
                                               const real2* __restrict__ ylist, real2* __restrict__ mbox /* this neighbour's mailbox [MB] */,
                                               real2* __restrict__ hist /* this lane's column of the level history, slot stride 32 */, int hbase,
                                               real out[5])
{
  constexpr int NE = TJ + 1;
  real ur[NE], ui[NE], br[NE], bi[NE];
  const real rsq = x * x + y * y + z * z, r = xsqrt(rsq);
  const real rscale0 = K.rfac0 * real(M_PI) / (rcut - K.rmin0), theta0 = (r - K.rmin0) * rscale0;
  real sn, cs; xsincos(theta0, &sn, &cs);
  const real z0 = r * cs / sn;
  const real r0inv = xrsqrt(rsq + z0 * z0);
  const real a_r = z0 * r0inv, a_i = -z * r0inv, b_r = y * r0inv, b_i = -x * r0inv;
  real G = real(0.0), abr = real(0.0), abi = real(0.0), bbr = real(0.0), bbi = real(0.0);
# pragma unroll
  for(int e = 0; e < NE; e++) { ur[e] = real(0.0); ui[e] = real(0.0); br[e] = real(0.0); bi[e] = real(0.0); }
  if( mb == 0 ) { ur[0] = real(1.0); G = real(0.5) * ylist[0].x; }      // level 0: u = 1 (middle element of its level: w = 1/2)

  // row mb at level J-1 is the mirror of row mb-1 (published to the mailbox at level J-1)
  auto birth = [&](auto jc)
  {
    constexpr int J = decltype(jc)::value;
    const int o = mbox_off(mb - 1);
#   pragma unroll
    for(int ma = 0; ma < J; ma++)
    {
      const int mp = J - 1 - ma;
      const real sgn = ((mb - 1 + mp) & 1) ? -real(1.0) : real(1.0);
      const real2 v = mbox[o + mp];
      ur[ma] = sgn * v.x; ui[ma] = -sgn * v.y;
    }
  };
  // one step down: nbar = br/bi (adjoint of level J, J+1 elements), old = ur/ui (level J-1, J elements) -> br/bi = adjoint of
  // level J-1 (J elements), or its mirror into the mailbox when the row was born at this level
  auto down = [&](auto jc)
  {
    constexpr int J = decltype(jc)::value;
    real nr[J + 1], ni[J + 1];
#   pragma unroll
    for(int ma = 0; ma <= J; ma++) { nr[ma] = br[ma]; ni[ma] = bi[ma]; }
#   pragma unroll
    for(int m = 0; m < J; m++)
    {
      const real qa = K.rootpq[J - m][J - mb], qb = K.rootpq[m + 1][J - mb];
      const real pr = qa * nr[m], pi = qa * ni[m], sr = qb * nr[m + 1], si = qb * ni[m + 1];
      // a * p - b * s, every term a fused multiply-add on the running value
      br[m] = fma(b_i, si, fma(-b_r, sr, fma(-a_i, pi, a_r * pr)));
      bi[m] = fma(-b_i, sr, fma(-b_r, si, fma(a_i, pr, a_r * pi)));
      abr = fma(pi, ui[m], fma(pr, ur[m], abr)); abi = fma(-pi, ur[m], fma(pr, ui[m], abi));
      bbr = fma(-si, ui[m], fma(-sr, ur[m], bbr)); bbi = fma(si, ur[m], fma(-sr, ui[m], bbi));
    }
    if( 2 * mb == J )
    {
      const int o = mbox_off(mb - 1);
#     pragma unroll
      for(int ma = 0; ma < J; ma++)
      {
        const int mp = J - 1 - ma;
        const real sgn = ((mb - 1 + mp) & 1) ? -real(1.0) : real(1.0);
        mbox[o + mp] = mk2<real>(sgn * br[ma], -sgn * bi[ma]);
      }
    }
  };
  auto seed_w = [&](int J, int ma) -> real { return 2 * mb == J ? (ma < mb ? real(1.0) : (ma == mb ? real(0.5) : real(0.0))) : real(1.0); };

  auto up = [&](auto jc)
  {
    constexpr int J = decltype(jc)::value;        // advance from level J-1 to level J < TJ, remember level J-1
    if( 2 * mb <= J )
    {
      if( 2 * mb == J ) birth(jc);
      const int hs = hbase + J * (J - 1) / 2;
#     pragma unroll
      for(int ma = 0; ma < J; ma++) hist[(hs + ma) * 32] = mk2<real>(ur[ma], ui[ma]);
#     pragma unroll
      for(int ma = J; ma >= 0; ma--)
      {
        real n_r = real(0.0), n_i = real(0.0);
        if( ma < J ) { const real q = K.rootpq[J - ma][J - mb]; n_r = q * (a_r * ur[ma] + a_i * ui[ma]); n_i = q * (a_r * ui[ma] - a_i * ur[ma]); }
        if( ma > 0 ) { const real q = K.rootpq[ma][J - mb]; n_r -= q * (b_r * ur[ma - 1] + b_i * ui[ma - 1]); n_i -= q * (b_r * ui[ma - 1] - b_i * ur[ma - 1]); }
        ur[ma] = n_r; ui[ma] = n_i;
      }
      const int base = K.idxu_block[J] + (J + 1) * mb;
#     pragma unroll
      for(int ma = 0; ma <= J; ma++)
      {
        const real2 Y = ylist[base + ma];
        if( 2 * mb == J ) G = fma(seed_w(J, ma), fma(ui[ma], Y.y, ur[ma] * Y.x), G);
        else G = fma(ui[ma], Y.y, fma(ur[ma], Y.x, G));
      }
      if( J == 2 * mb + 1 )
      {
        const int o = mbox_off(mb);
#       pragma unroll
        for(int ma = 0; ma <= J; ma++) mbox[o + ma] = mk2<real>(ur[ma], ui[ma]);
      }
    }
    __syncthreads();
  };
  if constexpr ( TJ >= 2 ) up(std::integral_constant<int, 1>{});
  if constexpr ( TJ >= 3 ) up(std::integral_constant<int, 2>{});
  if constexpr ( TJ >= 4 ) up(std::integral_constant<int, 3>{});
  if constexpr ( TJ >= 5 ) up(std::integral_constant<int, 4>{});
  if constexpr ( TJ >= 6 ) up(std::integral_constant<int, 5>{});
  if constexpr ( TJ >= 7 ) up(std::integral_constant<int, 6>{});
  if constexpr ( TJ >= 8 ) up(std::integral_constant<int, 7>{});

  // top level: the level TJ row is only needed for G, so level TJ-1 stays in the registers and the way down starts here
  {
    constexpr int J = TJ;
    if( 2 * mb <= J )
    {
      if( 2 * mb == J ) birth(std::integral_constant<int, J>{});
      const int base = K.idxu_block[J] + (J + 1) * mb;
#     pragma unroll
      for(int ma = 0; ma <= J; ma++)
      {
        real n_r = real(0.0), n_i = real(0.0);
        if( ma < J ) { const real q = K.rootpq[J - ma][J - mb]; n_r = q * (a_r * ur[ma] + a_i * ui[ma]); n_i = q * (a_r * ui[ma] - a_i * ur[ma]); }
        if( ma > 0 ) { const real q = K.rootpq[ma][J - mb]; n_r -= q * (b_r * ur[ma - 1] + b_i * ui[ma - 1]); n_i -= q * (b_r * ui[ma - 1] - b_i * ur[ma - 1]); }
        const real2 Y = ylist[base + ma];
        const real w = seed_w(J, ma);
        G += w * (n_r * Y.x + n_i * Y.y);
        br[ma] = w * Y.x; bi[ma] = w * Y.y;
      }
      down(std::integral_constant<int, J>{});
    }
    __syncthreads();
  }
  auto dn = [&](auto jc)
  {
    constexpr int J = decltype(jc)::value;        // adjoint of level J -> level J-1
    if( 2 * mb <= J )
    {
      const int base = K.idxu_block[J] + (J + 1) * mb;
      const bool mail = J == 2 * mb + 1;          // row mb+1 was born from this level: its adjoint comes back mirrored
      const int o = mbox_off(mb);
#     pragma unroll
      for(int ma = 0; ma <= J; ma++)
      {
        const real2 Y = ylist[base + ma];
        if( 2 * mb == J ) { const real w = seed_w(J, ma); br[ma] = fma(w, Y.x, br[ma]); bi[ma] = fma(w, Y.y, bi[ma]); }
        else { br[ma] += Y.x; bi[ma] += Y.y; }
        if( mail ) { const real2 m = mbox[o + ma]; br[ma] += m.x; bi[ma] += m.y; }
      }
      const int hs = hbase + J * (J - 1) / 2;
#     pragma unroll
      for(int ma = 0; ma < J; ma++) { const real2 v = hist[(hs + ma) * 32]; ur[ma] = v.x; ui[ma] = v.y; }
      down(jc);
    }
    __syncthreads();
  };
  if constexpr ( TJ >= 8 ) dn(std::integral_constant<int, 7>{});
  if constexpr ( TJ >= 7 ) dn(std::integral_constant<int, 6>{});
  if constexpr ( TJ >= 6 ) dn(std::integral_constant<int, 5>{});
  if constexpr ( TJ >= 5 ) dn(std::integral_constant<int, 4>{});
  if constexpr ( TJ >= 4 ) dn(std::integral_constant<int, 3>{});
  if constexpr ( TJ >= 3 ) dn(std::integral_constant<int, 2>{});
  if constexpr ( TJ >= 2 ) dn(std::integral_constant<int, 1>{});
  out[0] = G; out[1] = abr; out[2] = abi; out[3] = bbr; out[4] = bbi;
}

template<class real>
struct SnapArgsT
{
  typedef real real_t;
  const double* __restrict__ rx; const double* __restrict__ ry; const double* __restrict__ rz; const unsigned char* __restrict__ type;
  const unsigned long long* __restrict__ nbh_off; const unsigned* __restrict__ nbh_idx;
  const unsigned* __restrict__ atoms; unsigned n_atoms;
  const SnapZ* __restrict__ idxz; const real* __restrict__ cglist; const real* __restrict__ betaz;
  double *fx, *fy, *fz, *ep, *vir; int* err;
  unsigned long long* clk;     // optional per-phase cycle counters (tools/snap_bench.py --clocks), nullptr in production
  // split pipeline (snap_u -> snap_y -> snap_f): Utot and Y of the atoms of one chunk, AoSoA [atom / 32][jju][atom % 32]
  real2* ubuf; real2* ybuf; unsigned base;      // base = first central atom (position in the launch's atom list) of the chunk
  // in-range neighbours found by the Utot kernel, per chunk slot: 6 rows of SNAP_NN_MAX doubles (dx, dy, dz, wj, rc, index bits)
  double* nbtab; unsigned* nbcnt;
  const SnapZ* __restrict__ zsort; const real* __restrict__ betaz_sort; const int4* __restrict__ ytask; int n_ytask;
  // snap_y2_kernel: triples sorted by j, padded Clebsch-Gordan rows, beta per (element, triple), work items
  const SnapYTri* __restrict__ y2tri; const real* __restrict__ y2cg; const real* __restrict__ y2beta; const SnapYTask* __restrict__ y2task; int n_y2tri, n_y2task, n_y2cg;
};

// PHASE 0: fused (everything in one CTA, kept for reference / small runs); PHASE 1: Utot only -> A.ubuf; PHASE 3: reads Utot
// and Y of its atom back from A.ubuf / A.ybuf, then energy + force sweep.
template<class real, int TJ, bool XFORM, int PHASE>
__global__ void __launch_bounds__(32 * (TJ / 2 + 1)) snap_force_kernel(const SnapArgsT<real> A, const XForm X, const SnapConstT<real> K)
{
  constexpr int NR = TJ / 2 + 1, NT = 32 * NR;
  constexpr int MB = (TJ / 2) * (TJ / 2 + 1);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  real2* utot = reinterpret_cast<real2*>(smem_raw);                 // [idxu_max]
  real2* ylist = utot + K.idxu_max;                                   // [idxu_max]
  real2* mbox = ylist + K.idxu_max;                                   // [32][4 * MB]
  real* nb_x = reinterpret_cast<real*>(mbox + 32 * SNAP_MBOX_STRIDE(4 * (MB ? MB : 1)));   // [SNAP_NN_MAX] x 5
  real* nb_y = nb_x + SNAP_NN_MAX; real* nb_z = nb_y + SNAP_NN_MAX; real* nb_w = nb_z + SNAP_NN_MAX; real* nb_rc = nb_w + SNAP_NN_MAX;
  unsigned* nb_g = reinterpret_cast<unsigned*>(nb_rc + SNAP_NN_MAX);    // [SNAP_NN_MAX]
  real* red = reinterpret_cast<real*>(nb_g + SNAP_NN_MAX);          // [NR][32][3] + scratch
  __shared__ unsigned s_nn;
  const unsigned tid = threadIdx.x, lane = tid & 31u; const int mb = int(tid >> 5);
  const unsigned slot = A.base + blockIdx.x;
  const unsigned ai = A.atoms ? A.atoms[slot] : slot;
  const size_t soa = (size_t(blockIdx.x >> 5) * K.idxu_max) * 32 + (blockIdx.x & 31u);    // + jju * 32
  const double xa = A.rx[ai], ya = A.ry[ai], za = A.rz[ai];
  const int ei = A.type ? A.type[ai] : 0;

  long long tc = A.clk ? clock64() : 0;
  auto mark = [&](int ph) { if( A.clk && tid == 0 ) { const long long t = clock64(); atomicAdd(A.clk + ph, (unsigned long long)(t - tc)); tc = t; } };
  // ---- Utot = wself on the diagonal, Y = 0
  if( PHASE != 3 )
  {
    for(int k = tid; k < K.idxu_max; k += NT) { utot[k] = mk2<real>(real(0.0), real(0.0)); ylist[k] = mk2<real>(real(0.0), real(0.0)); }
    __syncthreads();
    for(int j = int(tid); j <= TJ; j += NT) for(int ma = 0; ma <= j; ma++) utot[K.idxu_block[j] + (j + 1) * ma + ma].x = K.wself;
  }
  else
  {
    for(int k = tid; k < K.idxu_max; k += NT) { utot[k] = A.ubuf[soa + size_t(k) * 32]; ylist[k] = A.ybuf[soa + size_t(k) * 32]; }
  }

  // ---- neighbour filter (warp 0, ballot compaction keeps the list order): rsq < cutsq_ij && rsq > 1e-20.  One call loads the
  // next batch of in-range neighbours (whole 32-entry chunks of the list, at most SNAP_NN_MAX) starting at list position `from`;
  // s_e = where the next batch starts (>= e1: the list is exhausted)
  __shared__ unsigned long long s_e;
  const unsigned long long e0 = A.nbh_off[ai], e1 = A.nbh_off[ai + 1];
  auto filter = [&](unsigned long long from)
  {
    if( mb == 0 )
    {
      unsigned nn = 0;
      unsigned long long e = from;
      for(; e < e1; e += 32)
      {
        const unsigned long long ee = e + lane;
        bool in = false; double dx = 0, dy = 0, dz = 0, rc = 0, wj = 0; unsigned g = 0;
        if( ee < e1 )
        {
          g = A.nbh_idx[ee];
          dx = A.rx[g] - xa; dy = A.ry[g] - ya; dz = A.rz[g] - za;
          apply_xform<XFORM>(X, dx, dy, dz);
          const int ej = A.type ? A.type[g] : 0;
          rc = (double(K.radelem[ei]) + double(K.radelem[ej])) * double(K.rcutfac); wj = K.wjelem[ej];
          const double d2 = dx * dx + dy * dy + dz * dz;
          in = d2 < rc * rc && d2 > 1e-20;
        }
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if( nn + __popc(m) > SNAP_NN_MAX ) break;                 // this chunk opens the next batch (same decision on every lane)
        const unsigned slot = nn + __popc(m & ((1u << lane) - 1u));
        if( in ) { nb_x[slot] = real(dx); nb_y[slot] = real(dy); nb_z[slot] = real(dz); nb_w[slot] = real(wj); nb_rc[slot] = real(rc); nb_g[slot] = g; }
        nn += __popc(m);
      }
      if( lane == 0 ) { s_nn = nn; s_e = e; }
    }
    __syncthreads();
  };
  // ---- sweep 1: Utot, batch by batch
  real dummy[3];
  if( PHASE != 3 )
  {
    unsigned tot = 0;
    for(unsigned long long from = e0; ; )
    {
      filter(from);
      const unsigned nn = s_nn; const unsigned long long nxt = s_e;
      if( PHASE == 1 && A.nbtab && mb == 0 )
      {
        double* t = A.nbtab + size_t(blockIdx.x) * 6 * SNAP_NN_TAB;
        for(unsigned i = lane; i < nn && tot + i < SNAP_NN_TAB; i += 32)
        {
          const unsigned o = tot + i;
          t[o] = nb_x[i]; t[SNAP_NN_TAB + o] = nb_y[i]; t[2 * SNAP_NN_TAB + o] = nb_z[i]; t[3 * SNAP_NN_TAB + o] = nb_w[i]; t[4 * SNAP_NN_TAB + o] = nb_rc[i];
          t[5 * SNAP_NN_TAB + o] = __longlong_as_double((long long)nb_g[i]);
        }
      }
      for(unsigned b0 = 0; b0 < nn; b0 += 32)
      {
        const unsigned n = b0 + lane; const bool valid = n < nn;
        const real x = valid ? nb_x[n] : real(1.0), y = valid ? nb_y[n] : real(0.0), z = valid ? nb_z[n] : real(0.0), w = valid ? nb_w[n] : real(0.0), rc = valid ? nb_rc[n] : real(4.0);
        snap_sweep<real, TJ, false>(K, mb, valid, x, y, z, w, rc, utot, ylist, mbox + lane * SNAP_MBOX_STRIDE(4 * (MB ? MB : 1)), dummy);
      }
      __syncthreads();                      // the next filter call overwrites the neighbour arrays
      tot += nn;
      if( nxt >= e1 ) break;
      from = nxt;
    }
    if( PHASE == 1 && A.nbtab && tid == 0 )
    {
      if( tot > SNAP_NN_TAB ) { atomicExch(A.err, 1); tot = SNAP_NN_TAB; }
      A.nbcnt[blockIdx.x] = tot;
    }
    if( PHASE == 0 ) mark(0);
  }
  __syncthreads();
  mark(1);
  // right half by inversion symmetry u[j-mb][j-ma] = (-1)^(mb+ma) conj(u[mb][ma]); middle row: second half from the first
  if( PHASE != 3 )
  for(int j = 1; j <= TJ; j++)
  {
    const int jb = K.idxu_block[j], half = (j + 1) * ((j + 1) / 2) + ((j % 2 == 0) ? j / 2 : 0);   // elements strictly before the mirror centre
    for(int k = int(tid); k < half; k += NT)
    {
      const int mbb = k / (j + 1), ma = k % (j + 1);
      const real sgn = ((mbb + ma) & 1) ? -real(1.0) : real(1.0);
      const real2 v = utot[jb + k];
      utot[jb + (j + 1) * (j - mbb) + (j - ma)] = mk2<real>(sgn * v.x, -sgn * v.y);
    }
  }
  __syncthreads();

  mark(2);
  if( PHASE == 1 )
  {
    for(int k = tid; k < K.idxu_max; k += NT) A.ubuf[soa + size_t(k) * 32] = utot[k];
    return;
  }
  // ---- Y = sum over idxz of betaj * Z  (compute_yi), shared-memory accumulation
  const real* betaz = A.betaz + size_t(ei) * K.idxz_max;
  if( PHASE == 0 )
  for(int jjz = int(tid); jjz < K.idxz_max; jjz += NT)
  {
    const SnapZ q = A.idxz[jjz];
    const real* cg = A.cglist + q.cgoff;
    real zr = real(0.0), zi = real(0.0);
    int jju1 = K.idxu_block[q.j1] + (q.j1 + 1) * q.mb1min, jju2 = K.idxu_block[q.j2] + (q.j2 + 1) * q.mb2max, icgb = q.mb1min * (q.j2 + 1) + q.mb2max;
    for(int ib = 0; ib < q.nb; ib++)
    {
      real sr = real(0.0), si = real(0.0);
      int ma1 = q.ma1min, ma2 = q.ma2max, icga = q.ma1min * (q.j2 + 1) + q.ma2max;
      for(int ia = 0; ia < q.na; ia++)
      {
        const real2 u1 = utot[jju1 + ma1], u2 = utot[jju2 + ma2];
        const real c = __ldg(cg + icga);
        sr += c * (u1.x * u2.x - u1.y * u2.y);
        si += c * (u1.x * u2.y + u1.y * u2.x);
        ma1++; ma2--; icga += q.j2;
      }
      const real c = __ldg(cg + icgb);
      zr += c * sr; zi += c * si;
      jju1 += q.j1 + 1; jju2 -= q.j2 + 1; icgb += q.j2;
    }
    const real bj = __ldg(betaz + jjz);
    atomicAdd(&ylist[q.jju].x, bj * zr); atomicAdd(&ylist[q.jju].y, bj * zi);
  }
  __syncthreads();

  mark(3);
  // ---- energy: e0 + (1/3) 2 sum_half Re(conj(Utot) Y) - sum_k beta_k bzero
  if( A.ep )
  {
    real s = real(0.0);
    for(int j = 0; j <= TJ; j++)
    {
      const int jb = K.idxu_block[j], cnt = (j + 1) * ((j + 1) / 2) + ((j % 2 == 0) ? j / 2 + 1 : 0);
      for(int k = int(tid); k < cnt; k += NT)
      {
        const real w = (j % 2 == 0 && k == cnt - 1) ? real(0.5) : real(1.0);
        s += w * (utot[jb + k].x * ylist[jb + k].x + utot[jb + k].y * ylist[jb + k].y);
      }
    }
#   pragma unroll
    for(int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if( lane == 0 ) red[mb] = s;
    __syncthreads();
    if( tid == 0 ) { real t = real(0.0); for(int w = 0; w < NR; w++) t += red[w]; A.ep[ai] += K.beta0[ei] + (2.0 / 3.0) * double(t) - K.bzero_e[ei]; }
    __syncthreads();
  }

  mark(4);
  // ---- sweep 2: dU/dr contracted with Y -> fij ; f_i += fij, f_j -= fij, virial -fij (x) rij on the centre
  real fix = real(0.0), fiy = real(0.0), fiz = real(0.0), v[9];
# pragma unroll
  for(int k = 0; k < 9; k++) v[k] = real(0.0);
  for(unsigned long long from = e0; ; )
  {
  filter(from);
  const unsigned nn = s_nn; const unsigned long long nxt = s_e;
  for(unsigned b0 = 0; b0 < nn; b0 += 32)
  {
    const unsigned n = b0 + lane; const bool valid = n < nn;
    const real x = valid ? nb_x[n] : real(1.0), y = valid ? nb_y[n] : real(0.0), z = valid ? nb_z[n] : real(0.0), w = valid ? nb_w[n] : real(0.0), rc = valid ? nb_rc[n] : real(4.0);
    real dedr[3];
    snap_sweep<real, TJ, true>(K, mb, valid, x, y, z, w, rc, utot, ylist, mbox + lane * SNAP_MBOX_STRIDE(4 * (MB ? MB : 1)), dedr);
    red[(mb * 32 + lane) * 3 + 0] = dedr[0]; red[(mb * 32 + lane) * 3 + 1] = dedr[1]; red[(mb * 32 + lane) * 3 + 2] = dedr[2];
    __syncthreads();
    if( mb == 0 && valid )
    {
      real f[3] = { real(0.0), real(0.0), real(0.0) };
      for(int w2 = 0; w2 < NR; w2++) for(int k = 0; k < 3; k++) f[k] += red[(w2 * 32 + lane) * 3 + k];
      for(int k = 0; k < 3; k++) f[k] *= real(2.0);
      fix += f[0]; fiy += f[1]; fiz += f[2];
      const unsigned g = nb_g[n];
      atomicAdd(A.fx + g, -double(f[0])); atomicAdd(A.fy + g, -double(f[1])); atomicAdd(A.fz + g, -double(f[2]));
      if( A.vir )
      {
        v[0] -= f[0] * x; v[1] -= f[0] * y; v[2] -= f[0] * z;
        v[3] -= f[1] * x; v[4] -= f[1] * y; v[5] -= f[1] * z;
        v[6] -= f[2] * x; v[7] -= f[2] * y; v[8] -= f[2] * z;
      }
    }
    __syncthreads();
  }
  if( nxt >= e1 ) break;
  from = nxt;
  }
  if( mb == 0 )
  {
#   pragma unroll
    for(int o = 16; o > 0; o >>= 1) { fix += __shfl_xor_sync(0xffffffffu, fix, o); fiy += __shfl_xor_sync(0xffffffffu, fiy, o); fiz += __shfl_xor_sync(0xffffffffu, fiz, o); }
    if( A.vir )
    {
#     pragma unroll
      for(int k = 0; k < 9; k++) { for(int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o); }
    }
    if( lane == 0 )
    {
      atomicAdd(A.fx + ai, double(fix)); atomicAdd(A.fy + ai, double(fiy)); atomicAdd(A.fz + ai, double(fiz));
      if( A.vir ) { double* p = A.vir + 9ull * ai; for(int k = 0; k < 9; k++) p[k] += double(v[k]); }
    }
  }
  mark(5);
}

// ---- Utot kernel of the split pipeline: CTA per atom, J/2+1 warps = rows, lane = neighbour -------------------------------
// The same steps as PHASE 1 of snap_force_kernel (filter, Utot sweep, mirror, AoSoA store, neighbour table for the force
// kernel) without the fused kernel's Y / derivative state: a fifth of its shared memory and under 80 registers, so five and
// more CTAs share an SM and hide each other's serial prologue (list walk: four dependent global loads) and level barriers.
template<class real, int TJ, bool XFORM>
__global__ void __launch_bounds__(32 * (TJ / 2 + 1), sizeof(real) == 4 ? 8 : 5) snap_u_kernel(const SnapArgsT<real> A, const XForm X, const SnapConstT<real> K)
{
  constexpr int NR = TJ / 2 + 1, NT = 32 * NR;
  constexpr int MB = (TJ / 2) * (TJ / 2 + 1), MBS = MB ? MB : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  real2* utot = reinterpret_cast<real2*>(smem_raw);                 // [idxu_max]
  real2* mbox = utot + K.idxu_max;                                    // [32][MBS] (odd stride)
  real* nb_x = reinterpret_cast<real*>(mbox + 32 * SNAP_MBOX_STRIDE(MBS));        // [SNAP_NN_MAX] x 5
  real* nb_y = nb_x + SNAP_NN_MAX; real* nb_z = nb_y + SNAP_NN_MAX; real* nb_w = nb_z + SNAP_NN_MAX; real* nb_rc = nb_w + SNAP_NN_MAX;
  unsigned* nb_g = reinterpret_cast<unsigned*>(nb_rc + SNAP_NN_MAX);    // [SNAP_NN_MAX]
  __shared__ unsigned s_nn;
  __shared__ unsigned long long s_e;
  const unsigned tid = threadIdx.x, lane = tid & 31u; const int mb = int(tid >> 5);
  const unsigned slot = A.base + blockIdx.x;
  const unsigned ai = A.atoms ? A.atoms[slot] : slot;
  const size_t soa = (size_t(blockIdx.x >> 5) * K.idxu_max) * 32 + (blockIdx.x & 31u);    // + jju * 32
  const double xa = A.rx[ai], ya = A.ry[ai], za = A.rz[ai];
  const int ei = A.type ? A.type[ai] : 0;
  // Utot = wself on the diagonals, zero elsewhere (one pass, no barrier of its own: the filter's barrier below orders it)
  for(int k = tid; k < K.idxu_max; k += NT)
  {
    int j = 0;
    while( j < TJ && k >= K.idxu_block[j + 1] ) ++j;
    const int r = k - K.idxu_block[j];
    utot[k] = mk2<real>(r % (j + 2) == 0 ? K.wself : real(0.0), real(0.0));      // r = (j+1) ma + ma
  }
  const unsigned long long e0 = A.nbh_off[ai], e1 = A.nbh_off[ai + 1];
  double* const tab = A.nbtab + size_t(blockIdx.x) * 6 * SNAP_NN_TAB;
  unsigned tot = 0;
  for(unsigned long long from = e0; ; )
  {
    // neighbour filter (warp 0, ballot compaction keeps the list order): rsq < cutsq_ij && rsq > 1e-20; whole 32-entry chunks
    // of the list, at most SNAP_NN_MAX in-range neighbours per batch
    if( mb == 0 )
    {
      unsigned nn = 0;
      unsigned long long e = from;
      for(; e < e1; e += 32)
      {
        const unsigned long long ee = e + lane;
        bool in = false; double dx = 0, dy = 0, dz = 0, rc = 0, wj = 0; unsigned g = 0;
        if( ee < e1 )
        {
          g = A.nbh_idx[ee];
          dx = A.rx[g] - xa; dy = A.ry[g] - ya; dz = A.rz[g] - za;
          apply_xform<XFORM>(X, dx, dy, dz);
          const int ej = A.type ? A.type[g] : 0;
          rc = (double(K.radelem[ei]) + double(K.radelem[ej])) * double(K.rcutfac); wj = K.wjelem[ej];
          const double d2 = dx * dx + dy * dy + dz * dz;
          in = d2 < rc * rc && d2 > 1e-20;
        }
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if( nn + __popc(m) > SNAP_NN_MAX ) break;                 // this chunk opens the next batch (same decision on every lane)
        const unsigned sl = nn + __popc(m & ((1u << lane) - 1u));
        if( in )
        {
          nb_x[sl] = real(dx); nb_y[sl] = real(dy); nb_z[sl] = real(dz); nb_w[sl] = real(wj); nb_rc[sl] = real(rc); nb_g[sl] = g;
          const unsigned o = tot + sl;
          if( o < SNAP_NN_TAB )
          {
            tab[o] = double(real(dx)); tab[SNAP_NN_TAB + o] = double(real(dy)); tab[2 * SNAP_NN_TAB + o] = double(real(dz));
            tab[3 * SNAP_NN_TAB + o] = double(real(wj)); tab[4 * SNAP_NN_TAB + o] = double(real(rc)); tab[5 * SNAP_NN_TAB + o] = __longlong_as_double((long long)g);
          }
        }
        nn += __popc(m);
      }
      if( lane == 0 ) { s_nn = nn; s_e = e; }
    }
    __syncthreads();
    const unsigned nn = s_nn; const unsigned long long nxt = s_e;
    real dummy[3];
    for(unsigned b0 = 0; b0 < nn; b0 += 32)
    {
      const unsigned n = b0 + lane; const bool valid = n < nn;
      const real x = valid ? nb_x[n] : real(1.0), y = valid ? nb_y[n] : real(0.0), z = valid ? nb_z[n] : real(0.0), w = valid ? nb_w[n] : real(0.0), rc = valid ? nb_rc[n] : real(4.0);
      snap_sweep<real, TJ, false>(K, mb, valid, x, y, z, w, rc, utot, nullptr, mbox + lane * SNAP_MBOX_STRIDE(MBS), dummy);
    }
    __syncthreads();                      // the next filter call overwrites the neighbour arrays
    tot += nn;
    if( nxt >= e1 ) break;
    from = nxt;
  }
  if( tid == 0 )
  {
    if( tot > SNAP_NN_TAB ) { atomicExch(A.err, 1); tot = SNAP_NN_TAB; }
    A.nbcnt[blockIdx.x] = tot;
  }
  // right half by inversion symmetry u[j-mb][j-ma] = (-1)^(mb+ma) conj(u[mb][ma]); middle row: second half from the first
  for(int j = 1; j <= TJ; j++)
  {
    const int jb = K.idxu_block[j], half = (j + 1) * ((j + 1) / 2) + ((j % 2 == 0) ? j / 2 : 0);   // elements strictly before the mirror centre
    for(int k = int(tid); k < half; k += NT)
    {
      const int mbb = k / (j + 1), ma = k % (j + 1);
      const real sgn = ((mbb + ma) & 1) ? -real(1.0) : real(1.0);
      const real2 v = utot[jb + k];
      utot[jb + (j + 1) * (j - mbb) + (j - ma)] = mk2<real>(sgn * v.x, -sgn * v.y);
    }
  }
  __syncthreads();
  for(int k = tid; k < K.idxu_max; k += NT) A.ubuf[soa + size_t(k) * 32] = utot[k];
}

// ---- force kernel, one CTA per (atom, Cartesian direction): J/2+1 warps = rows, lane = neighbour ------------------------
// Forward-mode sweep (snap_sweep_dir); the three directions of an atom are three independent CTAs of 160 threads (2J = 8):
// three of them fit the register file of an SM, so the serial prologue of one (neighbour filter: four dependent global
// loads; Y fetch) overlaps the sweeps of the others, and a barrier only joins 5 warps instead of real(15.)  The neighbour
// filter and the Y fetch are repeated per direction (cheap against the sweep); the energy is computed by direction real(0.)
template<class real, int TJ, bool XFORM>
__global__ void __launch_bounds__(32 * (TJ / 2 + 1), sizeof(real) == 4 ? 5 : 3) snap_fd_kernel(const SnapArgsT<real> A, const XForm X, const SnapConstT<real> K)
{
  constexpr int NR = TJ / 2 + 1, NT = 32 * NR;
  constexpr int MB = (TJ / 2) * (TJ / 2 + 1), MBS = 2 * (MB ? MB : 1);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  real2* ylist = reinterpret_cast<real2*>(smem_raw);                // [idxu_max]
  real2* mbox = ylist + K.idxu_max;                                   // [32][MBS] (odd stride)
  real* nb_x = reinterpret_cast<real*>(mbox + 32 * SNAP_MBOX_STRIDE(MBS));        // [SNAP_NN_MAX] x 5
  real* nb_y = nb_x + SNAP_NN_MAX; real* nb_z = nb_y + SNAP_NN_MAX; real* nb_w = nb_z + SNAP_NN_MAX; real* nb_rc = nb_w + SNAP_NN_MAX;
  unsigned* nb_g = reinterpret_cast<unsigned*>(nb_rc + SNAP_NN_MAX);    // [SNAP_NN_MAX]
  real* red = reinterpret_cast<real*>(nb_g + SNAP_NN_MAX);          // [NR][32]
  __shared__ unsigned s_nn;
  const unsigned tid = threadIdx.x, lane = tid & 31u; const int mb = int(tid >> 5);
  const unsigned at = blockIdx.x / 3u; const int kd = int(blockIdx.x % 3u);
  const unsigned slot = A.base + at;
  const unsigned ai = A.atoms ? A.atoms[slot] : slot;
  const size_t soa = (size_t(at >> 5) * K.idxu_max) * 32 + (at & 31u);
  const double xa = A.rx[ai], ya = A.ry[ai], za = A.rz[ai];
  const int ei = A.type ? A.type[ai] : 0;
  // the Utot kernel of this chunk left the in-range neighbours of every slot (at most SNAP_NN_TAB): coalesced reads instead of
  // the list walk, SNAP_NN_MAX at a time
  const unsigned ntot = A.nbcnt[at];
  const double* t = A.nbtab + size_t(at) * 6 * SNAP_NN_TAB;
  auto load_batch = [&](unsigned base)
  {
    const unsigned nn = min(unsigned(SNAP_NN_MAX), ntot - base);
    for(unsigned i = lane; i < nn; i += 32)
    {
      const unsigned o = base + i;
      nb_x[i] = real(t[o]); nb_y[i] = real(t[SNAP_NN_TAB + o]); nb_z[i] = real(t[2 * SNAP_NN_TAB + o]); nb_w[i] = real(t[3 * SNAP_NN_TAB + o]); nb_rc[i] = real(t[4 * SNAP_NN_TAB + o]);
      nb_g[i] = unsigned(__double_as_longlong(t[5 * SNAP_NN_TAB + o]));
    }
    if( lane == 0 ) s_nn = nn;
  };
  if( mb == 0 ) load_batch(0u);
  else
  {
    // the other rows fetch Y while row 0 reads the neighbour table
    for(int k = int(tid) - 32; k < K.idxu_max; k += NT - 32) ylist[k] = A.ybuf[soa + size_t(k) * 32];
  }
  __syncthreads();
  // ---- energy (direction 0 only): e0 + (1/3) 2 sum_half Re(conj(Utot) Y) - sum_k beta_k bzero ; Utot straight from global
  if( A.ep && kd == 0 )
  {
    real sE = real(0.0);
    for(int j = 0; j <= TJ; j++)
    {
      const int jb = K.idxu_block[j], cnt = (j + 1) * ((j + 1) / 2) + ((j % 2 == 0) ? j / 2 + 1 : 0);
      for(int k = int(tid); k < cnt; k += NT)
      {
        const real w = (j % 2 == 0 && k == cnt - 1) ? real(0.5) : real(1.0);
        const real2 u = A.ubuf[soa + size_t(jb + k) * 32];
        sE += w * (u.x * ylist[jb + k].x + u.y * ylist[jb + k].y);
      }
    }
#   pragma unroll
    for(int o = 16; o > 0; o >>= 1) sE += __shfl_xor_sync(0xffffffffu, sE, o);
    if( lane == 0 ) red[mb] = sE;
    __syncthreads();
    if( tid == 0 ) { real t = real(0.0); for(int w = 0; w < NR; w++) t += red[w]; A.ep[ai] += K.beta0[ei] + (2.0 / 3.0) * double(t) - K.bzero_e[ei]; }
    __syncthreads();
  }
  double* const fout = kd == 0 ? A.fx : (kd == 1 ? A.fy : A.fz);
  real fi = real(0.0), v0 = real(0.0), v1 = real(0.0), v2 = real(0.0);
  for(unsigned base = 0; base < ntot; base += SNAP_NN_MAX)
  {
  if( base ) { if( mb == 0 ) load_batch(base); __syncthreads(); }
  const unsigned nn = s_nn;
  for(unsigned b0 = 0; b0 < nn; b0 += 32)
  {
    const unsigned n = b0 + lane; const bool valid = n < nn;
    const real x = valid ? nb_x[n] : real(1.0), y = valid ? nb_y[n] : real(0.0), z = valid ? nb_z[n] : real(0.0), w = valid ? nb_w[n] : real(0.0), rc = valid ? nb_rc[n] : real(4.0);
    const real d = snap_sweep_dir<real, TJ>(K, mb, kd, valid, x, y, z, w, rc, ylist, mbox + lane * SNAP_MBOX_STRIDE(MBS));
    red[mb * 32 + lane] = d;
    __syncthreads();
    if( mb == 0 && valid )
    {
      real f = real(0.0);
      for(int w2 = 0; w2 < NR; w2++) f += red[w2 * 32 + lane];
      f *= real(2.0);
      fi += f;
      atomicAdd(fout + nb_g[n], -double(f));
      if( A.vir ) { v0 -= f * x; v1 -= f * y; v2 -= f * z; }
    }
    __syncthreads();
  }
  }
  if( mb == 0 )
  {
#   pragma unroll
    for(int o = 16; o > 0; o >>= 1) fi += __shfl_xor_sync(0xffffffffu, fi, o);
    if( A.vir )
    {
#     pragma unroll
      for(int o = 16; o > 0; o >>= 1) { v0 += __shfl_xor_sync(0xffffffffu, v0, o); v1 += __shfl_xor_sync(0xffffffffu, v1, o); v2 += __shfl_xor_sync(0xffffffffu, v2, o); }
    }
    if( lane == 0 )
    {
      atomicAdd(fout + ai, double(fi));
      if( A.vir ) { double* p = A.vir + 9ull * ai + 3 * kd; atomicAdd(p, double(v0)); atomicAdd(p + 1, double(v1)); atomicAdd(p + 2, double(v2)); }
    }
  }
}

// ---- force kernel, reverse mode: one CTA per atom, J/2+1 warps = rows, lane = neighbour --------------------------------
// Forward chain with the level history in shared memory ([slot][lane], conflict-free 16-byte columns), adjoint chain back
// down (snap_sweep_rev), then warp 0 turns { G, dG/da, dG/db } of each neighbour into the three force components.
template<int TJ> struct SnapHist
{
  // history slots of row mb: levels max(1, 2mb) .. TJ-1, level J holds J elements
  static constexpr int S(int J) { return J * (J - 1) / 2; }
  static constexpr int j0(int mb) { return 2 * mb > 1 ? 2 * mb : 1; }
  static constexpr int rowsize(int mb) { return j0(mb) < TJ ? S(TJ) - S(j0(mb)) : 0; }
  static constexpr int total() { int t = 0; for(int m = 0; 2 * m <= TJ; m++) t += rowsize(m); return t; }
};

template<class real, int TJ, bool XFORM>
__global__ void __launch_bounds__(32 * (TJ / 2 + 1), sizeof(real) == 4 ? 5 : 3) snap_fr_kernel(const SnapArgsT<real> A, const XForm X, const SnapConstT<real> K)
{
  constexpr int NR = TJ / 2 + 1, NT = 32 * NR;
  constexpr int MB = (TJ / 2) * (TJ / 2 + 1), MBS = MB ? MB : 1;
  constexpr int HN = SnapHist<TJ>::total() ? SnapHist<TJ>::total() : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  real2* ylist = reinterpret_cast<real2*>(smem_raw);                // [idxu_max]
  real2* hist = ylist + K.idxu_max;                                   // [HN][32]
  real2* mbox = hist + HN * 32;                                       // [32][MBS] (odd stride)
  real* nb_x = reinterpret_cast<real*>(mbox + 32 * SNAP_MBOX_STRIDE(MBS));        // [SNAP_NN_MAX] x 5
  real* nb_y = nb_x + SNAP_NN_MAX; real* nb_z = nb_y + SNAP_NN_MAX; real* nb_w = nb_z + SNAP_NN_MAX; real* nb_rc = nb_w + SNAP_NN_MAX;
  unsigned* nb_g = reinterpret_cast<unsigned*>(nb_rc + SNAP_NN_MAX);    // [SNAP_NN_MAX]
  real* red = reinterpret_cast<real*>(nb_g + SNAP_NN_MAX);          // [NR][5][32]
  __shared__ unsigned s_nn;
  const unsigned tid = threadIdx.x, lane = tid & 31u; const int mb = int(tid >> 5);
  const unsigned at = blockIdx.x;
  const unsigned slot = A.base + at;
  const unsigned ai = A.atoms ? A.atoms[slot] : slot;
  const size_t soa = (size_t(at >> 5) * K.idxu_max) * 32 + (at & 31u);
  const int ei = A.type ? A.type[ai] : 0;
  int hbase = 0;
  for(int m = 0; m < mb; m++) hbase += SnapHist<TJ>::rowsize(m);
  hbase -= SnapHist<TJ>::S(SnapHist<TJ>::j0(mb));
  // the Utot kernel of this chunk left the in-range neighbours of every slot (at most SNAP_NN_TAB), SNAP_NN_MAX at a time
  const unsigned ntot = A.nbcnt[at];
  const double* t = A.nbtab + size_t(at) * 6 * SNAP_NN_TAB;
  auto load_batch = [&](unsigned base)
  {
    const unsigned nn = min(unsigned(SNAP_NN_MAX), ntot - base);
    for(unsigned i = lane; i < nn; i += 32)
    {
      const unsigned o = base + i;
      nb_x[i] = real(t[o]); nb_y[i] = real(t[SNAP_NN_TAB + o]); nb_z[i] = real(t[2 * SNAP_NN_TAB + o]); nb_w[i] = real(t[3 * SNAP_NN_TAB + o]); nb_rc[i] = real(t[4 * SNAP_NN_TAB + o]);
      nb_g[i] = unsigned(__double_as_longlong(t[5 * SNAP_NN_TAB + o]));
    }
    if( lane == 0 ) s_nn = nn;
  };
  if( mb == 0 ) load_batch(0u);
  if( NR == 1 || mb != 0 )
  {
    // the other rows fetch Y while row 0 reads the neighbour table
    const int first = NR == 1 ? int(tid) : int(tid) - 32, step = NR == 1 ? NT : NT - 32;
    for(int k = first; k < K.idxu_max; k += step) ylist[k] = A.ybuf[soa + size_t(k) * 32];
  }
  __syncthreads();
  // ---- energy: e0 + (1/3) 2 sum_half Re(conj(Utot) Y) - sum_k beta_k bzero ; Utot straight from global
  if( A.ep )
  {
    real sE = real(0.0);
    for(int j = 0; j <= TJ; j++)
    {
      const int jb = K.idxu_block[j], cnt = (j + 1) * ((j + 1) / 2) + ((j % 2 == 0) ? j / 2 + 1 : 0);
      for(int k = int(tid); k < cnt; k += NT)
      {
        const real w = (j % 2 == 0 && k == cnt - 1) ? real(0.5) : real(1.0);
        const real2 u = A.ubuf[soa + size_t(jb + k) * 32];
        sE += w * (u.x * ylist[jb + k].x + u.y * ylist[jb + k].y);
      }
    }
#   pragma unroll
    for(int o = 16; o > 0; o >>= 1) sE += __shfl_xor_sync(0xffffffffu, sE, o);
    if( lane == 0 ) red[mb] = sE;
    __syncthreads();
    if( tid == 0 ) { real s = real(0.0); for(int w = 0; w < NR; w++) s += red[w]; A.ep[ai] += K.beta0[ei] + (2.0 / 3.0) * double(s) - K.bzero_e[ei]; }
    __syncthreads();
  }
  real fix = real(0.0), fiy = real(0.0), fiz = real(0.0), v[9];
# pragma unroll
  for(int k = 0; k < 9; k++) v[k] = real(0.0);
  for(unsigned base = 0; base < ntot; base += SNAP_NN_MAX)
  {
  if( base ) { if( mb == 0 ) load_batch(base); __syncthreads(); }
  const unsigned nn = s_nn;
  for(unsigned b0 = 0; b0 < nn; b0 += 32)
  {
    const unsigned n = b0 + lane; const bool valid = n < nn;
    const real x = valid ? nb_x[n] : real(1.0), y = valid ? nb_y[n] : real(0.0), z = valid ? nb_z[n] : real(0.0), rc = valid ? nb_rc[n] : real(4.0);
    real o5[5];
    snap_sweep_rev<real, TJ>(K, mb, x, y, z, rc, ylist, mbox + lane * SNAP_MBOX_STRIDE(MBS), hist + lane, hbase, o5);
#   pragma unroll
    for(int k = 0; k < 5; k++) red[(mb * 5 + k) * 32 + lane] = o5[k];
    __syncthreads();
    if( mb == 0 && valid )
    {
      real s5[5] = { real(0.0), real(0.0), real(0.0), real(0.0), real(0.0) };
      for(int w2 = 0; w2 < NR; w2++) for(int k = 0; k < 5; k++) s5[k] += red[(w2 * 5 + k) * 32 + lane];
      // chain rule a, b -> r (same expressions as the forward-mode sweeps)
      const real wj = nb_w[n];
      const real rsq = x * x + y * y + z * z, r = xsqrt(rsq);
      const real rscale0 = K.rfac0 * real(M_PI) / (rc - K.rmin0), theta0 = (r - K.rmin0) * rscale0;
      real sn, cs; xsincos(theta0, &sn, &cs);
      const real z0 = r * cs / sn;
      const real r0inv = xrsqrt(rsq + z0 * z0);
      const real rinv = real(1.0) / r;
      const real dz0dr = z0 * rinv - (r * rscale0) * (rsq + z0 * z0) / rsq;
      const real dr0invdr = -r0inv * r0inv * r0inv * (r + z0 * dz0dr);
      const real sfac = snap_sfac(K, r, rc) * wj, dsfac = snap_dsfac(K, r, rc) * wj;
      const real uvec[3] = { x * rinv, y * rinv, z * rinv };
      real f[3];
#     pragma unroll
      for(int k = 0; k < 3; k++)
      {
        const real dr0inv = dr0invdr * uvec[k], dz0 = dz0dr * uvec[k];
        const real da_r = dz0 * r0inv + z0 * dr0inv, da_i = -z * dr0inv + (k == 2 ? -r0inv : real(0.0));
        const real db_r = y * dr0inv + (k == 1 ? r0inv : real(0.0)), db_i = -x * dr0inv + (k == 0 ? -r0inv : real(0.0));
        f[k] = real(2.0) * (dsfac * uvec[k] * s5[0] + sfac * (s5[1] * da_r + s5[2] * da_i + s5[3] * db_r + s5[4] * db_i));
      }
      fix += f[0]; fiy += f[1]; fiz += f[2];
      const unsigned g = nb_g[n];
      atomicAdd(A.fx + g, -double(f[0])); atomicAdd(A.fy + g, -double(f[1])); atomicAdd(A.fz + g, -double(f[2]));
      if( A.vir )
      {
        v[0] -= f[0] * x; v[1] -= f[0] * y; v[2] -= f[0] * z;
        v[3] -= f[1] * x; v[4] -= f[1] * y; v[5] -= f[1] * z;
        v[6] -= f[2] * x; v[7] -= f[2] * y; v[8] -= f[2] * z;
      }
    }
    __syncthreads();
  }
  }
  if( mb == 0 )
  {
#   pragma unroll
    for(int o = 16; o > 0; o >>= 1) { fix += __shfl_xor_sync(0xffffffffu, fix, o); fiy += __shfl_xor_sync(0xffffffffu, fiy, o); fiz += __shfl_xor_sync(0xffffffffu, fiz, o); }
    if( A.vir )
    {
#     pragma unroll
      for(int k = 0; k < 9; k++) { for(int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o); }
    }
    if( lane == 0 )
    {
      atomicAdd(A.fx + ai, double(fix)); atomicAdd(A.fy + ai, double(fiy)); atomicAdd(A.fz + ai, double(fiz));
      if( A.vir ) { double* p = A.vir + 9ull * ai; for(int k = 0; k < 9; k++) p[k] += double(v[k]); }
    }
  }
}

// ---- compute_yi for 32 atoms at a time: lane = atom -------------------------------------------------------------------
// One CTA owns one AoSoA block of 32 central atoms: their Utot (idxu_max x 32 complex doubles, 146 KB at 2J = 8) arrives
// in shared memory with ONE TMA bulk copy, every warp-wide U access is then 32 consecutive 16-byte words (conflict-free),
// and all index / Clebsch-Gordan bookkeeping is warp-uniform.  Work items are the distinct Y elements (jju) with the
// list of idxz entries that feed them (sorted by cost, handed out through a shared counter), so each Y element has one
// writer: no atomics, no zero-fill.
template<class real, int TJ>
__global__ void __launch_bounds__(512, 1) snap_y_kernel(const SnapArgsT<real> A, const SnapConstT<real> K)
{
  extern __shared__ __align__(128) unsigned char ysm[];
  real2* U = reinterpret_cast<real2*>(ysm);                  // [idxu_max][32]
  __shared__ __align__(8) unsigned long long bar;
  __shared__ int next_task;
  const unsigned tid = threadIdx.x, lane = tid & 31u;
  const size_t blk = size_t(blockIdx.x) * K.idxu_max * 32;
  if( tid == 0 ) { mbar_init(&bar, 1); mbar_fence_init(); next_task = 0; }
  __syncthreads();
  if( tid == 0 )
  {
    const unsigned bytes = unsigned(K.idxu_max) * 32u * unsigned(sizeof(real2));
    mbar_arrive_expect_tx(&bar, bytes);
    bulk_g2s(U, A.ubuf + blk, bytes, &bar);
  }
  // element of this lane's atom (selects the beta row); lanes past the end of the atom list compute on stale data and
  // write into the padding of ybuf
  const unsigned slot = A.base + blockIdx.x * 32u + lane;
  int ei = 0;
  if( A.type && slot < A.n_atoms ) ei = A.type[A.atoms ? A.atoms[slot] : slot];
  const real* betaz = A.betaz_sort + size_t(ei) * K.idxz_max;
  mbar_wait(&bar, 0);
  for(;;)
  {
    int t = 0;
    if( lane == 0 ) t = atomicAdd(&next_task, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if( t >= A.n_ytask ) break;
    const int4 task = A.ytask[t];                      // x = jju, y = first entry in zsort, z = entry count
    real yr = real(0.0), yi = real(0.0);
    for(int e = task.y; e < task.y + task.z; e++)
    {
      const SnapZ q = A.zsort[e];
      const real* cg = A.cglist + q.cgoff;
      real zr = real(0.0), zi = real(0.0);
      int jju1 = K.idxu_block[q.j1] + (q.j1 + 1) * q.mb1min, jju2 = K.idxu_block[q.j2] + (q.j2 + 1) * q.mb2max, icgb = q.mb1min * (q.j2 + 1) + q.mb2max;
      for(int ib = 0; ib < q.nb; ib++)
      {
        real sr = real(0.0), si = real(0.0);
        const real2* u1p = U + size_t(jju1 + q.ma1min) * 32 + lane;
        const real2* u2p = U + size_t(jju2 + q.ma2max) * 32 + lane;
        int icga = q.ma1min * (q.j2 + 1) + q.ma2max;
        for(int ia = 0; ia < q.na; ia++)
        {
          const real2 u1 = *u1p, u2 = *u2p;
          const real c = __ldg(cg + icga);
          sr += c * (u1.x * u2.x - u1.y * u2.y);
          si += c * (u1.x * u2.y + u1.y * u2.x);
          u1p += 32; u2p -= 32; icga += q.j2;
        }
        const real c = __ldg(cg + icgb);
        zr += c * sr; zi += c * si;
        jju1 += q.j1 + 1; jju2 -= q.j2 + 1; icgb += q.j2;
      }
      const real bj = __ldg(betaz + e);
      yr += bj * zr; yi += bj * zi;
    }
    A.ybuf[blk + size_t(task.x) * 32 + lane] = mk2<real>(yr, yi);
  }
}

// ---- compute_yi, register-blocked (snap_y2_kernel) ------------------------------------------------------------------------
// Same CTA shape as snap_y_kernel (32 atoms, Utot block in shared memory, lane = atom), different work item: a block of W <= 5
// consecutive ma of ONE row (j, mb) of Y.  For a triple (j1, j2 <= j1, j) and a row pair (mb1, mb2 = mb + s - mb1) the ma sum is
// a correlation  z[ma] = sum_ma2 c[ma1][ma2] u1[ma1] u2[ma2],  ma1 = ma + s - ma2,  s = (j1 + j2 - j) / 2.  The loop runs over
// ma2 (the SHORT row: every step is useful for almost every output), one u2 element per step; the W outputs need the W
// consecutive u1 elements ma1 = t .. t + W - 1 with t falling by one per step: a sliding window in registers, rotated by
// unrolling the loop W times.  2 shared-memory loads per W complex multiply-adds instead of 2 per 1 (the first version ran at
// 85 % of the shared-memory pipe and 26 % of the FP64 pipe).  The Clebsch-Gordan coefficients are re-tabulated per triple as
// D[ma2][m] = c[m + s - ma2][ma2] (zero where ma1 falls outside 0..j1), so the W coefficients of a step are consecutive and a
// window position that sticks out of the u1 row needs no test (index clamped, coefficient zero).  The tables of all triples
// (27 KB at 2J = 8) sit in shared memory behind the Utot block; blocks start at even ma, so a step fetches its W coefficients
// as W/2 broadcast 16-byte loads.  (Constant memory was measured slower: 16 warps on different triples thrash its 2 KB L1.)
template<class real, int W>
__device__ __forceinline__ void snap_y_block(const SnapConstT<real>& K, const SnapYTask task, const SnapYTri* __restrict__ tri, const real* __restrict__ cgtab /* shared memory */,
                                             const real* __restrict__ betat, const real2* __restrict__ U, unsigned lane, real2* __restrict__ yout)
{
  real yr[W], yi[W];
# pragma unroll
  for(int d = 0; d < W; d++) { yr[d] = real(0.0); yi[d] = real(0.0); }
  const int mb = task.mb, ma0 = task.ma0;
  for(int it = 0; it < int(task.ntri); it++)
  {
    const SnapYTri q = tri[task.tri0 + it];
    const int j1 = q.j1, j2 = q.j2, s = q.s, P = q.P;
    const int mb1min = max(0, mb + s - j2), nb = min(j1, mb + s) - mb1min + 1;
    const int lo2 = max(0, ma0 + s - j1), n_it = min(j2, ma0 + W - 1 + s) - lo2 + 1;
    const int t0 = ma0 + s - lo2;                               // ma1 of output 0 at ma2 = lo2  (<= j1)
    const real bj = __ldg(betat + task.tri0 + it);
    for(int ib = 0; ib < nb; ib++)
    {
      const int r1 = mb1min + ib, r2 = mb + s - r1;
      const real cb = bj * cgtab[q.cgp + r2 * P + mb];      // c[r1][r2]
      const real2* u1row = U + size_t(K.idxu_block[j1] + r1 * (j1 + 1)) * 32 + lane;
      const real2* u2p = U + size_t(K.idxu_block[j2] + r2 * (j2 + 1) + lo2) * 32 + lane;
      int t = t0, cp = q.cgp + lo2 * P + ma0;                    // coefficient of output d at this step: D[cp + d]; next ma2: cp += P
      real2 w[W];
#     pragma unroll
      for(int d = 0; d < W; d++) w[d] = u1row[size_t(min(t + d, j1)) * 32];
      real sr[W], si[W];
#     pragma unroll
      for(int d = 0; d < W; d++) { sr[d] = real(0.0); si[d] = real(0.0); }
      for(int k = 0; k < n_it; k += W)
      {
#       pragma unroll
        for(int p = 0; p < W; p++)
        {
          if( k + p < n_it )
          {
            const real2 u2 = *u2p; u2p += 32;
            real c[W + 1];
#           pragma unroll
            for(int d = 0; d < W; d += 2) { const real2 cc = *reinterpret_cast<const real2*>(cgtab + cp + d); c[d] = cc.x; c[d + 1] = cc.y; }
#           pragma unroll
            for(int d = 0; d < W; d++)
            {
              const real2 u1 = w[(d - p + W) % W];
              sr[d] += c[d] * (u1.x * u2.x - u1.y * u2.y);
              si[d] += c[d] * (u1.x * u2.y + u1.y * u2.x);
            }
            cp += P; t -= 1;
            w[(W - 1 - p) % W] = u1row[size_t(max(t, 0)) * 32];
          }
        }
      }
#     pragma unroll
      for(int d = 0; d < W; d++) { yr[d] += cb * sr[d]; yi[d] += cb * si[d]; }
    }
  }
  const int jju = K.idxu_block[task.j] + (task.j + 1) * mb + ma0;
# pragma unroll
  for(int d = 0; d < W; d++) yout[size_t(jju + d) * 32] = mk2<real>(yr[d], yi[d]);
}

template<class real>
__global__ void __launch_bounds__(512, 1) snap_y2_kernel(const SnapArgsT<real> A, const SnapConstT<real> K)
{
  extern __shared__ __align__(128) unsigned char ysm[];
  real2* U = reinterpret_cast<real2*>(ysm);                  // [idxu_max][32]
  real* CGs = reinterpret_cast<real*>(U + size_t(K.idxu_max) * 32);     // [n_y2cg] coefficient tables of all triples
  __shared__ __align__(8) unsigned long long bar;
  __shared__ int next_task;
  const unsigned tid = threadIdx.x, lane = tid & 31u;
  const size_t blk = size_t(blockIdx.x) * K.idxu_max * 32;
  if( tid == 0 ) { mbar_init(&bar, 1); mbar_fence_init(); next_task = 0; }
  __syncthreads();
  if( tid == 0 )
  {
    const unsigned bytes = unsigned(K.idxu_max) * 32u * unsigned(sizeof(real2)), cgbytes = unsigned(A.n_y2cg) * unsigned(sizeof(real));
    mbar_arrive_expect_tx(&bar, bytes + cgbytes);
    bulk_g2s(U, A.ubuf + blk, bytes, &bar);
    bulk_g2s(CGs, A.y2cg, cgbytes, &bar);
  }
  // the second half of a middle row (2 mb = j, ma > mb) is never used with a non-zero weight, but the force sweep multiplies
  // it by that zero: it must be finite
  for(int j = 2; j <= K.twojmax; j += 2)
    for(int k = int(tid); k < (j / 2) * 32; k += 512)
      A.ybuf[blk + size_t(K.idxu_block[j] + (j + 1) * (j / 2) + j / 2 + 1 + (k >> 5)) * 32 + (k & 31)] = mk2<real>(real(0.0), real(0.0));
  // element of this lane's atom (selects the beta row); lanes past the end of the atom list compute on stale data and
  // write into the padding of ybuf
  const unsigned slot = A.base + blockIdx.x * 32u + lane;
  int ei = 0;
  if( A.type && slot < A.n_atoms ) ei = A.type[A.atoms ? A.atoms[slot] : slot];
  const real* betat = A.y2beta + size_t(ei) * A.n_y2tri;
  real2* yout = A.ybuf + blk + lane;
  mbar_wait(&bar, 0);
  for(;;)
  {
    int t = 0;
    if( lane == 0 ) t = atomicAdd(&next_task, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if( t >= A.n_y2task ) break;
    const SnapYTask task = A.y2task[t];
    switch( task.W )
    {
      case 1: snap_y_block<real, 1>(K, task, A.y2tri, CGs, betat, U, lane, yout); break;
      case 2: snap_y_block<real, 2>(K, task, A.y2tri, CGs, betat, U, lane, yout); break;
      case 3: snap_y_block<real, 3>(K, task, A.y2tri, CGs, betat, U, lane, yout); break;
      case 4: snap_y_block<real, 4>(K, task, A.y2tri, CGs, betat, U, lane, yout); break;
      default: snap_y_block<real, 5>(K, task, A.y2tri, CGs, betat, U, lane, yout); break;
    }
  }
}

// ---- host: index tables of the published algorithm ------------------------------------------------------------
static double fact(int n) { double f = 1.0; for(int i = 2; i <= n; i++) f *= i; return f; }

struct SnapTables
{
  int twojmax, jdim, idxu_max = 0, ncoeff = 0;
  std::vector<int> idxu_block, cg_block, b_block;
  std::vector<SnapZ> idxz; std::vector<double> cglist; std::vector<int> b_j;     // b_j[k] = j of component k
  int b3(int a, int b, int c) const { return (a * jdim + b) * jdim + c; }
  explicit SnapTables(int tj) : twojmax(tj), jdim(tj + 1)
  {
    idxu_block.assign(jdim, 0); cg_block.assign(size_t(jdim) * jdim * jdim, -1); b_block = cg_block;
    for(int j = 0; j <= tj; j++) { idxu_block[j] = idxu_max; idxu_max += (j + 1) * (j + 1); }
    // triples j2 <= j1, |j1-j2| <= j <= min(2J, j1+j2), same parity
    auto triples = [&](auto f) { for(int j1 = 0; j1 <= tj; j1++) for(int j2 = 0; j2 <= j1; j2++) for(int j = j1 - j2; j <= std::min(tj, j1 + j2); j += 2) f(j1, j2, j); };
    triples([&](int j1, int j2, int j)
    {
      cg_block[b3(j1, j2, j)] = int(cglist.size());
      for(int m1 = 0; m1 <= j1; m1++) for(int m2 = 0; m2 <= j2; m2++)
      {
        const int aa2 = 2 * m1 - j1, bb2 = 2 * m2 - j2, m = (aa2 + bb2 + j) / 2;
        if( m < 0 || m > j ) { cglist.push_back(0.0); continue; }
        double sum = 0.0;
        const int zlo = std::max(0, std::max(-(j - j2 + aa2) / 2, -(j - j1 - bb2) / 2)), zhi = std::min((j1 + j2 - j) / 2, std::min((j1 - aa2) / 2, (j2 + bb2) / 2));
        for(int z = zlo; z <= zhi; z++)
          sum += ((z & 1) ? -1.0 : 1.0) / (fact(z) * fact((j1 + j2 - j) / 2 - z) * fact((j1 - aa2) / 2 - z) * fact((j2 + bb2) / 2 - z) * fact((j - j2 + aa2) / 2 + z) * fact((j - j1 - bb2) / 2 + z));
        const int cc2 = 2 * m - j;
        const double dcg = std::sqrt(fact((j1 + j2 - j) / 2) * fact((j1 - j2 + j) / 2) * fact((-j1 + j2 + j) / 2) / fact((j1 + j2 + j) / 2 + 1));
        const double sf = std::sqrt(fact((j1 + aa2) / 2) * fact((j1 - aa2) / 2) * fact((j2 + bb2) / 2) * fact((j2 - bb2) / 2) * fact((j + cc2) / 2) * fact((j - cc2) / 2) * (j + 1));
        cglist.push_back(sum * dcg * sf);
      }
      if( j >= j1 ) { b_block[b3(j1, j2, j)] = ncoeff++; b_j.push_back(j); }
    });
    triples([&](int j1, int j2, int j)
    {
      for(int mb = 0; 2 * mb <= j; mb++) for(int ma = 0; ma <= j; ma++)
      {
        SnapZ z{}; z.j1 = (unsigned char)j1; z.j2 = (unsigned char)j2; z.j = (unsigned char)j;
        const int ma1min = std::max(0, (2 * ma - j - j2 + j1) / 2), mb1min = std::max(0, (2 * mb - j - j2 + j1) / 2);
        z.ma1min = (unsigned char)ma1min; z.ma2max = (unsigned char)((2 * ma - j - (2 * ma1min - j1) + j2) / 2);
        z.na = (unsigned char)(std::min(j1, (2 * ma - j + j2 + j1) / 2) - ma1min + 1);
        z.mb1min = (unsigned char)mb1min; z.mb2max = (unsigned char)((2 * mb - j - (2 * mb1min - j1) + j2) / 2);
        z.nb = (unsigned char)(std::min(j1, (2 * mb - j + j2 + j1) / 2) - mb1min + 1);
        z.jju = (unsigned short)(idxu_block[j] + (j + 1) * mb + ma);
        z.cgoff = cg_block[b3(j1, j2, j)];
        idxz.push_back(z);
      }
    });
  }
  // beta_k with the multiplicity and (j1+1)/(j+1) factors with which component k enters Y through z(j1,j2,j)
  double betaj(const SnapZ& z, const double* beta) const
  {
    const int j1 = z.j1, j2 = z.j2, j = z.j;
    if( j >= j1 ) { const double b = beta[b_block[b3(j1, j2, j)]]; return j1 == j ? (j2 == j ? 3.0 * b : 2.0 * b) : b; }
    if( j >= j2 ) { const double b = beta[b_block[b3(j, j2, j1)]]; return (j2 == j ? 2.0 * b : b) * (j1 + 1) / (j + 1.0); }
    return beta[b_block[b3(j2, j, j1)]] * (j1 + 1) / (j + 1.0);
  }
};

} // namespace xsb

using namespace xsb;

void xsb_snap_release(xsb_ctx* ctx)
{
  SnapDev* sd = g_snap_of(ctx);
  if( !sd ) return;
  struct { SnapDev* second; } itv{ sd }; auto* it = &itv;
  it->second->cglist32.release(); it->second->betaz32.release(); it->second->betaz_sort32.release(); it->second->idxz.release(); it->second->cglist.release(); it->second->betaz.release(); it->second->err.release();
  it->second->y2tri.release(); it->second->y2task.release(); it->second->y2cg.release(); it->second->y2beta.release(); it->second->y2cg32.release(); it->second->y2beta32.release();
  it->second->zsort.release(); it->second->betaz_sort.release(); it->second->ytask.release(); it->second->ubuf.release(); it->second->ybuf.release(); it->second->nbtab.release(); it->second->nbcnt.release(); it->second->clk.release();
  delete it->second; ctx->snap = nullptr;
}

constexpr unsigned SNAP_CHUNK = 65536;     // central atoms per pass of the split pipeline (Utot + Y staging: 2 x 285 x 16 B per atom)

template<class real, int TJ>
static int snap_launch(xsb_ctx* ctx, SnapDev* S, SnapArgsT<real> A, const SnapConstT<real>& KK)
{
  constexpr int NR = TJ / 2 + 1, NT = 32 * NR, MB = (TJ / 2) * (TJ / 2 + 1);
  const size_t smem = size_t(2 * S->K.idxu_max) * sizeof(typename R2<real>::type) + size_t(32) * SNAP_MBOX_STRIDE(4 * (MB ? MB : 1)) * sizeof(typename R2<real>::type) + SNAP_NN_MAX * (5 * sizeof(real) + sizeof(unsigned))
                    + size_t(NR) * 32 * 3 * sizeof(real) + 64;
  const XForm X = make_xform(ctx->grid);
  const bool xf = !ctx->grid.xform_is_identity;
  const size_t ysmem = size_t(S->K.idxu_max) * 32 * sizeof(typename R2<real>::type);
  const bool fused = getenv("XSB_SNAP_FUSED") != nullptr;        // development switch: the single-kernel version
  auto setattr = [&](auto kern, size_t bytes) -> int { XSB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes))); return XSB_OK; };
  int rc;
  if( fused )
  {
    if( xf ) { if( (rc = setattr(snap_force_kernel<real, TJ, true, 0>, smem)) ) return rc; snap_force_kernel<real, TJ, true, 0><<<A.n_atoms, NT, smem, ctx->stream>>>(A, X, KK); }
    else     { if( (rc = setattr(snap_force_kernel<real, TJ, false, 0>, smem)) ) return rc; snap_force_kernel<real, TJ, false, 0><<<A.n_atoms, NT, smem, ctx->stream>>>(A, X, KK); }
    XSB_LAUNCH_CHECK(ctx);
    return XSB_OK;
  }
  const unsigned chunk = std::min(A.n_atoms, SNAP_CHUNK);
  const size_t words = size_t((chunk + 31) / 32) * 32 * S->K.idxu_max;
  XSB_CUDA(ctx, S->ubuf.reserve(words)); XSB_CUDA(ctx, S->ybuf.reserve(words));
  A.ubuf = reinterpret_cast<typename R2<real>::type*>(S->ubuf.p); A.ybuf = reinterpret_cast<typename R2<real>::type*>(S->ybuf.p);      // sized for double2, float2 uses half
  XSB_CUDA(ctx, S->nbtab.reserve(size_t(chunk) * 6 * SNAP_NN_TAB)); XSB_CUDA(ctx, S->nbcnt.reserve(chunk));
  A.nbtab = S->nbtab.p; A.nbcnt = S->nbcnt.p;
  // forward mode: (direction, row) warps pay off once the per-thread state of the 3-direction sweep no longer fits the
  // register file (measured: 2J = 8 faster, 2J <= 6 slower than the one-thread-per-row sweep)
  constexpr bool DIRSPLIT = TJ >= 7;
  // force kernel: reverse mode (snap_fr_kernel) unless XSB_SNAP_FKERNEL=1 selects the forward-mode kernels for A/B: one CTA
  // per (atom, direction) at 2J >= 7, one thread per row carrying all three directions below
  const char* fk = getenv("XSB_SNAP_FKERNEL");
  const bool reverse = fk == nullptr || fk[0] == '0';
  const bool dircta = !reverse && DIRSPLIT;
  const size_t frsmem = size_t(S->K.idxu_max + 32 * (SnapHist<TJ>::total() ? SnapHist<TJ>::total() : 1)) * sizeof(typename R2<real>::type)
                      + size_t(32) * SNAP_MBOX_STRIDE(MB ? MB : 1) * sizeof(typename R2<real>::type) + SNAP_NN_MAX * (5 * sizeof(real) + sizeof(unsigned))
                      + size_t(NR) * 5 * 32 * sizeof(real) + 64;
  if( reverse ) { if( xf ) { if( (rc = setattr(snap_fr_kernel<real, TJ, true>, frsmem)) ) return rc; } else { if( (rc = setattr(snap_fr_kernel<real, TJ, false>, frsmem)) ) return rc; } }
  const size_t fdsmem = size_t(S->K.idxu_max) * sizeof(typename R2<real>::type) + size_t(32) * SNAP_MBOX_STRIDE(2 * (MB ? MB : 1)) * sizeof(typename R2<real>::type) + SNAP_NN_MAX * (5 * sizeof(real) + sizeof(unsigned))
                      + size_t(NR) * 32 * sizeof(real) + 64;
  if( dircta ) { if( xf ) { if( (rc = setattr(snap_fd_kernel<real, TJ, true>, fdsmem)) ) return rc; } else { if( (rc = setattr(snap_fd_kernel<real, TJ, false>, fdsmem)) ) return rc; } }
  if( xf ) { if( (rc = setattr(snap_force_kernel<real, TJ, true, 1>, smem)) ) return rc; if( (rc = setattr(snap_force_kernel<real, TJ, true, 3>, smem)) ) return rc; }
  else     { if( (rc = setattr(snap_force_kernel<real, TJ, false, 1>, smem)) ) return rc; if( (rc = setattr(snap_force_kernel<real, TJ, false, 3>, smem)) ) return rc; }
  if( (rc = setattr(snap_y_kernel<real, TJ>, ysmem)) ) return rc;
  const size_t y2smem = ysmem + size_t(S->n_y2cg) * sizeof(real);
  if( (rc = setattr(snap_y2_kernel<real>, y2smem)) ) return rc;
  const bool uold = getenv("XSB_SNAP_UKERNEL") != nullptr;        // A/B switch: PHASE 1 of the fused kernel
  const size_t usmem = size_t(S->K.idxu_max) * sizeof(typename R2<real>::type) + size_t(32) * SNAP_MBOX_STRIDE(MB ? MB : 1) * sizeof(typename R2<real>::type)
                     + SNAP_NN_MAX * (5 * sizeof(real) + sizeof(unsigned)) + 64;
  if( xf ) { if( (rc = setattr(snap_u_kernel<real, TJ, true>, usmem)) ) return rc; } else { if( (rc = setattr(snap_u_kernel<real, TJ, false>, usmem)) ) return rc; }
  const bool yold = getenv("XSB_SNAP_YKERNEL") != nullptr;        // A/B switch: the one-element-per-work-item kernel
  for(unsigned base = 0; base < A.n_atoms; base += chunk)
  {
    const unsigned cnt = std::min(chunk, A.n_atoms - base);
    A.base = base;
    if( uold )
    {
      if( xf ) snap_force_kernel<real, TJ, true, 1><<<cnt, NT, smem, ctx->stream>>>(A, X, KK);
      else     snap_force_kernel<real, TJ, false, 1><<<cnt, NT, smem, ctx->stream>>>(A, X, KK);
    }
    else if( xf ) snap_u_kernel<real, TJ, true><<<cnt, NT, usmem, ctx->stream>>>(A, X, KK);
    else          snap_u_kernel<real, TJ, false><<<cnt, NT, usmem, ctx->stream>>>(A, X, KK);
    XSB_LAUNCH_CHECK(ctx);
    if( yold ) snap_y_kernel<real, TJ><<<(cnt + 31) / 32, 512, ysmem, ctx->stream>>>(A, KK);
    else       snap_y2_kernel<real><<<(cnt + 31) / 32, 512, y2smem, ctx->stream>>>(A, KK);
    XSB_LAUNCH_CHECK(ctx);
    if( reverse )
    {
      if( xf ) snap_fr_kernel<real, TJ, true><<<cnt, NT, frsmem, ctx->stream>>>(A, X, KK);
      else     snap_fr_kernel<real, TJ, false><<<cnt, NT, frsmem, ctx->stream>>>(A, X, KK);
    }
    else if( dircta )
    {
      if( xf ) snap_fd_kernel<real, TJ, true><<<3 * cnt, NT, fdsmem, ctx->stream>>>(A, X, KK);
      else     snap_fd_kernel<real, TJ, false><<<3 * cnt, NT, fdsmem, ctx->stream>>>(A, X, KK);
    }
    else
    {
      if( xf ) snap_force_kernel<real, TJ, true, 3><<<cnt, NT, smem, ctx->stream>>>(A, X, KK);
      else     snap_force_kernel<real, TJ, false, 3><<<cnt, NT, smem, ctx->stream>>>(A, X, KK);
    }
    XSB_LAUNCH_CHECK(ctx);
  }
  return XSB_OK;
}

extern "C" {

int xsb_snap_ncoeff(int twojmax)
{
  if( twojmax < 0 || twojmax > 8 ) return -1;
  int n = 0;
  for(int j1 = 0; j1 <= twojmax; j1++) for(int j2 = 0; j2 <= j1; j2++) for(int j = j1 - j2; j <= std::min(twojmax, j1 + j2); j += 2) if( j >= j1 ) ++n;
  return n;
}

int xsb_snap_set(xsb_ctx* ctx, const xsb_snap_params* p)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, p != nullptr && p->radelem && p->wjelem && p->beta, XSB_ERR_INVALID, "snap: null parameters");
  XSB_REQUIRE(ctx, p->twojmax >= 1 && p->twojmax <= 8, XSB_ERR_UNSUPPORTED, "snap: twojmax must be in 1..8");
  XSB_REQUIRE(ctx, p->nelements >= 1 && p->nelements <= 8, XSB_ERR_INVALID, "snap: 1..8 elements");
  XSB_REQUIRE(ctx, p->quadraticflag == 0 && p->chemflag == 0 && p->switchinnerflag == 0, XSB_ERR_UNSUPPORTED, "snap: quadratic / chem / inner-switch variants are not implemented");
  XSB_REQUIRE(ctx, p->rcutfac > 0.0, XSB_ERR_INVALID, "snap: rcutfac must be > 0");
  XSB_CUDA(ctx, cudaSetDevice(ctx->device));
  SnapDev* S = g_snap_of(ctx);
  if( !S ) { S = new SnapDev; ctx->snap = S; }
  SnapTables T(p->twojmax);
  SnapConst& K = S->K; K = SnapConst{};
  for(int a = 1; a <= p->twojmax; a++) for(int b = 1; b <= p->twojmax; b++) K.rootpq[a][b] = std::sqrt(double(a) / b);
  for(int j = 0; j <= p->twojmax; j++) K.idxu_block[j] = T.idxu_block[j];
  K.twojmax = p->twojmax; K.idxu_max = T.idxu_max; K.idxz_max = int(T.idxz.size()); K.ncoeff = T.ncoeff; K.nelements = p->nelements;
  K.switchflag = p->switchflag; K.bzeroflag = p->bzeroflag; K.rfac0 = p->rfac0; K.rmin0 = p->rmin0; K.rcutfac = p->rcutfac; K.wself = 1.0;
  std::vector<double> betaz(size_t(p->nelements) * T.idxz.size());
  double radmax = 0.0;
  for(int e = 0; e < p->nelements; e++)
  {
    const double* be = p->beta + size_t(e) * (T.ncoeff + 1);
    K.radelem[e] = p->radelem[e]; K.wjelem[e] = p->wjelem[e]; K.beta0[e] = be[0]; K.bzero_e[e] = 0.0;
    radmax = std::max(radmax, p->radelem[e]);
    if( p->bzeroflag ) for(int k = 0; k < T.ncoeff; k++) K.bzero_e[e] += be[1 + k] * (K.wself * K.wself * K.wself) * (T.b_j[k] + 1);
    for(size_t z = 0; z < T.idxz.size(); z++) betaz[size_t(e) * T.idxz.size() + z] = T.betaj(T.idxz[z], be + 1);
  }
  S->rcut_max = 2.0 * radmax * p->rcutfac;
  XSB_CUDA(ctx, S->idxz.reserve(T.idxz.size())); XSB_CUDA(ctx, S->cglist.reserve(T.cglist.size())); XSB_CUDA(ctx, S->betaz.reserve(betaz.size())); XSB_CUDA(ctx, S->err.reserve(4));
  XSB_CUDA(ctx, cudaMemcpyAsync(S->idxz.p, T.idxz.data(), T.idxz.size() * sizeof(SnapZ), cudaMemcpyHostToDevice, ctx->stream));
  XSB_CUDA(ctx, cudaMemcpyAsync(S->cglist.p, T.cglist.data(), T.cglist.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  XSB_CUDA(ctx, cudaMemcpyAsync(S->betaz.p, betaz.data(), betaz.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  XSB_CUDA(ctx, cudaMemsetAsync(S->err.p, 0, 4 * sizeof(int), ctx->stream));
  // snap_y_kernel work items: one per distinct Y element, with the idxz entries that feed it; most expensive first
  std::vector<int4> tasks; std::vector<SnapZ> zsort; std::vector<double> bsort(betaz.size());
  {
    std::vector<int> order(T.idxz.size());
    for(size_t i = 0; i < order.size(); i++) order[i] = int(i);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return T.idxz[a].jju < T.idxz[b].jju; });
    std::vector<std::pair<long, int4>> tmp;
    for(size_t i = 0; i < order.size(); )
    {
      size_t e = i; long cost = 0;
      while( e < order.size() && T.idxz[order[e]].jju == T.idxz[order[i]].jju ) { cost += long(T.idxz[order[e]].na) * T.idxz[order[e]].nb + 4; ++e; }
      tmp.push_back({ cost, make_int4(int(T.idxz[order[i]].jju), int(i), int(e - i), 0) });
      i = e;
    }
    std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<long, int4>& a, const std::pair<long, int4>& b) { return a.first > b.first; });
    for(auto& t : tmp) tasks.push_back(t.second);
    for(size_t i = 0; i < order.size(); i++)
    {
      zsort.push_back(T.idxz[order[i]]);
      for(int e = 0; e < p->nelements; e++) bsort[size_t(e) * order.size() + i] = betaz[size_t(e) * order.size() + order[i]];
    }
  }
  // snap_y2_kernel tables: triples grouped by j, zero-padded Clebsch-Gordan rows, beta per (element, triple), work items =
  // blocks of <= 5 consecutive ma of one row (j, mb), most expensive first
  std::vector<SnapYTri> y2tri; std::vector<SnapYTask> y2task; std::vector<double> y2cg, y2beta;
  {
    struct TJ3 { int j1, j2, j; };
    std::vector<TJ3> tl;
    std::vector<int> first(p->twojmax + 2, 0);
    for(int j = 0; j <= p->twojmax; j++)
    {
      first[j] = int(tl.size());
      for(int j1 = 0; j1 <= p->twojmax; j1++) for(int j2 = 0; j2 <= j1; j2++)
        if( j >= j1 - j2 && j <= std::min(p->twojmax, j1 + j2) && ((j1 + j2 - j) & 1) == 0 ) tl.push_back({ j1, j2, j });
    }
    first[p->twojmax + 1] = int(tl.size());
    y2beta.assign(size_t(p->nelements) * tl.size(), 0.0);
    for(size_t t = 0; t < tl.size(); t++)
    {
      const TJ3 q = tl[t];
      const int P = (q.j + 2) & ~1, src = T.cg_block[T.b3(q.j1, q.j2, q.j)], sh = (q.j1 + q.j2 - q.j) / 2;
      SnapYTri y{ (unsigned char)q.j1, (unsigned char)q.j2, (unsigned char)sh, (unsigned char)P, int(y2cg.size()) };
      y2tri.push_back(y);
      for(int m2 = 0; m2 <= q.j2; m2++) for(int m = 0; m < P; m++)
      {
        const int m1 = m + sh - m2;
        y2cg.push_back(m <= q.j && m1 >= 0 && m1 <= q.j1 ? T.cglist[size_t(src) + m1 * (q.j2 + 1) + m2] : 0.0);
      }
      SnapZ z{}; z.j1 = (unsigned char)q.j1; z.j2 = (unsigned char)q.j2; z.j = (unsigned char)q.j;
      for(int e = 0; e < p->nelements; e++) y2beta[size_t(e) * tl.size() + t] = T.betaj(z, p->beta + size_t(e) * (T.ncoeff + 1) + 1);
    }
    std::vector<std::pair<long, SnapYTask>> tmp;
    for(int j = 0; j <= p->twojmax; j++) for(int mb = 0; 2 * mb <= j; mb++)
    {
      const int n = 2 * mb == j ? mb + 1 : j + 1;          // the second half of a middle row is never used (weight 0)
      int widths[3] = { n, 0, 0 };
      // blocks start at even ma: their coefficients are fetched as aligned pairs
      if( n == 6 ) { widths[0] = 4; widths[1] = 2; } else if( n == 7 ) { widths[0] = 4; widths[1] = 3; } else if( n == 8 ) { widths[0] = 4; widths[1] = 4; } else if( n == 9 ) { widths[0] = 4; widths[1] = 5; }
      for(int b = 0, ma0 = 0; b < 3 && widths[b]; ma0 += widths[b], b++)
      {
        const int W = widths[b];
        long cost = 0;
        for(int t = first[j]; t < first[j + 1]; t++)
        {
          const int j1 = tl[t].j1, j2 = tl[t].j2, s = (j1 + j2 - j) / 2;
          const int nb = std::min(j1, mb + s) - std::max(0, mb + s - j2) + 1, nit = std::min(j2, ma0 + W - 1 + s) - std::max(0, ma0 + s - j1) + 1;
          cost += long(nb) * (long(nit) * (6 * W + 6) + 4 * W + 12) + 10;
        }
        tmp.push_back({ cost, SnapYTask{ (unsigned char)j, (unsigned char)mb, (unsigned char)ma0, (unsigned char)W, (unsigned short)first[j], (unsigned short)(first[j + 1] - first[j]) } });
      }
    }
    std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<long, SnapYTask>& a, const std::pair<long, SnapYTask>& b) { return a.first > b.first; });
    for(auto& t : tmp) y2task.push_back(t.second);
  }
  S->n_y2tri = int(y2tri.size()); S->n_y2task = int(y2task.size());
  while( y2cg.size() % 4 != 2 ) y2cg.push_back(0.0);      // room for the pair load of the last odd block; bulk copies move multiples of 16 bytes (floats: 4 words)
  y2cg.push_back(0.0); y2cg.push_back(0.0);
  S->n_y2cg = int(y2cg.size());
  std::vector<float> y2cg32(y2cg.begin(), y2cg.end()), y2beta32(y2beta.begin(), y2beta.end());
  XSB_CUDA(ctx, S->y2tri.reserve(y2tri.size())); XSB_CUDA(ctx, S->y2task.reserve(y2task.size())); XSB_CUDA(ctx, S->y2cg.reserve(y2cg.size())); XSB_CUDA(ctx, S->y2beta.reserve(y2beta.size()));
  XSB_CUDA(ctx, S->y2cg32.reserve(y2cg32.size())); XSB_CUDA(ctx, S->y2beta32.reserve(y2beta32.size()));
  XSB_CUDA(ctx, cudaMemcpyAsync(S->y2tri.p, y2tri.data(), y2tri.size() * sizeof(SnapYTri), cudaMemcpyHostToDevice, ctx->stream));
  XSB_CUDA(ctx, cudaMemcpyAsync(S->y2task.p, y2task.data(), y2task.size() * sizeof(SnapYTask), cudaMemcpyHostToDevice, ctx->stream));
  XSB_CUDA(ctx, cudaMemcpyAsync(S->y2cg.p, y2cg.data(), y2cg.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  XSB_CUDA(ctx, cudaMemcpyAsync(S->y2beta.p, y2beta.data(), y2beta.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  XSB_CUDA(ctx, cudaMemcpyAsync(S->y2cg32.p, y2cg32.data(), y2cg32.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  XSB_CUDA(ctx, cudaMemcpyAsync(S->y2beta32.p, y2beta32.data(), y2beta32.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // the staging vectors of this block die at the end of the function; keep it simple
  S->n_ytask = int(tasks.size());
  XSB_CUDA(ctx, S->zsort.reserve(zsort.size())); XSB_CUDA(ctx, S->betaz_sort.reserve(bsort.size())); XSB_CUDA(ctx, S->ytask.reserve(tasks.size()));
  XSB_CUDA(ctx, cudaMemcpyAsync(S->zsort.p, zsort.data(), zsort.size() * sizeof(SnapZ), cudaMemcpyHostToDevice, ctx->stream));
  XSB_CUDA(ctx, cudaMemcpyAsync(S->betaz_sort.p, bsort.data(), bsort.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  XSB_CUDA(ctx, cudaMemcpyAsync(S->ytask.p, tasks.data(), tasks.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
  {
    // XSB_FLAG_MIXED (the reference's SNAP_FP32_MATH build, snap_force.cu:25-29): the same tables rounded once from FP64
    std::vector<float> c32(T.cglist.begin(), T.cglist.end()), b32(betaz.begin(), betaz.end()), bs32(bsort.begin(), bsort.end());
    XSB_CUDA(ctx, S->cglist32.reserve(c32.size())); XSB_CUDA(ctx, S->betaz32.reserve(b32.size())); XSB_CUDA(ctx, S->betaz_sort32.reserve(bs32.size()));
    XSB_CUDA(ctx, cudaMemcpyAsync(S->cglist32.p, c32.data(), c32.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    XSB_CUDA(ctx, cudaMemcpyAsync(S->betaz32.p, b32.data(), b32.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    XSB_CUDA(ctx, cudaMemcpyAsync(S->betaz_sort32.p, bs32.data(), bs32.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    SnapConstT<float>& F = S->K32; F = SnapConstT<float>{};
    for(int a = 0; a < 10; a++) for(int b = 0; b < 10; b++) F.rootpq[a][b] = float(K.rootpq[a][b]);
    for(int j = 0; j < 10; j++) F.idxu_block[j] = K.idxu_block[j];
    F.twojmax = K.twojmax; F.idxu_max = K.idxu_max; F.idxz_max = K.idxz_max; F.ncoeff = K.ncoeff; F.nelements = K.nelements; F.switchflag = K.switchflag; F.bzeroflag = K.bzeroflag;
    F.rfac0 = float(K.rfac0); F.rmin0 = float(K.rmin0); F.rcutfac = float(K.rcutfac); F.wself = float(K.wself);
    for(int e = 0; e < 8; e++) { F.radelem[e] = float(K.radelem[e]); F.wjelem[e] = float(K.wjelem[e]); F.beta0[e] = K.beta0[e]; F.bzero_e[e] = K.bzero_e[e]; }
    XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // the staging vectors die here
  }
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  S->set = true;
  return XSB_OK;
}

double xsb_snap_rcut_max(xsb_ctx* ctx) { SnapDev* S = ctx ? g_snap_of(ctx) : nullptr; return S && S->set ? S->rcut_max : 0.0; }

int xsb_snap_force(xsb_ctx* ctx, int flags)
{
  XSB_ENTER(ctx);
  SnapDev* S = g_snap_of(ctx);
  XSB_REQUIRE(ctx, S && S->set, XSB_ERR_STATE, "xsb_snap_set must be called first");
  XSB_REQUIRE(ctx, ctx->nbh_built, XSB_ERR_STATE, "chunk_neighbors must be built before a force operator");
  XSB_REQUIRE(ctx, S->rcut_max <= ctx->nbh_dist, XSB_ERR_INVALID, "snap cutoff exceeds the neighbour-list distance nbh_dist_lab");
  XSB_CUDA(ctx, cudaSetDevice(ctx->device));
  const bool ghost = flags & XSB_FLAG_GHOST, virial = flags & XSB_FLAG_VIRIAL;
  if( virial ) { int rc = xsb_internal_ensure_virial(ctx); if( rc ) return rc; }
  const bool mixed = (flags & XSB_FLAG_MIXED) != 0;      // FP32 Wigner / Clebsch-Gordan arithmetic, FP64 positions, forces and energies
  const unsigned n_atoms = unsigned(ghost ? ctx->n : ctx->n_own);
  if( n_atoms == 0 ) return XSB_OK;
  int rc = XSB_ERR_UNSUPPORTED;
  ctx->prof_begin(XSB_PROF_SNAP);
  auto go = [&](auto Areal, const auto& KK) -> int
  {
    typedef decltype(Areal) ArgsT;
    switch( S->K.twojmax )
    {
      case 1: return snap_launch<typename ArgsT::real_t, 1>(ctx, S, Areal, KK); case 2: return snap_launch<typename ArgsT::real_t, 2>(ctx, S, Areal, KK);
      case 3: return snap_launch<typename ArgsT::real_t, 3>(ctx, S, Areal, KK); case 4: return snap_launch<typename ArgsT::real_t, 4>(ctx, S, Areal, KK);
      case 5: return snap_launch<typename ArgsT::real_t, 5>(ctx, S, Areal, KK); case 6: return snap_launch<typename ArgsT::real_t, 6>(ctx, S, Areal, KK);
      case 7: return snap_launch<typename ArgsT::real_t, 7>(ctx, S, Areal, KK); case 8: return snap_launch<typename ArgsT::real_t, 8>(ctx, S, Areal, KK);
      default: return XSB_ERR_UNSUPPORTED;
    }
  };
  const unsigned char* types = S->K.nelements > 1 ? ctx->type.p : nullptr;
  const unsigned* sel = ghost ? nullptr : ctx->own_atoms.p;
  double* epp = (flags & XSB_FLAG_ENERGY) ? ctx->f64[XSB_F_EP].p : nullptr;
  double* virp = virial ? ctx->f64[XSB_F_VIRIAL].p : nullptr;
  if( mixed )
  {
    SnapArgsT<float> A{ ctx->f64[XSB_F_RX].p, ctx->f64[XSB_F_RY].p, ctx->f64[XSB_F_RZ].p, types, ctx->nbh_off.p, ctx->nbh_idx.p, sel, n_atoms,
                        S->idxz.p, S->cglist32.p, S->betaz32.p, ctx->f64[XSB_F_FX].p, ctx->f64[XSB_F_FY].p, ctx->f64[XSB_F_FZ].p, epp, virp, S->err.p,
                        S->clocks ? S->clk.p : nullptr, nullptr, nullptr, 0u, nullptr, nullptr, S->zsort.p, S->betaz_sort32.p, S->ytask.p, S->n_ytask, S->y2tri.p, S->y2cg32.p, S->y2beta32.p, S->y2task.p, S->n_y2tri, S->n_y2task, S->n_y2cg };
    rc = go(A, S->K32);
  }
  else
  {
    SnapArgsT<double> A{ ctx->f64[XSB_F_RX].p, ctx->f64[XSB_F_RY].p, ctx->f64[XSB_F_RZ].p, types, ctx->nbh_off.p, ctx->nbh_idx.p, sel, n_atoms,
                         S->idxz.p, S->cglist.p, S->betaz.p, ctx->f64[XSB_F_FX].p, ctx->f64[XSB_F_FY].p, ctx->f64[XSB_F_FZ].p, epp, virp, S->err.p,
                         S->clocks ? S->clk.p : nullptr, nullptr, nullptr, 0u, nullptr, nullptr, S->zsort.p, S->betaz_sort.p, S->ytask.p, S->n_ytask, S->y2tri.p, S->y2cg.p, S->y2beta.p, S->y2task.p, S->n_y2tri, S->n_y2task, S->n_y2cg };
    rc = go(A, S->K);
  }
  ctx->prof_end(XSB_PROF_SNAP);
  if( rc ) return rc;
  // an atom with more than SNAP_NN_MAX in-range neighbours had the excess pairs dropped: never hand such forces back as a
  // success (the read-back costs one stream sync per call; a SNAP call is tens of milliseconds of kernels)
  int over = 0;
  XSB_CUDA(ctx, cudaMemcpyAsync(&over, S->err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if( over )
  {
    S->overflowed = true;      // sticky until xsb_snap_overflow() is read
    XSB_CUDA(ctx, cudaMemsetAsync(S->err.p, 0, sizeof(int), ctx->stream));
    return ctx->fail(XSB_ERR_OVERFLOW, "snap_force: an atom has more than %d neighbours inside the SNAP cutoff; forces and energies of this call are incomplete", SNAP_NN_TAB);
  }
  return XSB_OK;
}

// development aid (not part of the ABI header): per-phase cycle sums of the snap kernel, summed over CTAs
extern "C" int xsbdbg_snap_clocks(xsb_ctx* ctx, int enable, unsigned long long* out8)
{
  SnapDev* S = g_snap_of(ctx);
  if( !S ) return XSB_ERR_STATE;
  if( enable && !S->clocks ) { if( S->clk.reserve(8) != cudaSuccess ) return XSB_ERR_CUDA; cudaMemsetAsync(S->clk.p, 0, 64, ctx->stream); S->clocks = true; }
  if( out8 && S->clocks ) { cudaMemcpyAsync(out8, S->clk.p, 64, cudaMemcpyDeviceToHost, ctx->stream); cudaStreamSynchronize(ctx->stream); cudaMemsetAsync(S->clk.p, 0, 64, ctx->stream); }
  if( !enable ) S->clocks = false;
  return XSB_OK;
}

// 1 when some atom had more than SNAP_NN_MAX in-range neighbours since the last call (forces are then incomplete)
int xsb_snap_overflow(xsb_ctx* ctx, int* flag)
{
  XSB_ENTER(ctx);
  SnapDev* S = g_snap_of(ctx);
  XSB_REQUIRE(ctx, S && S->set && flag, XSB_ERR_STATE, "xsb_snap_set must be called first");
  XSB_CUDA(ctx, cudaMemcpyAsync(flag, S->err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  XSB_CUDA(ctx, cudaMemsetAsync(S->err.p, 0, sizeof(int), ctx->stream));
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if( S->overflowed ) { *flag = 1; S->overflowed = false; }
  return XSB_OK;
}

} // extern "C"
