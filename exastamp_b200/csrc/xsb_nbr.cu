// xsb_nbr.cu -- chunk_neighbors operator.
//
// Device representation: flat CSR list (count / offset / u32 flat neighbour index) in canonical order
// (ascending neighbour cell, then ascending p_b) -- the order in which the reference stream is traversed.
// The reference's own uint16 per-cell stream (decoder: src/rigidmol/compute_pair_rigidmol.h:154-234) is
// produced on demand by xsb_chunk_neighbors_export() from the same list, honouring chunk_size and
// build_particle_offset (data/config/config_move_particles.msp:54-61).
//
// Build (tile path): nbr_count_kernel stages the cell block of a tile in shared memory, tests 32 candidates per step and
// keeps the survivor masks; after a scan of the counts nbr_expand_kernel replays the masks into the two index lists.
// Grids the tile path cannot serve fall back to two warp-per-particle sweeps (nbr_sweep_kernel: count, then fill) over
// the (2Ry+1)(2Rz+1) x-rows of cells around the particle's cell.
#include "xsb_tile.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_reduce.cuh>

namespace xsb
{

struct NbrParams
{
  GridView g;
  int Rx, Ry, Rz;
  double d2max;
  unsigned n;
};

// physical-space squared distance with the CPU's operation order (no FMA contraction) so that the
// membership test 0 < d2 < d2max is bit-identical to the host restatement.
template<bool XFORM>
__device__ __forceinline__ double nbh_d2(const GridView& g, double dx, double dy, double dz)
{
  if( XFORM )
  {
    const double* m = g.xf;
    const double x = __dadd_rn(__dadd_rn(__dmul_rn(m[0], dx), __dmul_rn(m[1], dy)), __dmul_rn(m[2], dz));
    const double y = __dadd_rn(__dadd_rn(__dmul_rn(m[3], dx), __dmul_rn(m[4], dy)), __dmul_rn(m[5], dz));
    const double z = __dadd_rn(__dadd_rn(__dmul_rn(m[6], dx), __dmul_rn(m[7], dy)), __dmul_rn(m[8], dz));
    dx = x; dy = y; dz = z;
  }
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// FILL=false: counts[a] ; FILL=true: idx[off[a] ...] in canonical order
template<bool XFORM, bool FILL>
__global__ void __launch_bounds__(256) nbr_sweep_kernel(NbrParams P, const unsigned* __restrict__ cell_start, const unsigned* __restrict__ cell_of,
                                                         const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
                                                         unsigned* __restrict__ counts, const unsigned long long* __restrict__ off, unsigned* __restrict__ idx,
                                                         unsigned long long* __restrict__ d2min_bits)
{
  const unsigned lane = threadIdx.x & 31u;
  const unsigned a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // no 32-bit overflow of the thread index above 2^27 atoms
  if( a >= P.n ) return;
  const unsigned ca = cell_of[a];
  const int nx = P.g.nx, ny = P.g.ny, nz = P.g.nz;
  const int ia = int(ca % unsigned(nx)), ja = int((ca / unsigned(nx)) % unsigned(ny)), ka = int(ca / (unsigned(nx) * unsigned(ny)));
  const double xa = rx[a], ya = ry[a], za = rz[a];
  const int ilo = max(0, ia - P.Rx), ihi = min(nx - 1, ia + P.Rx);
  unsigned cnt = 0;
  double dmin = 1.0e300;
  unsigned long long w = FILL ? off[a] : 0ull;
  for(int rk = -P.Rz; rk <= P.Rz; rk++)
  {
    const int kb = ka + rk; if( kb < 0 || kb >= nz ) continue;
    for(int rj = -P.Ry; rj <= P.Ry; rj++)
    {
      const int jb = ja + rj; if( jb < 0 || jb >= ny ) continue;
      const unsigned row = unsigned(nx) * (unsigned(jb) + unsigned(ny) * unsigned(kb));
      const unsigned b0 = cell_start[row + ilo], b1 = cell_start[row + ihi + 1];
      for(unsigned base = b0; base < b1; base += 32u)
      {
        const unsigned b = base + lane;
        bool keep = false;
        if( b < b1 && b != a )
        {
          const double d2 = nbh_d2<XFORM>(P.g, rx[b] - xa, ry[b] - ya, rz[b] - za);
          keep = d2 > 0.0 && d2 < P.d2max;
          if( !FILL && keep ) dmin = fmin(dmin, d2);
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if( FILL )
        {
          if( keep ) idx[w + __popc(m & ((1u << lane) - 1u))] = b;
          w += __popc(m);
        }
        else cnt += __popc(m);
      }
    }
  }
  if( !FILL )
  {
    if( lane == 0 ) counts[a] = cnt;
    // smallest pair distance of the list (positive doubles order like their bit patterns): sizes the shared-memory
    // window of the EAM spline tables (xsb_eam.cu)
#   pragma unroll
    for(int o = 16; o > 0; o >>= 1) dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    if( lane == 0 && dmin < 1.0e300 && (unsigned long long)__double_as_longlong(dmin) < *d2min_bits ) atomicMin(d2min_bits, (unsigned long long)__double_as_longlong(dmin));
  }
}

// ---- tile build: count sweep with survivor masks, then mask replay ------------------------------------------------------
// One CTA per tile (= cell): warp 0 lays out the tile's 27-cell block (same stage indexing as the force kernels,
// tile_meta_compute) and a table of 32-candidate "words" (stage index / flat index / end of the row, per word); the CTA
// copies the block's positions into shared memory once (AoS x,y,z: a warp's 64-bit loads at 24-byte stride are
// conflict-free), then its warps take the tile's atoms in turn.  For every word the 32 lanes test one candidate each with
// the reference's operation order (nbh_d2) and the ballot mask is kept: lane q of the warp holds the mask of word q, flushed as
// one coalesced store per 32 words.  nbr_expand_kernel replays the masks into the uint16 stage-index list (what the
// force kernels stream) and the u32 flat-index list (CSR view for the exporter / SNAP): no positions, no FP64 there.
constexpr int NBR_MAX_WORDS = 2048 / 32 + TILE_MAX_ROWS + 1;      // s_cap <= 2048 on this path

constexpr int NBR_MAX_TX = 8;                                       // widest tile (tile_plan)

// 32-candidate "words" of ONE central cell: the cells [i-Rx, i+Rx] of every staged row, i = that cell (for TX > 1 a
// sub-range of each staged row).  All atoms of the cell share the table; their survivor masks are indexed by word.
struct NbrWords
{
  unsigned short s[NBR_MAX_WORDS];     // stage index of bit 0
  unsigned short e[NBR_MAX_WORDS];     // stage index one past the last candidate of the word's row window
  unsigned g[NBR_MAX_WORDS];           // flat particle index of bit 0
  uint2 sc[NBR_MAX_WORDS];             // count sweep: { stage index of bit 0, valid candidates in the word (1..32) }
  unsigned n;
  unsigned a_end;                      // flat index one past the last atom of the cell
};

// executed by one full warp after tile_meta_compute(G, ..., M): table of cell i (a cell of the tile (ti, j, k))
__device__ __forceinline__ void nbr_words_compute(const TileGeom& G, const unsigned* __restrict__ cell_start, int i, int j, int k, const TileMeta& M, NbrWords& W)
{
  const unsigned lane = threadIdx.x & 31u;
  const int nry = 2 * G.Ry + 1, nrows = nry * (2 * G.Rz + 1);
  unsigned b = 0, e = 0;
  if( int(lane) < nrows )
  {
    const int kk = k + int(lane) / nry - G.Rz, jj = j + int(lane) % nry - G.Ry;
    if( jj >= 0 && jj < G.ny && kk >= 0 && kk < G.nz && M.s0[lane + 1] > M.s0[lane] )
    {
      const unsigned row = unsigned(G.nx) * (unsigned(jj) + unsigned(G.ny) * unsigned(kk));
      const unsigned gb = cell_start[row + max(0, i - G.Rx)], ge = cell_start[row + min(G.nx, i + G.Rx + 1)];
      if( ge > gb ) { b = gb - M.g0[lane]; e = ge - M.g0[lane]; }      // window of this cell inside the staged (widened) row
    }
  }
  const unsigned nw = (e - b + 31u) >> 5;
  unsigned pre = nw;
# pragma unroll
  for(int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, pre, o); if( int(lane) >= o ) pre += v; }
  if( int(lane) < nrows )
    for(unsigned q = 0; q < nw; q++)
    {
      W.s[pre - nw + q] = (unsigned short)(M.s0[lane] + b + 32u * q);
      W.e[pre - nw + q] = (unsigned short)(M.s0[lane] + e);
      W.g[pre - nw + q] = M.g0[lane] + b + 32u * q;
      W.sc[pre - nw + q] = make_uint2(M.s0[lane] + b + 32u * q, min(32u, e - b - 32u * q));
    }
  if( lane == 31 ) W.n = pre;
  if( lane == 0 ) W.a_end = cell_start[unsigned(G.nx) * (unsigned(j) + unsigned(G.ny) * unsigned(k)) + unsigned(i) + 1u];
}

// the CTA's warps fill the tables of the tile's cells (after M is visible to all of them); caller barriers afterwards
__device__ __forceinline__ unsigned nbr_tile_words(const TileGeom& G, const unsigned* __restrict__ cell_start, int ti, int j, int k, const TileMeta& M, NbrWords* Wc)
{
  const int i0 = ti * G.TX, nc = min(G.nx, i0 + G.TX) - i0;
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for(int t = warp; t < nc; t += nwarps) nbr_words_compute(G, cell_start, i0 + t, j, k, M, Wc[t]);
  return unsigned(nc);
}

__device__ __forceinline__ double lds_f64(unsigned addr) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v; }
__device__ __forceinline__ double lds_f64_8(unsigned addr) { double v; asm volatile("ld.shared.f64 %0, [%1+8];" : "=d"(v) : "r"(addr)); return v; }
__device__ __forceinline__ double lds_f64_16(unsigned addr) { double v; asm volatile("ld.shared.f64 %0, [%1+16];" : "=d"(v) : "r"(addr)); return v; }

// Distance test of the count sweep.  Out of ~1500 staged candidates per atom only ~13 % survive, so the sweep runs in
// FP32 on coordinates relative to the tile (12 B per staged atom instead of 24, FP32 pipe instead of FP64) and only
// a candidate whose FP32 distance falls inside a guard band around the two thresholds (0 and nbh_dist^2) is decided by
// the exact FP64 test in the reference's operation order (nbh_d2, positions re-read from global memory).  The band is
// a bound on |d2_fp32 - d2_exact| for any candidate within 1 % of the cut-off (host: nbr_fp32_band); everything
// farther away is classified correctly by a wide margin.  The resulting list is bit-identical to the all-FP64 sweep.
struct NbrF32 { float m[9]; float d2max, band; int xform; };

// count sweep of NA (1 or 2) consecutive atoms of ONE cell: the word descriptor and the candidate's coordinates are loaded once
// per word and tested against both central atoms (issue-bound loop: 39 -> ~31 instructions per atom and word)
template<int NA, bool XFORM>
__device__ __forceinline__ void nbr_count_atoms(const NbrWords& W, const unsigned a0, const unsigned sbase, const unsigned lane12, const unsigned c_off, const unsigned lane,
                                                const float lo_sure, const float hi_sure, const float hi_out, const NbrF32& F, const GridView& gv, const double d2max,
                                                const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
                                                unsigned* __restrict__ counts, unsigned* __restrict__ masks, const unsigned mask_stride, float& dmin_f)
{
  const unsigned nw = W.n;
  unsigned ada[NA], cnt[NA]; float xa[NA], ya[NA], za[NA]; unsigned* mrow[NA];
# pragma unroll
  for(int i = 0; i < NA; i++)
  {
    ada[i] = sbase + 12u * (a0 + unsigned(i) + c_off);
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xa[i]) : "r"(ada[i]));
    asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(ya[i]) : "r"(ada[i]));
    asm volatile("ld.shared.f32 %0, [%1+8];" : "=f"(za[i]) : "r"(ada[i]));
    cnt[i] = 0;
    mrow[i] = masks + size_t(a0 + unsigned(i)) * mask_stride;
  }
  const unsigned wsc = smem_u32(&W.sc[0]);
  for(unsigned q0 = 0; q0 < nw; q0 += 32u)
  {
    const unsigned qe = min(32u, nw - q0);
    unsigned held[NA];
#   pragma unroll
    for(int i = 0; i < NA; i++) held[i] = 0;
#   pragma unroll 2
    for(unsigned tq = 0; tq < qe; tq++)
    {
      unsigned ws, wc;
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ws), "=r"(wc) : "r"(wsc + 8u * (q0 + tq)));
      // branch-free body: lanes past the end of the word read a valid slot and are masked out
      const unsigned ad = lane12 + 12u * ws;
      float cx, cy, cz;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(cx) : "r"(ad));
      asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(cy) : "r"(ad));
      asm volatile("ld.shared.f32 %0, [%1+8];" : "=f"(cz) : "r"(ad));
      float d2f[NA]; bool keep[NA], maybe[NA]; bool any_maybe = false;
#     pragma unroll
      for(int i = 0; i < NA; i++)
      {
        float dx = cx - xa[i], dy = cy - ya[i], dz = cz - za[i];
        if( XFORM )
        {
          const float x = F.m[0] * dx + F.m[1] * dy + F.m[2] * dz, y = F.m[3] * dx + F.m[4] * dy + F.m[5] * dz, z = F.m[6] * dx + F.m[7] * dy + F.m[8] * dz;
          dx = x; dy = y; dz = z;
        }
        d2f[i] = dx * dx + dy * dy + dz * dz;
        const bool valid = (lane < wc) & (ad != ada[i]);
        keep[i] = valid & (d2f[i] >= lo_sure) & (d2f[i] < hi_sure);
        maybe[i] = valid & !keep[i] & (d2f[i] <= hi_out);
        any_maybe |= maybe[i];
      }
      unsigned m[NA];
#     pragma unroll
      for(int i = 0; i < NA; i++)
      {
        if( keep[i] ) dmin_f = fminf(dmin_f, d2f[i]);
        m[i] = __ballot_sync(0xffffffffu, keep[i]);
      }
      if( __any_sync(0xffffffffu, any_maybe) )
      {
        // guard band (rare): the exact FP64 test in the reference's operation order decides; its survivors join the mask
#       pragma unroll
        for(int i = 0; i < NA; i++)
        {
          bool ex = false;
          if( maybe[i] )
          {
            const unsigned b = W.g[q0 + tq] + lane, a = a0 + unsigned(i);
            const double d2 = nbh_d2<XFORM>(gv, rx[b] - rx[a], ry[b] - ry[a], rz[b] - rz[a]);
            ex = d2 > 0.0 && d2 < d2max;
            if( ex ) dmin_f = fminf(dmin_f, d2f[i]);
          }
          m[i] |= __ballot_sync(0xffffffffu, ex);
        }
      }
#     pragma unroll
      for(int i = 0; i < NA; i++)
      {
        cnt[i] += __popc(m[i]);
        held[i] = tq == lane ? m[i] : held[i];
      }
    }
#   pragma unroll
    for(int i = 0; i < NA; i++) if( lane < qe ) mrow[i][q0 + lane] = held[i];          // up to 32 words -> one coalesced store
  }
# pragma unroll
  for(int i = 0; i < NA; i++) if( lane == 0 ) counts[a0 + unsigned(i)] = cnt[i];
}

template<bool XFORM>
__global__ void __launch_bounds__(256) nbr_count_kernel(TileGeom G, GridView gv, double d2max, NbrF32 F, const unsigned* __restrict__ cell_start,
                                                         const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
                                                         unsigned* __restrict__ counts, unsigned* __restrict__ masks, unsigned mask_stride,
                                                         unsigned long long* __restrict__ d2min_bits)
{
  extern __shared__ __align__(16) unsigned char nbr_smem[];
  __shared__ TileMeta M;
  __shared__ NbrWords Wc[NBR_MAX_TX];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  int ti, j, k; tile_coords(G, blockIdx.x, ti, j, k);
  if( warp == 0 ) tile_meta_compute(G, cell_start, ti, j, k, M);
  __syncthreads();
  if( M.a_begin == M.a_end ) return;
  const unsigned ncell = nbr_tile_words(G, cell_start, ti, j, k, M, Wc);
  // coordinates relative to the first central atom of the tile, rounded to FP32 once
  const double ox = rx[M.a_begin], oy = ry[M.a_begin], oz = rz[M.a_begin];
  float* sxyz = reinterpret_cast<float*>(nbr_smem);
  for(unsigned r = warp; r < M.nrows; r += nwarps)
  {
    const unsigned s0 = M.s0[r], len = M.s0[r + 1] - s0, g0 = M.g0[r];
    for(unsigned t = lane; t < len; t += 32u)
    {
      float* p = sxyz + 3u * (s0 + t);
      p[0] = float(rx[g0 + t] - ox); p[1] = float(ry[g0 + t] - oy); p[2] = float(rz[g0 + t] - oz);
    }
  }
  __syncthreads();
  unsigned sbase = smem_u32(sxyz);
  const unsigned a_end = M.a_end, c_off = M.c_off;
  float lo_sure = F.band, hi_sure = F.d2max - F.band, hi_out = F.d2max + F.band;
  // keep the loop invariants in registers (ptxas otherwise rebuilds the shared-memory base and the thresholds per word)
  asm volatile("" : "+r"(sbase), "+f"(lo_sure), "+f"(hi_sure), "+f"(hi_out));
  const unsigned lane12 = sbase + 12u * lane;
  float dmin_f = 3.0e38f;            // smallest FP32 d2 among the kept pairs (sizes the EAM table window: a bound within 1e-3 is enough)
  // a warp takes two consecutive atoms at a time; a pair that straddles two cells of the tile (different word tables) is
  // swept one atom after the other
  for(unsigned a = M.a_begin + 2u * warp; a < a_end; a += 2u * nwarps)
  {
    unsigned t = 0;                                    // cell of the tile that holds atom a (warp-uniform)
    while( t + 1 < ncell && a >= Wc[t].a_end ) ++t;
    if( a + 1u < Wc[t].a_end )
      nbr_count_atoms<2, XFORM>(Wc[t], a, sbase, lane12, c_off, lane, lo_sure, hi_sure, hi_out, F, gv, d2max, rx, ry, rz, counts, masks, mask_stride, dmin_f);
    else
    {
      nbr_count_atoms<1, XFORM>(Wc[t], a, sbase, lane12, c_off, lane, lo_sure, hi_sure, hi_out, F, gv, d2max, rx, ry, rz, counts, masks, mask_stride, dmin_f);
      if( a + 1u < a_end )
      {
        while( t + 1 < ncell && a + 1u >= Wc[t].a_end ) ++t;
        nbr_count_atoms<1, XFORM>(Wc[t], a + 1u, sbase, lane12, c_off, lane, lo_sure, hi_sure, hi_out, F, gv, d2max, rx, ry, rz, counts, masks, mask_stride, dmin_f);
      }
    }
  }
# pragma unroll
  for(int o = 16; o > 0; o >>= 1) dmin_f = fminf(dmin_f, __shfl_xor_sync(0xffffffffu, dmin_f, o));
  if( lane == 0 && dmin_f < 3.0e38f )
  {
    const double lb = fmax(0.0, double(dmin_f) * (1.0 - 1.0e-3) - double(F.band));      // lower bound of the exact minimum
    const unsigned long long bits = (unsigned long long)__double_as_longlong(lb);
    if( bits < *d2min_bits ) atomicMin(d2min_bits, bits);
  }
}

// DEAL: the uint16 list of an atom is written in "bank-dealt" order instead of the canonical one (which the u32 list
// keeps): entries are dealt out in rounds, each round holding at most one entry per residue (stage index mod 16),
// ascending residue inside a round.  The force kernels gather x,y,z (8-byte words, SoA) from the stage with 16
// consecutive list entries per half-warp, so an aligned run of 16 entries with distinct residues is one conflict-free
// shared-memory wavefront per field instead of ~3 (ncu: >50 % of the wavefronts of the EAM passes were bank conflicts).
// lcap = per-warp capacity (entries) of the shared-memory work area.
template<bool DEAL>
__global__ void __launch_bounds__(256) nbr_expand_kernel(TileGeom G, const unsigned* __restrict__ cell_start, const unsigned long long* __restrict__ off,
                                                          const unsigned* __restrict__ masks, unsigned mask_stride,
                                                          unsigned* __restrict__ idx32, unsigned short* __restrict__ idx16, unsigned lcap)
{
  extern __shared__ __align__(16) unsigned short deal_smem[];
  __shared__ TileMeta M;
  __shared__ NbrWords Wc[NBR_MAX_TX];
  __shared__ unsigned rcnt[8][16];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  int ti, j, k; tile_coords(G, blockIdx.x, ti, j, k);
  if( warp == 0 ) tile_meta_compute(G, cell_start, ti, j, k, M);
  __syncthreads();
  if( M.a_begin == M.a_end ) return;
  const unsigned ncell = nbr_tile_words(G, cell_start, ti, j, k, M, Wc);
  __syncthreads();
  const unsigned lt = (1u << lane) - 1u;
  unsigned short* bidx = deal_smem + size_t(warp) * 2u * lcap;   // canonical stage indices of the current atom
  unsigned short* brk = bidx + lcap;                             // rank of each entry inside its residue class
  for(unsigned a = M.a_begin + warp; a < M.a_end; a += nwarps)
  {
    unsigned tc = 0;
    while( tc + 1 < ncell && a >= Wc[tc].a_end ) ++tc;
    const NbrWords& W = Wc[tc];
    const unsigned nw = W.n;
    const unsigned* mrow = masks + size_t(a) * mask_stride;
    unsigned short* o16 = idx16 + off[a];
    unsigned* o32 = idx32 + off[a];
    unsigned w = 0;
    for(unsigned q0 = 0; q0 < nw; q0 += 32u)
    {
      const unsigned mine = q0 + lane < nw ? mrow[q0 + lane] : 0u;      // 32 words per coalesced load
      unsigned todo = __ballot_sync(0xffffffffu, mine != 0u);             // most words have no survivor: visit the others only
      while( todo )
      {
        const unsigned t = __ffs(todo) - 1u; todo &= todo - 1u;
        const unsigned m = __shfl_sync(0xffffffffu, mine, t);
        if( m >> lane & 1u )
        {
          const unsigned o = w + __popc(m & lt);
          const unsigned short sv = (unsigned short)(unsigned(W.s[q0 + t]) + lane);
          if( DEAL ) bidx[o] = sv; else o16[o] = sv;
          o32[o] = W.g[q0 + t] + lane;
        }
        w += __popc(m);
      }
    }
    if( DEAL )
    {
      const unsigned L = w;
      if( lane < 16u ) rcnt[warp][lane] = 0u;
      __syncwarp();
      for(unsigned i0 = 0; i0 < L; i0 += 32u)
      {
        const unsigned i = i0 + lane; const bool act = i < L;
        const unsigned res = act ? (unsigned(bidx[i]) & 15u) : 16u + lane;    // idle lanes only match themselves
        const unsigned mm = __match_any_sync(0xffffffffu, res);
        if( act ) brk[i] = (unsigned short)(rcnt[warp][res] + __popc(mm & lt));
        __syncwarp();
        if( act && (mm & lt) == 0u ) rcnt[warp][res] += __popc(mm);
        __syncwarp();
      }
      const unsigned c = lane < 16u ? rcnt[warp][lane] : 0u;
      for(unsigned i0 = 0; i0 < L; i0 += 32u)
      {
        const unsigned i = i0 + lane; const bool act = i < L;
        const unsigned v = act ? bidx[i] : 0u, r = act ? brk[i] : 0u, res = v & 15u;
        unsigned pos = 0;
#       pragma unroll
        for(unsigned q = 0; q < 16u; q++)
        {
          const unsigned cq = __shfl_sync(0xffffffffu, c, q);
          pos += min(cq, r) + ((q < res && cq > r) ? 1u : 0u);
        }
        if( act ) o16[pos] = (unsigned short)v;
      }
      __syncwarp();
    }
  }
}

// ---- export to the reference uint16 stream ---------------------------------------------------------
__device__ __forceinline__ unsigned short encode_cell_index(int ri, int rj, int rk)
{
  return (unsigned short)((ri + 16) | ((rj + 16) << 5) | ((rk + 16) << 10));
}

// per-particle stream length: 1 + 2*groups + chunks
__global__ void stream_len_kernel(unsigned n, const unsigned long long* __restrict__ off, const unsigned* __restrict__ idx,
                                  const unsigned* __restrict__ cell_of, const unsigned* __restrict__ cell_start, int cs_log2,
                                  unsigned long long* __restrict__ len)
{
  const unsigned a = blockIdx.x * blockDim.x + threadIdx.x;
  if( a >= n ) return;
  unsigned groups = 0, chunks = 0, pc = 0xffffffffu, pk = 0xffffffffu;
  for(unsigned long long e = off[a]; e < off[a+1]; e++)
  {
    const unsigned b = idx[e], c = cell_of[b], k = (b - cell_start[c]) >> cs_log2;
    if( c != pc ) { ++groups; ++chunks; pc = c; pk = k; }
    else if( k != pk ) { ++chunks; pk = k; }
  }
  len[a] = 1ull + 2ull * groups + chunks;
}

__global__ void stream_off_kernel(unsigned ncells, const unsigned* __restrict__ cell_start, const unsigned long long* __restrict__ poff,
                                  int has_offsets, unsigned long long* __restrict__ stream_off)
{
  const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
  if( c > ncells ) return;
  const unsigned s = cell_start[c];
  stream_off[c] = poff[s] + (has_offsets ? 2ull * s : 0ull);
}

__global__ void stream_fill_kernel(unsigned n, GridView g, const unsigned long long* __restrict__ off, const unsigned* __restrict__ idx,
                                   const unsigned* __restrict__ cell_of, const unsigned* __restrict__ cell_start, int cs_log2, int has_offsets,
                                   const unsigned long long* __restrict__ poff, const unsigned long long* __restrict__ stream_off,
                                   unsigned short* __restrict__ data)
{
  const unsigned a = blockIdx.x * blockDim.x + threadIdx.x;
  if( a >= n ) return;
  const unsigned ca = cell_of[a], s = cell_start[ca], na = cell_start[ca + 1] - s, p = a - s;
  const unsigned long long rel = poff[a] - poff[s];
  unsigned short* cell_stream = data + stream_off[ca];
  if( has_offsets )
  {
    const unsigned o = 2u * na + unsigned(rel);
    cell_stream[2*p] = (unsigned short)(o & 0xFFFFu); cell_stream[2*p+1] = (unsigned short)(o >> 16);
    cell_stream += 2u * na;
  }
  unsigned short* w = cell_stream + rel;
  unsigned short* groups_pos = w++; unsigned short* nchunks_pos = nullptr;
  unsigned groups = 0, nchunks = 0, pc = 0xffffffffu, pk = 0xffffffffu;
  const int nx = g.nx, ny = g.ny;
  const int ia = int(ca % unsigned(nx)), ja = int((ca / unsigned(nx)) % unsigned(ny)), ka = int(ca / (unsigned(nx) * unsigned(ny)));
  for(unsigned long long e = off[a]; e < off[a+1]; e++)
  {
    const unsigned b = idx[e], c = cell_of[b], k = (b - cell_start[c]) >> cs_log2;
    if( c != pc )
    {
      if( nchunks_pos ) *nchunks_pos = (unsigned short)nchunks;
      const int ib = int(c % unsigned(nx)), jb = int((c / unsigned(nx)) % unsigned(ny)), kb = int(c / (unsigned(nx) * unsigned(ny)));
      *(w++) = encode_cell_index(ib - ia, jb - ja, kb - ka);
      nchunks_pos = w++; nchunks = 0; ++groups; pc = c; pk = 0xffffffffu;
    }
    if( k != pk ) { *(w++) = (unsigned short)k; ++nchunks; pk = k; }
  }
  if( nchunks_pos ) *nchunks_pos = (unsigned short)nchunks;
  *groups_pos = (unsigned short)groups;
}

__global__ void widen_kernel(unsigned n, const unsigned* __restrict__ in, unsigned long long* __restrict__ out)
{
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i < n ) out[i] = in[i];
}

static int exclusive_scan_u64(xsb_ctx* ctx, const unsigned long long* in, unsigned long long* out, size_t n)
{
  size_t tmp = 0;
  XSB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, n, ctx->stream));
  XSB_CUDA(ctx, ctx->scratch.reserve(tmp + 16));
  XSB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->scratch.p, tmp, in, out, n, ctx->stream));
  ctx->launches += 2;
  return XSB_OK;
}

// Tile geometry for this grid + search range: TX cells per tile along x, largest stage over all tiles (host copy of
// the cell offsets).  Returns false when the tile path cannot serve the list (search range > 2 cells in y/z, or a
// stage larger than a uint16 index / the shared-memory budget): the generic CSR kernels are used then.
// bound on |d2_fp32 - d2_exact| of the count sweep for a candidate within 1 % of the cut-off (see nbr_count_kernel):
// coordinates relative to a tile atom are below `ext` in magnitude, FP32 keeps them to 2^-24 * pow2(ext); the
// difference of two of them adds one rounding; the xform (row sums <= mrow) three products and two sums per component.
static NbrF32 nbr_fp32_band(const xsb_grid_desc& g, int TX, const int R[3], double dist)
{
  NbrF32 F{};
  double mrow = 0.0;
  for(int r = 0; r < 3; r++) { double s = 0; for(int c = 0; c < 3; c++) { s += std::fabs(g.xform[3*r + c]); F.m[3*r + c] = float(g.xform[3*r + c]); } mrow = std::max(mrow, s); }
  F.xform = g.xform_is_identity ? 0 : 1;
  if( g.xform_is_identity ) mrow = 1.0;
  const double ext = (TX + 2 * std::max(R[0], std::max(R[1], R[2])) + 1) * g.cell_size;
  double p2 = 1.0; while( p2 < ext ) p2 *= 2.0;
  const double eps = std::ldexp(1.0, -24);
  const double e_comp = 3.0 * eps * p2;                               // two roundings to FP32 + the subtraction
  const double dv = mrow * (e_comp + 8.0 * eps * p2);                 // per physical component, products/sums in FP32 (and the FP32 copy of the matrix)
  const double band = 2.0 * (2.0 * std::sqrt(3.0) * 1.01 * dist * dv + 16.0 * eps * dist * dist);
  F.d2max = float(dist * dist); F.band = float(band + 2.0 * eps * dist * dist);      // + the rounding of d2max itself
  return F;
}

static bool tile_plan_tx(xsb_ctx* ctx, const int R[3], int TX, unsigned align, TileGeom& G, unsigned& s_cap)
{
  const xsb_grid_desc& g = ctx->grid;
  G = TileGeom{};
  G.nx = g.dims[0]; G.ny = g.dims[1]; G.nz = g.dims[2]; G.gl = g.ghost_layers;
  G.Rx = R[0]; G.Ry = R[1]; G.Rz = R[2]; G.TX = TX; G.ghost = 1; G.align = align;
  G.tiles_x = (G.nx + G.TX - 1) / G.TX;
  s_cap = 0;
  const uint64_t am = align - 1;
  if( (2 * R[1] + 1) * (2 * R[2] + 1) > TILE_MAX_ROWS ) return false;
  const std::vector<uint64_t>& off = ctx->h_cell_off;
  // per x-row prefix: atoms in cells [i-Rx, i+TX+Rx) of row (j,k); stage = sum over the rows around (j,k)
  std::vector<unsigned> rowwin(size_t(G.tiles_x) * G.ny * G.nz);
  for(int k = 0; k < G.nz; k++) for(int j = 0; j < G.ny; j++)
  {
    const size_t row = size_t(G.nx) * (size_t(j) + size_t(G.ny) * k);
    for(int ti = 0; ti < G.tiles_x; ti++)
    {
      const int i0 = ti * G.TX, i1 = std::min(G.nx, i0 + G.TX);
      const uint64_t gb = off[row + std::max(0, i0 - G.Rx)], ge = off[row + std::min(G.nx, i1 + G.Rx)];
      rowwin[size_t(ti) + size_t(G.tiles_x) * (size_t(j) + size_t(G.ny) * k)] = ge > gb ? unsigned(((ge + am) & ~am) - (gb & ~am)) : 0u;   // widened, as tile_meta_compute
    }
  }
  for(int k = 0; k < G.nz; k++) for(int j = 0; j < G.ny; j++) for(int ti = 0; ti < G.tiles_x; ti++)
  {
    unsigned S = 0;
    for(int dk = -G.Rz; dk <= G.Rz; dk++) for(int dj = -G.Ry; dj <= G.Ry; dj++)
    {
      const int jj = j + dj, kk = k + dk;
      if( jj < 0 || jj >= G.ny || kk < 0 || kk >= G.nz ) continue;
      S += rowwin[size_t(ti) + size_t(G.tiles_x) * (size_t(jj) + size_t(G.ny) * kk)];
    }
    s_cap = std::max(s_cap, S);
  }
  s_cap = (s_cap + 15u) & ~15u;
  if( s_cap == 0 ) s_cap = 16;
  // 2 stage buffers of x,y,z,w + types must leave room for the operator tables: cap a buffer at 64 KiB; the build
  // kernels index at most 2048 staged atoms (NBR_MAX_WORDS)
  if( s_cap > 2048u || size_t(s_cap) * 33 > 64 * 1024 ) return false;
  G.s_cap = s_cap;
  return true;
}

// Tile width: up to 3 cells per tile when the stage limits allow.  A wider tile stages (TX+2Rx)/TX times fewer
// neighbour cells per central atom and gives the consumer warps more atoms between two mbarrier hand-overs: with one
// cell per tile and ~27 atoms per cell (LJ argon) ncu showed 18 % of the executed instructions in the barrier spin
// loops.  Measured on B200 (profiles/r01x_lj_tile_width.txt), LJ Ar 2 M atoms, pair pass / list build in ms:
// TX=1 1.356 / 5.65, TX=2 1.026 / 5.48, TX=3 0.978 / 5.88, TX=4 1.044 / 6.28, TX=6 1.037 / 7.28 -- beyond 3 the build
// and the balance over the persistent CTAs lose more than the pair pass gains.  Dense cells (EAM Cu, ~58 atoms)
// fill the stage with TX = 1.  XSB_TILE_TX=n forces a width (A/B runs).
static bool tile_plan(xsb_ctx* ctx, const int R[3], unsigned align, TileGeom& G, unsigned& s_cap)
{
  const xsb_grid_desc& g = ctx->grid;
  const char* fixed = getenv("XSB_TILE_TX");
  const int tx_max = fixed ? std::min(NBR_MAX_TX, std::max(1, atoi(fixed))) : 3;
  for(int TX = std::min(tx_max, std::max(1, g.dims[0])); TX >= 1; TX--)
  {
    const size_t tiles = size_t((g.dims[0] + TX - 1) / TX) * g.dims[1] * g.dims[2];
    if( TX > 1 && !fixed && tiles < size_t(ctx->sm_count) * 8 ) continue;
    if( tile_plan_tx(ctx, R, TX, align, G, s_cap) ) return true;
    if( (2 * R[1] + 1) * (2 * R[2] + 1) > TILE_MAX_ROWS ) return false;
  }
  return false;
}

} // namespace xsb

using namespace xsb;

extern "C" {

int xsb_chunk_neighbors_build(xsb_ctx* ctx, double nbh_dist_lab, const xsb_chunk_neighbors_config* cfg)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, ctx->h_cell_off.size() == ctx->ncells + 1, XSB_ERR_STATE, "grid/particles not set");
  XSB_REQUIRE(ctx, nbh_dist_lab > 0.0, XSB_ERR_INVALID, "nbh_dist_lab must be > 0");
  ctx->graph_gen++;
  if( cfg )
  {
    int cs = cfg->chunk_size;
    XSB_REQUIRE(ctx, cs >= 1 && cs <= 32 && (cs & (cs - 1)) == 0, XSB_ERR_INVALID, "chunk_size is not a power of two in 1..32");
    ctx->nbh_cfg = *cfg;
    if( ctx->nbh_cfg.stream_prealloc_factor < 1.0 ) ctx->nbh_cfg.stream_prealloc_factor = 1.0;
  }
  XSB_CUDA(ctx, cudaSetDevice(ctx->device));
  const unsigned n = unsigned(ctx->n);
  ctx->nbh_built = false; ctx->nbh_total = 0; ctx->nbh_max = 0; ctx->nbh_dist = nbh_dist_lab;
  ctx->sub_epoch = 0; ctx->sub_pw_kind = 0;      // a sub-list left by an earlier pass indexes the previous list
  XSB_CUDA(ctx, ctx->nbh_count.reserve(n + 1, XSB_GROW));
  XSB_CUDA(ctx, ctx->nbh_off.reserve(n + 2, XSB_GROW));
  XSB_CUDA(ctx, ctx->scratch64.reserve(n + 2, XSB_GROW));
  XSB_CUDA(ctx, cudaMemsetAsync(ctx->nbh_off.p, 0, 2 * sizeof(unsigned long long), ctx->stream));
  if( n == 0 ) { ctx->nbh_built = true; return XSB_OK; }
  XSB_CUDA(ctx, ctx->tmp64.reserve(16));
  unsigned long long* d2min = ctx->tmp64.p;
  XSB_CUDA(ctx, cudaMemsetAsync(d2min, 0xff, sizeof(unsigned long long), ctx->stream));
  ctx->tile_ok = false;
  ctx->prof_begin(XSB_PROF_NBR_BUILD);
  NbrParams P; P.g = ctx->view(); P.n = n; P.d2max = nbh_dist_lab * nbh_dist_lab;
  int R[3]; search_range(ctx->grid, nbh_dist_lab, R); P.Rx = R[0]; P.Ry = R[1]; P.Rz = R[2];
  {
    // the cells a list can reach must exist around every own cell: with ghost layers thinner than the search range the
    // own atoms next to the ghost border would silently get truncated lists (the reference sizes ghost_layers from
    // ghost_dist / cell_size, so the two always agree there); 15 = range of the 5-bit relative cell index of the stream
    int Ru[3]; search_range_unclamped(ctx->grid, nbh_dist_lab, Ru);
    for(int a = 0; a < 3; a++)
    {
      if( Ru[a] > 15 ) return ctx->fail(XSB_ERR_INVALID, "chunk_neighbors: nbh_dist_lab %g spans %d cells along axis %d (limit 15: relative cell index of the stream)", nbh_dist_lab, Ru[a], a);
      if( ctx->grid.ghost_layers > 0 && Ru[a] > ctx->grid.ghost_layers )
        return ctx->fail(XSB_ERR_INVALID, "chunk_neighbors: nbh_dist_lab %g needs %d cell layers along axis %d but the grid has ghost_layers = %d", nbh_dist_lab, Ru[a], a, ctx->grid.ghost_layers);
    }
  }
  const int block = 256; const unsigned grid = unsigned((uint64_t(n) * 32 + block - 1) / block);
  const double *rx = ctx->f64[XSB_F_RX].p, *ry = ctx->f64[XSB_F_RY].p, *rz = ctx->f64[XSB_F_RZ].p;
  // tile path: the same fixed tiling the force kernels use; one CTA per tile with its 27-cell block staged in shared memory
  TileGeom TG; unsigned s_cap = 0;
  // multi-species system (any type byte != 0): stage rows on 16-atom boundaries so the type bytes travel by TMA as well
  unsigned align = 2;
  {
    size_t tmp = 0; unsigned char* tmax = reinterpret_cast<unsigned char*>(ctx->tmp64.p + 1);
    XSB_CUDA(ctx, cub::DeviceReduce::Max(nullptr, tmp, ctx->type.p, tmax, int(n), ctx->stream));
    XSB_CUDA(ctx, ctx->scratch.reserve(tmp + 16));
    XSB_CUDA(ctx, cub::DeviceReduce::Max(ctx->scratch.p, tmp, ctx->type.p, tmax, int(n), ctx->stream));
    ctx->launches += 1;
    unsigned char h = 0;
    XSB_CUDA(ctx, cudaMemcpyAsync(&h, tmax, 1, cudaMemcpyDeviceToHost, ctx->stream));
    XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if( h != 0 ) align = 16;
  }
  const bool tile = tile_plan(ctx, R, align, TG, s_cap);
  size_t tile_smem = 0;
  unsigned mask_stride = 0;
  if( tile )
  {
    TG.ghost = 1; TG.ti_lo = 0; TG.ti_n = TG.tiles_x; TG.j_lo = 0; TG.j_n = TG.ny; TG.k_lo = 0; TG.k_n = TG.nz;
    TG.ntiles = unsigned(TG.ti_n) * unsigned(TG.j_n) * unsigned(TG.k_n); TG.nbuf = 1;
    tile_smem = (size_t(s_cap) + 32) * 12;      // + 32 slots: lanes past the end of the last word read (and discard) them
    const NbrF32 F32 = nbr_fp32_band(ctx->grid, TG.TX, R, nbh_dist_lab);
    // per device and cheap: set on every build rather than behind a process-global flag (several contexts / GPUs per process)
    cudaFuncSetAttribute(nbr_count_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(nbr_count_kernel<true >, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    // survivor masks of the count sweep (one word per 32-candidate step), replayed by nbr_expand_kernel
    mask_stride = ((s_cap + 31u) / 32u + unsigned((2 * R[1] + 1) * (2 * R[2] + 1)) + 31u) & ~31u;
    XSB_CUDA(ctx, ctx->nbh_masks.reserve(size_t(n) * mask_stride + 32, XSB_GROW));
    const int cblock = 256;      // 512-thread CTAs for large stages measured slower (C2: 6.96 -> 8.99 ms per rebuild)
    if( P.g.xform_identity ) nbr_count_kernel<false><<<TG.ntiles, cblock, tile_smem, ctx->stream>>>(TG, P.g, P.d2max, F32, ctx->cell_start.p, rx, ry, rz, ctx->nbh_count.p, ctx->nbh_masks.p, mask_stride, d2min);
    else                     nbr_count_kernel<true ><<<TG.ntiles, cblock, tile_smem, ctx->stream>>>(TG, P.g, P.d2max, F32, ctx->cell_start.p, rx, ry, rz, ctx->nbh_count.p, ctx->nbh_masks.p, mask_stride, d2min);
  }
  else if( P.g.xform_identity ) nbr_sweep_kernel<false,false><<<grid, block, 0, ctx->stream>>>(P, ctx->cell_start.p, ctx->cell_of.p, rx, ry, rz, ctx->nbh_count.p, nullptr, nullptr, d2min);
  else                          nbr_sweep_kernel<true ,false><<<grid, block, 0, ctx->stream>>>(P, ctx->cell_start.p, ctx->cell_of.p, rx, ry, rz, ctx->nbh_count.p, nullptr, nullptr, d2min);
  XSB_LAUNCH_CHECK(ctx);
  XSB_CUDA(ctx, cudaMemsetAsync(ctx->scratch64.p + n, 0, sizeof(unsigned long long), ctx->stream));
  widen_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, ctx->nbh_count.p, ctx->scratch64.p);
  XSB_LAUNCH_CHECK(ctx);
  int rc = exclusive_scan_u64(ctx, ctx->scratch64.p, ctx->nbh_off.p, size_t(n) + 1); if( rc ) return rc;
  unsigned long long total = 0;
  {
    size_t tmp = 0; unsigned* dmax = reinterpret_cast<unsigned*>(ctx->scratch64.p);   // scratch64 is free again after the scan
    XSB_CUDA(ctx, cub::DeviceReduce::Max(nullptr, tmp, ctx->nbh_count.p, dmax, int(n), ctx->stream));
    XSB_CUDA(ctx, ctx->scratch.reserve(tmp + 16));
    XSB_CUDA(ctx, cub::DeviceReduce::Max(ctx->scratch.p, tmp, ctx->nbh_count.p, dmax, int(n), ctx->stream));
    ctx->launches += 1;
    XSB_CUDA(ctx, cudaMemcpyAsync(&ctx->nbh_max, dmax, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
  }
  XSB_CUDA(ctx, cudaMemcpyAsync(&total, ctx->nbh_off.p + n, sizeof(total), cudaMemcpyDeviceToHost, ctx->stream));
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->nbh_total = total;
  XSB_CUDA(ctx, ctx->nbh_idx.reserve(size_t(total) + 32, ctx->nbh_cfg.stream_prealloc_factor));
  if( tile )
  {
    // tile-local uint16 view for the persistent tile kernels (xsb_tile.cuh) and the CSR view, written together
    XSB_CUDA(ctx, ctx->tl_idx.reserve(size_t(total) + 32, ctx->nbh_cfg.stream_prealloc_factor));
    // uint16 list in canonical order; bank-dealt order only on request (env XSB_TILE_DEAL=1).  Measured at C2 on B200
    // (profiles/r01s_*): wavefronts per position load 3.0 -> 1.41 on the full list (2.14 on a compacted sub-list), but
    // the EAM passes only gain 0.03 ms per step while the deal costs 2.7 ms per rebuild: not worth it for FP64 EAM.
    const unsigned lcap = (ctx->nbh_max + 31u) & ~31u;
    if( lcap && lcap <= 1024u && ctx->tile_deal )
      nbr_expand_kernel<true ><<<TG.ntiles, 256, size_t(8) * 2 * lcap * sizeof(unsigned short), ctx->stream>>>(TG, ctx->cell_start.p, ctx->nbh_off.p, ctx->nbh_masks.p, mask_stride, ctx->nbh_idx.p, ctx->tl_idx.p, lcap);
    else
      nbr_expand_kernel<false><<<TG.ntiles, 256, 0, ctx->stream>>>(TG, ctx->cell_start.p, ctx->nbh_off.p, ctx->nbh_masks.p, mask_stride, ctx->nbh_idx.p, ctx->tl_idx.p, 0u);
    if( ctx->nbh_cfg.free_scratch_memory ) ctx->nbh_masks.release();
  }
  else if( P.g.xform_identity ) nbr_sweep_kernel<false,true><<<grid, block, 0, ctx->stream>>>(P, ctx->cell_start.p, ctx->cell_of.p, rx, ry, rz, nullptr, ctx->nbh_off.p, ctx->nbh_idx.p, nullptr);
  else                          nbr_sweep_kernel<true ,true><<<grid, block, 0, ctx->stream>>>(P, ctx->cell_start.p, ctx->cell_of.p, rx, ry, rz, nullptr, ctx->nbh_off.p, ctx->nbh_idx.p, nullptr);
  XSB_LAUNCH_CHECK(ctx);
  {
    unsigned long long bits = 0;
    XSB_CUDA(ctx, cudaMemcpyAsync(&bits, d2min, sizeof(bits), cudaMemcpyDeviceToHost, ctx->stream));
    const TileGeom& G = TG; const bool ok = tile;
    XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double d2 = 0.0; if( bits != ~0ull ) std::memcpy(&d2, &bits, sizeof(d2));
    ctx->nbh_d2min = d2;
    ctx->tile_ok = ok; ctx->tile_TX = G.TX; ctx->tile_R[0] = R[0]; ctx->tile_R[1] = R[1]; ctx->tile_R[2] = R[2]; ctx->tile_s_cap = s_cap; ctx->tile_align = align;
  }
  ctx->prof_end(XSB_PROF_NBR_BUILD);
  ctx->nbh_built = true;
  return XSB_OK;
}

// host-only view of the FP32 guard band of the count sweep (no context, no GPU): CPU tests sample candidate pairs in
// FP32 and check |d2_fp32 - d2_exact| against it
double xsbdbg_nbr_fp32_band(double cell_size, const double* xform9, int tile_tx, int search_range, double nbh_dist)
{
  xsb_grid_desc g{};
  g.cell_size = cell_size;
  bool ident = true;
  for(int i = 0; i < 9; i++) { g.xform[i] = xform9 ? xform9[i] : ((i % 4 == 0) ? 1.0 : 0.0); ident = ident && g.xform[i] == ((i % 4 == 0) ? 1.0 : 0.0); }
  g.xform_is_identity = ident ? 1 : 0;
  const int R[3] = { search_range, search_range, search_range };
  return double(nbr_fp32_band(g, tile_tx, R, nbh_dist).band);
}

int xsb_chunk_neighbors_stats(xsb_ctx* ctx, uint64_t* total, uint32_t* maxn)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, ctx->nbh_built, XSB_ERR_STATE, "chunk_neighbors not built");
  if( total ) *total = ctx->nbh_total;
  if( maxn ) *maxn = ctx->nbh_max;
  return XSB_OK;
}

static int export_prepare(xsb_ctx* ctx, DevBuf<unsigned long long>& poff, DevBuf<unsigned long long>& soff, uint64_t* total_u16)
{
  XSB_REQUIRE(ctx, ctx->nbh_built, XSB_ERR_STATE, "chunk_neighbors not built");
  XSB_CUDA(ctx, cudaSetDevice(ctx->device));
  const unsigned n = unsigned(ctx->n), nc = unsigned(ctx->ncells);
  int cs_log2 = 0; while( (1 << cs_log2) < ctx->nbh_cfg.chunk_size ) ++cs_log2;
  XSB_CUDA(ctx, poff.reserve(size_t(n) + 2));
  XSB_CUDA(ctx, soff.reserve(size_t(nc) + 2));
  XSB_CUDA(ctx, ctx->scratch64.reserve(size_t(n) + 2));
  XSB_CUDA(ctx, cudaMemsetAsync(ctx->scratch64.p, 0, (size_t(n) + 1) * sizeof(unsigned long long), ctx->stream));
  if( n )
  {
    stream_len_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, ctx->nbh_off.p, ctx->nbh_idx.p, ctx->cell_of.p, ctx->cell_start.p, cs_log2, ctx->scratch64.p);
    XSB_LAUNCH_CHECK(ctx);
  }
  int rc = exclusive_scan_u64(ctx, ctx->scratch64.p, poff.p, size_t(n) + 1); if( rc ) return rc;
  stream_off_kernel<<<(nc + 1 + 255) / 256, 256, 0, ctx->stream>>>(nc, ctx->cell_start.p, poff.p, ctx->nbh_cfg.build_particle_offset, soff.p);
  XSB_LAUNCH_CHECK(ctx);
  unsigned long long tot = 0;
  XSB_CUDA(ctx, cudaMemcpyAsync(&tot, soff.p + nc, sizeof(tot), cudaMemcpyDeviceToHost, ctx->stream));
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *total_u16 = tot;
  return XSB_OK;
}

int xsb_chunk_neighbors_export_size(xsb_ctx* ctx, uint64_t* total_u16)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, total_u16 != nullptr, XSB_ERR_INVALID, "null output");
  DevBuf<unsigned long long> poff, soff;
  int rc = export_prepare(ctx, poff, soff, total_u16);
  poff.release(); soff.release();
  return rc;
}

int xsb_chunk_neighbors_export(xsb_ctx* ctx, uint64_t* stream_off, uint16_t* data)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, stream_off != nullptr && data != nullptr, XSB_ERR_INVALID, "null output");
  DevBuf<unsigned long long> poff, soff; DevBuf<unsigned short> dd;
  uint64_t tot = 0;
  int rc = export_prepare(ctx, poff, soff, &tot);
  if( rc == XSB_OK )
  {
    const unsigned n = unsigned(ctx->n);
    int cs_log2 = 0; while( (1 << cs_log2) < ctx->nbh_cfg.chunk_size ) ++cs_log2;
    cudaError_t e = dd.reserve(tot + 16);
    if( e != cudaSuccess ) rc = ctx->fail(XSB_ERR_CUDA, "export buffer: %s", cudaGetErrorString(e));
    else
    {
      if( n )
      {
        stream_fill_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, ctx->view(), ctx->nbh_off.p, ctx->nbh_idx.p, ctx->cell_of.p, ctx->cell_start.p, cs_log2,
                                                                      ctx->nbh_cfg.build_particle_offset, poff.p, soff.p, dd.p);
        ctx->launches++;
      }
      e = cudaMemcpyAsync(stream_off, soff.p, (ctx->ncells + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
      if( e == cudaSuccess && tot ) e = cudaMemcpyAsync(data, dd.p, tot * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream);
      if( e == cudaSuccess ) e = cudaStreamSynchronize(ctx->stream);
      if( e != cudaSuccess ) rc = ctx->fail(XSB_ERR_CUDA, "export: %s", cudaGetErrorString(e));
    }
  }
  poff.release(); soff.release(); dd.release();
  return rc;
}

int xsb_chunk_neighbors_download_flat(xsb_ctx* ctx, uint32_t* counts, uint64_t* offsets, uint32_t* idx)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, ctx->nbh_built, XSB_ERR_STATE, "chunk_neighbors not built");
  if( counts && ctx->n ) XSB_CUDA(ctx, cudaMemcpyAsync(counts, ctx->nbh_count.p, ctx->n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  if( offsets ) XSB_CUDA(ctx, cudaMemcpyAsync(offsets, ctx->nbh_off.p, (ctx->n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  if( idx && ctx->nbh_total ) XSB_CUDA(ctx, cudaMemcpyAsync(idx, ctx->nbh_idx.p, ctx->nbh_total * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return XSB_OK;
}

} // extern "C"
