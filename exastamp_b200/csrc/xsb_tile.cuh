// xsb_tile.cuh -- the B200 traversal engine behind every pair-pass operator (SURVEY.md 8a row a3,
// exanb::compute_cell_particle_pairs): persistent CTAs, one per SM, walk the cell grid tile by tile.
//
//   tile      = TX consecutive cells of one x-row (j,k); its central atoms are one contiguous flat range.
//   stage     = the (TX+2Rx) x (2Ry+1) x (2Rz+1) block of cells around the tile.  Cells of one x-row are
//               contiguous in the flat SoA, so the stage is (2Ry+1)(2Rz+1) contiguous runs ("rows") of atoms.
//               A dedicated producer warp copies them global -> shared with 1-D TMA bulk copies
//               (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) into a ring of stage buffers; rows are widened
//               to 16-byte boundaries (the pad slots hold real neighbouring atoms and are never indexed).
//               Consumer warps never meet at a CTA barrier: full/empty mbarriers per buffer, and a per-buffer
//               atom cursor from which warps grab central atoms, so a warp that runs out of work in tile n
//               starts on tile n+1 while slower warps finish.
//   list      = per central atom, uint16 indices INTO THE STAGE (ascending = the canonical (cell_b,p_b) order of
//               the reference stream), built once per chunk_neighbors call (xsb_nbr.cu).  2 B/entry of HBM
//               traffic instead of a 4-byte global index plus 3-4 scattered 8-byte gathers through L1/L2.
//   traversal = a group of TPA consecutive lanes owns one central atom and strides over its list; positions
//               (and the per-atom scalar w, e.g. rho_dEmb) come from shared memory; partial sums are combined
//               with xor-shuffles (Newton-off: every atom is written by exactly one group, no atomics).
#pragma once
#include "xsb_ctx.h"
#include "xsb_traverse.cuh"

namespace xsb
{

constexpr int TILE_MAX_ROWS = 25;      // (2Ry+1)(2Rz+1) <= 25  <=>  search range R <= 2 cells in y and z

struct TileGeom
{
  int nx, ny, nz, gl;
  int TX, Rx, Ry, Rz;
  int tiles_x;                 // tiles per x-row = ceil(nx / TX)
  // enumeration window (which tiles a launch visits): ghost=true -> everything, else the non-ghost interior
  int ti_lo, ti_n, j_lo, j_n, k_lo, k_n;
  unsigned ntiles;             // ti_n * j_n * k_n
  unsigned s_cap;              // stage capacity (atoms) of one buffer
  int nbuf;                    // stage buffers in the ring (2 or 3)
  int ghost;                   // central atoms of ghost cells are processed too
  unsigned align;              // staged rows start / end on multiples of `align` atoms: 2 (16-byte doubles for TMA), or 16
                               // when the list was built for a multi-species system, so the type bytes can be TMA-copied too
};

struct TileMeta                // lives in shared memory, one per stage buffer
{
  unsigned a_begin, a_end;     // flat range of central atoms to process
  unsigned c_off;              // stage index of central atom a = a + c_off (unsigned wrap-around arithmetic)
  unsigned S;                  // atoms in the stage
  unsigned nrows;
  unsigned g0[TILE_MAX_ROWS];      // flat index of the first staged atom of each row (even: 16-byte aligned doubles)
  unsigned s0[TILE_MAX_ROWS + 1];  // stage index of the first atom of each row (even); s0[nrows] = S
};

// executed by ONE FULL WARP (all 32 lanes converged).  (ti,j,k) = tile coordinates in the fixed tiling of the grid.
__device__ __forceinline__ void tile_meta_compute(const TileGeom& G, const unsigned* __restrict__ cell_start, int ti, int j, int k, TileMeta& M)
{
  const int lane = threadIdx.x & 31;
  const int i0 = ti * G.TX, i1 = min(G.nx, i0 + G.TX);
  const int nry = 2 * G.Ry + 1, nrows = nry * (2 * G.Rz + 1);
  unsigned len = 0, g = 0;
  if( lane < nrows )
  {
    const int kk = k + lane / nry - G.Rz, jj = j + lane % nry - G.Ry;
    if( jj >= 0 && jj < G.ny && kk >= 0 && kk < G.nz )
    {
      const unsigned row = unsigned(G.nx) * (unsigned(jj) + unsigned(G.ny) * unsigned(kk));
      const unsigned gb = cell_start[row + max(0, i0 - G.Rx)], ge = cell_start[row + min(G.nx, i1 + G.Rx)];
      const unsigned am = G.align - 1u;
      if( ge > gb ) { g = gb & ~am; len = ((ge + am) & ~am) - g; }   // widen to aligned bounds: TMA bulk copies need 16-byte alignment
    }
  }
  unsigned s = len;
# pragma unroll
  for(int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, s, o); if( lane >= o ) s += v; }
  const unsigned S = __shfl_sync(0xffffffffu, s, 31);
  if( lane < nrows ) { M.g0[lane] = g; M.s0[lane] = s - len; }
  if( lane == nrows - 1 ) M.s0[nrows] = S;
  if( lane == G.Rz * nry + G.Ry )   // the row that holds the tile itself
  {
    int ic0 = i0, ic1 = i1;
    bool empty = false;
    if( !G.ghost )
    {
      ic0 = max(i0, G.gl); ic1 = min(i1, G.nx - G.gl);
      empty = ic0 >= ic1 || j < G.gl || j >= G.ny - G.gl || k < G.gl || k >= G.nz - G.gl;
    }
    const unsigned row = unsigned(G.nx) * (unsigned(j) + unsigned(G.ny) * unsigned(k));
    const unsigned ab = empty ? 0u : cell_start[row + ic0], ae = empty ? 0u : cell_start[row + ic1];
    M.a_begin = ab; M.a_end = ae;
    M.c_off = (s - len) - g;
    M.S = (ab == ae) ? 0u : S;      // nothing to compute -> nothing to stage
    M.nrows = unsigned(nrows);
  }
}

// n-th tile visited by a launch -> coordinates in the fixed tiling
__device__ __forceinline__ void tile_coords(const TileGeom& G, unsigned t, int& ti, int& j, int& k)
{
  ti = G.ti_lo + int(t % unsigned(G.ti_n));
  const unsigned u = t / unsigned(G.ti_n);
  j = G.j_lo + int(u % unsigned(G.j_n));
  k = G.k_lo + int(u / unsigned(G.j_n));
}

// ---- mbarrier + TMA bulk copy (PTX; SASS: SYNCS.*, UBLKCP) --------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return unsigned(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{ asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }\n" :: "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes)
{ asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
  asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n"
               :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, bytes a multiple of 16, both addresses 16-byte aligned; completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// stage buffer: SoA doubles x,y,z,(w) of s_cap atoms each, then s_cap type bytes (when TYPES)
template<bool HAS_W, bool TYPES>
struct StageBuf
{
  double *x, *y, *z, *w; unsigned char* t;
  static __host__ __device__ size_t bytes(unsigned s_cap)
  {
    size_t b = size_t(s_cap) * 8 * (HAS_W ? 4 : 3);
    if( TYPES ) b += (size_t(s_cap) + 15) & ~size_t(15);
    return b;
  }
  __device__ __forceinline__ void bind(unsigned char* base, unsigned s_cap)
  {
    x = reinterpret_cast<double*>(base); y = x + s_cap; z = y + s_cap;
    w = HAS_W ? z + s_cap : nullptr;
    t = TYPES ? base + size_t(s_cap) * 8 * (HAS_W ? 4 : 3) : nullptr;
  }
};

struct TileFields { const double* __restrict__ rx; const double* __restrict__ ry; const double* __restrict__ rz; const double* __restrict__ w; const unsigned char* __restrict__ type; };

// ring of stage buffers in shared memory
constexpr int TILE_MAX_BUFS = 3;
struct TileRing
{
  unsigned long long full[TILE_MAX_BUFS];    // mbarrier: the stage of buffer b has landed (producer arrive + TMA bytes)
  unsigned long long empty[TILE_MAX_BUFS];   // mbarrier: every consumer warp is done with buffer b
  unsigned cursor[TILE_MAX_BUFS];            // next central atom (offset from a_begin) to hand out
  unsigned pad_;
  TileMeta meta[TILE_MAX_BUFS];
};

// Producer side, executed by one full warp: publish tile n = (ti,j,k) into buffer b.
template<bool HAS_W, bool TYPES>
__device__ __forceinline__ void tile_produce(const TileGeom& G, const unsigned* __restrict__ cell_start, const TileFields& F, TileRing& R, int b,
                                             StageBuf<HAS_W, TYPES>& B, int ti, int j, int k)
{
  const int lane = threadIdx.x & 31;
  TileMeta& M = R.meta[b];
  tile_meta_compute(G, cell_start, ti, j, k, M);
  __syncwarp();
  const unsigned S = M.S, nrows = M.nrows;
  const bool types_tma = TYPES && G.align >= 16u;
  if( TYPES && !types_tma )
  {
    // rows only aligned to 2 atoms (list built before the types became multi-species): byte copy row by row
    for(unsigned r = 0; r < nrows; r++)
    {
      const unsigned s0 = M.s0[r], cnt = M.s0[r + 1] - s0, g = M.g0[r];
      for(unsigned i = lane; i < cnt; i += 32) B.t[s0 + i] = F.type[g + i];
    }
  }
  __syncwarp();      // the byte-copied types of all lanes are ordered before lane 0's releasing arrive
  if( lane == 0 )
  {
    R.cursor[b] = 0;
    mbar_arrive_expect_tx(&R.full[b], S * 8u * (HAS_W ? 4u : 3u) + (types_tma ? S : 0u));    // releases meta, cursor (and byte-copied types after the syncwarp)
  }
  __syncwarp();
  if( S && unsigned(lane) < nrows )
  {
    const unsigned s0 = M.s0[lane], cnt = M.s0[lane + 1] - s0, g = M.g0[lane];
    if( cnt )
    {
      bulk_g2s(B.x + s0, F.rx + g, cnt * 8u, &R.full[b]);
      bulk_g2s(B.y + s0, F.ry + g, cnt * 8u, &R.full[b]);
      bulk_g2s(B.z + s0, F.rz + g, cnt * 8u, &R.full[b]);
      if( HAS_W ) bulk_g2s(B.w + s0, F.w + g, cnt * 8u, &R.full[b]);
      if( types_tma ) bulk_g2s(B.t + s0, F.type + g, cnt, &R.full[b]);
    }
  }
}

// shared-memory carve-up used by the host to size launches: [user tables][ring][queues][stage x nbuf]
template<bool HAS_W, bool TYPES>
inline size_t tile_smem_bytes(unsigned s_cap, size_t table_bytes, size_t queue_bytes, int nbuf)
{
  const size_t bb = (StageBuf<HAS_W, TYPES>::bytes(s_cap) + 127) & ~size_t(127);
  return ((table_bytes + 127) & ~size_t(127)) + ((sizeof(TileRing) + 127) & ~size_t(127)) + ((queue_bytes + 127) & ~size_t(127)) + size_t(nbuf) * bb;
}

// launch-side geometry: the fixed tiling recorded by chunk_neighbors + the window of tiles this launch visits
inline TileGeom make_tile_geom(const xsb_ctx* ctx, bool ghost)
{
  TileGeom G{};
  const xsb_grid_desc& g = ctx->grid;
  G.nx = g.dims[0]; G.ny = g.dims[1]; G.nz = g.dims[2]; G.gl = g.ghost_layers;
  G.TX = ctx->tile_TX; G.Rx = ctx->tile_R[0]; G.Ry = ctx->tile_R[1]; G.Rz = ctx->tile_R[2]; G.align = ctx->tile_align;
  G.tiles_x = (G.nx + G.TX - 1) / G.TX;
  G.ghost = ghost ? 1 : 0;
  G.s_cap = ctx->tile_s_cap;
  if( ghost ) { G.ti_lo = 0; G.ti_n = G.tiles_x; G.j_lo = 0; G.j_n = G.ny; G.k_lo = 0; G.k_n = G.nz; }
  else
  {
    G.ti_lo = G.gl / G.TX; G.ti_n = (G.nx - G.gl - 1) / G.TX - G.ti_lo + 1;
    G.j_lo = G.gl; G.j_n = G.ny - 2 * G.gl; G.k_lo = G.gl; G.k_n = G.nz - 2 * G.gl;
  }
  G.ntiles = unsigned(G.ti_n) * unsigned(G.j_n) * unsigned(G.k_n);
  return G;
}

struct TileList
{
  const unsigned long long* __restrict__ off;   // [n+1] entry offsets (shared with the CSR list)
  const unsigned short* __restrict__ idx;       // stage indices
  // in-range sub-list of the current positions: the entries with d2 <= rcut2, compacted in place (same offsets),
  // written by the first pair pass of a step (EAM rho / emb) and consumed by the second (EAM force)
  unsigned short* __restrict__ sub_idx;
  unsigned* __restrict__ sub_cnt;               // [n]
  double* __restrict__ pair_w;                  // per-pair cache aligned with sub_idx (xsb_tilepass.cuh), may be null
  size_t pw_plane;                              // offset (elements) of the second cached value of a pair, when an Op keeps two
  // inner skin (xsb_eam.cu): the sub-list keeps the entries with d2 <= list_rc2 >= rcut2, so that it stays a superset of the
  // in-range pairs for a few steps; entries beyond rcut2 carry a NaN in pair_w ("not live") and contribute nothing
  double list_rc2;
  // device-side mode switch: a launch whose want_mode differs from *mode returns at once (two launches per step, one runs)
  const int* __restrict__ mode; int want_mode;
};

// LIST_SUB_REWRITE: walk the sub-list left by an earlier step (no compaction), re-evaluate every entry on the current
// positions and rewrite its pair_w value (NaN when the pair is out of range now)
enum { LIST_FULL = 0, LIST_FULL_WRITE_SUB = 1, LIST_SUB = 2, LIST_SUB_REWRITE = 3 };

__device__ __forceinline__ double pw_dead() { return __longlong_as_double(0x7ff8000000000000ll); }      // quiet NaN: pair outside rcut

} // namespace xsb
