// xsb_tile.cuh -- the B200 traversal engine behind every pair-pass operator (SURVEY.md 8a row a3,
// exanb::compute_cell_particle_pairs): persistent CTAs, one per SM, walk the cell grid tile by tile.
//
//   tile      = TX consecutive cells of one x-row (j,k); its central atoms are one contiguous flat range.
//   stage     = the (TX+2Rx) x (2Ry+1) x (2Rz+1) block of cells around the tile.  Cells of one x-row are
//               contiguous in the flat SoA, so the stage is (2Ry+1)(2Rz+1) contiguous runs ("rows") of atoms;
//               they are copied global -> shared with cp.async (LDGSTS, 8-byte granules: no alignment
//               constraint on the runs), double buffered: tile n+1 lands while tile n is computed.
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
  int ghost;                   // central atoms of ghost cells are processed too
};

struct TileMeta                // lives in shared memory, one per stage buffer
{
  unsigned a_begin, a_end;     // flat range of central atoms to process
  unsigned c_off;              // stage index of central atom a = a + c_off (unsigned wrap-around arithmetic)
  unsigned S;                  // atoms in the stage
  unsigned nrows;
  unsigned g0[TILE_MAX_ROWS];      // flat index of the first atom of each row
  unsigned s0[TILE_MAX_ROWS + 1];  // stage index of the first atom of each row; s0[nrows] = S
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
      g = cell_start[row + max(0, i0 - G.Rx)];
      len = cell_start[row + min(G.nx, i1 + G.Rx)] - g;
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

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc)
{
  const unsigned d = unsigned(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template<int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

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

// all threads of the CTA: enqueue the asynchronous copies of one stage (no wait)
template<bool HAS_W, bool TYPES, int NT>
__device__ __forceinline__ void stage_issue(const TileMeta& M, const TileFields& F, StageBuf<HAS_W, TYPES>& B)
{
  const unsigned S = M.S;
  for(unsigned s = threadIdx.x; s < S; s += NT)
  {
    unsigned r = 0;
    while( s >= M.s0[r + 1] ) ++r;
    const unsigned g = M.g0[r] + (s - M.s0[r]);
    cp_async8(B.x + s, F.rx + g); cp_async8(B.y + s, F.ry + g); cp_async8(B.z + s, F.rz + g);
    if( HAS_W ) cp_async8(B.w + s, F.w + g);
    if( TYPES ) B.t[s] = F.type[g];     // 1-byte granule: plain load/store (visible after the next barrier)
  }
}

// The persistent tile loop.  `body(M, B)` is called by every thread of the CTA once per non-empty tile with the
// stage resident in shared memory; it must not return early (barriers follow).
template<bool HAS_W, bool TYPES, int NT, class Body>
__device__ __forceinline__ void tile_loop(const TileGeom& G, const unsigned* __restrict__ cell_start, const TileFields& F,
                                          unsigned char* stage_mem /* 2 buffers */, TileMeta* meta /* [3] in smem */, Body body)
{
  const size_t bb = (StageBuf<HAS_W, TYPES>::bytes(G.s_cap) + 15) & ~size_t(15);
  auto stage = [&](int b) { StageBuf<HAS_W, TYPES> B; B.bind(stage_mem + size_t(b) * bb, G.s_cap); return B; };
  // meta slots rotate mod 3: the slot written at the top of iteration n was last read by the body of iteration
  // n-2, which every thread has left before passing the barriers of iteration n-1 (stage buffers rotate mod 2).
  const bool w0 = threadIdx.x < 32;
  unsigned t = blockIdx.x;
  if( t >= G.ntiles ) return;
  if( w0 ) { int ti, j, k; tile_coords(G, t, ti, j, k); tile_meta_compute(G, cell_start, ti, j, k, meta[0]); }
  __syncthreads();
  { StageBuf<HAS_W, TYPES> B = stage(0); stage_issue<HAS_W, TYPES, NT>(meta[0], F, B); }
  cp_async_commit();
  int cur = 0, mcur = 0;
  for(; t < G.ntiles; t += gridDim.x)
  {
    const unsigned tn = t + gridDim.x;
    const int nxt = cur ^ 1, mnxt = mcur == 2 ? 0 : mcur + 1;
    if( tn < G.ntiles && w0 ) { int ti, j, k; tile_coords(G, tn, ti, j, k); tile_meta_compute(G, cell_start, ti, j, k, meta[mnxt]); }
    __syncthreads();                       // meta[mnxt] visible; every thread is done computing on buf[nxt]
    if( tn < G.ntiles ) { StageBuf<HAS_W, TYPES> B = stage(nxt); stage_issue<HAS_W, TYPES, NT>(meta[mnxt], F, B); }
    cp_async_commit();
    cp_async_wait<1>();                    // this thread's copies of the CURRENT tile have landed
    __syncthreads();                       // ... and everybody else's
    if( meta[mcur].a_begin < meta[mcur].a_end ) { StageBuf<HAS_W, TYPES> B = stage(cur); body(meta[mcur], B); }
    cur = nxt; mcur = mnxt;
  }
  cp_async_wait<0>();
}

// shared-memory carve-up used by the host to size launches: [user tables][meta x3][stage x2]
template<bool HAS_W, bool TYPES>
inline size_t tile_smem_bytes(unsigned s_cap, size_t table_bytes)
{
  const size_t bb = (StageBuf<HAS_W, TYPES>::bytes(s_cap) + 15) & ~size_t(15);
  return ((table_bytes + 15) & ~size_t(15)) + ((3 * sizeof(TileMeta) + 15) & ~size_t(15)) + 2 * bb;
}

// launch-side geometry: the fixed tiling recorded by chunk_neighbors + the window of tiles this launch visits
inline TileGeom make_tile_geom(const xsb_ctx* ctx, bool ghost)
{
  TileGeom G{};
  const xsb_grid_desc& g = ctx->grid;
  G.nx = g.dims[0]; G.ny = g.dims[1]; G.nz = g.dims[2]; G.gl = g.ghost_layers;
  G.TX = ctx->tile_TX; G.Rx = ctx->tile_R[0]; G.Ry = ctx->tile_R[1]; G.Rz = ctx->tile_R[2];
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
};

} // namespace xsb
