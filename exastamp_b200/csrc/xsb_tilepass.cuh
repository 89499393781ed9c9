// xsb_tilepass.cuh -- one persistent kernel template for every pair pass (LJ force, EAM rho, EAM force, ...):
// the operator-specific arithmetic is a small functor ("Op", the analogue of the reference's ForceOp / EmbOp
// functors handed to exanb::compute_cell_particle_pairs), the traversal is xsb_tile.cuh.
//
// Op interface
//   static constexpr bool HAS_W, TYPES;          stage the per-atom scalar w / the type bytes next to x,y,z
//   double rcut2;                                inclusive cutoff test d2 <= rcut2 (compute_pair_rigidmol.h:262)
//   size_t table_bytes() const;                  operator tables kept in shared memory for the whole launch
//   void   load_tables(unsigned char* smem, int nt);   cooperative fill by the whole CTA (caller barriers)
//   struct Acc;  void init(Acc&);  void start(Acc&, a, sa, B, tab);  void pair(Acc&, dx,dy,dz,d2, j, B, tab);   (tab = smem tables)
//   template<int TPA> void finish(Acc&, a, valid, sub);   group reduction + the single writer's stores
#pragma once
#include "xsb_tile.cuh"

namespace xsb
{

constexpr size_t TILE_SMEM_MAX = 227 * 1024;

template<int TPA, int NT, bool XFORM, class Op>
__global__ void __launch_bounds__(NT, 1) tile_pass_kernel(const TileGeom G, const unsigned* __restrict__ cell_start, const TileFields F, const TileList L,
                                                          const XForm X, const Op op)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const size_t tb = (op.table_bytes() + 15) & ~size_t(15);
  op.load_tables(smem, NT);
  TileMeta* meta = reinterpret_cast<TileMeta*>(smem + tb);
  unsigned char* stage_mem = smem + tb + ((3 * sizeof(TileMeta) + 15) & ~size_t(15));
  // (the first barrier inside tile_loop also publishes the tables)
  constexpr unsigned NG = NT / TPA;
  const unsigned g = threadIdx.x / TPA, sub = threadIdx.x % TPA;
  tile_loop<Op::HAS_W, Op::TYPES, NT>(G, cell_start, F, stage_mem, meta, [&](const TileMeta& M, const StageBuf<Op::HAS_W, Op::TYPES>& B)
  {
    const unsigned a_begin = M.a_begin, a_end = M.a_end, c_off = M.c_off;
    for(unsigned base = a_begin; base < a_end; base += NG)
    {
      const unsigned a = base + g;
      const bool valid = a < a_end;
      typename Op::Acc acc;
      op.init(acc);
      if( valid )
      {
        const unsigned sa = a + c_off;
        const double xa = B.x[sa], ya = B.y[sa], za = B.z[sa];
        op.start(acc, a, sa, B, smem);
        const unsigned long long e1 = L.off[a + 1];
        for(unsigned long long e = L.off[a] + sub; e < e1; e += TPA)
        {
          const unsigned j = __ldcs(L.idx + e);
          double dx = B.x[j] - xa, dy = B.y[j] - ya, dz = B.z[j] - za;
          apply_xform<XFORM>(X, dx, dy, dz);
          const double d2 = dx * dx + dy * dy + dz * dz;
          if( d2 <= op.rcut2 ) op.pair(acc, dx, dy, dz, d2, j, B, smem);
        }
      }
      op.template finish<TPA>(acc, a, valid, sub);
    }
  });
}

template<int TPA, int NT, class Op>
static int launch_tile_pass(xsb_ctx* ctx, bool ghost, const Op& op, const double* w)
{
  const TileGeom G = make_tile_geom(ctx, ghost);
  if( G.ntiles == 0 ) return XSB_OK;
  const size_t smem = tile_smem_bytes<Op::HAS_W, Op::TYPES>(G.s_cap, op.table_bytes());
  XSB_REQUIRE(ctx, smem <= TILE_SMEM_MAX, XSB_ERR_STATE, "tile pass: shared-memory plan exceeds 227 KiB (caller must pick a smaller table window)");
  const TileFields F{ ctx->f64[XSB_F_RX].p, ctx->f64[XSB_F_RY].p, ctx->f64[XSB_F_RZ].p, w, ctx->type.p };
  const TileList L{ ctx->nbh_off.p, ctx->tl_idx.p };
  const XForm X = make_xform(ctx->grid);
  const bool xf = !ctx->grid.xform_is_identity;
  auto go = [&](auto kern) -> int
  {
    XSB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int per_sm = 1;
    XSB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    if( per_sm < 1 ) per_sm = 1;
    const unsigned grid = std::min<unsigned>(G.ntiles, unsigned(ctx->sm_count * per_sm));
    kern<<<grid, NT, smem, ctx->stream>>>(G, ctx->cell_start.p, F, L, X, op);
    XSB_LAUNCH_CHECK(ctx);
    return XSB_OK;
  };
  return xf ? go(tile_pass_kernel<TPA, NT, true, Op>) : go(tile_pass_kernel<TPA, NT, false, Op>);
}

// 9-component virial accumulator shared by the force ops: vir += -1/2 f (x) dr, Mat3d row-major (ext tensor())
struct Vir9
{
  double v[9];
  __device__ __forceinline__ void zero() {
#   pragma unroll
    for(int k = 0; k < 9; k++) v[k] = 0.0; }
  __device__ __forceinline__ void add(double fx, double fy, double fz, double dx, double dy, double dz)
  {
    v[0] -= 0.5 * fx * dx; v[1] -= 0.5 * fx * dy; v[2] -= 0.5 * fx * dz;
    v[3] -= 0.5 * fy * dx; v[4] -= 0.5 * fy * dy; v[5] -= 0.5 * fy * dz;
    v[6] -= 0.5 * fz * dx; v[7] -= 0.5 * fz * dy; v[8] -= 0.5 * fz * dz;
  }
  template<int TPA> __device__ __forceinline__ void reduce() {
#   pragma unroll
    for(int k = 0; k < 9; k++) v[k] = group_sum<TPA>(v[k]); }
  __device__ __forceinline__ void store_add(double* vir, unsigned a) const { double* p = vir + 9ull * a;
#   pragma unroll
    for(int k = 0; k < 9; k++) p[k] += v[k]; }
};

} // namespace xsb
