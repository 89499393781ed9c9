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
// Optional per-pair cache (TileList::pair_w, PW_N = 1 or 2 doubles per in-range sub-list entry): an Op with PW_OUT = true
// returns from pair_d2() what the second pass of the step needs again for the same pair (EAM: rho'(r); double or double2),
// the queue path stores it next to the sub-list entry (coalesced 256-byte stores); an Op with PW_IN = true receives it in
// pair_pw(..., pv, pv2) when it walks that sub-list, instead of repeating the spline / transcendental evaluation.
// NO_AHEAD = true opts an Op out of the one-atom-ahead list-window prefetch (ops at the register limit).
#pragma once
#include "xsb_tile.cuh"
#include <type_traits>

namespace xsb
{

constexpr size_t TILE_SMEM_MAX = 227 * 1024;

// per-warp compaction queue of the d2-only passes: 64 slots of {d2, stage index}
constexpr int TILE_QUEUE_SLOTS = 64;
template<int NT> constexpr size_t tile_queue_bytes() { return size_t(NT / 32) * TILE_QUEUE_SLOTS * (sizeof(double) + sizeof(unsigned short)); }

template<class Op, class = void> struct op_pw_out : std::false_type {};
template<class Op> struct op_pw_out<Op, std::void_t<decltype(Op::PW_OUT)>> : std::bool_constant<Op::PW_OUT> {};
// number of cached values per pair (planes of TileList::pair_w): 1 unless the Op says PW_N = 2
template<class Op, class = void> struct op_pw_n : std::integral_constant<int, 1> {};
template<class Op> struct op_pw_n<Op, std::void_t<decltype(Op::PW_N)>> : std::integral_constant<int, Op::PW_N> {};
__device__ __forceinline__ void pw_store(double* __restrict__ pw, size_t, unsigned k, double v) { pw[k] = v; }
__device__ __forceinline__ void pw_store(double* __restrict__ pw, size_t plane, unsigned k, double2 v) { pw[k] = v.x; pw[plane + k] = v.y; }
// an Op at the register limit of its launch shape opts out of the one-atom-ahead window prefetch (NO_AHEAD = true)
template<class Op, class = void> struct op_no_ahead : std::false_type {};
template<class Op> struct op_no_ahead<Op, std::void_t<decltype(Op::NO_AHEAD)>> : std::bool_constant<Op::NO_AHEAD> {};
template<class Op, class = void> struct op_pw_in : std::false_type {};
template<class Op> struct op_pw_in<Op, std::void_t<decltype(Op::PW_IN)>> : std::bool_constant<Op::PW_IN> {};

template<int TPA, int NT, bool XFORM, int LMODE, class Op>
__global__ void __launch_bounds__(NT, 1) tile_pass_kernel(const TileGeom G, const unsigned* __restrict__ cell_start, const TileFields F, const TileList L,
                                                          const XForm X, const Op op)
{
  extern __shared__ __align__(128) unsigned char smem[];
  if( L.mode != nullptr && *L.mode != L.want_mode ) return;      // the other launch of this step serves the current mode
  constexpr bool QUEUE = Op::D2_ONLY && TPA == 32 && (LMODE == LIST_FULL_WRITE_SUB || LMODE == LIST_FULL);
  constexpr bool REWRITE = LMODE == LIST_SUB_REWRITE;
  constexpr bool PWO = op_pw_out<Op>::value && QUEUE && LMODE == LIST_FULL_WRITE_SUB;
  constexpr bool PWI = op_pw_in<Op>::value && LMODE == LIST_SUB;
  constexpr unsigned NWC = NT / 32 - 1;          // consumer warps; the last warp is the TMA producer
  constexpr unsigned GPW = 32 / TPA;             // central atoms per warp and grab
  typedef StageBuf<Op::HAS_W, Op::TYPES> Stage;
  const size_t tb = (op.table_bytes() + 127) & ~size_t(127);
  TileRing& R = *reinterpret_cast<TileRing*>(smem + tb);
  unsigned char* qmem = smem + tb + ((sizeof(TileRing) + 127) & ~size_t(127));
  unsigned char* stage_mem = qmem + (QUEUE ? ((tile_queue_bytes<NT>() + 127) & ~size_t(127)) : 0);
  const size_t bb = (Stage::bytes(G.s_cap) + 127) & ~size_t(127);
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const int nbuf = G.nbuf;
  if( threadIdx.x == 0 )
  {
    for(int b = 0; b < nbuf; b++) { mbar_init(&R.full[b], 1); mbar_init(&R.empty[b], NWC); }
    mbar_fence_init();
  }
  op.load_tables(smem, NT);
  __syncthreads();                                // the only CTA-wide barrier of the kernel

  if( warp == NWC )
  {
    // ---- producer warp: meta + TMA bulk copies of tile n into buffer n % nbuf, as soon as that buffer is free
    unsigned n = 0;
    for(unsigned t = blockIdx.x; t < G.ntiles; t += gridDim.x, ++n)
    {
      const int b = int(n % unsigned(nbuf));
      if( n >= unsigned(nbuf) ) mbar_wait(&R.empty[b], ((n / unsigned(nbuf)) - 1u) & 1u);
      Stage B; B.bind(stage_mem + size_t(b) * bb, G.s_cap);
      int ti, j, k; tile_coords(G, t, ti, j, k);
      tile_produce<Op::HAS_W, Op::TYPES>(G, cell_start, F, R, b, B, ti, j, k);
    }
    return;
  }

  // ---- consumer warps
  const unsigned sub = lane % TPA, gsel = lane / TPA;
  const unsigned gmask = TPA == 32 ? 0xffffffffu : (((1u << TPA) - 1u) << (gsel * TPA));
  unsigned n = 0;
  for(unsigned t = blockIdx.x; t < G.ntiles; t += gridDim.x, ++n)
  {
    const int b = int(n % unsigned(nbuf));
    mbar_wait(&R.full[b], (n / unsigned(nbuf)) & 1u);
    const TileMeta& M = R.meta[b];
    const unsigned a_begin = M.a_begin, na = M.a_end - M.a_begin, c_off = M.c_off;
    Stage B; B.bind(stage_mem + size_t(b) * bb, G.s_cap);
    // Atoms are grabbed one ahead: the list window (offset, length) of the next atom is loaded while the current one is
    // processed, so a warp starts an atom with one dependent global load (its first list entry) instead of two or three
    // (ncu: 39 % of the force pass's stall samples were in the per-atom prologue / epilogue).
    auto grab = [&]() -> unsigned { unsigned k = 0; if( lane == 0 ) k = atomicAdd(&R.cursor[b], GPW); return __shfl_sync(0xffffffffu, k, 0); };
    auto window = [&](unsigned k, unsigned long long& e0w, unsigned& lenw)
    {
      e0w = 0ull; lenw = 0u;
      if( k + gsel < na )
      {
        const unsigned aw = a_begin + k + gsel;
        e0w = L.off[aw];
        lenw = (LMODE == LIST_SUB || REWRITE) ? L.sub_cnt[aw] : unsigned(L.off[aw + 1] - e0w);
      }
    };
    unsigned k0 = grab();
    unsigned long long e0 = 0ull; unsigned len = 0u;
    window(k0, e0, len);
    constexpr bool AHEAD = !op_no_ahead<Op>::value;
    while( k0 < na )
    {
      unsigned k0n = 0; unsigned long long e0n = 0ull; unsigned lenn = 0u;
      if constexpr ( AHEAD ) { k0n = grab(); window(k0n, e0n, lenn); }
      const bool valid = k0 + gsel < na;
      const unsigned a = a_begin + k0 + gsel;
      typename Op::Acc acc;
      op.init(acc);
      if( valid )
      {
        const unsigned sa = a + c_off;
        const double xa = B.x[sa], ya = B.y[sa], za = B.z[sa];
        op.start(acc, a, sa, B, smem);
        // per-atom list window [e0, e0 + len): 64-bit base pointers once, 32-bit indices inside the loops
        if constexpr ( REWRITE )
        {
          // sub-list of an earlier step (a superset of the in-range pairs while the inner skin holds): every entry is
          // re-evaluated on the current positions and its cached value rewritten (pair_d2 returns NaN for a dead pair)
          const unsigned short* __restrict__ sp = L.sub_idx + e0;
          double* __restrict__ pwo = L.pair_w + e0;
          unsigned e = sub;
          unsigned jn = e < len ? __ldcs(sp + e) : 0u;
          while( e < len )
          {
            const unsigned j = jn, ec = e;
            e += TPA;
            if( e < len ) jn = __ldcs(sp + e);
            double dx = B.x[j] - xa, dy = B.y[j] - ya, dz = B.z[j] - za;
            apply_xform<XFORM>(X, dx, dy, dz);
            pw_store(pwo, L.pw_plane, ec, op.pair_d2(acc, dx * dx + dy * dy + dz * dz, j, B, smem));
          }
        }
        else if constexpr ( LMODE == LIST_SUB )
        {
          // dense: every entry is in range (filtered by the pass that wrote the sub-list on these positions)
          const unsigned short* __restrict__ sp = L.sub_idx + e0;
          const double* __restrict__ pwp = L.pair_w + e0;
          unsigned e = sub;
          unsigned jn = e < len ? __ldcs(sp + e) : 0u;
          constexpr bool PW2 = PWI && op_pw_n<Op>::value == 2;
          double pvn = 0.0, pvn2 = 0.0;
          if( PWI && e < len ) { pvn = __ldcs(pwp + e); if( PW2 ) pvn2 = __ldcs(pwp + L.pw_plane + e); }
          while( e < len )
          {
            const unsigned j = jn;
            const double pv = pvn, pv2 = pvn2;
            e += TPA;
            if( e < len ) { jn = __ldcs(sp + e); if( PWI ) pvn = __ldcs(pwp + e); if( PW2 ) pvn2 = __ldcs(pwp + L.pw_plane + e); }   // next entry in flight while this pair is evaluated
            double dx = B.x[j] - xa, dy = B.y[j] - ya, dz = B.z[j] - za;
            apply_xform<XFORM>(X, dx, dy, dz);
            const double d2 = dx * dx + dy * dy + dz * dz;
            if constexpr ( PWI ) { if( pv == pv ) op.pair_pw(acc, dx, dy, dz, d2, j, B, smem, pv, pv2); }      // NaN = dead pair (inner skin)
            else                 { if( d2 <= op.rcut2 ) op.pair(acc, dx, dy, dz, d2, j, B, smem); }
          }
        }
        else if constexpr ( QUEUE )
        {
          // one warp per atom: filter 32 entries per step, ballot-compact the in-range ones into the warp's queue
          // (and into the global sub-list), evaluate the functor on full warps only
          double* qd = reinterpret_cast<double*>(qmem) + warp * TILE_QUEUE_SLOTS;
          unsigned short* qj = reinterpret_cast<unsigned short*>(qmem + size_t(NT / 32) * TILE_QUEUE_SLOTS * sizeof(double)) + warp * TILE_QUEUE_SLOTS;
          const unsigned short* __restrict__ lp = L.idx + e0;
          unsigned short* __restrict__ wp = L.sub_idx + e0;
          const unsigned lt = (1u << sub) - 1u;
          // branch-free filter step with 32-bit shared addressing: lanes past the end of the list test stage slot 0 and are
          // masked out of the ballot
          unsigned xad = smem_u32(B.x), yad = xad + G.s_cap * 8u, zad = xad + G.s_cap * 16u;
          unsigned qda = smem_u32(qd), qja = smem_u32(qj);
          double rc2 = LMODE == LIST_FULL_WRITE_SUB ? L.list_rc2 : op.rcut2;      // membership of the sub-list (>= the Op's own cut-off)
          // keep the loop invariants in registers (ptxas otherwise re-reads s_cap / rcut2 from the constant bank and rebuilds
          // the three field bases every step)
          asm volatile("" : "+r"(xad), "+r"(yad), "+r"(zad), "+r"(qda), "+r"(qja), "+d"(rc2));
          // FIFO over a 64-slot ring: qt entries queued, qh evaluated so far; the k-th queued entry IS the k-th entry of
          // the atom's in-range sub-list, so a batch [qh, qh+32) maps to 32 consecutive sub-list positions
          double* __restrict__ pw = L.pair_w + e0;
          unsigned qt = 0, qh = 0;
          // list indices are fetched two steps ahead (a step is too short to cover an HBM round trip: ncu showed ~20 % of
          // the stall samples on the first use of the index)
          constexpr bool QJ = Op::TYPES;        // the functor only looks at j for the neighbour's type
          const unsigned short* lpn = lp + sub;
          unsigned jn = sub < len ? __ldcs(lpn) : 0u;
          unsigned jnn = sub + 32 < len ? __ldcs(lpn + 32) : 0u;
          for(unsigned e = sub; e < len + sub; e += 32)          // e - sub < len : same trip count on every lane
          {
            const unsigned j = jn;
            jn = jnn;
            lpn += 32;
            jnn = e + 64 < len ? __ldcs(lpn + 32) : 0u;
            const unsigned j8 = 8u * j;
            double dx, dy, dz;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(dx) : "r"(xad + j8));
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(dy) : "r"(yad + j8));
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(dz) : "r"(zad + j8));
            dx -= xa; dy -= ya; dz -= za;
            apply_xform<XFORM>(X, dx, dy, dz);
            const double d2 = dx * dx + dy * dy + dz * dz;
            const bool in = (e < len) & (d2 <= rc2);
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if( in )
            {
              const unsigned k = qt + __popc(m & lt), slot = k & 63u;
              asm volatile("st.shared.f64 [%0], %1;" :: "r"(qda + 8u * slot), "d"(d2) : "memory");
              if( QJ ) asm volatile("st.shared.u16 [%0], %1;" :: "r"(qja + 2u * slot), "h"((unsigned short)j) : "memory");
              if( LMODE == LIST_FULL_WRITE_SUB ) wp[k] = (unsigned short)j;
            }
            qt += __popc(m);
            __syncwarp();
            if( qt - qh >= 32 )
            {
              const unsigned slot = (qh + sub) & 63u;
              const unsigned qjv = QJ ? unsigned(qj[slot]) : 0u;
              if constexpr ( PWO ) pw_store(pw, L.pw_plane, qh + sub, op.pair_d2(acc, qd[slot], qjv, B, smem));
              else                 op.pair_d2(acc, qd[slot], qjv, B, smem);
              qh += 32;
              __syncwarp();
            }
          }
          if( sub < qt - qh )
          {
            const unsigned slot = (qh + sub) & 63u;
            const unsigned qjv = QJ ? unsigned(qj[slot]) : 0u;
            if constexpr ( PWO ) pw_store(pw, L.pw_plane, qh + sub, op.pair_d2(acc, qd[slot], qjv, B, smem));
            else                 op.pair_d2(acc, qd[slot], qjv, B, smem);
          }
          __syncwarp();
          if( LMODE == LIST_FULL_WRITE_SUB && sub == 0 ) L.sub_cnt[a] = qt;
        }
        else
        {
          const unsigned short* __restrict__ lp = L.idx + e0;
          unsigned short* __restrict__ wp = L.sub_idx + e0;
          unsigned cnt = 0;
          unsigned jn = sub < len ? __ldcs(lp + sub) : 0u;
          for(unsigned e = 0; e < len; e += TPA)
          {
            const unsigned ee = e + sub;
            const unsigned j = jn;
            if( ee + TPA < len ) jn = __ldcs(lp + ee + TPA);
            bool in = false;
            if( ee < len )
            {
              double dx = B.x[j] - xa, dy = B.y[j] - ya, dz = B.z[j] - za;
              apply_xform<XFORM>(X, dx, dy, dz);
              const double d2 = dx * dx + dy * dy + dz * dz;
              in = d2 <= (LMODE == LIST_FULL_WRITE_SUB ? L.list_rc2 : op.rcut2);
              if( d2 <= op.rcut2 ) op.pair(acc, dx, dy, dz, d2, j, B, smem);
            }
            if( LMODE == LIST_FULL_WRITE_SUB )
            {
              const unsigned m = __ballot_sync(gmask, in) & gmask;
              if( in ) wp[cnt + __popc(m & ((1u << lane) - 1u))] = (unsigned short)j;
              cnt += __popc(m);
            }
          }
          if( LMODE == LIST_FULL_WRITE_SUB && sub == 0 ) L.sub_cnt[a] = cnt;
        }
      }
      op.template finish<TPA>(acc, a, valid, sub);
      if constexpr ( AHEAD ) { k0 = k0n; e0 = e0n; len = lenn; }
      else { k0 = grab(); window(k0, e0, len); }
    }
    __syncwarp();
    if( lane == 0 ) mbar_arrive(&R.empty[b]);       // this warp will not touch buffer b again until it is refilled
  }
}

// lmode: LIST_FULL / LIST_FULL_WRITE_SUB / LIST_SUB
template<int TPA, int NT, class Op>
static int launch_tile_pass(xsb_ctx* ctx, bool ghost, const Op& op, const double* w, int lmode = LIST_FULL,
                            double list_rc2 = 0.0, const int* mode = nullptr, int want_mode = 0)
{
  TileGeom G = make_tile_geom(ctx, ghost);
  if( G.ntiles == 0 ) return XSB_OK;
  const bool queue = Op::D2_ONLY && TPA == 32 && (lmode == LIST_FULL || lmode == LIST_FULL_WRITE_SUB);
  const size_t qb = queue ? tile_queue_bytes<NT>() : 0;
  G.nbuf = tile_smem_bytes<Op::HAS_W, Op::TYPES>(G.s_cap, op.table_bytes(), qb, 3) <= TILE_SMEM_MAX ? 3 : 2;
  const size_t smem = tile_smem_bytes<Op::HAS_W, Op::TYPES>(G.s_cap, op.table_bytes(), qb, G.nbuf);
  XSB_REQUIRE(ctx, smem <= TILE_SMEM_MAX, XSB_ERR_STATE, "tile pass: shared-memory plan exceeds 227 KiB (caller must pick a smaller table window)");
  const TileFields F{ ctx->f64[XSB_F_RX].p, ctx->f64[XSB_F_RY].p, ctx->f64[XSB_F_RZ].p, w, ctx->type.p };
  if( lmode != LIST_FULL )
  {
    XSB_CUDA(ctx, ctx->sub_idx.reserve(size_t(ctx->nbh_total) + 32, ctx->nbh_cfg.stream_prealloc_factor));
    XSB_CUDA(ctx, ctx->sub_cnt.reserve(size_t(ctx->n) + 1, XSB_GROW));
  }
  const size_t pw_plane = (size_t(ctx->nbh_total) + 63) & ~size_t(31);      // second cached value of a pair lives one plane further
  if( op_pw_out<Op>::value && queue && lmode == LIST_FULL_WRITE_SUB )
    XSB_CUDA(ctx, ctx->pair_w.reserve(pw_plane * size_t(op_pw_n<Op>::value) + 32, ctx->nbh_cfg.stream_prealloc_factor));
  const TileList L{ ctx->nbh_off.p, ctx->tl_idx.p, ctx->sub_idx.p, ctx->sub_cnt.p, ctx->pair_w.p, pw_plane,
                    list_rc2 > op.rcut2 ? list_rc2 : op.rcut2, mode, want_mode };
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
  if constexpr ( op_pw_in<Op>::value )      // such an Op only exists for the walk over a sub-list that carries the cache
  {
    XSB_REQUIRE(ctx, lmode == LIST_SUB && ctx->pair_w.p != nullptr, XSB_ERR_STATE, "tile pass: per-pair cache requested without a valid sub-list");
    return xf ? go(tile_pass_kernel<TPA, NT, true, LIST_SUB, Op>) : go(tile_pass_kernel<TPA, NT, false, LIST_SUB, Op>);
  }
  else
  {
    if constexpr ( Op::D2_ONLY && op_pw_out<Op>::value && TPA != 32 )
    {
      if( lmode == LIST_SUB_REWRITE )
      {
        XSB_REQUIRE(ctx, ctx->pair_w.p != nullptr, XSB_ERR_STATE, "tile pass: sub-list re-evaluation without a per-pair cache");
        return xf ? go(tile_pass_kernel<TPA, NT, true, LIST_SUB_REWRITE, Op>) : go(tile_pass_kernel<TPA, NT, false, LIST_SUB_REWRITE, Op>);
      }
    }
    XSB_REQUIRE(ctx, lmode != LIST_SUB_REWRITE, XSB_ERR_STATE, "tile pass: this operator cannot re-evaluate a sub-list");
    if( lmode == LIST_SUB )                 return xf ? go(tile_pass_kernel<TPA, NT, true, LIST_SUB, Op>) : go(tile_pass_kernel<TPA, NT, false, LIST_SUB, Op>);
    if( lmode == LIST_FULL_WRITE_SUB )      return xf ? go(tile_pass_kernel<TPA, NT, true, LIST_FULL_WRITE_SUB, Op>) : go(tile_pass_kernel<TPA, NT, false, LIST_FULL_WRITE_SUB, Op>);
    return xf ? go(tile_pass_kernel<TPA, NT, true, LIST_FULL, Op>) : go(tile_pass_kernel<TPA, NT, false, LIST_FULL, Op>);
  }
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
    for(int k = 0; k < 9; k++) red_add(p + k, v[k]); }
};

} // namespace xsb
