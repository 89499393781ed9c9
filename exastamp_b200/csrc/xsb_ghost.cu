// xsb_ghost.cu -- ghost operators: ghost_comm_scheme, ghost_update_r / ghost_update_opt /
// ghost_update_all_no_fv (owner -> ghost copies) and update_force_energy_from_ghost (ghost -> owner add).
// Reference names: src/mpi/update_ghosts.cu:30-47, src/mpi/update_from_ghosts.cu:29-48; graph
// data/config/config_move_particles.msp:82-96,134-135.  The reference implementation is exaNBody's MPI
// point-to-point code; here one process drives one GPU and the exchange is NCCL P2P (ncclSend/ncclRecv in one
// group per call) over NVLink, with pack / unpack kernels at both ends.  Periodic images are materialised as
// ghost particles exactly like the reference does, even on a single rank (self peer = one fused kernel).
//
// Decomposition: static 3-D bricks of the global cell grid, rank (px,py,pz) owns cells
// [p*G/P, (p+1)*G/P) per axis.  Because the decomposition is a pure function of xsb_domain_desc, every rank
// derives its send lists from its peers' receive lists locally: only particle counts travel at scheme time.
#include "xsb_ctx.h"
#include <algorithm>
#include <dlfcn.h>
#include <cub/device/device_radix_sort.cuh>

extern "C" int xsb_internal_nccl_header_version();                                       // xsb_ncclwin.cu
extern "C" int xsb_internal_win_peer_ptrs(void* win, int nranks, void** out_dev, cudaStream_t stream);

namespace xsb
{

// ---- NCCL, bound at run time so that the library loads without it and shares the copy torch already loaded ---
typedef struct { char internal[128]; } NcclUniqueId;
struct NcclApi
{
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm**, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  // symmetric-memory windows (NCCL >= 2.27; optional: the peer-memory ghost transport needs them)
  int (*GetVersion)(int*) = nullptr;
  int (*MemAlloc)(void**, size_t) = nullptr;
  int (*MemFree)(void*) = nullptr;
  int (*CommWindowRegister)(ncclComm*, void*, size_t, void**, int) = nullptr;
  int (*CommWindowDeregister)(ncclComm*, void*) = nullptr;
  bool ok = false;
};
static NcclApi g_nccl;
static const int NCCL_UINT8 = 1, NCCL_UINT64 = 5, NCCL_FLOAT64 = 8, NCCL_MAX = 2, NCCL_SUM = 0;   // ncclDataType_t / ncclRedOp_t values (nccl.h)

static bool nccl_load(std::string& why)
{
  if( g_nccl.ok ) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if( !h ) h = dlopen("libnccl.so.2", RTLD_NOW);
  if( !h ) h = dlopen("libnccl.so", RTLD_NOW);
  if( !h ) { why = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
  g_nccl.lib = h;
# define XSB_SYM(name) *(void**)(&g_nccl.name) = dlsym(h, "nccl" #name); if( !g_nccl.name ) { why = "missing symbol nccl" #name; return false; }
  XSB_SYM(GetUniqueId) XSB_SYM(CommInitRank) XSB_SYM(CommDestroy) XSB_SYM(GroupStart) XSB_SYM(GroupEnd) XSB_SYM(Send) XSB_SYM(Recv)
  XSB_SYM(AllReduce) XSB_SYM(AllGather) XSB_SYM(GetErrorString)
# undef XSB_SYM
  *(void**)(&g_nccl.GetVersion) = dlsym(h, "ncclGetVersion");
  *(void**)(&g_nccl.MemAlloc) = dlsym(h, "ncclMemAlloc");
  *(void**)(&g_nccl.MemFree) = dlsym(h, "ncclMemFree");
  *(void**)(&g_nccl.CommWindowRegister) = dlsym(h, "ncclCommWindowRegister");
  *(void**)(&g_nccl.CommWindowDeregister) = dlsym(h, "ncclCommWindowDeregister");
  g_nccl.ok = true;
  return true;
}

#define XSB_NCCL(ctx, call) do { int r__ = (call); if( r__ != 0 ) \
  return (ctx)->fail(XSB_ERR_NCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__)); } while(0)

// ---- scheme ------------------------------------------------------------------------------------------------
struct GhostCell { int ghost_cell; int owner_rank; int owner_cell; int w[3]; };

struct GhostState
{
  xsb_domain_desc dom{};
  int nranks = 1;
  // per peer (rank order) segment offsets into the concatenated particle lists, n = nranks+1
  std::vector<unsigned> send_off, recv_off;
  unsigned n_send = 0, n_recv = 0;
  DevBuf<unsigned> send_idx;            // owner particle (flat)   [n_send]
  DevBuf<unsigned char> send_code;      // shift code 0..26        [n_send]
  DevBuf<unsigned> recv_idx;            // ghost particle (flat)   [n_recv]
  DevBuf<unsigned long long> send_buf, recv_buf;   // 8-byte words, [nfields][segment] per peer
  double shift[27][3];
  // cell-level exchange plan: depends on the decomposition only, built once (ghost_list walks every cell of every rank's
  // grid: ~2-3 ms of host time per rebuild at 8 ranks when it was recomputed)
  std::vector<GhostCell> plan_mine; std::vector< std::vector<GhostCell> > plan_to_peer;
  xsb_domain_desc plan_dom{}; int plan_gl = -1;

  // ---- peer-memory transport (one node, NVLink / NVSwitch): the receive block of every rank is one NCCL symmetric window
  // (ncclMemAlloc + ncclCommWindowRegister: CUDA VMM, only these blocks are peer-mapped); the pack kernel of an exchange
  // stores straight into the peers' receive buffers, a release flag per (source, buffer) tells the receiver's unpack kernel
  // that a segment has landed.  No NCCL call per exchange, no staging copy, no host involvement.
  bool p2p_tried = false, p2p_ok = false;
  std::string p2p_why;                                  // why the NCCL transport is used instead
  unsigned long long* p2p_block = nullptr;              // my block: [2][64] flags, then two receive buffers of p2p_cap words
  void* p2p_win = nullptr;                              // ncclWindow_t of the block
  size_t p2p_cap = 0;
  std::vector<unsigned long long*> p2p_peer;            // peers' blocks mapped into this process (nullptr for myself)
  unsigned long long p2p_epoch = 0;                     // exchanges done on this transport (identical on all ranks)
  std::vector<unsigned> peer_send_off, peer_recv_off;   // [P][P+1] send_off / recv_off of every rank (all-gathered by the scheme)
  bool p2p_fit = false;                                 // every rank's buffers hold the exchanges of the current scheme (decided collectively)
};

constexpr int P2P_FLAG_WORDS = 128;                     // [2 buffers][64 sources]


static inline int block_start(int r, int G, int P) { return int((long long)r * G / P); }

static inline int owner_of(int gc, int G, int P)
{
  int r = int(((long long)(gc + 1) * P - 1) / G);   // largest r with start(r) <= gc
  while( r > 0 && block_start(r, G, P) > gc ) --r;
  while( r + 1 < P && block_start(r + 1, G, P) <= gc ) ++r;
  return r;
}

// receive list of `rank`: every ghost cell of its local grid that mirrors an existing domain cell
static void ghost_list(const xsb_domain_desc& d, int gl, const int rc[3], std::vector<GhostCell>& out)
{
  int s[3], n[3], dims[3];
  for(int a = 0; a < 3; a++) { s[a] = block_start(rc[a], d.global_cells[a], d.rank_dims[a]); n[a] = block_start(rc[a] + 1, d.global_cells[a], d.rank_dims[a]) - s[a]; dims[a] = n[a] + 2 * gl; }
  out.clear();
  for(int k = 0; k < dims[2]; k++) for(int j = 0; j < dims[1]; j++) for(int i = 0; i < dims[0]; i++)
  {
    const int l[3] = { i, j, k };
    if( i >= gl && i < dims[0] - gl && j >= gl && j < dims[1] - gl && k >= gl && k < dims[2] - gl ) continue;
    GhostCell g; g.ghost_cell = i + dims[0] * (j + dims[1] * k);
    int orank[3], ocell[3]; bool exists = true;
    for(int a = 0; a < 3; a++)
    {
      const int G = d.global_cells[a];
      int gc = s[a] + l[a] - gl, w = 0;
      if( gc < 0 || gc >= G )
      {
        if( !d.periodic[a] ) { exists = false; break; }
        w = gc >= 0 ? gc / G : -((-gc + G - 1) / G);
        gc -= w * G;
      }
      g.w[a] = w;
      orank[a] = owner_of(gc, G, d.rank_dims[a]);
      const int os = block_start(orank[a], G, d.rank_dims[a]);
      ocell[a] = gc - os + gl;
    }
    if( !exists ) continue;
    g.owner_rank = orank[0] + d.rank_dims[0] * (orank[1] + d.rank_dims[1] * orank[2]);
    int odims[2];
    for(int a = 0; a < 2; a++) odims[a] = block_start(orank[a] + 1, d.global_cells[a], d.rank_dims[a]) - block_start(orank[a], d.global_cells[a], d.rank_dims[a]) + 2 * gl;
    g.owner_cell = ocell[0] + odims[0] * (ocell[1] + odims[1] * ocell[2]);
    out.push_back(g);
  }
  std::stable_sort(out.begin(), out.end(), [](const GhostCell& a, const GhostCell& b) { return a.owner_rank < b.owner_rank; });
}

static inline int shift_code(const int w[3]) { auto c = [](int v) { return v < 0 ? 0 : (v > 0 ? 2 : 1); }; return c(w[0]) + 3 * c(w[1]) + 9 * c(w[2]); }

// ---- kernels -----------------------------------------------------------------------------------------------
// kind: 0 double, 1..3 position axis (adds shift), 4 u64, 5 u8, 16+k component k of the 9-double per-atom virial (AoS Mat3d)
struct FieldPtrs { const void* src[16]; void* dst[16]; int kind[16]; int nf; };
struct ShiftTab { double s[27][3]; };

__device__ __forceinline__ size_t word_index(int kind, unsigned i) { return kind >= 16 ? 9ull * i + unsigned(kind - 16) : size_t(i); }
__device__ __forceinline__ unsigned long long load_word(const void* base, int kind, unsigned i)
{
  if( kind == 5 ) return (unsigned long long)static_cast<const unsigned char*>(base)[i];
  return static_cast<const unsigned long long*>(base)[word_index(kind, i)];
}
__device__ __forceinline__ void store_word(void* base, int kind, unsigned i, unsigned long long v)
{
  if( kind == 5 ) static_cast<unsigned char*>(base)[i] = (unsigned char)v;
  else static_cast<unsigned long long*>(base)[word_index(kind, i)] = v;
}

// owner -> wire : buf[peer segment][field][k] ; shift added to positions on the sender side
__global__ void ghost_pack_kernel(unsigned n, FieldPtrs F, ShiftTab S, const unsigned* __restrict__ idx, const unsigned char* __restrict__ code,
                                  const unsigned* __restrict__ seg_of, const unsigned* __restrict__ seg_off, unsigned long long* __restrict__ buf, bool from_ghost, unsigned me)
{
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if( k >= n ) return;
  if( seg_of[k] == me ) return;   // self segment is served by ghost_self_kernel
  const unsigned q = seg_of[k], s0 = seg_off[q], len = seg_off[q + 1] - s0, src = idx[k];
  unsigned long long* seg = buf + size_t(F.nf) * s0;
  for(int f = 0; f < F.nf; f++)
  {
    unsigned long long w = load_word(F.src[f], F.kind[f], src);
    if( !from_ghost && F.kind[f] >= 1 && F.kind[f] <= 3 ) w = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)w) + S.s[code[k]][F.kind[f] - 1]);
    seg[size_t(f) * len + (k - s0)] = w;
  }
}

struct P2PFlags { unsigned long long* flag[64]; };      // one address per peer: its flag word, or the base of my segment in its buffer

// owner -> peer memory: like ghost_pack_kernel, but the segment base is an address inside the RECEIVER's buffer
__global__ void ghost_pack_p2p_kernel(unsigned n, FieldPtrs F, ShiftTab S, const unsigned* __restrict__ idx, const unsigned char* __restrict__ code,
                                      const unsigned* __restrict__ seg_of, const unsigned* __restrict__ seg_off, const P2PFlags dst_base,
                                      bool from_ghost, unsigned me)
{
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if( k >= n ) return;
  const unsigned q = seg_of[k];
  if( q == me ) return;
  const unsigned s0 = seg_off[q], len = seg_off[q + 1] - s0, src = idx[k];
  unsigned long long* seg = dst_base.flag[q];
  for(int f = 0; f < F.nf; f++)
  {
    unsigned long long w = load_word(F.src[f], F.kind[f], src);
    if( !from_ghost && F.kind[f] >= 1 && F.kind[f] <= 3 ) w = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)w) + S.s[code[k]][F.kind[f] - 1]);
    seg[size_t(f) * len + (k - s0)] = w;       // 8-byte store over NVLink, consecutive k -> consecutive addresses
  }
}

// after the pack kernel of this exchange: tell every peer that my segment of epoch `epoch` is complete in its buffer
__global__ void ghost_signal_kernel(P2PFlags T, int P, int me, unsigned long long epoch)
{
  const int q = threadIdx.x;
  if( q >= P || q == me || T.flag[q] == nullptr ) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(T.flag[q]), "l"(epoch) : "memory");
}

__device__ __forceinline__ void p2p_wait_all(const unsigned long long* __restrict__ flags, const unsigned* __restrict__ in_off, int P, int me, unsigned long long epoch)
{
  if( threadIdx.x < unsigned(P) && int(threadIdx.x) != me && in_off[threadIdx.x + 1] > in_off[threadIdx.x] )
  {
    unsigned long long v;
    do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory"); } while( v < epoch );
  }
  __syncthreads();
}

// wire -> ghost (copy) or wire -> owner (add)
__global__ void ghost_unpack_kernel(unsigned n, FieldPtrs F, const unsigned* __restrict__ idx, const unsigned* __restrict__ seg_of,
                                    const unsigned* __restrict__ seg_off, const unsigned long long* __restrict__ buf, bool add, unsigned me,
                                    const unsigned long long* __restrict__ flags, int P, unsigned long long epoch)
{
  if( flags ) p2p_wait_all(flags, seg_off, P, int(me), epoch);      // peer-memory transport: every block waits for all incoming segments
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if( k >= n ) return;
  if( seg_of[k] == me ) return;
  const unsigned q = seg_of[k], s0 = seg_off[q], len = seg_off[q + 1] - s0, dst = idx[k];
  const unsigned long long* seg = buf + size_t(F.nf) * s0;
  for(int f = 0; f < F.nf; f++)
  {
    const unsigned long long w = seg[size_t(f) * len + (k - s0)];
    if( add ) atomicAdd(static_cast<double*>(F.dst[f]) + word_index(F.kind[f], dst), __longlong_as_double((long long)w));
    else store_word(F.dst[f], F.kind[f], dst, w);
  }
}

// single-rank fast path: ghost[k] = owner[k] (+ shift), no staging buffer
__global__ void ghost_self_kernel(unsigned n, FieldPtrs F, ShiftTab S, const unsigned* __restrict__ send_idx, const unsigned char* __restrict__ code,
                                  const unsigned* __restrict__ recv_idx, bool reverse_add)
{
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if( k >= n ) return;
  const unsigned o = send_idx[k], g = recv_idx[k];
  for(int f = 0; f < F.nf; f++)
  {
    if( reverse_add ) { atomicAdd(static_cast<double*>(F.dst[f]) + word_index(F.kind[f], o), static_cast<const double*>(F.src[f])[word_index(F.kind[f], g)]); continue; }
    unsigned long long w = load_word(F.src[f], F.kind[f], o);
    if( F.kind[f] >= 1 && F.kind[f] <= 3 ) w = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)w) + S.s[code[k]][F.kind[f] - 1]);
    store_word(F.dst[f], F.kind[f], g, w);
  }
}

// ghost scheme lists: one warp per (ghost cell, owner cell) entry expands its particle range
struct RangeEntry { unsigned start, count, out, seg_code; };   // seg_code = peer | code << 16
__global__ void __launch_bounds__(256) expand_ranges_kernel(unsigned n_entries, const RangeEntry* __restrict__ E, unsigned* __restrict__ idx,
                                                            unsigned char* __restrict__ code, unsigned* __restrict__ seg)
{
  const unsigned i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
  if( i >= n_entries ) return;
  const RangeEntry e = E[i];
  for(unsigned p = lane; p < e.count; p += 32u)
  {
    idx[e.out + p] = e.start + p;
    if( code ) code[e.out + p] = (unsigned char)(e.seg_code >> 16);
    seg[e.out + p] = e.seg_code & 0xffffu;
  }
}

static int make_fields(xsb_ctx* ctx, uint32_t mask, bool reverse, FieldPtrs& F)
{
  F.nf = 0;
  for(int f = 0; f < XSB_F_COUNT_; f++)
  {
    if( !(mask & (1u << f)) ) continue;
    if( f == XSB_F_VIRIAL )
    {
      // update_virial_force_energy_from_ghost (src/mpi/update_from_ghosts.cu:43): the 9 components travel as 9 words per atom
      XSB_REQUIRE(ctx, F.nf + 9 <= 16, XSB_ERR_INVALID, "at most 16 words per atom per ghost exchange");
      int rc = xsb_internal_ensure_virial(ctx); if( rc ) return rc;
      for(int k = 0; k < 9; k++) { F.src[F.nf] = ctx->f64[f].p; F.dst[F.nf] = ctx->f64[f].p; F.kind[F.nf] = 16 + k; ++F.nf; }
      continue;
    }
    XSB_REQUIRE(ctx, F.nf < 16, XSB_ERR_INVALID, "at most 16 words per atom per ghost exchange");
    void* p; int kind = 0;
    if( f == XSB_F_TYPE ) { p = ctx->type.p; kind = 5; }
    else if( f == XSB_F_ID ) { p = ctx->id.p; kind = 4; }
    else { p = ctx->f64[f].p; kind = (f == XSB_F_RX) ? 1 : (f == XSB_F_RY) ? 2 : (f == XSB_F_RZ) ? 3 : 0; }
    XSB_REQUIRE(ctx, !reverse || kind <= 3, XSB_ERR_INVALID, "only real-valued fields can be reduced from ghosts");
    F.src[F.nf] = p; F.dst[F.nf] = p; F.kind[F.nf] = kind; ++F.nf;
  }
  XSB_REQUIRE(ctx, F.nf > 0, XSB_ERR_INVALID, "empty field mask");
  return XSB_OK;
}

// exchange(): forward (owner->ghost copy) or reverse (ghost->owner add)
static int exchange(xsb_ctx* ctx, uint32_t mask, bool reverse)
{
  GhostState* G = ctx->ghost;
  XSB_REQUIRE(ctx, G != nullptr && ctx->ghost_valid, XSB_ERR_STATE, "xsb_ghost_comm_scheme must be called first (and again after every particle re-layout)");
  if( mask & ((1u << XSB_F_RX) | (1u << XSB_F_RY) | (1u << XSB_F_RZ)) ) ctx->pos_epoch++;
  FieldPtrs F; int rc = make_fields(ctx, mask, reverse, F); if( rc ) return rc;
  ShiftTab S; std::memcpy(S.s, G->shift, sizeof(S.s));
  const int me = ctx->rank, P = G->nranks;
  // self segment
  const unsigned s0 = G->send_off[me], ns = G->send_off[me + 1] - s0, r0 = G->recv_off[me];
  if( ns )
  {
    ghost_self_kernel<<<(ns + 255) / 256, 256, 0, ctx->stream>>>(ns, F, S, G->send_idx.p + s0, G->send_code.p + s0, G->recv_idx.p + r0, reverse);
    XSB_LAUNCH_CHECK(ctx);
  }
  if( P == 1 ) return XSB_OK;
  XSB_REQUIRE(ctx, ctx->comm != nullptr, XSB_ERR_STATE, "multi-rank ghost exchange needs xsb_comm_init");
  // remote peers: pack everything (self segment included, it is simply not sent), one grouped send/recv, unpack
  const unsigned n_out = reverse ? G->n_recv : G->n_send, n_in = reverse ? G->n_send : G->n_recv;
  const std::vector<unsigned>& out_off = reverse ? G->recv_off : G->send_off;
  const std::vector<unsigned>& in_off = reverse ? G->send_off : G->recv_off;
  const size_t words = size_t(F.nf) * std::max(G->n_send, G->n_recv) + 16;
  (void)words;
  if( G->p2p_ok && G->p2p_fit )
  {
    // ---- peer-memory transport
    const unsigned long long epoch = ++G->p2p_epoch;
    const int b = int(epoch & 1ull);
    const std::vector<unsigned>& peer_in = reverse ? G->peer_send_off : G->peer_recv_off;      // where my segment starts in peer q's list
    P2PFlags D{}, T{};
    for(int q = 0; q < P; q++)
    {
      D.flag[q] = nullptr; T.flag[q] = nullptr;
      if( q == me || out_off[q + 1] == out_off[q] ) continue;
      unsigned long long* blk = G->p2p_peer[q];
      D.flag[q] = blk + P2P_FLAG_WORDS + size_t(b) * G->p2p_cap + size_t(F.nf) * peer_in[size_t(q) * (P + 1) + me];
      T.flag[q] = blk + size_t(b) * 64 + me;
    }
    const unsigned n_out = reverse ? G->n_recv : G->n_send, n_in = reverse ? G->n_send : G->n_recv;
    if( n_out )
    {
      ghost_pack_p2p_kernel<<<(n_out + 255) / 256, 256, 0, ctx->stream>>>(n_out, F, S, reverse ? G->recv_idx.p : G->send_idx.p, G->send_code.p,
          reverse ? ctx->gseg_recv.p : ctx->gseg_send.p, reverse ? ctx->goff_recv.p : ctx->goff_send.p, D, reverse, unsigned(me));
      XSB_LAUNCH_CHECK(ctx);
    }
    ghost_signal_kernel<<<1, 64, 0, ctx->stream>>>(T, P, me, epoch);
    XSB_LAUNCH_CHECK(ctx);
    if( n_in )
    {
      ghost_unpack_kernel<<<(n_in + 255) / 256, 256, 0, ctx->stream>>>(n_in, F, reverse ? G->send_idx.p : G->recv_idx.p, reverse ? ctx->gseg_send.p : ctx->gseg_recv.p,
          reverse ? ctx->goff_send.p : ctx->goff_recv.p, G->p2p_block + P2P_FLAG_WORDS + size_t(b) * G->p2p_cap, reverse, unsigned(me),
          G->p2p_block + size_t(b) * 64, P, epoch);
      XSB_LAUNCH_CHECK(ctx);
    }
    return XSB_OK;
  }
  XSB_CUDA(ctx, G->send_buf.reserve(size_t(F.nf) * std::max(G->n_send, G->n_recv) + 16, XSB_GROW_GHOST));
  XSB_CUDA(ctx, G->recv_buf.reserve(size_t(F.nf) * std::max(G->n_send, G->n_recv) + 16, XSB_GROW_GHOST));
  const unsigned* d_out_seg = reverse ? ctx->gseg_recv.p : ctx->gseg_send.p;   // seg_of arrays (built with the scheme)
  const unsigned* d_in_seg = reverse ? ctx->gseg_send.p : ctx->gseg_recv.p;
  const unsigned* d_out_off = reverse ? ctx->goff_recv.p : ctx->goff_send.p;
  const unsigned* d_in_off = reverse ? ctx->goff_send.p : ctx->goff_recv.p;
  if( n_out )
  {
    ghost_pack_kernel<<<(n_out + 255) / 256, 256, 0, ctx->stream>>>(n_out, F, S, reverse ? G->recv_idx.p : G->send_idx.p, G->send_code.p,
                                                                     d_out_seg, d_out_off, G->send_buf.p, reverse, unsigned(me));
    XSB_LAUNCH_CHECK(ctx);
  }
  XSB_NCCL(ctx, g_nccl.GroupStart());
  for(int q = 0; q < P; q++)
  {
    if( q == me ) continue;
    const size_t so = size_t(F.nf) * out_off[q], sc = size_t(F.nf) * (out_off[q + 1] - out_off[q]);
    const size_t ro = size_t(F.nf) * in_off[q], rcnt = size_t(F.nf) * (in_off[q + 1] - in_off[q]);
    if( sc ) XSB_NCCL(ctx, g_nccl.Send(G->send_buf.p + so, sc, NCCL_UINT64, q, ctx->comm, ctx->stream));
    if( rcnt ) XSB_NCCL(ctx, g_nccl.Recv(G->recv_buf.p + ro, rcnt, NCCL_UINT64, q, ctx->comm, ctx->stream));
  }
  XSB_NCCL(ctx, g_nccl.GroupEnd());
  if( n_in )
  {
    ghost_unpack_kernel<<<(n_in + 255) / 256, 256, 0, ctx->stream>>>(n_in, F, reverse ? G->send_idx.p : G->recv_idx.p, d_in_seg, d_in_off, G->recv_buf.p, reverse, unsigned(me), nullptr, P, 0ull);
    XSB_LAUNCH_CHECK(ctx);
  }
  return XSB_OK;
}

// Peer-memory transport set-up.  First call: allocate my receive block, publish its CUDA IPC handle, map every peer's block.
// Every call (= every ghost_comm_scheme): all-gather the segment offsets and the buffer capacity of every rank, so that a
// sender knows where its segment starts in each receiver's buffer and ALL ranks take the same p2p-or-NCCL decision.
static int p2p_prepare(xsb_ctx* ctx, GhostState* G)
{
  const int P = G->nranks, me = ctx->rank;
  if( !G->p2p_tried )
  {
    G->p2p_tried = true;
    // round 1: can everybody try, and how large must the (symmetric: same size everywhere) block be
    int ok = 1;
    if( getenv("XSB_GHOST_NCCL") ) { ok = 0; G->p2p_why = "XSB_GHOST_NCCL is set"; }
    if( ok && P > 64 ) { ok = 0; G->p2p_why = "more than 64 ranks"; }
    int rtv = 0;
    if( ok && (!g_nccl.GetVersion || !g_nccl.MemAlloc || !g_nccl.CommWindowRegister || g_nccl.GetVersion(&rtv) != 0) ) { ok = 0; G->p2p_why = "this NCCL has no symmetric-memory windows"; }
    if( ok && (xsb_internal_nccl_header_version() == 0 || rtv != xsb_internal_nccl_header_version()) )
    { ok = 0; G->p2p_why = "NCCL device-API headers of the build (" + std::to_string(xsb_internal_nccl_header_version()) + ") do not match the loaded library (" + std::to_string(rtv) + ")"; }
    const size_t rec = 16;
    XSB_CUDA(ctx, ctx->scratch.reserve(rec * size_t(P + 1) + 64));
    unsigned long long h_rec[2] = { (unsigned long long)ok, (unsigned long long)std::max(G->n_send, G->n_recv) };
    std::vector<unsigned long long> all(2 * size_t(P));
    auto gather = [&]() -> int
    {
      XSB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch.p, h_rec, rec, cudaMemcpyHostToDevice, ctx->stream));
      XSB_NCCL(ctx, g_nccl.AllGather(ctx->scratch.p, ctx->scratch.p + rec, rec, NCCL_UINT8, ctx->comm, ctx->stream));
      XSB_CUDA(ctx, cudaMemcpyAsync(all.data(), ctx->scratch.p + rec, rec * size_t(P), cudaMemcpyDeviceToHost, ctx->stream));
      XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      return XSB_OK;
    };
    int rcg = gather(); if( rcg ) return rcg;
    unsigned long long need = 0;
    for(int q = 0; q < P; q++) { if( !all[2 * size_t(q)] ) { if( ok ) G->p2p_why = "a peer cannot use symmetric-memory windows"; ok = 0; } need = std::max(need, all[2 * size_t(q) + 1]); }
    if( ok )
    {
      // round 2 (collective): allocate + register.  16 words per atom is the widest exchange; 2x head-room for later schemes
      G->p2p_cap = std::max<size_t>(size_t(32) * size_t(need), size_t(1) << 20);
      const size_t bytes = (size_t(P2P_FLAG_WORDS) + 2 * G->p2p_cap) * sizeof(unsigned long long);
      void* blk = nullptr; int good = 1;
      if( g_nccl.MemAlloc(&blk, bytes) != 0 || !blk ) { good = 0; G->p2p_why = "ncclMemAlloc failed"; }
      G->p2p_block = static_cast<unsigned long long*>(blk);
      if( good ) cudaMemsetAsync(G->p2p_block, 0, P2P_FLAG_WORDS * sizeof(unsigned long long), ctx->stream);
      h_rec[0] = (unsigned long long)good;
      if( (rcg = gather()) ) return rcg;
      for(int q = 0; q < P; q++) if( !all[2 * size_t(q)] ) { if( good ) G->p2p_why = "a peer could not allocate its receive block"; good = 0; }
      if( good )
      {
        if( g_nccl.CommWindowRegister(ctx->comm, G->p2p_block, bytes, &G->p2p_win, /*NCCL_WIN_COLL_SYMMETRIC*/ 0x01) != 0 || !G->p2p_win ) { good = 0; G->p2p_why = "ncclCommWindowRegister(NCCL_WIN_COLL_SYMMETRIC) failed"; G->p2p_win = nullptr; }
        G->p2p_peer.assign(size_t(P), nullptr);
        if( good )
        {
          void** d_ptrs = reinterpret_cast<void**>(ctx->scratch.p);
          if( xsb_internal_win_peer_ptrs(G->p2p_win, P, d_ptrs, ctx->stream) != 0 ) { good = 0; G->p2p_why = "window address query failed"; }
          std::vector<void*> hp(size_t(P), nullptr);
          XSB_CUDA(ctx, cudaMemcpyAsync(hp.data(), d_ptrs, sizeof(void*) * size_t(P), cudaMemcpyDeviceToHost, ctx->stream));
          XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
          for(int q = 0; q < P; q++) { G->p2p_peer[size_t(q)] = static_cast<unsigned long long*>(hp[size_t(q)]); if( !hp[size_t(q)] ) { good = 0; G->p2p_why = "window address of a peer is null"; } }
        }
        // round 3: everybody registered and got addresses
        h_rec[0] = (unsigned long long)good;
        if( (rcg = gather()) ) return rcg;
        for(int q = 0; q < P; q++) if( !all[2 * size_t(q)] ) { if( good ) G->p2p_why = "a peer could not register its window"; good = 0; }
      }
      ok = good;
    }
    G->p2p_ok = ok != 0;
    if( !G->p2p_ok )
    {
      G->p2p_peer.clear();
      if( G->p2p_win && g_nccl.CommWindowDeregister ) { g_nccl.CommWindowDeregister(ctx->comm, G->p2p_win); G->p2p_win = nullptr; }
      if( G->p2p_block && g_nccl.MemFree ) { g_nccl.MemFree(G->p2p_block); G->p2p_block = nullptr; }
    }
  }
  G->p2p_fit = false;
  if( !G->p2p_ok ) return XSB_OK;
  // per scheme: [P][2(P+1)+2] u32 = send_off, recv_off, capacity (lo, hi) of every rank
  const size_t rec = size_t(2 * (P + 1) + 2);
  std::vector<unsigned> mine(rec), all(rec * size_t(P));
  for(int q = 0; q <= P; q++) { mine[size_t(q)] = G->send_off[size_t(q)]; mine[size_t(P + 1 + q)] = G->recv_off[size_t(q)]; }
  mine[rec - 2] = unsigned(G->p2p_cap & 0xffffffffull); mine[rec - 1] = unsigned(G->p2p_cap >> 32);
  XSB_CUDA(ctx, ctx->tmp32c.reserve(rec * size_t(P + 1) + 16));
  XSB_CUDA(ctx, cudaMemcpyAsync(ctx->tmp32c.p, mine.data(), rec * 4, cudaMemcpyHostToDevice, ctx->stream));
  XSB_NCCL(ctx, g_nccl.AllGather(ctx->tmp32c.p, ctx->tmp32c.p + rec, rec * 4, NCCL_UINT8, ctx->comm, ctx->stream));
  XSB_CUDA(ctx, cudaMemcpyAsync(all.data(), ctx->tmp32c.p + rec, all.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  G->peer_send_off.assign(size_t(P) * (P + 1), 0u); G->peer_recv_off.assign(size_t(P) * (P + 1), 0u);
  bool fit = true;
  for(int q = 0; q < P; q++)
  {
    const unsigned* r = all.data() + rec * size_t(q);
    for(int k = 0; k <= P; k++) { G->peer_send_off[size_t(q) * (P + 1) + k] = r[k]; G->peer_recv_off[size_t(q) * (P + 1) + k] = r[P + 1 + k]; }
    const size_t cap = size_t(r[rec - 2]) | (size_t(r[rec - 1]) << 32);
    if( size_t(16) * std::max(r[P], r[2 * P + 1]) + 16 > cap ) fit = false;        // 16 words per atom is the widest exchange
  }
  // my segment for peer q must be as long as what q expects from me
  for(int q = 0; q < P; q++)
  {
    if( q == me ) continue;
    const unsigned sl = G->send_off[size_t(q) + 1] - G->send_off[size_t(q)];
    const unsigned ql = G->peer_recv_off[size_t(q) * (P + 1) + me + 1] - G->peer_recv_off[size_t(q) * (P + 1) + me];
    XSB_REQUIRE(ctx, sl == ql, XSB_ERR_STATE, "ghost scheme: send / receive segment lengths of two ranks disagree");
  }
  G->p2p_fit = fit;
  return XSB_OK;
}

} // namespace xsb

using namespace xsb;

void xsb_ghost_release(xsb_ctx* ctx)
{
  if( ctx->ghost )
  {
    ctx->ghost->send_idx.release(); ctx->ghost->send_code.release(); ctx->ghost->recv_idx.release();
    ctx->ghost->send_buf.release(); ctx->ghost->recv_buf.release();
    if( ctx->ghost->p2p_win && g_nccl.CommWindowDeregister && ctx->comm ) g_nccl.CommWindowDeregister(ctx->comm, ctx->ghost->p2p_win);
    if( ctx->ghost->p2p_block && g_nccl.MemFree ) g_nccl.MemFree(ctx->ghost->p2p_block);
    delete ctx->ghost; ctx->ghost = nullptr;
  }
  ctx->old_cell_start.release(); ctx->tmp64.release(); ctx->tmp32a.release(); ctx->tmp32b.release(); ctx->tmp32c.release(); ctx->tmp32d.release(); ctx->gseg_send.release(); ctx->gseg_recv.release(); ctx->goff_send.release(); ctx->goff_recv.release(); ctx->backup.release();
  ctx->move_stage_b.release(); ctx->move_stage_c.release(); ctx->move_stage8_b.release(); ctx->move_stage8_c.release();
  if( ctx->comm && g_nccl.ok ) { g_nccl.CommDestroy(ctx->comm); ctx->comm = nullptr; }
}

extern "C" {

int xsb_comm_unique_id(void* id128)
{
  std::string why;
  if( !id128 || !nccl_load(why) ) return XSB_ERR_NCCL;
  NcclUniqueId id;
  if( g_nccl.GetUniqueId(&id) != 0 ) return XSB_ERR_NCCL;
  std::memcpy(id128, &id, 128);
  return XSB_OK;
}

int xsb_comm_init(xsb_ctx* ctx, int nranks, int rank, const void* id128)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, nranks >= 1 && rank >= 0 && rank < nranks, XSB_ERR_INVALID, "bad rank / nranks");
  ctx->nranks = nranks; ctx->rank = rank;
  if( nranks == 1 ) return XSB_OK;
  XSB_REQUIRE(ctx, id128 != nullptr, XSB_ERR_INVALID, "null ncclUniqueId");
  std::string why;
  if( !nccl_load(why) ) return ctx->fail(XSB_ERR_NCCL, "%s", why.c_str());
  XSB_CUDA(ctx, cudaSetDevice(ctx->device));
  NcclUniqueId id; std::memcpy(&id, id128, 128);
  XSB_NCCL(ctx, g_nccl.CommInitRank(&ctx->comm, nranks, id, rank));
  return XSB_OK;
}

int xsb_comm_allreduce_max(xsb_ctx* ctx, double* inout_host)
{
  XSB_ENTER(ctx);
  if( ctx->nranks == 1 ) return XSB_OK;
  XSB_REQUIRE(ctx, ctx->comm != nullptr, XSB_ERR_STATE, "xsb_comm_init must be called first");
  XSB_CUDA(ctx, ctx->scratch64.reserve(16));
  XSB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch64.p, inout_host, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  XSB_NCCL(ctx, g_nccl.AllReduce(ctx->scratch64.p, ctx->scratch64.p, 1, NCCL_FLOAT64, NCCL_MAX, ctx->comm, ctx->stream));
  XSB_CUDA(ctx, cudaMemcpyAsync(inout_host, ctx->scratch64.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return XSB_OK;
}

} // extern "C" (reopened below)

// ---- migrate_cell_particles across ranks (config_move_particles.msp:89-96; impl ext exaNBody, MPI) --------------------
namespace xsb
{
struct MigrateParams
{
  double g0[3], box[3], cell;     // grid-space corner of the global domain, its extent, cell edge
  int periodic[3], gcells[3], rdims[3];
};

// wraps positions into the global box, finds the brick that owns the particle's cell
__global__ void migrate_dest_kernel(unsigned n, MigrateParams M, double* __restrict__ rx, double* __restrict__ ry, double* __restrict__ rz,
                                    unsigned* __restrict__ key, unsigned* __restrict__ val, unsigned long long* __restrict__ counts)
{
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i >= n ) return;
  double r[3] = { rx[i], ry[i], rz[i] };
  int rk[3];
# pragma unroll
  for(int a = 0; a < 3; a++)
  {
    if( M.periodic[a] ) { r[a] -= floor((r[a] - M.g0[a]) / M.box[a]) * M.box[a]; if( r[a] - M.g0[a] >= M.box[a] ) r[a] = M.g0[a]; }
    int c = int(floor((r[a] - M.g0[a]) / M.cell));
    c = min(max(c, 0), M.gcells[a] - 1);
    int q = int(((long long)(c + 1) * M.rdims[a] - 1) / M.gcells[a]);
    while( q > 0 && int((long long)q * M.gcells[a] / M.rdims[a]) > c ) --q;
    while( q + 1 < M.rdims[a] && int((long long)(q + 1) * M.gcells[a] / M.rdims[a]) <= c ) ++q;
    rk[a] = q;
  }
  rx[i] = r[0]; ry[i] = r[1]; rz[i] = r[2];
  const unsigned dest = unsigned(rk[0] + M.rdims[0] * (rk[1] + M.rdims[1] * rk[2]));
  key[i] = dest; val[i] = i;
  atomicAdd(&counts[dest], 1ull);
}

// particles that stay on this rank: sorted positions [lo, lo + stay) -> their slots [r0, r0 + stay) of the new arrays
struct MigFields { const unsigned long long* src[7]; unsigned long long* dst[7]; };
__global__ void migrate_stay_kernel(unsigned stay, unsigned lo, unsigned r0, const unsigned* __restrict__ perm, MigFields F,
                                    const unsigned char* __restrict__ tsrc, unsigned char* __restrict__ tdst)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if( t >= stay ) return;
  const unsigned s = perm[lo + t];
# pragma unroll
  for(int k = 0; k < 7; k++) F.dst[k][r0 + t] = F.src[k][s];
  tdst[r0 + t] = tsrc[s];
}
// particles that leave: one 64-byte record each (7 fields + type), in destination order, so a peer gets ONE message
__global__ void migrate_pack_kernel(unsigned nl, unsigned lo, unsigned stay, const unsigned* __restrict__ perm, MigFields F,
                                    const unsigned char* __restrict__ tsrc, unsigned long long* __restrict__ rec)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if( t >= nl ) return;
  const unsigned s = perm[t < lo ? t : t + stay];
# pragma unroll
  for(int k = 0; k < 7; k++) rec[8ull * t + k] = F.src[k][s];
  rec[8ull * t + 7] = tsrc[s];
}
__global__ void migrate_unpack_kernel(unsigned nin, unsigned r0, unsigned stay, const unsigned long long* __restrict__ rec, MigFields F, unsigned char* __restrict__ tdst)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if( t >= nin ) return;
  const unsigned d = t < r0 ? t : t + stay;
# pragma unroll
  for(int k = 0; k < 7; k++) F.dst[k][d] = rec[8ull * t + k];
  tdst[d] = (unsigned char)rec[8ull * t + 7];
}
} // namespace xsb

int xsb_internal_migrate(xsb_ctx* ctx, const xsb_domain_desc* dom, unsigned n, double* const d[7], unsigned char* types,
                         unsigned* n_new, double* e[7], unsigned char** types_new)
{
  const int P = ctx->nranks, me = ctx->rank;
  XSB_REQUIRE(ctx, ctx->comm != nullptr, XSB_ERR_STATE, "multi-rank move_particles needs xsb_comm_init");
  XSB_REQUIRE(ctx, P <= 64, XSB_ERR_UNSUPPORTED, "more than 64 ranks");
  const xsb_grid_desc& g = ctx->grid;
  MigrateParams M;
  for(int a = 0; a < 3; a++)
  {
    M.g0[a] = g.origin[a] + g.ghost_layers * g.cell_size - block_start(dom->rank_coord[a], dom->global_cells[a], dom->rank_dims[a]) * g.cell_size;
    M.box[a] = dom->global_cells[a] * g.cell_size; M.periodic[a] = dom->periodic[a]; M.gcells[a] = dom->global_cells[a]; M.rdims[a] = dom->rank_dims[a];
  }
  M.cell = g.cell_size;
  XSB_CUDA(ctx, ctx->tmp32a.reserve(n + 16, XSB_GROW)); XSB_CUDA(ctx, ctx->tmp32b.reserve(n + 16, XSB_GROW));
  XSB_CUDA(ctx, ctx->tmp32c.reserve(n + 16, XSB_GROW)); XSB_CUDA(ctx, ctx->tmp32d.reserve(n + 16, XSB_GROW));
  unsigned *key = ctx->tmp32a.p, *val = ctx->tmp32b.p, *key2 = ctx->tmp32c.p, *perm = ctx->tmp32d.p;
  XSB_CUDA(ctx, ctx->scratch64.reserve(size_t(P) * (P + 1) + 16));
  unsigned long long* counts = ctx->scratch64.p;            // [P] mine, then [P][P] gathered
  unsigned long long* mat = counts + P;
  XSB_CUDA(ctx, cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * size_t(P) * (P + 1), ctx->stream));
  const unsigned grid = (n + 255) / 256;
  if( n )
  {
    migrate_dest_kernel<<<grid, 256, 0, ctx->stream>>>(n, M, d[0], d[1], d[2], key, val, counts);
    XSB_LAUNCH_CHECK(ctx);
    int end_bit = 1; while( (1 << end_bit) < P ) ++end_bit;
    size_t tmp = 0;
    XSB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp, key, key2, val, perm, int(n), 0, end_bit, ctx->stream));
    XSB_CUDA(ctx, ctx->scratch.reserve(tmp + 16));
    XSB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->scratch.p, tmp, key, key2, val, perm, int(n), 0, end_bit, ctx->stream));   // stable
    ctx->launches += 3;
  }
  XSB_NCCL(ctx, g_nccl.AllGather(counts, mat, size_t(P), NCCL_UINT64, ctx->comm, ctx->stream));
  std::vector<unsigned long long> hm(size_t(P) * P);
  XSB_CUDA(ctx, cudaMemcpyAsync(hm.data(), mat, hm.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::vector<size_t> soff(P + 1, 0), roff(P + 1, 0);
  for(int q = 0; q < P; q++) { soff[q + 1] = soff[q] + size_t(hm[size_t(me) * P + q]); roff[q + 1] = roff[q] + size_t(hm[size_t(q) * P + me]); }
  XSB_REQUIRE(ctx, soff[P] == n, XSB_ERR_STATE, "migrate: destination counts do not add up");
  const size_t nn = roff[P];
  XSB_REQUIRE(ctx, nn < 0xFFFFFFF0ull, XSB_ERR_OVERFLOW, "more than 2^32 particles per GPU after migration");
  // receive arrays: segment q = particles arriving from rank q, self included.  Stayers are gathered straight into
  // their segment; leavers travel as 64-byte records, one grouped send/recv per peer (not one per field).
  const size_t stay = soff[me + 1] - soff[me], nl = n - stay, nin = nn - stay;
  XSB_CUDA(ctx, ctx->move_stage_b.reserve(8 * std::max<size_t>(nl + nin, size_t(n) / 16 + 4096) + 16, 2.0));      // record buffers: sized once, well above the usual traffic (a re-allocation synchronises the device)
  XSB_CUDA(ctx, ctx->move_stage_c.reserve(7 * (nn + 1), XSB_GROW_GHOST)); XSB_CUDA(ctx, ctx->move_stage8_c.reserve(nn + 16, XSB_GROW_GHOST));
  for(int k = 0; k < 7; k++) e[k] = ctx->move_stage_c.p + size_t(k) * (nn + 1);
  unsigned char* et8 = ctx->move_stage8_c.p;
  unsigned long long* srec = reinterpret_cast<unsigned long long*>(ctx->move_stage_b.p);
  unsigned long long* rrec = srec + 8 * nl;
  MigFields F;
  for(int k = 0; k < 7; k++) { F.src[k] = reinterpret_cast<const unsigned long long*>(d[k]); F.dst[k] = reinterpret_cast<unsigned long long*>(e[k]); }
  if( stay ) { migrate_stay_kernel<<<unsigned((stay + 255) / 256), 256, 0, ctx->stream>>>(unsigned(stay), unsigned(soff[me]), unsigned(roff[me]), perm, F, types, et8); XSB_LAUNCH_CHECK(ctx); }
  if( nl ) { migrate_pack_kernel<<<unsigned((nl + 255) / 256), 256, 0, ctx->stream>>>(unsigned(nl), unsigned(soff[me]), unsigned(stay), perm, F, types, srec); XSB_LAUNCH_CHECK(ctx); }
  XSB_NCCL(ctx, g_nccl.GroupStart());
  for(int q = 0; q < P; q++)
  {
    if( q == me ) continue;
    const size_t sc = soff[q + 1] - soff[q], rcnt = roff[q + 1] - roff[q];
    const size_t cs = soff[q] - (q > me ? stay : 0), cr = roff[q] - (q > me ? stay : 0);      // offsets without the self segment
    if( sc ) XSB_NCCL(ctx, g_nccl.Send(srec + 8 * cs, 8 * sc, NCCL_UINT64, q, ctx->comm, ctx->stream));
    if( rcnt ) XSB_NCCL(ctx, g_nccl.Recv(rrec + 8 * cr, 8 * rcnt, NCCL_UINT64, q, ctx->comm, ctx->stream));
  }
  XSB_NCCL(ctx, g_nccl.GroupEnd());
  if( nin ) { migrate_unpack_kernel<<<unsigned((nin + 255) / 256), 256, 0, ctx->stream>>>(unsigned(nin), unsigned(roff[me]), unsigned(stay), rrec, F, et8); XSB_LAUNCH_CHECK(ctx); }
  ctx->migrated_out = n - stay; ctx->migrated_in = nn - stay;
  *n_new = unsigned(nn); *types_new = et8;
  return XSB_OK;
}

// sum-allreduce of `count` doubles in device memory (thermodynamic state: the reference's MPI_Allreduce(SUM))
int xsb_internal_allreduce_sum(xsb_ctx* ctx, double* dev_inout, int count)
{
  if( ctx->nranks == 1 ) return XSB_OK;
  XSB_REQUIRE(ctx, ctx->comm != nullptr, XSB_ERR_STATE, "xsb_comm_init must be called first");
  XSB_NCCL(ctx, g_nccl.AllReduce(dev_inout, dev_inout, size_t(count), NCCL_FLOAT64, NCCL_SUM, ctx->comm, ctx->stream));
  return XSB_OK;
}

int xsb_internal_allreduce_max(xsb_ctx* ctx, double* dev_inout, int count)
{
  if( ctx->nranks == 1 ) return XSB_OK;
  XSB_REQUIRE(ctx, ctx->comm != nullptr, XSB_ERR_STATE, "xsb_comm_init must be called first");
  XSB_NCCL(ctx, g_nccl.AllReduce(dev_inout, dev_inout, size_t(count), NCCL_FLOAT64, NCCL_MAX, ctx->comm, ctx->stream));
  return XSB_OK;
}

extern "C" {

// host-only view of the exchange plan (no context, no GPU): the receive list of brick `rank_coord`, i.e. for every
// ghost cell of its local grid that mirrors a domain cell: { ghost_cell, owner_rank, owner_cell (in the owner's local
// grid), wrap_x, wrap_y, wrap_z }, sorted by owner rank.  xsb_ghost_comm_scheme derives both its receive list and
// (from its peers' lists) its send lists from this function; tests drive it with gloo at world_size 2 on CPU.
int xsb_ghost_plan(const xsb_domain_desc* dom, int ghost_layers, const int32_t* rank_coord, int32_t* out6, uint64_t capacity, uint64_t* count)
{
  if( !dom || !rank_coord || !count || ghost_layers < 1 ) return XSB_ERR_INVALID;
  for(int a = 0; a < 3; a++)
    if( dom->rank_dims[a] < 1 || rank_coord[a] < 0 || rank_coord[a] >= dom->rank_dims[a] || dom->global_cells[a] < dom->rank_dims[a] ) return XSB_ERR_INVALID;
  std::vector<GhostCell> v;
  const int rc[3] = { rank_coord[0], rank_coord[1], rank_coord[2] };
  ghost_list(*dom, ghost_layers, rc, v);
  *count = v.size();
  if( !out6 ) return XSB_OK;
  if( capacity < v.size() ) return XSB_ERR_OVERFLOW;
  for(size_t i = 0; i < v.size(); i++)
  {
    int32_t* o = out6 + 6 * i;
    o[0] = v[i].ghost_cell; o[1] = v[i].owner_rank; o[2] = v[i].owner_cell; o[3] = v[i].w[0]; o[4] = v[i].w[1]; o[5] = v[i].w[2];
  }
  return XSB_OK;
}

int xsb_ghost_comm_scheme(xsb_ctx* ctx, const xsb_domain_desc* dom)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, dom != nullptr, XSB_ERR_INVALID, "null domain");
  XSB_REQUIRE(ctx, ctx->h_cell_off.size() == ctx->ncells + 1, XSB_ERR_STATE, "grid/particles not set");
  const int P = dom->rank_dims[0] * dom->rank_dims[1] * dom->rank_dims[2];
  XSB_REQUIRE(ctx, P == ctx->nranks, XSB_ERR_INVALID, "rank_dims product differs from the communicator size");
  const int me = dom->rank_coord[0] + dom->rank_dims[0] * (dom->rank_coord[1] + dom->rank_dims[1] * dom->rank_coord[2]);
  XSB_REQUIRE(ctx, me == ctx->rank, XSB_ERR_INVALID, "rank_coord does not match this rank (x fastest)");
  const int gl = ctx->grid.ghost_layers;
  for(int a = 0; a < 3; a++)
  {
    const int own = block_start(dom->rank_coord[a] + 1, dom->global_cells[a], dom->rank_dims[a]) - block_start(dom->rank_coord[a], dom->global_cells[a], dom->rank_dims[a]);
    XSB_REQUIRE(ctx, own + 2 * gl == ctx->grid.dims[a], XSB_ERR_INVALID, "local grid dims do not match the brick of this rank");
    XSB_REQUIRE(ctx, dom->global_cells[a] >= dom->rank_dims[a], XSB_ERR_INVALID, "fewer cells than ranks along an axis");
  }
  XSB_CUDA(ctx, cudaSetDevice(ctx->device));
  if( !ctx->ghost ) ctx->ghost = new GhostState;
  GhostState* G = ctx->ghost;
  G->dom = *dom; G->nranks = P;
  for(int c = 0; c < 27; c++) { const int w[3] = { c % 3 - 1, (c / 3) % 3 - 1, c / 9 - 1 }; for(int a = 0; a < 3; a++) G->shift[c][a] = 0.0; (void)w; }

  // receive list (mine) and send lists (entries of every peer's receive list that I own)
  if( G->plan_gl != gl || std::memcmp(&G->plan_dom, dom, sizeof(*dom)) != 0 || int(G->plan_to_peer.size()) != P )
  {
    ghost_list(*dom, gl, dom->rank_coord, G->plan_mine);
    G->plan_to_peer.assign(P, std::vector<GhostCell>());
    for(int q = 0; q < P; q++)
    {
      const int qc[3] = { q % dom->rank_dims[0], (q / dom->rank_dims[0]) % dom->rank_dims[1], q / (dom->rank_dims[0] * dom->rank_dims[1]) };
      std::vector<GhostCell> theirs; ghost_list(*dom, gl, qc, theirs);
      for(const GhostCell& g : theirs) if( g.owner_rank == me ) G->plan_to_peer[q].push_back(g);
    }
    G->plan_dom = *dom; G->plan_gl = gl;
  }
  const std::vector<GhostCell>& mine = G->plan_mine;
  const std::vector< std::vector<GhostCell> >& to_peer = G->plan_to_peer;
  const GridView gv = ctx->view();
  const std::vector<uint64_t>& off = ctx->h_cell_off;
  for(int q = 0; q < P; q++) for(const GhostCell& g : to_peer[q])
    XSB_REQUIRE(ctx, g.owner_cell >= 0 && uint64_t(g.owner_cell) < ctx->ncells && !gv.is_ghost_cell(unsigned(g.owner_cell)), XSB_ERR_STATE, "ghost scheme: owner cell is not an own cell");

  // particle counts: mine are known, peers' travel as u32 arrays (one per peer, cells in list order)
  std::vector<unsigned> send_counts, recv_counts(mine.size(), 0), sc_off(P + 1, 0), rc_off(P + 1, 0);
  for(int q = 0; q < P; q++) { for(const GhostCell& g : to_peer[q]) send_counts.push_back(unsigned(off[g.owner_cell + 1] - off[g.owner_cell])); sc_off[q + 1] = unsigned(send_counts.size()); }
  { size_t i = 0; for(int q = 0; q < P; q++) { while( i < mine.size() && mine[i].owner_rank == q ) ++i; rc_off[q + 1] = unsigned(i); } }
  for(unsigned i = rc_off[me]; i < rc_off[me + 1]; i++) recv_counts[i] = unsigned(off[mine[i].owner_cell + 1] - off[mine[i].owner_cell]);
  if( P > 1 )
  {
    XSB_REQUIRE(ctx, ctx->comm != nullptr, XSB_ERR_STATE, "multi-rank ghost scheme needs xsb_comm_init");
    XSB_CUDA(ctx, ctx->tmp32a.reserve(send_counts.size() + 16)); XSB_CUDA(ctx, ctx->tmp32b.reserve(recv_counts.size() + 16));
    if( !send_counts.empty() ) XSB_CUDA(ctx, cudaMemcpyAsync(ctx->tmp32a.p, send_counts.data(), send_counts.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    XSB_NCCL(ctx, g_nccl.GroupStart());
    for(int q = 0; q < P; q++)
    {
      if( q == me ) continue;
      // counts are sent as bytes so that no particular integer NCCL type value is assumed
      if( sc_off[q+1] > sc_off[q] ) XSB_NCCL(ctx, g_nccl.Send(ctx->tmp32a.p + sc_off[q], size_t(sc_off[q+1] - sc_off[q]) * 4, /*ncclUint8*/ 1, q, ctx->comm, ctx->stream));
      if( rc_off[q+1] > rc_off[q] ) XSB_NCCL(ctx, g_nccl.Recv(ctx->tmp32b.p + rc_off[q], size_t(rc_off[q+1] - rc_off[q]) * 4, /*ncclUint8*/ 1, q, ctx->comm, ctx->stream));
    }
    XSB_NCCL(ctx, g_nccl.GroupEnd());
    std::vector<unsigned> tmp(recv_counts.size());
    if( !tmp.empty() ) XSB_CUDA(ctx, cudaMemcpyAsync(tmp.data(), ctx->tmp32b.p, tmp.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for(int q = 0; q < P; q++) if( q != me ) for(unsigned i = rc_off[q]; i < rc_off[q+1]; i++) recv_counts[i] = tmp[i];
  }

  // new layout: own cells keep their counts, ghost cells take the received ones (others become empty)
  std::vector<uint64_t> cnt(ctx->ncells, 0), new_off(ctx->ncells + 1, 0);
  for(uint64_t c = 0; c < ctx->ncells; c++) if( !gv.is_ghost_cell(unsigned(c)) ) cnt[c] = off[c + 1] - off[c];
  for(size_t i = 0; i < mine.size(); i++) cnt[mine[i].ghost_cell] = recv_counts[i];
  for(uint64_t c = 0; c < ctx->ncells; c++) new_off[c + 1] = new_off[c] + cnt[c];
  int rc = xsb_internal_relayout(ctx, new_off.data()); if( rc ) return rc;

  // particle-level lists
  const double box[3] = { dom->box[0], dom->box[1], dom->box[2] };
  for(int c = 0; c < 27; c++) { G->shift[c][0] = (c % 3 - 1) * box[0]; G->shift[c][1] = ((c / 3) % 3 - 1) * box[1]; G->shift[c][2] = (c / 9 - 1) * box[2]; }
  std::vector<RangeEntry> es, er;
  G->send_off.assign(P + 1, 0); G->recv_off.assign(P + 1, 0);
  unsigned ns = 0, nr = 0;
  for(int q = 0; q < P; q++)
  {
    for(const GhostCell& g : to_peer[q])
    {
      XSB_REQUIRE(ctx, std::abs(g.w[0]) <= 1 && std::abs(g.w[1]) <= 1 && std::abs(g.w[2]) <= 1, XSB_ERR_UNSUPPORTED, "ghost layers wider than the periodic domain");
      const unsigned cnt = unsigned(new_off[g.owner_cell + 1] - new_off[g.owner_cell]);
      if( cnt ) es.push_back(RangeEntry{ unsigned(new_off[g.owner_cell]), cnt, ns, unsigned(q) | (unsigned(shift_code(g.w)) << 16) });
      ns += cnt;
    }
    G->send_off[q + 1] = ns;
  }
  for(int q = 0; q < P; q++)
  {
    for(unsigned i = rc_off[q]; i < rc_off[q + 1]; i++)
    {
      const unsigned cnt = unsigned(new_off[mine[i].ghost_cell + 1] - new_off[mine[i].ghost_cell]);
      if( cnt ) er.push_back(RangeEntry{ unsigned(new_off[mine[i].ghost_cell]), cnt, nr, unsigned(q) });
      nr += cnt;
    }
    G->recv_off[q + 1] = nr;
  }
  G->n_send = ns; G->n_recv = nr;
  XSB_REQUIRE(ctx, G->send_off[me + 1] - G->send_off[me] == G->recv_off[me + 1] - G->recv_off[me], XSB_ERR_STATE, "ghost scheme: self segment mismatch");
  XSB_CUDA(ctx, G->send_idx.reserve(size_t(ns) + 16, XSB_GROW_GHOST)); XSB_CUDA(ctx, G->send_code.reserve(size_t(ns) + 16, XSB_GROW_GHOST)); XSB_CUDA(ctx, G->recv_idx.reserve(size_t(nr) + 16, XSB_GROW_GHOST));
  XSB_CUDA(ctx, ctx->gseg_send.reserve(size_t(ns) + 16, XSB_GROW_GHOST)); XSB_CUDA(ctx, ctx->goff_send.reserve(P + 2));
  XSB_CUDA(ctx, ctx->gseg_recv.reserve(size_t(nr) + 16, XSB_GROW_GHOST)); XSB_CUDA(ctx, ctx->goff_recv.reserve(P + 2));
  {
    // per-cell entries travel (a few thousand), the per-particle lists are expanded on the device
    const size_t words = (es.size() + er.size()) * (sizeof(RangeEntry) / 8) + 2;
    XSB_CUDA(ctx, ctx->tmp64.reserve(words + 16, 1.5));
    RangeEntry* d_es = reinterpret_cast<RangeEntry*>(ctx->tmp64.p); RangeEntry* d_er = d_es + es.size();
    if( !es.empty() )
    {
      XSB_CUDA(ctx, cudaMemcpyAsync(d_es, es.data(), es.size() * sizeof(RangeEntry), cudaMemcpyHostToDevice, ctx->stream));
      expand_ranges_kernel<<<unsigned((es.size() * 32 + 255) / 256), 256, 0, ctx->stream>>>(unsigned(es.size()), d_es, G->send_idx.p, G->send_code.p, ctx->gseg_send.p);
      XSB_LAUNCH_CHECK(ctx);
    }
    if( !er.empty() )
    {
      XSB_CUDA(ctx, cudaMemcpyAsync(d_er, er.data(), er.size() * sizeof(RangeEntry), cudaMemcpyHostToDevice, ctx->stream));
      expand_ranges_kernel<<<unsigned((er.size() * 32 + 255) / 256), 256, 0, ctx->stream>>>(unsigned(er.size()), d_er, G->recv_idx.p, nullptr, ctx->gseg_recv.p);
      XSB_LAUNCH_CHECK(ctx);
    }
  }
  XSB_CUDA(ctx, cudaMemcpyAsync(ctx->goff_send.p, G->send_off.data(), (P + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  XSB_CUDA(ctx, cudaMemcpyAsync(ctx->goff_recv.p, G->recv_off.data(), (P + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if( P > 1 ) { int rcp = p2p_prepare(ctx, G); if( rcp ) return rcp; }
  ctx->ghost_valid = true;
  // ghost_update_all_no_fv: every persistent field travels once
  return exchange(ctx, (1u << XSB_F_RX) | (1u << XSB_F_RY) | (1u << XSB_F_RZ) | (1u << XSB_F_VX) | (1u << XSB_F_VY) | (1u << XSB_F_VZ) | (1u << XSB_F_TYPE) | (1u << XSB_F_ID), false);
}

// which transport the exchanges of the current scheme use: "p2p" (peer memory over NVLink, CUDA IPC), "nccl: <why>", "self"
int xsb_ghost_transport(xsb_ctx* ctx, char* out, size_t len)
{
  if( !ctx || !out || len == 0 ) return XSB_ERR_INVALID;
  std::string t = "self";
  if( ctx->ghost && ctx->ghost->nranks > 1 )
    t = (ctx->ghost->p2p_ok && ctx->ghost->p2p_fit) ? "p2p" : ("nccl: " + (ctx->ghost->p2p_ok ? std::string("exchange larger than the peer buffers") : ctx->ghost->p2p_why));
  std::strncpy(out, t.c_str(), len - 1); out[len - 1] = 0;
  return XSB_OK;
}

int xsb_ghost_update(xsb_ctx* ctx, uint32_t field_mask)
{
  XSB_ENTER(ctx);
  ctx->prof_begin(XSB_PROF_GHOST);
  const int rc = exchange(ctx, field_mask, false);
  ctx->prof_end(XSB_PROF_GHOST);
  return rc;
}

int xsb_ghost_reduce_add(xsb_ctx* ctx, uint32_t field_mask)
{
  XSB_ENTER(ctx);
  ctx->prof_begin(XSB_PROF_GHOST);
  const int rc = exchange(ctx, field_mask, true);
  ctx->prof_end(XSB_PROF_GHOST);
  return rc;
}

} // extern "C"
