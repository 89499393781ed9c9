// xsb_ctx.h -- internal state of one libxsb200 context (one per GPU).  Not part of the C ABI.
#pragma once
#include "../../include/xsb200.h"
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

struct ncclComm;

namespace xsb
{

// Head-room of the per-particle / per-ghost device arrays on (re)allocation.  Particle and ghost counts of a brick drift by
// a fraction of a percent between rebuilds (migration, density fluctuations): with too little slack a rebuild in the middle
// of a run hits cudaFree + cudaMalloc, which synchronises the device and costs milliseconds on ONE rank -- exactly the kind
// of straggler a 20-step timing window cannot absorb.  A B200 has 180 GB; the bench's working set is ~10 GB.
constexpr double XSB_GROW = 1.10, XSB_GROW_GHOST = 1.25;
constexpr int XSB_DISPL_RING = 8;

// RAII-less device buffer: grows geometrically, never shrinks until the context dies (steady-state MD loops
// must not touch cudaMalloc).
template<class T> struct DevBuf
{
  T* p = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t n, double factor = 1.0)
  {
    if( n <= cap ) return cudaSuccess;
    size_t want = size_t(double(n) * (factor < 1.0 ? 1.0 : factor));
    if( want < n ) want = n;
    if( p ) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if( e == cudaSuccess ) cap = want;
    return e;
  }
  // grow but keep the first `cap` elements (reserve() discards contents)
  cudaError_t reserve_keep(size_t n, double factor, cudaStream_t st)
  {
    if( n <= cap ) return cudaSuccess;
    size_t want = size_t(double(n) * (factor < 1.0 ? 1.0 : factor)); if( want < n ) want = n;
    T* q = nullptr;
    cudaError_t e = cudaMalloc((void**)&q, want * sizeof(T));
    if( e != cudaSuccess ) return e;
    if( p ) { e = cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st); cudaStreamSynchronize(st); cudaFree(p); }
    p = q; cap = want;
    return e;
  }
  void release() { if( p ) cudaFree(p); p = nullptr; cap = 0; }
};

// device-side view of the grid handed to kernels by value
struct GridView
{
  int nx, ny, nz, gl;
  double cell_size;
  double xf[9];
  int xform_identity;
  __host__ __device__ inline bool is_ghost_cell(unsigned c) const
  {
    const int i = int(c % unsigned(nx)), j = int((c / unsigned(nx)) % unsigned(ny)), k = int(c / (unsigned(nx) * unsigned(ny)));
    return i < gl || i >= nx - gl || j < gl || j >= ny - gl || k < gl || k >= nz - gl;
  }
};

// Device-resident control block of the inner-skin sub-list (xsb_eam.cu): the integrator kernels add every step's largest
// displacement (all-reduced over the ranks) to `acc`; the first EAM pass of a step re-filters the neighbour list (mode 0) when
// acc exceeds half the inner skin or the host demands it, otherwise it only re-evaluates the entries of the sub-list it left
// earlier (mode 1).  The decision never touches the host.
struct SubCtl { double acc; int mode; int pad_; unsigned long long builds, reuses; };

struct GhostState;
struct XFormInv { double m[9]; int identity; };

struct EamAlloyDev
{
  int nelements = 0, nr = 0, nrho = 0;
  double rdr = 0, rdrho = 0, rc = 0, rhomax = 0, conv_z2r = 0, conv_frho = 0;
  DevBuf<double> frho;      // reference layout [nel][nrho+1][8]
  DevBuf<double> rtab;      // r-tables in the reference layout: rhor [nel][nr+1][8] then z2r [npairs][nr+1][8]
  DevBuf<double> fc;        // the same r-tables as Hermite knots {f, c5} : [nel + npairs][nr+1][2] (xsb_eam.cu)
  DevBuf<float> fc32;       // mixed precision: knot pairs {f[m], c5[m], f[m+1], c5[m+1]} as floats: [nel + npairs][nr+1][4]
  bool set = false;
};

} // namespace xsb

struct xsb_ctx
{
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  int sm_count = 148;

  // a1 grid + particles
  bool grid_set = false;
  xsb_grid_desc grid{};
  uint64_t ncells = 0, n = 0;         // all cells / all particles (ghosts included)
  std::vector<uint64_t> h_cell_off;   // host copy
  xsb::DevBuf<unsigned> cell_start;   // [ncells+1] flat start of each cell
  xsb::DevBuf<unsigned> cell_of;      // [n] cell of each particle
  xsb::DevBuf<unsigned> own_atoms;    // flat indices of particles in non-ghost cells
  uint64_t n_own = 0;
  xsb::DevBuf<double> f64[XSB_F_TYPE]; // RX..RHO_DEMB (virial: 9n)
  xsb::DevBuf<uint8_t> type;
  xsb::DevBuf<uint64_t> id;
  bool virial_allocated = false;

  // a2 neighbour list (flat CSR, canonical order)
  bool nbh_built = false;
  double nbh_dist = 0.0;
  xsb_chunk_neighbors_config nbh_cfg{1, 1, 1, 0, 1.05};
  xsb::DevBuf<unsigned> nbh_count;            // [n]
  xsb::DevBuf<unsigned long long> nbh_off;    // [n+1]
  xsb::DevBuf<unsigned> nbh_idx;              // [total]
  uint64_t nbh_total = 0;
  unsigned nbh_max = 0;
  // tile-local view of the same list (xsb_tile.cuh): uint16 stage indices, same offsets as nbh_off
  bool tile_ok = false;                       // false -> force operators use the generic CSR-gather kernels
  bool tile_deal = false;                     // A/B switch (env XSB_TILE_DEAL=1): tl_idx in bank-dealt order (xsb_nbr.cu)
  int tile_TX = 1, tile_R[3] = {1, 1, 1};
  unsigned tile_align = 2;                    // row alignment (atoms) of the stage the list was built for (TileGeom::align)
  unsigned tile_s_cap = 0;                    // largest stage (atoms) over all tiles at build time
  double nbh_d2min = 0.0;                     // smallest pair distance^2 in the list at build time
  xsb::DevBuf<unsigned short> tl_idx;         // [total]
  xsb::DevBuf<unsigned> nbh_masks;            // build scratch: survivor masks of the count sweep (xsb_nbr.cu)
  // in-range sub-list written by the first pair pass of a step (EAM rho/emb) for the second (force): valid only while
  // positions, grid and cutoff are unchanged -- pos_epoch counts every API call that can move a particle
  xsb::DevBuf<unsigned short> sub_idx;        // [total]
  xsb::DevBuf<unsigned> sub_cnt;              // [n]
  xsb::DevBuf<double> pair_w;                 // [total] per-pair value cached by the pass that wrote the sub-list
  int sub_pw_kind = 0;                        // what pair_w holds for the current sub-list: 0 nothing, 1 eam_alloy rho'(r), 2 johnson rho'(r), 3 eam_alloy multi-element (rhojp, rhoip)
  double sub_pw_johnson[20] = {};             // parameter set (+ model id) behind a kind-2 cache
  int subcell_bits = 0;                       // particles of a cell sorted along a Morton curve of 2^bits sub-cells per axis when binning (env XSB_SUBCELL_SORT)
  int exp_tpa = 0;                            // A/B switch (env XSB_TPA=8): lanes per central atom of the non-virial FP64 force passes
  bool pair_cache_off = false;                // env XSB_NO_PAIR_CACHE=1: second pass re-evaluates instead (A/B profiling)
  uint64_t pos_epoch = 1, sub_epoch = 0;
  double sub_rcut = 0.0; bool sub_ghost = false;
  // inner skin: the sub-list keeps pairs up to rcut + inner_skin so that it can serve several steps (see SubCtl)
  double inner_skin = 0.0;                    // env XSB_INNER_SKIN (angstrom); 0 = filter the full list every step
  xsb::DevBuf<xsb::SubCtl> sub_ctl;
  uint64_t foreign_epoch = 1, sub_foreign = 0;   // position changes the displacement accounting does not see (uploads, re-binning, xform); value the sub-list was built under
  double sub_list_rc = 0.0;                   // membership radius of the current sub-list
  int sub_tables_id = 0, tables_id = 1, sub_pw_kind_built = 0;       // which table set / list build the sub-list belongs to
  bool pos_external = false;                  // a position device pointer was handed out: epochs cannot be trusted
  bool type_external = false;                 // same for the type bytes (the per-pair cache of a multi-element pass depends on them)
  bool sub_valid(double rcut, bool need_ghost) const
  { return !pos_external && sub_epoch == pos_epoch && sub_rcut == rcut && (sub_ghost || !need_ghost); }
  // a later operator of the same step with a SHORTER cut-off (a pair potential chained behind an EAM operator) may walk the
  // sub-list instead of the full neighbour list: it holds every pair within sub_rcut on the current positions
  bool sub_covers(double rcut, bool need_ghost) const
  { return !pos_external && !pair_sub_off && sub_epoch == pos_epoch && rcut <= sub_rcut && (sub_ghost || !need_ghost); }
  bool pair_sub_off = false;                  // env XSB_PAIR_NO_SUBLIST=1 (A/B)
  // Operator chain `compute_force: [eam_alloy_force, <pot>_multi_force]` (configs[4]): the force phase of eam_alloy_force is
  // not launched by its own call but by the NEXT entry point on this context -- on its own (any entry: XSB_ENTER), or, when
  // that entry is a pair operator with a cut-off <= the EAM one and the same flags, as ONE pass that evaluates both
  // potentials on the pairs it visits (positions gathered once instead of twice).  Calls were asynchronous already, so the
  // deferral is invisible; env XSB_NO_CHAIN_FUSION=1 launches at once (A/B).
  struct PendingEamForce { bool active = false; double rcut = 0.0; int phases = 0, flags = 0; } pending_eam;
  bool chain_fusion = true;
  uint64_t fused_chains = 0;                  // how many pair operators went into an EAM force pass (xsb_chain_stats)
  xsb::DevBuf<unsigned char> scratch;         // cub temp storage etc.
  xsb::DevBuf<unsigned long long> scratch64;  // misc u64 scratch

  // a8
  xsb::EamAlloyDev eam;

  // a9
  void* snap = nullptr;                       // xsb::SnapDev (xsb_snap.cu), owned by the context

  // a10
  ncclComm* comm = nullptr; int nranks = 1, rank = 0;
  xsb::GhostState* ghost = nullptr;           // built by xsb_ghost_comm_scheme (xsb_ghost.cu)
  bool ghost_valid = false;                   // its lists index the current particle layout
  xsb::DevBuf<unsigned> old_cell_start;       // relayout scratch
  xsb::DevBuf<unsigned long long> tmp64;      // relayout / sort scratch (8-byte words)
  xsb::DevBuf<unsigned> tmp32a, tmp32b, tmp32c, tmp32d;
  xsb::DevBuf<unsigned> gseg_send, gseg_recv, goff_send, goff_recv;   // ghost segment tables (device)
  xsb::DevBuf<double> move_stage; xsb::DevBuf<unsigned char> move_stage8;   // move_particles staging (persistent)
  xsb::DevBuf<double> move_stage_b, move_stage_c; xsb::DevBuf<unsigned char> move_stage8_b, move_stage8_c;   // cross-rank migration
  uint64_t migrated_out = 0, migrated_in = 0;   // particles that changed rank in the last xsb_particles_rebin
  uint64_t otb_clamped = 0;                     // particles the last binning clamped into a border cell (xsb_out_of_domain_count)
  xsb::DevBuf<double> backup; unsigned backup_n = 0xffffffffu;        // backup_r positions of own particles

  // xsb_verlet_boundary_async / xsb_displ_poll: ring of {max displacement^2, max step displacement^2} results
  double* displ_host = nullptr;               // pinned, [XSB_DISPL_RING][2]
  xsb::DevBuf<unsigned long long> displ_dev;
  cudaEvent_t displ_ev[8] = {};
  uint64_t displ_seq = 0;

  // asynchronous host <-> device field transfers (xsb_fields_upload_async / _download_async, xsb_core.cu): one copy stream
  // per direction so H2D and D2H use both DMA engines while the compute stream runs the passes
  cudaStream_t copy_up = nullptr, copy_down = nullptr;
  cudaEvent_t ev_up_done = nullptr, ev_up_free = nullptr, ev_down_ready = nullptr, ev_down_done = nullptr;
  bool up_pending = false, down_pending = false;
  xsb::DevBuf<double> stage_up, stage_down;

  // recorded steps (xsb_step_capture_begin / _end / xsb_step_replay, xsb_core.cu): the kernels a sequence of operator calls
  // enqueues, kept as an instantiated CUDA graph and re-issued with one launch
  struct StepGraph { cudaGraphExec_t exec = nullptr; uint64_t nodes = 0, gen = 0; bool verlet = false; };
  std::vector<StepGraph> step_graphs;
  bool capturing = false, cap_verlet = false;
  uint64_t cap_launch0 = 0;
  uint64_t graph_gen = 1;                     // bumped by everything that changes what a recorded step baked in (layout, list, cell matrix)

  // per-operator device timing (CUDA events on this context's stream), see xsb_profile_*
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev[XSB_PROF_COUNT_];   // begin/end pairs
  size_t prof_used[XSB_PROF_COUNT_] = {};
  cudaEvent_t timer_ev[2] = { nullptr, nullptr };
  void prof_begin(int tag)
  {
    if( !prof_on || capturing ) return;
    if( prof_used[tag] + 2 > prof_ev[tag].size() ) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); prof_ev[tag].push_back(a); prof_ev[tag].push_back(b); }
    cudaEventRecord(prof_ev[tag][prof_used[tag]], stream);
  }
  void prof_end(int tag)
  {
    if( !prof_on || capturing ) return;
    cudaEventRecord(prof_ev[tag][prof_used[tag] + 1], stream);
    prof_used[tag] += 2;
  }

  int fail(int code, const char* fmt, ...)
  {
    char buf[1024]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    err = buf; return code;
  }
  xsb::GridView view() const
  {
    xsb::GridView v; v.nx = grid.dims[0]; v.ny = grid.dims[1]; v.nz = grid.dims[2]; v.gl = grid.ghost_layers;
    v.cell_size = grid.cell_size; for(int i = 0; i < 9; i++) v.xf[i] = grid.xform[i]; v.xform_identity = grid.xform_is_identity;
    return v;
  }
};

#define XSB_CUDA(ctx, call) do { cudaError_t e__ = (call); if( e__ != cudaSuccess ) \
  return (ctx)->fail(XSB_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); } while(0)
// first statement of every extern "C" entry that touches the stream or allocates: one context per GPU, several contexts
// (devices) may live in one process, so the calling thread's current device must follow the context
// XSB_ENTER also launches a deferred EAM force phase (xsb_ctx::pending_eam); the pair operators, which may absorb it, use
// XSB_ENTER_KEEP and decide themselves
#define XSB_ENTER_KEEP(ctx) do { if( !(ctx) || !(ctx)->stream ) return XSB_ERR_STATE; XSB_CUDA(ctx, cudaSetDevice((ctx)->device)); } while(0)
#define XSB_ENTER(ctx) do { XSB_ENTER_KEEP(ctx); if( (ctx)->pending_eam.active ) { int rc__ = xsb_internal_flush_pending(ctx); if( rc__ ) return rc__; } } while(0)
#define XSB_LAUNCH_CHECK(ctx) do { (ctx)->launches++; cudaError_t e__ = cudaGetLastError(); if( e__ != cudaSuccess ) \
  return (ctx)->fail(XSB_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); } while(0)
#define XSB_REQUIRE(ctx, cond, code, msg) do { if( !(cond) ) return (ctx)->fail(code, "%s", msg); } while(0)

// internal helpers shared across translation units (C++ linkage, not exported by the header)
int  xsb_internal_flush_pending(xsb_ctx* ctx);                       // xsb_eam.cu: launch the deferred EAM force phase on its own
int  xsb_internal_ensure_virial(xsb_ctx* ctx);
int  xsb_internal_install_cells(xsb_ctx* ctx, const uint64_t* cell_off);
int  xsb_internal_relayout(xsb_ctx* ctx, const uint64_t* new_cell_off);
void xsb_ghost_release(xsb_ctx* ctx);
int  xsb_internal_assign_device(xsb_ctx* ctx, unsigned n, double* rx, double* ry, double* rz, const double* vx, const double* vy, const double* vz,
                                const unsigned char* type, const unsigned long long* id, const int wrap[3], const double box[3]);
// cross-rank part of move_particles (xsb_ghost.cu): n staged particles in d[0..6] (+types) -> particles this rank owns now
int  xsb_internal_migrate(xsb_ctx* ctx, const xsb_domain_desc* dom, unsigned n, double* const d[7], unsigned char* types,
                          unsigned* n_new, double* e[7], unsigned char** types_new);
void xsb_snap_release(xsb_ctx* ctx);
// inner-skin budget (SubCtl): every atom may have moved by up to `displacement` more (xsb_assign.cu)
int  xsb_internal_sub_account(xsb_ctx* ctx, double displacement);
// the same with the squared displacement read from device memory (all-reduced over the ranks first)
int  xsb_internal_sub_account_dev(xsb_ctx* ctx, unsigned long long* s2_dev);
int  xsb_internal_allreduce_max(xsb_ctx* ctx, double* dev_inout, int count);      // xsb_ghost.cu
// pinned result ring of xsb_verlet_boundary_async (xsb_assign.cu); idempotent
int  xsb_internal_displ_ring_init(xsb_ctx* ctx);

namespace xsb
{
// search range (cell layers per axis) covering every grid-space displacement of physical length < dist
void search_range(const xsb_grid_desc& g, double dist, int R[3]);
void search_range_unclamped(const xsb_grid_desc& g, double dist, int R[3]);
void inverse3(const double* m, double* inv);
}
