// xsb_core.cu -- context, grid + particle storage, zero_force_energy.
#include "xsb_ctx.h"
#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <new>

namespace xsb
{

void inverse3(const double* m, double* inv)
{
  const double det = m[0]*(m[4]*m[8]-m[5]*m[7]) - m[1]*(m[3]*m[8]-m[5]*m[6]) + m[2]*(m[3]*m[7]-m[4]*m[6]);
  const double id = 1.0 / det;
  inv[0] =  (m[4]*m[8]-m[5]*m[7])*id; inv[1] = -(m[1]*m[8]-m[2]*m[7])*id; inv[2] =  (m[1]*m[5]-m[2]*m[4])*id;
  inv[3] = -(m[3]*m[8]-m[5]*m[6])*id; inv[4] =  (m[0]*m[8]-m[2]*m[6])*id; inv[5] = -(m[0]*m[5]-m[2]*m[3])*id;
  inv[6] =  (m[3]*m[7]-m[4]*m[6])*id; inv[7] = -(m[0]*m[7]-m[1]*m[6])*id; inv[8] =  (m[0]*m[4]-m[1]*m[3])*id;
}

void search_range_unclamped(const xsb_grid_desc& g, double dist, int R[3])
{
  double inv[9] = {1,0,0,0,1,0,0,0,1};
  if( !g.xform_is_identity ) inverse3(g.xform, inv);
  for(int a = 0; a < 3; a++)
  {
    const double nrm = std::sqrt(inv[3*a]*inv[3*a] + inv[3*a+1]*inv[3*a+1] + inv[3*a+2]*inv[3*a+2]);
    R[a] = int(std::ceil(dist * nrm / g.cell_size));
    if( R[a] < 1 ) R[a] = 1;
  }
}

void search_range(const xsb_grid_desc& g, double dist, int R[3])
{
  search_range_unclamped(g, dist, R);
  for(int a = 0; a < 3; a++) if( R[a] > 15 ) R[a] = 15;   // range of the 5-bit relative cell index of the exported stream (callers reject more)
}

__global__ void zero_fields_kernel(const unsigned* __restrict__ atoms, unsigned n_atoms, unsigned n_all, bool all,
                                   double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz,
                                   double* __restrict__ ep, double* __restrict__ vir)
{
  const unsigned stride = gridDim.x * blockDim.x;
  const unsigned n = all ? n_all : n_atoms;
  for(unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
  {
    const unsigned a = all ? t : atoms[t];
    fx[a] = 0.0; fy[a] = 0.0; fz[a] = 0.0; ep[a] = 0.0;
    if( vir ) { double* v = vir + 9ull * a; for(int k = 0; k < 9; k++) v[k] = 0.0; }
  }
}

// own-only transfers: staging[f][t] <-> field_f[own_atoms[t]]
struct XferFields { double* p[8]; int nf; };
__global__ void xfer_scatter_kernel(unsigned n, const unsigned* __restrict__ atoms, XferFields F, const double* __restrict__ stage, size_t pitch)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if( t >= n ) return;
  const unsigned a = atoms[t];
  for(int f = 0; f < F.nf; f++) F.p[f][a] = stage[size_t(f) * pitch + t];
}
// the same scatter for an upload that brings all three position fields while an inner-skin sub-list is alive: the largest
// physical displacement |X (r_new - r_old)|^2 of the own atoms goes to *s2 (atomicMax on the bits of a non-negative double)
__global__ void xfer_scatter_displ_kernel(unsigned n, const unsigned* __restrict__ atoms, XferFields F, const double* __restrict__ stage, size_t pitch,
                                          int ix, int iy, int iz, XFormInv Xf, unsigned long long* __restrict__ s2)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  double d2 = 0.0;
  if( t < n )
  {
    const unsigned a = atoms[t];
    double ex = stage[size_t(ix) * pitch + t] - F.p[ix][a], ey = stage[size_t(iy) * pitch + t] - F.p[iy][a], ez = stage[size_t(iz) * pitch + t] - F.p[iz][a];
    if( !Xf.identity )
    {
      const double x = Xf.m[0]*ex + Xf.m[1]*ey + Xf.m[2]*ez, y = Xf.m[3]*ex + Xf.m[4]*ey + Xf.m[5]*ez, z = Xf.m[6]*ex + Xf.m[7]*ey + Xf.m[8]*ez;
      ex = x; ey = y; ez = z;
    }
    d2 = ex*ex + ey*ey + ez*ez;
    for(int f = 0; f < F.nf; f++) F.p[f][a] = stage[size_t(f) * pitch + t];
  }
  for(int o = 16; o > 0; o >>= 1) d2 = fmax(d2, __shfl_xor_sync(0xffffffffu, d2, o));
  __shared__ double s[8];
  if( (threadIdx.x & 31) == 0 ) s[threadIdx.x >> 5] = d2;
  __syncthreads();
  if( threadIdx.x == 0 )
  {
    double m = s[0];
    for(unsigned w = 1; w < (blockDim.x >> 5); w++) m = fmax(m, s[w]);
    if( !(m <= 0.0) ) atomicMax(s2, m == m ? (unsigned long long)__double_as_longlong(m) : 0x7ff0000000000000ull);      // NaN input: infinite displacement
  }
}
__global__ void xfer_gather_kernel(unsigned n, const unsigned* __restrict__ atoms, XferFields F, double* __restrict__ stage, size_t pitch)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if( t >= n ) return;
  const unsigned a = atoms[t];
  for(int f = 0; f < F.nf; f++) stage[size_t(f) * pitch + t] = F.p[f][a];
}

} // namespace xsb

using namespace xsb;

static int xfer_setup(xsb_ctx* ctx)
{
  if( ctx->copy_up ) return XSB_OK;
  XSB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_up, cudaStreamNonBlocking));
  XSB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_down, cudaStreamNonBlocking));
  cudaEvent_t* ev[4] = { &ctx->ev_up_done, &ctx->ev_up_free, &ctx->ev_down_ready, &ctx->ev_down_done };
  for(int i = 0; i < 4; i++) XSB_CUDA(ctx, cudaEventCreateWithFlags(ev[i], cudaEventDisableTiming));
  return XSB_OK;
}

static int xfer_fields(xsb_ctx* ctx, int nfields, const int* fields, xsb::XferFields& F)
{
  XSB_REQUIRE(ctx, nfields >= 1 && nfields <= 8 && fields != nullptr, XSB_ERR_INVALID, "1..8 fields per asynchronous transfer");
  XSB_REQUIRE(ctx, ctx->h_cell_off.size() == ctx->ncells + 1, XSB_ERR_STATE, "xsb_particles_set_cells must be called first");
  F.nf = nfields;
  for(int k = 0; k < nfields; k++)
  {
    const int f = fields[k];
    XSB_REQUIRE(ctx, f >= 0 && f < XSB_F_TYPE && f != XSB_F_VIRIAL, XSB_ERR_INVALID, "asynchronous transfers carry the scalar double fields (r, f, ep, v, rho_dEmb)");
    F.p[k] = ctx->f64[f].p;
  }
  return XSB_OK;
}

int xsb_internal_ensure_virial(xsb_ctx* ctx)
{
  if( ctx->virial_allocated && ctx->f64[XSB_F_VIRIAL].cap >= 9 * (ctx->n + 1) ) return XSB_OK;
  XSB_CUDA(ctx, ctx->f64[XSB_F_VIRIAL].reserve(9 * (ctx->n + 1), XSB_GROW));
  XSB_CUDA(ctx, cudaMemsetAsync(ctx->f64[XSB_F_VIRIAL].p, 0, 9 * (ctx->n + 1) * sizeof(double), ctx->stream));
  ctx->virial_allocated = true;
  return XSB_OK;
}

namespace xsb
{
// one warp per cell: cell_of[] of its particles and, for non-ghost cells, their slots in the own-particle list
__global__ void __launch_bounds__(256) cell_tables_kernel(unsigned ncells, const unsigned* __restrict__ cell_start, const unsigned* __restrict__ own_prefix,
                                                          unsigned* __restrict__ cell_of, unsigned* __restrict__ own_atoms)
{
  const unsigned c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
  if( c >= ncells ) return;
  const unsigned s = cell_start[c], e = cell_start[c + 1], o = own_prefix[c];
  const bool own = own_prefix[c + 1] - o == e - s;   // ghost cells add nothing to the own-particle prefix
  for(unsigned p = s + lane; p < e; p += 32u)
  {
    cell_of[p] = c;
    if( own ) own_atoms[o + (p - s)] = p;
  }
}
}

// cell tables (flat start, cell of each particle, list of particles in non-ghost cells): the per-cell prefix sums are
// done on the host (ncells entries), everything per particle on the device
int xsb_internal_install_cells(xsb_ctx* ctx, const uint64_t* off)
{
  ctx->graph_gen++;
  const uint64_t nc = ctx->ncells, n = off[nc];
  ctx->h_cell_off.assign(off, off + nc + 1);
  std::vector<unsigned> tab(2 * (nc + 1));            // [0, nc] cell start ; [nc+1, 2nc+1] own-particle prefix
  unsigned* start = tab.data(); unsigned* ownp = tab.data() + nc + 1;
  const GridView gv = ctx->view();
  uint64_t nown = 0;
  for(uint64_t c = 0; c < nc; c++)
  {
    start[c] = unsigned(off[c]); ownp[c] = unsigned(nown);
    if( !gv.is_ghost_cell(unsigned(c)) ) nown += off[c + 1] - off[c];
  }
  start[nc] = unsigned(off[nc]); ownp[nc] = unsigned(nown);
  ctx->n = n; ctx->n_own = nown; ctx->pos_epoch++; ctx->foreign_epoch++;
  ctx->backup_n = 0xffffffffu;      // the own particles were re-ordered: a backup_r of the old order compares different atoms
  ctx->ghost_valid = false;         // exchange lists index the old layout (xsb_ghost_comm_scheme sets it again)
  XSB_CUDA(ctx, ctx->cell_start.reserve(2 * (nc + 1)));
  XSB_CUDA(ctx, ctx->cell_of.reserve(n + 1, XSB_GROW));
  XSB_CUDA(ctx, ctx->own_atoms.reserve(nown + 1, XSB_GROW));
  XSB_CUDA(ctx, cudaMemcpyAsync(ctx->cell_start.p, tab.data(), tab.size() * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
  if( n )
  {
    xsb::cell_tables_kernel<<<unsigned((nc * 32 + 255) / 256), 256, 0, ctx->stream>>>(unsigned(nc), ctx->cell_start.p, ctx->cell_start.p + nc + 1,
                                                                                     ctx->cell_of.p, ctx->own_atoms.p);
    XSB_LAUNCH_CHECK(ctx);
  }
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // `tab` dies here
  ctx->nbh_built = false;
  return XSB_OK;
}

namespace xsb
{
template<class T>
__global__ void relayout_kernel(unsigned n_new, GridView g, const unsigned* __restrict__ cell_of, const unsigned* __restrict__ new_start,
                                const unsigned* __restrict__ old_start, const T* __restrict__ in, T* __restrict__ out)
{
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i >= n_new ) return;
  const unsigned c = cell_of[i];
  out[i] = g.is_ghost_cell(c) ? T(0) : in[old_start[c] + (i - new_start[c])];
}
}

// new cell offsets where every NON-ghost cell keeps its particle count: persistent fields (r, v, type, id) of own
// cells move to their new flat position, ghost slots are zero-filled, accumulators (f, ep, virial, rho_dEmb) are zeroed.
int xsb_internal_relayout(xsb_ctx* ctx, const uint64_t* new_off)
{
  const uint64_t nc = ctx->ncells;
  const GridView gv = ctx->view();
  for(uint64_t c = 0; c < nc; c++)
    if( !gv.is_ghost_cell(unsigned(c)) && new_off[c+1] - new_off[c] != ctx->h_cell_off[c+1] - ctx->h_cell_off[c] )
      return ctx->fail(XSB_ERR_INVALID, "relayout: particle count of own cell %llu changed", (unsigned long long)c);
  XSB_REQUIRE(ctx, new_off[nc] < 0xFFFFFFF0ull, XSB_ERR_OVERFLOW, "more than 2^32 particles per GPU");
  XSB_CUDA(ctx, ctx->old_cell_start.reserve(nc + 1));
  XSB_CUDA(ctx, cudaMemcpyAsync(ctx->old_cell_start.p, ctx->cell_start.p, (nc + 1) * sizeof(unsigned), cudaMemcpyDeviceToDevice, ctx->stream));
  int rc = xsb_internal_install_cells(ctx, new_off); if( rc ) return rc;
  const unsigned n = unsigned(ctx->n);
  const unsigned grid = (n + 255) / 256;
  XSB_CUDA(ctx, ctx->tmp64.reserve(size_t(n) + 16, XSB_GROW));
  const int moved[6] = { XSB_F_RX, XSB_F_RY, XSB_F_RZ, XSB_F_VX, XSB_F_VY, XSB_F_VZ };
  for(int k = 0; k < 6 && n; k++)
  {
    DevBuf<double>& b = ctx->f64[moved[k]];
    relayout_kernel<double><<<grid, 256, 0, ctx->stream>>>(n, gv, ctx->cell_of.p, ctx->cell_start.p, ctx->old_cell_start.p, b.p, reinterpret_cast<double*>(ctx->tmp64.p));
    XSB_LAUNCH_CHECK(ctx);
    XSB_CUDA(ctx, b.reserve_keep(n + 16, XSB_GROW, ctx->stream));      // stage rows are read up to the next 16-atom boundary
    XSB_CUDA(ctx, cudaMemcpyAsync(b.p, ctx->tmp64.p, size_t(n) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  if( n )
  {
    relayout_kernel<unsigned long long><<<grid, 256, 0, ctx->stream>>>(n, gv, ctx->cell_of.p, ctx->cell_start.p, ctx->old_cell_start.p,
                                                                       reinterpret_cast<const unsigned long long*>(ctx->id.p), ctx->tmp64.p);
    XSB_LAUNCH_CHECK(ctx);
    XSB_CUDA(ctx, ctx->id.reserve_keep(n + 1, XSB_GROW, ctx->stream));
    XSB_CUDA(ctx, cudaMemcpyAsync(ctx->id.p, ctx->tmp64.p, size_t(n) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    relayout_kernel<unsigned char><<<grid, 256, 0, ctx->stream>>>(n, gv, ctx->cell_of.p, ctx->cell_start.p, ctx->old_cell_start.p, ctx->type.p,
                                                                  reinterpret_cast<unsigned char*>(ctx->tmp64.p));
    XSB_LAUNCH_CHECK(ctx);
    XSB_CUDA(ctx, ctx->type.reserve_keep(n + 16, XSB_GROW, ctx->stream));
    XSB_CUDA(ctx, cudaMemcpyAsync(ctx->type.p, ctx->tmp64.p, size_t(n), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  const int zeroed[6] = { XSB_F_FX, XSB_F_FY, XSB_F_FZ, XSB_F_EP, XSB_F_RHO_DEMB, XSB_F_VIRIAL };
  for(int k = 0; k < 6; k++)
  {
    const int f = zeroed[k];
    if( f == XSB_F_VIRIAL && !ctx->virial_allocated ) continue;
    const size_t w = f == XSB_F_VIRIAL ? 9 : 1;
    XSB_CUDA(ctx, ctx->f64[f].reserve(w * (size_t(n) + 16), XSB_GROW));
    XSB_CUDA(ctx, cudaMemsetAsync(ctx->f64[f].p, 0, w * (size_t(n) + 1) * sizeof(double), ctx->stream));
  }
  return XSB_OK;
}

namespace xsb
{
// roofline denominators: 8 independent FMA chains per thread, 4096 steps
template<class real>
__global__ void __launch_bounds__(256) peak_fma_kernel(real* out, real a, real b, int iters)
{
  real v[8];
# pragma unroll
  for(int k = 0; k < 8; k++) v[k] = real(threadIdx.x + k);
  for(int i = 0; i < iters; i++)
  {
#   pragma unroll
    for(int k = 0; k < 8; k++) v[k] = v[k] * a + b;
  }
  real s = 0;
# pragma unroll
  for(int k = 0; k < 8; k++) s += v[k];
  if( s == real(-1) ) out[blockIdx.x * blockDim.x + threadIdx.x] = s;     // never true: keeps the chains alive
}
__global__ void __launch_bounds__(256) peak_copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n)
{
  for(size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) out[i] = in[i];
}
}

extern "C" {

const char* xsb_version(void) { return "xsb200 0.1 (sm_100a)"; }

int xsb_create(int device, xsb_ctx** out)
{
  if( !out ) return XSB_ERR_INVALID;
  *out = nullptr;
  xsb_ctx* ctx = new (std::nothrow) xsb_ctx;
  if( !ctx ) return XSB_ERR_INVALID;
  *out = ctx;   // returned even on failure so xsb_last_error() can be read; caller still must xsb_destroy it
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if( e != cudaSuccess || ndev == 0 )
    return ctx->fail(XSB_ERR_CUDA, "no CUDA device available (%s); libxsb200 has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if( device < 0 || device >= ndev ) return ctx->fail(XSB_ERR_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
  cudaDeviceProp prop;
  XSB_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
  if( prop.major != 10 )
    return ctx->fail(XSB_ERR_CUDA, "device %d is sm_%d%d; libxsb200 is built for sm_100a only", device, prop.major, prop.minor);
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->tile_deal = getenv("XSB_TILE_DEAL") != nullptr;   // A/B switches for profiling, not a fallback: same kernels
  ctx->pair_cache_off = getenv("XSB_NO_PAIR_CACHE") != nullptr;
  ctx->pair_sub_off = getenv("XSB_PAIR_NO_SUBLIST") != nullptr;
  ctx->chain_fusion = getenv("XSB_NO_CHAIN_FUSION") == nullptr;
  ctx->exp_tpa = getenv("XSB_TPA") ? atoi(getenv("XSB_TPA")) : 0;
  ctx->inner_skin = getenv("XSB_INNER_SKIN") ? std::max(0.0, atof(getenv("XSB_INNER_SKIN"))) : 0.0;
  ctx->subcell_bits = getenv("XSB_SUBCELL_SORT") ? std::min(3, std::max(0, atoi(getenv("XSB_SUBCELL_SORT")))) : 0;
  XSB_CUDA(ctx, cudaSetDevice(device));
  XSB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  return XSB_OK;
}

void xsb_destroy(xsb_ctx* ctx)
{
  if( !ctx ) return;
  if( ctx->stream )
  {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
  }
  ctx->cell_start.release(); ctx->cell_of.release(); ctx->own_atoms.release();
  for(auto& b : ctx->f64) b.release();
  ctx->type.release(); ctx->id.release();
  ctx->nbh_count.release(); ctx->nbh_off.release(); ctx->nbh_idx.release(); ctx->nbh_masks.release(); ctx->scratch.release(); ctx->scratch64.release();
  ctx->eam.frho.release(); ctx->eam.rtab.release(); ctx->eam.fc.release(); ctx->eam.fc32.release(); ctx->tl_idx.release(); ctx->sub_idx.release(); ctx->sub_cnt.release(); ctx->pair_w.release(); ctx->move_stage.release(); ctx->move_stage8.release();
  xsb_ghost_release(ctx);
  xsb_snap_release(ctx);
  ctx->stage_up.release(); ctx->stage_down.release(); ctx->displ_dev.release(); ctx->sub_ctl.release();
  if( ctx->displ_host ) { cudaFreeHost(ctx->displ_host); for(cudaEvent_t e : ctx->displ_ev) if( e ) cudaEventDestroy(e); }
  if( ctx->copy_up ) { cudaStreamSynchronize(ctx->copy_up); cudaStreamDestroy(ctx->copy_up); }
  if( ctx->copy_down ) { cudaStreamSynchronize(ctx->copy_down); cudaStreamDestroy(ctx->copy_down); }
  for(cudaEvent_t e : { ctx->ev_up_done, ctx->ev_up_free, ctx->ev_down_ready, ctx->ev_down_done }) if( e ) cudaEventDestroy(e);
  for(auto& v : ctx->prof_ev) for(cudaEvent_t e : v) cudaEventDestroy(e);
  for(auto& sg : ctx->step_graphs) if( sg.exec ) cudaGraphExecDestroy(sg.exec);
  if( ctx->stream ) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* xsb_last_error(const xsb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
uint64_t xsb_kernel_launch_count(const xsb_ctx* ctx)
{
  if( !ctx ) return 0;
  // a deferred EAM force phase (xsb_ctx::pending_eam) counts as launched by the call that requested it
  if( ctx->pending_eam.active && ctx->stream && cudaSetDevice(ctx->device) == cudaSuccess ) xsb_internal_flush_pending(const_cast<xsb_ctx*>(ctx));
  return ctx->launches;
}

int xsb_profile_enable(xsb_ctx* ctx, int on)
{
  XSB_ENTER(ctx);
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->prof_on = on != 0;
  for(int t = 0; t < XSB_PROF_COUNT_; t++) ctx->prof_used[t] = 0;
  return XSB_OK;
}

int xsb_profile_read(xsb_ctx* ctx, int tag, double* ms_total, uint64_t* intervals)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, tag >= 0 && tag < XSB_PROF_COUNT_, XSB_ERR_INVALID, "unknown profile tag");
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double tot = 0.0;
  for(size_t i = 0; i + 1 < ctx->prof_used[tag]; i += 2)
  {
    float ms = 0.f;
    XSB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->prof_ev[tag][i], ctx->prof_ev[tag][i + 1]));
    tot += ms;
  }
  if( ms_total ) *ms_total = tot;
  if( intervals ) *intervals = ctx->prof_used[tag] / 2;
  return XSB_OK;
}

// two-slot stopwatch on the context's stream: slot 0 = start, slot 1 = stop
int xsb_timer_record(xsb_ctx* ctx, int slot)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, slot == 0 || slot == 1, XSB_ERR_INVALID, "timer slot must be 0 or 1");
  if( !ctx->timer_ev[slot] ) XSB_CUDA(ctx, cudaEventCreate(&ctx->timer_ev[slot]));
  XSB_CUDA(ctx, cudaEventRecord(ctx->timer_ev[slot], ctx->stream));
  return XSB_OK;
}

int xsb_measure_peaks(xsb_ctx* ctx, double* fp64_tflops, double* fp32_tflops, double* hbm_gbs)
{
  XSB_ENTER(ctx);
  XSB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaEvent_t e0, e1;
  XSB_CUDA(ctx, cudaEventCreate(&e0)); XSB_CUDA(ctx, cudaEventCreate(&e1));
  const int iters = 4096, grid = ctx->sm_count * 8, block = 256;
  auto best_ms = [&](auto launch) -> double
  {
    double best = 1.0e30;
    for(int rep = 0; rep < 6; rep++)
    {
      cudaEventRecord(e0, ctx->stream); launch(); cudaEventRecord(e1, ctx->stream); cudaEventSynchronize(e1);
      float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
      if( rep > 0 && ms < best ) best = ms;
    }
    return best;
  };
  xsb::DevBuf<double> buf;
  XSB_CUDA(ctx, buf.reserve(size_t(1) << 28));          // 2 GiB: source and destination halves of 1 GiB each
  const double flop = 2.0 * 8.0 * iters * double(grid) * block;
  if( fp64_tflops ) { const double ms = best_ms([&]{ xsb::peak_fma_kernel<double><<<grid, block, 0, ctx->stream>>>(buf.p, 1.0000001, 1.0e-9, iters); }); *fp64_tflops = flop / (ms * 1.0e-3) * 1.0e-12; ctx->launches += 6; }
  if( fp32_tflops ) { const double ms = best_ms([&]{ xsb::peak_fma_kernel<float><<<grid, block, 0, ctx->stream>>>(reinterpret_cast<float*>(buf.p), 1.0000001f, 1.0e-9f, iters); }); *fp32_tflops = flop / (ms * 1.0e-3) * 1.0e-12; ctx->launches += 6; }
  if( hbm_gbs )
  {
    const size_t n16 = (size_t(1) << 30) / 16;
    const double ms = best_ms([&]{ xsb::peak_copy_kernel<<<ctx->sm_count * 16, block, 0, ctx->stream>>>(reinterpret_cast<const double2*>(buf.p), reinterpret_cast<double2*>(buf.p) + n16, n16); });
    *hbm_gbs = 2.0 * double(size_t(1) << 30) / (ms * 1.0e-3) * 1.0e-9; ctx->launches += 6;
  }
  cudaError_t e = cudaGetLastError();
  buf.release(); cudaEventDestroy(e0); cudaEventDestroy(e1);
  if( e != cudaSuccess ) return ctx->fail(XSB_ERR_CUDA, "xsb_measure_peaks: %s", cudaGetErrorString(e));
  return XSB_OK;
}

int xsb_timer_elapsed_ms(xsb_ctx* ctx, double* ms)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, ms && ctx->timer_ev[0] && ctx->timer_ev[1], XSB_ERR_STATE, "timer not recorded");
  XSB_CUDA(ctx, cudaEventSynchronize(ctx->timer_ev[1]));
  float f = 0.f;
  XSB_CUDA(ctx, cudaEventElapsedTime(&f, ctx->timer_ev[0], ctx->timer_ev[1]));
  *ms = f;
  return XSB_OK;
}

// ---- recorded steps -----------------------------------------------------------------------------------------------------
// A small system (C1: 131 k atoms, 0.1 ms of kernels per step) is bound by the launch path, not by the kernels: the operator
// calls of a regular step (integrator pass, ghost update, zero, force operators) are recorded once per neighbour-list
// generation as a CUDA graph and re-issued with one launch.  Everything an entry point passes to a kernel by value is frozen
// at recording time, so a recorded step is valid until the particle layout, the list or the cell matrix change (graph_gen).
int xsb_step_capture_begin(xsb_ctx* ctx)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, !ctx->capturing, XSB_ERR_STATE, "a step is already being recorded");
  XSB_REQUIRE(ctx, ctx->nranks == 1, XSB_ERR_UNSUPPORTED, "recorded steps are limited to one rank (the peer-memory ghost exchange passes its epoch by value)");
  XSB_REQUIRE(ctx, ctx->inner_skin == 0.0, XSB_ERR_UNSUPPORTED, "recorded steps and the EAM inner skin (host-side epochs choose the pass) exclude each other");
  int rc = xsb_internal_displ_ring_init(ctx); if( rc ) return rc;      // nothing may be allocated while the stream is capturing
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->cap_launch0 = ctx->launches; ctx->cap_verlet = false;
  XSB_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  ctx->capturing = true;
  return XSB_OK;
}

int xsb_step_capture_end(xsb_ctx* ctx, int* step_id)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, ctx->capturing, XSB_ERR_STATE, "xsb_step_capture_begin must be called first");
  ctx->capturing = false;
  const uint64_t nodes = ctx->launches - ctx->cap_launch0;
  ctx->launches = ctx->cap_launch0;                        // recorded, not executed
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
  if( e != cudaSuccess || g == nullptr )
  {
    cudaGetLastError();
    return ctx->fail(XSB_ERR_CUDA, "recording failed (%s): an entry point that waits for the device, reads a result back or allocates was called while recording",
                     cudaGetErrorString(e));
  }
  xsb_ctx::StepGraph sg; sg.nodes = nodes; sg.gen = ctx->graph_gen; sg.verlet = ctx->cap_verlet;
  e = cudaGraphInstantiate(&sg.exec, g, 0);
  cudaGraphDestroy(g);
  if( e != cudaSuccess ) return ctx->fail(XSB_ERR_CUDA, "cudaGraphInstantiate -> %s", cudaGetErrorString(e));
  int id = -1;
  for(size_t i = 0; i < ctx->step_graphs.size(); i++) if( !ctx->step_graphs[i].exec ) { id = int(i); break; }
  if( id < 0 ) { id = int(ctx->step_graphs.size()); ctx->step_graphs.push_back(sg); } else ctx->step_graphs[size_t(id)] = sg;
  if( step_id ) *step_id = id;
  return XSB_OK;
}

int xsb_step_replay(xsb_ctx* ctx, int step_id)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, !ctx->capturing, XSB_ERR_STATE, "a step is being recorded");
  XSB_REQUIRE(ctx, step_id >= 0 && size_t(step_id) < ctx->step_graphs.size() && ctx->step_graphs[size_t(step_id)].exec, XSB_ERR_INVALID, "unknown recorded step");
  const xsb_ctx::StepGraph& sg = ctx->step_graphs[size_t(step_id)];
  XSB_REQUIRE(ctx, sg.gen == ctx->graph_gen, XSB_ERR_STATE, "the particle layout, the neighbour list or the cell matrix changed since this step was recorded");
  XSB_CUDA(ctx, cudaGraphLaunch(sg.exec, ctx->stream));
  ctx->launches += sg.nodes;
  // host-side bookkeeping the recorded calls would have done: positions moved (any cached sub-list is stale for later
  // direct calls), and the integrator pass's maxima go into the result ring
  ctx->pos_epoch += 2; ctx->foreign_epoch++;
  if( sg.verlet )
  {
    const int slot = int(ctx->displ_seq % XSB_DISPL_RING);
    XSB_CUDA(ctx, cudaMemcpyAsync(ctx->displ_host + 2 * slot, ctx->displ_dev.p + 2 * XSB_DISPL_RING, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    XSB_CUDA(ctx, cudaEventRecord(ctx->displ_ev[slot], ctx->stream));
    ctx->displ_seq++;
  }
  return XSB_OK;
}

int xsb_step_release(xsb_ctx* ctx, int step_id)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, step_id >= 0 && size_t(step_id) < ctx->step_graphs.size() && ctx->step_graphs[size_t(step_id)].exec, XSB_ERR_INVALID, "unknown recorded step");
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaGraphExecDestroy(ctx->step_graphs[size_t(step_id)].exec);
  ctx->step_graphs[size_t(step_id)] = xsb_ctx::StepGraph{};
  return XSB_OK;
}

int xsb_sync(xsb_ctx* ctx)
{
  XSB_ENTER(ctx);
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return XSB_OK;
}

int xsb_grid_set(xsb_ctx* ctx, const xsb_grid_desc* g)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, g != nullptr, XSB_ERR_INVALID, "null grid");
  ctx->graph_gen++;
  XSB_REQUIRE(ctx, g->dims[0] > 0 && g->dims[1] > 0 && g->dims[2] > 0, XSB_ERR_INVALID, "grid dims must be positive");
  XSB_REQUIRE(ctx, g->ghost_layers >= 0 && 2 * g->ghost_layers < g->dims[0] && 2 * g->ghost_layers < g->dims[1] && 2 * g->ghost_layers < g->dims[2],
              XSB_ERR_INVALID, "ghost_layers inconsistent with dims");
  XSB_REQUIRE(ctx, g->cell_size > 0.0, XSB_ERR_INVALID, "cell_size must be > 0");
  const uint64_t nc = uint64_t(g->dims[0]) * uint64_t(g->dims[1]) * uint64_t(g->dims[2]);
  XSB_REQUIRE(ctx, nc < (1ull << 31), XSB_ERR_OVERFLOW, "too many cells");
  ctx->grid = *g; ctx->pos_epoch++; ctx->foreign_epoch++;
  ctx->ncells = nc;
  ctx->grid_set = true;
  ctx->nbh_built = false;
  ctx->n = 0;
  return XSB_OK;
}

int xsb_grid_set_xform(xsb_ctx* ctx, const double xform[9])
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, xform != nullptr, XSB_ERR_INVALID, "null xform");
  XSB_REQUIRE(ctx, ctx->grid_set, XSB_ERR_STATE, "xsb_grid_set must be called first");
  ctx->graph_gen++;
  bool ident = true;
  double dx2 = 0.0, inv[9] = {1,0,0,0,1,0,0,0,1}, inv2 = 0.0;
  if( !ctx->grid.xform_is_identity ) xsb::inverse3(ctx->grid.xform, inv);
  for(int i = 0; i < 9; i++)
  {
    XSB_REQUIRE(ctx, std::isfinite(xform[i]), XSB_ERR_INVALID, "xform is not finite");
    dx2 += (xform[i] - ctx->grid.xform[i]) * (xform[i] - ctx->grid.xform[i]); inv2 += inv[i] * inv[i];
    ctx->grid.xform[i] = xform[i];
    ident = ident && xform[i] == ((i % 4 == 0) ? 1.0 : 0.0);
  }
  ctx->grid.xform_is_identity = ident ? 1 : 0;
  ctx->pos_epoch++;                 // physical distances changed: the in-range sub-list of this step must be rewritten
  // Inner-skin accounting (SubCtl): a listed pair (|X r| < nbh_dist) changes its separation by at most |dX| |X^-1| nbh_dist
  // (Frobenius norms); that counts like both atoms moving half of it.  A barostat's per-step drift (1e-6 relative) costs
  // 1e-5 ang of the budget; a real deformation exhausts it and the next rho phase re-filters.
  if( ctx->sub_ctl.p && ctx->nbh_built ) { int rc = xsb_internal_sub_account(ctx, 0.5 * std::sqrt(dx2 * inv2) * ctx->nbh_dist); if( rc ) return rc; }
  else ctx->foreign_epoch++;
  return XSB_OK;
}

int xsb_particles_set_cells(xsb_ctx* ctx, const uint64_t* off)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, ctx->grid_set, XSB_ERR_STATE, "xsb_grid_set must be called first");
  XSB_REQUIRE(ctx, off != nullptr && off[0] == 0, XSB_ERR_INVALID, "cell_particle_offset[0] must be 0");
  const uint64_t nc = ctx->ncells;
  for(uint64_t c = 0; c < nc; c++) XSB_REQUIRE(ctx, off[c+1] >= off[c], XSB_ERR_INVALID, "cell_particle_offset must be non-decreasing");
  XSB_REQUIRE(ctx, off[nc] < 0xFFFFFFF0ull, XSB_ERR_OVERFLOW, "more than 2^32 particles per GPU");
  XSB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = xsb_internal_install_cells(ctx, off); if( rc ) return rc;
  const uint64_t n = ctx->n;
  for(int f = 0; f < XSB_F_TYPE; f++)
  {
    if( f == XSB_F_VIRIAL && !ctx->virial_allocated ) continue;   // allocated on first use
    const size_t w = f == XSB_F_VIRIAL ? 9 : 1;
    XSB_CUDA(ctx, ctx->f64[f].reserve(w * (n + 16), XSB_GROW));
    XSB_CUDA(ctx, cudaMemsetAsync(ctx->f64[f].p, 0, w * (n + 1) * sizeof(double), ctx->stream));
  }
  XSB_CUDA(ctx, ctx->type.reserve(n + 16, XSB_GROW));
  XSB_CUDA(ctx, cudaMemsetAsync(ctx->type.p, 0, n + 16, ctx->stream));
  XSB_CUDA(ctx, ctx->id.reserve(n + 1, XSB_GROW));
  XSB_CUDA(ctx, cudaMemsetAsync(ctx->id.p, 0, (n + 1) * sizeof(uint64_t), ctx->stream));
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return XSB_OK;
}

int xsb_cell_offsets_download(xsb_ctx* ctx, uint64_t* off)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, off != nullptr && ctx->h_cell_off.size() == ctx->ncells + 1, XSB_ERR_STATE, "no particles");
  std::memcpy(off, ctx->h_cell_off.data(), (ctx->ncells + 1) * sizeof(uint64_t));
  return XSB_OK;
}

uint64_t xsb_num_own_particles(const xsb_ctx* ctx) { return ctx ? ctx->n_own : 0; }

uint64_t xsb_num_particles(const xsb_ctx* ctx) { return ctx ? ctx->n : 0; }
uint64_t xsb_num_cells(const xsb_ctx* ctx) { return ctx ? ctx->ncells : 0; }

static int field_ptr(xsb_ctx* ctx, int field, void** p, size_t* bytes)
{
  XSB_REQUIRE(ctx, field >= 0 && field < XSB_F_COUNT_, XSB_ERR_INVALID, "unknown field");
  XSB_REQUIRE(ctx, ctx->h_cell_off.size() == ctx->ncells + 1, XSB_ERR_STATE, "xsb_particles_set_cells must be called first");
  if( field == XSB_F_TYPE ) { *p = ctx->type.p; *bytes = ctx->n; }
  else if( field == XSB_F_ID ) { *p = ctx->id.p; *bytes = ctx->n * sizeof(uint64_t); }
  else if( field == XSB_F_VIRIAL ) { int rc = xsb_internal_ensure_virial(ctx); if( rc ) return rc; *p = ctx->f64[field].p; *bytes = 9 * ctx->n * sizeof(double); }
  else { *p = ctx->f64[field].p; *bytes = ctx->n * sizeof(double); }
  return XSB_OK;
}

int xsb_field_upload(xsb_ctx* ctx, int field, const void* src)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, src != nullptr, XSB_ERR_INVALID, "null source");
  void* p = nullptr; size_t bytes = 0;
  int rc = field_ptr(ctx, field, &p, &bytes); if( rc ) return rc;
  if( bytes ) XSB_CUDA(ctx, cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  if( field == XSB_F_RX || field == XSB_F_RY || field == XSB_F_RZ ) { ctx->pos_epoch++; ctx->foreign_epoch++; }
  if( field == XSB_F_TYPE ) ctx->sub_pw_kind = 0;      // cached per-pair values depend on the neighbour's element
  return XSB_OK;
}

int xsb_field_download(xsb_ctx* ctx, int field, void* dst)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, dst != nullptr, XSB_ERR_INVALID, "null destination");
  void* p = nullptr; size_t bytes = 0;
  int rc = field_ptr(ctx, field, &p, &bytes); if( rc ) return rc;
  if( bytes ) XSB_CUDA(ctx, cudaMemcpyAsync(dst, p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return XSB_OK;
}

// ---- asynchronous transfers ------------------------------------------------------------------------------------------
// upload: [copy_up] wait(staging free) -> H2D into staging -> record(up_done)
//         [stream]  wait(up_done) -> scatter staging -> fields -> record(staging free)
// The H2D overlaps with everything enqueued on the compute stream before this call (the force passes of the step in
// flight still read the old positions); work enqueued after it sees the new values.
int xsb_fields_upload_async(xsb_ctx* ctx, int nfields, const int* fields, const void* const* host_src, int own_only)
{
  XSB_ENTER(ctx);
  xsb::XferFields F; int rc = xfer_fields(ctx, nfields, fields, F); if( rc ) return rc;
  XSB_REQUIRE(ctx, host_src != nullptr, XSB_ERR_INVALID, "null host array list");
  if( (rc = xfer_setup(ctx)) ) return rc;
  const size_t n = own_only ? ctx->n_own : ctx->n;
  if( n == 0 ) return XSB_OK;
  const size_t pitch = (n + 15) & ~size_t(15);
  if( ctx->stage_up.cap < pitch * 8 )
  {
    if( ctx->up_pending ) XSB_CUDA(ctx, cudaEventSynchronize(ctx->ev_up_free));
    XSB_CUDA(ctx, ctx->stage_up.reserve(pitch * 8, 1.05));
  }
  if( ctx->up_pending ) XSB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_up, ctx->ev_up_free, 0));
  for(int k = 0; k < nfields; k++)
  {
    XSB_REQUIRE(ctx, host_src[k] != nullptr, XSB_ERR_INVALID, "null host array");
    XSB_CUDA(ctx, cudaMemcpyAsync(ctx->stage_up.p + size_t(k) * pitch, host_src[k], n * sizeof(double), cudaMemcpyHostToDevice, ctx->copy_up));
  }
  XSB_CUDA(ctx, cudaEventRecord(ctx->ev_up_done, ctx->copy_up));
  XSB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_up_done, 0));
  // positions of the own atoms arriving while an inner-skin sub-list is alive (a host application that integrates and
  // hands the new positions over every step): their largest displacement is charged to the device-side budget, exactly
  // as the integrator pass does, instead of forcing the next rho phase to re-filter the neighbour list
  int ip[3] = { -1, -1, -1 };
  bool any_pos = false;
  for(int k = 0; k < nfields; k++) for(int c = 0; c < 3; c++) if( fields[k] == XSB_F_RX + c ) { ip[c] = k; any_pos = true; }
  const bool account = own_only && ip[0] >= 0 && ip[1] >= 0 && ip[2] >= 0 && ctx->sub_ctl.p != nullptr && ctx->inner_skin > 0.0;
  if( account )
  {
    if( (rc = xsb_internal_displ_ring_init(ctx)) ) return rc;
    unsigned long long* s2 = ctx->displ_dev.p + 2 * (XSB_DISPL_RING + 1);
    XSB_CUDA(ctx, cudaMemsetAsync(s2, 0, sizeof(unsigned long long), ctx->stream));
    xsb::XFormInv Xf; Xf.identity = ctx->grid.xform_is_identity; for(int i = 0; i < 9; i++) Xf.m[i] = ctx->grid.xform[i];
    xsb::xfer_scatter_displ_kernel<<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(unsigned(n), ctx->own_atoms.p, F, ctx->stage_up.p, pitch, ip[0], ip[1], ip[2], Xf, s2);
    XSB_LAUNCH_CHECK(ctx);
    if( (rc = xsb_internal_sub_account_dev(ctx, s2)) ) return rc;
  }
  else if( own_only )
  {
    xsb::xfer_scatter_kernel<<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(unsigned(n), ctx->own_atoms.p, F, ctx->stage_up.p, pitch);
    XSB_LAUNCH_CHECK(ctx);
  }
  else
    for(int k = 0; k < nfields; k++) XSB_CUDA(ctx, cudaMemcpyAsync(F.p[k], ctx->stage_up.p + size_t(k) * pitch, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  XSB_CUDA(ctx, cudaEventRecord(ctx->ev_up_free, ctx->stream));
  ctx->up_pending = true;
  if( any_pos ) { ctx->pos_epoch++; if( !account ) ctx->foreign_epoch++; }
  return XSB_OK;
}

// download: [stream]    wait(previous D2H done) -> gather fields -> staging -> record(down_ready)
//           [copy_down] wait(down_ready) -> D2H -> record(down_done)
// The snapshot is taken at this point of the compute stream; the D2H overlaps with what is enqueued afterwards.
// The host arrays are valid after xsb_copy_wait().
int xsb_fields_download_async(xsb_ctx* ctx, int nfields, const int* fields, void* const* host_dst, int own_only)
{
  XSB_ENTER(ctx);
  xsb::XferFields F; int rc = xfer_fields(ctx, nfields, fields, F); if( rc ) return rc;
  XSB_REQUIRE(ctx, host_dst != nullptr, XSB_ERR_INVALID, "null host array list");
  if( (rc = xfer_setup(ctx)) ) return rc;
  const size_t n = own_only ? ctx->n_own : ctx->n;
  if( n == 0 ) return XSB_OK;
  const size_t pitch = (n + 15) & ~size_t(15);
  if( ctx->stage_down.cap < pitch * 8 )
  {
    if( ctx->down_pending ) XSB_CUDA(ctx, cudaEventSynchronize(ctx->ev_down_done));
    XSB_CUDA(ctx, ctx->stage_down.reserve(pitch * 8, 1.05));
  }
  if( ctx->down_pending ) XSB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_down_done, 0));
  if( own_only )
  {
    xsb::xfer_gather_kernel<<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(unsigned(n), ctx->own_atoms.p, F, ctx->stage_down.p, pitch);
    XSB_LAUNCH_CHECK(ctx);
  }
  else
    for(int k = 0; k < nfields; k++) XSB_CUDA(ctx, cudaMemcpyAsync(ctx->stage_down.p + size_t(k) * pitch, F.p[k], n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  XSB_CUDA(ctx, cudaEventRecord(ctx->ev_down_ready, ctx->stream));
  XSB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_down, ctx->ev_down_ready, 0));
  for(int k = 0; k < nfields; k++)
  {
    XSB_REQUIRE(ctx, host_dst[k] != nullptr, XSB_ERR_INVALID, "null host array");
    XSB_CUDA(ctx, cudaMemcpyAsync(host_dst[k], ctx->stage_down.p + size_t(k) * pitch, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_down));
  }
  XSB_CUDA(ctx, cudaEventRecord(ctx->ev_down_done, ctx->copy_down));
  ctx->down_pending = true;
  return XSB_OK;
}

int xsb_copy_wait(xsb_ctx* ctx)
{
  XSB_ENTER(ctx);
  if( ctx->down_pending ) XSB_CUDA(ctx, cudaEventSynchronize(ctx->ev_down_done));
  if( ctx->up_pending ) XSB_CUDA(ctx, cudaEventSynchronize(ctx->ev_up_free));
  return XSB_OK;
}

void* xsb_field_device_ptr(xsb_ctx* ctx, int field)
{
  if( !ctx || !ctx->stream || cudaSetDevice(ctx->device) != cudaSuccess ) return nullptr;
  if( ctx->pending_eam.active && xsb_internal_flush_pending(ctx) ) return nullptr;
  void* p = nullptr; size_t bytes = 0;
  if( field_ptr(ctx, field, &p, &bytes) ) return nullptr;
  if( field == XSB_F_RX || field == XSB_F_RY || field == XSB_F_RZ ) ctx->pos_external = true;   // caller may move particles behind our back
  if( field == XSB_F_TYPE ) ctx->type_external = true;
  return p;
}

int xsb_zero_force_energy(xsb_ctx* ctx, int ghost)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, ctx->h_cell_off.size() == ctx->ncells + 1, XSB_ERR_STATE, "no particles");
  if( ctx->n == 0 ) return XSB_OK;
  const unsigned n = ghost ? unsigned(ctx->n) : unsigned(ctx->n_own);
  if( n == 0 ) return XSB_OK;
  const int block = 256;
  const int grid = int(std::min<uint64_t>((n + block - 1) / block, uint64_t(ctx->sm_count) * 8));
  zero_fields_kernel<<<grid, block, 0, ctx->stream>>>(ctx->own_atoms.p, unsigned(ctx->n_own), unsigned(ctx->n), ghost != 0,
      ctx->f64[XSB_F_FX].p, ctx->f64[XSB_F_FY].p, ctx->f64[XSB_F_FZ].p, ctx->f64[XSB_F_EP].p,
      ctx->virial_allocated ? ctx->f64[XSB_F_VIRIAL].p : nullptr);
  XSB_LAUNCH_CHECK(ctx);
  return XSB_OK;
}

} // extern "C"
