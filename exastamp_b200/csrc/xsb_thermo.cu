// xsb_thermo.cu -- simulation_thermodynamic_state (reference
// src/thermo_state/simulation_thermodynamic_state.cpp:81-230): sums over the particles of the OWN cells of
// virial (9), 1/2 m v(x)v (9), m v (3), 1/2 m v^2 per axis (3), ep, mass, count = the reference's 27-double buffer,
// followed by the MPI_Allreduce(SUM) of the reference done as one ncclAllReduce.  HBM-streaming (~60 B/atom, 132 with
// virial); two kernels (per-block partials in a fixed grid, then one block folds them in index order) so that the
// result is bit-reproducible from run to run.
#include "xsb_ctx.h"

int xsb_internal_allreduce_sum(xsb_ctx* ctx, double* dev_inout, int count);   // xsb_ghost.cu

namespace xsb
{

struct MassTab16 { double mass[16]; };
constexpr int THERMO_N = 27, THERMO_BLOCK = 256;

__global__ void __launch_bounds__(THERMO_BLOCK) thermo_partial_kernel(unsigned n, const unsigned* __restrict__ atoms, MassTab16 M,
    const unsigned char* __restrict__ type, const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz,
    const double* __restrict__ ep, const double* __restrict__ vir, double* __restrict__ partial)
{
  double acc[THERMO_N];
# pragma unroll
  for(int k = 0; k < THERMO_N; k++) acc[k] = 0.0;
  for(unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
  {
    const unsigned a = atoms[t];
    const double m = M.mass[type[a] & 15];
    const double v[3] = { vx[a], vy[a], vz[a] };
    if( vir ) {
#     pragma unroll
      for(int k = 0; k < 9; k++) acc[k] += vir[size_t(a) * 9 + k];
    }
#   pragma unroll
    for(int i = 0; i < 3; i++)
    {
#     pragma unroll
      for(int j = 0; j < 3; j++) acc[9 + 3 * i + j] += v[i] * v[j] * m;
      acc[18 + i] += v[i] * m;
      acc[21 + i] += v[i] * v[i] * m;
    }
    acc[24] += ep[a];
    acc[25] += m;
    acc[26] += 1.0;
  }
  __shared__ double s[THERMO_BLOCK / 32][THERMO_N];
# pragma unroll
  for(int k = 0; k < THERMO_N; k++)
  {
    double x = acc[k];
    for(int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if( (threadIdx.x & 31) == 0 ) s[threadIdx.x >> 5][k] = x;
  }
  __syncthreads();
  if( threadIdx.x < THERMO_N )
  {
    double x = 0.0;
    for(int w = 0; w < THERMO_BLOCK / 32; w++) x += s[w][threadIdx.x];
    partial[size_t(blockIdx.x) * THERMO_N + threadIdx.x] = x;
  }
}

__global__ void thermo_fold_kernel(unsigned nblocks, const double* __restrict__ partial, double* __restrict__ out)
{
  if( threadIdx.x >= THERMO_N ) return;
  double x = 0.0;
  for(unsigned b = 0; b < nblocks; b++) x += partial[size_t(b) * THERMO_N + threadIdx.x];
  // the reference halves the kinetic sums after the loop (simulation_thermodynamic_state.cpp:152-153)
  if( threadIdx.x >= 9 && threadIdx.x < 18 ) x *= 0.5;
  if( threadIdx.x >= 21 && threadIdx.x < 24 ) x *= 0.5;
  out[threadIdx.x] = x;
}

} // namespace xsb

using namespace xsb;

extern "C" int xsb_thermo_state(xsb_ctx* ctx, int n_types, const double* mass, double* out27)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, out27 != nullptr && mass != nullptr, XSB_ERR_INVALID, "null argument");
  XSB_REQUIRE(ctx, n_types >= 1 && n_types <= 16, XSB_ERR_INVALID, "n_types must be in 1..16");
  XSB_CUDA(ctx, cudaSetDevice(ctx->device));
  const unsigned n = unsigned(ctx->n_own);
  const unsigned nblocks = unsigned(ctx->sm_count) * 4u;
  XSB_CUDA(ctx, ctx->scratch64.reserve(size_t(nblocks + 2) * THERMO_N));
  double* partial = reinterpret_cast<double*>(ctx->scratch64.p);
  double* out = partial + size_t(nblocks) * THERMO_N;
  MassTab16 M; for(int t = 0; t < 16; t++) M.mass[t] = mass[t < n_types ? t : n_types - 1];
  ctx->prof_begin(XSB_PROF_INTEGRATE);
  thermo_partial_kernel<<<nblocks, THERMO_BLOCK, 0, ctx->stream>>>(n, ctx->own_atoms.p, M, ctx->type.p, ctx->f64[XSB_F_VX].p, ctx->f64[XSB_F_VY].p,
      ctx->f64[XSB_F_VZ].p, ctx->f64[XSB_F_EP].p, ctx->virial_allocated ? ctx->f64[XSB_F_VIRIAL].p : nullptr, partial);
  XSB_LAUNCH_CHECK(ctx);
  thermo_fold_kernel<<<1, 32, 0, ctx->stream>>>(nblocks, partial, out);
  XSB_LAUNCH_CHECK(ctx);
  ctx->prof_end(XSB_PROF_INTEGRATE);
  if( ctx->nranks > 1 ) { int rc = xsb_internal_allreduce_sum(ctx, out, THERMO_N); if( rc ) return rc; }
  XSB_CUDA(ctx, cudaMemcpyAsync(out27, out, THERMO_N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return XSB_OK;
}
