// xsb_assign.cu -- getting particles into the cell grid on the device, and the per-particle operators either side
// of the force path: push_f_v_r / push_f_v (Verlet halves, ext exaNBody), force_to_accel
// (src/compute/force_to_accel.cu:82-101), backup_r + particle_displ_over (neighbour rebuild trigger,
// data/config/config_move_particles.msp:19-23).  All are pure HBM streaming kernels.
//
// xsb_particles_assign bins an unsorted particle set into the OWN cells of the local grid with a stable radix sort
// on the cell index: inside a cell particles keep their input order, so the layout is deterministic and equals the
// host statement in tests/helpers.py.
#include "xsb_ctx.h"
#include <algorithm>
#include <cmath>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_reduce.cuh>

namespace xsb
{

struct BinParams
{
  double ox, oy, oz, inv_cell, cell;   // corner of the first OWN cell, 1/cell_size
  int nx, ny, nz, gl;
  double box[3]; int wrap[3];          // periodic wrap (only when this rank spans the whole axis)
  int sub_bits;                        // > 0: order the particles of a cell along a Morton curve of 2^sub_bits sub-cells per axis
};

// 3-D Morton code of (x,y,z) with b bits each (b <= 3)
__device__ __forceinline__ unsigned morton3(unsigned x, unsigned y, unsigned z, int b)
{
  unsigned m = 0;
  for(int i = 0; i < b; i++) m |= (((x >> i) & 1u) << (3 * i)) | (((y >> i) & 1u) << (3 * i + 1)) | (((z >> i) & 1u) << (3 * i + 2));
  return m;
}

// cell of a coordinate along one axis.  A particle outside the own cells is clamped into the border cell (the
// reference's move_particles keeps such "otb" particles aside, ext exaNBody): `how` gets 1 when that happened and 2
// when the particle is more than one cell outside, which the callers turn into an error -- the list build's search
// range and its FP32 guard band (xsb_nbr.cu) assume atoms lie within a cell of the cell that holds them.
__device__ __forceinline__ int own_cell_coord(double r, double o, double cell, int n_own, int& how)
{
  int c = int(floor((r - o) / cell));
  if( c < 0 || c >= n_own ) how = max(how, (c < -1 || c > n_own) ? 2 : 1);
  return min(max(c, 0), n_own - 1);
}

__global__ void bin_kernel(unsigned n, BinParams B, double* __restrict__ rx, double* __restrict__ ry, double* __restrict__ rz,
                           unsigned* __restrict__ key, unsigned* __restrict__ val, unsigned* __restrict__ counts)
{
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i >= n ) return;
  double x = rx[i], y = ry[i], z = rz[i];
  if( B.wrap[0] ) { x -= floor((x - B.ox) / B.box[0]) * B.box[0]; if( x - B.ox >= B.box[0] ) x = B.ox; rx[i] = x; }
  if( B.wrap[1] ) { y -= floor((y - B.oy) / B.box[1]) * B.box[1]; if( y - B.oy >= B.box[1] ) y = B.oy; ry[i] = y; }
  if( B.wrap[2] ) { z -= floor((z - B.oz) / B.box[2]) * B.box[2]; if( z - B.oz >= B.box[2] ) z = B.oz; rz[i] = z; }
  int how = 0;
  const int ci = own_cell_coord(x, B.ox, B.cell, B.nx - 2 * B.gl, how) + B.gl;
  const int cj = own_cell_coord(y, B.oy, B.cell, B.ny - 2 * B.gl, how) + B.gl;
  const int ck = own_cell_coord(z, B.oz, B.cell, B.nz - 2 * B.gl, how) + B.gl;
  const unsigned c = unsigned(ci) + unsigned(B.nx) * (unsigned(cj) + unsigned(B.ny) * unsigned(ck));
  unsigned k = c;
  if( B.sub_bits > 0 )
  {
    // sub-cell of the particle inside its cell: neighbours that a cut-off sphere takes out of a cell then form a few
    // contiguous index runs instead of a random subset, so the position gathers of the pair passes hit consecutive banks
    const double S = double(1 << B.sub_bits);
    auto sub = [&](double r, double o, int cc) { const double f = (r - o) / B.cell - double(cc - B.gl); return unsigned(min(max(int(f * S), 0), (1 << B.sub_bits) - 1)); };
    k = (c << (3 * B.sub_bits)) | morton3(sub(x, B.ox, ci), sub(y, B.oy, cj), sub(z, B.oz, ck), B.sub_bits);
  }
  key[i] = k; val[i] = i;
  atomicAdd(counts + c, 1u);
  if( how ) atomicAdd(counts + B.nx * B.ny * B.nz + (how - 1), 1u);      // [ncells]: clamped, [ncells+1]: lost (> 1 cell outside)
}

template<class T>
__global__ void permute_kernel(unsigned n, const unsigned* __restrict__ perm, const T* __restrict__ in, T* __restrict__ out)
{
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i < n ) out[i] = in[perm[i]];
}

__global__ void iota64_kernel(unsigned n, unsigned long long* out)
{
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i < n ) out[i] = i;
}

// ---- Verlet pieces ---------------------------------------------------------------------------------------------
// push_f_v_r: r += v dt + 1/2 a dt^2 (a stored in f after force_to_accel).  Positions live in grid space, velocities
// and accelerations in physical space: dr_grid = xform^-1 * dr_phys (the reference's INV_XFORM variant).
__global__ void push_f_v_r_kernel(unsigned n, const unsigned* __restrict__ atoms, double dt, double dt2h, XFormInv Xi,
                                  double* __restrict__ rx, double* __restrict__ ry, double* __restrict__ rz,
                                  const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz,
                                  const double* __restrict__ ax, const double* __restrict__ ay, const double* __restrict__ az)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if( t >= n ) return;
  const unsigned a = atoms[t];
  double dx = vx[a] * dt + ax[a] * dt2h, dy = vy[a] * dt + ay[a] * dt2h, dz = vz[a] * dt + az[a] * dt2h;
  if( !Xi.identity )
  {
    const double x = Xi.m[0]*dx + Xi.m[1]*dy + Xi.m[2]*dz, y = Xi.m[3]*dx + Xi.m[4]*dy + Xi.m[5]*dz, z = Xi.m[6]*dx + Xi.m[7]*dy + Xi.m[8]*dz;
    dx = x; dy = y; dz = z;
  }
  rx[a] += dx; ry[a] += dy; rz[a] += dz;
}

__global__ void push_f_v_kernel(unsigned n, const unsigned* __restrict__ atoms, double dth,
                                double* __restrict__ vx, double* __restrict__ vy, double* __restrict__ vz,
                                const double* __restrict__ ax, const double* __restrict__ ay, const double* __restrict__ az)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if( t >= n ) return;
  const unsigned a = atoms[t];
  vx[a] += ax[a] * dth; vy[a] += ay[a] * dth; vz[a] += az[a] * dth;
}

struct MassTab { double inv_mass[16]; };

__global__ void force_to_accel_kernel(unsigned n, const unsigned* __restrict__ atoms, MassTab M, const unsigned char* __restrict__ type,
                                      double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if( t >= n ) return;
  const unsigned a = atoms[t];
  const double im = M.inv_mass[type[a] & 15];
  fx[a] *= im; fy[a] *= im; fz[a] *= im;
}

__global__ void backup_r_kernel(unsigned n, const unsigned* __restrict__ atoms, const double* __restrict__ rx, const double* __restrict__ ry,
                                const double* __restrict__ rz, double* __restrict__ bx, double* __restrict__ by, double* __restrict__ bz)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if( t >= n ) return;
  const unsigned a = atoms[t];
  bx[t] = rx[a]; by[t] = ry[a]; bz[t] = rz[a];
}

// max over own particles of |xform*(r - r_backup)|^2, one atomicMax per block (values are non-negative doubles, so
// their bit patterns order like unsigned integers)
__global__ void displ_kernel(unsigned n, const unsigned* __restrict__ atoms, XFormInv X /* forward xform in m */, const double* __restrict__ rx,
                             const double* __restrict__ ry, const double* __restrict__ rz, const double* __restrict__ bx, const double* __restrict__ by,
                             const double* __restrict__ bz, unsigned long long* __restrict__ out)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  double d2 = 0.0;
  if( t < n )
  {
    const unsigned a = atoms[t];
    double dx = rx[a] - bx[t], dy = ry[a] - by[t], dz = rz[a] - bz[t];
    if( !X.identity )
    {
      const double x = X.m[0]*dx + X.m[1]*dy + X.m[2]*dz, y = X.m[3]*dx + X.m[4]*dy + X.m[5]*dz, z = X.m[6]*dx + X.m[7]*dy + X.m[8]*dz;
      dx = x; dy = y; dz = z;
    }
    d2 = dx*dx + dy*dy + dz*dz;
  }
  for(int o = 16; o > 0; o >>= 1) d2 = fmax(d2, __shfl_xor_sync(0xffffffffu, d2, o));
  __shared__ double s[8];
  if( (threadIdx.x & 31) == 0 ) s[threadIdx.x >> 5] = d2;
  __syncthreads();
  if( threadIdx.x == 0 )
  {
    double m = s[0];
    for(unsigned w = 1; w < (blockDim.x >> 5); w++) m = fmax(m, s[w]);
    atomicMax(out, (unsigned long long)__double_as_longlong(m));
  }
}

static void invert3(const double* m, double* inv)
{
  const double det = m[0]*(m[4]*m[8]-m[5]*m[7]) - m[1]*(m[3]*m[8]-m[5]*m[6]) + m[2]*(m[3]*m[7]-m[4]*m[6]);
  const double id = 1.0 / det;
  inv[0] =  (m[4]*m[8]-m[5]*m[7])*id; inv[1] = -(m[1]*m[8]-m[2]*m[7])*id; inv[2] =  (m[1]*m[5]-m[2]*m[4])*id;
  inv[3] = -(m[3]*m[8]-m[5]*m[6])*id; inv[4] =  (m[0]*m[8]-m[2]*m[6])*id; inv[5] = -(m[0]*m[5]-m[2]*m[3])*id;
  inv[6] =  (m[3]*m[7]-m[4]*m[6])*id; inv[7] = -(m[0]*m[7]-m[1]*m[6])*id; inv[8] =  (m[0]*m[4]-m[1]*m[3])*id;
}

} // namespace xsb

using namespace xsb;

// sort `n` particles given by device arrays into own cells and install them as the context's particle set
// (ghost cells empty).  in_* are device pointers; they may alias nothing owned by ctx->f64[].
int xsb_internal_assign_device(xsb_ctx* ctx, unsigned n, double* rx, double* ry, double* rz, const double* vx, const double* vy, const double* vz,
                         const unsigned char* type, const unsigned long long* id, const int wrap[3], const double box[3])
{
  const unsigned nc = unsigned(ctx->ncells);
  const xsb_grid_desc& g = ctx->grid;
  BinParams B; B.cell = g.cell_size; B.inv_cell = 1.0 / g.cell_size; B.gl = g.ghost_layers;
  B.ox = g.origin[0] + g.ghost_layers * g.cell_size; B.oy = g.origin[1] + g.ghost_layers * g.cell_size; B.oz = g.origin[2] + g.ghost_layers * g.cell_size;
  B.nx = g.dims[0]; B.ny = g.dims[1]; B.nz = g.dims[2];
  for(int a = 0; a < 3; a++) { B.wrap[a] = wrap ? wrap[a] : 0; B.box[a] = box ? box[a] : 0.0; }
  B.sub_bits = ctx->subcell_bits;
  while( B.sub_bits > 0 && (uint64_t(nc) << (3 * B.sub_bits)) >= (1ull << 32) ) --B.sub_bits;      // the sort key is 32 bits
  XSB_CUDA(ctx, ctx->tmp32a.reserve(n + 16, XSB_GROW)); XSB_CUDA(ctx, ctx->tmp32b.reserve(n + 16, XSB_GROW));
  XSB_CUDA(ctx, ctx->tmp32c.reserve(n + 16, XSB_GROW)); XSB_CUDA(ctx, ctx->tmp32d.reserve(std::max(n, nc) + 16, XSB_GROW));
  unsigned *key = ctx->tmp32a.p, *val = ctx->tmp32b.p, *key2 = ctx->tmp32c.p, *perm = ctx->tmp32d.p;
  XSB_CUDA(ctx, ctx->scratch64.reserve(size_t(nc) + 4));
  unsigned* counts = reinterpret_cast<unsigned*>(ctx->scratch64.p);
  XSB_CUDA(ctx, cudaMemsetAsync(counts, 0, (size_t(nc) + 2) * sizeof(unsigned), ctx->stream));
  if( n )
  {
    bin_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, B, rx, ry, rz, key, val, counts);
    XSB_LAUNCH_CHECK(ctx);
    int end_bit = 1; while( (1ull << end_bit) < nc ) ++end_bit;
    end_bit += 3 * B.sub_bits;
    size_t tmp = 0;
    XSB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp, key, key2, val, perm, int(n), 0, end_bit, ctx->stream));
    XSB_CUDA(ctx, ctx->scratch.reserve(tmp + 16));
    XSB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->scratch.p, tmp, key, key2, val, perm, int(n), 0, end_bit, ctx->stream));
    ctx->launches += 4;
  }
  std::vector<unsigned> hc(size_t(nc) + 2);
  XSB_CUDA(ctx, cudaMemcpyAsync(hc.data(), counts, (size_t(nc) + 2) * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->otb_clamped = hc[nc] + hc[nc + 1];
  if( hc[nc + 1] ) return ctx->fail(XSB_ERR_INVALID, "%u particle(s) lie more than one cell outside the own cells of this grid (lost particles: the list build assumes atoms sit in their cells)", hc[nc + 1]);
  std::vector<uint64_t> off(size_t(nc) + 1, 0);
  for(unsigned c = 0; c < nc; c++) off[c + 1] = off[c] + hc[c];
  if( off[nc] != n ) return ctx->fail(XSB_ERR_STATE, "assign: binned %llu of %u particles", (unsigned long long)off[nc], n);
  // perm lives in tmp32d: install_cells does not touch it
  int rc = xsb_internal_install_cells(ctx, off.data()); if( rc ) return rc;
  const unsigned grid = (n + 255) / 256;
  for(int f = 0; f < XSB_F_TYPE; f++)
  {
    if( f == XSB_F_VIRIAL && !ctx->virial_allocated ) continue;
    const size_t w = f == XSB_F_VIRIAL ? 9 : 1;
    XSB_CUDA(ctx, ctx->f64[f].reserve(w * (size_t(n) + 16), XSB_GROW));
    XSB_CUDA(ctx, cudaMemsetAsync(ctx->f64[f].p, 0, w * (size_t(n) + 1) * sizeof(double), ctx->stream));
  }
  XSB_CUDA(ctx, ctx->type.reserve(size_t(n) + 16, XSB_GROW)); XSB_CUDA(ctx, cudaMemsetAsync(ctx->type.p, 0, size_t(n) + 16, ctx->stream));
  XSB_CUDA(ctx, ctx->id.reserve(size_t(n) + 1, XSB_GROW));
  if( n )
  {
    const double* src[6] = { rx, ry, rz, vx, vy, vz }; const int dstf[6] = { XSB_F_RX, XSB_F_RY, XSB_F_RZ, XSB_F_VX, XSB_F_VY, XSB_F_VZ };
    for(int k = 0; k < 6; k++) if( src[k] ) { permute_kernel<double><<<grid, 256, 0, ctx->stream>>>(n, perm, src[k], ctx->f64[dstf[k]].p); XSB_LAUNCH_CHECK(ctx); }
    if( type ) { permute_kernel<unsigned char><<<grid, 256, 0, ctx->stream>>>(n, perm, type, ctx->type.p); XSB_LAUNCH_CHECK(ctx); }
    if( id ) { permute_kernel<unsigned long long><<<grid, 256, 0, ctx->stream>>>(n, perm, id, reinterpret_cast<unsigned long long*>(ctx->id.p)); XSB_LAUNCH_CHECK(ctx); }
  }
  return XSB_OK;
}

extern "C" {

int xsb_particles_assign(xsb_ctx* ctx, uint64_t n, const double* rx, const double* ry, const double* rz,
                         const double* vx, const double* vy, const double* vz, const uint8_t* type, const uint64_t* id)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, ctx->grid_set, XSB_ERR_STATE, "xsb_grid_set must be called first");
  XSB_REQUIRE(ctx, n < 0xFFFFFFF0ull, XSB_ERR_OVERFLOW, "more than 2^32 particles per GPU");
  XSB_REQUIRE(ctx, n == 0 || (rx && ry && rz), XSB_ERR_INVALID, "null positions");
  XSB_CUDA(ctx, cudaSetDevice(ctx->device));
  // stage host arrays on the device (8-byte words: 6 real fields + id, then type bytes)
  DevBuf<double>& st = ctx->move_stage; DevBuf<unsigned char>& stt = ctx->move_stage8;
  cudaError_t e = st.reserve(7 * (n + 1), XSB_GROW); if( e != cudaSuccess ) return ctx->fail(XSB_ERR_CUDA, "assign staging: %s", cudaGetErrorString(e));
  e = stt.reserve(n + 16, XSB_GROW); if( e != cudaSuccess ) return ctx->fail(XSB_ERR_CUDA, "assign staging: %s", cudaGetErrorString(e));
  const double* hsrc[6] = { rx, ry, rz, vx, vy, vz }; double* d[7];
  for(int k = 0; k < 7; k++) d[k] = st.p + size_t(k) * (n + 1);
  int rc = XSB_OK;
  for(int k = 0; k < 6 && n; k++)
    if( hsrc[k] ) { e = cudaMemcpyAsync(d[k], hsrc[k], n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream); if( e != cudaSuccess ) rc = ctx->fail(XSB_ERR_CUDA, "assign H2D: %s", cudaGetErrorString(e)); }
  if( n && id ) { e = cudaMemcpyAsync(d[6], id, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream); if( e != cudaSuccess ) rc = ctx->fail(XSB_ERR_CUDA, "assign H2D: %s", cudaGetErrorString(e)); }
  if( n && !id ) { iota64_kernel<<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(unsigned(n), reinterpret_cast<unsigned long long*>(d[6])); ctx->launches++; }
  if( n && type ) { e = cudaMemcpyAsync(stt.p, type, n, cudaMemcpyHostToDevice, ctx->stream); if( e != cudaSuccess ) rc = ctx->fail(XSB_ERR_CUDA, "assign H2D: %s", cudaGetErrorString(e)); }
  if( rc == XSB_OK )
    rc = xsb_internal_assign_device(ctx, unsigned(n), d[0], d[1], d[2], vx ? d[3] : nullptr, vy ? d[4] : nullptr, vz ? d[5] : nullptr,
                       type ? stt.p : nullptr, reinterpret_cast<unsigned long long*>(d[6]), nullptr, nullptr);
  cudaStreamSynchronize(ctx->stream);   // the caller's host arrays may die after the call
  return rc;
}

// move_particles for one rank spanning a periodic box: wrap positions into the box, re-bin the own particles into
// cells (stable), leaving ghost cells empty -- call xsb_ghost_comm_scheme and xsb_chunk_neighbors_build afterwards.
int xsb_particles_rebin(xsb_ctx* ctx, const xsb_domain_desc* dom)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, dom != nullptr, XSB_ERR_INVALID, "null domain");
  XSB_REQUIRE(ctx, ctx->h_cell_off.size() == ctx->ncells + 1, XSB_ERR_STATE, "no particles");
  const int P = dom->rank_dims[0] * dom->rank_dims[1] * dom->rank_dims[2];
  XSB_REQUIRE(ctx, P == ctx->nranks, XSB_ERR_INVALID, "rank_dims product differs from the communicator size");
  XSB_CUDA(ctx, cudaSetDevice(ctx->device));
  const unsigned n = unsigned(ctx->n_own);
  // staging buffers persist in the context: steady-state rebuilds must not touch cudaMalloc/cudaFree
  DevBuf<double>& st = ctx->move_stage; DevBuf<unsigned char>& stt = ctx->move_stage8;
  cudaError_t e = st.reserve(7 * (size_t(n) + 1), XSB_GROW); if( e != cudaSuccess ) return ctx->fail(XSB_ERR_CUDA, "rebin staging: %s", cudaGetErrorString(e));
  e = stt.reserve(size_t(n) + 16, XSB_GROW); if( e != cudaSuccess ) return ctx->fail(XSB_ERR_CUDA, "rebin staging: %s", cudaGetErrorString(e));
  double* d[7]; for(int k = 0; k < 7; k++) d[k] = st.p + size_t(k) * (n + 1);
  const int srcf[6] = { XSB_F_RX, XSB_F_RY, XSB_F_RZ, XSB_F_VX, XSB_F_VY, XSB_F_VZ };
  const unsigned grid = (n + 255) / 256;
  if( n )
  {
    for(int k = 0; k < 6; k++) permute_kernel<double><<<grid, 256, 0, ctx->stream>>>(n, ctx->own_atoms.p, ctx->f64[srcf[k]].p, d[k]);
    permute_kernel<unsigned long long><<<grid, 256, 0, ctx->stream>>>(n, ctx->own_atoms.p, reinterpret_cast<const unsigned long long*>(ctx->id.p), reinterpret_cast<unsigned long long*>(d[6]));
    permute_kernel<unsigned char><<<grid, 256, 0, ctx->stream>>>(n, ctx->own_atoms.p, ctx->type.p, stt.p);
    ctx->launches += 8;
  }
  ctx->prof_begin(XSB_PROF_MOVE);
  int rc;
  if( P == 1 )
  {
    const int wrap[3] = { dom->periodic[0], dom->periodic[1], dom->periodic[2] };
    rc = xsb_internal_assign_device(ctx, n, d[0], d[1], d[2], d[3], d[4], d[5], stt.p, reinterpret_cast<unsigned long long*>(d[6]), wrap, dom->box);
  }
  else
  {
    // migrate_cell_particles: particles whose (wrapped) position now belongs to another brick travel to that rank
    unsigned n_new = 0; double* en[7]; unsigned char* tn = nullptr;
    rc = xsb_internal_migrate(ctx, dom, n, d, stt.p, &n_new, en, &tn);
    if( rc == XSB_OK )
      rc = xsb_internal_assign_device(ctx, n_new, en[0], en[1], en[2], en[3], en[4], en[5], tn, reinterpret_cast<unsigned long long*>(en[6]), nullptr, nullptr);
  }
  ctx->prof_end(XSB_PROF_MOVE);
  return rc;
}

int xsb_out_of_domain_count(xsb_ctx* ctx, uint64_t* clamped)
{
  if( !ctx || !clamped ) return XSB_ERR_STATE;
  *clamped = ctx->otb_clamped;
  return XSB_OK;
}

int xsb_migration_stats(xsb_ctx* ctx, uint64_t* sent, uint64_t* received)
{
  if( !ctx ) return XSB_ERR_STATE;
  if( sent ) *sent = ctx->migrated_out;
  if( received ) *received = ctx->migrated_in;
  return XSB_OK;
}

} // extern "C"

namespace xsb
{
// The operators around a step boundary of the velocity-Verlet scheme (config_numerical_schemes.msp:23-52), per own atom in
// one pass: force_to_accel, push_f_v(dt/2)  |  push_f_v_r(dt), push_f_v(dt/2), and the displacement maximum of
// particle_displ_over.  Same arithmetic per atom as the separate kernels; 21 doubles of traffic per atom instead of 43
// and one launch instead of five.
__global__ void verlet_boundary_kernel(unsigned n, const unsigned* __restrict__ atoms, MassTab M, const unsigned char* __restrict__ type,
                                       double dt, double dth, double dt2h, XFormInv Xi, XFormInv Xf,
                                       double* __restrict__ rx, double* __restrict__ ry, double* __restrict__ rz,
                                       double* __restrict__ vx, double* __restrict__ vy, double* __restrict__ vz,
                                       double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz,
                                       const double* __restrict__ bx, const double* __restrict__ by, const double* __restrict__ bz,
                                       unsigned long long* __restrict__ out)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  double d2 = 0.0, s2 = 0.0;
  if( t < n )
  {
    const unsigned a = atoms[t];
    const double im = M.inv_mass[type[a] & 15];
    const double ax = fx[a] * im, ay = fy[a] * im, az = fz[a] * im;                    // force_to_accel
    fx[a] = ax; fy[a] = ay; fz[a] = az;
    double ux = vx[a], uy = vy[a], uz = vz[a];
    ux += ax * dth; uy += ay * dth; uz += az * dth;                                    // push_f_v: second half kick of the step
    double dx = ux * dt + ax * dt2h, dy = uy * dt + ay * dt2h, dz = uz * dt + az * dt2h;   // push_f_v_r of the next step
    s2 = dx*dx + dy*dy + dz*dz;                                                            // how far this step moves the atom
    if( !Xi.identity )
    {
      const double x = Xi.m[0]*dx + Xi.m[1]*dy + Xi.m[2]*dz, y = Xi.m[3]*dx + Xi.m[4]*dy + Xi.m[5]*dz, z = Xi.m[6]*dx + Xi.m[7]*dy + Xi.m[8]*dz;
      dx = x; dy = y; dz = z;
    }
    const double px = rx[a] + dx, py = ry[a] + dy, pz = rz[a] + dz;
    rx[a] = px; ry[a] = py; rz[a] = pz;
    ux += ax * dth; uy += ay * dth; uz += az * dth;                                    // push_f_v: first half kick
    vx[a] = ux; vy[a] = uy; vz[a] = uz;
    double ex = px - bx[t], ey = py - by[t], ez = pz - bz[t];                          // particle_displ_over
    if( !Xf.identity )
    {
      const double x = Xf.m[0]*ex + Xf.m[1]*ey + Xf.m[2]*ez, y = Xf.m[3]*ex + Xf.m[4]*ey + Xf.m[5]*ez, z = Xf.m[6]*ex + Xf.m[7]*ey + Xf.m[8]*ez;
      ex = x; ey = y; ez = z;
    }
    d2 = ex*ex + ey*ey + ez*ez;
  }
  for(int o = 16; o > 0; o >>= 1) { d2 = fmax(d2, __shfl_xor_sync(0xffffffffu, d2, o)); s2 = fmax(s2, __shfl_xor_sync(0xffffffffu, s2, o)); }
  __shared__ double s[8], ss[8];
  if( (threadIdx.x & 31) == 0 ) { s[threadIdx.x >> 5] = d2; ss[threadIdx.x >> 5] = s2; }
  __syncthreads();
  if( threadIdx.x == 0 )
  {
    double m = s[0], ms = ss[0];
    for(unsigned w = 1; w < (blockDim.x >> 5); w++) { m = fmax(m, s[w]); ms = fmax(ms, ss[w]); }
    if( m > 0.0 ) atomicMax(out, (unsigned long long)__double_as_longlong(m));
    if( ms > 0.0 ) atomicMax(out + 1, (unsigned long long)__double_as_longlong(ms));
  }
}
}

extern "C" {

// ---- Verlet pieces ---------------------------------------------------------------------------------------------
int xsb_push_f_v_r(xsb_ctx* ctx, double dt)
{
  XSB_ENTER(ctx);
  const unsigned n = unsigned(ctx->n_own); if( !n ) return XSB_OK;
  ctx->pos_epoch++; ctx->foreign_epoch++;      // no displacement accounting here: an inner-skin sub-list must be re-filtered
  XFormInv Xi; Xi.identity = ctx->grid.xform_is_identity; invert3(ctx->grid.xform, Xi.m);
  push_f_v_r_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, ctx->own_atoms.p, dt, 0.5 * dt * dt, Xi,
      ctx->f64[XSB_F_RX].p, ctx->f64[XSB_F_RY].p, ctx->f64[XSB_F_RZ].p, ctx->f64[XSB_F_VX].p, ctx->f64[XSB_F_VY].p, ctx->f64[XSB_F_VZ].p,
      ctx->f64[XSB_F_FX].p, ctx->f64[XSB_F_FY].p, ctx->f64[XSB_F_FZ].p);
  XSB_LAUNCH_CHECK(ctx);
  return XSB_OK;
}

int xsb_push_f_v(xsb_ctx* ctx, double dt)
{
  XSB_ENTER(ctx);
  const unsigned n = unsigned(ctx->n_own); if( !n ) return XSB_OK;
  push_f_v_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, ctx->own_atoms.p, dt, ctx->f64[XSB_F_VX].p, ctx->f64[XSB_F_VY].p, ctx->f64[XSB_F_VZ].p,
      ctx->f64[XSB_F_FX].p, ctx->f64[XSB_F_FY].p, ctx->f64[XSB_F_FZ].p);
  XSB_LAUNCH_CHECK(ctx);
  return XSB_OK;
}

int xsb_force_to_accel(xsb_ctx* ctx, int n_types, const double* mass)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, mass != nullptr && n_types >= 1 && n_types <= 16, XSB_ERR_INVALID, "1..16 species masses expected");
  const unsigned n = unsigned(ctx->n_own); if( !n ) return XSB_OK;
  MassTab M; for(int i = 0; i < 16; i++) M.inv_mass[i] = i < n_types ? 1.0 / mass[i] : 0.0;
  force_to_accel_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, ctx->own_atoms.p, M, ctx->type.p, ctx->f64[XSB_F_FX].p, ctx->f64[XSB_F_FY].p, ctx->f64[XSB_F_FZ].p);
  XSB_LAUNCH_CHECK(ctx);
  return XSB_OK;
}

} // extern "C"

namespace xsb
{
// inner-skin accounting (SubCtl): this step moved no atom (of any rank: s2 is all-reduced) further than sqrt(*s2)
__global__ void sub_accum_kernel(SubCtl* ctl, const unsigned long long* s2) { ctl->acc += sqrt(__longlong_as_double((long long)*s2)); }
__global__ void sub_accum_value_kernel(SubCtl* ctl, double d) { ctl->acc += d; }
}

int xsb_internal_sub_account(xsb_ctx* ctx, double displacement)
{
  if( !ctx->sub_ctl.p ) return XSB_OK;
  xsb::sub_accum_value_kernel<<<1, 1, 0, ctx->stream>>>(ctx->sub_ctl.p, displacement);
  XSB_LAUNCH_CHECK(ctx);
  return XSB_OK;
}

int xsb_internal_sub_account_dev(xsb_ctx* ctx, unsigned long long* s2_dev)
{
  if( !ctx->sub_ctl.p ) return XSB_OK;
  int rc = xsb_internal_allreduce_max(ctx, reinterpret_cast<double*>(s2_dev), 1); if( rc ) return rc;
  xsb::sub_accum_kernel<<<1, 1, 0, ctx->stream>>>(ctx->sub_ctl.p, s2_dev);
  XSB_LAUNCH_CHECK(ctx);
  return XSB_OK;
}

// the fused pass; out[0] = max |r - r_backup|^2, out[1] = max |step displacement|^2 (both zeroed here)
static int verlet_boundary_launch(xsb_ctx* ctx, int n_types, const double* mass, double dt, unsigned long long* out)
{
  XSB_REQUIRE(ctx, mass != nullptr && n_types >= 1 && n_types <= 16, XSB_ERR_INVALID, "1..16 species masses expected");
  const unsigned n = unsigned(ctx->n_own);
  XSB_REQUIRE(ctx, ctx->backup_n == n, XSB_ERR_STATE, "xsb_backup_r must be called after the last rebuild");
  XSB_CUDA(ctx, cudaMemsetAsync(out, 0, 2 * sizeof(unsigned long long), ctx->stream));
  ctx->pos_epoch++;
  if( n )
  {
    MassTab M; for(int i = 0; i < 16; i++) M.inv_mass[i] = i < n_types ? 1.0 / mass[i] : 0.0;
    XFormInv Xi; Xi.identity = ctx->grid.xform_is_identity; invert3(ctx->grid.xform, Xi.m);
    XFormInv Xf; Xf.identity = ctx->grid.xform_is_identity; for(int i = 0; i < 9; i++) Xf.m[i] = ctx->grid.xform[i];
    const double* b = ctx->backup.p;
    verlet_boundary_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, ctx->own_atoms.p, M, ctx->type.p, dt, 0.5 * dt, 0.5 * dt * dt, Xi, Xf,
        ctx->f64[XSB_F_RX].p, ctx->f64[XSB_F_RY].p, ctx->f64[XSB_F_RZ].p, ctx->f64[XSB_F_VX].p, ctx->f64[XSB_F_VY].p, ctx->f64[XSB_F_VZ].p,
        ctx->f64[XSB_F_FX].p, ctx->f64[XSB_F_FY].p, ctx->f64[XSB_F_FZ].p, b, b + n, b + 2 * size_t(n), out);
    XSB_LAUNCH_CHECK(ctx);
  }
  return XSB_OK;
}

int xsb_internal_allreduce_max(xsb_ctx* ctx, double* dev_inout, int count);      // xsb_ghost.cu

int xsb_internal_displ_ring_init(xsb_ctx* ctx)
{
  if( ctx->displ_host ) return XSB_OK;
  XSB_CUDA(ctx, cudaMallocHost((void**)&ctx->displ_host, sizeof(double) * 2 * XSB_DISPL_RING));
  XSB_CUDA(ctx, ctx->displ_dev.reserve(2 * (XSB_DISPL_RING + 2)));      // + the recording slot + the upload slot
  for(int i = 0; i < XSB_DISPL_RING; i++) XSB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->displ_ev[i], cudaEventDisableTiming));
  return XSB_OK;
}

extern "C" {

// Same pass without a host read-back: the two maxima are all-reduced (MAX) over the ranks on the stream and land in a ring
// of pinned host slots; xsb_displ_poll(lag) returns the pair recorded `lag` calls earlier, so a driver that decides on the
// previous step's value (plus the step displacement as a margin) never waits for the GPU and never couples the ranks'
// host threads through a blocking collective.
int xsb_verlet_boundary_async(xsb_ctx* ctx, int n_types, const double* mass, double dt)
{
  XSB_ENTER(ctx);
  int rc = xsb_internal_displ_ring_init(ctx); if( rc ) return rc;
  // while a step is being recorded the result goes to a fixed extra slot; xsb_step_replay copies it into the ring
  const int slot = ctx->capturing ? XSB_DISPL_RING : int(ctx->displ_seq % XSB_DISPL_RING);
  unsigned long long* out = ctx->displ_dev.p + 2 * slot;
  rc = verlet_boundary_launch(ctx, n_types, mass, dt, out); if( rc ) return rc;
  rc = xsb_internal_allreduce_max(ctx, reinterpret_cast<double*>(out), 2); if( rc ) return rc;      // squares are non-negative: MAX on the doubles
  if( ctx->sub_ctl.p ) { xsb::sub_accum_kernel<<<1, 1, 0, ctx->stream>>>(ctx->sub_ctl.p, out + 1); XSB_LAUNCH_CHECK(ctx); }
  if( ctx->capturing ) { ctx->cap_verlet = true; return XSB_OK; }
  XSB_CUDA(ctx, cudaMemcpyAsync(ctx->displ_host + 2 * slot, out, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  XSB_CUDA(ctx, cudaEventRecord(ctx->displ_ev[slot], ctx->stream));
  ctx->displ_seq++;
  return XSB_OK;
}

int xsb_displ_poll(xsb_ctx* ctx, int lag, double* max_displ, double* max_step_displ)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, lag >= 0 && lag < XSB_DISPL_RING - 1, XSB_ERR_INVALID, "lag must be in 0..6");
  XSB_REQUIRE(ctx, ctx->displ_seq > uint64_t(lag), XSB_ERR_STATE, "xsb_displ_poll: no xsb_verlet_boundary_async call that far back");
  const int slot = int((ctx->displ_seq - 1 - uint64_t(lag)) % XSB_DISPL_RING);
  XSB_CUDA(ctx, cudaEventSynchronize(ctx->displ_ev[slot]));
  if( max_displ ) *max_displ = std::sqrt(ctx->displ_host[2 * slot]);
  if( max_step_displ ) *max_step_displ = std::sqrt(ctx->displ_host[2 * slot + 1]);
  return XSB_OK;
}

int xsb_verlet_boundary(xsb_ctx* ctx, int n_types, const double* mass, double dt, double threshold, int* result, double* max_displ)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, result != nullptr, XSB_ERR_INVALID, "null result");
  XSB_CUDA(ctx, ctx->scratch64.reserve(16));
  int rcl = verlet_boundary_launch(ctx, n_types, mass, dt, ctx->scratch64.p); if( rcl ) return rcl;
  if( ctx->sub_ctl.p )
  {
    // blocking variant: only the displacement maximum is all-reduced (on the host, below), not the step displacement
    if( ctx->nranks > 1 ) ctx->foreign_epoch++;
    else { xsb::sub_accum_kernel<<<1, 1, 0, ctx->stream>>>(ctx->sub_ctl.p, ctx->scratch64.p + 1); XSB_LAUNCH_CHECK(ctx); }
  }
  double d2 = 0.0;
  XSB_CUDA(ctx, cudaMemcpyAsync(&d2, ctx->scratch64.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double d = std::sqrt(d2);
  int rc = xsb_comm_allreduce_max(ctx, &d); if( rc ) return rc;
  if( max_displ ) *max_displ = d;
  *result = d > threshold;
  return XSB_OK;
}

int xsb_backup_r(xsb_ctx* ctx)
{
  XSB_ENTER(ctx);
  const unsigned n = unsigned(ctx->n_own);
  XSB_CUDA(ctx, ctx->backup.reserve(3 * (size_t(n) + 1), XSB_GROW));
  ctx->backup_n = n;
  if( !n ) return XSB_OK;
  double* b = ctx->backup.p;
  backup_r_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, ctx->own_atoms.p, ctx->f64[XSB_F_RX].p, ctx->f64[XSB_F_RY].p, ctx->f64[XSB_F_RZ].p, b, b + n, b + 2 * size_t(n));
  XSB_LAUNCH_CHECK(ctx);
  return XSB_OK;
}

int xsb_particle_displ_over(xsb_ctx* ctx, double threshold, int* result, double* max_displ)
{
  XSB_ENTER(ctx);
  XSB_REQUIRE(ctx, result != nullptr, XSB_ERR_INVALID, "null result");
  const unsigned n = unsigned(ctx->n_own);
  XSB_REQUIRE(ctx, ctx->backup_n == n, XSB_ERR_STATE, "xsb_backup_r must be called after the last rebuild");
  XSB_CUDA(ctx, ctx->scratch64.reserve(16));
  XSB_CUDA(ctx, cudaMemsetAsync(ctx->scratch64.p, 0, sizeof(unsigned long long), ctx->stream));
  if( n )
  {
    XFormInv X; X.identity = ctx->grid.xform_is_identity; for(int i = 0; i < 9; i++) X.m[i] = ctx->grid.xform[i];
    const double* b = ctx->backup.p;
    displ_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, ctx->own_atoms.p, X, ctx->f64[XSB_F_RX].p, ctx->f64[XSB_F_RY].p, ctx->f64[XSB_F_RZ].p, b, b + n, b + 2 * size_t(n), ctx->scratch64.p);
    XSB_LAUNCH_CHECK(ctx);
  }
  double d2 = 0.0;
  XSB_CUDA(ctx, cudaMemcpyAsync(&d2, ctx->scratch64.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  XSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double d = std::sqrt(d2);
  int rc = xsb_comm_allreduce_max(ctx, &d); if( rc ) return rc;   // MPI_Allreduce(MAX) across sub-domains
  if( max_displ ) *max_displ = d;
  *result = d > threshold;
  return XSB_OK;
}

} // extern "C"
