// xsb_ncclwin.cu -- peer addresses of an NCCL symmetric window (NCCL >= 2.28 device API, nccl_device.h).
// ncclCommWindowRegister(..., NCCL_WIN_COLL_SYMMETRIC) maps the same-sized buffer of every rank of the node into one flat
// virtual range (CUDA VMM, only these buffers are peer-mapped -- unlike cudaIpcOpenMemHandle + cudaDeviceEnablePeerAccess,
// which exposes every allocation of the device and measurably slows down cudaMalloc and atomics-heavy kernels, see
// DESIGN.md section 5).  The window handle is only dereferenceable on the device: one tiny kernel asks for the address of
// offset 0 in every rank's copy; xsb_ghost.cu then uses plain pointers in its own pack kernels.
#include <cuda_runtime.h>
#ifdef XSB_HAVE_NCCL_DEVICE
#include <nccl.h>
#include <nccl_device.h>

namespace
{
__global__ void win_peer_ptrs_kernel(ncclWindow_t win, int nranks, void** out)
{
  const int q = threadIdx.x;
  if( q < nranks ) out[q] = ncclGetPeerPointer(win, 0, q);
}
}

extern "C" int xsb_internal_nccl_header_version() { return NCCL_VERSION_CODE; }

// out_dev: device array of nranks pointers
extern "C" int xsb_internal_win_peer_ptrs(void* win, int nranks, void** out_dev, cudaStream_t stream)
{
  if( nranks > 64 ) return 1;
  win_peer_ptrs_kernel<<<1, 64, 0, stream>>>(static_cast<ncclWindow_t>(win), nranks, out_dev);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
#else
extern "C" int xsb_internal_nccl_header_version() { return 0; }
extern "C" int xsb_internal_win_peer_ptrs(void*, int, void**, cudaStream_t) { return 1; }
#endif
