// tools/microbench.cu -- B200 primitive rates behind the design decisions of DESIGN.md section 3.1 (Newton-off vs half list):
// cycles per warp instruction, per SM, of
//   (a) LDS.64 / LDS.128 gathers with random per-lane addresses (what a pair pass does per listed pair),
//   (b) shared-memory FP64 atomicAdd with random addresses (what an in-stage half list would do per pair and component),
//   (c) red.global.add.f64 with random addresses (what a half list scattering f_j to global memory would do).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NT = 1024, ITERS = 2048, SLOTS = 4096;      // 32 KiB of doubles per CTA

__device__ __forceinline__ unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template<int MODE>
__global__ void __launch_bounds__(NT) bench_kernel(double* out, double* gbuf, unsigned gmask, long long* cycles)
{
  __shared__ __align__(16) double sm[SLOTS];
  for(int i = threadIdx.x; i < SLOTS; i += NT) sm[i] = double(i);
  __syncthreads();
  unsigned s = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 1u;
  double acc = 0.0, acc2 = 0.0;
  const long long t0 = clock64();
# pragma unroll 4
  for(int it = 0; it < ITERS; it++)
  {
    const unsigned r = lcg(s);
    if( MODE == 0 ) acc += sm[r & (SLOTS - 1)];                                                        // LDS.64 gather
    if( MODE == 1 ) { const double2 v = reinterpret_cast<const double2*>(sm)[r & (SLOTS / 2 - 1)]; acc += v.x; acc2 += v.y; }   // LDS.128 gather
    if( MODE == 2 ) atomicAdd(&sm[r & (SLOTS - 1)], 1.0);                                              // ATOMS CAS loop
    if( MODE == 3 ) asm volatile("red.global.add.f64 [%0], %1;" :: "l"(gbuf + (r & gmask)), "d"(1.0) : "memory");
    if( MODE == 4 ) acc += sm[(threadIdx.x + it) & (SLOTS - 1)];                                       // LDS.64, conflict-free
  }
  const long long t1 = clock64();
  if( acc + acc2 == -1.0 ) out[0] = acc;
  if( threadIdx.x == 0 ) cycles[blockIdx.x] = t1 - t0;
  if( MODE == 2 && threadIdx.x == 0 ) out[blockIdx.x + 1] = sm[5];
}

template<int MODE> static void run(const char* what, double* out, double* gbuf, unsigned gmask, long long* cyc, int nsm)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench_kernel<MODE><<<nsm, NT>>>(out, gbuf, gmask, cyc);
  cudaEventRecord(e0);
  bench_kernel<MODE><<<nsm, NT>>>(out, gbuf, gmask, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
  double mean = 0; for(int i = 0; i < nsm; i++) mean += double(h[i]); mean /= nsm;
  const double warp_instr = double(NT / 32) * ITERS;      // per SM
  printf("%-44s %8.3f ms  %7.2f cycles per warp instruction per SM  (%.1f G lane-ops/s chip-wide)\n", what, ms, mean / warp_instr,
         double(nsm) * NT * ITERS / (ms * 1e-3) * 1e-9);
  cudaError_t e = cudaGetLastError(); if( e != cudaSuccess ) printf("  CUDA error: %s\n", cudaGetErrorString(e));
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  printf("%s, %d SMs, one 1024-thread CTA per SM, %d iterations per thread\n", p.name, nsm, ITERS);
  double *out, *gbuf; long long* cyc;
  const size_t gn = size_t(1) << 23;      // 64 MiB of doubles: L2-resident scatter target, like the forces of one brick
  cudaMalloc(&out, 4096 * sizeof(double)); cudaMalloc(&gbuf, gn * sizeof(double)); cudaMalloc(&cyc, 256 * sizeof(long long));
  cudaMemset(gbuf, 0, gn * sizeof(double));
  run<4>("LDS.64, consecutive addresses", out, gbuf, unsigned(gn - 1), cyc, nsm);
  run<0>("LDS.64 gather, random addresses", out, gbuf, unsigned(gn - 1), cyc, nsm);
  run<1>("LDS.128 gather, random addresses", out, gbuf, unsigned(gn - 1), cyc, nsm);
  run<2>("atomicAdd(double) shared, random addresses", out, gbuf, unsigned(gn - 1), cyc, nsm);
  run<3>("red.global.add.f64, random in 64 MiB", out, gbuf, unsigned(gn - 1), cyc, nsm);
  run<3>("red.global.add.f64, random in 1 MiB", out, gbuf, unsigned((1u << 17) - 1), cyc, nsm);
  return 0;
}
