// tcgen05.mma issue-rate probe for CTA PAIRS (cta_group::2): cycles per M = 256 MMA (kind::f16, bf16, K = 16) as a function
// of N, next to the single-CTA M = 128 numbers of tools/umma_rate.cu.  One 2-CTA cluster per SM pair, operands resident
// in shared memory (each CTA holds its own 128 rows of A and HALF of the B rows), warp-uniform issue by the leader CTA,
// one multicast commit at the end.  Question answered: does the 48-cycle floor of an N = 64 MMA move when the MMA spans
// two SMs?  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I image_segmentation_b200/csrc -I include tools/umma_rate2.cu -o tools/umma_rate2 -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "unetk.h"
#include "tc_common.cuh"

using namespace unetk::tc;

struct Cfg { int n, iters, accs; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate2_kernel(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const uint32_t rank = blockIdx.x & 1;          // CTA rank inside the pair (cluster x-dim = 2)
  // A: 128 rows x 64 K (16 KiB, SW128 K-major); B: (N/2 <= 128) rows x 64 K (16 KiB) x 2 buffers
  uint32_t* w = (uint32_t*)smem;
  for (int i = threadIdx.x; i < (16 + 32) * 1024 / 4; i += blockDim.x) w[i] = 0x3c003c00u ^ (i * 2654435761u & 0x00ff00ffu);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc_2sm<512>(&tmem_ptr);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tm = tmem_ptr;
  long long cycles = 0;
  if (rank == 0 && threadIdx.x < 32) {
    const uint32_t idesc = make_idesc_bf16(256, c.n, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 16 * 1024);
    uint64_t da[4], db[8];
    for (int k = 0; k < 4; ++k) da[k] = make_smem_desc(a0 + k * 32, 0, 1024);
    for (int k = 0; k < 8; ++k) db[k] = make_smem_desc(b0 + (k >> 2) * 16 * 1024 + (k & 3) * 32, 0, 1024);
    const long long t0 = clock64();
    for (int it = 0; it < c.iters; it += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) umma_bf16_2sm_warp(tm + ((c.accs > 1 && (u & 1)) ? c.n : 0), da[u & 3], db[u], idesc, 1);
    }
    umma_commit_2sm_warp(&bar);
    mbar_wait(&bar, 0);
    cycles = clock64() - t0;
  } else if (rank == 1 && threadIdx.x == 0) {
    mbar_wait(&bar, 0);                           // the multicast commit arrives here too
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (rank == 0 && threadIdx.x == 0) out[blockIdx.x >> 1] = cycles;
  if (threadIdx.x < 32) { tcgen05_fence_after(); tmem_dealloc_2sm<512>(tm); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 74 * sizeof(long long));
  cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 4096;
  printf("cta_group::2, M = 256 over two SMs, warp-uniform issue from the leader CTA, 74 pairs\n");
  printf("%-6s %-5s %12s %16s %22s\n", "N", "accs", "cyc/MMA", "flop/clk/pair", "flop/clk/SM (vs 8192)");
  for (int accs : {1, 2})
    for (int n : {32, 64, 96, 128, 192, 256}) {
      if (accs * n > 512) continue;
      Cfg c{n, iters, accs};
      for (int rep = 0; rep < 2; ++rep) rate2_kernel<<<148, 128, 64 * 1024>>>(c, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N=%d: %s\n", n, cudaGetErrorString(e)); return 1; }
      long long h[74];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double s = 0;
      for (int i = 0; i < 74; ++i) s += h[i];
      const double cyc = s / 74 / iters;
      const double fpair = 2.0 * 256 * n * 16 / cyc;
      printf("%-6d %-5d %12.1f %16.0f %22.0f\n", n, accs, cyc, fpair, fpair / 2);
    }
  return 0;
}
