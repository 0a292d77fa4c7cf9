// tcgen05.mma issue-rate probe: cycles per MMA (kind::f16, bf16, K = 16) as a function of the tile shape and of where
// the A operand lives (shared memory descriptor vs tensor memory).  One CTA per SM, operands resident in shared
// memory (no TMA in the loop), one commit at the end.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I image_segmentation_b200/csrc -I include tools/umma_rate.cu -o tools/umma_rate -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "unetk.h"
#include "tc_common.cuh"

using namespace unetk::tc;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

// whole-warp issue: every lane executes the call with identical (warp-uniform) operands, one elected lane issues
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

struct Cfg { int m, n, a_tmem, iters, alt, issuers, elect; };  // alt: number of accumulators cycled through, one per MMA

__global__ void __launch_bounds__(128, 1) rate_kernel(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  // A: 128 rows x 64 K (16 KiB, SW128); B: 256 rows x 64 K (32 KiB) x 2 buffers
  uint32_t* w = (uint32_t*)smem;
  for (int i = threadIdx.x; i < (16 + 64) * 1024 / 4; i += blockDim.x) w[i] = 0x3c003c00u ^ (i * 2654435761u & 0x00ff00ffu);
  if (threadIdx.x == 0) { mbar_init(&bar, c.issuers); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_ptr);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = tmem_ptr;
  __shared__ long long tstart[4], tend[4];
  if (c.elect && threadIdx.x < 32) {
    const uint32_t idesc = make_idesc_bf16(c.m, c.n, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 16 * 1024);
    uint64_t da[4], db[8];
    for (int k = 0; k < 4; ++k) da[k] = make_smem_desc(a0 + k * 32, 0, 1024);
    for (int k = 0; k < 8; ++k) db[k] = make_smem_desc(b0 + (k >> 2) * 32 * 1024 + (k & 3) * 32, 0, 1024);
    long long t0 = clock64();
    for (int it = 0; it < c.iters; it += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) umma_bf16_elect(tm, da[u & 3], db[u], idesc, 1);
    }
    if (threadIdx.x == 0) umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) { tstart[0] = t0; tend[0] = t1; }
  } else if (!c.elect && (threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < c.issuers) {
    const int wi = threadIdx.x >> 5;
    const uint32_t idesc = make_idesc_bf16(c.m, c.n, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 16 * 1024);
    // descriptors precomputed; the loop body is 8 back-to-back MMAs (K slices 0..3 of B buffer 0, then of buffer 1)
    uint64_t da[4], db[8];
    for (int k = 0; k < 4; ++k) da[k] = make_smem_desc(a0 + k * 32, 0, 1024);
    for (int k = 0; k < 8; ++k) db[k] = make_smem_desc(b0 + (k >> 2) * 32 * 1024 + (k & 3) * 32, 0, 1024);
    const uint32_t d0 = tm + wi * 2 * c.n;
    const uint32_t d1 = c.alt > 1 ? d0 + c.n : d0;
    long long t0 = clock64();
    for (int it = 0; it < c.iters; it += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t d = (u & 1) ? d1 : d0;
        if (c.a_tmem) umma_bf16_ts(d, tm + (u & 3) * 8, db[u], idesc, 1);
        else umma_bf16(d, da[u & 3], db[u], idesc, 1);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    tstart[wi] = t0;
    tend[wi] = t1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long a = tstart[0], b = tend[0];
    for (int i = 1; i < c.issuers; ++i) { a = min(a, tstart[i]); b = max(b, tend[i]); }
    out[blockIdx.x] = b - a;
  }
  if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 4096;
  printf("%-6s %-6s %-7s %-5s %-5s %12s %10s\n", "M", "N", "A", "accs", "thr", "cyc/MMA", "flop/clk");
  for (int elect : {0, 1})
  for (int issuers : {1, 2})
  for (int alt : {1, 2})
  for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
    for (int m : {128, 64})
      for (int n : {32, 64, 96, 128, 192, 256}) {
        if (issuers * 2 * n > 512 || (a_tmem && (alt > 1 || issuers > 1)) || (issuers > 1 && alt > 1)) continue;
        if (elect && (issuers > 1 || alt > 1 || a_tmem)) continue;
        Cfg c{m, n, a_tmem, iters, alt, issuers, elect};
        for (int rep = 0; rep < 2; ++rep) rate_kernel<<<148, 128, 100 * 1024>>>(c, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("M=%d N=%d a_tmem=%d: %s\n", m, n, a_tmem, cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double s = 0;
        for (int i = 0; i < 148; ++i) s += h[i];
        const double cyc = s / 148 / (iters * issuers);
        printf("%-6d %-6d %-7s %-5d %-5s %12.1f %10.0f\n", m, n, a_tmem ? "tmem" : "smem", alt, elect ? "elect" : (issuers == 2 ? "2" : "1"), cyc, 2.0 * m * n * 16 / cyc);
      }
  return 0;
}
