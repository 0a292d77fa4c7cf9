// Gradient all-reduce through the NVSwitch (NVLS): the one exchange step of the data-parallel path.
//
// The reference has no distributed code; under stock DistributedDataParallel this step would be ncclAllReduce.  Here the
// flat fp32 gradient buffer of every rank lives in SYMMETRIC memory bound to one multicast object, and the all-reduce is
// a two-shot kernel that never stages data through a peer's SMs:
//   rank r owns the r-th slice:   v = multimem.ld_reduce.add.v4.f32 [mc + i]   (the switch reads the 16 bytes from all
//                                                                               GPUs and returns their sum)
//                                 multimem.st.v4.f32 [mc + i], v               (the switch writes the sum to all GPUs)
// so every GPU sends and receives S/N... bytes per phase instead of the 2(N-1)/N * S of a ring, and nothing is reduced
// on an SM.  The caller brackets the launch with system-scope barriers across the ranks (all local gradients written
// before, all slices stored after); PyTorch's symmetric-memory handle provides both the multicast pointer and the barrier.
//
// Roofline: NVLink.  Algorithmic bytes per launch and rank: 16 B * n_vec / world read-reduced + the same stored.
#include "common.cuh"

namespace unetk {

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float4* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

constexpr int kArUnroll = 4;

__global__ void __launch_bounds__(256) nvls_allreduce_kernel(float4* __restrict__ mc, int64_t begin, int64_t end, float scale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < end; i0 += stride * kArUnroll) {
    float4 v[kArUnroll];
#pragma unroll
    for (int u = 0; u < kArUnroll; ++u)
      if (i0 + u * stride < end) v[u] = multimem_ld_reduce_add(mc + i0 + u * stride);
#pragma unroll
    for (int u = 0; u < kArUnroll; ++u) {
      if (i0 + u * stride < end) {
        v[u].x *= scale; v[u].y *= scale; v[u].z *= scale; v[u].w *= scale;
        multimem_st(mc + i0 + u * stride, v[u]);
      }
    }
  }
}

}  // namespace unetk

using namespace unetk;

extern "C" int unetk_nvls_allreduce_f32(void* multicast_ptr, int64_t n_elems, int32_t rank, int32_t world, float scale,
                                        void* stream) {
  UNETK_REQUIRE(multicast_ptr != nullptr && n_elems > 0 && world >= 1 && rank >= 0 && rank < world, "nvls_allreduce: bad argument");
  UNETK_REQUIRE((reinterpret_cast<uintptr_t>(multicast_ptr) & 15) == 0 && n_elems % 4 == 0,
                "nvls_allreduce: the buffer must be 16-byte aligned and a multiple of 4 floats long");
  const int64_t n_vec = n_elems / 4;
  const int64_t chunk = (n_vec + world - 1) / world;
  const int64_t begin = chunk * rank, end = begin + chunk < n_vec ? begin + chunk : n_vec;
  if (begin >= end) return UNETK_OK;
  int64_t blocks = (end - begin + 256LL * kArUnroll - 1) / (256LL * kArUnroll);
  const int64_t cap = (int64_t)sm_count() * 4;
  if (blocks > cap) blocks = cap;
  nvls_allreduce_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(static_cast<float4*>(multicast_ptr), begin, end, scale);
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}
