// umma_probe: establishes, on real hardware, how tcgen05.mma addresses 128B-swizzled shared-memory operands when the
// descriptor's start address is shifted by whole 128-byte rows and when row groups are not 1024 bytes apart.
// (Needed for halo reuse in the 3x3 implicit GEMM: one shared-memory halo tile serving all 9 taps.)
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
// Run  :  tools/umma_probe   (prints, per configuration, which smem row/column each accumulator element came from)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e = (x);                                                                       \
    if (e != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);           \
      exit(1);                                                                                 \
    }                                                                                          \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Cfg {
  uint32_t start_bytes;  // descriptor start offset from the operand base
  uint32_t lbo, sbo;     // bytes
  uint32_t base_offset;  // 3-bit field
  uint32_t a_mn_major;   // 0: K-major A, 1: MN-major A
  uint32_t kadv_bytes;   // start-address advance per UMMA_K step
  uint32_t m;            // 128 or 64
};

constexpr int A_ROWS = 512;  // 64 KiB operand region

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap map_a,
                                                       const __grid_constant__ CUtensorMap map_b, Cfg cfg, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                       // A_ROWS x 128 B
  uint8_t* sb = smem + A_ROWS * 128;        // 64 x 128 B (identity)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 64 * 128);
  uint64_t* mma_bar = bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mma_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_ptr)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tmem_ptr;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(A_ROWS * 128 + 64 * 128));
    for (int i = 0; i < A_ROWS / 256; ++i)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
                       "r"(smem_u32(sa + i * 256 * 128)),
                   "l"(reinterpret_cast<uint64_t>(&map_a)), "r"(smem_u32(bar)), "r"(0), "r"(i * 256)
                   : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
                     "r"(smem_u32(sb)),
                 "l"(reinterpret_cast<uint64_t>(&map_b)), "r"(smem_u32(bar)), "r"(0), "r"(0)
                 : "memory");
    asm volatile(
        "{\n\t.reg .pred P1;\n\tW1:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D1;\n\tbra W1;\n\tD1:\n\t}" ::"r"(
            smem_u32(bar))
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;");
    auto desc = [](uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t bo) {
      uint64_t d = 0;
      d |= (uint64_t)((addr >> 4) & 0x3fff);
      d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
      d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)(bo & 7) << 49;
      d |= (uint64_t)2 << 61;
      return d;
    };
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (cfg.a_mn_major << 15) | (0u << 16) | ((64u >> 3) << 17) |
                           ((cfg.m >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      const uint64_t da = desc(smem_u32(sa) + cfg.start_bytes + k * cfg.kadv_bytes, cfg.lbo, cfg.sbo, cfg.base_offset);
      const uint64_t db = desc(smem_u32(sb) + k * 32, 0, 1024, 0);
      const uint32_t accum = k > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(accum)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mma_bar)) : "memory");
  }
  // everyone waits for the MMA
  asm volatile(
      "{\n\t.reg .pred P1;\n\tW2:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D2;\n\tbra W2;\n\tD2:\n\t}" ::"r"(
          smem_u32(mma_bar))
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c * 32)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, void* ptr, int rows, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", (int)r);
    exit(1);
  }
  return m;
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fn;
  std::vector<__nv_bfloat16> a_row(A_ROWS * 64), a_col(A_ROWS * 64), ident(64 * 64);
  for (int r = 0; r < A_ROWS; ++r)
    for (int c = 0; c < 64; ++c) {
      a_row[r * 64 + c] = __float2bfloat16((float)(r % 256));
      a_col[r * 64 + c] = __float2bfloat16((float)c);
    }
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 64; ++k) ident[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *d_row, *d_col, *d_id;
  float* d_out;
  CK(cudaMalloc(&d_row, a_row.size() * 2));
  CK(cudaMalloc(&d_col, a_col.size() * 2));
  CK(cudaMalloc(&d_id, ident.size() * 2));
  CK(cudaMalloc(&d_out, 128 * 64 * 4));
  CK(cudaMemcpy(d_row, a_row.data(), a_row.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_col, a_col.data(), a_col.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_id, ident.data(), ident.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap m_row = make_map(enc, d_row, A_ROWS, 256), m_col = make_map(enc, d_col, A_ROWS, 256),
              m_id = make_map(enc, d_id, 64, 64);
  const int smem = A_ROWS * 128 + 64 * 128 + 64 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));

  struct Named {
    const char* name;
    Cfg c;
  };
  // K-major: kadv 32 B.  MN-major A: rows = K, kadv = 16 rows * 128 B = 2048, LBO = slab stride (we use 128 rows * 128 B)
  std::vector<Named> cfgs = {
      {"K  start=0     sbo=1024 bo=0", {0, 0, 1024, 0, 0, 32, 128}},
      {"K  start=1row  sbo=1024 bo=0", {128, 0, 1024, 0, 0, 32, 128}},
      {"K  start=1row  sbo=1024 bo=1", {128, 0, 1024, 1, 0, 32, 128}},
      {"K  start=2row  sbo=1024 bo=0", {256, 0, 1024, 0, 0, 32, 128}},
      {"K  start=2row  sbo=1024 bo=2", {256, 0, 1024, 2, 0, 32, 128}},
      {"K  start=0     sbo=2048 bo=0", {0, 0, 2048, 0, 0, 32, 128}},
      {"K  start=1row  sbo=2048 bo=0", {128, 0, 2048, 0, 0, 32, 128}},
      {"K  start=1row  sbo=2048 bo=1", {128, 0, 2048, 1, 0, 32, 128}},
      {"K  start=17row sbo=2048 bo=0", {17 * 128, 0, 2048, 0, 0, 32, 128}},
      {"K  start=17row sbo=2048 bo=1", {17 * 128, 0, 2048, 1, 0, 32, 128}},
      {"K  start=0     sbo=1280 bo=0", {0, 0, 1280, 0, 0, 32, 128}},
      {"K  start=1row  sbo=1280 bo=0", {128, 0, 1280, 0, 0, 32, 128}},
      {"MN start=0     sbo=1024 lbo=16384 bo=0", {0, 16384, 1024, 0, 1, 2048, 128}},
      {"MN start=1row  sbo=1024 lbo=16384 bo=0", {128, 16384, 1024, 0, 1, 2048, 128}},
      {"MN start=1row  sbo=1024 lbo=16384 bo=1", {128, 16384, 1024, 1, 1, 2048, 128}},
      {"MN start=2row  sbo=1024 lbo=16384 bo=0", {256, 16384, 1024, 0, 1, 2048, 128}},
      {"MN start=0     sbo=1024 lbo=0     bo=0", {0, 0, 1024, 0, 1, 2048, 128}},
      {"K  M=64 start=0 sbo=1024 (TMEM lane layout of a 64-row accumulator)", {0, 0, 1024, 0, 0, 32, 64}},
      {"MN M=64 start=0 sbo=1024 lbo=16384", {0, 16384, 1024, 0, 1, 2048, 64}},
  };
  std::vector<float> o_row(128 * 64), o_col(128 * 64);
  for (auto& nc : cfgs) {
    probe_kernel<<<1, 128, smem>>>(m_row, m_id, nc.c, d_out);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(o_row.data(), d_out, o_row.size() * 4, cudaMemcpyDeviceToHost));
    probe_kernel<<<1, 128, smem>>>(m_col, m_id, nc.c, d_out);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(o_col.data(), d_out, o_col.size() * 4, cudaMemcpyDeviceToHost));
    printf("== %s\n", nc.name);
    if (!nc.c.a_mn_major) {
      // D[m][n] = A[row(m)][col(n)]: row-probe gives row(m) (must be equal across n), col-probe gives col(n)
      printf("   src row of m=0..17,64,127 : ");
      int ms[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 15, 16, 17, 64, 127};
      for (int m : ms) {
        bool uniform = true;
        for (int n = 1; n < 64; ++n) uniform &= o_row[m * 64 + n] == o_row[m * 64];
        printf("%d%s ", (int)o_row[m * 64], uniform ? "" : "*");
      }
      if (nc.c.m == 64) {
        printf("\n   TMEM lane -> source row (all 128 lanes, -1 = not written / other): ");
        for (int m = 0; m < 128; ++m) printf("%d ", (int)o_row[m * 64]);
      }
      printf("\n   src col of n=0..63 at m=0 : ");
      for (int n = 0; n < 64; n += 1) printf("%d ", (int)o_col[0 * 64 + n]);
      printf("\n   src col of n=0..63 at m=9 : ");
      for (int n = 0; n < 64; n += 1) printf("%d ", (int)o_col[9 * 64 + n]);
      printf("\n");
    } else {
      // D[m][n] = A[k=n -> smem row][col m]: row-probe gives the smem row used for k=n, col-probe the channel for m
      printf("   smem row used for k=n (n=0..20,63) at m=0: ");
      for (int n = 0; n < 21; ++n) printf("%d ", (int)o_row[0 * 64 + n]);
      printf("%d\n", (int)o_row[63]);
      printf("   channel of m=0..9,63,64,65,127 at n=0: ");
      int ms[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 63, 64, 65, 127};
      for (int m : ms) printf("%d ", (int)o_col[m * 64 + 0]);
      printf("\n   channel of m=0..9 at n=9: ");
      for (int m = 0; m < 10; ++m) printf("%d ", (int)o_col[m * 64 + 9]);
      printf("\n   row-probe at m=64,n=0..3 (slab 2 rows): %d %d %d %d\n", (int)o_row[64 * 64], (int)o_row[64 * 64 + 1],
             (int)o_row[64 * 64 + 2], (int)o_row[64 * 64 + 3]);
      if (nc.c.m == 64) {
        printf("   TMEM lane -> channel (col-probe, n=0), all 128 lanes: ");
        for (int m = 0; m < 128; ++m) printf("%d ", (int)o_col[m * 64]);
        printf("\n");
      }
    }
  }
  return 0;
}
