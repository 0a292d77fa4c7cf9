// bf16 implicit-GEMM contractions on the 5th-generation tensor cores (sm_100a):
//   TMA (cp.async.bulk.tensor, 128B swizzle, OOB zero fill = conv padding) -> shared-memory ring
//   -> tcgen05.mma (fp32 accumulators in TMEM, double buffered; issued warp-uniformly, see umma_bf16_warp in tc_common.cuh)
//   -> tcgen05.ld epilogue (bias, bf16 rounding, BatchNorm batch statistics or BatchNorm-backward sums, TMA stores:
//      NHWC / convT pixel shuffle into the concat slice).
//
// Persistent, warp-specialised CTAs (one per SM, usually paired into 2-CTA clusters for cta_group::2 MMAs):
// warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-11 = two epilogue groups
// (warp w of a group owns TMEM lanes 32*(w%4)..+31; the weight-gradient kernels have one group).
//
// Reference call sites replaced: unet/unet.py:16,19 (Conv2d 3x3 p1 forward, and its data gradient with the
// flipped/transposed weight pack), unet/unet.py:59 (ConvTranspose2d k2 s2 forward / data gradient), and the
// filter-gradient half of convolution_backward for both (tc_wgrad).
//
// Roofline: tensor pipe.  Algorithmic FLOPs per launch = 2 * rows * K * N_total.
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "conv_internal.cuh"
#include "tc_common.cuh"

namespace unetk {
namespace tc {

// =================================================================================================
// host helpers
// =================================================================================================
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_act_map(CUtensorMap* map, const unetk_tensor& t, int pw, int ph, int nb, int step, int oh, int ow) {
  EncodeTiledFn enc = get_encode_fn();
  UNETK_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  const uint64_t es = 2;
  char* base = static_cast<char*>(t.ptr) + ((int64_t)oh * t.w + ow) * t.ld * es;
  cuuint64_t dims[4] = {(cuuint64_t)t.c, (cuuint64_t)(t.w / step), (cuuint64_t)(t.h / step), (cuuint64_t)t.n};
  cuuint64_t strides[3] = {(cuuint64_t)step * t.ld * es, (cuuint64_t)step * t.w * t.ld * es,
                           (cuuint64_t)t.h * t.w * t.ld * es};
  cuuint32_t box[4] = {64, (cuuint32_t)pw, (cuuint32_t)ph, (cuuint32_t)nb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(activation) failed with CUresult %d (dims %d,%d,%d,%d ld %d box %d,%d,%d)", (int)r,
              t.c, t.w / step, t.h / step, t.n, t.ld, pw, ph, nb);
    return UNETK_ERR_CUDA;
  }
  return UNETK_OK;
}

int make_mat_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t k, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  UNETK_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)k * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(matrix) failed with CUresult %d (rows %lld k %lld box %d)", (int)r, (long long)rows,
              (long long)k, box_rows);
    return UNETK_ERR_CUDA;
  }
  return UNETK_OK;
}

static int pow2_floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}

PixelTile choose_pixel_tile(int n, int h, int w) {
  PixelTile t;
  t.pw = pow2_floor(w) < 16 ? pow2_floor(w) : 16;
  const int max_ph = 128 / t.pw;
  t.ph = pow2_floor(h) < max_ph ? pow2_floor(h) : max_ph;
  t.nb = 128 / (t.pw * t.ph);
  t.tiles_w = (w + t.pw - 1) / t.pw;
  t.tiles_h = (h + t.ph - 1) / t.ph;
  t.tiles_n = (n + t.nb - 1) / t.nb;
  return t;
}

// =================================================================================================
// forward-family kernels: y[rows, N] = A[rows, K] * B[N, K]^T, both operands K-major in shared memory
//   tc_conv_kernel       : one TMA box per (tap, 64-channel chunk) -- every mode; tc_conv2_kernel = its CTA-pair form
//   tc_conv_halo2_kernel : the default for 3x3: halo reuse (below) + CTA pairs, every N tile
//   tc_conv_halo_kernel  : single-CTA halo kernel (one pixel tile, or UNETK_TC_NO_HALO_PAIR); 3x3 only; ONE (18 x 10 pixel) halo box per 64-channel chunk feeds all 9 taps through
//                          row-shifted UMMA descriptors (start + (r*10+s)*128 B, SBO = 10*128 B).  tools/umma_probe.cu
//                          established on B200 that 128B swizzling is a function of the absolute shared-memory address,
//                          so shifted starts and a 1280-byte group stride address the TMA-written tile consistently.
//                          Cuts L2->SMEM operand traffic of the A side 6.4x (23 KB instead of 9 x 16 KB per chunk).
// =================================================================================================
// per-CTA cycle counters of the halo kernel (role idle times), read back with unetk_debug_counters(); 8 slots per CTA:
// [0] producer wait(empty A) [1] mma wait(full A) [2] mma wait(full B) [3] mma wait(tmem empty) [4] mma total
// [5] epilogue g0 wait(tmem full) [6] epilogue g0 inside epilogue_tile [7] epilogue g0 total
// (compiled in only with -DUNETK_DEBUG_COUNTERS; the default build carries neither the array nor clock reads)
#ifdef UNETK_DEBUG_COUNTERS
__device__ long long g_dbg[160 * 8];
constexpr bool kDbg = true;
#define DBG_T0() const long long _t0 = clock64()
#define DBG_ADD(var) var += clock64() - _t0
#define DBG_NOW() clock64()
#else
constexpr bool kDbg = false;
__device__ long long* const g_dbg = nullptr;   // never dereferenced: every use is behind `if (kDbg ...)`
#define DBG_T0() do { } while (0)
#define DBG_ADD(var) do { } while (0)
#define DBG_NOW() 0LL
#endif

constexpr int kTileM = 128;
constexpr int kBlockK = 64;                    // bf16 elements per K step = one 128-byte swizzle row
constexpr int kABytes = kTileM * kBlockK * 2;  // 16 KiB
constexpr int kNumThreads = 256;       // wgrad kernels: 4 control warps + 4 epilogue warps
constexpr int kConvThreads = 384;      // conv kernels: 4 control warps + 2 epilogue groups of 4 warps
constexpr int kStagingBytes = kTileM * 128;    // one 128-row x 64-channel bf16 output block
constexpr int kHaloRows = 18 * 10;             // (16+2) x (8+2) pixels
constexpr int kHaloBytes = kHaloRows * 128;    // 23040
constexpr int kHaloStride = 23 * 1024;         // halo buffers on 1 KiB boundaries
constexpr int kMaxBSlots = 24;
constexpr int kMaxAStages = 8;

struct ConvParams {
  CUtensorMap map_a[4];
  CUtensorMap map_b;
  CUtensorMap map_y[4];
  int mode, taps, chunks_per_tap;
  int pw, ph, nb, tiles_w, tiles_h;
  int num_m_tiles, num_n_tiles;
  int n, h, w;          // grid of the GEMM rows
  int cout;             // channels of y (mode 2: N_total = 4*cout)
  const float* bias;
  double* stat_sum;
  double* stat_sumsq;
  // shared-memory plan (bytes from the 1 KiB aligned base)
  int stages;           // per-tap kernel: ring depth; halo kernel: halo ring depth
  int b_slots;          // halo kernel: B ring depth (== K steps when resident)
  int resident_b;       // halo kernel: B loaded once and kept for every tile
  int num_staging;      // 1 or 2 epilogue staging blocks
  uint32_t off_b, off_staging, off_stats, off_bars;
  int stat_channels;    // size of each statistics array in shared memory (0 = none)
  // fused BatchNorm-backward reduction (data-gradient launches): y is dA of a conv+BN+ReLU layer whose raw output is z
  CUtensorMap map_z;    // same tiling as map_y[0]
  uint32_t off_zbuf;    // two 16 KiB blocks (one per epilogue group); 0 when unused
  const float* bn_scale;
  const float* bn_shift;
  const float* bn_mean;
  const float* bn_invstd;
  double* bn_sums;      // [2][cout]
};

// barrier block: [fullA 8][emptyA 8][fullB 24][emptyB 24][tmem_full 2][tmem_empty 2] + tmem pointer
constexpr int kBarBytes = (2 * kMaxAStages + 2 * kMaxBSlots + 6) * 8 + 16;

struct Bars {
  uint64_t* full_a;
  uint64_t* empty_a;
  uint64_t* full_b;
  uint64_t* empty_b;
  uint64_t* tmem_full;
  uint64_t* tmem_empty;
  uint64_t* zfull;   // z tile of the fused BatchNorm-backward reduction, one per epilogue group
  uint32_t* tmem_ptr;
  __device__ explicit Bars(uint8_t* base) {
    full_a = reinterpret_cast<uint64_t*>(base);
    empty_a = full_a + kMaxAStages;
    full_b = empty_a + kMaxAStages;
    empty_b = full_b + kMaxBSlots;
    tmem_full = empty_b + kMaxBSlots;
    tmem_empty = tmem_full + 2;
    zfull = tmem_empty + 2;
    tmem_ptr = reinterpret_cast<uint32_t*>(zfull + 2);
  }
};

// Epilogue of one accumulator tile (128 rows x BLOCK_N fp32 in TMEM), executed by ONE of the two epilogue groups
// (4 warps each; group g serves accumulator stage g, staging block g, named barrier 1+g, so two tiles drain
// concurrently):  TMEM -> registers (+bias) -> bf16 -> 128B-swizzled staging block in shared memory -> TMA store
// (clipped at the tensor edge).  BatchNorm batch statistics are column sums over the staged (bf16-rounded) block,
// kept in registers across tiles (a thread always owns the same channel pair) and flushed with fp64 atomics.
template <int BLOCK_N>
struct StatRegs {
  float v[BLOCK_N / 64][4];
  int n_tile;
  __device__ void clear() {
#pragma unroll
    for (int i = 0; i < BLOCK_N / 64; ++i) v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.f;
  }
  __device__ void flush(const ConvParams& p, int lane) {
    if (n_tile < 0) return;
#pragma unroll
    for (int i = 0; i < BLOCK_N / 64; ++i) {
      const int ch = n_tile * BLOCK_N + i * 64 + 2 * lane;
      if (p.bn_sums) {
        // v = {sum dy (ch), sum dy (ch+1), sum dy*z (ch), sum dy*z (ch+1)}; xhat = (z - mean) * invstd
        const float m0 = p.bn_mean[ch], m1 = p.bn_mean[ch + 1], i0 = p.bn_invstd[ch], i1 = p.bn_invstd[ch + 1];
        atomicAdd(p.bn_sums + ch, (double)v[i][0]);
        atomicAdd(p.bn_sums + ch + 1, (double)v[i][1]);
        atomicAdd(p.bn_sums + p.cout + ch, (double)(i0 * (v[i][2] - m0 * v[i][0])));
        atomicAdd(p.bn_sums + p.cout + ch + 1, (double)(i1 * (v[i][3] - m1 * v[i][1])));
      } else {
        atomicAdd(p.stat_sum + ch, (double)v[i][0]);
        atomicAdd(p.stat_sum + ch + 1, (double)v[i][1]);
        atomicAdd(p.stat_sumsq + ch, (double)v[i][2]);
        atomicAdd(p.stat_sumsq + ch + 1, (double)v[i][3]);
      }
    }
    clear();
  }
};

template <int BLOCK_N, bool HAS_BIAS, bool PAIR = false>
__device__ __forceinline__ void epilogue_tile(const ConvParams& p, uint8_t* staging, int bar_id, uint32_t tmem_acc, int q,
                                              int lane, bool valid_row, int n_tile, int w0, int h0, int n0,
                                              StatRegs<BLOCK_N>& st, uint64_t* tmem_empty_bar, uint8_t* zbuf = nullptr,
                                              uint64_t* zbar = nullptr, uint32_t* zphase = nullptr) {
  const int row = q * 32 + lane;
  const bool storer = (q == 0 && lane == 0);
  const bool bnred = p.bn_sums != nullptr;
  const bool do_stats = p.stat_sum != nullptr || bnred;
  const int col0 = n_tile * BLOCK_N;
  if (do_stats && st.n_tile != n_tile) {
    st.flush(p, lane);
    st.n_tile = n_tile;
  }
  const uint32_t sbase = smem_u32(staging);
  const uint32_t srow = sbase + row * 128;
#pragma unroll
  for (int blk = 0; blk < BLOCK_N / 64; ++blk) {
    // first output channel of this 64-wide block; ConvT fprop (mode 2): column = (quadrant, channel), one store map per
    // quadrant, so an N tile may span several quadrants
    int map_idx = 0, chb = col0 + blk * 64;
    if (p.mode == 2) {
      map_idx = chb / p.cout;
      chb -= map_idx * p.cout;
    }
    if (bnred) {
      named_bar_sync(bar_id, 128);   // every thread is done with the previous z block
      if (storer) {
        // z tile of the layer whose gradient this is (same pixels / channels as the output block), 128B-swizzled like staging
        mbar_arrive_expect_tx(zbar, kStagingBytes);
        tma_load_4d(zbuf, &p.map_z, zbar, chb, w0, h0, n0);
      }
    }
    uint32_t packed[32];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      // 32 columns at a time: at most 32 raw + 32 packed registers live
      uint32_t r[32];
      tmem_ld_32x32(tmem_acc + blk * 64 + half * 32, r);
      tmem_ld_wait();
      if (half == 1 && blk == BLOCK_N / 64 - 1) {
        // accumulator fully drained: hand the TMEM stage back to the MMA warp
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(tmem_empty_bar, 0); else mbar_arrive(tmem_empty_bar);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float lo = __uint_as_float(r[2 * j]), hi = __uint_as_float(r[2 * j + 1]);
        if (HAS_BIAS) {
          lo += __ldg(p.bias + chb + half * 32 + 2 * j);
          hi += __ldg(p.bias + chb + half * 32 + 2 * j + 1);
        }
        packed[half * 16 + j] = valid_row ? pack_bf16x2(lo, hi) : 0u;
      }
    }
    // the TMEM read and the conversion above overlap the TMA store of the previous block, which must have finished
    // READING the staging block before it is overwritten
    if (storer) tma_store_wait_read<0>();
    named_bar_sync(bar_id, 128);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t chunk = (uint32_t)j ^ (uint32_t)(row & 7);
      st_shared_v4(srow + chunk * 16, packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
    }
    fence_proxy_async_smem();
    named_bar_sync(bar_id, 128);
    if (storer) {
      tma_store_4d(&p.map_y[map_idx], staging, chb, w0, h0, n0);
      tma_store_commit();
    }
    if (do_stats) {
      // thread (q, lane): channel pair `lane` of this 64-channel block, rows q*32 .. q*32+31
      float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
      const uint32_t base = sbase + (uint32_t)(lane & 3) * 4 + (uint32_t)(q * 32) * 128;
      const uint32_t cgrp = (uint32_t)lane >> 2;
      if (bnred) {
        // BatchNorm-backward reduction: dy = dA * [relu(z*scale+shift) > 0]; accumulate sum dy and sum dy*z
        const int ch = chb + 2 * lane;
        const float sc0 = __ldg(p.bn_scale + ch), sc1 = __ldg(p.bn_scale + ch + 1);
        const float sh0 = __ldg(p.bn_shift + ch), sh1 = __ldg(p.bn_shift + ch + 1);
        const uint32_t zbase = smem_u32(zbuf) + (uint32_t)(lane & 3) * 4 + (uint32_t)(q * 32) * 128;
        mbar_wait(zbar, *zphase);
        *zphase ^= 1;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const uint32_t off = i * 128 + ((cgrp ^ (uint32_t)(i & 7)) << 4);
          const uint32_t wd = ld_shared_u32(base + off), wz = ld_shared_u32(zbase + off);
          const float dlo = __uint_as_float(wd << 16), dhi = __uint_as_float(wd & 0xffff0000u);
          const float zlo = __uint_as_float(wz << 16), zhi = __uint_as_float(wz & 0xffff0000u);
          const float ylo = fmaf(zlo, sc0, sh0) > 0.f ? dlo : 0.f;
          const float yhi = fmaf(zhi, sc1, sh1) > 0.f ? dhi : 0.f;
          s1a += ylo;
          s1b += yhi;
          s2a = fmaf(ylo, zlo, s2a);
          s2b = fmaf(yhi, zhi, s2b);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const uint32_t word = ld_shared_u32(base + i * 128 + ((cgrp ^ (uint32_t)(i & 7)) << 4));
          const float lo = __uint_as_float(word << 16), hi = __uint_as_float(word & 0xffff0000u);
          s1a += lo;
          s1b += hi;
          s2a = fmaf(lo, lo, s2a);
          s2b = fmaf(hi, hi, s2b);
        }
      }
      st.v[blk][0] += s1a;
      st.v[blk][1] += s1b;
      st.v[blk][2] += s2a;
      st.v[blk][3] += s2b;
    }
  }
}

__device__ __forceinline__ void init_common(const ConvParams& p, uint8_t* smem, Bars& bars, int warp, int lane, int a_stages,
                                            int b_slots) {
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.map_a[0]);
    prefetch_tensormap(&p.map_b);
    prefetch_tensormap(&p.map_y[0]);
    if (p.mode >= 2)
      for (int i = 1; i < 4; ++i) prefetch_tensormap(p.mode == 3 ? &p.map_a[i] : &p.map_y[i]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < a_stages; ++i) {
      mbar_init(&bars.full_a[i], 1);
      mbar_init(&bars.empty_a[i], 1);
    }
    for (int i = 0; i < b_slots; ++i) {
      mbar_init(&bars.full_b[i], 1);
      mbar_init(&bars.empty_b[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.tmem_full[i], 1);
      mbar_init(&bars.tmem_empty[i], 4);  // one arrive per epilogue warp
      mbar_init(&bars.zfull[i], 1);
    }
    fence_barrier_init();
  }
}

template <int BLOCK_N, bool HAS_BIAS>
__global__ void __launch_bounds__(kConvThreads, 1) tc_conv_kernel(const __grid_constant__ ConvParams p) {
  constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Bars bars(smem + p.off_bars);
  // warp index made provably warp-uniform for the compiler (role branches and everything computed in them stay uniform)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int ksteps = p.taps * p.chunks_per_tap;
  const int stages = p.stages;

  init_common(p, smem, bars, warp, lane, stages, 0);
  if (warp == 2) tmem_alloc<kTmemCols>(bars.tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *bars.tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        // m fastest: a CTA keeps its N tile (weights, statistics channels) for long runs of tiles
        const int n_tile = tile / p.num_m_tiles, m_tile = tile % p.num_m_tiles;
        const int tw = m_tile % p.tiles_w, th = (m_tile / p.tiles_w) % p.tiles_h, tn = m_tile / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.pw, h0 = th * p.ph, n0 = tn * p.nb;
        for (int s = 0; s < ksteps; ++s) {
          const int t = s / p.chunks_per_tap, chunk = s % p.chunks_per_tap;
          int dh = 0, dw = 0, mi = 0;
          if (p.mode == 1) {
            dh = t / 3 - 1;
            dw = t % 3 - 1;
          } else if (p.mode == 3) {
            mi = t;
          }
          mbar_wait(&bars.empty_a[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          mbar_arrive_expect_tx(&bars.full_a[stage], kStageBytes);
          tma_load_4d(sa, &p.map_a[mi], &bars.full_a[stage], chunk * kBlockK, w0 + dw, h0 + dh, n0);
          tma_load_2d(sa + kABytes, &p.map_b, &bars.full_a[stage], s * kBlockK, n_tile * BLOCK_N);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp, one elected lane per instruction) =====================
    {   // whole warp, converged: operands stay warp-uniform, one elected lane issues (see umma_bf16_warp)
      constexpr uint32_t idesc = make_idesc_bf16(kTileM, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&bars.tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int s = 0; s < ksteps; ++s) {
          mbar_wait(&bars.full_a[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t da = make_smem_desc(sa, 0, 1024);
          const uint64_t db = make_smem_desc(sa + kABytes, 0, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 bf16 = 32 bytes inside the 128B swizzle row: +2 in the (addr >> 4) field
            umma_bf16_warp(d_tmem, da + 2 * k, db + 2 * k, idesc, (s > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_warp(&bars.empty_a[stage]);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_warp(&bars.tmem_full[acc]);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue: group g = (warp-4)/4 drains the tiles with (it & 1) == g =====================
    const int g = (warp - 4) >> 2;
    const int q = (warp - 4) & 3;  // TMEM lane quadrant
    const int row = q * 32 + lane;
    const int pw_i = row % p.pw, ph_i = (row / p.pw) % p.ph, nb_i = row / (p.pw * p.ph);
    uint8_t* staging = smem + p.off_staging + g * kStagingBytes;
    uint8_t* zbuf = smem + p.off_zbuf + g * kStagingBytes;
    uint32_t zphase = 0;
    StatRegs<BLOCK_N> st;
    st.clear();
    st.n_tile = -1;
    int it = g;
    for (int tile = blockIdx.x + g * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x, it += 2) {
      const int n_tile = tile / p.num_m_tiles, m_tile = tile % p.num_m_tiles;
      const int tw = m_tile % p.tiles_w, th = (m_tile / p.tiles_w) % p.tiles_h, tn = m_tile / (p.tiles_w * p.tiles_h);
      const int w0 = tw * p.pw, h0 = th * p.ph, n0 = tn * p.nb;
      const bool valid = (w0 + pw_i) < p.w && (h0 + ph_i) < p.h && (n0 + nb_i) < p.n;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&bars.tmem_full[g], acc_phase);
      tcgen05_fence_after();
      epilogue_tile<BLOCK_N, HAS_BIAS>(p, staging, 1 + g, tmem_base + ((uint32_t)(q * 32) << 16) + g * BLOCK_N, q, lane, valid,
                                       n_tile, w0, h0, n0, st, &bars.tmem_empty[g], zbuf, &bars.zfull[g], &zphase);
    }
    if (p.stat_sum || p.bn_sums) st.flush(p, lane);
    if (q == 0 && lane == 0) tma_store_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// Work distribution of the CTA-pair kernels.  Work item = (N tile, pair of pixel tiles).  When the pairs split evenly
// over the N tiles, pair p keeps ONE N tile (weight rows, statistics channels) and the groups of pairs walk the pixel
// tiles in step, so a pixel tile is fetched by all N tiles at about the same time and comes from DRAM once (the input
// of a deep layer, 134 MB at batch 64, does not fit the L2: pixel-fastest order streamed it once per N tile).
struct PairSched {
  int iters, n_fixed, m0, mstep, m_pairs, pair_id, num_pairs;
  bool grouped;
  __device__ PairSched(int pair_id_, int num_pairs_, int m_pairs_, int num_n_tiles) {
    pair_id = pair_id_; num_pairs = num_pairs_; m_pairs = m_pairs_;
    grouped = num_n_tiles > 1 && num_pairs % num_n_tiles == 0;
    if (grouped) {
      mstep = num_pairs / num_n_tiles;
      n_fixed = pair_id / mstep;
      m0 = pair_id % mstep;
      iters = m0 < m_pairs ? (m_pairs - m0 + mstep - 1) / mstep : 0;
    } else {
      mstep = n_fixed = m0 = 0;
      const int total = m_pairs * num_n_tiles;
      iters = pair_id < total ? (total - pair_id + num_pairs - 1) / num_pairs : 0;
    }
  }
  __device__ __forceinline__ void get(int k, int& n_tile, int& m_pair) const {
    if (grouped) {
      n_tile = n_fixed;
      m_pair = m0 + k * mstep;
    } else {
      const int pt = pair_id + k * num_pairs;
      n_tile = pt / m_pairs;
      m_pair = pt % m_pairs;
    }
  }
};

// -------------------------------------------------------------------------------------------------
// CTA-pair variant of tc_conv_kernel (cta_group::2).  The in-kernel counters showed the single-CTA MMA bound by
// shared-memory operand fetch (~80 B/cycle: (4096 + 32 N)/80 cycles per M=128 MMA).  A pair of CTAs on adjacent SMs
// issues ONE tcgen05.mma of M = 256: each CTA supplies the A rows of its own 128-pixel tile and only HALF of the B
// (weight) rows, so the per-SM operand traffic per MMA drops to 4096 + 16 N bytes.
//   * cluster (2,1,1); pair p handles (N tile, pixel tiles 2j and 2j+1); CTA rank r owns pixel tile 2j+r
//   * full barriers live in the leader (rank 0): both producers' TMA loads complete_tx there; count 2
//   * MMA issued by the leader only; tcgen05.commit multicast releases ring slots / publishes accumulators in BOTH CTAs
//   * every CTA drains its own 128 TMEM lanes with its two epilogue groups; tmem_empty arrivals go to the leader
// -------------------------------------------------------------------------------------------------
template <int BLOCK_N, bool HAS_BIAS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
    tc_conv2_kernel(const __grid_constant__ ConvParams p) {
  constexpr int kBHalfBytes = (BLOCK_N / 2) * kBlockK * 2;
  constexpr int kStageBytes = kABytes + kBHalfBytes;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Bars bars(smem + p.off_bars);
  // warp index made provably warp-uniform for the compiler (role branches and everything computed in them stay uniform)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // rank in the (2,1,1) cluster == %cluster_ctarank; taken from blockIdx so the compiler knows it is uniform (an asm
  // output is treated as thread-varying and would put every UTCHMMA of the leader branch in a waterfall loop)
  const uint32_t rank = blockIdx.x & 1u;
  const bool leader = rank == 0;
  const int num_pairs = gridDim.x >> 1, pair_id = blockIdx.x >> 1;
  const int m_pairs = (p.num_m_tiles + 1) >> 1;
  const PairSched sched(pair_id, num_pairs, m_pairs, p.num_n_tiles);
  const int ksteps = p.taps * p.chunks_per_tap;
  const int stages = p.stages;

  cluster_sync_all();  // both CTAs of the pair are resident before any cross-CTA traffic / 2-SM TMEM allocation
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.map_a[0]);
    prefetch_tensormap(&p.map_b);
    prefetch_tensormap(&p.map_y[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&bars.full_a[i], 2);   // one arrival per producer of the pair (used in the leader only)
      mbar_init(&bars.empty_a[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.tmem_full[i], 1);
      mbar_init(&bars.tmem_empty[i], 8);  // 4 epilogue warps of each CTA (used in the leader only)
      mbar_init(&bars.zfull[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm<kTmemCols>(bars.tmem_ptr);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *bars.tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kk = 0; kk < sched.iters; ++kk) {
        int n_tile, m_pair;
        sched.get(kk, n_tile, m_pair);
        const int m_tile = 2 * m_pair + (int)rank;
        const int tw = m_tile % p.tiles_w, th = (m_tile / p.tiles_w) % p.tiles_h, tn = m_tile / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.pw, h0 = th * p.ph, n0 = tn * p.nb;   // a tile past the end is all out-of-bounds: zero fill
        for (int s = 0; s < ksteps; ++s) {
          const int t = s / p.chunks_per_tap, chunk = s % p.chunks_per_tap;
          int dh = 0, dw = 0, mi = 0;
          if (p.mode == 1) {
            dh = t / 3 - 1;
            dw = t % 3 - 1;
          } else if (p.mode == 3) {
            mi = t;
          }
          mbar_wait(&bars.empty_a[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          if (leader)
            mbar_arrive_expect_tx(&bars.full_a[stage], 2 * kStageBytes);
          else
            mbar_arrive_cluster(&bars.full_a[stage], 0);
          tma_load_4d_2sm(sa, &p.map_a[mi], &bars.full_a[stage], chunk * kBlockK, w0 + dw, h0 + dh, n0);
          tma_load_2d_2sm(sa + kABytes, &p.map_b, &bars.full_a[stage], s * kBlockK,
                          n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2));
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA; whole warp, one elected lane per instruction) =====================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int kk = 0; kk < sched.iters; ++kk, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&bars.tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int s = 0; s < ksteps; ++s) {
          mbar_wait(&bars.full_a[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t da = make_smem_desc(sa, 0, 1024);
          const uint64_t db = make_smem_desc(sa + kABytes, 0, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16_2sm_warp(d_tmem, da + 2 * k, db + 2 * k, idesc, (s > 0 || k > 0) ? 1u : 0u);
          umma_commit_2sm_warp(&bars.empty_a[stage]);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm_warp(&bars.tmem_full[acc]);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs; group g drains the pair tiles with (it & 1) == g) =====================
    const int g = (warp - 4) >> 2;
    const int q = (warp - 4) & 3;
    const int row = q * 32 + lane;
    const int pw_i = row % p.pw, ph_i = (row / p.pw) % p.ph, nb_i = row / (p.pw * p.ph);
    uint8_t* staging = smem + p.off_staging + g * kStagingBytes;
    uint8_t* zbuf = smem + p.off_zbuf + g * kStagingBytes;
    uint32_t zphase = 0;
    StatRegs<BLOCK_N> st;
    st.clear();
    st.n_tile = -1;
    int it = g;
    for (int kk = g; kk < sched.iters; kk += 2, it += 2) {
      int n_tile, m_pair;
      sched.get(kk, n_tile, m_pair);
      const int m_tile = 2 * m_pair + (int)rank;
      const int tw = m_tile % p.tiles_w, th = (m_tile / p.tiles_w) % p.tiles_h, tn = m_tile / (p.tiles_w * p.tiles_h);
      const int w0 = tw * p.pw, h0 = th * p.ph, n0 = tn * p.nb;
      const bool valid = (w0 + pw_i) < p.w && (h0 + ph_i) < p.h && (n0 + nb_i) < p.n;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&bars.tmem_full[g], acc_phase);
      tcgen05_fence_after();
      epilogue_tile<BLOCK_N, HAS_BIAS, true>(p, staging, 1 + g, tmem_base + ((uint32_t)(q * 32) << 16) + g * BLOCK_N, q, lane,
                                             valid, n_tile, w0, h0, n0, st, &bars.tmem_empty[g], zbuf, &bars.zfull[g], &zphase);
    }
    if (p.stat_sum || p.bn_sums) st.flush(p, lane);
    if (q == 0 && lane == 0) tma_store_wait_all();
  }

  tcgen05_fence_before();
  cluster_sync_all();   // the peer may still read this CTA's shared memory / signal its barriers until here
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc_2sm<kTmemCols>(tmem_base);
  }
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kConvThreads, 1) tc_conv_halo_kernel(const __grid_constant__ ConvParams p) {
  constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Bars bars(smem + p.off_bars);
  // warp index made provably warp-uniform for the compiler (role branches and everything computed in them stay uniform)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int cpt = p.chunks_per_tap;
  const int a_stages = p.stages, b_slots = p.b_slots;
  const bool resident = p.resident_b != 0;

  init_common(p, smem, bars, warp, lane, a_stages, b_slots);
  if (warp == 2) tmem_alloc<kTmemCols>(bars.tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *bars.tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      bool first = true;
      long long dbg_a = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_tile = tile / p.num_m_tiles, m_tile = tile % p.num_m_tiles;
        const int tw = m_tile % p.tiles_w, th = (m_tile / p.tiles_w) % p.tiles_h, tn = m_tile / (p.tiles_w * p.tiles_h);
        const int w0 = tw * 8, h0 = th * 16;
        for (int chunk = 0; chunk < cpt; ++chunk) {
          {
            DBG_T0();
            mbar_wait(&bars.empty_a[sa], pa ^ 1);
            DBG_ADD(dbg_a);
          }
          mbar_arrive_expect_tx(&bars.full_a[sa], kHaloBytes);
          tma_load_4d(smem + sa * kHaloStride, &p.map_a[0], &bars.full_a[sa], chunk * kBlockK, w0 - 1, h0 - 1, tn);
          if (++sa == a_stages) {
            sa = 0;
            pa ^= 1;
          }
          for (int tap = 0; tap < 9; ++tap) {
            if (resident) {
              if (first) {
                const int slot = chunk * 9 + tap;
                mbar_arrive_expect_tx(&bars.full_b[slot], kBBytes);
                tma_load_2d(smem + p.off_b + slot * kBBytes, &p.map_b, &bars.full_b[slot], (tap * cpt + chunk) * kBlockK,
                            n_tile * BLOCK_N);
              }
            } else {
              mbar_wait(&bars.empty_b[sb], pb ^ 1);
              mbar_arrive_expect_tx(&bars.full_b[sb], kBBytes);
              tma_load_2d(smem + p.off_b + sb * kBBytes, &p.map_b, &bars.full_b[sb], (tap * cpt + chunk) * kBlockK,
                          n_tile * BLOCK_N);
              if (++sb == b_slots) {
                sb = 0;
                pb ^= 1;
              }
            }
          }
        }
        first = false;
      }
      if (kDbg && lane == 0 && blockIdx.x < 160) g_dbg[blockIdx.x * 8 + 0] = dbg_a;
    }
    __syncwarp();
  } else if (warp == 1) {
    {   // whole warp, converged: operands stay warp-uniform, one elected lane issues (see umma_bf16_warp)
      constexpr uint32_t idesc = make_idesc_bf16(kTileM, BLOCK_N, 0, 0);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      bool first = true;
      int it = 0;
      long long dbg_fa = 0, dbg_fb = 0, dbg_te = 0;
      const long long dbg_start = DBG_NOW();
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        {
          DBG_T0();
          mbar_wait(&bars.tmem_empty[acc], acc_phase ^ 1);
          DBG_ADD(dbg_te);
        }
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int chunk = 0; chunk < cpt; ++chunk) {
          {
            DBG_T0();
            mbar_wait(&bars.full_a[sa], pa);
            DBG_ADD(dbg_fa);
          }
          tcgen05_fence_after();
          const uint32_t halo = smem_u32(smem + sa * kHaloStride);
          for (int tap = 0; tap < 9; ++tap) {
            uint32_t bslot;
            if (resident) {
              bslot = chunk * 9 + tap;
              if (first) {
                mbar_wait(&bars.full_b[bslot], 0);
                tcgen05_fence_after();
              }
            } else {
              bslot = sb;
              {
                DBG_T0();
                mbar_wait(&bars.full_b[sb], pb);
                DBG_ADD(dbg_fb);
              }
              tcgen05_fence_after();
            }
            // tap (r,s): output pixel (ph,pw) reads halo pixel (ph+r, pw+s) = halo row (ph+r)*10 + pw+s
            const uint32_t a0 = halo + ((tap / 3) * 10 + (tap % 3)) * 128;
            const uint64_t da = make_smem_desc(a0, 0, 10 * 128);
            const uint64_t db = make_smem_desc(smem_u32(smem + p.off_b + bslot * kBBytes), 0, 1024);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_bf16_warp(d_tmem, da + 2 * k, db + 2 * k, idesc, (chunk > 0 || tap > 0 || k > 0) ? 1u : 0u);
            if (!resident) {
              umma_commit_warp(&bars.empty_b[sb]);
              if (++sb == b_slots) {
                sb = 0;
                pb ^= 1;
              }
            }
          }
          umma_commit_warp(&bars.empty_a[sa]);
          if (++sa == a_stages) {
            sa = 0;
            pa ^= 1;
          }
        }
        umma_commit_warp(&bars.tmem_full[acc]);
        first = false;
      }
      if (kDbg && lane == 0 && blockIdx.x < 160) {
        g_dbg[blockIdx.x * 8 + 1] = dbg_fa;
        g_dbg[blockIdx.x * 8 + 2] = dbg_fb;
        g_dbg[blockIdx.x * 8 + 3] = dbg_te;
        g_dbg[blockIdx.x * 8 + 4] = DBG_NOW() - dbg_start;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int g = (warp - 4) >> 2;
    const int q = (warp - 4) & 3;
    const int row = q * 32 + lane;
    const int pw_i = row & 7, ph_i = row >> 3;
    long long dbg_tf = 0, dbg_epi = 0;
    const long long dbg_start = DBG_NOW();
    uint8_t* staging = smem + p.off_staging + g * kStagingBytes;
    uint8_t* zbuf = smem + p.off_zbuf + g * kStagingBytes;
    uint32_t zphase = 0;
    StatRegs<BLOCK_N> st;
    st.clear();
    st.n_tile = -1;
    int it = g;
    for (int tile = blockIdx.x + g * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x, it += 2) {
      const int n_tile = tile / p.num_m_tiles, m_tile = tile % p.num_m_tiles;
      const int tw = m_tile % p.tiles_w, th = (m_tile / p.tiles_w) % p.tiles_h, tn = m_tile / (p.tiles_w * p.tiles_h);
      const int w0 = tw * 8, h0 = th * 16;
      const bool valid = (w0 + pw_i) < p.w && (h0 + ph_i) < p.h;
      const uint32_t acc_phase = (it >> 1) & 1;
      {
        DBG_T0();
        mbar_wait(&bars.tmem_full[g], acc_phase);
        DBG_ADD(dbg_tf);
      }
      tcgen05_fence_after();
      {
        DBG_T0();
        epilogue_tile<BLOCK_N, false>(p, staging, 1 + g, tmem_base + ((uint32_t)(q * 32) << 16) + g * BLOCK_N, q, lane, valid,
                                      n_tile, w0, h0, tn, st, &bars.tmem_empty[g], zbuf, &bars.zfull[g], &zphase);
        DBG_ADD(dbg_epi);
      }
    }
    if (kDbg && g == 0 && q == 0 && lane == 0 && blockIdx.x < 160) {
      g_dbg[blockIdx.x * 8 + 5] = dbg_tf;
      g_dbg[blockIdx.x * 8 + 6] = dbg_epi;
      g_dbg[blockIdx.x * 8 + 7] = DBG_NOW() - dbg_start;
    }
    if (p.stat_sum || p.bn_sums) st.flush(p, lane);
    if (q == 0 && lane == 0) tma_store_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// -------------------------------------------------------------------------------------------------
// CTA-pair variant of the halo kernel: each CTA loads the halo of its own 16x8 pixel tile and HALF of the weight rows
// of every K step (resident or streamed); the leader issues M = 256 MMAs.  Same barrier protocol as tc_conv2_kernel.
// -------------------------------------------------------------------------------------------------
template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
    tc_conv_halo2_kernel(const __grid_constant__ ConvParams p) {
  constexpr int kBHalfBytes = (BLOCK_N / 2) * kBlockK * 2;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Bars bars(smem + p.off_bars);
  // warp index made provably warp-uniform for the compiler (role branches and everything computed in them stay uniform)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // rank in the (2,1,1) cluster == %cluster_ctarank; taken from blockIdx so the compiler knows it is uniform (an asm
  // output is treated as thread-varying and would put every UTCHMMA of the leader branch in a waterfall loop)
  const uint32_t rank = blockIdx.x & 1u;
  const bool leader = rank == 0;
  const int num_pairs = gridDim.x >> 1, pair_id = blockIdx.x >> 1;
  const int m_pairs = (p.num_m_tiles + 1) >> 1;
  const PairSched sched(pair_id, num_pairs, m_pairs, p.num_n_tiles);
  const int cpt = p.chunks_per_tap;
  const int a_stages = p.stages, b_slots = p.b_slots;
  const bool resident = p.resident_b != 0;

  cluster_sync_all();
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.map_a[0]);
    prefetch_tensormap(&p.map_b);
    prefetch_tensormap(&p.map_y[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < a_stages; ++i) {
      mbar_init(&bars.full_a[i], 2);
      mbar_init(&bars.empty_a[i], 1);
    }
    for (int i = 0; i < b_slots; ++i) {
      mbar_init(&bars.full_b[i], 2);
      mbar_init(&bars.empty_b[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.tmem_full[i], 1);
      mbar_init(&bars.tmem_empty[i], 8);
      mbar_init(&bars.zfull[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm<kTmemCols>(bars.tmem_ptr);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *bars.tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      bool first = true;
      for (int kk = 0; kk < sched.iters; ++kk) {
        int n_tile, m_pair;
        sched.get(kk, n_tile, m_pair);
        const int m_tile = 2 * m_pair + (int)rank;
        const int tw = m_tile % p.tiles_w, th = (m_tile / p.tiles_w) % p.tiles_h, tn = m_tile / (p.tiles_w * p.tiles_h);
        const int w0 = tw * 8, h0 = th * 16;
        const int brow = n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2);
        for (int chunk = 0; chunk < cpt; ++chunk) {
          mbar_wait(&bars.empty_a[sa], pa ^ 1);
          if (leader) mbar_arrive_expect_tx(&bars.full_a[sa], 2 * kHaloBytes); else mbar_arrive_cluster(&bars.full_a[sa], 0);
          tma_load_4d_2sm(smem + sa * kHaloStride, &p.map_a[0], &bars.full_a[sa], chunk * kBlockK, w0 - 1, h0 - 1, tn);
          if (++sa == a_stages) {
            sa = 0;
            pa ^= 1;
          }
          for (int tap = 0; tap < 9; ++tap) {
            if (resident) {
              if (first) {
                const int slot = chunk * 9 + tap;
                if (leader) mbar_arrive_expect_tx(&bars.full_b[slot], 2 * kBHalfBytes); else mbar_arrive_cluster(&bars.full_b[slot], 0);
                tma_load_2d_2sm(smem + p.off_b + slot * kBHalfBytes, &p.map_b, &bars.full_b[slot], (tap * cpt + chunk) * kBlockK, brow);
              }
            } else {
              mbar_wait(&bars.empty_b[sb], pb ^ 1);
              if (leader) mbar_arrive_expect_tx(&bars.full_b[sb], 2 * kBHalfBytes); else mbar_arrive_cluster(&bars.full_b[sb], 0);
              tma_load_2d_2sm(smem + p.off_b + sb * kBHalfBytes, &p.map_b, &bars.full_b[sb], (tap * cpt + chunk) * kBlockK, brow);
              if (++sb == b_slots) {
                sb = 0;
                pb ^= 1;
              }
            }
          }
        }
        first = false;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BLOCK_N, 0, 0);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      bool first = true;
      int it = 0;
      for (int kk = 0; kk < sched.iters; ++kk, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&bars.tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int chunk = 0; chunk < cpt; ++chunk) {
          mbar_wait(&bars.full_a[sa], pa);
          tcgen05_fence_after();
          const uint32_t halo = smem_u32(smem + sa * kHaloStride);
          for (int tap = 0; tap < 9; ++tap) {
            uint32_t bslot;
            if (resident) {
              bslot = chunk * 9 + tap;
              if (first) {
                mbar_wait(&bars.full_b[bslot], 0);
                tcgen05_fence_after();
              }
            } else {
              bslot = sb;
              mbar_wait(&bars.full_b[sb], pb);
              tcgen05_fence_after();
            }
            const uint32_t a0 = halo + ((tap / 3) * 10 + (tap % 3)) * 128;
            const uint64_t da = make_smem_desc(a0, 0, 10 * 128);
            const uint64_t db = make_smem_desc(smem_u32(smem + p.off_b + bslot * kBHalfBytes), 0, 1024);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_bf16_2sm_warp(d_tmem, da + 2 * k, db + 2 * k, idesc, (chunk > 0 || tap > 0 || k > 0) ? 1u : 0u);
            if (!resident) {
              umma_commit_2sm_warp(&bars.empty_b[sb]);
              if (++sb == b_slots) {
                sb = 0;
                pb ^= 1;
              }
            }
          }
          umma_commit_2sm_warp(&bars.empty_a[sa]);
          if (++sa == a_stages) {
            sa = 0;
            pa ^= 1;
          }
        }
        umma_commit_2sm_warp(&bars.tmem_full[acc]);
        first = false;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int g = (warp - 4) >> 2;
    const int q = (warp - 4) & 3;
    const int row = q * 32 + lane;
    const int pw_i = row & 7, ph_i = row >> 3;
    uint8_t* staging = smem + p.off_staging + g * kStagingBytes;
    uint8_t* zbuf = smem + p.off_zbuf + g * kStagingBytes;
    uint32_t zphase = 0;
    StatRegs<BLOCK_N> st;
    st.clear();
    st.n_tile = -1;
    int it = g;
    for (int kk = g; kk < sched.iters; kk += 2, it += 2) {
      int n_tile, m_pair;
      sched.get(kk, n_tile, m_pair);
      const int m_tile = 2 * m_pair + (int)rank;
      const int tw = m_tile % p.tiles_w, th = (m_tile / p.tiles_w) % p.tiles_h, tn = m_tile / (p.tiles_w * p.tiles_h);
      const int w0 = tw * 8, h0 = th * 16;
      const bool valid = (w0 + pw_i) < p.w && (h0 + ph_i) < p.h && tn < p.n;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&bars.tmem_full[g], acc_phase);
      tcgen05_fence_after();
      epilogue_tile<BLOCK_N, false, true>(p, staging, 1 + g, tmem_base + ((uint32_t)(q * 32) << 16) + g * BLOCK_N, q, lane,
                                          valid, n_tile, w0, h0, tn, st, &bars.tmem_empty[g], zbuf, &bars.zfull[g], &zphase);
    }
    if (p.stat_sum || p.bn_sums) st.flush(p, lane);
    if (q == 0 && lane == 0) tma_store_wait_all();
  }

  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc_2sm<kTmemCols>(tmem_base);
  }
}

// =================================================================================================
// weight-gradient kernel: dw[cu, t, cs] += sum_pixels U[p, cu] * S[gather(p, t), cs]
// both operands MN-major in shared memory (rows = pixels = K, 64-channel slabs of 128 B)
// =================================================================================================
struct WgradParams {
  CUtensorMap map_u;
  CUtensorMap map_s[4];
  int mode, taps;
  int pw, ph, nb, tiles_w, tiles_h;
  int num_ptiles;          // pixel tiles (K blocks of 128 pixels)
  int ptiles_per_split, splits;
  int cu_tiles, cs_tiles;  // output tiles of BM x BN
  int cu, cs;
  float* dw;
  // deterministic mode (UNETK_TC_DETERMINISTIC): every (output tile, split) work item STORES its partial sums to
  // partial[split][...] (same element layout as dw) and wgrad_reduce_kernel adds the splits to dw in index order
  float* partial;
  int64_t out_elems;
};

__device__ __forceinline__ void st_global_v4(float* gptr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gptr), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)),
               "f"(__uint_as_float(c)), "f"(__uint_as_float(d))
               : "memory");
}
// accumulate (atomics, default) or store (deterministic mode) four consecutive fp32 weight-gradient elements
__device__ __forceinline__ void wgrad_out_v4(bool det, float* gptr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  if (det) st_global_v4(gptr, a, b, c, d); else red_add_v4(gptr, a, b, c, d);
}

__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int64_t out_elems,
                                                           float* __restrict__ dw) {
  const int64_t n4 = out_elems / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 acc = reinterpret_cast<const float4*>(partial)[i];
    for (int s = 1; s < splits; ++s) {
      const float4 v = reinterpret_cast<const float4*>(partial + (int64_t)s * out_elems)[i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float4 d = reinterpret_cast<float4*>(dw)[i];
    d.x += acc.x; d.y += acc.y; d.z += acc.z; d.w += acc.w;
    reinterpret_cast<float4*>(dw)[i] = d;
  }
}

template <int BM_SLABS, int BN_SLABS>
struct WgradCfg {
  static constexpr int BLOCK_N = BN_SLABS * 64;
  static constexpr int kUBytes = BM_SLABS * kABytes;
  static constexpr int kSBytes = BN_SLABS * kABytes;
  static constexpr int kStageBytes = kUBytes + kSBytes;
  static constexpr int kStages = (200 * 1024) / kStageBytes;
  static constexpr int kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static constexpr int kBarOff = kStages * kStageBytes;
  static constexpr int kSmemBytes = kBarOff + (2 * kStages + 4) * 8 + 16 + 1024;
};

template <int BM_SLABS, int BN_SLABS>
__global__ void __launch_bounds__(kNumThreads, 1) tc_wgrad_kernel(const __grid_constant__ WgradParams p) {
  using Cfg = WgradCfg<BM_SLABS, BN_SLABS>;
  constexpr int BLOCK_N = Cfg::BLOCK_N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  // warp index made provably warp-uniform for the compiler (role branches and everything computed in them stay uniform)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // work item = (cu tile, cs tile, tap, split)
  const int num_items = p.cu_tiles * p.cs_tiles * p.taps * p.splits;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.map_u);
    prefetch_tensormap(&p.map_s[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // item decoding shared by all roles: split fastest so that concurrently running CTAs share operand tiles in L2
  auto decode = [&](int item, int& cu_t, int& cs_t, int& tap, int& split) {
    split = item % p.splits;
    int r = item / p.splits;
    tap = r % p.taps;
    r /= p.taps;
    cs_t = r % p.cs_tiles;
    cu_t = r / p.cs_tiles;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int cu_t, cs_t, tap, split;
        decode(item, cu_t, cs_t, tap, split);
        int dh = 0, dw = 0, mi = 0;
        if (p.mode == 1) {
          dh = tap / 3 - 1;
          dw = tap % 3 - 1;
        } else if (p.mode == 2) {
          mi = tap;
        }
        const int pt0 = split * p.ptiles_per_split;
        const int pt1 = min(pt0 + p.ptiles_per_split, p.num_ptiles);
        for (int pt = pt0; pt < pt1; ++pt) {
          const int tw = pt % p.tiles_w, th = (pt / p.tiles_w) % p.tiles_h, tn = pt / (p.tiles_w * p.tiles_h);
          const int w0 = tw * p.pw, h0 = th * p.ph, n0 = tn * p.nb;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* su = smem + stage * Cfg::kStageBytes;
          uint8_t* ss = su + Cfg::kUBytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
#pragma unroll
          for (int i = 0; i < BM_SLABS; ++i)
            tma_load_4d(su + i * kABytes, &p.map_u, &full_bar[stage], (cu_t * BM_SLABS + i) * 64, w0, h0, n0);
#pragma unroll
          for (int i = 0; i < BN_SLABS; ++i)
            tma_load_4d(ss + i * kABytes, &p.map_s[mi], &full_bar[stage], (cs_t * BN_SLABS + i) * 64, w0 + dw, h0 + dh, n0);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    {   // whole warp, converged: operands stay warp-uniform, one elected lane issues (see umma_bf16_warp)
      // M = 128 always; with a single 64-channel slab the descriptor's slab stride is 0 and rows 64..127 of the
      // accumulator duplicate rows 0..63 (ignored by the epilogue)
      constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 1, 1);
      constexpr uint32_t lbo_a = BM_SLABS == 2 ? kABytes : 0;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        int cu_t, cs_t, tap, split;
        decode(item, cu_t, cs_t, tap, split);
        const int pt0 = split * p.ptiles_per_split;
        const int pt1 = min(pt0 + p.ptiles_per_split, p.num_ptiles);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int pt = pt0; pt < pt1; ++pt) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t su = smem_u32(smem + stage * Cfg::kStageBytes);
          // MN-major, 128B swizzle: 8 K-rows x 128 B atoms, SBO = 1024 (next 8 pixels), LBO = next 64-channel slab
          const uint64_t da = make_smem_desc(su, lbo_a, 1024);
          const uint64_t db = make_smem_desc(su + Cfg::kUBytes, kABytes, 1024);
#pragma unroll
          for (int k = 0; k < kTileM / 16; ++k) {
            // 16 pixels = 16 rows of 128 B = 2048 bytes: +128 in the (addr >> 4) field
            umma_bf16_warp(d_tmem, da + 128 * k, db + 128 * k, idesc, (pt > pt0 || k > 0) ? 1u : 0u);
          }
          umma_commit_warp(&empty_bar[stage]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_warp(&tmem_full_bar[acc]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    const int row = q * 32 + lane;
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      int cu_t, cs_t, tap, split;
      decode(item, cu_t, cs_t, tap, split);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const int cu_idx = cu_t * (BM_SLABS * 64) + row;
      const bool valid = row < BM_SLABS * 64 && cu_idx < p.cu;
      const bool det = p.partial != nullptr;
      float* dst = (det ? p.partial + (int64_t)split * p.out_elems : p.dw) + ((int64_t)cu_idx * p.taps + tap) * p.cs + cs_t * BLOCK_N;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N + c * 32, r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) wgrad_out_v4(det, dst + c * 32 + j, r[j], r[j + 1], r[j + 2], r[j + 3]);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// =================================================================================================
// 3x3 weight gradient, halo form.  Work item = (cu tile, 64-channel cs slab, kernel row r, pixel split).
// Per 128-pixel block (16 x 8 patch) ONE band of the shifted operand (16 x 10 pixels, rows h0-1+r ..) is loaded; the
// three taps s = 0,1,2 of kernel row r are the same band shifted by one pixel, i.e. by one 128-byte shared-memory row.
// A single MMA of N = 192 therefore covers all three taps: its B descriptor uses LBO = 128 B (atom j = tap s=j) and
// SBO = 10*128 B (next image row inside the band).  Versus the per-tap kernel this is 3x fewer, 3x wider MMAs and
// 3x less operand traffic.   dw[cu, 3r+s, cs] += sum_p U[p, cu] * S[p + (r-1, s-1), cs]
// =================================================================================================
constexpr int kBandBytes = 16 * 10 * 128;  // 20480

template <int BM_SLABS>
__global__ void __launch_bounds__(kNumThreads, 1) tc_wgrad3x3_kernel(const __grid_constant__ WgradParams p) {
  constexpr int kUBytes = BM_SLABS * kABytes;
  constexpr int kStageBytes = kUBytes + kBandBytes;
  constexpr int kStages = (220 * 1024) / kStageBytes > 8 ? 8 : (220 * 1024) / kStageBytes;
  constexpr int BLOCK_N = 192;
  constexpr uint32_t kTmemCols = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  // warp index made provably warp-uniform for the compiler (role branches and everything computed in them stay uniform)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int num_items = p.cu_tiles * p.cs_tiles * 3 * p.splits;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.map_u);
    prefetch_tensormap(&p.map_s[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_ptr_smem);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  auto decode = [&](int item, int& cu_t, int& cs_t, int& r, int& split) {
    split = item % p.splits;
    int q = item / p.splits;
    r = q % 3;
    q /= 3;
    cs_t = q % p.cs_tiles;
    cu_t = q / p.cs_tiles;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int cu_t, cs_t, r, split;
        decode(item, cu_t, cs_t, r, split);
        const int pt0 = split * p.ptiles_per_split;
        const int pt1 = min(pt0 + p.ptiles_per_split, p.num_ptiles);
        for (int pt = pt0; pt < pt1; ++pt) {
          const int tw = pt % p.tiles_w, th = (pt / p.tiles_w) % p.tiles_h, tn = pt / (p.tiles_w * p.tiles_h);
          const int w0 = tw * 8, h0 = th * 16;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* su = smem + stage * kStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
#pragma unroll
          for (int i = 0; i < BM_SLABS; ++i)
            tma_load_4d(su + i * kABytes, &p.map_u, &full_bar[stage], (cu_t * BM_SLABS + i) * 64, w0, h0, tn);
          tma_load_4d(su + kUBytes, &p.map_s[0], &full_bar[stage], cs_t * 64, w0 - 1, h0 - 1 + r, tn);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {   // whole warp, converged: operands stay warp-uniform, one elected lane issues (see umma_bf16_warp)
      // a single 64-channel slab runs as an M = 64 MMA (tools/umma_probe.cu: accumulator row r then lives in TMEM lane
      // (r/16)*32 + r%16, i.e. the first 16 lanes of every 32-lane quadrant)
      constexpr uint32_t idesc = make_idesc_bf16(BM_SLABS == 2 ? 128 : 64, BLOCK_N, 1, 1);
      constexpr uint32_t lbo_a = BM_SLABS == 2 ? kABytes : 0;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        int cu_t, cs_t, r, split;
        decode(item, cu_t, cs_t, r, split);
        const int pt0 = split * p.ptiles_per_split;
        const int pt1 = min(pt0 + p.ptiles_per_split, p.num_ptiles);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int pt = pt0; pt < pt1; ++pt) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t su = smem_u32(smem + stage * kStageBytes);
          const uint32_t sband = su + kUBytes;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            // K block k = image rows 2k, 2k+1 of the patch: U rows 16k..16k+15; band rows (2k)*10 .. , next row +1280 B
            const uint64_t da = make_smem_desc(su + k * 2048, lbo_a, 1024);
            const uint64_t db = make_smem_desc(sband + k * 2560, 128, 1280);
            umma_bf16_warp(d_tmem, da, db, idesc, (pt > pt0 || k > 0) ? 1u : 0u);
          }
          umma_commit_warp(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_warp(&tmem_full_bar[acc]);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp - 4;
    const int row = q * 32 + lane;
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      int cu_t, cs_t, r, split;
      decode(item, cu_t, cs_t, r, split);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const int mrow = BM_SLABS == 2 ? row : q * 16 + lane;          // accumulator row held by this TMEM lane
      const int cu_idx = cu_t * (BM_SLABS * 64) + mrow;
      const bool valid = (BM_SLABS == 2 || lane < 16) && cu_idx < p.cu;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t rg[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + c * 32, rg);
        tmem_ld_wait();
        if (valid) {
          const int s_tap = c >> 1;
          const bool det = p.partial != nullptr;
          float* dst = (det ? p.partial + (int64_t)split * p.out_elems : p.dw) + ((int64_t)cu_idx * 9 + (r * 3 + s_tap)) * p.cs +
                       cs_t * 64 + (c & 1) * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4) wgrad_out_v4(det, dst + j, rg[j], rg[j + 1], rg[j + 2], rg[j + 3]);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// =================================================================================================
// 3x3 weight gradient for a 64-channel gradient operand (cu == 64: the 256x256 layers), all nine taps per stage.
// An M = 64 MMA runs at half the tensor rate, so instead of one M=64 x N=192 MMA per kernel row (the kernel above) the
// second half of an M = 128 tile is filled with the SAME 64 channels one image row further down:
//     A = [ U(p) ; U(p + 1 row) ]   (MN-major descriptor with LBO = 8 pixels = 1024 B: "atom 1" is the row-shifted view)
//     D[0:64]   = sum_p U[p]      * S[p + (0, s-1)] = dw[r=1, s]
//     D[64:128] = sum_p U[p + 1r] * S[p + (0, s-1)] = sum_p' U[p'] * S[p' + (-1, s-1)] = dw[r=0, s]
// (the terms the shift drops or adds multiply zero padding, so the result is exact) and a second, M = 64 MMA on the
// view shifted one row UP gives dw[r=2].  All three kernel rows therefore share ONE band of S (16 x 10 pixels) and
// one 18 x 8 pixel patch of U per 128-pixel block: 38 KB into shared memory per block instead of 3 x 36 KB, and
// 192 instead of 288 tensor cycles per K step.  Accumulators: 192 + 192 TMEM columns, single buffered (the epilogue of
// an item -- a few microseconds of fp32 atomics -- no longer overlaps the next item; items last ~100x longer).
// Work item = (64-channel cs slab, pixel split).
// (The same sharing for 128-channel tiles -- two kernel rows per stage, 2 x 192 TMEM columns -- was measured 5-15 % SLOWER
// than tc_wgrad3x3_kernel<2>: with the accumulators single buffered the reduction epilogue of every item is exposed, and
// those layers have short items.  Not kept.)
// =================================================================================================
constexpr int kUPatchBytes = 18 * 8 * 128;   // 18432: rows h0-1 .. h0+16 of the 8-pixel-wide patch

__global__ void __launch_bounds__(kNumThreads, 1) tc_wgrad3x3_c64_kernel(const __grid_constant__ WgradParams p) {
  constexpr int kStageBytes = kUPatchBytes + kBandBytes;   // 38912
  constexpr int kStages = 5;
  constexpr int BLOCK_N = 192;
  constexpr uint32_t kTmemCols = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int num_items = p.cs_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.map_u);
    prefetch_tensormap(&p.map_s[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(tmem_empty_bar, 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_ptr_smem);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int split = item % p.splits, cs_t = item / p.splits;
        const int pt0 = split * p.ptiles_per_split;
        const int pt1 = min(pt0 + p.ptiles_per_split, p.num_ptiles);
        for (int pt = pt0; pt < pt1; ++pt) {
          const int tw = pt % p.tiles_w, th = (pt / p.tiles_w) % p.tiles_h, tn = pt / (p.tiles_w * p.tiles_h);
          const int w0 = tw * 8, h0 = th * 16;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* su = smem + stage * kStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
          tma_load_4d(su, &p.map_u, &full_bar[stage], 0, w0, h0 - 1, tn);                       // 18 x 8 patch of U
          tma_load_4d(su + kUPatchBytes, &p.map_s[0], &full_bar[stage], cs_t * 64, w0 - 1, h0, tn);   // band, r = 1
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {   // whole warp, converged (see umma_bf16_warp)
      constexpr uint32_t idesc128 = make_idesc_bf16(128, BLOCK_N, 1, 1);
      constexpr uint32_t idesc64 = make_idesc_bf16(64, BLOCK_N, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int split = item % p.splits;
        const int pt0 = split * p.ptiles_per_split;
        const int pt1 = min(pt0 + p.ptiles_per_split, p.num_ptiles);
        mbar_wait(tmem_empty_bar, (uint32_t)(it & 1) ^ 1);
        tcgen05_fence_after();
        for (int pt = pt0; pt < pt1; ++pt) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t su = smem_u32(smem + stage * kStageBytes);
          const uint32_t sband = su + kUPatchBytes;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            // K block k = image rows 2k, 2k+1 of the 16 x 8 block.  U patch row j = image row h0-1+j (8 pixels = 1024 B each)
            const uint64_t db = make_smem_desc(sband + k * 2560, 128, 1280);
            const uint64_t da_mid = make_smem_desc(su + 1024 + k * 2048, 1024, 1024);   // atoms: rows h0.. and h0+1..
            const uint64_t da_up = make_smem_desc(su + k * 2048, 0, 1024);              // rows h0-1..
            const uint32_t accum = (pt > pt0 || k > 0) ? 1u : 0u;
            umma_bf16_warp(tmem_base, da_mid, db, idesc128, accum);
            umma_bf16_warp(tmem_base + 256, da_up, db, idesc64, accum);
          }
          umma_commit_warp(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_warp(tmem_full_bar);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp - 4;
    const int row = q * 32 + lane;
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int cs_t = item / p.splits, split = item % p.splits;
      const bool det = p.partial != nullptr;
      float* const out_base = det ? p.partial + (int64_t)split * p.out_elems : p.dw;
      mbar_wait(tmem_full_bar, (uint32_t)(it & 1));
      tcgen05_fence_after();
      // accumulator 1 (M = 128): lanes 0..63 = dw[r=1] of channel `lane`, lanes 64..127 = dw[r=0] of channel lane-64
      {
        const int cu_idx = row & 63, r = row < 64 ? 1 : 0;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t rg[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, rg);
          tmem_ld_wait();
          float* dst = out_base + ((int64_t)cu_idx * 9 + (r * 3 + (c >> 1))) * p.cs + cs_t * 64 + (c & 1) * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4) wgrad_out_v4(det, dst + j, rg[j], rg[j + 1], rg[j + 2], rg[j + 3]);
        }
      }
      // accumulator 2 (M = 64): row m lives in TMEM lane (m/16)*32 + m%16 -> dw[r=2]
      {
        const int cu_idx = q * 16 + lane;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t rg[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + 256 + c * 32, rg);
          tmem_ld_wait();
          if (lane < 16) {
            float* dst = out_base + ((int64_t)cu_idx * 9 + (6 + (c >> 1))) * p.cs + cs_t * 64 + (c & 1) * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 4) wgrad_out_v4(det, dst + j, rg[j], rg[j + 1], rg[j + 2], rg[j + 3]);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

template <int BM_SLABS>
constexpr int wgrad3x3_smem_bytes() {
  constexpr int stage = BM_SLABS * kABytes + kBandBytes;
  constexpr int stages = (220 * 1024) / stage > 8 ? 8 : (220 * 1024) / stage;
  return stages * stage + (2 * stages + 4) * 8 + 16 + 1024;
}

template <typename K>
static int set_smem_attr(K kernel, int bytes) {
  UNETK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return UNETK_OK;
}

constexpr int kMaxSmem = 227 * 1024;

template <int BLOCK_N>
static int launch_conv(const ConvParams& p, int smem_bytes, bool halo, bool pair, bool even_groups, cudaStream_t stream) {
  static int attr_rc = set_smem_attr(tc_conv_kernel<BLOCK_N, false>, kMaxSmem);
  static int attr_rc1 = set_smem_attr(tc_conv_kernel<BLOCK_N, true>, kMaxSmem);
  static int attr_rc2 = set_smem_attr(tc_conv_halo_kernel<BLOCK_N>, kMaxSmem);
  static int attr_rc3 = set_smem_attr(tc_conv2_kernel<BLOCK_N, false>, kMaxSmem);
  static int attr_rc4 = set_smem_attr(tc_conv2_kernel<BLOCK_N, true>, kMaxSmem);
  static int attr_rc5 = set_smem_attr(tc_conv_halo2_kernel<BLOCK_N>, kMaxSmem);
  if (attr_rc5) return attr_rc5;
  if (attr_rc) return attr_rc;
  if (attr_rc1) return attr_rc1;
  if (attr_rc2) return attr_rc2;
  if (attr_rc3) return attr_rc3;
  if (attr_rc4) return attr_rc4;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  if (pair) {
    const int ptiles = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
    int pairs = ptiles < sm_count() / 2 ? ptiles : sm_count() / 2;
    // grouped schedule (PairSched) needs the pairs to split evenly over the N tiles: give up at most a few pairs for it
    // (3x3 halo kernels only: measured +5..16 % on the Cout = 1024 layers, -7 % on the short ConvT launches)
    if (even_groups && halo && p.num_n_tiles > 1 && pairs >= 4 * p.num_n_tiles && pairs % p.num_n_tiles != 0)
      pairs -= pairs % p.num_n_tiles;
    if (halo)
      tc_conv_halo2_kernel<BLOCK_N><<<2 * pairs, kConvThreads, smem_bytes, stream>>>(p);
    else if (p.bias)
      tc_conv2_kernel<BLOCK_N, true><<<2 * pairs, kConvThreads, smem_bytes, stream>>>(p);
    else
      tc_conv2_kernel<BLOCK_N, false><<<2 * pairs, kConvThreads, smem_bytes, stream>>>(p);
    UNETK_LAUNCH_CHECK();
    return UNETK_OK;
  }
  const int grid = tiles < sm_count() ? tiles : sm_count();
  if (halo)
    tc_conv_halo_kernel<BLOCK_N><<<grid, kConvThreads, smem_bytes, stream>>>(p);
  else if (p.bias)
    tc_conv_kernel<BLOCK_N, true><<<grid, kConvThreads, smem_bytes, stream>>>(p);
  else
    tc_conv_kernel<BLOCK_N, false><<<grid, kConvThreads, smem_bytes, stream>>>(p);
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

template <int BM_SLABS, int BN_SLABS>
static int launch_wgrad(const WgradParams& p, cudaStream_t stream) {
  using Cfg = WgradCfg<BM_SLABS, BN_SLABS>;
  static int attr_rc = set_smem_attr(tc_wgrad_kernel<BM_SLABS, BN_SLABS>, Cfg::kSmemBytes);
  if (attr_rc) return attr_rc;
  const int items = p.cu_tiles * p.cs_tiles * p.taps * p.splits;
  const int grid = items < sm_count() ? items : sm_count();
  tc_wgrad_kernel<BM_SLABS, BN_SLABS><<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(p);
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

template <int BM_SLABS>
static int launch_wgrad3x3(const WgradParams& p, cudaStream_t stream) {
  constexpr int smem_bytes = wgrad3x3_smem_bytes<BM_SLABS>();
  static int attr_rc = set_smem_attr(tc_wgrad3x3_kernel<BM_SLABS>, smem_bytes);
  if (attr_rc) return attr_rc;
  const int items = p.cu_tiles * p.cs_tiles * 3 * p.splits;
  const int grid = items < sm_count() ? items : sm_count();
  tc_wgrad3x3_kernel<BM_SLABS><<<grid, kNumThreads, smem_bytes, stream>>>(p);
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

static int launch_wgrad3x3_c64(const WgradParams& p, cudaStream_t stream) {
  constexpr int smem_bytes = 5 * (kUPatchBytes + kBandBytes) + 256 + 1024;
  static int attr_rc = set_smem_attr(tc_wgrad3x3_c64_kernel, smem_bytes);
  if (attr_rc) return attr_rc;
  const int items = p.cs_tiles * p.splits;
  const int grid = items < sm_count() ? items : sm_count();
  tc_wgrad3x3_c64_kernel<<<grid, kNumThreads, smem_bytes, stream>>>(p);
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

}  // namespace tc

}  // namespace unetk

#ifdef UNETK_DEBUG_COUNTERS
// internal debugging aid of instrumented builds only (not part of include/unetk.h): per-CTA idle-cycle counters of the
// last halo conv launch
extern "C" __attribute__((visibility("default"))) int unetk_debug_counters(long long* host_out, int n) {
  if (n > 160 * 8) n = 160 * 8;
  return cudaMemcpyFromSymbol(host_out, unetk::tc::g_dbg, sizeof(long long) * n) == cudaSuccess ? 0 : -2;
}
#endif

namespace unetk {

// =================================================================================================
// dispatcher-facing entry points
// =================================================================================================
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool tc_conv_supported(const unetk_conv_args* a, const ConvGeom& g, const char** why) {
  if (a->x.dtype != UNETK_BF16) { *why = "bf16 only"; return false; }
  if (g.cin % 64 != 0) { *why = "Cin must be a multiple of 64"; return false; }
  if (g.cout % 64 != 0) { *why = "Cout must be a multiple of 64"; return false; }
  if (a->x.ld % 8 != 0 || a->y.ld % 8 != 0 || !aligned16(a->x.ptr) || !aligned16(a->y.ptr) || !aligned16(a->w)) {
    *why = "pointers must be 16B aligned and pixel strides multiples of 8";
    return false;
  }
  if (a->stat_sum && g.cout > 4096) { *why = "statistics support at most 4096 channels"; return false; }
  if (a->bn_z.ptr && (a->bn_z.ld % 8 != 0 || (reinterpret_cast<uintptr_t>(a->bn_z.ptr) & 15) != 0)) {
    *why = "bn_z must be 16B aligned with a pixel stride multiple of 8";
    return false;
  }
  if (a->bn_z.ptr && a->stat_sum) { *why = "forward statistics and backward reduction are mutually exclusive"; return false; }
  return true;
}

// Experiment switches travel in the upper bits of unetk_conv_args.algo / unetk_wgrad_args.algo (UNETK_TC_* in unetk.h):
// the library reads no environment variables and keeps no mutable global state.

int tc_conv(const unetk_conv_args* a, const ConvGeom& g, cudaStream_t stream) {
  using namespace tc;
  ConvParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  // N tile: the widest of 256/128/64 that divides the channel count of one output slice
  // (mode 2: columns are (quadrant, channel); the epilogue picks the store map per 64-column block, so a tile may span
  // quadrants and ConvT 128->64 runs as ONE N = 256 tile instead of four N = 64 tiles that each re-read the input)
  const int nslice = g.cout_total;
  const int block_n = nslice % 256 == 0 ? 256 : (nslice % 128 == 0 ? 128 : 64);
  const int b_bytes = block_n * kBlockK * 2;
  const int64_t ktotal = (int64_t)g.taps * g.cin;
  const int cpt = g.cin / 64;
  // halo variant: 3x3, N tile <= 128 (the layers whose operand traffic is L2-bound), image at least one tile big
  // halo reuse also for N = 256 (CTA pairs): measured 5-9% faster on the deep layers once the MMA issue path was fixed
  // (the per-tap pair kernel is then bound by L2->SMEM operand traffic, which the halo tile cuts by ~40%)
  const int flags = a->algo & ~UNETK_ALGO_MASK;
  const bool halo_n256 = !(flags & UNETK_TC_NO_HALO_N256);
  const bool halo = a->mode == 1 && (block_n <= 128 || halo_n256) && g.rows_h >= 16 && g.rows_w >= 8 && !(flags & UNETK_TC_NO_HALO);
  PixelTile pt;
  if (halo) {
    pt.pw = 8; pt.ph = 16; pt.nb = 1;
    pt.tiles_w = (g.rows_w + 7) / 8;
    pt.tiles_h = (g.rows_h + 15) / 16;
    pt.tiles_n = g.rows_n;
    if ((rc = make_act_map(&p.map_a[0], a->x, 10, 18, 1, 1, 0, 0))) return rc;
  } else {
    pt = choose_pixel_tile(g.rows_n, g.rows_h, g.rows_w);
    if (a->mode == 3) {
      for (int t = 0; t < 4; ++t)
        if ((rc = make_act_map(&p.map_a[t], a->x, pt.pw, pt.ph, pt.nb, 2, t >> 1, t & 1))) return rc;
    } else {
      if ((rc = make_act_map(&p.map_a[0], a->x, pt.pw, pt.ph, pt.nb, 1, 0, 0))) return rc;
    }
  }
  if (a->mode == 2) {
    for (int t = 0; t < 4; ++t)
      if ((rc = make_act_map(&p.map_y[t], a->y, pt.pw, pt.ph, pt.nb, 2, t >> 1, t & 1))) return rc;
  } else {
    if ((rc = make_act_map(&p.map_y[0], a->y, pt.pw, pt.ph, pt.nb, 1, 0, 0))) return rc;
  }
  // CTA pairs (cta_group::2) whenever there are at least two pixel tiles: each CTA then loads only half of the weight rows
  // per K step.  Measured on B200 (batch 64) with the warp-uniform MMA issue: pairs gain 4-10% on the per-tap kernel and
  // 8-17% on the halo kernel (N <= 128).  (Before the issue path was fixed the halo kernel LOST 10-30% with pairs: the
  // single-lane waterfall issue was the bottleneck and the leader CTA had to issue for both.)  UNETK_TC_NO_HALO_PAIR and
  // UNETK_TC_NO_PAIR (algo flags) keep the single-CTA kernels reachable for experiments.
  const bool halo_pair = !(flags & UNETK_TC_NO_HALO_PAIR);
  const bool pair = !(flags & UNETK_TC_NO_PAIR) && pt.num_tiles() >= 2 && (!halo || halo_pair || (halo_n256 && block_n == 256));
  const bool even_groups = !(flags & UNETK_TC_NO_EVEN_GROUPS);
  if ((rc = make_mat_map(&p.map_b, a->w, g.cout_total, ktotal, pair ? block_n / 2 : block_n))) return rc;
  p.mode = a->mode;
  p.taps = g.taps;
  p.chunks_per_tap = cpt;
  p.pw = pt.pw; p.ph = pt.ph; p.nb = pt.nb; p.tiles_w = pt.tiles_w; p.tiles_h = pt.tiles_h;
  UNETK_REQUIRE(pt.num_tiles() * (g.cout_total / block_n) < (1LL << 31), "conv(tc): too many tiles");
  p.num_m_tiles = (int)pt.num_tiles();
  p.num_n_tiles = g.cout_total / block_n;
  p.n = g.rows_n; p.h = g.rows_h; p.w = g.rows_w;
  p.cout = g.cout;
  p.bias = a->bias;
  p.stat_sum = a->stat_sum;
  p.stat_sumsq = a->stat_sumsq;
  p.stat_channels = 0;
  const bool bnred = a->bn_z.ptr != nullptr;
  if (bnred) {
    if ((rc = make_act_map(&p.map_z, a->bn_z, pt.pw, pt.ph, pt.nb, 1, 0, 0))) return rc;
    p.bn_scale = a->bn_scale; p.bn_shift = a->bn_shift; p.bn_mean = a->bn_mean; p.bn_invstd = a->bn_invstd;
    p.bn_sums = a->bn_sums;
  }
  // ---- shared-memory plan: operand ring(s) | 2 staging blocks (one per epilogue group) [| 2 z blocks] | barriers ----
  const int budget = kMaxSmem - 1024 /*alignment slack*/ - kBarBytes - (bnred ? 4 : 2) * kStagingBytes;
  p.num_staging = 2;
  int ring_bytes;
  if (!halo) {
    const int stage_bytes = kABytes + (pair ? b_bytes / 2 : b_bytes);
    p.stages = budget / stage_bytes;
    if (p.stages > kMaxAStages) p.stages = kMaxAStages;
    ring_bytes = p.stages * stage_bytes;
    p.off_b = 0;
  } else {
    const int ksteps = 9 * cpt;
    const int bb = pair ? b_bytes / 2 : b_bytes;   // a CTA of a pair holds half of the weight rows
    p.resident_b = (p.num_n_tiles == 1 && ksteps <= kMaxBSlots && ksteps * bb + 2 * kHaloStride <= budget) ? 1 : 0;
    if (p.resident_b) {
      p.b_slots = ksteps;
      p.stages = (budget - ksteps * bb) / kHaloStride;
      if (p.stages > 4) p.stages = 4;
    } else {
      p.stages = 3;
      p.b_slots = (budget - p.stages * kHaloStride) / bb;
      if (p.b_slots > 9) p.b_slots = 9;
      UNETK_REQUIRE(p.b_slots >= 2, "conv(tc halo): shared-memory plan failed");
    }
    p.off_b = p.stages * kHaloStride;
    ring_bytes = p.off_b + p.b_slots * bb;
  }
  p.off_staging = ring_bytes;
  p.off_stats = 0;
  p.off_zbuf = p.off_staging + 2 * kStagingBytes;
  p.off_bars = p.off_zbuf + (bnred ? 2 : 0) * kStagingBytes;
  const int smem_bytes = p.off_bars + kBarBytes + 1024;
  UNETK_REQUIRE(smem_bytes <= kMaxSmem && p.stages >= 2, "conv(tc): shared-memory plan failed (%d bytes, %d stages)", smem_bytes, p.stages);
  if (block_n == 256) return launch_conv<256>(p, smem_bytes, halo, pair, even_groups, stream);
  if (block_n == 128) return launch_conv<128>(p, smem_bytes, halo, pair, even_groups, stream);
  return launch_conv<64>(p, smem_bytes, halo, pair, even_groups, stream);
}

bool tc_wgrad_supported(const unetk_wgrad_args* a, int taps, const char** why) {
  (void)taps;
  if (a->u.dtype != UNETK_BF16) { *why = "bf16 only"; return false; }
  if (a->u.c % 64 != 0 || a->s.c % 64 != 0) { *why = "channel counts must be multiples of 64"; return false; }
  if (a->u.ld % 8 != 0 || a->s.ld % 8 != 0 || !aligned16(a->u.ptr) || !aligned16(a->s.ptr)) {
    *why = "pointers must be 16B aligned and pixel strides multiples of 8";
    return false;
  }
  return true;
}

// Split the pixel reduction of the weight gradient so that (output tiles x splits) fills the persistent grid in whole
// waves: the largest split count with at most 2 work items per SM (never rounding UP past a wave boundary, which would
// leave most SMs idle for a third round), at least 4 pixel tiles per item.
static void choose_splits(int64_t out_tiles, int num_ptiles, int* ptiles_per_split, int* splits_out) {
  // pick the split count whose work-item total fills the persistent grid best: efficiency = items / (rounds * SMs);
  // near-ties go to fewer splits (fewer fp32 atomics, longer K loops); >= 4 pixel tiles per item, <= ~3 items per SM
  const int64_t sms = sm_count();
  int64_t max_splits = (num_ptiles + 3) / 4;
  if (max_splits < 1) max_splits = 1;
  int64_t cap = (3 * sms + out_tiles - 1) / out_tiles;
  if (cap < 1) cap = 1;
  if (max_splits > cap) max_splits = cap;
  int best = 1;
  double best_eff = -1.0;
  for (int64_t sp = 1; sp <= max_splits; ++sp) {
    const int per = (int)((num_ptiles + sp - 1) / sp);
    const int64_t real = (num_ptiles + per - 1) / per;          // splits actually produced
    const int64_t items = out_tiles * real;
    const int64_t rounds = (items + sms - 1) / sms;
    // work per item is `per` pixel tiles; the grid finishes after rounds * per tile-times
    const double eff = (double)out_tiles * num_ptiles / ((double)rounds * sms * per);
    if (eff > best_eff + 0.03) {   // more splits only for a clear (> 3 %) gain in grid fill
      best_eff = eff;
      best = (int)real;
    }
  }
  int per = (num_ptiles + best - 1) / best;
  *ptiles_per_split = per;
  *splits_out = (num_ptiles + per - 1) / per;
}

// Number of pixel splits the dispatcher below will choose for this problem (and whether it takes the 3x3 halo form).
static int wgrad_splits(const unetk_wgrad_args* a, int taps) {
  using namespace tc;
  int per = 0, splits = 1;
  if (a->mode == 1 && a->u.h >= 16 && a->u.w >= 8 && !(a->algo & UNETK_TC_NO_HALO)) {
    const bool c64 = a->u.c == 64 && !(a->algo & UNETK_TC_NO_WGRAD_C64);
    const int bm_slabs = a->u.c % 128 == 0 ? 2 : 1;
    const int64_t ptiles = (int64_t)((a->u.w + 7) / 8) * ((a->u.h + 15) / 16) * a->u.n;
    const int64_t out_tiles = c64 ? (int64_t)(a->s.c / 64) : (int64_t)(a->u.c / (64 * bm_slabs)) * (a->s.c / 64) * 3;
    choose_splits(out_tiles, (int)ptiles, &per, &splits);
  } else {
    const PixelTile pt = choose_pixel_tile(a->u.n, a->u.h, a->u.w);
    const int bm_slabs = a->u.c % 128 == 0 ? 2 : 1, bn_slabs = a->s.c % 128 == 0 ? 2 : 1;
    const int64_t out_tiles = (int64_t)(a->u.c / (64 * bm_slabs)) * (a->s.c / (64 * bn_slabs)) * taps;
    choose_splits(out_tiles, (int)pt.num_tiles(), &per, &splits);
  }
  return splits;
}

int64_t tc_wgrad_partial_bytes(const unetk_wgrad_args* a, int taps) {
  const int splits = wgrad_splits(a, taps);
  return splits > 1 ? (int64_t)splits * a->u.c * taps * a->s.c * (int64_t)sizeof(float) : 0;
}

// deterministic mode: route the epilogues to the caller's partial buffer; returns 1 when active
static int setup_deterministic(const unetk_wgrad_args* a, int taps, tc::WgradParams& p) {
  p.out_elems = (int64_t)a->u.c * taps * a->s.c;
  p.partial = nullptr;
  if (!(a->algo & UNETK_TC_DETERMINISTIC) || p.splits <= 1) return 0;
  const int64_t need = (int64_t)p.splits * p.out_elems * (int64_t)sizeof(float);
  if (!a->partial || a->partial_bytes < need) {
    set_error("wgrad(tc): UNETK_TC_DETERMINISTIC needs a partial buffer of %lld bytes (unetk_wgrad_partial_bytes), got %lld",
              (long long)need, (long long)(a->partial ? a->partial_bytes : 0));
    return -1;
  }
  p.partial = a->partial;
  return 1;
}

static int finish_deterministic(const tc::WgradParams& p, cudaStream_t stream) {
  const int64_t n4 = p.out_elems / 4;
  int64_t blocks = (n4 + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  tc::wgrad_reduce_kernel<<<(unsigned)(blocks > 0 ? blocks : 1), 256, 0, stream>>>(p.partial, p.splits, p.out_elems, p.dw);
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

static int tc_wgrad3x3_halo(const unetk_wgrad_args* a, cudaStream_t stream) {
  using namespace tc;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  const bool c64 = a->u.c == 64 && !(a->algo & UNETK_TC_NO_WGRAD_C64);   // nine taps per stage (tc_wgrad3x3_c64_kernel): U patch with a row of halo
  if ((rc = make_act_map(&p.map_u, a->u, 8, c64 ? 18 : 16, 1, 1, 0, 0))) return rc;
  if ((rc = make_act_map(&p.map_s[0], a->s, 10, 16, 1, 1, 0, 0))) return rc;
  const int bm_slabs = a->u.c % 128 == 0 ? 2 : 1;
  p.mode = 1;
  p.taps = 9;
  p.pw = 8; p.ph = 16; p.nb = 1;
  p.tiles_w = (a->u.w + 7) / 8;
  p.tiles_h = (a->u.h + 15) / 16;
  const int64_t ptiles = (int64_t)p.tiles_w * p.tiles_h * a->u.n;
  UNETK_REQUIRE(ptiles < (1LL << 30), "wgrad(tc): too many pixel tiles");
  p.num_ptiles = (int)ptiles;
  p.cu = a->u.c; p.cs = a->s.c;
  p.cu_tiles = a->u.c / (64 * bm_slabs);
  p.cs_tiles = a->s.c / 64;
  const int64_t out_tiles = c64 ? (int64_t)p.cs_tiles : (int64_t)p.cu_tiles * p.cs_tiles * 3;
  choose_splits(out_tiles, p.num_ptiles, &p.ptiles_per_split, &p.splits);
  UNETK_REQUIRE(out_tiles * p.splits < (1LL << 31), "wgrad(tc): too many work items");
  p.dw = a->dw;
  const int det = setup_deterministic(a, 9, p);
  if (det < 0) return UNETK_ERR_INVALID;
  rc = c64 ? launch_wgrad3x3_c64(p, stream) : (bm_slabs == 2 ? launch_wgrad3x3<2>(p, stream) : launch_wgrad3x3<1>(p, stream));
  if (rc || !det) return rc;
  return finish_deterministic(p, stream);
}

int tc_wgrad(const unetk_wgrad_args* a, int taps, cudaStream_t stream) {
  using namespace tc;
  if (a->mode == 1 && a->u.h >= 16 && a->u.w >= 8 && !(a->algo & UNETK_TC_NO_HALO)) return tc_wgrad3x3_halo(a, stream);
  WgradParams p;
  memset(&p, 0, sizeof(p));
  const PixelTile pt = choose_pixel_tile(a->u.n, a->u.h, a->u.w);
  int rc;
  if ((rc = make_act_map(&p.map_u, a->u, pt.pw, pt.ph, pt.nb, 1, 0, 0))) return rc;
  if (a->mode == 2) {
    for (int t = 0; t < 4; ++t)
      if ((rc = make_act_map(&p.map_s[t], a->s, pt.pw, pt.ph, pt.nb, 2, t >> 1, t & 1))) return rc;
  } else {
    if ((rc = make_act_map(&p.map_s[0], a->s, pt.pw, pt.ph, pt.nb, 1, 0, 0))) return rc;
  }
  const int bm_slabs = a->u.c % 128 == 0 ? 2 : 1;
  const int bn_slabs = a->s.c % 128 == 0 ? 2 : 1;
  p.mode = a->mode;
  p.taps = taps;
  p.pw = pt.pw; p.ph = pt.ph; p.nb = pt.nb; p.tiles_w = pt.tiles_w; p.tiles_h = pt.tiles_h;
  UNETK_REQUIRE(pt.num_tiles() < (1LL << 30), "wgrad(tc): too many pixel tiles");
  p.num_ptiles = (int)pt.num_tiles();
  p.cu = a->u.c; p.cs = a->s.c;
  p.cu_tiles = a->u.c / (64 * bm_slabs);
  p.cs_tiles = a->s.c / (64 * bn_slabs);
  const int64_t out_tiles = (int64_t)p.cu_tiles * p.cs_tiles * taps;
  choose_splits(out_tiles, p.num_ptiles, &p.ptiles_per_split, &p.splits);
  UNETK_REQUIRE(out_tiles * p.splits < (1LL << 31), "wgrad(tc): too many work items");
  p.dw = a->dw;
  const int det = setup_deterministic(a, taps, p);
  if (det < 0) return UNETK_ERR_INVALID;
  if (bm_slabs == 2 && bn_slabs == 2) rc = launch_wgrad<2, 2>(p, stream);
  else if (bm_slabs == 2 && bn_slabs == 1) rc = launch_wgrad<2, 1>(p, stream);
  else if (bm_slabs == 1 && bn_slabs == 2) rc = launch_wgrad<1, 2>(p, stream);
  else rc = launch_wgrad<1, 1>(p, stream);
  if (rc || !det) return rc;
  return finish_deterministic(p, stream);
}

}  // namespace unetk
