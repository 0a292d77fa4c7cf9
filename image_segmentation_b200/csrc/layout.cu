// Layout kernels: NCHW fp32 input -> NHWC im2col operand of the first conv; weight / weight-gradient
// permutations between PyTorch's OIHW / IOHW fp32 parameters and the K-major operand packs.
#include "common.cuh"

namespace unetk {

// out[n,h,w,k] = x[n, k % cin, h + (k/cin)/3 - 1, w + (k/cin)%3 - 1]  for k < 9*cin, else 0
template <typename T>
__global__ void __launch_bounds__(256) im2col_first_kernel(const float* __restrict__ x, int n, int cin, int h, int w,
                                                           T* __restrict__ out, int kpad, int ld) {
  const int kg = kpad / 8;
  const int64_t total = (int64_t)n * h * w * kg;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % kg);
    int64_t p = i / kg;
    const int px = (int)(p % w);
    const int py = (int)((p / w) % h);
    const int img = (int)(p / ((int64_t)w * h));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = g * 8 + j;
      float val = 0.f;
      if (k < 9 * cin) {
        const int t = k / cin, ci = k % cin;
        const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) val = __ldg(x + (((int64_t)img * cin + ci) * h + yy) * w + xx);
      }
      v[j] = val;
    }
    store8(out + p * ld + g * 8, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) permute3_kernel(const float* __restrict__ src, T* __restrict__ dst, int d0, int d1,
                                                       int d2, int64_t ss0, int64_t ss1, int64_t ss2, int64_t ds0,
                                                       int64_t ds1, int64_t ds2) {
  // one thread per (i0, i1); the short dim i2 (taps) is looped
  const int64_t total = (int64_t)d0 * d1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i0 = i / d1, i1 = i % d1;
    const float* s = src + i0 * ss0 + i1 * ss1;
    T* d = dst + i0 * ds0 + i1 * ds1;
    for (int i2 = 0; i2 < d2; ++i2) d[i2 * ds2] = from_f<T>(s[i2 * ss2]);
  }
}


// ------------------------------------------------------------------------------------------------
// batched weight pack / unpack: one block per 32x32 channel tile, transposition through shared memory so that both
// the parameter side and the operand-pack side are accessed in contiguous runs
// ------------------------------------------------------------------------------------------------
constexpr int kWT = 32;

__device__ __forceinline__ void store2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void store2(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(a, b);
}

// One 32 x 32 channel tile of a 3x3 conv (TAPS = 9) or 2x2 ConvT (TAPS = 4) weight.  Every loop has a compile-time trip
// count and compile-time divisors, so the loads of a phase are all in flight before the first use; the parameter side moves
// as float4 when its tile rows are 16-byte aligned.
template <typename T, bool PACK, int TAPS>
__device__ __forceinline__ void weights_tile(const unetk_wjob& j, int a0, int b0, int tid, float (&sm)[kWT][kWT * 9 + 1]) {
  constexpr int RUN = kWT * TAPS;                   // contiguous floats per `a` row of the parameter tile
  constexpr int kVecIters = kWT * RUN / 4 / 256;    // 9 (3x3) / 4 (ConvT)
  static_assert(kWT * RUN % (4 * 256) == 0, "tile must split into whole float4 rounds");
  // parameter layout: [a][b][taps] with a = co (conv) / ci (convT);  A = rows, B = columns
  const int B = TAPS == 9 ? j.cin : j.cout;
  if (PACK) {
    const float* param = reinterpret_cast<const float*>(j.src);
    if ((reinterpret_cast<uintptr_t>(param) & 15) == 0 && ((size_t)B * TAPS) % 4 == 0) {
      float4 v[kVecIters];
#pragma unroll
      for (int it = 0; it < kVecIters; ++it) {
        const int i = tid + it * 256, al = i / (RUN / 4), e4 = i % (RUN / 4);
        v[it] = *reinterpret_cast<const float4*>(param + ((size_t)(a0 + al) * B + b0) * TAPS + 4 * e4);
      }
#pragma unroll
      for (int it = 0; it < kVecIters; ++it) {
        const int i = tid + it * 256, al = i / (RUN / 4), e4 = i % (RUN / 4);
        sm[al][4 * e4] = v[it].x;
        sm[al][4 * e4 + 1] = v[it].y;
        sm[al][4 * e4 + 2] = v[it].z;
        sm[al][4 * e4 + 3] = v[it].w;
      }
    } else {
#pragma unroll 12
      for (int i = tid; i < kWT * RUN; i += 256) {
        const int al = i / RUN, e = i % RUN;
        sm[al][e] = param[((size_t)(a0 + al) * B + b0) * TAPS + e];
      }
    }
    __syncthreads();
    T* wf = reinterpret_cast<T*>(j.dst0);
    T* wd = reinterpret_cast<T*>(j.dst1);
    // two adjacent elements per store (4 bytes of bf16 / 8 bytes of fp32)
#pragma unroll
    for (int i = tid; i < kWT * TAPS * (kWT / 2); i += 256) {
      const int cp = i % (kWT / 2), t = (i / (kWT / 2)) % TAPS, row = i / (TAPS * (kWT / 2));
      if (TAPS == 9) {
        // wf[co][t][ci]: consecutive threads -> consecutive ci pairs
        store2(wf + ((size_t)(a0 + row) * 9 + t) * j.cin + b0 + 2 * cp, sm[row][(2 * cp) * 9 + t], sm[row][(2 * cp + 1) * 9 + t]);
        // wd[ci][8-t][co]: consecutive threads -> consecutive co pairs
        store2(wd + ((size_t)(b0 + row) * 9 + (8 - t)) * j.cout + a0 + 2 * cp, sm[2 * cp][row * 9 + t], sm[2 * cp + 1][row * 9 + t]);
      } else {
        // convT: sm[ci][co*4 + ab];  wf[(ab*cout + co)][ci]: consecutive threads -> consecutive ci pairs
        store2(wf + ((size_t)t * j.cout + b0 + row) * j.cin + a0 + 2 * cp, sm[2 * cp][row * 4 + t], sm[2 * cp + 1][row * 4 + t]);
        // wd[ci][ab][co]: consecutive threads -> consecutive co pairs
        store2(wd + ((size_t)(a0 + row) * 4 + t) * j.cout + b0 + 2 * cp, sm[row][(2 * cp) * 4 + t], sm[row][(2 * cp + 1) * 4 + t]);
      }
    }
  } else {
    const float* ws = reinterpret_cast<const float*>(j.src);
    // packed gradient: ws[co][t][ci] (conv) / ws[ci][ab][co] (ConvT): 32 contiguous floats per (row, tap)
    const int pitch = TAPS == 9 ? j.cin : j.cout;
    if ((reinterpret_cast<uintptr_t>(ws) & 15) == 0 && pitch % 4 == 0) {
      float4 v[kVecIters];
#pragma unroll
      for (int it = 0; it < kVecIters; ++it) {
        const int i = tid + it * 256, c4 = i % (kWT / 4), t = (i / (kWT / 4)) % TAPS, row = i / (TAPS * (kWT / 4));
        v[it] = *reinterpret_cast<const float4*>(ws + ((size_t)(a0 + row) * TAPS + t) * pitch + b0 + 4 * c4);
      }
#pragma unroll
      for (int it = 0; it < kVecIters; ++it) {
        const int i = tid + it * 256, c4 = i % (kWT / 4), t = (i / (kWT / 4)) % TAPS, row = i / (TAPS * (kWT / 4));
        sm[row][(4 * c4) * TAPS + t] = v[it].x;
        sm[row][(4 * c4 + 1) * TAPS + t] = v[it].y;
        sm[row][(4 * c4 + 2) * TAPS + t] = v[it].z;
        sm[row][(4 * c4 + 3) * TAPS + t] = v[it].w;
      }
    } else {
#pragma unroll 12
      for (int i = tid; i < kWT * RUN; i += 256) {
        const int c = i % kWT, t = (i / kWT) % TAPS, row = i / RUN;
        sm[row][c * TAPS + t] = ws[((size_t)(a0 + row) * TAPS + t) * pitch + b0 + c];
      }
    }
    __syncthreads();
    float* grad = reinterpret_cast<float*>(j.dst0);
    if ((reinterpret_cast<uintptr_t>(grad) & 15) == 0 && ((size_t)B * TAPS) % 4 == 0) {
#pragma unroll
      for (int it = 0; it < kVecIters; ++it) {
        const int i = tid + it * 256, al = i / (RUN / 4), e4 = i % (RUN / 4);
        *reinterpret_cast<float4*>(grad + ((size_t)(a0 + al) * B + b0) * TAPS + 4 * e4) =
            make_float4(sm[al][4 * e4], sm[al][4 * e4 + 1], sm[al][4 * e4 + 2], sm[al][4 * e4 + 3]);
      }
    } else {
#pragma unroll 12
      for (int i = tid; i < kWT * RUN; i += 256) {
        const int al = i / RUN, e = i % RUN;
        grad[((size_t)(a0 + al) * B + b0) * TAPS + e] = sm[al][e];
      }
    }
  }
}

template <typename T, bool PACK>
__global__ void __launch_bounds__(256) weights_kernel(const unetk_wjob* __restrict__ jobs, const int32_t* __restrict__ tiles,
                                                      uint8_t* dst_base) {
  __shared__ float sm[kWT][kWT * 9 + 1];
  const int4 tl = reinterpret_cast<const int4*>(tiles)[blockIdx.x];
  unetk_wjob j = jobs[tl.x];
  if (dst_base) j.dst0 = dst_base + reinterpret_cast<uintptr_t>(j.dst0);
  const int a0 = tl.y, b0 = tl.z;
  const int tid = threadIdx.x;
  if (j.kind == 3) {
    // first layer in pixel-pair form (see engine.py): the GEMM row is a PAIR of horizontally adjacent pixels, K = 2 x kpad/2
    // (each pixel's padded 3x3xCin patch) and N = 2 x cout, with the block-diagonal weight  [[W 0] [0 W]]  of pitch kpad.
    // pack: W -> both diagonal blocks; unpack: the gradient is the sum of the two diagonal blocks of dW'.
    const int total = j.cout * j.cin * 9, half = j.kpad / 2;
    for (int i = tid; i < total; i += 256) {
      const int co = i / (j.cin * 9), r = i % (j.cin * 9), ci = r / 9, t = r % 9, k = t * j.cin + ci;
      if (PACK) {
        const T v = from_f<T>(reinterpret_cast<const float*>(j.src)[i]);
        reinterpret_cast<T*>(j.dst0)[(size_t)co * j.kpad + k] = v;
        reinterpret_cast<T*>(j.dst0)[(size_t)(j.cout + co) * j.kpad + half + k] = v;
      } else {
        const float* ws2 = reinterpret_cast<const float*>(j.src);
        reinterpret_cast<float*>(j.dst0)[i] = ws2[(size_t)co * j.kpad + k] + ws2[(size_t)(j.cout + co) * j.kpad + half + k];
      }
    }
    return;
  }
  if (j.kind == 2) {
    // tiny first layer: [co][ci][9] <-> [co][kpad], k = t*cin + ci
    const int total = j.cout * j.cin * 9;
    for (int i = tid; i < total; i += 256) {
      const int co = i / (j.cin * 9), r = i % (j.cin * 9), ci = r / 9, t = r % 9;
      if (PACK)
        reinterpret_cast<T*>(j.dst0)[(size_t)co * j.kpad + t * j.cin + ci] = from_f<T>(reinterpret_cast<const float*>(j.src)[i]);
      else
        reinterpret_cast<float*>(j.dst0)[i] = reinterpret_cast<const float*>(j.src)[(size_t)co * j.kpad + t * j.cin + ci];
    }
    return;
  }
  if (j.kind == 0) weights_tile<T, PACK, 9>(j, a0, b0, tid, sm); else weights_tile<T, PACK, 4>(j, a0, b0, tid, sm);
}

// NCHW fp32 image -> NHWC im2col operand, 8 consecutive K values per thread with the (tap, channel) decoding hoisted
// out of the pixel loop (thread's K group is fixed: the grid stride is a multiple of kg)
template <typename T>
__global__ void __launch_bounds__(256) im2col_first_kernel_v2(const float* __restrict__ x, int n, int cin, int h, int w,
                                                              T* __restrict__ out, int kpad, int ld) {
  const int kg = kpad / 8;                       // K groups per pixel
  const int g = threadIdx.x % kg;
  const int pix_lane = threadIdx.x / kg, lanes = blockDim.x / kg;
  int dy[8], dx[8], ci[8];
  bool kv[8];
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int k = g * 8 + jj;
    kv[jj] = k < 9 * cin;
    const int t = kv[jj] ? k / cin : 0;
    ci[jj] = kv[jj] ? k % cin : 0;
    dy[jj] = t / 3 - 1;
    dx[jj] = t % 3 - 1;
  }
  const int64_t npix = (int64_t)n * h * w;
  const int64_t hw = (int64_t)h * w;
  for (int64_t p = (int64_t)blockIdx.x * lanes + pix_lane; p < npix; p += (int64_t)gridDim.x * lanes) {
    const int px = (int)(p % w);
    const int py = (int)((p / w) % h);
    const int64_t img = p / hw;
    const float* xi = x + img * cin * hw;
    float v[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int yy = py + dy[jj], xx = px + dx[jj];
      const bool ok = kv[jj] && yy >= 0 && yy < h && xx >= 0 && xx < w;
      v[jj] = ok ? __ldg(xi + ci[jj] * hw + (int64_t)yy * w + xx) : 0.f;
    }
    store8(out + p * ld + g * 8, v);
  }
}

// Tiled variant: a block stages the (cin x 3 rows x TW+2 columns) fp32 input patch of a 1 x TW pixel strip in shared
// memory with coalesced row reads (zeros outside the image), then every thread assembles 8 consecutive K values of a
// pixel from it and writes 16 bytes; consecutive threads write consecutive 16-byte chunks.  The gather kernel above
// needed ~12 L1 wavefronts per load instruction and ran at 1.3 TB/s; this one is bound by the im2col write.
constexpr int kStripW = 64;   // pixels per strip row
constexpr int kStripH = 4;    // image rows per strip: every input row is staged once per 4 output rows
constexpr int kMaxFirstCin = 8;
template <typename T>
__global__ void __launch_bounds__(256) im2col_first_kernel_v3(const float* __restrict__ x, int n, int cin, int h, int w,
                                                              T* __restrict__ out, int kpad, int ld, int greal) {
  // tile rows: (ci, r) -> kStripW+2 columns, r = 0 .. kStripH+1; a trailing block of zeros serves the padding K values
  constexpr int kPitch = kStripW + 2, kRows = kStripH + 2;
  __shared__ float tile[kMaxFirstCin * kRows * kPitch + kRows * kPitch];
  const int kg = kpad / 8;
  // thread = (pixel of the strip row, REAL channel group gi < greal = ceil(9*cin/8)); it also writes the all-zero groups
  // gi + greal, gi + 2*greal, ... so that no warp diverges between "gather" and "zero fill" lanes
  const int gi = threadIdx.x % greal, pix_lane = threadIdx.x / greal, lanes = 256 / greal;
  const int tile_elems = cin * kRows * kPitch;
  const int zero_base = tile_elems;
  for (int e = threadIdx.x; e < kRows * kPitch; e += 256) tile[zero_base + e] = 0.f;
  int off[8];
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int k = gi * 8 + jj;
    if (k < 9 * cin) {
      const int t = k / cin, ci = k % cin;
      off[jj] = (ci * kRows + t / 3) * kPitch + t % 3;
    } else {
      off[jj] = zero_base;
    }
  }
  const int strips_w = (w + kStripW - 1) / kStripW, strips_h = (h + kStripH - 1) / kStripH;
  const int64_t strips = (int64_t)n * strips_h * strips_w;
  const int hw = h * w;
  const float zeros[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t sidx = blockIdx.x; sidx < strips; sidx += gridDim.x) {
    const int sx = (int)(sidx % strips_w);
    const int sy = (int)((sidx / strips_w) % strips_h);
    const int64_t img = sidx / ((int64_t)strips_w * strips_h);
    const int x0 = sx * kStripW, y0 = sy * kStripH;
    const float* xi = x + img * cin * (int64_t)hw;
    __syncthreads();   // previous strip fully consumed (and the zero block written, first time round)
    for (int e = threadIdx.x; e < tile_elems; e += 256) {
      const int col = e % kPitch, rc = e / kPitch;
      const int r = rc % kRows, ci = rc / kRows;
      const int yy = y0 + r - 1, xx = x0 + col - 1;
      tile[e] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(xi + ci * hw + yy * w + xx) : 0.f;
    }
    __syncthreads();
    const int npl = min(kStripW, w - x0), nrow = min(kStripH, h - y0);
    for (int r = 0; r < nrow; ++r) {
      T* orow = out + ((img * h + y0 + r) * (int64_t)w + x0) * ld;
      for (int pl = pix_lane; pl < npl; pl += lanes) {
        float v[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) v[jj] = tile[off[jj] + r * kPitch + pl];
        T* o = orow + (int64_t)pl * ld;
        store8(o + gi * 8, v);
        for (int gz = gi + greal; gz < kg; gz += greal) store8(o + gz * 8, zeros);
      }
    }
  }
}

}  // namespace unetk

using namespace unetk;

extern "C" {

int unetk_im2col3x3_first(const float* x_nchw, int32_t n, int32_t cin, int32_t h, int32_t w, const unetk_tensor* out,
                          void* stream) {
  UNETK_REQUIRE(x_nchw && out, "im2col_first: null argument");
  UNETK_REQUIRE(tensor_ok(*out) && vec8_ok(*out), "im2col_first: out must be NHWC with c%%8==0, ld%%8==0");
  UNETK_REQUIRE(out->n == n && out->h == h && out->w == w && out->c >= 9 * cin && cin > 0,
                "im2col_first: out must be [N,H,W,K>=9*Cin]");
  const int64_t items = (int64_t)n * h * w * (out->c / 8);
  int64_t blocks = (items + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  const int kg = out->c / 8;
  const int greal = (9 * cin + 7) / 8;   // channel groups that hold real K values
  if (256 % greal == 0 && cin <= kMaxFirstCin && (int64_t)cin * h * w < (1LL << 31)) {
    int64_t strips = (int64_t)n * ((h + kStripH - 1) / kStripH) * ((w + kStripW - 1) / kStripW);
    if (strips > cap) strips = cap;
    UNETK_DISPATCH_DTYPE(out->dtype, T, {
      im2col_first_kernel_v3<T><<<(int)strips, 256, 0, (cudaStream_t)stream>>>(x_nchw, n, cin, h, w, (T*)out->ptr, out->c, out->ld,
                                                                              greal);
    });
  } else if (256 % kg == 0) {
    const int lanes = 256 / kg;
    int64_t b2 = ((int64_t)n * h * w + lanes - 1) / lanes;
    if (b2 > cap) b2 = cap;
    UNETK_DISPATCH_DTYPE(out->dtype, T, {
      im2col_first_kernel_v2<T><<<(int)b2, 256, 0, (cudaStream_t)stream>>>(x_nchw, n, cin, h, w, (T*)out->ptr, out->c, out->ld);
    });
  } else {
    UNETK_DISPATCH_DTYPE(out->dtype, T, {
      im2col_first_kernel<T><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x_nchw, n, cin, h, w, (T*)out->ptr, out->c, out->ld);
    });
  }
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_permute3(const float* src, void* dst, int32_t dst_dtype, int32_t d0, int32_t d1, int32_t d2, int64_t ss0,
                   int64_t ss1, int64_t ss2, int64_t ds0, int64_t ds1, int64_t ds2, void* stream) {
  UNETK_REQUIRE(src && dst && d0 > 0 && d1 > 0 && d2 > 0, "permute3: bad argument");
  UNETK_REQUIRE(dst_dtype == UNETK_F32 || dst_dtype == UNETK_BF16, "permute3: bad dtype");
  const int64_t items = (int64_t)d0 * d1;
  int64_t blocks = (items + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  UNETK_DISPATCH_DTYPE(dst_dtype, T, {
    permute3_kernel<T><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, (T*)dst, d0, d1, d2, ss0, ss1, ss2, ds0, ds1, ds2);
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_weights_pack(const unetk_wjob* jobs, const int32_t* tiles, int32_t ntiles, int32_t dtype, void* stream) {
  UNETK_REQUIRE(jobs && tiles && ntiles > 0, "weights_pack: bad argument");
  UNETK_REQUIRE(dtype == UNETK_F32 || dtype == UNETK_BF16, "weights_pack: bad dtype");
  UNETK_DISPATCH_DTYPE(dtype, T, { weights_kernel<T, true><<<ntiles, 256, 0, (cudaStream_t)stream>>>(jobs, tiles, nullptr); });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_weights_unpack(const unetk_wjob* jobs, const int32_t* tiles, int32_t ntiles, void* dst_base, void* stream) {
  UNETK_REQUIRE(jobs && tiles && ntiles > 0, "weights_unpack: bad argument");
  weights_kernel<float, false><<<ntiles, 256, 0, (cudaStream_t)stream>>>(jobs, tiles, static_cast<uint8_t*>(dst_base));
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}
}
