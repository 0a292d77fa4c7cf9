// CUDA-core (FFMA) implicit-GEMM contractions: the fp32 parity tier and the cross-check for the tcgen05 kernels.
// Same operand layouts and modes as the tensor-core path (see unetk.h, unetk_conv / unetk_wgrad).
//
// Reference call sites replaced: unet/unet.py:16,19 (Conv2d 3x3 p1), :59 (ConvTranspose2d k2 s2) and their
// convolution_backward (data + filter gradients).  Accumulation is fp32 regardless of storage type.
#include "conv_internal.cuh"

namespace unetk {

constexpr int BM = 64, BN = 64, BK = 16;

struct GatherGeom {
  int n, h, w;    // grid of GEMM rows (output pixels for modes 0/1/3, input pixels for mode 2)
  int sh, sw;     // spatial size of the gathered tensor
  int mode;
};

// source pixel (in the gathered tensor) of GEMM row (img,y,x) for tap t; returns false when it is padding
__device__ __forceinline__ bool gather_pixel(const GatherGeom& g, int img, int y, int x, int t, int64_t& pix) {
  int yy, xx;
  if (g.mode == 1) {
    yy = y + t / 3 - 1;
    xx = x + t % 3 - 1;
    if (yy < 0 || yy >= g.sh || xx < 0 || xx >= g.sw) return false;
  } else if (g.mode == 3) {
    yy = 2 * y + (t >> 1);
    xx = 2 * x + (t & 1);
  } else {
    yy = y;
    xx = x;
  }
  pix = ((int64_t)img * g.sh + yy) * g.sw + xx;
  return true;
}

// y[m, n] = sum_k A[m, k] * B[n, k];  A gathered from x, B = packed weights [N][K]
template <typename T>
__global__ void __launch_bounds__(256) simt_conv_kernel(const T* __restrict__ x, int xld, int cin, const T* __restrict__ wt,
                                                        T* __restrict__ y, int yld, int cout_total, int cout,
                                                        GatherGeom g, int taps, const float* __restrict__ bias) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int64_t M = (int64_t)g.n * g.h * g.w;
  const int K = taps * cin;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;  // loader mapping: 64 rows x 4 k-quads
  const int ty = tid >> 4, tx = tid & 15;         // compute mapping: 16 x 16 threads, 4x4 outputs each

  // decode the loader's GEMM row once
  const int64_t mrow = m0 + lrow;
  const bool row_ok = mrow < M;
  int rimg = 0, ry = 0, rx = 0;
  if (row_ok) {
    rx = (int)(mrow % g.w);
    ry = (int)((mrow / g.w) % g.h);
    rimg = (int)(mrow / ((int64_t)g.w * g.h));
  }
  const int bn = n0 + lrow;
  const bool bn_ok = bn < cout_total;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    const int t = k0 / cin, c0 = k0 % cin;  // cin % 16 == 0 so a K-chunk never straddles taps
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    int64_t pix;
    if (row_ok && gather_pixel(g, rimg, ry, rx, t, pix)) load4(x + pix * xld + c0 + lk, av);
    if (bn_ok) load4(wt + (int64_t)bn * K + k0 + lk, bv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[lk + j][lrow] = av[j];
      Bs[lk + j][lrow] = bv[j];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int px = (int)(m % g.w), py = (int)((m / g.w) % g.h), img = (int)(m / ((int64_t)g.w * g.h));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n0 + tx * 4 + j;
      if (nn >= cout_total) continue;
      if (g.mode == 2) {
        // column = (a*2+b)*cout + co -> output pixel (2y+a, 2x+b)
        const int q = nn / cout, co = nn % cout;
        const int64_t op = ((int64_t)img * (2 * g.h) + 2 * py + (q >> 1)) * (2 * g.w) + 2 * px + (q & 1);
        y[op * yld + co] = from_f<T>(acc[i][j] + (bias ? bias[co] : 0.f));
      } else {
        y[m * yld + nn] = from_f<T>(acc[i][j] + (bias ? bias[nn] : 0.f));
      }
    }
  }
}

// dw[cu][t][cs] += sum_p U[p, cu] * S[gather(p,t), cs]   (M = cu, N = cs within tap t, K = pixels, split-K)
template <typename T>
__global__ void __launch_bounds__(256) simt_wgrad_kernel(const T* __restrict__ u, int uld, int cu, const T* __restrict__ s,
                                                         int sld, int cs, GatherGeom g, int taps, int pix_per_split,
                                                         float* __restrict__ dw) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int64_t P = (int64_t)g.n * g.h * g.w;
  const int m0 = blockIdx.x * BM;                 // cu tile
  const int ntile = blockIdx.y;                   // (tap, cs tile)
  const int cs_tiles = (cs + BN - 1) / BN;
  const int t = ntile / cs_tiles, n0 = (ntile % cs_tiles) * BN;
  const int64_t p_begin = (int64_t)blockIdx.z * pix_per_split;
  const int64_t p_end = min(p_begin + (int64_t)pix_per_split, P);
  const int tid = threadIdx.x;
  const int lp = tid >> 4, lc = (tid & 15) * 4;   // loader: 16 pixels x 16 channel-quads
  const int ty = tid >> 4, tx = tid & 15;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t p0 = p_begin; p0 < p_end; p0 += BK) {
    const int64_t p = p0 + lp;
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (p < p_end) {
      if (m0 + lc < cu) load4(u + p * uld + m0 + lc, av);
      const int px = (int)(p % g.w), py = (int)((p / g.w) % g.h), img = (int)(p / ((int64_t)g.w * g.h));
      int64_t pix;
      if (n0 + lc < cs && gather_pixel(g, img, py, px, t, pix)) load4(s + pix * sld + n0 + lc, bv);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[lp][lc + j] = av[j];
      Bs[lp][lc + j] = bv[j];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= cu) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n0 + tx * 4 + j;
      if (nn >= cs) continue;
      atomicAdd(dw + ((int64_t)m * taps + t) * cs + nn, acc[i][j]);
    }
  }
}

int simt_conv(const unetk_conv_args* a, const ConvGeom& cg, cudaStream_t stream) {
  GatherGeom g;
  g.n = cg.rows_n; g.h = cg.rows_h; g.w = cg.rows_w; g.sh = a->x.h; g.sw = a->x.w; g.mode = a->mode;
  const int64_t M = (int64_t)g.n * g.h * g.w;
  UNETK_REQUIRE(cg.cin % 16 == 0 && a->x.ld % 4 == 0, "conv(simt): Cin must be a multiple of 16 and ld of 4");
  UNETK_REQUIRE((M + BM - 1) / BM < (1LL << 31), "conv(simt): too many rows");
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((cg.cout_total + BN - 1) / BN));
  UNETK_DISPATCH_DTYPE(a->x.dtype, T, {
    simt_conv_kernel<T><<<grid, 256, 0, stream>>>((const T*)a->x.ptr, a->x.ld, cg.cin, (const T*)a->w, (T*)a->y.ptr, a->y.ld,
                                                  cg.cout_total, cg.cout, g, cg.taps, a->bias);
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int simt_wgrad(const unetk_wgrad_args* a, int taps, cudaStream_t stream) {
  GatherGeom g;
  g.n = a->u.n; g.h = a->u.h; g.w = a->u.w; g.sh = a->s.h; g.sw = a->s.w;
  g.mode = a->mode == 2 ? 3 : a->mode;  // stride-2 gather of the hi-res operand
  const int cu = a->u.c, cs = a->s.c;
  UNETK_REQUIRE(cu % 4 == 0 && cs % 4 == 0 && a->u.ld % 4 == 0 && a->s.ld % 4 == 0, "wgrad(simt): channels must be multiples of 4");
  const int64_t P = pixels(a->u);
  const int mt = (cu + BM - 1) / BM, nt = taps * ((cs + BN - 1) / BN);
  int64_t splits = (4LL * sm_count() + (int64_t)mt * nt - 1) / ((int64_t)mt * nt);
  const int64_t max_splits = (P + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int64_t pps = (P + splits - 1) / splits;
  pps = (pps + BK - 1) / BK * BK;
  splits = (P + pps - 1) / pps;
  dim3 grid(mt, nt, (unsigned)splits);
  UNETK_DISPATCH_DTYPE(a->u.dtype, T, {
    simt_wgrad_kernel<T><<<grid, 256, 0, stream>>>((const T*)a->u.ptr, a->u.ld, cu, (const T*)a->s.ptr, a->s.ld, cs, g, taps,
                                                   (int)pps, a->dw);
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

}  // namespace unetk
