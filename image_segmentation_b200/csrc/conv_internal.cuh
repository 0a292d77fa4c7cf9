// Internal interface between the contraction dispatcher (conv_api.cu) and its two implementations.
#pragma once
#include "common.cuh"

namespace unetk {

struct ConvGeom {
  int taps;        // 1, 9 or 4
  int cin;         // channels per tap of the gathered operand (K = taps*cin)
  int cout;        // channels of y
  int cout_total;  // GEMM N (4*cout for mode 2)
  int rows_n, rows_h, rows_w;  // spatial grid of the GEMM rows
};

int simt_conv(const unetk_conv_args* a, const ConvGeom& g, cudaStream_t stream);
int simt_wgrad(const unetk_wgrad_args* a, int taps, cudaStream_t stream);

// TMA + tcgen05 implementations (bf16 only)
bool tc_conv_supported(const unetk_conv_args* a, const ConvGeom& g, const char** why);
int tc_conv(const unetk_conv_args* a, const ConvGeom& g, cudaStream_t stream);
bool tc_wgrad_supported(const unetk_wgrad_args* a, int taps, const char** why);
int tc_wgrad(const unetk_wgrad_args* a, int taps, cudaStream_t stream);
int64_t tc_wgrad_partial_bytes(const unetk_wgrad_args* a, int taps);

}  // namespace unetk
