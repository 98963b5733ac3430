// Internal declarations shared by the kernel translation units and the C-ABI layer (api.cu).
#pragma once

#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/vp3d_b200.h"

namespace vp3d {

// Launch parameters of conv_gemm_kernel (see conv_gemm.cu). All strides are in elements of the named buffer.
struct ConvGemmParams {
  int a_seqs;           // sequences (outermost TMA coordinate of the activation view)
  int rows_out;         // valid output rows (frames) per sequence
  int m_tiles_per_seq;  // ceil(rows_out / 128)
  int n_tiles;          // n_pad / BLOCK_N
  int taps;             // filter taps
  int kblocks_per_tap;  // 128-byte K blocks per tap
  int tap_row_step;     // input-row distance between consecutive taps (dilation)
  int a_row_off;        // input row of tap 0 for output row 0 (may be negative: TMA zero-fills)

  const float* scale;   // per output channel, nullptr = identity
  const float* shift;
  int relu;

  const void* res;      // residual source (same element type as the activations), nullptr = none
  long long res_seq_stride;
  long long res_row_stride;
  int res_row_mul;      // residual row = out_row * res_row_mul + res_row_off
  int res_row_off;

  void* out;
  long long out_seq_stride;
  long long out_row_stride;
  int out_f32;          // 1: fp32 output (shrink layer / tf32 activations), 0: activation element type
  int n_valid;          // real output channels (<= n_pad); only consulted on the fp32 path
  int out_round_tf32;   // fp32 outputs are rounded (RN) to TF32 precision for a following TF32 layer

  float* stat_sum;      // optional per-channel sum / sum-of-squares of the raw accumulator (train-mode BN)
  float* stat_sqsum;
};

cudaError_t launch_conv_gemm(int dtype, int block_n, const CUtensorMap& tmA, const CUtensorMap& tmB,
                             const ConvGemmParams& p, int grid, cudaStream_t stream);

}  // namespace vp3d
