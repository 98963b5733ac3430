// Internal declarations shared by the kernel translation units and the C-ABI layer (api.cu).
#pragma once

#include <atomic>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/vp3d_b200.h"

namespace vp3d {

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE property of a kernel: `done` (one per kernel
// instantiation) holds one bit per device ordinal, so a process that drives several GPUs raises the limit on each of them.
inline cudaError_t set_max_smem_once(const void* func, int bytes, std::atomic<unsigned long long>& done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
  return e;
}

// bn_finalize folded into bn_act_fwd: sum == nullptr -> the kernel reads ready-made scale / shift instead
struct BnFinalizeParams {
  const double* sum; const double* sqsum;
  double inv_n;          // 1 / rows the sums were taken over
  float unbias;          // n / (n - 1): running_var takes the unbiased variance
  const float* gamma; const float* beta;
  float eps, momentum;
  float* running_mean; float* running_var; long long* nbt;
  float* scale_out; float* shift_out; float* mean_out; float* invstd_out;
  int c;
};

// Counter-based dropout description (dropout.cuh).
struct DropoutParams {
  float p;
  unsigned long long seed;
  unsigned long long stream;
  const unsigned long long* step_counter;
};

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
  int b_tap_col_step;   // MN-major weights only: weight columns per tap (c_in_pad of the forward layer)
  const int* dyn;       // optional device int[4] {a_row_off, res_row_off, out_row_off, out_row_off2 | -1} (streaming)

  const float* scale;   // per output channel, nullptr = identity
  const float* shift;
  int relu;

  const void* res;      // residual source (same element type as the activations), nullptr = none
  long long res_seq_stride;
  long long res_row_stride;
  int res_row_mul;      // residual row = out_row * res_row_mul + res_row_off
  int res_row_off;
  long long res_rows;   // > 0: rows outside [0, res_rows) add nothing
  int res_col_off;      // residual applies to output columns [res_col_off, res_col_off + res_cols), reading column
  int res_cols;         // col - res_col_off; res_cols == 0: all columns

  void* out;
  long long out_seq_stride;
  long long out_row_stride;
  int out_f32;          // 1: fp32 output (shrink layer / tf32 activations), 0: activation element type
  int out_tma;          // 16-bit outputs leave through TMA stores (tmC); fp32 outputs use direct stores
  int n_valid;          // real output channels (<= n_pad); only consulted on the fp32 path
  int out_round_tf32;   // fp32 outputs are rounded (RN) to TF32 precision for a following TF32 layer

  double* stat_sum;     // optional per-channel sum / sum-of-squares of the raw accumulator (train-mode BN)
  double* stat_sqsum;

  // train-mode BatchNorm finalize in the tail of the GEMM (with stat_sum): the last CTA to finish turns the sums into
  // scale / shift / mean / invstd and updates the running statistics (bn_tail.cuh). fin.sum == nullptr: off.
  BnFinalizeParams fin;
  unsigned int* fin_counter;   // zero on entry; counts finished CTAs

  // CTA-pair kernel: 1 = the grid has one cluster per tile and running clusters steal the tiles of clusters that were not
  // launched yet (cluster launch control) instead of walking a static persistent schedule -- a launch then adapts to
  // SMs that are busy with something else (NCCL's all-reduce CTAs during a data-parallel backward)
  int dyn_sched;

  // CTA-pair kernel only (conv_gemm2.cu, EPI = 1):
  DropoutParams drop;   // drop.p > 0: dropout after the ReLU, keyed by (flat output row seq * rows_out + t, 8-channel group)
  int side_mode;        // epilogue side input fetched by TMA (tmS, geometry of the output): 0 none,
                        // 1: out += side[seq][t + side_row_off][n]   2: out = side[seq][t][n] > 0 ? out * side_scale : 0
  int side_row_off;
  float side_scale;
  int direct_out;       // lean epilogue: 32-byte global stores straight from registers instead of staged TMA stores (set
                        // by the launcher when `out` and its strides are 32-byte aligned)
};

cudaError_t launch_conv_gemm(int dtype, int block_n, int w_mn_major, const CUtensorMap& tmA, const CUtensorMap& tmB,
                             const CUtensorMap& tmC, const ConvGemmParams& p, int grid, cudaStream_t stream);

// CTA-pair variant (conv_gemm2.cu): tcgen05.mma.cta_group::2, each CTA stages its A tile and half of the weight tile.
// `tmB_half` is the weight map with a box of 128 output channels x 64 elements of K.
bool conv_gemm_pair_supported(int dtype, int block_n, int w_mn_major, const ConvGemmParams& p);
cudaError_t launch_conv_gemm_pair(int dtype, int w_mn_major, const CUtensorMap& tmA, const CUtensorMap& tmB_half,
                                  const CUtensorMap& tmC, const CUtensorMap& tmS, const ConvGemmParams& p, int sm_count,
                                  cudaStream_t stream);

// Launch parameters of wgrad_gemm_kernel (see wgrad.cu).
struct WgradParams {
  int num_tiles;        // taps * co_tiles * ci_tiles
  int co_tiles;         // co_pad / BLOCK_M (128 or 256)
  int ci_tiles;         // ci_pad / BLOCK_N
  int seqs;
  int kb_per_seq;       // ceil(rows per sequence / 64)
  int num_slices;       // split of the row axis; items = tiles x slices, dealt round-robin to the CTAs
  int b_row_off;        // input row read against gradient row 0 by tap 0
  int b_tap_row_step;   // extra input rows per tap (dilation)
  int b_tap_col_step;   // extra input columns per tap (stride == width layers on the reshaped view)
  int valid_co, valid_ci;   // real output rows / columns of a tile grid that was padded to the tile size
  int dyn_sched;        // 1: grid = one CTA per item, running CTAs steal the items of CTAs not yet launched (wgrad.cu)
  float* out;           // packed fp32 [taps][co_pad][ci_pad]
  long long out_tap_stride;
  long long out_row_stride;
};
cudaError_t launch_wgrad(int dtype, int block_n, int block_m, const CUtensorMap& tmA, const CUtensorMap& tmB,
                         const WgradParams& p, int grid, cudaStream_t stream);
cudaError_t launch_wgrad_finish(const float* packed, float* dw, int c_out, int c_in, int taps, long long tap_stride,
                                long long row_stride, const float* gscale_buf, int sm_count, cudaStream_t stream);

// Fused low-latency streaming step (stream.cu): one cooperative kernel per frame for a few concurrent streams.
struct StreamLayer {
  const void* a;        // layer input: flat [rows][k_per_tap] view of a ring (or a plain buffer)
  const void* w;        // K-major weights [n][taps * k_per_tap] (BatchNorm scale folded in)
  const float* shift;   // [n] or nullptr
  const void* res;      // residual ring (16-bit, row stride res_row_stride) or nullptr
  void* out;            // 16-bit ring / buffer, or fp32 [streams][out_row_stride]
  int a_ring, res_ring, out_ring;   // ring indices whose positions apply (-1: rows start at 0 / no mirror copy)
  int k_per_tap, taps, tap_row_step;
  int n, n_valid, relu, out_f32;
  int res_row_stride, out_row_stride;
};
constexpr int kStreamMaxLayers = 12;
struct StreamStepParams {
  long long* step;                  // device frame counter (read at the start, incremented at the end)
  unsigned long long* barrier;      // 128 device words of grid-barrier counters: 0 when *step == 0, never reset in between
  const int* ring_len; const int* ring_dil; const int* ring_taps;
  int n_rings, rows_per_slot;
  const float* x_in;                // [n_streams][c_in] new frame
  void* ring0;
  int c_in, c_in_pad, n_streams, n_layers;
  StreamLayer layers[kStreamMaxLayers];
};
int stream_step_max_streams();
int stream_step_max_k();
cudaError_t launch_stream_step(int dtype, const StreamStepParams& p, int sm_count, cudaStream_t stream);

cudaError_t launch_counter_add(unsigned long long* counter, unsigned long long inc, cudaStream_t stream);
struct AdamParams {
  float* p; const float* g; float* m; float* v; float* vmax;
  long long n;
  float lr, beta1, beta2, eps, weight_decay;
  const float* step;    // device: step count of THIS update (already incremented)
  const float* lr_dev;  // optional device learning rate (overrides lr)
  int maximize;
  void* packed;         // optional 16-bit K-major operand [c_out_pad][taps][k_pad]
  int c_in, taps, k_pad;
};
cudaError_t launch_adam_pack(int dtype, const AdamParams& a, int sm_count, cudaStream_t stream);
// Several tensors in ONE launch (the 30 parameter tensors of a model: 10 convolution weights with their packed
// operands, 18 BatchNorm affine vectors, the shrink bias): consecutive block ranges are dealt to the tensors in proportion
// to their size. The hyper-parameters are shared, `step` / `packed` geometry are per tensor.
constexpr int kAdamMaxTensors = 32;
struct AdamTensor {
  float* p; const float* g; float* m; float* v; float* vmax;
  long long n;
  const float* step;
  void* packed;
  int c_in, taps, k_pad, reserved;
};
struct AdamMultiParams {
  AdamTensor t[kAdamMaxTensors];
  int block_start[kAdamMaxTensors + 1];   // tensor i owns blocks [block_start[i], block_start[i + 1])
  int count;
  float lr, beta1, beta2, eps, weight_decay;
  const float* lr_dev;
  int maximize;
};
cudaError_t launch_adam_multi(int dtype, const AdamMultiParams& a, cudaStream_t stream);
cudaError_t launch_stream_advance(long long* step, int n_rings, const int* ring_len, const int* ring_dil,
                                  const int* ring_taps, int rows_per_slot, int* table, int n_launch,
                                  const int* launch_desc, int* launch_table, cudaStream_t stream);
cudaError_t launch_ring_write(int dtype, const float* src, void* ring, const int* table, long long rows, int c, int c_pad,
                              int sm_count, cudaStream_t stream);
cudaError_t launch_bn_finalize(const double* sum, const double* sqsum, long long count, const float* gamma,
                               const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                               long long* nbt, float* scale, float* shift, float* mean, float* invstd, int c, int c_pad,
                               cudaStream_t stream);
cudaError_t launch_bn_act_fwd(int dtype, const void* z, const float* scale, const float* shift, const void* res,
                              long long seqs, long long rows_per_seq, long long res_seq_rows, int res_row_mul,
                              int res_row_off, int c_pad, const DropoutParams& dp, void* a,
                              const BnFinalizeParams& fin, int sm_count, cudaStream_t stream,
                              unsigned char* keep_mask = nullptr);
// backward passes that read the keep bits the forward stored (train.cu)
cudaError_t launch_bn_act_bwd_reduce_mask(int dtype, const void* g, const void* z, const unsigned char* keep,
                                          const float* mean, const float* invstd, float keep_scale, long long rows,
                                          int c_pad, double* sum_dy, double* sum_dy_xhat, int sm_count,
                                          cudaStream_t stream);
cudaError_t launch_bn_act_bwd_apply_mask(int dtype, const void* g, const void* z, const unsigned char* keep,
                                         const float* scale, const float* mean, const float* invstd, float keep_scale,
                                         long long rows, long long count, int c, int c_pad, const double* sum_dy,
                                         const double* sum_dy_xhat, const float* gscale_buf, void* dz, float* d_gamma,
                                         float* d_beta, int sm_count, cudaStream_t stream);
cudaError_t launch_col_stats(int dtype, const void* z, long long rows, int c_pad, double* sum, double* sqsum,
                             int sm_count, cudaStream_t stream);
cudaError_t launch_bn_act_bwd_reduce(int dtype, const void* g, const void* z, const float* scale, const float* shift,
                                     const float* mean, const float* invstd, long long rows, int c_pad,
                                     const DropoutParams& dp, double* sum_dy, double* sum_dy_xhat, int sm_count,
                                     cudaStream_t stream);
cudaError_t launch_bn_act_bwd_apply(int dtype, const void* g, const void* z, const float* scale, const float* shift,
                                    const float* mean, const float* invstd, long long rows, long long count, int c,
                                    int c_pad, const DropoutParams& dp, const double* sum_dy, const double* sum_dy_xhat,
                                    const float* gscale_buf, void* dz, float* d_gamma, float* d_beta, int sm_count,
                                    cudaStream_t stream);
cudaError_t launch_expand_bn_stats(int dtype, const float* G, const void* w, int k_total, int ones_col,
                                   const float* gamma, const float* beta, float eps, float momentum,
                                   float* running_mean, float* running_var, long long* nbt, float* scale, float* shift,
                                   float* mean, float* invstd, float* wg, int c, int c_pad, cudaStream_t stream);
cudaError_t launch_expand_bwd_finish(int dtype, const float* P, const float* wg, const float* G, const void* w, int k_total,
                                     int ones_col, const float* scale, const float* mean, const float* invstd,
                                     const float* gscale_buf, int c, int c_pad, int c_in, int c_in_pad, int taps, float* dw,
                                     float* d_gamma, float* d_beta, cudaStream_t stream);
cudaError_t launch_grad_scale(const float* dy, long long n, float* gscale_buf, int sm_count, cudaStream_t stream);
cudaError_t launch_grad_pack_rows(int dtype, const float* src, void* dst, long long rows, int c, int c_pad,
                                  const float* gscale_buf, float* col_sum, int sm_count, cudaStream_t stream);

// Peer-memory gradient exchange (allreduce.cu). Pointer tables by value: at most kArMaxRanks ranks of one NVLink domain.
constexpr int kArMaxRanks = 16;
constexpr int kArMaxCtas = 64;
struct AllReduceParams {
  float* mc;                      // multicast mapping of the exchange buffer, or nullptr (plain peer loads / stores)
  float* peers[kArMaxRanks];      // this process's mapping of every rank's exchange buffer
  uint32_t* flags[kArMaxRanks];   // ... of every rank's flag words [ctas][world]
  int rank, world;
  long long off, n;               // fp32 elements, multiples of 4
  float scale;
  unsigned long long timeout_ns;
};
cudaError_t launch_peer_allreduce(const AllReduceParams& p, int ctas, cudaStream_t stream);

}  // namespace vp3d
