// Train-mode BatchNorm of the EXPAND layer without ever touching its (rows x 1024) output for statistics.
//
// Reference: x = drop(relu(expand_bn(expand_conv(x)))) (common/models/TemporalModel.py:127 / :189) and its autograd
// backward (run.py:485). The expand convolution is linear with a tiny contraction length K = taps * c_in (102 for 17
// joints), so everything BatchNorm needs is a function of the K x K Gram matrix of the layer INPUT:
//
//   z = X w_c  (X: rows x K view of the input, w_c: the K weights of channel c)
//   sum_r z        = w_c . s                    s = X^T 1          (column sums)
//   sum_r z^2      = w_c^T G w_c                G = X^T X
//   sum_r z x_k    = (W G)[c][k]
//
// Forward: G (one 256 x 256 tcgen05 tile, vp3d_wgrad with both operands = X; a constant-one input column supplies s and
// the row count) -> expand_bn_stats_kernel -> scale / shift, so the GEMM epilogue applies BatchNorm + ReLU + dropout
// itself and the raw output z is never stored or re-read (saves the 170 MB write + 340 MB apply pass at batch 1024).
// Backward: the data-gradient GEMM of the next layer gates its result by (a > 0) in its epilogue (conv_gemm2.cu, side
// mode 2) -> gm = g * relu/dropout mask; P = gm^T X is the ordinary weight-gradient GEMM of the layer; then with
//   Sg = sum_r gm (= P[c][ones column]),  Sgz = sum_k W[c][k] P[c][k] (= sum_r gm z),
//   d_beta = Sg,  d_gamma = invstd (Sgz - mean Sg),
//   dW[c][k] = gamma invstd ( P[c][k] - (Sg / n) s_k - (d_gamma / n) invstd ((W G)[c][k] - mean s_k) )
// which is exactly sum_r dz[r][c] x[r][k] with dz the BatchNorm backward -- no reduce pass, no apply pass, no dz matrix.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "kernels.h"
#include "pdl.cuh"

namespace vp3d {

namespace {

constexpr int kExCh = 8;        // channels per block
constexpr int kExThreads = 256; // one thread per column of the 256-wide Gram tile

template <int DT>
__device__ __forceinline__ float load_w(const void* w, long long i) {
  if (DT == VP3D_F16) return __half2float(static_cast<const __half*>(w)[i]);
  return __bfloat162float(static_cast<const __nv_bfloat16*>(w)[i]);
}

// sum over the 256 threads of a block of kExCh values each -> every thread gets the totals (red: [kExCh][8] floats)
__device__ __forceinline__ void block_sum(float (&v)[kExCh], float (*red)[8]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < kExCh; ++c) {
    float x = v[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[c][warp] = x;
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < kExCh; ++c) {
    float x = 0.f;
#pragma unroll
    for (int w = 0; w < kExThreads / 32; ++w) x += red[c][w];
    v[c] = x;
  }
  __syncthreads();
}

// grid = c_pad / kExCh blocks. G: fp32 [256][256] (rows / columns >= k_total are zero), ones_col: the input column that
// holds 1.0 (its weight is zero). Writes scale / shift / mean / invstd [c_pad], wg [c_pad][256] = (W G)[c][k], updates
// the running statistics like F.batch_norm(training=True).
template <int DT>
__global__ void __launch_bounds__(kExThreads)
expand_bn_stats_kernel(const float* __restrict__ G, const void* __restrict__ w, int k_total, int ones_col,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                       float* __restrict__ running_mean, float* __restrict__ running_var, long long* __restrict__ nbt,
                       float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                       float* __restrict__ invstd_out, float* __restrict__ wg, int c, int tick_inline) {
  pdl_enter();
  // weights of the block's 8 channels, [column][channel]: the inner loop below reads one column's 8 weights as two
  // broadcast LDS.128 (a [channel][column] layout needs 8 scalar LDS per column, and the kernel was bound by the LSU's
  // instruction rate: 30 us)
  __shared__ __align__(16) float ws[256][kExCh];
  __shared__ float red[kExCh][8];
  static_assert(kExCh == 8, "two float4 per column");
  const int k = threadIdx.x;
  const int c0 = blockIdx.x * kExCh;
  if (momentum < 0.f) momentum = 1.f / (float)((nbt != nullptr ? *nbt : 0) + 1);   // cumulative average (momentum=None)
  if (tick_inline && blockIdx.x == 0 && k == 0 && nbt != nullptr) *nbt += 1;       // nobody reads the count in this mode
  float wk[kExCh];                                  // this thread's column of the 8 weight rows
#pragma unroll
  for (int cc = 0; cc < kExCh; ++cc) {
    wk[cc] = (k < k_total && c0 + cc < c) ? load_w<DT>(w, (long long)(c0 + cc) * k_total + k) : 0.f;
    ws[k][cc] = wk[cc];
  }
  __syncthreads();
  const float n = G[ones_col * 256 + ones_col];
  const float inv_n = 1.f / n;
  const float s_k = G[ones_col * 256 + k];          // column sum of input column k (0 for k >= k_total)
  // t[cc] = sum_j G[j][k] w[cc][j]  (G symmetric: walking column k keeps the warp's loads coalesced)
  float t[kExCh], dot_s[kExCh];
#pragma unroll
  for (int cc = 0; cc < kExCh; ++cc) t[cc] = 0.f;
  // (k_total is a run-time value: without the unroll hint one L2 load is in flight per thread and the 192 dependent
  // round trips cost 30 us)
#pragma unroll 16
  for (int j = 0; j < k_total; ++j) {
    const float g = __ldg(G + j * 256 + k);
    const float4 w0 = *reinterpret_cast<const float4*>(&ws[j][0]);
    const float4 w1 = *reinterpret_cast<const float4*>(&ws[j][4]);
    t[0] = fmaf(g, w0.x, t[0]);
    t[1] = fmaf(g, w0.y, t[1]);
    t[2] = fmaf(g, w0.z, t[2]);
    t[3] = fmaf(g, w0.w, t[3]);
    t[4] = fmaf(g, w1.x, t[4]);
    t[5] = fmaf(g, w1.y, t[5]);
    t[6] = fmaf(g, w1.z, t[6]);
    t[7] = fmaf(g, w1.w, t[7]);
  }
#pragma unroll
  for (int cc = 0; cc < kExCh; ++cc) {
    wg[(long long)(c0 + cc) * 256 + k] = t[cc];
    dot_s[cc] = wk[cc] * s_k;                       // -> w . s
  }
  block_sum(dot_s, red);
  // centred quadratic form: var = (1/n) sum_k w_k (t_k - s_k (w . s) / n)  -- the subtraction happens per column, before
  // the sum, so the cancellation of E[z^2] - E[z]^2 is spread over K small terms
  float q[kExCh];
#pragma unroll
  for (int cc = 0; cc < kExCh; ++cc) q[cc] = wk[cc] * (t[cc] - s_k * dot_s[cc] * inv_n);
  block_sum(q, red);
  if (k < kExCh) {
    const int ch = c0 + k;
    float m = 0.f, var = 0.f;
#pragma unroll
    for (int cc = 0; cc < kExCh; ++cc)
      if (cc == k) {
        m = dot_s[cc] * inv_n;
        var = fmaxf(q[cc] * inv_n, 0.f);
      }
    if (ch < c) {
      const float invstd = 1.f / sqrtf(var + eps);
      const float sc = gamma[ch] * invstd;
      scale[ch] = sc;
      shift[ch] = beta[ch] - m * sc;
      mean_out[ch] = m;
      invstd_out[ch] = invstd;
      if (running_mean != nullptr) {
        running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * m;
        running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (var * (n / (n - 1.f)));
      }
    } else {
      scale[ch] = shift[ch] = mean_out[ch] = invstd_out[ch] = 0.f;
    }
  }
}

__global__ void counter_tick_kernel(long long* nbt) {
  pdl_enter();
  *nbt += 1;
}

// grid = c_pad / kExCh blocks, thread = column k. P: fp32 [c_pad][256] = gm^T X (scaled by gscale), wg from the forward.
// dw: nn.Conv1d layout (c_out, c_in, taps); column k = tap * c_in_pad + ci.
template <int DT>
__global__ void __launch_bounds__(kExThreads)
expand_bwd_finish_kernel(const float* __restrict__ P, const float* __restrict__ wg, const float* __restrict__ G,
                         const void* __restrict__ w, int k_total, int ones_col, const float* __restrict__ scale,
                         const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ gscale_buf, int c, int c_in, int c_in_pad, int taps,
                         float* __restrict__ dw, float* __restrict__ d_gamma, float* __restrict__ d_beta) {
  pdl_enter();
  __shared__ float red[kExCh][8];
  const int k = threadIdx.x;
  const int c0 = blockIdx.x * kExCh;
  const float inv_gs = gscale_buf != nullptr ? gscale_buf[1] : 1.f;
  const float n = G[ones_col * 256 + ones_col];
  const float inv_n = 1.f / n;
  const float s_k = G[ones_col * 256 + k];
  float p[kExCh], sgz[kExCh];
#pragma unroll
  for (int cc = 0; cc < kExCh; ++cc) {
    const int ch = c0 + cc;
    p[cc] = ch < c ? P[(long long)ch * 256 + k] : 0.f;
    sgz[cc] = (ch < c && k < k_total) ? p[cc] * load_w<DT>(w, (long long)ch * k_total + k) : 0.f;
  }
  block_sum(sgz, red);
  const int tap = k / c_in_pad, ci = k - tap * c_in_pad;
#pragma unroll
  for (int cc = 0; cc < kExCh; ++cc) {
    const int ch = c0 + cc;
    if (ch >= c) continue;
    const float sg = P[(long long)ch * 256 + ones_col];          // sum_r gm
    const float mu = mean[ch], is = invstd[ch];
    const float dg = is * (sgz[cc] - mu * sg);                   // sum_r gm xhat
    if (k == 0) {
      d_beta[ch] = sg * inv_gs;
      d_gamma[ch] = dg * inv_gs;
    }
    if (k < k_total && ci < c_in) {
      const float xhx = is * (wg[(long long)ch * 256 + k] - mu * s_k);   // sum_r xhat x_k
      const float v = scale[ch] * (p[cc] - sg * inv_n * s_k - dg * inv_n * xhx);
      dw[((long long)ch * c_in + ci) * taps + tap] = v * inv_gs;
    }
  }
}

}  // namespace

cudaError_t launch_expand_bn_stats(int dtype, const float* G, const void* w, int k_total, int ones_col,
                                   const float* gamma, const float* beta, float eps, float momentum,
                                   float* running_mean, float* running_var, long long* nbt, float* scale, float* shift,
                                   float* mean, float* invstd, float* wg, int c, int c_pad, cudaStream_t stream) {
  const int grid = c_pad / kExCh;
  const int tick_inline = momentum >= 0.f ? 1 : 0;
  if (dtype == VP3D_F16)
    launch_k(expand_bn_stats_kernel<VP3D_F16>, dim3(grid), dim3(kExThreads), 0, stream, G, w, k_total, ones_col, gamma, beta, eps, momentum,
                                                                      running_mean, running_var, nbt, scale, shift, mean,
                                                                      invstd, wg, c, tick_inline);
  else if (dtype == VP3D_BF16)
    launch_k(expand_bn_stats_kernel<VP3D_BF16>, dim3(grid), dim3(kExThreads), 0, stream, G, w, k_total, ones_col, gamma, beta, eps, momentum,
                                                                       running_mean, running_var, nbt, scale, shift, mean,
                                                                       invstd, wg, c, tick_inline);
  else
    return cudaErrorInvalidValue;
  if (nbt != nullptr && !tick_inline) launch_k(counter_tick_kernel, dim3(1), dim3(1), 0, stream, nbt);   // after every block has read the old count
  return cudaGetLastError();
}

cudaError_t launch_expand_bwd_finish(int dtype, const float* P, const float* wg, const float* G, const void* w, int k_total,
                                     int ones_col, const float* scale, const float* mean, const float* invstd,
                                     const float* gscale_buf, int c, int c_pad, int c_in, int c_in_pad, int taps, float* dw,
                                     float* d_gamma, float* d_beta, cudaStream_t stream) {
  const int grid = c_pad / kExCh;
  if (dtype == VP3D_F16)
    launch_k(expand_bwd_finish_kernel<VP3D_F16>, dim3(grid), dim3(kExThreads), 0, stream, P, wg, G, w, k_total, ones_col, scale, mean, invstd,
                                                                        gscale_buf, c, c_in, c_in_pad, taps, dw, d_gamma,
                                                                        d_beta);
  else if (dtype == VP3D_BF16)
    launch_k(expand_bwd_finish_kernel<VP3D_BF16>, dim3(grid), dim3(kExThreads), 0, stream, P, wg, G, w, k_total, ones_col, scale, mean,
                                                                         invstd, gscale_buf, c, c_in, c_in_pad, taps, dw,
                                                                         d_gamma, d_beta);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace vp3d
