// Train-mode BatchNorm statistics -> affine, shared by bn_finalize_kernel / bn_act_fwd_kernel (train.cu) and the tail of
// the convolution GEMMs (conv_gemm.cu, conv_gemm2.cu), so that every path gives the same bits.
#pragma once

#include "kernels.h"

namespace vp3d {

// Batch statistics of one channel from its double-precision sums. Only the cancellation-prone part (E[z^2] - E[z]^2)
// is done in double precision -- three operations; a double division / square root per channel in EVERY thread of the
// fused apply pass cost 55 us per training step on B200's thin FP64 pipe -- and invstd is an IEEE fp32
// 1 / sqrt(var + eps), the precision F.batch_norm itself normalises with.
struct BnChannel {
  float mean, invstd, var;
};
__device__ __forceinline__ BnChannel bn_channel_stats(double sum, double sqsum, double inv_n, float eps) {
  const double m = sum * inv_n;
  double var = fma(sqsum, inv_n, -m * m);   // biased, as F.batch_norm normalises with
  if (var < 0.0) var = 0.0;
  BnChannel r;
  r.mean = (float)m;
  r.var = (float)var;
  r.invstd = 1.f / sqrtf(r.var + eps);
  return r;
}

// One block: sums -> scale / shift / mean / invstd [c_pad] (+ running statistics, num_batches_tracked). momentum < 0:
// cumulative moving average, factor 1 / (batches seen so far + 1) (nn.BatchNorm1d(momentum=None)).
__device__ __forceinline__ void bn_finalize_block(const BnFinalizeParams& f, int c_pad, bool volatile_sums) {
  float momentum = f.momentum;
  if (momentum < 0.f) {
    const long long seen = f.nbt != nullptr ? *f.nbt : 0;
    momentum = 1.f / (float)(seen + 1);
    __syncthreads();
  }
  if (threadIdx.x == 0 && f.nbt != nullptr) *f.nbt += 1;
  // four channels per thread and iteration, every load issued before the first use (in the GEMM tail this block is the
  // last thing the whole launch waits for, and one dependent L2 round trip per channel added ~6 us)
  constexpr int U = 4;
  for (int base = 0; base < c_pad; base += U * (int)blockDim.x) {
    double s[U], q[U];
    float ga[U], be[U], rm[U], rv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * (int)blockDim.x + (int)threadIdx.x;
      const bool on = i < f.c;
      s[u] = on ? (volatile_sums ? __ldcg(f.sum + i) : f.sum[i]) : 0.0;
      q[u] = on ? (volatile_sums ? __ldcg(f.sqsum + i) : f.sqsum[i]) : 0.0;
      ga[u] = on ? f.gamma[i] : 0.f;
      be[u] = on ? f.beta[i] : 0.f;
      rm[u] = (on && f.running_mean != nullptr) ? f.running_mean[i] : 0.f;
      rv[u] = (on && f.running_mean != nullptr) ? f.running_var[i] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * (int)blockDim.x + (int)threadIdx.x;
      if (i >= c_pad) continue;
      if (i >= f.c) {
        f.scale_out[i] = 0.f;
        f.shift_out[i] = 0.f;
        f.mean_out[i] = 0.f;
        f.invstd_out[i] = 0.f;
        continue;
      }
      const BnChannel ch = bn_channel_stats(s[u], q[u], f.inv_n, f.eps);
      const float sc = __fmul_rn(ga[u], ch.invstd);
      f.scale_out[i] = sc;
      f.shift_out[i] = __fsub_rn(be[u], __fmul_rn(ch.mean, sc));
      f.mean_out[i] = ch.mean;
      f.invstd_out[i] = ch.invstd;
      if (f.running_mean != nullptr) {
        // explicit roundings: the same bits whichever kernel this is inlined into (no context-dependent fma contraction)
        f.running_mean[i] = __fadd_rn(__fmul_rn(1.f - momentum, rm[u]), __fmul_rn(momentum, ch.mean));
        f.running_var[i] = __fadd_rn(__fmul_rn(1.f - momentum, rv[u]),
                                     __fmul_rn(momentum, __fmul_rn(ch.var, f.unbias)));   // unbiased: n / (n - 1)
      }
    }
  }
}

// Tail of a GEMM that accumulated the sums with atomics: called by ALL threads of every CTA after the CTA's own atomics
// were issued and a __syncthreads() (the barrier orders them before thread 0's fence, which is cumulative). The last CTA
// to arrive finalizes; the others return. `counter` is zero on entry.
// `s_last`: one int of the kernel's (dynamic) shared memory -- the GEMM kernels have no byte of static shared memory to
// spare under the 227 KB limit.
__device__ __forceinline__ void bn_finalize_tail(const BnFinalizeParams& f, unsigned int* counter, int c_pad,
                                                 volatile int* s_last) {
  if (threadIdx.x == 0) {
    __threadfence();
    *s_last = atomicAdd(counter, 1u) == gridDim.x - 1 ? 1 : 0;
  }
  __syncthreads();
  if (!*s_last) return;
  __threadfence();
  bn_finalize_block(f, c_pad, true);
}

}  // namespace vp3d
