// Train-mode BatchNorm / ReLU / Dropout / residual kernels around the convolution GEMMs (all HBM-bandwidth bound).
//
// What they replace in the reference (common/models/TemporalModel.py in train() mode, backward at run.py:485):
//   forward   x = drop(relu(bn(conv(x))))            :127,134 / :189,194     bn_finalize + bn_act_fwd
//             x = res + drop(relu(bn(conv(x))))      :135 / :195             (residual rows res[t * mul + off])
//   backward  the autograd chain of the same ops                              bn_act_bwd_reduce + bn_act_bwd_apply
// The per-channel sum / sum of squares of the raw convolution output comes from the GEMM epilogue (conv_gemm.cu), so
// the forward touches each activation matrix twice (read z, write a) and the backward three times (2 x read g,z;
// write dz). Matrices are channels-last [rows][c_pad] in the 16-bit operand type; a thread owns 8 consecutive channels
// (one 16-byte vector) and walks rows, so every access is a full coalesced 16-byte lane.
//
// Dropout is counter based: the keep decision of (row, 8-channel group) is 8 x 16 bits of one Philox4x32-10 block keyed
// by (seed, stream), so the backward recomputes the forward mask instead of storing it.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "kernels.h"
#include "pdl.cuh"
#include "dropout.cuh"
#include "bn_tail.cuh"

namespace vp3d {

constexpr int kEwThreads = 256;

// ---------------------------------------------------------------------------------------------- helpers
template <int DT>
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 v;
    if (DT == VP3D_F16) v = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    else v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    f[2 * i] = v.x;
    f[2 * i + 1] = v.y;
  }
}
template <int DT>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (DT == VP3D_F16) {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    } else {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// helpers of the lean forward pass / the stored-mask backward passes
__device__ __forceinline__ uint32_t prmt_sign(uint32_t a, uint32_t sel) {   // prmt with zero as the second source
  uint32_t d;
  asm("prmt.b32 %0, %1, 0, %2;" : "=r"(d) : "r"(a), "r"(sel));
  return d;
}
__device__ __forceinline__ uint32_t prmt2(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
template <int DT>
__device__ __forceinline__ uint32_t pack2_relu(float a, float b) {   // {lo = relu(a), hi = relu(b)} in the operand type
  uint32_t d;
  if (DT == VP3D_F16) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}

// ---------------------------------------------------------------------------------------------- bn_finalize
// One block walks all channels (c_pad <= a few thousand): the step counter is read by every thread BEFORE thread 0 ticks
// it, which the cumulative-average mode (momentum < 0: nn.BatchNorm1d(momentum=None), factor 1 / num_batches_tracked
// after the tick) needs and a multi-block grid could not order.
__global__ void __launch_bounds__(1024)
bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sqsum, double inv_n, float unbias,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                   float* __restrict__ running_mean, float* __restrict__ running_var, long long* __restrict__ nbt,
                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                   float* __restrict__ invstd_out, int c, int c_pad) {
  pdl_enter();
  BnFinalizeParams f;
  f.sum = sum; f.sqsum = sqsum; f.inv_n = inv_n; f.unbias = unbias; f.gamma = gamma; f.beta = beta; f.eps = eps;
  f.momentum = momentum; f.running_mean = running_mean; f.running_var = running_var; f.nbt = nbt;
  f.scale_out = scale; f.shift_out = shift; f.mean_out = mean_out; f.invstd_out = invstd_out; f.c = c;
  bn_finalize_block(f, c_pad, false);
}

// ---------------------------------------------------------------------------------------------- row-walking kernels
// Shared shape of the HBM-bound kernels below: a thread owns one channel group (8 consecutive channels = one 16-byte
// vector), keeps that group's per-channel constants in registers and processes row PAIRS (rows 2P, 2P+1: one Philox
// block covers a pair), kPairUnroll pairs at a time with all loads issued before the first use. A block covers
// kEwGroups channel groups x kEwLanes pair lanes; in one iteration it touches 2 * kEwLanes * kPairUnroll ADJACENT rows,
// i.e. one contiguous span of HBM per channel slab.
constexpr int kEwGroups = 128;   // 1024 channels per block slab
constexpr int kEwLanes = 2;
constexpr int kRedLanes = 4;     // reductions: 512-thread blocks, ~1 per SM, so few double atomics reach L2
constexpr int kPairUnroll = 2;   // 4 rows in flight per thread
constexpr int kRowsPerIter = 2 * kEwLanes * kPairUnroll;

__device__ __forceinline__ void load8(const float* p, int grp, float (&o)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p) + 2 * grp);
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 2 * grp + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

// rows handled by this thread in iteration `it`: pair P = (it * gridDim.x + blockIdx.x) * kEwLanes * kPairUnroll +
// u * kEwLanes + lane, rows 2P and 2P + 1
template <int LANES>
struct RowWalkT {
  int gl, lane, grp;
  __device__ __forceinline__ RowWalkT() {
    gl = threadIdx.x % kEwGroups;
    lane = threadIdx.x / kEwGroups;
    grp = blockIdx.y * kEwGroups + gl;
  }
  __device__ __forceinline__ long long pair(long long it, int u) const {
    return (it * gridDim.x + blockIdx.x) * (LANES * kPairUnroll) + u * LANES + lane;
  }
};
using RowWalk = RowWalkT<kEwLanes>;

// bn_finalize folded into the apply pass (fin.sum != nullptr): every thread derives the scale / shift of its 8 channels
// from the batch sums itself (16 doubles, the arithmetic of bn_finalize_kernel to the bit), and the threads of the first
// row block additionally publish scale / shift / mean / invstd for the backward and update the running statistics --
// one launch (and one dependent-launch gap) less per layer.
template <int DT>
__global__ void __launch_bounds__(kEwGroups * kEwLanes, 3)
bn_act_fwd_kernel(const uint4* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                  const uint4* __restrict__ res, long long rows, long long rows_per_seq, long long res_seq_rows,
                  int res_row_mul, int res_row_off, int groups, DropoutParams dp, uint4* __restrict__ a,
                  const BnFinalizeParams fin, uint8_t* __restrict__ keep_mask) {
  pdl_enter_long<3>(false);   // level 3 (A/B builds): the HBM-bound passes keep their dependents back too
  const RowWalk w;
  if (w.grp >= groups) return;
  const DropCtx drop = make_drop(dp);
  const bool res_flat = res_seq_rows == rows_per_seq * res_row_mul;
  float sc[8], sh[8];
  if (fin.sum != nullptr) {
    const bool publish = blockIdx.x == 0 && w.lane == 0;
    if (publish && w.grp == 0 && fin.nbt != nullptr) *fin.nbt += 1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = w.grp * 8 + k;
      BnChannel ch{0.f, 0.f, 0.f};
      sc[k] = sh[k] = 0.f;
      if (i < fin.c) {
        ch = bn_channel_stats(fin.sum[i], fin.sqsum[i], fin.inv_n, fin.eps);
        sc[k] = __fmul_rn(fin.gamma[i], ch.invstd);
        sh[k] = __fsub_rn(fin.beta[i], __fmul_rn(ch.mean, sc[k]));
        if (publish && fin.running_mean != nullptr) {
          fin.running_mean[i] = __fadd_rn(__fmul_rn(1.f - fin.momentum, fin.running_mean[i]),
                                          __fmul_rn(fin.momentum, ch.mean));
          fin.running_var[i] = __fadd_rn(__fmul_rn(1.f - fin.momentum, fin.running_var[i]),
                                         __fmul_rn(fin.momentum, __fmul_rn(ch.var, fin.unbias)));
        }
      }
      if (publish) {
        fin.scale_out[i] = sc[k];
        fin.shift_out[i] = sh[k];
        fin.mean_out[i] = ch.mean;
        fin.invstd_out[i] = ch.invstd;
      }
    }
  } else {
    load8(scale, w.grp, sc);
    load8(shift, w.grp, sh);
  }
  // Lean arithmetic (the pass was issue bound at ~18 instructions per element and 30 % of the HBM peak): the keep scale of
  // the dropout is folded into the affine (relu(x) k == relu(x k), k > 0); the dropout decision is a byte-wise carry test on
  // the Philox words (bit 7 of byte i: channel i kept); without a residual the ReLU is a modifier of the 16-bit conversion
  // and the mask an AND on the packed pairs, with a residual the fp32 value is masked by a sign-replicating prmt before the
  // add (one rounding, as before). The keep bits (ReLU passed AND kept) for the backward are the byte sign bits gathered
  // by a multiply.
  const float ks = drop.on ? drop.keep_scale : 1.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] *= ks;
    sh[k] *= ks;
  }
  const uint32_t kadd = ((256u - drop.thresh) & 0xFFu) * 0x01010101u, kadd7 = kadd & 0x7F7F7F7Fu;
  const long long iters = (rows + (long long)gridDim.x * kRowsPerIter - 1) / ((long long)gridDim.x * kRowsPerIter);
  for (long long it = 0; it < iters; ++it) {
    uint4 zv[2 * kPairUnroll], rv[2 * kPairUnroll];
#pragma unroll
    for (int u = 0; u < 2 * kPairUnroll; ++u) {
      const long long row = 2 * w.pair(it, u >> 1) + (u & 1);
      if (row < rows) {
        zv[u] = __ldg(z + row * groups + w.grp);
        if (res != nullptr) {
          // residual row of output row (seq, t): seq * res_seq_rows + t * res_row_mul + res_row_off. When a residual
          // sequence is exactly res_row_mul x an output sequence (the strided 1f blocks) that is row * res_row_mul +
          // res_row_off -- no 64-bit division per row (it was ~6 instructions per element of this pass)
          long long rrow;
          if (res_flat) {
            rrow = row * res_row_mul + res_row_off;
          } else if (rows <= 0xFFFFFFFFll) {
            const unsigned int seq = (unsigned int)row / (unsigned int)rows_per_seq;
            const unsigned int t = (unsigned int)row - seq * (unsigned int)rows_per_seq;
            rrow = (long long)seq * res_seq_rows + (long long)t * res_row_mul + res_row_off;
          } else {
            const long long seq = row / rows_per_seq;
            rrow = seq * res_seq_rows + (row - seq * rows_per_seq) * res_row_mul + res_row_off;
          }
          rv[u] = __ldg(res + rrow * groups + w.grp);
        }
      }
    }
#pragma unroll
    for (int pu = 0; pu < kPairUnroll; ++pu) {
      const long long P = w.pair(it, pu);
      if (2 * P >= rows) continue;
      uint4 bits = make_uint4(0, 0, 0, 0);
      if (drop.on) bits = drop_bits(drop, P, w.grp);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long row = 2 * P + h;
        if (row < rows) {
          float v[8];
          unpack8<DT>(zv[2 * pu + h], v);
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], sc[k], sh[k]);
          // bit 7 of byte i of kl / kh: channel i / 4 + i kept by the dropout (byte >= thresh <=> byte + 256 - thresh carries)
          uint32_t kl = 0x80808080u, kh = 0x80808080u;
          if (drop.on) {
            const uint32_t w0 = h ? bits.z : bits.x, w1 = h ? bits.w : bits.y;
            kl = lop3_maj(w0, kadd, (w0 & 0x7F7F7F7Fu) + kadd7);
            kh = lop3_maj(w1, kadd, (w1 & 0x7F7F7F7Fu) + kadd7);
          }
          uint4 out;
          if (res == nullptr) {
            out.x = pack2_relu<DT>(v[0], v[1]) & prmt_sign(kl, 0x9988u);
            out.y = pack2_relu<DT>(v[2], v[3]) & prmt_sign(kl, 0xBBAAu);
            out.z = pack2_relu<DT>(v[4], v[5]) & prmt_sign(kh, 0x9988u);
            out.w = pack2_relu<DT>(v[6], v[7]) & prmt_sign(kh, 0xBBAAu);
          } else {
            float r[8];
            unpack8<DT>(rv[2 * pu + h], r);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              // sign of byte (k & 3) replicated over the word: all ones where the channel is kept
              const uint32_t m = prmt_sign(k < 4 ? kl : kh, 0x8888u + 0x1111u * (k & 3));
              r[k] += __uint_as_float(__float_as_uint(fmaxf(v[k], 0.f)) & m);
            }
            out = pack8<DT>(r);
          }
          if (keep_mask != nullptr) {
            // ReLU passed <=> sign bit of the pre-activation clear (and not -0): top bytes of the 8 values gathered into two
            // words, cleared from the dropout keep bits, then the four byte sign bits of each word -> one nibble
            const uint32_t sl = prmt2(prmt2(__float_as_uint(v[0]), __float_as_uint(v[1]), 0x0073u),
                                      prmt2(__float_as_uint(v[2]), __float_as_uint(v[3]), 0x0073u), 0x5410u);
            const uint32_t sh_ = prmt2(prmt2(__float_as_uint(v[4]), __float_as_uint(v[5]), 0x0073u),
                                       prmt2(__float_as_uint(v[6]), __float_as_uint(v[7]), 0x0073u), 0x5410u);
            const uint32_t fl = kl & ~sl & 0x80808080u, fh = kh & ~sh_ & 0x80808080u;
            keep_mask[row * groups + w.grp] = (uint8_t)(((fl * 0x00204081u) >> 28) | (((fh * 0x00204081u) >> 28) << 4));
          }
          a[row * groups + w.grp] = out;
        }
      }
    }
  }
}

// Reductions: the pair lanes of a block are combined in shared memory, then one double atomic per channel per block
// reaches L2; the grid is capped (reduce_grid) because same-address atomics serialise.
__device__ __forceinline__ void block_combine_and_add(const float (&a1)[8], const float (&a2)[8], int gl, int lane, int grp,
                                                      int groups, double* __restrict__ out1, double* __restrict__ out2) {
  __shared__ float part[2][kRedLanes][kEwGroups * 8 + 4];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    part[0][lane][gl * 8 + k] = a1[k];
    part[1][lane][gl * 8 + k] = a2[k];
  }
  __syncthreads();
  if (lane == 0 && grp < groups) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int r = 0; r < kRedLanes; ++r) {
        s1 += (double)part[0][r][gl * 8 + k];
        s2 += (double)part[1][r][gl * 8 + k];
      }
      atomicAdd(out1 + grp * 8 + k, s1);
      atomicAdd(out2 + grp * 8 + k, s2);
    }
  }
}

// per-channel sum / sum of squares of a stored matrix (train-mode BatchNorm statistics of layers whose GEMM is too
// short to hide the in-epilogue reduction)
template <int DT>
__global__ void __launch_bounds__(kEwGroups * kRedLanes)
col_stats_kernel(const uint4* __restrict__ z, long long rows, int groups, double* __restrict__ sum,
                 double* __restrict__ sqsum) {
  pdl_enter();
  constexpr int kU = 8;  // rows in flight per thread; a block reads kRedLanes * kU adjacent rows per iteration
  const int gl = threadIdx.x % kEwGroups, lane = threadIdx.x / kEwGroups;
  const int grp = blockIdx.y * kEwGroups + gl;
  float a1[8], a2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a1[k] = a2[k] = 0.f;
  if (grp < groups) {
    for (long long base = (long long)blockIdx.x * (kRedLanes * kU); base < rows;
         base += (long long)gridDim.x * (kRedLanes * kU)) {
      uint4 zv[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const long long row = base + u * kRedLanes + lane;
        zv[u] = row < rows ? __ldg(z + row * groups + grp) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        float v[8];
        unpack8<DT>(zv[u], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          a1[k] += v[k];
          a2[k] = fmaf(v[k], v[k], a2[k]);
        }
      }
    }
  }
  block_combine_and_add(a1, a2, gl, lane, grp, groups, sum, sqsum);
}

// ---------------------------------------------------------------------------------------------- backward
// dy = g * dropout multiplier * [z * scale + shift > 0];  xhat = (z - mean) * invstd
template <int DT>
__device__ __forceinline__ void dy_xhat(const uint4& gu, const uint4& zu, const DropCtx& drop, uint32_t w0, uint32_t w1,
                                        const float (&sc)[8], const float (&sh)[8], const float (&mu)[8],
                                        const float (&is)[8], float (&dy)[8], float (&xh)[8]) {
  float gv[8], zv[8], m[8];
  unpack8<DT>(gu, gv);
  unpack8<DT>(zu, zv);
  if (drop.on) {
    drop_mult8(drop, w0, w1, m);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = 1.f;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const bool on = fmaf(zv[k], sc[k], sh[k]) > 0.f;
    dy[k] = on ? gv[k] * m[k] : 0.f;
    xh[k] = (zv[k] - mu[k]) * is[k];
  }
}

template <int DT>
__global__ void __launch_bounds__(kEwGroups * kRedLanes, 1)
bn_act_bwd_reduce_kernel(const uint4* __restrict__ g, const uint4* __restrict__ z, const float* __restrict__ scale,
                         const float* __restrict__ shift, const float* __restrict__ mean,
                         const float* __restrict__ invstd, long long rows, int groups, DropoutParams dp,
                         double* __restrict__ sum_dy, double* __restrict__ sum_dy_xhat) {
  pdl_enter_long<3>(false);   // level 3 (A/B builds): the HBM-bound passes keep their dependents back too
  const RowWalkT<kRedLanes> w;
  float a1[8], a2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a1[k] = a2[k] = 0.f;
  if (w.grp < groups) {
    const DropCtx drop = make_drop(dp);
    float sc[8], sh[8], mu[8], is[8];
    load8(scale, w.grp, sc);
    load8(shift, w.grp, sh);
    load8(mean, w.grp, mu);
    load8(invstd, w.grp, is);
    const long long iters = (rows + (long long)gridDim.x * (2 * kRedLanes * kPairUnroll) - 1) / ((long long)gridDim.x * (2 * kRedLanes * kPairUnroll));
    for (long long it = 0; it < iters; ++it) {
      uint4 gv[2 * kPairUnroll], zv[2 * kPairUnroll];
#pragma unroll
      for (int u = 0; u < 2 * kPairUnroll; ++u) {
        const long long row = 2 * w.pair(it, u >> 1) + (u & 1);
        if (row < rows) {
          gv[u] = __ldg(g + row * groups + w.grp);
          zv[u] = __ldg(z + row * groups + w.grp);
        }
      }
#pragma unroll
      for (int pu = 0; pu < kPairUnroll; ++pu) {
        const long long P = w.pair(it, pu);
        if (2 * P >= rows) continue;
        uint4 bits = make_uint4(0, 0, 0, 0);
        if (drop.on) bits = drop_bits(drop, P, w.grp);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (2 * P + h < rows) {
            float dy[8], xh[8];
            dy_xhat<DT>(gv[2 * pu + h], zv[2 * pu + h], drop, h ? bits.z : bits.x, h ? bits.w : bits.y, sc, sh, mu, is,
                        dy, xh);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              a1[k] += dy[k];
              a2[k] = fmaf(dy[k], xh[k], a2[k]);
            }
          }
        }
      }
    }
  }
  block_combine_and_add(a1, a2, w.gl, w.lane, w.grp, groups, sum_dy, sum_dy_xhat);
}

template <int DT>
__global__ void __launch_bounds__(kEwGroups * kEwLanes, 2)
bn_act_bwd_apply_kernel(const uint4* __restrict__ g, const uint4* __restrict__ z, const float* __restrict__ scale,
                        const float* __restrict__ shift, const float* __restrict__ mean,
                        const float* __restrict__ invstd, long long rows, long long count, int c, int groups,
                        DropoutParams dp, const double* __restrict__ sum_dy, const double* __restrict__ sum_dy_xhat,
                        const float* __restrict__ gscale_buf, uint4* __restrict__ dz, float* __restrict__ d_gamma,
                        float* __restrict__ d_beta) {
  pdl_enter_long<3>(false);   // level 3 (A/B builds): the HBM-bound passes keep their dependents back too
  const RowWalk w;
  if (w.grp >= groups) return;
  const DropCtx drop = make_drop(dp);
  const double inv_n = 1.0 / (double)count;
  float sc[8], sh[8], mu[8], is[8], m1[8], m2[8];
  load8(scale, w.grp, sc);
  load8(shift, w.grp, sh);
  load8(mean, w.grp, mu);
  load8(invstd, w.grp, is);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    m1[k] = (float)(sum_dy[w.grp * 8 + k] * inv_n);
    m2[k] = (float)(sum_dy_xhat[w.grp * 8 + k] * inv_n);
  }
  // BatchNorm parameter gradients (un-scaled): written once, by pair lane 0 of the blocks of grid row 0
  if (blockIdx.x == 0 && w.lane == 0 && d_gamma != nullptr) {
    const double inv = gscale_buf != nullptr ? (double)gscale_buf[1] : 1.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = w.grp * 8 + k;
      if (ch < c) {
        d_gamma[ch] = (float)(sum_dy_xhat[ch] * inv);
        d_beta[ch] = (float)(sum_dy[ch] * inv);
      }
    }
  }
  const long long iters = (rows + (long long)gridDim.x * kRowsPerIter - 1) / ((long long)gridDim.x * kRowsPerIter);
  for (long long it = 0; it < iters; ++it) {
    uint4 gv[2 * kPairUnroll], zv[2 * kPairUnroll];
#pragma unroll
    for (int u = 0; u < 2 * kPairUnroll; ++u) {
      const long long row = 2 * w.pair(it, u >> 1) + (u & 1);
      if (row < rows) {
        gv[u] = __ldg(g + row * groups + w.grp);
        zv[u] = __ldg(z + row * groups + w.grp);
      }
    }
#pragma unroll
    for (int pu = 0; pu < kPairUnroll; ++pu) {
      const long long P = w.pair(it, pu);
      if (2 * P >= rows) continue;
      uint4 bits = make_uint4(0, 0, 0, 0);
      if (drop.on) bits = drop_bits(drop, P, w.grp);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long row = 2 * P + h;
        if (row < rows) {
          float dy[8], xh[8], o[8];
          dy_xhat<DT>(gv[2 * pu + h], zv[2 * pu + h], drop, h ? bits.z : bits.x, h ? bits.w : bits.y, sc, sh, mu, is, dy,
                      xh);
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = sc[k] * (dy[k] - m1[k] - xh[k] * m2[k]);
          dz[row * groups + w.grp] = pack8<DT>(o);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- backward, stored mask
// The forward can store the keep decision (ReLU passed AND not dropped) as one bit per element (vp3d_bn_act_fwd_mask:
// 1/16 of the 16-bit activation). With it the two backward passes need neither the Philox stream nor the affine
// comparison, and the rest folds into per-channel constants:
//   reduce   gm = keep ? g : 0;   S1 += gm,  S2 += gm * (z - mean)        sum_dy = ks S1, sum_dy_xhat = ks invstd S2
//   apply    dz = A gm + C z + B  with A = scale ks, C = -scale invstd m2, B = -scale m1 - C mean
//            (= scale * (ks gm - m1 - (z - mean) invstd m2), m1 / m2 = the batch means of dy / dy xhat)
// The mask is applied to the PACKED gradient: byte -> two words whose byte sign bits are the 8 keep bits (one multiply
// each), prmt with sign replication -> 16-bit lane masks, one AND per pair. ~7 instructions per element instead of 22-26
// (ncu: the recomputing kernels are issue bound at 40 % of the HBM peak), so both passes run at memory speed.
template <int DT>
__device__ __forceinline__ void masked_unpack8(const uint4& gu, uint32_t keep_byte, float (&gm)[8]) {
  // bit i of the low / high nibble -> bit 8 i + 7 (the sign bit of byte i); the four shifted copies do not overlap
  const uint32_t lo = (keep_byte & 0xFu) * 0x10204080u, hi = (keep_byte >> 4) * 0x10204080u;
  const uint4 m = make_uint4(gu.x & prmt_sign(lo, 0x9988u), gu.y & prmt_sign(lo, 0xBBAAu), gu.z & prmt_sign(hi, 0x9988u),
                             gu.w & prmt_sign(hi, 0xBBAAu));
  unpack8<DT>(m, gm);
}

template <int DT>
__global__ void __launch_bounds__(kEwGroups * kRedLanes, 1)
bn_act_bwd_reduce_mask_kernel(const uint4* __restrict__ g, const uint4* __restrict__ z, const uint8_t* __restrict__ keep,
                              const float* __restrict__ mean, const float* __restrict__ invstd, float keep_scale,
                              long long rows, int groups, double* __restrict__ sum_dy,
                              double* __restrict__ sum_dy_xhat) {
  pdl_enter_long<3>(false);
  const RowWalkT<kRedLanes> w;
  float a1[8], a2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a1[k] = a2[k] = 0.f;
  if (w.grp < groups) {
    float mu[8], is[8];
    load8(mean, w.grp, mu);
    load8(invstd, w.grp, is);
    const long long iters = (rows + (long long)gridDim.x * (2 * kRedLanes * kPairUnroll) - 1) / ((long long)gridDim.x * (2 * kRedLanes * kPairUnroll));
    for (long long it = 0; it < iters; ++it) {
      uint4 gv[2 * kPairUnroll], zv[2 * kPairUnroll];
      uint32_t kb[2 * kPairUnroll];
#pragma unroll
      for (int u = 0; u < 2 * kPairUnroll; ++u) {
        const long long row = 2 * w.pair(it, u >> 1) + (u & 1);
        if (row < rows) {
          gv[u] = __ldg(g + row * groups + w.grp);
          zv[u] = __ldg(z + row * groups + w.grp);
          kb[u] = __ldg(keep + row * groups + w.grp);
        }
      }
#pragma unroll
      for (int u = 0; u < 2 * kPairUnroll; ++u) {
        const long long row = 2 * w.pair(it, u >> 1) + (u & 1);
        if (row < rows) {
          float gm[8], zf[8];
          masked_unpack8<DT>(gv[u], kb[u], gm);
          unpack8<DT>(zv[u], zf);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            a1[k] += gm[k];
            a2[k] = fmaf(gm[k], zf[k] - mu[k], a2[k]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a1[k] *= keep_scale;
      a2[k] *= keep_scale * is[k];
    }
  }
  block_combine_and_add(a1, a2, w.gl, w.lane, w.grp, groups, sum_dy, sum_dy_xhat);
}

template <int DT>
__global__ void __launch_bounds__(kEwGroups * kEwLanes, 2)
bn_act_bwd_apply_mask_kernel(const uint4* __restrict__ g, const uint4* __restrict__ z, const uint8_t* __restrict__ keep,
                             const float* __restrict__ scale, const float* __restrict__ mean,
                             const float* __restrict__ invstd, float keep_scale, long long rows, long long count, int c,
                             int groups, const double* __restrict__ sum_dy, const double* __restrict__ sum_dy_xhat,
                             const float* __restrict__ gscale_buf, uint4* __restrict__ dz, float* __restrict__ d_gamma,
                             float* __restrict__ d_beta) {
  pdl_enter_long<3>(false);
  const RowWalk w;
  if (w.grp >= groups) return;
  const double inv_n = 1.0 / (double)count;
  float sc[8], mu[8], is[8], A[8], B[8], Cc[8];
  load8(scale, w.grp, sc);
  load8(mean, w.grp, mu);
  load8(invstd, w.grp, is);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float m1 = (float)(sum_dy[w.grp * 8 + k] * inv_n);
    const float m2 = (float)(sum_dy_xhat[w.grp * 8 + k] * inv_n);
    A[k] = sc[k] * keep_scale;
    Cc[k] = -sc[k] * is[k] * m2;
    B[k] = -sc[k] * m1 - Cc[k] * mu[k];
  }
  // BatchNorm parameter gradients (un-scaled): written once, by pair lane 0 of the blocks of grid row 0
  if (blockIdx.x == 0 && w.lane == 0 && d_gamma != nullptr) {
    const double inv = gscale_buf != nullptr ? (double)gscale_buf[1] : 1.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = w.grp * 8 + k;
      if (ch < c) {
        d_gamma[ch] = (float)(sum_dy_xhat[ch] * inv);
        d_beta[ch] = (float)(sum_dy[ch] * inv);
      }
    }
  }
  const long long iters = (rows + (long long)gridDim.x * kRowsPerIter - 1) / ((long long)gridDim.x * kRowsPerIter);
  for (long long it = 0; it < iters; ++it) {
    uint4 gv[2 * kPairUnroll], zv[2 * kPairUnroll];
    uint32_t kb[2 * kPairUnroll];
#pragma unroll
    for (int u = 0; u < 2 * kPairUnroll; ++u) {
      const long long row = 2 * w.pair(it, u >> 1) + (u & 1);
      if (row < rows) {
        gv[u] = __ldg(g + row * groups + w.grp);
        zv[u] = __ldg(z + row * groups + w.grp);
        kb[u] = __ldg(keep + row * groups + w.grp);
      }
    }
#pragma unroll
    for (int u = 0; u < 2 * kPairUnroll; ++u) {
      const long long row = 2 * w.pair(it, u >> 1) + (u & 1);
      if (row < rows) {
        float gm[8], zf[8], o[8];
        masked_unpack8<DT>(gv[u], kb[u], gm);
        unpack8<DT>(zv[u], zf);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaf(A[k], gm[k], fmaf(Cc[k], zf[k], B[k]));
        dz[row * groups + w.grp] = pack8<DT>(o);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- gradient scale
__global__ void __launch_bounds__(256)
grad_absmax_kernel(const float* __restrict__ dy, long long n, float* __restrict__ gscale_buf) {
  pdl_enter();
  float m = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = fabsf(dy[i]);
    if (v == v) m = fmaxf(m, v);  // NaN-safe: a NaN gradient stays a NaN downstream, it must not poison the scale
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(gscale_buf + 2), __float_as_uint(m));
}
__global__ void grad_scale_finish_kernel(float* __restrict__ gscale_buf) {
  pdl_enter();
  const float m = gscale_buf[2];
  float s = 1.f;
  if (m > 0.f && m < 3.0e38f) {
    int e;
    frexpf(64.f / m, &e);  // 64 / m = f * 2^e, f in [0.5, 1)  ->  floor(log2) = e - 1
    e -= 1;
    if (e > 60) e = 60;
    if (e < -60) e = -60;
    s = ldexpf(1.f, e);
  }
  gscale_buf[0] = s;
  gscale_buf[1] = 1.f / s;
}

template <int DT>
__global__ void __launch_bounds__(256)
grad_pack_rows_kernel(const float* __restrict__ src, void* __restrict__ dst, long long rows, int c, int c_pad,
                      const float* __restrict__ gscale_buf, float* __restrict__ col_sum) {
  pdl_enter();
  const float gs = gscale_buf != nullptr ? gscale_buf[0] : 1.f;
  // thread = column (blockDim.x >= c_pad handled by stride), rows strided over blocks: coalesced in both src and dst
  for (int k = threadIdx.x; k < c_pad; k += blockDim.x) {
    float acc = 0.f;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
      const float v = k < c ? __ldg(src + r * c + k) : 0.f;
      acc += v;
      if (DT == VP3D_F16) static_cast<__half*>(dst)[r * c_pad + k] = __float2half_rn(v * gs);
      else static_cast<__nv_bfloat16*>(dst)[r * c_pad + k] = __float2bfloat16_rn(v * gs);
    }
    if (col_sum != nullptr && k < c) atomicAdd(col_sum + k, acc);
  }
}

// ---------------------------------------------------------------------------------------------- fused Adam + re-pack
// optim.Adam(amsgrad=True) (run.py:662,487) on one convolution weight, fused with the re-pack of the
// updated weight into the K-major 16-bit operand the next forward reads. One pass: p, g, m, v, vmax in (20 B), p, m, v,
// vmax out (16 B) + 2 B packed per parameter, instead of torch's multi-tensor Adam pass followed by a pack kernel that
// reads p again. Arithmetic follows torch.optim.adam (capturable / fused form):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  vmax = max(vmax, v)
//   p -= lr / (1 - b1^t) * m / (sqrt(vmax) / sqrt(1 - b2^t) + eps)
__device__ __forceinline__ void adam_update(const AdamParams& a, float step_size, float bc2_sqrt, float& p, float g,
                                            float& m, float& v, float& x) {
  if (a.maximize) g = -g;
  if (a.weight_decay != 0.f) g = fmaf(a.weight_decay, p, g);
  m = fmaf(a.beta1, m, (1.f - a.beta1) * g);      // lerp form of exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(a.beta2, v, (1.f - a.beta2) * g * g);
  float denom_v = v;
  if (a.vmax != nullptr) {
    x = fmaxf(x, v);
    denom_v = x;
  }
  p -= step_size * (m / (sqrtf(denom_v) / bc2_sqrt + a.eps));
}

template <int DT>
__device__ __forceinline__ void adam_pack_body(const AdamParams& a, const long long first, const long long stride) {
  const float t = *a.step;
  const float lr = a.lr_dev != nullptr ? *a.lr_dev : a.lr;
  const float bc1 = 1.f - powf(a.beta1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(a.beta2, t));
  const float step_size = lr / bc1;
  const int row_len = a.c_in * a.taps;           // elements per output channel in the nn.Conv1d layout
  const long long n4 = a.n >> 2;                 // float4 groups (n is a multiple of 4 for every conv weight here)
  for (long long q = first; q < n4; q += stride) {
    float4 p4 = reinterpret_cast<float4*>(a.p)[q];
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.g) + q);
    float4 m4 = reinterpret_cast<float4*>(a.m)[q];
    float4 v4 = reinterpret_cast<float4*>(a.v)[q];
    float4 x4 = a.vmax != nullptr ? reinterpret_cast<float4*>(a.vmax)[q] : make_float4(0, 0, 0, 0);
    float pp[4] = {p4.x, p4.y, p4.z, p4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w},
          vv[4] = {v4.x, v4.y, v4.z, v4.w}, xx[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) adam_update(a, step_size, bc2_sqrt, pp[e], gg[e], mm[e], vv[e], xx[e]);
    reinterpret_cast<float4*>(a.p)[q] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(a.m)[q] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(a.v)[q] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    if (a.vmax != nullptr) reinterpret_cast<float4*>(a.vmax)[q] = make_float4(xx[0], xx[1], xx[2], xx[3]);
    if (a.packed != nullptr) {
      // element i = (co, ci, tap) of the (c_out, c_in, taps) weight -> packed[co][tap * k_pad + ci]
      const long long i0 = 4 * q;
      int co = (int)(i0 / row_len);
      int rem = (int)(i0 - (long long)co * row_len);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ci = rem / a.taps, tap = rem - ci * a.taps;
        const long long d = ((long long)co * a.taps + tap) * a.k_pad + ci;
        if (DT == VP3D_F16) static_cast<__half*>(a.packed)[d] = __float2half_rn(pp[e]);
        else static_cast<__nv_bfloat16*>(a.packed)[d] = __float2bfloat16_rn(pp[e]);
        if (++rem == row_len) {
          rem = 0;
          ++co;
        }
      }
    }
  }
  // scalar tail (small tensors whose length is not a multiple of 4; they carry no packed operand)
  for (long long i = 4 * n4 + first; i < a.n; i += stride) {
    float pv = a.p[i], mv = a.m[i], vv1 = a.v[i], xv = a.vmax != nullptr ? a.vmax[i] : 0.f;
    adam_update(a, step_size, bc2_sqrt, pv, a.g[i], mv, vv1, xv);
    a.p[i] = pv;
    a.m[i] = mv;
    a.v[i] = vv1;
    if (a.vmax != nullptr) a.vmax[i] = xv;
  }
}

template <int DT>
__global__ void __launch_bounds__(256)
adam_pack_kernel(AdamParams a) {
  pdl_enter_long<3>(false);   // level 3 (A/B builds): the HBM-bound passes keep their dependents back too
  adam_pack_body<DT>(a, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

template <int DT>
__global__ void __launch_bounds__(256)
adam_multi_kernel(const __grid_constant__ AdamMultiParams mp) {
  pdl_enter_long<3>(false);   // level 3 (A/B builds): the HBM-bound passes keep their dependents back too
  int i = 0;
  while (i + 1 < mp.count && (int)blockIdx.x >= mp.block_start[i + 1]) ++i;   // <= 32 entries, uniform per block
  const AdamTensor& T = mp.t[i];
  AdamParams a;
  a.p = T.p; a.g = T.g; a.m = T.m; a.v = T.v; a.vmax = T.vmax;
  a.n = T.n;
  a.lr = mp.lr; a.beta1 = mp.beta1; a.beta2 = mp.beta2; a.eps = mp.eps; a.weight_decay = mp.weight_decay;
  a.step = T.step; a.lr_dev = mp.lr_dev;
  a.maximize = mp.maximize;
  a.packed = T.packed; a.c_in = T.c_in; a.taps = T.taps; a.k_pad = T.k_pad;
  const int b0 = mp.block_start[i], nb = mp.block_start[i + 1] - b0;
  adam_pack_body<DT>(a, (long long)(blockIdx.x - b0) * blockDim.x + threadIdx.x, (long long)nb * blockDim.x);
}

cudaError_t launch_adam_multi(int dtype, const AdamMultiParams& a, cudaStream_t stream) {
  const int blocks = a.block_start[a.count];
  if (blocks <= 0) return cudaSuccess;
  if (dtype == VP3D_BF16) launch_k(adam_multi_kernel<VP3D_BF16>, dim3(blocks), dim3(256), 0, stream, a);
  else launch_k(adam_multi_kernel<VP3D_F16>, dim3(blocks), dim3(256), 0, stream, a);
  return cudaGetLastError();
}

cudaError_t launch_adam_pack(int dtype, const AdamParams& a, int sm_count, cudaStream_t stream) {
  long long blocks = ((a.n + 3) / 4 + 255) / 256;
  if (blocks > (long long)sm_count * 8) blocks = (long long)sm_count * 8;
  if (blocks < 1) blocks = 1;
  if (dtype == VP3D_BF16) launch_k(adam_pack_kernel<VP3D_BF16>, dim3((int)blocks), dim3(256), 0, stream, a);
  else launch_k(adam_pack_kernel<VP3D_F16>, dim3((int)blocks), dim3(256), 0, stream, a);
  return cudaGetLastError();
}

__global__ void counter_add_kernel(unsigned long long* counter, unsigned long long inc) {
  pdl_enter();
  *counter += inc;
}
cudaError_t launch_counter_add(unsigned long long* counter, unsigned long long inc, cudaStream_t stream) {
  launch_k(counter_add_kernel, dim3(1), dim3(1), 0, stream, counter, inc);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- launchers
// grid for the row-walking kernels: y covers the channel slabs, x strides over spans of kRowsPerIter adjacent rows
static dim3 row_walk_grid(long long rows, int groups, int sm_count, int per_sm) {
  const int gy = (groups + kEwGroups - 1) / kEwGroups;
  long long gx = (rows + kRowsPerIter - 1) / kRowsPerIter;
  const long long cap = ((long long)sm_count * per_sm + gy - 1) / gy;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)gy);
}

// reductions: 512-thread blocks, at most ~1 per SM over all channel slabs
static dim3 reduce_grid(long long rows, int rows_per_iter, int groups, int sm_count) {
  const int gy = (groups + kEwGroups - 1) / kEwGroups;
  long long gx = (rows + rows_per_iter - 1) / rows_per_iter;
  const long long cap = (sm_count + gy - 1) / gy;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)gy);
}

cudaError_t launch_bn_finalize(const double* sum, const double* sqsum, long long count, const float* gamma,
                               const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                               long long* nbt, float* scale, float* shift, float* mean, float* invstd, int c, int c_pad,
                               cudaStream_t stream) {
  const double n = (double)count;
  launch_k(bn_finalize_kernel, dim3(1), dim3(c_pad >= 1024 ? 1024 : ((c_pad + 31) / 32) * 32), 0, stream, sum, sqsum, 1.0 / n,
                                                              count > 1 ? (float)(n / (n - 1.0)) : 1.f, gamma, beta, eps, momentum,
                                                              running_mean, running_var, nbt, scale, shift, mean,
                                                              invstd, c, c_pad);
  return cudaGetLastError();
}

#define VP3D_DISPATCH_16(KERNEL, ...)                                                              \
  if (dtype == VP3D_F16) launch_k(KERNEL<VP3D_F16>, dim3(grid), dim3(block), 0, stream, __VA_ARGS__);                 \
  else if (dtype == VP3D_BF16) launch_k(KERNEL<VP3D_BF16>, dim3(grid), dim3(block), 0, stream, __VA_ARGS__);          \
  else return cudaErrorInvalidValue;                                                                \
  return cudaGetLastError();

cudaError_t launch_bn_act_fwd(int dtype, const void* z, const float* scale, const float* shift, const void* res,
                              long long seqs, long long rows_per_seq, long long res_seq_rows, int res_row_mul,
                              int res_row_off, int c_pad, const DropoutParams& dp, void* a,
                              const BnFinalizeParams& fin, int sm_count, cudaStream_t stream, unsigned char* keep_mask) {
  const long long rows = seqs * rows_per_seq;
  const int groups = c_pad / 8;
  const dim3 grid = row_walk_grid(rows, groups, sm_count, 8);
  const int block = kEwGroups * kEwLanes;
  VP3D_DISPATCH_16(bn_act_fwd_kernel, static_cast<const uint4*>(z), scale, shift, static_cast<const uint4*>(res), rows,
                   rows_per_seq, res_seq_rows, res_row_mul, res_row_off, groups, dp, static_cast<uint4*>(a), fin, keep_mask)
}

cudaError_t launch_col_stats(int dtype, const void* z, long long rows, int c_pad, double* sum, double* sqsum,
                             int sm_count, cudaStream_t stream) {
  const int groups = c_pad / 8;
  const dim3 grid = reduce_grid(rows, kRedLanes * 8, groups, sm_count);
  const int block = kEwGroups * kRedLanes;
  VP3D_DISPATCH_16(col_stats_kernel, static_cast<const uint4*>(z), rows, groups, sum, sqsum)
}

cudaError_t launch_bn_act_bwd_reduce(int dtype, const void* g, const void* z, const float* scale, const float* shift,
                                     const float* mean, const float* invstd, long long rows, int c_pad,
                                     const DropoutParams& dp, double* sum_dy, double* sum_dy_xhat, int sm_count,
                                     cudaStream_t stream) {
  const int groups = c_pad / 8;
  const dim3 grid = reduce_grid(rows, 2 * kRedLanes * kPairUnroll, groups, sm_count);
  const int block = kEwGroups * kRedLanes;
  VP3D_DISPATCH_16(bn_act_bwd_reduce_kernel, static_cast<const uint4*>(g), static_cast<const uint4*>(z), scale, shift,
                   mean, invstd, rows, groups, dp, sum_dy, sum_dy_xhat)
}

cudaError_t launch_bn_act_bwd_apply(int dtype, const void* g, const void* z, const float* scale, const float* shift,
                                    const float* mean, const float* invstd, long long rows, long long count, int c,
                                    int c_pad, const DropoutParams& dp, const double* sum_dy, const double* sum_dy_xhat,
                                    const float* gscale_buf, void* dz, float* d_gamma, float* d_beta, int sm_count,
                                    cudaStream_t stream) {
  const int groups = c_pad / 8;
  const dim3 grid = row_walk_grid(rows, groups, sm_count, 8);
  const int block = kEwGroups * kEwLanes;
  VP3D_DISPATCH_16(bn_act_bwd_apply_kernel, static_cast<const uint4*>(g), static_cast<const uint4*>(z), scale, shift,
                   mean, invstd, rows, count, c, groups, dp, sum_dy, sum_dy_xhat, gscale_buf, static_cast<uint4*>(dz),
                   d_gamma, d_beta)
}

cudaError_t launch_bn_act_bwd_reduce_mask(int dtype, const void* g, const void* z, const unsigned char* keep,
                                          const float* mean, const float* invstd, float keep_scale, long long rows,
                                          int c_pad, double* sum_dy, double* sum_dy_xhat, int sm_count,
                                          cudaStream_t stream) {
  const int groups = c_pad / 8;
  const dim3 grid = reduce_grid(rows, 2 * kRedLanes * kPairUnroll, groups, sm_count);
  const int block = kEwGroups * kRedLanes;
  VP3D_DISPATCH_16(bn_act_bwd_reduce_mask_kernel, static_cast<const uint4*>(g), static_cast<const uint4*>(z), keep, mean,
                   invstd, keep_scale, rows, groups, sum_dy, sum_dy_xhat)
}

cudaError_t launch_bn_act_bwd_apply_mask(int dtype, const void* g, const void* z, const unsigned char* keep,
                                         const float* scale, const float* mean, const float* invstd, float keep_scale,
                                         long long rows, long long count, int c, int c_pad, const double* sum_dy,
                                         const double* sum_dy_xhat, const float* gscale_buf, void* dz, float* d_gamma,
                                         float* d_beta, int sm_count, cudaStream_t stream) {
  const int groups = c_pad / 8;
  const dim3 grid = row_walk_grid(rows, groups, sm_count, 8);
  const int block = kEwGroups * kEwLanes;
  VP3D_DISPATCH_16(bn_act_bwd_apply_mask_kernel, static_cast<const uint4*>(g), static_cast<const uint4*>(z), keep, scale,
                   mean, invstd, keep_scale, rows, count, c, groups, sum_dy, sum_dy_xhat, gscale_buf,
                   static_cast<uint4*>(dz), d_gamma, d_beta)
}

cudaError_t launch_grad_scale(const float* dy, long long n, float* gscale_buf, int sm_count, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(gscale_buf, 0, 3 * sizeof(float), stream);
  if (e != cudaSuccess) return e;
  long long blocks = (n + 255) / 256;
  if (blocks > sm_count * 4) blocks = sm_count * 4;
  if (blocks < 1) blocks = 1;
  launch_k(grad_absmax_kernel, dim3((int)blocks), dim3(256), 0, stream, dy, n, gscale_buf);
  launch_k(grad_scale_finish_kernel, dim3(1), dim3(1), 0, stream, gscale_buf);
  return cudaGetLastError();
}

cudaError_t launch_grad_pack_rows(int dtype, const float* src, void* dst, long long rows, int c, int c_pad,
                                  const float* gscale_buf, float* col_sum, int sm_count, cudaStream_t stream) {
  long long blocks = rows;
  if (blocks > sm_count * 8) blocks = sm_count * 8;
  if (blocks < 1) blocks = 1;
  const int threads = c_pad < 256 ? ((c_pad + 31) / 32) * 32 : 256;
  if (dtype == VP3D_F16)
    launch_k(grad_pack_rows_kernel<VP3D_F16>, dim3((int)blocks), dim3(threads), 0, stream, src, dst, rows, c, c_pad, gscale_buf, col_sum);
  else if (dtype == VP3D_BF16)
    launch_k(grad_pack_rows_kernel<VP3D_BF16>, dim3((int)blocks), dim3(threads), 0, stream, src, dst, rows, c, c_pad, gscale_buf, col_sum);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace vp3d
