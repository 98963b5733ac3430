// Low-latency streaming step of the causal TemporalModel for a FEW concurrent streams (BASELINE configs[3]; reference
// common/models/TemporalModel.py:126-138 evaluated one new frame at a time, see vp3d_b200/streaming.py for the rings).
//
// With S <= 8 streams a frame is ten matrix-VECTOR products (S rows x K <= 3072 against 1024 x K weights): 34 MB of
// 16-bit weights stream out of L2 once per frame and there is nothing for a tensor core to do with an M = 128 tile that
// has one live row. The batch path pays ~11 us per layer for a launch, a TMEM / TMA pipeline fill and an epilogue (12
// dependent launches = 0.13 ms per frame whatever S). Here the whole frame is ONE cooperative kernel:
//
//   phase 0   the new frame's 2-D keypoints -> both copies of ring 0's current slot (what vp3d_ring_write does)
//   phase l   y[s][c] = act(sum_k W_l[c][k] x_l[s][k] + shift_l[c]) (+ residual): the layer input x_l (S x K, gathered from
//             the ring rows of the layer's taps) is staged in shared memory (as fp32) by every CTA, ONE WARP OWNS ONE OUTPUT CHANNEL
//             (148 SMs x 8 warps >= 1024 channels): lanes stride over K with 16-byte loads of the weight row, fp32
//             accumulation of the exact fp16 / bf16 products, butterfly reduction, lane 0 applies shift / ReLU / residual
//             and stores the 16-bit result into the next ring (slot and mirror) -- the rounding points of the GEMM path
//   between phases: a grid barrier (one atomic per CTA on one of eight monotonic 64-bit counters + acquire spin,
//             bounded by %globaltimer); the weight rows of the NEXT phase are requested before the barrier, so their L2 latency
//             hides behind it.
//
// Ring positions come from the device frame counter exactly as in stream_advance_kernel (elementwise.cu); the kernel
// increments it at the end, so consecutive launches need no host-side state and no CUDA graph.
#include <cstdio>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "kernels.h"

namespace vp3d {
namespace {

constexpr int kStThreads = 256;
constexpr int kStWarps = kStThreads / 32;
constexpr int kStMaxK = 3072;           // taps * c_in_pad of the widest layer
constexpr int kStIters = kStMaxK / 256; // 16-byte weight loads per lane per channel

template <int DT>
__device__ __forceinline__ float2 unpack2(uint32_t u);
template <>
__device__ __forceinline__ float2 unpack2<VP3D_F16>(uint32_t u) {
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}
template <>
__device__ __forceinline__ float2 unpack2<VP3D_BF16>(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}
template <int DT>
__device__ __forceinline__ uint16_t pack1(float v) {
  if (DT == VP3D_F16) return __half_as_ushort(__float2half_rn(v));
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
template <int DT>
__device__ __forceinline__ float unpack1(uint16_t h) {
  if (DT == VP3D_F16) return __half2float(__ushort_as_half(h));
  return __uint_as_float(static_cast<uint32_t>(h) << 16);
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long st_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint4 ldcg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }

// Grid barrier of the (cooperative, co-resident) grid. Same-address atomics serialise in L2 at ~27 cycles each, so 128
// arrivals on one word would cost ~2 us; the CTAs arrive on kBarWords counters (one 128-byte line each, CTA b on word
// b mod kBarWords) and eight lanes of warp 0 poll one word each. `arrivals` = how many times every CTA has arrived in
// total once this barrier is complete (monotonic counters, never reset between frames).
constexpr int kBarWords = 8;
constexpr int kBarStride = 16;   // u64 words between counters (128 bytes)
constexpr int kBarEpochWord = 120;   // frames this barrier buffer has served (its own 128-byte line)
__device__ void grid_barrier(unsigned long long* counters, unsigned long long arrivals) {
  __syncthreads();
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(counters + (blockIdx.x % kBarWords) * kBarStride, 1ull);
    }
    if (threadIdx.x < kBarWords) {
      // CTAs on word w: blockIdx = w, w + kBarWords, ...
      const unsigned long long members = (gridDim.x - threadIdx.x + kBarWords - 1) / kBarWords;
      const unsigned long long target = arrivals * members;
      const unsigned long long t0 = st_globaltimer();
      while (ld_acquire_u64(counters + threadIdx.x * kBarStride) < target) {
        if (st_globaltimer() - t0 > 2000000000ull) {   // 2 s: a CTA that never arrives traps the kernel instead of hanging
          printf("vp3d: stream_step grid barrier timed out (block %d)\n", blockIdx.x);
          __trap();
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
}

// CH = output channels per warp (1 in every shipped instantiation, see launch_dt).
template <int DT, int S, int CH>
__global__ void __launch_bounds__(kStThreads, 1) stream_step_kernel(const StreamStepParams p) {
  extern __shared__ __align__(16) uint8_t st_smem[];
  float* xs = reinterpret_cast<float*>(st_smem);   // [S][K] of the current layer, fp32
  __shared__ int ring[16][4];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t = *p.step;     // frame index of this step: nobody writes it before the last phase
  if (threadIdx.x < p.n_rings) {
    const int i = threadIdx.x;
    const int L = p.ring_len[i];
    const int q = static_cast<int>(t % L) + L;   // upper copy of the current slot: taps q - k*d never wrap
    ring[i][0] = (q - (p.ring_taps[i] - 1) * p.ring_dil[i]) * p.rows_per_slot;
    ring[i][1] = q * p.rows_per_slot;
    ring[i][2] = q * p.rows_per_slot;
    ring[i][3] = (q - L) * p.rows_per_slot;
  }
  __syncthreads();
  // The barrier counters are monotonic; how many frames they have seen is kept beside them (word kBarEpochWord), not taken
  // from the frame counter, so frames advanced by the GEMM path (vp3d_stream_advance) in between do no harm. Read here by
  // every CTA, written by CTA 0 after the frame's first barrier.
  const unsigned long long epoch = p.barrier[kBarEpochWord];
  const unsigned long long bar_base = epoch * p.n_layers;   // n_layers barriers per frame
  int bar = 0;

  // weight rows (and the epilogue's shift / residual values) of the first layer while phase 0 runs
  const int gwarp = (blockIdx.x * kStWarps + warp) * CH;        // first channel of this warp
  const int total_warps = gridDim.x * kStWarps * CH;            // channels covered by one pass of the grid
  uint4 wreg[CH][kStIters];
  float shreg[CH], resreg[CH];   // lane s < n_streams: residual of stream s; every lane: the channel's shift
  // Everything a layer needs that does NOT depend on the previous layer's output: its weight row, shift, and the residual
  // rows (the block input, written two barriers earlier). Requested BEFORE the barrier in front of the layer.
  auto request = [&](const StreamLayer& L, int c) {
    const int K = L.taps * L.k_per_tap;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int cj = c + j;
#pragma unroll
      for (int it = 0; it < kStIters; ++it) {
        const int k0 = it * 256 + lane * 8;
        wreg[j][it] = make_uint4(0u, 0u, 0u, 0u);
        if (cj < L.n && k0 < K)
          wreg[j][it] = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(L.w) + (size_t)cj * K + k0));
      }
      shreg[j] = (L.shift != nullptr && cj < L.n) ? __ldg(L.shift + cj) : 0.f;
      resreg[j] = 0.f;
      if (L.res != nullptr && cj < L.n && lane < p.n_streams) {
        const int res_off = L.res_ring >= 0 ? ring[L.res_ring][1] : 0;
        resreg[j] = unpack1<DT>(__ldcg(static_cast<const uint16_t*>(L.res) + (size_t)(res_off + lane) * L.res_row_stride + cj));
      }
    }
  };
  request(p.layers[0], gwarp);

  // ---- phase 0: new frame -> ring 0 (slot and mirror), columns >= c_in zero
  {
    const int total = p.n_streams * p.c_in_pad;
    uint16_t* r0 = static_cast<uint16_t*>(p.ring0);
    for (int i = blockIdx.x * kStThreads + threadIdx.x; i < total; i += gridDim.x * kStThreads) {
      const int s = i / p.c_in_pad, c = i - s * p.c_in_pad;
      const uint16_t v = c < p.c_in ? pack1<DT>(p.x_in[s * p.c_in + c]) : (uint16_t)0;
      r0[(size_t)(ring[0][1] + s) * p.c_in_pad + c] = v;
      r0[(size_t)(ring[0][3] + s) * p.c_in_pad + c] = v;
    }
  }
  grid_barrier(p.barrier, bar_base + static_cast<unsigned long long>(++bar));

  for (int l = 0; l < p.n_layers; ++l) {
    const StreamLayer& L = p.layers[l];
    const int K = L.taps * L.k_per_tap;
    const int a_off = L.a_ring >= 0 ? ring[L.a_ring][0] : 0;
    // layer input of every stream -> shared memory AS FP32 (converted once here instead of once per output channel);
    // 16-byte pieces of the 16-bit rows, rows of the taps are tap_row_step apart
    const int pieces_per_tap = L.k_per_tap / 8;
    const int pieces = S * L.taps * pieces_per_tap;
    for (int i = threadIdx.x; i < pieces; i += kStThreads) {
      const int s = i / (L.taps * pieces_per_tap);
      const int r = i - s * (L.taps * pieces_per_tap);
      const int tap = r / pieces_per_tap, pc = r - tap * pieces_per_tap;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (s < p.n_streams)
        v = ldcg16(static_cast<const uint16_t*>(L.a) + (size_t)(a_off + tap * L.tap_row_step + s) * L.k_per_tap + pc * 8);
      const float2 f0 = unpack2<DT>(v.x), f1 = unpack2<DT>(v.y), f2 = unpack2<DT>(v.z), f3 = unpack2<DT>(v.w);
      // the two halves of a piece go to two arrays, so that a warp's 16-byte reads are 16 bytes apart (no bank conflict)
      const int j = (tap * L.k_per_tap + pc * 8) >> 3;
      reinterpret_cast<float4*>(xs + (size_t)s * K)[j] = make_float4(f0.x, f0.y, f1.x, f1.y);
      reinterpret_cast<float4*>(xs + (size_t)s * K + (K >> 1))[j] = make_float4(f2.x, f2.y, f3.x, f3.y);
    }
    __syncthreads();

    for (int c = gwarp; c < L.n; c += total_warps) {
      if (c != gwarp) request(L, c);    // (more channels than warps: only with small grids)
      float acc[CH][S];
#pragma unroll
      for (int j = 0; j < CH; ++j)
#pragma unroll
        for (int s = 0; s < S; ++s) acc[j][s] = 0.f;
#pragma unroll
      for (int it = 0; it < kStIters; ++it) {
        const int k0 = it * 256 + lane * 8;
        if (k0 < K) {
          float2 w[CH][4];
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            w[j][0] = unpack2<DT>(wreg[j][it].x); w[j][1] = unpack2<DT>(wreg[j][it].y);
            w[j][2] = unpack2<DT>(wreg[j][it].z); w[j][3] = unpack2<DT>(wreg[j][it].w);
          }
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const float4 xa = reinterpret_cast<const float4*>(xs + (size_t)s * K)[k0 >> 3];
            const float4 xb = reinterpret_cast<const float4*>(xs + (size_t)s * K + (K >> 1))[k0 >> 3];
#pragma unroll
            for (int j = 0; j < CH; ++j) {
              float a = acc[j][s];
              a = fmaf(w[j][0].x, xa.x, a); a = fmaf(w[j][0].y, xa.y, a); a = fmaf(w[j][1].x, xa.z, a); a = fmaf(w[j][1].y, xa.w, a);
              a = fmaf(w[j][2].x, xb.x, a); a = fmaf(w[j][2].y, xb.y, a); a = fmaf(w[j][3].x, xb.z, a); a = fmaf(w[j][3].y, xb.w, a);
              acc[j][s] = a;
            }
          }
        }
      }
      // butterfly: every lane ends with every stream's total; lane s finishes stream s (its residual is in resreg)
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int cj = c + j;
        float mine = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc[j][s] += __shfl_xor_sync(0xffffffffu, acc[j][s], o);
          if (lane == s) mine = acc[j][s];
        }
        if (lane < p.n_streams && cj < L.n) {
          const int out_off = L.out_ring >= 0 ? ring[L.out_ring][2] : 0;
          const int out_off2 = L.out_ring >= 0 ? ring[L.out_ring][3] : -1;
          float v = mine + shreg[j];
          if (L.relu) v = fmaxf(v, 0.f);
          v += resreg[j];
          if (L.out_f32) {
            if (cj < L.n_valid) static_cast<float*>(L.out)[(size_t)lane * L.out_row_stride + cj] = v;
          } else {
            const uint16_t h = pack1<DT>(v);
            uint16_t* o = static_cast<uint16_t*>(L.out);
            o[(size_t)(out_off + lane) * L.out_row_stride + cj] = h;
            if (out_off2 >= 0) o[(size_t)(out_off2 + lane) * L.out_row_stride + cj] = h;
          }
        }
      }
    }
    if (l + 1 < p.n_layers) {
      request(p.layers[l + 1], gwarp);
      grid_barrier(p.barrier, bar_base + static_cast<unsigned long long>(++bar));
    }
  }
  // every CTA read *step before it arrived at the frame's first barrier, which this thread has passed
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *p.step = t + 1;
    p.barrier[kBarEpochWord] = epoch + 1;
  }
}

template <int DT, int S, int CH>
cudaError_t launch_s(const StreamStepParams& p, int grid, cudaStream_t stream) {
  const size_t smem = (size_t)S * kStMaxK * 4;
  auto kernel = stream_step_kernel<DT, S, CH>;
  static std::atomic<unsigned long long> attr_done{0};
  if (cudaError_t e = set_max_smem_once(reinterpret_cast<const void*>(kernel), (int)smem, attr_done)) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kStThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // co-residency of the grid (it spins on a grid barrier)
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, p);
}

// One channel per warp for every stream count. Two channels per warp (half the grid, the staged input read once for both)
// were measured slower: 4 streams 0.055 -> 0.066 ms, 8 streams 0.077 -> 0.093 ms per frame -- with several streams the
// layer time is FMA issue and weight-pull parallelism, not shared-memory reads.
template <int DT>
cudaError_t launch_dt(const StreamStepParams& p, int widest, int sm_count, cudaStream_t stream) {
  auto grid_for = [&](int ch) {
    int g = (widest + kStWarps * ch - 1) / (kStWarps * ch);
    return g > sm_count ? sm_count : g;
  };
  if (p.n_streams <= 1) return launch_s<DT, 1, 1>(p, grid_for(1), stream);
  if (p.n_streams <= 2) return launch_s<DT, 2, 1>(p, grid_for(1), stream);
  if (p.n_streams <= 4) return launch_s<DT, 4, 1>(p, grid_for(1), stream);
  return launch_s<DT, 8, 1>(p, grid_for(1), stream);
}

}  // namespace

int stream_step_max_streams() { return 8; }
int stream_step_max_k() { return kStMaxK; }

cudaError_t launch_stream_step(int dtype, const StreamStepParams& p, int sm_count, cudaStream_t stream) {
  // one warp per output channel of the widest layer, at most one CTA per SM (all resident at once: cooperative launch);
  // no more CTAs than that, every CTA is a participant of ten grid barriers per frame
  int widest = 1;
  for (int l = 0; l < p.n_layers; ++l) widest = p.layers[l].n > widest ? p.layers[l].n : widest;
  if (dtype == VP3D_BF16) return launch_dt<VP3D_BF16>(p, widest, sm_count, stream);
  return launch_dt<VP3D_F16>(p, widest, sm_count, stream);
}

}  // namespace vp3d
