// K4: weight gradient of the temporal convolutions as a tcgen05 / TMEM GEMM whose reduction runs over frames (sm_100a).
//
// What it replaces in the reference: the autograd weight gradient of every nn.Conv1d of the stack
// (common/models/TemporalModel.py:102,113-118 dilated, :168-181 strided; backward triggered at run.py:485):
//
//   dW[tap][co][ci] = sum_{seq, r}  dz[seq][r][co] * a[seq][r + tap * dilation][ci]          (dilated model)
//   dW[tap][co][ci] = sum_{r}       dz[r][co]      * a_view[r][tap * c_in + ci]              (stride == width, 1f model)
//
// Both operands live channels-last in HBM, i.e. the reduction index (the frame) is the *slow* index of both: they are
// "MN-major" tcgen05 operands. TMA fetches [64 frames x 128 bytes of channels] boxes with SWIZZLE_128B; the shared
// memory descriptor walks them with LBO = one box (next 64-channel group) and SBO = 1024 B (next 8 frames), so no
// transpose pass over HBM is needed.
//
// Work decomposition (split-K in lockstep): a work item is (row slice s, output tile 128 co x BN ci of one tap). The
// host picks the number of slices S so that tiles * S fills a whole number of waves of the persistent grid (one CTA
// per SM); item i = s * tiles + t goes to CTA i mod grid. All CTAs of a wave walk the same row range at the same
// pace, so every dz / a box that 8-12 tiles need is fetched from HBM once and hit in L2 by the others. A CTA
// accumulates its item in TMEM and adds the partial tile to the fp32 result with vector reductions
// (red.global.add.v4.f32); with S slices each output element receives S reductions.
//
// Pipeline per CTA (192 threads), as in conv_gemm.cu: warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue with
// two TMEM accumulator buffers so that the reduction of run i overlaps the MMAs of run i + 1.
#include "ptx.cuh"
#include "kernels.h"
#include "pdl.cuh"
#include "pdl.cuh"

namespace vp3d {

constexpr int kWgRows = 64;       // frames (reduction depth) per pipeline stage
constexpr int kWgThreads = 224;   // TMA producer, MMA issuer, 4 epilogue warps, tile scheduler
constexpr int kWgSchedWarp = 6;
constexpr int kWgClcSlots = 2;
constexpr int kWgClcConsumers = 1 + 1 + 4;   // producer, MMA thread, epilogue warps

// BM = output channels per tile. BM = 256 (two 128-row MMAs sharing the B tile, all 512 TMEM columns as ONE accumulator
// set) is the wide-layer configuration: both operands of a weight gradient stream from L2 / HBM with no reuse inside a
// CTA, so the mainloop is latency bound by bytes per flop and bytes in flight -- a 256 x 256 tile moves 64 KB per 1024
// MMA cycles (128 x 256: 48 KB per 512) and its three stages cover 3072 MMA cycles of latency instead of 2048. The
// price, no accumulator double buffering, is small here: a CTA runs one or two long items per launch.
template <int BN, int BM>
struct WgradCfg {
  static constexpr int kABytes = BM * kWgRows * 2;         // BM / 64 boxes of [64 frames][64 channels]
  static constexpr int kBBytes = BN * kWgRows * 2;         // BN / 64 boxes
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BM == 256) ? 3 : ((BN == 256) ? 4 : 8);
  static constexpr int kAccBufs = (BM == 256) ? 1 : 2;
  static constexpr int kAccCols = (BM / 128) * BN;         // TMEM columns of one accumulator set
  static constexpr int kTmemCols = (kAccBufs * kAccCols < 32) ? 32 : kAccBufs * kAccCols;
  static constexpr int kSmemBytes = kStages * kStageBytes + 512 + 1024;
};

// MN-major SWIZZLE_128B operand: 64-element (128-byte) channel groups `lbo` bytes apart, 8-frame groups 1024 B apart.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void decode_tile(int tile, const WgradParams& p, int& tap, int& co0, int& ci0) {
  const int ci_t = tile % p.ci_tiles;
  const int rest = tile / p.ci_tiles;
  const int co_t = rest % p.co_tiles;
  tap = rest / p.co_tiles;
  co0 = co_t;   // tile index; the kernel multiplies by its BM
  ci0 = ci_t;
}

template <int DT, int BN, int BM>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const WgradParams p) {
  using Cfg = WgradCfg<BN, BM>;
  constexpr int kMH = BM / 128;   // 128-row MMAs per tile
  constexpr uint32_t kFormat = (DT == VP3D_BF16) ? 1u : 0u;
  // instruction descriptor: fp32 accumulate, A and B both MN-major (bits 15, 16)
  constexpr uint32_t kIdesc = make_instr_desc(kFormat, 128, BN) | (1u << 15) | (1u << 16);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tmem_full_bar = bars + 2 * Cfg::kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint64_t* clc_full = bars + 2 * Cfg::kStages + 6;          // dynamic schedule (p.dyn_sched): response landed
  uint64_t* clc_empty = clc_full + kWgClcSlots;              // every consumer has read the response
  uint8_t* clc_resp = reinterpret_cast<uint8_t*>(bars) + 256;   // kWgClcSlots x 16 bytes
  static_assert((2 * Cfg::kStages + 6 + 2 * kWgClcSlots) * 8 <= 256, "barrier area layout");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool dyn = p.dyn_sched != 0;
  // Work items of this CTA. Static: blockIdx.x, + gridDim.x, ... Dynamic: the grid has one CTA per item; a running CTA
  // takes its own item, then the items of CTAs that have not been launched yet (cluster launch control) -- CTAs that
  // cannot be placed because something else (NCCL) holds their SM just never get work.
  struct ItemFeed {
    int slot = 0;
    uint32_t phase = 0;
  };
  auto next_item = [&](ItemFeed& f, int item, bool elected, bool whole_warp, int num_items_) -> int {
    if (!dyn) {
      item += gridDim.x;
      return item < num_items_ ? item : -1;
    }
    mbar_wait(&clc_full[f.slot], f.phase);
    int x;
    const bool ok = clc_query(clc_resp + 16 * f.slot, x);
    fence_proxy_async_smem();
    if (whole_warp) __syncwarp();
    if (elected) mbar_arrive(&clc_empty[f.slot]);
    if (++f.slot == kWgClcSlots) {
      f.slot = 0;
      f.phase ^= 1;
    }
    return ok ? x : -1;
  };

  // row blocks kb = seq * kb_per_seq + block-in-sequence; slice s owns kb in [s * kb_all / S, (s + 1) * kb_all / S)
  const long long kb_all = (long long)p.seqs * p.kb_per_seq;
  const int num_items = p.num_tiles * p.num_slices;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 128);
    }
    for (int s = 0; s < kWgClcSlots; ++s) {
      mbar_init(&clc_full[s], 1);
      mbar_init(&clc_empty[s], kWgClcConsumers);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_base_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  // everything above touched only this CTA's shared / tensor memory: it overlaps the tail of the previous kernel (pdl.cuh)
  pdl_enter_long<1>(false);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      ItemFeed feed;
      for (int item = blockIdx.x; item >= 0; item = next_item(feed, item, true, false, num_items)) {
        const int sl = item / p.num_tiles;
        const int tile = item - sl * p.num_tiles;
        const long long kb_lo = kb_all * sl / p.num_slices;
        const long long kb_hi = kb_all * (sl + 1) / p.num_slices;
        int tap, co0, ci_t;
        decode_tile(tile, p, tap, co0, ci_t);
        for (long long kb = kb_lo; kb < kb_hi; ++kb) {
          const int seq = (int)(kb / p.kb_per_seq);
          const int r0 = (int)(kb - (long long)seq * p.kb_per_seq) * kWgRows;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
#pragma unroll
          for (int g = 0; g < BM / 64; ++g)
            tma_load_3d(sa + g * (kWgRows * 128), &tmA, &full_bar[stage], co0 * BM + g * 64, r0, seq);
          uint8_t* sb = sa + Cfg::kABytes;
          const int b_col = ci_t * BN + tap * p.b_tap_col_step;
          const int b_row = r0 + p.b_row_off + tap * p.b_tap_row_step;
#pragma unroll
          for (int g = 0; g < BN / 64; ++g)
            tma_load_3d(sb + g * (kWgRows * 128), &tmB, &full_bar[stage], b_col + g * 64, b_row, seq);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      ItemFeed feed;
      for (int item = blockIdx.x; item >= 0; item = next_item(feed, item, true, false, num_items)) {
        const int sl = item / p.num_tiles;
        const long long kb_lo = kb_all * sl / p.num_slices;
        const long long kb_hi = kb_all * (sl + 1) / p.num_slices;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::kAccCols;
        for (long long kb = kb_lo; kb < kb_hi; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t bdesc = make_mnmajor_sw128_desc(sa + Cfg::kABytes, kWgRows * 128);
#pragma unroll
          for (int h = 0; h < kMH; ++h) {
            // output channels [128 h, 128 h + 128): two 64-channel boxes further into the A part of the stage
            const uint64_t adesc = make_mnmajor_sw128_desc(sa + h * 2 * (kWgRows * 128), kWgRows * 128);
#pragma unroll
            for (int k = 0; k < kWgRows / 16; ++k) {
              // 16 frames of reduction per MMA = two 8-frame groups = 2048 bytes further into every box
              umma_f16_ss(d_tmem + h * BN, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), kIdesc,
                          !(kb == kb_lo && k == 0));
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[acc]);
        if (Cfg::kAccBufs == 2) {
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        } else {
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == kWgSchedWarp) {
    // ------------------------------------------------------------------ tile scheduler (dynamic mode)
    if (dyn && lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      while (true) {
        mbar_wait(&clc_empty[slot], phase ^ 1);
        mbar_expect_tx(&clc_full[slot], 16);
        clc_try_cancel(clc_resp + 16 * slot, &clc_full[slot]);
        mbar_wait(&clc_full[slot], phase);
        int x;
        const bool ok = clc_query(clc_resp + 16 * slot, x);
        fence_proxy_async_smem();
        if (!ok) break;
        if (++slot == kWgClcSlots) {
          slot = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5): TMEM -> red.add
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    ItemFeed feed;
    for (int item = blockIdx.x; item >= 0; item = next_item(feed, item, lane == 0, true, num_items)) {
      const int sl = item / p.num_tiles;
      const int tile = item - sl * p.num_tiles;
      int tap, co0, ci_t;
      decode_tile(tile, p, tap, co0, ci_t);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int h = 0; h < kMH; ++h) {
        float* out_row = p.out + (long long)tap * p.out_tap_stride +
                         (long long)(co0 * BM + h * 128 + row) * p.out_row_stride + (long long)ci_t * BN;
        // rows / 32-column chunks that lie entirely in the padding of a narrow operand (valid_co / valid_ci) hold exact
        // zeros: their reductions are skipped (the Gram GEMM of the expand layer: 192 of 256 in both directions)
        if (co0 * BM + h * 128 + (row & ~31) >= p.valid_co) continue;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          if (ci_t * BN + c * 32 >= p.valid_ci) break;
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + acc * Cfg::kAccCols + h * BN + c * 32 + (static_cast<uint32_t>(quad * 32) << 16),
                             v);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            red_add_v4(out_row + c * 32 + 4 * j, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                       __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      if (Cfg::kAccBufs == 2) {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      } else {
        acc_phase ^= 1;
      }
    }
  }

  pdl_tail_trigger(false);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int DT, int BN, int BM>
static cudaError_t launch_wg(const CUtensorMap& tmA, const CUtensorMap& tmB, const WgradParams& p, int grid,
                             cudaStream_t stream) {
  using Cfg = WgradCfg<BN, BM>;
  static std::atomic<unsigned long long> attr_done{0};   // one bit per device ordinal
  if (cudaError_t e = set_max_smem_once(reinterpret_cast<const void*>(wgrad_gemm_kernel<DT, BN, BM>), Cfg::kSmemBytes,
                                        attr_done))
    return e;
  launch_k(wgrad_gemm_kernel<DT, BN, BM>, dim3(grid), dim3(kWgThreads), Cfg::kSmemBytes, stream, tmA, tmB, p);
  return cudaGetLastError();
}

cudaError_t launch_wgrad(int dtype, int block_n, int block_m, const CUtensorMap& tmA, const CUtensorMap& tmB,
                         const WgradParams& p, int grid, cudaStream_t stream) {
  if (block_n == 256 && block_m == 256) {
    if (dtype == VP3D_F16) return launch_wg<VP3D_F16, 256, 256>(tmA, tmB, p, grid, stream);
    if (dtype == VP3D_BF16) return launch_wg<VP3D_BF16, 256, 256>(tmA, tmB, p, grid, stream);
  } else if (block_n == 256) {
    if (dtype == VP3D_F16) return launch_wg<VP3D_F16, 256, 128>(tmA, tmB, p, grid, stream);
    if (dtype == VP3D_BF16) return launch_wg<VP3D_BF16, 256, 128>(tmA, tmB, p, grid, stream);
  } else if (block_n == 64) {
    if (dtype == VP3D_F16) return launch_wg<VP3D_F16, 64, 128>(tmA, tmB, p, grid, stream);
    if (dtype == VP3D_BF16) return launch_wg<VP3D_BF16, 64, 128>(tmA, tmB, p, grid, stream);
  }
  return cudaErrorInvalidValue;
}

// dw[co][ci][tap] = packed[tap * tap_stride + co * row_stride + ci] * inv_gscale   (nn.Conv1d weight layout).
// Thread = (co, ci); the taps of one (co, ci) are adjacent in dw, so a warp writes one contiguous span.
__global__ void __launch_bounds__(256)
wgrad_finish_kernel(const float* __restrict__ packed, float* __restrict__ dw, int c_out, int c_in, int taps,
                    long long tap_stride, long long row_stride, const float* __restrict__ gscale_buf) {
  pdl_enter_long<3>(false);
  const float inv = gscale_buf != nullptr ? gscale_buf[1] : 1.f;
  const int total = c_out * c_in;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i / c_in;
    const int ci = i - co * c_in;
    const float* src = packed + (long long)co * row_stride + ci;
    float* dst = dw + (long long)i * taps;
    for (int tap = 0; tap < taps; ++tap) dst[tap] = __ldg(src + tap * tap_stride) * inv;
  }
}

cudaError_t launch_wgrad_finish(const float* packed, float* dw, int c_out, int c_in, int taps, long long tap_stride,
                                long long row_stride, const float* gscale_buf, int sm_count, cudaStream_t stream) {
  const long long total = (long long)c_out * c_in;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)sm_count * 8) blocks = (long long)sm_count * 8;
  if (blocks < 1) blocks = 1;
  launch_k(wgrad_finish_kernel, dim3((int)blocks), dim3(256), 0, stream, packed, dw, c_out, c_in, taps, tap_stride, row_stride,
                                                       gscale_buf);
  return cudaGetLastError();
}

}  // namespace vp3d
