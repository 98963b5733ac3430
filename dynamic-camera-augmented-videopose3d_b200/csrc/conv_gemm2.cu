// K1 on CTA pairs (cta_group::2): the temporal-convolution block -- nn.Conv1d + folded BatchNorm + ReLU + residual
// slice-add (common/models/TemporalModel.py:126-138,188-198), the raw train-mode output with BatchNorm statistics, and
// the data gradient of autograd's conv backward -- as ONE tcgen05.mma of M = 256 per K step over two CTAs of a 2-cluster.
//
// Same math, layouts and epilogue as conv_gemm.cu; what changes is who stages what. A pair owns two adjacent 128-row
// tiles of one 256-wide column tile. Each CTA loads its own A tile (16 KB per stage) and HALF of the shared weight tile
// (128 of the 256 output channels, 16 KB), so a stage is 32 KB per SM instead of 48 KB: a third less L2 -> SM traffic
// and shared-memory fill per FLOP, and 5 stages where the single-CTA kernel has 4. Completion of all four loads of a
// stage is collected on the LEADER's full barrier (TMA .cta_group::2 may signal the peer's mbarrier); the leader's MMA
// thread issues tcgen05.mma.cta_group::2 and releases the stage / publishes the accumulators in both CTAs with multicast
// commits; both CTAs run the usual 8-warp epilogue on their own 128 TMEM lanes, and the follower's epilogue warps
// arrive remotely on the leader's accumulator-empty barrier.
//
// Scope: 16-bit operands (K-major weights: forward; MN-major: data gradient), 16-bit output through TMA stores, train-
// mode statistics, at most 1024 output channels when a scale / shift is applied (the whole per-channel table sits in
// shared memory because a pair changes column tile from tile to tile). fp32 output (shrink layer), streaming offsets
// and TF32 stay on conv_gemm_kernel, as do launches of less than two waves.
#include <cstdlib>

#include "ptx.cuh"
#include "kernels.h"
#include "pdl.cuh"
#include "dropout.cuh"
#include "bn_tail.cuh"

namespace vp3d {
namespace {

constexpr int kBM = 128;                  // rows per CTA (256 per pair)
constexpr int kBN = 256;                  // output channels per pair tile
constexpr int kKBytes = 128;              // one swizzle span of K per stage row
constexpr int kClcSlots = 3;              // responses in flight: the producer may be 3 tiles ahead of the slowest epilogue warp
constexpr int kABytes = kBM * kKBytes;            // 16 KB
constexpr int kBHalfBytes = (kBN / 2) * kKBytes;  // 16 KB: this CTA's half of the weight tile
constexpr int kStageBytes = kABytes + kBHalfBytes;
#ifndef VP3D_PAIR_OUTBUFS
#define VP3D_PAIR_OUTBUFS 2
#endif
constexpr int kBarBytes = 512;
constexpr int kAffineCols = 1024;
constexpr int kAffineBytes = 2 * kAffineCols * 4;
// Warp / shared-memory configuration per epilogue variant:
//   EPI 0  generic epilogue (residual through registers, train-mode statistics)   8 epilogue warps, 5 stages
//   EPI 1  + side input by TMA / fused dropout: three staging buffers per epilogue warp -- a side tile is fetched INTO
//          the staging buffer its result is later written over, two chunks ahead of its use
//   EPI 2  lean epilogue (affine, ReLU, dropout only)                               8 epilogue warps, 5 stages
//   EPI 3  lean epilogue for SHORT contractions (the expand layer, K = 192: the launch is bound by the epilogue, which is a
//          chain of dependent steps per 32-column chunk -- TMEM read, math, staging, fence, store -- so what it needs is
//          more chains in flight): 16 epilogue warps, four per TMEM lane quadrant, and 4 stages to make room for their
//          staging buffers
template <int EPI>
struct EpiCfg {
  static constexpr int kEpi = EPI == 3 ? 16 : 8;           // epilogue warps, kEpi / 4 per TMEM lane quadrant
  static constexpr int kParts = kEpi / 4;                  // column parts of a tile, one per warp of a quadrant
  static constexpr int kChunks = kBN / 32 / kParts;        // 32-column chunks per epilogue warp per tile
  static constexpr int kStages = EPI == 3 ? 4 : 5;
  static constexpr int kSchedWarp = 2 + kEpi;              // last warp: tile scheduler (cluster launch control), leader only
  static constexpr int kThreads = 64 + 32 * kEpi + 32;
  static constexpr int kClcConsumers = 2 + 1 + 2 * kEpi;   // both producers, the MMA thread, the epilogue warps of both CTAs
  static constexpr int kOutBufBytes = kEpi * 32 * 64;      // one 32 x 32 staging tile (64 B rows) per epilogue warp
  static constexpr int kOutBufs = EPI == 1 ? 3 : VP3D_PAIR_OUTBUFS;
  static constexpr int kOutStageBytes = kOutBufs * kOutBufBytes;
  static constexpr int kSideBars = EPI == 1 ? kEpi * kOutBufs : 0;
  static constexpr int kSmemBytes = kStages * kStageBytes + kOutStageBytes + kBarBytes + kAffineBytes;
  static_assert(2 * kStages + 4 + 2 + kSideBars <= 40 && (40 + 2 * kClcSlots) * 8 <= 384 && 384 + 16 * kClcSlots <= kBarBytes - 4,
                "barrier area layout");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory");
};
template <int EPI>
constexpr int smem_bytes() { return EpiCfg<EPI>::kSmemBytes; }
constexpr int kTmemCols = 2 * kBN;                // two accumulator buffers

template <int DT>
struct Fmt;
template <>
struct Fmt<VP3D_F16> {
  static constexpr uint32_t kFormat = 0;
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ float2 unpack(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
  // 0xFFFF in every 16-bit lane whose value is > 0
  static __device__ __forceinline__ uint32_t gt0_mask(uint32_t u) {
    return __hgt2_mask(*reinterpret_cast<__half2*>(&u), __half2(__ushort_as_half(0), __ushort_as_half(0)));
  }
};
template <>
struct Fmt<VP3D_BF16> {
  static constexpr uint32_t kFormat = 1;
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ float2 unpack(uint32_t u) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  }
  static __device__ __forceinline__ uint32_t gt0_mask(uint32_t u) {
    return __hgt2_mask(*reinterpret_cast<__nv_bfloat162*>(&u),
                       __nv_bfloat162(__ushort_as_bfloat16(0), __ushort_as_bfloat16(0)));
  }
};

// fp32 pair -> packed 16-bit pair with the ReLU inside the conversion (cvt.rn.relu: negative results become +0)
template <int DT>
__device__ __forceinline__ uint32_t pack_relu(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack_relu<VP3D_F16>(float a, float b) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
template <>
__device__ __forceinline__ uint32_t pack_relu<VP3D_BF16>(float a, float b) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ void st_global_v8(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f,
                                             uint32_t g, uint32_t h) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e),
               "r"(f), "r"(g), "r"(h)
               : "memory");
}
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// Random bytes of one epilogue chunk (this thread's row `grow`, 32 columns from col0 = four 8-channel groups): lo[j] /
// hi[j] hold the bytes of channels 8j .. 8j+3 / 8j+4 .. 8j+7 -- the mask of train.cu's passes, where one Philox block per
// (row pair, 8-channel group) holds the bits of both rows of the pair. Adjacent lanes own the two rows of a pair whenever
// the warp's first row is even: each of them then draws TWO of the chunk's four blocks and hands the partner its half
// (4 shuffles instead of 2 blocks).
__device__ __forceinline__ void drop_chunk_bits(const DropCtx& drop, long long grow, int lane, int col0, uint32_t (&lo)[4],
                                                uint32_t (&hi)[4]) {
  const bool h = (grow & 1) != 0;
  if (((grow - lane) & 1) == 0) {
    const int jg = (lane & 1) * 2;
    const uint4 b0 = drop_bits(drop, grow >> 1, (col0 >> 3) + jg);
    const uint4 b1 = drop_bits(drop, grow >> 1, (col0 >> 3) + jg + 1);
    const uint32_t m0 = h ? b0.z : b0.x, m1 = h ? b0.w : b0.y, m2 = h ? b1.z : b1.x, m3 = h ? b1.w : b1.y;
    const uint32_t r0 = __shfl_xor_sync(0xffffffffu, h ? b0.x : b0.z, 1);
    const uint32_t r1 = __shfl_xor_sync(0xffffffffu, h ? b0.y : b0.w, 1);
    const uint32_t r2 = __shfl_xor_sync(0xffffffffu, h ? b1.x : b1.z, 1);
    const uint32_t r3 = __shfl_xor_sync(0xffffffffu, h ? b1.y : b1.w, 1);
    const bool even = jg == 0;   // even lane drew groups 0, 1 and received 2, 3; odd lane the other way round
    lo[0] = even ? m0 : r0; hi[0] = even ? m1 : r1;
    lo[1] = even ? m2 : r2; hi[1] = even ? m3 : r3;
    lo[2] = even ? r0 : m0; hi[2] = even ? r1 : m1;
    lo[3] = even ? r2 : m2; hi[3] = even ? r3 : m3;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 bits = drop_bits(drop, grow >> 1, (col0 >> 3) + j);
      lo[j] = h ? bits.z : bits.x;
      hi[j] = h ? bits.w : bits.y;
    }
  }
}

// Lean epilogue (EPI = 2), affine + activation + conversion of one 32-column chunk: AFF 0 none, 1 + shift, 2 * scale + shift
// (tables in shared memory, read as broadcast LDS.128); the ReLU costs nothing (it is a modifier of the conversion).
template <int DT, bool RELU, int AFF>
__device__ __forceinline__ void lean_pack(const uint32_t (&x)[32], const float* sc, const float* sh, uint32_t (&pk)[16]) {
  using F = Fmt<DT>;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float a0 = __uint_as_float(x[4 * j + 0]), a1 = __uint_as_float(x[4 * j + 1]);
    float a2 = __uint_as_float(x[4 * j + 2]), a3 = __uint_as_float(x[4 * j + 3]);
    if (AFF == 2) {
      const float4 s4 = reinterpret_cast<const float4*>(sc)[j];
      const float4 h4 = reinterpret_cast<const float4*>(sh)[j];
      a0 = fmaf(a0, s4.x, h4.x); a1 = fmaf(a1, s4.y, h4.y); a2 = fmaf(a2, s4.z, h4.z); a3 = fmaf(a3, s4.w, h4.w);
    } else if (AFF == 1) {
      const float4 h4 = reinterpret_cast<const float4*>(sh)[j];
      a0 += h4.x; a1 += h4.y; a2 += h4.z; a3 += h4.w;
    }
    pk[2 * j + 0] = RELU ? pack_relu<DT>(a0, a1) : F::pack(a0, a1);
    pk[2 * j + 1] = RELU ? pack_relu<DT>(a2, a3) : F::pack(a2, a3);
  }
}

struct PairTile {
  int seq, t0, n0;
};
// pair tile `pt` (column tile fastest, so that the four column tiles of a row pair run at the same time and share A
// through L2) -> this CTA's 128-row tile
__device__ __forceinline__ PairTile decode_pair(int pt, const ConvGemmParams& p, int pairs_per_seq, int rank) {
  PairTile c;
  c.n0 = pt % p.n_tiles;
  const int mp = pt / p.n_tiles;
  c.seq = mp / pairs_per_seq;
  c.t0 = ((mp - c.seq * pairs_per_seq) * 2 + rank) * kBM;
  return c;
}

// MN-major SWIZZLE_128B weight operand (the forward-packed weights read as W^T by the data-gradient GEMM): 64-column
// groups `lbo` bytes apart, 8-row groups 1024 B apart (see conv_gemm.cu)
__device__ __forceinline__ uint64_t mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// EPI = 1 adds to the epilogue (train-mode fusions and the TMA residual path):
//   * dropout after the ReLU (p.drop), same counter-based mask as train.cu's bn_act_fwd pass;
//   * a side input with the geometry of the output (tmS), fetched by TMA into the staging tile two chunks ahead:
//     side_mode 1 adds it (residual rows without a register round trip), side_mode 2 gates the result by side > 0
//     (the ReLU / dropout mask of a layer recovered from its stored activation: dropped and clipped elements are 0).
template <int DT, bool BMN, int EPI>
__global__ void __launch_bounds__(EpiCfg<EPI>::kThreads, 1)
conv_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmS,
                      const ConvGemmParams p) {
  using F = Fmt<DT>;
  using Cfg = EpiCfg<EPI>;
  constexpr int kEpi = Cfg::kEpi, kStages = Cfg::kStages, kThreads = Cfg::kThreads, kSchedWarp = Cfg::kSchedWarp;
  constexpr int kClcConsumers = Cfg::kClcConsumers, kOutBufBytes = Cfg::kOutBufBytes;
  constexpr int kOutBufs = Cfg::kOutBufs;
  constexpr int kOutStageBytes = Cfg::kOutStageBytes;
  constexpr int kElemsPerKBlock = kKBytes / 2;
  constexpr uint32_t kIdesc = make_instr_desc(F::kFormat, 2 * kBM, kBN) | (BMN ? (1u << 16) : 0u);

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("vp3d: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* out_stage = smem + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + kOutStageBytes);
  uint64_t* full_bar = bars;                       // used in the leader only: bytes of BOTH CTAs' loads
  uint64_t* empty_bar = bars + kStages;            // per CTA, released by multicast commits
  uint64_t* tmem_full_bar = bars + 2 * kStages;    // per CTA, multicast commit
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // leader only: epilogue warps of both CTAs arrive
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint64_t* side_bar = tmem_empty_bar + 4;         // EPI: [epilogue warp][staging buffer], this CTA's own TMA loads
  uint64_t* clc_full = bars + 40;                  // dynamic schedule: response landed (per CTA, multicast complete_tx)
  uint64_t* clc_empty = bars + 40 + kClcSlots;     // leader only: every consumer of both CTAs has read the response
  uint8_t* clc_resp = reinterpret_cast<uint8_t*>(bars) + 384;   // kClcSlots x 16 bytes
  float* affine_smem = reinterpret_cast<float*>(out_stage + kOutStageBytes + kBarBytes);   // scale[1024] | shift[1024]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int pairs_per_seq = (p.m_tiles_per_seq + 1) / 2;
  const int total_pairs = p.a_seqs * pairs_per_seq * p.n_tiles;
  const int first_pair = (int)cluster_id_x();
  const int pair_step = (int)cluster_count_x();
  const int num_kb = p.taps * p.kblocks_per_tap;
  const bool dyn = p.dyn_sched != 0;

  // Tile sequence of this cluster. Static: first_pair, first_pair + pair_step, ... Dynamic: the cluster's own index, then
  // whatever the scheduler warp steals (see there); every consumer role walks the response ring on its own.
  struct TileFeed {
    int slot = 0;
    uint32_t phase = 0;
  };
  auto next_tile = [&](TileFeed& tf, int pt, bool elected, bool whole_warp) -> int {
    if (!dyn) {
      pt += pair_step;
      return pt < total_pairs ? pt : -1;
    }
    mbar_wait(&clc_full[tf.slot], tf.phase);
    int x;
    const bool ok = clc_query(clc_resp + 16 * tf.slot, x);
    fence_proxy_async_smem();     // this generic read is ordered before the next (async-proxy) response into the slot
    if (whole_warp) __syncwarp();   // (every lane of an epilogue warp has read the slot before lane 0 releases it)
    if (elected) mbar_arrive_cluster(map_to_cta(smem_u32(&clc_empty[tf.slot]), 0));
    if (++tf.slot == kClcSlots) {
      tf.slot = 0;
      tf.phase ^= 1;
    }
    return ok ? (x >> 1) : -1;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 2 * kEpi);
    }
    if (EPI == 1) {
      tma_prefetch_desc(&tmS);
      for (int s = 0; s < Cfg::kSideBars; ++s) mbar_init(&side_bar[s], 1);
    }
    for (int s = 0; s < kClcSlots; ++s) {
      mbar_init(&clc_full[s], 1);
      mbar_init(&clc_empty[s], kClcConsumers);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_base_slot, kTmemCols);
    tmem_relinquish_2cta();
  }
  // everything above touched only this CTA's shared / tensor memory: it overlaps the tail of the previous kernel (pdl.cuh)
  pdl_enter_long<2>(total_pairs <= pair_step);
  if (EPI >= 2) {
    // lean epilogue: the keep scale of the dropout is folded into the tables (relu(x) * k == relu(x * k) for k > 0)
    const float ks = p.drop.p > 0.f ? make_drop(p.drop).keep_scale : 1.f;
    if (p.shift != nullptr || p.scale != nullptr || p.drop.p > 0.f) {
      const int n_cols = p.n_tiles * kBN;
      for (int j = threadIdx.x; j < n_cols; j += kThreads) {
        affine_smem[j] = (p.scale != nullptr ? __ldg(p.scale + j) : 1.f) * ks;
        affine_smem[kAffineCols + j] = (p.shift != nullptr ? __ldg(p.shift + j) : 0.f) * ks;
      }
    }
  } else if (p.shift != nullptr) {
    const int n_cols = p.n_tiles * kBN;
    for (int j = threadIdx.x; j < n_cols; j += kThreads) {
      affine_smem[j] = p.scale != nullptr ? __ldg(p.scale + j) : 1.f;
      affine_smem[kAffineCols + j] = __ldg(p.shift + j);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before anything here can signal them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      TileFeed tf;
      for (int pt = first_pair; pt >= 0; pt = next_tile(tf, pt, true, false)) {
        const PairTile tc = decode_pair(pt, p, pairs_per_seq, rank);
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / p.kblocks_per_tap;
          const int kc = kb - tap * p.kblocks_per_tap;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * kStageBytes);
          const uint32_t full_leader = map_to_cta(smem_u32(&full_bar[stage]), 0);
          uint8_t* sa = smem + stage * kStageBytes;
          tma_load_3d_2cta(sa, &tmA, full_leader, kc * kElemsPerKBlock, tc.t0 + p.a_row_off + tap * p.tap_row_step,
                           tc.seq);
          if (BMN) {
            // k-rows kc*64.. of the [c_out][taps * c_in] weights, this CTA's 128 columns of the tap's column tile
#pragma unroll
            for (int g = 0; g < 2; ++g)
              tma_load_2d_2cta(sa + kABytes + g * 8192, &tmB, full_leader,
                               tap * p.b_tap_col_step + tc.n0 * kBN + rank * (kBN / 2) + g * 64, kc * kElemsPerKBlock);
          } else {
            tma_load_2d_2cta(sa + kABytes, &tmB, full_leader, kb * kElemsPerKBlock, tc.n0 * kBN + rank * (kBN / 2));
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      TileFeed tf;
      for (int pt = first_pair; pt >= 0; pt = next_tile(tf, pt, true, false)) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t adesc = make_kmajor_sw128_desc(sa);
          const uint64_t bdesc = BMN ? mnmajor_sw128_desc(sa + kABytes, 8192) : make_kmajor_sw128_desc(sa + kABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)   // 32 bytes of K: +2 in the address field (K-major), 16 k-rows = 2048 B (MN-major)
            umma_f16_ss_2cta(d_tmem, adesc + 2 * k, bdesc + (BMN ? (uint64_t)(128 * k) : (uint64_t)(2 * k)), kIdesc,
                             (kb | k) != 0);
          umma_commit_2cta(&empty_bar[stage], 3);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2cta(&tmem_full_bar[acc], 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp == kSchedWarp) {
    // ------------------------------------------------------------------ tile scheduler (leader CTA, dynamic mode)
    // One outstanding steal per free response slot: arm the slot's barrier in both CTAs, ask the hardware to cancel the
    // next cluster of this grid that has not started, and look at the answer only to learn when to stop (the first
    // refusal ends the sequence for every consumer as well).
    if (dyn && rank == 0 && lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      while (true) {
        mbar_wait(&clc_empty[slot], phase ^ 1);
        mbar_expect_tx(&clc_full[slot], 16);
        mbar_expect_tx_cluster(map_to_cta(smem_u32(&clc_full[slot]), 1), 16);
        clc_try_cancel_multicast(clc_resp + 16 * slot, &clc_full[slot]);
        mbar_wait(&clc_full[slot], phase);
        int x;
        const bool ok = clc_query(clc_resp + 16 * slot, x);
        fence_proxy_async_smem();
        if (!ok) break;
        if (++slot == kClcSlots) {
          slot = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (EPI >= 2) {
    // ------------------------------------------------------------------ lean epilogue (warps 2.., both CTAs)
    // Launches without residual, statistics or side input -- the expand layer (K = 192: the whole launch is epilogue
    // bound), the 3-tap inference layers, data gradients without fan-in, the lifter's Linear layers: TMEM -> (affine) ->
    // conversion with the ReLU inside -> dropout as an AND on the packed pairs -> staged tile -> TMA store. With 8 warps
    // the next chunk's tcgen05.ld is in flight during the math of the current one; the accumulator is handed back to the
    // MMA thread as soon as the tile's last chunk has left tensor memory (before its math).
    constexpr int kChunks = Cfg::kChunks;
    constexpr bool kPipeLd = EPI == 2;      // 16 warps: one chunk's registers per thread (102-register budget)
    const int quad = warp & 3;
    const int epi = warp - 2;
    const int half = epi >> 2;              // column part of the tile this warp owns (0 .. kParts - 1)
    const int row = quad * 32 + lane;
    unsigned out_buf = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    DropCtx drop;
    drop.on = false;
    if (p.drop.p > 0.f) drop = make_drop(p.drop);
    // keep iff byte >= thresh  <=>  byte + (256 - thresh) carries out of its byte: bit 7 of maj(x, y, (x & 7f..) + (y & 7f..))
    const uint32_t kadd = ((256u - drop.thresh) & 0xFFu) * 0x01010101u;
    const uint32_t kadd7 = kadd & 0x7F7F7F7Fu;
    const int aff = (p.scale != nullptr || drop.on) ? 2 : (p.shift != nullptr ? 1 : 0);
    const uint32_t empty_leader0 = map_to_cta(smem_u32(&tmem_empty_bar[0]), 0);
    const uint32_t empty_leader1 = map_to_cta(smem_u32(&tmem_empty_bar[1]), 0);
    TileFeed tf;
    for (int pt = first_pair; pt >= 0; pt = next_tile(tf, pt, lane == 0, true)) {
      const PairTile tc = decode_pair(pt, p, pairs_per_seq, rank);
      const long long grow = (long long)tc.seq * p.rows_out + tc.t0 + row;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + acc * kBN + half * kChunks * 32 + (static_cast<uint32_t>(quad * 32) << 16);
      // (8 warps, 168 registers: two chunks' registers per thread, the next load in flight across the math; 16 warps, 96
      // registers: one chunk's registers, load and wait adjacent -- four warps per scheduler cover the TMEM latency)
      uint32_t v[kPipeLd ? 2 : 1][32];
      if (kPipeLd) tmem_ld_32x32b_x32(taddr, v[0]);
#pragma unroll
      for (int cq = 0; cq < kChunks; ++cq) {
        uint32_t (&x)[32] = v[kPipeLd ? (cq & 1) : 0];
        if (!kPipeLd) tmem_ld_32x32b_x32(taddr + cq * 32, v[0]);
        tmem_wait_ld();
        if (kPipeLd && cq + 1 < kChunks) tmem_ld_32x32b_x32(taddr + (cq + 1) * 32, v[kPipeLd ? ((cq + 1) & 1) : 0]);
        if (cq + 1 == kChunks) {
          // this warp has read its share of the accumulator buffer: one arrival per warp on the LEADER's barrier
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(&tmem_empty_bar[acc]);
            else mbar_arrive_cluster(acc ? empty_leader1 : empty_leader0);
          }
        }
        const int c = half * kChunks + cq;
        const int col0 = tc.n0 * kBN + c * 32;
        const float* sc = affine_smem + col0;
        const float* sh = affine_smem + kAffineCols + col0;
        uint32_t pk[16];
        if (p.relu) {
          if (aff == 2) lean_pack<DT, true, 2>(x, sc, sh, pk);
          else if (aff == 1) lean_pack<DT, true, 1>(x, sc, sh, pk);
          else lean_pack<DT, true, 0>(x, sc, sh, pk);
        } else {
          if (aff == 2) lean_pack<DT, false, 2>(x, sc, sh, pk);
          else if (aff == 1) lean_pack<DT, false, 1>(x, sc, sh, pk);
          else lean_pack<DT, false, 0>(x, sc, sh, pk);
        }
        if (drop.on) {
          uint32_t lo[4], hi[4];
          drop_chunk_bits(drop, grow, lane, col0, lo, hi);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t ml = maj3(lo[j], kadd, (lo[j] & 0x7F7F7F7Fu) + kadd7);   // bit 7 of byte i: keep channel 8j + i
            const uint32_t mh = maj3(hi[j], kadd, (hi[j] & 0x7F7F7F7Fu) + kadd7);
            pk[4 * j + 0] &= prmt(ml, 0u, 0x9988u);   // sign of byte 0 over the low half, of byte 1 over the high half
            pk[4 * j + 1] &= prmt(ml, 0u, 0xBBAAu);
            pk[4 * j + 2] &= prmt(mh, 0u, 0x9988u);
            pk[4 * j + 3] &= prmt(mh, 0u, 0xBBAAu);
          }
        }
        if (p.direct_out) {
          // Straight from registers: a thread holds 64 contiguous bytes of its output row = two 32-byte (one sector each)
          // stores. No staging tile, proxy fence, warp barrier or TMA store -- and no epilogue traffic through the shared
          // memory the tensor pipe reads its operands from.
          if (tc.t0 + row < p.rows_out) {
            uint16_t* o = static_cast<uint16_t*>(p.out) + (long long)tc.seq * p.out_seq_stride +
                          (long long)(tc.t0 + row) * p.out_row_stride + col0;
            st_global_v8(o, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
            st_global_v8(o + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
          }
          continue;
        }
        const unsigned b = out_buf;
        out_buf = out_buf + 1 == kOutBufs ? 0 : out_buf + 1;
        uint8_t* my_stage = out_stage + b * kOutBufBytes + epi * (32 * 64);
        if (lane == 0) tma_store_wait_read<kOutBufs - 1>();   // the store that last used this buffer has drained it
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_shared_v4(my_stage + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), pk[4 * j + 0], pk[4 * j + 1], pk[4 * j + 2],
                       pk[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmC, my_stage, col0, tc.t0 + quad * 32, tc.seq);
          tma_store_commit();
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait_all<0>();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
    constexpr int kChunks = kBN / 64;   // 32-column chunks per epilogue warp
    const int quad = warp & 3;
    const int epi = warp - 2;
    const int half = epi >> 2;
    const int row = quad * 32 + lane;
    unsigned out_buf = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    // EPI: this warp's chunks form one stream gc = tiles done * kChunks + chunk; chunk gc is staged in buffer
    // gc % kOutBufs. A side tile is requested two chunks ahead of its use within a tile, the first two of a tile at the
    // top of the tile (before the wait for its accumulator, which hides their latency).
    const bool side_on = EPI && p.side_mode != 0;
    int gc = 0;
    auto side_issue = [&](const PairTile& tcs, int ci, int g) {   // lane 0: side tile of chunk ci of tile tcs, stream index g
      const int bb = g % kOutBufs;
      uint64_t* bar = &side_bar[epi * kOutBufs + bb];
      mbar_expect_tx(bar, 32 * 64);
      tma_load_3d(out_stage + bb * kOutBufBytes + epi * (32 * 64), &tmS, bar, tcs.n0 * kBN + (half * kChunks + ci) * 32,
                  tcs.t0 + quad * 32 + p.side_row_off, tcs.seq);
    };
    DropCtx drop;
    drop.on = false;
    if (EPI && p.drop.p > 0.f) drop = make_drop(p.drop);
    // train-mode BatchNorm statistics of the stored values, as in conv_gemm.cu: lane l owns column chunk * 32 + l of
    // this warp's column half, accumulated in registers and flushed when the CTA changes column tile
    float st_s[4] = {0.f, 0.f, 0.f, 0.f}, st_q[4] = {0.f, 0.f, 0.f, 0.f};
    int stat_n0 = -1;
    auto stat_flush = [&]() {
      if (stat_n0 >= 0) {
#pragma unroll
        for (int ci = 0; ci < kChunks; ++ci) {
          const int col = stat_n0 * kBN + (half * kChunks + ci) * 32 + lane;
          atomicAdd(p.stat_sum + col, (double)st_s[ci]);
          atomicAdd(p.stat_sqsum + col, (double)st_q[ci]);
          st_s[ci] = st_q[ci] = 0.f;
        }
      }
    };
    const uint32_t empty_leader0 = map_to_cta(smem_u32(&tmem_empty_bar[0]), 0);
    const uint32_t empty_leader1 = map_to_cta(smem_u32(&tmem_empty_bar[1]), 0);
    TileFeed tf;
    for (int pt = first_pair; pt >= 0; pt = next_tile(tf, pt, lane == 0, true)) {
      const PairTile tc = decode_pair(pt, p, pairs_per_seq, rank);
      if (side_on && lane == 0) {
        tma_store_wait_read<0>();     // the staging buffers of the previous tile's last chunks have been drained
        side_issue(tc, 0, gc);
        side_issue(tc, 1, gc + 1);
      }
      if (p.stat_sum != nullptr && tc.n0 != stat_n0) {
        stat_flush();
        stat_n0 = tc.n0;
      }
      const int t = tc.t0 + row;
      const bool row_ok = t < p.rows_out;
      const long long res_row = (long long)t * p.res_row_mul + p.res_row_off;
      const bool res_row_ok = p.res_rows <= 0 || (res_row >= 0 && res_row < p.res_rows);
      const long long res_off = (long long)tc.seq * p.res_seq_stride + res_row * p.res_row_stride - p.res_col_off;
      const bool res_any = p.res != nullptr && row_ok && res_row_ok;
      auto res_in_window = [&](int c) {
        const int col0 = tc.n0 * kBN + c * 32;
        return p.res_cols <= 0 || (col0 >= p.res_col_off && col0 < p.res_col_off + p.res_cols);
      };
      auto res_load = [&](int c, uint4 (&r)[4]) {
        if (res_any && res_in_window(c)) {
          const uint4* r4 = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(p.res) + res_off +
                                                           tc.n0 * kBN + c * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) r[j] = __ldg(r4 + j);
        }
      };
      // residual rows travel through registers two 32-column chunks ahead of their use (ncu: with one chunk of lead
      // the first use of a residual register was the top stall of the kernel, 10 % of all warp samples)
      uint4 rcur[4], rnext[4], rnext2[4];
      if (res_any) {
        const char* rp = reinterpret_cast<const char*>(static_cast<const uint16_t*>(p.res) + res_off + tc.n0 * kBN +
                                                       half * kChunks * 32);
#pragma unroll
        for (int k = 0; k < kChunks * 64; k += 128)
          if (res_in_window(half * kChunks + k / 64)) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + k));
      }
      res_load(half * kChunks, rcur);
      res_load(half * kChunks + 1, rnext);

      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();

#pragma unroll
      for (int cq = 0; cq < kChunks; ++cq) {
        const int c = half * kChunks + cq;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + acc * kBN + c * 32 + (static_cast<uint32_t>(quad * 32) << 16), v);
        if (c + 2 < (half + 1) * kChunks) res_load(c + 2, rnext2);
        tmem_wait_ld();
        const int col0 = tc.n0 * kBN + c * 32;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.scale != nullptr) {
          const float4* sc4 = reinterpret_cast<const float4*>(affine_smem + col0);
          const float4* sh4 = reinterpret_cast<const float4*>(affine_smem + kAffineCols + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 sc = sc4[j];
            const float4 sh = sh4[j];
            f[4 * j + 0] = fmaf(f[4 * j + 0], sc.x, sh.x);
            f[4 * j + 1] = fmaf(f[4 * j + 1], sc.y, sh.y);
            f[4 * j + 2] = fmaf(f[4 * j + 2], sc.z, sh.z);
            f[4 * j + 3] = fmaf(f[4 * j + 3], sc.w, sh.w);
          }
        } else if (p.shift != nullptr) {
          const float4* sh4 = reinterpret_cast<const float4*>(affine_smem + kAffineCols + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 sh = sh4[j];
            f[4 * j + 0] += sh.x;
            f[4 * j + 1] += sh.y;
            f[4 * j + 2] += sh.z;
            f[4 * j + 3] += sh.w;
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (EPI && drop.on) {
          const long long grow = (long long)tc.seq * p.rows_out + t;
          uint32_t lo[4], hi[4];
          drop_chunk_bits(drop, grow, lane, col0, lo, hi);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float m[8];
            drop_mult8(drop, lo[j], hi[j], m);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[8 * j + k] *= m[k];
          }
        }
        if (res_any && res_in_window(c)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 r = rcur[j];
            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 x = F::unpack(rr[e]);
              f[8 * j + 2 * e + 0] += x.x;
              f[8 * j + 2 * e + 1] += x.y;
            }
          }
        }
        if (EPI == 0 && p.direct_out) {
          // (launcher: no statistics) straight from registers, as in the lean epilogue
          if (row_ok) {
            uint16_t* o = static_cast<uint16_t*>(p.out) + (long long)tc.seq * p.out_seq_stride +
                          (long long)t * p.out_row_stride + col0;
            st_global_v8(o, F::pack(f[0], f[1]), F::pack(f[2], f[3]), F::pack(f[4], f[5]), F::pack(f[6], f[7]),
                         F::pack(f[8], f[9]), F::pack(f[10], f[11]), F::pack(f[12], f[13]), F::pack(f[14], f[15]));
            st_global_v8(o + 16, F::pack(f[16], f[17]), F::pack(f[18], f[19]), F::pack(f[20], f[21]), F::pack(f[22], f[23]),
                         F::pack(f[24], f[25]), F::pack(f[26], f[27]), F::pack(f[28], f[29]), F::pack(f[30], f[31]));
          }
          ++gc;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            rcur[j] = rnext[j];
            rnext[j] = rnext2[j];
          }
          continue;
        }
        // staged SWIZZLE_64B tile -> one TMA store per warp per 32 columns (rows past the sequence end are clipped)
        const unsigned b = out_buf;
        out_buf = out_buf + 1 == kOutBufs ? 0 : out_buf + 1;
        uint8_t* my_stage = out_stage + b * kOutBufBytes + epi * (32 * 64);
        if (side_on) {
          // the side tile of this chunk sits in the staging buffer (same SWIZZLE_64B box as the store): thread = row
          mbar_wait(&side_bar[epi * kOutBufs + b], (uint32_t)(gc / kOutBufs) & 1u);
          if (p.side_mode == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 r = *reinterpret_cast<const uint4*>(my_stage + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
              const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 x = F::unpack(rr[e]);
                f[8 * j + 2 * e + 0] += x.x;
                f[8 * j + 2 * e + 1] += x.y;
              }
            }
          } else {
            // gate: the comparison side > 0 happens on the packed pairs when the result is staged (below)
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= p.side_scale;
          }
        } else {
          if (lane == 0) tma_store_wait_read<kOutBufs - 1>();   // the store that last used this buffer has drained it
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint8_t* dst = my_stage + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
          uint32_t w0 = F::pack(f[8 * j + 0], f[8 * j + 1]), w1 = F::pack(f[8 * j + 2], f[8 * j + 3]);
          uint32_t w2 = F::pack(f[8 * j + 4], f[8 * j + 5]), w3 = F::pack(f[8 * j + 6], f[8 * j + 7]);
          if (side_on && p.side_mode == 2) {
            // out = side > 0 ? out : 0 as one compare-to-mask + AND per pair; the side values sit where the result goes
            const uint4 r = *reinterpret_cast<const uint4*>(dst);
            w0 &= F::gt0_mask(r.x);
            w1 &= F::gt0_mask(r.y);
            w2 &= F::gt0_mask(r.z);
            w3 &= F::gt0_mask(r.w);
          }
          st_shared_v4(dst, w0, w1, w2, w3);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (p.stat_sum != nullptr) {
          int valid = p.rows_out - (tc.t0 + quad * 32);
          valid = valid < 0 ? 0 : (valid > 32 ? 32 : valid);
          float cs = 0.f, cq = 0.f;
          const uint32_t col_off = (lane & 7) * 2, col_chunk = lane >> 3;
#pragma unroll 8
          for (int r = 0; r < valid; ++r) {
            const uint16_t h = *reinterpret_cast<const uint16_t*>(my_stage + r * 64 + ((col_chunk ^ ((r >> 1) & 3)) << 4) +
                                                                  col_off);
            const float x = (DT == VP3D_F16) ? __half2float(__ushort_as_half(h)) : __uint_as_float((uint32_t)h << 16);
            cs += x;
            cq = fmaf(x, x, cq);
          }
          const int ci = c - half * kChunks;
#pragma unroll
          for (int k = 0; k < kChunks; ++k)
            if (k == ci) {
              st_s[k] += cs;
              st_q[k] += cq;
            }
        }
        if (lane == 0) {
          tma_store_3d(&tmC, my_stage, col0, tc.t0 + quad * 32, tc.seq);
          tma_store_commit();
          if (side_on && c + 2 < (half + 1) * kChunks) {
            // chunk gc + 2 reuses the buffer of chunk gc - 1: its store (all but the newest one) has drained it
            tma_store_wait_read<1>();
            side_issue(tc, c + 2 - half * kChunks, gc + 2);
          }
        }
        ++gc;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          rcur[j] = rnext[j];
          rnext[j] = rnext2[j];
        }
      }
      // this warp has read its share of the accumulator buffer: one arrival per warp on the LEADER's barrier
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&tmem_empty_bar[acc]);
        else mbar_arrive_cluster(acc ? empty_leader1 : empty_leader0);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.stat_sum != nullptr) stat_flush();
    if (lane == 0) tma_store_wait_all<0>();
  }

  pdl_tail_trigger(total_pairs <= pair_step);
  // nobody leaves (or frees tensor memory) while the peer may still read this CTA's shared memory or signal its barriers
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
  if (p.fin.sum != nullptr)
    bn_finalize_tail(p.fin, p.fin_counter, p.n_tiles * kBN,
                     reinterpret_cast<volatile int*>(reinterpret_cast<uint8_t*>(bars) + kBarBytes - 4));
}

template <int DT, bool BMN, int EPI>
cudaError_t launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmS,
                        const ConvGemmParams& p, int clusters, cudaStream_t stream) {
  static std::atomic<unsigned long long> attr_done{0};   // one bit per device ordinal
  if (cudaError_t e = set_max_smem_once(reinterpret_cast<const void*>(conv_gemm_pair_kernel<DT, BMN, EPI>),
                                        smem_bytes<EPI>(), attr_done))
    return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(EpiCfg<EPI>::kThreads);
  cfg.dynamicSmemBytes = smem_bytes<EPI>();
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // pdl.cuh
  attr[1].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, conv_gemm_pair_kernel<DT, BMN, EPI>, tmA, tmB, tmC, tmS, p);
}

template <int DT, bool BMN>
cudaError_t launch_lean(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmS,
                        const ConvGemmParams& p, int clusters, cudaStream_t stream) {
  static const bool lean16_on = [] {
    const char* e = std::getenv("VP3D_LEAN16");
    return e == nullptr || e[0] != '0';
  }();
  if (lean16_on && p.taps * p.kblocks_per_tap <= 6)   // short contraction: epilogue bound, 16 epilogue warps
    return launch_pair<DT, BMN, 3>(tmA, tmB, tmC, tmS, p, clusters, stream);
  return launch_pair<DT, BMN, 2>(tmA, tmB, tmC, tmS, p, clusters, stream);
}

template <int DT, bool BMN>
cudaError_t launch_pair_epi(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmS,
                            const ConvGemmParams& p, int clusters, cudaStream_t stream) {
  // lean epilogue whenever nothing but affine / ReLU / dropout happens to the accumulator (VP3D_LEAN_EPI=0: A/B switch)
  static const bool lean_on = [] {
    const char* e = std::getenv("VP3D_LEAN_EPI");
    return e == nullptr || e[0] != '0';
  }();
  if (lean_on && p.side_mode == 0 && p.res == nullptr && p.stat_sum == nullptr) {
    static const bool direct_on = [] {
      const char* e = std::getenv("VP3D_DIRECT_OUT");
      return e != nullptr && (e[0] == '1' || e[0] == '2');     // off by default, see below
    }();
    ConvGemmParams q = p;
    // Measured per launch (ncu, 64 x 4338-frame inference): 3-tap layers 1159 / 1193 / 1207 / 1153 us with staged TMA
    // stores, 1136 / 1167 / 1178 / 1130 us with direct stores (the tensor pipe no longer shares the shared-memory
    // bandwidth with the staging traffic); the write-bound expand layer 128 -> 184 us (32-byte sector stores scattered
    // over 32 rows per instruction lose to TMA's 64-byte row segments when the launch is nothing but output). But the
    // whole inference step, where the chip sits at its power cap, does not follow the isolated launches: 7.80 ms with
    // direct stores on the MMA-bound launches against 7.75 ms with TMA stores (4 alternating runs each, same box). The
    // path stays as an experiment switch (VP3D_DIRECT_OUT=1; =2 extends it to the residual epilogue).
    q.direct_out = (direct_on && p.taps * p.kblocks_per_tap > 6 && (reinterpret_cast<uintptr_t>(p.out) & 31) == 0 &&
                    p.out_row_stride % 16 == 0 && p.out_seq_stride % 16 == 0) ? 1 : 0;
    return launch_lean<DT, BMN>(tmA, tmB, tmC, tmS, q, clusters, stream);
  }
  if (p.side_mode != 0 || p.drop.p > 0.f) return launch_pair<DT, BMN, 1>(tmA, tmB, tmC, tmS, p, clusters, stream);
  static const bool direct_generic = [] {
    const char* e = std::getenv("VP3D_DIRECT_OUT");
    return e != nullptr && e[0] == '2';      // the residual epilogue with direct stores: off by default (see the lean path)
  }();
  ConvGemmParams q = p;
  q.direct_out = (direct_generic && p.stat_sum == nullptr && (reinterpret_cast<uintptr_t>(p.out) & 31) == 0 &&
                  p.out_row_stride % 16 == 0 && p.out_seq_stride % 16 == 0) ? 1 : 0;
  return launch_pair<DT, BMN, 0>(tmA, tmB, tmC, tmS, q, clusters, stream);
}

}  // namespace

bool conv_gemm_pair_supported(int dtype, int block_n, int /*w_mn_major: both weight layouts are covered*/,
                              const ConvGemmParams& p) {
  if (dtype != VP3D_F16 && dtype != VP3D_BF16) return false;
  if (block_n != kBN || p.out_f32 || p.dyn != nullptr) return false;
  if ((p.shift != nullptr || p.drop.p > 0.f) && p.n_tiles * kBN > kAffineCols) return false;
  return true;
}

// K-major weights: tmB must be encoded with a box of 128 output channels (half a column tile) x 64 elements of K;
// MN-major weights: the [64 k-rows][64 columns] map of the single-CTA kernel
cudaError_t launch_conv_gemm_pair(int dtype, int w_mn_major, const CUtensorMap& tmA, const CUtensorMap& tmB,
                                  const CUtensorMap& tmC, const CUtensorMap& tmS, const ConvGemmParams& p, int sm_count,
                                  cudaStream_t stream) {
  const int pairs_per_seq = (p.m_tiles_per_seq + 1) / 2;
  const long long total_pairs = (long long)p.a_seqs * pairs_per_seq * p.n_tiles;
  int clusters = sm_count / 2;
  if (total_pairs < clusters || p.dyn_sched) clusters = (int)total_pairs;   // dynamic: one cluster per tile, most get stolen
  if (clusters < 1) return cudaSuccess;
  // statistics: a cluster count that is a multiple of the column-tile count keeps every CTA on one column tile, so its
  // per-channel sums stay in registers for the whole launch and are flushed once
  if (p.stat_sum != nullptr && clusters > p.n_tiles) clusters -= clusters % p.n_tiles;
  if (w_mn_major) {
    if (dtype == VP3D_BF16) return launch_pair_epi<VP3D_BF16, true>(tmA, tmB, tmC, tmS, p, clusters, stream);
    return launch_pair_epi<VP3D_F16, true>(tmA, tmB, tmC, tmS, p, clusters, stream);
  }
  if (dtype == VP3D_BF16) return launch_pair_epi<VP3D_BF16, false>(tmA, tmB, tmC, tmS, p, clusters, stream);
  return launch_pair_epi<VP3D_F16, false>(tmA, tmB, tmC, tmS, p, clusters, stream);
}

}  // namespace vp3d
