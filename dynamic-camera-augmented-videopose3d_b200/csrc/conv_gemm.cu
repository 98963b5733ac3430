// K1: temporal-convolution block as a tcgen05 / TMEM implicit GEMM (sm_100a).
//
// What it replaces in the reference: every nn.Conv1d of the TemporalModel stack together with the BatchNorm
// (eval: folded scale/shift), ReLU and residual slice-add that follow it
// (common/models/TemporalModel.py:126-138 for the dilated model, :188-198 for the strided 1f model).
//
// Data layout: activations are channels-last [seq][frame][channel] in HBM, so an output tile of 128 frames
// times one tap of a dilated (or strided) convolution is a plain 2-D box of the input: TMA fetches
// [128 frames, 128 bytes of channels] at frame offset  t0 + tap * dilation  straight into a SWIZZLE_128B
// shared-memory tile that tcgen05.mma consumes as a K-major operand. Frames past the end of a sequence are
// zero-filled by TMA, so ragged tails need no masking on the load side.
//
// Pipeline (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer   : ring of STAGES {A 16 KB, B BN*128 B} stages, full/empty mbarriers
//   warp 1      MMA issuer     : one thread issues 4 x tcgen05.mma (128 x BN x 32 bytes-of-K) per stage into one
//                                of two TMEM accumulator buffers, tcgen05.commit frees the stage / publishes the tile
//   warps 2..5  epilogue       : tcgen05.ld 32 lanes x 32 columns, y = acc*scale[c] + shift[c], ReLU, + residual,
//                                convert, 16-byte stores; overlaps the next tile's MMAs through the 2nd TMEM buffer
#include "ptx.cuh"
#include "kernels.h"
#include "pdl.cuh"
#include "bn_tail.cuh"

namespace vp3d {

constexpr int kBlockM = 128;
constexpr int kTileKBytes = 128;  // one swizzle span of K per stage row
#ifndef VP3D_K1_STAGES
#define VP3D_K1_STAGES 4
#endif
constexpr int kOutBufs = 2;                         // TMA-store staging buffers per epilogue warp (ring)
constexpr int kEpiWarps = 8;                        // two per TMEM lane quadrant, each takes half of the columns
constexpr int kNumThreads = 64 + 32 * kEpiWarps;    // + TMA producer warp + MMA issuer warp

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kTileKBytes;
  static constexpr int kBBytes = BN * kTileKBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? VP3D_K1_STAGES : 2 * VP3D_K1_STAGES;
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
  // TMA-store staging: per epilogue warp a ring of kOutBufs buffers of 32 rows x 64 B
  static constexpr int kOutBufBytes = kEpiWarps * 32 * 64;
  static constexpr int kOutStageBytes = kOutBufs * kOutBufBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kAffineBytes = 2 * BN * 4;    // scale / shift of the CTA's current column tile
  static constexpr int kSmemBytes = kStages * kStageBytes + kOutStageBytes + kBarBytes + kAffineBytes;
};

template <int DT>
struct ElemTraits;
template <>
struct ElemTraits<VP3D_F16> {
  static constexpr int kBytes = 2;
  static constexpr uint32_t kFormat = 0;
};
template <>
struct ElemTraits<VP3D_BF16> {
  static constexpr int kBytes = 2;
  static constexpr uint32_t kFormat = 1;
};
template <>
struct ElemTraits<VP3D_TF32> {
  static constexpr int kBytes = 4;
  static constexpr uint32_t kFormat = 2;
};

__device__ __forceinline__ uint32_t pack2(float a, float b, std::integral_constant<int, VP3D_F16>) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack2(float a, float b, std::integral_constant<int, VP3D_BF16>) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2(uint32_t u, std::integral_constant<int, VP3D_F16>) {
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}
__device__ __forceinline__ float2 unpack2(uint32_t u, std::integral_constant<int, VP3D_BF16>) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}

// round-to-nearest onto the TF32 grid, so that a following kind::tf32 MMA (which truncates) consumes it exactly
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

struct TileCoord {
  int seq, t0, n0;
};
__device__ __forceinline__ TileCoord decode_tile(int tile, const ConvGemmParams& p) {
  TileCoord c;
  const int n_tile = tile % p.n_tiles;
  const int m_tile = tile / p.n_tiles;
  c.seq = m_tile / p.m_tiles_per_seq;
  c.t0 = (m_tile - c.seq * p.m_tiles_per_seq) * kBlockM;
  c.n0 = n_tile;
  return c;
}

// MN-major SWIZZLE_128B operand (weights read as [k rows][n columns], n contiguous): 64-column groups `lbo` bytes
// apart, 8-row groups 1024 B apart -- what BN / 64 TMA boxes of [64 k-rows][64 columns] produce.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc_b(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// BMN = true: the weight operand is MN-major, i.e. the *forward-packed* weights [c_out][taps * c_in] are consumed as
// W^T by the data-gradient GEMM without a transposed copy (16-bit operand types only).
template <int DT, int BN, bool BMN>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const ConvGemmParams p) {
  using Cfg = GemmCfg<BN>;
  using ET = ElemTraits<DT>;
  static_assert(!BMN || ET::kBytes == 2, "MN-major weights: fp16 / bf16 only");
  constexpr int kElemsPerKBlock = kTileKBytes / ET::kBytes;  // 64 (16-bit) or 32 (tf32)
  constexpr int kMmasPerStage = 4;                           // 32 bytes of K per tcgen05.mma
  constexpr uint32_t kIdesc = make_instr_desc(ET::kFormat, kBlockM, BN) | (BMN ? (1u << 16) : 0u);

  // SWIZZLE_128B tiles need 1024-byte alignment; the kernel has no static shared memory, so the dynamic window starts at
  // the CTA's (1024-aligned) shared base -- checked below instead of paying 1 KB of slack
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("vp3d: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* out_stage = smem + Cfg::kStages * Cfg::kStageBytes;  // 1024-aligned: stages are multiples of 1 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + Cfg::kOutStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tmem_full_bar = bars + 2 * Cfg::kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* affine_smem = reinterpret_cast<float*>(out_stage + Cfg::kOutStageBytes + Cfg::kBarBytes);  // scale | shift

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.a_seqs * p.m_tiles_per_seq * p.n_tiles;
  const int num_kb = p.taps * p.kblocks_per_tap;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.out_tma) tma_prefetch_desc(&tmC);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 32 * kEpiWarps);
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
  pdl_enter_long<2>(total_tiles <= (int)gridDim.x);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int a_row_off = p.dyn != nullptr ? p.dyn[0] : p.a_row_off;  // streaming: ring position read on the device
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(tile, p);
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / p.kblocks_per_tap;
          const int kc = kb - tap * p.kblocks_per_tap;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          tma_load_3d(sa, &tmA, &full_bar[stage], kc * kElemsPerKBlock,
                      tc.t0 + a_row_off + tap * p.tap_row_step, tc.seq);
          if (BMN) {
            // rows kc*64.. of the [c_out][taps * c_in] weights, columns of this tap's N tile, 64 at a time
#pragma unroll
            for (int g = 0; g < BN / 64; ++g)
              tma_load_2d(sa + Cfg::kABytes + g * 8192, &tmB, &full_bar[stage],
                          tap * p.b_tap_col_step + tc.n0 * BN + g * 64, kc * kElemsPerKBlock);
          } else {
            tma_load_2d(sa + Cfg::kABytes, &tmB, &full_bar[stage], kb * kElemsPerKBlock, tc.n0 * BN);
          }
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
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t adesc = make_kmajor_sw128_desc(sa);
          const uint64_t bdesc = BMN ? make_mnmajor_sw128_desc_b(sa + Cfg::kABytes, 8192)
                                     : make_kmajor_sw128_desc(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kMmasPerStage; ++k) {
            // advance 32 bytes of K inside the swizzle span: +2 in the (address >> 4) field; an MN-major operand
            // advances 16 k-rows = two 8-row groups = 2048 bytes
            const uint64_t bk = BMN ? (uint64_t)(128 * k) : (uint64_t)(2 * k);
            if (ET::kBytes == 2)
              umma_f16_ss(d_tmem, adesc + 2 * k, bdesc + bk, kIdesc, (kb | k) != 0);
            else
              umma_tf32_ss(d_tmem, adesc + 2 * k, bdesc + bk, kIdesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    // The epilogue of a 128 x BN tile is latency bound (TMEM load -> math -> residual / store), so two warps share
    // each TMEM lane quadrant and split the columns: warp `epi` handles chunks [half * kChunks, (half + 1) * kChunks).
    constexpr int kChunks = BN / 64;              // 32-column chunks per epilogue warp
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may access (hardware: warp id % 4)
    const int epi = warp - 2;
    const int half = epi >> 2;
    const int row = quad * 32 + lane;
    unsigned out_buf = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    // Train-mode BatchNorm statistics: lane l owns column (chunk * 32 + l) of this warp's column half and keeps its
    // fp32 sum / sum of squares in registers across all tiles of the launch; flushed with double atomics when the CTA
    // moves to another column tile (with gridDim.x a multiple of n_tiles: once, at the end)
    float st_s[4] = {0.f, 0.f, 0.f, 0.f}, st_q[4] = {0.f, 0.f, 0.f, 0.f};
    static_assert(kChunks <= 4, "statistics accumulators");
    int stat_n0 = -1;
    auto stat_flush = [&]() {
      if (stat_n0 >= 0) {
#pragma unroll
        for (int ci = 0; ci < kChunks; ++ci) {
          const int col = stat_n0 * BN + (half * kChunks + ci) * 32 + lane;
          atomicAdd(p.stat_sum + col, (double)st_s[ci]);
          atomicAdd(p.stat_sqsum + col, (double)st_q[ci]);
          st_s[ci] = st_q[ci] = 0.f;
        }
      }
    };
    int affine_n0 = -1;
    // streaming (CausalStream under a CUDA graph): residual row, output row and mirror output row come from a device
    // table that a small kernel advances once per frame, so the captured launch never changes
    const int res_row_off_dyn = p.dyn != nullptr ? p.dyn[1] : p.res_row_off;
    const int out_row_off = p.dyn != nullptr ? p.dyn[2] : 0;
    const int out_row_off2 = p.dyn != nullptr ? p.dyn[3] : -1;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(tile, p);
      if (p.stat_sum != nullptr && tc.n0 != stat_n0) {
        stat_flush();
        stat_n0 = tc.n0;
      }
      if (p.shift != nullptr && tc.n0 != affine_n0) {
        // per-channel scale / shift of this column tile -> shared memory (once per launch when the grid is a multiple
        // of n_tiles). Named barrier 1 over the epilogue threads on both sides: nobody still reads the old tile's
        // values, everybody sees the new ones.
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
        const int e = threadIdx.x - 64;
        for (int j = e; j < BN; j += 32 * kEpiWarps) {
          if (p.scale != nullptr) affine_smem[j] = __ldg(p.scale + tc.n0 * BN + j);
          affine_smem[BN + j] = __ldg(p.shift + tc.n0 * BN + j);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
        affine_n0 = tc.n0;
      }
      const int t = tc.t0 + row;
      const bool row_ok = t < p.rows_out;
      const long long out_off = (long long)tc.seq * p.out_seq_stride + (long long)t * p.out_row_stride;
      const long long res_row = (long long)t * p.res_row_mul + res_row_off_dyn;
      const bool res_row_ok = p.res_rows <= 0 || (res_row >= 0 && res_row < p.res_rows);
      const long long res_off =
          (long long)tc.seq * p.res_seq_stride + res_row * p.res_row_stride - p.res_col_off;
      // 16-bit residual rows are prefetched one 32-column chunk ahead (the loads of chunk 0 go out before the wait for
      // the accumulator): their HBM / L2 latency was the dominant epilogue stall
      const bool res_any = ET::kBytes == 2 && p.res != nullptr && row_ok && res_row_ok;
      auto res_in_window = [&](int c) {
        const int col0 = tc.n0 * BN + c * 32;
        return p.res_cols <= 0 || (col0 >= p.res_col_off && col0 < p.res_col_off + p.res_cols);
      };
      auto res_load = [&](int c, uint4 (&r)[4]) {
        if (res_any && res_in_window(c)) {
          const uint4* r4 = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(p.res) + res_off +
                                                           tc.n0 * BN + c * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) r[j] = __ldg(r4 + j);
        }
      };
      uint4 rcur[4], rnext[4];
      if (res_any) {
        // the whole residual span of this thread's row (kChunks x 64 B) goes to L2 now; the register prefetch below is
        // only one chunk deep, which covers an L2 hit but not a DRAM miss
        const char* rp = reinterpret_cast<const char*>(static_cast<const uint16_t*>(p.res) + res_off + tc.n0 * BN +
                                                       half * kChunks * 32);
#pragma unroll
        for (int k = 0; k < kChunks * 64; k += 128)
          if (res_in_window(half * kChunks + k / 64)) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + k));
      }
      res_load(half * kChunks, rcur);

      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();

#pragma unroll
      for (int cq = 0; cq < kChunks; ++cq) {
        const int c = half * kChunks + cq;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + acc * BN + c * 32 + (static_cast<uint32_t>(quad * 32) << 16), v);
        if (c + 1 < (half + 1) * kChunks) res_load(c + 1, rnext);
        tmem_wait_ld();
        const int col0 = tc.n0 * BN + c * 32;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);

        if (p.scale != nullptr) {
          const float4* sc4 = reinterpret_cast<const float4*>(affine_smem + c * 32);
          const float4* sh4 = reinterpret_cast<const float4*>(affine_smem + BN + c * 32);
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
          // shift only: the per-channel scale was folded into the packed weight rows (eval-mode BatchNorm, bias)
          const float4* sh4 = reinterpret_cast<const float4*>(affine_smem + BN + c * 32);
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
        if (row_ok) {
          if (p.res != nullptr && res_row_ok &&
              (p.res_cols <= 0 || (col0 >= p.res_col_off && col0 < p.res_col_off + p.res_cols))) {
            if (ET::kBytes == 2) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 r = rcur[j];
                const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 x = unpack2(rr[e], std::integral_constant < int, DT == VP3D_TF32 ? VP3D_F16 : DT > {});
                  f[8 * j + 2 * e + 0] += x.x;
                  f[8 * j + 2 * e + 1] += x.y;
                }
              }
            } else {
              const float4* r4 = reinterpret_cast<const float4*>(static_cast<const float*>(p.res) + res_off + col0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 r = __ldg(r4 + j);
                f[4 * j + 0] += r.x;
                f[4 * j + 1] += r.y;
                f[4 * j + 2] += r.z;
                f[4 * j + 3] += r.w;
              }
            }
          }
          if (p.out_f32) {
            if (p.out_round_tf32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = round_tf32(f[j]);
            }
            float* o = static_cast<float*>(p.out) + out_off + col0;
            if (col0 + 32 <= p.n_valid && (p.out_row_stride & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                reinterpret_cast<float4*>(o)[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.n_valid) o[j] = f[j];
            }
          }
        }
        if (!p.out_f32) {
          // element-typed output (fp16 / bf16): the warp's 32 rows x 32 columns (64 B per row) are staged in shared
          // memory in the SWIZZLE_64B layout (conflict-free 16-byte stores) and leave as ONE coalesced TMA store;
          // TMA clips rows past the end of the sequence, so no row mask is needed here
          constexpr int D16 = (DT == VP3D_TF32) ? VP3D_F16 : DT;
          const unsigned b = out_buf++ & 1u;   // ring of 2 staging buffers
          uint8_t* my_stage = out_stage + b * Cfg::kOutBufBytes + epi * (32 * 64);
          if (lane == 0) tma_store_wait_read<1>();  // the store that last used this buffer has drained it
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            st_shared_v4(my_stage + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4),
                         pack2(f[8 * j + 0], f[8 * j + 1], std::integral_constant<int, D16>{}),
                         pack2(f[8 * j + 2], f[8 * j + 3], std::integral_constant<int, D16>{}),
                         pack2(f[8 * j + 4], f[8 * j + 5], std::integral_constant<int, D16>{}),
                         pack2(f[8 * j + 6], f[8 * j + 7], std::integral_constant<int, D16>{}));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (p.stat_sum != nullptr) {
            // Per-channel sum / sum of squares of the values as stored (16-bit): the tile is in shared memory now,
            // so lane l walks column l down the warp's rows -- 32 lanes read one 64-byte row per step, conflict free.
            // (The first version reduced the fp32 accumulators with a 62-shuffle transposing butterfly per chunk.)
            int valid = p.rows_out - (tc.t0 + quad * 32);
            valid = valid < 0 ? 0 : (valid > 32 ? 32 : valid);
            float cs = 0.f, cq = 0.f;
            const uint32_t col_off = (lane & 7) * 2, col_chunk = lane >> 3;
#pragma unroll 8
            for (int r = 0; r < valid; ++r) {
              const uint16_t h = *reinterpret_cast<const uint16_t*>(my_stage + r * 64 + ((col_chunk ^ ((r >> 1) & 3)) << 4) +
                                                                    col_off);
              const float v = (D16 == VP3D_F16) ? __half2float(__ushort_as_half(h)) : __uint_as_float((uint32_t)h << 16);
              cs += v;
              cq = fmaf(v, v, cq);
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
            tma_store_3d(&tmC, my_stage, tc.n0 * BN + c * 32, tc.t0 + quad * 32 + out_row_off, tc.seq);
            if (out_row_off2 >= 0)   // mirror slot of a streaming ring
              tma_store_3d(&tmC, my_stage, tc.n0 * BN + c * 32, tc.t0 + quad * 32 + out_row_off2, tc.seq);
            tma_store_commit();
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) rcur[j] = rnext[j];
      }
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.stat_sum != nullptr) stat_flush();
    if (lane == 0) tma_store_wait_all<0>();  // every output tile of this warp has reached global memory
  }

  pdl_tail_trigger(total_tiles <= (int)gridDim.x);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
  if (p.fin.sum != nullptr)
    bn_finalize_tail(p.fin, p.fin_counter, p.n_tiles * BN, reinterpret_cast<volatile int*>(reinterpret_cast<uint8_t*>(bars) + Cfg::kBarBytes - 4));
}

template <int DT, int BN, bool BMN>
static cudaError_t launch_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                              const ConvGemmParams& p, int grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static std::atomic<unsigned long long> attr_done{0};   // one bit per device ordinal
  if (cudaError_t e = set_max_smem_once(reinterpret_cast<const void*>(conv_gemm_kernel<DT, BN, BMN>), Cfg::kSmemBytes,
                                        attr_done))
    return e;
  launch_k(conv_gemm_kernel<DT, BN, BMN>, dim3(grid), dim3(kNumThreads), Cfg::kSmemBytes, stream, tmA, tmB, tmC, p);
  return cudaGetLastError();
}

cudaError_t launch_conv_gemm(int dtype, int block_n, int w_mn_major, const CUtensorMap& tmA, const CUtensorMap& tmB,
                             const CUtensorMap& tmC, const ConvGemmParams& p, int grid, cudaStream_t stream) {
  if (w_mn_major) {
    if (block_n == 256 && dtype == VP3D_F16) return launch_one<VP3D_F16, 256, true>(tmA, tmB, tmC, p, grid, stream);
    if (block_n == 256 && dtype == VP3D_BF16) return launch_one<VP3D_BF16, 256, true>(tmA, tmB, tmC, p, grid, stream);
    if (block_n == 64 && dtype == VP3D_F16) return launch_one<VP3D_F16, 64, true>(tmA, tmB, tmC, p, grid, stream);
    if (block_n == 64 && dtype == VP3D_BF16) return launch_one<VP3D_BF16, 64, true>(tmA, tmB, tmC, p, grid, stream);
    return cudaErrorInvalidValue;
  }
  if (block_n == 256) {
    if (dtype == VP3D_F16) return launch_one<VP3D_F16, 256, false>(tmA, tmB, tmC, p, grid, stream);
    if (dtype == VP3D_BF16) return launch_one<VP3D_BF16, 256, false>(tmA, tmB, tmC, p, grid, stream);
    if (dtype == VP3D_TF32) return launch_one<VP3D_TF32, 256, false>(tmA, tmB, tmC, p, grid, stream);
  } else if (block_n == 64) {
    if (dtype == VP3D_F16) return launch_one<VP3D_F16, 64, false>(tmA, tmB, tmC, p, grid, stream);
    if (dtype == VP3D_BF16) return launch_one<VP3D_BF16, 64, false>(tmA, tmB, tmC, p, grid, stream);
    if (dtype == VP3D_TF32) return launch_one<VP3D_TF32, 64, false>(tmA, tmB, tmC, p, grid, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace vp3d
