// Counter-based dropout shared by the HBM-bound train-mode passes (train.cu) and the GEMM epilogue (conv_gemm2.cu):
// the keep decision of an element is a pure function of (seed, layer stream, training step, row, channel), so the
// backward recomputes the forward mask instead of storing it, and a fused epilogue draws the same mask as the
// stand-alone pass would.
#pragma once

#include <cstdint>

#include "kernels.h"

namespace vp3d {

// Philox4x32-7 (Salmon et al., SC'11; 7 rounds is the fastest variant that passes BigCrush): counter (c0..c3), key.
__device__ __forceinline__ uint4 philox4x32_7(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// One Philox block = 16 random bytes = the keep decisions of 2 adjacent rows x 8 channels (rows 2P and 2P+1 of channel
// group grp): keep iff byte >= thresh, thresh = round(p * 256). The drop probability is therefore p quantised to
// 1/256 (exact for the reference default 0.25) and kept values are scaled by 256 / (256 - thresh), so E[mask] = 1.
struct DropCtx {
  bool on;
  uint32_t thresh;
  float keep_scale;
  uint2 key;
  uint32_t s_lo, s_hi;
};
__device__ __forceinline__ DropCtx make_drop(const DropoutParams& d) {
  DropCtx c;
  c.thresh = (uint32_t)(d.p * 256.f + 0.5f);
  c.on = c.thresh > 0;
  c.keep_scale = 256.f / (256.f - (float)c.thresh);
  c.key = make_uint2((uint32_t)d.seed, (uint32_t)(d.seed >> 32));
  // counter words 2, 3: layer stream and the training step (device counter, so graph replays differ)
  const unsigned long long step = d.step_counter != nullptr ? *d.step_counter : 0ull;
  c.s_lo = (uint32_t)d.stream ^ (uint32_t)(step >> 32);
  c.s_hi = (uint32_t)step;
  return c;
}
// random bytes of row pair `pair` (rows 2*pair, 2*pair + 1), channel group grp: .x,.y -> even row, .z,.w -> odd row
__device__ __forceinline__ uint4 drop_bits(const DropCtx& d, long long pair, int grp) {
  return philox4x32_7(make_uint4((uint32_t)pair, (uint32_t)((unsigned long long)pair >> 32) ^ (grp * 0x9E3779B1u),
                                 d.s_lo, d.s_hi), d.key);
}
__device__ __forceinline__ void drop_mult8(const DropCtx& d, uint32_t w0, uint32_t w1, float (&m)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = ((w0 >> (8 * i)) & 0xFFu) >= d.thresh ? d.keep_scale : 0.f;
    m[4 + i] = ((w1 >> (8 * i)) & 0xFFu) >= d.thresh ? d.keep_scale : 0.f;
  }
}

}  // namespace vp3d
