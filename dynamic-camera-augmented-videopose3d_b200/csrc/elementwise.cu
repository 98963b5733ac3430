// Layout / bookkeeping kernels around the conv GEMM (all HBM-bound, tiny next to the GEMMs):
//   pack_rows      fp32 [rows][c]  ->  operand type [rows][c_pad]   (model input (N,T,J*F) -> padded channels-last)
//   pack_weight    nn.Conv1d weight (c_out, c_in, taps) fp32 -> K-major packed operand [n_pad][taps][c_in_pad]
//                  (or its transpose for data-gradient GEMMs)
//   bn_fold        eval-mode BatchNorm1d -> per-channel scale/shift (TemporalModel.py:32,117,119 in eval())
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "kernels.h"
#include "pdl.cuh"

namespace vp3d {

template <int DT>
__device__ __forceinline__ void store_elem(void* dst, long long i, float v);
template <>
__device__ __forceinline__ void store_elem<VP3D_F16>(void* dst, long long i, float v) {
  static_cast<__half*>(dst)[i] = __float2half_rn(v);
}
template <>
__device__ __forceinline__ void store_elem<VP3D_BF16>(void* dst, long long i, float v) {
  static_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16_rn(v);
}
template <>
__device__ __forceinline__ void store_elem<VP3D_TF32>(void* dst, long long i, float v) {
  uint32_t u;  // fp32 container holding a TF32-representable value (round to nearest; the MMA would truncate)
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  static_cast<float*>(dst)[i] = __uint_as_float(u);
}

template <int DT>
__global__ void __launch_bounds__(256)
pack_rows_kernel(const float* __restrict__ src, void* __restrict__ dst, long long rows, int c, int c_pad) {
  pdl_enter();
  // blockDim.x = c_pad threads across a row (coalesced in src and dst), blockDim.y rows per block, grid-stride over rows
  const int k = threadIdx.x;
  for (long long r = (long long)blockIdx.x * blockDim.y + threadIdx.y; r < rows; r += (long long)gridDim.x * blockDim.y)
    for (int kk = k; kk < c_pad; kk += blockDim.x)
      store_elem<DT>(dst, r * c_pad + kk, kk < c ? __ldg(src + r * c + kk) : 0.f);
}

// 16-bit destination, c_pad % 8 == 0: a thread owns 8 consecutive output channels of one row = ONE 16-byte store
// (the element-per-thread kernel above issues 2-byte stores: 35 us for the 1024 x 243 x 34 training batch, where the
// bytes moved would take 10 us)
template <int DT>
__global__ void __launch_bounds__(256)
pack_rows_vec8_kernel(const float* __restrict__ src, uint4* __restrict__ dst, long long rows, int c, int groups,
                      int ones_col) {
  pdl_enter();
  const long long total = rows * groups;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / groups;
    const int k0 = (int)(i - r * groups) * 8;
    const float* s = src + r * c + k0;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = k0 + j < c ? __ldg(s + j) : (k0 + j == ones_col ? 1.f : 0.f);
    uint4 o;
    if (DT == VP3D_F16) {
      __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
      __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
      o = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                     *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
    } else {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
      o = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                     *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
    }
    dst[i] = o;
  }
}

// dst[n][tap][ci] = w[n][ci][tap]                  (transpose == 0; rows = output channels, K = tap*c_in_pad + ci)
// dst[tap*c_in_pad + ci][co] = w[co][ci][tap]      (transpose == 1; rows = (tap, ci) with c_in_pad = rows_pad / taps)
// dst[ci][tap*k_pad_per_tap + co] = w[co][ci][tap] (transpose == 2; rows = input channels, K = (tap, output channel))
template <int DT>
__global__ void __launch_bounds__(256)
pack_weight_kernel(const float* __restrict__ w, void* __restrict__ dst, int c_out, int c_in, int taps, int rows_pad,
                   int k_pad_per_tap, int transpose) {
  pdl_enter();
  const long long k_total = transpose == 1 ? k_pad_per_tap : (long long)taps * k_pad_per_tap;
  const long long total = (long long)rows_pad * k_total;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int c_in_pad = rows_pad / taps;  // transpose == 1 only
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / k_total;
    const long long k = i - r * k_total;
    float v = 0.f;
    if (transpose == 0) {
      const int tap = (int)(k / k_pad_per_tap);
      const int ci = (int)(k - (long long)tap * k_pad_per_tap);
      if (r < c_out && ci < c_in) v = __ldg(w + ((long long)r * c_in + ci) * taps + tap);
    } else if (transpose == 1) {
      const int tap = (int)(r / c_in_pad);
      const int ci = (int)(r - (long long)tap * c_in_pad);
      if (tap < taps && ci < c_in && k < c_out) v = __ldg(w + ((long long)k * c_in + ci) * taps + tap);
    } else {
      const int tap = (int)(k / k_pad_per_tap);
      const int co = (int)(k - (long long)tap * k_pad_per_tap);
      if (r < c_in && co < c_out) v = __ldg(w + ((long long)co * c_in + r) * taps + tap);
    }
    store_elem<DT>(dst, i, v);
  }
}

__global__ void __launch_bounds__(256)
bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
               const float* __restrict__ var, float eps, float* __restrict__ scale, float* __restrict__ shift, int c,
               int c_pad) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c_pad) return;
  if (i < c) {
    const float s = gamma[i] / sqrtf(var[i] + eps);
    scale[i] = s;
    shift[i] = beta[i] - mean[i] * s;
  } else {
    scale[i] = 0.f;
    shift[i] = 0.f;
  }
}

// transpose == 0 fast path: thread = (n, ci_pad index); reads the `taps` adjacent fp32 of w[n][ci][:] (a warp reads one
// contiguous span) and writes one element into each tap's K block (contiguous across the warp)
template <int DT>
__global__ void __launch_bounds__(256)
pack_weight_fwd_kernel(const float* __restrict__ w, const float* __restrict__ row_scale, void* __restrict__ dst,
                       int c_out, int c_in, int taps, int rows_pad, int k_pad_per_tap) {
  pdl_enter();
  const int total = rows_pad * k_pad_per_tap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / k_pad_per_tap;
    const int ci = i - n * k_pad_per_tap;
    const bool ok = n < c_out && ci < c_in;
    const float sc = (ok && row_scale != nullptr) ? __ldg(row_scale + n) : 1.f;
    const float* src = w + ((long long)n * c_in + ci) * taps;
    const long long d0 = (long long)n * taps * k_pad_per_tap + ci;
    for (int tap = 0; tap < taps; ++tap)
      store_elem<DT>(dst, d0 + (long long)tap * k_pad_per_tap, ok ? __ldg(src + tap) * sc : 0.f);
  }
}

// Streaming ring bookkeeping (vp3d_b200/streaming.py): see vp3d_stream_advance in include/vp3d_b200.h.
__global__ void stream_advance_kernel(long long* step, int n_rings, const int* ring_len, const int* ring_dil,
                                      const int* ring_taps, int rows_per_slot, int* table, int n_launch,
                                      const int* launch_desc, int* launch_table) {
  pdl_enter();
  __shared__ int ring[64][4];
  const long long t = *step;          // frame index of the step being issued
  const int i = threadIdx.x;
  if (i < n_rings) {
    const int L = ring_len[i];
    const int q = (int)(t % L) + L;   // upper copy of the current slot: taps q - k*d never wrap
    ring[i][0] = (q - (ring_taps[i] - 1) * ring_dil[i]) * rows_per_slot;
    ring[i][1] = q * rows_per_slot;
    ring[i][2] = q * rows_per_slot;
    ring[i][3] = (q - L) * rows_per_slot;
#pragma unroll
    for (int k = 0; k < 4; ++k) table[4 * i + k] = ring[i][k];
  }
  __syncthreads();
  if (i < n_launch) {
    // launch i reads ring launch_desc[3i] as its A operand, ring [3i+1] as residual, writes ring [3i+2] (-1: none)
    const int ra = launch_desc[3 * i], rr = launch_desc[3 * i + 1], ro = launch_desc[3 * i + 2];
    launch_table[4 * i + 0] = ra >= 0 ? ring[ra][0] : 0;
    launch_table[4 * i + 1] = rr >= 0 ? ring[rr][1] : 0;
    launch_table[4 * i + 2] = ro >= 0 ? ring[ro][2] : 0;
    launch_table[4 * i + 3] = ro >= 0 ? ring[ro][3] : -1;
  }
  if (i == 0) *step = t + 1;
}

template <int DT>
__global__ void __launch_bounds__(256)
ring_write_kernel(const float* __restrict__ src, void* __restrict__ ring, const int* __restrict__ entry, long long rows,
                  int c, int c_pad) {
  pdl_enter();
  const long long r_hi = entry[2], r_lo = entry[3];
  const long long total = rows * c_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c_pad;
    const int k = (int)(i - r * c_pad);
    const float v = k < c ? __ldg(src + r * c + k) : 0.f;
    store_elem<DT>(ring, (r_hi + r) * c_pad + k, v);
    store_elem<DT>(ring, (r_lo + r) * c_pad + k, v);
  }
}

cudaError_t launch_stream_advance(long long* step, int n_rings, const int* ring_len, const int* ring_dil,
                                  const int* ring_taps, int rows_per_slot, int* table, int n_launch,
                                  const int* launch_desc, int* launch_table, cudaStream_t stream) {
  launch_k(stream_advance_kernel, dim3(1), dim3(64), 0, stream, step, n_rings, ring_len, ring_dil, ring_taps, rows_per_slot, table,
                                              n_launch, launch_desc, launch_table);
  return cudaGetLastError();
}

static int ew_grid(long long total, int sm_count);
cudaError_t launch_ring_write(int dtype, const float* src, void* ring, const int* table, long long rows, int c, int c_pad,
                              int sm_count, cudaStream_t stream) {
  const int grid = ew_grid(rows * c_pad, sm_count);
  if (dtype == VP3D_F16) launch_k(ring_write_kernel<VP3D_F16>, dim3(grid), dim3(256), 0, stream, src, ring, table, rows, c, c_pad);
  else if (dtype == VP3D_BF16) launch_k(ring_write_kernel<VP3D_BF16>, dim3(grid), dim3(256), 0, stream, src, ring, table, rows, c, c_pad);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

static int ew_grid(long long total, int sm_count) {
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ones_col >= 0 (16-bit, c_pad % 8 == 0 only): that padding column holds 1.0 -- the Gram matrix of the packed rows then
// carries the column sums and the row count (expand.cu)
cudaError_t launch_pack_rows(int dtype, const float* src, void* dst, long long rows, int c, int c_pad, int sm_count,
                             cudaStream_t stream, int ones_col) {
  if ((dtype == VP3D_F16 || dtype == VP3D_BF16) && c_pad % 8 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const int groups = c_pad / 8;
    const int g = ew_grid(rows * groups, sm_count);
    if (dtype == VP3D_F16)
      launch_k(pack_rows_vec8_kernel<VP3D_F16>, dim3(g), dim3(256), 0, stream, src, static_cast<uint4*>(dst), rows, c, groups, ones_col);
    else
      launch_k(pack_rows_vec8_kernel<VP3D_BF16>, dim3(g), dim3(256), 0, stream, src, static_cast<uint4*>(dst), rows, c, groups, ones_col);
    return cudaGetLastError();
  }
  if (ones_col >= 0) return cudaErrorInvalidValue;
  const int bx = c_pad >= 256 ? 256 : ((c_pad + 31) / 32) * 32;
  const dim3 block(bx, 256 / bx > 0 ? 256 / bx : 1);
  long long gx = (rows + block.y - 1) / block.y;
  if (gx > (long long)sm_count * 16) gx = (long long)sm_count * 16;
  if (gx < 1) gx = 1;
  const int grid = (int)gx;
  if (dtype == VP3D_F16) launch_k(pack_rows_kernel<VP3D_F16>, dim3(grid), dim3(block), 0, stream, src, dst, rows, c, c_pad);
  else if (dtype == VP3D_BF16) launch_k(pack_rows_kernel<VP3D_BF16>, dim3(grid), dim3(block), 0, stream, src, dst, rows, c, c_pad);
  else if (dtype == VP3D_TF32) launch_k(pack_rows_kernel<VP3D_TF32>, dim3(grid), dim3(block), 0, stream, src, dst, rows, c, c_pad);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_pack_weight(int dtype, const float* w, void* dst, int c_out, int c_in, int taps, int rows_pad,
                               int k_pad_per_tap, int transpose, int sm_count, cudaStream_t stream,
                               const float* row_scale) {
  const long long k_total = transpose == 1 ? k_pad_per_tap : (long long)taps * k_pad_per_tap;
  if (transpose == 0 && (long long)rows_pad * k_pad_per_tap < (1LL << 31)) {
    const int g = ew_grid((long long)rows_pad * k_pad_per_tap, sm_count);
    if (dtype == VP3D_F16)
      launch_k(pack_weight_fwd_kernel<VP3D_F16>, dim3(g), dim3(256), 0, stream, w, row_scale, dst, c_out, c_in, taps, rows_pad, k_pad_per_tap);
    else if (dtype == VP3D_BF16)
      launch_k(pack_weight_fwd_kernel<VP3D_BF16>, dim3(g), dim3(256), 0, stream, w, row_scale, dst, c_out, c_in, taps, rows_pad, k_pad_per_tap);
    else if (dtype == VP3D_TF32)
      launch_k(pack_weight_fwd_kernel<VP3D_TF32>, dim3(g), dim3(256), 0, stream, w, row_scale, dst, c_out, c_in, taps, rows_pad, k_pad_per_tap);
    else
      return cudaErrorInvalidValue;
    return cudaGetLastError();
  }
  const int grid = ew_grid((long long)rows_pad * k_total, sm_count);
  if (dtype == VP3D_F16)
    launch_k(pack_weight_kernel<VP3D_F16>, dim3(grid), dim3(256), 0, stream, w, dst, c_out, c_in, taps, rows_pad, k_pad_per_tap, transpose);
  else if (dtype == VP3D_BF16)
    launch_k(pack_weight_kernel<VP3D_BF16>, dim3(grid), dim3(256), 0, stream, w, dst, c_out, c_in, taps, rows_pad, k_pad_per_tap, transpose);
  else if (dtype == VP3D_TF32)
    launch_k(pack_weight_kernel<VP3D_TF32>, dim3(grid), dim3(256), 0, stream, w, dst, c_out, c_in, taps, rows_pad, k_pad_per_tap, transpose);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                           float* scale, float* shift, int c, int c_pad, cudaStream_t stream) {
  launch_k(bn_fold_kernel, dim3((c_pad + 255) / 256), dim3(256), 0, stream, gamma, beta, mean, var, eps, scale, shift, c, c_pad);
  return cudaGetLastError();
}

}  // namespace vp3d
