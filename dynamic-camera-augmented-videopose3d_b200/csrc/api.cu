// C-ABI layer: argument validation, TMA tensor-map encoding, kernel launches. No torch types, no allocation.
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "kernels.h"

namespace vp3d {
cudaError_t launch_project_points(const float* X, float* out3, float* out2, long long n_pts, const float* q,
                                  const float* t, const float* cam, long long pts_per_q, long long pts_per_cam,
                                  int mode, int sm_count, cudaStream_t stream);
cudaError_t launch_project_bwd(const float* X, const float* cam, const float* g, long long n_pts, long long pts_per_cam,
                               int linear, float* gx, int sm_count, cudaStream_t stream);
cudaError_t launch_project_windows(const float* x, const float* q, const float* t, const float* cam,
                                   const long long* seq_start, const long long* seq_len, const int* sample_seq,
                                   const long long* sample_start, int batch, int joints, int chunk, int pad, int shift,
                                   int root_relative, int linear, float* out2, float* target3, float* cam3x4,
                                   int sm_count, cudaStream_t stream);
cudaError_t launch_mpjpe_fwd(const float* pred, const float* tgt, long long n_joints, const float* w, long long T,
                             long long J, long long s_n, long long s_t, long long s_j, double* partial, float* out,
                             int sm_count, cudaStream_t stream);
cudaError_t launch_mpjpe_bwd(const float* pred, const float* tgt, const float* grad_out, long long n_joints,
                             const float* w, long long T, long long J, long long s_n, long long s_t, long long s_j,
                             float* grad_pred, int sm_count, cudaStream_t stream);
cudaError_t launch_mpjpe_nd_fwd(const float* pred, const float* tgt, long long n_pts, int D, const float* w, long long T,
                                long long J, long long s_n, long long s_t, long long s_j, double* partial, float* out,
                                int sm_count, cudaStream_t stream);
cudaError_t launch_mpjpe_nd_bwd(const float* pred, const float* tgt, const float* grad_out, long long n_pts, int D,
                                const float* w, long long T, long long J, long long s_n, long long s_t, long long s_j,
                                float* grad_pred, int sm_count, cudaStream_t stream);
cudaError_t launch_n_mpjpe_fwd(const float* pred, const float* tgt, long long n_poses, int J, double* partial,
                               float* out, int sm_count, cudaStream_t stream);
cudaError_t launch_p_mpjpe_fwd(const float* pred, const float* tgt, long long n_poses, int J, double* partial, float* out,
                               int sm_count, cudaStream_t stream);
cudaError_t launch_velocity_error(const float* pred, const float* tgt, long long T, long long inner, int D,
                                  double* partial, float* out, int sm_count, cudaStream_t stream);
cudaError_t launch_n_mpjpe_bwd(const float* pred, const float* tgt, const float* grad_out, long long n_poses, int J,
                               float* grad_pred, int sm_count, cudaStream_t stream);
cudaError_t launch_reproj_fwd(const float* pose, const float* traj, long long n_pts, long long pts_per_traj,
                              const float* cam, long long pts_per_cam, int linear, const float* tgt, double* partial,
                              float* out, int sm_count, cudaStream_t stream);
cudaError_t launch_reproj_bwd(const float* pose, const float* traj, long long n_pts, long long pts_per_traj,
                              const float* cam, long long pts_per_cam, int linear, const float* tgt,
                              const float* grad_out, float* grad_pose, float* grad_traj, int sm_count,
                              cudaStream_t stream);
cudaError_t launch_pack_rows(int dtype, const float* src, void* dst, long long rows, int c, int c_pad, int sm_count,
                             cudaStream_t stream, int ones_col = -1);
cudaError_t launch_pack_weight(int dtype, const float* w, void* dst, int c_out, int c_in, int taps, int rows_pad,
                               int k_pad_per_tap, int transpose, int sm_count, cudaStream_t stream,
                               const float* row_scale = nullptr);
cudaError_t launch_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                           float* scale, float* shift, int c, int c_pad, cudaStream_t stream);
}  // namespace vp3d

namespace vp3d {
// Programmatic dependent launch for every kernel of the library (pdl.cuh). Initial value from VP3D_PDL ("0": off).
int g_pdl = [] {
  const char* e = std::getenv("VP3D_PDL");
  return (e != nullptr && std::strcmp(e, "0") == 0) ? 0 : 1;
}();
}  // namespace vp3d

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  return fail(VP3D_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

// Immutable per-device facts, resolved once (no other global state).
struct DeviceInfo {
  int sm_count = 0, sm_real = 0, cc_major = 0, cc_minor = 0;
  bool ok = false;
  bool pair_refused = false;   // THIS device refused a cluster launch once (MPS / partitioned GPU): single-CTA kernel from then on
};
// Upper bound on the SMs the persistent kernels size their grids for (0: all). Data-parallel training sets it a few
// SMs below the device's count so that NCCL's all-reduce CTAs find free SMs next to a running GEMM instead of queueing
// behind it (or, worse, a statically scheduled persistent GEMM queueing behind them).
// K1 on CTA pairs: 0 = never, 1 = launches of at least half a wave of tiles (default), 2 = whenever the launch is
// supported (tests on small shapes). Initial value from VP3D_K1_2CTA ("0" / "force").
int g_pair_mode = [] {
  const char* e = std::getenv("VP3D_K1_2CTA");
  if (e == nullptr) return 1;
  return std::strcmp(e, "force") == 0 ? 2 : (std::strcmp(e, "0") == 0 ? 0 : 1);
}();
// Tile schedule of the CTA-pair kernel for launches without statistics of at least two waves: 0 = static persistent
// (default), 1 = dynamic (cluster launch control, conv_gemm2.cu). Initial value from VP3D_SCHED ("dynamic").
int g_sched_mode = [] {
  const char* e = std::getenv("VP3D_SCHED");
  if (e == nullptr) return 0;
  return std::strcmp(e, "dynamic") == 0 ? 1 : (std::strcmp(e, "dynamic-all") == 0 ? 2 : 0);
}();
int g_sm_limit = [] {
  const char* e = std::getenv("VP3D_SM_LIMIT");
  return e != nullptr ? std::atoi(e) : 0;
}();
constexpr int kMaxDevices = 64;
DeviceInfo g_dev[kMaxDevices];
std::mutex g_dev_mu;

int device_info(DeviceInfo** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev < 0 || dev >= kMaxDevices) return fail(VP3D_ERR_INVALID, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lock(g_dev_mu);
  DeviceInfo& d = g_dev[dev];
  if (!d.ok) {
    if ((e = cudaDeviceGetAttribute(&d.sm_real, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev)) != cudaSuccess)
      return cuda_fail(e, "cudaDeviceGetAttribute");
    d.ok = true;
  }
  d.sm_count = (g_sm_limit > 0 && g_sm_limit < d.sm_real) ? g_sm_limit : d.sm_real;
  if (d.cc_major != 10)
    return fail(VP3D_ERR_UNSUPPORTED, "vp3d_b200 kernels are built for sm_100a only; device is sm_%d%d", d.cc_major,
                d.cc_minor);
  *out = &d;
  return VP3D_OK;
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_encode_once;

EncodeTiledFn get_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
  return g_encode;
}

vp3d::DropoutParams drop_of(const vp3d_dropout* d) {
  vp3d::DropoutParams dp;
  dp.p = d ? d->p : 0.f;
  dp.seed = d ? d->seed : 0ull;
  dp.stream = d ? d->stream : 0ull;
  dp.step_counter = d ? d->step_counter : nullptr;
  return dp;
}

int elem_bytes(int dtype) { return dtype == VP3D_TF32 ? 4 : 2; }
CUtensorMapDataType tm_dtype(int dtype) {
  switch (dtype) {
    case VP3D_F16: return CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    case VP3D_BF16: return CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    default: return CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  }
}

// Tiled map with a 128-byte inner box and SWIZZLE_128B; out-of-range elements read as zero.
int encode_map(CUtensorMap* tm, int dtype, int rank, const void* base, const cuuint64_t* dims,
               const cuuint64_t* strides_bytes, const cuuint32_t* box, const char* what,
               CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) return fail(VP3D_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, tm_dtype(dtype), (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(VP3D_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d (dims %llu,%llu,%llu)", what, (int)r,
                (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                (unsigned long long)(rank > 2 ? dims[2] : 0));
  return VP3D_OK;
}

}  // namespace

extern "C" {

int vp3d_version(void) { return 100; }

const char* vp3d_last_error(void) { return g_err; }

int vp3d_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  int sm = 0, maj = 0, min = 0;
  if ((e = cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess ||
      (e = cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev)) != cudaSuccess ||
      (e = cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev)) != cudaSuccess)
    return cuda_fail(e, "cudaDeviceGetAttribute");
  if (sm_count) *sm_count = sm;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  if (maj != 10) return fail(VP3D_ERR_UNSUPPORTED, "device is sm_%d%d, need sm_100", maj, min);
  return VP3D_OK;
}

int vp3d_set_pair_mode(int mode) {
  if (mode < 0 || mode > 2) return fail(VP3D_ERR_INVALID, "pair mode must be 0 (off), 1 (auto) or 2 (whenever supported)");
  g_pair_mode = mode;
  return VP3D_OK;
}

int vp3d_set_sched_mode(int mode) {
  if (mode < 0 || mode > 2)
    return fail(VP3D_ERR_INVALID, "sched mode must be 0 (static), 1 (dynamic) or 2 (dynamic for every pair launch)");
  g_sched_mode = mode;
  return VP3D_OK;
}

int vp3d_set_pdl(int on) {
  vp3d::g_pdl = on ? 1 : 0;
  return VP3D_OK;
}

int vp3d_set_sm_limit(int sms) {
  std::lock_guard<std::mutex> lock(g_dev_mu);
  g_sm_limit = sms > 0 ? sms : 0;
  return VP3D_OK;
}

int vp3d_conv_block_fwd(const vp3d_conv_args* a, void* stream) {
  if (a == nullptr) return fail(VP3D_ERR_INVALID, "args is NULL");
  if (a->dtype != VP3D_F16 && a->dtype != VP3D_BF16 && a->dtype != VP3D_TF32)
    return fail(VP3D_ERR_INVALID, "unknown dtype %d", a->dtype);
  if (a->block_n != 256 && a->block_n != 64) return fail(VP3D_ERR_INVALID, "block_n must be 256 or 64");
  const int eb = elem_bytes(a->dtype);
  const long long kblk = 128 / eb;  // elements of K per pipeline stage
  if (a->a == nullptr || a->w == nullptr || a->out == nullptr) return fail(VP3D_ERR_INVALID, "null a / w / out");
  if (a->a_seqs <= 0 || a->a_rows <= 0 || a->rows_out <= 0) return fail(VP3D_ERR_INVALID, "empty problem");
  if (a->taps < 1 || a->k_per_tap <= 0 || a->k_per_tap % kblk != 0)
    return fail(VP3D_ERR_INVALID, "k_per_tap (%lld) must be a positive multiple of %lld", a->k_per_tap, kblk);
  if (!a->w_mn_major && a->k_total != (long long)a->taps * a->k_per_tap)
    return fail(VP3D_ERR_INVALID, "k_total != taps * k_per_tap");
  if (a->w_mn_major) {
    if (a->dtype == VP3D_TF32) return fail(VP3D_ERR_INVALID, "w_mn_major needs fp16 / bf16 operands");
    if (a->w_row_stride < a->n_pad + (long long)(a->taps - 1) * a->w_tap_col_step || (a->w_row_stride * 2) % 16 != 0)
      return fail(VP3D_ERR_INVALID, "w_mn_major: bad w_row_stride");
  }
  if (a->a_kdim < a->k_per_tap) return fail(VP3D_ERR_INVALID, "a_kdim (%lld) < k_per_tap", a->a_kdim);
  if (a->n_pad <= 0 || a->n_pad % a->block_n != 0) return fail(VP3D_ERR_INVALID, "n_pad must be a multiple of block_n");
  if ((reinterpret_cast<uintptr_t>(a->a) & 15) || (reinterpret_cast<uintptr_t>(a->w) & 15) ||
      (reinterpret_cast<uintptr_t>(a->out) & 15))
    return fail(VP3D_ERR_INVALID, "a / w / out must be 16-byte aligned");
  if ((a->a_row_stride * eb) % 16 != 0 || (a->a_seq_stride * eb) % 16 != 0)
    return fail(VP3D_ERR_INVALID, "activation strides must be multiples of 16 bytes");
  if (a->scale != nullptr && a->shift == nullptr) return fail(VP3D_ERR_INVALID, "scale without shift");
  if ((a->stat_sum == nullptr) != (a->stat_sqsum == nullptr)) return fail(VP3D_ERR_INVALID, "stat_sum / stat_sqsum");
  if (a->stat_sum != nullptr && a->out_f32) return fail(VP3D_ERR_INVALID, "statistics need a 16-bit output (they are taken from the stored values)");
  if (a->dtype == VP3D_TF32 && !a->out_f32) return fail(VP3D_ERR_INVALID, "TF32 activations are fp32: set out_f32");
  if (!a->out_f32 && ((a->out_row_stride * 2) % 16 != 0 || (a->out_seq_stride * 2) % 16 != 0))
    return fail(VP3D_ERR_INVALID, "16-bit output strides must be multiples of 16 bytes");
  if (a->res != nullptr) {
    if ((reinterpret_cast<uintptr_t>(a->res) & 15) || (a->res_row_stride * eb) % 16 != 0 ||
        (a->res_seq_stride * eb) % 16 != 0)
      return fail(VP3D_ERR_INVALID, "residual must be 16-byte aligned with 16-byte strides");
  }
  if (a->out_f32 && (a->n_valid <= 0 || a->n_valid > a->n_pad)) return fail(VP3D_ERR_INVALID, "bad n_valid");
  if (a->dyn_offsets != nullptr && (a->out_f32 || a->a_seqs != 1 || a->out_rows_total < a->rows_out))
    return fail(VP3D_ERR_INVALID, "dyn_offsets: 16-bit output, one flat sequence and out_rows_total >= rows_out required");
  if (a->res_cols < 0 || a->res_col_off < 0 || a->res_cols % 32 != 0 || a->res_col_off % 32 != 0 ||
      a->res_col_off + a->res_cols > a->n_pad)
    return fail(VP3D_ERR_INVALID, "residual column window must be 32-aligned and inside n_pad");
  const bool want_drop = a->drop != nullptr && a->drop->p > 0.f;
  const bool want_epi = want_drop || a->side_mode != 0;
  if (want_drop && a->drop->p >= 1.f) return fail(VP3D_ERR_INVALID, "dropout p must be in [0, 1)");
  if (a->side_mode < 0 || a->side_mode > 2) return fail(VP3D_ERR_INVALID, "side_mode must be 0, 1 or 2");
  if (a->side_mode != 0) {
    if (a->side == nullptr || (reinterpret_cast<uintptr_t>(a->side) & 15) || (a->side_row_stride * 2) % 16 != 0 ||
        (a->side_seq_stride * 2) % 16 != 0 || a->side_rows <= 0)
      return fail(VP3D_ERR_INVALID, "side input must be 16-byte aligned with 16-byte strides and side_rows > 0");
  }

  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;

  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[3] = {(cuuint64_t)a->a_kdim, (cuuint64_t)a->a_rows, (cuuint64_t)a->a_seqs};
    cuuint64_t strides[2] = {(cuuint64_t)(a->a_row_stride * eb), (cuuint64_t)(a->a_seq_stride * eb)};
    if (a->a_seqs == 1) strides[1] = (cuuint64_t)(a->a_rows * a->a_row_stride * eb);  // unused extent, keep it legal
    cuuint32_t box[3] = {(cuuint32_t)kblk, 128, 1};
    if (int rc = encode_map(&tmA, a->dtype, 3, a->a, dims, strides, box, "activations")) return rc;
  }
  if (a->w_mn_major) {
    cuuint64_t dims[2] = {(cuuint64_t)a->w_row_stride, (cuuint64_t)a->k_per_tap};
    cuuint64_t strides[1] = {(cuuint64_t)(a->w_row_stride * eb)};
    cuuint32_t box[2] = {64, 64};
    if (int rc = encode_map(&tmB, a->dtype, 2, a->w, dims, strides, box, "weights (MN-major)")) return rc;
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)a->k_total, (cuuint64_t)a->n_pad};
    cuuint64_t strides[1] = {(cuuint64_t)(a->k_total * eb)};
    cuuint32_t box[2] = {(cuuint32_t)kblk, (cuuint32_t)a->block_n};
    if (int rc = encode_map(&tmB, a->dtype, 2, a->w, dims, strides, box, "weights")) return rc;
  }

  CUtensorMap tmC;
  memset(&tmC, 0, sizeof(tmC));
  if (!a->out_f32) {
    // output view [n_pad columns][rows_out][sequences]; one store = 32 columns x 32 rows (one epilogue warp, 64 B rows)
    const long long out_rows = a->dyn_offsets != nullptr ? a->out_rows_total : a->rows_out;
    cuuint64_t dims[3] = {(cuuint64_t)a->n_pad, (cuuint64_t)out_rows, (cuuint64_t)a->a_seqs};
    cuuint64_t strides[2] = {(cuuint64_t)(a->out_row_stride * 2), (cuuint64_t)(a->out_seq_stride * 2)};
    if (a->a_seqs == 1) strides[1] = (cuuint64_t)(out_rows * a->out_row_stride * 2);
    cuuint32_t box[3] = {32, 32, 1};
    if (int rc = encode_map(&tmC, a->dtype, 3, a->out, dims, strides, box, "output", CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
  }

  // epilogue side input: the view [n_pad columns][side_rows][sequences] read in the output's 32 x 32 boxes
  CUtensorMap tmS;
  memset(&tmS, 0, sizeof(tmS));
  if (a->side_mode != 0) {
    cuuint64_t dims[3] = {(cuuint64_t)a->n_pad, (cuuint64_t)a->side_rows, (cuuint64_t)a->a_seqs};
    cuuint64_t strides[2] = {(cuuint64_t)(a->side_row_stride * 2), (cuuint64_t)(a->side_seq_stride * 2)};
    if (a->a_seqs == 1) strides[1] = (cuuint64_t)(a->side_rows * a->side_row_stride * 2);
    cuuint32_t box[3] = {32, 32, 1};
    if (int rc = encode_map(&tmS, a->dtype, 3, a->side, dims, strides, box, "side input", CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
  }

  vp3d::ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  p.a_seqs = (int)a->a_seqs;
  p.rows_out = (int)a->rows_out;
  p.m_tiles_per_seq = (int)((a->rows_out + 127) / 128);
  p.n_tiles = (int)(a->n_pad / a->block_n);
  p.taps = a->taps;
  p.kblocks_per_tap = (int)(a->k_per_tap / kblk);
  p.tap_row_step = a->tap_row_step;
  p.a_row_off = (int)a->a_row_off;
  p.b_tap_col_step = (int)a->w_tap_col_step;
  p.dyn = a->dyn_offsets;
  p.scale = a->scale;
  p.shift = a->shift;
  p.relu = a->relu;
  p.res = a->res;
  p.res_seq_stride = a->res_seq_stride;
  p.res_row_stride = a->res_row_stride;
  p.res_row_mul = a->res_row_mul;
  p.res_row_off = a->res_row_off;
  p.res_rows = a->res_rows;
  p.res_col_off = (int)a->res_col_off;
  p.res_cols = (int)a->res_cols;
  p.out = a->out;
  p.out_seq_stride = a->out_seq_stride;
  p.out_row_stride = a->out_row_stride;
  p.out_f32 = a->out_f32;
  p.out_tma = a->out_f32 ? 0 : 1;
  p.n_valid = (int)(a->out_f32 ? a->n_valid : a->n_pad);
  p.out_round_tf32 = a->out_round_tf32;
  p.stat_sum = a->stat_sum;
  p.stat_sqsum = a->stat_sqsum;
  if (a->fin != nullptr) {
    const vp3d_bn_fin* f = a->fin;
    if (a->stat_sum == nullptr) return fail(VP3D_ERR_INVALID, "fin needs stat_sum / stat_sqsum");
    if (!f->gamma || !f->beta || !f->scale || !f->shift || !f->mean || !f->invstd || !f->done_counter || f->c <= 0 ||
        f->c > a->n_pad || (f->running_mean == nullptr) != (f->running_var == nullptr))
      return fail(VP3D_ERR_INVALID, "fin: null pointer or bad channel count");
    if (f->count <= 1)
      return fail(VP3D_ERR_INVALID, "Expected more than 1 value per channel when training (got %lld)", f->count);
    const double n = (double)f->count;
    p.fin = vp3d::BnFinalizeParams{a->stat_sum, a->stat_sqsum, 1.0 / n, (float)(n / (n - 1.0)), f->gamma, f->beta, f->eps,
                                   f->momentum, f->running_mean, f->running_var, f->num_batches_tracked, f->scale,
                                   f->shift, f->mean, f->invstd, f->c};
    p.fin_counter = f->done_counter;
  }
  if (want_drop) p.drop = drop_of(a->drop);
  p.side_mode = a->side_mode;
  p.side_row_off = a->side_row_off;
  p.side_scale = a->side_scale;

  const long long total_tiles = (long long)p.a_seqs * p.m_tiles_per_seq * p.n_tiles;
  if (total_tiles > 0x7fffffffLL) return fail(VP3D_ERR_INVALID, "too many tiles");
  int grid = (int)(total_tiles < dev->sm_count ? total_tiles : dev->sm_count);
  // a grid that is a multiple of n_tiles keeps every CTA on one column tile (weights and BN statistics stay put)
  if (grid > p.n_tiles && grid % p.n_tiles != 0) grid -= grid % p.n_tiles;
  // CTA-pair kernel (cta_group::2) for the layers it covers (g_pair_mode: vp3d_set_pair_mode / VP3D_K1_2CTA)
  // (fused dropout / side input exist in the pair kernel only: such a launch takes it whatever the mode and tile count)
  const int use_pairs = dev->pair_refused ? 0 : (want_epi ? 2 : g_pair_mode);
  const bool pair_ok = vp3d::conv_gemm_pair_supported(a->dtype, a->block_n, a->w_mn_major, p);
  if (want_epi && !(use_pairs && pair_ok))
    return fail(VP3D_ERR_UNSUPPORTED, "fused dropout / side input need the CTA-pair kernel (16-bit operands and output, "
                "block_n 256, no dyn_offsets, cluster launches available on this device)");
  // (mode 1: from half a wave of tiles on. Measured on the training step, where the layers of 96 - 288 tiles moved from
  // the single-CTA kernel to pairs: 1.766 -> 1.752 ms -- a pair reads each weight tile once for two row tiles, and
  // launches of this size are bound by L2 -> shared-memory traffic, not by the tensor pipe)
  if (use_pairs && pair_ok && (use_pairs == 2 || 2 * total_tiles >= (long long)dev->sm_count)) {
    p.dyn_sched = (a->stat_sum == nullptr &&
                   (g_sched_mode == 2 || (g_sched_mode == 1 && total_tiles >= 2LL * dev->sm_count))) ? 1 : 0;
    CUtensorMap tmBh = tmB;    // MN-major: the same [64 k-rows][64 columns] boxes, two per CTA
    if (!a->w_mn_major) {
      cuuint64_t dims[2] = {(cuuint64_t)a->k_total, (cuuint64_t)a->n_pad};
      cuuint64_t strides[1] = {(cuuint64_t)(a->k_total * eb)};
      cuuint32_t box[2] = {(cuuint32_t)kblk, (cuuint32_t)(a->block_n / 2)};
      if (int rc = encode_map(&tmBh, a->dtype, 2, a->w, dims, strides, box, "weights (half tile)")) return rc;
    }
    cudaError_t e2 = vp3d::launch_conv_gemm_pair(a->dtype, a->w_mn_major, tmA, tmBh, tmC, tmS, p, dev->sm_count,
                                                 static_cast<cudaStream_t>(stream));
    if (e2 == cudaSuccess) return VP3D_OK;
    if (want_epi) return cuda_fail(e2, "conv_gemm_pair launch (fused epilogue)");
    // a cluster launch can be refused where the plain one is not (MPS / partitioned devices): not an error of the call --
    // the single-CTA kernel below covers every case. The refusal is remembered for THIS device only; other devices of
    // the process and the global mode (vp3d_set_pair_mode) are left alone.
    (void)cudaGetLastError();
    dev->pair_refused = true;
    fprintf(stderr, "vp3d_b200: CTA-pair launch refused on this device (%s); it uses the single-CTA kernel from now on\n",
            cudaGetErrorString(e2));
  }
  cudaError_t e = vp3d::launch_conv_gemm(a->dtype, a->block_n, a->w_mn_major, tmA, tmB, tmC, p, grid, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "conv_gemm launch");
  return VP3D_OK;
}

int vp3d_pack_rows(int dtype, const float* src, void* dst, long long rows, int c, int c_pad, void* stream) {
  if (src == nullptr || dst == nullptr || rows < 0 || c <= 0 || c_pad < c) return fail(VP3D_ERR_INVALID, "pack_rows args");
  if (rows == 0) return VP3D_OK;
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_pack_rows(dtype, src, dst, rows, c, c_pad, dev->sm_count, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "pack_rows launch");
  return VP3D_OK;
}

int vp3d_pack_rows_ones(int dtype, const float* src, void* dst, long long rows, int c, int c_pad, int ones_col,
                        void* stream) {
  if (src == nullptr || dst == nullptr || rows < 0 || c <= 0 || c_pad < c) return fail(VP3D_ERR_INVALID, "pack_rows_ones args");
  if (dtype != VP3D_F16 && dtype != VP3D_BF16) return fail(VP3D_ERR_INVALID, "pack_rows_ones: dtype must be F16 or BF16");
  if (ones_col < c || ones_col >= c_pad) return fail(VP3D_ERR_INVALID, "pack_rows_ones: ones_col must be a padding column");
  if (c_pad % 8 != 0 || (reinterpret_cast<uintptr_t>(dst) & 15)) return fail(VP3D_ERR_INVALID, "pack_rows_ones: c_pad % 8, 16-byte dst");
  if (rows == 0) return VP3D_OK;
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_pack_rows(dtype, src, dst, rows, c, c_pad, dev->sm_count, static_cast<cudaStream_t>(stream),
                                         ones_col);
  if (e != cudaSuccess) return cuda_fail(e, "pack_rows_ones launch");
  return VP3D_OK;
}

int vp3d_expand_bn_stats(int dtype, const float* gram, const void* w, int k_total, int ones_col, const float* gamma,
                         const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                         long long* num_batches_tracked, float* scale, float* shift, float* mean, float* invstd,
                         float* wg, int c, int c_pad, void* stream) {
  if (dtype != VP3D_F16 && dtype != VP3D_BF16) return fail(VP3D_ERR_INVALID, "expand_bn_stats: dtype must be F16 or BF16");
  if (!gram || !w || !gamma || !beta || !scale || !shift || !mean || !invstd || !wg)
    return fail(VP3D_ERR_INVALID, "expand_bn_stats: null pointer");
  if (k_total <= 0 || k_total > 256 || ones_col < 0 || ones_col >= k_total || c <= 0 || c_pad < c || c_pad % 8 != 0)
    return fail(VP3D_ERR_INVALID, "expand_bn_stats: need 0 < k_total <= 256, ones_col < k_total, c <= c_pad, c_pad % 8 == 0");
  if ((running_mean == nullptr) != (running_var == nullptr)) return fail(VP3D_ERR_INVALID, "expand_bn_stats running stats");
  cudaError_t e = vp3d::launch_expand_bn_stats(dtype, gram, w, k_total, ones_col, gamma, beta, eps, momentum, running_mean,
                                               running_var, num_batches_tracked, scale, shift, mean, invstd, wg, c, c_pad,
                                               static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "expand_bn_stats launch");
  return VP3D_OK;
}

int vp3d_expand_bwd_finish(int dtype, const float* p_packed, const float* wg, const float* gram, const void* w, int k_total,
                           int ones_col, const float* scale, const float* mean, const float* invstd,
                           const float* gscale_buf, int c, int c_pad, int c_in, int c_in_pad, int taps, float* dw,
                           float* d_gamma, float* d_beta, void* stream) {
  if (dtype != VP3D_F16 && dtype != VP3D_BF16) return fail(VP3D_ERR_INVALID, "expand_bwd_finish: dtype must be F16 or BF16");
  if (!p_packed || !wg || !gram || !w || !scale || !mean || !invstd || !dw || !d_gamma || !d_beta)
    return fail(VP3D_ERR_INVALID, "expand_bwd_finish: null pointer");
  if (k_total <= 0 || k_total > 256 || ones_col < 0 || ones_col >= k_total || c <= 0 || c_pad < c || c_pad % 8 != 0 ||
      c_in <= 0 || c_in > c_in_pad || taps <= 0 || taps * c_in_pad != k_total)
    return fail(VP3D_ERR_INVALID, "expand_bwd_finish: inconsistent sizes");
  cudaError_t e = vp3d::launch_expand_bwd_finish(dtype, p_packed, wg, gram, w, k_total, ones_col, scale, mean, invstd,
                                                 gscale_buf, c, c_pad, c_in, c_in_pad, taps, dw, d_gamma, d_beta,
                                                 static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "expand_bwd_finish launch");
  return VP3D_OK;
}

int vp3d_pack_conv_weight(int dtype, const float* w, void* dst, int c_out, int c_in, int taps, int rows_pad,
                          int k_pad_per_tap, int transpose, void* stream) {
  if (w == nullptr || dst == nullptr || c_out <= 0 || c_in <= 0 || taps <= 0)
    return fail(VP3D_ERR_INVALID, "pack_conv_weight args");
  if (transpose < 0 || transpose > 2) return fail(VP3D_ERR_INVALID, "pack_conv_weight transpose must be 0, 1 or 2");
  if (transpose == 0 && (rows_pad < c_out || k_pad_per_tap < c_in)) return fail(VP3D_ERR_INVALID, "pack_conv_weight padding");
  if (transpose == 1 && (rows_pad % taps != 0 || rows_pad / taps < c_in || k_pad_per_tap < c_out))
    return fail(VP3D_ERR_INVALID, "pack_conv_weight (transpose 1) padding");
  if (transpose == 2 && (rows_pad < c_in || k_pad_per_tap < c_out))
    return fail(VP3D_ERR_INVALID, "pack_conv_weight (transpose 2) padding");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_pack_weight(dtype, w, dst, c_out, c_in, taps, rows_pad, k_pad_per_tap, transpose,
                                           dev->sm_count, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "pack_weight launch");
  return VP3D_OK;
}

int vp3d_pack_conv_weight_scaled(int dtype, const float* w, const float* row_scale, void* dst, int c_out, int c_in,
                                 int taps, int rows_pad, int k_pad_per_tap, void* stream) {
  if (w == nullptr || dst == nullptr || row_scale == nullptr || c_out <= 0 || c_in <= 0 || taps <= 0)
    return fail(VP3D_ERR_INVALID, "pack_conv_weight_scaled args");
  if (rows_pad < c_out || k_pad_per_tap < c_in) return fail(VP3D_ERR_INVALID, "pack_conv_weight_scaled padding");
  if ((long long)rows_pad * k_pad_per_tap >= (1LL << 31)) return fail(VP3D_ERR_INVALID, "pack_conv_weight_scaled: too large");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_pack_weight(dtype, w, dst, c_out, c_in, taps, rows_pad, k_pad_per_tap, 0, dev->sm_count,
                                           static_cast<cudaStream_t>(stream), row_scale);
  if (e != cudaSuccess) return cuda_fail(e, "pack_weight (scaled) launch");
  return VP3D_OK;
}

int vp3d_bn_fold(const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps,
                 float* scale, float* shift, int c, int c_pad, void* stream) {
  if (!gamma || !beta || !running_mean || !running_var || !scale || !shift || c <= 0 || c_pad < c)
    return fail(VP3D_ERR_INVALID, "bn_fold args");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_bn_fold(gamma, beta, running_mean, running_var, eps, scale, shift, c, c_pad,
                                       static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "bn_fold launch");
  return VP3D_OK;
}

int vp3d_project_points(const float* x, float* out3, float* out2, long long n_pts, const float* q, const float* t,
                        const float* cam, long long pts_per_q, long long pts_per_cam, int mode, void* stream) {
  if (n_pts < 0) return fail(VP3D_ERR_INVALID, "n_pts < 0");
  if (n_pts == 0) return VP3D_OK;
  if (x == nullptr) return fail(VP3D_ERR_INVALID, "x is NULL");
  const int xf = mode & (VP3D_PT_WORLD_TO_CAMERA | VP3D_PT_CAMERA_TO_WORLD | VP3D_PT_ROTATE);
  if (xf != 0 && (xf & (xf - 1)) != 0) return fail(VP3D_ERR_INVALID, "choose one rigid transform");
  if (xf && q == nullptr) return fail(VP3D_ERR_INVALID, "q is NULL");
  if ((mode & (VP3D_PT_WORLD_TO_CAMERA | VP3D_PT_CAMERA_TO_WORLD)) && t == nullptr)
    return fail(VP3D_ERR_INVALID, "t is NULL");
  if (xf && pts_per_q <= 0) return fail(VP3D_ERR_INVALID, "pts_per_q <= 0");
  if (mode & VP3D_PT_PROJECT) {
    if (cam == nullptr || out2 == nullptr) return fail(VP3D_ERR_INVALID, "projection needs cam and out2");
    if (pts_per_cam <= 0) return fail(VP3D_ERR_INVALID, "pts_per_cam <= 0");
  } else if (out2 != nullptr) {
    return fail(VP3D_ERR_INVALID, "out2 given without VP3D_PT_PROJECT");
  }
  if (out3 == nullptr && out2 == nullptr) return fail(VP3D_ERR_INVALID, "no output");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_project_points(x, out3, out2, n_pts, q, t, cam, pts_per_q, pts_per_cam, mode,
                                              dev->sm_count, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "project_points launch");
  return VP3D_OK;
}

int vp3d_project_bwd(const float* x, const float* cam, const float* grad_out2, long long n_pts, long long pts_per_cam,
                     int linear, float* grad_x, void* stream) {
  if (n_pts < 0) return fail(VP3D_ERR_INVALID, "n_pts < 0");
  if (n_pts == 0) return VP3D_OK;
  if (!x || !cam || !grad_out2 || !grad_x || pts_per_cam <= 0) return fail(VP3D_ERR_INVALID, "project_bwd args");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_project_bwd(x, cam, grad_out2, n_pts, pts_per_cam, linear, grad_x, dev->sm_count,
                                           static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "project_bwd launch");
  return VP3D_OK;
}

int vp3d_project_windows(const vp3d_window_args* a, void* stream) {
  if (a == nullptr) return fail(VP3D_ERR_INVALID, "args is NULL");
  if (!a->x_world || !a->q || !a->t || !a->cam || !a->seq_start || !a->seq_len || !a->sample_seq || !a->sample_start ||
      !a->out2)
    return fail(VP3D_ERR_INVALID, "project_windows: null pointer");
  if (a->batch <= 0 || a->joints <= 0 || a->chunk_length <= 0 || a->pad < 0)
    return fail(VP3D_ERR_INVALID, "project_windows: bad sizes");
  if ((long long)a->batch * (a->chunk_length + 2 * a->pad) * a->joints >= (1LL << 32))
    return fail(VP3D_ERR_INVALID, "project_windows: batch too large for 32-bit point indices");
  if (reinterpret_cast<uintptr_t>(a->q) & 15) return fail(VP3D_ERR_INVALID, "project_windows: q must be 16-byte aligned");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_project_windows(a->x_world, a->q, a->t, a->cam, a->seq_start, a->seq_len, a->sample_seq,
                                               a->sample_start, a->batch, a->joints, a->chunk_length, a->pad,
                                               a->causal_shift, a->root_relative, a->linear, a->out2, a->target3,
                                               a->cam3x4, dev->sm_count, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "project_windows launch");
  return VP3D_OK;
}

long long vp3d_loss_workspace_bytes(void) { return 8LL * 8 * 1024; }  // up to 8192 per-CTA partial sums (double)

int vp3d_mpjpe_fwd(const float* pred, const float* target, long long n_joints, const float* w, long long T, long long J,
                   long long w_stride_n, long long w_stride_t, long long w_stride_j, void* workspace, float* out,
                   void* stream) {
  if (!pred || !target || !workspace || !out || n_joints <= 0) return fail(VP3D_ERR_INVALID, "mpjpe_fwd args");
  if (w != nullptr && (T <= 0 || J <= 0 || n_joints % (T * J) != 0)) return fail(VP3D_ERR_INVALID, "mpjpe_fwd weight grid");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  if (dev->sm_count * 8 > 8192) return fail(VP3D_ERR_UNSUPPORTED, "workspace too small for this device");
  cudaError_t e = vp3d::launch_mpjpe_fwd(pred, target, n_joints, w, T, J, w_stride_n, w_stride_t, w_stride_j,
                                         static_cast<double*>(workspace), out, dev->sm_count,
                                         static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "mpjpe_fwd launch");
  return VP3D_OK;
}

int vp3d_mpjpe_bwd(const float* pred, const float* target, const float* grad_out, long long n_joints, const float* w,
                   long long T, long long J, long long w_stride_n, long long w_stride_t, long long w_stride_j,
                   float* grad_pred, void* stream) {
  if (!pred || !target || !grad_out || !grad_pred || n_joints <= 0) return fail(VP3D_ERR_INVALID, "mpjpe_bwd args");
  if (w != nullptr && (T <= 0 || J <= 0 || n_joints % (T * J) != 0)) return fail(VP3D_ERR_INVALID, "mpjpe_bwd weight grid");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_mpjpe_bwd(pred, target, grad_out, n_joints, w, T, J, w_stride_n, w_stride_t, w_stride_j,
                                         grad_pred, dev->sm_count, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "mpjpe_bwd launch");
  return VP3D_OK;
}

int vp3d_mpjpe_nd_fwd(const float* pred, const float* target, long long n_points, int dim, const float* w, long long T,
                      long long J, long long w_stride_n, long long w_stride_t, long long w_stride_j, void* workspace,
                      float* out, void* stream) {
  if (!pred || !target || !workspace || !out || n_points <= 0 || dim <= 0) return fail(VP3D_ERR_INVALID, "mpjpe_nd_fwd args");
  if (w != nullptr && (T <= 0 || J <= 0 || n_points % (T * J) != 0)) return fail(VP3D_ERR_INVALID, "mpjpe_nd_fwd weight grid");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_mpjpe_nd_fwd(pred, target, n_points, dim, w, T, J, w_stride_n, w_stride_t, w_stride_j,
                                            static_cast<double*>(workspace), out, dev->sm_count,
                                            static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "mpjpe_nd_fwd launch");
  return VP3D_OK;
}

int vp3d_mpjpe_nd_bwd(const float* pred, const float* target, const float* grad_out, long long n_points, int dim,
                      const float* w, long long T, long long J, long long w_stride_n, long long w_stride_t,
                      long long w_stride_j, float* grad_pred, void* stream) {
  if (!pred || !target || !grad_out || !grad_pred || n_points <= 0 || dim <= 0)
    return fail(VP3D_ERR_INVALID, "mpjpe_nd_bwd args");
  if (w != nullptr && (T <= 0 || J <= 0 || n_points % (T * J) != 0)) return fail(VP3D_ERR_INVALID, "mpjpe_nd_bwd weight grid");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_mpjpe_nd_bwd(pred, target, grad_out, n_points, dim, w, T, J, w_stride_n, w_stride_t,
                                            w_stride_j, grad_pred, dev->sm_count, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "mpjpe_nd_bwd launch");
  return VP3D_OK;
}

int vp3d_n_mpjpe_fwd(const float* pred, const float* target, long long n_poses, int J, void* workspace, float* out,
                     void* stream) {
  if (!pred || !target || !workspace || !out || n_poses <= 0 || J <= 0) return fail(VP3D_ERR_INVALID, "n_mpjpe_fwd args");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_n_mpjpe_fwd(pred, target, n_poses, J, static_cast<double*>(workspace), out, dev->sm_count,
                                           static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "n_mpjpe_fwd launch");
  return VP3D_OK;
}

int vp3d_p_mpjpe_fwd(const float* pred, const float* target, long long n_poses, int J, void* workspace, float* out,
                     void* stream) {
  if (!pred || !target || !workspace || !out || n_poses <= 0 || J <= 0) return fail(VP3D_ERR_INVALID, "p_mpjpe_fwd args");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_p_mpjpe_fwd(pred, target, n_poses, J, static_cast<double*>(workspace), out, dev->sm_count,
                                           static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "p_mpjpe_fwd launch");
  return VP3D_OK;
}

int vp3d_velocity_error(const float* pred, const float* target, long long T, long long inner, int dim, void* workspace,
                        float* out, void* stream) {
  if (!pred || !target || !workspace || !out || T < 2 || inner <= 0 || dim <= 0)
    return fail(VP3D_ERR_INVALID, "velocity_error args (needs at least two frames)");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_velocity_error(pred, target, T, inner, dim, static_cast<double*>(workspace), out,
                                              dev->sm_count, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "velocity_error launch");
  return VP3D_OK;
}

int vp3d_n_mpjpe_bwd(const float* pred, const float* target, const float* grad_out, long long n_poses, int J,
                     float* grad_pred, void* stream) {
  if (!pred || !target || !grad_out || !grad_pred || n_poses <= 0 || J <= 0) return fail(VP3D_ERR_INVALID, "n_mpjpe_bwd args");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_n_mpjpe_bwd(pred, target, grad_out, n_poses, J, grad_pred, dev->sm_count,
                                           static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "n_mpjpe_bwd launch");
  return VP3D_OK;
}

int vp3d_reproj_mpjpe_fwd(const float* pose, const float* traj, long long n_points, long long pts_per_traj,
                          const float* cam, long long pts_per_cam, int linear, const float* target2, void* workspace,
                          float* out, void* stream) {
  if (!pose || !cam || !target2 || !workspace || !out || n_points <= 0 || pts_per_cam <= 0)
    return fail(VP3D_ERR_INVALID, "reproj_mpjpe_fwd args");
  if (traj != nullptr && (pts_per_traj <= 0 || n_points % pts_per_traj != 0))
    return fail(VP3D_ERR_INVALID, "reproj_mpjpe_fwd: n_points must be a multiple of pts_per_traj");
  if (reinterpret_cast<uintptr_t>(target2) & 7) return fail(VP3D_ERR_INVALID, "reproj_mpjpe_fwd: target2 must be 8-byte aligned");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_reproj_fwd(pose, traj, n_points, pts_per_traj, cam, pts_per_cam, linear, target2,
                                          static_cast<double*>(workspace), out, dev->sm_count,
                                          static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "reproj_mpjpe_fwd launch");
  return VP3D_OK;
}

int vp3d_reproj_mpjpe_bwd(const float* pose, const float* traj, long long n_points, long long pts_per_traj,
                          const float* cam, long long pts_per_cam, int linear, const float* target2,
                          const float* grad_out, float* grad_pose, float* grad_traj, void* stream) {
  if (!pose || !cam || !target2 || !grad_out || n_points <= 0 || pts_per_cam <= 0 || (!grad_pose && !grad_traj))
    return fail(VP3D_ERR_INVALID, "reproj_mpjpe_bwd args");
  if (traj != nullptr && (pts_per_traj <= 0 || n_points % pts_per_traj != 0))
    return fail(VP3D_ERR_INVALID, "reproj_mpjpe_bwd: n_points must be a multiple of pts_per_traj");
  if (grad_traj != nullptr && traj == nullptr) return fail(VP3D_ERR_INVALID, "reproj_mpjpe_bwd: grad_traj without traj");
  if (reinterpret_cast<uintptr_t>(target2) & 7) return fail(VP3D_ERR_INVALID, "reproj_mpjpe_bwd: target2 must be 8-byte aligned");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_reproj_bwd(pose, traj, n_points, pts_per_traj, cam, pts_per_cam, linear, target2, grad_out,
                                          grad_pose, grad_traj, dev->sm_count, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "reproj_mpjpe_bwd launch");
  return VP3D_OK;
}

int vp3d_wgrad(const vp3d_wgrad_args* a, void* stream) {
  if (a == nullptr) return fail(VP3D_ERR_INVALID, "args is NULL");
  if (a->dtype != VP3D_F16 && a->dtype != VP3D_BF16) return fail(VP3D_ERR_INVALID, "wgrad: dtype must be F16 or BF16");
  if (a->block_n != 256 && a->block_n != 64) return fail(VP3D_ERR_INVALID, "wgrad: block_n must be 256 or 64");
  if (!a->dz || !a->a || !a->dw_packed) return fail(VP3D_ERR_INVALID, "wgrad: null dz / a / dw_packed");
  if (a->dz_seqs <= 0 || a->dz_rows <= 0 || a->a_rows <= 0 || a->taps < 1) return fail(VP3D_ERR_INVALID, "wgrad: empty problem");
  if (a->co_pad <= 0 || a->co_pad % 128 != 0) return fail(VP3D_ERR_INVALID, "wgrad: co_pad must be a multiple of 128");
  if (a->ci_pad <= 0 || a->ci_pad % a->block_n != 0) return fail(VP3D_ERR_INVALID, "wgrad: ci_pad must be a multiple of block_n");
  if ((reinterpret_cast<uintptr_t>(a->dz) & 15) || (reinterpret_cast<uintptr_t>(a->a) & 15) ||
      (reinterpret_cast<uintptr_t>(a->dw_packed) & 15))
    return fail(VP3D_ERR_INVALID, "wgrad: pointers must be 16-byte aligned");
  if ((a->dz_row_stride * 2) % 16 || (a->dz_seq_stride * 2) % 16 || (a->a_row_stride * 2) % 16 || (a->a_seq_stride * 2) % 16)
    return fail(VP3D_ERR_INVALID, "wgrad: strides must be multiples of 16 bytes");
  /* a_cols may be smaller than ci_pad (+ tap offsets): TMA zero-fills the missing input columns */
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;

  CUtensorMap tmA, tmB;
  {
    /* dz_cols < co_pad: the missing output-channel columns read as zero (TMA), like a_cols on the other operand */
    const long long dz_cols = a->dz_cols > 0 ? a->dz_cols : a->co_pad;
    if (dz_cols > a->co_pad) return fail(VP3D_ERR_INVALID, "wgrad: dz_cols > co_pad");
    cuuint64_t dims[3] = {(cuuint64_t)dz_cols, (cuuint64_t)a->dz_rows, (cuuint64_t)a->dz_seqs};
    cuuint64_t strides[2] = {(cuuint64_t)(a->dz_row_stride * 2), (cuuint64_t)(a->dz_seq_stride * 2)};
    if (a->dz_seqs == 1) strides[1] = (cuuint64_t)(a->dz_rows * a->dz_row_stride * 2);
    cuuint32_t box[3] = {64, 64, 1};
    if (int rc = encode_map(&tmA, a->dtype, 3, a->dz, dims, strides, box, "wgrad dz")) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)a->a_cols, (cuuint64_t)a->a_rows, (cuuint64_t)a->dz_seqs};
    cuuint64_t strides[2] = {(cuuint64_t)(a->a_row_stride * 2), (cuuint64_t)(a->a_seq_stride * 2)};
    if (a->dz_seqs == 1) strides[1] = (cuuint64_t)(a->a_rows * a->a_row_stride * 2);
    cuuint32_t box[3] = {64, 64, 1};
    if (int rc = encode_map(&tmB, a->dtype, 3, a->a, dims, strides, box, "wgrad a")) return rc;
  }
  vp3d::WgradParams p;
  memset(&p, 0, sizeof(p));
  // wide layers: 256 x 256 tiles (fewer bytes per flop, deeper latency cover); narrow ones keep 128-row tiles
  const int block_m = (a->block_n == 256 && a->co_pad % 256 == 0 && a->dz_seqs * ((a->dz_rows + 63) / 64) >= 32) ? 256 : 128;
  p.co_tiles = (int)(a->co_pad / block_m);
  p.ci_tiles = (int)(a->ci_pad / a->block_n);
  p.num_tiles = a->taps * p.co_tiles * p.ci_tiles;
  p.seqs = (int)a->dz_seqs;
  p.kb_per_seq = (int)((a->dz_rows + 63) / 64);
  {
    // Number of row slices S: items = tiles * S are dealt round-robin to one CTA per SM. Cost model in units of one
    // 64-row block of MMAs: waves(S) * (rows blocks per slice + flush), flush ~ 8 blocks (one 128 x BN fp32 reduction).
    const long long kb_all = (long long)p.seqs * p.kb_per_seq;
    const long long flush = 8;
    long long best = -1;
    int best_s = 1;
    const int s_max = a->max_slices > 0 ? a->max_slices : 64;
    // dynamic schedule (vp3d_set_sched_mode, data-parallel training): items are handed out one by one to whichever CTA
    // is free, so a launch degrades gracefully when some SMs are busy with something else -- provided there are a few
    // items per worker. Every extra slice costs a flush of the tile (not overlapped with the MMAs for the 256-row tile;
    // measured ~40 row blocks' worth), so: the SMALLEST slice count that gives two items per SM, never slices shallower
    // than 32 row blocks, the cap when neither can be met.
    p.dyn_sched = g_sched_mode != 0 ? 1 : 0;
    if (p.dyn_sched) {
      best_s = 1;
      for (int s_try = 1; s_try <= s_max && s_try <= kb_all; ++s_try) {
        best_s = s_try;
        const long long depth = (kb_all + s_try - 1) / s_try;
        if ((long long)p.num_tiles * s_try >= 2LL * dev->sm_count || depth <= 32) break;
      }
    }
    for (int s_try = 1; !p.dyn_sched && s_try <= s_max && s_try <= kb_all; ++s_try) {
      const long long items = (long long)p.num_tiles * s_try;
      const long long waves = (items + dev->sm_count - 1) / dev->sm_count;
      const long long cost = waves * ((kb_all + s_try - 1) / s_try + flush);
      if (best < 0 || cost < best) {
        best = cost;
        best_s = s_try;
      }
    }
    p.num_slices = best_s;
  }
  p.valid_co = (int)(a->dz_cols > 0 ? a->dz_cols : a->co_pad);
  p.valid_ci = (int)(a->a_cols < a->ci_pad && a->b_tap_col_step == 0 ? a->a_cols : a->ci_pad);
  p.b_row_off = (int)a->b_row_off;
  p.b_tap_row_step = a->b_tap_row_step;
  p.b_tap_col_step = (int)a->b_tap_col_step;
  p.out = a->dw_packed;
  p.out_tap_stride = a->co_pad * a->ci_pad;
  p.out_row_stride = a->ci_pad;
  const long long items = (long long)p.num_tiles * p.num_slices;
  const int grid = (int)((items < dev->sm_count || p.dyn_sched) ? items : dev->sm_count);
  cudaError_t e = vp3d::launch_wgrad(a->dtype, a->block_n, block_m, tmA, tmB, p, grid,
                                     static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "wgrad launch");
  return VP3D_OK;
}

int vp3d_wgrad_finish(const float* dw_packed, float* dw, int c_out, int c_in, int taps, long long tap_stride,
                      long long row_stride, const float* gscale_buf, void* stream) {
  if (!dw_packed || !dw || c_out <= 0 || c_in <= 0 || taps <= 0 || row_stride < c_in || tap_stride < 0)
    return fail(VP3D_ERR_INVALID, "wgrad_finish args");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_wgrad_finish(dw_packed, dw, c_out, c_in, taps, tap_stride, row_stride, gscale_buf,
                                            dev->sm_count, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "wgrad_finish launch");
  return VP3D_OK;
}

int vp3d_bn_finalize(const double* stat_sum, const double* stat_sqsum, long long count, const float* gamma,
                     const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                     long long* num_batches_tracked, float* scale, float* shift, float* mean, float* invstd, int c,
                     int c_pad, void* stream) {
  if (!stat_sum || !stat_sqsum || !gamma || !beta || !scale || !shift || !mean || !invstd || c <= 0 || c_pad < c)
    return fail(VP3D_ERR_INVALID, "bn_finalize args");
  if ((running_mean == nullptr) != (running_var == nullptr)) return fail(VP3D_ERR_INVALID, "bn_finalize running stats");
  if (count <= 1)
    return fail(VP3D_ERR_INVALID, "Expected more than 1 value per channel when training (got %lld)", count);
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_bn_finalize(stat_sum, stat_sqsum, count, gamma, beta, eps, momentum, running_mean,
                                           running_var, num_batches_tracked, scale, shift, mean, invstd, c, c_pad,
                                           static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "bn_finalize launch");
  return VP3D_OK;
}

namespace {
int check_ew(int dtype, int c_pad, const char* what) {
  if (dtype != VP3D_F16 && dtype != VP3D_BF16) return fail(VP3D_ERR_INVALID, "%s: dtype must be F16 or BF16", what);
  if (c_pad <= 0 || c_pad % 8 != 0) return fail(VP3D_ERR_INVALID, "%s: c_pad must be a positive multiple of 8", what);
  return VP3D_OK;
}
}  // namespace

int vp3d_bn_act_fwd(int dtype, const void* z, const float* scale, const float* shift, const void* res,
                    long long seqs, long long rows_per_seq, long long res_seq_rows, int res_row_mul, int res_row_off,
                    int c_pad, const vp3d_dropout* drop, void* a, void* stream) {
  if (int rc = check_ew(dtype, c_pad, "bn_act_fwd")) return rc;
  if (!z || !scale || !shift || !a || seqs <= 0 || rows_per_seq <= 0) return fail(VP3D_ERR_INVALID, "bn_act_fwd args");
  if (drop && (drop->p < 0.f || drop->p >= 1.f)) return fail(VP3D_ERR_INVALID, "dropout p must be in [0, 1)");
  if (res != nullptr &&
      (res_row_off < 0 || (rows_per_seq - 1) * (long long)res_row_mul + res_row_off >= res_seq_rows))
    return fail(VP3D_ERR_INVALID, "bn_act_fwd: residual rows out of range");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  vp3d::BnFinalizeParams fin;
  memset(&fin, 0, sizeof(fin));
  cudaError_t e = vp3d::launch_bn_act_fwd(dtype, z, scale, shift, res, seqs, rows_per_seq, res_seq_rows, res_row_mul,
                                          res_row_off, c_pad, drop_of(drop), a, fin, dev->sm_count,
                                          static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "bn_act_fwd launch");
  return VP3D_OK;
}

int vp3d_bn_act_fwd_mask(int dtype, const void* z, const float* scale, const float* shift, const void* res,
                         long long seqs, long long rows_per_seq, long long res_seq_rows, int res_row_mul, int res_row_off,
                         int c_pad, const vp3d_dropout* drop, void* a, unsigned char* keep_mask, void* stream) {
  if (int rc = check_ew(dtype, c_pad, "bn_act_fwd_mask")) return rc;
  if (!z || !scale || !shift || !a || !keep_mask || seqs <= 0 || rows_per_seq <= 0)
    return fail(VP3D_ERR_INVALID, "bn_act_fwd_mask args");
  if (drop && (drop->p < 0.f || drop->p >= 1.f)) return fail(VP3D_ERR_INVALID, "dropout p must be in [0, 1)");
  if (res != nullptr &&
      (res_row_off < 0 || (rows_per_seq - 1) * (long long)res_row_mul + res_row_off >= res_seq_rows))
    return fail(VP3D_ERR_INVALID, "bn_act_fwd_mask: residual rows out of range");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  vp3d::BnFinalizeParams fin;
  memset(&fin, 0, sizeof(fin));
  cudaError_t e = vp3d::launch_bn_act_fwd(dtype, z, scale, shift, res, seqs, rows_per_seq, res_seq_rows, res_row_mul,
                                          res_row_off, c_pad, drop_of(drop), a, fin, dev->sm_count,
                                          static_cast<cudaStream_t>(stream), keep_mask);
  if (e != cudaSuccess) return cuda_fail(e, "bn_act_fwd_mask launch");
  return VP3D_OK;
}

int vp3d_bn_act_bwd_reduce_mask(int dtype, const void* g, const void* z, const unsigned char* keep_mask, const float* mean,
                                const float* invstd, float keep_scale, long long rows, int c_pad, double* sum_dy,
                                double* sum_dy_xhat, void* stream) {
  if (int rc = check_ew(dtype, c_pad, "bn_act_bwd_reduce_mask")) return rc;
  if (!g || !z || !keep_mask || !mean || !invstd || !sum_dy || !sum_dy_xhat || rows <= 0 || !(keep_scale >= 1.f))
    return fail(VP3D_ERR_INVALID, "bn_act_bwd_reduce_mask args");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_bn_act_bwd_reduce_mask(dtype, g, z, keep_mask, mean, invstd, keep_scale, rows, c_pad, sum_dy,
                                                      sum_dy_xhat, dev->sm_count, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "bn_act_bwd_reduce_mask launch");
  return VP3D_OK;
}

int vp3d_bn_act_bwd_apply_mask(int dtype, const void* g, const void* z, const unsigned char* keep_mask, const float* scale,
                               const float* mean, const float* invstd, float keep_scale, long long rows, long long count,
                               int c, int c_pad, const double* sum_dy, const double* sum_dy_xhat, const float* gscale_buf,
                               void* dz, float* d_gamma, float* d_beta, void* stream) {
  if (int rc = check_ew(dtype, c_pad, "bn_act_bwd_apply_mask")) return rc;
  if (!g || !z || !keep_mask || !scale || !mean || !invstd || !sum_dy || !sum_dy_xhat || !dz || rows <= 0 || c <= 0 ||
      c > c_pad || count < rows || !(keep_scale >= 1.f))
    return fail(VP3D_ERR_INVALID, "bn_act_bwd_apply_mask args");
  if ((d_gamma == nullptr) != (d_beta == nullptr)) return fail(VP3D_ERR_INVALID, "bn_act_bwd_apply_mask d_gamma / d_beta");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_bn_act_bwd_apply_mask(dtype, g, z, keep_mask, scale, mean, invstd, keep_scale, rows, count, c,
                                                     c_pad, sum_dy, sum_dy_xhat, gscale_buf, dz, d_gamma, d_beta,
                                                     dev->sm_count, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "bn_act_bwd_apply_mask launch");
  return VP3D_OK;
}

int vp3d_bn_finalize_act_fwd(int dtype, const void* z, const double* stat_sum, const double* stat_sqsum, long long count,
                             const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                             float* running_var, long long* num_batches_tracked, float* scale, float* shift, float* mean,
                             float* invstd, int c, const void* res, long long seqs, long long rows_per_seq,
                             long long res_seq_rows, int res_row_mul, int res_row_off, int c_pad,
                             const vp3d_dropout* drop, void* a, void* stream) {
  if (int rc = check_ew(dtype, c_pad, "bn_finalize_act_fwd")) return rc;
  if (!z || !a || seqs <= 0 || rows_per_seq <= 0) return fail(VP3D_ERR_INVALID, "bn_finalize_act_fwd args");
  if (!stat_sum || !stat_sqsum || !gamma || !beta || !scale || !shift || !mean || !invstd || count <= 0 || c <= 0 ||
      c > c_pad || (running_mean == nullptr) != (running_var == nullptr))
    return fail(VP3D_ERR_INVALID, "bn_finalize_act_fwd: statistics arguments");
  if (count <= 1)
    return fail(VP3D_ERR_INVALID, "Expected more than 1 value per channel when training (got %lld)", count);
  if (drop && (drop->p < 0.f || drop->p >= 1.f)) return fail(VP3D_ERR_INVALID, "dropout p must be in [0, 1)");
  if (res != nullptr &&
      (res_row_off < 0 || (rows_per_seq - 1) * (long long)res_row_mul + res_row_off >= res_seq_rows))
    return fail(VP3D_ERR_INVALID, "bn_finalize_act_fwd: residual rows out of range");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  const double n = (double)count;
  vp3d::BnFinalizeParams fin{stat_sum, stat_sqsum, 1.0 / n, (float)(n / (n - 1.0)), gamma, beta, eps, momentum,
                             running_mean, running_var, num_batches_tracked, scale, shift, mean, invstd, c};
  cudaError_t e = vp3d::launch_bn_act_fwd(dtype, z, nullptr, nullptr, res, seqs, rows_per_seq, res_seq_rows, res_row_mul,
                                          res_row_off, c_pad, drop_of(drop), a, fin, dev->sm_count,
                                          static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "bn_finalize_act_fwd launch");
  return VP3D_OK;
}

int vp3d_col_stats(int dtype, const void* z, long long rows, int c_pad, double* sum, double* sqsum, void* stream) {
  if (int rc = check_ew(dtype, c_pad, "col_stats")) return rc;
  if (!z || !sum || !sqsum || rows <= 0) return fail(VP3D_ERR_INVALID, "col_stats args");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_col_stats(dtype, z, rows, c_pad, sum, sqsum, dev->sm_count,
                                         static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "col_stats launch");
  return VP3D_OK;
}

int vp3d_bn_act_bwd_reduce(int dtype, const void* g, const void* z, const float* scale, const float* shift,
                           const float* mean, const float* invstd, long long rows, int c_pad,
                           const vp3d_dropout* drop, double* sum_dy, double* sum_dy_xhat, void* stream) {
  if (int rc = check_ew(dtype, c_pad, "bn_act_bwd_reduce")) return rc;
  if (!g || !z || !scale || !shift || !mean || !invstd || !sum_dy || !sum_dy_xhat || rows <= 0)
    return fail(VP3D_ERR_INVALID, "bn_act_bwd_reduce args");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_bn_act_bwd_reduce(dtype, g, z, scale, shift, mean, invstd, rows, c_pad, drop_of(drop),
                                                 sum_dy, sum_dy_xhat, dev->sm_count, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "bn_act_bwd_reduce launch");
  return VP3D_OK;
}

int vp3d_bn_act_bwd_apply(int dtype, const void* g, const void* z, const float* scale, const float* shift,
                          const float* mean, const float* invstd, long long rows, long long count, int c, int c_pad,
                          const vp3d_dropout* drop, const double* sum_dy, const double* sum_dy_xhat,
                          const float* gscale_buf, void* dz, float* d_gamma, float* d_beta, void* stream) {
  if (int rc = check_ew(dtype, c_pad, "bn_act_bwd_apply")) return rc;
  if (!g || !z || !scale || !shift || !mean || !invstd || !sum_dy || !sum_dy_xhat || !dz || rows <= 0 || c <= 0 ||
      c > c_pad || count < rows)
    return fail(VP3D_ERR_INVALID, "bn_act_bwd_apply args");
  if ((d_gamma == nullptr) != (d_beta == nullptr)) return fail(VP3D_ERR_INVALID, "bn_act_bwd_apply d_gamma / d_beta");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_bn_act_bwd_apply(dtype, g, z, scale, shift, mean, invstd, rows, count, c, c_pad, drop_of(drop),
                                                sum_dy, sum_dy_xhat, gscale_buf, dz, d_gamma, d_beta, dev->sm_count,
                                                static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "bn_act_bwd_apply launch");
  return VP3D_OK;
}

int vp3d_counter_add(unsigned long long* counter, unsigned long long inc, void* stream) {
  if (!counter) return fail(VP3D_ERR_INVALID, "counter_add: null counter");
  cudaError_t e = vp3d::launch_counter_add(counter, inc, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "counter_add launch");
  return VP3D_OK;
}

int vp3d_grad_scale(const float* dy, long long n, float* gscale_buf, void* stream) {
  if (!dy || !gscale_buf || n <= 0) return fail(VP3D_ERR_INVALID, "grad_scale args");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_grad_scale(dy, n, gscale_buf, dev->sm_count, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "grad_scale launch");
  return VP3D_OK;
}

int vp3d_grad_pack_rows(int dtype, const float* src, void* dst, long long rows, int c, int c_pad,
                        const float* gscale_buf, float* col_sum, void* stream) {
  if (dtype != VP3D_F16 && dtype != VP3D_BF16) return fail(VP3D_ERR_INVALID, "grad_pack_rows: dtype must be F16 or BF16");
  if (!src || !dst || rows <= 0 || c <= 0 || c_pad < c) return fail(VP3D_ERR_INVALID, "grad_pack_rows args");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_grad_pack_rows(dtype, src, dst, rows, c, c_pad, gscale_buf, col_sum, dev->sm_count,
                                              static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "grad_pack_rows launch");
  return VP3D_OK;
}

int vp3d_stream_advance(long long* step, int n_rings, const int* ring_len, const int* ring_dil, const int* ring_taps,
                        int rows_per_slot, int* table, int n_launch, const int* launch_desc, int* launch_table,
                        void* stream) {
  if (!step || !ring_len || !ring_dil || !ring_taps || !table || n_rings <= 0 || n_rings > 64 || rows_per_slot <= 0)
    return fail(VP3D_ERR_INVALID, "stream_advance args");
  if (n_launch < 0 || n_launch > 64 || (n_launch > 0 && (!launch_desc || !launch_table)))
    return fail(VP3D_ERR_INVALID, "stream_advance launch table");
  cudaError_t e = vp3d::launch_stream_advance(step, n_rings, ring_len, ring_dil, ring_taps, rows_per_slot, table,
                                              n_launch, launch_desc, launch_table,
                                              static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "stream_advance launch");
  return VP3D_OK;
}

int vp3d_ring_write(int dtype, const float* src, void* ring, const int* table_entry, long long rows, int c, int c_pad,
                    void* stream) {
  if (dtype != VP3D_F16 && dtype != VP3D_BF16) return fail(VP3D_ERR_INVALID, "ring_write: dtype must be F16 or BF16");
  if (!src || !ring || !table_entry || rows <= 0 || c <= 0 || c_pad < c) return fail(VP3D_ERR_INVALID, "ring_write args");
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_ring_write(dtype, src, ring, table_entry, rows, c, c_pad, dev->sm_count,
                                          static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "ring_write launch");
  return VP3D_OK;
}

int vp3d_stream_step_fused(int dtype, long long* step, unsigned long long* barrier_counter, int n_rings, const int* ring_len,
                           const int* ring_dil, const int* ring_taps, int rows_per_slot, const float* x_in, int c_in,
                           int c_in_pad, void* ring0, int n_streams, const vp3d_stream_layer* layers, int n_layers,
                           void* stream) {
  if (dtype != VP3D_F16 && dtype != VP3D_BF16) return fail(VP3D_ERR_INVALID, "stream_step_fused: dtype must be F16 or BF16");
  if (!step || !barrier_counter || !ring_len || !ring_dil || !ring_taps || !x_in || !ring0 || !layers)
    return fail(VP3D_ERR_INVALID, "stream_step_fused: null pointer");
  if (n_rings <= 0 || n_rings > 16 || rows_per_slot <= 0 || c_in <= 0 || c_in_pad < c_in)
    return fail(VP3D_ERR_INVALID, "stream_step_fused: ring / input geometry");
  if (n_streams <= 0 || n_streams > vp3d::stream_step_max_streams())
    return fail(VP3D_ERR_UNSUPPORTED, "stream_step_fused serves 1..%d streams (got %d); larger batches take the GEMM path",
                vp3d::stream_step_max_streams(), n_streams);
  if (n_layers <= 0 || n_layers > vp3d::kStreamMaxLayers) return fail(VP3D_ERR_INVALID, "stream_step_fused: 1..12 layers");
  vp3d::StreamStepParams p;
  memset(&p, 0, sizeof(p));
  p.step = step;
  p.barrier = barrier_counter;
  p.ring_len = ring_len; p.ring_dil = ring_dil; p.ring_taps = ring_taps;
  p.n_rings = n_rings; p.rows_per_slot = rows_per_slot;
  p.x_in = x_in; p.ring0 = ring0; p.c_in = c_in; p.c_in_pad = c_in_pad; p.n_streams = n_streams; p.n_layers = n_layers;
  for (int l = 0; l < n_layers; ++l) {
    const vp3d_stream_layer& s = layers[l];
    if (!s.a || !s.w || !s.out || s.taps <= 0 || s.k_per_tap <= 0 || s.k_per_tap % 8 != 0 ||
        (long long)s.taps * s.k_per_tap > vp3d::stream_step_max_k() || s.n <= 0 || s.a_ring >= n_rings ||
        s.res_ring >= n_rings || s.out_ring >= n_rings || (s.res != nullptr && s.res_ring < 0))
      return fail(VP3D_ERR_INVALID, "stream_step_fused: layer %d", l);
    vp3d::StreamLayer& d = p.layers[l];
    d.a = s.a; d.w = s.w; d.shift = s.shift; d.res = s.res; d.out = s.out;
    d.a_ring = s.a_ring; d.res_ring = s.res_ring; d.out_ring = s.out_ring;
    d.k_per_tap = s.k_per_tap; d.taps = s.taps; d.tap_row_step = s.tap_row_step;
    d.n = s.n; d.n_valid = s.n_valid; d.relu = s.relu; d.out_f32 = s.out_f32;
    d.res_row_stride = s.res_row_stride; d.out_row_stride = s.out_row_stride;
  }
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  cudaError_t e = vp3d::launch_stream_step(dtype, p, dev->sm_count, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "stream_step_fused launch");
  return VP3D_OK;
}

static int check_adam_args(const vp3d_adam_args* a) {
  if (!a->p || !a->g || !a->m || !a->v || !a->step || a->n <= 0)
    return fail(VP3D_ERR_INVALID, "adam_step: null pointer or empty tensor");
  if (a->packed != nullptr && a->n % 4 != 0) return fail(VP3D_ERR_INVALID, "adam_step: packed weights need n % 4 == 0");
  if ((reinterpret_cast<uintptr_t>(a->p) | reinterpret_cast<uintptr_t>(a->g) | reinterpret_cast<uintptr_t>(a->m) |
       reinterpret_cast<uintptr_t>(a->v) | reinterpret_cast<uintptr_t>(a->vmax)) & 15)
    return fail(VP3D_ERR_INVALID, "adam_step: tensors must be 16-byte aligned");
  if (a->packed != nullptr) {
    if (a->dtype != VP3D_F16 && a->dtype != VP3D_BF16) return fail(VP3D_ERR_INVALID, "adam_step: packed dtype");
    if (a->c_in <= 0 || a->taps <= 0 || a->k_pad < a->c_in || a->n % ((long long)a->c_in * a->taps) != 0)
      return fail(VP3D_ERR_INVALID, "adam_step: weight geometry");
  }
  return VP3D_OK;
}

int vp3d_adam_step(const vp3d_adam_args* a, void* stream) {
  if (a == nullptr) return fail(VP3D_ERR_INVALID, "args is NULL");
  if (int rc = check_adam_args(a)) return rc;
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  vp3d::AdamParams q;
  q.p = a->p; q.g = a->g; q.m = a->m; q.v = a->v; q.vmax = a->vmax;
  q.n = a->n;
  q.lr = a->lr; q.beta1 = a->beta1; q.beta2 = a->beta2; q.eps = a->eps; q.weight_decay = a->weight_decay;
  q.step = a->step; q.lr_dev = a->lr_dev;
  q.maximize = a->maximize;
  q.packed = a->packed; q.c_in = a->c_in; q.taps = a->taps; q.k_pad = a->k_pad;
  cudaError_t e = vp3d::launch_adam_pack(a->dtype, q, dev->sm_count, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "adam_step launch");
  return VP3D_OK;
}

int vp3d_adam_step_multi(const vp3d_adam_args* args, int count, void* stream) {
  if (count < 0 || (count > 0 && args == nullptr)) return fail(VP3D_ERR_INVALID, "adam_step_multi: bad arguments");
  if (count == 0) return VP3D_OK;
  const vp3d_adam_args& h = args[0];
  int dtype = -1;
  for (int i = 0; i < count; ++i) {
    const vp3d_adam_args& a = args[i];
    if (int rc = check_adam_args(&a)) return rc;
    if (a.lr != h.lr || a.beta1 != h.beta1 || a.beta2 != h.beta2 || a.eps != h.eps || a.weight_decay != h.weight_decay ||
        a.lr_dev != h.lr_dev || a.maximize != h.maximize || (a.vmax == nullptr) != (h.vmax == nullptr))
      return fail(VP3D_ERR_INVALID, "adam_step_multi: tensors of one call share their hyper-parameters");
    if (a.packed != nullptr) {
      if (dtype >= 0 && a.dtype != dtype) return fail(VP3D_ERR_INVALID, "adam_step_multi: one packed dtype per call");
      dtype = a.dtype;
    }
  }
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  const long long budget = (long long)dev->sm_count * 8;   // blocks of 256 threads in flight, dealt by tensor size
  for (int first = 0; first < count; first += vp3d::kAdamMaxTensors) {
    const int n_t = count - first < vp3d::kAdamMaxTensors ? count - first : vp3d::kAdamMaxTensors;
    long long total = 0;
    for (int i = 0; i < n_t; ++i) total += args[first + i].n;
    vp3d::AdamMultiParams mp;
    memset(&mp, 0, sizeof(mp));
    mp.count = n_t;
    mp.lr = h.lr; mp.beta1 = h.beta1; mp.beta2 = h.beta2; mp.eps = h.eps; mp.weight_decay = h.weight_decay;
    mp.lr_dev = h.lr_dev;
    mp.maximize = h.maximize;
    int blocks = 0;
    for (int i = 0; i < n_t; ++i) {
      const vp3d_adam_args& a = args[first + i];
      vp3d::AdamTensor& T = mp.t[i];
      T.p = a.p; T.g = a.g; T.m = a.m; T.v = a.v; T.vmax = a.vmax;
      T.n = a.n;
      T.step = a.step;
      T.packed = a.packed; T.c_in = a.c_in; T.taps = a.taps; T.k_pad = a.k_pad;
      const long long need = ((a.n + 3) / 4 + 255) / 256;                 // blocks that give every thread one float4
      long long share = (budget * a.n + total - 1) / total;               // proportional share of the grid
      if (share > need) share = need;
      if (share < 1) share = 1;
      mp.block_start[i] = blocks;
      blocks += (int)share;
    }
    mp.block_start[n_t] = blocks;
    cudaError_t e = vp3d::launch_adam_multi(dtype >= 0 ? dtype : VP3D_F16, mp, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "adam_step_multi launch");
  }
  return VP3D_OK;
}

int vp3d_peer_allreduce_f32(const vp3d_allreduce_args* a, void* stream) {
  if (a == nullptr || a->peers == nullptr || a->flags == nullptr) return fail(VP3D_ERR_INVALID, "peer_allreduce: null arguments");
  if (a->world < 1 || a->world > vp3d::kArMaxRanks || a->rank < 0 || a->rank >= a->world)
    return fail(VP3D_ERR_INVALID, "peer_allreduce: rank %d of %d (at most %d ranks)", a->rank, a->world, vp3d::kArMaxRanks);
  if (a->ctas < 2 || a->ctas > vp3d::kArMaxCtas || a->ctas % 2 != 0)
    return fail(VP3D_ERR_INVALID, "peer_allreduce: ctas must be even, 2..%d", vp3d::kArMaxCtas);
  if (a->offset < 0 || a->count < 0 || a->offset % 4 != 0 || a->count % 4 != 0)
    return fail(VP3D_ERR_INVALID, "peer_allreduce: offset / count must be non-negative multiples of 4 floats");
  if (a->count == 0) return VP3D_OK;
  DeviceInfo* dev = nullptr;
  if (int rc = device_info(&dev)) return rc;
  vp3d::AllReduceParams p;
  memset(&p, 0, sizeof(p));
  p.mc = static_cast<float*>(a->multicast);
  for (int k = 0; k < a->world; ++k) {
    if (a->peers[k] == nullptr || a->flags[k] == nullptr) return fail(VP3D_ERR_INVALID, "peer_allreduce: null mapping of rank %d", k);
    if ((reinterpret_cast<uintptr_t>(a->peers[k]) & 15) != 0) return fail(VP3D_ERR_INVALID, "peer_allreduce: buffers must be 16-byte aligned");
    p.peers[k] = static_cast<float*>(a->peers[k]);
    p.flags[k] = static_cast<uint32_t*>(a->flags[k]);
  }
  if ((reinterpret_cast<uintptr_t>(a->multicast) & 15) != 0) return fail(VP3D_ERR_INVALID, "peer_allreduce: multicast address must be 16-byte aligned");
  p.rank = a->rank;
  p.world = a->world;
  p.off = a->offset;
  p.n = a->count;
  p.scale = a->scale;
  const double t = a->timeout_s > 0 ? a->timeout_s : 10.0;
  p.timeout_ns = static_cast<unsigned long long>(t * 1e9);
  cudaError_t e = vp3d::launch_peer_allreduce(p, a->ctas, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "peer_allreduce launch");
  return VP3D_OK;
}


}  // extern "C"
