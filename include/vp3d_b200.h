/* vp3d_b200.h -- C ABI of the B200 (sm_100a) hot path of Dynamic-Camera-Augmented-VideoPose3D.
 *
 * The reference has no FFI layer: its boundary for this path is the Python module API that run.py star-imports
 * (run.py:21-26). This header is the C boundary underneath our drop-in Python modules; each entry point names the
 * reference code it replaces. Conventions:
 *   - every function returns 0 (VP3D_OK) or a vp3d_status; it never throws and never allocates device memory;
 *     vp3d_last_error() returns a thread-local description of the last failure.
 *   - all pointers are raw device pointers owned by the caller unless stated; `stream` is a cudaStream_t passed
 *     as void* (NULL = default stream). Calls are asynchronous on that stream and re-entrant per stream.
 *   - strides are in elements of the named buffer.
 */
#ifndef VP3D_B200_H
#define VP3D_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum vp3d_status {
  VP3D_OK = 0,
  VP3D_ERR_INVALID = 1,     /* bad argument (shape, alignment, null pointer) */
  VP3D_ERR_CUDA = 2,        /* a CUDA runtime / driver call failed */
  VP3D_ERR_UNSUPPORTED = 3  /* not an sm_100 device, or feature not available */
} vp3d_status;

/* Operand type of the tensor-core contraction (activations and packed weights use the same type).
 * F16/BF16: 2-byte storage, tcgen05 kind::f16. TF32: 4-byte fp32 storage, tcgen05 kind::tf32.
 * Accumulation is always fp32 in tensor memory. */
typedef enum vp3d_dtype { VP3D_F16 = 0, VP3D_BF16 = 1, VP3D_TF32 = 2 } vp3d_dtype;

int vp3d_version(void);
const char* vp3d_last_error(void);
/* SM count and compute capability of the current device; VP3D_ERR_UNSUPPORTED unless cc == 10.x. */
int vp3d_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------------------------
 * K1  temporal convolution block: Conv1d (+ folded BatchNorm1d + ReLU + residual slice-add) as one implicit GEMM.
 * Replaces nn.Conv1d / nn.BatchNorm1d(eval) / nn.ReLU / `res + x` of common/models/TemporalModel.py:126-138
 * (dilated TemporalModel) and :188-198 (strided TemporalModelOptimized1f); with `w` packed transposed it is also the
 * data-gradient GEMM of the same layers.
 *
 *   out[s][t][n] = epilogue( sum_{tap, c} a[s][t + a_row_off + tap * tap_row_step][c] * w[n][tap * k_per_tap + c] )
 *   epilogue(x)  = relu?( x * scale[n] + shift[n] ) + res[s][t * res_row_mul + res_row_off][n]
 *
 * Activations are channels-last. Rows of `a` outside [0, a_rows) read as zero. A strided convolution whose stride
 * equals its width is expressed as taps = 1 on the reshaped view [s][t_out][taps * c].
 */
typedef struct vp3d_conv_args {
  int dtype;               /* vp3d_dtype */
  int block_n;             /* output-channel tile: 256, or 64 for narrow layers; n_pad must be a multiple */

  const void* a;           /* activations, element type = dtype; 16-byte aligned */
  long long a_seqs;        /* sequences */
  long long a_rows;        /* rows (frames) per sequence in this view */
  long long a_kdim;        /* channels per row visible to the contraction (= k_per_tap for taps > 1) */
  long long a_row_stride;  /* multiple of 16 bytes */
  long long a_seq_stride;
  long long a_row_off;     /* input row read by tap 0 of output row 0 (may be negative) */

  const void* w;           /* packed weights [n_pad][k_total], K contiguous (vp3d_pack_conv_weight) */
  long long n_pad;
  long long k_total;       /* taps * k_per_tap */
  int taps;
  int tap_row_step;        /* dilation, in rows */
  long long k_per_tap;     /* multiple of 128 bytes of K */

  long long rows_out;      /* output rows per sequence */
  void* out;
  int out_f32;             /* 1: fp32 output with n_valid real columns; 0: element-typed, n_pad columns written */
  long long out_row_stride;
  long long out_seq_stride;
  long long n_valid;
  int out_round_tf32;      /* fp32 output rounded (nearest) to TF32 so a following TF32 layer reads exact operands */

  const float* scale;      /* [n_pad] or NULL */
  const float* shift;      /* [n_pad]; required when scale is given */
  int relu;
  const void* res;         /* residual source, element type = dtype (fp32 for TF32), or NULL */
  long long res_row_stride;
  long long res_seq_stride;
  int res_row_mul;
  int res_row_off;

  float* stat_sum;         /* optional [n_pad] accumulators (+=) of the raw output and its square over valid rows */
  float* stat_sqsum;
} vp3d_conv_args;

int vp3d_conv_block_fwd(const vp3d_conv_args* args, void* stream);

/* fp32 [rows][c] -> dtype [rows][c_pad], zero padded (the (N,T,J*F) model input, TemporalModel.py:67-68). */
int vp3d_pack_rows(int dtype, const float* src, void* dst, long long rows, int c, int c_pad, void* stream);

/* nn.Conv1d weight (c_out, c_in, taps) fp32 (TemporalModel.py:33,102,113-118) -> packed GEMM operand.
 * transpose == 0: dst[rows_pad][taps * k_pad_per_tap], dst[n][tap * k_pad_per_tap + ci] = w[n][ci][tap]
 * transpose == 1: dst[rows_pad][k_pad_per_tap],        dst[tap * c_in + ci][co]          = w[co][ci][tap]  */
int vp3d_pack_conv_weight(int dtype, const float* w, void* dst, int c_out, int c_in, int taps, int rows_pad,
                          int k_pad_per_tap, int transpose, void* stream);

/* Eval-mode nn.BatchNorm1d (TemporalModel.py:32,117,119) -> scale = gamma / sqrt(var + eps), shift = beta - mean *
 * scale; entries [c, c_pad) are zeroed. */
int vp3d_bn_fold(const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps,
                 float* scale, float* shift, int c, int c_pad, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K5  camera geometry on points X[n_pts][3] (fp32). Replaces common/quaternion.py:10-35 and common/camera.py:28-90.
 * `mode` is a bit set:
 *   VP3D_PT_WORLD_TO_CAMERA  X <- qrot(conj(q), X - t)        camera.py:28-30
 *   VP3D_PT_CAMERA_TO_WORLD  X <- qrot(q, X) + t              camera.py:33-34
 *   VP3D_PT_ROTATE           X <- qrot(q, X)                  quaternion.py:10-24 (| VP3D_PT_CONJ: conj(q))
 *   VP3D_PT_PROJECT          out2 <- project_to_2d(X, cam)    camera.py:37-67   (| VP3D_PT_LINEAR: camera.py:69-90)
 * q[.][4] = (w,x,y,z) and t[.][3] are shared by pts_per_q consecutive points (J for one camera pose per frame,
 * T*J for a static camera); cam[.][9] = (fx,fy,cx,cy,k1,k2,k3,p1,p2) by pts_per_cam consecutive points.
 * out3 (transformed 3-D points) and out2 (2-D projections) may each be NULL.
 */
enum {
  VP3D_PT_WORLD_TO_CAMERA = 1,
  VP3D_PT_CAMERA_TO_WORLD = 2,
  VP3D_PT_ROTATE = 4,
  VP3D_PT_CONJ = 8,
  VP3D_PT_PROJECT = 16,
  VP3D_PT_LINEAR = 32
};
int vp3d_project_points(const float* x, float* out3, float* out2, long long n_pts, const float* q, const float* t,
                        const float* cam, long long pts_per_q, long long pts_per_cam, int mode, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K6  losses on joint grids pred/target[n_joints][3] fp32 (contiguous). Replaces common/loss.py:11-27,70-80.
 * `workspace` must hold vp3d_loss_workspace_bytes() bytes. `out` / `grad_out` are single device floats.
 * Weights (may be NULL): w is indexed over the (n, t, j) grid with n_joints = N*T*J through element strides
 * (0 = broadcast), matching `w * norm` broadcasting in loss.py:27.
 */
long long vp3d_loss_workspace_bytes(void);
int vp3d_mpjpe_fwd(const float* pred, const float* target, long long n_joints, const float* w, long long T, long long J,
                   long long w_stride_n, long long w_stride_t, long long w_stride_j, void* workspace, float* out,
                   void* stream);
int vp3d_mpjpe_bwd(const float* pred, const float* target, const float* grad_out, long long n_joints, const float* w,
                   long long T, long long J, long long w_stride_n, long long w_stride_t, long long w_stride_j,
                   float* grad_pred, void* stream);
int vp3d_n_mpjpe_fwd(const float* pred, const float* target, long long n_poses, int J, void* workspace, float* out,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VP3D_B200_H */
