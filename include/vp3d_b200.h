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
/* Grid-size bound for the persistent kernels: they size their grids for at most `sms` SMs (<= 0: all of them; the
 * environment variable VP3D_SM_LIMIT sets the initial value). Used by data-parallel training to leave a few SMs to
 * NCCL's all-reduce kernels (vp3d_b200.ddp.enable_grad_sync(reserve_sms=...)); no reference counterpart. */
int vp3d_set_sm_limit(int sms);
/* Tile schedule of the CTA-pair kernel (launches without statistics of at least two waves of tiles): 0 = static
 * persistent schedule, one cluster per SM pair walking tiles i, i + n, ... (default); 1 = dynamic: the grid has one
 * cluster per tile and running clusters steal the tiles of clusters that have not been launched yet (cluster launch
 * control, clusterlaunchcontrol.try_cancel). The dynamic schedule adapts to SMs that are busy with another kernel --
 * NCCL's all-reduce CTAs during a data-parallel backward (vp3d_b200.ddp switches it on). Initial value from the
 * environment variable VP3D_SCHED ("dynamic"). Mode 2 ("dynamic-all") applies it to every pair-kernel launch without
 * statistics whatever its size (tests). Results are identical in all modes. */
int vp3d_set_sched_mode(int mode);
/* Programmatic dependent launch: 1 (default; environment variable VP3D_PDL=0 turns it off) launches every kernel of the
 * library with cudaLaunchAttributeProgrammaticStreamSerialization, so the next kernel on the stream becomes resident and
 * runs its prologue while this one drains (each kernel orders its own global-memory accesses with griddepcontrol.wait);
 * 0 = plain stream order. Results are identical. No reference counterpart. */
int vp3d_set_pdl(int on);
/* Which K1 kernel vp3d_conv_block_fwd launches: 0 = always the single-CTA kernel, 1 = the CTA-pair kernel
 * (tcgen05.mma.cta_group::2) for supported launches of at least half a wave of tiles (default), 2 = for every supported
 * launch. Initial value from the environment variable VP3D_K1_2CTA ("0", "force"). */
int vp3d_set_pair_mode(int mode);

/* Dropout description shared by forward and backward: keep-mask = Philox4x32-10(seed, stream, row, channel group)
 * >= p; kept values are multiplied by 1 / (1 - p) (nn.Dropout, TemporalModel.py:28,127,134-135). p == 0: off. */
typedef struct vp3d_dropout {
  float p;
  unsigned long long seed;
  unsigned long long stream;              /* distinguishes layers */
  const unsigned long long* step_counter; /* optional device counter read by the kernel and mixed into the Philox
                                             counter: a captured CUDA graph draws a fresh mask on every replay */
} vp3d_dropout;

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
/* Arguments of vp3d_bn_finalize for the in-GEMM finalize (vp3d_conv_args.fin): same meaning, plus a counter. */
typedef struct vp3d_bn_fin {
  long long count;          /* rows the statistics are taken over (> 1) */
  const float* gamma;
  const float* beta;
  float eps;
  float momentum;           /* < 0: cumulative moving average (nn.BatchNorm1d(momentum=None)) */
  float* running_mean;      /* may be NULL (together with running_var) */
  float* running_var;
  long long* num_batches_tracked;   /* may be NULL */
  float* scale;             /* outputs, fp32 [n_pad] */
  float* shift;
  float* mean;
  float* invstd;
  int c;                    /* real channels (<= n_pad) */
  unsigned int* done_counter;   /* device word, ZERO on entry; the launch counts its finished CTAs in it */
} vp3d_bn_fin;

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
  int w_mn_major;          /* 1 (fp16 / bf16 only): `w` is [k_per_tap rows][w_row_stride columns] with the OUTPUT column
                              contiguous -- the forward-packed weights [c_out][taps * c_in_pad] read as W^T by the
                              data-gradient GEMM: out[.][n] += a[.][k] * w[k][tap * w_tap_col_step + n]. n_pad is then
                              the number of output columns and k_total is ignored. */
  long long w_row_stride;
  long long w_tap_col_step;

  long long rows_out;      /* output rows per sequence */
  void* out;
  int out_f32;             /* 1: fp32 output with n_valid real columns; 0: element-typed, n_pad columns written */
  long long out_row_stride;
  long long out_seq_stride;
  long long n_valid;
  int out_round_tf32;      /* fp32 output rounded (nearest) to TF32 so a following TF32 layer reads exact operands */

  const float* scale;      /* [n_pad] or NULL (NULL with shift given: epilogue(x) = relu?(x + shift[n]) -- the scale is
                              folded into the packed weight rows, see vp3d_pack_conv_weight_scaled) */
  const float* shift;      /* [n_pad]; required when scale is given */
  int relu;
  const void* res;         /* residual source, element type = dtype (fp32 for TF32), or NULL */
  long long res_row_stride;
  long long res_seq_stride;
  int res_row_mul;         /* residual row = out_row * res_row_mul + res_row_off */
  int res_row_off;
  long long res_rows;      /* > 0: residual rows outside [0, res_rows) contribute nothing (data-gradient fan-in) */
  long long res_col_off;   /* the residual is added to output columns [res_col_off, res_col_off + res_cols) only, */
  long long res_cols;      /* reading residual column (col - res_col_off); res_cols == 0: every column, offset 0  */

  const int* dyn_offsets;  /* optional DEVICE int[4] {a_row_off, res_row_off, out_row_off, out_row_off2 or -1}: when given,
                              the first two replace the fields above, out rows are shifted by out_row_off and every
                              16-bit output box is stored a second time at out_row_off2 (mirror slot). `out_rows_total`
                              must then give the number of rows of the whole output matrix (stores are clipped to it,
                              not to rows_out). Lets a captured CUDA graph walk ring buffers (vp3d_stream_advance). */
  long long out_rows_total;

  double* stat_sum;        /* optional [n_pad] accumulators (+=) of the output AS STORED (16-bit) and its square over the */
  double* stat_sqsum;      /* valid rows: train-mode BatchNorm statistics; fp32 per CTA, double across CTAs */

  const struct vp3d_bn_fin* fin; /* optional (with stat_sum): vp3d_bn_finalize in the tail of this launch -- the last CTA to
                              finish turns the sums into scale / shift / mean / invstd and updates the running
                              statistics, so the stand-alone finalize launch (and its launch gap) disappears */

  /* Fused train-mode epilogue (CTA-pair kernel: 16-bit operands and output, block_n 256, no dyn_offsets; the launch
   * takes that kernel whatever vp3d_set_pair_mode says and fails with VP3D_ERR_UNSUPPORTED where it cannot run). */
  const vp3d_dropout* drop;/* optional: nn.Dropout after the ReLU (TemporalModel.py:127 / :189), the same counter-based mask as
                              vp3d_bn_act_fwd draws for (flat output row s * rows_out + t, channel) */
  const void* side;        /* optional side input [s][side_rows][n_pad] in the operand type, fetched by TMA in the output's
                              tiles (column n of output row t reads side row t + side_row_off; rows outside read 0) */
  int side_mode;           /* 0: none. 1: out += side (a residual whose rows map 1:1; the generic `res` also covers strided
                              rows and column windows). 2: out = side > 0 ? out * side_scale : 0 -- the ReLU + dropout mask
                              of a layer recovered from its stored activation (clipped and dropped elements are 0),
                              applied to the gradient that reaches that activation (data-gradient GEMM of the next
                              layer); with stat_sum the column sums of the gated gradient come out of the same pass */
  int side_row_off;
  long long side_row_stride;
  long long side_seq_stride;
  long long side_rows;
  float side_scale;
} vp3d_conv_args;

int vp3d_conv_block_fwd(const vp3d_conv_args* args, void* stream);

/* fp32 [rows][c] -> dtype [rows][c_pad], zero padded (the (N,T,J*F) model input, TemporalModel.py:67-68). */
int vp3d_pack_rows(int dtype, const float* src, void* dst, long long rows, int c, int c_pad, void* stream);

/* The same with 1.0 in padding column `ones_col` (c <= ones_col < c_pad; F16 / BF16, c_pad % 8 == 0): the Gram matrix
 * of the packed rows then carries their column sums and the row count (vp3d_expand_bn_stats). The convolution weights
 * of a padding column are zero, so the layer output does not change. */
int vp3d_pack_rows_ones(int dtype, const float* src, void* dst, long long rows, int c, int c_pad, int ones_col,
                        void* stream);

/* nn.Conv1d weight (c_out, c_in, taps) fp32 (TemporalModel.py:33,102,113-118) -> packed GEMM operand.
 * transpose == 0: dst[rows_pad][taps * k_pad_per_tap], dst[n][tap * k_pad_per_tap + ci] = w[n][ci][tap]
 * transpose == 1: dst[rows_pad][k_pad_per_tap],        dst[tap * c_in_pad + ci][co]      = w[co][ci][tap]
 *                 (c_in_pad = rows_pad / taps: data-gradient operand of a stride == width convolution)
 * transpose == 2: dst[rows_pad][taps * k_pad_per_tap], dst[ci][tap * k_pad_per_tap + co] = w[co][ci][tap]
 *                 (data-gradient operand of a dilated convolution, taps walked with a negative row step) */
int vp3d_pack_conv_weight(int dtype, const float* w, void* dst, int c_out, int c_in, int taps, int rows_pad,
                          int k_pad_per_tap, int transpose, void* stream);

/* transpose == 0 packing with every output-channel row multiplied by row_scale[n] before rounding to the operand type:
 * eval-mode BatchNorm's gamma / sqrt(var + eps) folded into the weights, so the GEMM epilogue only adds the shift. */
int vp3d_pack_conv_weight_scaled(int dtype, const float* w, const float* row_scale, void* dst, int c_out, int c_in,
                                 int taps, int rows_pad, int k_pad_per_tap, void* stream);

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
  VP3D_PT_LINEAR = 32,
  VP3D_PT_FAST = 64   /* contracted (FMA) arithmetic and one reciprocal per point instead of the reference's operation
                         by operation evaluation: <= 1e-6 relative difference, ~2x fewer instructions (the default of
                         the fused world_to_image / streaming paths; the drop-in functions stay un-contracted) */
};
int vp3d_project_points(const float* x, float* out3, float* out2, long long n_pts, const float* q, const float* t,
                        const float* cam, long long pts_per_q, long long pts_per_cam, int mode, void* stream);

/* Backward of VP3D_PT_PROJECT (| VP3D_PT_LINEAR) wrt the camera-space points: grad_x[n_pts][3] = d sum(grad_out2 *
 * project_to_2d(x, cam)) / dx -- what autograd derives for camera.py:54-67 / :85-90 (torch.clamp passes the gradient
 * inside [-1, 1], bounds included). */
int vp3d_project_bwd(const float* x, const float* cam, const float* grad_out2, long long n_pts, long long pts_per_cam,
                     int linear, float* grad_x, void* stream);

/* Window feeder: the batch assembly of common/generators.py:102-132 fused with the dynamic-camera
 * projection. All sequences are resident on the device, concatenated: x_world[frames][joints][3], one camera pose per
 * frame q[frames][4] / t[frames][3], intrinsics cam[n_seq][9]; seq_start / seq_len give each sequence's frame range.
 * Sample b = (sample_seq[b], sample_start[b]) is the reference's (seq_i, start_3d) pair; window frame k reads source
 * frame clamp(start_3d - pad - causal_shift + k, 0, len - 1) (np.pad 'edge', generators.py:92-100).
 *   out2    [batch][chunk_length + 2 pad][joints][2]  project_to_2d(world_to_camera(X, q, t), cam)   (camera.py:28-67)
 *   target3 [batch][chunk_length][joints][3] or NULL  camera-space joints of the chunk frames, root-relative if asked
 *                                                     (run.py:72-74)
 *   cam3x4  [batch][window][3][4] or NULL             K @ [R | -R c] per window frame (generators.py:113-125) */
typedef struct vp3d_window_args {
  const float* x_world; const float* q; const float* t; const float* cam;
  const long long* seq_start; const long long* seq_len;
  const int* sample_seq; const long long* sample_start;
  int batch, joints, chunk_length, pad, causal_shift, root_relative, linear;
  float* out2; float* target3; float* cam3x4;
} vp3d_window_args;
int vp3d_project_windows(const vp3d_window_args* args, void* stream);

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
/* The same two for points of any dimension `dim` (pred/target[n_points][dim]): loss.py:17 takes the norm over the last
 * axis whatever its length (2-D reprojection errors). dim == 3 callers should use the vectorised pair above. */
int vp3d_mpjpe_nd_fwd(const float* pred, const float* target, long long n_points, int dim, const float* w, long long T,
                      long long J, long long w_stride_n, long long w_stride_t, long long w_stride_j, void* workspace,
                      float* out, void* stream);
int vp3d_mpjpe_nd_bwd(const float* pred, const float* target, const float* grad_out, long long n_points, int dim,
                      const float* w, long long T, long long J, long long w_stride_n, long long w_stride_t,
                      long long w_stride_j, float* grad_pred, void* stream);
int vp3d_n_mpjpe_fwd(const float* pred, const float* target, long long n_poses, int J, void* workspace, float* out,
                     void* stream);
/* Evaluation metrics on the device (the reference computes them in NumPy after a D2H copy, run.py:749-756):
 *   vp3d_p_mpjpe_fwd       loss.py:29-68   MPJPE after the optimal similarity transform per pose (3x3 SVD Procrustes with
 *                                          the reflection fix), poses pred/target[n_poses][J][3]
 *   vp3d_velocity_error    loss.py:82-91   mean | diff_t(pred) - diff_t(target) |_2 over arrays [T][inner][dim]
 * `workspace` as for vp3d_mpjpe_fwd; `out` one device float. */
int vp3d_p_mpjpe_fwd(const float* pred, const float* target, long long n_poses, int J, void* workspace, float* out,
                     void* stream);
int vp3d_velocity_error(const float* pred, const float* target, long long T, long long inner, int dim, void* workspace,
                        float* out, void* stream);
/* grad_pred = grad_out * d n_mpjpe / d pred (the scale factor is differentiated too, as autograd does for loss.py:77-80). */
int vp3d_n_mpjpe_bwd(const float* pred, const float* target, const float* grad_out, long long n_poses, int J,
                     float* grad_pred, void* stream);

/* Fused reprojection loss (north-star item 3, "projection kernel with fused MPJPE"):
 *   out = mean_i || project_to_2d(pose[i] + traj[i / pts_per_traj], cam[i / pts_per_cam]) - target2[i] ||_2
 * i.e. common/loss.py:11-17 applied to common/camera.py:37-67 (linear != 0: :69-90) in one pass over the points, without
 * the 2-D projection ever reaching memory. `traj` may be NULL. The backward writes d out / d pose (grad_pose, may be
 * NULL) and accumulates d out / d traj into grad_traj (ZERO on entry, may be NULL) with the clamp / z -> 0 conventions
 * of vp3d_project_bwd. workspace: vp3d_loss_workspace_bytes(). */
int vp3d_reproj_mpjpe_fwd(const float* pose, const float* traj, long long n_points, long long pts_per_traj,
                          const float* cam, long long pts_per_cam, int linear, const float* target2, void* workspace,
                          float* out, void* stream);
int vp3d_reproj_mpjpe_bwd(const float* pose, const float* traj, long long n_points, long long pts_per_traj,
                          const float* cam, long long pts_per_cam, int linear, const float* target2,
                          const float* grad_out, float* grad_pose, float* grad_traj, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Training path (TemporalModel.py:126-138 / :188-198 in train() mode and their autograd backward, run.py:473-485).
 * Layout: every activation / gradient matrix is channels-last [rows][c] in the 16-bit operand type (fp16 / bf16),
 * rows = sequences * frames. Gradients travel multiplied by a power-of-two `gscale` chosen on the device from
 * max|dL/dy| (vp3d_grad_scale) so that fp16 keeps its precision; parameter gradients are un-scaled when they are
 * written as fp32. VP3D_TF32 is not available on the training path.
 */

/* K4  weight gradient of a convolution: dW[tap][co][ci] += sum_{s, r} dz[s][r][co] * a[s][r + b_row_off + tap *
 * b_tap_row_step][ci + tap * b_tap_col_step]  (the autograd weight gradient of nn.Conv1d, TemporalModel.py:102,
 * 113-118,168-181). Both operands are read MN-major (reduction over rows) straight from the channels-last matrices;
 * work is split over (tile, row-block) units evenly across the SMs and combined with fp32 reductions into
 * `dw_packed` [taps][co_pad][ci_pad] fp32, which the caller zero-fills. block_n (ci tile) is 256 or 64. */
typedef struct vp3d_wgrad_args {
  int dtype;               /* VP3D_F16 or VP3D_BF16 */
  int block_n;
  const void* dz;          /* [seqs][rows][co_pad] gradient wrt the convolution output */
  long long dz_seqs, dz_rows, dz_row_stride, dz_seq_stride;
  long long co_pad;        /* multiple of 128 */
  const void* a;           /* layer input, viewed [seqs][a_rows][a_cols] */
  long long a_rows, a_cols, a_row_stride, a_seq_stride;  /* columns past a_cols read as zero */
  long long ci_pad;        /* input channels per tap (multiple of block_n) */
  int taps;
  long long b_row_off;
  int b_tap_row_step;      /* dilated convolution: tap k reads input row r + k * dilation */
  long long b_tap_col_step;/* stride == width convolution on the reshaped view: tap k reads columns k * c_in_pad + ci */
  float* dw_packed;        /* [taps][co_pad][ci_pad] fp32, accumulated with red.add */
  long long dz_cols;       /* 0: co_pad. Otherwise the real column count of `dz` (<= co_pad); the rest reads as zero */
  int max_slices;          /* 0: 64. Upper bound on the row slices (split-K depth) the launch may choose */
} vp3d_wgrad_args;
int vp3d_wgrad(const vp3d_wgrad_args* args, void* stream);

/* dw[co][ci][tap] = dw_packed[tap * tap_stride + co * row_stride + ci] * gscale_buf[1]  (nn.Conv1d weight layout,
 * fp32; gscale_buf = {gscale, 1 / gscale} on the device, NULL = 1). For vp3d_wgrad's [taps][co_pad][ci_pad] result
 * tap_stride = co_pad * ci_pad, row_stride = ci_pad; a narrow layer whose taps were contracted as ONE 256-wide tile
 * (taps = 1, b_tap_col_step = 0 on the [rows][taps * c_in_pad] view) has tap_stride = c_in_pad, row_stride = 256. */
int vp3d_wgrad_finish(const float* dw_packed, float* dw, int c_out, int c_in, int taps, long long tap_stride,
                      long long row_stride, const float* gscale_buf, void* stream);

/* Train-mode BatchNorm of the EXPAND layer (TemporalModel.py:127 / :189: expand_bn(expand_conv(x))) from the Gram matrix
 * of the layer input instead of a pass over the layer output: the convolution is linear with a contraction of only
 * k_total = taps * c_in_pad <= 256 columns, so with X the [rows][k_total] view of the packed input (one padding column
 * `ones_col` holding 1.0), gram = X^T X (fp32 [256][256], e.g. vp3d_wgrad with both operands = X) and w the packed
 * weights [c_pad][k_total]:  mean_c = w_c . s / n,  var_c = w_c^T (G / n - s s^T / n^2) w_c  (s = gram[ones_col][:],
 * n = gram[ones_col][ones_col]). Writes scale / shift / mean / invstd like vp3d_bn_finalize (running statistics and
 * num_batches_tracked updated the same way; momentum < 0 = cumulative average) and wg[c_pad][256] = W * gram, which
 * vp3d_expand_bwd_finish reads again. */
int vp3d_expand_bn_stats(int dtype, const float* gram, const void* w, int k_total, int ones_col, const float* gamma,
                         const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                         long long* num_batches_tracked, float* scale, float* shift, float* mean, float* invstd,
                         float* wg, int c, int c_pad, void* stream);
/* Backward of the same layer from p_packed[c_pad][256] = gm^T X (vp3d_wgrad), gm = the gradient wrt the layer's
 * activation gated by its ReLU / dropout mask and scaled by gscale_buf[0] (vp3d_conv_block_fwd side_mode 2):
 *   Sg = p[c][ones_col],  Sgz = sum_k w[c][k] p[c][k],  d_beta = Sg,  d_gamma = invstd (Sgz - mean Sg),
 *   dw[c][ci][tap] = scale_c ( p[c][k] - Sg s_k / n - d_gamma invstd (wg[c][k] - mean s_k) / n ),  k = tap * c_in_pad + ci
 * (all multiplied by gscale_buf[1]) -- the sum over rows of dz x_k with dz the BatchNorm backward, without materialising
 * dz and without a reduction pass over the activation gradient. */
int vp3d_expand_bwd_finish(int dtype, const float* p_packed, const float* wg, const float* gram, const void* w, int k_total,
                           int ones_col, const float* scale, const float* mean, const float* invstd,
                           const float* gscale_buf, int c, int c_pad, int c_in, int c_in_pad, int taps, float* dw,
                           float* d_gamma, float* d_beta, void* stream);

/* Train-mode nn.BatchNorm1d statistics (TemporalModel.py:32,117,119 in train()): from the per-channel sum / sum of
 * squares over `count` rows (accumulated by vp3d_conv_block_fwd) produce the forward affine scale = gamma * invstd,
 * shift = beta - mean * scale (entries [c, c_pad) zero), save mean / invstd for the backward, and update
 * running_mean / running_var (unbiased variance, `momentum`) and num_batches_tracked (int64, += 1) in place. */
int vp3d_bn_finalize(const double* stat_sum, const double* stat_sqsum, long long count, const float* gamma,
                     const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                     long long* num_batches_tracked, float* scale, float* shift, float* mean, float* invstd, int c,
                     int c_pad, void* stream);

/* sum[c] += sum_rows z[r][c], sqsum[c] += sum_rows z[r][c]^2 over a stored [rows][c_pad] matrix: the same statistics
 * as the GEMM epilogue's stat_sum / stat_sqsum, for layers whose contraction is too short to hide that reduction. */
int vp3d_col_stats(int dtype, const void* z, long long rows, int c_pad, double* sum, double* sqsum, void* stream);

/* *counter += inc on the stream (one thread): the per-step tick of step_counter, capturable in a CUDA graph. */
int vp3d_counter_add(unsigned long long* counter, unsigned long long inc, void* stream);

/* a[s][t][c] = dropout(relu(z[s][t][c] * scale[c] + shift[c])) + res[s][t * res_row_mul + res_row_off][c]
 * (TemporalModel.py:127,134-135 / :189,194-195). z, a: [seqs * rows_per_seq][c_pad]; res: [seqs][res_seq_rows][c_pad]
 * or NULL. */
int vp3d_bn_act_fwd(int dtype, const void* z, const float* scale, const float* shift, const void* res,
                    long long seqs, long long rows_per_seq, long long res_seq_rows, int res_row_mul, int res_row_off,
                    int c_pad, const vp3d_dropout* drop, void* a, void* stream);
/* vp3d_bn_finalize and vp3d_bn_act_fwd in one launch (the training forward of every layer when the statistics need no
 * all-reduce in between): every thread derives scale / shift of its channels from the sums; scale / shift / mean /
 * invstd (fp32 [c_pad]) are written for the backward, running statistics and num_batches_tracked updated as by
 * vp3d_bn_finalize. Results are bit-identical to the two-call sequence. */
int vp3d_bn_finalize_act_fwd(int dtype, const void* z, const double* stat_sum, const double* stat_sqsum, long long count,
                             const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                             float* running_var, long long* num_batches_tracked, float* scale, float* shift, float* mean,
                             float* invstd, int c, const void* res, long long seqs, long long rows_per_seq,
                             long long res_seq_rows, int res_row_mul, int res_row_off, int c_pad,
                             const vp3d_dropout* drop, void* a, void* stream);

/* Backward of the same chain, phase 1: with dy = g * dropout-mask / (1 - p) * [z * scale + shift > 0] and
 * xhat = (z - mean) * invstd, accumulate sum_dy[c] += sum_rows dy, sum_dy_xhat[c] += sum_rows dy * xhat (double). */
int vp3d_bn_act_bwd_reduce(int dtype, const void* g, const void* z, const float* scale, const float* shift,
                           const float* mean, const float* invstd, long long rows, int c_pad,
                           const vp3d_dropout* drop, double* sum_dy, double* sum_dy_xhat, void* stream);
/* phase 2: dz = scale * (dy - sum_dy / count - xhat * sum_dy_xhat / count) in the operand type (count = rows the
 * statistics were taken over: `rows`, or the global row count under SyncBN after the caller all-reduced the sums);
 * also writes the BatchNorm parameter gradients d_gamma[c] = sum_dy_xhat * gscale_buf[1], d_beta[c] = sum_dy *
 * gscale_buf[1]. */
int vp3d_bn_act_bwd_apply(int dtype, const void* g, const void* z, const float* scale, const float* shift,
                          const float* mean, const float* invstd, long long rows, long long count, int c, int c_pad,
                          const vp3d_dropout* drop, const double* sum_dy, const double* sum_dy_xhat,
                          const float* gscale_buf, void* dz, float* d_gamma, float* d_beta, void* stream);

/* The same three passes with the keep decision STORED instead of recomputed. vp3d_bn_act_fwd_mask also writes
 * keep_mask[r][c_pad / 8] (one byte per row and 8-channel group, bit k = channel 8 group + k passed the ReLU and was not
 * dropped: 1/16 of the 16-bit activation). The backward passes then need neither the dropout stream nor the affine
 * comparison and run at memory speed (~7 instead of 22-26 instructions per element): with gm = keep ? g : 0,
 *   reduce: sum_dy[c] += keep_scale * sum_rows gm,  sum_dy_xhat[c] += keep_scale * invstd[c] * sum_rows gm * (z - mean[c])
 *   apply:  dz = scale * (keep_scale * gm - sum_dy / count - (z - mean) * invstd * sum_dy_xhat / count), d_gamma / d_beta as
 *           vp3d_bn_act_bwd_apply.
 * keep_scale = 1 / (1 - p) as quantised by the dropout (256 / (256 - round(256 p)); 1 without dropout). Same results as
 * the recomputing passes up to fp32 rounding of the folded constants. */
int vp3d_bn_act_fwd_mask(int dtype, const void* z, const float* scale, const float* shift, const void* res,
                         long long seqs, long long rows_per_seq, long long res_seq_rows, int res_row_mul, int res_row_off,
                         int c_pad, const vp3d_dropout* drop, void* a, unsigned char* keep_mask, void* stream);
int vp3d_bn_act_bwd_reduce_mask(int dtype, const void* g, const void* z, const unsigned char* keep_mask, const float* mean,
                                const float* invstd, float keep_scale, long long rows, int c_pad, double* sum_dy,
                                double* sum_dy_xhat, void* stream);
int vp3d_bn_act_bwd_apply_mask(int dtype, const void* g, const void* z, const unsigned char* keep_mask, const float* scale,
                               const float* mean, const float* invstd, float keep_scale, long long rows, long long count,
                               int c, int c_pad, const double* sum_dy, const double* sum_dy_xhat, const float* gscale_buf,
                               void* dz, float* d_gamma, float* d_beta, void* stream);

/* gscale_buf[0] = 2^floor(log2(64 / max|dy|)) (1 if dy == 0), gscale_buf[1] = 1 / gscale_buf[0]; gscale_buf[2] is
 * scratch (max|dy|). dy: n fp32 values. */
int vp3d_grad_scale(const float* dy, long long n, float* gscale_buf, void* stream);
/* dst[r][k] = (k < c ? src[r][k] * gscale_buf[0] : 0) in the operand type; col_sum[k] += sum_r src[r][k] (fp32,
 * unscaled: the bias gradient of the shrink layer, TemporalModel.py:33), col_sum may be NULL. */
int vp3d_grad_pack_rows(int dtype, const float* src, void* dst, long long rows, int c, int c_pad,
                        const float* gscale_buf, float* col_sum, void* stream);

/* Streaming ring bookkeeping on the device (causal inference, BASELINE configs[3]). Ring i holds the last L_i = (taps_i -
 * 1) * dil_i + 1 input frames of convolution i twice: frame t lives in slots t mod L_i and t mod L_i + L_i, each slot
 * `rows_per_slot` rows. vp3d_stream_advance increments *step and writes, for every ring, table[i] = {a_row_off of tap 0
 * for this frame, row of the current frame (residual / write position, upper copy), the same row again as out_row_off,
 * row of the mirror copy}, all in rows. A GEMM launch usually reads one ring and writes another, so the same kernel
 * also composes launch_table[l] = {table[a_ring][0], table[res_ring][1], table[out_ring][2], table[out_ring][3]} from
 * launch_desc[l] = {a_ring, res_ring, out_ring} (-1 = none -> 0, 0, 0, -1): the int[4] blocks that
 * vp3d_conv_args.dyn_offsets points at.
 * vp3d_ring_write casts the new fp32 input rows [rows][c] into both copies of ring 0's current slot. */
int vp3d_stream_advance(long long* step, int n_rings, const int* ring_len, const int* ring_dil, const int* ring_taps,
                        int rows_per_slot, int* table, int n_launch, const int* launch_desc, int* launch_table,
                        void* stream);
int vp3d_ring_write(int dtype, const float* src, void* ring, const int* table_entry, long long rows, int c, int c_pad,
                    void* stream);

/* Low-latency streaming step for a FEW concurrent streams (n_streams <= 8; BASELINE configs[3]): the whole frame -- ring
 * bookkeeping (what vp3d_stream_advance computes), the write of the new 2-D keypoints into ring 0 (vp3d_ring_write) and
 * every layer of the causal TemporalModel (TemporalModel.py:126-138 one frame at a time) -- as ONE cooperative kernel
 * with grid barriers between the layers: with so few rows a layer is a matrix-vector product whose only cost is streaming
 * the weights out of L2, so one warp owns one output channel and nothing is launched between layers. Layer l computes
 *   out[s][c] = act(sum_k w[c][k] * a[a_pos + tap * tap_row_step + s][k mod k_per_tap] + shift[c]) (+ res[res_pos + s][c])
 * for s < n_streams, c < n, with a_pos / res_pos / out positions taken from the rings a_ring / res_ring / out_ring (-1: a
 * plain buffer starting at row 0; 16-bit outputs into a ring are stored in its slot AND its mirror). Same operands and
 * rounding points as the vp3d_conv_block_fwd launches it replaces (16-bit activations between layers, fp32 accumulate).
 * `layers` is a HOST array (n_layers <= 12, taps * k_per_tap <= 3072, k_per_tap % 8 == 0). `barrier_counter` points at
 * 128 device words (1 KB), zero-filled once by the caller and owned by this call from then on (barrier counters and the
 * number of frames they have served; independent of *step, which vp3d_stream_advance may advance in between). */
typedef struct vp3d_stream_layer {
  const void* a; const void* w; const float* shift; const void* res; void* out;
  int a_ring, res_ring, out_ring;
  int k_per_tap, taps, tap_row_step;
  int n, n_valid, relu, out_f32;
  int res_row_stride, out_row_stride;
} vp3d_stream_layer;
int vp3d_stream_step_fused(int dtype, long long* step, unsigned long long* barrier_counter, int n_rings, const int* ring_len,
                           const int* ring_dil, const int* ring_taps, int rows_per_slot, const float* x_in, int c_in,
                           int c_in_pad, void* ring0, int n_streams, const vp3d_stream_layer* layers, int n_layers,
                           void* stream);

/* Fused optimiser step for one convolution weight: torch.optim.Adam(amsgrad) as run.py:662 uses it, in
 * the arithmetic of torch's capturable implementation, plus the re-pack of the updated fp32 weight (c_out, c_in, taps)
 * into the 16-bit K-major operand packed[c_out_pad][taps][k_pad] the next forward reads (padding entries are left
 * untouched: the caller zero-fills the buffer once). `step` is a device float holding the number of THIS update (>= 1),
 * `lr_dev` an optional device learning rate -- both so that a captured CUDA graph needs no re-capture when they change.
 * Works on any fp32 tensor (BatchNorm affine, biases) with packed == NULL; vmax == NULL disables amsgrad. */
typedef struct vp3d_adam_args {
  float* p; const float* g; float* m; float* v; float* vmax;
  long long n;
  float lr, beta1, beta2, eps, weight_decay;
  const float* step; const float* lr_dev;
  int maximize;
  void* packed; int dtype, c_in, taps, k_pad;
} vp3d_adam_args;
int vp3d_adam_step(const vp3d_adam_args* args, void* stream);
/* The same update for `count` tensors in one launch (the 30 parameter tensors of a TemporalModel: one kernel instead of
 * 30 back-to-back ones, whose launch gaps cost more than their work for the BatchNorm vectors). All entries must share
 * lr / betas / eps / weight_decay / lr_dev / maximize / amsgrad and, where given, the packed dtype; `step` stays per
 * tensor (torch keeps one step counter per parameter, run.py:436-445 restores them from checkpoints). */
int vp3d_adam_step_multi(const vp3d_adam_args* args, int count, void* stream);

/* Gradient exchange of data-parallel training through NVLink / NVSwitch peer memory (no reference counterpart: the
 * reference trains in one process, run.py:473-487). In-place sum x scale of the fp32 slice
 * [offset, offset + count) of an exchange buffer that every rank of the NVLink domain has mapped: a two-shot all-reduce
 * in ONE kernel of `ctas` CTAs (rank barrier -> every rank reduces its 1 / world share and writes it to all replicas ->
 * rank barrier). `multicast` != NULL: the share is reduced by the switch (multimem.ld_reduce) and broadcast by it
 * (multimem.st); NULL: plain peer loads / stores through `peers`. One rank computes each element, every rank receives
 * the same bits. Every rank must issue the same sequence of calls (offset, count, ctas); calls on one stream are ordered.
 *   peers[k] / flags[k]  THIS process's mapping of rank k's exchange buffer / flag words (HOST arrays of `world` device
 *                        pointers, read during the call). Flag words: uint32 [ctas][world] per rank, zero before the
 *                        first call, touched by nothing else; consecutive calls and graph replays reuse them.
 *   timeout_s            a rank that does not arrive within this time traps the kernel (<= 0: 10 s). */
typedef struct vp3d_allreduce_args {
  void* multicast;
  void* const* peers;
  void* const* flags;
  int rank, world;
  long long offset, count;   /* fp32 elements, multiples of 4; the buffers 16-byte aligned */
  float scale;               /* 1 / world for the gradient average */
  int ctas;                  /* even (clusters of two CTAs), 2..64 */
  double timeout_s;
} vp3d_allreduce_args;
int vp3d_peer_allreduce_f32(const vp3d_allreduce_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VP3D_B200_H */
