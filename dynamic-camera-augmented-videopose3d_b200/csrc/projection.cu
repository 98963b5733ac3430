// K5: per-frame camera projection (HBM-bandwidth bound).
//
// Replaces, for CUDA tensors, the ATen elementwise chains behind
//   common/quaternion.py:10-24  qrot        v + 2*(w*(q x v) + q x (q x v))
//   common/quaternion.py:27-35  qinverse    conjugate of a unit quaternion
//   common/camera.py:28-34      world_to_camera / camera_to_world
//   common/camera.py:37-67      project_to_2d         (3 radial + 2 tangential distortion terms)
//   common/camera.py:69-90      project_to_2d_linear
// and fuses world -> camera -> image plane for the dynamic-camera case (one quaternion + translation per frame).
//
// Arithmetic follows the reference operation by operation in fp32 with contraction disabled (explicit
// __fmul_rn/__fadd_rn), so results agree with the unfused PyTorch/NumPy evaluation to the last bit or two,
// including the clamp saturation for z -> 0 and NaN propagation for 0/0.
//
// Access pattern: each thread owns 4 consecutive points = three aligned float4 loads (48 B) and writes two float4
// of 2-D output; the per-frame camera record (quaternion 4 + translation 3 + intrinsics 9 floats) is read through the
// read-only path and is shared by the J joints of a frame. A scalar path covers tails and unaligned views.
#include <cstdlib>

#include "kernels.h"
#include "pdl.cuh"
#include "ptx.cuh"

namespace vp3d {

struct Quat {
  float w, x, y, z;
};

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }

// quaternion.py:21-24
__device__ __forceinline__ float3 qrot_dev(const Quat q, const float3 v) {
  float3 uv, uuv, r;
  uv.x = sub(mul(q.y, v.z), mul(q.z, v.y));
  uv.y = sub(mul(q.z, v.x), mul(q.x, v.z));
  uv.z = sub(mul(q.x, v.y), mul(q.y, v.x));
  uuv.x = sub(mul(q.y, uv.z), mul(q.z, uv.y));
  uuv.y = sub(mul(q.z, uv.x), mul(q.x, uv.z));
  uuv.z = sub(mul(q.x, uv.y), mul(q.y, uv.x));
  r.x = add(v.x, mul(2.f, add(mul(q.w, uv.x), uuv.x)));
  r.y = add(v.y, mul(2.f, add(mul(q.w, uv.y), uuv.y)));
  r.z = add(v.z, mul(2.f, add(mul(q.w, uv.z), uuv.z)));
  return r;
}

// torch.clamp semantics: NaN stays NaN (fminf/fmaxf would drop it)
__device__ __forceinline__ float clamp_unit(float x) {
  if (x != x) return x;
  return fminf(fmaxf(x, -1.f), 1.f);
}

// camera.py:54-67 (linear != 0: camera.py:85-90)
__device__ __forceinline__ float2 project_dev(const float3 X, const float* __restrict__ cam, int linear) {
  const float fx = __ldg(cam + 0), fy = __ldg(cam + 1), cx = __ldg(cam + 2), cy = __ldg(cam + 3);
  const float xx = clamp_unit(__fdiv_rn(X.x, X.z));
  const float yy = clamp_unit(__fdiv_rn(X.y, X.z));
  float2 o;
  if (linear) {
    o.x = add(mul(fx, xx), cx);
    o.y = add(mul(fy, yy), cy);
    return o;
  }
  const float k1 = __ldg(cam + 4), k2 = __ldg(cam + 5), k3 = __ldg(cam + 6), p1 = __ldg(cam + 7), p2 = __ldg(cam + 8);
  const float r2 = add(mul(xx, xx), mul(yy, yy));
  const float r4 = mul(r2, r2);
  const float r6 = mul(r4, r2);
  const float radial = add(1.f, add(add(mul(k1, r2), mul(k2, r4)), mul(k3, r6)));
  const float tan = add(mul(p1, xx), mul(p2, yy));
  const float s = add(radial, tan);
  o.x = add(mul(fx, add(mul(xx, s), mul(p1, r2))), cx);
  o.y = add(mul(fy, add(mul(yy, s), mul(p2, r2))), cy);
  return o;
}

// One point through the selected stages.
//   mode bit 0: subtract translation then rotate by the conjugate quaternion (world -> camera)
//   mode bit 1: rotate by the quaternion then add translation (camera -> world)
//   mode bit 2: rotate only (qrot); bit 3 with it: use the conjugate
//   mode bit 4: project to 2-D; bit 5: linear projection
struct PointOps {
  const float* q;      // [n_q][4] (w,x,y,z)
  const float* t;      // [n_q][3]
  const float* cam;    // [n_cam][9]
  long long pts_per_q;    // consecutive points sharing one quaternion/translation record
  long long pts_per_cam;  // consecutive points sharing one intrinsics record
  int mode;
};

__device__ __forceinline__ float3 transform_point(const PointOps& o, long long idx, float3 X) {
  if (o.mode & 7) {
    const long long qi = idx / o.pts_per_q;
    const float* qp = o.q + 4 * qi;
    Quat q{__ldg(qp + 0), __ldg(qp + 1), __ldg(qp + 2), __ldg(qp + 3)};
    if (o.mode & 1) {
      const float* tt = o.t + 3 * qi;
      X.x = sub(X.x, __ldg(tt + 0));
      X.y = sub(X.y, __ldg(tt + 1));
      X.z = sub(X.z, __ldg(tt + 2));
      q.x = -q.x;
      q.y = -q.y;
      q.z = -q.z;
      X = qrot_dev(q, X);
    } else if (o.mode & 2) {
      const float* tt = o.t + 3 * qi;
      X = qrot_dev(q, X);
      X.x = add(X.x, __ldg(tt + 0));
      X.y = add(X.y, __ldg(tt + 1));
      X.z = add(X.z, __ldg(tt + 2));
    } else {
      if (o.mode & 8) {
        q.x = -q.x;
        q.y = -q.y;
        q.z = -q.z;
      }
      X = qrot_dev(q, X);
    }
  }
  return X;
}

// Camera records in registers: consecutive points mostly share their frame's quaternion / translation / intrinsics, so
// a thread divides once per 4-point group (32-bit when the point count allows) and afterwards only counts up,
// reloading a record when its index changes. (The first version divided twice per point in 64 bits and re-read the
// 16 camera floats per point: the kernel was instruction-bound at 37 % of HBM peak.)
struct CamRegs {
  Quat q;
  float3 t;
  float cam[9];
};

template <typename IndexT>
struct RecordCursor {
  IndexT idx, rem, per;
  __device__ __forceinline__ RecordCursor(IndexT point, IndexT per_) : per(per_) {
    idx = point / per_;
    rem = point - idx * per_;
  }
  // advances to the next point; true when the record index changed
  __device__ __forceinline__ bool next() {
    if (++rem == per) {
      rem = 0;
      ++idx;
      return true;
    }
    return false;
  }
};

__device__ __forceinline__ void load_pose(const PointOps& o, long long qi, CamRegs& r) {
  const float4 qq = __ldg(reinterpret_cast<const float4*>(o.q) + qi);   // rows of 4 floats: always 16-byte aligned
  // world -> camera and ROTATE|CONJ rotate by the conjugate: negate once per record, not once per point
  const bool conj = (o.mode & 1) || ((o.mode & 4) && (o.mode & 8));
  r.q = conj ? Quat{qq.x, -qq.y, -qq.z, -qq.w} : Quat{qq.x, qq.y, qq.z, qq.w};
  if (o.mode & 3) {
    const float* tt = o.t + 3 * qi;
    r.t = make_float3(__ldg(tt + 0), __ldg(tt + 1), __ldg(tt + 2));
  }
}
__device__ __forceinline__ void load_intrinsics(const PointOps& o, long long ci, CamRegs& r) {
  const float* c = o.cam + 9 * ci;
#pragma unroll
  for (int k = 0; k < 9; ++k) r.cam[k] = __ldg(c + k);
}

__device__ __forceinline__ float3 transform_regs(const PointOps& o, const CamRegs& r, float3 X) {
  if (o.mode & 7) {
    if (o.mode & 1) {
      X.x = sub(X.x, r.t.x);
      X.y = sub(X.y, r.t.y);
      X.z = sub(X.z, r.t.z);
      X = qrot_dev(r.q, X);
    } else if (o.mode & 2) {
      X = qrot_dev(r.q, X);
      X.x = add(X.x, r.t.x);
      X.y = add(X.y, r.t.y);
      X.z = add(X.z, r.t.z);
    } else {
      X = qrot_dev(r.q, X);
    }
  }
  return X;
}

// camera.py:54-67 with the intrinsics already in registers
__device__ __forceinline__ float2 project_regs(const float3 X, const float (&cam)[9], int linear) {
  const float xx = clamp_unit(__fdiv_rn(X.x, X.z));
  const float yy = clamp_unit(__fdiv_rn(X.y, X.z));
  float2 o;
  if (linear) {
    o.x = add(mul(cam[0], xx), cam[2]);
    o.y = add(mul(cam[1], yy), cam[3]);
    return o;
  }
  const float r2 = add(mul(xx, xx), mul(yy, yy));
  const float r4 = mul(r2, r2);
  const float r6 = mul(r4, r2);
  const float radial = add(1.f, add(add(mul(cam[4], r2), mul(cam[5], r4)), mul(cam[6], r6)));
  const float tan = add(mul(cam[7], xx), mul(cam[8], yy));
  const float s = add(radial, tan);
  o.x = add(mul(cam[0], add(mul(xx, s), mul(cam[7], r2))), cam[2]);
  o.y = add(mul(cam[1], add(mul(yy, s), mul(cam[8], r2))), cam[3]);
  return o;
}

// ---- contracted-arithmetic variants (mode bit VP3D_PT_FAST): the same formulas written for the compiler to fuse into
// FMAs, one reciprocal for both ratios. Results differ from the un-contracted path in the last bits (<= 1e-6 relative,
// inside the 1e-5 budget of BASELINE.json); special values behave the same (x/0 -> +-inf -> clamp, 0/0 -> NaN). The
// un-contracted path executes 158 instructions per point and is issue bound at 58 % of HBM peak (ncu: 67 % issue
// active, 30 % DRAM); this one is what the fused world_to_image / feeder / streaming entry points use by default.
__device__ __forceinline__ float3 qrot_fast(const Quat q, const float3 v) {
  const float ux = q.y * v.z - q.z * v.y, uy = q.z * v.x - q.x * v.z, uz = q.x * v.y - q.y * v.x;
  const float wx = q.y * uz - q.z * uy, wy = q.z * ux - q.x * uz, wz = q.x * uy - q.y * ux;
  return make_float3(fmaf(2.f, fmaf(q.w, ux, wx), v.x), fmaf(2.f, fmaf(q.w, uy, wy), v.y),
                     fmaf(2.f, fmaf(q.w, uz, wz), v.z));
}
__device__ __forceinline__ float3 transform_fast(const PointOps& o, const CamRegs& r, float3 X) {
  if (o.mode & 7) {
    if (o.mode & 1) {
      X = qrot_fast(r.q, make_float3(X.x - r.t.x, X.y - r.t.y, X.z - r.t.z));
    } else if (o.mode & 2) {
      X = qrot_fast(r.q, X);
      X = make_float3(X.x + r.t.x, X.y + r.t.y, X.z + r.t.z);
    } else {
      X = qrot_fast(r.q, X);
    }
  }
  return X;
}
__device__ __forceinline__ float2 project_fast(const float3 X, const float (&cam)[9], int linear) {
  const float iz = 1.f / X.z;   // IEEE reciprocal (no -use_fast_math): 0 -> inf, so x * iz keeps the reference's edge cases
  const float xx = clamp_unit(X.x * iz), yy = clamp_unit(X.y * iz);
  if (linear) return make_float2(fmaf(cam[0], xx, cam[2]), fmaf(cam[1], yy, cam[3]));
  const float r2 = fmaf(xx, xx, yy * yy);
  const float s = 1.f + r2 * fmaf(r2, fmaf(r2, cam[6], cam[5]), cam[4]) + fmaf(cam[7], xx, cam[8] * yy);
  return make_float2(fmaf(cam[0], fmaf(xx, s, cam[7] * r2), cam[2]), fmaf(cam[1], fmaf(yy, s, cam[8] * r2), cam[3]));
}

template <typename IndexT, bool FAST>
__global__ void __launch_bounds__(256, 4)
project_points_kernel(const float* __restrict__ X, float* __restrict__ out3, float* __restrict__ out2, long long n_pts,
                      PointOps o, int vec_ok) {
  pdl_enter();
  const long long n_quads = vec_ok ? (n_pts >> 2) : 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool has_pose = (o.mode & 7) != 0, has_proj = (o.mode & 16) != 0;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_quads; g += stride) {
    const float4* src = reinterpret_cast<const float4*>(X) + 3 * g;
    const float4 a = __ldg(src + 0), b = __ldg(src + 1), c = __ldg(src + 2);
    float3 P[4] = {{a.x, a.y, a.z}, {a.w, b.x, b.y}, {b.z, b.w, c.x}, {c.y, c.z, c.w}};
    float2 Q[4];
    CamRegs r;
    RecordCursor<IndexT> qc((IndexT)(4 * g), (IndexT)o.pts_per_q), cc((IndexT)(4 * g), (IndexT)o.pts_per_cam);
    if (has_pose) load_pose(o, (long long)qc.idx, r);
    if (has_proj) load_intrinsics(o, (long long)cc.idx, r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      P[i] = FAST ? transform_fast(o, r, P[i]) : transform_regs(o, r, P[i]);
      if (has_proj) Q[i] = FAST ? project_fast(P[i], r.cam, o.mode & 32) : project_regs(P[i], r.cam, o.mode & 32);
      if (i < 3) {
        if (has_pose && qc.next()) load_pose(o, (long long)qc.idx, r);
        if (has_proj && cc.next()) load_intrinsics(o, (long long)cc.idx, r);
      }
    }
    if (out3 != nullptr) {
      float4* d = reinterpret_cast<float4*>(out3) + 3 * g;
      d[0] = make_float4(P[0].x, P[0].y, P[0].z, P[1].x);
      d[1] = make_float4(P[1].y, P[1].z, P[2].x, P[2].y);
      d[2] = make_float4(P[2].z, P[3].x, P[3].y, P[3].z);
    }
    if (out2 != nullptr) {
      float4* d = reinterpret_cast<float4*>(out2) + 2 * g;
      d[0] = make_float4(Q[0].x, Q[0].y, Q[1].x, Q[1].y);
      d[1] = make_float4(Q[2].x, Q[2].y, Q[3].x, Q[3].y);
    }
  }
  // scalar tail (and the whole range when the buffers are not 16-byte aligned)
  for (long long idx = 4 * n_quads + (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n_pts; idx += stride) {
    float3 P = {X[3 * idx], X[3 * idx + 1], X[3 * idx + 2]};
    float2 Q = make_float2(0.f, 0.f);
    if (FAST) {   // same arithmetic as the vector path, so both agree bit for bit
      CamRegs r;
      if (has_pose) load_pose(o, idx / o.pts_per_q, r);
      P = transform_fast(o, r, P);
      if (out2 != nullptr) {
        load_intrinsics(o, idx / o.pts_per_cam, r);
        Q = project_fast(P, r.cam, o.mode & 32);
      }
    } else {
      P = transform_point(o, idx, P);
      if (out2 != nullptr) Q = project_dev(P, o.cam + 9 * (idx / o.pts_per_cam), o.mode & 32);
    }
    if (out3 != nullptr) {
      out3[3 * idx] = P.x;
      out3[3 * idx + 1] = P.y;
      out3[3 * idx + 2] = P.z;
    }
    if (out2 != nullptr) {
      out2[2 * idx] = Q.x;
      out2[2 * idx + 1] = Q.y;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Window feeder: the batch assembly of common/generators.py:102-132 (ChunkedGenerator.next_epoch) fused with the
// dynamic-camera projection. Sequences stay resident on the device as concatenated world-space joints with one camera
// pose per frame; for sample b and window frame k the source frame is
//     f = seq_start[s] + clamp(start_3d[b] - pad - causal_shift + k, 0, len[s] - 1)      (np.pad 'edge', :92-100)
// and the kernel writes project_to_2d(world_to_camera(X[f], q[f], t[f]), cam[s]) straight into the (B, window, J, 2)
// batch -- no (B, 243, J, 3) intermediate, no host loop. Optionally the camera-space target of the chunk frames
// (root-relative like run.py:72-74) and the per-frame K @ [R|t] matrices the sibling models consume (:119-125).
struct WindowParams {
  const float* x;        // [frames][J][3]
  const float* q;        // [frames][4]
  const float* t;        // [frames][3]
  const float* cam;      // [n_seq][9]
  const long long* seq_start;
  const long long* seq_len;
  const int* sample_seq;
  const long long* sample_start;
  int batch, joints, chunk, pad, shift, root_relative, linear, window;
};

__device__ __forceinline__ long long window_frame(const WindowParams& w, int s, long long local) {
  const long long len = w.seq_len[s];
  local = local < 0 ? 0 : (local >= len ? len - 1 : local);
  return w.seq_start[s] + local;
}

__device__ __forceinline__ float3 to_camera(const WindowParams& w, long long f, int j) {
  const float4 qq = __ldg(reinterpret_cast<const float4*>(w.q) + f);
  const Quat qc{qq.x, -qq.y, -qq.z, -qq.w};
  const float* xp = w.x + (f * w.joints + j) * 3;
  const float* tp = w.t + f * 3;
  float3 X = make_float3(sub(__ldg(xp), __ldg(tp)), sub(__ldg(xp + 1), __ldg(tp + 1)), sub(__ldg(xp + 2), __ldg(tp + 2)));
  return qrot_dev(qc, X);
}

__global__ void __launch_bounds__(256)
project_windows_kernel(WindowParams w, float* __restrict__ out2) {
  pdl_enter();
  const unsigned total = (unsigned)w.batch * w.window * w.joints;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned j = i % w.joints;
    const unsigned bk = i / w.joints;
    const unsigned k = bk % w.window;
    const unsigned b = bk / w.window;
    const int s = w.sample_seq[b];
    const long long f = window_frame(w, s, w.sample_start[b] - w.pad - w.shift + (long long)k);
    const float3 Xc = to_camera(w, f, (int)j);
    const float2 P = project_dev(Xc, w.cam + 9 * s, w.linear);
    reinterpret_cast<float2*>(out2)[i] = P;
  }
}

__global__ void __launch_bounds__(256)
window_targets_kernel(WindowParams w, float* __restrict__ target3) {
  pdl_enter();
  const unsigned total = (unsigned)w.batch * w.chunk * w.joints;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned j = i % w.joints;
    const unsigned bc = i / w.joints;
    const unsigned c = bc % w.chunk;
    const unsigned b = bc / w.chunk;
    const int s = w.sample_seq[b];
    const long long f = window_frame(w, s, w.sample_start[b] + (long long)c);
    float3 X = to_camera(w, f, (int)j);
    if (w.root_relative) {
      const float3 R = to_camera(w, f, 0);
      X = make_float3(sub(X.x, R.x), sub(X.y, R.y), sub(X.z, R.z));
    }
    target3[3 * i] = X.x;
    target3[3 * i + 1] = X.y;
    target3[3 * i + 2] = X.z;
  }
}

// K @ [R | -R c] per window frame: R = rotation of conj(q) (world -> camera), c = camera position t
__global__ void __launch_bounds__(256)
window_cameras_kernel(WindowParams w, float* __restrict__ cam3x4) {
  pdl_enter();
  const unsigned total = (unsigned)w.batch * w.window;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned k = i % w.window;
    const unsigned b = i / w.window;
    const int s = w.sample_seq[b];
    const long long f = window_frame(w, s, w.sample_start[b] - w.pad - w.shift + (long long)k);
    const float4 qq = __ldg(reinterpret_cast<const float4*>(w.q) + f);
    const Quat qc{qq.x, -qq.y, -qq.z, -qq.w};
    const float3 e0 = qrot_dev(qc, make_float3(1.f, 0.f, 0.f));   // columns of R
    const float3 e1 = qrot_dev(qc, make_float3(0.f, 1.f, 0.f));
    const float3 e2 = qrot_dev(qc, make_float3(0.f, 0.f, 1.f));
    const float* tp = w.t + f * 3;
    const float3 tc = qrot_dev(qc, make_float3(-__ldg(tp), -__ldg(tp + 1), -__ldg(tp + 2)));
    const float* cm = w.cam + 9 * s;
    const float fx = __ldg(cm), fy = __ldg(cm + 1), cx = __ldg(cm + 2), cy = __ldg(cm + 3);
    const float E[3][4] = {{e0.x, e1.x, e2.x, tc.x}, {e0.y, e1.y, e2.y, tc.y}, {e0.z, e1.z, e2.z, tc.z}};
    float* o = cam3x4 + (long long)i * 12;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      o[c] = add(mul(fx, E[0][c]), mul(cx, E[2][c]));
      o[4 + c] = add(mul(fy, E[1][c]), mul(cy, E[2][c]));
      o[8 + c] = E[2][c];
    }
  }
}

cudaError_t launch_project_windows(const float* x, const float* q, const float* t, const float* cam,
                                   const long long* seq_start, const long long* seq_len, const int* sample_seq,
                                   const long long* sample_start, int batch, int joints, int chunk, int pad, int shift,
                                   int root_relative, int linear, float* out2, float* target3, float* cam3x4,
                                   int sm_count, cudaStream_t stream) {
  WindowParams w{x, q, t, cam, seq_start, seq_len, sample_seq, sample_start, batch, joints, chunk, pad, shift,
                 root_relative, linear, chunk + 2 * pad};
  auto grid_for = [&](long long total) {
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)sm_count * 8) blocks = (long long)sm_count * 8;
    return (unsigned)(blocks < 1 ? 1 : blocks);
  };
  launch_k(project_windows_kernel, dim3(grid_for((long long)batch * w.window * joints)), dim3(256), 0, stream, w, out2);
  if (target3 != nullptr)
    launch_k(window_targets_kernel, dim3(grid_for((long long)batch * chunk * joints)), dim3(256), 0, stream, w, target3);
  if (cam3x4 != nullptr) launch_k(window_cameras_kernel, dim3(grid_for((long long)batch * w.window)), dim3(256), 0, stream, w, cam3x4);
  return cudaGetLastError();
}

// Backward of project_to_2d / project_to_2d_linear wrt the camera-space points (the reference's docstring calls the
// projection "differentiable", camera.py:39-40; autograd derives exactly this):  gx = d sum(g * proj(X)) / dX.
__global__ void __launch_bounds__(256)
project_bwd_kernel(const float* __restrict__ X, const float* __restrict__ cam, const float* __restrict__ g,
                   long long n_pts, long long pts_per_cam, int linear, float* __restrict__ gx) {
  pdl_enter();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pts; i += stride) {
    const float* c = cam + 9 * (i / pts_per_cam);
    const float x = X[3 * i], y = X[3 * i + 1], z = X[3 * i + 2];
    const float rx = x / z, ry = y / z;
    const float xx = clamp_unit(rx), yy = clamp_unit(ry);
    const float a = __ldg(c) * g[2 * i], b = __ldg(c + 1) * g[2 * i + 1];
    float gxx = a, gyy = b;
    if (!linear) {
      const float k1 = __ldg(c + 4), k2 = __ldg(c + 5), k3 = __ldg(c + 6), p1 = __ldg(c + 7), p2 = __ldg(c + 8);
      const float r2 = xx * xx + yy * yy;
      const float s = 1.f + r2 * (k1 + r2 * (k2 + r2 * k3)) + p1 * xx + p2 * yy;
      const float gs = a * xx + b * yy;
      const float gr2 = a * p1 + b * p2 + gs * (k1 + r2 * (2.f * k2 + 3.f * k3 * r2));
      gxx = a * s + gs * p1 + gr2 * 2.f * xx;
      gyy = b * s + gs * p2 + gr2 * 2.f * yy;
    }
    if (!(rx >= -1.f && rx <= 1.f)) gxx = 0.f;   // torch.clamp backward: pass inside [-1, 1], bounds included
    if (!(ry >= -1.f && ry <= 1.f)) gyy = 0.f;
    const float iz = 1.f / z;
    gx[3 * i] = gxx * iz;
    gx[3 * i + 1] = gyy * iz;
    gx[3 * i + 2] = -(gxx * x + gyy * y) * iz * iz;
  }
}

cudaError_t launch_project_bwd(const float* X, const float* cam, const float* g, long long n_pts, long long pts_per_cam,
                               int linear, float* gx, int sm_count, cudaStream_t stream) {
  long long blocks = (n_pts + 255) / 256;
  if (blocks > (long long)sm_count * 8) blocks = (long long)sm_count * 8;
  if (blocks < 1) blocks = 1;
  launch_k(project_bwd_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, X, cam, g, n_pts, pts_per_cam > 0 ? pts_per_cam : 1, linear, gx);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Dynamic-camera fast path: world -> camera -> image plane with ONE camera pose per frame (the per-frame form of
// camera.py:28-30 + camera.py:37-67), VP3D_PT_FAST arithmetic.
//
// project_points_kernel above is instruction bound (a quad of points per thread: record cursors, a quaternion rotation
// per point, an IEEE division, 3 strided 16-byte loads). This kernel spends its instructions differently:
//   * persistent CTAs walk tiles of F whole frames; a tile's joints, quaternions and translations are three contiguous
//     byte ranges, fetched by three bulk async copies (cp.async.bulk, the untiled TMA path) into a ring of STAGES
//     shared-memory buffers signalled on mbarriers -- no load instructions, no address arithmetic, and
//     (STAGES - 1) tiles per CTA in flight whatever the occupancy;
//   * per frame (not per point) one thread turns the conjugate quaternion into the 3x3 matrix of the same linear map
//     (I + 2w[q]x + 2[q]x^2, identical to qrot for any q, unit or not) and another fetches the frame's intrinsics;
//     both land in a 96-byte shared-memory record;
//   * per point: 3 conflict-free LDS (stride 3 words), the record as 6 broadcast LDS.128, 3 subtractions + 9 FMAs,
//     one MUFU reciprocal (<= 1 ulp), NaN-propagating min/max for torch.clamp, the distortion polynomial in Horner
//     form, one coalesced 8-byte store. ~50 instructions instead of ~90.
// Differences from the un-contracted path stay <= 1e-6 absolute on normalised image coordinates (tests: 1e-5 budget).
struct FrameParams {
  const float* x;      // [frames][J][3]
  const float* q;      // [frames][4]
  const float* t;      // [frames][3]
  const float* cam;    // [frames / frames_per_cam][9]
  float* out3;         // optional camera-space points
  float* out2;         // image-plane points
  unsigned n_frames, joints, frames_per_cam, tile_frames, joint_magic;
  int linear;
};

constexpr int kFrameThreads = 256;
constexpr int kFrameMaxTile = 128;       // frames per tile (<= kFrameThreads / 2: one record thread + one intrinsics thread)
constexpr int kFrameMaxPoints = 2304;    // points per tile (27 KB of joints per stage)

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // one MUFU; the non-ftz form adds a 6-instruction range fix-up
  return y;
}
// torch.clamp(x, -1, 1): NaN stays NaN
__device__ __forceinline__ float clamp_unit_nan(float x) {
  float y;
  asm("max.NaN.f32 %0, %1, 0fBF800000;\n\tmin.NaN.f32 %0, %0, 0f3F800000;" : "=f"(y) : "f"(x));
  return y;
}

template <int STAGES, bool HAS3>
__global__ void __launch_bounds__(kFrameThreads, 3)
project_frames_kernel(const FrameParams p) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char frame_smem[];
  const unsigned F = p.tile_frames, J = p.joints;
  const unsigned x_bytes = F * J * 12, q_bytes = F * 16, t_bytes = F * 12;   // all multiples of 16 (F % 4 == 0)
  const unsigned stage_bytes = x_bytes + q_bytes + t_bytes;
  float4* rec = reinterpret_cast<float4*>(frame_smem + STAGES * stage_bytes);   // [F][6]
  uint64_t* bars = reinterpret_cast<uint64_t*>(rec + 6 * F);
  const unsigned tid = threadIdx.x;
  const unsigned n_tiles = (p.n_frames + F - 1) / F;

  auto issue = [&](unsigned tile, unsigned s) {   // one thread; `tile` is a full tile
    unsigned char* dst = frame_smem + s * stage_bytes;
    const size_t f0 = (size_t)tile * F;
    mbar_expect_tx(&bars[s], stage_bytes);
    bulk_load_1d(dst, p.x + f0 * J * 3, x_bytes, &bars[s]);
    bulk_load_1d(dst + x_bytes, p.q + f0 * 4, q_bytes, &bars[s]);
    bulk_load_1d(dst + x_bytes + q_bytes, p.t + f0 * 3, t_bytes, &bars[s]);
  };
  auto is_full = [&](unsigned tile) { return (size_t)(tile + 1) * F <= p.n_frames; };

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
    for (unsigned s = 0; s < STAGES; ++s) {
      const unsigned tile = blockIdx.x + s * gridDim.x;
      if (tile < n_tiles && is_full(tile)) issue(tile, s);
    }
  }
  __syncthreads();

  unsigned it = 0;
  for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const unsigned s = it % STAGES, parity = (it / STAGES) & 1;
    const unsigned f0 = tile * F;
    const unsigned nf = min(F, p.n_frames - f0);
    const unsigned npts = nf * J;
    const float* sx = reinterpret_cast<const float*>(frame_smem + s * stage_bytes);
    const float4* sq = reinterpret_cast<const float4*>(frame_smem + s * stage_bytes + x_bytes);
    const float* st = reinterpret_cast<const float*>(frame_smem + s * stage_bytes + x_bytes + q_bytes);

    // intrinsics of this tile's frames: independent of the tile data, so requested before waiting for it
    float c[9];
    const bool cam_thread = tid >= kFrameMaxTile && tid - kFrameMaxTile < nf;
    if (cam_thread) {
      const float* cp = p.cam + 9 * (size_t)((f0 + tid - kFrameMaxTile) / p.frames_per_cam);
#pragma unroll
      for (int k = 0; k < 9; ++k) c[k] = __ldg(cp + k);
    }

    if (nf == F) {
      mbar_wait(&bars[s], parity);
    } else {   // ragged last tile (sizes not multiples of 16 bytes): ordinary loads; its stage is idle by construction
      float* wx = const_cast<float*>(sx);
      float* wq = reinterpret_cast<float*>(const_cast<float4*>(sq));
      float* wt = const_cast<float*>(st);
      for (unsigned i = tid; i < npts * 3; i += kFrameThreads) wx[i] = __ldg(p.x + (size_t)f0 * J * 3 + i);
      for (unsigned i = tid; i < nf * 4; i += kFrameThreads) wq[i] = __ldg(p.q + (size_t)f0 * 4 + i);
      for (unsigned i = tid; i < nf * 3; i += kFrameThreads) wt[i] = __ldg(p.t + (size_t)f0 * 3 + i);
      __syncthreads();
    }

    if (tid < nf) {   // rows of the matrix of v -> qrot(conj(q), v), each with its translation component
      const float4 qq = sq[tid];
      const float w = qq.x, x = -qq.y, y = -qq.z, z = -qq.w;
      const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
      rec[6 * tid + 0] = make_float4(fmaf(-2.f, yy + zz, 1.f), 2.f * (xy - wz), 2.f * (xz + wy), st[3 * tid + 0]);
      rec[6 * tid + 1] = make_float4(2.f * (xy + wz), fmaf(-2.f, xx + zz, 1.f), 2.f * (yz - wx), st[3 * tid + 1]);
      rec[6 * tid + 2] = make_float4(2.f * (xz - wy), 2.f * (yz + wx), fmaf(-2.f, xx + yy, 1.f), st[3 * tid + 2]);
    } else if (cam_thread) {
      const unsigned lf = tid - kFrameMaxTile;
      rec[6 * lf + 3] = make_float4(c[0], c[1], c[2], c[3]);
      rec[6 * lf + 4] = make_float4(c[4], c[5], c[6], c[7]);
      rec[6 * lf + 5] = make_float4(c[8], 0.f, 0.f, 0.f);
    }
    __syncthreads();

    float2* o2 = reinterpret_cast<float2*>(p.out2) + (size_t)f0 * J;
    float* o3 = HAS3 ? p.out3 + (size_t)f0 * J * 3 : nullptr;
#pragma unroll 2
    for (unsigned i = tid; i < npts; i += kFrameThreads) {
      const unsigned f = __umulhi(i, p.joint_magic);   // i / J, exact for i * J < 2^32 (J >= 2)
      const float4 r0 = rec[6 * f + 0], r1 = rec[6 * f + 1], r2 = rec[6 * f + 2];
      const float dx = sx[3 * i] - r0.w, dy = sx[3 * i + 1] - r1.w, dz = sx[3 * i + 2] - r2.w;
      const float X = fmaf(r0.z, dz, fmaf(r0.y, dy, r0.x * dx));
      const float Y = fmaf(r1.z, dz, fmaf(r1.y, dy, r1.x * dx));
      const float Z = fmaf(r2.z, dz, fmaf(r2.y, dy, r2.x * dx));
      if (HAS3) {
        o3[3 * i] = X;
        o3[3 * i + 1] = Y;
        o3[3 * i + 2] = Z;
      }
      {
        const float4 c0 = rec[6 * f + 3];
        const float iz = rcp_approx(Z);   // 0 -> inf: x * inf -> +-inf -> clamp, 0 * inf -> NaN, like x / 0 and 0 / 0
        const float u = clamp_unit_nan(X * iz), v = clamp_unit_nan(Y * iz);
        float2 o;
        if (p.linear) {
          o = make_float2(fmaf(c0.x, u, c0.z), fmaf(c0.y, v, c0.w));
        } else {
          const float4 c1 = rec[6 * f + 4];
          const float p2 = rec[6 * f + 5].x;
          const float r2s = fmaf(u, u, v * v);
          const float sden = 1.f + r2s * fmaf(r2s, fmaf(r2s, c1.z, c1.y), c1.x) + fmaf(c1.w, u, p2 * v);
          o = make_float2(fmaf(c0.x, fmaf(u, sden, c1.w * r2s), c0.z), fmaf(c0.y, fmaf(v, sden, p2 * r2s), c0.w));
        }
        o2[i] = o;
      }
    }
    __syncthreads();   // the stage and the records are free again

    if (tid == 0) {
      const unsigned next = tile + STAGES * gridDim.x;
      if (next < n_tiles && is_full(next)) issue(next, s);
    }
  }
}

// Tuning knobs for A/B runs: VP3D_PROJ_TUNE="stages,max_points_per_tile,ctas_per_sm"; "0" disables the frame kernel
// (everything takes project_points_kernel). Measured on B200 (tools/proj_probe.py, fraction of the 6565 GB/s copy peak at
// 1024 / 4096 windows of 243 x 17 joints): 3,2304,2 -> 0.59 / 0.77; 2,2304,3 -> 0.73 / 0.85; 2,1536,4 -> 0.73 / 0.86
// (default); 3,1152,4 -> 0.70 / 0.85: resident threads matter more than bytes in flight (the kernel still issues ~50
// instructions per point), so two stages and four 256-thread CTAs per SM (64 registers) win.
struct FrameTune {
  int stages = 2, max_points = 1536, ctas = 4;
  FrameTune() {
    if (const char* e = std::getenv("VP3D_PROJ_TUNE")) {
      int a = 0, b = 0, c = 0;
      const int n = std::sscanf(e, "%d,%d,%d", &a, &b, &c);
      if (n >= 1) stages = a;
      if (n >= 2 && b > 0) max_points = b;
      if (n >= 3 && c > 0) ctas = c;
    }
  }
};

static inline bool aligned16(const void* p);

// Returns cudaErrorNotSupported when the call does not fit the frame kernel (the caller then uses the generic one).
static cudaError_t try_launch_project_frames(const float* X, float* out3, float* out2, long long n_pts, const float* q,
                                             const float* t, const float* cam, long long pts_per_q,
                                             long long pts_per_cam, int mode, int sm_count, cudaStream_t stream) {
  static const FrameTune tune;
  const int want = VP3D_PT_WORLD_TO_CAMERA | VP3D_PT_PROJECT | VP3D_PT_FAST;
  if (tune.stages < 2 || tune.stages > 3) return cudaErrorNotSupported;
  if ((mode & ~VP3D_PT_LINEAR) != want || out2 == nullptr) return cudaErrorNotSupported;
  const long long J = pts_per_q;
  if (J < 2 || J > tune.max_points / 4 || n_pts % J != 0 || pts_per_cam % J != 0) return cudaErrorNotSupported;
  const long long n_frames = n_pts / J;
  if (n_frames >= (1LL << 31) / 2 || n_pts >= (1LL << 40)) return cudaErrorNotSupported;
  if (!aligned16(X) || !aligned16(q) || !aligned16(t) || (reinterpret_cast<uintptr_t>(out2) & 7) ||
      (reinterpret_cast<uintptr_t>(out3) & 3))
    return cudaErrorNotSupported;
  // tile size: whole frames, a multiple of 4 (16-byte granularity of the bulk copies), chosen so that the tile count
  // fills the last wave of the persistent grid
  long long fmax = tune.max_points / J;
  if (fmax > kFrameMaxTile) fmax = kFrameMaxTile;
  fmax &= ~3LL;
  const long long grid_cap = (long long)sm_count * tune.ctas;
  if (fmax < 4) return cudaErrorNotSupported;
  long long F = fmax, best = -1;
  for (long long f = fmax; f >= 4 && f >= fmax / 2; f -= 4) {   // frames the busiest CTA walks, + 4 per tile of overhead
    const long long tiles = (n_frames + f - 1) / f;
    const long long cost = ((tiles + grid_cap - 1) / grid_cap) * (f + 4);
    if (best < 0 || cost < best) best = cost, F = f;
  }
  const long long n_tiles = (n_frames + F - 1) / F;
  const unsigned grid = (unsigned)(n_tiles < grid_cap ? n_tiles : grid_cap);
  FrameParams p{X, q, t, cam, out3, out2, (unsigned)n_frames, (unsigned)J, (unsigned)(pts_per_cam / J), (unsigned)F,
                (unsigned)(((1ULL << 32) + J - 1) / J), (mode & VP3D_PT_LINEAR) ? 1 : 0};
  const size_t smem = (size_t)tune.stages * (F * J * 12 + F * 28) + F * 96 + 8 * tune.stages;
  if (smem > 200 * 1024) return cudaErrorNotSupported;
  static const cudaError_t configured = [] {   // opt in to > 48 KB of dynamic shared memory, once per process
    cudaError_t e = cudaSuccess;
    for (auto kernel : {project_frames_kernel<2, true>, project_frames_kernel<3, true>, project_frames_kernel<2, false>,
                        project_frames_kernel<3, false>}) {
      const cudaError_t r = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (r != cudaSuccess) e = r;
    }
    return e;
  }();
  if (configured != cudaSuccess) return configured;
  auto launch = [&](auto kernel) {
    launch_k(kernel, dim3(grid), dim3(kFrameThreads), smem, stream, p);
    return cudaGetLastError();
  };
  if (out3 != nullptr) return tune.stages == 2 ? launch(project_frames_kernel<2, true>) : launch(project_frames_kernel<3, true>);
  return tune.stages == 2 ? launch(project_frames_kernel<2, false>) : launch(project_frames_kernel<3, false>);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

cudaError_t launch_project_points(const float* X, float* out3, float* out2, long long n_pts, const float* q,
                                  const float* t, const float* cam, long long pts_per_q, long long pts_per_cam,
                                  int mode, int sm_count, cudaStream_t stream) {
  if (n_pts <= 0) return cudaSuccess;
  {
    const cudaError_t e = try_launch_project_frames(X, out3, out2, n_pts, q, t, cam, pts_per_q, pts_per_cam, mode,
                                                    sm_count, stream);
    if (e != cudaErrorNotSupported) return e;
  }
  PointOps o{q, t, cam, pts_per_q > 0 ? pts_per_q : 1, pts_per_cam > 0 ? pts_per_cam : 1, mode};
  const int vec_ok = aligned16(X) && (out3 == nullptr || aligned16(out3)) && (out2 == nullptr || aligned16(out2));
  const long long work = vec_ok ? ((n_pts + 3) >> 2) : n_pts;
  long long blocks = (work + 255) / 256;
  const long long cap = (long long)sm_count * 8;  // 8 resident CTAs of 256 threads per SM, grid-stride beyond
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const bool fast = (mode & 64) != 0;   // VP3D_PT_FAST
  if (n_pts < (1LL << 31)) {
    if (fast) launch_k(project_points_kernel<unsigned int, true>, dim3((unsigned)blocks), dim3(256), 0, stream, X, out3, out2, n_pts, o, vec_ok);
    else launch_k(project_points_kernel<unsigned int, false>, dim3((unsigned)blocks), dim3(256), 0, stream, X, out3, out2, n_pts, o, vec_ok);
  } else {
    launch_k(project_points_kernel<long long, false>, dim3((unsigned)blocks), dim3(256), 0, stream, X, out3, out2, n_pts, o, vec_ok);
  }
  return cudaGetLastError();
}

}  // namespace vp3d
