// K6: fused joint-error losses (HBM-bandwidth bound, 24 B per joint forward, +12 B backward).
//
// Replaces, for CUDA tensors, the sub / norm / mean chains of
//   common/loss.py:11-17  mpjpe            mean_{n,t,j} || pred - target ||_2
//   common/loss.py:21-27  weighted_mpjpe   mean_{n,t,j} w * || pred - target ||_2   (w broadcast over the joint grid)
//   common/loss.py:70-80  n_mpjpe          mpjpe(scale * pred, target), scale = mean_j<target,pred> / mean_j<pred,pred>
// Forward is a two-stage deterministic reduction (per-CTA partial sums in double, one finishing CTA), so the value
// does not depend on atomics ordering. Backward writes d loss / d pred = g/n * w * (pred - target) / ||pred - target||
// (zero where the distance is zero, as torch.linalg.norm's backward does).
#include "kernels.h"
#include "pdl.cuh"

namespace vp3d {

constexpr int kLossThreads = 256;

struct WeightView {
  const float* w;  // nullptr = unweighted
  long long T, J;  // joint grid (n, t, j) of the loss
  long long s_n, s_t, s_j;  // element strides of w over that grid (0 = broadcast)
};

__device__ __forceinline__ float weight_at(const WeightView& wv, long long idx) {
  if (wv.w == nullptr) return 1.f;
  const long long j = idx % wv.J;
  const long long nt = idx / wv.J;
  const long long t = nt % wv.T;
  const long long n = nt / wv.T;
  return __ldg(wv.w + n * wv.s_n + t * wv.s_t + j * wv.s_j);
}

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double warp_part[kLossThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x < 32) {
    r = (threadIdx.x < kLossThreads / 32) ? warp_part[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;  // valid in thread 0
}

__device__ __forceinline__ float dist3(float px, float py, float pz, float tx, float ty, float tz) {
  const float dx = px - tx, dy = py - ty, dz = pz - tz;
  return sqrtf(dx * dx + dy * dy + dz * dz);
}

__global__ void __launch_bounds__(kLossThreads)
mpjpe_partial_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, long long n_joints, WeightView wv,
                     int vec_ok, double* __restrict__ partial) {
  pdl_enter();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n_quads = vec_ok ? (n_joints >> 2) : 0;
  double acc = 0.0;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_quads; g += stride) {
    const float4* p4 = reinterpret_cast<const float4*>(pred) + 3 * g;
    const float4* t4 = reinterpret_cast<const float4*>(tgt) + 3 * g;
    const float4 a = __ldg(p4), b = __ldg(p4 + 1), c = __ldg(p4 + 2);
    const float4 d = __ldg(t4), e = __ldg(t4 + 1), f = __ldg(t4 + 2);
    const float e0 = dist3(a.x, a.y, a.z, d.x, d.y, d.z);
    const float e1 = dist3(a.w, b.x, b.y, d.w, e.x, e.y);
    const float e2 = dist3(b.z, b.w, c.x, e.z, e.w, f.x);
    const float e3 = dist3(c.y, c.z, c.w, f.y, f.z, f.w);
    if (wv.w == nullptr) {
      acc += (double)((e0 + e1) + (e2 + e3));
    } else {
      acc += (double)(weight_at(wv, 4 * g) * e0) + (double)(weight_at(wv, 4 * g + 1) * e1) +
             (double)(weight_at(wv, 4 * g + 2) * e2) + (double)(weight_at(wv, 4 * g + 3) * e3);
    }
  }
  for (long long i = 4 * n_quads + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_joints; i += stride) {
    const float e0 = dist3(pred[3 * i], pred[3 * i + 1], pred[3 * i + 2], tgt[3 * i], tgt[3 * i + 1], tgt[3 * i + 2]);
    acc += (double)(weight_at(wv, i) * e0);
  }
  const double s = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(kLossThreads)
mean_finish_kernel(const double* __restrict__ partial, int n_partial, double inv_count, float* __restrict__ out) {
  pdl_enter();
  double acc = 0.0;
  for (int i = threadIdx.x; i < n_partial; i += blockDim.x) acc += partial[i];
  const double s = block_sum(acc);
  if (threadIdx.x == 0) out[0] = (float)(s * inv_count);
}

__global__ void __launch_bounds__(kLossThreads)
mpjpe_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, const float* __restrict__ grad_out,
                 float inv_count, long long n_joints, WeightView wv, int vec_ok, float* __restrict__ grad_pred) {
  pdl_enter();
  const float g0 = __ldg(grad_out) * inv_count;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n_quads = vec_ok ? (n_joints >> 2) : 0;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_quads; g += stride) {
    const float4* p4 = reinterpret_cast<const float4*>(pred) + 3 * g;
    const float4* t4 = reinterpret_cast<const float4*>(tgt) + 3 * g;
    const float4 a = __ldg(p4), b = __ldg(p4 + 1), c = __ldg(p4 + 2);
    const float4 d = __ldg(t4), e = __ldg(t4 + 1), f = __ldg(t4 + 2);
    float v[12] = {a.x - d.x, a.y - d.y, a.z - d.z, a.w - d.w, b.x - e.x, b.y - e.y,
                   b.z - e.z, b.w - e.w, c.x - f.x, c.y - f.y, c.z - f.z, c.w - f.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float nrm = sqrtf(v[3 * i] * v[3 * i] + v[3 * i + 1] * v[3 * i + 1] + v[3 * i + 2] * v[3 * i + 2]);
      const float s = nrm > 0.f ? g0 * weight_at(wv, 4 * g + i) / nrm : 0.f;
      v[3 * i] *= s;
      v[3 * i + 1] *= s;
      v[3 * i + 2] *= s;
    }
    float4* o = reinterpret_cast<float4*>(grad_pred) + 3 * g;
    o[0] = make_float4(v[0], v[1], v[2], v[3]);
    o[1] = make_float4(v[4], v[5], v[6], v[7]);
    o[2] = make_float4(v[8], v[9], v[10], v[11]);
  }
  for (long long i = 4 * n_quads + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_joints; i += stride) {
    const float dx = pred[3 * i] - tgt[3 * i], dy = pred[3 * i + 1] - tgt[3 * i + 1], dz = pred[3 * i + 2] - tgt[3 * i + 2];
    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float s = nrm > 0.f ? g0 * weight_at(wv, i) / nrm : 0.f;
    grad_pred[3 * i] = dx * s;
    grad_pred[3 * i + 1] = dy * s;
    grad_pred[3 * i + 2] = dz * s;
  }
}

// n_mpjpe: one warp per (n, t) pose; lanes stride over the J joints of the pose.
__global__ void __launch_bounds__(kLossThreads)
n_mpjpe_partial_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, long long n_poses, int J,
                       double* __restrict__ partial) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  double acc = 0.0;
  for (long long pose = warp_global; pose < n_poses; pose += n_warps) {
    const float* p = pred + pose * J * 3;
    const float* t = tgt + pose * J * 3;
    float pp = 0.f, tp = 0.f;
    for (int j = lane; j < J; j += 32) {
      const float px = p[3 * j], py = p[3 * j + 1], pz = p[3 * j + 2];
      const float tx = t[3 * j], ty = t[3 * j + 1], tz = t[3 * j + 2];
      pp += px * px + py * py + pz * pz;
      tp += tx * px + ty * py + tz * pz;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      pp += __shfl_xor_sync(0xffffffffu, pp, o);
      tp += __shfl_xor_sync(0xffffffffu, tp, o);
    }
    const float scale = (tp / (float)J) / (pp / (float)J);
    float e = 0.f;
    for (int j = lane; j < J; j += 32)
      e += dist3(scale * p[3 * j], scale * p[3 * j + 1], scale * p[3 * j + 2], t[3 * j], t[3 * j + 1], t[3 * j + 2]);
    acc += (double)e;
  }
  const double s = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// Generic point dimension D (the reference's mpjpe is a norm over the last axis whatever its length, loss.py:17 --
// upstream uses it on 2-D reprojections): one thread per point, D strided reads.
__global__ void __launch_bounds__(kLossThreads)
mpjpe_nd_partial_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, long long n_pts, int D,
                        WeightView wv, double* __restrict__ partial) {
  pdl_enter();
  const long long stride = (long long)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pts; i += stride) {
    float s = 0.f;
    for (int d = 0; d < D; ++d) {
      const float v = pred[i * D + d] - tgt[i * D + d];
      s += v * v;
    }
    acc += (double)(weight_at(wv, i) * sqrtf(s));
  }
  const double r = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

__global__ void __launch_bounds__(kLossThreads)
mpjpe_nd_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, const float* __restrict__ grad_out,
                    float inv_count, long long n_pts, int D, WeightView wv, float* __restrict__ grad_pred) {
  pdl_enter();
  const float g0 = __ldg(grad_out) * inv_count;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pts; i += stride) {
    float s = 0.f;
    for (int d = 0; d < D; ++d) {
      const float v = pred[i * D + d] - tgt[i * D + d];
      s += v * v;
    }
    const float nrm = sqrtf(s);
    const float k = nrm > 0.f ? g0 * weight_at(wv, i) / nrm : 0.f;
    for (int d = 0; d < D; ++d) grad_pred[i * D + d] = (pred[i * D + d] - tgt[i * D + d]) * k;
  }
}

// Backward of n_mpjpe wrt the prediction (autograd of loss.py:77-80): with A = mean_j |P_j|^2, B = mean_j <T_j, P_j>,
// s = B / A, e_j = s P_j - T_j, u_j = e_j / |e_j| and w = sum_j <u_j, P_j>:
//   dL/dP_k = g / count * ( s u_k + w (T_k - 2 s P_k) / (J A) )
// One warp per (n, t) pose, lanes stride over the joints.
__global__ void __launch_bounds__(kLossThreads)
n_mpjpe_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, const float* __restrict__ grad_out,
                   float inv_count, long long n_poses, int J, float* __restrict__ grad_pred) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float g0 = __ldg(grad_out) * inv_count;
  for (long long pose = warp_global; pose < n_poses; pose += n_warps) {
    const float* p = pred + pose * J * 3;
    const float* t = tgt + pose * J * 3;
    float* o = grad_pred + pose * J * 3;
    float pp = 0.f, tp = 0.f;
    for (int j = lane; j < J; j += 32) {
      pp += p[3 * j] * p[3 * j] + p[3 * j + 1] * p[3 * j + 1] + p[3 * j + 2] * p[3 * j + 2];
      tp += t[3 * j] * p[3 * j] + t[3 * j + 1] * p[3 * j + 1] + t[3 * j + 2] * p[3 * j + 2];
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) {
      pp += __shfl_xor_sync(0xffffffffu, pp, k);
      tp += __shfl_xor_sync(0xffffffffu, tp, k);
    }
    const float A = pp / (float)J;
    const float s = (tp / (float)J) / A;
    float w = 0.f;
    for (int j = lane; j < J; j += 32) {
      const float ex = s * p[3 * j] - t[3 * j], ey = s * p[3 * j + 1] - t[3 * j + 1], ez = s * p[3 * j + 2] - t[3 * j + 2];
      const float nrm = sqrtf(ex * ex + ey * ey + ez * ez);
      if (nrm > 0.f) w += (ex * p[3 * j] + ey * p[3 * j + 1] + ez * p[3 * j + 2]) / nrm;
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) w += __shfl_xor_sync(0xffffffffu, w, k);
    const float c = w / ((float)J * A);
    for (int j = lane; j < J; j += 32) {
      const float px = p[3 * j], py = p[3 * j + 1], pz = p[3 * j + 2];
      const float ex = s * px - t[3 * j], ey = s * py - t[3 * j + 1], ez = s * pz - t[3 * j + 2];
      const float nrm = sqrtf(ex * ex + ey * ey + ez * ez);
      const float inv = nrm > 0.f ? s / nrm : 0.f;
      o[3 * j] = g0 * (ex * inv + c * (t[3 * j] - 2.f * s * px));
      o[3 * j + 1] = g0 * (ey * inv + c * (t[3 * j + 1] - 2.f * s * py));
      o[3 * j + 2] = g0 * (ez * inv + c * (t[3 * j + 2] - 2.f * s * pz));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Evaluation metrics on the device: p_mpjpe (loss.py:29-68, Procrustes-aligned MPJPE) and
// mean_velocity_error (loss.py:82-91). The reference computes both in NumPy on the host after a D2H copy per sequence
// (run.py:749-756). One thread per pose: centroids, norms, the 3x3 cross-covariance H = X0^T Y0, its SVD (cyclic Jacobi
// on H^T H in double, U recovered as H v / s), the reflection fix on the last singular vector, then the aligned error.
__device__ void jacobi_eig3(double A[3][3], double V[3][3], double w[3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off < 1e-30) break;
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      if (fabs(A[p][q]) < 1e-300) continue;
      const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
      const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
      for (int k = 0; k < 3; ++k) {   // A <- A J
        const double akp = A[k][p], akq = A[k][q];
        A[k][p] = c * akp - sn * akq;
        A[k][q] = sn * akp + c * akq;
      }
      for (int k = 0; k < 3; ++k) {   // A <- J^T A
        const double apk = A[p][k], aqk = A[q][k];
        A[p][k] = c * apk - sn * aqk;
        A[q][k] = sn * apk + c * aqk;
      }
      for (int k = 0; k < 3; ++k) {
        const double vkp = V[k][p], vkq = V[k][q];
        V[k][p] = c * vkp - sn * vkq;
        V[k][q] = sn * vkp + c * vkq;
      }
    }
  }
  for (int i = 0; i < 3; ++i) w[i] = A[i][i];
}

__global__ void __launch_bounds__(kLossThreads)
p_mpjpe_partial_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, long long n_poses, int J,
                       double* __restrict__ partial) {
  pdl_enter();
  const long long stride = (long long)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (long long pose = (long long)blockIdx.x * blockDim.x + threadIdx.x; pose < n_poses; pose += stride) {
    const float* Y = pred + pose * J * 3;   // "Y" = predicted, "X" = target, as in the reference
    const float* X = tgt + pose * J * 3;
    double muX[3] = {0, 0, 0}, muY[3] = {0, 0, 0};
    for (int j = 0; j < J; ++j)
      for (int d = 0; d < 3; ++d) {
        muX[d] += X[3 * j + d];
        muY[d] += Y[3 * j + d];
      }
    for (int d = 0; d < 3; ++d) {
      muX[d] /= J;
      muY[d] /= J;
    }
    double nX = 0, nY = 0, H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int j = 0; j < J; ++j) {
      double x[3], y[3];
      for (int d = 0; d < 3; ++d) {
        x[d] = X[3 * j + d] - muX[d];
        y[d] = Y[3 * j + d] - muY[d];
        nX += x[d] * x[d];
        nY += y[d] * y[d];
      }
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) H[a][b] += x[a] * y[b];
    }
    nX = sqrt(nX);
    nY = sqrt(nY);
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) H[a][b] /= nX * nY;   // H of the normalised point sets
    // SVD H = U diag(s) V^T
    double A[3][3], V[3][3], w[3];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) A[a][b] = H[0][a] * H[0][b] + H[1][a] * H[1][b] + H[2][a] * H[2][b];
    jacobi_eig3(A, V, w);
    int ord[3] = {0, 1, 2};   // descending singular values
    for (int i = 0; i < 2; ++i)
      for (int k = i + 1; k < 3; ++k)
        if (w[ord[k]] > w[ord[i]]) {
          const int tmp = ord[i];
          ord[i] = ord[k];
          ord[k] = tmp;
        }
    double sv[3], Vs[3][3], U[3][3];
    for (int i = 0; i < 3; ++i) {
      sv[i] = sqrt(w[ord[i]] > 0 ? w[ord[i]] : 0.0);
      for (int k = 0; k < 3; ++k) Vs[k][i] = V[k][ord[i]];
    }
    for (int i = 0; i < 3; ++i) {
      double u[3];
      for (int a = 0; a < 3; ++a) u[a] = H[a][0] * Vs[0][i] + H[a][1] * Vs[1][i] + H[a][2] * Vs[2][i];
      const double n = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
      if (i < 2 || n > 1e-12 * (sv[0] + 1e-300)) {
        for (int a = 0; a < 3; ++a) U[a][i] = n > 0 ? u[a] / n : (a == i ? 1.0 : 0.0);
      } else {   // rank-deficient H: complete the basis (the sign is settled by the reflection fix below)
        U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
        U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
        U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
      }
    }
    // R = V U^T, reflections removed by flipping the last singular vector / value
    double R[3][3];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) R[a][b] = Vs[a][0] * U[b][0] + Vs[a][1] * U[b][1] + Vs[a][2] * U[b][2];
    const double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                       R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
    if (det < 0) {
      for (int a = 0; a < 3; ++a) Vs[a][2] = -Vs[a][2];
      sv[2] = -sv[2];
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) R[a][b] = Vs[a][0] * U[b][0] + Vs[a][1] * U[b][1] + Vs[a][2] * U[b][2];
    }
    const double scale = (sv[0] + sv[1] + sv[2]) * nX / nY;
    double tvec[3];
    for (int b = 0; b < 3; ++b)
      tvec[b] = muX[b] - scale * (muY[0] * R[0][b] + muY[1] * R[1][b] + muY[2] * R[2][b]);
    double e = 0.0;
    for (int j = 0; j < J; ++j) {
      double d2 = 0.0;
      for (int b = 0; b < 3; ++b) {
        const double al = scale * (Y[3 * j] * R[0][b] + Y[3 * j + 1] * R[1][b] + Y[3 * j + 2] * R[2][b]) + tvec[b];
        const double df = al - X[3 * j + b];
        d2 += df * df;
      }
      e += sqrt(d2);
    }
    acc += e;
  }
  const double r = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

// mean over (t, point) of | (p[t+1] - p[t]) - (g[t+1] - g[t]) |_2 for arrays [T][inner points][D]
__global__ void __launch_bounds__(kLossThreads)
velocity_partial_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, long long T, long long inner, int D,
                        double* __restrict__ partial) {
  pdl_enter();
  const long long n = (T - 1) * inner;
  const long long stride = (long long)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long a = i * D, b = (i + inner) * D;
    float s = 0.f;
    for (int d = 0; d < D; ++d) {
      const float v = (pred[b + d] - pred[a + d]) - (tgt[b + d] - tgt[a + d]);
      s += v * v;
    }
    acc += (double)sqrtf(s);
  }
  const double r = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int loss_grid(long long n_joints, int sm_count) {
  long long blocks = ((n_joints + 3) / 4 + kLossThreads - 1) / kLossThreads;
  const long long cap = (long long)sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

cudaError_t launch_mpjpe_fwd(const float* pred, const float* tgt, long long n_joints, const float* w, long long T,
                             long long J, long long s_n, long long s_t, long long s_j, double* partial, float* out,
                             int sm_count, cudaStream_t stream) {
  WeightView wv{w, T > 0 ? T : 1, J > 0 ? J : 1, s_n, s_t, s_j};
  const int grid = loss_grid(n_joints, sm_count);
  const int vec_ok = aligned16(pred) && aligned16(tgt);
  launch_k(mpjpe_partial_kernel, dim3(grid), dim3(kLossThreads), 0, stream, pred, tgt, n_joints, wv, vec_ok, partial);
  launch_k(mean_finish_kernel, dim3(1), dim3(kLossThreads), 0, stream, partial, grid, n_joints > 0 ? 1.0 / (double)n_joints : 0.0, out);
  return cudaGetLastError();
}

cudaError_t launch_mpjpe_bwd(const float* pred, const float* tgt, const float* grad_out, long long n_joints,
                             const float* w, long long T, long long J, long long s_n, long long s_t, long long s_j,
                             float* grad_pred, int sm_count, cudaStream_t stream) {
  if (n_joints <= 0) return cudaSuccess;
  WeightView wv{w, T > 0 ? T : 1, J > 0 ? J : 1, s_n, s_t, s_j};
  const int grid = loss_grid(n_joints, sm_count);
  const int vec_ok = aligned16(pred) && aligned16(tgt) && aligned16(grad_pred);
  launch_k(mpjpe_bwd_kernel, dim3(grid), dim3(kLossThreads), 0, stream, pred, tgt, grad_out, 1.f / (float)n_joints, n_joints, wv, vec_ok,
                                                      grad_pred);
  return cudaGetLastError();
}

cudaError_t launch_mpjpe_nd_fwd(const float* pred, const float* tgt, long long n_pts, int D, const float* w, long long T,
                                long long J, long long s_n, long long s_t, long long s_j, double* partial, float* out,
                                int sm_count, cudaStream_t stream) {
  WeightView wv{w, T > 0 ? T : 1, J > 0 ? J : 1, s_n, s_t, s_j};
  const int grid = loss_grid(n_pts * 4, sm_count);
  launch_k(mpjpe_nd_partial_kernel, dim3(grid), dim3(kLossThreads), 0, stream, pred, tgt, n_pts, D, wv, partial);
  launch_k(mean_finish_kernel, dim3(1), dim3(kLossThreads), 0, stream, partial, grid, n_pts > 0 ? 1.0 / (double)n_pts : 0.0, out);
  return cudaGetLastError();
}

cudaError_t launch_mpjpe_nd_bwd(const float* pred, const float* tgt, const float* grad_out, long long n_pts, int D,
                                const float* w, long long T, long long J, long long s_n, long long s_t, long long s_j,
                                float* grad_pred, int sm_count, cudaStream_t stream) {
  if (n_pts <= 0) return cudaSuccess;
  WeightView wv{w, T > 0 ? T : 1, J > 0 ? J : 1, s_n, s_t, s_j};
  launch_k(mpjpe_nd_bwd_kernel, dim3(loss_grid(n_pts * 4, sm_count)), dim3(kLossThreads), 0, stream, pred, tgt, grad_out,
                                                                                   1.f / (float)n_pts, n_pts, D, wv,
                                                                                   grad_pred);
  return cudaGetLastError();
}

cudaError_t launch_n_mpjpe_fwd(const float* pred, const float* tgt, long long n_poses, int J, double* partial,
                               float* out, int sm_count, cudaStream_t stream) {
  long long blocks = (n_poses * 32 + kLossThreads - 1) / kLossThreads;
  const long long cap = (long long)sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  launch_k(n_mpjpe_partial_kernel, dim3((int)blocks), dim3(kLossThreads), 0, stream, pred, tgt, n_poses, J, partial);
  const double cnt = (double)n_poses * (double)J;
  launch_k(mean_finish_kernel, dim3(1), dim3(kLossThreads), 0, stream, partial, (int)blocks, cnt > 0 ? 1.0 / cnt : 0.0, out);
  return cudaGetLastError();
}

cudaError_t launch_p_mpjpe_fwd(const float* pred, const float* tgt, long long n_poses, int J, double* partial, float* out,
                               int sm_count, cudaStream_t stream) {
  long long blocks = (n_poses + kLossThreads - 1) / kLossThreads;
  const long long cap = (long long)sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  launch_k(p_mpjpe_partial_kernel, dim3((int)blocks), dim3(kLossThreads), 0, stream, pred, tgt, n_poses, J, partial);
  launch_k(mean_finish_kernel, dim3(1), dim3(kLossThreads), 0, stream, partial, (int)blocks, 1.0 / ((double)n_poses * J), out);
  return cudaGetLastError();
}

cudaError_t launch_velocity_error(const float* pred, const float* tgt, long long T, long long inner, int D,
                                  double* partial, float* out, int sm_count, cudaStream_t stream) {
  const long long n = (T - 1) * inner;
  const int grid = loss_grid(n * 4, sm_count);
  launch_k(velocity_partial_kernel, dim3(grid), dim3(kLossThreads), 0, stream, pred, tgt, T, inner, D, partial);
  launch_k(mean_finish_kernel, dim3(1), dim3(kLossThreads), 0, stream, partial, grid, n > 0 ? 1.0 / (double)n : 0.0, out);
  return cudaGetLastError();
}

cudaError_t launch_n_mpjpe_bwd(const float* pred, const float* tgt, const float* grad_out, long long n_poses, int J,
                               float* grad_pred, int sm_count, cudaStream_t stream) {
  long long blocks = (n_poses * 32 + kLossThreads - 1) / kLossThreads;
  const long long cap = (long long)sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const double cnt = (double)n_poses * (double)J;
  launch_k(n_mpjpe_bwd_kernel, dim3((int)blocks), dim3(kLossThreads), 0, stream, pred, tgt, grad_out, (float)(1.0 / cnt), n_poses, J,
                                                               grad_pred);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Fused reprojection loss: mpjpe(project_to_2d(pose + trajectory, camera_params), target_2d) in ONE pass over the points
// (common/camera.py:37-67 + common/loss.py:11-17; the reprojection term of upstream VideoPose3D's semi-supervised
// step, whose primitives are all this fork keeps). Unfused it is an add, the projection kernel, the 2-D error
// reduction and -- backwards -- the error gradient, the projection backward and the broadcast-sum of the trajectory
// gradient: 20 B / point written and read again between the kernels. Here the forward reads 12 B (pose) + 8 B (target)
// per point, the backward the same and writes 12 B.
struct ReprojArgs {
  const float* pose;     // [n_pts][3] camera-space points
  const float* traj;     // optional [n_pts / pts_per_traj][3], added to every point of its group
  const float* cam;      // [n_pts / pts_per_cam][9]
  const float* tgt;      // [n_pts][2]
  long long n_pts, pts_per_traj, pts_per_cam;
  int linear;
};

struct Reproj {
  float u, v;              // projected point
  float rx, ry, xx, yy;    // unclamped and clamped normalised coordinates
  float x, y, z;
};
__device__ __forceinline__ float clamp_unit_keep_nan(float x) {
  if (x != x) return x;
  return fminf(fmaxf(x, -1.f), 1.f);
}
__device__ __forceinline__ Reproj reproject(const ReprojArgs& a, long long i, const float*& c) {
  Reproj r;
  r.x = a.pose[3 * i];
  r.y = a.pose[3 * i + 1];
  r.z = a.pose[3 * i + 2];
  if (a.traj != nullptr) {
    const float* t = a.traj + 3 * (i / a.pts_per_traj);
    r.x += __ldg(t);
    r.y += __ldg(t + 1);
    r.z += __ldg(t + 2);
  }
  c = a.cam + 9 * (i / a.pts_per_cam);
  r.rx = r.x / r.z;
  r.ry = r.y / r.z;
  r.xx = clamp_unit_keep_nan(r.rx);
  r.yy = clamp_unit_keep_nan(r.ry);
  const float fx = __ldg(c), fy = __ldg(c + 1), cx = __ldg(c + 2), cy = __ldg(c + 3);
  if (a.linear) {
    r.u = fx * r.xx + cx;
    r.v = fy * r.yy + cy;
  } else {
    const float k1 = __ldg(c + 4), k2 = __ldg(c + 5), k3 = __ldg(c + 6), p1 = __ldg(c + 7), p2 = __ldg(c + 8);
    const float r2 = r.xx * r.xx + r.yy * r.yy;
    const float s = 1.f + r2 * (k1 + r2 * (k2 + r2 * k3)) + p1 * r.xx + p2 * r.yy;
    r.u = fx * (r.xx * s + p1 * r2) + cx;
    r.v = fy * (r.yy * s + p2 * r2) + cy;
  }
  return r;
}

__global__ void __launch_bounds__(kLossThreads)
reproj_partial_kernel(const ReprojArgs a, double* __restrict__ partial) {
  pdl_enter();
  const long long stride = (long long)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_pts; i += stride) {
    const float* c;
    const Reproj r = reproject(a, i, c);
    const float2 t = __ldg(reinterpret_cast<const float2*>(a.tgt) + i);
    const float du = r.u - t.x, dv = r.v - t.y;
    acc += (double)sqrtf(du * du + dv * dv);
  }
  const double s = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// grad_pose[i] = d loss / d pose[i]; grad_traj (zero on entry) += the same summed over the points of its group
__global__ void __launch_bounds__(kLossThreads)
reproj_bwd_kernel(const ReprojArgs a, const float* __restrict__ grad_out, float inv_count, float* __restrict__ grad_pose,
                  float* __restrict__ grad_traj) {
  pdl_enter();
  const float g0 = __ldg(grad_out) * inv_count;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_pts; i += stride) {
    const float* c;
    const Reproj r = reproject(a, i, c);
    const float2 t = __ldg(reinterpret_cast<const float2*>(a.tgt) + i);
    const float du = r.u - t.x, dv = r.v - t.y;
    const float nrm = sqrtf(du * du + dv * dv);
    const float k = nrm > 0.f ? g0 / nrm : 0.f;
    // chain of project_bwd_kernel (projection.cu) with g = (du, dv) * k
    const float ga = __ldg(c) * du * k, gb = __ldg(c + 1) * dv * k;
    float gxx = ga, gyy = gb;
    if (!a.linear) {
      const float k1 = __ldg(c + 4), k2 = __ldg(c + 5), k3 = __ldg(c + 6), p1 = __ldg(c + 7), p2 = __ldg(c + 8);
      const float r2 = r.xx * r.xx + r.yy * r.yy;
      const float s = 1.f + r2 * (k1 + r2 * (k2 + r2 * k3)) + p1 * r.xx + p2 * r.yy;
      const float gs = ga * r.xx + gb * r.yy;
      const float gr2 = ga * p1 + gb * p2 + gs * (k1 + r2 * (2.f * k2 + 3.f * k3 * r2));
      gxx = ga * s + gs * p1 + gr2 * 2.f * r.xx;
      gyy = gb * s + gs * p2 + gr2 * 2.f * r.yy;
    }
    if (!(r.rx >= -1.f && r.rx <= 1.f)) gxx = 0.f;   // torch.clamp backward
    if (!(r.ry >= -1.f && r.ry <= 1.f)) gyy = 0.f;
    const float iz = 1.f / r.z;
    const float gx = gxx * iz, gy = gyy * iz, gz = -(gxx * r.x + gyy * r.y) * iz * iz;
    if (grad_pose != nullptr) {
      grad_pose[3 * i] = gx;
      grad_pose[3 * i + 1] = gy;
      grad_pose[3 * i + 2] = gz;
    }
    if (grad_traj != nullptr) {
      float* gt = grad_traj + 3 * (i / a.pts_per_traj);
      atomicAdd(gt, gx);
      atomicAdd(gt + 1, gy);
      atomicAdd(gt + 2, gz);
    }
  }
}

cudaError_t launch_reproj_fwd(const float* pose, const float* traj, long long n_pts, long long pts_per_traj,
                              const float* cam, long long pts_per_cam, int linear, const float* tgt, double* partial,
                              float* out, int sm_count, cudaStream_t stream) {
  const ReprojArgs a{pose, traj, cam, tgt, n_pts, pts_per_traj > 0 ? pts_per_traj : 1, pts_per_cam > 0 ? pts_per_cam : 1,
                     linear};
  const int grid = loss_grid(n_pts * 4, sm_count);
  launch_k(reproj_partial_kernel, dim3(grid), dim3(kLossThreads), 0, stream, a, partial);
  launch_k(mean_finish_kernel, dim3(1), dim3(kLossThreads), 0, stream, partial, grid, n_pts > 0 ? 1.0 / (double)n_pts : 0.0, out);
  return cudaGetLastError();
}

cudaError_t launch_reproj_bwd(const float* pose, const float* traj, long long n_pts, long long pts_per_traj,
                              const float* cam, long long pts_per_cam, int linear, const float* tgt,
                              const float* grad_out, float* grad_pose, float* grad_traj, int sm_count,
                              cudaStream_t stream) {
  const ReprojArgs a{pose, traj, cam, tgt, n_pts, pts_per_traj > 0 ? pts_per_traj : 1, pts_per_cam > 0 ? pts_per_cam : 1,
                     linear};
  const int grid = loss_grid(n_pts * 4, sm_count);
  launch_k(reproj_bwd_kernel, dim3(grid), dim3(kLossThreads), 0, stream, a, grad_out, 1.f / (float)n_pts, grad_pose, grad_traj);
  return cudaGetLastError();
}

}  // namespace vp3d
