// Gradient exchange of data-parallel training over NVLink / NVSwitch peer memory (the reference is a single
// process and has no counterpart). Every rank maps every rank's exchange buffer (CUDA VMM handles exchanged once by the
// host side; with NVSwitch also ONE multicast address that stands for all replicas). An all-reduce of a slice
// [off, off + n) of that buffer is a two-shot exchange in ONE small kernel:
//
//   barrier over the ranks (flag words in peer memory)     every rank's slice is complete and visible
//   rank r owns 1 / world of the slice:
//     multicast:  v = multimem.ld_reduce.add [mc + i]       the SWITCH adds the replicas, one 16-byte answer returns
//                 multimem.st [mc + i], v * scale           and the switch writes the result into every replica
//     peer ptrs:  v = sum over ranks of ld [peer_k + i]     (no multicast object: plain P2P loads / stores)
//                 st [peer_k + i], v * scale  for every k
//   barrier over the ranks                                  every replica holds the complete result
//
// The kernel is a handful of CTAs (clusters of two, so that they occupy whole TPCs and leave the others to the CTA-pair
// GEMMs) with no shared memory. The persistent GEMMs of the backward size their grids for that many SMs fewer
// (vp3d_set_sm_limit), so the exchange of layer L runs BESIDE the GEMMs of layer L - 1 instead of taking turns with them
// -- which is what a library all-reduce with its shared-memory-hungry CTAs does to persistent 226 KB kernels
// (DESIGN.md section 5). One rank computes each element and every rank receives the same bits, so replicas stay
// bit-identical.
#include "kernels.h"
#include "pdl.cuh"
#include "ptx.cuh"

namespace vp3d {
namespace {

constexpr int kArThreads = 512;
constexpr int kArUnroll = 8;        // 16-byte requests in flight per thread (64 KB per CTA)

__device__ __forceinline__ uint32_t cas_acq_rel_sys(uint32_t* addr, uint32_t expect, uint32_t desired) {
  uint32_t old;
  asm volatile("atom.global.acq_rel.sys.cas.b32 %0, [%1], %2, %3;"
               : "=r"(old)
               : "l"(addr), "r"(expect), "r"(desired)
               : "memory");
  return old;
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Flag word [block][source rank] in every rank's flag area. A sender flips the receiver's word 0 -> 1 (spinning while
// the previous signal has not been consumed), the receiver flips its own word 1 -> 0: self-resetting, so consecutive
// barriers (and consecutive launches on one stream, and replays of a captured graph) share the words. Bounded: a rank
// that never arrives traps this kernel after `timeout_ns` instead of hanging the GPU.
__device__ void rank_barrier(const AllReduceParams& p) {
  __syncthreads();
  if (threadIdx.x < p.world) {
    const int peer = threadIdx.x;
    const unsigned long long t0 = global_ns();
    uint32_t* theirs = p.flags[peer] + blockIdx.x * p.world + p.rank;
    while (cas_acq_rel_sys(theirs, 0u, 1u) != 0u) {
      if (global_ns() - t0 > p.timeout_ns) {
        printf("vp3d: peer barrier timed out sending (rank %d -> %d, block %d)\n", p.rank, peer, blockIdx.x);
        __trap();
      }
    }
    uint32_t* mine = p.flags[p.rank] + blockIdx.x * p.world + peer;
    while (cas_acq_rel_sys(mine, 1u, 0u) != 1u) {
      if (global_ns() - t0 > p.timeout_ns) {
        printf("vp3d: peer barrier timed out waiting (rank %d <- %d, block %d)\n", p.rank, peer, blockIdx.x);
        __trap();
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float4 mc_ld_reduce(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 peer_ld(const float* ptr) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(ptr)
               : "memory");
  return v;
}
__device__ __forceinline__ void peer_st(float* ptr, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

template <bool MC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kArThreads, 1)
peer_allreduce_kernel(const AllReduceParams p) {
  // Programmatic dependent launch: wait for the producer of the slice, but do NOT release the dependents early. This
  // kernel blocks on OTHER ranks; a dependent released at its start -- the optimiser update of the slice, ~1200 blocks
  // -- would sit resident in griddepcontrol.wait on every SM for as long as the slowest rank takes, and the backward's
  // persistent GEMMs (one CTA per SM, 59 k registers) could not be scheduled beside it (see pdl_enter_long in pdl.cuh for
  // the measurement that located the effect in the GEMMs' own dependents).
  pdl_wait();
  rank_barrier(p);
  const long long vecs = p.n >> 2;
  const long long lo = vecs * p.rank / p.world, hi = vecs * (p.rank + 1) / p.world;
  const long long stride = static_cast<long long>(gridDim.x) * kArThreads;
  for (long long i0 = lo + static_cast<long long>(blockIdx.x) * kArThreads + threadIdx.x; i0 < hi;
       i0 += stride * kArUnroll) {
    float4 v[kArUnroll];
    if (MC) {
#pragma unroll
      for (int u = 0; u < kArUnroll; ++u) {
        const long long i = i0 + u * stride;
        if (i < hi) v[u] = mc_ld_reduce(p.mc + p.off + 4 * i);
      }
    } else {
#pragma unroll
      for (int u = 0; u < kArUnroll; ++u) v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < p.world; ++k) {     // fixed order; one rank computes an element, all ranks receive its bits
        const float* src = p.peers[k] + p.off;
#pragma unroll
        for (int u = 0; u < kArUnroll; ++u) {
          const long long i = i0 + u * stride;
          if (i < hi) {
            const float4 t = peer_ld(src + 4 * i);
            v[u].x += t.x; v[u].y += t.y; v[u].z += t.z; v[u].w += t.w;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kArUnroll; ++u) {
      const long long i = i0 + u * stride;
      if (i >= hi) continue;
      const float4 r = make_float4(v[u].x * p.scale, v[u].y * p.scale, v[u].z * p.scale, v[u].w * p.scale);
      if (MC) {
        mc_st(p.mc + p.off + 4 * i, r);
      } else {
        for (int k = 0; k < p.world; ++k) peer_st(p.peers[k] + p.off + 4 * i, r);
      }
    }
  }
  rank_barrier(p);
}

}  // namespace

cudaError_t launch_peer_allreduce(const AllReduceParams& p, int ctas, cudaStream_t stream) {
  if (p.mc != nullptr)
    launch_k(peer_allreduce_kernel<true>, dim3(ctas), dim3(kArThreads), 0, stream, p);
  else
    launch_k(peer_allreduce_kernel<false>, dim3(ctas), dim3(kArThreads), 0, stream, p);
  return cudaGetLastError();
}

}  // namespace vp3d
