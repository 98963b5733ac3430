// Programmatic dependent launch (PDL): every kernel of the library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the next kernel on the stream may become resident -- and run its
// prologue: barrier init, tensor-memory allocation, descriptor prefetch -- while this one drains its last CTAs, instead
// of being launched after the last CTA has retired. Each kernel therefore
//   * calls pdl_wait() (griddepcontrol.wait: the prerequisite grids have completed and their memory is visible) before
//     its first access to global memory, in EVERY thread and before any early return -- a grid none of whose threads
//     waited could complete before its predecessor and let ITS successor start too early;
//   * calls pdl_trigger() (griddepcontrol.launch_dependents) right after: the dependent grid may be scheduled as soon as
//     every CTA of this one has started (it still waits for this grid's completion in its own pdl_wait()).
// A step of the training loop is ~80 dependent launches of 3-150 us, most of them back to back on one stream (also as
// programmatic edges of the captured CUDA graph). vp3d_set_pdl(0) / VP3D_PDL=0 launches everything with the attribute
// off (plain stream order); the device-side instructions are then no-ops.
#pragma once

#include <cuda_runtime.h>

namespace vp3d {

extern int g_pdl;   // api.cu

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_wait();
  pdl_trigger();
}
// Long-running persistent kernels (LEVEL 1: the weight-gradient GEMM, 2: also the convolution GEMMs) do NOT release their
// dependents early. A dependent released at their start sits resident in griddepcontrol.wait for the whole launch; when it
// is a big grid of small blocks (the weight-gradient layout pass, Adam, a BatchNorm pass) it holds registers and thread
// slots that kernels of OTHER streams need -- the exchange kernel and the optimiser updates of a data-parallel backward,
// the BatchNorm passes meant to run beside the weight-gradient GEMM. Measured with A/B builds (VP3D_PDL_LATE_TRIGGER = 0 /
// 1 / 2, same box): one GPU 1.700 / 1.698 / 1.698 ms per training step (the gain of PDL comes from the chains of small
// kernels), two GPUs 1.971 / 1.896 / 1.808 ms. Level 3 (the HBM-bound passes -- BatchNorm, weight-gradient layout, Adam --
// keep their dependents back as well) loses that gain: 1.723-1.732 against 1.685-1.698 ms on one GPU. Default 2.
#ifndef VP3D_PDL_LATE_TRIGGER
#define VP3D_PDL_LATE_TRIGGER 2
#endif
// `short_launch`: at most one tile per CTA (the small layers, a streamed frame) -- nothing stays parked for long, and the
// dependent's early start is worth its ~1 us per boundary there (streaming GEMM path: 0.121 against 0.126 ms per frame).
// Tail trigger (A/B builds, VP3D_PDL_TAIL_TRIGGER=1): a long launch releases its dependents when a CTA's roles run out of
// work -- the producer warp first, after its last load -- so that the dependent's launch latency overlaps the last tiles'
// MMAs and epilogue instead of following the last CTA's exit, and nothing is parked for longer than about one tile.
// Measured on one GPU: training step 1.685-1.695 ms with, 1.673-1.684 without; inference 7.63 / 7.63 ms -- off.
#ifndef VP3D_PDL_TAIL_TRIGGER
#define VP3D_PDL_TAIL_TRIGGER 0
#endif
__device__ __forceinline__ void pdl_tail_trigger(bool short_launch) {
  if (VP3D_PDL_TAIL_TRIGGER && VP3D_PDL_LATE_TRIGGER > 0 && !short_launch) pdl_trigger();
}
template <int LEVEL>
__device__ __forceinline__ void pdl_enter_long(bool short_launch) {
  pdl_wait();
  if (VP3D_PDL_LATE_TRIGGER < LEVEL || short_launch) pdl_trigger();
}

// kernel<<<grid, block, smem, stream>>>(args...) with the PDL attribute (and optionally a cluster shape)
template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace vp3d
