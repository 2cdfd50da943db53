// ctk_kernel.cuh -- the refine kernel and its launcher (definitions; included by ctk_inst.cu only).
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

#include "ctk_launch.h"

namespace ctk {

// Persistent warps, one cluster per warp at a time; clusters are handed out by an atomic counter so
// that clusters of different cost do not leave the other warps of a block idle.
#ifndef CTK_BLOCK_WARPS
#define CTK_BLOCK_WARPS 4      // warps (clusters in flight) per block
#endif
#ifndef CTK_MIN_BLOCKS
#define CTK_MIN_BLOCKS 4       // resident blocks per SM the register budget is sized for
#endif

template <class C>
__global__ void __launch_bounds__(32 * CTK_BLOCK_WARPS, CTK_MIN_BLOCKS)
refine_kernel(const BatchArgs a) {
  ClusterSolver<C> solver(a, (uint32_t) (threadIdx.x >> 5) * (uint32_t) a.lay.total);
  int n_work = a.n_work;
  if (a.n_work_dev) n_work = min(n_work, *a.n_work_dev);   // e.g. the overflow list of a previous launch
  for (;;) {
    int w = 0;
    if ((threadIdx.x & 31) == 0) w = atomicAdd(a.counter, 1);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w >= n_work) break;
    solver.run(a.work_ids ? a.work_ids[w] : w);
  }
}

template <class C>
__global__ void __launch_bounds__(32 * CTK_BLOCK_WARPS, CTK_MIN_BLOCKS)
global_kernel(const BatchArgs a) {
  ClusterSolver<C> solver(a, (uint32_t) (threadIdx.x >> 5) * (uint32_t) a.lay.total);
  for (;;) {
    int w = 0;
    if ((threadIdx.x & 31) == 0) w = atomicAdd(a.counter, 1);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w >= a.n_work) break;
    solver.run_global(w);
  }
}

#define CTK_LAUNCH_CUDA(call)                                                                \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      snprintf(err, err_len, "%s: %s", #call, cudaGetErrorString(e_));                       \
      return CTK_E_CUDA;                                                                     \
    }                                                                                        \
  } while (0)

template <class C>
int launch_refine(const BatchArgs& args, cudaStream_t stream, char* err, size_t err_len) {
  int dev = 0, sms = 0, smem_max = 0;
  CTK_LAUNCH_CUDA(cudaGetDevice(&dev));
  CTK_LAUNCH_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CTK_LAUNCH_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  auto kernel_big = refine_kernel<C>;
  if (C::BIG) {
    // large clusters: one warp per block, arrays in the global workspace (args.big_blocks slices)
    int grid = args.big_blocks < args.n_work ? args.big_blocks : args.n_work;
    if (grid < 1) grid = 1;
    kernel_big<<<grid, 32, 0, stream>>>(args);
    CTK_LAUNCH_CUDA(cudaGetLastError());
    return 0;
  }
  const int per_warp = args.lay.total;
  if (per_warp > smem_max) {
    snprintf(err, err_len, "a cluster of %d features needs %d B of shared memory (limit %d B)",
             args.lay.n_max, per_warp, smem_max);
    return CTK_E_CAPACITY;
  }
  int warps = CTK_BLOCK_WARPS;
  while (warps > 1 && warps * per_warp > smem_max / 2) warps >>= 1;   // keep >= 2 blocks per SM
  while (warps > 1 && warps * per_warp > smem_max) warps >>= 1;
  const int smem = warps * per_warp;
  auto kernel = refine_kernel<C>;
  CTK_LAUNCH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int per_sm = 0;
  CTK_LAUNCH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem));
  if (per_sm < 1) {
    snprintf(err, err_len, "refine kernel does not fit on an SM (%d B shared)", smem);
    return CTK_E_CAPACITY;
  }
  int grid = per_sm * sms;
  const int need = (args.n_work + warps - 1) / warps;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  kernel<<<grid, warps * 32, smem, stream>>>(args);
  CTK_LAUNCH_CUDA(cudaGetLastError());
  return 0;
}

template <class C>
int launch_global(const BatchArgs& args, cudaStream_t stream, char* err, size_t err_len) {
  int dev = 0, sms = 0, smem_max = 0;
  CTK_LAUNCH_CUDA(cudaGetDevice(&dev));
  CTK_LAUNCH_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CTK_LAUNCH_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const int per_warp = args.lay.total;
  if (per_warp > smem_max) {
    snprintf(err, err_len, "a cluster of %d features needs %d B of shared memory (limit %d B)",
             args.lay.n_max, per_warp, smem_max);
    return CTK_E_CAPACITY;
  }
  int warps = CTK_BLOCK_WARPS;
  while (warps > 1 && warps * per_warp > smem_max) warps >>= 1;
  const int smem = warps * per_warp;
  auto kernel = global_kernel<C>;
  CTK_LAUNCH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int per_sm = 0;
  CTK_LAUNCH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem));
  if (per_sm < 1) {
    snprintf(err, err_len, "global kernel does not fit on an SM (%d B shared)", smem);
    return CTK_E_CAPACITY;
  }
  int grid = per_sm * sms;
  const int need = (args.n_work + warps - 1) / warps;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  kernel<<<grid, warps * 32, smem, stream>>>(args);
  CTK_LAUNCH_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ctk
