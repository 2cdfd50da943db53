// ctk_thread.cu -- the thread-per-cluster refine kernel (ctk_thread.cuh) and its launcher.
#include <cuda_runtime.h>
#include <stdio.h>

#include "ctk_thread.cuh"

namespace ctk {


template <class Real>
__global__ void __launch_bounds__(CTK_T_BLOCK) refine_thread_kernel(const BatchArgs a) {
  int n_work = a.n_work;
  if (a.n_work_dev) n_work = min(n_work, *a.n_work_dev);
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ Real constants[6 * CTK_T_NMAX * CTK_T_BLOCK];   // per-feature constants, one column per thread
  if (w >= n_work) return;
  ThreadSolver<Real> solver(a, constants + threadIdx.x);
  solver.run(a.work_ids ? a.work_ids[w] : w);
}

int launch_refine_threads(const BatchArgs& args, cudaStream_t stream, char* err, size_t err_len) {
  const int grid = (args.n_work + CTK_T_BLOCK - 1) / CTK_T_BLOCK;
  if (grid < 1) return 0;
  if (args.prob.compute_dtype == CTK_COMPUTE_F64)
    refine_thread_kernel<double><<<grid, CTK_T_BLOCK, 0, stream>>>(args);
  else
    refine_thread_kernel<float><<<grid, CTK_T_BLOCK, 0, stream>>>(args);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(err, err_len, "refine_thread_kernel launch: %s", cudaGetErrorString(e));
    return CTK_E_CUDA;
  }
  return 0;
}

}  // namespace ctk
