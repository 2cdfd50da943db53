// ctk_api.cu -- kernels and C ABI of libctk (include/ctk.h).  sm_100a only.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <stdlib.h>

#include "ctk_launch.h"

namespace {

thread_local char g_error[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

#define CTK_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      return fail(CTK_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_));                      \
  } while (0)

struct GlobalLauncher {
  const ctk::BatchArgs* args;
  cudaStream_t stream;
  int result;
  template <class C> void operator()() {
    if constexpr (C::EXTRA && !C::BIG)
      result = ctk::launch_global<C>(*args, stream, g_error, sizeof(g_error));
    else
      result = CTK_E_UNSUPPORTED;
  }
};

struct Launcher {
  const ctk::BatchArgs* args;
  cudaStream_t stream;
  int result;
  template <class C> void operator()() {
    result = ctk::launch_refine<C>(*args, stream, g_error, sizeof(g_error));
  }
};

// ------------------------------------------------------------------------------------------------
// frame maximum: one block per frame, 16-byte loads, HBM-bound
// ------------------------------------------------------------------------------------------------
template <class T> __device__ __forceinline__ double to_double(T v) { return (double) v; }

template <class T>
__global__ void __launch_bounds__(1024) frame_max_kernel(const void* const* frames, int64_t n_pixels,
                                                         double* out) {
  const T* src = reinterpret_cast<const T*>(frames[blockIdx.x]);
  constexpr int VEC = 16 / sizeof(T);
  double best = -INFINITY;
  const int64_t n_vec = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) ? n_pixels / VEC : 0;
  const uint4* v = reinterpret_cast<const uint4*>(src);
  for (int64_t i = threadIdx.x; i < n_vec; i += blockDim.x) {
    uint4 raw = __ldg(v + i);
    const T* e = reinterpret_cast<const T*>(&raw);
    T m = e[0];
#pragma unroll
    for (int k = 1; k < VEC; ++k) m = e[k] > m ? e[k] : m;
    best = fmax(best, to_double(m));
  }
  for (int64_t i = n_vec * VEC + threadIdx.x; i < n_pixels; i += blockDim.x)
    best = fmax(best, to_double(src[i]));
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, m));
  __shared__ double part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x < 32) {
    best = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : -INFINITY;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, m));
    if (threadIdx.x == 0) out[blockIdx.x] = best;
  }
}

}  // namespace

extern "C" {

int ctk_version(void) { return CTK_VERSION; }
size_t ctk_problem_bytes(void) { return sizeof(ctk_problem_t); }

const char* ctk_last_error(void) { return g_error; }

int ctk_frame_max(const void* const* d_frames, int32_t n_frames, int64_t n_pixels,
                  int32_t pixel_dtype, double* d_max_out, void* stream) {
  if (n_frames <= 0) return 0;
  if (!d_frames || !d_max_out || n_pixels <= 0) return fail(CTK_E_INVALID, "ctk_frame_max: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (pixel_dtype) {
    case CTK_PIXEL_U8: frame_max_kernel<uint8_t><<<n_frames, 1024, 0, st>>>(d_frames, n_pixels, d_max_out); break;
    case CTK_PIXEL_U16: frame_max_kernel<uint16_t><<<n_frames, 1024, 0, st>>>(d_frames, n_pixels, d_max_out); break;
    case CTK_PIXEL_F32: frame_max_kernel<float><<<n_frames, 1024, 0, st>>>(d_frames, n_pixels, d_max_out); break;
    case CTK_PIXEL_F64: frame_max_kernel<double><<<n_frames, 1024, 0, st>>>(d_frames, n_pixels, d_max_out); break;
    case CTK_PIXEL_I16: frame_max_kernel<int16_t><<<n_frames, 1024, 0, st>>>(d_frames, n_pixels, d_max_out); break;
    case CTK_PIXEL_I32: frame_max_kernel<int32_t><<<n_frames, 1024, 0, st>>>(d_frames, n_pixels, d_max_out); break;
    default: return fail(CTK_E_INVALID, "ctk_frame_max: unknown pixel dtype %d", pixel_dtype);
  }
  CTK_CUDA(cudaGetLastError());
  return 0;
}

size_t ctk_refine_workspace_bytes(void) { return 256; }

namespace {
// number of concurrently processed large clusters: bounded by 2 GiB of workspace
int big_block_count(const ctk::Layout& lay) {
  long long n = (2LL << 30) / (lay.total > 0 ? lay.total : 1);
  return (int) (n < 4 ? 4 : (n > 148 ? 148 : n));
}
}  // namespace

size_t ctk_refine_workspace_bytes_for(const ctk_problem_t* prob, int32_t max_cluster_features) {
  if (!prob || ctk::validate_problem(*prob)) return 0;
  ctk::Layout lay;
  if (!ctk::compute_layout(*prob, max_cluster_features, &lay)) return 0;
  if (max_cluster_features <= CTK_MAX_CLUSTER_FEATURES) return 256;
  return 256 + (size_t) big_block_count(lay) * (size_t) lay.total;
}

size_t ctk_refine_shared_bytes(const ctk_problem_t* prob, int32_t max_cluster_features) {
  if (!prob || ctk::validate_problem(*prob)) return 0;
  ctk::Layout lay;
  if (!ctk::compute_layout(*prob, max_cluster_features, &lay)) return 0;
  if (max_cluster_features > CTK_MAX_CLUSTER_FEATURES) return 0;   // no shared memory: workspace
  return lay.total <= 227 * 1024 ? (size_t) lay.total : 0;
}

int ctk_refine_batch(const ctk_problem_t* prob, const void* const* d_frames,
                     const int64_t* frame_shape, const double* d_frame_max, int32_t n_work,
                     const int32_t* d_work_ids, int32_t max_cluster_features,
                     const int32_t* d_cluster_frame, const int32_t* d_cluster_offset,
                     const double* d_params_in, const double* d_bounds_lo, const double* d_bounds_hi,
                     double* d_params_out, double* d_cost_out, int32_t* d_status_out,
                     int32_t* d_stats_out, void* d_workspace, void* stream) {
  return ctk_refine_batch_chained(prob, d_frames, frame_shape, d_frame_max, n_work, d_work_ids,
                                  max_cluster_features, d_cluster_frame, d_cluster_offset,
                                  d_params_in, d_bounds_lo, d_bounds_hi, d_params_out, d_cost_out,
                                  d_status_out, d_stats_out, d_workspace, nullptr, nullptr, 0, stream);
}

int ctk_refine_batch_chained(const ctk_problem_t* prob, const void* const* d_frames,
                             const int64_t* frame_shape, const double* d_frame_max, int32_t n_work,
                             const int32_t* d_work_ids, int32_t max_cluster_features,
                             const int32_t* d_cluster_frame, const int32_t* d_cluster_offset,
                             const double* d_params_in, const double* d_bounds_lo,
                             const double* d_bounds_hi, double* d_params_out, double* d_cost_out,
                             int32_t* d_status_out, int32_t* d_stats_out, void* d_workspace,
                             const int32_t* d_n_work, int32_t* d_overflow,
                             int32_t overflow_capacity, void* stream) {
  return ctk_refine_batch_ex(prob, d_frames, frame_shape, d_frame_max, n_work, d_work_ids,
                             max_cluster_features, d_cluster_frame, d_cluster_offset, d_params_in,
                             d_bounds_lo, d_bounds_hi, d_params_out, d_cost_out, d_status_out,
                             d_stats_out, d_workspace, d_n_work, d_overflow, overflow_capacity, 0,
                             stream);
}

int ctk_refine_batch_ex(const ctk_problem_t* prob, const void* const* d_frames,
                             const int64_t* frame_shape, const double* d_frame_max, int32_t n_work,
                             const int32_t* d_work_ids, int32_t max_cluster_features,
                             const int32_t* d_cluster_frame, const int32_t* d_cluster_offset,
                             const double* d_params_in, const double* d_bounds_lo,
                             const double* d_bounds_hi, double* d_params_out, double* d_cost_out,
                             int32_t* d_status_out, int32_t* d_stats_out, void* d_workspace,
                             const int32_t* d_n_work, int32_t* d_overflow,
                             int32_t overflow_capacity, int32_t flags, void* stream) {
  if (!prob) return fail(CTK_E_INVALID, "ctk_refine_batch: prob is NULL");
  if (const char* why = ctk::validate_problem(*prob)) return fail(CTK_E_INVALID, "ctk_refine_batch: %s", why);
  if (n_work < 0) return fail(CTK_E_INVALID, "ctk_refine_batch: n_work < 0");
  if (n_work == 0) return 0;
  if (!d_frames || !frame_shape || !d_frame_max || !d_cluster_frame || !d_cluster_offset ||
      !d_params_in || (!d_bounds_lo != !d_bounds_hi) || !d_params_out || !d_cost_out ||
      !d_status_out || !d_stats_out || !d_workspace)
    return fail(CTK_E_INVALID, "ctk_refine_batch: NULL pointer argument");
  ctk::BatchArgs a;
  memset(&a, 0, sizeof(a));
  a.prob = *prob;
  a.frames = d_frames;
  for (int k = 0; k < prob->ndim; ++k) {
    if (frame_shape[k] <= 0 || frame_shape[k] > (1 << 30))
      return fail(CTK_E_INVALID, "ctk_refine_batch: bad frame shape");
    a.shape[k] = frame_shape[k];
  }
  a.frame_max = d_frame_max;
  a.n_work = n_work;
  a.work_ids = d_work_ids;
  a.cluster_frame = d_cluster_frame;
  a.cluster_offset = d_cluster_offset;
  a.params_in = d_params_in;
  a.lo_in = d_bounds_lo;
  a.hi_in = d_bounds_hi;
  a.params_out = d_params_out;
  a.cost_out = d_cost_out;
  a.status_out = d_status_out;
  a.stats_out = d_stats_out;
  a.counter = static_cast<int32_t*>(d_workspace);
  if (overflow_capacity < 0) return fail(CTK_E_INVALID, "ctk_refine_batch: overflow_capacity < 0");
  a.n_work_dev = d_n_work;
  a.overflow = overflow_capacity > 0 ? d_overflow : nullptr;
  a.overflow_cap = overflow_capacity;
  if (!ctk::compute_layout(*prob, max_cluster_features, &a.lay))
    return fail(CTK_E_CAPACITY, "ctk_refine_batch: max_cluster_features %d out of range [1, %d]",
                max_cluster_features, CTK_MAX_BIG_FEATURES);
  const bool big = max_cluster_features > CTK_MAX_CLUSTER_FEATURES;
  if (big) {
    a.big_workspace = static_cast<char*>(d_workspace) + 256;
    a.big_blocks = big_block_count(a.lay);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CTK_CUDA(cudaMemsetAsync(d_workspace, 0, sizeof(int32_t), st));
  if (a.overflow && !(flags & CTK_LAUNCH_APPEND_OVERFLOW))
    CTK_CUDA(cudaMemsetAsync(a.overflow, 0, sizeof(int32_t), st));
  Launcher launcher{&a, st, 0};
  bool found = prob->compute_dtype == CTK_COMPUTE_F64
                   ? ctk::dispatch_config<double>(*prob, launcher, big)
                   : ctk::dispatch_config<float>(*prob, launcher, big);
  if (!found) return fail(CTK_E_UNSUPPORTED, "ctk_refine_batch: no kernel instance for this problem");
  return launcher.result;
}

int ctk_global_pass(const ctk_problem_t* prob, const void* const* d_frames,
                    const int64_t* frame_shape, double norm, int32_t n_clusters,
                    int32_t max_cluster_features, const int32_t* d_cluster_frame,
                    const int32_t* d_cluster_offset, const double* d_params_in,
                    const double* d_mask_centres, int32_t phase, double lambda, int32_t use_newton,
                    const double* d_global_step, double* d_params_out, double* d_accum,
                    double* d_cost_out, int32_t* d_status_out, void* d_workspace, void* stream) {
  if (!prob) return fail(CTK_E_INVALID, "ctk_global_pass: prob is NULL");
  if (const char* why = ctk::validate_problem(*prob, true)) return fail(CTK_E_INVALID, "ctk_global_pass: %s", why);
  if (prob->constraint_mask) return fail(CTK_E_UNSUPPORTED, "ctk_global_pass: constraints are not supported");
  if (n_clusters <= 0) return 0;
  if (!d_frames || !frame_shape || !d_cluster_frame || !d_cluster_offset || !d_params_in ||
      !d_params_out || !d_accum || !d_cost_out || !d_status_out || !d_workspace ||
      (phase != 1 && phase != 2) || (phase == 2 && !d_global_step) || !(norm > 0.) ||
      max_cluster_features > CTK_MAX_CLUSTER_FEATURES)
    return fail(CTK_E_INVALID, "ctk_global_pass: bad argument");
  ctk::BatchArgs a;
  memset(&a, 0, sizeof(a));
  a.prob = *prob;
  a.frames = d_frames;
  for (int k = 0; k < prob->ndim; ++k) {
    if (frame_shape[k] <= 0 || frame_shape[k] > (1 << 30))
      return fail(CTK_E_INVALID, "ctk_global_pass: bad frame shape");
    a.shape[k] = frame_shape[k];
  }
  a.n_work = n_clusters;
  a.cluster_frame = d_cluster_frame;
  a.cluster_offset = d_cluster_offset;
  a.params_in = d_params_in;
  a.params_out = d_params_out;
  a.cost_out = d_cost_out;
  a.status_out = d_status_out;
  a.counter = static_cast<int32_t*>(d_workspace);
  a.mask_centres = d_mask_centres;
  a.global_step = d_global_step;
  a.global_accum = d_accum;
  a.global_norm = norm;
  a.global_lambda = lambda;
  a.global_phase = phase;
  a.use_newton = use_newton;
  ctk_problem_t rigorous = *prob;
  rigorous.capacity_mode = 1;                      // no relaunch logic here: worst-case capacities
  if (!ctk::compute_layout(rigorous, max_cluster_features, &a.lay))
    return fail(CTK_E_CAPACITY, "ctk_global_pass: max_cluster_features %d out of range", max_cluster_features);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CTK_CUDA(cudaMemsetAsync(d_workspace, 0, sizeof(int32_t), st));
  // the full-flavour instances carry the pass; select them like a problem with constraints would
  ctk_problem_t select = *prob;
  select.constraint_mask = CTK_CONSTRAINT_DIMER;
  GlobalLauncher launcher{&a, st, 0};
  bool found = prob->compute_dtype == CTK_COMPUTE_F64 ? ctk::dispatch_config<double>(select, launcher, false)
                                                      : ctk::dispatch_config<float>(select, launcher, false);
  if (!found) return fail(CTK_E_UNSUPPORTED, "ctk_global_pass: no kernel instance for this problem");
  return launcher.result;
}

}  // extern "C"
