// ctk_label.cu -- find_clusters labels on the device (find.py:72-93): kernel + C ABI.  sm_100a only.
//
// One warp per frame (ctk_label.cuh), one warp per block, blocks pull frames from an atomic counter.
// A frame's point arrays live in shared memory when they fit; trees, pair lists and the set tables
// live in a per-block scratch in global memory (L1/L2 resident: ~100 KB per frame of config 2).
#include <cuda_runtime.h>
#include <stdio.h>

#include "ctk.h"
#include "ctk_label.cuh"

namespace {

struct LabelArgs {
  const double* pos[3];
  const int64_t* starts;
  const int64_t* stops;
  double separation[3];
  ctk_label::Caps caps;
  char* scratch;              // per block: stride bytes
  int64_t stride;
  int32_t* counter;
  int32_t* labels;
  int32_t* flags;
  int32_t n_frames, ndim;
  int32_t window;             // bytes of shared memory per block (the fast window of ctk_label.cuh)
};

__global__ void __launch_bounds__(32, 1) label_kernel(const LabelArgs a) {
  extern __shared__ __align__(16) char smem[];
  const int lane = threadIdx.x;
  const ctk_label::Scratch base = ctk_label::carve(a.scratch + (int64_t) blockIdx.x * a.stride, a.caps);
  const double* pos[3] = {a.pos[0], a.pos[1], a.pos[2]};
  const double sep[3] = {a.separation[0], a.separation[1], a.separation[2]};
  for (;;) {
    int f = 0;
    if (lane == 0) f = atomicAdd(a.counter, 1);
    f = __shfl_sync(0xffffffffu, f, 0);
    if (f >= a.n_frames) break;
    const int64_t row0 = a.starts[f];
    const int64_t cnt = a.stops[f] - row0;
    if (cnt <= 0) {
      if (lane == 0) *reinterpret_cast<volatile int32_t*>(a.flags + f) = ctk_label::FLAG_OK;
      continue;
    }
    int flag = ctk_label::FLAG_CAPACITY;
    if (cnt <= a.caps.points) {
      ctk_label::FrameLabeller fl;
      fl.s = base;
      fl.caps = a.caps;
      fl.timing = reinterpret_cast<unsigned long long*>(a.counter) + 8;      // header bytes 64..127
      ctk_label::use_window(fl.s, a.window > 0 ? smem + 256 : nullptr, a.window, (int) cnt, a.ndim);
      if (a.window > 0) fl.s.stage = reinterpret_cast<uint64_t*>(smem);
      flag = fl.run(pos, row0, (int) cnt, a.ndim, sep, 1.0, a.labels);
    }
    // labels and flags may live in mapped host memory, consumed frame by frame: the labels of the
    // frame must be visible before its flag
    __threadfence_system();
    __syncwarp();
    if (lane == 0) *reinterpret_cast<volatile int32_t*>(a.flags + f) = flag;
    __syncwarp();
  }
}

struct LabelPlan {
  ctk_label::Caps caps;
  int64_t stride;         // scratch bytes per block
  int64_t slots;          // blocks
  int32_t window;         // fast window bytes per block
  int32_t smem_bytes;     // window + staging
};

const int64_t kHeader = 256;            // the frame counter lives in front of the scratch
const int64_t kPairFactor = 6;          // pairs kept per frame: 6 * points + 1024 (more -> host path)

int make_plan(int64_t max_points, int32_t ndim, int64_t n_frames, LabelPlan* plan) {
  int device = 0, sms = 0, smem_optin = 0;
  if (cudaGetDevice(&device) != cudaSuccess) return CTK_E_CUDA;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return CTK_E_CUDA;
  if (cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess)
    return CTK_E_CUDA;
  if (max_points > (1 << 24)) max_points = 1 << 24;          // larger frames are flagged for the host path
  plan->caps = ctk_label::make_caps(max_points, kPairFactor);
  plan->stride = ctk_label::scratch_bytes(plan->caps);
  // the fast window: a frame of max_points completely, as long as two blocks still fit an SM;
  // larger frames keep their point arrays in the scratch and use the window for tables only
  int64_t window = ctk_label::window_bytes(plan->caps.points, ndim);
  const int64_t window_cap = 100 * 1024 < smem_optin - 1024 ? 100 * 1024 : smem_optin - 1024;
  if (window > window_cap) window = 64 * 1024 < window_cap ? 64 * 1024 : window_cap;
  window = window / 16 * 16;
  plan->window = (int32_t) window;
  plan->smem_bytes = (int32_t) (window + 256);
  int64_t per_sm = (225 * 1024) / (plan->smem_bytes + 1024);
  if (per_sm > 16) per_sm = 16;
  if (per_sm < 1) per_sm = 1;
  int64_t slots = (int64_t) sms * per_sm;
  if (slots > n_frames) slots = n_frames;
  const int64_t budget = (int64_t) 8 << 30;                  // scratch budget
  if (slots * plan->stride > budget) slots = budget / plan->stride;
  if (slots < 1) slots = 1;
  plan->slots = slots;
  return 0;
}

}  // namespace

extern "C" int ctk_label_frames_scratch(int64_t max_points, int32_t ndim, int64_t n_frames,
                                        int64_t* bytes_out) {
  if (max_points < 0 || ndim < 1 || ndim > 3 || n_frames < 0 || !bytes_out) return CTK_E_INVALID;
  LabelPlan plan;
  const int rc = make_plan(max_points, ndim, n_frames < 1 ? 1 : n_frames, &plan);
  if (rc) return rc;
  *bytes_out = kHeader + plan.slots * plan.stride;
  return 0;
}

extern "C" int ctk_label_frames(const double* const* d_pos_cols, int32_t ndim, const int64_t* d_starts,
                                const int64_t* d_stops, int64_t n_frames, int64_t max_points,
                                const double* separation, int32_t* d_labels,
                                int32_t* d_flags, void* d_scratch, int64_t scratch_bytes,
                                void* stream) {
  if (ndim < 1 || ndim > 3 || n_frames < 0 || max_points < 0 || !separation || !d_pos_cols)
    return CTK_E_INVALID;
  if (n_frames == 0) return 0;
  if (n_frames > (1 << 30) || !d_starts || !d_stops || !d_labels || !d_flags || !d_scratch)
    return CTK_E_INVALID;
  LabelPlan plan;
  int rc = make_plan(max_points, ndim, n_frames, &plan);
  if (rc) return rc;
  if (scratch_bytes < kHeader + plan.stride) return CTK_E_CAPACITY;
  int64_t slots = (scratch_bytes - kHeader) / plan.stride;
  if (slots > plan.slots) slots = plan.slots;
  LabelArgs a;
  for (int k = 0; k < 3; ++k) {
    a.pos[k] = k < ndim ? d_pos_cols[k] : nullptr;
    a.separation[k] = k < ndim ? separation[k] : 1.;
    if (k < ndim && (!d_pos_cols[k] || !(separation[k] == separation[k]))) return CTK_E_INVALID;
  }
  a.starts = d_starts;
  a.stops = d_stops;
  a.caps = plan.caps;
  a.counter = reinterpret_cast<int32_t*>(d_scratch);
  a.scratch = static_cast<char*>(d_scratch) + kHeader;
  a.stride = plan.stride;
  a.labels = d_labels;
  a.flags = d_flags;
  a.n_frames = (int32_t) n_frames;
  a.ndim = ndim;
  a.window = plan.window;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#ifdef CTK_LABEL_TIMING
  if (cudaMemsetAsync(a.counter, 0, 128, st) != cudaSuccess) return CTK_E_CUDA;
#else
  if (cudaMemsetAsync(a.counter, 0, sizeof(int32_t), st) != cudaSuccess) return CTK_E_CUDA;
#endif
  if (plan.smem_bytes > 48 * 1024 &&
      cudaFuncSetAttribute(label_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes) != cudaSuccess)
    return CTK_E_CUDA;
  label_kernel<<<(unsigned) slots, 32, plan.smem_bytes, st>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : CTK_E_CUDA;
}
