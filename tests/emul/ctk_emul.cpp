// ctk_emul.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the device solver source (clustertracking_b200/csrc/ctk_solver.cuh) as plain C++ with a
// one-lane "warp" so that the solver LOGIC (pixel sets, normal equations, bound handling,
// constraints, outer loop, failure statuses) can be exercised by the CPU tests in the GPU-less
// build container.  It is built on demand by tests/emul_backend.py into tests/emul/_build/ and is
// never loaded by the clustertracking_b200 package, bench.py or smoke(): the product path is the
// CUDA library and fails loudly without it.
#define CTK_EMUL 1
#include <stdlib.h>
#include <limits.h>
#include <algorithm>
using std::min;
using std::max;
#include "ctk_layout.h"

namespace {
struct RunAll {
  const ctk::BatchArgs* args;
  template <class C> void operator()() const {
    char* sm = static_cast<char*>(aligned_alloc(128, args->lay.total));
    for (int w = 0; w < args->n_work; ++w) {
      int cluster = args->work_ids ? args->work_ids[w] : w;
      memset(sm, 0xCD, args->lay.total);           // poison: stale data must never be relied on
      ctk::ClusterSolver<C> solver(*args, sm);
      solver.run(cluster);
    }
    free(sm);
  }
};
}  // namespace

extern "C" int ctk_emul_refine_batch(const ctk_problem_t* prob, const void* const* frames,
                                     const int64_t* frame_shape, const double* frame_max,
                                     int32_t n_work, const int32_t* work_ids,
                                     int32_t max_cluster_features, const int32_t* cluster_frame,
                                     const int32_t* cluster_offset, const double* params_in,
                                     const double* bounds_lo, const double* bounds_hi,
                                     double* params_out, double* cost_out, int32_t* status_out,
                                     int32_t* iters_out) {
  if (ctk::validate_problem(*prob)) return CTK_E_INVALID;
  ctk::BatchArgs a;
  memset(&a, 0, sizeof(a));
  a.prob = *prob;
  a.frames = frames;
  for (int k = 0; k < prob->ndim; ++k) a.shape[k] = frame_shape[k];
  a.frame_max = frame_max;
  a.n_work = n_work;
  a.work_ids = work_ids;
  a.cluster_frame = cluster_frame;
  a.cluster_offset = cluster_offset;
  a.params_in = params_in;
  a.lo_in = bounds_lo;
  a.hi_in = bounds_hi;
  a.params_out = params_out;
  a.cost_out = cost_out;
  a.status_out = status_out;
  a.stats_out = iters_out;
  if (!ctk::compute_layout(*prob, max_cluster_features, &a.lay)) return CTK_E_CAPACITY;
  RunAll run{&a};
  const bool big = max_cluster_features > CTK_MAX_CLUSTER_FEATURES;
  bool ok = prob->compute_dtype == CTK_COMPUTE_F64 ? ctk::dispatch_config<double>(*prob, run, big)
                                                   : ctk::dispatch_config<float>(*prob, run, big);
  return ok ? 0 : CTK_E_UNSUPPORTED;
}

extern "C" int ctk_emul_layout_bytes(const ctk_problem_t* prob, int32_t max_cluster_features) {
  ctk::Layout lay;
  if (!ctk::compute_layout(*prob, max_cluster_features, &lay)) return 0;
  return lay.total;
}

namespace {
struct RunGlobal {
  const ctk::BatchArgs* args;
  template <class C> void operator()() const {
    char* sm = static_cast<char*>(aligned_alloc(128, args->lay.total));
    for (int w = 0; w < args->n_work; ++w) {
      memset(sm, 0xCD, args->lay.total);
      ctk::ClusterSolver<C> solver(*args, sm);
      solver.run_global(w);
    }
    free(sm);
  }
};
}  // namespace

// host-emulated counterpart of ctk_global_pass (same arguments, host pointers)
extern "C" int ctk_emul_global_pass(const ctk_problem_t* prob, const void* const* frames,
                                    const int64_t* frame_shape, double norm, int32_t n_clusters,
                                    int32_t max_cluster_features, const int32_t* cluster_frame,
                                    const int32_t* cluster_offset, const double* params_in,
                                    const double* mask_centres, int32_t phase, double lambda,
                                    int32_t use_newton, const double* global_step,
                                    double* params_out, double* accum, double* cost_out,
                                    int32_t* status_out) {
  if (ctk::validate_problem(*prob, true)) return CTK_E_INVALID;
  ctk::BatchArgs a;
  memset(&a, 0, sizeof(a));
  a.prob = *prob;
  a.frames = frames;
  for (int k = 0; k < prob->ndim; ++k) a.shape[k] = frame_shape[k];
  a.n_work = n_clusters;
  a.cluster_frame = cluster_frame;
  a.cluster_offset = cluster_offset;
  a.params_in = params_in;
  a.params_out = params_out;
  a.cost_out = cost_out;
  a.status_out = status_out;
  a.mask_centres = mask_centres;
  a.global_step = global_step;
  a.global_accum = accum;
  a.global_norm = norm;
  a.global_lambda = lambda;
  a.global_phase = phase;
  a.use_newton = use_newton;
  ctk_problem_t rigorous = *prob;
  rigorous.capacity_mode = 1;
  if (!ctk::compute_layout(rigorous, max_cluster_features, &a.lay)) return CTK_E_CAPACITY;
  ctk_problem_t select = *prob;
  select.constraint_mask = CTK_CONSTRAINT_DIMER;
  RunGlobal run{&a};
  bool ok = prob->compute_dtype == CTK_COMPUTE_F64 ? ctk::dispatch_config<double>(select, run, false)
                                                   : ctk::dispatch_config<float>(select, run, false);
  return ok ? 0 : CTK_E_UNSUPPORTED;
}
