/*
 * ctk.h -- C ABI of libctk, the B200 (sm_100a) implementation of clustertracking's
 * least-squares refinement hot path.
 *
 * The reference (caspervdw/clustertracking) is pure Python and has no FFI of its own; its boundary
 * for this path is the public function `refine_leastsq` (clustertracking/refine.py:82-452).  The
 * entry points below are what a ctypes binding inside that function would call in place of its
 * per-cluster Python loop (refine.py:343-430): one call per batch of frames per device.  Each
 * declaration cites the reference lines it replaces.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C, no torch / C++ types; all sizes explicit;
 *   - every `d_` pointer is a DEVICE pointer owned by the caller; the library allocates nothing
 *     that outlives a call and never synchronises the stream (except where stated);
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = default stream);
 *   - return value 0 on success, a negative `CTK_E_*` code otherwise; `ctk_last_error()` returns a
 *     thread-local message.  Per-cluster numerical failures are NOT errors: they are reported in
 *     `status_out` (reference semantics: cost = NaN, parameters unchanged, refine.py:408-418).
 */
#ifndef CTK_H_
#define CTK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTK_VERSION 100          /* major*10000 + minor*100 + patch */
#define CTK_MAX_PARAMS 12        /* background, signal, <=3 pos, <=3 size, <=1 extra (+ spare) */
#define CTK_MAX_CLUSTER_FEATURES 32   /* features per cluster of the shared-memory kernels */
#define CTK_MAX_BIG_FEATURES 256      /* features per cluster of the large-cluster kernels (per-cluster
                                         arrays in a global-memory workspace, see ctk_refine_batch) */
#define CTK_MAX_RADIUS 30        /* mask radius per axis (pixel offsets are packed in 6 bits) */
#define CTK_MAX_TAPS 33          /* taps of the lowpass kernel per axis: half width <= 16 */
#define CTK_PROBE_MAX_FEATURES 8 /* ring / disc: clusters of up to this many features get the basin
                                    search (ctk_problem_t.probe_step); larger ones one minimisation */

/* parameter modes, same codes as fitfunc.py:9-11.  CTK_MODE_GLOBAL (one value for ALL features,
 * refine.py:319-332) is only valid for ctk_global_pass. */
enum { CTK_MODE_CONST = 0, CTK_MODE_VAR = 1, CTK_MODE_GLOBAL = 2, CTK_MODE_CLUSTER = 3 };
#define CTK_GLOBAL_HEADER 8      /* leading entries of the ctk_global_pass accumulator */
/* radial model families, fitfunc.py:112-146, 195-204 */
enum { CTK_FAMILY_GAUSS = 0, CTK_FAMILY_RING = 1, CTK_FAMILY_DISC = 2 };
/* pixel types of the frames */
enum { CTK_PIXEL_U8 = 0, CTK_PIXEL_U16 = 1, CTK_PIXEL_F32 = 2, CTK_PIXEL_F64 = 3,
       CTK_PIXEL_I16 = 4, CTK_PIXEL_I32 = 5 };
/* arithmetic of the pixel pass.  Parameters, bounds, the gradient (J^T r) and the objective are
 * float64 in both modes; the normal matrix J^T J, its Cholesky factor and the model values follow
 * the pixel arithmetic (float32 in CTK_COMPUTE_F32, float64 in CTK_COMPUTE_F64): the step only has
 * to be a descent direction, the fixed point is decided by the float64 gradient. */
enum { CTK_COMPUTE_F32 = 0, CTK_COMPUTE_F64 = 1 };
/* equality constraints, constraints.py:59-137; applied to clusters of exactly that size */
enum { CTK_CONSTRAINT_DIMER = 1, CTK_CONSTRAINT_TRIMER = 2, CTK_CONSTRAINT_TETRAMER = 4 };

/* per-cluster status (status_out).  0 = success; anything else = the reference's RefineException
 * path: cost NaN, parameters copied through unchanged. */
enum {
  CTK_OK = 0,
  CTK_FAIL_NONFINITE = 1,     /* non-finite initial parameters          refine.py:356-357 */
  CTK_FAIL_OUT_OF_IMAGE = 2,  /* no coordinate within the image + radius refine.py:33-34  */
  CTK_FAIL_NO_CONVERGENCE = 3,/* inner solver iteration limit           refine.py:376-377 */
  CTK_FAIL_RMS_DEV = 4,       /* rms_dev > max_rms_dev                  refine.py:391-394 */
  CTK_FAIL_BOUNDS = 5,        /* lower bound above upper bound                            */
  CTK_FAIL_TOO_LARGE = 6,     /* cluster exceeds the capacity of this launch (see ctk_refine_batch) */
  CTK_FAIL_NUMERIC = 7        /* non-finite value met during the fit    fitfunc.py:437-438 */
};

/* per-cluster counters written to stats_out[cluster * CTK_STATS + k] (accounting / roofline) */
#define CTK_STATS 8
enum {
  CTK_STAT_EVALS = 0,         /* objective evaluations = passes over the per-feature pixel lists */
  CTK_STAT_ACCUMS = 1,        /* normal-equation accumulations (accepted steps + 1 per restart)  */
  CTK_STAT_OUTER = 2,         /* outer re-mask iterations executed            refine.py:365      */
  CTK_STAT_PIXELS = 3,        /* union-mask pixels M of the last pixel set    refine.py:47       */
  CTK_STAT_ENTRIES = 4,       /* sum over features of mask pixels (last pixel set)               */
  CTK_STAT_PAIR_ENTRIES = 5,  /* pixels shared by two features, summed over pairs                */
  CTK_STAT_VARS = 6,          /* free variables V                                                */
  CTK_STAT_GRAD_ACCUMS = 7    /* gradient-only accumulations (chord iterations, factor reused)   */
};

/* library error codes (return values) */
enum {
  CTK_E_INVALID = -1,         /* bad argument */
  CTK_E_UNSUPPORTED = -2,     /* combination not built into the library */
  CTK_E_CUDA = -3,            /* a CUDA runtime call failed */
  CTK_E_CAPACITY = -4,        /* requested capacity does not fit the device's shared memory */
  CTK_E_NONFINITE = -5        /* host helpers: non-finite coordinates (scipy's kd-tree refuses them too) */
};

/* Problem description, host memory, plain old data.  Mirrors the keyword arguments of
 * refine_leastsq (refine.py:82-87) after FitFunctions.__init__ (fitfunc.py:325-413) resolved them. */
typedef struct {
  int32_t ndim;               /* 2 | 3                                   refine.py:254-283 */
  int32_t isotropic;          /* 1: one `size` column, 0: ndim columns   refine.py:287      */
  int32_t family;             /* CTK_FAMILY_*                            fitfunc.py:195-204 */
  int32_t n_params;           /* P = 2 + ndim + n_size + n_extra         fitfunc.py:353-354 */
  int32_t modes[CTK_MAX_PARAMS]; /* per column, order background, signal, pos.., size.., extra */
  int32_t radius[3];          /* mask radii diameter//2, axis order z,y,x (first ndim used) */
  int32_t pixel_dtype;        /* CTK_PIXEL_*                                                */
  int32_t compute_dtype;      /* CTK_COMPUTE_*                                              */
  int32_t max_iter;           /* outer re-mask iterations                refine.py:365      */
  int32_t lm_max_iter;        /* inner solver iteration cap (SLSQP maxiter, refine.py:243)  */
  double  max_shift;          /* refine.py:383-385 */
  double  max_rms_dev;        /* refine.py:391-394 */
  double  residual_factor;    /* refine.py:354, 379 */
  double  xtol;               /* inner solver step tolerance; <= 0 selects the default      */
  double  chord_tol;          /* reuse the factorised normal matrix once the scaled step is
                                 below this (0 = never); only the gradient is refreshed then  */
  int32_t constraint_mask;    /* OR of CTK_CONSTRAINT_*                                     */
  int32_t capacity_mode;      /* 0: size the per-cluster shared memory for the typical cluster of
                                 the class (overflow -> CTK_FAIL_TOO_LARGE, relaunch with 1);
                                 1: size it for the rigorous worst case                      */
  double  dimer_dist[3];      /* constraints.py:70-76, per axis */
  double  trimer_dist[3];     /* constraints.py:93-99, per axis */
  double  tetramer_dist[3];   /* constraints.py:127-137, per axis (2D: square, 3D: tetrahedron) */
  /* Bounds tables of FitFunctions.validate_bounds (fitfunc.py:492-533), [0] = lower, [1] = upper,
   * NaN = no bound.  Used when d_bounds_lo / d_bounds_hi are NULL: per feature and column
   *   low  = fmax(fmax(p - diff[0], p * (1 - rel[0])), abs[0]), NaN -> -inf
   *   high = fmin(fmin(p + diff[1], p * (1 + rel[1])), abs[1]), NaN -> +inf   (fitfunc.py:538-551) */
  double  bounds_abs[2][CTK_MAX_PARAMS];
  double  bounds_diff[2][CTK_MAX_PARAMS];
  double  bounds_rel[2][CTK_MAX_PARAMS];
  /* Lowpass of the cluster's sub-image before masking (`noise_size`, `threshold`; refine.py:36-40
   * -> preprocessing.py:12-49): separable gaussian over the cluster's bounding box, zero padded at
   * the BOX edge, then values <= threshold are set to 0. */
  int32_t lowpass;              /* 0: off (noise_size is None) */
  int32_t lowpass_half[3];      /* half width lw = int(4 sigma + 0.5) per axis; -1: axis not filtered */
  double  lowpass_threshold;    /* refine.py:38-39: 0 when `threshold` is None */
  double  lowpass_sigma[3];     /* noise_size per axis; the device builds the taps of
                                   trackpy.masks.gaussian_kernel(sigma, 4): exp(-x^2 / (2 sigma^2)),
                                   x = -lw .. lw, normalised to sum 1 */
  /* Basin search of the ring / disc families (piecewise-smooth objectives, fitfunc.py:20-26,
   * 121-131): after the first minimisation every centre is displaced by +-probe_step pixels along
   * every axis (and a disc with a free disc_size is switched to its gauss branch), the minimiser
   * runs again and a lower end point replaces the current one; at most probe_sweeps sweeps, stopping
   * when a sweep brings no improvement.  probe_step <= 0 or probe_sweeps <= 0: one minimisation only
   * (always the case for the smooth gauss family). */
  double  probe_step;
  int32_t probe_sweeps;
  int32_t reserved_;
} ctk_problem_t;

int ctk_version(void);
/* sizeof(ctk_problem_t) as compiled into the library: lets a binding check its struct mirror */
size_t ctk_problem_bytes(void);
const char* ctk_last_error(void);

/* Maximum pixel value of each frame, as float64 (replaces `frame.max()` at refine.py:354, which the
 * reference re-evaluates for every cluster).  HBM-bound streaming reduction.
 *   d_frames   [n_frames] device array of device pointers to C-contiguous frames
 *   n_pixels   pixels per frame
 *   d_max_out  [n_frames] float64 */
int ctk_frame_max(const void* const* d_frames, int32_t n_frames, int64_t n_pixels,
                  int32_t pixel_dtype, double* d_max_out, void* stream);

/* Bytes of scratch `ctk_refine_batch` needs in `d_workspace` for launches with
 * max_cluster_features <= CTK_MAX_CLUSTER_FEATURES. */
size_t ctk_refine_workspace_bytes(void);

/* Bytes of scratch a launch with this capacity needs.  Above CTK_MAX_CLUSTER_FEATURES the
 * per-cluster arrays (pixel lists, normal matrix, ...) live in this workspace instead of shared
 * memory; 0 = the request is out of range. */
size_t ctk_refine_workspace_bytes_for(const ctk_problem_t* prob, int32_t max_cluster_features);

/* Shared memory (bytes per cluster) a launch with this capacity would use, or 0 when it does not
 * fit the device limit (227 KB on sm_100a).  Lets the caller bin clusters by size. */
size_t ctk_refine_shared_bytes(const ctk_problem_t* prob, int32_t max_cluster_features);

/* Refine a batch of clusters: the body of the reference's loop over (frame, cluster) groups
 * (refine.py:343-430) including the pixel-set construction (refine.py:28-58, masks.py:30-68), the
 * objective (fitfunc.py:421-489), the bounds (fitfunc.py:535-558), the dimer/trimer constraints
 * (constraints.py:59-99), the outer re-mask loop (refine.py:365-388) and the rms check (391-394).
 * The reference minimises with scipy's SLSQP (refine.py:373-375); this library reaches the same
 * bound/equality-constrained minimum with a projected Levenberg-Marquardt iteration and an
 * augmented-Lagrangian treatment of the distance constraints, entirely on the device.
 *
 *   d_frames        [n_frames] device array of device pointers to C-contiguous frames
 *   frame_shape     [ndim] (host) frame shape, axis order z,y,x
 *   d_frame_max     [n_frames] float64, from ctk_frame_max
 *   n_work          number of clusters to process in this launch
 *   d_work_ids      [n_work] cluster indices to process (NULL = 0..n_work-1); put expensive first
 *   max_cluster_features  capacity of this launch; clusters with more features (or whose pixel
 *                   lists overflow the derived capacity, see prob->capacity_mode) get
 *                   CTK_FAIL_TOO_LARGE
 *   d_cluster_frame [n_clusters] index into d_frames
 *   d_cluster_offset[n_clusters + 1] feature ranges; features of a cluster are consecutive rows
 *   d_params_in     [n_features, P] float64 row-major, columns as in `modes`   (refine.py:345)
 *   d_bounds_lo/hi  [n_features, P] float64 per-feature bounds, +-inf allowed  (fitfunc.py:538-551);
 *                   both NULL = derive them on the device from prob->bounds_* and d_params_in
 *   d_params_out    [n_features, P] float64                                   (refine.py:380, 426)
 *   d_cost_out      [n_clusters] rms_dev, NaN on failure                      (refine.py:379, 427)
 *   d_status_out    [n_clusters] CTK_OK or CTK_FAIL_*
 *   d_stats_out     [n_clusters, CTK_STATS] int32 counters (CTK_STAT_*), for accounting
 *   d_workspace     ctk_refine_workspace_bytes_for(prob, max_cluster_features) bytes of device scratch
 */
int ctk_refine_batch(const ctk_problem_t* prob,
                     const void* const* d_frames, const int64_t* frame_shape,
                     const double* d_frame_max,
                     int32_t n_work, const int32_t* d_work_ids, int32_t max_cluster_features,
                     const int32_t* d_cluster_frame, const int32_t* d_cluster_offset,
                     const double* d_params_in, const double* d_bounds_lo, const double* d_bounds_hi,
                     double* d_params_out, double* d_cost_out, int32_t* d_status_out,
                     int32_t* d_stats_out, void* d_workspace, void* stream);

/* ctk_refine_batch with two device-side hooks, so that overflowing clusters can be relaunched
 * without a host round trip:
 *   d_n_work        optional device int32: the launch processes min(n_work, *d_n_work) work items
 *                   (n_work still sizes the grid).  Typically the count word of an overflow list.
 *   d_overflow      optional device int32 [1 + overflow_capacity]: word 0 is reset to 0 at the
 *                   start of the launch (stream-ordered) and counts the clusters that ended with
 *                   CTK_FAIL_TOO_LARGE; their indices are appended at words 1.. (order unspecified).
 *                   A follow-up launch with prob->capacity_mode = 1, d_work_ids = d_overflow + 1 and
 *                   d_n_work = d_overflow refines exactly those clusters. */
int ctk_refine_batch_chained(const ctk_problem_t* prob,
                     const void* const* d_frames, const int64_t* frame_shape,
                     const double* d_frame_max,
                     int32_t n_work, const int32_t* d_work_ids, int32_t max_cluster_features,
                     const int32_t* d_cluster_frame, const int32_t* d_cluster_offset,
                     const double* d_params_in, const double* d_bounds_lo, const double* d_bounds_hi,
                     double* d_params_out, double* d_cost_out, int32_t* d_status_out,
                     int32_t* d_stats_out, void* d_workspace,
                     const int32_t* d_n_work, int32_t* d_overflow, int32_t overflow_capacity,
                     void* stream);

/* ctk_refine_batch_chained with launch flags:
 *   CTK_LAUNCH_APPEND_OVERFLOW  do not reset word 0 of d_overflow: the clusters this launch cannot
 *                   hold are APPENDED to a list earlier launches on the stream started.  Lets every
 *                   size class hand its hardest clusters to ONE final large-cluster launch. */
enum { CTK_LAUNCH_APPEND_OVERFLOW = 1 };
int ctk_refine_batch_ex(const ctk_problem_t* prob,
                     const void* const* d_frames, const int64_t* frame_shape,
                     const double* d_frame_max,
                     int32_t n_work, const int32_t* d_work_ids, int32_t max_cluster_features,
                     const int32_t* d_cluster_frame, const int32_t* d_cluster_offset,
                     const double* d_params_in, const double* d_bounds_lo, const double* d_bounds_hi,
                     double* d_params_out, double* d_cost_out, int32_t* d_status_out,
                     int32_t* d_stats_out, void* d_workspace,
                     const int32_t* d_n_work, int32_t* d_overflow, int32_t overflow_capacity,
                     int32_t flags, void* stream);

/* One pass of a GLOBAL-level fit (param_mode 'global': some columns are shared by all features, so
 * the whole table is one problem; refine.py:319-332, 343-394 with level == 'global', objective
 * fitfunc.py:421-489 with groups).  The host iterates a damped Newton / Gauss-Newton method on the
 * block-arrow system; every iteration calls this twice over all clusters:
 *   phase 1  at d_params_in: per cluster the pixel set, residuals, normal equations, elimination of
 *            the cluster's own unknowns; adds to d_accum (which the caller zeroes):
 *              [0] F = sum_c nansum(diff^2) / M_c / norm          (the reference's objective)
 *              [4] clusters that failed (out of image, non-finite)   [5] singular eliminations
 *              [8 .. 8+G)            reduced right-hand side of the G shared unknowns
 *              [8+G .. 8+G+G(G+1)/2) their Schur complement, lower triangle row by row
 *            -- the ONLY reduction across clusters: all-reduce d_accum when frames are sharded;
 *   phase 2  given d_global_step [G] (the host solved the summed system): per cluster the step of
 *            its own unknowns by back substitution, the trial point projected on its bounds,
 *            written to d_params_out; adds [1] F at the trial point, [2] the decrease predicted by
 *            the quadratic model, [3] (max) the largest scaled step of a per-cluster unknown.
 * Shared unknowns are numbered in column order.  modes may contain CTK_MODE_GLOBAL; constraints
 * are not supported.  lambda: Marquardt damping of the per-cluster blocks; use_newton: add the
 * second-order term of the Hessian (gauss family).  d_mask_centres [n_features, ndim] or NULL
 * (= the positions in d_params_in): centres of the pixel sets (refine.py:365-388 re-centres them
 * between minimisations).  norm = max over frames of max()^2 / residual_factor (refine.py:325-331).
 * max_cluster_features <= CTK_MAX_CLUSTER_FEATURES. */
int ctk_global_pass(const ctk_problem_t* prob,
                    const void* const* d_frames, const int64_t* frame_shape, double norm,
                    int32_t n_clusters, int32_t max_cluster_features,
                    const int32_t* d_cluster_frame, const int32_t* d_cluster_offset,
                    const double* d_params_in, const double* d_mask_centres,
                    int32_t phase, double lambda, int32_t use_newton,
                    const double* d_global_step, double* d_params_out, double* d_accum,
                    double* d_cost_out, int32_t* d_status_out, void* d_workspace, void* stream);

/* Host helper (no GPU): cluster labels of one frame from the close pairs, visiting the pairs in the
 * given order with the reference's "the label of a's cluster survives" rule (find.py:41-48, 84-93).
 *   pairs [n_pairs, 2] int64; labels_out, sizes_out [n] int64 */
int ctk_label_clusters(const int64_t* pairs, int64_t n_pairs, int64_t n,
                       int64_t* labels_out, int64_t* sizes_out);

/* Host helper (no GPU): the order in which CPython iterates the set that
 * scipy.spatial.cKDTree.query_pairs(output_type='set') builds from `pairs` given in the order of
 * query_pairs(output_type='ndarray').  The reference visits the pairs in that order (find.py:87-91),
 * which fixes its cluster label values.  order_out [n_pairs] = insertion indices in iteration order. */
int ctk_pairs_set_order(const int64_t* pairs, int64_t n_pairs, int64_t* order_out);

/* Host helper (no GPU): the close pairs scipy.spatial.cKDTree(data).query_pairs(1,
 * output_type='ndarray') reports, in the same order (find.py:87; restatement of scipy's kd-tree for
 * this one call, verified against the installed scipy by find.py).  data [n, ndim] float64;
 * pairs_out [capacity, 2] may be NULL to query the count only. */
int ctk_query_pairs(const double* data, int64_t n, int32_t ndim, int64_t* pairs_out,
                    int64_t capacity, int64_t* n_pairs_out);

/* Device: the cluster labels of find_clusters for a whole video (find.py:72-93: cKDTree(pos /
 * separation).query_pairs(1), the pairs visited in the iteration order of the python set, the union
 * rule of Clusters.add), one warp per frame, label VALUES identical to the reference's (kd-tree,
 * std::nth_element, CPython set order restated on the device; see csrc/ctk_label.cuh).
 *   d_pos_cols [ndim]  HOST array of device pointers: table-order float64 columns, rows sorted by frame
 *   d_starts, d_stops [n_frames] int64 (device): frame f owns rows d_starts[f] .. d_stops[f]-1
 *   max_points         most rows of one frame;  separation [ndim] (host)
 *   d_labels [rows] int32   label of every row, local to its frame (as ctk_cluster_frames)
 *   d_flags  [n_frames] int32  0 = labelled; 1 = a scratch capacity was exceeded (label this frame
 *                      with ctk_cluster_pack_labelled / ctk_cluster_frames on the host); 2 = non-finite
 *   d_scratch          at least ctk_label_frames_scratch(...) bytes, 256-byte aligned
 * Asynchronous on `stream`.  The columns, d_starts / d_stops, d_labels and d_flags may be MAPPED
 * pinned host memory: a frame's labels are visible to the host before its flag is written
 * (__threadfence_system), so a host thread that initialised the flags to -1 can consume the frames
 * as they complete. */
int ctk_label_frames_scratch(int64_t max_points, int32_t ndim, int64_t n_frames, int64_t* bytes_out);
int ctk_label_frames(const double* const* d_pos_cols, int32_t ndim, const int64_t* d_starts,
                     const int64_t* d_stops, int64_t n_frames, int64_t max_points,
                     const double* separation, int32_t* d_labels, int32_t* d_flags,
                     void* d_scratch, int64_t scratch_bytes, void* stream);

/* Host helper: the rows at which a new frame starts in an int64 frame column (the groups of
 * find.py:122), in one pass; *sorted_out = 1 when the column never decreases.  starts_out holds up to
 * `capacity` entries, *n_runs_out the number of runs. */
int ctk_frame_runs(const int64_t* frames, int64_t n, int64_t* starts_out, int64_t capacity,
                   int64_t* n_runs_out, int32_t* sorted_out);

/* Host helper: sleep until none of flags[0 .. n-1] (the mapped per-frame flags of ctk_label_frames,
 * initialised to -1 by the caller) is negative, or timeout_us have passed (returns 1). */
int ctk_wait_flags(const int32_t* flags, int64_t n, int64_t timeout_us);

/* Host helper (no GPU): find_clusters for a whole video on host threads (find.py:72-129).
 *   pos [n, ndim] float64, rows sorted by frame; frame f owns rows starts[f] .. stops[f]-1
 *   separation [ndim]; n_threads worker threads (frames are independent)
 *   cluster_out [n]    label of every row, local to its frame (caller adds the running offset of
 *                      find.py:127-128 from span_out)
 *   size_out [n]       cluster_size
 *   by_cluster_out [n] row permutation that lists every frame's rows by label, original order kept
 *                      inside a label (the group order of refine.py:336)
 *   span_out [n_frames] largest label of the frame + 1 */
int ctk_cluster_frames(const double* pos, int64_t n, int32_t ndim, const int64_t* starts,
                       const int64_t* stops, int64_t n_frames, const double* separation,
                       int32_t n_threads, int64_t* cluster_out, int64_t* size_out,
                       int64_t* by_cluster_out, int64_t* span_out);

/* Host helpers (no GPU) of the packing pipeline: running cluster ids (find.py:127-128), group order
 * and group table (refine.py:336) of one labelled chunk of frames; the launch schedule (clusters by
 * size class, expensive first); the gather of the packed parameter rows (refine.py:345); the
 * write-back of the results (refine.py:408-427).  See
 * ctk_host.cpp for the argument lists. */
int ctk_group_chunk(const int64_t* local_labels, const int64_t* by_cluster, const int64_t* starts,
                    const int64_t* stops, const int64_t* spans, int64_t n_frames, int64_t next_id,
                    int64_t row_base, int32_t frame_base, int64_t* cluster_out, int64_t* order_out,
                    int32_t* group_offset_out, int32_t* group_frame_out, int64_t* n_groups_out,
                    int64_t* next_id_out);
int ctk_cluster_pack_frames(const double* pos, int64_t n, int32_t ndim, const int64_t* starts,
                            const int64_t* stops, int64_t n_frames, const double* separation,
                            int32_t n_threads, int64_t* cluster_out, int64_t* size_out,
                            int64_t* by_cluster_out, int64_t* span_out,
                            const double* const* columns, const double* scalars, int32_t n_cols,
                            int64_t row_base, double* params_out, int32_t* group_count_out,
                            int32_t* group_start_out);
int ctk_cluster_pack_columns(const double* pos, const double* const* pos_cols, int64_t n,
                             int32_t ndim, const int64_t* starts, const int64_t* stops,
                             int64_t n_frames, const double* separation, int32_t n_threads,
                             int64_t* cluster_out, int64_t* size_out, int64_t* by_cluster_out,
                             int64_t* span_out, const double* const* columns, const double* scalars,
                             int32_t n_cols, int64_t row_base, double* params_out,
                             int32_t* group_count_out, int32_t* group_start_out);
/* The same with the labels of the frames already computed on the device (ctk_label_frames):
 * labels_in [n] int32 for this call's rows, frame_flags [n_frames]; frames whose flag is not 0 are
 * labelled on the host as above. */
int ctk_cluster_pack_labelled(const double* pos, const double* const* pos_cols, int64_t n,
                              int32_t ndim, const int64_t* starts, const int64_t* stops,
                              int64_t n_frames, const double* separation, int32_t n_threads,
                              int64_t* cluster_out, int64_t* size_out, int64_t* by_cluster_out,
                              int64_t* span_out, const double* const* columns, const double* scalars,
                              int32_t n_cols, int64_t row_base, double* params_out,
                              int32_t* group_count_out, int32_t* group_start_out,
                              const int32_t* labels_in, const int32_t* frame_flags);
int ctk_concat_groups(const int64_t* starts, const int64_t* stops, const int32_t* group_count,
                      const int32_t* group_start, int64_t n_frames, int32_t frame_base,
                      int32_t* group_offset_out, int32_t* group_frame_out, int64_t* n_groups_out);
int ctk_apply_label_offsets(const int64_t* local, const int64_t* starts, const int64_t* stops,
                            const int64_t* frame_offset, int64_t n_frames, int32_t n_threads,
                            int64_t* cluster_out);
int ctk_schedule(const int32_t* cluster_offset, int64_t n_clusters, const int32_t* caps,
                 int32_t n_caps, const int32_t* class_target, int32_t* work_ids_out,
                 int64_t* class_count_out, int32_t* not_run_out, int64_t* n_not_run_out);
int ctk_gather_rows(const double* const* columns, const double* scalars, const int64_t* rows,
                    int64_t n, int32_t n_cols, double* out, int32_t n_threads);
int ctk_scatter_rows(const double* params, const double* params_in, const int64_t* rows,
                     int64_t row_base, int64_t n, int32_t n_cols, const int32_t* group_offset, const double* group_cost,
                     const int32_t* group_status, int64_t n_groups, double* const* columns,
                     double* cost_out, int32_t n_threads, int64_t* n_failed_out);

/* Local maxima of a batch of frames: the image part of clustertracking's grey_dilation
 * (find.py:219-270), the step that produces the initial coordinates of the refinement:
 *   threshold = np.percentile(image[image != 0], percentile)                 find.py:210-216
 *   dilation  = scipy.ndimage.grey_dilation(image, size, mode='constant')    find.py:255-259
 *   maxima    = (image == dilation) & (image > threshold), minus the pixels closer than `margin`
 *               to an edge, listed in C order                                find.py:260-270
 * (the pair-wise drop_close step, find.py:166-206, is host code).  Integer frames only
 * (CTK_PIXEL_U8, CTK_PIXEL_U16).  HBM-bound streaming kernels, asynchronous on `stream`.
 *   d_frames [n_frames] device pointers to C-contiguous frames; frame_shape [ndim], ndim 2 | 3
 *   size [ndim]    box of the dilation, int(2 * separation / sqrt(ndim))     find.py:252
 *   margin [ndim]  exclusion zone at the edges
 *   capacity       maxima stored per frame
 *   d_coords_out [n_frames, capacity, ndim] int32; d_values_out [n_frames, capacity] int32 (pixel
 *   value of every maximum); d_count_out [n_frames] int32 number found (may exceed capacity: call
 *   again with more room); d_threshold_out [n_frames] float64 (NaN: the frame is all black)
 *   d_workspace    ctk_find_workspace_bytes(n_frames, pixels per frame, pixel_dtype) bytes
 *   frames_aligned 1: every frame pointer is 4-byte aligned (enables the packed uint8 filter) */
size_t ctk_find_workspace_bytes(int32_t n_frames, int64_t n_pixels, int32_t pixel_dtype);
int ctk_find_maxima(const void* const* d_frames, int32_t n_frames, const int64_t* frame_shape,
                    int32_t ndim, int32_t pixel_dtype, const int32_t* size, double percentile,
                    const int32_t* margin, int32_t capacity, int32_t* d_coords_out,
                    int32_t* d_values_out, int32_t* d_count_out, double* d_threshold_out,
                    void* d_workspace, int32_t frames_aligned, void* stream);
const char* ctk_find_last_error(void);

/* drop_close (find.py:166-207) for the maxima of a batch of frames as written by ctk_find_maxima, on
 * host threads: keep_out [n_frames, capacity] uint8 flags, kept_out [n_frames] counts. */
int ctk_drop_close_frames(const int32_t* coords, const int32_t* values, const int32_t* counts,
                          int64_t n_frames, int32_t capacity, int32_t ndim, const double* separation,
                          int32_t n_threads, uint8_t* keep_out, int32_t* kept_out);

/* ctk_query_pairs with a radius: pairs closer than `r` (the SET scipy's
 * cKDTree(data, leafsize).query_pairs(r) reports for any leafsize; used by where_close,
 * find.py:166-199, whose result does not depend on the order of the pairs). */
int ctk_query_pairs_within(const double* data, int64_t n, int32_t ndim, double r,
                           int64_t* pairs_out, int64_t capacity, int64_t* n_pairs_out);

#ifdef __cplusplus
}
#endif
#endif  /* CTK_H_ */
