// ctk_solver.cuh -- one warp refines one cluster.
//
// The reference loops over (frame, cluster) groups in Python (refine.py:343-430) and hands every
// group to scipy's SLSQP.  Here one warp owns one cluster from the first pixel load to the final
// rms check; nothing goes back to the host in between.  Per outer re-mask iteration (refine.py:365):
//
//   build_pixels()  pixel set of the cluster (refine.py:28-58, masks.py:30-68): float64 ellipse tests
//                   in the reference's exact operation order (separable tables of ((idx-c)/r)^2),
//                   warp-ballot compaction of the union into shared memory, per-feature pixel lists,
//                   and lists of the pixels shared by two features;
//   evaluate()      one pass over the per-feature lists: model values (cached) and residuals
//                   (fitfunc.py:436-450);
//   accumulate()    J^T J and J^T r from the cached model values: per-feature blocks in registers,
//                   warp-shuffle reduction, scatter into the packed normal matrix (float64, shared);
//                   cross blocks from the shared-pixel lists (the analytic Jacobian of
//                   fitfunc.py:455-487; disc gets the analytic derivative the reference lacks);
//   solve()         active-set treatment of the box bounds (fitfunc.py:535-558), Marquardt damping,
//                   packed Cholesky in shared memory, augmented-Lagrangian rows for the dimer/trimer
//                   distance constraints (constraints.py:59-99).
//
// The same source compiles for the device (nvcc, 32 lanes) and, with -DCTK_EMUL, as plain C++ with a
// one-lane "warp".  The one-lane build exists ONLY so that tests/ can exercise the solver logic in
// the GPU-less build container; the package never loads it.
#pragma once

#include <math.h>
#include <stdint.h>

#include "ctk.h"

#if defined(CTK_EMUL) && defined(CTK_TRACE)
#include <stdio.h>
#define CTK_TRACEF(...) printf(__VA_ARGS__)
#else
#define CTK_TRACEF(...)
#endif

#ifdef CTK_EMUL
#define CTK_DEV inline
#define CTK_DEV_BIG inline
#define CTK_COLD inline
#define CTK_WARP 1
#else
#define CTK_DEV __device__ __forceinline__
#define CTK_DEV_BIG __device__ __forceinline__  // large phases: each has exactly ONE call site
// Rarely executed helpers (bounds from the tables, distance constraints).  They used to be out of
// line; since the kernels without constraints no longer contain them (Config::EXTRA), inlining them
// is what pays: an out-of-line call gives the whole kernel a stack frame (368 bytes) and costs the
// constrained fits 29 % (config 3: 8.6e6 -> 1.1e7 features/s).  -DCTK_COLD_NOINLINE restores calls.
#ifdef CTK_COLD_NOINLINE
#define CTK_COLD static __device__ __noinline__
#else
#define CTK_COLD static __device__ __forceinline__
#endif
#define CTK_WARP 32
#endif

// Kernel instances come in two flavours (Config::EXTRA): the lean one has the distance-constraint
// and lowpass paths compiled out.  They are rarely used, but inside the one big inlined kernel they
// cost every launch instruction-cache space and registers: the lean flavour is 22 % faster on
// config 2.  Inside ClusterSolver, CTK_NCON / CTK_LOWPASS read as compile-time zeros when lean.
// loop of a lane over "its" unknowns: v = lane, lane + 32, ...
#define CTK_FOR_V(v) _Pragma("unroll 1") for (int v = lane; v < V; v += CTK_WARP)
#define CTK_NCON (C::EXTRA ? n_con : 0)
#define CTK_LOWPASS (C::EXTRA && a.prob.lowpass)

namespace ctk {

// ------------------------------------------------------------------------------------------------
// lane primitives
// ------------------------------------------------------------------------------------------------
#ifdef CTK_EMUL
CTK_DEV int lane_id() { return 0; }
CTK_DEV void warp_sync() {}
CTK_DEV uint32_t ballot(bool p) { return p ? 1u : 0u; }
CTK_DEV uint32_t lanemask_lt() { return 0u; }
template <class T> CTK_DEV T shfl_xor(T v, int) { return v; }
template <class T> CTK_DEV T shfl(T v, int) { return v; }
CTK_DEV int popc(uint32_t v) { return __builtin_popcount(v); }
CTK_DEV int ctz(uint32_t v) { return __builtin_ctz(v); }
CTK_DEV double dsub(double a, double b) { return a - b; }   // built with -ffp-contract=off
CTK_DEV double ddiv(double a, double b) { return a / b; }
CTK_DEV double dmul(double a, double b) { return a * b; }
CTK_DEV double dadd(double a, double b) { return a + b; }
CTK_DEV float fast_exp(float x) { return expf(x); }
CTK_DEV int atomic_next(int32_t* c) { return (*c)++; }
CTK_DEV void atomic_add_d(double* p, double v) { *p += v; }
CTK_DEV void atomic_max_nonneg_d(double* p, double v) { if (v > *p) *p = v; }
#else
CTK_DEV int lane_id() { return threadIdx.x & 31; }
CTK_DEV void warp_sync() { __syncwarp(); }
CTK_DEV uint32_t ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
CTK_DEV uint32_t lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }
template <class T> CTK_DEV T shfl_xor(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <class T> CTK_DEV T shfl(T v, int s) { return __shfl_sync(0xffffffffu, v, s); }
CTK_DEV int popc(uint32_t v) { return __popc(v); }
CTK_DEV int ctz(uint32_t v) { return __ffs((int) v) - 1; }
// float64 without FMA contraction: the mask test must round exactly like numpy (refine.py:43)
CTK_DEV double dsub(double a, double b) { return __dsub_rn(a, b); }
CTK_DEV double ddiv(double a, double b) { return __ddiv_rn(a, b); }
CTK_DEV double dmul(double a, double b) { return __dmul_rn(a, b); }
CTK_DEV double dadd(double a, double b) { return __dadd_rn(a, b); }
CTK_DEV float fast_exp(float x) { return __expf(x); }
CTK_DEV int atomic_next(int32_t* c) { return atomicAdd(c, 1); }
CTK_DEV void atomic_add_d(double* p, double v) { atomicAdd(p, v); }
// maximum of non-negative doubles: their bit patterns order like unsigned integers
CTK_DEV void atomic_max_nonneg_d(double* p, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(p), (unsigned long long) __double_as_longlong(v));
}
#endif
CTK_DEV double fast_exp(double x) { return exp(x); }
#ifdef CTK_EMUL
CTK_DEV float fast_rsqrt(float x) { return 1.0f / sqrtf(x); }
CTK_DEV double fast_rsqrt(double x) { return 1.0 / sqrt(x); }
#else
CTK_DEV float fast_rsqrt(float x) { return rsqrtf(x); }
CTK_DEV double fast_rsqrt(double x) { return rsqrt(x); }
#endif

template <class T> CTK_DEV T warp_sum(T v) {
#pragma unroll
  for (int m = CTK_WARP / 2; m > 0; m >>= 1) v += shfl_xor(v, m);
  return v;
}
CTK_DEV int warp_min_i(int v) {
#pragma unroll
  for (int m = CTK_WARP / 2; m > 0; m >>= 1) { int o = shfl_xor(v, m); v = o < v ? o : v; }
  return v;
}
CTK_DEV int warp_max_i(int v) {
#pragma unroll
  for (int m = CTK_WARP / 2; m > 0; m >>= 1) { int o = shfl_xor(v, m); v = o > v ? o : v; }
  return v;
}
CTK_DEV double warp_max_d(double v) {
#pragma unroll
  for (int m = CTK_WARP / 2; m > 0; m >>= 1) { double o = shfl_xor(v, m); v = o > v ? o : v; }
  return v;
}
CTK_DEV bool warp_any(bool p) { return ballot(p) != 0u; }
CTK_DEV bool finite_d(double v) { return fabs(v) <= 1.79769313486231571e308; }  // false for NaN

// ------------------------------------------------------------------------------------------------
// shared-memory layout of one cluster (computed on the host, identical for every warp of a launch)
// ------------------------------------------------------------------------------------------------
struct Layout {
  int n_max;       // features per cluster this launch can hold
  int v_max;       // free variables
  int m_cap;       // union pixels
  int f_cap;       // pixels of one feature's mask
  int pair_cap;    // entries over all shared-pixel lists
  int npair_cap;   // number of shared-pixel lists
  int tab_len[3];  // 2 r_k + 3 entries per (feature, axis) table
  int tab_stride;  // sum of tab_len
  int real_bytes;  // 4 | 8
  // byte offsets into the cluster's slice
  int o_x, o_xt, o_x0, o_lo, o_hi, o_rhs, o_rhsf, o_d, o_dg, o_act;
  int o_xb;                // best point of the basin search (ring / disc only)
  int o_H, o_L;            // normal matrix / its damped factor, packed lower, COLUMN-major
  int o_idg;               // inverse diagonal of the factor
  int o_cs;                // [v_max + 1] start of each packed column
  int o_rc;                // [tri(v_max)] (row | col << 8) of each packed entry
  int o_cv;                // [n_max, P] variable index of (feature, column), -1 = constant
  int o_sidx;              // [n_max, sidx_stride] scatter targets of a feature's local block
  int sidx_stride;
  int o_con;               // multipliers mu[3], distances[3], penalty weight
  int mask_words;          // 32-bit words of the per-pixel feature mask (1 unless BIG)
  int o_tmpw;              // [warp lanes, mask_words] staging of a pixel's mask words (BIG)
  int o_cmode, o_cbase;    // [P] column modes / first variable of each column
  int o_ctab;              // [P, 6] bounds tables (diff, rel, abs) x (lower, upper)
  int o_taps;              // [3, CTK_MAX_TAPS] lowpass taps (only when the problem has a lowpass)
  int o_mc, o_fi, o_fr;
  int o_tab;      // aliases o_fe (tables are dead once the lists are built)
  int o_fe;
  int o_pval, o_pr, o_pbits, o_pcrd;
  int o_flist, o_pairs, o_phdr;
  int total;
};

CTK_DEV int tri(int v) { return v * (v + 1) / 2; }

// integer record of a feature (shared memory)
enum { FI_CI = 0, FI_TS = 3, FI_CNT = 6, FI_INB = 7, FI_STRIDE = 8 };
// real record of a feature: signal, frac[3], inverse size[3], extra
enum { FR_S = 0, FR_FRAC = 1, FR_IS = 4, FR_EX = 7, FR_STRIDE = 8 };

// flist entry: pixel index (14 bits) | three 6-bit offsets from the integer mask centre (biased +32)
CTK_DEV uint32_t pack_entry(int p, int o0, int o1, int o2) {
  return (uint32_t)p | ((uint32_t)(o0 + 32) << 14) | ((uint32_t)(o1 + 32) << 20) |
         ((uint32_t)(o2 + 32) << 26);
}
CTK_DEV int entry_pixel(uint32_t e) { return (int)(e & 0x3fffu); }
CTK_DEV int entry_off(uint32_t e, int k) { return (int)((e >> (14 + 6 * k)) & 0x3fu) - 32; }
// large-cluster variant: 32-bit pixel index | three 8-bit offsets (biased +128)
CTK_DEV uint64_t pack_entry64(int p, int o0, int o1, int o2) {
  return (uint64_t)(uint32_t)p | ((uint64_t)(uint32_t)(o0 + 128) << 32) |
         ((uint64_t)(uint32_t)(o1 + 128) << 40) | ((uint64_t)(uint32_t)(o2 + 128) << 48);
}
CTK_DEV int entry_pixel(uint64_t e) { return (int)(uint32_t)(e & 0xffffffffu); }
CTK_DEV int entry_off(uint64_t e, int k) { return (int)((e >> (32 + 8 * k)) & 0xffu) - 128; }
template <bool BIG> struct EntryOf { typedef uint32_t type; typedef uint16_t rc_type; };
template <> struct EntryOf<true> { typedef uint64_t type; typedef uint32_t rc_type; };
CTK_DEV uint16_t rc_pack(int r, int c, uint16_t) { return (uint16_t)(r | (c << 8)); }
CTK_DEV uint32_t rc_pack(int r, int c, uint32_t) { return (uint32_t)r | ((uint32_t)c << 16); }
CTK_DEV int rc_row(uint16_t v) { return v & 0xff; }
CTK_DEV int rc_col(uint16_t v) { return v >> 8; }
CTK_DEV int rc_row(uint32_t v) { return (int)(v & 0xffffu); }
CTK_DEV int rc_col(uint32_t v) { return (int)(v >> 16); }

// ------------------------------------------------------------------------------------------------
// kernel arguments
// ------------------------------------------------------------------------------------------------
struct BatchArgs {
  ctk_problem_t prob;
  const void* const* frames;
  int64_t shape[3];
  const double* frame_max;
  int n_work;
  const int32_t* work_ids;
  const int32_t* cluster_frame;
  const int32_t* cluster_offset;
  const double* params_in;
  const double* lo_in;
  const double* hi_in;
  double* params_out;
  double* cost_out;
  int32_t* status_out;
  int32_t* stats_out;        // [n_clusters, CTK_STATS] counters, see ctk.h
  int32_t* counter;
  int32_t* overflow;         // optional [1 + overflow_cap]: ids of clusters that ended TOO_LARGE
  int overflow_cap;
  const int32_t* n_work_dev; // optional device-side work count (<= n_work)
  char* big_workspace;       // BIG kernels: one slice of lay.total bytes per block
  int big_blocks;            // number of slices
  // global-level fits (ctk_global_pass): see ClusterSolver::run_global
  const double* mask_centres;  // optional [n_features, ndim]: centres of the pixel sets
  const double* global_step;   // phase 2: step of the global unknowns [G]
  double* global_accum;        // [CTK_GLOBAL_HEADER + G + G (G + 1) / 2] sums over the clusters
  double global_norm, global_lambda;
  int global_phase, use_newton;
  Layout lay;
};

// ------------------------------------------------------------------------------------------------
// cold helpers (not inlined): bounds from the tables, distance constraints
// ------------------------------------------------------------------------------------------------
// fitfunc.py:538-551; fmax/fmin skip NaN exactly like np.fmax/np.fmin, all-NaN -> unbounded
CTK_COLD double bound_from_tables(double p, double diff, double rel, double ab, int upper) {
  if (upper) {
    double v = fmin(fmin(p + diff, p * (1. + rel)), ab);
    return v == v ? v : INFINITY;
  }
  double v = fmax(fmax(p - diff, p * (1. - rel)), ab);
  return v == v ? v : -INFINITY;
}

// Distance constraints (constraints.py:59-99).  Constraint j couples features (p, q): dimer (0,1);
// trimer (0,1), (1,2), (0,2); tetramer see con_pair.  con[0..5] multipliers, con[6..8] distances.
struct ConView {
  const double* x;
  const int* cv;        // [n, P] variable index of (feature, column)
  const double* con;
  int n, P, nd, n_con;
  double w;             // penalty weight
};
CTK_COLD int con_pos_var(const ConView v, int k, int i) { return v.cv[i * v.P + 2 + k]; }
// squared distance of features p, q in units of the constraint distance
CTK_COLD double con_dist2(const ConView v, int p, int q) {
  double s = 0.;
  for (int k = 0; k < v.nd; ++k) {
    double d = (v.x[con_pos_var(v, k, p)] - v.x[con_pos_var(v, k, q)]) / v.con[6 + k];
    s += d * d;
  }
  return s;
}
// Feature pair of constraint j.  dimer (0,1); trimer (0,1),(1,2),(0,2) (constraints.py:79-83);
// tetramer in 3D: all six pairs (constraints.py:117-125); tetramer in 2D: the four SHORTEST of the
// six pair distances, i.e. the sides of the square (constraints.py:102-114).
CTK_COLD void con_pair(const ConView v, int j, int* p, int* q) {
  const int pa[6] = {0, 1, 0, 1, 0, 2}, pb[6] = {1, 2, 2, 3, 3, 3};
  if (v.n == 4 && v.nd == 2) {
    double d[6];
    int order[6];
    for (int a = 0; a < 6; ++a) { d[a] = con_dist2(v, pa[a], pb[a]); order[a] = a; }
    for (int a = 1; a < 6; ++a)                      // insertion sort, stable
      for (int b = a; b > 0 && d[order[b]] < d[order[b - 1]]; --b) {
        int t = order[b]; order[b] = order[b - 1]; order[b - 1] = t;
      }
    j = order[j];
  }
  *p = pa[j];
  *q = pb[j];
}
CTK_COLD double con_value(const ConView v, int j) {
  int p, q;
  con_pair(v, j, &p, &q);
  return 1. - con_dist2(v, p, q);
}
CTK_COLD double con_penalty(const ConView v) {
  double out = 0.;
  for (int j = 0; j < v.n_con; ++j) {
    double c = con_value(v, j);
    out += v.con[j] * c + 0.5 * v.w * c * c;
  }
  return out;
}
CTK_COLD double con_violation(const ConView v) {
  double out = 0.;
  for (int j = 0; j < v.n_con; ++j) out = fmax(out, fabs(con_value(v, j)));
  return out;
}
// gradient of constraint j: entry u (< 2 nd) belongs to variable *idx, value *g
CTK_COLD void con_grad(const ConView v, int j, int u, int* idx, double* g) {
  int p, q;
  con_pair(v, j, &p, &q);
  const int k = u >> 1;
  const double dist = v.con[6 + k];
  const double d = (v.x[con_pos_var(v, k, p)] - v.x[con_pos_var(v, k, q)]) / (dist * dist);
  *idx = con_pos_var(v, k, (u & 1) ? q : p);
  *g = (u & 1) ? 2. * d : -2. * d;
}
// add w A^T A to the packed (column-major lower, column starts cs) matrix Kp and
// -(mu + w c) A^T to rhs; one lane calls this
template <class Real>
CTK_COLD void con_add_rows(const ConView v, Real* Kp, const int* cs, double* rhs) {
  for (int j = 0; j < v.n_con; ++j) {
    const double lam = v.con[j] + v.w * con_value(v, j);
    for (int a = 0; a < 2 * v.nd; ++a) {
      int ia; double ga;
      con_grad(v, j, a, &ia, &ga);
      rhs[ia] -= lam * ga;
      for (int b = 0; b < 2 * v.nd; ++b) {
        int ib; double gb;
        con_grad(v, j, b, &ib, &gb);
        if (ib <= ia) Kp[cs[ib] + ia - ib] += (Real) (v.w * ga * gb);
      }
    }
  }
}
// -0.5 w sum_j (A_j s)^2: the constraint rows' share of the predicted decrease
CTK_COLD double con_quadratic(const ConView v, const double* s) {
  double out = 0.;
  for (int j = 0; j < v.n_con; ++j) {
    double as = 0.;
    for (int a = 0; a < 2 * v.nd; ++a) {
      int ia; double ga;
      con_grad(v, j, a, &ia, &ga);
      as += ga * s[ia];
    }
    out -= 0.5 * v.w * as * as;
  }
  return out;
}

// compile-time configuration of a kernel instance
template <class Real_, int ND_, bool ISO_, int FAM_, bool SZ_, bool EX_, bool BIG_ = false,
          bool EXTRA_ = true>
struct Config {
  typedef Real_ Real;
  // EXTRA: the instance carries the distance constraints and the lowpass (see CTK_NCON above)
  static const bool EXTRA = EXTRA_;
  // BIG: clusters of more than 32 features.  Same algorithm, but the per-cluster arrays live in a
  // global-memory workspace instead of shared memory, the per-pixel feature mask has several
  // words, and the packed indices are wider.
  static const bool BIG = BIG_;
  static const int ND = ND_;
  static const bool ISO = ISO_;
  static const int FAM = FAM_;
  static const bool SZ = SZ_;   // size column(s) carry derivatives
  static const bool EX = EX_;   // extra column (thickness / disc_size) carries a derivative
  static const int NS = ISO_ ? 1 : ND_;
  static const int NE = (FAM_ == CTK_FAMILY_GAUSS) ? 0 : 1;
  static const int P = 2 + ND_ + NS + NE;
  static const int LD = 1 + ND_ + (SZ_ ? NS : 0) + ((EX_ && NE) ? 1 : 0);   // derivative slots
  static const int LT = LD * (LD + 1) / 2;
};

template <class C>
struct ClusterSolver {
  typedef typename C::Real Real;
  typedef typename EntryOf<C::BIG>::type Entry;
  typedef typename EntryOf<C::BIG>::rc_type RcT;
  enum { ND = C::ND, P = C::P, LD = C::LD, LT = C::LT, NS = C::NS };

  // Kernel arguments BY VALUE: with every access at a compile-time index the compiler keeps them
  // in the constant bank (a reference would turn each read into a generic load).
  const BatchArgs a;
  int lane;
  // The cluster's slice of shared memory.  On the device it is always addressed as an offset from
  // the kernel's extern __shared__ array, so that the compiler emits shared-space loads/stores
  // (LDS/STS with 32-bit addresses) instead of generic ones.
#ifdef CTK_EMUL
  char* sm_;
  CTK_DEV char* slice() const { return sm_; }
#else
  uint32_t sm_off;
  CTK_DEV char* slice() const {
    if (C::BIG) return a.big_workspace + (size_t) blockIdx.x * (size_t) a.lay.total;
    extern __shared__ __align__(128) char ctk_smem[];
    return ctk_smem + sm_off;
  }
#endif

  // cluster
  int n, feat0, V, V_loc, M, npairs;   // V_loc: unknowns private to the cluster (globals come last)
  int blo[3], bdim[3];
  int cached_V;                  // V the packed-index tables were built for (-1 = none)
  int shared_columns;            // parameter columns (besides background) shared within the cluster
  const void* frame;
  double fmax_;
  int evals, accums, grad_accums, outers, n_entries, n_pair_entries;
  // residual statistics of the last evaluate()
  double sum_r, n_valid;
  // augmented Lagrangian (multipliers and distances live in shared memory, see CON())
  int n_con;
  double pen_w;
  // accumulate() adds the second-order term of the Hessian (gauss family, see second_order())
  bool newton;

#ifdef CTK_EMUL
  CTK_DEV ClusterSolver(const BatchArgs& args, char* smem)
      : a(args), lane(lane_id()), sm_(smem), cached_V(-1) { init_column_tables(); }
#else
  CTK_DEV ClusterSolver(const BatchArgs& args, uint32_t smem_offset)
      : a(args), lane(lane_id()), sm_off(smem_offset), cached_V(-1) { init_column_tables(); }
#endif

  // ---- typed views ------------------------------------------------------------------------------
  CTK_DEV double* dvec(int off) const { return reinterpret_cast<double*>(slice() + off); }
  CTK_DEV double* X() const { return dvec(a.lay.o_x); }
  CTK_DEV double* XT() const { return dvec(a.lay.o_xt); }
  CTK_DEV double* X0() const { return dvec(a.lay.o_x0); }
  CTK_DEV double* XB() const { return dvec(a.lay.o_xb); }
  CTK_DEV double* LO() const { return dvec(a.lay.o_lo); }
  CTK_DEV double* HI() const { return dvec(a.lay.o_hi); }
  CTK_DEV double* RHS() const { return dvec(a.lay.o_rhs); }
  CTK_DEV double* D() const { return dvec(a.lay.o_d); }
  CTK_DEV double* DG() const { return dvec(a.lay.o_dg); }
  CTK_DEV int* ACT() const { return reinterpret_cast<int*>(slice() + a.lay.o_act); }
  // normal matrix, its factor and the factor's inverse diagonal are kept in the pixel arithmetic
  // type: a step of limited accuracy still converges to the same point (the gradient decides)
  CTK_DEV Real* Hm() const { return reinterpret_cast<Real*>(slice() + a.lay.o_H); }
  CTK_DEV Real* Lm() const { return reinterpret_cast<Real*>(slice() + a.lay.o_L); }
  CTK_DEV Real* IDG() const { return reinterpret_cast<Real*>(slice() + a.lay.o_idg); }
  CTK_DEV int* CS() const { return reinterpret_cast<int*>(slice() + a.lay.o_cs); }
  CTK_DEV RcT* RC() const { return reinterpret_cast<RcT*>(slice() + a.lay.o_rc); }
  CTK_DEV int NW() const { return C::BIG ? a.lay.mask_words : 1; }
  CTK_DEV int* CV() const { return reinterpret_cast<int*>(slice() + a.lay.o_cv); }
  CTK_DEV int* SIDX() const { return reinterpret_cast<int*>(slice() + a.lay.o_sidx); }
  CTK_DEV double* CON() const { return dvec(a.lay.o_con); }        // mu[0..5], dist[6..8]
  CTK_DEV int* CMODE() const { return reinterpret_cast<int*>(slice() + a.lay.o_cmode); }
  CTK_DEV int* CBASE() const { return reinterpret_cast<int*>(slice() + a.lay.o_cbase); }
  CTK_DEV double* CTAB() const { return dvec(a.lay.o_ctab); }
  CTK_DEV double* TAPS() const { return dvec(a.lay.o_taps); }
  CTK_DEV double* MC() const { return dvec(a.lay.o_mc); }
  CTK_DEV int* FI() const { return reinterpret_cast<int*>(slice() + a.lay.o_fi); }
  CTK_DEV Real* FR() const { return reinterpret_cast<Real*>(slice() + a.lay.o_fr); }
  CTK_DEV double* TAB() const { return dvec(a.lay.o_tab); }
  CTK_DEV Real* FE() const { return reinterpret_cast<Real*>(slice() + a.lay.o_fe); }
  CTK_DEV Real* PR() const { return reinterpret_cast<Real*>(slice() + a.lay.o_pr); }
  CTK_DEV uint32_t* PBITS() const { return reinterpret_cast<uint32_t*>(slice() + a.lay.o_pbits); }
  CTK_DEV uint32_t* PCRD() const { return reinterpret_cast<uint32_t*>(slice() + a.lay.o_pcrd); }
  CTK_DEV Entry* FLIST() const { return reinterpret_cast<Entry*>(slice() + a.lay.o_flist); }
  CTK_DEV uint32_t* PAIRS() const { return reinterpret_cast<uint32_t*>(slice() + a.lay.o_pairs); }
  CTK_DEV int* PHDR() const { return reinterpret_cast<int*>(slice() + a.lay.o_phdr); }

  CTK_DEV int mode(int col) const { return a.prob.modes[col]; }
  // variable index of (column, feature), -1 when the column is constant
  CTK_DEV int var_of(int col, int i) const { return CV()[i * P + col]; }
  // packed index of entry (r, c), r >= c, of a lower triangle stored column by column
  CTK_DEV int pk(int r, int c) const { return CS()[c] + r - c; }
  CTK_DEV int pk_sym(int u, int v) const { return u >= v ? pk(u, v) : pk(v, u); }
  // column of derivative slot `s`
  CTK_DEV static int slot_col(int s) {
    if (s <= ND) return 1 + s;                       // signal, positions
    if (C::SZ && s < 1 + ND + NS) return 2 + ND + (s - 1 - ND);
    return 2 + ND + NS;                              // extra
  }

  // pixel values are staged in shared memory in their native width
  // type of the staged values: the frame's own, or the arithmetic type when they are filtered
  CTK_DEV int staged_dtype() const {
    return CTK_LOWPASS ? (sizeof(Real) == 4 ? CTK_PIXEL_F32 : CTK_PIXEL_F64) : a.prob.pixel_dtype;
  }
  CTK_DEV double frame_value(int64_t i) const {
    switch (a.prob.pixel_dtype) {
      case CTK_PIXEL_U8: return (double) reinterpret_cast<const uint8_t*>(frame)[i];
      case CTK_PIXEL_U16: return (double) reinterpret_cast<const uint16_t*>(frame)[i];
      case CTK_PIXEL_F32: return (double) reinterpret_cast<const float*>(frame)[i];
      case CTK_PIXEL_F64: return reinterpret_cast<const double*>(frame)[i];
      case CTK_PIXEL_I16: return (double) reinterpret_cast<const int16_t*>(frame)[i];
      default: return (double) reinterpret_cast<const int32_t*>(frame)[i];
    }
  }
  // Lowpass-filtered value of box pixel c (refine.py:36-40 -> preprocessing.py:39-44): gaussian taps
  // along axis 0 first, then axis 1 (then 2), over the cluster's BOX with zeros beyond its edge, then
  // the threshold cut.  Float64 like the reference.
  CTK_DEV double lowpass_value(const int (&c)[3]) const {
    const double* taps = TAPS();
    int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}, hw[3] = {-1, -1, -1};
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      hw[k] = k == 0 ? a.prob.lowpass_half[0] : (k == 1 ? a.prob.lowpass_half[1] : a.prob.lowpass_half[2]);
      if (hw[k] >= 0) { lo[k] = max(-hw[k], -c[k]); hi[k] = min(hw[k], bdim[k] - 1 - c[k]); }
    }
    double acc2 = 0.;
    for (int d2 = lo[2]; d2 <= hi[2]; ++d2) {
      double acc1 = 0.;
      for (int d1 = lo[1]; d1 <= hi[1]; ++d1) {
        double acc0 = 0.;
        for (int d0 = lo[0]; d0 <= hi[0]; ++d0) {
          const int d[3] = {d0, d1, d2};
          int64_t gi = 0;
#pragma unroll
          for (int k = 0; k < ND; ++k) gi = gi * a.shape[k] + (blo[k] + c[k] + d[k]);
          acc0 += (hw[0] >= 0 ? taps[d0 + hw[0]] : 1.) * frame_value(gi);
        }
        acc1 += (ND > 1 && hw[1] >= 0 ? taps[CTK_MAX_TAPS + d1 + hw[1]] : 1.) * acc0;
      }
      acc2 += (ND > 2 && hw[2] >= 0 ? taps[2 * CTK_MAX_TAPS + d2 + hw[2]] : 1.) * acc1;
    }
    return acc2 > a.prob.lowpass_threshold ? acc2 : 0.;
  }
  // The pixel of a box position is FETCHED at the top of the box walk (bits of its native width in
  // a 64-bit register) and STORED after the coverage tests, so that the global-memory latency hides
  // behind them instead of stalling the warp once per 32 box pixels (7 % of all stall samples).
  CTK_DEV uint64_t fetch_pixel(int64_t idx) const {
    if (CTK_LOWPASS) return 0;
    switch (a.prob.pixel_dtype) {
      case CTK_PIXEL_U8: return reinterpret_cast<const uint8_t*>(frame)[idx];
      case CTK_PIXEL_U16: case CTK_PIXEL_I16: return reinterpret_cast<const uint16_t*>(frame)[idx];
      case CTK_PIXEL_F64: return reinterpret_cast<const uint64_t*>(frame)[idx];
      default: return reinterpret_cast<const uint32_t*>(frame)[idx];
    }
  }
  CTK_DEV void stage_pixel(uint64_t raw, int p, const int (&c)[3]) const {
    void* dst = slice() + a.lay.o_pval;
    if (CTK_LOWPASS) {
      reinterpret_cast<Real*>(dst)[p] = (Real) lowpass_value(c);
      return;
    }
    switch (a.prob.pixel_dtype) {
      case CTK_PIXEL_U8: reinterpret_cast<uint8_t*>(dst)[p] = (uint8_t) raw; break;
      case CTK_PIXEL_U16: case CTK_PIXEL_I16: reinterpret_cast<uint16_t*>(dst)[p] = (uint16_t) raw; break;
      case CTK_PIXEL_F64: reinterpret_cast<uint64_t*>(dst)[p] = raw; break;
      default: reinterpret_cast<uint32_t*>(dst)[p] = (uint32_t) raw; break;
    }
  }
  CTK_DEV Real pixel_value(int p) const {
    const void* src = slice() + a.lay.o_pval;
    switch (staged_dtype()) {
      case CTK_PIXEL_U8: return (Real) reinterpret_cast<const uint8_t*>(src)[p];
      case CTK_PIXEL_U16: return (Real) reinterpret_cast<const uint16_t*>(src)[p];
      case CTK_PIXEL_F32: return (Real) reinterpret_cast<const float*>(src)[p];
      case CTK_PIXEL_F64: return (Real) reinterpret_cast<const double*>(src)[p];
      case CTK_PIXEL_I16: return (Real) reinterpret_cast<const int16_t*>(src)[p];
      default: return (Real) reinterpret_cast<const int32_t*>(src)[p];
    }
  }

  // Per-launch column tables in shared memory (modes and the bounds tables), filled once per warp:
  // the per-cluster set-up then runs as plain loops over (feature, column) instead of code
  // unrolled per column.
  CTK_DEV void init_column_tables() {
    int* cmode = CMODE();
    double* ctab = CTAB();
#pragma unroll
    for (int c = 0; c < P; ++c) {
      if (lane == 0) {
        cmode[c] = a.prob.modes[c];
        ctab[c * 6 + 0] = a.prob.bounds_diff[0][c]; ctab[c * 6 + 1] = a.prob.bounds_rel[0][c];
        ctab[c * 6 + 2] = a.prob.bounds_abs[0][c];  ctab[c * 6 + 3] = a.prob.bounds_diff[1][c];
        ctab[c * 6 + 4] = a.prob.bounds_rel[1][c];  ctab[c * 6 + 5] = a.prob.bounds_abs[1][c];
      }
    }
    if (CTK_LOWPASS) {
      // taps of trackpy.masks.gaussian_kernel(sigma, truncate=4): exp(x^2 / (-2 sigma^2)) on
      // [-lw, lw], normalised; CTK_MAX_TAPS <= 2 lanes' worth of entries per axis
      double* taps = TAPS();
#pragma unroll
      for (int k = 0; k < ND; ++k) {
        const int hw = k == 0 ? a.prob.lowpass_half[0] : (k == 1 ? a.prob.lowpass_half[1] : a.prob.lowpass_half[2]);
        const double sg = k == 0 ? a.prob.lowpass_sigma[0] : (k == 1 ? a.prob.lowpass_sigma[1] : a.prob.lowpass_sigma[2]);
        if (hw < 0) continue;
        double part = 0.;
#pragma unroll 1
        for (int t = lane; t <= 2 * hw; t += CTK_WARP) {
          const double x = (double) (t - hw);
          const double w = exp(x * x / (-2. * sg * sg));
          taps[k * CTK_MAX_TAPS + t] = w;
          part += w;
        }
        const double total = warp_sum(part);
        warp_sync();
#pragma unroll 1
        for (int t = lane; t <= 2 * hw; t += CTK_WARP) taps[k * CTK_MAX_TAPS + t] /= total;
      }
    }
    shared_columns = 0;
#pragma unroll
    for (int c = 1; c < P; ++c)
      shared_columns += (mode(c) == CTK_MODE_CLUSTER || mode(c) == CTK_MODE_GLOBAL) ? 1 : 0;
    warp_sync();
  }

  CTK_DEV_BIG int setup_variables() {
    // variable numbering: columns in order; a 'var' column takes n entries, a 'cluster' column one
    int* cv = CV();
    const int* cmode = CMODE();
    int* cbase = CBASE();
    const double* ctab = CTAB();
    int v = 0;
    for (int c = 0; c < P; ++c) {
      const int m = cmode[c];
      if (m == CTK_MODE_GLOBAL) continue;
      if (lane == 0) cbase[c] = v;
      v += m == CTK_MODE_VAR ? n : (m == CTK_MODE_CLUSTER ? 1 : 0);
    }
    V_loc = v;
    // columns shared by ALL features of the table ('global', refine.py:319-332): one unknown each,
    // numbered after the cluster's own so that eliminating the first V_loc leaves their Schur block
    for (int c = 0; c < P; ++c) {
      if (cmode[c] != CTK_MODE_GLOBAL) continue;
      if (lane == 0) cbase[c] = v;
      v += 1;
    }
    V = v;
    if (V > a.lay.v_max || V > (C::BIG ? 65535 : 255)) return CTK_FAIL_TOO_LARGE;
    warp_sync();
    const double* pin = a.params_in + (int64_t)feat0 * P;
    const bool tables = a.lo_in == nullptr;            // bounds from the problem's tables
    const double* lin = tables ? nullptr : a.lo_in + (int64_t)feat0 * P;
    const double* hin = tables ? nullptr : a.hi_in + (int64_t)feat0 * P;
    double *x0 = X0(), *lo = LO(), *hi = HI();
    bool bad = false;
#pragma unroll 1
    for (int t = lane; t < n * P; t += CTK_WARP) {
      const int i = t / P, c = t - i * P;
      const int m = cmode[c];
      const double p = pin[t];
      bad |= !finite_d(p);
      cv[t] = m == CTK_MODE_VAR ? cbase[c] + i
                                : ((m == CTK_MODE_CLUSTER || m == CTK_MODE_GLOBAL) ? cbase[c] : -1);
      if (m == CTK_MODE_VAR) {
        const int vi = cbase[c] + i;
        const double* tb = ctab + c * 6;
        x0[vi] = p;
        lo[vi] = tables ? bound_from_tables(p, tb[0], tb[1], tb[2], 0) : lin[t];
        hi[vi] = tables ? bound_from_tables(p, tb[3], tb[4], tb[5], 1) : hin[t];
      }
    }
    if (warp_any(bad)) return CTK_FAIL_NONFINITE;
#pragma unroll 1
    for (int c = lane; c < P; c += CTK_WARP) {
      if (cmode[c] != CTK_MODE_CLUSTER && cmode[c] != CTK_MODE_GLOBAL) continue;
      const double* tb = ctab + c * 6;                  // shared entry: mean start, widest bound
      double s = 0., l = INFINITY, h = -INFINITY;
      for (int i = 0; i < n; ++i) {
        const double p = pin[i * P + c];
        s += p;
        l = fmin(l, tables ? bound_from_tables(p, tb[0], tb[1], tb[2], 0) : lin[i * P + c]);
        h = fmax(h, tables ? bound_from_tables(p, tb[3], tb[4], tb[5], 1) : hin[i * P + c]);
      }
      const int vi = cbase[c];
      x0[vi] = s / n; lo[vi] = l; hi[vi] = h;
      if (cmode[c] == CTK_MODE_GLOBAL) {               // one value for the whole table, bounds: the host's job
        x0[vi] = pin[c]; lo[vi] = -INFINITY; hi[vi] = INFINITY;
      }
    }
    warp_sync();
    bad = false;
    CTK_FOR_V(v2) {
      bad |= !(lo[v2] <= hi[v2]);
      x0[v2] = fmin(fmax(x0[v2], lo[v2]), hi[v2]);     // scipy clips the start into the box
    }
    if (warp_any(bad)) return CTK_FAIL_BOUNDS;
    // packed-index tables (depend on V only; consecutive clusters of a launch mostly share V)
    if (V != cached_V) {
      int* cs = CS();
      RcT* rc = RC();
#pragma unroll 1
      for (int c = lane; c <= V; c += CTK_WARP) cs[c] = c * V - c * (c - 1) / 2;
      warp_sync();
#pragma unroll 1
      for (int t = lane; t < V * V; t += CTK_WARP) {
        int r = t / V, c = t - r * V;
        if (r >= c) rc[cs[c] + r - c] = rc_pack(r, c, RcT());
      }
      cached_V = V;
    }
    warp_sync();
    // scatter targets of every feature's local block: [LT] block entries, [LD] products with the
    // background column, [LD] right-hand side entries
    int* sidx = SIDX();
    const int vb = cv[0];
#pragma unroll 1
    for (int t = lane; t < n * (LT + 2 * LD); t += CTK_WARP) {
      const int i = t / (LT + 2 * LD), k = t - i * (LT + 2 * LD);
      int target = -1;
      if (k < LT) {
        int u = 0, rem = k;                       // k = u (u + 1) / 2 + w, w <= u
        while (rem > u) { rem -= u + 1; ++u; }
        const int vu = cv[i * P + slot_col(u)], vw = cv[i * P + slot_col(rem)];
        if (vu >= 0 && vw >= 0) target = pk_sym(vu, vw);
      } else if (k < LT + LD) {
        const int vu = cv[i * P + slot_col(k - LT)];
        if (vu >= 0 && vb >= 0) target = pk_sym(vu, vb);
      } else {
        target = cv[i * P + slot_col(k - LT - LD)];
      }
      sidx[i * a.lay.sidx_stride + k] = target;
    }
    warp_sync();
    return CTK_OK;
  }

  // ellipse test of box pixel c against feature i from the separable tables (refine.py:43-44)
  CTK_DEV bool covers(int i, const int (&c)[3], const int* fi, const double* tab) const {
    const int* f = fi + i * FI_STRIDE;
    const double* tb = tab + i * a.lay.tab_stride;
    double s = 0.;
    bool in = true;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      int e = c[k] - f[FI_TS + k];
      in = in && (e >= 0) && (e < a.lay.tab_len[k]);
      if (in) s = (k == 0) ? tb[e] : dadd(s, tb[e]);
      tb += a.lay.tab_len[k];
    }
    return in && s <= 1.0;
  }
  CTK_DEV static bool covered(const uint32_t* pbits, int p, int i, int nw) {
    if (!C::BIG) return (pbits[p] >> i) & 1u;
    return (pbits[(size_t) p * nw + (i >> 5)] >> (i & 31)) & 1u;
  }
  CTK_DEV static void store_entry(uint32_t* dst, int p, int o0, int o1, int o2) {
    *dst = pack_entry(p, o0, o1, o2);
  }
  CTK_DEV static void store_entry(uint64_t* dst, int p, int o0, int o1, int o2) {
    *dst = pack_entry64(p, o0, o1, o2);
  }

  // ---- pixel set (refine.py:28-58, masks.py:30-68) ----------------------------------------------
  CTK_DEV_BIG int build_pixels() {
    const double* mc = MC();
    int* fi = FI();
    // integer centres (round half to even) and the in-bounds test of masks.py:42-46
    int mn[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, mx[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
#pragma unroll 1
    for (int i = lane; i < n; i += CTK_WARP) {
      bool inb = true;
      int ci[3] = {0, 0, 0};
#pragma unroll
      for (int k = 0; k < ND; ++k) {
        double r = rint(mc[i * 3 + k]);
        r = fmin(fmax(r, -1.0e9), 1.0e9);
        ci[k] = (int) r;
        fi[i * FI_STRIDE + FI_CI + k] = ci[k];
        inb = inb && (ci[k] >= -a.prob.radius[k]) && (ci[k] < (int) a.shape[k] + a.prob.radius[k]);
      }
      fi[i * FI_STRIDE + FI_INB] = inb ? 1 : 0;
      if (inb) {
#pragma unroll
        for (int k = 0; k < ND; ++k) { mn[k] = min(mn[k], ci[k]); mx[k] = max(mx[k], ci[k]); }
      }
    }
    int64_t total = 1;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      int lo_k = warp_min_i(mn[k]), hi_k = warp_max_i(mx[k]);
      if (lo_k == INT32_MAX) return CTK_FAIL_OUT_OF_IMAGE;
      blo[k] = max(0, lo_k - a.prob.radius[k]);
      int bhi = min((int) a.shape[k], hi_k + a.prob.radius[k] + 1);
      bdim[k] = bhi - blo[k];
      total *= bdim[k];
    }
    if (total <= 0 || total > (int64_t) 1 << 30) return CTK_FAIL_TOO_LARGE;
#pragma unroll
    for (int k = 0; k < ND; ++k) if (!C::BIG && bdim[k] > 1023) return CTK_FAIL_TOO_LARGE;
    warp_sync();
    // separable tables: tab[i][k][e] = (((ts + e) - (c - origin)) / r)^2, float64, numpy's order
    double* tab = TAB();
#pragma unroll 1
    for (int i = lane; i < n; i += CTK_WARP) {
#pragma unroll
      for (int k = 0; k < ND; ++k) {
        double crel = dsub(mc[i * 3 + k], (double) blo[k]);
        double s = floor(crel - (double) a.prob.radius[k]);
        s = fmin(fmax(s, -1.0e9), 1.0e9);
        fi[i * FI_STRIDE + FI_TS + k] = (int) s;
      }
    }
    warp_sync();
    {
      int tab_off = 0;
#pragma unroll
      for (int k = 0; k < ND; ++k) {
        const int len = a.lay.tab_len[k];
#pragma unroll 1
        for (int t = lane; t < n * len; t += CTK_WARP) {
          const int i = t / len, e = t - i * len;
          const double crel = dsub(mc[i * 3 + k], (double) blo[k]);
          const double idx = (double) (fi[i * FI_STRIDE + FI_TS + k] + e);
          const double q = ddiv(dsub(idx, crel), (double) a.prob.radius[k]);
          tab[i * a.lay.tab_stride + tab_off + e] = dmul(q, q);
        }
        tab_off += len;
      }
    }
    warp_sync();
    // walk the box in C order, ballot-compact the union
    uint32_t *pbits = PBITS(), *pcrd = PCRD();
    const int nw = NW();
    uint32_t* tmpw = reinterpret_cast<uint32_t*>(slice() + a.lay.o_tmpw) + lane * nw;   // BIG only
    int count = 0;
    const int itotal = (int) total;
    for (int q0 = 0; q0 < itotal; q0 += CTK_WARP) {
      int q = q0 + lane;
      uint32_t bits = 0u;            // the mask itself (one word), or "any word set" (BIG)
      int c[3] = {0, 0, 0};
      uint64_t raw = 0;
      if (q < itotal) {
        int rem = q;
#pragma unroll
        for (int k = ND - 1; k >= 0; --k) { c[k] = rem % bdim[k]; rem /= bdim[k]; }
        int64_t gi = 0;
#pragma unroll
        for (int k = 0; k < ND; ++k) gi = gi * a.shape[k] + (blo[k] + c[k]);
        raw = fetch_pixel(gi);
        if (!C::BIG) {
          for (int i = 0; i < n; ++i)
            if (covers(i, c, fi, tab)) bits |= (1u << i);
        } else {
          for (int w = 0; w < nw; ++w) {
            uint32_t word = 0u;
            const int i1 = min(n, 32 * w + 32);
            for (int i = 32 * w; i < i1; ++i)
              if (covers(i, c, fi, tab)) word |= (1u << (i & 31));
            tmpw[w] = word;
            bits |= word;
          }
        }
      }
      uint32_t ball = ballot(bits != 0u);
      if (bits != 0u) {
        int pos = count + popc(ball & lanemask_lt());
        if (pos < a.lay.m_cap) {
          stage_pixel(raw, pos, c);
          if (!C::BIG) {
            pbits[pos] = bits;
            pcrd[pos] = (uint32_t) c[0] | ((uint32_t) c[1] << 10) | ((uint32_t) c[2] << 20);
          } else {
            for (int w = 0; w < nw; ++w) pbits[(size_t) pos * nw + w] = tmpw[w];
            pcrd[pos] = (uint32_t) q;              // linear box index, decoded when needed
          }
        }
      }
      count += popc(ball);
    }
    M = count;
    if (M > a.lay.m_cap || M == 0) return M == 0 ? CTK_FAIL_OUT_OF_IMAGE : CTK_FAIL_TOO_LARGE;
    warp_sync();
    // Per-feature pixel lists, in union order, and -- feature by feature -- the lists of the pixels
    // a feature shares with every earlier one: entries (t_i | t_j << 16) grouped per pair.  While
    // feature j's list is written, the position t_j of every pixel in it goes to a per-pixel rank
    // array (in the model-value cache, unused until the first evaluate()), so that a shared pixel
    // found through feature i's list gets its t_j by one load instead of a binary search.
    Entry* flist = FLIST();
    uint16_t* rank = reinterpret_cast<uint16_t*>(FE());
    uint32_t* pairs = PAIRS();
    int* phdr = PHDR();
    int np = 0, ptotal = 0;
    for (int j = 0; j < n; ++j) {
      const int* f_j = fi + j * FI_STRIDE;
      int oc[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) oc[k] = (k < ND) ? f_j[FI_CI + k] - blo[k] : 0;
      int cnt_j = 0;
      for (int p0 = 0; p0 < M; p0 += CTK_WARP) {
        int p = p0 + lane;
        bool has = (p < M) && covered(pbits, p, j, nw);
        uint32_t ball = ballot(has);
        if (has) {
          int t = cnt_j + popc(ball & lanemask_lt());
          rank[p] = (uint16_t) t;
          if (t < a.lay.f_cap) {
            uint32_t crd = pcrd[p];
            int o[3] = {0, 0, 0};
            if (!C::BIG) {
              o[0] = (int) (crd & 1023u); o[1] = (int) ((crd >> 10) & 1023u);
              o[2] = (int) ((crd >> 20) & 1023u);
            } else {
              int rem = (int) crd;
#pragma unroll
              for (int k = ND - 1; k >= 0; --k) { o[k] = rem % bdim[k]; rem /= bdim[k]; }
            }
            store_entry(flist + (size_t) j * a.lay.f_cap + t, p, o[0] - oc[0],
                        ND > 1 ? o[1] - oc[1] : 0, ND > 2 ? o[2] - oc[2] : 0);
          }
        }
        cnt_j += popc(ball);
      }
      if (lane == 0) fi[j * FI_STRIDE + FI_CNT] = cnt_j;
      if (cnt_j > a.lay.f_cap) return CTK_FAIL_TOO_LARGE;
      warp_sync();
      for (int i = 0; i < j; ++i) {
        const int* f_i = fi + i * FI_STRIDE;
        bool apart = false;
#pragma unroll
        for (int k = 0; k < ND; ++k)
          apart |= abs(f_i[FI_CI + k] - f_j[FI_CI + k]) > 2 * a.prob.radius[k] + 2;
        if (apart) continue;
        const int cnt_i = f_i[FI_CNT];
        int cnt = 0;
        for (int t0 = 0; t0 < cnt_i; t0 += CTK_WARP) {
          int t = t0 + lane;
          int p = 0;
          bool has = false;
          if (t < cnt_i) {
            p = entry_pixel(flist[(size_t) i * a.lay.f_cap + t]);
            has = covered(pbits, p, j, nw);
          }
          uint32_t ball = ballot(has);
          if (has) {
            int pos = ptotal + cnt + popc(ball & lanemask_lt());
            if (pos < a.lay.pair_cap) pairs[pos] = (uint32_t) t | ((uint32_t) rank[p] << 16);
          }
          cnt += popc(ball);
        }
        if (cnt > 0) {
          if (np >= a.lay.npair_cap || ptotal + cnt > a.lay.pair_cap) return CTK_FAIL_TOO_LARGE;
          if (lane == 0) {
            phdr[np * 4 + 0] = i; phdr[np * 4 + 1] = j; phdr[np * 4 + 2] = ptotal;
            phdr[np * 4 + 3] = cnt;
          }
          ++np;
          ptotal += cnt;
        }
      }
      warp_sync();                       // the rank array is rewritten by the next feature
    }
    npairs = np;
    n_pair_entries = ptotal;
    n_entries = 0;
    for (int i = 0; i < n; ++i) n_entries += fi[i * FI_STRIDE + FI_CNT];
    warp_sync();
    return CTK_OK;
  }

  // ---- per-feature constants of the pixel pass from the variable vector ------------------------
  CTK_DEV double value_of(const double* x, int col, int i) const {
    int v = var_of(col, i);
    return v < 0 ? a.params_in[(int64_t) (feat0 + i) * P + col] : x[v];
  }

  CTK_DEV void load_features(const double* x) {
    Real* fr = FR();
    const int* fi = FI();
#pragma unroll 1
    for (int i = lane; i < n; i += CTK_WARP) {
      Real* r = fr + i * FR_STRIDE;
      r[FR_S] = (Real) value_of(x, 1, i);
#pragma unroll
      for (int k = 0; k < ND; ++k) {
        r[FR_FRAC + k] = (Real) (value_of(x, 2 + k, i) - (double) fi[i * FI_STRIDE + FI_CI + k]);
        double size = value_of(x, 2 + ND + (C::ISO ? 0 : k), i);
        r[FR_IS + k] = (Real) (1.0 / size);
      }
      r[FR_EX] = (C::NE > 0) ? (Real) value_of(x, 2 + ND + NS, i) : (Real) 0;
    }
    warp_sync();
  }

  struct Feat { Real s, frac[3], is[3], ex; };
  CTK_DEV Feat feat(int i) const {
    const Real* r = FR() + i * FR_STRIDE;
    Feat f;
    f.s = r[FR_S];
#pragma unroll
    for (int k = 0; k < 3; ++k) { f.frac[k] = r[FR_FRAC + k]; f.is[k] = r[FR_IS + k]; }
    f.ex = r[FR_EX];
    return f;
  }

  // geometry of one (pixel, feature): q_k = (x_k - c_k)/size_k, r2 = sum q_k^2, d2 = pixel dist^2
  struct Geo { Real q[3], r2, d2; };
  CTK_DEV Geo geometry(Entry e, const Feat& f) const {
    Geo g;
    g.r2 = 0; g.d2 = 0;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      Real d = (Real) entry_off(e, k) - f.frac[k];
      g.q[k] = d * f.is[k];
      g.r2 += g.q[k] * g.q[k];
      g.d2 += d * d;
    }
    return g;
  }

  // model value; `drop` = the reference produces NaN here (pixel leaves every sum, fitfunc.py:449)
  CTK_DEV Real model_value(const Geo& g, const Feat& f, bool& drop) const {
    drop = false;
    const Real half_nd = (Real) (0.5 * ND);
    if (C::FAM == CTK_FAMILY_GAUSS) return fast_exp(-half_nd * g.r2);
    const bool safe_nan = g.d2 < (Real) 1;                 // the *_safe r2 variants, fitfunc.py:20-26
    if (C::FAM == CTK_FAMILY_RING) {
      drop = safe_nan;
      Real u = (sqrt(g.r2) - (Real) 1 + f.ex) / f.ex;
      return fast_exp(-half_nd * u * u);
    }
    // disc, fitfunc.py:121-131
    Real d = f.ex;
    if (d <= (Real) 0) { drop = safe_nan; return fast_exp(-half_nd * g.r2); }
    if (d >= (Real) 1) d = (Real) 0.999;
    if (safe_nan || !(g.r2 > d * d)) return (Real) 1;
    Real u = (sqrt(g.r2) - d) / ((Real) 1 - d);
    return fast_exp(-half_nd * u * u);
  }

  // derivatives of s*g wrt the LD slots, given the cached model value g
  CTK_DEV void model_derivs(const Geo& g, const Feat& f, Real gv, Real* m) const {
    const Real nd = (Real) ND;
    Real W;              // -2 * s * dg/dr2
    Real dex = 0;        // s * dg/dextra
    if (C::FAM == CTK_FAMILY_GAUSS) {
      W = f.s * nd * gv;
    } else if (C::FAM == CTK_FAMILY_RING) {
      Real rr = sqrt(g.r2), t = f.ex;
      Real u = (rr - (Real) 1 + t) / t;
      W = f.s * nd * gv * u / (rr * t);
      dex = f.s * gv * nd * u * (u - (Real) 1) / t;
    } else {
      Real d = f.ex;
      if (d <= (Real) 0) {
        W = f.s * nd * gv;
      } else {
        bool clamp = d >= (Real) 1;
        if (clamp) d = (Real) 0.999;
        if (g.d2 < (Real) 1 || !(g.r2 > d * d)) {
          W = 0;
        } else {
          Real rr = sqrt(g.r2), om = (Real) 1 - d;
          Real u = (rr - d) / om;
          W = f.s * nd * gv * u / (rr * om);
          dex = clamp ? (Real) 0 : f.s * gv * nd * u * ((Real) 1 - rr) / (om * om);
        }
      }
    }
    m[0] = gv;
#pragma unroll
    for (int k = 0; k < ND; ++k) m[1 + k] = W * g.q[k] * f.is[k];
    if (C::SZ) {
      if (C::ISO) m[1 + ND] = W * g.r2 * f.is[0];
      else {
#pragma unroll
        for (int k = 0; k < ND; ++k) m[1 + ND + k] = W * g.q[k] * g.q[k] * f.is[k];
      }
    }
    if (C::EX && C::NE) m[LD - 1] = dex;
  }

  // Second-order term of the Hessian of 0.5 sum r^2 for one (pixel, feature) of the gauss family:
  // acc[u (u + 1) / 2 + w] -= r * d2(s g)/d(slot u)d(slot w).  With g = exp(-E),
  // E = nd/2 sum_k (d_k / size_k)^2:  d2(s g)/dtheta dphi = s g (E_theta E_phi - E_theta,phi) and
  // d2(s g)/ds dtheta = -g E_theta.  The model is a SUM over features, so this term is block
  // diagonal per feature: it rides on the accumulators of J^T J at no extra reduction cost.
  // Gauss-Newton alone converges linearly at a rate set by the residual (clusters with a poor start
  // or a neighbour's light in the mask needed > 100 iterations, where SLSQP's BFGS needs ~20); with
  // the exact Hessian the damped iteration converges quadratically.
  CTK_DEV void second_order(const Geo& g, const Feat& f, Real gv, Real r, Real* acc) const {
    const Real nd = (Real) ND;
    const Real rg = r * gv, rG = rg * f.s;
    Real e[LD];
    e[0] = 0;
#pragma unroll
    for (int k = 0; k < ND; ++k) e[1 + k] = -nd * g.q[k] * f.is[k];
    if (C::SZ) {
      if (C::ISO) e[1 + ND] = -nd * g.r2 * f.is[0];
      else {
#pragma unroll
        for (int k = 0; k < ND; ++k) e[1 + ND + k] = -nd * g.q[k] * g.q[k] * f.is[k];
      }
    }
#pragma unroll
    for (int u = 1; u < LD; ++u) {
      acc[u * (u + 1) / 2] += rg * e[u];
#pragma unroll
      for (int w = 1; w <= u; ++w) {
        Real euw = 0;                                   // E_theta,phi
        if (u <= ND) {
          if (w == u) euw = nd * f.is[u - 1 < 3 ? u - 1 : 0] * f.is[u - 1 < 3 ? u - 1 : 0];
        } else if (C::ISO) {                            // u = the size slot
          if (w <= ND) euw = (Real) 2 * nd * g.q[w - 1 < 3 ? w - 1 : 0] * f.is[0] * f.is[0];
          else euw = (Real) 3 * nd * g.r2 * f.is[0] * f.is[0];
        } else {                                        // u = size slot of axis k
          const int k = u - 1 - ND >= 0 && u - 1 - ND < 3 ? u - 1 - ND : 0;
          if (w == 1 + k) euw = (Real) 2 * nd * g.q[k] * f.is[k] * f.is[k];
          else if (w == u) euw = (Real) 3 * nd * g.q[k] * g.q[k] * f.is[k] * f.is[k];
        }
        acc[u * (u + 1) / 2 + w] -= rG * (e[u] * e[w] - euw);
      }
    }
  }

  // ---- objective: 0.5 * sum of squared residuals (fitfunc.py:436-450, without the 1/M/norm) ----
  CTK_DEV_BIG double evaluate(const double* x) {
    ++evals;
    load_features(x);
    Real* pr = PR();
#pragma unroll 1
    for (int p = lane; p < M; p += CTK_WARP) pr[p] = 0;
    warp_sync();
    const Entry* flist = FLIST();
    Real* fe = FE();
    const int* fi = FI();
    for (int i = 0; i < n; ++i) {
      const Feat f = feat(i);
      const int cnt = fi[i * FI_STRIDE + FI_CNT];
      const Entry* fl = flist + i * a.lay.f_cap;
      Real* ge = fe + i * a.lay.f_cap;
#pragma unroll 1
      for (int t = lane; t < cnt; t += CTK_WARP) {
        Entry e = fl[t];
        Geo g = geometry(e, f);
        bool drop;
        Real gv = model_value(g, f, drop);
        ge[t] = gv;
        int p = entry_pixel(e);
        pr[p] += drop ? (Real) NAN : f.s * gv;
      }
      warp_sync();
    }
    const Real bg = (Real) value_of(x, 0, 0);
    double acc = 0., sr = 0., nv = 0.;
#pragma unroll 1
    for (int p = lane; p < M; p += CTK_WARP) {
      Real r = pixel_value(p) - bg - pr[p];
      pr[p] = r;
      if (r == r) { acc += (double) r * (double) r; sr += (double) r; nv += 1.; }
    }
    acc = warp_sum(acc);
    sum_r = warp_sum(sr);
    n_valid = warp_sum(nv);
    warp_sync();
    return 0.5 * acc;
  }

  // value of a compile-time sized register array at a run-time position (no local memory)
  template <int N> CTK_DEV static Real pick(const Real (&arr)[N], int k) {
    Real v = 0;
#pragma unroll
    for (int q = 0; q < N; ++q) v = (q == k) ? arr[q] : v;
    return v;
  }

  // Warp reduction of the register array v[OFF .. OFF + CNT): a halving butterfly (each step a lane
  // keeps one half of its entries and hands the other half to its partner) needs about CNT
  // shuffles where CNT plain warp sums need 5 CNT.  The total of entry k ends up on one owning
  // lane; f(k, total, owner) is called ONCE on every lane, in converged control flow.
  template <int OFF, int CNT, int N, class F>
  CTK_DEV void reduce_each(Real (&v)[N], F&& f) const {
#ifdef CTK_EMUL
    for (int k = 0; k < CNT; ++k) f(OFF + k, v[OFF + k], true);
#else
    static_assert(CNT >= 1 && CNT <= 32, "reduce_each: split longer arrays");
    constexpr int KP = CNT <= 1 ? 1 : CNT <= 2 ? 2 : CNT <= 4 ? 4 : CNT <= 8 ? 8 : CNT <= 16 ? 16 : 32;
    constexpr int STEPS = KP == 1 ? 0 : KP == 2 ? 1 : KP == 4 ? 2 : KP == 8 ? 3 : KP == 16 ? 4 : 5;
    int index = 0;
#pragma unroll
    for (int st = 0; st < STEPS; ++st) {
      const int h = KP >> (st + 1), m = 16 >> st;
      const bool up = (lane & m) != 0;
#pragma unroll
      for (int i = 0; i < h; ++i) {
        const Real lo = v[OFF + i];
        const Real hi = (i + h < CNT) ? v[OFF + i + h] : (Real) 0;
        v[OFF + i] = (up ? hi : lo) + shfl_xor(up ? lo : hi, m);
      }
      index += up ? h : 0;
    }
    Real r = v[OFF];
#pragma unroll
    for (int m = 16 / KP; m > 0; m >>= 1) r += shfl_xor(r, m);
    f(OFF + index, r, (lane & (32 / KP - 1)) == 0 && index < CNT);
#endif
  }
  template <int OFF, int CNT, int N, class F>
  CTK_DEV void reduce_all(Real (&v)[N], F&& f) const {
    if constexpr (CNT > 32) {
      reduce_each<OFF, 32>(v, f);
      reduce_all<OFF + 32, CNT - 32>(v, f);
    } else {
      reduce_each<OFF, CNT>(v, f);
    }
  }

  // ---- normal equations from the caches of the last evaluate() ---------------------------------
  // H = sum m m^T (packed lower, column-major), RHS = sum m r  (= -gradient of 0.5 sum r^2)
  // grad_only: refresh only RHS (the gradient); H and its factor are kept from the last full pass
  // (chord iteration, used close to the minimum where H has stopped changing).
  CTK_DEV_BIG void accumulate(bool grad_only) {
    if (grad_only) ++grad_accums; else ++accums;
    Real* H = Hm();
    double* rhs = RHS();
    const int nt = CS()[V];
    if (!grad_only) for (int t = lane; t < nt; t += CTK_WARP) H[t] = 0;
    CTK_FOR_V(v) rhs[v] = 0.;
    warp_sync();
    const int* cv = CV();
    const int vb = cv[0];
    if (vb >= 0 && lane == 0) { if (!grad_only) H[pk(vb, vb)] = (Real) n_valid; rhs[vb] = sum_r; }
    warp_sync();
    const Entry* flist = FLIST();
    const Real* fe = FE();
    const Real* pr = PR();
    const int* fi = FI();
    const int* sidx = SIDX();
    for (int i = 0; i < n; ++i) {
      const Feat f = feat(i);
      const int cnt = fi[i * FI_STRIDE + FI_CNT];
      const Entry* fl = flist + i * a.lay.f_cap;
      const Real* ge = fe + i * a.lay.f_cap;
      Real acc[LT + 2 * LD];       // [LT] m_u m_w, [LD] m_u, [LD] m_u r
#pragma unroll
      for (int k = 0; k < LT + 2 * LD; ++k) acc[k] = 0;
#pragma unroll 1
      for (int t = lane; t < cnt; t += CTK_WARP) {
        Entry e = fl[t];
        Real r = pr[entry_pixel(e)];
        if (!(r == r)) continue;
        Geo g = geometry(e, f);
        Real m[LD];
        model_derivs(g, f, ge[t], m);
        if (grad_only) {
#pragma unroll
          for (int u = 0; u < LD; ++u) acc[LT + LD + u] += m[u] * r;
        } else {
          int k = 0;
#pragma unroll
          for (int u = 0; u < LD; ++u) {
            acc[LT + u] += m[u];
            acc[LT + LD + u] += m[u] * r;
#pragma unroll
            for (int w = 0; w <= u; ++w) acc[k++] += m[u] * m[w];
          }
          if (C::FAM == CTK_FAMILY_GAUSS && newton) second_order(g, f, ge[t], r, acc);
        }
      }
      // the owning lane of every sum adds it to its target
      const int* sx = sidx + i * a.lay.sidx_stride;
      auto add = [&](int k, Real val, bool owner) {
        if (!owner) return;
        const int target = sx[k];
        if (target >= 0) {
          if (k < LT + LD) H[target] += val; else rhs[target] += (double) val;
        }
      };
      if (grad_only) reduce_all<LT + LD, LD>(acc, add);
      else reduce_all<0, LT + 2 * LD>(acc, add);
      warp_sync();
    }
    if (grad_only) return;
    // cross blocks over the pixels two features share
    const uint32_t* pairs = PAIRS();
    const int* phdr = PHDR();
    for (int q = 0; q < npairs; ++q) {
      const int i = phdr[q * 4], j = phdr[q * 4 + 1], start = phdr[q * 4 + 2], cnt = phdr[q * 4 + 3];
      const Feat f_i = feat(i), f_j = feat(j);
      const Entry *fl_i = flist + i * a.lay.f_cap, *fl_j = flist + j * a.lay.f_cap;
      const Real *ge_i = fe + i * a.lay.f_cap, *ge_j = fe + j * a.lay.f_cap;
      Real B[LD * LD];
#pragma unroll
      for (int k = 0; k < LD * LD; ++k) B[k] = 0;
#pragma unroll 1
      for (int t = lane; t < cnt; t += CTK_WARP) {
        uint32_t pe = pairs[start + t];
        int ti = (int) (pe & 0xffffu), tj = (int) (pe >> 16);
        Entry ei = fl_i[ti], ej = fl_j[tj];
        Real r = pr[entry_pixel(ei)];
        if (!(r == r)) continue;
        Real mi[LD], mj[LD];
        model_derivs(geometry(ei, f_i), f_i, ge_i[ti], mi);
        model_derivs(geometry(ej, f_j), f_j, ge_j[tj], mj);
#pragma unroll
        for (int u = 0; u < LD; ++u)
#pragma unroll
          for (int w = 0; w < LD; ++w) B[u * LD + w] += mi[u] * mj[w];
      }
      // entry (u, w) goes to (var(i,u), var(j,w)); two entries can share a target only when both
      // columns are shared within the cluster, so the adds go one lane at a time in that case
      const bool serial = shared_columns > 1;
      auto add = [&](int k, Real val, bool owner) {
        int target = -1;
        if (owner) {
          const int u = k / LD, w = k - u * LD;
          const int vu = cv[i * P + slot_col(u)], vw = cv[j * P + slot_col(w)];
          if (vu >= 0 && vw >= 0) {
            target = pk_sym(vu, vw);
            if (vu == vw) val *= (Real) 2;
          }
        }
        if (serial) {
#pragma unroll 1
          for (int turn = 0; turn < CTK_WARP; ++turn) {
            if (turn == lane && target >= 0) H[target] += val;
            warp_sync();
          }
        } else if (target >= 0) {
          H[target] += val;
        }
      };
      reduce_all<0, LD * LD>(B, add);
      warp_sync();
    }
  }

  // ---- distance constraints as augmented-Lagrangian rows: thin views on the cold helpers -------
  CTK_DEV ConView con_view(const double* x) const {
    ConView v;
    v.x = x; v.cv = CV(); v.con = CON(); v.n = n; v.P = P; v.nd = ND; v.n_con = n_con; v.w = pen_w;
    return v;
  }
  CTK_DEV double penalty(const double* x) const { return CTK_NCON ? con_penalty(con_view(x)) : 0.; }
  CTK_DEV double con_violation(const double* x) const {
    return CTK_NCON ? ctk::con_violation(con_view(x)) : 0.;
  }

  // ---- damped, bound-aware step --------------------------------------------------------------------
  // Builds K = H (+ constraint rows) in Lm and the full right-hand side in rhs_full, freezes the
  // active set, adds lambda*diag, factorises (Cholesky, column-major packed, every lane works on the
  // trailing block) with the forward substitution folded in, then back-substitutes.  On return D()
  // holds the step.  Returns false on breakdown.
  // reuse: keep the factor of the previous call (same active set required, else refactorise).
  // n_elim >= 0 (global-level fits): eliminate only the first n_elim unknowns -- the cluster's own
  // -- and return: the trailing block of the factor then holds the (scaled) Schur complement of
  // the shared unknowns and the tail of D() their reduced right-hand side.  Those unknowns get no
  // damping here (the host damps the summed block).
  CTK_DEV_BIG bool solve(double lambda, double* rhs_full, bool reuse, int n_elim = -1) {
    const Real* H = Hm();
    Real* Kf = Lm();
    double* d = D();
    double* sc = DG();                 // Jacobi scaling 1/sqrt(K_vv)
    Real* idg = IDG();
    int* act = ACT();
    const int* cs = CS();
    const RcT* rc = RC();
    const double *x = X(), *lo = LO(), *hi = HI();
    const int nt = cs[V];
#pragma unroll 1
    for (int v = lane; v < V; v += CTK_WARP) rhs_full[v] = RHS()[v];
    warp_sync();
    if (reuse) {
      // chord step: same matrix, same scaling, new right-hand side; the frozen set must not move
      bool moved = false;
#pragma unroll 1
      for (int v = lane; v < V; v += CTK_WARP) {
        const double g = rhs_full[v];
        const bool frozen = (x[v] <= lo[v] && g < 0.) || (x[v] >= hi[v] && g > 0.) || !(lo[v] < hi[v]);
        moved |= frozen != (act[v] != 0);
        d[v] = frozen ? 0. : g * sc[v];
      }
      if (warp_any(moved)) reuse = false;
      warp_sync();
    }
    if (reuse) {
      for (int j = 0; j < V; ++j) {                        // forward substitution with L
        const double yj = d[j] * (double) idg[j];
        warp_sync();
        if (lane == 0) d[j] = yj;
#pragma unroll 1
        for (int r = j + 1 + lane; r < V; r += CTK_WARP) d[r] -= (double) Kf[cs[j] + r - j] * yj;
        warp_sync();
      }
    } else {
#pragma unroll 1
      for (int t = lane; t < nt; t += CTK_WARP) Kf[t] = H[t];
      warp_sync();
      if (CTK_NCON > 0) {
        if (lane == 0) con_add_rows(con_view(x), Kf, cs, rhs_full);
        warp_sync();
      }
      double dmax = 0.;
#pragma unroll 1
      for (int v = lane; v < V; v += CTK_WARP) dmax = fmax(dmax, (double) Kf[cs[v]]);
      dmax = warp_max_d(dmax);
      const double floor_ = fmax(dmax * 1e-14, 1e-30);
#pragma unroll 1
      for (int v = lane; v < V; v += CTK_WARP) {
        const double g = rhs_full[v];
        const bool frozen = (x[v] <= lo[v] && g < 0.) || (x[v] >= hi[v] && g > 0.) || !(lo[v] < hi[v]);
        act[v] = frozen ? 1 : 0;
        const double s = (double) fast_rsqrt((Real) fmax((double) Kf[cs[v]], floor_));
        sc[v] = s;
        d[v] = frozen ? 0. : g * s;
      }
      warp_sync();
      // scaled, damped system: (S K S + lambda I)(S^-1 step) = S rhs; frozen rows become identity
      const Real lam1 = (Real) (1. + lambda);
#pragma unroll 1
      for (int t = lane; t < nt; t += CTK_WARP) {
        const int r = rc_row(rc[t]), c = rc_col(rc[t]);
        Real v;
        if (act[r] || act[c]) v = (r == c) ? (Real) 1 : (Real) 0;
        else if (r == c) v = (n_elim >= 0 && r >= n_elim) ? (Real) 1 : lam1;
        else v = Kf[t] * (Real) (sc[r] * sc[c]);
        Kf[t] = v;
      }
      warp_sync();
      const int n_cols = n_elim >= 0 ? n_elim : V;
      for (int j = 0; j < n_cols; ++j) {
        const int cj = cs[j];
        const Real piv = Kf[cj];
        if (!(piv > (Real) 1e-7)) return false;              // also catches NaN
        const Real inv = fast_rsqrt(piv);
        const double yj = d[j] * (double) inv;               // forward substitution, row j
        warp_sync();
#pragma unroll 1
        for (int r = j + lane; r < V; r += CTK_WARP) Kf[cj + r - j] *= inv;
        if (lane == 0) { idg[j] = inv; d[j] = yj; }
        warp_sync();
#pragma unroll 1
        for (int t = cs[j + 1] + lane; t < nt; t += CTK_WARP) {
          const int r = rc_row(rc[t]), c = rc_col(rc[t]);
          Kf[t] -= Kf[cj + r - j] * Kf[cj + c - j];
        }
#pragma unroll 1
        for (int r = j + 1 + lane; r < V; r += CTK_WARP) d[r] -= (double) Kf[cj + r - j] * yj;
        warp_sync();
      }
    }
    if (n_elim >= 0) return true;
    for (int j = V - 1; j >= 0; --j) {                     // back substitution with L^T
      const double zj = d[j] * (double) idg[j];
      warp_sync();
      if (lane == 0) d[j] = zj;
#pragma unroll 1
      for (int r = lane; r < j; r += CTK_WARP) d[r] -= (double) Kf[cs[r] + j - r] * zj;
      warp_sync();
    }
#pragma unroll 1
    for (int v = lane; v < V; v += CTK_WARP) d[v] *= sc[v];
    warp_sync();
    return true;
  }

  // predicted decrease of the (augmented) objective for step s: rhs.s - 0.5 s^T K s
  CTK_DEV_BIG double predicted(const double* s, const double* rhs_full) const {
    const Real* H = Hm();
    const RcT* rc = RC();
    const int nt = CS()[V];
    double acc = 0.;
#pragma unroll 1
    for (int t = lane; t < nt; t += CTK_WARP) {
      const int r = rc_row(rc[t]), c = rc_col(rc[t]);
      acc -= (r == c ? 0.5 : 1.) * (double) H[t] * s[r] * s[c];
    }
    CTK_FOR_V(u) acc += s[u] * rhs_full[u];
    acc = warp_sum(acc);
    // constraint rows: the gradient part is already in rhs_full; add -0.5 w (A s)^2
    if (CTK_NCON > 0) acc += con_quadratic(con_view(X()), s);
    return acc;
  }

  CTK_DEV bool is_pos_var(int v) const {
    bool hit = false;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      const int m = mode(2 + k);
      const int b = CV()[2 + k];                     // feature 0 holds the first index of the column
      hit = hit || (m != CTK_MODE_CONST && v >= b && v < b + (m == CTK_MODE_VAR ? n : 1));
    }
    return hit;
  }

  // ---- projected Levenberg-Marquardt with augmented-Lagrangian constraints ---------------------
  // Minimises from X() (already holding the start vector).  Returns status; *f_data = 0.5 sum r^2.
  // Written as one loop in which evaluate(), accumulate(), solve() and predicted() each appear
  // once, so that the (fully inlined) kernel holds a single copy of every phase.
  CTK_DEV_BIG int minimise(double* f_data) {
    const bool f32 = sizeof(Real) == 4;
    const double xtol = a.prob.xtol > 0. ? a.prob.xtol : (f32 ? 2e-6 : 1e-9);
    const double eps_f = f32 ? 4e-6 : 1e-13;      // resolution of the objective
    const double ctol = f32 ? 1e-8 : 1e-10;
    double *x = X(), *xt = XT(), *d = D();
    double* rhs_full = dvec(a.lay.o_rhsf);            // rhs incl. constraint terms, before freezing
    double lambda = 1e-3, nu = 2.;
    if (lane == 0) for (int j = 0; j < 6; ++j) CON()[j] = 0.;
    CTK_FOR_V(v) xt[v] = x[v];
    warp_sync();
    pen_w = 0.;
    double pen_w0 = 0.;
    double fd = 0., fa = 0., pred = 0., worst = 0.;
    double c_prev = 0.;
    int al_rounds = 0, rejects = 0;
    double prev_small_step = INFINITY;
    double last_step = INFINITY;           // scaled size of the last accepted step
    // chord iterations: once the steps are small the normal matrix has stopped changing, so only the
    // gradient is refreshed and the previous factor is reused (unconstrained clusters only)
    const double chord_tol = a.prob.chord_tol;
    bool first = true, force = true, need_eval = true, chord_next = false;
    // Exact Hessian (gauss family): on from the first accepted step (the start point itself, where
    // the residual is largest and the corrected matrix most likely indefinite, gets plain
    // Gauss-Newton); off for good if the corrected matrix ever fails to factorise.  Measured on
    // config 2: 15 % fewer objective evaluations per cluster (float32), 27 % (float64).
    newton = false;
    bool newton_allowed = C::FAM == CTK_FAMILY_GAUSS;
    for (int it = 0; it <= a.prob.lm_max_iter; ++it) {
#ifndef CTK_EMUL
      // keep the flag opaque: knowing its value on the back edges, the compiler threads those
      // jumps and duplicates the whole tail of the loop (solve, predicted) -- 30 KB of code in a
      // kernel that is bound by instruction fetch
      { int flag = need_eval ? 1 : 0; asm volatile("" : "+r"(flag)); need_eval = flag != 0; }
#endif
      if (need_eval) {
        const double fdt = evaluate(xt);
        const double fat = fdt + (first ? 0. : penalty(xt));
        bool accept;
        if (force) {
          if (!finite_d(fdt)) return CTK_FAIL_NUMERIC;
          accept = true;
        } else {
          // below the resolution of the objective the comparison fat < fa is rounding noise: trust
          // the quadratic model there (the gradient stays accurate long after the objective is flat)
          const bool noise = pred > 0. && pred <= eps_f * fabs(fa) &&
                             fabs(fat - fa) <= 8. * eps_f * fabs(fa);
          CTK_TRACEF("   pred %.3g fat-fa %.3g noise %d\n", pred, fat - fa, (int) noise);
          accept = finite_d(fat) && pred > 0. && (fat < fa || noise);
          if (accept) {
            if (!noise) {
              const double rho = (fa - fat) / pred;
              const double t = 2. * rho - 1.;
              lambda = fmax(lambda * fmax(1. / 3., 1. - t * t * t), 1e-12);
            } else {
              if (worst > 0.9 * prev_small_step) lambda *= 4.;    // not contracting: damp harder
              prev_small_step = worst;
            }
            nu = 2.;
            rejects = 0;
            last_step = worst;
            newton = newton_allowed;
          } else {
            lambda *= nu;
            nu *= 2.;
            if (++rejects > 40 || lambda > 1e18) {
              // no representable descent step is left: x is a numerical minimiser
              *f_data = fd;
              return (CTK_NCON == 0 || con_violation(x) <= 1e-6) ? CTK_OK : CTK_FAIL_NO_CONVERGENCE;
            }
            if (chord_next) {
              // the step from the stale matrix failed: rebuild everything at x (the caches hold the
              // rejected point, so x is evaluated again)
              CTK_FOR_V(v) xt[v] = x[v];
              warp_sync();
              force = true;
              chord_next = false;
              continue;
            }
          }
        }
        if (accept) {
          CTK_FOR_V(v) x[v] = xt[v];
          warp_sync();
          fd = fdt;
          fa = fat;
          const bool cheap = !force && CTK_NCON == 0 && worst < chord_tol;
          accumulate(cheap);
          chord_next = cheap;
          force = false;
          if (first && CTK_NCON > 0) {
            // penalty weight relative to the curvature of the data term in the position variables
            double hmax = 0.;
            CTK_FOR_V(v)
              if (is_pos_var(v)) hmax = fmax(hmax, (double) Hm()[CS()[v]]);
            hmax = warp_max_d(hmax);
            double a2 = 0.;
#pragma unroll
            for (int k = 0; k < ND; ++k) a2 = fmax(a2, 8. / (CON()[6 + k] * CON()[6 + k]));
            pen_w = pen_w0 = 100. * fmax(hmax, 1e-30) / a2;
            fa = fd + penalty(x);
            c_prev = con_violation(x);
          }
          first = false;
        }
      }
      need_eval = true;
      if (!solve(lambda, rhs_full, chord_next)) {
        if (newton) {
          // the corrected matrix is not positive definite here: back to Gauss-Newton at x
          newton = false;
          newton_allowed = false;
          CTK_FOR_V(v) xt[v] = x[v];
          warp_sync();
          force = true;
          chord_next = false;
          continue;
        }
        lambda = fmax(lambda * 10., 1e-8);
        if (++rejects > 60) { *f_data = fd; return CTK_FAIL_NUMERIC; }
        need_eval = false;
        continue;
      }
      // trial point, projected on the box
      worst = 0.;
      CTK_FOR_V(v) {
        const double t = fmin(fmax(x[v] + d[v], LO()[v]), HI()[v]);
        xt[v] = t;
        const double s = t - x[v];
        d[v] = s;
        const float scale = is_pos_var(v) ? 1.f : fmaxf(1.f, fabsf((float) x[v]));
        worst = fmax(worst, (double) (fabsf((float) s) / scale));
      }
      worst = warp_max_d(worst);
      warp_sync();
      if (!finite_d(worst)) { *f_data = fd; return CTK_FAIL_NUMERIC; }
      CTK_TRACEF("it %d lambda %.3g worst %.3g fa %.10g cv %.3g w %.3g al %d newton %d\n", it, lambda, worst, fa,
                 CTK_NCON ? con_violation(x) : 0., pen_w, al_rounds, (int) newton);
#if defined(CTK_EMUL) && defined(CTK_TRACE)
      for (int v = 0; v < V && v < 8; ++v)
        printf("      x[%d] %.6f step %.3g act %d lo %.4g hi %.4g rhs %.3g\n", v, x[v], d[v], ACT()[v], LO()[v], HI()[v], rhs_full[v]);
#endif
      if (worst <= xtol) {
        // stationary for the current multipliers
        if (CTK_NCON == 0) { *f_data = fd; return CTK_OK; }
        // take the (sub-tolerance) step: it carries the Newton correction towards c(x) = 0; the
        // data term is flat at this scale, so its caches stay valid
        CTK_FOR_V(v) x[v] = xt[v];
        warp_sync();
        const double cv = con_violation(x);
        if (cv <= ctol || al_rounds >= 40 || (al_rounds > 2 && cv >= 0.5 * c_prev && cv <= 1e-6)) {
          *f_data = fd;
          return CTK_OK;
        }
        if (lane == 0) {
          const ConView cvw = con_view(x);
          for (int j = 0; j < CTK_NCON; ++j) CON()[j] += pen_w * con_value(cvw, j);
        }
        warp_sync();
        if (al_rounds > 0 && cv > 0.25 * c_prev && pen_w < 1e6 * pen_w0) pen_w *= 10.;
        c_prev = cv;
        ++al_rounds;
        fa = fd + penalty(x);
        lambda = fmin(lambda, 1e-3);
        need_eval = false;
        continue;
      }
      pred = predicted(d, rhs_full);
    }
    // Iteration limit.  Objectives with jumps (ring/disc "safe" pixels entering or leaving the sums,
    // fitfunc.py:20-26) can pin the minimiser against a discontinuity, where the damped steps shrink
    // geometrically without ever meeting xtol; a point that only moves by < 1e-3 is accepted.
    *f_data = fd;
    return last_step <= 1e-3 ? CTK_OK : CTK_FAIL_NO_CONVERGENCE;
  }

  // ---- basin search (ring / disc) ---------------------------------------------------------------
  // The ring and disc objectives jump whenever a "safe" pixel (closer than one pixel to a centre,
  // fitfunc.py:20-26) enters or leaves the sums, and the disc objective jumps at disc_size = 0
  // (fitfunc.py:121-131: gauss branch, safe pixels dropped instead of forced to 1).  They are
  // piecewise smooth with many shallow basins; SLSQP's line search hops between them, a damped
  // Gauss-Newton iteration stays in the basin it starts in.  So that the fit never ends ABOVE the
  // reference's minimum, the neighbouring basins are probed: from the minimum found, every centre is
  // displaced along every axis (and the disc is switched to its gauss branch), the minimiser runs
  // again, and a lower end point replaces the current one; repeated while a sweep improves.
  // The smooth gauss objective has no such structure: one run.
  // Probe list: [axis shifts +-step | axis shifts +-step/5 | safe-pixel crossings | shape parameter]
  enum { N_NEIGH = ND == 2 ? 9 : 27 };
  CTK_DEV int extra_count() const {
    if (!(C::EX && C::NE)) return 0;
    const int m = mode(2 + ND + NS);
    return m == CTK_MODE_VAR ? n : (m == CTK_MODE_CLUSTER ? 1 : 0);
  }
  CTK_DEV int probe_count() const {
    if (C::FAM == CTK_FAMILY_GAUSS) return 0;
    return (4 * ND + N_NEIGH) * n + (C::FAM == CTK_FAMILY_DISC ? 3 : 2) * extra_count();
  }
  // X() = XB() + probe q; false when the probe does not apply (constant column, bound in the way,
  // no pixel near the unit sphere in that direction)
  CTK_DEV bool apply_probe(int q) {
    double *x = X();
    const double *xb = XB(), *lo = LO(), *hi = HI();
    CTK_FOR_V(v) x[v] = xb[v];
    warp_sync();
    bool ok = false;
    if (q < 4 * ND * n) {
      // centres: +-probe_step along every axis, then +-probe_step / 5
      const double step = q < 2 * ND * n ? a.prob.probe_step : 0.2 * a.prob.probe_step;
      if (q >= 2 * ND * n) q -= 2 * ND * n;
      const int i = q / (2 * ND), k = (q >> 1) % ND;
      const int v = var_of(2 + k, i);
      if (v >= 0 && mode(2 + k) == CTK_MODE_VAR) {
        const double t = fmin(fmax(xb[v] + ((q & 1) ? step : -step), lo[v]), hi[v]);
        ok = t != xb[v];
        warp_sync();
        if (lane == 0) x[v] = t;
      }
      warp_sync();
      return ok;
    }
    q -= 4 * ND * n;
    if (q < N_NEIGH * n) {
      // safe-pixel crossings: a pixel whose distance to the centre is close to 1 enters or leaves
      // the sums when the centre moves a little (fitfunc.py:20-26); put the centre just on the
      // other side of that unit sphere, moving along the line through the pixel
      const int i = q / N_NEIGH;
      int o = q - i * N_NEIGH;
      double c[3] = {0., 0., 0.}, dlt[3] = {0., 0., 0.};
      double d2 = 0.;
      bool free_pos = true;
#pragma unroll
      for (int k = ND - 1; k >= 0; --k) {
        c[k] = value_of(xb, 2 + k, i);
        const double pix = rint(c[k]) + (double) (o % 3 - 1);
        o /= 3;
        dlt[k] = c[k] - pix;
        d2 += dlt[k] * dlt[k];
        free_pos = free_pos && mode(2 + k) == CTK_MODE_VAR;
      }
      const double dist = sqrt(d2);
      if (free_pos && dist > 0.75 && dist < 1.25) {
        const double target = dist >= 1. ? 1. - 1e-3 : 1. + 1e-3;
        ok = true;
#pragma unroll
        for (int k = 0; k < ND; ++k) {
          const int v = var_of(2 + k, i);
          const double t = c[k] + dlt[k] * (target / dist - 1.);
          ok = ok && t >= lo[v] && t <= hi[v];
        }
        warp_sync();
        if (ok && lane == 0) {
#pragma unroll
          for (int k = 0; k < ND; ++k) x[var_of(2 + k, i)] = c[k] + dlt[k] * (target / dist - 1.);
        }
      }
      warp_sync();
      return ok;
    }
    q -= N_NEIGH * n;
    {
      // the shape parameter (thickness / disc_size) of one feature: x 0.75, x 1.33, and for the disc
      // its gauss branch disc_size <= 0 (flat in disc_size, safe pixels dropped, fitfunc.py:121-131)
      const int per = C::FAM == CTK_FAMILY_DISC ? 3 : 2;
      const int j = q / per, kind = q - j * per;
      const int v = var_of(2 + ND + NS, 0) + j;
      double t;
      if (kind == 2) t = fmin(xb[v], -0.25);
      else t = xb[v] * (kind ? 4. / 3. : 0.75);
      t = fmin(fmax(t, lo[v]), hi[v]);
      ok = kind == 2 ? (t <= 0. && xb[v] > 0.) : (t != xb[v]);
      warp_sync();
      if (ok && lane == 0) x[v] = t;
    }
    warp_sync();
    return ok;
  }

  CTK_DEV_BIG int search(double* f_data) {
    const int n_probe = (a.prob.probe_step > 0. && a.prob.probe_sweeps > 0 &&
                         n <= CTK_PROBE_MAX_FEATURES) ? probe_count() : 0;
    const int total = 1 + a.prob.probe_sweeps * n_probe;
    double best = INFINITY;
    bool improved = false;
    int status = CTK_OK;
    for (int k = 0; k < total; ++k) {
      bool go = true;
      if (k == 0) {
        CTK_FOR_V(v) X()[v] = X0()[v];
        warp_sync();
      } else {
        go = apply_probe((k - 1) % n_probe);
      }
      if (go) {
        double fd = 0.;
        const int st = minimise(&fd);
        if (k == 0) {
          status = st;
          if (st != CTK_OK || n_probe == 0) { *f_data = fd; return st; }
        }
        if (st == CTK_OK && (k == 0 || fd < best * (1. - 1e-9))) {
          best = fd;
          improved = k > 0;
          CTK_FOR_V(v) XB()[v] = X()[v];
          warp_sync();
        }
      }
      if (k > 0 && (k - 1) % n_probe == n_probe - 1) {     // end of a sweep
        if (!improved) break;
        improved = false;
      }
    }
    CTK_FOR_V(v) X()[v] = XB()[v];
    warp_sync();
    *f_data = best;
    return status;
  }

  // ---- global-level fits (refine.py:319-332): one pass over one cluster ---------------------------
  // Some columns are shared by ALL features of the table, so the whole table is ONE problem:
  //   F = sum_c w_c f_c(x_c, g),  f_c = 0.5 sum r^2 of cluster c, w_c = 2 / (M_c norm)
  // (fitfunc.py:436-450 divides every cluster's sum by its own pixel count).  Its normal matrix is
  // a block arrow: per-cluster blocks A_c, couplings B_c to the shared unknowns g, and sum_c C_c.
  // The host iterates; every iteration launches two passes over all clusters:
  //   phase 1  at the current point: pixel set, residuals, normal equations (as in the per-cluster
  //            fit), elimination of the cluster's own unknowns, and w_c (C_c - B_c^T A_c^-1 B_c),
  //            w_c (g_g - B_c^T A_c^-1 g_c), w_c f_c added to the global accumulators -- the only
  //            reduction across clusters (an all-reduce when frames are sharded over ranks);
  //   phase 2  with the step of the shared unknowns (the host solved the small summed system): the
  //            cluster's own step by back substitution, the trial point projected on the box, its
  //            objective, the predicted decrease and the step size, and the trial parameters.
  enum { GLOBAL_F = 0, GLOBAL_FT = 1, GLOBAL_PRED = 2, GLOBAL_STEP = 3, GLOBAL_FAILED = 4,
         GLOBAL_SINGULAR = 5, GLOBAL_HEADER = CTK_GLOBAL_HEADER };
  CTK_DEV void run_global(int cluster) {
    feat0 = a.cluster_offset[cluster];
    n = a.cluster_offset[cluster + 1] - feat0;
    const int fidx = a.cluster_frame[cluster];
    frame = a.frames[fidx];
    fmax_ = 0.;
    evals = accums = grad_accums = outers = n_entries = n_pair_entries = 0;
    M = 0; V = 0; n_con = 0; pen_w = 0.;
    double* acc = a.global_accum;
    int status = (n <= 0 || n > a.lay.n_max) ? CTK_FAIL_TOO_LARGE : CTK_OK;
    if (status == CTK_OK) status = setup_variables();
    if (status == CTK_OK) {
      double* mc = MC();
#pragma unroll 1
      for (int i = lane; i < n; i += CTK_WARP) {
#pragma unroll
        for (int k = 0; k < ND; ++k)
          mc[i * 3 + k] = a.mask_centres ? a.mask_centres[(int64_t) (feat0 + i) * ND + k]
                                         : a.params_in[(int64_t) (feat0 + i) * P + 2 + k];
      }
      warp_sync();
      status = build_pixels();
    }
    double fd = 0.;
    if (status == CTK_OK) {
      CTK_FOR_V(v) { X()[v] = X0()[v]; XT()[v] = X0()[v]; }
      warp_sync();
      fd = evaluate(X());
      if (!finite_d(fd)) status = CTK_FAIL_NUMERIC;
    }
    if (status != CTK_OK) {
      if (lane == 0) { atomic_add_d(acc + GLOBAL_FAILED, 1.); a.status_out[cluster] = status; }
      warp_sync();
      return;
    }
    const double w = 2. / ((double) M * a.global_norm);
    const int G = V - V_loc;
    newton = C::FAM == CTK_FAMILY_GAUSS && a.use_newton != 0;
    accumulate(false);
    double* rhs_full = dvec(a.lay.o_rhsf);
    if (!solve(a.global_lambda, rhs_full, false, V_loc)) {
      if (lane == 0) { atomic_add_d(acc + GLOBAL_SINGULAR, 1.); a.status_out[cluster] = CTK_OK; }
      warp_sync();
      return;
    }
    double* d = D();
    const double* sc = DG();
    const Real* Kf = Lm();
    const int* cs = CS();
    if (a.global_phase == 1) {
      if (lane == 0) atomic_add_d(acc + GLOBAL_F, w * fd);
#pragma unroll 1
      for (int t = lane; t < G * (G + 3) / 2; t += CTK_WARP) {
        if (t < G) {                                   // reduced right-hand side, unscaled
          atomic_add_d(acc + GLOBAL_HEADER + t, w * d[V_loc + t] / sc[V_loc + t]);
        } else {                                       // Schur block entry (u, v), u >= v
          int u = 0, rem = t - G;
          while (rem > u) { rem -= u + 1; ++u; }
          const int ru = V_loc + u, rv = V_loc + rem;
          atomic_add_d(acc + GLOBAL_HEADER + t, w * (double) Kf[cs[rv] + ru - rv] / (sc[ru] * sc[rv]));
        }
      }
    } else {
      // step of the cluster's own unknowns with the shared part fixed: L_ll^T y_l = d_l - L_gl^T y_g
      warp_sync();
      for (int j = V - 1; j >= 0; --j) {
        const double zj = j >= V_loc ? a.global_step[j - V_loc] / sc[j] : d[j] * (double) IDG()[j];
        warp_sync();
        if (lane == 0) d[j] = zj;
        const int top = j < V_loc ? j : V_loc;
#pragma unroll 1
        for (int r = lane; r < top; r += CTK_WARP) d[r] -= (double) Kf[cs[r] + j - r] * zj;
        warp_sync();
      }
      CTK_FOR_V(v) d[v] *= sc[v];
      warp_sync();
      double worst = 0.;
      double *x = X(), *xt = XT();
      CTK_FOR_V(v) {
        // shared unknowns take the host's step exactly (every cluster must end with the same value)
        const double t = v >= V_loc ? x[v] + a.global_step[v - V_loc]
                                    : fmin(fmax(x[v] + d[v], LO()[v]), HI()[v]);
        xt[v] = t;
        const double st = t - x[v];
        d[v] = st;
        if (v < V_loc) {
          const double scale = is_pos_var(v) ? 1. : fmax(1., fabs(x[v]));
          worst = fmax(worst, fabs(st) / scale);
        }
      }
      worst = warp_max_d(worst);
      warp_sync();
      const double pred = predicted(d, rhs_full);
      const double ft = evaluate(xt);
      if (lane == 0) {
        atomic_add_d(acc + GLOBAL_FT, finite_d(ft) ? w * ft : INFINITY);
        atomic_add_d(acc + GLOBAL_PRED, w * pred);
        if (finite_d(worst)) atomic_max_nonneg_d(acc + GLOBAL_STEP, worst);
        else atomic_add_d(acc + GLOBAL_FT, INFINITY);
      }
      double* pout = a.params_out + (int64_t) feat0 * P;
#pragma unroll 1
      for (int t = lane; t < n * P; t += CTK_WARP) {
        const int i = t / P, c = t - i * P;
        pout[t] = value_of(xt, c, i);
      }
    }
    if (lane == 0) {
      a.status_out[cluster] = CTK_OK;
      a.cost_out[cluster] = w * fd;
    }
    warp_sync();
  }

  // ---- whole cluster (refine.py:343-430) -------------------------------------------------------
  CTK_DEV void run(int cluster) {
    feat0 = a.cluster_offset[cluster];
    n = a.cluster_offset[cluster + 1] - feat0;
    const int fidx = a.cluster_frame[cluster];
    frame = a.frames[fidx];
    fmax_ = a.frame_max[fidx];
    evals = accums = grad_accums = outers = n_entries = n_pair_entries = 0;
    M = 0;
    V = 0;
    int status = CTK_OK;
    double cost = NAN;
    if (n <= 0 || n > a.lay.n_max || (!C::BIG && n > CTK_MAX_CLUSTER_FEATURES))
      status = CTK_FAIL_TOO_LARGE;
    if (status == CTK_OK) status = setup_variables();
    // constraints apply to clusters of exactly their size, with free per-feature positions
    n_con = 0;
    if (status == CTK_OK) {
      bool pos_var = true;
#pragma unroll
      for (int k = 0; k < ND; ++k) pos_var = pos_var && mode(2 + k) == CTK_MODE_VAR;
      if (pos_var && n == 2 && (a.prob.constraint_mask & CTK_CONSTRAINT_DIMER)) {
        n_con = 1;
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < ND; ++k) CON()[6 + k] = a.prob.dimer_dist[k];
        }
      } else if (pos_var && n == 4 && (a.prob.constraint_mask & CTK_CONSTRAINT_TETRAMER)) {
        n_con = ND == 2 ? 4 : 6;
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < ND; ++k) CON()[6 + k] = a.prob.tetramer_dist[k];
        }
      } else if (pos_var && n == 3 && (a.prob.constraint_mask & CTK_CONSTRAINT_TRIMER)) {
        n_con = 3;
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < ND; ++k) CON()[6 + k] = a.prob.trimer_dist[k];
        }
      }
    }
    double fd = 0.;
    if (status == CTK_OK) {
      double* mc = MC();
#pragma unroll 1
      for (int i = lane; i < n; i += CTK_WARP) {
#pragma unroll
        for (int k = 0; k < ND; ++k) mc[i * 3 + k] = a.params_in[(int64_t) (feat0 + i) * P + 2 + k];
      }
      warp_sync();
      for (int outer = 0; outer < a.prob.max_iter; ++outer) {
        ++outers;
        status = build_pixels();
        if (status != CTK_OK) break;
        if (C::FAM == CTK_FAMILY_GAUSS) {      // smooth objective: one minimisation, no basin search
          CTK_FOR_V(v) X()[v] = X0()[v];       // restarts from X0, refine.py:361-365
          warp_sync();
          status = minimise(&fd);
        } else {
          status = search(&fd);
        }
        if (status != CTK_OK) break;
        // accept when every feature stayed within max_shift of its mask centre, refine.py:383-385
        bool moved = false;
#pragma unroll 1
        for (int i = lane; i < n; i += CTK_WARP) {
          double s = 0.;
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            double dlt = value_of(X(), 2 + k, i) - mc[i * 3 + k];
            s += dlt * dlt;
          }
          moved |= !(s < a.prob.max_shift * a.prob.max_shift);
        }
        moved = warp_any(moved);
        if (!moved) break;
        warp_sync();
#pragma unroll 1
        for (int i = lane; i < n; i += CTK_WARP) {
#pragma unroll
          for (int k = 0; k < ND; ++k) mc[i * 3 + k] = value_of(X(), 2 + k, i);
        }
        warp_sync();
      }
    }
    if (status == CTK_OK) {
      // rms_dev = sqrt(fun / residual_factor), fun = sum diff^2 / M / norm   (refine.py:354, 379)
      double norm = fmax_ * fmax_ / a.prob.residual_factor;
      double fun = 2. * fd / (double) M / norm;
      cost = sqrt(fun / a.prob.residual_factor);
      if (!finite_d(cost)) status = CTK_FAIL_NUMERIC;
      else if (cost > a.prob.max_rms_dev) status = CTK_FAIL_RMS_DEV;
    }
    // write back (refine.py:408-427): failure leaves the parameters untouched and cost = NaN
    if (n > 0) {
      const double* pin = a.params_in + (int64_t) feat0 * P;
      double* pout = a.params_out + (int64_t) feat0 * P;
#pragma unroll 1
      for (int t = lane; t < n * P; t += CTK_WARP) {
        int i = t / P, c = t - i * P;
        double v = pin[t];
        if (status == CTK_OK) v = value_of(X(), c, i);
        pout[t] = v;
      }
    }
    if (lane == 0) {
      a.cost_out[cluster] = status == CTK_OK ? cost : NAN;
      a.status_out[cluster] = status;
      int32_t* st = a.stats_out + (int64_t) cluster * CTK_STATS;
      st[CTK_STAT_EVALS] = evals; st[CTK_STAT_ACCUMS] = accums; st[CTK_STAT_OUTER] = outers;
      st[CTK_STAT_PIXELS] = M; st[CTK_STAT_ENTRIES] = n_entries;
      st[CTK_STAT_PAIR_ENTRIES] = n_pair_entries; st[CTK_STAT_VARS] = V;
      st[CTK_STAT_GRAD_ACCUMS] = grad_accums;
      if (status == CTK_FAIL_TOO_LARGE && a.overflow != nullptr) {
        const int k = atomic_next(a.overflow);
        if (k < a.overflow_cap) a.overflow[1 + k] = cluster;
      }
    }
    warp_sync();
  }
};

}  // namespace ctk
