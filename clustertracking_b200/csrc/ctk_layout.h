// ctk_layout.h -- host-side sizing of the per-cluster shared-memory slice.
#pragma once

#include <string.h>

#include "ctk_solver.cuh"

namespace ctk {

// Upper bound on the pixels one feature's mask can hold (refine.py:43-44).  The mask centre is
// within half a pixel of its rounded value per axis, so offset o can be inside only if
// sum_k (max(|o_k| - 0.5, 0) / r_k)^2 <= 1.
inline int mask_capacity(const int32_t* radius, int ndim) {
  int r[3] = {0, 0, 0};
  for (int k = 0; k < ndim; ++k) r[3 - ndim + k] = radius[k];
  int count = 0;
  for (int z = -r[0] - 1; z <= r[0] + 1; ++z)
    for (int y = -r[1] - 1; y <= r[1] + 1; ++y)
      for (int x = -r[2] - 1; x <= r[2] + 1; ++x) {
        const int o[3] = {z, y, x};
        double s = 0.;
        bool out = false;
        for (int k = 0; k < 3; ++k) {
          double a = fabs((double) o[k]) - 0.5;
          if (a < 0.) a = 0.;
          if (r[k] == 0) { out = out || o[k] != 0; continue; }
          s += (a / r[k]) * (a / r[k]);
        }
        if (!out && s <= 1.0 + 1e-9) ++count;
      }
  return count;
}

// Largest pixel count of one mask over a 16^ndim grid of sub-pixel centre offsets, plus a margin.
// Used for the optimistic capacities: a cluster that overflows them gets CTK_FAIL_TOO_LARGE and
// is relaunched by the caller with a larger capacity (see ctk_refine_batch).
inline int mask_typical(const int32_t* radius, int ndim) {
  int r[3] = {0, 0, 0};
  for (int k = 0; k < ndim; ++k) r[3 - ndim + k] = radius[k];
  const int steps = 16;
  int best = 0;
  for (int a = 0; a < (r[0] ? steps : 1); ++a)
    for (int b = 0; b < (r[1] ? steps : 1); ++b)
      for (int c = 0; c < (r[2] ? steps : 1); ++c) {
        const double off[3] = {a / (double) steps - 0.5, b / (double) steps - 0.5,
                               c / (double) steps - 0.5};
        int count = 0;
        for (int z = -r[0] - 1; z <= r[0] + 1; ++z)
          for (int y = -r[1] - 1; y <= r[1] + 1; ++y)
            for (int x = -r[2] - 1; x <= r[2] + 1; ++x) {
              const int o[3] = {z, y, x};
              double s = 0.;
              bool out = false;
              for (int k = 0; k < 3; ++k) {
                if (r[k] == 0) { out = out || o[k] != 0; continue; }
                const double q = (o[k] - off[k]) / r[k];
                s += q * q;
              }
              if (!out && s <= 1.0) ++count;
            }
        if (count > best) best = count;
      }
  return best;
}

inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// Fills `lay` for clusters of up to n_max features.  Returns false on an invalid request.
// n_max > CTK_MAX_CLUSTER_FEATURES selects the large-cluster layout (global-memory workspace).
inline bool compute_layout(const ctk_problem_t& p, int n_max, Layout* lay) {
  memset(lay, 0, sizeof(*lay));
  if (n_max < 1 || n_max > CTK_MAX_BIG_FEATURES) return false;
  const bool big = n_max > CTK_MAX_CLUSTER_FEATURES;
  const int nd = p.ndim;
  int v_max = 0;
  for (int c = 0; c < p.n_params; ++c) {
    if (p.modes[c] == CTK_MODE_VAR) v_max += n_max;
    else if (p.modes[c] == CTK_MODE_CLUSTER || p.modes[c] == CTK_MODE_GLOBAL) v_max += 1;
  }
  if (v_max < 1) v_max = 1;
  if (v_max > (big ? 65535 : 255)) return false;        // packed (row, column) fields
  const int rb = p.compute_dtype == CTK_COMPUTE_F64 ? 8 : 4;
  lay->n_max = n_max;
  lay->v_max = v_max;
  // Capacities.  The largest launch class provisions for the rigorous worst case; smaller classes
  // provision for what clusters of that size typically need and rely on the relaunch of overflowing
  // clusters (status CTK_FAIL_TOO_LARGE) with a larger class.
  const bool rigorous = n_max >= CTK_MAX_CLUSTER_FEATURES || p.capacity_mode == 1;
  const int entry_bytes = big ? 8 : 4;
  lay->mask_words = big ? (n_max + 31) / 32 : 1;
  // the two mask sizes only depend on the radii: remember the last answer (one launch per size
  // class asks again with the same problem)
  static thread_local int memo_key[4] = {-1, -1, -1, -1}, memo_bound = 0, memo_typ = 0;
  const int key[4] = {nd, p.radius[0], nd > 1 ? p.radius[1] : 0, nd > 2 ? p.radius[2] : 0};
  if (memcmp(key, memo_key, sizeof(key)) != 0) {
    memo_bound = mask_capacity(p.radius, nd);
    memo_typ = mask_typical(p.radius, nd);
    memcpy(memo_key, key, sizeof(key));
  }
  const int f_bound = memo_bound;
  int f_typ = memo_typ + 2;
  if (f_typ > f_bound) f_typ = f_bound;
  lay->f_cap = rigorous ? f_bound : f_typ;
  if (lay->f_cap > 16384) return false;
  if (big && (long long) n_max * lay->f_cap > (1 << 24)) return false;
  lay->m_cap = n_max * lay->f_cap;
  if (!rigorous && n_max >= 3) lay->m_cap = (3 * n_max * lay->f_cap + 3) / 4 + 8;
  if (!big && lay->m_cap > 16384) lay->m_cap = 16384;  // pixel index is packed in 14 bits
  lay->pair_cap = n_max > 1 ? n_max * lay->f_cap : 1;
  if (!rigorous && n_max >= 3) lay->pair_cap = (3 * n_max * lay->f_cap + 4) / 5;
  int np = n_max * (n_max - 1) / 2;
  lay->npair_cap = np < 1 ? 1 : (np < 4 * n_max ? np : 4 * n_max);
  if (big) {                      // percolated clusters: several features per pixel are common
    lay->pair_cap = 4 * lay->m_cap;
    lay->npair_cap = np < 24 * n_max ? np : 24 * n_max;
  }
  lay->tab_stride = 0;
  for (int k = 0; k < 3; ++k) {
    lay->tab_len[k] = k < nd ? 2 * p.radius[k] + 3 : 0;
    lay->tab_stride += lay->tab_len[k];
  }
  lay->real_bytes = rb;
  // pixel values are staged in their native width, filtered values (lowpass) in the arithmetic type
  const int px_bytes = p.lowpass ? rb
                       : p.pixel_dtype == CTK_PIXEL_U8 ? 1
                       : (p.pixel_dtype == CTK_PIXEL_U16 || p.pixel_dtype == CTK_PIXEL_I16) ? 2
                       : p.pixel_dtype == CTK_PIXEL_F64 ? 8 : 4;
  int o = 0;
  auto take = [&o](int bytes) { int at = o; o = align_up(o + bytes, 16); return at; };
  // persistent over the whole cluster
  lay->o_x0 = take(v_max * 8);
  lay->o_lo = take(v_max * 8);
  lay->o_hi = take(v_max * 8);
  lay->o_xb = take(p.family != CTK_FAMILY_GAUSS ? v_max * 8 : 0);
  lay->o_cs = take((v_max + 1) * 4);
  lay->o_rc = take(v_max * (v_max + 1) / 2 * (big ? 4 : 2));
  lay->o_cv = take(n_max * p.n_params * 4);
  const int ld_max = p.n_params - 1;
  lay->sidx_stride = ld_max * (ld_max + 1) / 2 + 2 * ld_max;
  lay->o_sidx = take(n_max * lay->sidx_stride * 4);
  lay->o_con = take(16 * 8);
  lay->o_cmode = take(p.n_params * 4);
  lay->o_cbase = take(p.n_params * 4);
  lay->o_ctab = take(p.n_params * 6 * 8);
  lay->o_taps = take(p.lowpass ? 3 * CTK_MAX_TAPS * 8 : 0);
  lay->o_mc = take(n_max * 3 * 8);
  lay->o_fi = take(n_max * FI_STRIDE * 4);
  lay->o_fr = take(n_max * FR_STRIDE * rb);
  // solver scratch; the per-pixel feature bitmasks (needed only while the lists are built) share it
  const int scratch = o;
  lay->o_x = take(v_max * 8);
  lay->o_xt = take(v_max * 8);
  lay->o_rhs = take(v_max * 8);
  lay->o_rhsf = take(v_max * 8);
  lay->o_d = take(v_max * 8);
  lay->o_dg = take(v_max * 8);
  lay->o_act = take(v_max * 4);
  lay->o_H = take(v_max * (v_max + 1) / 2 * rb);
  lay->o_L = take(v_max * (v_max + 1) / 2 * rb);
  lay->o_idg = take(v_max * rb);
  lay->o_pbits = scratch;
  if (o - scratch < lay->m_cap * 4 * lay->mask_words)
    o = align_up(scratch + lay->m_cap * 4 * lay->mask_words, 16);
  lay->o_tmpw = take(big ? 32 * lay->mask_words * 4 : 16);
  // pixel lists
  const int tab_bytes = n_max * lay->tab_stride * 8;
  const int fe_bytes = n_max * lay->f_cap * rb;
  lay->o_fe = take(tab_bytes > fe_bytes ? tab_bytes : fe_bytes);
  lay->o_tab = lay->o_fe;
  lay->o_pval = take(lay->m_cap * px_bytes);
  lay->o_pr = take(lay->m_cap * (rb > 4 ? rb : 4));
  lay->o_pcrd = lay->o_pr;                 // box coordinates are dead once the lists exist
  lay->o_flist = take(n_max * lay->f_cap * entry_bytes);
  lay->o_pairs = take(lay->pair_cap * 4);
  lay->o_phdr = take(lay->npair_cap * 16);
  lay->total = align_up(o, 128);
  return true;
}

inline const char* validate_problem(const ctk_problem_t& p, bool allow_global = false) {
  if (p.ndim != 2 && p.ndim != 3) return "ndim must be 2 or 3";
  if (p.family < CTK_FAMILY_GAUSS || p.family > CTK_FAMILY_DISC) return "unknown family";
  const int ns = p.isotropic ? 1 : p.ndim;
  const int ne = p.family == CTK_FAMILY_GAUSS ? 0 : 1;
  if (p.n_params != 2 + p.ndim + ns + ne) return "n_params does not match ndim/isotropic/family";
  for (int c = 0; c < p.n_params; ++c)
    if (p.modes[c] != CTK_MODE_CONST && p.modes[c] != CTK_MODE_VAR && p.modes[c] != CTK_MODE_CLUSTER &&
        !(allow_global && p.modes[c] == CTK_MODE_GLOBAL))
      return "parameter modes must be const, var or cluster (global: ctk_global_pass only)";
  if (p.modes[0] == CTK_MODE_VAR) return "background cannot vary per feature (fitfunc.py:389-392)";
  for (int k = 0; k < p.ndim; ++k)
    if (p.radius[k] < 1 || p.radius[k] > CTK_MAX_RADIUS) return "radius out of range [1, 30]";
  if (p.pixel_dtype < CTK_PIXEL_U8 || p.pixel_dtype > CTK_PIXEL_I32) return "unknown pixel dtype";
  if (p.compute_dtype != CTK_COMPUTE_F32 && p.compute_dtype != CTK_COMPUTE_F64)
    return "unknown compute dtype";
  if (p.max_iter < 1 || p.lm_max_iter < 1) return "iteration limits must be positive";
  if (p.lowpass)
    for (int k = 0; k < p.ndim; ++k)
      if (p.lowpass_half[k] < -1 || 2 * p.lowpass_half[k] + 1 > CTK_MAX_TAPS)
        return "lowpass kernel too wide (half width must be <= 16)";
  if (!(p.residual_factor > 0.)) return "residual_factor must be positive";
  if (p.constraint_mask & CTK_CONSTRAINT_DIMER)
    for (int k = 0; k < p.ndim; ++k) if (!(p.dimer_dist[k] > 0.)) return "dimer distance must be > 0";
  if (p.constraint_mask & CTK_CONSTRAINT_TRIMER)
    for (int k = 0; k < p.ndim; ++k) if (!(p.trimer_dist[k] > 0.)) return "trimer distance must be > 0";
  if (p.constraint_mask & CTK_CONSTRAINT_TETRAMER)
    for (int k = 0; k < p.ndim; ++k)
      if (!(p.tetramer_dist[k] > 0.)) return "tetramer distance must be > 0";
  return nullptr;
}

// Runs `fn.template operator()<Config>()` for the kernel instance matching the problem.
// size_free / extra_free select the instances that carry derivative slots for those columns.
template <class Real, class Fn>
inline bool dispatch_config(const ctk_problem_t& p, Fn&& fn, bool big = false) {
  const int nd = p.ndim;
  if (big) {
    // large clusters: one instance per (geometry, family) that carries every derivative slot;
    // constant columns are simply not scattered
#define CTK_BIG_CASE(ND, ISO, FAM)                                                            \
  if (nd == ND && (p.isotropic != 0) == ISO && p.family == FAM) {                             \
    fn.template operator()<Config<Real, ND, ISO, FAM, true, FAM != CTK_FAMILY_GAUSS, true> >(); \
    return true;                                                                              \
  }
#define CTK_BIG_GEOM(FAM) CTK_BIG_CASE(2, true, FAM) CTK_BIG_CASE(2, false, FAM)              \
                          CTK_BIG_CASE(3, true, FAM) CTK_BIG_CASE(3, false, FAM)
    CTK_BIG_GEOM(CTK_FAMILY_GAUSS) CTK_BIG_GEOM(CTK_FAMILY_RING) CTK_BIG_GEOM(CTK_FAMILY_DISC)
#undef CTK_BIG_GEOM
#undef CTK_BIG_CASE
    return false;
  }
  const int ns = p.isotropic ? 1 : nd;
  const bool extra = p.constraint_mask != 0 || p.lowpass != 0;   // else: the lean instances
  bool size_free = false;
  for (int k = 0; k < ns; ++k) size_free |= p.modes[2 + nd + k] != CTK_MODE_CONST;
  const bool extra_free = p.family != CTK_FAMILY_GAUSS && p.modes[2 + nd + ns] != CTK_MODE_CONST;
#define CTK_CASE(ND, ISO, FAM, SZ, EX)                                                       \
  if (nd == ND && (p.isotropic != 0) == ISO && p.family == FAM && size_free == SZ &&         \
      extra_free == EX) {                                                                    \
    if (extra) fn.template operator()<Config<Real, ND, ISO, FAM, SZ, EX, false, true> >();   \
    else fn.template operator()<Config<Real, ND, ISO, FAM, SZ, EX, false, false> >();        \
    return true;                                                                             \
  }
#define CTK_CASES_GEOM(FAM, SZ, EX)                                                          \
  CTK_CASE(2, true, FAM, SZ, EX) CTK_CASE(2, false, FAM, SZ, EX)                             \
  CTK_CASE(3, true, FAM, SZ, EX) CTK_CASE(3, false, FAM, SZ, EX)
  CTK_CASES_GEOM(CTK_FAMILY_GAUSS, false, false)
  CTK_CASES_GEOM(CTK_FAMILY_GAUSS, true, false)
  CTK_CASES_GEOM(CTK_FAMILY_RING, false, false)
  CTK_CASES_GEOM(CTK_FAMILY_RING, true, false)
  CTK_CASES_GEOM(CTK_FAMILY_RING, false, true)
  CTK_CASES_GEOM(CTK_FAMILY_RING, true, true)
  CTK_CASES_GEOM(CTK_FAMILY_DISC, false, false)
  CTK_CASES_GEOM(CTK_FAMILY_DISC, true, false)
  CTK_CASES_GEOM(CTK_FAMILY_DISC, false, true)
  CTK_CASES_GEOM(CTK_FAMILY_DISC, true, true)
#undef CTK_CASES_GEOM
#undef CTK_CASE
  return false;
}

}  // namespace ctk
