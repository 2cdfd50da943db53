// ctk_thread.cuh -- one THREAD refines one cluster.
//
// The warp-per-cluster solver (ctk_solver.cuh) spends most of its instructions on what a warp needs
// to cooperate on a problem of 7..19 unknowns and a few hundred pixels: reductions, barriers,
// packed index tables, lanes idling in the factorisation.  For the path's dominant case -- the
// reference's default model (2D isotropic gauss, `signal` and position free per feature, one
// background per cluster, constant size; refine.py:82-87 defaults, fitfunc.py:356-387), clusters of
// up to 8 features, no constraints -- this file runs the SAME algorithm with one thread per cluster:
// every sum is a private register, every vector a private (interleaved, coalesced) local array, no
// barrier anywhere.  32 clusters advance per warp instruction instead of one.
//
// Per-thread state lives in local memory, which the hardware interleaves across the lanes of a warp:
// an access is coalesced only when every lane uses the SAME index.  The data structures are
// therefore per-feature ENTRY LISTS walked by a loop counter (packed pixel offset from the integer
// mask centre + the set of features covering the pixel, the pixel value); nothing in
// the iteration is addressed by pixel position.  One fused pass per Levenberg-Marquardt iteration
// computes the objective and the normal equations at the trial point into a spare
// buffer that is committed when the step is accepted.  The rules (pixel sets in float64 without FMA
// contraction, objective, bounds, LM loop with chord iterations, outer re-mask loop, statuses)
// follow ctk_solver.cuh function by function.
//
// Clusters that do not fit (bounding box above CTK_T_BOX pixels) report CTK_FAIL_TOO_LARGE and are
// relaunched by the caller with capacity_mode = 1, which always selects the warp kernel.
#pragma once

#include "ctk_solver.cuh"

namespace ctk {

#define CTK_T_NMAX 8                          // features per cluster
#define CTK_T_VMAX (1 + 3 * CTK_T_NMAX)       // background + (signal, y, x) per feature
#define CTK_T_TRI (CTK_T_VMAX * (CTK_T_VMAX + 1) / 2)
#define CTK_T_RMAX 15                         // mask radius
#define CTK_T_TAB (2 * CTK_T_RMAX + 3)
#define CTK_T_BOX 1600                        // pixels of the cluster's bounding box
#define CTK_T_ECAP 832                        // entries (feature, mask pixel) of a cluster
#define CTK_T_BLOCK 128                       // threads per block (sizes the shared constants)

// host + device: does this launch qualify for the thread-per-cluster kernel?
inline bool thread_eligible(const ctk_problem_t& p, int n_max) {
  if (p.capacity_mode != 0) return false;     // relaunches of overflowing clusters: warp kernel
  if (p.ndim != 2 || !p.isotropic || p.family != CTK_FAMILY_GAUSS || p.n_params != 5) return false;
  if (p.lowpass || p.constraint_mask) return false;
  if (p.modes[0] != CTK_MODE_CLUSTER || p.modes[1] != CTK_MODE_VAR || p.modes[2] != CTK_MODE_VAR ||
      p.modes[3] != CTK_MODE_VAR || p.modes[4] != CTK_MODE_CONST)
    return false;
  if (p.radius[0] > CTK_T_RMAX || p.radius[1] > CTK_T_RMAX) return false;
  return n_max >= 1 && n_max <= CTK_T_NMAX;
}

template <class Real>
struct ThreadSolver {
  enum { P = 5 };
  const BatchArgs& a;

  int n, feat0, V, M, E, Q;
  int blo[2], bdim[2], total;
  const void* frame;
  int evals, accums, grad_accums, outers;
  int curH, curR;                                      // committed halves of the double buffers

  double x[CTK_T_VMAX], xt[CTK_T_VMAX], x0[CTK_T_VMAX], lo[CTK_T_VMAX], hi[CTK_T_VMAX];
  double rb[2][CTK_T_VMAX], rhsf[CTK_T_VMAX], d[CTK_T_VMAX], sc[CTK_T_VMAX];
  unsigned char act[CTK_T_VMAX];
  Real Hb[2][CTK_T_TRI], L[CTK_T_TRI], idg[CTK_T_VMAX];   // packed lower triangles, row-major
  double mc[CTK_T_NMAX][2];
  int ci[CTK_T_NMAX][2], ts[CTK_T_NMAX][2];
  int estart[CTK_T_NMAX + 1], eshared[CTK_T_NMAX];     // feature i: exclusive entries, then shared
  uint32_t EL[CTK_T_ECAP];                             // (dy + 32) | (dx + 32) << 6 | cover bits << 12
  Real Iv[CTK_T_ECAP];                                 // pixel value of every entry
  // build-only scratch: cover bits of the box pixels, column spans of the mask rows
  unsigned char B[CTK_T_BOX];
  short sp_lo[CTK_T_NMAX][CTK_T_TAB], sp_hi[CTK_T_NMAX][CTK_T_TAB];

  // Per-feature constants of the pixel pass (signal, frac y, frac x, 1/size, centre offsets to the
  // feature whose entries are walked): small, hot, indexed by feature -> shared memory, one
  // column per thread (bank = thread), instead of local memory.
#ifdef CTK_EMUL
  Real fc_[6][CTK_T_NMAX];
  CTK_DEV Real& fc(int c, int j) { return fc_[c][j]; }
#else
  Real* fc_;
  CTK_DEV Real& fc(int c, int j) { return fc_[(c * CTK_T_NMAX + j) * CTK_T_BLOCK]; }
#endif

#ifdef CTK_EMUL
  CTK_DEV explicit ThreadSolver(const BatchArgs& args) : a(args) {}
#else
  CTK_DEV ThreadSolver(const BatchArgs& args, Real* shared_column) : a(args), fc_(shared_column) {}
#endif

  CTK_DEV static int tri_at(int r, int c) { return r * (r + 1) / 2 + c; }          // r >= c
  CTK_DEV static int sym_at(int u, int v) { return u >= v ? tri_at(u, v) : tri_at(v, u); }
  CTK_DEV int var_s(int i) const { return 1 + i; }
  CTK_DEV int var_y(int i) const { return 1 + n + i; }
  CTK_DEV int var_x(int i) const { return 1 + 2 * n + i; }
  CTK_DEV int var_of(int i, int u) const { return 1 + u * n + i; }                 // u: 0 s, 1 y, 2 x
  CTK_DEV bool is_pos_var(int v) const { return v > n; }

  CTK_DEV Real frame_value(int64_t i) const {
    switch (a.prob.pixel_dtype) {
      case CTK_PIXEL_U8: return (Real) reinterpret_cast<const uint8_t*>(frame)[i];
      case CTK_PIXEL_U16: return (Real) reinterpret_cast<const uint16_t*>(frame)[i];
      case CTK_PIXEL_F32: return (Real) reinterpret_cast<const float*>(frame)[i];
      case CTK_PIXEL_F64: return (Real) reinterpret_cast<const double*>(frame)[i];
      case CTK_PIXEL_I16: return (Real) reinterpret_cast<const int16_t*>(frame)[i];
      default: return (Real) reinterpret_cast<const int32_t*>(frame)[i];
    }
  }

  // ---- variables, start vector and bounds (fitfunc.py:207-263, 535-558) ------------------------
  CTK_DEV int setup_variables() {
    V = 1 + 3 * n;
    const double* pin = a.params_in + (int64_t) feat0 * P;
    const bool tables = a.lo_in == nullptr;
    const double* lin = tables ? nullptr : a.lo_in + (int64_t) feat0 * P;
    const double* hin = tables ? nullptr : a.hi_in + (int64_t) feat0 * P;
    bool bad = false;
    double bsum = 0., blow = INFINITY, bhigh = -INFINITY;
    for (int i = 0; i < n; ++i) {
      for (int c = 0; c < P; ++c) bad |= !finite_d(pin[i * P + c]);
      // background: one entry per cluster, mean start, widest bound
      {
        const double p = pin[i * P];
        bsum += p;
        blow = fmin(blow, tables ? bound_from_tables(p, a.prob.bounds_diff[0][0], a.prob.bounds_rel[0][0],
                                                     a.prob.bounds_abs[0][0], 0) : lin[i * P]);
        bhigh = fmax(bhigh, tables ? bound_from_tables(p, a.prob.bounds_diff[1][0], a.prob.bounds_rel[1][0],
                                                       a.prob.bounds_abs[1][0], 1) : hin[i * P]);
      }
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int c = 1 + u, v = var_of(i, u);
        const double p = pin[i * P + c];
        x0[v] = p;
        if (tables) {
          const double dl = u == 0 ? a.prob.bounds_diff[0][1] : (u == 1 ? a.prob.bounds_diff[0][2] : a.prob.bounds_diff[0][3]);
          const double rl = u == 0 ? a.prob.bounds_rel[0][1] : (u == 1 ? a.prob.bounds_rel[0][2] : a.prob.bounds_rel[0][3]);
          const double al = u == 0 ? a.prob.bounds_abs[0][1] : (u == 1 ? a.prob.bounds_abs[0][2] : a.prob.bounds_abs[0][3]);
          const double dh = u == 0 ? a.prob.bounds_diff[1][1] : (u == 1 ? a.prob.bounds_diff[1][2] : a.prob.bounds_diff[1][3]);
          const double rh = u == 0 ? a.prob.bounds_rel[1][1] : (u == 1 ? a.prob.bounds_rel[1][2] : a.prob.bounds_rel[1][3]);
          const double ah = u == 0 ? a.prob.bounds_abs[1][1] : (u == 1 ? a.prob.bounds_abs[1][2] : a.prob.bounds_abs[1][3]);
          lo[v] = bound_from_tables(p, dl, rl, al, 0);
          hi[v] = bound_from_tables(p, dh, rh, ah, 1);
        } else {
          lo[v] = lin[i * P + c];
          hi[v] = hin[i * P + c];
        }
      }
      fc(3, i) = (Real) (1.0 / pin[i * P + 4]);
    }
    if (bad) return CTK_FAIL_NONFINITE;
    x0[0] = bsum / n; lo[0] = blow; hi[0] = bhigh;
    for (int v = 0; v < V; ++v) {
      if (!(lo[v] <= hi[v])) return CTK_FAIL_BOUNDS;
      x0[v] = fmin(fmax(x0[v], lo[v]), hi[v]);             // scipy clips the start into the box
    }
    return CTK_OK;
  }

  // ---- pixel set (refine.py:28-58, masks.py:30-68) as per-feature entry lists -------------------
  CTK_DEV int build_pixels() {
    int mn[2] = {INT32_MAX, INT32_MAX}, mx[2] = {INT32_MIN, INT32_MIN};
    for (int i = 0; i < n; ++i) {
      bool inb = true;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        double r = rint(mc[i][k]);
        r = fmin(fmax(r, -1.0e9), 1.0e9);
        ci[i][k] = (int) r;
        const int rad = k == 0 ? a.prob.radius[0] : a.prob.radius[1];
        const int shp = (int) (k == 0 ? a.shape[0] : a.shape[1]);
        inb = inb && (ci[i][k] >= -rad) && (ci[i][k] < shp + rad);
      }
      if (inb) {
#pragma unroll
        for (int k = 0; k < 2; ++k) { mn[k] = min(mn[k], ci[i][k]); mx[k] = max(mx[k], ci[i][k]); }
      }
    }
    if (mn[0] == INT32_MAX) return CTK_FAIL_OUT_OF_IMAGE;
    int64_t tot = 1;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int rad = k == 0 ? a.prob.radius[0] : a.prob.radius[1];
      const int shp = (int) (k == 0 ? a.shape[0] : a.shape[1]);
      blo[k] = max(0, mn[k] - rad);
      const int bhi = min(shp, mx[k] + rad + 1);
      bdim[k] = bhi - blo[k];
      tot *= bdim[k];
    }
    if (tot <= 0 || tot > CTK_T_BOX) return CTK_FAIL_TOO_LARGE;
    total = (int) tot;
    for (int p = 0; p < total; ++p) B[p] = 0;
    const int len0 = 2 * a.prob.radius[0] + 3, len1 = 2 * a.prob.radius[1] + 3;
    for (int i = 0; i < n; ++i) {
      // separable tables ((idx - (c - origin)) / r)^2 in float64, numpy's operation order
      double tab1[CTK_T_TAB];
      const double crel0 = dsub(mc[i][0], (double) blo[0]), crel1 = dsub(mc[i][1], (double) blo[1]);
      {
        double s0 = floor(crel0 - (double) a.prob.radius[0]), s1 = floor(crel1 - (double) a.prob.radius[1]);
        s0 = fmin(fmax(s0, -1.0e9), 1.0e9);
        s1 = fmin(fmax(s1, -1.0e9), 1.0e9);
        ts[i][0] = (int) s0;
        ts[i][1] = (int) s1;
      }
      for (int e = 0; e < len1; ++e) {
        const double q = ddiv(dsub((double) (ts[i][1] + e), crel1), (double) a.prob.radius[1]);
        tab1[e] = dmul(q, q);
      }
      for (int e0 = 0; e0 < len0; ++e0) {
        const int cy = ts[i][0] + e0;
        int first = 1, last = 0;
        if (cy >= 0 && cy < bdim[0]) {
          const double q = ddiv(dsub((double) cy, crel0), (double) a.prob.radius[0]);
          const double t0 = dmul(q, q);
          for (int e1 = 0; e1 < len1; ++e1) {
            const int cx = ts[i][1] + e1;
            if (cx < 0 || cx >= bdim[1]) continue;
            if (dadd(t0, tab1[e1]) <= 1.0) {
              if (first > last) first = cx;
              last = cx;
              B[cy * bdim[1] + cx] |= (unsigned char) (1u << i);
            }
          }
        }
        sp_lo[i][e0] = (short) first;
        sp_hi[i][e0] = (short) last;
      }
    }
    // entry lists: feature i's mask pixels with the cover set and the pixel value; the pixels only
    // feature i covers come first, the pixels it shares with other features after them (so that
    // the lanes of a warp do not drag each other through the shared-pixel code)
    E = 0; M = 0; Q = 0;
    for (int i = 0; i < n; ++i) {
      estart[i] = E;
      const int oy = ci[i][0] - blo[0], ox = ci[i][1] - blo[1];      // integer centre in box coords
      const unsigned self = 1u << i, lower = self - 1u;
      for (int sweep = 0; sweep < 2; ++sweep) {
        if (sweep == 1) eshared[i] = E;
        for (int e0 = 0; e0 < len0; ++e0) {
          const int c0 = sp_lo[i][e0], c1 = sp_hi[i][e0];
          if (c1 < c0) continue;
          const int cy = ts[i][0] + e0;
          if (E + (c1 - c0 + 1) > CTK_T_ECAP) return CTK_FAIL_TOO_LARGE;
          const int64_t row = (int64_t) (blo[0] + cy) * a.shape[1] + blo[1];
          for (int cx = c0; cx <= c1; ++cx) {
            const unsigned bits = B[cy * bdim[1] + cx];
            if ((bits == self) != (sweep == 0)) continue;
            EL[E] = (uint32_t) (cy - oy + 32) | ((uint32_t) (cx - ox + 32) << 6) | (bits << 12);
            Iv[E] = frame_value(row + cx);
            ++E;
            if (!(bits & lower)) {                         // the pixel's first feature counts it
              ++M;
              const int k = popc(bits);
              Q += k * (k - 1) / 2;
            }
          }
        }
      }
    }
    estart[n] = E;
    if (M == 0) return CTK_FAIL_OUT_OF_IMAGE;
    return CTK_OK;
  }

  // ---- one fused pass at `v`: objective 0.5 sum r^2 (fitfunc.py:436-450 without 1/M/norm),
  // -gradient (always) and the normal matrix (unless grad_only) into the SPARE halves of the
  // double buffers.
  CTK_DEV double pass(const double* v, bool grad_only) {
    ++evals;
    Real* Hn = Hb[curH ^ 1];
    double* rn = rb[curR ^ 1];
    const int nt = tri_at(V - 1, V - 1) + 1;
    if (!grad_only) for (int t = 0; t < nt; ++t) Hn[t] = 0;
    for (int u = 0; u < V; ++u) rn[u] = 0.;
    for (int i = 0; i < n; ++i) {
      fc(0, i) = (Real) v[var_s(i)];
      fc(1, i) = (Real) (v[var_y(i)] - (double) ci[i][0]);
      fc(2, i) = (Real) (v[var_x(i)] - (double) ci[i][1]);
    }
    const Real bg = (Real) v[0];
    double acc = 0., sr = 0.;
    for (int i = 0; i < n; ++i) {
      const Real s = fc(0, i), fy = fc(1, i), fx = fc(2, i), is = fc(3, i);
      const Real w2 = (Real) 2 * s * is;                   // W q is = (s ndim g) (d is) is
      for (int j = 0; j < n; ++j) {                        // integer centre of i relative to j
        fc(4, j) = (Real) (ci[i][0] - ci[j][0]);
        fc(5, j) = (Real) (ci[i][1] - ci[j][1]);
      }
      const unsigned self = 1u << i, lower = self - 1u;
      Real a00 = 0, a10 = 0, a11 = 0, a20 = 0, a21 = 0, a22 = 0, s0 = 0, s1 = 0, s2 = 0, g0 = 0, g1 = 0, g2 = 0;
      // pixels of this feature alone
      const int t_sh = eshared[i], t1 = estart[i + 1];
#pragma unroll 4
      for (int t = estart[i]; t < t_sh; ++t) {
        const uint32_t e = EL[t];
        const Real dy = (Real) ((int) (e & 63u) - 32), dx = (Real) ((int) ((e >> 6) & 63u) - 32);
        const Real qy = (dy - fy) * is, qx = (dx - fx) * is;
        const Real gv = fast_exp(-(qy * qy + qx * qx));     // exp(-0.5 ndim r2), ndim = 2
        const Real r = Iv[t] - bg - s * gv;
        acc += (double) r * (double) r;
        sr += (double) r;
        const Real m0 = gv, wg = w2 * gv, m1 = wg * qy, m2 = wg * qx;
        g0 += m0 * r; g1 += m1 * r; g2 += m2 * r;
        if (!grad_only) {
          s0 += m0; s1 += m1; s2 += m2;
          a00 += m0 * m0; a10 += m1 * m0; a11 += m1 * m1;
          a20 += m2 * m0; a21 += m2 * m1; a22 += m2 * m2;
        }
      }
      // pixels shared with other features: their model values are evaluated here as well (the
      // per-feature constants sit in shared memory, so the feature index can be a run-time value)
#pragma unroll 2
      for (int t = t_sh; t < t1; ++t) {
        const uint32_t e = EL[t];
        const Real dy = (Real) ((int) (e & 63u) - 32), dx = (Real) ((int) ((e >> 6) & 63u) - 32);
        const unsigned bits = (e >> 12) & 255u;
        const Real qy = (dy - fy) * is, qx = (dx - fx) * is;
        const Real gv = fast_exp(-(qy * qy + qx * qx));
        Real r = Iv[t] - bg - s * gv;
        unsigned others = bits & ~self;
        while (others) {
          const int j = ctz(others);
          others &= others - 1u;
          const Real isj = fc(3, j);
          const Real qyj = (dy + fc(4, j) - fc(1, j)) * isj, qxj = (dx + fc(5, j) - fc(2, j)) * isj;
          r -= fc(0, j) * fast_exp(-(qyj * qyj + qxj * qxj));
        }
        if (!(bits & lower)) { acc += (double) r * (double) r; sr += (double) r; }
        const Real m0 = gv, wg = w2 * gv, m1 = wg * qy, m2 = wg * qx;
        g0 += m0 * r; g1 += m1 * r; g2 += m2 * r;
        if (!grad_only) {
          s0 += m0; s1 += m1; s2 += m2;
          a00 += m0 * m0; a10 += m1 * m0; a11 += m1 * m1;
          a20 += m2 * m0; a21 += m2 * m1; a22 += m2 * m2;
        }
      }
      const int vs = var_s(i), vy = var_y(i), vx = var_x(i);
      rn[vs] += (double) g0; rn[vy] += (double) g1; rn[vx] += (double) g2;
      if (!grad_only) {
        Hn[tri_at(vs, 0)] += s0; Hn[tri_at(vy, 0)] += s1; Hn[tri_at(vx, 0)] += s2;
        Hn[tri_at(vs, vs)] += a00; Hn[tri_at(vy, vs)] += a10; Hn[tri_at(vy, vy)] += a11;
        Hn[tri_at(vx, vs)] += a20; Hn[tri_at(vx, vy)] += a21; Hn[tri_at(vx, vx)] += a22;
      }
    }
    rn[0] = sr;
    if (grad_only) { ++grad_accums; return 0.5 * acc; }
    ++accums;
    Hn[0] = (Real) M;
    // cross blocks over the pixels two features share
    for (int i = 0; i < n - 1; ++i) {
      for (int j = i + 1; j < n; ++j) {
        if (abs(ci[i][0] - ci[j][0]) > 2 * a.prob.radius[0] + 2 ||
            abs(ci[i][1] - ci[j][1]) > 2 * a.prob.radius[1] + 2)
          continue;
        const Real si = fc(0, i), fyi = fc(1, i), fxi = fc(2, i), isi = fc(3, i);
        const Real sj = fc(0, j), isj = fc(3, j);
        const Real fyj = fc(1, j) - (Real) (ci[i][0] - ci[j][0]), fxj = fc(2, j) - (Real) (ci[i][1] - ci[j][1]);
        const Real w2i = (Real) 2 * si * isi, w2j = (Real) 2 * sj * isj;
        Real b[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) b[k] = 0;
        bool any = false;
        const int t1 = estart[i + 1];
        for (int t = eshared[i]; t < t1; ++t) {
          const uint32_t e = EL[t];
          if (!((e >> (12 + j)) & 1u)) continue;
          any = true;
          const Real dy = (Real) ((int) (e & 63u) - 32), dx = (Real) ((int) ((e >> 6) & 63u) - 32);
          const Real qyi = (dy - fyi) * isi, qxi = (dx - fxi) * isi;
          const Real qyj = (dy - fyj) * isj, qxj = (dx - fxj) * isj;
          const Real gi = fast_exp(-(qyi * qyi + qxi * qxi)), gj = fast_exp(-(qyj * qyj + qxj * qxj));
          (void) si;
          const Real mi[3] = {gi, w2i * gi * qyi, w2i * gi * qxi};
          const Real mj[3] = {gj, w2j * gj * qyj, w2j * gj * qxj};
#pragma unroll
          for (int u = 0; u < 3; ++u)
#pragma unroll
            for (int w = 0; w < 3; ++w) b[u * 3 + w] += mi[u] * mj[w];
        }
        if (!any) continue;
#pragma unroll
        for (int u = 0; u < 3; ++u)
#pragma unroll
          for (int w = 0; w < 3; ++w) Hn[sym_at(var_of(i, u), var_of(j, w))] += b[u * 3 + w];
      }
    }
    return 0.5 * acc;
  }

  // ---- damped, bound-aware step (see ClusterSolver::solve) --------------------------------------
  CTK_DEV bool solve(double lambda, bool reuse) {
    const Real* H = Hb[curH];
    const double* rhs = rb[curR];
    for (int v = 0; v < V; ++v) rhsf[v] = rhs[v];
    if (reuse) {
      bool moved = false;
      for (int v = 0; v < V; ++v) {
        const double g = rhsf[v];
        const bool frozen = (x[v] <= lo[v] && g < 0.) || (x[v] >= hi[v] && g > 0.) || !(lo[v] < hi[v]);
        moved |= frozen != (act[v] != 0);
        d[v] = frozen ? 0. : g * sc[v];
      }
      if (moved) reuse = false;
    }
    if (reuse) {
      for (int j = 0; j < V; ++j) {                        // forward substitution with L
        const double yj = d[j] * (double) idg[j];
        d[j] = yj;
        for (int r = j + 1; r < V; ++r) d[r] -= (double) L[tri_at(r, j)] * yj;
      }
    } else {
      double dmax = 0.;
      for (int v = 0; v < V; ++v) dmax = fmax(dmax, (double) H[tri_at(v, v)]);
      const double floor_ = fmax(dmax * 1e-14, 1e-30);
      for (int v = 0; v < V; ++v) {
        const double g = rhsf[v];
        const bool frozen = (x[v] <= lo[v] && g < 0.) || (x[v] >= hi[v] && g > 0.) || !(lo[v] < hi[v]);
        act[v] = frozen ? 1 : 0;
        const double s = (double) fast_rsqrt((Real) fmax((double) H[tri_at(v, v)], floor_));
        sc[v] = s;
        d[v] = frozen ? 0. : g * s;
      }
      // scaled, damped system: (S K S + lambda I)(S^-1 step) = S rhs; frozen rows become identity
      const Real lam1 = (Real) (1. + lambda);
      for (int r = 0; r < V; ++r) {
        for (int c = 0; c <= r; ++c) {
          Real v;
          if (act[r] || act[c]) v = (r == c) ? (Real) 1 : (Real) 0;
          else if (r == c) v = lam1;
          else v = H[tri_at(r, c)] * (Real) (sc[r] * sc[c]);
          L[tri_at(r, c)] = v;
        }
      }
      for (int j = 0; j < V; ++j) {
        const Real piv = L[tri_at(j, j)];
        if (!(piv > (Real) 1e-7)) return false;              // also catches NaN
        const Real inv = fast_rsqrt(piv);
        const double yj = d[j] * (double) inv;               // forward substitution, row j
        for (int r = j; r < V; ++r) L[tri_at(r, j)] *= inv;
        idg[j] = inv;
        d[j] = yj;
        for (int c = j + 1; c < V; ++c) {
          const Real lcj = L[tri_at(c, j)];
          for (int r = c; r < V; ++r) L[tri_at(r, c)] -= L[tri_at(r, j)] * lcj;
        }
        for (int r = j + 1; r < V; ++r) d[r] -= (double) L[tri_at(r, j)] * yj;
      }
    }
    for (int j = V - 1; j >= 0; --j) {                     // back substitution with L^T
      const double zj = d[j] * (double) idg[j];
      d[j] = zj;
      for (int r = 0; r < j; ++r) d[r] -= (double) L[tri_at(j, r)] * zj;
    }
    for (int v = 0; v < V; ++v) d[v] *= sc[v];
    return true;
  }

  // predicted decrease of the objective for step s: rhs.s - 0.5 s^T H s
  CTK_DEV double predicted(const double* s) const {
    const Real* H = Hb[curH];
    double acc = 0.;
    for (int r = 0; r < V; ++r) {
      for (int c = 0; c < r; ++c) acc -= (double) H[tri_at(r, c)] * s[r] * s[c];
      acc -= 0.5 * (double) H[tri_at(r, r)] * s[r] * s[r];
      acc += s[r] * rhsf[r];
    }
    return acc;
  }

  // ---- projected Levenberg-Marquardt (ClusterSolver::minimise without constraint rows) ----------
  CTK_DEV int minimise(double* f_data) {
    const bool f32 = sizeof(Real) == 4;
    const double xtol = a.prob.xtol > 0. ? a.prob.xtol : (f32 ? 2e-6 : 1e-9);
    const double eps_f = f32 ? 4e-6 : 1e-13;
    const double chord_tol = a.prob.chord_tol;
    double lambda = 1e-3, nu = 2.;
    for (int v = 0; v < V; ++v) xt[v] = x[v];
    double fd = 0., fa = 0., pred = 0., worst = 0.;
    int rejects = 0;
    double prev_small_step = INFINITY, last_step = INFINITY;
    bool force = true, need_eval = true, chord_next = false;
    curH = curR = 0;
    for (int it = 0; it <= a.prob.lm_max_iter; ++it) {
      if (need_eval) {
        // the normal equations at the trial point are computed along with the objective and
        // committed below if the step is accepted; close to the minimum only the gradient is
        // refreshed and the factorised matrix reused (chord iteration)
        const bool cheap = !force && worst < chord_tol;
        const double fat = pass(xt, cheap);
        bool accept;
        if (force) {
          if (!finite_d(fat)) return CTK_FAIL_NUMERIC;
          accept = true;
        } else {
          const bool noise = pred > 0. && pred <= eps_f * fabs(fa) &&
                             fabs(fat - fa) <= 8. * eps_f * fabs(fa);
          accept = finite_d(fat) && pred > 0. && (fat < fa || noise);
          if (accept) {
            if (!noise) {
              const double rho = (fa - fat) / pred;
              const double t = 2. * rho - 1.;
              lambda = fmax(lambda * fmax(1. / 3., 1. - t * t * t), 1e-12);
            } else {
              if (worst > 0.9 * prev_small_step) lambda *= 4.;
              prev_small_step = worst;
            }
            nu = 2.;
            rejects = 0;
            last_step = worst;
          } else {
            lambda *= nu;
            nu *= 2.;
            if (++rejects > 40 || lambda > 1e18) { *f_data = fd; return CTK_OK; }
            if (chord_next) {
              for (int v = 0; v < V; ++v) xt[v] = x[v];
              force = true;
              chord_next = false;
              continue;
            }
          }
        }
        if (accept) {
          for (int v = 0; v < V; ++v) x[v] = xt[v];
          fd = fa = fat;
          curR ^= 1;
          if (!cheap) curH ^= 1;
          chord_next = cheap;
          force = false;
        }
      }
      need_eval = true;
      if (!solve(lambda, chord_next)) {
        lambda = fmax(lambda * 10., 1e-8);
        if (++rejects > 60) { *f_data = fd; return CTK_FAIL_NUMERIC; }
        need_eval = false;
        continue;
      }
      worst = 0.;
      for (int v = 0; v < V; ++v) {
        const double t = fmin(fmax(x[v] + d[v], lo[v]), hi[v]);
        xt[v] = t;
        const double s = t - x[v];
        d[v] = s;
        const float scale = is_pos_var(v) ? 1.f : fmaxf(1.f, fabsf((float) x[v]));
        worst = fmax(worst, (double) (fabsf((float) s) / scale));
      }
      if (!finite_d(worst)) { *f_data = fd; return CTK_FAIL_NUMERIC; }
      if (worst <= xtol) { *f_data = fd; return CTK_OK; }
      pred = predicted(d);
    }
    *f_data = fd;
    return last_step <= 1e-3 ? CTK_OK : CTK_FAIL_NO_CONVERGENCE;
  }

  // ---- whole cluster (refine.py:343-430) -------------------------------------------------------
  CTK_DEV void run(int cluster) {
    feat0 = a.cluster_offset[cluster];
    n = a.cluster_offset[cluster + 1] - feat0;
    const int fidx = a.cluster_frame[cluster];
    frame = a.frames[fidx];
    const double fmax_ = a.frame_max[fidx];
    evals = accums = grad_accums = outers = 0;
    M = E = Q = 0;
    V = 0;
    int status = CTK_OK;
    double cost = NAN, fd = 0.;
    if (n <= 0 || n > CTK_T_NMAX || n > a.lay.n_max) status = CTK_FAIL_TOO_LARGE;
    if (status == CTK_OK) status = setup_variables();
    if (status == CTK_OK) {
      for (int i = 0; i < n; ++i) {
        mc[i][0] = a.params_in[(int64_t) (feat0 + i) * P + 2];
        mc[i][1] = a.params_in[(int64_t) (feat0 + i) * P + 3];
      }
      for (int outer = 0; outer < a.prob.max_iter; ++outer) {
        ++outers;
        status = build_pixels();
        if (status != CTK_OK) break;
        for (int v = 0; v < V; ++v) x[v] = x0[v];            // restart, refine.py:361-365
        status = minimise(&fd);
        if (status != CTK_OK) break;
        bool moved = false;                                   // refine.py:383-385
        for (int i = 0; i < n; ++i) {
          const double dy = x[var_y(i)] - mc[i][0], dx = x[var_x(i)] - mc[i][1];
          moved |= !(dy * dy + dx * dx < a.prob.max_shift * a.prob.max_shift);
        }
        if (!moved) break;
        for (int i = 0; i < n; ++i) { mc[i][0] = x[var_y(i)]; mc[i][1] = x[var_x(i)]; }
      }
    }
    if (status == CTK_OK) {
      const double norm = fmax_ * fmax_ / a.prob.residual_factor;       // refine.py:354, 379
      const double fun = 2. * fd / (double) M / norm;
      cost = sqrt(fun / a.prob.residual_factor);
      if (!finite_d(cost)) status = CTK_FAIL_NUMERIC;
      else if (cost > a.prob.max_rms_dev) status = CTK_FAIL_RMS_DEV;
    }
    if (n > 0) {                                   // refine.py:408-427: failures keep their input
      const double* pin = a.params_in + (int64_t) feat0 * P;
      double* pout = a.params_out + (int64_t) feat0 * P;
      const bool ok = status == CTK_OK;
      for (int i = 0; i < n; ++i) {
        pout[i * P + 0] = ok ? x[0] : pin[i * P + 0];
        pout[i * P + 1] = ok ? x[var_s(i)] : pin[i * P + 1];
        pout[i * P + 2] = ok ? x[var_y(i)] : pin[i * P + 2];
        pout[i * P + 3] = ok ? x[var_x(i)] : pin[i * P + 3];
        pout[i * P + 4] = pin[i * P + 4];
      }
    }
    a.cost_out[cluster] = status == CTK_OK ? cost : NAN;
    a.status_out[cluster] = status;
    int32_t* st = a.stats_out + (int64_t) cluster * CTK_STATS;
    st[CTK_STAT_EVALS] = evals; st[CTK_STAT_ACCUMS] = accums; st[CTK_STAT_OUTER] = outers;
    st[CTK_STAT_PIXELS] = M; st[CTK_STAT_ENTRIES] = E; st[CTK_STAT_PAIR_ENTRIES] = Q;
    st[CTK_STAT_VARS] = V; st[CTK_STAT_GRAD_ACCUMS] = grad_accums;
    if (status == CTK_FAIL_TOO_LARGE && a.overflow != nullptr) {
      const int k = atomic_next(a.overflow);
      if (k < a.overflow_cap) a.overflow[1 + k] = cluster;
    }
  }
};

}  // namespace ctk
