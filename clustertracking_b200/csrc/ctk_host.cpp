// ctk_host.cpp -- host-only helpers of libctk (no CUDA).
#include <stdint.h>

#include <vector>

#include "ctk.h"

// Union of close pairs with the reference's rule (find.py:41-48): when the clusters of a and b
// merge, the label of a's cluster survives and every member of b's cluster is relabelled.
extern "C" int ctk_label_clusters(const int64_t* pairs, int64_t n_pairs, int64_t n,
                                  int64_t* labels_out, int64_t* sizes_out) {
  if (n < 0 || n_pairs < 0 || (n_pairs > 0 && !pairs) || (n > 0 && (!labels_out || !sizes_out)))
    return CTK_E_INVALID;
  std::vector<int64_t> next(n, -1), tail(n), count(n, 1);
  for (int64_t i = 0; i < n; ++i) { labels_out[i] = i; tail[i] = i; }
  for (int64_t k = 0; k < n_pairs; ++k) {
    const int64_t a = pairs[2 * k], b = pairs[2 * k + 1];
    if (a < 0 || a >= n || b < 0 || b >= n) return CTK_E_INVALID;
    const int64_t keep = labels_out[a], drop = labels_out[b];
    if (keep == drop) continue;
    for (int64_t m = drop; m >= 0; m = next[m]) labels_out[m] = keep;   // drop's chain starts at drop
    next[tail[keep]] = drop;
    tail[keep] = tail[drop];
    count[keep] += count[drop];
  }
  for (int64_t i = 0; i < n; ++i) sizes_out[i] = count[labels_out[i]];
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Iteration order of the python set that scipy's cKDTree.query_pairs(output_type='set') builds.
//
// The reference visits the close pairs in the iteration order of that set (find.py:87-91) and its
// cluster LABELS depend on the order (membership does not).  scipy fills the set by adding the
// tuples (i, j) in the order of its result vector -- the order query_pairs(output_type='ndarray')
// returns -- so the iteration order is a pure function of that sequence and of CPython's set
// implementation (Objects/setobject.c: open addressing, LINEAR_PROBES 9, PERTURB_SHIFT 5, growth
// to 4x used when fill*5 >= mask*3) and tuple hash (Objects/tupleobject.c, xxHash-style, 3.8+).
// This function replays exactly that, without creating a single python object.  find.py checks
// it against a real python set once per process and falls back to building the set if the
// interpreter ever disagrees.
//   pairs [n_pairs, 2] int64 (distinct pairs, non-negative entries); order_out [n_pairs] receives
//   the insertion indices in iteration order.
namespace {
inline uint64_t tuple2_hash(uint64_t a, uint64_t b) {
  const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL,
                 P5 = 2870177450012600261ULL;
  uint64_t acc = P5;
  const uint64_t lanes[2] = {a, b};
  for (int k = 0; k < 2; ++k) {
    acc += lanes[k] * P2;
    acc = (acc << 31) | (acc >> 33);
    acc *= P1;
  }
  acc += 2ULL ^ (P5 ^ 3527539ULL);
  if (acc == (uint64_t) -1) acc = 1546275796ULL;
  return acc;
}

struct SetSlot { int64_t item; uint64_t hash; };   // item < 0: unused

inline void insert_clean(std::vector<SetSlot>& table, uint64_t mask, int64_t item, uint64_t hash) {
  uint64_t perturb = hash, i = hash & mask;
  for (;;) {
    if (table[i].item < 0) { table[i] = {item, hash}; return; }
    if (i + 9 <= mask) {
      for (uint64_t j = 1; j <= 9; ++j)
        if (table[i + j].item < 0) { table[i + j] = {item, hash}; return; }
    }
    perturb >>= 5;
    i = (i * 5 + 1 + perturb) & mask;
  }
}
}  // namespace

extern "C" int ctk_pairs_set_order(const int64_t* pairs, int64_t n_pairs, int64_t* order_out) {
  if (n_pairs < 0 || (n_pairs > 0 && (!pairs || !order_out))) return CTK_E_INVALID;
  uint64_t mask = 7;
  std::vector<SetSlot> table(mask + 1, SetSlot{-1, 0});
  uint64_t fill = 0;
  for (int64_t k = 0; k < n_pairs; ++k) {
    if (pairs[2 * k] < 0 || pairs[2 * k + 1] < 0) return CTK_E_INVALID;
    const uint64_t hash = tuple2_hash((uint64_t) pairs[2 * k], (uint64_t) pairs[2 * k + 1]);
    // set_add_entry: probe (the keys are distinct, so an occupied slot is simply skipped)
    uint64_t perturb = hash, i = hash & mask;
    bool placed = false;
    while (!placed) {
      const int probes = (i + 9 <= mask) ? 9 : 0;
      for (int j = 0; j <= probes; ++j) {
        if (table[i + j].item < 0) { table[i + j] = {k, hash}; placed = true; break; }
      }
      if (!placed) { perturb >>= 5; i = (i * 5 + 1 + perturb) & mask; }
    }
    ++fill;
    if (fill * 5 >= mask * 3) {                      // set_table_resize(used > 50000 ? 2x : 4x)
      const uint64_t minused = fill > 50000 ? fill * 2 : fill * 4;
      uint64_t newsize = 8;
      while (newsize <= minused) newsize <<= 1;
      std::vector<SetSlot> bigger(newsize, SetSlot{-1, 0});
      for (const SetSlot& s : table)
        if (s.item >= 0) insert_clean(bigger, newsize - 1, s.item, s.hash);
      table.swap(bigger);
      mask = newsize - 1;
    }
  }
  int64_t out = 0;
  for (const SetSlot& s : table)
    if (s.item >= 0) order_out[out++] = s.item;
  return out == n_pairs ? 0 : CTK_E_INVALID;
}

// ------------------------------------------------------------------------------------------------
// Cluster labelling of a whole video on host threads (find.py:72-129).
//
// The reference obtains the close pairs of a frame from scipy.spatial.cKDTree.query_pairs and its
// label VALUES depend on the order in which that call reports the pairs (see above).  To label a
// video without one python call per frame, the few hundred lines below restate the published
// algorithm of scipy's kd-tree for exactly the call the reference makes -- cKDTree(data) with the
// default leafsize 16, compact nodes and median splits (scipy/spatial/ckdtree/src/build.cxx),
// query_pairs(r=1, p=2, eps=0) (query_pairs.cxx with the rectangle-distance tracker of
// rectangle.h) -- so that the pairs come out in the same order.  scipy is a third-party dependency
// of the reference (unpinned, `setup.py:22`); find.py verifies this restatement against the
// installed scipy once per process and uses scipy itself if they ever disagree.
// ------------------------------------------------------------------------------------------------
#include <algorithm>
#include <atomic>
#include <cmath>
#include <thread>

namespace {

// max of two finite doubles (std::fmax is an out-of-line libm call; the inputs here are never NaN:
// ctk_cluster_pack_columns rejects non-finite positions like scipy does)
static inline double dmax2(double x, double y) { return x > y ? x : y; }

struct KdNode {
  int split_dim;          // -1: leaf
  double split;
  int start, end;
  int less, greater;
  double lo[3], hi[3];    // leaves: tight bounds of the points (exact pruning, see PairQuery)
};

// A point in tree order.  scipy permutes an index array with std::nth_element and an indirect
// comparison; permuting the records themselves with the same comparison results gives the same
// permutation (the algorithm only sees the outcomes of the comparisons) without the indirection.
struct KdPoint { double c[3]; int64_t idx; };

struct KdTree {
  int n, m;
  bool finite;                     // scipy refuses non-finite data; so do the callers of this tree
  std::vector<KdPoint> pts;
  std::vector<KdNode> nodes;
  double mins[3], maxes[3];

  // data [n, m] (or, when `cols` is given, m column arrays read from row `row0` on); every
  // coordinate is divided by scale[k] (numpy's pos / separation)
  void init(const double* data, int n_, int m_, const double* scale,
            const double* const* cols = nullptr, int64_t row0 = 0) {
    n = n_; m = m_;
    finite = true;
    pts.resize(n);
    for (int i = 0; i < n; ++i) {
      KdPoint& p = pts[i];
      p.idx = i;
      for (int k = 0; k < 3; ++k) {
        if (k >= m) { p.c[k] = 0.; continue; }
        const double v = cols ? cols[k][row0 + i] : data[(int64_t) i * m + k];
        p.c[k] = scale ? v / scale[k] : v;
        finite = finite && std::isfinite(p.c[k]);
      }
    }
    nodes.clear();
    nodes.reserve(n / 4 + 8);
    for (int k = 0; k < m; ++k) { mins[k] = maxes[k] = n ? pts[0].c[k] : 0.; }
    for (int i = 1; i < n; ++i)
      for (int k = 0; k < m; ++k) {
        const double v = pts[i].c[k];
        maxes[k] = maxes[k] > v ? maxes[k] : v;
        mins[k] = mins[k] < v ? mins[k] : v;
      }
    if (n > 0 && finite) build(0, n);
  }

  int build(int start, int end) {
    const int node_index = (int) nodes.size();
    nodes.push_back(KdNode{-1, 0., start, end, -1, -1, {0., 0., 0.}, {0., 0., 0.}});
    // bounds over the node's points (scipy's compact nodes recompute them for the split decision)
    double mx[3], mn[3];
    for (int k = 0; k < m; ++k) { mx[k] = pts[start].c[k]; mn[k] = pts[start].c[k]; }
    for (int j = start + 1; j < end; ++j)
      for (int k = 0; k < m; ++k) {
        const double v = pts[j].c[k];
        mx[k] = mx[k] > v ? mx[k] : v;
        mn[k] = mn[k] < v ? mn[k] : v;
      }
    for (int k = 0; k < m; ++k) { nodes[node_index].lo[k] = mn[k]; nodes[node_index].hi[k] = mx[k]; }
    if (end - start <= 16) return node_index;
    int d = 0;
    double size = 0.;
    for (int k = 0; k < m; ++k)
      if (mx[k] - mn[k] > size) { d = k; size = mx[k] - mn[k]; }
    if (mx[d] == mn[d]) return node_index;            // all points identical: leaf
    // median split
    const int half = (end - start) / 2;
    std::nth_element(pts.begin() + start, pts.begin() + start + half, pts.begin() + end,
                     [d](const KdPoint& a, const KdPoint& b) { return a.c[d] < b.c[d]; });
    double split = pts[start + half].c[d];
    // a median equal to the node's minimum would leave the lower side empty: scipy then splits
    // just above it, so that every point equal to the minimum goes to the lower side
    if (split == mn[d]) split = std::nextafter(split, HUGE_VAL);
    int p = start, q = end - 1;
    while (p <= q) {
      if (pts[p].c[d] < split) ++p;
      else if (pts[q].c[d] >= split) --q;
      else { std::swap(pts[p], pts[q]); ++p; --q; }
    }
    const int l = build(start, p);
    const int g = build(p, end);
    KdNode& nd = nodes[node_index];
    nd.split_dim = d;
    nd.split = split;
    nd.less = l;
    nd.greater = g;
    return node_index;
  }
};

// rectangle-rectangle distance tracker for p = 2 (squared distances), eps = 0
struct RectTracker {
  int m;
  double r1mn[3], r1mx[3], r2mn[3], r2mx[3];
  double min_d, max_d, upper, limit;
  struct Item { int which, dim; double min_d, max_d, mn, mx; };
  Item stack[256];                 // two pushes per level of two trees of depth <= ~40
  int depth;

  void interval(int k, double* mn, double* mx) const {
    *mn = dmax2(0., dmax2(r1mn[k] - r2mx[k], r2mn[k] - r1mx[k]));
    *mx = dmax2(r1mx[k] - r2mn[k], r2mx[k] - r1mn[k]);
  }
  void rect_rect(double* mn, double* mx) const {
    *mn = 0.; *mx = 0.;
    for (int k = 0; k < m; ++k) {
      double a, b;
      interval(k, &a, &b);
      *mn += a * a;
      *mx += b * b;
    }
  }
  void init(const KdTree& t, double r) {
    m = t.m;
    for (int k = 0; k < m; ++k) {
      r1mn[k] = r2mn[k] = t.mins[k];
      r1mx[k] = r2mx[k] = t.maxes[k];
    }
    upper = r * r;
    rect_rect(&min_d, &max_d);
    limit = max_d;
    depth = 0;
  }
  void push(int which, bool less, int dim, double split) {
    double* mn = which == 1 ? r1mn : r2mn;
    double* mx = which == 1 ? r1mx : r2mx;
    stack[depth++] = Item{which, dim, min_d, max_d, mn[dim], mx[dim]};
    double min1, max1, min2, max2;
    interval(dim, &min1, &max1);
    min1 *= min1; max1 *= max1;
    if (less) mx[dim] = split; else mn[dim] = split;
    interval(dim, &min2, &max2);
    min2 *= min2; max2 *= max2;
    bool sub = (min1 != 0 && min1 < limit) || max1 < limit;
    sub = sub || (min2 != 0 && min2 < limit) || max2 < limit;
    sub = sub || min_d < limit || max_d < limit;
    if (sub) rect_rect(&min_d, &max_d);
    else { min_d += (min2 - min1); max_d += (max2 - max1); }
  }
  void pop() {
    const Item it = stack[--depth];
    min_d = it.min_d; max_d = it.max_d;
    if (it.which == 1) { r1mn[it.dim] = it.mn; r1mx[it.dim] = it.mx; }
    else { r2mn[it.dim] = it.mn; r2mx[it.dim] = it.mx; }
  }
};

struct PairQuery {
  const KdTree* t;
  RectTracker tr;
  std::vector<int64_t>* out;      // flat (i, j), i < j

  void add(int64_t i, int64_t j) {
    if (i > j) std::swap(i, j);
    out->push_back(i);
    out->push_back(j);
  }
  void no_checking(int n1, int n2) {
    const KdNode& a = t->nodes[n1];
    const KdNode& b = t->nodes[n2];
    if (a.split_dim == -1) {
      if (b.split_dim == -1) {
        for (int i = a.start; i < a.end; ++i) {
          const int min_j = n1 == n2 ? i + 1 : b.start;
          for (int j = min_j; j < b.end; ++j) add(t->pts[i].idx, t->pts[j].idx);
        }
      } else {
        no_checking(n1, b.less);
        no_checking(n1, b.greater);
      }
    } else if (n1 == n2) {
      no_checking(a.less, b.less);
      no_checking(a.less, b.greater);
      no_checking(a.greater, b.greater);
    } else {
      no_checking(a.less, n2);
      no_checking(a.greater, n2);
    }
  }
  // scipy tests every point pair of two leaves; pairs that cannot be close are skipped here with
  // bounds that are exact lower bounds of the same floating-point sum (monotone rounding), so the
  // pairs reported and their order do not change
  void leaves(int n1, int n2) {
    const KdNode& a = t->nodes[n1];
    const KdNode& b = t->nodes[n2];
    const int m = t->m;
    const double upper = tr.upper;
    if (n1 != n2) {
      double s = 0.;
      for (int k = 0; k < m; ++k) {
        const double g = dmax2(0., dmax2(a.lo[k] - b.hi[k], b.lo[k] - a.hi[k]));
        s += g * g;
      }
      if (s > upper) return;
    }
    const KdPoint* pts = t->pts.data();
    for (int i = a.start; i < a.end; ++i) {
      const KdPoint& u = pts[i];
      if (n1 != n2) {
        double s = 0.;
        for (int k = 0; k < m; ++k) {
          const double g = dmax2(0., dmax2(u.c[k] - b.hi[k], b.lo[k] - u.c[k]));
          s += g * g;
        }
        if (s > upper) continue;
      }
      const int min_j = n1 == n2 ? i + 1 : b.start;
      for (int j = min_j; j < b.end; ++j) {
        const KdPoint& v = pts[j];
        double s = 0.;
        for (int k = 0; k < m; ++k) { const double d = u.c[k] - v.c[k]; s += d * d; }
        if (s <= upper) add(u.idx, v.idx);
      }
    }
  }
  void checking(int n1, int n2) {
    if (tr.min_d > tr.upper) return;
    if (tr.max_d < tr.upper) { no_checking(n1, n2); return; }
    const KdNode& a = t->nodes[n1];
    const KdNode& b = t->nodes[n2];
    if (a.split_dim == -1) {
      if (b.split_dim == -1) {
        leaves(n1, n2);
      } else {
        tr.push(2, true, b.split_dim, b.split);
        checking(n1, b.less);
        tr.pop();
        tr.push(2, false, b.split_dim, b.split);
        checking(n1, b.greater);
        tr.pop();
      }
    } else if (b.split_dim == -1) {
      tr.push(1, true, a.split_dim, a.split);
      checking(a.less, n2);
      tr.pop();
      tr.push(1, false, a.split_dim, a.split);
      checking(a.greater, n2);
      tr.pop();
    } else {
      tr.push(1, true, a.split_dim, a.split);
      tr.push(2, true, b.split_dim, b.split);
      checking(a.less, b.less);
      tr.pop();
      tr.push(2, false, b.split_dim, b.split);
      checking(a.less, b.greater);
      tr.pop();
      tr.pop();
      tr.push(1, false, a.split_dim, a.split);
      if (n1 != n2) {
        tr.push(2, true, b.split_dim, b.split);
        checking(a.greater, b.less);
        tr.pop();
      }
      tr.push(2, false, b.split_dim, b.split);
      checking(a.greater, b.greater);
      tr.pop();
      tr.pop();
    }
  }
};

// per-thread scratch of the frame labelling
struct FrameScratch {
  std::vector<int64_t> pairs, order, ordered, count, sizes;
  KdTree tree;
  PairQuery query;
};

}  // namespace

namespace {
template <class Fn>
void parallel_ranges(int64_t n, int n_threads, Fn&& fn) {
  int nt = n_threads < 1 ? 1 : n_threads;
  if (n < (int64_t) nt * 8192) nt = (int) (n / 8192) + 1;
  if (nt <= 1) { fn((int64_t) 0, n); return; }
  std::vector<std::thread> pool;
  const int64_t step = (n + nt - 1) / nt;
  for (int k = 0; k < nt; ++k) {
    const int64_t a = k * step, b = a + step < n ? a + step : n;
    if (a >= b) break;
    pool.emplace_back([=, &fn]() { fn(a, b); });
  }
  for (auto& th : pool) th.join();
}
}  // namespace

extern "C" int ctk_query_pairs(const double* data, int64_t n, int32_t ndim, int64_t* pairs_out,
                               int64_t capacity, int64_t* n_pairs_out) {
  return ctk_query_pairs_within(data, n, ndim, 1.0, pairs_out, capacity, n_pairs_out);
}

extern "C" int ctk_query_pairs_within(const double* data, int64_t n, int32_t ndim, double r,
                                      int64_t* pairs_out, int64_t capacity, int64_t* n_pairs_out) {
  if (n < 0 || ndim < 1 || ndim > 3 || (n > 0 && !data) || !n_pairs_out || !(r >= 0.)) return CTK_E_INVALID;
  KdTree tree;
  tree.init(data, (int) n, ndim, nullptr);
  if (!tree.finite) return CTK_E_NONFINITE;
  std::vector<int64_t> pairs;
  PairQuery q;
  q.t = &tree;
  q.out = &pairs;
  q.tr.init(tree, r);
  if (n > 0) q.checking(0, 0);
  *n_pairs_out = (int64_t) pairs.size() / 2;
  if (pairs_out) {
    if ((int64_t) pairs.size() / 2 > capacity) return CTK_E_CAPACITY;
    std::copy(pairs.begin(), pairs.end(), pairs_out);
  }
  return 0;
}

extern "C" int ctk_cluster_frames(const double* pos, int64_t n, int32_t ndim, const int64_t* starts,
                                  const int64_t* stops, int64_t n_frames, const double* separation,
                                  int32_t n_threads, int64_t* cluster_out, int64_t* size_out,
                                  int64_t* by_cluster_out, int64_t* span_out) {
  return ctk_cluster_pack_frames(pos, n, ndim, starts, stops, n_frames, separation, n_threads,
                                 cluster_out, size_out, by_cluster_out, span_out, nullptr, nullptr,
                                 0, 0, nullptr, nullptr, nullptr);
}

// ctk_cluster_frames that also, while a frame's rows are hot in the worker's cache, (a) gathers the
// packed parameter rows of the frame in (cluster, row) order (refine.py:345) and (b) writes the
// frame's group table: group_count_out[f] clusters, the row (index into this call's rows) at which
// each starts in group_start_out[starts[f] ...].
//   columns [n_cols] table-order column arrays (NULL entry: the constant scalars[j]); the table row
//   of this call's row i is row_base + i; params_out [n, n_cols] packed rows.  columns == NULL
//   skips the gather; group_*_out == NULL skips the group table.
extern "C" int ctk_cluster_pack_frames(const double* pos, int64_t n, int32_t ndim,
                                       const int64_t* starts, const int64_t* stops, int64_t n_frames,
                                       const double* separation, int32_t n_threads,
                                       int64_t* cluster_out, int64_t* size_out,
                                       int64_t* by_cluster_out, int64_t* span_out,
                                       const double* const* columns, const double* scalars,
                                       int32_t n_cols, int64_t row_base, double* params_out,
                                       int32_t* group_count_out, int32_t* group_start_out) {
  return ctk_cluster_pack_columns(pos, nullptr, n, ndim, starts, stops, n_frames, separation,
                                  n_threads, cluster_out, size_out, by_cluster_out, span_out, columns,
                                  scalars, n_cols, row_base, params_out, group_count_out,
                                  group_start_out);
}

// The same with the positions given as `ndim` table-order column arrays (pos_cols[k][row_base + i]
// is coordinate k of this call's row i) instead of one packed [n, ndim] array.
//
// ctk_cluster_pack_labelled: the same, with the labels of the frames already computed on the
// device (ctk_label_frames): labels_in [n] int32 (this call's rows), frame_flags [n_frames]; a frame
// whose flag is not 0 (scratch capacity exceeded on the device) is labelled here instead.
static int cluster_pack_impl(const double* pos, const double* const* pos_cols, int64_t n,
                             int32_t ndim, const int64_t* starts, const int64_t* stops,
                             int64_t n_frames, const double* separation,
                             int32_t n_threads, int64_t* cluster_out, int64_t* size_out,
                             int64_t* by_cluster_out, int64_t* span_out,
                             const double* const* columns, const double* scalars,
                             int32_t n_cols, int64_t row_base, double* params_out,
                             int32_t* group_count_out, int32_t* group_start_out,
                             const int32_t* labels_in, const int32_t* frame_flags) {
  if (n < 0 || n_frames < 0 || ndim < 1 || ndim > 3 || !separation) return CTK_E_INVALID;
  if (columns && (n_cols < 1 || !scalars || !params_out)) return CTK_E_INVALID;
  if ((group_count_out == nullptr) != (group_start_out == nullptr)) return CTK_E_INVALID;
  if (n_frames == 0) return 0;
  if ((!pos && !pos_cols) || !starts || !stops || !cluster_out || !size_out || !by_cluster_out ||
      !span_out)
    return CTK_E_INVALID;
  for (int64_t f = 0; f < n_frames; ++f)
    if (starts[f] < 0 || stops[f] < starts[f] || stops[f] > n || stops[f] - starts[f] > (1 << 30))
      return CTK_E_INVALID;
  std::atomic<int64_t> next(0);
  std::atomic<int> failed(0);
  auto worker = [&]() {
    FrameScratch s;
    for (;;) {
      const int64_t f = next.fetch_add(1);
      if (f >= n_frames) break;
      const int64_t a = starts[f], b = stops[f];
      const int cnt = (int) (b - a);
      if (cnt == 0) {
        span_out[f] = 0;
        if (group_count_out) group_count_out[f] = 0;
        continue;
      }
      const bool given = labels_in && frame_flags[f] == 0;
      if (!given) {
        if (labels_in && frame_flags[f] == 2) { failed = 2; continue; }
        s.tree.init(pos ? pos + a * ndim : nullptr, cnt, ndim, separation, pos_cols, row_base + a);
        if (!s.tree.finite) { failed = 2; continue; }
        s.pairs.clear();
        s.query.t = &s.tree;
        s.query.out = &s.pairs;
        s.query.tr.init(s.tree, 1.0);
        s.query.checking(0, 0);
        const int64_t np = (int64_t) s.pairs.size() / 2;
        s.order.resize(np);
        s.ordered.resize(2 * np);
        if (ctk_pairs_set_order(s.pairs.data(), np, s.order.data()) != 0) { failed = 1; continue; }
        for (int64_t k = 0; k < np; ++k) {
          s.ordered[2 * k] = s.pairs[2 * s.order[k]];
          s.ordered[2 * k + 1] = s.pairs[2 * s.order[k] + 1];
        }
        if (ctk_label_clusters(s.ordered.data(), np, cnt, cluster_out + a, size_out + a) != 0) {
          failed = 1;
          continue;
        }
      }
      // Four passes over the frame (labels are point indices of the frame, so everything is a
      // counting sort): (1) labels -> histogram; (2) prefix sums: where each label's rows start,
      // its size, the group table; (3) stable scatter of the rows by label, the packed parameter
      // row of every feature written on the way (columns read in table order); (4) sizes.
      s.count.assign((size_t) cnt + 1, 0);
      int64_t* start = s.count.data();               // start[id + 1]: histogram, then offsets
      int64_t top = 0;
      bool ok = true;
      for (int i = 0; i < cnt; ++i) {
        const int64_t id = given ? (int64_t) labels_in[a + i] : cluster_out[a + i];
        if (id < 0 || id >= cnt) { ok = false; break; }
        if (given) cluster_out[a + i] = id;
        ++start[id + 1];
        top = id > top ? id : top;
      }
      if (!ok) { failed = 1; continue; }
      span_out[f] = top + 1;
      s.sizes.resize((size_t) cnt);
      int groups = 0;
      for (int id = 0; id < cnt; ++id) {
        const int64_t size = start[id + 1];
        s.sizes[id] = size;
        if (size > 0 && group_count_out) group_start_out[a + groups++] = (int32_t) (a + start[id]);
        start[id + 1] += start[id];
      }
      if (group_count_out) group_count_out[f] = groups;
      for (int i = 0; i < cnt; ++i) {
        const int64_t id = cluster_out[a + i];
        const int64_t k = a + start[id]++;
        by_cluster_out[k] = a + i;
        size_out[a + i] = s.sizes[id];
        if (columns) {                               // packed rows in (cluster, row) order
          const int64_t row = row_base + a + i;
          double* dst = params_out + k * n_cols;
          for (int j = 0; j < n_cols; ++j) dst[j] = columns[j] ? columns[j][row] : scalars[j];
        }
      }
    }
  };
  int nt = n_threads < 1 ? 1 : n_threads;
  if (nt > n_frames) nt = (int) n_frames;
  if (nt == 1) {
    worker();
  } else {
    std::vector<std::thread> pool;
    for (int k = 0; k < nt; ++k) pool.emplace_back(worker);
    for (auto& th : pool) th.join();
  }
  return failed == 2 ? CTK_E_NONFINITE : (failed ? CTK_E_INVALID : 0);
}

extern "C" int ctk_cluster_pack_columns(const double* pos, const double* const* pos_cols, int64_t n,
                                        int32_t ndim, const int64_t* starts, const int64_t* stops,
                                        int64_t n_frames, const double* separation,
                                        int32_t n_threads, int64_t* cluster_out, int64_t* size_out,
                                        int64_t* by_cluster_out, int64_t* span_out,
                                        const double* const* columns, const double* scalars,
                                        int32_t n_cols, int64_t row_base, double* params_out,
                                        int32_t* group_count_out, int32_t* group_start_out) {
  return cluster_pack_impl(pos, pos_cols, n, ndim, starts, stops, n_frames, separation, n_threads,
                           cluster_out, size_out, by_cluster_out, span_out, columns, scalars, n_cols,
                           row_base, params_out, group_count_out, group_start_out, nullptr, nullptr);
}

extern "C" int ctk_cluster_pack_labelled(const double* pos, const double* const* pos_cols, int64_t n,
                                         int32_t ndim, const int64_t* starts, const int64_t* stops,
                                         int64_t n_frames, const double* separation,
                                         int32_t n_threads, int64_t* cluster_out, int64_t* size_out,
                                         int64_t* by_cluster_out, int64_t* span_out,
                                         const double* const* columns, const double* scalars,
                                         int32_t n_cols, int64_t row_base, double* params_out,
                                         int32_t* group_count_out, int32_t* group_start_out,
                                         const int32_t* labels_in, const int32_t* frame_flags) {
  if ((labels_in == nullptr) != (frame_flags == nullptr)) return CTK_E_INVALID;
  return cluster_pack_impl(pos, pos_cols, n, ndim, starts, stops, n_frames, separation, n_threads,
                           cluster_out, size_out, by_cluster_out, span_out, columns, scalars, n_cols,
                           row_base, params_out, group_count_out, group_start_out, labels_in,
                           frame_flags);
}

// ------------------------------------------------------------------------------------------------
// Packing helpers of the host pipeline (refine.py:336-345, 426-427): plain loops on host threads
// in place of numpy fancy indexing on one core.
// ------------------------------------------------------------------------------------------------

// Running cluster ids, group order and group table of one labelled chunk of frames.
//   local_labels, by_cluster [m]  outputs of ctk_cluster_frames for the chunk (chunk-local rows)
//   starts, stops [n_frames]       chunk-local row range of each frame; spans [n_frames]
//   next_id                        running offset of find.py:127-128 before this chunk
//   row_base                       table row of the chunk's first row; frame_base: index of its first frame
//   cluster_out [m] labels with the running offset; order_out [m] table rows by (frame, cluster);
//   group_offset_out [m + 1] int32; group_frame_out [m] int32; *n_groups_out; *next_id_out
extern "C" int ctk_group_chunk(const int64_t* local_labels, const int64_t* by_cluster,
                               const int64_t* starts, const int64_t* stops, const int64_t* spans,
                               int64_t n_frames, int64_t next_id, int64_t row_base,
                               int32_t frame_base, int64_t* cluster_out, int64_t* order_out,
                               int32_t* group_offset_out, int32_t* group_frame_out,
                               int64_t* n_groups_out, int64_t* next_id_out) {
  if (n_frames < 0 || !n_groups_out || !next_id_out) return CTK_E_INVALID;
  int64_t groups = 0;
  for (int64_t f = 0; f < n_frames; ++f) {
    const int64_t a = starts[f], b = stops[f];
    for (int64_t i = a; i < b; ++i) cluster_out[i] = local_labels[i] + next_id;
    int64_t prev = -1;
    for (int64_t i = a; i < b; ++i) {
      const int64_t row = by_cluster[i];
      order_out[i] = row + row_base;
      const int64_t label = local_labels[row];
      if (label != prev) {
        group_offset_out[groups] = (int32_t) i;
        group_frame_out[groups] = frame_base + (int32_t) f;
        ++groups;
        prev = label;
      }
    }
    next_id += spans[f];
  }
  group_offset_out[groups] = n_frames ? (int32_t) stops[n_frames - 1] : 0;
  *n_groups_out = groups;
  *next_id_out = next_id;
  return 0;
}

// Group table of a chunk from the per-frame tables of ctk_cluster_pack_frames.
//   group_offset_out [total groups + 1] int32, group_frame_out [total groups] int32
extern "C" int ctk_concat_groups(const int64_t* starts, const int64_t* stops,
                                 const int32_t* group_count, const int32_t* group_start,
                                 int64_t n_frames, int32_t frame_base, int32_t* group_offset_out,
                                 int32_t* group_frame_out, int64_t* n_groups_out) {
  if (n_frames < 0 || !n_groups_out) return CTK_E_INVALID;
  int64_t g = 0;
  for (int64_t f = 0; f < n_frames; ++f) {
    const int32_t* src = group_start + starts[f];
    for (int32_t k = 0; k < group_count[f]; ++k) {
      group_offset_out[g] = src[k];
      group_frame_out[g] = frame_base + (int32_t) f;
      ++g;
    }
  }
  group_offset_out[g] = n_frames ? (int32_t) stops[n_frames - 1] : 0;
  *n_groups_out = g;
  return 0;
}

// cluster_out[i] = local[i] + offset of row i's frame (find.py:127-128), on host threads
extern "C" int ctk_apply_label_offsets(const int64_t* local, const int64_t* starts,
                                       const int64_t* stops, const int64_t* frame_offset,
                                       int64_t n_frames, int32_t n_threads, int64_t* cluster_out) {
  if (n_frames < 0 || (n_frames > 0 && (!local || !starts || !stops || !frame_offset || !cluster_out)))
    return CTK_E_INVALID;
  parallel_ranges(n_frames, n_threads, [=](int64_t fa, int64_t fb) {
    for (int64_t f = fa; f < fb; ++f)
      for (int64_t i = starts[f]; i < stops[f]; ++i) cluster_out[i] = local[i] + frame_offset[f];
  });
  return 0;
}

// out[r, j] = columns[j] ? columns[j][rows[r]] : scalars[j]
extern "C" int ctk_gather_rows(const double* const* columns, const double* scalars,
                               const int64_t* rows, int64_t n, int32_t n_cols, double* out,
                               int32_t n_threads) {
  if (n < 0 || n_cols < 1 || !columns || !scalars || (n > 0 && (!rows || !out))) return CTK_E_INVALID;
  parallel_ranges(n, n_threads, [=](int64_t a, int64_t b) {
    for (int64_t r = a; r < b; ++r) {
      const int64_t row = rows[r];
      double* dst = out + r * n_cols;
      for (int j = 0; j < n_cols; ++j) dst[j] = columns[j] ? columns[j][row] : scalars[j];
    }
  });
  return 0;
}

// Write-back of one chunk (refine.py:408-427): columns[j][rows[r]] = params[r, j] for clusters with
// status 0, the untouched input (params_in) for failed ones; cost_out[rows[r]] = the cluster's cost
// or NaN.  Returns the number of failed clusters in *n_failed_out.
extern "C" int ctk_scatter_rows(const double* params, const double* params_in, const int64_t* rows,
                                int64_t row_base, int64_t n, int32_t n_cols, const int32_t* group_offset,
                                const double* group_cost, const int32_t* group_status,
                                int64_t n_groups, double* const* columns, double* cost_out,
                                int32_t n_threads, int64_t* n_failed_out) {
  if (n < 0 || n_cols < 1 || n_groups < 0 || !columns || !cost_out || !n_failed_out)
    return CTK_E_INVALID;
  std::atomic<int64_t> failed(0);
  parallel_ranges(n_groups, n_threads, [=, &failed](int64_t a, int64_t b) {
    int64_t bad = 0;
    for (int64_t g = a; g < b; ++g) {
      const bool ok = group_status[g] == 0;
      const double cost = ok ? group_cost[g] : NAN;
      const double* src = ok ? params : params_in;
      bad += ok ? 0 : 1;
      for (int64_t r = group_offset[g]; r < group_offset[g + 1]; ++r) {
        const int64_t row = row_base + rows[r];
        for (int j = 0; j < n_cols; ++j) columns[j][row] = src[r * n_cols + j];
        cost_out[row] = cost;
      }
    }
    failed += bad;
  });
  *n_failed_out = failed.load();
  return 0;
}

// Launch schedule of one plan: clusters sorted by (size class, size descending) with a counting
// sort.  class_target[k] = class a cluster of class k runs in (k itself, a larger class when class
// k's arrays do not fit shared memory) or -1 (cannot run).
//   caps [n_caps] ascending capacities; work_ids_out [n_clusters]; class_count_out [n_caps] clusters
//   scheduled per TARGET class (work ids are grouped by target class, ascending);
//   not_run_out [n_clusters] ids that cannot run (count in *n_not_run_out)
extern "C" int ctk_schedule(const int32_t* cluster_offset, int64_t n_clusters, const int32_t* caps,
                            int32_t n_caps, const int32_t* class_target, int32_t* work_ids_out,
                            int64_t* class_count_out, int32_t* not_run_out, int64_t* n_not_run_out) {
  if (n_clusters < 0 || n_caps < 1 || n_caps > 64 || !cluster_offset || !caps || !class_target ||
      !work_ids_out || !class_count_out || !not_run_out || !n_not_run_out)
    return CTK_E_INVALID;
  const int B = 64;                                   // size buckets inside a class
  std::vector<int64_t> bucket((size_t) n_caps * B + 1, 0);
  std::vector<int32_t> key((size_t) n_clusters);
  int64_t not_run = 0;
  for (int64_t c = 0; c < n_clusters; ++c) {
    const int size = cluster_offset[c + 1] - cluster_offset[c];
    int k = 0;
    while (k < n_caps && caps[k] < size) ++k;
    const int target = k < n_caps ? class_target[k] : -1;
    if (target < 0) {
      key[c] = -1;
      not_run_out[not_run++] = (int32_t) c;
      continue;
    }
    key[c] = target * B + (B - 1 - (size < B - 1 ? size : B - 1));
    ++bucket[key[c] + 1];
  }
  for (size_t b = 1; b < bucket.size(); ++b) bucket[b] += bucket[b - 1];
  for (int k = 0; k < n_caps; ++k) class_count_out[k] = bucket[(size_t) (k + 1) * B] - bucket[(size_t) k * B];
  for (int64_t c = 0; c < n_clusters; ++c)
    if (key[c] >= 0) work_ids_out[bucket[key[c]]++] = (int32_t) c;
  *n_not_run_out = not_run;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// drop_close for a batch of frames on host threads (find.py:166-207): of every pair of maxima
// closer than `separation` (scaled distance <= 1 - 1e-7) the dimmer one is dropped, ties by the
// smaller sum of scaled coordinates; all pairs are judged at once (a feature goes if it loses ANY
// pair), so the order of the pairs does not matter.
//   coords [n_frames, capacity, ndim] int32 and values [n_frames, capacity] int32 as written by
//   ctk_find_maxima; counts [n_frames]; keep_out [n_frames, capacity] uint8 (1 = kept);
//   kept_out [n_frames] number kept
// ------------------------------------------------------------------------------------------------
extern "C" int ctk_drop_close_frames(const int32_t* coords, const int32_t* values,
                                     const int32_t* counts, int64_t n_frames, int32_t capacity,
                                     int32_t ndim, const double* separation, int32_t n_threads,
                                     uint8_t* keep_out, int32_t* kept_out) {
  if (n_frames < 0 || capacity < 1 || ndim < 1 || ndim > 3 || !separation) return CTK_E_INVALID;
  if (n_frames == 0) return 0;
  if (!coords || !values || !counts || !keep_out || !kept_out) return CTK_E_INVALID;
  bool no_op = false;
  for (int k = 0; k < ndim; ++k) no_op = no_op || separation[k] == 0.;     // find.py:176-177
  std::atomic<int64_t> next(0);
  auto worker = [&]() {
    std::vector<double> scaled, total;
    std::vector<int64_t> pairs;
    KdTree tree;
    PairQuery query;
    for (;;) {
      const int64_t f = next.fetch_add(1);
      if (f >= n_frames) break;
      const int cnt = counts[f] < capacity ? counts[f] : capacity;
      uint8_t* keep = keep_out + (size_t) f * capacity;
      for (int i = 0; i < cnt; ++i) keep[i] = 1;
      kept_out[f] = cnt;
      if (cnt < 2 || no_op) continue;
      const int32_t* c = coords + (size_t) f * capacity * ndim;
      const int32_t* v = values + (size_t) f * capacity;
      scaled.resize((size_t) cnt * ndim);
      total.resize(cnt);
      for (int i = 0; i < cnt; ++i) {
        double s = 0.;
        for (int k = 0; k < ndim; ++k) {
          const double q = (double) c[(size_t) i * ndim + k] / separation[k];
          scaled[(size_t) i * ndim + k] = q;
          s = k == 0 ? q : s + q;
        }
        total[i] = s;
      }
      tree.init(scaled.data(), cnt, ndim, nullptr);
      pairs.clear();
      query.t = &tree;
      query.out = &pairs;
      query.tr.init(tree, 1 - 1e-7);
      query.checking(0, 0);
      for (size_t k = 0; k + 1 < pairs.size(); k += 2) {
        const int64_t i0 = pairs[k], i1 = pairs[k + 1];           // i0 < i1
        int64_t drop;
        if (v[i0] != v[i1]) drop = v[i0] > v[i1] ? i1 : i0;
        else drop = total[i0] > total[i1] ? i1 : i0;
        keep[drop] = 0;
      }
      int kept = 0;
      for (int i = 0; i < cnt; ++i) kept += keep[i];
      kept_out[f] = kept;
    }
  };
  int nt = n_threads < 1 ? 1 : n_threads;
  if (nt > n_frames) nt = (int) n_frames;
  if (nt == 1) {
    worker();
  } else {
    std::vector<std::thread> pool;
    for (int k = 0; k < nt; ++k) pool.emplace_back(worker);
    for (auto& th : pool) th.join();
  }
  return 0;
}

// Wait (sleeping, without the interpreter lock: the caller is a ctypes call) until none of
// flags[0 .. n-1] is negative -- the per-frame flags ctk_label_frames writes into mapped host memory
// -- or until timeout_us have passed.  Returns 0 when all are set, 1 on timeout.
#include <chrono>
extern "C" int ctk_wait_flags(const int32_t* flags, int64_t n, int64_t timeout_us) {
  if (n < 0 || (n > 0 && !flags)) return CTK_E_INVALID;
  const volatile int32_t* f = flags;
  const auto t0 = std::chrono::steady_clock::now();
  int64_t done = 0;
  for (;;) {
    while (done < n && f[done] >= 0) ++done;
    if (done >= n) return 0;
    std::this_thread::sleep_for(std::chrono::microseconds(50));
    if (std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() > timeout_us)
      return 1;
  }
}

// Rows at which a new frame starts in a frame column (find.py:122: groupby(frame)), in one pass.
//   starts_out [capacity] receives the first row of every run of equal values;
//   *n_runs_out the number of runs (may exceed capacity: then only the count is valid);
//   *sorted_out 1 when every change is an increase (the table is sorted by frame)
extern "C" int ctk_frame_runs(const int64_t* frames, int64_t n, int64_t* starts_out, int64_t capacity,
                              int64_t* n_runs_out, int32_t* sorted_out) {
  if (n < 0 || (n > 0 && !frames) || !n_runs_out || !sorted_out || capacity < 0 ||
      (capacity > 0 && !starts_out))
    return CTK_E_INVALID;
  int64_t runs = 0;
  int32_t sorted = 1;
  if (n > 0) {
    if (capacity > 0) starts_out[0] = 0;
    runs = 1;
    for (int64_t i = 1; i < n; ++i) {
      if (frames[i] != frames[i - 1]) {
        if (frames[i] < frames[i - 1]) sorted = 0;
        if (runs < capacity) starts_out[runs] = i;
        ++runs;
      }
    }
  }
  *n_runs_out = runs;
  *sorted_out = sorted;
  return 0;
}
