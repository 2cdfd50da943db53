// ctk_label.cuh -- cluster labelling of one frame by one warp (find.py:72-93), on the device.
//
// The reference's label VALUES depend on the order in which scipy.spatial.cKDTree.query_pairs
// reports the close pairs and on CPython's set iteration order (find.py:87-91; see ctk_host.cpp,
// which holds the verified host restatement).  This file is the same chain for the device:
//
//   build        scipy's kd-tree for cKDTree(pos / separation): leafsize 16, compact nodes, median
//                split through libstdc++'s std::nth_element (introselect: median-of-three,
//                unguarded Hoare partition, heap select at the depth limit, insertion sort of <= 3)
//                restated step by step so that the points end up in the same order inside every leaf;
//   query        query_pairs(1): the dual-tree traversal with the rectangle-distance tracker, as a
//                state machine with explicit stacks; leaf x leaf blocks are tested by all lanes and
//                written in (i, j) order with a ballot;
//   set order    CPython's set of the (i, j) tuples (open addressing, 9 linear probes, perturb
//                shift 5, growth to 4x / 2x used at 3/5 load; xxHash-style tuple hash);
//   union        the reference's rule (find.py:41-48): the label of a's cluster survives.
//
// Frames are independent: ONE WARP PER FRAME, and inside a frame the work is spread over the lanes
// wherever the RESULT does not depend on the order of execution:
//   * the tree is built level by level, one lane per node (each lane runs the sequential selection
//     on its own node; the root, alone on its level, shares its scans between the lanes);
//   * the dual-tree traversal is expanded breadth first, one lane per (node, node) visit; every
//     visit carries its own rectangle-distance state, children are compacted in order, and leaf
//     blocks stay in the list, so the final list is the depth-first order of scipy's recursion;
//   * leaf x leaf blocks are tested by all lanes and written in (i, j) order with a ballot;
//   * hashes of the pairs are computed by all lanes; the inserts into the set table and the union
//     rule are sequential (lane 0) on shared memory.
// Every step is exact: float64 without FMA contraction, the same comparisons as the host
// restatement.  A frame that exceeds a capacity is FLAGGED (frame_flag != 0) and labelled by the
// host path instead.
//
// The same source compiles, with -DCTK_EMUL, as plain C++ with a one-lane warp (tests only).
#pragma once

#include <math.h>
#include <stdint.h>

#ifdef CTK_EMUL
#define LBL_DEV inline
#define LBL_WARP 1
#define LBL_UNROLL
#else
#define LBL_DEV __device__ __forceinline__
#define LBL_WARP 32
#define LBL_UNROLL _Pragma("unroll")
#endif

namespace ctk_label {

#ifdef CTK_EMUL
LBL_DEV int lane_id() { return 0; }
LBL_DEV void warp_sync() {}
LBL_DEV uint32_t ballot(bool p) { return p ? 1u : 0u; }
LBL_DEV uint32_t lanemask_lt() { return 0u; }
template <class T> LBL_DEV T shfl(T v, int) { return v; }
template <class T> LBL_DEV T shfl_xor(T v, int) { return v; }
LBL_DEV int popc(uint32_t v) { return __builtin_popcount(v); }
LBL_DEV int ctz(uint32_t v) { return __builtin_ctz(v); }
LBL_DEV int clz(uint32_t v) { return __builtin_clz(v); }
LBL_DEV double dsub(double a, double b) { return a - b; }     // built with -ffp-contract=off
LBL_DEV double dmul(double a, double b) { return a * b; }
LBL_DEV double dadd(double a, double b) { return a + b; }
LBL_DEV double ddiv(double a, double b) { return a / b; }
LBL_DEV double next_up(double v) { return nextafter(v, HUGE_VAL); }
LBL_DEV bool is_finite(double v) { return std::isfinite(v); }
#else
LBL_DEV int lane_id() { return threadIdx.x & 31; }
LBL_DEV void warp_sync() { __syncwarp(); }
LBL_DEV uint32_t ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
LBL_DEV uint32_t lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }
template <class T> LBL_DEV T shfl(T v, int s) { return __shfl_sync(0xffffffffu, v, s); }
template <class T> LBL_DEV T shfl_xor(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
LBL_DEV int popc(uint32_t v) { return __popc(v); }
LBL_DEV int ctz(uint32_t v) { return __ffs((int) v) - 1; }
LBL_DEV int clz(uint32_t v) { return __clz((int) v); }
LBL_DEV double dsub(double a, double b) { return __dsub_rn(a, b); }
LBL_DEV double dmul(double a, double b) { return __dmul_rn(a, b); }
LBL_DEV double dadd(double a, double b) { return __dadd_rn(a, b); }
LBL_DEV double ddiv(double a, double b) { return __ddiv_rn(a, b); }
LBL_DEV double next_up(double v) { return nextafter(v, (double) INFINITY); }
LBL_DEV bool is_finite(double v) { return isfinite(v); }
#endif

LBL_DEV double dmax2(double x, double y) { return x > y ? x : y; }

// -DCTK_LABEL_TIMING: cycles per phase, summed over frames (profiles/tools/label_bench.py)
#if defined(CTK_LABEL_TIMING) && !defined(CTK_EMUL)
#define LBL_TICK(slot) do { const long long now_ = clock64(); if (lane == 0 && timing) atomicAdd(timing + (slot), (unsigned long long) (now_ - tick_)); tick_ = now_; } while (0)
#define LBL_TICK_DECL long long tick_ = clock64()
#else
#define LBL_TICK(slot) do { } while (0)
#define LBL_TICK_DECL do { } while (0)
#endif

enum { FLAG_OK = 0, FLAG_CAPACITY = 1, FLAG_NONFINITE = 2 };
enum { LEAF_SIZE = 16 };
enum { MODE_CHECK = 0, MODE_NOCHECK = 1, MODE_LEAVES = 2, MODE_LEAVES_ALL = 3 };

struct Node {                 // 32 bytes
  double split;
  int32_t start, end;
  int32_t less, greater;
  int32_t split_dim;          // -1: leaf
  int32_t pad_;
};
struct Box { double lo[3], hi[3]; };     // tight bounds of a node's points

struct Pair { int32_t i, j; };

// One visit of the dual-tree traversal with the state of scipy's RectRectDistanceTracker at its
// entry (rectangle.h: both rectangles, min / max squared distance).  128 bytes.
struct Task {
  int32_t n1, n2, mode, pad_;
  double min_d, max_d;
  double r1mn[3], r1mx[3], r2mn[3], r2mx[3];
};

// Capacities of one warp's scratch and the scratch itself (pointers into one allocation).
struct Caps {
  int32_t points;     // most points of a frame
  int32_t nodes;      // 2 * points (every split leaves both sides non-empty)
  int32_t tasks;      // visits kept per level of the traversal
  int32_t pairs;      // close pairs kept per frame; more -> FLAG_CAPACITY
  int32_t table;      // slots of each of the two set tables (power of two > 4 * pairs)
};

struct Scratch {
  double* c[3];       // coordinates in tree order, one array per axis
  int32_t* idx;       // original row (frame-local) of the point at each tree position
  int32_t *label, *next, *tail;
  Node* nodes;
  Box* boxes;
  Task* tasks[2];
  Pair* pairs;
  int32_t* table[2];
  uint64_t* stage;     // [32] hashes of a batch of pairs on their way to lane 0
  // shared-memory window (device): [fast, fast + fast_bytes) holds the point arrays and the nodes
  // while the tree is in use, then the set tables and the union arrays
  char* fast;
  int64_t fast_bytes;
  int32_t fast_nodes;  // nodes that fit the window (0: nodes live in the scratch)
};

#ifndef CTK_EMUL
__host__ __device__
#endif
inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

inline Caps make_caps(int64_t max_points, int64_t pair_factor) {
  Caps c;
  c.points = (int32_t) (max_points < 1 ? 1 : max_points);
  c.nodes = 2 * c.points;
  c.tasks = 2 * c.points + 2048;
  int64_t pairs = pair_factor * c.points + 1024;
  if (pairs > (1 << 27)) pairs = 1 << 27;
  c.pairs = (int32_t) pairs;
  int64_t table = 8;
  while (table <= 4 * pairs) table <<= 1;
  c.table = (int32_t) table;
  return c;
}

// byte size of one warp's scratch; carve() follows the same order
inline int64_t scratch_bytes(const Caps& c) {
  int64_t b = 0;
  b += 3 * align_up((int64_t) c.points * 8, 128);
  b += 4 * align_up((int64_t) c.points * 4, 128);
  b += align_up((int64_t) c.nodes * (int64_t) sizeof(Node), 128);
  b += align_up((int64_t) c.nodes * (int64_t) sizeof(Box), 128);
  b += 2 * align_up((int64_t) c.tasks * (int64_t) sizeof(Task), 128);
  b += align_up((int64_t) c.pairs * (int64_t) sizeof(Pair), 128);
  b += 2 * align_up((int64_t) c.table * 4, 128);
  b += 256;
  return b;
}

#ifdef CTK_EMUL
inline
#else
__host__ __device__ inline
#endif
Scratch carve(char* base, const Caps& c) {
  Scratch s;
  char* p = base;
  auto take = [&p](int64_t bytes) { char* q = p; p += (bytes + 127) / 128 * 128; return q; };
  LBL_UNROLL
  for (int k = 0; k < 3; ++k) s.c[k] = reinterpret_cast<double*>(take((int64_t) c.points * 8));
  s.idx = reinterpret_cast<int32_t*>(take((int64_t) c.points * 4));
  s.label = reinterpret_cast<int32_t*>(take((int64_t) c.points * 4));
  s.next = reinterpret_cast<int32_t*>(take((int64_t) c.points * 4));
  s.tail = reinterpret_cast<int32_t*>(take((int64_t) c.points * 4));
  s.nodes = reinterpret_cast<Node*>(take((int64_t) c.nodes * (int64_t) sizeof(Node)));
  s.boxes = reinterpret_cast<Box*>(take((int64_t) c.nodes * (int64_t) sizeof(Box)));
  s.tasks[0] = reinterpret_cast<Task*>(take((int64_t) c.tasks * (int64_t) sizeof(Task)));
  s.tasks[1] = reinterpret_cast<Task*>(take((int64_t) c.tasks * (int64_t) sizeof(Task)));
  s.pairs = reinterpret_cast<Pair*>(take((int64_t) c.pairs * (int64_t) sizeof(Pair)));
  s.table[0] = reinterpret_cast<int32_t*>(take((int64_t) c.table * 4));
  s.table[1] = reinterpret_cast<int32_t*>(take((int64_t) c.table * 4));
  s.stage = reinterpret_cast<uint64_t*>(take(256));
  s.fast = nullptr;
  s.fast_bytes = 0;
  s.fast_nodes = 0;
  return s;
}

// bytes of fast window that hold a frame of n points completely (points + the expected nodes)
#ifndef CTK_EMUL
__host__ __device__
#endif
inline int64_t window_bytes(int64_t n, int m) {
  return align_up(n * (8 * m + 4), 16) + (n / 4 + 64) * (int64_t) sizeof(Node) + 64;
}

// Point the frame's hot arrays into the fast window `w` (shared memory on the device) as far as
// they fit: the point arrays first, then the nodes (expected count; more -> FLAG_CAPACITY).
LBL_DEV void use_window(Scratch& s, char* w, int64_t bytes, int n, int m) {
  s.fast = w;
  s.fast_bytes = bytes / 16 * 16;
  s.fast_nodes = 0;
  const int64_t pts = align_up((int64_t) n * (8 * m + 4), 16);
  if (!w || pts > s.fast_bytes) return;
  char* p = w;
  LBL_UNROLL
  for (int k = 0; k < 3; ++k)
    if (k < m) { s.c[k] = reinterpret_cast<double*>(p); p += (int64_t) n * 8; }
  s.idx = reinterpret_cast<int32_t*>(p);
  const int64_t nodes = (s.fast_bytes - pts) / (int64_t) sizeof(Node);
  if (nodes >= n / 4 + 64) {
    s.nodes = reinterpret_cast<Node*>(w + pts);
    s.fast_nodes = (int32_t) nodes;
  }
}

// ------------------------------------------------------------------------------------------------
// std::nth_element (libstdc++ bits/stl_algo.h, bits/stl_heap.h) on the point records, comparing
// coordinate `key` only.  Runs on ONE lane.
// ------------------------------------------------------------------------------------------------
struct Points {
  double* c[3];
  int32_t* idx;
  double* key;      // = c[d]
  int m;
};

struct Record { double c[3]; int32_t idx; };

LBL_DEV Record load_rec(const Points& p, int i) {
  Record r;
  LBL_UNROLL
  for (int k = 0; k < 3; ++k) r.c[k] = k < p.m ? p.c[k][i] : 0.;
  r.idx = p.idx[i];
  return r;
}
LBL_DEV void store_rec(const Points& p, int i, const Record& r) {
  LBL_UNROLL
  for (int k = 0; k < 3; ++k) if (k < p.m) p.c[k][i] = r.c[k];
  p.idx[i] = r.idx;
}
LBL_DEV void move_rec(const Points& p, int dst, int src) {
  LBL_UNROLL
  for (int k = 0; k < 3; ++k) if (k < p.m) p.c[k][dst] = p.c[k][src];
  p.idx[dst] = p.idx[src];
}
LBL_DEV void swap_rec(const Points& p, int i, int j) {
  LBL_UNROLL
  for (int k = 0; k < 3; ++k) if (k < p.m) { const double t = p.c[k][i]; p.c[k][i] = p.c[k][j]; p.c[k][j] = t; }
  const int32_t t = p.idx[i]; p.idx[i] = p.idx[j]; p.idx[j] = t;
}
LBL_DEV double rec_key(const Points& p, const Record& r) {
  return p.key == p.c[0] ? r.c[0] : (p.key == p.c[1] ? r.c[1] : r.c[2]);
}

// __push_heap / __adjust_heap / __make_heap / __pop_heap / __heap_select on [first, ...)
LBL_DEV void push_heap(const Points& p, int first, int hole, int top, const Record& value) {
  const double vk = rec_key(p, value);
  int parent = (hole - 1) / 2;
  while (hole > top && p.key[first + parent] < vk) {
    move_rec(p, first + hole, first + parent);
    hole = parent;
    parent = (hole - 1) / 2;
  }
  store_rec(p, first + hole, value);
}
LBL_DEV void adjust_heap(const Points& p, int first, int hole, int len, const Record& value) {
  const int top = hole;
  int second = hole;
  while (second < (len - 1) / 2) {
    second = 2 * (second + 1);
    if (p.key[first + second] < p.key[first + (second - 1)]) --second;
    move_rec(p, first + hole, first + second);
    hole = second;
  }
  if ((len & 1) == 0 && second == (len - 2) / 2) {
    second = 2 * (second + 1);
    move_rec(p, first + hole, first + (second - 1));
    hole = second - 1;
  }
  push_heap(p, first, hole, top, value);
}
LBL_DEV void heap_select(const Points& p, int first, int middle, int last) {
  const int len = middle - first;
  if (len >= 2) {                                   // __make_heap
    int parent = (len - 2) / 2;
    for (;;) {
      const Record value = load_rec(p, first + parent);
      adjust_heap(p, first, parent, len, value);
      if (parent == 0) break;
      --parent;
    }
  }
  for (int i = middle; i < last; ++i)
    if (p.key[i] < p.key[first]) {                  // __pop_heap(first, middle, i)
      const Record value = load_rec(p, i);
      move_rec(p, i, first);
      adjust_heap(p, first, 0, len, value);
    }
}
LBL_DEV void insertion_sort(const Points& p, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    const Record val = load_rec(p, i);
    const double vk = rec_key(p, val);
    if (vk < p.key[first]) {
      for (int j = i; j > first; --j) move_rec(p, j, j - 1);       // move_backward
      store_rec(p, first, val);
    } else {                                                       // __unguarded_linear_insert
      int hole = i, next = i - 1;
      while (vk < p.key[next]) { move_rec(p, hole, next); hole = next; --next; }
      store_rec(p, hole, val);
    }
  }
}
LBL_DEV void nth_element(const Points& p, int first, int nth, int last) {
  if (first == last || nth == last) return;
  int depth_limit = 2 * (31 - clz((uint32_t) (last - first)));      // std::__lg(n) * 2
  const double* key = p.key;
  while (last - first > 3) {
    if (depth_limit == 0) {
      heap_select(p, first, nth + 1, last);
      swap_rec(p, first, nth);
      return;
    }
    --depth_limit;
    // __unguarded_partition_pivot: median of (first + 1, mid, last - 1) to first
    const int mid = first + (last - first) / 2;
    {
      const int a = first + 1, b = mid, c = last - 1;
      const double ka = key[a], kb = key[b], kc = key[c];
      int med;
      if (ka < kb) med = (kb < kc) ? b : ((ka < kc) ? c : a);
      else med = (ka < kc) ? a : ((kb < kc) ? c : b);
      swap_rec(p, first, med);
    }
    const double pv = key[first];
    int lo = first + 1, hi = last;
    for (;;) {                                                      // __unguarded_partition
      while (key[lo] < pv) ++lo;
      --hi;
      while (pv < key[hi]) --hi;
      if (!(lo < hi)) break;
      swap_rec(p, lo, hi);
      ++lo;
    }
    if (lo <= nth) first = lo; else last = lo;
  }
  insertion_sort(p, first, last);
}

// ------------------------------------------------------------------------------------------------
// one frame
// ------------------------------------------------------------------------------------------------
struct Rect {                 // the tracker state of one visit, in registers
  double r1mn[3], r1mx[3], r2mn[3], r2mx[3];
  double min_d, max_d;
};

LBL_DEV double get3(const double* v, int k) { return k == 0 ? v[0] : (k == 1 ? v[1] : v[2]); }
LBL_DEV void set3(double* v, int k, double x) { if (k == 0) v[0] = x; else if (k == 1) v[1] = x; else v[2] = x; }

struct FrameLabeller {
  Scratch s;
  Caps caps;
  int m, n, lane;
  int n_nodes;
  int flag;
  double mins[3], maxes[3];
  unsigned long long* timing = nullptr;     // [8] cycle counters (CTK_LABEL_TIMING builds)

  // ---- tree ----------------------------------------------------------------------------------
  // bounds of points [start, end) on all lanes (every lane gets the result)
  LBL_DEV void bounds_warp(int start, int end, double* mn, double* mx) const {
    LBL_UNROLL
    for (int k = 0; k < 3; ++k) { mn[k] = 0.; mx[k] = 0.; }
LBL_UNROLL
    for (int k = 0; k < 3; ++k) {
      if (k >= m) continue;
      const double* c = s.c[k];
      double lo = c[start], hi = lo;
      for (int i = start + lane; i < end; i += LBL_WARP) {
        const double v = c[i];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
      }
      for (int w = LBL_WARP / 2; w > 0; w >>= 1) {
        const double a = shfl_xor(lo, w), b = shfl_xor(hi, w);
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
      }
      mn[k] = lo; mx[k] = hi;
    }
  }
  // the same on one lane
  LBL_DEV void bounds_lane(int start, int end, double* mn, double* mx) const {
    LBL_UNROLL
    for (int k = 0; k < 3; ++k) { mn[k] = 0.; mx[k] = 0.; }
LBL_UNROLL
    for (int k = 0; k < 3; ++k) {
      if (k >= m) continue;
      const double* c = s.c[k];
      double lo = c[start], hi = lo;
      for (int i = start + 1; i < end; ++i) {
        const double v = c[i];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
      }
      mn[k] = lo; mx[k] = hi;
    }
  }

  // first index in [p, q] whose key is not < split (q + 1 if none); all lanes
  LBL_DEV int scan_up(const double* key, int p, int q, double split) const {
    for (int base = p; base <= q; base += LBL_WARP) {
      const int i = base + lane;
      const uint32_t b = ballot(i <= q && !(key[i] < split));
      if (b) return base + ctz(b);
    }
    return q + 1;
  }
  // last index in [p, q] whose key is < split (p - 1 if none); all lanes
  LBL_DEV int scan_down(const double* key, int p, int q, double split) const {
    for (int base = q; base >= p; base -= LBL_WARP) {
      const int i = base - lane;
      const uint32_t b = ballot(i >= p && key[i] < split);
      if (b) return base - ctz(b);
    }
    return p - 1;
  }

  LBL_DEV Points points(int d) const {
    Points pts;
    LBL_UNROLL
    for (int k = 0; k < 3; ++k) pts.c[k] = s.c[k];
    pts.idx = s.idx; pts.key = d == 0 ? s.c[0] : (d == 1 ? s.c[1] : s.c[2]); pts.m = m;
    return pts;
  }

  // split dimension of a node from its bounds (-1: leaf)
  LBL_DEV int split_dim(int cnt, const double* mn, const double* mx) const {
    int d = 0;
    double size = 0.;
    LBL_UNROLL
    for (int k = 0; k < 3; ++k)
      if (k < m && mx[k] - mn[k] > size) { d = k; size = mx[k] - mn[k]; }
    if (cnt <= LEAF_SIZE || get3(mx, d) == get3(mn, d)) return -1;
    return d;
  }

  // Nodes are numbered level by level; the two children of a node are adjacent.
  LBL_DEV void build() {
    // ---- root: alone on its level, so its scans are shared between the lanes
#if defined(CTK_LABEL_TIMING) && !defined(CTK_EMUL)
    const long long build0_ = clock64();
#endif
    double mn[3], mx[3];
    bounds_warp(0, n, mn, mx);
    int level_begin = 0, level_end = 1;
    n_nodes = 1;
    {
      const int d = split_dim(n, mn, mx);
      int p = 0;
      double split = 0.;
      if (d >= 0) {
        const Points pts = points(d);
        if (lane == 0) nth_element(pts, 0, n / 2, n);
        warp_sync();
        const double* key = pts.key;
        split = key[n / 2];
        if (split == get3(mn, d)) split = next_up(split);
        int q = n - 1;
        for (;;) {
          p = scan_up(key, p, q, split);
          if (p > q) break;
          q = scan_down(key, p, q, split);
          if (p > q) break;
          if (lane == 0) swap_rec(pts, p, q);
          warp_sync();
          ++p; --q;
        }
      }
      if (lane == 0) {
        Node nd;
        nd.split = split; nd.start = 0; nd.end = n; nd.split_dim = d; nd.pad_ = 0;
        nd.less = d >= 0 ? 1 : -1; nd.greater = d >= 0 ? 2 : -1;
        s.nodes[0] = nd;
        Box bx;
        LBL_UNROLL
        for (int k = 0; k < 3; ++k) { bx.lo[k] = mn[k]; bx.hi[k] = mx[k]; }
        s.boxes[0] = bx;
        if (d >= 0) {
          Node c;
          c.split = 0.; c.split_dim = -1; c.less = -1; c.greater = -1; c.pad_ = 0;
          c.start = 0; c.end = p; s.nodes[1] = c;
          c.start = p; c.end = n; s.nodes[2] = c;
        }
      }
      if (d >= 0) { n_nodes = 3; level_begin = 1; level_end = 3; } else { level_begin = 1; level_end = 1; }
      warp_sync();
    }
#if defined(CTK_LABEL_TIMING) && !defined(CTK_EMUL)
    if (lane == 0 && timing) atomicAdd(timing + 5, (unsigned long long) (clock64() - build0_));
#endif
    // ---- the other levels: one lane per node
    const int node_cap = s.fast_nodes > 0 ? s.fast_nodes : caps.nodes;
    while (level_begin < level_end) {
      for (int base = level_begin; base < level_end; base += LBL_WARP) {
        const int node = base + lane;
        const bool active = node < level_end;
        int d = -1, p = 0, start = 0, end = 0;
        double split = 0.;
        if (active) {
          start = s.nodes[node].start; end = s.nodes[node].end;
          bounds_lane(start, end, mn, mx);
          Box bx;
          LBL_UNROLL
          for (int k = 0; k < 3; ++k) { bx.lo[k] = mn[k]; bx.hi[k] = mx[k]; }
          s.boxes[node] = bx;
          d = split_dim(end - start, mn, mx);
          if (d >= 0) {
            const Points pts = points(d);
            const int half = (end - start) / 2;
            nth_element(pts, start, start + half, end);
            const double* key = pts.key;
            split = key[start + half];
            // a median equal to the node's minimum would leave the lower side empty: scipy splits
            // just above it, so that every point equal to the minimum goes to the lower side
            if (split == get3(mn, d)) split = next_up(split);
            p = start;
            int q = end - 1;
            while (p <= q) {
              if (key[p] < split) ++p;
              else if (key[q] >= split) --q;
              else { swap_rec(pts, p, q); ++p; --q; }
            }
          }
        }
        const uint32_t bits = ballot(d >= 0);
        const int total = 2 * popc(bits);
        if (n_nodes + total > node_cap) { flag = FLAG_CAPACITY; return; }
        if (d >= 0) {
          const int child = n_nodes + 2 * popc(bits & lanemask_lt());
          s.nodes[node].split = split;
          s.nodes[node].split_dim = d;
          s.nodes[node].less = child;
          s.nodes[node].greater = child + 1;
          Node c;
          c.split = 0.; c.split_dim = -1; c.less = -1; c.greater = -1; c.pad_ = 0;
          c.start = start; c.end = p; s.nodes[child] = c;
          c.start = p; c.end = end; s.nodes[child + 1] = c;
        }
        n_nodes += total;
      }
      warp_sync();
      level_begin = level_end;
      level_end = n_nodes;
    }
  }

  // ---- pair query ------------------------------------------------------------------------------
  LBL_DEV void rect_rect(Rect& t) const {
    double a0 = 0., b0 = 0.;
LBL_UNROLL
    for (int k = 0; k < 3; ++k) {
      if (k < m) {
        const double a = dmax2(0., dmax2(dsub(t.r1mn[k], t.r2mx[k]), dsub(t.r2mn[k], t.r1mx[k])));
        const double b = dmax2(dsub(t.r1mx[k], t.r2mn[k]), dsub(t.r2mx[k], t.r1mn[k]));
        a0 = dadd(a0, dmul(a, a));
        b0 = dadd(b0, dmul(b, b));
      }
    }
    t.min_d = a0; t.max_d = b0;
  }
  LBL_DEV void interval_dim(const Rect& t, int k, double* mn, double* mx) const {
    const double a_mn = get3(t.r1mn, k), a_mx = get3(t.r1mx, k);
    const double b_mn = get3(t.r2mn, k), b_mx = get3(t.r2mx, k);
    *mn = dmax2(0., dmax2(dsub(a_mn, b_mx), dsub(b_mn, a_mx)));
    *mx = dmax2(dsub(a_mx, b_mn), dsub(b_mx, a_mn));
  }
  // RectRectDistanceTracker::push for p = 2 (rectangle.h), on a state held by value
  LBL_DEV void track_push(Rect& t, double limit, int which, bool less, int dim, double split) const {
    double* mn = which == 1 ? t.r1mn : t.r2mn;
    double* mx = which == 1 ? t.r1mx : t.r2mx;
    double min1, max1, min2, max2;
    interval_dim(t, dim, &min1, &max1);
    min1 = dmul(min1, min1); max1 = dmul(max1, max1);
    if (less) set3(mx, dim, split); else set3(mn, dim, split);
    interval_dim(t, dim, &min2, &max2);
    min2 = dmul(min2, min2); max2 = dmul(max2, max2);
    bool sub = (min1 != 0 && min1 < limit) || max1 < limit;
    sub = sub || (min2 != 0 && min2 < limit) || max2 < limit;
    sub = sub || t.min_d < limit || t.max_d < limit;
    if (sub) rect_rect(t);
    else { t.min_d = dadd(t.min_d, dsub(min2, min1)); t.max_d = dadd(t.max_d, dsub(max2, max1)); }
  }

  LBL_DEV static void load_rect(const Task* src, Rect& r) {
    LBL_UNROLL
    for (int k = 0; k < 3; ++k) { r.r1mn[k] = src->r1mn[k]; r.r1mx[k] = src->r1mx[k]; r.r2mn[k] = src->r2mn[k]; r.r2mx[k] = src->r2mx[k]; }
    r.min_d = src->min_d; r.max_d = src->max_d;
  }
  LBL_DEV static void store_task(Task* dst, int n1, int n2, int mode, const Rect& r) {
    dst->n1 = n1; dst->n2 = n2; dst->mode = mode; dst->pad_ = 0;
    dst->min_d = r.min_d; dst->max_d = r.max_d;
    LBL_UNROLL
    for (int k = 0; k < 3; ++k) { dst->r1mn[k] = r.r1mn[k]; dst->r1mx[k] = r.r1mx[k]; dst->r2mn[k] = r.r2mn[k]; dst->r2mx[k] = r.r2mx[k]; }
  }

  // entry tests of a visit (query_pairs.cxx traverse_checking): -1 = pruned, else the mode it runs in
  LBL_DEV int enter(int n1, int n2, int mode, const Rect& r, double upper) const {
    if (mode == MODE_CHECK) {
      if (r.min_d > upper) return -1;
      if (r.max_d < upper) mode = MODE_NOCHECK;
    }
    if (s.nodes[n1].split_dim == -1 && s.nodes[n2].split_dim == -1) {
      if (mode == MODE_NOCHECK) return MODE_LEAVES_ALL;
      if (n1 != n2) {                    // exact lower bound of the same float64 sums: skip the block
        const Box a = s.boxes[n1], b = s.boxes[n2];
        double q = 0.;
        LBL_UNROLL
        for (int k = 0; k < 3; ++k)
          if (k < m) {
            const double g = dmax2(0., dmax2(dsub(a.lo[k], b.hi[k]), dsub(b.lo[k], a.hi[k])));
            q = dadd(q, dmul(g, g));
          }
        if (q > upper) return -1;
      }
      return MODE_LEAVES;
    }
    return mode;
  }

  // All lanes: the pairs of leaf block (n1, n2) in (i, j) order; `checked`: test the distance.
  // The lanes tile the block row-major (w lanes per row of b, 32 / w rows of a per step); the
  // float64 tests of GROUP steps are issued together (they are independent; one test alone is a
  // chain of dependent float64 operations), then the hits are written in order with ballots.
  LBL_DEV void leaf_block(int n1, int n2, bool checked, double upper, int& np) {
    enum { GROUP = 4 };
    const int a0 = s.nodes[n1].start, na = s.nodes[n1].end - a0;
    const int b0 = s.nodes[n2].start, nb = s.nodes[n2].end - b0;
    const bool same = n1 == n2;
    const double *c0 = s.c[0], *c1 = s.c[1], *c2 = s.c[2];
    int w = LBL_WARP;                           // lanes per row
    if (nb <= LBL_WARP / 2) { w = 1; while (w < nb) w <<= 1; }
    const int rows = LBL_WARP / w;
    const int r = lane / w, jj0 = lane % w;
    // several column chunks (nb > 32): a row must finish all its chunks before the next row starts
    const int group = nb > w ? 1 : GROUP;
    for (int i0 = 0; i0 < na; i0 += rows * group) {
      for (int jb = 0; jb < nb; jb += w) {
        const int j = b0 + jb + jj0;
        const bool col_ok = jb + jj0 < nb;
        double v0 = 0., v1 = 0., v2 = 0.;
        int vi = 0;
        if (col_ok) { v0 = c0[j]; if (m > 1) v1 = c1[j]; if (m > 2) v2 = c2[j]; vi = s.idx[j]; }
        bool hit[GROUP];
        int ui[GROUP];
        LBL_UNROLL
        for (int g = 0; g < GROUP; ++g) {
          const int i = a0 + i0 + g * rows + r;
          hit[g] = g < group && col_ok && i0 + g * rows + r < na && (!same || j > i);
          ui[g] = 0;
          if (hit[g]) {
            ui[g] = s.idx[i];
            if (checked) {
              double d = dsub(c0[i], v0);
              double q = dadd(0., dmul(d, d));
              if (m > 1) { d = dsub(c1[i], v1); q = dadd(q, dmul(d, d)); }
              if (m > 2) { d = dsub(c2[i], v2); q = dadd(q, dmul(d, d)); }
              hit[g] = q <= upper;
            }
          }
        }
        LBL_UNROLL
        for (int g = 0; g < GROUP; ++g) {
          if (g >= group) continue;
          const uint32_t bits = ballot(hit[g]);
          if (hit[g]) {
            const int pos = np + popc(bits & lanemask_lt());
            if (pos < caps.pairs) s.pairs[pos] = ui[g] < vi ? Pair{ui[g], vi} : Pair{vi, ui[g]};
          }
          np += popc(bits);
        }
      }
    }
  }

  // query_pairs(r): breadth-first expansion of traverse_checking / traverse_no_checking.  A visit
  // and its children are independent of their siblings (the tracker's pop restores the saved state
  // exactly), so every lane expands one visit; children are written in order and leaf blocks stay
  // in the list, which therefore ends as the leaf blocks in the order of the recursion.
  LBL_DEV int query(double r) {
    const double upper = dmul(r, r);
    Rect root;
    LBL_UNROLL
    for (int k = 0; k < 3; ++k) {
      root.r1mn[k] = root.r2mn[k] = mins[k];
      root.r1mx[k] = root.r2mx[k] = maxes[k];
    }
    rect_rect(root);
    const double limit = root.max_d;
    int cur = 0, n_cur = 0;
    {
      const int mode = enter(0, 0, MODE_CHECK, root, upper);
      if (mode >= 0) {
        if (lane == 0) store_task(s.tasks[0], 0, 0, mode, root);
        n_cur = 1;
      }
      warp_sync();
    }
    for (;;) {
      const Task* from = cur ? s.tasks[1] : s.tasks[0];
      Task* to = cur ? s.tasks[0] : s.tasks[1];
      int n_next = 0;
      bool any_open = false;
      for (int base = 0; base < n_cur; base += LBL_WARP) {
        const int t = base + lane;
        const bool active = t < n_cur;
        int n1 = 0, n2 = 0, mode = MODE_LEAVES;
        if (active) { n1 = from[t].n1; n2 = from[t].n2; mode = from[t].mode; }
        const bool open = active && mode < MODE_LEAVES;
        // children in the order of the recursion: slot = 2 * (side of a) + (side of b)
        int c1[4], c2[4], cm[4];
        double cmin[4], cmax[4];
        for (int j = 0; j < 4; ++j) { c1[j] = 0; c2[j] = 0; cm[j] = -1; cmin[j] = 0.; cmax[j] = 0.; }
        bool desc_a = false, desc_b = false;
        int a_dim = -1, b_dim = -1;
        double a_split = 0., b_split = 0.;
        Rect parent;
        if (open) {
          const Node a = s.nodes[n1];
          const Node b = s.nodes[n2];
          a_dim = a.split_dim; b_dim = b.split_dim; a_split = a.split; b_split = b.split;
          if (mode == MODE_CHECK) { desc_a = a_dim != -1; desc_b = b_dim != -1; }
          else if (a_dim == -1) { desc_b = true; }
          else if (n1 == n2) { desc_a = true; desc_b = true; }
          else { desc_a = true; }
          load_rect(from + t, parent);
LBL_UNROLL
          for (int sa = 0; sa < 2; ++sa) {
            if (!desc_a && sa == 1) continue;
            Rect mid = parent;
            if (desc_a && mode == MODE_CHECK) track_push(mid, limit, 1, sa == 0, a_dim, a_split);
            const int child1 = desc_a ? (sa == 0 ? a.less : a.greater) : n1;
LBL_UNROLL
            for (int sb = 0; sb < 2; ++sb) {
              if (!desc_b && sb == 1) continue;
              if (desc_a && desc_b && n1 == n2 && sa == 1 && sb == 0) continue;
              Rect child = mid;
              if (desc_b && mode == MODE_CHECK) track_push(child, limit, 2, sb == 0, b_dim, b_split);
              const int child2 = desc_b ? (sb == 0 ? b.less : b.greater) : n2;
              const int slot = 2 * sa + sb;
              c1[slot] = child1; c2[slot] = child2;
              cm[slot] = enter(child1, child2, mode, child, upper);
              cmin[slot] = child.min_d; cmax[slot] = child.max_d;
            }
          }
        }
        int mine = 0;
        if (active && !open) mine = 1;                      // a leaf block: carried along
        else mine = (cm[0] >= 0) + (cm[1] >= 0) + (cm[2] >= 0) + (cm[3] >= 0);
        // exclusive prefix of the counts over the lanes
        int incl = mine;
        for (int w = 1; w < LBL_WARP; w <<= 1) {
          const int v = shfl_up_int(incl, w);
          if (lane >= w) incl += v;
        }
        const int total = shfl(incl, LBL_WARP - 1);
        int pos = n_next + incl - mine;
        if (n_next + total > caps.tasks) { flag = FLAG_CAPACITY; return 0; }
        if (active && !open) {
          to[pos].n1 = n1; to[pos].n2 = n2; to[pos].mode = mode; to[pos].pad_ = 0;
        } else if (open) {
LBL_UNROLL
          for (int j = 0; j < 4; ++j) {
            if (cm[j] < 0) continue;
            Rect child = parent;
            if (desc_a) { if ((j >> 1) == 0) set3(child.r1mx, a_dim, a_split); else set3(child.r1mn, a_dim, a_split); }
            if (desc_b) { if ((j & 1) == 0) set3(child.r2mx, b_dim, b_split); else set3(child.r2mn, b_dim, b_split); }
            child.min_d = cmin[j]; child.max_d = cmax[j];
            store_task(to + pos, c1[j], c2[j], cm[j], child);
            any_open = any_open || cm[j] < MODE_LEAVES;
            ++pos;
          }
        }
        n_next += total;
      }
      warp_sync();
      cur ^= 1;
      n_cur = n_next;
      if (!ballot(any_open)) break;
    }
    // the leaf blocks, in order
#if defined(CTK_LABEL_TIMING) && !defined(CTK_EMUL)
    const long long leaf0_ = clock64();
#endif
    int np = 0;
    const Task* list = cur ? s.tasks[1] : s.tasks[0];
    for (int base = 0; base < n_cur; base += LBL_WARP) {
      const int t = base + lane;
      int n1 = 0, n2 = 0, mode = 0;
      if (t < n_cur) { n1 = list[t].n1; n2 = list[t].n2; mode = list[t].mode; }
      const int count = n_cur - base < LBL_WARP ? n_cur - base : LBL_WARP;
      for (int k = 0; k < count; ++k)
        leaf_block(shfl(n1, k), shfl(n2, k), shfl(mode, k) == MODE_LEAVES, upper, np);
      if (np > caps.pairs) { flag = FLAG_CAPACITY; break; }
    }
    warp_sync();
#if defined(CTK_LABEL_TIMING) && !defined(CTK_EMUL)
    if (lane == 0 && timing) { atomicAdd(timing + 6, (unsigned long long) (clock64() - leaf0_)); atomicAdd(timing + 7, (unsigned long long) n_cur * 1000000ULL); }
#endif
    return np;
  }

  LBL_DEV int shfl_up_int(int v, int delta) const {
#ifdef CTK_EMUL
    (void) delta;
    return v;
#else
    return __shfl_up_sync(0xffffffffu, v, delta);
#endif
  }

  // ---- CPython set order + union -------------------------------------------------------------
  LBL_DEV static uint64_t tuple2_hash(uint64_t a, uint64_t b) {
    const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL,
                   P5 = 2870177450012600261ULL;
    uint64_t acc = P5;
    acc += a * P2; acc = (acc << 31) | (acc >> 33); acc *= P1;
    acc += b * P2; acc = (acc << 31) | (acc >> 33); acc *= P1;
    acc += 2ULL ^ (P5 ^ 3527539ULL);
    if (acc == (uint64_t) -1) acc = 1546275796ULL;
    return acc;
  }
  // set_add_entry for a key known to be absent / set_insert_clean: the same probe sequence
  LBL_DEV static void set_insert(int32_t* table, uint64_t mask, int item, uint64_t hash) {
    uint64_t perturb = hash, i = hash & mask;
    for (;;) {
      if (table[i] < 0) { table[i] = item; return; }
      if (i + 9 <= mask) {
        for (uint64_t j = 1; j <= 9; ++j)
          if (table[i + j] < 0) { table[i + j] = item; return; }
      }
      perturb >>= 5;
      i = (i * 5 + 1 + perturb) & mask;
    }
  }

  // Where a table of `size` slots lives: in the fast window (alternating ends, so that the table
  // being filled never overlaps the one being read, and leaving `keep` bytes for the union arrays)
  // or, once it does not fit, in the scratch.
  LBL_DEV int32_t* table_home(uint64_t size, uint64_t other_size, int which, int64_t keep, bool* fast_ok) const {
    const int64_t bytes = (int64_t) size * 4, other = (int64_t) other_size * 4;
    if (*fast_ok && s.fast && bytes + other <= s.fast_bytes && bytes + keep <= s.fast_bytes)
      return reinterpret_cast<int32_t*>(which == 0 ? s.fast : s.fast + (s.fast_bytes - bytes));
    *fast_ok = false;
    return which ? s.table[1] : s.table[0];
  }

  // -> the table that holds the final set, *mask_out its mask, *which_out its end of the window
  LBL_DEV int32_t* set_order(int np, int64_t keep, uint64_t* mask_out, int* which_out) {
    int cur = 0;
    uint64_t mask = 7;
    bool fast_ok = true;
    int32_t* table = table_home(8, 0, 0, keep, &fast_ok);
    for (int i = lane; i < 8; i += LBL_WARP) table[i] = -1;
    warp_sync();
    uint64_t fill = 0;
    for (int base = 0; base < np; base += LBL_WARP) {
      // hashes of the next pairs on all lanes, inserts in order on lane 0
      uint64_t hash = 0;
      if (base + lane < np) { const Pair p = s.pairs[base + lane]; hash = tuple2_hash((uint64_t) p.i, (uint64_t) p.j); }
      const int count = np - base < LBL_WARP ? np - base : LBL_WARP;
      s.stage[lane] = hash;
      warp_sync();
      int done = 0;
      while (done < count) {
        if (lane == 0) {
          while (done < count) {
            set_insert(table, mask, base + done, s.stage[done]);
            ++done; ++fill;
            if (fill * 5 >= mask * 3) break;
          }
        }
        done = shfl(done, 0);
        fill = shfl(fill, 0);
        warp_sync();
        if (!(fill * 5 >= mask * 3)) continue;
        // set_table_resize(used > 50000 ? 2 * used : 4 * used): re-insert in slot order
        const uint64_t minused = fill > 50000 ? fill * 2 : fill * 4;
        uint64_t newsize = 8;
        while (newsize <= minused) newsize <<= 1;
        int32_t* to = table_home(newsize, mask + 1, cur ^ 1, keep, &fast_ok);
        for (uint64_t i = lane; i < newsize; i += LBL_WARP) to[i] = -1;
        warp_sync();
        for (uint64_t b2 = 0; b2 <= mask; b2 += LBL_WARP) {
          const int item = b2 + lane <= mask ? table[b2 + lane] : -1;
          uint64_t h2 = 0;
          if (item >= 0) { const Pair p = s.pairs[item]; h2 = tuple2_hash((uint64_t) p.i, (uint64_t) p.j); }
          uint32_t bits = ballot(item >= 0);
          while (bits) {
            const int src = ctz(bits);
            bits &= bits - 1;
            const int it = shfl(item, src);
            const uint64_t hh = shfl(h2, src);
            if (lane == 0) set_insert(to, newsize - 1, it, hh);
          }
        }
        warp_sync();
        cur ^= 1;
        table = to;
        mask = newsize - 1;
      }
    }
    *mask_out = mask;
    *which_out = cur;
    return table;
  }
  LBL_DEV void label_union(int np) {
    // union arrays: in the fast window when the final table leaves room (decided with `keep`)
    const int64_t keep = (int64_t) n * 12 + 64;
    uint64_t mask = 7;
    int which = 0;
    LBL_TICK_DECL;
    int32_t* table = np > 0 ? set_order(np, keep, &mask, &which) : nullptr;
    LBL_TICK(3);
    int32_t *label = s.label, *next = s.next, *tail = s.tail;
    if (s.fast && keep <= s.fast_bytes) {
      const bool table_fast = np > 0 && reinterpret_cast<char*>(table) >= s.fast &&
                              reinterpret_cast<char*>(table) < s.fast + s.fast_bytes;
      const int64_t table_bytes = table_fast ? (int64_t) (mask + 1) * 4 : 0;
      if (table_bytes + keep <= s.fast_bytes) {
        // the table sits at one end of the window (which == 0: the low end): take the other end
        char* home = (table_fast && which == 0) ? s.fast + (s.fast_bytes - keep) : s.fast;
        home = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(home) + 15) & ~(uintptr_t) 15);
        label = reinterpret_cast<int32_t*>(home);
        next = label + n;
        tail = next + n;
      }
    }
    for (int i = lane; i < n; i += LBL_WARP) { label[i] = i; next[i] = -1; tail[i] = i; }
    warp_sync();
    if (np > 0) {
      for (uint64_t base = 0; base <= mask; base += LBL_WARP) {
        const int item = base + lane <= mask ? table[base + lane] : -1;
        Pair p = Pair{0, 0};
        if (item >= 0) p = s.pairs[item];
        uint32_t bits = ballot(item >= 0);
        while (bits) {
          const int src = ctz(bits);
          bits &= bits - 1;
          const int pi = shfl(p.i, src), pj = shfl(p.j, src);
          if (lane == 0) {
            const int keep_id = label[pi], drop = label[pj];
            if (keep_id != drop) {
              for (int q = drop; q >= 0; q = next[q]) label[q] = keep_id;
              next[tail[keep_id]] = drop;
              tail[keep_id] = tail[drop];
            }
          }
        }
      }
      warp_sync();
    }
    out_label_ = label;
  }
  int32_t* out_label_;

  // ---- driver ------------------------------------------------------------------------------------
  int n_pairs_ = 0;        // pairs found by the last run (s.pairs holds them in scipy's order)

  // pos[k]: table-order column k; the frame owns rows a .. a + n - 1.  labels_out [rows] int32.
  LBL_DEV int run(const double* const* pos, int64_t a, int n_, int m_, const double* separation,
                  double r, int32_t* labels_out) {
    n = n_; m = m_; lane = lane_id(); flag = FLAG_OK;
    if (n > caps.points) return FLAG_CAPACITY;
    LBL_TICK_DECL;
    bool finite = true;
    LBL_UNROLL
    for (int k = 0; k < 3; ++k) { mins[k] = 0.; maxes[k] = 0.; }
    LBL_UNROLL
    for (int k = 0; k < 3; ++k) {
      if (k >= m) continue;
      const double* src = pos[k] + a;
      const double sep = separation[k];
      double* c = s.c[k];
      for (int i = lane; i < n; i += LBL_WARP) {
        const double v = ddiv(src[i], sep);
        c[i] = v;
        finite = finite && is_finite(v);
      }
    }
    for (int i = lane; i < n; i += LBL_WARP) s.idx[i] = i;
    warp_sync();
    if (ballot(!finite)) return FLAG_NONFINITE;
    bounds_warp(0, n, mins, maxes);
    LBL_TICK(0);
    build();
    LBL_TICK(1);
    if (flag) return flag;
    const int np = query(r);
    n_pairs_ = np;
    LBL_TICK(2);
    if (flag) return flag;
    label_union(np);
    for (int i = lane; i < n; i += LBL_WARP) labels_out[a + i] = out_label_[i];
    warp_sync();
    LBL_TICK(4);                             // union + output (includes slot 3 = set order)
    return FLAG_OK;
  }
};

}  // namespace ctk_label
