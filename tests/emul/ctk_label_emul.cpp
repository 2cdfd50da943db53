// ctk_label_emul.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the device labelling source (clustertracking_b200/csrc/ctk_label.cuh) as plain C++ with a
// one-lane "warp", so that the CPU tests can hold it to the host restatement (ctk_cluster_frames,
// itself verified against scipy) in the GPU-less build container.  Never loaded by the package.
#define CTK_EMUL 1
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "ctk_label.cuh"

extern "C" int ctk_emul_label_frames(const double* const* pos_cols, int32_t ndim,
                                     const int64_t* starts, const int64_t* stops, int64_t n_frames,
                                     const double* separation, int64_t pair_factor, int64_t window,
                                     int32_t* labels_out, int32_t* sizes_out, int32_t* flags_out) {
  int64_t max_points = 1;
  for (int64_t f = 0; f < n_frames; ++f) max_points = std::max(max_points, stops[f] - starts[f]);
  const ctk_label::Caps caps = ctk_label::make_caps(max_points, pair_factor);
  const int64_t bytes = ctk_label::scratch_bytes(caps);
  char* mem = static_cast<char*>(aligned_alloc(128, (size_t) bytes));
  // window < 0: as large as the kernel would make it; 0: none; else that many bytes
  if (window < 0) window = ctk_label::window_bytes(max_points, ndim);
  char* fast = window > 0 ? static_cast<char*>(aligned_alloc(128, (size_t) ctk_label::align_up(window, 128))) : nullptr;
  for (int64_t f = 0; f < n_frames; ++f) {
    memset(mem, 0xCD, (size_t) bytes);                 // poison: stale data must never be relied on
    if (fast) memset(fast, 0xCD, (size_t) window);
    ctk_label::FrameLabeller fl;
    fl.s = ctk_label::carve(mem, caps);
    fl.caps = caps;
    const int n = (int) (stops[f] - starts[f]);
    if (fast) ctk_label::use_window(fl.s, fast, window, n, ndim);
    flags_out[f] = n == 0 ? 0 : fl.run(pos_cols, starts[f], n, ndim, separation, 1.0, labels_out);
    if (sizes_out && flags_out[f] == 0) {
      std::vector<int32_t> count((size_t) n, 0);
      for (int i = 0; i < n; ++i) ++count[labels_out[starts[f] + i]];
      for (int i = 0; i < n; ++i) sizes_out[starts[f] + i] = count[labels_out[starts[f] + i]];
    }
  }
  free(mem);
  free(fast);
  return 0;
}

// The restated std::nth_element against the real one, on inputs that drive the real one into its
// depth limit (heap select): McIlroy's adversary ("A killer adversary for quicksort") decides the
// key values lazily while std::nth_element runs, which yields a concrete worst-case input.
namespace {
struct Adversary {
  std::vector<int> val;
  int nsolid, candidate, gas;
  explicit Adversary(int n) : val(n), nsolid(0), candidate(0), gas(n) { std::fill(val.begin(), val.end(), gas); }
  bool less(int x, int y) {
    if (val[x] == gas && val[y] == gas) {
      if (x == candidate) val[x] = nsolid++; else val[y] = nsolid++;
    }
    if (val[x] == gas) candidate = x; else if (val[y] == gas) candidate = y;
    return val[x] < val[y];
  }
};
}  // namespace

// -> 0 when the restatement leaves the records in the same order as std::nth_element for (a) an
// adversarial input of n keys and (b) `random_cases` random inputs with many ties.
extern "C" int ctk_emul_nth_element_check(int32_t n, int32_t nth, int32_t random_cases, uint32_t seed,
                                          int32_t* hit_heap_out) {
  std::vector<std::vector<double>> inputs;
  {
    Adversary adv(n);
    std::vector<int> items(n);
    for (int i = 0; i < n; ++i) items[i] = i;
    std::nth_element(items.begin(), items.begin() + nth, items.end(),
                     [&adv](int x, int y) { return adv.less(x, y); });
    std::vector<double> keys(n);
    for (int i = 0; i < n; ++i) keys[i] = (double) adv.val[i];
    inputs.push_back(keys);
  }
  uint32_t state = seed ? seed : 1u;
  auto rnd = [&state]() { state ^= state << 13; state ^= state >> 17; state ^= state << 5; return state; };
  for (int c = 0; c < random_cases; ++c) {
    std::vector<double> keys(n);
    const uint32_t range = 1 + rnd() % (uint32_t) (2 * n);
    for (int i = 0; i < n; ++i) keys[i] = (double) (rnd() % range);
    inputs.push_back(keys);
  }
  int heap_hits = 0;
  for (size_t c = 0; c < inputs.size(); ++c) {
    const std::vector<double>& keys = inputs[c];
    struct Rec { double key; int idx; };
    std::vector<Rec> want(n);
    for (int i = 0; i < n; ++i) want[i] = Rec{keys[i], i};
    long comparisons = 0;
    std::nth_element(want.begin(), want.begin() + nth, want.end(),
                     [&comparisons](const Rec& a, const Rec& b) { ++comparisons; return a.key < b.key; });
    if (comparisons > 6L * n) ++heap_hits;              // far beyond the average: the fallback ran
    std::vector<double> c0(keys), c1(n, 0.), c2(n, 0.);
    std::vector<int32_t> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    ctk_label::Points p;
    p.c[0] = c0.data(); p.c[1] = c1.data(); p.c[2] = c2.data();
    p.idx = idx.data(); p.key = c0.data(); p.m = 1;
    ctk_label::nth_element(p, 0, nth, n);
    for (int i = 0; i < n; ++i)
      if (idx[i] != want[i].idx || c0[i] != want[i].key) return (int) c + 1;
  }
  if (hit_heap_out) *hit_heap_out = heap_hits;
  return 0;
}

// The close pairs of ONE point set in the order the device code reports them (= the order
// scipy's query_pairs(output_type='ndarray') reports them).  pairs_out [capacity, 2] int64.
extern "C" int ctk_emul_query_pairs(const double* const* pos_cols, int32_t ndim, int64_t n,
                                    int64_t pair_factor, int64_t window, int64_t* pairs_out,
                                    int64_t capacity, int64_t* n_pairs_out) {
  const ctk_label::Caps caps = ctk_label::make_caps(n, pair_factor);
  const int64_t bytes = ctk_label::scratch_bytes(caps);
  char* mem = static_cast<char*>(aligned_alloc(128, (size_t) bytes));
  if (window < 0) window = ctk_label::window_bytes(n, ndim);
  char* fast = window > 0 ? static_cast<char*>(aligned_alloc(128, (size_t) ctk_label::align_up(window, 128))) : nullptr;
  memset(mem, 0xCD, (size_t) bytes);
  ctk_label::FrameLabeller fl;
  fl.s = ctk_label::carve(mem, caps);
  fl.caps = caps;
  if (fast) ctk_label::use_window(fl.s, fast, window, (int) n, ndim);
  const ctk_label::Pair* pairs = fl.s.pairs;
  const double ones[3] = {1., 1., 1.};
  std::vector<int32_t> labels((size_t) n);
  const int flag = fl.run(pos_cols, 0, (int) n, ndim, ones, 1.0, labels.data());
  *n_pairs_out = fl.n_pairs_;
  if (flag == 0 && fl.n_pairs_ <= capacity)
    for (int k = 0; k < fl.n_pairs_; ++k) { pairs_out[2 * k] = pairs[k].i; pairs_out[2 * k + 1] = pairs[k].j; }
  free(mem);
  free(fast);
  return flag;
}
