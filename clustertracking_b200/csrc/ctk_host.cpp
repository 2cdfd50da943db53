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
