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
