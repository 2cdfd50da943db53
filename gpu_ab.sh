#!/bin/bash
# occupancy experiment: pad the dynamic shared memory so that fewer blocks fit an SM
for pad in 0 26 63 160; do
  echo "=== CTK_SMEM_PAD=$pad"
  CTK_SMEM_PAD=$pad python profiles/tools/class_times.py 300 2>&1 | grep -E "main|total"
done
