#!/bin/bash
for mode in 1 0; do
echo "CTK_LABEL_DEVICE=$mode"
CTK_LABEL_DEVICE=$mode CTK_BENCH_STEPS=1 python bench.py --gpus 1 --steps 15 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "e2e step" | awk -F'[(,]' '{printf "%s ", $2} END {print ""}'
done
nproc
