#!/bin/bash
python -m pytest tests/test_gpu_label.py -x -q 2>&1 | tail -1
for b in 2 3 4; do
echo "blocks per SM $b"
CTK_LABEL_BLOCKS_PER_SM=$b python profiles/tools/label_bench.py 1000 2>&1 | tail -1 | cut -c1-130
CTK_LABEL_BLOCKS_PER_SM=$b CTK_BENCH_STEPS=1 python bench.py --gpus 1 --steps 12 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "e2e step" | awk -F'[(,]' '{printf "%s ", $2} END {print ""}'
done
