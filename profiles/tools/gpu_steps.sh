#!/bin/bash
python -m pytest tests/test_gpu_label.py tests/test_gpu_sharding.py -x -q 2>&1 | tail -1
CTK_BENCH_STEPS=1 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "e2e step" | awk -F'[(,]' '{printf "%s ", $2} END {print ""}'
CTK_BENCH_STEPS=1 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "e2e step" | tail -1 | cut -c1-420
