#!/bin/bash
CTK_BENCH_STEPS=1 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "e2e step" | awk -F'[(,]' '{printf "%s ", $2} END {print ""}'
CTK_SWITCH_INTERVAL=0.005 CTK_BENCH_STEPS=1 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "e2e step" | awk -F'[(,]' '{printf "%s ", $2} END {print ""}'
CTK_BENCH_STEPS=1 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "e2e step" | tail -1 | cut -c1-400
