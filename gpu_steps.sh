#!/bin/bash
CTK_BENCH_STEPS=1 python bench.py --gpus 1 --steps 6 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "e2e step" | tail -2
