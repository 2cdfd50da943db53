#!/bin/bash
set -x
python -m pytest tests/test_gpu_label.py -x -q 2>&1 | grep -E "Error|assert|FAILED|passed|failed" | head -20
CTK_LIB_PATH=/root/repo/profiles/tools/_build/libctk_timing.so python profiles/tools/label_bench.py 1000 2>&1 | tail -1 | tee gpurun_out/r02_label_bench_timing.json
python profiles/tools/label_bench.py 1000 2>&1 | tail -1 | tee gpurun_out/r02_label_bench.json
