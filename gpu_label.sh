#!/bin/bash
set -x
python -m pytest tests/test_gpu_label.py -x -q 2>&1 | grep -E "Error|assert|FAILED|passed|failed" | head -20
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err
tail -c 1500 gpurun_out/r02b_bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02b_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['host_ms'], d['e2e']['labelling'], 'pageable', d['e2e'].get('pageable'))
PY
CTK_LABEL_DEVICE=0 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_n1_hostlabels.json 2> gpurun_out/r02b_bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02b_bench_n1_hostlabels.json').read().strip().splitlines()[-1])
print('HOST LABELS value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['host_ms'], d['e2e']['labelling'], 'pageable', d['e2e'].get('pageable'))
PY
