#!/bin/bash
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_accuracy.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -4
python bench.py --gpus 1 --steps 10 --warmup 3 --cpu-frames 1 > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
tail -c 600 gpurun_out/r02c_bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02c_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'f64', d['value_f64']['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['host_ms'], 'pageable', d['e2e'].get('pageable', {}).get('value'))
print(d['roofline']['frac'], d['parity_vs_oracle'])
PY
