#!/bin/bash
set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err
tail -c 400 gpurun_out/r02d_bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02d_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'f64', d['value_f64']['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['host_ms'], d['e2e']['labelling'], 'pageable', d['e2e'].get('pageable', {}).get('value'))
print(d['roofline']['frac'], d['gpu_launches'], d['parity_vs_oracle'])
PY
