python bench.py --frames 300 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['value']/1e6,2), 'M/s  ms', round(d['ms_per_step'],3), ' e2e', round(d['e2e']['value']/1e6,2), 'M/s', 'failed', d['config']['failed_clusters'])"
