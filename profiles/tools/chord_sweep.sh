# resident throughput of config 2 (300 frames) for several chord tolerances
for tol in 0.02 0.05 0.1 0.2 0.5 1.0; do
  echo -n "chord_tol $tol: "
  CTK_CHORD_TOL=$tol python bench.py --frames 300 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['value']/1e6,2), 'M/s  ms', round(d['ms_per_step'],3), 'evals', round(d['config']['mean_evaluations_per_cluster'],2), 'failed', d['config']['failed_clusters'])"
done
