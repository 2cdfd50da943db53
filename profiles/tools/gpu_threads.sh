#!/bin/bash
# usage: gpu_threads.sh N T1 T2 ... -- sharded bench at N GPUs with CTK_HOST_THREADS = T
N=$1; shift
for T in "$@"; do
CTK_HOST_THREADS=$T python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$T bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r02t_bench_n${N}_t$T.json 2> gpurun_out/r02t_bench_n${N}_t$T.err
python - <<PY
import json
name='gpurun_out/r02t_bench_n${N}_t$T.json'
try:
    d = json.loads(open(name).read().strip().splitlines()[-1]); e = d['e2e']
    print('threads $T', 'e2e %.3g' % e['value'], '%.1f ms' % e['ms_per_step'], e['host_ms'])
except Exception as exc:
    print(name, 'FAILED', exc)
PY
done
