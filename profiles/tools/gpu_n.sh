#!/bin/bash
# usage: gpu_n.sh N  -- sharded bench at N GPUs, device labels vs host labels
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r02b_bench_n$N.json 2> gpurun_out/r02b_bench_n$N.err
#CTK_LABEL_DEVICE=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r02b_bench_n${N}_hostlabels.json 2> gpurun_out/r02b_bench_n$N.err2
python - <<PY
import json
for name in ('gpurun_out/r02b_bench_n$N.json', 'gpurun_out/r02b_bench_n${N}_hostlabels.json'):
    try:
        d = json.loads(open(name).read().strip().splitlines()[-1])
        e = d['e2e']
        print(name, 'value %.3g' % d['value'], 'e2e %.3g' % e['value'], '%.1f ms' % e['ms_per_step'], e['host_ms'], e.get('labelling'), 'strong', e.get('strong'), e.get('sharded_ms_rank0'))
    except Exception as exc:
        print(name, 'FAILED', exc)
PY
tail -3 gpurun_out/r02b_bench_n$N.err
nproc
