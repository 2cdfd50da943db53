#!/bin/bash
# round-2 measurement pass on one GPU: bench, launch list, full captures of the dominant refine launch and of the label kernel
set -x
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -c 300 gpurun_out/r02_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:refine_kernel|frame_max_kernel|label_kernel|global_kernel" -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_list.log 2>&1
tail -2 gpurun_out/r02_ncu_list.log | cut -c1-300
# dominant launch: class 6 main = 7th refine_kernel launch of a step (index 6)
ncu --set full --clock-control none --import-source on -k regex:refine_kernel -s 6 -c 1 -f -o gpurun_out/r02_refine_c6 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_full.log 2>&1
tail -2 gpurun_out/r02_ncu_full.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:label_kernel -c 1 -f -o gpurun_out/r02_label python profiles/tools/label_bench.py 1000 > gpurun_out/r02_ncu_label.log 2>&1
tail -2 gpurun_out/r02_ncu_label.log | cut -c1-300
python profiles/tools/config_bench.py 100 20 100 > gpurun_out/r02_config_bench.jsonl 2> gpurun_out/r02_config_bench.err
cat gpurun_out/r02_config_bench.jsonl | cut -c1-400
