#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:refine_kernel|frame_max_kernel|label_kernel|global_kernel' -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_list.log 2>&1
tail -1 gpurun_out/r02_ncu_list.log | cut -c1-200
wc -l gpurun_out/r02_launches.csv
