#!/bin/bash
# ncu capture of one refine launch: ./profiles/tools/gpu_prof.sh <tag> <launch index> [lib]
tag=$1; idx=$2; lib=${3:-libctk.so}
CTK_LIB_PATH=$PWD/clustertracking_b200/$lib ncu --set full --clock-control none --import-source on -k regex:refine_kernel -s $idx -c 1 -f -o gpurun_out/prof_$tag python profiles/tools/class_times.py 300 > gpurun_out/prof_$tag.log 2>&1
tail -3 gpurun_out/prof_$tag.log
