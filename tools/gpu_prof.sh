#!/bin/bash
# ncu full captures of bin_points / reduce_tiles on cfg2; args: order (scan|shuffled) tag
ORDER=${1:-scan}; TAG=${2:-x}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bin_points -s 2 -c 1 -o gpurun_out/prof_bin_$TAG \
    python tools/quick_bench.py --cfg 2 --orders $ORDER --algos binned --reps 1 > gpurun_out/ncu_bin.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:reduce_tiles -s 2 -c 1 -o gpurun_out/prof_red_$TAG \
    python tools/quick_bench.py --cfg 2 --orders $ORDER --algos binned --reps 1 > gpurun_out/ncu_red.txt 2>&1
ls -la gpurun_out/*.ncu-rep
