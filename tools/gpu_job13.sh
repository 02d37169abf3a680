#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l.csv \
      python tools/quick_bench.py --cfg 2 --algos binned --reps 1 > gpurun_out/ncu_run.txt 2>&1
grep -E "bin_points|reduce_tiles|index_chunks|scan_tiles" gpurun_out/l.csv | awk -F'","' '{print substr($5,1,46), $NF}' | sed -n '5,8p;13,16p'
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json | cut -c1-2000; tail -n 3 gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.txt 2>&1
LM_BEV_TILE_H_LOG2=6 bash tools/gpu_prof.sh scan v8
