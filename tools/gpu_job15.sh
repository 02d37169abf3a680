#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/quick_bench.py --cfg 4 --algos binned 2>&1 | grep -v generated
timeout 600 python tools/quick_bench.py --cfg 5 2>&1 | grep -v generated
timeout 600 python tools/quick_bench.py --cfg 1 --algos binned 2>&1 | grep -v generated
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l4.csv \
      python tools/quick_bench.py --cfg 4 --orders scan --algos binned --reps 1 > gpurun_out/ncu_run.txt 2>&1
grep -E "bin_points|reduce_tiles|index_chunks|scan_tiles" gpurun_out/l4.csv | awk -F'","' '{print substr($5,1,46), $NF}' | tail -8
