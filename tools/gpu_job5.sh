#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -4 gpurun_out/pytest_gpu.txt
run() { # name, env...
  echo "=== $1"; shift
  env "$@" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l.csv \
      python tools/quick_bench.py --cfg 2 --algos binned --reps 1 > gpurun_out/ncu_run.txt 2>&1
  grep -E "bin_points|reduce_tiles" gpurun_out/l.csv | awk -F'","' '{print substr($5,1,46), $NF}' | sed -n '3,4p;7,8p'
}
run "th6 agg1" LM_BEV_TILE_H_LOG2=6 LM_BEV_WARP_AGG=1
run "th6 agg0" LM_BEV_TILE_H_LOG2=6 LM_BEV_WARP_AGG=0
run "th7 agg1" LM_BEV_TILE_H_LOG2=7 LM_BEV_WARP_AGG=1
run "th7 agg0" LM_BEV_TILE_H_LOG2=7 LM_BEV_WARP_AGG=0
LM_BEV_TILE_H_LOG2=7 timeout 600 python tools/quick_bench.py --cfg 2 --algos binned 2>&1 | grep -v generated
