#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { # name, nvcc-extra, env...
  echo "=== $1"; local extra="$2"; shift; shift
  LM_BEV_NVCC_EXTRA="$extra" python -c "from lanemapping_b200.build import build_native; build_native(force=True)" || return
  env "$@" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l.csv \
      python tools/quick_bench.py --cfg 2 --algos binned --reps 1 > gpurun_out/ncu_run.txt 2>&1
  grep -E "bin_points|reduce_tiles" gpurun_out/l.csv | awk -F'","' '{print substr($5,1,46), $NF}' | sed -n '3,4p;7,8p'
}
run "256x4 min4 th6" "" LM_BEV_TILE_H_LOG2=6
run "256x4 min4 th7" "" LM_BEV_TILE_H_LOG2=7
run "256x4 min5 th6" "-DLM_BIN_MIN_CTAS=5" LM_BEV_TILE_H_LOG2=6
run "128x8 min6 th6" "-DLM_BIN_THREADS=128 -DLM_BIN_PPT=8 -DLM_BIN_MIN_CTAS=6" LM_BEV_TILE_H_LOG2=6
run "512x2 min2 th6" "-DLM_BIN_THREADS=512 -DLM_BIN_PPT=2 -DLM_BIN_MIN_CTAS=2" LM_BEV_TILE_H_LOG2=6
run "256x2 min6 th6" "-DLM_BIN_THREADS=256 -DLM_BIN_PPT=2 -DLM_BIN_MIN_CTAS=6" LM_BEV_TILE_H_LOG2=6
python -c "from lanemapping_b200.build import build_native; build_native(force=True)"
LM_BEV_TILE_H_LOG2=6 bash tools/gpu_prof.sh scan v6
