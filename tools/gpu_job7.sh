#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { # name, nvcc-extra, env...
  echo "=== $1"; local extra="$2"; shift; shift
  LM_BEV_NVCC_EXTRA="$extra" python -c "from lanemapping_b200.build import build_native; build_native(force=True)" || return
  env "$@" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l.csv \
      python tools/quick_bench.py --cfg 2 --algos binned --reps 1 > gpurun_out/ncu_run.txt 2>&1
  grep -E "bin_points|reduce_tiles" gpurun_out/l.csv | awk -F'","' '{print substr($5,1,46), $NF}' | sed -n '3,4p;7,8p'
}
run "direct 512x4 th6" "" LM_BEV_STAGED_STORE=0 LM_BEV_TILE_H_LOG2=6
run "staged 512x4 th6" "" LM_BEV_STAGED_STORE=1 LM_BEV_TILE_H_LOG2=6
run "direct 512x4 th7" "" LM_BEV_STAGED_STORE=0 LM_BEV_TILE_H_LOG2=7
run "direct 256x4 min4 th6" "-DLM_BIN_THREADS=256 -DLM_BIN_PPT=4 -DLM_BIN_MIN_CTAS=4" LM_BEV_STAGED_STORE=0 LM_BEV_TILE_H_LOG2=6
run "direct 256x8 min3 th6" "-DLM_BIN_THREADS=256 -DLM_BIN_PPT=8 -DLM_BIN_MIN_CTAS=3" LM_BEV_STAGED_STORE=0 LM_BEV_TILE_H_LOG2=6
run "direct 512x4 min3 th6" "-DLM_BIN_MIN_CTAS=3" LM_BEV_STAGED_STORE=0 LM_BEV_TILE_H_LOG2=6
python -c "from lanemapping_b200.build import build_native; build_native(force=True)"
