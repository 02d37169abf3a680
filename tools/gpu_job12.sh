#!/bin/bash
mkdir -p gpurun_out
run() { # name, nvcc-extra, env...
  echo "=== $1"; local extra="$2"; shift; shift
  LM_BEV_NVCC_EXTRA="$extra" python -c "from lanemapping_b200.build import build_native; build_native(force=True)" || return
  env "$@" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l.csv \
      python tools/quick_bench.py --cfg 2 --algos binned --reps 1 > gpurun_out/ncu_run.txt 2>&1
  grep -E "bin_points|reduce_tiles" gpurun_out/l.csv | awk -F'","' '{print substr($5,1,46), $NF}' | sed -n '3,4p;7,8p'
  tail -2 gpurun_out/ncu_run.txt | cut -c1-120
}
run "chunk512 slots4 min3 (default)" ""
run "chunk1024 slots2 min4" "-DLM_CHUNK_LOG2=10 -DLM_BIN_MIN_CTAS=4"
run "chunk1024 slots2 min4 th7" "-DLM_CHUNK_LOG2=10 -DLM_BIN_MIN_CTAS=4" LM_BEV_TILE_H_LOG2=7
run "chunk1024 slots2 min5" "-DLM_CHUNK_LOG2=10 -DLM_BIN_MIN_CTAS=5"
LM_BEV_NVCC_EXTRA="-DLM_CHUNK_LOG2=10 -DLM_BIN_MIN_CTAS=4" python -c "from lanemapping_b200.build import build_native; build_native(force=True)"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "from lanemapping_b200.build import build_native; build_native(force=True)"
