#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -5 gpurun_out/pytest_gpu.txt
timeout 600 python tools/quick_bench.py --cfg 2 --algos binned > gpurun_out/quick_cfg2.txt 2>&1
LM_BEV_TILE_H_LOG2=7 timeout 600 python tools/quick_bench.py --cfg 2 --algos binned >> gpurun_out/quick_cfg2.txt 2>&1
cat gpurun_out/quick_cfg2.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg2.csv \
    python tools/quick_bench.py --cfg 2 --algos binned --reps 1 > gpurun_out/ncu_run.txt 2>&1
grep -E "bin_points|reduce_tiles|scan_tiles|index_chunks" gpurun_out/launches_cfg2.csv | awk -F'","' '{print substr($5,1,40), $NF}' | tail -8
LM_BEV_TILE_H_LOG2=7 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg2_th7.csv \
    python tools/quick_bench.py --cfg 2 --algos binned --reps 1 > gpurun_out/ncu_run.txt 2>&1
grep -E "bin_points|reduce_tiles|scan_tiles|index_chunks" gpurun_out/launches_cfg2_th7.csv | awk -F'","' '{print substr($5,1,40), $NF}' | tail -8
bash tools/gpu_prof.sh scan v3
