#!/bin/bash
# first GPU call: microbench, parity tests, quick timings, launch list
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 ./tools/ubench > gpurun_out/ubench.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
timeout 600 python tools/quick_bench.py --cfg 2 > gpurun_out/quick_cfg2.txt 2>&1
timeout 300 python tools/quick_bench.py --cfg 1 >> gpurun_out/quick_cfg2.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg2.csv \
    python tools/quick_bench.py --cfg 2 --orders scan --algos binned --reps 1 > gpurun_out/ncu_run.txt 2>&1
tail -5 gpurun_out/pytest_gpu.txt
cat gpurun_out/ubench.txt gpurun_out/quick_cfg2.txt
