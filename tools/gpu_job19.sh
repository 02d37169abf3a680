#!/bin/bash
# post-processing kernels: tests; then ncu evidence of the current build (launch list of the bench command + full captures)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_post.py -x -q 2>&1 | tail -15
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_v9_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
tail -2 gpurun_out/b_ncu.log | cut -c1-300
bash tools/gpu_prof.sh scan v9
