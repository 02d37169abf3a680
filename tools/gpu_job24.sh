#!/bin/bash
# round-1 checkpoint on the current kernels (v10): full gpu suite, bench both arms, launch list, ncu full of bin + reduce
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1_v10.json 2> gpurun_out/bench_n1_v10.err
cut -c1-1500 gpurun_out/bench_n1_v10.json; tail -n 3 gpurun_out/bench_n1_v10.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_v10.json 2> gpurun_out/bench_ref_v10.err
cut -c1-800 gpurun_out/bench_ref_v10.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v10_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
bash tools/gpu_prof.sh scan v10
python tools/ncu_summary.py gpurun_out/prof_bin_v10.ncu-rep > gpurun_out/v10_bin_summary.txt 2>&1
python tools/ncu_summary.py gpurun_out/prof_red_v10.ncu-rep > gpurun_out/v10_red_summary.txt 2>&1
head -c 1500 gpurun_out/v10_bin_summary.txt
