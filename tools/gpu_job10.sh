#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_n1.err
timeout 600 python tools/quick_bench.py --cfg 4 --algos binned 2>&1 | grep -v generated
timeout 600 python tools/quick_bench.py --cfg 1 --algos binned 2>&1 | grep -v generated
