#!/bin/bash
# v11 final build (store-side L2 hint compiled out again): gpu suite + bench line
mkdir -p gpurun_out
LM_SKIP_FULLSIZE=1 timeout 60 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
timeout 80 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1_v11b.json 2> gpurun_out/bench_n1_v11b.err
cut -c1-1800 gpurun_out/bench_n1_v11b.json; tail -n 2 gpurun_out/bench_n1_v11b.err
