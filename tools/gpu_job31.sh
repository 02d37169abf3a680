#!/bin/bash
# final N=1 checkpoint of the v11 build: gpu suite (full-size cases last), smoke, bench both arms
mkdir -p gpurun_out
LM_SKIP_FULLSIZE=1 timeout 170 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1_v11.json 2> gpurun_out/bench_n1_v11.err
cut -c1-1800 gpurun_out/bench_n1_v11.json; tail -n 2 gpurun_out/bench_n1_v11.err
timeout 60 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_v11.json 2> gpurun_out/bench_ref_v11.err
cut -c1-300 gpurun_out/bench_ref_v11.json
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 120 python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -1
