#!/bin/bash
# N=1: full-size parity tests (new), one strip rank replayed on one GPU (stage split), other configs' timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -5
timeout 400 python tools/strip_step_profile.py > gpurun_out/strip_prof_v10.txt 2>&1; tail -5 gpurun_out/strip_prof_v10.txt
( timeout 300 python tools/quick_bench.py --cfg 1 --orders scan --algos binned
  timeout 300 python tools/quick_bench.py --cfg 2 --orders shuffled --algos binned
  timeout 300 python tools/quick_bench.py --cfg 4 --orders scan --algos binned
  timeout 300 python tools/quick_bench.py --cfg 5 ) > gpurun_out/quick_bench_v10.txt 2>&1
cat gpurun_out/quick_bench_v10.txt
