#!/bin/bash
# per-CTA chunk regions (8 B of bin state per tile): full gpu suite, cfg 2/4 timings, single-rank strip step
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python tools/quick_bench.py --cfg 2 --algos binned 2>&1 | grep -v generated
timeout 600 python tools/quick_bench.py --cfg 4 --algos binned --orders scan 2>&1 | grep -v generated
timeout 600 python tools/quick_bench.py --cfg 1 --algos binned --orders scan 2>&1 | grep -v generated
timeout 600 python tools/strip_step_profile.py 2>&1 | tail -2
