#!/bin/bash
# post tests (colour jitter, io3), full gpu suite, reduce L2-prefetch timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_post.py -x -q 2>&1 | tail -15
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_post.py 2>&1 | tail -4
timeout 600 python tools/quick_bench.py --cfg 2 --algos binned 2>&1 | grep -v generated
timeout 600 python tools/quick_bench.py --cfg 4 --algos binned --orders scan 2>&1 | grep -v generated
timeout 600 python tools/quick_bench.py --cfg 5 2>&1 | grep -v generated
