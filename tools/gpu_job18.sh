#!/bin/bash
# LAS front end + batch fix: tests, then LAS timings at cfg2 scale
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_las.py tests/test_gpu_batch.py -x -q 2>&1 | tail -15
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_batch.py --deselect tests/test_gpu_las.py 2>&1 | tail -4
timeout 600 python tools/quick_bench.py --cfg 2 --algos binned --orders scan --las 2>&1 | grep -v generated
