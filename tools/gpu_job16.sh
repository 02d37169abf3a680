#!/bin/bash
# baseline of the restored checkpoint: gpu tests, bench, quick_bench over configs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv
cat MEASURED_PEAKS.json 2>/dev/null
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_n1.json
timeout 600 python tools/quick_bench.py --cfg 2 2>&1 | grep -v generated
timeout 600 python tools/quick_bench.py --cfg 1 --algos binned 2>&1 | grep -v generated
timeout 600 python tools/quick_bench.py --cfg 4 --algos binned 2>&1 | grep -v generated
timeout 600 python tools/quick_bench.py --cfg 5 2>&1 | grep -v generated
