#!/bin/bash
# 2 GPUs, v11 (equal strips, 3 bin CTAs/SM on wide rasters): NCCL parity test + bench --gpus 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_strips_nccl.py -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_v11.json 2> gpurun_out/bench_n2_v11.err
cat gpurun_out/bench_n2_v11.json; tail -n 3 gpurun_out/bench_n2_v11.err
