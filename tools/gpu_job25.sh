#!/bin/bash
# 2 GPUs: NCCL strip parity test, bench --gpus 2 (both arms), evidence for SURVEY 8(e)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_strips_nccl.py -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_v10.json 2> gpurun_out/bench_n2_v10.err
cat gpurun_out/bench_n2_v10.json; tail -n 5 gpurun_out/bench_n2_v10.err
