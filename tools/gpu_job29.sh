#!/bin/bash
# 4 GPUs: interior ranks (two neighbours each) over NCCL, full per-rank workload
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4_v11.json 2> gpurun_out/bench_n4_v11.err
cat gpurun_out/bench_n4_v11.json; tail -n 3 gpurun_out/bench_n4_v11.err
