#!/bin/bash
# 2 GPUs, v11 + mosaic gathered on rank 0: NCCL parity test (incl. rooted gather) + bench --gpus 2
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_strips_nccl.py -x -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_v11_root.json 2> gpurun_out/bench_n2_v11_root.err
cat gpurun_out/bench_n2_v11_root.json; tail -n 3 gpurun_out/bench_n2_v11_root.err
