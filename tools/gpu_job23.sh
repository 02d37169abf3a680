#!/bin/bash
# N=2 after moving the exchange to the side stream + chunk regions; strip stage split
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_strips_nccl.py -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
cat gpurun_out/bench_n2.json | cut -c1-400; tail -n 3 gpurun_out/bench_n2.err
timeout 600 python tools/strip_step_profile.py 2>&1 | tail -4
