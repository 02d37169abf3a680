#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
cat gpurun_out/bench_n2.json; tail -n 3 gpurun_out/bench_n2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_strip.csv \
    python tools/strip_step_profile.py > gpurun_out/strip_prof.txt 2>&1
grep -E "Kernel|kernel" gpurun_out/launches_strip.csv | awk -F'","' '{print substr($5,1,60), $NF}' | tail -14
cat gpurun_out/strip_prof.txt | tail -5
