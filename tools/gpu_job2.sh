#!/bin/bash
# bench line + smoke + ncu full captures of bin_points / reduce_tiles
set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python bench.py --order shuffled --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_shuffled.json 2>> gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bin_points -s 2 -c 1 -o gpurun_out/prof_bin \
    python tools/quick_bench.py --cfg 2 --orders scan --algos binned --reps 1 > gpurun_out/ncu_bin.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:reduce_tiles -s 2 -c 1 -o gpurun_out/prof_red \
    python tools/quick_bench.py --cfg 2 --orders scan --algos binned --reps 1 > gpurun_out/ncu_red.txt 2>&1
cat gpurun_out/smoke.txt gpurun_out/bench_n1.json gpurun_out/bench_n1_shuffled.json gpurun_out/bench_ref.json
tail -3 gpurun_out/bench_n1.err gpurun_out/bench_ref.err
ls -la gpurun_out
