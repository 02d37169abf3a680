#!/bin/bash
# One parametrised GPU job runner (replaces the round-1 one-shot scripts).  Usage, from the repo root on a GPU box:
#   gpurun --timeout 900 -- 'bash tools/gpu_job.sh <job> [args...]'
# Everything a job wants kept goes to gpurun_out/ (merged back by gpurun); copy what is evidence into profiles/.
set -u
mkdir -p gpurun_out
JOB=${1:-help}; shift || true
case "$JOB" in
  tests)        # GPU suite; LM_SKIP_FULLSIZE=1 skips the 1e8-point parity runs
    timeout ${T:-900} python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -15 ;;
  bench)        # bench line -> gpurun_out/bench_<tag>.json ; args: tag [bench.py args...]
    TAG=${1:-x}; shift || true
    timeout ${T:-600} python bench.py "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
    cut -c1-2500 gpurun_out/bench_$TAG.json; tail -n 3 gpurun_out/bench_$TAG.err ;;
  l2ring)       # tools/l2ring: does a recycled record ring stay in L2 under the point stream? (timing + DRAM bytes)
    [ -x tools/l2ring ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/l2ring tools/l2ring.cu
    for R in 8 16 32 64; do for P in 0 1 2 3; do ./tools/l2ring $R 0.5 $P | grep "^ring"; done; done | tee gpurun_out/l2ring_time.txt
    ./tools/l2ring 32 0.5 0 | grep -v "^ring" | tee -a gpurun_out/l2ring_time.txt
    [ "${NCU:-1}" = 1 ] || exit 0
    for R in 16 32 64; do for P in 0 1 2; do
      timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
          -k regex:stream_ring -s 2 -c 1 --csv ./tools/l2ring $R 0.5 $P 2>/dev/null | grep -E "stream_ring" | \
          awk -F'","' -v r=$R -v p=$P '{print "ring_MB=" r, "policy=" p, $(NF-2), $(NF-1), $NF}'
    done; done | tee gpurun_out/l2ring_dram.txt ;;
  ncu)          # full capture of one kernel; args: kernel-regex tag [quick_bench args...]
    K=$1; TAG=$2; shift 2
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o gpurun_out/prof_$TAG \
        python tools/quick_bench.py "$@" > gpurun_out/ncu_$TAG.txt 2>&1
    ls -la gpurun_out/prof_$TAG.ncu-rep ;;
  launches)     # per-launch durations of a bench run; args: tag [bench.py args...]
    TAG=${1:-x}; shift || true
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 "$@" > gpurun_out/launches_$TAG.log 2>&1
    tail -n 40 gpurun_out/launches_$TAG.csv | cut -c1-220 ;;
  *) echo "jobs: tests | bench | l2ring | ncu | launches"; exit 2 ;;
esac
