#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/hint_sweep.py > gpurun_out/hint_sweep.txt 2>&1
cat gpurun_out/hint_sweep.txt
