#!/bin/bash
# One --set full capture of the grouped gate+up decode launch (release library) and its per-instruction source page as CSV:
# where every warp role spends its samples.  Output: gpurun_out/gate_up_full.ncu-rep, gate_up_source.csv, gate_up_full_raw.csv
mkdir -p gpurun_out
QUICK="--steps 1 --warmup 1 --no-cpu-baseline --no-gemm --no-extras --no-triton-ref --no-model-step --layers 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_umma -s 2 -c 1 -o gpurun_out/gate_up_full -f python bench.py $QUICK > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/gate_up_full.ncu-rep --page source --csv > gpurun_out/gate_up_source.csv 2>/dev/null
ncu -i gpurun_out/gate_up_full.ncu-rep --page raw --csv > gpurun_out/gate_up_full_raw.csv 2>/dev/null
wc -l gpurun_out/gate_up_source.csv
