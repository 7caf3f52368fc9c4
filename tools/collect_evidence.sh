#!/bin/bash
# Round evidence on the GPU box: smoke, bench line, per-kernel timings, ncu launch list, per-launch DRAM traffic of one layer,
# one --set full capture, compute-sanitizer logs.  Outputs under gpurun_out/ (copied to profiles/rNN/ and summarised there).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 1300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 300 gpurun_out/bench_final.err
timeout 400 python tools/kernel_bench.py auto > gpurun_out/kernel_bench.log 2>&1
QUICK="--steps 1 --warmup 1 --no-cpu-baseline --no-gemm --no-extras --no-triton-ref --no-model-step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fwd_umma -c 256 --csv --log-file gpurun_out/launches.csv python bench.py $QUICK > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fwd_umma -c 4 --csv --page raw --log-file gpurun_out/traffic_raw.csv python bench.py $QUICK --layers 1 > gpurun_out/ncu_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_umma -s 2 -c 1 -o gpurun_out/gate_up_full -f python bench.py $QUICK --layers 1 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/gate_up_full.ncu-rep --page raw --csv > gpurun_out/gate_up_full_raw.csv 2>/dev/null
tools/sanitize.sh gpurun_out
ls -la gpurun_out | tail -20
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d.get("w1a16_gemm",{}).get("tflops"), d["clocks"], d["gpu_launches"], d["cpu_baseline"])
PY
