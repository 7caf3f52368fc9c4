#!/bin/bash
# compute-sanitizer over small launches of every forward-kernel variant (tools/sanitize_cases.py); run on the GPU box:
#   tools/sanitize.sh [out_dir]         -> out_dir/sanitizer_{memcheck,racecheck,synccheck,initcheck}.txt
out=${1:-gpurun_out}
mkdir -p "$out"
for tool in memcheck synccheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py > "$out/sanitizer_$tool.txt" 2>&1
  echo "== $tool: exit $? =="; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|all cases ran|Error|error" "$out/sanitizer_$tool.txt" | head -8
done
