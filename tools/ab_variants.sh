#!/bin/bash
# Compile-time A/B of the RELEASE kernel: build variants here (no GPU needed), time them on the GPU box.
#   here:        tools/ab_variants.sh build name1:-DBD_X=1 name2:-DBD_Y=2,-DBD_Z   -> bitdelta_b200/variants/<name>.so
#   on the box:  tools/ab_variants.sh run "T,m,K,N ..."                            -> one line per variant and shape
# The run swaps each variant in for libbitdelta_b200.so inside the box's scratch copy of the repo and restores it.
set -e
cd "$(dirname "$0")/.."
V=bitdelta_b200/variants
if [ "$1" = build ]; then
  shift; mkdir -p $V
  for spec in "$@"; do
    name=${spec%%:*}; defs=${spec#*:}
    python bitdelta_b200/build.py --out=$PWD/$V/$name.so ${defs//,/ } | tail -1
  done
else
  shapes=${2:-"6,1,4096,14336 6,1,14336,4096 6,1,4096,4096 1,1,4096,14336"}
  cp bitdelta_b200/libbitdelta_b200.so /tmp/bd_release.so
  for so in /tmp/bd_release.so $V/*.so; do
    cp $so bitdelta_b200/libbitdelta_b200.so
    timeout 150 python tools/kernel_bench.py auto $shapes 2>&1 | grep '"us"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$(basename $so .so)'.ljust(14), 'T', d['T'], 'm', d['m'], 'K', d['K'], 'N', d['N'], 'us', d['us'], 'frac', d['frac_of_6574'])"
  done
  cp /tmp/bd_release.so bitdelta_b200/libbitdelta_b200.so
fi
