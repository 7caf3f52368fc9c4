#!/bin/bash
# A/B runs of the tcgen05 decode kernel with the bring-up knobs (bring-up library only): which stage bounds a launch?
# usage: tools/ab_flags.sh "T,m,K,N ..." flag flag ...
shapes=${1:-"6,1,4096,14336 1,1,4096,14336"}; shift
for f in ${@:-0 16 32 48 64 112 1}; do
  BD_BRINGUP_LIB=1 BD_DBG_FLAGS=$f timeout 120 python tools/kernel_bench.py umma $shapes 2>&1 | grep '"us"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('flags', d['flags'].rjust(3), 'T', d['T'], 'K', d['K'], 'N', d['N'], 'us', d['us'], 'frac', d['frac_of_6574'])"
done
