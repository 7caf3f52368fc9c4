#!/usr/bin/env python
"""Per-block sample distribution of an `ncu --page source --csv` dump of fwd_umma_kernel: which code (hence which warp role
and phase) the warps spend their time in.  usage: ncu_roles.py gate_up_source.csv [block]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
isrc, iex, ismp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
names = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
idx = [h.index(n) for n in names]
recs = [(n, r) for n, r in enumerate(rows[2:]) if len(r) >= len(h)]
tot = sum(int(r[ismp] or 0) for _, r in recs)
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 32
KEYS = ("UTMALDG", "UTCQMMA", "UTCHMMA", "UTCBAR", "STTM", "LDTM", "BAR.", "SYNCS", "F2FP", "LDG", "ATOM", "STG", "MEMBAR", "VIMNMX3", "LDS", "STS")
print("total samples", tot)
for i in range(0, len(recs), blk):
    c = recs[i:i + blk]
    s = sum(int(r[ismp] or 0) for _, r in c)
    if s < tot * 0.004:
        continue
    ex = max(int(r[iex] or 0) for _, r in c)
    agg = collections.Counter()
    for _, r in c:
        for nm, j in zip(names, idx):
            agg[nm[6:]] += int(r[j] or 0)
    key = sorted({k for _, r in c for k in KEYS if k in r[isrc]})
    top = ", ".join(f"{k}:{v}" for k, v in agg.most_common(4))
    print(f"{i:5d} {s:5d} ({100*s/tot:4.1f}%) maxexec {ex:8d}  [{top}]  {key}")
