#!/usr/bin/env python
"""Turns an ncu raw CSV of one decoder layer's 7 fused launches (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum)
into profiles/traffic.json.  usage: ncu_traffic.py <raw.csv> <out.json>"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
start = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[start]
recs = {}
for r in rows[start + 1:]:
    if len(r) < len(h):
        continue
    d = dict(zip(h, r))
    k = int(d["ID"])
    recs.setdefault(k, {})[d["Metric Name"]] = (float(d["Metric Value"]), d["Metric Unit"])
names = ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1}
out = {"launches": []}
tot = 0
for i, k in enumerate(sorted(recs)[:7]):
    m = recs[k]
    rd = m["dram__bytes_read.sum"][0] * scale[m["dram__bytes_read.sum"][1]]
    wr = m["dram__bytes_write.sum"][0] * scale[m["dram__bytes_write.sum"][1]]
    t = m["gpu__time_duration.sum"][0] * scale.get(m["gpu__time_duration.sum"][1], 1)
    out["launches"].append({"name": names[i], "dram_read_bytes": rd, "dram_write_bytes": wr, "duration_ns_under_ncu": t})
    tot += rd + wr
out["dram_bytes_per_launch_avg"] = tot / max(len(out["launches"]), 1)
out["note"] = "one decoder layer (7 launches) of bench.py --layers 1 under ncu, T=6 tenants; cold-cache, serialised"
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out)[:400])
