#!/usr/bin/env python
"""Turns an ncu raw CSV of one decoder layer's 7 fused launches (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum)
into profiles/traffic.json.  usage: ncu_traffic.py <raw.csv> <out.json>"""
import csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
start = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[start]
recs = {}
WANT = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")
if "Metric Name" in h:  # long format (default --csv): one row per (launch, metric)
    for r in rows[start + 1:]:
        if len(r) < len(h) or not r[0].strip().isdigit():
            continue
        d = dict(zip(h, r))
        recs.setdefault(int(d["ID"]), {})[d["Metric Name"]] = (float(d["Metric Value"].replace(",", "")), d["Metric Unit"])
else:  # wide format (--page raw): one row per launch, a units row under the header
    units = rows[start + 1]
    col = {n: i for i, n in enumerate(h)}
    for r in rows[start + 2:]:
        if len(r) < len(h) or not r[0].strip().isdigit():
            continue
        recs[int(r[0])] = {m: (float(r[col[m]].replace(",", "")), units[col[m]]) for m in WANT}
names7 = ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"]
names4 = ["q+k+v_proj (grouped)", "o_proj", "gate+up_proj (grouped)", "down_proj"]
names = names4 if len(recs) <= 4 else names7
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1}
out = {"launches": []}
tot = 0
for i, k in enumerate(sorted(recs)[:len(names)]):
    m = recs[k]
    rd = m["dram__bytes_read.sum"][0] * scale[m["dram__bytes_read.sum"][1]]
    wr = m["dram__bytes_write.sum"][0] * scale[m["dram__bytes_write.sum"][1]]
    t = m["gpu__time_duration.sum"][0] * scale.get(m["gpu__time_duration.sum"][1], 1)
    out["launches"].append({"name": names[i], "dram_read_bytes": rd, "dram_write_bytes": wr, "duration_ns_under_ncu": t})
    tot += rd + wr
out["dram_bytes_per_launch_avg"] = tot / max(len(out["launches"]), 1)
alg = {"q+k+v_proj (grouped)": 2*6144*4096 + 6*6144*4096//8, "o_proj": 2*4096*4096 + 6*4096*4096//8, "gate+up_proj (grouped)": 2*28672*4096 + 6*28672*4096//8, "down_proj": 2*4096*14336 + 6*4096*14336//8}
for l in out["launches"]:
    if l["name"] in alg:
        l["algorithmic_bytes"] = alg[l["name"]]
        l["traffic_over_algorithmic"] = (l["dram_read_bytes"] + l["dram_write_bytes"]) / alg[l["name"]]
out["note"] = "one decoder layer of bench.py --layers 1 under ncu (grouped launches), T=6 tenants; cold-cache, serialised"
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out)[:400])
