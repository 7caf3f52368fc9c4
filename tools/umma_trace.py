#!/usr/bin/env python
"""Prints the per-unit timeline (clock64 deltas) of CTA 0 of the tcgen05 kernel for one launch.  usage: umma_trace.py T m K N"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("BD_BRINGUP_LIB", "1")  # the library with the trace and the A/B knobs (build.py --bringup)
import torch
import bitdelta_b200 as bd
from bitdelta_b200 import _lib
from bitdelta_b200.diff import _fused_forward
T, m, K, N = [int(v) for v in sys.argv[1:5]] if len(sys.argv) >= 5 else (6, 1, 4096, 14336)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
w = (torch.randn(N, K, generator=g, device=dev) * 0.02).bfloat16()
ms = torch.randint(-(2**31), 2**31 - 1, (T, K // 32, N), generator=g, device=dev, dtype=torch.int64).to(torch.int32)
coeff = torch.full((T,), 0.002, device=dev)
x = torch.randn(T, m, K, generator=g, device=dev).bfloat16()
for _ in range(3):
    _fused_forward(x, w, ms, coeff, T, "umma", static_operands=True)
_lib.lib.bd_debug_set_flags(int(os.environ.get("BD_DBG_FLAGS", "0")), 0)  # A/B knobs of the bring-up library (see bd_umma.cu)
buf = torch.zeros(64 * 16 + 4 * 160, dtype=torch.int64, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_lib.lib.bd_debug_set_trace(buf.data_ptr())
_fused_forward(x, w, ms, coeff, T, "umma", static_operands=True)
ev0.record()
_fused_forward(x, w, ms, coeff, T, "umma", static_operands=True)
ev1.record()
torch.cuda.synchronize()
print(f"event time of one traced launch: {ev0.elapsed_time(ev1)*1e3:.1f} us")
_lib.lib.bd_debug_set_trace(None)
allt = buf.cpu()
life = allt[1024:].view(160, 4)
live = life[life[:, 0] > 0]
e0 = live[:, 0].min().item()
ent = (live[:, 0] - e0).float() / 1e3
ex = (live[:, 1] - e0).float() / 1e3
print(f"grid: {live.shape[0]} CTAs; entry us min/median/max = {ent.min():.2f}/{ent.median():.2f}/{ent.max():.2f}; exit us min/median/max = {ex.min():.2f}/{ex.median():.2f}/{ex.max():.2f}; lifetime median {(ex-ent).median():.2f} max {(ex-ent).max():.2f}")
if os.environ.get("BD_TRACE_CTAS"):  # per-CTA lifetimes: which CTAs end late, and on which SMs
    lt_ = ((life[:, 1] - life[:, 0]).float() / 1e3).tolist()
    order = sorted(range(live.shape[0]), key=lambda c: lt_[c])
    tail_ = ((life[:, 1] - life[:, 3]).float() / 1e3).tolist()  # last unit handed over -> exit (drain, epilogue, fix-up)
    print("CTA lifetime/tail (us) sorted by lifetime: " + " ".join(f"{c}@sm{int(life[c, 2])}:{lt_[c]:.1f}/{tail_[c]:.1f}" for c in order))
t = allt[:1024].view(64, 16)
t0 = t[0, 8].item()
k = t[63]
print(f"kernel (CTA 0): entry->setup {k[1]-k[0]} cyc, setup->first TMA {t0-k[1]}, first TMA->last unit done {k[2]-t0}, last epilogue {k[3]-k[2]}, ->teardown {k[4]-k[3]}; total {k[4]-k[0]} cyc = {(k[9]-k[8])/1e3:.2f} us (globaltimer)")
names = ["U:start", "U:released", "M:synced", "U:unpk", "U:fenced", "M:top", "M:fenced", "M:commit", "P:empty", "M:issued", "X:done", "U:arrived", "M:conv", "U:end", "S:full", "S:release"]
rows = [[t[it, s_].item() for s_ in range(14)] for it in range(8, 56) if t[it, 1].item() and t[it + 2, 0].item()]
if rows:
    import statistics as st_
    own = [b[0] - a[0] for a, b in zip(rows, rows[1:])]
    print(f"traced unpack warp, per own unit (median): wait+fence {st_.median(r[1]-r[0] for r in rows):.0f}, unpack {st_.median(r[3]-r[1] for r in rows):.0f}, "
          f"wait::st+fence {st_.median(r[4]-r[3] for r in rows):.0f}, arrive {st_.median(r[11]-r[4] for r in rows):.0f}, bookkeeping {st_.median(r[13]-r[11] for r in rows):.0f}; own-unit period {st_.median(own):.0f}")
print(f"T={T} m={m} K={K} N={N}; cycles relative to the producer's first TMA issue")
print("unit " + " ".join(n.rjust(9) for n in names))
for it in range(64):
    if not any(t[it, s].item() for s in range(16)):
        break
    row = [(t[it, s].item() - t0) if t[it, s].item() else -1 for s in range(16)]
    print(f"{it:4d} " + " ".join(str(v).rjust(9) for v in row))
