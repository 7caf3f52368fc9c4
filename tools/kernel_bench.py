#!/usr/bin/env python
"""Kernel-only timing of the fused forward for a list of shapes: launches are captured in a CUDA graph over rotating
operand sets larger than L2, so the number is device time per launch without Python overhead.
usage: kernel_bench.py [kernel] [T,m,K,N ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("BD_DBG_FLAGS", "0") != "0":
    os.environ.setdefault("BD_BRINGUP_LIB", "1")  # A/B knobs exist only in the bring-up library (build.py --bringup); plain timings use the release library
import torch
import bitdelta_b200 as bd
from bitdelta_b200.diff import _fused_forward

dev = torch.device("cuda:0")
kernel = sys.argv[1] if len(sys.argv) > 1 else "auto"
from bitdelta_b200 import _lib
if _lib.BRINGUP:
    _lib.lib.bd_debug_set_flags(int(os.environ.get("BD_DBG_FLAGS", "0")), int(os.environ.get("BD_LOAD_GROUP", "1")))
shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]] or [
    (6, 1, 4096, 4096), (6, 1, 4096, 1024), (6, 1, 4096, 14336), (6, 1, 14336, 4096),
    (1, 1, 4096, 4096), (1, 1, 4096, 14336), (1, 16, 4096, 4096), (1, 128, 4096, 4096), (3, 1, 4096, 14336)]
out = []
for (T, m, K, N) in shapes:
    bytes_alg = 2 * N * K + T * N * K // 8 + 2 * T * m * (K + N)
    n_sets = int(os.environ.get("BD_NSETS", "0")) or max(2, int(600e6 // bytes_alg) + 1)  # BD_NSETS=1: operands stay in L2 (consumer-side floor)
    g = torch.Generator(device=dev).manual_seed(0)
    ws = [(torch.randn(N, K, generator=g, device=dev) * 0.02).bfloat16() for _ in range(n_sets)]
    ms = [torch.randint(-(2**31), 2**31 - 1, (T, K // 32, N), generator=g, device=dev, dtype=torch.int64).to(torch.int32) for _ in range(n_sets)]
    coeff = torch.full((T,), 0.002, device=dev)
    x = torch.randn(T, m, K, generator=g, device=dev).bfloat16()
    s = torch.cuda.Stream()
    reps = max(n_sets, 16)
    with torch.cuda.stream(s):
        for i in range(n_sets):
            _fused_forward(x, ws[i], ms[i], coeff, T, kernel, static_operands=True)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for i in range(reps):
                _fused_forward(x, ws[i % n_sets], ms[i % n_sets], coeff, T, kernel, static_operands=True)
        for _ in range(3):
            gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            gr.replay()
        e1.record(s)
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * reps)
    rec = {"kernel": kernel, "flags": os.environ.get("BD_DBG_FLAGS", "0"), "group": os.environ.get("BD_LOAD_GROUP", "default"), "T": T, "m": m, "K": K, "N": N, "us": round(us, 2), "GBps": round(bytes_alg / us / 1e3, 1),
           "frac_of_6574": round(bytes_alg / us / 1e3 / 6574.1, 3), "tflops": round(4 * T * m * N * K / us / 1e6, 2)}
    print(json.dumps(rec), flush=True)
    out.append(rec)
    del ws, ms
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/kernel_bench_{kernel}.json", "w"), indent=1)
