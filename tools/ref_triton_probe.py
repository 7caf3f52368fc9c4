#!/usr/bin/env python
"""Measures the REFERENCE's own GPU path (Triton binary_bmm + cuBLAS + pointwise, demo_backend.py:93-98) next to ours on
the same B200, for the Mistral-7B linear shapes with 6 tenants.  Comparison tool only:
it imports the unmodified reference kernel file from baseline/_ref (pip-installed from /root/reference, git-ignored).
"""
import importlib.util
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bitdelta_b200 as bd  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_binary_gemm_kernel", os.path.join(ROOT, "baseline/_ref/bitdelta/binary_gemm_kernel.py"))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

dev = torch.device("cuda:0")
T = 6
SHAPES = [("q/o_proj", 4096, 4096), ("k/v_proj", 1024, 4096), ("gate/up_proj", 14336, 4096), ("down_proj", 4096, 14336)]


def bench(fn, n_sets, iters=30):
    for i in range(5):
        fn(i % n_sets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % n_sets)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


out = []
for dtype in (torch.bfloat16, torch.float16):
    for name, N, K in SHAPES:
        bytes_alg = 2 * N * K + T * N * K // 8
        n_sets = max(2, int(400e6 // bytes_alg) + 1)  # rotate through > 3x L2 worth of operands
        gen = torch.Generator(device=dev).manual_seed(0)
        ws = [(torch.randn(N, K, generator=gen, device=dev) * 0.02).to(dtype) for _ in range(n_sets)]
        ms = [torch.randint(-(2**31), 2**31 - 1, (T, K // 32, N), generator=gen, device=dev, dtype=torch.int64).to(torch.int32) for _ in range(n_sets)]
        coeff = torch.full((T,), 0.002, device=dev, dtype=dtype)
        x = torch.randn(T, 1, K, generator=gen, device=dev).to(dtype)
        lins = []
        for w in ws:
            lin = torch.nn.Linear(K, N, bias=False, device=dev, dtype=dtype)
            lin.weight.data = w
            lins.append(lin)
        ours = [bd.DiffCompressModule(lins[i], ms[i], coeff) for i in range(n_sets)]

        def f_ref(i):
            return lins[i](x) + ref.binary_bmm(x, ms[i]) * coeff[:, None, None]

        def f_ours(i):
            return ours[i](x)

        def f_cublas(i):
            return lins[i](x)

        rec = {"dtype": str(dtype), "shape": name, "N": N, "K": K}
        try:
            y_ref = f_ref(0)
            y_ours = f_ours(0)
            rec["rel_ours_vs_ref"] = ((y_ours.float() - y_ref.float()).abs().mean() / y_ref.float().abs().mean()).item()
            rec["ref_us"] = bench(f_ref, n_sets)
            rec["ref_triton_only_us"] = bench(lambda i: ref.binary_bmm(x, ms[i]), n_sets)
        except Exception as e:  # the 2.0-era kernel may not compile under triton 3.6
            rec["ref_error"] = repr(e)[:300]
        rec["cublas_base_only_us"] = bench(f_cublas, n_sets)
        rec["ours_us"] = bench(f_ours, n_sets)
        rec["ours_GBps"] = bytes_alg / rec["ours_us"] / 1e3
        if "ref_us" in rec:
            rec["ref_GBps"] = bytes_alg / rec["ref_us"] / 1e3
            rec["speedup"] = rec["ref_us"] / rec["ours_us"]
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del ws, ms, lins, ours
        torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ref_triton_probe.json"), "w"), indent=1)
