#!/usr/bin/env python
"""Bring-up aid: compares the tcgen05 kernel with the SIMT kernel and an fp64 torch reference on a list of shapes."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("BD_BRINGUP_LIB", "1")  # the library with the trace and the A/B knobs (build.py --bringup)
import torch
import bitdelta_b200 as bd
from bitdelta_b200.diff import _fused_forward
from bitdelta_b200.binary_gemm_kernel import _bmm_impl

dev = torch.device("cuda:0")
cases = [  # T, m, K, N, has_base
    (1, 1, 64, 128, True), (1, 1, 64, 128, False), (1, 1, 128, 128, True), (1, 3, 256, 256, True), (2, 1, 128, 128, True),
    (6, 1, 4096, 4096, True), (6, 1, 4096, 1024, True), (6, 1, 14336, 4096, True), (6, 1, 4096, 14336, True),
    (6, 2, 512, 384, True), (1, 16, 4096, 4096, True), (1, 128, 1024, 512, True), (4, 16, 256, 256, False), (1, 5, 96, 200, True),
    (1, 300, 1024, 512, True), (1, 2048, 4096, 4096, True), (1, 1000, 512, 1280, False), (2, 200, 256, 256, True),
]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
for dtype in (torch.bfloat16, torch.float16):
    for (T, m, K, N, has_base) in cases:
        g = torch.Generator(device=dev).manual_seed(T * 1000 + m * 100 + K + N)
        x = torch.randn(T, m, K, generator=g, device=dev).to(dtype)
        w = (torch.randn(N, K, generator=g, device=dev) * 0.02).to(dtype)
        masks = torch.randint(-(2**31), 2**31 - 1, (T, K // 32, N), generator=g, device=dev, dtype=torch.int64).to(torch.int32)
        coeff = torch.rand(T, generator=g, device=dev) * 0.003 + 0.001
        signs = bd.unpack(masks).double() * 2 - 1
        delta = torch.bmm(x.double(), signs)
        exact = (x.double() @ w.double().T + coeff.double()[:, None, None] * delta) if has_base else delta
        rec = {"dtype": str(dtype)[6:], "T": T, "m": m, "K": K, "N": N, "base": has_base}
        for kern in ("simt", "umma"):
            try:
                if has_base:
                    y = _fused_forward(x, w, masks, coeff, T, kern)
                else:
                    y = _bmm_impl(x, masks, masks.shape[1] * masks.shape[2], kern)
                torch.cuda.synchronize()
                err = (y.double() - exact).abs()
                rec[kern + "_rel"] = (err.mean() / exact.abs().mean()).item()
                rec[kern + "_max"] = err.max().item()
                if kern == "umma" and rec[kern + "_rel"] > 3e-3:
                    # localise: which rows / columns are wrong
                    bad = (err > 0.02 * exact.abs().mean() + 0.02 * exact.abs())
                    rec["bad_frac"] = bad.double().mean().item()
                    rec["bad_rows"] = bad.any(-1).flatten().nonzero().flatten()[:8].tolist()
                    cols = bad.any(0).any(0).nonzero().flatten()
                    rec["bad_cols_head"] = cols[:8].tolist(); rec["bad_cols_n"] = cols.numel()
                    if has_base:
                        base_only = x.double() @ w.double().T
                        rec["vs_base_only_rel"] = ((y.double() - base_only).abs().mean() / exact.abs().mean()).item()
                        rec["vs_delta_only_rel"] = ((y.double() - coeff.double()[:, None, None] * delta).abs().mean() / exact.abs().mean()).item()
            except Exception as e:
                rec[kern + "_error"] = repr(e)[:200]
                if "CUDA" in repr(e) or "cuda" in repr(e):
                    print(json.dumps(rec), flush=True)
                    sys.exit(1)
        print(json.dumps(rec), flush=True)
