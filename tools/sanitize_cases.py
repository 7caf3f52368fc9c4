#!/usr/bin/env python
"""Small launches of every forward-kernel variant for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):

    compute-sanitizer --tool memcheck python tools/sanitize_cases.py        (tools/sanitize.sh runs all tools)

Cases: tcgen05 decode launch on the 8-bit path (stream-K with split-K fix-up), the 16-bit path (fp16, and bf16 with several
rows per tenant), a grouped q/k/v launch, a prefill-size launch with row chunks, the delta-only product, the fp32 partial-sum
output, the SIMT kernel on a ragged shape, the codec / compress / fold kernels and the per-tenant leaves.  Every result is
checked against a float64 reference so that a sanitizer-clean run is also a correct one."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bitdelta_b200 as bd
from bitdelta_b200.diff import _fused_forward, _fused_forward_grouped

dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)


def problem(T, m, K, N, dtype=torch.bfloat16):
    w = (torch.randn(N, K, generator=gen, device=dev) * 0.02).to(dtype)
    masks = torch.randint(-(2**31), 2**31 - 1, (T, K // 32, N), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
    coeff = (torch.rand(T, generator=gen, device=dev) * 0.003 + 0.0005).to(dtype)
    x = torch.randn(T, m, K, generator=gen, device=dev).to(dtype)
    return x, w, masks, coeff


def exact(x, w, masks, coeff):
    s = bd.unpack(masks).double() * 2 - 1
    return x.double() @ w.double().T + coeff.double()[:, None, None] * torch.bmm(x.double(), s)


def check(name, y, ref, tol=2e-3):
    torch.cuda.synchronize()
    rel = ((y.double() - ref).abs().mean() / ref.abs().mean()).item()
    print(f"{name}: mean-rel {rel:.2e}", flush=True)
    assert rel < tol, name


cases = [
    ("umma decode 8-bit path, split-K (T=6 m=1 K=512 N=384)", (6, 1, 512, 384, torch.bfloat16), "umma"),
    ("umma decode 8-bit path, many tiles (T=3 m=1 K=256 N=2048)", (3, 1, 256, 2048, torch.bfloat16), "umma"),
    ("umma 16-bit path fp16 (T=4 m=1 K=512 N=256)", (4, 1, 512, 256, torch.float16), "umma"),
    ("umma 16-bit path bf16 rows (T=3 m=5 K=256 N=384)", (3, 5, 256, 384, torch.bfloat16), "umma"),
    ("umma prefill row chunks (T=1 m=300 K=256 N=256)", (1, 300, 256, 256, torch.bfloat16), "umma"),
    ("simt ragged (T=2 m=3 K=96 N=50)", (2, 3, 96, 50, torch.bfloat16), "simt"),
]
for name, (T, m, K, N, dt), kern in cases:
    x, w, masks, coeff = problem(T, m, K, N, dt)
    for _ in range(2):  # twice: the second launch runs on the workspace the first one left behind
        y = _fused_forward(x, w, masks, coeff, T, kern, static_operands=True)
    check(name, y, exact(x, w, masks, coeff))

x, w, masks, coeff = problem(6, 1, 512, 384)
y32 = _fused_forward(x, w, masks, coeff, 6, "umma", out_fp32=True)
check("umma fp32 partial-sum output", y32, exact(x, w, masks, coeff), 1e-5)
c = bd.binary_bmm(x, masks, kernel="umma")
check("umma delta only (binary_bmm)", c, torch.bmm(x.double(), bd.unpack(masks).double() * 2 - 1))

x = torch.randn(4, 1, 256, generator=gen, device=dev).bfloat16()
segs = [problem(4, 1, 256, n)[1:] for n in (256, 128, 128)]
ys = _fused_forward_grouped(x, [s[0] for s in segs], [s[1] for s in segs], [s[2] for s in segs], 4, "umma", static_operands=True)
for i, (w, masks, coeff) in enumerate(segs):
    check(f"umma grouped q/k/v segment {i}", ys[i], exact(x, w, masks, coeff))

bits = torch.rand(3, 256, 200, generator=gen, device=dev) > 0.5
assert torch.equal(bd.unpack(bd.pack(bits)), bits)
base = (torch.randn(128, 256, generator=gen, device=dev) * 0.02).bfloat16()
fine = (base.float() + 0.002 * torch.randn(128, 256, generator=gen, device=dev)).bfloat16()
mod = bd.BinaryDiff(base, fine)
with torch.no_grad():
    y = mod(torch.randn(1, 7, 256, generator=gen, device=dev).bfloat16())
torch.cuda.synchronize()
heads = [(torch.randn(v, 256, generator=gen, device=dev) * 0.02).bfloat16() for v in (300, 302)]
lm = bd.DataParallelModule(torch.nn.Linear(256, 300, bias=False, device=dev, dtype=torch.bfloat16), heads)
out = lm(torch.randn(2, 1, 256, generator=gen, device=dev).bfloat16())
torch.cuda.synchronize()
print("all cases ran", flush=True)
