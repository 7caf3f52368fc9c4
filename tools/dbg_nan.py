import torch, sys
sys.path.insert(0, "/root/repo")
import bitdelta_b200 as bd
from bitdelta_b200.diff import _fused_forward
dev = torch.device("cuda:0")
for (T, K, N) in [(8, 8192, 8192), (8, 8192, 1024), (8, 28672, 8192), (6, 4096, 4096), (8, 4096, 4096), (10, 4096, 4096), (7, 8192, 8192)]:
    gen = torch.Generator(device=dev).manual_seed(777)
    w = torch.empty(N, K, device=dev, dtype=torch.bfloat16).normal_(0.0, 0.02, generator=gen)
    masks = torch.randint(-(2**31), 2**31 - 1, (T, K // 32, N), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
    coeffs = (torch.rand(T, generator=gen, device=dev) * 0.002 + 0.001).to(torch.bfloat16)
    x = torch.randn(T, 1, K, generator=gen, device=dev).bfloat16()
    for i in range(3):
        y = _fused_forward(x, w, masks, coeffs, T, "auto", static_operands=True)
        torch.cuda.synchronize()
        bad = ~torch.isfinite(y)
        print(T, K, N, "run", i, "non-finite", int(bad.sum()), "per tenant", bad.sum(dim=(1, 2)).tolist(), flush=True)
    signs = bd.unpack(masks).double() * 2 - 1
    exact = x.double() @ w.double().T + coeffs.double()[:, None, None] * torch.bmm(x.double(), signs)
    print("   mean-rel", ((y.double() - exact).abs().mean() / exact.abs().mean()).item())
    if bad.any():
        idx = bad.nonzero()[:8].tolist()
        print("   first bad", idx)
