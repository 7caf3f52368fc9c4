"""Host-side model of the activation split used by the 8-bit delta path of fwd_umma_kernel (bd_umma.cu, xperm_job):
every bf16 activation goes to one of five static exponent buckets and is written as three e5m2 pieces of x * 2^(51-24b).
The model uses the kernel's integer arithmetic and torch's float8_e5m2 rounding (round-to-nearest-even, like
cvt.rn.satfinite.e5m2x2.f32) over ALL 65536 bf16 bit patterns: the split must be exact on [2^-60, 2^60) and degrade
gradually below.  This pins the claim DESIGN.md makes about the path; the GPU tests check the kernel itself."""
import numpy as np
import torch

EXP_LO, BUCKETS = 127 - 60, 5


def split(bits: torch.Tensor):
    """bits: int32 tensor of bf16 patterns -> (bucket, pieces[3] as fp32 of the scaled value, bad mask)"""
    eb = (bits >> 7) & 0xFF
    b = torch.clamp((torch.clamp(eb - EXP_LO, min=0) * 171) >> 12, max=BUCKETS - 1)
    bad = eb >= EXP_LO + 24 * BUCKETS
    x = (bits << 16).to(torch.int32).view(torch.float32)
    scale = ((127 + 51 - 24 * b) << 23).to(torch.int32).view(torch.float32)
    f = x * scale
    pieces = []
    for _ in range(3):
        p = f.to(torch.float8_e5m2).to(torch.float32)
        pieces.append(p)
        f = f - p
    return b, pieces, bad


def test_bucket_index_is_floor_div_24():
    v = torch.arange(0, 256)
    assert torch.equal((v * 171) >> 12, v // 24)


def test_split_is_exact_for_every_bf16_in_range():
    bits = torch.arange(0, 1 << 16, dtype=torch.int32)
    b, pieces, bad = split(bits)
    x = (bits << 16).to(torch.int32).view(torch.float32).double()
    eb = (bits >> 7) & 0xFF
    recon = (pieces[0].double() + pieces[1].double() + pieces[2].double()) * torch.pow(torch.tensor(2.0, dtype=torch.float64), (24 * b - 51).double())
    in_range = (eb >= EXP_LO) & (eb < EXP_LO + 24 * BUCKETS)
    assert in_range.sum() == 2 * 120 * 128  # 120 binades, 128 mantissas, two signs
    assert torch.equal(recon[in_range], x[in_range])
    assert not bad[in_range].any()
    # every piece is a finite e5m2 value (no saturation inside the window)
    for p in pieces:
        assert torch.isfinite(p[in_range]).all() and p[in_range].abs().max() <= 32768.0
    # below the window: bucket 0, absolute error below 2^-76 (half of e5m2's smallest subnormal 2^-16, times 2^-51 ... rounded thrice)
    small = (eb < EXP_LO)
    assert (b[small] == 0).all()
    assert (recon[small] - x[small]).abs().max() <= 2.0 ** -67
    # zero stays zero, out-of-range / inf / nan are flagged
    assert recon[0] == 0 and recon[0x8000] == 0
    assert bad[eb >= EXP_LO + 24 * BUCKETS].all() and bad[0x7F80] and bad[0x7FC0]


def test_delta_sum_matches_fp64_on_wide_rows():
    """A decode-sized delta product with the split operands, accumulated per bucket like the kernel's accumulator columns."""
    g = torch.Generator().manual_seed(0)
    K, N = 4096, 64
    signs = (torch.rand(K, N, generator=g) > 0.5).double() * 2 - 1
    for scale in (1e-12, 1e-6, 1e-3, 1.0, 1e3, 6e4, 1e9):
        x = (torch.randn(K, generator=g) * scale)
        x[::97] *= 3000.0  # outlier channels
        x = x.bfloat16()
        bits = x.view(torch.int16).to(torch.int32) & 0xFFFF
        b, pieces, bad = split(bits)
        assert not bad.any()
        acc = torch.zeros(N, dtype=torch.float64)
        for bk in range(BUCKETS):
            sel = (b == bk)
            col = sum((p.double() * sel) @ signs for p in pieces)
            acc += col * 2.0 ** (24 * bk - 51)
        exact = x.double() @ signs
        assert torch.equal(acc, exact)
