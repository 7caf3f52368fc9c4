"""Host-side model of the activation split used by the 8-bit delta path of fwd_umma_kernel (bd_umma.cu, row-scale pass +
xperm_job): an activation row is multiplied by 2^(14 - emax) (emax = the row's largest exponent inside the CTA's K range) and
every element is written as three e5m2 pieces.  The model uses the kernel's integer arithmetic and torch's float8_e5m2
rounding (round-to-nearest-even, like cvt.rn.satfinite.e5m2x2.f32) over ALL 65536 bf16 bit patterns: the split must be exact
for every element within 2^-23 of the row maximum and lose at most 2^-31 of the row maximum below that -- whatever the
magnitude of the row.  This pins the claim DESIGN.md makes about the path; the GPU tests check the kernel itself."""
import torch

TOP = 14


def scale_field(emax: int) -> int:
    return min(max(254 + TOP - emax, 1), 253)


def split(bits: torch.Tensor, emax: int):
    """bits: int32 tensor of bf16 patterns -> (pieces[3] as fp32 of the scaled value, descale as python float)"""
    sf = scale_field(emax)
    x = (bits << 16).to(torch.int32).view(torch.float32)
    scale = torch.tensor(sf << 23, dtype=torch.int32).view(torch.float32)
    f = x * scale
    pieces = []
    for _ in range(3):
        p = f.to(torch.float8_e5m2).to(torch.float32)
        pieces.append(p)
        f = f - p
    descale = torch.tensor((254 - sf) << 23, dtype=torch.int32).view(torch.float32).item()
    return pieces, descale


def _finite_patterns():
    bits = torch.arange(0, 1 << 16, dtype=torch.int32)
    eb = (bits >> 7) & 0xFF
    return bits[eb != 255], eb[eb != 255]


def test_scale_and_descale_are_reciprocal_normals():
    for emax in range(0, 255):
        sf = scale_field(emax)
        assert 1 <= sf <= 253 and 1 <= 254 - sf <= 253
        a = torch.tensor(sf << 23, dtype=torch.int32).view(torch.float32).double()
        b = torch.tensor((254 - sf) << 23, dtype=torch.int32).view(torch.float32).double()
        assert (a * b).item() == 1.0


def test_split_is_exact_near_the_row_maximum_for_every_row_scale():
    bits, eb = _finite_patterns()
    x = (bits << 16).to(torch.int32).view(torch.float32).double()
    for emax in (1, 14, 15, 40, 100, 127, 140, 200, 254):
        sel = eb <= emax                      # a row whose largest exponent is emax holds only such elements
        pieces, descale = split(bits[sel], emax)
        recon = (pieces[0].double() + pieces[1].double() + pieces[2].double()) * descale
        xs, es = x[sel], eb[sel]
        for p in pieces:                      # no saturation: the first piece stays below e5m2's largest finite value
            assert torch.isfinite(p).all() and p.abs().max() <= 57344.0
        if 15 <= emax:                        # (rows below 2^-112 keep less headroom: the scale is clamped)
            near = es >= max(emax - 23, 1)    # normal numbers within 2^-23 of the row maximum: exact
            assert torch.equal(recon[near], xs[near])
            rowmax = 2.0 ** (emax - 127)
            assert (recon - xs).abs().max() <= rowmax * 2.0 ** -31
        else:
            assert (recon - xs).abs().max() <= 2.0 ** -126 * 2.0 ** -8
        assert recon[xs == 0].abs().max() == 0


def test_delta_sum_matches_fp64_on_wide_rows():
    """A decode-sized delta product with the split operands, accumulated like the kernel's three accumulator columns."""
    g = torch.Generator().manual_seed(0)
    K, N = 4096, 64
    signs = (torch.rand(K, N, generator=g) > 0.5).double() * 2 - 1
    for scale in (1e-30, 1e-12, 1e-6, 1e-3, 1.0, 1e3, 6e4, 1e9, 1e30):
        x = (torch.randn(K, generator=g) * scale)
        x[::97] *= 3000.0  # outlier channels
        x = x.bfloat16()
        bits = x.view(torch.int16).to(torch.int32) & 0xFFFF
        emax = int(((bits >> 7) & 0xFF).max())
        pieces, descale = split(bits, emax)
        acc = sum(p.double() @ signs for p in pieces) * descale
        exact = x.double() @ signs
        rowmax = x.double().abs().max()
        assert (acc - exact).abs().max() <= K * rowmax * 2.0 ** -31
        assert ((acc - exact).abs().mean() / exact.abs().mean()).item() < 1e-6
