"""GPU parity tests: the CUDA path (through the Python surface -> ctypes -> C ABI) against the numpy oracle and the
golden vectors produced by the reference.

Tolerances (stated once, used everywhere):
  * integer / bit work (pack, unpack, compress masks, fold): bit-exact.
  * floating point: both products are accumulated in fp32 and the sum is rounded ONCE to the activation dtype, so against
    the float64 truth every element must satisfy  |y - exact| <= (h + 1e-3) * |exact| + 1e-3 * mean|exact|
    where h is the half-ulp relative rounding step of the output dtype (unit roundoff: 2^-8 for bf16, 2^-11 for fp16): 1e-3 relative for
    the arithmetic (north_star), plus the unavoidable output quantisation, plus an absolute floor for cancelled sums.
  * the notebook's own metric mean|y-ref| / mean|ref| (cells 22-24) must stay < 1e-3 for fp16 outputs and < 2e-3 for
    bf16 outputs (one bf16 rounding alone is ~1.3e-3 by that metric).
"""
import os

import numpy as np
import pytest
import torch

from oracle import bitdelta_oracle as O

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():  # keeps collection cheap on the CPU box; the marker deselects these anyway
    pytest.skip("needs a CUDA device", allow_module_level=True)

import bitdelta_b200 as bd  # noqa: E402

DEV = torch.device("cuda:0")
KERNELS = ["simt", "auto"]


def bf(bits):
    return O.bf16_bits_to_f32(bits)


def t_bf16(bits: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(bits.view(np.int16).copy()).view(torch.bfloat16)


def to_np(t: torch.Tensor) -> np.ndarray:
    return t.detach().float().cpu().numpy()


def assert_close_to_exact(y: torch.Tensor, exact: np.ndarray, what=""):
    h = 2.0**-8 if y.dtype == torch.bfloat16 else 2.0**-11
    got = to_np(y).astype(np.float64)
    err = np.abs(got - exact)
    bound = (h + 1e-3) * np.abs(exact) + 1e-3 * np.abs(exact).mean()
    bad = err > bound
    assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} elements out of tolerance, worst {err.max():.4g}"
    rel = O.rel_mean_abs_err(got, exact)
    assert rel < (2e-3 if y.dtype == torch.bfloat16 else 1e-3), f"{what}: mean-rel {rel:.3g}"


# ------------------------------------------------------------------------------------------------ codec
@pytest.mark.parametrize("n_bits", [8, 16, 32, 64])
def test_codec_matches_reference_vectors(golden, n_bits):
    g = golden("codec.npz")
    bits = torch.from_numpy(g["bits"]).to(DEV)
    packed = bd.pack(bits, n_bits)
    assert packed.device.type == "cuda" and np.array_equal(packed.cpu().numpy(), g[f"packed{n_bits}"])
    assert torch.equal(bd.unpack(packed, n_bits), bits)


def test_codec_known_answers_and_negative_words(golden):
    g = golden("codec.npz")
    assert np.array_equal(bd.pack(torch.from_numpy(g["kat_bits"]).to(DEV)).cpu().numpy(), g["kat_packed"])
    words = torch.from_numpy(g["words"]).to(DEV)
    assert np.array_equal(bd.unpack(words).cpu().numpy(), g["words_unpacked"])


def test_codec_full_size_round_trip():
    # Mistral-7B down_proj signs for 6 tenants: [6, 14336/32, 4096] words
    gen = torch.Generator(device=DEV).manual_seed(7)
    words = torch.randint(-(2**31), 2**31 - 1, (6, 448, 4096), generator=gen, device=DEV, dtype=torch.int64).to(torch.int32)
    bits = bd.unpack(words)
    assert bits.shape == (6, 14336, 4096) and bits.dtype == torch.bool
    assert torch.equal(bd.pack(bits), words)
    # popcount is preserved
    w64 = words.to(torch.int64) & 0xFFFFFFFF
    pop = sum(((w64 >> i) & 1).sum().item() for i in range(32))
    assert pop == bits.sum().item()


def test_codec_edge_cases():
    assert bd.pack(torch.zeros(0, 32, 5, dtype=torch.bool, device=DEV)).shape == (0, 1, 5)
    assert bd.unpack(torch.zeros(2, 0, 7, dtype=torch.int32, device=DEV)).shape == (2, 0, 7)
    with pytest.raises(AssertionError, match="K must be divisible by n_bits"):
        bd.pack(torch.zeros(33, 4, dtype=torch.bool, device=DEV))
    # non-contiguous input (a transposed view, as BinaryDiff.__init__ passes it)
    x = torch.rand(96, 64, device=DEV) > 0.5
    assert np.array_equal(bd.pack(x.T).cpu().numpy(), O.pack(x.T.cpu().numpy()))


# ------------------------------------------------------------------------------------------------ compress / fold
def test_compress_matches_reference_ctor(golden):
    g = golden("binarydiff_ctor.npz")
    m = bd.BinaryDiff(t_bf16(g["base"]).to(DEV), t_bf16(g["finetune"]).to(DEV))
    assert np.array_equal(m.mask.cpu().numpy(), g["mask"])
    assert abs(m.coeff.item() - float(g["coeff"])) <= 1e-6 * float(g["coeff"])
    assert list(m.state_dict().keys()) == list(g["state_keys"])
    assert m.base.shape == (96, 48) and m.base.stride() == (1, 96)
    assert m.coeff.dtype == torch.float32 and m.coeff.dim() == 0 and m.coeff.requires_grad


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_compress_full_size_properties(dtype):
    torch.manual_seed(0)
    N, K = 4096, 4096
    base = (torch.randn(N, K, device=DEV) * 0.02).to(dtype)
    fine = (base.float() + torch.randn(N, K, device=DEV) * 0.002).to(dtype)
    fine[5, :100] = base[5, :100]
    m = bd.BinaryDiff(base, fine)
    diff = fine - base
    want_bits = ~(diff < 0)
    assert torch.equal(bd.unpack(m.mask), want_bits.T)
    want_coeff = diff.double().abs().mean().item()
    assert abs(m.coeff.item() - want_coeff) <= 1e-6 * want_coeff
    # against the oracle on a slab
    mask_o, coeff_o = O.binarydiff_compress(to_np(base[:64]), to_np(fine[:64]))
    assert np.array_equal(m.mask[:, :64].cpu().numpy(), mask_o)


def test_fold_matches_load_diff(golden):
    g = golden("tiny_llama.npz")
    d = torch.load(os.path.join(os.path.dirname(__file__), "golden", "tiny_llama_diff.pt"), weights_only=False)
    for name in ["model.layers.0.self_attn.q_proj", "model.layers.1.mlp.down_proj", "model.layers.1.self_attn.k_proj"]:
        w = t_bf16(g["base::" + name]).to(DEV).clone()
        bd.fold_into(w, d[name + ".mask"].to(DEV), d[name + ".coeff"])
        got = w.cpu().view(torch.int16).numpy().view(np.uint16)
        assert np.array_equal(got, g["folded::" + name]), name


# ------------------------------------------------------------------------------------------------ binary_matmul / binary_bmm
@pytest.mark.parametrize("kernel", KERNELS)
def test_notebook_cell7_binary_matmul(kernel):
    # notebook cell 7: seed 0, a = randn(512,512) fp16, b = randn(512,512) > 0.5, allclose(atol=1e-2, rtol=0)
    torch.manual_seed(0)
    a = torch.randn((512, 512), device=DEV, dtype=torch.float16)
    b = torch.randn((512, 512), device=DEV) > 0.5
    c = bd.binary_matmul(a, bd.pack(b), kernel=kernel)
    ref = torch.matmul(a.float(), b.float() * 2 - 1)
    # fp16 output spacing at |c| ~ 64 is 0.03-0.06, so the notebook's atol=1e-2 only holds between two fp16 results;
    # against the fp32 product we allow half an fp16 ulp on top of it.
    assert torch.allclose(c.float(), ref, atol=1e-2 + 0.5 * 2.0**-4, rtol=0)
    assert_close_to_exact(c, O.binary_matmul_exact(to_np(a), bd.pack(b).cpu().numpy()), "cell7")


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_notebook_cell13_binary_bmm(kernel, dtype):
    # notebook cells 13-14: a = randn(16,128,512), b = randn(16,512,1024) > 0.5, allclose(rtol=1e-2); recorded max rel 0.0062
    torch.manual_seed(0)
    a = torch.randn((16, 128, 512), device=DEV, dtype=dtype)
    b = torch.randn((16, 512, 1024), device=DEV) > 0.5
    words = bd.pack(b)
    c = bd.binary_bmm(a, words, kernel=kernel)
    assert c.shape == (16, 128, 1024) and c.dtype == dtype
    assert_close_to_exact(c, O.binary_bmm_exact(to_np(a), words.cpu().numpy()), "cell13")


def test_binary_bmm_matches_triton_interpreter_vectors(golden):
    g = golden("triton_interp.npz")
    a3 = torch.from_numpy(g["a3"]).to(DEV)
    c3 = bd.binary_bmm(a3, torch.from_numpy(g["b3"]).to(DEV))
    # the reference kernel rounds fp32 -> fp16 once for fp16 inputs, as we do: results agree to the last fp16 bit up to
    # fp32 summation order
    diff = (c3.float().cpu() - torch.from_numpy(g["c3"]).float()).abs()
    assert (diff <= 2.0**-10 * torch.from_numpy(g["c3"]).float().abs() + 1e-3).all()
    c2 = bd.binary_matmul(torch.from_numpy(g["a"]).to(DEV), torch.from_numpy(g["b"]).to(DEV))
    assert O.rel_mean_abs_err(to_np(c2), g["c"].astype(np.float32)) < 2e-4


def test_binary_bmm_assertions():
    a = torch.zeros(2, 4, 64, device=DEV, dtype=torch.bfloat16)
    b = torch.zeros(2, 2, 8, device=DEV, dtype=torch.int32)
    with pytest.raises(AssertionError, match="Matrix A must be 3D"):
        bd.binary_bmm(a[0], b)
    with pytest.raises(AssertionError, match="Incompatible dimensions"):
        bd.binary_bmm(a[:, :, :32], b)
    with pytest.raises(AssertionError, match="Incompatible batch dimensions"):
        bd.binary_bmm(a, b[:1])
    with pytest.raises(AssertionError, match="Matrix A must be contiguous"):
        bd.binary_bmm(a.transpose(0, 1).contiguous().transpose(0, 1), b)
    with pytest.raises(AssertionError, match="same device"):
        bd.binary_bmm(a, b.cpu())
    assert bd.binary_bmm(a[:, :0], b).shape == (2, 0, 8)


# ------------------------------------------------------------------------------------------------ BinaryDiff.forward
@pytest.mark.parametrize("kernel", KERNELS)
def test_binarydiff_forward_reference_vectors(golden, kernel):
    g = golden("binarydiff_forward.npz")
    m = bd.BinaryDiff(t_bf16(g["base"]).to(DEV), t_bf16(g["finetune"]).to(DEV))
    m.kernel = kernel
    assert np.array_equal(m.mask.cpu().numpy(), g["mask"])
    x = t_bf16(g["x"]).to(DEV)
    y = m(x)
    assert y.shape == (2, 5, 160) and y.dtype == torch.bfloat16
    exact = O.binarydiff_forward_exact(bf(g["x"]), bf(g["base"]), g["mask"], g["coeff"])
    assert_close_to_exact(y, exact, "golden forward")
    # the reference's own CPU evaluations of diff.py:39
    assert O.rel_mean_abs_err(to_np(y), g["y_f32"]) < 2e-3
    assert O.rel_mean_abs_err(to_np(y), bf(g["y_bf16"])) < 4e-3


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("M", [1, 16, 128, 300])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_config1_single_4096_linear(kernel, M, dtype):
    # BASELINE config 1 / SURVEY 8d: seed 0, W_base = randn*0.02, W_fine = W_base + randn*0.002, x = randn(1, M, 4096)
    torch.manual_seed(0)
    base = (torch.randn(4096, 4096) * 0.02).to(dtype)
    fine = (base.float() + torch.randn(4096, 4096) * 0.002).to(dtype)
    x = torch.randn(1, M, 4096).to(dtype)
    m = bd.BinaryDiff(base.to(DEV), fine.to(DEV))
    m.kernel = kernel
    y = m(x.to(DEV))
    mask = m.mask.cpu().numpy()
    mask_o, coeff_o = O.binarydiff_compress(to_np(base), to_np(fine))
    assert np.array_equal(mask, mask_o)
    assert np.array_equal(bd.unpack(m.mask).cpu().numpy(), O.unpack(mask_o))  # bit-exact int32 unpack
    exact = O.binarydiff_forward_exact(to_np(x), to_np(base), mask, m.coeff.item())
    assert_close_to_exact(y, exact, f"config1 M={M}")
    if dtype == torch.bfloat16:
        emu = O.binarydiff_forward(to_np(x), to_np(base), mask, m.coeff.item())
        assert O.rel_mean_abs_err(to_np(y), emu) < 4e-3


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("N,K,rows", [(50, 32, 3), (64, 64, 1), (200, 96, 9), (1024, 4096, 6), (72, 2080, 17), (132, 160, 128), (4100, 4128, 2), (640, 1024, 700), (4096, 4096, 1500)])
def test_forward_ragged_shapes(kernel, N, K, rows):
    torch.manual_seed(N + K + rows)
    base = (torch.randn(N, K) * 0.05).bfloat16()
    fine = (base.float() + torch.randn(N, K) * 0.01).bfloat16()
    x = torch.randn(rows, K).bfloat16()
    m = bd.BinaryDiff(base.to(DEV), fine.to(DEV))
    m.kernel = kernel
    y = m(x.to(DEV))
    exact = O.binarydiff_forward_exact(to_np(x), to_np(base), m.mask.cpu().numpy(), m.coeff.item())
    assert_close_to_exact(y, exact, f"ragged {N}x{K}x{rows}")


@pytest.mark.parametrize("kernel", ["simt", "umma"])
def test_forward_is_deterministic_and_workspace_is_left_clean(kernel):
    # split-K partials are combined in a fixed order, so repeated launches are bit-identical; the tile counters in the
    # workspace are reset by the last CTA, so back-to-back launches of different shapes on one workspace stay correct
    torch.manual_seed(1)
    base = (torch.randn(1024, 4096, device=DEV) * 0.02).bfloat16()
    fine = (base.float() + torch.randn_like(base.float()) * 0.002).bfloat16()
    m = bd.BinaryDiff(base, fine)
    m.kernel = kernel
    base2 = (torch.randn(4096, 1024, device=DEV) * 0.02).bfloat16()
    m2 = bd.BinaryDiff(base2, (base2.float() + torch.randn_like(base2.float()) * 0.002).bfloat16())
    m2.kernel = kernel
    x = torch.randn(1, 6, 4096, device=DEV).bfloat16()
    x2 = torch.randn(1, 3, 1024, device=DEV).bfloat16()
    y0, z0 = m(x), m2(x2)
    for _ in range(5):
        assert torch.equal(m(x), y0)
        assert torch.equal(m2(x2), z0)


def test_auto_selects_the_tcgen05_kernel_for_baseline_shapes():
    from bitdelta_b200 import _lib

    sel = _lib.lib.bd_select_kernel
    for T, m, K, N in [(6, 1, 4096, 4096), (6, 1, 4096, 1024), (6, 1, 4096, 14336), (6, 1, 14336, 4096),  # Mistral-7B + 6 deltas
                       (1, 1, 4096, 11008), (1, 1, 11008, 4096), (1, 64, 4096, 4096),                     # Llama-2-7B + 1 delta
                       (4, 1, 5120, 13824), (4, 16, 5120, 5120), (1, 128, 8192, 1024)]:
        assert sel(0, T, m, K, N, 1) == _lib.KERNEL_UMMA, (T, m, K, N)
        assert sel(1, T, m, K, N, 0) == _lib.KERNEL_UMMA, (T, m, K, N)
    assert sel(0, 1, 1, 32, 50, 1) == _lib.KERNEL_SIMT       # N % 4 != 0
    # more rows / tenants than one launch takes are decomposed into several tcgen05 launches
    assert sel(0, 1, 4096, 4096, 4096, 1) == _lib.KERNEL_UMMA and sel(0, 8, 1, 8192, 8192, 1) == _lib.KERNEL_UMMA
    with pytest.raises(RuntimeError, match="does not support"):
        bd.binary_bmm(torch.zeros(1, 1, 32, device=DEV, dtype=torch.bfloat16), torch.zeros(1, 1, 50, device=DEV, dtype=torch.int32), kernel="umma")


# ------------------------------------------------------------------------------------------------ multi-tenant modules
@pytest.mark.parametrize("kernel", KERNELS)
def test_diffcompress_reference_vectors(golden, kernel):
    g = golden("demo_modules.npz")
    lin = torch.nn.Linear(128, 96, bias=False).to(DEV, torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_(t_bf16(g["weight"]).to(DEV))
    mod = bd.DiffCompressModule(lin, torch.from_numpy(g["masks"]).to(DEV), t_bf16(g["coeffs"]).to(DEV))
    mod.kernel = kernel
    y = mod(t_bf16(g["x"]).to(DEV))
    exact = O.diffcompress_forward_exact(bf(g["x"]), bf(g["weight"]), g["masks"], bf(g["coeffs"]))
    assert_close_to_exact(y, exact, "DiffCompressModule golden")
    assert O.rel_mean_abs_err(to_np(y), bf(g["y"])) < 4e-3  # vs the reference module's bf16 output


def test_dataparallel_reference_vectors(golden):
    g = golden("demo_modules.npz")
    x = t_bf16(g["x"]).to(DEV)
    head = torch.nn.Linear(128, 50, bias=False).to(DEV, torch.bfloat16)
    ws = [t_bf16(g[f"head_w{i}"]).to(DEV) for i in range(3)]
    logits = bd.DataParallelModule(head, ws)(x)
    ref = bf(g["logits"])
    assert logits.shape == (3, 2, 52)
    assert torch.all(logits[0, :, 50:] == torch.finfo(torch.bfloat16).min)
    assert O.rel_mean_abs_err(to_np(logits)[:, :, :50], ref[:, :, :50]) < 2e-3
    assert O.rel_mean_abs_err(to_np(logits)[1:], ref[1:]) < 2e-3
    emb = torch.nn.Embedding(52, 128).to(DEV, torch.bfloat16)
    out = bd.DataParallelModule(emb, [t_bf16(e).to(DEV) for e in g["emb_w"]])(torch.from_numpy(g["ids"]).to(DEV))
    assert np.array_equal(to_np(out), bf(g["emb_out"]))


@pytest.mark.parametrize("delta_grad_x", [False, True])
def test_binarydiff_scale_gradients(delta_grad_x):
    """Scale distillation (train.py:60-97): gradients of the fused op against autograd over the reference's composition
    y = x @ base + coeff * D, where D = binary_bmm(x, mask) is a constant for autograd (no grad_fn in the reference)."""
    torch.manual_seed(11)
    N, K, M = 256, 512, 24
    base = (torch.randn(N, K, device=DEV) * 0.02).bfloat16()
    fine = (base.float() + torch.randn(N, K, device=DEV) * 0.002).bfloat16()
    mod = bd.BinaryDiff(base, fine)
    mod.delta_grad_x = delta_grad_x
    x = torch.randn(2, M // 2, K, device=DEV).bfloat16().requires_grad_(True)
    gy = torch.randn(2, M // 2, N, device=DEV).bfloat16()
    y = mod(x)
    assert y.grad_fn is not None
    y.backward(gy)
    g_coeff, g_x = mod.coeff.grad.clone(), x.grad.clone()
    # fp64 autograd over the same composition
    sign = (bd.unpack(mod.mask).double() * 2 - 1)  # [K, N]
    xr = x.detach().double().requires_grad_(True)
    cr = mod.coeff.detach().double().requires_grad_(True)
    d = xr @ sign
    yr = xr @ base.double().T + cr * (d if delta_grad_x else d.detach())
    yr.backward(gy.double())
    assert abs(g_coeff.item() - cr.grad.item()) <= 2e-3 * abs(cr.grad.item()) + 1e-3 * (gy.double().abs() * d.detach().abs()).sum().item() ** 0.5
    rel = ((g_x.double() - xr.grad).abs().mean() / xr.grad.abs().mean()).item()
    assert rel < 4e-3, rel
    with torch.no_grad():
        assert mod(x).grad_fn is None


def test_tiny_llama_multi_tenant_logits_match_reference_fold(golden):
    """Model level: a 2-layer Llama served through register_diff_compress (fused 1-bit-delta linears + the native
    per-tenant leaves) must reproduce the logits the REFERENCE computed after folding the same diff.pt into the weights
    (load_diff, diff.py:81-106; recorded by gen_golden.py).  The two differ only by the bf16 rounding of the folded weight."""
    from transformers import LlamaConfig, LlamaForCausalLM

    from bitdelta_b200 import _lib
    from bitdelta_b200 import demo_backend as db

    g = golden("tiny_llama.npz")
    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=96, max_position_embeddings=64, tie_word_embeddings=False)
    sd = {k[len("basesd::"):]: t_bf16(g[k]) for k in g.files if k.startswith("basesd::")}
    model = LlamaForCausalLM(cfg).to(torch.bfloat16)
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    path = os.path.join(os.path.dirname(__file__), "golden", "tiny_llama_diff.pt")
    ckpts = []
    for _ in range(2):  # two tenants carrying the same delta: row t of the batch is served by checkpoint t
        d = torch.load(path, weights_only=False)
        ckpts.append({k: (v.detach().to(DEV).to(torch.bfloat16) if v.is_floating_point() else v.to(DEV)) for k, v in d.items()})
    db.cached_modules.clear()
    try:
        db.register_diff_compress(model, ckpts)
        assert db.fuse_sibling_projections(model) == 4
        assert isinstance(model.lm_head, db.DataParallelModule) and isinstance(model.model.norm, db.DataParallelModule)
        ids = torch.from_numpy(g["ids"]).to(DEV)
        n0 = _lib.launch_count()
        with torch.no_grad():
            logits = model(ids).logits.float().cpu().numpy()
        assert _lib.launch_count() > n0
        ref = g["folded_logits"]
        assert logits.shape == ref.shape
        assert O.rel_mean_abs_err(logits, ref) < 3e-2
        assert (logits.argmax(-1) == ref.argmax(-1)).mean() >= 0.9
        with torch.no_grad():  # decode-size call: one token per tenant, lm_head on the native ragged kernel
            one = model(ids[:, :1]).logits.float().cpu().numpy()
        assert O.rel_mean_abs_err(one, ref[:, :1]) < 3e-2
    finally:
        db.unregister_diff_compress(model)
        db.cached_modules.clear()
    assert isinstance(model.lm_head, torch.nn.Linear)


def test_decode_loop_over_registered_model_matches_folded_model(golden):
    """f-3: the greedy multi-tenant decode loop (demo_backend.py:190-258) over a model served through the fused modules must
    produce the tokens HF's greedy generate produces on the model with the SAME diff folded into its weights (load_diff)."""
    from transformers import LlamaConfig, LlamaForCausalLM

    from bitdelta_b200 import demo_backend as db

    g = golden("tiny_llama.npz")
    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=96, max_position_embeddings=64, tie_word_embeddings=False)
    sd = {k[len("basesd::"):]: t_bf16(g[k]) for k in g.files if k.startswith("basesd::")}
    path = os.path.join(os.path.dirname(__file__), "golden", "tiny_llama_diff.pt")
    folded = LlamaForCausalLM(cfg).to(torch.bfloat16)
    folded.load_state_dict(sd)
    bd.load_diff(folded, path)
    folded = folded.to(DEV).eval()
    model = LlamaForCausalLM(cfg).to(torch.bfloat16)
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    ckpts = []
    for _ in range(2):
        d = torch.load(path, weights_only=False)
        ckpts.append({k: (v.detach().to(DEV).to(torch.bfloat16) if v.is_floating_point() else v.to(DEV)) for k, v in d.items()})
    ids = torch.from_numpy(g["ids"]).to(DEV)[:, :8]
    mask = torch.ones_like(ids)
    n = 6
    with torch.no_grad():
        ref = folded.generate(ids, attention_mask=mask, max_new_tokens=n, do_sample=False, pad_token_id=0, eos_token_id=None)[:, 8:]
        ref_logits = folded(ids).logits[:, -1].float()
    db.cached_modules.clear()
    try:
        db.register_diff_compress(model, ckpts)
        db.fuse_sibling_projections(model)
        ours = bd.greedy_decode(model, ids, mask, n)
    finally:
        db.unregister_diff_compress(model)
        db.cached_modules.clear()
    assert ours.shape == (2, n)
    # bf16 near-ties can flip an argmax between the folded and the unfolded arithmetic (and a flipped token changes the rest
    # of a random-init model's sequence): require the first token to agree unless the top-2 margin is within bf16 noise
    top2 = ref_logits.topk(2, dim=-1).values
    margin = (top2[:, 0] - top2[:, 1]) / top2[:, 0].abs().clamp_min(1e-6)
    for t in range(2):
        if margin[t] > 0.05:
            assert ours[t, 0] == ref[t, 0]
    assert int(ours.min()) >= 0 and int(ours.max()) < 96  # never an index from the finfo.min padding


def test_graphed_decode_step_matches_the_eager_loop(golden):
    """f-3: the whole decode step (embedding, fused BinaryDiff launches, per-tenant norms and lm_heads, argmax, bookkeeping)
    captured as ONE CUDA graph over a static KV cache must produce exactly the tokens of the eager loop on the same modules."""
    from transformers import LlamaConfig, LlamaForCausalLM

    from bitdelta_b200 import demo_backend as db
    from bitdelta_b200.decode import GraphedDecoder

    g = golden("tiny_llama.npz")
    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=96, max_position_embeddings=64, tie_word_embeddings=False)
    sd = {k[len("basesd::"):]: t_bf16(g[k]) for k in g.files if k.startswith("basesd::")}
    path = os.path.join(os.path.dirname(__file__), "golden", "tiny_llama_diff.pt")
    model = LlamaForCausalLM(cfg).to(torch.bfloat16)
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    ckpts = []
    for i in range(3):
        d = torch.load(path, weights_only=False)
        ck = {k: (v.detach().to(DEV).to(torch.bfloat16) if v.is_floating_point() else v.to(DEV)) for k, v in d.items()}
        for k in ck:  # make the tenants differ
            if k.endswith(".coeff"):
                ck[k] = ck[k] * (1.0 + 0.5 * i)
        ckpts.append(ck)
    ids = torch.from_numpy(g["ids"]).to(DEV)[:1, :8].repeat(3, 1)
    mask = torch.ones_like(ids)
    mask[1, :3] = 0
    ids = ids * mask
    n = 10
    db.cached_modules.clear()
    try:
        db.register_diff_compress(model, ckpts)
        db.fuse_sibling_projections(model)
        dynamic = bd.greedy_decode(model, ids, mask, n)  # the reference's loop: growing cache, re-concatenated mask
        dec = GraphedDecoder(model, max_cache_len=32)
        static_eager = dec.decode(ids, mask, n, use_graph=False).clone()
        graphed = dec.decode(ids, mask, n)
        assert dec.graph is not None
        again = dec.decode(ids, mask, n)
    finally:
        db.unregister_diff_compress(model)
        db.cached_modules.clear()
    # same modules, same kernels, same shapes: the graph changes the launch mechanism only
    assert torch.equal(graphed, static_eager), f"graphed {graphed.tolist()} vs eager static-cache steps {static_eager.tolist()}"
    assert torch.equal(graphed, again)
    # against the growing-cache loop the attention kernel sees another key length (summation order): the first tokens agree,
    # later ones may part at a bf16 near-tie of a random-init model
    assert torch.equal(graphed[:, 0], dynamic[:, 0])
    assert (graphed == dynamic).float().mean() > 0.5


def t_16(bits: np.ndarray, tag: str) -> torch.Tensor:
    return torch.from_numpy(bits.view(np.int16).copy()).view(torch.bfloat16 if tag == "bf16" else torch.float16)


@pytest.mark.parametrize("tag", ["bf16", "fp16"])
@pytest.mark.parametrize("m", [1, 3])
def test_tenant_leaves_reference_vectors(golden, tag, m):
    """DataParallelModule's three native leaf kernels against the reference's own module outputs (gen_tenant_leaves)."""
    from transformers.models.llama.modeling_llama import LlamaRMSNorm

    from bitdelta_b200 import _lib

    g = golden("tenant_leaves.npz")
    dt = torch.bfloat16 if tag == "bf16" else torch.float16
    dec = bf if tag == "bf16" else (lambda b: np.asarray(b).view(np.float16).astype(np.float32))
    ulp = 2.0**-8 if tag == "bf16" else 2.0**-11
    pre = f"{tag}_m{m}_"
    x = t_16(g[pre + "x"], tag).to(DEV)
    vocab = (70, 75, 64)
    # lm_head: ragged vocab, finfo.min padding
    head = torch.nn.Linear(256, vocab[0], bias=False).to(DEV, dt)
    ws = [t_16(g[pre + f"head_w{t}"], tag).to(DEV) for t in range(3)]
    n0 = _lib.launch_count()
    logits = bd.DataParallelModule(head, ws)(x)
    assert _lib.launch_count() == n0 + 1, "lm_head leaf must be ONE native launch"
    ref = dec(g[pre + "logits"])
    got = to_np(logits)
    assert got.shape == ref.shape == (3, m, 75)
    for t, v in enumerate(vocab):
        assert np.array_equal(got[t, :, v:], ref[t, :, v:]) and np.all(ref[t, :, v:] == float(torch.finfo(dt).min))
        assert np.all(np.abs(got[t, :, :v] - ref[t, :, :v]) <= 2 * ulp * np.abs(ref[t, :, :v]) + 1e-6)
    # RMSNorm with per-tenant weights (HF LlamaRMSNorm is what the reference's model instantiates)
    norm = LlamaRMSNorm(256, eps=float(g["eps"])).to(DEV, dt)
    nws = [t_16(w, tag).to(DEV) for w in g[pre + "norm_w"]]
    n0 = _lib.launch_count()
    normed = bd.DataParallelModule(norm, nws)(x)
    assert _lib.launch_count() == n0 + 1
    refn = dec(g[pre + "normed"])
    gotn = to_np(normed)
    assert np.all(np.abs(gotn - refn) <= 2 * ulp * np.abs(refn) + 1e-6)
    assert (gotn == refn).mean() > 0.98
    # embed_tokens
    emb = torch.nn.Embedding(75, 256).to(DEV, dt)
    ews = [t_16(g[pre + f"emb_w{t}"], tag).to(DEV) for t in range(3)]
    n0 = _lib.launch_count()
    e = bd.DataParallelModule(emb, ews)(torch.from_numpy(g[pre + "ids"]).to(DEV))
    assert _lib.launch_count() == n0 + 1
    assert np.array_equal(to_np(e), dec(g[pre + "emb_out"]))


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("m", [1, 2, 4])
def test_tenant_linear_lm_head_size(dt, m):
    """Mistral-7B lm_heads of the six demo tenants (vocab 32000 x2, 32002 x4), decode rows, against the oracle."""
    torch.manual_seed(3)
    K, vocab = 4096, (32000, 32000, 32002, 32002, 32002, 32002)
    x = torch.randn(6, m, K, device=DEV).to(dt)
    ws = [(torch.randn(v, K, device=DEV) * 0.02).to(dt) for v in vocab]
    head = torch.nn.Linear(K, vocab[0], bias=False).to(DEV, dt)
    y = bd.DataParallelModule(head, ws)(x)
    assert y.shape == (6, m, 32002)
    assert torch.all(y[:2, :, 32000:] == torch.finfo(dt).min)
    for t, v in enumerate(vocab):
        exact = to_np(x[t]).astype(np.float64) @ to_np(ws[t]).astype(np.float64).T
        assert_close_to_exact(y[t, :, :v], exact, f"lm_head tenant {t}")
    # same launch under CUDA-graph capture and replay
    mod = bd.DataParallelModule(head, ws)
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        mod(x)
        torch.cuda.current_stream().synchronize()
        with torch.cuda.graph(graph, stream=side):
            yg = mod(x)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(yg, y)


def test_tenant_linear_shared_bias_and_many_tenants():
    """A wrapped Linear with a bias (shared by the tenants, equal widths) and more tenants than one launch group (32)."""
    torch.manual_seed(4)
    T, K, N = 37, 264, 45
    x = torch.randn(T, 1, K, device=DEV).bfloat16()
    ws = [(torch.randn(N, K, device=DEV) * 0.05).bfloat16() for _ in range(T)]
    head = torch.nn.Linear(K, N, bias=True).to(DEV, torch.bfloat16)
    y = bd.DataParallelModule(head, ws)(x)
    for t in range(T):
        exact = to_np(x[t]).astype(np.float64) @ to_np(ws[t]).astype(np.float64).T + to_np(head.bias).astype(np.float64)
        assert_close_to_exact(y[t], exact, f"bias tenant {t}")


def test_dataparallel_generic_module_keeps_reference_loop():
    """A leaf kind without a native kernel (LayerNorm with bias) goes through the reference's per-tenant loop."""
    torch.manual_seed(5)
    ln = torch.nn.LayerNorm(64).to(DEV, torch.bfloat16)
    ws = [torch.randn(64, device=DEV).bfloat16() for _ in range(3)]
    x = torch.randn(3, 2, 64, device=DEV).bfloat16()
    y = bd.DataParallelModule(ln, ws)(x)
    for t in range(3):
        ref = torch.nn.functional.layer_norm(x[t], (64,), ws[t], ln.bias, ln.eps)
        assert torch.equal(y[t], ref)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("N,K,m,T", [(1024, 4096, 1, 6), (4096, 14336, 1, 6), (4096, 4096, 3, 6), (1024, 8192, 1, 8), (512, 1024, 20, 3), (384, 512, 150, 2)])
def test_mistral_shapes_six_tenants(kernel, N, K, m, T):
    # BASELINE config 3 shapes (k_proj, down_proj, q_proj), T = 6 tenants, decode rows; config 5's 8 tenants (two
    # launches of the tcgen05 kernel) and a multi-tenant case with more than 16 rows per tenant (one launch per tenant)
    gen = torch.Generator(device=DEV).manual_seed(N + K + m)
    w = (torch.randn(N, K, generator=gen, device=DEV) * 0.02).bfloat16()
    masks = torch.randint(-(2**31), 2**31 - 1, (T, K // 32, N), generator=gen, device=DEV, dtype=torch.int64).to(torch.int32)
    coeffs = (torch.rand(T, generator=gen, device=DEV) * 0.003 + 0.0005).bfloat16()
    x = torch.randn(T, m, K, generator=gen, device=DEV).bfloat16()
    lin = torch.nn.Linear(K, N, bias=False, device=DEV, dtype=torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_(w)
    mod = bd.DiffCompressModule(lin, masks, coeffs)
    mod.kernel = kernel
    y = mod(x)
    # float64 truth on the GPU from the bit-exact unpack (checked against the oracle on a slab below)
    signs = bd.unpack(masks).double() * 2 - 1
    exact = x.double() @ w.double().T + coeffs.double()[:, None, None] * torch.bmm(x.double(), signs)
    assert_close_to_exact(y, exact.cpu().numpy(), f"mistral {N}x{K}")
    sl = slice(0, 96)
    ex_o = O.diffcompress_forward_exact(to_np(x), to_np(w[sl]), masks[:, :, sl].cpu().numpy(), to_np(coeffs))
    assert np.allclose(exact[:, :, sl].cpu().numpy(), ex_o, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("N,K,T", [(40000, 64, 6), (30000, 128, 1), (19000, 192, 6), (9000, 320, 3), (33000, 512, 8), (32000, 4096, 6),
                                   (20000, 256, 5), (6000, 1024, 7), (12000, 512, 10), (3000, 2048, 2), (5000, 768, 4)])  # T = 5, 7, 10: the MMA issuer's generic loop
def test_decode_many_runs_per_cta(N, K, T):
    # Short K against many weight-row tiles: every CTA of the tcgen05 kernel works through several (tile, K run)s -- runs of
    # ONE unit (K = 64: a warp group then has no unit of its own in every other run), whole-tile runs in the middle of a
    # CTA's range, partial runs at both ends.  Exercises the deferred read-out (next run's first unit handed over before
    # the previous run's accumulators are read) and the accumulators-read handshake with the MMA warp at every boundary.
    gen = torch.Generator(device=DEV).manual_seed(N + K + T)
    w = (torch.randn(N, K, generator=gen, device=DEV) * 0.05).bfloat16()
    masks = torch.randint(-(2**31), 2**31 - 1, (T, K // 32, N), generator=gen, device=DEV, dtype=torch.int64).to(torch.int32)
    coeffs = (torch.rand(T, generator=gen, device=DEV) * 0.01 + 0.001).bfloat16()
    x = torch.randn(T, 1, K, generator=gen, device=DEV).bfloat16()
    lin = torch.nn.Linear(K, N, bias=False, device=DEV, dtype=torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_(w)
    mod = bd.DiffCompressModule(lin, masks, coeffs)
    mod.kernel = "umma"
    signs = bd.unpack(masks).double() * 2 - 1
    exact = (x.double() @ w.double().T + coeffs.double()[:, None, None] * torch.bmm(x.double(), signs)).cpu().numpy()
    for rep in range(3):  # back to back: the workspace (arrival counters, slots) left by one launch serves the next
        assert_close_to_exact(mod(x), exact, f"many runs {N}x{K} T={T} launch {rep}")


@pytest.mark.parametrize("shapes,T,m", [([4096, 1024, 1024], 6, 1), ([14336, 14336], 6, 1), ([256, 384, 200], 3, 2), ([1024, 1024], 8, 1)])
def test_grouped_launch_matches_individual_modules(shapes, T, m):
    # q/k/v and gate/up called back to back on the same input share ONE launch (SiblingGroup); results must match the
    # float64 truth like the individual launches do, and a member called on a different input must not see cached data
    K = 4096 if shapes[0] >= 1024 else 512
    gen = torch.Generator(device=DEV).manual_seed(sum(shapes) + T)
    mods, exacts = [], []
    x = torch.randn(T, m, K, generator=gen, device=DEV).bfloat16()
    x2 = torch.randn(T, m, K, generator=gen, device=DEV).bfloat16()
    for N in shapes:
        lin = torch.nn.Linear(K, N, bias=False, device=DEV, dtype=torch.bfloat16)
        with torch.no_grad():
            lin.weight.normal_(0.0, 0.02, generator=gen)
        masks = torch.randint(-(2**31), 2**31 - 1, (T, K // 32, N), generator=gen, device=DEV, dtype=torch.int64).to(torch.int32)
        coeffs = (torch.rand(T, generator=gen, device=DEV) * 0.003 + 0.0005).bfloat16()
        mods.append(bd.DiffCompressModule(lin, masks, coeffs))
    def truth(mod, xin):
        signs = bd.unpack(mod.mask).double() * 2 - 1
        return (xin.double() @ mod.module.weight.detach().double().T + mod.coeff.double()[:, None, None] * torch.bmm(xin.double(), signs)).cpu().numpy()
    singles = [mod(x) for mod in mods]
    n0 = bd._lib.launch_count()
    bd.group_projections(mods)
    grouped = [mod(x) for mod in mods]
    launches = bd._lib.launch_count() - n0
    if all(N % 128 == 0 for N in shapes[:-1]) and T <= 6:
        assert launches == 1, launches
    for mod, yg, ys in zip(mods, grouped, singles):
        assert_close_to_exact(yg, truth(mod, x), "grouped")
        assert (yg.float() - ys.float()).abs().max() <= 2.0**-7 * ys.float().abs().max()  # at most an output ulp apart
    # out-of-pattern use: only the second member, on another input -> fresh launch, correct result
    y2 = mods[1](x2)
    assert_close_to_exact(y2, truth(mods[1], x2), "grouped, different input")
    y0 = mods[0](x)
    assert_close_to_exact(y0, truth(mods[0], x), "grouped, first member again")


def test_fuse_sibling_projections_on_a_decoder_like_block():
    import torch.nn as nn

    class Attn(nn.Module):
        def __init__(self):
            super().__init__()
            self.q_proj = nn.Linear(256, 256, bias=False)
            self.k_proj = nn.Linear(256, 128, bias=False)
            self.v_proj = nn.Linear(256, 128, bias=False)
            self.o_proj = nn.Linear(256, 256, bias=False)

        def forward(self, x):
            return self.o_proj(self.q_proj(x) + torch.cat([self.k_proj(x), self.v_proj(x)], dim=-1))

    torch.manual_seed(0)
    model = nn.Sequential()
    model.add_module("self_attn", Attn())
    model = model.to(DEV, torch.bfloat16)
    T = 2
    names = ["q_proj", "k_proj", "v_proj", "o_proj"]
    ckpts = [{f"self_attn.{n}.mask": bd.pack(torch.rand(256, getattr(model.self_attn, n).out_features, device=DEV) > 0.5)
              for n in names} for _ in range(T)]
    for c in ckpts:
        for n in names:
            c[f"self_attn.{n}.coeff"] = torch.tensor(0.004, device=DEV)
    bd.demo_backend.cached_modules.clear()
    x = torch.randn(T, 3, 256, device=DEV).bfloat16()
    with torch.no_grad():
        bd.register_diff_compress(model, ckpts)
        y_plain = model(x)
        assert bd.fuse_sibling_projections(model) == 1
        n0 = bd._lib.launch_count()
        y_fused = model(x)
        assert bd._lib.launch_count() - n0 == 2  # q/k/v in one launch + o_proj
        # a non-contiguous input is made contiguous once per round, not once per member
        xt = x.transpose(0, 1).contiguous().transpose(0, 1)
        assert not xt.is_contiguous()
        n0 = bd._lib.launch_count()
        y_nc = model(xt)
        assert bd._lib.launch_count() - n0 == 2 and torch.equal(y_nc, y_fused)
    # A static-buffer decode loop under inference_mode: q_proj consumes the buffer, the buffer is updated IN PLACE before
    # k_proj / v_proj of that round are ever called, and the block runs again.  Inference tensors carry no version counter,
    # so only the block's pre-forward hook can tell the rounds apart: the second forward must see the new activations.
    with torch.inference_mode():
        buf = x.clone()
        q_only = model.self_attn.q_proj(buf)     # starts a round and leaves k/v outputs cached
        buf.copy_(x * 0.5)                        # same storage, no version bump visible under inference_mode
        y_new = model(buf)
        buf2 = (x * 0.5).clone()
        bd.unregister_diff_compress(model)
        bd.register_diff_compress(model, ckpts)
        y_ref = model(buf2)
        bd.unregister_diff_compress(model)
    assert torch.equal(q_only, model.self_attn.q_proj(x) * 0 + q_only)  # (keeps q_only alive)
    assert (y_fused.float() - y_plain.float()).abs().max() <= 2.0**-6 * y_plain.float().abs().max()
    assert torch.equal(y_new, y_ref), "stale sibling outputs were reused after an in-place update of the activation buffer"
    bd.demo_backend.cached_modules.clear()


def test_linearity_in_the_coefficient():
    # size-independent property at full size: y(2c) - y(c) == y(c) - y(0)  up to bf16 rounding
    T, N, K = 6, 4096, 4096
    gen = torch.Generator(device=DEV).manual_seed(3)
    w = (torch.randn(N, K, generator=gen, device=DEV) * 0.02).bfloat16()
    masks = torch.randint(-(2**31), 2**31 - 1, (T, K // 32, N), generator=gen, device=DEV, dtype=torch.int64).to(torch.int32)
    x = torch.randn(T, 1, K, generator=gen, device=DEV).bfloat16()
    lin = torch.nn.Linear(K, N, bias=False, device=DEV, dtype=torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_(w)
    c = torch.full((T,), 0.002, device=DEV)
    y0 = bd.DiffCompressModule(lin, masks, c * 0)(x).float()
    y1 = bd.DiffCompressModule(lin, masks, c)(x).float()
    y2 = bd.DiffCompressModule(lin, masks, c * 2)(x).float()
    base_only = (x.float() @ w.float().T)
    assert O.rel_mean_abs_err(to_np(y0), to_np(base_only)) < 2e-3
    d1, d2 = (y1 - y0), (y2 - y1)
    delta = 0.002 * torch.bmm(x.float(), bd.unpack(masks).float() * 2 - 1)
    assert (d1 - delta).abs().mean() / delta.abs().mean() < 0.1  # differences of bf16-rounded outputs: coarse
    assert (d2 - delta).abs().mean() / delta.abs().mean() < 0.1


def test_register_and_unregister_diff_compress():
    import torch.nn as nn

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.q_proj = nn.Linear(64, 96, bias=False)
            self.norm = nn.LayerNorm(64, bias=False)

        def forward(self, x):
            return self.q_proj(self.norm(x))

    class Model(nn.Module):
        def __init__(self):
            super().__init__()
            self.block = Block()

        def forward(self, x):
            return self.block(x)

    torch.manual_seed(0)
    model = Model().to(DEV, torch.bfloat16)
    T = 3
    ckpts = []
    for _ in range(T):
        ckpts.append({
            "block.q_proj.mask": bd.pack(torch.rand(64, 96, device=DEV) > 0.5),
            "block.q_proj.coeff": torch.tensor(0.01, device=DEV),
            "block.norm.weight": torch.rand(64, device=DEV).bfloat16(),
        })
    ref_ckpts = [dict(c) for c in ckpts]
    bd.demo_backend.cached_modules.clear()
    x = torch.randn(T, 4, 64, device=DEV).bfloat16()
    with bd.DiffCompress(model, ckpts):
        assert isinstance(model.block.q_proj, bd.DiffCompressModule)
        assert isinstance(model.block.norm, bd.DataParallelModule)
        assert model.block.q_proj.mask.shape == (T, 2, 96)
        assert "block.q_proj.mask" not in ckpts[0]  # popped like the reference does
        y = model(x)
    assert isinstance(model.block.q_proj, nn.Linear) and isinstance(model.block.norm, nn.LayerNorm)
    for t in range(T):
        h = torch.nn.functional.layer_norm(x[t].float(), (64,), ref_ckpts[t]["block.norm.weight"].float()).bfloat16()
        signs = bd.unpack(ref_ckpts[t]["block.q_proj.mask"]).double() * 2 - 1
        exact = h.double() @ model.block.q_proj.weight.detach().double().T + 0.01 * (h.double() @ signs)
        assert_close_to_exact(y[t].detach(), exact.cpu().numpy(), f"tenant {t}")
    bd.demo_backend.cached_modules.clear()


def test_cuda_graph_capture_of_the_fused_forward():
    T, N, K = 6, 1024, 4096
    gen = torch.Generator(device=DEV).manual_seed(5)
    w = (torch.randn(N, K, generator=gen, device=DEV) * 0.02).bfloat16()
    masks = torch.randint(-(2**31), 2**31 - 1, (T, K // 32, N), generator=gen, device=DEV, dtype=torch.int64).to(torch.int32)
    coeffs = torch.full((T,), 0.002, device=DEV)
    x = torch.randn(T, 1, K, generator=gen, device=DEV).bfloat16()
    lin = torch.nn.Linear(K, N, bias=False, device=DEV, dtype=torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_(w)
    mod = bd.DiffCompressModule(lin, masks, coeffs)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        y_eager = mod(x)  # warm-up on the capture stream (allocates its workspace)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            y_graph = mod(x)
        graph.replay()
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(y_graph, y_eager)


# ------------------------------------------------------------------------------------------------ activation range
# The decode launches of the tcgen05 kernel (bf16, one row per tenant) feed the delta product from 8-bit operands: the
# activations are scaled by a power of two taken from the row's largest exponent and split into three e5m2 pieces
# (bd_umma.cu, row-scale pass + xperm_job; host-side model in tests/test_d8_split_model.py).  These cases
# leave the randn scale every other test uses: the result must stay inside the SAME tolerance whatever the magnitude of
# the row, like the reference, which accumulates the unrounded activation (binary_gemm_kernel.py:260-278).
SCALES = [1e-30, 1e-12, 1e-6, 1e-4, 1e-3, 1.0, 1e3, 6e4, 1e9, 1e30]


def _tenant_problem(T, m, K, N, seed):
    gen = torch.Generator(device=DEV).manual_seed(seed)
    w = (torch.randn(N, K, generator=gen, device=DEV) * 0.02).bfloat16()
    masks = torch.randint(-(2**31), 2**31 - 1, (T, K // 32, N), generator=gen, device=DEV, dtype=torch.int64).to(torch.int32)
    coeffs = (torch.rand(T, generator=gen, device=DEV) * 0.003 + 0.0005).bfloat16()
    x = torch.randn(T, m, K, generator=gen, device=DEV)
    return gen, w, masks, coeffs, x


@pytest.mark.parametrize("kernel", ["umma", "simt"])
@pytest.mark.parametrize("m", [1, 3])
@pytest.mark.parametrize("scale", SCALES)
def test_activation_scale_does_not_change_the_error(kernel, m, scale):
    T, K, N = 6, 4096, 1024
    _, w, masks, coeffs, x = _tenant_problem(T, m, K, N, 7 + m)
    x = (x * scale).bfloat16()
    signs = bd.unpack(masks).double() * 2 - 1
    # delta only (binary_bmm): nothing hides an error of the sign product
    c = bd.binary_bmm(x, masks, kernel=kernel)
    exact_d = torch.bmm(x.double(), signs)
    assert_close_to_exact(c, exact_d.cpu().numpy(), f"binary_bmm scale {scale:g}")
    # DiffCompressModule, with a coefficient large enough for the delta term to matter (0.5 .. 1.5)
    big = (coeffs.float() * 500).bfloat16()
    lin = torch.nn.Linear(K, N, bias=False, device=DEV, dtype=torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_(w)
    for cf in (coeffs, big):
        mod = bd.DiffCompressModule(lin, masks, cf)
        mod.kernel = kernel
        y = mod(x)
        exact = x.double() @ w.double().T + cf.double()[:, None, None] * exact_d
        assert_close_to_exact(y, exact.cpu().numpy(), f"DiffCompressModule scale {scale:g}")


@pytest.mark.parametrize("scale", [1e-7, 1e-6, 1e-4, 1e-2, 1.0, 1e2, 1e4])
def test_activation_scale_fp16_decode(scale):
    # fp16 activations (the dtype the reference's demo runs in, demo_backend.py:24) take the same 8-bit decode path with four
    # pieces per element; the scales reach from fp16's subnormals to just below its overflow
    T, m, K, N = 6, 1, 4096, 1024
    gen, w, masks, coeffs, x = _tenant_problem(T, m, K, N, 41)
    w, coeffs = w.half(), coeffs.half()
    x = (x * scale).clamp(-60000.0, 60000.0).half()
    signs = bd.unpack(masks).double() * 2 - 1
    exact_d = torch.bmm(x.double(), signs)
    c = bd.binary_bmm(x, masks, kernel="umma")
    assert torch.isfinite(c.float()).all() or scale >= 1e4  # sums of 4096 values near the top of fp16 may overflow the OUTPUT
    if scale < 1e4:
        assert_close_to_exact(c, exact_d.cpu().numpy(), f"binary_bmm fp16 scale {scale:g}")
    lin = torch.nn.Linear(K, N, bias=False, device=DEV, dtype=torch.float16)
    with torch.no_grad():
        lin.weight.copy_(w)
    for cf in (coeffs, (coeffs.float() * 300).half()):
        mod = bd.DiffCompressModule(lin, masks, cf)
        mod.kernel = "umma"
        y = mod(x)
        exact = x.double() @ w.double().T + cf.double()[:, None, None] * exact_d
        # (the OUTPUT must be representable: skip sums that overflow fp16 or sit in its subnormal range, where half an ulp
        # of the result is already more than the relative tolerance)
        if exact.abs().max() < 6e4 and exact.abs().mean() > 1e-3:
            assert_close_to_exact(y, exact.cpu().numpy(), f"DiffCompressModule fp16 scale {scale:g}")
        elif exact.abs().max() < 6e4:
            assert (y.double() - exact).abs().max() <= 2.0**-24 * 0.51 + 1e-3 * exact.abs().max()


@pytest.mark.parametrize("kernel", ["umma", "simt"])
def test_wide_dynamic_range_rows(kernel):
    # every row mixes magnitudes from 1e-9 to 1e9 (exponents drawn uniformly) plus a few massive outlier channels
    T, m, K, N = 6, 1, 4096, 512
    gen, w, masks, coeffs, x = _tenant_problem(T, m, K, N, 11)
    expo = torch.rand(T, m, K, generator=gen, device=DEV) * 18 - 9
    x = (x * torch.pow(10.0, expo))
    x[:, :, ::251] *= 1e4
    x = x.bfloat16()
    signs = bd.unpack(masks).double() * 2 - 1
    exact_d = torch.bmm(x.double(), signs)
    c = bd.binary_bmm(x, masks, kernel=kernel)
    assert_close_to_exact(c, exact_d.cpu().numpy(), "binary_bmm wide rows")
    # ... and rows whose elements differ by tenant: tenant t lives around 10^(4t - 10)
    x2 = (torch.randn(T, m, K, generator=gen, device=DEV) * torch.pow(10.0, 4.0 * torch.arange(T, device=DEV) - 10.0)[:, None, None]).bfloat16()
    lin = torch.nn.Linear(K, N, bias=False, device=DEV, dtype=torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_(w)
    mod = bd.DiffCompressModule(lin, masks, (coeffs.float() * 300).bfloat16())
    mod.kernel = kernel
    y = mod(x2)
    exact = x2.double() @ w.double().T + mod.coeff.double()[:, None, None] * torch.bmm(x2.double(), signs)
    got = y.double()
    for t in range(T):  # per tenant: the tolerance's absolute floor must not be set by the largest tenant
        assert_close_to_exact(y[t], exact[t].cpu().numpy(), f"tenant {t}")
    assert torch.isfinite(got).all()


def test_non_finite_and_huge_activations_are_not_clipped():
    # inf / NaN in one tenant's row: that tenant's outputs are non-finite (never a silently saturated number), the other
    # tenants are unaffected; a huge but finite activation (2^70, far outside e5m2's own range) is just another row maximum
    T, m, K, N = 6, 1, 1024, 256
    _, w, masks, coeffs, x = _tenant_problem(T, m, K, N, 13)
    x = x.bfloat16()
    lin = torch.nn.Linear(K, N, bias=False, device=DEV, dtype=torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_(w)
    mod = bd.DiffCompressModule(lin, masks, coeffs)
    mod.kernel = "umma"
    y_ref = mod(x).clone()
    for t, bad in ((1, float("inf")), (3, float("nan")), (4, float("-inf"))):
        xb = x.clone()
        xb[t, 0, 517] = bad
        others = [i for i in range(T) if i != t]
        for y in (mod(xb), bd.binary_bmm(xb, masks, kernel="umma")):
            assert not torch.isfinite(y[t]).any(), f"tenant {t} with {bad}"
            assert torch.isfinite(y[others]).all()
        assert torch.equal(mod(xb)[others], y_ref[others])
    xb = x.clone()
    xb[4, 0, 517] = 2.0**70
    xb[2, 0, 3] = -(2.0**-90)
    signs = bd.unpack(masks).double() * 2 - 1
    exact_d = torch.bmm(xb.double(), signs)
    c = bd.binary_bmm(xb, masks, kernel="umma")
    assert torch.isfinite(c.float()).all()
    for t in range(T):
        assert_close_to_exact(c[t], exact_d[t].cpu().numpy(), f"tenant {t} with a 2^70 outlier in tenant 4")


def test_magnitude_changes_from_block_to_block():
    # The 8-bit decode path scales a row by ONE power of two per CTA (taken from the largest exponent inside the CTA's K
    # range) and different CTAs see different K ranges.  Here every 64-wide K block has its own magnitude (some blocks are
    # all zero), so neighbouring CTAs pick different scales and the split of the small blocks runs far below the row maximum.
    T, m, K, N = 6, 1, 8192, 640
    gen, w, masks, coeffs, x = _tenant_problem(T, m, K, N, 17)
    expo = torch.tensor([-12.0, -6.0, 0.0, 6.0, 11.0], device=DEV)[torch.randint(0, 5, (T, K // 64), generator=gen, device=DEV)]
    blk = torch.pow(10.0, expo) * (torch.rand(T, K // 64, generator=gen, device=DEV) > 0.2)  # a fifth of the blocks is zero
    x = (x * blk.repeat_interleave(64, dim=1)[:, None, :]).bfloat16()
    signs = bd.unpack(masks).double() * 2 - 1
    exact_d = torch.bmm(x.double(), signs)
    for kernel in ("umma", "simt"):
        c = bd.binary_bmm(x, masks, kernel=kernel)
        assert_close_to_exact(c, exact_d.cpu().numpy(), f"binary_bmm per-block scales ({kernel})")
    lin = torch.nn.Linear(K, N, bias=False, device=DEV, dtype=torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_(w)
    mod = bd.DiffCompressModule(lin, masks, (coeffs.float() * 300).bfloat16())
    mod.kernel = "umma"
    y1, y2 = mod(x), mod(x)
    exact = x.double() @ w.double().T + mod.coeff.double()[:, None, None] * exact_d
    assert_close_to_exact(y1, exact.cpu().numpy(), "DiffCompressModule per-block scales")
    assert torch.equal(y1, y2)


# ------------------------------------------------------------------------------------------------ BASELINE shapes
# Every (N, K) of the models BASELINE.json names, through the reference-named modules, against the float64 truth built from
# the bit-exact unpack: Llama-2-7B (config 2, one delta: BinaryDiff at decode and prefill sizes), Llama-2-13B + 4 deltas
# (config 4), Llama-2-70B + 8 deltas as the per-rank shards of an 8-way tensor-parallel split (config 5: column-parallel
# q/k/v/gate/up slice N, row-parallel o/down slice K).
def _exact_forward(x, w, masks, coeffs):
    signs = bd.unpack(masks).double() * 2 - 1
    d = torch.bmm(x.double(), signs) if masks.dim() == 3 else x.double() @ signs
    c = coeffs.double().reshape(-1)
    return x.double() @ w.double().T + (c[:, None, None] if masks.dim() == 3 else c) * d


@pytest.mark.parametrize("N,K", [(4096, 4096), (11008, 4096), (4096, 11008)])
@pytest.mark.parametrize("rows", [1, 2048])
def test_llama2_7b_shapes_binarydiff(N, K, rows):
    gen = torch.Generator(device=DEV).manual_seed(N * 3 + K + rows)
    base = (torch.randn(N, K, generator=gen, device=DEV) * 0.02).bfloat16()
    fine = (base.float() + torch.randn(N, K, generator=gen, device=DEV) * 0.002).bfloat16()
    mod = bd.BinaryDiff(base, fine)
    assert bd._lib.lib.bd_select_kernel(bd._lib.BD_BF16, 1, rows, K, N, 1) == bd._lib.KERNEL_UMMA
    x = torch.randn(1, rows, K, generator=gen, device=DEV).bfloat16()
    with torch.no_grad():
        y = mod(x)
    exact = _exact_forward(x[0], base, mod.mask, mod.coeff.detach())
    assert_close_to_exact(y[0], exact.cpu().numpy(), f"llama-2-7b {N}x{K} rows {rows}")


@pytest.mark.parametrize("N,K,T,m", [
    (5120, 5120, 4, 1), (13824, 5120, 4, 1), (5120, 13824, 4, 1), (5120, 5120, 4, 64),       # Llama-2-13B + 4 deltas
    (1024, 8192, 8, 1), (128, 8192, 8, 1), (3584, 8192, 8, 1), (8192, 1024, 8, 1), (8192, 3584, 8, 1),  # 70B, TP = 8 shards
    (8192, 8192, 8, 1), (28672 // 4, 8192, 8, 1),                                            # 70B, TP = 1 / 4
])
def test_llama2_13b_and_70b_shard_shapes(N, K, T, m):
    gen = torch.Generator(device=DEV).manual_seed(N + 7 * K + T)
    w = (torch.randn(N, K, generator=gen, device=DEV) * 0.02).bfloat16()
    masks = torch.randint(-(2**31), 2**31 - 1, (T, K // 32, N), generator=gen, device=DEV, dtype=torch.int64).to(torch.int32)
    coeffs = (torch.rand(T, generator=gen, device=DEV) * 0.003 + 0.0005).bfloat16()
    x = torch.randn(T, m, K, generator=gen, device=DEV).bfloat16()
    lin = torch.nn.Linear(K, N, bias=False, device=DEV, dtype=torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_(w)
    mod = bd.DiffCompressModule(lin, masks, coeffs)
    y = mod(x)
    exact = _exact_forward(x, w, masks, coeffs)
    assert_close_to_exact(y, exact.cpu().numpy(), f"{N}x{K} T={T} m={m}")


def test_multi_tenant_prefill_13b():
    # config 4: prefill of a few hundred tokens per tenant through the 13B attention projection (the full 4096-token
    # prefill is the same code path, one 128-row chunk after the other)
    N, K, T, m = 5120, 5120, 4, 300
    gen = torch.Generator(device=DEV).manual_seed(4)
    w = (torch.randn(N, K, generator=gen, device=DEV) * 0.02).bfloat16()
    masks = torch.randint(-(2**31), 2**31 - 1, (T, K // 32, N), generator=gen, device=DEV, dtype=torch.int64).to(torch.int32)
    coeffs = (torch.rand(T, generator=gen, device=DEV) * 0.003 + 0.0005).bfloat16()
    x = torch.randn(T, m, K, generator=gen, device=DEV).bfloat16()
    lin = torch.nn.Linear(K, N, bias=False, device=DEV, dtype=torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_(w)
    y = bd.DiffCompressModule(lin, masks, coeffs)(x)
    assert_close_to_exact(y, _exact_forward(x, w, masks, coeffs).cpu().numpy(), "13B multi-tenant prefill")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("T,m,K,N", [(4, 600, 1024, 5120), (3, 130, 512, 8192), (2, 40, 2048, 384), (5, 129, 256, 128)])
def test_multi_tenant_prefill_in_one_launch(T, m, K, N, dtype):
    # more than 16 rows per tenant: all tenants' row chunks run in ONE launch (tile = (tenant, N tile, row chunk)); the first
    # two shapes have enough tiles for the strided whole-tile schedule, the others take the stream-K path with split-K fix-up
    gen = torch.Generator(device=DEV).manual_seed(T * 1000 + m)
    w = (torch.randn(N, K, generator=gen, device=DEV) * 0.02).to(dtype)
    masks = torch.randint(-(2**31), 2**31 - 1, (T, K // 32, N), generator=gen, device=DEV, dtype=torch.int64).to(torch.int32)
    coeffs = (torch.rand(T, generator=gen, device=DEV) * 0.3 + 0.05).to(dtype)  # large: a wrong tenant's signs or scale shows
    x = torch.randn(T, m, K, generator=gen, device=DEV).to(dtype)
    lin = torch.nn.Linear(K, N, bias=False, device=DEV, dtype=dtype)
    with torch.no_grad():
        lin.weight.copy_(w)
    mod = bd.DiffCompressModule(lin, masks, coeffs)
    mod.kernel = "umma"
    n0 = bd._lib.launch_count()
    y = mod(x)
    assert bd._lib.launch_count() - n0 == 1
    assert_close_to_exact(y, _exact_forward(x, w, masks, coeffs).cpu().numpy(), f"T={T} m={m}")
    assert torch.equal(y, mod(x))


# ------------------------------------------------------------------------------------------------ launch flags
@pytest.mark.parametrize("kernel", ["umma", "simt"])
@pytest.mark.parametrize("T,m,K,N", [(8, 1, 1024, 8192), (2, 40, 512, 384), (1, 300, 1024, 512)])
def test_fp32_partial_sum_output(kernel, T, m, K, N):
    # BD_FLAG_FP32_OUT: the unrounded fp32 sums (row-parallel tensor-parallel shards); rounding them gives the normal output
    from bitdelta_b200.diff import _fused_forward

    gen, w, masks, coeffs, x = _tenant_problem(T, m, K, N, 23)
    x = x.bfloat16()
    y32 = _fused_forward(x, w, masks, coeffs, T, kernel, out_fp32=True)
    y16 = _fused_forward(x, w, masks, coeffs, T, kernel)
    assert y32.dtype == torch.float32 and y32.shape == y16.shape
    assert torch.equal(y32.bfloat16(), y16)
    exact = _exact_forward(x, w, masks, coeffs)
    rel = ((y32.double() - exact).abs().mean() / exact.abs().mean()).item()
    assert rel < 1e-5, rel  # fp32 accumulation only


def test_operands_written_by_the_preceding_kernel_are_seen():
    # binary_bmm(a, pack(b)) (the reference notebook's call pattern) and freshly copied weights are produced by
    # the kernel right before the forward on the same stream; without BD_FLAG_STATIC_OPERANDS the launch must not prefetch
    # them ahead of that kernel.  Alternate two different operand sets in one buffer, back to back.
    T, m, K, N = 6, 1, 4096, 4096
    gen = torch.Generator(device=DEV).manual_seed(29)
    a = torch.randn(T, m, K, generator=gen, device=DEV).bfloat16()
    bits = [torch.rand(T, K, N, generator=gen, device=DEV) > 0.5 for _ in range(2)]
    want = [torch.bmm(a.double(), b.double() * 2 - 1) for b in bits]
    for i in range(12):
        c = bd.binary_bmm(a, bd.pack(bits[i % 2]), kernel="umma")
        assert_close_to_exact(c, want[i % 2].cpu().numpy(), f"iteration {i}")
    w2 = [(torch.randn(N, K, generator=gen, device=DEV) * 0.02).bfloat16() for _ in range(2)]
    masks = bd.pack(bits[0])
    coeffs = torch.full((T,), 0.002, device=DEV)
    lin = torch.nn.Linear(K, N, bias=False, device=DEV, dtype=torch.bfloat16)
    mod = bd.DiffCompressModule(lin, masks, coeffs)
    buf = torch.empty(K, N, device=DEV, dtype=torch.bfloat16)
    for i in range(8):
        buf.copy_(w2[i % 2].T)
        lin.weight.data = buf.T  # non-contiguous view: forward makes a fresh contiguous copy on the stream
        y = mod(a)
        exact = _exact_forward(a, w2[i % 2], masks, coeffs)
        assert_close_to_exact(y, exact.cpu().numpy(), f"weight copy {i}")


def test_captured_graph_survives_a_larger_eager_call():
    # the per-stream workspace is allocated once at its maximum; a decode graph captured first keeps working after
    # a prefill-sized eager call on the same stream (the workspace used to be re-allocated, leaving the graph a stale pointer)
    T, K, N = 6, 4096, 1024
    _, w, masks, coeffs, x = _tenant_problem(T, 1, K, N, 31)
    x = x.bfloat16()
    lin = torch.nn.Linear(K, N, bias=False, device=DEV, dtype=torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_(w)
    mod = bd.DiffCompressModule(lin, masks, coeffs)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        y_eager = mod(x).clone()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            y_graph = mod(x)
        ws_before = bd._lib.workspace(DEV, 6, N).data_ptr()
        big = torch.randn(1, 2000, K, device=DEV).bfloat16()
        fine = (w.float() + 0.002 * torch.randn_like(w.float())).bfloat16()
        with torch.no_grad():
            bd.BinaryDiff(w, fine)(big)
        assert bd._lib.workspace(DEV, 2000, N).data_ptr() == ws_before
        for _ in range(3):
            graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(y_graph, y_eager)
