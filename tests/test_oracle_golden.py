"""Pins oracle/bitdelta_oracle.py to vectors produced by the reference itself (tests/golden/gen_golden.py)."""
import numpy as np
import pytest

from oracle import bitdelta_oracle as O


def bf(bits):
    return O.bf16_bits_to_f32(bits)


def test_bf16_rounding_helpers():
    import torch

    x = np.random.default_rng(0).standard_normal(4096).astype(np.float32) * 3
    x[:4] = [0.0, -0.0, 1.00390625, 65504.0]  # includes a tie (1 + 2^-8)
    want = torch.from_numpy(x).bfloat16().float().numpy()
    assert np.array_equal(O.round_to_bf16(x), want)
    assert np.array_equal(O.bf16_bits_to_f32(O.f32_to_bf16_bits(x)), want)


@pytest.mark.parametrize("n_bits", [8, 16, 32, 64])
def test_pack_unpack_match_reference(golden, n_bits):
    g = golden("codec.npz")
    packed = O.pack(g["bits"], n_bits)
    assert packed.dtype == g[f"packed{n_bits}"].dtype
    assert np.array_equal(packed, g[f"packed{n_bits}"])
    assert np.array_equal(O.unpack(g[f"packed{n_bits}"], n_bits), g["bits"])


def test_pack_known_answers(golden):
    g = golden("codec.npz")
    assert np.array_equal(O.pack(g["kat_bits"]), g["kat_packed"])
    kp = g["kat_packed"][:, 0, 0]
    assert kp[0] == 1 and kp[31] == -2147483648 and kp[32] == -1 and kp[33] == 0x55555555
    for i in range(32):
        assert kp[i] == np.int32(np.uint32(1 << i).astype(np.uint32).view(np.int32))


def test_unpack_negative_words(golden):
    g = golden("codec.npz")
    assert np.array_equal(O.unpack(g["words"]), g["words_unpacked"])
    assert np.array_equal(O.pack(g["words_unpacked"]), g["words"])


def test_pack_asserts_like_reference():
    with pytest.raises(AssertionError, match="K must be divisible by n_bits"):
        O.pack(np.zeros((33, 4), bool))


def test_binarydiff_ctor(golden):
    g = golden("binarydiff_ctor.npz")
    mask, coeff = O.binarydiff_compress(bf(g["base"]), bf(g["finetune"]))
    assert np.array_equal(mask, g["mask"])  # bit-exact, incl. the diff == 0 -> +1 rows
    assert abs(float(coeff) - float(g["coeff"])) <= 1e-6 * float(g["coeff"])
    assert list(g["state_keys"]) == ["coeff", "mask", "base"]
    # rows 0/1 were forced to diff == 0: the reference maps them to bit 1
    signs = O.unpack(g["mask"])
    assert signs[:8, 0].all() and signs[:4, 1].all()


def test_kernel_rounding_chain_vs_triton_interpreter(golden):
    g = golden("triton_interp.npz")
    c = O.binary_matmul(g["a"].astype(np.float32), g["b"], out="fp16")
    assert np.array_equal(c.astype(np.float16), g["c"])  # the Triton body on CPU: fp32 acc -> fp16, bit-exact
    c3 = O.binary_bmm(g["a3"].astype(np.float32), g["b3"], out="fp16")
    assert np.array_equal(c3.astype(np.float16), g["c3"])
    exact = O.binary_bmm_exact(g["a3"].astype(np.float32), g["b3"])
    assert O.rel_mean_abs_err(g["c3"], exact) < 1e-3


def test_binarydiff_forward_vs_reference_cpu_path(golden):
    g = golden("binarydiff_forward.npz")
    base, x = bf(g["base"]), bf(g["x"])
    mask, coeff = O.binarydiff_compress(base, bf(g["finetune"]))
    assert np.array_equal(mask, g["mask"])
    exact = O.binarydiff_forward_exact(x, base, g["mask"], g["coeff"])
    # the reference's fp32 CPU evaluation of diff.py:39 agrees with the float64 truth to fp32 accuracy
    assert O.rel_mean_abs_err(g["y_f32"], exact) < 1e-6
    # its bf16 evaluation and the oracle's emulated rounding chain both sit within bf16 rounding of the truth
    assert O.rel_mean_abs_err(bf(g["y_bf16"]), exact) < 4e-3
    emu = O.binarydiff_forward(x, base, g["mask"], g["coeff"])
    assert O.rel_mean_abs_err(emu, exact) < 4e-3
    assert O.rel_mean_abs_err(emu, bf(g["y_bf16"])) < 4e-3


def test_demo_modules(golden):
    g = golden("demo_modules.npz")
    x, w = bf(g["x"]), bf(g["weight"])
    coeffs = bf(g["coeffs"])
    y = O.diffcompress_forward(x, w, g["masks"], coeffs)
    exact = O.diffcompress_forward_exact(x, w, g["masks"], coeffs)
    assert O.rel_mean_abs_err(bf(g["y"]), exact) < 4e-3
    # same rounding chain -> equal up to the order of fp32 accumulation inside the two GEMMs
    assert O.rel_mean_abs_err(y, bf(g["y"])) < 2e-3
    logits = O.dataparallel_forward(x, [bf(g["head_w0"]), bf(g["head_w1"]), bf(g["head_w2"])], "linear")
    ref_logits = bf(g["logits"])
    assert logits.shape == ref_logits.shape == (3, 2, 52)
    fill = ref_logits[0, :, 50:]
    assert np.all(fill == np.float32(-3.3895313892515355e38))  # finfo(bf16).min padding of the narrow vocab
    assert np.array_equal(logits[0, :, 50:], fill)
    assert O.rel_mean_abs_err(logits[:, :, :50], ref_logits[:, :, :50]) < 2e-3
    emb = O.dataparallel_forward(g["ids"], [bf(e) for e in g["emb_w"]], "embedding")
    assert np.array_equal(emb, bf(g["emb_out"]))


@pytest.mark.parametrize("tag", ["bf16", "fp16"])
@pytest.mark.parametrize("m", [1, 3])
def test_tenant_leaves(golden, tag, m):
    """DataParallelModule around lm_head (ragged vocab), embed_tokens and the HF RMSNorm, as run by the reference."""
    g = golden("tenant_leaves.npz")
    dec = bf if tag == "bf16" else (lambda b: np.asarray(b).view(np.float16).astype(np.float32))
    ulp = 2.0 ** -8 if tag == "bf16" else 2.0 ** -11
    pre = f"{tag}_m{m}_"
    x = dec(g[pre + "x"])
    ws = [dec(g[pre + f"head_w{t}"]) for t in range(3)]
    ref = dec(g[pre + "logits"])
    got = O.dataparallel_forward(x, ws, "linear", out=tag)
    assert got.shape == ref.shape == (3, m, 75)
    fill = np.float32(-3.3895313892515355e38) if tag == "bf16" else np.float32(-65504.0)
    for t, v in enumerate((70, 75, 64)):
        assert np.all(ref[t, :, v:] == fill) and np.array_equal(got[t, :, v:], ref[t, :, v:])
        assert np.all(np.abs(got[t, :, :v] - ref[t, :, :v]) <= 2 * ulp * np.abs(ref[t, :, :v]) + 1e-6)
    normed = O.dataparallel_forward(x, list(dec(g[pre + "norm_w"])), "rmsnorm", eps=float(g["eps"]), out=tag)
    refn = dec(g[pre + "normed"])
    assert np.all(np.abs(normed - refn) <= 2 * ulp * np.abs(refn) + 1e-6)
    assert (normed == refn).mean() > 0.98
    emb = O.dataparallel_forward(g[pre + "ids"], [dec(g[pre + f"emb_w{t}"]) for t in range(3)], "embedding", out=tag)
    assert np.array_equal(emb, dec(g[pre + "emb_out"]))


def test_fold_matches_load_diff(golden):
    import torch

    g = golden("tiny_llama.npz")
    import os

    d = torch.load(os.path.join(os.path.dirname(__file__), "golden", "tiny_llama_diff.pt"), weights_only=False)
    for name in ["model.layers.0.self_attn.q_proj", "model.layers.1.mlp.down_proj", "model.layers.1.self_attn.k_proj"]:
        base, fine = bf(g["base::" + name]), bf(g["fine::" + name])
        mask, coeff = O.binarydiff_compress(base, fine)
        assert np.array_equal(mask, d[name + ".mask"].numpy())
        assert abs(float(coeff) - d[name + ".coeff"].item()) <= 1e-6 * float(coeff)
        folded = O.fold_delta(base, d[name + ".mask"].numpy(), d[name + ".coeff"].item())
        assert np.array_equal(O.f32_to_bf16_bits(folded), g["folded::" + name])
    # full-precision leaves are replaced, not folded (diff.py:96-97)
    assert np.array_equal(g["folded::lm_head"], g["fine::lm_head"])


def test_diff_pt_format(golden):
    g = golden("tiny_llama.npz")
    keys, dtypes, shapes = list(g["keys"]), list(g["dtypes"]), list(g["shapes"])
    info = dict(zip(keys, zip(dtypes, shapes)))
    assert info["model.layers.0.self_attn.q_proj.mask"] == ("torch.int32", "(2, 64)")
    assert info["model.layers.0.mlp.down_proj.mask"] == ("torch.int32", "(4, 64)")
    assert info["model.layers.0.self_attn.q_proj.coeff"] == ("torch.float32", "()")
    assert info["lm_head.weight"][0] == "torch.bfloat16" and info["model.embed_tokens.weight"][0] == "torch.bfloat16"
    assert "model.norm.weight" in info and "model.layers.1.input_layernorm.weight" in info
    assert not any(k.endswith(".base") for k in keys)
    assert len(keys) == 2 * 7 * 2 + 2 + 2 * 2 + 1
