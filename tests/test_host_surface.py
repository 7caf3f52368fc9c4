"""Host-side behaviour of the reference-compatible Python surface (no GPU needed)."""
import os

import numpy as np
import pytest
import torch

import bitdelta_b200 as bd
from oracle import bitdelta_oracle as O


def t_bf16(bits):
    return torch.from_numpy(bits.view(np.int16).copy()).view(torch.bfloat16)


def test_reference_names_and_signatures():
    import inspect

    assert list(inspect.signature(bd.pack).parameters) == ["x", "n_bits"]
    assert list(inspect.signature(bd.unpack).parameters) == ["x", "n_bits"]
    assert list(inspect.signature(bd.binary_matmul).parameters)[:4] == ["a", "b", "n_bits", "activation"]
    assert list(inspect.signature(bd.binary_bmm).parameters)[:4] == ["a", "b", "n_bits", "activation"]
    assert list(inspect.signature(bd.BinaryDiff.__init__).parameters) == ["self", "base", "finetune"]
    assert list(inspect.signature(bd.DiffCompressModule.__init__).parameters) == ["self", "module", "mask_list", "coeff_list"]
    assert list(inspect.signature(bd.DataParallelModule.__init__).parameters) == ["self", "module", "weight_list"]
    assert list(inspect.signature(bd.register_diff_compress).parameters) == ["model", "checkpoint_list"]
    for name in ["compress_diff", "save_diff", "load_diff", "save_full_model", "unregister_diff_compress", "DiffCompress"]:
        assert hasattr(bd, name)


@pytest.mark.parametrize("n_bits", [8, 16, 32, 64])
def test_host_codec_matches_reference_vectors(golden, n_bits):
    g = golden("codec.npz")
    bits = torch.from_numpy(g["bits"])
    packed = bd.pack(bits, n_bits)
    assert packed.dtype == {8: torch.uint8, 16: torch.int16, 32: torch.int32, 64: torch.int64}[n_bits]
    assert np.array_equal(packed.numpy(), g[f"packed{n_bits}"])
    assert torch.equal(bd.unpack(packed, n_bits), bits)


def test_host_codec_fuzz_against_oracle():
    rng = np.random.default_rng(0)
    for _ in range(25):
        lead = tuple(rng.integers(1, 3, size=rng.integers(0, 3)))
        J, N = int(rng.integers(1, 5)), int(rng.integers(1, 70))
        bits = rng.random(lead + (J * 32, N)) > rng.random()
        packed = bd.pack(torch.from_numpy(bits))
        assert np.array_equal(packed.numpy(), O.pack(bits))
        assert np.array_equal(bd.unpack(packed).numpy(), bits)


def test_pack_assertion_and_transposed_input():
    with pytest.raises(AssertionError, match="K must be divisible by n_bits"):
        bd.pack(torch.zeros(33, 4, dtype=torch.bool))
    x = torch.rand(96, 64) > 0.5
    assert np.array_equal(bd.pack(x.T).numpy(), O.pack(x.T.numpy()))
    assert bd.pack(torch.zeros(0, 32, 5, dtype=torch.bool)).shape == (0, 1, 5)


def test_gemm_refuses_cpu_tensors_loudly():
    a = torch.zeros(1, 1, 32, dtype=torch.bfloat16)
    b = torch.zeros(1, 1, 4, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bd.binary_bmm(a, b)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bd.binary_matmul(a[0], b[0])
    m = bd.BinaryDiff(torch.zeros(8, 32, dtype=torch.bfloat16), torch.ones(8, 32, dtype=torch.bfloat16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 1, 32, dtype=torch.bfloat16))
    with pytest.raises(AssertionError, match="Matrix A must be 3D"):
        bd.binary_bmm(a[0], b)
    with pytest.raises(AssertionError, match="Incompatible dimensions"):
        bd.binary_bmm(torch.zeros(1, 1, 64, dtype=torch.bfloat16), b)


def test_binarydiff_ctor_on_host_matches_reference(golden):
    g = golden("binarydiff_ctor.npz")
    m = bd.BinaryDiff(t_bf16(g["base"]), t_bf16(g["finetune"]))
    assert np.array_equal(m.mask.numpy(), g["mask"])
    assert abs(m.coeff.item() - float(g["coeff"])) <= 1e-6 * float(g["coeff"])
    assert list(m.state_dict().keys()) == list(g["state_keys"])
    assert m.base.stride() == (1, 96) and m.coeff.requires_grad and m.coeff.dtype == torch.float32


def _tiny_models(golden):
    from transformers import LlamaConfig, LlamaForCausalLM

    g = golden("tiny_llama.npz")
    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=96, max_position_embeddings=64, tie_word_embeddings=False)
    sd = {}
    for k in g.files:
        if k.startswith("basesd::"):
            v = g[k]
            sd[k[len("basesd::"):]] = t_bf16(v) if v.dtype == np.uint16 else torch.from_numpy(v)
    base = LlamaForCausalLM(cfg).to(torch.bfloat16)
    base.load_state_dict(sd)
    return g, cfg, base


def test_diff_pt_round_trip_against_reference_file(golden, tmp_path):
    """compress_diff + save_diff write the same dict the reference wrote; load_diff folds it to the same weights."""
    from transformers import LlamaForCausalLM

    g, cfg, base = _tiny_models(golden)
    ref = torch.load(os.path.join(os.path.dirname(__file__), "golden", "tiny_llama_diff.pt"), weights_only=False)
    # rebuild the fine-tuned model from the reference diff: full leaves from the file, projections from base + stored fine slabs
    fine = LlamaForCausalLM(cfg).to(torch.bfloat16)
    fine.load_state_dict(base.state_dict())
    folded = LlamaForCausalLM(cfg).to(torch.bfloat16)
    folded.load_state_dict(base.state_dict())
    bd.load_diff(folded, os.path.join(os.path.dirname(__file__), "golden", "tiny_llama_diff.pt"))
    for name in ["model.layers.0.self_attn.q_proj", "model.layers.1.mlp.down_proj", "model.layers.1.self_attn.k_proj"]:
        got = folded.get_submodule(name).weight.detach().view(torch.int16).numpy().view(np.uint16)
        assert np.array_equal(got, g["folded::" + name]), name
    assert np.array_equal(folded.lm_head.weight.detach().view(torch.int16).numpy().view(np.uint16), g["folded::lm_head"])
    ids = torch.from_numpy(g["ids"])
    with torch.no_grad():
        logits = folded(ids).logits.float().numpy()
    assert np.allclose(logits, g["folded_logits"], rtol=0, atol=1e-6)

    # now the writer: compress a model whose projections carry the reference's fine weights where we have them
    with torch.no_grad():
        for name in ["model.layers.0.self_attn.q_proj", "model.layers.1.mlp.down_proj", "model.layers.1.self_attn.k_proj"]:
            fine.get_submodule(name).weight.copy_(t_bf16(g["fine::" + name]))
    comp = LlamaForCausalLM(cfg).to(torch.bfloat16)
    comp.load_state_dict(fine.state_dict())
    bd.compress_diff(base, fine, comp)
    out = tmp_path / "diff.pt"
    bd.save_diff(comp, str(out))
    mine = torch.load(str(out), weights_only=False)
    assert list(mine.keys()) == list(ref.keys())
    for k in ref:
        assert mine[k].dtype == ref[k].dtype and mine[k].shape == ref[k].shape, k
    for name in ["model.layers.0.self_attn.q_proj", "model.layers.1.mlp.down_proj", "model.layers.1.self_attn.k_proj"]:
        assert torch.equal(mine[name + ".mask"], ref[name + ".mask"])
        assert abs(mine[name + ".coeff"].item() - ref[name + ".coeff"].item()) <= 1e-6 * ref[name + ".coeff"].item()


def test_register_unregister_on_host_modules():
    import torch.nn as nn

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.q_proj = nn.Linear(64, 32, bias=False)
            self.norm = nn.LayerNorm(64, bias=False)

    model = M()
    orig_w = model.norm.weight.data
    ckpts = [{"q_proj.mask": bd.pack(torch.rand(64, 32) > 0.5), "q_proj.coeff": torch.tensor(0.1 * (i + 1)),
              "norm.weight": torch.full((64,), float(i))} for i in range(2)]
    bd.demo_backend.cached_modules.clear()
    bd.register_diff_compress(model, ckpts)
    assert isinstance(model.q_proj, bd.DiffCompressModule) and isinstance(model.norm, bd.DataParallelModule)
    assert model.q_proj.mask.shape == (2, 2, 32) and model.q_proj.coeff.shape == (2,)
    assert "q_proj.mask" not in ckpts[0] and "norm.weight" in ckpts[0]
    out = model.norm(torch.ones(2, 3, 64) * torch.arange(64))
    assert out.shape == (2, 3, 64) and torch.all(out[0] == 0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.q_proj(torch.zeros(2, 1, 64))
    bd.unregister_diff_compress(model)
    assert isinstance(model.q_proj, nn.Linear) and isinstance(model.norm, nn.LayerNorm)
    assert model.norm.weight.data.data_ptr() == orig_w.data_ptr()
    bd.demo_backend.cached_modules.clear()
    with pytest.raises(AssertionError, match="Only support linear layer"):
        bd.register_diff_compress(M(), [{"norm.mask": torch.zeros(2, 64, dtype=torch.int32), "norm.coeff": torch.tensor(1.0)}])
    bd.demo_backend.cached_modules.clear()


@pytest.mark.parametrize("tag", ["bf16", "fp16"])
def test_dataparallel_host_path_matches_reference_vectors(golden, tag):
    """On host tensors DataParallelModule runs the reference's generic loop (the native kernels are CUDA-only): its
    outputs must equal the reference module's (tests/golden/gen_golden.py::gen_tenant_leaves) bit for bit."""
    from transformers.models.llama.modeling_llama import LlamaRMSNorm

    g = golden("tenant_leaves.npz")
    dt = torch.bfloat16 if tag == "bf16" else torch.float16
    t16 = lambda b: torch.from_numpy(b.view(np.int16).copy()).view(dt)  # noqa: E731
    bits = lambda t: t.contiguous().view(torch.int16).numpy().view(np.uint16)  # noqa: E731
    for m in (1, 3):
        pre = f"{tag}_m{m}_"
        x = t16(g[pre + "x"])
        head = torch.nn.Linear(256, 70, bias=False).to(dt)
        with torch.no_grad():
            logits = bd.DataParallelModule(head, [t16(g[pre + f"head_w{t}"]) for t in range(3)])(x)
            normed = bd.DataParallelModule(LlamaRMSNorm(256, eps=float(g["eps"])).to(dt), [t16(w) for w in g[pre + "norm_w"]])(x)
            emb = bd.DataParallelModule(torch.nn.Embedding(75, 256).to(dt), [t16(g[pre + f"emb_w{t}"]) for t in range(3)])(
                torch.from_numpy(g[pre + "ids"]))
        assert np.array_equal(bits(logits), g[pre + "logits"])
        assert np.array_equal(bits(normed), g[pre + "normed"])
        assert np.array_equal(bits(emb), g[pre + "emb_out"])


def test_binarydiff_weight_view_cache_also_for_inference_tensors():
    """BinaryDiff keeps the reference's base = W.T view; if a state_dict load made `base` a contiguous [K, N] buffer, the
    [N, K] operand the kernel needs is rebuilt once and cached -- also under torch.inference_mode (no version counter)."""
    base = (torch.randn(48, 96) * 0.02).bfloat16()
    fine = (base.float() + torch.randn(48, 96) * 0.002).bfloat16()
    m = bd.BinaryDiff(base, fine)
    assert m._weight_nk().data_ptr() == base.data_ptr()  # zero-copy while the strides are the reference's
    m.base = m.base.contiguous()
    w = m._weight_nk()
    assert w.is_contiguous() and torch.equal(w, base) and m._weight_nk() is w
    with torch.inference_mode():
        m2 = bd.BinaryDiff(base, fine)
        m2.base = m2.base.contiguous()
        assert torch.equal(m2._weight_nk(), base)


def test_register_unregister_on_nested_hf_model(golden):
    """The same wrap / unwrap over a real (2-layer) HF Llama and the reference's own diff.pt, on the host: every projection
    becomes a DiffCompressModule with stacked [T, K/32, N] masks, every full-precision leaf a DataParallelModule, and
    unregister restores the original modules and weights; a second register re-uses the cached stacks."""
    from bitdelta_b200 import demo_backend as db

    g, cfg, model = _tiny_models(golden)
    path = os.path.join(os.path.dirname(__file__), "golden", "tiny_llama_diff.pt")
    before = {n: p.data_ptr() for n, p in model.named_parameters()}
    db.cached_modules.clear()
    try:
        for round_ in range(2):
            ckpts = [torch.load(path, weights_only=False) for _ in range(3)]
            db.register_diff_compress(model, ckpts)
            assert db.fuse_sibling_projections(model) == 4
            kinds = {n: type(m).__name__ for n, m in model.named_modules()}
            assert kinds["model.layers.1.mlp.down_proj"] == "DiffCompressModule" and kinds["lm_head"] == "DataParallelModule"
            assert kinds["model.embed_tokens"] == "DataParallelModule" and kinds["model.layers.0.input_layernorm"] == "DataParallelModule"
            q = model.model.layers[0].self_attn.q_proj
            assert q.mask.shape == (3, 2, 64) and q.coeff.shape == (3,) and q._group is not None
            assert len(db.cached_modules) == 14
            # entries are removed from the tenant dicts only when they are stacked, i.e. on the first register (reference :131-141)
            assert ("model.layers.0.self_attn.q_proj.mask" in ckpts[0]) == (round_ == 1)
            db.unregister_diff_compress(model)
            assert isinstance(model.lm_head, torch.nn.Linear) and isinstance(model.model.layers[0].self_attn.q_proj, torch.nn.Linear)
            assert {n: p.data_ptr() for n, p in model.named_parameters()} == before
            if round_ == 0:  # the reference's cache: the next register finds the stacks although the fresh dicts carry them again
                assert set(db.cached_modules) == {n for n, k in kinds.items() if k == "DiffCompressModule"}
    finally:
        db.cached_modules.clear()


def test_load_checkpoints_stacked_equals_load_then_register(golden):
    """f-1: the memory-mapped loader puts every tenant's sign words straight into the stacked [T, K/32, N] tensor and
    publishes it in ``cached_modules``; registering from it builds exactly the modules the reference's load-to-CPU,
    move, stack, pop sequence builds (demo_backend.py:26-35, :131-141)."""
    from bitdelta_b200 import demo_backend as db

    g, cfg, model = _tiny_models(golden)
    path = os.path.join(os.path.dirname(__file__), "golden", "tiny_llama_diff.pt")
    db.cached_modules.clear()
    try:
        ckpts = []
        for _ in range(3):
            d = torch.load(path, weights_only=False)
            ckpts.append({k: (v.detach().to(torch.bfloat16) if v.is_floating_point() else v) for k, v in d.items()})
        db.register_diff_compress(model, ckpts)
        want = {n: (m.mask.clone(), m.coeff.clone()) for n, m in model.named_modules() if isinstance(m, db.DiffCompressModule)}
        want_w = {n: [w.clone() for w in m.weight_list] for n, m in model.named_modules() if isinstance(m, db.DataParallelModule)}
        db.unregister_diff_compress(model)
        db.cached_modules.clear()
        rest = db.load_checkpoints_stacked([path] * 3, "cpu", torch.bfloat16)
        assert set(db.cached_modules) == set(want) and len(rest) == 3
        assert not any(k.endswith((".mask", ".coeff")) for k in rest[0])
        db.register_diff_compress(model, rest)
        for n, m in model.named_modules():
            if isinstance(m, db.DiffCompressModule):
                assert torch.equal(m.mask, want[n][0]) and torch.equal(m.coeff, want[n][1]) and m.mask.is_contiguous()
            elif isinstance(m, db.DataParallelModule):
                assert all(torch.equal(a, b) for a, b in zip(m.weight_list, want_w[n]))
        assert {n for n, m in model.named_modules() if isinstance(m, db.DiffCompressModule)} == set(want)
        db.unregister_diff_compress(model)
    finally:
        db.cached_modules.clear()
