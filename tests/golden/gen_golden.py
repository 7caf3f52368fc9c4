#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the REFERENCE itself.

Run in the build container only (needs /root/reference; the GPU box has no reference):

    python tests/golden/gen_golden.py

What is executed is the reference's own code, imported from /root/reference -- nothing is copied:
  * bitdelta/binary_gemm_kernel.py  pack / unpack                      (CPU torch)
  * bitdelta/binary_gemm_kernel.py  binary_matmul_kernel / binary_bmm_kernel bodies under TRITON_INTERPRET=1 (fp16)
  * bitdelta/diff.py                BinaryDiff.__init__, compress_diff, save_diff, load_diff (CPU torch;
                                    `accelerate` is absent here and is stubbed -- it is only used by utils.get_model)
  * demo/demo_backend.py:62-98      DataParallelModule / DiffCompressModule class bodies, exec'd from the file's
                                    source lines with `binary_bmm` bound to the reference's CPU identity
                                    x @ (unpack(mask)*2-1) (diff.py:93 / notebook cell 7), because the Triton launcher
                                    needs a CUDA device.
bf16 tensors are stored as their raw uint16 bit patterns (numpy has no bf16).
"""
import os
import sys
import types

os.environ["TRITON_INTERPRET"] = "1"
REF = "/root/reference"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

import numpy as np
import torch
import torch.nn as nn
import transformers  # noqa: F401  (must be imported before the accelerate stub)

acc_stub = types.ModuleType("accelerate")
acc_stub.infer_auto_device_map = lambda *a, **k: None
acc_stub.init_empty_weights = lambda *a, **k: None
sys.modules.setdefault("accelerate", acc_stub)

import bitdelta.binary_gemm_kernel as ref_k  # noqa: E402
import bitdelta.diff as ref_diff  # noqa: E402


def bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.detach().contiguous().view(torch.int16).numpy().view(np.uint16).copy()


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print(f"wrote {path} ({os.path.getsize(path)} bytes)")


# ---------------------------------------------------------------- 1. codec
def gen_codec():
    g = torch.Generator().manual_seed(1234)
    out = {}
    bits = torch.rand(2, 3, 128, 40, generator=g) > 0.5
    out["bits"] = bits.numpy()
    for nb in (8, 16, 32, 64):
        p = ref_k.pack(bits, n_bits=nb)
        out[f"packed{nb}"] = p.numpy()
        assert torch.equal(ref_k.unpack(p, n_bits=nb), bits)
    # known answers: single bit i of a 32-group, all ones, alternating
    kat_bits = torch.zeros(34, 32, 1, dtype=torch.bool)
    for i in range(32):
        kat_bits[i, i, 0] = True
    kat_bits[32, :, 0] = True
    kat_bits[33, ::2, 0] = True
    out["kat_bits"] = kat_bits.numpy()
    out["kat_packed"] = ref_k.pack(kat_bits).numpy()
    # unpack of arbitrary words incl. negative ones
    words = torch.randint(-(2**31), 2**31 - 1, (3, 4, 17), generator=g, dtype=torch.int64).to(torch.int32)
    out["words"] = words.numpy()
    out["words_unpacked"] = ref_k.unpack(words).numpy()
    save("codec.npz", **out)


# ---------------------------------------------------------------- 2. BinaryDiff ctor
def gen_ctor():
    torch.manual_seed(0)
    N, K = 48, 96
    base = (torch.randn(N, K) * 0.02).bfloat16()
    fine = (base.float() + torch.randn(N, K) * 0.002).bfloat16()
    fine[0, :8] = base[0, :8]  # diff == +0  -> bit 1 (diff.py:14-15)
    fine[1, :4] = 0.0
    base[1, :4] = 0.0  # 0 - 0
    m = ref_diff.BinaryDiff(base.clone(), fine.clone())
    assert m.base.stride() == (1, K)
    save(
        "binarydiff_ctor.npz",
        base=bf16_bits(base),
        finetune=bf16_bits(fine),
        mask=m.mask.numpy(),
        coeff=np.float32(m.coeff.item()),
        state_keys=np.array(list(m.state_dict().keys())),
    )
    return base, fine, m


# ---------------------------------------------------------------- 3. Triton kernel bodies on the CPU interpreter
def gen_triton_interp():
    import triton

    torch.manual_seed(0)
    M, K, N = 24, 128, 96
    a = torch.randn(M, K, dtype=torch.float16)
    b = ref_k.pack(torch.randn(K, N) > 0.5)  # notebook cell 7 distribution
    c = torch.empty(M, N, dtype=torch.float16)
    meta = dict(BLOCK_SIZE_M=16, BLOCK_SIZE_N=32, BLOCK_SIZE_K=32, GROUP_SIZE_M=8, ACTIVATION="")
    grid = (triton.cdiv(M, 16) * triton.cdiv(N, 32),)
    ref_k.binary_matmul_kernel.fn[grid](
        a, b, c, M, N, K, 32, a.stride(0), a.stride(1), b.stride(0), b.stride(1), c.stride(0), c.stride(1), **meta
    )
    Bt = 3
    a3 = torch.randn(Bt, M, K, dtype=torch.float16)
    b3 = ref_k.pack(torch.randn(Bt, K, N) > 0.5)
    c3 = torch.empty(Bt, M, N, dtype=torch.float16)
    ref_k.binary_bmm_kernel.fn[(grid[0], Bt)](
        a3, b3, c3, M, N, K, 32,
        a3.stride(1), a3.stride(2), b3.stride(1), b3.stride(2), c3.stride(1), c3.stride(2),
        a3.stride(0), b3.stride(0), c3.stride(0), **meta,
    )
    save(
        "triton_interp.npz",
        a=a.numpy(), b=b.numpy(), c=c.numpy(),
        a3=a3.numpy(), b3=b3.numpy(), c3=c3.numpy(),
    )


# ---------------------------------------------------------------- 4. BinaryDiff forward, reference CPU unpack path
def gen_forward():
    torch.manual_seed(0)
    N, K = 160, 256
    base = (torch.randn(N, K) * 0.02).bfloat16()
    fine = (base.float() + torch.randn(N, K) * 0.002).bfloat16()
    m = ref_diff.BinaryDiff(base.clone(), fine.clone())
    x = torch.randn(2, 5, K).bfloat16()
    signs = ref_k.unpack(m.mask) * 2 - 1  # diff.py:93 identity
    with torch.no_grad():
        y_bf16 = x @ m.base + m.coeff * (x @ signs.to(torch.bfloat16))
        y_f32 = x.float() @ m.base.float() + m.coeff * (x.float() @ signs.float())
    save(
        "binarydiff_forward.npz",
        base=bf16_bits(base), finetune=bf16_bits(fine), x=bf16_bits(x),
        mask=m.mask.numpy(), coeff=np.float32(m.coeff.item()),
        y_bf16=bf16_bits(y_bf16), y_f32=y_f32.numpy(),
    )


# ---------------------------------------------------------------- 5. demo backend modules
def load_demo_classes():
    src = open(os.path.join(REF, "demo/demo_backend.py")).read().splitlines()
    body = "\n".join(src[61:179])  # lines 62-179: the two module classes + register/unregister + DiffCompress

    def cpu_binary_bmm(a, b, n_bits=32, activation=""):
        s = (ref_k.unpack(b) * 2 - 1).to(torch.float32)
        return torch.bmm(a.float(), s).to(torch.float16).to(a.dtype)  # fp32 acc -> fp16 -> a.dtype (kernel :260,:287,:314)

    import gc

    ns = dict(torch=torch, nn=nn, gc=gc, binary_bmm=cpu_binary_bmm)
    exec(compile(body, "demo_backend_slice", "exec"), ns)
    return ns


def gen_demo():
    ns = load_demo_classes()
    torch.manual_seed(0)
    T, m, K, N = 3, 2, 128, 96
    lin = nn.Linear(K, N, bias=False).to(torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_((torch.randn(N, K) * 0.02).bfloat16())
    masks = ref_k.pack(torch.rand(T, K, N) > 0.5)
    coeffs = (torch.rand(T) * 0.004).bfloat16()
    x = torch.randn(T, m, K).bfloat16()
    mod = ns["DiffCompressModule"](lin, masks, coeffs)
    with torch.no_grad():
        y = mod(x)
    # DataParallelModule with ragged vocab
    head = nn.Linear(K, 50, bias=False).to(torch.bfloat16)
    ws = [(torch.randn(v, K) * 0.05).bfloat16() for v in (50, 52, 52)]
    dp = ns["DataParallelModule"](head, ws)
    with torch.no_grad():
        logits = dp(x)
    emb = nn.Embedding(52, K).to(torch.bfloat16)
    ews = [(torch.randn(52, K)).bfloat16() for _ in range(T)]
    ids = torch.randint(0, 50, (T, m))
    dpe = ns["DataParallelModule"](emb, ews)
    with torch.no_grad():
        e = dpe(ids)
    save(
        "demo_modules.npz",
        weight=bf16_bits(lin.weight), masks=masks.numpy(), coeffs=bf16_bits(coeffs), x=bf16_bits(x), y=bf16_bits(y),
        head_w0=bf16_bits(ws[0]), head_w1=bf16_bits(ws[1]), head_w2=bf16_bits(ws[2]), logits=bf16_bits(logits),
        emb_w=np.stack([bf16_bits(w) for w in ews]), ids=ids.numpy(), emb_out=bf16_bits(e),
    )


# ---------------------------------------------------------------- 6. diff.pt through the reference's save_diff / load_diff
def gen_diff_pt():
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg = LlamaConfig(
        hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
        num_key_value_heads=2, vocab_size=96, max_position_embeddings=64, tie_word_embeddings=False,
    )
    torch.manual_seed(0)
    base = LlamaForCausalLM(cfg).to(torch.bfloat16)
    fine = LlamaForCausalLM(cfg).to(torch.bfloat16)
    fine.load_state_dict(base.state_dict())
    with torch.no_grad():
        for n, p in fine.named_parameters():
            p.add_((torch.randn_like(p.float()) * 0.002).to(p.dtype))
    comp = LlamaForCausalLM(cfg).to(torch.bfloat16)
    comp.load_state_dict(fine.state_dict())
    ref_diff.compress_diff(base, fine, comp)
    path = os.path.join(HERE, "tiny_llama_diff.pt")
    ref_diff.save_diff(comp, path)
    d = torch.load(path, weights_only=False)
    keys = list(d.keys())
    folded = LlamaForCausalLM(cfg).to(torch.bfloat16)
    folded.load_state_dict(base.state_dict())
    ref_diff.load_diff(folded, path)
    sel = ["model.layers.0.self_attn.q_proj", "model.layers.1.mlp.down_proj", "model.layers.1.self_attn.k_proj"]
    arrs = {}
    for s in sel:
        arrs["base::" + s] = bf16_bits(base.get_submodule(s).weight)
        arrs["fine::" + s] = bf16_bits(fine.get_submodule(s).weight)
        arrs["folded::" + s] = bf16_bits(folded.get_submodule(s).weight)
    arrs["folded::lm_head"] = bf16_bits(folded.lm_head.weight)
    arrs["fine::lm_head"] = bf16_bits(fine.lm_head.weight)
    # full base state so tests can rebuild the model without depending on HF init
    for k, v in base.state_dict().items():
        arrs["basesd::" + k] = bf16_bits(v) if v.dtype == torch.bfloat16 else v.numpy()
    # the compressed model's logits on fixed tokens (reference module forward needs CUDA -> use fold path)
    ids = torch.randint(0, 96, (2, 12))
    with torch.no_grad():
        arrs["folded_logits"] = folded(ids).logits.float().numpy()
    arrs["ids"] = ids.numpy()
    save("tiny_llama.npz", keys=np.array(keys), dtypes=np.array([str(d[k].dtype) for k in keys]),
         shapes=np.array([str(tuple(d[k].shape)) for k in keys]), **arrs)
    print("diff.pt keys:", len(keys), keys[:4], "...")


# ---------------------------------------------------------------- 7. per-tenant dense leaves (DataParallelModule, demo_backend.py:62-79)
def gen_tenant_leaves():
    """The reference's DataParallelModule around the three leaf kinds it wraps in the demo: lm_head (ragged vocab),
    embed_tokens and the HF RMSNorm (transformers' LlamaRMSNorm, the class the decoder layers instantiate)."""
    from transformers.models.llama.modeling_llama import LlamaRMSNorm

    ns = load_demo_classes()
    arrs = {}
    for tag, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        torch.manual_seed(7)
        T, K = 3, 256
        vocab = (70, 75, 64)
        for m in (1, 3):
            x = torch.randn(T, m, K).to(dt)
            head = nn.Linear(K, vocab[0], bias=False).to(dt)
            ws = [(torch.randn(v, K) * 0.05).to(dt) for v in vocab]
            with torch.no_grad():
                logits = ns["DataParallelModule"](head, ws)(x)
            norm = LlamaRMSNorm(K, eps=1e-5).to(dt)
            nws = [(1.0 + 0.1 * torch.randn(K)).to(dt) for _ in range(T)]
            with torch.no_grad():
                normed = ns["DataParallelModule"](norm, nws)(x)
            emb = nn.Embedding(max(vocab), K).to(dt)
            ews = [torch.randn(v, K).to(dt) for v in vocab]
            ids = torch.stack([torch.randint(0, v, (m,)) for v in vocab])
            with torch.no_grad():
                e = ns["DataParallelModule"](emb, ews)(ids)
            pre = f"{tag}_m{m}_"
            arrs[pre + "x"] = bf16_bits(x)
            for t in range(T):
                arrs[pre + f"head_w{t}"] = bf16_bits(ws[t])
                arrs[pre + f"emb_w{t}"] = bf16_bits(ews[t])
            arrs[pre + "logits"] = bf16_bits(logits)
            arrs[pre + "norm_w"] = np.stack([bf16_bits(w) for w in nws])
            arrs[pre + "normed"] = bf16_bits(normed)
            arrs[pre + "ids"] = ids.numpy()
            arrs[pre + "emb_out"] = bf16_bits(e)
    arrs["eps"] = np.float32(1e-5)
    save("tenant_leaves.npz", **arrs)


GENERATORS = dict(codec=gen_codec, ctor=gen_ctor, triton_interp=gen_triton_interp, forward=gen_forward, demo=gen_demo,
                  diff_pt=gen_diff_pt, tenant_leaves=gen_tenant_leaves)

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(GENERATORS)):
        GENERATORS[name]()
