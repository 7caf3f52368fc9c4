"""Drop-in for the reference module ``bitdelta/diff.py``: ``BinaryDiff`` and the ``diff.pt`` glue.

    BinaryDiff(base, finetune)        reference diff.py:8-39   (buffers mask/base, parameter coeff, forward(x))
    compress_diff / save_diff / load_diff / save_full_model   reference diff.py:41-116

``BinaryDiff.forward`` is ONE launch of the fused sm_100a kernel
``y = x.W_base^T + coeff * (x . (2*unpack(mask)-1))`` instead of the reference's mask.repeat copy + cuBLAS GEMM +
Triton bmm + two pointwise kernels (diff.py:38-39).  The sign matrix is never repeated per batch row: the kernel
broadcasts it (tenant stride 0).
"""
from __future__ import annotations

import gc

import torch
import torch.nn as nn

from . import _lib
from .binary_gemm_kernel import pack, unpack  # noqa: F401  (re-exported like the reference module)


def _fused_forward(x: torch.Tensor, w_nk: torch.Tensor, masks: torch.Tensor, coeff: torch.Tensor, T: int, kernel="auto",
                   static_operands: bool = False, out_fp32: bool = False) -> torch.Tensor:
    """x: (T, m, K) contiguous; w_nk: (N, K) contiguous; masks: (K/32, N) or (T, K/32, N) int32; coeff: (T,) or 0-dim.

    ``static_operands``: w_nk and masks are long-lived buffers that nothing queued on the stream is still writing
    (BD_FLAG_STATIC_OPERANDS: the launch may prefetch them while the stream's preceding kernel drains).  ``out_fp32``:
    return the unrounded fp32 sums (BD_FLAG_FP32_OUT; tensor-parallel partials)."""
    if not x.is_cuda:
        raise RuntimeError(
            "bitdelta_b200: the fused BinaryDiff forward is implemented only as sm_100a CUDA kernels "
            f"(libbitdelta_b200.so); got activations on {x.device}. There is no CPU fallback."
        )
    m, K = x.shape[1], x.shape[2]
    N = w_nk.shape[0]
    assert w_nk.shape[1] == K and masks.shape[-2] * 32 == K and masks.shape[-1] == N, "Incompatible dimensions"
    assert x.dtype == w_nk.dtype, "activations and base weight must share a dtype"
    assert masks.dtype == torch.int32 and masks.is_contiguous()
    assert x.device == w_nk.device == masks.device == coeff.device, "A and B must be on the same device"
    stride = masks.shape[-2] * masks.shape[-1] if masks.dim() == 3 else 0
    if masks.dim() == 3:
        assert masks.shape[0] == T, "Incompatible batch dimensions"
    if coeff.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        coeff = coeff.float()
    coeff = coeff.detach().reshape(-1)
    if coeff.numel() == 1 and T > 1:
        coeff = coeff.expand(T)
    coeff = coeff.contiguous()
    assert coeff.numel() == T
    y = torch.empty((T, m, N), device=x.device, dtype=torch.float32 if out_fp32 else x.dtype)
    if y.numel() == 0:
        return y
    with torch.cuda.device(x.device):
        ws = _lib.workspace(x.device, T * m, N)
        _lib.check(
            _lib.lib.bd_binarydiff_fwd_batched(
                x.data_ptr(), w_nk.data_ptr(), masks.data_ptr(), coeff.data_ptr(), _lib.dtype_code(coeff.dtype), y.data_ptr(),
                _lib.dtype_code(x.dtype), T, m, K, N, stride, ws.data_ptr(), ws.numel(),
                _lib.kernel_code(kernel, static_operands, out_fp32), _lib.stream_ptr(x.device),
            )
        )
    return y


def _fused_forward_grouped(x: torch.Tensor, weights, masks_list, coeffs, T: int, kernel="auto", static_operands: bool = False):
    """Several BinaryDiff linears on the SAME activations in one launch (q/k/v, gate/up).  x: (T, m, K); weights[s]: (N_s, K);
    masks_list[s]: (T, K/32, N_s) int32; coeffs[s]: (T,).  Returns [y_s (T, m, N_s)].  ``static_operands`` as in _fused_forward."""
    import ctypes

    if not x.is_cuda:
        raise RuntimeError(f"bitdelta_b200: grouped fused forward needs CUDA tensors (got {x.device}). There is no CPU fallback.")
    nseg = len(weights)
    assert 1 <= nseg <= 3 and len(masks_list) == nseg and len(coeffs) == nseg
    m, K = x.shape[1], x.shape[2]
    cdt = coeffs[0].dtype if coeffs[0].dtype in (torch.float32, torch.bfloat16, torch.float16) else torch.float32
    cs = []
    for c in coeffs:
        c = c.detach().to(cdt).reshape(-1)
        if c.numel() == 1 and T > 1:
            c = c.expand(T)
        cs.append(c.contiguous())
    ys, strides, Ns = [], [], []
    for w, mk in zip(weights, masks_list):
        assert w.is_contiguous() and w.shape[1] == K and w.dtype == x.dtype and mk.dtype == torch.int32 and mk.is_contiguous()
        assert mk.shape[-2] * 32 == K and mk.shape[-1] == w.shape[0] and (mk.dim() == 2 or mk.shape[0] == T), "Incompatible dimensions"
        Ns.append(w.shape[0])
        strides.append(mk.shape[-2] * mk.shape[-1] if mk.dim() == 3 else 0)
        ys.append(torch.empty((T, m, w.shape[0]), device=x.device, dtype=x.dtype))
    if x.numel() == 0:
        return ys
    vp = ctypes.c_void_p
    arr = lambda ts: (vp * nseg)(*[t.data_ptr() for t in ts])  # noqa: E731
    i64 = lambda vs: (ctypes.c_int64 * nseg)(*vs)  # noqa: E731
    with torch.cuda.device(x.device):
        ws = _lib.workspace(x.device, T * m, max(Ns))
        _lib.check(
            _lib.lib.bd_binarydiff_fwd_grouped(
                x.data_ptr(), nseg, arr(weights), arr(masks_list), arr(cs), _lib.dtype_code(cdt), arr(ys), i64(Ns), i64(strides),
                _lib.dtype_code(x.dtype), T, m, K, ws.data_ptr(), ws.numel(), _lib.kernel_code(kernel, static_operands),
                _lib.stream_ptr(x.device),
            )
        )
    return ys


class _BinaryDiffFunction(torch.autograd.Function):
    """Scale distillation through the fused op (reference train.py:60-97 trains ``coeff`` through ``BinaryDiff.forward``).

    The reference's ``binary_bmm`` output has no ``grad_fn``, so autograd sees ``y = x @ base + coeff * D`` with ``D`` a
    constant:  d/dcoeff = sum(grad_y * D),  d/dx = grad_y @ base.T  (the delta term does NOT back-propagate into x).
    ``delta_grad_x=True`` adds the true term ``coeff * grad_y @ sign.T`` instead of dropping it.
    """

    @staticmethod
    def forward(ctx, rows, w_nk, mask, coeff, kernel, delta_grad_x):
        ctx.save_for_backward(rows, w_nk, mask, coeff)
        ctx.kernel, ctx.delta_grad_x = kernel, delta_grad_x
        return _fused_forward(rows, w_nk, mask, coeff, 1, kernel)

    @staticmethod
    def backward(ctx, grad_y):
        from .binary_gemm_kernel import binary_bmm, unpack

        rows, w_nk, mask, coeff = ctx.saved_tensors
        grad_y = grad_y.contiguous()
        grad_rows = grad_coeff = None
        if ctx.needs_input_grad[3]:
            d = binary_bmm(rows, mask[None])  # x . sign, fp32 accumulate, one rounding (the reference's Delta-branch output)
            grad_coeff = (grad_y.float() * d.float()).sum().to(coeff.dtype).reshape(coeff.shape)
        if ctx.needs_input_grad[0]:
            grad_rows = torch.matmul(grad_y, w_nk)  # plain library GEMM: [1,M,N] @ [N,K]
            if ctx.delta_grad_x:
                sign = (unpack(mask).to(rows.dtype) * 2 - 1)  # [K, N]
                grad_rows = grad_rows + (coeff.float() * torch.matmul(grad_y, sign.t()).float()).to(rows.dtype)
        return grad_rows, None, None, grad_coeff, None, None


class BinaryDiff(nn.Module):
    """16-bit base weight + 1-bit delta linear layer (reference ``BinaryDiff``, diff.py:8-39).

    State: ``mask`` int32 ``[K/32, N]`` (buffer), ``base`` = ``W_base.T`` ``[K, N]`` stride ``(1, K)`` view (buffer),
    ``coeff`` fp32 0-dim ``nn.Parameter`` -- identical names, shapes and dtypes, so ``state_dict`` and ``save_diff``
    output are interchangeable with the reference's.
    """

    kernel = "auto"
    delta_grad_x = False  # reference behaviour: the delta branch carries no gradient w.r.t. the activations

    def __init__(self, base: torch.Tensor, finetune: torch.Tensor):
        super().__init__()
        assert base.shape == finetune.shape and base.dim() == 2
        N, K = base.shape
        assert K % 32 == 0, "K must be divisible by n_bits"
        if base.is_cuda and base.dtype in (torch.bfloat16, torch.float16):
            base_c = base.contiguous()
            fine_c = finetune.to(base.device, base.dtype).contiguous()
            mask = torch.empty((K // 32, N), dtype=torch.int32, device=base.device)
            quantile = torch.empty((), dtype=torch.float32, device=base.device)
            scratch = torch.zeros((), dtype=torch.float64, device=base.device)
            with torch.cuda.device(base.device):
                _lib.check(
                    _lib.lib.bd_compress(
                        base_c.data_ptr(), fine_c.data_ptr(), _lib.dtype_code(base.dtype), N, K, mask.data_ptr(),
                        quantile.data_ptr(), scratch.data_ptr(), _lib.stream_ptr(base.device),
                    )
                )
            base = base_c
        else:
            # host tensors (or fp32 weights): same arithmetic as diff.py:11-16, bits packed by the library's host codec
            diff = finetune - base
            quantile = diff.float().abs().mean()
            mask = pack((~(diff < 0)).T)
        self.register_buffer("mask", mask)
        self.register_buffer("base", base.T)
        self.register_parameter(
            "coeff",
            nn.Parameter(quantile.detach().clone().to(dtype=torch.float32, device=base.device).requires_grad_(True)),
        )
        self._w_cache = None

    def _weight_nk(self, with_static: bool = False):
        """[N, K] row-major view of the base weight (zero-copy while ``base`` keeps the reference's (1, K) strides).
        ``with_static`` also returns whether the tensor is a long-lived buffer (False right after a copy was made: the
        copy kernel is still queued on the stream, so the forward must not prefetch it early)."""
        w = self.base.t()
        static = True
        if not w.is_contiguous():
            key = (self.base.data_ptr(), 0 if self.base.is_inference() else self.base._version, self.base.device)
            if self._w_cache is None or self._w_cache[0] != key:
                self._w_cache = (key, w.contiguous())
                static = False
            w = self._w_cache[1]
        return (w, static) if with_static else w

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        # [B, seq, in] @ [in, out] + coeff * ([B, seq, in] @ sign[in, out])   (diff.py:33-39)
        lead = x.shape[:-1]
        rows = x.reshape(1, -1, x.shape[-1]).contiguous()
        if torch.is_grad_enabled() and (rows.requires_grad or self.coeff.requires_grad):
            y = _BinaryDiffFunction.apply(rows, self._weight_nk(), self.mask, self.coeff, self.kernel, self.delta_grad_x)
        else:
            w, static = self._weight_nk(with_static=True)
            y = _fused_forward(rows, w, self.mask, self.coeff, 1, self.kernel, static_operands=static)
        return y.reshape(*lead, y.shape[-1])


def _projection_sites(model):
    """(parent module, parent path, child name) of every leaf the reference compresses: children whose name contains
    ``proj`` under modules whose path contains ``mlp`` or ``self_attn`` (reference diff.py:60-64), in traversal order."""
    sites = []
    for path, parent in model.named_modules():
        if "mlp" not in path and "self_attn" not in path:
            continue
        sites.extend((parent, path, child) for child, _ in parent.named_children() if "proj" in child)
    return sites


def _release_cached_memory():
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def compress_diff(base_model, finetuned_model, finetuned_compressed_model):
    """Replace every projection of ``finetuned_compressed_model`` by ``BinaryDiff(base weight, fine-tuned weight)`` built on
    the device the projection lives on (reference diff.py:41-64).  The dense layer is dropped before the compressed one is
    attached so that a 7B model never holds both."""
    for parent, path, child in _projection_sites(finetuned_compressed_model):
        where = getattr(parent, child).weight.device
        leaf = f"{path}.{child}"
        w_base = base_model.get_submodule(leaf).weight.detach().to(where)
        w_fine = finetuned_model.get_submodule(leaf).weight.detach().to(where)
        replacement = BinaryDiff(base=w_base, finetune=w_fine).to(where)
        del w_base, w_fine
        setattr(parent, child, None)  # frees the dense projection
        _release_cached_memory()
        setattr(parent, child, replacement)


def save_diff(finetuned_compressed_model, save_dir):
    """Write ``diff.pt`` in the reference's layout (diff.py:66-79): for every BinaryDiff ``<path>.mask`` (int32 [K/32,N]) and
    ``<path>.coeff`` (fp32 0-dim), then every trainable parameter under its own name (which stores the coeffs again under
    the same keys, and the full embeddings / norms / lm_head).  Everything is moved to the CPU; key order is the reference's."""
    model = finetuned_compressed_model
    entries = {}
    for path, mod in model.named_modules():
        if not isinstance(mod, BinaryDiff):
            continue
        entries[f"{path}.mask"] = mod.mask.cpu()
        entries[f"{path}.coeff"] = mod.coeff.cpu()
    entries.update((pname, p.cpu()) for pname, p in model.named_parameters() if p.requires_grad)
    torch.save(entries, save_dir)


@torch.no_grad()
def load_diff(model, diff_dir):
    """Fold a ``diff.pt`` into a dense model (reference diff.py:81-106).  Per module path: a ``.mask`` entry adds
    ``((2*unpack(mask)-1)*coeff).T`` to the weight (one device kernel, no [K,N] temporaries); a ``.weight`` entry replaces
    the parameter (cast to the model's dtype); an ``.A``/``.B`` LoRA pair adds ``(A@B).T``.  The config's vocabulary size
    follows the (possibly replaced) lm_head."""
    where = model.device
    # diff.pt is a flat dict of tensors and Parameters (often downloaded from third parties): the safe unpickler is enough
    stored = torch.load(diff_dir, map_location="cpu", weights_only=True)

    def fetch(key):
        return stored[key].to(where)

    for path, mod in model.named_modules():
        if f"{path}.mask" in stored:
            fold_into(mod.weight, fetch(f"{path}.mask"), fetch(f"{path}.coeff"))
        elif f"{path}.weight" in stored:
            mod.weight = nn.Parameter(fetch(f"{path}.weight").to(mod.weight.dtype))
        elif f"{path}.A" in stored:
            low_rank = fetch(f"{path}.A") @ fetch(f"{path}.B")
            mod.weight.add_(low_rank.T.to(mod.weight.dtype))
    model.config.vocab_size = model.lm_head.weight.shape[0]


@torch.no_grad()
def fold_into(weight: torch.Tensor, mask: torch.Tensor, coeff: torch.Tensor) -> None:
    """In place ``weight[N,K] += round(((2*unpack(mask)-1) * coeff).T)`` in the weight dtype (diff.py:93-95)."""
    N, K = weight.shape
    assert mask.shape == (K // 32, N) and mask.dtype == torch.int32
    if weight.is_cuda and weight.dtype in (torch.bfloat16, torch.float16) and weight.is_contiguous():
        c = coeff.detach().to(device=weight.device, dtype=torch.float32).reshape(1).contiguous()
        m = mask.to(weight.device).contiguous()
        with torch.cuda.device(weight.device):
            _lib.check(_lib.lib.bd_fold(weight.data_ptr(), m.data_ptr(), c.data_ptr(), _lib.dtype_code(weight.dtype), N, K,
                                        _lib.stream_ptr(weight.device)))
    else:
        delta = (unpack(mask) * 2 - 1) * coeff
        weight.add_(delta.T.to(weight.dtype))


def save_full_model(base_model_name, finetuned_model_name, diff_dir, save_dir, device):
    """Materialise the fine-tuned model from base + ``diff.pt`` and save it with the fine-tune's tokenizer (diff.py:108-116)."""
    import transformers

    dense = transformers.AutoModelForCausalLM.from_pretrained(base_model_name, torch_dtype=torch.bfloat16, low_cpu_mem_usage=True)
    dense = dense.to(device)
    load_diff(dense, diff_dir)
    dense.save_pretrained(save_dir)
    transformers.AutoTokenizer.from_pretrained(finetuned_model_name).save_pretrained(save_dir)
    del dense
