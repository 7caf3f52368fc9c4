"""Drop-in for the multi-tenant modules of the reference's ``demo/demo_backend.py`` (lines 62-179).

    DataParallelModule(module, weight_list)                reference :62-79
    DiffCompressModule(module, mask_list, coeff_list)      reference :82-98
    register_diff_compress / unregister_diff_compress      reference :107-166
    DiffCompress (context manager)                         reference :170-179

Convention kept from the reference (:101-102): batch size == number of checkpoints, and batch row ``i`` is served by
checkpoint ``i``.  ``DiffCompressModule.forward`` is one launch of the fused batched kernel: the shared base product
and all tenants' 1-bit deltas, with the per-tenant coefficient applied in the fp32 epilogue (the reference's
"TODO: Fuse coeff", :96).  The FastAPI server, prompt templating and tokenizers of the demo are out of scope.
"""
from __future__ import annotations

import ctypes
import gc

import torch
import torch.nn as nn

from . import _lib
from .diff import _fused_forward, _fused_forward_grouped


TENANT_LINEAR_MAX_ROWS = 4  # BD_TENANT_LINEAR_MAX_ROWS: decode-size launches; larger m is a plain library GEMM per tenant


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _native_leaf_ok(x, weights):
    w0 = weights[0]
    return (x.is_cuda and w0.dtype in (torch.bfloat16, torch.float16)
            and all(w.device == x.device and w.dtype == w0.dtype and w.is_contiguous() for w in weights))


def tenant_linear(x, weights, bias=None):
    """One launch of ``bd_tenant_linear``: ``y[t] = x[t] @ weights[t].T`` right-padded with ``finfo.min`` to the widest
    tenant (reference :72-79).  x: [T, m, K] with m <= 4."""
    T, m, K = x.shape
    sizes = [int(w.shape[0]) for w in weights]
    ldy = max(sizes)
    y = torch.empty((T, m, ldy), device=x.device, dtype=x.dtype)
    n_out = (ctypes.c_int64 * T)(*sizes)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib.bd_tenant_linear(x.data_ptr(), _ptr_array(weights), n_out, bias.data_ptr() if bias is not None else None,
                                             y.data_ptr(), _lib.dtype_code(x.dtype), T, m, K, ldy, _lib.stream_ptr(x.device)))
    return y


def tenant_rmsnorm(x, weights, eps):
    """One launch of ``bd_tenant_rmsnorm`` (HF Llama/Mistral RMSNorm arithmetic with tenant t's weight on row t)."""
    T, H = x.shape[0], x.shape[-1]
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib.bd_tenant_rmsnorm(x.data_ptr(), _ptr_array(weights), y.data_ptr(), _lib.dtype_code(x.dtype), T,
                                              x.numel() // (T * H), H, float(eps), _lib.stream_ptr(x.device)))
    return y


def tenant_embed(ids, weights):
    """One launch of ``bd_tenant_embed``: ``y[t, i] = weights[t][ids[t, i]]``."""
    T, H = ids.shape[0], int(weights[0].shape[1])
    y = torch.empty(tuple(ids.shape) + (H,), device=ids.device, dtype=weights[0].dtype)
    n_rows = (ctypes.c_int64 * T)(*[int(w.shape[0]) for w in weights])
    with torch.cuda.device(ids.device):
        _lib.check(_lib.lib.bd_tenant_embed(ids.data_ptr(), _ptr_array(weights), n_rows, y.data_ptr(), _lib.dtype_code(y.dtype), T,
                                            ids.numel() // T, H, _lib.stream_ptr(ids.device)))
    return y


class DataParallelModule(nn.Module):
    """Per-tenant full-precision leaf (embed_tokens, RMSNorm, lm_head): row i uses ``weight_list[i]`` (:62-79).

    Outputs of different width (ragged vocabularies) are right-padded with ``finfo(dtype).min`` exactly like the
    reference's nested-tensor padding (:78-79).  On a CUDA device the three leaf kinds the demo wraps run as ONE native
    launch over all tenants (``bd_tenant_linear`` / ``bd_tenant_rmsnorm`` / ``bd_tenant_embed``); any other module type
    keeps the reference's generic loop (swap ``weight.data``, call the module on row i).
    """

    def __init__(self, module, weight_list):
        super().__init__()
        self.module = module
        self.weight_list = weight_list
        self.original_weight = module.weight.data

    def _native(self, x):
        mod, ws = self.module, self.weight_list
        T = len(ws)
        if not x.is_cuda or x.shape[0] != T or x.dim() < 2:
            return None
        if isinstance(mod, nn.Embedding):
            if (x.dtype == torch.int64 and mod.max_norm is None and _native_leaf_ok(x, ws)
                    and len({int(w.shape[1]) for w in ws}) == 1):
                return tenant_embed(x.contiguous(), ws)
            return None
        if not (x.dtype == ws[0].dtype and _native_leaf_ok(x, ws)):
            return None
        if isinstance(mod, nn.Linear):
            K = x.shape[-1]
            m = x.numel() // (T * K)
            sizes = {int(w.shape[0]) for w in ws}
            bias = mod.bias
            if (1 <= m <= TENANT_LINEAR_MAX_ROWS and K % 8 == 0 and all(w.shape[1] == K and w.data_ptr() % 16 == 0 for w in ws)
                    and (bias is None or (len(sizes) == 1 and bias.dtype == x.dtype and bias.is_cuda))):
                y = tenant_linear(x.contiguous().view(T, m, K), ws, bias)
                return y.view(tuple(x.shape[:-1]) + (y.shape[-1],))
            return None
        if hasattr(mod, "variance_epsilon") and ws[0].dim() == 1 and type(mod).__name__.endswith(("RMSNorm", "LayerNorm")) \
                and not hasattr(mod, "bias") and all(w.shape[0] == x.shape[-1] for w in ws):
            return tenant_rmsnorm(x.contiguous(), ws, mod.variance_epsilon)
        return None

    def forward(self, hidden_states):
        y = self._native(hidden_states)
        if y is not None:
            return y
        outputs = []
        for i in range(len(self.weight_list)):
            self.module.weight.data = self.weight_list[i]
            outputs.append(self.module(hidden_states[i, None])[0])
        width = max(o.shape[-1] for o in outputs)
        if all(o.shape[-1] == width for o in outputs):
            return torch.stack(outputs, dim=0)
        out = outputs[0].new_full((len(outputs),) + tuple(outputs[0].shape[:-1]) + (width,), torch.finfo(outputs[0].dtype).min)
        for i, o in enumerate(outputs):
            out[i, ..., : o.shape[-1]] = o
        return out


class SiblingGroup:
    """DiffCompressModules that are always called back to back on the SAME input (q/k/v of an attention block, gate/up of an
    MLP -- exactly how the HF decoder layer calls them).  The first call of a round launches ONE grouped kernel for all
    members; the other members' calls return their cached output.  A member called with a different input (or twice)
    simply triggers a new launch, so results never depend on the caller following the pattern."""

    def __init__(self, members):
        self.members = list(members)
        self._key = None
        self._x = None  # keeps the activations alive while outputs are cached, so the key cannot be recycled
        self._outs = {}

    def invalidate(self, *_):
        """Drops cached outputs.  `fuse_sibling_projections` calls it from a forward pre-hook of the parent block, so every
        forward of the block starts a new round whatever happened to the activation buffer in between (a static-buffer
        decode loop under torch.inference_mode updates it in place, and inference tensors carry no version counter)."""
        self._key, self._x, self._outs = None, None, {}

    def forward_for(self, who, hidden_states):
        # The key is built from the tensor the CALLER passed (all members receive the same object from the decoder layer),
        # so a non-contiguous input is made contiguous once per round, not once per member.
        x0 = hidden_states
        ver = 0 if x0.is_inference() else x0._version  # inference tensors have no version counter: see invalidate()
        key = (x0.data_ptr(), ver, tuple(x0.shape), tuple(x0.stride()), x0.dtype, x0.device)
        if self._key != key or id(who) not in self._outs:
            T = who.mask.shape[0]
            x = x0.contiguous()
            static = all(m.module.weight.is_contiguous() for m in self.members)  # a fresh copy is still queued on the stream
            ws = [m.module.weight if m.module.weight.is_contiguous() else m.module.weight.contiguous() for m in self.members]
            ys = _fused_forward_grouped(x, ws, [m.mask for m in self.members], [m.coeff for m in self.members], T, who.kernel,
                                        static_operands=static)
            self._key, self._x = key, x0
            self._outs = {id(m): y for m, y in zip(self.members, ys)}
        y = self._outs.pop(id(who))
        if not self._outs:
            self._key, self._x = None, None
        return y


def group_projections(members):
    """Make `members` (DiffCompressModules with the same input features and tenants) share one grouped launch."""
    members = list(members)
    assert all(isinstance(m, DiffCompressModule) for m in members) and 2 <= len(members) <= 3
    K = members[0].module.weight.shape[1]
    assert all(m.module.weight.shape[1] == K and m.mask.shape[0] == members[0].mask.shape[0] for m in members)
    assert all(getattr(m.module, "bias", None) is None for m in members), "grouped launch supports bias-free linears"
    g = SiblingGroup(members)
    for m in members:
        m._group = g
    return g


def fuse_sibling_projections(model):
    """Opt-in: group q_proj/k_proj/v_proj and gate_proj/up_proj DiffCompressModules that share a parent (after
    register_diff_compress).  The HF decoder layer calls them one after the other on the same hidden states."""
    n = 0
    for _, parent in model.named_modules():
        kids = dict(parent.named_children())
        for names in (("q_proj", "k_proj", "v_proj"), ("gate_proj", "up_proj")):
            mods = [kids.get(nm) for nm in names]
            if all(isinstance(mm, DiffCompressModule) for mm in mods) and getattr(mods[0], "_group", None) is None:
                g = group_projections(mods)
                g._hook = parent.register_forward_pre_hook(g.invalidate)  # a new round with every forward of the block
                n += 1
    return n


class DiffCompressModule(nn.Module):
    """Shared ``nn.Linear`` + per-tenant 1-bit deltas (:82-98): ``y[t] = module(x[t]) + coeff[t] * (x[t] . sign_t)``."""

    kernel = "auto"
    _group = None  # set by group_projections / fuse_sibling_projections

    def __init__(self, module, mask_list, coeff_list):
        super().__init__()
        self.module = module
        self.mask = mask_list    # int32 [T, K/32, N], stacked once by register_diff_compress
        self.coeff = coeff_list  # [T], model dtype or fp32

    def forward(self, hidden_states):
        # hidden_states: (T, seq, K)
        T = self.mask.shape[0]
        assert hidden_states.dim() == 3 and hidden_states.shape[0] == T, "Incompatible batch dimensions"
        if self._group is not None and hidden_states.is_cuda:
            return self._group.forward_for(self, hidden_states)
        x = hidden_states.contiguous()
        w = self.module.weight
        static = w.is_contiguous()  # module-owned buffers; a fresh .contiguous() copy is still queued on the stream
        if not static:
            w = w.contiguous()
        y = _fused_forward(x, w, self.mask, self.coeff, T, self.kernel, static_operands=static)
        if getattr(self.module, "bias", None) is not None:
            y = y + self.module.bias
        return y


# Assume batch size = len(checkpoint_list); sample i uses checkpoint_list[i].
# Cache of stacked masks / coeffs, keyed by module path (reference :105).
cached_modules = {}


def _leaf_sites(model):
    """(path, leaf, parent module, attribute name) of every module without children, in traversal order."""
    sites = []
    for path, mod in model.named_modules():
        if path == "" or next(mod.named_children(), None) is not None:
            continue
        parent_path, _, attr = path.rpartition(".")
        sites.append((path, mod, model.get_submodule(parent_path), attr))
    return sites


def _stack_tenant_deltas(path, checkpoint_list):
    """Masks [T,K/32,N] and coefficients [T] of one projection over all tenants, stacked once; the per-tenant entries are
    removed from the checkpoint dicts so the memory is not held twice (reference :131-141)."""
    masks = torch.stack([ck[f"{path}.mask"] for ck in checkpoint_list], dim=0).contiguous()
    coeffs = torch.stack([ck[f"{path}.coeff"] for ck in checkpoint_list], dim=0)
    for ck in checkpoint_list:
        del ck[f"{path}.mask"], ck[f"{path}.coeff"]
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
    return masks, coeffs


def load_checkpoints_stacked(paths, device, dtype=torch.float16):
    """diff.pt files -> what ``register_diff_compress`` consumes, without the host-side detour of the reference's start-up
    (demo_backend.py:26-35 loads every file to the CPU, moves each tensor to the GPU, and :131-141 then stacks the tenants'
    masks on the device next to the per-tenant copies before popping them).

    The files are memory-mapped (``torch.load(mmap=True, weights_only=True)``); every projection's sign words go with ONE
    host-to-device copy per tenant straight into their slice of the final ``[T, K/32, N]`` tensor (the layout the fused
    kernel reads), the scales into the ``[T]`` tensor, and both are published in ``cached_modules`` under the module path
    exactly as ``register_diff_compress`` would have cached them.  Returns the checkpoint list (one dict per tenant) holding
    only what is left: the tenants' full-precision leaves (``*.weight``) on ``device`` in ``dtype``.  Nothing else changes:
    ``register_diff_compress(model, load_checkpoints_stacked(paths, device))``.
    """
    files = [torch.load(p, map_location="cpu", mmap=True, weights_only=True) for p in paths]
    T = len(files)
    out = [dict() for _ in range(T)]
    for key in files[0]:
        if key.endswith(".mask"):
            path = key[: -len(".mask")]
            first = files[0][key]
            masks = torch.empty((T,) + tuple(first.shape), dtype=torch.int32, device=device)
            for t, f in enumerate(files):
                masks[t].copy_(f[key], non_blocking=True)
            coeffs = torch.stack([f[f"{path}.coeff"].detach().reshape(()) for f in files]).to(device=device, dtype=dtype)
            cached_modules[path] = (masks, coeffs)
        elif key.endswith(".coeff"):
            continue
        else:
            for t, f in enumerate(files):
                out[t][key] = f[key].detach().to(device=device, dtype=dtype)
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)
    return out


def register_diff_compress(model, checkpoint_list):
    """Wrap every leaf whose path appears in the tenants' checkpoints (reference :107-153).

    ``<path>.weight`` present -> ``DataParallelModule`` over the tenants' full weights (embed_tokens, norms, lm_head);
    ``<path>.mask`` present (or stacked earlier and kept in ``cached_modules``) -> ``DiffCompressModule`` over the stacked
    masks / coefficients.  Anything else is left alone.
    """
    first = checkpoint_list[0]
    for path, leaf, parent, attr in _leaf_sites(model):
        if f"{path}.weight" in first:
            wrapped = DataParallelModule(leaf, [ck[f"{path}.weight"] for ck in checkpoint_list])
        elif f"{path}.mask" in first or path in cached_modules:
            assert isinstance(leaf, nn.Linear), "Only support linear layer"
            if path not in cached_modules:
                cached_modules[path] = _stack_tenant_deltas(path, checkpoint_list)
            wrapped = DiffCompressModule(leaf, *cached_modules[path])
        else:
            continue
        setattr(parent, attr, wrapped)


def unregister_diff_compress(model):
    """Put the original leaves back (reference :156-166); a DataParallelModule also restores the weight it swapped."""
    for path, mod in list(model.named_modules()):
        if isinstance(mod, DataParallelModule):
            mod.module.weight.data = mod.original_weight
        elif isinstance(mod, DiffCompressModule):
            hook = getattr(mod._group, "_hook", None) if mod._group is not None else None
            if hook is not None:  # the parent block's round-invalidation hook of fuse_sibling_projections
                hook.remove()
                mod._group._hook = None
            mod._group = None
        else:
            continue
        parent_path, _, attr = path.rpartition(".")
        setattr(model.get_submodule(parent_path), attr, mod.module)


class DiffCompress:
    """Context manager form (reference :170-179)."""

    def __init__(self, model, checkpoint_list):
        self.model = model
        self.checkpoint_list = checkpoint_list

    def __enter__(self):
        register_diff_compress(self.model, self.checkpoint_list)

    def __exit__(self, exc_type, exc_value, traceback):
        unregister_diff_compress(self.model)
