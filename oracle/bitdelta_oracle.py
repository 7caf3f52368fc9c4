"""CPU oracle for the BitDelta W1A16 hot path -- TEST INFRASTRUCTURE ONLY.

This module restates, in numpy, the arithmetic of the reference's hot path so the CUDA
product can be checked against it.  Nothing under ``bitdelta_b200/`` may import it; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs do.

Parity status: PINNED.  ``tests/golden/*.npz`` were produced by importing the reference itself
(``/root/reference/bitdelta/binary_gemm_kernel.py`` ``pack``/``unpack``, ``bitdelta/diff.py``
``BinaryDiff``/``save_diff``/``load_diff``, the Triton kernel body executed under
``TRITON_INTERPRET=1``, and the class definitions of ``demo/demo_backend.py:62-98``) with the
committed script ``tests/golden/gen_golden.py``; ``tests/test_oracle_golden.py`` checks every
function below against those vectors.

Every function cites the reference lines it follows (paths relative to the reference root).
bf16 values travel as float32 arrays whose low 16 mantissa bits are zero (numpy has no bf16).
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# dtype helpers
# --------------------------------------------------------------------------------------


def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """Round float32 -> bfloat16 (round-to-nearest-even), returned as float32.

    This is what ``tensor.to(torch.bfloat16)`` does for finite values; NaN is preserved.
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32)
    bias = ((u >> 16) & 1) + np.uint32(0x7FFF)
    r = ((u + bias) & np.uint32(0xFFFF0000)).astype(np.uint32)
    out = r.view(np.float32).copy()
    nan = np.isnan(x)
    if nan.any():
        out[nan] = np.nan
    return out


def bf16_bits_to_f32(bits: np.ndarray) -> np.ndarray:
    """uint16 raw bf16 patterns -> float32 values."""
    return (bits.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> uint16 raw bf16 patterns (RNE)."""
    return (round_to_bf16(x).view(np.uint32) >> 16).astype(np.uint16)


def round_to_fp16(x: np.ndarray) -> np.ndarray:
    """float32 -> float16 -> float32 (RNE, overflow to inf), the ``accumulator.to(tl.float16)`` step."""
    with np.errstate(over="ignore"):
        return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float32)


_INT_DTYPES = {8: np.uint8, 16: np.int16, 32: np.int32, 64: np.int64}
_UINT_DTYPES = {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}

# --------------------------------------------------------------------------------------
# a1 / a2: the bit codec
# --------------------------------------------------------------------------------------


def pack(bits: np.ndarray, n_bits: int = 32) -> np.ndarray:
    """bool ``(*, K, N)`` -> int ``(*, K//n_bits, N)``; bit ``i`` (LSB = 0) of word ``[j, n]`` is ``bits[n_bits*j+i, n]``.

    Follows ``bitdelta/binary_gemm_kernel.py:6-32``: the reference shifts each bool by its
    position inside the ``n_bits`` group, sums the group in int64 and casts to the word dtype
    (so for ``n_bits=32`` bit 31 lands in the int32 sign bit).  Same assertion text as :13.
    """
    bits = np.asarray(bits)
    assert bits.shape[-2] % n_bits == 0, "K must be divisible by n_bits"
    lead = bits.shape[:-2]
    K, N = bits.shape[-2:]
    g = bits.reshape(-1, K // n_bits, n_bits, N).astype(np.uint64)
    weights = (np.uint64(1) << np.arange(n_bits, dtype=np.uint64))[None, None, :, None]
    words = (g * weights).sum(axis=2, dtype=np.uint64)  # disjoint bits: sum == OR
    words = words.astype(_UINT_DTYPES[n_bits]).view(_INT_DTYPES[n_bits])
    return words.reshape(*lead, K // n_bits, N)


def unpack(words: np.ndarray, n_bits: int = 32) -> np.ndarray:
    """int ``(*, K//n_bits, N)`` -> bool ``(*, K, N)``; inverse of :func:`pack`.

    Follows ``bitdelta/binary_gemm_kernel.py:34-46`` (``(x >> shift) & 1`` with an arithmetic
    shift on the signed word; the ``& 1`` makes sign extension irrelevant).
    """
    words = np.asarray(words)
    lead = words.shape[:-2]
    J, N = words.shape[-2:]
    u = words.reshape(-1, J, 1, N).astype(np.int64)
    shifts = np.arange(n_bits, dtype=np.int64)[None, None, :, None]
    b = (u >> shifts) & 1
    return b.reshape(*lead, J * n_bits, N).astype(bool)


def signs_pm1(words: np.ndarray, dtype=np.float32) -> np.ndarray:
    """``unpack(words)*2-1`` as a float matrix ``(*, K, N)``: bit 1 -> +1, bit 0 -> -1.

    ``binary_gemm_kernel.py:128-129`` / ``:271-272`` (kernel) and ``diff.py:93`` (fold path).
    """
    return unpack(words).astype(dtype) * dtype(2) - dtype(1)


# --------------------------------------------------------------------------------------
# a3 / a4: the W1A16 GEMMs
# --------------------------------------------------------------------------------------


def binary_matmul_exact(a: np.ndarray, b_words: np.ndarray) -> np.ndarray:
    """float64 truth of ``C[M,N] = A[M,K] . (2*unpack(B)-1)`` (no output rounding)."""
    assert a.shape[1] == b_words.shape[0] * 32, "Incompatible dimensions"
    return np.asarray(a, dtype=np.float64) @ signs_pm1(b_words, np.float64)


def binary_matmul(a: np.ndarray, b_words: np.ndarray, out: str = "bf16") -> np.ndarray:
    """The reference kernel's rounding chain: fp32 accumulate, ``.to(tl.float16)``, store as ``a.dtype``.

    ``binary_gemm_kernel.py:118`` (fp32 accumulator), ``:143`` (cast to fp16), ``:167`` (output tensor
    has ``a.dtype``).  ``out`` names ``a.dtype``: "fp16" -> result is the fp16 value; "bf16" -> the fp16
    value rounded again to bf16 (the double rounding SURVEY.md section 2b documents).
    """
    acc = (np.asarray(a, dtype=np.float32) @ signs_pm1(b_words, np.float32)).astype(np.float32)
    c = round_to_fp16(acc)
    if out == "bf16":
        c = round_to_bf16(c)
    elif out != "fp16":
        raise ValueError(out)
    return c


def binary_bmm_exact(a: np.ndarray, b_words: np.ndarray) -> np.ndarray:
    """float64 truth of the batched product, ``C[b] = A[b] . (2*unpack(B[b])-1)`` (``binary_gemm_kernel.py:297-335``)."""
    assert a.ndim == 3, "Matrix A must be 3D"
    assert b_words.ndim == 3, "Matrix B must be 3D"
    assert a.shape[2] == b_words.shape[1] * 32, "Incompatible dimensions"
    assert a.shape[0] == b_words.shape[0], "Incompatible batch dimensions"
    return np.stack([binary_matmul_exact(a[i], b_words[i]) for i in range(a.shape[0])])


def binary_bmm(a: np.ndarray, b_words: np.ndarray, out: str = "bf16") -> np.ndarray:
    """Batched kernel with the reference's rounding chain (``binary_gemm_kernel.py:260,287,314``)."""
    assert a.ndim == 3, "Matrix A must be 3D"
    assert b_words.ndim == 3, "Matrix B must be 3D"
    assert a.shape[2] == b_words.shape[1] * 32, "Incompatible dimensions"
    assert a.shape[0] == b_words.shape[0], "Incompatible batch dimensions"
    return np.stack([binary_matmul(a[i], b_words[i], out) for i in range(a.shape[0])])


# --------------------------------------------------------------------------------------
# a5 / a6: BinaryDiff
# --------------------------------------------------------------------------------------


def binarydiff_compress(base: np.ndarray, finetune: np.ndarray):
    """``BinaryDiff.__init__`` (``diff.py:9-31``): returns ``(mask int32 [K/32,N], coeff float32 scalar)``.

    ``base``/``finetune`` are ``[N, K]`` bf16-valued float32 arrays.  ``diff = finetune - base`` is
    computed in the weight dtype (one bf16 rounding, :11); ``coeff = mean(|diff|)`` in float32 (:12);
    the bit is 1 unless ``diff < 0`` -- so ``diff == 0`` and ``-0.0`` map to +1 (:14-15); the bool
    matrix is transposed to ``[K, N]`` before packing (:16).
    """
    diff = round_to_bf16(np.asarray(finetune, np.float32) - np.asarray(base, np.float32))
    # torch's float32 mean over a big tensor is a pairwise/vectorised sum; the value is only
    # reproducible to ~1 ulp*log(n), so callers compare coeff with a relative tolerance.
    coeff = np.float32(np.abs(diff).astype(np.float64).mean())
    bits = ~(diff < 0)
    return pack(bits.T), coeff


def binarydiff_forward_exact(x, base, mask, coeff) -> np.ndarray:
    """float64 truth of ``y = x.W_base^T + coeff * (x . (2*unpack(mask)-1))`` (``diff.py:33-39``).

    x: ``(..., K)``; base: ``[N, K]`` (the module keeps ``base.T`` as a view, :19); mask ``[K/32, N]``.
    """
    x64 = np.asarray(x, np.float64)
    return x64 @ np.asarray(base, np.float64).T + np.float64(coeff) * (x64 @ signs_pm1(mask, np.float64))


def binarydiff_forward(x, base, mask, coeff, out: str = "bf16") -> np.ndarray:
    """``BinaryDiff.forward`` with the reference's rounding order (``diff.py:38-39``).

    1. ``x @ self.base`` -> fp32 accumulate, rounded to the activation dtype;
    2. ``binary_bmm`` -> fp32 accumulate, fp16, activation dtype (:func:`binary_bmm`);
    3. ``self.coeff * (...)`` -> fp32 0-dim parameter times 16-bit tensor, computed in fp32, rounded;
    4. the add, rounded.
    """
    rnd = round_to_bf16 if out == "bf16" else round_to_fp16
    x32 = np.asarray(x, np.float32)
    lead = x32.shape[:-1]
    x2 = x32.reshape(-1, x32.shape[-1])
    t_base = rnd(x2 @ np.asarray(base, np.float32).T)
    t_delta = binary_matmul(x2, mask, out)
    t_scaled = rnd(np.float32(coeff) * t_delta)
    y = rnd(t_base + t_scaled)
    return y.reshape(*lead, -1)


def fold_delta(base: np.ndarray, mask: np.ndarray, coeff, out: str = "bf16") -> np.ndarray:
    """``load_diff`` fold (``diff.py:93-95``): ``W += ((unpack(mask)*2-1) * coeff).T`` cast to the weight dtype, then added in it."""
    rnd = round_to_bf16 if out == "bf16" else round_to_fp16
    delta = rnd((signs_pm1(mask, np.float32) * np.float32(coeff)).T)
    return rnd(np.asarray(base, np.float32) + delta)


# --------------------------------------------------------------------------------------
# a7 / a8: the multi-tenant modules of the demo backend
# --------------------------------------------------------------------------------------


def diffcompress_forward_exact(x, weight, masks, coeffs) -> np.ndarray:
    """float64 truth of ``DiffCompressModule.forward`` (``demo/demo_backend.py:93-98``).

    x: ``[T, m, K]`` (row t of the batch belongs to tenant t, :101-102); weight: shared ``nn.Linear``
    weight ``[N, K]``; masks ``[T, K/32, N]``; coeffs ``[T]``.
    """
    x64 = np.asarray(x, np.float64)
    y = x64 @ np.asarray(weight, np.float64).T
    for t in range(x64.shape[0]):
        y[t] += np.float64(coeffs[t]) * (x64[t] @ signs_pm1(masks[t], np.float64))
    return y


def diffcompress_forward(x, weight, masks, coeffs, out: str = "bf16") -> np.ndarray:
    """``DiffCompressModule.forward`` with the reference's rounding order (``demo_backend.py:95-98``).

    ``self.module(x)`` rounds to the activation dtype; ``binary_bmm`` rounds through fp16; the product with
    ``coeff[:, None, None]`` (a tensor of the model dtype, ``demo_backend.py:37-39``) and the final add each round again.
    """
    rnd = round_to_bf16 if out == "bf16" else round_to_fp16
    x32 = np.asarray(x, np.float32)
    t_base = rnd(x32 @ np.asarray(weight, np.float32).T)
    t_delta = binary_bmm(x32, masks, out)
    c = rnd(np.asarray(coeffs, np.float32))[:, None, None]
    return rnd(t_base + rnd(t_delta * c))


def dataparallel_forward(x, weights, kind: str = "linear", eps: float = 1e-6, out: str = "bf16") -> np.ndarray:
    """``DataParallelModule.forward`` (``demo_backend.py:69-79``): tenant t's row goes through tenant t's weight.

    kind "linear": ``weights[t]`` is ``[V_t, K]``, ``x[t]`` is ``[m, K]`` -> ``[m, V_t]``; outputs of different
    width are right-padded with ``finfo(dtype).min`` to the widest (``torch.nested.to_padded_tensor``, :78-79).
    kind "embedding": ``x[t]`` holds integer ids.  kind "rmsnorm": Llama/Mistral RMSNorm with weight ``weights[t]``.
    """
    rnd = round_to_bf16 if out == "bf16" else round_to_fp16
    outs = []
    for t, w in enumerate(weights):
        w32 = np.asarray(w, np.float32)
        if kind == "linear":
            outs.append(rnd(np.asarray(x[t], np.float32) @ w32.T))
        elif kind == "embedding":
            outs.append(w32[np.asarray(x[t], np.int64)])
        elif kind == "rmsnorm":
            h = np.asarray(x[t], np.float32)
            var = (h * h).mean(-1, keepdims=True, dtype=np.float32)
            outs.append(rnd(w32 * rnd(h * (1.0 / np.sqrt(var + np.float32(eps))).astype(np.float32))))
        else:
            raise ValueError(kind)
    width = max(o.shape[-1] for o in outs)
    fill = np.float32(-3.3895313892515355e38) if out == "bf16" else np.float32(-65504.0)
    padded = np.full((len(outs),) + outs[0].shape[:-1] + (width,), fill, np.float32)
    for t, o in enumerate(outs):
        padded[t, ..., : o.shape[-1]] = o
    return padded


# --------------------------------------------------------------------------------------
# Error metrics used by the reference's own checks
# --------------------------------------------------------------------------------------


def rel_mean_abs_err(got: np.ndarray, ref: np.ndarray) -> float:
    """``(got-ref).abs().mean() / ref.abs().mean()`` -- the notebook's benchmark assertion (cells 22-24)."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.abs(got - ref).mean() / max(np.abs(ref).mean(), 1e-30))
