"""Drop-in for the reference module ``bitdelta/binary_gemm_kernel.py``: same names, signatures and assertions.

    pack(x, n_bits=32)                         reference :6-32
    unpack(x, n_bits=32)                       reference :34-46
    binary_matmul(a, b, n_bits=32, activation="")   reference :153-184
    binary_bmm(a, b, n_bits=32, activation="")      reference :297-335

The Triton kernels are replaced by libbitdelta_b200.so (hand-written sm_100a CUDA behind a C ABI, see
include/bitdelta_b200.h).  CUDA tensors run on the device on the current stream of ``a.device``; CPU tensors are
accepted by pack/unpack only (the reference's diff.pt path packs and unpacks on the CPU) and go through the library's
host codec.  There is no PyTorch fallback for the GEMMs: they need a CUDA device.
"""
from __future__ import annotations

import torch

from . import _lib

_WORD_DTYPES = {8: torch.uint8, 16: torch.int16, 32: torch.int32, 64: torch.int64}


def _as_bool_bytes(x: torch.Tensor) -> torch.Tensor:
    if x.dtype != torch.bool:
        x = x != 0
    return x.contiguous()


def pack(x: torch.Tensor, n_bits: int = 32) -> torch.Tensor:
    """pack n_bits of x into a single integer.

    x: bool tensor (*, K, N)  ->  int tensor (*, K // n_bits, N); bit i (LSB first) of word [j, n] is x[n_bits*j+i, n].
    """
    assert x.shape[-2] % n_bits == 0, "K must be divisible by n_bits"
    if n_bits not in _WORD_DTYPES:
        raise ValueError(f"n_bits must be one of {sorted(_WORD_DTYPES)}, got {n_bits}")
    lead = tuple(x.shape[:-2])
    K, N = x.shape[-2], x.shape[-1]
    batch = 1
    for d in lead:
        batch *= d
    bits = _as_bool_bytes(x)
    out = torch.empty(lead + (K // n_bits, N), dtype=_WORD_DTYPES[n_bits], device=x.device)
    if out.numel() == 0:
        return out
    if x.is_cuda:
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib.bd_pack(bits.data_ptr(), out.data_ptr(), n_bits, batch, K, N, _lib.stream_ptr(x.device)))
    else:
        _lib.check(_lib.lib.bd_pack_host(bits.data_ptr(), out.data_ptr(), n_bits, batch, K, N))
    return out


def unpack(x: torch.Tensor, n_bits: int = 32) -> torch.Tensor:
    """unpack each integer of x into n_bits booleans.

    x: int tensor (*, K // n_bits, N)  ->  bool tensor (*, K, N)
    """
    if n_bits not in _WORD_DTYPES:
        raise ValueError(f"n_bits must be one of {sorted(_WORD_DTYPES)}, got {n_bits}")
    if x.dtype != _WORD_DTYPES[n_bits]:
        # the reference shifts whatever integer dtype it is given; only the low n_bits matter
        x = x.to(_WORD_DTYPES[n_bits])
    x = x.contiguous()
    lead = tuple(x.shape[:-2])
    J, N = x.shape[-2], x.shape[-1]
    batch = 1
    for d in lead:
        batch *= d
    out = torch.empty(lead + (J * n_bits, N), dtype=torch.bool, device=x.device)
    if out.numel() == 0:
        return out
    if x.is_cuda:
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib.bd_unpack(x.data_ptr(), out.data_ptr(), n_bits, batch, J, N, _lib.stream_ptr(x.device)))
    else:
        _lib.check(_lib.lib.bd_unpack_host(x.data_ptr(), out.data_ptr(), n_bits, batch, J, N))
    return out


def _require_cuda(t: torch.Tensor, who: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{who}: the W1A16 GEMM is implemented only as sm_100a CUDA kernels (libbitdelta_b200.so); "
            f"got a tensor on {t.device}. There is no CPU fallback."
        )


def _bmm_impl(a3: torch.Tensor, b_words: torch.Tensor, b_batch_stride: int, kernel="auto") -> torch.Tensor:
    B, M, K = a3.shape
    N = b_words.shape[-1]
    c = torch.empty((B, M, N), device=a3.device, dtype=a3.dtype)
    if c.numel() == 0:
        return c
    with torch.cuda.device(a3.device):
        ws = _lib.workspace(a3.device, B * M, N)
        _lib.check(
            _lib.lib.bd_binary_bmm(
                a3.data_ptr(), b_words.data_ptr(), c.data_ptr(), _lib.dtype_code(a3.dtype), B, M, K, N, b_batch_stride,
                ws.data_ptr(), ws.numel(), _lib.kernel_code(kernel), _lib.stream_ptr(a3.device),
            )
        )
    return c


def binary_matmul(a: torch.Tensor, b: torch.Tensor, n_bits: int = 32, activation: str = "", *, kernel="auto") -> torch.Tensor:
    """
    a: float tensor (M, K)
    b: int tensor (K // n_bits, N), packed booleans
    returns (M, N) = a @ (2*unpack(b) - 1), dtype of a.  `activation` is accepted and ignored, as in the reference.
    """
    assert a.shape[1] == b.shape[0] * n_bits, "Incompatible dimensions"
    assert a.is_contiguous(), "Matrix A must be contiguous"
    assert b.is_contiguous(), "Matrix B must be contiguous"
    if n_bits != 32 or b.dtype != torch.int32:
        raise NotImplementedError("binary_matmul: only int32 words (n_bits=32) are implemented, as used by the reference")
    _require_cuda(a, "binary_matmul")
    assert a.device == b.device, "A and B must be on the same device"
    return _bmm_impl(a.unsqueeze(0), b, 0, kernel)[0]


def binary_bmm(a: torch.Tensor, b: torch.Tensor, n_bits: int = 32, activation: str = "", *, kernel="auto") -> torch.Tensor:
    """
    a: float tensor (B, M, K)
    b: int tensor (B, K // n_bits, N), packed booleans
    returns (B, M, N) with c[i] = a[i] @ (2*unpack(b[i]) - 1), dtype of a.
    """
    assert a.dim() == 3, "Matrix A must be 3D"
    assert b.dim() == 3, "Matrix B must be 3D"
    assert a.shape[2] == b.shape[1] * n_bits, "Incompatible dimensions"
    assert a.shape[0] == b.shape[0], "Incompatible batch dimensions"
    assert a.is_contiguous(), "Matrix A must be contiguous"
    assert b.is_contiguous(), "Matrix B must be contiguous"
    assert a.device == b.device, "A and B must be on the same device"
    if n_bits != 32 or b.dtype != torch.int32:
        raise NotImplementedError("binary_bmm: only int32 words (n_bits=32) are implemented, as used by the reference")
    _require_cuda(a, "binary_bmm")
    return _bmm_impl(a, b, b.shape[1] * b.shape[2], kernel)
