"""ctypes binding of libbitdelta_b200.so (the C ABI declared in include/bitdelta_b200.h).

The library is the product: there is no Python/PyTorch fallback.  If the shared object is missing the import fails
loudly with the command that builds it.  PyTorch is used only for device memory, streams and tensors' raw pointers.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libbitdelta_b200.so")
# tools/ (kernel A/B runs, the clock64 trace) opt into the separately built bring-up library; the release library has
# neither the knobs nor the entry points, so nothing in the environment can change what a forward computes.
BRINGUP = os.environ.get("BD_BRINGUP_LIB") == "1"
if BRINGUP:
    LIB_PATH = os.path.join(_PKG, "libbitdelta_b200_bringup.so")

BD_BF16, BD_FP16, BD_FP32 = 0, 1, 2
KERNEL_AUTO, KERNEL_SIMT, KERNEL_UMMA = 0, 1, 2
_KERNELS = {"auto": KERNEL_AUTO, "simt": KERNEL_SIMT, "umma": KERNEL_UMMA}
_DTYPES = {torch.bfloat16: BD_BF16, torch.float16: BD_FP16, torch.float32: BD_FP32}

EXPORTS = [
    "bd_abi_version", "bd_last_error", "bd_launch_count",
    "bd_pack", "bd_unpack", "bd_pack_host", "bd_unpack_host",
    "bd_compress", "bd_fold",
    "bd_binary_bmm", "bd_binarydiff_fwd_batched", "bd_binarydiff_fwd_grouped",
    "bd_tenant_linear", "bd_tenant_rmsnorm", "bd_tenant_embed",
    "bd_workspace_bytes", "bd_select_kernel",
    "bd_tp_buffer_bytes", "bd_tp_buffer_create", "bd_tp_buffer_open", "bd_tp_buffer_close", "bd_tp_allreduce",
]


class BitDeltaLibraryError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library is the only implementation of this package. "
            "Build it with `python bitdelta_b200/build.py` (needs nvcc, no GPU required)."
        )
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, i64, i32, sz = c.c_void_p, c.c_int64, c.c_int, c.c_size_t
    lib.bd_abi_version.restype = i32
    lib.bd_last_error.restype = c.c_char_p
    lib.bd_launch_count.restype = c.c_uint64
    lib.bd_pack.argtypes = [vp, vp, i32, i64, i64, i64, vp]
    lib.bd_unpack.argtypes = [vp, vp, i32, i64, i64, i64, vp]
    lib.bd_pack_host.argtypes = [vp, vp, i32, i64, i64, i64]
    lib.bd_unpack_host.argtypes = [vp, vp, i32, i64, i64, i64]
    lib.bd_compress.argtypes = [vp, vp, i32, i64, i64, vp, vp, vp, vp]
    lib.bd_fold.argtypes = [vp, vp, vp, i32, i64, i64, vp]
    lib.bd_binary_bmm.argtypes = [vp, vp, vp, i32, i64, i64, i64, i64, i64, vp, sz, i32, vp]
    lib.bd_binarydiff_fwd_batched.argtypes = [vp, vp, vp, vp, i32, vp, i32, i64, i64, i64, i64, i64, vp, sz, i32, vp]
    lib.bd_binarydiff_fwd_grouped.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, i64, i64, i64, vp, sz, i32, vp]
    lib.bd_tenant_linear.argtypes = [vp, vp, vp, vp, vp, i32, i64, i64, i64, i64, vp]
    lib.bd_tenant_rmsnorm.argtypes = [vp, vp, vp, i32, i64, i64, i64, c.c_float, vp]
    lib.bd_tenant_embed.argtypes = [vp, vp, vp, vp, i32, i64, i64, i64, vp]
    lib.bd_workspace_bytes.argtypes = [i64, i64]
    lib.bd_workspace_bytes.restype = sz
    lib.bd_select_kernel.argtypes = [i32, i64, i64, i64, i64, i32]
    lib.bd_tp_buffer_bytes.argtypes = [i64, i32]
    lib.bd_tp_buffer_bytes.restype = sz
    lib.bd_tp_buffer_create.argtypes = [sz, c.POINTER(vp), vp]
    lib.bd_tp_buffer_open.argtypes = [vp, c.POINTER(vp)]
    lib.bd_tp_buffer_close.argtypes = [vp, i32]
    lib.bd_tp_allreduce.argtypes = [c.POINTER(vp), sz, i32, i32, vp, i64, vp, i32, vp]
    if BRINGUP:
        lib.bd_debug_set_trace.argtypes = [vp]
        lib.bd_debug_set_trace.restype = None
        lib.bd_debug_set_flags.argtypes = [i32, i32]
        lib.bd_debug_set_flags.restype = None
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is c.c_int and name not in ("bd_abi_version", "bd_select_kernel"):
            fn.restype = i32
    return lib


lib = _load()


def check(status: int) -> None:
    """Maps a non-zero C status to RuntimeError carrying bd_last_error() (SURVEY.md section 8b error convention)."""
    if status != 0:
        msg = lib.bd_last_error().decode("utf-8", "replace")
        raise BitDeltaLibraryError(f"bitdelta_b200 C ABI error {status}: {msg}")


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError:
        raise TypeError(f"bitdelta_b200 supports bfloat16/float16 activations, got {dt}") from None


FLAG_STATIC_OPERANDS, FLAG_FP32_OUT = 1 << 8, 1 << 9  # bd_launch_flag


def kernel_code(name, static_operands: bool = False, fp32_out: bool = False) -> int:
    """`kernel` argument of the forward entry points: bd_kernel selector | bd_launch_flag bits."""
    code = name if isinstance(name, int) else _KERNELS[name]
    return code | (FLAG_STATIC_OPERANDS if static_operands else 0) | (FLAG_FP32_OUT if fp32_out else 0)


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def launch_count() -> int:
    return int(lib.bd_launch_count())


# ---- per-(device, stream) zero-initialised workspaces; the kernels leave them zeroed ----
_ws_lock = threading.Lock()
_workspaces: dict = {}
_WS_MAX_ROWS = 1 << 20  # bd_workspace_bytes clamps the row count to the kernels' per-launch maximum


def workspace(device: torch.device, rows: int, n: int) -> torch.Tensor:
    """The workspace of (device, current stream).  bd_workspace_bytes is bounded (row chunks are capped at 128 rows, N does
    not enter), so the maximum is allocated once and the buffer is NEVER replaced: a CUDA graph captured earlier keeps a
    valid pointer whatever is launched on the stream afterwards (a later, larger eager call used to free the buffer a
    captured graph still wrote its split-K partials and tile counters to)."""
    del rows, n  # every problem size fits the one buffer
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream_ptr(device))
    with _ws_lock:
        ws = _workspaces.get(key)
        if ws is None:
            if torch.cuda.is_current_stream_capturing():
                raise BitDeltaLibraryError(
                    "the first call on a stream allocates its workspace; run one warm-up call before capturing a CUDA graph"
                )
            ws = torch.zeros(int(lib.bd_workspace_bytes(_WS_MAX_ROWS, 1)), dtype=torch.uint8, device=device)
            _workspaces[key] = ws
    return ws
