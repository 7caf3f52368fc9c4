"""The C-ABI library loads on a machine without a GPU and exports every symbol include/bitdelta_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bitdelta_b200.h")).read()
    return sorted(set(re.findall(r"^BD_API\s+[\w\s\*]+?\b(bd_\w+)\s*\(", src, flags=re.M)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ["bd_pack", "bd_unpack", "bd_binary_bmm", "bd_binarydiff_fwd_batched", "bd_compress", "bd_fold", "bd_last_error"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    from bitdelta_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == declared_symbols()
    assert _lib.lib.bd_abi_version() == 1


def test_errors_are_status_codes_with_messages():
    from bitdelta_b200 import _lib

    rc = _lib.lib.bd_pack_host(None, None, 32, 1, 33, 4)
    assert rc == -1 and b"K must be divisible by n_bits" in _lib.lib.bd_last_error()
    rc = _lib.lib.bd_pack_host(None, None, 7, 1, 32, 4)
    assert rc == -1 and b"n_bits" in _lib.lib.bd_last_error()
    # argument validation of the forward entry points happens before any CUDA call
    rc = _lib.lib.bd_binarydiff_fwd_batched(16, 16, 16, 16, 2, 16, 0, 1, 1, 33, 8, 0, None, 0, 0, None)
    assert rc == -1 and b"K must be divisible" in _lib.lib.bd_last_error()
    rc = _lib.lib.bd_binarydiff_fwd_batched(16, None, 16, 16, 2, 16, 0, 1, 1, 32, 8, 0, None, 0, 0, None)
    assert rc == -1 and b"w is required" in _lib.lib.bd_last_error()
    rc = _lib.lib.bd_binary_bmm(16, 16, 16, 5, 1, 1, 32, 8, 0, None, 0, 0, None)
    assert rc == -2
    # per-tenant leaves: same convention (validation before any CUDA call)
    import ctypes

    ptrs = (ctypes.c_void_p * 2)(256, 512)
    n_out = (ctypes.c_int64 * 2)(8, 8)
    rc = _lib.lib.bd_tenant_linear(256, ptrs, n_out, None, 256, 0, 2, 5, 64, 8, None)
    assert rc == -1 and b"rows per tenant" in _lib.lib.bd_last_error()
    rc = _lib.lib.bd_tenant_linear(256, ptrs, n_out, None, 256, 0, 2, 1, 60, 8, None)
    assert rc == -1 and b"multiple of 8" in _lib.lib.bd_last_error()
    n_bad = (ctypes.c_int64 * 2)(8, 9)
    rc = _lib.lib.bd_tenant_linear(256, ptrs, n_bad, None, 256, 0, 2, 1, 64, 8, None)
    assert rc == -1 and b"out of range" in _lib.lib.bd_last_error()
    rc = _lib.lib.bd_tenant_rmsnorm(256, ptrs, 256, 2, 2, 1, 64, 1e-5, None)
    assert rc == -1 and b"dtype" in _lib.lib.bd_last_error()
    rc = _lib.lib.bd_tenant_embed(None, ptrs, n_out, 256, 0, 2, 1, 64, None)
    assert rc == -1 and b"null pointer" in _lib.lib.bd_last_error()
    assert _lib.lib.bd_workspace_bytes(6, 14336) >= 8192
    assert _lib.lib.bd_select_kernel(0, 6, 1, 4096, 4096, 1) in (1, 2)
    assert _lib.lib.bd_select_kernel(0, 1, 1, 32, 50, 1) == 1  # odd shapes go to the general kernel


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "bitdelta_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt, f"{f} mentions the oracle"
