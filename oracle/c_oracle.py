"""ctypes loader for oracle/bitdelta_oracle.c (TEST / CPU-BASELINE INFRASTRUCTURE ONLY; see that file's header)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "bitdelta_oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "_build/liboracle.so"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        vp, i64 = ctypes.c_void_p, ctypes.c_int64
        _lib.oracle_fwd_batched_bf16.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i64, ctypes.c_int]
        _lib.oracle_fwd_batched_bf16.restype = None
        _lib.oracle_unpack_i32.argtypes = [vp, vp, i64, i64]
        _lib.oracle_pack_i32.argtypes = [vp, vp, i64, i64]
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def _cgroup_cpu_limit():
    """CPUs this container may use according to its cgroup CPU quota (v2 cpu.max or v1 cfs quota), or None."""
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            return max(1, -(-int(quota) // int(period)))
    except (OSError, ValueError):
        pass
    try:
        quota = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if quota > 0 and period > 0:
            return max(1, -(-quota // period))
    except (OSError, ValueError):
        pass
    return None


def max_threads() -> int:
    """Host threads the baseline can really use: the OpenMP processor count, capped by the scheduler affinity and by the
    cgroup CPU quota (on the B200 boxes nproc says 128 but the quota is 16 CPUs: 128 threads get throttled to 0.4 s per
    layer where 16 threads take 0.03 s)."""
    import os

    n = int(lib().oracle_max_threads())
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    lim = _cgroup_cpu_limit()
    if lim is not None:
        n = min(n, lim)
    return max(1, n)


def fwd_batched_bf16(x_bits, w_bits, masks, coeff, threads: int = 0) -> np.ndarray:
    """x_bits [T,m,K] uint16, w_bits [N,K] uint16 or None, masks [T,K/32,N] int32, coeff [T] float32 or None -> y [T,m,N] float32."""
    x_bits = np.ascontiguousarray(x_bits, np.uint16)
    masks = np.ascontiguousarray(masks, np.int32)
    T, m, K = x_bits.shape
    N = masks.shape[-1]
    assert masks.shape == (T, K // 32, N)
    y = np.empty((T, m, N), np.float32)
    wp = None
    if w_bits is not None:
        w_bits = np.ascontiguousarray(w_bits, np.uint16)
        assert w_bits.shape == (N, K)
        wp = w_bits.ctypes.data
    cp = None
    if coeff is not None:
        coeff = np.ascontiguousarray(coeff, np.float32)
        cp = coeff.ctypes.data
    lib().oracle_fwd_batched_bf16(x_bits.ctypes.data, wp, masks.ctypes.data, cp, y.ctypes.data, T, m, K, N, threads)
    return y


def unpack_i32(words: np.ndarray) -> np.ndarray:
    words = np.ascontiguousarray(words, np.int32)
    J, N = words.shape
    bits = np.empty((J * 32, N), np.uint8)
    lib().oracle_unpack_i32(words.ctypes.data, bits.ctypes.data, J, N)
    return bits.astype(bool)


def pack_i32(bits: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(bits, np.uint8)
    K, N = b.shape
    words = np.empty((K // 32, N), np.int32)
    lib().oracle_pack_i32(b.ctypes.data, words.ctypes.data, K, N)
    return words
