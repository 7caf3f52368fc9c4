"""Multi-GPU use of the BitDelta hot path (one process per GPU, torch.distributed for the plumbing).

The reference has no distributed code at all; BASELINE.json asks for two partitionings:

* **tenant sharding** (configs 3-4): tenants are independent units that share only the read-only base weights.  Every
  rank keeps a replica of W_base and serves a contiguous slice of the tenants -- no collective on the data path, only
  an optional host-side gather of the results (`gather_tenant_outputs`).
* **tensor parallelism** (config 5, Llama-2-70B): Megatron split of every BinaryDiff linear.  Column-parallel layers
  (q/k/v/gate/up) slice N of both the base weight and the sign words; row-parallel layers (o/down) slice K, which for
  the sign words means slicing whole 32-bit word rows (K_shard % 32 == 0).  The per-matrix scale alpha is shared by all
  shards.  A row-parallel forward ends with ONE sum all-reduce (NCCL over NVLink on the GPU box).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib
from .diff import _fused_forward


# --------------------------------------------------------------------------------------------- tenant sharding
def tenant_partition(num_tenants: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced split: returns (start, count) per rank; the first `num_tenants % world_size` ranks get one more."""
    base, rem = divmod(num_tenants, world_size)
    out, start = [], 0
    for r in range(world_size):
        cnt = base + (1 if r < rem else 0)
        out.append((start, cnt))
        start += cnt
    return out


def shard_checkpoints(checkpoint_list: Sequence[dict], rank: int, world_size: int) -> List[dict]:
    """The slice of `checkpoint_list` (one diff.pt dict per tenant, as `register_diff_compress` takes it) this rank serves."""
    start, cnt = tenant_partition(len(checkpoint_list), world_size)[rank]
    return list(checkpoint_list[start:start + cnt])


def gather_tenant_outputs(local: torch.Tensor, num_tenants: int, group=None) -> torch.Tensor:
    """Host-side convenience: all-gather per-tenant results `[T_local, ...]` into `[T, ...]` on every rank (ragged counts
    are padded to the largest shard for the collective and trimmed afterwards).  Not on the data path of a decode step."""
    world = dist.get_world_size(group)
    parts = tenant_partition(num_tenants, world)
    max_cnt = max(c for _, c in parts)
    pad = local.new_zeros((max_cnt,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, (_, c) in zip(bufs, parts)], dim=0)


# --------------------------------------------------------------------------------------------- tensor parallelism
def split_column_parallel(weight: torch.Tensor, masks: torch.Tensor, rank: int, world_size: int):
    """Slice the OUTPUT features: weight [N, K] -> [N/ws, K]; masks [..., K/32, N] -> [..., K/32, N/ws] (made contiguous)."""
    N = weight.shape[0]
    assert N % world_size == 0, "N must divide evenly across tensor-parallel ranks"
    n0, n1 = rank * N // world_size, (rank + 1) * N // world_size
    return weight[n0:n1].contiguous(), masks[..., n0:n1].contiguous()


def split_row_parallel(weight: torch.Tensor, masks: torch.Tensor, rank: int, world_size: int):
    """Slice the INPUT features: weight [N, K] -> [N, K/ws]; masks [..., K/32, N] -> [..., (K/32)/ws, N].

    K/ws must be a multiple of 32 so that a shard owns whole sign words (8192/8 and 28672/8 are)."""
    K = weight.shape[1]
    assert K % (32 * world_size) == 0, "K / world_size must be a multiple of 32 (whole sign words per shard)"
    k0, k1 = rank * K // world_size, (rank + 1) * K // world_size
    return weight[:, k0:k1].contiguous(), masks[..., k0 // 32:k1 // 32, :].contiguous()


class PeerExchange:
    """Exchange buffers of one tensor-parallel group for `bd_tp_allreduce` (include/bitdelta_b200.h, section e): each rank
    owns a buffer in device memory, exports it with CUDA IPC and maps the other ranks' buffers; the row-parallel linears
    of the group then sum their fp32 partials with one kernel over NVLink peer memory instead of an NCCL call.

    One process per GPU on one NVSwitch box (the ranks of `group` must be able to map each other's memory).  Collective
    to construct: every rank of the group must create it at the same point."""

    def __init__(self, max_elems: int, group=None, device=None):
        import ctypes

        assert dist.is_initialized(), "PeerExchange needs an initialised process group"
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.bytes = int(_lib.lib.bd_tp_buffer_bytes(int(max_elems), self.world))
        assert self.bytes > 0, "bad exchange size / world size"
        self.max_elems = int(max_elems)
        handle = ctypes.create_string_buffer(64)
        own = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.bd_tp_buffer_create(self.bytes, ctypes.byref(own), handle))
            handles = [None] * self.world
            dist.all_gather_object(handles, handle.raw, group=group)
            self._ptrs, self._opened = [], []
            for r, h in enumerate(handles):
                if r == self.rank:
                    self._ptrs.append(own.value)
                    continue
                p = ctypes.c_void_p()
                _lib.check(_lib.lib.bd_tp_buffer_open(ctypes.create_string_buffer(h, 64), ctypes.byref(p)))
                self._ptrs.append(p.value)
                self._opened.append(p.value)
        self._own = own.value
        self._arr = (ctypes.c_void_p * self.world)(*self._ptrs)
        dist.barrier(group=group)  # every rank has mapped every buffer before the first exchange

    def all_reduce(self, partial: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
        """Sum of the ranks' fp32 `partial` tensors, rounded once to `dtype`; asynchronous on the current stream."""
        assert partial.dtype == torch.float32 and partial.is_contiguous() and partial.numel() % 4 == 0
        assert partial.numel() <= self.max_elems, "exchange buffer too small for this tensor"
        y = torch.empty(partial.shape, device=partial.device, dtype=dtype)
        code = _lib.BD_FP32 if dtype == torch.float32 else _lib.dtype_code(dtype)
        with torch.cuda.device(partial.device):
            _lib.check(_lib.lib.bd_tp_allreduce(self._arr, self.bytes, self.rank, self.world, partial.data_ptr(), partial.numel(),
                                                y.data_ptr(), code, _lib.stream_ptr(partial.device)))
        return y

    def close(self):
        if getattr(self, "_own", None) is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # nobody is still reading a peer's buffer
        with torch.cuda.device(self.device):
            for p in self._opened:
                _lib.lib.bd_tp_buffer_close(p, 1)
            _lib.lib.bd_tp_buffer_close(self._own, 0)
        self._own, self._opened = None, []


class TensorParallelDiffLinear(nn.Module):
    """One tensor-parallel shard of a multi-tenant BinaryDiff linear (`DiffCompressModule` semantics, demo_backend.py:93-98).

    mode "column": output is this rank's slice of N (no communication).
    mode "row":    input is this rank's slice of K; the partial products are summed with one all-reduce.
    """

    def __init__(self, weight: torch.Tensor, masks: torch.Tensor, coeffs: torch.Tensor, mode: str, group=None, kernel="auto",
                 exchange: "PeerExchange | None" = None):
        super().__init__()
        assert mode in ("column", "row")
        self.mode, self.group, self.kernel = mode, group, kernel
        self.exchange = exchange  # row mode: sum over NVLink peer memory (bd_tp_allreduce) instead of an NCCL all-reduce
        self.register_buffer("weight", weight.contiguous())
        self.register_buffer("mask", masks.contiguous())   # [T, K_shard/32, N_shard]
        self.register_buffer("coeff", coeffs)

    @classmethod
    def from_full(cls, weight, masks, coeffs, mode: str, rank: int, world_size: int, group=None, exchange=None):
        w, m = (split_column_parallel if mode == "column" else split_row_parallel)(weight, masks, rank, world_size)
        return cls(w, m, coeffs, mode, group, exchange=exchange)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        # x: [T, m, K_shard]
        T = self.mask.shape[0]
        if self.mode == "row" and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            # The kernel hands back its UNROUNDED fp32 partial sums (BD_FLAG_FP32_OUT); they are summed across the ranks in
            # fp32 and rounded once, like the unsharded kernel rounds its fp32 accumulator once.
            y32 = _fused_forward(x.contiguous(), self.weight, self.mask, self.coeff, T, self.kernel, static_operands=True, out_fp32=True)
            if self.exchange is not None and y32.numel() % 4 == 0 and y32.numel() <= self.exchange.max_elems:
                return self.exchange.all_reduce(y32, x.dtype)
            dist.all_reduce(y32, op=dist.ReduceOp.SUM, group=self.group)
            return y32.to(x.dtype)
        return _fused_forward(x.contiguous(), self.weight, self.mask, self.coeff, T, self.kernel, static_operands=True)
