"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: tenant partitioning/gather and the tensor-parallel splits.
The per-shard arithmetic is done with the oracle (the CUDA kernels need a GPU); what is under test is the sharding."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bitdelta_b200 import parallel as P
        from oracle import bitdelta_oracle as O

        rng = np.random.default_rng(0)  # same data on every rank
        T, m, K, N = 5, 2, 128, 96
        x = O.round_to_bf16(rng.standard_normal((T, m, K)).astype(np.float32))
        w = O.round_to_bf16((rng.standard_normal((N, K)) * 0.05).astype(np.float32))
        masks = rng.integers(-(2**31), 2**31 - 1, (T, K // 32, N)).astype(np.int32)
        coeffs = (rng.random(T) * 0.01).astype(np.float32)
        full = O.diffcompress_forward_exact(x, w, masks, coeffs)

        # ---- tenant sharding: each rank evaluates its own tenants, results are gathered
        parts = P.tenant_partition(T, world)
        assert sum(c for _, c in parts) == T and parts[0][0] == 0 and all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        s, c = parts[rank]
        ckpts = [{"id": t} for t in range(T)]
        assert [d["id"] for d in P.shard_checkpoints(ckpts, rank, world)] == list(range(s, s + c))
        local = O.diffcompress_forward_exact(x[s:s + c], w, masks[s:s + c], coeffs[s:s + c])
        gathered = P.gather_tenant_outputs(torch.from_numpy(local), T)
        assert np.allclose(gathered.numpy(), full, rtol=1e-12, atol=1e-12)

        # ---- tensor parallel, column split: concatenating the shards' outputs along N gives the full output
        wt, mt = torch.from_numpy(w), torch.from_numpy(masks)
        wc, mc = P.split_column_parallel(wt, mt, rank, world)
        assert wc.shape == (N // world, K) and mc.shape == (T, K // 32, N // world) and mc.is_contiguous()
        yc = torch.from_numpy(O.diffcompress_forward_exact(x, wc.numpy(), mc.numpy(), coeffs))
        bufs = [torch.empty_like(yc) for _ in range(world)]
        dist.all_gather(bufs, yc)
        assert np.allclose(torch.cat(bufs, dim=-1).numpy(), full, rtol=1e-12, atol=1e-12)

        # ---- tensor parallel, row split: the shards' partial outputs sum (all-reduce) to the full output
        wr, mr = P.split_row_parallel(wt, mt, rank, world)
        k0, k1 = rank * K // world, (rank + 1) * K // world
        assert wr.shape == (N, K // world) and mr.shape == (T, K // 32 // world, N)
        yr = torch.from_numpy(O.diffcompress_forward_exact(x[:, :, k0:k1], wr.numpy(), mr.numpy(), coeffs))
        dist.all_reduce(yr)
        assert np.allclose(yr.numpy(), full, rtol=1e-10, atol=1e-10)
        with pytest.raises(AssertionError):
            P.split_row_parallel(wt[:, :96], mt[:, :3], rank, world)  # 96 / 2 = 48 is not a whole number of sign words
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback

        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_tenant_partition_properties():
    sys.path.insert(0, ROOT)
    from bitdelta_b200 import parallel as P

    for T in range(1, 20):
        for w in range(1, 9):
            parts = P.tenant_partition(T, w)
            assert len(parts) == w and sum(c for _, c in parts) == T
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
