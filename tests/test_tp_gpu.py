"""Tensor-parallel BinaryDiff linears on real GPUs (BASELINE config 5): needs >= 2 GPUs on one box, one process per GPU.

Row-parallel shards of a 70B-shaped down_proj slice are summed (a) by bd_tp_allreduce over NVLink peer memory and (b) by an NCCL
all-reduce of the same fp32 partials; both must match the unsharded module within the 1e-3 bar, match each other bit for bit
on every rank, and keep doing so over many back-to-back exchanges and CUDA-graph replays (the exchange protocol is
double-buffered by a device-resident epoch)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip("needs at least two CUDA devices", allow_module_level=True)


def _worker(rank, world, port, results):
    import torch.distributed as dist

    import bitdelta_b200 as bd
    from bitdelta_b200.parallel import PeerExchange, TensorParallelDiffLinear

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        T, K, N = 8, 28672 // 4, 2048
        gen = torch.Generator(device=dev).manual_seed(5)  # same tensors on every rank, sharded locally
        w = (torch.randn(N, K, generator=gen, device=dev) * 0.02).bfloat16()
        masks = torch.randint(-(2**31), 2**31 - 1, (T, K // 32, N), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
        coeffs = (torch.rand(T, generator=gen, device=dev) * 0.003 + 0.0005).bfloat16()
        x = torch.randn(T, 1, K, generator=gen, device=dev).bfloat16()
        xs = x[..., rank * K // world:(rank + 1) * K // world].contiguous()
        lin = torch.nn.Linear(K, N, bias=False, device=dev, dtype=torch.bfloat16)
        lin.weight.data = w
        y_full = bd.DiffCompressModule(lin, masks, coeffs)(x)
        signs = bd.unpack(masks).double() * 2 - 1
        exact = x.double() @ w.double().T + coeffs.double()[:, None, None] * torch.bmm(x.double(), signs)
        ex = PeerExchange(T * N)
        shard = TensorParallelDiffLinear.from_full(w, masks, coeffs, "row", rank, world, exchange=ex)
        y_peer = shard(xs)
        shard.exchange = None
        y_nccl = shard(xs)
        shard.exchange = ex
        torch.cuda.synchronize(dev)
        rel = lambda a: ((a.double() - exact).abs().mean() / exact.abs().mean()).item()  # noqa: E731
        assert torch.equal(y_peer, y_nccl) or world > 2, "peer-memory sum and NCCL sum of two fp32 partials must agree bit for bit"
        assert rel(y_peer) < 2e-3 and rel(y_nccl) < 2e-3 and rel(y_full) < 2e-3
        assert ((y_peer.float() - y_full.float()).abs().mean() / y_full.float().abs().mean()).item() < 1e-3
        # every rank holds the same bits
        gathered = [torch.empty_like(y_peer) for _ in range(world)]
        dist.all_gather(gathered, y_peer)
        assert all(torch.equal(g, y_peer) for g in gathered)
        # protocol stress: many exchanges back to back with changing data, odd and even counts, then graph replays
        for it in range(41):
            xi = (xs.float() * (1.0 + 0.01 * it)).bfloat16()
            yi = shard(xi)
            shard.exchange = None
            yn = shard(xi)
            shard.exchange = ex
            assert torch.equal(yi, yn) or world > 2, f"iteration {it}"
        s = torch.cuda.Stream()
        torch.cuda.synchronize(dev)
        with torch.cuda.stream(s):
            shard(xs)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                a = shard(xs)
                b = shard(xs)
                c = shard(xs)  # odd number of exchanges per replay: the parity of the epoch, not of the call, picks the slot
            for _ in range(7):
                g.replay()
        torch.cuda.synchronize(dev)
        assert torch.equal(a, y_peer) and torch.equal(b, y_peer) and torch.equal(c, y_peer)
        ex.close()
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_row_parallel_peer_exchange_matches_nccl_and_unsharded(world):
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, 29533, results), nprocs=world, join=True)
    assert [results.get(r) for r in range(world)] == ["ok"] * world
