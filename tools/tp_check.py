#!/usr/bin/env python
"""2-GPU check of the tensor-parallel BinaryDiff linears (config 5 style): column-parallel then row-parallel with one NCCL
all-reduce, against the unsharded single-GPU result.  Launch: python -m torch.distributed.run --nproc-per-node 2 tools/tp_check.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bitdelta_b200 as bd
from bitdelta_b200.parallel import TensorParallelDiffLinear

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
T, m, H, I = 8, 1, 8192, 28672 // 4   # Llama-2-70B hidden, a quarter of its MLP width to keep the check quick
g = torch.Generator(device=dev).manual_seed(0)  # identical full tensors on every rank
w_up = (torch.randn(I, H, generator=g, device=dev) * 0.02).bfloat16()
m_up = torch.randint(-(2**31), 2**31 - 1, (T, H // 32, I), generator=g, device=dev, dtype=torch.int64).to(torch.int32)
w_dn = (torch.randn(H, I, generator=g, device=dev) * 0.02).bfloat16()
m_dn = torch.randint(-(2**31), 2**31 - 1, (T, I // 32, H), generator=g, device=dev, dtype=torch.int64).to(torch.int32)
c_up = torch.rand(T, generator=g, device=dev) * 0.002 + 0.001
c_dn = torch.rand(T, generator=g, device=dev) * 0.002 + 0.001
x = torch.randn(T, m, H, generator=g, device=dev).bfloat16()
up = TensorParallelDiffLinear.from_full(w_up, m_up, c_up, "column", rank, world)
dn = TensorParallelDiffLinear.from_full(w_dn, m_dn, c_dn, "row", rank, world)
h_local = up(x)                                  # [T, m, I / world]
y = dn(torch.nn.functional.silu(h_local))        # one all-reduce inside
# unsharded reference on this GPU
lin_up = torch.nn.Linear(H, I, bias=False, device=dev, dtype=torch.bfloat16); lin_up.weight.data = w_up
lin_dn = torch.nn.Linear(I, H, bias=False, device=dev, dtype=torch.bfloat16); lin_dn.weight.data = w_dn
h_full = bd.DiffCompressModule(lin_up, m_up, c_up)(x)
y_full = bd.DiffCompressModule(lin_dn, m_dn, c_dn)(torch.nn.functional.silu(h_full))
i0, i1 = rank * I // world, (rank + 1) * I // world
err_col = (h_local.float() - h_full[..., i0:i1].float()).abs().max().item()
rel_row = ((y.float() - y_full.float()).abs().mean() / y_full.float().abs().mean()).item()
# the shards use a different stream-K partition than the unsharded launch, so fp32 summation order (and thus the last bf16
# bit of a few outputs) may differ
rel_col = ((h_local.float() - h_full[..., i0:i1].float()).abs().mean() / h_full.float().abs().mean()).item()
ok = rel_col < 1e-3 and rel_row < 2e-3
# timing of the row-parallel layer incl. all-reduce
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
hh = torch.nn.functional.silu(h_local)
for _ in range(5): dn(hh)
e0.record()
for _ in range(20): dn(hh)
e1.record(); torch.cuda.synchronize()
res = {"rank": rank, "world": world, "column_parallel_max_abs_diff": err_col, "column_parallel_mean_rel": rel_col, "row_parallel_mean_rel_vs_unsharded": rel_row, "ok": ok,
       "row_parallel_layer_us_incl_allreduce": e0.elapsed_time(e1) * 1e3 / 20}
print(json.dumps(res), flush=True)
dist.barrier(); dist.destroy_process_group()
if rank == 0:
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(res, open('gpurun_out/tp_check.json', 'w'))
sys.exit(0 if ok else 1)
