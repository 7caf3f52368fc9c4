#!/usr/bin/env python
"""Benchmark of the BitDelta hot path on B200: Mistral-7B + 6 deltas, batched decode (BASELINE.json metric).

A "step" is one decode step's worth of the hot path: all 224 BinaryDiff linears of Mistral-7B (32 layers x
q/k/v/o/gate/up/down) evaluated for 6 tenants x 1 new token through `DiffCompressModule.forward`
(demo/demo_backend.py:93-98 semantics): q/k/v and gate/up, which a decoder layer calls back to back on the same hidden
states, share one launch (128 launches per step; --no-group gives 224), reading 13.96 GB of bf16 base weights and
6 x 0.87 GB of sign words.  Attention, norms and the per-tenant lm_heads are not part of the W1A16 path and are not
executed; the config says so.  Weights are random-init, inputs synthetic.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 (torchrun, one rank per GPU): every rank holds a replica of W_base and serves its own 6 tenants (tenant sharding,
no collective on the data path) -> weak scaling; value = 6*N tokens / max-over-ranks step time.

`--impl reference` times the reference's CPU path for the same workload (the oracle's C port of the unpack-matmul
path, all host threads) on a bounded sample -- one decoder layer per step -- and scales it to a whole decode step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Mistral-7B decoder layer: (name, N_out, K_in)
MISTRAL_LINEARS = [
    ("q_proj", 4096, 4096), ("k_proj", 1024, 4096), ("v_proj", 1024, 4096), ("o_proj", 4096, 4096),
    ("gate_proj", 14336, 4096), ("up_proj", 14336, 4096), ("down_proj", 4096, 14336),
]
LAYERS = 32
TENANTS = 6


def step_bytes(T: int, m: int, layers: int = LAYERS) -> int:
    """Algorithmic bytes of one step: 2NK + T*NK/8 + 2*T*m*(K+N) per linear."""
    per_layer = sum(2 * n * k + T * n * k // 8 + 2 * T * m * (k + n) for _, n, k in MISTRAL_LINEARS)
    return per_layer * layers


def step_flops(T: int, m: int, layers: int = LAYERS) -> int:
    return sum(4 * T * m * n * k for _, n, k in MISTRAL_LINEARS) * layers


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), "measured"
        except Exception:
            pass
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None

    def start(self, wait_first: float = 3.0):
        """Starts nvidia-smi -lms 20 and returns once its first sample has arrived (its start-up can take longer than a short
        timed region), so the loop is sampling for the whole of the region that follows."""
        import threading

        self.lines = []
        self._first = threading.Event()
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def reader():
            for line in self.proc.stdout:
                self.lines.append(line)
                self._first.set()

        self._thread = threading.Thread(target=reader, daemon=True)
        self._thread.start()
        self._first.wait(wait_first)

    def mark(self):
        """Drops the samples taken so far (idle GPU, warm-up): called right before the timed region."""
        if self.proc is not None:
            self.lines.clear()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self._thread.join(timeout=2)
        out = "".join(self.lines)
        sm, smax, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = the upper half of the samples (the sampler also sees idle time around the region)
        sm.sort()
        load = sm[len(sm) // 2:]
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
def cpu_layer_problems(seed: int = 0):
    """Synthetic operands of ONE Mistral decoder layer (7 linears, 6 tenants, 1 token each) for the CPU baseline."""
    import numpy as np

    rng = np.random.default_rng(seed)
    probs = []
    for _, n, k in MISTRAL_LINEARS:
        x = (rng.integers(0, 1 << 14, (TENANTS, 1, k)).astype(np.uint16)) | 0x3C00       # bf16 in [1, 2)
        w = (rng.integers(0, 1 << 14, (n, k)).astype(np.uint16)) | 0x3800
        masks = rng.integers(-(2**31), 2**31 - 1, (TENANTS, k // 32, n)).astype(np.int32)
        coeff = np.full(TENANTS, 0.002, np.float32)
        probs.append((x, w, masks, coeff))
    return probs


def cpu_layer_sample(threads: int = 0, reps: int = 1, seed: int = 0, probs=None):
    """Times the oracle's C port on ONE Mistral decoder layer on the host cores.  Returns (seconds per layer, cores used).
    Only bench.py's baseline legs may execute oracle/ code."""
    from oracle import c_oracle as C

    fresh = probs is None
    if fresh:
        probs = cpu_layer_problems(seed)
    cores = threads or C.max_threads()
    threads = cores  # explicit: OpenMP's own default ignores the cgroup CPU quota
    if fresh:
        C.fwd_batched_bf16(*probs[1], threads=threads)  # warm the thread pool / page in
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        for p in probs:
            C.fwd_batched_bf16(*p, threads=threads)
        times.append(time.perf_counter() - t0)
    return min(times), cores


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    t_layers = []
    probs = cpu_layer_problems(0)  # generated once (600 MB): a step is one pass of the C port over the layer's operands
    for i in range(args.warmup + args.steps):
        t, cores = cpu_layer_sample(threads=0, reps=1, probs=probs)
        if i >= args.warmup:
            t_layers.append(t)
    t_step = statistics.mean(t_layers) * LAYERS
    value = TENANTS / t_step
    sample = f"1 of {LAYERS} decoder layers per step (7 linears x {TENANTS} tenants x 1 token, operands resident), scaled x{LAYERS}"
    line = {
        "impl": "reference", "metric": "tokens/sec Mistral-7B+6delta batched decode (BinaryDiff linears)", "value": value,
        "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(1),
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(world: int):
    return {
        "workload": "mistral7b_6delta_decode_linears",
        "detail": f"{LAYERS} layers x 7 BinaryDiff linears, {TENANTS} tenants x 1 token per GPU, hidden 4096 / inter 14336 / kv 1024",
        "tenants_per_gpu": TENANTS, "tokens_per_step_per_gpu": TENANTS, "parallelism": f"tenant-sharded x{world} (no collective)",
        "l2_policy": "inputs larger than L2 (19.2 GB of weights+signs streamed per step)",
        "not_included": "attention, norms, per-tenant lm_head (outside the W1A16 path)",
    }


# ----------------------------------------------------------------------------------------------------------------------
def load_reference_gpu_modules():
    """The reference's OWN GPU path, unmodified, from baseline/_ref (pip-installed from /root/reference):
    `binary_bmm` (Triton, bitdelta/binary_gemm_kernel.py:297-335) and the `DiffCompressModule` class body of
    demo/demo_backend.py:82-98 (the demo file itself cannot be imported: it loads Mistral-7B at module level, :23)."""
    import ast
    import importlib.util

    import torch
    import torch.nn as nn

    ref_root = os.path.join(ROOT, "baseline", "_ref")
    spec = importlib.util.spec_from_file_location("_ref_binary_gemm_kernel", os.path.join(ref_root, "bitdelta", "binary_gemm_kernel.py"))
    kmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(kmod)
    src = open(os.path.join(ref_root, "demo", "demo_backend.py")).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "DiffCompressModule")
    ns = {"torch": torch, "nn": nn, "binary_bmm": kmod.binary_bmm}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "baseline/_ref/demo/demo_backend.py", "exec"), ns)
    return ns["DiffCompressModule"], kmod


def run_triton_reference(torch, mods, xs, y_ours, stream, dev, steps: int):
    """Times the reference's DiffCompressModule stack (cuBLAS nn.Linear + Triton binary_bmm + two pointwise kernels per
    linear) on the SAME weights, sign words and activations as our arm, under the same CUDA-graph replay harness."""
    try:
        RefModule, _ = load_reference_gpu_modules()
    except Exception as e:  # baseline/_ref missing or triton import failure
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    x_h, x_a, x_m = xs
    try:
        with torch.cuda.stream(stream):
            ref = [{name: RefModule(m.module, m.mask, m.coeff) for name, m in layer.items()} for layer in mods]

            def step():
                y = None
                for layer in ref:
                    layer["q_proj"](x_h); layer["k_proj"](x_h); layer["v_proj"](x_h)
                    layer["o_proj"](x_a)
                    layer["gate_proj"](x_h); layer["up_proj"](x_h)
                    y = layer["down_proj"](x_m)
                return y

            t0 = time.perf_counter()
            y_ref = step()  # Triton autotunes each (M, N, K) on its first call
            torch.cuda.synchronize(dev)
            tune_s = time.perf_counter() - t0
            step()
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                y_static = step()
            for _ in range(2):
                graph.replay()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                graph.replay()
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            rel = ((y_ours.float() - y_ref.float()).abs().mean() / y_ref.float().abs().mean()).item()
        return {"ms_per_step": ms, "tokens_s": TENANTS / (ms * 1e-3), "steps": steps, "harness": "CUDA graph replay, same weights / signs / activations as `value`",
                "path": "baseline/_ref: demo_backend.DiffCompressModule (nn.Linear cuBLAS + Triton binary_bmm + 2 pointwise), 224 linears per step, ungrouped",
                "launches_per_step": 224 * 4, "rel_ours_vs_reference": rel, "autotune_s": tune_s}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


def _group(bd, layer):
    # what bd.fuse_sibling_projections does on a real decoder layer: projections the layer calls back to back on the same
    # hidden states share one launch
    bd.group_projections([layer["q_proj"], layer["k_proj"], layer["v_proj"]])
    bd.group_projections([layer["gate_proj"], layer["up_proj"]])


def build_model(torch, bd, dev, layers: int, seed: int, grouped: bool = True, tenants: int = TENANTS):
    """Random-init Mistral-7B-shaped stack of DiffCompressModules with `tenants` deltas on `dev`."""
    gen = torch.Generator(device=dev).manual_seed(seed)
    mods = []
    for _ in range(layers):
        layer = {}
        for name, n, k in MISTRAL_LINEARS:
            lin = torch.nn.Linear(k, n, bias=False, device=dev, dtype=torch.bfloat16)
            with torch.no_grad():
                lin.weight.normal_(0.0, 0.02, generator=gen)
            masks = torch.randint(-(2**31), 2**31 - 1, (tenants, k // 32, n), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
            coeffs = (torch.rand(tenants, generator=gen, device=dev) * 0.002 + 0.001).to(torch.bfloat16)
            layer[name] = bd.DiffCompressModule(lin, masks, coeffs)
        if grouped:
            _group(bd, layer)
        mods.append(layer)
    return mods


def tenant_slice(bd, mods, tenants: int, grouped: bool = True):
    """The same stack serving only its first `tenants` deltas: new wrappers over leading slices (views) of the stacked sign
    words and coefficients -- what a rank of a tenant-sharded deployment holds."""
    out = []
    for layer in mods:
        sub = {name: bd.DiffCompressModule(m.module, m.mask[:tenants], m.coeff[:tenants]) for name, m in layer.items()}
        if grouped:
            _group(bd, sub)
        out.append(sub)
    return out


def mistral_step(layer_mods, x_h, x_a, x_m):
    y = None
    for layer in layer_mods:
        layer["q_proj"](x_h); layer["k_proj"](x_h); layer["v_proj"](x_h)
        layer["o_proj"](x_a)
        layer["gate_proj"](x_h); layer["up_proj"](x_h)
        y = layer["down_proj"](x_m)
    return y


def graph_time_ms(torch, fn, stream, dev, steps: int, warmup: int, barrier=None):
    """Captures fn() in a CUDA graph on `stream` (fn has been run eagerly before) and returns (ms per replay, result)."""
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=stream):
        out = fn()
    for _ in range(warmup):
        graph.replay()
    (barrier or (lambda: torch.cuda.synchronize(dev)))()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        graph.replay()
    e1.record(stream)
    (barrier or (lambda: torch.cuda.synchronize(dev)))()
    return e0.elapsed_time(e1) / steps, out


# Llama-2-70B decoder layer (BASELINE config 5): (name, N_out, K_in, tensor-parallel mode)
LLAMA70B_LINEARS = [
    ("q_proj", 8192, 8192, "column"), ("k_proj", 1024, 8192, "column"), ("v_proj", 1024, 8192, "column"), ("o_proj", 8192, 8192, "row"),
    ("gate_proj", 28672, 8192, "column"), ("up_proj", 28672, 8192, "column"), ("down_proj", 8192, 28672, "row"),
]
TP_TENANTS = 8


def run_tp(torch, bd, dist, dev, rank: int, world: int, stream, steps: int, warmup: int, barrier, tp_layers: int):
    """BASELINE config 5: Llama-2-70B + 8 deltas decoder layers, every BinaryDiff linear split Megatron-style over the
    `world` ranks (column-parallel q/k/v/gate/up: slices of N; row-parallel o/down: slices of K) with ONE sum all-reduce
    of the fp32 partials after each row-parallel linear (2 per layer).  Times `tp_layers` layers per step (weights + signs
    per rank exceed L2), with and without the collectives, and checks layer 0 against the unsharded modules."""
    from bitdelta_b200.diff import _fused_forward, _fused_forward_grouped
    from bitdelta_b200.parallel import PeerExchange, TensorParallelDiffLinear, split_column_parallel, split_row_parallel

    T = TP_TENANTS
    exchange = PeerExchange(T * 8192) if world > 1 else None
    gen = torch.Generator(device=dev).manual_seed(777)  # same seed on every rank: identical full tensors, sharded locally
    layers, check = [], None
    for li in range(tp_layers):
        layer = {}
        for name, n, k, mode in LLAMA70B_LINEARS:
            w = torch.empty(n, k, device=dev, dtype=torch.bfloat16).normal_(0.0, 0.02, generator=gen)
            masks = torch.randint(-(2**31), 2**31 - 1, (T, k // 32, n), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
            coeffs = (torch.rand(T, generator=gen, device=dev) * 0.002 + 0.001).to(torch.bfloat16)
            ws, ms = (split_column_parallel if mode == "column" else split_row_parallel)(w, masks, rank, world)
            layer[name] = TensorParallelDiffLinear(ws, ms, coeffs, mode)
            if li == 0 and mode == "row":
                check = check or {}
                check[name] = (w, masks, coeffs)
            del w, masks
        layers.append(layer)
    x_h = torch.randn(T, 1, 8192, generator=gen, device=dev).bfloat16()
    x_a_full = torch.randn(T, 1, 8192, generator=gen, device=dev).bfloat16()
    x_m_full = torch.randn(T, 1, 28672, generator=gen, device=dev).bfloat16()
    x_a = x_a_full[..., rank * 8192 // world:(rank + 1) * 8192 // world].contiguous()
    x_m = x_m_full[..., rank * 28672 // world:(rank + 1) * 28672 // world].contiguous()

    torch.cuda.synchronize(dev)  # the operands were generated on the default stream; the legs below run on `stream`

    def col(layer, names, x):
        ms_ = [layer[nm] for nm in names]
        return _fused_forward_grouped(x, [m.weight for m in ms_], [m.mask for m in ms_], [m.coeff for m in ms_], T, "auto", static_operands=True)

    def row(mod, x, how):
        if how == "none":  # the shard's own work only: fp32 partial sums rounded locally, no exchange
            return _fused_forward(x, mod.weight, mod.mask, mod.coeff, T, "auto", static_operands=True, out_fp32=True).to(torch.bfloat16)
        mod.exchange = exchange if how == "peer" else None  # "peer": bd_tp_allreduce over NVLink peer memory; "nccl": dist.all_reduce
        return mod(x)

    def step(how: str):
        y = None
        for layer in layers:
            col(layer, ("q_proj", "k_proj", "v_proj"), x_h)
            row(layer["o_proj"], x_a, how)
            col(layer, ("gate_proj", "up_proj"), x_h)
            y = row(layer["down_proj"], x_m, how)
        return y

    with torch.cuda.stream(stream):
        # parity of the sharded row-parallel linears against the unsharded module (layer 0, every rank holds the full tensors)
        parity = {}
        for name, xf, xs in (("o_proj", x_a_full, x_a), ("down_proj", x_m_full, x_m)):
            w, masks, coeffs = check[name]
            lin = torch.nn.Linear(w.shape[1], w.shape[0], bias=False, device=dev, dtype=torch.bfloat16)
            lin.weight.data = w
            y_full = bd.DiffCompressModule(lin, masks, coeffs)(xf)
            parity[name] = {}
            for how in (("peer", "nccl") if world > 1 else ("nccl",)):
                y_tp = row(layers[0][name], xs, how)
                torch.cuda.synchronize(dev)
                parity[name]["tp1" if world == 1 else how] = {
                    "mean_rel": ((y_tp.float() - y_full.float()).abs().mean() / y_full.float().abs().mean()).item(),
                    "bit_equal_frac": (y_tp == y_full).float().mean().item()}
        check = None
        torch.cuda.empty_cache()
        hows = ["peer", "nccl", "none"] if world > 1 else ["none"]
        ms = {}
        for how in hows:
            step(how)
            torch.cuda.synchronize(dev)
            ms[how], _ = graph_time_ms(torch, lambda: step(how), stream, dev, steps, warmup, barrier)
    t = torch.tensor([ms[h] for h in hows], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = {h: t[i].item() for i, h in enumerate(hows)}
    per_rank_bytes = sum(2 * n * k + T * n * k // 8 for _, n, k, _ in LLAMA70B_LINEARS) // world
    best = "peer" if world > 1 else "none"
    us_layer = ms[best] * 1e3 / tp_layers
    out = {
        "workload": f"Llama-2-70B + {T} deltas decoder layers, tensor-parallel x{world}: column-parallel q/k/v (one launch) and gate/up (one launch), row-parallel o and down, {T} tenants x 1 token",
        "layers_per_step": tp_layers, "us_per_layer": us_layer, "us_per_layer_without_exchange": ms["none"] * 1e3 / tp_layers,
        "exchanges_per_layer": 2 if world > 1 else 0,
        "collective": ("bd_tp_allreduce: one kernel over NVLink peer memory (push fp32 partials [8, 8192] to every rank, flag, sum in rank order, round once)"
                       if world > 1 else "none (tp = 1)"),
        "tokens_s_80_layers": T / (us_layer * 80 * 1e-6), "per_rank_bytes_per_layer": per_rank_bytes,
        "hbm_gbps_per_rank": per_rank_bytes / (us_layer * 1e-6) / 1e9, "hbm_frac_per_rank": per_rank_bytes / (us_layer * 1e-6) / 1e9 / measured_peaks()[0],
        "parity_vs_unsharded": parity,
    }
    if world > 1:
        out["exchange_us_each"] = (ms["peer"] - ms["none"]) * 1e3 / (2 * tp_layers)
        out["exchange_share_of_layer"] = (ms["peer"] - ms["none"]) / ms["peer"]
        out["nccl_variant"] = {"us_per_layer": ms["nccl"] * 1e3 / tp_layers, "allreduce_us_each": (ms["nccl"] - ms["none"]) * 1e3 / (2 * tp_layers),
                               "what": "same shards, dist.all_reduce (NCCL) of the fp32 partials + a rounding kernel"}
        exchange.close()
    return out


def run_model_step(torch, bd, dev, steps: int, warmup: int, prompt_len: int = 64, max_cache_len: int = 512):
    """SURVEY.md 8 row f-3 next to the headline: a WHOLE decode step of a random-init Mistral-7B served to 6 tenants through the
    reference's registration surface (register_diff_compress: DiffCompressModule projections + DataParallelModule embedding /
    norms / lm_heads with the demo's ragged vocabularies, fuse_sibling_projections) -- HF attention (SDPA over a static KV
    cache), RoPE, residuals, SiLU, our fused linears and per-tenant leaves, greedy argmax -- captured as ONE CUDA graph
    (bitdelta_b200.decode.GraphedDecoder; the reference's loop is demo/demo_backend.py:190-258)."""
    import gc

    from transformers import MistralConfig, MistralForCausalLM

    from bitdelta_b200 import _lib
    from bitdelta_b200 import demo_backend as db
    from bitdelta_b200.decode import GraphedDecoder

    vocab = (32000, 32000, 32002, 32002, 32002, 32002)  # the six demo tenants
    cfg = MistralConfig(hidden_size=4096, intermediate_size=14336, num_hidden_layers=LAYERS, num_attention_heads=32, num_key_value_heads=8,
                        vocab_size=vocab[0], max_position_embeddings=4096, sliding_window=None, tie_word_embeddings=False)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(dev):
            model = MistralForCausalLM(cfg).eval()
    finally:
        torch.set_default_dtype(old)
    gen = torch.Generator(device=dev).manual_seed(4242)
    sites = [(f"model.layers.{i}.{blk}.{name}", n, k) for i in range(LAYERS)
             for blk, names in (("self_attn", ("q_proj", "k_proj", "v_proj", "o_proj")), ("mlp", ("gate_proj", "up_proj", "down_proj")))
             for name in names for nm, n, k in MISTRAL_LINEARS if nm == name]
    ckpts = []
    for t in range(TENANTS):
        ck = {}
        for path, n, k in sites:
            ck[path + ".mask"] = torch.randint(-(2**31), 2**31 - 1, (k // 32, n), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
            ck[path + ".coeff"] = (torch.rand((), generator=gen, device=dev) * 0.002 + 0.001).to(torch.bfloat16)
        ck["model.embed_tokens.weight"] = (torch.randn(vocab[t], 4096, generator=gen, device=dev) * 0.02).bfloat16()
        ck["lm_head.weight"] = (torch.randn(vocab[t], 4096, generator=gen, device=dev) * 0.02).bfloat16()
        for i in range(LAYERS):
            for nm in ("input_layernorm", "post_attention_layernorm"):
                ck[f"model.layers.{i}.{nm}.weight"] = (1.0 + 0.05 * torch.randn(4096, generator=gen, device=dev)).bfloat16()
        ck["model.norm.weight"] = (1.0 + 0.05 * torch.randn(4096, generator=gen, device=dev)).bfloat16()
        ckpts.append(ck)
    db.cached_modules.clear()
    out = None
    try:
        db.register_diff_compress(model, ckpts)
        groups = db.fuse_sibling_projections(model)
        ids = torch.randint(1, 32000, (TENANTS, prompt_len), generator=gen, device=dev)
        am = torch.ones_like(ids)
        am[1, :7] = 0  # left padding, like the reference's prompt batching
        am[4, :19] = 0
        ids = ids * am
        torch.cuda.synchronize(dev)
        dec = GraphedDecoder(model, max_cache_len=max_cache_len)
        t0 = time.perf_counter()
        dec.prefill(ids, am)
        torch.cuda.synchronize(dev)
        prefill_ms = (time.perf_counter() - t0) * 1e3
        # eager static-cache steps (Python re-enters every module, like the reference's loop does)
        for _ in range(2):
            dec.step()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(5):
            dec.step()
        torch.cuda.synchronize(dev)
        eager_ms = (time.perf_counter() - t0) * 1e3 / 5
        eager_tokens = dec.step().clone()
        dec.prefill(ids, am)
        n0 = _lib.launch_count()
        dec.capture()
        ours_per_step = (_lib.launch_count() - n0) // 3  # two warm-up steps + the captured one
        for _ in range(7):
            dec.step()
        graph_tokens = dec.step().clone()  # the 8th step after the prefill, like eager_tokens above
        for _ in range(warmup):
            dec.step()
        torch.cuda.synchronize(dev)
        s = torch.cuda.current_stream(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(steps):
            dec.step()
        e1.record(s)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        out = {"ms_per_step": ms, "tokens_s": TENANTS / (ms * 1e-3), "steps": steps,
               "what": "whole decode step of HF MistralForCausalLM (random init, 32 layers) for 6 tenants as one CUDA graph: embedding, RMSNorms, "
                       "SDPA attention over a static KV cache, RoPE, residuals, SiLU, 224 BinaryDiff linears (fused launches), 6 ragged lm_heads, argmax",
               "prompt_len": prompt_len, "kv_positions": max_cache_len, "vocab": list(vocab), "sibling_groups": groups,
               "bitdelta_kernels_per_step": ours_per_step, "eager_static_cache_ms_per_step": eager_ms, "prefill_ms_eager": prefill_ms,
               "graph_equals_eager_tokens": bool(torch.equal(eager_tokens, graph_tokens))}
    except Exception as e:
        out = {"unavailable": f"{type(e).__name__}: {e}"[:400]}
    finally:
        db.unregister_diff_compress(model)
        db.cached_modules.clear()
        del model, ckpts
        gc.collect()
        torch.cuda.empty_cache()
    return out


# BASELINE configs 2 and 4: (name, N_out, K_in) per decoder layer
LLAMA7B_LINEARS = [("q_proj", 4096, 4096), ("k_proj", 4096, 4096), ("v_proj", 4096, 4096), ("o_proj", 4096, 4096),
                   ("gate_proj", 11008, 4096), ("up_proj", 11008, 4096), ("down_proj", 4096, 11008)]
LLAMA13B_LINEARS = [("q_proj", 5120, 5120), ("k_proj", 5120, 5120), ("v_proj", 5120, 5120), ("o_proj", 5120, 5120),
                    ("gate_proj", 13824, 5120), ("up_proj", 13824, 5120), ("down_proj", 5120, 13824)]


def run_other_config(torch, bd, dist, dev, rank, world, stream, barrier, name, linears, layers, tenants_total, steps, warmup, prefill_tokens):
    """The hot path of another BASELINE config: all BinaryDiff linears of a random-init model of that shape, `tenants_total`
    deltas split over the ranks (delta sharding, every rank keeps the whole W_base), one decode step = one new token per
    tenant (CUDA graph, q/k/v and gate/up grouped) and, optionally, a prefill pass of `prefill_tokens` tokens per tenant."""
    if tenants_total % world != 0:
        return None
    T = tenants_total // world
    hid = linears[0][2]
    inter = linears[-1][2]
    gen = torch.Generator(device=dev).manual_seed(31 + rank)
    mods = []
    for _ in range(layers):
        layer = {}
        for nm, n, k in linears:
            lin = torch.nn.Linear(k, n, bias=False, device=dev, dtype=torch.bfloat16)
            with torch.no_grad():
                lin.weight.normal_(0.0, 0.02, generator=gen)
            masks = torch.randint(-(2**31), 2**31 - 1, (T, k // 32, n), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
            coeffs = (torch.rand(T, generator=gen, device=dev) * 0.002 + 0.001).to(torch.bfloat16)
            layer[nm] = bd.DiffCompressModule(lin, masks, coeffs)
        _group(bd, layer)
        mods.append(layer)
    x_h = torch.randn(T, 1, hid, generator=gen, device=dev).bfloat16()
    x_m = torch.randn(T, 1, inter, generator=gen, device=dev).bfloat16()
    torch.cuda.synchronize(dev)
    bytes_step = sum(2 * n * k + T * n * k // 8 + 2 * T * (k + n) for _, n, k in linears) * layers
    out = {"model": name, "layers": layers, "tenants_total": tenants_total, "tenants_per_gpu": T, "n_gpus": world}
    with torch.cuda.stream(stream):
        mistral_step(mods, x_h, x_h, x_m)
        torch.cuda.synchronize(dev)
        ms, _ = graph_time_ms(torch, lambda: mistral_step(mods, x_h, x_h, x_m), stream, dev, steps, warmup, barrier)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t[0].item()
        out["decode"] = {"ms_per_step": ms, "tokens_s": tenants_total / (ms * 1e-3), "bytes_per_gpu_per_step": bytes_step,
                         "hbm_frac_per_gpu": bytes_step / (ms * 1e-3) / 1e9 / measured_peaks()[0]}
        if prefill_tokens:
            xp_h = torch.randn(T, prefill_tokens, hid, generator=gen, device=dev).bfloat16()
            xp_m = torch.randn(T, prefill_tokens, inter, generator=gen, device=dev).bfloat16()
            mistral_step(mods[:1], xp_h, xp_h, xp_m)  # warm-up on one layer
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            mistral_step(mods, xp_h, xp_h, xp_m)
            e1.record(stream)
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            pms = t[0].item()
            flops = sum(4 * tenants_total * prefill_tokens * n * k for _, n, k in linears) * layers
            out["prefill"] = {"tokens_per_tenant": prefill_tokens, "ms": pms, "tflops_all_gpus": flops / (pms * 1e-3) / 1e12,
                              "tokens_s": tenants_total * prefill_tokens / (pms * 1e-3),
                              "frac_of_bf16_peak_per_gpu": flops / world / (pms * 1e-3) / 1e12 / measured_peaks()[1]}
    del mods
    torch.cuda.empty_cache()
    return out


def run_tenant_strong(torch, bd, dist, dev, rank, world, mods, stream, steps, warmup, barrier, grouped, gen):
    """Strong scaling of tenant sharding: a FIXED set of 8 tenants split over the ranks (8 / world each), every rank still
    streaming its whole W_base replica -- the sub-linear curve SURVEY.md 8e predicts (per-GPU bytes 2NK + (T/G) NK/8)."""
    total = 8
    if total % world != 0:
        return None
    t_local = total // world
    sub = tenant_slice(bd, mods, t_local, grouped)
    x_h = torch.randn(t_local, 1, 4096, generator=gen, device=dev).bfloat16()
    x_a = torch.randn(t_local, 1, 4096, generator=gen, device=dev).bfloat16()
    x_m = torch.randn(t_local, 1, 14336, generator=gen, device=dev).bfloat16()
    torch.cuda.synchronize(dev)
    with torch.cuda.stream(stream):
        mistral_step(sub, x_h, x_a, x_m)
        torch.cuda.synchronize(dev)
        ms, _ = graph_time_ms(torch, lambda: mistral_step(sub, x_h, x_a, x_m), stream, dev, steps, warmup, barrier)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t[0].item()
    return {"tenants_total": total, "tenants_per_gpu": t_local, "ms_per_step": ms, "tokens_s": total / (ms * 1e-3), "scaling": "strong",
            "bytes_per_gpu_per_step": step_bytes(t_local, 1, len(mods)),
            "hbm_frac_per_gpu": step_bytes(t_local, 1, len(mods)) / (ms * 1e-3) / 1e9 / measured_peaks()[0]}


def run_ours(args, rank: int, local_rank: int, world: int):
    import torch

    import bitdelta_b200 as bd
    from bitdelta_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU baseline)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    layers = args.layers
    # the strong-scaling leg serves 8 tenants in total: only a single GPU needs more than the headline's 6 resident
    t_build = max(TENANTS, 8 // world) if not args.no_extras else TENANTS
    mods_all = build_model(torch, bd, dev, layers, seed=1234 + rank, grouped=not args.no_group, tenants=t_build)
    mods = mods_all if t_build == TENANTS else tenant_slice(bd, mods_all, TENANTS, grouped=not args.no_group)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    # static synthetic activations (one new token per tenant); outputs are not chained because the omitted norms
    # would be needed to keep magnitudes bounded
    x_h = torch.randn(TENANTS, 1, 4096, generator=gen, device=dev).bfloat16()
    x_a = torch.randn(TENANTS, 1, 4096, generator=gen, device=dev).bfloat16()
    x_m = torch.randn(TENANTS, 1, 14336, generator=gen, device=dev).bfloat16()
    host_in = [t.cpu().pin_memory() for t in (x_h, x_a, x_m)]
    host_out = torch.empty(TENANTS, 1, 4096, dtype=torch.bfloat16).pin_memory()

    def step():
        return mistral_step(mods, x_h, x_a, x_m)

    stream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize(dev)  # the model was generated on the default stream
    with torch.cuda.stream(stream):
        y_eager = step()  # also allocates the workspace of this stream
        torch.cuda.synchronize(dev)
        n0 = _lib.launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            y_static = step()
        launches_per_step = _lib.launch_count() - n0
        graph.replay()
        torch.cuda.synchronize(dev)
        assert torch.equal(y_static, y_eager), "graph replay disagrees with eager execution"

        def barrier():
            torch.cuda.synchronize(dev)
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize(dev)

        # ---- device-resident timing -------------------------------------------------------------------------
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()  # returns once nvidia-smi delivers samples; the warm-up below keeps the GPU busy meanwhile
        for _ in range(args.warmup):
            graph.replay()
        barrier()
        if rank == 0:
            sampler.mark()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            graph.replay()
        e1.record(stream)
        barrier()
        ms_total = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None

        # ---- end to end: pinned host inputs -> H2D -> step -> D2H of the result, every step ------------------
        def e2e_step():
            x_h.copy_(host_in[0], non_blocking=True)
            x_a.copy_(host_in[1], non_blocking=True)
            x_m.copy_(host_in[2], non_blocking=True)
            graph.replay()
            host_out.copy_(y_static, non_blocking=True)
            stream.synchronize()  # the caller needs the tokens before it can build the next step

        for _ in range(args.warmup):
            e2e_step()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(args.steps):
            e2e_step()
        f1.record(stream)
        barrier()
        ms_e2e_total = f0.elapsed_time(f1)

    # ---- the other half of the headline metric: W1A16 GEMM throughput at prefill size (one BinaryDiff linear, 4096 tokens) ----
    gemm = None
    if not args.no_gemm:  # every rank runs its own replica of the GEMM (no collective): aggregate = sum over the GPUs
        with torch.cuda.stream(stream):
            Mg, Ng, Kg = 4096, 4096, 4096
            lin = mods[0]["q_proj"]
            xg = torch.randn(1, Mg, Kg, generator=gen, device=dev).bfloat16()
            one = bd.DiffCompressModule(lin.module, lin.mask[:1].contiguous(), lin.coeff[:1].contiguous())
            for _ in range(3):
                one(xg)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            for _ in range(10):
                one(xg)
            g1.record(stream)
            torch.cuda.synchronize(dev)
            us_t = torch.tensor([g0.elapsed_time(g1) * 1e3 / 10], device=dev, dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(us_t, op=dist.ReduceOp.MAX)
            us = us_t.item()
            tf = 4.0 * Mg * Ng * Kg / us / 1e6
            gemm = {"shape": f"M={Mg} tokens x N={Ng} x K={Kg}, 1 delta (BinaryDiff prefill), one replica per GPU", "us": us, "tflops": tf,
                    "tflops_all_gpus": tf * world, "n_gpus": world,
                    "flops_counted": "4*M*N*K (base product + sign product, as the reference's notebook counts them)"}

    # ---- next row: the per-tenant dense leaves of the same decode step, each ONE native launch over the tenants ----
    leaves = None
    if rank == 0 and not args.no_gemm:
        with torch.cuda.stream(stream):
            vocab = (32000, 32000, 32002, 32002, 32002, 32002)  # the six demo tenants
            heads = [(torch.randn(v, 4096, generator=gen, device=dev) * 0.02).bfloat16() for v in vocab]
            head = bd.DataParallelModule(torch.nn.Linear(4096, vocab[0], bias=False, device=dev, dtype=torch.bfloat16), heads)
            class MistralRMSNorm(torch.nn.Module):  # the HF module's attributes; its forward is the native kernel's arithmetic
                def __init__(self):
                    super().__init__()
                    self.weight = torch.nn.Parameter(torch.ones(4096, device=dev, dtype=torch.bfloat16))
                    self.variance_epsilon = 1e-5

            norm_mod = MistralRMSNorm()
            norm = bd.DataParallelModule(norm_mod, [torch.ones(4096, device=dev, dtype=torch.bfloat16) for _ in vocab])
            emb = bd.DataParallelModule(torch.nn.Embedding(vocab[0], 4096, device=dev, dtype=torch.bfloat16), heads)
            xh = torch.randn(TENANTS, 1, 4096, generator=gen, device=dev).bfloat16()
            ids = torch.randint(0, 32000, (TENANTS, 1), device=dev)

            def timed(fn, n):  # device time per call: n calls captured in a CUDA graph, replayed
                for _ in range(3):
                    fn()
                torch.cuda.synchronize(dev)
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=stream):
                    for _ in range(n):
                        fn()
                gr.replay()
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record(stream)
                for _ in range(3):
                    gr.replay()
                t1.record(stream)
                torch.cuda.synchronize(dev)
                return t0.elapsed_time(t1) * 1e3 / (3 * n)

            us_head = timed(lambda: head(xh), 20)
            us_norm = timed(lambda: norm(xh), 50)
            us_emb = timed(lambda: emb(ids), 50)
            head_bytes = sum(v * 4096 * 2 for v in vocab) + TENANTS * 4096 * 2 + TENANTS * max(vocab) * 2
            leaves = {"lm_head": {"us": us_head, "algorithmic_bytes": head_bytes, "gbps": head_bytes / us_head / 1e3,
                                  "vocab": list(vocab), "note": "weights (1.57 GB) exceed L2: every launch streams from HBM; a read-only stream can exceed the copy-measured peak"},
                      "rmsnorm_us": us_norm, "embed_us": us_emb,
                      "step_extra_ms": (us_head + 65 * us_norm + us_emb) / 1e3,
                      "note": "not part of `value`: 1 lm_head + 65 RMSNorm + 1 embedding launches per decode step"}
            del heads, head, emb

    # ---- BASELINE configs 4-5 mechanics on this box: strong scaling of tenant sharding, tensor-parallel 70B layers ----
    strong = tp = triton_ref = None
    if not args.no_extras:
        strong = run_tenant_strong(torch, bd, dist, dev, rank, world, mods_all, stream, max(args.steps // 2, 5), args.warmup, barrier,
                                   not args.no_group, gen)
        torch.cuda.empty_cache()
        tp = run_tp(torch, bd, dist, dev, rank, world, stream, max(args.steps, 10), args.warmup, barrier, args.tp_layers)
        torch.cuda.empty_cache()
    other_configs = None
    if not args.no_extras:
        other_configs = []
        # config 2: Llama-2-7B + one delta on one GPU (replicated per rank at N > 1); config 4: Llama-2-13B + 4 deltas,
        # prefill 4096 + decode, delta-sharded over 1 / 2 / 4 GPUs
        if world == 1:
            other_configs.append(run_other_config(torch, bd, dist, dev, rank, world, stream, barrier, "Llama-2-7B + 1 delta (config 2)",
                                                  LLAMA7B_LINEARS, 32, 1, max(args.steps // 2, 5), args.warmup, 2048))
        if world in (1, 2, 4):
            other_configs.append(run_other_config(torch, bd, dist, dev, rank, world, stream, barrier, "Llama-2-13B + 4 deltas (config 4)",
                                                  LLAMA13B_LINEARS, 40, 4, max(args.steps // 2, 5), args.warmup, 4096))
        other_configs = [c for c in other_configs if c is not None]
    # ---- the reference's own GPU path (Triton) under the same harness: single-GPU runs only (it is not a scaling leg) ----
    if rank == 0 and world == 1 and not args.no_triton_ref:
        triton_ref = run_triton_reference(torch, mods, (x_h, x_a, x_m), y_static, stream, dev, max(args.steps // 5, 3))
    model_step = None
    if rank == 0 and world == 1 and not args.no_model_step:
        model_step = run_model_step(torch, bd, dev, max(args.steps, 20), args.warmup)

    times = torch.tensor([ms_total, ms_e2e_total], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_step = times[0].item() / args.steps
    ms_e2e = times[1].item() / args.steps

    if rank == 0:
        tokens = TENANTS * world
        value = tokens / (ms_step * 1e-3)
        e2e_value = tokens / (ms_e2e * 1e-3)
        hbm_peak, tf_peak, peak_kind = measured_peaks()
        bytes_step = step_bytes(TENANTS, 1, layers)
        achieved = bytes_step / (ms_step * 1e-3) / 1e9
        kernel = {1: "simt", 2: "umma"}[_lib.lib.bd_select_kernel(0, TENANTS, 1, 4096, 4096, 1)]
        cpu = None
        if not args.no_cpu_baseline:
            t_layer, cores = cpu_layer_sample(threads=0, reps=2)
            cpu = {"value": TENANTS / (t_layer * LAYERS), "unit": "tokens/s", "cores": cores, "kind": "port",
                   "sample": f"1 of {LAYERS} decoder layers (7 linears x {TENANTS} tenants x 1 token), best of 2, scaled x{LAYERS}"}
        cfg = workload_config(world)
        cfg["kernel"] = kernel
        cfg["grouped_launches"] = "q/k/v and gate/up share a launch (SiblingGroup)" if not args.no_group else "off"
        if layers != LAYERS:
            cfg["workload"] += f"_{layers}layers_DEBUG"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu capture
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch_avg")
            except Exception:
                traffic = None
        line = {
            "metric": "tokens/sec Mistral-7B+6delta batched decode (BinaryDiff linears)",
            "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": cfg,
            "e2e": {"value": e2e_value, "unit": "tokens/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": sum(t.numel() * 2 for t in host_in), "d2h_bytes_per_step": host_out.numel() * 2},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch_avg": bytes_step / max(launches_per_step, 1), "peak_kind": peak_kind, "kernel": f"bd fused forward ({kernel})",
                         "algorithmic_bytes_per_step": bytes_step, "launches_per_step": launches_per_step,
                         "w1a16_tflops": step_flops(TENANTS, 1, layers) / (ms_step * 1e-3) / 1e12,
                         "w1a16_frac_of_bf16_peak": step_flops(TENANTS, 1, layers) / (ms_step * 1e-3) / 1e12 / tf_peak},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        if gemm is not None:
            gemm["frac_of_measured_bf16_peak"] = gemm["tflops"] / tf_peak
            gemm["peak_tflops"] = tf_peak
            line["w1a16_gemm"] = gemm
        if leaves is not None:
            leaves["lm_head"]["frac"] = leaves["lm_head"]["gbps"] / hbm_peak
            line["tenant_leaves"] = leaves
        if strong is not None:
            line["tenant_strong_scaling"] = strong
        if other_configs:
            line["other_configs"] = other_configs
        if tp is not None:
            line["tp"] = tp
        if model_step is not None:
            line["model_step"] = model_step
        if triton_ref is not None:
            if "ms_per_step" in triton_ref:
                triton_ref["ours_over_reference"] = triton_ref["ms_per_step"] / ms_step
            line["triton_reference"] = triton_ref
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--layers", type=int, default=LAYERS, help="debug only: fewer layers (the result is labelled DEBUG)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gemm", action="store_true", help="skip the prefill-size W1A16 GEMM measurement")
    ap.add_argument("--no-group", action="store_true", help="one launch per linear (no q/k/v and gate/up grouping)")
    ap.add_argument("--no-extras", action="store_true", help="skip the strong-scaling and tensor-parallel legs")
    ap.add_argument("--no-model-step", action="store_true", help="skip the whole-model decode step leg (single-GPU runs)")
    ap.add_argument("--no-triton-ref", action="store_true", help="skip timing the reference's Triton path (single-GPU runs)")
    ap.add_argument("--tp-layers", type=int, default=4, help="70B decoder layers per step of the tensor-parallel leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
