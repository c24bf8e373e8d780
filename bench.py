#!/usr/bin/env python3
"""Headline benchmark: attention forward TFLOP/s (hd=128) on N B200s -- BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c5]

N > 1 is launched by the driver as `python -m torch.distributed.run --nproc-per-node N ... bench.py
--gpus N ...` (one rank per GPU). (batch, head) problems are independent, so ranks share nothing on
the data path: each rank owns its own batch shard (weak scaling: the per-GPU workload is fixed), NCCL
is used only for the barriers and the MAX-reduce of the elapsed time.

A "step" is one forward pass of the hot path over one batch of synthetic input:
  workload c2 (default) = BASELINE config[1]: bf16, batch 8, 32 heads, seqlen 4096, head_dim 128,
  causal (Llama-3-8B attention shape) per GPU.
FLOP convention: 4 * D * (unmasked q-k pairs), causal counted as S*S/2.

JSON keys beyond the base contract: `roofline` (dominant kernel vs measured bf16 peak),
`cpu_baseline` (the reference's own CPU attention, oracle/_ref, on a bounded sample), `e2e`
(host pinned buffers -> H2D -> kernel -> D2H inside the timed region), `clocks`, `gpu_launches`,
`sustained` (the headline step back to back for >= 2.5 s against the sustained measured peak) and `configs`
(the other BASELINE.json workloads, each with value / ms / roofline: C1, C3, C4 at B = 1 / 8 / 64 on rank 0, and
C5 -- global batch 64 sharded over the N ranks, strong scaling, device-timed max over ranks).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "flash-attention-v100_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "attention_fwd_tflops_hd128"
UNIT = "TFLOP/s"

WORKLOADS = {
    # name: (batch per GPU, heads, kv heads, seqlen, head_dim, causal, window)
    "c2": dict(batch=8, heads=32, heads_k=32, seqlen=4096, head_dim=128, causal=True, window=(-1, -1),
               desc="bf16 B=8 H=32 Hk=32 S=4096 D=128 causal per GPU (BASELINE config 2, Llama-3-8B shape)"),
    "c2gqa": dict(batch=8, heads=32, heads_k=8, seqlen=4096, head_dim=128, causal=True, window=(-1, -1),
                  desc="bf16 B=8 H=32 Hk=8 S=4096 D=128 causal per GPU (config 2 with Llama-3-8B GQA)"),
    "d256": dict(batch=8, heads=16, heads_k=16, seqlen=4096, head_dim=256, causal=True, window=(-1, -1),
                 desc="bf16 B=8 H=16 Hk=16 S=4096 D=256 causal per GPU (SURVEY 8(f)-3: the 256-wide tile)"),
    "c5": dict(batch=64, heads=32, heads_k=32, seqlen=8192, head_dim=128, causal=True, window=(4096, 0), strong=True,
               desc="bf16 B=64 (global, sharded over the ranks) H=32 S=8192 D=128 causal + window 4096 (BASELINE config 5)"),
}


def algorithmic_flops(w) -> float:
    S, D = w["seqlen"], w["head_dim"]
    wl = w["window"][0]
    if wl >= 0 and wl < S:
        pairs = wl * (wl + 1) // 2 + (S - wl) * (wl + 1)  # exact count, SURVEY 8d (config 5)
    elif w["causal"]:
        pairs = S * S // 2  # FA convention
    else:
        pairs = S * S
    return 4.0 * D * pairs * w["batch"] * w["heads"]


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["bf16_tflops"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    return 1590.0, 1400.0, "fallback"


class ClockSampler:
    """Samples SM clock, power and throttle reasons through NVML while the timed region runs
    (every ~2 ms; `nvidia-smi` is far too slow for a region of tens of milliseconds)."""

    def __init__(self, index: int):
        self.index = index
        self.sm, self.power, self.reasons = [], [], set()
        self.sm_max = None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self._nvml = None

    def _run(self):
        n = self._nvml
        names = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown if hasattr(n, "nvmlClocksEventReasonHwSlowdown") else n.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", None) or n.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", None) or n.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", None) or n.nvmlClocksThrottleReasonSwPowerCap}
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
                self.power.append(n.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
                mask = get_reasons(self._h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.002)

    def __enter__(self):
        if self._nvml is not None:
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._nvml is not None:
            self._t.join(timeout=2)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.sm), "sm_min_mhz": min(self.sm), "sm_max_mhz": self.sm_max,
                "power_w_max": max(self.power) if self.power else None, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


# ------------------------------------------------------------------------------------------- reference arm
def gpu_local_cpus(index: int):
    """CPUs NVML reports as local to GPU `index` (its NUMA node), restricted to the CPUs this process may run on;
    None when the box gives no topology (no NVML, a VM without NUMA information, an empty intersection)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:  # noqa: BLE001
        return None


class numa_local:
    """Runs the calling thread on the CPUs local to a GPU for the duration of the block, so that pinned host buffers
    allocated inside it land on that GPU's NUMA node (first touch). With 8 ranks on a two-socket host, buffers on the
    far socket put every H2D / D2H copy on the inter-socket link. Affinity is restored on exit; a no-op without
    topology information."""

    def __init__(self, index: int):
        self.cpus = gpu_local_cpus(index)
        self.prev = None
        self.applied = False

    def __enter__(self):
        try:
            self.prev = os.sched_getaffinity(0)
            if self.cpus and self.cpus != self.prev:
                os.sched_setaffinity(0, self.cpus)
                self.applied = True
        except Exception:  # noqa: BLE001
            self.applied = False
        return self

    def __exit__(self, *a):
        if self.applied:
            try:
                os.sched_setaffinity(0, self.prev)
            except Exception:  # noqa: BLE001
                pass
        return False


def cpu_reference_sample(w, threads: int):
    """One bounded sample of the workload on the host cores with the reference's own CPU attention
    (oracle/_ref, compiled from reference utils/sass/mma_swizzle/forward_kernel.cu:346-370); falls back to
    the oracle's C port. Sample = `threads` (batch, head) problems of the workload's shape, one per thread."""
    import numpy as np

    from oracle import native

    S, D = w["seqlen"], w["head_dim"]
    heads = max(1, threads)
    rng = np.random.default_rng(421)
    q = rng.standard_normal((heads, S, D), dtype=np.float32)
    k = rng.standard_normal((heads, S, D), dtype=np.float32)
    v = rng.standard_normal((heads, S, D), dtype=np.float32)
    windowed = w["window"][0] >= 0
    if native.ref_available() and not windowed:
        kind = "reference"
        t0 = time.perf_counter()
        native.ref_cpu_attention(q, k, v, D ** -0.5, w["causal"], threads=threads)
        dt = time.perf_counter() - t0
    else:
        kind = "port"  # the reference's cpu_attention has no window: use the C restatement
        qq = np.ascontiguousarray(q.transpose(1, 0, 2))
        kk = np.ascontiguousarray(k.transpose(1, 0, 2))
        vv = np.ascontiguousarray(v.transpose(1, 0, 2))
        t0 = time.perf_counter()
        native.c_oracle_attention(qq, kk, vv, D ** -0.5, w["window"][0], 0 if w["causal"] else w["window"][1],
                                  threads=threads)
        dt = time.perf_counter() - t0
    flops = algorithmic_flops(dict(w, batch=1, heads=heads))
    return flops / dt / 1e12, dt, kind, f"{heads} of {w['batch'] * w['heads']} (batch,head) problems of the workload, one per host thread"


def cpu_sdpa_sample(w, threads: int):
    """The north star's named yardstick beside the reference's own loop: torch SDPA on the host cores, on one batch
    element of the workload (all heads; bf16 inputs upcast to fp32 as the reference's test.py:18-34 does)."""
    import torch
    import torch.nn.functional as F

    S, D, H, Hk = w["seqlen"], w["head_dim"], w["heads"], w["heads_k"]
    torch.manual_seed(421)
    q = torch.randn(1, H, S, D, dtype=torch.float32)
    k = torch.randn(1, Hk, S, D, dtype=torch.float32)
    v = torch.randn(1, Hk, S, D, dtype=torch.float32)
    mask = None
    if w["window"][0] >= 0:  # causal + left window as an explicit boolean mask
        i = torch.arange(S)[:, None]
        j = torch.arange(S)[None, :]
        mask = (j <= i) & (j >= i - w["window"][0])
    kwargs = dict(attn_mask=mask) if mask is not None else dict(is_causal=bool(w["causal"]))
    if mask is not None:  # the masked path materialises the scores: keep the sample to one GQA group / 4 heads
        g = H // Hk
        hs = max(4 // g, 1) * g
        q, k, v, H, Hk = q[:, :hs], k[:, :hs // g], v[:, :hs // g], hs, hs // g
    if Hk != H:
        kwargs["enable_gqa"] = True
    torch.set_num_threads(threads)
    F.scaled_dot_product_attention(q[:, :, :256], k[:, :, :256], v[:, :, :256], is_causal=True, **({"enable_gqa": True} if Hk != H else {}))
    t0 = time.perf_counter()
    F.scaled_dot_product_attention(q, k, v, **kwargs)
    dt = time.perf_counter() - t0
    flops = algorithmic_flops(dict(w, batch=1, heads=H))
    return {"value": flops / dt / 1e12, "unit": UNIT, "cores": threads, "kind": "torch SDPA fp32 on the host",
            "sample": f"1 of {w['batch']} batch elements of the workload, {H} of {w['heads']} heads", "sample_seconds": dt}


def run_reference(args, w, rank: int):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for _ in range(min(args.warmup, 1)):
        pass  # a CPU loop needs no warm-up beyond page-in; keep the run bounded
    vals, times = [], []
    steps = max(1, min(args.steps, 3))
    for _ in range(steps):
        tf, dt, kind, sample = cpu_reference_sample(w, threads)
        vals.append(tf)
        times.append(dt)
    value = statistics.median(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": 0, "ms_per_step": statistics.median(times) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "note": "CPU arm: each step is a bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------- the other configs
def _time_events(fn, iters, warm=3):
    """Mean milliseconds per call over `iters` back-to-back calls (CUDA events on the current stream)."""
    import torch

    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def _graph_replay_ms(fn, iters):
    """The same call captured once in a CUDA graph and replayed (what a serving / training loop would issue)."""
    import torch

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            fn(0)
    torch.cuda.current_stream().wait_stream(side)
    return _time_events(lambda i: g.replay(), iters)


def bench_c1(dev):
    """BASELINE config 1: flash_attn_func fp16 B=2 H=8 S=512 D=64 non-causal (the reference's own test.py case).
    1.07 GFLOP per call: launch-bound by construction, so the record is microseconds per call (eager through the
    Python API, and replayed from a CUDA graph), with the tensor-roofline fraction shown for what it is."""
    import torch
    from flash_attn_v100 import flash_attn_func

    torch.manual_seed(421)
    q, k, v = (torch.randn(2, 512, 8, 64, device=dev, dtype=torch.float16) for _ in range(3))
    fn = lambda i: flash_attn_func(q, k, v)
    ms = _time_events(fn, 200, warm=10)
    gms = _graph_replay_ms(fn, 200)
    flops = 4.0 * 64 * 512 * 512 * 2 * 8
    peak, _, _ = measured_peaks()
    return {"workload": "C1 fp16 B=2 H=8 S=512 D=64 non-causal, flash_attn_func", "value": flops / ms / 1e9, "unit": UNIT,
            "ms": ms, "us_per_call_eager": ms * 1e3, "us_per_call_graph": gms * 1e3, "value_graph": flops / gms / 1e9,
            "roofline": {"bound": "launch latency (1.07 GFLOP per call)", "frac": flops / gms / 1e9 / peak, "peak": peak,
                         "unit": "TFLOP/s", "note": "fraction of the bf16/fp16 tensor peak at graph-replay latency"}}


def bench_c3(dev):
    """BASELINE config 3: flash_attn_varlen_func bf16, 64 packed sequences (randint(1, 2049), seed 0), H=32, D=128,
    causal. FLOPs = 4 D H sum_i s_i (s_i + 1) / 2."""
    import torch
    from flash_attn_v100 import flash_attn_varlen_func

    lens = torch.randint(1, 2049, (64,), generator=torch.Generator().manual_seed(0))
    H, D = 32, 128
    T = int(lens.sum())
    torch.manual_seed(421)
    q = torch.randn(T, H, D, device=dev, dtype=torch.bfloat16)
    k, v = torch.randn_like(q), torch.randn_like(q)
    cu = torch.nn.functional.pad(lens.cumsum(0), (1, 0)).int().to(dev)
    mx = int(lens.max())
    ms = _time_events(lambda i: flash_attn_varlen_func(q, k, v, cu, cu, mx, mx, causal=True), 30)
    flops = 4.0 * D * H * float((lens.double() * (lens.double() + 1) / 2).sum())
    peak, _, _ = measured_peaks()
    return {"workload": f"C3 bf16 varlen 64 packed sequences <= 2048 ({T} tokens) H=32 D=128 causal", "value": flops / ms / 1e9,
            "unit": UNIT, "ms": ms, "roofline": {"bound": "tensor", "achieved": flops / ms / 1e9, "peak": peak,
                                                 "unit": "TFLOP/s", "frac": flops / ms / 1e9 / peak}}


def bench_widened(dev):
    """SURVEY 8(f) rows at the config-2 shape, for the driver record: the same shape non-causal, the backward (dense,
    through autograd: row-dot + dK/dV pass + dQ pass; 10 D FLOP per unmasked pair = 2.5 x the forward), and the
    score-modifier variants of the forward. bf16 B=8 H=32 S=4096 D=128."""
    import torch
    from flash_attn_v100 import flash_attn_func

    B, S, H, D = 8, 4096, 32, 128
    torch.manual_seed(421)
    q, k, v = (torch.randn(B, S, H, D, device=dev, dtype=torch.bfloat16, requires_grad=True) for _ in range(3))
    do = torch.randn(B, S, H, D, device=dev, dtype=torch.bfloat16)
    slopes = (torch.rand(H, device=dev) * 0.2).float()
    peak, _, _ = measured_peaks()
    fl = 4.0 * D * B * H * S * S

    def rec(ms, flops, bound="tensor"):
        return {"value": flops / ms / 1e9, "unit": UNIT, "ms": ms, "roofline": {"bound": bound, "frac": flops / ms / 1e9 / peak, "peak": peak}}

    out = {}
    with torch.no_grad():
        out["fwd_noncausal"] = rec(_time_events(lambda i: flash_attn_func(q, k, v), 10), fl)
        out["fwd_alibi"] = rec(_time_events(lambda i: flash_attn_func(q, k, v, causal=True, alibi_slopes=slopes), 10), fl / 2)
        out["fwd_softcap30"] = rec(_time_events(lambda i: flash_attn_func(q, k, v, causal=True, softcap=30.0), 10), fl / 2)
        out["fwd_window1024"] = rec(_time_events(lambda i: flash_attn_func(q, k, v, causal=True, window_size=(1024, 0)), 10),
                                    4.0 * D * B * H * (1024 * 1025 // 2 + (S - 1024) * 1025))
        out["fwd_dropout0.1"] = rec(_time_events(lambda i: flash_attn_func(q, k, v, causal=True, dropout_p=0.1), 5), fl / 2)
    o = flash_attn_func(q, k, v, causal=True)
    out["bwd_causal"] = rec(_time_events(lambda i: torch.autograd.grad(o, (q, k, v), do, retain_graph=True), 10), 2.5 * fl / 2)
    del q, k, v, o, do
    # the other head dims at the same token count and model width (H * D = 4096): bf16 causal forward
    with torch.no_grad():
        for d in (64, 256):
            h = 4096 // d
            qd, kd, vd = (torch.randn(B, S, h, d, device=dev, dtype=torch.bfloat16) for _ in range(3))
            out[f"fwd_causal_d{d}"] = rec(_time_events(lambda i: flash_attn_func(qd, kd, vd, causal=True), 10), 4.0 * d * B * h * S * S / 2)
    return {"workload": "config-2 shape (bf16 B=8 H=32 S=4096 D=128): non-causal forward, feature variants, backward; "
                        "head_dim 64 (H=64) and 256 (H=16) causal forward", **out}


def bench_c4(dev, B, Hk=8):
    """BASELINE config 4: flash_attn_with_kvcache bf16 decode, q_seqlen 1, kv_seqlen 8192, 32 heads, D=128, rotary,
    paged block_table (page 256). HBM-bound: achieved GB/s = attended K+V bytes / time. The step is the
    whole public call: append + rotary + split-KV attention + combine. Caches rotate so L2 cannot hold them."""
    import torch
    from flash_attn_v100 import flash_attn_with_kvcache

    dt, H, D, Sk, page = torch.bfloat16, 32, 128, 8192, 256
    n_pages = B * (Sk // page)
    n_caches = max(2, min(8, int(3 * 2 ** 30 // (n_pages * page * Hk * D * 2 * 2)) + 1))
    caches = [(torch.randn(n_pages, page, Hk, D, device=dev, dtype=dt), torch.randn(n_pages, page, Hk, D, device=dev, dtype=dt))
              for _ in range(n_caches)]
    bt = torch.randperm(n_pages, generator=torch.Generator().manual_seed(0)).view(B, -1).int().to(dev)
    lens = torch.full((B,), Sk - 1, dtype=torch.int32, device=dev)
    q = torch.randn(B, 1, H, D, device=dev, dtype=dt)
    kn, vn = (torch.randn(B, 1, Hk, D, device=dev, dtype=dt) for _ in range(2))
    inv = 1.0 / (10000 ** (torch.arange(0, D, 2, dtype=torch.float32) / D))
    ang = torch.outer(torch.arange(Sk, dtype=torch.float32), inv)
    cos, sin = ang.cos().to(dt).to(dev), ang.sin().to(dt).to(dev)

    def step(i):
        kc, vc = caches[i % n_caches]
        return flash_attn_with_kvcache(q, kc, vc, kn, vn, rotary_cos=cos, rotary_sin=sin, cache_seqlens=lens,
                                       block_table=bt, causal=True, rotary_interleaved=False)

    ms = _time_events(step, 40, warm=2 * n_caches)
    graphs = []
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(n_caches):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                step(i)
            graphs.append(g)
    torch.cuda.current_stream().wait_stream(side)
    gms = _time_events(lambda i: graphs[i % n_caches].replay(), 40, warm=n_caches)
    nbytes = 2.0 * B * Sk * Hk * D * 2
    hbm = 6547.5
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        hbm = float(json.load(open(path)).get("hbm_gbs", hbm))
    return {"workload": f"C4 bf16 decode B={B} Sq=1 Sk=8192 H=32 Hk={Hk} D=128 rotary, paged (page 256)",
            "value": nbytes / gms / 1e6, "unit": "GB/s", "ms": gms, "us_per_step_eager": ms * 1e3, "us_per_step_graph": gms * 1e3,
            "value_eager": nbytes / ms / 1e6,
            "roofline": {"bound": "hbm", "achieved": nbytes / gms / 1e6, "peak": hbm, "unit": "GB/s", "frac": nbytes / gms / 1e6 / hbm,
                         "frac_eager": nbytes / ms / 1e6 / hbm, "note": "graph replay; K+V bytes attended per step / time"}}


def bench_c5(dev, rank, world, dist, barrier, steps=8):
    """BASELINE config 5: bf16 B=64 (global) H=32 S=8192 D=128 causal + sliding window 4096, the batch sharded over the
    ranks (strong scaling: the global batch is fixed). Device-timed, barrier on both sides, max over ranks."""
    import torch
    import sharding
    from flash_attn_v100 import flash_attn_func

    w = WORKLOADS["c5"]
    b0, b1 = sharding.shard_range(w["batch"], world, rank)
    B, H, S, D = b1 - b0, w["heads"], w["seqlen"], w["head_dim"]
    torch.manual_seed(4210 + rank)
    q = torch.randn(B, S, H, D, device=dev, dtype=torch.bfloat16)
    k, v = torch.randn_like(q), torch.randn_like(q)
    fn = lambda: flash_attn_func(q, k, v, causal=True, window_size=w["window"])
    for _ in range(3):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index or 0) as clocks:
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
    ms = sharding.max_over_ranks(e0.elapsed_time(e1), dist, dev) / steps
    flops_all = algorithmic_flops(w)  # the whole global batch
    peak, peak_sus, _ = measured_peaks()
    per_gpu = flops_all / world / ms / 1e9
    return {"workload": w["desc"], "value": flops_all / ms / 1e9, "unit": UNIT, "ms": ms, "n_gpus": world, "scaling": "strong",
            "batch_per_gpu": B, "steps": steps,
            "roofline": {"bound": "tensor", "achieved": per_gpu, "peak": peak, "unit": "TFLOP/s", "frac": per_gpu / peak,
                         "frac_of_sustained": per_gpu / peak_sus,
                         "note": "per GPU; steps of tens of milliseconds run under the power cap, so the sustained peak is the fair denominator"},
            "clocks": clocks.summary()}


def bench_sustained(step, flops_rank, dev, seconds=2.5):
    """The headline step repeated back to back for >= `seconds` (the 30-step headline lasts ~30 ms and never reaches
    the GPU's power state): TFLOP/s against the SUSTAINED measured peak, with its own clock record."""
    import torch

    n = 64
    ms = _time_events(lambda i: step(), n, warm=0)
    iters = max(n, int(seconds * 1e3 / ms) + 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index or 0) as clocks:
        e0.record()
        for _ in range(iters):
            step()
        e1.record()
        torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1)
    _, peak_sus, src = measured_peaks()
    val = flops_rank * iters / (total_ms * 1e-3) / 1e12
    return {"value": val, "unit": UNIT, "steps": iters, "seconds": total_ms * 1e-3, "ms_per_step": total_ms / iters,
            "roofline": {"bound": "tensor", "achieved": val, "peak": peak_sus, "unit": "TFLOP/s", "frac": val / peak_sus,
                         "peak_source": f"{src} bf16 sustained (MEASURED_PEAKS.json)"},
            "clocks": clocks.summary()}


# ------------------------------------------------------------------------------------------- our arm
def run_ours(args, w, rank: int, world: int, local_rank: int):
    import torch

    import flash_attn_v100_cuda as op
    from flash_attn_v100 import flash_attn_func

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback in the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION; stdout must carry only the JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group(backend="nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    import sharding

    strong = bool(w.get("strong"))
    if strong:  # fixed global batch, each rank takes its contiguous slice (config 5)
        b0, b1 = sharding.shard_range(w["batch"], world, rank)
        w = dict(w, batch=b1 - b0)
    B, H, Hk, S, D = w["batch"], w["heads"], w["heads_k"], w["seqlen"], w["head_dim"]
    causal, window = w["causal"], w["window"]
    torch.manual_seed(421 + rank)  # each rank generates its own shard locally
    q = torch.randn(B, S, H, D, device=dev, dtype=torch.bfloat16)
    k = torch.randn(B, S, Hk, D, device=dev, dtype=torch.bfloat16)
    v = torch.randn(B, S, Hk, D, device=dev, dtype=torch.bfloat16)
    flops_rank = algorithmic_flops(w)

    def step():
        return flash_attn_func(q, k, v, causal=causal, window_size=window)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: exactly K steps, CUDA events on the launching (current) stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches0 = op.launch_count()
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev[0].record()
        for i in range(args.steps):
            step()
            ev[i + 1].record()
        barrier()
    launches = op.launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    max_ms = sharding.max_over_ranks(total_ms, dist, dev)
    flops_all = sum(sharding.gather_checksums(flops_rank, dist, dev))  # whole-job work per step
    value = flops_all * args.steps / (max_ms * 1e-3) / 1e12

    # ---- e2e: host pinned buffers -> H2D -> kernel -> D2H through the public API, pipelined over batch elements.
    # "3s" (default): a copy-in stream, a compute stream and a copy-out stream chained by events over NB buffer sets,
    # so the H2D engine runs back to back; "2s": the first pipeline (two streams, each H2D -> kernel -> D2H).
    # tools/e2e_probe.py measures both beside the raw PCIe copy rates.
    # the random data is made first, outside the bound block: torch's intra-op worker threads inherit the affinity of
    # the thread that first needs them, and the CPU baseline below must keep every host core
    host_src = [torch.randn(x.shape, dtype=torch.bfloat16) for x in (q, k, v)]
    with numa_local(local_rank) as numa:
        hq, hk, hv = (t.pin_memory() for t in host_src)
        hout = torch.empty(q.shape, dtype=torch.bfloat16).pin_memory()
    del host_src
    print(f"[bench rank {rank}] pinned e2e buffers: gpu-local cpus={sorted(numa.cpus) if numa.cpus else None} "
          f"bound={numa.applied}", file=sys.stderr, flush=True)
    pipe = os.environ.get("FA_E2E_PIPE", "3s")
    NB = 2 if pipe == "2s" else 3
    dq = [torch.empty_like(q[:1]) for _ in range(NB)]
    dk = [torch.empty_like(k[:1]) for _ in range(NB)]
    dv = [torch.empty_like(v[:1]) for _ in range(NB)]
    if pipe == "2s":
        streams = [torch.cuda.Stream(device=dev) for _ in range(2)]

        def e2e_step():
            for b in range(B):
                st = streams[b % 2]
                with torch.cuda.stream(st):
                    dq[b % 2].copy_(hq[b:b + 1], non_blocking=True)
                    dk[b % 2].copy_(hk[b:b + 1], non_blocking=True)
                    dv[b % 2].copy_(hv[b:b + 1], non_blocking=True)
                    o = flash_attn_func(dq[b % 2], dk[b % 2], dv[b % 2], causal=causal, window_size=window)
                    hout[b:b + 1].copy_(o, non_blocking=True)
            for st in streams:
                st.synchronize()
        e2e_note = "pipelined per batch element on 2 streams"
    else:
        s_in, s_c, s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
        loaded = [torch.cuda.Event() for _ in range(NB)]
        consumed = [torch.cuda.Event() for _ in range(NB)]

        def e2e_step():
            for b in range(B):
                i = b % NB
                with torch.cuda.stream(s_in):
                    if b >= NB:
                        s_in.wait_event(consumed[i])  # the kernel that read this buffer set has finished
                    dq[i].copy_(hq[b:b + 1], non_blocking=True)
                    dk[i].copy_(hk[b:b + 1], non_blocking=True)
                    dv[i].copy_(hv[b:b + 1], non_blocking=True)
                    loaded[i].record(s_in)
                with torch.cuda.stream(s_c):
                    s_c.wait_event(loaded[i])
                    o = flash_attn_func(dq[i], dk[i], dv[i], causal=causal, window_size=window)
                    consumed[i].record(s_c)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(consumed[i])
                    o.record_stream(s_out)
                    hout[b:b + 1].copy_(o, non_blocking=True)
            for st in (s_in, s_c, s_out):
                st.synchronize()
        e2e_note = f"pipelined per batch element: copy-in / compute / copy-out streams over {NB} buffer sets"

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_value = flops_all * e2e_steps / sharding.max_over_ranks(e2e_s, dist, dev) / 1e12
    h2d = sum(x.numel() * 2 for x in (hq, hk, hv))
    d2h = hout.numel() * 2

    # ---- the rest of BASELINE.json's configs and the sustained leg (the headline above is unchanged by them)
    extra = {}
    if not args.no_configs:
        def guarded(name, fn):
            try:
                extra[name] = fn()
            except Exception as e:  # noqa: BLE001  (a side record must never cost the headline line)
                extra[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.synchronize()

        if rank == 0 and args.workload == "c2":
            guarded("sustained", lambda: bench_sustained(step, flops_rank, dev))
        del hq, hk, hv, hout, dq, dk, dv, q, k, v
        torch.cuda.empty_cache()
        if rank == 0:
            guarded("C1", lambda: bench_c1(dev))
            guarded("C3", lambda: bench_c3(dev))
            guarded("widened", lambda: bench_widened(dev))
            for Bd in (1, 8, 64):
                guarded(f"C4_B{Bd}", lambda: bench_c4(dev, Bd))
            torch.cuda.empty_cache()
        if args.workload != "c5":
            # every rank takes part: config 5 is the one workload BASELINE.json shards over the GPUs
            try:
                extra["C5"] = bench_c5(dev, rank, world, dist, barrier)
            except Exception as e:  # noqa: BLE001
                extra["C5"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_sustained, peak_src = measured_peaks()
    kern_ms = statistics.mean(per_launch_ms)
    achieved = flops_rank / (kern_ms * 1e-3) / 1e12
    threads = os.cpu_count() or 1
    cpu_baseline = None  # timed on rank 0 at N=1 only (the scaling runs would repeat the same CPU work)
    if world == 1:
        cpu_tf, cpu_dt, cpu_kind, cpu_sample = cpu_reference_sample(w, threads)
        cpu_baseline = {"value": cpu_tf, "unit": UNIT, "cores": threads, "kind": cpu_kind, "sample": cpu_sample,
                        "sample_seconds": cpu_dt}
        try:
            cpu_baseline["torch_sdpa"] = cpu_sdpa_sample(w, threads)
        except Exception as e:  # noqa: BLE001  (an extra yardstick must never cost the bench line)
            cpu_baseline["torch_sdpa"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02", "traffic_r02.json")
    if not os.path.exists(tpath):
        tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": w["desc"], "global_batch": WORKLOADS[args.workload]["batch"] * (1 if strong else world), "seq_len": S, "heads": H, "heads_k": Hk,
                   "head_dim": D, "parallelism": f"batch-sharded x{world}, no data-path collective",
                   "l2": "inputs+output 1 GiB per step exceed the 126 MB L2 (no flush needed)",
                   "flops_per_step_per_gpu": flops_rank},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": f"{peak_src} bf16 burst (MEASURED_PEAKS.json)",
                     "frac_of_sustained": achieved / peak_sustained, "frac_of_nominal_2250": achieved / 2250.0,
                     "kernel": "fa_fwd_sm100_kernel<128,bf16>", "kernel_ms": kern_ms},
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": "pinned host q,k,v -> H2D, kernel, out -> D2H, " + e2e_note,
                "pinned_numa_local": bool(numa.applied)},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
    }
    if extra:
        line["sustained"] = extra.pop("sustained", None)
        line["configs"] = extra
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    if dist is not None:
        dist.destroy_process_group()


_JSON_OUT = sys.stdout


def _reserve_stdout_for_the_json_line():
    """stdout must carry exactly one JSON line, but native libraries write there too (NCCL prints its version banner on
    stdout whatever NCCL_DEBUG is set to after import). Keep a private handle on the real stdout for the JSON line
    and point file descriptor 1 at stderr for everything else."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def main():
    _reserve_stdout_for_the_json_line()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-configs", action="store_true", help="headline only: skip the C1/C3/C4/C5 and sustained records")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, rank)
        return
    run_ours(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
