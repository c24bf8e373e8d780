#!/usr/bin/env python
"""What bounds bench.py's `e2e` figure (config 2 through the public API with HOST buffers)?

Times, on one GPU, with pinned host tensors of config 2's shape:
  * the raw H2D copy of q,k,v (805 MB), the raw D2H copy of out (268 MB), and both directions at once;
  * variant "2s": bench.py's original pipeline (two streams, each H2D -> kernel -> D2H per batch element);
  * variant "3s": a copy-in stream, a compute stream and a copy-out stream chained by events over NB buffer sets,
    so the H2D engine never waits for a D2H of the same stream.
Writes gpurun_out/e2e_probe.json. Timing: wall clock around whole steps bracketed by torch.cuda.synchronize()
(the copies are part of the measured quantity, so host time is the right clock here).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v100_b200"))

import torch  # noqa: E402

from flash_attn_v100 import flash_attn_func  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    tot = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = min(best, dt)
        tot += dt
    return {"mean_ms": tot / reps * 1e3, "best_ms": best * 1e3}


def main():
    dev = torch.device("cuda", 0)
    B, S, H, D = 8, 4096, 32, 128
    flops = 4.0 * D * B * H * S * S / 2
    hq, hk, hv = (torch.randn(B, S, H, D, dtype=torch.bfloat16).pin_memory() for _ in range(3))
    hout = torch.empty(B, S, H, D, dtype=torch.bfloat16).pin_memory()
    res = {}

    # ---- raw copies
    gq, gk, gv = (torch.empty(B, S, H, D, dtype=torch.bfloat16, device=dev) for _ in range(3))
    go = torch.randn(B, S, H, D, dtype=torch.bfloat16, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def h2d():
        with torch.cuda.stream(s_in):
            gq.copy_(hq, non_blocking=True); gk.copy_(hk, non_blocking=True); gv.copy_(hv, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s_out):
            hout.copy_(go, non_blocking=True)

    def both():
        h2d(); d2h()

    nbytes_in, nbytes_out = 3 * hq.numel() * 2, hout.numel() * 2
    for name, fn, nb in (("h2d", h2d, nbytes_in), ("d2h", d2h, nbytes_out), ("both", both, nbytes_in + nbytes_out)):
        r = timed(fn)
        r["GBps"] = nb / (r["best_ms"] * 1e-3) / 1e9
        res[name] = r
    del gq, gk, gv, go

    # ---- variant 2s (bench.py round-1 pipeline)
    streams = [torch.cuda.Stream(dev) for _ in range(2)]
    dq = [torch.empty(1, S, H, D, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    dk = [torch.empty_like(dq[0]) for _ in range(2)]
    dv = [torch.empty_like(dq[0]) for _ in range(2)]

    def step_2s():
        for b in range(B):
            st = streams[b % 2]
            with torch.cuda.stream(st):
                dq[b % 2].copy_(hq[b:b + 1], non_blocking=True)
                dk[b % 2].copy_(hk[b:b + 1], non_blocking=True)
                dv[b % 2].copy_(hv[b:b + 1], non_blocking=True)
                o = flash_attn_func(dq[b % 2], dk[b % 2], dv[b % 2], causal=True)
                hout[b:b + 1].copy_(o, non_blocking=True)
        for st in streams:
            st.synchronize()

    r = timed(step_2s)
    r["TFLOPs"] = flops / (r["mean_ms"] * 1e-3) / 1e12
    res["2s"] = r

    # ---- variant 3s: copy-in / compute / copy-out streams over NB buffer sets
    for NB in (2, 3, 4):
        s_c = torch.cuda.Stream(dev)
        bq = [torch.empty(1, S, H, D, dtype=torch.bfloat16, device=dev) for _ in range(NB)]
        bk = [torch.empty_like(bq[0]) for _ in range(NB)]
        bv = [torch.empty_like(bq[0]) for _ in range(NB)]
        loaded = [torch.cuda.Event() for _ in range(NB)]
        consumed = [torch.cuda.Event() for _ in range(NB)]
        computed = [torch.cuda.Event() for _ in range(NB)]

        def step_3s():
            for b in range(B):
                i = b % NB
                with torch.cuda.stream(s_in):
                    if b >= NB:
                        s_in.wait_event(consumed[i])
                    bq[i].copy_(hq[b:b + 1], non_blocking=True)
                    bk[i].copy_(hk[b:b + 1], non_blocking=True)
                    bv[i].copy_(hv[b:b + 1], non_blocking=True)
                    loaded[i].record(s_in)
                with torch.cuda.stream(s_c):
                    s_c.wait_event(loaded[i])
                    o = flash_attn_func(bq[i], bk[i], bv[i], causal=True)
                    consumed[i].record(s_c)
                    computed[i].record(s_c)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(computed[i])
                    o.record_stream(s_out)
                    hout[b:b + 1].copy_(o, non_blocking=True)
            s_in.synchronize(); s_c.synchronize(); s_out.synchronize()

        r = timed(step_3s)
        r["TFLOPs"] = flops / (r["mean_ms"] * 1e-3) / 1e12
        res[f"3s_nb{NB}"] = r

    # parity of the pipelined result against one device-resident call (same inputs)
    ref = flash_attn_func(hq.to(dev), hk.to(dev), hv.to(dev), causal=True)
    res["pipelined_equals_resident"] = bool(torch.equal(ref.cpu(), hout))

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "e2e_probe.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
