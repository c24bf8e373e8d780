#!/usr/bin/env python3
"""BASELINE config 4 timing: paged KV-cache decode (Sq=1, Sk=8192, H=32, D=128, rotary, page=256).
Reports achieved HBM GB/s = attended K+V bytes / time for B in {1, 64, 256}, Hk in {8, 32}."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v100_b200"))
from flash_attn_v100 import flash_attn_with_kvcache  # noqa: E402


def rotary_tables(seqlen, rot, dtype):
    inv = 1.0 / (10000 ** (torch.arange(0, rot, 2, dtype=torch.float32) / rot))
    ang = torch.outer(torch.arange(seqlen, dtype=torch.float32), inv)
    return ang.cos().to(dtype).cuda(), ang.sin().to(dtype).cuda()


def run(B, Hk, H=32, D=128, Sk=8192, page=256, iters=20, n_caches=4):
    dt = torch.bfloat16
    n_pages = B * (Sk // page)
    caches = []
    for i in range(n_caches):  # rotate caches so the 126 MB L2 cannot hold the working set
        kc = torch.randn(n_pages, page, Hk, D, device="cuda", dtype=dt)
        vc = torch.randn(n_pages, page, Hk, D, device="cuda", dtype=dt)
        caches.append((kc, vc))
        if kc.numel() * 2 * 2 * (i + 1) > 3 * 2 ** 30:
            break
    bt = torch.randperm(n_pages, generator=torch.Generator().manual_seed(0)).view(B, -1).int().cuda()
    lens = torch.full((B,), Sk - 1, dtype=torch.int32, device="cuda")
    q = torch.randn(B, 1, H, D, device="cuda", dtype=dt)
    kn = torch.randn(B, 1, Hk, D, device="cuda", dtype=dt)
    vn = torch.randn(B, 1, Hk, D, device="cuda", dtype=dt)
    cos, sin = rotary_tables(Sk, D, dt)

    def step(i):
        kc, vc = caches[i % len(caches)]
        return flash_attn_with_kvcache(q, kc, vc, kn, vn, rotary_cos=cos, rotary_sin=sin, cache_seqlens=lens,
                                       block_table=bt, causal=True, rotary_interleaved=False)

    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nbytes = 2 * B * Sk * Hk * D * 2
    rec = {"B": B, "Hk": Hk, "us": ms * 1e3, "GBps": nbytes / ms / 1e6, "frac_of_6464.9": nbytes / ms / 1e6 / 6464.9,
           "caches_rotated": len(caches)}
    # the same steps replayed from CUDA graphs (one per cache buffer), as a serving loop would issue them
    graphs = []
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(len(caches)):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                step(i)
            graphs.append(g)
    torch.cuda.current_stream().wait_stream(side)
    for g in graphs:
        g.replay()
    torch.cuda.synchronize()
    e0.record()
    for i in range(iters):
        graphs[i % len(graphs)].replay()
    e1.record()
    torch.cuda.synchronize()
    gms = e0.elapsed_time(e1) / iters
    rec.update(graph_us=gms * 1e3, graph_GBps=nbytes / gms / 1e6)
    return rec


if __name__ == "__main__":
    res = []
    for B, Hk in [(64, 8), (64, 32), (1, 8), (256, 8), (8, 8)]:
        r = run(B, Hk)
        print(json.dumps(r), flush=True)
        res.append(r)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "decode_bench.json"), "w"), indent=1)
