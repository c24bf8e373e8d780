#!/usr/bin/env python
"""Raw host <-> device copy rates with N ranks copying at once (is bench.py's `e2e` at N = 8 at the
host's ceiling?). Launch under torch.distributed.run; every rank pins config 2's q,k,v (805 MB) and out (268 MB), and after
a barrier times H2D alone, D2H alone and both directions at once (CUDA events on the rank's own streams, MAX over ranks
through a gloo all-reduce). Rank 0 prints one JSON line with the per-rank and the aggregate GB/s."""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo", init_method="env://")
    B, S, H, D = 8, 4096, 32, 128
    host_in = [torch.empty(B, S, H, D, dtype=torch.bfloat16).pin_memory() for _ in range(3)]
    host_out = torch.empty(B, S, H, D, dtype=torch.bfloat16).pin_memory()
    dev_in = [torch.empty(B, S, H, D, dtype=torch.bfloat16, device="cuda") for _ in range(3)]
    dev_out = torch.empty(B, S, H, D, dtype=torch.bfloat16, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    in_bytes, out_bytes = 3 * host_in[0].numel() * 2, host_out.numel() * 2

    def run(do_in, do_out, reps=5):
        best = 1e9
        for _ in range(reps + 1):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s_in.wait_event(e0)
            s_out.wait_event(e0)
            if do_in:
                with torch.cuda.stream(s_in):
                    for h, d in zip(host_in, dev_in):
                        d.copy_(h, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s_out):
                    host_out.copy_(dev_out, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s_in)
            torch.cuda.current_stream().wait_stream(s_out)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        t = torch.tensor([best], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    res = {"n_ranks": world}
    for name, di, do in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        ms = run(di, do)
        nbytes = (in_bytes if di else 0) + (out_bytes if do else 0)
        res[name] = {"ms_max_over_ranks": round(ms, 3), "GBps_per_rank": round(nbytes / ms / 1e6, 1),
                     "GBps_aggregate": round(world * nbytes / ms / 1e6, 1)}
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
