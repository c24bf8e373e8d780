"""Host-side multi-GPU logic on CPU: sharding helpers and the gloo (world_size 2) reduction path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sharding


def test_shard_range_covers_everything_once():
    for total in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                a, b = sharding.shard_range(total, world, r)
                got += list(range(a, b))
            assert got == list(range(total))
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def test_batch_then_gqa_group_sharding():
    assert sharding.shard_batch_or_heads(64, 32, 8, 8, 3) == ((24, 32), (0, 32))  # BASELINE config 5
    # batch 1, 8 ranks, 8 kv heads of 4 query heads each: one GQA group per rank
    seen = []
    for r in range(8):
        (b0, b1), (h0, h1) = sharding.shard_batch_or_heads(1, 32, 8, 8, r)
        assert (b0, b1) == (0, 1) and (h1 - h0) == 4 and h0 % 4 == 0
        seen += list(range(h0, h1))
    assert seen == list(range(32))
    with pytest.raises(ValueError):
        sharding.shard_batch_or_heads(1, 8, 2, 8, 0)


def test_balanced_varlen_shards():
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(1, 2049, (64,), generator=g).tolist()  # BASELINE config 3 lengths
    shards = sharding.balanced_varlen_shards(lens, 8)
    assert sorted(i for s in shards for i in s) == list(range(64))
    loads = [sum(lens[i] ** 2 for i in s) for s in shards]
    assert max(loads) / (sum(loads) / 8) < 1.05


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b0, b1 = sharding.shard_range(64, world, rank)
        elapsed = 10.0 + rank  # pretend rank 1 was slower
        dist.barrier()
        mx = sharding.max_over_ranks(elapsed, dist)
        sums = sharding.gather_checksums(float(b1 - b0), dist)
        if rank == 0:
            out.put((mx, sums))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_max_reduce_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    mx, sums = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert mx == 11.0 and sums == [32.0, 32.0]


def test_single_process_paths():
    assert sharding.max_over_ranks(3.5) == 3.5
    assert sharding.gather_checksums(2.0) == [2.0]


def test_bench_reference_arm_prints_one_json_line(tmp_path):
    """`bench.py --impl reference` (CPU arm of the contract) on a shrunk workload: one JSON line, the
    reference's own cpu_attention when oracle/_ref is built, else the C port."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--workload', 'c2']\n"
        "import bench\n"
        "bench.WORKLOADS['c2'] = dict(bench.WORKLOADS['c2'], seqlen=256, batch=1, heads=2, heads_k=2)\n"
        "bench.main()\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "TFLOP/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0
