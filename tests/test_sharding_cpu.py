"""CPU, 2 gloo ranks: the N>1 host logic (batch sharding + max-over-ranks timing) that
bench.py uses on N GPUs.  The per-rank compute here is the CPU oracle, standing in for the
CUDA path, so the test checks the plumbing: shards are disjoint, cover the batch, and the
sharded result equals the single-process one."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from golf_b200.sharding import shard_range

    for n in (0, 1, 7, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import synthetic_controls
    from golf_b200.sharding import max_over_ranks, shard_range, sum_over_ranks
    from oracle import golf_oracle as O

    torch.set_num_threads(1)
    O.set_num_threads(1)
    B, Tn, H, M = 5, 2400, 240, 12
    gain, a = synthetic_controls(B, Tn // H + 1, M, seed=1)
    ex = torch.randn(B, Tn, generator=torch.Generator().manual_seed(2))
    lo, hi = shard_range(B, rank, world)
    y = O.lpc_ss_fused(ex[lo:hi], gain[lo:hi], a[lo:hi], H)
    full = torch.zeros(B, y.shape[1])
    full[lo:hi] = y
    dist.all_reduce(full)  # disjoint shards: the sum is the concatenation
    slowest = max_over_ranks(float(rank + 1))
    total = sum_over_ranks(float(hi - lo))
    if rank == 0:
        ref = O.lpc_ss_fused(ex, gain, a, H)
        out.put((bool(torch.equal(full, ref)), slowest, total))
    dist.destroy_process_group()


def test_two_rank_batch_sharding_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    same, slowest, total = out.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert same and slowest == 2.0 and total == 5.0
