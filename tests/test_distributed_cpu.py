"""N > 1 host logic on CPU: world_size-2 ``gloo`` process group (sharding, count exchange, gather layout)."""

from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from thor_slam_b200.ingest.distributed import CloudGather, exchange_counts, gather_layout, shard_frame_sets


def test_sharding_is_a_partition():
    for n in (0, 1, 7, 64):
        for world in (1, 2, 4, 8):
            shards = [shard_frame_sets(n, r, world) for r in range(world)]
            assert sorted(i for s in shards for i in s) == list(range(n))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    with pytest.raises(ValueError):
        shard_frame_sets(4, 2, 2)
    assert gather_layout([12, 0, 24]) == [0, 12, 12]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_sets: int, out: dict) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard_frame_sets(n_sets, rank, world)
        # every frame set yields a "cloud" whose length depends on the frame-set index (ragged, like compacted clouds)
        clouds = [torch.full((10 + 3 * i, 3), float(i)) for i in mine]
        local = torch.cat(clouds) if clouds else torch.zeros((0, 3))
        sizes = exchange_counts(local.numel() * 4)
        assert sizes == [sum((10 + 3 * i) * 12 for i in shard_frame_sets(n_sets, r, world)) for r in range(world)]
        fused = CloudGather(None, rank, world, root=0).gather(local)
        if rank == 0:
            fused = fused.view(-1, 3)
            want = torch.cat([torch.full((10 + 3 * i, 3), float(i)) for r in range(world) for i in shard_frame_sets(n_sets, r, world)])
            out["ok"] = bool(torch.equal(fused, want))
            out["n"] = int(fused.shape[0])
        else:
            assert fused is None
        # variable-length record lists (the gloo face of gather_records): capacity 64, rank r holds 5 + 9 r records
        n = 5 + 9 * rank
        rec = torch.zeros(64, dtype=torch.int64)
        rec[:n] = torch.arange(n) + 1000 * rank
        fused, counts = CloudGather(None, rank, world, root=0).gather_records(rec, torch.tensor([n], dtype=torch.int32))
        assert counts == [5 + 9 * r for r in range(world)]
        if rank == 0:
            out["rec_ok"] = bool(torch.equal(fused, torch.cat([torch.arange(5 + 9 * r) + 1000 * r for r in range(world)])))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather():
    world, n_sets = 2, 7
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_sets, out), nprocs=world, join=True)
    assert out["ok"] and out["n"] == sum(10 + 3 * i for i in range(n_sets))
    assert out["rec_ok"]
