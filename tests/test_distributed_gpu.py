"""N = 2 on real GPUs: ragged cloud gather over NCCL inside the library and the gather fused into the
back-projection kernel through a peer-mapped buffer.  Needs two visible GPUs (skipped otherwise); bench.py
checks the same paths at N = 2, 4, 8 (``matches_local``)."""

from __future__ import annotations

import os
import socket

import numpy as np
import pytest


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out: dict) -> None:
    import torch
    import torch.distributed as dist

    from oracle import backproject as ob
    from oracle import conventions as conv
    from tests import cases
    from thor_slam_b200.camera.synthetic import make_depth
    from thor_slam_b200.ingest import formats as F
    from thor_slam_b200.ingest.context import IngestContext, StreamSpec
    from thor_slam_b200.ingest.distributed import CloudGather, PeerCloudBuffer, shard_frame_sets

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ctx = IngestContext(rank)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        w, h, n_sets = 320, 200, 5
        s, _ = cases.stereo_maps(w, h, seed=3)
        intr = s.get_intrinsics()[0]
        m = conv.body_T_camera(cases.random_pose(np.random.default_rng(4)), s.get_extrinsics()[0].to_4x4_matrix(), "rdf")
        ctx.upload_projection(0, intr.matrix, m, (w, h))
        depth_all = np.stack([make_depth(np.random.default_rng(100 + i), w, h) for i in range(n_sets)])
        mine = shard_frame_sets(n_sets, rank, world)  # 3 + 2 frame sets: ragged
        depth = torch.from_numpy(depth_all[mine].view(np.int16)).cuda().view(torch.uint16)
        xyz = torch.zeros((len(mine), h, w, 3), dtype=torch.float32, device="cuda")
        mask = torch.zeros((len(mine), h, w), dtype=torch.uint8, device="cuda")
        ctx.backproject(0, depth, xyz, mask)
        # (a) NCCL gather of ragged dense clouds
        fused = CloudGather(ctx, rank, world, root=0).gather(xyz)
        torch.cuda.synchronize()
        if rank == 0:
            order = [i for r in range(world) for i in shard_frame_sets(n_sets, r, world)]
            got = fused.view(n_sets, h, w, 3).cpu().numpy()
            ok = True
            for k, i in enumerate(order):
                pts, _, _ = ob.backproject(depth_all[i], intr.matrix, m)
                ok = ok and ob.points_close(got[k], pts)[0]
            out["gather_ok"] = bool(ok)
        # (b) gather fused into the kernel: fixed-size slices of a buffer that lives on rank 0
        per_rank = (3, h, w, 3)
        peer = PeerCloudBuffer(ctx, rank, world, per_rank, root=0)
        dst = peer.slice_for(rank).slice0(0, len(mine))
        ctx.backproject(0, depth, dst, mask)
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            whole = peer.as_tensor().cpu().numpy()
            ok = True
            for r in range(world):
                for k, i in enumerate(shard_frame_sets(n_sets, r, world)):
                    pts, _, _ = ob.backproject(depth_all[i], intr.matrix, m)
                    ok = ok and ob.points_close(whole[r, k], pts)[0]
            out["peer_ok"] = bool(ok)
        dist.barrier()
        peer.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_gather_and_peer_store():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out.get("gather_ok") and out.get("peer_ok")
