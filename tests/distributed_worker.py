"""Two-rank check of every multi-GPU path of the ingest stage, against the oracle (TEST INFRASTRUCTURE).

Run by ``tests/test_distributed_gpu.py`` and, when two GPUs are visible, by ``__graft_entry__.smoke()``:
(a) NCCL gather of ragged dense clouds on the exchange stream, (b) gather fused into the back-projection kernel through a
peer-mapped buffer, (c) variable-length NCCL gather of voxel lists, asynchronous (the next batch's kernel is enqueued before
the wait), (d) the same exchange as peer-store kernels (``RecordExchange``) over several rounds, so inbox slots are reused.
"""

from __future__ import annotations

import os
import socket

import numpy as np


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def worker(rank: int, world: int, port: int, out: dict) -> None:
    import torch
    import torch.distributed as dist

    from oracle import backproject as ob
    from oracle import conventions as conv
    from oracle import voxel as ov
    from tests import cases
    from thor_slam_b200.camera.synthetic import make_depth, make_depth_scene
    from thor_slam_b200.ingest.context import IngestContext
    from thor_slam_b200.ingest.distributed import CloudGather, PeerCloudBuffer, RecordExchange, shard_frame_sets

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
        gat = CloudGather(ctx, rank, world, root=0)
        fused = gat.gather(xyz)
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

        # (c) + (d): voxel lists of three rounds of frame sets; round k of rank r holds sets [k * 4 + r * 2, + 2)
        rounds, per = 3, 2
        ctx.set_voxel_grid(0.05, 10000)
        scene_all = np.stack([make_depth_scene(np.random.default_rng(500 + i), w, h, focal_px=intr.matrix[0, 0]) for i in range(rounds * world * per)])

        def want_round(k: int) -> np.ndarray:
            parts = []
            for r in range(world):
                for j in range(per):
                    i = k * world * per + r * per + j
                    parts.append(ov.voxel_records([(scene_all[i], intr.matrix, m)], 0.05, 10000, set_id=j, tag=r))
            return np.sort(np.concatenate(parts))

        cap = 1 << 16
        rec = [torch.zeros(cap, dtype=torch.int64, device="cuda") for _ in range(2)]
        nrec = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(2)]
        dev_depth = [torch.from_numpy(scene_all[k * world * per + rank * per: k * world * per + (rank + 1) * per].view(np.int16)).cuda().view(torch.uint16)
                     for k in range(rounds)]
        # (c) NCCL, variable length, pipelined: round k + 1's kernel is enqueued before round k's counts are collected, and
        #     round k's exchange runs under it
        gathered = [torch.zeros(world * cap, dtype=torch.int64, device="cuda") if rank == 0 else None for _ in range(rounds)]
        counts_k: list = [None] * rounds
        fences = [0] * rounds

        def send(j: int) -> None:
            _, counts_k[j] = gat.records_send(rec[j % 2], gathered[j], wait=False)
            fences[j] = ctx.exchange_fence()

        for k in range(rounds):
            if k >= 2:
                ctx.exchange_wait(fences[k - 2], on_stream=True)  # rec[k % 2] has left
            ctx.voxel_cloud([(0, dev_depth[k])], rec[k % 2], nrec[k % 2], tag=rank)
            if k >= 1:
                send(k - 1)
            gat.records_begin(nrec[k % 2])
        send(rounds - 1)
        gat.wait()
        torch.cuda.synchronize()
        if rank == 0:
            ok_c = True
            for k in range(rounds):
                got = np.sort(gathered[k][: sum(counts_k[k])].cpu().numpy().view(np.uint64))
                ok_c = ok_c and np.array_equal(got, want_round(k))
            out["records_nccl_ok"] = bool(ok_c)
        dist.barrier()
        # (d) peer-store exchange, two inbox slots, three rounds, nothing waited for on the host until the end; once with the
        #     TMA bulk-copy kernel, once with the 16-byte-store kernel
        for use_tma in (1, 0):
            ctx.set_option(ctx.OPT_PUSH_TMA, use_tma)
            ex = RecordExchange(ctx, rank, world, capacity=world * cap, root=0, slots=2)
            taken = [torch.zeros(world * cap, dtype=torch.int64, device="cuda") for _ in range(rounds)] if rank == 0 else None
            status = [torch.zeros(2, dtype=torch.int32, device="cuda") for _ in range(rounds)] if rank == 0 else None
            own = []
            for k in range(rounds):
                if k >= 2:
                    ctx.exchange_wait(fences[k - 2], on_stream=True)  # rec[k % 2] was pushed two rounds ago
                ctx.voxel_cloud([(0, dev_depth[k])], rec[k % 2], nrec[k % 2], tag=rank)
                if rank == 0:
                    own.append((rec[k % 2].clone(), nrec[k % 2].clone()))  # the root's own list stays local
                fences[k] = ex.push(rec[k % 2], nrec[k % 2])
                if rank == 0:
                    ex.take(taken[k], status[k])
            ex.wait()
            torch.cuda.synchronize()
            if rank == 0:
                ok_d = True
                for k in range(rounds):
                    n, err = (int(x) for x in status[k].cpu().numpy())
                    got = taken[k][:n].cpu().numpy().view(np.uint64)
                    mine = own[k][0][: int(own[k][1].item())].cpu().numpy().view(np.uint64)
                    got = np.sort(np.concatenate([got[got != 0], mine]))  # zero records pad odd lists to 16 bytes
                    ok_d = ok_d and err == 0 and np.array_equal(got, want_round(k))
                out["records_push_tma_ok" if use_tma else "records_push_ok"] = bool(ok_d)
            ex.close()
        ctx.set_option(ctx.OPT_PUSH_TMA, 1)
        ctx.close()
    finally:
        dist.destroy_process_group()


def run_two_ranks(timeout_s: float = 300.0) -> dict:
    """Spawn the two-rank check; raises if a rank fails or hangs past ``timeout_s``."""
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    out = mgr.dict()
    ctx = mp.spawn(worker, args=(2, free_port(), out), nprocs=2, join=False)
    import time

    t0 = time.time()
    while not ctx.join(timeout=5.0):
        if time.time() - t0 > timeout_s:
            for p in ctx.processes:
                p.kill()
            raise TimeoutError("two-rank check did not finish")
    return dict(out)
