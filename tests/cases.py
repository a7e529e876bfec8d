"""Parity cases shared by the CPU-emulation tests and the ``-m gpu`` tests.

Every case drives the C ABI (through ``IngestContext``) on seeded inputs and compares with the
oracle (``oracle/``) and/or the committed golden fixtures.  Bars: bit-exact for u8 / masks / counts;
``|p - p_ref|_inf <= 1e-5 * max(|p_ref|_inf, 1e-3)`` for 3-D points (north_star tolerance).
"""

from __future__ import annotations

import numpy as np

from oracle import backproject as ob
from oracle import conventions as conv
from oracle import convert as oc
from oracle import rectify as orc
from oracle import voxel as ov
from thor_slam_b200.camera.synthetic import SyntheticCameraConfig, SyntheticCameraSource, make_depth, make_depth_scene, make_image
from thor_slam_b200.ingest import formats as F
from thor_slam_b200.ingest.context import StreamSpec

POINT_RTOL = 1e-5
POINT_FLOOR = 1e-3

CONVERSIONS = [("bgr8", "rgb8"), ("bgr8", "mono8"), ("nv12", "mono8"), ("nv12", "rgb8"), ("nv12", "bgr8"), ("mono8", "mono8")]
RECTIFY_CONVERSIONS = [("mono8", "mono8"), ("nv12", "mono8"), ("bgr8", "rgb8"), ("bgr8", "mono8"), ("nv12", "rgb8")]


def oracle_convert(img: np.ndarray, s: str, d: str) -> np.ndarray:
    if (s, d) == ("nv12", "bgr8"):
        return oc.nv12_to_bgr_cv(img)
    if (s, d) == ("mono8", "mono8"):
        return img
    return oc.convert(img, s, d)


def make_batch(rng: np.random.Generator, fmt: str, w: int, h: int, n: int) -> np.ndarray:
    if n == 0:
        return np.zeros((0, *F.frame_shape(F.fmt(fmt), w, h)), dtype=np.uint8)
    return np.stack([make_image(rng, fmt, w, h) for _ in range(n)])


def check_convert(be, s: str, d: str, w: int, h: int, n: int = 2, seed: int = 0) -> None:
    rng = np.random.default_rng(seed)
    src = make_batch(rng, s, w, h, n)
    if n and s == "nv12":  # full-range random chroma AND luma, so saturation paths are hit
        src[0] = rng.integers(0, 256, size=src[0].shape, dtype=np.uint8)
    dst = be.zeros((n, *F.frame_shape(F.fmt(d), w, h)), np.uint8)
    be.ctx.convert(be.dev(src), dst, s, d, w, h)
    got = be.host(dst)
    for i in range(n):
        want = oracle_convert(src[i], s, d)
        assert np.array_equal(got[i], want), f"{s}->{d} {w}x{h} frame {i}: {(got[i] != want).sum()} bytes differ"


def stereo_maps(w: int, h: int, seed: int = 3, distortion: str = "rational14"):
    """(source, [(mapx, mapy) left, right]) from a synthetic stereo calibration via the oracle (cv2)."""
    s = SyntheticCameraSource(SyntheticCameraConfig(name="oak0", resolution=(w, h), pool=1, seed=seed, distortion=distortion))
    (il, ir), (el, er) = s.get_intrinsics(), s.get_extrinsics()
    r1, r2, p1, p2 = orc.stereo_rectify_cv(il.matrix, il.coeffs, ir.matrix, ir.coeffs, (w, h), el.to_4x4_matrix(), er.to_4x4_matrix())
    maps = [orc.undistort_rectify_map_cv(il.matrix, il.coeffs, r1, p1, (w, h)),
            orc.undistort_rectify_map_cv(ir.matrix, ir.coeffs, r2, p2, (w, h))]
    return s, maps


def edge_maps(dst_w: int, dst_h: int) -> tuple[np.ndarray, np.ndarray]:
    """An affine map that leaves the source on every side (border taps, fully-outside pixels)."""
    yy, xx = np.mgrid[0:dst_h, 0:dst_w].astype(np.float32)
    return (xx * 1.13 - 13.3 + 0.05 * yy).astype(np.float32), (yy * 1.21 - 14.7 - 0.03 * xx).astype(np.float32)


def shear_maps(dst_w: int, dst_h: int, shear: float = 0.16) -> tuple[np.ndarray, np.ndarray]:
    """A map whose source row changes every ``1 / shear`` output pixels: more than 32 exception pairs per (tile, warp) of the
    pair-window kernel, so its overflow list is exercised (a fisheye camera does the same near the image corners)."""
    yy, xx = np.mgrid[0:dst_h, 0:dst_w].astype(np.float32)
    return (xx + 3.3).astype(np.float32), (yy + shear * xx + 0.4).astype(np.float32)


def check_rectify(be, cam: int, mapx, mapy, s: str, d: str, src_w: int, src_h: int, n: int = 3, seed: int = 1,
                  expect_variant: int | None = None, expect_exceptions: bool = False, expect_overflow: bool = False) -> None:
    rng = np.random.default_rng(seed)
    dst_h, dst_w = mapx.shape
    be.ctx.upload_rectify_map(cam, mapx, mapy, (src_w, src_h))
    src = make_batch(rng, s, src_w, src_h, n)
    if s != "mono8" and n:
        src[0] = rng.integers(0, 256, size=src[0].shape, dtype=np.uint8)
    dst = be.zeros((n, *F.frame_shape(F.fmt(d), dst_w, dst_h)), np.uint8)
    wants = [orc.remap_cv(np.ascontiguousarray(oracle_convert(src[i], s, d)), mapx, mapy) for i in range(n)]
    # every kernel variant must give the same bytes: TMA-pipelined (tile height 32 and 16), thread-staged, generic
    # (the TMA kernel also with 1 and 2 frames of the batch per LUT fetch)
    mono = d == "mono8" and s in ("mono8", "nv12")
    variants = ([(4, 32, 8), (4, 16, 2), (4, 32, 1), (4, 24, 3), (3, 32, 8), (3, 16, 2), (3, 24, 3), (3, 32, 1), (2, 32, 8), (1, 32, 8)] if mono
                else [(4, 32, 8), (1, 32, 8)])
    stages = {8: 4, 2: 3, 3: 2, 1: 6}
    try:
        for variant, th, fpu in variants:
            be.ctx.set_option(be.ctx.OPT_MONO_VARIANT, variant)
            be.ctx.set_option(be.ctx.OPT_TMA_TILE_H, th)
            be.ctx.set_option(be.ctx.OPT_FRAMES_PER_UNIT, fpu)
            be.ctx.set_option(be.ctx.OPT_STAGES, stages[fpu])
            be.ctx.set_option(be.ctx.OPT_LUT_PREFETCH, 0 if fpu == 3 else 1)  # defaults (16 frames, 2 stages, no prefetch) run everywhere else
            if mono and expect_variant is not None and variant == 4:
                plan = be.ctx.rectify_plan(cam)
                assert plan["variant"] == expect_variant, f"slot {cam} would run kernel variant {plan}"
                if expect_exceptions:
                    assert plan["exceptions_per_warp"] > 0, "this map was chosen to exercise the exception path"
                if expect_overflow and th == 32:
                    assert plan["exceptions_per_warp"] == 32 and plan["overflow_pixels"] > 0, f"this map was chosen to overflow the exception lists: {plan}"
            if (s, d) == ("bgr8", "rgb8") and expect_variant == 4 and variant == 4:
                assert be.ctx.rectify_plan(cam)["colour_variant"] == 5, "BGR8 -> RGB8 must run the 3-channel window kernel here"
            dst = be.zeros((n, *F.frame_shape(F.fmt(d), dst_w, dst_h)), np.uint8)
            be.ctx.rectify(cam, be.dev(src), dst, s, d)
            got = be.host(dst)
            for i in range(n):
                assert np.array_equal(got[i], wants[i]), (
                    f"rectify {s}->{d} variant={variant} th={th} frame {i}: {(got[i] != wants[i]).sum()} bytes differ")
        if mono:
            # the pair-window kernel's other layout (two pixels per window: what maps fall back to when four per window overflow)
            be.ctx.set_option(be.ctx.OPT_MONO_VARIANT, 4)
            be.ctx.set_option(be.ctx.OPT_TMA_TILE_H, 32)
            be.ctx.set_option(be.ctx.OPT_FRAMES_PER_UNIT, 0)
            be.ctx.set_option(be.ctx.OPT_RECTIFY_QUAD, 0)
            be.ctx.upload_rectify_map(cam, mapx, mapy, (src_w, src_h))
            assert be.ctx.rectify_plan(cam)["pixels_per_window"] in (0, 2)
            dst = be.zeros((n, *F.frame_shape(F.fmt(d), dst_w, dst_h)), np.uint8)
            be.ctx.rectify(cam, be.dev(src), dst, s, d)
            got = be.host(dst)
            for i in range(n):
                assert np.array_equal(got[i], wants[i]), f"rectify {s}->{d} pair layout frame {i}: {(got[i] != wants[i]).sum()} bytes differ"
    finally:
        be.ctx.set_option(be.ctx.OPT_RECTIFY_QUAD, 1)
        if mono:
            be.ctx.upload_rectify_map(cam, mapx, mapy, (src_w, src_h))  # the slot is left in the default layout
        be.ctx.set_option(be.ctx.OPT_MONO_VARIANT, 4)
        be.ctx.set_option(be.ctx.OPT_TMA_TILE_H, 32)
        be.ctx.set_option(be.ctx.OPT_FRAMES_PER_UNIT, 16)
        be.ctx.set_option(be.ctx.OPT_FRAMES_PER_UNIT, 0)
        be.ctx.set_option(be.ctx.OPT_STAGES, 6)
        be.ctx.set_option(be.ctx.OPT_LUT_PREFETCH, 0)
    mask = be.zeros((dst_h, dst_w), np.uint8)
    be.ctx.get_valid_mask(cam, mask)
    assert np.array_equal(be.host(mask), orc.valid_mask(mapx, mapy, (src_w, src_h)))


def random_pose(rng: np.random.Generator) -> np.ndarray:
    from scipy.spatial.transform import Rotation

    t = np.eye(4)
    t[:3, :3] = Rotation.from_rotvec(rng.uniform(-1.5, 1.5, 3)).as_matrix()
    t[:3, 3] = rng.uniform(-0.5, 0.5, 3)
    return t


def check_backproject(be, cam: int, w: int, h: int, n: int = 2, seed: int = 2, rig_frame: str = "rdf", depth=None) -> None:
    rng = np.random.default_rng(seed)
    s = SyntheticCameraSource(SyntheticCameraConfig(name="oak0", resolution=(w, h), pool=1, seed=seed))
    intr, ext = s.get_intrinsics()[0], s.get_extrinsics()[0]
    m = conv.body_T_camera(random_pose(rng), ext.to_4x4_matrix(), rig_frame)
    be.ctx.upload_projection(cam, intr.matrix, m, (w, h))
    if depth is None:
        depth = np.stack([make_depth(rng, w, h) for _ in range(n)]) if n else np.zeros((0, h, w), np.uint16)
    n = depth.shape[0]
    xyz = be.zeros((n, h, w, 3), np.float32)
    mask = be.zeros((n, h, w), np.uint8)
    count = be.dev(np.full(max(n, 1), 0xDEADBEEF, dtype=np.uint32))  # must be overwritten, not accumulated
    be.ctx.backproject(cam, be.dev(depth), xyz, mask, count)
    gx, gm, gc = be.host(xyz), be.host(mask), be.host(count)
    for i in range(n):
        pts, msk, cnt = ob.backproject(depth[i], intr.matrix, m)
        ok, worst = ob.points_close(gx[i], pts, POINT_RTOL, POINT_FLOOR)
        assert ok, f"points off by {worst:.3e} (> {POINT_RTOL})"
        assert np.array_equal(gm[i], msk)
        assert int(gc[i]) == cnt
        assert not np.any(gx[i][msk == 0]), "invalid pixels must be written as (0,0,0)"


def check_depth_stats(be, w: int, h: int, n: int = 3, seed: int = 8) -> None:
    """``ti_depth_stats`` against ``oracle.backproject.depth_stats`` (examples/rgbd_stream.py:270-276): exact integers."""
    rng = np.random.default_rng(seed)
    depth = np.stack([make_depth(rng, w, h) for _ in range(n)])
    depth[0] = 0                       # nothing valid
    if n > 1:
        depth[1, h // 2:, :] = 65535   # saturated half
    stats = be.dev(np.full((n, 6), 0xDEADBEEF, dtype=np.uint32))
    be.ctx.depth_stats(be.dev(depth), stats)
    got = be.host(stats)
    for i in range(n):
        want = ob.depth_stats(depth[i])
        total = int(got[i, 4]) | (int(got[i, 5]) << 32)
        assert int(got[i, 0]) == want["count"], (i, got[i], want)
        if want["count"]:
            assert int(got[i, 1]) == want["min"] and int(got[i, 2]) == want["max"]
            assert total == int(depth[i].astype(np.uint64).sum()) and abs(total / want["count"] - want["mean"]) < 1e-6
        else:
            assert int(got[i, 1]) == 0xFFFFFFFF and total == 0


def check_backproject_colour(be, cam: int, w: int, h: int, rw: int, rh: int, n: int = 2, seed: int = 6, on_half_pixels: bool = False) -> None:
    """``ti_backproject_colour``: the cloud / mask / count of ``ti_backproject`` AND the colours of ``ti_register_colour`` from one
    pass over the depth image.  ``on_half_pixels``: an RGB camera identical to the depth camera but for half a pixel of
    principal point - every projection lands exactly on x.5, the worst case for the reciprocal fast path's guard band."""
    rng = np.random.default_rng(seed)
    s = SyntheticCameraSource(SyntheticCameraConfig(name="oak0", resolution=(w, h), enable_rgbd=True, rgb_resolution=(rw, rh), depth_resolution=(w, h),
                                                    pool=1, seed=seed))
    ri, di = s.get_rgbd_intrinsics()
    re, de = s.get_rgbd_extrinsics()
    k_rgb, rgb_T_depth = ri.matrix, np.linalg.inv(re.to_4x4_matrix()) @ de.to_4x4_matrix()
    if on_half_pixels:
        assert (rw, rh) == (w, h)
        k_rgb = di.matrix.copy()
        k_rgb[0, 2] += 0.5
        k_rgb[1, 2] += 0.5
        rgb_T_depth = np.eye(4)
    m = conv.body_T_camera(random_pose(rng), de.to_4x4_matrix(), "rdf")
    be.ctx.upload_projection(cam, di.matrix, m, (w, h))
    be.ctx.upload_registration(cam, di.matrix, (w, h), k_rgb, (rw, rh), rgb_T_depth)
    depth = np.stack([make_depth_scene(rng, w, h) if i % 2 else make_depth(rng, w, h) for i in range(n)])
    rgb = rng.integers(0, 256, size=(n, rh, rw, 3), dtype=np.uint8)
    xyz, mask = be.zeros((n, h, w, 3), np.float32), be.zeros((n, h, w), np.uint8)
    count = be.dev(np.full(n, 0xDEADBEEF, dtype=np.uint32))
    colour = be.dev(np.full((n, h, w, 3), 0x77, dtype=np.uint8))
    be.ctx.backproject_colour(cam, be.dev(depth), be.dev(rgb), xyz, colour, mask, count)
    gx, gm, gc, gcol = be.host(xyz), be.host(mask), be.host(count), be.host(colour)
    for i in range(n):
        pts, msk, cnt = ob.backproject(depth[i], di.matrix, m)
        ok, worst = ob.points_close(gx[i], pts, POINT_RTOL, POINT_FLOOR)
        assert ok, f"points off by {worst:.3e}"
        assert np.array_equal(gm[i], msk) and int(gc[i]) == cnt
        want = ob.register_colour(depth[i], di.matrix, rgb_T_depth, k_rgb, rgb[i])
        assert np.array_equal(gcol[i], want), f"frame {i}: {(gcol[i] != want).any(axis=-1).sum()} colours differ"
    # and the stand-alone kernel agrees with the fused one
    alone = be.dev(np.zeros((n, h, w, 3), np.uint8))
    be.ctx.register_colour(cam, be.dev(depth), be.dev(rgb), alone)
    assert np.array_equal(be.host(alone), gcol)


def check_voxel(be, cam0: int, sizes: list[tuple[int, int]], n: int = 2, seed: int = 4, voxel: float = 0.05, max_depth_mm: int = 10000,
                scene: str = "room", set_base: int = 0, tag: int = 0, capacity: int | None = None, depth=None) -> int:
    """``ti_voxel_cloud`` over ``len(sizes)`` cameras x ``n`` frame sets against ``np.unique`` of the oracle's keys: the SET of
    records, the total and the per-set counts must be exact; ``ti_voxel_points`` must give the voxel centres bit for bit.
    Returns the number of records."""
    rng = np.random.default_rng(seed)
    cams = []
    for i, (w, h) in enumerate(sizes):
        s = SyntheticCameraSource(SyntheticCameraConfig(name=f"oak{i}", resolution=(w, h), pool=1, seed=seed + i))
        intr, ext = s.get_intrinsics()[0], s.get_extrinsics()[0]
        m = conv.body_T_camera(random_pose(rng), ext.to_4x4_matrix(), "rdf")
        be.ctx.upload_projection(cam0 + i, intr.matrix, m, (w, h))
        if depth is not None:
            d = depth[i]
        elif scene == "room":
            d = np.stack([make_depth_scene(rng, w, h, focal_px=intr.matrix[0, 0]) for _ in range(n)]) if n else np.zeros((0, h, w), np.uint16)
        else:
            d = np.stack([make_depth(rng, w, h) for _ in range(n)]) if n else np.zeros((0, h, w), np.uint16)
        cams.append((d, intr.matrix, m))
    n = cams[0][0].shape[0]
    want_sets = [ov.voxel_records([(d[b], k, m) for d, k, m in cams], voxel, max_depth_mm, set_base + b, tag) for b in range(n)]
    want = np.concatenate(want_sets) if want_sets else np.zeros(0, np.uint64)
    cap = capacity if capacity is not None else max(len(want) + 7, 1)
    be.ctx.set_voxel_grid(voxel, max_depth_mm)
    records = be.dev(np.full(cap, 0x5555555555555555, dtype=np.int64))
    n_rec = be.dev(np.full(1, 0xDEADBEEF, dtype=np.uint32))  # overwritten, not accumulated
    counts = be.dev(np.full(max(n, 1), 0xDEADBEEF, dtype=np.uint32))
    for _ in range(2):  # a second launch reuses the hash set: entries of the first must read as free
        be.ctx.voxel_cloud([(cam0 + i, be.dev(cams[i][0])) for i in range(len(cams))], records, n_rec, counts, set_base=set_base, tag=tag)
    got_n = int(be.host(n_rec)[0])
    got_counts = be.host(counts)
    got = be.host(records).view(np.uint64)
    if capacity is None:
        assert got_n == len(want), f"{got_n} records, oracle has {len(want)} distinct voxels"
        for b in range(n):
            assert int(got_counts[b]) == len(want_sets[b]), f"set {b}: {int(got_counts[b])} vs {len(want_sets[b])}"
        assert np.array_equal(np.sort(got[:got_n]), np.sort(want)), "record sets differ"
        assert np.all(got[got_n:] == 0x5555555555555555), "wrote past the end of the list"
        xyz = be.zeros((cap, 3), np.float32)
        be.ctx.voxel_points(records, n_rec, xyz)
        assert np.array_equal(be.host(xyz)[:got_n], ov.record_points(got[:got_n], voxel))
        assert not be.host(xyz)[got_n:].any()
    else:  # truncated list: the count says so (a lower bound of the distinct voxels), everything written is a genuine, distinct record
        assert cap < got_n <= len(want)
        assert len(np.unique(got)) == cap and np.isin(got, want).all()
    return got_n


def check_host_pipeline(be, cam: int, w: int, h: int, batches: list[int], chunk: int, pinned: bool, submit: bool) -> None:
    """Host buffers in, host buffers out (``ti_ingest_host`` / ``ti_ingest_host_submit`` + ``_wait``): one rectify and
    one back-projection stream per batch; with ``submit`` every batch is enqueued before the first wait."""
    rng = np.random.default_rng(10 + cam)
    s, maps = stereo_maps(w, h, seed=10)
    be.ctx.upload_rectify_map(cam, *maps[0], (w, h))
    intr = s.get_intrinsics()[0]
    m = conv.body_T_camera(None, s.get_extrinsics()[0].to_4x4_matrix(), "rdf")
    be.ctx.upload_projection(cam, intr.matrix, m, (w, h))

    def host(a: np.ndarray):
        if not pinned:
            return a
        import torch

        return torch.from_numpy(a).pin_memory()

    def view(x) -> np.ndarray:
        return x if isinstance(x, np.ndarray) else x.numpy()

    jobs = []
    for n in batches:
        left = host(make_batch(rng, "mono8", w, h, n))
        depth_np = np.stack([make_depth(rng, w, h) for _ in range(n)]) if n else np.zeros((0, h, w), np.uint16)
        depth = host(depth_np.view(np.int16))
        out = dict(o_l=host(np.zeros((n, h, w), np.uint8)), xyz=host(np.zeros((n, h, w, 3), np.float32)),
                   mask=host(np.zeros((n, h, w), np.uint8)), count=host(np.zeros((n,), np.int32)))
        specs = [
            StreamSpec(F.KIND_RECTIFY, left, out["o_l"], F.MONO8, F.MONO8, camera=cam),
            StreamSpec(F.KIND_BACKPROJECT, depth, out["xyz"], F.DEPTH16, F.XYZ32F, camera=cam, mask=out["mask"], count=out["count"]),
        ]
        jobs.append((n, left, depth_np, out, specs))
    if submit:
        tickets = [be.ctx.ingest_host_submit(j[4], chunk=chunk) for j in jobs]
        assert all(b > a for a, b in zip(tickets, tickets[1:])) or len(set(tickets)) == 1 == len(tickets)
        for t in reversed(tickets):  # waiting out of order is allowed
            be.ctx.ingest_host_wait(t)
    else:
        for j in jobs:
            be.ctx.ingest_host(j[4], chunk=chunk)
    for n, left, depth_np, out, _ in jobs:
        for i in range(n):
            assert np.array_equal(view(out["o_l"])[i], orc.remap_cv(view(left)[i], *maps[0]))
            pts, msk, cnt = ob.backproject(depth_np[i], intr.matrix, m)
            assert ob.points_close(view(out["xyz"])[i], pts)[0]
            assert np.array_equal(view(out["mask"])[i], msk) and int(view(out["count"])[i]) == cnt
