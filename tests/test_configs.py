"""BASELINE.json configs 2-5 at their full sizes through the C ABI (``ti_ingest``), against the oracle.

Config 1 (2-stream rig through ``CameraRig`` on the CPU) is ``tests/test_ingest_rig.py``.  Full-size
frames are compared frame by frame where the oracle finishes in seconds, and through size-independent
properties elsewhere (batch == per-frame bytes, checksum of checksums, mask == depth > 0, count ==
mask.sum(), rigid-motion invariance of point distances).
"""

from __future__ import annotations

import zlib

import numpy as np
import pytest

from oracle import backproject as ob
from oracle import conventions as conv
from oracle import rectify as orc
from tests import cases
from thor_slam_b200.camera.synthetic import SyntheticCameraConfig, SyntheticCameraSource, make_depth, make_image
from thor_slam_b200.ingest import formats as F
from thor_slam_b200.ingest.context import StreamSpec

W, H = 1280, 800


def _stereo_source(i: int, pixel_format: str = "mono8", resolution=(W, H), **kw) -> SyntheticCameraSource:
    return SyntheticCameraSource(SyntheticCameraConfig(name=f"oak{i}", resolution=resolution, pixel_format=pixel_format, pool=2,
                                                       seed=1337 + 17 * i, time_offset=0.001 * i, **kw))


def _maps(src: SyntheticCameraSource, size):
    (il, ir), (el, er) = src.get_intrinsics(), src.get_extrinsics()
    r1, r2, p1, p2 = orc.stereo_rectify_cv(il.matrix, il.coeffs, ir.matrix, ir.coeffs, size, el.to_4x4_matrix(), er.to_4x4_matrix())
    return [orc.undistort_rectify_map_cv(il.matrix, il.coeffs, r1, p1, size), orc.undistort_rectify_map_cv(ir.matrix, ir.coeffs, r2, p2, size)]


# ---- config 2: 4 x OAK-D Pro stereo rig, 8 mono streams 1280x800, convert + rectify ---------------------
@pytest.mark.gpu
def test_config2_eight_mono_streams_one_launch(gpu_backend):
    be = gpu_backend
    n = 3
    specs, wants, outs = [], [], []
    for i in range(4):
        src = _stereo_source(i)
        maps = _maps(src, (W, H))
        for cam in range(2):
            slot = 2 * i + cam
            be.ctx.upload_rectify_map(slot, *maps[cam], (W, H))
            assert be.ctx.rectify_plan(slot)["variant"] == 4, "the headline rig must run the pair-window kernel"
            frames = np.stack([src._pool[b % 2][cam] for b in range(n)])
            out = be.zeros((n, H, W), np.uint8)
            specs.append(StreamSpec(F.KIND_RECTIFY, be.dev(frames), out, F.MONO8, F.MONO8, camera=slot))
            wants.append([orc.remap_cv(frames[b], *maps[cam]) for b in range(2)])
            outs.append(out)
    launches0 = be.ctx.launch_count
    be.ctx.ingest(specs)
    # ... plus ONE per-pixel launch for all slots whose exception lists overflowed (a few hundred pixels per frame set)
    plans = [be.ctx.rectify_plan(slot) for slot in range(8)]
    repairs = int(any(p["overflow_pixels"] > 0 for p in plans))
    layouts = len({(p["pitch"], p["pixels_per_window"]) for p in plans})  # slots of one layout share a launch; this rig has one
    assert layouts == 1 and plans[0]["pixels_per_window"] == 4, plans  # the headline rig runs the quad layout
    assert be.ctx.launch_count - launches0 == 1 + repairs, "8 streams x n frame sets must be ONE remap kernel launch"
    for s, out in enumerate(outs):
        got = be.host(out)
        for b in range(n):
            assert np.array_equal(got[b], wants[s][b % 2]), f"stream {s} frame {b}"


@pytest.mark.gpu
def test_config2_nv12_variant(gpu_backend):
    """Same rig delivering NV12 (1200 x 1280 buffers): mono8 is the luma plane, rectified."""
    be = gpu_backend
    src = _stereo_source(0, pixel_format="nv12")
    maps = _maps(src, (W, H))
    be.ctx.upload_rectify_map(10, *maps[0], (W, H))
    rng = np.random.default_rng(5)
    frames = np.stack([make_image(rng, "nv12", W, H) for _ in range(2)])
    out = be.zeros((2, H, W), np.uint8)
    be.ctx.ingest([StreamSpec(F.KIND_RECTIFY, be.dev(frames), out, F.NV12, F.MONO8, camera=10)])
    got = be.host(out)
    for b in range(2):
        assert np.array_equal(got[b], orc.remap_cv(np.ascontiguousarray(frames[b, :H]), *maps[0]))


# ---- config 3: 4-camera RGB-D, 1920x1080 BGR -> rgb8 + 1280x800 depth -> FLU body-frame cloud -----------
@pytest.mark.gpu
def test_config3_rgbd_to_flu_cloud(gpu_backend):
    be = gpu_backend
    rng = np.random.default_rng(33)
    specs, checks = [], []
    for i in range(4):
        src = SyntheticCameraSource(SyntheticCameraConfig(name=f"oak{i}", enable_rgbd=True, rgb_resolution=(1920, 1080),
                                                          depth_resolution=(W, H), pool=1, seed=400 + i))
        _ri, di = src.get_rgbd_intrinsics()
        rig_pose = cases.random_pose(rng)
        m = conv.body_T_camera(rig_pose, src.get_rgbd_extrinsics()[1].to_4x4_matrix(), "rdf")  # RDF rig poses -> FLU body frame
        be.ctx.upload_projection(20 + i, di.matrix, m, (W, H))
        bgr = make_image(rng, "bgr8", 1920, 1080)[None]
        depth = make_depth(rng, W, H)[None]
        rgb = be.zeros((1, 1080, 1920, 3), np.uint8)
        xyz, mask, count = be.zeros((1, H, W, 3), np.float32), be.zeros((1, H, W), np.uint8), be.zeros((1,), np.uint32)
        specs.append(StreamSpec(F.KIND_CONVERT, be.dev(bgr), rgb, F.BGR8, F.RGB8, width=1920, height=1080))
        specs.append(StreamSpec(F.KIND_BACKPROJECT, be.dev(depth), xyz, F.DEPTH16, F.XYZ32F, camera=20 + i, mask=mask, count=count))
        checks.append((bgr, depth, di.matrix, m, rgb, xyz, mask, count))
    launches0 = be.ctx.launch_count
    be.ctx.ingest(specs)
    assert be.ctx.launch_count - launches0 == 2, "one conversion launch + one back-projection launch for the whole frame set"
    for bgr, depth, k, m, rgb, xyz, mask, count in checks:
        assert np.array_equal(be.host(rgb)[0], bgr[0][..., ::-1])  # bit-exact channel swap (cv2.cvtColor BGR2RGB)
        pts, msk, cnt = ob.backproject(depth[0], k, m)
        gx, gm, gc = be.host(xyz)[0], be.host(mask)[0], int(be.host(count)[0])
        ok, worst = ob.points_close(gx, pts, cases.POINT_RTOL, cases.POINT_FLOOR)
        assert ok, f"points off by {worst:.3e}"
        assert np.array_equal(gm, msk) and np.array_equal(gm, (depth[0] > 0).astype(np.uint8)) and gc == cnt == int(gm.sum())
        # README known answer generalised: RDF -> FLU is a proper rotation, so distances between valid points survive it
        v = np.argwhere(gm)[:: max(1, int(gm.sum()) // 500)]
        cam_pts, _, _ = ob.backproject(depth[0], k, np.eye(4)[:3])
        a, b = gx[v[:-1, 0], v[:-1, 1]], gx[v[1:, 0], v[1:, 1]]
        ca, cb = cam_pts[v[:-1, 0], v[:-1, 1]], cam_pts[v[1:, 0], v[1:, 1]]
        np.testing.assert_allclose(np.linalg.norm(a - b, axis=1), np.linalg.norm(ca - cb, axis=1), rtol=1e-4, atol=1e-4)


# ---- config 4: mixed rig, 2 x OAK-D Pro + 2 x OAK-D Long Range, one launch per kind --------------------
def _config4(be, scale: int) -> None:
    """scale 1 = BASELINE sizes; scale 4 = every dimension / 4 (CPU emulation)."""
    rng = np.random.default_rng(44)
    pw, ph = 1280 // scale, 800 // scale      # Pro stereo MONO + RGB-D
    lw, lh = 1920 // scale, 1200 // scale     # LR stereo COLOR + RGB
    specs, checks, slot = [], [], 0
    for i in range(4):
        long_range = i >= 2
        w, h = (lw, lh) if long_range else (pw, ph)
        src = SyntheticCameraSource(SyntheticCameraConfig(
            name=f"{'lr' if long_range else 'pro'}{i}", resolution=(w, h), pixel_format="bgr8" if long_range else "mono8",
            enable_rgbd=True, rgb_resolution=(w, h), depth_resolution=(pw, ph), pool=1, seed=900 + i, read_imu=True))
        maps = _maps(src, (w, h))
        # per-model IMU frame convention (thor_slam scripts/run_slam.py:254-276): Pro IMU is DRB, LR IMU is already RDF
        rig_pose = cases.random_pose(rng)
        imu = conv.imu_world_extrinsics(rig_pose, np.eye(4), "rdf" if long_range else "drb")
        want_rot = rig_pose[:3, :3] @ (np.eye(3) if long_range else conv.DRB_TO_RDF[:3, :3])
        np.testing.assert_allclose(imu[:3, :3], want_rot, atol=1e-12)
        fmt_in, fmt_out = (F.BGR8, F.RGB8) if long_range else (F.MONO8, F.MONO8)
        for cam in range(2):
            be.ctx.upload_rectify_map(slot, *maps[cam], (w, h))
            img = make_image(rng, "bgr8" if long_range else "mono8", w, h)[None]
            out = be.zeros((1, h, w, 3) if long_range else (1, h, w), np.uint8)
            specs.append(StreamSpec(F.KIND_RECTIFY, be.dev(img), out, fmt_in, fmt_out, camera=slot))
            ref_src = np.ascontiguousarray(img[0][..., ::-1]) if long_range else img[0]
            checks.append(("rect", out, orc.remap_cv(ref_src, *maps[cam])))
            slot += 1
        _ri, di = src.get_rgbd_intrinsics()
        m = conv.body_T_camera(rig_pose, src.get_rgbd_extrinsics()[1].to_4x4_matrix(), "rdf")
        be.ctx.upload_projection(slot, di.matrix, m, (pw, ph))
        bgr = make_image(rng, "bgr8", w, h)[None]
        depth = make_depth(rng, pw, ph)[None]
        rgb = be.zeros((1, h, w, 3), np.uint8)
        xyz, mask, count = be.zeros((1, ph, pw, 3), np.float32), be.zeros((1, ph, pw), np.uint8), be.zeros((1,), np.uint32)
        specs.append(StreamSpec(F.KIND_CONVERT, be.dev(bgr), rgb, F.BGR8, F.RGB8, width=w, height=h))
        specs.append(StreamSpec(F.KIND_BACKPROJECT, be.dev(depth), xyz, F.DEPTH16, F.XYZ32F, camera=slot, mask=mask, count=count))
        checks.append(("rgb", rgb, bgr[0][..., ::-1]))
        checks.append(("cloud", (xyz, mask, count), ob.backproject(depth[0], di.matrix, m)))
        slot += 1
    be.ctx.ingest(specs)  # ragged shapes, four conversions, per-camera calibration: one call
    for kind, got, want in checks:
        if kind == "cloud":
            xyz, mask, count = got
            pts, msk, cnt = want
            ok, worst = ob.points_close(be.host(xyz)[0], pts, cases.POINT_RTOL, cases.POINT_FLOOR)
            assert ok, f"points off by {worst:.3e}"
            assert np.array_equal(be.host(mask)[0], msk) and int(be.host(count)[0]) == cnt
        else:
            g = be.host(got)[0]
            assert np.array_equal(g, want), f"{kind}: {(g != want).sum()} bytes differ"


def test_config4_mixed_rig_emulated(emu_backend):
    _config4(emu_backend, scale=4)


@pytest.mark.gpu
def test_config4_mixed_rig_full_size(gpu_backend):
    _config4(gpu_backend, scale=1)


# ---- config 5: batched replay, 64 frame sets x 8 streams (4 mono + 4 depth) ----------------------------
@pytest.mark.gpu
def test_config5_batched_replay_properties(gpu_backend):
    """64 frame sets in one call == the same frame sets one at a time; frames cycle with period 2, so the
    checksum of checksums has a closed form; clouds obey mask / count identities at full size."""
    import torch

    be = gpu_backend
    n = 64
    rng = np.random.default_rng(55)
    specs, keep = [], []
    for i in range(4):
        src = _stereo_source(i)
        maps = _maps(src, (W, H))
        be.ctx.upload_rectify_map(30 + i, *maps[0], (W, H))
        intr = src.get_intrinsics()[0]
        m = conv.body_T_camera(cases.random_pose(rng), src.get_extrinsics()[0].to_4x4_matrix(), "rdf")
        be.ctx.upload_projection(30 + i, intr.matrix, m, (W, H))
        two = np.stack([src._pool[b][0] for b in range(2)])
        left = torch.from_numpy(two).cuda().repeat(n // 2, 1, 1).contiguous()
        d2 = np.stack([make_depth(rng, W, H) for _ in range(2)])
        depth = torch.from_numpy(d2.view(np.int16)).cuda().view(torch.uint16).repeat(n // 2, 1, 1).contiguous()
        out = torch.zeros((n, H, W), dtype=torch.uint8, device="cuda")
        xyz = torch.zeros((n, H, W, 3), dtype=torch.float32, device="cuda")
        mask = torch.zeros((n, H, W), dtype=torch.uint8, device="cuda")
        count = torch.zeros((n,), dtype=torch.int32, device="cuda")
        specs.append(StreamSpec(F.KIND_RECTIFY, left, out, F.MONO8, F.MONO8, camera=30 + i))
        specs.append(StreamSpec(F.KIND_BACKPROJECT, depth, xyz, F.DEPTH16, F.XYZ32F, camera=30 + i, mask=mask, count=count))
        keep.append((two, maps[0], d2, intr.matrix, m, out, xyz, mask, count))
    be.ctx.ingest(specs)
    be.ctx.sync()
    torch.cuda.synchronize()
    for two, mp, d2, k, m, out, xyz, mask, count in keep:
        got = out.cpu().numpy()
        want = [orc.remap_cv(two[b], *mp) for b in range(2)]
        crcs = [zlib.crc32(got[b].tobytes()) for b in range(n)]
        assert crcs == [zlib.crc32(want[b % 2].tobytes()) for b in range(n)], "a frame of the batch differs from cv2.remap"
        assert zlib.crc32(np.asarray(crcs, np.uint32).tobytes()) == zlib.crc32(np.asarray(crcs[:2] * (n // 2), np.uint32).tobytes())
        gm, gc = mask.cpu().numpy(), count.cpu().numpy()
        for b in (0, 1, n - 1):
            pts, msk, cnt = ob.backproject(d2[b % 2], k, m)
            assert ob.points_close(xyz[b].cpu().numpy(), pts, cases.POINT_RTOL, cases.POINT_FLOOR)[0]
            assert np.array_equal(gm[b], msk) and int(gc[b]) == cnt
        assert np.array_equal(gc, gm.reshape(n, -1).sum(axis=1).astype(gc.dtype))       # count == mask.sum() for every frame
        assert torch.equal(xyz[0::2], xyz[0:1].expand(n // 2, -1, -1, -1)) and torch.equal(xyz[1::2], xyz[1:2].expand(n // 2, -1, -1, -1))
        assert not bool(torch.any(xyz[mask == 0] != 0)), "invalid pixels are written as (0,0,0)"


# ---- kernel selection: maps the pair-window kernel cannot take fall back loudly-visible, still exact -----
def _rotated_maps(w: int, h: int, deg: float):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    c, s = np.cos(np.deg2rad(deg)), np.sin(np.deg2rad(deg))
    cx, cy = w / 2, h / 2
    return ((xx - cx) * c - (yy - cy) * s + cx).astype(np.float32), ((xx - cx) * s + (yy - cy) * c + cy).astype(np.float32)


def test_layout_choice_of_the_pair_window_kernel_emulated(emu_backend):
    """The layout of a slot is decided when its map is uploaded: four pixels per window where nearly all exceptions fit the lists
    (a mildly rotated stereo map), two where they do not (a sheared map that changes source row every six pixels), two when the host
    asks for pairs; the exception capacity follows TI_OPT_RECTIFY_QUAD.  Every choice gives the same bytes (check_rectify runs both)."""
    ctx = emu_backend.ctx
    mx, my = _rotated_maps(256, 96, 1.5)
    try:
        ctx.upload_rectify_map(21, mx, my, (256, 96))
        plan = ctx.rectify_plan(21)
        assert plan["variant"] == 4 and plan["pixels_per_window"] == 4 and 0 < plan["exceptions_per_warp"] <= 24, plan
        ctx.set_option(ctx.OPT_RECTIFY_QUAD, 8)  # quads, at most 8 exception entries per (tile, warp): the rest overflows
        ctx.upload_rectify_map(21, mx, my, (256, 96))
        plan8 = ctx.rectify_plan(21)
        assert plan8["pixels_per_window"] == 2 or (plan8["exceptions_per_warp"] <= 8 and plan8["overflow_pixels"] > plan["overflow_pixels"]), plan8
        ctx.set_option(ctx.OPT_RECTIFY_QUAD, 0)
        ctx.upload_rectify_map(21, mx, my, (256, 96))
        assert ctx.rectify_plan(21)["pixels_per_window"] == 2
        ctx.set_option(ctx.OPT_RECTIFY_QUAD, 1)
        sx, sy = cases.shear_maps(256, 96)
        ctx.upload_rectify_map(21, sx, sy, (256 + 16, 160))
        plan_s = ctx.rectify_plan(21)
        assert plan_s["variant"] < 4 or plan_s["pixels_per_window"] == 2, plan_s
    finally:
        ctx.set_option(ctx.OPT_RECTIFY_QUAD, 1)
    cases.check_rectify(emu_backend, 21, mx, my, "mono8", "mono8", 256, 96, n=2, expect_variant=4)
    # the ring depth follows the shared memory the host asks the remap grids to leave free; the bytes do not
    try:
        for kb in (0, 36, 128):
            ctx.set_option(ctx.OPT_SMEM_HEADROOM_KB, kb)
            cases.check_rectify(emu_backend, 21, mx, my, "mono8", "mono8", 256, 96, n=2, expect_variant=4)
    finally:
        ctx.set_option(ctx.OPT_SMEM_HEADROOM_KB, ctx.SMEM_HEADROOM_DEFAULT_KB)


def test_rotated_map_overflows_the_exception_table_emulated(emu_backend):
    """20 degrees of roll: a third of the pixel pairs straddle two source rows.  With 32-row tiles the surplus over the 32
    exceptions a (tile, warp) holds exceeds 3 % of the image -> the slot leaves the pair-window kernel, and the plan says so.
    With 16-row tiles it stays, with full exception lists and an overflow list - and with exception entries whose
    destination offset is -1 (the pixel left of the repairing lane's first one), which an all-ones "unused" marker once
    swallowed: the result must be exact on every variant."""
    mx, my = _rotated_maps(256, 96, 20.0)
    emu_backend.ctx.upload_rectify_map(12, mx, my, (256, 96))
    assert emu_backend.ctx.rectify_plan(12)["variant"] < 4
    emu_backend.ctx.set_option(emu_backend.ctx.OPT_TMA_TILE_H, 16)
    plan16 = emu_backend.ctx.rectify_plan(12)
    emu_backend.ctx.set_option(emu_backend.ctx.OPT_TMA_TILE_H, 32)
    assert plan16["variant"] == 4 and plan16["exceptions_per_warp"] == 32 and plan16["overflow_pixels"] > 0, plan16
    cases.check_rectify(emu_backend, 12, mx, my, "mono8", "mono8", 256, 96, n=2)


def test_small_roll_stays_on_the_pair_kernel_emulated(emu_backend):
    mx, my = _rotated_maps(256, 96, 1.5)
    cases.check_rectify(emu_backend, 13, mx, my, "mono8", "mono8", 256, 96, n=2, expect_variant=4, expect_exceptions=True)


@pytest.mark.gpu
def test_rotated_maps_gpu(gpu_backend):
    for slot, deg, variant in ((14, 20.0, None), (15, 1.5, 4), (16, 3.0, None)):
        mx, my = _rotated_maps(W, H, deg)
        cases.check_rectify(gpu_backend, slot, mx, my, "mono8", "mono8", W, H, n=2, expect_variant=variant)


# ---- randomised maps: scale, shear, roll, barrel distortion, offsets that push the source box over every border --------
def _random_map(rng: np.random.Generator, w: int, h: int):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    cx, cy = w / 2 + rng.uniform(-20, 20), h / 2 + rng.uniform(-20, 20)
    s = rng.uniform(0.8, 1.25)
    roll = np.deg2rad(rng.uniform(-4, 4))
    k1 = rng.uniform(-0.25, 0.25)
    x, y = (xx - cx) / w, (yy - cy) / w
    r2 = x * x + y * y
    f = 1 + k1 * r2
    xd, yd = x * f, y * f
    c, sn = np.cos(roll), np.sin(roll)
    mx = (xd * c - yd * sn) * w * s + cx + rng.uniform(-15, 15) + rng.uniform(-0.05, 0.05) * yy
    my = (xd * sn + yd * c) * w * s * rng.uniform(0.9, 1.1) + cy + rng.uniform(-15, 15)
    return mx.astype(np.float32), my.astype(np.float32)


def _random_maps_case(be, n_maps: int, w: int, h: int, seed: int) -> None:
    rng = np.random.default_rng(seed)
    ran_pair = 0
    for i in range(n_maps):
        mx, my = _random_map(rng, w, h)
        cases.check_rectify(be, 18, mx, my, "mono8", "mono8", w, h, n=2, seed=seed + i)
        ran_pair += be.ctx.rectify_plan(18)["variant"] == 4
    assert ran_pair >= n_maps // 2, "most of these maps are meant to qualify for the pair-window kernel"


def test_random_maps_emulated(emu_backend):
    _random_maps_case(emu_backend, 3, 256, 64, seed=77)


@pytest.mark.gpu
def test_random_maps_gpu(gpu_backend):
    _random_maps_case(gpu_backend, 12, 640, 416, seed=78)
    cases.check_rectify(gpu_backend, 18, *_random_map(np.random.default_rng(5), 1920, 1200), "bgr8", "rgb8", 1920, 1200, n=1)
