"""``-m gpu``: parity of the CUDA path (through the C ABI) against the oracle and golden fixtures."""

from __future__ import annotations

import numpy as np
import pytest

from oracle import backproject as ob
from oracle import conventions as conv
from oracle import rectify as orc
from tests import cases
from tests.conftest import GOLDEN
from thor_slam_b200.camera.synthetic import make_depth
from thor_slam_b200.ingest import formats as F
from thor_slam_b200.ingest.context import StreamSpec

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("s,d", cases.CONVERSIONS)
@pytest.mark.parametrize("w,h", [(64, 32), (50, 22), (1280, 800)])
def test_convert(gpu_backend, s, d, w, h):
    cases.check_convert(gpu_backend, s, d, w, h, n=3)


def test_convert_1080p_bgr(gpu_backend):
    cases.check_convert(gpu_backend, "bgr8", "rgb8", 1920, 1080, n=2)


def test_convert_empty_batch(gpu_backend):
    cases.check_convert(gpu_backend, "bgr8", "rgb8", 64, 32, n=0)


def test_convert_golden(gpu_backend):
    g = np.load(GOLDEN / "cv_arith.npz")
    be = gpu_backend
    for src_key, s, d, want_key in [("bgr", "bgr8", "rgb8", "bgr2rgb"), ("bgr", "bgr8", "mono8", "bgr2gray"),
                                    ("nv12", "nv12", "rgb8", "nv122rgb"), ("nv12", "nv12", "bgr8", "nv122bgr"),
                                    ("nv12", "nv12", "mono8", "nv122gray"), ("nv12lim", "nv12", "rgb8", "nv12lim2rgb")]:
        src = g[src_key][None]
        want = g[want_key]
        dst = be.zeros((1, *want.shape), np.uint8)
        be.ctx.convert(be.dev(src), dst, s, d, 48, 32)
        assert np.array_equal(be.host(dst)[0], want), want_key


@pytest.mark.parametrize("s,d", cases.RECTIFY_CONVERSIONS)
def test_rectify_stereo_maps_small(gpu_backend, s, d):
    _, maps = cases.stereo_maps(192, 96)
    cases.check_rectify(gpu_backend, 0, *maps[1], s, d, 192, 96)


@pytest.mark.parametrize("s,d", cases.RECTIFY_CONVERSIONS)
def test_rectify_border(gpu_backend, s, d):
    mx, my = cases.edge_maps(160, 64)
    cases.check_rectify(gpu_backend, 1, mx, my, s, d, 160, 64)


@pytest.mark.parametrize("distortion", ["rational14", "plumb_bob5", "fisheye4", "none"])
def test_rectify_full_size_mono(gpu_backend, distortion):
    """BASELINE config 2 stream shape, every distortion model the reference's CameraInfo rule knows."""
    _, maps = cases.stereo_maps(1280, 800, seed=11, distortion=distortion)
    for cam, (mx, my) in enumerate(maps):
        # no silent fall-back: every distortion model of the reference's CameraInfo rule runs the pair-window kernel
        cases.check_rectify(gpu_backend, 8 + cam, mx, my, "mono8", "mono8", 1280, 800, n=3, expect_variant=4)


def test_rectify_overflowing_exception_lists(gpu_backend):
    mx, my = cases.shear_maps(1280, 800)
    cases.check_rectify(gpu_backend, 12, mx, my, "mono8", "mono8", 1296, 1024, n=2, expect_variant=4, expect_overflow=True)


def test_rectify_full_size_colour(gpu_backend):
    _, maps = cases.stereo_maps(1920, 1200, seed=5)
    cases.check_rectify(gpu_backend, 10, *maps[0], "bgr8", "rgb8", 1920, 1200, n=2)
    cases.check_rectify(gpu_backend, 10, *maps[0], "bgr8", "mono8", 1920, 1200, n=1)
    cases.check_rectify(gpu_backend, 10, *maps[0], "nv12", "rgb8", 1920, 1200, n=1)


def test_rectify_largest_sensor_mode(gpu_backend):
    """4224 x 3136, the largest sensor mode the reference's driver lists (luxonis.py:36-44): BGR8 -> RGB8 and mono rectify on
    the window kernels, bit-exact."""
    _, maps = cases.stereo_maps(4224, 3136, seed=7)
    cases.check_rectify(gpu_backend, 13, *maps[0], "bgr8", "rgb8", 4224, 3136, n=1, expect_variant=4)
    cases.check_rectify(gpu_backend, 13, *maps[1], "mono8", "mono8", 4224, 3136, n=2, expect_variant=4)


def test_rectify_resize_ragged_outside(gpu_backend):
    yy, xx = np.mgrid[0:51, 0:99].astype(np.float32)
    cases.check_rectify(gpu_backend, 2, xx * 1.1 + 0.3, yy * 1.05 + 0.7, "mono8", "mono8", 110, 60)
    yy, xx = np.mgrid[0:400, 0:640].astype(np.float32)
    cases.check_rectify(gpu_backend, 3, xx * 2.0 + 0.25, yy * 2.0 + 0.75, "mono8", "mono8", 1280, 800, expect_variant=4)  # 2x downscale (slam_config.yaml:7)
    assert gpu_backend.ctx.rectify_plan(3)["pitch"] == 320, "the 2 x downscale map must run the wide-pitch pair-window kernel"
    cases.check_rectify(gpu_backend, 3, xx * 2.0 + 0.25, yy * 2.0 + 0.75, "bgr8", "rgb8", 1280, 800, n=2, expect_variant=4)  # colour: 1024-byte rows
    yy, xx = np.mgrid[0:600, 0:960].astype(np.float32)
    cases.check_rectify(gpu_backend, 5, xx * 2.0 + 0.6, yy * 2.0 + 0.3, "bgr8", "rgb8", 1920, 1200, n=2, expect_variant=4)  # the LR colour stereo stream, halved
    mx = np.full((32, 128), -50.0, np.float32)
    cases.check_rectify(gpu_backend, 4, mx, mx.copy(), "mono8", "mono8", 128, 32)


def test_rectify_odd_output_width_on_fast_kernels(gpu_backend):
    yy, xx = np.mgrid[0:51, 0:99].astype(np.float32)
    cases.check_rectify(gpu_backend, 7, xx * 1.1 + 0.3, yy * 1.05 + 0.7, "mono8", "mono8", 112, 60)
    yy, xx = np.mgrid[0:40, 0:130].astype(np.float32)
    cases.check_rectify(gpu_backend, 7, xx * 0.8 + 1.5, yy * 1.2 + 0.25, "mono8", "mono8", 112, 60)


def test_rectify_golden(gpu_backend):
    g = np.load(GOLDEN / "cv_arith.npz")
    be = gpu_backend
    for side in "lr":
        be.ctx.upload_rectify_map(20, g[f"mapx_{side}"], g[f"mapy_{side}"], (160, 100))
        dst = be.zeros((1, 100, 160), np.uint8)
        be.ctx.rectify(20, be.dev(g[f"img_{side}"][None]), dst, "mono8", "mono8")
        assert np.array_equal(be.host(dst)[0], g[f"rect_{side}"])
    be.ctx.upload_rectify_map(21, g["edge_mapx"], g["edge_mapy"], (48, 32))
    dst = be.zeros((1, 32, 48), np.uint8)
    be.ctx.rectify(21, be.dev(g["bgr2gray"][None]), dst, "mono8", "mono8")
    assert np.array_equal(be.host(dst)[0], g["edge_rect_u8"])
    dst3 = be.zeros((1, 32, 48, 3), np.uint8)
    be.ctx.rectify(21, be.dev(g["bgr"][None]), dst3, "bgr8", "rgb8")
    assert np.array_equal(be.host(dst3)[0], g["edge_rect_c3"][..., ::-1])


def test_rectify_identity_is_copy(gpu_backend):
    """Size-independent property: the identity map reproduces the input bit for bit."""
    yy, xx = np.mgrid[0:800, 0:1280].astype(np.float32)
    be = gpu_backend
    be.ctx.upload_rectify_map(22, xx, yy, (1280, 800))
    rng = np.random.default_rng(4)
    src = rng.integers(0, 256, size=(4, 800, 1280), dtype=np.uint8)
    dst = be.zeros(src.shape, np.uint8)
    be.ctx.rectify(22, be.dev(src), dst, "mono8", "mono8")
    assert np.array_equal(be.host(dst), src)


@pytest.mark.parametrize("w,h", [(64, 40), (50, 22), (1280, 800)])
@pytest.mark.parametrize("frame", ["rdf", "flu"])
def test_backproject(gpu_backend, w, h, frame):
    cases.check_backproject(gpu_backend, 30, w, h, n=2, rig_frame=frame)


def test_backproject_extremes(gpu_backend):
    depth = np.zeros((3, 16, 64), np.uint16)
    depth[1] = 65535
    depth[2, ::2, 1::3] = 1
    cases.check_backproject(gpu_backend, 31, 64, 16, depth=depth)


def test_backproject_golden_depth_and_readme_vector(gpu_backend):
    """Pinned inputs: golden depth frame; README known answer rdf_to_flu @ [1,0,0,1] = [0,-1,0,1]."""
    g = np.load(GOLDEN / "cv_arith.npz")
    be = gpu_backend
    # K = identity-ish so that pixel (u=1, v=0) at d = 1000 mm back-projects to RDF (1, 0, 1)
    k = np.array([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1]])
    be.ctx.upload_projection(32, k, conv.RDF_TO_FLU, (8, 1))
    depth = np.zeros((1, 1, 8), np.uint16)
    depth[0, 0, 1] = 1000
    xyz = be.zeros((1, 1, 8, 3), np.float32)
    be.ctx.backproject(32, be.dev(depth), xyz)
    np.testing.assert_allclose(be.host(xyz)[0, 0, 1], [1.0, -1.0, 0.0], atol=1e-6)  # (x,y,z)_rdf=(1,0,1) -> flu (1,-1,0)
    cases.check_backproject(gpu_backend, 33, 48, 32, depth=g["depth"][None])


def test_ingest_fused_matches_single_calls(gpu_backend):
    be = gpu_backend
    w, h, n = 640, 400, 3
    rng = np.random.default_rng(9)
    s, maps = cases.stereo_maps(w, h, seed=9)
    be.ctx.upload_rectify_map(40, *maps[0], (w, h))
    be.ctx.upload_rectify_map(41, *maps[1], (w, h))
    intr = s.get_intrinsics()[0]
    m = conv.body_T_camera(cases.random_pose(rng), s.get_extrinsics()[0].to_4x4_matrix(), "rdf")
    be.ctx.upload_projection(40, intr.matrix, m, (w, h))
    left = cases.make_batch(rng, "mono8", w, h, n)
    right = cases.make_batch(rng, "nv12", w, h, n)
    bgr = cases.make_batch(rng, "bgr8", w, h, n)
    depth = np.stack([make_depth(rng, w, h) for _ in range(n)])
    o_l, o_r = be.zeros((n, h, w), np.uint8), be.zeros((n, h, w), np.uint8)
    o_rgb = be.zeros((n, h, w, 3), np.uint8)
    xyz, mask, count = be.zeros((n, h, w, 3), np.float32), be.zeros((n, h, w), np.uint8), be.zeros((n,), np.uint32)
    be.ctx.ingest([
        StreamSpec(F.KIND_RECTIFY, be.dev(left), o_l, F.MONO8, F.MONO8, camera=40),
        StreamSpec(F.KIND_RECTIFY, be.dev(right), o_r, F.NV12, F.MONO8, camera=41),
        StreamSpec(F.KIND_CONVERT, be.dev(bgr), o_rgb, F.BGR8, F.RGB8, width=w, height=h),
        StreamSpec(F.KIND_BACKPROJECT, be.dev(depth), xyz, F.DEPTH16, F.XYZ32F, camera=40, mask=mask, count=count),
    ])
    gl, gr, grgb, gx, gm, gc = (be.host(x) for x in (o_l, o_r, o_rgb, xyz, mask, count))
    for i in range(n):
        assert np.array_equal(gl[i], orc.remap_cv(left[i], *maps[0]))
        assert np.array_equal(gr[i], orc.remap_cv(np.ascontiguousarray(right[i, :h]), *maps[1]))
        assert np.array_equal(grgb[i], bgr[i, ..., ::-1])
        pts, msk, cnt = ob.backproject(depth[i], intr.matrix, m)
        assert ob.points_close(gx[i], pts)[0] and np.array_equal(gm[i], msk) and int(gc[i]) == cnt


@pytest.mark.parametrize("n,chunk", [(7, 3), (21, 8), (2, 8), (24, 8)])
def test_ingest_host_pipeline(gpu_backend, n, chunk):
    """Host buffers in, host buffers out, chunked H2D / kernels / D2H; ramped and ragged chunks."""
    cases.check_host_pipeline(gpu_backend, 42, 640, 400, [n], chunk, pinned=True, submit=False)


@pytest.mark.parametrize("batches,chunk", [([5, 9, 3], 4), ([2] * 11, 2), ([8, 0, 8], 8)])
def test_ingest_host_submit_wait(gpu_backend, batches, chunk):
    """Several batches in flight at once (more than the ticket ring holds in the second case); waits out of order."""
    cases.check_host_pipeline(gpu_backend, 43, 640, 400, batches, chunk, pinned=True, submit=True)


def test_ingest_host_wait_errors(gpu_backend):
    gpu_backend.ctx.ingest_host_wait(0)  # the ticket of an empty submission
    with pytest.raises(ValueError):
        gpu_backend.ctx.ingest_host_wait(1 << 40)


def test_errors(gpu_backend):
    import torch

    ctx = gpu_backend.ctx
    a = torch.zeros((1, 8, 16), dtype=torch.uint8, device="cuda")
    with pytest.raises(RuntimeError):
        ctx.rectify(60, a, a.clone(), "mono8", "mono8")  # slot never uploaded
    with pytest.raises(ValueError):
        ctx.convert(a, a.clone(), "mono8", "rgb8", 16, 8)  # unsupported conversion
    with pytest.raises(ValueError):
        ctx.convert(a.cpu(), a.clone(), "mono8", "mono8", 16, 8)  # host tensor on the device API


@pytest.mark.parametrize("scene", ["room", "noise"])
def test_voxel_cloud_config5_full_size(gpu_backend, scene):
    """BASELINE config 5 frame sets: 4 depth cameras of 1280x800 fused per frame set into one voxel list (0.05 m, 10 m cap)."""
    n = cases.check_voxel(gpu_backend, 30, [(1280, 800)] * 4, n=2, scene=scene, seed=21)
    assert n > 0


def test_voxel_cloud_small_ragged_and_edges(gpu_backend):
    cases.check_voxel(gpu_backend, 34, [(128, 64), (100, 45)], n=3, scene="room")
    cases.check_voxel(gpu_backend, 34, [(128, 64), (100, 45)], n=3, scene="noise", voxel=0.02, max_depth_mm=0)
    depth = np.zeros((3, 32, 64), np.uint16)
    depth[1] = 65535
    depth[1, 5, 7] = 10000
    depth[2] = 1
    cases.check_voxel(gpu_backend, 36, [(64, 32)], depth=[depth], set_base=2045, tag=255)
    cases.check_voxel(gpu_backend, 36, [(64, 32)], n=0)
    cases.check_voxel(gpu_backend, 36, [(128, 64)], n=1, scene="noise", capacity=100)
    with pytest.raises(ValueError):
        cases.check_voxel(gpu_backend, 36, [(64, 32)], n=1, voxel=0.001, max_depth_mm=0)


def test_voxel_cloud_many_launches_epoch_wrap(gpu_backend):
    """More than 255 launches on one table: the epoch byte wraps and the table is cleared."""
    for i in range(130):  # two launches per check
        cases.check_voxel(gpu_backend, 37, [(64, 32)], n=1, scene="noise", seed=100 + i)


def test_backproject_with_fused_colour(gpu_backend):
    """Fused depth -> xyz + mask + count + colour: the reciprocal fast path with its guard band must select
    the very pixels the float32 oracle (IEEE divisions) selects - full size, a small ragged size, and the all-on-x.5 worst case."""
    cases.check_backproject_colour(gpu_backend, 40, 1280, 800, 1920, 1080, n=3)
    cases.check_backproject_colour(gpu_backend, 41, 64, 40, 96, 54)
    cases.check_backproject_colour(gpu_backend, 42, 640, 400, 640, 400, on_half_pixels=True)
    cases.check_backproject_colour(gpu_backend, 43, 50, 22, 70, 30)


@pytest.mark.parametrize("w,h", [(1280, 800), (50, 22)])
def test_depth_stats(gpu_backend, w, h):
    cases.check_depth_stats(gpu_backend, w, h, n=4)
