"""Kernel sources run thread-by-thread on the CPU (tests/emu) against the oracle - small sizes.

This is a pre-flight for the ``-m gpu`` parity tests: same cases, same C ABI, same kernel source
text, no GPU.  It is test infrastructure; the product library is the nvcc build only.
"""

from __future__ import annotations

import numpy as np
import pytest

from tests import cases
from thor_slam_b200.ingest import formats as F
from thor_slam_b200.ingest.context import StreamSpec


@pytest.mark.parametrize("s,d", cases.CONVERSIONS)
@pytest.mark.parametrize("w,h", [(64, 32), (50, 22)])  # vector path / scalar path
def test_convert(emu_backend, s, d, w, h):
    cases.check_convert(emu_backend, s, d, w, h, n=2)


def test_convert_empty_batch(emu_backend):
    cases.check_convert(emu_backend, "bgr8", "rgb8", 64, 32, n=0)


@pytest.mark.parametrize("s,d", cases.RECTIFY_CONVERSIONS)
def test_rectify_stereo_maps(emu_backend, s, d):
    _, maps = cases.stereo_maps(192, 96)
    cases.check_rectify(emu_backend, 0, *maps[0], s, d, 192, 96, expect_variant=4)


@pytest.mark.parametrize("s,d", cases.RECTIFY_CONVERSIONS)
def test_rectify_border(emu_backend, s, d):
    mx, my = cases.edge_maps(160, 64)
    cases.check_rectify(emu_backend, 1, mx, my, s, d, 160, 64, expect_variant=4, expect_exceptions=True)


def test_rectify_overflowing_exception_lists(emu_backend):
    """More than 32 exception pairs in a (tile, warp): the pair-window kernel keeps the slot, the surplus pixels are repaired
    by the per-pixel pass after it."""
    mx, my = cases.shear_maps(256, 64)
    cases.check_rectify(emu_backend, 9, mx, my, "mono8", "mono8", 272, 112, n=2, expect_variant=4, expect_overflow=True)


@pytest.mark.parametrize("s,d", [("bgr8", "mono8"), ("nv12", "rgb8")])
def test_rectify_two_pass_in_chunks(emu_backend, s, d):
    """The two-pass conversions cut the batch into chunks whose scratch stays in the L2: a 1 KB budget makes every frame its
    own chunk (5 frames -> 5 convert + 5 remap launches), the bytes must not care."""
    _, maps = cases.stereo_maps(192, 96)
    ctx = emu_backend.ctx
    ctx.set_option(ctx.OPT_L2_SCRATCH_KB, 1)
    try:
        launches0 = ctx.launch_count
        cases.check_rectify(emu_backend, 0, *maps[0], s, d, 192, 96, n=5, expect_variant=4)
        assert ctx.launch_count - launches0 >= 10
    finally:
        ctx.set_option(ctx.OPT_L2_SCRATCH_KB, 0)


def test_more_two_pass_streams_than_one_conversion_launch_takes(emu_backend):
    """34 BGR8 -> MONO8 rectify streams in one ti_ingest call: the gray pre-pass is chunked by TI_MAX_STREAMS (32) and the remap by
    the 16 jobs a window-kernel launch takes - round 1 failed here with 'too many convert streams'."""
    _, maps = cases.stereo_maps(192, 96)
    ctx = emu_backend.ctx
    ctx.upload_rectify_map(0, *maps[0], (192, 96))
    rng = np.random.default_rng(3)
    src = cases.make_batch(rng, "bgr8", 192, 96, 1)
    want = cases.orc.remap_cv(np.ascontiguousarray(cases.oracle_convert(src[0], "bgr8", "mono8")), *maps[0])
    outs = [np.zeros((1, 96, 192), np.uint8) for _ in range(34)]
    ctx.ingest([StreamSpec(F.KIND_RECTIFY, src, o, F.BGR8, F.MONO8, camera=0) for o in outs])
    for o in outs:
        assert np.array_equal(o[0], want)


def test_rectify_resize_and_ragged(emu_backend):
    yy, xx = np.mgrid[0:51, 0:99].astype(np.float32)
    cases.check_rectify(emu_backend, 2, xx * 1.1 + 0.3, yy * 1.05 + 0.7, "mono8", "mono8", 110, 60)  # direct kernel
    yy, xx = np.mgrid[0:72, 0:200].astype(np.float32)
    cases.check_rectify(emu_backend, 3, xx * 0.75 + 3.25, yy * 0.8 + 0.5, "mono8", "mono8", 160, 64)  # tiled, ragged tiles


def test_rectify_odd_output_width_on_fast_kernels(emu_backend):
    """dst_w odd (no 16-bit stores possible) while the source still qualifies for the TMA / thread-staged kernels."""
    yy, xx = np.mgrid[0:51, 0:99].astype(np.float32)
    cases.check_rectify(emu_backend, 7, xx * 1.1 + 0.3, yy * 1.05 + 0.7, "mono8", "mono8", 112, 60)
    yy, xx = np.mgrid[0:40, 0:130].astype(np.float32)  # two tiles wide, second one 2 pixels
    cases.check_rectify(emu_backend, 7, xx * 0.8 + 1.5, yy * 1.2 + 0.25, "mono8", "mono8", 112, 60)


def test_rectify_downscale_runs_the_wide_pitch_kernel(emu_backend):
    """A 2 x downscale map (the output_resolution of config/slam_config.yaml:7,26 done on the host): 128 output pixels sample 256
    source pixels, more than the 192-byte rows of the standard boxes - the slot plans the pair-window kernel with 320-byte rows."""
    yy, xx = np.mgrid[0:64, 0:160].astype(np.float32)
    mx, my = (xx * 2.0 + 0.25 + 0.004 * yy).astype(np.float32), (yy * 2.0 + 0.75 - 0.003 * xx).astype(np.float32)
    emu_backend.ctx.upload_rectify_map(15, mx, my, (320, 128))
    plan = emu_backend.ctx.rectify_plan(15)
    assert plan["variant"] == 4 and plan["pitch"] == 320, plan
    cases.check_rectify(emu_backend, 15, mx, my, "mono8", "mono8", 320, 128, n=2, expect_variant=4)
    cases.check_rectify(emu_backend, 15, mx, my, "nv12", "mono8", 320, 128, n=2, expect_variant=4)
    cases.check_rectify(emu_backend, 15, mx, my, "bgr8", "mono8", 320, 128, n=1, expect_variant=4)
    cases.check_rectify(emu_backend, 15, mx, my, "bgr8", "rgb8", 320, 128, n=2, expect_variant=4)  # asserts the 3-channel window kernel (1024-byte rows)
    cases.check_rectify(emu_backend, 15, mx, my, "nv12", "rgb8", 320, 128, n=1, expect_variant=4)


def test_rectify_beyond_2046_pixels(emu_backend):
    """Source coordinates above the 11-bit limit of round 1's packed LUT (the driver lists 4000 x 3000 and 4224 x 3136 sensor
    modes, luxonis.py:36-44): a 2560-wide strip, mono and colour."""
    yy, xx = np.mgrid[0:40, 0:2560].astype(np.float32)
    mx, my = (xx * 0.998 + 2.3 + 0.01 * yy).astype(np.float32), (yy * 1.02 + 0.4 + 0.0007 * xx).astype(np.float32)
    cases.check_rectify(emu_backend, 14, mx, my, "mono8", "mono8", 2560, 48, n=1, expect_variant=4)
    cases.check_rectify(emu_backend, 14, mx, my, "bgr8", "rgb8", 2560, 48, n=1, expect_variant=4)
    with pytest.raises(ValueError):
        emu_backend.ctx.upload_rectify_map(14, mx, my, (8200, 48))


def test_rectify_all_outside(emu_backend):
    mx = np.full((32, 128), -50.0, np.float32)
    cases.check_rectify(emu_backend, 4, mx, mx.copy(), "mono8", "mono8", 128, 32)


@pytest.mark.parametrize("w,h", [(64, 40), (50, 22)])
@pytest.mark.parametrize("frame", ["rdf", "flu"])
def test_backproject(emu_backend, w, h, frame):
    cases.check_backproject(emu_backend, 5, w, h, rig_frame=frame)


def test_backproject_extremes(emu_backend):
    depth = np.zeros((3, 16, 64), np.uint16)
    depth[1] = 65535
    depth[2, ::2, 1::3] = 1
    cases.check_backproject(emu_backend, 6, 64, 16, depth=depth)


@pytest.mark.parametrize("w,h", [(64, 40), (50, 22)])
def test_depth_stats(emu_backend, w, h):
    cases.check_depth_stats(emu_backend, w, h)


def test_backproject_with_fused_colour(emu_backend):
    cases.check_backproject_colour(emu_backend, 7, 64, 40, 96, 54)
    cases.check_backproject_colour(emu_backend, 7, 64, 40, 64, 40, on_half_pixels=True)
    cases.check_backproject_colour(emu_backend, 7, 50, 22, 70, 30)  # unaligned width: scalar kernel + stand-alone registration


@pytest.mark.parametrize("scene", ["room", "noise"])
def test_voxel_cloud(emu_backend, scene):
    """Two cameras of different size (vector and scalar load paths, ragged tiles) fused per frame set."""
    cases.check_voxel(emu_backend, 20, [(128, 64), (100, 45)], n=2, scene=scene)


def test_voxel_cloud_edges(emu_backend):
    depth = np.zeros((3, 32, 64), np.uint16)  # set 0: nothing valid; set 1: all beyond the cap but one pixel; set 2: all one voxel
    depth[1] = 65535
    depth[1, 5, 7] = 10000
    depth[2] = 1
    cases.check_voxel(emu_backend, 22, [(64, 32)], depth=[depth], set_base=2045, tag=255)
    cases.check_voxel(emu_backend, 22, [(64, 32)], n=0)
    cases.check_voxel(emu_backend, 22, [(128, 64)], n=1, scene="noise", capacity=100)  # truncated list, full count
    with pytest.raises(ValueError):
        cases.check_voxel(emu_backend, 22, [(64, 32)], n=1, voxel=0.001, max_depth_mm=0)  # 65 m / 1 mm overflows the key fields
    with pytest.raises(ValueError):
        cases.check_voxel(emu_backend, 22, [(64, 32)], n=2, set_base=2047)


@pytest.mark.parametrize("submit", [False, True])
def test_host_pipeline_chunk_schedule(emu_backend, submit):
    """ti_ingest_host / _submit / _wait: ramped, ragged and empty batches through the three chunk slots."""
    cases.check_host_pipeline(emu_backend, 44, 64, 32, [13, 0, 2, 7] if submit else [13, 2], 4, pinned=False, submit=submit)


def test_errors(emu_backend):
    ctx = emu_backend.ctx
    a = np.zeros((1, 8, 16), np.uint8)
    with pytest.raises(RuntimeError):  # TI_ESTATE: slot never uploaded
        ctx.rectify(40, a, a.copy(), "mono8", "mono8")
    with pytest.raises(ValueError):  # TI_EINVAL: unsupported conversion
        ctx.convert(a, a.copy(), "mono8", "rgb8", 16, 8)
    with pytest.raises(ValueError):
        ctx.upload_rectify_map(99, np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32), (4, 4))
    with pytest.raises(ValueError):
        ctx.convert(np.zeros((1, 9, 6), np.uint8), np.zeros((1, 6, 6, 3), np.uint8), "nv12", "rgb8", 5, 6)  # odd NV12


def test_copy_async_and_plans(emu_backend):
    """ti_copy_async (the rig's staging copies): whole buffers, pre-resolved plans, size mismatch, empty copy."""
    import torch

    ctx = emu_backend.ctx
    src = torch.arange(4096, dtype=torch.uint8).reshape(4, 1024)
    dst = torch.zeros_like(src)
    ctx.copy_async(dst[1], src[2])
    assert torch.equal(dst[1], src[2]) and not dst[0].any()
    plan = ctx.copy_plan(dst[3], src[0])
    ctx.copy_planned(plan)
    assert torch.equal(dst[3], src[0])
    ctx.copy_async(dst[0, :0], src[0, :0])  # nothing to copy: accepted
    with pytest.raises(ValueError):
        ctx.copy_async(dst[0], src[0, :512])
    with pytest.raises(ValueError):
        ctx.copy_plan(dst, src[0])
