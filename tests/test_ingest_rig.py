"""``IngestRig`` as a drop-in for ``CameraRig``: same frame-set selection, ingested pixels match the oracle.

Runs on the CPU emulation backend (no GPU) and, marked ``gpu``, on the real library.
"""

from __future__ import annotations

import numpy as np
import pytest

from oracle import backproject as ob
from oracle import conventions as conv
from oracle import convert as oc
from oracle import rectify as orc
from thor_slam_b200.camera import CameraRig, Extrinsics
from thor_slam_b200.camera.synthetic import SyntheticCameraConfig, SyntheticCameraSource
from thor_slam_b200.ingest.rig import IngestRig
from thor_slam_b200.slam import RecordingSlamEngine, SlamConfig

W, H = 192, 96


def make_sources():
    return [
        SyntheticCameraSource(SyntheticCameraConfig(name="oak1", resolution=(W, H), pixel_format="mono8", seed=31, pool=3,
                                                    enable_rgbd=True, rgb_resolution=(128, 72), depth_resolution=(W, H),
                                                    time_offset=0.002)),
        SyntheticCameraSource(SyntheticCameraConfig(name="oak0", resolution=(W, H), pixel_format="bgr8", seed=32, pool=3,
                                                    distortion="plumb_bob5", time_offset=0.0005)),
        SyntheticCameraSource(SyntheticCameraConfig(name="oak2", resolution=(W, H), pixel_format="nv12", seed=33, pool=3, fps=15.0)),
    ]


def rig_poses(names):
    from scipy.spatial.transform import Rotation

    rng = np.random.default_rng(5)
    out = {}
    for n in names:
        m = np.eye(4)
        m[:3, :3] = Rotation.from_rotvec(rng.uniform(-1, 1, 3)).as_matrix()
        m[:3, 3] = rng.uniform(-0.3, 0.3, 3)
        out[n] = Extrinsics.from_4x4_matrix(m)
    return out


def oracle_maps(src):
    (il, ir), (el, er) = src.get_intrinsics(), src.get_extrinsics()
    r1, r2, p1, p2 = orc.stereo_rectify_cv(il.matrix, il.coeffs, ir.matrix, ir.coeffs, (W, H), el.to_4x4_matrix(), er.to_4x4_matrix())
    return [orc.undistort_rectify_map_cv(il.matrix, il.coeffs, r1, p1, (W, H)), orc.undistort_rectify_map_cv(ir.matrix, ir.coeffs, r2, p2, (W, H))]


def run_case(ctx):
    poses = rig_poses(["oak0", "oak1", "oak2"])
    plain = CameraRig(make_sources(), queue_size=4, rig_extrinsics=poses)
    rig = IngestRig(make_sources(), queue_size=4, rig_extrinsics=poses, context=ctx)
    assert rig.get_synchronized_frames() is None  # not running -> None, like the reference
    engine = RecordingSlamEngine()
    engine.initialize(rig.calibration, SlamConfig(num_cameras=6))
    fresh = {s.name: s for s in make_sources()}
    for s in fresh.values():
        s.start()
    raw_by_seq = {n: {} for n in fresh}
    with plain, rig:
        for step in range(6):
            ref = plain.get_synchronized_frames()
            got = rig.get_synchronized_frames(with_clouds=(step == 3))
            # identical selection
            assert got.timestamp == ref.timestamp and got.max_time_delta == ref.max_time_delta
            assert list(got.frame_sets) == list(ref.frame_sets)
            for name in ref.frame_sets:
                assert [f.sequence_num for f in got.frame_sets[name].frames] == [f.sequence_num for f in ref.frame_sets[name].frames]
                maps = oracle_maps(fresh[name])
                fmt = fresh[name].cfg.pixel_format
                for i, (fg, fr) in enumerate(zip(got.frame_sets[name].frames, ref.frame_sets[name].frames)):
                    raw = fr.image
                    if fmt == "mono8":
                        want = orc.remap_cv(raw, *maps[i])
                    elif fmt == "bgr8":
                        want = orc.remap_cv(oc.bgr_to_rgb_cv(raw), *maps[i])
                    else:
                        want = orc.remap_cv(oc.nv12_to_rgb_cv(raw), *maps[i])
                    img = np.asarray(fg.image)
                    assert img.shape == want.shape and np.array_equal(img, want), (name, i, step)
                    assert fg.timestamp == fr.timestamp and fg.camera_name == fr.camera_name
            engine.process_frames(got)
            if step == 3:
                assert got.clouds is not None and list(got.clouds) == ["oak1"]
                c = got.clouds["oak1"]
                src = fresh["oak1"]
                _ri, di = src.get_rgbd_intrinsics()
                m = conv.body_T_camera(poses["oak1"].to_4x4_matrix(), src.get_rgbd_extrinsics()[1].to_4x4_matrix(), "rdf")
                depth = np.asarray(c["depth"].image)
                pts, mask, cnt = ob.backproject(depth, di.matrix, m)
                assert ob.points_close(np.asarray(c["points"]), pts)[0]
                assert np.array_equal(np.asarray(c["mask"]), mask) and int(c["count"]) == cnt
                rgb_src = src._rgbd_pool[int(c["rgb"].sequence_num) % 3][0]
                assert np.array_equal(np.asarray(c["rgb"].image), rgb_src[..., ::-1])
                ri, _di = src.get_rgbd_intrinsics()
                re, de = src.get_rgbd_extrinsics()
                want_col = ob.register_colour(depth, di.matrix, np.linalg.inv(re.to_4x4_matrix()) @ de.to_4x4_matrix(), ri.matrix,
                                              np.ascontiguousarray(rgb_src[..., ::-1]))
                assert np.array_equal(np.asarray(c["colours"]), want_col)
    # the consumer saw 6 streams in the reference's global order: sorted source name, then cam_idx
    assert [(c.source_name, c.cam_idx) for c in engine.cameras] == [("oak0", 0), ("oak0", 1), ("oak1", 0), ("oak1", 1), ("oak2", 0), ("oak2", 1)]
    assert [r.encoding for r in engine.history[-1]] == ["rgb8", "rgb8", "mono8", "mono8", "rgb8", "rgb8"]
    assert rig.get_queue_depths() == {"oak1": 0, "oak0": 0, "oak2": 0}  # stop() cleared the queues
    # re-loading rig poses re-uploads the body transforms
    rig.load_rig_extrinsics({"oak1": Extrinsics.from_4x4_matrix(np.eye(4))})
    assert np.array_equal(rig.get_rig_extrinsics("oak1").to_4x4_matrix(), np.eye(4))
    with pytest.raises(ValueError):
        rig.load_rig_extrinsics({"ghost": Extrinsics.from_4x4_matrix(np.eye(4))})


def test_ingest_rig_emulated(emu_backend):
    run_case(emu_backend.ctx)


@pytest.mark.gpu
def test_ingest_rig_gpu(gpu_backend):
    run_case(gpu_backend.ctx)


def test_config1_through_rig_emulated(emu_backend):
    """BASELINE config 1: one OAK-D, 640x400 mono + 640x400 depth through the rig on CPU (emulated kernels)."""
    src = SyntheticCameraSource(SyntheticCameraConfig(name="oak0", resolution=(640, 400), enable_rgbd=True, rgb_resolution=(640, 400),
                                                      depth_resolution=(640, 400), pool=1))
    rig = IngestRig([src], queue_size=2, context=emu_backend.ctx)
    with rig:
        sync = rig.get_synchronized_frames(with_clouds=True)
    maps = oracle_maps_size(src, 640, 400)
    left = np.asarray(sync.frame_sets["oak0"].frames[0].image)
    assert np.array_equal(left, orc.remap_cv(src._pool[0][0], *maps[0]))
    c = sync.clouds["oak0"]
    _ri, di = src.get_rgbd_intrinsics()
    m = conv.body_T_camera(np.eye(4), src.get_rgbd_extrinsics()[1].to_4x4_matrix(), "rdf")
    pts, mask, cnt = ob.backproject(src._rgbd_pool[0][1], di.matrix, m)
    assert ob.points_close(np.asarray(c["points"]), pts)[0] and int(c["count"]) == cnt


def oracle_maps_size(src, w, h):
    (il, ir), (el, er) = src.get_intrinsics(), src.get_extrinsics()
    r1, r2, p1, p2 = orc.stereo_rectify_cv(il.matrix, il.coeffs, ir.matrix, ir.coeffs, (w, h), el.to_4x4_matrix(), er.to_4x4_matrix())
    return [orc.undistort_rectify_map_cv(il.matrix, il.coeffs, r1, p1, (w, h)), orc.undistort_rectify_map_cv(ir.matrix, ir.coeffs, r2, p2, (w, h))]


def _lifetime_case(ctx) -> None:
    """Slot discipline of the pinned ring: frames stay valid for ``queue_size`` polls, a re-used slot is refused, copies live
    for ever, a frame set matched twice is ingested once, and the rig hands out the calibration of the pixels it returns."""
    src = SyntheticCameraSource(SyntheticCameraConfig(name="oak0", resolution=(W, H), pixel_format="mono8", seed=31, pool=3))
    maps = oracle_maps(src)
    q = 3
    rig = IngestRig([SyntheticCameraSource(SyntheticCameraConfig(name="oak0", resolution=(W, H), pixel_format="mono8", seed=31, pool=3))],
                    queue_size=q, context=ctx)
    with rig:
        first = rig.get_synchronized_frames()
        img0 = first.frame_sets["oak0"].frames[0].image
        kept = np.array(img0)          # a copy: unlimited lifetime
        view = np.asarray(img0)        # a view of the slot's pinned mirror
        assert np.array_equal(view, orc.remap_cv(src._pool[0][0], *maps[0]))
        launches0 = ctx.launch_count
        for _ in range(q - 1):         # the slot is still the queue's: the view stays what it was
            rig.get_synchronized_frames()
            assert np.array_equal(np.asarray(img0), kept)
        assert ctx.launch_count - launches0 == q - 1, "one ingest launch per NEW frame set"
        rig.get_synchronized_frames()  # poll number queue_size: the first slot is staged again
        with pytest.raises(RuntimeError, match="re-used"):
            np.asarray(img0)
        late = first.frame_sets["oak0"].frames[1].image  # never touched before the slot went: refused as well
        with pytest.raises(RuntimeError, match="re-used"):
            np.asarray(late)
        assert np.array_equal(kept, orc.remap_cv(src._pool[0][0], *maps[0]))
        # calibration of the returned (rectified) pixels vs the raw cameras
        raw, rect = rig.calibration.intrinsics["oak0"], rig.rectified_intrinsics("oak0")
        (il, ir), (el, er) = src.get_intrinsics(), src.get_extrinsics()
        r1, r2, p1, p2 = orc.stereo_rectify_cv(il.matrix, il.coeffs, ir.matrix, ir.coeffs, (W, H), el.to_4x4_matrix(), er.to_4x4_matrix())
        assert np.array_equal(raw[0].matrix, il.matrix) and np.any(raw[0].coeffs != 0)
        assert np.allclose(rect[0].matrix, p1[:3, :3]) and np.allclose(rect[1].matrix, p2[:3, :3]) and not np.any(rect[0].coeffs)
        rp = rig.rectification("oak0")
        assert np.allclose(rp[0]["R"], r1) and np.allclose(rp[1]["P"], p2)
    plain = IngestRig([SyntheticCameraSource(SyntheticCameraConfig(name="oak0", resolution=(W, H), pixel_format="mono8", seed=31, pool=3))],
                      queue_size=q, context=ctx, rectify=False)
    assert plain.rectification("oak0") is None and np.array_equal(plain.rectified_intrinsics("oak0")[0].matrix, il.matrix)


def test_ring_slot_lifetime_emulated(emu_backend):
    _lifetime_case(emu_backend.ctx)


@pytest.mark.gpu
def test_ring_slot_lifetime_gpu(gpu_backend):
    _lifetime_case(gpu_backend.ctx)


@pytest.mark.gpu
def test_config1_through_rig_gpu_full_size(gpu_backend):
    """BASELINE config 1 at its own size on the GPU: one OAK-D, 640x400 mono + 640x400 depth through the rig."""
    src = SyntheticCameraSource(SyntheticCameraConfig(name="oak0", resolution=(640, 400), enable_rgbd=True, rgb_resolution=(640, 400),
                                                      depth_resolution=(640, 400), pool=2))
    rig = IngestRig([src], queue_size=4, context=gpu_backend.ctx)
    maps = oracle_maps_size(src, 640, 400)
    with rig:
        for step in range(5):
            sync = rig.get_synchronized_frames(with_clouds=True)
            for cam in range(2):
                got = np.asarray(sync.frame_sets["oak0"].frames[cam].image)
                assert np.array_equal(got, orc.remap_cv(src._pool[step % 2][cam], *maps[cam])), (step, cam)
            c = sync.clouds["oak0"]
            _ri, di = src.get_rgbd_intrinsics()
            m = conv.body_T_camera(np.eye(4), src.get_rgbd_extrinsics()[1].to_4x4_matrix(), "rdf")
            pts, mask, cnt = ob.backproject(np.asarray(c["depth"].image), di.matrix, m)
            assert ob.points_close(np.asarray(c["points"]), pts)[0] and int(c["count"]) == cnt
            assert np.array_equal(np.asarray(c["mask"]), mask)
