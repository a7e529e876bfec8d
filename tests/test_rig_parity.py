"""Our ``CameraRig`` / containers against traces recorded from the reference's ``CameraRig``.

``tests/golden/rig_sync.json`` was produced by running ``thor_slam.camera.rig.CameraRig`` (the
reference, unmodified) on the scenario below; here the same scenario goes through
``thor_slam_b200.camera.rig.CameraRig`` and must pick the same frames with the same numbers.
"""

from __future__ import annotations

import json

import numpy as np
import pytest

from tests.conftest import GOLDEN
from thor_slam_b200.camera import CameraRig, Extrinsics, FrameSet, IMUExtrinsics, RigCalibration
from thor_slam_b200.camera.rig import _Ring
from thor_slam_b200.camera.synthetic import SyntheticCameraConfig, SyntheticCameraSource
from thor_slam_b200.camera.types import IPv4

SCENARIO = [("oak_b", 30.0, 0.0031, True, True), ("oak_a", 30.0, 0.0007, True, False), ("oak_c", 20.0, 0.0190, False, False)]
QUEUE, STEPS = 5, 24


def make_sources():
    return [
        SyntheticCameraSource(SyntheticCameraConfig(name=name, fps=fps, time_offset=off, stereo=stereo, read_imu=imu,
                                                    resolution=(64, 40), pixel_format="mono8" if stereo else "bgr8",
                                                    seed=11 + i, pool=2))
        for i, (name, fps, off, stereo, imu) in enumerate(SCENARIO)
    ]


def entry(sync):
    if sync is None:
        return None
    return {
        "timestamp": sync.timestamp,
        "max_time_delta": sync.max_time_delta,
        "order": list(sync.frame_sets.keys()),
        "picked": {n: [f.sequence_num for f in fs.frames] for n, fs in sync.frame_sets.items()},
        "fs_timestamp": {n: fs.timestamp for n, fs in sync.frame_sets.items()},
        "sensor_timestamp": sync.sensor_timestamp,
        "n_all_frames": len(sync.get_all_frames()),
    }


def test_sync_trace_matches_reference():
    ref = json.loads((GOLDEN / "rig_sync.json").read_text())
    rig = CameraRig(make_sources(), queue_size=QUEUE, imu_source="oak_b")
    assert entry(rig.get_synchronized_frames()) == ref["before_start"]  # None while stopped
    rig.start()
    latest = []
    for step in range(STEPS):
        assert entry(rig.get_synchronized_frames()) == ref["sync"][step], f"step {step}"
        assert rig.get_queue_depths() == ref["depths"][step]
        if step % 6 == 5:
            latest.append(entry(rig.get_latest_frames()))
    assert latest == ref["latest"]
    assert rig.prune_old_frames(0.05) == ref["pruned"]
    assert rig.get_queue_depths() == ref["depths_after_prune"]
    rig.stop()
    assert rig.get_queue_depths() == ref["depths_after_stop"]
    assert entry(rig.get_synchronized_frames()) == ref["after_stop"]


def test_error_conventions_match_reference():
    ref = json.loads((GOLDEN / "rig_sync.json").read_text())["errors"]
    srcs = make_sources()
    with pytest.raises(ValueError):
        CameraRig(srcs, imu_source="nope")
    with pytest.raises(ValueError):
        CameraRig(srcs, imu_source="oak_a")  # has no sensor data
    rig = CameraRig(srcs)
    with pytest.raises(ValueError):
        rig.load_rig_extrinsics({"ghost": Extrinsics(np.eye(3), np.zeros(3))})
    with pytest.raises(ValueError):
        Extrinsics.from_4x4_matrix(np.eye(3))
    with pytest.raises(ValueError):
        FrameSet.from_frames([], "x")
    assert set(ref.values()) == {"ValueError"}
    with pytest.raises(RuntimeError):
        srcs[0].get_latest_frames()  # before start(): luxonis.py:765-766
    with pytest.raises(ValueError):
        IPv4("300.1.1.1")
    assert str(IPv4("192.168.2.25")) == "192.168.2.25"


def test_world_extrinsics_match_reference():
    g = np.load(GOLDEN / "calibration.npz")
    cal = RigCalibration(
        intrinsics={},
        extrinsics={n: [Extrinsics.from_4x4_matrix(g[f"cam_{n}_{i}"]) for i in range(k)] for n, k in (("a", 2), ("b", 1), ("c", 2))},
        rig_extrinsics={n: Extrinsics.from_4x4_matrix(g[f"rig_{n}"]) for n in ("a", "b")},
    )
    for n, k in (("a", 2), ("b", 1), ("c", 2)):
        for i, e in enumerate(cal.get_world_extrinsics(n)):
            assert np.array_equal(e.to_4x4_matrix(), g[f"world_{n}_{i}"])
    assert cal.get_world_extrinsics("zzz") is None


def test_context_manager_and_calibration_reload():
    srcs = make_sources()
    with CameraRig(srcs, queue_size=3) as rig:
        assert rig.is_running() and rig.get_source_names() == ["oak_b", "oak_a", "oak_c"]
        assert rig.get_source("oak_a") is srcs[1] and rig.get_source("none") is None
        assert np.array_equal(rig.get_rig_extrinsics("oak_a").to_4x4_matrix(), np.eye(4))  # identity default
        m = np.eye(4)
        m[:3, 3] = [1, 2, 3]
        rig.load_rig_extrinsics({"oak_a": Extrinsics.from_4x4_matrix(m)},
                                IMUExtrinsics("oak_b", Extrinsics.from_4x4_matrix(np.eye(4))))
        w = rig.get_world_extrinsics("oak_a")[0].to_4x4_matrix()
        assert np.allclose(w, m @ srcs[1].get_extrinsics()[0].to_4x4_matrix())
        for _ in range(5):
            rig.get_synchronized_frames()
        assert max(rig.get_queue_depths().values()) == 3  # bounded like deque(maxlen)
        rig.clear_queues()
        assert set(rig.get_queue_depths().values()) == {0}
    assert not rig.is_running()


def test_ring_behaves_like_bounded_deque():
    from collections import deque

    r, d = _Ring(4), deque(maxlen=4)
    rng = np.random.default_rng(0)
    for step in range(200):
        op = rng.integers(0, 10)
        if op < 7:
            r.push(step), d.append(step)
        elif op < 9 and d:
            assert r.pop_oldest() == d.popleft()
        else:
            r.clear(), d.clear()
        assert list(r) == list(d) and len(r) == len(d) and bool(r) == bool(d)
        if d:
            assert r[-1] == d[-1] and r[0] == d[0]
