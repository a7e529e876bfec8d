"""SURVEY section 8(f) rows built so far, against golden vectors recorded from the reference."""

from __future__ import annotations

import json

import numpy as np
import pytest

from tests.conftest import GOLDEN
from thor_slam_b200.camera import Extrinsics
from thor_slam_b200.camera.rig import CameraRig
from thor_slam_b200.camera.synthetic import SyntheticCameraConfig, SyntheticCameraSource
from thor_slam_b200.camera.utils import load_rig_extrinsics_from_urdf, parse_urdf_transform, rpy_to_matrix
from thor_slam_b200.slam import extract_cameras
from thor_slam_b200.slam.camera_info import camera_info_raw, camera_info_rectified, static_transforms


def test_urdf_rig_extrinsics_match_reference(tmp_path):
    u = json.loads((GOLDEN / "urdf.json").read_text())
    joints = "".join(
        f'<joint name="{j["joint"]}" type="fixed"><parent link="base_link"/><child link="{j["link"]}"/>'
        f'<origin xyz="{j["xyz"]}" rpy="{j["rpy"]}"/></joint>' for j in u["joints"].values())
    other = '<joint name="x" type="fixed"><parent link="not_base"/><child link="link_Camera_1_centroid"/><origin xyz="9 9 9" rpy="0 0 0"/></joint>'
    path = tmp_path / "rig.urdf"
    path.write_text(f'<robot name="r"><link name="base_link"/>{other}{joints}</robot>')
    got = load_rig_extrinsics_from_urdf(path, {**u["camera_map"], "ghost": "no_such_link"})
    assert set(got) == set(u["matrices"])  # the unknown link only warns
    for src, e in got.items():
        assert np.allclose(e.to_4x4_matrix(), np.array(u["matrices"][src]), rtol=0, atol=1e-15)
    with pytest.raises(FileNotFoundError):
        load_rig_extrinsics_from_urdf(tmp_path / "nope.urdf", {})


def test_author_requested_transform_check():
    """thor_slam/camera/utils.py:99-100: '1 m in x, 0.5 m in y, 0.25 m in z ... roll pitch yaw'."""
    import xml.etree.ElementTree as ET

    u = json.loads((GOLDEN / "urdf.json").read_text())["author_case"]
    j = ET.fromstring(f'<joint name="t" type="fixed"><origin xyz="{u["xyz"]}" rpy="{u["rpy"]}"/></joint>')
    m = parse_urdf_transform(j)
    assert np.allclose(m, np.array(u["matrix"]), rtol=0, atol=1e-15)
    assert np.allclose(m[:3, 3], [1.0, 0.5, 0.25])
    assert np.array_equal(parse_urdf_transform(ET.fromstring('<joint name="t" type="fixed"/>')), np.eye(4))
    # the URDF-standard order is a different matrix unless two of the angles vanish
    assert not np.allclose(rpy_to_matrix([0.1, -0.2, 0.3], "urdf"), rpy_to_matrix([0.1, -0.2, 0.3], "reference"))
    assert np.allclose(rpy_to_matrix([0.4, 0, 0], "urdf"), rpy_to_matrix([0.4, 0, 0], "reference"))


def test_camera_info_and_tf_match_reference_adapter():
    meta = json.loads((GOLDEN / "isaac_adapter.json").read_text())
    g = np.load(GOLDEN / "isaac_adapter.npz")
    a = SyntheticCameraSource(SyntheticCameraConfig(name="192.168.2.25", resolution=(96, 64), pixel_format="mono8", seed=5, pool=1))
    b = SyntheticCameraSource(SyntheticCameraConfig(name="192.168.2.21", resolution=(96, 64), pixel_format="bgr8", seed=6, pool=1, distortion="plumb_bob5"))
    rig = CameraRig([a, b], queue_size=2, rig_extrinsics={s.name: Extrinsics.from_4x4_matrix(g[f"rig_ext_{s.name}"]) for s in (a, b)})
    cams = extract_cameras(rig.calibration, 4)
    assert [(c.source_name, c.cam_idx) for c in cams] == [(c["source_name"], c["cam_idx"]) for c in meta["cameras"]]
    for i, cam in enumerate(cams):
        left = cams[i - 1] if cam.cam_idx == 1 and i > 0 and cams[i - 1].source_name == cam.source_name else None
        info = camera_info_raw(cam, left)
        ref = meta["infos"][i]
        assert (info.width, info.height, info.distortion_model) == (ref["width"], ref["height"], ref["distortion_model"])
        assert info.d == ref["d"] and info.k == ref["k"] and info.r == ref["r"]
        assert np.allclose(info.p, ref["p"], rtol=1e-13, atol=1e-15)
    tfs = static_transforms(cams, imu_extrinsics=rig.calibration.imu_extrinsics)
    assert len(tfs) == len(meta["tf"])
    for t, ref in zip(tfs, meta["tf"]):
        assert (t.parent, t.child) == (ref["parent"], ref["child"])
        assert np.allclose(t.translation, ref["t"], atol=1e-15) and np.allclose(t.rotation_xyzw, ref["q"], atol=1e-12)
    std = static_transforms(cams, ros_standard_optical=True)
    assert np.allclose(np.abs(std[1].rotation_xyzw), [0.5, 0.5, 0.5, 0.5])


def test_rectified_camera_info():
    from thor_slam_b200.ingest.calib import stereo_rectification

    s = SyntheticCameraSource(SyntheticCameraConfig(name="o", resolution=(160, 100), seed=21, pool=1))
    (il, ir), (el, er) = s.get_intrinsics(), s.get_extrinsics()
    r1, r2, p1, p2 = stereo_rectification(il, ir, el, er, (160, 100))
    li, ri = camera_info_rectified(160, 100, r1, p1), camera_info_rectified(160, 100, r2, p2)
    assert li.d == [0.0] * 5 and li.k == np.asarray(p1)[:, :3].flatten().tolist()
    assert li.p[3] == 0.0 and ri.p[3] < 0  # ROS stereo convention: Tx = -fx * baseline on the right camera
    assert np.isclose(ri.p[3], -p2[0, 0] * 0.075, rtol=1e-3)
