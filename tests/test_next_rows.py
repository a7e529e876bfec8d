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


# ---- (f) row 1: rig description (YAML schema) -> sources + URDF poses -> ingest rig ------------------------------
def test_pipeline_config_matches_reference_from_dict():
    """Every field and default of scripts/run_pipeline.py:85-163 on the shipped YAML and the schema's corner cases."""
    import dataclasses

    from thor_slam_b200.ingest.pipeline_config import CAMERA_MAP, PipelineConfig

    g = json.loads((GOLDEN / "pipeline_config.json").read_text())
    assert CAMERA_MAP == g["camera_map_run_pipeline"] == g["camera_map_run_slam"]
    for name, case in g["cases"].items():
        want = dict(case["config"])
        default_urdf = want.pop("urdf_is_default_brackets")
        num = want.pop("num_cameras")
        cfg = PipelineConfig.from_dict(case["input"], default_urdf=__file__)  # any existing file plays the default URDF
        got = json.loads(json.dumps(dataclasses.asdict(cfg)))  # tuples -> lists, like the golden file
        if default_urdf:
            assert got["urdf_path"] == __file__, name
            got["urdf_path"] = "<default>"
        assert got == want, name
        assert cfg.calculate_num_cameras() == num, name


def test_rig_description_to_ingested_frames_emulated(emu_backend, tmp_path):
    """YAML + URDF -> IngestRig: poses come from the URDF (reference Euler order), frames come out rectified,
    the cloud is in the FLU body frame of that pose."""
    from oracle import backproject as ob
    from oracle import conventions as conv
    from oracle import rectify as orc
    from thor_slam_b200.ingest.pipeline_config import PipelineConfig, build_ingest_rig

    u = json.loads((GOLDEN / "urdf.json").read_text())
    joints = "".join(
        f'<joint name="{j["joint"]}" type="fixed"><parent link="base_link"/><child link="{j["link"]}"/>'
        f'<origin xyz="{j["xyz"]}" rpy="{j["rpy"]}"/></joint>' for j in u["joints"].values())
    urdf = tmp_path / "rig.urdf"
    urdf.write_text(f'<robot name="r"><link name="base_link"/>{joints}</robot>')
    yaml_path = tmp_path / "slam_config.yaml"
    yaml_path.write_text(
        "cameras:\n"
        '  - ip: "192.168.2.21"\n    stereo: true\n    resolution: [640, 400]\n    output_resolution: [192, 96]\n    sensor_type: "MONO"\n'
        '  - ip: "192.168.2.25"\n    stereo: true\n    resolution: [192, 96]\n    sensor_type: "MONO"\n    enable_rgbd: true\n'
        "    rgb_output_resolution: [128, 64]\n"
        f'urdf_path: "{urdf}"\nrig_queue_size: 3\nnvblox_cameras:\n  - "192.168.2.25"\n')
    cfg = PipelineConfig.from_yaml(yaml_path)
    assert cfg.calculate_num_cameras() == 4 and cfg.rig_queue_size == 3 and cfg.nvblox_cameras == ["192.168.2.25"]
    rig = build_ingest_rig(cfg, context=emu_backend.ctx)
    assert rig.queue_size == 3 and rig.get_source_names() == ["192.168.2.21", "192.168.2.25"]
    for ip in rig.get_source_names():
        assert np.allclose(rig.get_rig_extrinsics(ip).to_4x4_matrix(), np.array(u["matrices"][ip]), rtol=0, atol=1e-15)
    with rig:
        sync = rig.get_synchronized_frames(with_clouds=True)
    for ip in rig.get_source_names():
        src = rig.get_source(ip)
        (il, ir), (el, er) = src.get_intrinsics(), src.get_extrinsics()
        assert (il.width, il.height) == (192, 96)  # output_resolution wins over resolution for the SLAM streams
        r1, _r2, p1, _p2 = orc.stereo_rectify_cv(il.matrix, il.coeffs, ir.matrix, ir.coeffs, (192, 96), el.to_4x4_matrix(), er.to_4x4_matrix())
        mapx, mapy = orc.undistort_rectify_map_cv(il.matrix, il.coeffs, r1, p1, (192, 96))
        left = np.asarray(sync.frame_sets[ip].frames[0].image)
        assert np.array_equal(left, orc.remap_cv(src._pool[0][0], mapx, mapy))
    assert set(sync.clouds) == {"192.168.2.25"}  # only the nvblox camera delivers RGB-D
    src = rig.get_source("192.168.2.25")
    _ri, di = src.get_rgbd_intrinsics()
    m = conv.body_T_camera(np.array(u["matrices"]["192.168.2.25"]), src.get_rgbd_extrinsics()[1].to_4x4_matrix(), "rdf")
    pts, _mask, cnt = ob.backproject(src._rgbd_pool[0][1], di.matrix, m)
    c = sync.clouds["192.168.2.25"]
    assert ob.points_close(np.asarray(c["points"]), pts)[0] and int(c["count"]) == cnt


# ---- (f) row 3: depth -> RGB registration (one colour per depth pixel) -----------------------------------------------
def _registration_case(be, dw: int, dh: int, rw: int, rh: int, n: int = 2, seed: int = 21) -> None:
    from oracle import backproject as ob
    from thor_slam_b200.camera.synthetic import make_depth, make_image

    rng = np.random.default_rng(seed)
    src = SyntheticCameraSource(SyntheticCameraConfig(name="oak0", enable_rgbd=True, rgb_resolution=(rw, rh), depth_resolution=(dw, dh),
                                                      resolution=(dw, dh), pool=1, seed=seed))
    ri, di = src.get_rgbd_intrinsics()
    re, de = src.get_rgbd_extrinsics()  # both -> CAM_A (thor_slam/camera/drivers/luxonis.py:1068-1091)
    rgb_T_depth = np.linalg.inv(re.to_4x4_matrix()) @ de.to_4x4_matrix()
    be.ctx.upload_registration(50, di.matrix, (dw, dh), ri.matrix, (rw, rh), rgb_T_depth)
    depth = np.stack([make_depth(rng, dw, dh) for _ in range(n)])
    depth[0, :2, :] = 1  # 1 mm: far outside the RGB frustum or right at the camera - must not crash, colour (0,0,0) or a border pixel
    rgb = np.stack([make_image(rng, "bgr8", rw, rh)[..., ::-1].copy() for _ in range(n)])
    colour = be.zeros((n, dh, dw, 3), np.uint8)
    be.ctx.register_colour(50, be.dev(depth), be.dev(rgb), colour)
    got = be.host(colour)
    for i in range(n):
        want = ob.register_colour(depth[i], di.matrix, rgb_T_depth, ri.matrix, rgb[i])
        assert np.array_equal(got[i], want), f"frame {i}: {(got[i] != want).any(axis=-1).sum()} pixels differ"
        assert not got[i][depth[i] == 0].any(), "no depth, no colour"
    assert (got.reshape(-1, 3).any(axis=1)).mean() > 0.5, "most valid points of this rig land inside the RGB image"
    # a sensor pair looking in very different directions: most points fall behind the RGB camera or outside its image
    from scipy.spatial.transform import Rotation

    for k, yaw in enumerate((35.0, 100.0, 180.0)):
        skew = np.eye(4)
        skew[:3, :3] = Rotation.from_euler("y", yaw, degrees=True).as_matrix()
        skew[:3, 3] = (0.3, -0.1, 0.2)
        be.ctx.upload_registration(53, di.matrix, (dw, dh), ri.matrix, (rw, rh), skew)
        colour3 = be.zeros((1, dh, dw, 3), np.uint8)
        be.ctx.register_colour(53, be.dev(depth[:1]), be.dev(rgb[:1]), colour3)
        want = ob.register_colour(depth[0], di.matrix, skew, ri.matrix, rgb[0])
        assert np.array_equal(be.host(colour3)[0], want), f"yaw {yaw}"
        if yaw == 180.0:
            assert not want.any()
    # identical sensors (same K, same size, identity extrinsics): every valid pixel takes the colour at its own position
    be.ctx.upload_registration(51, di.matrix, (dw, dh), di.matrix, (dw, dh), np.eye(4))
    same = np.ascontiguousarray(rgb[:, :dh, :dw]) if (rh >= dh and rw >= dw) else make_image(rng, "bgr8", dw, dh)[None].repeat(n, 0)
    colour2 = be.zeros((n, dh, dw, 3), np.uint8)
    be.ctx.register_colour(51, be.dev(depth), be.dev(same), colour2)
    got2 = be.host(colour2)
    valid = depth > 0
    assert np.array_equal(got2[valid], same[valid]) and not got2[~valid].any()


def test_depth_rgb_registration_emulated(emu_backend):
    _registration_case(emu_backend, 96, 64, 128, 72)


@pytest.mark.gpu
def test_depth_rgb_registration_gpu(gpu_backend):
    _registration_case(gpu_backend, 1280, 800, 1920, 1080)
    _registration_case(gpu_backend, 640, 400, 640, 400, n=3, seed=22)
    with pytest.raises(RuntimeError):
        gpu_backend.ctx.register_colour(52, gpu_backend.zeros((1, 8, 8), np.uint16), gpu_backend.zeros((1, 8, 8, 3), np.uint8),
                                        gpu_backend.zeros((1, 8, 8, 3), np.uint8))  # slot never uploaded


# ---- (f) row 4: IMU samples in the convention of the IMU extrinsics -----------------------------------------------------
def test_imu_sample_rotation_matches_the_extrinsics_convention():
    from oracle import conventions as conv
    from thor_slam_b200.ingest.calib import rotate_imu_sample

    rng = np.random.default_rng(8)
    sample = {"accelerometer": [0.3, -9.7, 1.1], "gyroscope": [0.01, -0.02, 0.5], "timestamp": 12.5}
    assert rotate_imu_sample(sample, "rdf") is sample and rotate_imu_sample(None, "drb") is None  # reference behaviour: untouched
    got = rotate_imu_sample(sample, "drb")
    assert got["timestamp"] == 12.5 and sample["accelerometer"] == [0.3, -9.7, 1.1]  # pass-through, no mutation
    # x_rdf = y_drb, y_rdf = x_drb, z_rdf = -z_drb (scripts/run_slam.py:254-266)
    assert got["accelerometer"] == [-9.7, 0.3, -1.1] and got["gyroscope"] == [-0.02, 0.01, -0.5]
    # a Pro IMU read through (DRB extrinsics, raw sample) and through (RDF extrinsics, rotated sample) is the same world vector
    rig_pose, t_imu = np.eye(4), np.eye(4)
    from scipy.spatial.transform import Rotation

    rig_pose[:3, :3] = Rotation.from_rotvec(rng.uniform(-1, 1, 3)).as_matrix()
    t_imu[:3, :3] = Rotation.from_rotvec(rng.uniform(-0.2, 0.2, 3)).as_matrix()
    drb_world = conv.imu_world_extrinsics(rig_pose, t_imu, "drb")[:3, :3]
    for key in ("accelerometer", "gyroscope"):
        raw_in_imu_axes = np.linalg.inv(t_imu[:3, :3]) @ conv.DRB_TO_RDF[:3, :3].T @ np.asarray(got[key])  # undo: back to what the chip reported
        np.testing.assert_allclose(raw_in_imu_axes, np.linalg.inv(t_imu[:3, :3]) @ np.asarray(sample[key]), atol=1e-12)
        np.testing.assert_allclose(drb_world @ np.linalg.inv(t_imu[:3, :3]) @ np.asarray(sample[key]),
                                   rig_pose[:3, :3] @ conv.DRB_TO_RDF[:3, :3] @ np.asarray(sample[key]), atol=1e-12)
    assert abs(np.linalg.det(conv.DRB_TO_RDF[:3, :3]) - 1.0) < 1e-15  # proper rotation: the gyro transforms like the accelerometer
    with pytest.raises(ValueError):
        rotate_imu_sample(sample, "ulb")


def test_ingest_rig_rotates_pro_imu_samples(emu_backend):
    from thor_slam_b200.ingest.rig import IngestRig

    def make(name):
        return SyntheticCameraSource(SyntheticCameraConfig(name=name, resolution=(192, 96), pool=1, read_imu=True, seed=3))

    rig_drb = IngestRig([make("pro")], queue_size=2, imu_source="pro", imu_frame="drb", context=emu_backend.ctx)
    rig_rdf = IngestRig([make("pro")], queue_size=2, imu_source="pro", context=emu_backend.ctx)
    with rig_drb, rig_rdf:
        a, b = rig_drb.get_synchronized_frames(), rig_rdf.get_synchronized_frames()
    assert a.sensor_data is not None and b.sensor_data is not None and a.sensor_timestamp == b.sensor_timestamp
    for key in ("accelerometer", "gyroscope"):
        x, y, z = b.sensor_data[key][:3]
        assert a.sensor_data[key][:3] == [y, x, -z]
    with pytest.raises(ValueError):
        IngestRig([make("pro")], imu_source="pro", imu_frame="flu", context=emu_backend.ctx)


class _RecordedCalib:
    """The fake ``CalibrationHandler`` of ``tests/golden/make_golden.py``, rebuilt from what the golden file recorded."""

    def __init__(self, g: dict) -> None:
        self.g = g

    def getCameraIntrinsics(self, sock, w, h):
        k = np.array(self.g["k_native"][sock])
        nw, nh = self.g["native_resolution"][sock]
        k[0] *= w / nw
        k[1] *= h / nh
        return k.tolist()

    def getDistortionCoefficients(self, sock):
        return self.g["distortion"][sock]

    def getCameraExtrinsics(self, src, dst):
        assert dst == "CAM_A"
        return self.g["to_cam_a_cm"][src]

    def getImuToCameraExtrinsics(self, sock):
        return self.g["imu_to_cam_a_cm"]


def test_luxonis_calibration_conventions_match_the_driver():
    """sensor -> output scaling, cm -> m, CAM_A as the reference, RGB-D intrinsics rules - against the
    outputs of the reference driver's own getters (``tests/golden/luxonis_calibration.json``)."""
    import json

    from thor_slam_b200.camera import luxonis_calib as lc

    g = json.loads((GOLDEN / "luxonis_calibration.json").read_text())
    calib = _RecordedCalib(g)
    mono = g["mono_sensor_resolution"]

    def same_intr(got, want) -> None:
        assert (got.width, got.height) == (want["width"], want["height"])
        assert np.array_equal(got.matrix, np.array(want["matrix"])) and np.array_equal(got.coeffs, np.array(want["coeffs"]))

    for tag in ("stereo_native", "stereo_half", "stereo_anisotropic"):
        c = g["cases"][tag]
        for got, want in zip(lc.slam_intrinsics(calib, True, mono, c["output_resolution"]), c["intrinsics"], strict=True):
            same_intr(got, want)
        for got, want in zip(lc.slam_extrinsics(calib, True), c["extrinsics"], strict=True):
            assert np.array_equal(got.to_4x4_matrix(), np.array(want))
        assert np.array_equal(lc.sensor_extrinsics(calib).to_4x4_matrix(), np.array(c["sensor_extrinsics"]))
    c = g["cases"]["single"]
    same_intr(lc.slam_intrinsics(calib, False, mono, c["output_resolution"])[0], c["intrinsics"][0])
    assert np.array_equal(lc.slam_extrinsics(calib, False)[0].to_4x4_matrix(), np.eye(4)) and c["extrinsics"] == [np.eye(4).tolist()]
    for tag in ("rgbd_aligned", "rgbd_not_aligned", "rgbd_aligned_mismatched"):
        c = g["cases"][tag]
        ri, di = lc.rgbd_intrinsics(calib, c["rgb_sensor_resolution"], c["rgb_output_resolution"], c["depth_output_resolution"], mono, c["depth_align_to_rgb"])
        same_intr(ri, c["rgb_intrinsics"])
        same_intr(di, c["depth_intrinsics"])
        re, de = lc.rgbd_extrinsics(calib)
        assert np.array_equal(re.to_4x4_matrix(), np.array(c["rgb_extrinsics"])) and np.array_equal(de.to_4x4_matrix(), np.array(c["depth_extrinsics"]))
    # the conventions themselves, spelled out: 3.75 cm of baseline is 0.0375 m, left is at -x of CAM_A
    left = lc.slam_extrinsics(calib, True)[0]
    assert left.translation[0] == g["to_cam_a_cm"]["CAM_B"][0][3] / 100.0 < 0
    half = lc.slam_intrinsics(calib, True, mono, (640, 400))[0].matrix
    full = lc.slam_intrinsics(calib, True, mono, (1280, 800))[0].matrix
    assert np.allclose(half[:2], full[:2] / 2)

    class _NoImu(_RecordedCalib):
        def getImuToCameraExtrinsics(self, sock):
            raise RuntimeError("no IMU extrinsics")

    assert np.array_equal(lc.sensor_extrinsics(_NoImu(g)).to_4x4_matrix(), np.eye(4))
