"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run only in the build container (``/root/reference`` is not on the GPU box):

    python tests/golden/make_golden.py

The reference's pure-Python modules are imported and executed for real; the
packages it needs but this image lacks (ROS 2: ``rclpy``, ``cv_bridge``,
``tf2_ros``, message packages; ``depthai``) are replaced by recording stubs, so
what lands in the fixtures is exactly what the reference would have published:

* ``rig_sync.json``        - ``thor_slam.camera.rig.CameraRig`` frame-set selection traces
* ``calibration.npz``      - ``Extrinsics`` round trips + ``RigCalibration.get_world_extrinsics``
* ``isaac_adapter.npz/json`` - ``IsaacRosAdapter``: stream order, mono8 / BGR2RGB images,
                             CameraInfo K/D/R/P, static TFs, ``RDF_TO_FLU_MATRIX``
* ``rgbd_publisher.json``  - ``scripts.run_pipeline.RGBDPublisher`` CameraInfo + encodings
* ``urdf.json``            - ``load_rig_extrinsics_from_urdf`` on ``examples/assets/brackets.urdf``
* ``cv_arith.npz``         - OpenCV 4.13 outputs (cvtColor / initUndistortRectifyMap / remap) on
                             small seeded inputs, frozen so the GPU box checks against bytes
                             produced here rather than against its own OpenCV build.
"""

from __future__ import annotations

import json
import sys
import types
import xml.etree.ElementTree as ET
from pathlib import Path
from unittest import mock

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REF))


# ----------------------------------------------------------------------------------------------
# recording stubs for ROS 2 / depthai
# ----------------------------------------------------------------------------------------------
class _Msg:
    """Attribute bag: ``msg.header.stamp.sec = 1`` just works; lists are preallocated like rosidl."""

    def __init__(self) -> None:
        object.__setattr__(self, "_d", {})

    def __getattr__(self, name: str):
        d = object.__getattribute__(self, "_d")
        if name not in d:
            d[name] = [0.0] * 36 if name.endswith("covariance") else _Msg()
        return d[name]

    def __setattr__(self, name: str, value) -> None:
        object.__getattribute__(self, "_d")[name] = value

    def to_plain(self):
        out = {}
        for k, v in object.__getattribute__(self, "_d").items():
            out[k] = v.to_plain() if isinstance(v, _Msg) else v
        return out


PUBLISHED: dict[str, list] = {}
TRANSFORMS: list = []


class _Publisher:
    def __init__(self, topic: str) -> None:
        self.topic = topic
        PUBLISHED.setdefault(topic, [])

    def publish(self, msg) -> None:
        PUBLISHED[self.topic].append(msg)


class _Clock:
    def now(self):
        m = mock.MagicMock()
        m.to_msg.return_value = "t0"
        return m


class _Node:
    def __init__(self, name: str = "node", *a, **k) -> None:
        self.name = name

    def create_publisher(self, _type, topic, _qos):
        return _Publisher(topic)

    def create_subscription(self, *a, **k):
        return None

    def get_clock(self):
        return _Clock()

    def get_logger(self):
        return mock.MagicMock()

    def destroy_node(self):
        pass


class _Bridge:
    def cv2_to_imgmsg(self, img, encoding="passthrough"):
        m = _Msg()
        m.encoding = encoding
        m.image = np.array(img, copy=True)
        return m


class _TfBroadcaster:
    def __init__(self, node) -> None:
        pass

    def sendTransform(self, transforms) -> None:  # noqa: N802 (ROS name)
        TRANSFORMS.extend(transforms)


def _install_stubs() -> None:
    def mod(name: str, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    rclpy = mod("rclpy", ok=lambda: True, init=lambda *a, **k: None, spin=lambda node: None, shutdown=lambda: None)
    rclpy.node = mod("rclpy.node", Node=_Node)
    rclpy.publisher = mod("rclpy.publisher", Publisher=_Publisher)
    rclpy.qos = mod(
        "rclpy.qos",
        qos_profile_sensor_data="sensor_qos",
        QoSProfile=lambda **k: k,
        ReliabilityPolicy=mock.MagicMock(),
        DurabilityPolicy=mock.MagicMock(),
        HistoryPolicy=mock.MagicMock(),
    )
    mod("builtin_interfaces"), mod("builtin_interfaces.msg", Time=_Msg)
    mod("cv_bridge", CvBridge=_Bridge)
    mod("geometry_msgs"), mod("geometry_msgs.msg", TransformStamped=_Msg)
    mod("nav_msgs"), mod("nav_msgs.msg", Odometry=_Msg)
    mod("sensor_msgs"), mod("sensor_msgs.msg", CameraInfo=_Msg, Image=_Msg, Imu=_Msg)
    mod("tf2_ros", StaticTransformBroadcaster=_TfBroadcaster)
    sys.modules["depthai"] = mock.MagicMock(name="depthai")
    for extra in ("colorlogging", "askin"):
        sys.modules[extra] = mock.MagicMock(name=extra)


_install_stubs()

import cv2  # noqa: E402

from thor_slam.camera import rig as ref_rig  # noqa: E402
from thor_slam.camera import types as ref_types  # noqa: E402

from thor_slam_b200.camera.synthetic import (  # noqa: E402
    SyntheticCameraConfig,
    SyntheticCameraSource,
    make_depth,
    make_image,
)


def ref_source(inner) -> "ref_types.CameraSource":
    """Wrap one of OUR synthetic sources in a genuine reference ``CameraSource`` subclass."""

    class Wrapped(ref_types.CameraSource):
        @property
        def name(self):
            return inner.name

        def start(self):
            inner.start()

        def stop(self):
            inner.stop()

        def get_latest_frames(self):
            return [ref_types.CameraFrame(f.image, f.timestamp, f.sequence_num, f.camera_name) for f in inner.get_latest_frames()]

        def try_get_latest_frames(self):
            fr = inner.try_get_latest_frames()
            return None if fr is None else [ref_types.CameraFrame(f.image, f.timestamp, f.sequence_num, f.camera_name) for f in fr]

        def get_intrinsics(self):
            return [ref_types.Intrinsics(i.width, i.height, i.matrix, i.coeffs) for i in inner.get_intrinsics()]

        def get_extrinsics(self):
            return [ref_types.Extrinsics(e.rotation, e.translation) for e in inner.get_extrinsics()]

        def get_sensor_extrinsics(self):
            e = inner.get_sensor_extrinsics()
            return None if e is None else ref_types.Extrinsics(e.rotation, e.translation)

        def get_timestamped_sensor_data(self):
            return inner.get_timestamped_sensor_data()

        @property
        def has_sensor_data(self):
            return inner.has_sensor_data

    return Wrapped()


# ----------------------------------------------------------------------------------------------
# scenario shared with tests/test_rig_parity.py
# ----------------------------------------------------------------------------------------------
SYNC_SCENARIO = [
    # name, fps, time_offset, stereo, read_imu
    ("oak_b", 30.0, 0.0031, True, True),
    ("oak_a", 30.0, 0.0007, True, False),
    ("oak_c", 20.0, 0.0190, False, False),
]
SYNC_QUEUE = 5
SYNC_STEPS = 24


def scenario_sources():
    out = []
    for i, (name, fps, off, stereo, imu) in enumerate(SYNC_SCENARIO):
        out.append(
            SyntheticCameraSource(
                SyntheticCameraConfig(
                    name=name, fps=fps, time_offset=off, stereo=stereo, read_imu=imu, resolution=(64, 40),
                    pixel_format="mono8" if stereo else "bgr8", seed=11 + i, pool=2,
                )
            )
        )
    return out


def _trace_entry(sync) -> dict | None:
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


def gen_rig_sync() -> None:
    srcs = [ref_source(s) for s in scenario_sources()]
    rig = ref_rig.CameraRig(srcs, queue_size=SYNC_QUEUE, imu_source="oak_b")
    trace = {"before_start": _trace_entry(rig.get_synchronized_frames()), "sync": [], "latest": [], "depths": []}
    rig.start()
    for step in range(SYNC_STEPS):
        trace["sync"].append(_trace_entry(rig.get_synchronized_frames()))
        trace["depths"].append(rig.get_queue_depths())
        if step % 6 == 5:
            trace["latest"].append(_trace_entry(rig.get_latest_frames()))
    trace["pruned"] = rig.prune_old_frames(0.05)
    trace["depths_after_prune"] = rig.get_queue_depths()
    rig.stop()
    trace["depths_after_stop"] = rig.get_queue_depths()
    trace["after_stop"] = _trace_entry(rig.get_synchronized_frames())
    # error behaviour
    errs = {}
    try:
        ref_rig.CameraRig(srcs, imu_source="nope")
    except ValueError as e:
        errs["imu_unknown"] = type(e).__name__
    try:
        ref_rig.CameraRig(srcs, imu_source="oak_a")
    except ValueError as e:
        errs["imu_no_data"] = type(e).__name__
    try:
        rig.load_rig_extrinsics({"ghost": ref_types.Extrinsics(np.eye(3), np.zeros(3))})
    except ValueError as e:
        errs["unknown_source"] = type(e).__name__
    try:
        ref_types.Extrinsics.from_4x4_matrix(np.eye(3))
    except ValueError as e:
        errs["bad_4x4"] = type(e).__name__
    try:
        ref_types.FrameSet.from_frames([], "x")
    except ValueError as e:
        errs["empty_frameset"] = type(e).__name__
    trace["errors"] = errs
    (HERE / "rig_sync.json").write_text(json.dumps(trace, indent=1))


def gen_calibration() -> None:
    rng = np.random.default_rng(1337)
    from scipy.spatial.transform import Rotation

    def rand_T():
        m = np.eye(4)
        m[:3, :3] = Rotation.from_rotvec(rng.uniform(-1, 1, 3)).as_matrix()
        m[:3, 3] = rng.uniform(-0.5, 0.5, 3)
        return m

    rig_T = {"a": rand_T(), "b": rand_T()}
    cam_T = {"a": [rand_T(), rand_T()], "b": [rand_T()], "c": [rand_T(), rand_T()]}
    cal = ref_rig.RigCalibration(
        intrinsics={},
        extrinsics={n: [ref_types.Extrinsics.from_4x4_matrix(t) for t in ts] for n, ts in cam_T.items()},
        rig_extrinsics={n: ref_types.Extrinsics.from_4x4_matrix(t) for n, t in rig_T.items()},
    )
    out = {}
    for n in cam_T:
        for i, e in enumerate(cal.get_world_extrinsics(n)):
            out[f"world_{n}_{i}"] = e.to_4x4_matrix()
    for n, t in rig_T.items():
        out[f"rig_{n}"] = t
    for n, ts in cam_T.items():
        for i, t in enumerate(ts):
            out[f"cam_{n}_{i}"] = t
    out["unknown_is_none"] = np.array(cal.get_world_extrinsics("zzz") is None)
    np.savez(HERE / "calibration.npz", **out)


def _stereo_rig_sources():
    a = SyntheticCameraSource(SyntheticCameraConfig(name="192.168.2.25", resolution=(96, 64), pixel_format="mono8", seed=5, pool=1))
    b = SyntheticCameraSource(SyntheticCameraConfig(name="192.168.2.21", resolution=(96, 64), pixel_format="bgr8", seed=6, pool=1, distortion="plumb_bob5"))
    return [a, b]


def gen_isaac_adapter() -> None:
    from thor_slam.slam.adapters import isaac_ros

    PUBLISHED.clear()
    TRANSFORMS.clear()
    inner = _stereo_rig_sources()
    rng = np.random.default_rng(3)
    from scipy.spatial.transform import Rotation

    rig_ext = {}
    for s in inner:
        m = np.eye(4)
        m[:3, :3] = Rotation.from_rotvec(rng.uniform(-1, 1, 3)).as_matrix()
        m[:3, 3] = rng.uniform(-0.3, 0.3, 3)
        rig_ext[s.name] = ref_types.Extrinsics.from_4x4_matrix(m)
    rig = ref_rig.CameraRig([ref_source(s) for s in inner], queue_size=4, rig_extrinsics=rig_ext)
    adapter = isaac_ros.IsaacRosAdapter(num_cameras=4)
    adapter.initialize(rig.calibration)
    rig.start()
    sync = rig.get_synchronized_frames()
    adapter.process_frames(sync)
    rig.stop()

    arrays = {"RDF_TO_FLU_MATRIX": np.array(isaac_ros.RDF_TO_FLU_MATRIX)}
    arrays["readme_known_answer"] = isaac_ros.RDF_TO_FLU_MATRIX @ np.array([1, 0, 0, 1])
    meta: dict = {"cameras": [], "images": [], "infos": [], "tf": []}
    for i, cam in enumerate(adapter._cameras):
        meta["cameras"].append({"source_name": cam.source_name, "cam_idx": cam.cam_idx})
        arrays[f"cam_ext_{i}"] = cam.extrinsics.to_4x4_matrix()
    for i in range(4):
        msg = PUBLISHED[f"/visual_slam/image_{i}"][0]
        arrays[f"image_{i}"] = msg.image
        meta["images"].append({"encoding": msg.encoding, "frame_id": msg.header.frame_id,
                               "sec": msg.header.stamp.sec, "nanosec": msg.header.stamp.nanosec})
        info = PUBLISHED[f"/visual_slam/camera_info_{i}"][0]
        meta["infos"].append({"width": info.width, "height": info.height, "distortion_model": info.distortion_model,
                              "d": list(info.d), "k": list(info.k), "r": list(info.r), "p": list(info.p)})
    for t in TRANSFORMS:
        meta["tf"].append({
            "parent": t.header.frame_id, "child": t.child_frame_id,
            "t": [t.transform.translation.x, t.transform.translation.y, t.transform.translation.z],
            "q": [t.transform.rotation.x, t.transform.rotation.y, t.transform.rotation.z, t.transform.rotation.w],
        })
    # inputs needed to replay the case without the reference
    for s in inner:
        arrays[f"rig_ext_{s.name}"] = rig_ext[s.name].to_4x4_matrix()
    meta["sources"] = [s.name for s in inner]
    meta["sync_timestamp"] = sync.timestamp
    np.savez_compressed(HERE / "isaac_adapter.npz", **arrays)
    (HERE / "isaac_adapter.json").write_text(json.dumps(meta, indent=1))


def gen_rgbd_publisher() -> None:
    from scripts import run_pipeline

    PUBLISHED.clear()
    src = SyntheticCameraSource(
        SyntheticCameraConfig(name="192.168.2.25", resolution=(96, 64), enable_rgbd=True, rgb_resolution=(128, 72),
                              depth_resolution=(96, 64), seed=9, pool=1)
    )
    src.start()
    cam = mock.MagicMock()
    ri, di = src.get_rgbd_intrinsics()
    cam.get_rgbd_intrinsics.return_value = (
        ref_types.Intrinsics(ri.width, ri.height, ri.matrix, ri.coeffs),
        ref_types.Intrinsics(di.width, di.height, di.matrix, di.coeffs),
    )
    pub = run_pipeline.RGBDPublisher(cam, 2)
    rgb, depth = src.get_latest_rgbd_frames()
    pub.publish_rgbd(rgb, depth)
    ns = "/camera_2"
    rgb_msg = PUBLISHED[f"{ns}/rgb/image_raw"][0]
    depth_msg = PUBLISHED[f"{ns}/depth/image_raw"][0]
    meta = {
        "rgb_encoding": rgb_msg.encoding, "depth_encoding": depth_msg.encoding,
        "frame_id": rgb_msg.header.frame_id,
        "rgb_is_bgr_reversed": bool((rgb_msg.image == rgb.image[..., ::-1]).all()),
        "depth_unchanged": bool((depth_msg.image == depth.image).all()),
        "depth_dtype": str(depth_msg.image.dtype),
    }
    for key, topic in (("rgb_info", f"{ns}/rgb/camera_info"), ("depth_info", f"{ns}/depth/camera_info")):
        info = PUBLISHED[topic][0]
        meta[key] = {"width": info.width, "height": info.height, "distortion_model": info.distortion_model,
                     "d": list(info.d), "k": list(info.k), "r": list(info.r), "p": list(info.p)}
    meta["rgb_K"] = ri.matrix.tolist()
    meta["rgb_coeffs"] = ri.coeffs.tolist()
    meta["depth_K"] = di.matrix.tolist()
    meta["depth_coeffs"] = di.coeffs.tolist()
    (HERE / "rgbd_publisher.json").write_text(json.dumps(meta, indent=1))


class _FakeCalib:
    """What ``dai.CalibrationHandler`` answers, with the conventions the driver relies on: intrinsics are given at the
    resolution asked for (DepthAI scales the stored matrix), distortion has 14 coefficients, extrinsics are 4x4 with the
    translation in CENTIMETRES (``drivers/luxonis.py:675-726``)."""

    def __init__(self, sockets: dict, k_native: dict, native: dict, dist: dict, to_a_cm: dict, imu_to_a_cm: np.ndarray) -> None:
        self.sockets, self.k_native, self.native, self.dist, self.to_a_cm, self.imu_to_a_cm = sockets, k_native, native, dist, to_a_cm, imu_to_a_cm
        self.calls: list = []

    def _name(self, sock) -> str:
        return next(n for n, s in self.sockets.items() if s is sock)

    def getCameraIntrinsics(self, sock, w, h):
        n = self._name(sock)
        self.calls.append(("intrinsics", n, int(w), int(h)))
        k = self.k_native[n].copy()
        nw, nh = self.native[n]
        k[0] *= w / nw
        k[1] *= h / nh
        return k.tolist()

    def getDistortionCoefficients(self, sock):
        return self.dist[self._name(sock)].tolist()

    def getCameraExtrinsics(self, src, dst):
        a, b = self._name(src), self._name(dst)
        assert b == "CAM_A", "the driver only ever asks for X -> CAM_A"
        self.calls.append(("extrinsics", a, b))
        return self.to_a_cm[a].tolist()

    def getImuToCameraExtrinsics(self, sock):
        assert self._name(sock) == "CAM_A"
        return self.imu_to_a_cm.tolist()


def gen_luxonis_calibration() -> None:
    """Rows a4 / a5 of SURVEY section 8: the driver's own calibration getters (``drivers/luxonis.py:596-726,974-1091``) run on
    a fake ``_calib_data`` - sensor-resolution K scaled to the output, centimetres to metres, which socket is the reference."""
    import depthai as dai  # the MagicMock installed above: sockets are stable attribute objects
    from thor_slam.camera.drivers import luxonis as lux

    rng = np.random.default_rng(77)
    sockets = {"CAM_A": dai.CameraBoardSocket.CAM_A, "CAM_B": dai.CameraBoardSocket.CAM_B, "CAM_C": dai.CameraBoardSocket.CAM_C}
    native = {"CAM_A": (1920, 1200), "CAM_B": (1280, 800), "CAM_C": (1280, 800)}
    k_native, dist, to_a = {}, {}, {}
    for n, (w, h) in native.items():
        f = 0.62 * w * (1 + rng.uniform(-0.01, 0.01))
        k_native[n] = np.array([[f, 0, w / 2 + rng.uniform(-8, 8)], [0, f * (1 + rng.uniform(-0.002, 0.002)), h / 2 + rng.uniform(-8, 8)], [0, 0, 1.0]])
        dist[n] = rng.uniform(-0.05, 0.05, size=14)
    from scipy.spatial.transform import Rotation

    for n, tx in (("CAM_B", -3.75), ("CAM_C", 3.75)):
        m = np.eye(4)
        m[:3, :3] = Rotation.from_rotvec(rng.uniform(-0.01, 0.01, 3)).as_matrix()
        m[:3, 3] = [tx, rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05)]  # centimetres
        to_a[n] = m
    imu = np.eye(4)
    imu[:3, :3] = Rotation.from_rotvec(rng.uniform(-0.2, 0.2, 3)).as_matrix()
    imu[:3, 3] = [0.7, -1.3, 0.4]  # centimetres
    calib = _FakeCalib(sockets, k_native, native, dist, to_a, imu)

    def source(stereo: bool, out_res, rgbd) -> "lux.LuxonisCameraSource":
        s = lux.LuxonisCameraSource.__new__(lux.LuxonisCameraSource)  # no device: the getters only need cfg + _calib_data
        s.cfg = lux.LuxonisCameraConfig(ip="192.168.2.21", fps=30, stereo=stereo, mono_sensor_resolution=lux.LuxonisResolution(1280, 800),
                                        output_resolution=lux.LuxonisResolution(*out_res), rgbd_camera_config=rgbd)
        s._calib_data, s._intrinsics, s._extrinsics = calib, None, None
        return s

    def intr(i) -> dict:
        return {"width": i.width, "height": i.height, "matrix": np.asarray(i.matrix).tolist(), "coeffs": np.asarray(i.coeffs).tolist()}

    cases = {}
    for tag, out_res in (("stereo_native", (1280, 800)), ("stereo_half", (640, 400)), ("stereo_anisotropic", (640, 480))):
        s = source(True, out_res, None)
        cases[tag] = {"output_resolution": list(out_res), "intrinsics": [intr(i) for i in s.get_intrinsics()],
                      "extrinsics": [e.to_4x4_matrix().tolist() for e in s.get_extrinsics()],
                      "sensor_extrinsics": s.get_sensor_extrinsics().to_4x4_matrix().tolist()}
    s = source(False, (1280, 800), None)
    cases["single"] = {"output_resolution": [1280, 800], "intrinsics": [intr(i) for i in s.get_intrinsics()],
                       "extrinsics": [e.to_4x4_matrix().tolist() for e in s.get_extrinsics()]}
    for tag, aligned, rgb_out, depth_out in (("rgbd_aligned", True, (1280, 800), (1280, 800)), ("rgbd_not_aligned", False, (1280, 720), (640, 400)),
                                              ("rgbd_aligned_mismatched", True, (1280, 800), (640, 400))):
        rgbd = lux.LuxonisRGBDCameraConfig(enable_rgbd=True, rgb_sensor_resolution=lux.LuxonisResolution(1920, 1200),
                                           rgb_output_resolution=lux.LuxonisResolution(*rgb_out),
                                           depth_output_resolution=lux.LuxonisResolution(*depth_out), depth_align_to_rgb=aligned)
        s = source(True, (640, 400), rgbd)
        ri, di = s.get_rgbd_intrinsics()
        re, de = s.get_rgbd_extrinsics()
        cases[tag] = {"depth_align_to_rgb": aligned, "rgb_sensor_resolution": [1920, 1200], "rgb_output_resolution": list(rgb_out),
                      "depth_output_resolution": list(depth_out), "rgb_intrinsics": intr(ri), "depth_intrinsics": intr(di),
                      "rgb_extrinsics": re.to_4x4_matrix().tolist(), "depth_extrinsics": de.to_4x4_matrix().tolist()}
    meta = {"what": "outputs of thor_slam.camera.drivers.luxonis.LuxonisCameraSource calibration getters on a fake CalibrationHandler",
            "native_resolution": {k: list(v) for k, v in native.items()}, "k_native": {k: v.tolist() for k, v in k_native.items()},
            "distortion": {k: v.tolist() for k, v in dist.items()}, "to_cam_a_cm": {k: v.tolist() for k, v in to_a.items()},
            "imu_to_cam_a_cm": imu.tolist(), "mono_sensor_resolution": [1280, 800], "cases": cases}
    (HERE / "luxonis_calibration.json").write_text(json.dumps(meta, indent=1))


def gen_urdf() -> None:
    from scripts import run_slam
    from thor_slam.camera import utils as ref_utils

    urdf = REF / "examples" / "assets" / "brackets.urdf"
    ext = ref_utils.load_rig_extrinsics_from_urdf(urdf, run_slam.CAMERA_MAP)
    root = ET.parse(urdf).getroot()
    joints = {}
    for source, link in run_slam.CAMERA_MAP.items():
        for j in root.findall("joint"):
            c, p = j.find("child"), j.find("parent")
            if c is not None and c.get("link") == link and p is not None and p.get("link") == "base_link":
                o = j.find("origin")
                joints[source] = {"link": link, "joint": j.get("name"), "xyz": o.get("xyz"), "rpy": o.get("rpy")}
                break
    out = {
        "camera_map": run_slam.CAMERA_MAP,
        "joints": joints,
        "matrices": {s: e.to_4x4_matrix().tolist() for s, e in ext.items()},
    }
    # the author's requested check (utils.py:99-100): 1 m x, 0.5 m y, 0.25 m z + roll/pitch/yaw
    j = ET.fromstring('<joint name="t" type="fixed"><origin xyz="1 0.5 0.25" rpy="0.1 -0.2 0.3"/></joint>')
    out["author_case"] = {"xyz": "1 0.5 0.25", "rpy": "0.1 -0.2 0.3", "matrix": ref_utils.parse_urdf_transform(j).tolist()}
    j2 = ET.fromstring('<joint name="t" type="fixed"></joint>')
    out["no_origin_is_identity"] = bool((ref_utils.parse_urdf_transform(j2) == np.eye(4)).all())
    (HERE / "urdf.json").write_text(json.dumps(out, indent=1))


def gen_pipeline_config() -> None:
    """scripts/run_pipeline.py:85-163 - PipelineConfig.from_dict on the shipped YAML and on the schema's corner cases."""
    import dataclasses

    import yaml
    from scripts import run_pipeline, run_slam

    cases = {
        "slam_config.yaml": yaml.safe_load((REF / "config" / "slam_config.yaml").read_text()),
        "empty": {},
        "deprecated_rgbd_camera_ip": {"cameras": [{"ip": "10.0.0.1"}, {"ip": "10.0.0.2", "stereo": False}], "rgbd_camera_ip": "10.0.0.2"},
        "enable_rgbd_flags": {"cameras": [{"ip": "a", "enable_rgbd": True, "sensor_type": "mono"},
                                          {"ip": "b", "resolution": [1920, 1200], "output_resolution": [640, 400]},
                                          {"ip": "c", "enable_rgbd": True, "rgb_output_resolution": [1280, 720]}],
                              "fps": 15, "rig_queue_size": 4, "urdf_path": "/tmp/x.urdf"},
        "nvblox_not_a_list": {"cameras": [{"ip": "a"}], "nvblox_cameras": "a"},
    }
    out = {"camera_map_run_pipeline": run_pipeline.CAMERA_MAP, "camera_map_run_slam": run_slam.CAMERA_MAP, "cases": {}}
    for name, data in cases.items():
        cfg = run_pipeline.PipelineConfig.from_dict(data)
        d = dataclasses.asdict(cfg)
        d["urdf_is_default_brackets"] = d["urdf_path"].endswith("examples/assets/brackets.urdf")
        if d["urdf_is_default_brackets"]:
            d["urdf_path"] = "<default>"
        d["num_cameras"] = cfg.calculate_num_cameras()
        out["cases"][name] = {"input": data, "config": d}
    (HERE / "pipeline_config.json").write_text(json.dumps(out, indent=1))


def gen_cv_arith() -> None:
    from oracle import rectify

    rng = np.random.default_rng(1337)
    out = {}
    bgr = make_image(rng, "bgr8", 48, 32)
    nv12 = make_image(rng, "nv12", 48, 32)
    nv12_lim = nv12.copy()
    nv12_lim[:32] = rng.integers(16, 236, size=(32, 48), dtype=np.uint8)
    nv12_lim[32:] = rng.integers(16, 241, size=(16, 48), dtype=np.uint8)
    out["bgr"] = bgr
    out["bgr2rgb"] = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
    out["bgr2gray"] = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    for tag, buf in (("nv12", nv12), ("nv12lim", nv12_lim)):
        out[tag] = buf
        out[f"{tag}2gray"] = cv2.cvtColor(buf, cv2.COLOR_YUV2GRAY_NV12)
        out[f"{tag}2rgb"] = cv2.cvtColor(buf, cv2.COLOR_YUV2RGB_NV12)
        out[f"{tag}2bgr"] = cv2.cvtColor(buf, cv2.COLOR_YUV2BGR_NV12)

    # a small stereo pair: calibration -> stereoRectify -> maps -> remap
    src = SyntheticCameraSource(SyntheticCameraConfig(name="g", resolution=(160, 100), seed=21, pool=1))
    (il, ir_), (el, er) = src.get_intrinsics(), src.get_extrinsics()
    r1, r2, p1, p2 = rectify.stereo_rectify_cv(il.matrix, il.coeffs, ir_.matrix, ir_.coeffs, (160, 100), el.to_4x4_matrix(), er.to_4x4_matrix())
    for side, intr, r, p, img in (("l", il, r1, p1, src._pool[0][0]), ("r", ir_, r2, p2, src._pool[0][1])):
        mx, my = cv2.initUndistortRectifyMap(intr.matrix, intr.coeffs[:8], r, p, (160, 100), cv2.CV_32FC1)
        out[f"K_{side}"], out[f"D_{side}"], out[f"R_{side}"], out[f"P_{side}"] = intr.matrix, intr.coeffs, r, p
        out[f"mapx_{side}"], out[f"mapy_{side}"] = mx, my
        out[f"img_{side}"] = img
        out[f"rect_{side}"] = cv2.remap(img, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    out["T_left_to_ref"], out["T_right_to_ref"] = el.to_4x4_matrix(), er.to_4x4_matrix()
    # a map that leaves the image on every side (border taps), u8 / 3-channel / f32
    yy, xx = np.mgrid[0:32, 0:48].astype(np.float32)
    mx = (xx * 1.13 - 3.3 + 0.05 * yy).astype(np.float32)
    my = (yy * 1.21 - 4.7 - 0.03 * xx).astype(np.float32)
    gray = out["bgr2gray"]
    out["edge_mapx"], out["edge_mapy"] = mx, my
    out["edge_rect_u8"] = cv2.remap(gray, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    out["edge_rect_c3"] = cv2.remap(bgr, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    out["edge_rect_f32"] = cv2.remap(gray.astype(np.float32), mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    out["depth"] = make_depth(rng, 48, 32)
    np.savez_compressed(HERE / "cv_arith.npz", **out)


if __name__ == "__main__":
    gen_rig_sync()
    gen_calibration()
    gen_isaac_adapter()
    gen_rgbd_publisher()
    gen_luxonis_calibration()
    gen_urdf()
    gen_pipeline_config()
    gen_cv_arith()
    for f in sorted(HERE.iterdir()):
        print(f"{f.name:28s} {f.stat().st_size:8d} B")
