"""Oracle: coordinate conventions and calibration composition (float64).

Follows ``thor_slam/camera/types.py:41-69`` (4x4 direction),
``thor_slam/camera/rig.py:35-70`` (``world_T_camera = rig_T_source @
source_T_camera``), ``thor_slam/slam/adapters/isaac_ros.py:42-49`` and
``README.md:187-201`` (``RDF_TO_FLU_MATRIX`` applied as ``M @ p``),
``isaac_ros.py:138-157`` (global stream order), ``:364-411`` (CameraInfo),
``thor_slam/camera/utils.py:101-126`` (URDF joint origin -> 4x4, scipy
*intrinsic* "XYZ" Euler order - reference behaviour),
``scripts/run_slam.py:254-276`` (IMU DRB -> RDF).
TEST INFRASTRUCTURE - see ``oracle/__init__.py``.
"""

from __future__ import annotations

import xml.etree.ElementTree as ET

import numpy as np
from scipy.spatial.transform import Rotation

from oracle.rectify import select_distortion

RDF_TO_FLU = np.array(
    [
        [0, 0, 1, 0],
        [-1, 0, 0, 0],
        [0, -1, 0, 0],
        [0, 0, 0, 1],
    ],
    dtype=np.float64,
)

DRB_TO_RDF = np.array(
    [
        [0, 1, 0, 0],
        [1, 0, 0, 0],
        [0, 0, -1, 0],
        [0, 0, 0, 1],
    ],
    dtype=np.float64,
)


def to_4x4(rotation: np.ndarray, translation: np.ndarray) -> np.ndarray:
    m = np.eye(4)
    m[:3, :3] = rotation
    m[:3, 3] = np.asarray(translation).reshape(3)
    return m


def world_T_camera(rig_T_source: np.ndarray | None, source_T_camera: np.ndarray) -> np.ndarray:
    """rig.py:55-68 - missing rig pose means the camera extrinsics are returned as-is."""
    return source_T_camera if rig_T_source is None else rig_T_source @ source_T_camera


def body_T_camera(rig_T_source: np.ndarray | None, source_T_camera: np.ndarray, rig_frame: str = "rdf") -> np.ndarray:
    """Transform applied to every back-projected point.

    ``rig_frame="rdf"``: rig poses are expressed in the Luxonis RDF convention
    (README.md:169) and the body frame is FLU -> ``RDF_TO_FLU @ world_T_camera``.
    ``rig_frame="flu"``: rig poses already are FLU ``base_link`` poses -> no extra rotation.
    """
    w = world_T_camera(rig_T_source, source_T_camera)
    if rig_frame == "rdf":
        return RDF_TO_FLU @ w
    if rig_frame == "flu":
        return w
    raise ValueError(rig_frame)


def stream_order(intrinsics: dict[str, list], num_cameras: int) -> list[tuple[str, int]]:
    """isaac_ros.py:143-157 - sorted source names x cam_idx, capped at num_cameras."""
    out: list[tuple[str, int]] = []
    for name in sorted(intrinsics):
        for idx in range(len(intrinsics[name])):
            if len(out) >= num_cameras:
                break
            out.append((name, idx))
    return out


def camera_info(k: np.ndarray, coeffs: np.ndarray, width: int, height: int) -> dict:
    """Raw-image CameraInfo (isaac_ros.py:364-389, run_pipeline.py:258-292): R = I, P = [K|0]."""
    model, d = select_distortion(coeffs)
    p = np.zeros((3, 4))
    p[:3, :3] = k
    return {
        "width": width,
        "height": height,
        "distortion_model": model,
        "d": [float(x) for x in d],
        "k": np.asarray(k, dtype=np.float64).flatten().tolist(),
        "r": np.eye(3).flatten().tolist(),
        "p": p.flatten().tolist(),
    }


def right_camera_tx(rot_l: np.ndarray, t_l: np.ndarray, t_r: np.ndarray, fx_right: float) -> tuple[float, float]:
    """(baseline, P[0,3]) of the right camera of a pair - isaac_ros.py:392-405."""
    baseline = float((np.asarray(rot_l).T @ (np.asarray(t_r) - np.asarray(t_l)))[0])
    return baseline, -fx_right * baseline


def urdf_origin_to_matrix(xyz: str, rpy: str) -> np.ndarray:
    """utils.py:101-126 - NB scipy ``from_euler("XYZ")`` is *intrinsic* XYZ (reference behaviour)."""
    m = np.eye(4)
    m[:3, :3] = Rotation.from_euler("XYZ", [float(v) for v in rpy.split()], degrees=False).as_matrix()
    m[:3, 3] = [float(v) for v in xyz.split()]
    return m


def urdf_rig_extrinsics(urdf_path: str, camera_map: dict[str, str]) -> dict[str, np.ndarray]:
    """utils.py:129-178 - first ``base_link -> link`` fixed joint per mapped link."""
    root = ET.parse(urdf_path).getroot()
    out: dict[str, np.ndarray] = {}
    for source, link in camera_map.items():
        for joint in root.findall("joint"):
            child, parent = joint.find("child"), joint.find("parent")
            if child is None or child.get("link") != link:
                continue
            if parent is None or parent.get("link") != "base_link":
                continue
            origin = joint.find("origin")
            out[source] = (
                np.eye(4) if origin is None else urdf_origin_to_matrix(origin.get("xyz", "0 0 0"), origin.get("rpy", "0 0 0"))
            )
            break
    return out


def imu_world_extrinsics(rig_T_source: np.ndarray | None, source_T_imu: np.ndarray, imu_frame: str) -> np.ndarray:
    """run_slam.py:254-276 - OAK-D Pro IMU is DRB (rotate into RDF first), OAK-D LR IMU already RDF."""
    m = (DRB_TO_RDF if imu_frame == "drb" else np.eye(4)) @ source_T_imu
    return m if rig_T_source is None else rig_T_source @ m
