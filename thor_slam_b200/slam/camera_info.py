"""Consumer-side message preparation as plain data (no ROS): CameraInfo and static transforms.

Restates the maths of ``IsaacRosAdapter.process_frames`` (``thor_slam/slam/adapters/isaac_ros.py:364-411``)
and ``_publish_tf`` (``:159-226``) so that a ROS shim can publish what the ingest stage produces:

* :func:`camera_info_raw`   - the reference's rule for *unrectified* images: distortion model by
  coefficient count, ``R = I``, ``P = [K|0]``, right camera ``P[0,3] = -fx * baseline``;
* :func:`camera_info_rectified` - what goes with OUR rectified output (``rectified_images:=true``):
  ``D = 0``, ``K = P[:, :3]``, ``R = R_i``, ``P = P_i`` from the stereo rectification;
* :func:`static_transforms` - ``base_link -> camera_i`` and ``camera_i -> camera_i_optical_frame``; the
  optical-frame rotation is the reference's (``flu_to_rdf``, quaternion ``[0.5,-0.5,0.5,0.5]`` - quirk (i)
  of SURVEY.md section 8c) unless ``ros_standard_optical=True``.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from thor_slam_b200.camera.calibration import Intrinsics
from thor_slam_b200.ingest.calib import FLU_TO_RDF_MATRIX, RDF_TO_FLU_MATRIX, distortion_model
from thor_slam_b200.slam.interface import CameraConfig, _matrix_to_quat


@dataclass
class CameraInfoData:
    width: int
    height: int
    distortion_model: str
    d: list[float]
    k: list[float]  # 3x3 row-major
    r: list[float]  # 3x3 row-major
    p: list[float]  # 3x4 row-major


def camera_info_raw(cam: CameraConfig, left_of_pair: CameraConfig | None = None) -> CameraInfoData:
    """CameraInfo of an unrectified stream; pass the left camera for the right one of a stereo pair."""
    intr: Intrinsics = cam.intrinsics
    model, d = distortion_model(intr.coeffs)
    p = np.zeros((3, 4))
    p[:3, :3] = intr.matrix
    if left_of_pair is not None:
        t_lr = left_of_pair.extrinsics.rotation.T @ (cam.extrinsics.translation - left_of_pair.extrinsics.translation)
        p[0, 3] = -float(intr.matrix[0, 0]) * float(t_lr[0])
    return CameraInfoData(intr.width, intr.height, model, [float(x) for x in d], np.asarray(intr.matrix, float).flatten().tolist(),
                          np.eye(3).flatten().tolist(), p.flatten().tolist())


def camera_info_rectified(width: int, height: int, r_rect: np.ndarray, p_rect: np.ndarray) -> CameraInfoData:
    """CameraInfo that describes the ingest stage's rectified output of one camera."""
    p = np.zeros((3, 4))
    p[:, : np.asarray(p_rect).shape[1]] = p_rect
    return CameraInfoData(width, height, "plumb_bob", [0.0] * 5, p[:, :3].flatten().tolist(),
                          np.asarray(r_rect, float).flatten().tolist(), p.flatten().tolist())


@dataclass
class TransformData:
    parent: str
    child: str
    translation: list[float]
    rotation_xyzw: list[float]


def static_transforms(cameras: list[CameraConfig], ros_standard_optical: bool = False, imu_extrinsics=None) -> list[TransformData]:
    """``imu_extrinsics``: an ``IMUExtrinsics`` (already in the base frame) adds ``base_link -> imu_link`` (isaac_ros.py:228-262)."""
    optical = (RDF_TO_FLU_MATRIX if ros_standard_optical else FLU_TO_RDF_MATRIX)[:3, :3]
    q_opt = [float(v) for v in _matrix_to_quat(optical)]
    out: list[TransformData] = []
    for i, cam in enumerate(cameras):
        q = [float(v) for v in _matrix_to_quat(cam.extrinsics.rotation)]
        out.append(TransformData("base_link", f"camera_{i}", [float(v) for v in cam.extrinsics.translation], q))
        out.append(TransformData(f"camera_{i}", f"camera_{i}_optical_frame", [0.0, 0.0, 0.0], q_opt))
    if imu_extrinsics is not None:
        e = imu_extrinsics.extrinsics
        out.append(TransformData("base_link", "imu_link", [float(v) for v in e.translation], [float(v) for v in _matrix_to_quat(e.rotation)]))
    return out
