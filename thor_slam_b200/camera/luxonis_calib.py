"""Calibration conventions of the reference's OAK driver, as code.

``thor_slam/camera/drivers/luxonis.py`` turns what a DepthAI ``CalibrationHandler`` stores into the ``Intrinsics`` /
``Extrinsics`` every later stage consumes.  Those conventions decide what the remap LUTs and body transforms of the ingest
stage mean, so they are restated here against any object that answers the three ``CalibrationHandler`` calls the driver
makes - the real handler on a robot, a recorded one in tests:

* intrinsics are read at the SENSOR resolution and scaled to the published one, x and y independently
  (``fx, cx *= out_w / sensor_w``; ``fy, cy *= out_h / sensor_h`` - ``luxonis.py:620-627``; the device letterboxes, the driver
  stretches: reference behaviour, kept);
* stereo sources publish ``[left = CAM_B, right = CAM_C]``, single sources ``[CAM_A]`` (``:596-673``);
* extrinsics are ``X -> CAM_A`` 4x4 matrices whose translation DepthAI stores in centimetres: ``/ 100`` (``:675-726``);
  a single camera is its own reference (identity);
* RGB-D: RGB = CAM_A at its own sensor / output resolution, identity extrinsics; depth shares the RGB intrinsics when
  ``depth_align_to_rgb`` (rescaled if the two published sizes differ) and is CAM_B at the depth output size otherwise;
  depth extrinsics = CAM_B -> CAM_A in metres either way (``:974-1091``).

Pinned on the reference itself: ``tests/golden/make_golden.py`` runs the driver's getters on a fake handler and
``tests/test_next_rows.py::test_luxonis_calibration_conventions_match_the_driver`` compares.
"""

from __future__ import annotations

from typing import Any, Mapping, Sequence

import numpy as np

from thor_slam_b200.camera.calibration import Extrinsics, Intrinsics

SOCKETS = {"CAM_A": "CAM_A", "CAM_B": "CAM_B", "CAM_C": "CAM_C"}  # pass dai.CameraBoardSocket members on a robot
CM_PER_M = 100.0


def scaled_intrinsics(calib: Any, socket: Any, sensor_res: Sequence[int], out_res: Sequence[int]) -> Intrinsics:
    """K of ``socket`` at ``sensor_res`` scaled to ``out_res`` + the handler's distortion coefficients (all 14 for an OAK)."""
    k = np.array(calib.getCameraIntrinsics(socket, int(sensor_res[0]), int(sensor_res[1])), dtype=np.float64)
    sx, sy = out_res[0] / sensor_res[0], out_res[1] / sensor_res[1]
    k = k.copy()
    k[0, 0] *= sx
    k[1, 1] *= sy
    k[0, 2] *= sx
    k[1, 2] *= sy
    return Intrinsics(width=int(out_res[0]), height=int(out_res[1]), matrix=k, coeffs=np.array(calib.getDistortionCoefficients(socket), dtype=np.float64))


def slam_intrinsics(calib: Any, stereo: bool, mono_sensor_res: Sequence[int], out_res: Sequence[int],
                    sockets: Mapping[str, Any] = SOCKETS) -> list[Intrinsics]:
    """``get_intrinsics()``: ``[left, right]`` of a stereo source, ``[CAM_A]`` of a single one, at the published resolution."""
    names = ("CAM_B", "CAM_C") if stereo else ("CAM_A",)
    return [scaled_intrinsics(calib, sockets[n], mono_sensor_res, out_res) for n in names]


def to_reference_metres(matrix_cm: Any) -> Extrinsics:
    m = np.array(matrix_cm, dtype=np.float64)
    m[:3, 3] /= CM_PER_M
    return Extrinsics.from_4x4_matrix(m)


def slam_extrinsics(calib: Any, stereo: bool, sockets: Mapping[str, Any] = SOCKETS) -> list[Extrinsics]:
    """``get_extrinsics()``: left -> CAM_A and right -> CAM_A in metres; identity for a single camera."""
    if not stereo:
        return [Extrinsics.from_4x4_matrix(np.eye(4))]
    return [to_reference_metres(calib.getCameraExtrinsics(sockets[n], sockets["CAM_A"])) for n in ("CAM_B", "CAM_C")]


def sensor_extrinsics(calib: Any, sockets: Mapping[str, Any] = SOCKETS) -> Extrinsics:
    """``get_sensor_extrinsics()``: IMU -> CAM_A in metres; identity when the handler has none (the driver logs and goes on)."""
    try:
        m = calib.getImuToCameraExtrinsics(sockets["CAM_A"])
    except RuntimeError:
        return Extrinsics.from_4x4_matrix(np.eye(4))
    return to_reference_metres(m)


def rgbd_intrinsics(calib: Any, rgb_sensor_res: Sequence[int], rgb_out: Sequence[int], depth_out: Sequence[int], mono_sensor_res: Sequence[int],
                    depth_align_to_rgb: bool, sockets: Mapping[str, Any] = SOCKETS) -> tuple[Intrinsics, Intrinsics]:
    """``get_rgbd_intrinsics()`` -> (rgb, depth) at their published resolutions."""
    rgb = scaled_intrinsics(calib, sockets["CAM_A"], rgb_sensor_res, rgb_out)
    if depth_align_to_rgb:
        k = rgb.matrix.copy()
        if tuple(depth_out) != tuple(rgb_out):  # validation normally forbids it; the driver rescales
            sx, sy = depth_out[0] / rgb_out[0], depth_out[1] / rgb_out[1]
            k[0, 0] *= sx
            k[1, 1] *= sy
            k[0, 2] *= sx
            k[1, 2] *= sy
        depth = Intrinsics(width=int(depth_out[0]), height=int(depth_out[1]), matrix=k, coeffs=rgb.coeffs.copy())
    else:
        depth = scaled_intrinsics(calib, sockets["CAM_B"], mono_sensor_res, depth_out)
    return rgb, depth


def rgbd_extrinsics(calib: Any, sockets: Mapping[str, Any] = SOCKETS) -> tuple[Extrinsics, Extrinsics]:
    """``get_rgbd_extrinsics()`` -> (identity, CAM_B -> CAM_A in metres)."""
    return Extrinsics.from_4x4_matrix(np.eye(4)), to_reference_metres(calib.getCameraExtrinsics(sockets["CAM_B"], sockets["CAM_A"]))
