"""Host-side calibration maths of the ingest stage (float64, one-off per rig).

Turns the reference's calibration containers into what the kernels consume:

* remap maps (float32 ``mapx, mapy``, OpenCV convention) from ``Intrinsics`` + ``Extrinsics`` -
  the undistortion the reference leaves to cuVSLAM by publishing ``CameraInfo{D,K,R=I,P}`` with
  ``rectified_images:=false`` (``thor_slam/slam/adapters/isaac_ros.py:364-411``, ``Makefile:77-80``);
* ``body_T_camera`` 4x4s from ``RigCalibration.get_world_extrinsics`` (``thor_slam/camera/rig.py:35-70``)
  with the RDF->FLU rotation of ``isaac_ros.py:42-49`` folded in.

Conventions kept from the reference:

* which distortion coefficients count (``isaac_ros.py:370-383``): ``len >= 8`` -> rational model on
  the first 8 (k1 k2 p1 p2 k3 k4 k5 k6), 5 -> plumb_bob, 4 -> equidistant (fisheye), anything else
  -> zero-padded plumb_bob;
* stereo ``Extrinsics`` are left->CAM_A and right->CAM_A in metres (``drivers/luxonis.py:675-709``).

The per-pixel map is built here in float64 numpy following ``cv::initUndistortRectifyMap``
operation for operation (tests check it against OpenCV bit for bit); the four small rectification
matrices come from OpenCV's ``stereoRectify`` - the reference's own declared dependency
(``thor_slam/requirements.txt:4``).
"""

from __future__ import annotations

from typing import Sequence

import numpy as np

from thor_slam_b200.camera.calibration import Extrinsics, Intrinsics

RDF_TO_FLU_MATRIX = np.array(
    [
        [0.0, 0.0, 1.0, 0.0],  # x_flu =  z_rdf (forward)
        [-1.0, 0.0, 0.0, 0.0],  # y_flu = -x_rdf (left)
        [0.0, -1.0, 0.0, 0.0],  # z_flu = -y_rdf (up)
        [0.0, 0.0, 0.0, 1.0],
    ]
)
FLU_TO_RDF_MATRIX = RDF_TO_FLU_MATRIX.T.copy()  # proper rotation: inverse == transpose

DRB_TO_RDF_MATRIX = np.array(  # OAK-D Pro IMU axes -> camera axes (scripts/run_slam.py:254-266)
    [
        [0.0, 1.0, 0.0, 0.0],
        [1.0, 0.0, 0.0, 0.0],
        [0.0, 0.0, -1.0, 0.0],
        [0.0, 0.0, 0.0, 1.0],
    ]
)


def rotate_imu_sample(sample: dict | None, imu_frame: str = "rdf") -> dict | None:
    """IMU sample with its vectors expressed in the camera (RDF) axes.

    The reference folds the OAK-D Pro's DRB -> RDF rotation into the IMU *extrinsics* only
    (``scripts/run_slam.py:254-276``, ``scripts/run_pipeline.py:543-565``) and publishes the samples themselves as they
    come (``slam/adapters/isaac_ros.py:284-298``); a consumer that wants samples and extrinsics in one convention needs
    the same rotation on ``accelerometer`` / ``gyroscope``.  ``DRB_TO_RDF`` is a proper rotation (det +1), so the
    angular rate transforms like the acceleration.  ``imu_frame="rdf"`` (OAK-D Long Range) returns the sample unchanged;
    every other key (timestamps, ...) is passed through."""
    if sample is None or imu_frame == "rdf":
        return sample
    if imu_frame != "drb":
        raise ValueError(f"unknown IMU frame {imu_frame!r} (expected 'drb' or 'rdf')")
    r = DRB_TO_RDF_MATRIX[:3, :3]
    out = dict(sample)
    for key in ("accelerometer", "gyroscope"):
        v = sample.get(key)
        if v is not None and len(v) >= 3:
            out[key] = [float(x) for x in r @ np.asarray(v[:3], dtype=np.float64)] + list(v[3:])
    return out


def distortion_model(coeffs: np.ndarray | Sequence[float]) -> tuple[str, np.ndarray]:
    """(ROS model name, coefficients that take part) - the reference's CameraInfo rule."""
    d = [float(x) for x in np.asarray(coeffs, dtype=np.float64).reshape(-1)]
    if len(d) >= 8:
        return "rational_polynomial", np.array(d[:8])
    if len(d) == 5:
        return "plumb_bob", np.array(d)
    if len(d) == 4:
        return "equidistant", np.array(d)
    return "plumb_bob", np.array((d + [0.0] * 5)[:5])


def undistort_rectify_map(
    k: np.ndarray, coeffs: np.ndarray, r: np.ndarray | None, p: np.ndarray | None, size: tuple[int, int]
) -> tuple[np.ndarray, np.ndarray]:
    """float32 ``(mapx, mapy)`` of shape ``(h, w)``: where output pixel (u, v) samples the source.

    ``r``: rectifying rotation (None = identity); ``p``: new projection 3x3 / 3x4 (None = ``k``).
    """
    model, d = distortion_model(coeffs)
    k = np.asarray(k, dtype=np.float64)
    r = np.eye(3) if r is None else np.asarray(r, dtype=np.float64)
    p = k if p is None else np.asarray(p, dtype=np.float64)
    inv = np.linalg.inv(p[:3, :3] @ r)
    w, h = int(size[0]), int(size[1])
    u = np.arange(w, dtype=np.float64)[None, :]
    v = np.arange(h, dtype=np.float64)[:, None]
    xh = v * inv[0, 1] + inv[0, 2] + u * inv[0, 0]
    yh = v * inv[1, 1] + inv[1, 2] + u * inv[1, 0]
    wh = v * inv[2, 1] + inv[2, 2] + u * inv[2, 0]
    fx, fy, cx, cy = k[0, 0], k[1, 1], k[0, 2], k[1, 2]
    if model == "equidistant":
        x, y = xh / wh, yh / wh
        rad = np.sqrt(x * x + y * y)
        theta = np.arctan(rad)
        t2 = theta * theta
        theta_d = theta * (1 + d[0] * t2 + d[1] * t2 * t2 + d[2] * t2 * t2 * t2 + d[3] * t2 * t2 * t2 * t2)
        safe = np.where(rad == 0, 1.0, rad)
        scale = np.where(rad == 0, 1.0, theta_d / safe)
        return (fx * x * scale + cx).astype(np.float32), (fy * y * scale + cy).astype(np.float32)
    d = np.concatenate([d, np.zeros(12 - len(d))])
    k1, k2, p1, p2, k3, k4, k5, k6, s1, s2, s3, s4 = d
    iw = 1.0 / wh
    x, y = xh * iw, yh * iw
    x2, y2 = x * x, y * y
    r2 = x2 + y2
    xy2 = 2 * x * y
    kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((k6 * r2 + k5) * r2 + k4) * r2)
    xd = x * kr + p1 * xy2 + p2 * (r2 + 2 * x2) + s1 * r2 + s2 * r2 * r2
    yd = y * kr + p1 * (r2 + 2 * y2) + p2 * xy2 + s3 * r2 + s4 * r2 * r2
    return (fx * xd + cx).astype(np.float32), (fy * yd + cy).astype(np.float32)


def stereo_rectification(
    intr_left: Intrinsics, intr_right: Intrinsics, ext_left: Extrinsics, ext_right: Extrinsics, size: tuple[int, int]
) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """(R1, R2, P1, P2) - ``cv2.stereoRectify(..., CALIB_ZERO_DISPARITY, alpha=0)`` on the pair."""
    import cv2  # opencv-python: a dependency of the reference itself

    ml, dl = distortion_model(intr_left.coeffs)
    mr, dr = distortion_model(intr_right.coeffs)
    left_to_right = np.linalg.inv(ext_right.to_4x4_matrix()) @ ext_left.to_4x4_matrix()
    rot, trans = left_to_right[:3, :3].copy(), left_to_right[:3, 3].copy()
    kl, kr_ = np.asarray(intr_left.matrix, np.float64), np.asarray(intr_right.matrix, np.float64)
    if ml == "equidistant" and mr == "equidistant":
        r1, r2, p1, p2, _q = cv2.fisheye.stereoRectify(kl, dl.reshape(4, 1), kr_, dr.reshape(4, 1), size, rot,
                                                       trans.reshape(3, 1), flags=cv2.CALIB_ZERO_DISPARITY, balance=0.0)
    else:
        r1, r2, p1, p2, *_ = cv2.stereoRectify(kl, dl, kr_, dr, size, rot, trans, flags=cv2.CALIB_ZERO_DISPARITY, alpha=0)
    return r1, r2, p1, p2


def stereo_rectify_maps(
    intrinsics: Sequence[Intrinsics], extrinsics: Sequence[Extrinsics], size: tuple[int, int], with_rp: bool = False
):
    """Remap maps of a source: ``[left, right]`` for a stereo pair, ``[undistort-only]`` for a single camera.

    ``with_rp``: also return, per stream, ``{"R": 3x3, "P": 3x4}`` - what a rectified ``CameraInfo`` carries (for a single camera
    ``R = I`` and ``P = [K | 0]``: undistortion only)."""
    if len(intrinsics) == 2:
        r1, r2, p1, p2 = stereo_rectification(intrinsics[0], intrinsics[1], extrinsics[0], extrinsics[1], size)
        maps = [
            undistort_rectify_map(intrinsics[0].matrix, intrinsics[0].coeffs, r1, p1, size),
            undistort_rectify_map(intrinsics[1].matrix, intrinsics[1].coeffs, r2, p2, size),
        ]
        rp = [{"R": np.asarray(r1, np.float64), "P": np.asarray(p1, np.float64)}, {"R": np.asarray(r2, np.float64), "P": np.asarray(p2, np.float64)}]
    else:
        maps = [undistort_rectify_map(i.matrix, i.coeffs, None, None, size) for i in intrinsics]
        rp = [{"R": np.eye(3), "P": np.hstack([np.asarray(i.matrix, np.float64), np.zeros((3, 1))])} for i in intrinsics]
    return (maps, rp) if with_rp else maps


def body_T_camera(rig_T_source: np.ndarray | None, source_T_camera: np.ndarray, rig_frame: str = "rdf") -> np.ndarray:
    """4x4 applied to every back-projected point (``p_body = M @ world_T_camera @ p_cam``).

    ``rig_frame="rdf"``: rig poses follow the Luxonis RDF convention (reference README) and the body
    frame is FLU, so ``M = RDF_TO_FLU_MATRIX``; ``"flu"``: rig poses already are FLU ``base_link`` poses.
    A missing rig pose means the camera extrinsics are used as they are (``rig.py:55-58``).
    """
    world = np.asarray(source_T_camera, dtype=np.float64)
    if rig_T_source is not None:
        world = np.asarray(rig_T_source, dtype=np.float64) @ world
    if rig_frame == "rdf":
        return RDF_TO_FLU_MATRIX @ world
    if rig_frame == "flu":
        return world
    raise ValueError(f"rig_frame must be 'rdf' or 'flu', got {rig_frame!r}")
