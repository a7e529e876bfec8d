"""Oracle: depth -> 3-D points in the body frame (float64) + validity mask/count.

The reference publishes ``16UC1`` millimetre depth plus ``CameraInfo`` and lets
nvblox back-project (``scripts/run_pipeline.py:247-256``,
``launch/thor_nvblox.launch.py:53-81``); restated here as a pinhole model on the
*depth image's* ``K`` (no distortion term - nvblox uses ``K`` only):

    z = d_mm * 1e-3 ; x = (u - cx) / fx * z ; y = (v - cy) / fy * z   (RDF optical frame)
    valid = d_mm > 0                                   (examples/rgbd_stream.py:121-123)
    p_body = body_T_camera @ [x, y, z, 1]              (oracle/conventions.py)

Invalid pixels are written as (0, 0, 0).  Stats follow
``examples/rgbd_stream.py:270-276`` (mean/min/max over valid pixels).
TEST INFRASTRUCTURE - see ``oracle/__init__.py``.
"""

from __future__ import annotations

import numpy as np


def backproject(depth_mm: np.ndarray, k: np.ndarray, body_T_cam: np.ndarray) -> tuple[np.ndarray, np.ndarray, int]:
    """(points HxWx3 float64, mask HxW u8, valid count)."""
    h, w = depth_mm.shape
    fx, fy, cx, cy = float(k[0, 0]), float(k[1, 1]), float(k[0, 2]), float(k[1, 2])
    u = np.arange(w, dtype=np.float64)[None, :]
    v = np.arange(h, dtype=np.float64)[:, None]
    z = depth_mm.astype(np.float64) * 1e-3
    x = (u - cx) / fx * z
    y = (v - cy) / fy * z
    cam = np.stack([x, y, z], axis=-1)
    m = np.asarray(body_T_cam, dtype=np.float64)
    pts = cam @ m[:3, :3].T + m[:3, 3]
    valid = depth_mm > 0
    pts[~valid] = 0.0
    return pts, valid.astype(np.uint8), int(valid.sum())


def depth_stats(depth_mm: np.ndarray) -> dict:
    valid = depth_mm[depth_mm > 0]
    if valid.size == 0:
        return {"count": 0, "mean": 0.0, "min": 0, "max": 0}
    return {"count": int(valid.size), "mean": float(valid.mean()), "min": int(valid.min()), "max": int(valid.max())}


def points_close(got: np.ndarray, ref: np.ndarray, rtol: float = 1e-5, floor: float = 1e-3) -> tuple[bool, float]:
    """north_star tolerance: ``|got - ref|_inf <= rtol * max(|ref|_inf, floor)`` per point."""
    got = np.asarray(got, dtype=np.float64).reshape(-1, 3)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1, 3)
    err = np.abs(got - ref).max(axis=1)
    scale = np.maximum(np.abs(ref).max(axis=1), floor)
    worst = float((err / scale).max()) if len(err) else 0.0
    return worst <= rtol, worst
