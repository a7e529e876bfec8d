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


def register_colour(depth_mm: np.ndarray, k_depth: np.ndarray, rgb_T_depth: np.ndarray, k_rgb: np.ndarray, rgb: np.ndarray) -> np.ndarray:
    """One colour per depth pixel (HxWx3 u8): the per-pixel form of the depth / RGB association nvblox makes from the two
    images the reference publishes (``scripts/run_pipeline.py:218-256``) with the extrinsics of
    ``drivers/luxonis.py:1068-1091``.  Pinhole on both images' ``K``, nearest RGB pixel (round-half-even), (0,0,0) where
    depth is 0, the point is behind the RGB camera or outside the image.

    Restated in float32 with one rounding per multiply / add / divide, in the order of
    ``thor_slam_b200/csrc/ti_register.cu`` - the constants are formed in float64 and rounded once, exactly as
    ``ti_upload_registration`` does - so the selected pixels are bit-identical, not merely close."""
    f32 = np.float32
    h, w = depth_mm.shape
    rh, rw = rgb.shape[:2]
    kd, kr = np.asarray(k_depth, np.float64), np.asarray(k_rgb, np.float64)
    m = np.asarray(rgb_T_depth, np.float64)
    a = np.stack([m[:3, 0] / kd[0, 0], m[:3, 1] / kd[1, 1], m[:3, 2]], axis=1).astype(f32)  # rows: x, y, z of the RGB frame
    t = m[:3, 3].astype(f32)
    cx, cy = f32(kd[0, 2]), f32(kd[1, 2])
    rfx, rfy, rcx, rcy = f32(kr[0, 0]), f32(kr[1, 1]), f32(kr[0, 2]), f32(kr[1, 2])
    fu = np.arange(w, dtype=f32)[None, :] - cx
    fv = np.arange(h, dtype=f32)[:, None] - cy
    z = depth_mm.astype(f32) * f32(0.001)
    with np.errstate(all="ignore"):
        p = [((a[i, 0] * fu + a[i, 1] * fv) + a[i, 2]) * z + t[i] for i in range(3)]
        ok = (depth_mm > 0) & (p[2] > 0)
        ur = (p[0] / p[2]) * rfx + rcx
        vr = (p[1] / p[2]) * rfy + rcy
        ok &= (ur > -1) & (ur < f32(rw)) & (vr > -1) & (vr < f32(rh))
        iu = np.rint(np.where(ok, ur, 0)).astype(np.int64)
        iv = np.rint(np.where(ok, vr, 0)).astype(np.int64)
    ok &= (iu >= 0) & (iu < rw) & (iv >= 0) & (iv < rh)
    out = np.zeros((h, w, 3), np.uint8)
    out[ok] = rgb[iv[ok], iu[ok]]
    return out


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
