"""Oracle: stereo rectification / undistortion maps and the bilinear remap.

The reference never rectifies on the host (it publishes raw images plus
``CameraInfo{D,K,R=I,P}`` and lets cuVSLAM undistort -
``thor_slam/slam/adapters/isaac_ros.py:364-411``), so this stage is restated
with OpenCV following the reference's *conventions*:

* which coefficients count: ``len(D) >= 8`` -> rational model on the first 8,
  ``5`` -> plumb_bob, ``4`` -> equidistant (fisheye), else zero-padded
  plumb_bob (``isaac_ros.py:370-383``, ``scripts/run_pipeline.py:268-280``);
* stereo geometry: ``Extrinsics`` are left->CAM_A and right->CAM_A in metres
  (``thor_slam/camera/drivers/luxonis.py:675-709``), so the left->right
  transform OpenCV wants is ``inv(T_right) @ T_left``.

``*_cv`` = OpenCV; ``*_np`` = float64 / integer numpy restatement of the same
arithmetic (what the CUDA kernel is written against).
TEST INFRASTRUCTURE - see ``oracle/__init__.py``.
"""

from __future__ import annotations

import cv2
import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS  # 32: OpenCV quantises remap coordinates to 1/32 px
INTER_REMAP_COEF_BITS = 15


def select_distortion(coeffs: np.ndarray) -> tuple[str, np.ndarray]:
    """(model name, coefficient vector handed to OpenCV) per isaac_ros.py:370-383."""
    d = [float(x) for x in np.asarray(coeffs).flatten()]
    if len(d) >= 8:
        return "rational_polynomial", np.array(d[:8])
    if len(d) == 5:
        return "plumb_bob", np.array(d)
    if len(d) == 4:
        return "equidistant", np.array(d)
    return "plumb_bob", np.array((d + [0, 0, 0, 0, 0])[:5])


def left_to_right(t_left_to_ref: np.ndarray, t_right_to_ref: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """(R, T) with ``x_right = R x_left + T`` from the two camera->CAM_A 4x4s."""
    m = np.linalg.inv(t_right_to_ref) @ t_left_to_ref
    return m[:3, :3].copy(), m[:3, 3].copy()


def stereo_rectify_cv(
    k_l: np.ndarray, d_l: np.ndarray, k_r: np.ndarray, d_r: np.ndarray, size: tuple[int, int],
    t_left_to_ref: np.ndarray, t_right_to_ref: np.ndarray,
) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """(R1, R2, P1, P2) from ``cv2.stereoRectify(..., CALIB_ZERO_DISPARITY, alpha=0)``."""
    model_l, dl = select_distortion(d_l)
    model_r, dr = select_distortion(d_r)
    rot, trans = left_to_right(t_left_to_ref, t_right_to_ref)
    if model_l == "equidistant" and model_r == "equidistant":
        r1, r2, p1, p2, _q = cv2.fisheye.stereoRectify(
            k_l, dl.reshape(4, 1), k_r, dr.reshape(4, 1), size, rot, trans.reshape(3, 1),
            flags=cv2.CALIB_ZERO_DISPARITY, balance=0.0,
        )
    else:
        r1, r2, p1, p2, *_ = cv2.stereoRectify(
            k_l, dl, k_r, dr, size, rot, trans, flags=cv2.CALIB_ZERO_DISPARITY, alpha=0
        )
    return r1, r2, p1, p2


def undistort_rectify_map_cv(
    k: np.ndarray, coeffs: np.ndarray, r: np.ndarray | None, p: np.ndarray | None, size: tuple[int, int]
) -> tuple[np.ndarray, np.ndarray]:
    """float32 (mapx, mapy): for every *output* pixel, where to sample the source."""
    model, d = select_distortion(coeffs)
    r = np.eye(3) if r is None else r
    p = k if p is None else p
    if model == "equidistant":
        return cv2.fisheye.initUndistortRectifyMap(k, d.reshape(4, 1), r, p, size, cv2.CV_32FC1)
    return cv2.initUndistortRectifyMap(k, d, r, p, size, cv2.CV_32FC1)


def undistort_rectify_map_np(
    k: np.ndarray, coeffs: np.ndarray, r: np.ndarray | None, p: np.ndarray | None, size: tuple[int, int]
) -> tuple[np.ndarray, np.ndarray]:
    """float64 restatement of ``cv::initUndistortRectifyMap`` (rational / plumb_bob models)."""
    model, d = select_distortion(coeffs)
    if model == "equidistant":
        return _fisheye_map_np(k, d, r, p, size)
    d = np.concatenate([d, np.zeros(14 - len(d))])
    k1, k2, p1, p2, k3, k4, k5, k6, s1, s2, s3, s4 = d[:12]
    r = np.eye(3) if r is None else np.asarray(r, dtype=np.float64)
    p = np.asarray(k if p is None else p, dtype=np.float64)
    ir = np.linalg.inv(p[:3, :3] @ r)
    w, h = size
    u = np.arange(w, dtype=np.float64)[None, :]
    v = np.arange(h, dtype=np.float64)[:, None]
    _x = v * ir[0, 1] + ir[0, 2] + u * ir[0, 0]
    _y = v * ir[1, 1] + ir[1, 2] + u * ir[1, 0]
    _w = v * ir[2, 1] + ir[2, 2] + u * ir[2, 0]
    iw = 1.0 / _w
    x = _x * iw
    y = _y * iw
    x2, y2 = x * x, y * y
    r2 = x2 + y2
    _2xy = 2 * x * y
    kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((k6 * r2 + k5) * r2 + k4) * r2)
    xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2) + s1 * r2 + s2 * r2 * r2
    yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy + s3 * r2 + s4 * r2 * r2
    fx, fy, cx, cy = k[0, 0], k[1, 1], k[0, 2], k[1, 2]
    return (fx * xd + cx).astype(np.float32), (fy * yd + cy).astype(np.float32)


def _fisheye_map_np(k, d, r, p, size):
    r = np.eye(3) if r is None else np.asarray(r, dtype=np.float64)
    p = np.asarray(k if p is None else p, dtype=np.float64)
    ir = np.linalg.inv(p[:3, :3] @ r)
    w, h = size
    u = np.arange(w, dtype=np.float64)[None, :]
    v = np.arange(h, dtype=np.float64)[:, None]
    _x = v * ir[0, 1] + ir[0, 2] + u * ir[0, 0]
    _y = v * ir[1, 1] + ir[1, 2] + u * ir[1, 0]
    _w = v * ir[2, 1] + ir[2, 2] + u * ir[2, 0]
    x = _x / _w
    y = _y / _w
    rr = np.sqrt(x * x + y * y)
    theta = np.arctan(rr)
    t2 = theta * theta
    theta_d = theta * (1 + d[0] * t2 + d[1] * t2**2 + d[2] * t2**3 + d[3] * t2**4)
    scale = np.where(rr == 0, 1.0, theta_d / np.where(rr == 0, 1.0, rr))
    fx, fy, cx, cy = k[0, 0], k[1, 1], k[0, 2], k[1, 2]
    return (fx * x * scale + cx).astype(np.float32), (fy * y * scale + cy).astype(np.float32)


def quantize_map(mapx: np.ndarray, mapy: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """``cvRound(map * 32)`` (round-half-even, computed in float32 like OpenCV) as int32."""
    ix = np.rint(mapx.astype(np.float32) * np.float32(INTER_TAB_SIZE)).astype(np.int64)
    iy = np.rint(mapy.astype(np.float32) * np.float32(INTER_TAB_SIZE)).astype(np.int64)
    # OpenCV stores the integer part as int16 with saturation (remap.cpp: saturate_cast<short>)
    x0 = np.clip(ix >> INTER_BITS, -32768, 32767)
    y0 = np.clip(iy >> INTER_BITS, -32768, 32767)
    return (x0 * INTER_TAB_SIZE + (ix & 31)).astype(np.int32), (y0 * INTER_TAB_SIZE + (iy & 31)).astype(np.int32)


def remap_cv(src: np.ndarray, mapx: np.ndarray, mapy: np.ndarray) -> np.ndarray:
    return cv2.remap(src, mapx, mapy, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)


def _taps(src: np.ndarray, x0: np.ndarray, y0: np.ndarray) -> list[np.ndarray]:
    h, w = src.shape[:2]
    out = []
    for dy in (0, 1):
        for dx in (0, 1):
            xs, ys = x0 + dx, y0 + dy
            inside = (xs >= 0) & (xs < w) & (ys >= 0) & (ys < h)
            t = src[np.clip(ys, 0, h - 1), np.clip(xs, 0, w - 1)]
            if src.ndim == 3:
                inside = inside[..., None]
            out.append(np.where(inside, t, 0))
    return out


def remap_u8_np(src: np.ndarray, mapx: np.ndarray, mapy: np.ndarray) -> np.ndarray:
    """Integer restatement of ``cv2.remap`` INTER_LINEAR / BORDER_CONSTANT(0) for u8 images."""
    ix, iy = quantize_map(mapx, mapy)
    x0, y0 = ix >> INTER_BITS, iy >> INTER_BITS
    fx, fy = (ix & 31).astype(np.int64), (iy & 31).astype(np.int64)
    t00, t01, t10, t11 = (t.astype(np.int64) for t in _taps(src, x0, y0))
    if src.ndim == 3:
        fx, fy = fx[..., None], fy[..., None]
    # weights are (32-fx)(32-fy)/1024 scaled to 2^15: exact multiples of 32, so the sum is exact
    s = t00 * (32 - fx) * (32 - fy) + t01 * fx * (32 - fy) + t10 * (32 - fx) * fy + t11 * fx * fy
    return np.clip((s * 32 + (1 << (INTER_REMAP_COEF_BITS - 1))) >> INTER_REMAP_COEF_BITS, 0, 255).astype(np.uint8)


def remap_f32_np(src: np.ndarray, mapx: np.ndarray, mapy: np.ndarray) -> np.ndarray:
    """float restatement of ``cv2.remap`` for f32 images (same 1/32-px quantised taps)."""
    ix, iy = quantize_map(mapx, mapy)
    x0, y0 = ix >> INTER_BITS, iy >> INTER_BITS
    fx = (ix & 31).astype(np.float64) / 32
    fy = (iy & 31).astype(np.float64) / 32
    t00, t01, t10, t11 = (t.astype(np.float64) for t in _taps(src, x0, y0))
    if src.ndim == 3:
        fx, fy = fx[..., None], fy[..., None]
    return (t00 * (1 - fx) * (1 - fy) + t01 * fx * (1 - fy) + t10 * (1 - fx) * fy + t11 * fx * fy).astype(np.float32)


def valid_mask(mapx: np.ndarray, mapy: np.ndarray, src_size: tuple[int, int]) -> np.ndarray:
    """u8 mask: 1 where all four bilinear taps fall inside the ``(w, h)`` source image."""
    ix, iy = quantize_map(mapx, mapy)
    x0, y0 = ix >> INTER_BITS, iy >> INTER_BITS
    w, h = src_size
    return ((x0 >= 0) & (x0 + 1 <= w - 1) & (y0 >= 0) & (y0 + 1 <= h - 1)).astype(np.uint8)
