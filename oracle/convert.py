"""Oracle: image format conversion (u8, bit-exact).

Follows: ``cv2.cvtColor(img, cv2.COLOR_BGR2RGB)`` at
``thor_slam/slam/adapters/isaac_ros.py:357`` and ``scripts/run_pipeline.py:234``;
mono8 pass-through at ``isaac_ros.py:352-353``; the NV12 -> BGR/GRAY conversion
that ``dai.ImgFrame.getCvFrame()`` performs for the driver
(``thor_slam/camera/drivers/luxonis.py:773,788,806``; depthai is an un-pinned,
un-vendored wheel, restated here as OpenCV's ``COLOR_YUV2*_NV12``).

Every function exists twice: ``*_cv`` calls OpenCV the way the reference does,
``*_np`` restates the integer arithmetic in numpy so the CUDA kernels have a
formula to be checked against.  TEST INFRASTRUCTURE - see ``oracle/__init__.py``.
"""

from __future__ import annotations

import cv2
import numpy as np

# cv2.COLOR_BGR2GRAY, 15-bit fixed point (OpenCV 4.x color_rgb.simd.hpp: BY15, GY15, RY15)
GRAY_B, GRAY_G, GRAY_R, GRAY_SHIFT = 3735, 19235, 9798, 15

# cv2.COLOR_YUV2BGR_NV12: BT.601 limited range, 20-bit fixed point (color_yuv.simd.hpp)
YUV_SHIFT = 20
YUV_CY, YUV_CVR, YUV_CVG, YUV_CUG, YUV_CUB = 1220542, 1673527, -852492, -409993, 2116026


def bgr_to_rgb_cv(img: np.ndarray) -> np.ndarray:
    return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)


def bgr_to_rgb_np(img: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(img[..., ::-1])


def bgr_to_gray_cv(img: np.ndarray) -> np.ndarray:
    return cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)


def bgr_to_gray_np(img: np.ndarray) -> np.ndarray:
    b = img[..., 0].astype(np.int32)
    g = img[..., 1].astype(np.int32)
    r = img[..., 2].astype(np.int32)
    return ((b * GRAY_B + g * GRAY_G + r * GRAY_R + (1 << (GRAY_SHIFT - 1))) >> GRAY_SHIFT).astype(np.uint8)


def nv12_to_gray_cv(buf: np.ndarray) -> np.ndarray:
    return cv2.cvtColor(buf, cv2.COLOR_YUV2GRAY_NV12)


def nv12_to_gray_np(buf: np.ndarray) -> np.ndarray:
    h = buf.shape[0] * 2 // 3
    return np.ascontiguousarray(buf[:h])


def nv12_to_bgr_cv(buf: np.ndarray) -> np.ndarray:
    return cv2.cvtColor(buf, cv2.COLOR_YUV2BGR_NV12)


def nv12_to_rgb_cv(buf: np.ndarray) -> np.ndarray:
    return cv2.cvtColor(buf, cv2.COLOR_YUV2RGB_NV12)


def _nv12_planes(buf: np.ndarray) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    h = buf.shape[0] * 2 // 3
    w = buf.shape[1]
    y = buf[:h].astype(np.int64)
    uv = buf[h:].reshape(h // 2, w // 2, 2).astype(np.int64)
    u = np.repeat(np.repeat(uv[..., 0], 2, axis=0), 2, axis=1)
    v = np.repeat(np.repeat(uv[..., 1], 2, axis=0), 2, axis=1)
    return y, u, v


def nv12_to_rgb_np(buf: np.ndarray) -> np.ndarray:
    y, u, v = _nv12_planes(buf)
    yy = np.maximum(0, y - 16) * YUV_CY
    u = u - 128
    v = v - 128
    half = 1 << (YUV_SHIFT - 1)
    r = (yy + YUV_CVR * v + half) >> YUV_SHIFT
    g = (yy + YUV_CVG * v + YUV_CUG * u + half) >> YUV_SHIFT
    b = (yy + YUV_CUB * u + half) >> YUV_SHIFT
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)


def nv12_to_bgr_np(buf: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(nv12_to_rgb_np(buf)[..., ::-1])


# --- the stream-level rule of the ingest stage ---------------------------------
def convert(img: np.ndarray, src_fmt: str, dst_fmt: str) -> np.ndarray:
    """src_fmt in {mono8, bgr8, nv12}; dst_fmt in {mono8, rgb8}."""
    if src_fmt == "mono8":
        if dst_fmt != "mono8":
            raise ValueError("mono8 input can only be published as mono8 (isaac_ros.py:352-353)")
        return img
    if src_fmt == "bgr8":
        return bgr_to_rgb_cv(img) if dst_fmt == "rgb8" else bgr_to_gray_cv(img)
    if src_fmt == "nv12":
        return nv12_to_rgb_cv(img) if dst_fmt == "rgb8" else nv12_to_gray_cv(img)
    raise ValueError(f"unknown source format {src_fmt!r}")
