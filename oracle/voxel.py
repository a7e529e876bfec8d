"""Oracle of the voxel down-sampled cloud (``ti_voxel_cloud``)  -  TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

``oracle/voxel.c`` restates the key arithmetic in C (IEEE double, fused multiply-adds in a fixed order - the reason it
is C: numpy has no ``fma``); this module compiles it on first use, packs keys into the record format of
``include/thoringest.h`` and removes duplicates with ``np.unique``.  ``voxel_keys_f64`` is the independent float64
numpy route (``floor(backproject(...) / voxel_size)``) the C keys are checked against.

The reference's parameters: ``voxel_size`` 0.05 m and a 10 m integration distance
(``launch/thor_nvblox.launch.py:26-31``); cloud type ``N x 3`` (``thor_slam/slam/interface.py:134-138``).
"""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from oracle import backproject as ob

HERE = Path(__file__).resolve().parent
SRC = HERE / "voxel.c"
OUT = HERE / "_build" / "libvoxel_oracle.so"
KEY_BIAS = 16384

_lib = None


def build(force: bool = False) -> Path:
    if force or not OUT.exists() or OUT.stat().st_mtime < SRC.stat().st_mtime:
        OUT.parent.mkdir(exist_ok=True)
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", str(SRC), "-o", str(OUT), "-lm"], check=True)
    return OUT


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.oracle_voxel_keys.restype = C.c_int64
        _lib.oracle_voxel_keys.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double,
                                           C.c_uint32, C.c_void_p, C.c_void_p]
    return _lib


def _km(k: np.ndarray, body_T_cam: np.ndarray):
    k = np.asarray(k, np.float64)
    m = np.asarray(body_T_cam, np.float64)[:3, :4].reshape(-1)
    return (C.c_double * 4)(k[0, 0], k[1, 1], k[0, 2], k[1, 2]), (C.c_double * 12)(*m)


def voxel_keys(depth_mm: np.ndarray, k: np.ndarray, body_T_cam: np.ndarray, voxel: float, max_depth_mm: int = 0) -> tuple[np.ndarray, np.ndarray]:
    """(keys HxWx3 int32, valid HxW bool) - the contract arithmetic (C, fma)."""
    depth_mm = np.ascontiguousarray(depth_mm, dtype=np.uint16)
    h, w = depth_mm.shape
    keys = np.zeros((h, w, 3), np.int32)
    valid = np.zeros((h, w), np.uint8)
    kk, mm = _km(k, body_T_cam)
    _load().oracle_voxel_keys(depth_mm.ctypes.data, w, h, kk, mm, float(voxel), int(max_depth_mm) or 65535, keys.ctypes.data, valid.ctypes.data)
    return keys, valid.astype(bool)


def voxel_keys_f64(depth_mm: np.ndarray, k: np.ndarray, body_T_cam: np.ndarray, voxel: float) -> tuple[np.ndarray, np.ndarray]:
    """Independent route: float64 numpy back-projection, then floor(p / voxel).  Also returns the distance of p / voxel to
    the nearest integer per pixel (keys may legitimately differ from the fma route only where that is ~1e-9)."""
    pts, _, _ = ob.backproject(depth_mm, k, body_T_cam)
    q = pts / voxel
    return np.floor(q).astype(np.int64), np.abs(q - np.rint(q)).min(axis=-1)


def pack_records(keys: np.ndarray, set_id: int, tag: int = 0) -> np.ndarray:
    keys = np.asarray(keys, np.int64).reshape(-1, 3)
    if len(keys) and (np.abs(keys).max() >= KEY_BIAS):
        raise ValueError("voxel key outside the 15-bit field")
    b = (keys + KEY_BIAS).astype(np.uint64)
    return (np.uint64(tag) << np.uint64(56)) | (np.uint64(set_id) << np.uint64(45)) | (b[:, 0] << np.uint64(30)) | (b[:, 1] << np.uint64(15)) | b[:, 2]


def unpack_records(records: np.ndarray) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(tag, set, keys Nx3 int64)"""
    r = np.asarray(records, np.uint64)
    keys = np.stack([(r >> np.uint64(30)) & np.uint64(0x7FFF), (r >> np.uint64(15)) & np.uint64(0x7FFF), r & np.uint64(0x7FFF)], axis=-1).astype(np.int64) - KEY_BIAS
    return (r >> np.uint64(56)).astype(np.int64), ((r >> np.uint64(45)) & np.uint64(0x7FF)).astype(np.int64), keys


def voxel_records(cameras: list[tuple[np.ndarray, np.ndarray, np.ndarray]], voxel: float, max_depth_mm: int = 0, set_id: int = 0,
                  tag: int = 0) -> np.ndarray:
    """Sorted distinct records of ONE frame set: ``cameras`` = [(depth HxW u16, K 3x3, body_T_cam 4x4), ...]."""
    parts = []
    for depth, k, m in cameras:
        keys, valid = voxel_keys(depth, k, m, voxel, max_depth_mm)
        parts.append(pack_records(keys[valid], set_id, tag))
    return np.unique(np.concatenate(parts)) if parts else np.zeros(0, np.uint64)


def record_points(records: np.ndarray, voxel: float) -> np.ndarray:
    """Voxel centres, N x 3 float32 (``ti_voxel_points``): ``(k + 0.5) * voxel`` in double, rounded once."""
    _, _, keys = unpack_records(records)
    return ((keys.astype(np.float64) + 0.5) * voxel).astype(np.float32)
