"""Calibration containers of the ingest path.

Mirrors the calibration half of the reference's ``thor_slam/camera/types.py``
(``Intrinsics`` :31-38, ``Extrinsics`` :41-69, ``IMUExtrinsics`` :72-81): same
class names, field names, field order and matrix direction, so objects built
for the reference rig can be handed to this package unchanged.

Conventions (all float64 on the host, exactly as the reference keeps them):

* ``Intrinsics.matrix`` is the 3x3 pinhole ``K``; ``coeffs`` is whatever the
  driver returned (14 numbers on an OAK: k1 k2 p1 p2 k3 k4 k5 k6 s1..s4 tx ty).
* ``Extrinsics`` is ``parent_T_camera``: ``[R t; 0 1] @ p_camera = p_parent``
  with ``t`` in metres.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np

try:  # Python >= 3.11
    from typing import Self
except ImportError:  # pragma: no cover
    from typing_extensions import Self


@dataclass
class Intrinsics:
    """Pinhole model of one stream at its *published* resolution."""

    width: int
    height: int
    matrix: np.ndarray  # 3x3 K
    coeffs: np.ndarray  # distortion coefficients, OpenCV order

    # -- helpers that do not exist in the reference (additive, never required) --
    @property
    def fx(self) -> float:
        return float(np.asarray(self.matrix)[0, 0])

    @property
    def fy(self) -> float:
        return float(np.asarray(self.matrix)[1, 1])

    @property
    def cx(self) -> float:
        return float(np.asarray(self.matrix)[0, 2])

    @property
    def cy(self) -> float:
        return float(np.asarray(self.matrix)[1, 2])


@dataclass
class Extrinsics:
    """Rigid transform ``parent_T_camera`` (rotation 3x3, translation in metres)."""

    rotation: np.ndarray
    translation: np.ndarray

    @classmethod
    def from_4x4_matrix(cls, matrix: np.ndarray | Sequence[Sequence[float]]) -> Self:
        """Split a homogeneous 4x4 into (R, t); rejects any other shape with ValueError."""
        m = np.array(matrix)
        if m.shape != (4, 4):
            raise ValueError(f"Expected 4x4 matrix, got shape {m.shape}")
        return cls(rotation=m[:3, :3], translation=m[:3, 3])

    def to_4x4_matrix(self) -> np.ndarray:
        """Homogeneous 4x4 ``[R t; 0 1]`` (float64)."""
        out = np.eye(4)
        out[:3, :3] = self.rotation
        out[:3, 3] = self.translation
        return out

    @classmethod
    def identity(cls) -> Self:
        return cls(rotation=np.eye(3), translation=np.zeros(3))


@dataclass
class IMUExtrinsics:
    """IMU pose (already in the world/base frame) and the source that carries the IMU."""

    source_name: str
    extrinsics: Extrinsics

    def to_4x4_matrix(self) -> np.ndarray:
        return self.extrinsics.to_4x4_matrix()
