"""Consumer-side API: what a SLAM engine sees of the ingest path.

Signature mirror of the reference's ``thor_slam/slam/interface.py``
(``TrackingState`` :16-23, ``CameraConfig`` :26-33, ``SlamPose`` :36-103,
``MapPoint`` :106-120, ``SlamMap`` :123-138, ``SlamConfig`` :141-165,
``SlamEngine`` :168-270).  The pose estimator itself (cuVSLAM) is external and
out of scope; this module only keeps the contract so an engine written against
the reference can be fed by :class:`thor_slam_b200.ingest.rig.IngestRig`.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from enum import Enum, auto
from types import TracebackType

import numpy as np

try:
    from typing import Self
except ImportError:  # pragma: no cover
    from typing_extensions import Self

from thor_slam_b200.camera.calibration import Extrinsics, Intrinsics
from thor_slam_b200.camera.frames import SynchronizedFrameSet
from thor_slam_b200.camera.rig import RigCalibration


class TrackingState(Enum):
    NOT_INITIALIZED = auto()
    INITIALIZING = auto()
    TRACKING = auto()
    LOST = auto()
    RELOCALIZING = auto()


@dataclass
class CameraConfig:
    """One stream as the engine indexes it (global order: sorted source name, then cam_idx)."""

    intrinsics: Intrinsics
    extrinsics: Extrinsics
    source_name: str
    cam_idx: int


def _quat_to_matrix(q: np.ndarray) -> np.ndarray:
    """[qx, qy, qz, qw] -> 3x3 (normalises first, like scipy's ``Rotation.from_quat``)."""
    x, y, z, w = np.asarray(q, dtype=np.float64) / np.linalg.norm(q)
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ]
    )


def _matrix_to_quat(m: np.ndarray) -> np.ndarray:
    """3x3 -> [qx, qy, qz, qw]; delegates to scipy so the sign/branch choice matches the reference."""
    from scipy.spatial.transform import Rotation

    return Rotation.from_matrix(np.asarray(m, dtype=np.float64)).as_quat()


@dataclass
class SlamPose:
    """Pose estimate: position (m), quaternion [qx, qy, qz, qw], timestamp (s)."""

    position: np.ndarray
    rotation: np.ndarray
    timestamp: float
    tracking_state: TrackingState = TrackingState.TRACKING
    confidence: float = 1.0
    covariance: np.ndarray | None = None

    def to_4x4_matrix(self) -> np.ndarray:
        out = np.eye(4)
        out[:3, :3] = _quat_to_matrix(self.rotation)
        out[:3, 3] = self.position
        return out

    @classmethod
    def from_4x4_matrix(
        cls,
        matrix: np.ndarray,
        timestamp: float,
        tracking_state: TrackingState = TrackingState.TRACKING,
        confidence: float = 1.0,
    ) -> Self:
        m = np.asarray(matrix)
        return cls(
            position=m[:3, 3],
            rotation=_matrix_to_quat(m[:3, :3]),
            timestamp=timestamp,
            tracking_state=tracking_state,
            confidence=confidence,
        )

    @classmethod
    def identity(cls, timestamp: float = 0.0) -> Self:
        return cls(position=np.zeros(3), rotation=np.array([0.0, 0.0, 0.0, 1.0]), timestamp=timestamp)


@dataclass
class MapPoint:
    position: np.ndarray
    color: np.ndarray | None = None
    normal: np.ndarray | None = None
    observations: int = 1


@dataclass
class SlamMap:
    points: list[MapPoint] = field(default_factory=list)
    keyframe_poses: list[SlamPose] = field(default_factory=list)
    timestamp: float = 0.0

    def to_point_cloud(self) -> np.ndarray:
        """N x 3 float array - the shape contract every cloud in this package follows."""
        if not self.points:
            return np.empty((0, 3))
        return np.array([p.position for p in self.points])


@dataclass
class SlamConfig:
    num_cameras: int = 2
    rectified_images: bool = True
    enable_loop_closure: bool = True
    enable_mapping: bool = True
    max_map_size: int = 100000
    expected_fps: float = 30.0


class SlamEngine(ABC):
    """Engine contract; context-manager exit calls ``shutdown``."""

    @abstractmethod
    def initialize(self, calibration: RigCalibration, config: SlamConfig | None = None) -> None: ...

    @abstractmethod
    def process_frames(self, frame_set: SynchronizedFrameSet) -> SlamPose | None: ...

    @abstractmethod
    def get_tracking_state(self) -> TrackingState: ...

    @abstractmethod
    def get_map(self) -> SlamMap: ...

    @abstractmethod
    def reset(self) -> None: ...

    @abstractmethod
    def shutdown(self) -> None: ...

    def save_map(self, path: str) -> bool:
        raise NotImplementedError("This SLAM engine does not support map saving")

    def load_map(self, path: str) -> bool:
        raise NotImplementedError("This SLAM engine does not support map loading")

    def relocalize(self) -> bool:
        raise NotImplementedError("This SLAM engine does not support relocalization")

    def __enter__(self) -> Self:
        return self

    def __exit__(
        self,
        exc_type: type[BaseException] | None,
        exc_val: BaseException | None,
        exc_tb: TracebackType | None,
    ) -> None:
        self.shutdown()
