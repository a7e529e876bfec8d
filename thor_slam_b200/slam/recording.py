"""A ``SlamEngine`` that records what it is fed (test double for cuVSLAM).

It follows the consumer contract of the reference adapter
(``thor_slam/slam/adapters/isaac_ros.py:138-157`` camera ordering,
``:327-362`` per-frame handling): global stream order is sorted source name
then ``cam_idx`` capped at ``num_cameras``; 2-D images are ``mono8``; 3-D images
are expected to be ``rgb8`` *already* when they come from the ingest stage
(``frames_are_ingested=True``) or BGR when they come straight from a driver.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from thor_slam_b200.camera.calibration import Extrinsics
from thor_slam_b200.camera.frames import SynchronizedFrameSet
from thor_slam_b200.camera.rig import RigCalibration
from thor_slam_b200.slam.interface import CameraConfig, SlamConfig, SlamEngine, SlamMap, SlamPose, TrackingState


def extract_cameras(cal: RigCalibration, num_cameras: int) -> list[CameraConfig]:
    """Flatten a rig calibration into the engine's global stream order."""
    out: list[CameraConfig] = []
    for source in sorted(cal.intrinsics):
        world = cal.get_world_extrinsics(source) or cal.extrinsics.get(source, [])
        for idx, intr in enumerate(cal.intrinsics[source]):
            if len(out) >= num_cameras:
                break
            extr = world[idx] if idx < len(world) else Extrinsics(np.eye(3), np.zeros(3))
            out.append(CameraConfig(intr, extr, source, idx))
    return out


@dataclass
class RecordedImage:
    index: int
    encoding: str
    image: np.ndarray
    timestamp: float
    frame_id: str


@dataclass
class RecordingSlamEngine(SlamEngine):
    """Keeps the last ``keep`` frame sets it was given, already laid out per global stream."""

    keep: int = 4
    history: list[list[RecordedImage]] = field(default_factory=list)
    clouds: list[dict] = field(default_factory=list)
    _cameras: list[CameraConfig] = field(default_factory=list)
    _state: TrackingState = TrackingState.NOT_INITIALIZED

    def initialize(self, calibration: RigCalibration, config: SlamConfig | None = None) -> None:
        cfg = config or SlamConfig()
        self._cameras = extract_cameras(calibration, cfg.num_cameras)
        self._state = TrackingState.INITIALIZING

    @property
    def cameras(self) -> list[CameraConfig]:
        return self._cameras

    def process_frames(self, frame_set: SynchronizedFrameSet) -> SlamPose | None:
        if self._state is TrackingState.NOT_INITIALIZED:
            raise RuntimeError("Not initialized")
        row: list[RecordedImage] = []
        for i, cam in enumerate(self._cameras):
            fs = frame_set.frame_sets.get(cam.source_name)
            if fs is None or cam.cam_idx >= len(fs.frames):
                continue
            frame = fs.frames[cam.cam_idx]
            img = np.asarray(frame.image)
            row.append(RecordedImage(i, "mono8" if img.ndim == 2 else "rgb8", img, frame.timestamp, f"camera_{i}"))
        self.history.append(row)
        del self.history[: -self.keep]
        if getattr(frame_set, "clouds", None) is not None:
            self.clouds.append(frame_set.clouds)
            del self.clouds[: -self.keep]
        self._state = TrackingState.TRACKING
        return SlamPose.identity(frame_set.timestamp)

    def get_tracking_state(self) -> TrackingState:
        return self._state

    def get_map(self) -> SlamMap:
        return SlamMap()

    def reset(self) -> None:
        self.history.clear()
        self.clouds.clear()

    def shutdown(self) -> None:
        self._state = TrackingState.NOT_INITIALIZED
