"""Consumer side (API mirror of ``thor_slam.slam`` minus the ROS adapter)."""

from thor_slam_b200.slam.interface import (
    CameraConfig,
    MapPoint,
    SlamConfig,
    SlamEngine,
    SlamMap,
    SlamPose,
    TrackingState,
)
from thor_slam_b200.slam.recording import RecordingSlamEngine, extract_cameras

__all__ = [
    "CameraConfig",
    "MapPoint",
    "RecordingSlamEngine",
    "SlamConfig",
    "SlamEngine",
    "SlamMap",
    "SlamPose",
    "TrackingState",
    "extract_cameras",
]
