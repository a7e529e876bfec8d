"""Camera side of the ingest path (API mirror of ``thor_slam.camera``)."""

from thor_slam_b200.camera.rig import CameraRig, RigCalibration
from thor_slam_b200.camera.types import (
    CameraFrame,
    CameraSource,
    DeviceImage,
    Extrinsics,
    FrameSet,
    IMUExtrinsics,
    Intrinsics,
    SynchronizedFrameSet,
)

__all__ = [
    "CameraFrame",
    "CameraRig",
    "CameraSource",
    "DeviceImage",
    "Extrinsics",
    "FrameSet",
    "IMUExtrinsics",
    "Intrinsics",
    "RigCalibration",
    "SynchronizedFrameSet",
]
