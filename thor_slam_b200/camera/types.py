"""Import-path mirror of the reference's ``thor_slam.camera.types``.

``from thor_slam.camera.types import X`` becomes
``from thor_slam_b200.camera.types import X`` for every public ``X``.
"""

from thor_slam_b200.camera.calibration import Extrinsics, IMUExtrinsics, Intrinsics
from thor_slam_b200.camera.frames import (
    CameraFrame,
    CameraSensorType,
    CameraSource,
    DeviceImage,
    FrameSet,
    IMUData,
    IPv4,
    SensorData,
    SynchronizedFrameSet,
)

__all__ = [
    "CameraFrame",
    "CameraSensorType",
    "CameraSource",
    "DeviceImage",
    "Extrinsics",
    "FrameSet",
    "IMUData",
    "IMUExtrinsics",
    "IPv4",
    "Intrinsics",
    "SensorData",
    "SynchronizedFrameSet",
]
