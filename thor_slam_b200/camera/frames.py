"""Frame containers and the ``CameraSource`` plugin interface.

API mirror of the reference's ``thor_slam/camera/types.py``:
``CameraFrame`` :84-91, ``SensorData``/``IMUData`` :94-128, ``CameraSource``
:131-210, ``FrameSet`` :213-254, ``SynchronizedFrameSet`` :257-307, ``IPv4``
:13-28.  Names, fields, defaults and return-``None`` behaviour are kept so a
``SlamEngine`` written for the reference consumes these objects unchanged.

What is new (and invisible to a reference-style consumer):

* ``CameraFrame.image`` may be a :class:`DeviceImage` - an ndarray-compatible
  view of a buffer that lives in HBM; it copies to the host only when a numpy
  consumer touches it (``np.asarray(frame.image)``).
* ``SynchronizedFrameSet.clouds`` optionally carries the body-frame point
  clouds produced by the ingest stage (``None`` for plain frame sets).
"""

from __future__ import annotations

import re
from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from typing import Any, Literal

import numpy as np

try:
    from typing import Self
except ImportError:  # pragma: no cover
    from typing_extensions import Self

from thor_slam_b200.camera.calibration import Extrinsics, Intrinsics

CameraSensorType = Literal["COLOR", "MONO"]

_IPV4 = re.compile(r"^((25[0-5]|2[0-4][0-9]|[01]?[0-9][0-9]?)\.){3}(25[0-5]|2[0-4][0-9]|[01]?[0-9][0-9]?)$")


class IPv4(str):
    """A dotted-quad string that validates itself (ValueError on anything else)."""

    def __init__(self, ip: str) -> None:
        if _IPV4.match(ip) is None:
            raise ValueError(f"Invalid IPv4 address: {ip}")
        self._ip = ip

    def __str__(self) -> str:
        return self._ip

    @property
    def ip(self) -> str:
        return self._ip


class DeviceImage:
    """An image that lives in GPU memory but quacks like ``np.ndarray``.

    ``tensor`` is the carrier (a torch CUDA tensor, HxW or HxWxC); ``shape``,
    ``dtype``, ``ndim`` and ``__array__`` make it acceptable wherever the
    reference's consumers expect ``CameraFrame.image`` (they only look at
    ``len(img.shape)`` and hand the array to OpenCV - isaac_ros.py:351-358).

    LIFETIME.  The reference hands out fresh arrays (``getCvFrame()`` copies); here the pixels live in a ring slot of the rig
    that produced them and stay valid for ``queue_size`` further polls of that rig - the window in which the reference's
    queue would still hold the frame.  ``np.asarray(image)`` is a zero-copy view of the slot's pinned host mirror (the
    device -> host copy was enqueued when the frame set left the rig); ``np.array(image)`` / ``image.copy()`` /
    ``image.tensor.clone()`` give pixels of unlimited lifetime.  Touching a frame whose slot has since been re-used raises
    ``RuntimeError`` instead of returning somebody else's pixels.
    """

    __slots__ = ("tensor", "_host", "_ready", "_mirror", "_mirror_ready", "_guard")

    def __init__(self, tensor: Any, ready_event: Any = None, host_mirror: Any = None, mirror_event: Any = None, guard: Any = None) -> None:
        self.tensor = tensor
        self._host: np.ndarray | None = None
        self._ready = ready_event
        self._mirror = host_mirror        # pinned host tensor the rig is copying `tensor` into (or None)
        self._mirror_ready = mirror_event
        self._guard = guard               # callable -> bool: the slot still holds this frame

    @property
    def shape(self) -> tuple[int, ...]:
        return tuple(self.tensor.shape)

    @property
    def ndim(self) -> int:
        return len(self.tensor.shape)

    @property
    def dtype(self) -> np.dtype:
        return np.dtype(str(self.tensor.dtype).replace("torch.", ""))

    def _check_alive(self) -> None:
        if self._guard is not None and not self._guard():
            raise RuntimeError("this frame's ring slot has been re-used (frames stay valid for queue_size polls of their rig); "
                               "copy frames you keep longer: np.array(image) or image.tensor.clone()")

    def wait(self) -> None:
        """Block until the pixels are in ``tensor`` (device side)."""
        if self._ready is not None:
            self._ready.synchronize()
            self._ready = None

    def numpy(self) -> np.ndarray:
        if self._host is None:
            self._check_alive()
            if self._mirror is not None:
                if self._mirror_ready is not None:
                    self._mirror_ready.synchronize()
                    self._mirror_ready = None
                self._check_alive()
                m = self._mirror
                self._host = m.view(torch_int16()).numpy().view(np.uint16) if str(m.dtype) == "torch.uint16" else m.numpy()
            else:
                self.wait()
                self._host = self.tensor.cpu().numpy()
        elif self._mirror is not None:
            self._check_alive()
        return self._host

    def copy(self) -> np.ndarray:
        return np.array(self.numpy(), copy=True)

    def __array__(self, dtype: Any = None, copy: Any = None) -> np.ndarray:
        arr = self.numpy()
        if dtype is not None:
            arr = arr.astype(dtype, copy=False)
        return np.array(arr, copy=True) if copy else arr

    def __len__(self) -> int:
        return self.shape[0]

    def __getitem__(self, idx: Any) -> Any:
        return self.numpy()[idx]


def torch_int16() -> Any:
    import torch

    return torch.int16


@dataclass
class CameraFrame:
    """One image with its capture time (seconds, host clock), counter and stream name."""

    image: Any  # np.ndarray (reference) or DeviceImage
    timestamp: float
    sequence_num: int
    camera_name: str


class SensorData(ABC):
    """Abstract non-image sample (the reference only has IMU)."""

    @abstractmethod
    def get_timestamp(self) -> float: ...

    @abstractmethod
    def get_sequence_num(self) -> int: ...

    @abstractmethod
    def get_data(self) -> dict: ...


class IMUData(SensorData):
    """Accelerometer + gyroscope sample."""

    accelerometer: np.ndarray
    gyroscope: np.ndarray
    timestamp: float
    sequence_num: int

    def get_timestamp(self) -> float:
        return self.timestamp

    def get_sequence_num(self) -> int:
        return self.sequence_num

    def get_data(self) -> dict:
        return {"accelerometer": self.accelerometer, "gyroscope": self.gyroscope}


class CameraSource(ABC):
    """Driver plugin interface; identical member set to the reference ABC."""

    @property
    @abstractmethod
    def name(self) -> str: ...

    @abstractmethod
    def start(self) -> None: ...

    @abstractmethod
    def stop(self) -> None: ...

    @abstractmethod
    def get_latest_frames(self) -> list[CameraFrame]:
        """Blocking: next ``[left, right]`` or ``[rgb]``."""

    @abstractmethod
    def try_get_latest_frames(self) -> list[CameraFrame] | None:
        """Non-blocking variant: ``None`` when nothing is ready."""

    @abstractmethod
    def get_intrinsics(self) -> list[Intrinsics]: ...

    @abstractmethod
    def get_extrinsics(self) -> list[Extrinsics]: ...

    @abstractmethod
    def get_sensor_extrinsics(self) -> Extrinsics | None:
        """Pose of a non-camera sensor (IMU) in the source frame, or ``None``."""

    @abstractmethod
    def get_timestamped_sensor_data(self) -> tuple[dict | None, float | None]: ...

    def try_get_timestamped_sensor_data(self) -> tuple[dict | None, float | None]:
        """Never raises: ``(None, None)`` when the source has no sensor or the read fails."""
        if not self.has_sensor_data:
            return None, None
        try:
            return self.get_timestamped_sensor_data()
        except Exception:  # the reference swallows driver errors here too
            return None, None

    @property
    @abstractmethod
    def has_sensor_data(self) -> bool: ...


@dataclass
class FrameSet:
    """Frames of one source captured together; ``timestamp`` is the first frame's."""

    timestamp: float
    frames: list[CameraFrame]
    source_name: str
    sensor_data: dict | None = None
    sensor_timestamp: float | None = None

    @classmethod
    def from_frames(cls, frames: list[CameraFrame], source_name: str) -> Self:
        if not frames:
            raise ValueError("Cannot create FrameSet from empty frame list")
        return cls(timestamp=frames[0].timestamp, frames=frames, source_name=source_name)

    def get_timestamps(self) -> list[float]:
        return [f.timestamp for f in self.frames]

    def get_max_timestamp(self) -> float:
        return max(self.get_timestamps())

    def get_min_timestamp(self) -> float:
        return min(self.get_timestamps())

    def get_timestamp_spread(self) -> float:
        ts = self.get_timestamps()
        return max(ts) - min(ts)


@dataclass
class SynchronizedFrameSet:
    """One frame set per source, all matched to ``timestamp``."""

    timestamp: float
    frame_sets: dict[str, FrameSet]
    max_time_delta: float
    sensor_data: dict | None = None
    sensor_timestamp: float | None = None
    # --- additive fields (ignored by reference-style consumers) -------------
    clouds: dict[str, Any] | None = field(default=None, repr=False)

    def get_all_frames(self) -> list[CameraFrame]:
        return [f for fs in self.frame_sets.values() for f in fs.frames]

    def get_frames_for_source(self, source_name: str) -> list[CameraFrame] | None:
        fs = self.frame_sets.get(source_name)
        return None if fs is None else fs.frames

    def get_all_timestamps(self) -> dict[str, list[float]]:
        return {name: fs.get_timestamps() for name, fs in self.frame_sets.items()}

    def get_timestamp_for_frame(self, source_name: str, frame_index: int) -> float | None:
        fs = self.frame_sets.get(source_name)
        if fs is None or not 0 <= frame_index < len(fs.frames):
            return None
        return fs.frames[frame_index].timestamp
