"""Rig calibration and the host-side frame-set synchroniser.

Behavioural mirror of the reference's ``thor_slam/camera/rig.py``:

* ``RigCalibration.get_world_extrinsics`` (:35-70):
  ``world_T_camera = rig_T_source @ source_T_camera`` in float64; a source
  without a rig pose gets its camera extrinsics back unchanged (with a warning).
* ``CameraRig`` (:73-520): blocking poll of every source in construction order,
  one bounded queue per source (``queue_size`` newest frame sets), reference
  timestamp = min over sources of their newest timestamp, per-source pick =
  first queue entry with the smallest ``|ts - ref|``, nothing is consumed,
  ``None`` (not an exception) while stopped or while any queue is empty.

The queue itself is a fixed-capacity ring (``_Ring``) rather than a deque so the
GPU subclass (:class:`thor_slam_b200.ingest.rig.IngestRig`) can tie every slot
to a pinned host buffer / device buffer pair that is reused for the whole run.
"""

from __future__ import annotations

import logging
from dataclasses import dataclass, field
from threading import Lock
from types import TracebackType
from typing import Generic, Iterator, Sequence, TypeVar

import numpy as np

try:
    from typing import Self
except ImportError:  # pragma: no cover
    from typing_extensions import Self

from thor_slam_b200.camera.calibration import Extrinsics, IMUExtrinsics, Intrinsics
from thor_slam_b200.camera.frames import CameraSource, FrameSet, SynchronizedFrameSet

logger = logging.getLogger(__name__)

T = TypeVar("T")


class _Ring(Generic[T]):
    """Fixed-capacity FIFO that overwrites its oldest entry (== ``deque(maxlen=n)``).

    ``slot_of_newest`` exposes the physical slot index so a parallel array of
    pinned buffers can be indexed with it.
    """

    __slots__ = ("_items", "_head", "_count", "capacity")

    def __init__(self, capacity: int) -> None:
        if capacity <= 0:
            raise ValueError("ring capacity must be positive")
        self.capacity = capacity
        self._items: list[T | None] = [None] * capacity
        self._head = 0  # physical index of the oldest entry
        self._count = 0

    def push(self, item: T) -> int:
        """Append; returns the physical slot that now holds ``item``."""
        slot = (self._head + self._count) % self.capacity
        if self._count == self.capacity:
            slot = self._head
            self._head = (self._head + 1) % self.capacity
        else:
            self._count += 1
        self._items[slot] = item
        return slot

    def next_slot(self) -> int:
        """Physical slot the next ``push`` will write."""
        return self._head if self._count == self.capacity else (self._head + self._count) % self.capacity

    def pop_oldest(self) -> T:
        if not self._count:
            raise IndexError("pop from empty ring")
        item = self._items[self._head]
        self._items[self._head] = None
        self._head = (self._head + 1) % self.capacity
        self._count -= 1
        return item  # type: ignore[return-value]

    def clear(self) -> None:
        self._items = [None] * self.capacity
        self._head = 0
        self._count = 0

    def __len__(self) -> int:
        return self._count

    def __bool__(self) -> bool:
        return self._count > 0

    def __iter__(self) -> Iterator[T]:  # oldest -> newest, like a deque
        for i in range(self._count):
            yield self._items[(self._head + i) % self.capacity]  # type: ignore[misc]

    def __getitem__(self, i: int) -> T:
        if i < 0:
            i += self._count
        if not 0 <= i < self._count:
            raise IndexError("ring index out of range")
        return self._items[(self._head + i) % self.capacity]  # type: ignore[return-value]

    @property
    def slot_of_newest(self) -> int:
        return (self._head + self._count - 1) % self.capacity


@dataclass
class RigCalibration:
    """Everything the consumers need to place every stream in the rig frame."""

    intrinsics: dict[str, list[Intrinsics]]
    extrinsics: dict[str, list[Extrinsics]]
    source_names: list[str] = field(default_factory=list)
    rig_extrinsics: dict[str, Extrinsics] = field(default_factory=dict)
    imu_extrinsics: IMUExtrinsics | None = None

    def get_world_extrinsics(self, source_name: str) -> list[Extrinsics] | None:
        """``world_T_camera`` for every stream of ``source_name`` (``None`` if unknown)."""
        cams = self.extrinsics.get(source_name)
        if cams is None:
            return None
        rig_pose = self.rig_extrinsics.get(source_name)
        if rig_pose is None:
            logger.warning("No rig extrinsics defined for source %s, returning camera extrinsics as-is", source_name)
            return cams
        world_T_source = rig_pose.to_4x4_matrix()
        return [Extrinsics.from_4x4_matrix(world_T_source @ c.to_4x4_matrix()) for c in cams]


def _identity_imu(source: str | None) -> IMUExtrinsics:
    return IMUExtrinsics(source_name=source if source is not None else "", extrinsics=Extrinsics.from_4x4_matrix(np.eye(4)))


class CameraRig:
    """Synchronises several :class:`CameraSource` plugins into frame sets."""

    def __init__(
        self,
        sources: Sequence[CameraSource],
        queue_size: int = 30,
        rig_extrinsics: dict[str, Extrinsics] | None = None,
        imu_extrinsics: IMUExtrinsics | None = None,
        imu_source: str | None = None,
    ) -> None:
        self.sources: dict[str, CameraSource] = {s.name: s for s in sources}
        self.queue_size = queue_size
        self._frame_queues: dict[str, _Ring[FrameSet]] = {n: _Ring(queue_size) for n in self.sources}
        self._imu_queue: _Ring[tuple[float, dict]] = _Ring(queue_size)
        self._lock = Lock()
        self._running = False
        self._imu_source = imu_source

        if imu_source is not None:
            if imu_source not in self.sources:
                raise ValueError(
                    f"IMU source '{imu_source}' not found in sources. Available sources: {list(self.sources.keys())}"
                )
            if not self.sources[imu_source].has_sensor_data:
                raise ValueError(
                    f"IMU source '{imu_source}' does not have sensor data enabled. "
                    "Set read_imu=True when creating the camera source."
                )
            logger.info("Using '%s' as IMU source", imu_source)

        if not rig_extrinsics:
            logger.warning("No rig extrinsics provided, using identity transformation for all sources")
            rig_extrinsics = {n: Extrinsics.from_4x4_matrix(np.eye(4)) for n in self.sources}
        if not imu_extrinsics:
            logger.warning("No imu extrinsics provided, using identity transformation for the IMU")
            imu_extrinsics = _identity_imu(imu_source)

        self._calibration = self._build_calibration(rig_extrinsics, imu_extrinsics)

    # -- lifecycle ---------------------------------------------------------
    def __enter__(self) -> Self:
        self.start()
        return self

    def __exit__(
        self,
        exc_type: type[BaseException] | None,
        exc_val: BaseException | None,
        exc_tb: TracebackType | None,
    ) -> None:
        self.stop()

    def start(self) -> None:
        if self._running:
            return
        for s in self.sources.values():
            s.start()
        self._running = True

    def stop(self) -> None:
        if not self._running:
            return
        for s in self.sources.values():
            s.stop()
        self._running = False
        self.clear_queues()

    def is_running(self) -> bool:
        return self._running

    # -- calibration -------------------------------------------------------
    def _build_calibration(self, rig_extrinsics: dict[str, Extrinsics], imu_extrinsics: IMUExtrinsics) -> RigCalibration:
        return RigCalibration(
            intrinsics={n: s.get_intrinsics() for n, s in self.sources.items()},
            extrinsics={n: s.get_extrinsics() for n, s in self.sources.items()},
            source_names=list(self.sources),
            rig_extrinsics=rig_extrinsics,
            imu_extrinsics=imu_extrinsics,
        )

    @property
    def calibration(self) -> RigCalibration:
        return self._calibration

    def load_rig_extrinsics(
        self, rig_extrinsics: dict[str, Extrinsics], imu_extrinsics: IMUExtrinsics | None = None
    ) -> None:
        """Replace/extend the per-source rig poses; unknown source names raise ValueError."""
        for n in rig_extrinsics:
            if n not in self.sources:
                raise ValueError(f"Unknown source: {n}")
        merged = dict(self._calibration.rig_extrinsics)
        merged.update(rig_extrinsics)
        imu = imu_extrinsics if imu_extrinsics is not None else (self._calibration.imu_extrinsics or _identity_imu(self._imu_source))
        self._calibration = self._build_calibration(merged, imu)
        self._on_calibration_changed()

    def _on_calibration_changed(self) -> None:
        """Hook for subclasses that keep device-side copies of the calibration."""

    def get_rig_extrinsics(self, source_name: str) -> Extrinsics | None:
        return self._calibration.rig_extrinsics.get(source_name)

    def get_world_extrinsics(self, source_name: str) -> list[Extrinsics] | None:
        return self._calibration.get_world_extrinsics(source_name)

    # -- polling and matching ----------------------------------------------
    def _wrap_frames(self, name: str, frames: list) -> FrameSet:
        """Turn one driver read into a queue entry (subclasses stage to pinned memory here)."""
        return FrameSet.from_frames(frames, source_name=name)

    def _poll_cameras(self) -> None:
        for name, source in self.sources.items():
            if name == self._imu_source:
                data, ts = source.try_get_timestamped_sensor_data()
                if data is not None and ts is not None:
                    self._imu_queue.push((ts, data))
            frames = source.get_latest_frames()
            if frames:
                entry = self._wrap_frames(name, frames)
                with self._lock:
                    self._frame_queues[name].push(entry)

    @staticmethod
    def _find_closest_frame_set(queue: "_Ring[FrameSet]", target_timestamp: float) -> FrameSet | None:
        best: FrameSet | None = None
        best_dt = float("inf")
        for fs in queue:  # strict '<' keeps the first (oldest) minimum, as Python's min() does
            dt = abs(fs.timestamp - target_timestamp)
            if dt < best_dt:
                best, best_dt = fs, dt
        return best

    @staticmethod
    def _find_closest_imu_data(
        queue: "_Ring[tuple[float, dict]]", target_timestamp: float
    ) -> tuple[float | None, dict | None]:
        best: tuple[float, dict] | None = None
        best_dt = float("inf")
        for item in queue:
            dt = abs(item[0] - target_timestamp)
            if dt < best_dt:
                best, best_dt = item, dt
        return (None, None) if best is None else (best[0], best[1])

    def _get_reference_timestamp(self) -> float | None:
        with self._lock:
            newest = []
            for q in self._frame_queues.values():
                if not q:
                    return None
                newest.append(q[-1].timestamp)
        return min(newest)

    def get_synchronized_frames(self, max_wait_ms: float = 100.0) -> SynchronizedFrameSet | None:
        """Poll every source once and return the best-matched frame set (or ``None``)."""
        if not self._running:
            return None
        self._poll_cameras()
        ref = self._get_reference_timestamp()
        if ref is None:
            logger.warning("No reference timestamp found, not all cameras have frames yet")
            return None

        chosen: dict[str, FrameSet] = {}
        worst = 0.0
        with self._lock:
            for name, q in self._frame_queues.items():
                fs = self._find_closest_frame_set(q, ref)
                if fs is None:
                    return None
                chosen[name] = fs
                worst = max(worst, abs(fs.timestamp - ref))

        sensor_data: dict | None = None
        sensor_ts: float | None = None
        if self._imu_source is not None:
            ts, data = self._find_closest_imu_data(self._imu_queue, ref)
            if data is not None:
                sensor_data, sensor_ts = data, ts

        return self._finish(
            SynchronizedFrameSet(
                timestamp=ref,
                frame_sets=chosen,
                max_time_delta=worst,
                sensor_data=sensor_data,
                sensor_timestamp=sensor_ts,
            )
        )

    def get_latest_frames(self) -> SynchronizedFrameSet | None:
        """Newest frame set of every source, no matching; reference ts = newest of them."""
        if not self._running:
            return None
        self._poll_cameras()
        latest: dict[str, FrameSet] = {}
        with self._lock:
            for name, q in self._frame_queues.items():
                if not q:
                    logger.warning("Camera %s has no frames yet", name)
                    return None
                latest[name] = q[-1]
        stamps = [fs.timestamp for fs in latest.values()]
        ref = max(stamps) if stamps else 0.0
        spread = max(stamps) - min(stamps) if stamps else 0.0

        sensor_data: dict | None = None
        sensor_ts: float | None = None
        if self._imu_source is not None and self._imu_queue:
            ts, data = self._imu_queue[-1]
            if data is not None:
                sensor_data, sensor_ts = data, ts
        return self._finish(
            SynchronizedFrameSet(
                timestamp=ref,
                frame_sets=latest,
                max_time_delta=spread,
                sensor_data=sensor_data,
                sensor_timestamp=sensor_ts,
            )
        )

    def _finish(self, sync: SynchronizedFrameSet) -> SynchronizedFrameSet:
        """Last step before a frame set leaves the rig (the GPU subclass ingests here)."""
        return sync

    # -- introspection -----------------------------------------------------
    def get_source_names(self) -> list[str]:
        return list(self.sources)

    def get_source(self, name: str) -> CameraSource | None:
        return self.sources.get(name)

    def clear_queues(self) -> None:
        with self._lock:
            for q in self._frame_queues.values():
                q.clear()

    def get_queue_depths(self) -> dict[str, int]:
        with self._lock:
            return {n: len(q) for n, q in self._frame_queues.items()}

    def prune_old_frames(self, max_age_seconds: float = 1.0) -> int:
        """Drop queue entries older than ``newest - max_age_seconds``; returns how many."""
        with self._lock:
            newest = max((q[-1].timestamp for q in self._frame_queues.values() if q), default=None)
            if newest is None:
                return 0
            cutoff = newest - max_age_seconds
            dropped = 0
            for q in self._frame_queues.values():
                while q and q[0].timestamp < cutoff:
                    q.pop_oldest()
                    dropped += 1
        return dropped
