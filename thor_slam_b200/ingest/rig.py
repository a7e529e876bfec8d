"""``IngestRig`` - the drop-in behind ``CameraRig.get_synchronized_frames()``.

Same constructor, same methods, same ``None``/exception behaviour as the reference's ``CameraRig``
(``thor_slam/camera/rig.py:73-520`` - it *is* a subclass of our API mirror), but every frame set
that leaves the rig has been through the GPU ingest stage:

* SLAM streams: format conversion (mono8 pass-through, BGR8 -> rgb8, NV12 -> mono8 / rgb8) fused with
  the stereo-rectification / undistortion remap (what the reference leaves to cuVSLAM by publishing
  raw images with ``rectified_images:=false``);
* RGB-D sources (duck-typed like ``LuxonisCameraSource``, ``drivers/luxonis.py:871-1091``): BGR8 ->
  rgb8, and depth (u16 mm) -> body-frame FLU point cloud + valid mask + valid count (what the
  reference leaves to nvblox).

What changes relative to the reference (BASELINE.json north_star, "subsystems that change"):

* frame-set assembly: the per-source ``deque(maxlen=queue_size)`` (``rig.py:113``) becomes a ring of
  ``queue_size`` pinned host slots paired with ``queue_size`` device slots per stream; a driver read
  is copied once into its pinned slot and uploaded asynchronously on a copy stream, so by the time a
  frame set is matched its pixels are already in HBM;
* calibration: ``Intrinsics`` / ``Extrinsics`` (``camera/types.py``) are turned once into remap LUTs
  and 3x4 body transforms on the device (re-done on ``load_rig_extrinsics``).

``CameraFrame.image`` of a returned frame is a :class:`DeviceImage` (ndarray-compatible; the device -> host copy into the slot's
pinned mirror is enqueued before the frame set leaves the rig, ``np.asarray(image)`` only waits for it);
``SynchronizedFrameSet.clouds`` maps source name -> ``{"points", "mask", "count"}``.

Slot discipline (what replaces the reference's "every frame is a fresh array"): a slot's device input is re-staged only after
the ingest kernels that read it have finished (an event per slot), its outputs stay valid for ``queue_size`` polls, and a
:class:`DeviceImage` refuses to hand out pixels of a slot that has been re-used (``frames.py``).

Calibration of what is returned: with ``rectify=True`` the pixels are rectified and undistorted, so the matching calibration
is ``rectified_intrinsics(name)`` / ``rectification(name)`` (``D = 0``, ``K = P[:3,:3]``, ``R``, ``P`` - what a consumer
publishes with ``rectified_images:=true``); ``calibration`` keeps describing the raw cameras, as in the reference.
"""

from __future__ import annotations

import logging
from dataclasses import dataclass, field
from typing import Any, Sequence

import numpy as np
import torch

from thor_slam_b200.camera.calibration import Extrinsics, IMUExtrinsics, Intrinsics
from thor_slam_b200.camera.frames import CameraFrame, CameraSource, DeviceImage, FrameSet, SynchronizedFrameSet
from thor_slam_b200.camera.rig import CameraRig
from thor_slam_b200.ingest import formats as F
from thor_slam_b200.ingest.calib import body_T_camera, rotate_imu_sample, stereo_rectify_maps
from thor_slam_b200.ingest.context import IngestContext, StreamSpec

logger = logging.getLogger(__name__)

_TORCH_DTYPE = {F.MONO8: torch.uint8, F.BGR8: torch.uint8, F.RGB8: torch.uint8, F.NV12: torch.uint8,
                F.DEPTH16: torch.uint16, F.XYZ32F: torch.float32}


@dataclass
class _Stream:
    """One stream of one source: its formats, calibration slot and slot rings."""

    source: str
    index: int            # position in the source's frame list (0 left / 1 right / 0 rgb)
    camera: int           # calibration slot in the ingest context
    src_format: int
    dst_format: int
    src_size: tuple[int, int]   # (w, h)
    dst_size: tuple[int, int]
    kind: int
    host: Any = None      # [queue_size, ...] pinned
    dev: Any = None       # [queue_size, ...] device
    out: Any = None       # [queue_size, ...] device, ingest output
    out_host: Any = None  # [queue_size, ...] pinned mirror of `out`
    mask: Any = None
    count: Any = None
    gen: list = field(default_factory=list)        # per slot: how many times it has been staged
    up_plan: list = field(default_factory=list)    # per slot: (device address, pinned address, bytes) of the upload
    down_plan: list = field(default_factory=list)  # per slot: the download of the ingest output into its pinned mirror


@dataclass
class _SlotFrameSet(FrameSet):
    """Queue entry that remembers which ring slot holds its pixels."""

    slot: int = -1
    uploaded: Any = field(default=None, repr=False)  # event: H2D of this slot finished
    gen: int = 0                                      # staging generation of the slot when this entry was made
    result: Any = field(default=None, repr=False)    # (ready event, mirror event): the slot has been ingested already


class IngestRig(CameraRig):
    def __init__(
        self,
        sources: Sequence[CameraSource],
        queue_size: int = 30,
        rig_extrinsics: dict[str, Extrinsics] | None = None,
        imu_extrinsics: IMUExtrinsics | None = None,
        imu_source: str | None = None,
        *,
        device: int = 0,
        rectify: bool = True,
        rig_frame: str = "rdf",
        color_output: str = "rgb8",
        imu_frame: str = "rdf",
        context: IngestContext | None = None,
    ) -> None:
        """Extra (keyword-only) arguments over the reference constructor:

        ``device``: CUDA device index; ``rectify``: remap SLAM streams (False = conversion only);
        ``rig_frame``: ``"rdf"`` (rig poses in the Luxonis convention, clouds rotated to FLU with
        ``RDF_TO_FLU_MATRIX``) or ``"flu"``; ``color_output``: ``"rgb8"`` (what the reference's adapter
        publishes for 3-channel frames) or ``"mono8"``; ``imu_frame``: axes of the IMU source's samples, ``"drb"``
        (OAK-D Pro: samples are rotated into the camera's RDF axes, see ``calib.rotate_imu_sample``) or ``"rdf"``
        (OAK-D Long Range, reference behaviour: untouched); ``context``: share an ``IngestContext``.
        """
        super().__init__(sources, queue_size, rig_extrinsics, imu_extrinsics, imu_source)
        self._ctx = context if context is not None else IngestContext(device)
        self._emulated = self._ctx.lib.is_emulation
        self._device = torch.device("cpu") if self._emulated else torch.device("cuda", device)
        self._rectify = rectify
        self._rig_frame = rig_frame
        self._color_output = F.fmt(color_output)
        if imu_frame not in ("rdf", "drb"):
            raise ValueError(f"unknown IMU frame {imu_frame!r} (expected 'drb' or 'rdf')")
        self._imu_frame = imu_frame
        self._streams: dict[str, list[_Stream]] = {}
        self._rgbd: dict[str, tuple[_Stream, _Stream]] = {}
        self._colours: dict[str, Any] = {}
        self._next_camera = 0
        self._copy_stream = None if self._emulated else torch.cuda.Stream(device=self._device)
        self._down_stream = None if self._emulated else torch.cuda.Stream(device=self._device)
        self._copy_handle = None if self._emulated else self._copy_stream.cuda_stream   # raw cudaStream_t of the two
        self._down_handle = None if self._emulated else self._down_stream.cuda_stream
        self._slot_ingested: dict[tuple[str, int], Any] = {}   # (source, slot) -> event: the kernels that read the slot are done
        self._prepared: dict[tuple, Any] = {}                  # slots of a frame set -> packed ti_stream array
        self._rect: dict[str, list[dict]] = {}                 # per source, per stream: {"R", "P"} of the rectification
        self._build_streams()
        self._upload_calibration()

    # -- setup -----------------------------------------------------------------
    def _alloc(self, shape: tuple[int, ...], fmt: int, pinned: bool = False) -> Any:
        dt = _TORCH_DTYPE[fmt]
        if pinned:
            t = torch.zeros(shape, dtype=dt)
            return t if self._emulated else t.pin_memory()
        return torch.zeros(shape, dtype=dt, device=self._device)

    def _new_camera(self) -> int:
        cam = self._next_camera
        self._next_camera += 1
        return cam

    def _build_streams(self) -> None:
        q = self.queue_size
        for name, src in self.sources.items():
            intr = self._calibration.intrinsics[name]
            fmts = getattr(src, "get_stream_formats", None)
            names = fmts() if callable(fmts) else [None] * len(intr)
            streams = []
            for i, (it, fname) in enumerate(zip(intr, names)):
                sfmt = F.fmt(fname) if fname is not None else None
                size = (int(it.width), int(it.height))
                st = _Stream(name, i, self._new_camera(), sfmt if sfmt is not None else -1, -1, size, size,
                             F.KIND_RECTIFY if self._rectify else F.KIND_CONVERT)
                streams.append(st)
            self._streams[name] = streams
            if getattr(src, "has_rgbd_streams", False):
                ri, di = src.get_rgbd_intrinsics()
                rgb = _Stream(name, 0, self._new_camera(), F.BGR8, F.RGB8, (ri.width, ri.height), (ri.width, ri.height), F.KIND_CONVERT)
                dep = _Stream(name, 1, self._new_camera(), F.DEPTH16, F.XYZ32F, (di.width, di.height), (di.width, di.height), F.KIND_BACKPROJECT)
                for st in (rgb, dep):
                    self._alloc_stream(st, 2)
                self._rgbd[name] = (rgb, dep)
        del q

    def _alloc_stream(self, st: _Stream, slots: int) -> None:
        w, h = st.src_size
        dw, dh = st.dst_size
        st.host = self._alloc((slots, *F.frame_shape(st.src_format, w, h)), st.src_format, pinned=True)
        st.dev = self._alloc((slots, *F.frame_shape(st.src_format, w, h)), st.src_format)
        st.out = self._alloc((slots, *F.frame_shape(st.dst_format, dw, dh)), st.dst_format)
        st.out_host = self._alloc((slots, *F.frame_shape(st.dst_format, dw, dh)), st.dst_format, pinned=True)
        st.gen = [0] * slots
        # the staging copies of every slot, resolved once: (dst address, src address, bytes)
        st.up_plan = [self._ctx.copy_plan(st.dev[i], st.host[i]) for i in range(slots)]
        st.down_plan = [self._ctx.copy_plan(st.out_host[i], st.out[i]) for i in range(slots)]
        if st.kind == F.KIND_BACKPROJECT:
            st.mask = self._alloc((slots, dh, dw), F.MONO8)
            st.count = torch.zeros((slots,), dtype=torch.int32, device=self._device)

    def _resolve_format(self, st: _Stream, image: np.ndarray) -> None:
        """First frame of a stream: fix its wire format (the reference decides by ndim, isaac_ros.py:351-358)."""
        if st.src_format < 0:
            st.src_format = F.infer_format(image)
        if st.src_format == F.MONO8 or (st.src_format == F.NV12 and self._color_output == F.MONO8):
            st.dst_format = F.MONO8
        elif st.src_format == F.NV12:
            st.dst_format = self._color_output
        else:
            st.dst_format = self._color_output  # BGR8 -> rgb8 (reference) or mono8
        self._alloc_stream(st, self.queue_size)

    def _upload_calibration(self) -> None:
        cal = self._calibration
        for name, streams in self._streams.items():
            intr, extr = cal.intrinsics[name], cal.extrinsics[name]
            if self._rectify:
                size = streams[0].src_size
                maps, rp = stereo_rectify_maps(intr, extr, size, with_rp=True)
                self._rect[name] = rp
                for st, (mx, my) in zip(streams, maps):
                    self._ctx.upload_rectify_map(st.camera, mx, my, st.src_size)
                self._prepared.clear()
        for name, (_rgb, dep) in self._rgbd.items():
            src = self.sources[name]
            _ri, di = src.get_rgbd_intrinsics()
            _re, de = src.get_rgbd_extrinsics()
            rig_pose = cal.rig_extrinsics.get(name)
            m = body_T_camera(None if rig_pose is None else rig_pose.to_4x4_matrix(), de.to_4x4_matrix(), self._rig_frame)
            self._ctx.upload_projection(dep.camera, di.matrix, m, dep.src_size)
            # depth -> RGB registration: both extrinsics map into the source frame (CAM_A), luxonis.py:1068-1091
            rgb_T_depth = np.linalg.inv(_re.to_4x4_matrix()) @ de.to_4x4_matrix()
            self._ctx.upload_registration(dep.camera, di.matrix, dep.src_size, _ri.matrix, _rgb.src_size, rgb_T_depth)

    def _on_calibration_changed(self) -> None:
        if hasattr(self, "_streams"):
            self._upload_calibration()

    # -- frame-set assembly: pinned ring + async upload -------------------------------
    def _stage_host(self, st: _Stream, slot: int, image: np.ndarray) -> None:
        """CPU half of staging: one copy of the driver's frame into the slot's pinned buffer.  (torch's copy splits a frame
        over the process's OpenMP threads: 8 x 1 MB take 0.19 ms here; the same copies handed to a thread pool - measured - take
        0.9 ms, the pool's workers then fight over those threads.)"""
        host = st.host[slot]
        src = torch.from_numpy(np.ascontiguousarray(image))
        if src.dtype != host.dtype:
            src = src.view(host.dtype)
        if tuple(src.shape) != tuple(host.shape):
            raise ValueError(f"{st.source}[{st.index}]: frame shape {tuple(src.shape)} does not match the calibrated {tuple(host.shape)}")
        host.copy_(src)

    def _stage_upload(self, st: _Stream, slot: int) -> None:
        """GPU half: pinned slot -> device slot on the copy stream, behind the kernels that still read the device slot."""
        st.gen[slot] += 1
        if self._emulated:
            st.dev[slot].copy_(st.host[slot])
            return
        # through the library on the copy stream's raw handle: a torch stream context + copy_ costs 40 us per frame (measured,
        # tools/rig_profile.py), this call 2
        self._ctx.copy_planned(st.up_plan[slot], self._copy_handle)

    def _stage(self, st: _Stream, slot: int, image: np.ndarray) -> None:
        self._stage_host(st, slot, image)
        self._stage_upload(st, slot)

    def _wait_slot_free(self, name: str, slot: int) -> None:
        """Before a slot is re-staged: the ingest kernels that read its device buffer (and the download of its outputs) are done."""
        ev = self._slot_ingested.pop((name, slot), None)
        if ev is not None and not self._emulated:
            ready, mirror = ev
            self._copy_stream.wait_event(ready)  # write-after-read on st.dev[slot]
            mirror.synchronize()                 # the pinned input / output mirrors of the slot are host-visible state

    def _wrap_frames(self, name: str, frames: list) -> FrameSet:
        streams = self._streams[name]
        slot = self._frame_queues[name].next_slot()
        self._wait_slot_free(name, slot)
        for st, fr in zip(streams, frames):
            if st.host is None:
                self._resolve_format(st, fr.image)
        staged = []
        for st, fr in zip(streams, frames):
            self._stage_host(st, slot, np.asarray(fr.image))
            self._stage_upload(st, slot)  # runs while the next stream's frame is being copied into its pinned slot
            view = st.host[slot].numpy() if st.host.dtype != torch.uint16 else st.host[slot].view(torch.int16).numpy().view(np.uint16)
            staged.append(CameraFrame(view, fr.timestamp, fr.sequence_num, fr.camera_name))
        ev = None
        if not self._emulated:
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        return _SlotFrameSet(timestamp=staged[0].timestamp, frames=staged, source_name=name, slot=slot, uploaded=ev, gen=streams[0].gen[slot])

    # -- the ingest stage itself ------------------------------------------------------
    def _finish(self, sync: SynchronizedFrameSet) -> SynchronizedFrameSet:
        todo: list[tuple[str, Any]] = []
        for name, fs in sync.frame_sets.items():
            if getattr(fs, "slot", -1) < 0:
                return sync  # not one of ours (e.g. injected by a test): pass through untouched
            if fs.result is None:  # nothing is popped from the queues: a frame set can be matched twice, it is ingested once
                todo.append((name, fs))
        if todo:
            key = tuple((name, fs.slot) for name, fs in todo)
            prep = self._prepared.get(key)
            if prep is None:
                specs = [StreamSpec(st.kind, st.dev[fs.slot:fs.slot + 1], st.out[fs.slot:fs.slot + 1], st.src_format, st.dst_format,
                                    camera=st.camera, width=st.src_size[0], height=st.src_size[1])
                         for name, fs in todo for st in self._streams[name]]
                prep = self._prepared[key] = self._ctx.prepare(specs)
            ready = mirror = None
            if not self._emulated:
                cur = torch.cuda.current_stream(self._device)
                self._ctx.set_stream(cur.cuda_stream)
                for _, fs in todo:
                    if fs.uploaded is not None:
                        cur.wait_event(fs.uploaded)
            self._ctx.ingest_prepared(prep)
            if not self._emulated:
                ready = torch.cuda.Event()
                ready.record(cur)
                self._down_stream.wait_event(ready)
                for name, fs in todo:  # outputs -> pinned mirrors on the download stream, under whatever the caller does next
                    for st in self._streams[name]:
                        self._ctx.copy_planned(st.down_plan[fs.slot], self._down_handle)
                mirror = torch.cuda.Event()
                mirror.record(self._down_stream)
            else:
                for name, fs in todo:
                    for st in self._streams[name]:
                        st.out_host[fs.slot].copy_(st.out[fs.slot])
            for name, fs in todo:
                fs.result = (ready, mirror)
                self._slot_ingested[(name, fs.slot)] = (ready, mirror)
        out_sets: dict[str, FrameSet] = {}
        for name, fs in sync.frame_sets.items():
            ready, mirror = fs.result
            frames = []
            for st, fr in zip(self._streams[name], fs.frames):
                guard = (lambda st=st, slot=fs.slot, gen=fs.gen: st.gen[slot] == gen)
                frames.append(CameraFrame(DeviceImage(st.out[fs.slot], ready, st.out_host[fs.slot], mirror, guard), fr.timestamp, fr.sequence_num, fr.camera_name))
            out_sets[name] = FrameSet(fs.timestamp, frames, fs.source_name, fs.sensor_data, fs.sensor_timestamp)
        return SynchronizedFrameSet(sync.timestamp, out_sets, sync.max_time_delta, rotate_imu_sample(sync.sensor_data, self._imu_frame),
                                    sync.sensor_timestamp)

    # -- calibration of the RETURNED pixels ----------------------------------------------------------
    def rectification(self, source_name: str) -> list[dict] | None:
        """Per stream of ``source_name``: ``{"R": 3x3, "P": 3x4}`` of the stereo rectification the returned pixels went through
        (``cv2.stereoRectify`` with ``CALIB_ZERO_DISPARITY``, ``alpha=0``), or ``None`` with ``rectify=False``."""
        return self._rect.get(source_name) if self._rectify else None

    def rectified_intrinsics(self, source_name: str) -> list[Intrinsics]:
        """``Intrinsics`` that describe what ``get_synchronized_frames()`` returns: ``K = P[:3,:3]`` and no distortion when the rig
        rectifies (``CameraInfo`` with ``D = 0``, ``R``, ``P`` and ``rectified_images:=true`` - the rectified counterpart of
        ``isaac_ros.py:364-411``), the raw intrinsics otherwise."""
        raw = self._calibration.intrinsics[source_name]
        if not self._rectify:
            return list(raw)
        return [Intrinsics(width=it.width, height=it.height, matrix=np.array(rp["P"])[:3, :3].copy(), coeffs=np.zeros(5))
                for it, rp in zip(raw, self._rect[source_name])]

    # -- RGB-D (bypasses the synchroniser in the reference too: run_pipeline.py:624-631) --------
    def get_rgbd(self, source_name: str, blocking: bool = False) -> dict | None:
        """Latest RGB-D pair of ``source_name`` through the ingest stage.

        Returns ``{"rgb": CameraFrame(rgb8), "depth": CameraFrame(u16 mm, untouched), "points": DeviceImage
        HxWx3 f32 body frame, "colours": DeviceImage HxWx3 u8 (the RGB pixel each depth pixel projects to),
        "mask": DeviceImage HxW u8, "count": int tensor}`` or ``None`` when the
        source has no new pair (or is not an RGB-D source / the rig is stopped).
        """
        if not self._running or source_name not in self._rgbd:
            return None
        src = self.sources[source_name]
        pair = src.get_latest_rgbd_frames() if blocking else src.try_get_latest_rgbd_frames()
        if pair is None:
            return None
        rgb_f, dep_f = pair
        rgb, dep = self._rgbd[source_name]
        slot = int(rgb_f.sequence_num) % 2
        self._wait_slot_free(source_name + "/rgbd", slot)
        self._stage(rgb, slot, np.asarray(rgb_f.image))
        self._stage(dep, slot, np.asarray(dep_f.image))
        if not self._emulated:
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
            cur = torch.cuda.current_stream(self._device)
            cur.wait_event(ev)
            self._ctx.set_stream(cur.cuda_stream)
        self._ctx.ingest([
            StreamSpec(F.KIND_CONVERT, rgb.dev[slot:slot + 1], rgb.out[slot:slot + 1], F.BGR8, F.RGB8, width=rgb.src_size[0], height=rgb.src_size[1]),
        ])
        colours = self._colours.setdefault(source_name, self._alloc((2, dep.src_size[1], dep.src_size[0], 3), F.RGB8))
        # cloud + mask + count + the colour of every depth pixel in one pass over the depth image (ti_backproject_colour)
        self._ctx.backproject_colour(dep.camera, dep.dev[slot:slot + 1], rgb.out[slot:slot + 1], dep.out[slot:slot + 1], colours[slot:slot + 1],
                                     dep.mask[slot:slot + 1], dep.count[slot:slot + 1])
        ready = None
        if not self._emulated:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(self._device))
            self._slot_ingested[(source_name + "/rgbd", slot)] = (ready, ready)
        return {
            "rgb": CameraFrame(DeviceImage(rgb.out[slot], ready), rgb_f.timestamp, rgb_f.sequence_num, rgb_f.camera_name),
            "depth": dep_f,
            "colours": DeviceImage(colours[slot], ready),
            "points": DeviceImage(dep.out[slot], ready),
            "mask": DeviceImage(dep.mask[slot], ready),
            "count": dep.count[slot],
        }

    def get_synchronized_frames(self, max_wait_ms: float = 100.0, with_clouds: bool = False) -> SynchronizedFrameSet | None:
        """Reference signature plus ``with_clouds``: attach the RGB-D clouds of every RGB-D source."""
        sync = super().get_synchronized_frames(max_wait_ms)
        if sync is not None and with_clouds and self._rgbd:
            clouds = {}
            for name in self._rgbd:
                got = self.get_rgbd(name)
                if got is not None:
                    clouds[name] = got
            sync.clouds = clouds or None
        return sync

    @property
    def ingest_context(self) -> IngestContext:
        return self._ctx

    def stop(self) -> None:
        super().stop()
        self._slot_ingested.clear()
        if not self._emulated:
            torch.cuda.synchronize(self._device)
