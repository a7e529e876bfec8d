"""``IngestContext`` - thin Python face of one ``ti_ctx`` (one per process and GPU).

PyTorch tensors are only buffer carriers here: every call hands raw ``data_ptr()`` addresses to
the C ABI.  Error mapping follows the reference's conventions:
``TI_EINVAL`` -> ``ValueError``; everything else -> ``RuntimeError``.
"""

from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Any, Sequence

import numpy as np

from thor_slam_b200.ingest import formats as F
from thor_slam_b200.ingest._lib import TI_EINVAL, TI_OK, IngestLibrary, TiDepthStream, TiStream, default_library


def _is_torch(x: Any) -> bool:
    return hasattr(x, "data_ptr")


@dataclass
class StreamSpec:
    """One stream of a frame-set batch (Python view of ``struct ti_stream``).

    ``src`` / ``dst`` / ``mask`` / ``count`` are batched arrays whose first dimension is the
    frame index; strides default to the tensors' own batch stride.
    """

    kind: int
    src: Any
    dst: Any
    src_format: int
    dst_format: int
    camera: int = 0
    width: int = 0
    height: int = 0
    mask: Any = None
    count: Any = None


class IngestContext:
    def __init__(self, device: int = 0, library: IngestLibrary | None = None) -> None:
        self.lib = library if library is not None else default_library()
        self.device = device
        handle = C.c_void_p()
        rc = self.lib.ti_create(device, C.byref(handle))
        if rc != TI_OK:
            msg = self.lib.ti_last_error(None).decode()
            raise (ValueError if rc == TI_EINVAL else RuntimeError)(msg)
        self._h = handle
        self._cams: dict[int, dict] = {}

    # -- plumbing ------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.ti_destroy(self._h)
            self._h = None

    def __del__(self) -> None:  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self) -> "IngestContext":
        return self

    def __exit__(self, *exc: object) -> None:
        self.close()

    def _check(self, rc: int) -> None:
        if rc == TI_OK:
            return
        msg = self.lib.ti_last_error(self._h).decode()
        raise (ValueError if rc == TI_EINVAL else RuntimeError)(f"thoringest: {msg}")

    def _ptr(self, x: Any, host_ok: bool = False) -> int | None:
        """Raw address of a buffer carrier; device tensors only, unless ``host_ok``."""
        if x is None:
            return None
        if _is_torch(x):
            if not x.is_contiguous():
                raise ValueError("buffers handed to the ingest library must be contiguous")
            if not x.is_cuda and not (host_ok or self.lib.is_emulation):
                raise ValueError("expected a CUDA tensor (the ingest stage has no CPU path)")
            return x.data_ptr()
        if isinstance(x, np.ndarray):
            if not (host_ok or self.lib.is_emulation):
                raise ValueError("numpy arrays are host memory; pass a CUDA tensor")
            if not x.flags["C_CONTIGUOUS"]:
                raise ValueError("buffers handed to the ingest library must be contiguous")
            return x.ctypes.data
        if isinstance(x, int):
            return x
        raise TypeError(f"unsupported buffer carrier {type(x).__name__}")

    @staticmethod
    def _batch_stride(x: Any) -> int:
        """Bytes between consecutive frames (dimension 0) of a batched array."""
        if _is_torch(x):
            return x.stride(0) * x.element_size() if x.dim() > 0 and x.shape[0] > 1 else math.prod(x.shape[1:]) * x.element_size()
        return x.strides[0] if x.ndim > 0 and x.shape[0] > 1 else math.prod(x.shape[1:]) * x.itemsize

    def set_stream(self, cuda_stream: int | None) -> None:
        """``cuda_stream``: e.g. ``torch.cuda.current_stream().cuda_stream`` (``None``/0: legacy default)."""
        self._check(self.lib.ti_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def sync(self) -> None:
        self._check(self.lib.ti_sync(self._h))

    OPT_FORCE_GENERIC_RECTIFY = 1
    OPT_CTAS_PER_SM = 2
    OPT_MONO_VARIANT = 3
    OPT_TMA_TILE_H = 4
    OPT_DEBUG = 5
    OPT_FRAMES_PER_UNIT = 6
    OPT_STAGES = 7
    OPT_LUT_PREFETCH = 8

    def set_option(self, option: int, value: int) -> None:
        """Tuning / test switches of the library; results never depend on them."""
        self._check(self.lib.ti_set_option(self._h, option, value))

    @property
    def launch_count(self) -> int:
        return int(self.lib.ti_launch_count(self._h))

    @property
    def sm_count(self) -> int:
        return int(self.lib.ti_device_sm_count(self._h))

    # -- calibration upload ----------------------------------------------------
    def upload_rectify_map(self, camera: int, mapx: np.ndarray, mapy: np.ndarray, src_size: tuple[int, int]) -> None:
        """``mapx``/``mapy``: float32 ``dst_h x dst_w`` (OpenCV convention); ``src_size`` = (w, h)."""
        mapx = np.ascontiguousarray(mapx, dtype=np.float32)
        mapy = np.ascontiguousarray(mapy, dtype=np.float32)
        if mapx.shape != mapy.shape or mapx.ndim != 2:
            raise ValueError(f"mapx/mapy must be equal-shaped 2-D arrays, got {mapx.shape} and {mapy.shape}")
        dst_h, dst_w = mapx.shape
        self._check(
            self.lib.ti_upload_rectify_map(self._h, camera, dst_w, dst_h, int(src_size[0]), int(src_size[1]),
                                           mapx.ctypes.data, mapy.ctypes.data)
        )
        self._cams.setdefault(camera, {}).update(dst=(dst_w, dst_h), src=(int(src_size[0]), int(src_size[1])))

    def upload_projection(self, camera: int, k: np.ndarray, body_T_cam: np.ndarray, size: tuple[int, int]) -> None:
        """``k``: 3x3 intrinsics of the depth image; ``body_T_cam``: 4x4 (or 3x4) float64; ``size`` = (w, h)."""
        k = np.asarray(k, dtype=np.float64)
        m = np.asarray(body_T_cam, dtype=np.float64)
        if k.shape != (3, 3) or m.shape not in ((4, 4), (3, 4)):
            raise ValueError("k must be 3x3 and body_T_cam 4x4 or 3x4")
        kk = (C.c_double * 4)(k[0, 0], k[1, 1], k[0, 2], k[1, 2])
        mm = (C.c_double * 12)(*m[:3, :4].reshape(-1))
        self._check(self.lib.ti_upload_projection(self._h, camera, int(size[0]), int(size[1]), kk, mm))
        self._cams.setdefault(camera, {}).update(proj=(int(size[0]), int(size[1])))

    def upload_registration(self, camera: int, k_depth: np.ndarray, depth_size: tuple[int, int], k_rgb: np.ndarray,
                            rgb_size: tuple[int, int], rgb_T_depth: np.ndarray) -> None:
        """Depth -> RGB registration constants: 3x3 intrinsics of both images, sizes (w, h), 4x4 (or 3x4) ``rgb_T_depth``."""
        kd, kr = np.asarray(k_depth, dtype=np.float64), np.asarray(k_rgb, dtype=np.float64)
        m = np.asarray(rgb_T_depth, dtype=np.float64)
        if kd.shape != (3, 3) or kr.shape != (3, 3) or m.shape not in ((4, 4), (3, 4)):
            raise ValueError("intrinsics must be 3x3 and rgb_T_depth 4x4 or 3x4")
        self._check(self.lib.ti_upload_registration(
            self._h, camera, int(depth_size[0]), int(depth_size[1]), (C.c_double * 4)(kd[0, 0], kd[1, 1], kd[0, 2], kd[1, 2]),
            int(rgb_size[0]), int(rgb_size[1]), (C.c_double * 4)(kr[0, 0], kr[1, 1], kr[0, 2], kr[1, 2]),
            (C.c_double * 12)(*m[:3, :4].reshape(-1))))

    def register_colour(self, camera: int, depth: Any, rgb: Any, colour: Any) -> Any:
        """One RGB8 colour per depth pixel (``ti_register_colour``): ``depth`` [n,H,W] u16, ``rgb`` [n,Hr,Wr,3], ``colour`` [n,H,W,3]."""
        n = int(depth.shape[0])
        self._check(self.lib.ti_register_colour(self._h, camera, self._ptr(depth), self._ptr(rgb), self._ptr(colour), n,
                                                self._batch_stride(depth), self._batch_stride(rgb), self._batch_stride(colour)))
        return colour

    def camera_info(self, camera: int) -> dict:
        return dict(self._cams.get(camera, {}))

    def rectify_plan(self, camera: int) -> dict:
        """Which mono remap kernel slot ``camera`` runs under the current options (``ti_rectify_plan``)."""
        import ctypes

        out = (ctypes.c_int32 * 8)()
        self._check(self.lib.ti_rectify_plan(self._h, camera, out))
        return {"variant": int(out[0]), "tile_h": int(out[1]), "rows": int(out[2]), "exceptions_per_warp": int(out[3]),
                "colour_variant": int(out[4]), "colour_rows": int(out[5]), "overflow_pixels": int(out[6]), "pitch": int(out[7]) & 0xFFFF,
                "pixels_per_window": (4 if int(out[7]) >> 16 else 2) if int(out[0]) == 4 else 0}

    def get_valid_mask(self, camera: int, out: Any) -> Any:
        self._check(self.lib.ti_get_valid_mask(self._h, camera, self._ptr(out)))
        return out

    # -- per-frame calls (device buffers) ---------------------------------------
    def convert(self, src: Any, dst: Any, src_format: int | str, dst_format: int | str, width: int, height: int) -> Any:
        """Batched format conversion; ``src``/``dst``: ``[n_batch, ...frame shape...]``."""
        n = int(src.shape[0])
        self._check(
            self.lib.ti_convert(self._h, F.fmt(src_format), F.fmt(dst_format), self._ptr(src), self._ptr(dst), width, height,
                                n, self._batch_stride(src), self._batch_stride(dst))
        )
        return dst

    def rectify(self, camera: int, src: Any, dst: Any, src_format: int | str, dst_format: int | str) -> Any:
        n = int(src.shape[0])
        self._check(
            self.lib.ti_rectify(self._h, camera, F.fmt(src_format), F.fmt(dst_format), self._ptr(src), self._ptr(dst), n,
                                self._batch_stride(src), self._batch_stride(dst))
        )
        return dst

    def backproject(self, camera: int, depth: Any, xyz: Any, mask: Any = None, count: Any = None) -> Any:
        n = int(depth.shape[0])
        self._check(
            self.lib.ti_backproject(self._h, camera, self._ptr(depth), self._ptr(xyz), self._ptr(mask), self._ptr(count), n,
                                    self._batch_stride(depth), self._batch_stride(xyz),
                                    self._batch_stride(mask) if mask is not None else 0)
        )
        return xyz

    def depth_stats(self, depth: Any, stats: Any) -> Any:
        """``ti_depth_stats``: ``depth`` [n, H, W] u16 -> ``stats`` [n, 6] u32/i32 = (count, min, max, 0, sum lo, sum hi) of depth > 0."""
        n, h, w = (int(x) for x in depth.shape)
        self._check(self.lib.ti_depth_stats(self._h, self._ptr(depth), w, h, n, self._batch_stride(depth), self._ptr(stats)))
        return stats

    def backproject_colour(self, camera: int, depth: Any, rgb: Any, xyz: Any, colour: Any, mask: Any = None, count: Any = None) -> Any:
        """``ti_backproject_colour``: clouds + mask + count + one RGB8 colour per depth pixel in one pass over the depth image."""
        n = int(depth.shape[0])
        self._check(
            self.lib.ti_backproject_colour(self._h, camera, self._ptr(depth), self._ptr(rgb), self._ptr(xyz), self._ptr(mask), self._ptr(count),
                                           self._ptr(colour), n, self._batch_stride(depth), self._batch_stride(rgb), self._batch_stride(xyz),
                                           self._batch_stride(mask) if mask is not None else 0, self._batch_stride(colour))
        )
        return xyz

    def _pack(self, streams: Sequence[StreamSpec], host: bool) -> tuple[Any, int]:
        arr = (TiStream * len(streams))()
        n_batch = None
        for i, s in enumerate(streams):
            nb = int(s.src.shape[0])
            if n_batch is None:
                n_batch = nb
            elif nb != n_batch:
                raise ValueError(f"stream {i} carries {nb} frames, earlier streams carry {n_batch}")
            t = arr[i]
            t.kind, t.camera = s.kind, s.camera
            t.src_format, t.dst_format = F.fmt(s.src_format), F.fmt(s.dst_format)
            t.width, t.height = s.width, s.height
            t.src, t.dst = self._ptr(s.src, host), self._ptr(s.dst, host)
            t.src_frame_stride, t.dst_frame_stride = self._batch_stride(s.src), self._batch_stride(s.dst)
            t.mask = self._ptr(s.mask, host)
            t.mask_frame_stride = self._batch_stride(s.mask) if s.mask is not None else 0
            t.count = self._ptr(s.count, host)
        return arr, (n_batch or 0)

    def ingest(self, streams: Sequence[StreamSpec]) -> None:
        """Whole frame-set batch, device buffers, at most one launch per stream kind."""
        arr, n_batch = self._pack(streams, host=False)
        self._check(self.lib.ti_ingest(self._h, arr, len(streams), n_batch))

    def prepare(self, streams: Sequence[StreamSpec]) -> tuple[Any, int, int, tuple]:
        """Pack a stream list once; ``ingest_prepared`` then costs one foreign call (the live rig replays the same few slot
        combinations for ever).  The buffers named in ``streams`` are kept alive by the returned object."""
        arr, n_batch = self._pack(streams, host=False)
        return arr, len(streams), n_batch, tuple(streams)

    def ingest_prepared(self, prepared: tuple[Any, int, int, tuple]) -> None:
        self._check(self.lib.ti_ingest(self._h, prepared[0], prepared[1], prepared[2]))

    def ingest_host(self, streams: Sequence[StreamSpec], chunk: int = 8) -> None:
        """Same with (pinned) host buffers: H2D, kernels and D2H pipelined ``chunk`` frame sets at a time."""
        arr, n_batch = self._pack(streams, host=True)
        self._check(self.lib.ti_ingest_host(self._h, arr, len(streams), n_batch, chunk))

    def ingest_host_submit(self, streams: Sequence[StreamSpec], chunk: int = 8) -> int:
        """Enqueue a host-buffer batch behind earlier submissions and return its ticket at once.

        The buffers named in ``streams`` must stay alive and untouched until :meth:`ingest_host_wait` returns for the
        ticket.  A capture loop that double-buffers its host frames overlaps batch k+1's upload with batch k's download.
        """
        arr, n_batch = self._pack(streams, host=True)
        ticket = C.c_uint64(0)
        self._check(self.lib.ti_ingest_host_submit(self._h, arr, len(streams), n_batch, chunk, C.byref(ticket)))
        return int(ticket.value)

    def ingest_host_wait(self, ticket: int) -> None:
        """Block until the batch submitted under ``ticket`` has landed in its destination buffers."""
        self._check(self.lib.ti_ingest_host_wait(self._h, ticket))

    def copy_async(self, dst: Any, src: Any, stream: int | None = None) -> None:
        """One asynchronous copy between pinned-host / device tensors of equal byte size on CUDA stream ``stream`` (a raw handle;
        None = the context's stream) - ``ti_copy_async``, the rig's staging copies without a framework stream switch."""
        nbytes = src.numel() * src.element_size()
        if dst.numel() * dst.element_size() != nbytes:
            raise ValueError("copy_async: source and destination differ in size")
        self._check(self.lib.ti_copy_async(self._h, C.c_void_p(self._ptr(dst, host_ok=True)), C.c_void_p(self._ptr(src, host_ok=True)), nbytes,
                                           C.c_void_p(stream) if stream else None))

    def copy_plan(self, dst: Any, src: Any) -> tuple[int, int, int]:
        """(dst address, src address, bytes) of a copy that will be issued many times: :meth:`copy_planned` then costs one foreign call."""
        nbytes = src.numel() * src.element_size()
        if dst.numel() * dst.element_size() != nbytes:
            raise ValueError("copy_plan: source and destination differ in size")
        return int(self._ptr(dst, host_ok=True)), int(self._ptr(src, host_ok=True)), int(nbytes)

    def copy_planned(self, plan: tuple[int, int, int], stream: int | None = None) -> None:
        rc = self.lib.ti_copy_async(self._h, plan[0], plan[1], plan[2], stream)
        if rc != TI_OK:
            self._check(rc)

    # -- voxel down-sampled cloud ------------------------------------------------
    def set_voxel_grid(self, voxel_size_m: float = 0.05, max_depth_mm: int = 10000) -> None:
        """Grid of the rig-wide cloud; defaults are nvblox's (``launch/thor_nvblox.launch.py:26-31``). ``max_depth_mm`` 0 = no cap."""
        self._check(self.lib.ti_set_voxel_grid(self._h, float(voxel_size_m), int(max_depth_mm)))
        self.voxel_size = float(voxel_size_m)

    def voxel_cloud(self, depth_streams: Sequence[tuple[int, Any]], records: Any, n_records: Any, set_counts: Any = None,
                    set_base: int = 0, tag: int = 0) -> None:
        """``ti_voxel_cloud``: ``depth_streams`` = [(camera slot, depth [n_batch, H, W] u16), ...]; ``records``: u64/i64 [capacity];
        ``n_records``: u32/i32 [1]; ``set_counts``: u32/i32 [n_batch] or ``None``.  One record per occupied voxel per frame set."""
        arr = (TiDepthStream * max(1, len(depth_streams)))()
        n_batch = None
        for i, (cam, depth) in enumerate(depth_streams):
            nb = int(depth.shape[0])
            if n_batch is None:
                n_batch = nb
            elif nb != n_batch:
                raise ValueError(f"depth stream {i} carries {nb} frames, earlier streams carry {n_batch}")
            arr[i].camera, arr[i].depth, arr[i].depth_frame_stride = int(cam), self._ptr(depth), self._batch_stride(depth)
        capacity = int(records.shape[0]) if records is not None else 0
        self._check(self.lib.ti_voxel_cloud(self._h, arr, len(depth_streams), n_batch or 0, int(set_base), int(tag), self._ptr(records), capacity,
                                            self._ptr(n_records), self._ptr(set_counts)))

    def voxel_points(self, records: Any, n_records: Any, xyz: Any) -> Any:
        """``ti_voxel_points``: voxel centres ``[capacity, 3]`` f32 of the first ``*n_records`` records (``N x 3`` cloud contract)."""
        self._check(self.lib.ti_voxel_points(self._h, self._ptr(records), self._ptr(n_records), int(xyz.shape[0]), self._ptr(xyz)))
        return xyz

    # -- multi-GPU ---------------------------------------------------------------
    def nccl_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        rc = self.lib.ti_nccl_unique_id(buf)
        if rc != TI_OK:
            raise RuntimeError(self.lib.ti_last_error(None).decode())
        return buf.raw

    def nccl_init(self, unique_id: bytes, rank: int, world: int) -> None:
        if len(unique_id) != 128:
            raise ValueError("NCCL unique id must be 128 bytes")
        self._check(self.lib.ti_nccl_init(self._h, C.create_string_buffer(unique_id, 128), rank, world))

    def gather_clouds(self, local: Any, gathered: Any, bytes_per_rank: Sequence[int], root: int = 0) -> None:
        arr = (C.c_uint64 * len(bytes_per_rank))(*[int(b) for b in bytes_per_rank])
        self._check(self.lib.ti_gather_clouds(self._h, self._ptr(local), self._ptr(gathered), arr, root))

    def gather_wait(self, on_stream: bool = False) -> None:
        """Wait for the last exchange (``ti_gather_wait``): the ingest stream waits if ``on_stream``, else this thread blocks."""
        self._check(self.lib.ti_gather_wait(self._h, 1 if on_stream else 0))

    def gather_counts(self, n_local: Any, world: int) -> list[int]:
        """All ranks' ``*n_local`` (a device u32) on the host - sizes of a variable-length gather (``ti_gather_counts``)."""
        out = (C.c_uint32 * world)()
        self._check(self.lib.ti_gather_counts(self._h, self._ptr(n_local), out))
        return [int(x) for x in out]

    def gather_counts_begin(self, n_local: Any) -> None:
        self._check(self.lib.ti_gather_counts_begin(self._h, self._ptr(n_local)))

    def gather_counts_finish(self, world: int) -> list[int]:
        out = (C.c_uint32 * world)()
        self._check(self.lib.ti_gather_counts_finish(self._h, out))
        return [int(x) for x in out]

    def gather_records(self, records: Any, gathered: Any, counts: Sequence[int], root: int = 0) -> None:
        """Variable-length gather of voxel lists sized by ``gather_counts*`` (``ti_gather_records``), on the exchange stream."""
        arr = (C.c_uint32 * len(counts))(*[int(c) for c in counts])
        self._check(self.lib.ti_gather_records(self._h, self._ptr(records), self._ptr(gathered), arr, root))

    def exchange_fence(self) -> int:
        f = C.c_uint64(0)
        self._check(self.lib.ti_exchange_fence(self._h, C.byref(f)))
        return int(f.value)

    def exchange_wait(self, fence: int, on_stream: bool = False) -> None:
        self._check(self.lib.ti_exchange_wait(self._h, int(fence), 1 if on_stream else 0))

    def inbox_init(self, inbox: int) -> None:
        self._check(self.lib.ti_inbox_init(self._h, C.c_void_p(inbox)))

    def cloud_push(self, records: Any, n_records: Any, inbox: int, inbox_capacity: int, gen: int) -> None:
        """Append a ``ti_voxel_cloud`` list to a (peer-mapped) inbox with peer stores on the exchange stream (``ti_cloud_push``)."""
        self._check(self.lib.ti_cloud_push(self._h, self._ptr(records), self._ptr(n_records), int(records.shape[0]), C.c_void_p(inbox), int(inbox_capacity), int(gen)))

    def inbox_take(self, inbox: int, inbox_capacity: int, world: int, dst: Any, status: Any) -> None:
        """Root: take one generation of an inbox into ``dst`` (u64/i64 [capacity]); ``status`` u32/i32 [2] = (count, error)."""
        cap = int(dst.shape[0]) if dst is not None else 0
        self._check(self.lib.ti_inbox_take(self._h, C.c_void_p(inbox), int(inbox_capacity), int(world), self._ptr(dst), cap, self._ptr(status)))

    OPT_PUSH_BLOCKS = 9
    OPT_L2_SCRATCH_KB = 10
    OPT_PUSH_TMA = 11
    OPT_RECTIFY_QUAD = 12
    OPT_SMEM_HEADROOM_KB = 13
    SMEM_HEADROOM_DEFAULT_KB = 20
    SMEM_HEADROOM_NCCL_KB = 36

    def nccl_barrier(self) -> None:
        self._check(self.lib.ti_nccl_barrier(self._h))

    def peer_alloc(self, nbytes: int) -> tuple[int, bytes]:
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        self._check(self.lib.ti_peer_alloc(self._h, nbytes, C.byref(ptr), handle))
        return int(ptr.value), handle.raw

    def peer_open(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        self._check(self.lib.ti_peer_open(self._h, C.create_string_buffer(handle, 64), C.byref(ptr)))
        return int(ptr.value)

    def peer_close(self, ptr: int) -> None:
        self._check(self.lib.ti_peer_close(self._h, C.c_void_p(ptr)))

    def peer_free(self, ptr: int) -> None:
        self._check(self.lib.ti_peer_free(self._h, C.c_void_p(ptr)))
