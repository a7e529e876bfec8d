"""ctypes binding of ``libthoringest.so`` (the C ABI in ``include/thoringest.h``).

There is exactly one product library: ``thor_slam_b200/libthoringest.so``, built in-tree for
sm_100a by ``thor_slam_b200/csrc/Makefile``.  If it is missing, importing the ingest stage fails
loudly - there is no CPU fallback.  (Tests may inject another ``ctypes.CDLL`` that exports the same
ABI - the CPU emulation under ``tests/emu`` - through ``IngestLibrary(cdll=...)``; nothing in this
package ever does.)
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent.parent / "libthoringest.so"
ABI_VERSION = 2
INBOX_HEADER_BYTES = 128  # TI_INBOX_HEADER_BYTES

TI_OK, TI_EINVAL, TI_ECUDA, TI_ENCCL, TI_ESTATE, TI_ENOMEM = range(6)


class TiStream(C.Structure):
    """``struct ti_stream`` - field order and types must match the header exactly."""

    _fields_ = [
        ("kind", C.c_int32),
        ("camera", C.c_int32),
        ("src_format", C.c_int32),
        ("dst_format", C.c_int32),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("src", C.c_void_p),
        ("dst", C.c_void_p),
        ("src_frame_stride", C.c_uint64),
        ("dst_frame_stride", C.c_uint64),
        ("mask", C.c_void_p),
        ("mask_frame_stride", C.c_uint64),
        ("count", C.c_void_p),
    ]


class TiDepthStream(C.Structure):
    """``struct ti_depth_stream``."""

    _fields_ = [("camera", C.c_int32), ("reserved", C.c_int32), ("depth", C.c_void_p), ("depth_frame_stride", C.c_uint64)]


# name -> (restype, argtypes); every symbol declared in include/thoringest.h
SIGNATURES: dict[str, tuple] = {
    "ti_abi_version": (C.c_int, []),
    "ti_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "ti_destroy": (C.c_int, [C.c_void_p]),
    "ti_last_error": (C.c_char_p, [C.c_void_p]),
    "ti_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ti_sync": (C.c_int, [C.c_void_p]),
    "ti_launch_count": (C.c_uint64, [C.c_void_p]),
    "ti_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "ti_device_sm_count": (C.c_int, [C.c_void_p]),
    "ti_upload_rectify_map": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ti_upload_projection": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ti_rectify_plan": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32)]),
    "ti_upload_registration": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int,
                                         C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ti_register_colour": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64]),
    "ti_get_valid_mask": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ti_convert": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64]),
    "ti_rectify": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64]),
    "ti_backproject": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64]),
    "ti_depth_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p]),
    "ti_backproject_colour": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]),
    "ti_set_voxel_grid": (C.c_int, [C.c_void_p, C.c_double, C.c_uint32]),
    "ti_voxel_cloud": (C.c_int, [C.c_void_p, C.POINTER(TiDepthStream), C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64,
                                 C.c_void_p, C.c_void_p]),
    "ti_voxel_points": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "ti_ingest": (C.c_int, [C.c_void_p, C.POINTER(TiStream), C.c_int, C.c_int]),
    "ti_ingest_host": (C.c_int, [C.c_void_p, C.POINTER(TiStream), C.c_int, C.c_int, C.c_int]),
    "ti_ingest_host_submit": (C.c_int, [C.c_void_p, C.POINTER(TiStream), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]),
    "ti_ingest_host_wait": (C.c_int, [C.c_void_p, C.c_uint64]),
    "ti_copy_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "ti_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "ti_nccl_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "ti_gather_clouds": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_int]),
    "ti_gather_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "ti_gather_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]),
    "ti_gather_counts_begin": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ti_gather_counts_finish": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32)]),
    "ti_gather_records": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.c_int]),
    "ti_exchange_fence": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "ti_exchange_wait": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int]),
    "ti_nccl_barrier": (C.c_int, [C.c_void_p]),
    "ti_inbox_init": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ti_cloud_push": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32]),
    "ti_inbox_take": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]),
    "ti_peer_alloc": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.c_void_p]),
    "ti_peer_open": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "ti_peer_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ti_peer_free": (C.c_int, [C.c_void_p, C.c_void_p]),
}


class IngestLibraryError(ImportError):
    pass


class IngestLibrary:
    """Typed handle on the shared library."""

    def __init__(self, cdll: C.CDLL | None = None) -> None:
        if cdll is None:
            if not LIB_PATH.exists():
                raise IngestLibraryError(
                    f"{LIB_PATH} is missing. Build it with `make -C thor_slam_b200/csrc` "
                    "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
                    "The ingest stage runs on the GPU only; there is no CPU fallback."
                )
            cdll = C.CDLL(str(LIB_PATH))
        self.cdll = cdll
        for name, (restype, argtypes) in SIGNATURES.items():
            try:
                fn = getattr(cdll, name)
            except AttributeError as exc:
                raise IngestLibraryError(f"{getattr(cdll, '_name', cdll)} does not export {name}") from exc
            fn.restype = restype
            fn.argtypes = argtypes
            setattr(self, name, fn)
        got = self.ti_abi_version()
        if got != ABI_VERSION:
            raise IngestLibraryError(f"ABI version mismatch: library {got}, binding {ABI_VERSION}")
        self.is_emulation = hasattr(cdll, "ti_emu_marker")


_default: IngestLibrary | None = None


def default_library() -> IngestLibrary:
    global _default
    if _default is None:
        _default = IngestLibrary()
    return _default
