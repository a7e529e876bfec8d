"""Wire formats and stream kinds (mirror of the enums in ``include/thoringest.h``)."""

from __future__ import annotations

MONO8, BGR8, RGB8, NV12, DEPTH16, XYZ32F = range(6)
KIND_CONVERT, KIND_RECTIFY, KIND_BACKPROJECT = range(3)

FORMAT_BY_NAME = {"mono8": MONO8, "bgr8": BGR8, "rgb8": RGB8, "nv12": NV12, "depth16": DEPTH16, "xyz32f": XYZ32F}
NAME_BY_FORMAT = {v: k for k, v in FORMAT_BY_NAME.items()}


def fmt(value: int | str) -> int:
    if isinstance(value, str):
        try:
            return FORMAT_BY_NAME[value]
        except KeyError:
            raise ValueError(f"unknown format {value!r}; expected one of {sorted(FORMAT_BY_NAME)}") from None
    return int(value)


def frame_shape(f: int, width: int, height: int) -> tuple[int, ...]:
    """Array shape of one frame of format ``f`` with ``width x height`` pixels."""
    if f == MONO8 or f == DEPTH16:
        return (height, width)
    if f in (BGR8, RGB8):
        return (height, width, 3)
    if f == NV12:
        return (height * 3 // 2, width)
    if f == XYZ32F:
        return (height, width, 3)
    raise ValueError(f"unknown format {f}")


def frame_bytes(f: int, width: int, height: int) -> int:
    per_px = {MONO8: 1, BGR8: 3, RGB8: 3, DEPTH16: 2, XYZ32F: 12}
    if f == NV12:
        return width * height * 3 // 2
    return width * height * per_px[f]


def infer_format(image) -> int:
    """The reference tags nothing (isaac_ros.py:351-358 decides by ndim); extend that rule by dtype."""
    shape = tuple(image.shape)
    dtype = str(image.dtype).replace("torch.", "")
    if dtype == "uint16":
        return DEPTH16
    if len(shape) == 2:
        return MONO8
    if len(shape) == 3 and shape[2] == 3:
        return BGR8  # what getCvFrame() yields for colour streams
    raise ValueError(f"cannot infer a wire format from shape {shape} dtype {dtype}")
