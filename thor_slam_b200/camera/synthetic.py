"""Seeded synthetic ``CameraSource`` - stands in for the OAK driver.

The reference's only driver (``thor_slam/camera/drivers/luxonis.py``) needs
DepthAI hardware, so the benchmark and the parity tests use this plugin
instead.  It follows the driver's *output conventions*:

* stereo sources return ``[left, right]`` named ``"{name}_left"`` /
  ``"{name}_right"``, single sources return ``[rgb]`` named ``"{name}_rgb"``
  (luxonis.py:759-819); MONO streams are ``HxW`` u8, COLOR streams ``HxWx3`` u8
  in **BGR** order (what ``getCvFrame()`` yields);
* RGB-D extras (duck-typed, not part of the ABC - luxonis.py:871-1091):
  ``has_rgbd_streams``, ``get_latest_rgbd_frames`` -> ``(rgb BGR u8, depth u16 mm,
  0 = invalid)`` named ``"{name}_rgb"`` / ``"{name}_depth"``;
* calibration: ``Intrinsics`` at the *published* resolution with 14 OAK-style
  coefficients, stereo ``Extrinsics`` = left->CAM_A and right->CAM_A in metres
  (luxonis.py:675-709), RGB = identity, depth = CAM_B->CAM_A (:1068-1091);
* ``get_latest_frames`` before ``start`` raises ``RuntimeError`` (:765-766).

In addition a source may publish raw **NV12** buffers (``(H*3/2) x W`` u8) -
the camera's native format before ``getCvFrame()`` - and says so through
``get_stream_formats()`` so the ingest stage does the NV12 conversion itself.
Values follow SURVEY.md section 8(d): ``np.random.default_rng(seed)``, smooth
gradient + uniform noise, depth uniform 300..10000 mm with 20 % zeros and a few
65535.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from thor_slam_b200.camera.calibration import Extrinsics, Intrinsics
from thor_slam_b200.camera.frames import CameraFrame, CameraSource

PIXEL_FORMATS = ("mono8", "bgr8", "nv12")


@dataclass
class SyntheticCameraConfig:
    name: str
    stereo: bool = True
    pixel_format: str = "mono8"  # of the SLAM streams: mono8 | bgr8 | nv12
    resolution: tuple[int, int] = (1280, 800)  # (width, height) of the SLAM streams
    enable_rgbd: bool = False
    rgb_resolution: tuple[int, int] = (1920, 1080)
    depth_resolution: tuple[int, int] = (1280, 800)
    depth_align_to_rgb: bool = False
    fps: float = 30.0
    time_offset: float = 0.0  # seconds added to every timestamp of this source (<= 4 ms in the survey's rig)
    seed: int = 1337
    pool: int = 3  # distinct frames generated up front and cycled
    baseline_m: float = 0.075
    read_imu: bool = False
    imu_rate_hz: float = 400.0
    distortion: str = "rational14"  # rational14 | plumb_bob5 | fisheye4 | none


def make_intrinsics(rng: np.random.Generator, width: int, height: int, distortion: str = "rational14") -> Intrinsics:
    """OAK-like pinhole + distortion at (width, height); focal ~ 0.625 * width, +-1 %."""
    f = 0.625 * width
    fx = f * (1 + rng.uniform(-0.01, 0.01))
    fy = f * (1 + rng.uniform(-0.01, 0.01))
    cx = width / 2 + rng.uniform(-0.01, 0.01) * width
    cy = height / 2 + rng.uniform(-0.01, 0.01) * height
    k = np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]])
    if distortion == "rational14":
        # numerator and denominator nearly cancel, as in factory OAK calibrations
        k4, k5, k6 = rng.uniform(0.2, 0.5), rng.uniform(-0.1, 0.1), rng.uniform(-0.02, 0.02)
        k1, k2, k3 = k4 - rng.uniform(0.05, 0.12), k5 + rng.uniform(-0.03, 0.05), k6 + rng.uniform(-0.01, 0.01)
        p1, p2 = rng.uniform(-5e-4, 5e-4, size=2)
        s = rng.uniform(-2e-4, 2e-4, size=4)
        d = np.array([k1, k2, p1, p2, k3, k4, k5, k6, *s, 0.0, 0.0])
    elif distortion == "plumb_bob5":
        d = np.array([rng.uniform(-0.12, -0.05), rng.uniform(0.0, 0.05), *rng.uniform(-5e-4, 5e-4, size=2), rng.uniform(-0.01, 0.01)])
    elif distortion == "fisheye4":
        d = rng.uniform(-0.02, 0.02, size=4)
    elif distortion == "none":
        d = np.zeros(5)
    else:
        raise ValueError(f"unknown distortion model {distortion!r}")
    return Intrinsics(width=width, height=height, matrix=k, coeffs=d)


def _small_rotation(rng: np.random.Generator, max_rad: float) -> np.ndarray:
    w = rng.uniform(-max_rad, max_rad, size=3)
    th = float(np.linalg.norm(w))
    if th == 0.0:
        return np.eye(3)
    kx, ky, kz = w / th
    kmat = np.array([[0, -kz, ky], [kz, 0, -kx], [-ky, kx, 0]])
    return np.eye(3) + np.sin(th) * kmat + (1 - np.cos(th)) * (kmat @ kmat)


def _gradient_noise(rng: np.random.Generator, h: int, w: int, channels: int = 1) -> np.ndarray:
    yy, xx = np.mgrid[0:h, 0:w]
    out = np.empty((h, w, channels), dtype=np.uint8)
    for c in range(channels):
        phase = rng.uniform(0, 2 * np.pi)
        grad = 96 + 64 * np.sin(xx / w * 2 * np.pi * (1 + c) + phase) + 48 * (yy / h)
        noise = rng.integers(0, 64, size=(h, w))
        out[..., c] = np.clip(grad + noise, 0, 255).astype(np.uint8)
    return out[..., 0] if channels == 1 else out


def make_image(rng: np.random.Generator, fmt: str, width: int, height: int) -> np.ndarray:
    if fmt == "mono8":
        return _gradient_noise(rng, height, width)
    if fmt == "bgr8":
        return _gradient_noise(rng, height, width, 3)
    if fmt == "nv12":
        if width % 2 or height % 2:
            raise ValueError("NV12 needs even width and height")
        buf = np.empty((height * 3 // 2, width), dtype=np.uint8)
        buf[:height] = _gradient_noise(rng, height, width)
        buf[height:] = rng.integers(0, 256, size=(height // 2, width), dtype=np.uint8)  # full-range U,V
        return buf
    raise ValueError(f"unknown pixel format {fmt!r}")


def make_depth(rng: np.random.Generator, width: int, height: int) -> np.ndarray:
    d = rng.integers(300, 10001, size=(height, width)).astype(np.uint16)
    d[rng.random((height, width)) < 0.2] = 0
    hot = rng.integers(0, height * width, size=max(4, height * width // 50000))
    d.reshape(-1)[hot] = 65535
    return d


def make_depth_scene(rng: np.random.Generator, width: int, height: int, focal_px: float | None = None, hole_fraction: float = 0.2) -> np.ndarray:
    """A depth image of SURFACES (``make_depth`` is per-pixel noise, the worst case for any spatial down-sampling): a room
    seen from inside - floor, ceiling, two side walls, a back wall - with a few boxes in it, multiplicative stereo-like
    noise of 0.2 %, coherent holes covering about ``hole_fraction`` of the image and a few saturated pixels.
    Millimetres, 0 = invalid, like ``get_latest_rgbd_frames`` (``drivers/luxonis.py:876-921``)."""
    f = float(focal_px if focal_px is not None else 0.625 * width)
    x = (np.arange(width, dtype=np.float64)[None, :] - width / 2) / f
    y = (np.arange(height, dtype=np.float64)[:, None] - height / 2) / f
    half_w, up, down, back = rng.uniform(2.0, 3.5), rng.uniform(1.0, 1.8), rng.uniform(0.8, 1.4), rng.uniform(5.0, 9.0)
    with np.errstate(divide="ignore"):
        z = np.minimum(np.minimum(half_w / np.abs(x), np.where(y > 0, down / np.abs(y), up / np.abs(y))), back)
    z = np.broadcast_to(z, (height, width)).copy()
    for _ in range(int(rng.integers(3, 7))):  # boxes: fronto-parallel faces nearer than the room
        bw, bh = int(rng.integers(width // 16, width // 4)), int(rng.integers(height // 10, height // 3))
        u0, v0 = int(rng.integers(0, width - bw)), int(rng.integers(0, height - bh))
        zb = rng.uniform(0.6, 4.0)
        z[v0:v0 + bh, u0:u0 + bw] = np.minimum(z[v0:v0 + bh, u0:u0 + bw], zb)
    z *= 1.0 + 0.002 * rng.standard_normal((height, width))
    d = np.clip(np.rint(z * 1000.0), 1, 65534).astype(np.uint16)
    if hole_fraction > 0:
        coarse = rng.random(((height + 31) // 32, (width + 31) // 32)) < hole_fraction
        d[np.kron(coarse, np.ones((32, 32), bool))[:height, :width]] = 0
    hot = rng.integers(0, height * width, size=max(4, height * width // 50000))
    d.reshape(-1)[hot] = 65535
    return d


class SyntheticCameraSource(CameraSource):
    """Deterministic frames, calibration and clocks for one (stereo or single) camera."""

    def __init__(self, cfg: SyntheticCameraConfig) -> None:
        if cfg.pixel_format not in PIXEL_FORMATS:
            raise ValueError(f"pixel_format must be one of {PIXEL_FORMATS}, got {cfg.pixel_format!r}")
        if cfg.enable_rgbd and not cfg.stereo:
            raise ValueError("RGB-D needs a stereo source (depth comes from the stereo pair)")
        self.cfg = cfg
        self._running = False
        self._seq = 0
        self._rgbd_seq = 0
        self._imu_seq = 0
        rng = np.random.default_rng(cfg.seed)
        w, h = cfg.resolution
        n_streams = 2 if cfg.stereo else 1

        self._intrinsics = [make_intrinsics(rng, w, h, cfg.distortion) for _ in range(n_streams)]
        if cfg.stereo:
            half = cfg.baseline_m / 2
            left = np.eye(4)
            left[:3, :3] = _small_rotation(rng, 0.004)
            left[:3, 3] = [-half, rng.uniform(-2e-4, 2e-4), rng.uniform(-2e-4, 2e-4)]
            right = np.eye(4)
            right[:3, :3] = _small_rotation(rng, 0.004)
            right[:3, 3] = [half, rng.uniform(-2e-4, 2e-4), rng.uniform(-2e-4, 2e-4)]
            self._extrinsics = [Extrinsics.from_4x4_matrix(left), Extrinsics.from_4x4_matrix(right)]
        else:
            self._extrinsics = [Extrinsics.from_4x4_matrix(np.eye(4))]

        self._pool = [[make_image(rng, cfg.pixel_format, w, h) for _ in range(n_streams)] for _ in range(cfg.pool)]

        self._rgbd_pool: list[tuple[np.ndarray, np.ndarray]] = []
        self._rgbd_intrinsics: tuple[Intrinsics, Intrinsics] | None = None
        if cfg.enable_rgbd:
            rw, rh = cfg.rgb_resolution
            dw, dh = cfg.depth_resolution
            rgb_intr = make_intrinsics(rng, rw, rh, cfg.distortion)
            if cfg.depth_align_to_rgb:
                # depth shares the RGB camera model, rescaled if the sizes differ (luxonis.py:1018-1032)
                km = rgb_intr.matrix.copy()
                km[0, 0] *= dw / rw
                km[0, 2] *= dw / rw
                km[1, 1] *= dh / rh
                km[1, 2] *= dh / rh
                depth_intr = Intrinsics(dw, dh, km, rgb_intr.coeffs.copy())
            else:
                # unaligned depth lives in the left mono camera (CAM_B) (luxonis.py:1033-1051)
                li = self._intrinsics[0]
                km = li.matrix.copy()
                km[0, 0] *= dw / w
                km[0, 2] *= dw / w
                km[1, 1] *= dh / h
                km[1, 2] *= dh / h
                depth_intr = Intrinsics(dw, dh, km, li.coeffs.copy())
            self._rgbd_intrinsics = (rgb_intr, depth_intr)
            self._rgbd_pool = [(make_image(rng, "bgr8", rw, rh), make_depth(rng, dw, dh)) for _ in range(cfg.pool)]

        self._imu_rng = np.random.default_rng(cfg.seed + 7)

    # -- CameraSource ------------------------------------------------------
    @property
    def name(self) -> str:
        return self.cfg.name

    def start(self) -> None:
        self._running = True

    def stop(self) -> None:
        self._running = False

    def is_running(self) -> bool:
        return self._running

    def _stamp(self, seq: int) -> float:
        return seq / self.cfg.fps + self.cfg.time_offset

    def get_latest_frames(self) -> list[CameraFrame]:
        if not self._running:
            raise RuntimeError("Camera source not started. Call start() first.")
        seq = self._seq
        self._seq += 1
        images = self._pool[seq % len(self._pool)]
        ts = self._stamp(seq)
        if self.cfg.stereo:
            # the right sensor is read a hair later than the left one on real hardware
            return [
                CameraFrame(images[0], ts, seq, f"{self.name}_left"),
                CameraFrame(images[1], ts + 1e-4, seq, f"{self.name}_right"),
            ]
        return [CameraFrame(images[0], ts, seq, f"{self.name}_rgb")]

    def try_get_latest_frames(self) -> list[CameraFrame] | None:
        return self.get_latest_frames() if self._running else None

    def get_intrinsics(self) -> list[Intrinsics]:
        return self._intrinsics

    def get_extrinsics(self) -> list[Extrinsics]:
        return self._extrinsics

    def get_stream_formats(self) -> list[str]:
        """Extra (not in the reference ABC): wire format of each SLAM stream."""
        return [self.cfg.pixel_format] * (2 if self.cfg.stereo else 1)

    def get_sensor_extrinsics(self) -> Extrinsics | None:
        m = np.eye(4)
        m[:3, 3] = [0.0, -0.005, -0.01]
        return Extrinsics.from_4x4_matrix(m)

    @property
    def has_sensor_data(self) -> bool:
        return self.cfg.read_imu

    def get_timestamped_sensor_data(self) -> tuple[dict | None, float | None]:
        if not self.cfg.read_imu:
            return None, None
        seq = self._imu_seq
        self._imu_seq += 1
        ts = seq / self.cfg.imu_rate_hz + self.cfg.time_offset
        sample = {
            "accelerometer": self._imu_rng.normal([0.0, 9.81, 0.0], 0.02),
            "gyroscope": self._imu_rng.normal(0.0, 0.002, size=3),
            "timestamp": ts,
            "sequence_num": seq,
        }
        return sample, ts  # the latest packet, flat, like LuxonisCameraSource (luxonis.py:1098-1160)

    # -- RGB-D extras (duck-typed like LuxonisCameraSource) -----------------
    @property
    def has_rgbd_streams(self) -> bool:
        return self.cfg.stereo and self.cfg.enable_rgbd

    def get_latest_rgbd_frames(self) -> tuple[CameraFrame, CameraFrame]:
        if not self._running:
            raise RuntimeError("Camera source not started. Call start() first.")
        if not self.has_rgbd_streams:
            raise RuntimeError("RGB-D streams not enabled. Set enable_rgbd=True and stereo=True.")
        seq = self._rgbd_seq
        self._rgbd_seq += 1
        rgb, depth = self._rgbd_pool[seq % len(self._rgbd_pool)]
        ts = self._stamp(seq)
        return (
            CameraFrame(rgb, ts, seq, f"{self.name}_rgb"),
            CameraFrame(depth, ts, seq, f"{self.name}_depth"),
        )

    def try_get_latest_rgbd_frames(self) -> tuple[CameraFrame, CameraFrame] | None:
        if not self._running or not self.has_rgbd_streams:
            return None
        return self.get_latest_rgbd_frames()

    def get_rgbd_intrinsics(self) -> tuple[Intrinsics, Intrinsics]:
        if not self.has_rgbd_streams or self._rgbd_intrinsics is None:
            raise RuntimeError("RGB-D streams not enabled. Set enable_rgbd=True and stereo=True.")
        return self._rgbd_intrinsics

    def get_rgbd_extrinsics(self) -> tuple[Extrinsics, Extrinsics]:
        if not self.has_rgbd_streams:
            raise RuntimeError("RGB-D streams not enabled. Set enable_rgbd=True and stereo=True.")
        return Extrinsics.from_4x4_matrix(np.eye(4)), self._extrinsics[0]


def make_rig_sources(
    n_cameras: int = 4,
    resolution: tuple[int, int] = (1280, 800),
    pixel_format: str = "mono8",
    enable_rgbd: bool = False,
    seed: int = 1337,
    **kw: object,
) -> list[SyntheticCameraSource]:
    """The survey's ``n x OAK-D`` rig: names ``oak0..``, per-source clock skew <= 4 ms."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_cameras):
        cfg = SyntheticCameraConfig(
            name=f"oak{i}",
            resolution=resolution,
            pixel_format=pixel_format,
            enable_rgbd=enable_rgbd,
            time_offset=float(rng.uniform(0, 0.004)),
            seed=seed + 101 * i,
            **kw,  # type: ignore[arg-type]
        )
        out.append(SyntheticCameraSource(cfg))
    return out
