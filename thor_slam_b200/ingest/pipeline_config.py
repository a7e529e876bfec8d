"""Rig description -> ingest rig (calibration ingest).

Mirror of the reference's YAML schema - ``scripts/run_pipeline.py:67-163`` (``CameraConfig`` /
``PipelineConfig.from_dict``) and ``config/slam_config.yaml`` - field for field and default for default
(pinned on ``tests/golden/pipeline_config.json``, produced by running the reference's own ``from_dict``),
plus the step the reference spreads over ``scripts/run_pipeline.py:488-610``: turn the description into
sources, rig poses from the URDF (``camera/utils.py:101-178``) and a rig.  Here the rig is an ``IngestRig``,
so the calibration goes straight into remap LUTs and projection constants on the GPU.

No OAK hardware exists in this environment: ``build_sources`` makes ``SyntheticCameraSource`` objects that
follow the same naming / resolution / format rules as ``LuxonisCameraSource`` would for each entry; a real
driver plugs in through ``source_factory``.
"""

from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Any, Callable

from thor_slam_b200.camera.synthetic import SyntheticCameraConfig, SyntheticCameraSource
from thor_slam_b200.camera.types import CameraSource, Extrinsics
from thor_slam_b200.camera.utils import load_rig_extrinsics_from_urdf

# scripts/run_pipeline.py:58-64 (same table in scripts/run_slam.py:45-50): camera IP -> URDF link of its bracket
CAMERA_MAP: dict[str, str] = {
    "192.168.2.25": "link_Camera_1_centroid",  # front low
    "192.168.2.21": "link_Camera_2_centroid",  # right
    "192.168.2.23": "link_Camera_3_centroid",  # up
    "192.168.2.22": "link_Camera_4_centroid",  # left
}


def _pair(v: Any) -> tuple[int, int] | None:
    return None if v is None else (v[0], v[1])


@dataclass
class CameraEntry:
    """One camera of the description (the reference's ``CameraConfig``, ``run_pipeline.py:67-82``)."""

    ip: str
    stereo: bool
    resolution: tuple[int, int]  # (width, height) of the stereo / SLAM streams
    sensor_type: str  # "COLOR" or "MONO"
    output_resolution: tuple[int, int] | None = None
    enable_rgbd: bool = False
    rgb_sensor_resolution: tuple[int, int] | None = None
    rgb_output_resolution: tuple[int, int] | None = None


@dataclass
class PipelineConfig:
    """The reference's ``PipelineConfig`` (``run_pipeline.py:85-163``)."""

    cameras: list[CameraEntry]
    fps: int = 30
    display: bool = False
    urdf_path: str = ""
    imu_report_rate: int = 400
    queue_size: int = 8
    rig_queue_size: int = 30
    rgbd_camera_ip: str | None = None  # deprecated in the reference: use nvblox_cameras
    nvblox_cameras: list[str] | None = None

    @classmethod
    def from_dict(cls, data: dict[str, Any], default_urdf: str | Path | None = None) -> "PipelineConfig":
        """``default_urdf`` stands in for the reference's ``examples/assets/brackets.urdf`` default (used when
        ``urdf_path`` is empty and the file exists)."""
        cameras = [
            CameraEntry(
                ip=c["ip"],
                stereo=c.get("stereo", True),
                resolution=_pair(c.get("resolution", [1280, 800])),
                sensor_type=c.get("sensor_type", "COLOR").upper(),
                output_resolution=_pair(c.get("output_resolution")),
                enable_rgbd=c.get("enable_rgbd", False),
                rgb_sensor_resolution=_pair(c.get("rgb_sensor_resolution")),
                rgb_output_resolution=_pair(c.get("rgb_output_resolution")),
            )
            for c in data.get("cameras", [])
        ]
        urdf_path = data.get("urdf_path", "")
        if not urdf_path and default_urdf is not None and Path(default_urdf).exists():
            urdf_path = str(default_urdf)
        nvblox = data.get("nvblox_cameras")
        if nvblox is None:
            legacy = data.get("rgbd_camera_ip")
            nvblox = [legacy] if legacy else [c["ip"] for c in data.get("cameras", []) if c.get("enable_rgbd", False)]
        return cls(
            cameras=cameras,
            fps=data.get("fps", 30),
            display=data.get("display", False),
            urdf_path=urdf_path,
            imu_report_rate=data.get("imu_report_rate", 400),
            queue_size=data.get("queue_size", 8),
            rig_queue_size=data.get("rig_queue_size", 30),
            rgbd_camera_ip=data.get("rgbd_camera_ip"),
            nvblox_cameras=nvblox if isinstance(nvblox, list) else None,
        )

    @classmethod
    def from_yaml(cls, path: str | Path, default_urdf: str | Path | None = None) -> "PipelineConfig":
        import yaml

        return cls.from_dict(yaml.safe_load(Path(path).read_text()) or {}, default_urdf)

    def calculate_num_cameras(self) -> int:
        """Streams cuVSLAM would see: 2 per stereo camera, 1 otherwise (``run_pipeline.py:161-163``)."""
        return sum(2 if c.stereo else 1 for c in self.cameras)


def build_sources(cfg: PipelineConfig, source_factory: Callable[[CameraEntry, PipelineConfig], CameraSource] | None = None,
                  seed: int = 1337) -> list[CameraSource]:
    """One source per camera entry, named by its IP like the Luxonis driver (``luxonis.py:759-819``):
    SLAM streams at ``output_resolution or resolution``, MONO -> mono8 / COLOR -> bgr8, RGB-D streams when the
    camera is in ``nvblox_cameras`` (or has ``enable_rgbd``), RGB at ``rgb_output_resolution or resolution``."""
    out: list[CameraSource] = []
    nvblox = set(cfg.nvblox_cameras or [])
    for i, c in enumerate(cfg.cameras):
        if source_factory is not None:
            out.append(source_factory(c, cfg))
            continue
        slam_res = c.output_resolution or c.resolution
        rgbd = c.enable_rgbd or c.ip in nvblox
        rgb_res = c.rgb_output_resolution or slam_res
        out.append(SyntheticCameraSource(SyntheticCameraConfig(
            name=c.ip, stereo=c.stereo, pixel_format="mono8" if c.sensor_type == "MONO" else "bgr8", resolution=slam_res,
            enable_rgbd=rgbd, rgb_resolution=rgb_res, depth_resolution=slam_res, fps=float(cfg.fps), seed=seed + 101 * i,
            time_offset=0.001 * i, read_imu=(i == 0), imu_rate_hz=float(cfg.imu_report_rate))))
    return out


def rig_extrinsics_from_config(cfg: PipelineConfig, camera_map: dict[str, str] | None = None) -> dict[str, Extrinsics]:
    """``base_link_T_source`` per camera IP from the description's URDF (empty dict without a URDF, as the reference
    then runs with identity rig poses)."""
    if not cfg.urdf_path:
        return {}
    wanted = {c.ip for c in cfg.cameras}
    cmap = {ip: link for ip, link in (camera_map or CAMERA_MAP).items() if ip in wanted}
    return load_rig_extrinsics_from_urdf(cfg.urdf_path, cmap)


def build_ingest_rig(cfg: PipelineConfig, *, camera_map: dict[str, str] | None = None,
                     source_factory: Callable[[CameraEntry, PipelineConfig], CameraSource] | None = None, **rig_kwargs: Any):
    """Description -> ``IngestRig``: sources, URDF rig poses, queue size; the rig's constructor uploads every
    camera's remap LUT and projection (``IngestRig._upload_calibration``)."""
    from thor_slam_b200.ingest.rig import IngestRig

    sources = build_sources(cfg, source_factory)
    imu_source = next((s.name for s in sources if getattr(s, "has_sensor_data", False)), None)
    return IngestRig(sources, queue_size=cfg.rig_queue_size, rig_extrinsics=rig_extrinsics_from_config(cfg, camera_map),
                     imu_source=imu_source, **rig_kwargs)
