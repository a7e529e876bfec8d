"""Rig description -> rig poses (API mirror of the calibration half of ``thor_slam/camera/utils.py``).

``parse_urdf_transform`` (reference :101-126) and ``load_rig_extrinsics_from_urdf`` (:129-178) with the
same names, arguments, warnings and error behaviour.  Device discovery / interactive prompts of the
reference module need DepthAI hardware and are out of scope.

Euler order: the reference builds the rotation with scipy ``Rotation.from_euler("XYZ", rpy)`` - that is
*intrinsic* X-Y-Z, ``R = Rx(r) @ Ry(p) @ Rz(y)`` - although its comment (and the URDF standard) say
fixed-axis / extrinsic, ``R = Rz(y) @ Ry(p) @ Rx(r)``.  The two differ for every camera of
``examples/assets/brackets.urdf``.  ``euler="reference"`` (default) reproduces the reference bit for bit
(golden vectors in ``tests/golden/urdf.json``); ``euler="urdf"`` gives the URDF-standard matrix.
"""

from __future__ import annotations

import logging
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np

from thor_slam_b200.camera.calibration import Extrinsics

logger = logging.getLogger(__name__)


def _rx(a: float) -> np.ndarray:
    c, s = np.cos(a), np.sin(a)
    return np.array([[1.0, 0, 0], [0, c, -s], [0, s, c]])


def _ry(a: float) -> np.ndarray:
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1.0, 0], [-s, 0, c]])


def _rz(a: float) -> np.ndarray:
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])


def rpy_to_matrix(rpy: tuple[float, float, float] | list[float], euler: str = "reference") -> np.ndarray:
    r, p, y = (float(v) for v in rpy)
    if euler == "reference":  # scipy intrinsic "XYZ"
        return _rx(r) @ _ry(p) @ _rz(y)
    if euler == "urdf":  # fixed-axis roll, pitch, yaw
        return _rz(y) @ _ry(p) @ _rx(r)
    raise ValueError(f"euler must be 'reference' or 'urdf', got {euler!r}")


def parse_urdf_transform(joint_elem: ET.Element, euler: str = "reference") -> np.ndarray:
    """4x4 ``parent_T_child`` of a URDF joint ``<origin xyz=... rpy=...>`` (identity, with a warning, if absent)."""
    origin = joint_elem.find("origin")
    if origin is None:
        logger.warning("Joint %s has no origin tag, assuming identity.", joint_elem.get("name"))
        return np.eye(4)
    xyz = np.array([float(v) for v in origin.get("xyz", "0 0 0").split()])
    rpy = [float(v) for v in origin.get("rpy", "0 0 0").split()]
    m = np.eye(4)
    m[:3, :3] = rpy_to_matrix(rpy, euler)
    m[:3, 3] = xyz
    return m


def load_rig_extrinsics_from_urdf(urdf_path: str | Path, camera_map: dict[str, str], euler: str = "reference") -> dict[str, Extrinsics]:
    """``{source name: base_link_T_source}`` from a star-topology URDF (joints whose parent is ``base_link``).

    ``camera_map``: source name -> URDF link name.  Missing links only log a warning, a missing file
    raises ``FileNotFoundError`` - both as in the reference.
    """
    urdf_path = Path(urdf_path)
    if not urdf_path.exists():
        raise FileNotFoundError(f"URDF not found at {urdf_path}")
    root = ET.parse(urdf_path).getroot()
    out: dict[str, Extrinsics] = {}
    for source, link in camera_map.items():
        for joint in root.findall("joint"):
            child = joint.find("child")
            if child is None or child.get("link") != link:
                continue
            parent = joint.find("parent")
            if parent is None or parent.get("link") != "base_link":
                logger.warning("Skipping joint %s: parent is not base_link", joint.get("name"))
                continue
            out[source] = Extrinsics.from_4x4_matrix(parse_urdf_transform(joint, euler))
            logger.info("Loaded extrinsics for %s (found link: %s)", source, link)
            break
        else:
            logger.warning("Could not find URDF link matching '%s' for source %s", link, source)
    return out
