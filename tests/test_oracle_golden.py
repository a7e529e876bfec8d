"""The oracle against every golden vector generated from the reference (tests/golden/make_golden.py).

These are the pins of the oracle's *conventions*: 4x4 direction, pose composition, stream order,
BGR->RGB / mono8 pass-through, CameraInfo rule, baseline term, RDF_TO_FLU, URDF joint -> 4x4.
"""

from __future__ import annotations

import json

import numpy as np
import pytest

from oracle import backproject as ob
from oracle import conventions as conv
from oracle import convert as oc
from oracle import rectify as orc
from tests.conftest import GOLDEN


def test_readme_known_answer():
    """README.md:187-201 - the only known-answer vector in the reference."""
    g = np.load(GOLDEN / "isaac_adapter.npz")
    assert np.array_equal(conv.RDF_TO_FLU, g["RDF_TO_FLU_MATRIX"])
    assert np.array_equal(conv.RDF_TO_FLU @ np.array([1, 0, 0, 1]), [0, -1, 0, 1])
    assert np.array_equal(g["readme_known_answer"], [0, -1, 0, 1])
    assert np.isclose(np.linalg.det(conv.RDF_TO_FLU[:3, :3]), 1.0)


def test_world_extrinsics_composition():
    g = np.load(GOLDEN / "calibration.npz")
    for name, n in (("a", 2), ("b", 1)):
        for i in range(n):
            got = conv.world_T_camera(g[f"rig_{name}"], g[f"cam_{name}_{i}"])
            assert np.array_equal(got, g[f"world_{name}_{i}"])  # same float64 matmul -> bit-exact
    for i in range(2):  # source "c" has no rig pose: extrinsics come back unchanged
        assert np.array_equal(conv.world_T_camera(None, g[f"cam_c_{i}"]), g[f"world_c_{i}"])
    assert bool(g["unknown_is_none"])


def test_adapter_stream_order_images_and_camera_info():
    meta = json.loads((GOLDEN / "isaac_adapter.json").read_text())
    g = np.load(GOLDEN / "isaac_adapter.npz")
    # stream order: sorted source names x cam_idx
    order = conv.stream_order({n: [0, 1] for n in meta["sources"]}, 4)
    assert order == [(c["source_name"], c["cam_idx"]) for c in meta["cameras"]]
    # world extrinsics of each global camera
    from thor_slam_b200.camera.synthetic import SyntheticCameraConfig, SyntheticCameraSource

    a = SyntheticCameraSource(SyntheticCameraConfig(name="192.168.2.25", resolution=(96, 64), pixel_format="mono8", seed=5, pool=1))
    b = SyntheticCameraSource(SyntheticCameraConfig(name="192.168.2.21", resolution=(96, 64), pixel_format="bgr8", seed=6, pool=1, distortion="plumb_bob5"))
    by_name = {s.name: s for s in (a, b)}
    for s in (a, b):
        s.start()
    frames = {s.name: s.get_latest_frames() for s in (a, b)}
    for i, (name, idx) in enumerate(order):
        src = by_name[name]
        world = conv.world_T_camera(g[f"rig_ext_{name}"], src.get_extrinsics()[idx].to_4x4_matrix())
        assert np.allclose(world, g[f"cam_ext_{i}"], atol=0, rtol=0)
        img = frames[name][idx].image
        want_enc = meta["images"][i]["encoding"]
        if img.ndim == 2:
            assert want_enc == "mono8" and np.array_equal(img, g[f"image_{i}"])
        else:
            assert want_enc == "rgb8" and np.array_equal(oc.bgr_to_rgb_cv(img), g[f"image_{i}"])
            assert np.array_equal(oc.bgr_to_rgb_np(img), g[f"image_{i}"])
        intr = src.get_intrinsics()[idx]
        info = conv.camera_info(intr.matrix, intr.coeffs, intr.width, intr.height)
        ref = meta["infos"][i]
        for key in ("width", "height", "distortion_model", "d", "k", "r"):
            assert info[key] == ref[key], key
        p = np.array(info["p"]).reshape(3, 4)
        if idx == 1:
            left_world = conv.world_T_camera(g[f"rig_ext_{name}"], src.get_extrinsics()[0].to_4x4_matrix())
            _, tx = conv.right_camera_tx(left_world[:3, :3], left_world[:3, 3], world[:3, 3], intr.matrix[0, 0])
            p[0, 3] = tx
        assert np.allclose(p.flatten(), ref["p"], rtol=1e-13, atol=1e-15)
    # optical-frame TF: the reference publishes flu_to_rdf (quirk (i) of SURVEY 8c), quaternion [0.5,-0.5,0.5,0.5]
    optical = [t for t in meta["tf"] if t["child"].endswith("_optical_frame")]
    assert all(np.allclose(t["q"], [0.5, -0.5, 0.5, 0.5]) and t["t"] == [0.0, 0.0, 0.0] for t in optical)


def test_rgbd_publisher_rules():
    m = json.loads((GOLDEN / "rgbd_publisher.json").read_text())
    assert m["rgb_encoding"] == "rgb8" and m["depth_encoding"] == "16UC1"
    assert m["rgb_is_bgr_reversed"] and m["depth_unchanged"] and m["depth_dtype"] == "uint16"
    assert m["frame_id"] == "camera_2_optical_frame"
    for key, kk, dd in (("rgb_info", "rgb_K", "rgb_coeffs"), ("depth_info", "depth_K", "depth_coeffs")):
        info = conv.camera_info(np.array(m[kk]), np.array(m[dd]), m[key]["width"], m[key]["height"])
        for f in ("distortion_model", "d", "k", "r", "p"):
            assert info[f] == m[key][f], (key, f)


def test_urdf_rig_extrinsics(tmp_path):
    u = json.loads((GOLDEN / "urdf.json").read_text())
    for src, j in u["joints"].items():
        assert np.allclose(conv.urdf_origin_to_matrix(j["xyz"], j["rpy"]), np.array(u["matrices"][src]), rtol=0, atol=1e-15)
    a = u["author_case"]  # utils.py:99-100: "1 m in x, 0.5 m in y, 0.25 m in z ... roll pitch yaw"
    assert np.allclose(conv.urdf_origin_to_matrix(a["xyz"], a["rpy"]), np.array(a["matrix"]), rtol=0, atol=1e-15)
    # whole-file path: write a URDF with the same joints and go through the file parser
    joints = "".join(
        f'<joint name="{j["joint"]}" type="fixed"><parent link="base_link"/><child link="{j["link"]}"/>'
        f'<origin xyz="{j["xyz"]}" rpy="{j["rpy"]}"/></joint>' for j in u["joints"].values())
    path = tmp_path / "rig.urdf"
    path.write_text(f'<robot name="r"><link name="base_link"/>{joints}</robot>')
    got = conv.urdf_rig_extrinsics(str(path), u["camera_map"])
    assert set(got) == set(u["matrices"])
    for src, m in got.items():
        assert np.allclose(m, np.array(u["matrices"][src]), rtol=0, atol=1e-15)
    assert u["no_origin_is_identity"]


def test_cv_arithmetic_fixtures():
    """Frozen OpenCV outputs vs the numpy restatements (and vs OpenCV on this machine)."""
    g = np.load(GOLDEN / "cv_arith.npz")
    assert np.array_equal(oc.bgr_to_rgb_np(g["bgr"]), g["bgr2rgb"])
    assert np.array_equal(oc.bgr_to_gray_np(g["bgr"]), g["bgr2gray"])
    assert np.array_equal(oc.bgr_to_gray_cv(g["bgr"]), g["bgr2gray"])
    for tag in ("nv12", "nv12lim"):
        assert np.array_equal(oc.nv12_to_gray_np(g[tag]), g[f"{tag}2gray"])
        assert np.array_equal(oc.nv12_to_rgb_np(g[tag]), g[f"{tag}2rgb"])
        assert np.array_equal(oc.nv12_to_bgr_np(g[tag]), g[f"{tag}2bgr"])
        assert np.array_equal(oc.nv12_to_rgb_cv(g[tag]), g[f"{tag}2rgb"])
    for side in "lr":
        mx, my = orc.undistort_rectify_map_np(g[f"K_{side}"], g[f"D_{side}"], g[f"R_{side}"], g[f"P_{side}"], (160, 100))
        qa, qb = orc.quantize_map(mx, my), orc.quantize_map(g[f"mapx_{side}"], g[f"mapy_{side}"])
        assert np.array_equal(qa[0], qb[0]) and np.array_equal(qa[1], qb[1])
        assert np.array_equal(orc.remap_u8_np(g[f"img_{side}"], g[f"mapx_{side}"], g[f"mapy_{side}"]), g[f"rect_{side}"])
        assert np.array_equal(orc.remap_cv(g[f"img_{side}"], g[f"mapx_{side}"], g[f"mapy_{side}"]), g[f"rect_{side}"])
    r1, r2, p1, p2 = orc.stereo_rectify_cv(g["K_l"], g["D_l"], g["K_r"], g["D_r"], (160, 100), g["T_left_to_ref"], g["T_right_to_ref"])
    assert np.allclose(r1, g["R_l"], atol=1e-12) and np.allclose(p2, g["P_r"], atol=1e-9)
    assert np.array_equal(orc.remap_u8_np(g["bgr2gray"], g["edge_mapx"], g["edge_mapy"]), g["edge_rect_u8"])
    assert np.array_equal(orc.remap_u8_np(g["bgr"], g["edge_mapx"], g["edge_mapy"]), g["edge_rect_c3"])
    f32 = orc.remap_f32_np(g["bgr2gray"].astype(np.float32), g["edge_mapx"], g["edge_mapy"])
    assert np.allclose(f32, g["edge_rect_f32"], rtol=1e-5, atol=1e-4)  # north_star fp32 tolerance


def test_backprojection_oracle_known_points():
    k = np.array([[800.0, 0, 640], [0, 800.0, 400], [0, 0, 1]])
    depth = np.zeros((800, 1280), np.uint16)
    depth[400, 640] = 2000  # principal point, 2 m
    depth[400, 1040] = 1000  # 400 px right of it at 1 m -> x = 0.5 m
    depth[0, 0] = 0
    pts, mask, cnt = ob.backproject(depth, k, conv.RDF_TO_FLU)
    assert cnt == 2 and mask.sum() == 2
    assert np.allclose(pts[400, 640], [2.0, 0.0, 0.0])  # straight ahead -> +x in FLU
    assert np.allclose(pts[400, 1040], [1.0, -0.5, 0.0])  # right of centre -> -y in FLU
    assert ob.depth_stats(depth) == {"count": 2, "mean": 1500.0, "min": 1000, "max": 2000}
    assert ob.depth_stats(np.zeros((4, 4), np.uint16))["count"] == 0


@pytest.mark.parametrize("n,model", [(14, "rational_polynomial"), (8, "rational_polynomial"), (5, "plumb_bob"), (4, "equidistant"), (2, "plumb_bob"), (0, "plumb_bob")])
def test_distortion_model_rule(n, model):
    m, d = orc.select_distortion(np.arange(1, n + 1, dtype=float))
    assert m == model
    assert len(d) == {"rational_polynomial": 8, "plumb_bob": 5, "equidistant": 4}[model]


def test_voxel_oracle_fma_route_agrees_with_float64_numpy():
    """The C restatement (fma, constants folded) against floor(backproject(...) / voxel) in plain float64 numpy: identical keys
    wherever p / voxel is not within 1e-9 of an integer, and never more than one voxel apart."""
    from oracle import voxel as ov
    from tests import cases
    from thor_slam_b200.camera.synthetic import SyntheticCameraConfig, SyntheticCameraSource, make_depth, make_depth_scene

    rng = np.random.default_rng(7)
    for w, h, scene in [(320, 200, "room"), (320, 200, "noise"), (101, 57, "noise")]:
        s = SyntheticCameraSource(SyntheticCameraConfig(name="oak0", resolution=(w, h), pool=1, seed=7))
        k = s.get_intrinsics()[0].matrix
        m = cases.conv.body_T_camera(cases.random_pose(rng), s.get_extrinsics()[0].to_4x4_matrix(), "rdf")
        d = make_depth_scene(rng, w, h) if scene == "room" else make_depth(rng, w, h)
        keys, valid = ov.voxel_keys(d, k, m, 0.05, 10000)
        assert np.array_equal(valid, (d > 0) & (d <= 10000))
        ref, margin = ov.voxel_keys_f64(d, k, m, 0.05)
        clear = valid & (margin > 1e-9)
        assert clear.sum() > 0.99 * valid.sum()
        assert np.array_equal(keys[clear], ref[clear])
        assert np.abs(keys[valid] - ref[valid]).max() <= 1
    # identity pose, principal-point pixel: x = y = 0 exactly, z = d mm -> key (0, 0, floor(d / 50))
    k = np.array([[100.0, 0, 8], [0, 100.0, 4], [0, 0, 1]])
    d = np.full((8, 16), 1234, np.uint16)
    keys, _ = ov.voxel_keys(d, k, np.eye(4), 0.05, 0)
    assert tuple(keys[4, 8]) == (0, 0, 24)
    rec = ov.pack_records(keys[4:5, 8], set_id=3, tag=2)
    assert ov.unpack_records(rec)[0][0] == 2 and ov.unpack_records(rec)[1][0] == 3 and tuple(ov.unpack_records(rec)[2][0]) == (0, 0, 24)
    assert np.allclose(ov.record_points(rec, 0.05), [[0.025, 0.025, 1.225]])
