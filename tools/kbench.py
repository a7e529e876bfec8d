#!/usr/bin/env python
"""Kernel micro-benchmarks on one GPU (CUDA events, inputs >> L2): one line per kernel / variant.

    python tools/kbench.py [--batch 32] [--iters 20]

Prints achieved GB/s on the ALGORITHMIC bytes of each kernel and the
fraction of MEASURED_PEAKS.json hbm_gbs.  Development tool; bench.py is the judged benchmark.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from thor_slam_b200.camera.synthetic import SyntheticCameraConfig, SyntheticCameraSource, make_depth, make_depth_scene  # noqa: E402
from thor_slam_b200.ingest import formats as F  # noqa: E402
from thor_slam_b200.ingest.calib import body_T_camera, stereo_rectify_maps  # noqa: E402
from thor_slam_b200.ingest.context import IngestContext, StreamSpec  # noqa: E402


def timeit(fn, iters: int, warm: int = 3) -> float:
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters  # ms


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--one", action="store_true", help="rectify: first variant only (for ncu)")
    args = ap.parse_args()
    peak = 6454.3
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        peak = float(json.loads(p.read_text())["hbm_gbs"])
    torch.cuda.set_device(0)
    ctx = IngestContext(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    W, H, B = 1280, 800, args.batch
    rows = []

    def report(name: str, ms: float, algo_bytes: float, px: float) -> None:
        gbs = algo_bytes / (ms * 1e-3) / 1e9
        rows.append((name, ms, gbs, gbs / peak, px / (ms * 1e-3) / 1e9))
        print(f"{name:44s} {ms:9.4f} ms  {gbs:8.1f} GB/s  frac {gbs / peak:6.3f}  {px / (ms * 1e-3) / 1e9:8.2f} GPix/s", flush=True)

    rng = np.random.default_rng(0)
    src = SyntheticCameraSource(SyntheticCameraConfig(name="oak0", resolution=(W, H), pool=2, enable_rgbd=False))
    maps = stereo_rectify_maps(src.get_intrinsics(), src.get_extrinsics(), (W, H))
    NS = 8
    for cam in range(NS):
        ctx.upload_rectify_map(cam, *maps[cam % 2], (W, H))
    frames = [torch.from_numpy(np.stack([src._pool[b % 2][s % 2] for b in range(B)])).cuda() for s in range(NS)]
    outs = [torch.empty_like(f) for f in frames]
    specs = [StreamSpec(F.KIND_RECTIFY, frames[s], outs[s], F.MONO8, F.MONO8, camera=s) for s in range(NS)]
    px = NS * B * W * H

    if not args.only or "rect" in args.only:
        print("plan", ctx.rectify_plan(0), flush=True)
        for variant, th, fpu, stages, pf in ((4, 32, 0, 4, 0), (4, 32, 0, 3, 0), (4, 32, 0, 5, 0), (4, 32, 0, 6, 0), (4, 24, 0, 3, 0), (4, 24, 0, 4, 0),
                                             (4, 24, 0, 2, 0), (4, 16, 0, 4, 0), (3, 32, 16, 2, 0))[:1 if args.one else None]:
            ctx.set_option(ctx.OPT_MONO_VARIANT, variant)
            ctx.set_option(ctx.OPT_TMA_TILE_H, th)
            ctx.set_option(ctx.OPT_FRAMES_PER_UNIT, fpu)
            ctx.set_option(ctx.OPT_STAGES, stages)
            ctx.set_option(ctx.OPT_LUT_PREFETCH, pf)
            report(f"rectify mono v{variant} th={th} fpu={fpu} S={stages} prefetch={pf}", timeit(lambda: ctx.ingest(specs), args.iters), 2 * px, px)
        ctx.set_option(ctx.OPT_CTAS_PER_SM, 0)
        ctx.set_option(ctx.OPT_FRAMES_PER_UNIT, 16)
        ctx.set_option(ctx.OPT_FRAMES_PER_UNIT, 0)
        ctx.set_option(ctx.OPT_STAGES, 6)
        ctx.set_option(ctx.OPT_LUT_PREFETCH, 0)
        ctx.set_option(ctx.OPT_MONO_VARIANT, 4)
        ctx.set_option(ctx.OPT_TMA_TILE_H, 32)
        if args.one:
            prep = ctx.prepare(specs)  # what the live rig does: the stream array packed once, one foreign call per frame set
            report("  same, prepared stream array (ingest_prepared)", timeit(lambda: ctx.ingest_prepared(prep), args.iters), 2 * px, px)
            ctx.close()
            return
        for dbg, what in ((1, "loads only (no blend)"), (2, "blend only (no loads)"), (3, "pipeline only"), (8, "half the window loads (wrong pixels)")):
            ctx.set_option(ctx.OPT_DEBUG, dbg)
            report(f"rectify mono v4 DEBUG {what}", timeit(lambda: ctx.ingest(specs), args.iters), 2 * px, px)
        ctx.set_option(ctx.OPT_DEBUG, 0)
        for per_sm in (1, 2, 3):
            ctx.set_option(ctx.OPT_CTAS_PER_SM, per_sm)
            report(f"rectify mono v4 th=32 ctas/sm={per_sm}", timeit(lambda: ctx.ingest(specs), args.iters), 2 * px, px)
        ctx.set_option(ctx.OPT_CTAS_PER_SM, 0)
        ctx.set_option(ctx.OPT_FORCE_GENERIC_RECTIFY, 1)
        report("rectify mono generic (v1 tiled)", timeit(lambda: ctx.ingest(specs), max(3, args.iters // 4)), 2 * px, px)
        ctx.set_option(ctx.OPT_FORCE_GENERIC_RECTIFY, 0)
        # plain device copy of the same bytes = what "1.0" looks like for this traffic
        report("torch copy_ (same bytes)", timeit(lambda: [o.copy_(f) for o, f in zip(outs, frames)], args.iters), 2 * px, px)

    if not args.only or "quad" in args.only:
        # the pair-window kernel's two layouts on the same maps (the layout is chosen when the map is uploaded)
        for quad in (1, 0, 1):
            ctx.set_option(ctx.OPT_RECTIFY_QUAD, quad)
            for cam in range(NS):
                ctx.upload_rectify_map(cam, *maps[cam % 2], (W, H))
            print("plan", ctx.rectify_plan(0), ctx.rectify_plan(1), flush=True)
            for stages in (3, 4, 5, 6):
                ctx.set_option(ctx.OPT_STAGES, stages)
                report(f"rectify mono v4 quad={quad} S={stages}", timeit(lambda: ctx.ingest(specs), args.iters), 2 * px, px)
            ctx.set_option(ctx.OPT_STAGES, 6)
        ctx.set_option(ctx.OPT_RECTIFY_QUAD, 1)

    if not args.only or "downscale" in args.only:
        # the 2 x downscale of config/slam_config.yaml's output_resolution done on the host: 1280x800 -> 640x400 (8 mono streams) and
        # 1920x1200 -> 960x600 BGR -> RGB (4 colour streams); algorithmic bytes = source read + result written
        yy, xx = np.mgrid[0:400, 0:640].astype(np.float32)
        for cam in range(NS):
            ctx.upload_rectify_map(24 + cam, xx * 2.0 + 0.25 + 0.002 * yy, yy * 2.0 + 0.75 - 0.002 * xx, (W, H))
        print("plan", ctx.rectify_plan(24), flush=True)
        douts = [torch.empty((B, 400, 640), dtype=torch.uint8, device="cuda") for _ in range(NS)]
        dspecs = [StreamSpec(F.KIND_RECTIFY, frames[s], douts[s], F.MONO8, F.MONO8, camera=24 + s) for s in range(NS)]
        report("rectify mono 1280x800 -> 640x400 (wide-pitch pair windows)", timeit(lambda: ctx.ingest(dspecs), args.iters), 1.25 * px, px)
        ctx.set_option(ctx.OPT_MONO_VARIANT, 3)
        report("  the same on the round-1 route (v3 / v2)", timeit(lambda: ctx.ingest(dspecs), args.iters), 1.25 * px, px)
        ctx.set_option(ctx.OPT_MONO_VARIANT, 4)
        # every distortion model of the reference's CameraInfo rule (isaac_ros.py:370-383) at the headline shape
        for model in ("fisheye4", "plumb_bob5", "none"):
            fsrc = SyntheticCameraSource(SyntheticCameraConfig(name="f0", resolution=(W, H), pool=1, enable_rgbd=False, distortion=model, seed=11))
            fmaps = stereo_rectify_maps(fsrc.get_intrinsics(), fsrc.get_extrinsics(), (W, H))
            for cam in range(NS):
                ctx.upload_rectify_map(40 + cam, *fmaps[cam % 2], (W, H))
            fspecs = [StreamSpec(F.KIND_RECTIFY, frames[s], outs[s], F.MONO8, F.MONO8, camera=40 + s) for s in range(NS)]
            pl = ctx.rectify_plan(40)
            report(f"rectify mono 1280x800 {model}: variant {pl['variant']}, {pl['exceptions_per_warp']} exc/warp, {pl['overflow_pixels']} overflow px",
                   timeit(lambda: ctx.ingest(fspecs), args.iters), 2 * px, px)
        CW, CH, NB, NC = 1920, 1200, max(2, B // 2), 4
        yy, xx = np.mgrid[0:600, 0:960].astype(np.float32)
        for cam in range(NC):
            ctx.upload_rectify_map(32 + cam, xx * 2.0 + 0.6 + 0.002 * yy, yy * 2.0 + 0.3 - 0.002 * xx, (CW, CH))
        print("plan", ctx.rectify_plan(32), flush=True)
        cin = [torch.randint(0, 256, (NB, CH, CW, 3), dtype=torch.uint8, device="cuda") for _ in range(NC)]
        cout = [torch.empty((NB, 600, 960, 3), dtype=torch.uint8, device="cuda") for _ in range(NC)]
        cspecs = [StreamSpec(F.KIND_RECTIFY, cin[i], cout[i], F.BGR8, F.RGB8, camera=32 + i) for i in range(NC)]
        cpx = NC * NB * CW * CH
        report("rectify bgr8->rgb8 1920x1200 -> 960x600 (wide-pitch 3-channel windows)", timeit(lambda: ctx.ingest(cspecs), args.iters), 3.75 * cpx, cpx)
        ctx.set_option(ctx.OPT_FORCE_GENERIC_RECTIFY, 1)
        report("  the same on the round-1 route (generic tiled)", timeit(lambda: ctx.ingest(cspecs), max(3, args.iters // 4)), 3.75 * cpx, cpx)
        ctx.set_option(ctx.OPT_FORCE_GENERIC_RECTIFY, 0)
        del cin, cout, douts

    if not args.only or "colour" in args.only:
        CW, CH, NB, NC = 1920, 1200, max(2, B // 2), 4  # the long-range cameras' colour stereo streams (config 4)
        csrc = SyntheticCameraSource(SyntheticCameraConfig(name="lr0", resolution=(CW, CH), pixel_format="bgr8", pool=1, enable_rgbd=False))
        cmaps = stereo_rectify_maps(csrc.get_intrinsics(), csrc.get_extrinsics(), (CW, CH))
        for cam in range(NC):
            ctx.upload_rectify_map(16 + cam, *cmaps[cam % 2], (CW, CH))
        print("plan", ctx.rectify_plan(16), flush=True)
        cin = [torch.randint(0, 256, (NB, CH, CW, 3), dtype=torch.uint8, device="cuda") for _ in range(NC)]
        cout = [torch.empty_like(t) for t in cin]
        gout = [torch.empty((NB, CH, CW), dtype=torch.uint8, device="cuda") for _ in range(NC)]
        cspecs = [StreamSpec(F.KIND_RECTIFY, cin[i], cout[i], F.BGR8, F.RGB8, camera=16 + i) for i in range(NC)]
        gspecs = [StreamSpec(F.KIND_RECTIFY, cin[i], gout[i], F.BGR8, F.MONO8, camera=16 + i) for i in range(NC)]
        cpx = NC * NB * CW * CH
        report("rectify bgr8->rgb8 1920x1200 (3-channel windows)", timeit(lambda: ctx.ingest(cspecs), args.iters), 6 * cpx, cpx)
        report("rectify bgr8->mono8 1920x1200 (gray pass + v4, L2-sized chunks)", timeit(lambda: ctx.ingest(gspecs), args.iters), 4 * cpx, cpx)
        ctx.set_option(ctx.OPT_L2_SCRATCH_KB, 4 << 20)
        report("  the same with the whole batch as one chunk (round 1)", timeit(lambda: ctx.ingest(gspecs), args.iters), 4 * cpx, cpx)
        ctx.set_option(ctx.OPT_L2_SCRATCH_KB, 0)
        nvin = [torch.randint(0, 256, (NB, CH * 3 // 2, CW), dtype=torch.uint8, device="cuda") for _ in range(NC)]
        nspecs = [StreamSpec(F.KIND_RECTIFY, nvin[i], cout[i], F.NV12, F.RGB8, camera=16 + i) for i in range(NC)]
        report("rectify nv12->rgb8 1920x1200 (bgr pass + 3-channel windows, L2-sized chunks)", timeit(lambda: ctx.ingest(nspecs), args.iters), 4.5 * cpx, cpx)
        ctx.set_option(ctx.OPT_L2_SCRATCH_KB, 4 << 20)
        report("  the same with the whole batch as one chunk (round 1)", timeit(lambda: ctx.ingest(nspecs), args.iters), 4.5 * cpx, cpx)
        ctx.set_option(ctx.OPT_L2_SCRATCH_KB, 0)
        for kb in (8 << 10, 16 << 10, 24 << 10, 64 << 10):
            ctx.set_option(ctx.OPT_L2_SCRATCH_KB, kb)
            report(f"  bgr8->mono8, {kb >> 10} MB of scratch per chunk", timeit(lambda: ctx.ingest(gspecs), args.iters), 4 * cpx, cpx)
        ctx.set_option(ctx.OPT_L2_SCRATCH_KB, 0)
        del nvin
        ctx.set_option(ctx.OPT_FORCE_GENERIC_RECTIFY, 1)
        report("rectify bgr8->rgb8 generic tiled", timeit(lambda: ctx.ingest(cspecs), max(3, args.iters // 4)), 6 * cpx, cpx)
        report("rectify bgr8->mono8 generic direct", timeit(lambda: ctx.ingest(gspecs), max(3, args.iters // 4)), 4 * cpx, cpx)
        ctx.set_option(ctx.OPT_FORCE_GENERIC_RECTIFY, 0)
        report("torch copy_ bgr (6 B/px)", timeit(lambda: [o.copy_(i) for o, i in zip(cout, cin)], args.iters), 6 * cpx, cpx)
        del cin, cout, gout

    if not args.only or "bp" in args.only:
        intr = src.get_intrinsics()[0]
        m = body_T_camera(None, src.get_extrinsics()[0].to_4x4_matrix(), "rdf")
        ND = 4
        for cam in range(ND):
            ctx.upload_projection(cam, intr.matrix, m, (W, H))
        d2 = np.stack([make_depth(rng, W, H) for _ in range(2)]).view(np.int16)
        depth = [torch.from_numpy(d2).cuda().view(torch.uint16).repeat((B + 1) // 2, 1, 1)[:B].contiguous() for _ in range(ND)]
        xyz = [torch.empty((B, H, W, 3), dtype=torch.float32, device="cuda") for _ in range(ND)]
        mask = [torch.empty((B, H, W), dtype=torch.uint8, device="cuda") for _ in range(ND)]
        cnt = [torch.zeros((B,), dtype=torch.int32, device="cuda") for _ in range(ND)]
        bspecs = [StreamSpec(F.KIND_BACKPROJECT, depth[i], xyz[i], F.DEPTH16, F.XYZ32F, camera=i, mask=mask[i], count=cnt[i]) for i in range(ND)]
        dpx = ND * B * W * H
        report("backproject depth->xyz+mask+count", timeit(lambda: ctx.ingest(bspecs), args.iters), 15 * dpx, dpx)
        report("torch copy_ xyz (24 B/px traffic)", timeit(lambda: [xyz[i].copy_(xyz[(i + 1) % ND]) for i in range(ND)], args.iters), 24 * dpx, dpx)
        report("torch fill_ xyz (12 B/px, write only)", timeit(lambda: [xyz[i].fill_(1.0) for i in range(ND)], args.iters), 12 * dpx, dpx)
        rgb_img = torch.randint(0, 256, (B, 1080, 1920, 3), dtype=torch.uint8, device="cuda")
        colour = torch.empty((B, H, W, 3), dtype=torch.uint8, device="cuda")
        ri, di = SyntheticCameraSource(SyntheticCameraConfig(name="r", enable_rgbd=True, rgb_resolution=(1920, 1080), depth_resolution=(W, H),
                                                             pool=1)).get_rgbd_intrinsics()
        t_rd = np.eye(4)
        t_rd[0, 3] = -0.0375
        ctx.upload_registration(0, di.matrix, (W, H), ri.matrix, (1920, 1080), t_rd)
        report("register depth->rgb colour (depth 2 + colour 3 B/px + the 1080p RGB image once)",
               timeit(lambda: ctx.register_colour(0, depth[0], rgb_img, colour), args.iters), 5 * B * W * H + B * 1920 * 1080 * 3, B * W * H)
        xyz1 = torch.empty((B, H, W, 3), dtype=torch.float32, device="cuda")
        report("backproject + colour fused (2 in + 12 + 1 + 3 out = 18 B/px; the RGB image read on top)",
               timeit(lambda: ctx.backproject_colour(0, depth[0], rgb_img, xyz1, colour, mask[0], cnt[0]), args.iters), 18 * B * W * H, B * W * H)
        report("  the same as two kernels (backproject, then register colour)",
               timeit(lambda: (ctx.backproject(0, depth[0], xyz1, mask[0], cnt[0]), ctx.register_colour(0, depth[0], rgb_img, colour)), args.iters),
               18 * B * W * H, B * W * H)
        del rgb_img, colour, xyz1
        for per_sm in (2, 3, 4, 6, 8):
            ctx.set_option(ctx.OPT_CTAS_PER_SM, per_sm)
            report(f"backproject ctas/sm={per_sm}", timeit(lambda: ctx.ingest(bspecs), args.iters), 15 * dpx, dpx)
        ctx.set_option(ctx.OPT_CTAS_PER_SM, 0)
        del xyz, mask, depth

    if not args.only or "voxel" in args.only:
        intr = src.get_intrinsics()[0]
        m = body_T_camera(None, src.get_extrinsics()[0].to_4x4_matrix(), "rdf")
        ND, VB = 4, max(2, B // 2)
        for cam in range(ND):
            ctx.upload_projection(cam, intr.matrix, m, (W, H))
        ctx.set_voxel_grid(0.05, 10000)
        vpx = ND * VB * W * H
        rec = torch.empty(vpx, dtype=torch.int64, device="cuda")
        nrec = torch.zeros(1, dtype=torch.int32, device="cuda")
        cnts = torch.zeros(VB, dtype=torch.int32, device="cuda")
        for scene, gen in (("room", lambda: make_depth_scene(rng, W, H, focal_px=intr.matrix[0, 0])), ("noise", lambda: make_depth(rng, W, H))):
            d4 = np.stack([gen() for _ in range(4)]).view(np.int16)
            vdepth = [torch.from_numpy(np.roll(d4, i, axis=0)).cuda().view(torch.uint16).repeat((VB + 3) // 4, 1, 1)[:VB].contiguous() for i in range(ND)]
            streams = [(i, vdepth[i]) for i in range(ND)]
            for per_sm, dbg in ((0, 0), (0, 8), (0, 0), (0, 8), (0, 4), (0, 1), (0, 2), (0, 3))[:4 if args.one else None]:
                if scene == "noise" and dbg in (2, 3):
                    continue
                ctx.set_option(ctx.OPT_CTAS_PER_SM, per_sm)
                ctx.set_option(ctx.OPT_DEBUG, dbg)
                ms = timeit(lambda: ctx.voxel_cloud(streams, rec, nrec, cnts), 3 if scene == "noise" else args.iters)
                n = int(nrec.item())
                report(f"voxel cloud {scene} ctas/sm={per_sm} debug={dbg}: {n} voxels of {vpx} px (x{vpx * 0.8 / max(n, 1):.1f})", ms, 2 * vpx + 8 * n, vpx)
            ctx.set_option(ctx.OPT_CTAS_PER_SM, 0)
            ctx.set_option(ctx.OPT_DEBUG, 0)
            del vdepth
        del rec

    if not args.only or "conv" in args.only:
        CW, CH, NB = 1920, 1080, max(2, B)  # >= 32 frames: 199 MB in, far beyond L2
        bgr = torch.randint(0, 256, (NB, CH, CW, 3), dtype=torch.uint8, device="cuda")
        rgb = torch.empty_like(bgr)
        gray = torch.empty((NB, CH, CW), dtype=torch.uint8, device="cuda")
        nv = torch.randint(0, 256, (NB, CH * 3 // 2, CW), dtype=torch.uint8, device="cuda")
        cpx = NB * CW * CH
        report("convert bgr8->rgb8 1080p", timeit(lambda: ctx.convert(bgr, rgb, "bgr8", "rgb8", CW, CH), args.iters), 6 * cpx, cpx)
        report("convert bgr8->mono8 1080p", timeit(lambda: ctx.convert(bgr, gray, "bgr8", "mono8", CW, CH), args.iters), 4 * cpx, cpx)
        report("convert nv12->rgb8 1080p", timeit(lambda: ctx.convert(nv, rgb, "nv12", "rgb8", CW, CH), args.iters), 4.5 * cpx, cpx)
        report("convert nv12->mono8 1080p", timeit(lambda: ctx.convert(nv, gray, "nv12", "mono8", CW, CH), args.iters), 2 * cpx, cpx)
        report("torch copy_ bgr (6 B/px)", timeit(lambda: rgb.copy_(bgr), args.iters), 6 * cpx, cpx)
    ctx.close()


if __name__ == "__main__":
    main()
