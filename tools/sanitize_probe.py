"""Smallest run of every kernel family (for compute-sanitizer / debugging):

    compute-sanitizer --tool memcheck python tools/sanitize_probe.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from thor_slam_b200.ingest.context import IngestContext  # noqa: E402

ctx = IngestContext(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
W, H = 384, 96
yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
mapx, mapy = xx * 1.04 - 9.3 + 0.03 * yy, yy * 1.07 - 5.6 - 0.02 * xx  # leaves the image on every side, exceptions on most rows
ctx.upload_rectify_map(0, mapx, mapy, (W, H))
print("plan", ctx.rectify_plan(0))
for n in (1, 3):
    mono = torch.randint(0, 256, (n, H, W), dtype=torch.uint8, device="cuda")
    out = torch.zeros_like(mono)
    for variant in (4, 3, 2, 1):
        ctx.set_option(ctx.OPT_MONO_VARIANT, variant)
        ctx.rectify(0, mono, out, "mono8", "mono8")
    ctx.set_option(ctx.OPT_MONO_VARIANT, 4)
    bgr = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
    rgb = torch.zeros_like(bgr)
    ctx.rectify(0, bgr, rgb, "bgr8", "rgb8")      # 3-channel window kernel
    ctx.rectify(0, bgr, out, "bgr8", "mono8")     # gray pass + pair-window kernel
    nv12 = torch.randint(0, 256, (n, H * 3 // 2, W), dtype=torch.uint8, device="cuda")
    ctx.rectify(0, nv12, rgb, "nv12", "rgb8")     # conversion pass + 3-channel window kernel
    ctx.convert(nv12, rgb, "nv12", "rgb8", W, H)
    k = np.array([[300.0, 0, W / 2], [0, 300.0, H / 2], [0, 0, 1]])
    ctx.upload_projection(0, k, np.eye(4), (W, H))
    ctx.upload_registration(0, k, (W, H), k, (W, H), np.eye(4))
    depth = torch.randint(0, 5000, (n, H, W), dtype=torch.int32, device="cuda").to(torch.int16).view(torch.uint16)
    xyz = torch.zeros((n, H, W, 3), dtype=torch.float32, device="cuda")
    mask = torch.zeros((n, H, W), dtype=torch.uint8, device="cuda")
    count = torch.zeros((n,), dtype=torch.int32, device="cuda")
    ctx.backproject(0, depth, xyz, mask, count)
    colour = torch.zeros((n, H, W, 3), dtype=torch.uint8, device="cuda")
    ctx.register_colour(0, depth, rgb, colour)
    ctx.sync()
torch.cuda.synchronize()
print("ok", int(out.sum()), int(rgb.sum()), int(count.sum()), int(colour.sum()))
ctx.close()
