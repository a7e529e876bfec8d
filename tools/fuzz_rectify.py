#!/usr/bin/env python
"""Randomised parity sweep of the rectify kernels against the oracle (cv2.remap) on the GPU.

    python tools/fuzz_rectify.py [--cases 120] [--seed 1]

Every case draws an output size (any width / height, not only multiples of the tile), a different source size, a map
(scale, shear, roll, barrel distortion, offsets that push source boxes over every border), a conversion and a batch
size, runs every kernel variant the slot qualifies for (tests/cases.py:check_rectify) and compares bytes.
Development tool: it imports the test helpers, hence the oracle - never part of the product path.
"""
from __future__ import annotations

import argparse
import collections
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def random_map(rng: np.random.Generator, dst_w: int, dst_h: int, src_w: int, src_h: int):
    yy, xx = np.mgrid[0:dst_h, 0:dst_w].astype(np.float64)
    cx, cy = dst_w / 2 + rng.uniform(-20, 20), dst_h / 2 + rng.uniform(-20, 20)
    sx = src_w / dst_w * rng.uniform(0.85, 1.2)
    sy = src_h / dst_h * rng.uniform(0.85, 1.2)
    roll = np.deg2rad(rng.uniform(-3, 3) if rng.random() < 0.8 else rng.uniform(-30, 30))
    k1 = rng.uniform(-0.3, 0.3)
    x, y = (xx - cx) / dst_w, (yy - cy) / dst_w
    f = 1 + k1 * (x * x + y * y)
    xd, yd = x * f, y * f
    c, sn = np.cos(roll), np.sin(roll)
    mx = (xd * c - yd * sn) * dst_w * sx + src_w / 2 + rng.uniform(-15, 15) + rng.uniform(-0.05, 0.05) * yy
    my = (xd * sn + yd * c) * dst_w * sy + src_h / 2 + rng.uniform(-15, 15)
    return mx.astype(np.float32), my.astype(np.float32)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=120)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--emu", action="store_true", help="run the CPU emulation build of the kernels (tests/emu), small sizes only")
    args = ap.parse_args()
    from tests import cases
    from tests.conftest import Backend
    from thor_slam_b200.ingest.context import IngestContext

    if args.emu:
        import ctypes

        from tests.emu.build_emu import build
        from thor_slam_b200.ingest._lib import IngestLibrary

        ctx = IngestContext(0, IngestLibrary(ctypes.CDLL(str(build()))))
        be = Backend("emu", ctx)
    else:
        import torch

        torch.cuda.set_device(0)
        ctx = IngestContext(0)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        be = Backend("gpu", ctx)
    rng = np.random.default_rng(args.seed)
    convs = [("mono8", "mono8")] * 5 + [("bgr8", "rgb8")] * 2 + [("bgr8", "mono8"), ("nv12", "mono8"), ("nv12", "rgb8")]
    plans: collections.Counter = collections.Counter()
    overflow: list[int] = []
    t0 = time.time()
    for i in range(args.cases):
        s, d = convs[int(rng.integers(len(convs)))]
        big = rng.random() < 0.25 and not args.emu
        dst_w = int(rng.integers(33, 2048 if big else (300 if args.emu else 700)))
        dst_h = int(rng.integers(9, 1300 if big else (70 if args.emu else 300)))
        ratio = rng.choice([1.0, 1.0, 0.5, 2.0, 1.5])
        src_w = min(4224, max(16, int(dst_w * ratio) + int(rng.integers(-5, 6))))  # up to the driver's largest sensor mode (the library takes 8190)
        src_h = min(3136, max(8, int(dst_h * ratio) + int(rng.integers(-5, 6))))
        if rng.random() < 0.8:
            src_w = max(16, src_w & ~15)  # the TMA kernels need a 16-byte row pitch (every sensor mode has one)
        if s == "nv12":
            src_w, src_h = src_w & ~1, src_h & ~1
        mx, my = random_map(rng, dst_w, dst_h, src_w, src_h)
        n = int(rng.integers(1, 4))
        try:
            cases.check_rectify(be, 20, mx, my, s, d, src_w, src_h, n=n, seed=args.seed * 1000 + i)
        except AssertionError as e:
            print(f"case {i}: {s}->{d} dst {dst_w}x{dst_h} src {src_w}x{src_h} n={n}: MISMATCH {e}", flush=True)
            raise SystemExit(1)
        p = ctx.rectify_plan(20)
        plans[(s, d, (f"{p['variant']} ({p['pixels_per_window']} px/window)" if p["variant"] == 4 else p["variant"]) if d == "mono8" else p["colour_variant"])] += 1
        if p["overflow_pixels"]:
            overflow.append(p["overflow_pixels"])
    print(f"{args.cases} cases, every variant bit-exact against cv2.remap, {time.time() - t0:.0f} s")
    print(f"  {len(overflow)} slots ran the pair-window kernel WITH an overflow list (up to {max(overflow, default=0)} pixels repaired after the kernel)")
    for k, v in sorted(plans.items(), key=str):
        print(f"  {k[0]:>5s} -> {k[1]:<5s} default kernel variant {k[2]}: {v} cases")


if __name__ == "__main__":
    main()
