#!/usr/bin/env python
"""Randomised parity sweep of the conversion and back-projection kernels against the oracle on the GPU.

    python tools/fuzz_convert_backproject.py [--cases 150] [--seed 1] [--emu]

Conversions: every supported pair at any width / height (odd widths too; NV12 needs even sizes), batches 0..3.
Back-projection: any size, random rig pose, RDF and FLU rig frames, random depth with holes and saturated pixels;
points within 1e-5 relative, masks and counts exact (tests/cases.py).  Development tool - imports the oracle.
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


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=150)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--emu", action="store_true")
    args = ap.parse_args()
    from tests import cases
    from tests.conftest import Backend
    from thor_slam_b200.ingest.context import IngestContext

    if args.emu:
        import ctypes

        from tests.emu.build_emu import build
        from thor_slam_b200.ingest._lib import IngestLibrary

        be = Backend("emu", IngestContext(0, IngestLibrary(ctypes.CDLL(str(build())))))
    else:
        import torch

        torch.cuda.set_device(0)
        ctx = IngestContext(0)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        be = Backend("gpu", ctx)
    rng = np.random.default_rng(args.seed)
    done: collections.Counter = collections.Counter()
    t0 = time.time()
    wmax, hmax = (200, 60) if args.emu else (2048, 1300)
    for i in range(args.cases):
        big = rng.random() < 0.2
        w = int(rng.integers(1, wmax if big else max(2, wmax // 4)))
        h = int(rng.integers(1, hmax if big else max(2, hmax // 4)))
        n = int(rng.integers(0, 4))
        what = "bp" if rng.random() < 0.4 else "conv"
        try:
            if what == "conv":
                s, d = cases.CONVERSIONS[int(rng.integers(len(cases.CONVERSIONS)))]
                if s == "nv12":
                    w, h = max(2, w & ~1), max(2, h & ~1)
                cases.check_convert(be, s, d, w, h, n=n, seed=args.seed * 1000 + i)
                done[f"convert {s}->{d}"] += 1
            else:
                frame = "flu" if rng.random() < 0.5 else "rdf"
                depth = None
                if n and rng.random() < 0.3:
                    depth = rng.integers(0, 65536, size=(n, h, w)).astype(np.uint16)  # full range, few holes
                cases.check_backproject(be, 21, w, h, n=n, seed=args.seed * 1000 + i, rig_frame=frame, depth=depth)
                done[f"backproject ({frame})"] += 1
        except AssertionError as e:
            print(f"case {i}: {what} {w}x{h} n={n}: MISMATCH {e}", flush=True)
            raise SystemExit(1)
    print(f"{args.cases} cases match the oracle, {time.time() - t0:.0f} s")
    for k, v in sorted(done.items()):
        print(f"  {k}: {v} cases")


if __name__ == "__main__":
    main()
