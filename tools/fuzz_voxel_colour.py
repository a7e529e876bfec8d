#!/usr/bin/env python
"""Randomised parity sweep of the round-2 kernels against the oracle on the GPU (or the CPU emulation).

    python tools/fuzz_voxel_colour.py [--cases 100] [--seed 1] [--emu]

* ``ti_voxel_cloud``: 1-4 cameras of any size (vector and scalar depth loads, ragged tiles), 0-3 frame sets, surfaces or noise,
  voxel sizes 2 cm - 20 cm, with and without the depth cap, random set base / tag: record SET, total and per-set counts exact,
  ``ti_voxel_points`` bit for bit (tests/cases.py:check_voxel).
* ``ti_backproject_colour``: any depth / RGB size, random registration: cloud within 1e-5, mask / count exact, colours identical
  to the float32 oracle and to the stand-alone kernel (check_backproject_colour).
* ``ti_depth_stats``: exact integers.
Development tool - imports the test helpers, hence the oracle.
"""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
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
    t0, n_vox, n_col, records, n_refused = time.time(), 0, 0, 0, 0
    top = 160 if args.emu else 900
    for i in range(args.cases):
        kind = i % 3
        try:
            if kind < 2:
                sizes = [(int(rng.integers(9, top)) if rng.random() < 0.5 else int(rng.integers(2, top // 8)) * 8, int(rng.integers(3, top // 2)))
                         for _ in range(int(rng.integers(1, 5)))]
                records += cases.check_voxel(be, 40, sizes, n=int(rng.integers(0, 4)), seed=args.seed * 1000 + i, voxel=float(rng.choice([0.02, 0.05, 0.05, 0.2])),
                                             max_depth_mm=int(rng.choice([0, 10000, 4000])), scene=str(rng.choice(["room", "noise"])),
                                             set_base=int(rng.integers(0, 2040)), tag=int(rng.integers(0, 256)))
                n_vox += 1
            else:
                w = int(rng.integers(2, top // 8)) * 8 if rng.random() < 0.7 else int(rng.integers(9, top))
                h = int(rng.integers(3, top // 2))
                cases.check_backproject_colour(be, 44, w, h, int(rng.integers(16, 2 * top)), int(rng.integers(8, top)), n=int(rng.integers(1, 3)), seed=args.seed * 1000 + i)
                cases.check_depth_stats(be, w, h, n=2, seed=args.seed * 1000 + i)
                n_col += 1
        except ValueError as e:  # a 2 cm grid, no depth cap, a wide camera far from the origin: the library says the 15-bit fields do not reach
            if "key fields" not in str(e):
                raise
            n_refused += 1
        except AssertionError as e:
            print(f"case {i} (kind {kind}): MISMATCH {e}", flush=True)
            raise SystemExit(1)
    print(f"{n_vox} voxel-cloud cases ({records} records, every set exact), {n_col} fused colour + depth-stats cases: all match the oracle ({n_refused} grids refused as out of key range), {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
