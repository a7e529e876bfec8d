#!/usr/bin/env python
"""Host<->device copy ceilings of this box (pinned memory): H2D alone, D2H alone, both at once.

    python tools/pcie_probe.py [--mb 256]

The e2e number of bench.py moves 1 B/px in and 1 B/px out through ``ti_ingest_host``; this is the
ceiling that path can reach.  Development tool.
"""
from __future__ import annotations

import argparse
import time

import torch


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--pieces", type=int, default=1, help="split every copy into this many back-to-back pieces")
    ap.add_argument("--streams", type=int, default=1, help="round-robin the pieces of one direction over this many streams")
    args = ap.parse_args()
    n = args.mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    up = [torch.cuda.Stream() for _ in range(args.streams)]
    down = [torch.cuda.Stream() for _ in range(args.streams)]
    step = n // args.pieces

    def run(h2d: bool, d2h: bool) -> float:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.iters):
            for k in range(args.pieces):
                sl = slice(k * step, (k + 1) * step)
                if h2d:
                    with torch.cuda.stream(up[k % args.streams]):
                        d_a[sl].copy_(h_in[sl], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(down[k % args.streams]):
                        h_out[sl].copy_(d_b[sl], non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / args.iters

    def run_dependent() -> float:
        """Piece k comes back only after it went up (the shape of ti_ingest_host: upload -> kernel -> download per chunk)."""
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.iters):
            for k in range(args.pieces):
                sl = slice(k * step, (k + 1) * step)
                with torch.cuda.stream(up[0]):
                    d_a[sl].copy_(h_in[sl], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                with torch.cuda.stream(down[0]):
                    down[0].wait_event(ev)
                    h_out[sl].copy_(d_a[sl], non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / args.iters

    run_dependent()
    dt = run_dependent()
    print(f"{'up, then down, per piece':26s} {n / dt / 1e9:7.1f} GB/s per direction ({args.mb} MB in {args.pieces} piece(s), {dt * 1e3:.2f} ms)")
    for name, a, b in (("H2D alone", True, False), ("D2H alone", False, True), ("H2D + D2H concurrently", True, True)):
        run(a, b)
        dt = run(a, b)
        print(f"{name:26s} {n / dt / 1e9:7.1f} GB/s per direction ({args.mb} MB in {args.pieces} piece(s) over {args.streams} stream(s), {dt * 1e3:.2f} ms)")


if __name__ == "__main__":
    main()
