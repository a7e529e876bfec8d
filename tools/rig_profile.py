#!/usr/bin/env python
"""Where one ``IngestRig.get_synchronized_frames()`` + ``np.asarray`` of all frames spends its time (B200 box):
``python tools/rig_profile.py`` prints the per-call latency and cProfile's top functions by cumulative time."""
from __future__ import annotations

import cProfile
import pstats
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main() -> None:
    import torch

    from thor_slam_b200.camera.synthetic import make_rig_sources
    from thor_slam_b200.ingest.rig import IngestRig

    sources = make_rig_sources(4, resolution=(1280, 800), pixel_format="mono8", seed=1337, pool=2)
    rig = IngestRig(sources, queue_size=10)
    rig.start()

    def loop(n: int) -> list[float]:
        lat = []
        for _ in range(n):
            t1 = time.perf_counter()
            fs = rig.get_synchronized_frames()
            t2 = time.perf_counter()
            [np.asarray(f.image) for f in fs.get_all_frames()]
            t3 = time.perf_counter()
            lat.append((t2 - t1, t3 - t2))
        return lat

    loop(20)
    torch.cuda.synchronize()
    lat = np.array(loop(300))
    print(f"get_synchronized_frames {np.median(lat[:, 0]) * 1e3:.3f} ms, np.asarray x 8 {np.median(lat[:, 1]) * 1e3:.3f} ms, "
          f"total {np.median(lat.sum(1)) * 1e3:.3f} ms (medians of 300 calls)")
    pr = cProfile.Profile()
    pr.enable()
    loop(300)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
    rig.stop()


if __name__ == "__main__":
    main()
