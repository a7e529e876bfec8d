"""Smallest possible run of the TMA-pipelined rectify kernel (for compute-sanitizer / debugging)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from thor_slam_b200.ingest.context import IngestContext  # noqa: E402

ctx = IngestContext(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
W, H = 256, 64
yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
ctx.upload_rectify_map(0, xx * 0.97 + 1.3, yy * 0.98 + 0.6, (W, H))
src = torch.randint(0, 256, (2, H, W), dtype=torch.uint8, device="cuda")
dst = torch.zeros_like(src)
ctx.set_option(ctx.OPT_MONO_VARIANT, int(sys.argv[1]) if len(sys.argv) > 1 else 3)
ctx.set_option(ctx.OPT_DEBUG, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
ctx.rectify(0, src, dst, "mono8", "mono8")
torch.cuda.synchronize()
print("ok", int(dst.sum()))
