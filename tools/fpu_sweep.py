import sys, json
sys.path.insert(0, ".")
import numpy as np, torch
from tools.kbench import timeit
from thor_slam_b200.camera.synthetic import SyntheticCameraConfig, SyntheticCameraSource
from thor_slam_b200.ingest import formats as F
from thor_slam_b200.ingest.calib import stereo_rectify_maps
from thor_slam_b200.ingest.context import IngestContext, StreamSpec
import bench
torch.cuda.set_device(0)
ctx = IngestContext(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
sources, maps = bench.build_rig()
for cam, (mx, my) in enumerate(maps): ctx.upload_rectify_map(cam, mx, my, (1280, 800))
B = 64
pf = bench.host_frames(sources, 2)
src = [torch.from_numpy(pf[s]).cuda().repeat(B // 2, 1, 1).contiguous() for s in range(8)]
dst = [torch.empty_like(t) for t in src]
specs = [StreamSpec(F.KIND_RECTIFY, src[s], dst[s], F.MONO8, F.MONO8, camera=s) for s in range(8)]
px = 8 * B * 1280 * 800
import collections
res = collections.defaultdict(list)
cfgs = [(st, f) for st in (3, 4) for f in (0, 16, 22, 32, 64)]
for rep in range(4):
    for st, f in cfgs:
        ctx.set_option(ctx.OPT_FRAMES_PER_UNIT, f); ctx.set_option(ctx.OPT_STAGES, st)
        res[(st, f)].append(timeit(lambda: ctx.ingest(specs), 40, warm=3))
for (st, f), v in res.items():
    ms = sorted(v)[len(v) // 2]
    print(f"S={st} fpu={f:3d}  median {ms:.4f} ms  frac {2*px/(ms*1e-3)/1e9/6454.3:.3f}   all {[round(x, 4) for x in v]}", flush=True)
