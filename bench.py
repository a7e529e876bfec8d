#!/usr/bin/env python
"""Benchmark of the ingest hot path (BASELINE.json metric: frame-sets/s, 8 x 1280x800).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)

Workload = BASELINE config 2: a 4 x OAK-D Pro stereo rig, 8 mono8 streams of 1280x800 per frame set,
format conversion (mono8 pass-through) fused with the stereo-rectification remap.  One *step* is one
pass of the hot path over a batch of ``--batch`` frame sets (default 64 = 524 MB in + 524 MB out per
step, far larger than the 126 MB L2, so no step is served from cache).

* ``value``   : frame-sets/s with inputs resident in HBM (CUDA events on the launch stream).
* ``e2e``     : same metric through ``IngestContext.ingest_host_submit`` / ``ingest_host_wait`` - pinned HOST buffers
                in and out, host->device and device->host copies of every step inside the timed region, two
                batches in flight; ``blocking_call_value`` is the same loop through the blocking ``ingest_host``.
* ``roofline``: algorithmic bytes (2 B/px: 1 read + 1 written) of one launch
                of the rectify kernel / its average duration, against MEASURED_PEAKS.json ``hbm_gbs``.
* ``cpu_baseline``: the oracle (cv2.remap, all host threads) on a bounded sample, rank 0, N=1 only.

With N > 1 every rank processes its own batch (weak scaling, no data-path collective - config 2 has no
exchange step); ``extras.config5`` additionally times the 4 x (mono rectify + depth -> cloud) frame set
and the NCCL gather of the clouds separately.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

W, H = 1280, 800
N_CAMERAS = 4  # stereo sources -> 8 streams
STREAMS = 2 * N_CAMERAS
PX_PER_SET = STREAMS * W * H
ALGO_BYTES_PER_PX = 2  # mono8 -> rectified mono8
# dram__bytes_read.sum + dram__bytes_write.sum of ONE rectify_mono_pair_kernel launch over 64 frame sets, from the
# `ncu --set full` capture of this bench (profiles/r01_rect_v4_ncu_raw.txt): 584.93 MB + 485.15 MB
NCU_TRAFFIC_BYTES_PER_FRAME_SET = (584.931840e6 + 485.148160e6) / 64
KERNEL_BY_VARIANT = {4: "rectify_mono_pair_kernel<32,false,1280>", 3: "rectify_mono_tma_kernel<32,false>", 2: "rectify_mono_kernel", 1: "rectify_tile_kernel<1>"}
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peak() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback 6650 GB/s (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / power / throttle reasons of one GPU, polled through NVML on a thread during the timed region."""

    REASONS = {
        "hw_slowdown": 0x8,            # nvmlClocksThrottleReasonHwSlowdown
        "sw_power_cap": 0x4,           # nvmlClocksThrottleReasonSwPowerCap
        "sw_thermal_slowdown": 0x20,   # nvmlClocksThrottleReasonSwThermalSlowdown
        "hw_thermal_slowdown": 0x40,   # nvmlClocksThrottleReasonHwThermalSlowdown
        "hw_power_brake_slowdown": 0x80,
    }

    def __init__(self, index: int, period_s: float = 0.004) -> None:
        self.index, self.period = index, period_s
        self.samples: list[tuple[float, int, float, int]] = []  # (t, sm_mhz, power_w, reasons)
        self._stop = threading.Event()
        self._thread: threading.Thread | None = None
        self.sm_max = None
        self.error: str | None = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # pragma: no cover - depends on the box
            self._nv = None
            self.error = f"NVML unavailable: {exc}"

    def _poll(self) -> None:
        nv = self._nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.samples.append((time.perf_counter(), int(mhz), float(pw), int(rs)))
            except Exception as exc:  # pragma: no cover
                self.error = str(exc)
                return
            self._stop.wait(self.period)

    def start(self) -> None:
        if self._nv is None:
            return
        self._thread = threading.Thread(target=self._poll, daemon=True)
        self._thread.start()

    def stop(self, t0: float | None = None, t1: float | None = None) -> dict:
        """Summary over the samples taken in [t0, t1] (perf_counter times); all samples if the window holds none."""
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=1.0)
        if self._nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [self.error or "no samples"], "samples": 0}
        inside = [x for x in self.samples if (t0 is None or x[0] >= t0) and (t1 is None or x[0] <= t1)]
        window = "timed region"
        if not inside:
            inside, window = self.samples, "warm-up + timed region (timed region shorter than one NVML poll)"
        reasons = set()
        for _, _, _, rs in inside:
            for name, bit in self.REASONS.items():
                if rs & bit:
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median([x[1] for x in inside])),
            "sm_max_mhz": self.sm_max,
            "power_w_max": max(x[2] for x in inside),
            "samples": len(inside),
            "window": window,
            "reasons": sorted(reasons),
        }


# ------------------------------------------------------------------------------------------------
# workload construction (shared by both arms)
# ------------------------------------------------------------------------------------------------
def build_rig(seed: int = 1337):
    """4 synthetic OAK-D Pro stereo sources + their rectification maps (host float32)."""
    from thor_slam_b200.camera.synthetic import make_rig_sources
    from thor_slam_b200.ingest.calib import stereo_rectify_maps

    sources = make_rig_sources(N_CAMERAS, resolution=(W, H), pixel_format="mono8", seed=seed, pool=2)
    maps = []
    for s in sources:
        maps.extend(stereo_rectify_maps(s.get_intrinsics(), s.get_extrinsics(), (W, H)))
    return sources, maps  # maps[2*i + {0,1}] = (mapx, mapy) of source i left/right


def host_frames(sources, n_sets: int) -> list[np.ndarray]:
    """Per stream, ``n_sets`` frames [n_sets, H, W] u8 cycling through each source's seeded pool."""
    out = []
    for s in sources:
        pool = s._pool
        for cam in range(2):
            out.append(np.stack([pool[b % len(pool)][cam] for b in range(n_sets)]))
    return out


def cpu_pass(frames: list[np.ndarray], maps, n_sets: int) -> None:
    """The reference-side CPU path for this workload: cv2.remap per stream per frame set (oracle)."""
    from oracle import rectify as orc

    for b in range(n_sets):
        for s in range(STREAMS):
            orc.remap_cv(frames[s][b], maps[s][0], maps[s][1])


def time_cpu(frames, maps, budget_s: float, max_sets: int) -> tuple[float, int, float]:
    """(frame-sets/s, sets processed, seconds) on a bounded sample."""
    cpu_pass(frames, maps, 1)  # warm OpenCV's thread pool
    done, t0 = 0, time.perf_counter()
    while done < max_sets and (time.perf_counter() - t0) < budget_s:
        cpu_pass([f[done % f.shape[0]: done % f.shape[0] + 1] for f in frames], maps, 1)
        done += 1
    dt = time.perf_counter() - t0
    return done / dt, done, dt


# ------------------------------------------------------------------------------------------------
def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import cv2

    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)
    sources, maps = build_rig()
    sets_per_step = max(1, args.ref_sets)
    frames = host_frames(sources, min(sets_per_step, 4))
    for _ in range(args.warmup):
        cpu_pass([f[:1] for f in frames], maps, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for b in range(sets_per_step):
            cpu_pass([f[b % f.shape[0]: b % f.shape[0] + 1] for f in frames], maps, 1)
    dt = time.perf_counter() - t0
    value = args.steps * sets_per_step / dt
    line = {
        "impl": "reference",
        "metric": "frame_sets_per_sec",
        "value": value,
        "unit": "frame-sets/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u8",
        "data": "synthetic",
        "config": workload_config(sets_per_step, args.gpus) | {"note": "CPU path: cv2.remap INTER_LINEAR per stream (oracle port of the reference-side path), all host threads"},
        "mpix_per_sec": value * PX_PER_SET / 1e6,
        "cpu_baseline": {"value": value, "unit": "frame-sets/s", "cores": cv2.getNumThreads(), "kind": "port",
                         "sample": f"{args.steps} steps x {sets_per_step} frame sets of 8 x 1280x800 mono8, cv2.remap"},
        "e2e": {"value": value, "unit": "frame-sets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(batch: int, gpus: int) -> dict:
    return {
        "workload": "BASELINE config 2: 4x OAK-D Pro stereo rig, 8 mono8 streams 1280x800, convert+rectify (stereoRectify maps, rational-8 distortion)",
        "streams": STREAMS,
        "width": W,
        "height": H,
        "frame_sets_per_step": batch,
        "sharding": f"frame-set batches per rank x{gpus} (no data-path collective)",
        "l2_policy": "inputs larger than L2 (batch in+out >= 1 GB vs 126 MB L2); no flush needed",
    }


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from thor_slam_b200.ingest import formats as F
    from thor_slam_b200.ingest.context import IngestContext, StreamSpec

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier() -> None:
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = IngestContext(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    sources, maps = build_rig()
    for cam, (mx, my) in enumerate(maps):
        ctx.upload_rectify_map(cam, mx, my, (W, H))

    plan = ctx.rectify_plan(0)
    if plan["variant"] != 4:
        raise SystemExit(f"bench: the benchmark rig must run the pair-window kernel, got {plan}")
    B = args.batch
    pool_frames = host_frames(sources, 2)  # two distinct frames per stream, tiled over the batch on device
    d_src, d_dst = [], []
    for s in range(STREAMS):
        pf = torch.from_numpy(pool_frames[s]).cuda()
        d_src.append(pf.repeat((B + 1) // 2, 1, 1)[:B].contiguous())
        d_dst.append(torch.empty((B, H, W), dtype=torch.uint8, device="cuda"))
    specs = [StreamSpec(F.KIND_RECTIFY, d_src[s], d_dst[s], F.MONO8, F.MONO8, camera=s) for s in range(STREAMS)]

    # ---- device-resident throughput ("value") ---------------------------------------------------
    for _ in range(args.warmup):
        ctx.ingest(specs)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        ctx.ingest(specs)
    barrier()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        ctx.ingest(specs)
    ev1.record(stream)
    barrier()
    t_end = time.perf_counter()
    launches = ctx.launch_count - launches0
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = world * B * args.steps / (ms_max * 1e-3)

    got_value_frame = d_dst[3][1].cpu().numpy() if rank == 0 else None  # checked against the oracle in the cpu_baseline leg

    # ---- end to end through the host-buffer API ("e2e") -----------------------------------------
    # A capture loop double-buffers its host frames: batch k+1 is submitted before batch k is waited for, so one
    # step's download overlaps the next step's upload.  Every step still uploads its own inputs from pinned host
    # memory and reads its own result back into pinned host memory inside the timed region.  The same loop through
    # the blocking call (one batch at a time, nothing overlapped across calls) is reported next to it.
    Be = min(B, args.e2e_batch)
    from thor_slam_b200.ingest.hostmem import near_gpu

    with near_gpu(local_rank) as place:  # pinned pages on the GPU's own NUMA node (matters with one process per GPU)
        if args.no_numa:
            os.sched_setaffinity(0, place.before)
        h_src = [[torch.from_numpy(np.ascontiguousarray(np.tile(pool_frames[s], ((Be + 1) // 2, 1, 1))[:Be])).pin_memory() for s in range(STREAMS)]
                 for _ in range(2)]
        h_dst = [[torch.zeros((Be, H, W), dtype=torch.uint8).pin_memory() for _ in range(STREAMS)] for _ in range(2)]
    host_cpus = len(place.cpus)
    hspecs = [[StreamSpec(F.KIND_RECTIFY, h_src[k][s], h_dst[k][s], F.MONO8, F.MONO8, camera=s) for s in range(STREAMS)] for k in range(2)]
    e2e_steps = max(4, min(args.steps, 10))

    def e2e_loop(steps: int, blocking: bool) -> None:
        prev = None
        for k in range(steps):
            if blocking:
                ctx.ingest_host(hspecs[k % 2], chunk=args.chunk)
                continue
            ticket = ctx.ingest_host_submit(hspecs[k % 2], chunk=args.chunk)
            if prev is not None:
                ctx.ingest_host_wait(prev)  # buffers k-1 are the caller's again before they are resubmitted as k+1
            prev = ticket
        if prev is not None:
            ctx.ingest_host_wait(prev)

    def e2e_time(blocking: bool) -> float:
        e2e_loop(2, blocking)
        barrier()
        t0 = time.perf_counter()
        e2e_loop(e2e_steps, blocking)
        torch.cuda.synchronize()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if distributed:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item())

    e2e_blocking_value = world * Be * e2e_steps / e2e_time(blocking=True)
    for k in range(2):
        for t in h_dst[k]:
            t.zero_()
    launches_e2e0 = ctx.launch_count
    e2e_value = world * Be * e2e_steps / e2e_time(blocking=False)
    launches_e2e = (ctx.launch_count - launches_e2e0) * e2e_steps // (e2e_steps + 2)  # the two warm-up steps launch too
    got_e2e_frame = h_dst[(e2e_steps - 1) % 2][5][0].numpy().copy() if rank == 0 else None

    extras: dict = {}
    if args.extras:
        extras = run_extras(args, ctx, sources, maps, rank, world, distributed, barrier)

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import cv2

        cv2.setNumThreads(os.cpu_count() or 1)
        # the oracle runs only in this leg: it is the CPU baseline and, on the same frames, the checker of both GPU numbers
        from oracle import rectify as orc

        if not np.array_equal(got_value_frame, orc.remap_cv(pool_frames[3][1], maps[3][0], maps[3][1])):
            raise SystemExit("bench: GPU output differs from cv2.remap - refusing to report a number")
        if not np.array_equal(got_e2e_frame, orc.remap_cv(pool_frames[5][0], maps[5][0], maps[5][1])):
            raise SystemExit("bench: e2e output differs from cv2.remap - refusing to report a number")
        cps, done, secs = time_cpu(pool_frames, maps, args.cpu_budget, 4096)
        cpu_baseline = {"value": cps, "unit": "frame-sets/s", "cores": cv2.getNumThreads(), "kind": "port",
                        "sample": f"{done} frame sets of 8 x 1280x800 mono8 in {secs:.1f} s, cv2.remap INTER_LINEAR (OpenCV {cv2.__version__})"}

    if rank == 0:
        peak, peak_src = measured_peak()
        algo_bytes = B * PX_PER_SET * ALGO_BYTES_PER_PX
        achieved = algo_bytes / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": "frame_sets_per_sec",
            "value": value,
            "unit": "frame-sets/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "u8",
            "data": "synthetic",
            "config": workload_config(B, world),
            "mpix_per_sec": value * PX_PER_SET / 1e6,
            "e2e": {"value": e2e_value, "unit": "frame-sets/s", "h2d_bytes_per_step": Be * PX_PER_SET, "d2h_bytes_per_step": Be * PX_PER_SET,
                    "frame_sets_per_step": Be, "steps": e2e_steps, "chunk": args.chunk, "gpu_launches": launches_e2e,
                    "api": "ingest_host_submit / ingest_host_wait, two batches in flight (double-buffered pinned host frames)",
                    "blocking_call_value": e2e_blocking_value, "host_buffers_on_gpu_local_cpus": 0 if args.no_numa else host_cpus},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES_PER_FRAME_SET * B, "traffic_source": "ncu --set full, profiles/r01_rect_v4_ncu_raw.txt (scaled per frame set)",
                         "kernel": KERNEL_BY_VARIANT[plan["variant"]], "kernel_plan": plan, "algorithmic_bytes_per_launch": algo_bytes,
                         "peak_source": peak_src, "frac_of_8000_datasheet": achieved / 8000.0,
                         "note": "per-rank launch; duration = max-over-ranks ms_per_step (one launch per step)"},
            "cpu_baseline": cpu_baseline,
        }
        if extras:
            line["extras"] = extras
        print(json.dumps(line))
    ctx.close()
    if distributed:
        dist.destroy_process_group()


def run_extras(args, ctx, sources, maps, rank, world, distributed, barrier) -> dict:
    """Config 5 frame sets (4 mono rectify + 4 depth -> cloud) and, for N > 1, the cloud gather."""
    import torch
    import torch.distributed as dist

    from thor_slam_b200.camera.synthetic import make_depth
    from thor_slam_b200.ingest import formats as F
    from thor_slam_b200.ingest.calib import body_T_camera
    from thor_slam_b200.ingest.context import StreamSpec

    B = max(1, args.batch // 4)
    rng = np.random.default_rng(1337 + rank)
    specs = []
    keep = []
    clouds = torch.empty((N_CAMERAS, B, H, W, 3), dtype=torch.float32, device="cuda")
    for i, s in enumerate(sources):
        left = torch.from_numpy(np.stack([s._pool[b % 2][0] for b in range(B)])).cuda()
        out = torch.empty_like(left)
        depth = torch.from_numpy(np.stack([make_depth(rng, W, H) for _ in range(2)]).view(np.int16)).cuda().view(torch.uint16)
        depth = depth.repeat((B + 1) // 2, 1, 1)[:B].contiguous()
        mask = torch.empty((B, H, W), dtype=torch.uint8, device="cuda")
        count = torch.zeros((B,), dtype=torch.int32, device="cuda")
        intr = s.get_intrinsics()[0]
        ctx.upload_projection(2 * i, intr.matrix, body_T_camera(None, s.get_extrinsics()[0].to_4x4_matrix(), "rdf"), (W, H))
        specs.append(StreamSpec(F.KIND_RECTIFY, left, out, F.MONO8, F.MONO8, camera=2 * i))
        specs.append(StreamSpec(F.KIND_BACKPROJECT, depth, clouds[i], F.DEPTH16, F.XYZ32F, camera=2 * i, mask=mask, count=count))
        keep.append((left, out, depth, mask, count))
    stream = torch.cuda.current_stream()
    for _ in range(3):
        ctx.ingest(specs)
    barrier()
    steps = max(3, min(args.steps, 10))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        ctx.ingest(specs)
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    bytes_per_set = N_CAMERAS * W * H * (2 + 15)
    peak, _ = measured_peak()
    out = {"config5": {"frame_sets_per_step_per_rank": B, "ms_per_step": ms, "frame_sets_per_sec": world * B / (ms * 1e-3),
                       "hbm_gbs": B * bytes_per_set / (ms * 1e-3) / 1e9, "frac": B * bytes_per_set / (ms * 1e-3) / 1e9 / peak,
                       "algorithmic_bytes_per_frame_set": bytes_per_set}}
    if not distributed:
        out.update(run_config34(args, ctx, sources, keep, clouds, peak))
    if distributed:
        from thor_slam_b200.ingest.distributed import CloudGather, PeerCloudBuffer, RawDeviceBuffer

        # (a) compute, then ONE grouped ncclSend/ncclRecv gather of the dense clouds on rank 0
        gat = CloudGather(ctx, rank, world, root=0)
        nbytes = clouds.numel() * 4
        gathered = torch.empty((world, *clouds.shape), dtype=torch.float32, device="cuda") if rank == 0 else None
        for _ in range(2):
            gat.gather(clouds, gathered)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(steps):
            gat.gather(clouds, gathered)
        g1.record(stream)
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gms = float(tg.item()) / steps
        out["config5"]["gather_nccl"] = {"ms": gms, "bytes_into_root": nbytes * (world - 1),
                                         "root_ingress_gbs": nbytes * (world - 1) / (gms * 1e-3) / 1e9,
                                         "kind": "grouped ncclSend/ncclRecv of dense clouds after the kernel (includes the count all-gather)"}
        if rank == 0:
            out["config5"]["gather_nccl"]["matches_local"] = bool(torch.equal(gathered[0], clouds))
        # (b) gather fused into the kernel: xyz stores go straight into rank 0's buffer over NVLink
        peer = PeerCloudBuffer(ctx, rank, world, tuple(clouds.shape), root=0)
        mine = peer.slice_for(rank)
        fused_specs = []
        for i in range(len(specs)):
            sp = specs[i]
            if sp.kind == F.KIND_BACKPROJECT:
                cam_i = i // 2
                dst = RawDeviceBuffer(mine.ptr + cam_i * clouds[0].numel() * 4, tuple(clouds[0].shape), 4)
                sp = StreamSpec(F.KIND_BACKPROJECT, sp.src, dst, F.DEPTH16, F.XYZ32F, camera=sp.camera, mask=sp.mask, count=sp.count)
            fused_specs.append(sp)
        for _ in range(2):
            ctx.ingest(fused_specs)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(steps):
            ctx.ingest(fused_specs)
        f1.record(stream)
        barrier()
        tf = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        fms = float(tf.item()) / steps
        out["config5"]["fused_peer_store"] = {"ms_per_step": fms, "frame_sets_per_sec": world * B / (fms * 1e-3),
                                              "kind": "back-projection kernels store xyz into rank 0's buffer through CUDA-IPC peer mapping (gather fused into the kernel); barrier only"}
        if rank == 0:
            fused = peer.as_tensor()
            out["config5"]["fused_peer_store"]["matches_local"] = bool(torch.equal(fused[0], clouds))
        barrier()
        peer.close()
    return out


def run_config34(args, ctx, sources, keep5, clouds, peak) -> dict:
    """BASELINE configs 3 and 4 as whole frame sets, one ``ingest`` call per step (N = 1; kernels only, device-resident).

    Config 3: 4 x (1920x1080 BGR -> rgb8, 1280x800 depth -> body-frame cloud).  Config 4: 2 x "OAK-D Pro" (stereo mono
    1280x800 rectified, 1280x800 BGR -> rgb8, depth -> cloud) + 2 x "OAK-D LR" (stereo BGR 1920x1200 -> rgb8 rectified,
    1920x1200 BGR -> rgb8, depth -> cloud).  Algorithmic bytes per SURVEY section 8(d).
    """
    import torch

    from thor_slam_b200.camera.synthetic import SyntheticCameraConfig, SyntheticCameraSource
    from thor_slam_b200.ingest import formats as F
    from thor_slam_b200.ingest.calib import stereo_rectify_maps
    from thor_slam_b200.ingest.context import StreamSpec

    stream = torch.cuda.current_stream()
    steps = max(3, min(args.steps, 10))

    def timed(specs) -> float:
        for _ in range(3):
            ctx.ingest(specs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            ctx.ingest(specs)
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1) / steps

    def report(ms: float, n_sets: int, bytes_per_set: int) -> dict:
        gbs = n_sets * bytes_per_set / (ms * 1e-3) / 1e9
        return {"frame_sets_per_step": n_sets, "ms_per_step": ms, "frame_sets_per_sec": n_sets / (ms * 1e-3), "hbm_gbs": gbs,
                "frac": gbs / peak, "algorithmic_bytes_per_frame_set": bytes_per_set}

    def bgr(n, h, w):
        return torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")

    out = {}
    # ---- config 3 ----
    B3 = max(2, args.batch // 4)
    specs = []
    for i in range(N_CAMERAS):
        _, _, depth, mask, count = keep5[i]
        rgb_in = bgr(B3, 1080, 1920)
        specs.append(StreamSpec(F.KIND_CONVERT, rgb_in, torch.empty_like(rgb_in), F.BGR8, F.RGB8, width=1920, height=1080))
        specs.append(StreamSpec(F.KIND_BACKPROJECT, depth[:B3], clouds[i][:B3], F.DEPTH16, F.XYZ32F, camera=2 * i, mask=mask[:B3], count=count[:B3]))
    out["config3"] = report(timed(specs), B3, N_CAMERAS * (1920 * 1080 * 6 + W * H * 15))
    del specs
    # ---- config 4 ----
    B4 = max(2, args.batch // 8)
    lr = SyntheticCameraSource(SyntheticCameraConfig(name="lr0", resolution=(1920, 1200), pixel_format="bgr8", pool=1, enable_rgbd=False))
    lr_maps = stereo_rectify_maps(lr.get_intrinsics(), lr.get_extrinsics(), (1920, 1200))
    for k in range(2):
        for cam in range(2):
            ctx.upload_rectify_map(8 + 2 * k + cam, *lr_maps[cam], (1920, 1200))
    specs = []
    for i in range(2):  # OAK-D Pro: slots 2i, 2i+1 hold the mono stereo maps, slot 2i the depth projection
        left, _, depth, mask, count = keep5[i]
        for cam in range(2):
            specs.append(StreamSpec(F.KIND_RECTIFY, left[:B4], torch.empty_like(left[:B4]), F.MONO8, F.MONO8, camera=2 * i + cam))
        rgb_in = bgr(B4, H, W)
        specs.append(StreamSpec(F.KIND_CONVERT, rgb_in, torch.empty_like(rgb_in), F.BGR8, F.RGB8, width=W, height=H))
        specs.append(StreamSpec(F.KIND_BACKPROJECT, depth[:B4], clouds[i][:B4], F.DEPTH16, F.XYZ32F, camera=2 * i, mask=mask[:B4], count=count[:B4]))
    for k in range(2):  # OAK-D LR
        _, _, depth, mask, count = keep5[2 + k]
        for cam in range(2):
            col = bgr(B4, 1200, 1920)
            specs.append(StreamSpec(F.KIND_RECTIFY, col, torch.empty_like(col), F.BGR8, F.RGB8, camera=8 + 2 * k + cam))
        rgb_in = bgr(B4, 1200, 1920)
        specs.append(StreamSpec(F.KIND_CONVERT, rgb_in, torch.empty_like(rgb_in), F.BGR8, F.RGB8, width=1920, height=1200))
        specs.append(StreamSpec(F.KIND_BACKPROJECT, depth[:B4], clouds[2 + k][:B4], F.DEPTH16, F.XYZ32F, camera=2 * (2 + k), mask=mask[:B4], count=count[:B4]))
    pro = 2 * W * H * 2 + W * H * 6 + W * H * 15
    lrb = 2 * 1920 * 1200 * 6 + 1920 * 1200 * 6 + W * H * 15
    out["config4"] = report(timed(specs), B4, 2 * pro + 2 * lrb)
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frame sets per step (per rank)")
    ap.add_argument("--e2e-batch", type=int, default=64, dest="e2e_batch")
    ap.add_argument("--chunk", type=int, default=8, help="frame sets per H2D/compute/D2H pipeline stage")
    ap.add_argument("--cpu-budget", type=float, default=10.0, dest="cpu_budget", help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--ref-sets", type=int, default=8, dest="ref_sets", help="frame sets per step of the reference arm")
    ap.add_argument("--no-cpu", action="store_true", dest="no_cpu")
    ap.add_argument("--no-numa", action="store_true", dest="no_numa", help="do not place the pinned host buffers on the GPU's NUMA node")
    ap.add_argument("--extras", action="store_true", help="also time config 5 (rectify + back-projection) and the NCCL gather")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
