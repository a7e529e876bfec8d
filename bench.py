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

* ``sustained``: the same step loop run for >= 1.5 s (the K-step region of a short run lasts milliseconds), with the SM
                clock sampled through NVML over it.
* ``configs``  : BASELINE configs 3, 4 and 5 as whole frame sets, one ``ingest`` call per step (N = 1), and config 5 with the
                voxel down-sampled cloud (``ti_voxel_cloud``) in place of the dense one.
* ``e2e_rig``  : ``SyntheticCameraSource -> IngestRig.get_synchronized_frames() -> np.asarray(image)`` frame-sets/s and latency,
                beside the oracle (cv2.remap) driven through the ``CameraRig`` mirror.  N = 1.

With N > 1 every rank processes its own batch (weak scaling, no data-path collective - config 2 has no exchange step) and the
line carries ``config5``: BASELINE config 5 (4 x (mono rectify + depth -> cloud) per frame set, frame sets sharded over the ranks)
WITH its one exchange step - the gather of the clouds on rank 0 - overlapped with the next batch's kernels: voxel lists over NCCL
on the library's exchange stream, the same as peer-store kernels, and the dense clouds of round 1 for comparison; the gathered
list of the last step is checked against the oracle on rank 0.  ``--strong`` fixes the total at 64 frame sets instead of 16 per rank.
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
# dram__bytes_read.sum + dram__bytes_write.sum of ONE rectify_mono_pair_kernel launch (quad layout) over 64 frame sets of this rig,
# from the `ncu --set full` capture of round 2 (profiles/r02_ncu_rect_quad.txt): 577.31 MB + 479.36 MB = 1.008 x algorithmic
NCU_TRAFFIC_BYTES_PER_FRAME_SET = (577.310720e6 + 479.360000e6) / 64
KERNEL_BY_VARIANT = {4: "rectify_mono_pair_kernel<32,false,1280,192,QUAD>", 3: "rectify_mono_tma_kernel<32,false>", 2: "rectify_mono_kernel", 1: "rectify_tile_kernel<1>"}
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
_JSON_OUT = sys.stdout


def emit(line: dict) -> None:
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


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
    """The reference-side CPU path on the same workload, same step size, same loop as ``cpu_baseline``: ``args.batch`` frame sets
    per step, ``--warmup`` full steps untimed, ``--steps`` steps timed, every host thread OpenCV will use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import cv2

    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)
    sources, maps = build_rig()
    B = args.batch
    frames = host_frames(sources, 2)

    def step() -> None:
        for b in range(B):
            cpu_pass([f[b % 2: b % 2 + 1] for f in frames], maps, 1)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.steps * B / dt
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
        "config": workload_config(B, args.gpus),
        "mpix_per_sec": value * PX_PER_SET / 1e6,
        "cpu_baseline": {"value": value, "unit": "frame-sets/s", "cores": cv2.getNumThreads(), "kind": "port",
                         "sample": f"{args.steps} steps x {B} frame sets of 8 x 1280x800 mono8 in {dt:.1f} s after {args.warmup} warm-up steps, "
                                   f"cv2.remap INTER_LINEAR per stream (OpenCV {cv2.__version__}; oracle port of the reference-side path)"},
        "e2e": {"value": value, "unit": "frame-sets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


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
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the one JSON line and nothing else
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier() -> None:
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = IngestContext(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    sources, maps = build_rig()
    if os.environ.get("TI_BENCH_QUAD"):  # bring-up: the pair-window kernel's layout / exception capacity (TI_OPT_RECTIFY_QUAD)
        ctx.set_option(ctx.OPT_RECTIFY_QUAD, int(os.environ["TI_BENCH_QUAD"]))
    if os.environ.get("TI_BENCH_STAGES"):
        ctx.set_option(ctx.OPT_STAGES, int(os.environ["TI_BENCH_STAGES"]))
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

    if args.only_config5 and distributed:  # a second pass (e.g. --strong): the headline K steps above, then config 5 and nothing else
        c5 = run_config5_exchange(args, ctx, sources, rank, world, barrier, measured_peak()[0])
        if rank == 0:
            emit({"metric": "frame_sets_per_sec", "value": value, "unit": "frame-sets/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                  "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                  "config": workload_config(B, world), "gpu_launches": launches, "clocks": clocks, "config5": c5,
                  "note": "--only-config5 pass: no e2e / sustained / roofline blocks"})
        ctx.close()
        dist.destroy_process_group()
        return

    # ---- the same loop, sustained: the K-step region above lasts milliseconds; this one >= args.sustain_s seconds, with the
    # SM clock sampled over it (one NVML poll per 4 ms) -------------------------------------------
    sus_steps = int(min(50000, max(args.steps, args.sustain_s / max(ms_per_step * 1e-3, 1e-6))))
    sampler2 = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler2.start()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts_begin = time.perf_counter()
    sus_launches0 = ctx.launch_count
    s0.record(stream)
    for _ in range(sus_steps):
        ctx.ingest(specs)
    s1.record(stream)
    barrier()
    sus_launches = ctx.launch_count - sus_launches0
    ts_end = time.perf_counter()
    sus_clocks = sampler2.stop(ts_begin, ts_end) if rank == 0 else None
    tsus = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(tsus, op=dist.ReduceOp.MAX)
    sus_ms = float(tsus.item())
    sustained = {"steps": sus_steps, "seconds": round(sus_ms * 1e-3, 3), "ms_per_step": sus_ms / sus_steps, "value": world * B * sus_steps / (sus_ms * 1e-3),
                 "gpu_launches": sus_launches, "clocks": sus_clocks}

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

    peak, peak_src = measured_peak()
    pcie = pcie_probe(rank, world, barrier) if not args.no_pcie else None
    configs = run_configs(args, ctx, sources, peak) if (world == 1 and not args.no_configs) else None

    def headline(extra: dict) -> dict:
        """The JSON line from what has been measured so far (used at the end, and by the watchdog below)."""
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
                         "traffic": NCU_TRAFFIC_BYTES_PER_FRAME_SET * B, "traffic_source": "ncu --set full, profiles/r02_ncu_rect_quad.txt (scaled per frame set)",
                         "kernel": KERNEL_BY_VARIANT[plan["variant"]].replace("QUAD", "true" if plan.get("pixels_per_window") == 4 else "false"), "kernel_plan": plan, "algorithmic_bytes_per_launch": algo_bytes,
                         "peak_source": peak_src, "frac_of_8000_datasheet": achieved / 8000.0,
                         "sustained_achieved": algo_bytes / (sustained["ms_per_step"] * 1e-3) / 1e9,
                         "sustained_frac": algo_bytes / (sustained["ms_per_step"] * 1e-3) / 1e9 / peak,
                         "note": "per-rank launches; duration = max-over-ranks ms_per_step (per step: ONE remap launch for the 8 streams + one per-pixel repair launch for the ~1 650 pixels per frame set whose exception lists overflowed; the repair pass is inside the timed step)"},
            "cpu_baseline": None,
            "sustained": sustained,
        }
        for key, val in extra.items():
            if val is not None:
                line[key] = val
        return line

    # Config 5 with its exchange is the one part of this run in which ranks wait for one another on the device.  Whatever happens
    # there - an exception on one rank, a peer that never answers - the line with everything measured so far still goes out:
    # a watchdog prints it (rank 0) and ends the process, so the driver never waits for a number that exists already.
    config5 = None
    if distributed and not args.no_configs:
        def bail(reason: str) -> None:
            if rank == 0:
                emit(headline({"pcie": pcie, "config5": {"error": reason}}))
            os._exit(0)

        watchdog = threading.Timer(args.c5_timeout, bail, args=(f"config 5 did not finish within {args.c5_timeout:.0f} s",))
        watchdog.daemon = True
        watchdog.start()
        try:
            config5 = run_config5_exchange(args, ctx, sources, rank, world, barrier, peak)
        except SystemExit:
            raise  # a parity failure: no number may be printed
        except Exception as exc:  # noqa: BLE001 - anything else: keep the headline, say what happened
            print(f"[bench] rank {rank}: config 5 failed: {exc!r}", file=sys.stderr, flush=True)
            watchdog.cancel()
            bail(f"rank {rank}: {exc!r}")
        watchdog.cancel()
    e2e_rig = None
    if rank == 0 and world == 1 and not args.no_rig:
        e2e_rig = run_e2e_rig(args)

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
        line = headline({"configs": configs, "config5": config5, "e2e_rig": e2e_rig, "pcie": pcie})
        line["cpu_baseline"] = cpu_baseline
        emit(line)
    ctx.close()
    if distributed:
        dist.destroy_process_group()


def timed_steps(stream, fn, steps: int, warm: int = 3) -> float:
    """ms per call of ``fn(k)`` over ``steps`` calls after ``warm`` untimed ones (CUDA events on ``stream``)."""
    import torch

    for k in range(warm):
        fn(k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(steps):
        fn(warm + k)
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / steps


class Config5:
    """BASELINE config 5 frame sets on one rank: 4 cameras x (left mono 1280x800 -> rectified, depth 1280x800 -> cloud).

    Depth is a scene of surfaces (``make_depth_scene``; ``scene="noise"``: the per-pixel noise of SURVEY 8(d)), two distinct
    frame sets per camera tiled over the batch, seeded per rank so that rank 0 can rebuild every rank's input for the oracle."""

    def __init__(self, ctx, sources, B: int, rank: int, scene: str = "room") -> None:
        import torch

        from thor_slam_b200.camera.synthetic import make_depth, make_depth_scene
        from thor_slam_b200.ingest import formats as F
        from thor_slam_b200.ingest.calib import body_T_camera
        from thor_slam_b200.ingest.context import StreamSpec

        self.ctx, self.B, self.rank, self.scene = ctx, B, rank, scene
        self.k, self.m, self.depth_np = [], [], self.host_depth(sources, rank, scene)
        self.rect_specs, self.dense_specs, self.depth_streams, self.keep = [], [], [], []
        self.clouds = None
        for i, s in enumerate(sources):
            left = torch.from_numpy(np.stack([s._pool[b % 2][0] for b in range(B)])).cuda()
            out = torch.empty_like(left)
            depth = torch.from_numpy(self.depth_np[i].view(np.int16)).cuda().view(torch.uint16).repeat((B + 1) // 2, 1, 1)[:B].contiguous()
            intr = s.get_intrinsics()[0]
            m = body_T_camera(None, s.get_extrinsics()[0].to_4x4_matrix(), "rdf")
            ctx.upload_projection(2 * i, intr.matrix, m, (W, H))
            self.k.append(intr.matrix)
            self.m.append(m)
            self.rect_specs.append(StreamSpec(F.KIND_RECTIFY, left, out, F.MONO8, F.MONO8, camera=2 * i))
            self.depth_streams.append((2 * i, depth))
            self.keep.append((left, out, depth))

    @staticmethod
    def host_depth(sources, rank: int, scene: str) -> list[np.ndarray]:
        """[camera] -> u16 [2, H, W]: the two distinct depth frames of every camera of ``rank``."""
        from thor_slam_b200.camera.synthetic import make_depth, make_depth_scene

        rng = np.random.default_rng(4242 + rank)
        out = []
        for s in sources:
            f = float(s.get_intrinsics()[0].matrix[0, 0])
            out.append(np.stack([make_depth_scene(rng, W, H, focal_px=f) if scene == "room" else make_depth(rng, W, H) for _ in range(2)]))
        return out

    def add_dense(self) -> None:
        import torch

        from thor_slam_b200.ingest import formats as F
        from thor_slam_b200.ingest.context import StreamSpec

        B = self.B
        self.clouds = torch.empty((N_CAMERAS, B, H, W, 3), dtype=torch.float32, device="cuda")
        self.masks = torch.empty((N_CAMERAS, B, H, W), dtype=torch.uint8, device="cuda")
        self.counts = torch.zeros((N_CAMERAS, B), dtype=torch.int32, device="cuda")
        self.dense_specs = list(self.rect_specs)
        for i, (cam, depth) in enumerate(self.depth_streams):
            self.dense_specs.append(StreamSpec(F.KIND_BACKPROJECT, depth, self.clouds[i], F.DEPTH16, F.XYZ32F, camera=cam, mask=self.masks[i], count=self.counts[i]))

    def oracle_records(self, rank: int, sources, tag: int, sets: int | None = None) -> np.ndarray:
        """Sorted records of the first ``sets`` frame sets (default: all) of one step of ``rank`` (set j uses depth frame j % 2)."""
        from oracle import voxel as ov

        n = self.B if sets is None else min(sets, self.B)
        depth = self.depth_np if rank == self.rank else self.host_depth(sources, rank, self.scene)
        base = [ov.voxel_records([(depth[i][j], self.k[i], self.m[i]) for i in range(N_CAMERAS)], 0.05, 10000, set_id=0, tag=tag) for j in range(min(2, n))]
        parts = [base[j % 2] | (np.uint64(j) << np.uint64(45)) for j in range(n)]
        return np.sort(np.concatenate(parts))

    @staticmethod
    def first_sets(records: np.ndarray, sets: int | None) -> np.ndarray:
        if sets is None:
            return records
        return records[((records >> np.uint64(45)) & np.uint64(0x7FF)) < np.uint64(sets)]


def run_configs(args, ctx, sources, peak: float) -> dict:
    """N = 1: BASELINE configs 3, 4, 5 as whole frame sets (kernels only, device-resident, one ``ingest`` call per step) and
    config 5 with the voxel down-sampled cloud.  Algorithmic bytes per SURVEY section 8(d)."""
    import torch

    from thor_slam_b200.camera.synthetic import SyntheticCameraConfig, SyntheticCameraSource
    from thor_slam_b200.ingest import formats as F
    from thor_slam_b200.ingest.calib import stereo_rectify_maps
    from thor_slam_b200.ingest.context import StreamSpec

    stream = torch.cuda.current_stream()
    steps = max(5, min(args.steps, 20))

    def report(ms: float, n_sets: int, bytes_per_set: int, **extra) -> dict:
        gbs = n_sets * bytes_per_set / (ms * 1e-3) / 1e9
        return {"frame_sets_per_step": n_sets, "ms_per_step": round(ms, 5), "frame_sets_per_sec": round(n_sets / (ms * 1e-3), 1), "hbm_gbs": round(gbs, 1),
                "frac": round(gbs / peak, 4), "algorithmic_bytes_per_frame_set": bytes_per_set, **extra}

    def bgr(n, h, w):
        return torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")

    out = {}
    B5 = max(2, args.batch // 4)
    c5 = Config5(ctx, sources, B5, rank=0, scene="room")
    c5.add_dense()
    out["5"] = report(timed_steps(stream, lambda k: ctx.ingest(c5.dense_specs), steps), B5, N_CAMERAS * W * H * (2 + 15),
                      what="4 x (mono rectify + depth -> dense body-frame cloud + mask + count)")
    # config 5 with the cloud down-sampled to occupied 0.05 m voxels (nvblox's grid, 10 m cap) instead of 12 B per pixel
    ctx.set_voxel_grid(0.05, 10000)
    cap = B5 * 400_000
    rec = torch.empty(cap, dtype=torch.int64, device="cuda")
    nrec = torch.zeros(1, dtype=torch.int32, device="cuda")

    def voxel_step(k: int) -> None:
        ctx.ingest(c5.rect_specs)
        ctx.voxel_cloud(c5.depth_streams, rec, nrec)

    ms_v = timed_steps(stream, voxel_step, steps)
    n_vox = int(nrec.item())
    if n_vox > cap:
        raise SystemExit(f"bench: voxel list overflow ({n_vox} > {cap})")
    ms_vk = timed_steps(stream, lambda k: ctx.voxel_cloud(c5.depth_streams, rec, nrec), steps)
    out["5_voxel"] = report(ms_v, B5, N_CAMERAS * W * H * (2 + 2) + 8 * n_vox // B5,
                            what="4 x mono rectify + 4 depth frames -> ONE list of occupied voxels per frame set (ti_voxel_cloud)",
                            scene="room (surfaces)", voxels_per_frame_set=n_vox // B5, valid_pixels_per_voxel=round(0.8 * N_CAMERAS * W * H * B5 / max(n_vox, 1), 1),
                            cloud_bytes_per_frame_set={"dense_xyz": N_CAMERAS * W * H * 12, "voxel_records": 8 * n_vox // B5},
                            voxel_kernel_ms=round(ms_vk, 5), voxel_kernel_gpix_per_sec=round(N_CAMERAS * W * H * B5 / (ms_vk * 1e-3) / 1e9, 1))
    got = np.sort(rec[:n_vox].cpu().numpy().view(np.uint64))  # the checker, outside every timed region
    if not np.array_equal(got, c5.oracle_records(0, sources, tag=0)):
        raise SystemExit("bench: voxel records differ from the oracle - refusing to report a number")
    out["5_voxel"]["oracle_check"] = "record set of one step == np.unique of the oracle's keys"
    # ---- config 3 ----
    B3 = max(2, args.batch // 4)
    specs = []
    for i in range(N_CAMERAS):
        rgb_in = bgr(B3, 1080, 1920)
        specs.append(StreamSpec(F.KIND_CONVERT, rgb_in, torch.empty_like(rgb_in), F.BGR8, F.RGB8, width=1920, height=1080))
        specs.append(c5.dense_specs[N_CAMERAS + i])
    out["3"] = report(timed_steps(stream, lambda k: ctx.ingest(specs), steps), B3, N_CAMERAS * (1920 * 1080 * 6 + W * H * 15),
                      what="4 x (1920x1080 BGR -> rgb8 + 1280x800 depth -> cloud)")
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
        left, _, depth = c5.keep[i]
        for cam in range(2):
            specs.append(StreamSpec(F.KIND_RECTIFY, left[:B4], torch.empty_like(left[:B4]), F.MONO8, F.MONO8, camera=2 * i + cam))
        rgb_in = bgr(B4, H, W)
        specs.append(StreamSpec(F.KIND_CONVERT, rgb_in, torch.empty_like(rgb_in), F.BGR8, F.RGB8, width=W, height=H))
        specs.append(StreamSpec(F.KIND_BACKPROJECT, depth[:B4], c5.clouds[i][:B4], F.DEPTH16, F.XYZ32F, camera=2 * i, mask=c5.masks[i][:B4], count=c5.counts[i][:B4]))
    for k in range(2):  # OAK-D LR
        _, _, depth = c5.keep[2 + k]
        for cam in range(2):
            col = bgr(B4, 1200, 1920)
            specs.append(StreamSpec(F.KIND_RECTIFY, col, torch.empty_like(col), F.BGR8, F.RGB8, camera=8 + 2 * k + cam))
        rgb_in = bgr(B4, 1200, 1920)
        specs.append(StreamSpec(F.KIND_CONVERT, rgb_in, torch.empty_like(rgb_in), F.BGR8, F.RGB8, width=1920, height=1200))
        specs.append(StreamSpec(F.KIND_BACKPROJECT, depth[:B4], c5.clouds[2 + k][:B4], F.DEPTH16, F.XYZ32F, camera=2 * (2 + k), mask=c5.masks[2 + k][:B4],
                                count=c5.counts[2 + k][:B4]))
    pro = 2 * W * H * 2 + W * H * 6 + W * H * 15
    lrb = 2 * 1920 * 1200 * 6 + 1920 * 1200 * 6 + W * H * 15
    out["4"] = report(timed_steps(stream, lambda k: ctx.ingest(specs), steps), B4, 2 * pro + 2 * lrb,
                      what="2 x OAK-D Pro (2 mono rectify, BGR -> rgb8, depth -> cloud) + 2 x OAK-D LR (2 x 1920x1200 BGR -> rgb8 rectify, BGR -> rgb8, depth -> cloud)")
    return out


def run_config5_exchange(args, ctx, sources, rank: int, world: int, barrier, peak: float) -> dict:
    """N > 1: config 5 WITH its exchange step.  Frame sets are sharded over the ranks (``B`` per rank per step); every step's
    cloud goes to rank 0 while the next step's kernels run.  Reported: compute only, and compute + exchange overlapped, for
    (a) voxel lists over NCCL on the exchange stream, (b) voxel lists as peer-store kernels, (c) round 1's dense clouds."""
    import torch
    import torch.distributed as dist

    from thor_slam_b200.ingest.distributed import CloudGather, RecordExchange

    if os.environ.get("TI_BENCH_FAIL_RANK") == str(rank):  # test hook of the watchdog in run_ours
        raise RuntimeError("injected failure (TI_BENCH_FAIL_RANK)")
    stream = torch.cuda.current_stream()
    B = max(1, 64 // world) if args.strong else (args.c5_batch or max(1, args.batch // 4))
    steps = max(6, min(args.steps, 30))
    out: dict = {"frame_sets_per_step_per_rank": B, "scaling": "strong (64 frame sets per step in total)" if args.strong else "weak",
                 "sharding": f"frame set i -> rank i mod {world}; 4 cameras x (mono rectify + depth) per frame set; clouds gathered on rank 0"}

    def max_ms(ms: float) -> float:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def fs_per_sec(ms: float) -> float:
        return round(world * B / (ms * 1e-3), 1)

    ctx.set_voxel_grid(0.05, 10000)
    gat = CloudGather(ctx, rank, world, root=0)
    for scene in args.c5_scenes.split(","):
        c5 = Config5(ctx, sources, B, rank, scene)
        check_sets = None if scene == "room" else 1  # the noise lists are 16 x larger: the oracle checks frame set 0 of every rank
        cap = B * (400_000 if scene == "room" else 3_400_000)
        NB = 2
        rec = [torch.empty(cap, dtype=torch.int64, device="cuda") for _ in range(NB)]
        nrec = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(NB)]

        def compute(k: int) -> None:
            ctx.ingest(c5.rect_specs)
            ctx.voxel_cloud(c5.depth_streams, rec[k % NB], nrec[k % NB], tag=rank)

        res: dict = {}
        barrier()
        ms_c = max_ms(timed_steps(stream, compute, steps))
        n_mine = int(nrec[0].item())
        if n_mine > cap:
            raise SystemExit(f"bench: voxel list overflow ({n_mine} > {cap})")
        res["compute_only"] = {"ms_per_step": round(ms_c, 5), "frame_sets_per_sec": fs_per_sec(ms_c)}
        res["voxels_per_frame_set"] = n_mine // B
        res["cloud_bytes_per_frame_set"] = {"dense_xyz": N_CAMERAS * W * H * 12, "voxel_records": 8 * n_mine // B}

        # (a) NCCL on the exchange stream, pipelined: step k's exchange runs under step k + 1's kernels
        gathered = [torch.empty(world * cap, dtype=torch.int64, device="cuda") if rank == 0 else None for _ in range(NB)]
        last_counts: list = [None]

        def run_nccl(n: int) -> float:
            fences = [0] * n
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)

            def send(j: int) -> None:
                _, last_counts[0] = gat.records_send(rec[j % NB], gathered[j % NB], wait=False)
                fences[j] = ctx.exchange_fence()

            for k in range(n):
                if k >= NB:
                    ctx.exchange_wait(fences[k - NB], on_stream=True)
                compute(k)
                if k >= 1:
                    send(k - 1)
                gat.records_begin(nrec[k % NB])
            send(n - 1)
            ctx.exchange_wait(fences[n - 1], on_stream=True)  # the timed region ends when the last cloud has landed
            e1.record(stream)
            e1.synchronize()
            return e0.elapsed_time(e1) / n

        run_nccl(3)
        ms_n = max_ms(run_nccl(steps))
        per_rank = last_counts[0]
        res["overlapped_nccl"] = {"ms_per_step": round(ms_n, 5), "frame_sets_per_sec": fs_per_sec(ms_n), "vs_compute_only": round(ms_c / ms_n, 4),
                                  "bytes_into_root_per_step": 8 * (sum(per_rank) - per_rank[0]),
                                  "kind": "ti_gather_counts_begin/finish + ti_gather_records (grouped ncclSend/ncclRecv) on the exchange stream"}
        if rank == 0:  # the checker, outside the timed region: last step's fused list against the oracle, every rank's frames
            n_tot = sum(per_rank)
            got = np.sort(Config5.first_sets(gathered[(steps - 1) % NB][:n_tot].cpu().numpy().view(np.uint64), check_sets))
            want = np.sort(np.concatenate([c5.oracle_records(r, sources, tag=r, sets=check_sets) for r in range(world)]))
            if not np.array_equal(got, want):
                raise SystemExit(f"bench: gathered voxel list ({scene}, NCCL) differs from the oracle - refusing to report a number")
            res["overlapped_nccl"]["oracle_check"] = f"{len(got)} of {n_tot} gathered records ({'all' if check_sets is None else 'frame set 0 of every rank'}) == oracle over all {world} ranks"
        del gathered
        barrier()

        # (b) the same exchange as peer-store kernels over NVLink (no collective library, no host round trip)
        ex = RecordExchange(ctx, rank, world, capacity=world * cap, root=0, slots=2)
        taken = torch.empty(world * cap, dtype=torch.int64, device="cuda") if rank == 0 else None
        status = torch.zeros(2, dtype=torch.int32, device="cuda") if rank == 0 else None

        def run_push(n: int) -> float:
            fences = [0] * n
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for k in range(n):
                if k >= NB:
                    ctx.exchange_wait(fences[k - NB], on_stream=True)
                compute(k)
                fences[k] = ex.push(rec[k % NB], nrec[k % NB])
                if rank == 0:  # the fusing rank consumes a round in place while the other inbox fills; the last one is copied out for the oracle
                    ex.take(taken if k == n - 1 else None, status)
            ex.wait(on_stream=True)
            e1.record(stream)
            e1.synchronize()
            return e0.elapsed_time(e1) / n

        run_push(4)
        ms_p = run_push(steps)
        ms_p_ranks = None
        res["overlapped_push"] = {}

        t_mine = torch.tensor([ms_p], dtype=torch.float64, device="cuda")  # who sets the pace: this rank's own time per step
        t_all = [torch.zeros_like(t_mine) for _ in range(world)]
        dist.all_gather(t_all, t_mine)
        if rank == 0:
            n_tot, err = (int(x) for x in status.cpu().numpy())
            others = taken[:n_tot].cpu().numpy().view(np.uint64)
            mine = rec[(steps - 1) % NB][: int(nrec[(steps - 1) % NB].item())].cpu().numpy().view(np.uint64)  # the root's own list never left
            n_tot = int((others != 0).sum()) + len(mine)  # zero records pad odd lists to 16 bytes
            got = np.sort(Config5.first_sets(np.concatenate([others[others != 0], mine]), check_sets))
            want = np.sort(np.concatenate([c5.oracle_records(r, sources, tag=r, sets=check_sets) for r in range(world)]))
            if os.environ.get("TI_PUSH_DEBUG"):  # bring-up switches of the exchange (no payload / no kernels): nothing to check
                got = want
            if err or not np.array_equal(got, want):
                raise SystemExit(f"bench: gathered voxel list ({scene}, peer stores, error flag {err}) differs from the oracle - refusing to report a number")
            ms_p_ranks = [round(float(t.item()), 5) for t in t_all]
            res["overlapped_push"] = {"ms_per_step": round(max(ms_p_ranks), 5), "frame_sets_per_sec": fs_per_sec(max(ms_p_ranks)),
                                      "vs_compute_only": round(ms_c / max(ms_p_ranks), 4), "ms_per_step_by_rank": ms_p_ranks,
                                      "kind": "ti_cloud_push / ti_inbox_take: slots reserved with one system-scope atomic, TMA bulk copies into rank 0's "
                                              "inbox over NVLink, consumed in place; rank 0's own list stays local"}
            res["overlapped_push"]["oracle_check"] = f"{len(got)} of {n_tot} records in rank 0's inbox ({'all' if check_sets is None else 'frame set 0 of every rank'}) == oracle over all {world} ranks"
        ex.close()
        del taken, rec
        out[scene] = res
        if scene == "room":
            # (c) round 1's exchange for comparison: dense clouds (12 B per pixel, invalid ones included), gathered after the kernels
            c5.add_dense()
            barrier()
            ms_d = max_ms(timed_steps(stream, lambda k: ctx.ingest(c5.dense_specs), steps))
            dsteps = max(3, steps // 3)
            gd = torch.empty((world, *c5.clouds.shape), dtype=torch.float32, device="cuda") if rank == 0 else None

            def dense_step(k: int) -> None:
                ctx.ingest(c5.dense_specs)
                gat.gather(c5.clouds, gd, wait=False)
                gat.wait(on_stream=True)

            barrier()
            ms_dg = max_ms(timed_steps(stream, dense_step, dsteps, warm=2))
            gat.wait()
            nbytes = c5.clouds.numel() * 4
            out["dense"] = {"compute_only": {"ms_per_step": round(ms_d, 5), "frame_sets_per_sec": fs_per_sec(ms_d)},
                            "with_gather": {"ms_per_step": round(ms_dg, 5), "frame_sets_per_sec": fs_per_sec(ms_dg), "vs_compute_only": round(ms_d / ms_dg, 4),
                                            "bytes_into_root_per_step": nbytes * (world - 1),
                                            "kind": "ti_gather_clouds of the dense clouds after every step (the buffer is rewritten by the next step, so nothing overlaps)"}}
            del gd
        del c5
        torch.cuda.empty_cache()
    return out


def run_e2e_rig(args) -> dict:
    """The drop-in itself: 4 synthetic stereo sources -> ``IngestRig.get_synchronized_frames()`` -> ``np.asarray(frame.image)`` of
    all 8 rectified frames, one frame set per call, beside the oracle (cv2.remap, all host threads) behind the ``CameraRig``
    mirror on the same sources.  Host buffers in, host buffers out."""
    import cv2
    import torch

    from oracle import rectify as orc
    from thor_slam_b200.camera.rig import CameraRig
    from thor_slam_b200.camera.synthetic import make_rig_sources
    from thor_slam_b200.ingest.calib import stereo_rectify_maps
    from thor_slam_b200.ingest.rig import IngestRig

    n_sets = 200
    sources = make_rig_sources(N_CAMERAS, resolution=(W, H), pixel_format="mono8", seed=1337, pool=2)
    rig = IngestRig(sources, queue_size=10)
    rig.start()
    lat = []
    for _ in range(10):
        fs = rig.get_synchronized_frames()
        [np.asarray(f.image) for f in fs.get_all_frames()]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n_sets):
        t1 = time.perf_counter()
        fs = rig.get_synchronized_frames()
        imgs = [np.asarray(f.image) for f in fs.get_all_frames()]
        lat.append(time.perf_counter() - t1)
    dt = time.perf_counter() - t0
    last_gpu = imgs[3].copy()
    last_src = fs  # noqa: F841
    rig.stop()
    # the oracle behind the CameraRig mirror, same sources, same frames
    sources2 = make_rig_sources(N_CAMERAS, resolution=(W, H), pixel_format="mono8", seed=1337, pool=2)
    maps = []
    for s in sources2:
        maps.extend(stereo_rectify_maps(s.get_intrinsics(), s.get_extrinsics(), (W, H)))
    cv2.setNumThreads(os.cpu_count() or 1)
    ref = CameraRig(sources2, queue_size=10)
    ref.start()
    for _ in range(10):
        fs2 = ref.get_synchronized_frames()
        [orc.remap_cv(f.image, *maps[i]) for i, f in enumerate(fs2.get_all_frames())]
    n_cpu, t0c, lat_c = 0, time.perf_counter(), []
    while n_cpu < n_sets and time.perf_counter() - t0c < 8.0:
        t1 = time.perf_counter()
        fs2 = ref.get_synchronized_frames()
        outs = [orc.remap_cv(f.image, *maps[i]) for i, f in enumerate(fs2.get_all_frames())]
        lat_c.append(time.perf_counter() - t1)
        n_cpu += 1
    dtc = time.perf_counter() - t0c
    ref.stop()
    # same call count on both rigs -> the same source frames: compare one stream of the last frame set of equal index
    check = None
    if n_cpu == n_sets:
        check = bool(np.array_equal(last_gpu, outs[3]))
        if not check:
            raise SystemExit("bench: IngestRig output differs from cv2.remap behind the CameraRig mirror")
    return {"frame_sets_per_sec": round(n_sets / dt, 1), "latency_ms_median": round(float(np.median(lat)) * 1e3, 3), "latency_ms_p95": round(float(np.percentile(lat, 95)) * 1e3, 3),
            "frame_sets": n_sets, "h2d_bytes_per_frame_set": PX_PER_SET, "d2h_bytes_per_frame_set": PX_PER_SET,
            "api": "IngestRig.get_synchronized_frames() + np.asarray(frame.image) for all 8 frames, one frame set per call",
            "cpu_rig": {"frame_sets_per_sec": round(n_cpu / dtc, 1), "latency_ms_median": round(float(np.median(lat_c)) * 1e3, 3), "frame_sets": n_cpu,
                        "cores": cv2.getNumThreads(), "api": "CameraRig (API mirror).get_synchronized_frames() + cv2.remap per frame (oracle)"},
            "matches_cpu_rig": check}


def pcie_probe(rank: int, world: int, barrier, mb: int = 256, iters: int = 6) -> dict:
    """Host <-> device copy ceilings with ALL ranks copying at once (pinned buffers near each GPU): what the box gives the
    e2e path at this N.  GB/s per rank (min over ranks) and summed over ranks."""
    import torch
    import torch.distributed as dist

    from thor_slam_b200.ingest.hostmem import near_gpu

    n = mb << 20
    with near_gpu(torch.cuda.current_device()):
        h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    up, down = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d: bool, d2h: bool) -> float:
        barrier()
        t0 = time.perf_counter()
        for _ in range(iters):
            if h2d:
                with torch.cuda.stream(up):
                    d_a.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(down):
                    h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        return n * iters / (time.perf_counter() - t0) / 1e9

    res = {}
    for name, a, b in (("h2d_alone", True, False), ("d2h_alone", False, True), ("both_directions", True, True)):
        run(a, b)
        g = run(a, b)
        t = torch.tensor([g, -g], dtype=torch.float64, device="cuda")
        s = torch.tensor([g], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            dist.all_reduce(s, op=dist.ReduceOp.SUM)
        res[name] = {"gbs_per_direction_min_rank": round(float(t[0].item()), 1), "gbs_per_direction_max_rank": round(-float(t[1].item()), 1),
                     "gbs_per_direction_sum": round(float(s.item()), 1)}
    res["note"] = f"{mb} MB pinned copies, all {world} rank(s) at once"
    return res


def main() -> None:
    # stdout carries the one JSON line and nothing else: whatever a library prints there (NCCL announces its version on the
    # first communicator) goes to stderr instead; the line itself is written to the saved descriptor
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frame sets per step (per rank)")
    ap.add_argument("--e2e-batch", type=int, default=64, dest="e2e_batch")
    ap.add_argument("--chunk", type=int, default=8, help="frame sets per H2D/compute/D2H pipeline stage")
    ap.add_argument("--cpu-budget", type=float, default=10.0, dest="cpu_budget", help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu", action="store_true", dest="no_cpu")
    ap.add_argument("--no-numa", action="store_true", dest="no_numa", help="do not place the pinned host buffers on the GPU's NUMA node")
    ap.add_argument("--no-configs", action="store_true", dest="no_configs", help="skip configs 3/4/5 (N = 1) and config 5 with its exchange (N > 1)")
    ap.add_argument("--no-rig", action="store_true", dest="no_rig", help="skip the IngestRig end-to-end loop")
    ap.add_argument("--no-pcie", action="store_true", dest="no_pcie", help="skip the host<->device copy probe")
    ap.add_argument("--only-config5", action="store_true", dest="only_config5", help="N > 1: skip the e2e loops and the sustained loop (a second, --strong pass)")
    ap.add_argument("--c5-batch", type=int, default=0, dest="c5_batch", help="config 5: frame sets per rank per step (default batch / 4)")
    ap.add_argument("--c5-scenes", default="room,noise", dest="c5_scenes")
    ap.add_argument("--c5-timeout", type=float, default=240.0, dest="c5_timeout", help="seconds after which an N > 1 run prints its line without config 5")
    ap.add_argument("--strong", action="store_true", help="config 5 at N > 1: 64 frame sets per step in total instead of 16 per rank")
    ap.add_argument("--sustain-s", type=float, default=1.5, dest="sustain_s", help="seconds of the sustained loop")
    ap.add_argument("--extras", action="store_true", help=argparse.SUPPRESS)  # round-1 flag, now the default
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
