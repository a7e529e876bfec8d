"""Host-side placement for the host-buffer pipeline (``ti_ingest_host``).

Pinned frame buffers are placed on the NUMA node of the thread that allocates them.  With one process per GPU
 every rank should allocate on the node its GPU hangs off, otherwise its uploads and downloads
cross the socket interconnect and contend with the other ranks'.  ``bind_to_gpu`` pins the calling process to the
CPUs NVML reports as local to the GPU; call it before allocating pinned memory.
"""

from __future__ import annotations

import os


def gpu_local_cpus(device_index: int) -> set[int]:
    """CPUs NVML reports as closest to the GPU (physical index as NVML counts it); empty when NVML cannot say."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_cpus = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpus + 63) // 64)
    except Exception:
        return set()
    cpus: set[int] = set()
    for w, word in enumerate(words):
        for b in range(64):
            if (int(word) >> b) & 1:
                cpus.add(64 * w + b)
    return cpus


def nvml_index(cuda_index: int) -> int:
    """NVML index of a CUDA device ordinal (they differ under CUDA_VISIBLE_DEVICES)."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "").strip()
    if vis:
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if cuda_index < len(ids) and ids[cuda_index].isdigit():
            return int(ids[cuda_index])
    return cuda_index


def bind_to_gpu(cuda_index: int) -> list[int]:
    """Restrict this process to the CPUs local to the GPU.  Returns the CPUs now allowed (unchanged set if NVML has
    no answer or the local CPUs are outside the process's cpuset)."""
    allowed = os.sched_getaffinity(0)
    local = gpu_local_cpus(nvml_index(cuda_index)) & allowed
    if local and local != allowed:
        os.sched_setaffinity(0, local)
        allowed = local
    return sorted(allowed)


class near_gpu:
    """``with near_gpu(i): buf = torch.empty(...).pin_memory()`` - bind for the allocation only, then restore the
    previous CPU set (the pages stay where they were first placed; the copies are DMA and do not care where the
    thread runs afterwards)."""

    def __init__(self, cuda_index: int):
        self.cuda_index = cuda_index
        self.before: set[int] | None = None
        self.cpus: list[int] = []

    def __enter__(self) -> "near_gpu":
        self.before = os.sched_getaffinity(0)
        self.cpus = bind_to_gpu(self.cuda_index)
        return self

    def __exit__(self, *exc) -> None:
        if self.before is not None:
            os.sched_setaffinity(0, self.before)
