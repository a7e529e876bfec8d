"""Multi-GPU side of the ingest path: one process per GPU, frame sets sharded, ONE cloud gather.

The path shards with no data-path collective: every stream of every frame set is independent (the
LUTs and body transforms are per-camera constants replicated on every GPU).  The only exchange step
is the gather of the per-GPU body-frame point clouds on the fusing rank:

* ``shard_frame_sets``     - frame set ``i`` -> rank ``i % world`` (perfect balance, any batch size);
* ``exchange_counts``      - tiny all-gather of per-rank byte counts over ``torch.distributed`` (plumbing);
* ``CloudGather.gather``   - grouped ``ncclSend``/``ncclRecv`` inside ``libthoringest.so`` on the library's own
                             exchange stream (``ti_gather_clouds``: it starts behind the kernels enqueued so far
                             and overlaps the ones enqueued next), or ``torch.distributed`` send/recv when the
                             process group is ``gloo`` (CPU tests of the host logic);
* ``CloudGather.gather_records`` - the same for the variable-length voxel lists of ``ti_voxel_cloud``: the device
                             counts are all-gathered on the exchange stream (``ti_gather_counts``), no torch collective;
* ``RecordExchange``       - the exchange as our own kernels: producers append their lists to an inbox in the root's
                             HBM with peer stores over NVLink, reserving slots with one system-scope atomic
                             (``ti_cloud_push`` / ``ti_inbox_take``); no collective library and no host round trip;
* ``PeerCloudBuffer``      - the gather FUSED into the producing kernel: the root allocates the fused
                             cloud buffer, exports a CUDA-IPC handle, every rank maps it and hands its slice
                             to ``ti_backproject`` as the xyz destination, so the kernel's stores go straight
                             over NVLink into the root's HBM; only a barrier remains.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Sequence

import torch
import torch.distributed as dist

from thor_slam_b200.ingest.context import IngestContext


def shard_frame_sets(n_sets: int, rank: int, world: int) -> list[int]:
    """Indices of the frame sets rank ``rank`` of ``world`` processes (round robin)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_sets, world))


def gather_layout(bytes_per_rank: Sequence[int]) -> list[int]:
    """Byte offset of every rank's slice in the gathered buffer (rank order, back to back)."""
    offs, acc = [], 0
    for b in bytes_per_rank:
        if b < 0:
            raise ValueError("negative slice size")
        offs.append(acc)
        acc += int(b)
    return offs


def exchange_counts(local_bytes: int, group: Any = None) -> list[int]:
    """All ranks learn every rank's slice size (needed before a variable-length gather)."""
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor([int(local_bytes)], dtype=torch.int64, device=dev)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [int(t.item()) for t in out]


@dataclass
class RawDeviceBuffer:
    """A device allocation the library owns (peer buffers), dressed as a buffer carrier."""

    ptr: int
    shape: tuple[int, ...]
    itemsize: int
    is_cuda: bool = True

    def data_ptr(self) -> int:
        return self.ptr

    def is_contiguous(self) -> bool:
        return True

    def dim(self) -> int:
        return len(self.shape)

    def stride(self, d: int) -> int:
        s = 1
        for x in self.shape[d + 1:]:
            s *= x
        return s

    def element_size(self) -> int:
        return self.itemsize

    def slice0(self, start: int, stop: int) -> "RawDeviceBuffer":
        return RawDeviceBuffer(self.ptr + start * self.stride(0) * self.itemsize, (stop - start, *self.shape[1:]), self.itemsize)


class CloudGather:
    """Gather of per-rank clouds (dense ``[..., 3]`` f32 blocks or compacted point lists) on ``root``."""

    def __init__(self, ctx: IngestContext | None, rank: int, world: int, root: int = 0, group: Any = None) -> None:
        self.ctx, self.rank, self.world, self.root, self.group = ctx, rank, world, root, group
        self.backend = dist.get_backend(group) if dist.is_initialized() else "none"
        self._nccl_ready = False
        self._keep: Any = None
        self._pending_counts: list[int] | None = None

    def _ensure_nccl(self) -> None:
        if self._nccl_ready:
            return
        assert self.ctx is not None
        uid = [self.ctx.nccl_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(uid, src=0, group=self.group)
        self.ctx.nccl_init(uid[0], self.rank, self.world)
        self._nccl_ready = True

    def gather(self, local: torch.Tensor, gathered: torch.Tensor | None = None, wait: bool = True) -> torch.Tensor | None:
        """Returns the fused buffer on ``root`` (``None`` elsewhere).  ``local`` may differ in length per rank.

        ``wait=False`` (NCCL only) returns as soon as the exchange is enqueued on the library's exchange stream; the caller
        goes on enqueueing the next batch and calls :meth:`wait` before touching ``gathered`` or overwriting ``local``."""
        nbytes = local.numel() * local.element_size()
        sizes = exchange_counts(nbytes, self.group)
        return self._gather_sized(local, gathered, sizes, wait)

    def _gather_sized(self, local: torch.Tensor, gathered: torch.Tensor | None, sizes: list[int], wait: bool) -> torch.Tensor | None:
        total = sum(sizes)
        nbytes = sizes[self.rank]
        if self.backend == "nccl" and local.is_cuda:
            self._ensure_nccl()
            if self.rank == self.root and gathered is None:
                gathered = torch.empty(total // local.element_size(), dtype=local.dtype, device=local.device)
            self._keep = local.contiguous()  # the send reads it asynchronously: keep it alive until wait()
            self.ctx.gather_clouds(self._keep, gathered, sizes, self.root)
            if wait:
                self.wait()
            return gathered if self.rank == self.root else None
        # gloo / CPU: host-logic path used by the CPU tests
        flat = local.contiguous().view(-1)[: nbytes // local.element_size()]
        if self.rank == self.root:
            parts = [torch.empty(s // local.element_size(), dtype=local.dtype) for s in sizes]
            parts[self.root] = flat
            for r in range(self.world):
                if r != self.root and sizes[r]:
                    dist.recv(parts[r], src=r, group=self.group)
            return torch.cat(parts) if gathered is None else gathered.view(-1)[: total // local.element_size()].copy_(torch.cat(parts))
        if nbytes:
            dist.send(flat, dst=self.root, group=self.group)
        return None

    def wait(self, on_stream: bool = False) -> None:
        """Block (or make the ingest stream wait) until the last ``gather(..., wait=False)`` has landed."""
        if self.backend == "nccl" and self.ctx is not None and self._nccl_ready:
            self.ctx.gather_wait(on_stream)
            if not on_stream:
                self._keep = None

    def gather_records(self, records: torch.Tensor, n_records: torch.Tensor, gathered: torch.Tensor | None = None,
                       wait: bool = True) -> tuple[torch.Tensor | None, list[int]]:
        """Variable-length gather of ``ti_voxel_cloud`` lists: rank r contributes ``records[:n_records]`` (a DEVICE count).
        Returns (fused list on root else ``None``, records per rank).  = :meth:`records_begin` + :meth:`records_send`."""
        self.records_begin(n_records)
        return self.records_send(records, gathered, wait)

    def records_begin(self, n_records: torch.Tensor) -> None:
        """First half: all-gather the device counts on the exchange stream, behind what the ingest stream holds now.  A pipelined
        caller enqueues its next batch between this call and :meth:`records_send`, which blocks for the counts."""
        if self.backend == "nccl" and n_records.is_cuda:
            self._ensure_nccl()
            self.ctx.gather_counts_begin(n_records)
            self._pending_counts = None
        else:
            mine = torch.tensor([int(n_records.view(-1)[0].item())], dtype=torch.int64)
            out = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(out, mine, group=self.group)
            self._pending_counts = [int(t.item()) for t in out]

    def records_send(self, records: torch.Tensor, gathered: torch.Tensor | None = None, wait: bool = True) -> tuple[torch.Tensor | None, list[int]]:
        """Second half: collect the counts, then the grouped send / receive of exactly that many records per rank."""
        cap = int(records.shape[0])
        if self._pending_counts is None:
            counts = [min(c, cap) for c in self.ctx.gather_counts_finish(self.world)]  # a truncated list still reports its full count
            total = sum(counts)
            if self.rank == self.root and gathered is None:
                gathered = torch.empty(total, dtype=records.dtype, device=records.device)
            self._keep = records
            self.ctx.gather_records(records, gathered, counts, self.root)
            if wait:
                self.wait()
            return (gathered if self.rank == self.root else None), counts
        counts = [min(c, cap) for c in self._pending_counts]
        sizes = [c * records.element_size() for c in counts]
        return self._gather_sized(records, gathered, sizes, wait), counts


class RecordExchange:
    """The cloud exchange as the library's own kernels over peer memory (``thor_slam_b200/csrc/ti_push.cu``).

    The root owns ``slots`` inboxes (used round robin, so a producer does not wait for the root to drain the previous round);
    every OTHER rank appends its ``ti_voxel_cloud`` list with :meth:`push` (on the root the call only takes a fence: its own
    list is already where the fusion runs); the root collects a round with :meth:`take`.  A round's cloud is therefore the
    inbox plus the root's own list.  Everything runs on the library's exchange stream behind an event of the ingest stream.
    Runs of different ranks start 16-byte aligned in the inbox: odd lists are padded with one zero record - skip zeros."""

    def __init__(self, ctx: IngestContext, rank: int, world: int, capacity: int, root: int = 0, slots: int = 2, group: Any = None) -> None:
        from thor_slam_b200.ingest._lib import INBOX_HEADER_BYTES

        self.ctx, self.rank, self.world, self.root, self.capacity, self.slots = ctx, rank, world, root, int(capacity), slots
        self.group = group
        nbytes = INBOX_HEADER_BYTES + 8 * self.capacity
        handles: list[Any] = [None] * slots
        self._owned: list[int] = []
        self._mapped: list[int] = []
        if rank == root:
            for k in range(slots):
                ptr, h = ctx.peer_alloc(nbytes)
                ctx.inbox_init(ptr)
                self._owned.append(ptr)
                handles[k] = h
        dist.broadcast_object_list(handles, src=root, group=group)
        if rank == root:
            self.inbox = list(self._owned)
        else:
            self._mapped = [ctx.peer_open(h) for h in handles]
            self.inbox = list(self._mapped)
        self.round = 0
        self.taken = 0

    def push(self, records: Any, n_records: Any) -> int:
        """Append this rank's list for the current round; rounds advance with every call.  Returns a fence: once it has
        passed (``ctx.exchange_wait(fence, on_stream=True)``), ``records`` may be overwritten."""
        k = self.round
        if self.rank != self.root:
            self.ctx.cloud_push(records, n_records, self.inbox[k % self.slots], self.capacity, k // self.slots)
        self.round += 1
        return self.ctx.exchange_fence()

    def take(self, dst: Any, status: Any) -> None:
        """Root only: the next round's fused list into ``dst`` (u64 [>= capacity]), ``status`` = (count, error flag)."""
        assert self.rank == self.root
        k = self.taken
        self.ctx.inbox_take(self.inbox[k % self.slots], self.capacity, self.world - 1, dst, status)
        self.taken += 1

    def wait(self, on_stream: bool = False) -> None:
        self.ctx.gather_wait(on_stream)

    def close(self) -> None:
        """Collective: nobody may still be storing into the root's inboxes when they are freed."""
        self.ctx.gather_wait(False)
        if dist.is_initialized():
            dist.barrier(group=self.group)
        for p in self._mapped:
            self.ctx.peer_close(p)
        self._mapped = []
        if dist.is_initialized():
            dist.barrier(group=self.group)
        for p in self._owned:
            self.ctx.peer_free(p)
        self._owned = []


class PeerCloudBuffer:
    """Fused gather: every rank's back-projection writes its slice of ONE buffer that lives on ``root``."""

    def __init__(self, ctx: IngestContext, rank: int, world: int, shape_per_rank: tuple[int, ...], root: int = 0, group: Any = None) -> None:
        self.ctx, self.rank, self.world, self.root = ctx, rank, world, root
        self.shape_per_rank = tuple(shape_per_rank)
        per_rank = 4
        for x in shape_per_rank:
            per_rank *= x
        self.bytes_per_rank = per_rank
        handle = [None]
        self._owned = self._mapped = None
        if rank == root:
            self._owned, h = ctx.peer_alloc(per_rank * world)
            handle = [h]
        dist.broadcast_object_list(handle, src=root, group=group)
        base = self._owned if rank == root else ctx.peer_open(handle[0])
        if rank != root:
            self._mapped = base
        self.base = base

    def slice_for(self, rank: int) -> RawDeviceBuffer:
        return RawDeviceBuffer(self.base + rank * self.bytes_per_rank, self.shape_per_rank, 4)

    def whole(self) -> RawDeviceBuffer:
        return RawDeviceBuffer(self.base, (self.world, *self.shape_per_rank), 4)

    def as_tensor(self) -> torch.Tensor:
        """Root only: the fused buffer as a torch tensor (zero-copy through __cuda_array_interface__)."""
        assert self.rank == self.root

        class _CAI:
            pass

        obj = _CAI()
        obj.__cuda_array_interface__ = {"shape": (self.world, *self.shape_per_rank), "typestr": "<f4",
                                        "data": (self.base, False), "version": 3, "strides": None}
        return torch.as_tensor(obj, device=torch.device("cuda", self.ctx.device))

    def close(self) -> None:
        """Collective: peers unmap first, then the owner frees (a barrier on either side, so no store is in flight)."""
        self.ctx.sync()
        if dist.is_initialized():
            dist.barrier()
        if self._mapped is not None:
            self.ctx.peer_close(self._mapped)
            self._mapped = None
        if dist.is_initialized():
            dist.barrier()
        if self._owned is not None:
            self.ctx.peer_free(self._owned)
            self._owned = None
