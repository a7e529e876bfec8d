"""Multi-GPU side of the ingest path: one process per GPU, frame sets sharded, ONE cloud gather.

The path shards with no data-path collective: every stream of every frame set is independent (the
LUTs and body transforms are per-camera constants replicated on every GPU).  The only exchange step
is the gather of the per-GPU body-frame point clouds on the fusing rank:

* ``shard_frame_sets``     - frame set ``i`` -> rank ``i % world`` (perfect balance, any batch size);
* ``exchange_counts``      - tiny all-gather of per-rank byte counts over ``torch.distributed`` (plumbing);
* ``CloudGather.gather``   - grouped ``ncclSend``/``ncclRecv`` inside ``libthoringest.so`` on the ingest
                             stream (``ti_gather_clouds``), or ``torch.distributed.gather`` when the process
                             group is ``gloo`` (CPU tests of the host logic);
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

    def _ensure_nccl(self) -> None:
        if self._nccl_ready:
            return
        assert self.ctx is not None
        uid = [self.ctx.nccl_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(uid, src=0, group=self.group)
        self.ctx.nccl_init(uid[0], self.rank, self.world)
        self._nccl_ready = True

    def gather(self, local: torch.Tensor, gathered: torch.Tensor | None = None) -> torch.Tensor | None:
        """Returns the fused buffer on ``root`` (``None`` elsewhere).  ``local`` may differ in length per rank."""
        nbytes = local.numel() * local.element_size()
        sizes = exchange_counts(nbytes, self.group)
        total = sum(sizes)
        if self.backend == "nccl" and local.is_cuda:
            self._ensure_nccl()
            if self.rank == self.root and gathered is None:
                gathered = torch.empty(total // local.element_size(), dtype=local.dtype, device=local.device)
            self.ctx.gather_clouds(local.contiguous(), gathered, sizes, self.root)
            return gathered if self.rank == self.root else None
        # gloo / CPU: host-logic path used by the CPU tests
        flat = local.contiguous().view(-1)
        if self.rank == self.root:
            parts = [torch.empty(s // local.element_size(), dtype=local.dtype) for s in sizes]
            parts[self.root] = flat
            for r in range(self.world):
                if r != self.root and sizes[r]:
                    dist.recv(parts[r], src=r, group=self.group)
            return torch.cat(parts) if gathered is None else gathered.view(-1).copy_(torch.cat(parts))
        if nbytes:
            dist.send(flat, dst=self.root, group=self.group)
        return None


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
        if self._mapped is not None:
            self.ctx.peer_close(self._mapped)
            self._mapped = None
        if self._owned is not None:
            self.ctx.peer_free(self._owned)
            self._owned = None
