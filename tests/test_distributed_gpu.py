"""N = 2 on real GPUs: every multi-GPU path of the library against the oracle (``tests/distributed_worker.py``): NCCL gather
of ragged dense clouds on the exchange stream, the gather fused into the back-projection kernel through a peer-mapped buffer,
the variable-length voxel-list gather (NCCL, asynchronous) and the same exchange as peer-store kernels.  Needs two visible
GPUs (skipped otherwise; ``__graft_entry__.smoke()`` runs the same check whenever two GPUs are visible, and ``bench.py`` checks
its gathered lists against the oracle at every N > 1)."""

from __future__ import annotations

import pytest


@pytest.mark.gpu
def test_two_gpu_gather_and_peer_store():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from tests.distributed_worker import run_two_ranks

    out = run_two_ranks()
    assert out.get("gather_ok") and out.get("peer_ok"), out
    assert out.get("records_nccl_ok") and out.get("records_push_ok") and out.get("records_push_tma_ok"), out
