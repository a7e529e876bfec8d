"""pytest configuration: the ``gpu`` marker and the two backends the parity cases run on.

* ``emu``  - tests/emu: the library's own kernel sources compiled by g++ and run on CPU threads
             (catches indexing / packing bugs without a GPU; test infrastructure only);
* ``gpu``  - the real ``libthoringest.so`` on ``cuda:0`` (tests marked ``@pytest.mark.gpu``).
"""

from __future__ import annotations

import ctypes
import random
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config: pytest.Config) -> None:
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running")


@pytest.fixture(autouse=True)
def _seed() -> None:
    random.seed(1337)  # mirrors the reference's tests/conftest.py:9-11
    np.random.seed(1337)


class Backend:
    """Array plumbing for one backend: host numpy in, backend buffers, host numpy out."""

    def __init__(self, name: str, ctx) -> None:
        self.name = name
        self.ctx = ctx

    def dev(self, arr: np.ndarray):
        if self.name == "emu":
            return np.ascontiguousarray(arr).copy()
        import torch

        if arr.dtype == np.uint16:
            return torch.from_numpy(arr.view(np.int16).copy()).cuda().view(torch.uint16)
        if arr.dtype == np.uint32:
            return torch.from_numpy(arr.view(np.int32).copy()).cuda().view(torch.uint32)
        return torch.from_numpy(np.ascontiguousarray(arr).copy()).cuda()

    def zeros(self, shape, dtype):
        return self.dev(np.zeros(shape, dtype=dtype))

    def host(self, buf) -> np.ndarray:
        if self.name == "emu":
            return np.asarray(buf)
        import torch

        self.ctx.sync()
        torch.cuda.synchronize()
        if buf.dtype == torch.uint16:
            return buf.view(torch.int16).cpu().numpy().view(np.uint16)
        if buf.dtype == torch.uint32:
            return buf.view(torch.int32).cpu().numpy().view(np.uint32)
        return buf.cpu().numpy()


@pytest.fixture(scope="session")
def emu_backend() -> Backend:
    from tests.emu.build_emu import build

    from thor_slam_b200.ingest._lib import IngestLibrary
    from thor_slam_b200.ingest.context import IngestContext

    lib = IngestLibrary(ctypes.CDLL(str(build())))
    assert lib.is_emulation
    return Backend("emu", IngestContext(0, lib))


@pytest.fixture(scope="session")
def gpu_backend() -> Backend:
    import torch

    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    from thor_slam_b200.ingest.context import IngestContext

    torch.cuda.set_device(0)
    ctx = IngestContext(0)  # the in-tree libthoringest.so; raises if it was not built
    assert not ctx.lib.is_emulation
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    return Backend("gpu", ctx)
