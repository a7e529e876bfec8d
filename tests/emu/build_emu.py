"""Build the CPU emulation of libthoringest (TEST INFRASTRUCTURE ONLY).

``g++ -DTI_EMULATE`` compiles the very kernel sources that nvcc compiles for sm_100a
(thor_slam_b200/csrc/*.cu, minus the NCCL file) against ``cuda_emu.h``; every CUDA thread becomes
a std::thread.  The result, ``tests/emu/_build/libthoringest_emu.so``, exports the same C ABI plus
the marker symbol ``ti_emu_marker`` and is loaded ONLY by tests (dependency-injected into
``IngestLibrary``); the product loader never looks for it.
"""

from __future__ import annotations

import hashlib
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE.parent.parent / "thor_slam_b200" / "csrc"
OUT_DIR = HERE / "_build"
SOURCES = ["ti_api.cu", "ti_convert.cu", "ti_rectify.cu", "ti_rectify_pair.cu", "ti_rectify_c3.cu", "ti_backproject.cu", "ti_register.cu", "ti_voxel.cu", "ti_tma.cu"]
CUDA_INC = "/usr/local/cuda/include"


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE / "cuda_emu.h", HERE / "emu_extra.cpp",
                                                                        HERE.parent.parent / "include" / "thoringest.h"]):
        h.update(f.read_bytes())
    return h.hexdigest()


def build(force: bool = False) -> Path:
    OUT_DIR.mkdir(exist_ok=True)
    out = OUT_DIR / "libthoringest_emu.so"
    stamp = OUT_DIR / "stamp"
    digest = _digest()
    if out.exists() and stamp.exists() and stamp.read_text() == digest and not force:
        return out
    cmd = ["g++", "-std=c++20", "-O2", "-g", "-shared", "-fPIC", "-pthread", "-DTI_EMULATE", "-Wno-attributes",
           "-Wno-unknown-pragmas", f"-I{HERE}", f"-I{CUDA_INC}", "-o", str(out)]
    for s in SOURCES:
        cmd += ["-x", "c++", str(CSRC / s)]
    cmd += ["-x", "c++", str(HERE / "emu_extra.cpp")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulation build failed:\n" + res.stdout + res.stderr)
    stamp.write_text(digest)
    return out


if __name__ == "__main__":
    print(build(force=True))
