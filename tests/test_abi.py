"""The C-ABI boundary: header, binding and built library agree (no GPU needed, no compute calls)."""

from __future__ import annotations

import ctypes
import re
import subprocess
from pathlib import Path

import pytest

from thor_slam_b200.ingest import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "thoringest.h"


def header_functions() -> list[str]:
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return re.findall(r"^\s*(?:const\s+)?(?:int|uint64_t|char\s*\*)\s+\*?(ti_\w+)\s*\(", text, flags=re.M)


def test_binding_covers_every_declared_symbol():
    declared = set(header_functions())
    assert len(declared) >= 24
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    if not _lib.LIB_PATH.exists():
        import __graft_entry__

        __graft_entry__.build()
    lib = _lib.IngestLibrary()  # dlopen, resolves every symbol, checks the ABI version
    assert lib.ti_abi_version() == _lib.ABI_VERSION
    assert not lib.is_emulation
    nm = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\sT\s+(ti_\w+)", nm))
    assert set(header_functions()) <= exported


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layout_matches_header():
    """sizeof / offsets of ti_stream as the C compiler sees them."""
    src = ROOT / "tests" / "emu" / "_build" / "layout.c"
    src.parent.mkdir(exist_ok=True)
    fields = [f[0] for f in _lib.TiStream._fields_]
    body = "".join(f'printf("{f} %zu\\n", offsetof(ti_stream, {f}));' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "thoringest.h"\nint main(){printf("size %zu\\n", sizeof(ti_stream));' + body + "return 0;}")
    exe = src.with_suffix("")
    subprocess.run(["gcc", "-std=c99", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)], check=True)  # header is plain C
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    assert int(got["size"]) == ctypes.sizeof(_lib.TiStream)
    for f in fields:
        assert int(got[f]) == getattr(_lib.TiStream, f).offset, f


def test_no_device_fails_loudly():
    """Without a GPU the product path must raise - never fall back to a CPU implementation."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from thor_slam_b200.ingest.context import IngestContext

    with pytest.raises(RuntimeError, match="no CUDA device|CPU fallback"):
        IngestContext(0)


def test_product_package_never_imports_the_oracle_or_the_emulation():
    for path in (ROOT / "thor_slam_b200").rglob("*.py"):
        text = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
        assert "tests.emu" not in text and "libthoringest_emu" not in text, path
