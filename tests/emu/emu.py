"""Loader of tests/emu/libfg_emu.so — the CPU emulation of the RFC5424 device logic (TEST INFRASTRUCTURE).

The library is the product's walker sources (flowgger_b200/csrc/fg_r5fast.cuh, fg_rfc5424.cuh) compiled with g++ plus a
driver that replays the CTA rounds of parse5424_kernel one lane at a time.  Only tests import this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
_lib = None


def build(force: bool = False) -> Path:
    so = HERE / "libfg_emu.so"
    csrc = REPO / "flowgger_b200" / "csrc"
    srcs = [HERE / "emu_r5.cpp", HERE / "cuda_shim.h", REPO / "include" / "flowgger_cuda.h"] + sorted(csrc.glob("*.cuh")) + \
        sorted(csrc.glob("*.h"))
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", str(so),
                        str(HERE / "emu_r5.cpp"), "-I", str(REPO / "include")], check=True)
    return so


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.emu5424_classify16.restype = C.c_uint32
        _lib.emu5424_classify16.argtypes = [C.c_void_p]
    return _lib


def classify16(block: bytes) -> int:
    assert len(block) == 16
    buf = C.create_string_buffer(block, 16)
    return int(lib().emu5424_classify16(buf))


def decode_dump(native, data: np.ndarray, offsets: np.ndarray, tile_bytes: int = 13312, strip_eol: bool = False,
                invalid: np.ndarray | None = None):
    """Emulated RFC5424 decode -> (canonical dumps, dump offsets, FgBatchOut fields as a dict) through the product's
    host materialiser (fgh_dump_records)."""
    from flowgger_b200.native import FgBatchOut, dump_records
    out = FgBatchOut()
    n = len(offsets) - 1
    lib().emu5424_decode(C.c_void_p(data.ctypes.data), C.c_void_p(offsets.ctypes.data), n, tile_bytes, 1 if strip_eol else 0,
                         C.c_void_p(invalid.ctypes.data) if invalid is not None else None, C.byref(out))
    try:
        buf, offs = dump_records(native.FMT_RFC5424, out, data, offsets)
        info = {"n_entries8": out.n_entries8, "n_wide": out.n_wide, "arena_bytes": out.arena_bytes, "n_entries": out.n_entries}
    finally:
        lib().emu5424_free(C.byref(out))
    return buf, offs, info
