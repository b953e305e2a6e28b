"""Build the in-tree native libraries (sm_100a only).

    python -m flowgger_b200.build [--force]

Produces, under flowgger_b200/lib/:
  libflowgger_cuda.so   the C-ABI decoder (CUDA kernels + host pipeline), include/flowgger_cuda.h
  libflowgger_host.so   C++ mirror of the reference's Decoder/Record/Splitter interface on top of the C ABI
  libfg_gen.so          synthetic log generators for tests and bench (not product code)
The oracle (oracle/liboracle.so) is test infrastructure and is built by its own Makefile.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
REPO = ROOT.parent
CSRC = ROOT / "csrc"
LIB = ROOT / "lib"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-O3",
    "--fmad=false",          # keep IEEE double semantics explicit (timestamp recipe)
    "-Xptxas", "-v",
]
CXX = os.environ.get("CXX", "g++")
CXX_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-Wall", "-Wextra", "-pthread", "-march=x86-64-v3"]


def _newer(target: Path, sources: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(s.stat().st_mtime > t for s in sources)


def _run(cmd: list[str]) -> None:
    print("+", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    out = r.stdout
    if r.returncode != 0:
        sys.stdout.write(out)
        raise RuntimeError(f"build step failed: {' '.join(cmd)}")
    # keep the ptxas resource lines: they are the first thing to read before GPU time
    for line in out.splitlines():
        if "registers" in line or "spill" in line or "error" in line.lower():
            print("   ", line.strip())


def build_cuda(force: bool = False) -> Path:
    LIB.mkdir(exist_ok=True)
    target = LIB / "libflowgger_cuda.so"
    srcs = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [
        REPO / "include" / "flowgger_cuda.h"]
    if force or _newer(target, srcs):
        cus = [str(p) for p in sorted(CSRC.glob("*.cu"))]
        _run([NVCC, *NVCC_FLAGS, "-shared", "-o", str(target), *cus, "-I", str(REPO / "include")])
    return target


def build_host(force: bool = False) -> Path | None:
    hdir = CSRC / "host"
    if not hdir.exists():
        return None
    target = LIB / "libflowgger_host.so"
    srcs = sorted(hdir.glob("*.cpp")) + sorted(hdir.glob("*.hpp")) + [REPO / "include" / "flowgger_cuda.h"]
    if force or _newer(target, srcs):
        _run([CXX, *CXX_FLAGS, "-shared", "-o", str(target), *[str(p) for p in sorted(hdir.glob("*.cpp"))],
              "-I", str(REPO / "include"), f"-L{LIB}", "-lflowgger_cuda", "-Wl,-rpath,$ORIGIN"])
    return target


def build_gen(force: bool = False) -> Path | None:
    gdir = CSRC / "gen"
    if not gdir.exists():
        return None
    target = LIB / "libfg_gen.so"
    srcs = sorted(gdir.glob("*.cpp")) + sorted(gdir.glob("*.hpp"))
    if force or _newer(target, srcs):
        _run([CXX, *CXX_FLAGS, "-shared", "-o", str(target), *[str(p) for p in sorted(gdir.glob("*.cpp"))]])
    return target


def build_oracle(force: bool = False) -> Path:
    odir = REPO / "oracle"
    target = odir / "liboracle.so"
    srcs = [odir / "oracle.cpp", odir / "capi.cpp", odir / "encoder.cpp", odir / "rfc3164.cpp", odir / "oracle.hpp"]
    if force or _newer(target, srcs):
        _run(["make", "-C", str(odir), "-B" if force else "-s", "liboracle.so"])
    return target


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_host(force)
    build_gen(force)
    build_oracle(force)


if __name__ == "__main__":
    build_all("--force" in sys.argv)
    print("ok")
