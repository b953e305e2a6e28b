"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE (see oracle/oracle.hpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_lib = None


def build(force: bool = False) -> Path:
    so = HERE / "liboracle.so"
    srcs = [HERE / "oracle.cpp", HERE / "capi.cpp", HERE / "encoder.cpp", HERE / "rfc3164.cpp", HERE / "oracle.hpp"]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(HERE), "-s", "-B", "liboracle.so"], check=True)
    return so


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        so = HERE / "liboracle.so"
        if not so.exists():
            build()
        L = C.CDLL(str(so))
        L.fgo_ltsv_config_new.restype = C.c_void_p
        L.fgo_ltsv_config_free.argtypes = [C.c_void_p]
        L.fgo_ltsv_config_add_schema.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.fgo_ltsv_config_enable_schema.argtypes = [C.c_void_p]
        L.fgo_ltsv_config_set_suffix.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        L.fgo_free.argtypes = [C.c_void_p]
        L.fgo_rfc3164_config_new.restype = C.c_void_p
        L.fgo_rfc3164_config_new.argtypes = [C.c_int]
        L.fgo_rfc3164_config_free.argtypes = [C.c_void_p]
        L.fgo_rfc3164_config_add_zone.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p]
        L.fgo_decode_dump.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.fgo_decode_bench.restype = C.c_double
        L.fgo_decode_bench.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_int64)]
        L.fgo_decode_debug.restype = C.c_void_p
        L.fgo_decode_debug.argtypes = [C.c_int, C.c_void_p, C.c_char_p, C.c_int64]
        L.fgo_g15.restype = C.c_void_p
        L.fgo_gelf_encoder_test.restype = C.c_void_p
        L.fgo_format_f64.restype = C.c_void_p
        L.fgo_format_f64.argtypes = [C.c_double]
        L.fgo_decode_encode_gelf.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                             C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.fgo_parse_f64.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_double)]
        L.fgo_rfc3339.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_double)]
        L.fgo_english.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.POINTER(C.c_double)]
        _lib = L
    return _lib


_TYPES = {"string": 0, "bool": 1, "f64": 2, "i64": 3, "u64": 4}


class LtsvConfig:
    """ltsv_decoder.rs:24-83 (type names and suffix keys are case-insensitive)."""

    def __init__(self, schema: dict[str, str] | None = None, suffixes: dict[str, str] | None = None):
        L = lib()
        self.h = C.c_void_p(L.fgo_ltsv_config_new())
        if schema is not None:
            L.fgo_ltsv_config_enable_schema(self.h)
            for k, t in schema.items():
                L.fgo_ltsv_config_add_schema(self.h, k.encode(), _TYPES[t.lower()])
        for t, s in (suffixes or {}).items():
            L.fgo_ltsv_config_set_suffix(self.h, _TYPES[t.lower()], s.encode())

    def __del__(self):
        try:
            lib().fgo_ltsv_config_free(self.h)
        except Exception:
            pass


class Rfc3164Config:
    """The two inputs RFC3164Decoder takes from its environment (rfc3164_decoder.rs:175 current year, :196 zone database).
    `zones`: name -> (transitions, offsets) as oracle/tzread.py produces them; default = the system's TZif files."""

    _system_zones = None

    def __init__(self, year: int, zones: dict | None = None):
        L = lib()
        if zones is None:
            if Rfc3164Config._system_zones is None:
                import tzread
                Rfc3164Config._system_zones = tzread.load_zones()
            zones = Rfc3164Config._system_zones
        self.year = year
        self.zones = zones
        self.h = C.c_void_p(L.fgo_rfc3164_config_new(year))
        for name, (trans, offs) in zones.items():
            t = np.asarray(trans, dtype=np.int64)
            o = np.asarray(offs, dtype=np.int32)
            assert len(o) == len(t) + 1
            L.fgo_rfc3164_config_add_zone(self.h, name.encode(), len(t), C.c_void_p(t.ctypes.data), C.c_void_p(o.ctypes.data))

    def __del__(self):
        try:
            lib().fgo_rfc3164_config_free(self.h)
        except Exception:
            pass


def decode_debug(fmt: int, line: bytes | str, cfg: "LtsvConfig | Rfc3164Config | None" = None) -> str:
    b = line.encode() if isinstance(line, str) else line
    p = lib().fgo_decode_debug(fmt, cfg.h if cfg else None, b, len(b))
    s = C.string_at(p).decode()
    lib().fgo_free(C.c_void_p(p))
    return s


def g15(which: int) -> str:
    p = lib().fgo_g15(which)
    s = C.string_at(p).decode()
    lib().fgo_free(C.c_void_p(p))
    return s


def gelf_encoder_test(which: int) -> str:
    """JSON of the Record built by the reference's encoder test #which (gelf_encoder.rs:125,152,175,215)."""
    p = lib().fgo_gelf_encoder_test(which)
    s = C.string_at(p).decode()
    lib().fgo_free(C.c_void_p(p))
    return s


def format_f64(v: float) -> str:
    p = lib().fgo_format_f64(v)
    s = C.string_at(p).decode()
    lib().fgo_free(C.c_void_p(p))
    return s


def decode_encode_gelf(fmt: int, data: np.ndarray, offsets: np.ndarray, extra: dict[str, str] | None = None,
                       cfg: "LtsvConfig | None" = None, nthreads: int = 8) -> tuple[bytes, np.ndarray]:
    """decode + GelfEncoder::encode per line (line_splitter.rs:50-52): (concatenated JSON, int64 offsets[n+1]); lines the
    decoder rejects contribute an empty record."""
    n = len(offsets) - 1
    ex = list((extra or {}).items())
    keys = (C.c_char_p * max(len(ex), 1))(*[k.encode() for k, _ in ex])
    vals = (C.c_char_p * max(len(ex), 1))(*[v.encode() for _, v in ex])
    pb, po = C.c_void_p(), C.c_void_p()
    lib().fgo_decode_encode_gelf(fmt, cfg.h if cfg else None, C.c_void_p(data.ctypes.data), C.c_void_p(offsets.ctypes.data), n,
                                 nthreads, len(ex), keys, vals, C.byref(pb), C.byref(po))
    try:
        offs = np.ctypeslib.as_array(C.cast(po, C.POINTER(C.c_int64)), shape=(n + 1,)).copy()
        buf = C.string_at(pb, int(offs[-1]))
    finally:
        lib().fgo_free(pb)
        lib().fgo_free(po)
    return buf, offs


def decode_dump(fmt: int, data: np.ndarray, offsets: np.ndarray, cfg: LtsvConfig | None = None,
                nthreads: int = 8) -> tuple[bytes, np.ndarray]:
    assert data.dtype == np.uint8 and offsets.dtype == np.int32
    n = len(offsets) - 1
    pb, po = C.c_void_p(), C.c_void_p()
    lib().fgo_decode_dump(fmt, cfg.h if cfg else None, C.c_void_p(data.ctypes.data), C.c_void_p(offsets.ctypes.data),
                          n, nthreads, C.byref(pb), C.byref(po))
    try:
        offs = np.ctypeslib.as_array(C.cast(po, C.POINTER(C.c_int64)), shape=(n + 1,)).copy()
        buf = C.string_at(pb, int(offs[-1]))
    finally:
        lib().fgo_free(pb)
        lib().fgo_free(po)
    return buf, offs


def decode_bench(fmt: int, data: np.ndarray, offsets: np.ndarray, cfg: LtsvConfig | None = None,
                 nthreads: int = 1) -> tuple[float, int]:
    n = len(offsets) - 1
    ok = C.c_int64()
    s = lib().fgo_decode_bench(fmt, cfg.h if cfg else None, C.c_void_p(data.ctypes.data),
                               C.c_void_p(offsets.ctypes.data), n, nthreads, C.byref(ok))
    return float(s), int(ok.value)


def parse_f64(s: bytes):
    d = C.c_double()
    return d.value if lib().fgo_parse_f64(s, len(s), C.byref(d)) else None


def rfc3339(s: bytes):
    d = C.c_double()
    return d.value if lib().fgo_rfc3339(s, len(s), C.byref(d)) else None


def english(s: bytes, with_subsecond: bool):
    d = C.c_double()
    return d.value if lib().fgo_english(s, len(s), 1 if with_subsecond else 0, C.byref(d)) else None


def pack(lines: list[bytes]) -> tuple[np.ndarray, np.ndarray]:
    offs = np.zeros(len(lines) + 1, dtype=np.int32)
    np.cumsum([len(x) for x in lines], out=offs[1:])
    data = np.frombuffer(b"".join(lines) or b"\0", dtype=np.uint8).copy()
    return data, offs
