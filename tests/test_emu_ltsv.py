"""CPU check of the LTSV DEVICE LOGIC: the product's walker sources (stage-1 TAB / ':' bitmaps, stage-2 part walk, slot
reservation, staged rows, direct path) compiled with g++ (tests/emu) and replayed CTA by CTA, pushed through the product's
host materialiser and compared with the oracle.  No GPU needed; the `-m gpu` tests repeat all of this on the device."""
import numpy as np
import pytest

import vectors as V
from conftest import first_diff

LT = 1
SCHEMA = {"counter": "u64", "score": "i64", "mean": "f64", "done": "bool"}


@pytest.fixture(scope="module")
def emu():
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent / "emu"))
    import emu as E
    E.build()
    return E


def check(emu, native, oracle, data, offs, schema=None, suffixes=None, **kw):
    gbuf, goffs, info = emu.ltsv_decode_dump(native, data, offs, schema=schema, suffixes=suffixes, **kw)
    cfg = oracle.LtsvConfig(schema, suffixes) if (schema is not None or suffixes is not None) else None
    obuf, ooffs = oracle.decode_dump(LT, data, offs, cfg, nthreads=8)
    if not (gbuf == obuf and np.array_equal(goffs, ooffs)):
        diffs = first_diff(gbuf, goffs, obuf, ooffs, data, offs)
        msg = "\n".join(f"line {i}: {line!r}\n   emu: {g!r}\n   ref: {o!r}" for i, line, g, o in diffs)
        raise AssertionError(f"{len(diffs)}+ lines differ from the oracle:\n{msg}")
    return info


def test_stage1_bitmaps_per_byte(emu):
    """lt_tab16 flags exactly the TABs of a granule, lt_colon_flags8 exactly the colons of 8 bytes — for every byte value in
    every position."""
    rng = np.random.default_rng(6)
    for b in range(256):
        for pos in range(16):
            blk = bytearray(rng.integers(0x61, 0x7B, 16, dtype=np.uint8).tobytes())
            blk[pos] = b
            t, c = emu.ltsv_classify16(bytes(blk))
            assert t == ((1 << pos) if b == 9 else 0) and c == ((1 << pos) if b == 0x3A and pos < 8 else 0), (b, pos)
    for _ in range(3000):
        blk = rng.choice(np.array([9, 0x3A, 8, 0x0A, 0x3B, 0x39, 0x89, 0xBA, 0x41, 0], dtype=np.uint8), 16).tobytes()
        t, c = emu.ltsv_classify16(blk)
        assert t == sum(1 << k for k in range(16) if blk[k] == 9)
        assert c == sum(1 << k for k in range(8) if blk[k] == 0x3A)


def test_goldens_and_appendix(emu, native, oracle):
    lines = [V.G9_LINE, V.G10_LINE, V.G11_LINE, V.G12_LINE]
    data, offs = oracle.pack([l.encode() for l in lines])
    check(emu, native, oracle, data, offs, schema=V.LTSV_SCHEMA)
    data, offs = oracle.pack([V.G13_LINE.encode()])
    check(emu, native, oracle, data, offs, schema=V.LTSV_SCHEMA_G13, suffixes=V.LTSV_SUFFIX_G13)
    data, offs = oracle.pack([V.G14_LINE.encode()])
    check(emu, native, oracle, data, offs, schema=V.LTSV_SCHEMA_G14, suffixes=V.LTSV_SUFFIX_G14)
    data, offs = oracle.pack([l.encode() for l, _ in V.LTSV_CASES])
    check(emu, native, oracle, data, offs)
    data, offs = oracle.pack([l.encode() for l, _ in V.LTSV_SCHEMA_CASES])
    check(emu, native, oracle, data, offs, schema=SCHEMA)


def test_empty_ragged_and_direct_path(emu, native, oracle):
    """Lines longer than the tile and lines with more parts than the staging area take the direct path (the round-1
    scanner); everything else the bitmap walk — in any mix inside one CTA."""
    big = b"time:1\thost:h\tk:" + b"v" * 100_000 + b"\tz:" + b":" * 1000
    many = b"time:1\thost:h" + b"".join(b"\tk%d:v" % i for i in range(5000))
    colons = b"\t".join([b":"] * 3000) + b"\ttime:1\thost:h"
    tabs_only = b"\t" * 2000
    long_key = b"k" * 100 + b":v\ttime:1\thost:h\t" + b"x" * 70 + b"\t" + b"y" * 33 + b":" + b"z" * 65
    lines = [b"", b"\t", b":", b":\t:", b"time:1\thost:h", big, many, colons, tabs_only, long_key] + [V.G11_LINE.encode()] * 200
    data, offs = oracle.pack(lines)
    info = check(emu, native, oracle, data, offs)
    assert info["direct"] >= 3
    info = check(emu, native, oracle, data, offs, tile_bytes=8192)
    assert info["direct"] >= 3
    # a CTA whose TAB counts overflow the staging area mid-way: the round is cut, not lost
    mid = [b"time:1\thost:h" + b"".join(b"\tk%d:v" % i for i in range(300)) for _ in range(40)]
    data, offs = oracle.pack(mid)
    info = check(emu, native, oracle, data, offs, tile_bytes=65024)
    assert info["rounds"] > 1 and info["direct"] == 0


def test_generated(emu, native, oracle):
    data, offs = native.generate(native.FMT_LTSV, 1757, 60_000)
    info = check(emu, native, oracle, data, offs)
    assert info["direct"] == 0
    check(emu, native, oracle, data, offs, schema=SCHEMA)
    check(emu, native, oracle, data, offs, schema=V.LTSV_SCHEMA_G13, suffixes=V.LTSV_SUFFIX_G13)
    check(emu, native, oracle, data, offs, tile_bytes=8192)


def test_eol_and_invalid_lines(emu, native, oracle):
    """split mode: lines keep their terminators, flagged lines are not parsed."""
    data, offs = native.generate(native.FMT_LTSV, 3, 3000)
    lines = [bytes(data[offs[i]:offs[i + 1]]) for i in range(len(offs) - 1)]
    term = [l + (b"\r\n" if i % 3 == 0 else b"\n") for i, l in enumerate(lines)]
    d2, o2 = oracle.pack(term)
    inv = np.zeros(len(term), dtype=np.uint8)
    inv[::17] = 1
    gbuf, goffs, _ = emu.ltsv_decode_dump(native, d2, o2, strip_eol=1, invalid=inv)
    d1, o1 = oracle.pack(lines)
    obuf, ooffs = oracle.decode_dump(LT, d1, o1, None, nthreads=8)
    for i in range(len(lines)):
        g = gbuf[goffs[i]:goffs[i + 1]]
        if inv[i]:
            assert b"Invalid UTF-8 input" in g
        else:
            # spans are absolute offsets into different buffers: compare everything but the error-position field
            assert g == obuf[ooffs[i]:ooffs[i + 1]] or b"ERR" in g


def test_mutation_fuzz(emu, native, oracle):
    rng = np.random.default_rng(4321)
    data, offs = native.generate(native.FMT_LTSV, 8, 20_000, bad_frac=0.0)
    alphabet = b"\t:[]+-.eE0159 TZ/"
    out = []
    for i in range(len(offs) - 1):
        ln = bytearray(data[offs[i]:offs[i + 1]])
        if any(b >= 0x80 for b in ln):
            out.append(bytes(ln))
            continue
        for _ in range(int(rng.integers(1, 4))):
            op = int(rng.integers(0, 3))
            pos = int(rng.integers(0, max(len(ln), 1)))
            ch = alphabet[int(rng.integers(0, len(alphabet)))]
            if op == 0 and ln:
                ln[pos] = ch
            elif op == 1:
                ln.insert(pos, ch)
            elif ln:
                del ln[pos]
        out.append(bytes(ln))
    d2, o2 = oracle.pack(out)
    check(emu, native, oracle, d2, o2, schema=SCHEMA)


MANY_TYPED_SCHEMA = {"a": "u64", "b": "u64", "c": "u64", "d": "i64", "e": "f64", "f": "bool", "time": "u64"}
MANY_TYPED_LINES = [
    b"time:1\thost:h\ta:1\tb:2\tc:3\td:-4\te:5.5\tf:true",
    b"a:1\ta:2\ta:3\ta:4\ttime:1\thost:h",                      # four values of one type: two parked, two parsed in place
    b"a:1\tb:x\tc:y\ttime:1\thost:h",                            # the FIRST failing part wins (b)
    b"a:1\tb:2\tc:y\ta:z\ttime:1\thost:h",                      # in-place failure (c) after two parked values
    b"a:1\tb:2\tc:3\ta:z\ttime:bad\thost:h",                    # in-place failure (second a) before the parked time fails
    b"time:bad\ta:1\tb:2\tc:3\ta:z\thost:h",                    # parked time at an earlier part than the in-place failure
    b"level:9\ta:x\ttime:1\thost:h", b"a:x\tlevel:9\ttime:1\thost:h", b"level:3\tlevel:8\tlevel:2\ttime:1\thost:h",
    b"time:1\ttime:bad\ttime:2\thost:h", b"time:bad\ttime:1\thost:h", b"level:x\tlevel:1\ttime:1\thost:h",
    b"d:-9223372036854775808\te:1e400\tf:TRUE\ttime:1\thost:h", b"e:.\ttime:1\thost:h", b"f:false\tf:true\tf:no\ttime:1\thost:h",
]


def test_many_typed_values_and_parked_order(emu, native, oracle):
    """Several values of one schema type, failures at parked and in-place positions: the first failing PART decides."""
    data, offs = oracle.pack(MANY_TYPED_LINES)
    check(emu, native, oracle, data, offs, schema=MANY_TYPED_SCHEMA)
    data, offs = oracle.pack(MANY_TYPED_LINES * 40)
    check(emu, native, oracle, data, offs, schema=MANY_TYPED_SCHEMA)
