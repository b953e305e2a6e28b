"""CPU check of the GELF DEVICE LOGIC: the product's walker sources (stage-1 string bitmap, stage-2 member walk for regular
lines, phase 2, the exact parser behind the slow list) compiled with g++ (tests/emu) and replayed CTA by CTA, pushed
through the product's host materialiser and compared with the oracle.  No GPU needed; the `-m gpu` tests repeat all of
this on the device."""
import numpy as np
import pytest

import vectors as V
from conftest import first_diff

GE = 2


@pytest.fixture(scope="module")
def emu():
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent / "emu"))
    import emu as E
    E.build()
    return E


def check(emu, native, oracle, data, offs, **kw):
    data = np.concatenate([data, np.zeros(64, dtype=np.uint8)])  # the exact parser reads whole 16-byte blocks
    gbuf, goffs, info = emu.gelf_decode_dump(native, data, offs, **kw)
    obuf, ooffs = oracle.decode_dump(GE, data, offs, None, nthreads=8)
    if not (gbuf == obuf and np.array_equal(goffs, ooffs)):
        diffs = first_diff(gbuf, goffs, obuf, ooffs, data, offs)
        msg = "\n".join(f"line {i}: {line!r}\n   emu: {g!r}\n   ref: {o!r}" for i, line, g, o in diffs)
        raise AssertionError(f"{len(diffs)}+ lines differ from the oracle:\n{msg}")
    assert info["bound_violations"] == 0  # rows kept <= members that are not reserved keys (the slots reserved for them)
    return info


def test_stage1_bitmaps_per_byte(emu):
    """gf_bits16 flags exactly the quotes, the backslashes, the commas and "some byte below 0x20" — for every byte value in
    every position."""
    rng = np.random.default_rng(7)
    for b in range(256):
        for pos in range(16):
            blk = bytearray(rng.integers(0x61, 0x7B, 16, dtype=np.uint8).tobytes())
            blk[pos] = b
            q, bs, p, c = emu.gelf_bits16(bytes(blk))
            bit = 1 << pos
            assert (q, bs, p, c) == (bit if b == 0x22 else 0, bit if b == 0x5C else 0, bit if b == 0x2C else 0, 1 if b < 0x20 else 0), (b, pos)
    for _ in range(3000):
        blk = rng.integers(0, 256, 16, dtype=np.uint8).tobytes()
        q, bs, p, c = emu.gelf_bits16(blk)
        assert q == sum(1 << k for k in range(16) if blk[k] == 0x22)
        assert bs == sum(1 << k for k in range(16) if blk[k] == 0x5C)
        assert p == sum(1 << k for k in range(16) if blk[k] == 0x2C)
        assert c == int(any(x < 0x20 for x in blk))


def test_goldens_and_appendix(emu, native, oracle):
    lines = [V.G3_LINE.encode()] + [l.encode() for l, _ in V.GELF_ERRORS] + [l.encode() for l, _ in V.GELF_CASES]
    data, offs = oracle.pack(lines)
    check(emu, native, oracle, data, offs)


def test_regular_shapes_stay_on_the_fast_path(emu, native, oracle):
    lines = [
        b'{"host":"h","timestamp":1}', b'  { "host" : "h" , "timestamp" : 1.5 , "_a" : "x" }  ', b'{}', b'{ }',
        b'{"host":"h","version":"1.1","level":3,"short_message":"s","full_message":"f","_x":true,"_y":null,"_z":false,"n":-12}',
        b'{"host":"a\\"b\\\\c\\/d\\b\\f\\n\\r\\t","_u":"\\u00e9\\ud83d\\ude80x","timestamp":1e3}',
        b'{"host":"h","host":"second wins","_k":1,"_k":2,"timestamp":3,"timestamp":4}',
        b'{"hostx":"h","host":"h","timestamq":1,"full_messagf":"x","short_messagE":"y","versioN":"z","leveL":1,"leve":2}',
        b'{"host":"h","_s":"' + b"long " * 60 + b'"}',
    ]
    data, offs = oracle.pack(lines)
    info = check(emu, native, oracle, data, offs)
    assert info["slow"] == 0


def test_irregular_shapes_take_the_exact_parser(emu, native, oracle):
    lines = [
        b'', b' ', b'{', b'}', b'null', b'[1,2,3]', b'"str"', b'12',
        b'{"host":"h",\t"timestamp":1}', b'{"host":"h"\n}', b'{"ho\\u0073t":"h"}', b'{"host":"h","_n":{"a":1}}', b'{"host":"h","_a":[1,2]}',
        b'{"host":"h","_s":"raw\nnewline"}', b'{"host":"h","_s":"raw\ttab"}', b'{"host":"h","_s":"bad \\x escape"}',
        b'{"host":"h","_s":"\\ud800 lone"}', b'{"host":"h","_s":"\\udc00"}', b'{"host":"h","_s":"\\u12g4"}', b'{"host":"h","_s":"unterminated}',
        b'{"host":"h","_n":01}', b'{"host":"h","_n":1.}', b'{"host":"h","_n":-}', b'{"host":"h","_n":tru}', b'{"host":"h",}', b'{"host":"h"} x',
        b'{"host":"h" "x":1}', b'{"host" "h"}', b'{host:"h"}', b'{"host":"h","_s":"\\',
        ('{"host":"h","timestamp":1' + "".join(',"k%d":%d' % (i, i) for i in range(40)) + "}").encode(),
    ]
    data, offs = oracle.pack(lines)
    info = check(emu, native, oracle, data, offs)
    assert info["slow"] == len(lines)


def test_numbers(emu, native, oracle):
    rng = np.random.default_rng(5)
    nums = ["0", "-0", "1", "-1", "18446744073709551615", "18446744073709551616", "-9223372036854775808", "-9223372036854775809", "0.0",
            "1E+2", "1e-2", "123456789012345678901234567890", "1.7976931348623157e308", "1.8e308", "4.9e-324", "1e309", "0e400",
            "1.5e2147483647", "1.5e-2147483649", "18446744073709551615.5", "7.2057594037927933e16"]
    for _ in range(3000):
        nd = int(rng.integers(1, 25))
        digits = str(int(rng.integers(1, 10))) + "".join(str(int(x)) for x in rng.integers(0, 10, nd - 1))
        s = digits
        if rng.random() < 0.6:
            pos = int(rng.integers(1, nd + 1))
            s = digits[:pos] + "." + (digits[pos:] or "0")
        if rng.random() < 0.5:
            s += "e%d" % int(rng.integers(-330, 310))
        if rng.random() < 0.2:
            s = "-" + s
        nums.append(s)
    lines = [('{"host":"h","timestamp":%s,"n":%s}' % (n, n)).encode() for n in nums]
    data, offs = oracle.pack(lines)
    check(emu, native, oracle, data, offs)


def test_generated_and_ragged(emu, native, oracle):
    data, offs = native.generate(native.FMT_GELF, 0x6E1F, 40_000)
    info = check(emu, native, oracle, data, offs)
    assert info["slow"] < 0.03 * 40_000  # malformed lines and the few irregular shapes only
    check(emu, native, oracle, data, offs, tile_bytes=8192)
    data, offs = native.generate(native.FMT_GELF, 77, 20_000, bad_frac=1.0)
    check(emu, native, oracle, data, offs)
    big = '{"host":"h","timestamp":1,"k":"' + "v" * 100_000 + '\\n","z":[' + ",".join(["1"] * 5000) + "]}"
    many = '{"host":"h","timestamp":1' + "".join(',"k%d":%d' % (7919 * i % 3001, i) for i in range(3000)) + "}"
    wide = [('{"host":"h","timestamp":1' + "".join(',"_k%d":%d' % (i, i) for i in range(20)) + "}").encode()] * 100
    data, offs = oracle.pack([big.encode(), many.encode(), V.G3_LINE.encode()] + wide + [V.G3_LINE.encode()] * 150)
    info = check(emu, native, oracle, data, offs)
    assert info["rounds"] > 4  # 100 lines x 20 rows: the staging area cuts the rounds


def test_mutation_fuzz(emu, native, oracle):
    rng = np.random.default_rng(2024)
    data, offs = native.generate(native.FMT_GELF, 9, 20_000, bad_frac=0.0)
    alphabet = b'{}[]":,\\ \n\tu0e-.ntf_'
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
    check(emu, native, oracle, d2, o2)
