"""Parity of the CUDA LTSV path (through the C ABI) against the oracle. GPU only."""
import numpy as np
import pytest

import vectors as V
from conftest import assert_parity

pytestmark = pytest.mark.gpu
LT = 1
SCHEMA = {"counter": "u64", "score": "i64", "mean": "f64", "done": "bool"}
SUFFIX = {"u64": "_u64", "i64": "_i64", "F64": "_f64", "Bool": "_bool"}


@pytest.fixture(scope="module")
def plain(native):
    d = native.BatchDecoder(native.FMT_LTSV, max_batch_bytes=512 << 20, max_batch_lines=2 << 20)
    yield d
    d.close()


@pytest.fixture(scope="module")
def typed(native):
    d = native.BatchDecoder(native.FMT_LTSV, max_batch_bytes=512 << 20, max_batch_lines=2 << 20, ltsv_schema=SCHEMA)
    yield d
    d.close()


@pytest.fixture(scope="module")
def suffixed(native):
    d = native.BatchDecoder(native.FMT_LTSV, max_batch_bytes=512 << 20, max_batch_lines=2 << 20,
                            ltsv_schema=V.LTSV_SCHEMA_G13, ltsv_suffixes=V.LTSV_SUFFIX_G13)
    yield d
    d.close()


def test_goldens_g9_g12(typed, oracle):
    lines = [V.G9_LINE, V.G10_LINE, V.G11_LINE, V.G12_LINE]
    data, offs = oracle.pack([l.encode() for l in lines])
    res = assert_parity(typed, oracle, LT, data, offs, oracle.LtsvConfig(V.LTSV_SCHEMA))
    assert res.status.tolist() == [0, 0, 0, 0]
    want = [1438790025.99, 1438790025.637824, 971211336.3, 1438790025.637824]
    assert res.ts.view(np.uint64).tolist() == np.array(want).view(np.uint64).tolist()
    assert int(res.meta[2] >> 16 & 0xFF) == 3
    assert res.sd[:, 1].tolist() == [3, 3, 7, 3]


def test_goldens_g13_g14_suffixes(suffixed, oracle, native):
    data, offs = oracle.pack([V.G13_LINE.encode()])
    assert_parity(suffixed, oracle, LT, data, offs, oracle.LtsvConfig(V.LTSV_SCHEMA_G13, V.LTSV_SUFFIX_G13))
    d2 = native.BatchDecoder(native.FMT_LTSV, ltsv_schema=V.LTSV_SCHEMA_G14, ltsv_suffixes=V.LTSV_SUFFIX_G14)
    try:
        data, offs = oracle.pack([V.G14_LINE.encode()])
        assert_parity(d2, oracle, LT, data, offs, oracle.LtsvConfig(V.LTSV_SCHEMA_G14, V.LTSV_SUFFIX_G14))
    finally:
        d2.close()


def test_appendix_vectors(plain, typed, oracle, native):
    data, offs = oracle.pack([l.encode() for l, _ in V.LTSV_CASES])
    res = assert_parity(plain, oracle, LT, data, offs)
    for i, (line, err) in enumerate(V.LTSV_CASES):
        assert native.error_string(LT, int(res.status[i])) == err, line
    data, offs = oracle.pack([l.encode() for l, _ in V.LTSV_SCHEMA_CASES])
    res = assert_parity(typed, oracle, LT, data, offs, oracle.LtsvConfig(SCHEMA))
    for i, (line, err) in enumerate(V.LTSV_SCHEMA_CASES):
        assert native.error_string(LT, int(res.status[i])) == err, line


HARD_DECIMALS = [
    "2.2250738585072011e-308", "2.2250738585072012e-308", "2.2250738585072014e-308", "4.9e-324", "2.4703282292062327e-324",
    "2.4703282292062328e-324", "1.7976931348623157e308", "1.7976931348623158e308", "1.7976931348623159e308", "1e309",
    "9007199254740993", "9007199254740992.5", "9007199254740993.0000000000000000000000000000001", "0.1", "0.3", "1e23",
    "8.41e21", "6.0221409e+23", "123456789012345678901234567890", "0.000000000000000000000000000000000000000001",
    "1438790025.6378241", "1438790025.63782412345678901234567890123456789", "3.141592653589793238462643383279502884197",
    "1.00000000000000011102230246251565404236316680908203125", "1.00000000000000011102230246251565404236316680908203124",
    "1.00000000000000011102230246251565404236316680908203126", "0." + "0" * 400 + "1", "1" + "0" * 400, "1" + "0" * 308,
    "0." + "0" * 322 + "25", "0." + "0" * 323 + "25", "0." + "0" * 323 + "24", "5e-324", "2e-324", "3e-324",
    "17976931348623157" + "0" * 292, "17976931348623158" + "0" * 292, "1" * 800, "0." + "9" * 800, "1e-400", "1e400",
    "+.5e1", "-5.E-1", "00000000000000000000001.5", "1e+0000000000000000000000000000002", "1e-00000000000000000000000000000000",
    "72057594037927945", "7.2057594037927933e16", "2.808895523222369e306", "9.5e-27", "6.8985865317742005e122",
]


def test_correctly_rounded_decimals(plain, oracle):
    """Rust f64::from_str is correctly rounded; the device path is Clinger + an exact big-integer path."""
    rng = np.random.default_rng(99)
    vals = list(HARD_DECIMALS)
    for _ in range(20000):
        nd = int(rng.integers(1, 40))
        digits = "".join(str(int(x)) for x in rng.integers(0, 10, nd))
        pos = int(rng.integers(0, nd + 1))
        s = digits[:pos] + "." + digits[pos:] if rng.random() < 0.7 else digits
        if rng.random() < 0.6:
            s += "e%d" % int(rng.integers(-345, 320))
        vals.append(s)
    # halfway cases around random doubles (exact decimal expansions of midpoints, +- one final digit)
    from decimal import Decimal, getcontext
    getcontext().prec = 1200
    for _ in range(3000):
        bits = int(rng.integers(1, 0x7FEFFFFFFFFFFFFF))
        x = np.array([bits], dtype=np.uint64).view(np.float64)[0]
        y = np.nextafter(x, np.inf)
        if not np.isfinite(y):
            continue
        mid = (Decimal(float(x)) + Decimal(float(y))) / 2
        s = format(mid, "e")
        vals.append(s)
        m, e = s.split("e")
        vals.append(m + "1e" + e)
        vals.append(m[:-1] + ("0" if m[-1] != "0" else "1") + "e" + e)
    lines = [("time:" + v + "\thost:h").encode() for v in vals]
    data, offs = oracle.pack(lines)
    res = assert_parity(plain, oracle, LT, data, offs)
    assert int((res.status != 0).sum()) == 0


def test_empty_and_ragged(plain, oracle):
    big = b"time:1\thost:h\tk:" + b"v" * 400_000 + b"\tz:" + b":" * 1000
    many = b"time:1\thost:h" + b"".join(b"\tk%d:v" % i for i in range(5000))
    colons = b"\t".join([b":"] * 3000) + b"\ttime:1\thost:h"
    lines = [b"", b"\t", b":", b":\t:", b"time:1\thost:h", big, many, colons] + [V.G11_LINE.encode()] * 200
    data, offs = oracle.pack(lines)
    assert_parity(plain, oracle, LT, data, offs)
    assert_parity(plain, oracle, LT, data, offs, resident=True)


def test_generated(plain, typed, suffixed, oracle, native):
    data, offs = native.generate(native.FMT_LTSV, 1757, 300_000)
    res = assert_parity(plain, oracle, LT, data, offs)
    assert 500 < int((res.status != 0).sum()) < 3000
    assert_parity(typed, oracle, LT, data, offs, oracle.LtsvConfig(SCHEMA))
    assert_parity(suffixed, oracle, LT, data, offs, oracle.LtsvConfig(V.LTSV_SCHEMA_G13, V.LTSV_SUFFIX_G13), resident=True)


def test_mutation_fuzz(typed, oracle, native):
    rng = np.random.default_rng(4321)
    data, offs = native.generate(native.FMT_LTSV, 8, 60_000, bad_frac=0.0)
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
    assert_parity(typed, oracle, LT, d2, o2, oracle.LtsvConfig(SCHEMA))


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


def test_many_typed_values_and_parked_order(oracle, native):
    """Several values of one schema type, failures at parked and in-place positions: the first failing PART decides."""
    d = native.BatchDecoder(native.FMT_LTSV, ltsv_schema=MANY_TYPED_SCHEMA)
    try:
        data, offs = oracle.pack(MANY_TYPED_LINES * 40)
        assert_parity(d, oracle, LT, data, offs, oracle.LtsvConfig(MANY_TYPED_SCHEMA))
    finally:
        d.close()
