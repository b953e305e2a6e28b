"""Parity of the CUDA GELF path (through the C ABI) against the oracle. GPU only."""
import json

import numpy as np
import pytest

import vectors as V
from conftest import assert_parity

pytestmark = pytest.mark.gpu
GE = 2


@pytest.fixture(scope="module")
def dec(native):
    d = native.BatchDecoder(native.FMT_GELF, max_batch_bytes=768 << 20, max_batch_lines=2 << 20)
    yield d
    d.close()


def test_golden_g3(dec, oracle):
    data, offs = oracle.pack([V.G3_LINE.encode()])
    res = assert_parity(dec, oracle, GE, data, offs)
    assert res.status.tolist() == [0]
    assert res.ts.view(np.uint64).tolist() == np.array([1385053862.3072]).view(np.uint64).tolist()
    assert int(res.meta[0] >> 16 & 0xFF) == 1 and int(res.meta[0] >> 8 & 0xFF) == 0xFF
    assert res.sd[0, 1] == 3
    names = [bytes(data[o:o + l]) for o, l in res.entry_name[res.sd[0, 0]:res.sd[0, 0] + 3]]
    assert names == [b"_some_env_var", b"_some_info", b"_user_id"]  # sorted-key order
    assert int(res.entry_val[res.sd[0, 0] + 2]) == 9001 and int(res.entry_meta[res.sd[0, 0] + 2] & 7) == 4


def test_golden_g4_g8_errors(dec, oracle, native):
    data, offs = oracle.pack([l.encode() for l, _ in V.GELF_ERRORS])
    res = assert_parity(dec, oracle, GE, data, offs)
    for i, (line, err) in enumerate(V.GELF_ERRORS):
        assert native.error_string(GE, int(res.status[i])) == err, line


def test_appendix_vectors(dec, oracle, native):
    data, offs = oracle.pack([l.encode() for l, _ in V.GELF_CASES])
    res = assert_parity(dec, oracle, GE, data, offs)
    for i, (line, err) in enumerate(V.GELF_CASES):
        assert native.error_string(GE, int(res.status[i])) == err, line


def test_numbers(dec, oracle):
    """serde_json 0.8 number typing and its (not correctly rounded) float assembly."""
    rng = np.random.default_rng(5)
    nums = ["0", "-0", "1", "-1", "18446744073709551615", "18446744073709551616", "184467440737095516150", "-9223372036854775808",
            "-9223372036854775809", "0.0", "-0.0", "1e0", "1E+2", "1e-2", "123456789012345678901234567890", "1.7976931348623157e308",
            "1.8e308", "4.9e-324", "1e-320", "0.1", "0.3", "2.5", "1385053862.3072", "9007199254740993", "9007199254740993.0",
            "123456789.123456789123456789", "0.000000000000000000000000000001", "1e22", "1e23", "5e-324", "2e-324", "1e308", "1e309",
            "0e400", "-0e-400", "1.0e+00000000000000000000001", "1.5e2147483647", "1.5e-2147483649", "0.0e2147483648", "18446744073709551615.5",
            "99999999999999999999.999e-5", "1234567890123456789012e-30", "7.2057594037927933e16"]
    for _ in range(20000):
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
    assert_parity(dec, oracle, GE, data, offs)


def test_empty_ragged_and_deep(dec, oracle):
    big = '{"host":"h","timestamp":1,"k":"' + "v" * 400_000 + '\\n","z":[' + ",".join(["1"] * 5000) + "]}"
    many = '{"host":"h","timestamp":1' + "".join(',"k%d":%d' % (7919 * i % 3001, i) for i in range(3000)) + "}"
    lines = [b"", b" ", b"{", b"}", b"{}", b"null", big.encode(), many.encode(), V.G3_LINE.encode()] + [V.G3_LINE.encode()] * 150
    data, offs = oracle.pack(lines)
    assert_parity(dec, oracle, GE, data, offs)
    assert_parity(dec, oracle, GE, data, offs, resident=True)


def test_generated(dec, oracle, native):
    data, offs = native.generate(native.FMT_GELF, 0x6E1F, 300_000)
    res = assert_parity(dec, oracle, GE, data, offs)
    assert 500 < int((res.status != 0).sum()) < 3500
    data, offs = native.generate(native.FMT_GELF, 77, 100_000, bad_frac=1.0)
    assert_parity(dec, oracle, GE, data, offs, resident=True)


def test_python_json_cross_check(dec, oracle, native):
    """Independent check of the string-unescape path: Python's json agrees on every decoded string field."""
    data, offs = native.generate(native.FMT_GELF, 31, 20_000, bad_frac=0.0)
    res = dec.decode(data, offs)
    buf, o = dec.dump(res, data, offs)
    for i in range(0, 20_000, 7):
        obj = json.loads(bytes(data[offs[i]:offs[i + 1]]).decode())
        d = buf[o[i]:o[i + 1]]
        host = obj["host"].encode()
        assert b";host=%d:%s;" % (len(host), host) in d
        fm = obj["full_message"].encode()
        assert b";full=%d:%s;" % (len(fm), fm) in d


def test_mutation_fuzz(dec, oracle, native):
    rng = np.random.default_rng(2024)
    data, offs = native.generate(native.FMT_GELF, 9, 60_000, bad_frac=0.0)
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
    res = assert_parity(dec, oracle, GE, d2, o2)
    assert int((res.status != 0).sum()) > 3000
