"""Parity of the CUDA RFC5424 path (through the C ABI) against the oracle. GPU only."""
import numpy as np
import pytest

import vectors as V
from conftest import assert_parity

pytestmark = pytest.mark.gpu
R5 = 0


@pytest.fixture(scope="module")
def dec(native):
    d = native.BatchDecoder(native.FMT_RFC5424, max_batch_bytes=600 << 20, max_batch_lines=3 << 20, chunk_lines=1 << 18)
    yield d
    d.close()


def test_golden_g1_g2(dec, oracle, native):
    data, offs = oracle.pack([V.G1_LINE.encode(), V.G2_LINE.encode()])
    res = assert_parity(dec, oracle, R5, data, offs)
    assert res.status.tolist() == [0, 0]
    assert res.ts.view(np.uint64).tolist() == [0x41D5708C6268D21C] * 2  # 1438790025.637824, bit-exact
    assert (res.meta >> 8 & 0xFF).tolist() == [2, 2] and (res.meta >> 16 & 0xFF).tolist() == [7, 7]
    span = lambda a, i: bytes(data[a[i, 0]:a[i, 0] + a[i, 1]])
    sp = res.spans5424(offs)  # the compact 32-byte rows decoded like the fg_row5424_* helpers of the C header
    assert span(sp["hostname"], 0) == b"testhostname" and span(sp["appname"], 0) == b"appname"
    assert span(sp["procid"], 0) == b"69" and span(sp["msgid"], 0) == b"42" and span(sp["msg"], 1) == b"test message"
    assert sp["sd"][:, 1].tolist() == [3, 6]  # header + 2 pairs ; 2 headers + 4 pairs
    # software="te\\st sc\\"ript" was unescaped on the device (rfc5424_decoder.rs:105-125): arena records are [u16 length][bytes]
    arena = bytes(res.arena[:int(res.raw.arena_bytes)])
    recs, at = [], 0
    while at < len(arena):
        l = int.from_bytes(arena[at:at + 2], "little")
        recs.append(arena[at + 2:at + 2 + l])
        at += (2 + l + 1) & ~1
    assert recs == [b'te\\st sc"ript'] * 2
    esc = [int(e) for e in res.entries8 if (int(e) >> 62) & 1]
    assert len(esc) == 2


def test_appendix_vectors(dec, oracle, native):
    lines = [l.encode() for l, _ in V.RFC5424_CASES]
    data, offs = oracle.pack(lines)
    res = assert_parity(dec, oracle, R5, data, offs)
    for i, (line, err) in enumerate(V.RFC5424_CASES):
        got = native.error_string(R5, int(res.status[i]))
        assert got == err, (line, got, err)


def test_empty_batch_and_empty_lines(dec, oracle):
    data, offs = oracle.pack([])
    res = dec.decode(data, offs)
    assert res.n == 0
    data, offs = oracle.pack([b"", b"", V.G1_LINE.encode(), b""])
    assert_parity(dec, oracle, R5, data, offs)


def test_ragged_and_long_lines(dec, oracle):
    """lines longer than the shared-memory tile take the direct-from-global path; mixed with short ones."""
    big_msg = b"x" * 300_000
    lines = [V.G1_LINE.encode(), b"<13>1 " + V.TS.encode() + b" h a p m - " + big_msg,
             b"<13>1 " + V.TS.encode() + b" h a p m [id k=\"" + b"v" * 250_000 + b"\\\"\"] tail  ",
             V.G2_LINE.encode(), b"<13>1 " + V.TS.encode() + b" " + b"h" * 70_000 + b" a p m - z"] + [V.G2_LINE.encode()] * 300
    data, offs = oracle.pack(lines)
    assert_parity(dec, oracle, R5, data, offs)
    assert_parity(dec, oracle, R5, data, offs, resident=True)


def test_generated_1m_lines(dec, oracle, native):
    data, offs = native.generate(native.FMT_RFC5424, 5424, 1_000_000, bad_frac=0.005)
    res = assert_parity(dec, oracle, R5, data, offs)
    bad = int((res.status != 0).sum())
    assert 2000 < bad < 9000
    assert dec.kernel_launches() > 0


def test_generated_all_bad_and_resident(dec, oracle, native):
    data, offs = native.generate(native.FMT_RFC5424, 99, 200_000, bad_frac=1.0)
    assert_parity(dec, oracle, R5, data, offs, resident=True)


def test_mutation_fuzz(dec, oracle, native):
    """Random byte edits of valid lines (kept valid UTF-8 by using ASCII edits on ASCII lines)."""
    rng = np.random.default_rng(1234)
    data, offs = native.generate(native.FMT_RFC5424, 7, 100_000, bad_frac=0.0)
    lines = [bytearray(data[offs[i]:offs[i + 1]]) for i in range(len(offs) - 1)]
    alphabet = b' []"\\=<>-1:TZ+.\t'
    out = []
    for ln in lines:
        if any(b >= 0x80 for b in ln):
            out.append(bytes(ln))
            continue
        k = int(rng.integers(1, 4))
        for _ in range(k):
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
    res = assert_parity(dec, oracle, R5, d2, o2)
    assert int((res.status != 0).sum()) > 10_000


def test_pageable_and_pinned_inputs_agree(dec, oracle, native):
    data, offs = native.generate(native.FMT_RFC5424, 11, 300_000)
    r1 = dec.decode(data, offs)
    ts1, meta1 = r1.ts.copy(), r1.meta.copy()
    pb = dec.host_alloc(len(data))
    po = dec.host_alloc(offs.nbytes, dtype=np.int32)
    pb[:] = data
    po[:] = offs
    r2 = dec.decode(pb, po)
    assert np.array_equal(ts1.view(np.uint64), r2.ts.view(np.uint64)) and np.array_equal(meta1, r2.meta)


def test_roundtrip_property_full_msg(dec, native):
    """Size-independent property: for Ok rows, full_msg is the (BOM-stripped, right-trimmed) line and every
    span lies inside its own line; holds at any batch size without the oracle."""
    data, offs = native.generate(native.FMT_RFC5424, 3, 3_000_000, bad_frac=0.005)
    res = dec.decode(data, offs)
    ok = (res.status == 0) & ((res.meta >> 24) & 0x80 == 0)  # compact rows (the few FG_FLAG_WIDE rows carry absolute spans)
    assert int(ok.sum()) > 2_950_000
    lo, hi = offs[:-1][ok], offs[1:][ok]
    sp = res.spans5424(offs)
    for col in (sp["hostname"], sp["appname"], sp["procid"], sp["msgid"], sp["full_msg"]):
        o, l = col[ok, 0], col[ok, 1]
        assert (o >= lo).all() and (l >= 0).all() and (o + l <= hi).all()
    fo, fl = sp["full_msg"][ok, 0], sp["full_msg"][ok, 1]
    assert (fo == lo).all()
    m = sp["msg"][ok]
    has = m[:, 0] >= 0
    assert (m[has, 0] + m[has, 1] <= fo[has] + fl[has]).all()
    assert np.isfinite(res.ts[ok]).all() and (res.ts[ok] > 1.4e9).all() and (res.ts[ok] < 2.1e9).all()


def test_many_resident_passes_are_idempotent(dec, oracle, native):
    """fg_parse_resident_n (what bench.py times): K back-to-back passes leave exactly the result of one pass."""
    data, offs = native.generate(native.FMT_RFC5424, 13, 200_000)
    dec.upload(data, offs)
    dec.parse_resident()
    ms = dec.parse_resident_many(7)
    assert ms > 0
    res = dec.download()
    g, go = dec.dump(res, data, offs)
    r, ro = oracle.decode_dump(0, data, offs)
    assert g == r and np.array_equal(go, ro)


def test_full_size_batch_parity(oracle, native):
    """BASELINE.json configs[1] at full size: 10 M generated lines (bench.py's seed and length shape), every Record compared
    with the oracle (canonical dumps, f64 bits and error strings included)."""
    n = 10_000_000
    data, offs = native.generate(native.FMT_RFC5424, 5424, n, mean_len=169.2, bad_frac=0.005, nthreads=32)
    big = native.BatchDecoder(native.FMT_RFC5424, max_batch_bytes=int(offs[-1]) + (1 << 20), max_batch_lines=n)
    try:
        res = big.decode(data, offs)  # ONE decode of the whole 10 M-line batch ...
        assert res.n == n
        # ... compared with the oracle line by line, in 1 M-line slices only to bound the size of the dump buffers
        step = 1_000_000
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            gbuf, goffs = big.dump(res, data, offs, nthreads=32, lo=lo, hi=hi)
            base = int(offs[lo])
            sub_offs = (offs[lo:hi + 1] - base).astype(np.int32)
            obuf, ooffs = oracle.decode_dump(R5, data[base:int(offs[hi])], sub_offs, nthreads=32)
            assert gbuf == obuf and np.array_equal(goffs, ooffs), f"lines {lo}:{hi} of the 10 M-line decode differ from the oracle"
        assert int((res.status != 0).sum()) > 30_000
    finally:
        big.close()
