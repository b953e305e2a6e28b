"""N2: decode (RFC5424) + GelfEncoder::encode fused on the device, against the oracle's restatement of both stages
(oracle decoder -> Record -> oracle/encoder.cpp, which is pinned to the reference's own encoder tests).  GPU only."""
import numpy as np
import pytest

import vectors as V

pytestmark = pytest.mark.gpu
R5 = 0


@pytest.fixture(scope="module")
def dec(native):
    d = native.BatchDecoder(native.FMT_RFC5424, max_batch_bytes=400 << 20, max_batch_lines=2 << 20, chunk_lines=1 << 17)
    yield d
    d.close()


def check(dec, oracle, data, offs, extra=None):
    dec.set_gelf_extra(extra or {})
    buf, o, status, _ = dec.decode_encode_gelf(data, offs)
    obuf, oo = oracle.decode_encode_gelf(R5, data, offs, extra or {}, nthreads=16)
    n = len(offs) - 1
    if buf != obuf or not np.array_equal(o, oo):
        for i in range(n):
            a, b = buf[o[i]:o[i + 1]], obuf[oo[i]:oo[i + 1]]
            assert a == b, (i, bytes(data[offs[i]:offs[i + 1]]), a, b)
        raise AssertionError("offsets differ")
    # a line the decoder rejects: empty record + the decoder's status
    d_dump, d_offs = oracle.decode_dump(R5, data, offs)
    for i in range(min(n, 5000)):
        assert (status[i] != 0) == d_dump[d_offs[i]:d_offs[i] + 2].startswith(b"E:")
    return buf, o, status


def test_goldens_and_vectors(dec, oracle):
    lines = [V.G1_LINE.encode(), V.G2_LINE.encode()] + [l.encode() for l, _ in V.RFC5424_CASES]
    data, offs = oracle.pack(lines)
    buf, o, status = check(dec, oracle, data, offs)
    rec = buf[o[1]:o[2]].decode()
    # G2: two elements — the last sd_id wins, keys in byte order, the escaped value is unescaped then JSON-escaped
    assert rec.startswith('{"_key":"value","_key2":"value2","_software":"te\\\\st sc\\"ript","_swVersion":"0.0.1","application_name":"appname"')
    assert '"sd_id":"master@456"' in rec and '"timestamp":1438790025.637824' in rec and rec.endswith('"version":"1.1"}')
    check(dec, oracle, data, offs, {"secret-token": "secret", "host": "overridden", "_key": "extra wins", "a\"b": "c\\d\n"})


def test_generated_and_strings(dec, oracle, native):
    data, offs = native.generate(native.FMT_RFC5424, 5424, 400_000, bad_frac=0.01)
    check(dec, oracle, data, offs)
    check(dec, oracle, data, offs, {"zone": "eu-1", "_aa": "b"})
    tricky = [b"<13>1 " + V.TS.encode() + b' h a p m [i k="v" k="w" a="1"][j k="z"] tab\there "quoted" back\\slash',
              b"<13>1 " + V.TS.encode() + b" h a p m - ",
              b"<13>1 " + V.TS.encode() + b"  a p m - empty host",
              b"<13>1 2015-08-05T15:53:45.123456789+01:30 h a p m - nanos",
              b"<13>1 2015-08-05T15:53:45Z h a p m - integral seconds",
              b"<13>1 " + V.TS.encode() + b' h a p m [id  a="1" ] irregular (slow path) \xc3\xa9',
              b"\xef\xbb\xbf<13>1 " + V.TS.encode() + b" h a p m - bom"]
    d2, o2 = oracle.pack(tricky * 50)
    check(dec, oracle, d2, o2)


def test_output_buffer_regrow(native, oracle):
    d = native.BatchDecoder(native.FMT_RFC5424, max_batch_bytes=1 << 20, max_batch_lines=40_000)
    try:
        # tiny lines: the encoded records are far larger than 2x the input + 200 B per line
        lines = [b"<13>1 2015-08-05T15:53:45Z h a p m [i " + b" ".join(b'k%d="\\"\\"\\"\\""' % k for k in range(30)) + b"] m"] * 2000
        data, offs = oracle.pack(lines)
        check(d, oracle, data, offs)
    finally:
        d.close()


def test_splitter_rfc5424_to_gelf(native, oracle):
    """Config #1 with output.format = "gelf": BatchingLineSplitter sends exactly the bytes LineSplitter + RFC5424Decoder +
    GelfEncoder would, and prints the same stderr line for every rejected line (line_splitter.rs:37-52)."""
    data, offs = native.generate(native.FMT_RFC5424, 21, 10_000, bad_frac=0.02)
    lines = [bytes(data[offs[i]:offs[i + 1]]) for i in range(10_000)]
    text = b"\n".join(lines) + b"\n"
    d = native.BatchDecoder(native.FMT_RFC5424, max_batch_bytes=64 << 20, max_batch_lines=1 << 16)
    try:
        records, err = native.splitter_run_gelf(d, text, {"env": "prod"}, max_lines=3000, max_bytes=1 << 20)
    finally:
        d.close()
    ebuf, eo = oracle.decode_encode_gelf(R5, data, offs, {"env": "prod"})
    dbuf, do = oracle.decode_dump(R5, data, offs)
    want_records, want_err = [], []
    for i, l in enumerate(lines):
        dump = dbuf[do[i]:do[i + 1]]
        if dump.startswith(b"E:"):
            want_err.append(dump[2:dump.index(b";out=")] + b": [" + l.decode().strip().encode() + b"]")
        else:
            want_records.append(ebuf[eo[i]:eo[i + 1]])
    assert records.split(b"\n")[:-1] == want_records
    assert err.split(b"\n")[:-1] == want_err
