"""Device-side line framing + UTF-8 validation + decode (fg_split_decode) against the CPU
restatement of LineSplitter's framing (oracle/pysplit.py) and the decoder oracle. GPU only."""
import numpy as np
import pytest

import vectors as V

pytestmark = pytest.mark.gpu


def check(dec, oracle, fmt, stream: bytes):
    import pysplit
    offs, lines, valid = pysplit.split_lines(stream)
    arr = np.frombuffer(stream or b"\0", dtype=np.uint8).copy()[: len(stream)]
    buf, bo, line_offs, _ = dec.split_dump(arr if len(stream) else np.zeros(0, np.uint8))
    assert np.array_equal(line_offs, offs), (line_offs[:10], offs[:10])
    good = [l for l, v in zip(lines, valid) if v]
    d, o = oracle.pack(good)
    obuf, oo = oracle.decode_dump(fmt, d, o)
    k = 0
    for i, (l, v) in enumerate(zip(lines, valid)):
        got = buf[bo[i]:bo[i + 1]]
        if not v:
            assert got == b"E:Invalid UTF-8 input;out=0", (i, l, got)
        else:
            assert got == obuf[oo[k]:oo[k + 1]], (i, l, got, obuf[oo[k]:oo[k + 1]])
            k += 1


def test_framing_edge_cases(native, oracle):
    dec = native.BatchDecoder(native.FMT_RFC5424, max_batch_bytes=64 << 20, max_batch_lines=1 << 18)
    try:
        g = V.G1_LINE.encode()
        for stream in [b"", b"\n", b"\n\n", g, g + b"\n", g + b"\r\n", g + b"\r", g + b"\r\r\n", b"\r\n" + g, g + b"\n" + g,
                       g + b"\n\n" + g + b"\n", b"\xff\n" + g + b"\n", g + b"\n\xc3", g + b"\n\xc3\n\xa9" + g + b"\n",
                       "<13>1 2015-08-05T15:53:45Z h a p m - café 日本\U0001F680\n".encode(),
                       b"<13>1 x \xed\xa0\x80\n" + g + b"\n\xf4\x90\x80\x80\n\xe0\x80\x80\n\xc0\xaf\n" + g,
                       b"a" * 20000 + b"\n" + g + b"\n" + b"b" * 9000]:
            check(dec, oracle, 0, stream)
    finally:
        dec.close()


def test_generated_stream_all_formats(native, oracle):
    rng = np.random.default_rng(3)
    for fmt in (native.FMT_RFC5424, native.FMT_LTSV, native.FMT_GELF):
        data, offs = native.generate(fmt, 21, 60_000)
        parts = []
        for i in range(len(offs) - 1):
            l = bytes(data[offs[i]:offs[i + 1]])
            if fmt == native.FMT_GELF and b"\n" in l:
                l = l.replace(b"\n", b" ")  # a raw LF would be a line break in a stream
            r = rng.random()
            if r < 0.01:
                l = l[: len(l) // 2] + b"\xfe" + l[len(l) // 2:]          # invalid byte
            elif r < 0.02:
                l = l + b"\xe2\x82"                                          # truncated sequence at the end of the line
            parts.append(l + (b"\r\n" if rng.random() < 0.1 else b"\n"))
        stream = b"".join(parts)
        if fmt == native.FMT_LTSV:
            stream = stream[:-1]  # unterminated last line
        dec = native.BatchDecoder(fmt, max_batch_bytes=len(stream) + (1 << 20), max_batch_lines=1 << 17)
        try:
            check(dec, oracle, fmt, stream)
        finally:
            dec.close()
