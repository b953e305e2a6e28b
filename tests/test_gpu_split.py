"""Device-side line framing + UTF-8 validation + decode (fg_split_decode) against the CPU
restatement of LineSplitter's framing (oracle/pysplit.py) and the decoder oracle. GPU only."""
import numpy as np
import pytest

import vectors as V

pytestmark = pytest.mark.gpu


def check(dec, oracle, fmt, stream: bytes, framing: int = 0):
    import pysplit
    offs, lines, valid = (pysplit.split_nul if framing else pysplit.split_lines)(stream)
    arr = np.frombuffer(stream or b"\0", dtype=np.uint8).copy()[: len(stream)]
    buf, bo, line_offs, _ = dec.split_dump(arr if len(stream) else np.zeros(0, np.uint8), framing)
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


def test_multi_chunk_stream_and_chunk_boundaries(native, oracle):
    """fg_split_decode pipelines the stream in 64 MiB chunks: lines, multi-byte characters and invalid sequences that
    straddle a chunk boundary must come out exactly as in the single-chunk case."""
    import pysplit
    base = ("<13>1 2015-08-05T15:53:45Z h a p m - " + "x" * 24).encode()      # 61 bytes
    assert len(base) == 61
    line = base + b"ab\n"                                                         # 64 bytes incl. terminator
    n_fill = (64 << 20) // 64
    dec = native.BatchDecoder(native.FMT_RFC5424, max_batch_bytes=200 << 20, max_batch_lines=3 << 20)
    try:
        for shift, tail in [(0, "é"), (1, "é"), (2, "日"), (3, "\U0001F680"), (1, None), (2, b"\xe2\x82"), (3, b"\xf0\x9f")]:
            first = b"<13>1 2015-08-05T15:53:45Z h a p m - " + b"y" * (26 + shift) + b"\n"   # shifts every later line
            if tail is None:
                special = base + b"\xc3\n"            # lead byte right before the '\n': invalid line
            elif isinstance(tail, bytes):
                special = base + tail + b"\n"         # truncated sequence: invalid line
            else:
                special = base + tail.encode() + b"\n"
            # the special line sits so that its non-ASCII bytes fall around the 64 MiB boundary
            k = n_fill - 2
            stream = first + line * k + special + line * 5 + b"<13>1 2015-08-05T15:53:45Z h a p m - end"
            stream = stream + line * (n_fill // 2)    # a third chunk and an unterminated... (ends with '\n' here)
            arr = np.frombuffer(stream, dtype=np.uint8)
            buf, bo, line_offs, _ = dec.split_dump(arr)
            offs, lines, valid = pysplit.split_lines(stream)
            assert np.array_equal(line_offs, offs)
            idx = k + 1
            assert lines[idx].startswith(base)
            got = buf[bo[idx]:bo[idx + 1]]
            if valid[idx]:
                d, o = oracle.pack([lines[idx]])
                ob, oo = oracle.decode_dump(0, d, o)
                assert got == ob
            else:
                assert got == b"E:Invalid UTF-8 input;out=0"
            # every other line is valid
            bad = [j for j in range(len(lines)) if buf[bo[j]:bo[j] + 2] != b"R:"]
            assert bad == ([] if valid[idx] else [idx])
    finally:
        dec.close()


def test_nul_framing(native, oracle):
    """input.framing = "nul" (NulSplitter, nul_splitter.rs:18-40) on the device: records end at a NUL byte; '\\n' and
    '\\r' are ordinary bytes of the record."""
    dec = native.BatchDecoder(native.FMT_RFC5424, max_batch_bytes=64 << 20, max_batch_lines=1 << 18)
    try:
        g = V.G1_LINE.encode()
        for stream in [b"", b"\0", b"\0\0", g, g + b"\0", g + b"\r\n\0", g + b"\0" + g, g + b"\0\0" + g + b"\0", b"\xff\0" + g + b"\0",
                       g + b"\0\xc3", g + b"\n" + g + b"\0", b"a" * 20000 + b"\0" + g + b"\0" + b"b" * 9000]:
            check(dec, oracle, 0, stream, framing=1)
        data, offs = native.generate(native.FMT_RFC5424, 23, 100_000)
        stream = b"".join(bytes(data[offs[i]:offs[i + 1]]) + b"\0" for i in range(100_000))
        check(dec, oracle, 0, stream, framing=1)
    finally:
        dec.close()
