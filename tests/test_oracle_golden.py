"""Pin the CPU oracle against every known-answer test the reference holds for the Decoder path
 and against the derived behaviour vectors of Appendix A."""
import struct

import pytest

import vectors as V

R5, LT, GE = 0, 1, 2


def f64_hex(x):
    return struct.pack(">d", x).hex()


def test_g1_rfc5424(oracle):  # rfc5424_decoder.rs:245-278
    s = oracle.decode_debug(R5, V.G1_LINE)
    assert s.startswith("Ok(Record { ts: 1438790025.637824, hostname: \"testhostname\", facility: Some(2), severity: Some(7), "
                        "appname: Some(\"appname\"), procid: Some(\"69\"), msgid: Some(\"42\"), msg: Some(\"test message\")")
    assert 'sd: Some([StructuredData { sd_id: Some("origin@123"), pairs: [("_software", String("te\\\\st sc\\"ript")), ("_swVersion", String("0.0.1"))] }])' in s
    data, offs = oracle.pack([V.G1_LINE.encode()])
    buf, _ = oracle.decode_dump(R5, data, offs)
    assert b"ts=41d5708c6268d21c;" in buf  # exact f64 bits of 1438790025.637824 (two-rounding recipe)
    assert f64_hex(1438790025.637824) == "41d5708c6268d21c"


def test_g2_rfc5424_multiple_sd(oracle):  # rfc5424_decoder.rs:281-314
    s = oracle.decode_debug(R5, V.G2_LINE)
    assert s.count("StructuredData {") == 2
    assert 'StructuredData { sd_id: Some("master@456"), pairs: [("_key", String("value")), ("_key2", String("value2"))] }' in s
    assert 'msg: Some("test message")' in s


def test_g3_gelf(oracle):  # gelf_decoder.rs:134-170
    s = oracle.decode_debug(GE, V.G3_LINE)
    assert s.startswith('Ok(Record { ts: 1385053862.3072, hostname: "example.org", facility: None, severity: Some(1), ')
    assert 'msg: Some("A short message that helps you identify what is going on")' in s
    assert 'full_msg: Some("Backtrace here\\n\\nmore stuff")' in s
    assert 'pairs: [("_some_env_var", String("bar")), ("_some_info", String("foo")), ("_user_id", U64(9001))]' in s


@pytest.mark.parametrize("line,err", V.GELF_ERRORS)
def test_g4_g8_gelf_errors(oracle, line, err):  # gelf_decoder.rs:173-205
    assert oracle.decode_debug(GE, line) == f'Err("{err}")'


def test_g9_g12_ltsv_timestamps(oracle):  # ltsv_decoder.rs:369-393,475-487
    cfg = oracle.LtsvConfig(V.LTSV_SCHEMA)
    for line, ts in [(V.G9_LINE, 1438790025.99), (V.G10_LINE, 1438790025.637824), (V.G12_LINE, 1438790025.637824)]:
        data, offs = oracle.pack([line.encode()])
        buf, _ = oracle.decode_dump(LT, data, offs, cfg)
        assert f"ts={f64_hex(ts)};".encode() in buf, (line, buf)


def test_g11_ltsv_typed(oracle):  # ltsv_decoder.rs:396-472
    cfg = oracle.LtsvConfig(V.LTSV_SCHEMA)
    s = oracle.decode_debug(LT, V.G11_LINE, cfg)
    assert s.startswith('Ok(Record { ts: 971211336.3, hostname: "testhostname", facility: None, severity: Some(3), ')
    assert 'msg: Some("this is a test")' in s
    for frag in ['("_name1", String("value1"))', '("_name 2", String(" value 2"))', '("_n3", String("v3"))',
                 '("_counter", U64(42))', '("_score", I64(-1))', '("_mean", F64(0.42))', '("_done", Bool(true))']:
        assert frag in s
    data, offs = oracle.pack([V.G11_LINE.encode()])
    buf, _ = oracle.decode_dump(LT, data, offs, cfg)
    assert b"ts=41ccf1c124266666;" in buf


def test_g13_g14_ltsv_suffixes(oracle):  # ltsv_decoder.rs:270-366
    s = oracle.decode_debug(LT, V.G13_LINE, oracle.LtsvConfig(V.LTSV_SCHEMA_G13, V.LTSV_SUFFIX_G13))
    for frag in ['("_counter_u64", U64(42))', '("_score_i64", I64(-1))', '("_mean_f64", F64(0.42))', '("_done_bool", Bool(true))']:
        assert frag in s
    s = oracle.decode_debug(LT, V.G14_LINE, oracle.LtsvConfig(V.LTSV_SCHEMA_G14, V.LTSV_SUFFIX_G14))
    for frag in ['("_counter_u64", U64(42))', '("_score_i64", I64(-1))', '("_mean_f64", F64(0.42))', '("_done_bool", Bool(true))']:
        assert frag in s
    assert "_u64_u64" not in s


def test_g15_record_display(oracle):  # record.rs:94-132
    assert oracle.g15(0) == '[someid a="a string" b="123456" c="true" d="123.456" e="-123456" f]'
    assert oracle.g15(1) == ('StructuredData { sd_id: Some("someid"), pairs: [("a", String("a string")), ("b", U64(123456)), '
                             '("c", Bool(true)), ("d", F64(123.456)), ("e", I64(-123456)), ("_f", Null)] }')
    assert oracle.g15(2) == ('Record { ts: 123.456, hostname: "hostname", facility: Some(3), severity: Some(8), appname: Some("app"), '
                             'procid: Some("123"), msgid: None, msg: Some("msg"), full_msg: None, sd: None }')


def _check(oracle, fmt, cases, cfg=None):
    for line, err in cases:
        s = oracle.decode_debug(fmt, line, cfg)
        if err is None:
            assert s.startswith("Ok("), (line, s)
        else:
            assert s == f'Err("{err}")', (line, s)


def test_appendix_rfc5424(oracle):
    _check(oracle, R5, V.RFC5424_CASES)
    s = oracle.decode_debug(R5, V.RFC5424_CASES[0][0])  # E1: "-" stays Some("-")
    assert 'facility: Some(4), severity: Some(2), appname: Some("su"), procid: Some("-"), msgid: Some("ID47"), msg: Some("\'su root\' failed")' in s
    assert s.endswith("sd: None })")
    s = oracle.decode_debug(R5, V.H + '[id a="\\\\"] m')  # E26
    assert '("_a", String("\\\\"))' in s
    s = oracle.decode_debug(R5, V.H + "- 　hello ")  # E27
    assert 'msg: Some("hello")' in s
    s = oracle.decode_debug(R5, "﻿<13>1 " + V.TS + " h a p m - x")  # E12: BOM not part of full_msg
    assert 'full_msg: Some("<13>1 ' in s


def test_appendix_ltsv(oracle):
    _check(oracle, LT, V.LTSV_CASES)
    _check(oracle, LT, V.LTSV_SCHEMA_CASES, oracle.LtsvConfig(V.LTSV_SCHEMA))
    data, offs = oracle.pack([b"time:1\thost:h\tfoo", b"foo\tlevel:9\tbar", b""])
    buf, o = oracle.decode_dump(LT, data, offs)
    assert buf[o[0]:o[1]].endswith(b";out=1;28:Missing value for name 'foo'")
    assert buf[o[1]:o[2]] == b"E:Severity level should be <= 7;out=1;28:Missing value for name 'foo'"
    assert buf[o[2]:o[3]] == b"E:Missing timestamp;out=1;25:Missing value for name ''"
    s = oracle.decode_debug(LT, "time:1\thost:a\thost:b")
    assert 'hostname: "b"' in s


def test_appendix_gelf(oracle):
    _check(oracle, GE, V.GELF_CASES)
    s = oracle.decode_debug(GE, '{"host":"h","x":null,"y":true,"z":-3,"w":2.5,"timestamp":0}')  # J7 sorted-key order
    assert 'pairs: [("_w", F64(2.5)), ("_x", Null), ("_y", Bool(true)), ("_z", I64(-3))]' in s
    s = oracle.decode_debug(GE, '{"host":"h","_x":1,"x":2,"timestamp":0}')  # J8
    assert 'pairs: [("_x", U64(1)), ("_x", U64(2))]' in s
    s = oracle.decode_debug(GE, '{"host":"h","timestamp":0,"k":"a\nb\\\nc"}')
    assert '("_k", String("a\\nb\\\\nc"))' in s


def test_timestamp_two_roundings_differ_from_decimal_parse(oracle):
    """SURVEY.md §7 hard part 1: fl(fl(nanos)/1e9) != correctly rounded decimal in ~27% of stamps."""
    import random
    from datetime import datetime, timezone
    rnd = random.Random(7)
    differ = 0
    n = 2000
    for _ in range(n):
        secs = rnd.randrange(1420070400, 2051222400)
        us = rnd.randrange(1_000_000)
        dt = datetime.fromtimestamp(secs, tz=timezone.utc)
        s = dt.strftime("%Y-%m-%dT%H:%M:%S") + f".{us:06d}Z"
        got = oracle.rfc3339(s.encode())
        want = float(secs * 1_000_000_000 + us * 1000) / 1e9  # Python int->float is RNE like Rust's `as f64`
        assert got == want, s
        differ += got != float(f"{secs}.{us:06d}")
    assert 0.15 < differ / n < 0.40


def test_rust_f64_grammar(oracle):
    ok = {"1": 1.0, "+1.5": 1.5, "-.5": -0.5, "5.": 5.0, "1e3": 1000.0, "1E-2": 0.01, "inf": float("inf"),
          "-Infinity": float("-inf"), "1438790025.99": 1438790025.99, "9007199254740993": 9007199254740992.0,
          "0.000000000000000000000000000000000000000000000001": 1e-48, "1e400": float("inf"), "4.9e-324": 5e-324}
    for s, v in ok.items():
        assert oracle.parse_f64(s.encode()) == v, s
    for s in ["", "+", "-", ".", "e5", "1e", "1e+", "0x10", "1_0", " 1", "1 ", "infinit", "nane", "1.5.2", "--1"]:
        assert oracle.parse_f64(s.encode()) is None, s
    assert oracle.parse_f64(b"NaN") != oracle.parse_f64(b"NaN")


def test_golden_fixture_file_matches_vectors_module():
    """tests/golden/reference_vectors.json (the reference's own test inputs, with file:line) and tests/vectors.py agree."""
    import json
    from pathlib import Path
    g = json.loads((Path(__file__).parent / "golden" / "reference_vectors.json").read_text())
    assert g["G1"]["line"] == V.G1_LINE and g["G2"]["line"] == V.G2_LINE and g["G3"]["line"] == V.G3_LINE
    assert g["G9"]["line"] == V.G9_LINE and g["G11"]["line"] == V.G11_LINE and g["G14"]["line"] == V.G14_LINE
    assert [c["err"] for c in g["G4-G8"]["cases"]] == [e for _, e in V.GELF_ERRORS]
    assert f64_hex(g["G1"]["expect"]["ts"]) == g["G1"]["expect"]["ts_bits"]
    assert f64_hex(g["G11"]["expect"]["ts"]) == g["G11"]["expect"]["ts_bits"]


# --- encoder/gelf_encoder.rs: the reference's own encoder tests pin the oracle's GelfEncoder restatement (N2) -------
GELF_ENC_EXPECTED = [
    # gelf_encoder.rs:125
    r'{"_some_info":"foo","application_name":"appname","full_message":"Backtrace here\n\nmore stuff","host":"example.org","level":1,"process_id":"44","sd_id":"someid","secret-token":"secret","short_message":"A short message that helps you identify what is going on","timestamp":1385053862.3072,"version":"1.1"}',
    # gelf_encoder.rs:152
    r'{"host":"unknown","level":1,"short_message":"A short message that helps you identify what is going on","timestamp":1385053862.3072,"version":"1.1"}',
    # gelf_encoder.rs:175
    r'{"a_key":"bar","host":"unknown","level":1,"short_message":"A short message that helps you identify what is going on","timestamp":1385053862.3072,"version":"1.1"}',
    # gelf_encoder.rs:215
    r'{"_some_info":"foo","application_name":"appname","full_message":"Backtrace here\n\nmore stuff","host":"example.org","info":123.456,"level":1,"process_id":"44","sd_id":"someid2","secret-token":"secret","short_message":"A short message that helps you identify what is going on","timestamp":1385053862.3072,"version":"1.1"}',
]


def test_gelf_encoder_reference_tests(oracle):
    for k, want in enumerate(GELF_ENC_EXPECTED):
        assert oracle.gelf_encoder_test(k) == want


def test_f64_writer_roundtrips_and_matches_device_logic(oracle):
    """dtoa (Grisu2 + prettify) is restated twice — oracle/encoder.cpp (128-bit products, powers of ten recomputed with
    big-number arithmetic) and flowgger_b200/csrc/fg_dtoa.cuh (device code, generated table) — and both must agree on
    every value, round-trip exactly, and print the shapes serde_json 0.8 prints."""
    import ctypes as C
    import random
    import struct
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent / "emu"))
    import emu
    L = emu.lib()
    L.emu_json_f64.argtypes = [C.c_double, C.c_char_p]

    def dev(v):
        b = C.create_string_buffer(40)
        n = L.emu_json_f64(v, b)
        return b.raw[:n].decode()

    fixed = {1385053862.3072: "1385053862.3072", 123.456: "123.456", 1438790025.0: "1438790025.0", 0.0: "0.0", 1e21: "1e21",
             1e20: "100000000000000000000.0", 1e-7: "1e-7", 0.000001: "0.000001", 0.1: "0.1", 5e-324: "5e-324",
             1.7976931348623157e308: "1.7976931348623157e308", -2.5: "-2.5", 1438790025.637824: "1438790025.637824"}
    for v, want in fixed.items():
        assert oracle.format_f64(v) == want and dev(v) == want, (v, oracle.format_f64(v), dev(v))
    assert oracle.format_f64(float("nan")) == "null" and dev(float("inf")) == "null" and dev(-0.0) == "-0.0"
    rnd = random.Random(7)
    for i in range(60_000):
        if i % 3 == 0:
            v = struct.unpack("<d", struct.pack("<Q", rnd.getrandbits(64)))[0]
            if v != v or abs(v) == float("inf"):
                continue
        elif i % 3 == 1:
            v = round(rnd.uniform(1.4e9, 2.1e9), rnd.choice([0, 3, 6, 9]))
        else:
            v = rnd.uniform(-1e6, 1e6)
        a, b = oracle.format_f64(v), dev(v)
        assert a == b and float(a) == v, (v, a, b)


# ---- RFC3164 ----------------------------------------------------------------------------------
R3 = 3


def _utc_seconds(y, mo, d, h, mi, s):
    import calendar
    return float(calendar.timegm((y, mo, d, h, mi, s)))


@pytest.mark.parametrize("year", [2026, 2021, 2024])
def test_g16_g26_rfc3164(oracle, year):  # rfc3164_decoder.rs:218-425, the reference's eleven tests
    cfg = oracle.Rfc3164Config(year)
    for name, ref, line, exp in V.RFC3164_GOLDEN:
        s = oracle.decode_debug(R3, line, cfg)
        if exp is None:
            assert s.startswith("Err("), (name, s)
            continue
        y, mo, d, h, mi, sec = exp["date"]
        ts = _utc_seconds(y if y is not None else year, mo, d, h, mi, sec)  # ts_from_date_time / ts_from_partial_date_time
        fac = "None" if exp["fac"] is None else f"Some({exp['fac']})"
        sev = "None" if exp["sev"] is None else f"Some({exp['sev']})"
        assert s.startswith(f'Ok(Record {{ ts: {ts!r}, hostname: "{exp["host"]}", facility: {fac}, severity: {sev}, '
                            "appname: None, procid: None, msgid: None, "), (name, s)
        assert s.endswith("sd: None })"), (name, s)
        data, offs = oracle.pack([line.encode()])
        buf, _ = oracle.decode_dump(R3, data, offs, cfg)
        if exp["msg"] is not None:
            m = exp["msg"].encode()
            assert b";msg=%d:%s;" % (len(m), m) in buf, (name, buf)
        full = (exp["full"] if exp["full"] is not None else line).encode()
        assert b";full=%d:%s;" % (len(full), full) in buf, (name, buf)
        assert f"ts={f64_hex(ts)};".encode() in buf


def test_rfc3164_derived_cases(oracle):
    cfg = oracle.Rfc3164Config(2026)
    for line, err in V.RFC3164_CASES:
        s = oracle.decode_debug(R3, line, cfg)
        if err is None:
            assert s.startswith("Ok("), (line, s)
        else:
            assert s == f'Err("{err}")', (line, s)
    # spot values: zone arithmetic and the re-joined message
    def rec(line):
        return oracle.decode_debug(R3, line, cfg)
    assert 'ts: 1596726924.0, hostname: "h"' in rec("2020 Aug 6 11:15:24 America/New_York h m")      # EDT, UTC-4
    assert 'ts: 1596730524.0, hostname: "h"' in rec("2020 Aug 6 11:15:24 Etc/GMT+5 h m")             # POSIX sign: UTC-5
    assert 'ts: 1596712524.0, hostname: "utc"' in rec("2020 Aug 6 11:15:24 utc h m")
    assert 'ts: 1615707000.0' in rec("2021 Mar 14 02:30:00 America/New_York h m")                    # 02:30 EST (UTC-5) = 07:30Z
    assert 'ts: 1636263000.0' in rec("2021 Nov 7 01:30:00 America/New_York h m")                     # the first 01:30 (EDT) = 05:30Z
    assert 'ts: 2540282400.0' in rec("2050 Jul 1 12:00:00 Europe/Paris h m")                         # CEST from the footer rule
    assert 'msg: Some("m n o")' in rec("Aug\t6 11:15:24 h　m  n\r\no ")
    assert 'msg: Some("a b")' in rec("Aug 6 11:15:24 h a  b")
    assert 'hostname: "a b"' in rec("a b: 2020 Aug 6 11:15:24: m: n: o ") and 'msg: Some("m: n: o ")' in rec("a b: 2020 Aug 6 11:15:24: m: n: o ")
    assert 'facility: Some(1), severity: Some(5)' in rec("<<13>Aug 6 11:15:24 h m")


def test_tz_tables_match_zoneinfo(oracle):
    """oracle/tzread.py (TZif reader + POSIX footer expansion + the local-time rule of oracle/rfc3164.cpp) against the
    standard library's zoneinfo (fold=0) on random local times and around transitions, 1901-2400."""
    import random
    from datetime import datetime, timedelta
    from zoneinfo import ZoneInfo
    import tzread
    zones = tzread.load_zones()
    assert len(zones) > 300 and "UTC" in zones and "America/Sao_Paulo" in zones
    rnd = random.Random(3)
    lo, hi = -2208988800 + 86400 * 400, 13569465600
    for name in rnd.sample(sorted(zones), 60) + ["America/New_York", "Europe/Dublin", "Australia/Lord_Howe", "Africa/Casablanca"]:
        zi, z = ZoneInfo(name), zones[name]
        probes = [rnd.randrange(lo, hi) for _ in range(60)]
        tr, of = z
        for k in rnd.sample(range(len(tr)), min(len(tr), 12)):
            if lo < tr[k] < hi:
                probes += [tr[k] + o + d for o in (of[k], of[k + 1]) for d in (-3601, -1, 0, 1, 1800, 3599, 3601)]
        for local in probes:
            want = (datetime(1970, 1, 1) + timedelta(seconds=local)).replace(tzinfo=zi).utcoffset().total_seconds()
            assert tzread.offset_at_local(z, local) == want, (name, local)


def test_rfc3164_three_way(oracle, native):
    """C++ oracle (oracle/rfc3164.cpp + the TZif tables of tzread.py) against the independent Python restatement
    (oracle/pyrfc3164.py: regular expressions, datetime, zoneinfo) on the vectors and on generated lines, dump for dump."""
    import pyrfc3164
    year = 2026
    cfg = oracle.Rfc3164Config(year)
    lines = [l.encode() for _, _, l, _ in V.RFC3164_GOLDEN] + [l.encode() for l, _ in V.RFC3164_CASES]
    data, offs = native.generate(native.FMT_RFC3164, 2, 60_000, bad_frac=0.05)
    lines += [bytes(data[offs[i]:offs[i + 1]]) for i in range(len(offs) - 1)]
    import tzread
    for k, nm in enumerate(sorted(tzread.load_zones())):       # every zone, across its history and into the footer rules
        lines.append(f"{1900 + (k * 7) % 300} {'Mar Apr Oct Nov'.split()[k % 4]} {1 + k % 28} 0{k % 10}:30:00 {nm} h m".encode())
    d, o = oracle.pack(lines)
    buf, bo = oracle.decode_dump(R3, d, o, cfg)
    skipped = 0
    for i, l in enumerate(lines):
        r = pyrfc3164.decode(l.decode(), year)
        if r is pyrfc3164.UNSUPPORTED:
            skipped += 1
            continue
        assert pyrfc3164.dump(r) == buf[bo[i]:bo[i + 1]], (l, pyrfc3164.dump(r), buf[bo[i]:bo[i + 1]])
    assert skipped <= 4                                         # the vectors with year 0000 / negative years
