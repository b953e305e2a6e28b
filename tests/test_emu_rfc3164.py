"""CPU check of the RFC3164 DEVICE LOGIC: the product's device source (fg_rfc3164.cuh) and zone-table
builder (fg_tz.cu) compiled with g++ (tests/emu) and replayed CTA by CTA the way parse3164_kernel runs them, pushed through
the product's host materialiser and compared with the oracle.  parse3164_kernel has no warp-level interplay (one thread
parses one line on its own), so this replay exercises all of its logic; the `-m gpu` tests repeat it on the device."""
import numpy as np
import pytest

import vectors as V
from conftest import first_diff

R3 = 3
YEAR = 2026


@pytest.fixture(scope="module")
def emu():
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent / "emu"))
    import emu as E
    E.build()
    return E


def check(emu, native, oracle, data, offs, year=YEAR, **kw):
    """both walkers of fg_rfc3164.cuh — r3164_parse_lockstep (what the kernel is built with) and r3164_parse_line — against
    the oracle"""
    obuf, ooffs = oracle.decode_dump(R3, data, offs, oracle.Rfc3164Config(year), nthreads=8)
    info = None
    for lockstep in (False, True):
        gbuf, goffs, inf = emu.r3164_decode_dump(native, data, offs, year, lockstep=lockstep, **kw)
        if not (gbuf == obuf and np.array_equal(goffs, ooffs)):
            diffs = first_diff(gbuf, goffs, obuf, ooffs, data, offs)
            msg = "\n".join(f"line {i}: {line!r}\n   emu: {g!r}\n   ref: {o!r}" for i, line, g, o in diffs)
            raise AssertionError(f"lockstep={lockstep}: {len(diffs)}+ lines differ from the oracle:\n{msg}")
        assert info is None or info == inf
        info = inf
    return info


def test_goldens_and_derived_cases(emu, native, oracle):
    lines = [l.encode() for _, _, l, _ in V.RFC3164_GOLDEN] + [l.encode() for l, _ in V.RFC3164_CASES]
    data, offs = oracle.pack(lines)
    for year in (2026, 2024, 2021):
        check(emu, native, oracle, data, offs, year=year)
    check(emu, native, oracle, data, offs, tile_bytes=512)     # most lines straight from the input buffer
    check(emu, native, oracle, data, offs, year=999)           # a year that does not print as four digits: no year-less date parses
    check(emu, native, oracle, data, offs, year=12345)


def test_error_strings(emu, native, oracle):
    """every RFC3164 status renders the reference's text (fg_error_string), checked through the dump of the derived cases"""
    data, offs = oracle.pack([l.encode() for l, _ in V.RFC3164_CASES])
    gbuf, goffs, _ = emu.r3164_decode_dump(native, data, offs, YEAR)
    for k, (line, err) in enumerate(V.RFC3164_CASES):
        d = gbuf[goffs[k]:goffs[k + 1]]
        if err is None:
            assert d.startswith(b"R:"), (line, d)
        else:
            assert d == b"E:" + err.encode() + b";out=0", (line, d)


def test_generated(emu, native, oracle):
    data, offs = native.generate(native.FMT_RFC3164, 3164, 200_000, bad_frac=0.02)
    info = check(emu, native, oracle, data, offs)
    assert info["arena_bytes"] > 100_000 and info["from_tile"] > 190_000
    data, offs = native.generate(native.FMT_RFC3164, 31, 30_000, bad_frac=1.0)
    check(emu, native, oracle, data, offs, year=2024)
    data, offs = native.generate(native.FMT_RFC3164, 64, 30_000, mean_len=600.0)
    check(emu, native, oracle, data, offs, tile_bytes=40960)


def test_arena_regrow_and_small_tiles(emu, native, oracle):
    data, offs = native.generate(native.FMT_RFC3164, 5, 20_000, bad_frac=0.01)
    info = check(emu, native, oracle, data, offs, arena_cap=64)
    assert info["redo"] == 1                                   # the bump allocator ran past the capacity once: regrow + redo
    info = check(emu, native, oracle, data, offs, tile_bytes=1024)
    assert info["from_global"] > 10_000
    long_line = b"<13>Aug  6 11:15:24 host tag: " + b"x y  " * 20_000     # 100 KB message, re-joined
    data, offs = oracle.pack([long_line, b"<13>Aug  6 11:15:24 host tag: short", long_line[:70_000]])
    info = check(emu, native, oracle, data, offs)
    assert info["from_global"] == 3 and info["arena_bytes"] > 80_000


def test_mutation_fuzz(emu, native, oracle):
    rng = np.random.default_rng(3164)
    data, offs = native.generate(native.FMT_RFC3164, 9, 40_000, bad_frac=0.0)
    alphabet = b" :<>+-0123456789\tAugJanUTC/_\r"
    out = []
    for i in range(len(offs) - 1):
        ln = bytearray(data[offs[i]:offs[i + 1]])
        if any(b >= 0x80 for b in ln):
            out.append(bytes(ln))          # keep the UTF-8 lines valid
            continue
        head = min(len(ln), 48)            # the date, zone and hostname live in the first bytes
        for _ in range(int(rng.integers(1, 4))):
            k = int(rng.integers(0, min(head, len(ln)) if rng.random() < 0.8 else len(ln)))
            op = int(rng.integers(0, 3))
            c = alphabet[int(rng.integers(0, len(alphabet)))]
            if op == 0:
                ln[k] = c
            elif op == 1:
                ln.insert(k, c)
            elif len(ln) > 1:
                del ln[k]
        out.append(bytes(ln))
    d2, o2 = oracle.pack(out)
    check(emu, native, oracle, d2, o2)
    # every zone identifier, and near misses of them, in the token position where a zone is looked up
    import tzread
    names = sorted(tzread.load_zones())
    lines = []
    for k, nm in enumerate(names):
        lines.append(f"<{k % 192}>{1990 + k % 60} Mar {1 + k % 28} 0{k % 10}:30:00 {nm} host tag: m {k}".encode())
        lines.append(f"Oct 31 01:30:00 {nm[:-1]} host m".encode())
        lines.append(f"Oct 31 01:30:00 {nm}x host m".encode())
        lines.append(f"h{k}: 2031 Nov 2 01:30:00 {nm.swapcase()}: m".encode())
    d3, o3 = oracle.pack(lines)
    check(emu, native, oracle, d3, o3)


def test_split_mode_terminators(emu, native, oracle):
    lines = [V.RFC3164_GOLDEN[1][2].encode(), V.RFC3164_GOLDEN[9][2].encode(), b"", b"abc", "h: 2019 Mar 27 12:09:39: m  ".encode()]
    d0, o0 = oracle.pack(lines)
    obuf, ooffs = oracle.decode_dump(R3, d0, o0, oracle.Rfc3164Config(YEAR))
    want = [obuf[ooffs[i]:ooffs[i + 1]] for i in range(len(lines))]
    for mode, term in ((1, [b"\n", b"\r\n"]), (2, [b"\0", b"\0"])):
        raw = [l + term[k % 2] for k, l in enumerate(lines)]
        data, offs = oracle.pack(raw)
        gbuf, goffs, _ = emu.r3164_decode_dump(native, data, offs, YEAR, strip_eol=mode)
        # spans differ by construction (terminators inside the stream); the decoded Records must not
        assert [gbuf[goffs[i]:goffs[i + 1]] for i in range(len(lines))] == want
    inv = np.array([0, 1, 0, 0, 0], dtype=np.uint8)
    raw = [l + b"\n" for l in lines]
    data, offs = oracle.pack(raw)
    gbuf, goffs, _ = emu.r3164_decode_dump(native, data, offs, YEAR, strip_eol=1, invalid=inv)
    assert gbuf[goffs[1]:goffs[2]] == b"E:Invalid UTF-8 input;out=0" and gbuf[goffs[0]:goffs[1]] == want[0]


def test_lockstep_convergence_in_a_32_lane_emulation(emu, native, oracle):
    """r3164_parse_lockstep with a warp emulated as 32 concurrent host threads (tests/emu: libfg_emu_warp.so): fg_any is a
    rendezvous of the lanes, and at every rendezvous all 32 lanes must be at the SAME vote (same source line), all must make
    the same number of votes, and the results (arena allocation by concurrent atomics included) must equal the oracle's."""
    lines = [l.encode() for _, _, l, _ in V.RFC3164_GOLDEN] + [l.encode() for l, _ in V.RFC3164_CASES]
    data, offs = oracle.pack(lines)
    d2, o2 = native.generate(native.FMT_RFC3164, 11, 1536, bad_frac=0.08)
    for dat, off in ((data, offs), (d2, o2)):
        gbuf, goffs, info = emu.r3164_decode_dump(native, dat, off, YEAR, warp=True)
        obuf, ooffs = oracle.decode_dump(R3, dat, off, oracle.Rfc3164Config(YEAR))
        assert info["vote_mismatches"] == 0 and info["votes"] > 500, info
        assert gbuf == obuf and np.array_equal(goffs, ooffs)
    assert sum(emu.r3164_site_votes().values()) > 10_000


def test_zone_tables_three_way(emu, native):
    """The product's TZif reader + POSIX footer expansion + packed search (fg_tz.cu / fg_rfc3164.cuh), reached both through
    the emulation build and through the C ABI's host-side query (fg_tz_lookup in libflowgger_cuda.so), against the oracle's
    independent Python reader (oracle/tzread.py, itself checked against zoneinfo in test_oracle_golden.py): every zone,
    at every transition boundary and at random local times."""
    import random
    import tzread
    zones = tzread.load_zones()
    assert emu.tz_count() == len(zones) == native.tz_count()
    rnd = random.Random(8)
    for name, z in zones.items():
        tr, of = z
        probes = [rnd.randrange(-3_000_000_000, 14_000_000_000) for _ in range(8)] + [-(1 << 40), 1 << 40]
        for k in range(len(tr)):
            if k % 7 == 0 or k > len(tr) - 6:
                probes += [tr[k] + o + d for o in (of[k], of[k + 1]) for d in (-1, 0, 1)]
        for local in probes:
            want = tzread.offset_at_local(z, local)
            assert emu.tz_lookup(name, local) == want, (name, local)
        for local in probes[:12]:
            assert native.tz_lookup(name, local) == tzread.offset_at_local(z, local), (name, local)
    for bogus in ("utc", "UTC ", "Europe", "Europe/", "Mars/Phobos", "posixrules", "localtime", "", "Z" * 40):
        assert emu.tz_lookup(bogus, 0) is None and native.tz_lookup(bogus, 0) is None, bogus


def test_token_soup_three_way(emu, native, oracle):
    """Lines assembled at random from the pieces the decoder distinguishes (PRI forms, years, months, days, times, zone
    names and near misses, separators of every White_Space kind, ": " in odd places): device logic (both walkers) ==
    C++ oracle == the independent Python restatement (oracle/pyrfc3164.py)."""
    import random
    import pyrfc3164
    rnd = random.Random(31640)
    pri = ["", "", "<13>", "<0>", "<191>", "<255>", "<256>", "<+7>", "<<5>", "<>", "<1 2>", "<13", "< 13>", "<013>"]
    year = ["", "", "", "2020 ", "1999 ", "0001 ", "9999 ", "+2020 ", "20200 ", "202 ", "2024 ", "2023 "]
    mon = ["Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec", "jan", "Sept", "Au", "Aug:"]
    day = ["1", "6", " 6", "06", "28", "29", "30", "31", "32", "0", "00", "7th", ""]
    tm = ["11:15:24", "00:00:00", "23:59:59", "24:00:00", "23:60:00", "23:59:60", "1:15:24", "11:15", "11:15:24:", "11-15-24", "11:15:24.5"]
    zone = ["", "", "", "UTC ", "GMT ", "Europe/Paris ", "America/New_York ", "America/Argentina/Buenos_Aires ", "utc ", "Europe/paris ",
            "Etc/GMT-14 ", "Asia/Kolkata ", "EST5EDT ", "Zulu ", "Mars/Phobos ", "Australia/Lord_Howe ", "Europe/Paris: "]
    host = ["host", "web-01.example.org", "h", "été", "UTC", "10.0.0.1", "host:", ""]
    sep = [" ", " ", " ", "  ", "\t", " ", " ", "　", " \t ", "\r\n", "\x0b", "\x1f", "​"]
    # most picks are well-formed (so that every later stage of the decoder is reached), the rest are the near misses
    mon = mon[:12] * 5 + mon
    day = ["1", "6", " 6", "06", "28"] * 4 + day
    tm = ["11:15:24", "00:00:00", "23:59:59"] * 6 + tm
    pri = ["", "<13>", "<0>", "<191>"] * 4 + pri
    year = ["", "", "2020 ", "1999 ", "2024 "] * 3 + year
    words = ["error", "GET", "/x", "200", "a:", ": ", "b", "naïve", "日本", "x y", "[1]", "tag[2]:", "", "\x01"]
    lines = []
    for _ in range(30_000):
        kind = rnd.random()
        date = rnd.choice(year) + rnd.choice(mon) + rnd.choice(sep) + rnd.choice(day) + rnd.choice(sep) + rnd.choice(tm)
        msg = rnd.choice(sep if rnd.random() < 0.2 else [" "]).join(rnd.choice(words) for _ in range(rnd.randrange(0, 9)))
        tail = rnd.choice(["", "", "", " ", "\n", " \t", " "])
        if kind < 0.7:
            l = rnd.choice(pri) + date + rnd.choice(sep) + rnd.choice(zone) + rnd.choice(host) + rnd.choice(sep) + msg + tail
        elif kind < 0.95:
            l = rnd.choice(pri) + rnd.choice(host) + rnd.choice([": ", ": ", ":", " : ", ":  "]) + date + rnd.choice(["", " "]) + rnd.choice(zone).strip() + \
                rnd.choice([": ", ": ", ":", " :", ""]) + msg + tail
        else:
            l = rnd.choice(pri) + msg + tail
        lines.append(l.encode())
    data, offs = oracle.pack(lines)
    check(emu, native, oracle, data, offs)
    obuf, ooffs = oracle.decode_dump(R3, data, offs, oracle.Rfc3164Config(YEAR))
    ok = 0
    for i, l in enumerate(lines):
        if b"9999" in l:
            continue  # past the year 2400 the zone tables keep their last offset (fg_tz.h: kTzLastYear), zoneinfo keeps applying the rule
        r = pyrfc3164.decode(l.decode(), YEAR)
        if r is pyrfc3164.UNSUPPORTED:
            continue
        assert pyrfc3164.dump(r) == obuf[ooffs[i]:ooffs[i + 1]], (l, pyrfc3164.dump(r), obuf[ooffs[i]:ooffs[i + 1]])
        ok += isinstance(r, dict)
    assert ok > 5000  # a good share of the soup decodes
