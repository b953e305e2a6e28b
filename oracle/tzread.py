"""Zone tables for the RFC3164 oracle — TEST INFRASTRUCTURE (see oracle/oracle.hpp).

The reference resolves the optional zone name of an RFC3164 timestamp with time_tz::timezones::get_by_name and
assume_timezone (rfc3164_decoder.rs:196-203); time-tz compiles the IANA database in.  The oracle takes the same database
from the system's TZif files (RFC 8536), read here in Python independently of the product's C++ reader
(flowgger_b200/csrc/fg_tz.cu), and tests/test_tz_tables.py checks both against the standard library's `zoneinfo`.

A zone is (trans, offs): ascending UTC transition seconds and len(trans) + 1 UTC offsets, offs[k] in force on
[trans[k-1], trans[k]).  Explicit transitions come from the 64-bit data block; the years after the last one up to
LAST_YEAR are generated from the POSIX TZ footer.
"""
from __future__ import annotations

import os
import re
import struct
from pathlib import Path

LAST_YEAR = 2400  # footer rules are expanded up to and including this year; later times keep the last offset
SKIP_NAMES = {"posixrules", "localtime"}  # not IANA identifiers
SKIP_DIRS = {"posix", "right"}


def tzdir() -> Path:
    return Path(os.environ.get("TZDIR") or "/usr/share/zoneinfo")


def _days_from_civil(y: int, m: int, d: int) -> int:
    y -= m <= 2
    era = (y if y >= 0 else y - 399) // 400
    yoe = y - era * 400
    doy = (153 * (m + (-3 if m > 2 else 9)) + 2) // 5 + d - 1
    return era * 146097 + yoe * 365 + yoe // 4 - yoe // 100 + doy - 719468


def _is_leap(y: int) -> bool:
    return (y % 4 == 0 and y % 100 != 0) or y % 400 == 0


_MDAYS = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]


def _posix_offset(s: str, i: int) -> tuple[int, int]:
    """[+-]hh[:mm[:ss]] at s[i:] -> (seconds, next index)."""
    sign = 1
    if i < len(s) and s[i] in "+-":
        sign = -1 if s[i] == "-" else 1
        i += 1
    m = re.match(r"(\d{1,3})(?::(\d{1,2}))?(?::(\d{1,2}))?", s[i:])
    if not m:
        raise ValueError(f"bad offset in TZ string {s!r}")
    h, mi, se = int(m.group(1)), int(m.group(2) or 0), int(m.group(3) or 0)
    return sign * (h * 3600 + mi * 60 + se), i + m.end()


def _posix_name(s: str, i: int) -> int:
    if i < len(s) and s[i] == "<":
        return s.index(">", i) + 1
    j = i
    while j < len(s) and s[j].isalpha():
        j += 1
    return j


def parse_posix_tz(s: str):
    """-> (std_utoff, None) or (std_utoff, (dst_utoff, start_rule, end_rule)); a rule is (kind, a, b, c, time_seconds)."""
    i = _posix_name(s, 0)
    off, i = _posix_offset(s, i)
    std = -off
    if i >= len(s):
        return std, None
    i = _posix_name(s, i)
    dst = std + 3600
    if i < len(s) and s[i] != ",":
        off, i = _posix_offset(s, i)
        dst = -off
    if i >= len(s):
        return std, None  # a DST name without rules: no transitions can be generated
    rules = []
    for part in s[i + 1:].split(","):
        date, _, tm = part.partition("/")
        t = 7200
        if tm:
            t, _ = _posix_offset(tm, 0)
        if date.startswith("M"):
            m, w, d = (int(x) for x in date[1:].split("."))
            rules.append(("M", m, w, d, t))
        elif date.startswith("J"):
            rules.append(("J", int(date[1:]), 0, 0, t))
        else:
            rules.append(("N", int(date), 0, 0, t))
    return std, (dst, rules[0], rules[1])


def _rule_day(rule, year: int) -> int:
    """days since the epoch of the rule's date in `year`."""
    kind, a, b, c, _ = rule
    jan1 = _days_from_civil(year, 1, 1)
    if kind == "J":  # 1..365, 29 February is never counted
        return jan1 + a - 1 + (1 if _is_leap(year) and a >= 60 else 0)
    if kind == "N":  # 0..365, leap days counted
        return jan1 + a
    m, w, d = a, b, c
    first = _days_from_civil(year, m, 1)
    wd_first = (first + 4) % 7  # 1970-01-01 was a Thursday (4), 0 = Sunday
    day = 1 + (d - wd_first) % 7 + (w - 1) * 7
    mdays = _MDAYS[m - 1] + (1 if m == 2 and _is_leap(year) else 0)
    if day > mdays:
        day -= 7
    return first + day - 1


def read_tzif(path: Path):
    """-> (trans, offs) or None if `path` is not a version >= 2 TZif file."""
    data = path.read_bytes()
    if data[:4] != b"TZif" or data[4:5] in (b"\0", b""):
        return None
    isut, isstd, leap, timecnt, typecnt, charcnt = struct.unpack(">6I", data[20:44])
    p = 44 + timecnt * 4 + timecnt + typecnt * 6 + charcnt + leap * 8 + isstd + isut
    if data[p:p + 4] != b"TZif":
        return None
    isut, isstd, leap, timecnt, typecnt, charcnt = struct.unpack(">6I", data[p + 20:p + 44])
    p += 44
    times = list(struct.unpack(f">{timecnt}q", data[p:p + 8 * timecnt]))
    p += 8 * timecnt
    idx = list(data[p:p + timecnt])
    p += timecnt
    utoff = [struct.unpack(">i", data[p + 6 * k:p + 6 * k + 4])[0] for k in range(typecnt)]
    p += typecnt * 6 + charcnt + leap * 12 + isstd + isut
    footer = data[p:].split(b"\n")[1].decode("ascii") if data[p:p + 1] == b"\n" else ""
    trans = times
    offs = [utoff[0]] + [utoff[i] for i in idx]  # RFC 8536 3.2: before the first transition, time type 0
    if footer:
        std, dst_rules = parse_posix_tz(footer)
        if dst_rules is not None:
            dst, start, end = dst_rules
            last = trans[-1] if trans else -(1 << 62)
            y0 = 1970
            if trans:
                days = last // 86400
                y0 = 1970 + int(days // 366) - 1
            extra = []
            for y in range(max(y0, 1900), LAST_YEAR + 1):
                extra.append((_rule_day(start, y) * 86400 + start[4] - std, dst))  # start time is in standard time
                extra.append((_rule_day(end, y) * 86400 + end[4] - dst, std))      # end time is in daylight time
            extra.sort()
            for t, o in extra:
                if t > last:
                    trans.append(t)
                    offs.append(o)
        elif trans and offs[-1] != std:
            pass  # the footer only restates the last type for the files zic writes; nothing to add
    return trans, offs


def load_zones(root: Path | None = None) -> dict[str, tuple[list[int], list[int]]]:
    root = root or tzdir()
    zones = {}
    for dirpath, dirnames, filenames in os.walk(root):
        rel = Path(dirpath).relative_to(root)
        if rel.parts and rel.parts[0] in SKIP_DIRS:
            dirnames[:] = []
            continue
        for fn in filenames:
            name = str(rel / fn) if rel.parts else fn
            if name in SKIP_NAMES:
                continue
            try:
                z = read_tzif(Path(dirpath) / fn)
            except (OSError, ValueError, struct.error, IndexError):
                z = None
            if z is not None:
                zones[name] = z
    return zones


def offset_at_local(zone, local: int) -> int:
    """The rule oracle/rfc3164.cpp applies, restated for the zoneinfo cross-check."""
    trans, offs = zone
    n = len(trans)
    for k in range(n + 1):
        if (k == 0 or local >= trans[k - 1] + offs[k]) and (k == n or local < trans[k] + offs[k]):
            return offs[k]
    for k in range(n):
        if trans[k] + offs[k] <= local < trans[k] + offs[k + 1]:
            return offs[k]
    return offs[n]
