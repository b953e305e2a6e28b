// fg_rfc3164.cuh — RFC3164Decoder::decode for ONE line (decoder/rfc3164_decoder.rs:31-213), device logic.
//
// The reference tokenises into owned Vecs and re-joins; here a line is walked in place:
//   parse_strip_pri (:126-153)      r3_strip_pri
//   decode_rfc_standard (:55-87)    the first six whitespace tokens are located (split_whitespace = Unicode White_Space),
//                                   the date is read off them, then the rest of the line is checked ONCE, four bytes per
//                                   step: if it holds nothing but single ASCII spaces between its tokens the re-joined
//                                   message (`tokens.join(" ")`, :67) IS a span of the input; otherwise the tokens are
//                                   copied, joined by one space, into the arena
//   decode_rfc_custom (:89-123)     two ": " separators; split + join with the same separator is the identity, so the
//                                   message is always a span
//   parse_date (:163-213)           time 0.3 format description "[year] [month repr:short] [day padding:none]
//                                   [hour]:[minute]:[second]" applied to the tokens directly (a token holds no space, the
//                                   description puts one space between its items, so token k must match item k exactly)
//   get_by_name / assume_timezone   tz_find / tz_offset_local over the packed zone table (fg_tz.cu)
// The error returned is that of the last form tried (:40-47).  tests/emu compiles this file with g++.
#pragma once
#include <stdint.h>

#include "fg_common.cuh"
#include "fg_kernels.cuh"
#include "fg_status.h"

#ifdef FG_HOST_EMU
#define FG_HD static inline
#elif defined(__CUDACC__)
#define FG_HD __host__ __device__ __forceinline__
#else
#define FG_HD static inline
#endif

namespace fg {

// ---- zone table ---------------------------------------------------------------------------------------------------
FG_HD unsigned long long tz_fnv1a(const uint8_t* p, int a, int b) {
    unsigned long long h = 0xcbf29ce484222325ull;
    for (int i = a; i < b; ++i) {
        h ^= (unsigned long long)p[i];
        h *= 0x100000001b3ull;
    }
    return h;
}
// get_by_name (exact, case-sensitive): zone of the identifier p[a, b), or -1
FG_HD int tz_find(const TzDeviceTable& T, const uint8_t* p, int a, int b) {
    const int len = b - a;
    if (T.n_names == 0 || len < T.min_len || len > T.max_len) return -1;
    const uint32_t c0 = p[a];
    if (!((T.first_mask[c0 >> 5] >> (c0 & 31u)) & 1u)) return -1;  // ordinary hostnames stop here
    const unsigned long long h = tz_fnv1a(p, a, b);
    int lo = 0, hi = T.n_names;  // first k with hash[k] >= h
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (T.hash[mid] < h) lo = mid + 1;
        else hi = mid;
    }
    for (int k = lo; k < T.n_names && T.hash[k] == h; ++k) {
        const int na = T.name_off[k], nb = T.name_off[k + 1];
        if (nb - na != len) continue;
        bool same = true;
        for (int i = 0; i < len; ++i) same = same && T.names[na + i] == p[a + i];
        if (same) return T.zone[k];
    }
    return -1;
}
// assume_timezone: the UTC offset in force at LOCAL second `local` — the last span whose key is <= local.  fg_tz.cu sets
// key = transition + max(offset before, offset after): a local time that occurs twice takes the earlier offset, one the
// jump skips the offset before the jump.
FG_HD int tz_offset_local(const TzDeviceTable& T, int z, long long local) {
    int lo = T.first[z], hi = T.first[z + 1];
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (T.key[mid] <= local) lo = mid;
        else hi = mid;
    }
    return T.off[lo];
}

#if defined(FG_HOST_EMU) || defined(__CUDACC__)

struct R3Out {
    uint32_t status, facility, severity, flags;
    double ts;
    int host_o, host_l, msg_o, msg_l, full_l;  // relative to the line start; msg_o indexes the arena with FG_FLAG_MSG_ARENA
};
constexpr uint32_t kR3FlagMsgArena = 0x40u;  // include/flowgger_cuda.h: FG_FLAG_MSG_ARENA

FG_DEV uint32_t r3_ld32(const uint8_t* q) {  // q is 4-byte aligned
#ifdef FG_HOST_EMU
    uint32_t w;
    memcpy(&w, q, 4);
    return w;
#else
    return *reinterpret_cast<const uint32_t*>(q);
#endif
}

// next token of str::split_whitespace at or after pos, inside [pos, end)
FG_DEV bool r3_token(bytes_t p, int& pos, int end, int& a, int& b) {
    while (pos < end) {
        const int w = ws_len_front(p, pos, end);
        if (!w) break;
        pos += w;
    }
    if (pos >= end) return false;
    a = pos;
    while (pos < end) {
        const uint32_t c = p[pos];
        if (c < 0x80u) {
            if (is_ascii_ws(c)) break;
        } else if (ws_len_front(p, pos, end)) {  // a continuation byte never starts a White_Space character
            break;
        }
        ++pos;
    }
    b = pos;
    return true;
}

// [month repr:short], case-sensitive
FG_DEV int r3_month(bytes_t p, int a, int b) {
    if (b - a != 3) return 0;
    const uint32_t k = ((uint32_t)p[a] << 16) | ((uint32_t)p[a + 1] << 8) | (uint32_t)p[a + 2];
    switch (k) {
        case 0x4A616Eu: return 1;   // Jan
        case 0x466562u: return 2;   // Feb
        case 0x4D6172u: return 3;   // Mar
        case 0x417072u: return 4;   // Apr
        case 0x4D6179u: return 5;   // May
        case 0x4A756Eu: return 6;   // Jun
        case 0x4A756Cu: return 7;   // Jul
        case 0x417567u: return 8;   // Aug
        case 0x536570u: return 9;   // Sep
        case 0x4F6374u: return 10;  // Oct
        case 0x4E6F76u: return 11;  // Nov
        case 0x446563u: return 12;  // Dec
        default: return 0;
    }
}
// [year]: optional sign, exactly four digits (time 0.3 without large-dates)
FG_DEV bool r3_year(bytes_t p, int a, int b, int& year) {
    bool neg = false;
    if (a < b && (p[a] == '+' || p[a] == '-')) {
        neg = p[a] == '-';
        ++a;
    }
    if (b - a != 4) return false;
    int i = a, y;
    if (!four_digits(p, i, b, y)) return false;
    year = neg ? -y : y;
    return true;
}
// "[month repr:short] [day padding:none] [hour]:[minute]:[second]" on three tokens + Date::from_calendar_date /
// Time::from_hms -> seconds since the epoch of that civil time read as UTC
FG_DEV bool r3_date(bytes_t p, int year, const int* ta, const int* tb, long long& local) {
    const int month = r3_month(p, ta[0], tb[0]);
    if (!month) return false;
    const int dl = tb[1] - ta[1];
    if (dl < 1 || dl > 2) return false;  // one or two digits, then the separating space
    int day = 0;
    for (int i = ta[1]; i < tb[1]; ++i) {
        const uint32_t d = (uint32_t)p[i] - '0';
        if (d > 9u) return false;
        day = day * 10 + (int)d;
    }
    if (day == 0) return false;  // NonZeroU8
    if (tb[2] - ta[2] != 8) return false;
    int i = ta[2], hh, mm, ss;
    if (!two_digits(p, i, tb[2], hh) || p[i] != ':') return false;
    ++i;
    if (!two_digits(p, i, tb[2], mm) || p[i] != ':') return false;
    ++i;
    if (!two_digits(p, i, tb[2], ss)) return false;
    if (day > days_in_month(year, month)) return false;
    if (hh > 23 || mm > 59 || ss > 59) return false;
    local = (long long)days_from_civil(year, month, day) * 86400 + hh * 3600 + mm * 60 + ss;
    return true;
}

// parse_date_token (:155-161) + parse_date (:163-213) over the first `ntok` (<= 6 known) tokens; idx = tokens consumed
FG_DEV uint32_t r3_parse_date_token(bytes_t p, const R3164DeviceConfig& cfg, int ntok, const int* ta, const int* tb, double& ts,
                                    int& idx) {
    if (ntok < 3) return FG_E3_TIME_FORMAT;
    long long local = 0;
    // without a year: format!("{} {}", current_year, ...) parses only when the year prints as four digits
    bool ok = cfg.year >= 1000 && cfg.year <= 9999 && r3_date(p, cfg.year, ta, tb, local);
    idx = 3;
    if (!ok) {
        if (ntok < 4) return FG_E3_WITH_YEAR;
        int year;
        if (!r3_year(p, ta[0], tb[0], year) || !r3_date(p, year, ta + 1, tb + 1, local)) return FG_E3_DATE;
        idx = 4;
    }
    if (ntok > idx) {
        // (constant indices keep the token arrays in registers)
        const int za = idx == 3 ? ta[3] : ta[4], zb = idx == 3 ? tb[3] : tb[4];
        const int z = tz_find(cfg.tz, p, za, zb);
        if (z >= 0) {
            local -= (long long)tz_offset_local(cfg.tz, z, local);
            idx += 1;
        }
    }
    ts = unix_nanos_to_ts(local, 0u);
    return FG_ST_OK;
}

// true: p[a, b) holds only ASCII bytes >= 0x20 and no two adjacent spaces (so split_whitespace + join(" ") returns it as is,
// given that it neither starts nor ends with a space)
FG_DEV bool r3_plain_run(bytes_t p, int a, int b) {
    bool prev_sp = false;
    int i = a;
    while (i < b && ((size_t)(p + i) & 3u) != 0u) {
        const uint32_t c = p[i];
        if (c < 0x20u || c >= 0x80u) return false;
        const bool sp = c == 0x20u;
        if (sp && prev_sp) return false;
        prev_sp = sp;
        ++i;
    }
    for (; i + 4 <= b; i += 4) {
        const uint32_t w = r3_ld32(p + i);
        if ((w & 0x80808080u) != 0u || ((w - 0x20202020u) & ~w & 0x80808080u) != 0u) return false;  // a byte >= 0x80 or < 0x20
        const uint32_t t = w ^ 0x20202020u;
        const uint32_t z = ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u;  // bit 8k+7: byte k is a space (exact)
        if ((z & (z >> 8)) != 0u || (prev_sp && (z & 0x80u) != 0u)) return false;
        prev_sp = (z >> 31) != 0u;
    }
    for (; i < b; ++i) {
        const uint32_t c = p[i];
        if (c < 0x20u || c >= 0x80u) return false;
        const bool sp = c == 0x20u;
        if (sp && prev_sp) return false;
        prev_sp = sp;
    }
    return true;
}

// `_log_tokens[1..].join(" ")` (:67) for the text after the hostname token, which ends at `from`; `end` = the line without
// its trailing whitespace
FG_DEV void r3_message(bytes_t p, int from, int end, const R3164DeviceConfig& cfg, R3Out& r) {
    int pos = from, a, b;
    if (!r3_token(p, pos, end, a, b)) {  // no token left: Some("")
        r.msg_o = from;
        r.msg_l = 0;
        return;
    }
    if (r3_plain_run(p, a, end)) {
        r.msg_o = a;
        r.msg_l = end - a;
        return;
    }
    // exact pass: is every gap one ASCII space after all (non-ASCII text, control bytes that are not White_Space)?
    const int first = a;
    bool regular = true;
    uint32_t joined = (uint32_t)(b - a);
    int last_b = b;
    for (;;) {
        int a2, b2;
        if (!r3_token(p, pos, end, a2, b2)) break;
        regular = regular && a2 - last_b == 1 && p[last_b] == 0x20u;
        joined += 1u + (uint32_t)(b2 - a2);
        last_b = b2;
    }
    if (regular) {
        r.msg_o = first;
        r.msg_l = last_b - first;
        return;
    }
    const uint32_t at = atomicAdd(cfg.arena_counter, joined);
    r.msg_o = (int)at;
    r.msg_l = (int)joined;
    r.flags |= kR3FlagMsgArena;
    if ((unsigned long long)at + joined > (unsigned long long)cfg.arena_cap) return;  // the host regrows the arena and redoes the batch
    uint8_t* o = cfg.arena + at;
    pos = first;
    bool lead = true;
    for (;;) {
        int a2, b2;
        if (!r3_token(p, pos, end, a2, b2)) break;
        if (!lead) *o++ = 0x20u;
        lead = false;
        for (int i = a2; i < b2; ++i) *o++ = p[i];
    }
}

// Decoder::decode (:31-48) for the line p[0, len)
FG_DEV void r3164_parse_line(bytes_t p, int len, const R3164DeviceConfig& cfg, R3Out& r) {
    r.status = FG_ST_OK;
    r.facility = r.severity = 0xFFu;
    r.flags = 0u;
    r.ts = 0.0;
    r.host_o = r.msg_o = -1;
    r.host_l = r.msg_l = 0;
    r.full_l = 0;
    // parse_strip_pri (:126-153)
    int m0 = 0;
    if (len > 0 && p[0] == '<') {
        int gt = 1;
        while (gt < len && p[gt] != '>') ++gt;
        if (gt >= len) { r.status = FG_E3_PRI_MALFORMED; return; }
        int a = 0, b = gt + 1;
        while (a < b && p[a] == '<') ++a;       // trim_start_matches('<')
        while (b > a && p[b - 1] == '>') --b;   // trim_end_matches('>')
        uint32_t pri;
        if (!parse_u8(p, a, b, pri)) { r.status = FG_E3_PRI_INVALID; return; }
        r.facility = pri >> 3;
        r.severity = pri & 7u;
        m0 = gt + 1;
    }
    const int end = trim_end(p, 0, len);  // full_msg = line.trim_end() (:77, :112)
    // decode_rfc_standard (:55-87): needs four tokens; the date, an optional zone and the hostname are among the first six
    {
        int ta[6] = {0, 0, 0, 0, 0, 0}, tb[6] = {0, 0, 0, 0, 0, 0}, n = 0, pos = m0;
#pragma unroll
        for (int k = 0; k < 6; ++k)
            if (n == k && r3_token(p, pos, len, ta[k], tb[k])) n = k + 1;
        if (n > 3) {
            double ts;
            int idx;
            if (r3_parse_date_token(p, cfg, n, ta, tb, ts, idx) == FG_ST_OK) {
                if (idx >= n) { r.status = FG_E3_PANIC; return; }  // `_log_tokens[0]` on an empty Vec (:64)
                const int ha = idx == 3 ? ta[3] : (idx == 4 ? ta[4] : ta[5]), hb = idx == 3 ? tb[3] : (idx == 4 ? tb[4] : tb[5]);
                r.ts = ts;
                r.host_o = ha;
                r.host_l = hb - ha;
                r.full_l = end;
                r3_message(p, hb, end, cfg, r);
                return;
            }
        }
    }
    // decode_rfc_custom (:89-123): <hostname>: <date>: <message>
    int s1 = -1, s2 = -1;
    for (int i = m0; i + 1 < len; ++i) {
        if (p[i] == ':' && p[i + 1] == ' ') {
            if (s1 < 0) {
                s1 = i;
                ++i;  // str::split matches do not overlap
            } else {
                s2 = i;
                break;
            }
        }
    }
    if (s2 < 0) { r.status = FG_E3_CUSTOM; return; }
    int ta[5] = {0, 0, 0, 0, 0}, tb[5] = {0, 0, 0, 0, 0}, n = 0, pos = s1 + 2;
#pragma unroll
    for (int k = 0; k < 5; ++k)
        if (n == k && r3_token(p, pos, s2, ta[k], tb[k])) n = k + 1;
    double ts;
    int idx;
    const uint32_t st = r3_parse_date_token(p, cfg, n, ta, tb, ts, idx);
    if (st != FG_ST_OK) { r.status = st; return; }
    r.ts = ts;
    r.host_o = m0;
    r.host_l = s1 - m0;
    r.msg_o = s2 + 2;  // tokens[2..].join(": "): everything behind the second separator, untrimmed
    r.msg_l = len - (s2 + 2);
    r.full_l = end;
}

#endif  // FG_HOST_EMU || __CUDACC__

}  // namespace fg
