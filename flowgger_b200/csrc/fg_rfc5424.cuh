// fg_rfc5424.cuh — one RFC5424 line -> Record fields, on device.
//
// B200-native replacement for RFC5424Decoder::decode
// (/root/reference/src/flowgger/decoder/rfc5424_decoder.rs:18-49) and its helpers
// BOM::parse :63-71, parse_pri_version :74-92, rfc3339_to_unix :94-99,
// parse_data :127-161, parse_msg :163-172, parse_sd_data :174-242.
// The unescape of SD values (:105-125) is deferred to Record materialisation:
// the table carries the raw value span plus FG_EM_UNESCAPE.
//
// One thread owns one line whose bytes are already staged in shared memory by
// the CTA-wide bulk copy (fg_kernels.cu); all positions are relative to the
// line start.
#pragma once
#include "fg_common.cuh"
#include "fg_status.h"

namespace fg {

struct LineResult {
    double ts;
    uint32_t status;
    uint32_t facility, severity, flags;
    // spans relative to the line start; off < 0 => None
    int host_o, host_l, app_o, app_l, proc_o, proc_l, mid_o, mid_l, msg_o, msg_l, full_o, full_l;
    uint32_t n_entries;  // SD headers + pairs
    int sd_pos;          // where the SD text starts (position of the first '['), for the emit pass
};

struct EntrySink {
    int2* name;
    unsigned long long* val;
    uint8_t* meta;
};

// The structured-data walk (parse_data :134-158 + parse_sd_data :174-242) as
// an explicit 5-state machine.  The reference's 6-tuple match collapses to:
//   OUT  : !in_name, name None, !in_value        NAME : in_name
//   EQ   : name Some, !in_value (only '"' legal) VAL  : in_value, !esc
//   VESC : in_value, esc
// EMIT=false counts entries and finds the message; EMIT=true re-walks an
// already validated SD text and writes the side-table rows.
template <bool EMIT>
FG_DEV uint32_t rfc5424_sd_walk(bytes_t p, int len, int pos, int line_off, uint32_t& n_entries, int& msg_from,
                                const EntrySink& sink, uint32_t ebase) {
    uint32_t n = 0;
    for (;;) {
        // parse_sd_data(line, pos + 1): sd_id = up to the first ' '
        int s = pos + 1;
        while (s < len && p[s] != ' ') ++s;
        if (s >= len) return FG_E5_MISSING_SD;  // :177
        const uint32_t header = n++;
        uint32_t pairs = 0;
        int i = s + 1;
        int state = 0;  // 0 OUT, 1 NAME, 2 EQ, 3 VAL, 4 VESC
        int name_start = 0, name_end = 0, value_start = 0;
        bool has_bs = false;
        int after = -1;
        for (; i < len; ++i) {
            const uint32_t c = p[i];
            if (state == 3) {
                if (c == '\\') {
                    state = 4;
                    has_bs = true;
                } else if (c == '"') {
                    if (EMIT) {
                        const uint32_t e = ebase + n;
                        sink.name[e] = make_int2(line_off + name_start, name_end - name_start);
                        sink.val[e] = (unsigned long long)(uint32_t)(line_off + value_start) |
                                      ((unsigned long long)(uint32_t)(i - value_start) << 32);
                        sink.meta[e] = (uint8_t)(0u /*FG_TAG_STRING*/ | (has_bs ? 0x08u : 0u));
                    }
                    ++n;
                    ++pairs;
                    state = 0;
                }
            } else if (state == 4) {
                state = 3;
            } else if (state == 0) {
                if (c == ' ' || c == '"') {
                } else if (c == ']') {
                    after = i + 1;
                    break;
                } else if (c >= 33u && c <= 126u && c != '=') {  // is_sd_name :188-192 ('"' and ']' handled above)
                    state = 1;
                    name_start = i;
                } else {
                    return FG_E5_SD_FORMAT;  // :235
                }
            } else if (state == 1) {
                if (c == '=') {
                    name_end = i;
                    state = 2;
                } else if (!(c >= 33u && c <= 126u && c != '"' && c != ']')) {
                    return FG_E5_SD_FORMAT;
                }
            } else {  // state 2: only '"' opens the value (:212)
                if (c != '"') return FG_E5_SD_FORMAT;
                state = 3;
                value_start = i + 1;
                has_bs = false;
            }
        }
        if (after < 0) return FG_E5_SD_NO_END;  // :239
        if (EMIT) {
            const uint32_t e = ebase + header;
            sink.name[e] = make_int2(line_off + pos + 1, s - (pos + 1));
            sink.val[e] = pairs;
            sink.meta[e] = 7u;  // FG_TAG_SD_HEADER
        }
        if (after >= len) return FG_E5_MISSING_MSG;  // :148
        const uint32_t c = p[after];
        if (c == '[') {
            pos = after;
            continue;
        }
        if (c != ' ') return FG_E5_MALFORMED;  // :154
        msg_from = after;
        n_entries = n;
        return FG_ST_OK;
    }
}

FG_DEV void rfc5424_parse_line(bytes_t p, int len, LineResult& r) {
    r.ts = 0.0;
    r.facility = 0xFFu;
    r.severity = 0xFFu;
    r.flags = 0;
    r.host_o = r.app_o = r.proc_o = r.mid_o = r.msg_o = r.full_o = -1;
    r.host_l = r.app_l = r.proc_l = r.mid_l = r.msg_l = r.full_l = 0;
    r.n_entries = 0;
    r.sd_pos = -1;

    // BOM::parse :63-71
    int b = 0;
    if (len >= 3 && p[0] == 0xEFu && p[1] == 0xBBu && p[2] == 0xBFu) {
        b = 3;
    } else if (!(len > 0 && p[0] == '<')) {
        r.status = FG_E5_BOM;
        return;
    }
    // splitn(7, ' ') :23 — positions of the first six spaces
    int sp[6];
    int nsp = 0;
    {
        int i = b;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            sp[k] = len;
            if (nsp == k) {
                while (i < len && p[i] != ' ') ++i;
                if (i < len) {
                    sp[k] = i;
                    ++nsp;
                    ++i;
                }
            }
        }
    }
    // parse_pri_version :74-92 on part0 = [b, sp[0])
    {
        const int e0 = sp[0];
        if (!(b < e0 && p[b] == '<')) {
            r.status = FG_E5_PRI_BRACKETS;
            return;
        }
        int gt = b + 1;
        while (gt < e0 && p[gt] != '>') ++gt;
        uint32_t pri;
        if (!parse_u8(p, b + 1, gt, pri)) {
            r.status = FG_E5_INVALID_PRI;
            return;
        }
        if (gt >= e0) {
            r.status = FG_E5_MISSING_VERSION;
            return;
        }
        if (!(e0 - gt == 2 && p[gt + 1] == '1')) {
            r.status = FG_E5_UNSUPPORTED_VERSION;
            return;
        }
        r.facility = pri >> 3;
        r.severity = pri & 7u;
    }
    if (nsp < 1) {
        r.status = FG_E5_MISSING_TS;
        return;
    }
    if (!parse_rfc3339(p, sp[0] + 1, sp[1], r.ts)) {
        r.status = FG_E5_BAD_TS;
        return;
    }
    if (nsp < 6) {
        r.status = FG_E5_MISSING_HOST + (uint32_t)(nsp - 1);  // :26-30 in order
        return;
    }
    // parse_data :127-161 on part6 = [sp[5]+1, len)
    const int d = sp[5] + 1;
    if (d >= len) {
        r.status = FG_E5_MISSING_MSG;  // :129
        return;
    }
    const uint32_t c0 = p[d];
    int msg_from;
    if (c0 == '-') {
        msg_from = d + 1;
    } else if (c0 == '[') {
        EntrySink none = {nullptr, nullptr, nullptr};
        uint32_t st = rfc5424_sd_walk<false>(p, len, d, 0, r.n_entries, msg_from, none, 0);
        if (st != FG_ST_OK) {
            r.status = st;
            r.n_entries = 0;
            return;
        }
        r.sd_pos = d;
    } else {
        r.status = FG_E5_MALFORMED;  // :159
        return;
    }
    // parse_msg :163-172
    {
        int hi = trim_end(p, msg_from, len);
        int lo = trim_start(p, msg_from, hi);
        if (hi > lo) {
            r.msg_o = lo;
            r.msg_l = hi - lo;
        }
    }
    r.host_o = sp[1] + 1;
    r.host_l = sp[2] - sp[1] - 1;
    r.app_o = sp[2] + 1;
    r.app_l = sp[3] - sp[2] - 1;
    r.proc_o = sp[3] + 1;
    r.proc_l = sp[4] - sp[3] - 1;
    r.mid_o = sp[4] + 1;
    r.mid_l = sp[5] - sp[4] - 1;
    r.full_o = b;  // line.trim_end() of the BOM-stripped line :46
    r.full_l = trim_end(p, b, len) - b;
    r.status = FG_ST_OK;
}

FG_DEV void rfc5424_emit(bytes_t p, int len, int line_off, const LineResult& r, const EntrySink& sink, uint32_t ebase) {
    uint32_t n;
    int mf;
    rfc5424_sd_walk<true>(p, len, r.sd_pos, line_off, n, mf, sink, ebase);
}

}  // namespace fg
