// fg_r5fast.cuh — RFC5424 fast path: structural bitmap + bit-walk over a shared-memory tile.
//
// B200-native replacement for RFC5424Decoder::decode
// (/root/reference/src/flowgger/decoder/rfc5424_decoder.rs:18-49) and its helpers BOM::parse :63-71,
// parse_pri_version :74-92, rfc3339_to_unix :94-99, parse_data :127-161, parse_msg :163-172,
// parse_sd_data :174-242.  Two stages per CTA tile (a contiguous span of lines staged by one TMA bulk copy):
//
//   stage 1  r5_classify16: every thread takes 16-byte granules of the flat tile (LDS.128, conflict-free, all 32 lanes
//            busy) and writes one bit per byte into a bitmap: I = "this byte may end a token".  I is a cheap SUPERSET of
//            the bytes the grammar cares about (7 SWAR ops per 4 bytes, exact per byte, no cross-byte carries):
//                b <= 0x22            control bytes, ' ', '!', '"'
//                (b & 0x1E) == 0x1C   0x1C 0x1D '<' '=' '\\' ']' '|' '}'
//                b >= 0x7F            DEL and every non-ASCII byte
//            so every byte that is NOT flagged is a legal SD-NAME character and can never end a header field, an sd_id,
//            a name or a value; the flagged bytes that turn out to be ordinary ('!', '<', '|', '}', ...) cost the walker
//            one extra step.
//   stage 2  r5_walk: one thread per line hops from flagged byte to flagged byte (two bitmap words, a funnel shift and
//            a find-first-set per hop), peeks at the byte and drives the reference's state machine token by token.  The
//            32 lines of a warp advance in lock step (every loop is a warp-uniform `while (any)`), so a warp pays the
//            maximum number of TOKENS over its lanes, not of bytes.
//
// Structured-data rows are staged as 8-byte packed entries (u16 positions relative to the line start) in the line's OWN
// already-consumed bytes of the tile: slot k may be written once the cursor has passed its last byte, so no scratch table
// exists on this path.  A line whose rows do not fit behind the cursor (dozens of 4-byte pairs) or that is 64 KiB or
// longer is handed to the wide kernel (fg_rfc5424.cuh, the round-1 scanner reading global memory).
//
// The unescape of SD values (:105-125) is done by unescape5424_kernel (fg_parse5424.cu) into the batch's arena; this
// walker only reserves the extension slot behind every pair whose value holds a backslash.
#pragma once
#include "fg_common.cuh"
#include "fg_status.h"

namespace fg {

// ---- stage 1 ----------------------------------------------------------------------------------------------------
// 0x80 in every byte of w that is "interesting" (see the header comment); exact per byte.
FG_DEV uint32_t r5_flags(uint32_t w) {
    const uint32_t low = w & 0x7F7F7F7Fu;
    const uint32_t a2 = low + 0x5D5D5D5Du;                    // bit 7: low >= 0x23   (max 0xDC: no carry out of the byte)
    const uint32_t a1 = low + 0x01010101u;                    // bit 7: low == 0x7F
    const uint32_t x = (w & 0x1E1E1E1Eu) ^ 0x1C1C1C1Cu;       // 0 iff (b & 0x1E) == 0x1C
    const uint32_t a3 = x + 0x7F7F7F7Fu;                      // bit 7: x != 0        (x <= 0x1E: no carry)
    return (~a2 | a1 | w | ~a3) & 0x80808080u;
}
// the four 0x80 flags of a word -> bits 28..31 (byte j -> bit 28 + j); bits 24..27 of the product are always 0
FG_DEV uint32_t r5_nibble_top(uint32_t f) { return f * 0x00204081u; }
FG_DEV uint32_t r5_classify16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    const uint32_t n0 = r5_nibble_top(r5_flags(w0)), n1 = r5_nibble_top(r5_flags(w1));
    const uint32_t n2 = r5_nibble_top(r5_flags(w2)), n3 = r5_nibble_top(r5_flags(w3));
    return (n0 >> 28) | ((n1 >> 24) & 0xF0u) | ((n2 >> 20) & 0xF00u) | ((n3 >> 16) & 0xF000u);
}

// ---- stage 2 ----------------------------------------------------------------------------------------------------
// 32 bitmap bits starting at tile position t (bit 0 = byte t).  bm needs one readable word past the last granule.
FG_DEV uint32_t r5_window(const uint32_t* bm, int t) {
    const int k = t >> 5;
    return fg_funnel_r(bm[k], bm[k + 1], (uint32_t)t & 31u);
}

// packed 8-byte side-table rows (positions relative to the line start, < 65536)
//   pair  : name_start | name_end << 16 | value_end << 32 | flags << 48   (value starts at name_end + 2)
//   header: sd_id start | sd_id end << 16 | #pairs << 32 | 0x8000 << 48
//   ext   : the row after a pair with FG_E8_ESC: arena offset | unescaped length << 32 (written by unescape5424_kernel)
constexpr unsigned long long kE8Esc = 1ull << 48;
constexpr unsigned long long kE8Header = 0x8000ull << 48;
FG_DEV unsigned long long r5_pack_pair8(int ns, int ne, int ve, bool esc) {
    return (unsigned long long)(uint32_t)ns | ((unsigned long long)(uint32_t)ne << 16) | ((unsigned long long)(uint32_t)ve << 32) |
           (esc ? kE8Esc : 0ull);
}
FG_DEV unsigned long long r5_pack_header8(int es, int id_end, uint32_t pairs) {
    return (unsigned long long)(uint32_t)es | ((unsigned long long)(uint32_t)id_end << 16) | ((unsigned long long)pairs << 32) | kE8Header;
}

struct R5Fast {
    double ts;
    uint32_t status, facility, severity, flags;
    int sp1, sp2, sp3, sp4, sp5;  // relative to the line start: hostname = [sp1+1, sp2), appname = [sp2+1, sp3), ...
    int msg_o, msg_l;             // msg span; msg_l == 0 => msg None and msg_o = end of full_msg
    uint32_t n_entries;           // 8-byte rows staged at stage[0 .. n_entries)
    unsigned long long* stage;
    bool wide;  // not representable / not stageable here: the wide kernel redoes this line
    bool esc;   // some value holds a backslash (the line goes on the unescape work list)
};

// unescape_sd_value (rfc5424_decoder.rs:105-125) over the raw value bytes v[0, len): `\"` `\\` `\]` drop the backslash,
// any other escaped byte keeps it, a trailing lone backslash is dropped.  Returns the unescaped length; writes when out != nullptr.
FG_DEV int r5_unescape(const uint8_t* v, int len, uint8_t* out) {
    int o = 0;
    bool esc = false;
    for (int k = 0; k < len; ++k) {
        const uint8_t c = v[k];
        if (!esc) {
            if (c == '\\') esc = true;
            else {
                if (out) out[o] = c;
                ++o;
            }
        } else {
            if (c != '"' && c != '\\' && c != ']') {
                if (out) out[o] = '\\';
                ++o;
            }
            if (out) out[o] = c;
            ++o;
            esc = false;
        }
    }
    return o;
}

constexpr uint32_t kFlagBom = 0x40u;   // FG_FLAG_BOM
constexpr uint32_t kFlagWide = 0x80u;  // FG_FLAG_WIDE

// T: tile bytes (shared memory), bm: its bitmap; the line is T[ls, le).  Idle lanes pass ls == le.
// ALL lanes of a warp must call this together.
FG_DEV void r5_walk(uint8_t* T, const uint32_t* bm, int ls, int le, R5Fast& r) {
    r.ts = 0.0;
    r.facility = 0xFFu;
    r.severity = 0xFFu;
    r.flags = 0;
    r.sp1 = r.sp2 = r.sp3 = r.sp4 = r.sp5 = 0;
    r.msg_o = r.msg_l = 0;
    r.n_entries = 0;
    r.stage = nullptr;
    r.wide = false;
    r.esc = false;
    uint32_t status = FG_ST_OK;

    // ---- BOM::parse :63-71 ---------------------------------------------------------------------
    int b = ls;
    if (le - ls >= 3 && T[ls] == 0xEFu && T[ls + 1] == 0xBBu && T[ls + 2] == 0xBFu) {
        b = ls + 3;
        r.flags |= kFlagBom;
    } else if (!(le > ls && T[ls] == '<')) {
        status = FG_E5_BOM;
    }

    // ---- splitn(7, ' ') :23 — the first six spaces, one bitmap hop each --------------------------
    int sp[6];
    {
        int t = b;
        bool act = status == FG_ST_OK;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            bool s = act;
            for (;;) {
                if (s) {
                    const uint32_t W = r5_window(bm, t);
                    t += W ? fg_ffs(W) - 1 : 32;
                    if (t >= le) {
                        s = false;
                        act = false;
                    } else if (W) {
                        if (T[t] == ' ') s = false;
                        else ++t;
                    }
                }
                if (!fg_any(s)) break;
            }
            sp[k] = act ? t : le;
            if (act) ++t;
        }
    }
    const int nsp = (sp[0] < le) + (sp[1] < le) + (sp[2] < le) + (sp[3] < le) + (sp[4] < le) + (sp[5] < le);

    // ---- parse_pri_version :74-92 on part0 = [b, sp0) --------------------------------------------
    if (status == FG_ST_OK) {
        const int e0 = sp[0];
        if (!(b < e0 && T[b] == '<')) {
            status = FG_E5_PRI_BRACKETS;
        } else {
            int gt = b + 1;
            while (gt < e0 && T[gt] != '>') ++gt;
            uint32_t pri = 0;
            if (!parse_u8(T, b + 1, gt, pri)) status = FG_E5_INVALID_PRI;
            else if (gt >= e0) status = FG_E5_MISSING_VERSION;
            else if (!(e0 - gt == 2 && T[gt + 1] == '1')) status = FG_E5_UNSUPPORTED_VERSION;
            else {
                r.facility = pri >> 3;
                r.severity = pri & 7u;
            }
        }
    }
    fg_syncwarp();
    // ---- timestamp :25, :94-103 ---------------------------------------------------------------
    if (status == FG_ST_OK) {
        if (nsp < 1) status = FG_E5_MISSING_TS;
        else if (!parse_rfc3339(T, sp[0] + 1, sp[1], r.ts)) status = FG_E5_BAD_TS;
        else if (nsp < 6) status = FG_E5_MISSING_HOST + (uint32_t)(nsp - 1);  // :26-30 in order
    }
    fg_syncwarp();

    // ---- parse_data :127-161 on part6 = [sp5+1, le) ------------------------------------------------
    const int d = sp[5] + 1;
    int msg_from = le;
    bool walk = false;
    if (status == FG_ST_OK) {
        if (d >= le) status = FG_E5_MISSING_MSG;  // :129
        else {
            const uint32_t c0 = T[d];
            if (c0 == '-') msg_from = d + 1;
            else if (c0 == '[') walk = true;
            else status = FG_E5_MALFORMED;  // :159
        }
    }
    {
        // Structured data, token-nested and in lock step: one outer iteration handles (per lane) either an sd_id or one
        // name="value" pair.  st_id: at the first byte of an sd_id (:175-177); !st_id: between params (!in_name, name None,
        // !in_value).  Rows are staged behind the cursor in the line's own bytes (slot k ends at sbase + 8 (k + 1)).
        const int sbase = (ls + 7) & ~7;
        unsigned long long* stg = (unsigned long long*)(T + sbase);
        uint32_t n = 1, pairs = 0, hdr = 0;
        int i = d + 1, elem_start = d + 1, id_end = 0;
        bool st_id = true, active = walk, wide = false, any_esc = false;
        while (fg_any(active)) {
            // (A) sd_id: up to the first ' ' (unvalidated, may contain ']')
            {
                const bool scan = active && st_id;
                bool s = scan;
                for (;;) {
                    if (s) {
                        const uint32_t W = r5_window(bm, i);
                        i += W ? fg_ffs(W) - 1 : 32;
                        if (i >= le) s = false;
                        else if (W) {
                            if (T[i] == ' ') s = false;
                            else ++i;
                        }
                    }
                    if (!fg_any(s)) break;
                }
                if (scan) {
                    if (i >= le) { active = false; status = FG_E5_MISSING_SD; }  // :177
                    else { id_end = i; ++i; st_id = false; }
                }
            }
            // (B) between params: skip ' ' and stray '"' (:194, :232)
            {
                bool s = active;
                for (;;) {
                    bool more = false;
                    if (s) {
                        if (i >= le) s = false;
                        else {
                            const uint32_t c = T[i];
                            more = (c == ' ') || (c == '"');
                            if (more) ++i;
                            else s = false;
                        }
                    }
                    if (!fg_any(more)) break;
                }
                if (active && i >= le) { active = false; status = FG_E5_SD_NO_END; }  // :239
            }
            // classify the byte that ended (B)
            bool do_name = false;
            int name_start = 0, name_end = 0;
            if (active) {
                const uint32_t c = T[i];
                if (c == ']') {  // :197 end of this element, then :145-155
                    if (sbase + 8 * (int)n > i + 1) { wide = true; active = false; }
                    else {
                        stg[hdr] = r5_pack_header8(elem_start - ls, id_end - ls, pairs);
                        if (i + 1 >= le) { active = false; status = FG_E5_MISSING_MSG; }  // :148
                        else {
                            const uint32_t c2 = T[i + 1];
                            if (c2 == '[') { elem_start = i + 2; i += 2; hdr = n++; pairs = 0; st_id = true; }
                            else if (c2 == ' ') { msg_from = i + 1; active = false; }
                            else { active = false; status = FG_E5_MALFORMED; }  // :154
                        }
                    }
                } else if (c >= 33u && c <= 126u && c != '=') {  // is_sd_name :188-192 ('"' and ']' excluded above)
                    do_name = true;
                    name_start = i;
                    ++i;
                } else {
                    active = false;
                    status = FG_E5_SD_FORMAT;  // :235
                }
            }
            // (C) name: hop to the next flagged byte; the flagged bytes that are legal name characters are stepped over
            {
                bool s = do_name;
                for (;;) {
                    if (s) {
                        const uint32_t W = r5_window(bm, i);
                        i += W ? fg_ffs(W) - 1 : 32;
                        if (i >= le) s = false;
                        else if (W) {
                            const uint32_t c = T[i];
                            if (c == '!' || c == '<' || c == '\\' || c == '|' || c == '}') ++i;
                            else s = false;
                        }
                    }
                    if (!fg_any(s)) break;
                }
            }
            bool do_val = false;
            if (do_name) {
                if (i >= le) { active = false; status = FG_E5_SD_NO_END; }
                else if (T[i] != '=') { active = false; status = FG_E5_SD_FORMAT; }
                else {
                    name_end = i;
                    ++i;
                    if (i >= le) { active = false; status = FG_E5_SD_NO_END; }
                    else if (T[i] != '"') { active = false; status = FG_E5_SD_FORMAT; }  // :212 is the only arm
                    else { ++i; do_val = true; }
                }
            }
            // (D) value: up to the first unescaped '"' (:216, :217, :231)
            bool has_bs = false;
            {
                bool s = do_val;
                for (;;) {
                    if (s) {
                        const uint32_t W = r5_window(bm, i);
                        i += W ? fg_ffs(W) - 1 : 32;
                        if (i >= le) s = false;
                        else if (W) {
                            const uint32_t c = T[i];
                            if (c == '"') s = false;
                            else if (c == '\\') { has_bs = true; i += 2; }  // the escaped byte is skipped
                            else ++i;
                        }
                    }
                    if (!fg_any(s)) break;
                }
            }
            if (do_val) {
                if (i >= le) { active = false; status = FG_E5_SD_NO_END; }
                else {
                    const int need = has_bs ? 2 : 1;
                    if (sbase + 8 * ((int)n + need) > i + 1) { wide = true; active = false; }
                    else {
                        stg[n] = r5_pack_pair8(name_start - ls, name_end - ls, i - ls, has_bs);
                        n += (uint32_t)need;
                        any_esc |= has_bs;
                        ++pairs;
                        ++i;
                    }
                }
            }
        }
        if (walk && status == FG_ST_OK && !wide) {
            r.n_entries = n;
            r.stage = stg;
            r.esc = any_esc;
        }
        r.wide = wide;
    }

    // ---- parse_msg :163-172, Record assembly :32-47 ------------------------------------------------
    if (status == FG_ST_OK) {
        // msg = line[msg_from..].trim(); full_msg = line.trim_end() of the BOM-stripped line (:46).  The byte before
        // msg_from is '-' or ']', so when the rest is all whitespace full_msg ends at msg_from.
        const int hi = trim_end(T, msg_from, le);
        const int lo = trim_start(T, msg_from, hi);
        if (hi > lo) {
            r.msg_o = lo - ls;
            r.msg_l = hi - lo;
        } else {
            r.msg_o = msg_from - ls;
            r.msg_l = 0;
        }
        r.sp1 = sp[1] - ls;
        r.sp2 = sp[2] - ls;
        r.sp3 = sp[3] - ls;
        r.sp4 = sp[4] - ls;
        r.sp5 = sp[5] - ls;
    }
    r.status = status;
    fg_syncwarp();
}

}  // namespace fg
