// fg_gelffast.cuh — GELF on the bitmap pipeline: one structural bitmap + one-member-per-step walk over a shared-memory tile,
// for REGULAR lines; everything else goes to the exact parser of fg_gelf.cuh.
//
// B200-native replacement for GelfDecoder::decode (/root/reference/src/flowgger/decoder/gelf_decoder.rs:34-125).  A regular
// line is what every GELF sender emits: ONE flat JSON object
//     { "key" : value , "key" : value ... }        value = string | number | true | false | null
// with nothing but spaces between the tokens, no escape inside a key, no raw control byte anywhere, at most
// kMaxLocalMembers members.  The walker proves that shape as it goes — with serde_json's own rules for what it accepts
// (string escapes incl. \uXXXX surrogate pairs: read.rs parse_escape; numbers: json_number of fg_gelf.cuh) — and the first
// byte that does not fit (a nested container, a TAB or LF between tokens, a raw control byte inside a string, any syntax
// error ...) hands the line to gelf_parse_line, which restates the whole grammar, the newline-retry of :44-46 and every
// error string.  A line accepted here is parsed to exactly the members the full parser would collect.
//
//   stage 1  gf_classify16: every thread takes 32-byte granules of the flat tile and writes one word of the bitmap
//            X = '"' | '\\' | byte < 0x20 (exact per byte): the only bytes that can end or alter a JSON string.
//   stage 2  gf_walk: one thread per line, ONE member per loop iteration for all 32 lines of a warp.  A string body is not
//            read: its end is the next X bit that is a quote (a per-lane word iterator skips 32 bytes per step), backslash
//            hits are validated where they stand.  Bytes BETWEEN tokens are read directly (a handful per member).
//   phase 2  (gelf_finish_local, shared with the exact parser) sorts the members by key, keeps the last duplicate and
//            applies the per-key rules of gelf_decoder.rs:51-107.
#pragma once
#include "fg_common.cuh"
#include "fg_gelf.cuh"
#include "fg_ltsvfast.cuh"  // lt_eq_flags, lt_load8
#include "fg_r5fast.cuh"
#include "fg_status.h"

namespace fg {

// ---- stage 1 ----------------------------------------------------------------------------------------------------
FG_DEV uint32_t gf_flags(uint32_t w) {
    const uint32_t t = w & 0xE0E0E0E0u;                                             // 0 iff the byte is < 0x20
    const uint32_t ctrl = ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u;
    return lt_eq_flags(w, 0x22222222u) | lt_eq_flags(w, 0x5C5C5C5Cu) | ctrl;
}
FG_DEV uint32_t gf_classify16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    return r5_gather16(gf_flags(w0), gf_flags(w1), gf_flags(w2), gf_flags(w3));
}

// ---- stage 2 ----------------------------------------------------------------------------------------------------
FG_DEV bool gf_hex4(const uint8_t* T, int at, uint32_t& n) {
    n = 0;
    for (int k = 0; k < 4; ++k) {
        const uint32_t c = T[at + k];
        uint32_t v;
        if (c - '0' <= 9u) v = c - '0';
        else if ((c | 0x20u) - 'a' <= 5u) v = (c | 0x20u) - 'a' + 10u;
        else return false;
        n = n * 16u + v;
    }
    return true;
}
// h = position of a backslash inside a string: position after the whole escape sequence (read.rs parse_escape), or -1 when
// serde_json rejects it (the exact parser then reports the error)
FG_DEV int gf_escape_end(const uint8_t* T, int h, int le) {
    if (h + 1 >= le) return -1;
    const uint32_t x = T[h + 1];
    if (x == '"' || x == '\\' || x == '/' || x == 'b' || x == 'f' || x == 'n' || x == 'r' || x == 't') return h + 2;
    if (x != 'u') return -1;
    uint32_t n1;
    if (h + 6 > le || !gf_hex4(T, h + 2, n1)) return -1;
    if (n1 >= 0xDC00u && n1 <= 0xDFFFu) return -1;  // LoneLeadingSurrogateInHexEscape
    if (n1 >= 0xD800u && n1 <= 0xDBFFu) {
        uint32_t n2;
        if (h + 12 > le || T[h + 6] != '\\' || T[h + 7] != 'u' || !gf_hex4(T, h + 8, n2)) return -1;
        if (n2 < 0xDC00u || n2 > 0xDFFFu) return -1;
        return h + 12;
    }
    return h + 6;
}

// A JSON string body starting at s (the byte after the opening quote).  All lanes call; lanes with act = true scan.  Returns
// the position of the closing quote, or -1 (not a regular string: no closing quote inside the line, a rejected escape, a
// raw control byte).  has_bs = the body holds an escape.
FG_DEV int gf_string(const uint8_t* T, const uint32_t* bmX, int s, int le, bool act, bool& has_bs) {
    int xw = s >> 5;
    uint32_t xrem = act ? bmX[xw] & (0xFFFFFFFFu << (s & 31)) : 0u;
    int end = -1;
    bool more = act;
    has_bs = false;
    while (fg_any(more)) {
        for (;;) {  // next X bit: 32 bytes of string body per step
            const bool need = more && xrem == 0u && ((xw + 1) << 5) < le;
            if (!fg_any(need)) break;
            if (need) {
                ++xw;
                xrem = bmX[xw];
            }
        }
        if (more) {
            const int h = (xw << 5) + fg_ffs(xrem) - 1;
            if (xrem == 0u || h >= le) {
                more = false;  // EOFWhileParsingString
            } else {
                const uint32_t c = T[h];
                if (c == '"') {
                    end = h;
                    more = false;
                } else if (c == '\\') {
                    has_bs = true;
                    const int nx = gf_escape_end(T, h, le);
                    if (nx < 0) {
                        more = false;
                    } else {  // the iterator continues behind the escape sequence (which may hold X bytes itself: \" \\)
                        if ((nx >> 5) != xw) {
                            xw = nx >> 5;
                            xrem = bmX[xw];
                        }
                        xrem &= 0xFFFFFFFFu << (nx & 31);
                    }
                } else {
                    more = false;  // raw control byte: InvalidUnicodeCodePoint -> newline retry / error (exact parser)
                }
            }
        }
    }
    return end;
}

FG_DEV void gf_skip_spaces(const uint8_t* T, int& pos, int le, bool act) {
    for (;;) {
        const bool more = act && pos < le && T[pos] == ' ';
        if (!fg_any(more)) break;
        if (more) ++pos;
    }
}

// one of the six keys gelf_decoder.rs:51-96 consumes (raw bytes; keys with escapes never get here)
FG_DEV bool gf_is_reserved(const uint8_t* k, int len) {
    const unsigned long long a = lt_load8(k), b = lt_load8(k + 8);
    switch (len) {
        case 4: return (uint32_t)a == 0x74736F68u;                                                        // host
        case 5: return (a & 0x000000FFFFFFFFFFull) == 0x0000006C6576656Cull;                              // level
        case 7: return (a & 0x00FFFFFFFFFFFFFFull) == 0x006E6F6973726576ull;                              // version
        case 9: return a == 0x6D617473656D6974ull && (b & 0xFFull) == 0x70ull;                            // timestam|p
        case 12: return a == 0x73656D5F6C6C7566ull && (uint32_t)b == 0x65676173u;                         // full_mes|sage
        case 13: return a == 0x656D5F74726F6873ull && (b & 0x000000FFFFFFFFFFull) == 0x0000006567617373ull;  // short_me|ssage
        default: return false;
    }
}

// All 32 lanes call (idle lanes with active_line = false).  T = the tile, [ls, le) the line inside it, line_off = the
// absolute offset of the line (spans in M are absolute, like the exact parser's).  Returns true when the line is regular:
// M then holds its members in document order and n_plain counts those that are not reserved keys (an upper bound of the
// side-table rows).  false: the line must be parsed by gelf_parse_line.
FG_DEV bool gf_walk(const uint8_t* T, const uint32_t* bmX, int ls, int le, bool active_line, int line_off, Members& M, uint32_t& n_plain) {
    M.m = 0;
    M.spilled = false;
    n_plain = 0;
    bool reg = active_line;
    int pos = ls;
    gf_skip_spaces(T, pos, le, reg);
    if (reg) {
        if (pos < le && T[pos] == '{') ++pos;
        else reg = false;
    }
    gf_skip_spaces(T, pos, le, reg);
    bool open = reg;  // inside the object: a key comes next
    if (reg && pos < le && T[pos] == '}') {
        ++pos;
        open = false;
    }
    while (fg_any(reg && open)) {
        const bool run = reg && open;
        // "key"
        const bool k_ok = run && pos < le && T[pos] == '"';
        const int key_s = pos + 1;
        bool kbs = false;
        const int key_e = gf_string(T, bmX, key_s, le, k_ok, kbs);
        if (run) {
            if (!k_ok || key_e < 0 || kbs) reg = false;  // KeyMustBeAString / escapes in a key: exact parser
            else pos = key_e + 1;
        }
        gf_skip_spaces(T, pos, le, run && reg);
        if (run && reg) {
            if (pos < le && T[pos] == ':') ++pos;
            else reg = false;
        }
        gf_skip_spaces(T, pos, le, run && reg);
        // value
        if (run && reg && pos >= le) reg = false;
        const uint32_t c = (run && reg) ? T[pos] : 0u;
        const bool is_str = run && reg && c == '"';
        const int vs = pos + 1;
        bool vbs = false;
        const int ve = gf_string(T, bmX, vs, le, is_str, vbs);
        uint32_t tag = JT_NULL, vflags = 0;
        uint64_t bits = 0;
        if (run && reg) {
            if (is_str) {
                if (ve < 0) {
                    reg = false;
                } else {
                    tag = JT_STRING;
                    bits = (uint64_t)(uint32_t)(line_off + (vs - ls)) | ((uint64_t)(uint32_t)(ve - vs) << 32);
                    vflags = vbs ? 0x08u : 0u;  // FG_EM_UNESCAPE
                    pos = ve + 1;
                }
            } else {
                Json j;
                j.p = T;
                j.len = le;
                j.i = pos;
                j.mode2 = false;
                if (c == '-') {
                    ++j.i;
                    if (!json_number(j, false, tag, bits)) reg = false;
                } else if (c - '0' <= 9u) {
                    if (!json_number(j, true, tag, bits)) reg = false;
                } else if (c == 'n') {
                    ++j.i;
                    if (!json_lit(j, "ull", 3)) reg = false;
                    tag = JT_NULL;
                } else if (c == 't') {
                    ++j.i;
                    if (!json_lit(j, "rue", 3)) reg = false;
                    tag = JT_BOOL;
                    bits = 1;
                } else if (c == 'f') {
                    ++j.i;
                    if (!json_lit(j, "alse", 4)) reg = false;
                    tag = JT_BOOL;
                    bits = 0;
                } else {
                    reg = false;  // a container as a member value, or ExpectedSomeValue: exact parser
                }
                pos = j.i;
            }
        }
        if (run && reg) {
            if (M.m >= (uint32_t)kMaxLocalMembers) {
                reg = false;
            } else {
                M.name[M.m] = make_int2(line_off + (key_s - ls), key_e - key_s);
                M.val[M.m] = bits;
                M.meta[M.m] = (uint8_t)(tag | vflags);
                ++M.m;
                if (!gf_is_reserved(T + key_s, key_e - key_s)) ++n_plain;
            }
        }
        gf_skip_spaces(T, pos, le, run && reg);
        if (run && reg) {
            const uint32_t d = pos < le ? T[pos] : 0u;
            if (d == ',') ++pos;
            else if (d == '}') { ++pos; open = false; }
            else reg = false;  // ExpectedObjectCommaOrEnd
        }
        gf_skip_spaces(T, pos, le, run && reg && open);
    }
    gf_skip_spaces(T, pos, le, reg);
    if (reg && pos != le) reg = false;  // TrailingCharacters
    return reg;
}

}  // namespace fg
