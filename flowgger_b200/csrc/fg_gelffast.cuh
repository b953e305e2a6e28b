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
//   stage 2  gf_walk: one thread per line, ONE member per loop iteration for all 32 lines of a warp — but only to find
//            where the members are: key span, value span, string or not.  A string body is not read: its end is the next X
//            bit that is a quote (a per-lane word iterator skips 32 bytes per step; a backslash hit skips its byte).  Bytes
//            BETWEEN tokens are read directly (a handful per member).
//   members  gf_member: one thread per MEMBER, for all members of the CTA round, 256 threads wide: the escapes of a string
//            value are validated, a number goes through json_number, the reserved keys are recognised.  (Doing this inside
//            the walk ran ~3 of 32 lanes wide: 1258 warp-instructions per line, profiles/r2_notes.md.)
//   phase 2  (gelf_finish_local, shared with the exact parser) one thread per line sorts the members by key, keeps the last
//            duplicate and applies the per-key rules of gelf_decoder.rs:51-107.
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
// the position of the closing quote, or -1 (no closing quote inside the line, a raw control byte).  A backslash skips the
// byte behind it, whatever it is: gf_member validates the escapes (a line with an invalid one is not regular, so the
// shortcut cannot change what a regular line parses to).  has_bs = the body holds a backslash.
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
                    const int nx = h + 2;
                    if (nx > le) {
                        more = false;
                    } else {  // the iterator continues behind the escaped byte (which may be an X byte itself: \" \\)
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

// GKEY_* of a raw key (keys with escapes never get here): the six keys gelf_decoder.rs:51-96 consumes, else GKEY_OTHER
FG_DEV int gf_key_kind(const uint8_t* k, int len) {
    const unsigned long long a = lt_load8(k), b = lt_load8(k + 8);
    switch (len) {
        case 4: return (uint32_t)a == 0x74736F68u ? GKEY_HOST : GKEY_OTHER;
        case 5: return (a & 0x000000FFFFFFFFFFull) == 0x0000006C6576656Cull ? GKEY_LEVEL : GKEY_OTHER;
        case 7: return (a & 0x00FFFFFFFFFFFFFFull) == 0x006E6F6973726576ull ? GKEY_VERSION : GKEY_OTHER;
        case 9: return (a == 0x6D617473656D6974ull && (b & 0xFFull) == 0x70ull) ? GKEY_TIMESTAMP : GKEY_OTHER;              // timestam|p
        case 12: return (a == 0x73656D5F6C6C7566ull && (uint32_t)b == 0x65676173u) ? GKEY_FULL : GKEY_OTHER;                // full_mes|sage
        case 13: return (a == 0x656D5F74726F6873ull && (b & 0x000000FFFFFFFFFFull) == 0x0000006567617373ull) ? GKEY_SHORT : GKEY_OTHER;  // short_me|ssage
        default: return GKEY_OTHER;
    }
}
// the first 8 key bytes as a big-endian number, zero-padded: orders like the bytes do (key bytes are never 0)
FG_DEV unsigned long long gf_key_prefix(const uint8_t* k, int len) {
    unsigned long long a = lt_load8(k);
    if (len < 8) a &= (1ull << (8 * len)) - 1ull;
    const uint32_t lo = (uint32_t)a, hi = (uint32_t)(a >> 32);
#ifdef FG_HOST_EMU
    return ((unsigned long long)__builtin_bswap32(lo) << 32) | __builtin_bswap32(hi);
#else
    return ((unsigned long long)__byte_perm(lo, 0u, 0x0123) << 32) | __byte_perm(hi, 0u, 0x0123);
#endif
}

// where the members of a line are (tile positions, < 65536): key [ks, ke), value [vs, ve) — for a string the body
constexpr uint32_t GK_TOKEN = 0, GK_STRING = 1, GK_STRING_BS = 2;
struct GfSpans {
    unsigned long long sp[kMaxLocalMembers];  // ks | ke << 16 | vs << 32 | ve << 48
    uint8_t kind[kMaxLocalMembers];
    uint32_t m;
};

// All 32 lanes call (idle lanes with active_line = false).  T = the tile, [ls, le) the line inside it.  Returns true when
// the line has the regular SHAPE (its members are in G, in document order); gf_member still has to accept every member.
FG_DEV bool gf_walk(const uint8_t* T, const uint32_t* bmX, int ls, int le, bool active_line, GfSpans& G) {
    G.m = 0;
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
        // "key": the next X byte behind the opening quote must be the closing one (no escapes in keys, <= 32 bytes)
        int key_s = pos + 1, key_e = 0;
        if (run) {
            const bool k_ok = pos < le && T[pos] == '"';
            const uint32_t W = k_ok ? r5_window(bmX, key_s) : 0u;
            const int h = key_s + fg_ffs(W) - 1;
            if (W == 0u || h >= le || T[h] != '"') reg = false;  // KeyMustBeAString, escapes / control bytes, long keys: exact parser
            else { key_e = h; pos = h + 1; }
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
        const int vs = is_str ? pos + 1 : pos;
        bool vbs = false;
        int ve = gf_string(T, bmX, vs, le, is_str, vbs);
        if (run && reg && is_str) {
            if (ve < 0) reg = false;
            else pos = ve + 1;
        }
        // any other token runs up to the next ',', '}' or space
        const bool is_tok = run && reg && !is_str;
        int te = pos;
        for (;;) {
            bool more = false;
            if (is_tok && te < le) {
                const uint32_t d = T[te];
                more = !(d == ',' || d == '}' || d == ' ');
            }
            if (!fg_any(more)) break;
            if (more) ++te;
        }
        if (is_tok) {
            // number, true, false, null; a container as a member value or anything else: exact parser
            if (te == pos || !(c == '-' || c - '0' <= 9u || c == 't' || c == 'f' || c == 'n')) reg = false;
            ve = te;
            pos = te;
        }
        if (run && reg) {
            if (G.m >= (uint32_t)kMaxLocalMembers) {
                reg = false;
            } else {
                G.sp[G.m] = (unsigned long long)(uint32_t)key_s | ((unsigned long long)(uint32_t)key_e << 16) |
                            ((unsigned long long)(uint32_t)vs << 32) | ((unsigned long long)(uint32_t)ve << 48);
                G.kind[G.m] = (uint8_t)(is_str ? (vbs ? GK_STRING_BS : GK_STRING) : GK_TOKEN);
                ++G.m;
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

// One member (any thread): validates the value the way serde_json does and produces the member as the exact parser would
// (name span absolute, value, tag | flags).  false: the line is not regular after all.  plain = not a reserved key.
FG_DEV bool gf_member(const uint8_t* T, const uint32_t* bmX, unsigned long long sp, uint32_t kind, int line_off, int ls, int2& name,
                      unsigned long long& val, uint32_t& meta, bool& plain) {
    const int ks = (int)(sp & 0xFFFFu), ke = (int)((sp >> 16) & 0xFFFFu), vs = (int)((sp >> 32) & 0xFFFFu), ve = (int)(sp >> 48);
    name = make_int2(line_off + (ks - ls), ke - ks);
    const int kk = gf_key_kind(T + ks, ke - ks);
    plain = kk == GKEY_OTHER;
    const uint32_t kbits = (uint32_t)kk << 5;  // bits 5..7 of the member's meta: the key class, for phase 2
    if (kind != GK_TOKEN) {
        if (kind == GK_STRING_BS) {  // every backslash of the body starts a valid escape (read.rs parse_escape)
            int pos = vs;
            for (;;) {
                int xw = pos >> 5;
                uint32_t xrem = bmX[xw] & (0xFFFFFFFFu << (pos & 31));
                while (xrem == 0u && ((xw + 1) << 5) < ve) xrem = bmX[++xw];
                const int h = (xw << 5) + fg_ffs(xrem) - 1;
                if (xrem == 0u || h >= ve) break;
                pos = gf_escape_end(T, h, ve);  // T[h] is a backslash: quotes inside the body are escaped, control bytes excluded
                if (pos < 0) return false;
            }
        }
        val = (unsigned long long)(uint32_t)(line_off + (vs - ls)) | ((unsigned long long)(uint32_t)(ve - vs) << 32);
        meta = JT_STRING | (kind == GK_STRING_BS ? 0x08u : 0u) | kbits;  // FG_EM_UNESCAPE
        return true;
    }
    Json j;
    j.p = T;
    j.len = ve;
    j.i = vs;
    j.mode2 = false;
    const uint32_t c = T[vs];
    uint32_t tag = JT_NULL;
    uint64_t bits = 0;
    bool ok;
    if (c == '-') {
        ++j.i;
        ok = json_number(j, false, tag, bits);
    } else if (c - '0' <= 9u) {
        ok = json_number(j, true, tag, bits);
    } else if (c == 'n') {
        ++j.i;
        ok = json_lit(j, "ull", 3);
    } else if (c == 't') {
        ++j.i;
        ok = json_lit(j, "rue", 3);
        tag = JT_BOOL;
        bits = 1;
    } else {
        ++j.i;
        ok = json_lit(j, "alse", 4);
        tag = JT_BOOL;
    }
    if (!ok || j.i != ve) return false;
    val = bits;
    meta = tag | kbits;
    return true;
}


// Phase 2 for members produced by gf_member (raw keys, key class in bits 5..7 of meta): BTreeMap order = stable insertion
// sort of an index permutation — by the 8-byte key prefix, the bytes only when two prefixes tie —, the last duplicate wins,
// then the rule of the key class.  Same result as gelf_finish_local (fg_gelf.cuh) on the same members.
FG_DEV void gf_finish(bytes_t p, int line_off, const Members& M, LineResult& r, GelfAcc& g, const EntrySink& sink, uint32_t sbase) {
    const uint32_t m = M.m;
    unsigned long long pre[kMaxLocalMembers];
    uint8_t ord[kMaxLocalMembers];
    auto cmp = [&](uint32_t x, uint32_t y) {  // key x vs key y
        if (pre[x] != pre[y]) return pre[x] < pre[y] ? -1 : 1;
        const int2 a = M.name[x], b = M.name[y];
        if (a.y <= 8 && b.y <= 8) return 0;  // both keys are their prefix
        return raw_key_cmp(p, a.x - line_off, a.x - line_off + a.y, b.x - line_off, b.x - line_off + b.y);
    };
    for (uint32_t a = 0; a < m; ++a) {
        pre[a] = gf_key_prefix(p + (M.name[a].x - line_off), M.name[a].y);
        int b = (int)a - 1;
        while (b >= 0 && cmp(ord[b], a) > 0) {
            ord[b + 1] = ord[b];
            --b;
        }
        ord[b + 1] = (uint8_t)a;
    }
    uint32_t gi = 0;
    while (gi < m && g.status == FG_ST_OK) {
        uint32_t ge = gi + 1;  // group of equal keys [gi, ge): the last inserted value wins
        while (ge < m && cmp(ord[gi], ord[ge]) == 0) ++ge;
        const uint32_t w = ord[ge - 1];
        gelf_apply_kind((int)(M.meta[w] >> 5), p, line_off, false, M.name[w], M.val[w], M.meta[w] & 0x0Fu, r, g, sink, sbase);
        gi = ge;
    }
}

}  // namespace fg
