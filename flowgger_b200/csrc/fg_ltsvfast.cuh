// fg_ltsvfast.cuh — LTSV on the bitmap pipeline: two structural bitmaps + one-part-per-step walk over a shared-memory tile.
//
// B200-native replacement for LTSVDecoder::decode (/root/reference/src/flowgger/decoder/ltsv_decoder.rs:87-221); the
// value parsers (parse_ts :263-267, the typed schema values :138-195) are the ones of fg_ltsv.cuh, called on tile bytes.
//
//   stage 1  lt_classify16: every thread takes 32-byte granules of the flat tile (two LDS.128, all 32 lanes busy) and writes
//            one word each of two bitmaps, exact per byte: T = TAB (line.split('\t') :94), C = ':' (splitn(2, ':') :95).
//   stage 2  ltsv_walk: one thread per line, ONE tab-separated part per loop iteration for all 32 lines of a warp.  The end
//            of the part is the next set bit of T (a per-lane iterator: word + remaining bits, no byte is read), the key
//            ends at the first set bit of C inside the part (one funnel-shifted 32-bit window).  What the round-1 scanner
//            found with 16-byte SWAR scans per lane and per part (645 warp-instructions per line, 3.5x read amplification
//            from 32 lanes pulling 32 different lines through L1) costs a find-first-set here.
//
// Side-table rows are staged as 8-byte packed entries in shared memory — the line's slots are reserved from its TAB
// count (#parts = #tabs + 1 >= #pairs), known from T before the walk — and leave the SM as coalesced column stores.
#pragma once
#include "fg_common.cuh"
#include "fg_ltsv.cuh"
#include "fg_r5fast.cuh"
#include "fg_status.h"

namespace fg {

// ---- stage 1 ----------------------------------------------------------------------------------------------------
// 0x80 in every byte of w that equals the byte replicated in pat; exact per byte (no cross-byte carries)
FG_DEV uint32_t lt_eq_flags(uint32_t w, uint32_t pat) {
    const uint32_t x = w ^ pat;
    const uint32_t y = ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x;  // bit 7: byte != 0
    return ~y & 0x80808080u;
}
FG_DEV void lt_classify16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t& t16, uint32_t& c16) {
    t16 = r5_gather16(lt_eq_flags(w0, 0x09090909u), lt_eq_flags(w1, 0x09090909u), lt_eq_flags(w2, 0x09090909u), lt_eq_flags(w3, 0x09090909u));
    c16 = r5_gather16(lt_eq_flags(w0, 0x3A3A3A3Au), lt_eq_flags(w1, 0x3A3A3A3Au), lt_eq_flags(w2, 0x3A3A3A3Au), lt_eq_flags(w3, 0x3A3A3A3Au));
}

// TABs inside tile range [ls, le)
FG_DEV int lt_count_tabs(const uint32_t* bmT, int ls, int le) {
    if (le <= ls) return 0;
    const int w0 = ls >> 5, w1 = (le - 1) >> 5;
    int cnt = 0;
    for (int w = w0; w <= w1; ++w) {
        uint32_t m = bmT[w];
        if (w == w0) m &= 0xFFFFFFFFu << (ls & 31);
        if (w == w1 && (le & 31)) m &= 0xFFFFFFFFu >> (32 - (le & 31));
        cnt += __popc(m);
    }
    return cnt;
}

// staged side-table row: key start (tile-relative) | key length << 16 | value length << 32 | meta << 56; the value starts
// one byte after the key (the colon).  meta = fg_ltsv_type | FG_EM_SUFFIX; typed values sit in the parallel `stage_val`.
FG_DEV unsigned long long lt_pack_entry(int ka, int kn, int vl, uint32_t meta) {
    return (unsigned long long)(uint32_t)ka | ((unsigned long long)(uint32_t)kn << 16) | ((unsigned long long)(uint32_t)vl << 32) |
           ((unsigned long long)meta << 56);
}

// ---- stage 2 ----------------------------------------------------------------------------------------------------
// the first 8 bytes at p (any alignment), little-endian; reads up to 11 bytes past p (tile + bitmaps are contiguous)
FG_DEV unsigned long long lt_load8(const uint8_t* p) {
#ifdef FG_HOST_EMU
    unsigned long long v;
    memcpy(&v, p, 8);
    return v;
#else
    const uint32_t sh = ((uint32_t)(size_t)p & 3u) * 8u;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p - ((size_t)p & 3u));
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    return (unsigned long long)__funnelshift_r(w0, w1, sh) | ((unsigned long long)__funnelshift_r(w1, w2, sh) << 32);
#endif
}

// The schema (ltsv_decoder.rs:19-22) as the walker reads it: from shared memory when the kernel staged it (`names` etc.
// then point into the CTA's copy), with two cheap filters in front of the byte compares — a key can only be a schema key
// if its length and its first byte occur in the schema.
struct LtsvSchemaView {
    const uint8_t* names;
    const int32_t* name_off;
    const int32_t* types;
    int n;
    uint32_t len_mask;        // bit min(len, 31)
    const uint32_t* first;    // [8] 256-bit set of first bytes (nullptr: no filter)
};
// the two filters of a schema (one thread)
FG_DEV void lt_schema_filters(const uint8_t* names, const int32_t* name_off, int n, uint32_t& len_mask, uint32_t* first /* [8] */) {
    len_mask = 0;
    for (int k = 0; k < 8; ++k) first[k] = 0;
    for (int k = 0; k < n; ++k) {
        const int o = name_off[k], l = name_off[k + 1] - o;
        len_mask |= 1u << (l < 31 ? l : 31);
        if (l > 0) first[names[o] >> 5] |= 1u << (names[o] & 31u);
    }
}
FG_DEV int lt_schema_type(bytes_t p, int a, int n, unsigned long long k8, const LtsvSchemaView& S) {
    if (!((S.len_mask >> (n < 31 ? n : 31)) & 1u)) return 0;
    if (S.first != nullptr && n > 0) {
        const uint32_t c = (uint32_t)(k8 & 0xFFu);
        if (!((S.first[c >> 5] >> (c & 31u)) & 1u)) return 0;
    }
    for (int k = 0; k < S.n; ++k) {
        const int o = S.name_off[k], l = S.name_off[k + 1] - o;
        if (l != n) continue;
        bool eq = true;
        for (int j = 0; j < n && eq; ++j) eq = S.names[o + j] == p[a + j];
        if (eq) return S.types[k];
    }
    return 0;
}

// All 32 lanes of a warp must call this (idle lanes with active_line = false).  T = the tile, [ls, le) the line inside it
// (le - ls < 65536).  `stage` / `stage_val` = this line's reserved slots (>= #tabs + 1).  Result spans are relative to ls,
// exactly as ltsv_parse_line (fg_ltsv.cuh) reports them.
//
// The part loop holds NO per-key work that only a few lanes need: the four reserved keys are recognised branch-free from
// the first 8 key bytes; `host` / `message` are two predicated moves; the values of `time`, `level` and of typed schema
// keys are only PARKED (position, length, part) and parsed after the loop in lock step — one phase per kind, so all lanes
// run the same parser.  (The first version parsed them where they stood: ~2 of 32 lanes active, 386 warp-instructions per
// line, profiles/r2_notes.md.)  Evaluation order is preserved: the reference returns at the FIRST failing part, so a parked
// item is evaluated iff no error was found at an earlier part, and an earlier failure replaces a later one.
template <bool TYPED>
FG_DEV void ltsv_walk(const uint8_t* T, const uint32_t* bmT, const uint32_t* bmC, int ls, int le, bool active_line,
                      const LtsvDeviceConfig& cfg, const LtsvSchemaView& S, LineResult& r, unsigned long long* stage,
                      unsigned long long* stage_val) {
    r.ts = 0.0;
    r.facility = 0xFFu;
    r.severity = 0xFFu;
    r.flags = 0;
    r.host_o = r.app_o = r.proc_o = r.mid_o = r.msg_o = r.full_o = -1;
    r.host_l = r.app_l = r.proc_l = r.mid_l = r.msg_l = r.full_l = 0;
    r.n_entries = 0;
    uint32_t status = FG_ST_OK, n = 0, flags = 0;
    bool have_ts = false;
    int err_pos = 0;  // tile position of the failing part
    bool err_set = false;
    int part = ls;  // start of the current part
    bool active = active_line;
    // iterator over the TABs at or after `part`
    int tw = ls >> 5;
    uint32_t trem = active ? bmT[tw] & (0xFFFFFFFFu << (ls & 31)) : 0u;
    // parked values: `time`, `level`, and two typed values per schema type (slot = 2 * (type - 1) + {0, 1})
    int ts_a = -1, ts_b = -1, ts_part = 0;
    int lv_a = -1, lv_b = -1, lv_part = 0;
    int pk_va[8], pk_pt[8];
    uint32_t pk_info[8];  // value length | row << 16; 0xFFFFFFFF = empty
#pragma unroll
    for (int q = 0; q < 8; ++q) { pk_va[q] = 0; pk_pt[q] = 0; pk_info[q] = 0xFFFFFFFFu; }
    const bytes_t p = T;
    while (fg_any(active)) {  // line.split('\t') :94
        // end of this part: the next TAB below le, else le
        for (;;) {
            const bool need = active && trem == 0u && ((tw + 1) << 5) < le;
            if (!fg_any(need)) break;
            if (need) {
                ++tw;
                trem = bmT[tw];
            }
        }
        int pe = le;
        if (active && trem) {
            const int t = (tw << 5) + fg_ffs(trem) - 1;
            if (t < le) {
                pe = t;
                trem &= trem - 1u;
            }
        }
        // splitn(2, ':') :95 — the first ':' of the part
        int cpos = part;
        uint32_t wc = active ? r5_window(bmC, part) : 0u;
        for (;;) {  // keys longer than 32 bytes, or long parts without a colon
            const bool need = active && wc == 0u && cpos + 32 < pe;
            if (!fg_any(need)) break;
            if (need) {
                cpos += 32;
                wc = r5_window(bmC, cpos);
            }
        }
        int colon = -1;
        if (active && wc) {
            const int c = cpos + fg_ffs(wc) - 1;
            if (c < pe) colon = c;
        }
        if (active) {
            if (colon < 0) {
                flags |= 0x02u;  // FG_FLAG_MISSING_VALUE: println! at :99 is replayed by the host
            } else {
                const int ka = part, kn = colon - part, va = colon + 1, vb = pe;
                const unsigned long long k8 = lt_load8(p + ka);
                const uint32_t k4 = (uint32_t)k8;
                const bool is_time = kn == 4 && k4 == 0x656D6974u;                                       // "time"  :104
                const bool is_host = kn == 4 && k4 == 0x74736F68u;                                       // "host"
                const bool is_msg = kn == 7 && (k8 & 0x00FFFFFFFFFFFFFFull) == 0x006567617373656Dull;    // "message"
                const bool is_level = kn == 5 && (k8 & 0x000000FFFFFFFFFFull) == 0x0000006C6576656Cull;  // "level"  :114
                if (is_host) { r.host_o = va - ls; r.host_l = vb - va; }
                if (is_msg) { r.msg_o = va - ls; r.msg_l = vb - va; }
                if (is_time) {
                    if (ts_a >= 0) {  // an earlier `time` is still parked: it is evaluated first (a failure returns there)
                        int a = ts_a, b = ts_b;
                        if (b - a >= 2 && p[a] == '[' && p[b - 1] == ']') { ++a; --b; }
                        if (ltsv_parse_ts(p, a, b, r.ts)) have_ts = true;
                        else { status = FG_EL_TS; err_pos = ts_part; err_set = true; }
                    }
                    ts_a = status == FG_ST_OK ? va : -1;
                    ts_b = vb;
                    ts_part = part;
                } else if (is_level) {
                    if (lv_a >= 0) {  // same for an earlier `level` (:114-121)
                        uint32_t sev;
                        if (!parse_u8(p, lv_a, lv_b, sev)) { status = FG_EL_SEV; err_pos = lv_part; err_set = true; }
                        else if (sev > 7u) { status = FG_EL_SEV_HIGH; err_pos = lv_part; err_set = true; }
                        else r.severity = sev;
                    }
                    lv_a = status == FG_ST_OK ? va : -1;
                    lv_b = vb;
                    lv_part = part;
                } else if (!is_host && !is_msg) {  // :122-199
                    uint32_t meta = 0;
                    bool deferred = false;
                    unsigned long long val = 0;
                    if (TYPED) {
                        const int type = lt_schema_type(p, ka, kn, k8, S);
                        if (type != 0) {
                            meta = (uint32_t)type;  // FG_TAG_* == fg_ltsv_type
                            if (ltsv_needs_suffix(p, ka, kn, type, cfg)) meta |= 0x20u;  // FG_EM_SUFFIX
                            int q = -1;  // first free slot of this type (static indices: the arrays stay in registers)
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                                if (q < 0 && (u >> 1) == type - 1 && pk_info[u] == 0xFFFFFFFFu) q = u;
                            if (q >= 0 && n < 65535u) {
#pragma unroll
                                for (int u = 0; u < 8; ++u)
                                    if (u == q) { pk_va[u] = va; pk_pt[u] = part; pk_info[u] = (uint32_t)(vb - va) | (n << 16); }
                                deferred = true;
                            } else {
                                status = ltsv_parse_typed(p, va, vb, type, val);
                            }
                        }
                    }
                    if (status == FG_ST_OK) {
                        stage[n] = lt_pack_entry(ka, kn, vb - va, meta);
                        if (TYPED && !deferred) stage_val[n] = val;
                        ++n;
                    }
                }
            }
            if (status != FG_ST_OK) {
                if (!err_set) err_pos = part;  // the error belongs to the current part (unless an earlier parked value failed)
                active = false;
            } else if (pe >= le) {
                active = false;
            } else {
                part = pe + 1;
            }
        }
    }
    // Parked work, in lock step.  All of it is pure, so it is evaluated even when a later part already failed.
    {
        const bool pend = active_line && ts_a >= 0 && (status == FG_ST_OK || ts_part < err_pos);
        if (fg_any(pend)) {
            if (pend) {
                int a = ts_a, b = ts_b;
                if (b - a >= 2 && p[a] == '[' && p[b - 1] == ']') { ++a; --b; }  // :105-109
                double t;
                if (ltsv_parse_ts(p, a, b, t)) { r.ts = t; have_ts = true; }
                else { status = FG_EL_TS; err_pos = ts_part; }
            }
        }
    }
    {
        const bool pend = active_line && lv_a >= 0 && (status == FG_ST_OK || lv_part < err_pos);
        if (fg_any(pend)) {
            if (pend) {
                uint32_t sev;
                if (!parse_u8(p, lv_a, lv_b, sev)) { status = FG_EL_SEV; err_pos = lv_part; }
                else if (sev > 7u) { status = FG_EL_SEV_HIGH; err_pos = lv_part; }
                else r.severity = sev;
            }
        }
    }
    if (TYPED) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {  // slots 2t, 2t+1 hold values of schema type t + 1: one parser per phase
            const uint32_t info = pk_info[q];
            const bool has = active_line && info != 0xFFFFFFFFu && (status == FG_ST_OK || pk_pt[q] < err_pos);
            if (fg_any(has)) {
                if (has) {
                    unsigned long long val = 0;
                    const uint32_t st = ltsv_parse_typed(p, pk_va[q], pk_va[q] + (int)(info & 0xFFFFu), q / 2 + 1, val);
                    if (st == FG_ST_OK) stage_val[info >> 16] = val;
                    else { status = st; err_pos = pk_pt[q]; }
                }
            }
        }
    }
    int err_rel = err_pos - ls;
    if (active_line && status == FG_ST_OK) {
        if (!have_ts) { status = FG_EL_MISSING_TS; err_rel = (le - ls) + 1; }             // :205
        else if (r.host_o < 0) { status = FG_EL_MISSING_HOST; err_rel = (le - ls) + 1; }  // :206
    }
    if (status == FG_ST_OK) {
        r.full_o = 0;  // full_msg = the whole line, untrimmed :219
        r.full_l = le - ls;
        r.n_entries = n;
    } else {
        r.full_o = err_rel;
    }
    r.flags = flags;
    r.status = status;
    fg_syncwarp();
}

}  // namespace fg
