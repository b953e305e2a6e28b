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
// All 32 lanes of a warp must call this (idle lanes with active_line = false).  T = the tile, [ls, le) the line inside it
// (le - ls < 65536).  `stage` / `stage_val` = this line's reserved slots (>= #tabs + 1).  Result spans are relative to ls,
// exactly as ltsv_parse_line (fg_ltsv.cuh) reports them.
template <bool TYPED>
FG_DEV void ltsv_walk(const uint8_t* T, const uint32_t* bmT, const uint32_t* bmC, int ls, int le, bool active_line,
                      const LtsvDeviceConfig& cfg, LineResult& r, unsigned long long* stage, unsigned long long* stage_val) {
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
    // the `time` value and up to four typed values are parsed AFTER the part loop, in lock step (fg_ltsv.cuh explains why)
    int ts_a = -1, ts_b = -1, ts_part = 0;
    uint32_t np = 0;
    int t_va0 = 0, t_va1 = 0, t_va2 = 0, t_va3 = 0, t_pt0 = 0, t_pt1 = 0, t_pt2 = 0, t_pt3 = 0;
    uint32_t t_pk0 = 0, t_pk1 = 0, t_pk2 = 0, t_pk3 = 0;
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
                // the four reserved keys differ in (length, first byte): one cheap test rejects ordinary keys
                const uint32_t k0 = kn > 0 ? p[ka] : 0u;
                const bool maybe = (kn == 4 && (k0 == 't' || k0 == 'h')) || (kn == 7 && k0 == 'm') || (kn == 5 && k0 == 'l');
                if (maybe && key_is(p, ka, kn, "time", 4)) {  // :104-111
                    if (ts_a >= 0) {  // an earlier `time` is still pending: it is evaluated first (a failure returns there)
                        if (ltsv_parse_ts(p, ts_a, ts_b, r.ts)) have_ts = true;
                        else { status = FG_EL_TS; err_pos = ts_part; err_set = true; }
                    }
                    if (status == FG_ST_OK) {
                        ts_a = va;
                        ts_b = vb;
                        ts_part = part;
                        if (ts_b - ts_a >= 2 && p[ts_a] == '[' && p[ts_b - 1] == ']') { ++ts_a; --ts_b; }
                    } else {
                        ts_a = -1;
                    }
                } else if (maybe && key_is(p, ka, kn, "host", 4)) {
                    r.host_o = va - ls;
                    r.host_l = vb - va;
                } else if (maybe && key_is(p, ka, kn, "message", 7)) {
                    r.msg_o = va - ls;
                    r.msg_l = vb - va;
                } else if (maybe && key_is(p, ka, kn, "level", 5)) {  // :114-121
                    uint32_t sev;
                    if (!parse_u8(p, va, vb, sev)) status = FG_EL_SEV;
                    else if (sev > 7u) status = FG_EL_SEV_HIGH;
                    else r.severity = sev;
                } else {  // :122-199
                    uint32_t meta = 0;
                    bool deferred = false;
                    unsigned long long val = 0;
                    if (TYPED) {
                        const int type = cfg.has_schema ? ltsv_schema_type(p, ka, kn, cfg) : 0;
                        meta = (uint32_t)type;  // FG_TAG_* == fg_ltsv_type
                        if (type != 0 && ltsv_needs_suffix(p, ka, kn, type, cfg)) meta |= 0x20u;  // FG_EM_SUFFIX
                        if (type != 0) {
                            if (np < 4u && n < 256u) {
                                const uint32_t packed = (uint32_t)(vb - va) | ((uint32_t)type << 20) | (n << 24);
                                if (np == 0u) { t_va0 = va; t_pk0 = packed; t_pt0 = part; }
                                else if (np == 1u) { t_va1 = va; t_pk1 = packed; t_pt1 = part; }
                                else if (np == 2u) { t_va2 = va; t_pk2 = packed; t_pt2 = part; }
                                else { t_va3 = va; t_pk3 = packed; t_pt3 = part; }
                                ++np;
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
                if (!err_set) err_pos = part;  // the error belongs to the current part (unless an earlier `time` already failed)
                active = false;
            } else if (pe >= le) {
                active = false;
            } else {
                part = pe + 1;
            }
        }
    }
    // Deferred work, in lock step.  All of it is pure, so it is evaluated even when a later part already failed; the
    // reference returns at the FIRST failing part, i.e. the candidate error with the smallest part position wins.
    {
        const bool pend = active_line && ts_a >= 0 && (status == FG_ST_OK || ts_part < err_pos);
        if (fg_any(pend)) {
            if (pend) {
                double t;
                if (ltsv_parse_ts(p, ts_a, ts_b, t)) { r.ts = t; have_ts = true; }
                else { status = FG_EL_TS; err_pos = ts_part; }
            }
        }
    }
    if (TYPED) {
#pragma unroll
        for (uint32_t sl = 0; sl < 4u; ++sl) {
            const int va = sl == 0u ? t_va0 : (sl == 1u ? t_va1 : (sl == 2u ? t_va2 : t_va3));
            const uint32_t pk = sl == 0u ? t_pk0 : (sl == 1u ? t_pk1 : (sl == 2u ? t_pk2 : t_pk3));
            const int pt = sl == 0u ? t_pt0 : (sl == 1u ? t_pt1 : (sl == 2u ? t_pt2 : t_pt3));
            const bool has = active_line && sl < np && (status == FG_ST_OK || pt < err_pos);
            if (fg_any(has)) {
                if (has) {
                    unsigned long long val = 0;
                    const uint32_t st = ltsv_parse_typed(p, va, va + (int)(pk & 0xFFFFFu), (int)((pk >> 20) & 7u), val);
                    if (st == FG_ST_OK) stage_val[pk >> 24] = val;
                    else { status = st; err_pos = pt; }
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
