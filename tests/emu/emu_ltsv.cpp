// emu_ltsv.cpp — CPU emulation of the LTSV device logic (TEST INFRASTRUCTURE, see cuda_shim.h).
//
// Compiles the product's walker sources (fg_ltsvfast.cuh: stage-1 TAB / ':' bitmaps, stage-2 part walk; fg_ltsv.cuh: the
// value parsers and the direct-path scanner) with g++ and replays what parse_ltsv_kernel does with them — CTA rounds over
// a staging tile, slot reservation from the TAB counts, staged 8-byte rows, side-table placement — one lane at a time.
// The result has the layout of fg_batch_out (columnar rows + side table), so the CPU test-suite can push it through the
// product's materialiser and compare with the oracle without a GPU.
#define FG_HOST_EMU 1
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/flowgger_cuda.h"
#include "../../flowgger_b200/csrc/fg_ltsvfast.cuh"

namespace {
constexpr int kLines = fg::kLtsvLinesPerCta;
constexpr int kSlots = fg::kLtsvStageSlots;

struct Tables {
    std::vector<double> ts;
    std::vector<uint32_t> meta;
    std::vector<fg_span> host, msg, full, sd, none;
    std::vector<fg_span> ename;
    std::vector<uint64_t> eval;
    std::vector<uint8_t> emeta;
    int rounds = 0, direct = 0;
};

void put_row(Tables& t, int i, int o0, const fg::LineResult& res, uint32_t begin, uint32_t n) {
    const bool ok = res.status == FG_ST_OK;
    t.ts[i] = res.ts;
    t.meta[i] = res.status | (res.facility << 8) | (res.severity << 16) | (res.flags << 24);
    t.host[i] = fg_span{res.host_o >= 0 ? o0 + res.host_o : -1, res.host_l};
    t.msg[i] = fg_span{res.msg_o >= 0 ? o0 + res.msg_o : -1, res.msg_l};
    t.full[i] = ok ? fg_span{res.full_o >= 0 ? o0 + res.full_o : -1, res.full_l} : fg_span{o0 + std::max(res.full_o, 0), 0};
    t.sd[i] = fg_span{(int32_t)begin, (int32_t)n};
}
}  // namespace

extern "C" {

// the two 16-bit masks of stage 1 (TAB | ':' << 16), exposed so tests can pin the SWAR identity per byte
uint32_t emu_ltsv_classify16(const uint8_t* p) {
    uint32_t w[4], t, c;
    memcpy(w, p, 16);
    fg::lt_classify16(w[0], w[1], w[2], w[3], t, c);
    return t | (c << 16);
}

// schema: n_schema names (NUL-terminated) + types; suffix[5] indexed by fg_ltsv_type (nullptr = none)
int emu_ltsv_decode(const uint8_t* bytes, const int32_t* offsets, int32_t n, int32_t tile_bytes, int32_t strip_eol,
                    const uint8_t* invalid, int32_t has_schema, int32_t n_schema, const char* const* names, const int32_t* types,
                    const char* const* suffix, fg_batch_out* out, int32_t* info /* [2]: rounds, direct-path lines */) {
    Tables* T = new Tables();
    const size_t nn = (size_t)std::max(n, 1);
    T->ts.resize(nn); T->meta.resize(nn); T->host.resize(nn); T->msg.resize(nn); T->full.resize(nn); T->sd.resize(nn);
    T->none.assign(nn, fg_span{-1, 0});
    // LtsvDeviceConfig, as fg_create lays it out
    std::string name_blob, suffix_blob;
    std::vector<int32_t> name_off{0}, tys;
    for (int k = 0; k < n_schema; ++k) {
        name_blob += names[k];
        name_off.push_back((int32_t)name_blob.size());
        tys.push_back(types[k]);
    }
    fg::LtsvDeviceConfig cfg{};
    cfg.has_schema = (has_schema || n_schema > 0) ? 1 : 0;
    cfg.n_schema = n_schema;
    cfg.suffix_off[0] = 0;
    for (int t = 0; t < 5; ++t) {
        if (t > 0 && suffix && suffix[t]) {
            cfg.suffix_present |= 1u << t;
            suffix_blob += suffix[t];
        }
        cfg.suffix_off[t + 1] = (int32_t)suffix_blob.size();
    }
    cfg.names = (const uint8_t*)name_blob.data();
    cfg.name_off = name_off.data();
    cfg.types = tys.data();
    cfg.suffix = (const uint8_t*)suffix_blob.data();
    const bool typed = cfg.has_schema != 0;
    uint32_t first[8], len_mask;
    fg::lt_schema_filters(cfg.names, cfg.name_off, cfg.n_schema, len_mask, first);
    const fg::LtsvSchemaView S = {cfg.names, cfg.name_off, cfg.types, cfg.n_schema, len_mask, first};

    const int64_t total_bytes = n > 0 ? offsets[n] : 0;
    std::vector<uint8_t> tile((size_t)tile_bytes + 64);
    const int bm_words = tile_bytes / 32 + 4;
    std::vector<uint32_t> bmT((size_t)bm_words + 4), bmC((size_t)bm_words + 4);
    std::vector<unsigned long long> stage((size_t)kSlots), stage_val((size_t)kSlots);
    // scratch table of the direct path
    std::vector<fg_span> tname;
    std::vector<uint64_t> tval;
    std::vector<uint8_t> tmeta;

    auto direct_line = [&](int i) {
        ++T->direct;
        const int d0 = offsets[i];
        int len = offsets[i + 1] - d0;
        bool bad = false;
        if (strip_eol && len > 0) {
            if (strip_eol == 2) {
                if (bytes[d0 + len - 1] == 0) --len;
            } else if (bytes[d0 + len - 1] == '\n') {
                --len;
                if (len > 0 && bytes[d0 + len - 1] == '\r') --len;
            }
            if (invalid && invalid[i]) { bad = true; len = 0; }
        }
        tname.assign((size_t)len / 2 + 2, fg_span{0, 0});
        tval.assign(tname.size(), 0);
        tmeta.assign(tname.size(), 0);
        fg::EntrySink tmp = {(int2*)tname.data(), (unsigned long long*)tval.data(), tmeta.data()};
        fg::LineResult res;
        fg::ltsv_parse_line(bytes + d0, len, d0, 0u, true, cfg, res, tmp);
        if (bad) { res.status = FG_ES_INVALID_UTF8; res.n_entries = 0; res.full_o = 0; }
        const uint32_t my_n = res.status == FG_ST_OK ? res.n_entries : 0u;
        const uint32_t begin = my_n ? (uint32_t)T->ename.size() : 0u;
        for (uint32_t k = 0; k < my_n; ++k) {
            T->ename.push_back(tname[k]);
            T->eval.push_back(tval[k]);
            T->emeta.push_back(tmeta[k]);
        }
        put_row(*T, i, d0, res, begin, my_n);
    };

    for (int first = 0; first < n; first += kLines) {
        const int last = std::min(n, first + kLines);
        int cur = first;
        bool direct_next = false;
        while (cur < last) {
            const int base = offsets[cur] & ~15;
            int r = 0;
            if (!direct_next)
                while (cur + r < last && offsets[cur + r + 1] - base <= tile_bytes) ++r;
            if (r == 0) {
                direct_next = false;
                direct_line(cur);
                ++cur;
                continue;
            }
            ++T->rounds;
            const int oend = offsets[cur + r];
            const uint32_t nbytes = (uint32_t)((oend - base) + 15) & ~15u;
            for (uint32_t k = 0; k < nbytes; ++k) tile[k] = (int64_t)base + k < total_bytes ? bytes[base + k] : 0;  // the bulk copy
            const int nword = (int)((nbytes + 31u) >> 5);
            for (int g = 0; g < nword; ++g) {
                const uint32_t a = emu_ltsv_classify16(tile.data() + 32 * g), b = emu_ltsv_classify16(tile.data() + 32 * g + 16);
                bmT[g] = (a & 0xFFFFu) | (b << 16);
                bmC[g] = (a >> 16) | (b & 0xFFFF0000u);
            }
            for (int k = 0; k < 4; ++k) bmT[nword + k] = bmC[nword + k] = 0;
            // slot reservation
            std::vector<int> ls((size_t)r), le((size_t)r);
            std::vector<uint32_t> nb((size_t)r), slot0((size_t)r);
            std::vector<char> bad((size_t)r, 0);
            uint32_t run = 0;
            int r2 = 0;
            for (int tid = 0; tid < r; ++tid) {
                const int i = cur + tid;
                ls[tid] = offsets[i] - base;
                le[tid] = offsets[i + 1] - base;
                if (strip_eol && le[tid] > ls[tid]) {
                    if (strip_eol == 2) {
                        if (tile[le[tid] - 1] == 0) --le[tid];
                    } else if (tile[le[tid] - 1] == '\n') {
                        --le[tid];
                        if (le[tid] > ls[tid] && tile[le[tid] - 1] == '\r') --le[tid];
                    }
                    if (invalid && invalid[i]) bad[tid] = 1;
                }
                nb[tid] = bad[tid] ? 0u : (uint32_t)fg::lt_count_tabs(bmT.data(), ls[tid], le[tid]) + 1u;
                slot0[tid] = run;
                run += nb[tid];
                if (run <= (uint32_t)kSlots && r2 == tid) r2 = tid + 1;
            }
            if (r2 == 0) {
                direct_next = true;
                continue;
            }
            r = r2;
            for (int tid = 0; tid < r; ++tid) {
                const int i = cur + tid;
                fg::LineResult res;
                const bool walk = !bad[tid];
                if (typed) fg::ltsv_walk<true>(tile.data(), bmT.data(), bmC.data(), ls[tid], walk ? le[tid] : ls[tid], walk, cfg, S, res,
                                               stage.data() + slot0[tid], stage_val.data() + slot0[tid]);
                else fg::ltsv_walk<false>(tile.data(), bmT.data(), bmC.data(), ls[tid], walk ? le[tid] : ls[tid], walk, cfg, S, res,
                                          stage.data() + slot0[tid], stage_val.data() + slot0[tid]);
                if (bad[tid]) { res.status = FG_ES_INVALID_UTF8; res.n_entries = 0; res.full_o = 0; }
                const uint32_t my_n = (walk && res.status == FG_ST_OK) ? res.n_entries : 0u;
                const uint32_t begin = my_n ? (uint32_t)T->ename.size() : 0u;
                for (uint32_t k = 0; k < my_n; ++k) {  // the copy-out of parse_ltsv_kernel
                    const unsigned long long e = stage[slot0[tid] + k];
                    const int ka = base + (int)(e & 0xFFFFu), kn = (int)((e >> 16) & 0xFFFFu);
                    const uint32_t meta = (uint32_t)(e >> 56);
                    unsigned long long v = (unsigned long long)(uint32_t)(ka + kn + 1) | (((e >> 32) & 0xFFFFFFull) << 32);
                    if (typed && (meta & 0x07u) != 0u) v = stage_val[slot0[tid] + k];
                    T->ename.push_back(fg_span{ka, kn});
                    T->eval.push_back(v);
                    T->emeta.push_back((uint8_t)meta);
                }
                put_row(*T, i, offsets[i], res, begin, my_n);
            }
            cur += r;
        }
    }
    memset(out, 0, sizeof *out);
    out->n = n;
    out->n_entries = (int32_t)T->ename.size();
    out->ts = T->ts.data();
    out->meta = T->meta.data();
    out->hostname = T->host.data();
    out->appname = T->none.data();
    out->procid = T->none.data();
    out->msgid = T->none.data();
    out->msg = T->msg.data();
    out->full_msg = T->full.data();
    out->sd = T->sd.data();
    out->entry_name = T->ename.data();
    out->entry_val = T->eval.data();
    out->entry_meta = T->emeta.data();
    out->line_offsets = (const int32_t*)T;  // owner handle for emu_ltsv_free
    if (info) { info[0] = T->rounds; info[1] = T->direct; }
    return 0;
}

void emu_ltsv_free(fg_batch_out* out) {
    delete (Tables*)out->line_offsets;
    memset(out, 0, sizeof *out);
}

}  // extern "C"
