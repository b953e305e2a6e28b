// emu_r3164.cpp — CPU emulation of the RFC3164 device logic (TEST INFRASTRUCTURE, see cuda_shim.h).
//
// Compiles the product's device source (fg_rfc3164.cuh: r3164_parse_line and the zone-table search) and its host-side
// zone-table builder (fg_tz.cu, which is plain C++) with g++ and replays what parse3164_kernel does with them: per CTA of
// 64 lines the contiguous byte span is copied into a 16-byte-aligned staging tile at the same in-granule offset the TMA copy
// gives it, every line that fits is parsed from the tile, the rest from the input buffer; re-joined messages go through the
// same bump allocator (with the same overflow / regrow / redo protocol the C ABI runs).  The kernel has no warp-level
// interplay — a line is parsed by one thread alone — so this replay covers the whole of its logic.
#define FG_HOST_EMU 1
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/flowgger_cuda.h"
#include "../../flowgger_b200/csrc/fg_rfc3164.cuh"
#include "../../flowgger_b200/csrc/fg_tz.h"

namespace {
constexpr int kLines = fg::kR3164LinesPerCta;

struct Tables {
    std::vector<double> ts;
    std::vector<uint32_t> meta;
    std::vector<fg_span> host, msg, full, sd;
    std::vector<uint8_t> arena;
    int from_tile = 0, from_global = 0, redo = 0;
};

fg::TzHostTable g_tz;
std::string g_tz_dir = "\x01";
}  // namespace

extern "C" {

// `year`: what now_utc().year() stands for.  tzdir: NULL = the system database.  arena_cap: initial capacity (small values
// exercise the regrow + redo path).  info: [3] lines parsed from the tile, from the input buffer, redo passes.
int emu_r3164_decode(const uint8_t* bytes, const int32_t* offsets, int32_t n, int32_t tile_bytes, int32_t strip_eol,
                     const uint8_t* invalid, int32_t year, const char* tzdir, uint32_t arena_cap, int32_t lockstep, fg_batch_out* out,
                     int32_t* info) {
    const std::string dir = tzdir ? tzdir : "";
    if (!g_tz.loaded || dir != g_tz_dir) {
        std::string err;
        if (!fg::tz_load_dir(tzdir, g_tz, err)) return -1;
        g_tz_dir = dir;
    }
    Tables* T = new Tables();
    const size_t nn = (size_t)std::max(n, 1);
    T->ts.resize(nn); T->meta.resize(nn); T->host.resize(nn); T->msg.resize(nn); T->full.resize(nn); T->sd.assign(nn, fg_span{0, 0});
    std::vector<uint8_t> tile((size_t)tile_bytes + 64);
    uint8_t* tile0 = tile.data() + ((16 - ((uintptr_t)tile.data() & 15)) & 15);  // 16-byte aligned like the shared-memory tile
    T->arena.resize(std::max<uint32_t>(arena_cap, 1));
    uint32_t counter = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        counter = 0;
        T->from_tile = T->from_global = 0;
        fg::R3164DeviceConfig cfg{};
        cfg.year = year;
        cfg.tz = g_tz.view();
        cfg.arena = T->arena.data();
        cfg.arena_cap = (uint32_t)T->arena.size();
        cfg.arena_counter = &counter;
        for (int first = 0; first < n; first += kLines) {
            const int last = std::min(n, first + kLines);
            const int base = offsets[first] & ~15;
            int r = 0;
            while (first + r < last && offsets[first + r + 1] - base <= tile_bytes) ++r;
            if (r > 0) {
                const uint32_t nbytes = (uint32_t)((offsets[first + r] - base) + 15) & ~15u;
                // the device buffer is padded; here only the bytes that exist are copied, the rest of the granule is zero
                const int64_t avail = (int64_t)offsets[n] - base;
                memset(tile0, 0, nbytes);
                memcpy(tile0, bytes + base, (size_t)std::min<int64_t>(nbytes, avail));
            }
            for (int i = first; i < last; ++i) {
                const int o0 = offsets[i], o1 = offsets[i + 1];
                const bool fits = i - first < r;
                const uint8_t* lp = fits ? tile0 + (o0 - base) : bytes + o0;
                int len = o1 - o0;
                bool bad = false;
                if (strip_eol && len > 0) {
                    if (strip_eol == 2) {
                        if (lp[len - 1] == 0) --len;
                    } else if (lp[len - 1] == '\n') {
                        --len;
                        if (len > 0 && lp[len - 1] == '\r') --len;
                    }
                    bad = invalid != nullptr && invalid[i] != 0;
                }
                fg::R3Out res;
                if (lockstep) {  // the walker the kernel is built with (FG_R3_LOCKSTEP = 1), a warp being this one lane
                    fg::r3164_parse_lockstep(lp, len, !bad, cfg, res);
                    if (bad) res.status = FG_ES_INVALID_UTF8;
                } else if (bad) {
                    res.status = FG_ES_INVALID_UTF8;
                    res.facility = res.severity = 0xFFu;
                    res.flags = 0u;
                    res.ts = 0.0;
                } else {
                    fg::r3164_parse_line(lp, len, cfg, res);
                }
                (fits ? T->from_tile : T->from_global)++;
                const bool ok = res.status == FG_ST_OK;
                T->ts[i] = ok ? res.ts : 0.0;
                T->meta[i] = res.status | (res.facility << 8) | (res.severity << 16) | (res.flags << 24);
                T->host[i] = ok ? fg_span{o0 + res.host_o, res.host_l} : fg_span{-1, 0};
                T->msg[i] = ok ? fg_span{(res.flags & fg::kR3FlagMsgArena) ? res.msg_o : o0 + res.msg_o, res.msg_l} : fg_span{-1, 0};
                T->full[i] = ok ? fg_span{o0, res.full_l} : fg_span{-1, 0};
            }
        }
        if (counter <= T->arena.size()) break;
        T->arena.assign((size_t)counter + counter / 8 + 1024, 0);  // regrow_tables, then the batch is redone
        ++T->redo;
    }
    memset(out, 0, sizeof *out);
    out->n = n;
    out->ts = T->ts.data();
    out->meta = T->meta.data();
    out->hostname = T->host.data();
    out->msg = T->msg.data();
    out->full_msg = T->full.data();
    out->sd = T->sd.data();
    out->arena = T->arena.data();
    out->arena_bytes = (int64_t)counter;
    out->line_offsets = (const int32_t*)T;  // owner handle for emu_r3164_free
    if (info) { info[0] = T->from_tile; info[1] = T->from_global; info[2] = T->redo; }
    return 0;
}

void emu_r3164_free(fg_batch_out* out) {
    delete (Tables*)out->line_offsets;
    memset(out, 0, sizeof *out);
}

// the packed zone table as the kernel searches it: 1 + offset stored, or 0 when `name` is no identifier
int emu_tz_lookup(const char* tzdir, const char* name, int64_t local, int32_t* offset) {
    const std::string dir = tzdir ? tzdir : "";
    if (!g_tz.loaded || dir != g_tz_dir) {
        std::string err;
        if (!fg::tz_load_dir(tzdir, g_tz, err)) return -1;
        g_tz_dir = dir;
    }
    const fg::TzDeviceTable V = g_tz.view();
    const int z = fg::tz_find(V, (const uint8_t*)name, 0, (int)strlen(name));
    if (z < 0) return 0;
    *offset = fg::tz_offset_local(V, z, (long long)local);
    return 1;
}
int emu_tz_count(const char* tzdir) {
    int32_t o;
    if (emu_tz_lookup(tzdir, "UTC", 0, &o) < 0) return -1;
    return g_tz.n_names();
}

}  // extern "C"
