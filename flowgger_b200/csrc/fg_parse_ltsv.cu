// fg_parse_ltsv.cu — the LTSV decoder on sm_100a: bytes -> row columns + side table, on the bitmap pipeline.
//
//   parse_ltsv_kernel   one CTA = LINES consecutive lines.  (1) thread 0 issues ONE TMA bulk copy (cp.async.bulk, SASS UBLKCP)
//                       of the CTA's contiguous byte span into the shared-memory tile — HBM is read once, coalesced, with no
//                       per-thread global loads; (2) all threads sweep the tile 32 bytes per step and write the TAB and ':'
//                       bitmaps (fg_ltsvfast.cuh stage 1); (3) each line's staging slots are reserved from its TAB count by a
//                       CTA scan; (4) one thread per line walks its parts over the bitmaps (stage 2, lock step) and stages
//                       8-byte packed rows; (5) a second CTA scan + ONE global atomic place the rows, which all threads copy
//                       out slot by slot — consecutive threads write consecutive rows of the three side-table columns.
//   A line that does not fit the tile (or whose TAB count exceeds the staging area) is parsed by the round-1 scanner
//   (fg_ltsv.cuh) straight from global memory, through the scratch table.
#include "fg_kernels.cuh"

#include "fg_common.cuh"
#include "fg_ltsvfast.cuh"
#include "fg_status.h"
#include "fg_tma.cuh"

namespace fg {

namespace {

constexpr int kLines = kLtsvLinesPerCta;
constexpr int kStageSlots = kLtsvStageSlots;
constexpr int kSchemaKeys = 64, kSchemaBlob = 1024, kSuffixBlob = 64;  // larger schemas are read from global memory

// scratch table -> side table (the direct path only)
__device__ __forceinline__ void copy_rows_direct(uint32_t src, uint32_t dst, uint32_t n, const EntrySink& sink, const EntrySink& tmp) {
    for (uint32_t k = 0; k < n; ++k) {
        sink.name[dst + k] = tmp.name[src + k];
        sink.val[dst + k] = tmp.val[src + k];
        sink.meta[dst + k] = tmp.meta[src + k];
    }
}

__device__ __forceinline__ void write_row(const ParseParams& P, int i, int o0, const LineResult& res, uint32_t my_begin, uint32_t my_n) {
    const bool ok = res.status == FG_ST_OK;
    P.ts[i] = res.ts;
    P.meta[i] = res.status | (res.facility << 8) | (res.severity << 16) | (res.flags << 24);
    P.host[i] = make_int2(res.host_o >= 0 ? o0 + res.host_o : -1, res.host_l);
    P.msg[i] = make_int2(res.msg_o >= 0 ? o0 + res.msg_o : -1, res.msg_l);
    // error rows: full.off = absolute byte offset of the failing part (LTSV side effects), len 0
    P.full[i] = ok ? make_int2(res.full_o >= 0 ? o0 + res.full_o : -1, res.full_l) : make_int2(o0 + max(res.full_o, 0), 0);
    P.sd[i] = make_int2((int)my_begin, (int)my_n);
}

template <int LINES, bool TYPED>
__global__ void __launch_bounds__(LINES) parse_ltsv_kernel(const __grid_constant__ ParseParams P) {
    extern __shared__ __align__(128) uint8_t tile[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t scan_ws[33];
    __shared__ uint32_t s_ebase, s_slots;
    __shared__ uint32_t line_slot[LINES];   // first staging slot | rows << 16
    __shared__ uint32_t line_dense[LINES];  // exclusive sum of the rows of the lines before this one
    // TYPED: the CTA's copy of the schema and the suffixes (a few hundred bytes) — the walker compares keys against
    // shared memory instead of pulling the schema through L1 from every lane
    __shared__ uint8_t s_names[TYPED ? kSchemaBlob : 4];
    __shared__ int32_t s_name_off[TYPED ? kSchemaKeys + 1 : 1];
    __shared__ int32_t s_types[TYPED ? kSchemaKeys : 1];
    __shared__ uint8_t s_suffix[TYPED ? kSuffixBlob : 4];
    __shared__ uint32_t s_first[8];
    __shared__ uint32_t s_len_mask;

    const int tid = threadIdx.x;
    const int first = blockIdx.x * LINES;
    const int last = min(P.n, first + LINES);
    // behind the tile: the two bitmaps (tile_bytes / 32 + 4 words each), the staged rows, the slot -> line map
    const int bm_words = P.tile_bytes / 32 + 4;
    uint32_t* bmT = reinterpret_cast<uint32_t*>(tile + P.tile_bytes);
    uint32_t* bmC = bmT + bm_words;
    unsigned long long* stage = reinterpret_cast<unsigned long long*>(bmC + bm_words);
    unsigned long long* stage_val = stage + kStageSlots;  // TYPED only
    uint8_t* slot_line = reinterpret_cast<uint8_t*>(stage + (TYPED ? 2 : 1) * kStageSlots);
    if (*P.bad_offsets) return;  // CTA-uniform
    if (tid == 0) mbar_init(&mbar, 1);
    LtsvDeviceConfig cfg = P.ltsv;
    LtsvSchemaView S = {cfg.names, cfg.name_off, cfg.types, cfg.n_schema, 0xFFFFFFFFu, nullptr};
    if (TYPED) {
        const int nblob = cfg.n_schema > 0 ? cfg.name_off[cfg.n_schema] : 0, nsuf = cfg.suffix_off[5];
        if (cfg.n_schema <= kSchemaKeys && nblob <= kSchemaBlob && nsuf <= kSuffixBlob) {  // CTA-uniform
            for (int k = tid; k < nblob; k += LINES) s_names[k] = cfg.names[k];
            for (int k = tid; k <= cfg.n_schema; k += LINES) s_name_off[k] = cfg.name_off[k];
            for (int k = tid; k < cfg.n_schema; k += LINES) s_types[k] = cfg.types[k];
            for (int k = tid; k < nsuf; k += LINES) s_suffix[k] = cfg.suffix[k];
            __syncthreads();
            if (tid == 0) {
                uint32_t lm;
                lt_schema_filters(s_names, s_name_off, cfg.n_schema, lm, s_first);
                s_len_mask = lm;
            }
            __syncthreads();
            cfg.suffix = s_suffix;
            S = LtsvSchemaView{s_names, s_name_off, s_types, cfg.n_schema, s_len_mask, s_first};
        }
    }
    __syncthreads();

    const EntrySink sink = {P.entry_name, P.entry_val, P.entry_meta};
    const EntrySink tmp = {P.tmp_name, P.tmp_val, P.tmp_meta};
    uint32_t parity = 0;
    int cur = first;
    bool direct_next = false;  // CTA-uniform: line `cur` must take the direct path (its TAB count exceeds the staging area)
    while (cur < last) {
        const int i = cur + tid;
        const int o0 = __ldg(P.offsets + min(i, last));
        const int o1 = __ldg(P.offsets + min(i + 1, last));
        const int ocur = __ldg(P.offsets + cur);
        const int base = ocur & ~15;
        // tile_bytes <= 65024 (launch_parse_ltsv): a line inside the tile is shorter than 64 KiB, positions fit 16 bits
        const bool fits = !direct_next && (i < last) && (o1 - base <= P.tile_bytes);
        int r = __syncthreads_count(fits);  // offsets are monotone: `fits` is a prefix property
        if (r == 0) {
            // ---- direct path: line `cur` alone, read from global memory by the round-1 scanner ----------------------
            direct_next = false;
            const bool act = tid == 0;
            const int d0 = ocur;
            int len = act ? __ldg(P.offsets + cur + 1) - d0 : 0;
            bool bad_utf8 = false;
            if (P.strip_eol && len > 0) {
                const uint8_t* lp = P.bytes + d0;
                if (P.strip_eol == 2) {
                    if (lp[len - 1] == 0) --len;
                } else if (lp[len - 1] == '\n') {
                    --len;
                    if (len > 0 && lp[len - 1] == '\r') --len;
                }
                if (P.line_invalid != nullptr && P.line_invalid[cur]) {
                    bad_utf8 = true;
                    len = 0;
                }
            }
            LineResult res;
            const uint32_t sidx = (uint32_t)d0 / 2u + (uint32_t)(P.line0 + cur);  // Format<1>::scratch_index
            ltsv_parse_line(P.bytes + d0, len, d0, sidx, act, P.ltsv, res, tmp);
            if (bad_utf8) {
                res.status = FG_ES_INVALID_UTF8;
                res.n_entries = 0;
                res.full_o = 0;
            }
            if (act) {
                const uint32_t my_n = res.status == FG_ST_OK ? res.n_entries : 0u;
                uint32_t my_begin = 0;
                if (my_n) {
                    const uint32_t eb = atomicAdd(P.entry_counter, my_n);
                    if ((unsigned long long)eb + my_n <= (unsigned long long)P.entry_cap) {
                        my_begin = eb;
                        copy_rows_direct(sidx, eb, my_n, sink, tmp);
                    }
                }
                write_row(P, cur, d0, res, my_begin, my_n);
            }
            __syncthreads();
            cur += 1;
            continue;
        }
        const int oend = __ldg(P.offsets + cur + r);
        const uint32_t nbytes = (uint32_t)((oend - base) + 15) & ~15u;
        if (tid == 0) {
            fence_proxy_async();  // generic-proxy accesses of the previous round happen-before this async write
            mbar_expect_tx(&mbar, nbytes);
            bulk_g2s(tile, P.bytes + base, nbytes, &mbar);
        }
        mbar_wait(&mbar, parity);
        parity ^= 1u;

        // ---- stage 1: TAB and ':' bitmaps of the whole tile, 32 bytes (= one word of each) per thread per step -------
        const int nword = (int)((nbytes + 31u) >> 5);  // the tile allocation is a multiple of 512 bytes: reading the odd granule is safe
        for (int g = tid; g < nword; g += LINES) {
            const uint4 v0 = reinterpret_cast<const uint4*>(tile)[2 * g], v1 = reinterpret_cast<const uint4*>(tile)[2 * g + 1];
            uint32_t t0, c0, t1, c1;
            lt_classify16(v0.x, v0.y, v0.z, v0.w, t0, c0);
            lt_classify16(v1.x, v1.y, v1.z, v1.w, t1, c1);
            bmT[g] = t0 | (t1 << 16);
            bmC[g] = c0 | (c1 << 16);
        }
        if (tid < 4) {  // the walker's windows read up to two words past the last one
            bmT[nword + tid] = 0;
            bmC[nword + tid] = 0;
        }
        __syncthreads();

        // ---- the lines of this round; staging slots from the TAB counts -------------------------------------------
        bool active = tid < r;
        int ls = active ? o0 - base : 0;
        int le = active ? o1 - base : 0;
        bool bad_utf8 = false;
        if (P.strip_eol && le > ls) {
            // BufRead::lines: drop the '\n' and one '\r' before it (line_splitter.rs:17); invalid UTF-8 lines are skipped (:22-25)
            if (P.strip_eol == 2) {  // BufRead::split(0): only the NUL terminator goes (nul_splitter.rs:18)
                if (tile[le - 1] == 0) --le;
            } else if (tile[le - 1] == '\n') {
                --le;
                if (le > ls && tile[le - 1] == '\r') --le;
            }
            if (P.line_invalid != nullptr && P.line_invalid[i]) bad_utf8 = true;
        }
        const bool walk = active && !bad_utf8;
        const uint32_t nb = walk ? (uint32_t)lt_count_tabs(bmT, ls, le) + 1u : 0u;  // #parts >= #pairs
        uint32_t slots_total;
        const uint32_t slot0 = block_exclusive_scan(nb, scan_ws, slots_total);
        if (slots_total > (uint32_t)kStageSlots) {  // CTA-uniform, rare: keep the lines whose slots fit, redo the rest next round
            const int r2 = __syncthreads_count(active && slot0 + nb <= (uint32_t)kStageSlots);
            if (r2 == 0) {  // the first line alone has more parts than the staging area: direct path
                direct_next = true;
                __syncthreads();
                continue;
            }
            r = r2;
            active = tid < r;
        }
        const bool walk2 = walk && active;

        // ---- stage 2: one thread per line ------------------------------------------------------------------------
        LineResult res;
        ltsv_walk<TYPED>(tile, bmT, bmC, ls, walk2 ? le : ls, walk2, cfg, S, res, stage + slot0, stage_val + slot0);
        if (bad_utf8) {
            res.status = FG_ES_INVALID_UTF8;
            res.n_entries = 0;
            res.full_o = 0;
        }
        const uint32_t my_n = (walk2 && res.status == FG_ST_OK) ? res.n_entries : 0u;
        uint32_t total;
        const uint32_t excl = block_exclusive_scan(my_n, scan_ws, total);
        line_slot[tid] = slot0 | (my_n << 16);
        line_dense[tid] = excl;
        if (walk2)
            for (uint32_t k = 0; k < nb; ++k) slot_line[slot0 + k] = (uint8_t)tid;
        if (tid == r - 1) s_slots = slot0 + nb;  // slots in use this round
        if (tid == 0 && total) s_ebase = atomicAdd(P.entry_counter, total);
        __syncthreads();
        uint32_t my_begin = 0;
        if (total) {  // CTA-uniform
            const uint32_t ebase = s_ebase;
            const bool ovf = (unsigned long long)ebase + total > (unsigned long long)P.entry_cap;
            if (!ovf) {
                if (my_n) my_begin = ebase + excl;
                // every staged row -> the three side-table columns; slot s of line l is row (s - first slot of l)
                const uint32_t nslots = s_slots;
                for (uint32_t s = (uint32_t)tid; s < nslots; s += (uint32_t)LINES) {
                    const uint32_t l = slot_line[s];
                    const uint32_t ls_n = line_slot[l];
                    const uint32_t k = s - (ls_n & 0xFFFFu);
                    if (k < (ls_n >> 16)) {
                        const unsigned long long e = stage[s];
                        const uint32_t j = ebase + line_dense[l] + k;
                        const int ka = base + (int)(e & 0xFFFFu), kn = (int)((e >> 16) & 0xFFFFu);
                        const uint32_t meta = (uint32_t)(e >> 56);
                        sink.name[j] = make_int2(ka, kn);
                        unsigned long long v = (unsigned long long)(uint32_t)(ka + kn + 1) | (((e >> 32) & 0xFFFFFFull) << 32);
                        if (TYPED && (meta & 0x07u) != 0u) v = stage_val[s];
                        sink.val[j] = v;
                        sink.meta[j] = (uint8_t)meta;
                    }
                }
            }
        }
        if (active) write_row(P, i, o0, res, my_begin, my_n);
        __syncthreads();  // tile, bitmaps, staging and scan scratch are reused by the next round
        cur += r;
    }
}

}  // namespace

int parse_ltsv_smem_bytes(int tile_bytes, bool typed) {
    return tile_bytes + 2 * (tile_bytes / 32 + 4) * 4 + kStageSlots * 8 * (typed ? 2 : 1) + kStageSlots + 16;
}

cudaError_t configure_parse_ltsv(int max_tile_bytes) {
    cudaError_t e = cudaFuncSetAttribute(parse_ltsv_kernel<kLines, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         parse_ltsv_smem_bytes(max_tile_bytes, false));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(parse_ltsv_kernel<kLines, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                parse_ltsv_smem_bytes(max_tile_bytes, true));
}

cudaError_t launch_parse_ltsv(const ParseParams& p, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    if (p.tile_bytes <= 0 || p.tile_bytes > kLtsvMaxTile || (p.tile_bytes & 511)) return cudaErrorInvalidValue;
    const int grid = (p.n + kLines - 1) / kLines;
    const bool typed = p.ltsv.has_schema != 0;
    if (typed) parse_ltsv_kernel<kLines, true><<<grid, kLines, parse_ltsv_smem_bytes(p.tile_bytes, true), stream>>>(p);
    else parse_ltsv_kernel<kLines, false><<<grid, kLines, parse_ltsv_smem_bytes(p.tile_bytes, false), stream>>>(p);
    return cudaGetLastError();
}

}  // namespace fg
