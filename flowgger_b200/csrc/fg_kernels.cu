// fg_kernels.cu — the batched parse kernels (sm_100a).
//
// Execution model:
//   * one CTA = kLinesPerCta consecutive lines; the CTA's contiguous byte span
//     [offsets[first] & ~15, offsets[first+r]) is staged into shared memory with
//     ONE TMA bulk copy (cp.async.bulk, SASS UBLKCP) completing on an mbarrier,
//     so HBM is read once, fully coalesced, with no per-thread load instructions;
//   * one thread = one line: it walks its bytes in shared memory with the
//     format's sequential state machine and produces the Record fields;
//   * the variable-length structured-data rows are placed by a CTA-wide scan of
//     per-line counts plus ONE global atomic per CTA, then emitted by a second
//     walk over the (still resident) shared-memory bytes;
//   * the fixed-width row columns are written SoA, i.e. fully coalesced.
// Spans whose bytes do not fit the tile are handled in several rounds; a single
// line longer than the tile is parsed straight from global memory.
#include "fg_kernels.cuh"

#include "fg_common.cuh"
#include "fg_tma.cuh"
#include "fg_rfc5424.cuh"
#include "fg_gelf.cuh"
#include "fg_status.h"

#include <cstdio>
#include <cstdlib>

namespace fg {

// scratch table -> side table (loads first, then stores: independent L2 round trips in flight)
FG_DEV void copy_rows(uint32_t src, uint32_t dst, uint32_t n, const EntrySink& sink, const EntrySink& tmp) {
    uint32_t k = 0;
    for (; k + 4 <= n; k += 4) {
        int2 a[4];
        unsigned long long b[4];
        uint8_t c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a[u] = tmp.name[src + k + u]; b[u] = tmp.val[src + k + u]; c[u] = tmp.meta[src + k + u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) { sink.name[dst + k + u] = a[u]; sink.val[dst + k + u] = b[u]; sink.meta[dst + k + u] = c[u]; }
    }
    for (; k < n; ++k) {
        sink.name[dst + k] = tmp.name[src + k];
        sink.val[dst + k] = tmp.val[src + k];
        sink.meta[dst + k] = tmp.meta[src + k];
    }
}

template <int FMT>
struct Format;

struct NoShared {};

template <>
struct Format<2> {  // GELF
    typedef NoShared Shared;
    static FG_DEV void init_shared(Shared&) {}
    // a top-level member needs >= 5 input bytes (`"":0,`)
    static FG_DEV uint32_t scratch_index(int line_off, int) { return (uint32_t)line_off / 3u; }
    static FG_DEV void parse(bytes_t p, int len, int line_off, int line_idx, bool active, bool /*in_smem*/, Shared&, LineResult& r,
                             const EntrySink& tmp, const ParseParams&) {
        gelf_parse_line(p, len, line_off, scratch_index(line_off, line_idx), active, r, tmp);
    }
    static FG_DEV void expand(const LineResult&, int, uint32_t src, uint32_t dst, uint32_t n, const EntrySink& sink, const EntrySink& tmp) {
        copy_rows(src, dst, n, sink, tmp);
    }
};

// STAGE = true : the CTA's byte span is bulk-copied into shared memory first (short lines: RFC5424).
// STAGE = false: threads read their lines straight from global memory through L1 (long lines: at ~500 B/line the tile
//                would cap an SM at 12 resident warps; without it the register file is the limit).
template <int FMT, bool STAGE, int MINB = kMinCtasPerSm, int LINES = kLinesPerCta>
__global__ void __launch_bounds__(LINES, MINB) parse_kernel(const __grid_constant__ ParseParams P) {
    extern __shared__ __align__(128) uint8_t tile[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t scan_ws[33];
    __shared__ uint32_t s_ebase;
    __shared__ typename Format<FMT>::Shared fsh;

    const int tid = threadIdx.x;
    const int first = blockIdx.x * LINES;
    const int last = min(P.n, first + LINES);
    if (*P.bad_offsets) return;  // CTA-uniform
    if (tid == 0) mbar_init(&mbar, 1);
    Format<FMT>::init_shared(fsh);
    __syncthreads();

    const EntrySink sink = {P.entry_name, P.entry_val, P.entry_meta};
    const EntrySink tmp = {P.tmp_name, P.tmp_val, P.tmp_meta};
    uint32_t parity = 0;
    int cur = first;
    while (cur < last) {
        const int i = cur + tid;
        const int o0 = __ldg(P.offsets + min(i, last));
        const int o1 = __ldg(P.offsets + min(i + 1, last));
        const int ocur = __ldg(P.offsets + cur);
        const int base = ocur & ~15;
        const bool fits = (i < last) && (!STAGE || o1 - base <= P.tile_bytes);
        int r = __syncthreads_count(fits);  // offsets are monotone: `fits` is a prefix property
        const bool direct = !STAGE || (r == 0);  // r == 0: the first pending line alone exceeds the tile
        if (!STAGE) {
            // every pending line of the CTA is parsed from global memory in this single round
        } else if (direct) {
            r = 1;
        } else {
            if (tid == 0) {
                const int oend = __ldg(P.offsets + cur + r);
                const uint32_t nbytes = (uint32_t)((oend - base) + 15) & ~15u;
                fence_proxy_async();  // generic-proxy reads of the previous round happen-before this async write
                mbar_expect_tx(&mbar, nbytes);
                bulk_g2s(tile, P.bytes + base, nbytes, &mbar);
            }
            mbar_wait(&mbar, parity);
            parity ^= 1u;
        }
        const bool active = tid < r;
        int len = active ? o1 - o0 : 0;  // idle lanes run the lock-step phases on an empty line
        bool bad_utf8 = false;
        if (P.strip_eol && len > 0) {
            // BufRead::lines: drop the '\n' and one '\r' before it (line_splitter.rs:17); invalid UTF-8 lines are skipped (:22-25)
            const uint8_t* lp = direct ? P.bytes + o0 : tile + (o0 - base);
            if (P.strip_eol == 2) {  // BufRead::split(0): only the NUL terminator goes (nul_splitter.rs:18)
                if (lp[len - 1] == 0) --len;
            } else if (lp[len - 1] == '\n') {
                --len;
                if (len > 0 && lp[len - 1] == '\r') --len;
            }
            if (P.line_invalid != nullptr && P.line_invalid[i]) {
                bad_utf8 = true;
                len = 0;
            }
        }
        LineResult res;
        const int lidx = P.line0 + i;  // index of the line inside the batch (unique scratch slot)
        if (!direct) Format<FMT>::parse(tile + (o0 - base), len, o0, lidx, active, true, fsh, res, tmp, P);
        else Format<FMT>::parse(P.bytes + o0, len, o0, lidx, active, false, fsh, res, tmp, P);
        if (bad_utf8) {
            res.status = FG_ES_INVALID_UTF8;
            res.n_entries = 0;
            res.full_o = 0;
        }
        const uint32_t my_n = (active && res.status == FG_ST_OK) ? res.n_entries : 0u;
        uint32_t total;
        const uint32_t excl = block_exclusive_scan(my_n, scan_ws, total);
        uint32_t my_begin = 0;
        if (total) {  // CTA-uniform
            if (tid == 0) s_ebase = atomicAdd(P.entry_counter, total);
            __syncthreads();
            const uint32_t ebase = s_ebase;
            const bool ovf = (unsigned long long)ebase + total > (unsigned long long)P.entry_cap;
            if (my_n && !ovf) {
                // compact this line's staged rows from the scratch table into the side table
                my_begin = ebase + excl;
                Format<FMT>::expand(res, o0, Format<FMT>::scratch_index(o0, lidx), my_begin, my_n, sink, tmp);
            }
        }
        if (active) {
            const bool ok = res.status == FG_ST_OK;
            P.ts[i] = res.ts;
            P.meta[i] = res.status | (res.facility << 8) | (res.severity << 16) | (res.flags << 24);
            P.host[i] = make_int2(res.host_o >= 0 ? o0 + res.host_o : -1, res.host_l);
            if (false) {
                P.app[i] = make_int2(res.app_o >= 0 ? o0 + res.app_o : -1, res.app_l);
                P.proc[i] = make_int2(res.proc_o >= 0 ? o0 + res.proc_o : -1, res.proc_l);
                P.msgid[i] = make_int2(res.mid_o >= 0 ? o0 + res.mid_o : -1, res.mid_l);
            }
            P.msg[i] = make_int2(res.msg_o >= 0 ? o0 + res.msg_o : -1, res.msg_l);
            // error rows: full.off = absolute byte offset of the failing part (LTSV side effects), len 0
            P.full[i] = ok ? make_int2(res.full_o >= 0 ? o0 + res.full_o : -1, res.full_l)
                           : make_int2(o0 + max(res.full_o, 0), 0);
            P.sd[i] = make_int2((int)my_begin, (int)my_n);
        }
        __syncthreads();  // tile and scan scratch are reused by the next round
        cur += r;
    }
}

static int g_max_tile = 48 * 1024;

cudaError_t configure_kernels(int max_tile_bytes, int max_tile5424) {
    g_max_tile = max_tile_bytes;

    {
        static Pow10Table t;
        for (int k = 0; k <= 308; ++k) {
            char buf[16];
            snprintf(buf, sizeof buf, "1e%d", k);
            t.v[k] = strtod(buf, nullptr);  // correctly rounded decimal literals, like rustc's POW10 table
        }
        cudaError_t e1 = cudaMemcpyToSymbol(g_pow10, &t, sizeof t);
        if (e1 != cudaSuccess) return e1;
    }
    cudaError_t e = configure_parse5424(max_tile5424);
    if (e != cudaSuccess) return e;
    e = configure_gelf_encode(max_tile5424);
    if (e != cudaSuccess) return e;
    e = configure_parse_ltsv(kLtsvMaxTile);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(parse_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_tile_bytes);
    return e;
}

__global__ void __launch_bounds__(256) check_offsets_kernel(const int32_t* __restrict__ offsets, int n, long long max_bytes,
                                                           uint32_t* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const int a = offsets[i];
    bool bad = a < 0 || (long long)a > max_bytes;
    if (i < n) bad = bad || offsets[i + 1] < a;
    if (bad) atomicOr(flag, 1u);
}

cudaError_t launch_check_offsets(const int32_t* d_offsets, int n, long long max_bytes, uint32_t* d_flag, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    check_offsets_kernel<<<(n + 1 + 255) / 256, 256, 0, stream>>>(d_offsets, n, max_bytes, d_flag);
    return cudaGetLastError();
}

cudaError_t launch_parse(int fmt, const ParseParams& p, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    const int lines = lines_per_cta(fmt);
    const int grid = (p.n + lines - 1) / lines;
    switch (fmt) {
        case 1: return launch_parse_ltsv(p, stream);
        case 2:
            if (p.tile_bytes > 0) parse_kernel<2, true><<<grid, kLinesPerCta, p.tile_bytes, stream>>>(p);
            else parse_kernel<2, false, kGelfUnstagedCtasPerSm><<<grid, kLinesPerCta, 0, stream>>>(p);
            break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

const char* kernel_build_info() {
    return "flowgger_b200 parse kernels: sm_100a, RFC5424: structural bitmap + bit-walk over TMA-bulk-staged CTA tiles, "
           "kernels=[parse5424_kernel, post5424_kernel, gelf_size_kernel, gelf_write_kernel, parse_ltsv_kernel, parse_kernel<gelf>]";
}

}  // namespace fg
