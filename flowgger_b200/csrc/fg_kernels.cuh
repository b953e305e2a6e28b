// fg_kernels.cuh — launch parameters shared by fg_kernels.cu and fg_abi.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fg {

// LTSV decoder configuration in device memory (ltsv_decoder.rs:19-22)
struct LtsvDeviceConfig {
    int32_t has_schema;
    int32_t n_schema;
    const uint8_t* names;       // concatenated schema key bytes
    const int32_t* name_off;    // [n_schema+1]
    const int32_t* types;       // [n_schema] fg_ltsv_type
    const uint8_t* suffix;      // concatenated suffix bytes
    int32_t suffix_off[6];      // per fg_ltsv_type: [t]..[t+1]; empty span + present bit
    uint32_t suffix_present;    // bit t set if a suffix is configured for type t
};

struct ParseParams {
    const uint8_t* bytes;     // device copy of the caller's byte buffer (base of all spans)
    const int32_t* offsets;   // [n+1] line offsets into bytes
    int32_t n;                // lines in this launch
    int32_t line0;            // batch index of the first line of this launch
    int32_t tile_bytes;       // dynamic shared memory staging tile, multiple of 16
    // row columns, element 0 = first line of this launch
    double* ts;
    uint32_t* meta;
    int2* host;
    int2* app;
    int2* proc;
    int2* msgid;
    int2* msg;
    int2* full;
    int2* sd;
    // structured-data side table
    int2* entry_name;
    unsigned long long* entry_val;
    uint8_t* entry_meta;
    // provisional rows of a line, indexed by its byte offset (see Format<>::scratch_index)
    int2* tmp_name;
    unsigned long long* tmp_val;
    uint8_t* tmp_meta;
    // split mode (fg_split_decode): every line still carries its "\n" / "\r\n" terminator, and lines flagged invalid
    // by the UTF-8 pass are not parsed
    const uint8_t* line_invalid;  // [n] or nullptr
    int32_t strip_eol;
    uint32_t* entry_counter;  // running total (atomic bump, one add per CTA round)
    uint32_t entry_cap;
    LtsvDeviceConfig ltsv;
};

constexpr int kLinesPerCta = 128;   // lines (= threads) per CTA (256 was measured slower: bigger barriers, same warps/SM)
constexpr int kMinCtasPerSm = 7;    // <= 72 registers/thread; shared memory (tile ~26 KB at 180 B/line) allows 7 CTAs
// GELF read straight from global memory is latency-bound (divergent tokenizer, ~7 active lanes): 16 CTAs/SM at 32
// registers beat 12 at 40 and 7 at 72 (10.5 / 10.9 / 12.3 ms per 3.5 M lines); LTSV is the opposite (5.8 -> 8.8 ms: L1 working set)
constexpr int kGelfUnstagedCtasPerSm = 16;
// RFC5424 (short lines, staged tile): 64-line CTAs at 14 per SM — same warps/SM, tile waits and barriers half as wide
// (2.36 -> 2.31 ms per 10 M lines; 13 KB tile + 1.8 KB static + 1 KB reserved per CTA = 225 KB of the SM's 227 KB)
constexpr int kRfc5424LinesPerCta = 64;
constexpr int kRfc5424CtasPerSm = 14;
constexpr int lines_per_cta(int fmt) { return fmt == 0 ? kRfc5424LinesPerCta : kLinesPerCta; }

cudaError_t launch_parse(int fmt, const ParseParams& p, cudaStream_t stream);
cudaError_t configure_kernels(int max_tile_bytes);
const char* kernel_build_info();

// device-side line framing + UTF-8 validation (fg_split.cu)
cudaError_t launch_split_chunk(const uint8_t* d_bytes, long long nbytes, long long c0, long long c1, int is_last, uint32_t* d_seg,
                               uint32_t* d_run, int32_t* d_cum_k, int32_t* d_offsets, int32_t* d_n_lines, int max_lines,
                               uint8_t* d_invalid, cudaStream_t stream);
int split_segments(long long nbytes);

}  // namespace fg
