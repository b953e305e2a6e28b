// fg_split.cu — device-side line framing + UTF-8 validation.
//
// Replaces the per-record work of LineSplitter::run (/root/reference/src/flowgger/splitter/line_splitter.rs:17-25):
// `BufRead::lines` (split at '\n', drop it and one preceding '\r', a last line without '\n' is still yielded) and the
// UTF-8 check of `String` (invalid line => "Invalid UTF-8 input", line skipped).  Input: a raw byte stream resident in
// HBM.  Output: int32 line-start offsets (line i = stream[offsets[i], offsets[i+1]) INCLUDING its terminator, which the
// parse kernels strip) and one invalid-UTF-8 flag per line.  These kernels are plain data-parallel byte scans (every
// byte read once per pass, 16-byte vector loads): two passes for the offsets (count, scan, fill) and one for UTF-8.
#include <cuda_runtime.h>
#include <stdint.h>

#include "fg_kernels.cuh"

namespace fg {

namespace {

constexpr int kSegBytes = 8192;   // one warp owns one segment
constexpr int kWarpsPerCta = 8;

__device__ __forceinline__ uint32_t nl_flags(uint32_t w) {  // 0x80 in every byte == '\n' (exact for all bytes)
    const uint32_t x = w ^ 0x0A0A0A0Au;
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
}

// bytes beyond nbytes are zero (the host clears a tail), so whole 16-byte loads are safe up to the padded end
__global__ void __launch_bounds__(kWarpsPerCta * 32) count_newlines_kernel(const uint8_t* __restrict__ bytes, long long nbytes,
                                                                            uint32_t* __restrict__ seg_counts, int nseg) {
    const int seg = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (seg >= nseg) return;
    const uint32_t lane = threadIdx.x & 31u;
    const long long base = (long long)seg * kSegBytes;
    uint32_t cnt = 0;
#pragma unroll 4
    for (int row = 0; row < kSegBytes / 512; ++row) {
        const long long pos = base + row * 512 + lane * 16;
        if (pos < nbytes) {
            const uint4 v = *reinterpret_cast<const uint4*>(bytes + pos);
            cnt += __popc(nl_flags(v.x)) + __popc(nl_flags(v.y)) + __popc(nl_flags(v.z)) + __popc(nl_flags(v.w));
        }
    }
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if (lane == 0) seg_counts[seg] = cnt;
}

// single CTA: exclusive scan of the segment counts; also the line count and the two end offsets
__global__ void __launch_bounds__(1024) scan_segments_kernel(uint32_t* __restrict__ seg_counts, int nseg, const uint8_t* __restrict__ bytes,
                                                             long long nbytes, int32_t* __restrict__ offsets, int32_t* __restrict__ n_lines,
                                                             int max_lines) {
    __shared__ uint32_t part[1024];
    const int t = threadIdx.x;
    const int per = (nseg + 1023) / 1024;
    const int lo = t * per, hi = min(nseg, lo + per);
    uint32_t s = 0;
    for (int k = lo; k < hi; ++k) s += seg_counts[k];
    part[t] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const uint32_t v = t >= d ? part[t - d] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t run = part[t] - s;  // exclusive prefix of this thread's range
    for (int k = lo; k < hi; ++k) {
        const uint32_t c = seg_counts[k];
        seg_counts[k] = run;
        run += c;
    }
    if (t == 1023) {
        const uint32_t newlines = part[1023];
        const bool tail = nbytes > 0 && bytes[nbytes - 1] != '\n';  // BufRead::lines yields an unterminated last line
        const long long n = (long long)newlines + (tail ? 1 : 0);
        *n_lines = n > max_lines ? -1 : (int32_t)n;
        offsets[0] = 0;
        if (n <= max_lines && tail) offsets[n] = (int32_t)nbytes;
    }
}

__global__ void __launch_bounds__(kWarpsPerCta * 32) fill_offsets_kernel(const uint8_t* __restrict__ bytes, long long nbytes,
                                                                          const uint32_t* __restrict__ seg_base, int nseg,
                                                                          int32_t* __restrict__ offsets, const int32_t* __restrict__ n_lines) {
    if (*n_lines < 0) return;  // more lines than the context can hold: the host reports FG_E_CAPACITY
    const int seg = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (seg >= nseg) return;
    const uint32_t lane = threadIdx.x & 31u;
    const long long base = (long long)seg * kSegBytes;
    uint32_t run = seg_base[seg];  // newlines before this segment
    for (int row = 0; row < kSegBytes / 512; ++row) {
        const long long pos = base + row * 512 + lane * 16;
        uint32_t z[4] = {0u, 0u, 0u, 0u};
        if (pos < nbytes) {
            const uint4 v = *reinterpret_cast<const uint4*>(bytes + pos);
            z[0] = nl_flags(v.x); z[1] = nl_flags(v.y); z[2] = nl_flags(v.z); z[3] = nl_flags(v.w);
        }
        const uint32_t c = __popc(z[0]) + __popc(z[1]) + __popc(z[2]) + __popc(z[3]);
        uint32_t inc = c;  // inclusive warp scan
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= (uint32_t)d) inc += y;
        }
        uint32_t k = run + inc - c;  // rank of this lane's first newline
#pragma unroll
        for (int wI = 0; wI < 4; ++wI) {
            uint32_t f = z[wI];
            while (f) {
                const int b = (__ffs((int)f) - 1) >> 3;
                offsets[k + 1] = (int32_t)(pos + wI * 4 + b + 1);  // the next line starts after this '\n'
                ++k;
                f &= f - 1;
            }
        }
        run += __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
}

// core::str::from_utf8 acceptance, data-parallel: every thread owns 16 bytes, validates the sequences that START in
// its chunk (reading up to 3 bytes beyond it) and checks that every continuation byte in its chunk is claimed by a
// lead byte at most 3 positions back.  Chunks of pure ASCII (the common case) cost four compares.
__device__ __forceinline__ int utf8_need(uint32_t c) {  // continuation bytes a lead byte announces; -1 = invalid lead
    if (c < 0x80u) return 0;
    if (c >= 0xC2u && c <= 0xDFu) return 1;
    if (c >= 0xE0u && c <= 0xEFu) return 2;
    if (c >= 0xF0u && c <= 0xF4u) return 3;
    return -1;  // 0x80..0xC1 (continuation or overlong lead), 0xF5..0xFF
}

__global__ void __launch_bounds__(256) validate_utf8_kernel(const uint8_t* __restrict__ bytes, long long nbytes,
                                                            const int32_t* __restrict__ offsets, const int32_t* __restrict__ n_lines,
                                                            uint8_t* __restrict__ invalid) {
    const int n = *n_lines;
    if (n <= 0) return;
    const long long pos = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (pos >= nbytes) return;
    const uint4 v = *reinterpret_cast<const uint4*>(bytes + pos);
    if (((v.x | v.y | v.z | v.w) & 0x80808080u) == 0u) return;  // ASCII only
    // an error flags the line that holds the offending byte (several lines may meet inside one chunk)
    auto flag = [&](long long bad) {
        int lo = 0, hi = n;  // offsets[lo] <= bad < offsets[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if ((long long)offsets[mid] <= bad) lo = mid;
            else hi = mid;
        }
        invalid[lo] = 1;
    };
    const long long end = pos + 16 < nbytes ? pos + 16 : nbytes;
    for (long long q = pos; q < end; ++q) {
        const uint32_t c = bytes[q];
        if (c < 0x80u) continue;
        if ((c & 0xC0u) == 0x80u) {
            // continuation byte: some lead within the 3 previous bytes must cover it
            bool claimed = false;
            for (int back = 1; back <= 3 && q - back >= 0; ++back) {
                const uint32_t l = bytes[q - back];
                if ((l & 0xC0u) == 0x80u) continue;  // another continuation byte: keep looking
                claimed = utf8_need(l) >= back;
                break;
            }
            if (!claimed) flag(q);
            continue;
        }
        const int need = utf8_need(c);
        if (need < 0) { flag(q); continue; }
        // continuation bytes (bytes past the end of the stream read as 0 and fail the check)
        uint32_t b1 = q + 1 < nbytes ? bytes[q + 1] : 0u;
        bool ok = (b1 & 0xC0u) == 0x80u;
        if (c == 0xE0u) ok = ok && b1 >= 0xA0u;             // overlong 3-byte
        else if (c == 0xEDu) ok = ok && b1 <= 0x9Fu;        // UTF-16 surrogates
        else if (c == 0xF0u) ok = ok && b1 >= 0x90u;        // overlong 4-byte
        else if (c == 0xF4u) ok = ok && b1 <= 0x8Fu;        // > U+10FFFF
        for (int k = 2; k <= need && ok; ++k) {
            const uint32_t bk = q + k < nbytes ? bytes[q + k] : 0u;
            ok = (bk & 0xC0u) == 0x80u;
        }
        if (!ok) flag(q);
    }
}

}  // namespace

cudaError_t launch_split(const uint8_t* d_bytes, long long nbytes, uint32_t* d_seg, int32_t* d_offsets, int32_t* d_n_lines,
                         int max_lines, uint8_t* d_invalid, cudaStream_t stream) {
    const int nseg = (int)((nbytes + kSegBytes - 1) / kSegBytes);
    const int grid = (nseg + kWarpsPerCta - 1) / kWarpsPerCta;
    if (nseg > 0) count_newlines_kernel<<<grid, kWarpsPerCta * 32, 0, stream>>>(d_bytes, nbytes, d_seg, nseg);
    scan_segments_kernel<<<1, 1024, 0, stream>>>(d_seg, nseg, d_bytes, nbytes, d_offsets, d_n_lines, max_lines);
    if (nseg > 0) {
        fill_offsets_kernel<<<grid, kWarpsPerCta * 32, 0, stream>>>(d_bytes, nbytes, d_seg, nseg, d_offsets, d_n_lines);
        const long long chunks = (nbytes + 15) / 16;
        validate_utf8_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, stream>>>(d_bytes, nbytes, d_offsets, d_n_lines, d_invalid);
    }
    return cudaGetLastError();
}

int split_segments(long long nbytes) { return (int)((nbytes + kSegBytes - 1) / kSegBytes); }

}  // namespace fg
