// fg_dtoa.cuh — f64 -> decimal text exactly as the reference's GELF encoder prints Record.ts and F64 values.
//
// The reference serialises through serde_json "~0.8" (Cargo.toml:51; gelf_encoder.rs:113 `serde_json::to_vec`), whose
// f64 path is `dtoa::write` (crate dtoa 0.2, un-vendored): Grisu2 (Loitsch 2010, in the formulation of RapidJSON's
// dtoa.h that the crate ports) followed by `prettify`: plain decimals for 10^-6 <= v < 10^21 (an integral value gets
// ".0"), exponent form `d.ddde±x` otherwise; NaN and the infinities are written as `null` by serde_json.  Grisu2 is NOT
// always the shortest representation, so it is restated step by step (DiyFp product rounded half up, the 87 cached
// powers of fg_dtoa_table.inc, digit generation with grisu_round) rather than replaced by a shortest-digits routine.
// Pinned by the reference's own encoder tests (gelf_encoder.rs:125,215: 1385053862.3072 and 123.456); everything else
// is "parity unpinned" (published algorithm of an un-vendored crate) and cross-checked against Python's repr in tests.
#pragma once
#include <stdint.h>

#include "fg_simt.cuh"

#ifdef FG_HOST_EMU
#define FG_DTOA_CONST const
#else
#define FG_DTOA_CONST __device__ const
#endif

namespace fg {

#include "fg_dtoa_table.inc"

struct DiyFp {
    uint64_t f;
    int e;
};

FG_DEV uint64_t dtoa_umulhi(uint64_t a, uint64_t b) { return __umul64hi(a, b); }

// product rounded half up to 64 bits (diyfp.rs Mul: tmp += 1 << 31 on the portable path, identical to h + (l >> 63))
FG_DEV DiyFp diy_mul(DiyFp x, DiyFp y) {
    const uint64_t h = dtoa_umulhi(x.f, y.f), l = x.f * y.f;
    DiyFp r;
    r.f = h + (l >> 63);
    r.e = x.e + y.e + 64;
    return r;
}
FG_DEV int dtoa_clz64(uint64_t v) { return __clzll((long long)v); }

FG_DEV uint32_t dtoa_pow10(int k) {
    uint32_t p = 1;
    for (int i = 0; i < k; ++i) p *= 10u;
    return p;
}

FG_DEV void grisu_round(uint8_t* buf, int len, uint64_t delta, uint64_t rest, uint64_t ten_kappa, uint64_t wp_w) {
    while (rest < wp_w && delta - rest >= ten_kappa && (rest + ten_kappa < wp_w || wp_w - rest > rest + ten_kappa - wp_w)) {
        buf[len - 1]--;
        rest += ten_kappa;
    }
}

// digits of v (> 0, finite) into buf (no NUL), returns their count; *K = decimal exponent: v = digits * 10^K
FG_DEV int grisu2(double value, uint8_t* buf, int* K) {
    uint64_t bits;
#ifdef FG_HOST_EMU
    memcpy(&bits, &value, 8);
#else
    bits = (uint64_t)__double_as_longlong(value);
#endif
    const int biased = (int)((bits >> 52) & 0x7FFu);
    const uint64_t frac = bits & 0x000FFFFFFFFFFFFFull;
    DiyFp v;
    if (biased) { v.f = frac + 0x0010000000000000ull; v.e = biased - 0x433; }
    else { v.f = frac; v.e = 1 - 0x433; }
    // normalized_boundaries
    DiyFp pl;
    pl.f = (v.f << 1) + 1;
    pl.e = v.e - 1;
    {
        const int s = dtoa_clz64(pl.f) - 10;  // until bit 53 (hidden << 1) is set
        pl.f <<= s;
        pl.e -= s;
        pl.f <<= 10;  // DIY_SIGNIFICAND_SIZE - SIGNIFICAND_SIZE - 2
        pl.e -= 10;
    }
    DiyFp mi;
    if (v.f == 0x0010000000000000ull) { mi.f = (v.f << 2) - 1; mi.e = v.e - 2; }
    else { mi.f = (v.f << 1) - 1; mi.e = v.e - 1; }
    mi.f <<= mi.e - pl.e;
    mi.e = pl.e;
    // get_cached_power(w_p.e)
    const double dk = (double)(-61 - pl.e) * 0.30102999566398114 + 347.0;
    int k = (int)dk;
    if (dk - (double)k > 0.0) ++k;
    const int index = (k >> 3) + 1;
    *K = -(-348 + (index << 3));
    DiyFp c;
    c.f = kDtoaPowF[index];
    c.e = kDtoaPowE[index];
    // w = normalize(v) * c_mk
    DiyFp vn = v;
    {
        const int s = dtoa_clz64(vn.f);
        vn.f <<= s;
        vn.e -= s;
    }
    const DiyFp W = diy_mul(vn, c);
    DiyFp Wp = diy_mul(pl, c);
    DiyFp Wm = diy_mul(mi, c);
    Wm.f += 1;
    Wp.f -= 1;
    // digit_gen(W, Wp, Wp.f - Wm.f)
    uint64_t delta = Wp.f - Wm.f;
    const int sh = -Wp.e;  // 32..60
    const uint64_t one_f = 1ull << sh;
    const uint64_t wp_w = Wp.f - W.f;
    uint32_t p1 = (uint32_t)(Wp.f >> sh);
    uint64_t p2 = Wp.f & (one_f - 1);
    int kappa = p1 < 10u ? 1 : p1 < 100u ? 2 : p1 < 1000u ? 3 : p1 < 10000u ? 4 : p1 < 100000u ? 5 : p1 < 1000000u ? 6
              : p1 < 10000000u ? 7 : p1 < 100000000u ? 8 : p1 < 1000000000u ? 9 : 10;
    int len = 0;
    while (kappa > 0) {
        const uint32_t div = dtoa_pow10(kappa - 1);
        const uint32_t d = p1 / div;
        p1 -= d * div;
        if (d || len) buf[len++] = (uint8_t)('0' + d);
        --kappa;
        const uint64_t tmp = ((uint64_t)p1 << sh) + p2;
        if (tmp <= delta) {
            *K += kappa;
            grisu_round(buf, len, delta, tmp, (uint64_t)dtoa_pow10(kappa) << sh, wp_w);
            return len;
        }
    }
    for (;;) {
        p2 *= 10u;
        delta *= 10u;
        const uint32_t d = (uint32_t)(p2 >> sh);
        if (d || len) buf[len++] = (uint8_t)('0' + d);
        p2 &= one_f - 1;
        --kappa;
        if (p2 < delta) {
            *K += kappa;
            const int idx = -kappa;
            grisu_round(buf, len, delta, p2, one_f, wp_w * (idx < 9 ? (uint64_t)dtoa_pow10(idx) : 0ull));
            return len;
        }
    }
}

FG_DEV int dtoa_write_exponent(int k, uint8_t* out) {
    int n = 0;
    if (k < 0) { out[n++] = '-'; k = -k; }
    if (k >= 100) { out[n++] = (uint8_t)('0' + k / 100); k %= 100; out[n++] = (uint8_t)('0' + k / 10); out[n++] = (uint8_t)('0' + k % 10); }
    else if (k >= 10) { out[n++] = (uint8_t)('0' + k / 10); out[n++] = (uint8_t)('0' + k % 10); }
    else out[n++] = (uint8_t)('0' + k);
    return n;
}

// serde_json 0.8 `Value::F64` / dtoa::write: returns the number of bytes written to out (at most 25)
FG_DEV int json_f64(double value, uint8_t* out) {
    uint64_t bits;
#ifdef FG_HOST_EMU
    memcpy(&bits, &value, 8);
#else
    bits = (uint64_t)__double_as_longlong(value);
#endif
    if (((bits >> 52) & 0x7FFu) == 0x7FFu) {  // NaN / infinite -> null
        out[0] = 'n'; out[1] = 'u'; out[2] = 'l'; out[3] = 'l';
        return 4;
    }
    int n = 0;
    if (bits >> 63) { out[n++] = '-'; value = -value; }
    if ((bits << 1) == 0) {
        out[n++] = '0'; out[n++] = '.'; out[n++] = '0';
        return n;
    }
    uint8_t dig[20];
    int K;
    const int length = grisu2(value, dig, &K);
    uint8_t* b = out + n;
    const int kk = length + K;  // 10^(kk-1) <= v < 10^kk
    if (0 <= K && kk <= 21) {  // 1234e7 -> 12340000000.0
        for (int i = 0; i < length; ++i) b[i] = dig[i];
        for (int i = length; i < kk; ++i) b[i] = '0';
        b[kk] = '.';
        b[kk + 1] = '0';
        return n + kk + 2;
    }
    if (0 < kk && kk <= 21) {  // 1234e-2 -> 12.34
        for (int i = 0; i < kk; ++i) b[i] = dig[i];
        b[kk] = '.';
        for (int i = kk; i < length; ++i) b[i + 1] = dig[i];
        return n + length + 1;
    }
    if (-6 < kk && kk <= 0) {  // 1234e-6 -> 0.001234
        const int off = 2 - kk;
        b[0] = '0';
        b[1] = '.';
        for (int i = 2; i < off; ++i) b[i] = '0';
        for (int i = 0; i < length; ++i) b[off + i] = dig[i];
        return n + length + off;
    }
    if (length == 1) {  // 1e30
        b[0] = dig[0];
        b[1] = 'e';
        return n + 2 + dtoa_write_exponent(kk - 1, b + 2);
    }
    b[0] = dig[0];  // 1234e30 -> 1.234e33
    b[1] = '.';
    for (int i = 1; i < length; ++i) b[i + 1] = dig[i];
    b[length + 1] = 'e';
    return n + length + 2 + dtoa_write_exponent(kk - 1, b + length + 2);
}

}  // namespace fg
