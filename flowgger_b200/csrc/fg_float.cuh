// fg_float.cuh — decimal text -> f64 on device, bit-exact with the two algorithms the reference uses:
//
//  (1) Rust `f64::from_str` (core::num::dec2flt), used by the LTSV decoder for `time:<decimal>`
//      (ltsv_decoder.rs:256-261) and typed `f64` values (:157-160): correctly rounded (RNE),
//      grammar [+-](inf|infinity|nan|digits[.digits][(e|E)[+-]digits]).  Implemented as Clinger's exact
//      fast path (significand <= 2^53, |exp10| <= 22) plus an EXACT big-integer path for everything
//      else (up to 768 significant digits + sticky, like dec2flt's slow path), so the result is the
//      unique correctly rounded double for every input — no approximation, no CPU fallback.
//
//  (2) serde_json ~0.8 number assembly, used by the GELF decoder (gelf_decoder.rs:42): u64 significand,
//      then repeated * or / by a power of ten from a 309-entry table (NOT correctly rounded; two
//      roundings are possible) — `visit_f64_from_parts`.
#pragma once
#include "fg_common.cuh"

namespace fg {

// 10^0 .. 10^22 are exact doubles
static __device__ __constant__ double kPow10Exact[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                                  1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

constexpr int kBigLimbs = 128;  // 4096 bits: enough for 768 digits * 10^310 and for the scaled division case

struct BigNum {
    uint32_t v[kBigLimbs];
    int n;  // limbs in use
};

static __device__ __noinline__ void big_mul_small_add(BigNum& b, uint32_t m, uint32_t add) {
    uint64_t carry = add;
    for (int k = 0; k < b.n; ++k) {
        const uint64_t t = (uint64_t)b.v[k] * m + carry;
        b.v[k] = (uint32_t)t;
        carry = t >> 32;
    }
    if (carry && b.n < kBigLimbs) b.v[b.n++] = (uint32_t)carry;
}
// b /= d ; returns remainder
static __device__ __noinline__ uint32_t big_div_small(BigNum& b, uint32_t d) {
    uint64_t rem = 0;
    for (int k = b.n - 1; k >= 0; --k) {
        const uint64_t cur = (rem << 32) | b.v[k];
        b.v[k] = (uint32_t)(cur / d);
        rem = cur % d;
    }
    while (b.n > 0 && b.v[b.n - 1] == 0) --b.n;
    return (uint32_t)rem;
}
static __device__ __noinline__ void big_shl(BigNum& b, int s) {
    if (b.n == 0 || s == 0) return;
    const int ws = s >> 5, bs = s & 31;
    int nn = b.n + ws + 1;
    if (nn > kBigLimbs) nn = kBigLimbs;
    for (int k = nn - 1; k >= 0; --k) {
        const int src = k - ws;
        uint32_t lo = (src >= 0 && src < b.n) ? b.v[src] : 0u;
        uint32_t lo2 = (src - 1 >= 0 && src - 1 < b.n) ? b.v[src - 1] : 0u;
        b.v[k] = bs ? ((lo << bs) | (lo2 >> (32 - bs))) : lo;
    }
    b.n = nn;
    while (b.n > 0 && b.v[b.n - 1] == 0) --b.n;
}
FG_DEV int big_bitlen(const BigNum& b) {
    if (b.n == 0) return 0;
    return 32 * (b.n - 1) + (32 - __clz((int)b.v[b.n - 1]));
}
FG_DEV uint32_t big_bit(const BigNum& b, int i) {  // bit i (0 = LSB); 0 outside
    if (i < 0 || (i >> 5) >= b.n) return 0u;
    return (b.v[i >> 5] >> (i & 31)) & 1u;
}
// any bit set strictly below bit i ?
static __device__ __noinline__ bool big_any_below(const BigNum& b, int i) {
    if (i <= 0) return false;
    const int w = i >> 5, r = i & 31;
    for (int k = 0; k < w && k < b.n; ++k)
        if (b.v[k]) return true;
    if (r && w < b.n && (b.v[w] & ((1u << r) - 1u))) return true;
    return false;
}
// bits [i, i+64) as u64
FG_DEV uint64_t big_extract64(const BigNum& b, int i) {
    uint64_t out = 0;
    for (int k = 0; k < 64; ++k) out |= (uint64_t)big_bit(b, i + k) << k;
    return out;
}

// Exact: value = (digits of [a,b) skipping '.', as an integer) * 10^q10, sign applied by the caller.
// `p[a..b)` holds only digits and at most one '.'; leading zeros allowed.
static __device__ __noinline__ double big_decimal_to_f64(bytes_t p, int a, int b, int q10) {
    BigNum N;
    N.n = 0;
    int nd = 0;           // significant digits accumulated
    bool sticky = false;  // non-zero digits beyond the 768th
    uint32_t chunk = 0, chunk_mul = 1;
    for (int i = a; i < b; ++i) {
        const uint32_t c = p[i];
        if (c == '.') continue;
        const uint32_t dgt = c - '0';
        if (nd == 0 && dgt == 0) continue;  // leading zero: no effect on the integer
        if (nd < 768) {
            chunk = chunk * 10u + dgt;
            chunk_mul *= 10u;
            ++nd;
            if (chunk_mul == 1000000000u) {
                big_mul_small_add(N, chunk_mul, chunk);
                chunk = 0;
                chunk_mul = 1;
            }
        } else {
            sticky |= dgt != 0;
            ++q10;  // dropped digit: the kept integer is 10x smaller
        }
    }
    if (chunk_mul != 1) big_mul_small_add(N, chunk_mul, chunk);
    if (N.n == 0) return 0.0;
    // decimal magnitude: value in [10^(dec-1), 10^dec)
    const int dec = nd + q10;
    if (dec > 310) return __longlong_as_double(0x7FF0000000000000LL);
    if (dec < -326) return 0.0;
    int k = 0;  // value = N * 2^-k
    if (q10 >= 0) {
        int q = q10;
        while (q >= 9) { big_mul_small_add(N, 1000000000u, 0); q -= 9; }
        uint32_t m = 1;
        while (q-- > 0) m *= 10u;
        if (m > 1) big_mul_small_add(N, m, 0);
    } else {
        const int aq = -q10;
        // scale so that the quotient keeps >= 66 bits: 10^aq < 2^(3.3219281*aq + 1)
        const int need = 66 + (int)(((long long)aq * 217706LL) >> 16) + 2;
        const int bl0 = big_bitlen(N);
        k = need > bl0 ? need - bl0 : 0;
        big_shl(N, k);
        int q = aq;
        while (q >= 9) { sticky |= big_div_small(N, 1000000000u) != 0; q -= 9; }
        uint32_t m = 1;
        while (q-- > 0) m *= 10u;
        if (m > 1) sticky |= big_div_small(N, m) != 0;
        if (N.n == 0) return 0.0;
    }
    const int bl = big_bitlen(N);
    int E = bl - 1 - k;  // floor(log2(value))
    // s = number of low bits of N to drop so that the result is an integer count of ulps
    int s;
    if (E >= -1022) s = bl - 53;          // normal: keep 53 bits
    else s = k - 1074;                    // subnormal: units of 2^-1074
    uint64_t M;
    if (s <= 0) {
        M = big_extract64(N, 0) << (-s);  // exact (bl <= 53 here)
    } else {
        M = (s >= bl + 1) ? 0ull : big_extract64(N, s);
        const uint32_t rb = big_bit(N, s - 1);
        const bool rest = sticky || big_any_below(N, s - 1);
        if (rb && (rest || (M & 1ull))) ++M;  // round to nearest, ties to even
    }
    if (E >= -1022) {
        if (M == (1ull << 53)) { M >>= 1; ++E; }
        if (E > 1023) return __longlong_as_double(0x7FF0000000000000LL);
        const uint64_t bits = ((uint64_t)(E + 1023) << 52) | (M & ((1ull << 52) - 1ull));
        return __longlong_as_double((long long)bits);
    }
    // subnormal (or rounded up into the smallest normal: the encodings are contiguous)
    return __longlong_as_double((long long)M);
}


// ---- Eisel-Lemire (core::num::dec2flt::lemire::compute_float) ----------------------------------------------
// Decides almost every <= 19-digit input in ~60 instructions; returns false when it cannot (the caller then takes
// the exact big-integer path, like dec2flt's slow path).  Table: 128-bit truncated 5^q, q = -342 .. 308.
struct Pow5 {
    unsigned long long hi, lo;
};
static __device__ const Pow5 kPow5[651] = {
#include "fg_pow5_table.inc"
};

FG_DEV bool eisel_lemire(long long q, uint64_t w, uint64_t& bits) {
    if (w == 0 || q < -342) { bits = 0; return true; }
    if (q > 308) { bits = 0x7FF0000000000000ull; return true; }
    const int lz = __clzll((long long)w);
    w <<= lz;
    // compute_product_approx(q, w, 52 + 3)
    const Pow5 t = kPow5[q + 342];
    uint64_t first_lo = w * t.hi, first_hi = __umul64hi(w, t.hi);
    const uint64_t mask = 0xFFFFFFFFFFFFFFFFull >> 55;
    if ((first_hi & mask) == mask) {
        const uint64_t second_hi = __umul64hi(w, t.lo);
        first_lo += second_hi;
        if (second_hi > first_lo) ++first_hi;
    }
    const uint64_t lo = first_lo, hi = first_hi;
    if (lo == 0xFFFFFFFFFFFFFFFFull && !(q >= -27 && q <= 55)) return false;
    const int upperbit = (int)(hi >> 63);
    uint64_t mantissa = hi >> (upperbit + 64 - 52 - 3);
    int power2 = (int)((((int)q * (152170 + 65536)) >> 16) + 63) + upperbit - lz + 1023;
    if (power2 <= 0) {
        if (-power2 + 1 >= 64) { bits = 0; return true; }
        mantissa >>= -power2 + 1;
        mantissa += mantissa & 1ull;
        mantissa >>= 1;
        power2 = (mantissa >= (1ull << 52)) ? 1 : 0;
        bits = mantissa + ((uint64_t)power2 << 52) - ((power2 ? (1ull << 52) : 0ull));
        // BiasedFp{f: mantissa, e: power2}: when the subnormal rounds up to 2^52 the hidden bit IS the exponent field
        bits = power2 ? (1ull << 52) | (mantissa & ((1ull << 52) - 1ull)) : mantissa;
        return true;
    }
    if (lo <= 1 && q >= -4 && q <= 23 && (mantissa & 3ull) == 1ull && (mantissa << (upperbit + 64 - 52 - 3)) == hi) mantissa &= ~1ull;
    mantissa += mantissa & 1ull;
    mantissa >>= 1;
    if (mantissa >= (2ull << 52)) {
        mantissa = 1ull << 52;
        ++power2;
    }
    mantissa &= ~(1ull << 52);
    if (power2 >= 0x7FF) { bits = 0x7FF0000000000000ull; return true; }
    bits = mantissa | ((uint64_t)power2 << 52);
    return true;
}

FG_DEV bool ieq3(bytes_t p, int i, char a, char b, char c) {
    return (p[i] | 0x20u) == (uint32_t)a && (p[i + 1] | 0x20u) == (uint32_t)b && (p[i + 2] | 0x20u) == (uint32_t)c;
}

// Rust f64::from_str over [a,b).  Returns false on a grammar error.
static __device__ __noinline__ bool parse_f64_rust(bytes_t p, int a, int b, double& out) {
    if (a >= b) return false;
    int i = a;
    bool neg = false;
    if (p[i] == '+' || p[i] == '-') { neg = p[i] == '-'; ++i; }
    if (i >= b) return false;
    const int n = b - i;
    if (n == 3 && ieq3(p, i, 'i', 'n', 'f')) { out = neg ? -__longlong_as_double(0x7FF0000000000000LL) : __longlong_as_double(0x7FF0000000000000LL); return true; }
    if (n == 8 && ieq3(p, i, 'i', 'n', 'f') && ieq3(p, i + 3, 'i', 'n', 'i') && (p[i + 6] | 0x20u) == 't' && (p[i + 7] | 0x20u) == 'y') {
        out = neg ? -__longlong_as_double(0x7FF0000000000000LL) : __longlong_as_double(0x7FF0000000000000LL);
        return true;
    }
    if (n == 3 && ieq3(p, i, 'n', 'a', 'n')) {
        out = __longlong_as_double((long long)(0x7FF8000000000000ull | (neg ? 0x8000000000000000ull : 0ull)));
        return true;
    }
    const int mant_start = i;
    uint64_t w = 0;
    int nd = 0;         // significant digits seen (after leading zeros)
    int nfrac = 0;      // digits after the '.'
    int ndig = 0;       // all digits
    bool seen_dot = false;
    for (; i < b; ++i) {
        const uint32_t c = p[i];
        if (c == '.') {
            if (seen_dot) return false;
            seen_dot = true;
            continue;
        }
        const uint32_t dgt = c - '0';
        if (dgt > 9u) break;
        ++ndig;
        if (seen_dot) ++nfrac;
        if (nd > 0 || dgt != 0) {
            if (nd < 19) w = w * 10ull + dgt;
            ++nd;
        }
    }
    if (ndig == 0) return false;
    const int mant_end = i;
    long long e10 = 0;
    if (i < b) {
        if ((p[i] | 0x20u) != 'e') return false;
        ++i;
        bool eneg = false;
        if (i < b && (p[i] == '+' || p[i] == '-')) { eneg = p[i] == '-'; ++i; }
        if (i >= b) return false;
        for (; i < b; ++i) {
            const uint32_t dgt = (uint32_t)p[i] - '0';
            if (dgt > 9u) return false;
            if (e10 < 1000000) e10 = e10 * 10 + dgt;
        }
        if (eneg) e10 = -e10;
    }
    double v;
    if (nd == 0) {
        v = 0.0;
    } else {
        const long long q = e10 - nfrac;  // value = (all digits as integer) * 10^q
        if (nd <= 19 && w <= (1ull << 53) && q >= -22 && q <= 22) {
            // Clinger: both operands exact, a single IEEE rounding
            const double dw = (double)(long long)w;
            v = q < 0 ? __ddiv_rn(dw, kPow10Exact[-q]) : __dmul_rn(dw, kPow10Exact[q]);
        } else {
            // Eisel-Lemire on the (at most) 19 leading digits; with more digits the truncated and the bumped
            // significand must agree (dec2flt::dec2flt), otherwise — or when the approximation is undecided — exact path
            const long long q19 = q + (nd > 19 ? nd - 19 : 0);
            uint64_t b1 = 0, b2 = 0;
            bool ok = eisel_lemire(q19, w, b1);
            if (ok && nd > 19) ok = eisel_lemire(q19, w + 1, b2) && b1 == b2;
            if (ok) {
                v = __longlong_as_double((long long)b1);
            } else {
                int qq = q > 100000 ? 100000 : (q < -100000 ? -100000 : (int)q);
                v = big_decimal_to_f64(p, mant_start, mant_end, qq);
            }
        }
    }
    out = neg ? -v : v;
    return true;
}

// ---- serde_json ~0.8 `visit_f64_from_parts` -------------------------------------------------------
// pow10 table 1e0..1e308 lives in global memory (uploaded once by configure_kernels)
struct Pow10Table {
    double v[309];
};
static __device__ Pow10Table g_pow10;  // one copy per translation unit; only fg_parse_gelf.cu reads it and uploads its own (configure_parse_gelf)

// returns false for NumberOutOfRange
FG_DEV bool serde_f64_from_parts(bool pos, uint64_t significand, int exponent, double& out) {
    double f = __ull2double_rn(significand);
    for (;;) {
        const long long ae = exponent < 0 ? -(long long)exponent : (long long)exponent;
        if (ae <= 308) {
            const double pw = g_pow10.v[ae];
            if (exponent >= 0) {
                f = __dmul_rn(f, pw);
                if (isinf(f)) return false;
            } else {
                f = __ddiv_rn(f, pw);
            }
            break;
        }
        if (f == 0.0) break;
        if (exponent >= 0) return false;
        f = __ddiv_rn(f, 1e308);
        exponent += 308;
    }
    out = pos ? f : -f;
    return true;
}

}  // namespace fg
