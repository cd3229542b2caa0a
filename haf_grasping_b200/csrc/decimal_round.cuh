// decimal_round.cuh -- exact emulation of the reference's two decimal text round trips.
//
// The reference passes every feature value through text twice before libsvm sees it:
//   1. write_featurevector prints the float with "%.4g"            (reference II2FV.cpp:133)
//      and svm-scale parses it back with sscanf("%lf")             (svm-scale.c:178, :270)
//   2. svm-scale prints the scaled double with "%g"                (svm-scale.c:350)
//      and svm-predict parses it back with strtod                  (svm-predict.c:108)
// text4(f) = strtod(sprintf("%.4g", (double)f)) and text6(v) = strtod(sprintf("%g", v)) are
// reproduced here bit-exactly without any string: round the binary value to N significant decimal
// digits (round-half-even on the EXACT value, as glibc printf does) giving R * 10^q, then return
// the double nearest to R * 10^q (what a correctly rounded strtod returns).
//
// Fast path (all values the pipeline produces in practice): the scaled value x*10^p is formed
// exactly as an unevaluated double-double (one FMA), rounded with an exact tie test, and the result
// R / 10^p (or R * 10^-p) is ONE correctly rounded IEEE operation on two exact operands (Clinger's
// fast path), valid for |p| <= 22.
// Slow path (|x| >= 1e4 for text4 / >= 1e6 for text6, or tiny magnitudes): fixed-width big-integer
// arithmetic, exact for every finite float (text4) and for doubles with 1e-60 <= |v| <= 1e60 (text6).
//
// Compiles for host (g++, used by the CPU unit test tests/test_decimal_round.py) and device (nvcc).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define HAFDEC_HD __host__ __device__ __forceinline__
#define HAFDEC_HD_NOINLINE __host__ __device__ __noinline__
#else
#define HAFDEC_HD inline
#define HAFDEC_HD_NOINLINE inline
#endif

namespace hafdec {

// 10^k for 0 <= k <= 22: all exactly representable in binary64.
#if defined(__CUDACC__)
__device__ __constant__ double kPow10Dev[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                                1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
#endif
HAFDEC_HD double pow10_exact(int k) {
#if defined(__CUDA_ARCH__)
    return kPow10Dev[k];
#else
    static const double t[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    return t[k];
#endif
}

HAFDEC_HD double fma_exact(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
HAFDEC_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
HAFDEC_HD double div_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
HAFDEC_HD uint64_t dbl_bits(double v) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(v);
#else
    uint64_t u;
    memcpy(&u, &v, 8);
    return u;
#endif
}
HAFDEC_HD double bits_dbl(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double v;
    memcpy(&v, &u, 8);
    return v;
#endif
}

// ------------------------------------------------------------------------------------------
// Fixed-width unsigned big integer (slow path only).  320 bits cover every case listed above.
// ------------------------------------------------------------------------------------------
struct Big {
    static const int NW = 10;
    uint32_t w[NW];
    bool overflow;
};
HAFDEC_HD void big_set(Big& b, uint64_t v) {
    for (int i = 0; i < Big::NW; i++) b.w[i] = 0;
    b.w[0] = (uint32_t)v;
    b.w[1] = (uint32_t)(v >> 32);
    b.overflow = false;
}
HAFDEC_HD void big_mul_small(Big& b, uint32_t m) {
    uint64_t carry = 0;
    for (int i = 0; i < Big::NW; i++) {
        uint64_t t = (uint64_t)b.w[i] * m + carry;
        b.w[i] = (uint32_t)t;
        carry = t >> 32;
    }
    if (carry) b.overflow = true;
}
HAFDEC_HD void big_mul_pow5(Big& b, int k) {
    while (k >= 13) { big_mul_small(b, 1220703125u); k -= 13; }  // 5^13
    uint32_t m = 1;
    for (int i = 0; i < k; i++) m *= 5u;
    if (m != 1) big_mul_small(b, m);
}
HAFDEC_HD int big_bitlen(const Big& a) {
    for (int i = Big::NW - 1; i >= 0; i--)
        if (a.w[i]) {
            int l = 0;
            uint32_t v = a.w[i];
            while (v) { l++; v >>= 1; }
            return i * 32 + l;
        }
    return 0;
}
HAFDEC_HD void big_shl(Big& b, int s) {
    if (s <= 0) return;
    if (big_bitlen(b) + s > Big::NW * 32) { b.overflow = true; return; }
    int ws = s >> 5, bs = s & 31;
    for (int i = Big::NW - 1; i >= 0; i--) {
        uint32_t v = 0;
        int src = i - ws;
        if (src >= 0) {
            v = b.w[src] << bs;
            if (bs && src - 1 >= 0) v |= b.w[src - 1] >> (32 - bs);
        }
        b.w[i] = v;
    }
}
HAFDEC_HD int big_cmp(const Big& a, const Big& b) {
    for (int i = Big::NW - 1; i >= 0; i--) {
        if (a.w[i] != b.w[i]) return a.w[i] > b.w[i] ? 1 : -1;
    }
    return 0;
}
HAFDEC_HD void big_sub(Big& a, const Big& b) {  // a -= b, requires a >= b
    uint64_t borrow = 0;
    for (int i = 0; i < Big::NW; i++) {
        uint64_t t = (uint64_t)a.w[i] - b.w[i] - borrow;
        a.w[i] = (uint32_t)t;
        borrow = (t >> 63) & 1;
    }
}
HAFDEC_HD bool big_is_zero(const Big& a) {
    for (int i = 0; i < Big::NW; i++)
        if (a.w[i]) return false;
    return true;
}
// q = floor(n / d) for a quotient known to be < 2^qbits; n becomes the remainder.
HAFDEC_HD uint64_t big_divrem_small_quot(Big& n, const Big& d, int qbits) {
    uint64_t q = 0;
    for (int bit = qbits - 1; bit >= 0; bit--) {
        Big t = d;
        big_shl(t, bit);
        if (!t.overflow && big_cmp(n, t) >= 0) {
            big_sub(n, t);
            q |= (1ull << bit);
        }
    }
    return q;
}

// value = M * 2^s (M > 0).  Rounds to `digits` significant decimal digits, half-even on the exact
// value: returns R (10^(digits-1) <= R <= 10^digits) and q with rounded value = R * 10^q.
// ok=false when the operands do not fit the fixed width (outside the documented range).
HAFDEC_HD_NOINLINE bool round_sig_big(uint64_t M, int s, int digits, uint64_t* R_out, int* q_out) {
    int bl = 0;
    {
        uint64_t v = M;
        while (v) { bl++; v >>= 1; }
    }
    // decimal exponent estimate of M*2^s from its bit length; fixed up by the loop below
    int e2 = bl - 1 + s;
    int E = (int)floor((double)e2 * 0.30102999566398120);
    uint64_t lo_lim = 1, hi_lim;
    for (int i = 1; i < digits; i++) lo_lim *= 10;
    hi_lim = lo_lim * 10;
    for (int iter = 0; iter < 6; iter++) {
        int q = E - (digits - 1);
        Big num, den;
        big_set(num, M);
        big_set(den, 1);
        if (q >= 0) big_mul_pow5(den, q); else big_mul_pow5(num, -q);
        int sh = s - q;
        if (sh >= 0) big_shl(num, sh); else big_shl(den, -sh);
        if (num.overflow || den.overflow) return false;
        // quotient < 2^24 whenever E is within +-1 of the truth (digits <= 6 -> R < 1e7 < 2^24)
        Big top = den;
        big_shl(top, 24);
        if (!top.overflow && big_cmp(num, top) >= 0) { E++; continue; }
        uint64_t R = big_divrem_small_quot(num, den, 24);
        Big twice = num;
        big_shl(twice, 1);
        int c = big_cmp(twice, den);
        if (c > 0 || (c == 0 && (R & 1))) R++;
        if (R > hi_lim) { E++; continue; }
        if (R < lo_lim) { E--; continue; }
        *R_out = R;
        *q_out = q;
        return true;
    }
    return false;
}

// nearest double to R * 10^q (R < 2^24), ties to even: what a correctly rounded strtod returns.
HAFDEC_HD_NOINLINE bool decimal_to_double_big(uint64_t R, int q, double* out) {
    if (R == 0) { *out = 0.0; return true; }
    Big num, den;
    big_set(num, R);
    big_set(den, 1);
    int e2 = 0;  // value = (num/den) * 2^e2
    if (q >= 0) { big_mul_pow5(num, q); e2 = q; }
    else { big_mul_pow5(den, -q); e2 = q; }
    if (num.overflow || den.overflow) return false;
    // scale so that the integer quotient has 56..57 bits
    int t = big_bitlen(den) - big_bitlen(num) + 57;
    if (t > 0) { big_shl(num, t); e2 -= t; }
    else if (t < 0) { big_shl(den, -t); e2 += -t; }
    if (num.overflow || den.overflow) return false;
    uint64_t Q = big_divrem_small_quot(num, den, 58);
    bool sticky = !big_is_zero(num);
    int ql = 0;
    {
        uint64_t v = Q;
        while (v) { ql++; v >>= 1; }
    }
    int drop = ql - 53;  // >= 3 by construction
    if (drop < 1) return false;
    uint64_t keep = Q >> drop;
    uint64_t rem = Q & ((1ull << drop) - 1);
    uint64_t half = 1ull << (drop - 1);
    if (rem > half || (rem == half && (sticky || (keep & 1)))) keep++;
    int ex = e2 + drop;  // value = keep * 2^ex, keep <= 2^53
    if (ex + 53 > 1023 || ex < -1021) return false;  // outside the documented range (never for |v| in 1e-60..1e60)
    // build the double exactly: keep has <= 54 bits; (double)keep exact when keep <= 2^53
    double d = (double)keep;
    uint64_t bits = dbl_bits(d);
    int64_t be = (int64_t)((bits >> 52) & 0x7FF) + ex;
    if (be <= 0 || be >= 2047) return false;
    bits = (bits & ~(0x7FFull << 52)) | ((uint64_t)be << 52);
    *out = bits_dbl(bits);
    return true;
}

// Round a (positive, finite) to `digits` significant digits and return the nearest double.
// a must be exactly representable as given (float promoted to double for text4; any double for text6).
// *inexact_range is set when the value lies outside the range the emulation covers exactly.
template <int DIGITS>
HAFDEC_HD double round_sig_to_double(double a, bool* unsupported) {
    const uint64_t bits = dbl_bits(a);
    const int be = (int)((bits >> 52) & 0x7FF);
    const int e2 = be - 1023;  // a in [2^e2, 2^(e2+1)) for normal doubles
    const double lo_lim = (DIGITS == 4) ? 1000.0 : 100000.0;
    const double hi_lim = lo_lim * 10.0;
    if (be != 0) {
        // E_est <= floor(log10 a) + (rarely) 1; the loop fixes it either way
        int E = (e2 * 1233) >> 12;
        for (int iter = 0; iter < 4; iter++) {
            int p = (DIGITS - 1) - E;  // scale by 10^p
            if (p < 0 || p > 22) break;
            double P = pow10_exact(p);
            double hi = mul_rn(a, P);
            double lo = fma_exact(a, P, -hi);  // a*P == hi + lo exactly
            double r = rint(hi);
            double d = hi - r;  // exact
            if (d == 0.5) { if (lo > 0.0) r += 1.0; }
            else if (d == -0.5) { if (lo < 0.0) r -= 1.0; }
            if (r > hi_lim) { E++; continue; }
            if (r < lo_lim) { E--; continue; }
            return div_rn(r, P);  // one correctly rounded operation on exact operands
        }
    }
    // slow path
    uint64_t M;
    int s;
    if (be == 0) { M = bits & 0xFFFFFFFFFFFFFull; s = -1074; }
    else { M = (bits & 0xFFFFFFFFFFFFFull) | (1ull << 52); s = be - 1075; }
    while ((M & 1) == 0) { M >>= 1; s++; }
    uint64_t R;
    int q;
    double out;
    if (round_sig_big(M, s, DIGITS, &R, &q) && decimal_to_double_big(R, q, &out)) return out;
    if (unsupported) *unsupported = true;
    return a;
}

// ------------------------------------------------------------------------------------------
// decimal -> binary32, correctly rounded (what strtof / `istream >> float` return): the ASCII PCD reader.
// value = M * 10^q (M != 0; sticky: non-zero digits were dropped behind M's 19).  Slow path: exact big-integer
// division, round half-even at 24 bits (fewer below 2^-126: subnormals), overflow -> +inf.
// ------------------------------------------------------------------------------------------
HAFDEC_HD_NOINLINE bool decimal_to_float_big(uint64_t M, int q, bool sticky_in, float* out) {
    if (q > 60) { *out = INFINITY; return true; }            // M >= 1: >= 1e61 > FLT_MAX
    if (q < -90) { *out = 0.0f; return true; }               // M < 2^64 ~ 1.8e19: < 1.8e-71, far below 2^-150
    Big num, den;
    big_set(num, M);
    big_set(den, 1);
    int e2 = q;  // value = (num / den) * 2^e2
    if (q >= 0) big_mul_pow5(num, q); else big_mul_pow5(den, -q);
    if (num.overflow || den.overflow) return false;
    int t = big_bitlen(den) - big_bitlen(num) + 57;   // integer quotient of 57..58 bits
    if (t > 0) { big_shl(num, t); e2 -= t; }
    else if (t < 0) { big_shl(den, -t); e2 += -t; }
    if (num.overflow || den.overflow) return false;
    const uint64_t Q = big_divrem_small_quot(num, den, 58);
    const bool sticky = sticky_in || !big_is_zero(num);
    int ql = 0;
    { uint64_t v = Q; while (v) { ql++; v >>= 1; } }
    const int E2 = e2 + ql - 1;                       // value in [2^E2, 2^(E2+1))
    if (E2 > 128) { *out = INFINITY; return true; }
    int drop = ql - 24;                               // normal: 24 significant bits
    if (E2 < -126) drop = -149 - e2;                  // subnormal: nothing below 2^-149
    if (drop >= 63) { *out = 0.0f; return true; }     // below 2^-150 by a wide margin
    if (drop < 1) return false;
    uint64_t keep = Q >> drop;
    const uint64_t rem = Q & ((1ull << drop) - 1), half = 1ull << (drop - 1);
    if (rem > half || (rem == half && (sticky || (keep & 1)))) keep++;
    *out = ldexpf((float)keep, e2 + drop);            // keep <= 2^24: exact; ldexpf: exact or +inf
    return true;
}
// Fast path: M < 2^53 and |q| <= 22 -> v = M * 10^q or M / 10^-q is ONE correctly rounded double operation on exact
// operands (Clinger); (float)v is then correct unless v sits exactly on the midpoint of two floats (the double rounding
// may have moved the value there) or in the subnormal / overflow range: those go to the exact path.
HAFDEC_HD float decimal_to_float(uint64_t M, int q, bool sticky, bool* unsupported) {
    if (M == 0) return 0.0f;
    if (!sticky && M < (1ull << 53) && q >= -22 && q <= 22) {
        const double d = (double)M;
        const double v = q >= 0 ? mul_rn(d, pow10_exact(q)) : div_rn(d, pow10_exact(-q));
        const uint64_t b = dbl_bits(v);
        const int be = (int)((b >> 52) & 0x7FF) - 1023;
        if (be >= -126 && be <= 126 && (b & 0x1FFFFFFFull) != 0x10000000ull) return (float)v;
    }
    float f;
    if (decimal_to_float_big(M, q, sticky, &f)) return f;
    if (unsupported) *unsupported = true;
    return (float)((double)M * pow(10.0, (double)q));
}
// One token of an ASCII PCD record as pcl::PCDReader reads a FLOAT32 field (PCL io/pcd_io.h copyStringValue<float>: "nan" ->
// quiet NaN, else `istringstream >> float`, else (float)atof): [+-] digits [. digits] [e|E [+-] digits] of the longest valid
// prefix, correctly rounded; "inf" / "infinity" / "nan" in any case; anything else 0.  p .. end delimit the token.
HAFDEC_HD float parse_float_token(const unsigned char* p, const unsigned char* end, bool* unsupported) {
    bool neg = false;
    if (p < end && (*p == '+' || *p == '-')) { neg = *p == '-'; p++; }
    if (end - p >= 3) {
        const unsigned char a = p[0] | 0x20, b = p[1] | 0x20, c = p[2] | 0x20;
        if (a == 'n' && b == 'a' && c == 'n') return neg ? -NAN : NAN;
        if (a == 'i' && b == 'n' && c == 'f') return neg ? -INFINITY : INFINITY;
    }
    uint64_t M = 0;
    int nd = 0, q = 0;          // significant digits taken into M (<= 19), decimal exponent of M's last digit
    bool any = false, sticky = false, seen_nonzero = false;
    for (; p < end && *p >= '0' && *p <= '9'; p++) {
        any = true;
        const int dg = *p - '0';
        if (dg) seen_nonzero = true;
        if (!seen_nonzero) continue;
        if (nd < 19) { M = M * 10 + (uint64_t)dg; nd++; }
        else { q++; if (dg) sticky = true; }
    }
    if (p < end && *p == '.') {
        p++;
        for (; p < end && *p >= '0' && *p <= '9'; p++) {
            any = true;
            const int dg = *p - '0';
            if (dg) seen_nonzero = true;
            if (!seen_nonzero) { q--; continue; }
            if (nd < 19) { M = M * 10 + (uint64_t)dg; nd++; q--; }
            else if (dg) sticky = true;
        }
    }
    if (!any) return 0.0f;   // no conversion: atof gives +0
    if (p < end && (*p == 'e' || *p == 'E')) {
        const unsigned char* s = p + 1;
        bool eneg = false;
        if (s < end && (*s == '+' || *s == '-')) { eneg = *s == '-'; s++; }
        if (s < end && *s >= '0' && *s <= '9') {
            int ex = 0;
            for (; s < end && *s >= '0' && *s <= '9'; s++) if (ex < 100000) ex = ex * 10 + (*s - '0');
            q += eneg ? -ex : ex;
        }
    }
    float f = decimal_to_float(M, q, sticky, unsupported);
    if (sticky) {
        // more than 19 significant digits: the value lies in (M, M + 1) * 10^q.  Rounding is monotonic, so when both ends give
        // the same float that float is the answer; otherwise a rounding boundary lies within 1e-19 (relative) of the value
        // and the dropped digits would decide: reported, not guessed (PCL writes <= 9 significant digits)
        const float f2 = decimal_to_float(M + 1, q, false, unsupported);
        if (!(f == f2) && unsupported) *unsupported = true;
    }
    return neg ? -f : f;
}

// strtod(sprintf("%.4g", (double)f))
HAFDEC_HD double text4(float f, bool* unsupported = nullptr) {
    double x = (double)f;
    if (!(fabs(x) < INFINITY) || x == 0.0) return x;  // "inf"/"nan"/"0"/"-0" round-trip to themselves
    double r = round_sig_to_double<4>(fabs(x), unsupported);
    return x < 0 ? -r : r;
}
// strtod(sprintf("%g", v))
HAFDEC_HD double text6(double v, bool* unsupported = nullptr) {
    if (!(fabs(v) < INFINITY) || v == 0.0) return v;
    double r = round_sig_to_double<6>(fabs(v), unsupported);
    return v < 0 ? -r : r;
}

}  // namespace hafdec
