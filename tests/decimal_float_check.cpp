// CPU check of hafdec::parse_float_token (haf_grasping_b200/csrc/decimal_round.cuh) against glibc strtof, which is what
// pcl::PCDReader's `istringstream >> float` ends in: for every generated decimal string the bit patterns must be equal.
// usage: decimal_float_check <n_random> <seed>   -> prints "mismatches=<k> checked=<n> unsupported=<u>"
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../haf_grasping_b200/csrc/decimal_round.cuh"

static uint64_t sm64(uint64_t& s) { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static long mism = 0, checked = 0, unsup = 0, flagged_long = 0;
static void check(const char* str) {
    const float ref = strtof(str, NULL);
    bool u = false;
    const float got = hafdec::parse_float_token((const unsigned char*)str, (const unsigned char*)str + strlen(str), &u);
    checked++;
    if (u) {   // allowed only for > 19 significant digits with a rounding boundary inside the last kept digit (see parse_float_token)
        int nd = 0; bool nz = false;
        for (const char* c = str; *c && *c != 'e' && *c != 'E'; c++) if (*c >= '0' && *c <= '9') { if (*c != '0') nz = true; if (nz) nd++; }
        if (nd > 19) { flagged_long++; return; }
        unsup++;
    }
    if (memcmp(&ref, &got, 4) != 0 && !(ref != ref && got != got)) { if (mism < 20) printf("MISMATCH '%s' ref=%.9g (%a) got=%.9g (%a)\n", str, ref, ref, got, got); mism++; }
}
int main(int argc, char** argv) {
    long n = argc > 1 ? atol(argv[1]) : 1000000;
    uint64_t seed = argc > 2 ? strtoull(argv[2], 0, 10) : 1;
    char buf[128];
    // 1. what PCL writes: %.{7,8,9,10}g of random floats of every binade (incl. subnormals), and of realistic coordinates
    for (long i = 0; i < n; i++) {
        uint32_t b = (uint32_t)sm64(seed); float f; memcpy(&f, &b, 4);
        if (f != f || f - f != 0) continue;
        snprintf(buf, sizeof buf, "%.*g", 6 + (int)(i % 5), (double)f); check(buf);
        const double u = (double)(sm64(seed) >> 11) / 9007199254740992.0;
        snprintf(buf, sizeof buf, "%.*g", 7 + (int)(i % 4), (u * 2 - 1) * 3.0); check(buf);
        snprintf(buf, sizeof buf, "%.*f", 3 + (int)(i % 9), (u * 2 - 1) * 10.0); check(buf);
        snprintf(buf, sizeof buf, "%.*e", 5 + (int)(i % 14), (u + 0.01) * 1e-17 * (double)(1 + i % 97)); check(buf);
    }
    // 2. exact midpoints between adjacent floats and their decimal neighbours (17+ digit strings): ties-to-even and the
    //    double-rounding trap of the fast path
    for (long i = 0; i < n / 4; i++) {
        uint32_t b = (uint32_t)sm64(seed) & 0x7F7FFFFFu; float f; memcpy(&f, &b, 4);
        uint32_t b2 = b + 1; float g; memcpy(&g, &b2, 4);
        if (g - g != 0) continue;
        const double mid = ((double)f + (double)g) * 0.5;     // exact in double
        snprintf(buf, sizeof buf, "%.60g", mid); check(buf);   // the exact decimal expansion of the midpoint (<= 60 digits shown)
        snprintf(buf, sizeof buf, "%.17g", mid); check(buf);
        snprintf(buf, sizeof buf, "%.17g", nextafter(mid, 1e300)); check(buf);
        snprintf(buf, sizeof buf, "%.17g", nextafter(mid, -1e300)); check(buf);
        snprintf(buf, sizeof buf, "%.19g", mid); check(buf);
        snprintf(buf, sizeof buf, "%.25g", mid); check(buf);
    }
    // 3. syntax: signs, leading zeros, no integer part, exponent forms, prefixes, specials, range ends
    const char* forms[] = {"0", "-0", "+0.0", "000123.4500", ".5", "-.5e1", "5.", "5.e-1", "1e5", "1E+5", "1e-5", "1e", "1e+", "1.5abc", "abc", "", "-", "+",
                           "nan", "NaN", "-nan", "inf", "-inf", "Infinity", "3.4028235e38", "3.4028236e38", "3.40282357e38", "3.5e38", "1e39", "-1e39", "1e60", "1e61", "1e400",
                           "1.17549435e-38", "1.17549428e-38", "1.4e-45", "0.7e-45", "0.700649232162408535e-45", "0.7006492321624086e-45", "1e-46", "1e-60", "1e-90", "1e-91", "1e-400",
                           "16777217", "16777219", "33554434", "33554438", "9007199254740993", "18446744073709551615", "18446744073709551616", "123456789012345678901234567890",
                           "0.000000000000000000000000000000000000000000001", "100000000000000000000000000000000000000", "-5.5302301e-18", "0.0007522106", "1.00000005960464477539062500000000000001"};
    for (size_t k = 0; k < sizeof forms / sizeof forms[0]; k++) check(forms[k]);
    for (int e = -50; e <= 40; e++) for (int m = 1; m < 1000; m += 7) { snprintf(buf, sizeof buf, "%d.%03de%d", m % 10, m, e); check(buf); }
    printf("mismatches=%ld checked=%ld unsupported=%ld flagged_long_tokens=%ld\n", mism, checked, unsup, flagged_long);
    return mism != 0 || unsup != 0;
}
