// CPU check of haf_grasping_b200/csrc/decimal_round.cuh against glibc: for sampled floats / doubles,
// text4(f) must equal strtod(snprintf("%.4g", f)) and text6(v) must equal strtod(snprintf("%g", v)) bit for bit.
// usage: decimal_round_check <n_random> <seed>   -> prints "mismatches=<k> checked=<n> unsupported=<u>"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include "../haf_grasping_b200/csrc/decimal_round.cuh"

static uint64_t sm64(uint64_t& s) { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static long mism = 0, checked = 0, unsup = 0, unsup4 = 0;
static void check4(float f) {
    char buf[64]; snprintf(buf, sizeof buf, "%.4g", (double)f);
    double ref = strtod(buf, NULL); bool u = false; double got = hafdec::text4(f, &u);
    checked++; if (u) { unsup++; unsup4++; }
    if (memcmp(&ref, &got, 8) != 0 && !(ref != ref && got != got)) { if (mism < 20) printf("text4 MISMATCH f=%.9g (%a) ref=%.17g got=%.17g\n", f, f, ref, got); mism++; }
}
static void check6(double v, bool must_support) {
    char buf[64]; snprintf(buf, sizeof buf, "%g", v);
    double ref = strtod(buf, NULL); bool u = false; double got = hafdec::text6(v, &u);
    checked++; if (u) { unsup++; if (!must_support) return; }
    if (memcmp(&ref, &got, 8) != 0 && !(ref != ref && got != got)) { if (mism < 20) printf("text6 MISMATCH v=%.17g (%a) ref=%.17g got=%.17g\n", v, v, ref, got); mism++; }
}
int main(int argc, char** argv) {
    if (argc > 3 && !strcmp(argv[1], "exhaustive")) {  // every float bit pattern of slice part/nparts
        uint64_t part = strtoull(argv[2], 0, 10), nparts = strtoull(argv[3], 0, 10);
        for (uint64_t b = part; b < (1ull << 32); b += nparts) { uint32_t bb = (uint32_t)b; float f; memcpy(&f, &bb, 4); check4(f); }
        printf("mismatches=%ld checked=%ld unsupported=%ld unsupported_text4=%ld\n", mism, checked, unsup, unsup4);
        return mism != 0 || unsup4 != 0;
    }
    long n = argc > 1 ? atol(argv[1]) : 1000000; uint64_t seed = argc > 2 ? strtoull(argv[2], 0, 10) : 1;
    // 1. random bit patterns: every exponent incl. denormals, inf, nan
    for (long i = 0; i < n; i++) { uint32_t b = (uint32_t)sm64(seed); float f; memcpy(&f, &b, 4); check4(f); }
    // 2. values near decimal rounding ties at 4 digits: (k + 0.5) * 10^e  +- few ulps, e in the float range
    for (int e = -44; e <= 34; e++) for (int k = 1000; k < 10000; k += (n > 100000 ? 7 : 97)) {
        char buf[64]; snprintf(buf, sizeof buf, "%d.5e%d", k, e - 3); float f = strtof(buf, NULL);
        if (!(f > 0) || f > 3e38f) continue;
        uint32_t b; memcpy(&b, &f, 4);
        for (int d = -2; d <= 2; d++) { uint32_t bb = b + d; float g; memcpy(&g, &bb, 4); check4(g); check4(-g); }
    }
    // 3. realistic magnitudes (feature values): uniform * 10^[-8, 4]
    for (long i = 0; i < n; i++) { double u = (double)(sm64(seed) >> 11) / 9007199254740992.0; int e = (int)(sm64(seed) % 13) - 8; check4((float)((u * 2 - 1) * pow(10.0, e))); }
    // exact powers of ten and small integers
    for (int e = -45; e <= 38; e++) { char buf[32]; snprintf(buf, sizeof buf, "1e%d", e); check4(strtof(buf, NULL)); snprintf(buf, sizeof buf, "9.9995e%d", e); check4(strtof(buf, NULL)); snprintf(buf, sizeof buf, "9.9994e%d", e); check4(strtof(buf, NULL)); }
    for (int i = -20000; i <= 20000; i++) check4((float)i * 0.5f);
    // text6 on doubles: scaled values are mostly in [-1.5, 1.5]; also random magnitudes 1e-60..1e60 and ties
    for (long i = 0; i < n; i++) { double u = (double)(sm64(seed) >> 11) / 9007199254740992.0; check6(u * 3 - 1.5, true); }
    for (long i = 0; i < n; i++) { double u = (double)(sm64(seed) >> 11) / 9007199254740992.0; int e = (int)(sm64(seed) % 121) - 60; check6((u + 0.1) * pow(10.0, e) * ((i & 1) ? -1 : 1), true); }
    for (int e = -60; e <= 60; e++) for (int k = 100000; k < 1000000; k += (n > 100000 ? 613 : 9973)) {
        char buf[64]; snprintf(buf, sizeof buf, "%d.5e%d", k, e - 5); double v = strtod(buf, NULL);
        uint64_t b; memcpy(&b, &v, 8);
        for (int d = -2; d <= 2; d++) { uint64_t bb = b + d; double g; memcpy(&g, &bb, 8); check6(g, true); }
    }
    for (long i = 0; i < n / 4; i++) { uint64_t b = sm64(seed); double v; memcpy(&v, &b, 8); check6(v, false); }  // any double: exact or flagged
    // results of -1 + 2*(v-min)/(max-min) type arithmetic near zero
    for (long i = 0; i < n; i++) { double u = (double)(sm64(seed) >> 11) / 9007199254740992.0; double t = -1.0 + 2.0 * (0.5 + (u - 0.5) * 1e-12); check6(t, true); }
    printf("mismatches=%ld checked=%ld unsupported=%ld unsupported_text4=%ld\n", mism, checked, unsup, unsup4);
    return mism != 0 || unsup4 != 0;
}
