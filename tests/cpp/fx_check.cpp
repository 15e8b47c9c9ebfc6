// Host-side check of csrc/fx.h — the arithmetic behind the lean kernel's fixed-point double sums (FX).  The limb updates below mirror
// fx_add in csrc/agg_direct.cuh with plain adds in place of the shared-memory atomics (an atomic add returns the old value: same carry).
// Prints one line per case; exit code 0 iff every case holds.   g++ -O2 -std=c++17 -I csrc tests/cpp/fx_check.cpp -o fx_check
#include "fx.h"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include <algorithm>
#include <cmath>
using namespace bk;

struct Slot { uint32_t ext = 0, mid = 0, hi = 0; double exact = 0.0; long n_main = 0, n_fine = 0, n_exact = 0; };

static void add(Slot& s, double x, double scale, uint32_t fx_lo) {
    uint32_t lo; uint64_t up;
    const int kind = fx_split(x, scale, fx_lo, lo, up);
    if (kind == FX_MAIN) {
        const uint32_t old = s.mid; s.mid += lo;
        const uint32_t h = (uint32_t)up + ((uint32_t)(old + lo) < old ? 1u : 0u);
        if (h) s.hi += h;
        s.n_main++;
    } else if (kind == FX_FINE) {
        const uint32_t old = s.ext; s.ext += lo;
        const uint64_t t = up + ((uint32_t)(old + lo) < old ? 1u : 0u);
        const uint32_t m = (uint32_t)t;
        const uint32_t old2 = s.mid; s.mid += m;
        const uint32_t h = (uint32_t)(t >> 32) + ((uint32_t)(old2 + m) < old2 ? 1u : 0u);
        if (h) s.hi += h;
        s.n_fine++;
    } else if (kind == FX_EXACT) { s.exact += x; s.n_exact++; }
}
static double total(const Slot& s, int F) {
    const long long top = (long long)(((uint64_t)s.hi << 32) | s.mid);
    return fx_combine(top, s.ext, F) + s.exact;
}
static uint32_t sample_emax(const std::vector<double>& v, size_t nsamp) {
    uint32_t emax = 0;
    for (size_t i = 0; i < nsamp; i++) {
        const double x = v[i * v.size() / nsamp];
        uint64_t b; memcpy(&b, &x, 8);
        uint32_t e = (uint32_t)(b >> 52) & 0x7FFu;
        if (e == 0x7FFu) e = 0;
        emax = std::max(emax, e);
    }
    return emax;
}
static int fails = 0;
static void check(bool ok, const char* what) { if (!ok) { fails++; printf("FAIL %s\n", what); } }

// one column of values summed into `groups` slots; reference = long double sum (64-bit mantissa) per group
static void run_case(const char* name, std::vector<double> v, int groups, uint64_t rows_cta, unsigned seed) {
    std::mt19937_64 rng(seed);
    const int M = fx_magnitude_bits(rows_cta | FX_MIN_ROWS);
    const uint32_t emax = sample_emax(v, std::min<size_t>(640, v.size()));
    const int F = fx_scale_exp(M, emax);
    const double scale = fx_pow2(F);
    const uint32_t fx_lo = fx_floor_exp(M);
    std::vector<int> g(v.size());
    for (auto& x : g) x = (int)(rng() % (uint64_t)groups);
    std::vector<Slot> a((size_t)groups), b((size_t)groups);
    std::vector<long double> ref((size_t)groups, 0.0L), refabs((size_t)groups, 0.0L);
    for (size_t i = 0; i < v.size(); i++) { add(a[(size_t)g[i]], v[i], scale, fx_lo); ref[(size_t)g[i]] += (long double)v[i]; refabs[(size_t)g[i]] += fabsl((long double)v[i]); }
    std::vector<size_t> order(v.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::shuffle(order.begin(), order.end(), rng);
    for (size_t i : order) add(b[(size_t)g[i]], v[i], scale, fx_lo);
    double worst = 0.0; long nm = 0, nf = 0, ne = 0; bool same = true;
    const long double bound = ldexpl(1.0L, -(M - FX_MAIN_BINADES));   // relative to sum |x|: every value keeps >= M-15 significant bits
    for (int k = 0; k < groups; k++) {
        const double ta = total(a[(size_t)k], F), tb = total(b[(size_t)k], F);
        // the limbs never depend on the order of the rows; the exact (double) side path does, like any double sum
        if (a[(size_t)k].ext != b[(size_t)k].ext || a[(size_t)k].mid != b[(size_t)k].mid || a[(size_t)k].hi != b[(size_t)k].hi) same = false;
        (void)tb;
        nm += a[(size_t)k].n_main; nf += a[(size_t)k].n_fine; ne += a[(size_t)k].n_exact;
        if (refabs[(size_t)k] > 0) {
            const long double rel = fabsl((long double)ta - ref[(size_t)k]) / refabs[(size_t)k];
            if (std::isfinite((double)rel)) worst = std::max(worst, (double)rel);
            if (std::isfinite((double)ref[(size_t)k]) && !(rel <= bound + 1e-15L)) { fails++; printf("FAIL %s: group %d rel %.3Le > bound %.3Le\n", name, k, rel, bound); break; }
        }
    }
    check(same, "limbs independent of the order of the rows");
    printf("%-34s M=%2d F=%5d  main %9ld fine %8ld exact %7ld  worst |err|/sum|x| %.2e (bound %.2e)\n", name, M, F, nm, nf, ne, worst, (double)bound);
}

int main() {
    std::mt19937_64 rng(12345);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::uniform_real_distribution<double> ud(0.0, 1.0);
    const size_t N = 2000000;
    // classification of the special values
    {
        const int M = fx_magnitude_bits(700000 | FX_MIN_ROWS); const int F = fx_scale_exp(M, 1022); const double sc = fx_pow2(F); const uint32_t lo0 = fx_floor_exp(M);
        uint32_t lo; uint64_t up;
        check(M == 42, "700k rows per CTA leave 42 magnitude bits");
        check(fx_split(0.0, sc, lo0, lo, up) == FX_ZERO && fx_split(-0.0, sc, lo0, lo, up) == FX_ZERO, "zeros add nothing");
        check(fx_split(NAN, sc, lo0, lo, up) == FX_EXACT && fx_split(INFINITY, sc, lo0, lo, up) == FX_EXACT && fx_split(-INFINITY, sc, lo0, lo, up) == FX_EXACT, "NaN / Inf take the exact path");
        check(fx_split(0.75, sc, lo0, lo, up) == FX_MAIN, "a value near the sampled maximum is a main value");
        check(fx_split(3.99, sc, lo0, lo, up) == FX_MAIN && fx_split(4.0, sc, lo0, lo, up) == FX_EXACT, "2^FX_MARGIN times the sampled binade is the upper edge");
        check(fx_split(ldexp(0.75, -14), sc, lo0, lo, up) == FX_FINE && fx_split(ldexp(0.75, -44), sc, lo0, lo, up) == FX_FINE, "fine range");
        check(fx_split(ldexp(0.75, -45), sc, lo0, lo, up) == FX_EXACT && fx_split(5e-324, sc, lo0, lo, up) == FX_EXACT && fx_split(1e300, sc, lo0, lo, up) == FX_EXACT, "beyond the ranges: exact path");
        check(fx_split(-0.75, sc, lo0, lo, up) == FX_MAIN && (int64_t)up < 0, "negative values sign-extend");
        // round trip of single values: exact when the value has few bits
        for (double x : {0.5, -0.5, 0.75, 1.0 / 1024, -3.0 / 4096, 1.0 / (1 << 30), 1e-9, -1e-9, 2.5}) {
            Slot s; add(s, x, sc, lo0);
            const double back = total(s, F);
            check(fabs(back - x) <= ldexp(fabs(x), -(M - FX_MAIN_BINADES)), "single value round trip");
        }
        // exact cancellation and carries across all three limbs
        Slot s;
        for (int i = 0; i < 100000; i++) { add(s, 0.999999999, sc, lo0); add(s, 1e-8, sc, lo0); }
        for (int i = 0; i < 100000; i++) { add(s, -0.999999999, sc, lo0); add(s, -1e-8, sc, lo0); }
        check(s.ext == 0 && s.mid == 0 && s.hi == 0, "sum of x and -x is exactly zero (carries and borrows through every limb)");
        // the head room: rows_cta additions of the largest main value do not overflow
        Slot t; const double big = 3.999999;
        for (int i = 0; i < 700000; i++) add(t, big, sc, lo0);
        check(fabs(total(t, F) - 700000.0 * big) <= 700000.0 * big * 1e-9, "700k additions of the largest value stay in range");
        Slot u;
        for (int i = 0; i < 700000; i++) add(u, -big, sc, lo0);
        check(fabs(total(u, F) + 700000.0 * big) <= 700000.0 * big * 1e-9, "700k additions of the most negative value stay in range");
    }
    check(fx_magnitude_bits(1 | FX_MIN_ROWS) == FX_MAX_M && fx_magnitude_bits(2564 | FX_MIN_ROWS) == FX_MAX_M && fx_magnitude_bits(((uint64_t)1 << 40) | FX_MIN_ROWS) == 21, "magnitude bits: capped for small launches");
    {   // a small launch (few rows per CTA): the largest fine value times 2^32 must still round inside 63 bits
        const int M = fx_magnitude_bits(2564 | FX_MIN_ROWS); const int F = fx_scale_exp(M, 1022); const double sc = fx_pow2(F); const uint32_t lo0 = fx_floor_exp(M);
        uint32_t lo; uint64_t up;
        const double top_fine = ldexp(1.0 - ldexp(1.0, -53), M - FX_MAIN_BINADES - F);   // just below the main range's floor
        check(fx_split(top_fine, sc, lo0, lo, up) == FX_FINE && fx_split(-top_fine, sc, lo0, lo, up) == FX_FINE, "largest fine value");
        for (double x : {top_fine, -top_fine, top_fine / 3, -top_fine / 7}) {
            Slot s; add(s, x, sc, lo0);
            check(fabs(total(s, F) - x) <= ldexp(fabs(x), -40), "fine values of a small launch survive the second scaling");
        }
    }
    std::vector<double> v(N);
    for (auto& x : v) x = nd(rng) * 1e3;
    run_case("N(0, 1) * 1e3, small launch", v, 37, 2564, 20);
    for (auto& x : v) x = exp(nd(rng) * 6.0) * ((rng() & 1) ? 1.0 : -1.0);
    run_case("+-lognormal sigma 6, small launch", v, 37, 2564, 21);
    for (auto& x : v) x = ud(rng);
    run_case("uniform [0, 1)  (C2's 0_3)", v, 1000, 700000, 1);
    for (auto& x : v) x = nd(rng) * 1e3;
    run_case("N(0, 1) * 1e3   (C2's 0_4)", v, 1000, 700000, 2);
    run_case("N(0, 1) * 1e3, one group", v, 1, 2000000 + 4, 3);
    for (auto& x : v) x = exp(nd(rng) * 2.0);
    run_case("lognormal sigma 2 (heavy tail)", v, 100, 700000, 4);
    for (auto& x : v) x = exp(nd(rng) * 6.0);
    run_case("lognormal sigma 6 (52 binades)", v, 100, 700000, 5);
    for (size_t i = 0; i < N; i++) v[i] = (double)i * 1e-3;
    run_case("sorted ascending", v, 10, 700000, 6);
    for (size_t i = 0; i < N; i++) v[i] = (i % 1000 == 0) ? 1e12 : ud(rng) * 1e-6;
    run_case("outliers 1e18 x the bulk", v, 10, 700000, 7);
    for (auto& x : v) x = (double)((long long)(rng() % 2001) - 1000);
    run_case("integers -1000 .. 1000", v, 50, 700000, 8);
    for (auto& x : v) x = (rng() % 10 == 0) ? ud(rng) : 0.0;
    run_case("90 % zeros", v, 50, 700000, 9);
    for (auto& x : v) x = ud(rng) * 1e-300;
    run_case("tiny values (1e-300)", v, 50, 700000, 10);
    for (auto& x : v) x = ud(rng) * 1e300;
    run_case("huge values (1e300)", v, 50, 700000, 11);
    for (size_t i = 0; i < N; i++) v[i] = (i % 7 == 0) ? NAN : ud(rng);
    run_case("NaNs among the values", v, 50, 700000, 12);
    for (auto& x : v) x = 5e-324 * (double)(rng() % 1000);
    run_case("denormals", v, 50, 700000, 13);
    printf(fails ? "FAILED: %d\n" : "fx_check ok\n", fails);
    return fails ? 1 : 0;
}
