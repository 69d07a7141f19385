// Test program: lbm::vtk_put (what lbm_output_save writes) against a real `ostream << value << ' '`
// (what the reference's outputSave writes, ldc.cu:603-607) on N pseudo-random bit patterns of float
// and double plus the special values, values next to rounding ties and next to powers of ten.  Prints the
// number of mismatches; exit code 0 iff none.
//   fmt_check N             N rounds against the ostream (slow: the ostream is the slow part)
//   fmt_check --exhaustive  lbm::vtk_write against std::to_chars on all 2^32 float patterns and 2^32 doubles
//                           (threads; minutes), std::to_chars being pinned to the ostream by the first mode
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <cmath>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "vtk_format.h"

template <typename V>
static long check(V v, long &shown) {
    std::string a;
    lbm::vtk_put(a, v);
    std::ostringstream os;
    os << v << ' ';
    if (a == os.str()) return 0;
    if (shown++ < 10) std::printf("mismatch: to_chars '%s' ostream '%s'\n", a.c_str(), os.str().c_str());
    return 1;
}

template <typename V>
static long check_fast(V v) {
    char a[lbm::VTK_MAX_CHARS], b[lbm::VTK_MAX_CHARS];
    const size_t na = (size_t)(lbm::vtk_write(a, v) - a), nb = (size_t)(lbm::vtk_write_slow(b, v) - b);
    if (na == nb && !std::memcmp(a, b, na)) return 0;
    a[na] = b[nb] = 0;
    std::printf("mismatch: fast '%s' to_chars '%s'\n", a, b);
    return 1;
}

static int exhaustive() {
    const unsigned nthr = std::max(1u, std::thread::hardware_concurrency());
    std::vector<long> bad(nthr, 0);
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthr; t++)
        pool.emplace_back([t, nthr, &bad]() {
            long b = 0;
            for (uint64_t u = t; u < (1ull << 32); u += nthr) {
                const uint32_t w = (uint32_t)u;
                float f;
                std::memcpy(&f, &w, 4);
                if (f != f) continue;
                b += check_fast(f);
                // a double per pattern: the float's bits spread over the double's exponent and mantissa
                uint64_t s = u * 0x9E3779B97F4A7C15ull;
                s ^= s >> 29, s *= 0xBF58476D1CE4E5B9ull, s ^= s >> 32;
                double d;
                std::memcpy(&d, &s, 8);
                if (d == d) b += check_fast(d);
                if (b > 20) break;
            }
            bad[t] = b;
        });
    for (auto &th : pool) th.join();
    long total = 0;
    for (long b : bad) total += b;
    std::printf("exhaustive: all float patterns and 2^32 doubles, %ld mismatches\n", total);
    return total ? 1 : 0;
}

int main(int argc, char **argv) {
    if (argc > 1 && !std::strcmp(argv[1], "--exhaustive")) return exhaustive();
    const long n = argc > 1 ? std::atol(argv[1]) : 3000000;
    long bad = 0, shown = 0;
    uint64_t s = 0x243F6A8885A308D3ull;
    auto next = [&s]() {
        s ^= s << 13, s ^= s >> 7, s ^= s << 17;
        return s;
    };
    for (long i = 0; i < n; i++) {
        uint32_t u = (uint32_t)next();
        float f;
        std::memcpy(&f, &u, 4);
        if (f != f) continue;  // NaN payloads print alike but are not produced by the solver
        bad += check(f, shown);
        // values of the magnitude the solver prints: velocities * C_U, densities * C_rho
        float g = (float)((double)(int32_t)(uint32_t)next() / 2147483648.0) * (i % 3 == 0 ? 1060.0f : (i % 3 == 1 ? 2.4705f : 1e-4f));
        bad += check(g, shown);
        uint64_t w = next();
        double d;
        std::memcpy(&d, &w, 8);
        if (d == d) bad += check(d, shown);
        bad += check((double)g * 3.3333, shown);
    }
    const float sp[] = {0.0f, -0.0f, 1.0f, 0.1f, 1e-5f, 9.999995e-5f, 999999.5f, 1e6f, 123456.7f, 1e-38f, 1e-45f,
                        std::numeric_limits<float>::max(), std::numeric_limits<float>::infinity(),
                        -std::numeric_limits<float>::infinity(), 0.0001f, 0.00001f, 100000.0f, 1000000.0f, 0.5f};
    for (float v : sp) bad += check(v, shown), bad += check((double)v, shown);
    // rounding ties of the six-digit form and their neighbours, powers of ten and their neighbours
    for (int j = -44; j <= 38; j++)
        for (int r = 0; r < 400; r++) {
            const double k = 100000.0 + (double)(next() % 900000ull) + 0.5;
            const double t = k * std::pow(10.0, j - 5);
            const float tf = (float)t;
            for (float v : {tf, std::nextafterf(tf, 0.0f), std::nextafterf(tf, 1e38f)}) bad += check(v, shown), bad += check(-v, shown);
            for (double v : {t, std::nextafter(t, 0.0), std::nextafter(t, 1e300)}) bad += check(v, shown);
            if (r == 0) {
                const double pw = std::pow(10.0, j);
                const float pf = (float)pw;
                for (float v : {pf, std::nextafterf(pf, 0.0f), std::nextafterf(pf, 1e38f)}) bad += check(v, shown);
                for (double v : {pw, std::nextafter(pw, 0.0), std::nextafter(pw, 1e300)}) bad += check(v, shown);
            }
        }
    for (int j = -300; j <= 300; j += 7) {
        const double pw = std::pow(10.0, j);
        for (double v : {pw, std::nextafter(pw, 0.0), std::nextafter(pw, 1e308), 9.999995 * pw, 1.2345649999999 * pw}) bad += check(v, shown);
    }
    std::printf("checked %ld rounds, %ld mismatches\n", n, bad);
    return bad ? 1 : 0;
}
