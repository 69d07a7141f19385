// Text form of one value in the reference's ASCII VTK bodies.  The reference prints with
// `ofs << value << ' '` (ldc.cu:603-607, bifurcation.cu:1140-1150): default ostream formatting =
// printf("%g") with 6 significant digits, which is what std::to_chars(general, 6) produces -- without a
// locale lookup and a virtual call per value.
//
// vtk_write() produces the same characters several times faster for the values a run prints: it scales the
// value to an integer of six digits in double arithmetic, and only when the scaled value lies so close to a
// rounding tie that double arithmetic could not decide the digit (or the value is not finite, or outside
// 1e-280 .. 1e280) does it hand over to std::to_chars.  At 64^3 a dump is 0.65 - 1 M values and the
// reference dumps every 500 steps: with to_chars alone the drivers spent more CPU time formatting than the
// GPU spent stepping.
//
// Host only; shared with tests/cpp/fmt_check.cpp, which compares both with a real ostream on millions of
// bit patterns and vtk_write with to_chars on every float there is (--exhaustive).
#pragma once
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>

namespace lbm {

constexpr int VTK_MAX_CHARS = 32;  // per value, blank included (to_chars of a double: 14 at most)

namespace fmt_detail {
constexpr int P10_MAX = 300;
struct Pow10 {
    double v[2 * P10_MAX + 1];
    Pow10() {
        char s[16];
        for (int k = -P10_MAX; k <= P10_MAX; k++) {  // correctly rounded powers of ten
            std::snprintf(s, sizeof s, "1e%d", k);
            v[k + P10_MAX] = std::strtod(s, nullptr);
        }
    }
    double operator()(int k) const { return v[k + P10_MAX]; }
};
inline const Pow10 &pow10_table() {
    static const Pow10 t;
    return t;
}
struct Pairs {
    char c[200];
    constexpr Pairs() : c() {
        for (int i = 0; i < 100; i++) c[2 * i] = (char)('0' + i / 10), c[2 * i + 1] = (char)('0' + i % 10);
    }
};
constexpr Pairs PAIRS{};
}  // namespace fmt_detail

template <typename V>
inline char *vtk_write_slow(char *p, V v) {
    auto r = std::to_chars(p, p + VTK_MAX_CHARS - 1, v, std::chars_format::general, 6);
    *r.ptr = ' ';
    return r.ptr + 1;
}

// `%g ` of a float or double at p (VTK_MAX_CHARS bytes available); returns the end
template <typename V>
inline char *vtk_write(char *p, V value) {
    const double v0 = (double)value;  // exact for float
    uint64_t bits;
    std::memcpy(&bits, &v0, 8);
    const int e2 = (int)((bits >> 52) & 0x7ff) - 1023;
    if ((bits << 1) == 0) {  // +-0
        if (bits >> 63) *p++ = '-';
        *p++ = '0', *p++ = ' ';
        return p;
    }
    if (e2 < -930 || e2 > 930) return vtk_write_slow(p, value);  // inf, nan, subnormal doubles, |v| beyond 1e+-280
    char *const start = p;
    *p = '-', p += bits >> 63;  // no branch on the sign: it is as good as random in a velocity field
    const double v = std::fabs(v0);
    static const fmt_detail::Pow10 &P10 = fmt_detail::pow10_table();
    // floor(log10 v) is e10 or e10 + 1
    int e10 = (e2 * 1233) >> 12;
    double scaled = v * P10(5 - e10);
    if (scaled >= 1e6) scaled = v * P10(5 - ++e10);
    else if (scaled < 1e5) scaled = v * P10(5 - --e10);
    if (!(scaled >= 1e5 && scaled < 1e6)) return vtk_write_slow(start, value);
    uint32_t n = (uint32_t)scaled;  // truncation = floor, scaled > 0
    const double fr = scaled - (double)n;
    // scaled carries a relative error below 3e-16 (one rounded table entry, one product): < 1e-9 absolute
    if (std::fabs(fr - 0.5) < 1e-6) return vtk_write_slow(start, value);
    n += fr > 0.5 ? 1u : 0u;
    if (n >= 1000000u) n = 100000u, e10++;
    char dg[12] = {};
    const uint32_t a = n / 10000u, bc = n - a * 10000u, b = bc / 100u, c = bc - b * 100u;
    std::memcpy(dg, fmt_detail::PAIRS.c + 2 * a, 2), std::memcpy(dg + 2, fmt_detail::PAIRS.c + 2 * b, 2);
    std::memcpy(dg + 4, fmt_detail::PAIRS.c + 2 * c, 2);
    // significant digits left after the trailing zeros are dropped: the pair table again, no loop
    const int nd = c ? 6 - (c % 10u == 0) : b ? 4 - (b % 10u == 0) : 2 - (a % 10u == 0);
    // whole groups of six are copied and the cursor advanced by the count that is wanted: every value has
    // VTK_MAX_CHARS bytes to write into
    if (e10 >= -4 && e10 < 6) {
        if (e10 >= 0) {
            std::memcpy(p, dg, 6), p += e10 + 1;
            if (nd > e10 + 1) {
                *p++ = '.';
                std::memcpy(p, dg + e10 + 1, 6), p += nd - (e10 + 1);
            }
        } else {
            std::memcpy(p, "0.0000", 6), p += 1 - e10;
            std::memcpy(p, dg, 6), p += nd;
        }
    } else {
        *p++ = dg[0];
        if (nd > 1) {
            *p++ = '.';
            std::memcpy(p, dg + 1, 5), p += nd - 1;
        }
        *p++ = 'e';
        int x = e10;
        if (x < 0) *p++ = '-', x = -x;
        else *p++ = '+';
        if (x >= 100) *p++ = (char)('0' + x / 100), x %= 100;
        std::memcpy(p, fmt_detail::PAIRS.c + 2 * x, 2), p += 2;
    }
    *p++ = ' ';
    return p;
}

template <typename V>
inline void vtk_put(std::string &buf, V v) {
    char tmp[VTK_MAX_CHARS];
    buf.append(tmp, (size_t)(vtk_write(tmp, v) - tmp));
}

}  // namespace lbm
