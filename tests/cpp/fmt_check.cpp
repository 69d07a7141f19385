// Test program: lbm::vtk_put (what lbm_output_save writes) against a real `ostream << value << ' '`
// (what the reference's outputSave writes, ldc.cu:603-607) on N pseudo-random bit patterns of float
// and double plus the special values.  Prints the number of mismatches; exit code 0 iff none.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <sstream>
#include <string>

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

int main(int argc, char **argv) {
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
    std::printf("checked %ld rounds, %ld mismatches\n", n, bad);
    return bad ? 1 : 0;
}
