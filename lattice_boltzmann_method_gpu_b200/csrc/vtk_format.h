// Text form of one value in the reference's ASCII VTK bodies.  The reference prints with
// `ofs << value << ' '` (ldc.cu:603-607, bifurcation.cu:1140-1150): default ostream formatting =
// printf("%g") with 6 significant digits, which is what std::to_chars(general, 6) produces -- without a
// locale lookup and a virtual call per value.  Host only; shared with tests/cpp/fmt_check.cpp, which
// compares it with a real ostream on millions of bit patterns.
#pragma once
#include <charconv>
#include <string>

namespace lbm {

template <typename V>
inline void vtk_put(std::string &buf, V v) {
    char tmp[48];
    auto r = std::to_chars(tmp, tmp + sizeof tmp, v, std::chars_format::general, 6);
    buf.append(tmp, r.ptr);
    buf.push_back(' ');
}

}  // namespace lbm
