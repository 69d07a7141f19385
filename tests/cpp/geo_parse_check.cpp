// Test program: lbm::parse_int_tokens (what lbm_geo_pre reads a geo.txt with, csrc/geo_text.h) against the
// reference's own way of reading it -- fscanf(f, "%d ", &tmp) token by token (bif.cu:50-61) -- on files with
// single blanks, mixed whitespace, signs, multi-digit labels, missing / surplus / malformed tokens.
//   geo_parse_check <dir>   writes its files there; prints the number of mismatches; exit code 0 iff none
//   geo_parse_check <dir> --time N   also times both readers on an N-token file
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "geo_text.h"

static long read_fscanf(const char *path, long expect, std::vector<int> &out) {
    out.assign((size_t)expect, 0);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    long cnt = 0;
    int tmp;
    for (long t = 0; t < expect; t++) {
        if (fscanf(f, "%d ", &tmp) != 1) break;
        out[(size_t)t] = tmp, cnt++;
    }
    fclose(f);
    return cnt;
}
static long read_mapped(const char *path, long expect, std::vector<int> &out, int threads) {
    out.assign((size_t)expect, 0);
    lbm::MappedFile f;
    if (!f.open_ro(path)) return -1;
    int *o = out.data();
    return std::min(expect, lbm::parse_int_tokens(f.p, f.n, expect, threads, [o](long t, int v) { o[t] = v; }));
}
static int compare(const std::string &path, const std::string &text, long expect, const char *what) {
    FILE *f = fopen(path.c_str(), "w");
    fwrite(text.data(), 1, text.size(), f);
    fclose(f);
    std::vector<int> a, b;
    const long ca = read_fscanf(path.c_str(), expect, a);
    int bad = 0;
    for (int threads : {1, 3, 16}) {
        const long cb = read_mapped(path.c_str(), expect, b, threads);
        // after a short / malformed file only the tokens both readers accepted are compared
        const long n = std::min(ca, cb);
        if (ca != cb || memcmp(a.data(), b.data(), (size_t)n * sizeof(int)) != 0) {
            printf("mismatch (%s, %d threads): fscanf read %ld tokens, parser %ld\n", what, threads, ca, cb);
            bad++;
        }
    }
    return bad;
}

int main(int argc, char **argv) {
    const std::string dir = argc > 1 ? argv[1] : ".";
    int bad = 0;
    unsigned long long s = 0x9E3779B97F4A7C15ull;
    auto next = [&s]() {
        s ^= s << 13, s ^= s >> 7, s ^= s << 17;
        return s;
    };
    // the MATLAB writer's form: one digit, one blank
    for (long n : {1L, 7L, 4096L, 3000001L}) {
        std::string t;
        for (long i = 0; i < n; i++) t += (char)('0' + next() % 8), t += ' ';
        bad += compare(dir + "/geo_a.txt", t, n, "digit blank");
        bad += compare(dir + "/geo_b.txt", t, n + 5, "short file");
        bad += compare(dir + "/geo_c.txt", t, n > 3 ? n - 3 : n, "surplus tokens");
    }
    {  // signs, several digits, every kind of whitespace, no trailing blank, leading blanks
        std::string t = "  \n";
        const long n = 2500000;
        for (long i = 0; i < n; i++) {
            const int v = (int)(next() % 2001) - 1000;
            t += std::to_string(v);
            if (v > 0 && next() % 7 == 0) t.insert(t.size() - std::to_string(v).size(), "+");
            const char *ws[] = {" ", "\n", "\t", "  ", " \r\n", "\n\n\n"};
            if (i + 1 < n) t += ws[next() % 6];
        }
        bad += compare(dir + "/geo_d.txt", t, n, "mixed");
    }
    for (const char *txt : {"1 2 3 x 4 5", "1 2 3- 4", "1 2 3 4.5 6", "", "   ", "12", "-", "1 2 -", "7 8 9abc"}) bad += compare(dir + "/geo_e.txt", txt, 6, txt);
    if (argc > 3 && !strcmp(argv[2], "--time")) {
        const long n = atol(argv[3]);
        std::string t((size_t)n * 2, ' ');
        for (long i = 0; i < n; i++) t[(size_t)i * 2] = (char)('0' + next() % 5);
        const std::string path = dir + "/geo_t.txt";
        FILE *f = fopen(path.c_str(), "w");
        fwrite(t.data(), 1, t.size(), f);
        fclose(f);
        std::vector<int> a, b;
        auto t0 = std::chrono::steady_clock::now();
        read_fscanf(path.c_str(), n, a);
        auto t1 = std::chrono::steady_clock::now();
        read_mapped(path.c_str(), n, b, 16);
        auto t2 = std::chrono::steady_clock::now();
        printf("%ld tokens: fscanf %.2f s, mapped + threads %.3f s, equal %d\n", n, std::chrono::duration<double>(t1 - t0).count(),
               std::chrono::duration<double>(t2 - t1).count(), (int)(a == b));
        remove(path.c_str());
    }
    for (const char *nm : {"/geo_a.txt", "/geo_b.txt", "/geo_c.txt", "/geo_d.txt", "/geo_e.txt"}) remove((dir + nm).c_str());
    printf("%d mismatches\n", bad);
    return bad ? 1 : 0;
}
