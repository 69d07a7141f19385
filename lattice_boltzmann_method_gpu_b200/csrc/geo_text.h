// Reader for the reference's geo.txt: whitespace-separated integer labels, one per lattice node, written by the
// MATLAB front end and read by the reference with `fscanf(f, "%d ", ...)` in a triple loop (bif.cu:50-61 x fastest,
// cor.cu:45-56 y fastest).  fscanf costs 50 - 100 ns per token -- seconds for the coronary box (31.5 M tokens),
// minutes at 1024^3 -- so the file is mapped and parsed by several threads instead: the byte range is cut at
// whitespace, a first pass counts the tokens of every piece (which gives each piece the index of its first
// token), a second pass converts them.  Same acceptance as the fscanf loop: optional sign, decimal digits, any
// amount of blanks / tabs / newlines between tokens, tokens beyond the expected count ignored; anything else
// stops the parse there (after the digits in front of it, as fscanf would) and the caller sees a short count.
// Host only; shared with tests/cpp/geo_parse_check.cpp, which compares it with the fscanf loop.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <thread>
#include <vector>

namespace lbm {

struct MappedFile {
    const char *p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool open_ro(const char *path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        n = (size_t)st.st_size;
        if (n == 0) return true;
        void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return false;
        madvise(m, n, MADV_SEQUENTIAL);
        p = (const char *)m;
        return true;
    }
    ~MappedFile() {
        if (p) munmap((void *)p, n);
        if (fd >= 0) ::close(fd);
    }
};

inline bool geo_is_space(char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// Calls store(t, value) for token t = 0 .. min(tokens, max_tokens) - 1 (from several threads, each t once) and
// returns the number of well-formed tokens before the first malformed one (all of them if there is none).
template <typename Store>
long parse_int_tokens(const char *p, size_t n, long max_tokens, int max_threads, Store store) {
    if (n == 0) return 0;
    const int T = (int)std::max<size_t>(1, std::min<size_t>({(size_t)std::max(1, max_threads), (size_t)std::max(1u, std::thread::hardware_concurrency()), n / (1u << 20) + 1}));
    std::vector<size_t> cut((size_t)T + 1);
    cut[0] = 0, cut[(size_t)T] = n;
    for (int k = 1; k < T; k++) {
        size_t b = std::max(cut[(size_t)k - 1], n / (size_t)T * (size_t)k);
        while (b < n && !geo_is_space(p[b])) b++;  // never inside a token
        cut[(size_t)k] = b;
    }
    std::vector<long> first((size_t)T + 1, 0);
    auto run = [&](auto fn) {
        std::vector<std::thread> pool;
        for (int k = 1; k < T; k++) pool.emplace_back(fn, k);
        fn(0);
        for (auto &th : pool) th.join();
    };
    run([&](int k) {
        long c = 0;
        bool in = false;
        for (size_t i = cut[(size_t)k]; i < cut[(size_t)k + 1]; i++) {
            const bool s = geo_is_space(p[i]);
            c += (!s && !in);
            in = !s;
        }
        first[(size_t)k + 1] = c;
    });
    for (int k = 0; k < T; k++) first[(size_t)k + 1] += first[(size_t)k];
    std::atomic<long> bad(first[(size_t)T]);  // index of the first malformed token
    run([&](int k) {
        long t = first[(size_t)k];
        size_t i = cut[(size_t)k];
        const size_t e = cut[(size_t)k + 1];
        while (i < e && t < max_tokens) {
            while (i < e && geo_is_space(p[i])) i++;
            if (i >= e) break;
            bool neg = false;
            if (p[i] == '-' || p[i] == '+') neg = p[i] == '-', i++;
            long v = 0;
            int nd = 0;
            while (i < e && p[i] >= '0' && p[i] <= '9' && nd < 10) v = v * 10 + (p[i] - '0'), i++, nd++;
            // like fscanf("%d"): the digits in front of a malformed tail still count as a token, then reading stops
            const bool tail = i < e && !geo_is_space(p[i]);
            if (nd) store(t, (int)(neg ? -v : v)), t++;
            if (nd == 0 || tail) {
                long cur = bad.load();
                while (t < cur && !bad.compare_exchange_weak(cur, t)) {
                }
                return;
            }
        }
    });
    return std::min(bad.load(), first[(size_t)T]);
}

}  // namespace lbm
