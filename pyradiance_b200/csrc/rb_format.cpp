// rb_format.cpp -- ASCII matrix output at C speed.
//
// rcontrib writes "%.6e\t" per component and "\n" per record (rt/rc2.c:304-312), rtrace "%e\t"
// (rt/rtrace.c:907-918 puta()) -- the same conversion.  In Python that is ~0.7 us per value: 30 s for
// the 100k x 145 x 3 matrix the GPU computes in 1.3 s.  Here: std::to_chars (correctly rounded, the
// text printf gives) on all host cores, each thread formatting a block of rows into its own buffer.
#include <algorithm>
#include <charconv>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "../../include/rb200.h"

namespace {
// style 0: "v\tv\t...v\t\n" (rcontrib / rtrace); style 1: "r g b\tr g b\t...r g b\n" (cm_write, cmatrix.c:492-498)
template <class T>
void format_rows(const T* v, size_t r0, size_t r1, size_t per_row, int style, std::string& out) {
    out.reserve((r1 - r0) * (per_row * 15 + 1));
    char buf[64];
    for (size_t r = r0; r < r1; r++) {
        const T* p = v + r * per_row;
        for (size_t k = 0; k < per_row; k++) {
            // std::to_chars(scientific, 6) == printf("%e"): correctly rounded, two-digit exponent at least
            auto res = std::to_chars(buf, buf + sizeof(buf) - 1, (double)p[k], std::chars_format::scientific, 6);
            if (style == 0) *res.ptr++ = '\t';
            else *res.ptr++ = (k % 3 != 2) ? ' ' : (k + 1 == per_row ? '\n' : '\t');
            out.append(buf, (size_t)(res.ptr - buf));
        }
        if (style == 0) out.push_back('\n');
    }
}
}  // namespace

// Returns the number of bytes the text takes; writes it when `out` is large enough (outlen >= result).
extern "C" size_t rb_format_ascii(const void* values, int is_double, size_t nrows, size_t per_row, int style, char* out,
                                  size_t outlen) {
    if (nrows == 0) return 0;
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = std::max<size_t>(1, std::min<size_t>({(size_t)(hw ? hw : 4), (size_t)32, (nrows * per_row) / 20000 + 1}));
    std::vector<std::string> parts(nt);
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++) {
        size_t r0 = nrows * t / nt, r1 = nrows * (t + 1) / nt;
        th.emplace_back([=, &parts]() {
            if (is_double) format_rows((const double*)values, r0, r1, per_row, style, parts[t]);
            else format_rows((const float*)values, r0, r1, per_row, style, parts[t]);
        });
    }
    for (auto& x : th) x.join();
    size_t total = 0;
    for (auto& p : parts) total += p.size();
    if (out && outlen >= total) {
        size_t off = 0;
        for (auto& p : parts) { memcpy(out + off, p.data(), p.size()); off += p.size(); }
    }
    return total;
}
