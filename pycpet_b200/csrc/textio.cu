// textio.cu -- host-side writers for the path's text outputs (no device code).
//
// After the kernels, PyCPET's per-frame cost is dominated by np.savetxt: `.top` files are written
// with the default "%.18e" (CPET/source/CPET.py:123), `_efield.dat` / `_esp.dat` with "%.3f" after a
// 7-line header (CPET/utils/io.py:50-109).  np.savetxt formats row by row in Python (~2-3 s per
// million rows).  cpet_write_rows produces byte-identical text with snprintf on all host cores:
// every element is widened to double exactly (float32 / float16 -> float64 is exact, which is what
// `fmt % tuple(row)` does through Python floats) and printed with the C conversion that Python's
// % operator specifies, so the files are interchangeable with the reference's.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include <cuda_fp16.h>

#include "cpet_internal.h"

namespace cpet {

static inline double elem_as_double(const void* data, int dtype, size_t i) {
    if (dtype == 0) return (double)reinterpret_cast<const float*>(data)[i];
    if (dtype == 1) return reinterpret_cast<const double*>(data)[i];
    // IEEE binary16 -> double, exact
    const uint16_t h = reinterpret_cast<const uint16_t*>(data)[i];
    const int sign = h >> 15, ex = (h >> 10) & 31, man = h & 1023;
    double v;
    if (ex == 0) v = ldexp((double)man, -24);
    else if (ex == 31) v = man ? NAN : INFINITY;
    else v = ldexp((double)(man | 1024), ex - 25);
    return sign ? -v : v;
}

static void format_span(const void* data, int dtype, size_t row0, size_t row1, int n_cols, const char* fmt,
                        std::string* out) {
    out->reserve((row1 - row0) * (size_t)n_cols * 26);
    char buf[512];
    for (size_t r = row0; r < row1; ++r) {
        for (int c = 0; c < n_cols; ++c) {
            const double v = elem_as_double(data, dtype, r * (size_t)n_cols + c);
            int len;
            if (v != v) len = snprintf(buf, sizeof(buf), "nan");     // Python never prints "-nan"
            else len = snprintf(buf, sizeof(buf), fmt, v);
            if (len < 0) len = 0;
            if (len > (int)sizeof(buf) - 1) len = (int)sizeof(buf) - 1;
            out->append(buf, (size_t)len);
            out->push_back(c + 1 < n_cols ? ' ' : '\n');
        }
    }
}

}  // namespace cpet

using namespace cpet;

extern "C" int cpet_write_rows(const char* path, const char* header, const void* data, int dtype, int64_t n_rows,
                               int n_cols, const char* fmt, int n_threads) {
    CPET_REQUIRE(path && fmt && (data || n_rows == 0), CPET_ERR_INVALID, "cpet_write_rows: NULL argument");
    CPET_REQUIRE(dtype >= 0 && dtype <= 2 && n_rows >= 0 && n_cols >= 1, CPET_ERR_INVALID, "bad array description");
    // exactly one floating conversion, e.g. "%.18e" or "%.3f" (what np.savetxt accepts per column)
    const char* pct = strchr(fmt, '%');
    CPET_REQUIRE(pct && !strchr(pct + 1, '%') && strlen(fmt) < 16 && strpbrk(pct, "eEfFgG"), CPET_ERR_INVALID,
                 "fmt must hold one e/f/g conversion");
    FILE* fh = fopen(path, "wb");
    CPET_REQUIRE(fh != nullptr, CPET_ERR_INVALID, "cannot open '%s' for writing", path);
    if (header && *header) fwrite(header, 1, strlen(header), fh);
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if ((int64_t)nt > n_rows / 4096 + 1) nt = (int)(n_rows / 4096 + 1);
    // bounded memory: format in waves of nt spans of <= 64k rows
    const size_t span = 65536;
    std::vector<std::string> bufs((size_t)nt);
    for (size_t base = 0; base < (size_t)n_rows; base += span * (size_t)nt) {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) {
            const size_t a = base + span * (size_t)t;
            if (a >= (size_t)n_rows) break;
            const size_t b = a + span < (size_t)n_rows ? a + span : (size_t)n_rows;
            bufs[(size_t)t].clear();
            th.emplace_back(format_span, data, dtype, a, b, n_cols, fmt, &bufs[(size_t)t]);
        }
        for (auto& x : th) x.join();
        for (size_t t = 0; t < th.size(); ++t) fwrite(bufs[t].data(), 1, bufs[t].size(), fh);
    }
    const bool ok = fclose(fh) == 0;
    CPET_REQUIRE(ok, CPET_ERR_INVALID, "write to '%s' failed", path);
    return CPET_OK;
}
