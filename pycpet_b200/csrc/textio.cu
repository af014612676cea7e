// textio.cu -- host-side writers and readers for the path's text outputs (no device code).
//
// After the kernels, PyCPET's per-frame cost is dominated by np.savetxt: `.top` files are written
// with the default "%.18e" (CPET/source/CPET.py:123), `_efield.dat` / `_esp.dat` with "%.3f" after a
// 7-line header (CPET/utils/io.py:50-109).  np.savetxt formats row by row in Python (~2-3 s per
// million rows).  cpet_write_rows produces byte-identical text with std::to_chars / snprintf on all host cores:
// every element is widened to double exactly (float32 / float16 -> float64 is exact, which is what
// `fmt % tuple(row)` does through Python floats) and printed with the C conversion that Python's
// % operator specifies, so the files are interchangeable with the reference's.
//
// The way back in costs as much: make_histograms (CPET/utils/calculator.py:596-718) walks every
// `.top` file three times with `for line in fh: float(line.split()[k])`.  cpet_count_rows /
// cpet_read_rows do the same parse -- '#' lines skipped, the leading columns of every other line
// converted with a correctly rounded decimal->double conversion (std::from_chars, the value
// Python's float() returns) -- on all host cores, straight into the caller's float64 array.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <charconv>
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

// "%.<N>e" / "%.<N>f" (every format the path writes) go through std::to_chars, which prints the same
// exact decimal expansion as printf about three times faster; any other conversion keeps snprintf.
static bool simple_format(const char* fmt, std::chars_format* kind, int* precision) {
    if (fmt[0] != '%' || fmt[1] != '.') return false;
    const char* p = fmt + 2;
    int prec = 0, digits = 0;
    while (*p >= '0' && *p <= '9' && digits < 3) { prec = prec * 10 + (*p - '0'); ++p; ++digits; }
    if (!digits || prec > 60 || (*p != 'e' && *p != 'f') || p[1] != '\0') return false;
    *kind = *p == 'e' ? std::chars_format::scientific : std::chars_format::fixed;
    *precision = prec;
    return true;
}

static void format_span(const void* data, int dtype, size_t row0, size_t row1, int n_cols, const char* fmt,
                        std::string* out) {
    out->reserve((row1 - row0) * (size_t)n_cols * 26);
    char buf[512];
    std::chars_format kind = std::chars_format::general;
    int prec = 0;
    const bool fast = simple_format(fmt, &kind, &prec);
    for (size_t r = row0; r < row1; ++r) {
        for (int c = 0; c < n_cols; ++c) {
            const double v = elem_as_double(data, dtype, r * (size_t)n_cols + c);
            int len;
            if (v != v) len = snprintf(buf, sizeof(buf), fmt, copysign(NAN, 1.0));   // Python never prints "-nan"
            else if (fast) {
                const std::to_chars_result res = std::to_chars(buf, buf + sizeof(buf) - 1, v, kind, prec);
                len = res.ec == std::errc() ? (int)(res.ptr - buf) : snprintf(buf, sizeof(buf), fmt, v);
            } else len = snprintf(buf, sizeof(buf), fmt, v);
            if (len < 0) len = 0;
            if (len > (int)sizeof(buf) - 1) len = (int)sizeof(buf) - 1;
            out->append(buf, (size_t)len);
            out->push_back(c + 1 < n_cols ? ' ' : '\n');
        }
    }
}

}  // namespace cpet

namespace cpet {

// ---- reader ------------------------------------------------------------------------------------
static inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// [a, b) holds whole lines.  A data line is one that does not start with '#' (the reference's
// `line.startswith("#")`, UC:605) and is not blank.  With out == nullptr the lines are only counted.
// Returns the number of data lines, or -(1 + index of the first malformed data line of the span).
static int64_t parse_span(const char* a, const char* b, int n_cols, double* out) {
    int64_t rows = 0;
    while (a < b) {
        const char* eol = static_cast<const char*>(memchr(a, '\n', (size_t)(b - a)));
        if (!eol) eol = b;
        const char* p = a;
        a = eol + 1;
        if (*p == '#') continue;
        while (p < eol && is_blank(*p)) ++p;
        if (p == eol) continue;
        if (out) {
            for (int c = 0; c < n_cols; ++c) {
                while (p < eol && is_blank(*p)) ++p;
                if (p < eol && *p == '+') ++p;                       // float("+1.0") is legal, from_chars refuses it
                double v;
                const std::from_chars_result r = std::from_chars(p, eol, v, std::chars_format::general);
                if (r.ec == std::errc::result_out_of_range) {
                    // float() returns +-inf on overflow and (signed) zero / denormal on underflow
                    std::string tok(p, r.ptr);
                    v = strtod(tok.c_str(), nullptr);
                } else if (r.ec != std::errc() || (r.ptr < eol && !is_blank(*r.ptr))) {
                    return -(rows + 1);
                }
                out[(size_t)rows * (size_t)n_cols + (size_t)c] = v;
                p = r.ptr;
            }
        }
        ++rows;
    }
    return rows;
}

struct TextFile {
    std::vector<char> buf;
    std::vector<const char*> cut;        // span boundaries, each at the start of a line
    int load(const char* path, int n_threads) {
        FILE* fh = fopen(path, "rb");
        CPET_REQUIRE(fh != nullptr, CPET_ERR_INVALID, "cannot open '%s' for reading", path);
        fseek(fh, 0, SEEK_END);
        const long size = ftell(fh);
        fseek(fh, 0, SEEK_SET);
        if (size < 0) fclose(fh);
        CPET_REQUIRE(size >= 0, CPET_ERR_INVALID, "cannot size '%s'", path);
        buf.resize((size_t)size + 1);
        const size_t got = size ? fread(buf.data(), 1, (size_t)size, fh) : 0;
        fclose(fh);
        CPET_REQUIRE(got == (size_t)size, CPET_ERR_INVALID, "short read on '%s'", path);
        buf[(size_t)size] = '\n';
        int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
        if (nt < 1) nt = 1;
        if ((size_t)nt > (size_t)size / (1u << 20) + 1) nt = (int)((size_t)size / (1u << 20) + 1);
        const char* base = buf.data();
        const char* end = base + size;
        cut.assign(1, base);
        for (int t = 1; t < nt; ++t) {
            const char* p = base + (size_t)size * (size_t)t / (size_t)nt;
            if (p <= cut.back()) continue;
            const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
            if (!nl || nl + 1 >= end) break;
            if (nl + 1 > cut.back()) cut.push_back(nl + 1);
        }
        cut.push_back(end);
        return CPET_OK;
    }
    // counts[i] = data lines of span i (out == nullptr) / parses span i to out + offsets[i] * n_cols
    void run(int n_cols, double* out, const std::vector<int64_t>* offsets, std::vector<int64_t>* result) {
        const size_t n = cut.size() - 1;
        result->assign(n, 0);
        std::vector<std::thread> th;
        for (size_t i = 0; i < n; ++i)
            th.emplace_back([=]() {
                double* o = out ? out + (size_t)(*offsets)[i] * (size_t)n_cols : nullptr;
                (*result)[i] = parse_span(cut[i], cut[i + 1], n_cols, o);
            });
        for (auto& x : th) x.join();
    }
};

}  // namespace cpet

using namespace cpet;

extern "C" int cpet_count_rows(const char* path, int64_t* n_rows, int n_threads) {
    CPET_REQUIRE(path && n_rows, CPET_ERR_INVALID, "cpet_count_rows: NULL argument");
    TextFile f;
    if (int rc = f.load(path, n_threads)) return rc;
    std::vector<int64_t> counts;
    f.run(0, nullptr, nullptr, &counts);
    int64_t total = 0;
    for (int64_t c : counts) total += c;
    *n_rows = total;
    return CPET_OK;
}

extern "C" int cpet_read_rows(const char* path, int n_cols, int64_t n_rows, double* out, int n_threads) {
    CPET_REQUIRE(path && (out || n_rows == 0), CPET_ERR_INVALID, "cpet_read_rows: NULL argument");
    CPET_REQUIRE(n_cols >= 1 && n_rows >= 0, CPET_ERR_INVALID, "bad array description");
    TextFile f;
    if (int rc = f.load(path, n_threads)) return rc;
    std::vector<int64_t> counts, offsets, parsed;
    f.run(0, nullptr, nullptr, &counts);
    offsets.assign(counts.size(), 0);
    int64_t total = 0;
    for (size_t i = 0; i < counts.size(); ++i) { offsets[i] = total; total += counts[i]; }
    CPET_REQUIRE(total == n_rows, CPET_ERR_INVALID, "'%s' holds %lld data lines, the caller expects %lld", path,
                 (long long)total, (long long)n_rows);
    if (n_rows == 0) return CPET_OK;
    f.run(n_cols, out, &offsets, &parsed);
    for (size_t i = 0; i < parsed.size(); ++i)
        CPET_REQUIRE(parsed[i] >= 0, CPET_ERR_INVALID, "'%s': data line %lld does not hold %d numbers", path,
                     (long long)(offsets[i] - parsed[i]), n_cols);
    return CPET_OK;
}

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
