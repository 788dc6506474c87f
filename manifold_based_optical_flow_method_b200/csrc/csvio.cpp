// On-disk formats either side of the hot path: the reference moves
// its (T,N) signal and its (T-1,2N) velocity fields through pandas CSV files
//   load_potentials : pd.read_csv(path, sep=',', header='infer', index_col=0).values   (cof:203-207)
//   reshape_and_save_data : pd.DataFrame(data.reshape(n,-1)).to_csv(path)              (cof:314-320)
// At 164k vertices and 1000 frames that is 3 GB in and 6.5 GB out of text, minutes of pandas time
// next to a 4.5 s solve.  This file reads and writes exactly that dialect with all host threads:
//   write: header ",0,1,...,C-1", rows "r,v0,v1,..." with every double printed like Python's repr
//          (shortest round-trip digits; fixed notation for 1e-4 <= |v| < 1e16, else d.ddde+XX),
//          NaN as an empty field (pandas na_rep=''), inf / -inf;
//   read : header line skipped, first field of each row (the index) skipped, fields parsed with
//          std::from_chars (correctly rounded, i.e. pandas float_precision='round_trip'), empty
//          fields -> NaN.
#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <limits>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

#include "mof_b200.h"
#include "mof_error.h"

namespace {

// Python repr(float) into out (at most 32 chars); returns the length.
int format_repr(double v, char* out) {
    if (std::isnan(v)) return 0;                                   // pandas: empty field
    if (std::isinf(v)) {
        const char* s = v > 0 ? "inf" : "-inf";
        int n = (int)strlen(s);
        memcpy(out, s, n);
        return n;
    }
    char sci[40];
    auto res = std::to_chars(sci, sci + sizeof sci, v, std::chars_format::scientific);   // shortest round trip
    const char* p = sci;
    char* o = out;
    if (*p == '-') { *o++ = '-'; ++p; }
    // digits d[.ddd], exponent e[+-]XX
    char digits[24];
    int nd = 0;
    digits[nd++] = *p++;
    if (*p == '.') { ++p; while (*p != 'e') digits[nd++] = *p++; }
    ++p;                                                            // 'e'
    int esign = (*p == '-') ? -1 : 1;
    ++p;
    int ex = 0;
    while (p < res.ptr) ex = ex * 10 + (*p++ - '0');
    ex *= esign;
    if (ex >= -4 && ex < 16) {                                      // fixed notation (float_repr_style 'short')
        if (ex < 0) {
            *o++ = '0'; *o++ = '.';
            for (int k = 0; k < -ex - 1; ++k) *o++ = '0';
            for (int k = 0; k < nd; ++k) *o++ = digits[k];
        } else {
            for (int k = 0; k <= ex; ++k) *o++ = k < nd ? digits[k] : '0';
            *o++ = '.';
            if (nd > ex + 1) for (int k = ex + 1; k < nd; ++k) *o++ = digits[k];
            else *o++ = '0';
        }
    } else {                                                        // d[.ddd]e+XX, at least two exponent digits
        *o++ = digits[0];
        if (nd > 1) { *o++ = '.'; for (int k = 1; k < nd; ++k) *o++ = digits[k]; }
        *o++ = 'e';
        *o++ = ex < 0 ? '-' : '+';
        int ae = ex < 0 ? -ex : ex;
        char eb[8];
        int ne = 0;
        while (ae) { eb[ne++] = char('0' + ae % 10); ae /= 10; }
        while (ne < 2) eb[ne++] = '0';
        while (ne) *o++ = eb[--ne];
    }
    return int(o - out);
}

int put_int(int64_t v, char* out) {
    auto r = std::to_chars(out, out + 24, v);
    return int(r.ptr - out);
}

int pick_threads(int n_threads, int64_t rows) {
    int n = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (n < 1) n = 1;
    if ((int64_t)n > rows) n = (int)std::max<int64_t>(1, rows);
    return std::min(n, 64);
}

bool write_all(int fd, const char* p, size_t n) {
    while (n) {
        ssize_t w = ::write(fd, p, n);
        if (w <= 0) return false;
        p += w;
        n -= (size_t)w;
    }
    return true;
}

}  // namespace

extern "C" int mof_csv_write(const char* path, const double* data, int64_t rows, int64_t cols, int n_threads) {
    if (!path || (!data && rows * cols > 0) || rows < 0 || cols < 0) return mof_set_error(-1, "mof_csv_write: bad arguments");
    int fd = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return mof_set_error(-4, "mof_csv_write: cannot open %s", path);
    std::string head;
    head.reserve((size_t)cols * 7 + 2);
    char tmp[40];
    for (int64_t c = 0; c < cols; ++c) { head.push_back(','); head.append(tmp, put_int(c, tmp)); }
    if (cols == 0) head.append("\"\"");                             // pandas writes '""' for an empty column index
    head.push_back('\n');
    bool ok = write_all(fd, head.data(), head.size());
    const int nt = pick_threads(n_threads, rows);
    // rows are formatted in waves of nt chunks (bounded memory) and written in order
    const int64_t chunk_rows = std::max<int64_t>(1, std::min<int64_t>(64, (int64_t)(64 << 20) / std::max<int64_t>(1, cols * 24)));
    std::vector<std::string> bufs(nt);
    for (int64_t r0 = 0; ok && r0 < rows; r0 += chunk_rows * nt) {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) {
            const int64_t a = r0 + t * chunk_rows, b = std::min(rows, a + chunk_rows);
            bufs[t].clear();
            if (a >= b) continue;
            th.emplace_back([&, t, a, b] {
                std::string& s = bufs[t];
                s.resize((size_t)(b - a) * ((size_t)cols * 26 + 24));        // worst case: 25 chars + ',' per field
                char* o = &s[0];
                for (int64_t r = a; r < b; ++r) {
                    o += put_int(r, o);
                    const double* row = data + r * cols;
                    for (int64_t c = 0; c < cols; ++c) {
                        *o++ = ',';
                        o += format_repr(row[c], o);
                    }
                    *o++ = '\n';
                }
                s.resize((size_t)(o - s.data()));
            });
        }
        for (auto& x : th) x.join();
        for (int t = 0; ok && t < nt; ++t) ok = write_all(fd, bufs[t].data(), bufs[t].size());
    }
    if (::close(fd) != 0) ok = false;
    return ok ? 0 : mof_set_error(-4, "mof_csv_write: write to %s failed", path);
}

namespace {
struct Mapped {
    const char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    ~Mapped() {
        if (p && n) munmap((void*)p, n);
        if (fd >= 0) ::close(fd);
    }
};

int map_file(const char* path, Mapped& m) {
    m.fd = ::open(path, O_RDONLY);
    if (m.fd < 0) return mof_set_error(-4, "cannot open %s", path);
    struct stat st;
    if (fstat(m.fd, &st) != 0) return mof_set_error(-4, "cannot stat %s", path);
    m.n = (size_t)st.st_size;
    if (m.n == 0) return 0;
    void* q = mmap(nullptr, m.n, PROT_READ, MAP_PRIVATE, m.fd, 0);
    if (q == MAP_FAILED) { m.n = 0; return mof_set_error(-4, "cannot map %s", path); }
    m.p = (const char*)q;
    madvise(q, m.n, MADV_SEQUENTIAL);
    return 0;
}

// start offsets of the data lines (after the header line); empty trailing line ignored
void line_starts(const Mapped& m, std::vector<size_t>& starts, size_t& header_end) {
    const char* e = m.p + m.n;
    const char* h = (const char*)memchr(m.p, '\n', m.n);
    header_end = h ? size_t(h - m.p) : m.n;
    const char* q = h ? h + 1 : e;
    while (q < e) {
        starts.push_back(size_t(q - m.p));
        const char* nl = (const char*)memchr(q, '\n', size_t(e - q));
        q = nl ? nl + 1 : e;
    }
}
}  // namespace

extern "C" int mof_csv_dims(const char* path, int64_t* rows, int64_t* cols) {
    if (!path || !rows || !cols) return mof_set_error(-1, "mof_csv_dims: bad arguments");
    Mapped m;
    if (int rc = map_file(path, m)) return rc;
    std::vector<size_t> starts;
    size_t hend = 0;
    if (m.n) line_starts(m, starts, hend);
    int64_t commas = 0;
    for (size_t k = 0; k < hend; ++k) commas += m.p[k] == ',';
    *rows = (int64_t)starts.size();
    *cols = commas;                                                 // first field is the index column
    return 0;
}

extern "C" int mof_csv_read(const char* path, double* data, int64_t rows, int64_t cols, int n_threads) {
    if (!path || (!data && rows * cols > 0) || rows < 0 || cols < 0) return mof_set_error(-1, "mof_csv_read: bad arguments");
    Mapped m;
    if (int rc = map_file(path, m)) return rc;
    std::vector<size_t> starts;
    size_t hend = 0;
    if (m.n) line_starts(m, starts, hend);
    if ((int64_t)starts.size() != rows) return mof_set_error(-5, "mof_csv_read: %s has %lld data rows, expected %lld", path, (long long)starts.size(), (long long)rows);
    const int nt = pick_threads(n_threads, rows);
    std::atomic<int64_t> bad{-1};
    std::vector<std::thread> th;
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int t = 0; t < nt; ++t) {
        th.emplace_back([&, t] {
            for (int64_t r = t; r < rows; r += nt) {
                const char* q = m.p + starts[r];
                const char* e = (r + 1 < rows) ? m.p + starts[r + 1] : m.p + m.n;
                while (e > q && (e[-1] == '\n' || e[-1] == '\r')) --e;
                const char* c = (const char*)memchr(q, ',', size_t(e - q));      // skip the index field
                double* out = data + r * cols;
                int64_t k = 0;
                while (c && k < cols) {
                    const char* f = c + 1;
                    const char* nx = (const char*)memchr(f, ',', size_t(e - f));
                    const char* fe = nx ? nx : e;
                    while (f < fe && *f == ' ') ++f;
                    if (f == fe) {
                        out[k] = nan;
                    } else {
                        const char* g = (*f == '+') ? f + 1 : f;
                        auto res = std::from_chars(g, fe, out[k]);
                        if (res.ec != std::errc() ) {
                            // from_chars does not take "inf"/"nan" spellings with every libstdc++; handle them
                            std::string s(g, fe);
                            if (s == "inf" || s == "Inf" || s == "infinity") out[k] = INFINITY;
                            else if (s == "-inf" || s == "-Inf" || s == "-infinity") out[k] = -INFINITY;
                            else if (s == "nan" || s == "NaN" || s == "NA" || s == "null") out[k] = nan;
                            else { bad.store(r); return; }
                        }
                    }
                    ++k;
                    c = nx;
                }
                if (k != cols || c) { bad.store(r); return; }
            }
        });
    }
    for (auto& x : th) x.join();
    if (bad.load() >= 0) return mof_set_error(-5, "mof_csv_read: cannot parse data row %lld of %s (expected %lld numeric fields)", (long long)bad.load(), path, (long long)cols);
    return 0;
}
