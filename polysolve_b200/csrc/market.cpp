// Matrix Market reader / writer behind include/psb200_io.h (host only). Parsing is done on the whole file in memory
// with strtol / strtod (no iostream extraction in the entry loop): a 70 M-entry fixture reads at disk speed.
#include "../../include/psb200.h"
#include "../../include/psb200_io.h"

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

struct psb200_market
{
    int64_t rows = 0, cols = 0;
    std::vector<int32_t> outer, inner;
    std::vector<double> vals;
};

namespace {
thread_local std::string g_err;

int fail(const std::string &msg)
{
    g_err = msg;
    return PSB200_ERR_INVALID;
}

bool read_file(const char *path, std::string &buf)
{
    FILE *f = std::fopen(path, "rb");
    if (!f)
        return false;
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    buf.resize(sz > 0 ? (size_t)sz : 0);
    const size_t got = sz > 0 ? std::fread(&buf[0], 1, (size_t)sz, f) : 0;
    std::fclose(f);
    buf.resize(got);
    return true;
}

std::string lower(std::string s)
{
    for (char &c : s)
        c = (char)std::tolower((unsigned char)c);
    return s;
}

// next line that is neither blank nor a comment; returns false at end of buffer
bool next_data_line(const char *&p, const char *end, const char *&line_end)
{
    while (p < end)
    {
        const char *e = (const char *)std::memchr(p, '\n', (size_t)(end - p));
        if (!e)
            e = end;
        const char *q = p;
        while (q < e && (*q == ' ' || *q == '\t' || *q == '\r'))
            ++q;
        if (q < e && *q != '%')
        {
            p = q;
            line_end = e;
            return true;
        }
        p = e < end ? e + 1 : end;
    }
    return false;
}

struct Entry
{
    int32_t r, c;
    double v;
};

// column-major compressed storage with ascending rows and summed duplicates (SparseMatrix::setFromTriplets)
void compress(int64_t rows, int64_t cols, std::vector<Entry> &e, psb200_market &m)
{
    m.rows = rows;
    m.cols = cols;
    std::vector<int64_t> cnt((size_t)cols + 1, 0);
    for (const Entry &t : e)
        cnt[(size_t)t.c + 1]++;
    for (int64_t c = 0; c < cols; ++c)
        cnt[(size_t)c + 1] += cnt[(size_t)c];
    // counting sort by column (stable), then sort rows inside each column (stable, so duplicates add in file order)
    std::vector<Entry> byc(e.size());
    {
        std::vector<int64_t> cur(cnt.begin(), cnt.end() - 1);
        for (const Entry &t : e)
            byc[(size_t)cur[(size_t)t.c]++] = t;
    }
    m.outer.assign((size_t)cols + 1, 0);
    m.inner.clear();
    m.vals.clear();
    m.inner.reserve(e.size());
    m.vals.reserve(e.size());
    for (int64_t c = 0; c < cols; ++c)
    {
        auto b = byc.begin() + cnt[(size_t)c], en = byc.begin() + cnt[(size_t)c + 1];
        std::stable_sort(b, en, [](const Entry &x, const Entry &y) { return x.r < y.r; });
        for (auto it = b; it != en; ++it)
        {
            if (it != b && (it - 1)->r == it->r)
                m.vals.back() += it->v;
            else
            {
                m.inner.push_back(it->r);
                m.vals.push_back(it->v);
            }
        }
        m.outer[(size_t)c + 1] = (int32_t)m.inner.size();
    }
}
} // namespace

extern "C" {

const char *psb200_market_last_error(void) { return g_err.c_str(); }

static int market_load_impl(const char *path, int symmetric_mode, psb200_market_handle *out, int64_t *rows, int64_t *cols, int64_t *nnz)
{
    if (!path || !out)
        return fail("psb200_market_load: null argument");
    *out = nullptr;
    std::string buf;
    if (!read_file(path, buf))
        return fail(std::string("psb200_market_load: cannot open ") + path);
    const char *p = buf.data(), *end = buf.data() + buf.size();
    bool pattern = false, header_sym = false;
    if (buf.compare(0, 14, "%%MatrixMarket") == 0)
    {
        const char *e = (const char *)std::memchr(p, '\n', buf.size());
        const std::string h = lower(std::string(p, e ? e : end));
        if (h.find("coordinate") == std::string::npos)
            return fail("psb200_market_load: only the coordinate format holds a sparse matrix (use psb200_market_load_vector for array files)");
        if (h.find("complex") != std::string::npos)
            return fail("psb200_market_load: complex matrices are not supported");
        if (h.find("skew-symmetric") != std::string::npos || h.find("hermitian") != std::string::npos)
            return fail("psb200_market_load: skew-symmetric / hermitian storage is not supported");
        pattern = h.find("pattern") != std::string::npos;
        header_sym = h.find("symmetric") != std::string::npos;
    }
    const bool mirror = symmetric_mode == 1 || (symmetric_mode < 0 && header_sym);
    const char *le = nullptr;
    if (!next_data_line(p, end, le))
        return fail("psb200_market_load: missing size line");
    char *q = nullptr;
    const long long M = std::strtoll(p, &q, 10);
    const long long N = std::strtoll(q, &q, 10);
    const long long L = std::strtoll(q, &q, 10);
    if (q > le || M < 0 || N < 0 || L < 0 || M > 0x7ffffffeLL || N > 0x7ffffffeLL)
        return fail("psb200_market_load: bad size line");
    p = le < end ? le + 1 : end;
    // an entry line needs at least 4 bytes ("1 1\n"): a size line that announces more than the file can hold is wrong,
    // and must not be trusted with a reservation
    if (L > (long long)(buf.size() / 4) + 1)
        return fail("psb200_market_load: the size line announces " + std::to_string(L) + " entries, the file is too short for that");
    std::vector<Entry> entries;
    entries.reserve((size_t)(mirror ? 2 * L : L));
    long long count = 0;
    while (next_data_line(p, end, le))
    {
        const long long i = std::strtoll(p, &q, 10);
        if (q == p)
            return fail("psb200_market_load: unreadable entry line " + std::to_string(count + 1));
        const char *q0 = q;
        const long long j = std::strtoll(q0, &q, 10);
        if (q == q0)
            return fail("psb200_market_load: unreadable entry line " + std::to_string(count + 1));
        double v = 1.0;
        if (!pattern)
        {
            const char *q1 = q;
            v = std::strtod(q1, &q);
            if (q == q1)
                return fail("psb200_market_load: missing value on entry line " + std::to_string(count + 1));
        }
        if (q > le)
            return fail("psb200_market_load: too few fields on entry line " + std::to_string(count + 1)); // strto* ran into the next line
        if (i < 1 || j < 1 || i > M || j > N)
            return fail("psb200_market_load: index out of range on entry line " + std::to_string(count + 1));
        entries.push_back({(int32_t)(i - 1), (int32_t)(j - 1), v});
        if (mirror && i != j)
        {
            if (j > M || i > N)
                return fail("psb200_market_load: a symmetric file needs a square matrix");
            entries.push_back({(int32_t)(j - 1), (int32_t)(i - 1), v});
        }
        ++count;
        p = le < end ? le + 1 : end;
    }
    if (count != L)
        return fail("psb200_market_load: the size line announces " + std::to_string(L) + " entries, the file holds " + std::to_string(count));
    if (entries.size() > 0x7ffffffeull)
        return fail("psb200_market_load: more than 2^31 entries (int32 index range, reference Types.hpp:11-15)");
    std::unique_ptr<psb200_market> m(new psb200_market());
    compress(M, N, entries, *m);
    const int64_t stored = (int64_t)m->inner.size();
    *out = m.release();
    if (rows)
        *rows = M;
    if (cols)
        *cols = N;
    if (nnz)
        *nnz = stored;
    return PSB200_OK;
}

int psb200_market_load(const char *path, int symmetric_mode, psb200_market_handle *out, int64_t *rows, int64_t *cols, int64_t *nnz)
{
    try
    {
        return market_load_impl(path, symmetric_mode, out, rows, cols, nnz);
    }
    catch (const std::exception &e) // e.g. bad_alloc for a size line that announces 2^31 columns
    {
        if (out)
            *out = nullptr;
        return fail(std::string("psb200_market_load: ") + e.what());
    }
}

int psb200_market_get_csc(psb200_market_handle m, int32_t *outer, int32_t *inner, double *vals)
{
    if (!m)
        return fail("psb200_market_get_csc: null handle");
    if (outer)
        std::copy(m->outer.begin(), m->outer.end(), outer);
    if (inner)
        std::copy(m->inner.begin(), m->inner.end(), inner);
    if (vals)
        std::copy(m->vals.begin(), m->vals.end(), vals);
    return PSB200_OK;
}

int psb200_market_free(psb200_market_handle m)
{
    delete m;
    return PSB200_OK;
}

int psb200_market_save(const char *path, int64_t rows, int64_t cols, const int32_t *outer, const int32_t *inner, const double *vals, int symmetric)
{
    if (!path || !outer || rows < 0 || cols < 0 || (outer[cols] > 0 && (!inner || !vals)))
        return fail("psb200_market_save: null or negative argument");
    FILE *f = std::fopen(path, "wb");
    if (!f)
        return fail(std::string("psb200_market_save: cannot open ") + path);
    long long nnz = 0;
    for (int64_t c = 0; c < cols; ++c)
        for (int32_t k = outer[c]; k < outer[c + 1]; ++k)
            nnz += (!symmetric || inner[k] >= c) ? 1 : 0;
    std::fprintf(f, "%%%%MatrixMarket matrix coordinate real %s\n%lld %lld %lld\n", symmetric ? "symmetric" : "general", (long long)rows,
                 (long long)cols, nnz);
    for (int64_t c = 0; c < cols; ++c)
        for (int32_t k = outer[c]; k < outer[c + 1]; ++k)
            if (!symmetric || inner[k] >= c)
                std::fprintf(f, "%d %lld %.17g\n", inner[k] + 1, (long long)c + 1, vals[k]);
    const bool ok = std::fclose(f) == 0;
    return ok ? PSB200_OK : fail("psb200_market_save: write failed");
}

int psb200_market_load_vector(const char *path, double *out, int64_t cap, int64_t *n)
{
    if (!path)
        return fail("psb200_market_load_vector: null path");
    std::string buf;
    if (!read_file(path, buf))
        return fail(std::string("psb200_market_load_vector: cannot open ") + path);
    const char *p = buf.data(), *end = buf.data() + buf.size(), *le = nullptr;
    if (!next_data_line(p, end, le))
        return fail("psb200_market_load_vector: missing size line");
    char *q = nullptr;
    const long long M = std::strtoll(p, &q, 10);
    const long long C = std::strtoll(q, &q, 10);
    if (M < 0 || C != 1)
        return fail("psb200_market_load_vector: expected an n x 1 array");
    if (M > (long long)(buf.size() / 2) + 1) // a value needs at least 2 bytes ("1\n"): do not let the caller allocate for a lie
        return fail("psb200_market_load_vector: the size line announces " + std::to_string(M) + " values, the file is too short for that");
    if (n)
        *n = M;
    if (!out)
        return PSB200_OK; // size query
    if (cap < M)
        return fail("psb200_market_load_vector: buffer too small");
    p = le < end ? le + 1 : end;
    long long k = 0;
    while (k < M && next_data_line(p, end, le))
    {
        const char *q0 = p;
        out[k] = std::strtod(q0, &q);
        if (q == q0)
            return fail("psb200_market_load_vector: unreadable value " + std::to_string(k + 1));
        ++k;
        p = le < end ? le + 1 : end;
    }
    return k == M ? PSB200_OK : fail("psb200_market_load_vector: file ends after " + std::to_string(k) + " values");
}

int psb200_market_save_vector(const char *path, const double *v, int64_t n)
{
    if (!path || n < 0 || (n > 0 && !v))
        return fail("psb200_market_save_vector: null or negative argument");
    FILE *f = std::fopen(path, "wb");
    if (!f)
        return fail(std::string("psb200_market_save_vector: cannot open ") + path);
    std::fprintf(f, "%%%%MatrixMarket matrix array real general\n%lld 1\n", (long long)n);
    for (int64_t i = 0; i < n; ++i)
        std::fprintf(f, "%.17g\n", v[i]);
    return std::fclose(f) == 0 ? PSB200_OK : fail("psb200_market_save_vector: write failed");
}

} // extern "C"
