#include "sketchfile.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include <zlib.h>

namespace spsp_host {

typedef unsigned __int128 u128;

namespace {

inline uint64_t rc_bits64(uint64_t x)
{
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    return __builtin_bswap64(x) ^ 0xAAAAAAAAAAAAAAAAULL;
}
inline uint64_t canon(uint64_t x, int n)
{
    uint64_t r = rc_bits64(x) >> (64 - 2 * n);
    return x < r ? x : r;
}
inline u128 canon(u128 x, int n)
{
    u128 r = (((u128)rc_bits64((uint64_t)x) << 64) | rc_bits64((uint64_t)(x >> 64))) >> (128 - 2 * n);
    return x < r ? x : r;
}

inline void push_key(SketchElems &o, uint64_t v) { o.klo.push_back(v); }
inline void push_key(SketchElems &o, u128 v) { o.klo.push_back((uint64_t)v); o.khi.push_back((uint64_t)(v >> 64)); }

// All k-mers of a run of 2-bit codes (canonical) appended to keys.
template <class Key>
void kmers_of(const std::vector<uint8_t> &codes, int k, Key kmask, std::vector<Key> &keys)
{
    if ((int)codes.size() < k) return;
    Key cur = 0;
    for (int i = 0; i < k - 1; i++) cur = (cur << 2) | codes[i];
    for (size_t i = (size_t)k - 1; i < codes.size(); i++) {
        cur = ((cur << 2) | codes[i]) & kmask;
        keys.push_back(canon(cur, k));
    }
}

template <class Key>
void decode_body(const uint8_t *p, size_t n, size_t pos, int k, int m, SketchElems &out)
{
    const int d = k - m;
    const Key kmask = (((Key)1) << (2 * k)) - 1;
    std::vector<uint8_t> sk, mcodes((size_t)m);
    std::vector<Key> keys;
    while (pos + (size_t)m <= n) {
        uint32_t minimizer = 0;
        for (int i = 0; i < m; i++) {
            mcodes[i] = (p[pos + i] >> 1) & 3;           // str2num, utils.cpp:158-165
            minimizer = (minimizer << 2) | mcodes[i];
        }
        pos += (size_t)m;
        uint32_t sz = 0;
        if (pos + 4 <= n) { memcpy(&sz, p + pos, 4); pos += 4; } else { pos = n; }
        if ((size_t)sz > n - pos) sz = (uint32_t)(n - pos);
        keys.clear();
        // maximal super-k-mers: [mod byte][4 bases / byte]; every 2d bases are
        // prefix(d) + suffix(d) around the minimizer (strDecompressor utils.cpp:71-111,
        // inject_minimizer Comparator.cpp:78-92)
        if (sz > 1) {
            const uint8_t *q = p + pos;
            unsigned mod = q[0];
            size_t nbases = (mod == 0) ? ((size_t)sz - 1) * 4 : ((size_t)sz - 2) * 4 + mod;
            const size_t full_bases = (mod == 0) ? nbases : ((size_t)sz - 2) * 4;
            auto code_at = [&](size_t i) -> uint8_t {
                if (i < full_bases) return (q[1 + (i >> 2)] >> (6 - 2 * (i & 3))) & 3;
                return (q[sz - 1] >> (2 * (mod - (i - full_bases)))) & 3;   // trailing partial byte
            };
            for (size_t b = 0; b + 2 * (size_t)d <= nbases; b += 2 * (size_t)d) {
                sk.clear();
                for (int i = 0; i < d; i++) sk.push_back(code_at(b + i));
                sk.insert(sk.end(), mcodes.begin(), mcodes.end());
                for (int i = 0; i < d; i++) sk.push_back(code_at(b + d + i));
                kmers_of<Key>(sk, k, kmask, keys);
            }
        }
        pos += sz;
        // non-maximal super-k-mers: "prefix\nsuffix\n" pairs until two empty lines
        for (;;) {
            size_t a0 = pos; while (pos < n && p[pos] != '\n') pos++;
            size_t a1 = pos; if (pos < n) pos++;
            size_t b0 = pos; while (pos < n && p[pos] != '\n') pos++;
            size_t b1 = pos; if (pos < n) pos++;
            if (a1 == a0 && b1 == b0) break;
            sk.clear();
            for (size_t i = a0; i < a1; i++) sk.push_back((p[i] >> 1) & 3);
            sk.insert(sk.end(), mcodes.begin(), mcodes.end());
            for (size_t i = b0; i < b1; i++) sk.push_back((p[i] >> 1) & 3);
            kmers_of<Key>(sk, k, kmask, keys);
            if (pos >= n) break;
        }
        std::sort(keys.begin(), keys.end());
        keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
        for (const Key &x : keys) { out.minim.push_back(minimizer); push_key(out, x); }
    }
}

}  // namespace

bool decode_sketch(const uint8_t *p, size_t n, SketchElems &out, std::string *err)
{
    out = SketchElems();
    size_t pos = 0;
    while (pos < n && p[pos] != '\n') pos++;
    if (pos >= n || pos > 120) { if (err) *err = "no header line"; return false; }
    char hdr[128];
    memcpy(hdr, p, pos);
    hdr[pos] = 0;
    int L = 0, m = 0;
    if (sscanf(hdr, "%d %d", &L, &m) != 2 || m < 1 || m > 15 || L <= m) {
        if (err) *err = "bad header";
        return false;
    }
    int k = (L + m) / 2;                 // Comparator.cpp:34
    if (k > 63 || k <= m) { if (err) *err = "unsupported k"; return false; }
    out.k = k; out.m = m;
    pos++;
    if (k <= 32) decode_body<uint64_t>(p, n, pos, k, m, out);
    else decode_body<u128>(p, n, pos, k, m, out);
    return true;
}

static void csv_header(const std::vector<std::string> &names, bool jaccard, std::vector<uint8_t> &out)
{
    const uint32_t n = (uint32_t)names.size();
    for (uint32_t i = 0; i < n; i++) {
        out.insert(out.end(), names[i].begin(), names[i].end());
        out.push_back(i + 1 == n ? '\n' : ',');
    }
    if (!jaccard) out.push_back('\n');                    // Comparator.cpp:373
}

// rows [i0, i1) of the matrix as text (Comparator.cpp:375-407, :425-459)
static void csv_rows(uint32_t n, uint32_t i0, uint32_t i1, const uint32_t *inter, uint64_t ld, bool row_major_full,
                     const std::vector<uint64_t> &sizes, bool jaccard, unsigned precision, double min_threshold,
                     std::vector<uint8_t> &out)
{
    char tmp[64];
    for (uint32_t i = i0; i < i1; i++)
        for (uint32_t j = 0; j < n; j++) {
            if (i == j) out.push_back('1');
            else {
                // all-vs-all results hold the pair at (min,max); query results hold row i fully
                uint32_t c = row_major_full ? inter[(uint64_t)i * ld + j]
                                            : (i < j ? inter[(uint64_t)i * ld + j] : inter[(uint64_t)j * ld + i]);
                if (!c) out.push_back('0');
                else {
                    double sc = jaccard ? (double)c / (double)(sizes[i] + sizes[j] - c) : (double)c / (double)sizes[i];
                    if (sc < min_threshold) out.push_back('0');
                    else {
                        const int l = snprintf(tmp, sizeof tmp, "%.*g", (int)precision, sc);
                        out.insert(out.end(), tmp, tmp + l);
                    }
                }
            }
            out.push_back(j + 1 == n ? '\n' : ',');
        }
}

void format_csv(const std::vector<std::string> &names, uint32_t query_size, const uint32_t *inter, uint64_t ld,
                bool row_major_full, const std::vector<uint64_t> &sizes, bool jaccard, unsigned precision,
                double min_threshold, std::vector<uint8_t> &out)
{
    const uint32_t n = (uint32_t)names.size();
    csv_header(names, jaccard, out);
    csv_rows(n, 0, std::min(n, query_size), inter, ld, row_major_full, sizes, jaccard, precision, min_threshold, out);
}

// One complete gzip member holding p[0, n).
static bool gzip_member(const uint8_t *p, size_t n, int level, std::vector<uint8_t> &out)
{
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    out.resize(deflateBound(&zs, (uLong)n) + 64);
    size_t off = 0, produced = 0;
    int rc = Z_OK;
    do {                                                   // (avail_in is 32 bits wide)
        const size_t piece = std::min<size_t>(n - off, (size_t)1 << 30);
        zs.next_in = const_cast<Bytef *>(p + off);
        zs.avail_in = (uInt)piece;
        off += piece;
        const int flush = off == n ? Z_FINISH : Z_NO_FLUSH;
        do {
            zs.next_out = out.data() + produced;
            zs.avail_out = (uInt)std::min<size_t>(out.size() - produced, (size_t)1 << 30);
            const size_t before = zs.avail_out;
            rc = deflate(&zs, flush);
            produced += before - zs.avail_out;
            if (rc == Z_STREAM_ERROR) { deflateEnd(&zs); return false; }
            if (produced == out.size()) out.resize(out.size() * 2);
        } while (zs.avail_in || (flush == Z_FINISH && rc != Z_STREAM_END));
    } while (off < n);
    deflateEnd(&zs);
    out.resize(produced);
    return rc == Z_STREAM_END;
}

bool write_csv_gz(const std::string &path, const std::vector<std::string> &names, uint32_t query_size, const uint32_t *inter,
                  uint64_t ld, bool row_major_full, const std::vector<uint64_t> &sizes, bool jaccard, unsigned precision,
                  double min_threshold, int threads, uint64_t *text_bytes)
{
    const uint32_t n = (uint32_t)names.size(), rows = std::min(n, query_size);
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    uint64_t total = 0;
    bool ok = true;
    auto put_member = [&](const std::vector<uint8_t> &text) {
        std::vector<uint8_t> z;
        if (!gzip_member(text.data(), text.size(), 1, z) || fwrite(z.data(), 1, z.size(), f) != z.size()) ok = false;
        total += text.size();
    };
    {
        std::vector<uint8_t> hdr;
        csv_header(names, jaccard, hdr);
        put_member(hdr);                                   // (an empty matrix still yields a valid gzip file)
    }
    // blocks of about 4 MB of text, `threads` of them formatted and compressed at a time, written in order: the
    // file is a sequence of gzip members (what `zcat`, zlib's gzread and the reference's zstr all read as one
    // stream) and host memory holds one wave of blocks, never the whole matrix as text
    const unsigned nw = (unsigned)std::max(1, std::min(threads, 64));
    const uint32_t per_block = (uint32_t)std::max<uint64_t>(1, ((uint64_t)4 << 20) / ((uint64_t)n * 6 + 1));
    for (uint32_t i0 = 0; i0 < rows && ok; i0 += per_block * nw) {
        const uint32_t blocks = std::min<uint32_t>(nw, (rows - i0 + per_block - 1) / per_block);
        std::vector<std::vector<uint8_t>> z(blocks);
        std::vector<uint64_t> tb(blocks, 0);
        std::vector<char> good(blocks, 1);
        std::vector<std::thread> pool;
        auto work = [&](uint32_t b) {
            std::vector<uint8_t> text;
            const uint32_t a = i0 + b * per_block, e = std::min(rows, a + per_block);
            text.reserve((size_t)(e - a) * n * 4);
            csv_rows(n, a, e, inter, ld, row_major_full, sizes, jaccard, precision, min_threshold, text);
            tb[b] = text.size();
            if (!gzip_member(text.data(), text.size(), 1, z[b])) good[b] = 0;
        };
        for (uint32_t b = 1; b < blocks; b++) pool.emplace_back(work, b);
        work(0);
        for (auto &t : pool) t.join();
        for (uint32_t b = 0; b < blocks && ok; b++) {
            if (!good[b] || fwrite(z[b].data(), 1, z[b].size(), f) != z[b].size()) ok = false;
            total += tb[b];
        }
    }
    if (fclose(f) != 0) ok = false;
    if (text_bytes) *text_bytes = total;
    return ok;
}

}  // namespace spsp_host
