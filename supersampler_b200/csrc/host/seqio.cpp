#include "seqio.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__AVX2__)
#include <immintrin.h>
#endif

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "spsp.h"

namespace spsp_host {

// ---------------------------------------------------------------- WordBuf

WordBuf::~WordBuf() { detach(); }

void WordBuf::detach()
{
    if (p_ && !external_) {
        if (pinned_) spsp_host_free(p_); else free(p_);
    }
    p_ = nullptr; cap_ = 0; external_ = false;
}

void WordBuf::attach(uint32_t *p, uint64_t words)
{
    detach();
    p_ = p; cap_ = words; external_ = true;
}

void WordBuf::reserve(uint64_t words)
{
    if (words <= cap_) return;
    if (external_) throw std::runtime_error("packed region too small for this input");
    uint64_t ncap = cap_ ? cap_ : 1024;
    while (ncap < words) ncap += ncap / 2 + 1024;
    uint32_t *np = nullptr;
    if (pinned_) {
        void *v = nullptr;
        if (spsp_host_alloc(&v, ncap * sizeof(uint32_t)) != 0)
            throw std::runtime_error(std::string("pinned allocation failed: ") + spsp_last_error());
        np = static_cast<uint32_t *>(v);
    } else {
        np = static_cast<uint32_t *>(malloc(ncap * sizeof(uint32_t)));
        if (!np) throw std::bad_alloc();
    }
    if (p_) {
        memcpy(np, p_, cap_ * sizeof(uint32_t));
        if (pinned_) spsp_host_free(p_); else free(p_);
    }
    p_ = np;
    cap_ = ncap;
}

// ------------------------------------------------------------ FastaPacker

namespace {
struct Lut {
    uint8_t v[256];
    Lut()
    {
        memset(v, 4, sizeof v);
        // reference utils.cpp:13-16: code = (c/2)%4 on upper-cased ACGT
        v['A'] = v['a'] = 0; v['C'] = v['c'] = 1; v['T'] = v['t'] = 2; v['G'] = v['g'] = 3;
    }
};
const Lut kLut;
}  // namespace

FastaPacker::FastaPacker(PackedInput &out, uint32_t min_len) : out_(out), min_len_(min_len)
{
    out_.clear();
    out_.words.reserve(spsp_packed_words(0));
}

inline void FastaPacker::flush_word()
{
    out_.words.data()[word_idx_++] = acc_;
    acc_ = 0;
    fill_ = 0;
}

void FastaPacker::end_record()
{
    uint64_t len = out_.n_bases - rec_start_;
    if (len < min_len_) {
        // reference SubSampler.cpp:340-343: records shorter than k are skipped
        if (len || state_ != HEADER || any_input_) out_.dropped_records++;
        out_.n_bases = rec_start_;
        acc_ = ck_acc_; fill_ = ck_fill_; word_idx_ = ck_word_idx_;
    } else {
        out_.rec_off.push_back(out_.n_bases);
    }
    rec_start_ = out_.n_bases;
    ck_acc_ = acc_; ck_fill_ = fill_; ck_word_idx_ = word_idx_;
}

// Append 16 packed bases (first base in the MSBs of v) behind `pb` pending bits.
static inline void append16(uint32_t *w, uint64_t &widx, uint32_t &acc, int pb, uint32_t v)
{
    const uint64_t x = ((uint64_t)acc << 32) | v;
    w[widx++] = (uint32_t)(x >> pb);
    acc = v & ((1u << pb) - 1u);
}

// Append `nbits` (< 64, even) packed bases held in the low bits of `bits` behind the pending ones.
static inline void append_bits(uint32_t *w, uint64_t &widx, uint32_t &acc, int &fill, uint64_t bits, int nbits)
{
    unsigned __int128 x = ((unsigned __int128)acc << nbits) | bits;
    int total = 2 * fill + nbits;
    while (total >= 32) {
        w[widx++] = (uint32_t)(x >> (total - 32));
        total -= 32;
    }
    acc = (uint32_t)x & (uint32_t)((1ull << total) - 1);
    fill = total / 2;
}

// Sequence bytes of a FASTA record from p: valid bases are appended, '\n' ends the
// line (returns the position behind it, *eol = true), every other byte is deleted
// (clean_dna, utils.cpp:675-702).  32 bytes per step with AVX2: validity by a
// nibble LUT, codes (c>>1)&3 packed 4 per byte by two multiply-adds; a block that
// holds another byte contributes its leading valid bases and restarts behind it.
static inline const uint8_t *pack_seq(const uint8_t *p, const uint8_t *end, uint32_t *w, uint64_t &widx, uint32_t &acc,
                                      int &fill, uint64_t &nb, bool *eol)
{
    *eol = false;
#if defined(__AVX2__)
    const __m256i lut = _mm256_setr_epi8(-1, 'A', -1, 'C', 'T', -1, -1, 'G', -1, -1, -1, -1, -1, -1, -1, -1,
                                         -1, 'A', -1, 'C', 'T', -1, -1, 'G', -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i up = _mm256_set1_epi8((char)0xDF), three = _mm256_set1_epi8(3);
    const __m256i w41 = _mm256_set1_epi16(0x0104), w161 = _mm256_set1_epi32(0x00010010);
    const __m256i pick = _mm256_setr_epi8(12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                          12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    while (end - p >= 32) {
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p));
        const __m256i ok = _mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, c), _mm256_and_si256(c, up));
        const uint32_t okm = (uint32_t)_mm256_movemask_epi8(ok);
        const __m256i codes = _mm256_and_si256(_mm256_srli_epi16(c, 1), three);
        const __m256i b4 = _mm256_madd_epi16(_mm256_maddubs_epi16(codes, w41), w161);   // one byte per 4 bases
        const __m256i pk = _mm256_shuffle_epi8(b4, pick);
        const uint32_t v0 = (uint32_t)_mm256_extract_epi32(pk, 0), v1 = (uint32_t)_mm256_extract_epi32(pk, 4);
        if (okm == 0xFFFFFFFFu) {
            const int pb = 2 * fill;
            const uint64_t x0 = ((uint64_t)acc << 32) | v0;
            w[widx] = (uint32_t)(x0 >> pb);
            const uint64_t x1 = ((uint64_t)(v0 & ((1u << pb) - 1u)) << 32) | v1;
            w[widx + 1] = (uint32_t)(x1 >> pb);
            widx += 2;
            acc = v1 & ((1u << pb) - 1u);
            nb += 32;
            p += 32;
            continue;
        }
        const unsigned n = (unsigned)__builtin_ctz(~okm);            // leading valid bases, < 32
        if (n) {
            const uint64_t v = ((uint64_t)v0 << 32) | v1;
            append_bits(w, widx, acc, fill, v >> (64 - 2 * n), 2 * (int)n);
            nb += n;
        }
        p += n;
        if (*p == '\n') { *eol = true; return p + 1; }
        p++;                                                         // a deleted byte
    }
#endif
    for (; p < end; p++) {
        const uint8_t c = *p;
        const uint8_t code = kLut.v[c];
        if (code < 4) {
            acc = (acc << 2) | code;
            nb++;
            if (++fill == 16) { w[widx++] = acc; acc = 0; fill = 0; }
        } else if (c == '\n') {
            *eol = true;
            return p + 1;
        }
    }
    return end;
}

void FastaPacker::feed(const uint8_t *p, size_t n)
{
    if (!n) return;
    any_input_ = true;
    out_.words.reserve(word_idx_ + n / 16 + 2);       // every byte is at most one base
    uint32_t *w = out_.words.data();
    const uint8_t *end = p + n;
    uint32_t acc = acc_;
    int fill = fill_;
    uint64_t widx = word_idx_, nb = out_.n_bases;
    while (p < end) {
        if (state_ == SEQ) {
            bool eol;
            p = pack_seq(p, end, w, widx, acc, fill, nb, &eol);
            if (eol) {
                // next line of the same record unless it starts a new one
                if (p < end && *p != '>') continue;
                state_ = LINE_START;
            }
        } else if (state_ == HEADER) {
            const void *nl = memchr(p, '\n', (size_t)(end - p));
            if (!nl) { p = end; break; }
            p = static_cast<const uint8_t *>(nl) + 1;
            state_ = LINE_START;
        } else {  // LINE_START
            if (*p == '>') {
                acc_ = acc; fill_ = fill; word_idx_ = widx; out_.n_bases = nb;
                end_record();
                acc = acc_; fill = fill_; widx = word_idx_; nb = out_.n_bases;
                state_ = HEADER;
            } else {
                state_ = SEQ;
            }
        }
    }
    acc_ = acc; fill_ = fill; word_idx_ = widx; out_.n_bases = nb;
}

void FastaPacker::finish()
{
    end_record();
    // left-align the last partial word, then zero padding for the kernels
    uint64_t need = spsp_packed_words(out_.n_bases);
    out_.words.reserve(need);
    uint32_t *w = out_.words.data();
    uint64_t widx = word_idx_;
    if (fill_) w[widx++] = acc_ << (2 * (16 - fill_));
    for (; widx < need; widx++) w[widx] = 0;
}

// ------------------------------------------------------------------ files

static bool is_gzip(int fd)
{
    unsigned char mg[2] = {0, 0};
    ssize_t r = pread(fd, mg, 2, 0);
    return r == 2 && mg[0] == 0x1f && mg[1] == 0x8b;
}

bool pack_fasta_file(const std::string &path, uint32_t min_len, PackedInput &out, uint64_t *file_bytes)
{
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || S_ISDIR(st.st_mode)) { close(fd); return false; }
    if (file_bytes) *file_bytes = (uint64_t)st.st_size;
    FastaPacker pk(out, min_len);
    const size_t CH = 4u << 20;
    std::vector<uint8_t> buf(CH);
    if (is_gzip(fd)) {
        gzFile gz = gzdopen(fd, "rb");
        if (!gz) { close(fd); return false; }
        gzbuffer(gz, 1u << 20);
        for (;;) {
            int r = gzread(gz, buf.data(), (unsigned)CH);
            if (r <= 0) break;
            pk.feed(buf.data(), (size_t)r);
        }
        gzclose(gz);
    } else {
        out.words.reserve((uint64_t)st.st_size / 16 + 64);
        for (;;) {
            ssize_t r = read(fd, buf.data(), CH);
            if (r <= 0) break;
            pk.feed(buf.data(), (size_t)r);
        }
        close(fd);
    }
    pk.finish();
    return true;
}

void pack_fasta_buffer(const uint8_t *p, size_t n, uint32_t min_len, PackedInput &out)
{
    FastaPacker pk(out, min_len);
    pk.feed(p, n);
    pk.finish();
}

bool read_file_maybe_gz(const std::string &path, std::vector<uint8_t> &out)
{
    out.clear();
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || S_ISDIR(st.st_mode)) { close(fd); return false; }
    if (is_gzip(fd)) {
        gzFile gz = gzdopen(fd, "rb");
        if (!gz) { close(fd); return false; }
        gzbuffer(gz, 1u << 18);
        size_t cap = (size_t)st.st_size * 4 + 4096;
        out.resize(cap);
        size_t n = 0;
        for (;;) {
            if (n == out.size()) out.resize(out.size() * 2);
            int r = gzread(gz, out.data() + n, (unsigned)std::min<size_t>(out.size() - n, 1u << 30));
            if (r <= 0) break;
            n += (size_t)r;
        }
        gzclose(gz);
        out.resize(n);
    } else {
        out.resize((size_t)st.st_size);
        size_t n = 0;
        while (n < out.size()) {
            ssize_t r = read(fd, out.data() + n, out.size() - n);
            if (r <= 0) break;
            n += (size_t)r;
        }
        out.resize(n);
        close(fd);
    }
    return true;
}

bool write_gz(const std::string &path, const uint8_t *p, size_t n, int level)
{
    char mode[8];
    snprintf(mode, sizeof mode, "wb%d", level);
    gzFile gz = gzopen(path.c_str(), mode);
    if (!gz) return false;
    size_t off = 0;
    while (off < n) {
        unsigned ch = (unsigned)std::min<size_t>(n - off, 1u << 30);
        if (gzwrite(gz, p + off, ch) <= 0) { gzclose(gz); return false; }
        off += ch;
    }
    return gzclose(gz) == Z_OK;
}

}  // namespace spsp_host
