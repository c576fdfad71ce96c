#include "seqio.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "spsp.h"

namespace spsp_host {

// ---------------------------------------------------------------- WordBuf

WordBuf::~WordBuf()
{
    if (!p_) return;
    if (pinned_) spsp_host_free(p_); else free(p_);
}

void WordBuf::reserve(uint64_t words)
{
    if (words <= cap_) return;
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
    out_.words.reserve(1024);
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

void FastaPacker::feed(const uint8_t *p, size_t n)
{
    if (!n) return;
    any_input_ = true;
    out_.words.reserve(word_idx_ + n / 16 + 16);
    uint32_t *w = out_.words.data();
    const uint8_t *end = p + n;
    uint32_t acc = acc_;
    int fill = fill_;
    uint64_t widx = word_idx_, nb = out_.n_bases;
    while (p < end) {
        if (state_ == SEQ) {
            // hot loop: bases of one line
            while (p < end) {
                uint8_t c = *p++;
                uint8_t code = kLut.v[c];
                if (code < 4) {
                    acc = (acc << 2) | code;
                    nb++;
                    if (++fill == 16) { w[widx++] = acc; acc = 0; fill = 0; }
                } else if (c == '\n') {
                    state_ = LINE_START;
                    break;
                }
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
