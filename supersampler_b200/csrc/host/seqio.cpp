#include "seqio.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
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

// While a buffer is being filled its words hold base j of the word at bits
// 2j..2j+1 ("low first", what PEXT compaction produces); finish() turns every word
// into the output order (first base in the top bits) in one vectorised sweep.
static inline void emit64(uint32_t *w, uint64_t &widx, uint64_t acc)
{
    memcpy(w + widx, &acc, 8);
    widx += 2;
}
// dst[i] = src[i] with its 16 two-bit groups reversed, i in [0, n).  The destination is the (pinned) staging
// buffer, written exactly once: non-temporal stores keep it from being read first (no read-for-ownership) and
// from displacing the text that is being streamed through the caches.
static inline uint32_t word_to_msb_first(uint32_t x)
{
    x = __builtin_bswap32(x);
    x = ((x & 0x0F0F0F0Fu) << 4) | ((x >> 4) & 0x0F0F0F0Fu);
    return ((x & 0x33333333u) << 2) | ((x >> 2) & 0x33333333u);
}
static void words_to_msb_first(const uint32_t *src, uint32_t *dst, uint64_t n)
{
    uint64_t i = 0;
#if defined(__AVX2__)
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31u)) { dst[i] = word_to_msb_first(src[i]); i++; }
    const __m256i rev4 = _mm256_setr_epi8(0, 4, 8, 12, 1, 5, 9, 13, 2, 6, 10, 14, 3, 7, 11, 15,
                                          0, 4, 8, 12, 1, 5, 9, 13, 2, 6, 10, 14, 3, 7, 11, 15);   // nibble ab -> ba (2-bit groups)
    const __m256i bswap = _mm256_setr_epi8(3, 2, 1, 0, 7, 6, 5, 4, 11, 10, 9, 8, 15, 14, 13, 12,
                                           3, 2, 1, 0, 7, 6, 5, 4, 11, 10, 9, 8, 15, 14, 13, 12);
    const __m256i lo4 = _mm256_set1_epi8(0x0F);
    static const bool nt = getenv("SPSP_PACK_NO_NT") == nullptr;          // experiments: plain stores
    const bool streamed = nt && i + 8 <= n;
    for (; i + 8 <= n; i += 8) {
        __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        const __m256i l = _mm256_shuffle_epi8(rev4, _mm256_and_si256(x, lo4));
        const __m256i h = _mm256_shuffle_epi8(rev4, _mm256_and_si256(_mm256_srli_epi16(x, 4), lo4));
        x = _mm256_or_si256(_mm256_slli_epi16(l, 4), h);                   // groups reversed inside every byte
        x = _mm256_shuffle_epi8(x, bswap);
        if (nt) _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), x);
        else _mm256_store_si256(reinterpret_cast<__m256i *>(dst + i), x);
    }
    if (streamed) _mm_sfence();                                            // before the words are handed to a copy
#endif
    for (; i < n; i++) dst[i] = word_to_msb_first(src[i]);
}

// The block loops run ~50 uops per 32 bytes, so the out-of-order window only covers a few cache lines of
// text: an explicit prefetch this far ahead keeps enough misses in flight (measured on the bench box: the
// hardware streamer alone leaves the loop latency-bound at ~6 GB/s per thread).
// With the output leaving through non-temporal stores the best distance moved from 2 KB to 4 KB (e2e on one box:
// 1 KB 75, 2 KB 84-85, 4 KB 96-97, 6-16 KB 94 Gbp/s; prefetching into L2 only or non-temporally was slower).
// SPSP_PACK_PF_DIST overrides it for tuning on another host.
constexpr int PACK_PREFETCH = 4096;
static const int g_pf_dist = getenv("SPSP_PACK_PF_DIST") ? atoi(getenv("SPSP_PACK_PF_DIST")) : PACK_PREFETCH;
static inline void pack_prefetch(const void *p)
{
    _mm_prefetch(static_cast<const char *>(p) + g_pf_dist, _MM_HINT_T0);
}

// Whole 32-byte blocks of sequence text that hold no '>' (a possible record
// start, left to the byte-wise state machine): every byte that is not a base --
// line ends, N, IUPAC codes, CR -- is deleted (clean_dna, utils.cpp:675-702)
// without a branch: validity by a nibble LUT, codes (c>>1)&3 gathered 4 per byte
// by two multiply-adds, PEXT squeezes the deleted positions out of the 64-bit
// code word.  Returns the first byte not consumed.
static inline const uint8_t *pack_blocks(const uint8_t *p, const uint8_t *end, uint32_t *w, uint64_t &widx, uint64_t &acc,
                                         int &fill, uint64_t &nb)
{
#if defined(__AVX2__) && defined(__BMI2__)
    const __m256i lut = _mm256_setr_epi8(-1, 'A', -1, 'C', 'T', -1, -1, 'G', -1, -1, -1, -1, -1, -1, -1, -1,
                                         -1, 'A', -1, 'C', 'T', -1, -1, 'G', -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i up = _mm256_set1_epi8((char)0xDF), three = _mm256_set1_epi8(3), gt = _mm256_set1_epi8('>');
    const __m256i ones = _mm256_set1_epi8((char)0xFF);
    const __m256i w14 = _mm256_set1_epi16(0x0401), w116 = _mm256_set1_epi32(0x00100001);
    const __m256i pick = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                          0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    while (end - p >= 32) {
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p));
        pack_prefetch(p);
        const uint32_t okm = (uint32_t)_mm256_movemask_epi8(
            _mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, c), _mm256_and_si256(c, up)));
        // '>' (and 0xFF, see LINE_START below) is not a base: only a block with a deleted byte can hold one
        if (okm != 0xFFFFFFFFu && _mm256_movemask_epi8(_mm256_or_si256(_mm256_cmpeq_epi8(c, gt), _mm256_cmpeq_epi8(c, ones)))) break;
        const __m256i codes = _mm256_and_si256(_mm256_srli_epi16(c, 1), three);
        const __m256i b4 = _mm256_madd_epi16(_mm256_maddubs_epi16(codes, w14), w116);   // byte = c0 + 4 c1 + 16 c2 + 64 c3
        const __m256i pk = _mm256_shuffle_epi8(b4, pick);
        const uint64_t all = (uint64_t)(uint32_t)_mm256_extract_epi32(pk, 0) |
                             ((uint64_t)(uint32_t)_mm256_extract_epi32(pk, 4) << 32);      // code of byte j at bits 2j
        const uint64_t keep = _pdep_u64(okm, 0x5555555555555555ULL) * 3;                   // 2 mask bits per valid byte
        const uint64_t v = _pext_u64(all, keep);
        const int nbits = 2 * __builtin_popcount(okm);
        acc |= v << fill;                                                                  // fill < 64
        if (fill + nbits >= 64) {
            emit64(w, widx, acc);
            acc = fill ? v >> (64 - fill) : 0;
            fill -= 64;
        }
        fill += nbits;
        nb += (uint64_t)(nbits >> 1);
        p += 32;
    }
#endif
    (void)end; (void)w; (void)widx; (void)acc; (void)fill; (void)nb;
    return p;
}

#if defined(__x86_64__)
// The same block loop with AVX-512 (VBMI2), chosen at run time: 64 bytes per step, validity and '>' as mask
// registers, VPCOMPRESSB squeezes the deleted bytes out before the codes are formed (no PDEP/PEXT chain), so a
// step costs about what the 32-byte AVX2 step costs.
__attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi,avx512vbmi2,bmi,bmi2,lzcnt,popcnt")))
static const uint8_t *pack_blocks_avx512(const uint8_t *p, const uint8_t *end, uint32_t *w, uint64_t &widx, uint64_t &acc,
                                         int &fill, uint64_t &nb)
{
    const __m512i lut = _mm512_broadcast_i32x4(_mm_setr_epi8(-1, 'A', -1, 'C', 'T', -1, -1, 'G', -1, -1, -1, -1, -1, -1, -1, -1));
    const __m512i up = _mm512_set1_epi8((char)0xDF), three = _mm512_set1_epi8(3), gt = _mm512_set1_epi8('>');
    const __m512i w14 = _mm512_set1_epi16(0x0401), w116 = _mm512_set1_epi32(0x00100001);
    while (end - p >= 64) {
        const __m512i c = _mm512_loadu_si512(p);
        pack_prefetch(p);
        const uint64_t ok = _mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(lut, c), _mm512_and_si512(c, up));
        // '>' (and 0xFF) is not a base: only a block with a deleted byte can hold one
        if (ok != ~0ULL && (_mm512_cmpeq_epi8_mask(c, gt) | _mm512_cmpeq_epi8_mask(c, _mm512_set1_epi8((char)0xFF)))) break;
        const __m512i z = _mm512_maskz_compress_epi8(ok, c);                                // bases first, zeros behind
        const __m512i codes = _mm512_and_si512(_mm512_srli_epi16(z, 1), three);
        const __m512i b4 = _mm512_madd_epi16(_mm512_maddubs_epi16(codes, w14), w116);       // byte = c0 + 4 c1 + 16 c2 + 64 c3
        const __m128i pk = _mm512_cvtepi32_epi8(b4);                                        // code of base j at bits 2j of 128
        const uint64_t v0 = (uint64_t)_mm_cvtsi128_si64(pk), v1 = (uint64_t)_mm_extract_epi64(pk, 1);
        const int n = __builtin_popcountll(ok);
        const int n0 = n < 32 ? 2 * n : 64, n1 = 2 * n - n0;                                // bits taken from v0 / v1
        acc |= v0 << fill;
        if (fill + n0 >= 64) {
            memcpy(w + widx, &acc, 8); widx += 2;
            acc = fill ? v0 >> (64 - fill) : 0;
            fill -= 64;
        }
        fill += n0;
        acc |= v1 << fill;                                                                  // v1 == 0 when n1 == 0
        if (fill + n1 >= 64) {
            memcpy(w + widx, &acc, 8); widx += 2;
            acc = fill ? v1 >> (64 - fill) : 0;
            fill -= 64;
        }
        fill += n1;
        nb += (uint64_t)n;
        p += 64;
    }
    return p;
}

static bool have_avx512_packer()
{
    static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                           __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512vbmi") &&
                           __builtin_cpu_supports("avx512vbmi2") && !getenv("SPSP_NO_AVX512");
    return ok;
}
#else
static bool have_avx512_packer() { return false; }
#endif

// Words that are not final yet live in scratch_ (scratch_[0] is output word converted_), in the packer's
// "low first" bit order; commit() / finish() move them to the output buffer in output order.  Long inputs are
// fed in pieces so that the scratch stays cache-resident between the two passes.
void FastaPacker::feed(const uint8_t *p, size_t n)
{
    constexpr size_t PIECE = (size_t)512 << 10;
    while (n > PIECE) {
        feed_piece(p, PIECE);
        commit();
        p += PIECE; n -= PIECE;
    }
    feed_piece(p, n);
}

void FastaPacker::feed_piece(const uint8_t *p, size_t n)
{
    if (!n) return;
    any_input_ = true;
    out_.words.reserve(word_idx_ + n / 16 + 4);       // every byte is at most one base
    if (scratch_.size() < word_idx_ - converted_ + n / 16 + 8) scratch_.resize(word_idx_ - converted_ + n / 16 + 8);
    // w + widx addresses scratch word widx - converted_ (converted_ does not move inside this call)
    uint32_t *w = reinterpret_cast<uint32_t *>(reinterpret_cast<uintptr_t>(scratch_.data()) - converted_ * sizeof(uint32_t));
    const uint8_t *const begin = p, *end = p + n;
    uint64_t acc = acc_;
    int fill = fill_;
    uint64_t widx = word_idx_, nb = out_.n_bases;
    const bool wide = have_avx512_packer();
    (void)wide;
    while (p < end) {
        if (state_ == SEQ) {
            const uint8_t *q = p;
#if defined(__x86_64__)
            if (wide) q = pack_blocks_avx512(q, end, w, widx, acc, fill, nb);
#endif
            q = pack_blocks(q, end, w, widx, acc, fill, nb);
            if (q != p) {
                p = q;
                if (p[-1] == '\n') state_ = LINE_START;          // the blocks ended with a line
                continue;
            }
            // byte-wise to the end of this line (block with a '>' in it, or the tail of the buffer)
            while (p < end) {
                const uint8_t c = *p++;
                const uint8_t code = kLut.v[c];
                if (code < 4) {
                    acc |= (uint64_t)code << fill;
                    fill += 2;
                    nb++;
                    if (fill == 64) { emit64(w, widx, acc); acc = 0; fill = 0; }
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
            // a line that starts with '>' -- or with the byte 0xFF, which the reference's `char c = peek(); c != EOF`
            // (utils.cpp:709-713) cannot tell from the end of the file -- is a header line and starts a record
            if (*p == '>' || *p == 0xFF) {
                acc_ = acc; fill_ = fill; word_idx_ = widx; out_.n_bases = nb;
                end_record();
                acc = acc_; fill = fill_; widx = word_idx_; nb = out_.n_bases;
                state_ = HEADER;
            } else {
                state_ = SEQ;
            }
        }
    }
    (void)begin;
    acc_ = acc; fill_ = fill; word_idx_ = widx; out_.n_bases = nb;
}

uint64_t FastaPacker::commit()
{
    // a record shorter than min_len is taken back when it ends (end_record): everything older than the
    // current record's start is final, and so is everything once the record is long enough to stay
    uint64_t stable = (out_.n_bases - rec_start_ >= min_len_) ? word_idx_ : ck_word_idx_;
    stable &= ~(uint64_t)1;                                     // whole 64-bit units
    if (stable > converted_) {
        const uint64_t n = stable - converted_, rest = word_idx_ - stable;
        words_to_msb_first(scratch_.data(), out_.words.data() + converted_, n);
        if (rest) memmove(scratch_.data(), scratch_.data() + n, rest * sizeof(uint32_t));
        converted_ = stable;
    }
    return converted_;
}

void FastaPacker::finish()
{
    end_record();
    uint64_t need = spsp_packed_words(out_.n_bases);
    out_.words.reserve(std::max<uint64_t>(need, word_idx_ + 2));
    if (scratch_.size() < word_idx_ - converted_ + 2) scratch_.resize(word_idx_ - converted_ + 2);
    uint32_t *w = reinterpret_cast<uint32_t *>(reinterpret_cast<uintptr_t>(scratch_.data()) - converted_ * sizeof(uint32_t));
    uint64_t widx = word_idx_;
    if (fill_) emit64(w, widx, acc_);                 // bits above fill_ are zero: left-aligned after the sweep
    if (widx > converted_) words_to_msb_first(scratch_.data(), out_.words.data() + converted_, widx - converted_);
    converted_ = widx;
    w = out_.words.data();
    widx = (out_.n_bases + 15) / 16;                  // zero padding for the kernels
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
        bool bad = false;
        for (;;) {
            int r = gzread(gz, buf.data(), (unsigned)CH);
            if (r < 0) bad = true;                 // truncated / corrupt stream: the reference's zstr throws here
            if (r <= 0) break;
            pk.feed(buf.data(), (size_t)r);
        }
        if (gzclose(gz) != Z_OK) bad = true;
        if (bad) return false;
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
            if (r < 0) { gzclose(gz); out.clear(); return false; }     // truncated / corrupt: not a partial success
            if (r == 0) break;
            n += (size_t)r;
        }
        if (gzclose(gz) != Z_OK) { out.clear(); return false; }
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
