#include "postpass.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "seqio.h"

namespace spsp_host {

typedef unsigned __int128 u128;

uint64_t compute_threshold(int k, int m, double s)
{
    if (!(s > 1)) return ~(uint64_t)0;
    uint64_t w = (uint64_t)(k - m + 1);
    long double frac = (long double)1 / s;
    long double root = powl((long double)1 - frac, (long double)1 / w);
    long double res = ((long double)1 - root) * ((uint64_t)1 << 63);
    return (uint64_t)res * 2;
}

namespace {

inline uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
// XXH64 of an 8-byte value, seed 1312 (include/xxhash64.h:115-148,158-163).
inline uint64_t mmer_hash(uint64_t x)
{
    const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL, P3 = 1609587929392839161ULL,
                   P4 = 9650029242287828579ULL, P5 = 2870177450012600261ULL;
    uint64_t r = 1312ULL + P5 + 8ULL;
    r ^= rotl(x * P2, 31) * P1;
    r = rotl(r, 27) * P1 + P4;
    r ^= r >> 33; r *= P2; r ^= r >> 29; r *= P3; r ^= r >> 32;
    return r;
}

inline uint64_t rc_bits64(uint64_t x)
{
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    return __builtin_bswap64(x) ^ 0xAAAAAAAAAAAAAAAAULL;
}
inline uint64_t revcomp(uint64_t x, int n) { return rc_bits64(x) >> (64 - 2 * n); }
inline u128 revcomp(u128 x, int n)
{
    u128 r = ((u128)rc_bits64((uint64_t)x) << 64) | rc_bits64((uint64_t)(x >> 64));
    return r >> (128 - 2 * n);
}

const char kBase[4] = {'A', 'C', 'T', 'G'};

struct RHit {               // one hit inside the current record
    uint64_t pos;           // relative to the record
    uint64_t hash;
    uint32_t canon;
    bool rev;
};

struct Piece {
    uint64_t first, last;   // k-mer indices (inclusive) relative to the record
    uint32_t minimizer;
    bool rev;
};

// State of the tracked minimizer (hits only; `valid` false = a non-selected m-mer).
struct Track {
    bool valid = false;
    uint32_t canon = 0;
    uint64_t hash = 0, posmin = 0;
    bool rev = false;
};

// regular_minimizer_pos restricted to hits (SubSampler.cpp:81-169): scan the
// window of k-mer c right to left; keeps the reference's position quirks.
Track rescan(const std::vector<RHit> &h, size_t lo, size_t hi, uint64_t c, uint64_t d)
{
    Track t;
    uint64_t position = 0;
    for (size_t i = hi; i-- > lo;) {
        const RHit &x = h[i];
        uint64_t j = c + d - x.pos;
        if (j == 0) {
            t.valid = true; t.canon = x.canon; t.hash = x.hash; t.rev = x.rev;
            position = x.rev ? 0 : d;                                        // :88-93
        } else if (!t.valid || t.hash > x.hash) {
            t.valid = true; t.canon = x.canon; t.hash = x.hash; t.rev = x.rev;
            position = d - j;                                                // :117-128
        } else if (x.canon == t.canon && x.rev == t.rev) {                   // :149-164
            if (t.rev && position > j) position = j;
            if (!t.rev && position > d - j) position = d - j;
        }
    }
    t.posmin = c + position;
    return t;
}

// Sparse replay of SubSampler.cpp:352-454 for one record of n bases.
void replay_record(const std::vector<RHit> &h, uint64_t n, int k, int m, std::vector<Piece> &pieces)
{
    if (h.empty()) return;
    const uint64_t d = (uint64_t)(k - m), K = n - k + 1;
    size_t lo = 0, hi = 0;                       // hits with pos in [c, c+d] are h[lo, hi)
    auto window = [&](uint64_t c) {
        while (hi < h.size() && h[hi].pos <= c + d) hi++;
        while (lo < hi && h[lo].pos < c) lo++;
    };
    window(0);
    Track cur = rescan(h, lo, hi, 0, d);                                     // :359-365
    bool old_valid = cur.valid, old_rev = cur.rev, is_rev = cur.rev;
    uint32_t old_min = cur.canon;
    uint64_t last = 0, c = 1;
    while (c < K) {
        if (!cur.valid) {
            // nothing observable happens until the next hit enters on the right
            if (hi >= h.size()) break;
            uint64_t nc = h[hi].pos - d;        // pos > c-1+d here, so nc >= c
            if (nc > c) c = nc;
            if (c >= K) break;
        }
        window(c);
        const uint64_t p = c + d;
        bool dump = false;
        const RHit *ent = (hi > lo && h[hi - 1].pos == p) ? &h[hi - 1] : nullptr;
        if (ent && (!cur.valid || ent->hash < cur.hash)) {                   // :374-388
            cur.valid = true; cur.canon = ent->canon; cur.hash = ent->hash; cur.posmin = p;
            cur.rev = ent->rev; is_rev = ent->rev;
        } else if (cur.valid && c - 1 >= cur.posmin) {                       // :391-398
            cur = rescan(h, lo, hi, c, d);
            if (cur.valid) is_rev = cur.rev;
            dump = true;
        }
        bool changed = (old_valid != cur.valid) || (cur.valid && old_min != cur.canon);
        if (changed || dump) {                                               // :401-435
            if (old_valid) pieces.push_back(Piece{last, c - 1, old_min, old_rev});
            last = c;
            old_valid = cur.valid; old_min = cur.canon; old_rev = is_rev;
        }
        c++;
    }
    if (old_valid) pieces.push_back(Piece{last, K - 1, old_min, old_rev});   // :441-450
}

template <class Key>
struct Entry {
    Key key;
    uint32_t minimizer;
    uint32_t order;
    uint8_t pos_min;
};

template <class Key>
struct Uniq {
    Key key;
    uint32_t first_order;
    uint8_t count, pos_min, seen;
};

template <class Key>
class SketchBuilder {
public:
    SketchBuilder(const uint32_t *packed, const SketchParams &prm) : w_(packed), prm_(prm)
    {
        k_ = prm.k; m_ = prm.m; d_ = k_ - m_;
        kmask_ = k_ * 2 == (int)sizeof(Key) * 8 ? ~(Key)0 : (((Key)1) << (2 * k_)) - 1;
        mmask_ = (1u << (2 * m_)) - 1u;
    }

    // handle_superkmer (SubSampler.cpp:243-302) on one piece of the record at `rec`.
    void take_piece(uint64_t rec, const Piece &pc)
    {
        const uint64_t nk = pc.last - pc.first + 1;
        st.selected_superkmers++;
        st.selected_kmers += nk;
        if (nk == (uint64_t)d_ + 1) st.maximal_superkmers++;
        tmp_.resize(nk);
        uint64_t b = rec + pc.first;
        Key key = 0;
        for (int i = 0; i < k_ - 1; i++) key = (key << 2) | base_at(w_, b + i);
        for (uint64_t t = 0; t < nk; t++) {
            key = ((key << 2) | base_at(w_, b + t + k_ - 1)) & kmask_;
            tmp_[t] = key;
        }
        for (uint64_t t = 0; t < nk; t++) {
            // oriented k-mers left to right: genome order is reversed for rev pieces
            Key kk = pc.rev ? revcomp(tmp_[nk - 1 - t], k_) : tmp_[t];
            unsigned pos = 255;                   // (uint8_t)string::npos, never seen in practice
            for (int q = 0; q <= d_; q++)
                if ((uint32_t)((kk >> (2 * (d_ - q))) & mmask_) == pc.minimizer) { pos = (unsigned)q; break; }
            entries_.push_back(Entry<Key>{kk, pc.minimizer, (uint32_t)entries_.size(), (uint8_t)pos});
        }
    }

    void write(std::vector<uint8_t> &out)
    {
        char hdr[160];
        int hl = snprintf(hdr, sizeof hdr, "%d %d %llu %f\n", 2 * k_ - m_, m_, (unsigned long long)st.selected_kmers,
                          prm_.s);                                            // :459
        out.insert(out.end(), hdr, hdr + hl);
        // buckets in ascending minimizer order, entries in emission order
        std::sort(entries_.begin(), entries_.end(), [](const Entry<Key> &a, const Entry<Key> &b) {
            return a.minimizer != b.minimizer ? a.minimizer < b.minimizer : a.order < b.order;
        });
        size_t i = 0;
        while (i < entries_.size()) {
            size_t j = i;
            while (j < entries_.size() && entries_[j].minimizer == entries_[i].minimizer) j++;
            write_bucket(i, j, out);
            st.buckets++;
            i = j;
        }
    }

    SketchStats st;

private:
    void num2txt(Key v, int n, char *dst)
    {
        for (int i = n - 1; i >= 0; i--) { dst[i] = kBase[(unsigned)(v & 3)]; v >>= 2; }
    }

    Uniq<Key> *lookup(Key key)
    {
        size_t lo = 0, hi = by_key_.size();
        while (lo < hi) {
            size_t mid = (lo + hi) >> 1;
            if (by_key_[mid].key < key) lo = mid + 1; else hi = mid;
        }
        return (lo < by_key_.size() && by_key_[lo].key == key) ? &by_key_[lo] : nullptr;
    }

    // find_next (SubSampler.cpp:566-602): probe order A,T,C,G.
    Uniq<Key> *step(Key cur, bool left)
    {
        static const unsigned order[4] = {0, 2, 1, 3};
        for (unsigned o : order) {
            Key nx = left ? (Key)((cur >> 2) | ((Key)o << (2 * k_ - 2))) : (Key)(((cur << 2) | o) & kmask_);
            Uniq<Key> *u = lookup(nx);
            if (u && !u->seen && u->count >= prm_.abundance) { u->seen = 1; return u; }
        }
        return nullptr;
    }

    void write_bucket(size_t b, size_t e, std::vector<uint8_t> &out)
    {
        const uint32_t minimizer = entries_[b].minimizer;
        // de-duplicate: first occurrence keeps pos_min, count wraps at 256
        idx_.resize(e - b);
        for (size_t i = 0; i < e - b; i++) idx_[i] = (uint32_t)(b + i);
        std::sort(idx_.begin(), idx_.end(), [&](uint32_t x, uint32_t y) {
            return entries_[x].key != entries_[y].key ? entries_[x].key < entries_[y].key : x < y;
        });
        by_key_.clear();
        for (size_t i = 0; i < idx_.size();) {
            size_t j = i;
            while (j < idx_.size() && entries_[idx_[j]].key == entries_[idx_[i]].key) j++;
            const Entry<Key> &f = entries_[idx_[i]];
            by_key_.push_back(Uniq<Key>{f.key, f.order, (uint8_t)((j - i) & 0xFF), f.pos_min, 0});
            i = j;
        }
        st.distinct_kmers += by_key_.size();
        ins_.resize(by_key_.size());
        for (size_t i = 0; i < ins_.size(); i++) ins_[i] = (uint32_t)i;
        std::sort(ins_.begin(), ins_.end(),
                  [&](uint32_t x, uint32_t y) { return by_key_[x].first_order < by_key_[y].first_order; });

        char mtxt[16];
        num2txt((Key)minimizer, m_, mtxt);
        out.insert(out.end(), mtxt, mtxt + m_);                               // :465-466
        maxs_.clear();
        text_.clear();
        const int full = 2 * k_ - m_;
        size_t cursor = 0;
        for (;;) {
            // find_first_kmer (:604-620): first unseen entry in insertion order
            while (cursor < ins_.size() &&
                   (by_key_[ins_[cursor]].seen || by_key_[ins_[cursor]].count < prm_.abundance))
                cursor++;
            if (cursor >= ins_.size()) break;
            Uniq<Key> *start = &by_key_[ins_[cursor]];
            start->seen = 1;
            // reconstruct_superkmer (:512-564)
            char buf[300];
            int lo = 150, hi = 150 + k_;
            num2txt(start->key, k_, buf + lo);
            uint64_t n_left = (uint64_t)d_ - start->pos_min, n_right = start->pos_min;
            Key cur = start->key;
            while (hi - lo != full) {
                if (n_left != 0) {
                    Uniq<Key> *nx = lo > 0 ? step(cur, true) : nullptr;
                    n_left--;
                    if (nx) buf[--lo] = kBase[(unsigned)(nx->key >> (2 * k_ - 2)) & 3];
                    else n_left = 0;
                    cur = (n_left == 0) ? start->key : nx->key;
                } else if (n_right != 0) {
                    Uniq<Key> *nx = step(cur, false);
                    n_right--;
                    if (!nx) break;
                    buf[hi++] = kBase[(unsigned)(nx->key & 3)];
                    cur = nx->key;
                } else {
                    break;
                }
            }
            st.out_superkmers++;
            const int len = hi - lo;
            if (len == full) {                                               // :479-485
                st.out_maximal++;
                maxs_.insert(maxs_.end(), buf + lo, buf + lo + d_);
                maxs_.insert(maxs_.end(), buf + lo + k_, buf + lo + k_ + d_);
            } else {                                                         // :486-494
                int q = -1;
                for (int t = 0; t + m_ <= len; t++)
                    if (!memcmp(buf + lo + t, mtxt, (size_t)m_)) { q = t; break; }
                if (q < 0) {
                    text_.insert(text_.end(), buf + lo, buf + hi);
                    text_.push_back('\n');
                    text_.push_back('\n');
                } else {
                    text_.insert(text_.end(), buf + lo, buf + lo + q);
                    text_.push_back('\n');
                    text_.insert(text_.end(), buf + lo + q + m_, buf + hi);
                    text_.push_back('\n');
                }
            }
        }
        // strCompressor (utils.cpp:48-68), accumulator starting at zero
        packed_.clear();
        if (!maxs_.empty()) {
            unsigned mod = (unsigned)(maxs_.size() % 4);
            packed_.push_back((uint8_t)mod);
            uint8_t c = 0;
            for (size_t i = 0; i < maxs_.size(); i++) {
                c = (uint8_t)(c + (((uint8_t)maxs_[i] >> 1) & 3));
                if ((i + 1) % 4 == 0) { packed_.push_back(c); c = 0; }
                c = (uint8_t)(c << 2);
            }
            if (mod) packed_.push_back(c);
        }
        uint32_t sz = (uint32_t)packed_.size();                               // :498-503
        const uint8_t *szp = reinterpret_cast<const uint8_t *>(&sz);
        out.insert(out.end(), szp, szp + 4);
        out.insert(out.end(), packed_.begin(), packed_.end());
        out.insert(out.end(), text_.begin(), text_.end());
        out.push_back('\n');
        out.push_back('\n');
    }

    const uint32_t *w_;
    SketchParams prm_;
    int k_, m_, d_;
    Key kmask_;
    uint32_t mmask_;
    std::vector<Key> tmp_;
    std::vector<Entry<Key>> entries_;
    std::vector<uint32_t> idx_, ins_;
    std::vector<Uniq<Key>> by_key_;
    std::vector<char> maxs_, text_;
    std::vector<uint8_t> packed_;
};

template <class Key>
void build_sketch_t(const uint32_t *packed, const std::vector<uint64_t> &rec_off, std::vector<spsp_hit> &hits,
                    const SketchParams &prm, std::vector<uint8_t> &out, SketchStats *stats)
{
    std::sort(hits.begin(), hits.end(), [](const spsp_hit &a, const spsp_hit &b) { return a.pos < b.pos; });
    SketchBuilder<Key> sb(packed, prm);
    std::vector<RHit> rh;
    std::vector<Piece> pieces;
    size_t hp = 0;
    const size_t n_rec = rec_off.empty() ? 0 : rec_off.size() - 1;
    uint64_t used = 0;
    for (size_t r = 0; r < n_rec; r++) {
        const uint64_t b = rec_off[r], e = rec_off[r + 1];
        sb.st.records++;
        sb.st.bases += e - b;
        rh.clear();
        while (hp < hits.size() && hits[hp].pos < e) {
            const spsp_hit &x = hits[hp++];
            if (x.pos < b || x.pos + (uint64_t)prm.m > e) continue;       // straddles a record boundary
            rh.push_back(RHit{x.pos - b, mmer_hash(x.canon), x.canon, x.rev != 0});
        }
        if (rh.empty() || e - b < (uint64_t)prm.k) continue;
        used += rh.size();
        pieces.clear();
        replay_record(rh, e - b, prm.k, prm.m, pieces);
        for (const Piece &pc : pieces) sb.take_piece(b, pc);
    }
    sb.write(out);
    if (stats) {
        *stats = sb.st;
        stats->hits = hits.size();
        stats->hits_used = used;
    }
}

}  // namespace

void build_sketch(const uint32_t *packed, const std::vector<uint64_t> &rec_off, std::vector<spsp_hit> &hits,
                  const SketchParams &prm, std::vector<uint8_t> &out, SketchStats *stats)
{
    if (prm.k <= 32) build_sketch_t<uint64_t>(packed, rec_off, hits, prm, out, stats);
    else build_sketch_t<u128>(packed, rec_off, hits, prm, out, stats);
}

}  // namespace spsp_host
