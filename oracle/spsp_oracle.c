/*
 * spsp_oracle.c -- CPU restatement of SuperSampler's sketch-and-compare path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may build, load or call this file.  The product (supersampler_b200/)
 * never links it and has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement
 * against tests/golden/ vectors that were produced by the unmodified reference
 * binaries (oracle/_ref, built by oracle/Makefile from /root/reference) with
 * tools/make_golden.py; when oracle/_ref is present the tests also run it live.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the reference root).  The scan is the *dense* literal state machine (one hash
 * per base plus rescans), i.e. the independent cross-check of the sparse
 * hit-list formulation used by the CUDA path.
 *
 * Undefined behaviour of the reference that its official build resolves as
 * "zero/false" is restated as zero/false here: strCompressor's accumulator
 * (utils.cpp:55), kmer_info::seen (SubSampler.cpp:283-286), dump / is_rev at
 * record start (SubSampler.cpp:353).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------ bytes */

typedef struct {
    uint8_t *p;
    size_t n, cap;
} buf_t;

static void buf_reserve(buf_t *b, size_t extra)
{
    if (b->n + extra <= b->cap) return;
    size_t c = b->cap ? b->cap : 256;
    while (c < b->n + extra) c *= 2;
    b->p = (uint8_t *)realloc(b->p, c);
    b->cap = c;
}
static void buf_put(buf_t *b, const void *src, size_t n)
{
    buf_reserve(b, n);
    memcpy(b->p + b->n, src, n);
    b->n += n;
}
static void buf_putc(buf_t *b, uint8_t c) { buf_put(b, &c, 1); }

/* ------------------------------------------------------------ 2-bit codec */

/* utils.cpp:13-16 nuc2int: (c/2)%4 -> A0 C1 T2 G3 (valid on cleaned text). */
static inline unsigned base_code(uint8_t c) { return (c >> 1) & 3u; }
/* utils.cpp:26-45 int2nuc / utils.cpp:168-183 num2str alphabet. */
static const char CODE2BASE[4] = {'A', 'C', 'T', 'G'};

/* utils.cpp:158-165 str2num: first base in the most significant bits. */
static u128 text_to_num(const uint8_t *s, size_t n)
{
    u128 v = 0;
    for (size_t i = 0; i < n; i++) v = (v << 2) + base_code(s[i]);
    return v;
}
/* utils.cpp:168-183 num2str. */
static void num_to_text(u128 v, unsigned n, uint8_t *out)
{
    for (unsigned i = 0; i < n; i++) {
        out[n - 1 - i] = (uint8_t)CODE2BASE[(unsigned)(v & 3)];
        v >>= 2;
    }
}

/* utils.cpp:449-462 rcbc: reverse complement of an n-mer held in a u64
 * (complement = code^2, then reverse the 2-bit digits, right-align). */
static uint64_t revcomp64(uint64_t x, unsigned n)
{
    uint64_t r = 0;
    for (unsigned i = 0; i < n; i++) {
        r = (r << 2) | ((x & 3u) ^ 2u);
        x >>= 2;
    }
    return r;
}
/* utils.cpp:397-438 rcb: same for a 128-bit k-mer. */
static u128 revcomp128(u128 x, unsigned n)
{
    u128 r = 0;
    for (unsigned i = 0; i < n; i++) {
        r = (r << 2) | (u128)(((unsigned)(x & 3u)) ^ 2u);
        x >>= 2;
    }
    return r;
}
/* utils.cpp:465-467 canonize(u64). */
static inline uint64_t canon64(uint64_t x, unsigned n)
{
    uint64_t r = revcomp64(x, n);
    return x < r ? x : r;
}
/* utils.cpp:470-472 canonize(u128). */
static inline u128 canon128(u128 x, unsigned n)
{
    u128 r = revcomp128(x, n);
    return x < r ? x : r;
}

/* ------------------------------------------------------------------ hash */

static inline uint64_t rotl64(uint64_t x, unsigned r) { return (x << r) | (x >> (64 - r)); }

/* SubSampler.cpp:64-67 unrevhash -> include/xxhash64.h:158-163 (static hash),
 * :115-148 (finalisation with totalLength 8 < 32: result = seed + Prime5 + 8,
 * one 8-byte round, avalanche), :188-191 processSingle.  The argument is the
 * canonical m-mer stored as a little-endian u64.  Everything after :67 in
 * unrevhash (the decycling classes) is unreachable. */
uint64_t spo_hash(uint64_t x)
{
    const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL,
                   P3 = 1609587929392839161ULL, P4 = 9650029242287828579ULL,
                   P5 = 2870177450012600261ULL;
    uint64_t r = 1312ULL + P5 + 8ULL;
    uint64_t lane = rotl64(x * P2, 31) * P1;
    r = rotl64(r ^ lane, 27) * P1 + P4;
    r ^= r >> 33;
    r *= P2;
    r ^= r >> 29;
    r *= P3;
    r ^= r >> 32;
    return r;
}

/* SubSampler.cpp:622-631 compute_threshold, selected at SubSampler.h:79-83
 * (s <= 1 -> all ones).  x87 long double on purpose. */
uint64_t spo_threshold(unsigned k, unsigned m, double s)
{
    if (!(s > 1)) return ~(uint64_t)0;
    uint64_t w = (uint64_t)k - m + 1;
    long double frac = (long double)1 / s;
    long double root = powl((long double)1 - frac, (long double)1 / w);
    long double res = ((long double)1 - root) * ((uint64_t)1 << 63);
    return (uint64_t)res * 2;
}

/* ------------------------------------------------------- FASTA record cut */

/* utils.cpp:706-718 getLineFasta + utils.cpp:675-702 clean_dna.
 * Reads one record starting at *pos: the first line is dropped whatever it
 * holds, then lines are concatenated until a line starts with '>' (or EOF);
 * every char outside ACGTacgt is deleted and the rest upper-cased.
 * Returns 0 once the stream would report eof() before the call
 * (SubSampler.cpp:334 loop condition). */
static int next_record(const uint8_t *txt, size_t n, size_t *pos, int *eof, buf_t *rec)
{
    if (*eof) return 0;
    rec->n = 0;
    size_t p = *pos;
    /* header getline */
    while (p < n && txt[p] != '\n') p++;
    if (p < n) p++; else *eof = 1;              /* hit EOF inside getline */
    for (;;) {
        if (p >= n) { *eof = 1; break; }        /* peek() == EOF sets eofbit */
        /* utils.cpp:709-713 holds peek() in a (signed) char and compares it with EOF: a line that starts with the
         * byte 0xFF ends the record exactly like one that starts with '>' (and is then consumed as a header line) --
         * verified against the reference binary, tests/golden case "ffline" */
        if (txt[p] == '>' || txt[p] == 0xFF) break;
        size_t q = p;
        while (q < n && txt[q] != '\n') q++;
        for (size_t i = p; i < q; i++) {
            uint8_t c = txt[i];
            switch (c) {
            case 'A': case 'C': case 'G': case 'T': buf_putc(rec, c); break;
            case 'a': case 'c': case 'g': case 't': buf_putc(rec, (uint8_t)(c - 32)); break;
            default: break;
            }
        }
        if (q < n) p = q + 1; else { p = n; }
    }
    *pos = p;
    return 1;
}

/* --------------------------------------------- insertion-ordered k-mer map */

/* SubSampler.h:23-27 kmer_info; the container is ankerl::unordered_dense
 * (include/unordered_dense.h:429: values live in a vector, iteration order =
 * insertion order), which SubSampler.cpp:604-620 relies on. */
typedef struct {
    u128 key;
    uint8_t count, pos_min, seen;
} kent_t;

typedef struct {
    uint32_t minimizer;
    kent_t *e;
    uint32_t n, cap;
    uint32_t *slot;     /* open addressing, value = entry index + 1 */
    uint32_t nslot;
} bucket_t;

static inline uint64_t mix128(u128 k)
{
    uint64_t h = (uint64_t)k ^ ((uint64_t)(k >> 64) * 0x9E3779B97F4A7C15ULL);
    h ^= h >> 32; h *= 0xD6E8FEB86659FD93ULL; h ^= h >> 32;
    return h;
}
static void bucket_rehash(bucket_t *b, uint32_t nslot)
{
    free(b->slot);
    b->slot = (uint32_t *)calloc(nslot, sizeof(uint32_t));
    b->nslot = nslot;
    for (uint32_t i = 0; i < b->n; i++) {
        uint32_t s = (uint32_t)(mix128(b->e[i].key) & (nslot - 1));
        while (b->slot[s]) s = (s + 1) & (nslot - 1);
        b->slot[s] = i + 1;
    }
}
static kent_t *bucket_find(bucket_t *b, u128 key)
{
    if (!b->nslot) return NULL;
    uint32_t s = (uint32_t)(mix128(key) & (b->nslot - 1));
    while (b->slot[s]) {
        kent_t *e = &b->e[b->slot[s] - 1];
        if (e->key == key) return e;
        s = (s + 1) & (b->nslot - 1);
    }
    return NULL;
}
static kent_t *bucket_append(bucket_t *b, u128 key)
{
    if (b->n == b->cap) {
        b->cap = b->cap ? b->cap * 2 : 8;
        b->e = (kent_t *)realloc(b->e, b->cap * sizeof(kent_t));
    }
    if ((b->n + 1) * 2 > b->nslot) bucket_rehash(b, b->nslot ? b->nslot * 2 : 16);
    kent_t *e = &b->e[b->n];
    e->key = key; e->count = 0; e->pos_min = 0; e->seen = 0;
    uint32_t s = (uint32_t)(mix128(key) & (b->nslot - 1));
    while (b->slot[s]) s = (s + 1) & (b->nslot - 1);
    b->slot[s] = ++b->n;
    return e;
}

/* SubSampler.h:62 minimizer_map: std::map<uint32_t, ...> (ascending keys). */
typedef struct {
    bucket_t *b;
    uint32_t n, cap;
    uint32_t *slot;
    uint32_t nslot;
} bmap_t;

static bucket_t *bmap_get(bmap_t *m, uint32_t minimizer)
{
    if (!m->nslot) {
        m->nslot = 1024;
        m->slot = (uint32_t *)calloc(m->nslot, sizeof(uint32_t));
    }
    uint32_t s = (minimizer * 2654435761u) & (m->nslot - 1);
    while (m->slot[s]) {
        if (m->b[m->slot[s] - 1].minimizer == minimizer) return &m->b[m->slot[s] - 1];
        s = (s + 1) & (m->nslot - 1);
    }
    if (m->n == m->cap) {
        m->cap = m->cap ? m->cap * 2 : 64;
        m->b = (bucket_t *)realloc(m->b, m->cap * sizeof(bucket_t));
    }
    bucket_t *b = &m->b[m->n];
    memset(b, 0, sizeof *b);
    b->minimizer = minimizer;
    m->slot[s] = ++m->n;
    if (m->n * 2 > m->nslot) {
        uint32_t ns = m->nslot * 2;
        uint32_t *sl = (uint32_t *)calloc(ns, sizeof(uint32_t));
        for (uint32_t i = 0; i < m->n; i++) {
            uint32_t t = (m->b[i].minimizer * 2654435761u) & (ns - 1);
            while (sl[t]) t = (t + 1) & (ns - 1);
            sl[t] = i + 1;
        }
        free(m->slot);
        m->slot = sl; m->nslot = ns;
    }
    return b;
}
static void bmap_free(bmap_t *m)
{
    for (uint32_t i = 0; i < m->n; i++) { free(m->b[i].e); free(m->b[i].slot); }
    free(m->b); free(m->slot);
    memset(m, 0, sizeof *m);
}

/* ----------------------------------------------------------- sketch state */

typedef struct {
    uint64_t records, bases;              /* records >= k, cleaned bases in them */
    uint64_t total_kmers, total_superkmers;
    uint64_t selected_kmers, selected_superkmers, maximal_superkmers;
    uint64_t buckets, distinct_kmers, out_superkmers, out_maximal;
    uint64_t pb_events;                   /* SubSampler.cpp:265-269 "PB" */
} spo_stats;

typedef struct {
    unsigned k, m, d, abundance;
    uint64_t thr, mmask;
    u128 kmask;
    bmap_t map;
    spo_stats st;
    /* optional trace of selected k-mer occurrences (record, start) */
    uint64_t *sel; size_t nsel, capsel;
    int trace;
} sk_t;

/* SubSampler.cpp:81-169 regular_minimizer_pos: scan the d+1 m-mers of one
 * k-mer from the right-most to the left-most; strict '<' on the hash; the
 * returned position keeps the reference's quirks (reverse at the right end
 * -> 0, :88-93; same-orientation ties :149-164). */
static uint64_t rescan_kmer(const sk_t *S, u128 seq, uint64_t *position, int *is_rev)
{
    const unsigned m = S->m, d = S->d;
    uint64_t fw = (uint64_t)seq & S->mmask;
    uint64_t best = canon64(fw, m);
    int rev = (best != fw);
    uint64_t pos = rev ? 0 : d;
    uint64_t hbest = spo_hash(best);
    for (unsigned j = 1; j <= d; j++) {
        seq >>= 2;
        fw = (uint64_t)seq & S->mmask;
        uint64_t cn = canon64(fw, m);
        int lrev = (cn != fw);
        uint64_t h = spo_hash(cn);
        if (hbest > h) {
            pos = d - j; best = cn; rev = lrev; hbest = h;
        } else if (cn == best && lrev == rev) {
            if (rev && pos > j) pos = j;
            if (!rev && pos > (uint64_t)d - j) pos = d - j;
        }
    }
    *position = pos;
    *is_rev = rev;
    return best;
}

/* SubSampler.cpp:243-302 handle_superkmer: orient the piece, then insert each
 * of its k-mers (as oriented, not canonicalised) into the minimizer's bucket:
 * first sight -> {count 1, leftmost position of the minimizer text, unseen},
 * otherwise count++ on a uint8 (wraps at 256). */
static void take_piece(sk_t *S, const uint8_t *piece, size_t len, uint32_t minimizer, int rev,
                       uint64_t rec_id, uint64_t start)
{
    const unsigned k = S->k, m = S->m;
    uint8_t *tmp = (uint8_t *)malloc(len);
    if (rev) {                                   /* utils.cpp:142-148 revComp */
        for (size_t i = 0; i < len; i++) {
            uint8_t c = piece[len - 1 - i];
            tmp[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
        }
    } else {
        memcpy(tmp, piece, len);
    }
    S->st.selected_superkmers++;
    S->st.selected_kmers += len - k + 1;
    if (len == 2u * k - m) S->st.maximal_superkmers++;
    uint8_t mtxt[16];
    num_to_text(minimizer, m, mtxt);
    bucket_t *b = bmap_get(&S->map, minimizer);
    for (size_t i = 0; i + k <= len; i++) {
        unsigned pos = 255;                      /* (uint8_t)string::npos */
        for (unsigned q = 0; q + m <= k; q++)
            if (!memcmp(tmp + i + q, mtxt, m)) { pos = q; break; }
        if (pos == 255) S->st.pb_events++;
        u128 key = text_to_num(tmp + i, k);
        kent_t *e = bucket_find(b, key);
        if (e) {
            e->count++;
        } else {
            e = bucket_append(b, key);
            e->count = 1; e->pos_min = (uint8_t)pos; e->seen = 0;
        }
        if (S->trace) {
            if (S->nsel + 2 > S->capsel) {
                S->capsel = S->capsel ? S->capsel * 2 : 1024;
                S->sel = (uint64_t *)realloc(S->sel, S->capsel * sizeof(uint64_t));
            }
            S->sel[S->nsel++] = rec_id;
            S->sel[S->nsel++] = rev ? start + (len - k - i) : start + i;
        }
    }
    free(tmp);
}

/* SubSampler.cpp:352-454: the per-record minimizer state machine. */
static void scan_record(sk_t *S, const uint8_t *ref, size_t n, uint64_t rec_id)
{
    const unsigned k = S->k, m = S->m;
    if (n < k) return;                            /* :340-343 */
    S->st.records++;
    S->st.bases += n;
    int is_rev = 0, old_rev = 0, dump = 0;
    uint64_t last = 0, position_min = 0, i = 0;
    u128 seq = text_to_num(ref, k);                                   /* :359 */
    uint64_t fw = (uint64_t)text_to_num(ref + k - m, m);              /* :360 */
    uint64_t rc = revcomp64(fw, m);                                   /* :361 */
    uint32_t minimizer = (uint32_t)rescan_kmer(S, seq, &position_min, &old_rev); /* :363 */
    uint32_t old_minimizer = minimizer;
    uint64_t hash_min = spo_hash(minimizer);                          /* :365 */
    for (; i + k < n; i++) {                                          /* :367 */
        unsigned c = base_code(ref[i + k]);
        seq = ((seq << 2) + c) & S->kmask;                            /* :29-34 */
        fw = ((fw << 2) + c) & S->mmask;                              /* :36-41 */
        rc = (rc >> 2) + ((uint64_t)(c ^ 2u) << (2 * m - 2));         /* :49-53 */
        uint64_t cn = fw < rc ? fw : rc;
        uint64_t h = spo_hash(cn);
        if (h < hash_min) {                                           /* :374-388 */
            minimizer = (uint32_t)cn; hash_min = h;
            position_min = i + k - m + 1;
            is_rev = (cn != fw);
        } else if (i >= position_min) {                               /* :391-398 */
            minimizer = (uint32_t)rescan_kmer(S, seq, &position_min, &is_rev);
            dump = 1;
            hash_min = spo_hash(minimizer);
            position_min += i + 1;
        }
        if (old_minimizer != minimizer || dump) {                     /* :401 */
            dump = 0;
            if (spo_hash(old_minimizer) <= S->thr)                    /* :405 */
                take_piece(S, ref + last, i + k - last, old_minimizer, old_rev, rec_id, last);
            S->st.total_kmers += i - last + 1;
            S->st.total_superkmers++;
            last = i + 1;
            old_minimizer = minimizer;
            old_rev = is_rev;
        }
    }
    if (n - last > (size_t)k - 1) {                                   /* :441-454 */
        if (spo_hash(old_minimizer) <= S->thr)
            take_piece(S, ref + last, i + k - last, old_minimizer, old_rev, rec_id, last);
        S->st.total_kmers += i - last + 1;
        S->st.total_superkmers++;
    }
}

/* SubSampler.cpp:604-620 find_first_kmer. */
static kent_t *first_unseen(bucket_t *b, unsigned abundance)
{
    for (uint32_t i = 0; i < b->n; i++)
        if (!b->e[i].seen && b->e[i].count >= abundance) { b->e[i].seen = 1; return &b->e[i]; }
    return NULL;
}
/* SubSampler.cpp:566-602 find_next: probe order A,T,C,G (:568). */
static int step_next(const sk_t *S, bucket_t *b, u128 cur, int left, u128 *out)
{
    static const unsigned order[4] = {0, 2, 1, 3};
    for (int t = 0; t < 4; t++) {
        u128 nx;
        if (left) nx = (cur >> 2) + ((u128)order[t] << (2 * S->k - 2));
        else      nx = ((cur << 2) + order[t]) & S->kmask;
        kent_t *e = bucket_find(b, nx);
        if (e && !e->seen && e->count >= S->abundance) { e->seen = 1; *out = nx; return 1; }
    }
    return 0;
}

/* utils.cpp:48-68 strCompressor with the accumulator starting at zero. */
static void pack_bases(const buf_t *txt, buf_t *out)
{
    out->n = 0;
    if (!txt->n) return;
    unsigned mod = (unsigned)(txt->n % 4);
    buf_putc(out, (uint8_t)mod);
    uint8_t c = 0;
    for (size_t i = 0; i < txt->n; i++) {
        c = (uint8_t)(c + base_code(txt->p[i]));
        if ((i + 1) % 4 == 0) { buf_putc(out, c); c = 0; }
        c = (uint8_t)(c << 2);
    }
    if (mod) buf_putc(out, c);
}

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

/* SubSampler.cpp:459-504 (writer loop) + :512-564 reconstruct_superkmer. */
static void write_sketch(sk_t *S, double s, buf_t *out)
{
    const unsigned k = S->k, m = S->m, d = S->d;
    char hdr[128];
    int hl = snprintf(hdr, sizeof hdr, "%u %u %llu %f\n", 2 * k - m, m,
                      (unsigned long long)S->st.selected_kmers, s);
    buf_put(out, hdr, (size_t)hl);
    /* ascending minimizer order (std::map) */
    uint32_t nb = S->map.n;
    uint32_t *ord = (uint32_t *)malloc((nb ? nb : 1) * sizeof(uint32_t));
    uint64_t *okey = (uint64_t *)malloc((nb ? nb : 1) * sizeof(uint64_t));
    for (uint32_t i = 0; i < nb; i++) okey[i] = ((uint64_t)S->map.b[i].minimizer << 32) | i;
    qsort(okey, nb, sizeof(uint64_t), cmp_u64);
    for (uint32_t i = 0; i < nb; i++) ord[i] = (uint32_t)okey[i];
    free(okey);
    buf_t maxs = {0}, txt = {0}, packed = {0}, sk = {0};
    uint8_t mtxt[16], ktxt[64];
    S->st.buckets = nb;
    for (uint32_t bi = 0; bi < nb; bi++) {
        bucket_t *b = &S->map.b[ord[bi]];
        num_to_text(b->minimizer, m, mtxt);
        buf_put(out, mtxt, m);                                        /* :465-466 */
        maxs.n = txt.n = 0;
        S->st.distinct_kmers += b->n;
        for (;;) {
            kent_t *st = first_unseen(b, S->abundance);               /* :472-476 */
            if (!st) break;
            /* :512-564 */
            u128 start = st->key, cur = start, nx;
            uint64_t n_left = (uint64_t)d - st->pos_min, n_right = st->pos_min;
            sk.n = 0;
            num_to_text(start, k, ktxt);
            buf_put(&sk, ktxt, k);
            while (sk.n != 2u * k - m) {
                if (n_left != 0) {
                    int ok = step_next(S, b, cur, 1, &nx);
                    n_left--;
                    if (ok) {
                        buf_reserve(&sk, 1);
                        memmove(sk.p + 1, sk.p, sk.n);
                        sk.p[0] = (uint8_t)CODE2BASE[(unsigned)(nx >> (2 * k - 2)) & 3];
                        sk.n++;
                    } else {
                        n_left = 0;
                    }
                    cur = (n_left == 0) ? start : nx;
                } else if (n_right != 0) {
                    int ok = step_next(S, b, cur, 0, &nx);
                    n_right--;
                    if (!ok) break;
                    buf_putc(&sk, (uint8_t)CODE2BASE[(unsigned)(nx & 3)]);
                    cur = nx;
                } else {
                    break;
                }
            }
            S->st.out_superkmers++;
            if (sk.n == 2u * k - m) {                                 /* :479-485 */
                S->st.out_maximal++;
                buf_put(&maxs, sk.p, d);
                buf_put(&maxs, sk.p + k, d);
            } else {                                                  /* :486-494 */
                size_t q = 0;
                int found = 0;
                for (; q + m <= sk.n; q++)
                    if (!memcmp(sk.p + q, mtxt, m)) { found = 1; break; }
                if (!found) q = sk.n;             /* npos: substr(0,npos) = whole string */
                buf_put(&txt, sk.p, q);
                buf_putc(&txt, '\n');
                if (found) buf_put(&txt, sk.p + q + m, sk.n - q - m);
                buf_putc(&txt, '\n');
            }
        }
        pack_bases(&maxs, &packed);                                   /* :498-503 */
        uint32_t sz = (uint32_t)packed.n;
        buf_put(out, &sz, 4);
        buf_put(out, packed.p, packed.n);
        buf_put(out, txt.p, txt.n);
        buf_put(out, "\n\n", 2);
    }
    free(ord); free(maxs.p); free(txt.p); free(packed.p); free(sk.p);
}

static void sk_init(sk_t *S, unsigned k, unsigned m, double s, unsigned abundance)
{
    memset(S, 0, sizeof *S);
    S->k = k; S->m = m; S->d = k - m; S->abundance = abundance;
    S->thr = spo_threshold(k, m, s);
    S->mmask = ((uint64_t)1 << (2 * m)) - 1;
    S->kmask = (((u128)1) << (2 * k)) - 1;
}

/* Sketch one FASTA text (already inflated) -> sketch bytes (before gzip).
 * SubSampler.cpp:306-510 parse_fasta_test. */
int spo_sketch(const uint8_t *fasta, size_t n, unsigned k, unsigned m, double s,
               unsigned abundance, uint8_t **out, size_t *out_len, spo_stats *stats,
               uint64_t **sel_out, size_t *nsel_out)
{
    if (k > 63 || m > 15 || m >= k || !(k & 1) || !(m & 1)) return -1;
    sk_t S;
    sk_init(&S, k, m, s, abundance);
    S.trace = sel_out != NULL;
    buf_t rec = {0};
    size_t pos = 0;
    int eof = 0;
    uint64_t rid = 0;
    while (next_record(fasta, n, &pos, &eof, &rec)) {
        scan_record(&S, rec.p, rec.n, rid);
        rid++;
    }
    buf_t o = {0};
    write_sketch(&S, s, &o);
    *out = o.p; *out_len = o.n;
    if (stats) *stats = S.st;
    if (sel_out) { *sel_out = S.sel; *nsel_out = S.nsel / 2; } else free(S.sel);
    bmap_free(&S.map);
    free(rec.p);
    return 0;
}

/* Cleaned records of a FASTA text: concatenated bases + offsets (n_rec+1),
 * including records shorter than k (callers filter).  For kernel parity tests. */
int spo_clean(const uint8_t *fasta, size_t n, uint8_t **bases, uint64_t **offs, size_t *n_rec)
{
    buf_t rec = {0}, all = {0};
    size_t pos = 0, cap = 16, nr = 0;
    int eof = 0;
    uint64_t *o = (uint64_t *)malloc(cap * sizeof(uint64_t));
    o[0] = 0;
    while (next_record(fasta, n, &pos, &eof, &rec)) {
        buf_put(&all, rec.p, rec.n);
        if (nr + 2 > cap) { cap *= 2; o = (uint64_t *)realloc(o, cap * sizeof(uint64_t)); }
        o[++nr] = all.n;
    }
    free(rec.p);
    if (!all.p) all.p = (uint8_t *)malloc(1);
    *bases = all.p; *offs = o; *n_rec = nr;
    return 0;
}

/* Closed-form hit list of one cleaned record: positions p
 * whose canonical m-mer hash is <= T.  Returns count; fills up to cap. */
size_t spo_hits(const uint8_t *seq, size_t n, unsigned m, uint64_t thr,
                uint64_t *pos, uint32_t *canon, uint8_t *rev, uint64_t *hash, size_t cap)
{
    if (n < m) return 0;
    uint64_t mmask = ((uint64_t)1 << (2 * m)) - 1, fw = 0, rc = 0;
    size_t cnt = 0;
    for (size_t i = 0; i < n; i++) {
        unsigned c = base_code(seq[i]);
        fw = ((fw << 2) + c) & mmask;
        rc = (rc >> 2) + ((uint64_t)(c ^ 2u) << (2 * m - 2));
        if (i + 1 < m) continue;
        uint64_t cn = fw < rc ? fw : rc;
        uint64_t h = spo_hash(cn);
        if (h <= thr) {
            if (cnt < cap) {
                pos[cnt] = i + 1 - m; canon[cnt] = (uint32_t)cn;
                rev[cnt] = (uint8_t)(cn != fw); hash[cnt] = h;
            }
            cnt++;
        }
    }
    return cnt;
}

void spo_free(void *p) { free(p); }

/* ================================================================ compare */

typedef struct {
    uint64_t minimizer;
    u128 *km;           /* distinct canonical k-mers of the bucket */
    uint32_t n;
} cbucket_t;

typedef struct {
    cbucket_t *b;
    uint32_t n, cap;
    unsigned k, m;
} csketch_t;

static int cmp_u128(const void *a, const void *b)
{
    u128 x = *(const u128 *)a, y = *(const u128 *)b;
    return x < y ? -1 : x > y;
}

/* utils.cpp:71-111 strDecompressor. */
static void unpack_bases(const uint8_t *p, size_t n, buf_t *out)
{
    out->n = 0;
    if (!n) return;
    unsigned mod = p[0];
    size_t last = mod == 0 ? n : n - 1;
    for (size_t i = 1; i < last; i++)
        for (int sh = 6; sh >= 0; sh -= 2) buf_putc(out, (uint8_t)CODE2BASE[(p[i] >> sh) & 3]);
    if (mod != 0 && last >= 1 && last < n) {
        uint8_t f[8] = {0};
        uint8_t v = p[last];
        for (unsigned i = 0; i < mod + 1 && i < 8; i++) { f[mod - i] = (uint8_t)CODE2BASE[v & 3]; v >>= 2; }
        for (unsigned i = 0; i < mod; i++) buf_putc(out, f[i]);
    }
}

static void cb_add(u128 **arr, uint32_t *n, uint32_t *cap, u128 v)
{
    if (*n == *cap) { *cap = *cap ? *cap * 2 : 64; *arr = (u128 *)realloc(*arr, *cap * sizeof(u128)); }
    (*arr)[(*n)++] = v;
}

/* Decode one sketch (bytes after gunzip) into buckets of distinct canonical
 * k-mers.  Comparator.cpp:23-37 get_header_info, :291-323 increment_files,
 * :97-154 / :177-264 bucket decode, :78-92 inject_minimizer. */
static int decode_sketch(const uint8_t *p, size_t n, csketch_t *S)
{
    memset(S, 0, sizeof *S);
    size_t pos = 0;
    while (pos < n && p[pos] != '\n') pos++;
    if (pos >= n) return -1;
    unsigned L = 0, m = 0;
    if (sscanf((const char *)p, "%u %u", &L, &m) != 2) return -1;
    unsigned k = (L + m) / 2, d = (L - m) / 2;
    S->k = k; S->m = m;
    pos++;
    u128 kmask = (((u128)1) << (2 * k)) - 1;
    buf_t txt = {0}, full = {0};
    while (pos + m <= n) {
        if (S->n == S->cap) { S->cap = S->cap ? S->cap * 2 : 64; S->b = (cbucket_t *)realloc(S->b, S->cap * sizeof(cbucket_t)); }
        cbucket_t *b = &S->b[S->n++];
        b->minimizer = (uint64_t)text_to_num(p + pos, m);
        const uint8_t *mtxt = p + pos;
        pos += m;
        uint32_t sz = 0;
        if (pos + 4 > n) { b->km = NULL; b->n = 0; break; }
        memcpy(&sz, p + pos, 4);
        pos += 4;
        if (pos + sz > n) sz = (uint32_t)(n - pos);
        unpack_bases(p + pos, sz, &txt);
        pos += sz;
        u128 *arr = NULL; uint32_t na = 0, ca = 0;
        /* maximal super-k-mers: prefix(d) + minimizer + suffix(d) each */
        full.n = 0;
        if (txt.n) {
            for (size_t i = 0; i < txt.n; i += d) {
                size_t a = txt.n - i < d ? txt.n - i : d;
                buf_put(&full, txt.p + i, a);
                i += d;
                buf_put(&full, mtxt, m);
                if (i < txt.n) { a = txt.n - i < d ? txt.n - i : d; buf_put(&full, txt.p + i, a); }
            }
        }
        if (full.n >= k) {
            size_t i = 0;
            while (i + k <= full.n) {
                u128 cur = text_to_num(full.p + i, k - 1);
                for (unsigned j = 0; j < d + 1; j++) {
                    if (i + k - 1 >= full.n) break;   /* .at() would throw; not reachable on valid files */
                    cur = ((cur << 2) + base_code(full.p[i + k - 1])) & kmask;
                    cb_add(&arr, &na, &ca, canon128(cur, k));
                    i++;
                }
                i += k - 1;
            }
        }
        /* non-maximal: pairs of text lines until two empty lines */
        for (;;) {
            size_t a0 = pos; while (pos < n && p[pos] != '\n') pos++;
            size_t a1 = pos; if (pos < n) pos++;
            size_t b0 = pos; while (pos < n && p[pos] != '\n') pos++;
            size_t b1 = pos; if (pos < n) pos++;
            if (a1 == a0 && b1 == b0) break;
            full.n = 0;
            buf_put(&full, p + a0, a1 - a0);
            buf_put(&full, mtxt, m);
            buf_put(&full, p + b0, b1 - b0);
            if (full.n >= k) {
                u128 cur = text_to_num(full.p, k - 1);
                for (size_t i = 0; i + k <= full.n; i++) {
                    cur = ((cur << 2) + base_code(full.p[i + k - 1])) & kmask;
                    cb_add(&arr, &na, &ca, canon128(cur, k));
                }
            }
            if (pos >= n) break;
        }
        /* distinct */
        if (na) {
            qsort(arr, na, sizeof(u128), cmp_u128);
            uint32_t w = 1;
            for (uint32_t i = 1; i < na; i++) if (arr[i] != arr[w - 1]) arr[w++] = arr[i];
            na = w;
        }
        b->km = arr; b->n = na;
    }
    free(txt.p); free(full.p);
    return 0;
}

/* Comparator.cpp:39-74 compare_sketches, :328-359 findMin, :97-154
 * skip_bucket, :177-264 count_intersection, :269-287 compute_scores.
 * inter is n*n row-major, only i<j entries are written (score_A key i*n+j);
 * sizes[i] = nb_kmer_seen_infile[i]. */
int spo_compare(const uint8_t *const *sk, const size_t *len, unsigned n, unsigned query_size,
                uint32_t *inter, uint64_t *sizes, unsigned *k_out, unsigned *m_out)
{
    csketch_t *S = (csketch_t *)calloc(n ? n : 1, sizeof(csketch_t));
    uint32_t *cur = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
    for (unsigned i = 0; i < n; i++)
        if (decode_sketch(sk[i], len[i], &S[i]) != 0) { free(S); free(cur); return -1; }
    memset(inter, 0, (size_t)n * n * sizeof(uint32_t));
    memset(sizes, 0, (size_t)n * sizeof(uint64_t));
    if (n) { if (k_out) *k_out = S[n - 1].k; if (m_out) *m_out = S[n - 1].m; }
    unsigned *idx = (unsigned *)malloc((n ? n : 1) * sizeof(unsigned));
    for (;;) {
        uint64_t mn = ~(uint64_t)0;
        unsigned ni = 0;
        int qfound = 0;
        for (unsigned i = 0; i < n; i++) {
            if (cur[i] >= S[i].n) continue;
            uint64_t v = S[i].b[cur[i]].minimizer;
            if (v < mn) { mn = v; ni = 0; idx[ni++] = i; qfound = i < query_size; }
            else if (v == mn) { idx[ni++] = i; if (i < query_size) qfound = 1; }
        }
        if (!ni) break;
        for (unsigned a = 0; a < ni; a++) sizes[idx[a]] += S[idx[a]].b[cur[idx[a]]].n;
        if (ni > 1 && qfound) {
            /* shared k-mers: every pair of files holding the same canonical k-mer */
            for (unsigned a = 0; a < ni; a++)
                for (unsigned b = a + 1; b < ni; b++) {
                    cbucket_t *x = &S[idx[a]].b[cur[idx[a]]], *y = &S[idx[b]].b[cur[idx[b]]];
                    uint32_t p = 0, q = 0, c = 0;
                    while (p < x->n && q < y->n) {
                        if (x->km[p] < y->km[q]) p++;
                        else if (y->km[q] < x->km[p]) q++;
                        else { c++; p++; q++; }
                    }
                    inter[(size_t)idx[a] * n + idx[b]] += c;
                }
        }
        for (unsigned a = 0; a < ni; a++) cur[idx[a]]++;
    }
    for (unsigned i = 0; i < n; i++) {
        for (uint32_t b = 0; b < S[i].n; b++) free(S[i].b[b].km);
        free(S[i].b);
    }
    free(S); free(cur); free(idx);
    return 0;
}

/* Comparator.cpp:362-408 print_containment (jaccard=0) and :412-460
 * print_jaccard (jaccard=1): header of names, containment has one blank line,
 * rows i<query_size; diag "1"; zero count "0"; value < min_threshold "0";
 * else setprecision(p) default-float == printf("%.*g"). */
int spo_csv(const char *const *names, unsigned n, unsigned query_size, const uint32_t *inter,
            const uint64_t *sizes, int jaccard, unsigned precision, double min_threshold,
            uint8_t **out, size_t *out_len)
{
    buf_t o = {0};
    for (unsigned i = 0; i < n; i++) {
        buf_put(&o, names[i], strlen(names[i]));
        buf_putc(&o, i + 1 == n ? '\n' : ',');
    }
    if (!jaccard) buf_putc(&o, '\n');
    char tmp[64];
    for (unsigned i = 0; i < n && i < query_size; i++)
        for (unsigned j = 0; j < n; j++) {
            if (i == j) buf_putc(&o, '1');
            else {
                uint32_t c = i < j ? inter[(size_t)i * n + j] : inter[(size_t)j * n + i];
                if (!c) buf_putc(&o, '0');
                else {
                    double sc = jaccard ? (double)c / (double)(sizes[i] + sizes[j] - c)
                                        : (double)c / (double)sizes[i];
                    if (sc < min_threshold) buf_putc(&o, '0');
                    else { int l = snprintf(tmp, sizeof tmp, "%.*g", (int)precision, sc); buf_put(&o, tmp, (size_t)l); }
                }
            }
            buf_putc(&o, j + 1 == n ? '\n' : ',');
        }
    *out = o.p; *out_len = o.n;
    return 0;
}
