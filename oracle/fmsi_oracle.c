/* fmsi_oracle.c — TEST INFRASTRUCTURE ONLY. See fmsi_oracle.h for scope and the pinning status.
 *
 * A plain-C, single-threaded restatement of the reference's query path. Every function names
 * the reference file:line it follows (paths relative to the reference root). It deliberately
 * keeps the reference's data representation (three wavelet-tree bitvectors, RRR<63> mask decoded
 * on the fly, plain kLCP bits) so that it is an independent check of the product's GPU layout.
 */
#define _POSIX_C_SOURCE 200809L
#include "fmsi_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* src/kmers.h:3-20 — nucleotideToInt: A/a=0 C/c=1 G/g=2 T/t=3, everything else 4.             */
static int nucleotide_to_int(unsigned char ch) {
    switch (ch) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}
/* src/kmers.h:21-38 — complementaryNucleotide keeps case, maps non-ACGT to 'N'. */
static char complementary_nucleotide(unsigned char ch) {
    switch (ch) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    default: return 'N';
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Plain bit vector = sdsl::int_vector<1>; bit p is bit (p & 63) of little-endian word p >> 6
 * (sdsl-lite/include/sdsl/int_vector.hpp:1563-1595 for the on-disk form).                     */
typedef struct {
    uint64_t nbits;
    uint64_t *w;      /* (nbits>>6)+2 words, zero padded: rank(size()) may touch word nbits>>6
                         (sdsl pads the same way, memory_management.hpp:351-355)               */
    uint64_t *cum;    /* cum[b] = popcount of words [0, 8b): prefix count per 512 bits          */
} obv;

static int popcnt64(uint64_t x) { return __builtin_popcountll(x); }

static int obv_alloc(obv *b, uint64_t nbits) {
    b->nbits = nbits;
    b->w = (uint64_t *)calloc((nbits >> 6) + 2, 8);
    b->cum = NULL;
    return b->w ? 0 : -1;
}
static void obv_free(obv *b) {
    free(b->w);
    free(b->cum);
    b->w = b->cum = NULL;
    b->nbits = 0;
}
static inline int obv_get(const obv *b, uint64_t p) { return (int)((b->w[p >> 6] >> (p & 63)) & 1); }
static inline void obv_set(obv *b, uint64_t p, int v) {
    if (v) b->w[p >> 6] |= (1ull << (p & 63));
    else b->w[p >> 6] &= ~(1ull << (p & 63));
}
/* get_int: len (<=64) bits starting at bit position pos, LSB first (sdsl bits::read_int). */
static uint64_t obv_get_int(const obv *b, uint64_t pos, unsigned len) {
    if (len == 0) return 0;
    uint64_t wi = pos >> 6;
    unsigned off = (unsigned)(pos & 63);
    uint64_t x = b->w[wi] >> off;
    if (off + len > 64) x |= b->w[wi + 1] << (64 - off);
    if (len < 64) x &= ((1ull << len) - 1);
    return x;
}
static void obv_set_int(obv *b, uint64_t pos, uint64_t x, unsigned len) {
    for (unsigned t = 0; t < len; ++t) obv_set(b, pos + t, (int)((x >> t) & 1));
}
/* Rank support. Semantics of sdsl::rank_support_v5<1>::rank (rank_support_v5.hpp:116-134):
 * number of set bits in [0, idx), 0 <= idx <= size(). The sampling structure is our own. */
static int obv_build_rank(obv *b) {
    uint64_t nwords = (b->nbits >> 6) + 2;
    uint64_t nblk = nwords / 8 + 2;
    b->cum = (uint64_t *)calloc(nblk, 8);
    if (!b->cum) return -1;
    uint64_t run = 0;
    for (uint64_t wi = 0; wi < nwords; ++wi) {
        if ((wi & 7) == 0) b->cum[wi >> 3] = run;
        run += (uint64_t)popcnt64(b->w[wi]);
    }
    return 0;
}
static uint64_t obv_rank(const obv *b, uint64_t idx) {
    uint64_t wi = idx >> 6;
    uint64_t r = b->cum[wi >> 3];
    for (uint64_t t = (wi & ~7ull); t < wi; ++t) r += (uint64_t)popcnt64(b->w[t]);
    unsigned off = (unsigned)(idx & 63);
    if (off) r += (uint64_t)popcnt64(b->w[wi] & ((1ull << off) - 1));
    return r;
}

/* sdsl::int_vector<0>: u64 bit length, u8 width, packed little-endian values. */
typedef struct {
    obv bits;
    unsigned width;
    uint64_t n; /* number of elements */
} oiv;
static uint64_t oiv_get(const oiv *v, uint64_t i) { return obv_get_int(&v->bits, i * v->width, v->width); }
static void oiv_set(oiv *v, uint64_t i, uint64_t x) { obv_set_int(&v->bits, i * v->width, x, v->width); }
static int oiv_alloc(oiv *v, uint64_t n, unsigned width) {
    v->n = n;
    v->width = width;
    return obv_alloc(&v->bits, n * width);
}

/* ------------------------------------------------------------------------------------------ */
/* RRR<63> (sdsl-lite/include/sdsl/rrr_vector.hpp, rrr_helper.hpp), t_bs = 63, t_k = 32.        */
enum { RRR_BS = 63, RRR_K = 32 };
static uint64_t g_binom[65][65];
static unsigned g_space[64];
static int g_binom_ready = 0;
static unsigned hi_bit(uint64_t x) { return 63u - (unsigned)__builtin_clzll(x); } /* bits::hi, x>0 */

/* rrr_helper.hpp:174-214 (binomial_table) and :240-256 (space[]). */
static void binom_init(void) {
    if (g_binom_ready) return;
    for (int k = 0; k <= 64; ++k) g_binom[k][k] = 1;
    for (int k = 0; k <= 64; ++k) g_binom[0][k] = 0;
    for (int nn = 0; nn <= 64; ++nn) g_binom[nn][0] = 1;
    for (int nn = 1; nn <= 64; ++nn)
        for (int k = 1; k <= 64; ++k) g_binom[nn][k] = g_binom[nn - 1][k - 1] + g_binom[nn - 1][k];
    for (int k = 0; k <= RRR_BS; ++k)
        g_space[k] = (g_binom[RRR_BS][k] == 1) ? 0 : hi_bit(g_binom[RRR_BS][k]) + 1;
    g_binom_ready = 1;
}

typedef struct {
    uint64_t size;
    oiv bt;      /* width 6 */
    obv btnr;
    oiv btnrp;
    oiv rank;
    obv invert;
} orrr;

static void orrr_free(orrr *r) {
    obv_free(&r->bt.bits);
    obv_free(&r->btnr);
    obv_free(&r->btnrp.bits);
    obv_free(&r->rank.bits);
    obv_free(&r->invert);
}

/* rrr_helper.hpp:304-320 bin_to_nr: offset of a 63-bit word inside its popcount class. */
static uint64_t rrr_bin_to_nr(uint64_t bin) {
    if (bin == 0 || bin == ((1ull << RRR_BS) - 1)) return 0;
    uint64_t nr = 0;
    unsigned k = (unsigned)popcnt64(bin), nn = RRR_BS;
    while (bin != 0) {
        if (bin & 1ull) {
            nr += g_binom[nn - 1][k];
            --k;
        }
        bin >>= 1;
        --nn;
    }
    return nr;
}
/* rrr_helper.hpp:323-371 decode_bit (linear branch; the binary-search branch is an equivalent
 * optimisation). */
static int rrr_decode_bit(unsigned k, uint64_t nr, unsigned off) {
    if (k == RRR_BS) return 1;
    if (k == 0) return 0;
    if (k == 1) return (RRR_BS - nr - 1) == off;
    unsigned nn = RRR_BS;
    unsigned i = 0;
    while (k > 1) {
        if (i > off) return 0;
        if (nr >= g_binom[nn - 1][k]) {
            nr -= g_binom[nn - 1][k];
            --k;
            if (i == off) return 1;
        }
        --nn;
        ++i;
    }
    return (RRR_BS - nr - 1) == off;
}
/* rrr_helper.hpp:411-460 decode_popcount: ones among the first off bits of the block. */
static unsigned rrr_decode_popcount(unsigned k, uint64_t nr, unsigned off) {
    if (k == RRR_BS) return off;
    if (k == 0) return 0;
    if (k == 1) return (RRR_BS - nr - 1) < off;
    unsigned result = 0, nn = RRR_BS, i = 0;
    while (k > 1) {
        if (i >= off) return result;
        if (nr >= g_binom[nn - 1][k]) {
            nr -= g_binom[nn - 1][k];
            --k;
            ++result;
        }
        --nn;
        ++i;
    }
    return result + ((RRR_BS - nr - 1) < off);
}
/* rrr_vector.hpp:257-277 operator[]. */
static int orrr_get(const orrr *r, uint64_t i) {
    uint64_t bt_idx = i / RRR_BS;
    unsigned bt = (unsigned)oiv_get(&r->bt, bt_idx);
    uint64_t sample_pos = bt_idx / RRR_K;
    if (obv_get(&r->invert, sample_pos)) bt = RRR_BS - bt;
    if (bt == 0 || bt == RRR_BS) return bt > 0;
    unsigned off = (unsigned)(i % RRR_BS);
    uint64_t btnrp = oiv_get(&r->btnrp, sample_pos);
    for (uint64_t j = sample_pos * RRR_K; j < bt_idx; ++j) btnrp += g_space[oiv_get(&r->bt, j)];
    unsigned btnrlen = g_space[bt];
    uint64_t btnr = obv_get_int(&r->btnr, btnrp, btnrlen);
    return rrr_decode_bit(bt, btnr, off);
}
/* rrr_vector.hpp:445-482 rank_support_rrr<1,63>::rank. */
static uint64_t orrr_rank(const orrr *r, uint64_t i) {
    uint64_t bt_idx = i / RRR_BS;
    uint64_t sample_pos = bt_idx / RRR_K;
    uint64_t btnrp = oiv_get(&r->btnrp, sample_pos);
    uint64_t rank = oiv_get(&r->rank, sample_pos);
    if (sample_pos + 1 < r->rank.n) {
        uint64_t diff = oiv_get(&r->rank, sample_pos + 1) - rank;
        if (diff == 0) return rank;
        if (diff == (uint64_t)RRR_BS * RRR_K) return rank + i - sample_pos * RRR_K * RRR_BS;
    }
    int inv = obv_get(&r->invert, sample_pos);
    for (uint64_t j = sample_pos * RRR_K; j < bt_idx; ++j) {
        unsigned b = (unsigned)oiv_get(&r->bt, j);
        rank += inv ? RRR_BS - b : b;
        btnrp += g_space[b];
    }
    unsigned off = (unsigned)(i % RRR_BS);
    if (!off) return rank;
    unsigned bt = (unsigned)oiv_get(&r->bt, bt_idx);
    if (inv) bt = RRR_BS - bt;
    uint64_t btnr = obv_get_int(&r->btnr, btnrp, g_space[bt]);
    return rank + rrr_decode_popcount(bt, btnr, off);
}

/* rrr_vector.hpp:150-250 — constructor from a plain bit vector. */
static int orrr_build(orrr *r, const obv *bv) {
    binom_init();
    memset(r, 0, sizeof(*r));
    uint64_t m_size = bv->nbits;
    r->size = m_size;
    uint64_t nblocks = (m_size + RRR_BS) / RRR_BS;
    if (oiv_alloc(&r->bt, nblocks, hi_bit(RRR_BS) + 1)) return -1;
    uint64_t pos = 0, i = 0, x, btnr_pos = 0, sum_rank = 0;
    while (pos + RRR_BS <= m_size) {
        x = (uint64_t)popcnt64(obv_get_int(bv, pos, RRR_BS));
        oiv_set(&r->bt, i++, x);
        sum_rank += x;
        btnr_pos += g_space[x];
        pos += RRR_BS;
    }
    if (pos < m_size) {
        x = (uint64_t)popcnt64(obv_get_int(bv, pos, (unsigned)(m_size - pos)));
        oiv_set(&r->bt, i++, x);
        sum_rank += x;
        btnr_pos += g_space[x];
    }
    uint64_t nsb = (nblocks + RRR_K - 1) / RRR_K;
    if (obv_alloc(&r->btnr, btnr_pos > 64 ? btnr_pos : 64)) return -1;
    /* bits::hi(0) == 0 in sdsl (bits.hpp), hence the x ? hi : 0 guards. */
    if (oiv_alloc(&r->btnrp, nsb, (btnr_pos ? hi_bit(btnr_pos) : 0) + 1)) return -1;
    if (oiv_alloc(&r->rank, nsb + ((m_size % ((uint64_t)RRR_K * RRR_BS)) > 0),
                  (sum_rank ? hi_bit(sum_rank) : 0) + 1))
        return -1;
    if (obv_alloc(&r->invert, nsb)) return -1;

    pos = 0;
    i = 0;
    btnr_pos = 0;
    sum_rank = 0;
    int invert = 0;
    while (pos + RRR_BS <= m_size) {
        if ((i % RRR_K) == 0) {
            oiv_set(&r->btnrp, i / RRR_K, btnr_pos);
            oiv_set(&r->rank, i / RRR_K, sum_rank);
            if (i + RRR_K <= nblocks) {
                uint64_t gt_half = 0;
                for (uint64_t j = i; j < i + RRR_K; ++j)
                    if (oiv_get(&r->bt, j) > RRR_BS / 2) ++gt_half;
                if (gt_half > (RRR_K / 2)) {
                    obv_set(&r->invert, i / RRR_K, 1);
                    for (uint64_t j = i; j < i + RRR_K; ++j)
                        oiv_set(&r->bt, j, RRR_BS - oiv_get(&r->bt, j));
                    invert = 1;
                } else {
                    invert = 0;
                }
            } else {
                invert = 0;
            }
        }
        x = oiv_get(&r->bt, i++);
        unsigned sp = g_space[x];
        sum_rank += invert ? (RRR_BS - x) : x;
        if (sp) {
            uint64_t bin = obv_get_int(bv, pos, RRR_BS);
            obv_set_int(&r->btnr, btnr_pos, rrr_bin_to_nr(bin), sp);
        }
        btnr_pos += sp;
        pos += RRR_BS;
    }
    if (pos < m_size) {
        if ((i % RRR_K) == 0) {
            oiv_set(&r->btnrp, i / RRR_K, btnr_pos);
            oiv_set(&r->rank, i / RRR_K, sum_rank);
            obv_set(&r->invert, i / RRR_K, 0);
            invert = 0;
        }
        x = oiv_get(&r->bt, i++);
        unsigned sp = g_space[x];
        sum_rank += invert ? (RRR_BS - x) : x;
        if (sp) {
            uint64_t bin = obv_get_int(bv, pos, (unsigned)(m_size - pos));
            obv_set_int(&r->btnr, btnr_pos, rrr_bin_to_nr(bin), sp);
        }
        btnr_pos += sp;
    }
    oiv_set(&r->rank, r->rank.n - 1, sum_rank);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* On-disk forms: int_vector.hpp:593-610 (header), :1563-1595 (data); rrr_vector.hpp:350-373.   */
typedef struct {
    const uint8_t *p;
    size_t n, pos;
    int err;
} rd;
static uint64_t rd_u64(rd *r) {
    if (r->pos + 8 > r->n) {
        r->err = 1;
        return 0;
    }
    uint64_t v;
    memcpy(&v, r->p + r->pos, 8);
    r->pos += 8;
    return v;
}
static unsigned rd_u8(rd *r) {
    if (r->pos + 1 > r->n) {
        r->err = 1;
        return 0;
    }
    return r->p[r->pos++];
}
static int rd_bv(rd *r, obv *b) {
    uint64_t nbits = rd_u64(r);
    if (r->err) return -1;
    uint64_t nwords = (nbits + 63) >> 6;
    if (r->pos + nwords * 8 > r->n) {
        r->err = 1;
        return -1;
    }
    if (obv_alloc(b, nbits)) return -1;
    memcpy(b->w, r->p + r->pos, nwords * 8);
    r->pos += nwords * 8;
    return 0;
}
static int rd_iv(rd *r, oiv *v) {
    uint64_t nbits = rd_u64(r);
    unsigned width = rd_u8(r);
    if (r->err || width == 0 || width > 64) {
        r->err = 1;
        return -1;
    }
    uint64_t nwords = (nbits + 63) >> 6;
    if (r->pos + nwords * 8 > r->n) {
        r->err = 1;
        return -1;
    }
    v->width = width;
    v->n = nbits / width;
    if (obv_alloc(&v->bits, nbits)) return -1;
    memcpy(v->bits.w, r->p + r->pos, nwords * 8);
    r->pos += nwords * 8;
    return 0;
}
static uint8_t *read_file(const char *path, size_t *n) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *buf = (uint8_t *)malloc((size_t)sz + 1);
    if (buf && fread(buf, 1, (size_t)sz, f) != (size_t)sz) {
        free(buf);
        buf = NULL;
    }
    fclose(f);
    *n = (size_t)sz;
    return buf;
}
static int load_bv_file(const char *path, obv *b) {
    size_t n;
    uint8_t *buf = read_file(path, &n);
    if (!buf) return -1;
    rd r = {buf, n, 0, 0};
    int rc = rd_bv(&r, b);
    free(buf);
    return rc;
}

/* serialisers (for the encoder pin test) */
typedef struct {
    uint8_t *p;
    size_t n, cap;
} wr;
static void wr_bytes(wr *w, const void *src, size_t n) {
    if (w->n + n > w->cap) {
        w->cap = (w->n + n) * 2 + 64;
        w->p = (uint8_t *)realloc(w->p, w->cap);
    }
    memcpy(w->p + w->n, src, n);
    w->n += n;
}
static void wr_bv(wr *w, const obv *b) {
    wr_bytes(w, &b->nbits, 8);
    wr_bytes(w, b->w, ((b->nbits + 63) >> 6) * 8);
}
static void wr_iv(wr *w, const oiv *v) {
    uint8_t width = (uint8_t)v->width;
    wr_bytes(w, &v->bits.nbits, 8);
    wr_bytes(w, &width, 1);
    wr_bytes(w, v->bits.w, ((v->bits.nbits + 63) >> 6) * 8);
}

/* ------------------------------------------------------------------------------------------ */
/* src/fms_index.h:18-49 strand_predictor. */
typedef struct {
    int score;
    int result_scores[2];
    int previous;
} opred;
static int clipped(int x) { return x < -7 ? -7 : (x > 7 ? 7 : x); }
static void pred_reset(opred *p) {
    p->score = 0;
    p->result_scores[0] = p->result_scores[1] = 0;
    p->previous = -1;
}
static void pred_log_result(opred *p, int f, int b) {
    int d = f - b;
    p->score = clipped(p->score + d);
    if (p->previous != -1) p->result_scores[p->previous] = clipped(p->result_scores[p->previous] + d);
    p->previous = f > b;
}
static int pred_predict_swap(const opred *p) {
    if (p->previous != -1 && p->result_scores[p->previous] != 0) return p->result_scores[p->previous] < 0;
    return p->score < 0;
}

/* src/fms_index.h:52-66 struct fms_index. */
struct fmsi_oracle_index {
    obv ac_gt, ac, gt, klcp;
    orrr mask;
    uint64_t counts[4];
    uint64_t dollar_position;
    int k;
    int has_klcp;
    opred predictor;
    fmsi_oracle_counters ctr;
};

static int finish_index(fmsi_oracle_index *x) {
    if (obv_build_rank(&x->ac_gt) || obv_build_rank(&x->ac) || obv_build_rank(&x->gt)) return -1;
    pred_reset(&x->predictor);
    memset(&x->ctr, 0, sizeof(x->ctr));
    return 0;
}

fmsi_oracle_index *fmsi_oracle_load(const char *prefix, int use_klcp) {
    binom_init();
    fmsi_oracle_index *x = (fmsi_oracle_index *)calloc(1, sizeof(*x));
    if (!x) return NULL;
    size_t pl = strlen(prefix);
    char *path = (char *)malloc(pl + 32);
    int ok = 1;
    sprintf(path, "%s.fmsi.ac_gt", prefix);
    ok = ok && load_bv_file(path, &x->ac_gt) == 0;
    sprintf(path, "%s.fmsi.ac", prefix);
    ok = ok && load_bv_file(path, &x->ac) == 0;
    sprintf(path, "%s.fmsi.gt", prefix);
    ok = ok && load_bv_file(path, &x->gt) == 0;
    if (ok) {
        sprintf(path, "%s.fmsi.mask", prefix);
        size_t n;
        uint8_t *buf = read_file(path, &n);
        if (!buf) ok = 0;
        else {
            rd r = {buf, n, 0, 0};
            x->mask.size = rd_u64(&r);
            ok = !r.err && rd_iv(&r, &x->mask.bt) == 0 && rd_bv(&r, &x->mask.btnr) == 0 &&
                 rd_iv(&r, &x->mask.btnrp) == 0 && rd_iv(&r, &x->mask.rank) == 0 &&
                 rd_bv(&r, &x->mask.invert) == 0 && r.pos == n;
            free(buf);
        }
    }
    if (ok && use_klcp) {
        sprintf(path, "%s.fmsi.klcp", prefix);
        FILE *f = fopen(path, "rb");
        if (f) {
            fclose(f);
            ok = load_bv_file(path, &x->klcp) == 0;
            x->has_klcp = ok && x->klcp.nbits > 0;
        }
    }
    if (ok) {
        sprintf(path, "%s.fmsi.misc", prefix);
        FILE *f = fopen(path, "r");
        if (!f) ok = 0;
        else {
            unsigned long long d, c[4];
            int k;
            if (fscanf(f, "%llu %llu %llu %llu %llu %d", &d, &c[0], &c[1], &c[2], &c[3], &k) != 6) ok = 0;
            else {
                x->dollar_position = d;
                for (int t = 0; t < 4; ++t) x->counts[t] = c[t];
                x->k = k;
            }
            fclose(f);
        }
    }
    free(path);
    if (!ok || finish_index(x)) {
        fmsi_oracle_free(x);
        return NULL;
    }
    return x;
}

static int bits_to_obv(obv *b, const uint8_t *bits, size_t n) {
    if (obv_alloc(b, n)) return -1;
    for (size_t i = 0; i < n; ++i)
        if (bits[i]) obv_set(b, i, 1);
    return 0;
}

fmsi_oracle_index *fmsi_oracle_from_bits(const uint8_t *ac_gt, size_t n_ac_gt, const uint8_t *ac,
                                         size_t n_ac, const uint8_t *gt, size_t n_gt,
                                         const uint8_t *mask, size_t n_mask,
                                         const uint64_t counts[4], uint64_t dollar_position,
                                         const uint8_t *klcp, size_t n_klcp, int k) {
    binom_init();
    fmsi_oracle_index *x = (fmsi_oracle_index *)calloc(1, sizeof(*x));
    if (!x) return NULL;
    obv m;
    int ok = bits_to_obv(&x->ac_gt, ac_gt, n_ac_gt) == 0 && bits_to_obv(&x->ac, ac, n_ac) == 0 &&
             bits_to_obv(&x->gt, gt, n_gt) == 0 && bits_to_obv(&m, mask, n_mask) == 0;
    if (ok) {
        ok = orrr_build(&x->mask, &m) == 0;
        obv_free(&m);
    }
    if (ok && klcp && n_klcp) {
        ok = bits_to_obv(&x->klcp, klcp, n_klcp) == 0;
        x->has_klcp = 1;
    }
    for (int t = 0; t < 4; ++t) x->counts[t] = counts[t];
    x->dollar_position = dollar_position;
    x->k = k;
    if (!ok || finish_index(x)) {
        fmsi_oracle_free(x);
        return NULL;
    }
    return x;
}

void fmsi_oracle_free(fmsi_oracle_index *x) {
    if (!x) return;
    obv_free(&x->ac_gt);
    obv_free(&x->ac);
    obv_free(&x->gt);
    obv_free(&x->klcp);
    orrr_free(&x->mask);
    free(x);
}

uint64_t fmsi_oracle_size(const fmsi_oracle_index *x) { return x->mask.size; }
int fmsi_oracle_k(const fmsi_oracle_index *x) { return x->k; }
int fmsi_oracle_has_klcp(const fmsi_oracle_index *x) { return x->has_klcp; }
uint64_t fmsi_oracle_count(const fmsi_oracle_index *x, int c) { return x->counts[c & 3]; }
uint64_t fmsi_oracle_dollar(const fmsi_oracle_index *x) { return x->dollar_position; }
void fmsi_oracle_reset_predictor(fmsi_oracle_index *x) { pred_reset(&x->predictor); }
void fmsi_oracle_counters_reset(fmsi_oracle_index *x) { memset(&x->ctr, 0, sizeof(x->ctr)); }
void fmsi_oracle_counters_get(const fmsi_oracle_index *x, fmsi_oracle_counters *o) { *o = x->ctr; }
int fmsi_oracle_mask_bit(const fmsi_oracle_index *x, uint64_t i) { return orrr_get(&x->mask, i); }
uint64_t fmsi_oracle_mask_rank(const fmsi_oracle_index *x, uint64_t i) { return orrr_rank(&x->mask, i); }
int fmsi_oracle_klcp_bit(const fmsi_oracle_index *x, uint64_t i) { return obv_get(&x->klcp, i); }

/* ------------------------------------------------------------------------------------------ */
/* src/fms_index.h:68-86 rank(index, i, c). */
uint64_t fmsi_oracle_rank(const fmsi_oracle_index *x, uint64_t i, int c) {
    uint64_t gt_position = obv_rank(&x->ac_gt, i);
    if (c >= 2) {
        uint64_t t_position = obv_rank(&x->gt, gt_position);
        return c == 2 ? gt_position - t_position : t_position;
    }
    uint64_t c_position = obv_rank(&x->ac, i - gt_position);
    if (c == 0) return i - gt_position - c_position - (i >= x->dollar_position + 1);
    return c_position;
}
/* src/fms_index.h:88-95 access(index, i). */
int fmsi_oracle_access(const fmsi_oracle_index *x, uint64_t i) {
    uint64_t gt_position = obv_rank(&x->ac_gt, i);
    if (obv_get(&x->ac_gt, i)) return 2 + obv_get(&x->gt, gt_position);
    return obv_get(&x->ac, i - gt_position);
}
/* src/fms_index.h:98-103 update_range. */
void fmsi_oracle_update_range(fmsi_oracle_index *x, uint64_t *i, uint64_t *j, int c) {
    if (*j == *i) return;
    x->ctr.lf_steps++;
    x->ctr.rank_sectors += 1 + ((*i >> 6) != (*j >> 6));
    uint64_t count = x->counts[c];
    *i = count + fmsi_oracle_rank(x, *i, c);
    *j = count + fmsi_oracle_rank(x, *j, c);
}
/* src/fms_index.h:106-109 extend_range_with_klcp. */
void fmsi_oracle_extend_range_with_klcp(fmsi_oracle_index *x, uint64_t *i, uint64_t *j) {
    x->ctr.klcp_steps++;
    while (obv_get(&x->klcp, *j - 1)) (*j)++;
    while (obv_get(&x->klcp, *i - 1)) (*i)--;
}
/* src/fms_index.h:117-124 get_range_with_pattern. */
void fmsi_oracle_get_range_with_pattern(fmsi_oracle_index *x, uint64_t *sa_start, uint64_t *sa_end,
                                        const char *pattern, int k) {
    *sa_start = 0;
    *sa_end = x->mask.size;
    x->ctr.strand_searches++;
    for (int i = k - 1; i >= 0 && *sa_start != *sa_end; --i)
        fmsi_oracle_update_range(x, sa_start, sa_end, nucleotide_to_int((unsigned char)pattern[i]));
}
/* src/fms_index.h:126-144 infer_presence<maximized_ones>. */
int fmsi_oracle_infer_presence(fmsi_oracle_index *x, uint64_t sa_start, uint64_t sa_end, int max_ones) {
    if (max_ones) {
        if (sa_start != sa_end) {
            x->ctr.mask_sectors += 1;
            return orrr_get(&x->mask, sa_start);
        }
        return -1;
    }
    if (sa_start != sa_end) x->ctr.mask_sectors += 1 + ((sa_start >> 6) != (sa_end >> 6));
    for (uint64_t i = sa_start; i < sa_end; ++i)
        if (orrr_get(&x->mask, i)) return 1;
    return sa_start == sa_end ? -1 : 0;
}
/* src/fms_index.h:146-156 kmer_order / kmer_order_if_present. */
int64_t fmsi_oracle_kmer_order_if_present(fmsi_oracle_index *x, uint64_t sa_start, uint64_t sa_end) {
    if (fmsi_oracle_infer_presence(x, sa_start, sa_end, 0) == 1) return (int64_t)orrr_rank(&x->mask, sa_start);
    return -1;
}
/* src/fms_index.h:158-169 single_query_or<> / single_query_order. */
static int single_query_or(fmsi_oracle_index *x, const char *pattern, int k, int max_ones) {
    uint64_t s, e;
    fmsi_oracle_get_range_with_pattern(x, &s, &e, pattern, k);
    return fmsi_oracle_infer_presence(x, s, e, max_ones);
}
static int64_t single_query_order(fmsi_oracle_index *x, const char *pattern, int k) {
    uint64_t s, e;
    fmsi_oracle_get_range_with_pattern(x, &s, &e, pattern, k);
    return fmsi_oracle_kmer_order_if_present(x, s, e);
}

/* ------------------------------------------------------------------------------------------ */
void fmsi_oracle_buf_free(fmsi_oracle_buf *b) {
    free(b->s);
    b->s = NULL;
    b->len = b->cap = 0;
}
static void buf_put(fmsi_oracle_buf *b, const char *s, size_t n) {
    if (b->len + n + 1 > b->cap) {
        b->cap = (b->len + n + 1) * 2 + 256;
        b->s = (char *)realloc(b->s, b->cap);
    }
    memcpy(b->s + b->len, s, n);
    b->len += n;
    b->s[b->len] = 0;
}
static void buf_putc(fmsi_oracle_buf *b, char c) { buf_put(b, &c, 1); }
static void buf_puti(fmsi_oracle_buf *b, long long v) {
    char t[32];
    int n = snprintf(t, sizeof t, "%lld", v);
    buf_put(b, t, (size_t)n);
}

/* src/kmers.h:53-59 ReverseComplementString. */
static char *reverse_complement_string(const char *s, size_t len) {
    char *r = (char *)malloc(len + 1);
    for (size_t i = 0; i < len; ++i) r[i] = complementary_nucleotide((unsigned char)s[len - i - 1]);
    r[len] = 0;
    return r;
}

/* src/fms_index.h:181-254 query_kmers_streaming<maximized_ones>. */
static void query_kmers_streaming(fmsi_oracle_index *x, int max_ones, const char *sequence,
                                  const char *rc_sequence, size_t sequence_length, int k,
                                  int output_orders, fmsi_oracle_buf *of) {
    size_t nres = sequence_length - (size_t)k + 1;
    int64_t *result = (int64_t *)malloc(nres * sizeof(int64_t));
    for (size_t i = 0; i < nres; ++i) result[i] = -1;
    int should_swap = pred_predict_swap(&x->predictor);
    if (should_swap) {
        const char *t = sequence;
        sequence = rc_sequence;
        rc_sequence = t;
    }
    int fwd_pred = 0, bwd_pred = 0;
    uint64_t sa_start = (uint64_t)-1, sa_end = (uint64_t)-1;
    for (size_t i = 0; i + (size_t)k <= sequence_length; ++i) {
        size_t i_back = sequence_length - (size_t)k - i;
        if (sa_start == sa_end) {
            fmsi_oracle_get_range_with_pattern(x, &sa_start, &sa_end, sequence + i_back, k);
        } else {
            fmsi_oracle_extend_range_with_klcp(x, &sa_start, &sa_end);
            fmsi_oracle_update_range(x, &sa_start, &sa_end, nucleotide_to_int((unsigned char)sequence[i_back]));
        }
        if (output_orders) {
            result[i_back] = fmsi_oracle_kmer_order_if_present(x, sa_start, sa_end);
            if (result[i_back] >= 0) fwd_pred++;
            else fwd_pred--;
        } else {
            result[i_back] = fmsi_oracle_infer_presence(x, sa_start, sa_end, max_ones);
            fwd_pred += (int)result[i_back];
        }
    }
    sa_start = sa_end = (uint64_t)-1;
    for (size_t i = 0; i + (size_t)k <= sequence_length; ++i) {
        if ((result[i] >= 0 && output_orders) || result[i] == 1 || (result[i] == 0 && max_ones)) {
            sa_start = sa_end = (uint64_t)-1;
            continue;
        }
        size_t i_back = sequence_length - (size_t)k - i;
        if (sa_start == sa_end) {
            fmsi_oracle_get_range_with_pattern(x, &sa_start, &sa_end, rc_sequence + i_back, k);
        } else {
            fmsi_oracle_extend_range_with_klcp(x, &sa_start, &sa_end);
            fmsi_oracle_update_range(x, &sa_start, &sa_end, nucleotide_to_int((unsigned char)rc_sequence[i_back]));
        }
        int64_t res;
        if (output_orders) {
            res = fmsi_oracle_kmer_order_if_present(x, sa_start, sa_end);
            if (res >= 0) bwd_pred++;
            else bwd_pred--;
        } else {
            res = fmsi_oracle_infer_presence(x, sa_start, sa_end, max_ones);
            bwd_pred += (int)res;
        }
        if (res > result[i]) result[i] = res;
    }
    if (should_swap) {
        for (size_t a = 0, b = nres - 1; a < b; ++a, --b) {
            int64_t t = result[a];
            result[a] = result[b];
            result[b] = t;
        }
        int t = fwd_pred;
        fwd_pred = bwd_pred;
        bwd_pred = t;
    }
    pred_log_result(&x->predictor, fwd_pred, bwd_pred);
    for (size_t i = 0; i < nres; ++i) {
        if (output_orders) {
            if (i > 0) buf_putc(of, ',');
            buf_puti(of, result[i]);
        } else {
            buf_putc(of, result[i] == 1 ? '1' : '0');
        }
    }
    x->ctr.kmers += nres;
    free(result);
}

/* src/fms_index.h:263-331 query_kmers_single<mode> (orr / all; general is out of scope). */
static void query_kmers_single(fmsi_oracle_index *x, int mode, const char *sequence,
                               const char *rc_sequence, size_t sequence_length, int k,
                               fmsi_oracle_buf *of, int output_orders) {
    for (size_t i = 0; i + (size_t)k <= sequence_length; ++i) {
        const char *kmer = sequence + i;
        const char *rc_kmer = rc_sequence + (sequence_length - (size_t)k - i);
        int should_swap = pred_predict_swap(&x->predictor);
        if (should_swap) {
            const char *t = kmer;
            kmer = rc_kmer;
            rc_kmer = t;
        }
        int fwd_pred = 0, bwd_pred = 0;
        int64_t got;
        if (output_orders) got = single_query_order(x, kmer, k);
        else got = single_query_or(x, kmer, k, mode == FMSI_ORACLE_MODE_ALL);
        fwd_pred = (int)got; /* int = int64 truncation as in the reference (:282) */
        if (output_orders) {
            if (fwd_pred >= 0) fwd_pred = 1;
            else {
                got = single_query_order(x, rc_kmer, k);
                bwd_pred = got >= 0 ? 1 : -1;
            }
        } else if (mode == FMSI_ORACLE_MODE_OR) {
            if (got != 1) {
                got = single_query_or(x, rc_kmer, k, 0);
                bwd_pred = (int)got;
            }
        } else {
            if (got == -1) {
                got = single_query_or(x, rc_kmer, k, 1);
                bwd_pred = (int)got;
            }
        }
        if (output_orders) {
            if (i > 0) buf_putc(of, ',');
            buf_puti(of, got);
        } else {
            buf_putc(of, got == 1 ? '1' : '0');
        }
        if (should_swap) {
            int t = fwd_pred;
            fwd_pred = bwd_pred;
            bwd_pred = t;
        }
        pred_log_result(&x->predictor, fwd_pred, bwd_pred);
        x->ctr.kmers++;
    }
}

/* src/fms_index.h:333-342 query_kmers<mode>. */
void fmsi_oracle_query_kmers(fmsi_oracle_index *x, int mode, const char *seq, size_t len, int k,
                             int has_klcp, int output_orders, fmsi_oracle_buf *out) {
    char *rc = reverse_complement_string(seq, len);
    if (has_klcp) query_kmers_streaming(x, mode == FMSI_ORACLE_MODE_ALL, seq, rc, len, k, output_orders, out);
    else query_kmers_single(x, mode, seq, rc, len, k, out, output_orders);
    free(rc);
}

/* ------------------------------------------------------------------------------------------ */
/* src/kseq.h:173-226 kseq_read over a flat memory buffer (the 16 KiB refill logic of
 * ks_getc/ks_getuntil2, :68-151, is transparent to the result). */
typedef struct {
    char *s;
    size_t l, m;
} kstr;
typedef struct {
    const char *p;
    size_t n, pos;
    int last_char;
    kstr name, comment, seq, qual;
} kseq_mem;
enum { SEP_SPACE = 0, SEP_LINE = 2 };

static int ksm_getc(kseq_mem *ks) {
    if (ks->pos >= ks->n) return -1;
    return (unsigned char)ks->p[ks->pos++];
}
static void kstr_reserve(kstr *s, size_t need) {
    if (s->m < need) {
        s->m = need * 2 + 16;
        s->s = (char *)realloc(s->s, s->m);
    }
}
static int64_t ksm_getuntil2(kseq_mem *ks, int delimiter, kstr *str, int *dret, int append) {
    if (dret) *dret = 0;
    if (!append) str->l = 0;
    if (ks->pos >= ks->n) {
        /* !gotany && eof */
        kstr_reserve(str, str->l + 1);
        str->s[str->l] = 0;
        return -1;
    }
    size_t i = ks->pos;
    if (delimiter == SEP_LINE) {
        const char *sep = (const char *)memchr(ks->p + ks->pos, '\n', ks->n - ks->pos);
        i = sep ? (size_t)(sep - ks->p) : ks->n;
    } else {
        for (; i < ks->n; ++i)
            if (isspace((unsigned char)ks->p[i])) break;
    }
    kstr_reserve(str, str->l + (i - ks->pos) + 2);
    memcpy(str->s + str->l, ks->p + ks->pos, i - ks->pos);
    str->l += i - ks->pos;
    if (i < ks->n) {
        if (dret) *dret = (unsigned char)ks->p[i];
        ks->pos = i + 1;
    } else {
        ks->pos = ks->n;
    }
    if (delimiter == SEP_LINE && str->l > 1 && str->s[str->l - 1] == '\r') --str->l;
    str->s[str->l] = 0;
    return (int64_t)str->l;
}
static int64_t ksm_read(kseq_mem *ks) {
    int c;
    int64_t r;
    if (ks->last_char == 0) {
        while ((c = ksm_getc(ks)) >= 0 && c != '>' && c != '@') {}
        if (c < 0) return c;
        ks->last_char = c;
    }
    ks->comment.l = ks->seq.l = ks->qual.l = 0;
    if ((r = ksm_getuntil2(ks, SEP_SPACE, &ks->name, &c, 0)) < 0) return r;
    if (c != '\n') ksm_getuntil2(ks, SEP_LINE, &ks->comment, 0, 0);
    kstr_reserve(&ks->seq, 256);
    while ((c = ksm_getc(ks)) >= 0 && c != '>' && c != '+' && c != '@') {
        if (c == '\n') continue;
        kstr_reserve(&ks->seq, ks->seq.l + 2);
        ks->seq.s[ks->seq.l++] = (char)c;
        ksm_getuntil2(ks, SEP_LINE, &ks->seq, 0, 1);
    }
    if (c == '>' || c == '@') ks->last_char = c;
    kstr_reserve(&ks->seq, ks->seq.l + 2);
    ks->seq.s[ks->seq.l] = 0;
    if (c != '+') return (int64_t)ks->seq.l;
    while ((c = ksm_getc(ks)) >= 0 && c != '\n') {}
    if (c == -1) return -2;
    while (ksm_getuntil2(ks, SEP_LINE, &ks->qual, 0, 1) >= 0 && ks->qual.l < ks->seq.l) {}
    ks->last_char = 0;
    if (ks->seq.l != ks->qual.l) return -2;
    return (int64_t)ks->seq.l;
}

/* src/parser.h:57-64 next_invalid_character_or_end. */
static size_t next_invalid_character_or_end(const char *s, size_t length) {
    for (size_t i = 0; i < length; ++i)
        if (nucleotide_to_int((unsigned char)s[i]) == 4) return i;
    return length;
}

/* src/main.cpp:328-373 — the record loop of ms_query. */
int64_t fmsi_oracle_ms_query(fmsi_oracle_index *x, const char *text, size_t text_len, int k, int mode,
                             int has_klcp, int output_orders, fmsi_oracle_buf *out) {
    kseq_mem ks;
    memset(&ks, 0, sizeof ks);
    ks.p = text;
    ks.n = text_len;
    int64_t sequence_length, nrec = 0;
    while ((sequence_length = ksm_read(&ks)) >= 0) {
        ++nrec;
        int64_t max_chunk = 400;
        int64_t sq = 2 * (int64_t)sqrt((double)sequence_length);
        if (sq < max_chunk) max_chunk = sq;
        if (max_chunk < 10) max_chunk = 10;
        max_chunk += k;
        buf_put(out, ks.name.s ? ks.name.s : "", ks.name.l);
        buf_putc(out, '\t');
        const char *sequence = ks.seq.s;
        int output_comma = 0;
        while (sequence_length > 0) {
            int64_t current_length = (int64_t)next_invalid_character_or_end(sequence, (size_t)sequence_length);
            while (current_length >= k) {
                if (output_orders && output_comma) buf_putc(out, ',');
                output_comma = 1;
                int64_t chunk_length = current_length < max_chunk ? current_length : max_chunk;
                fmsi_oracle_query_kmers(x, mode, sequence, (size_t)chunk_length, k, has_klcp, output_orders, out);
                sequence += chunk_length - k + 1;
                current_length -= chunk_length - k + 1;
                sequence_length -= chunk_length - k + 1;
            }
            sequence_length -= current_length + 1;
            sequence += current_length + 1;
            if (sequence_length >= 0) {
                int64_t lim = k < current_length + 1 ? k : current_length + 1;
                for (int64_t i = 0; i < lim; ++i) {
                    if (output_orders) {
                        if (output_comma) buf_putc(out, ',');
                        output_comma = 1;
                        buf_put(out, "-1", 2);
                    } else {
                        buf_putc(out, '0');
                    }
                }
            }
        }
        buf_putc(out, '\n');
    }
    free(ks.name.s);
    free(ks.comment.s);
    free(ks.seq.s);
    free(ks.qual.s);
    return nrec;
}

/* ------------------------------------------------------------------------------------------ */
void fmsi_oracle_kmer_both_strands(fmsi_oracle_index *x, const char *kmer, int k, int mode,
                                   int output_orders, int64_t *fwd, int64_t *rc) {
    char *r = reverse_complement_string(kmer, (size_t)k);
    if (output_orders) {
        *fwd = single_query_order(x, kmer, k);
        *rc = single_query_order(x, r, k);
    } else {
        *fwd = single_query_or(x, kmer, k, mode == FMSI_ORACLE_MODE_ALL);
        *rc = single_query_or(x, r, k, mode == FMSI_ORACLE_MODE_ALL);
    }
    free(r);
}

void fmsi_oracle_query_packed(fmsi_oracle_index *x, int mode, int output_orders, const uint64_t *kmers,
                              size_t n, int k, int64_t *results) {
    static const char L[4] = {'A', 'C', 'G', 'T'};
    char fwd[33], rc[33];
    for (size_t q = 0; q < n; ++q) {
        uint64_t v = kmers[q];
        for (int t = 0; t < k; ++t) {
            int c = (int)((v >> (2 * (k - 1 - t))) & 3);
            fwd[t] = L[c];
            rc[k - 1 - t] = L[3 - c];
        }
        fwd[k] = rc[k] = 0;
        int64_t got;
        /* query_kmers_single with should_swap == false (:272-299) */
        if (output_orders) {
            got = single_query_order(x, fwd, k);
            if (got < 0) got = single_query_order(x, rc, k);
        } else if (mode == FMSI_ORACLE_MODE_OR) {
            got = single_query_or(x, fwd, k, 0);
            if (got != 1) got = single_query_or(x, rc, k, 0);
            got = got == 1;
        } else {
            got = single_query_or(x, fwd, k, 1);
            if (got == -1) got = single_query_or(x, rc, k, 1);
            got = got == 1;
        }
        results[q] = got;
        x->ctr.kmers++;
    }
}

uint8_t *fmsi_oracle_rrr_serialize(const uint8_t *bits, size_t nbits, size_t *out_len) {
    obv bv;
    orrr r;
    if (bits_to_obv(&bv, bits, nbits)) return NULL;
    if (orrr_build(&r, &bv)) {
        obv_free(&bv);
        return NULL;
    }
    wr w = {NULL, 0, 0};
    wr_bytes(&w, &r.size, 8);
    wr_iv(&w, &r.bt);
    wr_bv(&w, &r.btnr);
    wr_iv(&w, &r.btnrp);
    wr_iv(&w, &r.rank);
    wr_bv(&w, &r.invert);
    obv_free(&bv);
    orrr_free(&r);
    *out_len = w.n;
    return w.p;
}

uint8_t *fmsi_oracle_mask_bits(const fmsi_oracle_index *x) {
    uint8_t *o = (uint8_t *)malloc(x->mask.size + 1);
    for (uint64_t i = 0; i < x->mask.size; ++i) o[i] = (uint8_t)orrr_get(&x->mask, i);
    return o;
}
