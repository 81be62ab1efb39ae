/* fmsi_oracle.h — TEST INFRASTRUCTURE ONLY (never linked into or called by the product path).
 *
 * Plain-C restatement of the reference FMSI query path (OndrejSladky/fmsi v0.4.0):
 * src/fms_index.h (rank, update_range, extend_range_with_klcp, get_range_with_pattern,
 * infer_presence, kmer_order_if_present, strand_predictor, query_kmers_single,
 * query_kmers_streaming, load_index), src/main.cpp ms_query (record loop), src/parser.h,
 * src/kseq.h (record reader) and the parts of sdsl-lite 2.1.0 the path touches
 * (int_vector / rrr_vector<63> serialisation, rank semantics, RRR block coding).
 *
 * Parity is PINNED: tests/test_oracle.py checks this file against every golden vector of the
 * reference's own tests for the path (tests/fms_index_test.h, tests/testfiles/) and against
 * outputs of the unmodified reference binary (oracle/_ref/fmsi) captured in tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * use anything declared here.
 */
#ifndef FMSI_ORACLE_H
#define FMSI_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fmsi_oracle_index fmsi_oracle_index;

/* query_mode of src/fms_index.h:256-260 (general / -f functions are out of scope). */
enum { FMSI_ORACLE_MODE_OR = 0, FMSI_ORACLE_MODE_ALL = 1 };

/* load_index, src/fms_index.h:502-526. Returns NULL on I/O or format error. */
fmsi_oracle_index *fmsi_oracle_load(const char *prefix, int use_klcp);

/* In-memory fixture constructor mirroring the hand-built indexes of tests/fms_index_test.h:10-69.
 * Every *_bits array holds one byte (0/1) per bit. klcp may be NULL / n_klcp = 0. */
fmsi_oracle_index *fmsi_oracle_from_bits(const uint8_t *ac_gt, size_t n_ac_gt, const uint8_t *ac,
                                         size_t n_ac, const uint8_t *gt, size_t n_gt,
                                         const uint8_t *mask, size_t n_mask,
                                         const uint64_t counts[4], uint64_t dollar_position,
                                         const uint8_t *klcp, size_t n_klcp, int k);
void fmsi_oracle_free(fmsi_oracle_index *idx);

/* Accessors. */
uint64_t fmsi_oracle_size(const fmsi_oracle_index *idx); /* sa_transformed_mask.size() = n+1 */
int fmsi_oracle_k(const fmsi_oracle_index *idx);
int fmsi_oracle_has_klcp(const fmsi_oracle_index *idx);
uint64_t fmsi_oracle_count(const fmsi_oracle_index *idx, int c);
uint64_t fmsi_oracle_dollar(const fmsi_oracle_index *idx);
void fmsi_oracle_reset_predictor(fmsi_oracle_index *idx);

/* Primitive steps (same names as the reference). */
uint64_t fmsi_oracle_rank(const fmsi_oracle_index *idx, uint64_t i, int c);       /* fms_index.h:68 */
int fmsi_oracle_access(const fmsi_oracle_index *idx, uint64_t i);                 /* fms_index.h:88 */
void fmsi_oracle_update_range(fmsi_oracle_index *idx, uint64_t *i, uint64_t *j, int c);  /* :98 */
void fmsi_oracle_extend_range_with_klcp(fmsi_oracle_index *idx, uint64_t *i, uint64_t *j); /* :106 */
void fmsi_oracle_get_range_with_pattern(fmsi_oracle_index *idx, uint64_t *i, uint64_t *j,
                                        const char *pattern, int k);              /* :117 */
int fmsi_oracle_infer_presence(fmsi_oracle_index *idx, uint64_t i, uint64_t j, int max_ones); /* :126 */
int64_t fmsi_oracle_kmer_order_if_present(fmsi_oracle_index *idx, uint64_t i, uint64_t j);   /* :150 */
int fmsi_oracle_mask_bit(const fmsi_oracle_index *idx, uint64_t i);   /* rrr_vector::operator[] */
uint64_t fmsi_oracle_mask_rank(const fmsi_oracle_index *idx, uint64_t i); /* rank_support_rrr::rank */
int fmsi_oracle_klcp_bit(const fmsi_oracle_index *idx, uint64_t i);

/* Growable output buffer (the reference streams to std::ostream). */
typedef struct {
    char *s;
    size_t len, cap;
} fmsi_oracle_buf;
void fmsi_oracle_buf_free(fmsi_oracle_buf *b);

/* query_kmers<mode>, src/fms_index.h:333-342: one chunk, appends text to out. Mutates the
 * predictor exactly like the reference. */
void fmsi_oracle_query_kmers(fmsi_oracle_index *idx, int mode, const char *seq, size_t len, int k,
                             int has_klcp, int output_orders, fmsi_oracle_buf *out);

/* ms_query record loop, src/main.cpp:328-373, over an in-memory FASTA/FASTQ text (kseq
 * semantics, src/kseq.h:173-226). Appends "name\tresults\n" lines to out. Returns the number
 * of records processed. */
int64_t fmsi_oracle_ms_query(fmsi_oracle_index *idx, const char *text, size_t text_len, int k,
                             int mode, int has_klcp, int output_orders, fmsi_oracle_buf *out);

/* Predictor-free per-strand results for one k-mer (ASCII): what single_query_or<max_ones> /
 * single_query_order return on the k-mer and on its reverse complement. */
void fmsi_oracle_kmer_both_strands(fmsi_oracle_index *idx, const char *kmer, int k, int mode,
                                   int output_orders, int64_t *fwd, int64_t *rc);

/* Batch helper for parity tests: n packed k-mers (2 bits/base, first base in the highest used
 * bits, A=0 C=1 G=2 T=3) -> the value query_kmers_single prints with a NEUTRAL predictor
 * (forward strand first), as int64 (presence: 1/0; orders: id or -1). Does not touch the
 * predictor. */
void fmsi_oracle_query_packed(fmsi_oracle_index *idx, int mode, int output_orders,
                              const uint64_t *kmers, size_t n, int k, int64_t *results);

/* RRR<63> encoder restating rrr_vector's constructor (rrr_vector.hpp:150-250) and serialize
 * (:350-363); used to pin the coder against reference-written .mask files byte for byte.
 * bits: one byte per bit. Returns malloc'd buffer, *out_len bytes. */
uint8_t *fmsi_oracle_rrr_serialize(const uint8_t *bits, size_t nbits, size_t *out_len);
/* Decode every bit of the loaded mask into one byte per bit (caller frees). */
uint8_t *fmsi_oracle_mask_bits(const fmsi_oracle_index *idx);

/* Work counters, SURVEY.md §8(d): executed LF-steps (update_range with i != j), rank sectors
 * (1 + [i>>6 != j>>6] per LF-step), mask probes (one per strand search that ends non-empty;
 * 2 in or/orders mode when i>>?.. straddles — see .c), kLCP-extended steps. */
typedef struct {
    uint64_t lf_steps, rank_sectors, mask_sectors, klcp_steps, strand_searches, kmers;
} fmsi_oracle_counters;
void fmsi_oracle_counters_reset(fmsi_oracle_index *idx);
void fmsi_oracle_counters_get(const fmsi_oracle_index *idx, fmsi_oracle_counters *out);

#ifdef __cplusplus
}
#endif
#endif
