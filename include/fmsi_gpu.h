/* fmsi_gpu.h — C-ABI of libfmsi_gpu.so: the B200 (sm_100a) drop-in for FMSI's query path.
 *
 * The reference (OndrejSladky/fmsi v0.4.0) has no FFI; its query path is reached through the
 * function-level seam `load_index()` + `query_kmers<mode>()` called from `ms_query`
 * (reference src/main.cpp:302, :344-350). Every entry point below names the reference function
 * it replaces (paths relative to the reference root). INTEGRATION.md shows the binding a
 * maintainer adds to the reference's main.cpp.
 *
 * Conventions
 *  - plain C types only; all functions return FMSI_GPU_OK (0) or a negative error code, and
 *    fmsi_gpu_last_error() returns a thread-local message for the last failure;
 *  - k-mers are packed 2 bits per base, A=0 C=1 G=2 T=3 (reference src/kmers.h:3-20), first base
 *    in the highest used bits: kmer = sum base[t] << 2*(k-1-t), k <= 32 (longer k-mers are queried
 *    as text through fmsi_gpu_query_chunks);
 *  - SA intervals are half-open [i, j) over [0, N], N = n+1 = sa_transformed_mask.size();
 *  - `mem` says where the query/result buffers live: FMSI_GPU_MEM_HOST (the library stages them
 *    through its own pinned/device buffers and returns when results are in host memory) or
 *    FMSI_GPU_MEM_DEVICE (device pointers; the call enqueues on `stream` and does not synchronise);
 *  - there is NO CPU fallback: every call fails with FMSI_GPU_ERR_CUDA when no usable device exists.
 */
#ifndef FMSI_GPU_H
#define FMSI_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FMSI_GPU_ABI_VERSION 1

enum {
    FMSI_GPU_OK = 0,
    FMSI_GPU_ERR_ARG = -1,     /* bad argument */
    FMSI_GPU_ERR_IO = -2,      /* index files missing / malformed ("index not correctly loaded") */
    FMSI_GPU_ERR_CUDA = -3,    /* CUDA runtime failure or no device */
    FMSI_GPU_ERR_KLCP = -4,    /* streaming requested but the index has no kLCP array */
    FMSI_GPU_ERR_K = -5,       /* k does not match the index / k unsupported */
    FMSI_GPU_ERR_NOMEM = -6
};

/* query_mode, reference src/fms_index.h:256-260 (`general`: see fmsi_gpu_query_kmers_general). */
enum { FMSI_GPU_MODE_OR = 0, FMSI_GPU_MODE_ALL = 1 };
/* output_orders of query_kmers(): presence bits (`fmsi query`) or mask-rank ids (`fmsi lookup`). */
enum {
    FMSI_GPU_OUT_PRESENCE = 0,
    FMSI_GPU_OUT_ORDERS = 1,
    /* presence, one BIT per k-mer: bit q & 7 of byte q >> 3 is 1 iff the reference prints '1' for k-mer q (the
     * characters of the reference's output line, 8 per byte). STRANDS_LAZY only. results: uint8[(n + 7) / 8].
     * The device packs the bits, so a result crosses PCIe as 1/8 byte per k-mer. */
    FMSI_GPU_OUT_PRESENCE_BITS = 2
};
/* Text of fmsi_gpu_query_chunks*: ASCII ACGTacgt, or 2 bits per base (A=0 C=1 G=2 T=3, src/kmers.h:3-20), 32 bases per
 * 64-bit word, base b in bits [62 - 2 (b & 31), 63 - 2 (b & 31)] of word b >> 5 (first base highest, like a packed
 * k-mer); the unused low bits of the last word are ignored. */
enum { FMSI_GPU_TEXT_ASCII = 0, FMSI_GPU_TEXT_PACKED2 = 1 };
/* Strand policy.
 *  LAZY: forward strand first, reverse complement only if undecided — the reference's evaluation
 *        order with a neutral strand predictor (src/fms_index.h:268-299 with should_swap == false).
 *  BOTH: always evaluate both strands and return both values, so the host can replay the
 *        reference's stateful strand_predictor (src/fms_index.h:18-49) exactly, whatever it says. */
enum { FMSI_GPU_STRANDS_LAZY = 0, FMSI_GPU_STRANDS_BOTH = 1 };
enum { FMSI_GPU_MEM_HOST = 0, FMSI_GPU_MEM_DEVICE = 1 };

typedef struct fmsi_gpu_index fmsi_gpu_index;

typedef struct {
    int32_t prefix_t;      /* depth of the k-mer suffix lookup table; -1 = auto, 0 = none */
    int32_t sb_shift_log2; /* test hook: superblock size (log2 blocks); 0 = auto */
    int32_t dict;          /* k-mer dictionary tier for single-k-mer queries: -1 = auto (2, else 1, else 0 as memory allows; a note on
                            * stderr says so when auto ends below 2, and fmsi_gpu_index_info.dict tells which tier is resident),
                            * 0 = off (backward search), 1 = SA-ordered dictionary (one probe per strand search),
                            * 2 = strand-folded dictionary (one probe per k-mer) */
    int32_t multistep;     /* multi-step rank arrays of the backward-search kernels (m LF-steps per memory request):
                            * -1 = auto (2 when no dictionary tier is resident and they fit), 0 = off, 2 or 3 = bases
                            * per probe (2.3 / 9.1 bytes of device memory per BWT position; 2.7 / 10.7 for indexes of
                            * 2^32 positions and more, whose sectors carry a 64-bit counter) */
    int32_t fold_ids;      /* lookup ids of the strand-folded dictionary (8 bytes per distinct k-mer, read only by OUT_ORDERS
                            * queries): 0 = built on the first lookup, 1 = built with the tier, -1 = never (lookups then run
                            * on the backward-search kernels) */
    int32_t locality;      /* minimizer-bucketed dictionary for k-mers that come out of a text (reads, chunks; presence outputs):
                            * neighbouring k-mers share their memory requests. A second copy of the dictionary rows (8 bytes per
                            * distinct k-mer + 8 per bucket), opt-in: 0 = off, 1 = built at load, 2 = built by the first text
                            * call of 2^24 k-mers or more */
    int64_t reserved[4];
} fmsi_gpu_options;

typedef struct {
    uint64_t n_bwt;      /* N */
    uint64_t counts[4];  /* C-array of the reference's .misc */
    uint64_t dollar_position;
    uint64_t mask_ones;  /* number of represented occurrences = size of the lookup id space */
    uint64_t hbm_bytes;  /* device memory held by this index */
    int32_t k;
    int32_t has_klcp;
    int32_t prefix_t;
    int32_t wide;        /* 1 when N >= 2^32 (64-bit positions on device) */
    int32_t device;
    int32_t dict;        /* resident dictionary tier: 0 none, 1 SA-ordered, 2 strand-folded */
    int32_t dict_t;      /* bucket depth of that tier (bases) */
    int32_t multistep;   /* bases per probe of the resident multi-step rank arrays (0 = none) */
    int32_t fold_ids;    /* 1 when the strand-folded dictionary's lookup ids are resident */
    int32_t locality;    /* minimizer length of the resident minimizer-bucketed dictionary (0 = not resident) */
    int32_t reserved[2];
} fmsi_gpu_index_info;

const char *fmsi_gpu_last_error(void);
int fmsi_gpu_abi_version(void);
int fmsi_gpu_device_count(void);

/* load_index(fn, use_klcp) — reference src/fms_index.h:502-526. Reads
 * <prefix>.fmsi.{ac_gt,ac,gt,mask,klcp,misc}, converts them to the blocked GPU layout and uploads
 * it to `device`. opts may be NULL. */
int fmsi_gpu_index_load(const char *prefix, int use_klcp, int device, const fmsi_gpu_options *opts,
                        fmsi_gpu_index **out);
/* In-memory construction from raw bit arrays (one byte per bit) — the form of the hand-built
 * fixtures in the reference's unit tests (tests/fms_index_test.h:10-69). klcp may be NULL. */
int fmsi_gpu_index_from_bits(const uint8_t *ac_gt, size_t n_ac_gt, const uint8_t *ac, size_t n_ac,
                             const uint8_t *gt, size_t n_gt, const uint8_t *mask, size_t n_mask,
                             const uint64_t counts[4], uint64_t dollar_position, const uint8_t *klcp,
                             size_t n_klcp, int k, int device, const fmsi_gpu_options *opts,
                             fmsi_gpu_index **out);
/* construct<T>(ms, k, use_klcp) — reference src/fms_index.h:397-460 (+ construct_klcp :357-385),
 * on the GPU: suffix-sorts the mask-cased superstring `ms` (ACGTacgt, upper case = ON, n characters,
 * host or device memory per `mem`), and builds the device-resident index directly. The suffix
 * array is unique, so the result equals the reference's `fmsi index` output (fmsi_gpu_index_save
 * writes byte-identical files; the reference itself builds no kLCP array for k > 64, src/main.cpp:225-228).
 * Limits: n + 1 < 2^32, k <= FMSI_GPU_MAX_K. */
int fmsi_gpu_index_build(const char *ms, size_t n, int k, int with_klcp, int mem, int device,
                         const fmsi_gpu_options *opts, fmsi_gpu_index **out);
/* dump_index(index, fn) — reference src/fms_index.h:484-500: writes <prefix>.fmsi.{ac_gt,ac,gt,mask,
 * klcp,misc} in the reference's (sdsl) formats. Only for indexes made by fmsi_gpu_index_build. */
int fmsi_gpu_index_save(const fmsi_gpu_index *idx, const char *prefix);
int fmsi_gpu_index_free(fmsi_gpu_index *idx);
int fmsi_gpu_index_get_info(const fmsi_gpu_index *idx, fmsi_gpu_index_info *info);

/* ---- building blocks, one device thread per element; host arrays in and out ---------------- */
/* rank(index, i, c) — reference src/fms_index.h:68-86. */
int fmsi_gpu_rank(fmsi_gpu_index *idx, const uint64_t *i, const uint8_t *c, size_t n, uint64_t *out);
/* update_range(index, i, j, c) — src/fms_index.h:98-103. In place. */
int fmsi_gpu_update_range(fmsi_gpu_index *idx, uint64_t *i, uint64_t *j, const uint8_t *c, size_t n);
/* extend_range_with_klcp(index, i, j) — src/fms_index.h:106-109. In place. */
int fmsi_gpu_extend_range_with_klcp(fmsi_gpu_index *idx, uint64_t *i, uint64_t *j, size_t n);
/* get_range_with_pattern(index, sa_start, sa_end, pattern, k) — src/fms_index.h:117-124; patterns
 * are packed k-mers, any 1 <= k <= 32. use_table: 0 = plain k LF-steps, 1 = through the suffix
 * table (an empty interval may then be reported as any i == j). */
int fmsi_gpu_get_range_with_pattern(fmsi_gpu_index *idx, const uint64_t *kmers, int k, size_t n,
                                    int use_table, uint64_t *sa_start, uint64_t *sa_end);
/* infer_presence<maximized_ones>(index, sa_start, sa_end) — src/fms_index.h:126-144: -1/0/1. */
int fmsi_gpu_infer_presence(fmsi_gpu_index *idx, const uint64_t *sa_start, const uint64_t *sa_end,
                            size_t n, int maximized_ones, int8_t *out);
/* kmer_order_if_present(index, sa_start, sa_end) — src/fms_index.h:150-156. */
int fmsi_gpu_kmer_order_if_present(fmsi_gpu_index *idx, const uint64_t *sa_start,
                                   const uint64_t *sa_end, size_t n, int64_t *out);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* query_kmers_single<mode>() over n independent packed k-mers — reference src/fms_index.h:263-331
 * (one k-mer per element instead of one chunk per call).
 * results layout:
 *   PRESENCE, LAZY : uint8[n]     1 iff the reference prints '1'
 *   PRESENCE, BOTH : uint8[n]     (f+1) | (r+1) << 2, f/r = single_query_or<> on the k-mer / its RC
 *   ORDERS,   LAZY : int64[n]     the id the reference prints (-1 absent)
 *   ORDERS,   BOTH : int64[2n]    {f, r} = single_query_order on the k-mer / its RC
 * k must equal the index's k. */
int fmsi_gpu_query_kmers(fmsi_gpu_index *idx, int mode, int output, int strands,
                         const uint64_t *kmers, size_t n, int k, void *results, int mem,
                         void *stream);

/* query_kmers<mode>() over chunks of ACGT text — reference src/fms_index.h:333-342, i.e.
 * query_kmers_streaming (:181-254) when `streaming` != 0 (needs kLCP) else query_kmers_single.
 * bases: ASCII ACGTacgt only (the caller splits records at other characters like
 * ms_query, src/main.cpp:337-370); chunk c is bases[chunk_off[c] .. chunk_off[c]+chunk_len[c]),
 * chunk_len[c] >= k, and yields chunk_len[c]-k+1 results starting at result index res_off[c]
 * (res_off[] non-decreasing; result slots that no chunk covers are left unspecified).
 * With streaming != 0 a chunk holds at most FMSI_GPU_MAX_STREAM_KMERS k-mers and k (<= 32) must equal the index's k
 * (FMSI_GPU_ERR_K otherwise: the kLCP array describes the (k-1)-mers of that k, src/fms_index.h:357-385).
 * Host-mode calls validate every chunk (FMSI_GPU_ERR_ARG when one runs past the text); device-mode calls cannot,
 * and a chunk that runs past the text yields no results there. Device-mode calls on one index share its launch
 * scratch: serialise them (one stream at a time per index; replicas of a pool are independent).
 * results layout as for fmsi_gpu_query_kmers with n = total number of k-mers.
 * k may exceed 32 here (the reference's get_range_with_pattern, src/fms_index.h:117-124, takes any k
 * and `fmsi index` builds such indexes, src/main.cpp:225-233), up to FMSI_GPU_MAX_K: the k-mers are
 * then searched straight from the packed text. For k > 32 `streaming` only checks that the kLCP array
 * is loaded — kLCP interval reuse is a shortcut that never changes a per-strand value, so those
 * chunks may hold any number of k-mers and are answered by plain backward search. */
#define FMSI_GPU_MAX_STREAM_KMERS 64
#define FMSI_GPU_MAX_K 65536
int fmsi_gpu_query_chunks(fmsi_gpu_index *idx, int mode, int output, int strands, int streaming,
                          const char *bases, size_t n_bases, const uint64_t *chunk_off,
                          const uint32_t *chunk_len, const uint64_t *res_off, size_t n_chunks,
                          size_t n_results, int k, void *results, int mem, void *stream);

/* As fmsi_gpu_query_chunks with the text already 2-bit packed (FMSI_GPU_TEXT_PACKED2: text2 holds (n_bases + 31) / 32
 * words): a 150-base read then crosses PCIe as 0.31 bytes per k-mer instead of 1.25, and with
 * FMSI_GPU_OUT_PRESENCE_BITS its answers return as 0.125 bytes per k-mer. chunk_off / chunk_len are in bases. */
int fmsi_gpu_query_chunks_packed(fmsi_gpu_index *idx, int mode, int output, int strands, int streaming,
                                 const uint64_t *text2, size_t n_bases, const uint64_t *chunk_off,
                                 const uint32_t *chunk_len, const uint64_t *res_off, size_t n_chunks,
                                 size_t n_results, int k, void *results, int mem, void *stream);

/* The same over whole READS: read r is bases [read_off[r], read_off[r+1]) of the 2-bit packed text (read_off[] has
 * n_reads + 1 non-decreasing entries, the last <= n_bases), a read shorter than k yields nothing, and results come back to
 * back in read order — n_results = sum over reads of max(0, length - k + 1), which the caller computes to size `results`
 * (checked: FMSI_GPU_ERR_ARG). The device cuts the reads into the chunks its kernels take (<= FMSI_GPU_MAX_STREAM_KMERS
 * k-mers each for the streaming kernel), so a read costs 8 bytes of offsets on PCIe instead of 20 bytes per chunk and the
 * host validates one entry per read: what ms_query's record loop (src/main.cpp:337-354) hands to query_kmers(), batched.
 * With mem = DEVICE read_off is a device pointer as well (malformed entries then yield no results). */
int fmsi_gpu_query_reads_packed(fmsi_gpu_index *idx, int mode, int output, int strands, int streaming,
                                const uint64_t *text2, size_t n_bases, const uint64_t *read_off, size_t n_reads,
                                size_t n_results, int k, void *results, int mem, void *stream);

/* ---- f-MS framework: general demasking functions -------------------------------------------- */
/* query_kmers<query_mode::general>() — reference src/fms_index.h:317-327 with single_query_general
 * (:171-179) and the demasking functions of src/functions.h:7-57: the number of ON occurrences and of
 * all occurrences of the k-mer and of its reverse complement (a self-complementary k-mer is counted
 * once) are passed to f; results[q] = 1 iff the reference prints '1'. kLCP streaming does not apply
 * (the reference ignores -S in this mode) and the strand predictor is not involved. */
enum { FMSI_GPU_F_OR = 0, FMSI_GPU_F_AND = 1, FMSI_GPU_F_XOR = 2, FMSI_GPU_F_RANGE = 3 };
typedef struct {
    int32_t kind; /* FMSI_GPU_F_* */
    int32_t r, s; /* FMSI_GPU_F_RANGE: represented iff r <= #ON occurrences <= s ("INT-INT") */
    int32_t reserved;
} fmsi_gpu_function;
int fmsi_gpu_query_kmers_general(fmsi_gpu_index *idx, const fmsi_gpu_function *f, const uint64_t *kmers,
                                 size_t n, int k, uint8_t *results, int mem, void *stream);
int fmsi_gpu_query_chunks_general(fmsi_gpu_index *idx, const fmsi_gpu_function *f, const char *bases,
                                  size_t n_bases, const uint64_t *chunk_off, const uint32_t *chunk_len,
                                  const uint64_t *res_off, size_t n_chunks, size_t n_results, int k,
                                  uint8_t *results, int mem, void *stream);

/* ---- multi-GPU scheduler -------------------------------------------------------------------- */
/* The query path shards by independent units (k-mers; chunks for -S) with a full index replica per
 * GPU and no data-path collective (the reference is single-threaded: ms_query, src/main.cpp:238-375,
 * answers records one after another). A pool holds `primary` plus one replica per further entry of
 * devices[] — copied device to device (cudaMemcpyPeer: NVLink when peer access is available), so the
 * index files are parsed and converted once. An ordinal may repeat (several replicas on one GPU).
 * Calls take HOST buffers, split them into contiguous ranges (chunk ranges balanced by k-mer count,
 * chunks never split), drive every member from its own host thread and return when all results are
 * in `results`, in query order. fmsi_gpu_pool_free releases the replicas, not `primary`. */
typedef struct fmsi_gpu_pool fmsi_gpu_pool;
int fmsi_gpu_pool_create(fmsi_gpu_index *primary, const int *devices, int n_devices, fmsi_gpu_pool **out);
int fmsi_gpu_pool_size(const fmsi_gpu_pool *pool);
/* member m (0 = primary) for callers that schedule whole batches themselves; owned by the pool */
fmsi_gpu_index *fmsi_gpu_pool_member(fmsi_gpu_pool *pool, int m);
int fmsi_gpu_pool_free(fmsi_gpu_pool *pool);
/* as fmsi_gpu_query_kmers with mem = HOST */
int fmsi_gpu_pool_query_kmers(fmsi_gpu_pool *pool, int mode, int output, int strands, const uint64_t *kmers,
                              size_t n, int k, void *results);
/* as fmsi_gpu_query_chunks with mem = HOST and results back to back in chunk order
 * (res_off[c] = sum over earlier chunks of chunk_len - k + 1) */
int fmsi_gpu_pool_query_chunks(fmsi_gpu_pool *pool, int mode, int output, int strands, int streaming,
                               const char *bases, size_t n_bases, const uint64_t *chunk_off,
                               const uint32_t *chunk_len, size_t n_chunks, size_t n_results, int k,
                               void *results);

/* Number of kernel launches issued by this library on behalf of the calling process so far
 * (bench.py reports it as gpu_launches). */
uint64_t fmsi_gpu_launch_count(void);

/* Probe accounting for the roofline (bench.py): while `on`, the backward-search, streaming and strand-folded
 * dictionary kernels launched for this index add the number of DEPENDENT memory requests they issue (suffix-table
 * entries, rank / multi-step / aux / bucket / rows sectors; not the coalesced reads of the queries themselves) to a
 * running total. The call synchronises the device, stores the total so far in *total (may be NULL) and switches the
 * accounting on or off. Off by default: the timed kernels carry only a null-pointer test. */
int fmsi_gpu_count_probes(fmsi_gpu_index *idx, int on, uint64_t *total);

#ifdef __cplusplus
}
#endif
#endif
