// fmsi_gpu.cu — C-ABI of libfmsi_gpu.so (see include/fmsi_gpu.h). Host-side plumbing only:
// index upload, suffix-table construction, staging/pipelining of host buffers and kernel dispatch.
#include "../../include/fmsi_gpu.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <thrust/iterator/transform_iterator.h>

#include "device_index.cuh"
#include "dict.cuh"
#include "fold.cuh"
#include "loc.cuh"
#include "index_build.cuh"
#include "index_convert.cuh"
#include "index_layout.hpp"
#include "longk_kernels.cuh"
#include "multistep.cuh"
#include "query_kernels.cuh"
#include "stream_kernels.cuh"

using namespace fmsi;

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(FMSI_GPU_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));   \
    } while (0)

constexpr int kSlots = 3;
constexpr size_t kBatchKmers = 16u << 20;  // k-mers per pipelined host batch

// Per-launch device scratch: ctr[0] = work cursor, ctr[1] = overflow count of the dictionary kernel,
// ctr[2] = work cursor of the fixup launch; ovf = overflow list (one u32 per query of the launch).
struct LaunchScratch {
    unsigned long long *ctr = nullptr;
    void *ovf = nullptr;
    size_t ovf_cap = 0;
};

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    void *d_in = nullptr, *d_out = nullptr, *d_aux = nullptr;
    size_t in_cap = 0, out_cap = 0, aux_cap = 0;
    LaunchScratch ls;
};

}  // namespace

struct fmsi_gpu_index {
    int device = 0;
    int sm_count = 0;
    HostIndex meta;  // vectors released after upload
    DevIndex dev{};
    void *d_rank = nullptr, *d_aux = nullptr, *d_table = nullptr, *d_sb = nullptr, *d_counts = nullptr, *d_rows = nullptr;
    void *d_fbuckets = nullptr, *d_frows = nullptr, *d_fids = nullptr;  // strand-folded dictionary (fold.cuh)
    void *d_multi = nullptr;                                            // multi-step rank arrays (multistep.cuh)
    void *d_ldir = nullptr, *d_lrows = nullptr;                         // minimizer-bucketed dictionary (loc.cuh)
    size_t b_ldir = 0, b_lrows = 0;
    LocView loc{};
    int loc_policy = 0;        // fmsi_gpu_options.locality: 0 / -1 = never, 1 = at load, 2 = on the first large text call
    bool loc_failed = false;   // the build did not fit / failed once: text calls stay on the other tiers
    size_t b_multi = 0;
    size_t b_rank = 0, b_aux = 0, b_table = 0, b_sb = 0, b_rows = 0;  // bytes of the device arrays (replication)
    size_t b_fbuckets = 0, b_frows = 0, b_fids = 0;
    uint64_t hbm_bytes = 0;
    bool wide = false;
    DictView dict{};
    FoldView fold{};
    int fold_ids_policy = 0;        // fmsi_gpu_options.fold_ids: 0 = on the first lookup, 1 = with the tier, -1 = never
    bool fold_ids_failed = false;   // the lazy build ran out of memory once: lookups stay on the other tiers
    Slot slots[kSlots];
    cudaStream_t aux_stream = nullptr;       // second query stream of pipelined host-mode chunk calls
    std::vector<cudaEvent_t> piece_events;   // "text piece uploaded and packed" events of those calls
    cudaEvent_t span_done[2] = {nullptr, nullptr};  // "results of the latest span on query stream 0 / 1 are written" (bit-packed output)
    LaunchScratch user;  // scratch for MEM_DEVICE launches
    unsigned long long *d_probes = nullptr;  // fmsi_gpu_count_probes: running total of the kernels' dependent requests
    bool count_probes = false;
    void *d_user_bytes = nullptr;  // MEM_DEVICE calls with bit-packed output: the byte results before packing
    size_t user_bytes_cap = 0;
    // host copies of the BWT/mask/kLCP planes, kept only for indexes made by fmsi_gpu_index_build
    std::vector<uint64_t> plane_lo, plane_hi, plane_mask, plane_klcp;
};

namespace {

int ensure(void **p, size_t *cap, size_t need) {
    if (*cap >= need) return FMSI_GPU_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    size_t want = need + need / 4 + 256;
    CU(cudaMalloc(p, want));
    *cap = want;
    return FMSI_GPU_OK;
}

inline unsigned long long *probe_ctr(const fmsi_gpu_index *idx) { return idx->count_probes ? idx->d_probes : nullptr; }

inline unsigned blocks_for(size_t n, int block = 256) { return (unsigned)((n + block - 1) / block); }

template <typename Kernel>
int persistent_grid(const fmsi_gpu_index *idx, Kernel kern, int block) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    return idx->sm_count * per_sm;
}

u32 pick_chunk(size_t n, int grid, int block) {
    static const int forced = [] {  // $FMSI_GPU_CHUNK: k-mers a warp takes per grab (design experiment switch)
        const char *e = std::getenv("FMSI_GPU_CHUNK");
        return e ? std::atoi(e) : 0;
    }();
    if (forced >= 32) return (u32)(forced / 32 * 32);
    // ~32 grabs per warp: at the end of a launch warps run dry within one grab's duration of each other, and that
    // ramp-down is idle memory system (256 / 512 / 1408 k-mers per grab: 42.6 / 42.5 / 42.1 G k-mers/s at human scale)
    // ... but never fewer than 256 per grab: a grab is two dependent round trips (cursor, then the k-mers) during which
    // the warp has nothing in flight, and with 32-k-mer grabs a small launch (the 4-8 M k-mer batches of the host paths)
    // ran at a third of the rate of a large one (ncu: 355 us per 2^22 k-mers vs 1680 us per 2^26).
    const size_t warps = (size_t)grid * (block / 32);
    size_t c = n / (warps * 32 + 1);
    c = (c / 32) * 32;
    if (c < 256) c = 256;
    if (c > 2048) c = 2048;
    return (u32)c;
}

// north_star's "hot upper levels in L2-persisting windows", as an experiment switch (profiles/backward_ab.py):
// $FMSI_GPU_L2_PERSIST = table|rank|aux|multi[:MiB] pins the head of that array in L2 for the query launches of a
// stream (cudaAccessPolicyWindow, hit ratio 1, misses streaming). Off by default: with the suffix table in front
// of the search no array has a hot head left (every probe is uniformly random), and the A/B shows no gain.
void apply_l2_persist(const fmsi_gpu_index *idx, cudaStream_t st) {
    static const char *env = std::getenv("FMSI_GPU_L2_PERSIST");
    if (!env || !*env) return;
    static thread_local cudaStream_t done_for = (cudaStream_t)-1;
    if (done_for == st) return;
    done_for = st;
    std::string what(env);
    size_t mib = 96;
    const size_t colon = what.find(':');
    if (colon != std::string::npos) {
        mib = (size_t)std::atoll(what.c_str() + colon + 1);
        what = what.substr(0, colon);
    }
    const void *base = what == "table" ? idx->d_table : what == "rank" ? idx->d_rank : what == "aux" ? idx->d_aux : what == "multi" ? idx->d_multi : nullptr;
    size_t avail = what == "table" ? idx->b_table : what == "rank" ? idx->b_rank : what == "aux" ? idx->b_aux : what == "multi" ? idx->b_multi : 0;
    if (!base || !avail) return;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, idx->device) != cudaSuccess) return;
    size_t bytes = std::min(avail, mib << 20);
    bytes = std::min(bytes, (size_t)prop.accessPolicyMaxWindowSize);
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min(bytes, (size_t)prop.persistingL2CacheMaxSize));
    cudaStreamAttrValue attr;
    std::memset(&attr, 0, sizeof attr);
    attr.accessPolicyWindow.base_ptr = const_cast<void *>(base);
    attr.accessPolicyWindow.num_bytes = bytes;
    attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaError_t e = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
    fprintf(stderr, "[fmsi] L2 persisting window on %s: %zu MiB (%s)\n", what.c_str(), bytes >> 20, cudaGetErrorString(e));
    cudaGetLastError();
}

// Backward-search kernel over all n queries.
template <int MODE, int OUT, int STRANDS, bool WIDE>
int launch_query(const fmsi_gpu_index *idx, const DevIndex &d, const u64 *kmers, size_t n, void *out,
                 LaunchScratch &ls, cudaStream_t st) {
    auto kern = query_kmers_kernel<MODE, OUT, STRANDS, WIDE, false>;
    const int grid = persistent_grid(idx, kern, kQueryBlock);
    const u32 chunk = pick_chunk(n, grid, kQueryBlock);
    apply_l2_persist(idx, st);
    CU(cudaMemsetAsync(ls.ctr, 0, 4 * sizeof(unsigned long long), st));
    kern<<<grid, kQueryBlock, 0, st>>>(d, kmers, (u64)n, out, ls.ctr, chunk, nullptr, nullptr, GenF{0, 0, 0}, probe_ctr(idx));
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    return FMSI_GPU_OK;
}

int launch_general(const fmsi_gpu_index *idx, const DevIndex &d, const GenF &gf, const u64 *kmers, size_t n, void *out, LaunchScratch &ls,
                   cudaStream_t st) {
    CU(cudaMemsetAsync(ls.ctr, 0, 4 * sizeof(unsigned long long), st));
    if (idx->wide) {
        auto kern = query_kmers_kernel<K_MODE_GENERAL, K_OUT_PRESENCE, K_STRANDS_LAZY, true, false>;
        const int grid = persistent_grid(idx, kern, kQueryBlock);
        kern<<<grid, kQueryBlock, 0, st>>>(d, kmers, (u64)n, out, ls.ctr, pick_chunk(n, grid, kQueryBlock), nullptr, nullptr, gf, probe_ctr(idx));
    } else {
        auto kern = query_kmers_kernel<K_MODE_GENERAL, K_OUT_PRESENCE, K_STRANDS_LAZY, false, false>;
        const int grid = persistent_grid(idx, kern, kQueryBlock);
        kern<<<grid, kQueryBlock, 0, st>>>(d, kmers, (u64)n, out, ls.ctr, pick_chunk(n, grid, kQueryBlock), nullptr, nullptr, gf, probe_ctr(idx));
    }
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    return FMSI_GPU_OK;
}

// Dictionary kernel over all n queries (n < 2^32) + the backward-search fixup launch over the
// queries it put on the overflow list (usually none: that launch then exits at once).
template <int MODE, int OUT, int STRANDS>
int launch_dict(const fmsi_gpu_index *idx, const DevIndex &d, const u64 *kmers, size_t n, void *out,
                LaunchScratch &ls, cudaStream_t st) {
    int rc;
    if ((rc = ensure(&ls.ovf, &ls.ovf_cap, n * sizeof(u32)))) return rc;
    CU(cudaMemsetAsync(ls.ctr, 0, 4 * sizeof(unsigned long long), st));
    const bool pay64 = idx->dict.B > 16;
    int grid;
    if (pay64) {
        auto kern = dict_query_kernel<MODE, OUT, STRANDS, true>;
        grid = persistent_grid(idx, kern, kQueryBlock);
        kern<<<grid, kQueryBlock, 0, st>>>(d, idx->dict, kmers, (u64)n, out, ls.ctr, pick_chunk(n, grid, kQueryBlock), (u32 *)ls.ovf, ls.ctr + 1);
    } else {
        auto kern = dict_query_kernel<MODE, OUT, STRANDS, false>;
        grid = persistent_grid(idx, kern, kQueryBlock);
        kern<<<grid, kQueryBlock, 0, st>>>(d, idx->dict, kmers, (u64)n, out, ls.ctr, pick_chunk(n, grid, kQueryBlock), (u32 *)ls.ovf, ls.ctr + 1);
    }
    CU(cudaGetLastError());
    auto fix = query_kmers_kernel<MODE, OUT, STRANDS, false, true>;
    const int fgrid = persistent_grid(idx, fix, kQueryBlock);
    fix<<<fgrid, kQueryBlock, 0, st>>>(d, kmers, 0, out, ls.ctr + 2, 32u, (const u32 *)ls.ovf, ls.ctr + 1, GenF{0, 0, 0}, probe_ctr(idx));
    CU(cudaGetLastError());
    g_launches.fetch_add(2);
    return FMSI_GPU_OK;
}

// Strand-folded dictionary kernel over all n queries: one launch, every answer final.
bool fold_ld64() {  // $FMSI_GPU_LD64=0: whole-line L2 fills for bucket probes (design experiment switch)
    static const bool v = [] {
        const char *e = std::getenv("FMSI_GPU_LD64");
        return !(e && std::atoi(e) == 0);
    }();
    return v;
}

template <int MODE, int OUT, int STRANDS, bool PAY64, bool LD64>
int launch_fold_v(const fmsi_gpu_index *idx, const u64 *kmers, size_t n, void *out, LaunchScratch &ls, cudaStream_t st, const ReadSrc &rs) {
    auto kern = fold_query_kernel<MODE, OUT, STRANDS, PAY64, LD64>;
    const int grid = persistent_grid(idx, kern, kQueryBlock);
    CU(cudaMemsetAsync(ls.ctr, 0, 4 * sizeof(unsigned long long), st));
    kern<<<grid, kQueryBlock, 0, st>>>(idx->fold, kmers, (u64)n, out, ls.ctr, pick_chunk(n, grid, kQueryBlock), probe_ctr(idx), rs);
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    return FMSI_GPU_OK;
}

template <int MODE, int OUT, int STRANDS>
int launch_fold(const fmsi_gpu_index *idx, const u64 *kmers, size_t n, void *out, LaunchScratch &ls, cudaStream_t st, const ReadSrc &rs) {
    const bool pay64 = idx->fold.B > 16, ld64 = fold_ld64();
    if (pay64) return ld64 ? launch_fold_v<MODE, OUT, STRANDS, true, true>(idx, kmers, n, out, ls, st, rs)
                           : launch_fold_v<MODE, OUT, STRANDS, true, false>(idx, kmers, n, out, ls, st, rs);
    return ld64 ? launch_fold_v<MODE, OUT, STRANDS, false, true>(idx, kmers, n, out, ls, st, rs)
                : launch_fold_v<MODE, OUT, STRANDS, false, false>(idx, kmers, n, out, ls, st, rs);
}

// does the strand-folded dictionary answer this query shape? (then reads can feed it directly: ReadSrc)
bool fold_answers(const fmsi_gpu_index *idx, int k, int mode, int output) {
    return !idx->wide && mode != 2 && idx->fold.enabled && (u32)k == idx->fold.k && (output != FMSI_GPU_OUT_ORDERS || idx->fold.ids);
}

// Minimizer-bucketed dictionary (loc.cuh): text-derived queries with presence outputs.
bool loc_answers(const fmsi_gpu_index *idx, int k, int mode, int output) {
    return !idx->wide && mode != 2 && idx->loc.enabled && (u32)k == idx->loc.g.k && output == FMSI_GPU_OUT_PRESENCE;
}
template <int MODE, int STRANDS>
int launch_loc_v(const fmsi_gpu_index *idx, const u64 *packed, u64 n_bases, const u64 *coff, const u32 *clen, const u64 *roff, size_t n_chunks, void *out,
                 LaunchScratch &ls, cudaStream_t st) {
    auto kern = loc_stream_kernel<MODE, STRANDS>;
    const size_t smem = loc_stream_smem(idx->loc.g);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kLocBlock, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int grid = idx->sm_count * per_sm;
    const size_t warps = (size_t)grid * (kLocBlock / 32);
    size_t grab = n_chunks / (warps * 8 + 1);  // ~8 grabs per warp, whole tiles of 32 chunks
    grab = grab < 32 ? 32 : (grab > 1024 ? 1024 : grab / 32 * 32);
    CU(cudaMemsetAsync(ls.ctr, 0, 4 * sizeof(unsigned long long), st));
    kern<<<grid, kLocBlock, smem, st>>>(idx->loc, packed, n_bases, coff, clen, roff, (u64)n_chunks, (unsigned char *)out, ls.ctr, (u32)grab, probe_ctr(idx));
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    return FMSI_GPU_OK;
}
int launch_loc(const fmsi_gpu_index *idx, int mode, int strands, const u64 *packed, u64 n_bases, const u64 *coff, const u32 *clen, const u64 *roff,
               size_t n_chunks, void *out, LaunchScratch &ls, cudaStream_t st) {
    if (mode == FMSI_GPU_MODE_ALL)
        return strands == FMSI_GPU_STRANDS_BOTH ? launch_loc_v<K_MODE_ALL, K_STRANDS_BOTH>(idx, packed, n_bases, coff, clen, roff, n_chunks, out, ls, st)
                                                : launch_loc_v<K_MODE_ALL, K_STRANDS_LAZY>(idx, packed, n_bases, coff, clen, roff, n_chunks, out, ls, st);
    return strands == FMSI_GPU_STRANDS_BOTH ? launch_loc_v<K_MODE_OR, K_STRANDS_BOTH>(idx, packed, n_bases, coff, clen, roff, n_chunks, out, ls, st)
                                            : launch_loc_v<K_MODE_OR, K_STRANDS_LAZY>(idx, packed, n_bases, coff, clen, roff, n_chunks, out, ls, st);
}

template <int MODE, int OUT, int STRANDS>
int launch_query_w(const fmsi_gpu_index *idx, const DevIndex &d, const u64 *kmers, size_t n, void *out,
                   LaunchScratch &ls, cudaStream_t st, const ReadSrc &rs) {
    if (idx->wide) return launch_query<MODE, OUT, STRANDS, true>(idx, d, kmers, n, out, ls, st);
    if (idx->fold.enabled && d.k == idx->fold.k && (OUT != K_OUT_ORDERS || idx->fold.ids)) return launch_fold<MODE, OUT, STRANDS>(idx, kmers, n, out, ls, st, rs);
    if (idx->dict.enabled && d.k == idx->dict.k && d.t && n < (1ull << 32)) return launch_dict<MODE, OUT, STRANDS>(idx, d, kmers, n, out, ls, st);
    return launch_query<MODE, OUT, STRANDS, false>(idx, d, kmers, n, out, ls, st);
}

// mode FMSI_GPU_MODE_GENERAL_ (internal) carries its function in *gf
constexpr int FMSI_GPU_MODE_GENERAL_ = 2;
int dispatch_query(const fmsi_gpu_index *idx, const DevIndex &d, int mode, int output, int strands,
                   const u64 *kmers, size_t n, void *out, LaunchScratch &ls, cudaStream_t st, const GenF *gf = nullptr,
                   const ReadSrc &rs = ReadSrc{nullptr, nullptr, nullptr, 0, 0}) {
    if (mode == FMSI_GPU_MODE_GENERAL_) return launch_general(idx, d, *gf, kmers, n, out, ls, st);
    if (output == FMSI_GPU_OUT_ORDERS) {
        if (strands == FMSI_GPU_STRANDS_BOTH) return launch_query_w<K_MODE_OR, K_OUT_ORDERS, K_STRANDS_BOTH>(idx, d, kmers, n, out, ls, st, rs);
        return launch_query_w<K_MODE_OR, K_OUT_ORDERS, K_STRANDS_LAZY>(idx, d, kmers, n, out, ls, st, rs);
    }
    if (mode == FMSI_GPU_MODE_ALL) {
        if (strands == FMSI_GPU_STRANDS_BOTH) return launch_query_w<K_MODE_ALL, K_OUT_PRESENCE, K_STRANDS_BOTH>(idx, d, kmers, n, out, ls, st, rs);
        return launch_query_w<K_MODE_ALL, K_OUT_PRESENCE, K_STRANDS_LAZY>(idx, d, kmers, n, out, ls, st, rs);
    }
    if (strands == FMSI_GPU_STRANDS_BOTH) return launch_query_w<K_MODE_OR, K_OUT_PRESENCE, K_STRANDS_BOTH>(idx, d, kmers, n, out, ls, st, rs);
    return launch_query_w<K_MODE_OR, K_OUT_PRESENCE, K_STRANDS_LAZY>(idx, d, kmers, n, out, ls, st, rs);
}

// cub temp storage of one ExclusiveSum over n u32 counts, and the reads-mode device scratch behind the chunk arrays:
// read offsets, result / chunk prefix sums (n + 1 u64 each), per-read counts (2 x u32), scan temp.
struct U32ToU64 {  // 64-bit accumulation of 32-bit counts
    __host__ __device__ __forceinline__ u64 operator()(const u32 x) const { return (u64)x; }
};
typedef thrust::transform_iterator<U32ToU64, const u32 *> CountIter;
size_t reads_scan_temp_bytes(size_t n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, CountIter((const u32 *)nullptr, U32ToU64()), (u64 *)nullptr, (int)n);
    return bytes + 256;
}
size_t reads_scratch_bytes(size_t n) { return 3 * (n + 1) * 8 + 2 * ((n * 4 + 63) & ~size_t(63)) + reads_scan_temp_bytes(n) + 1024; }

size_t result_bytes(int output, int strands) {
    if (output == FMSI_GPU_OUT_PRESENCE) return 1;
    return strands == FMSI_GPU_STRANDS_BOTH ? 16 : 8;
}

// DevIndex for a query with k-mer length k (the table only helps when t <= k).
DevIndex dev_for_k(const fmsi_gpu_index *idx, int k) {
    DevIndex d = idx->dev;
    d.k = (u32)k;
    if (d.t > (u32)k) d.t = 0;
    return d;
}

// Suffix table at depth t as a plain {i, j} array (level by level; returns the full level).
template <bool WIDE>
int build_table_levels(fmsi_gpu_index *idx, u32 t, TableEntry<WIDE> **out_full) {
    typedef TableEntry<WIDE> E;
    const u64 total = 1ull << (2 * t);
    E *full = nullptr, *quarter = nullptr;
    CU(cudaMalloc(&full, total * sizeof(E)));
    if (t > 1) CU(cudaMalloc(&quarter, (total / 4) * sizeof(E)));
    DevIndex d = idx->dev;
    for (u32 s = 1; s <= t; ++s) {
        E *cur = ((t - s) % 2 == 0) ? full : quarter;
        const E *prev = ((t - s) % 2 == 0) ? quarter : full;
        const u64 cnt = 1ull << (2 * s);
        const int block = 256;
        const u64 grid = (cnt + block - 1) / block;
        table_level_kernel<WIDE><<<(unsigned)grid, block>>>(d, prev, cur, s);
        CU(cudaGetLastError());
        g_launches.fetch_add(1);
    }
    CU(cudaDeviceSynchronize());
    if (quarter) cudaFree(quarter);
    *out_full = full;
    return FMSI_GPU_OK;
}

template <bool WIDE>
int build_table(fmsi_gpu_index *idx, u32 t) {
    if (t == 0) return FMSI_GPU_OK;
    TableEntry<WIDE> *full = nullptr;
    int rc = build_table_levels<WIDE>(idx, t, &full);
    if (rc) return rc;
    idx->d_table = full;
    idx->dev.table = full;
    idx->dev.t = t;
    idx->dev.tshift = WIDE ? 4 : 3;
    idx->b_table = (1ull << (2 * t)) * sizeof(TableEntry<WIDE>);
    idx->hbm_bytes += idx->b_table;
    return FMSI_GPU_OK;
}

// Dictionary tier (dict.cuh): rows from the BWT, then 32-byte buckets in place of the {i, j} table.
int build_dict(fmsi_gpu_index *idx, u32 t) {
    const HostIndex &h = idx->meta;
    const u64 N = h.n;
    const u32 k = (u32)h.k, B = k - t;
    DevIndex d = idx->dev;
    u32 *psi = nullptr;
    u64 *rows = nullptr;
    CU(cudaMalloc(&psi, N * sizeof(u32)));
    if (cudaMalloc(&rows, (N + 8) * sizeof(u64)) != cudaSuccess) {  // + one sector: ROWS reads whole sectors
        cudaFree(psi);
        return fail(FMSI_GPU_ERR_NOMEM, "dictionary rows: out of device memory");
    }
    cudaMemset(rows + N, 0, 8 * sizeof(u64));
    psi_scatter_kernel<<<blocks_for(N), 256>>>(d, psi);
    rows_walk_kernel<<<blocks_for(N), 256>>>(d, psi, (u32)h.counts[1], (u32)h.counts[2], (u32)h.counts[3], t, B, rows);
    rows_invalidate_kernel<<<1, 32>>>(d, k, rows);
    g_launches.fetch_add(3);
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(psi);
    if (e != cudaSuccess) {
        cudaFree(rows);
        return fail(FMSI_GPU_ERR_CUDA, std::string("dictionary rows: ") + cudaGetErrorString(e));
    }
    TableEntry<false> *full = nullptr;
    int rc = build_table_levels<false>(idx, t, &full);
    if (rc) {
        cudaFree(rows);
        return rc;
    }
    const u64 total = 1ull << (2 * t);
    Bucket *buckets = nullptr;
    if (cudaMalloc(&buckets, total * sizeof(Bucket)) != cudaSuccess) {
        cudaFree(rows);
        cudaFree(full);
        return fail(FMSI_GPU_ERR_NOMEM, "dictionary buckets: out of device memory");
    }
    if (B > 16) bucket_fill_kernel<true><<<blocks_for(total), 256>>>(full, rows, total, buckets);
    else bucket_fill_kernel<false><<<blocks_for(total), 256>>>(full, rows, total, buckets);
    g_launches.fetch_add(1);
    e = cudaDeviceSynchronize();
    cudaFree(full);
    if (e != cudaSuccess) {
        cudaFree(rows);
        cudaFree(buckets);
        return fail(FMSI_GPU_ERR_CUDA, std::string("dictionary buckets: ") + cudaGetErrorString(e));
    }
    idx->d_table = buckets;
    idx->d_rows = rows;
    idx->dev.table = buckets;
    idx->dev.t = t;
    idx->dev.tshift = 5;
    idx->dict.rows = rows;
    idx->dict.B = B;
    idx->dict.k = k;
    idx->dict.enabled = 1;
    idx->b_table = total * sizeof(Bucket);
    idx->b_rows = (N + 8) * sizeof(u64);
    idx->hbm_bytes += idx->b_table + idx->b_rows;
    return FMSI_GPU_OK;
}

// Strand-folded dictionary (fold.cuh) at bucket depth t, plus a plain {i, j} suffix table (depth tt) for
// the backward-search kernels (general mode, queries with another k). Returns FMSI_GPU_ERR_NOMEM when the
// device cannot hold the build, so that the caller can fall back to a smaller tier.
int build_fold(fmsi_gpu_index *idx, u32 t, u32 tt) {
    const HostIndex &h = idx->meta;
    FoldArrays fa;
    uint64_t launches = 0;
    const bool with_ids = idx->fold_ids_policy == 1;
    try {
        build_fold_on_device(idx->dev, h.counts, (u32)h.k, t, true, with_ids, fa, &launches);
    } catch (const std::exception &e) {
        const bool oom = cudaGetLastError() == cudaErrorMemoryAllocation || std::string(e.what()).find("out of memory") != std::string::npos;
        return fail(oom ? FMSI_GPU_ERR_NOMEM : FMSI_GPU_ERR_CUDA, std::string("strand-folded dictionary: ") + e.what());
    }
    g_launches.fetch_add(launches);
    idx->d_fbuckets = fa.buckets;
    idx->d_frows = fa.orows;
    idx->d_fids = fa.ids;
    idx->b_fbuckets = (1ull << (2 * t)) * sizeof(FoldBucket);
    idx->b_frows = (fa.n_orows + 8) * sizeof(u64);
    idx->b_fids = fa.ids ? (fa.n_rows + 1) * sizeof(uint2) : 0;
    idx->hbm_bytes += idx->b_fbuckets + idx->b_frows + idx->b_fids;
    idx->fold.buckets = fa.buckets;
    idx->fold.orows = fa.orows;
    idx->fold.ids = fa.ids;
    idx->fold.n_rows = fa.n_rows;
    idx->fold.t = t;
    idx->fold.B = (u32)h.k - t;
    idx->fold.k = (u32)h.k;
    idx->fold.enabled = 1;
    return build_table<false>(idx, tt);
}

// `lookup` through the strand-folded dictionary needs ids[] (8 bytes per row), which `query` never reads: unless
// asked for with the tier (fold_ids = 1) it is built here, on the first lookup — the same passes over the BWT as
// the tier's own build, keeping only the ids. Out of memory is not an error: lookups then stay on the other tiers.
void ensure_fold_ids(fmsi_gpu_index *idx) {
    if (!idx->fold.enabled || idx->fold.ids || idx->fold_ids_policy < 0 || idx->fold_ids_failed) return;
    const HostIndex &h = idx->meta;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || fold_build_peak_bytes(h.n, idx->fold.t, true) - (32ull << (2 * idx->fold.t)) > free_b - free_b / 16) {
        idx->fold_ids_failed = true;
        fprintf(stderr, "[fmsi] note: not enough free device memory for the dictionary's lookup ids; lookups use the backward-search kernels\n");
        return;
    }
    FoldArrays fa;
    uint64_t launches = 0;
    const auto t0 = std::chrono::steady_clock::now();
    try {
        build_fold_on_device(idx->dev, h.counts, (u32)h.k, idx->fold.t, false, true, fa, &launches);
    } catch (const std::exception &e) {
        cudaGetLastError();
        idx->fold_ids_failed = true;
        fprintf(stderr, "[fmsi] note: lookup ids of the dictionary not built (%s); lookups use the backward-search kernels\n", e.what());
        return;
    }
    g_launches.fetch_add(launches);
    if (fa.n_rows != idx->fold.n_rows) {  // cannot happen: the row numbering is a function of the index alone
        cudaFree(fa.ids);
        idx->fold_ids_failed = true;
        return;
    }
    idx->d_fids = fa.ids;
    idx->b_fids = (fa.n_rows + 1) * sizeof(uint2);
    idx->hbm_bytes += idx->b_fids;
    idx->fold.ids = fa.ids;
    if (std::getenv("FMSI_GPU_TIMING"))
        fprintf(stderr, "[fmsi timing] lookup ids of the strand-folded dictionary built in %.3f s\n",
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
}

// The minimizer-bucketed dictionary (loc.cuh) answers k-mers that come out of a text — reads, chunks — at about half a
// memory request per k-mer, where the strand-folded dictionary needs one. It costs a second copy of the rows (8 bytes per
// distinct k-mer + 8 bytes per bucket), so it is opt-in: built at load (locality = 1) or here (locality = 2), by the first
// text call that is large enough to pay for the build. Out of memory is not an error: text calls then stay on the other tiers.
constexpr size_t kLocLazyResults = (size_t)1 << 24;
void ensure_loc(fmsi_gpu_index *idx, int k, size_t n_results, bool at_load) {
    if (idx->loc.enabled || idx->loc_failed || idx->loc_policy <= 0 || idx->wide) return;
    const HostIndex &h = idx->meta;
    if (k != h.k || h.k < 1 || h.k > 32 || h.n >= (1ull << 32) - 256) return;
    if (!at_load && (idx->loc_policy != 2 || n_results < kLocLazyResults)) return;
    u32 m = loc_pick_m(h.n, (u32)h.k);
    if (const char *e = std::getenv("FMSI_GPU_LOC_M")) m = (u32)std::atoi(e);
    u32 t = 1;
    while (t < 15 && t < m && (1ull << (2 * (t + 1))) <= h.n) ++t;
    if (const char *e = std::getenv("FMSI_GPU_LOC_T")) t = (u32)std::atoi(e);
    if (!loc_fits((u32)h.k, m, t)) {
        idx->loc_failed = true;
        return;
    }
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || loc_build_peak_bytes(h.n, t) > free_b - free_b / 16) {
        idx->loc_failed = true;
        fprintf(stderr, "[fmsi] note: not enough free device memory for the minimizer-bucketed dictionary; reads use the other tiers\n");
        return;
    }
    LocArrays la;
    uint64_t launches = 0;
    const auto t0 = std::chrono::steady_clock::now();
    try {
        build_loc_on_device(idx->dev, h.counts, (u32)h.k, m, t, la, &launches);
    } catch (const std::exception &e) {
        cudaGetLastError();
        idx->loc_failed = true;
        fprintf(stderr, "[fmsi] note: minimizer-bucketed dictionary not built (%s); reads use the other tiers\n", e.what());
        return;
    }
    g_launches.fetch_add(launches);
    idx->d_ldir = la.dir;
    idx->d_lrows = la.rows;
    idx->b_ldir = (8ull << (2 * t));
    idx->b_lrows = (la.n_rows + 8) * sizeof(u64);
    idx->hbm_bytes += idx->b_ldir + idx->b_lrows;
    idx->loc.dir = la.dir;
    idx->loc.rows = la.rows;
    idx->loc.n_rows = la.n_rows;
    idx->loc.g = loc_geom((u32)h.k, m, t);
    idx->loc.enabled = 1;
    if (std::getenv("FMSI_GPU_TIMING"))
        fprintf(stderr, "[fmsi timing] minimizer-bucketed dictionary (m = %u, t = %u, %llu rows) built in %.3f s\n", m, t,
                (unsigned long long)la.n_rows, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
}

int select_device(fmsi_gpu_index *idx) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(FMSI_GPU_ERR_CUDA, "no CUDA device available (libfmsi_gpu has no CPU fallback)");
    if (idx->device < 0 || idx->device >= ndev) return fail(FMSI_GPU_ERR_ARG, "device ordinal out of range");
    CU(cudaSetDevice(idx->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, idx->device));
    idx->sm_count = prop.multiProcessorCount;
    return FMSI_GPU_OK;
}

int alloc_scratch(LaunchScratch &ls) {
    CU(cudaMalloc(&ls.ctr, 4 * sizeof(unsigned long long)));
    return FMSI_GPU_OK;
}

int alloc_slots(fmsi_gpu_index *idx) {
    int rc;
    for (int s = 0; s < kSlots; ++s) {
        CU(cudaStreamCreateWithFlags(&idx->slots[s].stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&idx->slots[s].done, cudaEventDisableTiming));
        if ((rc = alloc_scratch(idx->slots[s].ls))) return rc;
    }
    return alloc_scratch(idx->user);
}

// Multi-step rank arrays for the backward-search kernels (multistep.cuh). opts->multistep / $FMSI_GPU_MULTISTEP:
// 0 = off, 2 / 3 = bases per probe, -1 = auto: 2 when no dictionary tier is resident (the dictionary tiers answer
// single k-mers themselves) and the arrays fit in a quarter of the free memory. Wide indexes take the 64-bit-counter
// sectors (192 rows each).
int setup_multistep(fmsi_gpu_index *idx, const fmsi_gpu_options *opts) {
    int want = opts ? opts->multistep : -1;
    if (const char *e = std::getenv("FMSI_GPU_MULTISTEP")) want = std::atoi(e);
    if (want == 0) return FMSI_GPU_OK;
    if (want != -1 && want != 2 && want != 3) return fail(FMSI_GPU_ERR_ARG, "multistep must be -1, 0, 2 or 3");
    const HostIndex &h = idx->meta;
    // a narrow layout holds u32 positions and counters: keep clear of 2^32 (the loader switches to wide well before)
    if (!idx->wide && h.n >= (1ull << 32) - 256) return FMSI_GPU_OK;
    const bool wide = idx->wide;
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    u32 m = 2;
    if (want == -1) {
        if (idx->fold.enabled || idx->dict.enabled) return FMSI_GPU_OK;
        if (multi_bytes(h.n, 2, wide) + multi_build_scratch_bytes(h.n, 2, wide) > free_b / 4) return FMSI_GPU_OK;
    } else {
        m = (u32)want;
        if (multi_bytes(h.n, m, wide) + multi_build_scratch_bytes(h.n, m, wide) > free_b - free_b / 16)
            return fail(FMSI_GPU_ERR_NOMEM, "multi-step rank arrays do not fit in device memory");
    }
    MultiBlock *multi = nullptr;
    u32 nblk = 0;
    uint64_t launches = 0;
    try {
        build_multi_on_device(idx->dev, wide, m, &multi, &nblk, &launches);
    } catch (const std::exception &e) {
        const bool oom = cudaGetLastError() == cudaErrorMemoryAllocation || std::string(e.what()).find("out of memory") != std::string::npos;
        if (oom && want == -1) return FMSI_GPU_OK;
        return fail(oom ? FMSI_GPU_ERR_NOMEM : FMSI_GPU_ERR_CUDA, std::string("multi-step rank arrays: ") + e.what());
    }
    g_launches.fetch_add(launches);
    idx->d_multi = multi;
    idx->b_multi = multi_bytes(h.n, m, wide);
    idx->hbm_bytes += idx->b_multi;
    idx->dev.multi = multi;
    idx->dev.multi_m = m;
    idx->dev.multi_nblk = nblk;
    return FMSI_GPU_OK;
}

// Common tail once d_rank / d_aux / d_sb / d_counts hold the layout: suffix table or dictionary,
// streams, launch scratch.
int finish_device_setup(fmsi_gpu_index *idx, const fmsi_gpu_options *opts) {
    HostIndex &h = idx->meta;
    DevIndex &d = idx->dev;
    d.rank = reinterpret_cast<const RankBlock *>(idx->d_rank);
    d.aux = reinterpret_cast<const AuxBlock *>(idx->d_aux);
    d.table = nullptr;
    d.multi = nullptr;
    d.multi_m = 0;
    d.multi_nblk = 0;
    d.sb_base = reinterpret_cast<const u64 *>(idx->d_sb);
    d.n = h.n;
    d.dollar = h.dollar;
    d.t = 0;
    d.tshift = 3;
    d.sb_shift = h.sb_shift >= 63 ? 63 : h.sb_shift;
    d.k = (u32)h.k;
    d.has_klcp = h.has_klcp;

    int t = opts ? opts->prefix_t : -1;
    if (const char *e = std::getenv("FMSI_GPU_PREFIX_T")) t = std::atoi(e);
    int want_dict = opts ? opts->dict : -1;
    if (const char *e = std::getenv("FMSI_GPU_DICT")) want_dict = std::atoi(e);
    idx->fold_ids_policy = opts ? opts->fold_ids : 0;
    if (const char *e = std::getenv("FMSI_GPU_FOLD_IDS")) idx->fold_ids_policy = std::atoi(e);
    idx->loc_policy = opts ? opts->locality : 0;
    if (const char *e = std::getenv("FMSI_GPU_LOCALITY")) idx->loc_policy = std::atoi(e);
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    if (const char *e = std::getenv("FMSI_GPU_FREE_CAP")) free_b = std::min<size_t>(free_b, (size_t)std::atoll(e));  // test hook: a busier GPU

    // Dictionary tiers: narrow indexes with k <= 32. Bucket depth = the largest t <= min(k, 15) with
    // 4^t <= 4N (on average at most ~4 and at least ~0.25 rows per bucket).
    //   dict = 2 / auto: the strand-folded dictionary (fold.cuh) when its build fits in the free memory
    //                    and what stays resident in 60 % of it;
    //   dict = 1 / auto: else the SA-ordered dictionary (dict.cuh): buckets (32 B each) plus rows (8 B per
    //                    SA row, + 4 B per row of build scratch) in half of the free memory;
    //   else the plain suffix table.
    const bool narrow = !idx->wide && h.k >= 1 && h.k <= 32 && h.n < (1ull << 32) - 256;
    auto auto_depth = [&]() {
        int td = t;
        if (td < 0) {
            td = 1;
            while (td < 15 && (1ull << (2 * (td + 1))) <= 4 * h.n) ++td;
        }
        if (td > h.k) td = h.k;
        if (td > 15) td = 15;
        return td;
    };
    auto table_depth = [&](int cap) {  // largest depth with 4^t <= N, capped by k, `cap` and a quarter of the free memory
        int tt = 0;
        while (tt < cap && (1ull << (2 * (tt + 1))) <= h.n) ++tt;
        if (tt > h.k) tt = h.k;
        const size_t esz = idx->wide ? 16 : 8;
        while (tt > 0 && ((1ull << (2 * tt)) * esz * 5 / 4) > free_b / 4) --tt;
        return tt;
    };
    int rc = FMSI_GPU_OK;
    bool done = false;
    if (narrow && (want_dict < 0 || want_dict == 2)) {
        int td = auto_depth();
        const bool with_ids = idx->fold_ids_policy == 1;
        auto fits = [&](int x) {
            return fold_build_peak_bytes(h.n, (u32)x, with_ids) <= free_b - free_b / 10 && fold_resident_bytes(h.n, (u32)x, with_ids) <= free_b / 10 * 6;
        };
        while (td > 1 && t < 0 && !fits(td)) --td;
        if (td >= 1 && h.k - td <= 30 && fits(td)) {
            rc = build_fold(idx, (u32)td, (u32)table_depth(13));
            if (rc == FMSI_GPU_OK) done = true;
            else if (rc != FMSI_GPU_ERR_NOMEM || want_dict == 2) return rc;
            else {  // release whatever the failed build left and try the next tier
                for (void **p : {&idx->d_fbuckets, &idx->d_frows, &idx->d_fids}) {
                    if (*p) cudaFree(*p);
                    *p = nullptr;
                }
                idx->fold = FoldView{};
                CU(cudaMemGetInfo(&free_b, &total_b));
            }
        }
    }
    if (!done && narrow && want_dict != 0) {
        int td = auto_depth();
        const size_t rows_b = (size_t)h.n * 12;
        while (td > 1 && t < 0 && ((1ull << (2 * td)) * 40 + rows_b) > free_b / 2) --td;
        if (td >= 1 && ((1ull << (2 * td)) * 40 + rows_b) <= free_b - free_b / 8) {
            rc = build_dict(idx, (u32)td);
            if (rc) return rc;
            done = true;
        }
    }
    if (!done) {
        // Suffix-table depth: auto = largest t with 4^t <= N (table no larger than ~2x the rank
        // array), capped by k, by 16 and by a quarter of the free device memory.
        if (t < 0) t = table_depth(16);
        if (t > h.k) t = h.k;
        if (t > 16) t = 16;
        const size_t esz = idx->wide ? 16 : 8;
        while (t > 0 && ((1ull << (2 * t)) * esz * 5 / 4) > free_b / 4) --t;
        rc = idx->wide ? build_table<true>(idx, (u32)t) : build_table<false>(idx, (u32)t);
    }
    if (rc) return rc;
    if (want_dict < 0 && narrow && !idx->fold.enabled)
        // auto asked for the fastest tier that fits: say so when it is not the one-probe dictionary (info.dict tells which)
        fprintf(stderr, "[fmsi] note: the strand-folded dictionary does not fit in the free device memory (%.1f GB free, %.1f GB needed while building); "
                        "using %s — single k-mer queries run several times slower\n",
                free_b / 1e9, fold_build_peak_bytes(h.n, (u32)auto_depth(), false) / 1e9, idx->dict.enabled ? "the SA-ordered dictionary" : "backward search");
    if ((rc = setup_multistep(idx, opts))) return rc;
    if (idx->loc_policy == 1) ensure_loc(idx, (int)h.k, 0, true);
    return alloc_slots(idx);
}

// Raw index material (host words as read from the files, or packed from bit arrays) -> device layout
// (index_convert.cuh) -> suffix table / dictionary, streams, scratch.
struct RawIndex {
    BitVec ac_gt, ac, gt, klcp, mask_plain;  // mask_plain used when has_rrr == false
    RrrFile rrr;
    bool has_rrr = false, has_klcp = false;
    uint64_t counts[4] = {0, 0, 0, 0};
    uint64_t dollar = 0;
    int k = 0;
};

std::chrono::steady_clock::time_point g_load_start;  // $FMSI_GPU_TIMING stage clock of the running fmsi_gpu_index_load

int convert_and_setup(fmsi_gpu_index *idx, const RawIndex &raw, const fmsi_gpu_options *opts) {
    int rc = select_device(idx);
    if (rc) return rc;
    HostIndex &h = idx->meta;
    h.n = raw.ac_gt.nbits;
    h.k = raw.k;
    h.dollar = raw.dollar;
    for (int c = 0; c < 4; ++c) h.counts[c] = raw.counts[c];
    const u64 N = h.n, nblk = (N >> 6) + 1;
    uint64_t launches = 0;
    try {
        if (N == 0) throw std::runtime_error("empty index");
        const u64 mask_bits = raw.has_rrr ? raw.rrr.size : raw.mask_plain.nbits;
        if (mask_bits != N) throw std::runtime_error("mask length does not match the BWT length");
        if (raw.has_klcp && raw.klcp.nbits != N) throw std::runtime_error("kLCP length does not match the BWT length");
        if (raw.dollar >= N) throw std::runtime_error("dollar_position out of range");
        h.has_klcp = raw.has_klcp;
        DevArr<u64> d_acgt, d_ac, d_gt, d_klcp, d_mask;
        upload_words(d_acgt, raw.ac_gt.w.data(), (N + 63) >> 6, nblk + 2);
        upload_words(d_ac, raw.ac.w.data(), (raw.ac.nbits + 63) >> 6, ((raw.ac.nbits + 63) >> 6) + 2);
        upload_words(d_gt, raw.gt.w.data(), (raw.gt.nbits + 63) >> 6, ((raw.gt.nbits + 63) >> 6) + 2);
        if (raw.has_klcp) upload_words(d_klcp, raw.klcp.w.data(), (N + 63) >> 6, nblk + 2);
        if (raw.has_rrr) rrr_decode_on_device(raw.rrr, nblk, d_mask, &launches);
        else upload_words(d_mask, raw.mask_plain.w.data(), (N + 63) >> 6, nblk + 2);
        ConvertedIndex cv;
        convert_on_device(N, d_acgt, d_ac, raw.ac.nbits, d_gt, raw.gt.nbits, d_mask.p, raw.has_klcp ? d_klcp.p : nullptr, raw.counts,
                          opts ? (unsigned)opts->sb_shift_log2 : 0u, cv, &launches);
        BCU(cudaDeviceSynchronize());
        if (std::getenv("FMSI_GPU_TIMING"))
            fprintf(stderr, "[fmsi timing] load: uploaded + converted on the device: %.3f s\n",
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - g_load_start).count());
        g_launches.fetch_add(launches);
        h.mask_ones = cv.mask_ones;
        h.sb_shift = cv.sb_shift;
        idx->wide = cv.sb_shift < 63;
        idx->b_rank = cv.rank.n * sizeof(RankBlock);
        idx->b_aux = cv.aux.n * sizeof(AuxBlock);
        idx->b_sb = cv.nsb * 32;
        idx->d_rank = cv.rank.p;
        idx->d_aux = cv.aux.p;
        idx->d_sb = cv.sb_base.p;
        cv.rank.p = nullptr;
        cv.aux.p = nullptr;
        cv.sb_base.p = nullptr;
    } catch (const std::exception &e) {
        return fail(FMSI_GPU_ERR_IO, e.what());
    }
    CU(cudaMalloc(&idx->d_counts, 32));
    CU(cudaMemcpy(idx->d_counts, h.counts, 32, cudaMemcpyHostToDevice));
    idx->hbm_bytes = idx->b_rank + idx->b_aux + idx->b_sb + 32;
    return finish_device_setup(idx, opts);
}

// Parse <prefix>.fmsi.{ac_gt,ac,gt,mask,klcp,misc} (formats: sdsl_io.hpp). A missing file is the
// reference's "index not correctly loaded" (main.cpp:304-307).
void read_raw_index_files(const std::string &prefix, bool use_klcp, RawIndex &raw) {
    const std::string base = prefix + ".fmsi";
    for (const char *ext : {".ac_gt", ".ac", ".gt", ".mask", ".misc"})
        if (!file_exists(base + ext)) throw std::runtime_error("index not correctly loaded: missing " + base + ext);
    // the four bit vectors are read concurrently, each straight into its word array (a human-scale index is
    // 4 x 388 MB; one thread copying file after file through a staging buffer took most of the load time)
    raw.has_klcp = false;
    const bool want_klcp = use_klcp && file_exists(base + ".klcp");
    std::string errors[4];
    auto reader = [&](int slot, const char *ext, BitVec *dst) {
        try {
            read_bitvec_file(base + ext, *dst);
        } catch (const std::exception &e) {
            errors[slot] = e.what();
        }
    };
    std::vector<std::thread> readers;
    readers.emplace_back(reader, 0, ".ac_gt", &raw.ac_gt);
    readers.emplace_back(reader, 1, ".ac", &raw.ac);
    readers.emplace_back(reader, 2, ".gt", &raw.gt);
    if (want_klcp) readers.emplace_back(reader, 3, ".klcp", &raw.klcp);
    std::string rrr_error;
    try {
        raw.rrr = read_rrr(base + ".mask");
    } catch (const std::exception &e) {
        rrr_error = e.what();
    }
    for (auto &t : readers) t.join();
    for (const std::string &e : errors)
        if (!e.empty()) throw std::runtime_error(e);
    if (!rrr_error.empty()) throw std::runtime_error(rrr_error);
    raw.has_rrr = true;
    if (want_klcp) raw.has_klcp = raw.klcp.nbits > 0;
    std::ifstream in(base + ".misc");
    if (!(in >> raw.dollar >> raw.counts[0] >> raw.counts[1] >> raw.counts[2] >> raw.counts[3] >> raw.k))
        throw std::runtime_error("malformed " + base + ".misc");
}

BitVec bits_to_vec(const uint8_t *bits, size_t n) {
    BitVec b;
    b.resize_bits(n);
    for (size_t p = 0; p < n; ++p)
        if (bits[p]) b.set(p);
    return b;
}

// RAII device scratch for the probe entry points
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    int alloc(size_t bytes) {
        CU(cudaMalloc(&p, bytes ? bytes : 8));
        return FMSI_GPU_OK;
    }
    int put(const void *src, size_t bytes) {
        int rc = alloc(bytes);
        if (rc) return rc;
        CU(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
        return FMSI_GPU_OK;
    }
    int get(void *dst, size_t bytes) {
        CU(cudaMemcpy(dst, p, bytes, cudaMemcpyDeviceToHost));
        return FMSI_GPU_OK;
    }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

}  // namespace

extern "C" {

const char *fmsi_gpu_last_error(void) { return g_err.c_str(); }
int fmsi_gpu_abi_version(void) { return FMSI_GPU_ABI_VERSION; }
int fmsi_gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
uint64_t fmsi_gpu_launch_count(void) { return g_launches.load(); }

int fmsi_gpu_count_probes(fmsi_gpu_index *idx, int on, uint64_t *total) {
    if (!idx) return fail(FMSI_GPU_ERR_ARG, "null index");
    CU(cudaSetDevice(idx->device));
    if (!idx->d_probes) {
        CU(cudaMalloc(&idx->d_probes, sizeof(unsigned long long)));
        CU(cudaMemset(idx->d_probes, 0, sizeof(unsigned long long)));
    }
    CU(cudaDeviceSynchronize());
    if (total) {
        unsigned long long v = 0;
        CU(cudaMemcpy(&v, idx->d_probes, sizeof v, cudaMemcpyDeviceToHost));
        *total = v;
    }
    idx->count_probes = on != 0;
    return FMSI_GPU_OK;
}

int fmsi_gpu_index_load(const char *prefix, int use_klcp, int device, const fmsi_gpu_options *opts,
                        fmsi_gpu_index **out) {
    if (!prefix || !out) return fail(FMSI_GPU_ERR_ARG, "null argument");
    *out = nullptr;
    std::unique_ptr<fmsi_gpu_index> idx(new fmsi_gpu_index());
    idx->device = device;
    RawIndex raw;
    // the CUDA context of the device (hundreds of ms in a fresh process) comes up while the files are read
    std::thread warm([device] {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) == cudaSuccess && device >= 0 && device < ndev && cudaSetDevice(device) == cudaSuccess) cudaFree(nullptr);
        cudaGetLastError();
    });
    const bool timing = std::getenv("FMSI_GPU_TIMING") != nullptr;
    const auto t0 = g_load_start = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    try {
        read_raw_index_files(prefix, use_klcp != 0, raw);
        if (timing) fprintf(stderr, "[fmsi timing] load: files read: %.3f s\n", since());
        warm.join();
        if (timing) fprintf(stderr, "[fmsi timing] load: CUDA context up: %.3f s\n", since());
        if (raw.rrr.size == 0) return fail(FMSI_GPU_ERR_IO, "index not correctly loaded (empty mask)");
    } catch (const std::exception &e) {
        if (warm.joinable()) warm.join();
        return fail(FMSI_GPU_ERR_IO, e.what());
    }
    int rc = convert_and_setup(idx.get(), raw, opts);
    if (timing) fprintf(stderr, "[fmsi timing] load: converted, table / dictionary built: %.3f s\n", since());
    if (rc) {
        fmsi_gpu_index_free(idx.release());
        return rc;
    }
    *out = idx.release();
    return FMSI_GPU_OK;
}

int fmsi_gpu_index_from_bits(const uint8_t *ac_gt, size_t n_ac_gt, const uint8_t *ac, size_t n_ac,
                             const uint8_t *gt, size_t n_gt, const uint8_t *mask, size_t n_mask,
                             const uint64_t counts[4], uint64_t dollar_position, const uint8_t *klcp,
                             size_t n_klcp, int k, int device, const fmsi_gpu_options *opts,
                             fmsi_gpu_index **out) {
    if (!ac_gt || !ac || !gt || !mask || !counts || !out) return fail(FMSI_GPU_ERR_ARG, "null argument");
    *out = nullptr;
    std::unique_ptr<fmsi_gpu_index> idx(new fmsi_gpu_index());
    idx->device = device;
    RawIndex raw;
    raw.ac_gt = bits_to_vec(ac_gt, n_ac_gt);
    raw.ac = bits_to_vec(ac, n_ac);
    raw.gt = bits_to_vec(gt, n_gt);
    raw.mask_plain = bits_to_vec(mask, n_mask);
    raw.has_klcp = klcp && n_klcp;
    if (raw.has_klcp) raw.klcp = bits_to_vec(klcp, n_klcp);
    for (int c = 0; c < 4; ++c) raw.counts[c] = counts[c];
    raw.dollar = dollar_position;
    raw.k = k;
    int rc = convert_and_setup(idx.get(), raw, opts);
    if (rc) {
        fmsi_gpu_index_free(idx.release());
        return rc;
    }
    *out = idx.release();
    return FMSI_GPU_OK;
}

int fmsi_gpu_index_build(const char *ms, size_t n, int k, int with_klcp, int mem, int device,
                         const fmsi_gpu_options *opts, fmsi_gpu_index **out) {
    if (!ms || !out) return fail(FMSI_GPU_ERR_ARG, "null argument");
    *out = nullptr;
    if (mem != FMSI_GPU_MEM_HOST && mem != FMSI_GPU_MEM_DEVICE) return fail(FMSI_GPU_ERR_ARG, "bad mem");
    std::unique_ptr<fmsi_gpu_index> idx(new fmsi_gpu_index());
    idx->device = device;
    int rc = select_device(idx.get());
    if (rc) return rc;
    uint64_t launches = 0;
    try {
        DevArr<char> staged;
        const char *d_ms = ms;
        if (mem == FMSI_GPU_MEM_HOST) {
            staged.alloc(n);
            BCU(cudaMemcpy(staged.p, ms, n, cudaMemcpyHostToDevice));
            d_ms = staged.p;
        }
        BuiltIndex b;
        build_index_on_device(d_ms, n, k, with_klcp != 0, true, b, &launches);
        g_launches.fetch_add(launches);
        HostIndex &h = idx->meta;
        h.n = b.n_bwt;
        h.k = k;
        h.has_klcp = with_klcp != 0;
        for (int c = 0; c < 4; ++c) h.counts[c] = b.counts[c];
        h.dollar = b.dollar;
        h.mask_ones = b.mask_ones;
        h.sb_shift = 63;
        h.sb_base.assign(4, 0);
        idx->wide = false;
        idx->d_rank = b.rank.p;
        idx->d_aux = b.aux.p;
        idx->hbm_bytes = (b.rank.n + b.aux.n) * 32 + 64;
        idx->b_rank = b.rank.n * 32;
        idx->b_aux = b.aux.n * 32;
        idx->b_sb = 32;
        b.rank.p = nullptr;
        b.aux.p = nullptr;
        idx->plane_lo.swap(b.lo);
        idx->plane_hi.swap(b.hi);
        idx->plane_mask.swap(b.mask);
        idx->plane_klcp.swap(b.klcp);
    } catch (const std::exception &e) {
        fmsi_gpu_index_free(idx.release());
        return fail(FMSI_GPU_ERR_CUDA, std::string("index build: ") + e.what());
    }
    HostIndex &h = idx->meta;
    auto cu_fail = [&](const char *what) {
        fmsi_gpu_index_free(idx.release());
        return fail(FMSI_GPU_ERR_CUDA, what);
    };
    if (cudaMalloc(&idx->d_sb, 32) != cudaSuccess || cudaMalloc(&idx->d_counts, 32) != cudaSuccess) return cu_fail("cudaMalloc");
    if (cudaMemcpy(idx->d_sb, h.sb_base.data(), 32, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(idx->d_counts, h.counts, 32, cudaMemcpyHostToDevice) != cudaSuccess)
        return cu_fail("cudaMemcpy");
    rc = finish_device_setup(idx.get(), opts);
    if (rc) {
        fmsi_gpu_index_free(idx.release());
        return rc;
    }
    *out = idx.release();
    return FMSI_GPU_OK;
}

int fmsi_gpu_index_save(const fmsi_gpu_index *idx, const char *prefix) {
    if (!idx || !prefix) return fail(FMSI_GPU_ERR_ARG, "null argument");
    if (idx->plane_hi.empty()) return fail(FMSI_GPU_ERR_ARG, "fmsi_gpu_index_save: index was not made by fmsi_gpu_index_build");
    try {
        const HostIndex &h = idx->meta;
        const uint64_t N = h.n, nw = (N + 63) >> 6;
        const std::string base = std::string(prefix) + ".fmsi";
        // counts = {1, #A+1, #A+#C+1, ...} -> |ac| = counts[2], |gt| = N - counts[2] (fms_index.h:434-451).
        // ac / gt = the low plane compacted to the A/C/$ slots and to the G/T slots, in row order: per 1024-block
        // granule the slot counts, a prefix sum over granules, then all host threads append their ranges.
        BitVec ac, gt;
        ac.resize_bits(h.counts[2]);
        gt.resize_bits(N - h.counts[2]);
        const uint64_t kGran = 1024, ngran = (nw + kGran - 1) / kGran;
        auto block_hi = [&](uint64_t b) {
            const unsigned valid = (unsigned)std::min<uint64_t>(64, N - b * 64);
            return idx->plane_hi[b] & (valid == 64 ? ~0ull : ((1ull << valid) - 1));
        };
        auto block_valid = [&](uint64_t b) { return (uint64_t)std::min<uint64_t>(64, N - b * 64); };
        std::vector<uint64_t> g_start(ngran + 1, 0);  // G/T slots before granule g
        parallel_ranges(ngran, 1, [&](uint64_t g0, uint64_t g1) {
            for (uint64_t g = g0; g < g1; ++g) {
                uint64_t c = 0;
                for (uint64_t b = g * kGran; b < std::min(nw, (g + 1) * kGran); ++b) c += (uint64_t)__builtin_popcountll(block_hi(b));
                g_start[g + 1] = c;
            }
        });
        for (uint64_t g = 0; g < ngran; ++g) g_start[g + 1] += g_start[g];
        if (g_start[ngran] != gt.nbits) throw std::runtime_error("plane sizes inconsistent with counts");
        parallel_ranges(ngran, 1, [&](uint64_t g0, uint64_t g1) {
            uint64_t gp = g_start[g0], ap = g0 * kGran * 64 - gp;
            for (uint64_t b = g0 * kGran; b < std::min(nw, g1 * kGran); ++b) {
                const uint64_t hi = block_hi(b), lo = idx->plane_lo[b];
                const uint64_t vmask = block_valid(b) == 64 ? ~0ull : ((1ull << block_valid(b)) - 1);
                uint64_t sel = ~hi & vmask, bits = 0;
                unsigned cnt = 0;
                while (sel) {  // A/C/$ slots in order
                    bits |= ((lo >> __builtin_ctzll(sel)) & 1ull) << cnt++;
                    sel &= sel - 1;
                }
                ac.set_int_atomic(ap, bits, cnt);
                ap += cnt;
                sel = hi;
                bits = 0;
                cnt = 0;
                while (sel) {
                    bits |= ((lo >> __builtin_ctzll(sel)) & 1ull) << cnt++;
                    sel &= sel - 1;
                }
                gt.set_int_atomic(gp, bits, cnt);
                gp += cnt;
            }
        });
        // the five files are written concurrently (the mask's RRR<63> coder runs on all threads first)
        BitVec mask;
        mask.resize_bits(N);
        std::memcpy(mask.w.data(), idx->plane_mask.data(), nw * 8);
        const RrrFile rrr = rrr_encode(mask);
        mask = BitVec();
        std::vector<std::thread> writers;
        std::string werr[5];
        auto guarded = [&](int slot, std::function<void()> fn) {
            writers.emplace_back([&werr, slot, fn] {
                try {
                    fn();
                } catch (const std::exception &e) {
                    werr[slot] = e.what();
                }
            });
        };
        auto write_plane = [&](const std::string &path, const std::vector<uint64_t> &plane) {
            ByteWriter w(path);
            w.u64(N);
            w.bytes(plane.data(), nw * 8);
        };
        guarded(0, [&] { write_plane(base + ".ac_gt", idx->plane_hi); });
        guarded(1, [&] {
            ByteWriter w(base + ".ac");
            w.bitvec(ac);
        });
        guarded(2, [&] {
            ByteWriter w(base + ".gt");
            w.bitvec(gt);
        });
        guarded(3, [&] { write_rrr(base + ".mask", rrr); });
        if (h.has_klcp) guarded(4, [&] { write_plane(base + ".klcp", idx->plane_klcp); });
        else std::remove((base + ".klcp").c_str());
        for (auto &t : writers) t.join();
        for (const std::string &e : werr)
            if (!e.empty()) throw std::runtime_error(e);
        FILE *f = std::fopen((base + ".misc").c_str(), "w");
        if (!f) throw std::runtime_error("cannot create " + base + ".misc");
        std::fprintf(f, "%llu\n%llu\n%llu\n%llu\n%llu\n%d\n", (unsigned long long)h.dollar, (unsigned long long)h.counts[0],
                     (unsigned long long)h.counts[1], (unsigned long long)h.counts[2], (unsigned long long)h.counts[3], h.k);
        std::fclose(f);
    } catch (const std::exception &e) {
        return fail(FMSI_GPU_ERR_IO, e.what());
    }
    return FMSI_GPU_OK;
}

int fmsi_gpu_index_free(fmsi_gpu_index *idx) {
    if (!idx) return FMSI_GPU_OK;
    cudaSetDevice(idx->device);
    for (void *p : {idx->d_rank, idx->d_aux, idx->d_table, idx->d_sb, idx->d_counts, idx->d_rows, idx->d_fbuckets, idx->d_frows, idx->d_fids,
                    idx->d_multi, idx->d_ldir, idx->d_lrows, idx->d_user_bytes, (void *)idx->d_probes, (void *)idx->user.ctr, idx->user.ovf})
        if (p) cudaFree(p);
    if (idx->aux_stream) cudaStreamDestroy(idx->aux_stream);
    for (cudaEvent_t ev : idx->piece_events) cudaEventDestroy(ev);
    for (cudaEvent_t ev : idx->span_done)
        if (ev) cudaEventDestroy(ev);
    for (auto &s : idx->slots) {
        for (void *p : {s.d_in, s.d_out, s.d_aux, (void *)s.ls.ctr, s.ls.ovf})
            if (p) cudaFree(p);
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    delete idx;
    return FMSI_GPU_OK;
}

int fmsi_gpu_index_get_info(const fmsi_gpu_index *idx, fmsi_gpu_index_info *info) {
    if (!idx || !info) return fail(FMSI_GPU_ERR_ARG, "null argument");
    std::memset(info, 0, sizeof(*info));
    info->n_bwt = idx->meta.n;
    for (int c = 0; c < 4; ++c) info->counts[c] = idx->meta.counts[c];
    info->dollar_position = idx->meta.dollar;
    info->mask_ones = idx->meta.mask_ones;
    info->hbm_bytes = idx->hbm_bytes;
    info->k = idx->meta.k;
    info->has_klcp = idx->meta.has_klcp;
    info->prefix_t = (int32_t)idx->dev.t;
    info->dict = idx->fold.enabled ? 2 : (int32_t)idx->dict.enabled;
    info->dict_t = idx->fold.enabled ? (int32_t)idx->fold.t : idx->dict.enabled ? (int32_t)idx->dev.t : 0;
    info->wide = idx->wide;
    info->device = idx->device;
    info->multistep = (int32_t)idx->dev.multi_m;
    info->fold_ids = idx->fold.enabled && idx->fold.ids ? 1 : 0;
    info->locality = idx->loc.enabled ? (int32_t)idx->loc.g.m : 0;
    return FMSI_GPU_OK;
}

// ---------------------------------------------------------------------------------- probes
#define PROBE_PROLOGUE()                                                  \
    if (!idx) return fail(FMSI_GPU_ERR_ARG, "null index");                \
    if (n == 0) return FMSI_GPU_OK;                                       \
    CU(cudaSetDevice(idx->device))

int fmsi_gpu_rank(fmsi_gpu_index *idx, const uint64_t *i, const uint8_t *c, size_t n, uint64_t *out) {
    PROBE_PROLOGUE();
    for (size_t q = 0; q < n; ++q)
        if (i[q] > idx->meta.n || c[q] > 3) return fail(FMSI_GPU_ERR_ARG, "rank argument out of range");
    DevBuf di, dc, dout;
    int rc;
    if ((rc = di.put(i, n * 8)) || (rc = dc.put(c, n)) || (rc = dout.alloc(n * 8))) return rc;
    if (idx->wide)
        probe_rank_kernel<true><<<blocks_for(n), 256>>>(idx->dev, di.as<u64>(), dc.as<unsigned char>(), n, (const u64 *)idx->d_counts, dout.as<u64>());
    else
        probe_rank_kernel<false><<<blocks_for(n), 256>>>(idx->dev, di.as<u64>(), dc.as<unsigned char>(), n, (const u64 *)idx->d_counts, dout.as<u64>());
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    return dout.get(out, n * 8);
}

int fmsi_gpu_update_range(fmsi_gpu_index *idx, uint64_t *i, uint64_t *j, const uint8_t *c, size_t n) {
    PROBE_PROLOGUE();
    for (size_t q = 0; q < n; ++q)
        if (i[q] > idx->meta.n || j[q] > idx->meta.n || c[q] > 3) return fail(FMSI_GPU_ERR_ARG, "update_range argument out of range");
    DevBuf di, dj, dc;
    int rc;
    if ((rc = di.put(i, n * 8)) || (rc = dj.put(j, n * 8)) || (rc = dc.put(c, n))) return rc;
    if (idx->wide) probe_update_range_kernel<true><<<blocks_for(n), 256>>>(idx->dev, di.as<u64>(), dj.as<u64>(), dc.as<unsigned char>(), n);
    else probe_update_range_kernel<false><<<blocks_for(n), 256>>>(idx->dev, di.as<u64>(), dj.as<u64>(), dc.as<unsigned char>(), n);
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    if ((rc = di.get(i, n * 8)) || (rc = dj.get(j, n * 8))) return rc;
    return FMSI_GPU_OK;
}

int fmsi_gpu_extend_range_with_klcp(fmsi_gpu_index *idx, uint64_t *i, uint64_t *j, size_t n) {
    PROBE_PROLOGUE();
    if (!idx->meta.has_klcp) return fail(FMSI_GPU_ERR_KLCP, "index loaded without kLCP");
    for (size_t q = 0; q < n; ++q)
        if (i[q] == 0 || i[q] >= j[q] || j[q] > idx->meta.n) return fail(FMSI_GPU_ERR_ARG, "extend_range_with_klcp needs 1 <= i < j <= N");
    DevBuf di, dj;
    int rc;
    if ((rc = di.put(i, n * 8)) || (rc = dj.put(j, n * 8))) return rc;
    probe_extend_kernel<<<blocks_for(n), 256>>>(idx->dev, di.as<u64>(), dj.as<u64>(), n);
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    if ((rc = di.get(i, n * 8)) || (rc = dj.get(j, n * 8))) return rc;
    return FMSI_GPU_OK;
}

int fmsi_gpu_get_range_with_pattern(fmsi_gpu_index *idx, const uint64_t *kmers, int k, size_t n,
                                    int use_table, uint64_t *sa_start, uint64_t *sa_end) {
    PROBE_PROLOGUE();
    if (k < 1 || k > 32) return fail(FMSI_GPU_ERR_K, "k must be in [1, 32]");
    DevBuf dk, di, dj;
    int rc;
    if ((rc = dk.put(kmers, n * 8)) || (rc = di.alloc(n * 8)) || (rc = dj.alloc(n * 8))) return rc;
    if (idx->wide) probe_get_range_kernel<true><<<blocks_for(n), 256>>>(idx->dev, dk.as<u64>(), (u32)k, n, use_table, di.as<u64>(), dj.as<u64>());
    else probe_get_range_kernel<false><<<blocks_for(n), 256>>>(idx->dev, dk.as<u64>(), (u32)k, n, use_table, di.as<u64>(), dj.as<u64>());
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    if ((rc = di.get(sa_start, n * 8)) || (rc = dj.get(sa_end, n * 8))) return rc;
    return FMSI_GPU_OK;
}

int fmsi_gpu_infer_presence(fmsi_gpu_index *idx, const uint64_t *sa_start, const uint64_t *sa_end,
                            size_t n, int maximized_ones, int8_t *out) {
    PROBE_PROLOGUE();
    for (size_t q = 0; q < n; ++q)
        if (sa_start[q] > sa_end[q] || sa_end[q] > idx->meta.n) return fail(FMSI_GPU_ERR_ARG, "interval out of range");
    DevBuf di, dj, dout;
    int rc;
    if ((rc = di.put(sa_start, n * 8)) || (rc = dj.put(sa_end, n * 8)) || (rc = dout.alloc(n))) return rc;
    probe_presence_kernel<<<blocks_for(n), 256>>>(idx->dev, di.as<u64>(), dj.as<u64>(), n, maximized_ones, dout.as<signed char>());
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    return dout.get(out, n);
}

int fmsi_gpu_kmer_order_if_present(fmsi_gpu_index *idx, const uint64_t *sa_start,
                                   const uint64_t *sa_end, size_t n, int64_t *out) {
    PROBE_PROLOGUE();
    for (size_t q = 0; q < n; ++q)
        if (sa_start[q] > sa_end[q] || sa_end[q] > idx->meta.n) return fail(FMSI_GPU_ERR_ARG, "interval out of range");
    DevBuf di, dj, dout;
    int rc;
    if ((rc = di.put(sa_start, n * 8)) || (rc = dj.put(sa_end, n * 8)) || (rc = dout.alloc(n * 8))) return rc;
    probe_order_kernel<<<blocks_for(n), 256>>>(idx->dev, di.as<u64>(), dj.as<u64>(), n, dout.as<long long>());
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    return dout.get(out, n * 8);
}

// ---------------------------------------------------------------------------------- hot path
}  // extern "C"

namespace {

bool to_genf(const fmsi_gpu_function *f, GenF &g) {
    if (!f || f->kind < FMSI_GPU_F_OR || f->kind > FMSI_GPU_F_RANGE || f->r < 0 || f->s < 0) return false;
    g.kind = f->kind;
    g.r = (unsigned)f->r;
    g.s = (unsigned)f->s;
    return true;
}

int query_kmers_impl(fmsi_gpu_index *idx, int mode, int output, int strands, const GenF *gf,
                     const uint64_t *kmers, size_t n, int k, void *results, int mem, void *stream) {
    if (!idx) return fail(FMSI_GPU_ERR_ARG, "null index");
    if (k < 1 || k > 32) return fail(FMSI_GPU_ERR_K, "k must be in [1, 32] for packed k-mers");
    const bool bits = output == FMSI_GPU_OUT_PRESENCE_BITS;  // presence bytes, then packed 8 per byte on the device
    if (bits) output = FMSI_GPU_OUT_PRESENCE;
    if ((mode != FMSI_GPU_MODE_OR && mode != FMSI_GPU_MODE_ALL && !(mode == FMSI_GPU_MODE_GENERAL_ && gf)) || (output != FMSI_GPU_OUT_PRESENCE && output != FMSI_GPU_OUT_ORDERS) ||
        (strands != FMSI_GPU_STRANDS_LAZY && strands != FMSI_GPU_STRANDS_BOTH) || (bits && strands != FMSI_GPU_STRANDS_LAZY))
        return fail(FMSI_GPU_ERR_ARG, "bad mode/output/strands");
    if (n == 0) return FMSI_GPU_OK;
    if (!kmers || !results) return fail(FMSI_GPU_ERR_ARG, "null buffer");
    CU(cudaSetDevice(idx->device));
    if (output == FMSI_GPU_OUT_ORDERS && (u32)k == idx->fold.k) ensure_fold_ids(idx);
    const DevIndex d = dev_for_k(idx, k);
    const size_t rbytes = result_bytes(output, strands);
    int rc;

    if (mem == FMSI_GPU_MEM_DEVICE) {
        cudaStream_t st = (cudaStream_t)stream;
        if (!bits) return dispatch_query(idx, d, mode, output, strands, kmers, n, results, idx->user, st, gf);
        if ((rc = ensure(&idx->d_user_bytes, &idx->user_bytes_cap, n))) return rc;
        if ((rc = dispatch_query(idx, d, mode, output, strands, kmers, n, idx->d_user_bytes, idx->user, st, gf))) return rc;
        pack_presence_bits_kernel<<<blocks_for((n + 7) / 8), 256, 0, st>>>((const unsigned char *)idx->d_user_bytes, (u64)n, (unsigned char *)results);
        CU(cudaGetLastError());
        g_launches.fetch_add(1);
        return FMSI_GPU_OK;
    }
    if (mem != FMSI_GPU_MEM_HOST) return fail(FMSI_GPU_ERR_ARG, "bad mem");

    // Host buffers: double-buffered batches so H2D, kernel and D2H of neighbouring batches overlap. The call is
    // bound by the H2D copies (8 B per k-mer over PCIe); what is not overlapped is the kernel + D2H of the last
    // batch, so a call is cut into ~16 batches (at least 2 Mi k-mers each: 16 MB copies still run at link speed;
    // a multiple of 64 k-mers, so that bit-packed results of a batch start on a byte boundary).
    const size_t batch = (std::min(kBatchKmers, std::max<size_t>((size_t)1 << 21, (n + 15) / 16)) + 63) & ~size_t(63);
    size_t done = 0;
    int b = 0;
    while (done < n) {
        Slot &s = idx->slots[b % kSlots];
        const size_t m = std::min(batch, n - done);
        CU(cudaEventSynchronize(s.done));
        if ((rc = ensure(&s.d_in, &s.in_cap, m * 8)) || (rc = ensure(&s.d_out, &s.out_cap, m * rbytes))) return rc;
        if (bits && (rc = ensure(&s.d_aux, &s.aux_cap, (m + 7) / 8))) return rc;
        CU(cudaMemcpyAsync(s.d_in, kmers + done, m * 8, cudaMemcpyHostToDevice, s.stream));
        if ((rc = dispatch_query(idx, d, mode, output, strands, (const u64 *)s.d_in, m, s.d_out, s.ls, s.stream, gf))) return rc;
        if (bits) {
            pack_presence_bits_kernel<<<blocks_for((m + 7) / 8), 256, 0, s.stream>>>((const unsigned char *)s.d_out, (u64)m, (unsigned char *)s.d_aux);
            CU(cudaGetLastError());
            g_launches.fetch_add(1);
            CU(cudaMemcpyAsync((char *)results + done / 8, s.d_aux, (m + 7) / 8, cudaMemcpyDeviceToHost, s.stream));
        } else {
            CU(cudaMemcpyAsync((char *)results + done * rbytes, s.d_out, m * rbytes, cudaMemcpyDeviceToHost, s.stream));
        }
        CU(cudaEventRecord(s.done, s.stream));
        done += m;
        ++b;
    }
    for (auto &s : idx->slots) CU(cudaStreamSynchronize(s.stream));
    return FMSI_GPU_OK;
}

int query_chunks_impl(fmsi_gpu_index *idx, int mode, int output, int strands, int streaming, const GenF *gf,
                      const void *text, int text_format, size_t n_bases, const uint64_t *chunk_off,
                      const uint32_t *chunk_len, const uint64_t *res_off, size_t n_chunks,
                      size_t n_results, int k, void *results, int mem, void *stream) {
    if (!idx) return fail(FMSI_GPU_ERR_ARG, "null index");
    if (k < 1 || k > FMSI_GPU_MAX_K) return fail(FMSI_GPU_ERR_K, "k must be in [1, FMSI_GPU_MAX_K]");
    const bool bits = output == FMSI_GPU_OUT_PRESENCE_BITS;  // presence bytes, packed 8 per byte once all are in
    if (bits) output = FMSI_GPU_OUT_PRESENCE;
    if ((mode != FMSI_GPU_MODE_OR && mode != FMSI_GPU_MODE_ALL && !(mode == FMSI_GPU_MODE_GENERAL_ && gf)) || (output != FMSI_GPU_OUT_PRESENCE && output != FMSI_GPU_OUT_ORDERS) ||
        (strands != FMSI_GPU_STRANDS_LAZY && strands != FMSI_GPU_STRANDS_BOTH) || (bits && strands != FMSI_GPU_STRANDS_LAZY))
        return fail(FMSI_GPU_ERR_ARG, "bad mode/output/strands");
    if (text_format != FMSI_GPU_TEXT_ASCII && text_format != FMSI_GPU_TEXT_PACKED2) return fail(FMSI_GPU_ERR_ARG, "bad text format");
    if (streaming && !idx->meta.has_klcp) return fail(FMSI_GPU_ERR_KLCP, "kLCP array was not loaded for the given index");
    // the kLCP array describes (k-1)-mers of the index's own k (construct_klcp, fms_index.h:357-385)
    if (streaming && k <= 32 && k != idx->meta.k) return fail(FMSI_GPU_ERR_K, "streaming queries need k equal to the index's k");
    if (n_chunks == 0 || n_results == 0) return FMSI_GPU_OK;
    // reads mode (fmsi_gpu_query_reads_packed): chunk_off = read_off[n_reads + 1], n_chunks = n_reads, no chunk arrays —
    // the device cuts the reads into chunks itself
    const bool reads_mode = chunk_len == nullptr && res_off == nullptr;
    if (!text || !chunk_off || (!reads_mode && (!chunk_len || !res_off)) || !results) return fail(FMSI_GPU_ERR_ARG, "null buffer");
    CU(cudaSetDevice(idx->device));
    if (output == FMSI_GPU_OUT_ORDERS && (u32)k == idx->fold.k) ensure_fold_ids(idx);
    if (output == FMSI_GPU_OUT_PRESENCE && mode != FMSI_GPU_MODE_GENERAL_) ensure_loc(idx, k, n_results, false);
    const DevIndex d = dev_for_k(idx, k);
    const size_t rbytes = result_bytes(output, strands);
    const bool on_host = mem == FMSI_GPU_MEM_HOST;
    if (!on_host && mem != FMSI_GPU_MEM_DEVICE) return fail(FMSI_GPU_ERR_ARG, "bad mem");
    const bool packed_in = text_format == FMSI_GPU_TEXT_PACKED2;
    const char *bases = packed_in ? nullptr : (const char *)text;
    const u64 *words_in = packed_in ? (const u64 *)text : nullptr;
    const size_t n_words_in = (n_bases + 31) / 32;

    Slot &s = idx->slots[0];
    cudaStream_t st = on_host ? s.stream : (cudaStream_t)stream;
    const char *d_bases = bases;
    const u64 *d_off = chunk_off, *d_res = res_off;
    const u32 *d_len = chunk_len;
    void *d_results = results;
    char *reads_scratch = nullptr;
    int rc;
    // scratch: [packed bases | (host mode) bases, offsets, lens, res_off | packed k-mers (non-streaming)]
    // Per-k-mer strand values do not depend on how they are computed (kLCP interval reuse is only a
    // shortcut), so when the dictionary tier is resident streamed chunks take it too: ~2 requests per
    // k-mer instead of an aux probe + an LF-step per k-mer and strand (profiles/r01d_modes_*.json).
    // k > 32: queries are start positions into the packed text (longk_kernels.cuh), with or without -S.
    const bool longk = k > 32;
    const bool via_loc = !longk && loc_answers(idx, k, mode, output);
    const bool via_kmers = longk || !streaming || via_loc || (!idx->wide && ((idx->fold.enabled && (u32)k == idx->fold.k && (output != FMSI_GPU_OUT_ORDERS || idx->fold.ids)) ||
                                                         (idx->dict.enabled && (u32)k == idx->dict.k && d.t && n_results < (1ull << 32))));
    const size_t n_words = n_words_in + 4;
    size_t aux_need = n_words * 8;
    const size_t kmers_off = aux_need;
    if (via_kmers) aux_need += n_results * 8;
    // Host mode, large ordered inputs: the text goes up in pieces on one stream while the chunks that are complete
    // run (and their results come back) on two others, so H2D, kernels and D2H overlap within the call.
    static const size_t kPiece = [] {  // bases per piece (a multiple of 32); $FMSI_GPU_PIECE_MIB: design experiment switch
        const char *e = std::getenv("FMSI_GPU_PIECE_MIB");
        const size_t mib = e ? (size_t)std::atoll(e) : 0;
        return (mib >= 1 && mib <= 1024 ? mib : 8) << 20;
    }();
    const bool pipelined = on_host && n_bases > 2 * kPiece;
    // reads mode: the streaming kernel takes chunks of <= 64 k-mers, every other path one chunk per read
    const size_t n_reads = reads_mode ? n_chunks : 0;
    // (the minimizer-bucketed tier walks tiles of 32 k-mers: reads cut into chunks of 32 give it full tiles)
    const u32 read_max_kmers = !reads_mode ? 0u : via_loc ? 32u : !via_kmers ? FMSI_GPU_MAX_STREAM_KMERS : 0u;
    if (reads_mode) n_chunks = read_max_kmers ? n_reads + n_results / read_max_kmers + 1 : n_reads;  // capacity of the device chunk arrays
    auto bad_stream_chunk = [&](size_t c) { return chunk_len[c] < (u32)k || chunk_len[c] - (u32)k + 1 > FMSI_GPU_MAX_STREAM_KMERS; };
    const char *kBadStreamChunk = "streaming chunks must hold between 1 and FMSI_GPU_MAX_STREAM_KMERS k-mers";
    if (on_host) {
        for (auto &sl : idx->slots) CU(cudaEventSynchronize(sl.done));
        if (pipelined && !idx->aux_stream) CU(cudaStreamCreateWithFlags(&idx->aux_stream, cudaStreamNonBlocking));
        const size_t text_stage = packed_in ? 0 : ((n_bases + 15) & ~size_t(15));
        const size_t in_need = text_stage + n_chunks * (8 + 8 + 4) + 64 + (reads_mode ? reads_scratch_bytes(n_reads) : 0);
        // bit-packed output: the packed bits sit behind the byte results
        const size_t out_need = n_results * rbytes + (bits ? ((n_results + 7) / 8 + 64) : 0);
        if ((rc = ensure(&s.d_in, &s.in_cap, in_need)) || (rc = ensure(&s.d_out, &s.out_cap, out_need))) return rc;
        char *p = (char *)s.d_in;
        d_bases = p;
        p += text_stage;
        d_off = (const u64 *)p;
        p += n_chunks * 8;
        d_res = (const u64 *)p;
        p += n_chunks * 8;
        d_len = (const u32 *)p;
        p += ((n_chunks * 4 + 63) & ~size_t(63));
        reads_scratch = p;
        d_results = s.d_out;
    } else {
        if (bits) {
            if ((rc = ensure(&idx->d_user_bytes, &idx->user_bytes_cap, n_results))) return rc;
            d_results = idx->d_user_bytes;
        }
        if (reads_mode) {  // device-mode reads: the chunk arrays live in the slot's input scratch
            const size_t need = n_chunks * (8 + 8 + 4) + 128 + reads_scratch_bytes(n_reads);
            if ((rc = ensure(&s.d_in, &s.in_cap, need))) return rc;
            char *p = (char *)s.d_in;
            d_off = (const u64 *)p;
            p += n_chunks * 8;
            d_res = (const u64 *)p;
            p += n_chunks * 8;
            d_len = (const u32 *)p;
            p += ((n_chunks * 4 + 63) & ~size_t(63));
            reads_scratch = p;
        }
    }
    // reads -> device chunk arrays, on stream q (reads mode): read offsets up (host mode), counts, two prefix sums, expansion
    auto expand_reads = [&](cudaStream_t q) -> int {
        char *p = reads_scratch;
        u64 *d_roff = (u64 *)p;
        p += (n_reads + 1) * 8;
        u64 *d_rbase = (u64 *)p;
        p += (n_reads + 1) * 8;
        u64 *d_cbase = (u64 *)p;
        p += (n_reads + 1) * 8;
        u32 *d_nk = (u32 *)p;
        p += ((n_reads * 4 + 63) & ~size_t(63));
        u32 *d_nch = (u32 *)p;
        p += ((n_reads * 4 + 63) & ~size_t(63));
        void *d_tmp = (void *)(((uintptr_t)p + 255) & ~(uintptr_t)255);
        size_t tmp_bytes = reads_scan_temp_bytes(n_reads);
        const u64 *roff_dev = chunk_off;
        if (on_host) {
            CU(cudaMemcpyAsync(d_roff, chunk_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, q));
            roff_dev = d_roff;
        }
        read_counts_kernel<<<blocks_for(n_reads), 256, 0, q>>>(roff_dev, (u64)n_reads, (u64)n_bases, (u32)k, read_max_kmers, d_nk, d_nch);
        CU(cudaGetLastError());
        CU(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, CountIter(d_nk, U32ToU64()), d_rbase, (int)n_reads, q));
        tmp_bytes = reads_scan_temp_bytes(n_reads);
        CU(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, CountIter(d_nch, U32ToU64()), d_cbase, (int)n_reads, q));
        expand_reads_kernel<<<blocks_for(n_reads), 256, 0, q>>>(roff_dev, (u64)n_reads, (u64)n_bases, (u32)k, read_max_kmers, d_nk, d_rbase, d_cbase,
                                                                 (u64 *)d_off, (u32 *)d_len, (u64 *)d_res);
        CU(cudaGetLastError());
        g_launches.fetch_add(4);
        return FMSI_GPU_OK;
    };
    // host view of a read (reads mode, host buffers): results and device chunks it yields; false if malformed
    auto read_counts = [&](size_t r, u64 &nk, u64 &nch) -> bool {
        const u64 a = chunk_off[r], b = chunk_off[r + 1];
        if (b < a || b > n_bases) return false;
        const u64 len = b - a;
        nk = len >= (u64)k ? len - (u64)k + 1 : 0;
        nch = read_max_kmers ? (nk + read_max_kmers - 1) / read_max_kmers : 1;
        return true;
    };
    // chunk metadata [c0, c1) to the device (host mode)
    auto upload_chunks = [&](size_t c0, size_t c1, cudaStream_t q) -> int {
        if (c1 <= c0) return FMSI_GPU_OK;
        CU(cudaMemcpyAsync((void *)(d_off + c0), chunk_off + c0, (c1 - c0) * 8, cudaMemcpyHostToDevice, q));
        CU(cudaMemcpyAsync((void *)(d_res + c0), res_off + c0, (c1 - c0) * 8, cudaMemcpyHostToDevice, q));
        CU(cudaMemcpyAsync((void *)(d_len + c0), chunk_len + c0, (c1 - c0) * 4, cudaMemcpyHostToDevice, q));
        return FMSI_GPU_OK;
    };
    if ((rc = ensure(&s.d_aux, &s.aux_cap, aux_need))) return rc;
    u64 *d_packed = (u64 *)s.d_aux;
    u64 *d_slots = (u64 *)((char *)s.d_aux + kmers_off);  // k-mers (or, for k > 32, start positions) of the result slots

    // Packed words [w0, w1) of the text on stream q: ASCII text is uploaded (host mode) and packed on the device,
    // 2-bit text (FMSI_GPU_TEXT_PACKED2: the same layout) is copied as it is; words past the text are zeroed.
    auto stage_words = [&](size_t w0, size_t w1, cudaStream_t q) -> int {
        if (w1 <= w0) return FMSI_GPU_OK;
        if (packed_in) {
            const size_t e = std::min(w1, n_words_in);
            if (e > w0) CU(cudaMemcpyAsync(d_packed + w0, words_in + w0, (e - w0) * 8, on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, q));
            if (w1 > std::max(e, w0)) CU(cudaMemsetAsync(d_packed + std::max(e, w0), 0, (w1 - std::max(e, w0)) * 8, q));
            return FMSI_GPU_OK;
        }
        if (on_host) {
            const size_t b0 = 32 * w0, b1 = std::min(n_bases, 32 * w1);
            if (b1 > b0) CU(cudaMemcpyAsync((char *)d_bases + b0, bases + b0, b1 - b0, cudaMemcpyHostToDevice, q));
        }
        pack_bases_kernel<<<blocks_for(w1 - w0), 256, 0, q>>>(d_bases + 32 * w0, (u64)(n_bases - std::min(n_bases, 32 * w0)), d_packed + w0, (u64)(w1 - w0));
        CU(cudaGetLastError());
        g_launches.fetch_add(1);
        return FMSI_GPU_OK;
    };

    // The queries of chunks [c0, c1) = result slots [r0, r1), on stream `q` with launch scratch `ls`.
    auto copy_back = [&](size_t r0, size_t r1, cudaStream_t q) -> int {
        if (r1 > r0 && !bits) CU(cudaMemcpyAsync((char *)results + r0 * rbytes, (char *)d_results + r0 * rbytes, (r1 - r0) * rbytes, cudaMemcpyDeviceToHost, q));
        return FMSI_GPU_OK;
    };
    // bit-packed output: all byte results are in d_results (stream q ordered after them) -> bits -> caller
    auto finish_bits = [&](cudaStream_t q) -> int {
        if (!bits) return FMSI_GPU_OK;
        unsigned char *d_bits = on_host ? (unsigned char *)d_results + ((n_results + 63) & ~size_t(63)) : (unsigned char *)results;
        pack_presence_bits_kernel<<<blocks_for((n_results + 7) / 8), 256, 0, q>>>((const unsigned char *)d_results, (u64)n_results, d_bits);
        CU(cudaGetLastError());
        g_launches.fetch_add(1);
        if (on_host) CU(cudaMemcpyAsync(results, d_bits, (n_results + 7) / 8, cudaMemcpyDeviceToHost, q));
        return FMSI_GPU_OK;
    };
    // Bit-packed output of a pipelined call: once span `span` (results [.., r1)) has been launched on q, every result byte
    // that is complete — [bits_done, r1 / 8), all of them after the last span — is packed and sent back on q, behind
    // that span's kernel, so only the last span's bits return after the queries. The first byte may hold results of
    // the previous span, which ran on the other query stream: q waits for that span's event.
    size_t bits_done = 0;
    auto span_bits = [&](size_t span, size_t r1, bool last_span, cudaStream_t q) -> int {
        if (!bits) return FMSI_GPU_OK;
        for (int e = 0; e < 2; ++e)
            if (!idx->span_done[e]) CU(cudaEventCreateWithFlags(&idx->span_done[e], cudaEventDisableTiming));
        const size_t byte_hi = last_span ? (n_results + 7) / 8 : r1 / 8;
        if (byte_hi > bits_done) {
            if (span > 0) CU(cudaStreamWaitEvent(q, idx->span_done[(span - 1) & 1], 0));
            unsigned char *d_bits = (unsigned char *)d_results + ((n_results + 63) & ~size_t(63));
            const size_t q0 = bits_done * 8, q1 = std::min(byte_hi * 8, n_results);
            pack_presence_bits_kernel<<<blocks_for(byte_hi - bits_done), 256, 0, q>>>((const unsigned char *)d_results + q0, (u64)(q1 - q0), d_bits + bits_done);
            CU(cudaGetLastError());
            g_launches.fetch_add(1);
            CU(cudaMemcpyAsync((char *)results + bits_done, d_bits + bits_done, byte_hi - bits_done, cudaMemcpyDeviceToHost, q));
            bits_done = byte_hi;
        }
        CU(cudaEventRecord(idx->span_done[span & 1], q));
        return FMSI_GPU_OK;
    };
    auto run_span = [&](size_t c0, size_t c1, size_t r0, size_t r1, cudaStream_t q, LaunchScratch &ls) -> int {
        if (c1 <= c0 || r1 <= r0) return FMSI_GPU_OK;
        void *out_span = (char *)d_results + r0 * rbytes;
        if (longk) {
            extract_starts_kernel<<<slot_blocks(r1 - r0), 256, 0, q>>>(d_off, d_len, d_res, (u64)c0, (u64)c1, (u64)r0, (u64)r1, (u32)k, (u64)n_bases, d_slots);
            CU(cudaGetLastError());
            int e = dispatch_long(idx->wide, idx->sm_count, d, mode, output, strands, gf, d_packed, d_slots + r0, r1 - r0, out_span, ls.ctr, q);
            if (e) return fail(FMSI_GPU_ERR_CUDA, std::string("long-k kernel launch: ") + cudaGetErrorString((cudaError_t)e));
            g_launches.fetch_add(2);
        } else if (via_loc) {
            // neighbouring k-mers of a text share their bucket in the minimizer-bucketed dictionary (loc.cuh): one lane per
            // chunk (reads mode: per read), results written at their slots of d_results
            int e = launch_loc(idx, mode, strands, d_packed, (u64)n_bases, d_off + c0, d_len + c0, d_res + c0, c1 - c0, d_results, ls, q);
            if (e) return e;
        } else if (via_kmers && reads_mode && fold_answers(idx, k, mode, output)) {
            // the dictionary kernel cuts its k-mers out of the reads itself (one chunk per read: d_off = first base,
            // d_res = first result slot of every read)
            const ReadSrc rs{d_packed, d_off, d_res, (u64)n_reads, (u64)r0};
            int e = dispatch_query(idx, d, mode, output, strands, nullptr, r1 - r0, out_span, ls, q, gf, rs);
            if (e) return e;
        } else if (via_kmers) {
            extract_kmers_kernel<<<slot_blocks(r1 - r0), 256, 0, q>>>(d_packed, d_off, d_len, d_res, (u64)c0, (u64)c1, (u64)r0, (u64)r1, (u32)k, (u64)n_bases, d_slots);
            CU(cudaGetLastError());
            g_launches.fetch_add(1);
            int e = dispatch_query(idx, d, mode, output, strands, d_slots + r0, r1 - r0, out_span, ls, q, gf);
            if (e) return e;
        } else {
            int e = dispatch_stream(idx->wide, idx->sm_count, d, mode, output, strands, d_packed, (u64)n_bases, d_off + c0, d_len + c0, d_res + c0, c1 - c0,
                                    d_results, ls.ctr, q, probe_ctr(idx));
            if (e) return fail(FMSI_GPU_ERR_CUDA, std::string("streaming kernel launch: ") + cudaGetErrorString((cudaError_t)e));
            g_launches.fetch_add(1);
        }
        return FMSI_GPU_OK;
    };

    auto single_batch = [&]() -> int {
        size_t chunks_used = n_chunks;
        if (reads_mode) {
            if (on_host) {
                u64 R = 0, C = 0;
                for (size_t r = 0; r < n_reads; ++r) {
                    u64 nk, nch;
                    if (!read_counts(r, nk, nch)) return fail(FMSI_GPU_ERR_ARG, "read offsets must be non-decreasing and end inside the text");
                    R += nk;
                    C += nch;
                }
                if (R != n_results) return fail(FMSI_GPU_ERR_ARG, "n_results does not match the reads");
                chunks_used = (size_t)C;
            }
            // device-mode streamed reads: the chunk count is only known on the device; the unused tail of the chunk
            // arrays (sized for the worst case) holds empty chunks, which the kernel skips — no synchronisation
            if (!on_host && read_max_kmers) CU(cudaMemsetAsync((void *)d_len, 0, n_chunks * 4, st));
            int e = expand_reads(st);
            if (e) return e;
        } else if (on_host) {
            for (size_t c = 0; c < n_chunks; ++c) {
                if (chunk_off[c] + chunk_len[c] > n_bases) return fail(FMSI_GPU_ERR_ARG, "chunk exceeds the text");
                if (streaming && !longk && bad_stream_chunk(c)) return fail(FMSI_GPU_ERR_ARG, kBadStreamChunk);
            }
            int e = upload_chunks(0, n_chunks, st);
            if (e) return e;
        }
        int e = stage_words(0, n_words, st);
        if (e) return e;
        e = run_span(0, chunks_used, 0, n_results, st, on_host ? s.ls : idx->user);
        if (e) return e;
        if ((e = finish_bits(st))) return e;
        if (on_host) {
            if ((e = copy_back(0, n_results, st))) return e;
            CU(cudaEventRecord(s.done, st));
            CU(cudaStreamSynchronize(st));
        }
        return FMSI_GPU_OK;
    };
    if (!pipelined) return single_batch();

    // Pieces [p0, p1) of the text (32-base aligned, so no packed word is written twice) on the copy stream `st`,
    // each followed by the metadata of the chunks that now lie wholly inside the uploaded prefix; those chunks go
    // to one of two query streams. The results of a span are copied back after the NEXT span has been launched:
    // with pageable host buffers both copies block the calling thread, and this order keeps the GPU busy meanwhile.
    // Chunks are validated as they are scheduled; chunks out of text order end the pipeline and the call starts
    // over as a single batch (same results).
    static const bool trace = std::getenv("FMSI_GPU_TRACE") != nullptr;  // host-side timeline of the pipelined call
    const auto tr0 = std::chrono::steady_clock::now();
    auto tr_us = [&] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tr0).count(); };
    double tr_validate = 0;
    cudaStream_t qs[2] = {idx->slots[1].stream, idx->aux_stream};
    LaunchScratch *qls[2] = {&idx->slots[1].ls, &s.ls};
    size_t c_done = 0, r_done = 0, span = 0;
    size_t back_r0 = 0, back_r1 = 0;  // span whose results are still on the device
    cudaStream_t back_q = nullptr;
    bool ordered = true;
    size_t rd_done = 0;  // reads mode: reads whose chunks have been scheduled
    u64 rd_R = 0, rd_C = 0;  //             results / device chunks of those reads
    if (reads_mode && (rc = expand_reads(st))) return rc;
    for (size_t p0 = 0; p0 < n_bases && ordered; p0 += kPiece) {
        const size_t p1 = std::min(n_bases, p0 + kPiece);
        const bool last = p1 == n_bases;
        const size_t w0 = p0 / 32, w1 = last ? n_words : p1 / 32;
        if ((rc = stage_words(w0, w1, st))) return rc;
        size_t c1 = c_done;
        const double tv0 = trace ? tr_us() : 0;
        if (reads_mode) {  // the reads that lie wholly inside the uploaded prefix
            for (; rd_done < n_reads; ++rd_done) {
                u64 nk, nch;
                if (!read_counts(rd_done, nk, nch)) return fail(FMSI_GPU_ERR_ARG, "read offsets must be non-decreasing and end inside the text");
                if (!last && chunk_off[rd_done + 1] > 32 * w1) break;
                rd_R += nk;
                rd_C += nch;
            }
            if (trace) tr_validate += tr_us() - tv0;
            if (rd_C == c_done && rd_R == r_done) continue;
            if (last && rd_R != n_results) return fail(FMSI_GPU_ERR_ARG, "n_results does not match the reads");
            if (rd_R > n_results) return fail(FMSI_GPU_ERR_ARG, "n_results does not match the reads");
            if (idx->piece_events.size() <= span) {
                cudaEvent_t ev;
                CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                idx->piece_events.push_back(ev);
            }
            CU(cudaEventRecord(idx->piece_events[span], st));
            cudaStream_t q = qs[span & 1];
            CU(cudaStreamWaitEvent(q, idx->piece_events[span], 0));
            if ((rc = run_span(c_done, (size_t)rd_C, r_done, (size_t)rd_R, q, *qls[span & 1]))) return rc;
            if ((rc = span_bits(span, (size_t)rd_R, last && rd_done == n_reads, q))) return rc;
            if (back_q && (rc = copy_back(back_r0, back_r1, back_q))) return rc;
            back_r0 = r_done;
            back_r1 = (size_t)rd_R;
            back_q = q;
            c_done = (size_t)rd_C;
            r_done = (size_t)rd_R;
            ++span;
            continue;
        }
        for (; c1 < n_chunks; ++c1) {
            const uint64_t end = chunk_off[c1] + chunk_len[c1];
            if (c1 > 0 && (chunk_off[c1] < chunk_off[c1 - 1] || end < chunk_off[c1 - 1] + chunk_len[c1 - 1] || res_off[c1] < res_off[c1 - 1])) {
                ordered = false;
                break;
            }
            if (end > n_bases) return fail(FMSI_GPU_ERR_ARG, "chunk exceeds the text");
            if (!last && end > 32 * w1) break;
            if (streaming && !longk && bad_stream_chunk(c1)) return fail(FMSI_GPU_ERR_ARG, kBadStreamChunk);
        }
        if (trace) tr_validate += tr_us() - tv0;
        if (!ordered || c1 == c_done) continue;
        if ((rc = upload_chunks(c_done, c1, st))) return rc;
        const size_t r1 = std::max(r_done, c1 < n_chunks ? (size_t)std::min<uint64_t>(res_off[c1], n_results) : n_results);
        if (idx->piece_events.size() <= span) {
            cudaEvent_t ev;
            CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            idx->piece_events.push_back(ev);
        }
        CU(cudaEventRecord(idx->piece_events[span], st));
        cudaStream_t q = qs[span & 1];
        CU(cudaStreamWaitEvent(q, idx->piece_events[span], 0));
        if ((rc = run_span(c_done, c1, r_done, r1, q, *qls[span & 1]))) return rc;
        if ((rc = span_bits(span, r1, last && c1 == n_chunks, q))) return rc;
        if (back_q && (rc = copy_back(back_r0, back_r1, back_q))) return rc;
        back_r0 = r_done;
        back_r1 = r1;
        back_q = q;
        c_done = c1;
        r_done = r1;
        ++span;
    }
    if (ordered && back_q && (rc = copy_back(back_r0, back_r1, back_q))) return rc;
    const double tr_enq = trace ? tr_us() : 0;
    CU(cudaEventRecord(s.done, st));
    CU(cudaEventRecord(idx->slots[1].done, qs[0]));
    CU(cudaStreamSynchronize(st));
    CU(cudaStreamSynchronize(qs[0]));
    CU(cudaStreamSynchronize(qs[1]));
    if (trace)
        fprintf(stderr, "[fmsi trace] chunks call: %zu chunks, %zu results, %zu spans: all enqueued at %.0f us (chunk validation %.0f us), drained at %.0f us\n",
                n_chunks, n_results, span, tr_enq, tr_validate, tr_us());
    if (!ordered) return single_batch();
    if (bits && bits_done < (n_results + 7) / 8) {  // (result slots no chunk covers at the end) pack and fetch everything
        if ((rc = finish_bits(st))) return rc;
        CU(cudaStreamSynchronize(st));
    }
    return FMSI_GPU_OK;
}

}  // namespace

extern "C" {

int fmsi_gpu_query_kmers(fmsi_gpu_index *idx, int mode, int output, int strands, const uint64_t *kmers, size_t n, int k,
                         void *results, int mem, void *stream) {
    if (mode != FMSI_GPU_MODE_OR && mode != FMSI_GPU_MODE_ALL) return fail(FMSI_GPU_ERR_ARG, "bad mode/output/strands");
    return query_kmers_impl(idx, mode, output, strands, nullptr, kmers, n, k, results, mem, stream);
}

int fmsi_gpu_query_chunks(fmsi_gpu_index *idx, int mode, int output, int strands, int streaming, const char *bases,
                          size_t n_bases, const uint64_t *chunk_off, const uint32_t *chunk_len, const uint64_t *res_off,
                          size_t n_chunks, size_t n_results, int k, void *results, int mem, void *stream) {
    if (mode != FMSI_GPU_MODE_OR && mode != FMSI_GPU_MODE_ALL) return fail(FMSI_GPU_ERR_ARG, "bad mode/output/strands");
    return query_chunks_impl(idx, mode, output, strands, streaming, nullptr, bases, FMSI_GPU_TEXT_ASCII, n_bases, chunk_off, chunk_len, res_off, n_chunks,
                             n_results, k, results, mem, stream);
}

int fmsi_gpu_query_chunks_packed(fmsi_gpu_index *idx, int mode, int output, int strands, int streaming, const uint64_t *text2,
                                 size_t n_bases, const uint64_t *chunk_off, const uint32_t *chunk_len, const uint64_t *res_off,
                                 size_t n_chunks, size_t n_results, int k, void *results, int mem, void *stream) {
    if (mode != FMSI_GPU_MODE_OR && mode != FMSI_GPU_MODE_ALL) return fail(FMSI_GPU_ERR_ARG, "bad mode/output/strands");
    return query_chunks_impl(idx, mode, output, strands, streaming, nullptr, text2, FMSI_GPU_TEXT_PACKED2, n_bases, chunk_off, chunk_len, res_off,
                             n_chunks, n_results, k, results, mem, stream);
}

int fmsi_gpu_query_reads_packed(fmsi_gpu_index *idx, int mode, int output, int strands, int streaming, const uint64_t *text2, size_t n_bases,
                                const uint64_t *read_off, size_t n_reads, size_t n_results, int k, void *results, int mem, void *stream) {
    if (mode != FMSI_GPU_MODE_OR && mode != FMSI_GPU_MODE_ALL) return fail(FMSI_GPU_ERR_ARG, "bad mode/output/strands");
    if (n_reads >= (1ull << 31)) return fail(FMSI_GPU_ERR_ARG, "too many reads in one call");
    return query_chunks_impl(idx, mode, output, strands, streaming, nullptr, text2, FMSI_GPU_TEXT_PACKED2, n_bases, read_off, nullptr, nullptr, n_reads,
                             n_results, k, results, mem, stream);
}

int fmsi_gpu_query_kmers_general(fmsi_gpu_index *idx, const fmsi_gpu_function *f, const uint64_t *kmers, size_t n, int k,
                                 uint8_t *results, int mem, void *stream) {
    GenF gf;
    if (!to_genf(f, gf)) return fail(FMSI_GPU_ERR_ARG, "bad demasking function");
    return query_kmers_impl(idx, FMSI_GPU_MODE_GENERAL_, FMSI_GPU_OUT_PRESENCE, FMSI_GPU_STRANDS_LAZY, &gf, kmers, n, k, results, mem, stream);
}

int fmsi_gpu_query_chunks_general(fmsi_gpu_index *idx, const fmsi_gpu_function *f, const char *bases, size_t n_bases,
                                  const uint64_t *chunk_off, const uint32_t *chunk_len, const uint64_t *res_off, size_t n_chunks,
                                  size_t n_results, int k, uint8_t *results, int mem, void *stream) {
    GenF gf;
    if (!to_genf(f, gf)) return fail(FMSI_GPU_ERR_ARG, "bad demasking function");
    return query_chunks_impl(idx, FMSI_GPU_MODE_GENERAL_, FMSI_GPU_OUT_PRESENCE, FMSI_GPU_STRANDS_LAZY, 0, &gf, bases, FMSI_GPU_TEXT_ASCII, n_bases, chunk_off,
                             chunk_len, res_off, n_chunks, n_results, k, results, mem, stream);
}

// ---------------------------------------------------------------------------------- multi-GPU pool
}  // extern "C"

struct fmsi_gpu_pool {
    std::vector<fmsi_gpu_index *> members;  // members[0] is the caller's primary
    fmsi_gpu_index *primary = nullptr;      // not owned: fmsi_gpu_pool_free releases every other member
};

namespace {

// Device-to-device copy of one array of the primary onto `dev` (NVLink peer copy when the driver
// allows peer access, staged through the host by the driver otherwise).
int replicate_array(void **dst, int dev, const void *src, int src_dev, size_t bytes) {
    *dst = nullptr;
    if (!src || !bytes) return FMSI_GPU_OK;
    CU(cudaSetDevice(dev));
    CU(cudaMalloc(dst, bytes));
    if (dev == src_dev) CU(cudaMemcpy(*dst, src, bytes, cudaMemcpyDeviceToDevice));
    else CU(cudaMemcpyPeer(*dst, dev, src, src_dev, bytes));
    return FMSI_GPU_OK;
}

int replicate_index(const fmsi_gpu_index *src, int dev, fmsi_gpu_index **out) {
    std::unique_ptr<fmsi_gpu_index> r(new fmsi_gpu_index());
    r->device = dev;
    int rc = select_device(r.get());
    if (rc) return rc;
    if (dev != src->device) {  // enable NVLink/PCIe peer access in both directions when available
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, dev, src->device) == cudaSuccess && can) {
            cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(FMSI_GPU_ERR_CUDA, cudaGetErrorString(e));
            cudaGetLastError();
        }
    }
    r->meta = src->meta;
    r->wide = src->wide;
    r->dev = src->dev;
    r->dict = src->dict;
    r->fold = src->fold;
    r->fold_ids_policy = src->fold_ids_policy;
    r->loc = src->loc;
    r->loc_policy = src->loc_policy;
    r->loc_failed = src->loc_failed;
    r->b_ldir = src->b_ldir;
    r->b_lrows = src->b_lrows;
    r->b_fbuckets = src->b_fbuckets;
    r->b_frows = src->b_frows;
    r->b_fids = src->b_fids;
    r->hbm_bytes = src->hbm_bytes;
    r->b_rank = src->b_rank;
    r->b_aux = src->b_aux;
    r->b_table = src->b_table;
    r->b_sb = src->b_sb;
    r->b_rows = src->b_rows;
    r->b_multi = src->b_multi;
    auto bail = [&](int code) {
        fmsi_gpu_index_free(r.release());
        return code;
    };
    if ((rc = replicate_array(&r->d_rank, dev, src->d_rank, src->device, src->b_rank)) ||
        (rc = replicate_array(&r->d_aux, dev, src->d_aux, src->device, src->b_aux)) ||
        (rc = replicate_array(&r->d_table, dev, src->d_table, src->device, src->b_table)) ||
        (rc = replicate_array(&r->d_sb, dev, src->d_sb, src->device, src->b_sb)) ||
        (rc = replicate_array(&r->d_counts, dev, src->d_counts, src->device, 32)) ||
        (rc = replicate_array(&r->d_rows, dev, src->d_rows, src->device, src->b_rows)) ||
        (rc = replicate_array(&r->d_fbuckets, dev, src->d_fbuckets, src->device, src->b_fbuckets)) ||
        (rc = replicate_array(&r->d_frows, dev, src->d_frows, src->device, src->b_frows)) ||
        (rc = replicate_array(&r->d_fids, dev, src->d_fids, src->device, src->b_fids)) ||
        (rc = replicate_array(&r->d_multi, dev, src->d_multi, src->device, src->b_multi)) ||
        (rc = replicate_array(&r->d_ldir, dev, src->d_ldir, src->device, src->b_ldir)) ||
        (rc = replicate_array(&r->d_lrows, dev, src->d_lrows, src->device, src->b_lrows)))
        return bail(rc);
    r->dev.rank = reinterpret_cast<const RankBlock *>(r->d_rank);
    r->dev.aux = reinterpret_cast<const AuxBlock *>(r->d_aux);
    r->dev.table = r->d_table;
    r->dev.multi = reinterpret_cast<const MultiBlock *>(r->d_multi);
    r->dev.sb_base = reinterpret_cast<const u64 *>(r->d_sb);
    r->dict.rows = reinterpret_cast<const u64 *>(r->d_rows);
    r->fold.buckets = r->d_fbuckets;
    r->fold.orows = reinterpret_cast<const u64 *>(r->d_frows);
    r->fold.ids = reinterpret_cast<const uint2 *>(r->d_fids);
    r->loc.dir = reinterpret_cast<const u64 *>(r->d_ldir);
    r->loc.rows = reinterpret_cast<const u64 *>(r->d_lrows);
    if (cudaSetDevice(dev) != cudaSuccess) return bail(fail(FMSI_GPU_ERR_CUDA, "cudaSetDevice"));
    if ((rc = alloc_slots(r.get()))) return bail(rc);
    *out = r.release();
    return FMSI_GPU_OK;
}

// Run fn(member, shard) on one host thread per member; first failure wins.
template <typename Fn>
int pool_run(fmsi_gpu_pool *pool, size_t n_shards, Fn fn) {
    std::vector<int> rcs(n_shards, FMSI_GPU_OK);
    std::vector<std::string> errs(n_shards);
    std::vector<std::thread> th;
    for (size_t s = 0; s < n_shards; ++s)
        th.emplace_back([&, s] {
            rcs[s] = fn(pool->members[s], s);
            if (rcs[s]) errs[s] = fmsi_gpu_last_error();  // thread-local in the worker
        });
    for (auto &t : th) t.join();
    for (size_t s = 0; s < n_shards; ++s)
        if (rcs[s]) return fail(rcs[s], "pool member " + std::to_string(s) + ": " + errs[s]);
    return FMSI_GPU_OK;
}

}  // namespace

extern "C" {

int fmsi_gpu_pool_create(fmsi_gpu_index *primary, const int *devices, int n_devices, fmsi_gpu_pool **out) {
    if (!primary || !devices || n_devices < 1 || !out) return fail(FMSI_GPU_ERR_ARG, "bad pool arguments");
    *out = nullptr;
    std::unique_ptr<fmsi_gpu_pool> pool(new fmsi_gpu_pool());
    pool->primary = primary;  // ownership is by identity, not by position: a failure below may come before the rotation
    bool primary_used = false;
    for (int m = 0; m < n_devices; ++m) {
        if (devices[m] == primary->device && !primary_used) {
            pool->members.push_back(primary);
            primary_used = true;
            continue;
        }
        fmsi_gpu_index *r = nullptr;
        int rc = replicate_index(primary, devices[m], &r);
        if (rc) {
            fmsi_gpu_pool_free(pool.release());
            return rc;
        }
        pool->members.push_back(r);
    }
    if (!primary_used) {  // keep the invariant members[0] == primary by rotating it in front
        pool->members.insert(pool->members.begin(), primary);
    }
    // the primary must sit at position 0 so that pool_free knows which member it does not own
    for (size_t m = 1; m < pool->members.size(); ++m)
        if (pool->members[m] == primary) std::swap(pool->members[0], pool->members[m]);
    CU(cudaSetDevice(primary->device));
    *out = pool.release();
    return FMSI_GPU_OK;
}

int fmsi_gpu_pool_size(const fmsi_gpu_pool *pool) { return pool ? (int)pool->members.size() : 0; }

fmsi_gpu_index *fmsi_gpu_pool_member(fmsi_gpu_pool *pool, int m) {
    if (!pool || m < 0 || (size_t)m >= pool->members.size()) return nullptr;
    return pool->members[(size_t)m];
}

int fmsi_gpu_pool_free(fmsi_gpu_pool *pool) {
    if (!pool) return FMSI_GPU_OK;
    for (fmsi_gpu_index *m : pool->members)
        if (m != pool->primary) fmsi_gpu_index_free(m);
    delete pool;
    return FMSI_GPU_OK;
}

int fmsi_gpu_pool_query_kmers(fmsi_gpu_pool *pool, int mode, int output, int strands, const uint64_t *kmers,
                              size_t n, int k, void *results) {
    if (!pool || pool->members.empty()) return fail(FMSI_GPU_ERR_ARG, "null pool");
    if (n == 0) return FMSI_GPU_OK;
    if (!kmers || !results) return fail(FMSI_GPU_ERR_ARG, "null buffer");
    if ((output != FMSI_GPU_OUT_PRESENCE && output != FMSI_GPU_OUT_ORDERS) ||
        (strands != FMSI_GPU_STRANDS_LAZY && strands != FMSI_GPU_STRANDS_BOTH))
        return fail(FMSI_GPU_ERR_ARG, "bad mode/output/strands");
    const size_t W = pool->members.size(), rbytes = result_bytes(output, strands);
    const size_t base = n / W, extra = n % W;  // contiguous ranges, sizes differing by at most one
    std::vector<size_t> begin(W + 1, 0);
    for (size_t s = 0; s < W; ++s) begin[s + 1] = begin[s] + base + (s < extra ? 1 : 0);
    return pool_run(pool, W, [&](fmsi_gpu_index *m, size_t s) {
        const size_t b = begin[s], cnt = begin[s + 1] - b;
        if (!cnt) return (int)FMSI_GPU_OK;
        return fmsi_gpu_query_kmers(m, mode, output, strands, kmers + b, cnt, k, (char *)results + b * rbytes, FMSI_GPU_MEM_HOST, nullptr);
    });
}

int fmsi_gpu_pool_query_chunks(fmsi_gpu_pool *pool, int mode, int output, int strands, int streaming, const char *bases,
                               size_t n_bases, const uint64_t *chunk_off, const uint32_t *chunk_len, size_t n_chunks,
                               size_t n_results, int k, void *results) {
    if (!pool || pool->members.empty()) return fail(FMSI_GPU_ERR_ARG, "null pool");
    if (n_chunks == 0 || n_results == 0) return FMSI_GPU_OK;
    if (!bases || !chunk_off || !chunk_len || !results) return fail(FMSI_GPU_ERR_ARG, "null buffer");
    if (k < 1 || k > FMSI_GPU_MAX_K) return fail(FMSI_GPU_ERR_K, "k must be in [1, FMSI_GPU_MAX_K]");
    if ((output != FMSI_GPU_OUT_PRESENCE && output != FMSI_GPU_OUT_ORDERS) ||
        (strands != FMSI_GPU_STRANDS_LAZY && strands != FMSI_GPU_STRANDS_BOTH))
        return fail(FMSI_GPU_ERR_ARG, "bad mode/output/strands");
    // results of chunk c start at roff[c] = sum of (chunk_len - k + 1) over earlier chunks
    std::vector<uint64_t> roff(n_chunks + 1, 0);
    for (size_t c = 0; c < n_chunks; ++c) {
        if (chunk_len[c] < (uint32_t)k || chunk_off[c] + chunk_len[c] > n_bases) return fail(FMSI_GPU_ERR_ARG, "chunk out of range or shorter than k");
        roff[c + 1] = roff[c] + (chunk_len[c] - (uint32_t)k + 1);
    }
    if (roff[n_chunks] != n_results) return fail(FMSI_GPU_ERR_ARG, "n_results does not match the chunks");
    const size_t W = pool->members.size(), rbytes = result_bytes(output, strands);
    std::vector<size_t> cb(W + 1, 0);  // chunk ranges balanced by k-mer count; chunks are never split
    for (size_t s = 1; s < W; ++s) {
        const uint64_t target = n_results / W * s;
        cb[s] = (size_t)(std::lower_bound(roff.begin(), roff.end(), target) - roff.begin());
        if (cb[s] > n_chunks) cb[s] = n_chunks;
        if (cb[s] < cb[s - 1]) cb[s] = cb[s - 1];
    }
    cb[W] = n_chunks;
    return pool_run(pool, W, [&](fmsi_gpu_index *m, size_t s) {
        const size_t b = cb[s], e = cb[s + 1];
        if (b == e) return (int)FMSI_GPU_OK;
        uint64_t lo = ~0ull, hi = 0;
        for (size_t c = b; c < e; ++c) {
            lo = std::min<uint64_t>(lo, chunk_off[c]);
            hi = std::max<uint64_t>(hi, chunk_off[c] + chunk_len[c]);
        }
        std::vector<uint64_t> off(e - b), ro(e - b);
        for (size_t c = b; c < e; ++c) {
            off[c - b] = chunk_off[c] - lo;
            ro[c - b] = roff[c] - roff[b];
        }
        return fmsi_gpu_query_chunks(m, mode, output, strands, streaming, bases + lo, (size_t)(hi - lo), off.data(), chunk_len + b,
                                     ro.data(), e - b, (size_t)(roff[e] - roff[b]), k, (char *)results + roff[b] * rbytes,
                                     FMSI_GPU_MEM_HOST, nullptr);
    });
}

}  // extern "C"
