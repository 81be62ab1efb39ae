// fold.cuh — the strand-folded k-mer dictionary: one memory request per k-mer, both strands at once.
//
// Why. On B200 a dependent random read costs one DRAM row activation (~37-44 G/s per GPU, measured
// with tools/randbw2.cu) whatever it returns, so the cost of a query is the number of dependent
// probes. The SA-ordered dictionary of dict.cuh needs one probe per STRAND search — 1.75 per k-mer
// on the bench's mix, 2 for absent k-mers and for STRANDS_BOTH. Everything query_kmers_single
// (src/fms_index.h:263-331) can say about a k-mer q is a function of two SA intervals, the one of q and
// the one of its reverse complement, and only through three facts per interval (infer_presence /
// kmer_order_if_present, :126-156): is it non-empty, is mask[sa_start] set, is any mask bit in it set
// (+ rank1(sa_start) for lookup). This tier stores exactly those facts, once per distinct k-mer
// of the index, keyed by the k-mer in a strand-neutral orientation:
//
//   orient(q)  = q if q*C <= rc(q)*C (mod 2^64) else rc(q)   (C odd: a bijection, so ties <=> q == rc(q);
//                a hash order, unlike the lexicographic minimum, keeps the first t bases uniform)
//   key(q)     = a bijective mix of orient(q) on 2k bits (fold_key) — a hash table, so the load of a bucket does
//                not depend on how the index's k-mers share prefixes
//   bucket[x], x = top 2t bits of key(q)  (32 B = one sector, 4^t of them):
//       u32 start, end      rows [start, end) of this bucket in rows[] / ids[] (sorted by payload, distinct)
//       u32 flags           4 bits per inline row: sf | sr << 2
//       PAY32: u32 pay[5]       the low 2B bits (B = k - t <= 16) of the keys of the first 5 rows
//       PAY64: u32 pad; u64 pay[2]   (16 < B <= 30) first 2 rows
//   rows[g]  (8 B)  pay << 4 | sf | sr << 2   — read only when the bucket holds more rows than fit inline
//   ids[g]   (8 B)  {id_f, id_r} = rank1(sa_start) of either orientation (lookup only)
//   sf / sr: state of the interval of orient(q) / of its reverse complement:
//       0 empty, 1 non-empty without ON occurrence, 2 ON occurrence(s) but mask[sa_start] == 0, 3 mask[sa_start] == 1
//
// A query computes orient(q), reads one bucket, compares <= 5 payloads; a row match is unique, so there is
// no scan over runs and no overflow list: a bucket with more rows than fit is binary-searched in rows[]
// (one sector = 4 rows per step). or / -O / lookup and LAZY / BOTH strands are all decided from (sf, sr).
//
// Built on the device from the BWT alone (file-loaded and device-built indexes alike): psi = inverse
// LF-mapping, the k-mer of every SA row by walking psi (SA order = sorted k-mers, so equal k-mers are
// runs = their SA intervals), one entry per run (state from the mask ranks at the run's ends), entries
// radix-sorted by orient(k-mer), the two orientations of a k-mer merged into one row.
#pragma once
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "dict.cuh"
#include "index_build.cuh"

namespace fmsi {

constexpr u32 kFoldCap32 = 5, kFoldCap64 = 2;
constexpr u64 kFoldMul = 0x9E3779B97F4A7C15ull;
constexpr u32 kFoldNoId = 0xFFFFFFFFu;

struct FoldView {
    const void *buckets;  // [4^t] 32-byte sectors
    const u64 *rows;      // [n_rows + 8]
    const uint2 *ids;     // [n_rows]
    u64 n_rows;
    u32 t, B, k, enabled;
};

__device__ __forceinline__ bool fold_swapped(u64 q, u64 rc) { return rc * kFoldMul < q * kFoldMul; }

// The table key of an oriented k-mer: a bijection of [0, 4^k) (multiply by an odd constant, fold the upper half
// into the lower, multiply again, all mod 4^k), so rows stay unique while the bucket index (the top 2t bits) no
// longer follows the k-mer's first t bases. Indexes whose k-mers share long prefixes (pangenomes: thousands of
// near-copies of one genome) would otherwise pile their rows into the few buckets of the genome's own t-mers.
constexpr u64 kFoldMul2 = 0xD6E8FEB86659FD93ull;
__device__ __forceinline__ u64 fold_key(u64 c, u32 k) {
    const u64 m = k < 32 ? (1ull << (2 * k)) - 1ull : ~0ull;
    c = (c * kFoldMul) & m;
    c ^= c >> k;
    c = (c * kFoldMul2) & m;
    c ^= c >> k;
    return c;
}

// ------------------------------------------------------------------------------------------- build
// kmers[r] = the first k characters of the suffix of SA row r (padded with A past the sentinel);
// valid bit r = that suffix has >= k characters. F column from counts[].
__global__ void fold_kmers_kernel(const u64 n, const u32 *__restrict__ psi, const u32 c1, const u32 c2, const u32 c3, const u32 k,
                                  u64 *__restrict__ kmers, u32 *__restrict__ validbits) {
    const u64 r = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    bool valid = false;
    if (r < n) {
        u32 cur = (u32)r;
        u64 km = 0;
        bool ended = false;
        for (u32 s = 0; s < k; ++s) {
            if (cur == 0) ended = true;  // row 0 is the sentinel's suffix
            const u32 c = ended ? 0u : (cur >= c3) ? 3u : (cur >= c2) ? 2u : (cur >= c1) ? 1u : 0u;
            km = (km << 2) | c;
            if (!ended && s + 1 < k) cur = __ldg(psi + cur);
        }
        kmers[r] = km;
        valid = !ended;
    }
    const unsigned b = __ballot_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31u) == 0 && r < n) validbits[r >> 5] = b;
}

struct FoldRunHead {  // row r starts a run of equal k-mers of equal validity
    const u64 *kmers;
    const u32 *validbits;
    __host__ __device__ __forceinline__ bool operator()(const u32 r) const {
        if (r == 0) return true;
        const bool v = (validbits[r >> 5] >> (r & 31u)) & 1u, pv = (validbits[(r - 1) >> 5] >> ((r - 1) & 31u)) & 1u;
        return v != pv || kmers[r] != kmers[r - 1];
    }
};
struct FoldKeyHead {  // entry m starts a group of equal keys
    const u64 *keys;
    __host__ __device__ __forceinline__ bool operator()(const u32 m) const { return m == 0 || keys[m] != keys[m - 1]; }
};

__device__ __forceinline__ u64 fold_rank1(const DevIndex &d, u64 p) {  // ones in mask[0, p), p in [0, N]
    const AuxBlock &a = d.aux[p >> 6];
    return a.mask_cum + (u64)__popcll(a.mask & low_mask((u32)p & 63u));
}

// One entry per run h = SA interval [heads[h], heads[h+1]) of a distinct k-mer: key = orient(k-mer),
// val = id | state << 32 | (the run is the reverse complement of its key) << 34. Runs of invalid rows
// get state 0 and drop out when the rows are assembled.
__global__ void fold_entries_kernel(const DevIndex d, const u64 *__restrict__ kmers, const u32 *__restrict__ validbits,
                                    const u32 *__restrict__ heads, const u64 n_heads, const u32 k, u64 *__restrict__ keys,
                                    u64 *__restrict__ vals) {
    const u64 h = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (h >= n_heads) return;
    const u64 i = heads[h], j = (h + 1 < n_heads) ? (u64)heads[h + 1] : d.n;
    const u64 q = kmers[i];
    const bool valid = (validbits[i >> 5] >> (i & 31u)) & 1u;
    const u64 rc = revcomp_packed(q, k);
    const bool sw = fold_swapped(q, rc);
    u64 state = 0, id = 0;
    if (valid) {
        const u64 ri = fold_rank1(d, i), rj = fold_rank1(d, j);
        const bool first = (d.aux[i >> 6].mask >> (i & 63)) & 1ull;
        state = rj > ri ? (first ? 3 : 2) : 1;
        id = ri;
    }
    keys[h] = fold_key(sw ? rc : q, k);
    vals[h] = id | (state << 32) | ((u64)sw << 34) | ((u64)(q == rc) << 35);
}

struct alignas(32) FoldBucket {
    u32 start, end, flags, w[5];
};
static_assert(sizeof(FoldBucket) == 32, "one sector");

// One row per group g = entries [gs[g], gs[g+1]) of equal key (the run of the key itself and/or the run
// of its reverse complement; a self-complementary k-mer has one run that stands for both).
__global__ void fold_rows_kernel(const u64 *__restrict__ keys, const u64 *__restrict__ vals, const u32 *__restrict__ gs,
                                 const u64 n_groups, const u64 n_entries, const u32 k, const u32 B, u64 *__restrict__ rows,
                                 uint2 *__restrict__ ids, FoldBucket *__restrict__ buckets) {
    const u64 g = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const u64 e0 = gs[g], e1 = (g + 1 < n_groups) ? (u64)gs[g + 1] : n_entries;
    const u64 key = keys[e0];
    u32 sf = 0, sr = 0, idf = kFoldNoId, idr = kFoldNoId;
    bool self_rc = false;
    for (u64 e = e0; e < e1; ++e) {
        const u64 v = vals[e];
        const u32 st = (u32)(v >> 32) & 3u;
        self_rc |= (v >> 35) & 1ull;
        if (!st) continue;
        if ((v >> 34) & 1ull) {
            sr = st;
            idr = (u32)v;
        } else {
            sf = st;
            idf = (u32)v;
        }
    }
    if (self_rc) {
        sr = sf;
        idr = idf;
    }
    if (sf < 2) idf = kFoldNoId;
    if (sr < 2) idr = kFoldNoId;
    const u64 pmask = B ? ((1ull << (2 * B)) - 1ull) : 0ull;
    rows[g] = ((key & pmask) << 4) | sf | (sr << 2);
    ids[g] = make_uint2(idf, idr);
    const u64 x = key >> (2 * B);
    const bool first = g == 0 || (keys[gs[g - 1]] >> (2 * B)) != x;
    const bool last = g + 1 == n_groups || (keys[e1] >> (2 * B)) != x;
    if (first) buckets[x].start = (u32)g;
    if (last) buckets[x].end = (u32)(g + 1);
}

template <bool PAY64>
__global__ void fold_bucket_fill_kernel(const u64 *__restrict__ rows, const u64 total, FoldBucket *__restrict__ buckets) {
    const u64 x = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (x >= total) return;
    FoldBucket b = buckets[x];
    b.flags = 0;
    for (int s = 0; s < 5; ++s) b.w[s] = 0;
    const u32 cap = PAY64 ? kFoldCap64 : kFoldCap32;
    const u32 cnt = b.end - b.start;
    const u32 m = cnt < cap ? cnt : cap;
    for (u32 s = 0; s < m; ++s) {
        const u64 row = rows[(u64)b.start + s];
        b.flags |= (u32)(row & 15ull) << (4 * s);
        const u64 pay = row >> 4;
        if (PAY64) {
            b.w[1 + 2 * s] = (u32)pay;
            b.w[2 + 2 * s] = (u32)(pay >> 32);
        } else {
            b.w[s] = (u32)pay;
        }
    }
    buckets[x] = b;
}

// out[] = the indices r in [0, count) with pred(r), in order; returns how many. Done in slices of 2^30
// indices (one cub::DeviceSelect::If each) so that every slice stays within 32-bit offsets.
template <typename Pred>
inline u64 fold_select_heads(u64 count, u32 *out, Pred pred) {
    DevArr<u64> d_num(1);
    const u64 slice = 1ull << 30;
    size_t tmp_bytes = 0;
    BCU(cub::DeviceSelect::If(nullptr, tmp_bytes, thrust::counting_iterator<u32>(0), out, d_num.p, (int)(count < slice ? count : slice), pred));
    DevArr<unsigned char> tmp(tmp_bytes);
    u64 selected = 0;
    for (u64 base = 0; base < count; base += slice) {
        const u64 m = count - base < slice ? count - base : slice;
        size_t bytes = tmp_bytes;
        BCU(cub::DeviceSelect::If(tmp.p, bytes, thrust::counting_iterator<u32>((u32)base), out + selected, d_num.p, (int)m, pred));
        u64 num = 0;
        BCU(cudaMemcpy(&num, d_num.p, 8, cudaMemcpyDeviceToHost));
        selected += num;
    }
    return selected;
}

struct FoldArrays {  // device arrays of a built tier (ownership passes to the caller)
    FoldBucket *buckets = nullptr;
    u64 *rows = nullptr;
    uint2 *ids = nullptr;
    u64 n_rows = 0;
};

// Peak device memory of build_fold_on_device beyond the index itself (bytes), and what stays resident.
inline u64 fold_build_peak_bytes(u64 N, u32 t) { return 36ull * N + (32ull << (2 * t)) + (64ull << 20); }
inline u64 fold_resident_bytes(u64 N, u32 t) { return 16ull * N + (32ull << (2 * t)); }

// Throws std::runtime_error (out of memory included); nothing is leaked then.
inline void build_fold_on_device(const DevIndex &d, const u64 counts[4], u32 k, u32 t, FoldArrays &out, uint64_t *launches) {
    const u64 N = d.n;
    const u32 B = k - t;
    const u64 total = 1ull << (2 * t);
    auto stage = [](const char *what) {  // surfaces asynchronous errors with the step that caused them
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
    };
    DevArr<FoldBucket> buckets(total);
    BCU(cudaMemset(buckets.p, 0, total * sizeof(FoldBucket)));
    DevArr<u64> keys, vals;
    u64 M = 0;
    {
        DevArr<u64> kmers(N);
        DevArr<u32> validbits((N >> 5) + 2);
        {
            DevArr<u32> psi(N);
            psi_scatter_kernel<<<nblocks_for(N), 256>>>(d, psi.p);
            stage("psi");
            fold_kmers_kernel<<<nblocks_for(N), 256>>>(N, psi.p, (u32)counts[1], (u32)counts[2], (u32)counts[3], k, kmers.p, validbits.p);
            stage("k-mers of the SA rows");
        }
        DevArr<u32> heads(N);
        M = fold_select_heads(N, heads.p, FoldRunHead{kmers.p, validbits.p});
        stage("run heads");
        keys.alloc(M);
        vals.alloc(M);
        fold_entries_kernel<<<nblocks_for(M), 256>>>(d, kmers.p, validbits.p, heads.p, M, k, keys.p, vals.p);
        stage("entries");
    }
    {
        DevArr<u64> keys_alt(M), vals_alt(M);
        radix_sort_pairs(keys, keys_alt, vals, vals_alt, M, (int)(2 * k));
        stage("sort");
    }
    DevArr<u32> gs(M);
    const u64 G = fold_select_heads(M, gs.p, FoldKeyHead{keys.p});
    stage("group heads");
    DevArr<u64> rows(G + 8);
    DevArr<uint2> ids(G + 1);
    BCU(cudaMemset(rows.p + G, 0xff, 8 * sizeof(u64)));
    fold_rows_kernel<<<nblocks_for(G), 256>>>(keys.p, vals.p, gs.p, G, M, k, B, rows.p, ids.p, buckets.p);
    stage("rows");
    if (B > 16) fold_bucket_fill_kernel<true><<<nblocks_for(total), 256>>>(rows.p, total, buckets.p);
    else fold_bucket_fill_kernel<false><<<nblocks_for(total), 256>>>(rows.p, total, buckets.p);
    stage("buckets");
    if (launches) *launches += 12;
    out.buckets = buckets.p;
    out.rows = rows.p;
    out.ids = ids.p;
    out.n_rows = G;
    buckets.p = nullptr;
    rows.p = nullptr;
    ids.p = nullptr;
}

// ------------------------------------------------------------------------------------------- query
enum { FP_BUCKET = 0, FP_SEARCH = 1, FP_IDS = 2 };

// A bucket / rows sector. LD64: ask L2 for a 64-byte fill instead of the whole 128-byte line.
template <bool LD64>
__device__ __forceinline__ void ld_fold_sector(const void *p, u64 &a, u64 &b, u64 &c, u64 &d) {
    if (LD64) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    } else {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    }
}

// value of one orientation's interval: infer_presence<maximized_ones> (fms_index.h:126-144)
template <int MODE>
__device__ __forceinline__ int fold_presence(u32 st) {
    if (st == 0) return -1;
    if (MODE == K_MODE_ALL) return st == 3 ? 1 : 0;
    return st >= 2 ? 1 : 0;
}

template <int MODE, int OUT, int STRANDS, bool PAY64, bool LD64>
__global__ void __launch_bounds__(kQueryBlock)
fold_query_kernel(const FoldView fv, const u64 *__restrict__ kmers, const u64 n, void *__restrict__ out,
                  unsigned long long *__restrict__ cursor, const u32 chunk, unsigned long long *__restrict__ probe_ctr) {
    const unsigned FULL = 0xffffffffu;
    const u32 lane = threadIdx.x & 31u;
    const u32 lt_mask = (1u << lane) - 1u;
    const u32 k = fv.k, B = fv.B;
    const u64 pmask = B ? ((1ull << (2 * B)) - 1ull) : 0ull;
    const u32 CAP = PAY64 ? kFoldCap64 : kFoldCap32;

    bool active = false, swapped = false;
    u32 phase = FP_BUCKET;
    u64 idx = 0, q = 0, bx = 0;
    u32 lo = 0, hi = 0, row = 0, st = 0;
    u64 cend = 0, wnext = 0, tile_base = 0, bufA = 0, bufB = 0;
    bool exhausted = false;
    u32 nprobe = 0;  // dependent memory requests issued by this lane (reported when probe_ctr is given)

    for (;;) {
        // ---------------------------------------------------------------- refill idle lanes
        const unsigned need = __ballot_sync(FULL, !active);
        if (need && !exhausted) {
            if (wnext >= cend) {
                unsigned long long c0 = 0;
                if (lane == 0) c0 = atomicAdd(cursor, (unsigned long long)chunk);
                c0 = __shfl_sync(FULL, c0, 0);
                if (c0 >= n) {
                    exhausted = true;
                } else {
                    wnext = tile_base = c0;
                    cend = (c0 + chunk < n) ? c0 + chunk : n;
                    bufA = (tile_base + lane < cend) ? kmers[tile_base + lane] : 0ull;
                    bufB = (tile_base + 32 + lane < cend) ? kmers[tile_base + 32 + lane] : 0ull;
                }
            }
            if (!exhausted) {
                const u32 pre = __popc(need & lt_mask);
                const u64 my = wnext + pre;
                const bool take = !active && my < cend;
                const u32 src = (u32)(my - tile_base);
                u64 km = __shfl_sync(FULL, bufA, src & 31u);
                if (__any_sync(FULL, take && src >= 32u)) {
                    const u64 kb = __shfl_sync(FULL, bufB, src & 31u);
                    if (src >= 32u) km = kb;
                }
                const u64 left = cend - wnext;
                const u32 want = __popc(need);
                wnext += (want < left) ? want : left;
                if (wnext - tile_base >= 32) {
                    tile_base += 32;
                    bufA = bufB;
                    bufB = (tile_base + 32 + lane < cend) ? kmers[tile_base + 32 + lane] : 0ull;
                }
                if (take) {
                    active = true;
                    idx = my;
                    if (k < 32) km &= (1ull << (2 * k)) - 1ull;
                    const u64 rc = revcomp_packed(km, k);
                    swapped = fold_swapped(km, rc);
                    const u64 c = fold_key(swapped ? rc : km, k);
                    q = c & pmask;
                    bx = c >> (2 * B);
                    phase = FP_BUCKET;
                }
            }
        }
        if (!__any_sync(FULL, active)) {
            if (exhausted) break;
            continue;
        }

        // ---------------------------------------------------------------- issue this round's loads
        const bool isB = active && phase == FP_BUCKET;
        const bool isS = active && phase == FP_SEARCH;
        const bool isI = active && phase == FP_IDS;
        const u32 r0 = (lo + ((hi - lo) >> 1)) & ~3u;  // rows sector probed by a search step
        u64 a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        if (isB) ld_fold_sector<LD64>(reinterpret_cast<const char *>(fv.buckets) + (bx << 5), a0, a1, a2, a3);
        if (isS) ld_fold_sector<LD64>(fv.rows + r0, a0, a1, a2, a3);
        if (isI) a0 = __ldg(reinterpret_cast<const u64 *>(fv.ids) + row);
        nprobe += (u32)(isB || isS || isI);

        // ---------------------------------------------------------------- consume
        bool done = false;  // (st, row) final
        if (isB) {
            const u32 start = (u32)a0, end = (u32)(a0 >> 32);
            const u32 cnt = end - start;
            const u32 flags = (u32)a1;
            const u32 m = cnt < CAP ? cnt : CAP;
            st = 0;
            bool found = false;
            u64 last = 0;
#pragma unroll
            for (u32 s = 0; s < CAP; ++s) {
                u64 pay;
                if (PAY64) pay = s == 0 ? a2 : a3;
                else pay = s == 0 ? (a1 >> 32) : s == 1 ? (a2 & 0xffffffffull) : s == 2 ? (a2 >> 32) : s == 3 ? (a3 & 0xffffffffull) : (a3 >> 32);
                if (s < m) {
                    last = pay;
                    if (pay == q) {
                        found = true;
                        st = (flags >> (4 * s)) & 15u;
                        row = start + s;
                    }
                }
            }
            if (found || cnt <= CAP || last > q) {
                done = true;
            } else {
                lo = start + CAP;
                hi = end;
                phase = FP_SEARCH;
            }
        } else if (isS) {
            // invariant: rows before lo have payload < q, rows from hi on have payload > q
            const u32 w0 = r0 > lo ? r0 : lo, w1 = (r0 + 4 < hi) ? r0 + 4 : hi;  // rows [w0, w1) of this sector count
            bool found = false;
            u64 first = 0, last = 0;
#pragma unroll
            for (u32 s = 0; s < 4; ++s) {
                const u64 rw = s == 0 ? a0 : s == 1 ? a1 : s == 2 ? a2 : a3;
                const u32 r = r0 + s;
                if (r >= w0 && r < w1) {
                    const u64 pay = rw >> 4;
                    if (r == w0) first = pay;
                    last = pay;
                    if (pay == q) {
                        found = true;
                        st = (u32)rw & 15u;
                        row = r;
                    }
                }
            }
            if (found) done = true;
            else if (last < q) lo = w1;
            else if (first > q) hi = w0;
            else done = true;  // q falls between two rows of this sector: absent (st stays 0)
            if (!done && lo >= hi) done = true;
        }

        if (done || isI) {
            u32 sf = st & 3u, sr = st >> 2;
            if (swapped) {
                const u32 x = sf;
                sf = sr;
                sr = x;
            }
            if (OUT == K_OUT_PRESENCE) {
                const int vf = fold_presence<MODE>(sf), vr = fold_presence<MODE>(sr);
                unsigned char v;
                if (STRANDS == K_STRANDS_BOTH) v = (unsigned char)((vf + 1) | ((vr + 1) << 2));
                else if (MODE == K_MODE_ALL) v = (unsigned char)((vf != -1 ? vf : vr) == 1);  // fms_index.h:294-298
                else v = (unsigned char)(vf == 1 || vr == 1);                                  // :289-293
                reinterpret_cast<unsigned char *>(out)[idx] = v;
                active = false;
            } else {
                const bool hf = sf >= 2, hr = sr >= 2;  // kmer_order_if_present >= 0 (fms_index.h:146-156)
                if (!isI && (hf || hr)) {
                    phase = FP_IDS;
                } else {
                    u32 idf = (u32)a0, idr = (u32)(a0 >> 32);
                    if (swapped) {
                        const u32 x = idf;
                        idf = idr;
                        idr = x;
                    }
                    const long long rf = hf ? (long long)idf : -1ll, rr = hr ? (long long)idr : -1ll;
                    if (STRANDS == K_STRANDS_BOTH) reinterpret_cast<longlong2 *>(out)[idx] = make_longlong2(rf, rr);
                    else reinterpret_cast<long long *>(out)[idx] = hf ? rf : rr;  // :283-288
                    active = false;
                }
            }
        }
    }
    count_probes(probe_ctr, nprobe);
}

}  // namespace fmsi
