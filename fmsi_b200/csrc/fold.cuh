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
//   row        = one distinct key: payload (the low 2B key bits, B = k - t) and state sf | sr << 2; the rows of a
//                bucket are sorted by payload and numbered globally in key order (row g <-> ids[g])
//   bucket[x], x = top 2t bits of key(q)  (32 B = one sector, 4^t of them):
//       u32 gstart          global number of the bucket's first row
//       u32 ovf             where its overflow rows start in orows[] (a multiple of 4; only if it has more rows than fit)
//       u32 meta            bits 0-19 (PAY64: 0-7): 4 state bits per inline row; bits 20-22: inline rows; bit 23: overflow
//       PAY32: u32 pay[5]       payloads of its first 5 rows (B <= 16)
//       PAY64: u32 pad; u64 pay[2]   (16 < B <= 30) first 2 rows
//   orows[]  (8 B entries, one 32-byte sector = 4 entries; only for buckets with more rows than fit inline — 7 % of
//            them at 2.9 rows per bucket): [count of overflow rows][row][row]... padded with ~0 to a multiple of 4, each
//            row = pay << 4 | sf | sr << 2. The first sector settles buckets with up to 3 overflow rows; larger ones
//            are binary-searched one sector per step.
//   ids[g]   (8 B)  {id_f, id_r} = rank1(sa_start) of either orientation — lookup only, and only built when asked for
//            (fmsi_gpu_options.fold_ids; on the first lookup otherwise)
//   sf / sr: state of the interval of orient(q) / of its reverse complement:
//       0 empty, 1 non-empty without ON occurrence, 2 ON occurrence(s) but mask[sa_start] == 0, 3 mask[sa_start] == 1
//
// A query computes orient(q), reads one bucket, compares <= 5 payloads; a row match is unique, so there is
// no scan over runs and no overflow list. or / -O / lookup and LAZY / BOTH strands are all decided from (sf, sr).
// Resident at human scale (N = 3.1e9, t = 15): buckets 34.4 GB + orows 2.6 GB (+ ids 24.8 GB for lookup) — the
// round-1 layout kept every row twice (rows[] 24.8 GB + ids[] 24.8 GB whatever the workload).
//
// Built on the device from the BWT alone (file-loaded and device-built indexes alike): psi = inverse
// LF-mapping, the k-mer of every SA row by walking psi (SA order = sorted k-mers, so equal k-mers are
// runs = their SA intervals), one entry per run (state from the mask ranks at the run's ends), entries
// radix-sorted by key, the two orientations of a k-mer merged into one row. The sort runs in P passes over
// ranges of buckets, so that its buffers (48 bytes per entry in flight) stay a fraction of the index.
#pragma once
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "dict.cuh"
#include "index_build.cuh"

namespace fmsi {

constexpr u32 kFoldCap32 = 5, kFoldCap64 = 2;
constexpr u64 kFoldMul = 0x9E3779B97F4A7C15ull;
constexpr u32 kFoldNoId = 0xFFFFFFFFu;
constexpr u32 kFoldMetaOvf = 1u << 23;

struct FoldView {
    const void *buckets;  // [4^t] 32-byte sectors
    const u64 *orows;     // overflow regions
    const uint2 *ids;     // [n_rows] or null (not built)
    u64 n_rows;
    u32 t, B, k, enabled;
};

__device__ __forceinline__ bool fold_swapped(u64 q, u64 rc) { return rc * kFoldMul < q * kFoldMul; }

// The table key of an oriented k-mer: a bijection of [0, 4^k) (multiply by an odd constant, fold the upper half
// into the lower, multiply again, all mod 4^k), so rows stay unique while the bucket index (the top 2t bits) no
// longer follows the k-mer's first t bases. Indexes whose k-mers share long prefixes (pangenomes: thousands of
// near-copies of one genome) would otherwise pile their rows into the few buckets of the genome's own t-mers.
constexpr u64 kFoldMul2 = 0xD6E8FEB86659FD93ull;
__host__ __device__ __forceinline__ u64 fold_key(u64 c, u32 k) {
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

// The same k-mers by pointer doubling instead of a k-1 step walk per row (30 dependent random reads at k = 31):
//   K_l[r] = the first l characters of the suffix of row r, P_l[r] = the row l characters further on;
//   K_2l[r] = K_l[r] . K_l[P_l[r]],  P_2l[r] = P_l[P_l[r]]            (l = 1, 2, 4, 8, 16: 8 random gathers in all)
// with psi made sticky at row 0 (the sentinel's suffix: psi'[0] = 0, and row 0 reads as A), which pads past the
// sentinel with A exactly like the walk. The rows whose suffix is shorter than k are the k rows LF reaches from row 0.
__device__ __forceinline__ u32 fold_fcol(u32 row, u32 c1, u32 c2, u32 c3) { return (row >= c3) ? 3u : (row >= c2) ? 2u : (row >= c1) ? 1u : 0u; }
__global__ void fold_sticky_kernel(u32 *__restrict__ psi) {
    if (blockIdx.x == 0 && threadIdx.x == 0) psi[0] = 0;
}
// one doubling step: kin == nullptr means l = 1 (characters come from the F column)
__global__ void fold_double_kernel(const u64 n, const u32 *__restrict__ pin, const u32 *__restrict__ kin, const u32 l, const u32 c1, const u32 c2,
                                   const u32 c3, u32 *__restrict__ pout, u32 *__restrict__ kout) {
    const u64 r = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (r >= n) return;
    const u32 p = __ldg(pin + r);
    const u32 a = kin ? __ldg(kin + r) : fold_fcol((u32)r, c1, c2, c3);
    const u32 b = kin ? __ldg(kin + p) : fold_fcol(p, c1, c2, c3);
    kout[r] = (a << (2 * l)) | b;
    pout[r] = __ldg(pin + p);
}
__global__ void fold_double_last_kernel(const u64 n, const u32 *__restrict__ p16, const u32 *__restrict__ k16, const u32 k, u64 *__restrict__ kmers) {
    const u64 r = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (r >= n) return;
    const u64 k32 = ((u64)__ldg(k16 + r) << 32) | (u64)__ldg(k16 + __ldg(p16 + r));
    kmers[r] = k32 >> (2 * (32 - k));
}
// validbits: all rows valid, except the k rows whose suffix holds fewer than k characters (LF-walk from row 0)
__global__ void fold_valid_fill_kernel(const u64 n, u32 *__restrict__ validbits) {
    const u64 w = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (w * 32 >= n) return;
    const u64 left = n - w * 32;
    validbits[w] = left >= 32 ? 0xFFFFFFFFu : ((1u << left) - 1u);
}
__global__ void fold_valid_clear_kernel(const DevIndex d, const u32 k, u32 *__restrict__ validbits) {
    if (blockIdx.x || threadIdx.x) return;
    u64 cur = 0;
    for (u32 s = 0; s < k; ++s) {
        validbits[cur >> 5] &= ~(1u << (cur & 31u));
        if (cur == d.dollar) break;  // that was the whole text
        u64 a0, a1, a2, a3;
        ld_sector_l1(d.rank + (cur >> 6), a0, a1, a2, a3);
        const u32 c = block_symbol(a2, a3, (u32)cur & 63u);
        cur = lf_map<false>(d, a0, a1, a2, a3, (u32)cur, c);
    }
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

__host__ __device__ __forceinline__ u64 fold_revcomp(u64 x, u32 k) {  // host-callable twin of revcomp_packed
    x = ~x;
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
    x = (x >> 32) | (x << 32);
    return x >> (64 - 2 * k);
}

struct FoldRunInPass {  // run h (a valid one) has its key in buckets [x_lo, x_hi)
    const u64 *kmers;
    const u32 *validbits;
    const u32 *heads;
    u32 k, B;
    u64 x_lo, x_hi;
    __host__ __device__ __forceinline__ bool operator()(const u32 h) const {
        const u32 i = heads[h];
        if (!((validbits[i >> 5] >> (i & 31u)) & 1u)) return false;
        const u64 q = kmers[i], rc = fold_revcomp(q, k);
        const u64 x = fold_key((rc * kFoldMul < q * kFoldMul) ? rc : q, k) >> (2 * B);
        return x >= x_lo && x < x_hi;
    }
};

// Sort passes over bucket ranges: bucket x belongs to pass x * P / 4^t.
__host__ __device__ __forceinline__ u64 fold_pass_begin(u64 p, u64 P, u64 total) { return (p * total + P - 1) / P; }
__global__ void fold_pass_hist_kernel(const FoldRunInPass pred, const u64 n_runs, const u64 total, const u32 P, unsigned long long *__restrict__ hist) {
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const u64 h = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (h < n_runs) {
        const u32 i = pred.heads[h];
        if ((pred.validbits[i >> 5] >> (i & 31u)) & 1u) {
            const u64 q = pred.kmers[i], rc = fold_revcomp(q, pred.k);
            const u64 x = fold_key((rc * kFoldMul < q * kFoldMul) ? rc : q, pred.k) >> (2 * pred.B);
            atomicAdd(&sh[(u32)(x * P / total)], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < P && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

__device__ __forceinline__ u64 fold_rank1(const DevIndex &d, u64 p) {  // ones in mask[0, p), p in [0, N]
    const AuxBlock &a = d.aux[p >> 6];
    return a.mask_cum + (u64)__popcll(a.mask & low_mask((u32)p & 63u));
}

// One entry per selected run h = sel[e] = SA interval [heads[h], heads[h+1]) of a distinct k-mer: key = key(k-mer),
// val = id | state << 32 | (the run is the reverse complement of its key's orientation) << 34 | self-complementary << 35.
__global__ void fold_entries_kernel(const DevIndex d, const u64 *__restrict__ kmers, const u32 *__restrict__ heads, const u64 n_heads,
                                    const u32 *__restrict__ sel, const u64 n_sel, const u32 k, u64 *__restrict__ keys, u64 *__restrict__ vals) {
    const u64 e = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (e >= n_sel) return;
    const u64 h = sel[e];
    const u64 i = heads[h], j = (h + 1 < n_heads) ? (u64)heads[h + 1] : d.n;
    const u64 q = kmers[i];
    const u64 rc = revcomp_packed(q, k);
    const bool sw = fold_swapped(q, rc);
    const u64 ri = fold_rank1(d, i), rj = fold_rank1(d, j);
    const bool first = (d.aux[i >> 6].mask >> (i & 63)) & 1ull;
    const u64 state = rj > ri ? (first ? 3 : 2) : 1;
    keys[e] = fold_key(sw ? rc : q, k);
    vals[e] = ri | (state << 32) | ((u64)sw << 34) | ((u64)(q == rc) << 35);
}

struct alignas(32) FoldBucket {
    u32 gstart, ovf, meta, w[5];
};
static_assert(sizeof(FoldBucket) == 32, "one sector");

// One row per group g = entries [gs[g], gs[g+1]) of equal key (the run of the key itself and/or the run of its
// reverse complement; a self-complementary k-mer has one run that stands for both). Row g of this pass is global
// row g0 + g. bfirst / bcount: first local row and number of rows of every bucket of the pass (zeroed before).
__global__ void fold_rows_kernel(const u64 *__restrict__ keys, const u64 *__restrict__ vals, const u32 *__restrict__ gs,
                                 const u64 n_groups, const u64 n_entries, const u32 B, const u64 g0, const u64 x_lo,
                                 u64 *__restrict__ rows, uint2 *__restrict__ ids, u32 *__restrict__ bfirst, u32 *__restrict__ bcount) {
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
    if (rows) rows[g] = ((key & pmask) << 4) | sf | (sr << 2);
    if (ids) ids[g0 + g] = make_uint2(idf, idr);
    if (bfirst) {
        const u64 x = key >> (2 * B);
        const bool first = g == 0 || (keys[gs[g - 1]] >> (2 * B)) != x;
        const bool last = g + 1 == n_groups || (keys[e1] >> (2 * B)) != x;
        if (first) bfirst[x - x_lo] = (u32)g;
        if (last) bcount[x - x_lo] = (u32)(g + 1);  // = first + count once `first` is subtracted below
    }
}

// entries a bucket needs in orows[]: count word + overflow rows, padded to whole sectors (0 when everything is inline)
__global__ void fold_ovf_sizes_kernel(const u32 *__restrict__ bfirst, const u32 *__restrict__ bcount, const u64 n_buckets, const u32 cap,
                                      u32 *__restrict__ sizes) {
    const u64 x = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (x >= n_buckets) return;
    const u32 cnt = bcount[x] ? bcount[x] - bfirst[x] : 0u;
    sizes[x] = cnt > cap ? ((1u + cnt - cap + 3u) & ~3u) : 0u;
}

template <bool PAY64>
__global__ void fold_bucket_fill_kernel(const u64 *__restrict__ rows, const u32 *__restrict__ bfirst, const u32 *__restrict__ bcount,
                                        const u64 *__restrict__ ovf_off, const u64 n_buckets, const u64 g0, const u64 o0,
                                        FoldBucket *__restrict__ buckets, u64 *__restrict__ orows) {
    const u64 x = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (x >= n_buckets) return;
    const u32 cap = PAY64 ? kFoldCap64 : kFoldCap32;
    const u32 first = bfirst[x];
    const u32 cnt = bcount[x] ? bcount[x] - first : 0u;
    FoldBucket b;
    b.gstart = (u32)(g0 + first);
    b.ovf = (u32)(o0 + ovf_off[x]);
    for (int s = 0; s < 5; ++s) b.w[s] = 0;
    const u32 m = cnt < cap ? cnt : cap;
    u32 meta = m << 20;
    for (u32 s = 0; s < m; ++s) {
        const u64 row = rows[(u64)first + s];
        meta |= (u32)(row & 15ull) << (4 * s);
        const u64 pay = row >> 4;
        if (PAY64) {
            b.w[1 + 2 * s] = (u32)pay;
            b.w[2 + 2 * s] = (u32)(pay >> 32);
        } else {
            b.w[s] = (u32)pay;
        }
    }
    if (cnt > cap) {
        meta |= kFoldMetaOvf;
        const u32 n_ovf = cnt - cap;
        u64 *o = orows + ovf_off[x];  // this pass's chunk
        o[0] = n_ovf;
        for (u32 s = 0; s < n_ovf; ++s) o[1 + s] = rows[(u64)first + cap + s];
        for (u32 s = 1 + n_ovf; s < ((1u + n_ovf + 3u) & ~3u); ++s) o[s] = ~0ull;
    }
    b.meta = meta;
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
    u64 *orows = nullptr;
    uint2 *ids = nullptr;
    u64 n_rows = 0, n_orows = 0;
};

// Sort passes: the buffers of a pass (selection 4 B, keys / values and their sort doubles 32 B, group heads 4 B,
// rows 8 B per entry) are kept near `kFoldPassBytes`.
constexpr u64 kFoldPassBytes = 12ull << 30;
inline u32 fold_passes(u64 N, u32 t) {
    u64 p = (48ull * N + kFoldPassBytes - 1) / kFoldPassBytes;
    const u64 buckets = 1ull << (2 * t);
    if (p > buckets) p = buckets;
    if (p > 256) p = 256;
    return (u32)(p < 1 ? 1 : p);
}
// Peak device memory of build_fold_on_device beyond the index itself (bytes), and what stays resident.
inline u64 fold_build_peak_bytes(u64 N, u32 t, bool with_ids) {
    const u64 passes = fold_passes(N, t);
    const u64 sort_stage = 12ull * N + N / 8 + (32ull << (2 * t)) + (with_ids ? 8ull * N : 0ull) + 48ull * N / passes + (12ull << (2 * t)) / passes + N / 2;
    const u64 kmer_stage = 20ull * N + N / 8;  // pointer doubling: k-mers 8N (holds K_l meanwhile), P_l twice, K_16
    return (sort_stage > kmer_stage ? sort_stage : kmer_stage) + (64ull << 20);
}
inline u64 fold_resident_bytes(u64 N, u32 t, bool with_ids) { return (32ull << (2 * t)) + N + (with_ids ? 8ull * N : 0ull); }

// The k-mer of every SA row (kmers[], validbits[]) and the first row of every run of equal k-mers (heads[0, M)): the
// distinct k-mers of the index with their SA intervals. Shared by the dictionary builders (fold.cuh, loc.cuh).
template <typename Stage, typename Lap>
inline void derive_kmer_runs(const DevIndex &d, const u64 counts[4], u32 k, DevArr<u64> &kmers, DevArr<u32> &validbits, DevArr<u32> &heads, u64 &M,
                             uint64_t &nl, Stage stage, Lap lap) {
    const u64 N = d.n;
    kmers.alloc(N);
    validbits.alloc((N >> 5) + 2);
    {
        const u32 c1 = (u32)counts[1], c2 = (u32)counts[2], c3 = (u32)counts[3];
        DevArr<u32> pa(N);
        psi_scatter_kernel<<<nblocks_for(N), 256>>>(d, pa.p);
        stage("psi");
        static const bool walk = std::getenv("FMSI_GPU_FOLD_WALK") != nullptr;  // the k-1 step walk per row (A/B switch)
        if (walk) {
            fold_kmers_kernel<<<nblocks_for(N), 256>>>(N, pa.p, c1, c2, c3, k, kmers.p, validbits.p);
            stage("k-mers of the SA rows");
            nl += 2;
        } else {
            fold_sticky_kernel<<<1, 32>>>(pa.p);
            // K_l lives in the upper / lower half of the (not yet used) kmers array, P_l in pa / pb
            DevArr<u32> pb(N);
            u32 *ka = reinterpret_cast<u32 *>(kmers.p), *kb = ka + N;
            const u32 *pin = pa.p, *kin = nullptr;
            u32 *pout = pb.p, *kout = ka;
            for (u32 l = 1; l < 16; l *= 2) {
                fold_double_kernel<<<nblocks_for(N), 256>>>(N, pin, kin, l, c1, c2, c3, pout, kout);
                stage("k-mer doubling");
                pin = pout;
                kin = kout;
                pout = pout == pb.p ? pa.p : pb.p;
                kout = kout == ka ? kb : ka;
            }
            // K_16 sits in one half of kmers[]: move it out before the 64-bit k-mers overwrite that array
            DevArr<u32> k16(N);
            BCU(cudaMemcpy(k16.p, kin, N * 4, cudaMemcpyDeviceToDevice));
            fold_double_last_kernel<<<nblocks_for(N), 256>>>(N, pin, k16.p, k, kmers.p);
            stage("k-mers of the SA rows");
            lap("pointer doubling");
            fold_valid_fill_kernel<<<nblocks_for((N + 31) / 32), 256>>>(N, validbits.p);
            fold_valid_clear_kernel<<<1, 32>>>(d, k, validbits.p);
            stage("valid rows");
            nl += 9;
        }
    }
    lap("k-mers of the SA rows");
    {
        DevArr<u32> all(N);
        M = fold_select_heads(N, all.p, FoldRunHead{kmers.p, validbits.p});
        stage("run heads");
        if (M * 10 < N * 9) {  // many rows per run (repetitive index): keep an exact-size copy
            heads.alloc(M);
            BCU(cudaMemcpy(heads.p, all.p, M * 4, cudaMemcpyDeviceToDevice));
        } else {
            heads.p = all.p;
            heads.n = all.n;
            all.p = nullptr;
        }
    }
}

// want_table: buckets + overflow rows; want_ids: the lookup ids (row numbering is the same whichever is built).
// Throws std::runtime_error (out of memory included); nothing is leaked then.
inline void build_fold_on_device(const DevIndex &d, const u64 counts[4], u32 k, u32 t, bool want_table, bool want_ids, FoldArrays &out,
                                 uint64_t *launches) {
    const u32 B = k - t;
    const u64 total = 1ull << (2 * t);
    const u32 cap = B > 16 ? kFoldCap64 : kFoldCap32;
    auto stage = [](const char *what) {  // surfaces asynchronous errors with the step that caused them
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
    };
    const bool timing = std::getenv("FMSI_GPU_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    size_t low_free = ~(size_t)0;  // lowest free device memory seen at the stage ends ($FMSI_GPU_TIMING): the build's peak
    auto lap = [&](const char *what) {
        if (!timing) return;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if (free_b < low_free) low_free = free_b;
        fprintf(stderr, "[fmsi timing] fold build: %s at %.3f s (device memory in use %.1f GB, peak so far %.1f GB)\n", what,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), (total_b - free_b) / 1e9, (total_b - low_free) / 1e9);
    };
    uint64_t nl = 0;
    DevArr<u64> kmers;
    DevArr<u32> validbits, heads;
    u64 M = 0;
    derive_kmer_runs(d, counts, k, kmers, validbits, heads, M, nl, stage, lap);
    lap("run heads");
    DevArr<FoldBucket> buckets;
    if (want_table) buckets.alloc(total);
    DevArr<uint2> ids;
    if (want_ids) ids.alloc(M + 1);  // rows <= runs
    std::vector<std::unique_ptr<DevArr<u64>>> ochunks;
    std::vector<u64> ochunk_len;
    const u32 P = fold_passes(M, t);
    // how many runs each pass will select (one streamed pass over the runs), so that its buffers are exact
    std::vector<unsigned long long> h_hist(P, 0);
    {
        DevArr<unsigned long long> hist(P);
        BCU(cudaMemset(hist.p, 0, P * sizeof(unsigned long long)));
        if (M) fold_pass_hist_kernel<<<nblocks_for(M), 256>>>(FoldRunInPass{kmers.p, validbits.p, heads.p, k, B, 0, total}, M, total, P, hist.p);
        stage("pass histogram");
        BCU(cudaMemcpy(h_hist.data(), hist.p, P * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        nl += 1;
    }
    // The buffers of a pass are allocated once, for the largest pass, and reused: allocating and freeing ~10 GB per
    // pass cost more than the sorting itself on boxes where cudaMalloc is slow (first touch of fresh device memory).
    u64 max_mp = 1, max_nb = 1;
    for (u32 p = 0; p < P; ++p) {
        max_mp = std::max<u64>(max_mp, h_hist[p]);
        max_nb = std::max<u64>(max_nb, fold_pass_begin(p + 1, P, total) - fold_pass_begin(p, P, total));
    }
    DevArr<u32> sel(max_mp + 1), gs(max_mp + 1);
    DevArr<u64> keys(max_mp), vals(max_mp), keys_alt(max_mp), vals_alt(max_mp);
    DevArr<u64> rows;
    DevArr<u32> bfirst, bcount, sizes;
    DevArr<u64> off;
    if (want_table) {
        rows.alloc(max_mp + 1);
        bfirst.alloc(max_nb);
        bcount.alloc(max_nb);
        sizes.alloc(max_nb);
        off.alloc(max_nb + 1);
    }
    u64 G0 = 0, O0 = 0;
    for (u32 p = 0; p < P; ++p) {
        const u64 x_lo = fold_pass_begin(p, P, total), x_hi = fold_pass_begin(p + 1, P, total), nb = x_hi - x_lo;
        if (want_table) {
            BCU(cudaMemset(bfirst.p, 0, nb * 4));
            BCU(cudaMemset(bcount.p, 0, nb * 4));
        }
        const u64 Mp = fold_select_heads(M, sel.p, FoldRunInPass{kmers.p, validbits.p, heads.p, k, B, x_lo, x_hi});
        stage("runs of the pass");
        if (Mp != h_hist[p]) throw std::runtime_error("pass histogram and selection disagree");
        if (Mp) fold_entries_kernel<<<nblocks_for(Mp), 256>>>(d, kmers.p, heads.p, M, sel.p, Mp, k, keys.p, vals.p);
        stage("entries");
        radix_sort_pairs(keys, keys_alt, vals, vals_alt, Mp, (int)(2 * k));
        stage("sort");
        if (p == 0 || p + 1 == P) lap(p == 0 ? "first pass sorted" : "last pass sorted");
        const u64 G = fold_select_heads(Mp, gs.p, FoldKeyHead{keys.p});
        stage("group heads");
        if (G)
            fold_rows_kernel<<<nblocks_for(G), 256>>>(keys.p, vals.p, gs.p, G, Mp, B, G0, x_lo, want_table ? rows.p : nullptr,
                                                      want_ids ? ids.p : nullptr, want_table ? bfirst.p : nullptr, want_table ? bcount.p : nullptr);
        stage("rows");
        nl += 5;
        if (want_table) {
            fold_ovf_sizes_kernel<<<nblocks_for(nb), 256>>>(bfirst.p, bcount.p, nb, cap, sizes.p);
            exclusive_sum_u64(sizes.p, off.p, nb);
            u64 last_off = 0;
            u32 last_size = 0;
            BCU(cudaMemcpy(&last_off, off.p + nb - 1, 8, cudaMemcpyDeviceToHost));
            BCU(cudaMemcpy(&last_size, sizes.p + nb - 1, 4, cudaMemcpyDeviceToHost));
            const u64 olen = last_off + last_size;
            if (O0 + olen >= (1ull << 32)) throw std::runtime_error("overflow rows exceed 32-bit offsets");
            ochunks.emplace_back(new DevArr<u64>(olen ? olen : 1));
            ochunk_len.push_back(olen);
            if (B > 16) fold_bucket_fill_kernel<true><<<nblocks_for(nb), 256>>>(rows.p, bfirst.p, bcount.p, off.p, nb, G0, O0, buckets.p + x_lo, ochunks.back()->p);
            else fold_bucket_fill_kernel<false><<<nblocks_for(nb), 256>>>(rows.p, bfirst.p, bcount.p, off.p, nb, G0, O0, buckets.p + x_lo, ochunks.back()->p);
            stage("buckets");
            nl += 3;
            O0 += olen;
        }
        G0 += G;
    }
    sel.release();
    gs.release();
    keys.release();
    vals.release();
    keys_alt.release();
    vals_alt.release();
    rows.release();
    lap("sort passes");
    if (G0 >= (1ull << 32)) throw std::runtime_error("more than 2^32 rows");
    kmers.release();
    heads.release();
    validbits.release();
    DevArr<u64> orows;
    if (want_table) {
        orows.alloc(O0 + 8);
        BCU(cudaMemset(orows.p + O0, 0xff, 8 * sizeof(u64)));
        u64 at = 0;
        for (size_t c = 0; c < ochunks.size(); ++c) {
            if (ochunk_len[c]) BCU(cudaMemcpy(orows.p + at, ochunks[c]->p, ochunk_len[c] * 8, cudaMemcpyDeviceToDevice));
            at += ochunk_len[c];
            ochunks[c].reset();
        }
    }
    if (want_ids && G0 * 10 < M * 9) {  // far fewer rows than runs: exact-size ids
        DevArr<uint2> exact(G0 + 1);
        BCU(cudaMemcpy(exact.p, ids.p, G0 * sizeof(uint2), cudaMemcpyDeviceToDevice));
        std::swap(exact.p, ids.p);
        std::swap(exact.n, ids.n);
    }
    lap("done");
    if (launches) *launches += nl;
    out.buckets = buckets.p;
    out.orows = orows.p;
    out.ids = ids.p;
    out.n_rows = G0;
    out.n_orows = O0;
    buckets.p = nullptr;
    orows.p = nullptr;
    ids.p = nullptr;
}

// ------------------------------------------------------------------------------------------- query
enum { FP_BUCKET = 0, FP_OVF = 1, FP_SEARCH = 2, FP_IDS = 3 };

// A bucket / overflow sector. LD64: ask L2 for a 64-byte fill instead of the whole 128-byte line.
template <bool LD64>
__device__ __forceinline__ void ld_fold_sector(const void *p, u64 &a, u64 &b, u64 &c, u64 &d) {
    if (LD64) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    } else {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    }
}

// Queries taken straight from reads (fmsi_gpu_query_reads_packed): query q of the launch is result slot slot0 + q, the
// k-mer at position slot - rbase[r] of read r = the last read with rbase[r] <= slot. text == nullptr: queries are
// packed k-mers in an array (the default). Saves the slot -> k-mer array that extract_kmers_kernel writes and this
// kernel would read back (16 bytes per k-mer of traffic, against 42 algorithmic ones).
struct ReadSrc {
    const u64 *text;   // 2-bit packed text
    const u64 *roff;   // [n_reads + 1] first base of every read
    const u64 *rbase;  // [n_reads] first result slot of every read (non-decreasing; equal for reads shorter than k)
    u64 n_reads;
    u64 slot0;
};
// read of `slot`, searching forward from read `from` (rbase[from] <= slot): gallop, then bisect
__device__ __forceinline__ u64 read_of_slot(const ReadSrc &rs, u64 slot, u64 from) {
    u64 lo = from, step = 1, hi;
    for (;;) {  // invariant: rbase[lo] <= slot
        hi = lo + step;
        if (hi >= rs.n_reads) {
            hi = rs.n_reads;
            break;
        }
        if (__ldg(rs.rbase + hi) > slot) break;
        lo = hi;
        step <<= 1;
    }
    while (hi - lo > 1) {  // rbase[lo] <= slot < rbase[hi] (hi == n_reads counts as +inf)
        const u64 mid = (lo + hi) >> 1;
        if (__ldg(rs.rbase + mid) <= slot) lo = mid;
        else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ u64 window64(const u64 *__restrict__ text, u64 s, u32 len) {  // len (<= 32) bases from base s
    const u64 w0 = __ldg(text + (s >> 5)), w1 = __ldg(text + (s >> 5) + 1);
    const u32 sh = 2u * ((u32)s & 31u);
    const u64 v = sh ? ((w0 << sh) | (w1 >> (64 - sh))) : w0;
    return v >> (64 - 2 * len);
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
                  unsigned long long *__restrict__ cursor, const u32 chunk, unsigned long long *__restrict__ probe_ctr, const ReadSrc rs) {
    const unsigned FULL = 0xffffffffu;
    const u32 lane = threadIdx.x & 31u;
    const u32 lt_mask = (1u << lane) - 1u;
    const u32 k = fv.k, B = fv.B;
    const u64 pmask = B ? ((1ull << (2 * B)) - 1ull) : 0ull;
    const u32 CAP = PAY64 ? kFoldCap64 : kFoldCap32;

    bool active = false, swapped = false;
    u32 phase = FP_BUCKET;
    u64 idx = 0, q = 0, bx = 0;
    // a bucket's overflow region: entries [ovf + 1, ovf + 1 + n_ovf) of orows[] hold its rows gbase + CAP, ...;
    // a search narrows [lo, hi) (entry numbers)
    u32 lo = 0, hi = 0, row = 0, st = 0, ovf = 0, gbase = 0;
    u64 cend = 0, wnext = 0, tile_base = 0, bufA = 0, bufB = 0;
    bool exhausted = false;
    u32 nprobe = 0;  // dependent memory requests issued by this lane (reported when probe_ctr is given)
    const bool from_reads = rs.text != nullptr;
    u64 rd_hint = 0;  // from_reads: a read at or before the one of the next tile's first query (warp-uniform)
    // query q of the launch, q < cend (0 beyond): from the k-mer array, or cut out of its read
    auto fetch = [&](u64 q) -> u64 {
        if (q >= cend) return 0ull;
        if (!from_reads) return kmers[q];
        const u64 slot = rs.slot0 + q;
        const u64 r = read_of_slot(rs, slot, rd_hint);
        return window64(rs.text, __ldg(rs.roff + r) + (slot - __ldg(rs.rbase + r)), k);
    };
    // after a tile load: the next tile starts at or after this tile's first query, whose read lane 0 just found
    auto advance_hint = [&](u64 first_q) {
        if (!from_reads) return;
        u64 r = 0;
        if (lane == 0 && first_q < cend) r = read_of_slot(rs, rs.slot0 + first_q, rd_hint);
        r = __shfl_sync(FULL, r, 0);
        if (first_q < cend) rd_hint = r;
    };

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
                    if (from_reads) {  // a new grab may lie anywhere: search from the first read
                        rd_hint = 0;
                        advance_hint(tile_base);
                    }
                    bufA = fetch(tile_base + lane);
                    bufB = fetch(tile_base + 32 + lane);
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
                    advance_hint(tile_base);
                    bufB = fetch(tile_base + 32 + lane);
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
        const bool isO = active && phase == FP_OVF;
        const bool isS = active && phase == FP_SEARCH;
        const bool isI = active && phase == FP_IDS;
        const u32 r0 = isO ? ovf : ((lo + ((hi - lo) >> 1)) & ~3u);  // overflow sector probed (entry number)
        u64 a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        if (isB) ld_fold_sector<LD64>(reinterpret_cast<const char *>(fv.buckets) + (bx << 5), a0, a1, a2, a3);
        if (isO || isS) ld_fold_sector<LD64>(fv.orows + r0, a0, a1, a2, a3);
        if (isI) a0 = __ldg(reinterpret_cast<const u64 *>(fv.ids) + row);
        nprobe += (u32)(isB || isO || isS || isI);

        // ---------------------------------------------------------------- consume
        bool done = false;  // (st, row) final
        if (isB) {
            gbase = (u32)a0;
            ovf = (u32)(a0 >> 32);
            const u32 meta = (u32)a1;
            const u32 m = (meta >> 20) & 7u;
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
                        st = (meta >> (4 * s)) & 15u;
                        row = gbase + s;
                    }
                }
            }
            // rows are sorted by payload: the inline ones are the bucket's smallest
            if (found || !(meta & kFoldMetaOvf) || last > q) done = true;
            else phase = FP_OVF;
        } else if (isO || isS) {
            // entries [w0, w1) of this sector are rows; invariant of a search: rows before lo have payload < q,
            // rows from hi on have payload > q
            u32 w0, w1;
            if (isO) {  // first sector of the region: a0 = number of overflow rows, then the first three
                lo = ovf + 1;
                hi = ovf + 1 + (u32)a0;
                w0 = lo;
                w1 = hi < ovf + 4 ? hi : ovf + 4;
            } else {
                w0 = r0 > lo ? r0 : lo;
                w1 = (r0 + 4 < hi) ? r0 + 4 : hi;
            }
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
                        row = gbase + CAP + (r - ovf - 1);
                    }
                }
            }
            if (found) done = true;
            else if (last < q) lo = w1;
            else if (first > q) hi = w0;
            else done = true;  // q falls between two rows of this sector: absent (st stays 0)
            if (!done && lo >= hi) done = true;
            if (!done) phase = FP_SEARCH;
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
