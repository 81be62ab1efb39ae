// loc.cuh — the minimizer-bucketed k-mer dictionary: neighbouring k-mers of a read share their memory requests.
//
// Why. The strand-folded dictionary (fold.cuh) answers a k-mer with ONE random 32-byte probe, and that probe is what a
// query costs: B200 serves ~36-44 G requests/s that miss L2, whatever they return. Single k-mers cannot do better. The
// k-mers of a READ can: consecutive k-mers overlap in k - 1 bases, and the lanes of a warp hold consecutive k-mers. If
// the bucket of a k-mer depends only on a substring that neighbours share — its MINIMIZER, the m-mer of smallest hash
// among the m-mers of the k-mer and of its reverse complement — then a run of ~(k - m + 2) / 2 neighbouring k-mers (a
// super-k-mer) lands in one bucket, their lanes ask for the same sectors, and the load unit merges equal addresses of a
// warp instruction into one request (tools/locbw.cu, profiles/r02u_locbw_b200.jsonl: 36 G k-mers/s for one random
// sector per lane, 67 / 79-98 G k-mers/s when 4 / 8 neighbouring lanes share a directory entry and a line of rows).
//
// Rows are the same as in fold.cuh — one per distinct k-mer of the index in a strand-neutral orientation, carrying the
// states (sf, sr) of the two SA intervals from which query_kmers_single (src/fms_index.h:263-331) derives every answer —
// but keyed and laid out differently:
//   pick(q)   the minimizer of {q, rc(q)}: the place whose canonical m-mer (the smaller of the m-mer and its reverse
//             complement) has the smallest ordering hash; o = the strand on which it reads canonical, pos = its position
//             in o (ties: loc_pick below — a rule that does not depend on which strand the caller holds)
//   h'        a bijection of the m-mer on 2m bits (not the ordering hash: a minimum is a small number)
//   bucket    top 2t bits of h'
//   R         pos . h' low 2m - 2t bits . the k - m bases of o around the m-mer        (2k - 2t + pbits bits)
//             (bucket, R) <-> o is one-to-one, so a row match is exact — no fingerprints
//   dir[x]    first row (32 bits) | row count (21) | simple (1) | smallest pos (5) | pos span (5)   8 bytes per bucket
//   rows[]    R << 4 | sf | sr << 2, sorted by (bucket, R)                                   8 bytes per distinct k-mer
// A query reads dir[bucket]; a pos outside the bucket's range is absent at once; otherwise the bucket's rows — sorted by
// pos first — are probed where pos interpolates to, a sector (4 rows) at a time, and bisected from there. 2-3 dependent
// requests for a k-mer on its own — worse than fold.cuh — but a fraction of a request per k-mer when the 32 lanes of a
// warp hold 32 consecutive k-mers of a read. So this tier only answers text-derived queries (reads / chunks, presence
// outputs); packed single k-mers stay on fold.cuh.
//
// Built on the device like fold.cuh (shared: the k-mer of every SA row by pointer doubling, the runs of equal k-mers);
// entries are sorted by their 64 low key bits in passes over aligned bucket ranges, which makes the bits a 2k + pbits > 64
// bit key loses constant within a pass.
#pragma once
#include "fold.cuh"

namespace fmsi {

struct LocGeom {
    u32 k, m, t;
    u32 w;      // k - m + 1 candidate positions per strand
    u32 pbits;  // bits of pos
    u32 fbits;  // 2 (k - m)
    u32 hlow;   // 2m - 2t: hash bits kept in R
    u32 rbits;  // hlow + pbits + fbits
};
__host__ __device__ __forceinline__ LocGeom loc_geom(u32 k, u32 m, u32 t) {
    LocGeom g;
    g.k = k;
    g.m = m;
    g.t = t;
    g.w = k - m + 1;
    g.pbits = 0;
    while ((1u << g.pbits) < g.w) ++g.pbits;
    g.fbits = 2 * (k - m);
    g.hlow = 2 * m - 2 * t;
    g.rbits = g.hlow + g.pbits + g.fbits;
    return g;
}
// minimizer length for an index of N rows and k-mers of length k: long enough that an m-mer rarely repeats in the text
// (4^m >= N), short enough to leave a window of up to 16 positions; 0 = the tier does not apply
inline u32 loc_pick_m(u64 N, u32 k) {
    u32 lg = 1;
    while (lg < 16 && (1ull << (2 * lg)) < N) ++lg;
    u32 m = k > 15 ? k - 15 : 1;
    if (m < lg) m = lg;
    if (m > 16) m = 16;
    if (m > k) m = k;
    return m;
}
// can the tier be built at bucket depth t? rows must hold R and 4 state bits; the hash must cover the bucket bits
inline bool loc_fits(u32 k, u32 m, u32 t) {
    if (k < 1 || k > 32 || m < 1 || m > 16 || m > k || t < 1 || t > m) return false;
    return loc_geom(k, m, t).rbits + 4 <= 64;
}

struct LocView {
    const u64 *dir;    // [4^t]
    const u64 *rows;   // [n_rows]
    u64 n_rows;
    LocGeom g;
    u32 enabled;
};

constexpr u32 kLocMul1 = 0x9E3779B1u, kLocMul3 = 0xC2B2AE35u, kLocMul4 = 0x27D4EB2Fu;
// ordering value of a canonical m-mer (an injection: equal values <=> equal m-mers)
__host__ __device__ __forceinline__ u32 loc_order(u32 c) { return c * kLocMul1; }
// a bijection of [0, 4^m), m <= 16: the key hash of a canonical m-mer
__host__ __device__ __forceinline__ u32 loc_spread(u32 x, u32 m, u32 mask) {
    x = (x * kLocMul3) & mask;
    x ^= x >> m;
    x = (x * kLocMul4) & mask;
    x ^= x >> m;
    return x;
}
// reverse complement of an m-mer (m <= 16)
__host__ __device__ __forceinline__ u32 loc_revcomp_m(u32 x, u32 m) {
    x = ~x;
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    x = (x >> 16) | (x << 16);
    return x >> (32 - 2 * m);
}

// The minimizer of the strand pair {q, rc}. At position p of q sits the m-mer x_p; the same place read on the other strand
// is y_p = revcomp(x_p), at position w - 1 - p of rc. The place with the smallest order(min(x_p, y_p)) wins, and the k-mer
// is oriented so that the winning m-mer reads as its canonical (smaller) form: (o, pos) = (q, p) if x_p < y_p, (rc, w-1-p)
// if y_p < x_p. Ties — the same canonical m-mer at several places, or a palindromic one (x_p == y_p: both orientations) —
// go to the candidate (o, pos) with the smallest pos, then to the strand with the smaller q * C: a rule stated in terms of
// the physical strands, so both strands of a pair pick the same (o, pos).
// Returns the canonical m-mer x, o, pos and sw (o is not the caller's q).
__host__ __device__ __forceinline__ void loc_pick(u64 q, u64 rc, const LocGeom &g, u32 &x, u32 &pos, bool &sw, u64 &o) {
    const u32 mask = g.m < 16 ? (1u << (2 * g.m)) - 1u : 0xFFFFFFFFu;
    u32 best = 0xFFFFFFFFu;
    for (u32 p = 0; p < g.w; ++p) {
        const u32 xf = (u32)(q >> (g.fbits - 2 * p)) & mask, yr = (u32)(rc >> (2 * p)) & mask;
        const u32 v = loc_order(xf < yr ? xf : yr);
        best = v < best ? v : best;
    }
    const u32 lq = (rc * kFoldMul < q * kFoldMul) ? 1u : 0u;  // label of strand q (0 = the strand with the smaller q * C)
    u32 bkey = 0xFFFFFFFFu;
    for (u32 p = 0; p < g.w; ++p) {
        const u32 xf = (u32)(q >> (g.fbits - 2 * p)) & mask, yr = (u32)(rc >> (2 * p)) & mask;
        if (loc_order(xf < yr ? xf : yr) != best) continue;
        if (xf <= yr) {  // candidate (q, p)
            const u32 key = 2 * p + lq;
            bkey = key < bkey ? key : bkey;
        }
        if (yr <= xf) {  // candidate (rc, w - 1 - p)
            const u32 key = 2 * (g.w - 1 - p) + (1u - lq);
            bkey = key < bkey ? key : bkey;
        }
    }
    pos = bkey >> 1;
    sw = (bkey & 1u) != lq;
    o = sw ? rc : q;
    x = (u32)(o >> (g.fbits - 2 * pos)) & mask;
}
// bucket and R of a pick
__host__ __device__ __forceinline__ void loc_key(u32 x, u32 pos, u64 o, const LocGeom &g, u32 &bucket, u64 &R) {
    const u32 mask = g.m < 16 ? (1u << (2 * g.m)) - 1u : 0xFFFFFFFFu;
    const u32 h = loc_spread(x, g.m, mask);
    const u32 rl = g.fbits - 2 * pos;  // bits of the bases behind the m-mer
    const u64 right = rl ? (o & ((1ull << rl) - 1ull)) : 0ull;
    const u64 left = pos ? (o >> (2 * (g.k - pos))) : 0ull;
    const u64 flanks = rl ? ((left << rl) | right) : left;
    const u64 hl = g.hlow ? (u64)(h & ((1u << g.hlow) - 1u)) : 0ull;  // hlow <= 30
    bucket = h >> g.hlow;
    R = ((u64)pos << (g.hlow + g.fbits)) | (hl << g.fbits) | flanks;
}
__host__ __device__ __forceinline__ u32 loc_pos_of_row(u64 row, const LocGeom &g) { return (u32)((row >> 4) >> (g.hlow + g.fbits)); }

// directory entry: first row (32) | rows (21) | simple (1) | smallest pos (5) | pos span (5). simple = exactly one row for
// every pos of the range, so the row of a pos is first + (pos - smallest pos), and a mismatch there means absent.
constexpr u32 kLocCountBits = 21;
__host__ __device__ __forceinline__ u64 loc_dir_entry(u32 first, u32 count, bool simple, u32 pmin, u32 pspan) {
    return (u64)first | ((u64)count << 32) | ((u64)simple << 53) | ((u64)pmin << 54) | ((u64)pspan << 59);
}

// ------------------------------------------------------------------------------------------- build
constexpr unsigned char kLocNoPass = 0xFF;

// pass of every run (its bucket's range; kLocNoPass for runs of suffixes shorter than k) + how many runs each pass takes
__global__ void loc_pass_kernel(const u64 *__restrict__ kmers, const u32 *__restrict__ validbits, const u32 *__restrict__ heads, const u64 n_runs,
                                const LocGeom g, const u32 pass_shift, unsigned char *__restrict__ pass_of, unsigned long long *__restrict__ hist) {
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const u64 r = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (r < n_runs) {
        const u32 i = heads[r];
        unsigned char p = kLocNoPass;
        if ((validbits[i >> 5] >> (i & 31u)) & 1u) {
            const u64 q = kmers[i];
            u32 h, pos, bucket;
            bool sw;
            u64 o, R;
            loc_pick(q, revcomp_packed(q, g.k), g, h, pos, sw, o);
            loc_key(h, pos, o, g, bucket, R);
            p = (unsigned char)(bucket >> pass_shift);
            atomicAdd(&sh[p], 1u);
        }
        pass_of[r] = p;
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}
struct LocRunInPass {
    const unsigned char *pass_of;
    unsigned char p;
    __host__ __device__ __forceinline__ bool operator()(const u32 r) const { return pass_of[r] == p; }
};

// One entry per selected run: key = the low 64 bits of bucket . R, val = state << 32 | (the run is the reverse complement
// of its row's orientation) << 34 | self-complementary << 35.
__global__ void loc_entries_kernel(const DevIndex d, const u64 *__restrict__ kmers, const u32 *__restrict__ heads, const u64 n_heads,
                                   const u32 *__restrict__ sel, const u64 n_sel, const LocGeom g, u64 *__restrict__ keys, u64 *__restrict__ vals) {
    const u64 e = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (e >= n_sel) return;
    const u64 r = sel[e];
    const u64 i = heads[r], j = (r + 1 < n_heads) ? (u64)heads[r + 1] : d.n;
    const u64 q = kmers[i];
    const u64 rc = revcomp_packed(q, g.k);
    u32 h, pos, bucket;
    bool sw;
    u64 o, R;
    loc_pick(q, rc, g, h, pos, sw, o);
    loc_key(h, pos, o, g, bucket, R);
    const u64 ri = fold_rank1(d, i), rj = fold_rank1(d, j);
    const bool first = (d.aux[i >> 6].mask >> (i & 63)) & 1ull;
    const u64 state = rj > ri ? (first ? 3 : 2) : 1;
    keys[e] = ((u64)bucket << g.rbits) | R;
    vals[e] = (state << 32) | ((u64)sw << 34) | ((u64)(q == rc) << 35);
}

// One row per group of equal keys (the run of the row's own orientation and / or the run of its reverse complement).
// bfirst / bcount (zeroed before): first local row and end of every bucket of the pass; lmask = bucket bits a key keeps.
__global__ void loc_rows_kernel(const u64 *__restrict__ keys, const u64 *__restrict__ vals, const u32 *__restrict__ gs, const u64 n_groups,
                                const u64 n_entries, const LocGeom g, const u64 x_lo, const u64 lmask, u64 *__restrict__ rows,
                                u32 *__restrict__ bfirst, u32 *__restrict__ bcount) {
    const u64 gi = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (gi >= n_groups) return;
    const u64 e0 = gs[gi], e1 = (gi + 1 < n_groups) ? (u64)gs[gi + 1] : n_entries;
    const u64 key = keys[e0];
    u32 sf = 0, sr = 0;
    bool self_rc = false;
    for (u64 e = e0; e < e1; ++e) {
        const u64 v = vals[e];
        const u32 st = (u32)(v >> 32) & 3u;
        self_rc |= (v >> 35) & 1ull;
        if ((v >> 34) & 1ull) sr = st;
        else sf = st;
    }
    if (self_rc) sr = sf;
    const u64 rmask = (1ull << g.rbits) - 1ull;
    rows[gi] = ((key & rmask) << 4) | sf | (sr << 2);
    const u64 x = ((key >> g.rbits) - x_lo) & lmask;
    const bool first = gi == 0 || (((keys[gs[gi - 1]] >> g.rbits) - x_lo) & lmask) != x;
    const bool last = gi + 1 == n_groups || (((keys[e1] >> g.rbits) - x_lo) & lmask) != x;
    if (first) bfirst[x] = (u32)gi;
    if (last) bcount[x] = (u32)(gi + 1);
}
// rows: this pass's rows (local numbering); a bucket with more rows than the entry can count raises *too_many
__global__ void loc_dir_kernel(const u32 *__restrict__ bfirst, const u32 *__restrict__ bcount, const u64 *__restrict__ rows, const u64 nb, const u64 g0,
                               const LocGeom g, u64 *__restrict__ dir, u32 *__restrict__ too_many) {
    const u64 x = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (x >= nb) return;
    const u32 cnt = bcount[x] ? bcount[x] - bfirst[x] : 0u;
    if (!cnt) {
        dir[x] = 0ull;
        return;
    }
    if (cnt >> kLocCountBits) atomicExch(too_many, 1u);
    const u32 pmin = loc_pos_of_row(rows[bfirst[x]], g), pmax = loc_pos_of_row(rows[bfirst[x] + cnt - 1], g);  // rows are sorted by pos first
    bool simple = cnt == pmax - pmin + 1;
    for (u32 j = 1; simple && j + 1 < cnt; ++j) simple = loc_pos_of_row(rows[bfirst[x] + j], g) == pmin + j;
    dir[x] = loc_dir_entry((u32)(g0 + bfirst[x]), cnt, simple, pmin, pmax - pmin);
}

struct LocArrays {  // device arrays of a built tier (ownership passes to the caller)
    u64 *dir = nullptr;
    u64 *rows = nullptr;
    u64 n_rows = 0;
};
inline u32 loc_passes(u64 M, const LocGeom &g) {
    const u32 key_bits = 2 * g.t + g.rbits;
    u32 p = 1;
    while ((u64)p * kFoldPassBytes < 48ull * M && p < 128) p <<= 1;
    if (key_bits > 64)
        while (p < (1u << (key_bits - 64))) p <<= 1;
    return p;
}
// Peak device memory of build_loc_on_device beyond the index itself (bytes), and what stays resident.
inline u64 loc_resident_bytes(u64 N, u32 t) { return (8ull << (2 * t)) + 8ull * N; }
inline u64 loc_build_peak_bytes(u64 N, u32 t) {
    const u64 sort_stage = 12ull * N + N / 8 + N + loc_resident_bytes(N, t) + 12ull * (1ull << 30) + 48ull * N / 8;
    const u64 kmer_stage = 20ull * N + N / 8;
    return (sort_stage > kmer_stage ? sort_stage : kmer_stage) + (64ull << 20);
}

// Throws std::runtime_error (out of memory included); nothing is leaked then.
inline void build_loc_on_device(const DevIndex &d, const u64 counts[4], u32 k, u32 m, u32 t, LocArrays &out, uint64_t *launches) {
    const LocGeom g = loc_geom(k, m, t);
    const u64 total = 1ull << (2 * t);
    auto stage = [](const char *what) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
    };
    const bool timing = std::getenv("FMSI_GPU_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        fprintf(stderr, "[fmsi timing] locality build: %s at %.3f s (device memory in use %.1f GB)\n", what,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), (total_b - free_b) / 1e9);
    };
    uint64_t nl = 0;
    DevArr<u64> kmers;
    DevArr<u32> validbits, heads;
    u64 M = 0;
    derive_kmer_runs(d, counts, k, kmers, validbits, heads, M, nl, stage, lap);
    lap("run heads");
    const u32 P = loc_passes(M, g);
    if (P > total) throw std::runtime_error("bucket depth too small for the sort passes");
    u32 pass_shift = 0;  // pass of bucket x = x >> pass_shift
    while ((total >> pass_shift) > P) ++pass_shift;
    const u64 nb = total / P;
    DevArr<unsigned char> pass_of(M + 1);
    std::vector<unsigned long long> h_hist(256, 0);
    {
        DevArr<unsigned long long> hist(256);
        BCU(cudaMemset(hist.p, 0, 256 * sizeof(unsigned long long)));
        if (M) loc_pass_kernel<<<nblocks_for(M), 256>>>(kmers.p, validbits.p, heads.p, M, g, pass_shift, pass_of.p, hist.p);
        stage("passes of the runs");
        BCU(cudaMemcpy(h_hist.data(), hist.p, 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        nl += 1;
    }
    validbits.release();
    u64 max_mp = 1, sum = 0;
    for (u32 p = 0; p < P; ++p) {
        max_mp = std::max<u64>(max_mp, h_hist[p]);
        sum += h_hist[p];
    }
    DevArr<u64> dir(total);
    DevArr<u32> too_many(1);
    BCU(cudaMemset(too_many.p, 0, 4));
    DevArr<u64> rows(sum + 8);  // rows <= valid runs; a probe reads up to 8 rows from a sector boundary
    DevArr<u32> sel(max_mp + 1), gs(max_mp + 1);
    DevArr<u64> keys(max_mp), vals(max_mp), keys_alt(max_mp), vals_alt(max_mp);
    DevArr<u32> bfirst(nb), bcount(nb);
    const u32 key_bits = 2 * g.t + g.rbits;
    const u64 lmask = nb - 1;  // a pass covers an aligned power-of-two range of buckets
    u64 G0 = 0;
    for (u32 p = 0; p < P; ++p) {
        const u64 x_lo = (u64)p * nb;
        BCU(cudaMemset(bfirst.p, 0, nb * 4));
        BCU(cudaMemset(bcount.p, 0, nb * 4));
        const u64 Mp = fold_select_heads(M, sel.p, LocRunInPass{pass_of.p, (unsigned char)p});
        stage("runs of the pass");
        if (Mp != h_hist[p]) throw std::runtime_error("pass histogram and selection disagree");
        if (Mp) loc_entries_kernel<<<nblocks_for(Mp), 256>>>(d, kmers.p, heads.p, M, sel.p, Mp, g, keys.p, vals.p);
        stage("entries");
        radix_sort_pairs(keys, keys_alt, vals, vals_alt, Mp, (int)(key_bits < 64 ? key_bits : 64));
        stage("sort");
        if (p == 0 || p + 1 == P) lap(p == 0 ? "first pass sorted" : "last pass sorted");
        const u64 G = fold_select_heads(Mp, gs.p, FoldKeyHead{keys.p});
        stage("group heads");
        if (G0 + G > sum) throw std::runtime_error("more rows than runs");
        if (G) loc_rows_kernel<<<nblocks_for(G), 256>>>(keys.p, vals.p, gs.p, G, Mp, g, x_lo, lmask, rows.p + G0, bfirst.p, bcount.p);
        stage("rows");
        loc_dir_kernel<<<nblocks_for(nb), 256>>>(bfirst.p, bcount.p, rows.p + G0, nb, G0, g, dir.p + x_lo, too_many.p);
        stage("directory");
        nl += 6;
        G0 += G;
    }
    if (G0 >= (1ull << 32)) throw std::runtime_error("more than 2^32 rows");
    {
        u32 h_too_many = 0;
        BCU(cudaMemcpy(&h_too_many, too_many.p, 4, cudaMemcpyDeviceToHost));
        if (h_too_many) throw std::runtime_error("a minimizer is shared by more than 2^21 distinct k-mers");
    }
    BCU(cudaMemset(rows.p + G0, 0xff, (sum + 8 - G0) * sizeof(u64)));  // a probe may run past the last row
    sel.release();
    gs.release();
    keys.release();
    vals.release();
    keys_alt.release();
    vals_alt.release();
    kmers.release();
    heads.release();
    pass_of.release();
    if (G0 * 10 < sum * 9) {  // far fewer rows than runs: exact-size rows
        DevArr<u64> exact(G0 + 8);
        BCU(cudaMemcpy(exact.p, rows.p, G0 * sizeof(u64), cudaMemcpyDeviceToDevice));
        BCU(cudaMemset(exact.p + G0, 0xff, 8 * sizeof(u64)));
        std::swap(exact.p, rows.p);
        std::swap(exact.n, rows.n);
    }
    lap("done");
    if (launches) *launches += nl;
    out.dir = dir.p;
    out.rows = rows.p;
    out.n_rows = G0;
    dir.p = nullptr;
    rows.p = nullptr;
}

// ------------------------------------------------------------------------------------------- query
__device__ __forceinline__ u64 ld_loc_dir(const u64 *p) {
    u64 r;
    asm volatile("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}

// The query kernel walks TILES: 32 consecutive k-mers of one chunk of text (a read), one per lane.
//   * Minimizers: lane l hashes ONE m-mer — the one that starts at the tile's base l (lanes below w - 1 a second one, at
//     base 32 + l) — and the minimum of every lane's window of w places comes out of a sliding-window minimum ACROSS the
//     lanes (log2 w rounds of shuffles), instead of 2 w hashes per lane. Ties — equal ordering bits in one window,
//     palindromic m-mers — are rare and go through loc_pick, the one exact definition (as do windows that are not a
//     power of two: small k).
//   * Lookups: a fixed two-load pipeline per tile — DIR (the bucket's entry: empty bucket or a pos outside its range is
//     absent at once; else where pos interpolates to), WIDE (8 rows around that place). Lanes that hold neighbouring
//     k-mers ask for the same entry and the same lines of rows in one instruction, which the load unit merges into one
//     request. A lane whose window does not settle its k-mer (a bucket of many rows) parks that search in a per-warp
//     queue in shared memory; whenever 32 are waiting the warp finishes them together by bisection.
//   * Overlap: WIDE of tile i and DIR of tile i + 1 are in flight while tile i + 2 is fetched and keyed.
constexpr int kLocBlock = 256;
#ifndef FMSI_LOC_MINBLOCKS
#define FMSI_LOC_MINBLOCKS 3
#endif
#ifndef FMSI_LOC_ROWS
#define FMSI_LOC_ROWS 8
#endif
constexpr u32 kLocRows = FMSI_LOC_ROWS;  // rows per probe: 8 (two sectors) or 4 (one)

__device__ __forceinline__ void loc_probe8(const u64 *rows, u32 r0, u64 (&v)[kLocRows]) {
    ld_sector_l1(rows + r0, v[0], v[1], v[2], v[3]);
    if (kLocRows == 8) ld_sector_l1(rows + r0 + 4, v[kLocRows - 4], v[kLocRows - 3], v[kLocRows - 2], v[kLocRows - 1]);
}
__device__ __forceinline__ u64 ld_loc_row(const u64 *p) {
    u64 r;
    asm volatile("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}
// first row of the 8-row window (sector aligned) around row c of [lo, hi)
__device__ __forceinline__ u32 loc_window(u32 c, u32 lo) {
    if (kLocRows == 4) return c & ~3u;
    u32 r0 = (c >= 2u ? c - 2u : 0u) & ~3u;
    const u32 l4 = lo & ~3u;
    return r0 < l4 ? l4 : r0;
}
// rows [r0, r0 + 8) against R4 = R << 4 inside [lo, hi): true = settled (st = the row's states, 0 = no such row); else
// [lo, hi) shrinks. A row is key << 4 | states, so key < R <=> row < R4 and key == R <=> R4 <= row < R4 + 16.
__device__ __forceinline__ bool loc_consume8(const u64 (&v)[kLocRows], u32 r0, u64 R4, u32 &lo, u32 &hi, u32 &st) {
    const u32 w0 = r0 > lo ? r0 : lo, w1 = (r0 + kLocRows < hi) ? r0 + kLocRows : hi;
    const u32 vm = ((1u << (w1 - r0)) - 1u) & ~((1u << (w0 - r0)) - 1u);  // rows of the window that belong to [lo, hi)
    u32 n_lt = 0;
    bool found = false;
#pragma unroll
    for (u32 e = 0; e < kLocRows; ++e) {
        const bool valid = (vm >> e) & 1u;
        n_lt += (u32)(valid && v[e] < R4);
        if (valid && (v[e] - R4) < 16ull) {
            found = true;
            st = (u32)v[e] & 15u;
        }
    }
    if (found) return true;
    if (n_lt == w1 - w0) lo = w1;  // every row of the window is smaller
    else if (n_lt == 0) hi = w0;   // every row is larger
    else return true;              // R falls between two rows: absent
    return lo >= hi;
}

template <int MODE, int STRANDS>
__device__ __forceinline__ unsigned char loc_result(u32 st, bool swapped) {
    u32 sf = st & 3u, sr = st >> 2;
    if (swapped) {
        const u32 x = sf;
        sf = sr;
        sr = x;
    }
    const int vf = fold_presence<MODE>(sf), vr = fold_presence<MODE>(sr);
    if (STRANDS == K_STRANDS_BOTH) return (unsigned char)((vf + 1) | ((vr + 1) << 2));
    if (MODE == K_MODE_ALL) return (unsigned char)((vf != -1 ? vf : vr) == 1);  // fms_index.h:294-298
    return (unsigned char)(vf == 1 || vr == 1);                                  // :289-293
}

constexpr u32 kLocTailCap = 64;  // parked searches per warp
struct LocTailQ {
    u64 q[kLocTailCap];   // result slot | swapped << 63
    u64 R4[kLocTailCap];
    u32 lo[kLocTailCap], hi[kLocTailCap];
};
// window entry: top 25 bits of the ordering value | tie << 6 | m-mer position inside the tile (0 .. 62)
constexpr u32 kLocTie = 64u, kLocInf = 0xFFFFFFFFu;
__device__ __forceinline__ u32 loc_entry(u32 ord, u32 p) { return (ord & ~127u) | p; }
// the better of two entries, a covering the earlier positions
__device__ __forceinline__ u32 loc_combine(u32 a, u32 b) {
    const u32 ka = a >> 7, kb = b >> 7;
    return ka < kb ? a : (kb < ka ? b : (a | kLocTie));
}

// Presence outputs only (K_OUT_PRESENCE). Chunk c = bases [coff[c], coff[c] + clen[c]) of the 2-bit packed text; its k-mers
// go to result slots roff[c], roff[c] + 1, ... (the layout of stream_kernel; chunks may be of any length, and a caller that
// cuts reads into chunks of 32 k-mers gets full tiles). Dynamic shared memory: loc_stream_smem().
template <int MODE, int STRANDS>
__global__ void __launch_bounds__(kLocBlock, FMSI_LOC_MINBLOCKS)
loc_stream_kernel(const LocView lv, const u64 *__restrict__ packed, const u64 n_bases, const u64 *__restrict__ coff,
                  const u32 *__restrict__ clen, const u64 *__restrict__ roff, const u64 n_chunks, unsigned char *__restrict__ out,
                  unsigned long long *__restrict__ cursor, const u32 grab, unsigned long long *__restrict__ probe_ctr) {
    extern __shared__ __align__(8) u32 loc_smem[];
    LocTailQ &tq = reinterpret_cast<LocTailQ *>(loc_smem)[threadIdx.x >> 5];
    const unsigned FULL = 0xffffffffu;
    const u32 lane = threadIdx.x & 31u;
    const u32 lt_mask = (1u << lane) - 1u;
    const LocGeom g = lv.g;
    const u32 k = g.k, W = g.w, m = g.m;
    const u32 mask = m < 16 ? (1u << (2 * m)) - 1u : 0xFFFFFFFFu;
    const bool pow2 = (W & (W - 1)) == 0;  // the sliding-window minimum below takes windows of 1, 2, 4, 8, 16 places

    // the warp's place in the work: chunks [cc, cend) of its grab, tiles of the current chunk from k-mer cb on (warp-uniform)
    u64 cc = 0, cend = 0, cs = 0, cres = 0;
    u32 cnk = 0, cb = 0;
    bool exhausted = false;
    u32 qcount = 0;  // parked searches (warp-uniform)
    u32 nreq = 0;    // requests after merging, counted by the leader lanes when probe_ctr is given

    auto count_requests = [&](bool mine, const void *addr) {  // distinct 128-byte lines of a warp instruction
        if (!probe_ctr) return;
        const unsigned act = __ballot_sync(FULL, mine);
        if (mine) {
            const unsigned same = __match_any_sync(act, (unsigned long long)addr >> 7);
            if ((u32)(__ffs(same) - 1) == lane) ++nreq;
        }
    };
    // finish up to 32 parked searches together (bisection, 8 rows per probe)
    auto drain = [&]() {
        const u32 nb = qcount < 32u ? qcount : 32u;
        const u32 e = qcount - nb + lane;
        bool busy = lane < nb;
        u64 q = 0, R4 = 0;
        u32 lo = 0, hi = 0, st = 0;
        if (busy) {
            q = tq.q[e];
            R4 = tq.R4[e];
            lo = tq.lo[e];
            hi = tq.hi[e];
        }
        qcount -= nb;
        while (__any_sync(FULL, busy)) {
            u64 v[kLocRows] = {};
            const u32 r0 = loc_window(lo + ((hi - lo) >> 1), lo);
            if (busy) loc_probe8(lv.rows, r0, v);
            count_requests(busy, lv.rows + r0);
            if (busy && loc_consume8(v, r0, R4, lo, hi, st)) {
                out[q & ~(1ull << 63)] = loc_result<MODE, STRANDS>(st, (q >> 63) != 0);
                busy = false;
            }
        }
        __syncwarp();
    };
    // the next tile: its k-mers, their minimizers and keys. false = no tiles left (warp-uniform)
    auto prepare = [&](u64 &res, bool &have, bool &sw, u32 &bucket, u32 &pos, u64 &R4) -> bool {
        while (cb >= cnk) {  // next chunk
            if (cc >= cend) {
                unsigned long long c0 = 0;
                if (lane == 0) c0 = atomicAdd(cursor, (unsigned long long)grab);
                c0 = __shfl_sync(FULL, c0, 0);
                if (c0 >= n_chunks) {
                    exhausted = true;
                    return false;
                }
                cc = c0;
                cend = (c0 + grab < n_chunks) ? c0 + grab : n_chunks;
            }
            const u64 s0 = __ldg(coff + cc);
            const u32 len = __ldg(clen + cc);
            cres = __ldg(roff + cc);
            cs = s0;
            cnk = (len >= k && s0 + len <= n_bases) ? len - k + 1 : 0u;
            cb = 0;
            ++cc;
        }
        const u64 s = cs + cb;
        const u32 nt = cnk - cb < 32u ? cnk - cb : 32u;  // k-mers of this tile; its m-mers sit at bases 0 .. nt + W - 2
        res = cres + cb;
        cb += 32;
        have = lane < nt;
        const u64 full = window64(packed, s + lane, 32);  // 32 bases from base `lane` of the tile (the text has that much slack)
        const u64 kf = full >> (64 - 2 * k), kr = revcomp_packed(kf, k);
        u32 x = 0, wm;
        u64 o, R;
        if (pow2) {
            // entries of the m-mers at bases lane and 32 + lane
            const u32 x1 = (u32)(full >> (64 - 2 * m)), y1 = (u32)kr & mask;
            const u64 f31 = __shfl_sync(FULL, full, 31);  // bases 31 .. 62
            const u32 x2 = (u32)((f31 << (2 * (1 + (lane & 15u)))) >> (64 - 2 * m)), y2 = loc_revcomp_m(x2, m);
            u32 e1 = lane < nt + W - 1 ? loc_entry(loc_order(x1 < y1 ? x1 : y1), lane) : kLocInf;
            u32 e2 = (lane < W - 1 && 32 + lane < nt + W - 1) ? loc_entry(loc_order(x2 < y2 ? x2 : y2), 32 + lane) : kLocInf;
            // sliding-window minimum across the lanes: after the round for d, e1 / e2 cover d + d places from their base on
            for (u32 d = 1; d < W; d <<= 1) {
                const u32 src = (lane + d) & 31u;
                const u32 a1 = __shfl_sync(FULL, e1, src), a2 = __shfl_sync(FULL, e2, src);
                const u32 n1 = lane + d < 32u ? a1 : a2;  // base lane + d
                const u32 n2 = lane + d < 32u ? a2 : kLocInf;  // base 32 + lane + d
                e1 = loc_combine(e1, n1);
                e2 = loc_combine(e2, n2);
            }
            wm = e1;
        } else {
            wm = kLocTie;
        }
        if (have) {
            u32 p = (wm & 63u) - lane;
            p = p < W ? p : 0u;  // (meaningless when the exact definition decides below)
            const u32 xf = (u32)(kf >> (g.fbits - 2 * p)) & mask, yr = (u32)(kr >> (2 * p)) & mask;
            if ((wm & kLocTie) || xf == yr) {  // the exact definition decides
                loc_pick(kf, kr, g, x, pos, sw, o);
            } else {
                const bool fwd = xf < yr;
                pos = fwd ? p : W - 1 - p;
                o = fwd ? kf : kr;
                sw = !fwd;
                x = fwd ? xf : yr;
            }
            loc_key(x, pos, o, g, bucket, R);
            R4 = R << 4;
        }
        return true;
    };

    // tile A: its DIR entry consumed, WIDE window chosen; tile B: prepared, DIR load issued
    u64 resA = 0, R4A = 0, resB = 0, R4B = 0, deB = 0;
    u32 loA = 0, hiA = 0, r0A = 0, bucketB = 0, posB = 0;
    bool okA = false, actA = false, swA = false, okB = false, haveB = false, swB = false;
    okB = prepare(resB, haveB, swB, bucketB, posB, R4B);
    if (okB && haveB) deB = ld_loc_dir(lv.dir + bucketB);
    if (okB) count_requests(haveB, lv.dir + bucketB);

    while (okA || okB) {
        // ---------------------------------------------------------------- WIDE of tile A goes out (DIR of tile B is in flight)
        u64 v[kLocRows] = {};
        if (okA && actA) loc_probe8(lv.rows, r0A, v);
        if (okA) count_requests(actA, lv.rows + r0A);
        // ---------------------------------------------------------------- meanwhile: tile C
        u64 resC = 0, R4C = 0;
        u32 bucketC = 0, posC = 0;
        bool haveC = false, swC = false, okC = false;
        if (!exhausted) okC = prepare(resC, haveC, swC, bucketC, posC, R4C);
        // ---------------------------------------------------------------- tile A: results, or park the search
        if (okA) {
            u32 st = 0;
            bool park = false;
            if (actA) {
                if (loc_consume8(v, r0A, R4A, loA, hiA, st)) out[resA + lane] = loc_result<MODE, STRANDS>(st, swA);
                else park = true;
            }
            const unsigned pmask = __ballot_sync(FULL, park);
            if (pmask) {
                if (qcount + (u32)__popc(pmask) > kLocTailCap) drain();  // leaves at most 32 parked
                if (park) {
                    const u32 e = qcount + (u32)__popc(pmask & lt_mask);
                    tq.q[e] = (resA + lane) | ((u64)swA << 63);
                    tq.R4[e] = R4A;
                    tq.lo[e] = loA;
                    tq.hi[e] = hiA;
                }
                qcount += (u32)__popc(pmask);
                __syncwarp();
                if (qcount >= 32u) drain();
            }
        }
        // ---------------------------------------------------------------- tile B: DIR entry -> becomes tile A
        okA = okB;
        if (okB) {
            resA = resB;
            R4A = R4B;
            swA = swB;
            actA = false;
            if (haveB) {
                const u32 first = (u32)deB, cnt = (u32)(deB >> 32) & ((1u << kLocCountBits) - 1u);
                const u32 pmin = (u32)(deB >> 54) & 31u, pspan = (u32)(deB >> 59);
                if (cnt == 0 || posB < pmin || posB > pmin + pspan) {
                    out[resB + lane] = loc_result<MODE, STRANDS>(0u, swB);  // empty bucket, or no row of it has the minimizer at this position
                } else {
                    actA = true;
                    loA = first;
                    hiA = first + cnt;
                    // rows are sorted by pos first: look where pos interpolates to
                    r0A = loc_window(first + ((2 * (posB - pmin) + 1) * cnt) / (2 * (pspan + 1)), loA);
                }
            }
        }
        // ---------------------------------------------------------------- tile C becomes tile B: its DIR load goes out
        okB = okC;
        resB = resC;
        R4B = R4C;
        swB = swC;
        haveB = haveC;
        bucketB = bucketC;
        posB = posC;
        deB = 0;
        if (okB && haveB) deB = ld_loc_dir(lv.dir + bucketB);
        if (okB) count_requests(haveB, lv.dir + bucketB);
    }
    while (qcount) drain();
    if (probe_ctr) count_probes(probe_ctr, nreq);
}

inline size_t loc_stream_smem(const LocGeom &, int block = kLocBlock) { return (size_t)(block / 32) * sizeof(LocTailQ); }

}  // namespace fmsi
