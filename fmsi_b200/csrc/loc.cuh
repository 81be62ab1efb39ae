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
//   pick(q)   the minimizer of {q, rc(q)}: among the m-mers of both strands the one whose ordering hash (its top 26 bits)
//             is smallest; ties go to the strand with the smaller q * C, then to the leftmost position — a rule that does
//             not depend on which strand the caller holds. o = that strand's k-mer, pos = the m-mer's position in it
//   h'        a bijection of the m-mer on 2m bits (not the ordering hash: a minimum is a small number)
//   bucket    top 2t bits of h'
//   R         pos . h' low 2m - 2t bits . the k - m bases of o around the m-mer        (2k - 2t + pbits bits)
//             (bucket, R) <-> o is one-to-one, so a row match is exact — no fingerprints
//   dir[x]    first row (32 bits) | row count (22) | smallest pos (5) | pos span (5)        8 bytes per bucket
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
constexpr u32 kLocCandBits = 6;  // 2 w <= 34 candidates
// ordering value of candidate `cand` (an m-mer x): top 26 bits of x * C, then the candidate's number
__host__ __device__ __forceinline__ u32 loc_order(u32 x, u32 cand) { return ((x * kLocMul1) & ~((1u << kLocCandBits) - 1u)) | cand; }
// a bijection of [0, 4^m), m <= 16: the key hash of an m-mer
__host__ __device__ __forceinline__ u32 loc_spread(u32 x, u32 m, u32 mask) {
    x = (x * kLocMul3) & mask;
    x ^= x >> m;
    x = (x * kLocMul4) & mask;
    x ^= x >> m;
    return x;
}

// The minimizer of {q, rc}: the strand it sits on (sw: not the caller's q), that strand's k-mer o, its position, its value.
__host__ __device__ __forceinline__ void loc_pick(u64 q, u64 rc, const LocGeom &g, u32 &x, u32 &pos, bool &sw, u64 &o) {
    const bool fs = rc * kFoldMul < q * kFoldMul;
    const u64 a = fs ? rc : q, b = fs ? q : rc;
    const u32 mask = g.m < 16 ? (1u << (2 * g.m)) - 1u : 0xFFFFFFFFu;
    u32 best = 0xFFFFFFFFu;
    for (u32 p = 0; p < g.w; ++p) {
        const u32 va = loc_order((u32)(a >> (g.fbits - 2 * p)) & mask, p);
        const u32 vb = loc_order((u32)(b >> (g.fbits - 2 * p)) & mask, g.w + p);
        best = va < best ? va : best;
        best = vb < best ? vb : best;
    }
    const u32 cand = best & ((1u << kLocCandBits) - 1u);
    const bool bs = cand >= g.w;
    pos = bs ? cand - g.w : cand;
    sw = fs != bs;
    o = bs ? b : a;
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

// directory entry
constexpr u32 kLocCountBits = 22;
__host__ __device__ __forceinline__ u64 loc_dir_entry(u32 first, u32 count, u32 pmin, u32 pspan) {
    return (u64)first | ((u64)count << 32) | ((u64)pmin << 54) | ((u64)pspan << 59);
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
    dir[x] = loc_dir_entry((u32)(g0 + bfirst[x]), cnt, pmin, pmax - pmin);
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
    DevArr<u64> rows(sum + 4);  // rows <= valid runs
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
        if (h_too_many) throw std::runtime_error("a minimizer is shared by more than 2^22 distinct k-mers");
    }
    BCU(cudaMemset(rows.p + G0, 0xff, (sum + 4 - G0) * sizeof(u64)));  // a sector read may run past the last row
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
        DevArr<u64> exact(G0 + 4);
        BCU(cudaMemcpy(exact.p, rows.p, G0 * sizeof(u64), cudaMemcpyDeviceToDevice));
        BCU(cudaMemset(exact.p + G0, 0xff, 4 * sizeof(u64)));
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

// One tile = 32 consecutive queries of the launch, one per lane, walked in lockstep: DIR (the bucket's entry), then SEARCH
// rounds until every lane has its row (or knows there is none). Lanes that hold neighbouring k-mers of a read ask for the
// same directory entry and the same row sectors in the same instruction, which the load unit merges.
struct LocTile {
    u64 R;       // the lane's key inside its bucket
    u32 lo, hi;  // SEARCH: rows still possible; DIR: lo = bucket
    u32 r0;      // SEARCH: first row of the sector to probe
    u32 st;      // the answer: sf | sr << 2 (0 = absent)
    u32 pos;
    bool active, swapped;
};

// Presence outputs only (K_OUT_PRESENCE). Queries come from a k-mer array whose neighbours are neighbouring k-mers of
// a text (extract_kmers_kernel) or straight from reads (ReadSrc), as in fold_query_kernel. Every warp keeps TWO tiles in
// flight (a lane has two independent requests outstanding).
template <int MODE, int STRANDS>
__global__ void __launch_bounds__(kQueryBlock)
loc_query_kernel(const LocView lv, const u64 *__restrict__ kmers, const u64 n, unsigned char *__restrict__ out,
                 unsigned long long *__restrict__ cursor, const u32 chunk, unsigned long long *__restrict__ probe_ctr, const ReadSrc rs) {
    const unsigned FULL = 0xffffffffu;
    const u32 lane = threadIdx.x & 31u;
    const LocGeom g = lv.g;
    const u32 k = g.k;
    const bool from_reads = rs.text != nullptr;

    u64 cnext = 0, cend = 0;  // the warp's grab: queries [cnext, cend) are still to be started (warp-uniform)
    u64 rd_hint = 0;          // from_reads: a read at or before the one of query cnext
    bool exhausted = false;

    LocTile T[2];
    u64 tbase[2] = {0, 0};  // first query of the tile
    u32 phase[2] = {2, 2};  // 0 = DIR, 1 = SEARCH, 2 = empty (warp-uniform)
    // the NEXT tile, prepared while the loads of the current ones are in flight: k-mers fetched, minimizers picked
    LocTile P;
    u64 pbase = 0;
    bool pvalid = false;

    auto count_requests = [&](bool mine, const void *addr) {  // requests after merging: distinct 128-byte lines of a warp instruction
        if (!probe_ctr) return;
        const unsigned act = __ballot_sync(FULL, mine);
        if (mine) {
            const unsigned same = __match_any_sync(act, (unsigned long long)addr >> 7);
            if ((u32)(__ffs(same) - 1) == lane) atomicAdd(probe_ctr, 1ull);
        }
    };
    auto write_result = [&](const LocTile &t, u64 q) {
        u32 sf = t.st & 3u, sr = t.st >> 2;
        if (t.swapped) {
            const u32 x = sf;
            sf = sr;
            sr = x;
        }
        const int vf = fold_presence<MODE>(sf), vr = fold_presence<MODE>(sr);
        unsigned char v;
        if (STRANDS == K_STRANDS_BOTH) v = (unsigned char)((vf + 1) | ((vr + 1) << 2));
        else if (MODE == K_MODE_ALL) v = (unsigned char)((vf != -1 ? vf : vr) == 1);  // fms_index.h:294-298
        else v = (unsigned char)(vf == 1 || vr == 1);                                  // :289-293
        out[q] = v;
    };
    // prepare the next tile: fetch its k-mers, pick the minimizers (P.lo = bucket)
    auto prepare = [&]() {
        if (cnext >= cend) {
            unsigned long long c0 = 0;
            if (lane == 0) c0 = atomicAdd(cursor, (unsigned long long)chunk);
            c0 = __shfl_sync(FULL, c0, 0);
            if (c0 >= n) {
                exhausted = true;
                return;
            }
            cnext = c0;
            cend = (c0 + chunk < n) ? c0 + chunk : n;
            rd_hint = 0;
        }
        const u64 q = cnext + lane;
        const bool have = q < cend;
        u64 km = 0;
        if (have) {
            if (!from_reads) {
                km = kmers[q];
            } else {
                const u64 slot = rs.slot0 + q;
                const u64 r = read_of_slot(rs, slot, rd_hint);
                km = window64(rs.text, __ldg(rs.roff + r) + (slot - __ldg(rs.rbase + r)), k);
                if (lane == 0) rd_hint = r;
            }
        }
        if (from_reads) rd_hint = __shfl_sync(FULL, rd_hint, 0);  // lane 0 holds the tile's first query
        P.active = have;
        P.st = 0;
        if (have) {
            if (k < 32) km &= (1ull << (2 * k)) - 1ull;
            u32 x;
            u64 o;
            loc_pick(km, revcomp_packed(km, k), g, x, P.pos, P.swapped, o);
            loc_key(x, P.pos, o, g, P.lo, P.R);
        }
        pbase = cnext;
        pvalid = true;
        cnext = (cnext + 32 < cend) ? cnext + 32 : cend;
    };

    for (;;) {
#pragma unroll
        for (int s = 0; s < 2; ++s)
            if (phase[s] == 2 && pvalid) {
                T[s] = P;
                tbase[s] = pbase;
                phase[s] = 0;
                pvalid = false;
            }
        if (phase[0] == 2 && phase[1] == 2 && exhausted) break;

        // ---------------------------------------------------------------- issue both tiles' loads
        u64 de[2] = {0, 0};
        u64 a0[2], a1[2], a2[2], a3[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            a0[s] = a1[s] = a2[s] = a3[s] = 0;
            const bool on = phase[s] != 2 && T[s].active;
            if (on && phase[s] == 0) de[s] = ld_loc_dir(lv.dir + T[s].lo);
            if (on && phase[s] == 1) ld_sector_l1(lv.rows + T[s].r0, a0[s], a1[s], a2[s], a3[s]);
            if (phase[s] != 2) count_requests(on, phase[s] == 0 ? (const void *)(lv.dir + T[s].lo) : (const void *)(lv.rows + T[s].r0));
        }
        // ---------------------------------------------------------------- meanwhile: the next tile
        if (!pvalid && !exhausted) prepare();

        // ---------------------------------------------------------------- consume
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (phase[s] == 2) continue;
            LocTile &t = T[s];
            bool done = false;
            if (t.active && phase[s] == 0) {
                const u32 first = (u32)de[s], cnt = (u32)(de[s] >> 32) & ((1u << kLocCountBits) - 1u);
                const u32 pmin = (u32)(de[s] >> 54) & 31u, pspan = (u32)(de[s] >> 59);
                if (cnt == 0 || t.pos < pmin || t.pos > pmin + pspan) {
                    done = true;  // empty bucket, or no row of it has the minimizer at this position
                } else {
                    t.lo = first;
                    t.hi = first + cnt;
                    // rows are sorted by pos first: start where pos interpolates to
                    const u32 guess = first + (u32)(((u64)(2 * (t.pos - pmin) + 1) * cnt) / (2 * (pspan + 1)));
                    t.r0 = (guess < t.hi ? guess : t.hi - 1) & ~3u;
                }
            } else if (t.active) {
                // invariant: rows before lo are smaller than R, rows from hi on are larger
                const u32 r0 = t.r0;
                const u32 w0 = r0 > t.lo ? r0 : t.lo, w1 = (r0 + 4 < t.hi) ? r0 + 4 : t.hi;
                bool found = false;
                u64 first = 0, last = 0;
#pragma unroll
                for (u32 e = 0; e < 4; ++e) {
                    const u64 rw = e == 0 ? a0[s] : e == 1 ? a1[s] : e == 2 ? a2[s] : a3[s];
                    const u32 r = r0 + e;
                    if (r >= w0 && r < w1) {
                        const u64 key = rw >> 4;
                        if (r == w0) first = key;
                        last = key;
                        if (key == t.R) {
                            found = true;
                            t.st = (u32)rw & 15u;
                        }
                    }
                }
                if (found) done = true;
                else if (last < t.R) t.lo = w1;
                else if (first > t.R) t.hi = w0;
                else done = true;  // R falls between two rows of this sector: absent
                if (!done && t.lo >= t.hi) done = true;
                if (!done) t.r0 = (t.lo + ((t.hi - t.lo) >> 1)) & ~3u;
            }
            if (done) {
                write_result(t, tbase[s] + lane);
                t.active = false;
            }
            if (phase[s] == 0) phase[s] = 1;
            if (!__any_sync(FULL, t.active)) phase[s] = 2;
        }
    }
}

}  // namespace fmsi
