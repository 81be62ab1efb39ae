// query_kernels.cuh — the backward-search kernels.
//
// query_kmers_kernel: single-k-mer queries (reference query_kmers_single, src/fms_index.h:263-331,
// minus the strand predictor, which the host replays when asked for STRANDS_BOTH).
//
// Execution model. The grid is persistent (SMs x resident CTAs). Each warp grabs chunks of
// `chunk` consecutive k-mers from a global atomic cursor; inside a chunk the 32 lanes are
// independent state machines: whenever a lane's search ends (interval empty, or k steps done and
// the mask probed) it is REFILLED in the same iteration from the warp's cursor — a ballot + popc
// prefix hands out the next indices, the k-mers themselves sit in two register tiles (the current
// and the next 32, loaded coalesced one tile ahead) and are fetched with a shuffle. So every lane
// always has exactly one dependent sector load in flight, no lane idles while its neighbours
// finish their longer searches, and results are written straight to their query index.
//
// A lane is in one of three phases, each costing one memory round trip:
//   TABLE : read {i, j} of the k-mer's last t bases from the suffix table (replaces t LF-steps)
//   STEP  : m LF-steps from one multi-step sector (multistep.cuh; m = 2 or 3 bases per probe) while at least m
//           steps remain, else one LF-step = rank sector of i; in either case a second sector when j lies in
//           another block
//   MASK  : probe the aux sector of the final interval (mask bit / mask rank)
// All loads of an iteration are issued before any is consumed, so the lanes of a warp overlap
// their misses even when they are in different phases.
#pragma once
#include "device_index.cuh"

namespace fmsi {

enum { PH_TABLE = 0, PH_STEP = 1, PH_MASK = 2 };
enum { K_MODE_OR = 0, K_MODE_ALL = 1, K_MODE_GENERAL = 2 };
enum { K_F_OR = 0, K_F_AND = 1, K_F_XOR = 2, K_F_RANGE = 3 };

// Demasking function of the f-MS framework (reference src/functions.h:7-57), applied to the number
// of ON occurrences and of all occurrences of a k-mer over both strands.
struct GenF {
    int kind;
    unsigned r, s;
};
__device__ __forceinline__ bool apply_f(const GenF &f, u64 ones, u64 total) {
    switch (f.kind) {
    case K_F_AND: return total && ones == total;       // f_and
    case K_F_XOR: return ones & 1ull;                  // f_xor
    case K_F_RANGE: return ones <= f.s && ones >= f.r; // f_r_to_s
    default: return ones != 0;                         // f_or
    }
}
enum { K_OUT_PRESENCE = 0, K_OUT_ORDERS = 1 };
enum { K_STRANDS_LAZY = 0, K_STRANDS_BOTH = 1 };

constexpr int kQueryBlock = 256;

// Probe accounting (fmsi_gpu_count_probes): the lanes' request counts, one atomic per warp at kernel exit.
__device__ __forceinline__ void count_probes(unsigned long long *probe_ctr, u32 nprobe) {
    if (!probe_ctr) return;
    u32 v = nprobe;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31u) == 0) atomicAdd(probe_ctr, (unsigned long long)v);
}

template <bool WIDE> struct TableEntry { u32 i, j; };
template <> struct TableEntry<true> { u64 i, j; };

template <bool WIDE>
__device__ __forceinline__ void ld_table(const DevIndex &d, u64 slot, typename PosT<WIDE>::type &i,
                                         typename PosT<WIDE>::type &j) {
    const char *p = reinterpret_cast<const char *>(d.table) + (slot << d.tshift);
    if (WIDE) {
        const ulonglong2 e = ld_pair64(p);
        i = (typename PosT<WIDE>::type)e.x;
        j = (typename PosT<WIDE>::type)e.y;
    } else {
        const uint2 e = ld_pair32(p);
        i = (typename PosT<WIDE>::type)e.x;
        j = (typename PosT<WIDE>::type)e.y;
    }
}

// Result of a finished strand search from its aux sector(s).
//   presence: -1 empty (never reaches here), 0 / 1;  orders: id or -1.
template <int MODE, int OUT>
__device__ __forceinline__ long long strand_result(u64 i, u64 j, u64 mask_i, u64 cum_i, u64 mask_j, u64 cum_j) {
    if (OUT == K_OUT_PRESENCE && MODE == K_MODE_ALL) {
        // infer_presence<true>: mask[sa_start] (fms_index.h:129-131)
        return (long long)((mask_i >> (i & 63)) & 1ull);
    }
    // infer_presence<false>: any ON occurrence in [i, j) (fms_index.h:133-137) == rank1(j) > rank1(i)
    const u64 ri = mask_rank_excl(mask_i, cum_i, (u32)i & 63u);
    const u64 rj = mask_rank_incl(mask_j, cum_j, (u32)(j - 1) & 63u);
    const bool any = rj > ri;
    if (OUT == K_OUT_PRESENCE) return any ? 1 : 0;
    // kmer_order_if_present: mask_rank(sa_start) if any else -1 (fms_index.h:146-156)
    return any ? (long long)ri : -1ll;
}

// INDIRECT: the launch answers only the queries listed in sel[0 .. *n_dev) (the dictionary kernel's
// overflow list, dict.cuh): query q of the launch is kmers[sel[q]] and its result goes to slot sel[q].
template <int MODE, int OUT, int STRANDS, bool WIDE, bool INDIRECT = false>
__global__ void __launch_bounds__(kQueryBlock)
query_kmers_kernel(const DevIndex d, const u64 *__restrict__ kmers, const u64 n_arg, void *__restrict__ out,
                   unsigned long long *__restrict__ cursor, const u32 chunk, const u32 *__restrict__ sel = nullptr,
                   const unsigned long long *__restrict__ n_dev = nullptr, const GenF gf = GenF{0, 0, 0},
                   unsigned long long *__restrict__ probe_ctr = nullptr) {
    typedef typename PosT<WIDE>::type pos_t;
    const u64 n = INDIRECT ? (u64)*n_dev : n_arg;
    const unsigned FULL = 0xffffffffu;
    const u32 lane = threadIdx.x & 31u;
    const u32 lt_mask = (1u << lane) - 1u;
    const u32 t = d.t, k = d.k;
    const u64 tmask = t ? ((t >= 32) ? ~0ull : ((1ull << (2 * t)) - 1ull)) : 0ull;
    const bool need_j = !(OUT == K_OUT_PRESENCE && MODE == K_MODE_ALL);
    const u32 mm = d.multi_m;
    const u64 xmask = (1ull << (2 * mm)) - 1ull;

    // lane state
    bool active = false;
    u32 phase = PH_TABLE, strand = 0, steps = 0;
    u64 kf = 0, pat = 0, idx = 0;
    pos_t i = 0, j = 0;
    long long res_f = 0;
    u64 g_ones = 0, g_total = 0;  // K_MODE_GENERAL: occurrences summed over the strands
    // warp state (uniform)
    u64 cend = 0, wnext = 0, tile_base = 0, bufA = 0, bufB = 0;
    u32 selA = 0, selB = 0;  // INDIRECT: result slots of the register tiles
    bool exhausted = false;
    u32 nprobe = 0;  // dependent memory requests issued by this lane (reported when probe_ctr is given)
    auto fetch = [&](u64 q, u64 &km, u32 &sl) {
        km = 0;
        sl = 0;
        if (q < cend) {
            if (INDIRECT) {
                sl = sel[q];
                km = kmers[sl];
            } else {
                km = kmers[q];
            }
        }
    };

    for (;;) {
        // ---------------------------------------------------------------- refill idle lanes
        const unsigned need = __ballot_sync(FULL, !active);
        if (need && !exhausted) {
            if (wnext >= cend) {  // warp-uniform: take the next chunk
                unsigned long long c0 = 0;
                if (lane == 0) c0 = atomicAdd(cursor, (unsigned long long)chunk);
                c0 = __shfl_sync(FULL, c0, 0);
                if (c0 >= n) {
                    exhausted = true;
                } else {
                    wnext = tile_base = c0;
                    cend = (c0 + chunk < n) ? c0 + chunk : n;
                    fetch(tile_base + lane, bufA, selA);
                    fetch(tile_base + 32 + lane, bufB, selB);
                }
            }
            if (!exhausted) {
                const u32 pre = __popc(need & lt_mask);
                const u64 my = wnext + pre;
                const bool take = !active && my < cend;
                const u32 src = (u32)(my - tile_base);  // < 64
                u64 km = __shfl_sync(FULL, bufA, src & 31u);
                u32 sl = INDIRECT ? __shfl_sync(FULL, selA, src & 31u) : 0u;
                if (__any_sync(FULL, take && src >= 32u)) {
                    const u64 kb = __shfl_sync(FULL, bufB, src & 31u);
                    const u32 sb = INDIRECT ? __shfl_sync(FULL, selB, src & 31u) : 0u;
                    if (src >= 32u) {
                        km = kb;
                        sl = sb;
                    }
                }
                const u64 left = cend - wnext;
                const u32 want = __popc(need);
                wnext += (want < left) ? want : left;
                if (wnext - tile_base >= 32) {  // uniform: slide the register tiles
                    tile_base += 32;
                    bufA = bufB;
                    selA = selB;
                    fetch(tile_base + 32 + lane, bufB, selB);
                }
                if (take) {
                    active = true;
                    idx = INDIRECT ? (u64)sl : my;
                    g_ones = g_total = 0;
                    kf = km;
                    pat = km;
                    strand = 0;
                    if (t) {
                        phase = PH_TABLE;
                    } else {
                        phase = PH_STEP;
                        i = 0;
                        j = (pos_t)d.n;
                        steps = k;
                    }
                }
            }
        }
        if (!__any_sync(FULL, active)) {
            if (exhausted) break;
            continue;
        }

        // ---------------------------------------------------------------- issue this round's loads
        const bool isT = active && phase == PH_TABLE;
        const bool isS = active && phase == PH_STEP;
        const bool isM = active && phase == PH_MASK;
        // multi-step probe: the next mm bases in one sector of the array of that mm-mer (multistep.cuh)
        const bool isX = isS && mm && steps >= mm;
        u64 bi, bj;
        const void *pa, *pb;
        if (isX) {
            bi = multi_block_of<WIDE>(i);
            bj = multi_block_of<WIDE>(j);
            const MultiBlock *base = d.multi + (pat & xmask) * (u64)d.multi_nblk;
            pa = base + bi;
            pb = base + bj;
        } else {
            bi = (u64)i >> 6;
            bj = isM ? (((u64)j - 1) >> 6) : ((u64)j >> 6);
            pa = isS ? (const void *)(d.rank + bi) : (const void *)(d.aux + bi);
            pb = isS ? (const void *)(d.rank + bj) : (const void *)(d.aux + bj);
        }
        const bool two = (isS || (isM && need_j)) && (bj != bi);
        u64 a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
        pos_t ti = 0, tj = 0;
        if (isT) ld_table<WIDE>(d, pat & tmask, ti, tj);
        if (isS || isM) {
            ld_sector(pa, a0, a1, a2, a3);
            if (two) ld_sector(pb, b0, b1, b2, b3);
        }
        nprobe += (u32)isT + (u32)(isS || isM) + (u32)two;

        // ---------------------------------------------------------------- consume
        bool done = false;       // this strand's search ended
        long long res = -1;      // its value
        if (isT) {
            i = ti;
            j = tj;
            pat >>= 2 * t;  // t < 32 whenever steps remain; harmless otherwise
            steps = k - t;
            if (i == j) done = true;
            else phase = steps ? PH_STEP : PH_MASK;
        } else if (isS) {
            if (!two) {
                b0 = a0; b1 = a1; b2 = a2; b3 = a3;
            }
            if (isX) {
                const u32 oi = (u32)((u64)i - bi * MultiGeom<WIDE>::rows), oj = (u32)((u64)j - bj * MultiGeom<WIDE>::rows);
                i = lf_multi_t<WIDE>(a0, a1, a2, a3, oi);
                j = lf_multi_t<WIDE>(b0, b1, b2, b3, oj);
                pat >>= 2 * mm;
                steps -= mm;
            } else {
                const u32 c = (u32)pat & 3u;
                pat >>= 2;
                const pos_t ni = lf_map<WIDE>(d, a0, a1, a2, a3, i, c);
                const pos_t nj = lf_map<WIDE>(d, b0, b1, b2, b3, j, c);
                i = ni;
                j = nj;
                --steps;
            }
            if (i == j) done = true;
            else if (steps == 0) phase = PH_MASK;
        } else if (isM) {
            if (!two) {
                b1 = a1; b2 = a2;
            }
            if (MODE == K_MODE_GENERAL) {  // single_query_general, fms_index.h:171-179
                g_ones += mask_rank_incl(b1, b2, (u32)((u64)j - 1) & 63u) - mask_rank_excl(a1, a2, (u32)i & 63u);
                g_total += (u64)j - (u64)i;
            } else {
                res = strand_result<MODE, OUT>((u64)i, (u64)j, a1, a2, b1, b2);
            }
            done = true;
        }

        if (done) {
            bool other;  // run the other strand next?
            if (MODE == K_MODE_GENERAL) {
                // both strands, except that a self-complementary k-mer is counted once (:318-323)
                other = strand == 0 && revcomp_packed(kf, k) != kf;
                if (!other) res = apply_f(gf, g_ones, g_total) ? 1 : 0;
            } else if (STRANDS == K_STRANDS_BOTH) {
                other = strand == 0;
                if (other) res_f = res;
            } else if (OUT == K_OUT_ORDERS) {
                other = strand == 0 && res < 0;   // fms_index.h:283-288
            } else if (MODE == K_MODE_OR) {
                other = strand == 0 && res != 1;  // :289-293
            } else {
                other = strand == 0 && res == -1; // :294-298
            }
            if (other) {
                strand = 1;
                pat = revcomp_packed(kf, k);
                if (t) {
                    phase = PH_TABLE;
                } else {
                    phase = PH_STEP;
                    i = 0;
                    j = (pos_t)d.n;
                    steps = k;
                }
            } else {
                if (OUT == K_OUT_PRESENCE) {
                    unsigned char v;
                    if (STRANDS == K_STRANDS_BOTH) v = (unsigned char)((res_f + 1) | ((res + 1) << 2));
                    else v = (unsigned char)(res == 1);
                    reinterpret_cast<unsigned char *>(out)[idx] = v;
                } else {
                    if (STRANDS == K_STRANDS_BOTH) {
                        reinterpret_cast<longlong2 *>(out)[idx] = make_longlong2(res_f, res);
                    } else {
                        reinterpret_cast<long long *>(out)[idx] = res;
                    }
                }
                active = false;
            }
        }
    }
    count_probes(probe_ctr, nprobe);
}

// ------------------------------------------------------------------------------------------------
// Suffix table, built level by level: level s holds the interval of every s-mer x (packed, first
// base highest) = update_range(level s-1 entry of x without its first base, first base).
template <bool WIDE>
__global__ void table_level_kernel(const DevIndex d, const TableEntry<WIDE> *__restrict__ prev,
                                   TableEntry<WIDE> *__restrict__ cur, const u32 s) {
    typedef typename PosT<WIDE>::type pos_t;
    const u64 x = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    const u64 total = 1ull << (2 * s);
    if (x >= total) return;
    const u32 c = (u32)(x >> (2 * (s - 1)));
    const u64 rest = x & ((1ull << (2 * (s - 1))) - 1ull);
    pos_t i, j;
    if (s == 1) {
        i = 0;
        j = (pos_t)d.n;
    } else {
        i = prev[rest].i;
        j = prev[rest].j;
    }
    if (i != j) {
        u64 a0, a1, a2, a3, b0, b1, b2, b3;
        ld_sector_l1(d.rank + ((u64)i >> 6), a0, a1, a2, a3);
        ld_sector_l1(d.rank + ((u64)j >> 6), b0, b1, b2, b3);
        const pos_t ni = lf_map<WIDE>(d, a0, a1, a2, a3, i, c);
        const pos_t nj = lf_map<WIDE>(d, b0, b1, b2, b3, j, c);
        i = ni;
        j = nj;
    }
    cur[x].i = i;
    cur[x].j = j;
}

// ------------------------------------------------------------------------------------------------
// Building-block probes (one thread per element), mirroring the reference's functions so that the
// reference's unit goldens can be run against the device code function by function.
template <bool WIDE>
__device__ __forceinline__ u64 dev_lf(const DevIndex &d, u64 i, u32 c) {
    typedef typename PosT<WIDE>::type pos_t;
    u64 a0, a1, a2, a3;
    ld_sector_l1(d.rank + (i >> 6), a0, a1, a2, a3);
    return (u64)lf_map<WIDE>(d, a0, a1, a2, a3, (pos_t)i, c);
}

template <bool WIDE>
__global__ void probe_rank_kernel(const DevIndex d, const u64 *i, const unsigned char *c, u64 n, const u64 *counts, u64 *out) {
    const u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (q >= n) return;
    // rank(i, c) = LF(i, c) - counts[c]
    out[q] = dev_lf<WIDE>(d, i[q], c[q] & 3u) - counts[c[q] & 3u];
}

template <bool WIDE>
__global__ void probe_update_range_kernel(const DevIndex d, u64 *i, u64 *j, const unsigned char *c, u64 n) {
    const u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (q >= n) return;
    if (i[q] == j[q]) return;  // fms_index.h:99
    const u64 ni = dev_lf<WIDE>(d, i[q], c[q] & 3u);
    const u64 nj = dev_lf<WIDE>(d, j[q], c[q] & 3u);
    i[q] = ni;
    j[q] = nj;
}

// extend_range_with_klcp (fms_index.h:106-109): j -> 1 + first zero of klcp at or after j-1,
// i -> 1 + last zero of klcp at or before i-1.
__device__ __forceinline__ void dev_extend_klcp(const DevIndex &d, u64 &i, u64 &j) {
    {
        u64 p = j - 1;
        for (;;) {
            const u64 w = d.aux[p >> 6].klcp;
            const u64 z = ~w & ~low_mask((u32)p & 63u);  // zeros at positions >= p in this word
            if (z) {
                j = (p & ~63ull) + (u64)(__ffsll((long long)z) - 1) + 1;
                break;
            }
            p = (p & ~63ull) + 64;
        }
    }
    {
        u64 p = i - 1;
        for (;;) {
            const u64 w = d.aux[p >> 6].klcp;
            const u64 z = ~w & ((2ull << (p & 63)) - 1ull);  // zeros at positions <= p in this word
            if (z) {
                i = (p & ~63ull) + (u64)(63 - __clzll((long long)z)) + 1;
                break;
            }
            p = (p & ~63ull) - 1;
        }
    }
}

__global__ void probe_extend_kernel(const DevIndex d, u64 *i, u64 *j, u64 n) {
    const u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (q >= n) return;
    u64 a = i[q], b = j[q];
    dev_extend_klcp(d, a, b);
    i[q] = a;
    j[q] = b;
}

template <bool WIDE>
__global__ void probe_get_range_kernel(const DevIndex d, const u64 *kmers, u32 k, u64 n, int use_table, u64 *oi, u64 *oj) {
    typedef typename PosT<WIDE>::type pos_t;
    const u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (q >= n) return;
    u64 pat = kmers[q];
    u64 i = 0, j = d.n;
    u32 steps = k;
    if (use_table && d.t && d.t <= k) {
        pos_t ti, tj;
        ld_table<WIDE>(d, pat & ((d.t >= 32) ? ~0ull : ((1ull << (2 * d.t)) - 1ull)), ti, tj);
        i = ti;
        j = tj;
        pat >>= 2 * d.t;
        steps = k - d.t;
    }
    for (; steps > 0 && i != j; --steps) {
        const u32 c = (u32)pat & 3u;
        pat >>= 2;
        const u64 ni = dev_lf<WIDE>(d, i, c), nj = dev_lf<WIDE>(d, j, c);
        i = ni;
        j = nj;
    }
    oi[q] = i;
    oj[q] = j;
}

__global__ void probe_presence_kernel(const DevIndex d, const u64 *si, const u64 *sj, u64 n, int max_ones, signed char *out) {
    const u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (q >= n) return;
    const u64 i = si[q], j = sj[q];
    if (i == j) {
        out[q] = -1;
        return;
    }
    const AuxBlock ai = d.aux[i >> 6], aj = d.aux[(j - 1) >> 6];
    long long r;
    if (max_ones) r = strand_result<K_MODE_ALL, K_OUT_PRESENCE>(i, j, ai.mask, ai.mask_cum, aj.mask, aj.mask_cum);
    else r = strand_result<K_MODE_OR, K_OUT_PRESENCE>(i, j, ai.mask, ai.mask_cum, aj.mask, aj.mask_cum);
    out[q] = (signed char)r;
}

__global__ void probe_order_kernel(const DevIndex d, const u64 *si, const u64 *sj, u64 n, long long *out) {
    const u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (q >= n) return;
    const u64 i = si[q], j = sj[q];
    if (i == j) {
        out[q] = -1;
        return;
    }
    const AuxBlock ai = d.aux[i >> 6], aj = d.aux[(j - 1) >> 6];
    out[q] = strand_result<K_MODE_OR, K_OUT_ORDERS>(i, j, ai.mask, ai.mask_cum, aj.mask, aj.mask_cum);
}

}  // namespace fmsi
