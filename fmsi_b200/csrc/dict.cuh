// dict.cuh — the k-mer dictionary tier: single-k-mer queries in ~1 memory request per strand.
//
// Why. On B200 a dependent random read costs one L2-miss REQUEST (~37 G/s per GPU, measured with
// tools/randbw2.cu) whether it returns 8 or 32 bytes, and DRAM moves a 128-byte line for it either
// way. The backward search of the reference (get_range_with_pattern, src/fms_index.h:117-124) needs
// k - t dependent requests after the suffix table; this tier answers the same question — the SA
// interval of the k-mer, then infer_presence / kmer_order_if_present on it (:126-156) — from data
// derived once at load time from the very same index (BWT + SA-transformed mask):
//
//   bucket[x], x = the k-mer's FIRST t bases (32 B = one sector, 4^t of them, replaces the {i, j}
//   suffix table; the backward kernels read its first 8 bytes):
//       u32 i, j        SA interval of the t-mer x
//       u32 meta        bit s: row i+s is valid (its suffix has >= k characters); bit 8+s: mask[i+s]
//       PAY32: u32 pay[5]   the next B = k - t <= 16 bases of the suffixes of rows i .. i+4
//       PAY64: u32 pad; u64 pay[2]   (16 < B <= 31) rows i, i+1
//   rows[r], r in [0, N) (8 B, SA order): pay << 2 | valid << 1 | mask[r] — the same for every row,
//   read only when a bucket holds more rows than fit in its sector.
//
// Rows of one bucket are suffixes sharing x, in lexicographic order, so the rows whose next B bases
// equal the k-mer's last B bases are exactly the k-mer's SA interval [i', j'):  -O presence =
// mask[i'] (first match), or-presence = any mask bit among the matches, lookup id = rank1(i') when
// any match is ON (one aux-sector probe). Payloads are non-decreasing inside a bucket (rows whose
// suffix ends before k characters are padded with A after the sentinel, which keeps them in order, and
// are flagged invalid), so a bucket with many rows is binary-searched for the first payload >= the
// k-mer's, one sector per step, before the linear scan. Only a k-mer whose matches run on for more than
// kDictMaxScan rows without settling the answer (a long run of OFF occurrences in or/lookup mode) is
// put on an overflow list that the backward-search kernel answers in a second launch
// (query_kmers_kernel<..., INDIRECT>), so results are exact for every input.
//
// The rows are derived from the BWT alone (works for file-loaded and device-built indexes alike):
// psi = inverse of the LF-mapping (one scatter pass), then every row walks psi k-1 times reading
// the first character of the visited rows (F column = a compare against counts[]).
#pragma once
#include "query_kernels.cuh"

namespace fmsi {

constexpr u32 kDictCap32 = 5, kDictCap64 = 2;
constexpr u32 kDictMaxScan = 64;  // rows scanned linearly from the first candidate before giving up to the fixup pass
constexpr u32 kDictBsearchMin = 16;  // windows larger than this are narrowed by binary search first

struct DictView {
    const u64 *rows;   // [N]
    u32 B;             // payload bases = k - t
    u32 k;             // the k the rows were built for
    u32 enabled;
};

// ------------------------------------------------------------------------------------------- build
__device__ __forceinline__ u32 block_symbol(u64 lo, u64 hi, u32 off) {
    return (u32)((lo >> off) & 1ull) | ((u32)((hi >> off) & 1ull) << 1);
}

// psi[LF(r)] = r for every row r (LF is a permutation of [0, N); the '$' row maps to row 0).
__global__ void psi_scatter_kernel(const DevIndex d, u32 *__restrict__ psi) {
    const u64 r = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (r >= d.n) return;
    u64 a0, a1, a2, a3;
    ld_sector_l1(d.rank + (r >> 6), a0, a1, a2, a3);
    u32 target = 0;
    if (r != d.dollar) {
        const u32 c = block_symbol(a2, a3, (u32)r & 63u);
        target = lf_map<false>(d, a0, a1, a2, a3, (u32)r, c);
    }
    psi[target] = (u32)r;
}

// rows[r] = (bases t .. k-1 of the suffix of row r) << 2 | valid | mask[r]. F column from counts.
__global__ void rows_walk_kernel(const DevIndex d, const u32 *__restrict__ psi, const u32 c1, const u32 c2, const u32 c3,
                                 const u32 t, const u32 B, u64 *__restrict__ rows) {
    const u64 r = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (r >= d.n) return;
    u32 cur = (u32)r;
    for (u32 s = 0; s < t; ++s) cur = __ldg(psi + cur);
    u64 pay = 0;
    bool ended = false;  // reached the sentinel: the rest is padding (A = 0), which keeps the bucket sorted
    for (u32 s = 0; s < B; ++s) {
        if (cur == 0) ended = true;
        const u32 c = ended ? 0u : (cur >= c3) ? 3u : (cur >= c2) ? 2u : (cur >= c1) ? 1u : 0u;
        pay = (pay << 2) | c;
        if (!ended && s + 1 < B) cur = __ldg(psi + cur);
    }
    const u64 m = (d.aux[r >> 6].mask >> (r & 63)) & 1ull;
    rows[r] = (pay << 2) | 2ull | m;
}

// The k rows whose suffixes are shorter than k (text positions n, n-1, ..., n-k+1): walk LF from row 0.
__global__ void rows_invalidate_kernel(const DevIndex d, const u32 k, u64 *__restrict__ rows) {
    if (blockIdx.x || threadIdx.x) return;
    u64 cur = 0;
    for (u32 s = 0; s < k; ++s) {
        rows[cur] &= ~2ull;
        if (cur == d.dollar) break;  // that was the whole text
        u64 a0, a1, a2, a3;
        ld_sector_l1(d.rank + (cur >> 6), a0, a1, a2, a3);
        const u32 c = block_symbol(a2, a3, (u32)cur & 63u);
        cur = lf_map<false>(d, a0, a1, a2, a3, (u32)cur, c);
    }
}

struct alignas(32) Bucket {
    u32 i, j, meta, w[5];
};
static_assert(sizeof(Bucket) == 32, "one sector");

template <bool PAY64>
__global__ void bucket_fill_kernel(const TableEntry<false> *__restrict__ tab, const u64 *__restrict__ rows, const u64 total,
                                   Bucket *__restrict__ buckets) {
    const u64 x = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (x >= total) return;
    const u32 i = tab[x].i, j = tab[x].j;
    Bucket b;
    b.i = i;
    b.j = j;
    b.meta = 0;
    for (int s = 0; s < 5; ++s) b.w[s] = 0;
    const u32 cap = PAY64 ? kDictCap64 : kDictCap32;
    const u32 m = (j - i < cap) ? (j - i) : cap;
    for (u32 s = 0; s < m; ++s) {
        const u64 row = rows[(u64)i + s];
        b.meta |= (u32)((row >> 1) & 1ull) << s;
        b.meta |= (u32)(row & 1ull) << (8 + s);
        const u64 pay = row >> 2;
        if (PAY64) {
            b.w[1 + 2 * s] = (u32)pay;
            b.w[2 + 2 * s] = (u32)(pay >> 32);
        } else {
            b.w[s] = (u32)pay;
        }
    }
    buckets[x] = b;
}

// ------------------------------------------------------------------------------------------- query
enum { DP_BUCKET = 0, DP_ROWS = 1, DP_AUX = 2, DP_BSEARCH = 3 };

template <int MODE, int OUT, int STRANDS, bool PAY64>
__global__ void __launch_bounds__(kQueryBlock)
dict_query_kernel(const DevIndex d, const DictView dv, const u64 *__restrict__ kmers, const u64 n, void *__restrict__ out,
                  unsigned long long *__restrict__ cursor, const u32 chunk, u32 *__restrict__ ovf_list,
                  unsigned long long *__restrict__ ovf_count) {
    const unsigned FULL = 0xffffffffu;
    const u32 lane = threadIdx.x & 31u;
    const u32 lt_mask = (1u << lane) - 1u;
    const u32 k = d.k, B = dv.B;
    const u64 pmask = B ? ((B >= 32) ? ~0ull : ((1ull << (2 * B)) - 1ull)) : 0ull;
    const u32 CAP = PAY64 ? kDictCap64 : kDictCap32;
    // What ends a strand search early: -O wants the first match only; or-presence any ON match.
    const bool first_only = (OUT == K_OUT_PRESENCE && MODE == K_MODE_ALL);

    bool active = false;
    u32 phase = DP_BUCKET, strand = 0;
    u64 kf = 0, pat = 0, idx = 0, q = 0;
    u32 i = 0, j = 0, rpos = 0, hi = 0;  // hi: binary-search window end, then the linear scan's give-up row
    u32 fm = 0;            // first matching row (valid when have)
    bool have = false, fm_mask = false, anymask = false;
    long long res_f = 0;
    u64 cend = 0, wnext = 0, tile_base = 0, bufA = 0, bufB = 0;
    bool exhausted = false;

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
                    kf = km;
                    pat = km;
                    strand = 0;
                    phase = DP_BUCKET;
                }
            }
        }
        if (!__any_sync(FULL, active)) {
            if (exhausted) break;
            continue;
        }

        // ---------------------------------------------------------------- issue this round's loads
        const bool isB = active && phase == DP_BUCKET;
        const bool isR = active && phase == DP_ROWS;
        const bool isA = active && phase == DP_AUX;
        const bool isS = active && phase == DP_BSEARCH;
        const u32 msec = (rpos + ((hi - rpos) >> 1)) >> 2;  // sector probed by a binary-search step
        u64 a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        const u32 rsec = rpos >> 2;                         // rows sector (4 rows of 8 B)
        if (isB) ld_sector(reinterpret_cast<const char *>(d.table) + ((pat >> (2 * B)) << 5), a0, a1, a2, a3);
        if (isR) {
            ld_sector(dv.rows + ((u64)rsec << 2), a0, a1, a2, a3);
        }
        if (isA) ld_sector(d.aux + (fm >> 6), a0, a1, a2, a3);
        if (isS) ld_sector(dv.rows + ((u64)msec << 2), a0, a1, a2, a3);

        // ---------------------------------------------------------------- consume
        bool done = false, overflow = false;
        long long res = -1;
        if (isB) {
            i = (u32)a0;
            j = (u32)(a0 >> 32);
            q = pat & pmask;
            have = false;
            fm_mask = anymask = false;
            const u32 cnt = j - i;
            const u32 meta = (u32)a1;
            bool past = false;  // a valid row with payload > q was seen: no match can follow
            const u32 m = cnt < CAP ? cnt : CAP;
#pragma unroll
            for (u32 s = 0; s < CAP; ++s) {
                u64 pay;
                if (PAY64) pay = s == 0 ? a2 : a3;
                else pay = s == 0 ? (a1 >> 32) : s == 1 ? (a2 & 0xffffffffull) : s == 2 ? (a2 >> 32) : s == 3 ? (a3 & 0xffffffffull) : (a3 >> 32);
                const bool valid = (meta >> s) & 1u, mb = (meta >> (8 + s)) & 1u;
                if (s < m) {
                    if (pay == q && valid) {
                        if (!have) {
                            have = true;
                            fm = i + s;
                            fm_mask = mb;
                        }
                        anymask |= mb;
                    } else if (pay > q) {
                        past = true;
                    }
                }
            }
            const bool conclusive = cnt <= CAP || past || (first_only ? have : anymask);
            if (conclusive) {
                done = true;
            } else {
                rpos = i + CAP;
                hi = j;
                // matches already seen continue right after the inline rows; otherwise find the first
                // payload >= q, by binary search when the bucket is large
                if (!have && hi - rpos > kDictBsearchMin) {
                    phase = DP_BSEARCH;
                } else {
                    phase = DP_ROWS;
                    hi = rpos + kDictMaxScan;
                }
            }
        } else if (isS) {
            // invariant: rows [i+CAP, rpos) have payload < q; rows >= hi have payload >= q (or hi == j)
            const u32 r0 = msec << 2;
            const u32 lo_r = r0 > rpos ? r0 : rpos, hi_r = (r0 + 4 < hi) ? r0 + 4 : hi;  // in-window rows [lo_r, hi_r)
            const u32 sf = lo_r - r0, sl = hi_r - 1 - r0;
            const u64 first = (sf == 0 ? a0 : sf == 1 ? a1 : sf == 2 ? a2 : a3) >> 2;
            const u64 last = (sl == 0 ? a0 : sl == 1 ? a1 : sl == 2 ? a2 : a3) >> 2;
            if (last < q) rpos = hi_r;
            else if (first >= q) hi = lo_r;
            else {
                rpos = lo_r;  // the first payload >= q lies inside this sector
                hi = lo_r;
            }
            if (hi - rpos <= kDictBsearchMin) {
                phase = DP_ROWS;
                hi = rpos + kDictMaxScan;
            }
        } else if (isR) {
            bool past = false;
#pragma unroll
            for (u32 s = 0; s < 4; ++s) {
                const u64 row = s == 0 ? a0 : s == 1 ? a1 : s == 2 ? a2 : a3;
                const u32 r = (rsec << 2) + s;
                if (r >= rpos && r < j) {
                    const u64 pay = row >> 2;
                    const bool mb = row & 1ull;
                    if (pay == q && (row & 2ull)) {
                        if (!have) {
                            have = true;
                            fm = r;
                            fm_mask = mb;
                        }
                        anymask |= mb;
                    } else if (pay > q) {
                        past = true;
                    }
                }
            }
            rpos = (rsec + 1) << 2;
            if (rpos >= j || past || (first_only ? have : anymask)) done = true;
            else if (rpos >= hi) overflow = true;
        } else if (isA) {
            res = (long long)mask_rank_excl(a1, a2, fm & 63u);
            done = true;
        }

        if (done && !isA && !overflow) {  // strand value from the matches (fms_index.h:126-156)
            if (!have) res = -1;
            else if (OUT == K_OUT_PRESENCE) res = first_only ? (fm_mask ? 1 : 0) : (anymask ? 1 : 0);
            else if (!anymask) res = -1;
            else {
                done = false;
                phase = DP_AUX;
            }
        }

        if (overflow) {
            ovf_list[atomicAdd(ovf_count, 1ull)] = (u32)idx;
            active = false;
        } else if (done) {
            bool other;
            if (STRANDS == K_STRANDS_BOTH) {
                other = strand == 0;
                if (other) res_f = res;
            } else if (OUT == K_OUT_ORDERS) {
                other = strand == 0 && res < 0;
            } else if (MODE == K_MODE_OR) {
                other = strand == 0 && res != 1;
            } else {
                other = strand == 0 && res == -1;
            }
            if (other) {
                strand = 1;
                pat = revcomp_packed(kf, k);
                phase = DP_BUCKET;
            } else {
                if (OUT == K_OUT_PRESENCE) {
                    unsigned char v;
                    if (STRANDS == K_STRANDS_BOTH) v = (unsigned char)((res_f + 1) | ((res + 1) << 2));
                    else v = (unsigned char)(res == 1);
                    reinterpret_cast<unsigned char *>(out)[idx] = v;
                } else {
                    if (STRANDS == K_STRANDS_BOTH) reinterpret_cast<longlong2 *>(out)[idx] = make_longlong2(res_f, res);
                    else reinterpret_cast<long long *>(out)[idx] = res;
                }
                active = false;
            }
        }
    }
}

}  // namespace fmsi
