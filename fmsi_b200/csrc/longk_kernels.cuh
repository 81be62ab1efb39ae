// longk_kernels.cuh — backward search for k > 32 (a k-mer no longer fits one 64-bit word).
//
// The reference's query path is agnostic to k: get_range_with_pattern (src/fms_index.h:117-124) walks
// the ASCII pattern right to left for any k, and `fmsi index` accepts any k (src/main.cpp:225-233;
// the kLCP array exists for k <= 64). Here a query is a START POSITION into the batch's 2-bit packed
// text (stream_kernels.cuh: pack_bases_kernel); the lane keeps a 32-base window of its pattern in one
// register and reloads it from the packed text (an L1/L2 hit: neighbouring k-mers of a read share
// their words) every 32 LF-steps:
//   forward strand : pattern[t] for t = k-1 .. 0        = text[s + t]
//   reverse strand : rc[t]      for t = k-1 .. 0        = complement(text[s + k-1-t])
// so the forward search consumes the k-mer from its end and the reverse-complement search from its
// start. Everything else — persistent grid, lane refill, TABLE / STEP / MASK phases, strand policy —
// is query_kmers_kernel's (query_kernels.cuh); results have the same layout.
//
// kLCP streaming (query_kmers_streaming, :181-254) is a shortcut, not a different answer: the
// per-strand value of a k-mer does not depend on how its interval was reached, so `-S` requests with
// k > 32 take this kernel too.
#pragma once
#include "query_kernels.cuh"
#include "stream_kernels.cuh"

namespace fmsi {

constexpr u64 kNoKmer = ~0ull;  // start position of a result slot that belongs to no k-mer

// The start position of every result slot's k-mer in the packed text (chunks overlap by k-1 bases, so slot r of
// chunk c starts at chunk_off[c] + (r - res_off[c])); kNoKmer for slots that belong to no k-mer.
__global__ void extract_starts_kernel(const u64 *__restrict__ coff, const u32 *__restrict__ clen,
                                      const u64 *__restrict__ roff, const u64 chunk_begin, const u64 n_chunks,
                                      const u64 slot_begin, const u64 n_results, const u32 k, const u64 n_bases,
                                      u64 *__restrict__ starts) {
    for_result_slots(coff, clen, roff, chunk_begin, n_chunks, slot_begin, n_results, k, n_bases, [&](u64 slot, u64 start) { starts[slot] = start; },
                     [&](u64 slot) { starts[slot] = kNoKmer; });
}

// The next (at most 32) pattern characters of a strand search that has consumed `consumed` of its k
// characters, packed so that the next one to use sits in the low two bits.
__device__ __forceinline__ u64 long_window(const u64 *__restrict__ text, u64 s, u32 k, u32 strand, u32 consumed, u32 &len) {
    const u32 rem = k - consumed;
    len = rem < 32u ? rem : 32u;
    if (strand == 0) return window(text, s + rem - len, len);
    return revcomp_packed(window(text, s + consumed, len), len);
}

// k-mer == its own reverse complement? (general mode counts such a k-mer once, fms_index.h:318-323)
__device__ __forceinline__ bool long_self_complementary(const u64 *__restrict__ text, u64 s, u32 k) {
    if (k & 1u) return false;
    for (u32 c = 0; c < k; c += 32) {
        u32 la, lb;
        const u64 a = long_window(text, s, k, 0, c, la);
        const u64 b = long_window(text, s, k, 1, c, lb);
        if (a != b) return false;
    }
    return true;
}

template <int MODE, int OUT, int STRANDS, bool WIDE>
__global__ void __launch_bounds__(kQueryBlock)
long_query_kernel(const DevIndex d, const u64 *__restrict__ text, const u64 *__restrict__ starts, const u64 n,
                  void *__restrict__ out, unsigned long long *__restrict__ cursor, const u32 chunk, const GenF gf) {
    typedef typename PosT<WIDE>::type pos_t;
    const unsigned FULL = 0xffffffffu;
    const u32 lane = threadIdx.x & 31u;
    const u32 lt_mask = (1u << lane) - 1u;
    const u32 t = d.t, k = d.k;  // t <= 16 < 32 < k
    const u64 tmask = t ? ((1ull << (2 * t)) - 1ull) : 0ull;
    const bool need_j = !(OUT == K_OUT_PRESENCE && MODE == K_MODE_ALL);
    const u32 mm = d.multi_m;
    const u64 xmask = (1ull << (2 * mm)) - 1ull;

    // lane state
    bool active = false;
    u32 phase = PH_TABLE, strand = 0, steps = 0, navail = 0;
    u64 s = 0, pat = 0, idx = 0;
    pos_t i = 0, j = 0;
    long long res_f = 0;
    u64 g_ones = 0, g_total = 0;
    // warp state (uniform)
    u64 cend = 0, wnext = 0, tile_base = 0, bufA = 0, bufB = 0;
    bool exhausted = false;

    auto begin_strand = [&](u32 which) {
        strand = which;
        pat = long_window(text, s, k, which, 0, navail);
        if (t) {
            phase = PH_TABLE;
        } else {
            phase = PH_STEP;
            i = 0;
            j = (pos_t)d.n;
            steps = k;
        }
    };
    auto write_result = [&](u64 slot, long long rf, long long rr) {
        if (OUT == K_OUT_PRESENCE) {
            unsigned char v;
            if (MODE != K_MODE_GENERAL && STRANDS == K_STRANDS_BOTH) v = (unsigned char)((rf + 1) | ((rr + 1) << 2));
            else v = (unsigned char)(rr == 1);
            reinterpret_cast<unsigned char *>(out)[slot] = v;
        } else if (STRANDS == K_STRANDS_BOTH) {
            reinterpret_cast<longlong2 *>(out)[slot] = make_longlong2(rf, rr);
        } else {
            reinterpret_cast<long long *>(out)[slot] = rr;
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
                    bufA = (tile_base + lane < cend) ? starts[tile_base + lane] : kNoKmer;
                    bufB = (tile_base + 32 + lane < cend) ? starts[tile_base + 32 + lane] : kNoKmer;
                }
            }
            if (!exhausted) {
                const u32 pre = __popc(need & lt_mask);
                const u64 my = wnext + pre;
                const bool take = !active && my < cend;
                const u32 src = (u32)(my - tile_base);  // < 64
                u64 st = __shfl_sync(FULL, bufA, src & 31u);
                if (__any_sync(FULL, take && src >= 32u)) {
                    const u64 sb = __shfl_sync(FULL, bufB, src & 31u);
                    if (src >= 32u) st = sb;
                }
                const u64 left = cend - wnext;
                const u32 want = __popc(need);
                wnext += (want < left) ? want : left;
                if (wnext - tile_base >= 32) {  // uniform: slide the register tiles
                    tile_base += 32;
                    bufA = bufB;
                    bufB = (tile_base + 32 + lane < cend) ? starts[tile_base + 32 + lane] : kNoKmer;
                }
                if (take) {
                    if (st == kNoKmer) {  // a gap of the result layout: nothing to search
                        write_result(my, -1, -1);
                    } else {
                        active = true;
                        idx = my;
                        s = st;
                        g_ones = g_total = 0;
                        begin_strand(0);
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
        // multi-step probe (multistep.cuh) while the register window still holds mm bases
        const bool isX = isS && mm && steps >= mm && navail >= mm;
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

        // ---------------------------------------------------------------- consume
        bool done = false;   // this strand's search ended
        long long res = -1;  // its value
        if (isT) {
            i = ti;
            j = tj;
            pat >>= 2 * t;
            navail -= t;
            steps = k - t;
            if (i == j) done = true;
            else phase = PH_STEP;  // k > t always
        } else if (isS) {
            if (!two) {
                b0 = a0; b1 = a1; b2 = a2; b3 = a3;
            }
            if (isX) {
                const u32 oi = (u32)((u64)i - bi * MultiGeom<WIDE>::rows), oj = (u32)((u64)j - bj * MultiGeom<WIDE>::rows);
                i = lf_multi_t<WIDE>(a0, a1, a2, a3, oi);
                j = lf_multi_t<WIDE>(b0, b1, b2, b3, oj);
                pat >>= 2 * mm;
                navail -= mm;
                steps -= mm;
            } else {
                const u32 c = (u32)pat & 3u;
                pat >>= 2;
                --navail;
                const pos_t ni = lf_map<WIDE>(d, a0, a1, a2, a3, i, c);
                const pos_t nj = lf_map<WIDE>(d, b0, b1, b2, b3, j, c);
                i = ni;
                j = nj;
                --steps;
            }
            if (i == j) done = true;
            else if (steps == 0) phase = PH_MASK;
            else if (navail == 0) pat = long_window(text, s, k, strand, k - steps, navail);
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
                other = strand == 0 && !long_self_complementary(text, s, k);
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
                begin_strand(1);
            } else {
                write_result(idx, res_f, res);
                active = false;
            }
        }
    }
}

template <int MODE, int OUT, int STRANDS, bool WIDE>
int launch_long(int sm_count, const DevIndex &d, const u64 *text, const u64 *starts, size_t n, void *out,
                unsigned long long *cursor, const GenF &gf, cudaStream_t st) {
    auto kern = long_query_kernel<MODE, OUT, STRANDS, WIDE>;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kQueryBlock, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int grid = sm_count * per_sm;
    const size_t warps = (size_t)grid * (kQueryBlock / 32);
    size_t chunk = n / (warps * 8 + 1);
    chunk = (chunk / 32) * 32;
    if (chunk < 32) chunk = 32;
    if (chunk > 2048) chunk = 2048;
    cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, kQueryBlock, 0, st>>>(d, text, starts, (u64)n, out, cursor, (u32)chunk, gf);
    return (int)cudaGetLastError();
}

template <int MODE, int OUT, int STRANDS>
int launch_long_w(bool wide, int sm_count, const DevIndex &d, const u64 *text, const u64 *starts, size_t n, void *out,
                  unsigned long long *cursor, const GenF &gf, cudaStream_t st) {
    if (wide) return launch_long<MODE, OUT, STRANDS, true>(sm_count, d, text, starts, n, out, cursor, gf, st);
    return launch_long<MODE, OUT, STRANDS, false>(sm_count, d, text, starts, n, out, cursor, gf, st);
}

// mode: K_MODE_OR / K_MODE_ALL / K_MODE_GENERAL (with gf). Returns a cudaError_t as int (0 = ok).
inline int dispatch_long(bool wide, int sm_count, const DevIndex &d, int mode, int output, int strands, const GenF *gf,
                         const u64 *text, const u64 *starts, size_t n, void *out, unsigned long long *cursor, cudaStream_t st) {
    const GenF none{0, 0, 0};
    if (mode == K_MODE_GENERAL) return launch_long_w<K_MODE_GENERAL, K_OUT_PRESENCE, K_STRANDS_LAZY>(wide, sm_count, d, text, starts, n, out, cursor, *gf, st);
    if (output == K_OUT_ORDERS) {
        if (strands == K_STRANDS_BOTH) return launch_long_w<K_MODE_OR, K_OUT_ORDERS, K_STRANDS_BOTH>(wide, sm_count, d, text, starts, n, out, cursor, none, st);
        return launch_long_w<K_MODE_OR, K_OUT_ORDERS, K_STRANDS_LAZY>(wide, sm_count, d, text, starts, n, out, cursor, none, st);
    }
    if (mode == K_MODE_ALL) {
        if (strands == K_STRANDS_BOTH) return launch_long_w<K_MODE_ALL, K_OUT_PRESENCE, K_STRANDS_BOTH>(wide, sm_count, d, text, starts, n, out, cursor, none, st);
        return launch_long_w<K_MODE_ALL, K_OUT_PRESENCE, K_STRANDS_LAZY>(wide, sm_count, d, text, starts, n, out, cursor, none, st);
    }
    if (strands == K_STRANDS_BOTH) return launch_long_w<K_MODE_OR, K_OUT_PRESENCE, K_STRANDS_BOTH>(wide, sm_count, d, text, starts, n, out, cursor, none, st);
    return launch_long_w<K_MODE_OR, K_OUT_PRESENCE, K_STRANDS_LAZY>(wide, sm_count, d, text, starts, n, out, cursor, none, st);
}

}  // namespace fmsi
