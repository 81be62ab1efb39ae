// stream_kernels.cuh — chunk helpers and the kLCP streaming kernel.
//
// stream_kernel restates query_kmers_streaming (reference src/fms_index.h:181-254) with one LANE per
// chunk of <= 64 k-mers: pass 0 walks the chunk right to left on the forward strand, pass 1 walks
// the reverse-complement strand over the positions pass 0 left undecided (all positions for
// STRANDS_BOTH). While the interval of the neighbouring k-mer is non-empty the next k-mer costs
//     MX   : one aux sector (two if [i, j) straddles a block) — the mask probe of the current k-mer
//            AND the kLCP extension (extend_range_with_klcp, :106-109) for the next one come from
//            the same sectors, because mask and kLCP bits of a 64-position block share a sector;
//     STEP : one LF-step (update_range, :98-103);
// after a miss the search restarts from the suffix table like a single query. Lanes are refilled
// with the next chunk as soon as they finish, exactly as in query_kmers_kernel. The chunk's bases
// (<= 95) live in three registers, so no base is re-read from memory.
//
// The engine may chunk differently from ms_query (main.cpp:329-331): per-k-mer strand results do
// not depend on chunk boundaries; the boundaries only feed the strand predictor, which the host
// replays on its own (predictor.hpp).
#pragma once
#include "query_kernels.cuh"

namespace fmsi {

constexpr int kStreamBlock = 256;
constexpr u32 kMaxStreamKmers = 64;

// ASCII ACGTacgt -> 2-bit code (A=0 C=1 G=2 T=3): bits 1-2 of the character are 00,01,11,10.
__device__ __forceinline__ u32 base_code(unsigned char ch) {
    const u32 x = (ch >> 1) & 3u;
    return x ^ (x >> 1);
}

// Four ASCII bases in a 32-bit word (first base in the lowest byte) -> their 2-bit codes in one byte, first base
// highest. Per byte: code = x ^ (x >> 1), x = (ch >> 1) & 3; the multiply gathers the four 2-bit fields (byte i's
// field lands at bit 30 - 2i; all other partial products fall below bit 24 without overlapping, or overflow).
__device__ __forceinline__ u32 pack4_bases(u32 w) {
    u32 y = (w >> 1) & 0x03030303u;
    y ^= (y >> 1) & 0x01010101u;
    return (y * 0x40100401u) >> 24;
}
__device__ __forceinline__ u64 pack16_bases(const uint4 x) {
    return ((u64)pack4_bases(x.x) << 24) | ((u64)pack4_bases(x.y) << 16) | ((u64)pack4_bases(x.z) << 8) | (u64)pack4_bases(x.w);
}

// Packed text: base b sits at bits [62 - 2(b & 31), 63 - 2(b & 31)] of word b >> 5, so that a window
// read is already in k-mer order (first base highest). One thread per word: two 16-byte loads when the
// text is 16-byte aligned (device scratch always is), bytes otherwise and at the tail.
__global__ void pack_bases_kernel(const char *__restrict__ bases, const u64 n_bases, u64 *__restrict__ packed,
                                  const u64 n_words) {
    const u64 w = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    u64 v = 0;
    const u64 b0 = w * 32;
    if ((reinterpret_cast<unsigned long long>(bases) & 15ull) == 0 && b0 + 32 <= n_bases) {
        const uint4 *p = reinterpret_cast<const uint4 *>(bases + b0);
        const uint4 x0 = __ldg(p), x1 = __ldg(p + 1);
        v = (pack16_bases(x0) << 32) | pack16_bases(x1);
    } else {
#pragma unroll 8
        for (u32 t = 0; t < 32; ++t) {
            const u64 b = b0 + t;
            const u32 code = (b < n_bases) ? base_code((unsigned char)bases[b]) : 0u;
            v = (v << 2) | code;
        }
    }
    packed[w] = v;
}

// Presence results, one byte (0 / 1) per k-mer -> one bit per k-mer, bit q & 7 of byte q >> 3 (the characters the
// reference prints, 8 per byte): FMSI_GPU_OUT_PRESENCE_BITS. One thread per output byte.
__global__ void pack_presence_bits_kernel(const unsigned char *__restrict__ res, const u64 n, unsigned char *__restrict__ bits) {
    const u64 b = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    const u64 q0 = b * 8;
    if (q0 >= n) return;
    u64 v = 0;
    if (q0 + 8 <= n && (reinterpret_cast<unsigned long long>(res) & 7ull) == 0) {
        v = *reinterpret_cast<const u64 *>(res + q0);
    } else {
        for (u32 t = 0; t < 8 && q0 + t < n; ++t) v |= (u64)res[q0 + t] << (8 * t);
    }
    bits[b] = (unsigned char)(((v & 0x0101010101010101ull) * 0x0102040810204080ull) >> 56);
}

// len (<= 32) bases starting at base s, as a packed k-mer.
__device__ __forceinline__ u64 window(const u64 *__restrict__ packed, u64 s, u32 len) {
    const u64 w0 = __ldg(packed + (s >> 5)), w1 = __ldg(packed + (s >> 5) + 1);
    const u32 sh = 2u * ((u32)s & 31u);
    const u64 v = sh ? ((w0 << sh) | (w1 >> (64 - sh))) : w0;
    return v >> (64 - 2 * len);
}

// Slot -> chunk resolution shared by extract_kmers_kernel / extract_starts_kernel (longk_kernels.cuh): result slot r
// belongs to the last chunk c with res_off[c] <= r (res_off[] non-decreasing). A binary search per slot is a
// log2(n_chunks)-deep chain of dependent loads per warp and made this the slowest kernel of the reads path
// (2.5 ms for 120 M slots). Here a warp owns kSlotsPerWarp = 31 * 32 consecutive slots: its 32 lanes search the 32
// sub-block boundaries at once (one deep chain per 992 slots, 32 of them in flight per warp), then every slot
// only searches between the chunks of its sub-block's two boundaries — a step or two, whatever the chunk sizes
// (one 5 Mbp chunk or ten million single-k-mer chunks). f(slot, start_base) stores the k-mer that starts at base
// `start_base`; g(slot) stores the value of slots that belong to no k-mer.
constexpr u32 kSlotsPerWarp = 31 * 32;

template <typename Fill, typename Gap>
__device__ __forceinline__ void for_result_slots(const u64 *__restrict__ coff, const u32 *__restrict__ clen, const u64 *__restrict__ roff,
                                                 const u64 chunk_begin, const u64 n_chunks, const u64 slot_begin, const u64 n_results,
                                                 const u32 k, const u64 n_bases, Fill f, Gap g) {
    // slots [slot_begin, n_results), which belong to chunks [chunk_begin, n_chunks) (a pipelined host call resolves
    // its result spans one after the other, each with the chunk range uploaded so far)
    const unsigned FULL = 0xffffffffu;
    const u32 lane = threadIdx.x & 31u;
    const u64 warp = (blockIdx.x * (u64)blockDim.x + threadIdx.x) >> 5;
    const u64 base = slot_begin + warp * kSlotsPerWarp;
    if (base >= n_results) return;  // uniform per warp
    u64 target = base + 32ull * lane;
    if (target >= n_results) target = n_results - 1;
    u64 lo = chunk_begin, hi = n_chunks;
    while (hi - lo > 1) {
        const u64 mid = (lo + hi) >> 1;
        if (__ldg(roff + mid) <= target) lo = mid;
        else hi = mid;
    }
    const u64 bound = lo;  // chunk of the first slot of sub-block `lane`
#pragma unroll 4
    for (u32 j = 0; j < 31; ++j) {
        u64 c0 = __shfl_sync(FULL, bound, j), c1 = __shfl_sync(FULL, bound, j + 1) + 1;
        const u64 r = base + 32ull * j + lane;
        if (r >= n_results) continue;  // no shuffles below this point
        while (c1 - c0 > 1) {
            const u64 mid = (c0 + c1) >> 1;
            if (__ldg(roff + mid) <= r) c0 = mid;
            else c1 = mid;
        }
        const u64 r0 = __ldg(roff + c0);
        const u32 len = __ldg(clen + c0);
        const u64 pos = r - r0;
        const u64 cs = __ldg(coff + c0);
        // a chunk that runs past the text (device-mode callers are not validated on the host) yields no k-mers
        if (r0 <= r && len >= k && pos + k <= len && cs + len <= n_bases) f(r, cs + pos);
        else g(r);
    }
}

inline unsigned slot_blocks(u64 n_results, int block = 256) {
    const u64 warps = (n_results + kSlotsPerWarp - 1) / kSlotsPerWarp;
    return (unsigned)((warps * 32 + block - 1) / block);
}

// Non-streaming chunks: the k-mer of every result slot, materialised for the single-query kernels, which
// then run over the flat array. Slots that belong to no k-mer (gaps) get the k-mer 0.
__global__ void extract_kmers_kernel(const u64 *__restrict__ packed, const u64 *__restrict__ coff,
                                     const u32 *__restrict__ clen, const u64 *__restrict__ roff, const u64 chunk_begin,
                                     const u64 n_chunks, const u64 slot_begin, const u64 n_results, const u32 k,
                                     const u64 n_bases, u64 *__restrict__ kmers) {
    for_result_slots(coff, clen, roff, chunk_begin, n_chunks, slot_begin, n_results, k, n_bases, [&](u64 slot, u64 start) { kmers[slot] = window(packed, start, k); },
                     [&](u64 slot) { kmers[slot] = 0; });
}

// Reads -> chunks, on the device (fmsi_gpu_query_reads_packed): read r = bases [roff[r], roff[r+1]) yields
// nk = max(0, len - k + 1) results, cut into chunks of at most `max_kmers` k-mers overlapping by k-1 (the streaming
// kernel's limit), or one chunk per read when max_kmers == 0 (the single-query kernels take chunks of any length; a read
// shorter than k becomes a chunk without results). rbase / cbase = exclusive prefix sums of nk / chunk counts.
__global__ void read_counts_kernel(const u64 *__restrict__ roff, const u64 n_reads, const u64 n_bases, const u32 k, const u32 max_kmers,
                                   u32 *__restrict__ nk, u32 *__restrict__ nch) {
    const u64 r = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const u64 a = roff[r], b = roff[r + 1];
    const u64 len = (b >= a && b <= n_bases) ? b - a : 0;  // a malformed read (device-mode callers) yields nothing
    const u64 m = len >= k ? len - k + 1 : 0;
    nk[r] = (u32)(m > 0xFFFFFF00ull ? 0xFFFFFF00ull : m);
    nch[r] = max_kmers ? (u32)((m + max_kmers - 1) / max_kmers) : 1u;
}
__global__ void expand_reads_kernel(const u64 *__restrict__ roff, const u64 n_reads, const u64 n_bases, const u32 k, const u32 max_kmers,
                                    const u32 *__restrict__ nk, const u64 *__restrict__ rbase, const u64 *__restrict__ cbase,
                                    u64 *__restrict__ coff, u32 *__restrict__ clen, u64 *__restrict__ cres) {
    const u64 r = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const u64 a = roff[r], m = nk[r], c0 = cbase[r], r0 = rbase[r];
    if (!max_kmers) {
        coff[c0] = a < n_bases ? a : 0;
        clen[c0] = m ? (u32)(m + k - 1) : 0u;
        cres[c0] = r0;
        return;
    }
    for (u64 p = 0, c = c0; p < m; p += max_kmers, ++c) {
        const u64 take = m - p < max_kmers ? m - p : max_kmers;
        coff[c] = a + p;
        clen[c] = (u32)(take + k - 1);
        cres[c] = r0 + p;
    }
}

enum { SP_TABLE = 0, SP_STEP = 1, SP_MX = 2 };

__device__ __forceinline__ u64 sel3(u32 w, u64 w0, u64 w1, u64 w2) { return w == 0 ? w0 : (w == 1 ? w1 : w2); }

template <int MODE, int OUT, int STRANDS, bool WIDE>
__global__ void __launch_bounds__(kStreamBlock)
stream_kernel(const DevIndex d, const u64 *__restrict__ packed, const u64 n_bases, const u64 *__restrict__ coff,
              const u32 *__restrict__ clen, const u64 *__restrict__ roff, const u64 n_chunks, void *out,
              unsigned long long *__restrict__ cursor, const u32 grab, unsigned long long *__restrict__ probe_ctr) {
    typedef typename PosT<WIDE>::type pos_t;
    const unsigned FULL = 0xffffffffu;
    const u32 lane = threadIdx.x & 31u;
    const u32 lt_mask = (1u << lane) - 1u;
    const u32 t = d.t, k = d.k;
    const u64 tmask = t ? ((t >= 32) ? ~0ull : ((1ull << (2 * t)) - 1ull)) : 0ull;
    const bool need_j = !(OUT == K_OUT_PRESENCE && MODE == K_MODE_ALL);
    const u32 mm = d.multi_m;
    const u64 xmask = (1ull << (2 * mm)) - 1ull;

    bool active = false;
    u32 phase = SP_TABLE, pass = 0, p = 0, nk = 0, steps = 0;
    u64 ro = 0, decided = 0, pat = 0, w0 = 0, w1 = 0, w2 = 0;
    pos_t i = 0, j = 0;
    u64 cnext = 0, cend = 0;
    bool exhausted = false;
    u32 nprobe = 0;  // dependent memory requests issued by this lane (reported when probe_ctr is given)

    // k-mer at chunk position q from the register-resident bases
    auto kmer_at = [&](u32 q) -> u64 {
        const u32 wi = q >> 5, sh = 2u * (q & 31u);
        const u64 a = sel3(wi, w0, w1, w2), b = sel3(wi + 1, w0, w1, w2);
        const u64 v = sh ? ((a << sh) | (b >> (64 - sh))) : a;
        return v >> (64 - 2 * k);
    };
    auto base_at = [&](u32 q) -> u32 {
        return (u32)(sel3(q >> 5, w0, w1, w2) >> (62 - 2 * (q & 31u))) & 3u;
    };
    auto begin_fresh = [&](u32 q) {
        const u64 km = kmer_at(q);
        pat = pass ? revcomp_packed(km, k) : km;
        if (t) {
            phase = SP_TABLE;
        } else {
            phase = SP_STEP;
            i = 0;
            j = (pos_t)d.n;
            steps = k;
        }
    };

    for (;;) {
        // ---------------------------------------------------------------- refill idle lanes
        const unsigned need = __ballot_sync(FULL, !active);
        if (need && !exhausted) {
            if (cnext >= cend) {
                unsigned long long c0 = 0;
                if (lane == 0) c0 = atomicAdd(cursor, (unsigned long long)grab);
                c0 = __shfl_sync(FULL, c0, 0);
                if (c0 >= n_chunks) exhausted = true;
                else {
                    cnext = c0;
                    cend = (c0 + grab < n_chunks) ? c0 + grab : n_chunks;
                }
            }
            if (!exhausted) {
                const u64 my = cnext + __popc(need & lt_mask);
                const bool take = !active && my < cend;
                const u64 left = cend - cnext;
                const u32 want = __popc(need);
                cnext += (want < left) ? want : left;
                if (take) {
                    const u64 cs = coff[my];
                    const u32 len = clen[my];
                    ro = roff[my];
                    nk = len - k + 1;
                    if (len >= k && nk <= kMaxStreamKmers && cs + len <= n_bases) {
                        // bases cs .. cs+95 as three aligned-to-chunk words
                        const u64 *pw = packed + (cs >> 5);
                        const u64 q0 = __ldg(pw), q1 = __ldg(pw + 1), q2 = __ldg(pw + 2);
                        const u32 sh = 2u * ((u32)cs & 31u);
                        if (sh) {
                            const u64 q3 = __ldg(pw + 3);
                            w0 = (q0 << sh) | (q1 >> (64 - sh));
                            w1 = (q1 << sh) | (q2 >> (64 - sh));
                            w2 = (q2 << sh) | (q3 >> (64 - sh));
                        } else {
                            w0 = q0;
                            w1 = q1;
                            w2 = q2;
                        }
                        active = true;
                        pass = 0;
                        decided = 0;
                        p = nk - 1;
                        begin_fresh(p);
                    }
                }
            }
        }
        if (!__any_sync(FULL, active)) {
            if (exhausted) break;
            continue;
        }

        // ---------------------------------------------------------------- loads
        const bool isT = active && phase == SP_TABLE;
        const bool isS = active && phase == SP_STEP;
        const bool isM = active && phase == SP_MX;
        // Will the next k-mer of this pass continue from this interval?
        bool will_cont = false;
        if (isM) {
            if (pass == 0) will_cont = p > 0;
            else will_cont = (p + 1 < nk) && (STRANDS == K_STRANDS_BOTH || !((decided >> (p + 1)) & 1ull));
        }
        // a fresh search (after a miss) takes multi-step probes like a single query (multistep.cuh); the one
        // step that continues a neighbour's interval is a plain LF-step
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
        const bool two = (isS || (isM && (need_j || will_cont))) && (bj != bi);
        u64 a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
        pos_t ti = 0, tj = 0;
        if (isT) ld_table<WIDE>(d, pat & tmask, ti, tj);
        if (isS || isM) {
            ld_sector_l1(pa, a0, a1, a2, a3);
            if (two) ld_sector_l1(pb, b0, b1, b2, b3);
        }
        nprobe += (u32)isT + (u32)(isS || isM) + (u32)two;

        // ---------------------------------------------------------------- consume
        bool finished = false;  // the k-mer at position p got its value for this pass
        bool nonempty = false;
        long long res = -1;
        if (isT) {
            i = ti;
            j = tj;
            pat >>= 2 * t;
            steps = k - t;
            if (i == j) finished = true;
            else phase = steps ? SP_STEP : SP_MX;
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
            if (i == j) finished = true;
            else if (steps == 0) phase = SP_MX;
        } else if (isM) {
            if (!two) {
                b0 = a0; b1 = a1; b2 = a2;
            }
            res = strand_result<MODE, OUT>((u64)i, (u64)j, a1, a2, b1, b2);
            finished = true;
            nonempty = true;
        }

        if (finished) {
            // ---- record (fms_index.h:200-207 / :224-233)
            const u64 slot = ro + p;
            if (pass == 0) {
                bool dec;
                if (OUT == K_OUT_ORDERS) dec = res >= 0;
                else if (MODE == K_MODE_ALL) dec = res != -1;
                else dec = res == 1;
                if (dec) decided |= 1ull << p;
                if (OUT == K_OUT_PRESENCE) {
                    reinterpret_cast<unsigned char *>(out)[slot] =
                        (STRANDS == K_STRANDS_BOTH) ? (unsigned char)(res + 1) : (unsigned char)(res == 1);
                } else if (STRANDS == K_STRANDS_BOTH) {
                    reinterpret_cast<long long *>(out)[2 * slot] = res;
                } else {
                    reinterpret_cast<long long *>(out)[slot] = res;
                }
            } else {
                if (OUT == K_OUT_PRESENCE) {
                    unsigned char *o = reinterpret_cast<unsigned char *>(out) + slot;
                    if (STRANDS == K_STRANDS_BOTH) *o = (unsigned char)(*o | ((res + 1) << 2));
                    else *o = (unsigned char)(res == 1);
                } else if (STRANDS == K_STRANDS_BOTH) {
                    reinterpret_cast<long long *>(out)[2 * slot + 1] = res;
                } else {
                    reinterpret_cast<long long *>(out)[slot] = res;
                }
            }
            // ---- next position of this pass, or next pass, or chunk done
            bool have_next = false, adjacent = false;
            u32 q = 0;
            if (pass == 0) {
                if (p > 0) {
                    q = p - 1;
                    have_next = adjacent = true;
                } else {
                    pass = 1;
                    const u64 und = (STRANDS == K_STRANDS_BOTH) ? ~0ull : ~decided;
                    const u64 cand = und & ((nk >= 64) ? ~0ull : ((1ull << nk) - 1ull));
                    if (cand) {
                        q = (u32)__ffsll((long long)cand) - 1;
                        have_next = true;
                    }
                }
            } else {
                const u64 und = (STRANDS == K_STRANDS_BOTH) ? ~0ull : ~decided;
                u64 cand = und & ((nk >= 64) ? ~0ull : ((1ull << nk) - 1ull));
                cand &= ~((2ull << p) - 1ull);  // positions > p
                if (cand) {
                    q = (u32)__ffsll((long long)cand) - 1;
                    have_next = true;
                    adjacent = q == p + 1;
                }
            }
            if (!have_next) {
                active = false;
            } else if (nonempty && adjacent) {
                // ---- extend_range_with_klcp from the sectors already in registers, then one step
                u64 ei = (u64)i, ej = (u64)j;
                bool slow = false;
                {
                    const u64 pj = ej - 1;  // block bj: klcp word b0 (== a0 when !two)
                    const u64 z = ~b0 & ~low_mask((u32)pj & 63u);
                    if (z) ej = (pj & ~63ull) + (u64)(__ffsll((long long)z) - 1) + 1;
                    else slow = true;
                }
                {
                    const u64 pi = ei - 1;
                    if ((pi >> 6) == bi) {
                        const u64 z = ~a0 & ((2ull << (pi & 63)) - 1ull);
                        if (z) ei = (pi & ~63ull) + (u64)(63 - __clzll((long long)z)) + 1;
                        else slow = true;
                    } else {
                        slow = true;
                    }
                }
                if (slow) {  // run of ones crosses a block boundary: rare, plain loop
                    ei = (u64)i;
                    ej = (u64)j;
                    dev_extend_klcp(d, ei, ej);
                }
                i = (pos_t)ei;
                j = (pos_t)ej;
                // pass 0 prepends base q; pass 1 prepends the complement of base q+k-1
                pat = pass == 0 ? base_at(q) : (3u - base_at(q + k - 1));
                steps = 1;
                phase = SP_STEP;
                p = q;
            } else {
                p = q;
                begin_fresh(q);
            }
        }
    }
    count_probes(probe_ctr, nprobe);
}

template <int MODE, int OUT, int STRANDS, bool WIDE>
int launch_stream(int sm_count, const DevIndex &d, const u64 *packed, u64 n_bases, const u64 *coff, const u32 *clen, const u64 *roff,
                  size_t n_chunks, void *out, unsigned long long *cursor, cudaStream_t st, unsigned long long *probe_ctr) {
    auto kern = stream_kernel<MODE, OUT, STRANDS, WIDE>;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kStreamBlock, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int grid = sm_count * per_sm;
    const size_t warps = (size_t)grid * (kStreamBlock / 32);
    size_t grab = n_chunks / (warps * 8 + 1);
    if (grab < 32) grab = 32;
    if (grab > 512) grab = 512;
    cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, kStreamBlock, 0, st>>>(d, packed, n_bases, coff, clen, roff, (u64)n_chunks, out, cursor, (u32)grab, probe_ctr);
    return (int)cudaGetLastError();
}

template <int MODE, int OUT, int STRANDS>
int launch_stream_w(bool wide, int sm_count, const DevIndex &d, const u64 *packed, u64 n_bases, const u64 *coff, const u32 *clen,
                    const u64 *roff, size_t n_chunks, void *out, unsigned long long *cursor, cudaStream_t st, unsigned long long *probe_ctr) {
    if (wide) return launch_stream<MODE, OUT, STRANDS, true>(sm_count, d, packed, n_bases, coff, clen, roff, n_chunks, out, cursor, st, probe_ctr);
    return launch_stream<MODE, OUT, STRANDS, false>(sm_count, d, packed, n_bases, coff, clen, roff, n_chunks, out, cursor, st, probe_ctr);
}

// returns a cudaError_t as int (0 = ok)
inline int dispatch_stream(bool wide, int sm_count, const DevIndex &d, int mode, int output, int strands,
                           const u64 *packed, u64 n_bases, const u64 *coff, const u32 *clen, const u64 *roff, size_t n_chunks,
                           void *out, unsigned long long *cursor, cudaStream_t st, unsigned long long *probe_ctr = nullptr) {
    if (output == K_OUT_ORDERS) {
        if (strands == K_STRANDS_BOTH) return launch_stream_w<K_MODE_OR, K_OUT_ORDERS, K_STRANDS_BOTH>(wide, sm_count, d, packed, n_bases, coff, clen, roff, n_chunks, out, cursor, st, probe_ctr);
        return launch_stream_w<K_MODE_OR, K_OUT_ORDERS, K_STRANDS_LAZY>(wide, sm_count, d, packed, n_bases, coff, clen, roff, n_chunks, out, cursor, st, probe_ctr);
    }
    if (mode == K_MODE_ALL) {
        if (strands == K_STRANDS_BOTH) return launch_stream_w<K_MODE_ALL, K_OUT_PRESENCE, K_STRANDS_BOTH>(wide, sm_count, d, packed, n_bases, coff, clen, roff, n_chunks, out, cursor, st, probe_ctr);
        return launch_stream_w<K_MODE_ALL, K_OUT_PRESENCE, K_STRANDS_LAZY>(wide, sm_count, d, packed, n_bases, coff, clen, roff, n_chunks, out, cursor, st, probe_ctr);
    }
    if (strands == K_STRANDS_BOTH) return launch_stream_w<K_MODE_OR, K_OUT_PRESENCE, K_STRANDS_BOTH>(wide, sm_count, d, packed, n_bases, coff, clen, roff, n_chunks, out, cursor, st, probe_ctr);
    return launch_stream_w<K_MODE_OR, K_OUT_PRESENCE, K_STRANDS_LAZY>(wide, sm_count, d, packed, n_bases, coff, clen, roff, n_chunks, out, cursor, st, probe_ctr);
}

}  // namespace fmsi
