// index_build.cuh — GPU construction of an FMS-index from a masked superstring.
//
// Scope: SURVEY §8(f) rank 3 — the reference's construct() (src/fms_index.h:397-460: QSufSort suffix
// array -> BWT split into ac_gt/ac/gt, SA-transformed mask, kLCP bits via construct_klcp :357-385)
// is single-threaded, 16 B/char and takes hours at 3.1 Gbp, which makes the human-scale configs
// unmeasurable. The suffix array is unique, so any correct suffix sorter yields the same index; this
// one produces the device layout directly and, through HostPlanes, the reference's files byte for
// byte (tests compare against `fmsi index` output).
//
// Suffix sorting: characters are coded $=0 < A=1 < C=2 < G=3 < T=4 in 3 bits, so a 63-bit key holds
// 21 characters and comparing keys IS comparing suffixes (the unique sentinel makes every suffix
// distinct, no special end-of-text rule). Round 0 buckets suffixes by first character and radix-sorts
// each bucket on the next 21 characters (cub::DeviceRadixSort); suffixes whose 22-character prefixes
// tie are refined in further rounds on the next 21 characters, sorted by (tie group, key), until no
// ties remain. For the benchmark's i.i.d. sequences round 1 already resolves everything.
#pragma once
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "device_index.cuh"

namespace fmsi {

#define BCU(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess) throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

template <typename T> struct DevArr {
    T *p = nullptr;
    size_t n = 0;
    DevArr() {}
    explicit DevArr(size_t n_) { alloc(n_); }
    DevArr(const DevArr &) = delete;
    DevArr &operator=(const DevArr &) = delete;
    void alloc(size_t n_) {
        release();
        n = n_;
        BCU(cudaMalloc(&p, (n_ ? n_ : 1) * sizeof(T)));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevArr() { release(); }
};

constexpr int kKeyChars = 21;

// ---- text ---------------------------------------------------------------------------------------
// Mask-cased ASCII -> packed 2-bit text (big-endian within a word, zero padded past n) and mask
// bits by text position (bit p & 63 of word p >> 6 = is_upper(ms[p]), fms_index.h:420).
__global__ void build_text_kernel(const char *__restrict__ ms, const u64 n, u64 *__restrict__ packed, const u64 n_words,
                                  u64 *__restrict__ maskbits, const u64 n_mask_words, int *__restrict__ bad) {
    const u64 w = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (w < n_words) {
        u64 v = 0;
        for (u32 t = 0; t < 32; ++t) {
            const u64 b = w * 32 + t;
            u32 code = 0;
            if (b < n) {
                const unsigned char ch = (unsigned char)ms[b];
                const unsigned char up = ch & 0xDF;
                if (up != 'A' && up != 'C' && up != 'G' && up != 'T') *bad = 1;
                const u32 x = (ch >> 1) & 3u;
                code = x ^ (x >> 1);
            }
            v = (v << 2) | code;
        }
        packed[w] = v;
    }
    if (w < n_mask_words) {
        u64 m = 0;
        for (u32 t = 0; t < 64; ++t) {
            const u64 b = w * 64 + t;
            if (b < n) {
                const unsigned char ch = (unsigned char)ms[b];
                if (ch >= 'A' && ch <= 'Z') m |= 1ull << t;
            }
        }
        maskbits[w] = m;
    }
}

__device__ __forceinline__ u32 text_code(const u64 *__restrict__ packed, u64 p) {
    return (u32)(packed[p >> 5] >> (62 - 2 * (p & 31))) & 3u;
}

// 21 characters starting at text position s as a 63-bit key ($=0, A..T = 1..4, 3 bits each).
__device__ __forceinline__ u64 suffix_key(const u64 *__restrict__ packed, u64 n, u64 s) {
    if (s >= n) return 0;
    const u64 w0 = packed[s >> 5], w1 = packed[(s >> 5) + 1];
    const u32 sh = 2u * ((u32)s & 31u);
    const u64 v = sh ? ((w0 << sh) | (w1 >> (64 - sh))) : w0;  // 32 chars, first char highest
    const u64 avail = n - s;
    u64 key = 0;
#pragma unroll
    for (int t = 0; t < kKeyChars; ++t) {
        const u64 c = (v >> (62 - 2 * t)) & 3ull;
        key = (key << 3) | (((u64)t < avail) ? c + 1 : 0ull);
    }
    return key;
}

__global__ void bucket_flags_kernel(const u64 *__restrict__ packed, u64 n, u64 base, u32 count, u32 c, unsigned char *__restrict__ flags) {
    const u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (q >= count) return;
    flags[q] = text_code(packed, base + q) == c;
}
__global__ void iota_kernel(u32 *__restrict__ out, u64 base, u32 count) {
    const u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (q < count) out[q] = (u32)(base + q);
}
__global__ void keys_kernel(const u64 *__restrict__ packed, u64 n, const u32 *__restrict__ pos, u64 count, u64 depth, u64 *__restrict__ keys) {
    const u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (q < count) keys[q] = suffix_key(packed, n, (u64)pos[q] + depth);
}
// tie[r] = 1 when the sorted item r has the same key as r-1 (same tie group)
__global__ void tie_flags_kernel(const u64 *__restrict__ keys, u64 count, unsigned char *__restrict__ tie) {
    const u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (q < count) tie[q] = q > 0 && keys[q] == keys[q - 1];
}
// For items of a bucket placed at SA slots [slot0, slot0+count): emit (group head slot, position)
// of every item that belongs to a tie group of size >= 2.
__global__ void head_slot_kernel(const unsigned char *__restrict__ tie, u64 count, u64 slot0, u64 *__restrict__ head) {
    const u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (q < count) head[q] = tie[q] ? 0ull : slot0 + q;  // max-scan turns this into the group head slot
}
__global__ void tied_flag_kernel(const unsigned char *__restrict__ tie, u64 count, unsigned char *__restrict__ tied) {
    const u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (q < count) tied[q] = tie[q] || (q + 1 < count && tie[q + 1]);
}
struct MaxOp {
    __device__ __forceinline__ u64 operator()(u64 a, u64 b) const { return a > b ? a : b; }
};
__global__ void pack_gq_kernel(const u64 *__restrict__ head, const u32 *__restrict__ pos, u64 count, u64 *__restrict__ gq) {
    const u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (q < count) gq[q] = (head[q] << 32) | (u64)pos[q];  // slots < 2^32 (narrow builder)
}
__global__ void unpack_pos_kernel(const u64 *__restrict__ gq, u64 count, u32 *__restrict__ pos, u32 *__restrict__ grp) {
    const u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (q < count) {
        pos[q] = (u32)gq[q];
        grp[q] = (u32)(gq[q] >> 32);
    }
}
__global__ void gather_u64_kernel(const u64 *__restrict__ src, const u32 *__restrict__ idx, u64 count, u64 *__restrict__ dst) {
    const u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (q < count) dst[q] = src[idx[q]];
}
// refinement: items sorted by (group, key). seg_start via max-scan of (new group ? q : 0).
__global__ void refine_flags_kernel(const u32 *__restrict__ grp, const u64 *__restrict__ keys, u64 count, u64 *__restrict__ segmark,
                                    unsigned char *__restrict__ tie) {
    const u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (q >= count) return;
    const bool newgrp = q == 0 || grp[q] != grp[q - 1];
    segmark[q] = newgrp ? q : 0ull;
    tie[q] = !newgrp && keys[q] == keys[q - 1];
}
__global__ void refine_place_kernel(const u32 *__restrict__ grp, const u64 *__restrict__ segstart, const u64 *__restrict__ gq, u64 count,
                                    u32 *__restrict__ sa, const unsigned char *__restrict__ tie, u64 *__restrict__ head) {
    const u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (q >= count) return;
    const u64 slot = (u64)grp[q] + (q - segstart[q]);
    sa[slot] = (u32)gq[q];
    head[q] = tie[q] ? 0ull : slot;
}

// ---- BWT / mask / kLCP planes from the suffix array ----------------------------------------------
// One warp per 64 SA slots. Outputs per block b: hi/lo planes, mask word, klcp word, symbol counts.
__global__ void planes_kernel(const u32 *__restrict__ sa, const u64 N, const u64 *__restrict__ packed, const u64 n,
                              const u64 *__restrict__ maskbits, const u32 k, const int with_klcp, u64 *__restrict__ lo_w,
                              u64 *__restrict__ hi_w, u64 *__restrict__ mask_w, u64 *__restrict__ klcp_w, u32 *__restrict__ cntA,
                              u32 *__restrict__ cntC, u32 *__restrict__ cntG, u32 *__restrict__ cntM, u64 *__restrict__ dollar,
                              const u64 nblk) {
    const u64 warp = (blockIdx.x * (u64)blockDim.x + threadIdx.x) >> 5;
    const u32 lane = threadIdx.x & 31u;
    if (warp >= nblk) return;
    u64 lo = 0, hi = 0, mk = 0, kl = 0;
    for (u32 half = 0; half < 2; ++half) {
        const u64 r = warp * 64 + half * 32 + lane;
        u32 sym = 0, mb = 0, kb = 0;
        if (r < N) {
            const u64 p = sa[r];
            if (p == 0) *dollar = r;  // BWT slot of '$', stored as A (fms_index.h:414-418)
            else sym = text_code(packed, p - 1);
            if (p < n) mb = (u32)(maskbits[p >> 6] >> (p & 63)) & 1u;
            if (with_klcp && r + 1 < N) {  // construct_klcp, fms_index.h:376-383
                const u64 p1 = sa[r + 1];
                const u64 km1 = k - 1;
                if (p + km1 <= n && p1 + km1 <= n) {
                    // equal (k-1)-mers at p and p1, compared 32 bases at a time (the reference keeps them in
                    // one uint64_t for k <= 32 and one __uint128_t for k <= 64, main.cpp:230-231)
                    kb = 1;
                    for (u64 off = 0; off < km1 && kb; off += 32) {
                        const u32 len = (u32)((km1 - off < 32) ? (km1 - off) : 32);
                        const u64 q = p + off, q1 = p1 + off;
                        const u64 w0 = packed[q >> 5], w1 = packed[(q >> 5) + 1];
                        const u32 s0 = 2u * ((u32)q & 31u);
                        const u64 a = (s0 ? ((w0 << s0) | (w1 >> (64 - s0))) : w0) >> (64 - 2 * len);
                        const u64 x0 = packed[q1 >> 5], x1 = packed[(q1 >> 5) + 1];
                        const u32 s1 = 2u * ((u32)q1 & 31u);
                        const u64 b = (s1 ? ((x0 << s1) | (x1 >> (64 - s1))) : x0) >> (64 - 2 * len);
                        kb = a == b;
                    }
                }
            }
        }
        const u32 blo = __ballot_sync(0xffffffffu, sym & 1u), bhi = __ballot_sync(0xffffffffu, sym & 2u);
        const u32 bm = __ballot_sync(0xffffffffu, mb), bk = __ballot_sync(0xffffffffu, kb);
        lo |= (u64)blo << (32 * half);
        hi |= (u64)bhi << (32 * half);
        mk |= (u64)bm << (32 * half);
        kl |= (u64)bk << (32 * half);
    }
    if (lane == 0) {
        const u64 base = warp * 64;
        const u32 valid = base >= N ? 0u : (u32)((N - base < 64) ? (N - base) : 64);
        const u64 vmask = valid == 64 ? ~0ull : ((1ull << valid) - 1ull);
        lo_w[warp] = lo;
        hi_w[warp] = hi;
        mask_w[warp] = mk;
        klcp_w[warp] = kl;
        cntA[warp] = (u32)__popcll(~hi & ~lo & vmask);
        cntC[warp] = (u32)__popcll(~hi & lo & vmask);
        cntG[warp] = (u32)__popcll(hi & ~lo & vmask);
        cntM[warp] = (u32)__popcll(mk);
    }
}

// Assemble the device layout (index_layout.hpp) from planes + exclusive prefix counts. Narrow only.
__global__ void assemble_blocks_kernel(const u64 *__restrict__ lo_w, const u64 *__restrict__ hi_w, const u64 *__restrict__ mask_w,
                                       const u64 *__restrict__ klcp_w, const u64 *__restrict__ exA, const u64 *__restrict__ exC,
                                       const u64 *__restrict__ exG, const u64 *__restrict__ exM, const u64 nblk, const u64 c0,
                                       const u64 c1, const u64 c2, const u64 c3, RankBlock *__restrict__ rank, AuxBlock *__restrict__ aux) {
    const u64 b = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const u64 a = exA[b], c = exC[b], g = exG[b];
    const u64 t = b * 64 - a - c - g;  // every earlier slot holds one of the four codes
    RankBlock rb;
    rb.cnt[0] = (uint32_t)(c0 + a);
    rb.cnt[1] = (uint32_t)(c1 + c);
    rb.cnt[2] = (uint32_t)(c2 + g);
    rb.cnt[3] = (uint32_t)(c3 + t);
    rb.lo = lo_w[b];
    rb.hi = hi_w[b];
    rank[b] = rb;
    AuxBlock ab;
    ab.klcp = klcp_w[b];
    ab.mask = mask_w[b];
    ab.mask_cum = exM[b];
    ab.spare = 0;
    aux[b] = ab;
}

struct BuiltIndex {
    u64 n_bwt = 0, dollar = 0, mask_ones = 0;
    u64 counts[4] = {0, 0, 0, 0};
    DevArr<RankBlock> rank;
    DevArr<AuxBlock> aux;
    // planes kept for fmsi_gpu_index_save (host copies)
    std::vector<u64> lo, hi, mask, klcp;
};

inline unsigned nblocks_for(u64 n, int block = 256) { return (unsigned)((n + block - 1) / block); }

// Sort (keys, vals) pairs on the low 63 bits.
template <typename V>
inline void radix_sort_pairs(DevArr<u64> &keys, DevArr<u64> &keys_alt, DevArr<V> &vals, DevArr<V> &vals_alt, u64 count, int end_bit = 63) {
    if (count == 0) return;
    cub::DoubleBuffer<u64> dk(keys.p, keys_alt.p);
    cub::DoubleBuffer<V> dv(vals.p, vals_alt.p);
    size_t tmp_bytes = 0;
    BCU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, count, 0, end_bit));
    DevArr<unsigned char> tmp(tmp_bytes);
    BCU(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, dk, dv, count, 0, end_bit));
    if (dk.Current() != keys.p) std::swap(keys.p, keys_alt.p);
    if (dv.Current() != vals.p) std::swap(vals.p, vals_alt.p);
}

inline void max_scan_inplace(u64 *data, u64 count) {
    if (count == 0) return;
    size_t tmp_bytes = 0;
    BCU(cub::DeviceScan::InclusiveScan(nullptr, tmp_bytes, data, data, MaxOp(), count));
    DevArr<unsigned char> tmp(tmp_bytes);
    BCU(cub::DeviceScan::InclusiveScan(tmp.p, tmp_bytes, data, data, MaxOp(), count));
}

// Exclusive prefix sums of 32-bit (or narrower) counts, accumulated in 64 bits: cub sums in the input's own type
// otherwise, and the mask of an index with more than 2^32 rows holds more than 2^32 ones (tests/test_gpu_wide.py).
template <typename InT>
struct CastToU64 {
    __host__ __device__ __forceinline__ u64 operator()(const InT &x) const { return (u64)x; }
};
template <typename InT>
inline void exclusive_sum_u64(const InT *in, u64 *out, u64 count) {
    thrust::transform_iterator<CastToU64<InT>, const InT *> it(in, CastToU64<InT>());
    size_t tmp_bytes = 0;
    BCU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, it, out, count));
    DevArr<unsigned char> tmp(tmp_bytes);
    BCU(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, it, out, count));
}

// Compact `vals[q]` where flags[q] != 0; returns the number selected.
template <typename T>
inline u64 select_flagged(const T *vals, const unsigned char *flags, T *out, u64 count) {
    if (count == 0) return 0;
    DevArr<u64> d_num(1);
    size_t tmp_bytes = 0;
    BCU(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, vals, flags, out, d_num.p, count));
    DevArr<unsigned char> tmp(tmp_bytes);
    BCU(cub::DeviceSelect::Flagged(tmp.p, tmp_bytes, vals, flags, out, d_num.p, count));
    u64 num = 0;
    BCU(cudaMemcpy(&num, d_num.p, 8, cudaMemcpyDeviceToHost));
    return num;
}

// Suffix array of text[0..n) + sentinel into sa[0..n] (device). n + 1 < 2^32.
inline void build_suffix_array(const u64 *d_packed, u64 n, u32 *d_sa, uint64_t *launches) {
    const u64 N = n + 1;
    u32 first = (u32)n;
    BCU(cudaMemcpy(d_sa, &first, 4, cudaMemcpyHostToDevice));  // SA[0] = n (the sentinel suffix)
    // tied items awaiting refinement, accumulated over buckets: gq = head_slot << 32 | pos
    std::vector<DevArr<u64> *> pending;
    std::vector<u64> pending_n;
    u64 slot0 = 1;
    const u64 kPiece = 1ull << 30;
    for (u32 c = 0; c < 4; ++c) {
        // positions p in [0, n) with text[p] == c, ascending (select in pieces below 2^31 items)
        DevArr<u32> pos(n / 2 + (1u << 20));  // a bucket rarely exceeds half the text; grown on demand below
        u64 cnt = 0;
        for (u64 base = 0; base < n; base += kPiece) {
            const u32 piece = (u32)std::min<u64>(kPiece, n - base);
            DevArr<unsigned char> flags(piece);
            DevArr<u32> ids(piece);
            bucket_flags_kernel<<<nblocks_for(piece), 256>>>(d_packed, n, base, piece, c, flags.p);
            iota_kernel<<<nblocks_for(piece), 256>>>(ids.p, base, piece);
            *launches += 2;
            DevArr<u32> sel(piece);
            const u64 got = select_flagged(ids.p, flags.p, sel.p, piece);
            if (cnt + got > pos.n) {
                DevArr<u32> bigger(cnt + got + (n - base - piece));
                BCU(cudaMemcpy(bigger.p, pos.p, cnt * 4, cudaMemcpyDeviceToDevice));
                std::swap(bigger.p, pos.p);
                std::swap(bigger.n, pos.n);
            }
            BCU(cudaMemcpy(pos.p + cnt, sel.p, got * 4, cudaMemcpyDeviceToDevice));
            cnt += got;
        }
        if (cnt == 0) continue;
        {
            DevArr<u64> keys(cnt), keys_alt(cnt);
            DevArr<u32> pos_alt(cnt);
            keys_kernel<<<nblocks_for(cnt), 256>>>(d_packed, n, pos.p, cnt, 1, keys.p);
            *launches += 1;
            radix_sort_pairs(keys, keys_alt, pos, pos_alt, cnt);
            BCU(cudaMemcpy(d_sa + slot0, pos.p, cnt * 4, cudaMemcpyDeviceToDevice));
            keys_alt.release();
            pos_alt.release();
            DevArr<unsigned char> tie(cnt), tied(cnt);
            tie_flags_kernel<<<nblocks_for(cnt), 256>>>(keys.p, cnt, tie.p);
            tied_flag_kernel<<<nblocks_for(cnt), 256>>>(tie.p, cnt, tied.p);
            DevArr<u64> head(cnt);
            head_slot_kernel<<<nblocks_for(cnt), 256>>>(tie.p, cnt, slot0, head.p);
            *launches += 3;
            keys.release();
            max_scan_inplace(head.p, cnt);
            DevArr<u64> gq(cnt);
            pack_gq_kernel<<<nblocks_for(cnt), 256>>>(head.p, pos.p, cnt, gq.p);
            *launches += 1;
            head.release();
            auto *sel = new DevArr<u64>(cnt);
            const u64 u = select_flagged(gq.p, tied.p, sel->p, cnt);
            if (u) {
                // shrink to fit
                auto *fit = new DevArr<u64>(u);
                BCU(cudaMemcpy(fit->p, sel->p, u * 8, cudaMemcpyDeviceToDevice));
                pending.push_back(fit);
                pending_n.push_back(u);
            }
            delete sel;
        }
        slot0 += cnt;
    }
    if (slot0 != N) throw std::runtime_error("suffix sort: bucket sizes do not add up");
    // ---- refinement rounds over all tied items
    u64 U = 0;
    for (u64 u : pending_n) U += u;
    DevArr<u64> gq(U ? U : 1);
    {
        u64 at = 0;
        for (size_t i = 0; i < pending.size(); ++i) {
            BCU(cudaMemcpy(gq.p + at, pending[i]->p, pending_n[i] * 8, cudaMemcpyDeviceToDevice));
            at += pending_n[i];
            delete pending[i];
        }
    }
    u64 depth = 1 + kKeyChars;
    while (U > 0) {
        if (U >= (1ull << 31)) throw std::runtime_error("suffix sort: too many tied suffixes for one refinement pass (highly repetitive input)");
        DevArr<u32> pos(U), grp(U);
        unpack_pos_kernel<<<nblocks_for(U), 256>>>(gq.p, U, pos.p, grp.p);
        DevArr<u64> keys(U), keys_alt(U), gq_alt(U);
        keys_kernel<<<nblocks_for(U), 256>>>(d_packed, n, pos.p, U, depth, keys.p);
        *launches += 2;
        // LSD: by key, then (stable) by group head slot carried in the high half of gq
        radix_sort_pairs(keys, keys_alt, gq, gq_alt, U);
        {   // second pass keyed on gq itself restricted to bits [32, 64): carries keys as values
            cub::DoubleBuffer<u64> dk(gq.p, gq_alt.p);
            cub::DoubleBuffer<u64> dv(keys.p, keys_alt.p);
            size_t tmp_bytes = 0;
            BCU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, U, 32, 64));
            DevArr<unsigned char> tmp(tmp_bytes);
            BCU(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, dk, dv, U, 32, 64));
            if (dk.Current() != gq.p) std::swap(gq.p, gq_alt.p);
            if (dv.Current() != keys.p) std::swap(keys.p, keys_alt.p);
        }
        keys_alt.release();
        unpack_pos_kernel<<<nblocks_for(U), 256>>>(gq.p, U, pos.p, grp.p);
        DevArr<u64> seg(U), head(U);
        DevArr<unsigned char> tie(U), tied(U);
        refine_flags_kernel<<<nblocks_for(U), 256>>>(grp.p, keys.p, U, seg.p, tie.p);
        *launches += 2;
        max_scan_inplace(seg.p, U);
        refine_place_kernel<<<nblocks_for(U), 256>>>(grp.p, seg.p, gq.p, U, d_sa, tie.p, head.p);
        tied_flag_kernel<<<nblocks_for(U), 256>>>(tie.p, U, tied.p);
        *launches += 2;
        max_scan_inplace(head.p, U);
        pack_gq_kernel<<<nblocks_for(U), 256>>>(head.p, pos.p, U, gq_alt.p);
        *launches += 1;
        const u64 u2 = select_flagged(gq_alt.p, tied.p, gq.p, U);
        U = u2;
        depth += kKeyChars;
    }
    BCU(cudaDeviceSynchronize());
}

// ms: mask-cased ASCII superstring on the device. Returns the device layout and host planes.
inline void build_index_on_device(const char *d_ms, u64 n, int k, bool with_klcp, bool keep_planes, BuiltIndex &out, uint64_t *launches) {
    if (n == 0) throw std::runtime_error("empty masked superstring");
    if (n + 1 >= (1ull << 32)) throw std::runtime_error("GPU index builder supports superstrings shorter than 2^32 - 1");
    if (k < 1 || k > 65536) throw std::runtime_error("GPU index builder supports 1 <= k <= 65536");
    const u64 N = n + 1, nblk = (N >> 6) + 1;
    const u64 n_words = (n + 31) / 32 + 4, n_mask_words = (n + 63) / 64 + 1;
    DevArr<u64> packed(n_words), maskbits(n_mask_words);
    DevArr<int> bad(1);
    BCU(cudaMemset(bad.p, 0, 4));
    build_text_kernel<<<nblocks_for(std::max(n_words, n_mask_words)), 256>>>(d_ms, n, packed.p, n_words, maskbits.p, n_mask_words, bad.p);
    *launches += 1;
    int h_bad = 0;
    BCU(cudaMemcpy(&h_bad, bad.p, 4, cudaMemcpyDeviceToHost));
    if (h_bad) throw std::runtime_error("masked superstring contains characters other than ACGTacgt");

    DevArr<u32> sa(N);
    build_suffix_array(packed.p, n, sa.p, launches);

    DevArr<u64> lo(nblk), hi(nblk), mk(nblk), kl(nblk), d_dollar(1);
    DevArr<u32> cA(nblk), cC(nblk), cG(nblk), cM(nblk);
    planes_kernel<<<nblocks_for(nblk * 32), 256>>>(sa.p, N, packed.p, n, maskbits.p, (u32)k, with_klcp ? 1 : 0, lo.p, hi.p, mk.p, kl.p, cA.p,
                                                   cC.p, cG.p, cM.p, d_dollar.p, nblk);
    *launches += 1;
    BCU(cudaDeviceSynchronize());
    sa.release();
    packed.release();
    maskbits.release();
    DevArr<u64> exA(nblk + 1), exC(nblk + 1), exG(nblk + 1), exM(nblk + 1);
    exclusive_sum_u64(cA.p, exA.p, nblk);
    exclusive_sum_u64(cC.p, exC.p, nblk);
    exclusive_sum_u64(cG.p, exG.p, nblk);
    exclusive_sum_u64(cM.p, exM.p, nblk);
    auto total = [&](DevArr<u64> &ex, DevArr<u32> &cnt) {
        u64 e = 0;
        u32 c = 0;
        BCU(cudaMemcpy(&e, ex.p + nblk - 1, 8, cudaMemcpyDeviceToHost));
        BCU(cudaMemcpy(&c, cnt.p + nblk - 1, 4, cudaMemcpyDeviceToHost));
        return e + c;
    };
    const u64 tA = total(exA, cA), tC = total(exC, cC), tG = total(exG, cG);
    out.mask_ones = total(exM, cM);
    out.n_bwt = N;
    out.counts[0] = 1;            // construct(), fms_index.h:451 (tA includes the '$' slot)
    out.counts[1] = tA;
    out.counts[2] = tA + tC;
    out.counts[3] = tA + tC + tG;
    BCU(cudaMemcpy(&out.dollar, d_dollar.p, 8, cudaMemcpyDeviceToHost));
    out.rank.alloc(nblk);
    out.aux.alloc(nblk);
    assemble_blocks_kernel<<<nblocks_for(nblk), 256>>>(lo.p, hi.p, mk.p, kl.p, exA.p, exC.p, exG.p, exM.p, nblk, out.counts[0], out.counts[1],
                                                       out.counts[2], out.counts[3], out.rank.p, out.aux.p);
    *launches += 1;
    BCU(cudaDeviceSynchronize());
    if (keep_planes) {
        out.lo.resize(nblk);
        out.hi.resize(nblk);
        out.mask.resize(nblk);
        out.klcp.resize(nblk);
        BCU(cudaMemcpy(out.lo.data(), lo.p, nblk * 8, cudaMemcpyDeviceToHost));
        BCU(cudaMemcpy(out.hi.data(), hi.p, nblk * 8, cudaMemcpyDeviceToHost));
        BCU(cudaMemcpy(out.mask.data(), mk.p, nblk * 8, cudaMemcpyDeviceToHost));
        BCU(cudaMemcpy(out.klcp.data(), kl.p, nblk * 8, cudaMemcpyDeviceToHost));
    }
}

}  // namespace fmsi
