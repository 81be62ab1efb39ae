// device_index.cuh — device-side view of the FMS-index and the rank / LF-mapping primitives.
// Everything here is integer work on 32-byte sectors; one LF probe = one 256-bit load.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "index_layout.hpp"

namespace fmsi {

typedef uint64_t u64;
typedef unsigned int u32;

// Multi-step rank sector (multistep.cuh): 224 SA rows of ONE m-mer x of preceding bases.
//   cnt  : C_m[x] + #{rows r < 224 b whose suffix is preceded by x}   (C_m[x] = #suffixes smaller than x)
//   bits : bit r % 224 set iff row r's suffix is preceded by x (first bit = lowest bit of bits[0])
// so m LF-steps (m applications of update_range, reference src/fms_index.h:98-103) cost one sector.
struct alignas(32) MultiBlock {
    uint32_t cnt;
    uint32_t bits[7];
};
static_assert(sizeof(MultiBlock) == 32, "one sector");
constexpr uint32_t kMultiRows = 224;
// The same sector for indexes of 2^32 rows and more: a 64-bit absolute counter leaves 192 rows per sector.
struct alignas(32) MultiBlockWide {
    uint64_t cnt;
    uint32_t bits[6];
};
static_assert(sizeof(MultiBlockWide) == 32, "one sector");
constexpr uint32_t kMultiRowsWide = 192;
template <bool WIDE> struct MultiGeom { static constexpr uint32_t rows = kMultiRows, hdr = 1; };
template <> struct MultiGeom<true> { static constexpr uint32_t rows = kMultiRowsWide, hdr = 2; };

struct DevIndex {
    const RankBlock *rank;
    const AuxBlock *aux;
    const MultiBlock *multi;  // [4^m][multi_nblk], x-major; null when the tier is not built (MultiBlockWide sectors when wide)
    u32 multi_m;              // bases per multi-step probe (0 = off, else 2 or 3)
    u32 multi_nblk;           // N / 224 + 1 (wide: N / 192 + 1)
    const void *table;    // 4^t entries, 1 << tshift bytes apart, each starting with {i, j}: u32 pairs
                          // (narrow; 32-byte dictionary buckets when tshift == 5) or u64 pairs (wide)
    const u64 *sb_base;   // [n_superblocks][4]
    u64 n;                // N = BWT length
    u64 dollar;
    u32 t;                // suffix-table depth (0 = no table)
    u32 tshift;           // log2 of the table entry stride in bytes
    u32 sb_shift;
    u32 k;
    u32 has_klcp;
};

template <bool WIDE> struct PosT { typedef u32 type; };
template <> struct PosT<true> { typedef u64 type; };

// L2 fill granularity of the random probes. A plain load makes L2 fetch the whole 128-byte line from DRAM for a
// 32-byte probe (ncu, r01b: 3.7x the requested bytes, DRAM at 0.89 of peak); `.L2::64B` asks for a 64-byte fill.
// -DFMSI_LD_PLAIN builds the A/B variant without the qualifier.
#ifdef FMSI_LD_PLAIN
#define FMSI_L2_HINT ""
#else
#define FMSI_L2_HINT ".L2::64B"
#endif

// One sector, bypassing L1 allocation (no reuse inside an SM for random probes).
__device__ __forceinline__ void ld_sector(const void *p, u64 &a, u64 &b, u64 &c, u64 &d) {
    asm volatile("ld.global.nc.L1::no_allocate" FMSI_L2_HINT ".v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
                 : "l"(p));
}
// One sector through L1 (streaming path: consecutive probes revisit the same block).
__device__ __forceinline__ void ld_sector_l1(const void *p, u64 &a, u64 &b, u64 &c, u64 &d) {
    asm volatile("ld.global.nc" FMSI_L2_HINT ".v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
                 : "l"(p));
}
// 8 / 16 bytes of a random table entry
__device__ __forceinline__ uint2 ld_pair32(const void *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate" FMSI_L2_HINT ".v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ ulonglong2 ld_pair64(const void *p) {
    ulonglong2 r;
    asm volatile("ld.global.nc.L1::no_allocate" FMSI_L2_HINT ".v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}

__device__ __forceinline__ u64 low_mask(u32 nbits) {  // nbits in [0, 63]
    return (1ull << nbits) - 1ull;
}

// LF(i, c) = counts[c] + rank(i, c) from the sector of block i>>6 (reference rank(),
// src/fms_index.h:68-86, and the `count + rank` of update_range, :100-102).
// w0 = cnt[0] | cnt[1] << 32, w1 = cnt[2] | cnt[3] << 32.
template <bool WIDE>
__device__ __forceinline__ typename PosT<WIDE>::type lf_map(const DevIndex &d, u64 w0, u64 w1, u64 lo, u64 hi,
                                                            typename PosT<WIDE>::type i, u32 c) {
    typedef typename PosT<WIDE>::type pos_t;
    const u64 w = (c & 2) ? w1 : w0;
    const u32 cnt = (c & 1) ? (u32)(w >> 32) : (u32)w;
    const u64 x = (c & 1) ? lo : ~lo;
    const u64 y = (c & 2) ? hi : ~hi;
    const u64 m = x & y & low_mask((u32)i & 63u);
    pos_t r = (pos_t)cnt + (pos_t)__popcll(m);
    if (WIDE) r += (pos_t)d.sb_base[(((u64)i >> 6) >> d.sb_shift) * 4 + c];
    // '$' is stored as A: ranks of A past the dollar slot are one too high (fms_index.h:81).
    r -= (pos_t)((c == 0) & ((u64)i > d.dollar));
    return r;
}

// ones among the first n bits of a 64-bit (32-bit) word, n any integer: 0 for n <= 0, all for n >= 64 (32)
__device__ __forceinline__ u32 popc_upto64(u64 w, int n) {
    if (n <= 0) return 0u;
    return (u32)__popcll(n >= 64 ? w : (w & ((1ull << n) - 1ull)));
}
__device__ __forceinline__ u32 popc_upto32(u32 w, int n) {
    if (n <= 0) return 0u;
    return (u32)__popc(n >= 32 ? w : (w & ((1u << n) - 1u)));
}

// m LF-steps at once from one multi-step sector (a = cnt | bits[0] << 32, b..d = bits[1..6]):
// cnt + #{set bits before offset i % 224}.
__device__ __forceinline__ u32 lf_multi(u64 a, u64 b, u64 c, u64 d, u32 off) {
    const int o = (int)off;
    return (u32)a + popc_upto32((u32)(a >> 32), o) + popc_upto64(b, o - 32) + popc_upto64(c, o - 96) + popc_upto64(d, o - 160);
}

// the wide sector: a = cnt (64 bits), b..d = 192 bits
__device__ __forceinline__ u64 lf_multi_wide(u64 a, u64 b, u64 c, u64 d, u32 off) {
    const int o = (int)off;
    return a + (u64)(popc_upto64(b, o) + popc_upto64(c, o - 64) + popc_upto64(d, o - 128));
}
template <bool WIDE>
__device__ __forceinline__ typename PosT<WIDE>::type lf_multi_t(u64 a, u64 b, u64 c, u64 d, u32 off) {
    if (WIDE) return (typename PosT<WIDE>::type)lf_multi_wide(a, b, c, d, off);
    return (typename PosT<WIDE>::type)lf_multi(a, b, c, d, off);
}
// block / offset of SA row i in a multi-step array
template <bool WIDE>
__device__ __forceinline__ u64 multi_block_of(typename PosT<WIDE>::type i) {
    if (WIDE) return (u64)i / MultiGeom<true>::rows;
    return (u64)((u32)i / MultiGeom<false>::rows);
}

// 2-bit reverse complement of a packed k-mer (first base in the highest used bits).
__device__ __forceinline__ u64 revcomp_packed(u64 x, u32 k) {
    x = __brevll(~x);
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    return x >> (64 - 2 * k);
}

// mask rank1 / bit helpers on an aux sector (klcp, mask, mask_cum, spare)
__device__ __forceinline__ u64 mask_rank_excl(u64 mask, u64 cum, u32 off) {  // ones in [64b, 64b+off)
    return cum + (u64)__popcll(mask & low_mask(off));
}
__device__ __forceinline__ u64 mask_rank_incl(u64 mask, u64 cum, u32 off) {  // ones in [64b, 64b+off]
    return cum + (u64)__popcll(mask & ((2ull << off) - 1ull));
}

}  // namespace fmsi
