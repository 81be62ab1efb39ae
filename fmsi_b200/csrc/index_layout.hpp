// index_layout.hpp — the GPU-resident layout of an FMS-index and the host-side converter that
// builds it from the reference's on-disk files (replaces load_index, reference
// src/fms_index.h:502-526; the state it rebuilds is struct fms_index, :52-66).
//
// Layout in HBM (all arrays cudaMalloc'ed, 256-byte aligned; element b covers BWT positions
// [64b, 64b+64)):
//
//   RankBlock rank[nblk]   32 B = one HBM/L2 sector = one LF-mapping probe:
//        cnt[c] : counts[c] + #{p < 64b : BWT[p] == c}  (the C-array is folded in; the '$' slot is
//                 stored as A exactly like the reference's `ac` vector, so LF(i,A) subtracts
//                 [i > dollar_position], fms_index.h:81); u32, relative to sb_base[b >> sb_shift][c]
//        lo, hi : bit p&63 of (hi,lo) is the 2-bit symbol at position p
//   AuxBlock  aux[nblk]    32 B = one sector, probed once per finished strand search and per
//                          kLCP extension:
//        klcp     : 64 bits of the kLCP vector (0 when the index is loaded without -S)
//        mask     : 64 bits of the SA-transformed mask (decoded from RRR<63> at load)
//        mask_cum : number of mask ones before position 64b   (rank1 for `lookup`)
//   nblk = floor(N/64)+1 so that rank(N, c), the first probe of every search, has a home
//   (N = n+1 = sa_transformed_mask.size()).
#pragma once
#include <cstdint>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "sdsl_io.hpp"

namespace fmsi {

struct alignas(32) RankBlock {
    uint32_t cnt[4];
    uint64_t lo;
    uint64_t hi;
};
struct alignas(32) AuxBlock {
    uint64_t klcp;
    uint64_t mask;
    uint64_t mask_cum;
    uint64_t spare;
};
static_assert(sizeof(RankBlock) == 32 && sizeof(AuxBlock) == 32, "one sector each");

struct HostIndex {
    uint64_t n = 0;  // BWT length N = n_superstring + 1
    int k = 0;
    bool has_klcp = false;
    uint64_t counts[4] = {0, 0, 0, 0};
    uint64_t dollar = 0;
    uint64_t mask_ones = 0;
    unsigned sb_shift = 63;         // superblock = 2^sb_shift blocks
    std::vector<uint64_t> sb_base;  // [n_superblocks][4]
    std::vector<RankBlock> rank;
    std::vector<AuxBlock> aux;
    bool wide() const { return sb_shift < 63; }
};

// Deposit the low popcount(sel) bits of `src` into the set positions of `sel` (software PDEP).
inline uint64_t deposit_bits(uint64_t src, uint64_t sel) {
    uint64_t out = 0;
    while (sel) {
        uint64_t low = sel & (0 - sel);
        if (src & 1) out |= low;
        src >>= 1;
        sel ^= low;
    }
    return out;
}

// Build the blocked layout from the wavelet-tree vectors, the plain mask and (optionally) kLCP.
// sb_shift_log2: superblock size in blocks (log2); 0 = choose (one superblock whenever N < 2^32).
inline HostIndex build_host_index(const BitVec &ac_gt, const BitVec &ac, const BitVec &gt, const BitVec &mask,
                                  const BitVec *klcp, const uint64_t counts[4], uint64_t dollar, int k,
                                  unsigned sb_shift_log2 = 0) {
    HostIndex h;
    h.n = ac_gt.nbits;
    h.k = k;
    h.dollar = dollar;
    for (int c = 0; c < 4; ++c) h.counts[c] = counts[c];
    if (h.n == 0) throw std::runtime_error("empty index");
    if (mask.nbits != h.n) throw std::runtime_error("mask length does not match the BWT length");
    if (klcp && klcp->nbits != 0 && klcp->nbits != h.n) throw std::runtime_error("kLCP length does not match the BWT length");
    if (dollar >= h.n) throw std::runtime_error("dollar_position out of range");
    h.has_klcp = klcp && klcp->nbits == h.n;

    const uint64_t nblk = (h.n >> 6) + 1;
    // u32 in-block counters hold values < 2^32: one superblock when counts fit, else 2^25 blocks
    // (2^31 positions) per superblock.
    if (sb_shift_log2) h.sb_shift = sb_shift_log2;
    else h.sb_shift = (h.n < (1ull << 32)) ? 63 : 25;
    const uint64_t nsb = h.sb_shift >= 63 ? 1 : ((nblk - 1) >> h.sb_shift) + 1;
    h.sb_base.assign(nsb * 4, 0);
    h.rank.resize(nblk);
    h.aux.resize(nblk);

    uint64_t occ[4] = {0, 0, 0, 0};
    uint64_t ac_pos = 0, gt_pos = 0, ones = 0;
    for (uint64_t b = 0; b < nblk; ++b) {
        const uint64_t p0 = b << 6;
        const unsigned valid = p0 >= h.n ? 0 : (unsigned)std::min<uint64_t>(64, h.n - p0);
        const uint64_t vmask = valid == 64 ? ~0ull : ((1ull << valid) - 1);
        const uint64_t g = valid ? (ac_gt.w[b] & vmask) : 0;  // 1 = G/T
        const uint64_t a = ~g & vmask;                          // 1 = A/C/$
        const unsigned ng = (unsigned)__builtin_popcountll(g), na = (unsigned)__builtin_popcountll(a);
        if (ac_pos + na > ac.nbits || gt_pos + ng > gt.nbits) throw std::runtime_error("ac/gt vectors shorter than ac_gt implies");
        const uint64_t acbits = na ? ac.get_int(ac_pos, na) : 0;
        const uint64_t gtbits = ng ? gt.get_int(gt_pos, ng) : 0;
        ac_pos += na;
        gt_pos += ng;
        const uint64_t lo = deposit_bits(acbits, a) | deposit_bits(gtbits, g);
        const uint64_t hi = g;

        const uint64_t sb = h.sb_shift >= 63 ? 0 : (b >> h.sb_shift);
        if (h.sb_shift < 63 && (b & ((1ull << h.sb_shift) - 1)) == 0)
            for (int c = 0; c < 4; ++c) h.sb_base[sb * 4 + c] = h.counts[c] + occ[c];
        RankBlock &rb = h.rank[b];
        for (int c = 0; c < 4; ++c) {
            uint64_t rel = h.counts[c] + occ[c] - h.sb_base[sb * 4 + c];
            if (rel >> 32) throw std::runtime_error("rank counter overflow (superblock too large)");
            rb.cnt[c] = (uint32_t)rel;
        }
        rb.lo = lo;
        rb.hi = hi;
        occ[0] += (unsigned)__builtin_popcountll(~hi & ~lo & vmask);
        occ[1] += (unsigned)__builtin_popcountll(~hi & lo & vmask);
        occ[2] += (unsigned)__builtin_popcountll(hi & ~lo & vmask);
        occ[3] += (unsigned)__builtin_popcountll(hi & lo & vmask);

        AuxBlock &ab = h.aux[b];
        const uint64_t m = valid ? (mask.w[b] & vmask) : 0;
        ab.mask = m;
        ab.mask_cum = ones;
        ab.klcp = (h.has_klcp && valid) ? (klcp->w[b] & vmask) : 0;
        ab.spare = 0;
        ones += (unsigned)__builtin_popcountll(m);
    }
    h.mask_ones = ones;
    if (ac_pos != ac.nbits || gt_pos != gt.nbits) throw std::runtime_error("ac/gt vector lengths inconsistent with ac_gt");
    // Self-check against .misc (construct(), fms_index.h:451: counts = {1, #A+1, #A+#C+1, #A+#C+#G+1};
    // occ[0] includes the '$' slot).
    if (h.counts[0] != 1 || h.counts[1] != occ[0] || h.counts[2] != occ[0] + occ[1] ||
        h.counts[3] != occ[0] + occ[1] + occ[2])
        throw std::runtime_error("counts in .misc do not match the BWT");
    return h;
}

struct IndexFiles {
    BitVec ac_gt, ac, gt, mask, klcp;
    uint64_t counts[4];
    uint64_t dollar;
    int k;
    bool klcp_present;
};

inline bool file_exists(const std::string &p) {
    std::ifstream f(p);
    return f.good();
}

// Parse <prefix>.fmsi.{ac_gt,ac,gt,mask,klcp,misc}. A missing/empty index yields mask.nbits == 0
// (the reference's only load check, main.cpp:304-307).
inline IndexFiles read_index_files(const std::string &prefix, bool use_klcp) {
    IndexFiles f;
    const std::string base = prefix + ".fmsi";
    for (const char *ext : {".ac_gt", ".ac", ".gt", ".mask", ".misc"})
        if (!file_exists(base + ext)) throw std::runtime_error("index not correctly loaded: missing " + base + ext);
    {
        ByteReader r(base + ".ac_gt");
        r.read_bitvec(f.ac_gt);
    }
    {
        ByteReader r(base + ".ac");
        r.read_bitvec(f.ac);
    }
    {
        ByteReader r(base + ".gt");
        r.read_bitvec(f.gt);
    }
    f.mask = rrr_decode_all(read_rrr(base + ".mask"));
    f.klcp_present = false;
    if (use_klcp && file_exists(base + ".klcp")) {
        ByteReader r(base + ".klcp");
        r.read_bitvec(f.klcp);
        f.klcp_present = f.klcp.nbits > 0;
    }
    std::ifstream in(base + ".misc");
    if (!(in >> f.dollar >> f.counts[0] >> f.counts[1] >> f.counts[2] >> f.counts[3] >> f.k))
        throw std::runtime_error("malformed " + base + ".misc");
    return f;
}

}  // namespace fmsi
