// index_layout.hpp — the GPU-resident layout of an FMS-index (the state of the reference's
// struct fms_index, src/fms_index.h:52-66). It is produced on the device: from the reference's
// on-disk files by index_convert.cuh (replaces load_index, :502-526) or from a masked superstring by
// index_build.cuh (replaces construct, :397-460).
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
    std::vector<uint64_t> sb_base;  // [n_superblocks][4] (device-built narrow indexes: one zero row)
};

inline bool file_exists(const std::string &p) {
    std::ifstream f(p);
    return f.good();
}

}  // namespace fmsi
