// index_convert.cuh — device-side conversion of the reference's index files into the GPU layout.
//
// Replaces what load_index (reference src/fms_index.h:502-526) does after reading the files —
// rebuilding rank supports in RAM — with: upload the raw sdsl words, then on the device
//   * decode the RRR<63> mask (class + offset in the combinatorial number system, per-superblock
//     `invert`, offset-stream samples every 32 blocks; sdsl rrr_vector.hpp:257-277, rrr_helper.hpp)
//     straight into plain 64-bit mask words, one thread per output word;
//   * re-interleave the wavelet-tree vectors ac_gt / ac / gt (fms_index.h:52-58) into the 2-bit
//     (hi, lo) planes: prefix sums of the ac_gt popcounts give every block its cursor into ac and gt,
//     a software bit-deposit places the bits;
//   * prefix-sum the per-block symbol and mask counts and assemble RankBlock / AuxBlock sectors
//     (index_layout.hpp), with 64-bit superblock bases when N >= 2^32.
// Host work is reading the files; a human-scale index converts in well under a second of GPU time.
#pragma once
#include "index_build.cuh"
#include "index_layout.hpp"

namespace fmsi {

__device__ __forceinline__ u64 dev_get_int(const u64 *__restrict__ w, u64 pos, u32 len) {  // len <= 64, LSB first
    if (len == 0) return 0;
    const u64 wi = pos >> 6;
    const u32 off = (u32)pos & 63u;
    u64 x = w[wi] >> off;
    if (off + len > 64) x |= w[wi + 1] << (64 - off);
    if (len < 64) x &= (1ull << len) - 1ull;
    return x;
}

__device__ __forceinline__ u64 dev_deposit(u64 src, u64 sel) {  // software PDEP
    u64 out = 0;
    while (sel) {
        const u64 low = sel & (0 - sel);
        if (src & 1ull) out |= low;
        src >>= 1;
        sel ^= low;
    }
    return out;
}

// ---- RRR<63> -------------------------------------------------------------------------------------
struct RrrDev {
    const u64 *bt;      // 6-bit classes, packed
    const u64 *btnr;    // offset stream
    const u64 *btnrp;   // stream position samples, `wp` bits each
    const u64 *invert;  // one bit per superblock
    u32 wp;
    u64 size;           // bits
};

// binom[n * 64 + k] = C(n, k), n, k < 64; space[k] = bits of the offset field of class k
__global__ void rrr_decode_kernel(const RrrDev r, const u64 *__restrict__ binom_g, const u32 *__restrict__ space_g,
                                  u64 *__restrict__ mask_words, const u64 n_words, u32 *__restrict__ ones_per_word) {
    __shared__ u64 binom[64 * 64];
    __shared__ u32 space[64];
    for (u32 t = threadIdx.x; t < 64 * 64; t += blockDim.x) binom[t] = binom_g[t];
    if (threadIdx.x < 64) space[threadIdx.x] = space_g[threadIdx.x];
    __syncthreads();
    const u64 w = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const u64 p0 = w * 64;
    u64 out = 0;
    if (p0 < r.size) {
        const u64 p1 = (p0 + 64 < r.size) ? p0 + 64 : r.size;  // exclusive
        for (u64 b = p0 / 63; b * 63 < p1; ++b) {
            const u64 sb = b >> 5;
            u64 off = dev_get_int(r.btnrp, sb * r.wp, r.wp);
            for (u64 e = sb << 5; e < b; ++e) off += space[(u32)dev_get_int(r.bt, e * 6, 6)];
            const u32 stored = (u32)dev_get_int(r.bt, b * 6, 6);
            const bool inv = (r.invert[sb >> 6] >> (sb & 63)) & 1ull;
            u32 ones = inv ? 63u - stored : stored;
            u64 nr = dev_get_int(r.btnr, off, space[stored]);
            u64 word = 0;
            if (ones == 63) word = (1ull << 63) - 1ull;
            else {
                u32 left = 63;
                for (u32 pos = 0; pos < 63 && ones > 0; ++pos, --left) {
                    const u64 below = binom[(left - 1) * 64 + ones];
                    if (nr >= below) {
                        word |= 1ull << pos;
                        nr -= below;
                        --ones;
                    }
                }
            }
            // block bits [63b, 63b+63) -> output bits [p0, p1)
            const u64 bb = b * 63;
            if (bb >= p0) out |= word << (bb - p0);
            else out |= word >> (p0 - bb);
        }
        const u32 valid = (u32)(p1 - p0);
        if (valid < 64) out &= (1ull << valid) - 1ull;
    }
    mask_words[w] = out;
    ones_per_word[w] = (u32)__popcll(out);
}

// ---- wavelet tree -> planes ------------------------------------------------------------------------
__global__ void acgt_counts_kernel(const u64 *__restrict__ acgt, const u64 N, const u64 nblk, u32 *__restrict__ n_ac, u32 *__restrict__ n_gt) {
    const u64 b = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const u64 p0 = b << 6;
    const u32 valid = p0 >= N ? 0u : (u32)((N - p0 < 64) ? (N - p0) : 64);
    const u64 vmask = valid == 64 ? ~0ull : ((1ull << valid) - 1ull);
    const u64 g = valid ? (acgt[b] & vmask) : 0ull;
    n_gt[b] = (u32)__popcll(g);
    n_ac[b] = (u32)__popcll(~g & vmask);
}

__global__ void planes_from_wt_kernel(const u64 *__restrict__ acgt, const u64 *__restrict__ ac, const u64 *__restrict__ gt,
                                      const u64 *__restrict__ mask_w, const u64 *__restrict__ ex_ac, const u64 *__restrict__ ex_gt,
                                      const u64 N, const u64 nblk, u64 *__restrict__ lo_w, u64 *__restrict__ hi_w,
                                      u32 *__restrict__ cA, u32 *__restrict__ cC, u32 *__restrict__ cG, u32 *__restrict__ cM) {
    const u64 b = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const u64 p0 = b << 6;
    const u32 valid = p0 >= N ? 0u : (u32)((N - p0 < 64) ? (N - p0) : 64);
    const u64 vmask = valid == 64 ? ~0ull : ((1ull << valid) - 1ull);
    const u64 g = valid ? (acgt[b] & vmask) : 0ull;
    const u64 a = ~g & vmask;
    const u64 acbits = dev_get_int(ac, ex_ac[b], (u32)__popcll(a));
    const u64 gtbits = dev_get_int(gt, ex_gt[b], (u32)__popcll(g));
    const u64 lo = dev_deposit(acbits, a) | dev_deposit(gtbits, g);
    lo_w[b] = lo;
    hi_w[b] = g;
    cA[b] = (u32)__popcll(~g & ~lo & vmask);
    cC[b] = (u32)__popcll(~g & lo & vmask);
    cG[b] = (u32)__popcll(g & ~lo & vmask);
    cM[b] = valid ? (u32)__popcll(mask_w[b] & vmask) : 0u;
}

// Superblock bases (absolute LF counts at the first block of every superblock), wide layout only.
__global__ void sb_base_kernel(const u64 *__restrict__ exA, const u64 *__restrict__ exC, const u64 *__restrict__ exG, const u64 nsb,
                               const u32 sb_shift, const u64 c0, const u64 c1, const u64 c2, const u64 c3, u64 *__restrict__ sb_base) {
    const u64 s = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (s >= nsb) return;
    const u64 b = s << sb_shift;
    const u64 a = exA[b], c = exC[b], g = exG[b], t = b * 64 - a - c - g;
    sb_base[s * 4 + 0] = c0 + a;
    sb_base[s * 4 + 1] = c1 + c;
    sb_base[s * 4 + 2] = c2 + g;
    sb_base[s * 4 + 3] = c3 + t;
}

__global__ void assemble_blocks_sb_kernel(const u64 *__restrict__ lo_w, const u64 *__restrict__ hi_w, const u64 *__restrict__ mask_w,
                                          const u64 *__restrict__ klcp_w, const u64 *__restrict__ exA, const u64 *__restrict__ exC,
                                          const u64 *__restrict__ exG, const u64 *__restrict__ exM, const u64 N, const u64 nblk, const u64 c0,
                                          const u64 c1, const u64 c2, const u64 c3, const u64 *__restrict__ sb_base, const u32 sb_shift,
                                          RankBlock *__restrict__ rank, AuxBlock *__restrict__ aux, int *__restrict__ overflow) {
    const u64 b = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const u64 p0 = b << 6;
    const u32 valid = p0 >= N ? 0u : (u32)((N - p0 < 64) ? (N - p0) : 64);
    const u64 vmask = valid == 64 ? ~0ull : ((1ull << valid) - 1ull);
    const u64 a = exA[b], c = exC[b], g = exG[b];
    // every earlier slot holds one of the four codes; slots past N hold none
    const u64 before = p0 < N ? p0 : N;
    const u64 t = before - a - c - g;
    u64 abs4[4] = {c0 + a, c1 + c, c2 + g, c3 + t};
    RankBlock rb;
    for (int s = 0; s < 4; ++s) {
        const u64 base = sb_shift >= 63 ? 0ull : sb_base[(b >> sb_shift) * 4 + s];
        const u64 rel = abs4[s] - base;
        if (rel >> 32) *overflow = 1;
        rb.cnt[s] = (uint32_t)rel;
    }
    rb.lo = lo_w[b];
    rb.hi = hi_w[b];
    rank[b] = rb;
    AuxBlock ab;
    ab.klcp = (klcp_w && valid) ? (klcp_w[b] & vmask) : 0ull;
    ab.mask = valid ? (mask_w[b] & vmask) : 0ull;
    ab.mask_cum = exM[b];
    ab.spare = 0;
    aux[b] = ab;
}

struct ConvertedIndex {
    DevArr<RankBlock> rank;
    DevArr<AuxBlock> aux;
    DevArr<u64> sb_base;       // [nsb * 4]
    u64 nsb = 1;
    unsigned sb_shift = 63;
    u64 mask_ones = 0;
    u64 occ[3] = {0, 0, 0};    // #A (incl. the '$' slot), #C, #G
};

// Upload helper: host words -> device array with `pad` zeroed words behind (get_int may touch one).
inline void upload_words(DevArr<u64> &d, const u64 *src, u64 n_words, u64 alloc_words) {
    d.alloc(alloc_words);
    BCU(cudaMemset(d.p, 0, alloc_words * 8));
    if (n_words) BCU(cudaMemcpy(d.p, src, n_words * 8, cudaMemcpyHostToDevice));
}

// d_mask: plain mask words (nblk), d_klcp may be null. sb_shift_log2: 0 = auto.
inline void convert_on_device(const u64 N, const DevArr<u64> &d_acgt, const DevArr<u64> &d_ac, const u64 n_ac_bits, const DevArr<u64> &d_gt,
                              const u64 n_gt_bits, const u64 *d_mask, const u64 *d_klcp, const u64 counts[4], unsigned sb_shift_log2,
                              ConvertedIndex &out, uint64_t *launches) {
    const u64 nblk = (N >> 6) + 1;
    out.sb_shift = sb_shift_log2 ? sb_shift_log2 : ((N < (1ull << 32)) ? 63 : 25);
    out.nsb = out.sb_shift >= 63 ? 1 : ((nblk - 1) >> out.sb_shift) + 1;
    DevArr<u32> n_ac(nblk), n_gt(nblk);
    acgt_counts_kernel<<<nblocks_for(nblk), 256>>>(d_acgt.p, N, nblk, n_ac.p, n_gt.p);
    DevArr<u64> ex_ac(nblk + 1), ex_gt(nblk + 1);
    exclusive_sum_u64(n_ac.p, ex_ac.p, nblk);
    exclusive_sum_u64(n_gt.p, ex_gt.p, nblk);
    auto total = [&](DevArr<u64> &ex, DevArr<u32> &cnt) {
        u64 e = 0;
        u32 c = 0;
        BCU(cudaMemcpy(&e, ex.p + nblk - 1, 8, cudaMemcpyDeviceToHost));
        BCU(cudaMemcpy(&c, cnt.p + nblk - 1, 4, cudaMemcpyDeviceToHost));
        return e + c;
    };
    if (total(ex_ac, n_ac) != n_ac_bits || total(ex_gt, n_gt) != n_gt_bits)
        throw std::runtime_error("ac/gt vector lengths inconsistent with ac_gt");
    DevArr<u64> lo(nblk), hi(nblk);
    DevArr<u32> cA(nblk), cC(nblk), cG(nblk), cM(nblk);
    planes_from_wt_kernel<<<nblocks_for(nblk), 256>>>(d_acgt.p, d_ac.p, d_gt.p, d_mask, ex_ac.p, ex_gt.p, N, nblk, lo.p, hi.p, cA.p, cC.p, cG.p, cM.p);
    *launches += 2;
    n_ac.release();
    n_gt.release();
    ex_ac.release();
    ex_gt.release();
    DevArr<u64> exA(nblk + 1), exC(nblk + 1), exG(nblk + 1), exM(nblk + 1);
    exclusive_sum_u64(cA.p, exA.p, nblk);
    exclusive_sum_u64(cC.p, exC.p, nblk);
    exclusive_sum_u64(cG.p, exG.p, nblk);
    exclusive_sum_u64(cM.p, exM.p, nblk);
    out.occ[0] = total(exA, cA);
    out.occ[1] = total(exC, cC);
    out.occ[2] = total(exG, cG);
    out.mask_ones = total(exM, cM);
    // Self-check against .misc (construct(), fms_index.h:451: counts = {1, #A+1, #A+#C+1, #A+#C+#G+1};
    // occ[0] includes the '$' slot).
    if (counts[0] != 1 || counts[1] != out.occ[0] || counts[2] != out.occ[0] + out.occ[1] || counts[3] != out.occ[0] + out.occ[1] + out.occ[2])
        throw std::runtime_error("counts in .misc do not match the BWT");
    out.sb_base.alloc(out.nsb * 4);
    BCU(cudaMemset(out.sb_base.p, 0, out.nsb * 32));
    if (out.sb_shift < 63) {
        sb_base_kernel<<<nblocks_for(out.nsb), 256>>>(exA.p, exC.p, exG.p, out.nsb, out.sb_shift, counts[0], counts[1], counts[2], counts[3], out.sb_base.p);
        *launches += 1;
    }
    out.rank.alloc(nblk);
    out.aux.alloc(nblk);
    DevArr<int> ovf(1);
    BCU(cudaMemset(ovf.p, 0, 4));
    assemble_blocks_sb_kernel<<<nblocks_for(nblk), 256>>>(lo.p, hi.p, d_mask, d_klcp, exA.p, exC.p, exG.p, exM.p, N, nblk, counts[0], counts[1],
                                                          counts[2], counts[3], out.sb_base.p, out.sb_shift >= 63 ? 63u : out.sb_shift, out.rank.p,
                                                          out.aux.p, ovf.p);
    *launches += 1;
    int h_ovf = 0;
    BCU(cudaMemcpy(&h_ovf, ovf.p, 4, cudaMemcpyDeviceToHost));
    if (h_ovf) throw std::runtime_error("rank counter overflow (superblock too large)");
}

// Decode an RRR<63> file image (already parsed into RrrFile on the host) into plain mask words on
// the device. Verifies the total against the last stored rank sample.
inline void rrr_decode_on_device(const RrrFile &f, const u64 nblk, DevArr<u64> &d_mask, uint64_t *launches) {
    const Binomials &B = Binomials::get();
    std::vector<u64> binom(64 * 64, 0);
    std::vector<u32> space(64, 0);
    for (int n = 0; n < 64; ++n)
        for (int k = 0; k < 64; ++k) binom[n * 64 + k] = B.c[n][k];
    for (int k = 0; k < 64; ++k) space[k] = B.space[k];
    DevArr<u64> d_binom(64 * 64);
    DevArr<u32> d_space(64);
    BCU(cudaMemcpy(d_binom.p, binom.data(), binom.size() * 8, cudaMemcpyHostToDevice));
    BCU(cudaMemcpy(d_space.p, space.data(), 64 * 4, cudaMemcpyHostToDevice));
    const u64 nblocks = (f.size + kRrrBlock) / kRrrBlock;
    if (f.bt.n < nblocks) throw std::runtime_error("rrr: bt array too short");
    const u64 nsuper = (nblocks + kRrrSample - 1) / kRrrSample;
    if (f.btnrp.n + 1 < nsuper || f.invert.nbits + 1 < nsuper) throw std::runtime_error("rrr: sample arrays too short");
    DevArr<u64> d_bt, d_btnr, d_btnrp, d_inv;
    upload_words(d_bt, f.bt.bits.w.data(), f.bt.bits.w.size(), f.bt.bits.w.size() + 2);
    upload_words(d_btnr, f.btnr.w.data(), f.btnr.w.size(), f.btnr.w.size() + 2);
    upload_words(d_btnrp, f.btnrp.bits.w.data(), f.btnrp.bits.w.size(), f.btnrp.bits.w.size() + 2);
    upload_words(d_inv, f.invert.w.data(), f.invert.w.size(), f.invert.w.size() + 2);
    RrrDev r;
    r.bt = d_bt.p;
    r.btnr = d_btnr.p;
    r.btnrp = d_btnrp.p;
    r.invert = d_inv.p;
    r.wp = f.btnrp.width;
    r.size = f.size;
    d_mask.alloc(nblk);
    DevArr<u32> ones(nblk);
    rrr_decode_kernel<<<nblocks_for(nblk), 256>>>(r, d_binom.p, d_space.p, d_mask.p, nblk, ones.p);
    *launches += 1;
    DevArr<u64> ex(nblk + 1);
    exclusive_sum_u64(ones.p, ex.p, nblk);
    u64 e = 0;
    u32 c = 0;
    BCU(cudaMemcpy(&e, ex.p + nblk - 1, 8, cudaMemcpyDeviceToHost));
    BCU(cudaMemcpy(&c, ones.p + nblk - 1, 4, cudaMemcpyDeviceToHost));
    if (f.rank.n == 0 || f.rank.get(f.rank.n - 1) != e + c) throw std::runtime_error("rrr: total mismatch");
}

}  // namespace fmsi
