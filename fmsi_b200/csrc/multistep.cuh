// multistep.cuh — multi-step rank arrays: m LF-steps of the backward search for one memory request.
//
// Why. The backward search of the reference (get_range_with_pattern, src/fms_index.h:117-124) is a chain of
// k dependent update_range steps (:98-103), each a rank on the BWT. On B200 a dependent random read costs one
// DRAM request (~37-44 G/s per GPU whatever it returns, tools/randbw2.cu), so the kernel's cost is its number
// of probes: after the suffix table, k - t of them per strand search (12.8 sectors per k-mer on the human-
// scale bench mix, profiles/r01b_human_query_kmers_metrics.txt). m steps at once need, for the m-mer x they
// prepend,  LF_m(i, x) = C_m[x] + #{rows r < i whose suffix is preceded by x}  — the same count the m single
// steps compose (each step counts the suffixes smaller than c . boundary). So for every m-mer x (4^m of
// them) this tier keeps a plain bitvector over the SA rows ("row r is preceded by x") cut into 32-byte
// sectors of one 32-bit absolute counter + 224 bits (device_index.cuh: MultiBlock). One probe then advances
// m bases, and the two interval ends share the sector unless they straddle a 224-row boundary.
//   x = c_1 | c_2 << 2 [| c_3 << 4]: c_1 = BWT[r] is the base right before the suffix, c_2 = BWT[LF(r)] the
//   one before that — the order in which the kernels consume a packed pattern (lowest two bits first).
//   Rows whose suffix starts less than m characters into the text are preceded by no m-mer (the '$' slot is
//   stored as A in the planes, as in the reference's `ac` vector, so it is tested by position).
// Size: 4^m * 32 / 224 bytes per row: m = 2 -> 2.29 B (7.1 GB at 3.1 Gbp), m = 3 -> 9.14 B (28 GB).
// Indexes of 2^32 rows and more take the same arrays with a 64-bit counter and 192 rows per sector
// (MultiBlockWide: 2.67 B per row at m = 2), so no superblock base has to be added to a probe.
// Built on the device from the rank blocks alone, so file-loaded and device-built indexes share it.
#pragma once
#include <type_traits>

#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "dict.cuh"
#include "index_build.cuh"

namespace fmsi {

constexpr unsigned char kMultiNoCode = 0xFF;

// code[r] = the m-mer preceding the suffix of row r (kMultiNoCode when fewer than m characters precede it)
template <bool WIDE>
__global__ void multi_codes_kernel(const DevIndex d, const u32 m, unsigned char *__restrict__ code) {
    typedef typename PosT<WIDE>::type pos_t;
    const u64 r = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (r >= d.n) return;
    pos_t cur = (pos_t)r;
    u32 x = 0;
    bool ok = true;
    for (u32 s = 0; s < m; ++s) {
        if ((u64)cur == d.dollar) {
            ok = false;
            break;
        }
        u64 a0, a1, a2, a3;
        ld_sector_l1(d.rank + ((u64)cur >> 6), a0, a1, a2, a3);
        const u32 c = block_symbol(a2, a3, (u32)cur & 63u);
        x |= c << (2 * s);
        cur = lf_map<WIDE>(d, a0, a1, a2, a3, cur, c);
    }
    code[r] = ok ? (unsigned char)x : kMultiNoCode;
}

// One warp per block of 224 (wide: 192) rows: the 4^m sectors of the block, counters = ones inside the block for now.
// A sector is 8 words: HDR counter words (1, wide: 2), then the bit words.
template <bool WIDE>
__global__ void multi_bits_kernel(const unsigned char *__restrict__ code, const u64 n, const u32 m, const u32 nblk,
                                  MultiBlock *__restrict__ multi) {
    constexpr u32 HDR = MultiGeom<WIDE>::hdr, W = 8 - HDR;
    const unsigned FULL = 0xffffffffu;
    const u32 lane = threadIdx.x & 31u;
    const u64 b = (blockIdx.x * (u64)blockDim.x + threadIdx.x) >> 5;
    if (b >= nblk) return;  // uniform per warp
    u32 cd[W];
#pragma unroll
    for (u32 w = 0; w < W; ++w) {
        const u64 r = b * MultiGeom<WIDE>::rows + 32u * w + lane;
        cd[w] = r < n ? (u32)code[r] : (u32)kMultiNoCode;
    }
    const u32 nx = 1u << (2 * m);
    for (u32 x0 = 0; x0 < nx; x0 += 32) {
        u32 mine[W];  // lane l keeps the sector of x0 + l
#pragma unroll
        for (u32 w = 0; w < W; ++w) mine[w] = 0;
        const u32 lim = nx - x0 < 32u ? nx - x0 : 32u;
        for (u32 xl = 0; xl < lim; ++xl) {
#pragma unroll
            for (u32 w = 0; w < W; ++w) {
                const u32 bw = __ballot_sync(FULL, cd[w] == x0 + xl);
                if (lane == xl) mine[w] = bw;
            }
        }
        if (lane < lim) {
            u32 v[8];
            u32 cnt = 0;
#pragma unroll
            for (u32 w = 0; w < W; ++w) {
                v[HDR + w] = mine[w];
                cnt += (u32)__popc(mine[w]);
            }
            v[0] = cnt;
            if (HDR == 2) v[1] = 0;
            uint4 *dst = reinterpret_cast<uint4 *>(multi + ((u64)(x0 + lane) * nblk + b));
            dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
            dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
        }
    }
}

// C_m[x] = #suffixes smaller than x = where m LF-steps take position 0 (the `i` of the depth-m suffix table)
template <bool WIDE>
__global__ void multi_cm_kernel(const DevIndex d, const u32 m, u64 *__restrict__ cm) {
    const u32 x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= (1u << (2 * m))) return;
    u64 i = 0;
    for (u32 s = 0; s < m; ++s) i = (u64)dev_lf<WIDE>(d, i, (x >> (2 * s)) & 3u);
    cm[x] = i;
}

template <typename T>
struct MultiCntAt {  // transform: sector index -> its in-block count (the low counter word of either layout)
    const MultiBlock *multi;
    __host__ __device__ __forceinline__ T operator()(const u64 s) const { return (T)multi[s].cnt; }
};

// counters: in-block ones -> C_m[x] + ones in the earlier blocks of the same x
template <bool WIDE, typename T>
__global__ void multi_counters_kernel(const T *__restrict__ ex, const u64 *__restrict__ cm, const u32 nblk, const u64 total,
                                      MultiBlock *__restrict__ multi) {
    const u64 s = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (s >= total) return;
    const u64 x = s / nblk;
    const u64 v = cm[x] + (u64)(ex[s] - ex[x * nblk]);
    if (WIDE) reinterpret_cast<MultiBlockWide *>(multi)[s].cnt = v;
    else multi[s].cnt = (u32)v;
}

inline u64 multi_rows_per_sector(bool wide) { return wide ? kMultiRowsWide : kMultiRows; }
inline u64 multi_bytes(u64 N, u32 m, bool wide = false) { return ((u64)1 << (2 * m)) * (N / multi_rows_per_sector(wide) + 1) * sizeof(MultiBlock); }
// transient device memory of build_multi_on_device beyond the arrays themselves
inline u64 multi_build_scratch_bytes(u64 N, u32 m, bool wide = false) {
    return N + ((u64)(wide ? 8 : 4) << (2 * m)) * (N / multi_rows_per_sector(wide) + 1) + (64ull << 20);
}

// Throws std::runtime_error; nothing is leaked then. WIDE: the index's layout (u64 positions, 64-bit counters).
template <bool WIDE>
inline void build_multi_on_device_t(const DevIndex &d, u32 m, MultiBlock **out, u32 *out_nblk, uint64_t *launches) {
    typedef typename std::conditional<WIDE, u64, u32>::type cnt_t;
    const u64 N = d.n;
    const u32 nblk = (u32)(N / MultiGeom<WIDE>::rows + 1);
    const u64 total = ((u64)1 << (2 * m)) * nblk;
    DevArr<MultiBlock> multi(total);
    {
        DevArr<unsigned char> code(N);
        multi_codes_kernel<WIDE><<<nblocks_for(N), 256>>>(d, m, code.p);
        BCU(cudaGetLastError());
        multi_bits_kernel<WIDE><<<nblocks_for((u64)nblk * 32), 256>>>(code.p, N, m, nblk, multi.p);
        BCU(cudaGetLastError());
        BCU(cudaDeviceSynchronize());
    }
    DevArr<u64> cm((size_t)1 << (2 * m));
    multi_cm_kernel<WIDE><<<1, 64>>>(d, m, cm.p);
    BCU(cudaGetLastError());
    DevArr<cnt_t> ex(total);
    {
        auto in = thrust::make_transform_iterator(thrust::counting_iterator<u64>(0), MultiCntAt<cnt_t>{multi.p});
        size_t tmp_bytes = 0;
        BCU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, ex.p, total));
        DevArr<unsigned char> tmp(tmp_bytes);
        BCU(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, in, ex.p, total));
    }
    multi_counters_kernel<WIDE, cnt_t><<<nblocks_for(total), 256>>>(ex.p, cm.p, nblk, total, multi.p);
    BCU(cudaGetLastError());
    BCU(cudaDeviceSynchronize());
    if (launches) *launches += 5;
    *out = multi.p;
    *out_nblk = nblk;
    multi.p = nullptr;
}
inline void build_multi_on_device(const DevIndex &d, bool wide, u32 m, MultiBlock **out, u32 *out_nblk, uint64_t *launches) {
    if (wide) build_multi_on_device_t<true>(d, m, out, out_nblk, launches);
    else build_multi_on_device_t<false>(d, m, out, out_nblk, launches);
}

}  // namespace fmsi
