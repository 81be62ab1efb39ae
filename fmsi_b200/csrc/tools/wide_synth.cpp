// wide_synth.cpp — test tool (not part of the product library): writes the six `.fmsi.*` files of a SYNTHETIC index with any
// number of rows, in particular more than 2^32 — the size at which the device layout switches to 64-bit positions and
// per-superblock counter bases, and which neither the GPU builder (n + 1 < 2^32) nor the reference's QSufSort (hours,
// 16 B per character) can produce in test time. The "BWT" is a seeded pseudo-random symbol sequence with one '$' slot; it
// need not be the transform of any text: rank, update_range, get_range_with_pattern, infer_presence, kmer_order and the
// kLCP extension are functions of these bit vectors alone, so the oracle (the reference's algorithm) and the GPU
// kernels must agree on them row for row. Patterns that occur are obtained by walking the LF-mapping
// (tests/test_gpu_wide.py does that with the oracle).
//   wide_synth <prefix> <n_rows> <k> <seed>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "../sdsl_io.hpp"

static inline uint64_t mix(uint64_t x) {  // splitmix64
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

int main(int argc, char **argv) {
    if (argc != 5) {
        std::fprintf(stderr, "usage: wide_synth <prefix> <n_rows> <k> <seed>\n");
        return 2;
    }
    try {
        const std::string base = std::string(argv[1]) + ".fmsi";
        const uint64_t N = std::strtoull(argv[2], nullptr, 10), seed = std::strtoull(argv[4], nullptr, 10);
        const int k = std::atoi(argv[3]);
        const uint64_t nw = (N + 63) >> 6;
        const uint64_t dollar = N / 3 + 17;
        fmsi::BitVec ac_gt, lo, mask, klcp;
        ac_gt.resize_bits(N);
        lo.resize_bits(N);
        mask.resize_bits(N);
        klcp.resize_bits(N);
        fmsi::parallel_ranges(nw, 1, [&](uint64_t w0, uint64_t w1) {
            for (uint64_t w = w0; w < w1; ++w) {
                const uint64_t valid = N - w * 64 >= 64 ? ~0ull : ((1ull << (N - w * 64)) - 1);
                ac_gt.w[w] = mix(seed + 4 * w) & valid;
                lo.w[w] = mix(seed + 4 * w + 1) & valid;
                // 63/64 ones: the lookup ids (mask ranks) of the last rows then exceed 2^32 as well
                mask.w[w] = ~(mix(seed + 4 * w + 2) & mix(seed * 7 + w) & mix(seed * 13 + w) & mix(seed * 17 + w) & mix(seed * 19 + w) & mix(seed * 23 + w)) & valid;
                klcp.w[w] = (mix(seed + 4 * w + 3) & mix(seed * 11 + w)) & valid;               // ~1/4 ones
            }
        });
        ac_gt.w[dollar >> 6] &= ~(1ull << (dollar & 63));  // the '$' slot is stored as A (fms_index.h:81)
        lo.w[dollar >> 6] &= ~(1ull << (dollar & 63));
        klcp.w[0] &= ~1ull;                                   // extend_range_with_klcp's scans end (fms_index.h:106-109)
        klcp.w[(N - 1) >> 6] &= ~(1ull << ((N - 1) & 63));
        // ac = the low bits of the A/C/$ slots in row order, gt = those of the G/T slots (fms_index.h:434-451)
        uint64_t n_gt = 0, n_c = 0, n_t = 0;
        for (uint64_t w = 0; w < nw; ++w) {
            n_gt += (uint64_t)__builtin_popcountll(ac_gt.w[w]);
            n_c += (uint64_t)__builtin_popcountll(~ac_gt.w[w] & lo.w[w]);
            n_t += (uint64_t)__builtin_popcountll(ac_gt.w[w] & lo.w[w]);
        }
        const uint64_t n_ac = N - n_gt, n_a = n_ac - n_c, n_g = n_gt - n_t;  // n_a counts the '$' slot
        fmsi::BitVec ac, gt;
        ac.resize_bits(n_ac);
        gt.resize_bits(n_gt);
        uint64_t ap = 0, gp = 0;
        for (uint64_t w = 0; w < nw; ++w) {
            const uint64_t valid = N - w * 64 >= 64 ? ~0ull : ((1ull << (N - w * 64)) - 1);
            for (int pass = 0; pass < 2; ++pass) {
                uint64_t sel = (pass ? ac_gt.w[w] : ~ac_gt.w[w]) & valid, bits = 0;
                unsigned cnt = 0;
                while (sel) {
                    bits |= ((lo.w[w] >> __builtin_ctzll(sel)) & 1ull) << cnt++;
                    sel &= sel - 1;
                }
                if (pass) {
                    gt.set_int(gp, bits, cnt);
                    gp += cnt;
                } else {
                    ac.set_int(ap, bits, cnt);
                    ap += cnt;
                }
            }
        }
        if (ap != n_ac || gp != n_gt) throw std::runtime_error("slot counts inconsistent");
        {
            fmsi::ByteWriter w(base + ".ac_gt");
            w.bitvec(ac_gt);
        }
        {
            fmsi::ByteWriter w(base + ".ac");
            w.bitvec(ac);
        }
        {
            fmsi::ByteWriter w(base + ".gt");
            w.bitvec(gt);
        }
        {
            fmsi::ByteWriter w(base + ".klcp");
            w.bitvec(klcp);
        }
        fmsi::write_rrr(base + ".mask", fmsi::rrr_encode(mask));
        FILE *f = std::fopen((base + ".misc").c_str(), "w");
        if (!f) throw std::runtime_error("cannot create " + base + ".misc");
        // counts = {1, #A+1, #A+#C+1, #A+#C+#G+1} with the '$' slot inside #A's slots (construct(), fms_index.h:451)
        std::fprintf(f, "%llu\n%llu\n%llu\n%llu\n%llu\n%d\n", (unsigned long long)dollar, 1ull, (unsigned long long)n_a, (unsigned long long)(n_a + n_c),
                     (unsigned long long)(n_a + n_c + n_g), k);
        std::fclose(f);
        std::printf("%llu %llu\n", (unsigned long long)N, (unsigned long long)dollar);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
