// loc_key_check.cu — host-only check of the key function of the minimizer-bucketed dictionary (loc.cuh); test tool
// (tests/test_loc_key.py). For every k-mer of a small alphabet space (exhaustive) or a seeded random sample:
//   * q and its reverse complement get the same (bucket, R), with opposite `sw` unless q is its own reverse complement;
//   * distinct strand pairs {q, rc} get distinct (bucket, R)  (exhaustive runs only);
//   * bucket < 4^t, R < 2^rbits, and the pick is a minimum: no m-mer of either strand orders below it.
// usage: loc_key_check k m t [samples seed]   -> prints one JSON line, exit code 1 on any violation
#include <cstdio>
#include <cstdlib>
#include <map>
#include <utility>

#include "../loc.cuh"

using namespace fmsi;

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    const u32 k = (u32)atoi(argv[1]), m = (u32)atoi(argv[2]), t = (u32)atoi(argv[3]);
    const u64 samples = argc > 4 ? strtoull(argv[4], 0, 10) : 0;
    u64 seed = argc > 5 ? strtoull(argv[5], 0, 10) : 1;
    if (!loc_fits(k, m, t)) {
        printf("{\"k\": %u, \"m\": %u, \"t\": %u, \"fits\": false}\n", k, m, t);
        return 0;
    }
    const LocGeom g = loc_geom(k, m, t);
    const u64 kmask = k < 32 ? (1ull << (2 * k)) - 1ull : ~0ull;
    const u32 mmask = m < 16 ? (1u << (2 * m)) - 1u : 0xFFFFFFFFu;
    std::map<std::pair<u32, u64>, u64> seen;  // (bucket, R) -> the smaller k-mer of the pair
    u64 bad = 0, n = 0, self = 0;
    auto next = [&]() {
        seed += 0x9E3779B97F4A7C15ull;
        u64 z = seed;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    const u64 total = samples ? samples : (1ull << (2 * k));
    for (u64 it = 0; it < total; ++it) {
        const u64 q = samples ? (next() & kmask) : it;
        const u64 rc = fold_revcomp(q, k);
        u32 h1, p1, b1, h2, p2, b2;
        bool s1, s2;
        u64 o1, o2, R1, R2;
        loc_pick(q, rc, g, h1, p1, s1, o1);
        loc_key(h1, p1, o1, g, b1, R1);
        loc_pick(rc, q, g, h2, p2, s2, o2);
        loc_key(h2, p2, o2, g, b2, R2);
        ++n;
        if (b1 != b2 || R1 != R2 || o1 != o2) ++bad;
        if (q == rc) {
            ++self;
            if (s1 || s2) ++bad;
        } else if (s1 == s2) {
            ++bad;
        }
        if (o1 != (s1 ? rc : q)) ++bad;
        if ((u64)b1 >> (2 * t)) ++bad;
        if (g.rbits < 64 && (R1 >> g.rbits)) ++bad;
        if (p1 >= g.w) ++bad;
        for (u32 p = 0; p < g.w; ++p) {  // the pick is a minimum over the canonical m-mers of every place
            const u32 xf = (u32)(q >> (g.fbits - 2 * p)) & mmask, yr = loc_revcomp_m(xf, m);
            if (yr != ((u32)(rc >> (2 * p)) & mmask)) ++bad;  // the same place read on the other strand
            if (loc_order(xf < yr ? xf : yr) < loc_order(h1)) ++bad;
        }
        if (((u32)(o1 >> (g.fbits - 2 * p1)) & mmask) != h1) ++bad;  // h1 = the m-mer of o at pos ...
        if (h1 > loc_revcomp_m(h1, m)) ++bad;                          // ... in its canonical form
        if (!samples) {
            const u64 canon = q < rc ? q : rc;
            auto r = seen.emplace(std::make_pair(b1, R1), canon);
            if (!r.second && r.first->second != canon) ++bad;
        }
    }
    printf("{\"k\": %u, \"m\": %u, \"t\": %u, \"fits\": true, \"w\": %u, \"pbits\": %u, \"rbits\": %u, \"kmers\": %llu, \"self_complementary\": %llu, "
           "\"distinct_rows\": %llu, \"violations\": %llu}\n",
           k, m, t, g.w, g.pbits, g.rbits, (unsigned long long)n, (unsigned long long)self, (unsigned long long)seen.size(), (unsigned long long)bad);
    return bad ? 1 : 0;
}
