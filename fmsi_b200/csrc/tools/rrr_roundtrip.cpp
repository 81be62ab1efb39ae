// rrr_roundtrip.cpp — test tool for the host-side sdsl format code (not part of the product library).
// Reads an rrr_vector<63> file (.fmsi.mask), decodes it to plain bits, encodes those bits again with the product's
// multi-threaded coder and writes the result: tests/test_sdsl_io.py requires the input's bytes back.
//   rrr_roundtrip <in.mask> <out.mask>
#include <cstdio>

#include "../sdsl_io.hpp"

int main(int argc, char **argv) {
    if (argc != 3) {
        std::fprintf(stderr, "usage: rrr_roundtrip <in.mask> <out.mask>\n");
        return 2;
    }
    try {
        const fmsi::RrrFile in = fmsi::read_rrr(argv[1]);
        const fmsi::BitVec bits = fmsi::rrr_decode_all(in);
        fmsi::write_rrr(argv[2], fmsi::rrr_encode(bits));
        uint64_t ones = 0;
        for (uint64_t w = 0; w < (bits.nbits + 63) / 64; ++w) ones += (uint64_t)__builtin_popcountll(bits.w[w]);
        std::printf("%llu %llu\n", (unsigned long long)bits.nbits, (unsigned long long)ones);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
