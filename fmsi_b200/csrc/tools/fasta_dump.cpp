// fasta_dump — prints the records fasta_blocks.hpp produces (block scanner + per-block parser), one
// "<name length> <sequence length> <name><sequence>\n" entry per record. tests/test_fasta_blocks.py
// compares this with the same dump made by the reference's own kseq (oracle/kseq_dump.cpp).
//   fasta_dump <block_bytes> <file | ->
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../fasta_blocks.hpp"

int main(int argc, char **argv) {
    if (argc != 3) return 2;
    fmsi::BlockSource src(argv[2], (size_t)std::atoll(argv[1]));
    std::vector<char> block;
    std::string name, seq;
    size_t blocks = 0;
    while (src.next(block)) {
        ++blocks;
        fmsi::MemRecordReader rd(block.data(), block.size());
        while (rd.next(name, seq) >= 0) {
            std::printf("%zu %zu ", name.size(), seq.size());
            std::fwrite(name.data(), 1, name.size(), stdout);
            std::fwrite(seq.data(), 1, seq.size(), stdout);
            std::fputc('\n', stdout);
        }
    }
    std::fprintf(stderr, "%zu blocks\n", blocks);
    return 0;
}
