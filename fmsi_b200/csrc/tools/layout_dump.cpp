// layout_dump — test tool (not part of the product): prints, for every record of a FASTA/FASTQ file, how many
// results the CLI's record layout (fmsi_cli.cpp: layout_record, the restatement of ms_query's record loop,
// reference src/main.cpp:328-373) will print for it: "<name>\t<results>\t<k-mer results>\n". The count includes the
// fillers the reference emits at invalid characters (the `ACGTN` -> 5 outputs quirk). tests/test_layout.py compares
// these counts with the line lengths of the reference binary's own output. No GPU is touched.
//   layout_dump <k> <streaming 0|1> <file>
#define main fmsi_cli_main
#include "../fmsi_cli.cpp"
#undef main

int main(int argc, char **argv) {
    if (argc != 4) return 2;
    const int k = std::atoi(argv[1]);
    const bool streaming = std::atoi(argv[2]) != 0;
    fmsi::BlockSource src(argv[3], 1 << 16);
    std::vector<char> block;
    std::string name, seq;
    while (src.next(block)) {
        fmsi::MemRecordReader rd(block.data(), block.size());
        Batch b;
        while (rd.next(name, seq) >= 0) layout_record(b, name, seq, k, streaming);
        uint64_t chunk_results = 0;
        for (size_t c = 0; c < b.chunk_len.size(); ++c) chunk_results += b.chunk_len[c] - (uint32_t)k + 1;
        if (chunk_results != b.n_results) return 3;  // GPU chunks must cover exactly the k-mer results
        uint64_t ref_results = 0;
        for (uint32_t m : b.ref_chunks) ref_results += m;
        if (ref_results != b.n_results) return 4;    // and so must the reference's chunks (predictor granularity)
        for (const Record &rec : b.records) {
            uint64_t total = 0, kmers = 0;
            for (size_t o = rec.op_begin; o < rec.op_end; ++o) {
                total += b.ops[o].count;
                if (b.ops[o].kmers) kmers += b.ops[o].count;
            }
            std::fwrite(b.names.data() + rec.name_begin, 1, rec.name_len, stdout);
            std::printf("\t%llu\t%llu\n", (unsigned long long)total, (unsigned long long)kmers);
        }
    }
    return 0;
}
