// layout_dump — test tool (not part of the product): prints, for every record of a FASTA/FASTQ file, how many
// results the CLI's record layout (fmsi_cli.cpp: layout_record, the restatement of ms_query's record loop,
// reference src/main.cpp:328-373) will print for it: "<name>\t<results>\t<k-mer results>\n". The count includes the
// fillers the reference emits at invalid characters (the `ACGTN` -> 5 outputs quirk). tests/test_layout.py compares
// these counts with the line lengths of the reference binary's own output. No GPU is touched.
#define main fmsi_cli_main
#include "../fmsi_cli.cpp"
#undef main

// one laid-out batch: consistency checks, then its records' counts added to `rows` (pieces of one record merge)
struct Row {
    std::string name;
    unsigned long long total = 0, kmers = 0;
    bool open = false;  // the record continues in the next batch
};
static int account(const Batch &b, int k, std::vector<Row> &rows) {
    uint64_t chunk_results = 0;
    for (size_t c = 0; c < b.chunk_len.size(); ++c) {
        chunk_results += b.chunk_len[c] - (uint32_t)k + 1;
        if (b.chunk_off[c] + b.chunk_len[c] > b.bases.size()) return 5;  // chunks lie inside the batch's text
    }
    if (chunk_results != b.n_results) return 3;  // GPU chunks must cover exactly the k-mer results
    uint64_t ref_results = 0;
    for (uint32_t m : b.ref_chunks) ref_results += m;
    if (ref_results != b.n_results) return 4;    // and so must the reference's chunks (predictor granularity)
    for (const Record &rec : b.records) {
        if (rec.cont_begin != (!rows.empty() && rows.back().open)) return 6;  // pieces chain up
        if (!rec.cont_begin) {
            rows.emplace_back();
            rows.back().name.assign(b.names.data() + rec.name_begin, rec.name_len);
        }
        Row &row = rows.back();
        if (rec.cont_begin && rec.comma0 != (row.total > 0)) return 7;
        for (size_t o = rec.op_begin; o < rec.op_end; ++o) {
            row.total += b.ops[o].count;
            if (b.ops[o].kmers) row.kmers += b.ops[o].count;
        }
        row.open = rec.cont_end;
    }
    return 0;
}

//   layout_dump <k> <streaming 0|1> <file> [piece-limit]   (piece-limit: cut records into pieces of that many results)
int main(int argc, char **argv) {
    if (argc != 4 && argc != 5) return 2;
    const int k = std::atoi(argv[1]);
    const bool streaming = std::atoi(argv[2]) != 0;
    const uint64_t limit = argc == 5 ? (uint64_t)std::atoll(argv[4]) : 0;
    fmsi::BlockSource src(argv[3], 1 << 16);
    std::vector<char> block;
    std::string name, seq;
    std::vector<Row> rows;
    int rc = 0;
    size_t pieces = 0;
    while (src.next(block)) {
        fmsi::MemRecordReader rd(block.data(), block.size());
        Batch b;
        auto emit = [&](Batch &x) {
            ++pieces;
            if (!rc) rc = account(x, k, rows);
            x = Batch();
        };
        while (rd.next(name, seq) >= 0) {
            if (limit) layout_record(b, name, seq, k, streaming, limit, emit);
            else layout_record(b, name, seq, k, streaming);
        }
        emit(b);
    }
    if (rc) return rc;
    if (!rows.empty() && rows.back().open) return 8;
    for (const Row &row : rows) {
        std::fwrite(row.name.data(), 1, row.name.size(), stdout);
        std::printf("\t%llu\t%llu\n", row.total, row.kmers);
    }
    if (limit) std::fprintf(stderr, "pieces %zu\n", pieces);
    return 0;
}
