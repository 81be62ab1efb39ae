// oracle/kseq_dump.cpp — TEST INFRASTRUCTURE ONLY. Dumps the records the reference's own reader
// (kseq.h through parser.h's KSEQ_INIT(gzFile, gzread) and OpenFile, both included from $(REF)/src at
// build time) yields for a file, in the format of fmsi_b200/csrc/tools/fasta_dump.cpp, reading until
// kseq_read turns negative exactly like ms_query's loop (src/main.cpp:328).
//   kseq_dump <file | ->
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>

#include "parser.h"

int main(int argc, char **argv) {
    if (argc != 2) return 2;
    std::string path = argv[1];
    gzFile fp = OpenFile(path);
    kseq_t *seq = kseq_init(fp);
    while (kseq_read(seq) >= 0) {
        std::printf("%zu %zu ", (size_t)seq->name.l, (size_t)seq->seq.l);
        std::fwrite(seq->name.s, 1, seq->name.l, stdout);
        std::fwrite(seq->seq.s, 1, seq->seq.l, stdout);
        std::fputc('\n', stdout);
    }
    kseq_destroy(seq);
    gzclose(fp);
    return 0;
}
