/* fmsi_oracle_main.c — TEST INFRASTRUCTURE ONLY.
 * Command-line wrapper around the plain-C oracle with the reference's `fmsi query` /
 * `fmsi lookup` flags (src/main.cpp:238-375), so differential tests can run
 *     oracle/_ref/fmsi query ...   vs   oracle/_build/fmsi_oracle query ...
 * Extra flag -C prints the SURVEY §8(d) work counters to stderr.
 */
#define _POSIX_C_SOURCE 200809L
#include "fmsi_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

static char *slurp(FILE *f, size_t *n) {
    size_t cap = 1 << 20, len = 0;
    char *b = (char *)malloc(cap);
    size_t r;
    while ((r = fread(b + len, 1, cap - len, f)) > 0) {
        len += r;
        if (len == cap) {
            cap *= 2;
            b = (char *)realloc(b, cap);
        }
    }
    *n = len;
    return b;
}

int main(int argc, char **argv) {
    if (argc < 3 || (strcmp(argv[1], "query") && strcmp(argv[1], "lookup"))) {
        fprintf(stderr, "usage: fmsi_oracle query|lookup [-q FILE] [-k INT] [-S] [-O] [-C] <index-prefix>\n");
        return 1;
    }
    int output_orders = strcmp(argv[1], "lookup") == 0;
    const char *prefix = argv[argc - 1];
    int sub_argc = argc - 2; /* drop the prefix, keep argv[1] as argv[0] for getopt */
    char **sub_argv = argv + 1;
    const char *qfn = "-";
    int k = 0, has_klcp = 0, mode = FMSI_ORACLE_MODE_OR, counters = 0, c;
    while ((c = getopt(sub_argc, sub_argv, "q:k:OSC")) >= 0) {
        switch (c) {
        case 'q': qfn = optarg; break;
        case 'k': k = atoi(optarg); break;
        case 'O': mode = FMSI_ORACLE_MODE_ALL; break;
        case 'S': has_klcp = 1; break;
        case 'C': counters = 1; break;
        default: return 1;
        }
    }
    fmsi_oracle_index *idx = fmsi_oracle_load(prefix, has_klcp);
    if (!idx || fmsi_oracle_size(idx) == 0) {
        fprintf(stderr, "ERROR: index not correctly loaded.\n");
        return 1;
    }
    if (has_klcp != fmsi_oracle_has_klcp(idx)) {
        fprintf(stderr, "ERROR: kLCP array was not constructed for the given index.\n");
        return 1;
    }
    if (k != 0 && k != fmsi_oracle_k(idx)) {
        fprintf(stderr, "ERROR: Mismatch. Provided k (%d) does not match the k of the index (%d).\n", k,
                fmsi_oracle_k(idx));
        return 1;
    }
    if (k == 0) k = fmsi_oracle_k(idx);
    FILE *f = strcmp(qfn, "-") ? fopen(qfn, "rb") : stdin;
    if (!f) {
        fprintf(stderr, "couldn't open file %s\n", qfn);
        return 2;
    }
    size_t n;
    char *text = slurp(f, &n);
    fmsi_oracle_buf out = {0, 0, 0};
    fmsi_oracle_ms_query(idx, text, n, k, mode, has_klcp, output_orders, &out);
    if (out.len) fwrite(out.s, 1, out.len, stdout);
    if (counters) {
        fmsi_oracle_counters ct;
        fmsi_oracle_counters_get(idx, &ct);
        fprintf(stderr,
                "{\"kmers\": %llu, \"lf_steps\": %llu, \"rank_sectors\": %llu, \"mask_sectors\": %llu, "
                "\"klcp_steps\": %llu, \"strand_searches\": %llu}\n",
                (unsigned long long)ct.kmers, (unsigned long long)ct.lf_steps,
                (unsigned long long)ct.rank_sectors, (unsigned long long)ct.mask_sectors,
                (unsigned long long)ct.klcp_steps, (unsigned long long)ct.strand_searches);
    }
    fmsi_oracle_buf_free(&out);
    free(text);
    fmsi_oracle_free(idx);
    return 0;
}
