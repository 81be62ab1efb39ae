// predictor_check.cpp — test tool (not part of the product): the CLI splits the exact -S strand-predictor replay into a
// per-chunk summary made on worker threads and an O(1)-per-chunk in-order pass (fmsi_cli.cpp: summarize / replay).
// This checks, on random strand values and chunkings, that the split gives the values and the predictor state of the
// plain sequential replay (predictor.hpp: replay_streaming_chunk, the restatement of fms_index.h:181-254).
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../predictor.hpp"
using namespace fmsi;
int main() {
    std::mt19937_64 rng(7);
    for (int trial = 0; trial < 2000; ++trial) {
        const bool orders = trial & 1;
        const QueryMode mode = (trial & 2) ? QueryMode::All : QueryMode::Or;
        const int style = (trial >> 2) % 3;  // 0: random, 1: mostly consistent strands, 2: heavy conflicts
        std::vector<uint32_t> chunks;
        size_t n = 0;
        const int nc = 1 + rng() % 200;
        for (int c = 0; c < nc; ++c) { chunks.push_back(1 + rng() % 70); n += chunks.back(); }
        std::vector<int64_t> f(n), r(n);
        for (size_t q = 0; q < n; ++q) {
            auto val = [&]() -> int64_t { return orders ? ((rng() % 3) ? (int64_t)(rng() % 1000) : -1) : (int64_t)(rng() % 3) - 1; };
            f[q] = val();
            r[q] = style == 1 && (rng() % 10) ? (orders ? -1 : -1) : val();
            if (style == 2 && !orders) { f[q] = rng() % 2; r[q] = 1 - f[q]; }
        }
        // old: sequential
        StrandPredictor p0;
        std::vector<int64_t> out0(n), out1(n);
        size_t q0 = 0;
        for (uint32_t m : chunks) {
            replay_streaming_chunk(p0, mode, orders, m, [&](size_t q) { return f[q0 + q]; }, [&](size_t q) { return r[q0 + q]; },
                                   [&](size_t q, int64_t v) { out0[q0 + q] = orders ? v : (v == 1); });
            q0 += m;
        }
        // new: summaries + O(1) replay
        std::vector<ChunkSummary> cs(chunks.size());
        q0 = 0;
        for (size_t c = 0; c < chunks.size(); ++c) {
            const uint32_t m = chunks[c];
            streaming_chunk_with_order(false, mode, orders, m, [&](size_t q) { return f[q0 + q]; }, [&](size_t q) { return r[q0 + q]; },
                                       [&](size_t q, int64_t v) { out1[q0 + q] = orders ? v : (v == 1); }, cs[c].fpr[0], cs[c].bpr[0]);
            bool differs = false;
            streaming_chunk_with_order(true, mode, orders, m, [&](size_t q) { return f[q0 + q]; }, [&](size_t q) { return r[q0 + q]; },
                                       [&](size_t q, int64_t v) { differs |= out1[q0 + q] != (orders ? v : (int64_t)(v == 1)); }, cs[c].fpr[1], cs[c].bpr[1]);
            cs[c].differs = differs;
            q0 += m;
        }
        StrandPredictor p1;
        q0 = 0;
        size_t redone = 0;
        for (size_t c = 0; c < chunks.size(); ++c) {
            const uint32_t m = chunks[c];
            const bool swap = p1.predict_swap();
            if (swap && cs[c].differs) {
                int a, b;
                streaming_chunk_with_order(true, mode, orders, m, [&](size_t q) { return f[q0 + q]; }, [&](size_t q) { return r[q0 + q]; },
                                           [&](size_t q, int64_t v) { out1[q0 + q] = orders ? v : (v == 1); }, a, b);
                ++redone;
            }
            p1.log_result(cs[c].fpr[swap], cs[c].bpr[swap]);
            q0 += m;
        }
        if (out0 != out1 || p0.score != p1.score || p0.previous != p1.previous || p0.result_scores[0] != p1.result_scores[0] || p0.result_scores[1] != p1.result_scores[1]) {
            printf("MISMATCH trial %d\n", trial);
            return 1;
        }
        if (trial < 6) printf("trial %d ok, chunks %d, redone %zu\n", trial, nc, redone);
    }
    printf("all ok\n");
    return 0;
}
