// predictor.hpp — host-side replay of the reference's stateful strand predictor.
//
// The reference threads a saturating-counter strand_predictor (src/fms_index.h:18-49) through every
// query: it decides which strand is searched first, and because `-O` and `lookup` stop at the first
// decided strand, the printed value can depend on the whole query history whenever a k-mer occurs
// on BOTH strands of the superstring. The GPU computes the history-free pair
// (f, r) = value on the k-mer / on its reverse complement (FMSI_GPU_STRANDS_BOTH); this file turns
// those pairs into exactly what the reference prints, by running the same state machine over them
// in query order. O(1) integer work per k-mer.
#pragma once
#include <algorithm>
#include <cstdint>

namespace fmsi {

struct StrandPredictor {  // reference strand_predictor, fms_index.h:18-49
    int score = 0;
    int result_scores[2] = {0, 0};
    int previous = -1;
    static int clipped(int x) { return std::max(-7, std::min(7, x)); }
    void log_result(int forward_result, int reverse_result) {
        const int difference = forward_result - reverse_result;
        score = clipped(score + difference);
        if (previous != -1) result_scores[previous] = clipped(result_scores[previous] + difference);
        previous = forward_result > reverse_result;
    }
    bool predict_swap() const {
        if (previous != -1 && result_scores[previous] != 0) return result_scores[previous] < 0;
        return score < 0;
    }
};

enum class QueryMode { Or, All };

// query_kmers_single for ONE k-mer given both strand values (fms_index.h:265-316).
// presence modes: f, r in {-1, 0, 1}; orders: id or -1. Returns the printed value.
inline int64_t replay_single(StrandPredictor &p, QueryMode mode, bool orders, int64_t f, int64_t r) {
    const bool swap = p.predict_swap();
    const int64_t first = swap ? r : f, second = swap ? f : r;
    int64_t got = first;
    int fpr = (int)got, bpr = 0;
    if (orders) {
        if (fpr >= 0) fpr = 1;
        else {
            got = second;
            bpr = got >= 0 ? 1 : -1;
        }
    } else if (mode == QueryMode::Or) {
        if (got != 1) {
            got = second;
            bpr = (int)got;
        }
    } else {
        if (got == -1) {
            got = second;
            bpr = (int)got;
        }
    }
    if (swap) std::swap(fpr, bpr);
    p.log_result(fpr, bpr);
    return got;
}

// query_kmers_streaming for ONE reference chunk of m k-mers (fms_index.h:181-254) with the strand order given:
// f[q], r[q] are the strand values of the k-mer at chunk position q; out[q] receives the merged value; (fpr, bpr)
// are the predictor inputs the reference passes to log_result for this chunk.
template <typename GetF, typename GetR, typename Put>
inline void streaming_chunk_with_order(bool swap, QueryMode mode, bool orders, size_t m, GetF f, GetR r, Put out, int &fpr_out, int &bpr_out) {
    const bool max_ones = mode == QueryMode::All;
    int fpr = 0, bpr = 0;
    for (size_t q = 0; q < m; ++q) {
        const int64_t first = swap ? r(q) : f(q), second = swap ? f(q) : r(q);
        int64_t res = first;
        fpr += orders ? (first >= 0 ? 1 : -1) : (int)first;
        const bool skip = (orders && first >= 0) || first == 1 || (first == 0 && max_ones);  // :212
        if (!skip) {
            bpr += orders ? (second >= 0 ? 1 : -1) : (int)second;
            res = std::max(first, second);
        }
        out(q, res);
    }
    if (swap) std::swap(fpr, bpr);
    fpr_out = fpr;
    bpr_out = bpr;
}

// The same chunk under the predictor: its state picks the order, the chunk's totals update it.
template <typename GetF, typename GetR, typename Put>
inline void replay_streaming_chunk(StrandPredictor &p, QueryMode mode, bool orders, size_t m, GetF f, GetR r, Put out) {
    int fpr, bpr;
    streaming_chunk_with_order(p.predict_swap(), mode, orders, m, f, r, out, fpr, bpr);
    p.log_result(fpr, bpr);
}

// What the predictor needs from a chunk, computed ahead of the in-order replay (by any thread): the log_result
// inputs under either strand order, and whether the merged values depend on the order at all (they do only where
// a k-mer is decided differently by its two strands). The in-order pass is then O(1) per chunk unless `differs`.
struct ChunkSummary {
    int32_t fpr[2], bpr[2];  // [swap]
    bool differs;
};

}  // namespace fmsi
