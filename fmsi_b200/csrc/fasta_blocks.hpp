// fasta_blocks.hpp — FASTA/FASTQ input for the host pipeline, split so that only a cheap scan is serial.
//
// The reference reads records one at a time with kseq (src/kseq.h:173-226, parser.h:14-27) on the one
// thread that also queries and prints. Here the input thread only finds where records END
// (scan_record: a memchr per line, nothing is copied or split) and hands out blocks of whole records;
// worker threads parse their block with the full kseq semantics (MemRecordReader). Both follow
// kseq_read statement by statement, so record boundaries, names and sequences are those of the
// reference for any input — including FASTQ, multi-line records, '\r', empty lines, garbage before a
// header and a truncated last record (tests/test_fasta_blocks.py checks this against the reference's
// own kseq on adversarial inputs at tiny block sizes).
#pragma once
#include <zlib.h>

#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace fmsi {

// kseq_read() over a memory block whose end is the end of the input.
class MemRecordReader {
  public:
    MemRecordReader(const char *p, size_t n) : p_(p), end_(p + n) {}
    // Sequence length (>= 0), or kseq's negative codes: -1 end of input, -2 truncated quality.
    int64_t next(std::string &name, std::string &seq) {
        int c;
        if (last_char_ == 0) {  // jump to the next header line (kseq.h:179-183)
            while ((c = getc()) >= 0 && c != '>' && c != '@') {}
            if (c < 0) return c;
            last_char_ = c;
        }
        seq.clear();
        const int64_t r = getuntil(kSpace, name, &c, false);
        if (r < 0) return r;
        if (c != '\n') getuntil(kLine, scratch_, nullptr, false);  // comment
        while ((c = getc()) >= 0 && c != '>' && c != '+' && c != '@') {
            if (c == '\n') continue;  // empty line
            seq.push_back((char)c);
            getuntil(kLine, seq, nullptr, true);
        }
        if (c == '>' || c == '@') last_char_ = c;
        if (c != '+') return (int64_t)seq.size();  // FASTA
        while ((c = getc()) >= 0 && c != '\n') {}  // rest of the '+' line
        if (c == -1) return -2;
        scratch_.clear();
        while (getuntil(kLine, scratch_, nullptr, true) >= 0 && scratch_.size() < seq.size()) {}
        last_char_ = 0;
        if (seq.size() != scratch_.size()) return -2;
        return (int64_t)seq.size();
    }

  private:
    enum { kSpace = 0, kLine = 2 };
    int getc() { return p_ < end_ ? (int)(unsigned char)*p_++ : -1; }
    // ks_getuntil2 (kseq.h:96-144): -1 iff nothing was left to read
    int64_t getuntil(int delimiter, std::string &str, int *dret, bool append) {
        if (dret) *dret = 0;
        if (!append) str.clear();
        if (p_ >= end_) return -1;
        const char *i;
        if (delimiter == kLine) {
            i = (const char *)std::memchr(p_, '\n', (size_t)(end_ - p_));
            if (!i) i = end_;
        } else {
            for (i = p_; i < end_; ++i)
                if (std::isspace((unsigned char)*i)) break;
        }
        str.append(p_, (size_t)(i - p_));
        if (i < end_) {
            if (dret) *dret = (unsigned char)*i;
            p_ = i + 1;
        } else {
            p_ = end_;
        }
        if (delimiter == kLine && str.size() > 1 && str.back() == '\r') str.pop_back();
        return (int64_t)str.size();
    }
    const char *p_, *end_;
    int last_char_ = 0;
    std::string scratch_;
};

struct ScanResult {
    enum Kind { RECORD, END, TRUNCATED, MORE } kind;
    size_t end;  // RECORD: one past the record (a following header character is NOT included);
                 // MORE: offset from which the data must be kept (start of the unfinished record)
};

namespace detail {
// Length bookkeeping of a kstring that lines are appended to by ks_getuntil2(KS_SEP_LINE, append):
// after every appended line one trailing '\r' is dropped when the string is longer than 1 (kseq.h:140).
struct LineAccumulator {
    uint64_t len = 0, trailing_cr = 0;
    void append(const char *b, const char *e) {
        const uint64_t n = (uint64_t)(e - b);
        if (n) {
            uint64_t cr = 0;
            while (cr < n && e[-1 - (ptrdiff_t)cr] == '\r') ++cr;
            trailing_cr = cr == n ? trailing_cr + n : cr;
            len += n;
        }
    }
    void strip() {
        if (len > 1 && trailing_cr > 0) {
            --len;
            --trailing_cr;
        }
    }
};
}  // namespace detail

// Where does the first record at or after buf[pos] end? Mirrors kseq_read started with last_char == 0.
// `final`: buf[n] is the end of the input; otherwise running out of data yields MORE.
inline ScanResult scan_record(const char *buf, size_t pos, size_t n, bool final) {
    size_t h = pos;
    while (h < n && buf[h] != '>' && buf[h] != '@') ++h;  // kseq.h:180
    if (h >= n) return {final ? ScanResult::END : ScanResult::MORE, n};
    size_t p = h + 1;
    if (p >= n) return {final ? ScanResult::END : ScanResult::MORE, h};  // nothing after the header char: ks_getuntil -> -1
    {   // name and comment: through the first '\n'
        const char *nl = (const char *)std::memchr(buf + p, '\n', n - p);
        if (!nl && !final) return {ScanResult::MORE, h};
        p = nl ? (size_t)(nl - buf) + 1 : n;
    }
    detail::LineAccumulator seq;
    for (;;) {  // sequence lines (kseq.h:190-194)
        if (p >= n) {
            if (!final) return {ScanResult::MORE, h};
            return {ScanResult::RECORD, n};  // FASTA record closed by the end of the input
        }
        const char c = buf[p];
        if (c == '>' || c == '@') return {ScanResult::RECORD, p};
        ++p;
        if (c == '+') break;
        if (c == '\n') continue;
        if (p >= n) {  // c was the last byte: ks_getuntil2 returns -1 before its '\r' rule
            if (!final) return {ScanResult::MORE, h};
            seq.append(buf + p - 1, buf + p);
            continue;
        }
        const char *nl = (const char *)std::memchr(buf + p, '\n', n - p);
        if (!nl && !final) return {ScanResult::MORE, h};
        const char *e = nl ? nl : buf + n;
        seq.append(buf + p - 1, e);
        seq.strip();
        p = nl ? (size_t)(nl - buf) + 1 : n;
    }
    {   // FASTQ: the rest of the '+' line (kseq.h:208-209)
        const char *nl = p < n ? (const char *)std::memchr(buf + p, '\n', n - p) : nullptr;
        if (!nl) return {final ? ScanResult::TRUNCATED : ScanResult::MORE, h};
        p = (size_t)(nl - buf) + 1;
    }
    detail::LineAccumulator qual;
    do {  // quality lines until as long as the sequence (kseq.h:210)
        if (p >= n) {
            if (!final) return {ScanResult::MORE, h};
            break;
        }
        const char *nl = (const char *)std::memchr(buf + p, '\n', n - p);
        if (!nl && !final) return {ScanResult::MORE, h};
        const char *e = nl ? nl : buf + n;
        qual.append(buf + p, e);
        qual.strip();
        p = nl ? (size_t)(nl - buf) + 1 : n;
    } while (qual.len < seq.len);
    if (qual.len != seq.len) return {ScanResult::TRUNCATED, h};  // kseq_read -> -2: the reference stops reading
    return {ScanResult::RECORD, p};
}

// Reads the query file (plain or gzip, file or stdin: parser.h:14-27) and cuts it into blocks of whole
// records of about `block_bytes`. A block is meant to be parsed by MemRecordReader on its own.
class BlockSource {
  public:
    BlockSource(const std::string &path, size_t block_bytes) : block_bytes_(block_bytes ? block_bytes : 1) {
        FILE *in = path == "-" ? stdin : std::fopen(path.c_str(), "r");
        if (!in) throw std::invalid_argument("couldn't open file " + path);  // uncaught in the reference too
        fp_ = gzdopen(fileno(in), "r");
        if (!fp_) throw std::invalid_argument("couldn't open file " + path);
        gzbuffer(fp_, 1 << 20);
    }
    ~BlockSource() {
        if (fp_) gzclose(fp_);
    }
    BlockSource(const BlockSource &) = delete;
    BlockSource &operator=(const BlockSource &) = delete;

    // The next block (>= 1 whole record) in `out`; false when the input is exhausted.
    bool next(std::vector<char> &out) {
        out.clear();
        if (done_) return false;
        out.swap(carry_);
        size_t scanned = 0;               // out[0, scanned) = whole records
        size_t target = block_bytes_;
        for (;;) {
            if (!eof_ && out.size() < target) fill(out, target);
            const size_t n = out.size();
            size_t keep = n;
            bool more = false;
            while (scanned < n || (eof_ && scanned == n)) {
                const ScanResult r = scan_record(out.data(), scanned, n, eof_);
                if (r.kind == ScanResult::RECORD) {
                    scanned = r.end;
                    if (scanned >= n) break;
                    continue;
                }
                if (r.kind == ScanResult::MORE) {
                    more = true;
                    keep = r.end;
                } else {
                    done_ = true;  // END, or TRUNCATED: the reference's loop ends at a negative kseq_read
                }
                break;
            }
            if (done_ || (eof_ && !more)) {
                done_ = true;
                out.resize(scanned);
                return scanned > 0;
            }
            if (scanned > 0) {  // emit the whole records, carry the unfinished tail
                carry_.assign(out.begin() + (ptrdiff_t)keep, out.end());
                out.resize(scanned);
                return true;
            }
            // no whole record yet: drop leading garbage, then read more (a record longer than the block)
            if (keep > 0) {
                out.erase(out.begin(), out.begin() + (ptrdiff_t)keep);
            }
            if (out.size() >= target) target *= 2;
        }
    }

  private:
    void fill(std::vector<char> &buf, size_t target) {
        size_t have = buf.size();
        buf.resize(target);
        while (have < target) {
            const size_t want = std::min<size_t>(target - have, 1u << 30);
            const int got = gzread(fp_, buf.data() + have, (unsigned)want);
            if (got <= 0) {
                eof_ = true;
                break;
            }
            have += (size_t)got;
        }
        buf.resize(have);
    }
    gzFile fp_ = nullptr;
    size_t block_bytes_;
    std::vector<char> carry_;
    bool eof_ = false, done_ = false;
};

}  // namespace fmsi
