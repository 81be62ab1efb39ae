// fmsi_cli.cpp — `fmsi query` / `fmsi lookup` / `fmsi index` front end over libfmsi_gpu.so.
//
// Drop-in for the query path of the reference CLI (reference src/main.cpp: ms_query :238-375,
// usage texts :102-130, dispatch :668-699): same flags (incl. -f and|xor|INT-INT), same index files,
// byte-identical stdout; and for `fmsi index` (ms_index :172-236) on the GPU builder, byte-identical files.
// Everything else the reference CLI does (export, merge, set operations, compact, clean) is out of scope
// for the GPU engine and is forwarded to the unchanged reference binary when $FMSI_REFERENCE_BIN points at one.
//
// Flow per block of records:  record-boundary scan (fasta_blocks.hpp; the only serial input step)
//   -> kseq-exact parse -> valid ACGT runs -> GPU chunks -> fmsi_gpu_query_chunks (both strands)
//   -> strand-predictor replay (predictor.hpp, exact reference semantics whatever the index/mask)
//   -> text formatter.
// $FMSI_GPU_STRANDS=lazy skips the replay (forward strand first, reverse complement only if
// undecided): identical output whenever the strand predictor cannot change results (or-mode
// always; -O on max-ones masks; lookup when no k-mer is ON on both strands), and much cheaper for -S.
#include <fmsi_gpu.h>
#include <malloc.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <functional>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "fasta_blocks.hpp"
#include "predictor.hpp"

namespace {

const char *kVersion = "0.4.0-b200";

int usage() {
    std::cerr << std::endl;
    std::cerr << "Program: FMSI - a tool for space-efficient k-mer set indexing via masked superstrings." << std::endl;
    std::cerr << "         (B200 GPU query engine; drop-in for `query` and `lookup`)" << std::endl;
    std::cerr << "Version: " << kVersion << std::endl << std::endl;
    std::cerr << "Usage:   fmsi <command> [options]" << std::endl << std::endl;
    std::cerr << "Command (GPU):" << std::endl;
    std::cerr << "    query   - Queries k-mers against an index." << std::endl;
    std::cerr << "    lookup  - Return unique hashes of present k-mers." << std::endl;
    std::cerr << "    index   - Creates a BWT based index of the given masked superstring (same files as the reference)." << std::endl << std::endl;
    std::cerr << "Command (forwarded to the reference binary named by $FMSI_REFERENCE_BIN):" << std::endl;
    std::cerr << "    export union inter diff symdiff merge compact clean" << std::endl << std::endl;
    return 1;
}

void usage_functions() {
    std::cerr << "  -f FUNCTION - Demasking function to determine k-mer presence; recognized functions:" << std::endl;
    std::cerr << "    or      - represented when at least 1 ON occurrence [default]" << std::endl;
    std::cerr << "    all     - all occurrence are either ON or OFF (equivalent to -O flag for queries)" << std::endl;
    std::cerr << "    and     - represented when no OFF occurrence" << std::endl;
    std::cerr << "    xor     - represented when an odd number of ON occurrences" << std::endl;
    std::cerr << "    INT-INT - represented when in the bounds" << std::endl;
}

int usage_query() {
    std::cerr << std::endl;
    std::cerr << "Usage:   fmsi query [options] <index-prefix>" << std::endl << std::endl;
    std::cerr << "Options (stable):" << std::endl;
    std::cerr << "  -q FILE - Path to FASTA/FASTQ with queries [default: stdin]" << std::endl;
    std::cerr << "  -k INT  - Size of k-mers [default: infer automatically from index]" << std::endl;
    std::cerr << "  -S      - Use kLCP array for streamed queries (increses memory consumption)" << std::endl;
    std::cerr << "  -O      - FMSI uses properties of max-one masked superstrings to speed up queries" << std::endl;
    std::cerr << "            Use only if a masked superstring with maximum number of ones is indexed." << std::endl;
    std::cerr << "Parameters (experimental, using f-MS framework):" << std::endl;
    usage_functions();
    std::cerr << std::endl;
    return 1;
}

int usage_lookup() {
    std::cerr << std::endl;
    std::cerr << "Usage:   fmsi lookup [options] <index-prefix>" << std::endl << std::endl;
    std::cerr << "Options (stable):" << std::endl;
    std::cerr << "  -q FILE - Path to FASTA/FASTQ with queries [default: stdin]" << std::endl;
    std::cerr << "  -k INT  - Size of k-mers [default: infer automatically from index]" << std::endl;
    std::cerr << "  -S      - Use kLCP array for streamed queries (increses memory consumption)" << std::endl;
    std::cerr << std::endl;
    return 1;
}

int usage_query(bool lookup) { return lookup ? usage_lookup() : usage_query(); }

// -f names the reference accepts (src/functions.h:23-57); only or/all run on the GPU.
bool known_function(const std::string &name) {
    if (name == "or" || name == "and" || name == "xor" || name == "all") return true;
    for (size_t i = 1; i + 1 < name.size(); ++i) {
        if (name[i] == '-') {
            bool valid = true;
            for (size_t j = 0; j < name.size(); ++j)
                if (j != i && (name[j] < '0' || name[j] > '9')) valid = false;
            if (valid) return true;
            break;
        }
    }
    return false;
}

int forward_to_reference(int argc, char *argv[]) {
    const char *ref = std::getenv("FMSI_REFERENCE_BIN");
    if (!ref || !*ref) {
        std::cerr << "ERROR: `" << (argc > 1 ? argv[1] : "") << "` is not part of the GPU query engine. "
                  << "Set FMSI_REFERENCE_BIN to the reference fmsi binary to forward it." << std::endl;
        return 1;
    }
    std::vector<char *> av(argv, argv + argc);
    av[0] = const_cast<char *>(ref);
    av.push_back(nullptr);
    execv(ref, av.data());
    std::cerr << "ERROR: cannot execute " << ref << std::endl;
    return 1;
}

inline bool is_acgt(unsigned char ch) {  // nucleotideToInt[ch] != 4, src/kmers.h:3-20
    switch (ch) {
    case 'A': case 'C': case 'G': case 'T': case 'a': case 'c': case 'g': case 't': return true;
    default: return false;
    }
}

struct Op {  // one stretch of a record's output
    uint64_t count;
    bool kmers;  // true: `count` consecutive k-mer results; false: `count` invalid-position fillers
};
struct Record {
    size_t name_begin, name_len;  // into Batch::names
    size_t op_begin, op_end;
    // a record too long for one batch is laid out in pieces (layout_record with a limit): a piece that continues an
    // earlier one prints no name, one that is continued prints no newline; comma0 = results were printed before it
    bool cont_begin = false, cont_end = false, comma0 = false;
};

struct Batch {
    std::vector<Record> records;
    std::string names;                  // record names back to back
    std::vector<Op> ops;
    std::string bases;                  // valid runs (>= k) back to back
    std::vector<uint64_t> chunk_off;    // GPU chunks
    std::vector<uint32_t> chunk_len;
    std::vector<uint64_t> res_off;
    std::vector<uint32_t> ref_chunks;   // k-mers per REFERENCE chunk, in order (predictor granularity)
    uint64_t n_results = 0;
    void clear() {
        records.clear();
        names.clear();
        ops.clear();
        bases.clear();
        chunk_off.clear();
        chunk_len.clear();
        res_off.clear();
        ref_chunks.clear();
        n_results = 0;
    }
};

// ms_query's record loop (main.cpp:328-373) turned into a layout: which k-mers exist, where the
// reference cuts its chunks, and how many filler results each invalid character produces.
// With `limit` > 0 a long record is cut into pieces: once the batch holds `limit` results, at the next boundary
// between two of the reference's chunks (so that the predictor sees the same chunks) the batch is handed to
// `emit` — which must leave it cleared — and the record continues in a fresh one. The reference holds a record in
// memory and walks it chunk by chunk (main.cpp:337-354); the pieces keep the result buffers bounded the same way.
void layout_record(Batch &b, const std::string &name, const std::string &s, int k, bool streaming, uint64_t limit = 0,
                   const std::function<void(Batch &)> &emit = nullptr) {
    Record rec;
    rec.name_begin = b.names.size();
    rec.name_len = std::strlen(name.c_str());  // C-string semantics of `cout << seq->name.s`
    b.names.append(name.data(), rec.name_len);
    rec.op_begin = b.ops.size();
    int64_t sequence_length = (int64_t)s.size();
    int64_t max_chunk = 400;
    max_chunk = k + std::max((int64_t)10, std::min(max_chunk, 2 * (int64_t)std::sqrt((double)sequence_length)));
    const char *sequence = s.data();
    const uint32_t gpu_max = streaming ? (uint32_t)(FMSI_GPU_MAX_STREAM_KMERS + k - 1) : 0xFFFFFF00u;
    bool printed = false;  // the record has produced results (comma state of `lookup` output)
    // k-mers [p0, p0 + m) of the valid run at `run`: text, GPU chunks (overlapping by k-1), one op
    auto add_segment = [&](const char *run, uint64_t p0, uint64_t m) {
        if (!m) return;
        const uint64_t base0 = b.bases.size();
        b.bases.append(run + p0, (size_t)(m + (uint64_t)k - 1));
        for (uint64_t p = 0; p < m;) {
            const uint64_t left = m + (uint64_t)k - 1 - p;
            const uint32_t len = (uint32_t)std::min<uint64_t>(left, gpu_max);
            b.chunk_off.push_back(base0 + p);
            b.chunk_len.push_back(len);
            b.res_off.push_back(b.n_results + p);
            p += len - k + 1;
        }
        b.ops.push_back({m, true});
        b.n_results += m;
        printed = true;
    };
    auto cut_piece = [&]() {  // close this piece of the record, hand the batch over, open the next piece
        rec.op_end = b.ops.size();
        rec.cont_end = true;
        b.records.push_back(rec);
        emit(b);
        rec = Record();
        rec.name_begin = b.names.size();
        rec.name_len = 0;
        rec.op_begin = b.ops.size();
        rec.cont_begin = true;
        rec.comma0 = printed;
    };
    while (sequence_length > 0) {
        int64_t current_length = 0;
        while (current_length < sequence_length && is_acgt((unsigned char)sequence[current_length])) ++current_length;
        if (current_length >= k) {
            const uint64_t run_kmers = (uint64_t)(current_length - k + 1);
            uint64_t seg0 = 0, done = 0;   // k-mers [seg0, done) of the run wait for their segment
            int64_t cur = current_length;  // reference chunks (main.cpp:340-354)
            while (cur >= k) {
                const int64_t chunk_length = std::min(cur, max_chunk);
                b.ref_chunks.push_back((uint32_t)(chunk_length - k + 1));
                cur -= chunk_length - k + 1;
                done += (uint64_t)(chunk_length - k + 1);
                if (limit && emit && cur >= k && b.n_results + (done - seg0) >= limit) {
                    add_segment(sequence, seg0, done - seg0);
                    seg0 = done;
                    cut_piece();
                }
            }
            add_segment(sequence, seg0, run_kmers - seg0);
            sequence += run_kmers;
            sequence_length -= (int64_t)run_kmers;
            current_length -= (int64_t)run_kmers;
        }
        // current_length < k characters of the run remain, then the invalid character (main.cpp:355-370)
        sequence_length -= current_length + 1;
        sequence += current_length + 1;
        if (sequence_length >= 0) {
            b.ops.push_back({(uint64_t)std::min<int64_t>(k, current_length + 1), false});
            printed = true;
            if (limit && emit && sequence_length > 0 && b.ops.size() - rec.op_begin >= limit) cut_piece();  // a record of invalid characters
        }
    }
    rec.op_end = b.ops.size();
    b.records.push_back(std::move(rec));
}

// ---------------------------------------------------------------------------------------------
// Host pipeline. The reference handles one record at a time on one thread (main.cpp:328-373); here
//   reader (main thread)  : reads the file and cuts it into blocks of whole records (fasta_blocks.hpp:
//                           a memchr per line) — starts while the index is still loading
//   workers (W threads)   : per block: kseq-exact parsing + layout; GPU query on a free index replica;
//                           later its text formatting
//   sequencer (1 thread)  : strand-predictor replay, strictly in block order (the predictor's state
//                           runs through every k-mer of the run) — skipped in lazy mode
//   writer (1 thread)     : stdout, in block order
// With $FMSI_GPU_DEVICES the replicas live on several GPUs and blocks go to whichever is free.
struct Job {
    uint64_t seq = 0;
    size_t units = 0;         // what the job counts towards the pipeline's in-flight bound
    std::vector<char> block;  // raw input: whole records
    Batch b;
    std::vector<uint8_t> raw8, fin8;
    std::vector<int64_t> raw64, fin64;
    std::vector<fmsi::ChunkSummary> summaries;  // per reference chunk (exact -S mode), made by a worker
    std::string out;
};

struct Config {
    int k = 0;
    bool streaming = false, orders = false, lazy = false;
    bool general = false;  // -f and|xor|INT-INT: counts over both strands, no predictor (fms_index.h:317-327)
    fmsi_gpu_function f{};
    fmsi::QueryMode mode = fmsi::QueryMode::Or;
};

class Pipeline {
  public:
    Pipeline(const Config &cfg, std::vector<fmsi_gpu_index *> members, int workers)
        : cfg_(cfg), members_(std::move(members)), member_busy_(members_.size(), false) {
        max_inflight_ = (size_t)(2 * workers + 2);
        for (int w = 0; w < workers; ++w) threads_.emplace_back([this] { worker(); });
        threads_.emplace_back([this] { sequencer(); });
        threads_.emplace_back([this] { writer(); });
    }
    // reader side: blocks while too many batches, or too much work (input bytes ~ results), are in flight
    void submit(Job *job) { enqueue(job, kParse, job->block.size()); }
    // a batch laid out by the reader itself (a piece of a record too long for one batch)
    void submit_parsed(Job *job) { enqueue(job, kQuery, (size_t)job->b.n_results + job->b.bases.size()); }
    bool finish() {
        {
            std::unique_lock<std::mutex> lk(mu_);
            closed_ = true;
            cv_.notify_all();
        }
        for (auto &t : threads_) t.join();
        return !failed_;
    }

  private:
    enum Stage { kParse, kQuery, kFormat };
    struct Task {
        Job *job;
        Stage stage;
    };
    void enqueue(Job *job, Stage stage, size_t units) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return failed_ || inflight_ == 0 || (inflight_ < max_inflight_ && inflight_units_ + units <= kMaxInflightUnits); });
        job->seq = submitted_++;
        job->units = units;
        ++inflight_;
        inflight_units_ += units;
        if (stage == kQuery) {  // queries leave in block order
            auto at = std::find_if(tasks_.begin(), tasks_.end(), [&](const Task &x) { return x.stage == kQuery && x.job->seq > job->seq; });
            tasks_.insert(at, {job, kQuery});
        } else {
            tasks_.push_back({job, stage});
        }
        cv_.notify_all();
    }
    void fail(const std::string &msg) {
        std::unique_lock<std::mutex> lk(mu_);
        if (!failed_) std::cerr << "ERROR: GPU query failed: " << msg << std::endl;
        failed_ = true;
        cv_.notify_all();
    }
    void worker() {
        for (;;) {
            Task t;
            int member = -1;
            {
                std::unique_lock<std::mutex> lk(mu_);
                for (;;) {
                    if (failed_) return;
                    // formatting first (drains memory), then queries when a replica is free, then parsing
                    auto it = std::find_if(tasks_.begin(), tasks_.end(), [](const Task &x) { return x.stage == kFormat; });
                    if (it == tasks_.end()) {
                        member = free_member();
                        if (member >= 0) it = std::find_if(tasks_.begin(), tasks_.end(), [](const Task &x) { return x.stage == kQuery; });
                    }
                    if (it == tasks_.end()) it = std::find_if(tasks_.begin(), tasks_.end(), [](const Task &x) { return x.stage == kParse; });
                    if (it != tasks_.end()) {
                        t = *it;
                        tasks_.erase(it);
                        if (t.stage == kQuery) member_busy_[member] = true;
                        break;
                    }
                    if (closed_ && written_ == submitted_) return;
                    cv_.wait(lk);
                }
            }
            if (t.stage == kFormat) {
                format(*t.job);
                std::unique_lock<std::mutex> lk(mu_);
                formatted_[t.job->seq] = t.job;
                cv_.notify_all();
            } else if (t.stage == kParse) {
                parse(*t.job);
                std::unique_lock<std::mutex> lk(mu_);
                // queries leave in block order so that a replica never idles behind a late block
                auto at = std::find_if(tasks_.begin(), tasks_.end(), [&](const Task &x) { return x.stage == kQuery && x.job->seq > t.job->seq; });
                tasks_.insert(at, {t.job, kQuery});
                cv_.notify_all();
            } else {
                const bool ok = query(*t.job, members_[member]);
                {
                    std::unique_lock<std::mutex> lk(mu_);
                    member_busy_[member] = false;
                    cv_.notify_all();
                }
                if (ok && !cfg_.lazy && cfg_.streaming) summarize(*t.job);  // the part of the replay that needs no predictor state
                std::unique_lock<std::mutex> lk(mu_);
                if (ok) queried_[t.job->seq] = t.job;
                cv_.notify_all();
            }
        }
    }
    int free_member() const {
        for (size_t m = 0; m < member_busy_.size(); ++m)
            if (!member_busy_[m]) return (int)m;
        return -1;
    }
    void sequencer() {
        uint64_t next = 0;
        for (;;) {
            Job *job = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return failed_ || queried_.count(next) || (closed_ && next == submitted_); });
                if (failed_ || !queried_.count(next)) return;
                job = queried_[next];
                queried_.erase(next);
            }
            if (!cfg_.lazy) replay(*job);
            {
                std::unique_lock<std::mutex> lk(mu_);
                tasks_.push_front({job, kFormat});
                cv_.notify_all();
            }
            ++next;
        }
    }
    void writer() {
        uint64_t next = 0;
        for (;;) {
            Job *job = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return failed_ || formatted_.count(next) || (closed_ && next == submitted_); });
                if (failed_ || !formatted_.count(next)) return;
                job = formatted_[next];
                formatted_.erase(next);
            }
            if (!job->out.empty()) std::fwrite(job->out.data(), 1, job->out.size(), stdout);
            const size_t units = job->units;
            delete job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                --inflight_;
                inflight_units_ -= units;
                ++written_;
                cv_.notify_all();
            }
            ++next;
        }
    }

    // ms_query's record loop (main.cpp:328-373) over one block of whole records
    void parse(Job &j) {
        fmsi::MemRecordReader reader(j.block.data(), j.block.size());
        std::string name, seq;
        while (reader.next(name, seq) >= 0) layout_record(j.b, name, seq, cfg_.k, cfg_.streaming);
        std::vector<char>().swap(j.block);
    }

    bool query(Job &j, fmsi_gpu_index *idx) {
        Batch &b = j.b;
        const uint64_t n = b.n_results;
        const int gmode = cfg_.mode == fmsi::QueryMode::All ? FMSI_GPU_MODE_ALL : FMSI_GPU_MODE_OR;
        const int gout = cfg_.orders ? FMSI_GPU_OUT_ORDERS : FMSI_GPU_OUT_PRESENCE;
        const int gstr = cfg_.lazy ? FMSI_GPU_STRANDS_LAZY : FMSI_GPU_STRANDS_BOTH;
        void *raw = nullptr;
        if (cfg_.orders) {
            j.raw64.resize(n * (cfg_.lazy ? 1 : 2));
            raw = j.raw64.data();
        } else {
            j.raw8.resize(n);
            raw = j.raw8.data();
        }
        if (n && cfg_.general) {
            const int rc = fmsi_gpu_query_chunks_general(idx, &cfg_.f, b.bases.data(), b.bases.size(), b.chunk_off.data(), b.chunk_len.data(),
                                                         b.res_off.data(), b.chunk_off.size(), n, cfg_.k, j.raw8.data(), FMSI_GPU_MEM_HOST, nullptr);
            if (rc != FMSI_GPU_OK) {
                fail(fmsi_gpu_last_error());
                return false;
            }
            if (!fix_mixed_case_palindromes(j, idx)) return false;
        } else if (n) {
            const int rc = fmsi_gpu_query_chunks(idx, gmode, gout, gstr, cfg_.streaming ? 1 : 0, b.bases.data(), b.bases.size(), b.chunk_off.data(),
                                                 b.chunk_len.data(), b.res_off.data(), b.chunk_off.size(), n, cfg_.k, raw, FMSI_GPU_MEM_HOST, nullptr);
            if (rc != FMSI_GPU_OK) {
                fail(fmsi_gpu_last_error());
                return false;
            }
        }
        std::string().swap(b.bases);  // no longer needed: free early
        return true;
    }

    // The reference decides "self-complementary: count once" by comparing the ASCII k-mer with its reverse complement
    // (AreStringsEqual, fms_index.h:319; ReverseComplementString keeps the case, kmers.h:21-59), so a palindromic k-mer
    // (even k) whose case pattern is not symmetric — `acGT` in a soft-masked query — is counted TWICE there. The
    // kernels compare bases, not case; those k-mers are found here and given the reference's value f(2 ones, 2 total):
    // xor -> 0, and / or unchanged, r-s -> a second query of just those k-mers with ceil(r/2) - floor(s/2).
    bool fix_mixed_case_palindromes(Job &j, fmsi_gpu_index *idx) {
        const int k = cfg_.k;
        if ((k & 1) || (cfg_.f.kind != FMSI_GPU_F_XOR && cfg_.f.kind != FMSI_GPU_F_RANGE)) return true;
        Batch &b = j.b;
        auto comp = [](unsigned char c) -> unsigned char {  // complement of an ACGTacgt letter, upper case
            switch (c & 0xDF) {
            case 'A': return 'T';
            case 'C': return 'G';
            case 'G': return 'C';
            default: return 'A';
            }
        };
        std::vector<uint64_t> off, slot;
        const char *t = b.bases.data();
        for (size_t c = 0; c < b.chunk_off.size(); ++c) {
            const uint64_t nk = b.chunk_len[c] - (uint32_t)k + 1;
            for (uint64_t q = 0; q < nk; ++q) {
                const unsigned char *s = (const unsigned char *)t + b.chunk_off[c] + q;
                bool pal = true, mixed = false;
                for (int i = 0; i < k / 2; ++i) {
                    if ((s[i] & 0xDF) != comp(s[k - 1 - i])) {
                        pal = false;
                        break;
                    }
                    mixed |= ((s[i] ^ s[k - 1 - i]) & 0x20) != 0;
                }
                if (pal && mixed) {
                    off.push_back(b.chunk_off[c] + q);
                    slot.push_back(b.res_off[c] + q);
                }
            }
        }
        if (off.empty()) return true;
        if (cfg_.f.kind == FMSI_GPU_F_XOR) {
            for (uint64_t sl : slot) j.raw8[sl] = 0;
            return true;
        }
        fmsi_gpu_function f2 = cfg_.f;
        f2.r = (cfg_.f.r + 1) / 2;
        f2.s = cfg_.f.s / 2;
        if (f2.r > f2.s) {
            for (uint64_t sl : slot) j.raw8[sl] = 0;
            return true;
        }
        std::vector<uint32_t> len(off.size(), (uint32_t)k);
        std::vector<uint64_t> ro(off.size());
        for (size_t q = 0; q < ro.size(); ++q) ro[q] = q;
        std::vector<uint8_t> res(off.size());
        const int rc = fmsi_gpu_query_chunks_general(idx, &f2, b.bases.data(), b.bases.size(), off.data(), len.data(), ro.data(), off.size(), off.size(),
                                                     k, res.data(), FMSI_GPU_MEM_HOST, nullptr);
        if (rc != FMSI_GPU_OK) {
            fail(fmsi_gpu_last_error());
            return false;
        }
        for (size_t q = 0; q < slot.size(); ++q) j.raw8[slot[q]] = res[q];
        return true;
    }

    // Exact -S mode, on a worker: the merged values of every reference chunk for the unswapped strand order, the
    // predictor inputs for either order, and whether the order changes any value (predictor.hpp: ChunkSummary).
    void summarize(Job &j) {
        Batch &b = j.b;
        const uint64_t n = b.n_results;
        const bool orders = cfg_.orders;
        if (orders) j.fin64.resize(n);
        else j.fin8.resize(n);
        j.summaries.resize(b.ref_chunks.size());
        const int64_t *r64 = j.raw64.data();
        const uint8_t *r8 = j.raw8.data();
        auto f_of = [&](uint64_t q) -> int64_t { return orders ? r64[2 * q] : (int64_t)(r8[q] & 3) - 1; };
        auto r_of = [&](uint64_t q) -> int64_t { return orders ? r64[2 * q + 1] : (int64_t)((r8[q] >> 2) & 3) - 1; };
        uint64_t q0 = 0;
        for (size_t c = 0; c < b.ref_chunks.size(); ++c) {
            const uint32_t m = b.ref_chunks[c];
            fmsi::ChunkSummary &cs = j.summaries[c];
            fmsi::streaming_chunk_with_order(
                false, cfg_.mode, orders, m, [&](size_t q) { return f_of(q0 + q); }, [&](size_t q) { return r_of(q0 + q); },
                [&](size_t q, int64_t v) {
                    if (orders) j.fin64[q0 + q] = v;
                    else j.fin8[q0 + q] = v == 1;
                },
                cs.fpr[0], cs.bpr[0]);
            bool differs = false;
            fmsi::streaming_chunk_with_order(
                true, cfg_.mode, orders, m, [&](size_t q) { return f_of(q0 + q); }, [&](size_t q) { return r_of(q0 + q); },
                [&](size_t q, int64_t v) { differs |= orders ? j.fin64[q0 + q] != v : j.fin8[q0 + q] != (uint8_t)(v == 1); }, cs.fpr[1], cs.bpr[1]);
            cs.differs = differs;
            q0 += m;
        }
    }

    // final values in query order from the both-strand results (exact mode), strictly in block order
    void replay(Job &j) {
        Batch &b = j.b;
        const uint64_t n = b.n_results;
        const bool orders = cfg_.orders;
        const int64_t *r64 = j.raw64.data();
        const uint8_t *r8 = j.raw8.data();
        auto f_of = [&](uint64_t q) -> int64_t { return orders ? r64[2 * q] : (int64_t)(r8[q] & 3) - 1; };
        auto r_of = [&](uint64_t q) -> int64_t { return orders ? r64[2 * q + 1] : (int64_t)((r8[q] >> 2) & 3) - 1; };
        uint64_t q0 = 0;
        if (cfg_.streaming) {
            // per chunk: the predictor picks the order, the summary supplies what that order logs; values were
            // merged for the unswapped order and are redone only where the other order changes them
            for (size_t c = 0; c < b.ref_chunks.size(); ++c) {
                const uint32_t m = b.ref_chunks[c];
                const fmsi::ChunkSummary &cs = j.summaries[c];
                const bool swap = predictor_.predict_swap();
                if (swap && cs.differs) {
                    int fpr, bpr;
                    fmsi::streaming_chunk_with_order(
                        true, cfg_.mode, orders, m, [&](size_t q) { return f_of(q0 + q); }, [&](size_t q) { return r_of(q0 + q); },
                        [&](size_t q, int64_t v) {
                            if (orders) j.fin64[q0 + q] = v;
                            else j.fin8[q0 + q] = v == 1;
                        },
                        fpr, bpr);
                }
                predictor_.log_result(cs.fpr[swap], cs.bpr[swap]);
                q0 += m;
            }
            return;
        }
        if (orders) j.fin64.resize(n);
        else j.fin8.resize(n);
        for (uint32_t m : b.ref_chunks) {
            for (uint32_t q = 0; q < m; ++q) {
                const int64_t v = fmsi::replay_single(predictor_, cfg_.mode, orders, f_of(q0 + q), r_of(q0 + q));
                if (orders) j.fin64[q0 + q] = v;
                else j.fin8[q0 + q] = v == 1;
            }
            q0 += m;
        }
    }

    // text (main.cpp:333, :341-342, :358-372 and fms_index.h:242-253 / :301-309)
    void format(Job &j) {
        Batch &b = j.b;
        const bool orders = cfg_.orders;
        const uint8_t *p8 = cfg_.lazy ? j.raw8.data() : j.fin8.data();
        const int64_t *p64 = cfg_.lazy ? j.raw64.data() : j.fin64.data();
        std::string &out = j.out;
        out.clear();
        out.reserve(b.names.size() + 2 * b.records.size() + (orders ? 8 : 1) * (size_t)b.n_results + 64);
        uint64_t q = 0;
        char num[24];
        for (const Record &rec : b.records) {
            if (!rec.cont_begin) {
                out.append(b.names.data() + rec.name_begin, rec.name_len);
                out.push_back('\t');
            }
            bool comma = rec.comma0;
            for (size_t o = rec.op_begin; o < rec.op_end; ++o) {
                const Op &op = b.ops[o];
                if (!orders) {
                    if (op.kmers) {
                        const size_t at = out.size();
                        out.resize(at + op.count);
                        for (uint64_t t = 0; t < op.count; ++t) out[at + t] = (char)('0' + p8[q + t]);
                        q += op.count;
                    } else {
                        out.append(op.count, '0');
                    }
                } else {
                    for (uint64_t t = 0; t < op.count; ++t) {
                        if (comma) out.push_back(',');
                        comma = true;
                        if (op.kmers) {
                            const int len = fast_itoa(p64[q + t], num);
                            out.append(num, (size_t)len);
                        } else {
                            out.append("-1");
                        }
                    }
                    if (op.kmers) q += op.count;
                }
            }
            if (!rec.cont_end) out.push_back('\n');
        }
    }
    static int fast_itoa(int64_t v, char *buf) {
        if (v < 0) {  // only -1 occurs
            buf[0] = '-';
            buf[1] = '1';
            return 2;
        }
        char tmp[24];
        int n = 0;
        uint64_t u = (uint64_t)v;
        do {
            tmp[n++] = (char)('0' + u % 10);
            u /= 10;
        } while (u);
        for (int t = 0; t < n; ++t) buf[t] = tmp[n - 1 - t];
        return n;
    }

    Config cfg_;
    std::vector<fmsi_gpu_index *> members_;
    std::vector<bool> member_busy_;
    fmsi::StrandPredictor predictor_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Task> tasks_;
    std::map<uint64_t, Job *> queried_, formatted_;
    std::vector<std::thread> threads_;
    size_t max_inflight_ = 4, inflight_ = 0;
    // bound on the work in flight (bytes of unparsed input, or results + text of laid-out batches): with `lookup -S`
    // a result costs ~35 bytes of host buffers, so this keeps the pipeline below ~20 GB whatever the record sizes
    static constexpr size_t kMaxInflightUnits = (size_t)512 << 20;
    size_t inflight_units_ = 0;
    uint64_t submitted_ = 0, written_ = 0;
    bool closed_ = false, failed_ = false;
};

// Reads ahead of the pipeline on its own thread (also while the index is still being loaded): blocks of
// whole records, bounded by `max_bytes` of queued input.
class BlockPrefetcher {
  public:
    BlockPrefetcher(const std::string &path, size_t block_bytes, size_t max_bytes)
        : st_(std::make_shared<State>(path, block_bytes, max_bytes)) {}
    ~BlockPrefetcher() {
        if (!thread_.joinable()) return;
        bool finished;
        {
            std::unique_lock<std::mutex> lk(st_->mu);
            st_->stop = true;
            finished = st_->done;
            st_->cv.notify_all();
        }
        if (finished) thread_.join();
        else thread_.detach();  // may sit in a blocking read of stdin; it owns its state
    }
    void start() {
        std::shared_ptr<State> st = st_;
        thread_ = std::thread([st] {
            std::vector<char> block;
            for (;;) {
                const bool got = st->src.next(block);
                std::unique_lock<std::mutex> lk(st->mu);
                if (!got) {
                    st->done = true;
                    st->cv.notify_all();
                    return;
                }
                st->cv.wait(lk, [&] { return st->stop || st->queued_bytes < st->max_bytes; });
                if (st->stop) {
                    st->done = true;
                    return;
                }
                st->queued_bytes += block.size();
                st->queue.emplace_back(std::move(block));
                block = std::vector<char>();
                st->cv.notify_all();
            }
        });
    }
    bool pop(std::vector<char> &out) {
        std::unique_lock<std::mutex> lk(st_->mu);
        st_->cv.wait(lk, [&] { return st_->done || !st_->queue.empty(); });
        if (st_->queue.empty()) return false;
        out = std::move(st_->queue.front());
        st_->queue.pop_front();
        st_->queued_bytes -= out.size();
        st_->cv.notify_all();
        return true;
    }

  private:
    struct State {
        State(const std::string &path, size_t block_bytes, size_t max_b) : src(path, block_bytes), max_bytes(max_b) {}
        fmsi::BlockSource src;
        std::mutex mu;
        std::condition_variable cv;
        std::deque<std::vector<char>> queue;
        size_t queued_bytes = 0, max_bytes;
        bool done = false, stop = false;
    };
    std::shared_ptr<State> st_;
    std::thread thread_;
};

int ms_query(int argc, char *argv[], bool output_orders) {
    bool usage = false;
    int c;
    int k = 0;
    std::string fn;
    if (argc > 1 && std::string(argv[argc - 1]) != "-h") {  // prefix must be last (main.cpp:244-247)
        fn = argv[argc - 1];
        argc--;
    }
    std::string query_fn = "-";
    std::string f_name = "or";
    bool has_klcp = false;
    while ((c = getopt(argc, argv, "f:hq:k:OS")) >= 0) {
        switch (c) {
        case 'f':
            if (!known_function(optarg)) {
                std::cerr << "ERROR: Function '" << optarg << "' not recognized." << std::endl;
                return usage_query();
            }
            f_name = optarg;
            break;
        case 'h': usage = true; break;
        case 'q': query_fn = optarg; break;
        case 'k': k = atoi(optarg); break;
        case 'O':
            if (f_name != "or") std::cerr << "WARNING: Parameter -O is ignored when parameter -f is specified." << std::endl;
            else f_name = "all";
            break;
        case 'S': has_klcp = true; break;
        default: return usage_query(output_orders);
        }
    }
    if (usage) {
        usage_query(output_orders);
        return 0;
    } else if (fn.empty()) {
        std::cerr << "ERROR: Path to the fasta file is a required argument." << std::endl;
        return usage_query(output_orders);
    }
    if (output_orders && f_name == "all") {
        std::cerr << "WARNING: The current version of FMSI has speed benefits only if output as (minimum) perfect hash function is not used. Additionally, if you desire minimum perfect hash function, please minimize the number of ones in the mask." << std::endl;
    } else if (f_name != "or" && output_orders) {
        std::cerr << "ERROR: FMSI as Minimum Perfect Hash Function is not allowed with f-masked superstrings in the current version." << std::endl;
        return usage_query(output_orders);
    }
    const bool general = f_name != "or" && f_name != "all";

    // $FMSI_GPU_DEVICES = "all" | "0,1,2,..." : shard every batch over replicas on these GPUs
    // (multi-GPU scheduler of the C-ABI); otherwise $FMSI_GPU_DEVICE (default 0) alone.
    int device = 0;
    if (const char *e = std::getenv("FMSI_GPU_DEVICE")) device = atoi(e);
    std::vector<int> devices;
    if (const char *e = std::getenv("FMSI_GPU_DEVICES")) {
        const std::string v = e;
        if (v == "all") {
            for (int dvc = 0; dvc < fmsi_gpu_device_count(); ++dvc) devices.push_back(dvc);
        } else {
            size_t at = 0;
            while (at < v.size()) {
                size_t end = v.find(',', at);
                if (end == std::string::npos) end = v.size();
                if (end > at) devices.push_back(atoi(v.substr(at, end - at).c_str()));
                at = end + 1;
            }
        }
        if (!devices.empty()) device = devices[0];
    }
    if (devices.size() <= 1) {
        // One GPU: hide the others from the CUDA runtime — driver initialisation is paid per visible
        // device (seconds on an 8-GPU box) and a short query run is dominated by it.
        const char *cvd = std::getenv("CUDA_VISIBLE_DEVICES");
        std::string pick;
        if (!cvd) {
            pick = std::to_string(device);
        } else {
            const std::string v = cvd;
            size_t at = 0;
            for (int ord = 0; at <= v.size(); ++ord) {
                size_t end = v.find(',', at);
                if (end == std::string::npos) end = v.size();
                if (ord == device) {
                    pick = v.substr(at, end - at);
                    break;
                }
                at = end + 1;
            }
        }
        if (!pick.empty()) {
            setenv("CUDA_VISIBLE_DEVICES", pick.c_str(), 1);
            device = 0;
        }
    }
    // The query file is opened (and, unless it is stdin, read ahead) while the index loads; a failure to
    // open it is reported where the reference reports it: after the index checks (main.cpp:302-327).
    size_t batch_bases = 16u << 20;
    if (const char *e = std::getenv("FMSI_GPU_BATCH_BASES")) batch_bases = (size_t)atoll(e);
    std::unique_ptr<BlockPrefetcher> input;
    std::string open_error;
    try {
        input.reset(new BlockPrefetcher(query_fn, batch_bases, (size_t)1 << 30));
    } catch (const std::invalid_argument &e) {
        open_error = e.what();
    }
    if (input && query_fn != "-") input->start();
    fmsi_gpu_index *idx = nullptr;
    // The dictionary tier costs ~2 s of build time at human scale and pays off only beyond ~10^10
    // k-mers per run, far above what FASTA parsing can feed: off unless $FMSI_GPU_DICT=1.
    const bool timing = std::getenv("FMSI_GPU_TIMING") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    fmsi_gpu_options opts;
    std::memset(&opts, 0, sizeof opts);
    opts.prefix_t = -1;
    opts.dict = 0;
    opts.multistep = -1;
    opts.locality = -1;  // like the dictionary tiers: the host parser cannot feed it fast enough to repay its build
    int rc = fmsi_gpu_index_load(fn.c_str(), has_klcp ? 1 : 0, device, &opts, &idx);
    if (rc == FMSI_GPU_ERR_IO) {
        std::cerr << "ERROR: index not correctly loaded. Ensure that you correctly call `fmsi index` before." << std::endl;
        return usage_query(output_orders);
    } else if (rc != FMSI_GPU_OK) {
        std::cerr << "ERROR: " << fmsi_gpu_last_error() << std::endl;
        return 1;
    }
    fmsi_gpu_index_info info;
    fmsi_gpu_index_get_info(idx, &info);
    if (has_klcp != (info.has_klcp != 0)) {
        std::cerr << "ERROR: kLCP array was not constructed for the given index. Either construct it again without the `-s` flag or use `query -s` which slows down streaming queries." << std::endl;
        return usage_query(output_orders);
    }
    const int index_k = info.k;
    if (k != 0 && k != index_k) {
        std::cerr << "ERROR: Mismatch. Provided k (" << k << ") does not match the k of the index (" << index_k << ")." << std::endl;
        return usage_query(output_orders);
    }
    if (k == 0) k = index_k;
    if (k < 1 || k > FMSI_GPU_MAX_K) {
        std::cerr << "ERROR: k must be between 1 and " << FMSI_GPU_MAX_K << " (index has k = " << k << ")." << std::endl;
        fmsi_gpu_index_free(idx);
        return 1;
    }

    fmsi_gpu_pool *pool = nullptr;
    std::vector<fmsi_gpu_index *> members{idx};
    if (devices.size() > 1) {
        rc = fmsi_gpu_pool_create(idx, devices.data(), (int)devices.size(), &pool);
        if (rc != FMSI_GPU_OK) {
            std::cerr << "ERROR: " << fmsi_gpu_last_error() << std::endl;
            return 1;
        }
        members.clear();
        for (int m = 0; m < fmsi_gpu_pool_size(pool); ++m) members.push_back(fmsi_gpu_pool_member(pool, m));
    }
    Config cfg;
    cfg.k = k;
    cfg.streaming = has_klcp;
    cfg.orders = output_orders;
    cfg.mode = f_name == "all" ? fmsi::QueryMode::All : fmsi::QueryMode::Or;
    // Plain `fmsi query` (`or` mode, presence output) prints 1 iff either strand has an ON occurrence, whichever strand the
    // predictor tries first (fms_index.h:289-293, :212-233): no replay is needed, one value per k-mer comes back.
    // `-O`, `lookup` and their `-S` forms stop at the first decided strand and do depend on the predictor's history
    // (SURVEY 8a row P): they get both strands and the exact replay unless $FMSI_GPU_STRANDS=lazy says the index makes it moot.
    cfg.lazy = !output_orders && f_name == "or";
    if (const char *e = std::getenv("FMSI_GPU_STRANDS")) cfg.lazy = std::string(e) == "lazy";
    if (general) {  // mask_function(), src/functions.h:23-57
        cfg.general = true;
        cfg.lazy = true;        // one value per k-mer, nothing to replay
        cfg.streaming = false;  // query_kmers() ignores kLCP in general mode (fms_index.h:337)
        if (f_name == "and") cfg.f.kind = FMSI_GPU_F_AND;
        else if (f_name == "xor") cfg.f.kind = FMSI_GPU_F_XOR;
        else {
            const size_t dash = f_name.find('-', 1);
            cfg.f.kind = FMSI_GPU_F_RANGE;
            cfg.f.r = std::stoi(f_name.substr(0, dash));
            cfg.f.s = std::stoi(f_name.substr(dash + 1));
        }
    }
    int workers = (int)std::min<unsigned>(8, std::max(2u, std::thread::hardware_concurrency() / 2));
    if (const char *e = std::getenv("FMSI_GPU_THREADS")) workers = std::max(1, atoi(e));
    workers = std::max<int>(workers, (int)members.size());

    if (timing) std::cerr << "[fmsi timing] index load + replicas: " << since(t_start) << " s" << std::endl;
    const auto t_query = std::chrono::steady_clock::now();
    bool ok;
    if (!input) throw std::invalid_argument(open_error);  // uncaught in the reference too (parser.h:21-23)
    if (query_fn == "-") input->start();
    {
        Pipeline pipe(cfg, members, workers);
        // A block beyond `giant` bytes holds a record longer than a batch should be (a chromosome, an assembly): the
        // reader lays it out itself, in pieces of ~`piece` results that enter the pipeline as they are cut, so that
        // result buffers stay bounded however long the record is (the text itself is held once, as in the reference).
        size_t giant = (size_t)64 << 20;
        uint64_t piece = (uint64_t)32 << 20;
        if (const char *e = std::getenv("FMSI_GPU_GIANT_BLOCK")) giant = (size_t)atoll(e);
        if (const char *e = std::getenv("FMSI_GPU_PIECE_RESULTS")) piece = (uint64_t)atoll(e);
        Job *job = new Job();
        while (input->pop(job->block)) {
            if (job->block.size() <= giant) {
                pipe.submit(job);
                job = new Job();
                continue;
            }
            fmsi::MemRecordReader reader(job->block.data(), job->block.size());
            std::string name, seq;
            Batch cur;
            auto emit = [&](Batch &b) {
                Job *pj = new Job();
                pj->b = std::move(b);
                b = Batch();
                pipe.submit_parsed(pj);
            };
            while (reader.next(name, seq) >= 0) {
                layout_record(cur, name, seq, cfg.k, cfg.streaming, piece, emit);
                if (cur.n_results >= piece) emit(cur);
            }
            if (!cur.records.empty()) emit(cur);
            std::vector<char>().swap(job->block);
        }
        delete job;
        if (timing) std::cerr << "[fmsi timing] reader done: " << since(t_query) << " s" << std::endl;
        ok = pipe.finish();
        if (timing) std::cerr << "[fmsi timing] pipeline drained: " << since(t_query) << " s" << std::endl;
    }
    input.reset();
    std::fflush(stdout);
    if (timing) std::cerr << "[fmsi timing] total: " << since(t_start) << " s" << std::endl;
    // Everything is written: leave without tearing down the device state piece by piece (freeing a human-scale index
    // and destroying the CUDA context cost 0.2-0.3 s, a quarter of a short run; the driver reclaims both at exit).
    if (!std::getenv("FMSI_GPU_CLEAN_EXIT")) {
        std::cerr.flush();
        _exit(ok ? 0 : 1);
    }
    fmsi_gpu_pool_free(pool);
    fmsi_gpu_index_free(idx);
    return ok ? 0 : 1;
}

int usage_index() {
    std::cerr << std::endl;
    std::cerr << "Usage:   fmsi index [options] <masked-superstring-input>" << std::endl << std::endl;
    std::cerr << "Options:" << std::endl;
    std::cerr << "    -k INT  - size of k-mers [recommended, default: number of mask trailing zeros - 1]" << std::endl;
    std::cerr << "    -x      - do not compute the kLCP array used for faster streaming queries." << std::endl << std::endl;
    std::cerr << "Note: `fmsi index` accepts only masked superstrings - these can be computed e.g. by KmerCamel from any FASTA file." << std::endl
              << std::endl;
    return 1;
}

// `fmsi index [-k K] [-x] MS.fa` — ms_index, reference src/main.cpp:172-236: same flags, messages and files
// (the suffix array is unique, so the GPU builder's `.fmsi.*` bytes equal the reference's; tests/test_gpu_build.py).
// Returns -2 for inputs the GPU builder does not take (n + 1 >= 2^32) when a reference binary can be forwarded to.
int ms_index(int argc, char *argv[]) {
    bool usage = false;
    std::string fn;
    if (argc > 1 && std::string(argv[argc - 1]) != "-h") {
        fn = argv[argc - 1];
        argc--;
    }
    int k = 0, c;
    bool no_streaming = false;
    while ((c = getopt(argc, argv, "hk:x")) >= 0) {
        switch (c) {
        case 'h': usage = true; break;
        case 'k': k = atoi(optarg); break;
        case 'x': no_streaming = true; break;
        default: return usage_index();
        }
    }
    if (usage) {
        usage_index();
        return 0;
    } else if (fn.empty()) {
        std::cerr << "ERROR: Path to the masked superstring is a required argument." << std::endl;
        return usage_index();
    }
    std::cerr << "Starting " << fn << std::endl;
    // read_masked_superstring (parser.h:41-55): the first FASTA/FASTQ entry of a plain or gzip file / stdin
    std::string ms, name;
    {
        FILE *in = fn == "-" ? stdin : std::fopen(fn.c_str(), "r");
        if (!in) throw std::invalid_argument("couldn't open file " + fn);  // uncaught in the reference too
        gzFile fp = gzdopen(fileno(in), "r");
        gzbuffer(fp, 1 << 20);
        std::vector<char> text;
        size_t have = 0;
        for (;;) {
            if (text.size() < have + (1u << 24)) text.resize(std::max<size_t>(2 * text.size(), (size_t)1 << 24));
            const int got = gzread(fp, text.data() + have, (unsigned)std::min<size_t>(text.size() - have, 1u << 30));
            if (got <= 0) break;
            have += (size_t)got;
        }
        gzclose(fp);
        fmsi::MemRecordReader reader(text.data(), have);
        const int64_t l = reader.next(name, ms);
        if (l < 0)
            throw std::invalid_argument("Error reading the fasta file. The fasta file should contain a single entry - the masked superstring.");
        std::string other;
        if (reader.next(name, other) >= 0)
            std::cerr << "Warning: The fasta file contains more than one entry. Only the first entry will be used." << std::endl;
    }
    if (ms.size() == 0) {
        std::cerr << "ERROR: The file '" << fn << "' is in incorrect format. It is supposed to be a fasta file with a single entry, the masked superstring"
                  << std::endl;
        return usage_index();
    }
    std::cerr << "Read masked superstring of length " << ms.size() << std::endl;
    int inferred_k = 1;  // infer_k (parser.h:31-37)
    while ((size_t)inferred_k < ms.size() && !(ms[ms.size() - inferred_k] >= 'A' && ms[ms.size() - inferred_k] <= 'Z')) inferred_k++;
    if (k == 0) {
        k = inferred_k;
        std::cerr << "Inferred k from the masked case convention: " << k << std::endl;
    }
    if (k != inferred_k)
        std::cerr << "WARNING: The provided k (" << k << ") does not match the k inferred from the mask convention (" << inferred_k
                  << "). The provided k is used but we recommend double checking that it is correct." << std::endl;
    if (k > 64 && !no_streaming) {
        std::cerr << "WARNING: Construction of kLCP array for streaming support is only available for k <= 64. The index will be constructed "
                     "without streaming support, which results in slower positive streaming queries."
                  << std::endl;
        no_streaming = true;
    }
    const char *ref_bin = std::getenv("FMSI_REFERENCE_BIN");
    const bool can_forward = ref_bin && *ref_bin;
    if (ms.size() + 1 >= (1ull << 32) && can_forward) return -2;
    int device = 0;
    if (const char *e = std::getenv("FMSI_GPU_DEVICE")) device = atoi(e);
    fmsi_gpu_options opts;
    std::memset(&opts, 0, sizeof opts);  // no suffix table, no dictionary: the index is only written out
    fmsi_gpu_index *idx = nullptr;
    int rc = fmsi_gpu_index_build(ms.data(), ms.size(), k, no_streaming ? 0 : 1, FMSI_GPU_MEM_HOST, device, &opts, &idx);
    if (rc != FMSI_GPU_OK) {
        std::cerr << "ERROR: " << fmsi_gpu_last_error() << std::endl;
        return can_forward ? -2 : 1;
    }
    std::cerr << "Constructed index" << std::endl;
    rc = fmsi_gpu_index_save(idx, fn.c_str());
    fmsi_gpu_index_free(idx);
    if (rc != FMSI_GPU_OK) {
        std::cerr << "ERROR: " << fmsi_gpu_last_error() << std::endl;
        return 1;
    }
    std::cerr << "Written index" << std::endl;
    return 0;
}

}  // namespace

int main(int argc, char *argv[]) {
    // blocks, layouts and output buffers are tens of MB each and short-lived: keep them in the heap
    // instead of mmap/munmap + page faults for every one of them
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, -1);
    if (argc < 2) return usage();
    const std::string op = argv[1];
    int ret;
    if (op == "query" || op == "lookup") {
        ret = ms_query(argc - 1, argv + 1, op == "lookup");
        if (ret == -2) return forward_to_reference(argc, argv);
        return ret;
    }
    if (op == "index") {
        ret = ms_index(argc - 1, argv + 1);
        if (ret == -2) return forward_to_reference(argc, argv);
        return ret;
    }
    if (op == "-v") {
        std::cout << kVersion << std::endl;
        return 0;
    }
    if (op == "-h") {
        usage();
        return 0;
    }
    if (op == "clean" || op == "merge" || op == "normalize" || op == "compact" || op == "export" || op == "union" ||
        op == "inter" || op == "diff" || op == "symdiff")
        return forward_to_reference(argc, argv);
    return usage();
}
