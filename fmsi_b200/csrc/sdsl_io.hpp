// sdsl_io.hpp — reader/writer for the sdsl-lite 2.1.0 on-disk forms FMSI's index files use,
// written from the format description, without linking sdsl.
//
//   .ac_gt / .ac / .gt / .klcp : int_vector<1>   = u64 bit_len, ceil(bit_len/64) LE u64 words
//                                (reference: sdsl int_vector.hpp:593-610 header, :1563-1595 data)
//   .mask                      : rrr_vector<63>  = u64 size, bt (int_vector<0>: u64 bit_len, u8 width,
//                                words), btnr (int_vector<1>), btnrp, rank (int_vector<0>), invert
//                                (int_vector<1>)   (reference: sdsl rrr_vector.hpp:350-373)
//   .misc                      : text: dollar_position, counts[0..3], k  (fms_index.h:493-499)
//
// The RRR<63> block code (class = popcount, offset = rank of the 63-bit word inside its class in
// the combinatorial number system) is decoded ONCE at load into plain bits; the GPU never sees RRR.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <functional>
#include <string>
#include <thread>
#include <vector>

namespace fmsi {

struct BitVec {
    uint64_t nbits = 0;
    std::vector<uint64_t> w;  // ceil(nbits/64)+1 words, zero padded
    bool get(uint64_t p) const { return (w[p >> 6] >> (p & 63)) & 1; }
    void set(uint64_t p) { w[p >> 6] |= 1ull << (p & 63); }
    void resize_bits(uint64_t n) {
        nbits = n;
        w.assign(((n + 63) >> 6) + 1, 0);
    }
    // len <= 64 bits starting at bit pos, least significant first
    uint64_t get_int(uint64_t pos, unsigned len) const {
        if (len == 0) return 0;
        uint64_t wi = pos >> 6;
        unsigned off = pos & 63;
        uint64_t x = w[wi] >> off;
        if (off + len > 64) x |= w[wi + 1] << (64 - off);
        if (len < 64) x &= (1ull << len) - 1;
        return x;
    }
    // ORs the low `len` (<= 64) bits of x into bits [pos, pos + len)
    void set_int(uint64_t pos, uint64_t x, unsigned len) {
        if (len == 0) return;
        if (len < 64) x &= (1ull << len) - 1;
        const uint64_t wi = pos >> 6;
        const unsigned off = pos & 63;
        w[wi] |= x << off;
        if (off + len > 64) w[wi + 1] |= x >> (64 - off);
    }
    // the same, for writers on different threads whose fields may share a word
    void set_int_atomic(uint64_t pos, uint64_t x, unsigned len) {
        if (len == 0) return;
        if (len < 64) x &= (1ull << len) - 1;
        const uint64_t wi = pos >> 6;
        const unsigned off = pos & 63;
        __atomic_fetch_or(&w[wi], x << off, __ATOMIC_RELAXED);
        if (off + len > 64) __atomic_fetch_or(&w[wi + 1], x >> (64 - off), __ATOMIC_RELAXED);
    }
};

// fn(begin, end) over [0, n) cut into contiguous ranges whose boundaries are multiples of `granule`, one host
// thread per range ($FMSI_GPU_THREADS, else the machine's hardware threads, at most 32). Exceptions propagate.
inline void parallel_ranges(uint64_t n, uint64_t granule, const std::function<void(uint64_t, uint64_t)> &fn) {
    unsigned T = std::thread::hardware_concurrency();
    if (const char *e = std::getenv("FMSI_GPU_THREADS")) T = (unsigned)std::atoi(e);
    if (T < 1) T = 1;
    if (T > 32) T = 32;
    const uint64_t units = (n + granule - 1) / granule;
    if (T > units) T = units ? (unsigned)units : 1;
    if (T == 1) {
        fn(0, n);
        return;
    }
    std::vector<std::thread> th;
    std::vector<std::string> err(T);
    for (unsigned t = 0; t < T; ++t) {
        const uint64_t a = std::min(n, units * t / T * granule), b = std::min(n, units * (t + 1) / T * granule);
        th.emplace_back([&, t, a, b] {
            try {
                if (a < b) fn(a, b);
            } catch (const std::exception &e) {
                err[t] = e.what();
            }
        });
    }
    for (auto &x : th) x.join();
    for (const std::string &e : err)
        if (!e.empty()) throw std::runtime_error(e);
}

// A file that holds exactly one int_vector<1> (.ac_gt / .ac / .gt / .klcp): the words are read straight into the
// vector (no intermediate copy of the file), so several files can be read by concurrent threads at disk / page
// cache speed.
inline void read_bitvec_file(const std::string &path, BitVec &b) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    uint64_t nbits = 0;
    bool ok = std::fread(&nbits, 8, 1, f) == 1;
    if (ok) {
        std::fseek(f, 0, SEEK_END);
        const uint64_t size = (uint64_t)std::ftell(f);
        const uint64_t n_words = (nbits >> 6) + ((nbits & 63) ? 1 : 0);
        ok = size >= 8 && (size - 8) / 8 >= n_words;
        if (ok) {
            std::fseek(f, 8, SEEK_SET);
            b.nbits = nbits;
            b.w.resize(n_words + 1);
            b.w[n_words] = 0;
            ok = n_words == 0 || std::fread(b.w.data(), 8, n_words, f) == n_words;
        }
    }
    std::fclose(f);
    if (!ok) throw std::runtime_error("truncated file " + path);
}

struct IntVec {
    BitVec bits;
    unsigned width = 0;
    uint64_t n = 0;
    uint64_t get(uint64_t i) const { return bits.get_int(i * width, width); }
    void alloc(uint64_t n_, unsigned width_) {
        n = n_;
        width = width_;
        bits.resize_bits(n_ * width_);
    }
    void set(uint64_t i, uint64_t x) { bits.set_int(i * width, x, width); }
};

class ByteReader {
  public:
    explicit ByteReader(const std::string &path) {
        FILE *f = std::fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("cannot open " + path);
        std::fseek(f, 0, SEEK_END);
        long sz = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        buf_.resize((size_t)sz);
        if (sz > 0 && std::fread(buf_.data(), 1, (size_t)sz, f) != (size_t)sz) {
            std::fclose(f);
            throw std::runtime_error("short read on " + path);
        }
        std::fclose(f);
        path_ = path;
    }
    uint64_t u64() {
        need(8);
        uint64_t v;
        std::memcpy(&v, buf_.data() + pos_, 8);
        pos_ += 8;
        return v;
    }
    unsigned u8() {
        need(1);
        return buf_[pos_++];
    }
    void words(uint64_t *dst, uint64_t n) {
        need(n * 8);
        std::memcpy(dst, buf_.data() + pos_, n * 8);
        pos_ += n * 8;
    }
    bool at_end() const { return pos_ == buf_.size(); }
    void read_bitvec(BitVec &b) {
        uint64_t nbits = u64();
        b.resize_bits(nbits);
        words(b.w.data(), (nbits + 63) >> 6);
    }
    void read_intvec(IntVec &v) {
        uint64_t nbits = u64();
        unsigned width = u8();
        if (width == 0 || width > 64) throw std::runtime_error("bad int_vector width in " + path_);
        v.width = width;
        v.n = nbits / width;
        v.bits.resize_bits(nbits);
        words(v.bits.w.data(), (nbits + 63) >> 6);
    }

  private:
    void need(uint64_t n) {
        if (pos_ + n > buf_.size()) throw std::runtime_error("truncated file " + path_);
    }
    std::vector<uint8_t> buf_;
    size_t pos_ = 0;
    std::string path_;
};

// ---- RRR<63>, samples every 32 blocks -----------------------------------------------------------
constexpr unsigned kRrrBlock = 63;
constexpr unsigned kRrrSample = 32;

struct Binomials {
    uint64_t c[65][65];
    unsigned space[64];  // bits of the offset field for a class (0 for the two uniform classes)
    Binomials() {
        for (int n = 0; n <= 64; ++n)
            for (int k = 0; k <= 64; ++k) c[n][k] = 0;
        for (int n = 0; n <= 64; ++n) {
            c[n][0] = 1;
            for (int k = 1; k <= n; ++k) c[n][k] = c[n - 1][k - 1] + (k <= n - 1 ? c[n - 1][k] : 0);
        }
        for (unsigned k = 0; k <= kRrrBlock; ++k) {
            uint64_t v = c[kRrrBlock][k];
            space[k] = (v == 1) ? 0 : (64 - (unsigned)__builtin_clzll(v));
        }
    }
    static const Binomials &get() {
        static Binomials b;
        return b;
    }
};

// Expand one block: `ones` set bits, `nr` = index inside the class. Bit t of the result is
// position t of the block. (Follows the coder described by sdsl rrr_helper.hpp:304-320 / :374-407.)
inline uint64_t rrr_unrank(unsigned ones, uint64_t nr) {
    if (ones == 0) return 0;
    if (ones == kRrrBlock) return (1ull << kRrrBlock) - 1;
    const Binomials &B = Binomials::get();
    uint64_t word = 0;
    unsigned left = kRrrBlock;
    for (unsigned pos = 0; pos < kRrrBlock && ones > 0; ++pos, --left) {
        uint64_t below = B.c[left - 1][ones];  // words of this class whose bit `pos` is 0
        if (nr >= below) {
            word |= 1ull << pos;
            nr -= below;
            --ones;
        }
    }
    return word;
}
inline uint64_t rrr_rank_of_word(uint64_t word) {
    const Binomials &B = Binomials::get();
    unsigned ones = (unsigned)__builtin_popcountll(word);
    if (ones == 0 || ones == kRrrBlock) return 0;
    uint64_t nr = 0;
    unsigned left = kRrrBlock;
    for (unsigned pos = 0; pos < kRrrBlock && ones > 0; ++pos, --left) {
        if ((word >> pos) & 1) {
            nr += B.c[left - 1][ones];
            --ones;
        }
    }
    return nr;
}

struct RrrFile {
    uint64_t size = 0;
    IntVec bt;
    BitVec btnr;
    IntVec btnrp;
    IntVec rank;
    BitVec invert;
};

inline RrrFile read_rrr(const std::string &path) {
    ByteReader r(path);
    RrrFile f;
    f.size = r.u64();
    r.read_intvec(f.bt);
    r.read_bitvec(f.btnr);
    r.read_intvec(f.btnrp);
    r.read_intvec(f.rank);
    r.read_bitvec(f.invert);
    if (!r.at_end()) throw std::runtime_error("trailing bytes in " + path);
    return f;
}

// Decode the whole vector into plain bits. Also verifies the stored samples (btnrp, rank) and the
// total, so a corrupt .mask is rejected at load instead of producing wrong answers.
inline BitVec rrr_decode_all(const RrrFile &f) {
    const Binomials &B = Binomials::get();
    BitVec out;
    out.resize_bits(f.size);
    uint64_t nblocks = (f.size + kRrrBlock) / kRrrBlock;
    if (f.bt.n < nblocks) throw std::runtime_error("rrr: bt array too short");
    uint64_t off = 0, ones_total = 0;
    for (uint64_t b = 0; b < nblocks; ++b) {
        uint64_t sb = b / kRrrSample;
        uint64_t pos0 = b * kRrrBlock;
        if (pos0 >= f.size) break;  // trailing dummy block when size % 63 == 0 (sdsl writes no sample for it)
        if (b % kRrrSample == 0) {
            if (sb >= f.btnrp.n || f.btnrp.get(sb) != off) throw std::runtime_error("rrr: btnrp sample mismatch");
            if (sb >= f.rank.n || f.rank.get(sb) != ones_total) throw std::runtime_error("rrr: rank sample mismatch");
        }
        unsigned stored = (unsigned)f.bt.get(b);
        unsigned ones = f.invert.get(sb) ? kRrrBlock - stored : stored;
        unsigned len = B.space[stored];
        uint64_t nr = f.btnr.get_int(off, len);
        off += len;
        uint64_t word = rrr_unrank(ones, nr);
        unsigned valid = (unsigned)std::min<uint64_t>(kRrrBlock, f.size - pos0);
        if (valid < kRrrBlock && (word >> valid)) throw std::runtime_error("rrr: padding bits set");
        ones_total += (unsigned)__builtin_popcountll(word);
        for (unsigned t = 0; t < valid; ++t)
            if ((word >> t) & 1) out.set(pos0 + t);
    }
    if (f.rank.n == 0 || f.rank.get(f.rank.n - 1) != ones_total) throw std::runtime_error("rrr: total mismatch");
    return out;
}


// ---- writers ----------------------------------------------------------------------------------
class ByteWriter {
  public:
    explicit ByteWriter(const std::string &path) : path_(path) {
        f_ = std::fopen(path.c_str(), "wb");
        if (!f_) throw std::runtime_error("cannot create " + path);
    }
    ~ByteWriter() {
        if (f_) std::fclose(f_);
    }
    void bytes(const void *p, size_t n) {
        if (n && std::fwrite(p, 1, n, f_) != n) throw std::runtime_error("short write on " + path_);
    }
    void u64(uint64_t v) { bytes(&v, 8); }
    void u8(uint8_t v) { bytes(&v, 1); }
    void bitvec(const BitVec &b) {
        u64(b.nbits);
        bytes(b.w.data(), ((b.nbits + 63) >> 6) * 8);
    }
    void intvec(const IntVec &v) {
        u64(v.bits.nbits);
        u8((uint8_t)v.width);
        bytes(v.bits.w.data(), ((v.bits.nbits + 63) >> 6) * 8);
    }

  private:
    FILE *f_ = nullptr;
    std::string path_;
};

inline unsigned bit_length_or_one(uint64_t x) { return x ? 64 - (unsigned)__builtin_clzll(x) : 1; }  // sdsl bits::hi(x) + 1

// Encode a plain bit vector as sdsl's rrr_vector<63> would (constructor, sdsl rrr_vector.hpp:150-250):
// 63-bit blocks stored as (class, offset in class); every 32 blocks one sample of the offset-stream
// position and of the running rank; a full superblock in which more than half of the blocks have
// more than 31 ones is stored complemented (classes only) and flagged in `invert`.
// Two passes over superblocks on all host threads (a human-scale mask has 49 M blocks): classes and per-superblock
// totals first, then — once the prefix sums give every superblock its offset-stream position and rank — the fields.
inline RrrFile rrr_encode(const BitVec &bv) {
    const Binomials &B = Binomials::get();
    RrrFile f;
    f.size = bv.nbits;
    const uint64_t nblocks = (bv.nbits + kRrrBlock) / kRrrBlock;  // incl. a dummy block when size % 63 == 0
    const uint64_t nsuper = (nblocks + kRrrSample - 1) / kRrrSample;
    std::vector<uint8_t> cls(nblocks, 0);
    std::vector<uint64_t> sb_off(nsuper + 1, 0), sb_ones(nsuper + 1, 0);  // per superblock, then exclusive prefix sums
    auto block_word = [&](uint64_t b) {
        const uint64_t pos = b * kRrrBlock;
        return bv.get_int(pos, (unsigned)std::min<uint64_t>(kRrrBlock, bv.nbits - pos));
    };
    // ranges are multiples of 64 superblocks: bt fields (32 x 6 bits = 3 words per superblock) and invert bits of
    // different threads then never share a word
    parallel_ranges(nsuper, 64, [&](uint64_t s0, uint64_t s1) {
        for (uint64_t sb = s0; sb < s1; ++sb) {
            const uint64_t b0 = sb * kRrrSample, b1 = std::min<uint64_t>(b0 + kRrrSample, nblocks);
            uint64_t bits = 0, ones = 0;
            for (uint64_t b = b0; b < b1 && b * kRrrBlock < bv.nbits; ++b) {
                cls[b] = (uint8_t)__builtin_popcountll(block_word(b));
                ones += cls[b];
                bits += B.space[cls[b]];
            }
            sb_off[sb] = bits;
            sb_ones[sb] = ones;
        }
    });
    uint64_t total_ones = 0, stream_bits = 0;
    for (uint64_t sb = 0; sb < nsuper; ++sb) {
        const uint64_t bits = sb_off[sb], ones = sb_ones[sb];
        sb_off[sb] = stream_bits;
        sb_ones[sb] = total_ones;
        stream_bits += bits;
        total_ones += ones;
    }
    f.bt.alloc(nblocks, 6);
    f.btnr.resize_bits(std::max<uint64_t>(stream_bits, 64));
    f.btnrp.alloc(nsuper, bit_length_or_one(stream_bits));
    f.rank.alloc(nsuper + ((bv.nbits % ((uint64_t)kRrrSample * kRrrBlock)) > 0 ? 1 : 0), bit_length_or_one(total_ones));
    f.invert.resize_bits(nsuper);
    parallel_ranges(nsuper, 64, [&](uint64_t s0, uint64_t s1) {
        for (uint64_t sb = s0; sb < s1; ++sb) {
            const uint64_t b0 = sb * kRrrSample, b1 = std::min<uint64_t>(b0 + kRrrSample, nblocks);
            if (b0 * kRrrBlock >= bv.nbits) break;  // superblock made only of the dummy block
            bool inv = false;
            if (b0 + kRrrSample <= nblocks) {  // only complete superblocks may be complemented
                unsigned heavy = 0;
                for (uint64_t b = b0; b < b1; ++b) heavy += cls[b] > kRrrBlock / 2;
                inv = heavy > kRrrSample / 2;
            }
            if (inv) f.invert.set(sb);
            uint64_t off = sb_off[sb];
            for (uint64_t b = b0; b < b1; ++b) {
                const unsigned stored = inv ? kRrrBlock - cls[b] : cls[b];
                f.bt.set(b, stored);
                if (b * kRrrBlock >= bv.nbits) continue;  // dummy block: class only
                const unsigned len = B.space[stored];
                if (len) f.btnr.set_int_atomic(off, rrr_rank_of_word(block_word(b)), len);
                off += len;
            }
        }
    });
    for (uint64_t sb = 0; sb < nsuper; ++sb) {  // samples: narrow fields that share words, written by one thread
        if (sb * kRrrSample * kRrrBlock >= bv.nbits) break;  // no sample for a superblock made only of the dummy block
        f.btnrp.set(sb, sb_off[sb]);
        f.rank.set(sb, sb_ones[sb]);
    }
    f.rank.set(f.rank.n - 1, total_ones);
    return f;
}

inline void write_rrr(const std::string &path, const RrrFile &f) {
    ByteWriter w(path);
    w.u64(f.size);
    w.intvec(f.bt);
    w.bitvec(f.btnr);
    w.intvec(f.btnrp);
    w.intvec(f.rank);
    w.bitvec(f.invert);
}

}  // namespace fmsi
