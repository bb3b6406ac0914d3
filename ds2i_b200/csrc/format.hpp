// Host-side reader of ds2i's on-disk formats (dependency-free C++17; no Boost, no succinct).
//
// Restates the layout produced by succinct::mapper::freeze (succinct/mapper.hpp:51-98): u64 flags,
// then members in map() order; PODs raw, mappable_vector<T> as u64 size + payload, non-PODs
// recursively; NO alignment padding anywhere, so every u64 is read with memcpy.
//   block_freq_index::map  block_freq_index.hpp:124-134
//   freq_index::map        freq_index.hpp:234-243, bitvector_collection.hpp:76-84
//   wand_data::map         wand_data.hpp:71-78
//   global_parameters::map global_parameters.hpp:14-24 (5 single bytes)
//   bit_vector::map        succinct/bit_vector.hpp:228-232 (u64 bits, then mappable_vector<u64>)
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace ds2i_gpu {

struct format_error : std::runtime_error {
    explicit format_error(std::string const& s) : std::runtime_error(s) {}
};

class byte_reader {
public:
    byte_reader(const uint8_t* p, size_t n) : m_p(p), m_n(n), m_pos(0) {}
    template <typename T> T get() {
        need(sizeof(T));
        T v; std::memcpy(&v, m_p + m_pos, sizeof(T)); m_pos += sizeof(T);
        return v;
    }
    const uint8_t* take(size_t bytes) {
        need(bytes);
        const uint8_t* r = m_p + m_pos; m_pos += bytes;
        return r;
    }
    size_t pos() const { return m_pos; }
    size_t remaining() const { return m_n - m_pos; }
private:
    void need(size_t b) const {
        if (b > m_n - m_pos) throw format_error("truncated file (need " + std::to_string(b) + " bytes at " + std::to_string(m_pos) + ")");
    }
    const uint8_t* m_p; size_t m_n, m_pos;
};

struct global_params {
    uint8_t ef_log_sampling0, ef_log_sampling1, rb_log_rank1_sampling, rb_log_sampling1, log_partition_size;
};

inline global_params read_params(byte_reader& r) {
    global_params p;
    p.ef_log_sampling0 = r.get<uint8_t>(); p.ef_log_sampling1 = r.get<uint8_t>();
    p.rb_log_rank1_sampling = r.get<uint8_t>(); p.rb_log_sampling1 = r.get<uint8_t>();
    p.log_partition_size = r.get<uint8_t>();
    return p;
}

// a succinct::bit_vector as it sits in the file: words are NOT 8-byte aligned there
struct bitvec_view {
    uint64_t bits = 0;
    uint64_t nwords = 0;
    const uint8_t* raw = nullptr;   // nwords * 8 bytes, unaligned
    uint64_t word(uint64_t i) const { uint64_t w; std::memcpy(&w, raw + 8 * i, 8); return w; }
};

inline bitvec_view read_bitvec(byte_reader& r) {
    bitvec_view v;
    v.bits = r.get<uint64_t>();
    v.nwords = r.get<uint64_t>();
    if (v.nwords > r.remaining() / 8) throw format_error("bit vector longer than file");
    if (v.bits > v.nwords * 64) throw format_error("bit vector size mismatch");
    v.raw = r.take(size_t(v.nwords) * 8);
    return v;
}

inline uint32_t msb64(uint64_t x) { return 63u - uint32_t(__builtin_clzll(x)); }
inline uint64_t ceil_log2_u64(uint64_t x) { return x > 1 ? msb64(x - 1) + 1 : 0; }   // util.hpp:30-33

// compact_elias_fano::offsets (compact_elias_fano.hpp:14-61)
struct ef_offsets {
    uint64_t universe, n, lower_bits, mask, higher_bits_length, pointer_size, pointers0, pointers1;
    uint64_t pointers0_offset, pointers1_offset, higher_bits_offset, lower_bits_offset, end;
    ef_offsets(uint64_t base, uint64_t universe_, uint64_t n_, uint32_t log_sampling0, uint32_t log_sampling1)
        : universe(universe_), n(n_)
    {
        lower_bits = universe > n ? msb64(universe / n) : 0;
        mask = (uint64_t(1) << lower_bits) - 1;
        higher_bits_length = n + (universe >> lower_bits) + 2;
        pointer_size = ceil_log2_u64(higher_bits_length);
        pointers0 = (higher_bits_length - n) >> log_sampling0;
        pointers1 = n >> log_sampling1;
        pointers0_offset = base;
        pointers1_offset = pointers0_offset + pointers0 * pointer_size;
        higher_bits_offset = pointers1_offset + pointers1 * pointer_size;
        lower_bits_offset = higher_bits_offset + higher_bits_length;
        end = lower_bits_offset + n * lower_bits;
    }
};

// bits [pos, pos+len) of a bit vector, len <= 64, LSB-first (succinct/bit_vector.hpp:251-268)
inline uint64_t get_bits(bitvec_view const& bv, uint64_t pos, uint32_t len) {
    if (!len) return 0;
    uint64_t block = pos >> 6, shift = pos & 63;
    uint64_t m = len == 64 ? ~uint64_t(0) : ((uint64_t(1) << len) - 1);
    if (shift + len <= 64) return (bv.word(block) >> shift) & m;
    return ((bv.word(block) >> shift) | (bv.word(block + 1) << (64 - shift))) & m;
}

// All n values of an Elias-Fano sequence in order (element i: high-bit position (v>>l)+i+1 set,
// low bits at lower_bits_offset + i*l; compact_elias_fano.hpp:105-118).
inline std::vector<uint64_t> ef_decode_all(bitvec_view const& bv, uint64_t base, uint64_t universe, uint64_t n,
                                           global_params const& p) {
    std::vector<uint64_t> out(n);
    if (!n) return out;
    ef_offsets of(base, universe, n, p.ef_log_sampling0, p.ef_log_sampling1);
    if (of.end > bv.bits) throw format_error("Elias-Fano sequence exceeds its bit vector");
    uint64_t pos = of.higher_bits_offset;           // absolute bit cursor in the high bits
    uint64_t w = bv.word(pos >> 6) & (~uint64_t(0) << (pos & 63));
    uint64_t wi = pos >> 6;
    for (uint64_t i = 0; i < n; ++i) {
        while (!w) {
            ++wi;
            if (wi >= bv.nwords) throw format_error("Elias-Fano high bits run past the end");
            w = bv.word(wi);
        }
        uint64_t bit = wi * 64 + uint64_t(__builtin_ctzll(w));
        w &= w - 1;
        uint64_t high = bit - of.higher_bits_offset - i - 1;
        uint64_t low = get_bits(bv, of.lower_bits_offset + i * of.lower_bits, uint32_t(of.lower_bits));
        out[i] = (high << of.lower_bits) | low;
    }
    return out;
}

// TightVariableByte::decode of one value (block_codecs.hpp:84-98): 7-bit groups LSB first, the LAST byte has bit 7 set
inline const uint8_t* tight_vbyte_decode(const uint8_t* in, const uint8_t* end, uint32_t* out) {
    uint32_t v = 0;
    for (unsigned shift = 0;; shift += 7) {
        if (in >= end || shift > 28) throw format_error("bad TightVariableByte value");
        uint8_t c = *in++;
        v += uint32_t(c & 127) << shift;
        if (c & 128) break;
    }
    *out = v;
    return in;
}

struct block_index_file {
    global_params params;
    uint64_t size = 0;        // number of posting lists
    uint64_t num_docs = 0;
    bitvec_view endpoints;    // compact_elias_fano of the list start offsets
    const uint8_t* lists = nullptr;
    uint64_t lists_bytes = 0;
};

inline block_index_file parse_block_index(const uint8_t* p, size_t n) {
    byte_reader r(p, n);
    block_index_file f;
    (void)r.get<uint64_t>();                 // mapper flags
    f.params = read_params(r);
    f.size = r.get<uint64_t>();
    f.num_docs = r.get<uint64_t>();
    f.endpoints = read_bitvec(r);
    f.lists_bytes = r.get<uint64_t>();
    f.lists = r.take(size_t(f.lists_bytes));
    if (f.num_docs == 0 || f.num_docs > 0xFFFFFFFFull) throw format_error("num_docs out of range");
    if (f.params.ef_log_sampling0 > 32 || f.params.ef_log_sampling1 > 32) throw format_error("bad global parameters");
    return f;
}

struct wand_file {
    uint64_t num_docs = 0, num_terms = 0;
    const uint8_t* norm_lens = nullptr;        // num_docs fp32, unaligned in the file
    const uint8_t* max_term_weight = nullptr;  // num_terms fp32
};

inline wand_file parse_wand(const uint8_t* p, size_t n) {
    byte_reader r(p, n);
    wand_file f;
    (void)r.get<uint64_t>();
    f.num_docs = r.get<uint64_t>();
    if (f.num_docs > r.remaining() / 4) throw format_error("wand data: norm_lens longer than file");
    f.norm_lens = r.take(size_t(f.num_docs) * 4);
    f.num_terms = r.get<uint64_t>();
    if (f.num_terms > r.remaining() / 4) throw format_error("wand data: max_term_weight longer than file");
    f.max_term_weight = r.take(size_t(f.num_terms) * 4);
    return f;
}

}  // namespace ds2i_gpu
