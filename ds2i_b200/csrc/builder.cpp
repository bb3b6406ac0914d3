// ds2i_build — multi-threaded, format-compatible builder of ds2i index files and wand data, plus the
// synthetic Zipfian collection generator of the benchmark.
//
// The files written here are byte-identical to what the reference's create_freq_index /
// create_wand_data write for the same collection (tests/test_builder.py checks that against the
// reference-built golden files), so the GPU query path and the reference CPU path read the SAME
// index file.  Formats restated from:
//   succinct::mapper::freeze            succinct/mapper.hpp:51-98
//   block_freq_index::builder           block_freq_index.hpp:18-70
//   block_posting_list::write           block_posting_list.hpp:14-53
//   optpfor_block::encode / findBestB   block_codecs.hpp:150-207; FastPFor newpfor.h:149-211, optpfor.h:55-107
//   Simple16 encoder                    FastPFor/headers/simple16.h:181-420
//   interpolative_block::encode         block_codecs.hpp:105-125; interpolative_coding.hpp:11-73
//   compact_elias_fano::write           compact_elias_fano.hpp:68-136
//   wand_data ctor / map                wand_data.hpp:16-52,71-78
//
// ATTRIBUTION.  Byte-identical output leaves no freedom in the encoders: every decision the reference takes (which b OptPFD
// picks, where a partition ends, which of three codes a partition gets, where a sampled pointer goes) has to be taken the same
// way.  The following functions are therefore close restatements of the reference's encoders — same algorithm, same
// arithmetic, in places the same variable roles — and are offline host tooling, not part of the query path:
//   optpfor_find_best_b / collect_exceptions / s16_encode   block_codecs.hpp:156-182, FastPFor newpfor.h:149-211, simple16.h:181-420
//   write_posting_list                                       block_posting_list.hpp:14-53
//   bit_writer32 / interpolative_encode                      interpolative_coding.hpp:11-73, block_codecs.hpp:105-125
//   ef_write                                                 compact_elias_fano.hpp:69-135
//   rb_write / partition_bits / choose_partitions / partitioned_write   compact_ranked_bitvector.hpp:56-115, indexed_sequence.hpp:24-84,
//                                                            optimal_partition.hpp:69-121, partitioned_sequence.hpp:22-120
// Everything around them (the parallel chunked build, the list sources, the synthetic generator, sharding, the file writer) is ours.
//
//   ds2i_build gen   <prefix> <num_docs> <num_terms> <seed> [scale=0.35] [nqueries=10000] [qseed]
//   ds2i_build index <block_optpfor|block_interpolative|opt> <collection prefix> <out.idx> [threads]
//   ds2i_build wand  <collection prefix> <out.wand> [threads]
//   ds2i_build synth <out prefix> <num_docs> <num_terms> <seed> [threads] [nqueries] [types=block_optpfor, ':'-separated]
//        -> <out>.block_optpfor.idx, <out>.wand, <out>.queries without materialising the collection
//   ds2i_build shard <block_optpfor|block_interpolative> <collection prefix> <out prefix> <G> [threads]
//        document-partitioned shards: shard g holds the documents [g*N/G, (g+1)*N/G) with local
//        docids, as an ordinary ds2i index <out>.<g>.idx + wand data <out>.<g>.wand (norm_lens computed with the
//        collection-wide average length, so BM25 scores equal the unsharded ones) + <out>.<g>.terms (u32 global
//        term id of every list of the shard: ds2i lists cannot be empty, so absent terms are left out)
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

typedef std::vector<uint8_t> bytes;

// ------------------------------------------------------------------------------------------------
// codecs (encode side)
static void vbyte_encode(uint32_t val, bytes& out) {      // TightVariableByte::encode_single
    while (val >= 128) { out.push_back(uint8_t(val & 127)); val >>= 7; }
    out.push_back(uint8_t(val | 128));
}

static inline uint32_t msb32(uint32_t x) { return 31u - uint32_t(__builtin_clz(x)); }

struct bit_writer32 {        // interpolative_coding.hpp:11-73, LSB-first into u32 words
    std::vector<uint32_t>& buf;
    size_t size = 0;
    explicit bit_writer32(std::vector<uint32_t>& b) : buf(b) { buf.clear(); }
    void write(uint32_t bits, uint32_t len) {
        if (!len) return;
        uint32_t pos = size % 32;
        size += len;
        if (pos == 0) buf.push_back(bits);
        else {
            buf.back() |= bits << pos;
            if (len > 32 - pos) buf.push_back(bits >> (32 - pos));
        }
    }
    void write_int(uint32_t val, uint32_t u) {
        uint32_t b = msb32(u);
        uint64_t m = (uint64_t(1) << (b + 1)) - u;
        if (val < m) write(val, b);
        else { val += uint32_t(m); write(val >> 1, b); write(val & 1, 1); }
    }
    void write_interpolative(const uint32_t* in, size_t n, uint32_t low, uint32_t high) {
        if (!n) return;
        size_t h = n / 2;
        uint32_t val = in[h];
        write_int(val - low, high - low + 1);
        write_interpolative(in, h, low, val);
        write_interpolative(in + h + 1, n - h - 1, val, high);
    }
};

struct codec_scratch {
    std::vector<uint32_t> inbuf = std::vector<uint32_t>(128), outbuf;
    uint32_t exceptions[2 * 128 + 32];
    uint32_t tobecoded[128];
};

static void interpolative_encode(const uint32_t* in, uint32_t sum_of_values, size_t n, bytes& out, codec_scratch& s) {
    s.inbuf[0] = in[0];
    for (size_t i = 1; i < n; ++i) s.inbuf[i] = s.inbuf[i - 1] + in[i];
    if (sum_of_values == uint32_t(-1)) {
        sum_of_values = s.inbuf[n - 1];
        vbyte_encode(sum_of_values, out);
    }
    bit_writer32 bw(s.outbuf);
    bw.write_interpolative(s.inbuf.data(), n - 1, 0, sum_of_values);
    const uint8_t* p = reinterpret_cast<const uint8_t*>(s.outbuf.data());
    out.insert(out.end(), p, p + (bw.size + 7) / 8);
}

// Simple16: selector -> runs of (count, bits); first layout that fits wins (simple16.h:181-420)
static const uint8_t S16_RUNS[16][6] = {
    {28, 1, 0, 0, 0, 0}, {7, 2, 14, 1, 0, 0}, {7, 1, 7, 2, 7, 1}, {14, 1, 7, 2, 0, 0}, {14, 2, 0, 0, 0, 0}, {1, 4, 8, 3, 0, 0},
    {1, 3, 4, 4, 3, 3},  {7, 4, 0, 0, 0, 0},  {4, 5, 2, 4, 0, 0}, {2, 4, 4, 5, 0, 0},  {3, 6, 2, 5, 0, 0},  {2, 5, 3, 6, 0, 0},
    {4, 7, 0, 0, 0, 0},  {1, 10, 2, 9, 0, 0}, {2, 14, 0, 0, 0, 0}, {1, 28, 0, 0, 0, 0}};

// returns the number of values the chosen word takes; *word receives the encoded word
static inline uint32_t s16_pack_one(const uint32_t* in, size_t remaining, uint32_t* word) {
    for (uint32_t sel = 0; sel < 16; ++sel) {
        const uint8_t* r = S16_RUNS[sel];
        size_t left = remaining, idx = 0;
        bool ok = true;
        for (int k = 0; k < 6 && ok; k += 2) {
            size_t c = std::min<size_t>(r[k], left);
            for (size_t i = 0; i < c; ++i)
                if (in[idx + i] >= (1u << r[k + 1])) { ok = false; break; }
            idx += c; left -= c;
        }
        if (!ok) {
            if (sel == 15) throw std::runtime_error("Simple16: value out of range");
            continue;
        }
        if (word) {
            uint32_t w = sel, fill = 0;
            left = remaining; idx = 0;
            for (int k = 0; k < 6; k += 2) {
                size_t c = std::min<size_t>(r[k], left);
                for (size_t i = 0; i < c; ++i) w = (w << r[k + 1]) | in[idx + i];
                fill += uint32_t(c) * r[k + 1];
                idx += c; left -= c;
            }
            w <<= 28 - fill;
            *word = w;
            return uint32_t(idx);
        }
        return uint32_t(std::min<size_t>(remaining, size_t(r[0]) + r[2] + r[4]));
    }
    return 0;
}

static size_t s16_fake_encode(const uint32_t* in, size_t n) {
    size_t words = 0;
    while (n) { uint32_t c = s16_pack_one(in, n, nullptr); in += c; n -= c; ++words; }
    return words;
}

static size_t s16_encode(const uint32_t* in, size_t n, uint32_t* out) {
    size_t words = 0;
    while (n) { uint32_t c = s16_pack_one(in, n, out + words); in += c; n -= c; ++words; }
    return words;
}

static const uint32_t POSS_LOGS[17] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 16, 20, 32};

static uint32_t collect_exceptions(uint32_t b, const uint32_t* in, uint32_t* exceptions) {
    // layout: [position gaps - 1 (first absolute) ... | (value >> b) - 1 ...]  (newpfor.h:170-190)
    uint32_t pos[128], val[128], n = 0;
    for (uint32_t i = 0; i < 128; ++i)
        if (in[i] >= (1u << b)) { pos[n] = i; val[n] = in[i] >> b; ++n; }
    for (uint32_t i = 0; i < n; ++i) {
        exceptions[i] = i ? pos[i] - pos[i - 1] - 1 : pos[0];
        exceptions[i + n] = val[i] - 1;
    }
    return n;
}

static uint32_t optpfor_try_b(uint32_t b, const uint32_t* in, codec_scratch& s) {   // OPTPFor::tryB, optpfor.h:55-107
    if (b == 32) return 128;
    uint32_t size = (128 * b + 31) / 32;
    uint32_t n = collect_exceptions(b, in, s.exceptions);
    if (n) size += uint32_t(s16_fake_encode(s.exceptions, 2 * n));
    return size;
}

static uint32_t optpfor_find_best_b(const uint32_t* in, codec_scratch& s) {        // block_codecs.hpp:156-182
    uint32_t acc = 0;
    for (uint32_t i = 0; i < 128; ++i) acc |= in[i];
    const uint32_t mb = acc ? 32 - uint32_t(__builtin_clz(acc)) : 0;
    uint32_t b = 0, bsize = 0xffffffffu, i = 0;
    while (mb > 28 + POSS_LOGS[i]) ++i;
    for (; i < 17; ++i) {
        if (POSS_LOGS[i] > mb && POSS_LOGS[i] >= mb) break;
        uint32_t csize = optpfor_try_b(POSS_LOGS[i], in, s);
        if (csize <= bsize) { b = POSS_LOGS[i]; bsize = csize; }
    }
    return b;
}

static void optpfor_encode(const uint32_t* in, uint32_t sum_of_values, size_t n, bytes& out, codec_scratch& s) {
    if (n < 128) { interpolative_encode(in, sum_of_values, n, out, s); return; }
    uint32_t buf[1 + 2 * 128 + 32 + 128];
    uint32_t b = optpfor_find_best_b(in, s);
    size_t words;
    if (b < 32) {                                                   // NewPFor::encodeBlock, newpfor.h:149-203
        uint32_t nexc = collect_exceptions(b, in, s.exceptions);
        for (uint32_t i = 0; i < 128; ++i) s.tobecoded[i] = (b == 0) ? 0 : (in[i] & ((1u << b) - 1));
        size_t excw = nexc ? s16_encode(s.exceptions, 2 * nexc, buf + 1) : 0;
        buf[0] = (b << 26) | (nexc << 16) | uint32_t(excw);
        uint32_t* o = buf + 1 + excw;
        for (uint32_t g = 0; g < 4; ++g) {                          // fastpackwithoutmask: 32 values, b bits, LSB-first
            uint64_t acc = 0; uint32_t have = 0, w = 0;
            for (uint32_t i = 0; i < 32; ++i) {
                acc |= uint64_t(s.tobecoded[32 * g + i]) << have;
                have += b;
                if (have >= 32) { o[w++] = uint32_t(acc); acc >>= 32; have -= 32; }
            }
            o += b;
        }
        words = size_t(o - buf);
    } else {
        buf[0] = 32u << 26;
        memcpy(buf + 1, in, 128 * 4);
        words = 129;
    }
    const uint8_t* p = reinterpret_cast<const uint8_t*>(buf);
    out.insert(out.end(), p, p + 4 * words);
}

enum codec_id { C_OPTPFOR, C_INTERPOLATIVE };

static void block_encode(codec_id c, const uint32_t* in, uint32_t sum, size_t n, bytes& out, codec_scratch& s) {
    if (c == C_OPTPFOR) optpfor_encode(in, sum, n, out, s);
    else interpolative_encode(in, sum, n, out, s);
}

// block_posting_list::write (block_posting_list.hpp:14-53)
static void write_posting_list(codec_id c, bytes& out, uint32_t n, const uint32_t* docs, const uint32_t* freqs, codec_scratch& s) {
    vbyte_encode(n, out);
    const uint64_t block_size = 128, blocks = (uint64_t(n) + block_size - 1) / block_size;
    size_t begin_block_maxs = out.size();
    size_t begin_block_endpoints = begin_block_maxs + 4 * blocks;
    size_t begin_blocks = begin_block_endpoints + 4 * (blocks - 1);
    out.resize(begin_blocks);
    uint32_t docs_buf[128], freqs_buf[128];
    uint32_t last_doc = uint32_t(-1), block_base = 0;
    size_t k = 0;
    for (size_t b = 0; b < blocks; ++b) {
        uint32_t cur = ((b + 1) * block_size <= n) ? uint32_t(block_size) : uint32_t(n % block_size);
        for (uint32_t i = 0; i < cur; ++i, ++k) {
            uint32_t doc = docs[k];
            docs_buf[i] = doc - last_doc - 1;
            last_doc = doc;
            freqs_buf[i] = freqs[k] - 1;
        }
        memcpy(&out[begin_block_maxs + 4 * b], &last_doc, 4);
        block_encode(c, docs_buf, last_doc - block_base - (cur - 1), cur, out, s);
        block_encode(c, freqs_buf, uint32_t(-1), cur, out, s);
        if (b != blocks - 1) {
            uint32_t e = uint32_t(out.size() - begin_blocks);
            memcpy(&out[begin_block_endpoints + 4 * b], &e, 4);
        }
        block_base = last_doc + 1;
    }
}

// ------------------------------------------------------------------------------------------------
// succinct::bit_vector_builder subset + compact_elias_fano::write
struct bitvec_builder {
    std::vector<uint64_t> w;
    uint64_t size = 0;
    void zero_extend(uint64_t n) { size += n; w.resize((size + 63) / 64, 0); }
    void set(uint64_t pos) { w[pos >> 6] |= uint64_t(1) << (pos & 63); }
    void set_bits(uint64_t pos, uint64_t bits, uint32_t len) {
        if (!len) return;
        uint64_t mask = len == 64 ? ~uint64_t(0) : ((uint64_t(1) << len) - 1);
        uint64_t word = pos >> 6, in = pos & 63;
        w[word] &= ~(mask << in);
        w[word] |= bits << in;
        if (in + len > 64) {
            uint64_t stored = 64 - in;
            w[word + 1] &= ~(mask >> stored);
            w[word + 1] |= bits >> stored;
        }
    }
};

static inline uint32_t msb64(uint64_t x) { return 63u - uint32_t(__builtin_clzll(x)); }
static inline uint64_t ceil_log2(uint64_t x) { return x > 1 ? msb64(x - 1) + 1 : 0; }

static void ef_write(bitvec_builder& bvb, const uint64_t* vals, uint64_t universe, uint64_t n, uint32_t log_s0, uint32_t log_s1) {
    uint64_t base = bvb.size;
    uint64_t lower_bits = universe > n ? msb64(universe / n) : 0;
    uint64_t mask = (uint64_t(1) << lower_bits) - 1;
    uint64_t higher_bits_length = n + (universe >> lower_bits) + 2;
    uint64_t pointer_size = ceil_log2(higher_bits_length);
    uint64_t pointers0 = (higher_bits_length - n) >> log_s0, pointers1 = n >> log_s1;
    uint64_t p0_off = base, p1_off = p0_off + pointers0 * pointer_size, hi_off = p1_off + pointers1 * pointer_size;
    uint64_t lo_off = hi_off + higher_bits_length, end = lo_off + n * lower_bits;
    bvb.zero_extend(end - base);
    uint64_t sample1_mask = (uint64_t(1) << log_s1) - 1;
    auto set_ptr0s = [&](uint64_t begin, uint64_t end_, uint64_t rank_end) {
        uint64_t begin_zeros = begin - rank_end, end_zeros = end_ - rank_end;
        uint64_t step = uint64_t(1) << log_s0;
        for (uint64_t ptr0 = (begin_zeros + step - 1) / step; (ptr0 << log_s0) < end_zeros; ++ptr0) {
            if (!ptr0) continue;
            bvb.set_bits(p0_off + (ptr0 - 1) * pointer_size, (ptr0 << log_s0) + rank_end, uint32_t(pointer_size));
        }
    };
    uint64_t last_high = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t v = vals[i];
        uint64_t high = (v >> lower_bits) + i + 1, low = v & mask;
        bvb.set(hi_off + high);
        bvb.set_bits(lo_off + i * lower_bits, low, uint32_t(lower_bits));
        if (i && (i & sample1_mask) == 0) bvb.set_bits(p1_off + ((i >> log_s1) - 1) * pointer_size, high, uint32_t(pointer_size));
        set_ptr0s(last_high + 1, high, i);
        last_high = high;
    }
    set_ptr0s(last_high + 1, higher_bits_length, n);
}

// ------------------------------------------------------------------------------------------------
// collections
struct mapped {
    const uint8_t* p = nullptr; size_t n = 0;
    explicit mapped(std::string const& path) {
        int fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) throw std::runtime_error("cannot open " + path);
        struct stat st; fstat(fd, &st); n = size_t(st.st_size);
        p = static_cast<const uint8_t*>(mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0));
        close(fd);
        if (p == MAP_FAILED) throw std::runtime_error("mmap failed " + path);
    }
    const uint32_t* u32() const { return reinterpret_cast<const uint32_t*>(p); }
    size_t words() const { return n / 4; }
};

// abstract source of posting lists so that `index` (from files) and `synth` (generated) share the builder
struct list_source {
    uint64_t num_docs = 0, num_lists = 0;
    std::vector<uint64_t> est_len;      // for load balancing
    // fetch list i into docs/freqs (resized by the callee)
    std::function<void(uint64_t, std::vector<uint32_t>&, std::vector<uint32_t>&)> get;
};

static list_source collection_source(std::string const& prefix, std::shared_ptr<mapped>& dm, std::shared_ptr<mapped>& fm) {
    dm.reset(new mapped(prefix + ".docs"));
    fm.reset(new mapped(prefix + ".freqs"));
    const uint32_t* d = dm->u32();
    const uint32_t* f = fm->u32();
    list_source src;
    if (dm->words() < 2 || d[0] != 1) throw std::runtime_error("bad .docs header");
    src.num_docs = d[1];
    auto starts = std::make_shared<std::vector<std::pair<uint64_t, uint64_t>>>();   // (docs word, freqs word)
    uint64_t pos = 2, fpos = 0;
    while (pos < dm->words()) {
        uint32_t n = d[pos];
        if (n) { starts->push_back({pos, fpos}); src.est_len.push_back(n); }   // zero-length sequences are skipped (binary_collection.hpp:134)
        pos += 1 + n; fpos += 1 + n;
    }
    src.num_lists = starts->size();
    src.get = [d, f, starts](uint64_t i, std::vector<uint32_t>& docs, std::vector<uint32_t>& freqs) {
        uint64_t p = (*starts)[i].first, q = (*starts)[i].second;
        uint32_t n = d[p];
        docs.assign(d + p + 1, d + p + 1 + n);
        freqs.assign(f + q + 1, f + q + 1 + n);
    };
    return src;
}

// ------------------------------------------------------------------------------------------------
// synthetic collection: Zipfian document frequencies, clustered docids, geometric freqs.
// Every list has its own counter-seeded xoshiro256** stream, so generation is deterministic for
// any thread count and lists can be regenerated independently.
struct rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t& x) {
        uint64_t z = (x += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    rng(uint64_t seed, uint64_t stream) {
        uint64_t x = seed ^ (stream * 0xd1342543de82ef95ull + 0x2545f4914f6cdd1dull);
        for (int i = 0; i < 4; ++i) s[i] = splitmix(x);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uniform() { return (double(next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }   // (0,1)
    double normal() { double u1 = uniform(), u2 = uniform(); return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2); }
    uint64_t geometric(double log1mp) { return uint64_t(std::log(uniform()) / log1mp) + 1; }   // >= 1
};

struct synth_params { uint64_t num_docs, num_terms, seed; double scale = 0.35, alpha = 0.75, sigma = 1.5; };

static uint64_t synth_df(synth_params const& sp, uint64_t rank /* 1-based */) {
    double v = std::floor(sp.scale * double(sp.num_docs) / std::pow(double(rank), sp.alpha) + 0.5);
    uint64_t hi = sp.num_docs / 2;
    if (v < 1) return 1;
    return v > double(hi) ? hi : uint64_t(v);
}

static void synth_list(synth_params const& sp, uint64_t term, std::vector<uint32_t>& docs, std::vector<uint32_t>& freqs) {
    rng r(sp.seed, term);
    const uint64_t N = sp.num_docs;
    const double p = double(synth_df(sp, term + 1)) / double(N);
    // segment length: 2048 docids, widened for sparse lists so that a segment expects >= 0.5 postings
    uint64_t seg = 2048;
    while (p * double(seg) < 0.5 && seg < N) seg <<= 1;
    docs.clear();
    const double half_var = 0.5 * sp.sigma * sp.sigma;
    for (uint64_t s0 = 0; s0 < N; s0 += seg) {
        uint64_t s1 = std::min(N, s0 + seg);
        double ps = p * std::exp(sp.sigma * r.normal() - half_var);       // lognormal intensity, mean 1
        if (ps > 0.995) ps = 0.995;
        if (ps < 1e-12) continue;
        double l1p = std::log1p(-ps);
        uint64_t pos = s0 + r.geometric(l1p) - 1;
        while (pos < s1) { docs.push_back(uint32_t(pos)); pos += r.geometric(l1p); }
    }
    if (docs.empty()) docs.push_back(uint32_t(r.next() % N));
    freqs.resize(docs.size());
    const double lq = std::log(0.4);                                     // freq ~ Geometric(0.6) >= 1, capped 2^14
    for (auto& f : freqs) { uint64_t g = r.geometric(lq); f = uint32_t(std::min<uint64_t>(g, 16384)); }
}

static void synth_queries(synth_params const& sp, uint64_t nq, uint64_t qseed, std::string const& path) {
    rng r(qseed, 0x51ed270b);
    static const double cum[8] = {0.09, 0.43, 0.68, 0.83, 0.91, 0.94, 0.97, 1.0};   // lengths 1..8 (histogram of T's queries)
    std::ofstream out(path);
    const double lnT = std::log(double(sp.num_terms));
    for (uint64_t q = 0; q < nq; ++q) {
        double u = r.uniform();
        int len = 1;
        while (len < 8 && u > cum[len - 1]) ++len;
        std::vector<uint32_t> t;
        for (int i = 0; i < len; ++i) {
            uint64_t rank = uint64_t(std::exp(r.uniform() * lnT));       // log-uniform in [1, T]
            if (rank < 1) rank = 1;
            if (rank > sp.num_terms) rank = sp.num_terms;
            t.push_back(uint32_t(rank - 1));
        }
        std::sort(t.begin(), t.end());
        t.erase(std::unique(t.begin(), t.end()), t.end());
        for (size_t i = 0; i < t.size(); ++i) out << (i ? "\t" : "") << t[i];
        out << "\n";
    }
}

// ------------------------------------------------------------------------------------------------
template <typename F>
static void parallel_ranges(std::vector<uint64_t> const& weight, unsigned threads, F fn) {
    // contiguous ranges of roughly equal total weight, many more ranges than threads (dynamic pick-up)
    uint64_t total = 0;
    for (auto w : weight) total += w + 64;
    uint64_t chunks = std::max<uint64_t>(1, std::min<uint64_t>(weight.size(), uint64_t(threads) * 16));
    std::vector<uint64_t> bounds{0};
    uint64_t acc = 0, target = (total + chunks - 1) / chunks;
    for (uint64_t i = 0; i < weight.size(); ++i) {
        acc += weight[i] + 64;
        if (acc >= target) { bounds.push_back(i + 1); acc = 0; }
    }
    if (bounds.back() != weight.size()) bounds.push_back(weight.size());
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; ++t)
        pool.emplace_back([&]() {
            while (true) {
                size_t c = next.fetch_add(1);
                if (c + 1 >= bounds.size()) break;
                fn(c, bounds[c], bounds[c + 1]);
            }
        });
    for (auto& th : pool) th.join();
    (void)fn;
}

struct built_chunk { bytes data; std::vector<uint64_t> list_bytes; };

static void put_u64(FILE* f, uint64_t v) { fwrite(&v, 8, 1, f); }

// block_freq_index file: flags | 5 params | m_size | m_num_docs | m_endpoints bit_vector | m_lists (block_freq_index.hpp:124-134)
static void build_block_index(list_source const& src, codec_id codec, std::string const& out_path, unsigned threads,
                              std::vector<std::atomic<uint32_t>>* doclen /* nullable: accumulates sum of freqs per doc */) {
    std::vector<std::pair<size_t, built_chunk>> chunks;
    std::vector<built_chunk> slots(std::max<size_t>(1, std::min<size_t>(src.num_lists, size_t(threads) * 16)) + 2);
    parallel_ranges(src.est_len, threads, [&](size_t c, uint64_t lo, uint64_t hi) {
        codec_scratch s;
        std::vector<uint32_t> docs, freqs;
        built_chunk& bc = slots[c];
        for (uint64_t i = lo; i < hi; ++i) {
            src.get(i, docs, freqs);
            if (docs.empty()) throw std::invalid_argument("List must be nonempty");
            size_t before = bc.data.size();
            write_posting_list(codec, bc.data, uint32_t(docs.size()), docs.data(), freqs.data(), s);
            bc.list_bytes.push_back(bc.data.size() - before);
            if (doclen) for (size_t k = 0; k < docs.size(); ++k) (*doclen)[docs[k]].fetch_add(freqs[k], std::memory_order_relaxed);
        }
    });
    std::vector<uint64_t> endpoints;
    endpoints.reserve(src.num_lists + 1);
    uint64_t total = 0;
    for (auto const& bc : slots)
        for (auto b : bc.list_bytes) { endpoints.push_back(total); total += b; }
    if (endpoints.size() != src.num_lists) throw std::runtime_error("internal: list count mismatch");
    bitvec_builder bvb;
    ef_write(bvb, endpoints.data(), total, src.num_lists, 9, 8);     // global_parameters defaults (global_parameters.hpp:6-12)
    FILE* f = fopen(out_path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + out_path);
    put_u64(f, 0);
    const uint8_t params[5] = {9, 8, 9, 8, 7};
    fwrite(params, 1, 5, f);
    put_u64(f, src.num_lists);
    put_u64(f, src.num_docs);
    put_u64(f, bvb.size);
    put_u64(f, bvb.w.size());
    fwrite(bvb.w.data(), 8, bvb.w.size(), f);
    put_u64(f, total);
    for (auto const& bc : slots) if (!bc.data.empty()) fwrite(bc.data.data(), 1, bc.data.size(), f);
    fclose(f);
}

// ------------------------------------------------------------------------------------------------
// `opt` index: freq_index<partitioned_sequence<indexed_sequence>, positive_sequence<partitioned_sequence<strict_sequence>>>
// (index_types.hpp:24-27).  The file must be byte-identical to create_freq_index's, so the SAME decisions have to be
// taken: the (1 + eps) shortest-path partitioner over the same cost function with the same integer / double arithmetic
// (optimal_partition.hpp:69-121, configuration.hpp:29-31: eps1 = 0.03, eps2 = 0.3, fix cost 64 bits), the cheapest of
// Elias-Fano / ranked bitvector / all-ones per partition (indexed_sequence.hpp:24-84, strict_sequence.hpp:32-98), and
// the bit layouts of Appendix B.6 of SURVEY.md (compact_elias_fano.hpp:14-135, compact_ranked_bitvector.hpp:14-115,
// partitioned_sequence.hpp:22-120, positive_sequence.hpp:14-30, freq_index.hpp:64-96, bitvector_collection.hpp:22-41).
// Restated from those specifications; tests/test_builder.py compares the result with the reference-built fixtures.
struct bit_sink : bitvec_builder {
    void push(uint64_t bits, uint32_t len) {                         // append_bits
        if (!len) return;
        uint64_t at = size;
        zero_extend(len);
        set_bits(at, bits, len);
    }
    void push_all(bitvec_builder const& o) {                          // append(bit_vector_builder)
        if (!o.size) return;
        const uint64_t at = size, shift = at & 63;
        zero_extend(o.size);
        uint64_t wi = at >> 6;
        const size_t nw = size_t((o.size + 63) / 64);
        if (!shift) { std::copy(o.w.begin(), o.w.begin() + nw, w.begin() + wi); return; }
        for (size_t i = 0; i < nw; ++i) {
            w[wi + i] |= o.w[i] << shift;
            if (wi + i + 1 < w.size()) w[wi + i + 1] |= o.w[i] >> (64 - shift);
        }
    }
    void gamma(uint64_t v) {                                           // write_gamma (integer_codes.hpp:6-13)
        uint64_t nn = v + 1, l = msb64(nn), hb = uint64_t(1) << l;
        push(hb, uint32_t(l + 1));
        push(nn ^ hb, uint32_t(l));
    }
    void delta(uint64_t v) {                                           // write_delta (:32-39)
        uint64_t nn = v + 1, l = msb64(nn), hb = uint64_t(1) << l;
        gamma(l);
        push(nn ^ hb, uint32_t(l));
    }
};

struct ef_params { uint32_t log_s0 = 9, log_s1 = 8, log_rank1 = 9, log_rb_s1 = 8; };    // global_parameters.hpp:6-12
static ef_params strict_params() { ef_params p; p.log_s0 = 63; p.log_rank1 = 63; return p; }   // strict_sequence.hpp:24-30

static inline uint64_t ef_bits(uint64_t universe, uint64_t n, ef_params const& p) {      // compact_elias_fano::offsets::end
    const uint64_t lower = universe > n ? msb64(universe / n) : 0;
    const uint64_t hlen = n + (universe >> lower) + 2;
    const uint64_t psize = ceil_log2(hlen);
    return ((hlen - n) >> p.log_s0) * psize + (n >> p.log_s1) * psize + hlen + n * lower;
}
static inline uint64_t rb_bits(uint64_t universe, uint64_t n, ef_params const& p) {      // compact_ranked_bitvector::offsets::end
    return (universe >> p.log_rank1) * ceil_log2(n + 1) + (n >> p.log_rb_s1) * ceil_log2(universe) + universe;
}

// compact_ranked_bitvector::write: rank samples every 2^log_rank1 positions, a select pointer every 2^log_rb_s1 ones, the bitmap
static void rb_write(bit_sink& out, const uint64_t* v, uint64_t universe, uint64_t n, ef_params const& p) {
    const uint64_t base = out.size, rsize = ceil_log2(n + 1), psize = ceil_log2(universe);
    const uint64_t nsamples = universe >> p.log_rank1, nptrs = n >> p.log_rb_s1;
    const uint64_t ptr_off = base + nsamples * rsize, bits_off = ptr_off + nptrs * psize;
    out.zero_extend(bits_off + universe - base);
    // sample k (k >= 1) = number of ones before position k << log_rank1
    auto samples_between = [&](uint64_t from, uint64_t to, uint64_t rank) {
        const uint64_t step = uint64_t(1) << p.log_rank1;
        for (uint64_t k = (from + step - 1) / step; (k << p.log_rank1) < to; ++k)
            if (k) out.set_bits(base + (k - 1) * rsize, rank, uint32_t(rsize));
    };
    const uint64_t smask = (uint64_t(1) << p.log_rb_s1) - 1;
    uint64_t prev = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t x = v[i];
        if (i && x <= prev) throw std::runtime_error("sequence is not strictly increasing");
        out.set(bits_off + x);
        if (i && !(i & smask)) out.set_bits(ptr_off + ((i >> p.log_rb_s1) - 1) * psize, x, uint32_t(psize));
        samples_between(prev + 1, x + 1, i);
        prev = x;
    }
    samples_between(prev + 1, universe, n);
}

// cost in bits of the cheapest representation of n values in [0, universe) (indexed_sequence / strict_sequence ::bitsize)
static inline uint64_t partition_bits(uint64_t universe, uint64_t n, bool strict, ef_params const& p, int* type = nullptr) {
    uint64_t best = universe == n ? 0 : ~uint64_t(0);
    int t = 2;                                                       // all_ones
    if (best) {
        const uint64_t ef = (strict ? ef_bits(universe - n + 1, n, p) : ef_bits(universe, n, p)) + 1;
        if (ef < best) { best = ef; t = 0; }
        const uint64_t rb = rb_bits(universe, n, p) + 1;
        if (rb < best) { best = rb; t = 1; }
    }
    if (type) *type = t;
    return best;
}

static void partition_write(bit_sink& out, const uint64_t* v, uint64_t universe, uint64_t n, bool strict, std::vector<uint64_t>& tmp) {
    const ef_params p = strict ? strict_params() : ef_params();
    int type;
    partition_bits(universe, n, strict, p, &type);
    if (type != 2) out.push(uint64_t(type), 1);                     // the type bit is absent for all-ones partitions
    if (type == 0) {
        if (!strict) { ef_write(out, v, universe, n, p.log_s0, p.log_s1); return; }
        tmp.resize(n);                                               // strict_elias_fano: v_i - i in a universe smaller by n - 1
        for (uint64_t i = 0; i < n; ++i) tmp[i] = v[i] - i;
        ef_write(out, tmp.data(), universe - n + 1, n, p.log_s0, p.log_s1);
    } else if (type == 1) {
        rb_write(out, v, universe, n, p);
    }
}

// End positions of the partitions chosen by the reference's approximate shortest-path search.  One window per cost
// bound lb, lb (1 + eps2), lb (1 + eps2)^2, ... < lb / eps1; every window slides over the sequence keeping its cost
// just above its bound; an edge (i -> window end) relaxes the path cost.  Types follow the reference exactly (32-bit
// element arithmetic for the window universe, the bound truncated to an integer after every multiplication).
static std::vector<uint32_t> choose_partitions(const uint64_t* v, uint64_t universe, uint64_t n, bool strict) {
    const ef_params p = strict ? strict_params() : ef_params();
    const double eps1 = 0.03, eps2 = 0.3;
    const uint64_t fix_cost = 64;
    auto cost = [&](uint64_t u, uint64_t m) { return partition_bits(u, m, strict, p) + fix_cost; };
    struct window { uint32_t start = 0, end = 0, lo = 0, hi = 0; uint64_t bound = 0; };
    const uint64_t whole = cost(universe, n);
    std::vector<uint64_t> best(n + 1, whole);
    best[0] = 0;
    std::vector<window> wins;
    const uint64_t lb = cost(1, 1);
    for (uint64_t bound = lb; double(bound) < double(lb) / eps1;) {
        window w; w.lo = uint32_t(v[0]); w.bound = bound;
        wins.push_back(w);
        if (bound >= whole) break;
        bound = uint64_t(double(bound) * (1 + eps2));
    }
    std::vector<uint32_t> from(n + 1, 0);
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t reach = uint64_t(i) + 1;
        for (window& w : wins) {
            while (w.end < reach) { w.hi = uint32_t(v[w.end]); ++w.end; }
            while (true) {
                const uint64_t c = cost(uint64_t(uint32_t(w.hi - w.lo + 1)), uint64_t(w.end - w.start));
                if (best[i] + c < best[w.end]) { best[w.end] = best[i] + c; from[w.end] = i; }
                reach = w.end;
                if (w.end == n || c >= w.bound) break;
                w.hi = uint32_t(v[w.end]); ++w.end;
            }
            w.lo = uint32_t(v[w.start]) + 1; ++w.start;
        }
    }
    std::vector<uint32_t> ends;
    for (uint32_t at = uint32_t(n); at; at = from[at]) ends.push_back(at);
    std::reverse(ends.begin(), ends.end());
    return ends;
}

// partitioned_sequence::write (partitioned_sequence.hpp:22-120)
static void partitioned_write(bit_sink& out, const uint64_t* v, uint64_t universe, uint64_t n, bool strict) {
    const std::vector<uint32_t> ends = choose_partitions(v, universe, n, strict);
    const uint64_t parts = ends.size();
    out.gamma(parts - 1);                                            // write_gamma_nonzero
    std::vector<uint64_t> rel, tmp;
    if (parts == 1) {
        const uint64_t base = v[0];
        rel.resize(n);
        for (uint64_t i = 0; i < n; ++i) rel[i] = v[i] - base;
        out.push(base, uint32_t(ceil_log2(universe)));
        if (n > 1) out.delta(base + rel.back() + 1 == universe ? 0 : rel.back());      // 0 = "tight": the universe ends with the last value
        partition_write(out, rel.data(), rel.back() + 1, n, strict, tmp);
        return;
    }
    bit_sink bodies;
    std::vector<uint64_t> body_ends, uppers{v[0]}, sizes(ends.begin(), ends.end());
    uint64_t base = v[0], at = 0;
    for (uint64_t pi = 0; pi < parts; ++pi) {
        rel.clear();
        for (; at < ends[pi]; ++at) rel.push_back(v[at] - base);
        partition_write(bodies, rel.data(), rel.back() + 1, rel.size(), strict, tmp);
        body_ends.push_back(bodies.size);
        uppers.push_back(v[at - 1]);
        base = v[at - 1] + 1;
    }
    const ef_params g;                                               // the two directories use the index-wide parameters
    bit_sink bsizes, buppers;
    ef_write(bsizes, sizes.data(), n, parts - 1, g.log_s0, g.log_s1);
    ef_write(buppers, uppers.data(), universe, parts + 1, g.log_s0, g.log_s1);
    const uint64_t ebits = ceil_log2(bodies.size + 1);
    out.gamma(ebits);
    out.push_all(bsizes);
    out.push_all(buppers);
    for (uint64_t pi = 0; pi + 1 < parts; ++pi) out.push(body_ends[pi], uint32_t(ebits));
    out.push_all(bodies);
}

struct built_bits { bit_sink docs, freqs; std::vector<uint64_t> docs_len, freqs_len; };

static void write_bit_collection(FILE* f, std::vector<built_bits> const& slots, bool freqs, uint64_t num_lists) {
    bit_sink all;
    std::vector<uint64_t> starts;
    starts.reserve(num_lists);
    uint64_t total = 0;
    for (auto const& s : slots) for (uint64_t l : (freqs ? s.freqs_len : s.docs_len)) { starts.push_back(total); total += l; }
    if (starts.size() != num_lists) throw std::runtime_error("internal: list count mismatch");
    all.w.reserve(size_t(total / 64 + 2));
    for (auto const& s : slots) all.push_all(freqs ? s.freqs : s.docs);
    bit_sink ends;
    ef_write(ends, starts.data(), total, num_lists, 9, 8);
    put_u64(f, num_lists);
    put_u64(f, ends.size); put_u64(f, ends.w.size()); fwrite(ends.w.data(), 8, ends.w.size(), f);
    put_u64(f, all.size); put_u64(f, all.w.size()); fwrite(all.w.data(), 8, all.w.size(), f);
}

// freq_index file: flags | 5 params | m_num_docs | docs collection | freqs collection (freq_index.hpp:234-243)
static void build_opt_index(list_source const& src, std::string const& out_path, unsigned threads,
                            std::vector<std::atomic<uint32_t>>* doclen) {
    std::vector<built_bits> slots(std::max<size_t>(1, std::min<size_t>(src.num_lists, size_t(threads) * 16)) + 2);
    parallel_ranges(src.est_len, threads, [&](size_t c, uint64_t lo, uint64_t hi) {
        std::vector<uint32_t> docs, freqs;
        std::vector<uint64_t> dv, fv;
        built_bits& bb = slots[c];
        for (uint64_t i = lo; i < hi; ++i) {
            src.get(i, docs, freqs);
            const uint64_t n = docs.size();
            if (!n) throw std::invalid_argument("List must be nonempty");
            dv.assign(docs.begin(), docs.end());
            fv.resize(n);
            uint64_t occ = 0;
            for (uint64_t k = 0; k < n; ++k) { occ += freqs[k]; fv[k] = occ; }       // positive_sequence: prefix sums, strictly increasing
            const uint64_t d0 = bb.docs.size, f0 = bb.freqs.size;
            bb.docs.gamma(occ - 1);                                                   // list header (freq_index.hpp:66-71)
            if (occ > 1) bb.docs.push(n, uint32_t(ceil_log2(occ + 1)));
            partitioned_write(bb.docs, dv.data(), src.num_docs, n, false);
            partitioned_write(bb.freqs, fv.data(), occ + 1, n, true);
            bb.docs_len.push_back(bb.docs.size - d0);
            bb.freqs_len.push_back(bb.freqs.size - f0);
            if (doclen) for (size_t k = 0; k < n; ++k) (*doclen)[docs[k]].fetch_add(freqs[k], std::memory_order_relaxed);
        }
    });
    FILE* f = fopen(out_path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + out_path);
    put_u64(f, 0);
    const uint8_t params[5] = {9, 8, 9, 8, 7};
    fwrite(params, 1, 5, f);
    put_u64(f, src.num_docs);
    write_bit_collection(f, slots, false, src.num_lists);
    write_bit_collection(f, slots, true, src.num_lists);
    fclose(f);
}

// bm25::doc_term_weight (bm25.hpp:11-15) — same expression, same compiler flags as the reference tool
static inline float doc_term_weight(uint64_t freq, float norm_len) {
    const float b = 0.5f, k1 = 1.2f;
    float f = float(freq);
    return f / (f + k1 * (1.0f - b + b * norm_len));
}

// wand_data (wand_data.hpp:16-52): norm_lens then per-list max doc_term_weight
static std::vector<float> compute_norm_lens(const uint32_t* sizes, uint64_t N) {
    std::vector<float> norm_lens(N);
    double lens_sum = 0;
    for (uint64_t i = 0; i < N; ++i) { float len = float(sizes[i]); norm_lens[i] = len; lens_sum += len; }
    float avg_len = float(lens_sum / double(N));
    for (uint64_t i = 0; i < N; ++i) norm_lens[i] /= avg_len;
    return norm_lens;
}

static void build_wand(list_source const& src, const uint32_t* sizes, std::string const& out_path, unsigned threads,
                       const float* given_norm_lens = nullptr) {
    const uint64_t N = src.num_docs;
    std::vector<float> norm_lens = given_norm_lens ? std::vector<float>(given_norm_lens, given_norm_lens + N) : compute_norm_lens(sizes, N);
    std::vector<float> max_term_weight(src.num_lists);
    parallel_ranges(src.est_len, threads, [&](size_t, uint64_t lo, uint64_t hi) {
        std::vector<uint32_t> docs, freqs;
        for (uint64_t i = lo; i < hi; ++i) {
            src.get(i, docs, freqs);
            float max_score = 0;
            for (size_t k = 0; k < docs.size(); ++k) max_score = std::max(max_score, doc_term_weight(freqs[k], norm_lens[docs[k]]));
            max_term_weight[i] = max_score;
        }
    });
    FILE* f = fopen(out_path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + out_path);
    put_u64(f, 0);
    put_u64(f, N);
    fwrite(norm_lens.data(), 4, N, f);
    put_u64(f, src.num_lists);
    fwrite(max_term_weight.data(), 4, src.num_lists, f);
    fclose(f);
}

static codec_id parse_codec(std::string const& t) {
    if (t == "block_optpfor") return C_OPTPFOR;
    if (t == "block_interpolative") return C_INTERPOLATIVE;
    throw std::invalid_argument("builder supports block_optpfor and block_interpolative, not " + t);
}

static list_source synth_source(synth_params sp) {
    list_source src;
    src.num_docs = sp.num_docs; src.num_lists = sp.num_terms;
    src.est_len.resize(sp.num_terms);
    for (uint64_t t = 0; t < sp.num_terms; ++t) src.est_len[t] = synth_df(sp, t + 1);
    src.get = [sp](uint64_t i, std::vector<uint32_t>& d, std::vector<uint32_t>& f) { synth_list(sp, i, d, f); };
    return src;
}

int main(int argc, char** argv) {
    try {
        std::string cmd = argc > 1 ? argv[1] : "";
        unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        if (cmd == "index" && argc >= 5) {
            std::shared_ptr<mapped> dm, fm;
            list_source src = collection_source(argv[3], dm, fm);
            unsigned threads = argc > 5 ? unsigned(atoi(argv[5])) : hw;
            if (std::string(argv[2]) == "opt") build_opt_index(src, argv[4], threads, nullptr);
            else build_block_index(src, parse_codec(argv[2]), argv[4], threads, nullptr);
            return 0;
        }
        if (cmd == "wand" && argc >= 4) {
            std::shared_ptr<mapped> dm, fm;
            list_source src = collection_source(argv[2], dm, fm);
            mapped sm(std::string(argv[2]) + ".sizes");
            if (sm.words() < 1 + src.num_docs) throw std::runtime_error("bad .sizes");
            build_wand(src, sm.u32() + 1, argv[3], argc > 4 ? unsigned(atoi(argv[4])) : hw);
            return 0;
        }
        if ((cmd == "gen" || cmd == "synth") && argc >= 6) {
            synth_params sp;
            sp.num_docs = strtoull(argv[3], nullptr, 10); sp.num_terms = strtoull(argv[4], nullptr, 10); sp.seed = strtoull(argv[5], nullptr, 10);
            std::string prefix = argv[2];
            list_source src = synth_source(sp);
            if (cmd == "gen") {
                if (argc > 6) sp.scale = atof(argv[6]);
                src = synth_source(sp);
                uint64_t nq = argc > 7 ? strtoull(argv[7], nullptr, 10) : 10000;
                uint64_t qseed = argc > 8 ? strtoull(argv[8], nullptr, 10) : sp.seed + 1;
                std::vector<uint32_t> doclen(sp.num_docs, 0), d, f;
                FILE* fd = fopen((prefix + ".docs").c_str(), "wb");
                FILE* ff = fopen((prefix + ".freqs").c_str(), "wb");
                if (!fd || !ff) throw std::runtime_error("cannot write collection");
                uint32_t hdr[2] = {1, uint32_t(sp.num_docs)};
                fwrite(hdr, 4, 2, fd);
                for (uint64_t t = 0; t < sp.num_terms; ++t) {
                    synth_list(sp, t, d, f);
                    uint32_t n = uint32_t(d.size());
                    fwrite(&n, 4, 1, fd); fwrite(d.data(), 4, n, fd);
                    fwrite(&n, 4, 1, ff); fwrite(f.data(), 4, n, ff);
                    for (size_t k = 0; k < d.size(); ++k) doclen[d[k]] += f[k];
                }
                fclose(fd); fclose(ff);
                FILE* fs = fopen((prefix + ".sizes").c_str(), "wb");
                uint32_t n = uint32_t(sp.num_docs);
                fwrite(&n, 4, 1, fs);
                for (auto& l : doclen) if (!l) l = 1;
                fwrite(doclen.data(), 4, doclen.size(), fs);
                fclose(fs);
                synth_queries(sp, nq, qseed, prefix + ".queries");
                return 0;
            }
            unsigned threads = argc > 6 ? unsigned(atoi(argv[6])) : hw;
            if (!threads) threads = hw;
            uint64_t nq = argc > 7 ? strtoull(argv[7], nullptr, 10) : 10000;
            std::string types = argc > 8 ? argv[8] : "block_optpfor";
            std::vector<std::atomic<uint32_t>> doclen(sp.num_docs);
            for (auto& a : doclen) a.store(0, std::memory_order_relaxed);
            bool first = true;
            size_t start = 0;
            while (start <= types.size()) {
                size_t end = types.find(':', start);
                if (end == std::string::npos) end = types.size();
                std::string t = types.substr(start, end - start);
                if (t == "opt") build_opt_index(src, prefix + ".opt.idx", threads, first ? &doclen : nullptr);
                else build_block_index(src, parse_codec(t), prefix + "." + t + ".idx", threads, first ? &doclen : nullptr);
                first = false;
                start = end + 1;
            }
            std::vector<uint32_t> sizes(sp.num_docs);
            for (uint64_t i = 0; i < sp.num_docs; ++i) { uint32_t l = doclen[i].load(std::memory_order_relaxed); sizes[i] = l ? l : 1; }
            build_wand(src, sizes.data(), prefix + ".wand", threads);
            synth_queries(sp, nq, sp.seed + 1, prefix + ".queries");
            return 0;
        }
        if (cmd == "shard" && argc >= 6) {
            std::shared_ptr<mapped> dm, fm;
            list_source all = collection_source(argv[3], dm, fm);
            mapped sm(std::string(argv[3]) + ".sizes");
            if (sm.words() < 1 + all.num_docs) throw std::runtime_error("bad .sizes");
            const uint64_t G = strtoull(argv[5], nullptr, 10);
            if (G < 1 || G > all.num_docs) throw std::invalid_argument("bad shard count");
            unsigned threads = argc > 6 ? unsigned(atoi(argv[6])) : hw;
            const std::vector<float> norm_lens = compute_norm_lens(sm.u32() + 1, all.num_docs);     // collection-wide average
            std::string out = argv[4];
            for (uint64_t g = 0; g < G; ++g) {
                const uint32_t lo = uint32_t(all.num_docs * g / G), hi = uint32_t(all.num_docs * (g + 1) / G);
                // the terms that occur in the shard, and how often
                std::vector<uint32_t> terms;
                list_source sh;
                std::vector<uint32_t> d, f;
                for (uint64_t t = 0; t < all.num_lists; ++t) {
                    all.get(t, d, f);
                    size_t a = std::lower_bound(d.begin(), d.end(), lo) - d.begin(), b = std::lower_bound(d.begin(), d.end(), hi) - d.begin();
                    if (b > a) { terms.push_back(uint32_t(t)); sh.est_len.push_back(b - a); }
                }
                if (terms.empty()) throw std::runtime_error("shard without postings");
                sh.num_docs = hi - lo; sh.num_lists = terms.size();
                auto tp = std::make_shared<std::vector<uint32_t>>(terms);
                sh.get = [all, tp, lo, hi](uint64_t i, std::vector<uint32_t>& docs, std::vector<uint32_t>& freqs) {
                    std::vector<uint32_t> d2, f2;
                    all.get((*tp)[i], d2, f2);
                    size_t a = std::lower_bound(d2.begin(), d2.end(), lo) - d2.begin(), b = std::lower_bound(d2.begin(), d2.end(), hi) - d2.begin();
                    docs.resize(b - a); freqs.assign(f2.begin() + a, f2.begin() + b);
                    for (size_t k = a; k < b; ++k) docs[k - a] = d2[k] - lo;
                };
                std::string base = out + "." + std::to_string(g);
                build_block_index(sh, parse_codec(argv[2]), base + ".idx", threads, nullptr);
                build_wand(sh, nullptr, base + ".wand", threads, norm_lens.data() + lo);
                FILE* ft = fopen((base + ".terms").c_str(), "wb");
                if (!ft) throw std::runtime_error("cannot write " + base + ".terms");
                fwrite(terms.data(), 4, terms.size(), ft);
                fclose(ft);
            }
            return 0;
        }
        if (cmd == "types") { printf("block_optpfor block_interpolative opt\n"); return 0; }      // index types this builder writes
        fprintf(stderr, "usage: ds2i_build gen|index|wand|synth|shard|types ... (see the header of builder.cpp)\n");
        return 1;
    } catch (std::exception const& e) {
        fprintf(stderr, "ds2i_build: %s\n", e.what());
        return 2;
    }
}
