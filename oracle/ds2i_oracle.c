/*
 * ORACLE (CPU restatement) — TEST INFRASTRUCTURE ONLY.  Nothing in the product path may include,
 * link or execute this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may.
 *
 * Plain-C restatement of the ds2i algorithms on the hot path for the block indexes, written from the
 * reference sources cited at each function (paths relative to the ds2i tree).  It exists beside the
 * stronger oracle — the UNMODIFIED reference compiled into oracle/_ref by oracle/Makefile — so that
 * the GPU box has a checker even if those binaries could not run there, and as an independent
 * reading of the formats.  PINNED: tests/test_oracle.py checks it bit-for-bit against the golden
 * vectors the reference itself produced (tests/golden/mini.expected.strict.bin, mini.collection.npz)
 * for every operator and every posting, for all nine index types of DS2I_INDEX_TYPES.
 * The Elias-Fano family (opt, uniform, single, ef) is restated as a full decode of every list (formats, partition
 * encodings, strict / positive transformations); the operators then run over the decoded arrays.
 *
 *   ds2i_oracle dump  <type> <index> <wand> <queries> <out.bin> <op[:op..]> [k]
 *   ds2i_oracle lists <type> <index> <out.bin>            every posting of every list
 *   ds2i_oracle bench <type> <index> <wand> <queries> <op>  single-thread timing (JSON on stdout)
 * Output formats equal oracle/drivers/ref_tool.cpp's.  Build: gcc -O2 -ffp-contract=off (scores are
 * then bit-identical to the reference built with -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

enum { OPTPFOR, VARINT, INTERPOLATIVE, MIXED, QMX };

/* ---- succinct::mapper layout (succinct/mapper.hpp:51-98), block_freq_index::map (block_freq_index.hpp:124-134) */
typedef struct {
    int codec;
    uint64_t size, num_docs;
    uint64_t ep_bits, ep_nwords;
    const uint8_t* ep_words;     /* unaligned u64 words of m_endpoints */
    const uint8_t* lists;
    uint64_t lists_bytes;
    uint64_t* list_start;        /* decoded EF endpoints */
} block_index;

typedef struct { uint64_t num_docs, num_terms; const uint8_t* norm_lens; const uint8_t* max_term_weight; } wand_data;

static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static float rdf(const uint8_t* p) { float v; memcpy(&v, p, 4); return v; }

static uint64_t bv_bits(const uint8_t* words, uint64_t pos, unsigned len) {   /* succinct/bit_vector.hpp:251-268 */
    if (!len) return 0;
    uint64_t w = rd64(words + 8 * (pos >> 6)), sh = pos & 63;
    uint64_t v = w >> sh;
    if (sh + len > 64) v |= rd64(words + 8 * ((pos >> 6) + 1)) << (64 - sh);
    return len == 64 ? v : v & (((uint64_t)1 << len) - 1);
}
static unsigned msb64(uint64_t x) { return 63 - (unsigned)__builtin_clzll(x); }
static uint64_t ceil_log2(uint64_t x) { return x > 1 ? msb64(x - 1) + 1 : 0; }   /* util.hpp:30-33 */

/* compact_elias_fano (compact_elias_fano.hpp:14-61,105-118): all n values of the sequence written at bit `off` */
static void ef_decode_at(const uint8_t* words, uint64_t off, uint64_t universe, uint64_t n, unsigned ls0, unsigned ls1, uint64_t* out) {
    uint64_t l = universe > n ? msb64(universe / n) : 0;
    uint64_t hbl = n + (universe >> l) + 2, psize = ceil_log2(hbl);
    uint64_t p0 = ls0 >= 63 ? 0 : (hbl - n) >> ls0, p1 = n >> ls1;
    uint64_t high_off = off + p0 * psize + p1 * psize, low_off = high_off + hbl;
    uint64_t pos = high_off;
    for (uint64_t i = 0; i < n; ++i) {
        while (!bv_bits(words, pos, 1)) ++pos;
        uint64_t high = pos - high_off - i - 1;
        out[i] = (high << l) | bv_bits(words, low_off + i * l, (unsigned)l);
        ++pos;
    }
}
static void ef_decode_all(const uint8_t* words, uint64_t universe, uint64_t n, unsigned ls0, unsigned ls1, uint64_t* out) {
    ef_decode_at(words, 0, universe, n, ls0, ls1, out);
}
/* bits the sequence occupies (compact_elias_fano.hpp:14-61 `end`) */
static uint64_t ef_bitsize(uint64_t universe, uint64_t n, unsigned ls0, unsigned ls1) {
    uint64_t l = universe > n ? msb64(universe / n) : 0;
    uint64_t hbl = n + (universe >> l) + 2, psize = ceil_log2(hbl);
    uint64_t p0 = ls0 >= 63 ? 0 : (hbl - n) >> ls0, p1 = n >> ls1;
    return p0 * psize + p1 * psize + hbl + n * l;
}

static int load_file(const char* path, uint8_t** data, size_t* n) {
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END); *n = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
    *data = (uint8_t*)malloc(*n + 64);
    if (fread(*data, 1, *n, f) != *n) { fclose(f); return -1; }
    memset(*data + *n, 0, 64);
    fclose(f);
    return 0;
}

static int open_index(block_index* ix, const char* type, const uint8_t* p) {
    if (!strcmp(type, "block_optpfor")) ix->codec = OPTPFOR;
    else if (!strcmp(type, "block_varint")) ix->codec = VARINT;
    else if (!strcmp(type, "block_interpolative")) ix->codec = INTERPOLATIVE;
    else if (!strcmp(type, "block_mixed")) ix->codec = MIXED;
    else if (!strcmp(type, "block_qmx")) ix->codec = QMX;
    else return -1;
    const uint8_t* c = p + 8;                /* flags */
    unsigned ls0 = c[0], ls1 = c[1]; c += 5; /* global_parameters: 5 single bytes */
    ix->size = rd64(c); c += 8;
    ix->num_docs = rd64(c); c += 8;
    ix->ep_bits = rd64(c); c += 8;
    ix->ep_nwords = rd64(c); c += 8;
    ix->ep_words = c; c += 8 * ix->ep_nwords;
    ix->lists_bytes = rd64(c); c += 8;
    ix->lists = c;
    ix->list_start = (uint64_t*)malloc(8 * (ix->size + 1));
    ef_decode_all(ix->ep_words, ix->lists_bytes, ix->size, ls0, ls1, ix->list_start);   /* block_freq_index.hpp:59-63 */
    ix->list_start[ix->size] = ix->lists_bytes;
    return 0;
}

static void open_wand(wand_data* w, const uint8_t* p) {   /* wand_data.hpp:71-78 */
    w->num_docs = rd64(p + 8); w->norm_lens = p + 16;
    w->num_terms = rd64(p + 16 + 4 * w->num_docs); w->max_term_weight = p + 24 + 4 * w->num_docs;
}

/* ---- codecs ---- */
static const uint8_t* vbyte_decode(const uint8_t* in, uint32_t* out) {   /* TightVariableByte, block_codecs.hpp:84-98 */
    uint32_t v = 0;
    for (unsigned shift = 0;; shift += 7) {
        uint8_t c = *in++;
        v += (uint32_t)(c & 127) << shift;
        if (c & 128) break;
    }
    *out = v;
    return in;
}

typedef struct { const uint8_t* in; uint64_t buf; unsigned avail; size_t pos; } bit_reader;   /* interpolative_coding.hpp:79-153 */
static uint32_t br_read(bit_reader* br, unsigned len) {
    if (!len) return 0;
    if (br->avail < len) { br->buf |= (uint64_t)rd32(br->in) << br->avail; br->in += 4; br->avail += 32; }
    uint32_t v = (uint32_t)(br->buf & (((uint64_t)1 << len) - 1));
    br->buf >>= len; br->avail -= len; br->pos += len;
    return v;
}
static uint32_t br_read_int(bit_reader* br, uint32_t u) {
    unsigned b = 31 - (unsigned)__builtin_clz(u);
    uint64_t m = ((uint64_t)1 << (b + 1)) - u;
    uint32_t val = br_read(br, b);
    if (val >= m) val = (uint32_t)(((uint64_t)(val << 1) + br_read(br, 1)) - m);
    return val;
}
static void br_interpolative(bit_reader* br, uint32_t* out, size_t n, uint32_t low, uint32_t high) {
    size_t h = n / 2;
    uint32_t val = low + br_read_int(br, high - low + 1);
    out[h] = val;
    if (n == 1) return;
    if (h) br_interpolative(br, out, h, low, val);
    if (n - h - 1) br_interpolative(br, out + h + 1, n - h - 1, val, high);
}
static const uint8_t* interpolative_decode(const uint8_t* in, uint32_t* out, uint32_t sum, size_t n) {   /* block_codecs.hpp:127-147 */
    if (sum == 0xffffffffu) in = vbyte_decode(in, &sum);
    out[n - 1] = sum;
    size_t nbytes = 0;
    if (n > 1) {
        bit_reader br = {in, 0, 0, 0};
        br_interpolative(&br, out, n - 1, 0, sum);
        for (size_t i = n - 1; i > 0; --i) out[i] -= out[i - 1];
        nbytes = (br.pos + 7) / 8;
    }
    return in + nbytes;
}

/* Simple16 (FastPFor/headers/simple16.h:730-1110): selector -> runs of (count, bits), MSB-first in 28 bits */
static const uint8_t S16[16][6] = {{28, 1, 0, 0, 0, 0}, {7, 2, 14, 1, 0, 0}, {7, 1, 7, 2, 7, 1}, {14, 1, 7, 2, 0, 0}, {14, 2, 0, 0, 0, 0},
                                   {1, 4, 8, 3, 0, 0},  {1, 3, 4, 4, 3, 3},  {7, 4, 0, 0, 0, 0}, {4, 5, 2, 4, 0, 0},  {2, 4, 4, 5, 0, 0},
                                   {3, 6, 2, 5, 0, 0},  {2, 5, 3, 6, 0, 0},  {4, 7, 0, 0, 0, 0}, {1, 10, 2, 9, 0, 0}, {2, 14, 0, 0, 0, 0},
                                   {1, 28, 0, 0, 0, 0}};
static void simple16_decode(const uint8_t* in, size_t nvalue, uint32_t* out) {   /* simple16.h:689-725 */
    size_t got = 0;
    while (got < nvalue) {
        uint32_t w = rd32(in); in += 4;
        const uint8_t* r = S16[w >> 28];
        unsigned shift = 28;
        for (int k = 0; k < 6; k += 2)
            for (unsigned i = 0; i < r[k]; ++i) { shift -= r[k + 1]; out[got++] = (w >> shift) & ((1u << r[k + 1]) - 1); }
    }
}
static const uint8_t* optpfor_decode(const uint8_t* in, uint32_t* out) {   /* NewPFor::decodeBlock, FastPFor/headers/newpfor.h:254-286 */
    uint32_t w0 = rd32(in);
    uint32_t b = w0 >> 26, nexc = (w0 >> 16) & 0x3ff, excw = w0 & 0xffff;
    in += 4;
    static uint32_t exc[2 * 128 + 64];
    if (excw) simple16_decode(in, 2 * nexc, exc);
    in += 4 * excw;
    for (unsigned g = 0; g < 4; ++g) {                       /* fastunpack: 32 values, b bits each, LSB-first */
        uint64_t acc = 0; unsigned have = 0; const uint8_t* wp = in;
        for (unsigned i = 0; i < 32; ++i) {
            while (have < b) { acc |= (uint64_t)rd32(wp) << have; wp += 4; have += 32; }
            out[32 * g + i] = b == 32 ? (uint32_t)acc : (uint32_t)(acc & (((uint64_t)1 << b) - 1));
            acc >>= b; have -= b;
        }
        in += 4 * b;
    }
    uint32_t lpos = (uint32_t)-1;
    for (uint32_t e = 0; e < nexc; ++e) { lpos += exc[e] + 1; out[lpos] |= (exc[e + nexc] + 1) << b; }
    return in;
}
static const uint8_t* varint_decode(const uint8_t* in, uint32_t* out, size_t n) {   /* VarIntG8IU.h:42-82; block_codecs.hpp:287-314 */
    size_t got = 0;
    while (got < n) {
        uint8_t desc = *in;
        const uint8_t* d = in + 1;
        uint32_t cur = 0; unsigned shift = 0;
        for (unsigned i = 0; i < 8; ++i) {
            cur |= (uint32_t)d[i] << shift; shift += 8;
            if (!((desc >> i) & 1)) { if (got < n) out[got] = cur; ++got; cur = 0; shift = 0; }
        }
        in += 9;
    }
    return in;
}
/* qmx_block::decode (block_codecs.hpp:336-349) over QMX::codec<128>::decode (qmx_codec.hpp:636-6115), scalar.
 * TightVByte(len), then len bytes: payload stripes from the front, key bytes from the back (consumed in reverse while the
 * payload cursor has not passed them, :655-656).  key = type << 4 | (16 - run).  A 128-bit stripe is four interleaved u32
 * lanes (value v in lane v & 3, row v >> 2); types 8 / 12 / 14 are plain 8 / 16 / 32-bit arrays; the 7-, 9-, 12- and 21-bit
 * types straddle a (lo, hi) stripe pair (:4833-4856, :5310-5338, :5700-5724, :5980-5998).  Only the first 128 values of
 * the block are kept (the codec may overshoot, block_codecs.hpp:319). */
static const uint8_t* qmx_decode(const uint8_t* in, uint32_t* out) {
    static const uint16_t COUNT[16] = {256, 128, 64, 40, 32, 24, 20, 36, 16, 28, 12, 20, 8, 12, 4, 0};
    static const uint8_t WIDTH[16] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16, 21, 32, 0};
    uint32_t len;
    in = vbyte_decode(in, &len);
    uint32_t p = 0, o = 0;
    for (uint32_t t = 0; t < len && o < 128; ++t) {
        if (p > len - 1 - t) break;                       /* the payload cursor has passed the key */
        uint32_t key = in[len - 1 - t], type = key >> 4, run = 16 - (key & 15);
        int wide = type == 7 || type == 9 || type == 11 || type == 13;
        uint32_t ub = type == 0 ? 0 : type == 15 ? 1 : wide ? 32 : 16, w = WIDTH[type];
        for (uint32_t r = 0; r < run; ++r) {
            const uint8_t* a = in + p + r * ub;
            for (uint32_t v = 0; v < COUNT[type]; ++v, ++o) {
                uint32_t val;
                if (type == 0) val = 1;                   /* "0 bits" decodes to 1 (:129-133,657-660) */
                else if (type == 8) val = a[v];
                else if (type == 12) val = rd32(a + 2 * v) & 0xffff;
                else if (type == 14) val = rd32(a + 4 * v);
                else {
                    uint32_t l = v & 3, row = v >> 2, mask = ((uint32_t)1 << w) - 1, lo = rd32(a + 4 * l);
                    if (wide) {
                        uint32_t hi = rd32(a + 16 + 4 * l), rs = 32 / w;
                        uint32_t resume = type == 7 ? 3 : type == 9 ? 4 : type == 11 ? 8 : 11;
                        if (row < rs) val = (lo >> (row * w)) & mask;
                        else if (row == rs) val = ((lo >> (rs * w)) | (hi << (32 - rs * w))) & mask;
                        else val = (hi >> (resume + w * (row - rs - 1))) & mask;
                    } else val = (lo >> (row * w)) & mask;
                }
                if (o < 128) out[o] = val;
            }
        }
        p += run * ub;
    }
    return in + len;
}

static const uint8_t* block_decode(int codec, const uint8_t* in, uint32_t* out, uint32_t sum, size_t n) {
    if (codec == INTERPOLATIVE || n < 128) return interpolative_decode(in, out, sum, n);   /* block_codecs.hpp:196-199,215-217 */
    if (codec == MIXED) {       /* mixed_block::decode, mixed_block.hpp:198-217: block_type byte (pfor 0, varint 1, interpolative 2) */
        int type = *in++;
        if (type == 2) return interpolative_decode(in, out, sum, n);
        codec = type == 0 ? OPTPFOR : VARINT;
    }
    if (codec == QMX) return qmx_decode(in, out);
    return codec == OPTPFOR ? optpfor_decode(in, out) : varint_decode(in, out, n);
}

/* ==== freq_index family: ef / single / uniform / opt (index_types.hpp:18-35) =========================================
 * Full decode of a list's docs and freqs sequences; the query operators below then run over the decoded arrays.  What is
 * restated is the READING of the formats (headers, partitions, the three partition encodings, the strict / positive
 * transformations), not the reference's skipping machinery (pointers, rank samples), which never changes a value. */
enum { V_OPT, V_UNIFORM, V_SINGLE, V_EF };
typedef struct {
    int variant;
    uint64_t size, num_docs;
    unsigned ls0, ls1, rb0, rb1, logp;           /* global_parameters (global_parameters.hpp:6-20) */
    const uint8_t *dwords, *fwords;              /* m_bitvectors of the two bitvector_collections */
    uint64_t dbits, fbits;
    uint64_t *dstart, *fstart;                   /* decoded m_endpoints (bitvector_collection.hpp:57-67) */
} ef_index;
static const ef_index* g_ef = 0;                 /* set: enumerators are backed by fully decoded EF-family lists */

typedef struct { const uint8_t* w; uint64_t pos; } bit_cursor;
static uint64_t bc_take(bit_cursor* c, unsigned len) { uint64_t v = bv_bits(c->w, c->pos, len); c->pos += len; return v; }
static uint64_t bc_gamma(bit_cursor* c) {        /* integer_codes.hpp:21-25 */
    unsigned l = 0;
    while (!bv_bits(c->w, c->pos, 1)) { ++c->pos; ++l; }
    ++c->pos;
    return (bc_take(c, l) | ((uint64_t)1 << l)) - 1;
}
static uint64_t bc_delta(bit_cursor* c) { uint64_t l = bc_gamma(c); return (bc_take(c, (unsigned)l) | ((uint64_t)1 << l)) - 1; }   /* :41-45 */

/* one indexed_sequence / strict_sequence (indexed_sequence.hpp:89-127, strict_sequence.hpp:50-137) or, raw, one
 * compact_elias_fano / strict_elias_fano (ef_index): n values in [0, universe) */
static void base_decode(const ef_index* ix, const uint8_t* w, uint64_t off, uint64_t universe, uint64_t n, int strict, uint64_t* out) {
    if (ix->variant == V_EF) {                   /* strict_elias_fano.hpp:20-36: v - i over universe - n + 1, global sampling */
        ef_decode_at(w, off, strict ? universe - n + 1 : universe, n, ix->ls0, ix->ls1, out);
        if (strict) for (uint64_t i = 0; i < n; ++i) out[i] += i;
        return;
    }
    if (universe == n) { for (uint64_t i = 0; i < n; ++i) out[i] = i; return; }      /* all_ones_sequence.hpp:25-75, no type bit */
    int type = (int)bv_bits(w, off, 1);
    off += 1;
    if (type == 0) {                             /* elias_fano; the strict variants never index zeros (strict_sequence.hpp:24-30) */
        ef_decode_at(w, off, strict ? universe - n + 1 : universe, n, strict ? 63 : ix->ls0, ix->ls1, out);
        if (strict) for (uint64_t i = 0; i < n; ++i) out[i] += i;
    } else {                                     /* compact_ranked_bitvector.hpp:14-50: rank samples | select pointers | bitmap */
        unsigned rs0 = strict ? 63 : ix->rb0;
        uint64_t samples = rs0 >= 63 ? 0 : universe >> rs0, p1 = n >> ix->rb1;
        uint64_t bits_off = off + samples * ceil_log2(n + 1) + p1 * ceil_log2(universe);
        uint64_t k = 0;
        for (uint64_t v = 0; v < universe && k < n; ++v)
            if (bv_bits(w, bits_off + v, 1)) out[k++] = v;
    }
}

/* partitioned_sequence (partitioned_sequence.hpp:21-178) / uniform_partitioned_sequence (uniform_partitioned_sequence.hpp:19-160) */
static void sequence_decode(const ef_index* ix, const uint8_t* w, uint64_t off, uint64_t universe, uint64_t n, int strict, uint64_t* out) {
    if (ix->variant == V_SINGLE || ix->variant == V_EF) { base_decode(ix, w, off, universe, n, strict, out); return; }
    bit_cursor it = {w, off};
    uint64_t partitions = bc_gamma(&it) + 1;
    if (partitions == 1) {
        uint64_t base = bc_take(&it, (unsigned)ceil_log2(universe)), ub = 0;
        if (n > 1) { uint64_t d = bc_delta(&it); ub = d ? d : universe - base - 1; }
        base_decode(ix, w, it.pos, ub + 1, n, strict, out);
        for (uint64_t i = 0; i < n; ++i) out[i] += base;
        return;
    }
    uint64_t endpoint_bits = bc_gamma(&it);
    uint64_t cur = it.pos;
    uint64_t* sizes = 0;
    if (ix->variant == V_OPT) {                  /* cumulative partition sizes: EF over n, partitions - 1 values */
        sizes = (uint64_t*)malloc(8 * partitions);
        ef_decode_at(w, cur, n, partitions - 1, ix->ls0, ix->ls1, sizes);
        cur += ef_bitsize(n, partitions - 1, ix->ls0, ix->ls1);
    }
    uint64_t* ubs = (uint64_t*)malloc(8 * (partitions + 1));
    ef_decode_at(w, cur, universe, partitions + 1, ix->ls0, ix->ls1, ubs);      /* first value, then every partition's last value */
    cur += ef_bitsize(universe, partitions + 1, ix->ls0, ix->ls1);
    uint64_t endpoints_off = cur, sequences_off = cur + endpoint_bits * (partitions - 1);
    uint64_t psize = (uint64_t)1 << ix->logp;
    for (uint64_t p = 0; p < partitions; ++p) {
        uint64_t endpoint = p ? bv_bits(w, endpoints_off + (p - 1) * endpoint_bits, (unsigned)endpoint_bits) : 0;
        uint64_t begin, end;
        if (sizes) { begin = p ? sizes[p - 1] : 0; end = p + 1 < partitions ? sizes[p] : n; }
        else { begin = p * psize; end = (p + 1) * psize < n ? (p + 1) * psize : n; }
        uint64_t base = ubs[p] + (p ? 1 : 0), ub = ubs[p + 1];
        base_decode(ix, w, sequences_off + endpoint, ub - base + 1, end - begin, strict, out + begin);
        for (uint64_t i = begin; i < end; ++i) out[i] += base;
    }
    free(sizes); free(ubs);
}

/* freq_index::map (freq_index.hpp:234-243) over succinct::mapper's layout */
static int open_ef_index(ef_index* ix, const char* type, const uint8_t* p) {
    if (!strcmp(type, "opt")) ix->variant = V_OPT;
    else if (!strcmp(type, "uniform")) ix->variant = V_UNIFORM;
    else if (!strcmp(type, "single")) ix->variant = V_SINGLE;
    else if (!strcmp(type, "ef")) ix->variant = V_EF;
    else return -1;
    const uint8_t* c = p + 8;
    ix->ls0 = c[0]; ix->ls1 = c[1]; ix->rb0 = c[2]; ix->rb1 = c[3]; ix->logp = c[4]; c += 5;
    ix->num_docs = rd64(c); c += 8;
    for (int which = 0; which < 2; ++which) {
        uint64_t size = rd64(c); c += 8;
        uint64_t ep_bits = rd64(c), ep_nwords = rd64(c + 8); c += 16; (void)ep_bits;
        const uint8_t* ep_words = c; c += 8 * ep_nwords;
        uint64_t bits = rd64(c), nwords = rd64(c + 8); c += 16;
        const uint8_t* words = c; c += 8 * nwords;
        uint64_t* start = (uint64_t*)malloc(8 * (size + 1));
        ef_decode_all(ep_words, bits, size, ix->ls0, ix->ls1, start);             /* bitvector_collection.hpp:40-46 */
        ix->size = size;
        if (which == 0) { ix->dwords = words; ix->dbits = bits; ix->dstart = start; }
        else { ix->fwords = words; ix->fbits = bits; ix->fstart = start; }
    }
    return 0;
}

/* freq_index::operator[] (freq_index.hpp:192-214): gamma(occurrences), n, the docs sequence; freqs are the differences of
 * a strict positive sequence over occurrences + 1 (positive_sequence.hpp:18-78) */
static uint32_t ef_list_decode(const ef_index* ix, uint64_t term, uint32_t** docs_out, uint32_t** freqs_out) {
    bit_cursor it = {ix->dwords, ix->dstart[term]};
    uint64_t occurrences = bc_gamma(&it) + 1, n = 1;
    if (occurrences > 1) n = bc_take(&it, (unsigned)ceil_log2(occurrences + 1));
    uint64_t* tmp = (uint64_t*)malloc(8 * n);
    uint32_t *docs = (uint32_t*)malloc(4 * n), *freqs = (uint32_t*)malloc(4 * n);
    sequence_decode(ix, ix->dwords, it.pos, ix->num_docs, n, 0, tmp);
    for (uint64_t i = 0; i < n; ++i) docs[i] = (uint32_t)tmp[i];
    sequence_decode(ix, ix->fwords, ix->fstart[term], occurrences + 1, n, 1, tmp);
    for (uint64_t i = 0; i < n; ++i) freqs[i] = (uint32_t)(tmp[i] - (i ? tmp[i - 1] : 0));
    free(tmp);
    *docs_out = docs; *freqs_out = freqs;
    return (uint32_t)n;
}

/* ---- block_posting_list::document_enumerator (block_posting_list.hpp:84-355) ---- */
typedef struct {
    int codec;
    uint32_t n, blocks;
    const uint8_t *block_maxs, *block_endpoints, *blocks_data;
    uint64_t universe;
    uint32_t cur_block, pos_in_block, cur_block_max, cur_block_size, cur_docid;
    const uint8_t* freqs_block_data;
    int freqs_decoded;
    uint32_t docs_buf[128], freqs_buf[128];
    uint32_t *adocs, *afreqs, apos;      /* EF family: the whole list, decoded (adocs != 0) */
} enumerator;

static uint32_t block_max(const enumerator* e, uint32_t b) { return rd32(e->block_maxs + 4 * (size_t)b); }
static void decode_docs_block(enumerator* e, uint64_t block) {   /* :292-319 */
    uint32_t endpoint = block ? rd32(e->block_endpoints + 4 * (block - 1)) : 0;
    e->cur_block_size = ((block + 1) * 128 <= e->n) ? 128 : (e->n % 128);
    uint32_t cur_base = (block ? block_max(e, (uint32_t)block - 1) : (uint32_t)-1) + 1;
    e->cur_block_max = block_max(e, (uint32_t)block);
    e->freqs_block_data = block_decode(e->codec, e->blocks_data + endpoint, e->docs_buf,
                                       e->cur_block_max - cur_base - (e->cur_block_size - 1), e->cur_block_size);
    e->docs_buf[0] += cur_base;
    e->cur_block = (uint32_t)block; e->pos_in_block = 0; e->cur_docid = e->docs_buf[0]; e->freqs_decoded = 0;
}
static void enum_open(enumerator* e, const block_index* ix, uint64_t term) {   /* :86-108; block_freq_index.hpp:85-94 */
    if (g_ef) {
        free(e->adocs); free(e->afreqs);
        e->n = ef_list_decode(g_ef, term, &e->adocs, &e->afreqs);
        e->universe = g_ef->num_docs; e->apos = 0; e->cur_docid = e->adocs[0];
        return;
    }
    const uint8_t* data = ix->lists + ix->list_start[term];
    e->codec = ix->codec;
    e->block_maxs = vbyte_decode(data, &e->n);
    e->blocks = (e->n + 127) / 128;
    e->block_endpoints = e->block_maxs + 4 * (size_t)e->blocks;
    e->blocks_data = e->block_endpoints + 4 * ((size_t)e->blocks - 1);
    e->universe = ix->num_docs;
    decode_docs_block(e, 0);
}
static void enum_next(enumerator* e) {   /* :110-122 */
    if (e->adocs) { ++e->apos; e->cur_docid = e->apos < e->n ? e->adocs[e->apos] : (uint32_t)e->universe; return; }
    ++e->pos_in_block;
    if (e->pos_in_block == e->cur_block_size) {
        if (e->cur_block + 1 == e->blocks) { e->cur_docid = (uint32_t)e->universe; return; }
        decode_docs_block(e, e->cur_block + 1);
    } else e->cur_docid += e->docs_buf[e->pos_in_block] + 1;
}
static void enum_next_geq(enumerator* e, uint64_t lower_bound) {   /* :124-146 */
    if (e->adocs) {     /* first posting >= lower_bound at or after the cursor (the operators never ask for less) */
        while (e->apos < e->n && e->adocs[e->apos] < lower_bound) ++e->apos;
        e->cur_docid = e->apos < e->n ? e->adocs[e->apos] : (uint32_t)e->universe;
        return;
    }
    if (lower_bound > e->cur_block_max) {
        if (lower_bound > block_max(e, e->blocks - 1)) { e->cur_docid = (uint32_t)e->universe; return; }
        uint64_t block = e->cur_block + 1;
        while (block_max(e, (uint32_t)block) < lower_bound) ++block;
        decode_docs_block(e, block);
    }
    while (e->cur_docid < lower_bound) e->cur_docid += e->docs_buf[++e->pos_in_block] + 1;
}
static uint64_t enum_freq(enumerator* e) {   /* :165-171,321-331 */
    if (e->adocs) return e->afreqs[e->apos];
    if (!e->freqs_decoded) {
        block_decode(e->codec, e->freqs_block_data, e->freqs_buf, (uint32_t)-1, e->cur_block_size);
        e->freqs_decoded = 1;
    }
    return (uint64_t)e->freqs_buf[e->pos_in_block] + 1;
}

/* ---- bm25 (bm25.hpp) and topk_queue (queries.hpp:152-197) ---- */
static float doc_term_weight(uint64_t freq, float norm_len) {
    float f = (float)freq;
    return f / (f + 1.2f * (1.0f - 0.5f + 0.5f * norm_len));
}
static float query_term_weight(uint64_t freq, uint64_t df, uint64_t num_docs) {
    float f = (float)freq, fdf = (float)df;
    float idf = logf(((float)num_docs - fdf + 0.5f) / (fdf + 0.5f));
    float eps = 1.0E-6f;
    return f * (eps > idf ? eps : idf) * (1.0f + 1.2f);
}
typedef struct { uint64_t k; size_t size; float q[64]; } topk_queue;   /* kept as a sorted array: same multiset as the heap */
static int topk_would_enter(const topk_queue* t, float s) { return t->size < t->k || s > t->q[t->size - 1]; }
static int topk_insert(topk_queue* t, float s) {
    if (!topk_would_enter(t, s)) return 0;
    size_t n = t->size < t->k ? t->size++ : t->size - 1;
    size_t i = n;
    while (i > 0 && t->q[i - 1] < s) { t->q[i] = t->q[i - 1]; --i; }
    t->q[i] = s;
    return 1;
}

/* ---- query operators (queries.hpp) ---- */
#define MAXT 64
typedef struct { enumerator e; float q_weight, max_weight; uint64_t term; } scored_enum;
static scored_enum g_enums[MAXT];

static int cmp_u32(const void* a, const void* b) { uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b; return x < y ? -1 : x > y; }

/* query_freqs (:136-150): sorted distinct terms with multiplicities */
static size_t query_freqs(const uint32_t* terms, size_t n, uint32_t* out_terms, uint32_t* out_freqs) {
    uint32_t tmp[1024];
    memcpy(tmp, terms, 4 * n);
    qsort(tmp, n, 4, cmp_u32);
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) {
        if (i == 0 || tmp[i] != tmp[i - 1]) { out_terms[m] = tmp[i]; out_freqs[m] = 1; ++m; }
        else out_freqs[m - 1] += 1;
    }
    return m;
}
static size_t setup(const block_index* ix, const wand_data* w, const uint32_t* terms, size_t n) {
    uint32_t t[MAXT], f[MAXT];
    size_t m = query_freqs(terms, n, t, f);
    for (size_t i = 0; i < m; ++i) {
        enum_open(&g_enums[i].e, ix, t[i]);
        g_enums[i].term = t[i];
        g_enums[i].q_weight = query_term_weight(f[i], g_enums[i].e.n, ix->num_docs);
        g_enums[i].max_weight = w ? g_enums[i].q_weight * rdf(w->max_term_weight + 4 * (size_t)t[i]) : 0.f;
    }
    return m;
}
/* std::sort on <= 16 elements is libstdc++'s insertion sort (stable); larger inputs are not exercised by the fixtures */
static void sort_idx(size_t* idx, size_t m, int by /*0 size, 1 max_weight, 2 docid*/) {
    for (size_t i = 1; i < m; ++i) {
        size_t v = idx[i], j = i;
        for (; j > 0; --j) {
            const scored_enum *a = &g_enums[v], *b = &g_enums[idx[j - 1]];
            int less = by == 0 ? a->e.n < b->e.n : by == 1 ? a->max_weight < b->max_weight : a->e.cur_docid < b->e.cur_docid;
            if (!less) break;
            idx[j] = idx[j - 1];
        }
        idx[j] = v;
    }
}

static uint64_t and_query(const block_index* ix, const wand_data* w, const uint32_t* terms, size_t n, int ranked, topk_queue* topk) {   /* :35-86, 322-401 */
    topk->size = 0;
    if (!n) return 0;
    size_t m = setup(ix, w, terms, n), idx[MAXT];
    for (size_t i = 0; i < m; ++i) idx[i] = i;
    sort_idx(idx, m, 0);
    uint64_t results = 0, candidate = g_enums[idx[0]].e.cur_docid;
    size_t i = 1;
    while (candidate < ix->num_docs) {
        for (; i < m; ++i) {
            enumerator* e = &g_enums[idx[i]].e;
            enum_next_geq(e, candidate);
            if (e->cur_docid != candidate) { candidate = e->cur_docid; i = 0; break; }
        }
        if (i == m) {
            results += 1;
            if (ranked) {
                float norm_len = rdf(w->norm_lens + 4 * candidate), score = 0;
                for (i = 0; i < m; ++i) score += g_enums[idx[i]].q_weight * doc_term_weight(enum_freq(&g_enums[idx[i]].e), norm_len);
                topk_insert(topk, score);
            }
            enum_next(&g_enums[idx[0]].e);
            candidate = g_enums[idx[0]].e.cur_docid;
            i = 1;
        }
    }
    return ranked ? topk->size : results;
}

static uint64_t or_query(const block_index* ix, const wand_data* w, const uint32_t* terms, size_t n, int ranked, topk_queue* topk) {   /* :88-131, 404-476 */
    topk->size = 0;
    if (!n) return 0;
    size_t m = setup(ix, w, terms, n);
    uint64_t results = 0, cur_doc = ix->num_docs;
    for (size_t i = 0; i < m; ++i) if (g_enums[i].e.cur_docid < cur_doc) cur_doc = g_enums[i].e.cur_docid;
    while (cur_doc < ix->num_docs) {
        results += 1;
        float score = 0, norm_len = ranked ? rdf(w->norm_lens + 4 * cur_doc) : 0.f;
        uint64_t next_doc = ix->num_docs;
        for (size_t i = 0; i < m; ++i) {
            enumerator* e = &g_enums[i].e;
            if (e->cur_docid == cur_doc) {
                if (ranked) score += g_enums[i].q_weight * doc_term_weight(enum_freq(e), norm_len);
                enum_next(e);
            }
            if (e->cur_docid < next_doc) next_doc = e->cur_docid;
        }
        if (ranked) topk_insert(topk, score);
        cur_doc = next_doc;
    }
    return ranked ? topk->size : results;
}

static uint64_t wand_query(const block_index* ix, const wand_data* w, const uint32_t* terms, size_t n, topk_queue* topk) {   /* :200-319 */
    topk->size = 0;
    if (!n) return 0;
    size_t m = setup(ix, w, terms, n), ord[MAXT];
    for (size_t i = 0; i < m; ++i) ord[i] = i;
    sort_idx(ord, m, 2);
    while (1) {
        float upper_bound = 0;
        size_t pivot; int found = 0;
        for (pivot = 0; pivot < m; ++pivot) {
            if (g_enums[ord[pivot]].e.cur_docid == ix->num_docs) break;
            upper_bound += g_enums[ord[pivot]].max_weight;
            if (topk_would_enter(topk, upper_bound)) { found = 1; break; }
        }
        if (!found) break;
        uint64_t pivot_id = g_enums[ord[pivot]].e.cur_docid;
        if (pivot_id == g_enums[ord[0]].e.cur_docid) {
            float score = 0, norm_len = rdf(w->norm_lens + 4 * pivot_id);
            for (size_t p = 0; p < m; ++p) {
                scored_enum* en = &g_enums[ord[p]];
                if (en->e.cur_docid != pivot_id) break;
                score += en->q_weight * doc_term_weight(enum_freq(&en->e), norm_len);
                enum_next(&en->e);
            }
            topk_insert(topk, score);
            sort_idx(ord, m, 2);
        } else {
            size_t next_list = pivot;
            for (; g_enums[ord[next_list]].e.cur_docid == pivot_id; --next_list) {}
            enum_next_geq(&g_enums[ord[next_list]].e, pivot_id);
            for (size_t i = next_list + 1; i < m; ++i) {
                if (g_enums[ord[i]].e.cur_docid < g_enums[ord[i - 1]].e.cur_docid) { size_t t = ord[i]; ord[i] = ord[i - 1]; ord[i - 1] = t; }
                else break;
            }
        }
    }
    return topk->size;
}

static uint64_t maxscore_query(const block_index* ix, const wand_data* w, const uint32_t* terms, size_t n, topk_queue* topk) {   /* :478-591 */
    topk->size = 0;
    if (!n) return 0;
    size_t m = setup(ix, w, terms, n), ord[MAXT];
    for (size_t i = 0; i < m; ++i) ord[i] = i;
    sort_idx(ord, m, 1);
    float ub[MAXT];
    ub[0] = g_enums[ord[0]].max_weight;
    for (size_t i = 1; i < m; ++i) ub[i] = ub[i - 1] + g_enums[ord[i]].max_weight;
    uint64_t non_essential = 0, cur_doc = ix->num_docs;
    for (size_t i = 0; i < m; ++i) if (g_enums[i].e.cur_docid < cur_doc) cur_doc = g_enums[i].e.cur_docid;
    while (non_essential < m && cur_doc < ix->num_docs) {
        float score = 0, norm_len = rdf(w->norm_lens + 4 * cur_doc);
        uint64_t next_doc = ix->num_docs;
        for (size_t i = non_essential; i < m; ++i) {
            scored_enum* en = &g_enums[ord[i]];
            if (en->e.cur_docid == cur_doc) { score += en->q_weight * doc_term_weight(enum_freq(&en->e), norm_len); enum_next(&en->e); }
            if (en->e.cur_docid < next_doc) next_doc = en->e.cur_docid;
        }
        for (size_t i = non_essential - 1; i + 1 > 0; --i) {
            if (!topk_would_enter(topk, score + ub[i])) break;
            scored_enum* en = &g_enums[ord[i]];
            enum_next_geq(&en->e, cur_doc);
            if (en->e.cur_docid == cur_doc) score += en->q_weight * doc_term_weight(enum_freq(&en->e), norm_len);
        }
        if (topk_insert(topk, score))
            while (non_essential < m && !topk_would_enter(topk, ub[non_essential])) non_essential += 1;
        cur_doc = next_doc;
    }
    return topk->size;
}

static uint64_t run_op(const char* op, const block_index* ix, const wand_data* w, const uint32_t* t, size_t n, topk_queue* topk) {
    if (!strcmp(op, "and")) return and_query(ix, w, t, n, 0, topk);
    if (!strcmp(op, "ranked_and")) return and_query(ix, w, t, n, 1, topk);
    if (!strcmp(op, "or")) return or_query(ix, w, t, n, 0, topk);
    if (!strcmp(op, "ranked_or")) return or_query(ix, w, t, n, 1, topk);
    if (!strcmp(op, "wand")) return wand_query(ix, w, t, n, topk);
    if (!strcmp(op, "maxscore")) return maxscore_query(ix, w, t, n, topk);
    fprintf(stderr, "unknown op %s\n", op);
    exit(1);
}

/* ---- driver ---- */
typedef struct { uint32_t* terms; size_t* begin; size_t n; } query_log;
static query_log read_queries(const char* path) {   /* read_query, queries.hpp:15-27 */
    query_log q = {(uint32_t*)malloc(4 << 22), (size_t*)malloc(sizeof(size_t) * (1 << 20)), 0};
    FILE* f = fopen(path, "r");
    if (!f) { perror(path); exit(1); }
    static char line[1 << 16];
    size_t nt = 0;
    q.begin[0] = 0;
    while (fgets(line, sizeof line, f)) {
        char* p = line;
        while (1) {
            char* end;
            unsigned long v = strtoul(p, &end, 10);
            if (end == p) break;
            q.terms[nt++] = (uint32_t)v;
            p = end;
        }
        q.begin[++q.n] = nt;
    }
    fclose(f);
    return q;
}

#ifndef DS2I_ORACLE_NO_MAIN
int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: ds2i_oracle dump|lists|bench ...\n"); return 1; }
    uint8_t* ibytes; size_t in;
    if (load_file(argv[3], &ibytes, &in)) { perror(argv[3]); return 1; }
    block_index ix;
    static ef_index efx;
    if (open_index(&ix, argv[2], ibytes)) {
        if (open_ef_index(&efx, argv[2], ibytes)) { fprintf(stderr, "unsupported index type %s\n", argv[2]); return 1; }
        g_ef = &efx;
        memset(&ix, 0, sizeof ix);
        ix.size = efx.size; ix.num_docs = efx.num_docs;
    }
    if (!strcmp(argv[1], "lists")) {
        FILE* f = fopen(argv[4], "wb");
        static enumerator e;
        for (uint64_t t = 0; t < ix.size; ++t) {
            enum_open(&e, &ix, t);
            uint64_t n = e.n;
            fwrite(&n, 8, 1, f);
            uint32_t* d = (uint32_t*)malloc(8 * n);
            for (uint64_t i = 0; i < n; ++i) { d[i] = e.cur_docid; d[n + i] = (uint32_t)enum_freq(&e); enum_next(&e); }
            if (e.cur_docid != ix.num_docs) { fprintf(stderr, "sentinel mismatch\n"); return 2; }
            fwrite(d, 4, 2 * n, f);
            free(d);
        }
        fclose(f);
        return 0;
    }
    uint8_t* wbytes; size_t wn;
    if (load_file(argv[4], &wbytes, &wn)) { perror(argv[4]); return 1; }
    wand_data w;
    open_wand(&w, wbytes);
    query_log q = read_queries(argv[5]);
    topk_queue topk;
    if (!strcmp(argv[1], "bench")) {
        topk.k = 10;
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        uint64_t acc = 0;
        for (size_t i = 0; i < q.n; ++i) acc += run_op(argv[6], &ix, &w, q.terms + q.begin[i], q.begin[i + 1] - q.begin[i], &topk);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        double s = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
        printf("{\"type\": \"%s\", \"query\": \"%s\", \"threads\": 1, \"queries\": %zu, \"seconds\": %.6f, \"qps\": %.3f, \"checksum\": %llu}\n",
               argv[2], argv[6], q.n, s, (double)q.n / s, (unsigned long long)acc);
        return 0;
    }
    /* dump */
    uint64_t k = argc > 8 ? strtoull(argv[8], 0, 10) : 10;
    topk.k = k;
    FILE* f = fopen(argv[6], "wb");
    char ops[256];
    strncpy(ops, argv[7], sizeof ops - 1); ops[sizeof ops - 1] = 0;
    uint64_t nops = 1;
    for (char* c = ops; *c; ++c) nops += *c == ':';
    uint64_t hdr[3] = {q.n, k, nops};
    fwrite(hdr, 8, 3, f);
    for (char* op = strtok(ops, ":"); op; op = strtok(0, ":")) {
        for (size_t i = 0; i < q.n; ++i) {
            uint64_t c = run_op(op, &ix, &w, q.terms + q.begin[i], q.begin[i + 1] - q.begin[i], &topk);
            float s[64] = {0};
            int ranked = !strcmp(op, "ranked_and") || !strcmp(op, "wand") || !strcmp(op, "maxscore") || !strcmp(op, "ranked_or");
            if (ranked) memcpy(s, topk.q, 4 * topk.size);
            fwrite(&c, 8, 1, f);
            fwrite(s, 4, k, f);
        }
    }
    fclose(f);
    return 0;
}
#endif
