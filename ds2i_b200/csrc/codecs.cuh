// Warp-cooperative decoders of ds2i's block codecs, reading the compressed bytes from a
// shared-memory window that TMA staged (device_common.cuh).  One warp decodes one block of up to
// 128 integers; every function is warp-collective (all 32 lanes call it with the same arguments)
// and returns, warp-uniformly, the number of bytes the block occupies — the reference decoders
// return the pointer past the block (block_codecs.hpp:127-147,210-226,287-314,336-349) and the
// freqs block starts exactly there (block_posting_list.hpp:304-307).
#pragma once
#include "device_common.cuh"

namespace ds2i_gpu {

enum : int { CODEC_OPTPFOR = 0, CODEC_VARINT = 1, CODEC_INTERPOLATIVE = 2, CODEC_QMX = 3,
             CODEC_MIXED = 4 /* mixed_block.hpp: one type byte in front of every full block */,
             CODEC_PEF = 5 /* not a block codec: 128-element windows of partitioned Elias-Fano sequences (pef.cuh) */, CODEC_ANY = 99 /* decided at run time from the index */ };

constexpr uint32_t BLOCK = 128;            // BlockCodec::block_size for every codec
constexpr uint32_t SCRATCH_WORDS = 32;     // interpolative stack: <= 7 pending right halves x 4 words

// ---- Simple16 layouts (FastPFor/headers/simple16.h:730-1110) ----------------------------------
// selector -> up to three runs (count, bits); values fill the 28 payload bits MSB-first.
// packed as n1 | b1<<5 | n2<<10 | b2<<15 | n3<<20 | b3<<25
#define S16(n1, b1, n2, b2, n3, b3) \
    (uint32_t(n1) | (uint32_t(b1) << 5) | (uint32_t(n2) << 10) | (uint32_t(b2) << 15) | (uint32_t(n3) << 20) | (uint32_t(b3) << 25))
__device__ __constant__ uint32_t c_s16_layout[16] = {
    S16(28, 1, 0, 0, 0, 0), S16(7, 2, 14, 1, 0, 0), S16(7, 1, 7, 2, 7, 1), S16(14, 1, 7, 2, 0, 0),
    S16(14, 2, 0, 0, 0, 0), S16(1, 4, 8, 3, 0, 0),  S16(1, 3, 4, 4, 3, 3), S16(7, 4, 0, 0, 0, 0),
    S16(4, 5, 2, 4, 0, 0),  S16(2, 4, 4, 5, 0, 0),  S16(3, 6, 2, 5, 0, 0), S16(2, 5, 3, 6, 0, 0),
    S16(4, 7, 0, 0, 0, 0),  S16(1, 10, 2, 9, 0, 0), S16(2, 14, 0, 0, 0, 0), S16(1, 28, 0, 0, 0, 0)};
#undef S16

// Per-CTA lookup tables in shared memory: tab[0..15] = values per selector, tab[16 + sel*28 + j] =
// (shift | width << 8) of the j-th value of a word with that selector.
constexpr uint32_t S16_TAB_WORDS = 16 + 16 * 28;
constexpr uint32_t S16_TAB_BYTES = S16_TAB_WORDS * 4;     // sits at offset 0 of g_smem in every kernel

__device__ __forceinline__ uint32_t s16_count(uint32_t lay) {
    return (lay & 31u) + ((lay >> 10) & 31u) + ((lay >> 20) & 31u);
}
__device__ __forceinline__ void s16_table_init(uint32_t* tab /* S16_TAB_WORDS words of shared memory */) {
    for (uint32_t t = threadIdx.x; t < S16_TAB_WORDS; t += blockDim.x) {
        if (t < 16) { tab[t] = s16_count(c_s16_layout[t]); continue; }
        uint32_t sel = (t - 16) / 28, j = (t - 16) % 28;
        uint32_t lay = c_s16_layout[sel];
        uint32_t n1 = lay & 31u, b1 = (lay >> 5) & 31u, n2 = (lay >> 10) & 31u, b2 = (lay >> 15) & 31u, n3 = (lay >> 20) & 31u, b3 = lay >> 25;
        uint32_t off = 0, w = 0;
        if (j < n1) { off = j * b1; w = b1; }
        else if (j < n1 + n2) { off = n1 * b1 + (j - n1) * b2; w = b2; }
        else if (j < n1 + n2 + n3) { off = n1 * b1 + n2 * b2 + (j - n1 - n2) * b3; w = b3; }
        tab[t] = w ? ((28u - off - w) | (w << 8)) : 0u;
    }
}
// j-th value of a Simple16 word (values fill the 28 payload bits MSB-first)
__device__ __forceinline__ uint32_t s16_value(const uint32_t* tab, uint32_t word, uint32_t j) {
    uint32_t t = tab[16u + (word >> 28) * 28u + j];
    return (word >> (t & 31u)) & ((1u << (t >> 8)) - 1u);
}

// ---- TightVariableByte, one value (block_codecs.hpp:84-98); single-lane helper -----------------
__device__ __forceinline__ uint32_t vbyte_decode(const uint32_t* win, uint32_t& off) {
    uint32_t v = 0;
#pragma unroll 1
    for (uint32_t shift = 0; shift <= 28; shift += 7) {
        uint32_t c = lds_u8(win, off++);
        v += (c & 127u) << shift;
        if (c & 128u) break;
    }
    return v;
}

// ---- OptPFD / NewPFD block of exactly 128 values (FastPFor/headers/newpfor.h:254-286) ---------
// Header word b<<26 | nExc<<16 | excWords; Simple16 exception stream; 4 groups of 32 values packed
// LSB-first at b bits (bitpackinghelpers.h fastunpack).  Lane l unpacks values 4l..4l+3 (one funnel
// shift when b <= 8); the byte phase of the (unaligned) block is folded into the bit position, so no
// realignment pass.  Exceptions: the Simple16 words sit one per lane; every lane locates the word
// holding "its" two values (position gap e and high bits nExc+e) by ranking e in a bitmap of the
// words' first value indices (two REDUX + a popcount; a shuffle search when there are more than 32
// exceptions) and extracts them with one table lookup each — no per-word expansion loop.
// WIDE_EXC adds a bitmap path for 33..64 exceptions (instead of the shuffle search): worth its code in the batched
// decode kernel (-3.4 %), not in the query kernels, where the extra instruction footprint costs more than it saves.
template <bool WIDE_EXC = false>
__device__ __noinline__ uint32_t decode_optpfor128(uint32_t win_off, uint32_t off, uint32_t out_off, uint32_t /*scratch_off*/) {
    const uint32_t* win = smem_words(win_off);
    uint32_t* out = smem_words(out_off);
    const uint32_t* s16tab = smem_words(0);
    const unsigned lane = lane_id();
    const uint32_t w0 = lds_u32(win, off);
    const uint32_t b = w0 >> 26;
    const uint32_t nexc = (w0 >> 16) & 0x3ffu;
    const uint32_t excw = w0 & 0xffffu;
    if (b >= 32) {   // raw block: newpfor.h:204-209
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) out[32 * j + lane] = lds_u32(win, off + 4u * (1u + 32u * j + lane));
        __syncwarp();
        return 4u * 129u;
    }
    // the four groups of 32 b-bit values are back to back (b words each): one LSB-first stream of 128
    // values.  Lane l takes values 4l..4l+3, i.e. the 4b bits at bit 4lb.
    {
        const uint32_t pbit = 8u * (off + 4u * (1u + excw)) + 4u * lane * b;
        uint4 v;
        if (b <= 8) {
            const uint32_t w = pbit >> 5;
            const uint32_t x = __funnelshift_r(win[w], win[w + 1], pbit & 31u);
            const uint32_t m = (1u << b) - 1u;
            v.x = x & m; v.y = (x >> b) & m; v.z = (x >> (2u * b)) & m; v.w = (x >> (3u * b)) & m;
        } else {
            v.x = lds_bits(win, pbit, b); v.y = lds_bits(win, pbit + b, b);
            v.z = lds_bits(win, pbit + 2u * b, b); v.w = lds_bits(win, pbit + 3u * b, b);
        }
        reinterpret_cast<uint4*>(out)[lane] = v;
    }
    if (nexc) {
        __syncwarp();
        if (excw <= 32) {
            const uint32_t word = lane < excw ? lds_u32(win, off + 4u * (1u + lane)) : 0u;
            const uint32_t cnt = lane < excw ? s16tab[word >> 28] : 0u;
            const uint32_t start = warp_inclusive_scan(cnt) - cnt;     // index of this word's first value
            if (nexc <= 32) {
                // common case: all 2*nexc <= 64 values.  Two 32-bit maps of the value indices at which a
                // word starts; the word holding value e is the number of starts <= e, minus one.
                const bool has = cnt != 0u;
                const uint32_t bm0 = __reduce_or_sync(FULL, (has && start < 32u) ? 1u << start : 0u);
                const uint32_t bm1 = __reduce_or_sync(FULL, (has && start >= 32u && start < 64u) ? 1u << (start - 32u) : 0u);
                const uint32_t pc0 = __popc(bm0);
                auto fetch = [&](uint32_t e) -> uint32_t {
                    const bool lo = e < 32u;
                    const uint32_t m = (lo ? bm0 : bm1) & (0xffffffffu >> (31u - (e & 31u)));
                    const uint32_t w = __popc(m) + (lo ? 0u : pc0) - 1u;
                    const uint32_t wv = __shfl_sync(FULL, word, w & 31u);
                    const uint32_t sv = __shfl_sync(FULL, start, w & 31u);
                    const uint32_t j = e - sv;
                    return s16_value(s16tab, wv, j < 28u ? j : 0u);
                };
                const bool valid = lane < nexc;
                const uint32_t g = fetch(valid ? lane : 0u) + 1u;
                const uint32_t hi = fetch(valid ? nexc + lane : 0u) + 1u;
                const uint32_t p = warp_inclusive_scan(valid ? g : 0u) - 1u;
                if (valid && p < BLOCK) out[p] |= hi << b;
            } else if (WIDE_EXC && nexc <= 64) {
                // up to 128 values: four maps of the start indices; the word holding value e = starts <= e, minus one
                const bool has = cnt != 0u;
                const uint32_t sw = start >> 5, sb = 1u << (start & 31u);
                const uint32_t bm0 = __reduce_or_sync(FULL, (has && sw == 0u) ? sb : 0u);
                const uint32_t bm1 = __reduce_or_sync(FULL, (has && sw == 1u) ? sb : 0u);
                const uint32_t bm2 = __reduce_or_sync(FULL, (has && sw == 2u) ? sb : 0u);
                const uint32_t bm3 = __reduce_or_sync(FULL, (has && sw == 3u) ? sb : 0u);
                const uint32_t pc0 = __popc(bm0), pc1 = pc0 + __popc(bm1), pc2 = pc1 + __popc(bm2);
                auto fetch = [&](uint32_t e) -> uint32_t {
                    const uint32_t k = e >> 5;
                    const uint32_t sel = k == 0u ? bm0 : k == 1u ? bm1 : k == 2u ? bm2 : bm3;
                    const uint32_t before = k == 0u ? 0u : k == 1u ? pc0 : k == 2u ? pc1 : pc2;
                    const uint32_t w = __popc(sel & (0xffffffffu >> (31u - (e & 31u)))) + before - 1u;
                    const uint32_t wv = __shfl_sync(FULL, word, w & 31u);
                    const uint32_t sv = __shfl_sync(FULL, start, w & 31u);
                    const uint32_t j = e - sv;
                    return s16_value(s16tab, wv, j < 28u ? j : 0u);
                };
                uint32_t carry = 0;
#pragma unroll 1
                for (uint32_t e0 = 0; e0 < nexc; e0 += 32) {
                    const uint32_t e = e0 + lane;
                    const bool valid = e < nexc;
                    const uint32_t g = fetch(valid ? e : 0u) + 1u;
                    const uint32_t hi = fetch(valid ? nexc + e : 0u) + 1u;
                    const uint32_t incl = warp_inclusive_scan(valid ? g : 0u);
                    const uint32_t p = carry + incl - 1u;
                    if (valid && p < BLOCK) out[p] |= hi << b;
                    carry += __shfl_sync(FULL, incl, 31);
                }
            } else {
                const uint32_t s_hi = excw > 1 ? 1u << (31 - __clz(excw - 1)) : 0u;   // search steps follow the word count
                auto fetch = [&](uint32_t e) -> uint32_t {
                    uint32_t w = 0;
                    for (uint32_t s = s_hi; s >= 1; s >>= 1) {
                        uint32_t st = __shfl_sync(FULL, start, (w + s) & 31u);
                        if (st <= e) w += s;
                    }
                    uint32_t wv = __shfl_sync(FULL, word, w);
                    uint32_t sv = __shfl_sync(FULL, start, w);
                    uint32_t j = e - sv;
                    return s16_value(s16tab, wv, j < 28u ? j : 0u);
                };
                uint32_t carry = 0;
                for (uint32_t e0 = 0; e0 < nexc; e0 += 32) {
                    uint32_t e = e0 + lane;
                    bool valid = e < nexc;
                    uint32_t g = fetch(valid ? e : 0u) + 1u;
                    uint32_t hi = fetch(valid ? nexc + e : 0u) + 1u;
                    uint32_t incl = warp_inclusive_scan(valid ? g : 0u);
                    uint32_t p = carry + incl - 1u;
                    if (valid && p < BLOCK) out[p] |= hi << b;
                    carry += __shfl_sync(FULL, incl, 31);
                }
            }
        } else if (lane == 0) {
            // long exception streams (> 32 Simple16 words, practically never): lane 0 walks the two halves
            // of the value stream (position gaps from value 0, high bits from value nexc) in lock-step
            const uint32_t wbase = off + 4u;
            uint32_t wa = 0, ja = 0, wb = 0, jb = 0, skipped = 0;
            while (true) {
                uint32_t cw = s16tab[lds_u32(win, wbase + 4u * wb) >> 28];
                if (skipped + cw > nexc || wb + 1u >= excw) { jb = nexc - skipped; break; }
                skipped += cw; ++wb;
            }
            uint32_t worda = lds_u32(win, wbase), ca = s16tab[worda >> 28];
            uint32_t wordb = lds_u32(win, wbase + 4u * wb), cb = s16tab[wordb >> 28];
            uint32_t pos = 0xffffffffu;
            for (uint32_t e = 0; e < nexc; ++e) {
                uint32_t g = s16_value(s16tab, worda, ja < 28u ? ja : 0u);
                if (++ja >= ca) { ++wa; worda = lds_u32(win, wbase + 4u * wa); ca = s16tab[worda >> 28]; ja = 0; }
                uint32_t h = s16_value(s16tab, wordb, jb < 28u ? jb : 0u);
                if (++jb >= cb) { ++wb; wordb = wb < excw ? lds_u32(win, wbase + 4u * wb) : 0u; cb = s16tab[wordb >> 28]; jb = 0; }
                pos += g + 1u;
                if (pos < BLOCK) out[pos] |= (h + 1u) << b;
            }
        }
    }
    __syncwarp();
    return 4u * (1u + excw + 4u * b);
}


// ---- varint-G8IU block of exactly 128 values (FastPFor/headers/VarIntG8IU.h:152-195; ds2i decode
// block_codecs.hpp:239-258,287-314).  9-byte groups: descriptor + 8 data bytes; bit i of the
// descriptor is 0 iff data byte i ends an integer.  The group count is not stored: lane g reads the
// descriptors of groups g and g+32, a warp scan of the per-group integer counts finds the group that
// completes the 128th value (bytes after it belong to the next block and are never interpreted), and
// every lane then expands its own groups.
__device__ __noinline__ uint32_t decode_varint128(uint32_t win_off, uint32_t off, uint32_t out_off) {
    const uint32_t* win = smem_words(win_off);
    uint32_t* out = smem_words(out_off);
    const unsigned lane = lane_id();
    uint32_t carry = 0, groups = 0;
    for (uint32_t round = 0; round < 2 && carry < BLOCK; ++round) {
        const uint32_t g = round * 32 + lane;
        const uint32_t gb = off + 9u * g;
        const uint32_t desc = lds_u8(win, gb);
        const uint32_t lo = lds_u32(win, gb + 1), hi = lds_u32(win, gb + 5);
        uint32_t cnt = 8u - __popc(desc);
        uint32_t incl = warp_inclusive_scan(cnt);
        uint32_t base = carry + incl - cnt;
        bool live = base < BLOCK;                      // groups that start before the 128th value
        if (live) {
            uint32_t idx = base, cur = 0, shift = 0;
#pragma unroll
            for (uint32_t i = 0; i < 8; ++i) {
                uint32_t byte = ((i < 4 ? lo : hi) >> (8u * (i & 3u))) & 0xffu;
                cur |= byte << shift;
                shift += 8;
                if (!((desc >> i) & 1u)) {
                    if (idx < BLOCK) out[idx] = cur;
                    ++idx; cur = 0; shift = 0;
                }
            }
        }
        groups += __popc(__ballot_sync(FULL, live));
        carry += __shfl_sync(FULL, incl, 31);
    }
    __syncwarp();
    return 9u * groups;
}

// ---- QMX block of exactly 128 values (qmx_codec.hpp:636-6115; ds2i wrapper block_codecs.hpp:317-350).
// TightVByte(len), then `len` bytes: payload stripes first, the key bytes in REVERSE order at the
// end.  key = type << 4 | (16 - run); the decoder consumes keys backwards while the payload cursor
// has not passed them (:655-656).  Lane t owns the t-th key from the end: two warp scans give every
// key's first payload byte and first output index; then each lane extracts "its" outputs by locating
// the key with a shuffle search.  128-bit stripes are four interleaved u32 lanes (value v in lane
// v & 3, row v >> 2); 8/16/32-bit types are sequential; the four 256-bit types (7, 9, 12, 21 bits)
// straddle a (lo, hi) stripe pair as tabulated below (:4833-4856, :5310-5338, :5700-5724, :5980-5998).
__device__ __noinline__ uint32_t decode_qmx128(uint32_t win_off, uint32_t off, uint32_t out_off) {
    const uint32_t* win = smem_words(win_off);
    uint32_t* out = smem_words(out_off);
    const unsigned lane = lane_id();
    uint32_t p0 = off;
    const uint32_t len = vbyte_decode(win, p0);
    // per type: values per stripe unit, payload bytes per unit, bit width
    auto type_count = [](uint32_t t) -> uint32_t {
        // {256,128,64,40,32,24,20,36,16,28,12,20,8,12,4,0}
        const uint32_t tab[16] = {256, 128, 64, 40, 32, 24, 20, 36, 16, 28, 12, 20, 8, 12, 4, 0};
        uint32_t r = 0;
#pragma unroll
        for (uint32_t i = 0; i < 16; ++i) r = (t == i) ? tab[i] : r;
        return r;
    };
    auto type_bytes = [](uint32_t t) -> uint32_t {
        if (t == 0) return 0u;
        if (t == 15) return 1u;
        return (t == 7 || t == 9 || t == 11 || t == 13) ? 32u : 16u;
    };
    auto type_width = [](uint32_t t) -> uint32_t {
        const uint32_t tab[16] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16, 21, 32, 0};
        uint32_t r = 0;
#pragma unroll
        for (uint32_t i = 0; i < 16; ++i) r = (t == i) ? tab[i] : r;
        return r;
    };

    uint32_t pcarry = 0, icarry = 0;
    for (uint32_t kbase = 0; kbase < len && icarry < BLOCK; kbase += 32) {
        const uint32_t T = kbase + lane;
        const bool inrange = T < len;
        const uint32_t key = inrange ? lds_u8(win, p0 + len - 1u - T) : 0xffu;
        const uint32_t type = key >> 4, run = 16u - (key & 15u);
        const uint32_t cnt = type_count(type);
        const uint32_t ub = type_bytes(type);
        uint32_t nbytes = inrange ? run * ub : 0u, nints = inrange ? run * cnt : 0u;
        uint32_t pincl = warp_inclusive_scan(nbytes);
        uint32_t pexcl = pcarry + pincl - nbytes;
        // "while (in <= keys)": key T is consumed iff the payload cursor has not passed it
        bool valid = inrange && pexcl <= len - 1u - T;
        unsigned vmask = __ballot_sync(FULL, valid);
        unsigned first_bad = __ffs(~vmask);                       // keys are consumed in order: stop at the first refusal
        uint32_t nvalid = first_bad ? first_bad - 1u : 32u;
        if (lane >= nvalid) nints = 0;
        uint32_t iincl = warp_inclusive_scan(nints);
        uint32_t iexcl = icarry + iincl - nints;                  // first output index of this key
        const uint32_t round_end = icarry + __shfl_sync(FULL, iincl, 31);
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t i = 32u * j + lane;
            // key holding output i (largest t with iexcl_t <= i among keys with ints)
            uint32_t t = 0;
#pragma unroll
            for (uint32_t s = 16; s >= 1; s >>= 1) {
                uint32_t st = __shfl_sync(FULL, iexcl, (t + s) & 31u);
                if (st <= i) t += s;
            }
            // (a key without values shares its start index with its successor, which the search prefers)
            uint32_t kb = __shfl_sync(FULL, iexcl, t), kn = __shfl_sync(FULL, nints, t);
            uint32_t ktype = __shfl_sync(FULL, type, t), kp = __shfl_sync(FULL, pexcl, t);
            bool mine = i >= icarry && i < round_end && i < BLOCK && i >= kb && i < kb + kn;
            if (mine) {
                const uint32_t c = type_count(ktype), w = type_width(ktype), rel = i - kb;
                const uint32_t r = rel / c, v = rel - r * c;
                const uint32_t addr = p0 + kp + r * type_bytes(ktype);
                uint32_t val;
                if (ktype == 0) val = 1u;                                            // "0 bits" decodes to 1 (:129-133,657-660)
                else if (ktype == 8) val = lds_u8(win, addr + v);
                else if (ktype == 12) val = lds_u32(win, addr + 2u * v) & 0xffffu;
                else if (ktype == 14) val = lds_u32(win, addr + 4u * v);
                else {
                    const uint32_t l = v & 3u, row = v >> 2;
                    const uint32_t mask = (1u << w) - 1u;
                    const uint32_t lo = lds_u32(win, addr + 4u * l);
                    if (ktype == 7 || ktype == 9 || ktype == 11 || ktype == 13) {
                        const uint32_t hi = lds_u32(win, addr + 16u + 4u * l);
                        const uint32_t rs = 32u / w;                                 // the straddling row
                        const uint32_t resume = ktype == 7 ? 3u : ktype == 9 ? 4u : ktype == 11 ? 8u : 11u;
                        if (row < rs) val = (lo >> (row * w)) & mask;
                        else if (row == rs) val = ((lo >> (rs * w)) | (hi << (32u - rs * w))) & mask;
                        else val = (hi >> (resume + w * (row - rs - 1u))) & mask;
                    } else {
                        val = (lo >> (row * w)) & mask;
                    }
                }
                out[i] = val;
            }
        }
        pcarry += __shfl_sync(FULL, pincl, 31);
        icarry = round_end;
        if (nvalid < 32u) break;
    }
    __syncwarp();
    return (p0 - off) + len;
}

// ---- Binary interpolative block, n <= 128 (interpolative_coding.hpp:93-146) -------------------
// Bit-serial: each code length depends on previously decoded values, so lane 0 decodes while the
// warp waits.  Writes the PREFIX SUMS P[0..n-1] (P[n-1] = sum) into out; callers turn them into
// docids (base + P[i] + i) or freqs (P[i] - P[i-1]) in parallel.
__device__ __noinline__ uint32_t decode_interpolative_prefix(uint32_t win_off, uint32_t off, uint32_t n,
                                                             uint32_t sum_of_values, uint32_t out_off, uint32_t scratch_off) {
    const uint32_t* win = smem_words(win_off);
    uint32_t* out = smem_words(out_off);
    uint32_t* scratch = smem_words(scratch_off);
    uint32_t consumed = 0;
    if (lane_id() == 0) {
        uint32_t pos = off;
        uint32_t sum = sum_of_values;
        if (sum == 0xffffffffu) sum = vbyte_decode(win, pos);
        out[n - 1] = sum;
        uint32_t bits = 0;
        if (n > 1) {
            const uint32_t bitbase = 8u * pos;
            uint32_t* stack = scratch;   // pending right halves: (base, cnt, low, high)
            int sp = 0;
            uint32_t base = 0, cnt = n - 1, low = 0, high = sum;
            while (true) {
                uint32_t h = cnt >> 1;
                uint32_t u = high - low + 1u;
                uint32_t val = 0;
                if (u > 1u) {
                    uint32_t nb = 31u - __clz(u);                               // msb(u)
                    uint32_t m = (nb == 31u) ? (0u - u) : ((2u << nb) - u);      // 2^(nb+1) - u
                    val = lds_bits(win, bitbase + bits, nb);
                    bits += nb;
                    if (val >= m) {
                        uint32_t one = lds_bits(win, bitbase + bits, 1);
                        bits += 1;
                        val = (val << 1) + one - m;
                    }
                }
                val += low;
                out[base + h] = val;
                uint32_t rc = cnt - h - 1u;
                if (h) {
                    if (rc) { stack[4 * sp] = base + h + 1u; stack[4 * sp + 1] = rc; stack[4 * sp + 2] = val; stack[4 * sp + 3] = high; ++sp; }
                    cnt = h; high = val;                  // descend left: (base, h, low, val)
                } else if (rc) {
                    base = base + 1u; cnt = rc; low = val; // h == 0: go right directly
                } else {
                    if (!sp) break;
                    --sp;
                    base = stack[4 * sp]; cnt = stack[4 * sp + 1]; low = stack[4 * sp + 2]; high = stack[4 * sp + 3];
                }
            }
        }
        consumed = (pos - off) + ((bits + 7u) >> 3);
    }
    __syncwarp();
    return __shfl_sync(FULL, consumed, 0);
}

// ---- Binary interpolative block, ONE LANE PER BLOCK --------------------------------------------
// The code is bit-serial inside a block, but blocks are independent: in the batched decode each lane
// of a warp takes its own block (list tails in every block index, every block of block_interpolative)
// straight from global memory.  P (prefix sums) goes to the lane's column of a shared-memory buffer in tree order;
// the warp then writes every block out together: docids = base + P[i] + i, freqs = P[i] - P[i-1] + 1.
struct GlobalBitReader {
    const uint32_t* word;   // next aligned word to load
    uint64_t buf;
    uint32_t avail;
    uint32_t consumed_bits;
    __device__ __forceinline__ void init(const uint8_t* p) {
        uintptr_t a = reinterpret_cast<uintptr_t>(p);
        word = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
        uint32_t skip = uint32_t(a & 3u) * 8u;
        buf = uint64_t(__ldg(word++)) >> skip;
        avail = 32u - skip;
        consumed_bits = 0;
    }
    __device__ __forceinline__ uint32_t read(uint32_t len) {
        if (!len) return 0u;
        if (avail < len) { buf |= uint64_t(__ldg(word++)) << avail; avail += 32u; }
        uint32_t v = uint32_t(buf & ((uint64_t(1) << len) - 1));
        buf >>= len; avail -= len; consumed_bits += len;
        return v;
    }
};

// P[0..n) of one block -> col[i * SERIAL_STRIDE] (the lane's column of a warp-shared, padded buffer).  The traversal
// (block_codecs.hpp:101-148 / interpolative_coding.hpp:113-153) is the reference's recursion made iterative; the
// bounds of a pending range are re-read from the values already decoded (low = P[base-1], high = P[base+cnt]), so
// the stack holds only (base, cnt) pairs and lives in a 128-bit shift register.  Returns the bytes consumed.
constexpr uint32_t SERIAL_STRIDE = 33;      // odd stride: the cooperative read-out of one column is conflict-free

__device__ __forceinline__ uint32_t decode_interpolative_lane(const uint8_t* in, uint32_t n, uint32_t sum_of_values, uint32_t* col) {
    const uint8_t* p = in;
    uint32_t sum = sum_of_values;
    if (sum == 0xffffffffu) {                    // TightVariableByte prefix (block_codecs.hpp:131-134)
        sum = 0;
        for (uint32_t shift = 0; shift <= 28; shift += 7) {
            uint32_t c = __ldg(p++);
            sum += (c & 127u) << shift;
            if (c & 128u) break;
        }
    }
    col[(n - 1) * SERIAL_STRIDE] = sum;
    uint32_t bits = 0;
    if (n > 1) {
        GlobalBitReader br;
        br.init(p);
        uint64_t stk_lo = 0, stk_hi = 0;         // pending right halves, 16 bits each: base | cnt << 8 (depth <= 7)
        uint32_t sp = 0;
        uint32_t base = 0, cnt = n - 1, low = 0, high = sum;
        while (true) {
            const uint32_t h = cnt >> 1;
            const uint32_t u = high - low + 1u;
            uint32_t val = 0;
            if (u > 1u) {
                const uint32_t nb = 31u - __clz(u);
                const uint32_t m = (nb == 31u) ? (0u - u) : ((2u << nb) - u);
                val = br.read(nb);
                if (val >= m) val = (val << 1) + br.read(1) - m;
            }
            val += low;
            col[(base + h) * SERIAL_STRIDE] = val;
            const uint32_t rc = cnt - h - 1u;
            if (h) {
                if (rc) {
                    stk_hi = (stk_hi << 16) | (stk_lo >> 48);
                    stk_lo = (stk_lo << 16) | uint64_t((base + h + 1u) | (rc << 8));
                    ++sp;
                }
                cnt = h; high = val;
            } else if (rc) {
                base = base + 1u; cnt = rc; low = val;
            } else {
                if (!sp) break;
                --sp;
                const uint32_t e = uint32_t(stk_lo) & 0xffffu;
                stk_lo = (stk_lo >> 16) | (stk_hi << 48);
                stk_hi >>= 16;
                base = e & 0xffu; cnt = e >> 8;
                low = col[(base - 1u) * SERIAL_STRIDE];          // base >= 1: a right half starts behind its parent
                high = col[(base + cnt) * SERIAL_STRIDE];
            }
        }
        bits = br.consumed_bits;
    }
    return uint32_t(p - in) + ((bits + 7u) >> 3);
}

}  // namespace ds2i_gpu
