// Device side of ds2i's partitioned Elias-Fano index (`opt`): freq_index<partitioned_sequence<>,
// positive_sequence<partitioned_sequence<strict_sequence>>> (index_types.hpp:30-33).
//
// The two big bit vectors (docs, freqs) are copied into HBM as they are on disk (64-bit words,
// LSB-first).  What the reference re-derives on every partition switch — partition sizes,
// upper bounds and endpoints, each an Elias-Fano / fixed-width sequence in the list header
// (partitioned_sequence.hpp:299-326) — is decoded ONCE at load time into a flat partition directory
// (PefPart), the analogue of block_maxs/block_endpoints of the block indexes, so that a warp finds a
// partition with the same ballot search it uses for block_max.  Partition bodies stay compressed and
// are decoded on the fly, 128 consecutive elements at a time ("chunk"):
//   * compact_elias_fano body (compact_elias_fano.hpp:14-61,105-126): the set bits of the high-bits
//     region are located 32 words per step (popcount + warp scan + select-in-word), starting from the
//     pointers1 sample of the chunk; low bits come from one unaligned field read per element;
//   * compact_ranked_bitvector body (compact_ranked_bitvector.hpp:14-50): same scan over the bitmap;
//   * all_ones (all_ones_sequence.hpp): arithmetic.
// next_geq inside a large partition uses pointers0 (EF, :291-336) / rank1 samples (bitvector).
#pragma once
#include "device_common.cuh"

namespace ds2i_gpu {

struct PefPart {            // one partition of one sequence (32 bytes: two 16-byte loads)
    uint64_t bit_off;       // absolute bit offset of the partition body (its type bit, when present)
    uint32_t begin;         // position of its first element inside the list
    uint32_t size;          // elements
    uint32_t base;          // value offset: stored values are relative to it
    uint32_t ub;            // its last value (absolute)
    uint32_t first_block;   // docs sequences: number of 128-element windows ("blocks") of the list before this partition
    uint32_t body_bits;     // bits of the body (type bit included): what a staged copy has to cover
};
static_assert(sizeof(PefPart) == 32, "PefPart layout");

// Where a decoder reads the bit vector from: HBM directly, or a window of it that TMA staged into shared memory
// (the words of the window sit at smem[0 .. nw), word w0 of the vector first; reads outside the window give 0 — the
// 32-words-per-step scans look past the end of a body, what they find there is never selected).
struct GlobalBits {
    const uint64_t* p;
    __device__ __forceinline__ uint64_t word(uint64_t w) const { return __ldg(p + w); }
};
struct StagedBits {
    uint32_t smem_off;      // byte offset of the window inside the kernel's dynamic shared memory (g_smem)
    uint32_t nw;
    uint64_t w0;
    __device__ __forceinline__ uint64_t word(uint64_t w) const {
        const uint64_t i = w - w0;
        return i < nw ? reinterpret_cast<const uint64_t*>(g_smem + smem_off)[i] : 0ull;
    }
};

struct PefListDir {
    uint64_t first_part;    // index into the PefPart array
    uint32_t nparts;
    uint32_t n;             // elements of the sequence
};

struct PefSeq {             // one bitvector_collection on device
    const uint64_t* bits;   // m_bitvectors words, 8-byte aligned copy with a zero tail
    const PefListDir* lists;
    const PefPart* parts;
    uint32_t log_sampling0, log_sampling1, rb_log_rank1_sampling, rb_log_sampling1;
    uint32_t raw_ef;        // ef_index: bodies are plain compact_elias_fano / strict_elias_fano (no type bit, zeros sampled)
};

struct PefIndexDev {
    PefSeq docs, freqs;
    uint64_t num_lists;
    uint32_t num_docs;
};

enum : uint32_t { PEF_EF = 0, PEF_RB = 1, PEF_AO = 2 };

__device__ __forceinline__ uint32_t ceil_log2_dev(uint64_t x) { return x > 1 ? 64u - uint32_t(__clzll(x - 1)) : 0u; }

// bits [pos, pos+len) of the vector, len <= 32
template <class Bits>
__device__ __forceinline__ uint32_t bv_get_bits(Bits const& bits, uint64_t pos, uint32_t len) {
    if (!len) return 0u;
    uint64_t w = bits.word(pos >> 6);
    uint32_t sh = uint32_t(pos & 63);
    uint64_t v = w >> sh;
    if (sh + len > 64) v |= bits.word((pos >> 6) + 1) << (64 - sh);
    return uint32_t(v) & (len >= 32 ? 0xffffffffu : ((1u << len) - 1u));
}
template <class Bits>
__device__ __forceinline__ uint64_t bv_get_bits64(Bits const& bits, uint64_t pos, uint32_t len) {   // len <= 57
    if (!len) return 0ull;
    uint64_t w = bits.word(pos >> 6);
    uint32_t sh = uint32_t(pos & 63);
    uint64_t v = w >> sh;
    if (sh + len > 64) v |= bits.word((pos >> 6) + 1) << (64 - sh);
    return v & ((uint64_t(1) << len) - 1);
}

// k-th (0-based) set bit of a 64-bit word: popcount bisection (the __fns intrinsic is a software loop — it was a quarter of
// all instructions of the window decoder)
__device__ __forceinline__ uint32_t select_in_word(uint64_t w, uint32_t k) {
    uint32_t x = uint32_t(w), r = 0;
    const uint32_t pl = __popc(x);
    if (k >= pl) { k -= pl; x = uint32_t(w >> 32); r = 32; }
#pragma unroll
    for (uint32_t s = 16; s >= 1; s >>= 1) {
        const uint32_t c = __popc(x & ((1u << s) - 1u));
        if (k >= c) { k -= c; x >>= s; r += s; }
    }
    return r;
}

// either of the two, chosen at run time (warp-uniformly): ONE instance of the de-inlined window decoder serves both
struct AnyBits {
    const uint64_t* p;
    uint64_t w0;
    uint32_t smem_off;
    uint32_t nw;            // 0: read HBM
    __device__ __forceinline__ uint64_t word(uint64_t w) const {
        if (nw) { const uint64_t i = w - w0; return i < nw ? reinterpret_cast<const uint64_t*>(g_smem + smem_off)[i] : 0ull; }
        return __ldg(p + w);
    }
};

// Layout of one partition body, resolved from (universe, n) like the reference constructors do.
struct PefBody {
    uint32_t type;
    uint32_t n, universe;
    uint32_t lower_bits;        // EF
    uint32_t pointer_size;
    uint64_t pointers0_off, pointers1_off, high_off, low_off;       // EF (absolute bit offsets)
    uint64_t rank_off, rb_ptr1_off, bitmap_off;                     // RB
    uint32_t rank_sample_size, n_rank_samples;
    uint32_t log_s0, log_s1;
    bool strict;                // strict_elias_fano: stored value = v - i over universe - n + 1
};

// indexed_sequence / strict_sequence dispatch (indexed_sequence.hpp:95-127, strict_sequence.hpp:104-137)
// known_type: the partition's type bit when the caller has it cached (0xffffffff: read it from the body)
template <class Seq, class Bits>
__device__ __forceinline__ PefBody pef_open_body(Seq const& seq, Bits const& bits, PefPart const& p, bool strict, uint32_t known_type = 0xffffffffu) {
    PefBody b;
    b.n = p.size;
    b.universe = p.ub - p.base + 1u;            // last relative value + 1
    b.strict = strict;
    uint64_t off = p.bit_off;
    if (seq.raw_ef) b.type = PEF_EF;            // compact_elias_fano.hpp:63-136 / strict_elias_fano.hpp:20-36 written directly
    else {
        if (b.universe == b.n) { b.type = PEF_AO; return b; }
        b.type = known_type != 0xffffffffu ? known_type : uint32_t(bits.word(p.bit_off >> 6) >> (p.bit_off & 63)) & 1u;
        off += 1;
    }
    if (b.type == PEF_EF) {
        // strict_sequence never indexes zeros (strict_sequence.hpp:24-30: ef_log_sampling0 = 63); strict_elias_fano
        // used on its own (ef_index) keeps the global sampling
        b.log_s0 = (strict && !seq.raw_ef) ? 63u : seq.log_sampling0;
        b.log_s1 = seq.log_sampling1;
        uint64_t u = strict ? uint64_t(b.universe) - b.n + 1 : uint64_t(b.universe);
        uint64_t n = b.n;
        // msb(u / n): a 32-bit division whenever both fit (the 64-bit one is an ~80-instruction routine)
        b.lower_bits = u > n ? ((u >> 32) ? 63u - uint32_t(__clzll(u / n)) : 31u - uint32_t(__clz(uint32_t(u) / uint32_t(n)))) : 0u;
        uint64_t hbl = n + (u >> b.lower_bits) + 2;
        b.pointer_size = ceil_log2_dev(hbl);
        uint64_t p0 = b.log_s0 >= 63 ? 0 : ((hbl - n) >> b.log_s0);
        uint64_t p1 = n >> b.log_s1;
        b.pointers0_off = off;
        b.pointers1_off = off + p0 * b.pointer_size;
        b.high_off = b.pointers1_off + p1 * b.pointer_size;
        b.low_off = b.high_off + hbl;
    } else {
        b.log_s0 = strict ? 63u : seq.rb_log_rank1_sampling;
        b.log_s1 = seq.rb_log_sampling1;
        b.rank_sample_size = ceil_log2_dev(uint64_t(b.n) + 1);
        b.pointer_size = ceil_log2_dev(b.universe);
        b.n_rank_samples = b.log_s0 >= 63 ? 0u : (b.universe >> b.log_s0);
        uint64_t p1 = b.n >> b.log_s1;
        b.rank_off = off;
        b.rb_ptr1_off = off + uint64_t(b.n_rank_samples) * b.rank_sample_size;
        b.bitmap_off = b.rb_ptr1_off + p1 * b.pointer_size;
    }
    return b;
}

// Positions (relative to `origin`) of the set bits with ordinals r0 .. r0+cnt-1 (ordinal 0 = first
// set bit at or after `start`), cnt <= 129: the first 128 go to out[0..), the 129th (the freqs path decodes one element
// before its window) is returned, warp-uniformly.  32 words per step.
template <class Bits>
__device__ __forceinline__ uint32_t pef_scan_ones(Bits const& bits, uint64_t start, uint64_t origin, uint32_t r0, uint32_t cnt, uint32_t* out) {
    const unsigned lane = lane_id();
    uint64_t wbase = start >> 6;
    uint32_t seen = 0;                    // set bits before the words of this step
    const uint32_t r1 = r0 + cnt;
    uint32_t extra = 0;
    bool first = true;
    while (seen < r1) {
        uint64_t w = bits.word(wbase + lane);
        if (first && lane == 0) w &= ~uint64_t(0) << (start & 63);
        first = false;
        uint32_t pc = __popcll(w);
        uint32_t incl = warp_inclusive_scan(pc);
        uint32_t excl = seen + incl - pc;
        uint32_t total = seen + __shfl_sync(FULL, incl, 31);
        // the word holding ordinal o: largest t with excl_t <= o.  (Words with no set bits share excl with their successor;
        // the search prefers the later one, which is the one that owns the ordinal.)
        auto locate = [&](uint32_t o, bool want, uint32_t& pos) {
            uint32_t t = 0;
#pragma unroll
            for (uint32_t s = 16; s >= 1; s >>= 1) {
                uint32_t e = __shfl_sync(FULL, excl, (t + s) & 31u);
                if (e <= o) t += s;
            }
            uint32_t we = __shfl_sync(FULL, excl, t);
            uint32_t wlo = __shfl_sync(FULL, uint32_t(w), t), whi = __shfl_sync(FULL, uint32_t(w >> 32), t);
            if (want) pos = uint32_t((wbase + t) * 64 + select_in_word((uint64_t(whi) << 32) | wlo, o - we) - origin);
            return want;
        };
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t o = r0 + lane + 32u * j;            // ordinal wanted
            uint32_t pos = 0;
            if (locate(o, (lane + 32u * j) < cnt && o >= seen && o < total, pos)) out[lane + 32u * j] = pos;
        }
        if (cnt > 128u) {                                       // warp-uniform
            const uint32_t o = r0 + 128u;
            locate(o, o >= seen && o < total, extra);
        }
        seen = total;
        wbase += 32;
    }
    __syncwarp();
    return extra;
}

// The same with the words of a step parked in shared memory (scratch_off: 32 x u64 words + 32 x u32 first ordinals = 384 B per
// warp): every lane resolves FOUR CONSECUTIVE ordinals — one search for the word of the first, one select to drop the bits
// before it, then next-set-bit steps — instead of four independent search + select rounds.  Output index = ordinal - r0 as above.
template <class Bits>
__device__ __forceinline__ uint32_t pef_scan_ones_smem(Bits const& bits, uint64_t start, uint64_t origin, uint32_t r0, uint32_t cnt, uint32_t* out, uint32_t scratch_off) {
    const unsigned lane = lane_id();
    uint64_t* wbuf = reinterpret_cast<uint64_t*>(g_smem + scratch_off);
    uint32_t* ebuf = reinterpret_cast<uint32_t*>(wbuf + 32);
    uint64_t wbase = start >> 6;
    uint32_t seen = 0, extra = 0;
    const uint32_t r1 = r0 + cnt;
    const uint32_t i0 = 4u * lane, i1 = min(cnt, i0 + 4u + ((lane == 31u && cnt > 128u) ? 1u : 0u));     // my output indices [i0, i1)
    bool first = true;
    while (seen < r1) {
        uint64_t w = bits.word(wbase + lane);
        if (first && lane == 0) w &= ~uint64_t(0) << (start & 63);
        first = false;
        const uint32_t pc = __popcll(w);
        const uint32_t incl = warp_inclusive_scan(pc);
        const uint32_t excl = seen + incl - pc;
        const uint32_t total = seen + __shfl_sync(FULL, incl, 31);
        __syncwarp();
        wbuf[lane] = w; ebuf[lane] = excl;
        __syncwarp();
        uint32_t o = max(r0 + i0, seen);
        const uint32_t hi = min(r0 + i1, total);
        if (i0 < i1 && o < hi) {
            // word holding ordinal o: the largest t with ebuf[t] <= o (words without set bits share their first ordinal with
            // their successor; the search lands on the last of them, the one that owns the ordinal)
            uint32_t t = 0;
#pragma unroll
            for (uint32_t s = 16; s >= 1; s >>= 1)
                if (ebuf[t + s] <= o) t += s;
            uint64_t cur = wbuf[t];
            const uint32_t k = o - ebuf[t];
            if (k) cur &= ~uint64_t(0) << select_in_word(cur, k);            // drop the k set bits before it
            for (; o < hi; ++o) {
                while (!cur) cur = wbuf[++t];
                const uint32_t p = uint32_t((wbase + t) * 64 + uint32_t(__ffsll((long long)cur) - 1) - origin);
                cur &= cur - 1;
                if (o - r0 < 128u) out[o - r0] = p; else extra = p;
            }
        }
        seen = total;
        wbase += 32;
    }
    __syncwarp();
    return cnt > 128u ? __shfl_sync(FULL, extra, 31) : 0u;
}

// count of ZERO bits wanted: position (relative to origin) of the zero with ordinal z (0 = first zero
// at or after start), z < 2^log_sampling0 + slack.  Returns warp-uniformly.
template <class Bits>
__device__ __forceinline__ uint64_t pef_select_zero(Bits const& bits, uint64_t start, uint32_t z) {
    const unsigned lane = lane_id();
    uint64_t wbase = start >> 6;
    uint32_t seen = 0;
    bool first = true;
    while (true) {
        uint64_t w = ~bits.word(wbase + lane);
        if (first && lane == 0) w &= ~uint64_t(0) << (start & 63);
        first = false;
        uint32_t pc = __popcll(w);
        uint32_t incl = warp_inclusive_scan(pc);
        uint32_t excl = seen + incl - pc;
        unsigned hit = __ballot_sync(FULL, z >= excl && z < excl + pc);
        if (hit) {
            uint32_t t = __ffs(hit) - 1;
            uint32_t e = __shfl_sync(FULL, excl, t);
            uint32_t wlo = __shfl_sync(FULL, uint32_t(w), t), whi = __shfl_sync(FULL, uint32_t(w >> 32), t);
            uint32_t bit = select_in_word((uint64_t(whi) << 32) | wlo, z - e);
            return (wbase + t) * 64 + bit;
        }
        seen += __shfl_sync(FULL, incl, 31);
        wbase += 32;
    }
}

// Elements [i0, i0+cnt) of a partition (cnt <= 128) as ABSOLUTE values (base added) into out[0..cnt).  with_prev (i0 > 0):
// element i0 - 1 is decoded along (one scan instead of two) and returned, warp-uniformly; otherwise 0 is returned.
template <class Bits>
__device__ __forceinline__ uint32_t pef_decode_range(Bits const& bits, PefPart const& p, PefBody const& b, uint32_t i0, uint32_t cnt, uint32_t* out,
                                                     bool with_prev = false, uint32_t scratch_off = 0 /* 384 B of shared memory: the faster scan */) {
    const unsigned lane = lane_id();
    if (b.type == PEF_AO) {
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            uint32_t e = lane + 32u * j;
            if (e < cnt) out[e] = p.base + i0 + e;
        }
        __syncwarp();
        return with_prev ? p.base + i0 - 1u : 0u;
    }
    const uint32_t shift = with_prev ? 1u : 0u;
    const uint32_t f0 = i0 - shift, total = cnt + shift;   // decoded elements: [f0, f0 + total), total <= 129
    const uint32_t s = f0 >> b.log_s1;                     // pointers1 sample at or before f0
    const bool ef = b.type == PEF_EF;
    uint64_t origin = ef ? b.high_off : b.bitmap_off;
    uint64_t start = origin;
    uint32_t r0 = f0;
    if (s) {
        // EF: high-bit position of element s << log_s1; bitvector: its value = its bit position
        uint64_t ptr = bv_get_bits64(bits, (ef ? b.pointers1_off : b.rb_ptr1_off) + uint64_t(s - 1) * b.pointer_size, b.pointer_size);
        start = origin + ptr;
        r0 = f0 - (s << b.log_s1);
    }
    const uint32_t extra = scratch_off ? pef_scan_ones_smem(bits, start, origin, r0, total, out, scratch_off) : pef_scan_ones(bits, start, origin, r0, total, out);
    // positions -> values; with_prev moves every value one slot down (element f0 goes to the return register), so all
    // positions are read before any value is written
    uint32_t pos[4];
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) { const uint32_t e = lane + 32u * j; pos[j] = e < total ? out[e] : 0u; }
    if (with_prev) __syncwarp();
    auto value = [&](uint32_t e, uint32_t ps) {
        if (!ef) return p.base + ps;
        const uint32_t i = f0 + e;
        const uint32_t high = ps - i - 1u;
        const uint32_t low = bv_get_bits(bits, b.low_off + uint64_t(i) * b.lower_bits, b.lower_bits);
        uint32_t v = (high << b.lower_bits) | low;
        if (b.strict) v += i;                              // strict_elias_fano.hpp:50-60
        return p.base + v;
    };
    uint32_t prev = 0;
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
        const uint32_t e = lane + 32u * j;
        if (e < total && e < 128u) {
            const uint32_t v = value(e, pos[j]);
            if (e >= shift) out[e - shift] = v; else prev = v;
        }
    }
    if (total > 128u && lane == 0) out[127] = value(128u, extra);
    __syncwarp();
    return with_prev ? __shfl_sync(FULL, prev, 0) : 0u;
}

// Index of the first element whose high part is >= (x >> l) (EF) / whose value is >= x (bitvector,
// all-ones): a lower bound for the rank of x; for EF the caller refines by comparing decoded values.
// x is relative to the partition base, x < universe.
template <class Bits>
__device__ __forceinline__ uint32_t pef_rank_hint(Bits const& bits, PefBody const& b, uint32_t x) {
    if (b.type == PEF_AO) return x;
    if (b.type == PEF_EF) {
        uint32_t high = x >> b.lower_bits;
        if (high == 0) return 0;
        // position of the zero that has `high` zeros before it (compact_elias_fano.hpp:291-336)
        uint32_t k = b.log_s0 >= 63 ? 0u : (high >> b.log_s0);
        uint64_t start = b.high_off;
        uint32_t skip = high;
        if (k) {
            uint64_t ptr = bv_get_bits64(bits, b.pointers0_off + uint64_t(k - 1) * b.pointer_size, b.pointer_size);
            start = b.high_off + ptr;
            skip = high - (k << b.log_s0);
        }
        uint64_t z = pef_select_zero(bits, start, skip);
        return uint32_t(z - b.high_off) - high;
    }
    // ranked bitvector: ones in [0, x) = sample + popcount of the remainder (compact_ranked_bitvector.hpp:256-302)
    const unsigned lane = lane_id();
    uint32_t k = b.log_s0 >= 63 ? 0u : (x >> b.log_s0);
    if (k > b.n_rank_samples) k = b.n_rank_samples;
    uint32_t rank = 0;
    uint64_t from = b.bitmap_off;
    if (k) {
        rank = uint32_t(bv_get_bits64(bits, b.rank_off + uint64_t(k - 1) * b.rank_sample_size, b.rank_sample_size));
        from = b.bitmap_off + (uint64_t(k) << b.log_s0);
    }
    const uint64_t to = b.bitmap_off + x;
    uint32_t acc = 0;
    for (uint64_t wb = from >> 6; wb * 64 < to; wb += 32) {
        uint64_t wi = wb + lane;
        uint64_t w = 0;
        if (wi * 64 < to) {
            w = bits.word(wi);
            if (wi == (from >> 6)) w &= ~uint64_t(0) << (from & 63);
            if ((wi + 1) * 64 > to) w &= (uint64_t(1) << (to & 63)) - 1;
        }
        acc += __popcll(w);
    }
    return rank + __reduce_add_sync(FULL, acc);
}

// what pef_open_body needs of a PefSeq, by value (argument of the de-inlined window decoder)
struct PefSeqParams {
    uint32_t log_sampling0, log_sampling1, rb_log_rank1_sampling, rb_log_sampling1, raw_ef;
};
__device__ __forceinline__ PefSeqParams pef_params(PefSeq const& s) {
    return PefSeqParams{s.log_sampling0, s.log_sampling1, s.rb_log_rank1_sampling, s.rb_log_sampling1, s.raw_ef};
}

// Elements [i0, i0 + cnt) of partition p (cnt <= 128) as absolute values into the shared-memory words at out_off.  De-inlined:
// the block-parallel query kernels decode windows at several call sites (driving list, probed lists, docs and freqs), and one
// inlined copy per site (~2.5 k instructions each) made them fetch-bound — 34 issue slots lost to instruction misses per issue.
// with_prev: also returns element i0 - 1 (i0 > 0).
__device__ __noinline__ uint32_t pef_window_values(PefSeqParams seq, AnyBits bits, PefPart p, bool strict, uint32_t i0, uint32_t cnt, uint32_t out_off,
                                                   bool with_prev, uint32_t scratch_off, uint32_t known_type = 0xffffffffu) {
    const PefBody b = pef_open_body(seq, bits, p, strict, known_type);
    return pef_decode_range(bits, p, b, i0, cnt, smem_words(out_off), with_prev, scratch_off);
}

// the same straight from HBM (what the literal enumerator and the full-decode kernels use)
__device__ __forceinline__ PefBody pef_open_body(PefSeq const& seq, PefPart const& p, bool strict) { return pef_open_body(seq, GlobalBits{seq.bits}, p, strict); }
__device__ __forceinline__ void pef_decode_range(PefSeq const& seq, PefPart const& p, PefBody const& b, uint32_t i0, uint32_t cnt, uint32_t* out) {
    pef_decode_range(GlobalBits{seq.bits}, p, b, i0, cnt, out, false);
}
__device__ __forceinline__ uint32_t pef_rank_hint(PefSeq const& seq, PefBody const& b, uint32_t x) { return pef_rank_hint(GlobalBits{seq.bits}, b, x); }

// one PefPart record: two 16-byte loads
__device__ __forceinline__ PefPart pef_load_part(const PefPart* parts, uint64_t i) {
    const uint4* q = reinterpret_cast<const uint4*>(parts + i);
    const uint4 a = __ldg(q), b = __ldg(q + 1);
    PefPart r;
    r.bit_off = (uint64_t(a.y) << 32) | a.x; r.begin = a.z; r.size = a.w;
    r.base = b.x; r.ub = b.y; r.first_block = b.z; r.body_bits = b.w;
    return r;
}

}  // namespace ds2i_gpu
