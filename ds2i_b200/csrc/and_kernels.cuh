// Block-at-a-time conjunctive evaluation: and_query / ranked_and_query (queries.hpp:35-86,322-401)
// re-thought for a warp.  The reference advances ONE candidate docid at a time through next_geq;
// here the 128 docids of a block of the shortest list are the candidates, carried 4 per lane, and
// every other list is probed for all of them at once (block_max search by ballot over the aligned
// block directory, 128-wide binary search of the decoded block in shared memory).  BM25 is accumulated
// per candidate in registers as the lists are probed, in the reference's order (lists by increasing
// size), so every score is bit-identical; only scores that beat the running threshold reach the
// (serial) top-k insert.
//
// Work item = (query, chunk of CH consecutive blocks of its shortest list): heavy queries are spread
// over many warps, each with a private top-k; merge_items_kernel folds the partial results.  The
// top-k multiset and the match count do not depend on the evaluation order, so results equal the
// reference's exactly.
//
// Shared memory per warp is kept small (560 B per query term + one staging window + one freqs
// buffer) and the kernel is instantiated per codec, because the path is instruction-issue bound:
// resident warps and instruction count are what the throughput follows.
#pragma once
#include <type_traits>

#include "query_kernels.cuh"

namespace ds2i_gpu {

// algorithmic-work counters
// Every use site has the warp context in a variable `c`; AndCtxT<false> compiles the counters away (the timed launches of the
// benchmark run without them — they cost registers in kernels that are register-bound — and one more, instrumented launch of
// the same batch collects them: the algorithmic work of a batch does not depend on the instance that measures it).
#define DS2I_STAT(x) if constexpr (std::remove_reference_t<decltype(c)>::stats) { x }

#ifdef DS2I_NIN_HIST
// experiment: distribution of the probe regimes.  [0..7] probe steps by candidates answered (1, 2-3, 4-7, .., 64-128), [8] driver blocks,
// [9..13] (driver block, list) visits by pending candidates at entry (1-8, 9-32, 33-64, 65-128), [16..23] visits by steps taken (1,2-3,4-7,...)
__device__ unsigned long long g_hist[32];
__device__ __forceinline__ uint32_t hist_bucket(uint32_t n) { return n ? 31u - __clz(n) : 0u; }
#define DS2I_HIST(i, v) do { if (lane_id() == 0) atomicAdd(&g_hist[i], (unsigned long long)(v)); } while (0)
#else
#define DS2I_HIST(i, v)
#endif

constexpr uint32_t AND_CHUNK_BLOCKS = 32;     // blocks of the shortest list per work item

// Double-buffered TMA for the driving list (its next block pair is copied into a second staging window while the current
// block is probed).  Built and measured on B200 in round 2 (-DDS2I_DRIVER_PREFETCH=1): results identical, ranked_and 16.2 ->
// 18.3 ms, and 9.0 -> 9.1 ms per 10k-query batch — the kernels are issue- and register-bound, not latency-bound (the copy's
// latency was already covered by the other resident warps), and the extra state costs spills.  Off by default.
#ifndef DS2I_DRIVER_PREFETCH
#define DS2I_DRIVER_PREFETCH 0
#endif
constexpr bool AND_DRIVER_PREFETCH = DS2I_DRIVER_PREFETCH != 0;
constexpr uint32_t AND_STAGE_WINDOWS = AND_DRIVER_PREFETCH ? 2 : 1;

// Work items are implicit: query sched[p] owns ceil(blocks of its shortest list / chunk_blocks) consecutive items; a warp
// maps a global item number to its query with a 32-ary search of the prefix array gstart (in processing order).
struct AndJob {
    const uint32_t* gstart;      // nq+1: items before the p-th query of the processing order
    const uint32_t* item_begin;  // nq+1: first result slot of query q (results are laid out in query order)
    uint32_t nitems;
    const uint8_t* qchunk;     // nq: blocks of the shortest list per item of query q (1..32: fewer where a block is expensive to probe)
    uint32_t* work_counter;
    uint32_t* item_counts;     // nitems: matches found by the item
    uint32_t* item_sizes;      // nitems: entries in the item's partial top-k
    float* item_scores;        // nitems * 2k: k scores, then the k docids they belong to
};

// last position p in [0, n) with a[p] <= x (a is non-decreasing, a[0] <= x): 32 probes per step
__device__ __forceinline__ uint32_t warp_upper_group(const uint32_t* a, uint32_t n, uint32_t x) {
    const unsigned lane = lane_id();
    uint32_t lo = 0, hi = n;            // answer in [lo, hi)
    while (hi - lo > 1) {
        const uint32_t span = hi - lo;
        const uint32_t step = (span + 31u) / 32u;
        const uint32_t p = lo + lane * step;
        const bool le = p < hi && __ldg(a + p) <= x;
        const unsigned m = __ballot_sync(FULL, le);          // lane 0 always set
        const uint32_t f = 31u - __clz(m);
        lo = lo + f * step;
        hi = min(hi, lo + step);
    }
    return lo;
}

// per query term: cursor + the decoded docids of the current block
struct AndList {
    uint64_t data_off;      // absolute byte offset of the list's block data inside m_lists (Elias-Fano: first partition of the docs sequence)
    uint32_t bfirst;        // the list's first entry in the block directory
    uint32_t nblocks;
    uint32_t n;
    uint32_t last_max;      // last docid of the list
    uint32_t term;          // the list (the Elias-Fano path looks its freqs sequence up by it)
    uint32_t pad1;
    // written together by lane 0 after every block decode (one 16-B store)
    uint32_t cur_block;     // 0xffffffff: not positioned yet
    uint32_t cur_max;       // last docid of the current block
    uint32_t cur_end;       // byte offset (from data_off) where the current block ends
    uint32_t freqs_off;     // byte offset (from data_off) of the current block's freqs
    uint32_t docs[BLOCK];   // absolute docids of the current block (0xffffffff beyond its size)
};
static_assert(sizeof(AndList) == 48 + 4 * BLOCK, "AndList layout");

struct AndWarp {
    float qw[MAX_TERMS];
    uint64_t bar;           // mbarrier of the probe staging window
    uint64_t dbar;          // mbarrier of the driving list's staging window (the next block pair is in flight while this one is probed)
};

// Elias-Fano path: the partition of the freqs sequence a list's last freqs window came from (lists are walked forward and
// freqs partitions are long, so the next window nearly always hits it: no directory search, no descriptor loads)
struct PefFreqSlot {
    PefPart part;
    uint32_t fp;            // index of the partition inside the list's freqs sequence; 0xffffffff: nothing cached
    uint32_t type;          // its type bit
    uint32_t pad[6];
};
static_assert(sizeof(PefFreqSlot) == 64, "PefFreqSlot layout");
constexpr uint32_t PEF_SCAN_SCRATCH_BYTES = 384;      // pef_scan_ones_smem: the words of a step and their first ordinals

__host__ __device__ constexpr size_t and_warp_smem_bytes(int slots, bool pef = false) {
    return sizeof(AndWarp) + size_t(slots) * sizeof(AndList) + BLOCK * 4 /* freqs */ + AND_STAGE_WINDOWS * STAGE_WORDS * 4 /* probe (+ driver) window */ +
           SCRATCH_WORDS * 4 + (pef ? size_t(slots) * sizeof(PefFreqSlot) + PEF_SCAN_SCRATCH_BYTES : 0);
}

// staging window <- bytes [start, end) of m_lists (the enclosing 16-B aligned range, one TMA bulk copy,
// completion on the warp's mbarrier); returns the offset of `start` inside the window
__device__ __forceinline__ uint32_t and_stage(const uint8_t* lists, uint64_t start, uint64_t end, uint32_t* stage, uint64_t* bar, uint32_t& phase) {
    const uint64_t a0 = start & ~uint64_t(15);
    uint32_t bytes = uint32_t(((end + 15) & ~uint64_t(15)) - a0);
    if (bytes > STAGE_BYTES) bytes = STAGE_BYTES;
    __syncwarp();   // every lane is done reading the previous window
    if (bytes) {
        if (lane_id() == 0) {
            mbar_expect_tx(bar, bytes);
            tma_load_1d(stage, lists + a0, bytes, bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
    }
    return uint32_t(start - a0);
}

template <int CODEC, bool WIDE_EXC = false>
__device__ __forceinline__ uint32_t and_decode_values(uint32_t stage_off, uint32_t off, uint32_t size, uint32_t sum_of_values, uint32_t out_off,
                                                      uint32_t stack_off, bool& prefix_out) {
    if (CODEC == CODEC_MIXED && size == BLOCK) {
        // mixed_block::decode (mixed_block.hpp:198-217): block_type byte, then that codec's block
        const uint32_t type = lds_u8(smem_words(stage_off), off);
        prefix_out = type == 2u;
        if (type == 1u) return 1u + decode_varint128(stage_off, off + 1u, out_off);
        if (type == 0u) return 1u + decode_optpfor128<WIDE_EXC>(stage_off, off + 1u, out_off, stack_off);
        return 1u + decode_interpolative_prefix(stage_off, off + 1u, size, sum_of_values, out_off, stack_off);
    }
    if (CODEC != CODEC_INTERPOLATIVE && CODEC != CODEC_MIXED && size == BLOCK) {
        prefix_out = false;
        if (CODEC == CODEC_OPTPFOR) return decode_optpfor128<WIDE_EXC>(stage_off, off, out_off, stack_off);
        if (CODEC == CODEC_VARINT) return decode_varint128(stage_off, off, out_off);
        return decode_qmx128(stage_off, off, out_off);
    }
    // n < block_size => every codec falls back to interpolative (block_codecs.hpp:196-199,215-217)
    prefix_out = true;
    return decode_interpolative_prefix(stage_off, off, size, sum_of_values, out_off, stack_off);
}

// cursor of list `term` into its slot (lane-private call: one lane per query term)
template <int CODEC>
__device__ __forceinline__ void and_list_setup(DevIndex const& idx, AndList* s, uint32_t term) {
    const ListDir d = idx.dir[term];
    const uint32_t bfirst = idx.bfirst[term];
    uint32_t nblocks;
    if (CODEC == CODEC_PEF) {
        nblocks = idx.bfirst[term + 1] - bfirst;                     // windows never straddle partitions: not ceil(n / 128)
        s->data_off = idx.pdocs.lists[term].first_part;
    } else {
        nblocks = (d.n + BLOCK - 1) / BLOCK;
        s->data_off = d.maxs_off + 4ull * nblocks + 4ull * (nblocks - 1);
    }
    s->bfirst = bfirst; s->nblocks = nblocks; s->n = d.n; s->term = term;
    s->last_max = __ldg(idx.bdir + bfirst + nblocks - 1).x;
    s->cur_block = 0xffffffffu;     // not positioned yet
    s->cur_max = 0; s->cur_end = 0; s->freqs_off = 0;
}

template <bool STATS>
struct AndCtxT {            // warp-uniform registers
    static constexpr bool stats = STATS;
    const uint8_t* lists;
    uint32_t* stage;
    uint64_t* bar;
    uint32_t stage_off, stack_off, ftmp_off;
    uint32_t fcache_off;    // Elias-Fano path: the warp's scan scratch (PEF_SCAN_SCRATCH_BYTES), then its PefFreqSlot array
    // double-buffered TMA: the driving list has its own staging window (right behind the probe window) and mbarrier (right
    // behind bar); block pair b + 1 is copied while block b is probed against the other lists
    uint32_t drv_slot;      // slot of the driving list
    uint32_t dwin_block;    // driver block whose pair is (or is being copied) in the driver window; 0xffffffff: none
    uint32_t dwin_delta;    // window offset of the list's data byte 0 (mod 2^32)
    uint32_t dphase;        // bit 0: phase of dbar; bit 1: the copy in flight has not been waited for yet
    uint32_t phase;
    uint32_t win_slot;      // list slot whose current block pair sits in the staging window (0xffffffff: none)
    uint32_t win_delta;     // window offset of that list's data byte 0 (mod 2^32)
    // algorithmic-work counters
    uint32_t c_docs_blocks, c_freqs_blocks, c_bytes_docs, c_bytes_freqs, c_maxs, c_scored;
};
typedef AndCtxT<true> AndCtx;

// ---- Elias-Fano index family: a "block" is a 128-element window of one partition of the docs sequence --------------------
// Window b of the list -> its partition (the directory's second word: index of the partition inside the list) and the
// elements [128 w, 128 w + 128) of it.  A body of up to STAGE_BYTES is copied into the warp's staging window with one TMA
// bulk copy and decoded from shared memory (both windows of a two-window partition reuse the copy); longer bodies (the
// single-partition sequences of very regular lists) are read in place.  De-inlined, scalars in and out: the kernels call it
// from two sites each and the code around the decoder is not small either (the kernels were instruction-fetch bound).
struct PefDocsWindow { uint32_t phase, staged_part, first_pos, cnt; };

__device__ __noinline__ PefDocsWindow pef_window_docs(PefSeq seq, uint64_t first_part, uint32_t b, uint32_t part_rel, uint32_t stage_off, uint32_t bar_off,
                                                      uint32_t phase, uint32_t staged_part /* partition in the staging window, 0xffffffff: none */, uint32_t docs_off,
                                                      uint32_t scratch_off) {
    const unsigned lane = lane_id();
    const PefPart p = pef_load_part(seq.parts, first_part + part_rel);
    const uint32_t i0 = (b - p.first_block) * BLOCK;
    const uint32_t cnt = min(BLOCK, p.size - i0);
    const uint64_t byte0 = (p.bit_off >> 3) & ~uint64_t(15);
    const uint64_t byte1 = (((p.bit_off + p.body_bits + 7) >> 3) + 15) & ~uint64_t(15);
    AnyBits bits{seq.bits, 0, 0, 0};
    if (p.body_bits && byte1 - byte0 <= STAGE_BYTES) {
        if (staged_part != part_rel) {
            uint64_t* bar = reinterpret_cast<uint64_t*>(g_smem + bar_off);
            __syncwarp();   // every lane is done reading the previous window
            if (lane == 0) {
                mbar_expect_tx(bar, uint32_t(byte1 - byte0));
                tma_load_1d(g_smem + stage_off, reinterpret_cast<const uint8_t*>(seq.bits) + byte0, uint32_t(byte1 - byte0), bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            staged_part = part_rel;
        }
        bits.w0 = byte0 >> 3; bits.smem_off = stage_off; bits.nw = uint32_t((byte1 - byte0) >> 3);
    }
    pef_window_values(pef_params(seq), bits, p, false, i0, cnt, docs_off, false, scratch_off);
    uint32_t* docs = smem_words(docs_off);
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) { const uint32_t e = lane + 32u * j; if (e >= cnt) docs[e] = 0xffffffffu; }
    __syncwarp();
    return PefDocsWindow{phase, staged_part, p.begin + i0, cnt};
}

// freq() - 1 of the postings at list positions [g0, g0 + cnt) -> the 128-word buffer at out_off.  positive_sequence over the strict
// prefix sums c[] (positive_sequence.hpp:48-66): freq(i) = c[i] - c[i-1].  The freqs sequence has its own partitions; a docs
// window may straddle two of them, so the range is decoded partition by partition.
__device__ __noinline__ void pef_window_freqs(PefSeq seq, uint32_t term, uint32_t g0, uint32_t cnt, uint32_t fcache_slot_off, uint32_t out_off, uint32_t scratch_off) {
    const unsigned lane = lane_id();
    uint32_t* out = smem_words(out_off);
    const AnyBits gb{seq.bits, 0, 0, 0};
    PefFreqSlot* fc = reinterpret_cast<PefFreqSlot*>(g_smem + fcache_slot_off);
    uint32_t fp = fc->fp;
    PefPart p = fc->part;
    uint32_t type = fc->type;
    uint64_t first_part = 0;
    bool have_first = false;
    bool cached = fp != 0xffffffffu && g0 >= p.begin && g0 - p.begin < p.size;
    if (!cached) {
        // partition of the freqs sequence holding position g0: last partition with begin <= g0 (32 probes per step)
        const PefListDir fl = seq.lists[term];
        first_part = fl.first_part; have_first = true;
        uint32_t lo = 0, hi = fl.nparts;
        while (hi - lo > 1) {
            const uint32_t span = hi - lo, step = (span + 31u) / 32u, pi = lo + lane * step;
            const bool le = pi < hi && __ldg(reinterpret_cast<const uint32_t*>(seq.parts + fl.first_part + pi) + 2) <= g0;      // PefPart::begin
            const unsigned m = __ballot_sync(FULL, le);
            const uint32_t f = 31u - __clz(m);
            lo = lo + f * step;
            hi = min(hi, lo + step);
        }
        fp = lo;
    }
    uint32_t g = g0, prev = 0;
    bool have_prev = false;
    while (g < g0 + cnt) {
        if (!cached) {
            if (!have_first) { first_part = seq.lists[term].first_part; have_first = true; }       // a window that runs past the cached partition
            p = pef_load_part(seq.parts, first_part + fp);
            type = (seq.raw_ef || p.ub - p.base + 1u == p.size) ? 0u : uint32_t(gb.word(p.bit_off >> 6) >> (p.bit_off & 63)) & 1u;
            __syncwarp();
            if (lane == 0) { fc->part = p; fc->fp = fp; fc->type = type; }
        }
        const uint32_t local = g - p.begin;
        const uint32_t take = min(g0 + cnt - g, p.size - local);
        // c[g0 - 1]: the value before a partition's first element is the previous partition's last value = base - 1
        // (partition 0: the sum before the first element is 0); otherwise the element before the window is decoded along
        const bool need_prev = !have_prev && local != 0;
        const uint32_t pv = pef_window_values(pef_params(seq), gb, p, true, local, take, out_off + 4u * (g - g0), need_prev, scratch_off, type);
        if (!have_prev) { prev = local ? pv : (fp ? p.base - 1u : 0u); have_prev = true; }
        g += take; ++fp;
        cached = false;
    }
    // prefix sums -> freq - 1, in place
    uint32_t v[5];
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) { const uint32_t e = 4u * lane + j; v[j + 1] = e < cnt ? out[e] : 0u; }
    v[0] = lane ? out[4u * lane - 1u] : prev;
    __syncwarp();
    uint4 r;
    r.x = v[1] - v[0] - 1u; r.y = v[2] - v[1] - 1u; r.z = v[3] - v[2] - 1u; r.w = v[4] - v[3] - 1u;
    reinterpret_cast<uint4*>(out)[lane] = r;
    __syncwarp();
}

// Start the copy of the driving list's block pair [e0, e1) (block b) into the driver window.  The window must not be read any
// more (the caller has decoded what it needs of the previous pair).
template <class List, class Ctx>
__device__ __forceinline__ void and_prefetch_driver(Ctx& c, const List* s, uint32_t b, uint32_t e0, uint32_t e1) {
    const uint64_t start = s->data_off + e0, end = s->data_off + e1;
    const uint64_t a0 = start & ~uint64_t(15);
    uint32_t bytes = uint32_t(((end + 15) & ~uint64_t(15)) - a0);
    if (bytes > STAGE_BYTES) bytes = STAGE_BYTES;
    __syncwarp();   // every lane is done reading the previous pair
    c.dwin_block = b; c.dwin_delta = uint32_t(start - a0) - e0;
    if (bytes) {
        if (lane_id() == 0) {
            uint64_t* dbar = c.bar + 1;
            mbar_expect_tx(dbar, bytes);
            tma_load_1d(c.stage + STAGE_WORDS, c.lists + a0, bytes, dbar);
        }
        c.dphase |= 2u;
    }
}
// the copy started by and_prefetch_driver has landed (waits once per copy)
template <class Ctx>
__device__ __forceinline__ void and_driver_ready(Ctx& c) {
    if (c.dphase & 2u) {
        mbar_wait(c.bar + 1, c.dphase & 1u);
        c.dphase = (c.dphase & 1u) ^ 1u;
    }
}

// block_posting_list.hpp:292-319 with the block's metadata in hand: [e0, e1) = byte range of the block
// pair inside the list's data, prev_max = block_max[b-1] (0xffffffff for b == 0), cur_max = block_max[b]
template <int CODEC, class List, class Ctx>
__device__ __forceinline__ void and_decode_docs(Ctx& c, DevIndex const& idx, List* s, uint32_t slot, uint32_t b, uint32_t e0, uint32_t e1, uint32_t prev_max, uint32_t cur_max) {
    if constexpr (CODEC == CODEC_PEF) {
        (void)e0; (void)prev_max;
        const PefDocsWindow w = pef_window_docs(idx.pdocs, s->data_off, b, e1, c.stage_off, smem_offset(c.bar), c.phase,
                                                c.win_slot == slot ? c.win_delta : 0xffffffffu, smem_offset(s->docs), c.fcache_off);
        c.phase = w.phase;
        if (w.staged_part != 0xffffffffu) { c.win_slot = slot; c.win_delta = w.staged_part; }
        // cur_end: the window's partition; freqs_off: list position of the window's first element; pad1: its size
        if (lane_id() == 0) { *reinterpret_cast<uint4*>(&s->cur_block) = make_uint4(b, cur_max, e1, w.first_pos); s->pad1 = w.cnt; }
        __syncwarp();
        DS2I_STAT(c.c_docs_blocks += 1; c.c_bytes_docs += w.cnt;)
        return;
    }
    const unsigned lane = lane_id();
    const uint32_t n = s->n;
    const uint32_t cur_base = prev_max + 1u;
    const uint32_t size = ((b + 1u) * BLOCK <= n) ? BLOCK : (n & (BLOCK - 1u));
    const uint64_t data_off = s->data_off;
    const bool in_dwin = slot == c.drv_slot && c.dwin_block == b;           // the pair was prefetched into the driver window
    uint32_t off, win_off;
    if (in_dwin) { and_driver_ready(c); off = e0 + c.dwin_delta; win_off = c.stage_off + STAGE_WORDS * 4u; }
    else { off = and_stage(c.lists, data_off + e0, data_off + e1, c.stage, c.bar, c.phase); win_off = c.stage_off; }
    bool prefix;
    const uint32_t consumed = and_decode_values<CODEC>(win_off, off, size, cur_max - cur_base - (size - 1u), smem_offset(s->docs), c.stack_off, prefix);
    if (prefix) {
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            uint32_t i = 32 * j + lane;
            s->docs[i] = i < size ? cur_base + s->docs[i] + i : 0xffffffffu;
        }
    } else {
        uint4 v = reinterpret_cast<uint4*>(s->docs)[lane];
        v.y += v.x; v.z += v.y; v.w += v.z;
        const uint32_t incl = warp_inclusive_scan(v.w);
        const uint32_t add = cur_base + (incl - v.w) + 4u * lane;     // docid_i = base + sum_{k<=i} gap_k + i
        v.x += add; v.y += add + 1u; v.z += add + 2u; v.w += add + 3u;
        reinterpret_cast<uint4*>(s->docs)[lane] = v;
    }
    if (lane == 0) *reinterpret_cast<uint4*>(&s->cur_block) = make_uint4(b, cur_max, e1, e0 + consumed);
    __syncwarp();
    if (!in_dwin) { c.win_slot = slot; c.win_delta = off - e0; }
    DS2I_STAT(c.c_docs_blocks += 1; c.c_bytes_docs += consumed;)
}

// freqs - 1 of the current block of `s` -> the 128-word buffer at out_off.  Right after the docs decode the
// block pair is still staged; a block carried over from an earlier candidate batch is staged again.
// Returns whether the buffer holds prefix sums (interpolative) instead of plain values.
template <int CODEC, class List, class Ctx>
__device__ __forceinline__ bool and_decode_freqs(Ctx& c, DevIndex const& idx, const List* s, uint32_t slot, uint32_t out_off) {
    if constexpr (CODEC == CODEC_PEF) {
        pef_window_freqs(idx.pfreqs, s->term, s->freqs_off, s->pad1, c.fcache_off + PEF_SCAN_SCRATCH_BYTES + slot * uint32_t(sizeof(PefFreqSlot)), out_off, c.fcache_off);
        DS2I_STAT(c.c_freqs_blocks += 1; c.c_bytes_freqs += s->pad1;)
        return false;
    }
    const uint32_t n = s->n, b = s->cur_block;
    const uint32_t size = ((b + 1u) * BLOCK <= n) ? BLOCK : (n & (BLOCK - 1u));
    const uint32_t freqs_off = s->freqs_off;
    const bool in_dwin = slot == c.drv_slot && c.dwin_block == b;
    if (!in_dwin && c.win_slot != slot) {
        const uint64_t data_off = s->data_off;
        const uint32_t off = and_stage(c.lists, data_off + freqs_off, data_off + s->cur_end, c.stage, c.bar, c.phase);
        c.win_slot = slot; c.win_delta = off - freqs_off;
    }
    bool prefix;
    const uint32_t consumed = in_dwin ? and_decode_values<CODEC>(c.stage_off + STAGE_WORDS * 4u, freqs_off + c.dwin_delta, size, 0xffffffffu, out_off, c.stack_off, prefix)
                                      : and_decode_values<CODEC>(c.stage_off, freqs_off + c.win_delta, size, 0xffffffffu, out_off, c.stack_off, prefix);
    DS2I_STAT(c.c_freqs_blocks += 1; c.c_bytes_freqs += consumed;)
    return prefix;
}

// first block index in [lo, nblocks) whose block_max >= bound (exists: bound <= last max), together
// with that block's metadata.  One 8-B aligned load per lane reads 32 consecutive (block_max,
// block_end) directory entries, so in the common short-skip case the block's byte range and base
// arrive with the probe; longer skips run a 32-ary search first.
struct BlockMeta { uint32_t block, e0, e1, prev_max, cur_max; };

template <class Ctx>
__device__ __forceinline__ BlockMeta and_find_block(Ctx& c, const uint2* bd, uint32_t nblocks, uint32_t lo, uint32_t lo_prev_max, uint32_t lo_prev_end,
                                                    uint32_t bound) {
    const unsigned lane = lane_id();
    uint32_t bi = lo + lane;
    uint2 en = bi < nblocks ? __ldg(bd + bi) : make_uint2(0xffffffffu, 0u);
    unsigned hit = __ballot_sync(FULL, en.x >= bound);
    DS2I_STAT(c.c_maxs += 32;)
    if (!hit) {
        uint32_t l2 = lo + 32, hi = nblocks - 1;     // invariant: max[hi] >= bound, every block < l2 has max < bound
        while (hi - l2 >= 31) {
            const uint32_t span = hi - l2;
            const uint32_t probe = l2 + uint32_t((uint64_t(span) * (lane + 1)) / 33);
            const uint32_t m = __ldg(bd + probe).x;
            const unsigned h = __ballot_sync(FULL, m >= bound);
            DS2I_STAT(c.c_maxs += 32;)
            if (h) {
                const uint32_t f = __ffs(h) - 1;
                const uint32_t nh = __shfl_sync(FULL, probe, f);
                if (f > 0) l2 = __shfl_sync(FULL, probe, f - 1) + 1;
                hi = nh;
            } else {
                l2 = __shfl_sync(FULL, probe, 31) + 1;
            }
        }
        // the answer lies in [l2, l2 + 30]; window from l2 - 1 so that its predecessor's entry comes along
        lo = l2 - 1;
        bi = lo + lane;
        en = bi < nblocks ? __ldg(bd + bi) : make_uint2(0xffffffffu, 0u);
        hit = __ballot_sync(FULL, en.x >= bound) & ~1u;
        DS2I_STAT(c.c_maxs += 32;)
    }
    const uint32_t f = __ffs(hit) - 1;
    BlockMeta r;
    r.block = lo + f;
    r.cur_max = __shfl_sync(FULL, en.x, f);
    r.e1 = __shfl_sync(FULL, en.y, f);
    const uint32_t pm = __shfl_sync(FULL, en.x, (f + 31) & 31);
    const uint32_t pe = __shfl_sync(FULL, en.y, (f + 31) & 31);
    r.prev_max = f ? pm : lo_prev_max;
    r.e0 = f ? pe : lo_prev_end;
    return r;
}

// position of the first element >= x in a sorted 128-entry block (padded with 0xffffffff)
__device__ __forceinline__ uint32_t lower_bound128(const uint32_t* d, uint32_t x) {
    uint32_t pos = 0;
#pragma unroll
    for (uint32_t s = 64; s >= 1; s >>= 1)
        if (d[pos + s - 1] < x) pos += s;
    return pos;
}

template <int CODEC, bool RANKED, int MIN_CTAS, bool STATS = true>
__global__ void __launch_bounds__(128, MIN_CTAS) and_block_kernel(DevIndex idx, DevWand wand, DevBatch batch, AndJob job, uint32_t k, int slots) {
    s16_table_init(smem_words(0));
    __syncthreads();

    const unsigned lane = lane_id();
    const unsigned warp = threadIdx.x >> 5;
    uint8_t* base = g_smem + S16_TAB_BYTES + warp * and_warp_smem_bytes(slots, CODEC == CODEC_PEF);
    AndWarp* ws = reinterpret_cast<AndWarp*>(base);
    AndList* st = reinterpret_cast<AndList*>(base + sizeof(AndWarp));
    uint32_t* ftmp = reinterpret_cast<uint32_t*>(base + sizeof(AndWarp) + size_t(slots) * sizeof(AndList));
    uint32_t* stage = ftmp + BLOCK;                 // probe window, then the driver window
    uint32_t* stack = stage + AND_STAGE_WINDOWS * STAGE_WORDS;

    AndCtxT<STATS> c;
    c.fcache_off = smem_offset(stack + SCRATCH_WORDS);
    c.drv_slot = AND_DRIVER_PREFETCH ? 0u : 0xffffffffu; c.dwin_block = 0xffffffffu; c.dwin_delta = 0; c.dphase = 0;
    c.lists = idx.lists; c.stage = stage; c.bar = &ws->bar;
    c.stage_off = smem_offset(stage); c.stack_off = smem_offset(stack); c.ftmp_off = smem_offset(ftmp);
    c.phase = 0; c.win_slot = 0xffffffffu; c.win_delta = 0;
    c.c_docs_blocks = c.c_freqs_blocks = c.c_bytes_docs = c.c_bytes_freqs = c.c_maxs = c.c_scored = 0;
    if (lane == 0) { mbar_init(c.bar, 1); mbar_init(c.bar + 1, 1); fence_mbar_init(); }
    __syncwarp();

    while (true) {
        uint32_t ii = 0;
        if (lane == 0) ii = atomicAdd(job.work_counter, 1u);
        ii = __shfl_sync(FULL, ii, 0);
        if (ii >= job.nitems) break;
        const uint32_t gpos = warp_upper_group(job.gstart, batch.nq, ii);
        const uint32_t q = batch.sched[gpos];
        const uint32_t chunk = ii - __ldg(job.gstart + gpos);
        const uint32_t chunk_blocks = __ldg(job.qchunk + q);
        const uint32_t first_block = chunk * chunk_blocks;
        ii = __ldg(job.item_begin + q) + chunk;      // result slot
        const uint32_t t0 = batch.q_begin[q];
        const uint32_t nt = batch.q_begin[q + 1] - t0;
        uint32_t matches = 0;
        TopK topk;
        topk.init(k);

        // slot i <- i-th list by increasing size (queries.hpp:357-360, the reference's own std::sort order)
        c.win_slot = 0xffffffffu;
        __syncwarp();
        if (lane < nt) {
            const uint32_t src = batch.ord_size[t0 + lane];
            if (RANKED) ws->qw[lane] = batch.q_weight[t0 + src];
            and_list_setup<CODEC>(idx, &st[lane], batch.term[t0 + src]);
            if (CODEC == CODEC_PEF) (reinterpret_cast<PefFreqSlot*>(g_smem + c.fcache_off + PEF_SCAN_SCRATCH_BYTES) + lane)->fp = 0xffffffffu;
        }
        __syncwarp();

        const uint32_t nb0 = st[0].nblocks;
        const uint32_t b_end = min(nb0, first_block + chunk_blocks);
        // directory entries of the whole chunk of the driving list, one block per lane, in one round trip
        uint32_t m_max = 0, m_end = 0, first_prev_max = 0xffffffffu, first_prev_end = 0;
        {
            static_assert(AND_CHUNK_BLOCKS <= 32, "one lane per block of the chunk");
            const uint2* bd0 = idx.bdir + st[0].bfirst;
            const uint32_t bi = first_block + lane;
            if (bi < b_end) { const uint2 en = __ldg(bd0 + bi); m_max = en.x; m_end = en.y; }
            if (first_block) {
                const uint2 en = __ldg(bd0 + first_block - 1);
                first_prev_max = en.x; first_prev_end = en.y;
            }
        }
        uint32_t b_begin = first_block;
        // ---- single-term queries (a tenth of the query log, but whole lists: a fifth of all driver blocks) ----
        if (!RANKED && CODEC != CODEC_PEF && nt == 1) {
            // and_query over one list counts its postings (queries.hpp:58-83 with an empty inner loop): nothing to decode
            matches = min(st[0].n, b_end * BLOCK) - first_block * BLOCK;
            b_begin = b_end;
        }
        bool exhausted = false;
        // byte range of driver block b inside the list's data: [end of block b - 1, end of block b)
        auto driver_range = [&](uint32_t b, uint32_t& e0, uint32_t& e1) {
            const uint32_t l = b - first_block;
            const uint32_t pe = __shfl_sync(FULL, m_end, (l + 31) & 31);
            e0 = l ? pe : first_prev_end; e1 = __shfl_sync(FULL, m_end, l);
        };
        c.dwin_block = 0xffffffffu;
        if (AND_DRIVER_PREFETCH && CODEC != CODEC_PEF && b_begin < b_end) {
            uint32_t e0, e1;
            driver_range(b_begin, e0, e1);
            and_prefetch_driver(c, &st[0], b_begin, e0, e1);
        }
        for (uint32_t b0 = b_begin; b0 < b_end && !exhausted; ++b0) {
            {
                const uint32_t l = b0 - first_block;
                const uint32_t pm = __shfl_sync(FULL, m_max, (l + 31) & 31), pe = __shfl_sync(FULL, m_end, (l + 31) & 31);
                and_decode_docs<CODEC>(c, idx, &st[0], 0u, b0, l ? pe : first_prev_end, __shfl_sync(FULL, m_end, l), l ? pm : first_prev_max,
                                       __shfl_sync(FULL, m_max, l));
            }
            const uint4 cv = reinterpret_cast<const uint4*>(st[0].docs)[lane];
            const uint32_t cand[4] = {cv.x, cv.y, cv.z, cv.w};
            uint32_t alive = 0;            // bit j: candidate 4*lane+j still matches every list probed so far
#pragma unroll
            for (int j = 0; j < 4; ++j) alive |= (cand[j] != 0xffffffffu) << j;

            // the driving list's own freqs (the block pair is staged right now), kept in registers
            uint32_t f0[4] = {0, 0, 0, 0};
            float norm_len[4] = {0.f, 0.f, 0.f, 0.f}, score[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t f0_bytes = 0;
            if (RANKED) {
                const uint32_t before = c.c_bytes_freqs;
                const bool prefix = and_decode_freqs<CODEC>(c, idx, &st[0], 0u, c.ftmp_off);
                f0_bytes = c.c_bytes_freqs - before;
                const uint4 fv = reinterpret_cast<const uint4*>(ftmp)[lane];
                f0[0] = fv.x; f0[1] = fv.y; f0[2] = fv.z; f0[3] = fv.w;
                if (prefix) {
                    const uint32_t prev = lane ? ftmp[4 * lane - 1] : 0u;
                    f0[3] -= f0[2]; f0[2] -= f0[1]; f0[1] -= f0[0]; f0[0] -= prev;
                }
                __syncwarp();
            }
            // everything this block needs of the driver window has been decoded: the next pair is copied while the other
            // lists are probed (the probes use the other staging window)
            if (AND_DRIVER_PREFETCH && CODEC != CODEC_PEF && b0 + 1 < b_end) {
                uint32_t e0, e1;
                driver_range(b0 + 1, e0, e1);
                and_prefetch_driver(c, &st[0], b0 + 1, e0, e1);
            }

            for (uint32_t i = 1; i < nt; ++i) {
                if (!__any_sync(FULL, alive)) break;
                AndList* s = &st[i];
                const uint2* bd = idx.bdir + s->bfirst;
                const uint32_t last_max = s->last_max;
                const float qwi = RANKED ? ws->qw[i] : 0.f;
                uint32_t pending = alive;  // alive candidates not yet looked up in list i
#ifdef DS2I_NIN_HIST
                uint32_t h_steps = 0;
                { const uint32_t np = __reduce_add_sync(FULL, __popc(pending)); DS2I_HIST(9 + (np <= 8 ? 0 : np <= 32 ? 1 : np <= 64 ? 2 : 3), 1); }
#endif
                while (true) {
                    uint32_t mine = 0xffffffffu;
#pragma unroll
                    for (int j = 3; j >= 0; --j) if (pending & (1u << j)) mine = cand[j];
                    const uint32_t cmin = __reduce_min_sync(FULL, mine);
                    if (cmin == 0xffffffffu) break;
                    if (cmin > last_max) {            // list i has nothing at or beyond cmin: those candidates die
                        alive &= ~pending;
                        exhausted = true;             // later blocks of list 0 only hold larger docids
                        break;
                    }
                    const uint32_t cur_block = s->cur_block;
                    if (cur_block == 0xffffffffu || cmin > s->cur_max) {
                        const bool fresh = cur_block == 0xffffffffu;
                        const BlockMeta bm = and_find_block(c, bd, s->nblocks, fresh ? 0u : cur_block + 1, fresh ? 0xffffffffu : s->cur_max,
                                                            fresh ? 0u : s->cur_end, cmin);
                        and_decode_docs<CODEC>(c, idx, s, i, bm.block, bm.e0, bm.e1, bm.prev_max, bm.cur_max);
                    }
                    const uint32_t cur_max = s->cur_max;
                    const uint32_t* d = s->docs;
                    // the candidates this block answers
                    uint32_t inb = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if ((pending & (1u << j)) && cand[j] <= cur_max) inb |= 1u << j;
                    pending &= ~inb;
                    const uint32_t nin = __reduce_add_sync(FULL, __popc(inb));
#ifdef DS2I_NIN_HIST
                    DS2I_HIST(hist_bucket(nin), 1); ++h_steps;
#endif
                    uint32_t hitmask = 0, pos[4] = {0, 0, 0, 0};
                    if (nin == 1) {
                        // the usual case when list i is much longer than the driving list: one candidate per block,
                        // compared with all 128 docids at once instead of a 7-step search on every lane
                        const uint4 v = reinterpret_cast<const uint4*>(d)[lane];
                        const uint32_t eq = (v.x == cmin ? 1u : 0u) | (v.y == cmin ? 2u : 0u) | (v.z == cmin ? 4u : 0u) | (v.w == cmin ? 8u : 0u);
                        const unsigned hb = __ballot_sync(FULL, eq != 0u);
                        if (hb) {
                            const uint32_t hl = __ffs(hb) - 1;
                            const uint32_t p = 4u * hl + __shfl_sync(FULL, uint32_t(__ffs(eq)) - 1u, hl);
                            hitmask = inb;
                            pos[0] = pos[1] = pos[2] = pos[3] = p;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (inb & (1u << j)) {
                                pos[j] = lower_bound128(d, cand[j]);
                                if (d[pos[j]] == cand[j]) hitmask |= 1u << j;
                            }
                    }
                    alive &= ~(inb & ~hitmask);
                    if (RANKED && __any_sync(FULL, hitmask)) {
                        if (i == 1) {
                            // first term of the sum (queries.hpp:374-379): the driving list's own weight; the
                            // norm_len gather is in flight while the freqs block is decoded
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (hitmask & (1u << j)) norm_len[j] = __ldg(wand.norm_lens + cand[j]);
                        }
                        const bool prefix = and_decode_freqs<CODEC>(c, idx, s, i, c.ftmp_off);
                        const float qw0 = ws->qw[0];
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (hitmask & (1u << j)) {
                                const uint32_t p = pos[j];
                                const uint32_t f = prefix ? ftmp[p] - (p ? ftmp[p - 1] : 0u) : ftmp[p];
                                if (i == 1) score[j] = qw0 * doc_term_weight(f0[j] + 1u, norm_len[j]);
                                score[j] += qwi * doc_term_weight(f + 1u, norm_len[j]);
                            }
                        __syncwarp();
                    }
                }
#ifdef DS2I_NIN_HIST
                DS2I_HIST(16 + hist_bucket(h_steps), 1); DS2I_HIST(24, h_steps);
#endif
            }
            DS2I_HIST(8, 1);

            const unsigned nalive = __popc(alive);
            const unsigned total = __reduce_add_sync(FULL, nalive);
            matches += total;
            if (RANKED) {
                if (!total) {
                    // the reference never decodes the freqs of a block without a match: not algorithmic work
                    DS2I_STAT(c.c_freqs_blocks -= 1; c.c_bytes_freqs -= f0_bytes;)
                } else {
                    DS2I_STAT(c.c_scored += total;)
                    if (nt == 1) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (alive & (1u << j)) norm_len[j] = __ldg(wand.norm_lens + cand[j]);
                        const float qw0 = ws->qw[0];
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (alive & (1u << j)) score[j] = qw0 * doc_term_weight(f0[j] + 1u, norm_len[j]);
                    }
                    // only scores that can enter the heap are inserted (serially, rare once the threshold is up)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        unsigned want = __ballot_sync(FULL, (alive & (1u << j)) && topk.would_enter(score[j]));
                        while (want) {
                            const int src = __ffs(want) - 1;
                            want &= want - 1;
                            topk.insert(__shfl_sync(FULL, score[j], src), __shfl_sync(FULL, cand[j], src));
                        }
                    }
                }
            }
        }

        if (AND_DRIVER_PREFETCH) and_driver_ready(c);       // a copy still in flight (the loop stopped early) must land before the window is reused
        if (lane == 0) { job.item_counts[ii] = matches; job.item_sizes[ii] = topk.size; }
        if (RANKED && lane < topk.size) {
            job.item_scores[size_t(ii) * 2 * k + lane] = topk.v;
            reinterpret_cast<uint32_t*>(job.item_scores)[size_t(ii) * 2 * k + k + lane] = topk.id;
        }
    }

    if (STATS && batch.stats && lane == 0) {
        atomicAdd(&batch.stats[0], (unsigned long long)c.c_docs_blocks);
        atomicAdd(&batch.stats[1], (unsigned long long)c.c_freqs_blocks);
        atomicAdd(&batch.stats[2], (unsigned long long)c.c_bytes_docs);
        atomicAdd(&batch.stats[3], (unsigned long long)c.c_bytes_freqs);
        atomicAdd(&batch.stats[4], (unsigned long long)c.c_maxs);
        atomicAdd(&batch.stats[5], (unsigned long long)c.c_scored);
    }
}

// fold the per-item partial results of each query: counts add up, top-k lists merge
__global__ void __launch_bounds__(128) merge_items_kernel(const uint32_t* item_begin /* nq+1 */, uint32_t nq, const uint32_t* item_counts,
                                                          const uint32_t* item_sizes, const float* item_scores, uint32_t k, bool ranked,
                                                          uint64_t* out_counts, float* out_scores, uint32_t* out_docids) {
    const unsigned lane = lane_id();
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const uint32_t i0 = item_begin[q], i1 = item_begin[q + 1];
    uint64_t count = 0;
    TopK topk;
    topk.init(k);
    for (uint32_t it = i0; it < i1; ++it) {
        count += item_counts[it];
        if (ranked) {
            uint32_t n = item_sizes[it];
            float v = lane < n ? item_scores[size_t(it) * 2 * k + lane] : 0.f;
            uint32_t vid = lane < n ? reinterpret_cast<const uint32_t*>(item_scores)[size_t(it) * 2 * k + k + lane] : 0xffffffffu;
            for (uint32_t j = 0; j < n; ++j) {
                float sc = __shfl_sync(FULL, v, j);
                if (!topk.would_enter(sc)) break;       // partial lists are sorted descending
                topk.insert(sc, __shfl_sync(FULL, vid, j));
            }
        }
    }
    if (lane == 0) out_counts[q] = ranked ? uint64_t(topk.size) : count;
    if (ranked && lane < k) {
        out_scores[size_t(q) * k + lane] = lane < topk.size ? topk.v : 0.f;
        out_docids[size_t(q) * k + lane] = lane < topk.size ? topk.id : 0xffffffffu;
    }
}

}  // namespace ds2i_gpu
