// Device-side block_posting_list::document_enumerator (block_posting_list.hpp:84-355): same
// next / next_geq / docid / freq / position / size semantics, one warp per enumerator call.
// State lives in shared memory (one ListState per query term); decoded docids are kept as absolute
// values (the reference keeps gaps and sums serially, block_posting_list.hpp:142-145), so
// next_geq inside a block is a 128-wide compare + one warp reduction.
#pragma once
#include "codecs.cuh"
#include "pef.cuh"

namespace ds2i_gpu {

constexpr uint32_t STAGE_BYTES = 1280;                 // staged window (16-B aligned superset of a block pair)
constexpr uint32_t STAGE_WORDS = STAGE_BYTES / 4 + 4;  // + slack for the w+1 word of unaligned reads
// Largest [docs | freqs] pair of 128-value blocks per codec: OptPFD raw blocks 2 x 4*129 = 1032 B (newpfor.h:204-209),
// varint-G8IU 2 x 64 groups x 9 B = 1152 B, QMX 2 x (512 B of 32-bit stripes + <= 32 keys + 2 length bytes) = 1092 B,
// interpolative (<= 128 values of < 32 bits) below those; + up to 15 B of alignment slack at either end of the window.
// ds2i_gpu_index_open refuses an index with a larger pair (DS2I_E_LIMIT), so the clamp in and_stage / stage_range never cuts data.
static_assert(STAGE_BYTES >= 2 * 576 + 30, "the staging window must hold the largest block pair of every codec");

struct ListDir {            // per posting list, built once at load time (host) from the EF endpoints
    uint64_t maxs_off;      // byte offset of block_maxs[] inside m_lists (just after TightVByte(n))
    uint32_t n;             // postings
    uint32_t data_bytes;    // bytes of block data (from the end of block_endpoints[] to the end of the list)
};

struct DevIndex {
    const uint8_t* lists;   // m_lists, 256-B aligned device copy with a zeroed tail pad
    const ListDir* dir;
    // aligned copy of every list's block_maxs[] / block_endpoints[] (byte-aligned only in the file,
    // block_posting_list.hpp:43,49), built once at load time: entry b of a list = (block_max[b], byte
    // offset inside the list's block data where block b ends), list t starts at bdir[bfirst[t]]
    const uint2* bdir;
    const uint32_t* bfirst;
    uint64_t num_lists;
    uint32_t num_docs;
    int codec;              // CODEC_*
    // Elias-Fano index family (codec == CODEC_PEF): the two bit-vector collections; bdir / bfirst then hold the window
    // directory of the docs sequences (last docid of every 128-element window, index of its partition inside the list;
    // bfirst has num_lists + 1 entries) and dir[] only carries n
    PefSeq pdocs, pfreqs;
};

struct ListState {
    uint64_t maxs_off;
    uint64_t data_off;      // absolute byte offset of the block data
    uint32_t n, nblocks, data_bytes;
    uint32_t cur_block, pos, cur_size, cur_max, cur_docid;
    uint32_t freqs_off;     // offset (from data_off) of the current block's freqs
    uint32_t block_end;     // offset (from data_off) of the end of the current block
    uint32_t freqs_ready;
    uint32_t pad;
    // used by the block-parallel kernels
    uint32_t prev_max;      // block_max of the block before the current one (0xffffffff for block 0)
    uint32_t last_max;      // last docid of the list
    uint32_t win_block;     // block the list sat on when the current window started
    uint32_t exhausted;
    uint32_t bfirst;        // the list's first entry in the block directory (DevIndex::bdir)
    uint32_t pad1, pad2, pad3;
    uint32_t docs[BLOCK];   // absolute docids of the current block (0xffffffff beyond cur_size)
    uint32_t freqs[BLOCK];  // freqs - 1 of the current block, valid when freqs_ready
};

// warp-private staging state + algorithmic-work counters (registers, warp-uniform)
struct WarpCtx {
    uint32_t* stage;        // STAGE_WORDS
    uint32_t* scratch;      // SCRATCH_WORDS
    uint64_t* bar;
    uint32_t phase;
    int codec;              // CODEC_* of the index (used when kernels are instantiated with CODEC_ANY)
    uint64_t win_start;     // absolute byte range currently staged
    uint32_t win_bytes;
    // counters (SURVEY.md §8d algorithmic bytes)
    uint32_t c_docs_blocks, c_freqs_blocks, c_docs_bytes, c_freqs_bytes, c_maxs, c_scored;
};

__device__ __forceinline__ void ctx_init(WarpCtx& c, uint32_t* stage, uint32_t* scratch, uint64_t* bar, int codec) {
    c.stage = stage; c.scratch = scratch; c.bar = bar; c.codec = codec;
    c.phase = 0; c.win_start = 0; c.win_bytes = 0;
    c.c_docs_blocks = c.c_freqs_blocks = c.c_docs_bytes = c.c_freqs_bytes = c.c_maxs = c.c_scored = 0;
    if (lane_id() == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncwarp();
}

// Make bytes [start, end) of m_lists available in the staging window; returns their byte offset
// inside the window.  One TMA bulk copy of the enclosing 16-B aligned range, completion on the
// warp's mbarrier.  A range already staged (e.g. the freqs block right behind a docs block that
// was just decoded) costs nothing.
__device__ __forceinline__ uint32_t stage_range(WarpCtx& c, const uint8_t* lists, uint64_t start, uint64_t end) {
    if (start >= c.win_start && end <= c.win_start + c.win_bytes) return uint32_t(start - c.win_start);
    uint64_t a0 = start & ~uint64_t(15);
    uint64_t a1 = (end + 15) & ~uint64_t(15);
    uint32_t bytes = uint32_t(a1 - a0);
    if (bytes > STAGE_BYTES) bytes = STAGE_BYTES;
    __syncwarp();   // every lane is done reading the previous window
    if (bytes) {
        if (lane_id() == 0) {
            mbar_expect_tx(c.bar, bytes);
            tma_load_1d(c.stage, lists + a0, bytes, c.bar);
        }
        mbar_wait(c.bar, c.phase);
        c.phase ^= 1u;
    }
    c.win_start = a0;
    c.win_bytes = bytes;
    return uint32_t(start - a0);
}

// values v[0..127] (gaps-1) in buf -> absolute docids base + sum_{k<=i} v[k] + i, in place
__device__ __forceinline__ void gaps_to_docids128(uint32_t* buf, uint32_t base) {
    const unsigned lane = lane_id();
    uint4 v = reinterpret_cast<uint4*>(buf)[lane];
    v.x += 1u; v.y += v.x + 1u; v.z += v.y + 1u; v.w += v.z + 1u;
    uint32_t incl = warp_inclusive_scan(v.w);
    uint32_t add = base + (incl - v.w) - 1u;
    v.x += add; v.y += add; v.z += add; v.w += add;
    reinterpret_cast<uint4*>(buf)[lane] = v;
    __syncwarp();
}

// one block of `size` values at window offset `off` -> buf; returns bytes consumed.
// prefix_out: interpolative leaves prefix sums (see codecs.cuh); others leave plain values.
template <int CODEC>
__device__ __forceinline__ uint32_t decode_values(WarpCtx& c, uint32_t off, uint32_t size, uint32_t sum_of_values,
                                                  uint32_t* buf, bool& prefix_out) {
    const int codec = (CODEC == CODEC_ANY) ? c.codec : CODEC;
    if (codec == CODEC_MIXED && size == BLOCK) {
        // mixed_block::decode (mixed_block.hpp:198-217): block_type byte, then that codec's block
        const uint32_t type = lds_u8(c.stage, off);
        prefix_out = type == 2u;
        if (type == 1u) return 1u + decode_varint128(smem_offset(c.stage), off + 1u, smem_offset(buf));
        if (type == 0u) return 1u + decode_optpfor128(smem_offset(c.stage), off + 1u, smem_offset(buf), smem_offset(c.scratch));
        return 1u + decode_interpolative_prefix(smem_offset(c.stage), off + 1u, size, sum_of_values, smem_offset(buf), smem_offset(c.scratch));
    }
    if (codec != CODEC_INTERPOLATIVE && codec != CODEC_MIXED && size == BLOCK) {
        prefix_out = false;
        if (codec == CODEC_OPTPFOR) return decode_optpfor128(smem_offset(c.stage), off, smem_offset(buf), smem_offset(c.scratch));
        if (codec == CODEC_VARINT) return decode_varint128(smem_offset(c.stage), off, smem_offset(buf));
        return decode_qmx128(smem_offset(c.stage), off, smem_offset(buf));
    }
    // n < block_size => every codec falls back to interpolative (block_codecs.hpp:196-199,215-217)
    prefix_out = true;
    return decode_interpolative_prefix(smem_offset(c.stage), off, size, sum_of_values, smem_offset(buf), smem_offset(c.scratch));
}

// Inlining policy of the block decoder wrappers.  Measured on B200 (round 1): __noinline__ here shrinks the
// kernels but costs more in spilled WarpCtx state than it saves in instruction-cache misses
// (ranked_and 26 -> 30 ms, wand 160 -> 183 ms per 10k-query batch), so they stay inlined.
#ifndef DS2I_DECODE_INLINE
#define DS2I_DECODE_INLINE __forceinline__
#endif

template <int CODEC>
struct BlockEnum {
    typedef ListState State;
    typedef DevIndex Index;
    // document_enumerator ctor + reset(): block_posting_list.hpp:86-108
    static __device__ __forceinline__ void open(WarpCtx& c, DevIndex const& idx, ListState* st, uint32_t term) {
        ListDir d = idx.dir[term];
        uint32_t nblocks = (d.n + BLOCK - 1) / BLOCK;
        if (lane_id() == 0) {
            st->maxs_off = d.maxs_off;
            st->data_off = d.maxs_off + 4ull * nblocks + 4ull * (nblocks - 1);
            st->n = d.n; st->nblocks = nblocks; st->data_bytes = d.data_bytes;
        }
        __syncwarp();
        decode_docs_block(c, idx, st, 0);
    }

    // block_posting_list.hpp:292-319
    static __device__ __forceinline__ void decode_docs_block(WarpCtx& c, DevIndex const& idx, ListState* st, uint32_t b) {
        const unsigned lane = lane_id();
        const uint32_t nblocks = st->nblocks;
        const uint8_t* maxs = idx.lists + st->maxs_off;
        const uint8_t* ends = maxs + 4ull * nblocks;
        // lanes 0..3 fetch endpoint[b-1], endpoint[b], max[b-1], max[b] with one converged load each
        uint32_t v = 0;
        {
            const bool is_end = lane < 2;
            const uint32_t which = lane & 1u;                       // 0: entry b-1, 1: entry b
            const bool have = lane < 4 && (which ? (is_end ? (b + 1 < nblocks) : true) : (b != 0));
            const uint8_t* p = (is_end ? ends : maxs) + 4ull * (b + which) - 4ull;
            if (have) v = ldg_u32_unaligned(p);
            else if (lane == 1) v = st->data_bytes;
            else if (lane == 2) v = 0xffffffffu;
        }
        const uint32_t e0 = __shfl_sync(FULL, v, 0), e1 = __shfl_sync(FULL, v, 1);
        const uint32_t prev_max = __shfl_sync(FULL, v, 2);
        const uint32_t cur_max = __shfl_sync(FULL, v, 3);
        decode_docs_block_meta(c, idx, st, b, e0, e1, prev_max, cur_max);
    }

    // the same with the block's metadata already in hand: e0/e1 = byte range of the block inside the
    // list's data, prev_max = block_max[b-1] (0xffffffff for b == 0), cur_max = block_max[b]
    static __device__ DS2I_DECODE_INLINE void decode_docs_block_meta(WarpCtx& c, DevIndex const& idx, ListState* st, uint32_t b,
                                                                  uint32_t e0, uint32_t e1, uint32_t prev_max, uint32_t cur_max) {
        const unsigned lane = lane_id();
        const uint32_t n = st->n;
        const uint32_t cur_base = prev_max + 1u;
        const uint32_t size = ((b + 1) * BLOCK <= n) ? BLOCK : (n % BLOCK);
        const uint64_t data_off = st->data_off;

        uint32_t off = stage_range(c, idx.lists, data_off + e0, data_off + e1);
        bool prefix;
        uint32_t consumed = decode_values<CODEC>(c, off, size, cur_max - cur_base - (size - 1u), st->docs, prefix);
        if (prefix) {
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) {
                uint32_t i = 32 * j + lane;
                st->docs[i] = i < size ? cur_base + st->docs[i] + i : 0xffffffffu;
            }
            __syncwarp();
        } else {
            gaps_to_docids128(st->docs, cur_base);
        }
        if (lane == 0) {
            st->cur_block = b; st->pos = 0; st->cur_size = size; st->cur_max = cur_max; st->prev_max = prev_max;
            st->cur_docid = st->docs[0];
            st->freqs_off = e0 + consumed; st->block_end = e1; st->freqs_ready = 0;
        }
        __syncwarp();
        c.c_docs_blocks += 1; c.c_docs_bytes += consumed;
    }

    // block_posting_list.hpp:321-331
    static __device__ __forceinline__ void decode_freqs_block(WarpCtx& c, DevIndex const& idx, ListState* st) {
        const unsigned lane = lane_id();
        const uint32_t size = st->cur_size;
        uint32_t off = stage_range(c, idx.lists, st->data_off + st->freqs_off, st->data_off + st->block_end);
        bool prefix;
        uint32_t consumed = decode_values<CODEC>(c, off, size, 0xffffffffu, st->freqs, prefix);
        if (prefix) {
            uint32_t d[4];
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) {
                uint32_t i = 32 * j + lane;
                d[j] = (i < size) ? st->freqs[i] - (i ? st->freqs[i - 1] : 0u) : 0u;
            }
            __syncwarp();
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) st->freqs[32 * j + lane] = d[j];
        }
        if (lane == 0) st->freqs_ready = 1;
        __syncwarp();
        c.c_freqs_blocks += 1; c.c_freqs_bytes += consumed;
    }

    static __device__ __forceinline__ uint32_t docid(const ListState* st) { return st->cur_docid; }

    // block_posting_list.hpp:110-122; returns the new docid()
    static __device__ __forceinline__ uint32_t next(WarpCtx& c, DevIndex const& idx, ListState* st) {
        uint32_t pos = st->pos + 1;
        if (pos == st->cur_size) {
            if (st->cur_block + 1 == st->nblocks) {
                __syncwarp();
                if (lane_id() == 0) { st->pos = pos; st->cur_docid = idx.num_docs; }
                __syncwarp();
                return idx.num_docs;
            }
            decode_docs_block(c, idx, st, st->cur_block + 1);
            return st->cur_docid;
        }
        uint32_t d = st->docs[pos & (BLOCK - 1)];
        __syncwarp();
        if (lane_id() == 0) { st->pos = pos; st->cur_docid = d; }
        __syncwarp();
        return d;
    }

    // block_posting_list.hpp:124-146; returns the new docid()
    static __device__ __forceinline__ uint32_t next_geq(WarpCtx& c, DevIndex const& idx, ListState* st, uint32_t lower_bound) {
        const unsigned lane = lane_id();
        const uint32_t cur = st->cur_docid;
        if (cur == idx.num_docs || cur >= lower_bound) return cur;   // past the end stays there; never moves backwards
        if (lower_bound > st->cur_max) {
            const uint32_t nblocks = st->nblocks;
            const uint8_t* maxs = idx.lists + st->maxs_off;
            // the reference checks the last block first, then scans block_max linearly (:129-137);
            // here 32 entries per step with a ballot — same block found.
            uint32_t last = ldg_u32_unaligned(maxs + 4ull * (nblocks - 1));
            c.c_maxs += 1;
            if (lower_bound > last) {
                __syncwarp();
                if (lane == 0) st->cur_docid = idx.num_docs;
                __syncwarp();
                return idx.num_docs;
            }
            uint32_t block = st->cur_block + 1;
            while (true) {
                uint32_t bi = block + lane;
                uint32_t m = bi < nblocks ? ldg_u32_unaligned(maxs + 4ull * bi) : 0xffffffffu;
                unsigned hit = __ballot_sync(FULL, m >= lower_bound);
                if (hit) { uint32_t f = __ffs(hit) - 1; c.c_maxs += f + 1; block += f; break; }
                c.c_maxs += 32;
                block += 32;
            }
            decode_docs_block(c, idx, st, block);
        }
        // first position with docid >= lower_bound (exists: lower_bound <= cur_max)
        uint4 v = reinterpret_cast<const uint4*>(st->docs)[lane];
        uint32_t cnt = (v.x < lower_bound) + (v.y < lower_bound) + (v.z < lower_bound) + (v.w < lower_bound);
        uint32_t pos = __reduce_add_sync(FULL, cnt);
        uint32_t cp = st->pos;
        if (pos < cp) pos = cp;
        uint32_t d = st->docs[pos & (BLOCK - 1)];
        __syncwarp();
        if (lane == 0) { st->pos = pos; st->cur_docid = d; }
        __syncwarp();
        return d;
    }

    // block_posting_list.hpp:165-171
    static __device__ __forceinline__ uint32_t freq(WarpCtx& c, DevIndex const& idx, ListState* st) {
        if (!st->freqs_ready) decode_freqs_block(c, idx, st);
        return st->freqs[st->pos & (BLOCK - 1)] + 1u;
    }

    static __device__ __forceinline__ uint64_t position(const ListState* st) { return uint64_t(st->cur_block) * BLOCK + st->pos; }
    static __device__ __forceinline__ uint32_t size(const ListState* st) { return st->n; }
};

}  // namespace ds2i_gpu
