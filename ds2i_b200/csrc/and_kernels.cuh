// Block-at-a-time conjunctive evaluation: and_query / ranked_and_query (queries.hpp:35-86,322-401)
// re-thought for a warp.  The reference advances ONE candidate docid at a time through next_geq;
// here the 128 docids of a block of the shortest list are the candidates, carried 4 per lane, and
// every other list is probed for all of them at once (block_max search by ballot, 128-wide binary
// search of the decoded block in shared memory).  Survivors are scored in parallel — BM25 summed in
// the reference's order (lists by increasing size), so every score is bit-identical — and only
// scores that beat the running threshold reach the (serial) top-k insert.
//
// Work item = (query, chunk of CH consecutive blocks of its shortest list): heavy queries are spread
// over many warps, each with a private top-k; merge_items_kernel folds the partial results.  The
// top-k multiset and the match count do not depend on the evaluation order, so results equal the
// reference's exactly.
#pragma once
#include "query_kernels.cuh"

namespace ds2i_gpu {

constexpr uint32_t AND_CHUNK_BLOCKS = 32;     // blocks of the shortest list per work item
constexpr int AND_SMALL_TERMS = 0;            // queries up to this many terms run in the high-occupancy launch

struct AndItem { uint32_t query, first_block; };

struct AndJob {
    const AndItem* items;      // in query order (item_begin[q] .. item_begin[q+1])
    const uint32_t* order;     // processing order: items of the costliest queries first
    uint32_t nitems;
    uint32_t chunk_blocks;     // blocks of the shortest list per item (<= 32)
    uint32_t* work_counter;
    uint32_t* item_counts;     // nitems: matches found by the item
    uint32_t* item_sizes;      // nitems: entries in the item's partial top-k
    float* item_scores;        // nitems * k
};

// first block index in [lo, nblocks) whose block_max >= bound (exists: bound <= last max), together
// with that block's metadata.  The first probe reads 32 consecutive block_max entries AND the
// matching block_endpoints in the same step, so in the common short-skip case the block's byte range
// and base arrive with the probe (one memory round trip instead of two); longer skips finish with a
// 32-ary search and a separate metadata fetch.
struct BlockMeta { uint32_t block, e0, e1, prev_max, cur_max; bool have; };

__device__ DS2I_DECODE_INLINE BlockMeta find_block(WarpCtx& c, const ListState* s, const uint8_t* maxs, uint32_t lo, uint32_t lo_prev_max, uint32_t bound) {
    const unsigned lane = lane_id();
    const uint32_t nblocks = s->nblocks;
    const uint8_t* ends = maxs + 4ull * nblocks;
    BlockMeta r;
    {
        uint32_t bi = lo + lane;
        uint32_t m = bi < nblocks ? ldg_u32_unaligned(maxs + 4ull * bi) : 0xffffffffu;
        uint32_t e = (bi < nblocks && bi) ? ldg_u32_unaligned(ends + 4ull * bi - 4ull) : 0u;    // start of block bi
        unsigned hit = __ballot_sync(FULL, m >= bound);
        c.c_maxs += 32;
        if (hit) {
            uint32_t f = __ffs(hit) - 1;
            r.block = lo + f;
            r.cur_max = __shfl_sync(FULL, m, f);
            r.e0 = __shfl_sync(FULL, e, f);
            uint32_t pm = __shfl_sync(FULL, m, (f + 31) & 31);
            r.prev_max = f ? pm : lo_prev_max;
            uint32_t en = __shfl_sync(FULL, e, (f + 1) & 31);
            r.have = true;
            if (r.block + 1 >= nblocks) r.e1 = s->data_bytes;
            else if (f < 31) r.e1 = en;
            else r.have = false;          // end offset not among the 32 probed entries
            return r;
        }
        lo += 32;
    }
    uint32_t hi = nblocks - 1;          // invariant: max[hi] >= bound, every block < lo has max < bound
    while (hi - lo >= 32) {
        uint32_t span = hi - lo;
        uint32_t probe = lo + uint32_t((uint64_t(span) * (lane + 1)) / 33);
        uint32_t m = ldg_u32_unaligned(maxs + 4ull * probe);
        unsigned hit = __ballot_sync(FULL, m >= bound);
        c.c_maxs += 32;
        if (hit) {
            uint32_t f = __ffs(hit) - 1;
            uint32_t nh = __shfl_sync(FULL, probe, f);
            if (f > 0) lo = __shfl_sync(FULL, probe, f - 1) + 1;
            hi = nh;
        } else {
            lo = __shfl_sync(FULL, probe, 31) + 1;
        }
    }
    uint32_t bi = lo + lane;
    uint32_t m = bi <= hi ? ldg_u32_unaligned(maxs + 4ull * bi) : 0xffffffffu;
    unsigned hit = __ballot_sync(FULL, m >= bound);
    c.c_maxs += hi - lo + 1;
    r.block = lo + (__ffs(hit) - 1);
    r.have = false; r.e0 = r.e1 = r.prev_max = r.cur_max = 0;
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

template <int CODEC, bool RANKED>
__global__ void __launch_bounds__(128) and_block_kernel(DevIndex idx, DevWand wand, DevBatch batch, AndJob job, uint32_t k, int slots) {
    s16_table_init(smem_words(0));
    __syncthreads();

    typedef BlockEnum<CODEC> E;
    const unsigned lane = lane_id();
    const unsigned warp = threadIdx.x >> 5;
    uint8_t* base = g_smem + S16_TAB_BYTES + warp * warp_smem_bytes(slots);
    WarpSmem* ws = reinterpret_cast<WarpSmem*>(base);
    ListState* st = reinterpret_cast<ListState*>(base + sizeof(WarpSmem));
    uint32_t* stage = reinterpret_cast<uint32_t*>(base + sizeof(WarpSmem) + size_t(slots) * sizeof(ListState));
    uint32_t* scratch = stage + STAGE_WORDS;

    WarpCtx c;
    ctx_init(c, stage, scratch, &ws->bar, idx.codec);

    while (true) {
        uint32_t ii = 0;
        if (lane == 0) ii = atomicAdd(job.work_counter, 1u);
        ii = __shfl_sync(FULL, ii, 0);
        if (ii >= job.nitems) break;
        ii = job.order[ii];
        const AndItem item = job.items[ii];
        const uint32_t q = item.query;
        const uint32_t t0 = batch.q_begin[q];
        const uint32_t nt = batch.q_begin[q + 1] - t0;
        uint32_t matches = 0;
        TopK topk;
        topk.init(k);

        // slot i <- i-th list by increasing size (queries.hpp:357-360, the reference's own std::sort order)
        __syncwarp();
        if (lane < nt) {
            uint32_t src = batch.ord_size[t0 + lane];
            if (RANKED) ws->qw[lane] = batch.q_weight[t0 + src];
            ListDir d = idx.dir[batch.term[t0 + src]];
            uint32_t nblocks = (d.n + BLOCK - 1) / BLOCK;
            ListState* s = &st[lane];
            s->maxs_off = d.maxs_off;
            s->data_off = d.maxs_off + 4ull * nblocks + 4ull * (nblocks - 1);
            s->n = d.n; s->nblocks = nblocks; s->data_bytes = d.data_bytes;
            s->cur_block = 0xffffffffu;     // not positioned yet
            s->cur_max = 0;
            s->pad = ldg_u32_unaligned(idx.lists + d.maxs_off + 4ull * (nblocks - 1));   // last docid of the list
        }
        __syncwarp();

        const uint32_t nb0 = st[0].nblocks;
        const uint32_t b_end = min(nb0, item.first_block + job.chunk_blocks);
        // metadata of the whole chunk of the driving list, one block per lane, fetched in one round trip
        uint32_t m_max = 0, m_start = 0, m_end = 0, m_first_prev;
        {
            static_assert(AND_CHUNK_BLOCKS <= 32, "one lane per block of the chunk");
            const uint8_t* maxs0 = idx.lists + st[0].maxs_off;
            const uint8_t* ends0 = maxs0 + 4ull * nb0;
            uint32_t bi = item.first_block + lane;
            if (bi < b_end) {
                m_max = ldg_u32_unaligned(maxs0 + 4ull * bi);
                m_start = bi ? ldg_u32_unaligned(ends0 + 4ull * bi - 4ull) : 0u;
                m_end = bi + 1 < nb0 ? ldg_u32_unaligned(ends0 + 4ull * bi) : st[0].data_bytes;
            }
            uint32_t pm = (lane == 0 && item.first_block) ? ldg_u32_unaligned(maxs0 + 4ull * item.first_block - 4ull) : 0xffffffffu;
            m_first_prev = __shfl_sync(FULL, pm, 0);
        }
        bool exhausted = false;
        for (uint32_t b0 = item.first_block; b0 < b_end && !exhausted; ++b0) {
            {
                uint32_t l = b0 - item.first_block;
                uint32_t pm = __shfl_sync(FULL, m_max, (l + 31) & 31);
                E::decode_docs_block_meta(c, idx, &st[0], b0, __shfl_sync(FULL, m_start, l), __shfl_sync(FULL, m_end, l),
                                          l ? pm : m_first_prev, __shfl_sync(FULL, m_max, l));
            }
            if (b0 + 1 < nb0) prefetch_l2(idx.lists + st[0].data_off + st[0].block_end + lane * 32u);   // next block of the driving list
            uint4 cv = reinterpret_cast<const uint4*>(st[0].docs)[lane];
            uint32_t cand[4] = {cv.x, cv.y, cv.z, cv.w};
            uint32_t alive = 0;            // bit j: candidate 4*lane+j still matches every list probed so far
#pragma unroll
            for (int j = 0; j < 4; ++j) alive |= (cand[j] != 0xffffffffu) << j;

            for (uint32_t i = 1; i < nt; ++i) {
                if (!__any_sync(FULL, alive)) break;
                ListState* s = &st[i];
                const uint8_t* maxs = idx.lists + s->maxs_off;
                const uint32_t last_max = s->pad;
                uint32_t pending = alive;  // alive candidates not yet looked up in list i
                while (true) {
                    uint32_t mine = 0xffffffffu;
#pragma unroll
                    for (int j = 3; j >= 0; --j) if (pending & (1u << j)) mine = cand[j];
                    uint32_t cmin = __reduce_min_sync(FULL, mine);
                    if (cmin == 0xffffffffu) break;
                    if (cmin > last_max) {            // list i has nothing at or beyond cmin: those candidates die
                        alive &= ~pending;
                        // later blocks of list 0 only hold larger docids
                        exhausted = true;
                        break;
                    }
                    uint32_t cur_block = s->cur_block;
                    if (cur_block == 0xffffffffu || cmin > s->cur_max) {
                        bool fresh = cur_block == 0xffffffffu;
                        BlockMeta bm = find_block(c, s, maxs, fresh ? 0u : cur_block + 1, fresh ? 0xffffffffu : s->cur_max, cmin);
                        if (bm.have) E::decode_docs_block_meta(c, idx, s, bm.block, bm.e0, bm.e1, bm.prev_max, bm.cur_max);
                        else E::decode_docs_block(c, idx, s, bm.block);
                    }
                    const uint32_t cur_max = s->cur_max;
                    const uint32_t* d = s->docs;
                    uint32_t hitmask = 0, pos[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        pos[j] = 0;
                        if ((pending & (1u << j)) && cand[j] <= cur_max) {
                            pos[j] = lower_bound128(d, cand[j]);
                            if (d[pos[j]] == cand[j]) hitmask |= 1u << j;
                            else alive &= ~(1u << j);
                            pending &= ~(1u << j);
                        }
                    }
                    if (RANKED && __any_sync(FULL, hitmask)) {
                        // freqs of this block: decoded into list 0's (otherwise idle) freqs buffer, the
                        // matched ones parked in list i's buffer under the candidate's slot
                        uint32_t* ftmp = st[0].freqs;
                        uint32_t off = stage_range(c, idx.lists, s->data_off + s->freqs_off, s->data_off + s->block_end);
                        bool prefix;
                        uint32_t size = s->cur_size;
                        uint32_t consumed = decode_values<CODEC>(c, off, size, 0xffffffffu, ftmp, prefix);
                        c.c_freqs_blocks += 1; c.c_freqs_bytes += consumed;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (hitmask & (1u << j)) {
                                uint32_t p = pos[j];
                                uint32_t f = prefix ? ftmp[p] - (p ? ftmp[p - 1] : 0u) : ftmp[p];
                                s->freqs[4 * lane + j] = f;
                            }
                        __syncwarp();
                    }
                }
            }

            unsigned nalive = __popc(alive);
            unsigned total = __reduce_add_sync(FULL, nalive);
            matches += total;
            if (RANKED && total) {
                // the driving list's own freqs, then BM25 in list order (queries.hpp:374-379)
                ListState* s0 = &st[0];
                uint32_t off = stage_range(c, idx.lists, s0->data_off + s0->freqs_off, s0->data_off + s0->block_end);
                bool prefix;
                uint32_t consumed = decode_values<CODEC>(c, off, s0->cur_size, 0xffffffffu, s0->freqs, prefix);
                c.c_freqs_blocks += 1; c.c_freqs_bytes += consumed;
                c.c_scored += total;
                float score[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    score[j] = 0.f;
                    if (alive & (1u << j)) {
                        uint32_t slot = 4 * lane + j;
                        float norm_len = __ldg(wand.norm_lens + cand[j]);
                        uint32_t f0 = prefix ? s0->freqs[slot] - (slot ? s0->freqs[slot - 1] : 0u) : s0->freqs[slot];
                        float sc = 0.f;
                        sc += ws->qw[0] * doc_term_weight(f0 + 1u, norm_len);
                        for (uint32_t i = 1; i < nt; ++i) sc += ws->qw[i] * doc_term_weight(st[i].freqs[slot] + 1u, norm_len);
                        score[j] = sc;
                    }
                }
                __syncwarp();
                // only scores that can enter the heap are inserted (serially, rare once the threshold is up)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    unsigned want = __ballot_sync(FULL, (alive & (1u << j)) && topk.would_enter(score[j]));
                    while (want) {
                        int src = __ffs(want) - 1;
                        want &= want - 1;
                        float sc = __shfl_sync(FULL, score[j], src);
                        topk.insert(sc);
                    }
                }
            }
        }

        if (lane == 0) { job.item_counts[ii] = matches; job.item_sizes[ii] = topk.size; }
        if (RANKED && lane < k) job.item_scores[size_t(ii) * k + lane] = lane < topk.size ? topk.v : 0.f;
    }

    if (batch.stats && lane == 0) {
        atomicAdd(&batch.stats[0], (unsigned long long)c.c_docs_blocks);
        atomicAdd(&batch.stats[1], (unsigned long long)c.c_freqs_blocks);
        atomicAdd(&batch.stats[2], (unsigned long long)c.c_docs_bytes);
        atomicAdd(&batch.stats[3], (unsigned long long)c.c_freqs_bytes);
        atomicAdd(&batch.stats[4], (unsigned long long)c.c_maxs);
        atomicAdd(&batch.stats[5], (unsigned long long)c.c_scored);
    }
}

// fold the per-item partial results of each query: counts add up, top-k lists merge
__global__ void __launch_bounds__(128) merge_items_kernel(const uint32_t* item_begin /* nq+1 */, uint32_t nq, const uint32_t* item_counts,
                                                          const uint32_t* item_sizes, const float* item_scores, uint32_t k, bool ranked,
                                                          uint64_t* out_counts, float* out_scores) {
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
            float v = lane < n ? item_scores[size_t(it) * k + lane] : 0.f;
            for (uint32_t j = 0; j < n; ++j) {
                float sc = __shfl_sync(FULL, v, j);
                if (!topk.would_enter(sc)) break;       // partial lists are sorted descending
                topk.insert(sc);
            }
        }
    }
    if (lane == 0) out_counts[q] = ranked ? uint64_t(topk.size) : count;
    if (ranked && lane < k) out_scores[size_t(q) * k + lane] = lane < topk.size ? topk.v : 0.f;
}

}  // namespace ds2i_gpu
