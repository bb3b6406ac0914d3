// Block-parallel dynamic-pruning top-k over the UNION of the query's lists: the device path behind
// wand_query and maxscore_query (queries.hpp:200-319, 478-591).
//
// Both reference operators are rank-safe: they return the k best BM25 scores of the documents that
// contain at least one query term, and differ only in how they skip documents that provably cannot
// enter the heap.  Their control flow is one document at a time with the heap threshold fed back
// after every insert — sequential.  This kernel keeps MaxScore's pruning logic (lists sorted by
// max_weight, prefix sums ub[], "essential" lists whose postings are the only candidates,
// non-essential lists probed from the highest bound down while score + ub[i] can still enter,
// queries.hpp:519-578) but evaluates it a docid TILE at a time:
//   * tile = UNION_TILE consecutive docids, with one fp32 accumulator per docid in shared memory;
//   * essential lists, in increasing max_weight order, add q_weight * doc_term_weight of their postings
//     inside the tile to the accumulators (one list at a time, so every document's sum is formed in
//     the reference's order, queries.hpp:545-554, without atomics);
//   * the touched accumulators are the candidates, 128 consecutive docids (4 per lane) per pass;
//     non-essential lists are probed block-at-a-time as in and_kernels.cuh, restricted to the
//     candidates for which would_enter(score + ub[i]) still holds;
//   * only scores above the running threshold reach the serial top-k insert.
// Tiles without essential postings are never visited (the next tile is the one holding the smallest
// unconsumed essential posting).  Pruning with a stale (lower) threshold is always safe, so the
// top-k multiset equals the reference's; which lists are essential depends on the threshold
// history, so scores can differ from the reference in the last bit — within the 1e-5 relative
// tolerance of the north star (the literal, bit-exact kernels stay available: DS2I_RUN_FAITHFUL).
//
// Work item = (query, docid range).  Items of one query share a monotone global threshold
// (atomicMax on the float bits), so splitting a heavy query over many warps keeps its pruning power.
#pragma once
#include "and_kernels.cuh"

namespace ds2i_gpu {

struct UnionItem { uint32_t query, lo, hi; };      // docid range [lo, hi)

struct UnionJob {
    const UnionItem* items;      // in query order
    const uint32_t* order;       // processing order
    uint32_t nitems;
    uint32_t* work_counter;
    uint32_t* query_threshold;   // nq: float bits of the best published k-th score of the query
    uint32_t* item_sizes;
    float* item_scores;          // nitems * k
};

// top-k with a floor shared across the items of a query
struct TopKShared {
    TopK t;
    float floor_;     // published k-th best score of the whole query so far (0 = none)
    __device__ __forceinline__ void init(uint32_t k) { t.init(k); floor_ = 0.f; }
    __device__ __forceinline__ float bar() const { return t.size < t.k ? floor_ : fmaxf(t.thr, floor_); }
    __device__ __forceinline__ bool would_enter(float s) const { return s > bar(); }
    __device__ __forceinline__ void insert(float s) { if (would_enter(s)) t.insert(s); }
};

constexpr uint32_t UNION_TILE = 512;            // docids per tile (accumulators: 2 KB per warp)

// per query term: cursor + the decoded docids and freqs of the current block
struct UnionList {
    uint64_t data_off;      // absolute byte offset of the list's block data inside m_lists
    uint32_t bfirst;        // the list's first entry in the block directory
    uint32_t nblocks;
    uint32_t n;
    uint32_t last_max;      // last docid of the list
    uint32_t pos;           // essential cursor: first unconsumed posting of the current block
    uint32_t done;          // as an essential list: nothing left inside the item's docid range
    // written together by lane 0 after every docs decode (one 16-B store)
    uint32_t cur_block;     // 0xffffffff: not positioned yet
    uint32_t cur_max;
    uint32_t cur_end;
    uint32_t freqs_off;
    uint32_t docs[BLOCK];   // absolute docids of the current block (0xffffffff beyond its size)
    uint32_t freqs[BLOCK];  // freqs - 1 of the current block (kept for essential lists only)
};
static_assert(sizeof(UnionList) == 48 + 8 * BLOCK, "UnionList layout");

struct UnionWarp {
    float qw[MAX_TERMS];
    float ub[MAX_TERMS];
    uint64_t bar;
    uint64_t pad;
};

__host__ __device__ constexpr size_t union_warp_smem_bytes(int slots) {
    return sizeof(UnionWarp) + size_t(slots) * sizeof(UnionList) + UNION_TILE * 4 /* accumulators */ + BLOCK * 4 /* freqs of a probed block */ +
           STAGE_WORDS * 4 + SCRATCH_WORDS * 4;
}

__device__ __forceinline__ uint32_t union_block_size(const UnionList* s) {
    const uint32_t n = s->n, b = s->cur_block;
    return ((b + 1u) * BLOCK <= n) ? BLOCK : (n & (BLOCK - 1u));
}

// essential list: docs + plain freqs of block bm.block, cursor at its first element
template <int CODEC>
__device__ __forceinline__ void union_load_block(AndCtx& c, UnionList* s, uint32_t slot, BlockMeta const& bm) {
    const unsigned lane = lane_id();
    and_decode_docs<CODEC>(c, s, slot, bm.block, bm.e0, bm.e1, bm.prev_max, bm.cur_max);
    const bool prefix = and_decode_freqs<CODEC>(c, s, slot, smem_offset(s->freqs));
    if (prefix) {       // interpolative leaves prefix sums
        const uint32_t size = union_block_size(s);
        uint32_t d[4];
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t i = 32 * j + lane;
            d[j] = (i < size) ? s->freqs[i] - (i ? s->freqs[i - 1] : 0u) : 0u;
        }
        __syncwarp();
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) s->freqs[32 * j + lane] = d[j];
    }
    if (lane == 0) s->pos = 0;
    __syncwarp();
}

template <int CODEC>
__global__ void __launch_bounds__(128, 4) union_block_kernel(DevIndex idx, DevWand wand, DevBatch batch, UnionJob job, uint32_t k, int slots) {
    s16_table_init(smem_words(0));
    __syncthreads();

    const unsigned lane = lane_id();
    const unsigned warp = threadIdx.x >> 5;
    // per warp: [UnionWarp | slots x UnionList | accumulators | freqs buffer | staging window | codec scratch]
    uint8_t* base = g_smem + S16_TAB_BYTES + warp * union_warp_smem_bytes(slots);
    UnionWarp* ws = reinterpret_cast<UnionWarp*>(base);
    UnionList* st = reinterpret_cast<UnionList*>(base + sizeof(UnionWarp));
    float* acc = reinterpret_cast<float*>(base + sizeof(UnionWarp) + size_t(slots) * sizeof(UnionList));
    uint32_t* ftmp = reinterpret_cast<uint32_t*>(acc + UNION_TILE);
    uint32_t* stage = ftmp + BLOCK;
    uint32_t* stack = stage + STAGE_WORDS;

    AndCtx c;
    c.lists = idx.lists; c.stage = stage; c.bar = &ws->bar;
    c.stage_off = smem_offset(stage); c.stack_off = smem_offset(stack); c.ftmp_off = smem_offset(ftmp);
    c.phase = 0; c.win_slot = 0xffffffffu; c.win_delta = 0;
    c.c_docs_blocks = c.c_freqs_blocks = c.c_bytes_docs = c.c_bytes_freqs = c.c_maxs = c.c_scored = 0;
    if (lane == 0) { mbar_init(c.bar, 1); fence_mbar_init(); }
#pragma unroll
    for (uint32_t p = 0; p < UNION_TILE / 128; ++p) reinterpret_cast<float4*>(acc)[32 * p + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    while (true) {
        uint32_t ii = 0;
        if (lane == 0) ii = atomicAdd(job.work_counter, 1u);
        ii = __shfl_sync(FULL, ii, 0);
        if (ii >= job.nitems) break;
        ii = job.order[ii];
        const UnionItem item = job.items[ii];
        const uint32_t q = item.query;
        const uint32_t t0 = batch.q_begin[q];
        const uint32_t nt = batch.q_begin[q + 1] - t0;
        const uint32_t lo = item.lo, hi = item.hi;
        TopKShared topk;
        topk.init(k);

        // slots in increasing max_weight order (queries.hpp:521-524), upper bounds by sequential prefix sum (:526-530)
        c.win_slot = 0xffffffffu;
        __syncwarp();
        if (lane < nt) {
            const uint32_t src = batch.ord_maxw[t0 + lane];
            ws->qw[lane] = batch.q_weight[t0 + src];
            ws->ub[lane] = batch.max_weight[t0 + src];
            const uint32_t term = batch.term[t0 + src];
            const ListDir d = idx.dir[term];
            const uint32_t nblocks = (d.n + BLOCK - 1) / BLOCK;
            const uint32_t bfirst = idx.bfirst[term];
            UnionList* s = &st[lane];
            s->data_off = d.maxs_off + 4ull * nblocks + 4ull * (nblocks - 1);
            s->bfirst = bfirst; s->nblocks = nblocks; s->n = d.n;
            s->last_max = __ldg(idx.bdir + bfirst + nblocks - 1).x;
            s->pos = 0;
            s->done = s->last_max < lo ? 1u : 0u;
            s->cur_block = 0xffffffffu; s->cur_max = 0; s->cur_end = 0; s->freqs_off = 0;
        }
        __syncwarp();
        if (lane == 0) {
            float a = ws->ub[0];
            for (uint32_t i = 1; i < nt; ++i) { a = a + ws->ub[i]; ws->ub[i] = a; }
        }
        __syncwarp();

        // lists [0, ne) are non-essential (queries.hpp:568-574); the shared floor may already exclude some
        uint32_t ne = 0;
        {
            uint32_t g = 0;
            if (lane == 0) g = *reinterpret_cast<volatile uint32_t*>(job.query_threshold + q);
            g = __shfl_sync(FULL, g, 0);
            topk.floor_ = fmaxf(topk.floor_, __uint_as_float(g));
        }
        while (ne < nt && !topk.would_enter(ws->ub[ne])) ne += 1;

        // essential lists: position at the first posting >= lo
        for (uint32_t e = ne; e < nt; ++e) {
            UnionList* s = &st[e];
            if (s->done) continue;
            const BlockMeta bm = and_find_block(c.c_maxs, idx.bdir + s->bfirst, s->nblocks, 0u, 0xffffffffu, 0u, lo);
            union_load_block<CODEC>(c, s, e, bm);
            const uint4 v = reinterpret_cast<const uint4*>(s->docs)[lane];
            const uint32_t p = __reduce_add_sync(FULL, (v.x < lo) + (v.y < lo) + (v.z < lo) + (v.w < lo));
            if (lane == 0) s->pos = p;
            __syncwarp();
        }

        while (ne < nt) {
            // next tile: the one holding the smallest unconsumed essential posting
            uint32_t nxt = 0xffffffffu;
            for (uint32_t e = ne; e < nt; ++e) {
                const UnionList* s = &st[e];
                if (!s->done) nxt = min(nxt, s->docs[s->pos]);
            }
            if (nxt >= hi) break;
            const uint32_t tile = nxt & ~(UNION_TILE - 1u);
            const uint32_t limit = min(tile + UNION_TILE, hi);

            // essential contributions, one list at a time in increasing max_weight order
            for (uint32_t e = ne; e < nt; ++e) {
                UnionList* s = &st[e];
                if (s->done) continue;
                const float qwe = ws->qw[e];
                while (true) {
                    const uint32_t pos = s->pos;
                    const uint4 dv = reinterpret_cast<const uint4*>(s->docs)[lane];
                    const uint4 fv = reinterpret_cast<const uint4*>(s->freqs)[lane];
                    const uint32_t d[4] = {dv.x, dv.y, dv.z, dv.w};
                    const uint32_t f[4] = {fv.x, fv.y, fv.z, fv.w};
                    uint32_t below = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool in = d[j] < limit;       // the 0xffffffff padding beyond the block's size never is
                        below += in;
                        if (in && 4 * lane + j >= pos) {
                            const float nl = __ldg(wand.norm_lens + d[j]);
                            acc[d[j] - tile] += qwe * doc_term_weight(f[j] + 1u, nl);
                        }
                    }
                    const uint32_t npos = __reduce_add_sync(FULL, below);
                    __syncwarp();
                    if (npos < union_block_size(s)) {       // the rest of the block lies beyond the tile
                        if (lane == 0) s->pos = npos;
                        break;
                    }
                    const uint32_t nb = s->cur_block + 1;
                    if (nb >= s->nblocks || s->cur_max + 1u >= hi) {
                        if (lane == 0) s->done = 1;
                        break;
                    }
                    const uint2 en = __ldg(idx.bdir + s->bfirst + nb);     // the next block's (block_max, end)
                    union_load_block<CODEC>(c, s, e, BlockMeta{nb, s->cur_end, en.y, s->cur_max, en.x});
                }
                __syncwarp();
            }

            // candidates: the touched accumulators, 128 consecutive docids per pass
#pragma unroll 1
            for (uint32_t p = 0; p < UNION_TILE / 128; ++p) {
                const float4 a = reinterpret_cast<const float4*>(acc)[32 * p + lane];
                float score[4] = {a.x, a.y, a.z, a.w};
                uint32_t alive = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) alive |= (score[j] > 0.f) << j;
                if (!__any_sync(FULL, alive)) continue;
                reinterpret_cast<float4*>(acc)[32 * p + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
                uint32_t cand[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) cand[j] = (alive & (1u << j)) ? tile + 128 * p + 4 * lane + j : 0xffffffffu;
                c.c_scored += __reduce_add_sync(FULL, __popc(alive));

                // non-essential lists from the highest bound down (queries.hpp:557-566); candidates are sorted,
                // so every list is probed in one forward pass
                uint32_t probing = alive;
                for (uint32_t i = ne; i-- > 0;) {
                    const float ubi = ws->ub[i];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if ((probing & (1u << j)) && !topk.would_enter(score[j] + ubi)) probing &= ~(1u << j);
                    if (!__any_sync(FULL, probing)) break;
                    UnionList* s = &st[i];
                    if (s->last_max < lo) continue;
                    const uint2* bd = idx.bdir + s->bfirst;
                    const uint32_t last_max = s->last_max;
                    const float qwi = ws->qw[i];
                    uint32_t pending = probing;
                    while (true) {
                        uint32_t mine = 0xffffffffu;
#pragma unroll
                        for (int j = 3; j >= 0; --j) if (pending & (1u << j)) mine = cand[j];
                        const uint32_t cmin = __reduce_min_sync(FULL, mine);
                        if (cmin == 0xffffffffu) break;
                        if (cmin > last_max) break;                 // nothing of list i at or beyond cmin
                        const uint32_t cur_block = s->cur_block;
                        if (cur_block == 0xffffffffu || cmin > s->cur_max) {
                            const bool fresh = cur_block == 0xffffffffu;
                            const BlockMeta bm = and_find_block(c.c_maxs, bd, s->nblocks, fresh ? 0u : cur_block + 1, fresh ? 0xffffffffu : s->cur_max,
                                                                fresh ? 0u : s->cur_end, cmin);
                            and_decode_docs<CODEC>(c, s, i, bm.block, bm.e0, bm.e1, bm.prev_max, bm.cur_max);
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
                                pending &= ~(1u << j);
                            }
                        }
                        if (__any_sync(FULL, hitmask)) {
                            float nl[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) nl[j] = (hitmask & (1u << j)) ? __ldg(wand.norm_lens + cand[j]) : 0.f;
                            const bool prefix = and_decode_freqs<CODEC>(c, s, i, c.ftmp_off);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (hitmask & (1u << j)) {
                                    const uint32_t pp = pos[j];
                                    const uint32_t fq = prefix ? ftmp[pp] - (pp ? ftmp[pp - 1] : 0u) : ftmp[pp];
                                    score[j] += qwi * doc_term_weight(fq + 1u, nl[j]);
                                }
                            __syncwarp();
                        }
                    }
                }

                // heap: only scores that can still enter
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    unsigned want = __ballot_sync(FULL, (alive & (1u << j)) && topk.would_enter(score[j]));
                    while (want) {
                        const int src = __ffs(want) - 1;
                        want &= want - 1;
                        topk.insert(__shfl_sync(FULL, score[j], src));
                    }
                }
            }
            __syncwarp();

            // publish the threshold, refresh the shared floor, grow the non-essential prefix
            if (topk.t.size == topk.t.k && topk.t.thr > topk.floor_) {
                if (lane == 0) atomicMax(job.query_threshold + q, __float_as_uint(topk.t.thr));
            }
            {
                uint32_t g = 0;
                if (lane == 0) g = *reinterpret_cast<volatile uint32_t*>(job.query_threshold + q);
                g = __shfl_sync(FULL, g, 0);
                topk.floor_ = fmaxf(topk.floor_, __uint_as_float(g));
            }
            while (ne < nt && !topk.would_enter(ws->ub[ne])) ne += 1;
        }

        if (lane == 0) job.item_sizes[ii] = topk.t.size;
        if (lane < k) job.item_scores[size_t(ii) * k + lane] = lane < topk.t.size ? topk.t.v : 0.f;
    }

    if (batch.stats && lane == 0) {
        atomicAdd(&batch.stats[0], (unsigned long long)c.c_docs_blocks);
        atomicAdd(&batch.stats[1], (unsigned long long)c.c_freqs_blocks);
        atomicAdd(&batch.stats[2], (unsigned long long)c.c_bytes_docs);
        atomicAdd(&batch.stats[3], (unsigned long long)c.c_bytes_freqs);
        atomicAdd(&batch.stats[4], (unsigned long long)c.c_maxs);
        atomicAdd(&batch.stats[5], (unsigned long long)c.c_scored);
    }
}

}  // namespace ds2i_gpu
