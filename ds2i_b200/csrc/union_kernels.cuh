// Block-parallel dynamic-pruning top-k over the UNION of the query's lists: the device path behind
// wand_query and maxscore_query (queries.hpp:200-319, 478-591).
//
// Both reference operators are rank-safe: they return the k best BM25 scores of the documents that
// contain at least one query term, and differ only in how they skip documents that provably cannot
// enter the heap.  Their control flow is one document at a time with the heap threshold fed back
// after every insert — sequential.  This kernel keeps MaxScore's pruning logic (lists sorted by
// max_weight, prefix sums ub[], "essential" lists whose postings are the only candidates,
// non-essential lists probed from the highest bound down while score + ub[i] can still enter,
// queries.hpp:519-578) but evaluates it 128 candidates at a time:
//   * a window = the docids up to the smallest current block_max of the essential lists; the
//     postings of each essential list inside the window are candidates, 4 per lane;
//   * essential contributions: 128-wide binary search of the other essential lists' decoded blocks
//     (a document is scored by the first essential list that contains it);
//   * non-essential lists: the block-at-a-time probing of and_kernels.cuh, restricted to the
//     candidates for which would_enter(score + ub[i]) still holds;
//   * only scores above the running threshold reach the serial top-k insert.
// Pruning with a stale (lower) threshold is always safe, so the top-k multiset equals the
// reference's; per-document sums follow MaxScore's order (essential lists by increasing max_weight,
// then non-essential ones downwards) but which lists are essential depends on the threshold
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

template <int CODEC>
struct UnionOps {
    typedef BlockEnum<CODEC> E;

    // freqs of the list's current block as plain values (freq - 1) in dst[0..size)
    static __device__ DS2I_DECODE_INLINE void decode_freqs_plain(WarpCtx& c, DevIndex const& idx, ListState* s, uint32_t* dst) {
        const unsigned lane = lane_id();
        uint32_t off = stage_range(c, idx.lists, s->data_off + s->freqs_off, s->data_off + s->block_end);
        bool prefix;
        uint32_t size = s->cur_size;
        uint32_t consumed = decode_values<CODEC>(c, off, size, 0xffffffffu, dst, prefix);
        c.c_freqs_blocks += 1; c.c_freqs_bytes += consumed;
        if (prefix) {
            uint32_t d[4];
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) {
                uint32_t i = 32 * j + lane;
                d[j] = (i < size) ? dst[i] - (i ? dst[i - 1] : 0u) : 0u;
            }
            __syncwarp();
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) dst[32 * j + lane] = d[j];
            __syncwarp();
        }
    }

    static __device__ __forceinline__ uint32_t count_less(const ListState* s, uint32_t bound) {
        uint4 v = reinterpret_cast<const uint4*>(s->docs)[lane_id()];
        uint32_t cnt = (v.x < bound) + (v.y < bound) + (v.z < bound) + (v.w < bound);
        return __reduce_add_sync(FULL, cnt);
    }

    // essential list: docs + freqs of block b, cursor at its first element
    static __device__ __forceinline__ void load_essential_block(WarpCtx& c, DevIndex const& idx, ListState* s, BlockMeta const& bm) {
        E::decode_docs_block_meta(c, idx, s, bm.block, bm.e0, bm.e1, bm.prev_max, bm.cur_max);
        const uint32_t b = bm.block;
        if (b + 1 < s->nblocks) prefetch_l2(idx.lists + s->data_off + s->block_end + lane_id() * 32u);
        decode_freqs_plain(c, idx, s, s->freqs);
        if (lane_id() == 0) s->freqs_ready = 1;
        __syncwarp();
    }
};

constexpr size_t UNION_MERGE_BYTES = 128 * 4 + 128 * 4 + 128 * 2;
__host__ __device__ constexpr size_t union_warp_smem_bytes(int slots) { return warp_smem_bytes(slots) + UNION_MERGE_BYTES; }

template <int CODEC>
__global__ void __launch_bounds__(128) union_block_kernel(DevIndex idx, DevWand wand, DevBatch batch, UnionJob job, uint32_t k, int slots) {
    s16_table_init(smem_words(0));
    __syncthreads();

    typedef BlockEnum<CODEC> E;
    typedef UnionOps<CODEC> U;
    const unsigned lane = lane_id();
    const unsigned warp = threadIdx.x >> 5;
    // per warp: [WarpSmem | slots x ListState | merge buffers | staging window | codec scratch]
    uint8_t* base = g_smem + S16_TAB_BYTES + warp * union_warp_smem_bytes(slots);
    WarpSmem* ws = reinterpret_cast<WarpSmem*>(base);
    ListState* st = reinterpret_cast<ListState*>(base + sizeof(WarpSmem));
    uint32_t* cdoc = reinterpret_cast<uint32_t*>(base + sizeof(WarpSmem) + size_t(slots) * sizeof(ListState));   // merged candidate docids
    uint32_t* ftmp = cdoc + 128;                                                                              // freqs of a probed block
    uint16_t* cinfo = reinterpret_cast<uint16_t*>(ftmp + 128);                                                // (list << 8) | slot per candidate
    uint32_t* stage = reinterpret_cast<uint32_t*>(base + sizeof(WarpSmem) + size_t(slots) * sizeof(ListState) + UNION_MERGE_BYTES);
    uint32_t* scratch = stage + STAGE_WORDS;

    WarpCtx c;
    ctx_init(c, stage, scratch, &ws->bar, idx.codec);

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
        __syncwarp();
        if (lane < nt) {
            uint32_t src = batch.ord_maxw[t0 + lane];
            ws->qw[lane] = batch.q_weight[t0 + src];
            ws->mw[lane] = batch.max_weight[t0 + src];
            ListDir d = idx.dir[batch.term[t0 + src]];
            uint32_t nblocks = (d.n + BLOCK - 1) / BLOCK;
            ListState* s = &st[lane];
            s->maxs_off = d.maxs_off;
            s->data_off = d.maxs_off + 4ull * nblocks + 4ull * (nblocks - 1);
            s->n = d.n; s->nblocks = nblocks; s->data_bytes = d.data_bytes;
            s->cur_block = 0xffffffffu; s->cur_max = 0; s->prev_max = 0xffffffffu; s->cur_size = 0; s->pos = 0;
            s->freqs_ready = 0; s->win_block = 0xffffffffu;
            s->bfirst = idx.bfirst[batch.term[t0 + src]];
            s->last_max = __ldg(idx.bdir + s->bfirst + nblocks - 1).x;
            s->exhausted = s->last_max < lo ? 1u : 0u;
        }
        __syncwarp();
        if (lane == 0) {
            float acc = ws->mw[0];
            ws->ub[0] = acc;
            for (uint32_t i = 1; i < nt; ++i) { acc = acc + ws->mw[i]; ws->ub[i] = acc; }
        }
        __syncwarp();

        uint32_t ne = 0;                 // lists [0, ne) are non-essential
        uint32_t positioned = 0;         // bit i: list i has been positioned as an essential list
        while (true) {
            // refresh the shared floor, grow the non-essential prefix (queries.hpp:568-574)
            {
                uint32_t g = 0;
                if (lane == 0) g = *reinterpret_cast<volatile uint32_t*>(job.query_threshold + q);
                g = __shfl_sync(FULL, g, 0);
                topk.floor_ = fmaxf(topk.floor_, __uint_as_float(g));
            }
            while (ne < nt && !topk.would_enter(ws->ub[ne])) ne += 1;
            if (ne == nt) break;

            // essential lists that have not been opened yet: position them at the start of the range
            for (uint32_t e = ne; e < nt; ++e) {
                if (positioned & (1u << e)) continue;
                positioned |= 1u << e;
                ListState* s = &st[e];
                if (s->exhausted) continue;
                uint32_t cb = s->cur_block;
                if (cb == 0xffffffffu || s->cur_max < lo) {
                    // never touched (or left behind as a probed list): find the block holding the first docid >= lo
                    bool fresh = cb == 0xffffffffu;
                    BlockMeta bm = and_find_block(c.c_maxs, idx.bdir + s->bfirst, s->nblocks, fresh ? 0u : cb + 1, fresh ? 0xffffffffu : s->cur_max,
                                                  fresh ? 0u : s->block_end, lo);
                    U::load_essential_block(c, idx, s, bm);
                    uint32_t p = U::count_less(s, lo);
                    if (lane == 0) s->pos = p;
                    __syncwarp();
                } else if (!s->freqs_ready) {
                    // was a probed (non-essential-style) list positioned inside the range: keep its cursor
                    U::decode_freqs_plain(c, idx, s, s->freqs);
                    if (lane == 0) s->freqs_ready = 1;
                    __syncwarp();
                }
            }

            // window: everything up to the smallest current block_max of the live essential lists, shrunk until
            // the essential postings inside it number at most 128 (one merged candidate batch per window)
            uint32_t w_hi = 0xffffffffu;
            bool any = false;
            for (uint32_t e = ne; e < nt; ++e) {
                const ListState* s = &st[e];
                if (s->exhausted) continue;
                any = true;
                w_hi = min(w_hi, s->cur_max);
            }
            if (!any) break;
            if (w_hi >= hi) w_hi = hi - 1;
            uint32_t total;
            while (true) {
                total = 0;
                uint32_t best_cnt = 0, best_e = ne;
                for (uint32_t e = ne; e < nt; ++e) {
                    ListState* s = &st[e];
                    uint32_t end_e = s->exhausted ? s->pos : U::count_less(s, w_hi + 1u);     // elements <= w_hi
                    uint32_t cnt = end_e > s->pos ? end_e - s->pos : 0u;
                    if (lane == 0) s->pad = s->pos + cnt;                                        // window end inside the block
                    if (cnt > best_cnt) { best_cnt = cnt; best_e = e; }
                    total += cnt;
                }
                __syncwarp();
                if (total <= 128u) break;
                uint32_t take = max(1u, best_cnt * 120u / total);
                w_hi = st[best_e].docs[st[best_e].pos + take - 1u];
            }

            if (total) {
                // merged, sorted candidate batch: rank of a posting = postings of the window that precede it
                // (ties between lists broken by list order, so copies of one document sit next to each other
                // in increasing max_weight order — the reference's summation order, queries.hpp:545-554)
                for (uint32_t e = ne; e < nt; ++e) {
                    const ListState* se = &st[e];
                    const uint32_t pos_e = se->pos, end_e = se->pad;
                    if (end_e <= pos_e) continue;
                    uint4 cv = reinterpret_cast<const uint4*>(se->docs)[lane];
                    const uint32_t x[4] = {cv.x, cv.y, cv.z, cv.w};
                    uint32_t rank[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) rank[j] = 4 * lane + j - pos_e;
                    for (uint32_t e2 = ne; e2 < nt; ++e2) {
                        if (e2 == e) continue;
                        const ListState* s2 = &st[e2];
                        const uint32_t pos2 = s2->pos, end2 = s2->pad;
                        if (end2 <= pos2) continue;
                        const uint32_t* d2 = s2->docs;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t slot = 4 * lane + j;
                            if (slot >= pos_e && slot < end_e) {
                                uint32_t lb = lower_bound128(d2, x[j]);
                                rank[j] += lb - pos2;
                                if (e2 < e && lb < end2 && d2[lb] == x[j]) rank[j] += 1u;
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t slot = 4 * lane + j;
                        if (slot >= pos_e && slot < end_e) { cdoc[rank[j] & 127u] = x[j]; cinfo[rank[j] & 127u] = uint16_t((e << 8) | slot); }
                    }
                }
                __syncwarp();

                // every lane owns 4 consecutive entries of the merged batch; an entry that repeats its
                // predecessor's docid is a copy and is folded into the first one
                uint32_t cand[4], alive = 0;
                float score[4], nl[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t r = 4 * lane + j;
                    cand[j] = r < total ? cdoc[r] : 0xffffffffu;
                    score[j] = 0.f; nl[j] = 0.f;
                    if (r < total && (r == 0 || cdoc[r - 1] != cand[j])) {
                        alive |= 1u << j;
                        nl[j] = __ldg(wand.norm_lens + cand[j]);
                        for (uint32_t t = 0; t < nt - ne; ++t) {
                            uint32_t r2 = r + t;
                            if (r2 >= total || cdoc[r2] != cand[j]) break;
                            uint32_t info = cinfo[r2];
                            score[j] += ws->qw[info >> 8] * doc_term_weight(st[info >> 8].freqs[info & 127u] + 1u, nl[j]);
                        }
                    }
                }
                c.c_scored += __reduce_add_sync(FULL, __popc(alive));

                // non-essential lists from the highest bound down (queries.hpp:557-566); candidates are sorted,
                // so every list is probed in one forward pass
                uint32_t probing = alive;
                for (uint32_t i = ne; i-- > 0;) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if ((probing & (1u << j)) && !topk.would_enter(score[j] + ws->ub[i])) probing &= ~(1u << j);
                    if (!__any_sync(FULL, probing)) break;
                    ListState* s = &st[i];
                    if (s->exhausted) continue;
                    const uint2* bd = idx.bdir + s->bfirst;
                    uint32_t pending = probing;
                    while (true) {
                        uint32_t mine = 0xffffffffu;
#pragma unroll
                        for (int j = 3; j >= 0; --j) if (pending & (1u << j)) mine = cand[j];
                        uint32_t cmin = __reduce_min_sync(FULL, mine);
                        if (cmin == 0xffffffffu) break;
                        if (cmin > s->last_max) break;                 // nothing of list i at or beyond cmin
                        uint32_t cur_block = s->cur_block;
                        if (cur_block == 0xffffffffu || cmin > s->cur_max) {
                            bool fresh = cur_block == 0xffffffffu;
                            BlockMeta bm = and_find_block(c.c_maxs, bd, s->nblocks, fresh ? 0u : cur_block + 1, fresh ? 0xffffffffu : s->cur_max,
                                                          fresh ? 0u : s->block_end, cmin);
                            E::decode_docs_block_meta(c, idx, s, bm.block, bm.e0, bm.e1, bm.prev_max, bm.cur_max);
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
                            U::decode_freqs_plain(c, idx, s, ftmp);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (hitmask & (1u << j)) score[j] += ws->qw[i] * doc_term_weight(ftmp[pos[j]] + 1u, nl[j]);
                            __syncwarp();
                        }
                    }
                }

                // heap: only scores that can still enter
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
                __syncwarp();
                if (lane >= ne && lane < nt) st[lane].pos = st[lane].pad;       // cursors move past the window
                __syncwarp();
            }

            // publish the threshold, then move consumed essential lists to their next block
            if (topk.t.size == topk.t.k && topk.t.thr > topk.floor_) {
                if (lane == 0) atomicMax(job.query_threshold + q, __float_as_uint(topk.t.thr));
            }
            bool range_done = (w_hi + 1u >= hi);
            for (uint32_t e = ne; e < nt; ++e) {
                ListState* s = &st[e];
                if (s->exhausted) continue;
                if (range_done) { if (lane == 0) s->exhausted = 1; continue; }
                if (s->pos >= s->cur_size) {
                    uint32_t nb = s->cur_block + 1;
                    if (nb >= s->nblocks || s->cur_max + 1u >= hi) { __syncwarp(); if (lane == 0) s->exhausted = 1; __syncwarp(); }
                    else {
                        const uint2 en = __ldg(idx.bdir + s->bfirst + nb);     // the next block's (block_max, end)
                        U::load_essential_block(c, idx, s, BlockMeta{nb, s->block_end, en.y, s->cur_max, en.x});
                    }
                }
            }
            __syncwarp();
            if (range_done) break;
        }

        if (lane == 0) job.item_sizes[ii] = topk.t.size;
        if (lane < k) job.item_scores[size_t(ii) * k + lane] = lane < topk.t.size ? topk.t.v : 0.f;
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

}  // namespace ds2i_gpu
