// Block-parallel dynamic-pruning top-k over the UNION of the query's lists: the device path behind
// wand_query and maxscore_query (queries.hpp:200-319, 478-591).
//
// Both reference operators are rank-safe: they return the k best BM25 scores of the documents that
// contain at least one query term, and differ only in how they skip documents that provably cannot
// enter the heap.  Their control flow is one document at a time with the heap threshold fed back
// after every insert — sequential.  This kernel keeps MaxScore's bounds (lists sorted by max_weight,
// prefix sums ub[], queries.hpp:519-530) but turns the document-at-a-time merge into independent,
// list-driven work:
//   * every document is OWNED by the highest-weight list that contains it.  A work item is
//     (query, driving list e, a run of consecutive blocks of e); the 128 docids of a block are the
//     candidates, 4 per lane, exactly as in and_kernels.cuh;
//   * lists above e are probed for docids only: a hit means the document belongs to that list's own
//     items, so the candidate is dropped here.  No merge of the essential lists is needed;
//   * a document whose lists all lie at or below e scores at most ub[e].  Once the query's shared
//     threshold reaches ub[e] the list is non-essential (queries.hpp:568-574) and every remaining
//     item of e returns at once — its documents that matter are owned by higher lists;
//   * survivors get their own weight (the driver's freqs are decoded only now) and are completed
//     against the lists below e from the highest bound down, while score + ub[i] can still enter
//     (queries.hpp:557-566);
//   * only scores above the running threshold reach the serial top-k insert.
// Items of one query share a monotone global threshold (atomicMax on the float bits); items are
// processed highest-weight lists first so that the long, low-weight lists usually find the
// threshold already above their bound.  Pruning with a stale (lower) threshold is always safe, so
// the top-k multiset equals the reference's.  Per-document sums run from the owning list downwards
// (the reference adds essential lists upwards first), so a score can differ from the reference in
// the last bits — far inside the 1e-5 relative tolerance of the north star; the literal, bit-exact
// kernels stay available (DS2I_RUN_FAITHFUL).
#pragma once
#include "and_kernels.cuh"

namespace ds2i_gpu {

// Work items are implicit: group g = (query, list slot) owns ceil(nblocks / item_blocks) consecutive items.  Groups are
// laid out in processing order; a warp maps a global item number to its group with a 32-ary search of the prefix
// array, so the host prepares O(query terms) words instead of one record per item.
struct UnionJob {
    const uint32_t* gstart;      // ngroups+1: items before group g, in processing order
    const uint32_t* gterm;       // ngroups: index of the group's term in the batch's per-term arrays (max_weight order slot = gslot)
    const uint32_t* gquery;      // ngroups
    const uint32_t* gbase;       // ngroups: first result slot of the group (results are laid out in query order)
    const float* ub;             // per query term, in max_weight order: the reference's upper_bounds[] (queries.hpp:526-530)
    const uint32_t* gblocks;     // ngroups: blocks of the driving list per item of the group (fewer where a block is expensive to complete)
    uint32_t ngroups, nitems;
    uint32_t* work_counter;
    uint32_t* query_threshold;   // nq: float bits of the best published k-th score of the query
    uint32_t* item_sizes;
    float* item_scores;          // nitems * 2k: k scores, then the k docids they belong to
};

// top-k with a floor shared across the items of a query
struct TopKShared {
    TopK t;
    float floor_;     // published k-th best score of the whole query so far (0 = none)
    __device__ __forceinline__ void init(uint32_t k) { t.init(k); floor_ = 0.f; }
    __device__ __forceinline__ float bar() const { return t.size < t.k ? floor_ : fmaxf(t.thr, floor_); }
    __device__ __forceinline__ bool would_enter(float s) const { return s > bar(); }
    __device__ __forceinline__ void insert(float s, uint32_t docid) { if (would_enter(s)) t.insert(s, docid); }
};

struct UnionWarp {
    float qw[MAX_TERMS];
    float ub[MAX_TERMS];    // upper bounds, inflated by one part in 2^19: a bound that is itself a rounded sum stays a bound
    uint64_t bar;           // mbarrier of the probe staging window
    uint64_t dbar;          // mbarrier of the driving list's staging window
};

__host__ __device__ constexpr size_t union_warp_smem_bytes(int slots, bool pef = false) {
    return sizeof(UnionWarp) + size_t(slots) * sizeof(AndList) + BLOCK * 4 /* freqs */ + AND_STAGE_WINDOWS * STAGE_WORDS * 4 /* probe (+ driver) window */ +
           SCRATCH_WORDS * 4 + (pef ? size_t(slots) * sizeof(PefFreqSlot) + PEF_SCAN_SCRATCH_BYTES : 0);
}

// Look the candidates in `alive` up in list s (slot i).  score_mode: hits add the list's BM25 term; otherwise hits
// are cleared from `alive` (the document is owned by list i).  Candidates are sorted, the list cursor only moves forward.
template <int CODEC, class Ctx>
__device__ __forceinline__ void union_probe(Ctx& c, DevIndex const& idx, AndList* s, uint32_t i, const uint32_t (&cand)[4], uint32_t& alive,
                                            bool score_mode, float qwi, const float (&norm_len)[4], float (&score)[4], const uint32_t* ftmp) {
    const unsigned lane = lane_id();
    const uint2* bd = idx.bdir + s->bfirst;
    const uint32_t last_max = s->last_max;
    uint32_t pending = alive;
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
            const BlockMeta bm = and_find_block(c, bd, s->nblocks, fresh ? 0u : cur_block + 1, fresh ? 0xffffffffu : s->cur_max,
                                                fresh ? 0u : s->cur_end, cmin);
            and_decode_docs<CODEC>(c, idx, s, i, bm.block, bm.e0, bm.e1, bm.prev_max, bm.cur_max);
        }
        const uint32_t cur_max = s->cur_max;
        const uint32_t* d = s->docs;
        // the candidates this block answers
        uint32_t inb = 0, top = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if ((pending & (1u << j)) && cand[j] <= cur_max) { inb |= 1u << j; top = cand[j]; }
        pending &= ~inb;
        const uint32_t nin = __reduce_add_sync(FULL, __popc(inb));
        const uint4 v = reinterpret_cast<const uint4*>(d)[lane];
        uint32_t hitmask = 0, pos[4] = {0, 0, 0, 0};
        if (nin == 1) {
            // a sparse list probing a dense one: one candidate per block, found by comparing it with all 128 docids at once
            const uint32_t eq = (v.x == cmin ? 1u : 0u) | (v.y == cmin ? 2u : 0u) | (v.z == cmin ? 4u : 0u) | (v.w == cmin ? 8u : 0u);
            const unsigned hb = __ballot_sync(FULL, eq != 0u);
            if (hb) {
                const uint32_t hl = __ffs(hb) - 1;
                const uint32_t p = 4u * hl + __shfl_sync(FULL, uint32_t(__ffs(eq)) - 1u, hl);
                hitmask = inb;
                pos[0] = pos[1] = pos[2] = pos[3] = p;
            }
        } else {
            // a dense list probing a sparse one: usually none of the block's docids falls inside the candidates' range
            const uint32_t chi = __reduce_max_sync(FULL, top);
            const bool in = (v.x >= cmin && v.x <= chi) || (v.y >= cmin && v.y <= chi) || (v.z >= cmin && v.z <= chi) || (v.w >= cmin && v.w <= chi);
            if (!__any_sync(FULL, in)) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (inb & (1u << j)) {
                    pos[j] = lower_bound128(d, cand[j]);
                    if (d[pos[j]] == cand[j]) hitmask |= 1u << j;
                }
        }
        if (!score_mode) alive &= ~hitmask;
        else if (__any_sync(FULL, hitmask)) {
            const bool prefix = and_decode_freqs<CODEC>(c, idx, s, i, c.ftmp_off);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (hitmask & (1u << j)) {
                    const uint32_t p = pos[j];
                    const uint32_t f = prefix ? ftmp[p] - (p ? ftmp[p - 1] : 0u) : ftmp[p];
                    score[j] += qwi * doc_term_weight(f + 1u, norm_len[j]);
                }
            __syncwarp();
        }
    }
}

// MODE: what is done with the documents a list owns.
//   UNION_TOPK        wand_query / maxscore_query: dynamic pruning against the query's shared threshold (above)
//   UNION_EXHAUSTIVE  ranked_or_query (queries.hpp:404-476): every document of the union is scored in full, no pruning —
//                     the exhaustive cross-check of the pruned operators, on the same machinery
//   UNION_COUNT       or_query (queries.hpp:88-131): the size of the union = documents each list owns, summed; docids only
enum : int { UNION_TOPK = 0, UNION_EXHAUSTIVE = 1, UNION_COUNT = 2 };

template <int CODEC, int MIN_CTAS, bool STATS = true, int MODE = UNION_TOPK>
__global__ void __launch_bounds__(128, MIN_CTAS) union_drive_kernel(DevIndex idx, DevWand wand, DevBatch batch, UnionJob job, uint32_t k, int slots) {
    s16_table_init(smem_words(0));
    __syncthreads();

    const unsigned lane = lane_id();
    const unsigned warp = threadIdx.x >> 5;
    uint8_t* base = g_smem + S16_TAB_BYTES + warp * union_warp_smem_bytes(slots, CODEC == CODEC_PEF);
    UnionWarp* ws = reinterpret_cast<UnionWarp*>(base);
    AndList* st = reinterpret_cast<AndList*>(base + sizeof(UnionWarp));
    uint32_t* ftmp = reinterpret_cast<uint32_t*>(base + sizeof(UnionWarp) + size_t(slots) * sizeof(AndList));
    uint32_t* stage = ftmp + BLOCK;                 // probe window, then the driver window
    uint32_t* stack = stage + AND_STAGE_WINDOWS * STAGE_WORDS;

    AndCtxT<STATS> c;
    c.fcache_off = smem_offset(stack + SCRATCH_WORDS);
    c.drv_slot = 0xffffffffu; c.dwin_block = 0xffffffffu; c.dwin_delta = 0; c.dphase = 0;
    c.lists = idx.lists; c.stage = stage; c.bar = &ws->bar;
    c.stage_off = smem_offset(stage); c.stack_off = smem_offset(stack); c.ftmp_off = smem_offset(ftmp);
    c.phase = 0; c.win_slot = 0xffffffffu; c.win_delta = 0;
    c.c_docs_blocks = c.c_freqs_blocks = c.c_bytes_docs = c.c_bytes_freqs = c.c_maxs = c.c_scored = 0;
    if (lane == 0) { mbar_init(c.bar, 1); mbar_init(c.bar + 1, 1); fence_mbar_init(); }
    __syncwarp();
    constexpr float INFLATE = 1.0f + 1.0f / 524288.0f;

    while (true) {
        uint32_t ii = 0;
        if (lane == 0) ii = atomicAdd(job.work_counter, 1u);
        ii = __shfl_sync(FULL, ii, 0);
        if (ii >= job.nitems) break;
        const uint32_t g = warp_upper_group(job.gstart, job.ngroups, ii);
        const uint32_t chunk = ii - __ldg(job.gstart + g);
        const uint32_t q = __ldg(job.gquery + g);
        const uint32_t item_blocks = __ldg(job.gblocks + g);
        const uint32_t first_block = chunk * item_blocks;
        const uint32_t rslot = __ldg(job.gbase + g) + chunk;       // where this item's partial top-k goes
        const uint32_t t0 = batch.q_begin[q];
        const uint32_t nt = batch.q_begin[q + 1] - t0;
        const uint32_t e = __ldg(job.gterm + g) - t0;
        const volatile uint32_t* thr_g = job.query_threshold + q;
        // other warps raise the shared threshold concurrently: ONE lane reads it and broadcasts, so the value that steers
        // warp-uniform control flow (continue / stop / bar()) is the same in all 32 lanes
        auto read_threshold = [&]() -> float {
            uint32_t t = 0;
            if (lane == 0) t = *thr_g;
            return __uint_as_float(__shfl_sync(FULL, t, 0));
        };

        TopKShared topk;
        topk.init(k);
        if (MODE == UNION_TOPK) topk.floor_ = read_threshold();
        const float ub_e = __ldg(job.ub + t0 + e) * INFLATE;
        if (MODE == UNION_TOPK && !(ub_e > topk.floor_)) {           // list e is non-essential already: nothing it owns can enter
            if (lane == 0) job.item_sizes[rslot] = 0;
            continue;
        }
        uint32_t owned = 0;                    // UNION_COUNT: documents of this item that no higher list holds

        // slot i <- i-th list by increasing max_weight (queries.hpp:521-524, the reference's own std::sort order)
        c.win_slot = 0xffffffffu;
        __syncwarp();
        if (lane < nt) {
            const uint32_t src = batch.ord_maxw[t0 + lane];
            ws->qw[lane] = batch.q_weight[t0 + src];
            ws->ub[lane] = __ldg(job.ub + t0 + lane) * INFLATE;
            and_list_setup<CODEC>(idx, &st[lane], batch.term[t0 + src]);
            if (CODEC == CODEC_PEF) (reinterpret_cast<PefFreqSlot*>(g_smem + c.fcache_off + PEF_SCAN_SCRATCH_BYTES) + lane)->fp = 0xffffffffu;
        }
        __syncwarp();

        AndList* sd = &st[e];
        const uint2* bd0 = idx.bdir + sd->bfirst;
        const float qw_e = ws->qw[e];
        const uint32_t b_end = min(sd->nblocks, first_block + item_blocks);
        float published = topk.floor_;
        bool stop = false;
        for (uint32_t c0 = first_block; c0 < b_end && !stop; c0 += 32) {
            const uint32_t c1 = min(b_end, c0 + 32u);
            // directory entries of 32 blocks of the driving list, one block per lane, in one round trip
            uint32_t m_max = 0, m_end = 0, first_prev_max = 0xffffffffu, first_prev_end = 0;
            {
                const uint32_t bi = c0 + lane;
                if (bi < c1) { const uint2 en = __ldg(bd0 + bi); m_max = en.x; m_end = en.y; }
                if (c0) { const uint2 en = __ldg(bd0 + c0 - 1); first_prev_max = en.x; first_prev_end = en.y; }
            }
            for (uint32_t b0 = c0; b0 < c1; ++b0) {
                if (MODE == UNION_TOPK) {
                    // refresh the shared floor (queries.hpp:568-574: the non-essential prefix only grows)
                    topk.floor_ = fmaxf(topk.floor_, read_threshold());
                    if (!(ub_e > topk.bar())) { stop = true; break; }
                }
                {
                    const uint32_t l = b0 - c0;
                    const uint32_t pm = __shfl_sync(FULL, m_max, (l + 31) & 31), pe = __shfl_sync(FULL, m_end, (l + 31) & 31);
                    and_decode_docs<CODEC>(c, idx, sd, e, b0, l ? pe : first_prev_end, __shfl_sync(FULL, m_end, l), l ? pm : first_prev_max,
                                           __shfl_sync(FULL, m_max, l));
                }
                const uint4 cv = reinterpret_cast<const uint4*>(sd->docs)[lane];
                const uint32_t cand[4] = {cv.x, cv.y, cv.z, cv.w};
                uint32_t alive = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) alive |= (cand[j] != 0xffffffffu) << j;
                float norm_len[4] = {0.f, 0.f, 0.f, 0.f}, score[4] = {0.f, 0.f, 0.f, 0.f};

                // every list from the highest bound down.  Lists above e own the documents they share with e (a hit
                // drops the candidate); at e the survivors get their own term; lists below e complete the score
                // while score + ub[i] can still enter (queries.hpp:557-566)
                for (uint32_t i = nt; i-- > 0;) {
                    if (MODE == UNION_COUNT && i == e) break;           // the ownership probes are all a count needs
                    if (i == e) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (alive & (1u << j)) norm_len[j] = __ldg(wand.norm_lens + cand[j]);
                        const bool prefix = and_decode_freqs<CODEC>(c, idx, sd, e, c.ftmp_off);
                        const uint4 fv = reinterpret_cast<const uint4*>(ftmp)[lane];
                        uint32_t f0[4] = {fv.x, fv.y, fv.z, fv.w};
                        if (prefix) {
                            const uint32_t prev = lane ? ftmp[4 * lane - 1] : 0u;
                            f0[3] -= f0[2]; f0[2] -= f0[1]; f0[1] -= f0[0]; f0[0] -= prev;
                        }
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (alive & (1u << j)) score[j] = qw_e * doc_term_weight(f0[j] + 1u, norm_len[j]);
                        DS2I_STAT(c.c_scored += __reduce_add_sync(FULL, __popc(alive));)
                        continue;
                    }
                    if (MODE == UNION_TOPK && i < e) {
                        const float bar = topk.bar(), ubi = ws->ub[i];
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if ((alive & (1u << j)) && !(score[j] + ubi > bar)) alive &= ~(1u << j);
                        if (!__any_sync(FULL, alive)) break;
                    }
                    union_probe<CODEC>(c, idx, &st[i], i, cand, alive, i < e, ws->qw[i], norm_len, score, ftmp);
                    if (i > e && !__any_sync(FULL, alive)) break;
                }

                if (MODE == UNION_COUNT) { owned += __reduce_add_sync(FULL, __popc(alive)); continue; }
                // heap: only scores that can still enter
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    unsigned want = __ballot_sync(FULL, (alive & (1u << j)) && topk.would_enter(score[j]));
                    while (want) {
                        const int src = __ffs(want) - 1;
                        want &= want - 1;
                        topk.insert(__shfl_sync(FULL, score[j], src), __shfl_sync(FULL, cand[j], src));
                    }
                }
                if (MODE == UNION_TOPK && topk.t.size == topk.t.k && topk.t.thr > published) {
                    published = topk.t.thr;
                    if (lane == 0) atomicMax(job.query_threshold + q, __float_as_uint(published));
                }
            }
        }

        if (lane == 0) job.item_sizes[rslot] = MODE == UNION_COUNT ? owned : topk.t.size;
        if (MODE != UNION_COUNT && lane < topk.t.size) {
            job.item_scores[size_t(rslot) * 2 * k + lane] = topk.t.v;
            reinterpret_cast<uint32_t*>(job.item_scores)[size_t(rslot) * 2 * k + k + lane] = topk.t.id;
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

// fold the per-item partial top-k lists of each query (most items of a pruned list are empty: 32 sizes per load)
__global__ void __launch_bounds__(128) merge_union_items_kernel(const uint32_t* item_begin /* nq+1 */, uint32_t nq, const uint32_t* item_sizes,
                                                                const float* item_scores, uint32_t k, uint64_t* out_counts, float* out_scores,
                                                                uint32_t* out_docids) {
    const unsigned lane = lane_id();
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const uint32_t i0 = item_begin[q], i1 = item_begin[q + 1];
    TopK topk;
    topk.init(k);
    for (uint32_t base = i0; base < i1; base += 32) {
        const uint32_t it_l = base + lane;
        const uint32_t n_l = it_l < i1 ? item_sizes[it_l] : 0u;
        unsigned nz = __ballot_sync(FULL, n_l != 0u);
        while (nz) {
            const int src = __ffs(nz) - 1;
            nz &= nz - 1;
            const uint32_t n = __shfl_sync(FULL, n_l, src);
            const uint32_t it = base + src;
            const float v = lane < n ? item_scores[size_t(it) * 2 * k + lane] : 0.f;
            const uint32_t vid = lane < n ? reinterpret_cast<const uint32_t*>(item_scores)[size_t(it) * 2 * k + k + lane] : 0xffffffffu;
            for (uint32_t j = 0; j < n; ++j) {
                const float sc = __shfl_sync(FULL, v, j);
                if (!topk.would_enter(sc)) break;       // partial lists are sorted descending
                topk.insert(sc, __shfl_sync(FULL, vid, j));
            }
        }
    }
    if (lane == 0) out_counts[q] = uint64_t(topk.size);
    if (lane < k) {
        out_scores[size_t(q) * k + lane] = lane < topk.size ? topk.v : 0.f;
        out_docids[size_t(q) * k + lane] = lane < topk.size ? topk.id : 0xffffffffu;
    }
}

// or_query: the union's size is the sum of what each (list, block run) item owns
__global__ void __launch_bounds__(128) merge_union_counts_kernel(const uint32_t* item_begin /* nq+1 */, uint32_t nq, const uint32_t* item_sizes, uint64_t* out_counts) {
    const unsigned lane = lane_id();
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    unsigned long long total = 0;
    for (uint32_t it = item_begin[q] + lane; it < item_begin[q + 1]; it += 32) total += item_sizes[it];
    for (int d = 16; d; d >>= 1) total += __shfl_xor_sync(FULL, total, d);
    if (lane == 0) out_counts[q] = total;
}

}  // namespace ds2i_gpu
