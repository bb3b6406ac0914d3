// Warp-per-query evaluation of ds2i's query operators (queries.hpp) over the device enumerator.
// Persistent CTAs pull queries (longest first) from a global counter; each warp owns a private
// slice of shared memory: one ListState per query term, a TMA staging window, codec scratch.
// BM25 is accumulated in the reference's summation order and the TU is compiled with
// --fmad=false so scores are bit-identical to a -ffp-contract=off build of the reference.
#pragma once
#include "block_enum.cuh"

namespace ds2i_gpu {

constexpr int MAX_TERMS = 16;
constexpr int MAX_K = 32;

enum : int { OP_AND = 0, OP_AND_FREQ = 1, OP_OR = 2, OP_OR_FREQ = 3, OP_RANKED_AND = 4, OP_WAND = 5, OP_MAXSCORE = 6, OP_RANKED_OR = 7 };

struct DevWand {
    const float* norm_lens;         // wand_data::norm_len   (wand_data.hpp:55-58)
    const float* max_term_weight;   // wand_data::max_term_weight (:60-63)
    float min_norm_len;             // smallest norm_len of the collection: bm25::doc_term_weight(f, .) is largest there
};

struct DevBatch {
    uint32_t nq;
    const uint32_t* q_begin;     // nq+1 offsets into the per-term arrays (distinct terms of each query)
    const uint32_t* term;        // distinct term ids in increasing order = query_freqs order (queries.hpp:136-150)
    const float* q_weight;       // bm25::query_term_weight, computed on the host (bm25.hpp:17-24)
    const float* max_weight;     // q_weight * max_term_weight[term] (queries.hpp:231)
    const uint8_t* ord_size;     // per query: term positions sorted by list size  (queries.hpp:357-360)
    const uint8_t* ord_maxw;     // per query: term positions sorted by max_weight (queries.hpp:521-524)
    const uint32_t* sched;       // processing order of the queries (costliest first)
    uint32_t* work_counter;
    uint64_t* out_counts;        // nq
    float* out_scores;           // nq * k
    uint32_t* out_docids;        // nq * k: docid of every score
    unsigned long long* stats;   // 8 counters or nullptr
};

// bm25::doc_term_weight (bm25.hpp:11-15); fp32, no contraction (TU built with --fmad=false)
__device__ __forceinline__ float doc_term_weight(uint32_t freq, float norm_len) {
    float f = float(freq);
    return f / (f + 1.2f * (0.5f + 0.5f * norm_len));
}

// topk_queue (queries.hpp:152-197): the k largest scores, kept sorted descending across lanes.  The reference keeps
// scores only; here every entry also carries its docid (what a caller needs to fetch the documents, and what a
// doc-partitioned deployment merges on) — it rides along and never influences which scores are kept.
struct TopK {
    float v;          // lane i holds the i-th largest score
    uint32_t id;      // ... and the docid it belongs to
    uint32_t size, k;
    float thr;        // k-th largest, valid when size == k
    __device__ __forceinline__ void init(uint32_t k_) { v = 0.f; id = 0xffffffffu; size = 0; k = k_; thr = 0.f; }
    __device__ __forceinline__ bool would_enter(float s) const { return size < k || s > thr; }
    __device__ __forceinline__ bool insert(float s, uint32_t docid) {
        if (!would_enter(s)) return false;
        const unsigned lane = lane_id();
        unsigned ge = __ballot_sync(FULL, lane < size && v >= s);
        uint32_t pos = __popc(ge);
        float up = __shfl_up_sync(FULL, v, 1);
        uint32_t upid = __shfl_up_sync(FULL, id, 1);
        uint32_t nsize = size < k ? size + 1 : k;
        if (lane > pos && lane < nsize) { v = up; id = upid; }
        if (lane == pos) { v = s; id = docid; }
        size = nsize;
        thr = __shfl_sync(FULL, v, k - 1);
        return true;
    }
};

// ordered_enums as a register: 4 bits per position (MAX_TERMS == 16)
struct Order {
    uint64_t o;
    __device__ __forceinline__ uint32_t get(uint32_t i) const { return uint32_t(o >> (4 * i)) & 15u; }
    __device__ __forceinline__ void set(uint32_t i, uint32_t s) { o = (o & ~(uint64_t(15) << (4 * i))) | (uint64_t(s) << (4 * i)); }
    __device__ __forceinline__ void swap(uint32_t i, uint32_t j) { uint32_t a = get(i), b = get(j); set(i, b); set(j, a); }
};

struct WarpSmem {   // scalar per-warp arrays that sit next to the ListStates
    float qw[MAX_TERMS];
    float mw[MAX_TERMS];
    float ub[MAX_TERMS];
    uint64_t bar;
    uint64_t pad;
};

__host__ __device__ constexpr size_t warp_smem_bytes(int slots) {
    return sizeof(WarpSmem) + size_t(slots) * sizeof(ListState) + STAGE_WORDS * 4 + SCRATCH_WORDS * 4;
}

template <typename State>
__host__ __device__ constexpr size_t warp_smem_bytes_t(int slots) {
    return sizeof(WarpSmem) + size_t(slots) * sizeof(State) + STAGE_WORDS * 4 + SCRATCH_WORDS * 4;
}

// The reference's operators, literally: one warp per query, one candidate at a time, through an
// enumerator E (BlockEnum for the block indexes, PefEnum for `opt`).
template <class E, int OP>
__device__ __forceinline__ void run_queries(typename E::Index const& idx, DevWand wand, DevBatch batch, uint32_t k, int slots, int codec) {
    s16_table_init(smem_words(0));
    __syncthreads();

    typedef typename E::State State;
    const unsigned lane = lane_id();
    const unsigned warp = threadIdx.x >> 5;
    uint8_t* base = g_smem + S16_TAB_BYTES + warp * warp_smem_bytes_t<State>(slots);
    WarpSmem* ws = reinterpret_cast<WarpSmem*>(base);
    State* st = reinterpret_cast<State*>(base + sizeof(WarpSmem));
    uint32_t* stage = reinterpret_cast<uint32_t*>(base + sizeof(WarpSmem) + size_t(slots) * sizeof(State));
    uint32_t* scratch = stage + STAGE_WORDS;

    WarpCtx c;
    ctx_init(c, stage, scratch, &ws->bar, codec);
    const uint32_t N = idx.num_docs;
    constexpr bool RANKED = (OP == OP_RANKED_AND || OP == OP_WAND || OP == OP_MAXSCORE || OP == OP_RANKED_OR);

    while (true) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(batch.work_counter, 1u);
        qi = __shfl_sync(FULL, qi, 0);
        if (qi >= batch.nq) break;
        const uint32_t q = batch.sched[qi];
        const uint32_t t0 = batch.q_begin[q];
        const uint32_t nt = batch.q_begin[q + 1] - t0;
        uint64_t results = 0;
        TopK topk;
        topk.init(k);

        if (nt > 0) {
            // slot i <- i-th list in the operator's own order
            __syncwarp();
            if (lane < nt) {
                uint32_t src = lane;
                if (OP == OP_AND || OP == OP_AND_FREQ || OP == OP_RANKED_AND) src = batch.ord_size[t0 + lane];
                if (OP == OP_MAXSCORE) src = batch.ord_maxw[t0 + lane];
                if (RANKED) {
                    ws->qw[lane] = batch.q_weight[t0 + src];
                    ws->mw[lane] = batch.max_weight[t0 + src];
                }
                ws->ub[lane] = __uint_as_float(batch.term[t0 + src]);   // borrow ub[] to pass the term ids
            }
            __syncwarp();
            for (uint32_t i = 0; i < nt; ++i) E::open(c, idx, &st[i], __float_as_uint(ws->ub[i]));

            if (OP == OP_AND || OP == OP_AND_FREQ || OP == OP_RANKED_AND) {
                // and_query (queries.hpp:58-83) / ranked_and_query (:362-387)
                uint32_t candidate = E::docid(&st[0]);
                uint32_t i = 1;
                while (candidate < N) {
                    for (; i < nt; ++i) {
                        uint32_t d = E::next_geq(c, idx, &st[i], candidate);
                        if (d != candidate) { candidate = d; i = 0; break; }
                    }
                    if (i == nt) {
                        results += 1;
                        if (OP == OP_AND_FREQ) {
                            for (i = 0; i < nt; ++i) (void)E::freq(c, idx, &st[i]);
                        }
                        if (OP == OP_RANKED_AND) {
                            float norm_len = wand.norm_lens[candidate];
                            float score = 0.f;
                            for (i = 0; i < nt; ++i) score += ws->qw[i] * doc_term_weight(E::freq(c, idx, &st[i]), norm_len);
                            topk.insert(score, candidate);
                            c.c_scored += 1;
                        }
                        candidate = E::next(c, idx, &st[0]);
                        i = 1;
                    }
                }
            } else if (OP == OP_OR || OP == OP_OR_FREQ || OP == OP_RANKED_OR) {
                // or_query (queries.hpp:105-128) / ranked_or_query (:440-467)
                uint32_t cur_doc = N;
                for (uint32_t i = 0; i < nt; ++i) cur_doc = min(cur_doc, E::docid(&st[i]));
                while (cur_doc < N) {
                    results += 1;
                    float score = 0.f, norm_len = 0.f;
                    if (OP == OP_RANKED_OR) { norm_len = wand.norm_lens[cur_doc]; c.c_scored += 1; }
                    uint32_t next_doc = N;
                    for (uint32_t i = 0; i < nt; ++i) {
                        uint32_t d = E::docid(&st[i]);
                        if (d == cur_doc) {
                            if (OP == OP_OR_FREQ) (void)E::freq(c, idx, &st[i]);
                            if (OP == OP_RANKED_OR) score += ws->qw[i] * doc_term_weight(E::freq(c, idx, &st[i]), norm_len);
                            d = E::next(c, idx, &st[i]);
                        }
                        next_doc = min(next_doc, d);
                    }
                    if (OP == OP_RANKED_OR) topk.insert(score, cur_doc);
                    cur_doc = next_doc;
                }
            } else if (OP == OP_WAND) {
                // wand_query (queries.hpp:236-305).  std::sort on <= 16 pointers is libstdc++'s
                // insertion sort, i.e. stable: restated here on a 4-bit-per-entry register.
                Order ord; ord.o = 0xfedcba9876543210ull;
                auto sort_enums = [&]() {
                    for (uint32_t i = 1; i < nt; ++i) {
                        uint32_t s = ord.get(i), d = st[s].cur_docid;
                        uint32_t j = i;
                        while (j > 0 && d < st[ord.get(j - 1)].cur_docid) { ord.set(j, ord.get(j - 1)); --j; }
                        ord.set(j, s);
                    }
                };
                sort_enums();
                while (true) {
                    float upper_bound = 0.f;
                    uint32_t pivot;
                    bool found = false;
                    for (pivot = 0; pivot < nt; ++pivot) {
                        uint32_t s = ord.get(pivot);
                        if (st[s].cur_docid == N) break;
                        upper_bound += ws->mw[s];
                        if (topk.would_enter(upper_bound)) { found = true; break; }
                    }
                    if (!found) break;
                    uint32_t pivot_id = st[ord.get(pivot)].cur_docid;
                    if (pivot_id == st[ord.get(0)].cur_docid) {
                        float score = 0.f;
                        float norm_len = wand.norm_lens[pivot_id];
                        for (uint32_t p = 0; p < nt; ++p) {
                            uint32_t s = ord.get(p);
                            if (st[s].cur_docid != pivot_id) break;
                            score += ws->qw[s] * doc_term_weight(E::freq(c, idx, &st[s]), norm_len);
                            E::next(c, idx, &st[s]);
                        }
                        topk.insert(score, pivot_id);
                        c.c_scored += 1;
                        sort_enums();
                    } else {
                        uint32_t next_list = pivot;
                        for (; st[ord.get(next_list)].cur_docid == pivot_id; --next_list) {}
                        E::next_geq(c, idx, &st[ord.get(next_list)], pivot_id);
                        for (uint32_t i = next_list + 1; i < nt; ++i) {
                            if (st[ord.get(i)].cur_docid < st[ord.get(i - 1)].cur_docid) ord.swap(i, i - 1);
                            else break;
                        }
                    }
                }
            } else if (OP == OP_MAXSCORE) {
                // maxscore_query (queries.hpp:519-578); slots are already in increasing max_weight order
                __syncwarp();
                if (lane == 0) {
                    float acc = ws->mw[0];
                    ws->ub[0] = acc;
                    for (uint32_t i = 1; i < nt; ++i) { acc = acc + ws->mw[i]; ws->ub[i] = acc; }
                }
                __syncwarp();
                uint32_t non_essential = 0;
                uint32_t cur_doc = N;
                for (uint32_t i = 0; i < nt; ++i) cur_doc = min(cur_doc, E::docid(&st[i]));
                while (non_essential < nt && cur_doc < N) {
                    float score = 0.f;
                    float norm_len = wand.norm_lens[cur_doc];
                    uint32_t next_doc = N;
                    for (uint32_t i = non_essential; i < nt; ++i) {
                        uint32_t d = E::docid(&st[i]);
                        if (d == cur_doc) {
                            score += ws->qw[i] * doc_term_weight(E::freq(c, idx, &st[i]), norm_len);
                            d = E::next(c, idx, &st[i]);
                        }
                        next_doc = min(next_doc, d);
                    }
                    for (uint32_t i = non_essential - 1; i + 1 > 0; --i) {
                        if (!topk.would_enter(score + ws->ub[i])) break;
                        uint32_t d = E::next_geq(c, idx, &st[i], cur_doc);
                        if (d == cur_doc) score += ws->qw[i] * doc_term_weight(E::freq(c, idx, &st[i]), norm_len);
                    }
                    c.c_scored += 1;
                    if (topk.insert(score, cur_doc)) {
                        while (non_essential < nt && !topk.would_enter(ws->ub[non_essential])) non_essential += 1;
                    }
                    cur_doc = next_doc;
                }
            }
        }

        if (lane == 0) batch.out_counts[q] = RANKED ? uint64_t(topk.size) : results;
        if (RANKED && lane < k) {
            batch.out_scores[size_t(q) * k + lane] = lane < topk.size ? topk.v : 0.f;
            batch.out_docids[size_t(q) * k + lane] = lane < topk.size ? topk.id : 0xffffffffu;
        }
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

template <int CODEC, int OP>
__global__ void __launch_bounds__(128) query_kernel(DevIndex idx, DevWand wand, DevBatch batch, uint32_t k, int slots) {
    run_queries<BlockEnum<CODEC>, OP>(idx, wand, batch, k, slots, idx.codec);
}

}  // namespace ds2i_gpu
