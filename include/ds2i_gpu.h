/*
 * ds2i_gpu.h — C ABI of the B200-native ds2i query path (libds2i_gpu.so).
 *
 * ds2i has no FFI layer of its own: its boundary is C++ template duck-typing (Index /
 * document_enumerator / QueryOperator concepts) plus the `queries` CLI contract.  Each entry point
 * below names the reference interface it stands in for (file:line relative to the ds2i tree) so a
 * maintainer can bind it from `queries.cpp`-style code (the C++ adapters that
 * model the reference concepts on top of this ABI are in include/ds2i_gpu.hpp).
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * DS2I_E_* code (ds2i_gpu_last_error() gives the message for the calling thread); the caller owns
 * all host buffers, the library owns all device memory; handles are thread-compatible (one call in
 * flight per handle).  Index and wand-data bytes are ds2i's own on-disk format
 * (succinct::mapper::freeze, succinct/mapper.hpp:51-98) — nothing is re-encoded.
 */
#ifndef DS2I_GPU_H_
#define DS2I_GPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ds2i_gpu_index ds2i_gpu_index;   /* an index resident in HBM  */
typedef struct ds2i_gpu_wand  ds2i_gpu_wand;    /* wand_data resident in HBM */
typedef struct ds2i_gpu_batch ds2i_gpu_batch;   /* a prepared query batch resident in HBM */

enum {
    DS2I_OK = 0,
    DS2I_E_ARG = -1,       /* invalid argument (std::invalid_argument in the reference)     */
    DS2I_E_FORMAT = -2,    /* not a ds2i index / wand file of the stated type               */
    DS2I_E_CUDA = -3,      /* CUDA runtime error (no CPU fallback exists: the call fails)   */
    DS2I_E_UNSUPPORTED = -4,
    DS2I_E_LIMIT = -5      /* a compile-time limit was exceeded (terms per query, k)        */
};

/* Query operators of queries.hpp, in the spelling queries.cpp:104-120 parses from argv. */
enum {
    DS2I_OP_AND = 0,        /* and_query<false>    queries.hpp:35-86   */
    DS2I_OP_AND_FREQ = 1,   /* and_query<true>     queries.hpp:35-86   */
    DS2I_OP_OR = 2,         /* or_query<false>     queries.hpp:88-131  */
    DS2I_OP_OR_FREQ = 3,    /* or_query<true>      queries.hpp:88-131  */
    DS2I_OP_RANKED_AND = 4, /* ranked_and_query    queries.hpp:322-401 */
    DS2I_OP_WAND = 5,       /* wand_query          queries.hpp:200-319 */
    DS2I_OP_MAXSCORE = 6,   /* maxscore_query      queries.hpp:478-591 */
    DS2I_OP_RANKED_OR = 7   /* ranked_or_query     queries.hpp:404-476 */
};

#define DS2I_GPU_MAX_TERMS 16   /* distinct terms per query handled on device */
#define DS2I_GPU_MAX_K     32   /* top-k capacity (queries.cpp:113-117 hard-wires k = 10) */

const char* ds2i_gpu_last_error(void);
int ds2i_gpu_op_from_name(const char* name);          /* "ranked_and" -> DS2I_OP_RANKED_AND; <0 if unknown */
int ds2i_gpu_index_type_known(const char* index_type); /* 1 for the entries of DS2I_INDEX_TYPES (index_types.hpp:41), else 0 */

/* ---- Index: replaces succinct::mapper::map(index, mapped_file) + the Index concept ------------
 * (queries.cpp:73-77; block_freq_index.hpp:73-134; freq_index.hpp:106-243).
 * index_type is the reference's type name, every entry of DS2I_INDEX_TYPES (index_types.hpp:41):
 * ef, single, uniform, opt, block_optpfor, block_varint, block_interpolative, block_mixed, block_qmx.
 * The compressed bytes are parsed on the host and copied ONCE into HBM of `device`. */
int ds2i_gpu_index_open(const void* file_bytes, size_t nbytes, const char* index_type, int device,
                        ds2i_gpu_index** out);
int ds2i_gpu_index_open_file(const char* path, const char* index_type, int device, ds2i_gpu_index** out);
void ds2i_gpu_index_close(ds2i_gpu_index*);
uint64_t ds2i_gpu_index_size(const ds2i_gpu_index*);       /* Index::size()      number of posting lists */
uint64_t ds2i_gpu_index_num_docs(const ds2i_gpu_index*);   /* Index::num_docs()                          */
uint64_t ds2i_gpu_index_device_bytes(const ds2i_gpu_index*);
/* document_enumerator::size() of index[term] for each term (block_posting_list.hpp:178-181). */
/* Document-partitioned deployment: the index holds the documents of ONE shard.  BM25 needs the
 * statistics of the whole collection — bm25::query_term_weight(qtf, df, num_docs) (bm25.hpp:17-24) and the order in
 * which ranked_and_query sums its terms (lists by increasing size, queries.hpp:357-360) — so the caller supplies, for
 * every list of this index, the document frequency over all shards, and the total number of documents.  With them
 * every score equals the one the reference computes on the unsharded collection.  df == NULL clears them. */
int ds2i_gpu_index_set_global_stats(ds2i_gpu_index*, const uint64_t* df, size_t nterms, uint64_t num_docs_total);
int ds2i_gpu_index_list_sizes(const ds2i_gpu_index*, const uint32_t* terms, size_t nterms, uint64_t* out_sizes);
/* Compressed bytes of index[term] in the file: block_maxs + block_endpoints + block data for the block indexes
 * (block_posting_list.hpp:14-53), the list's bits in the docs and freqs bit vectors for the Elias-Fano family
 * (bitvector_collection.hpp:57-67) — the algorithmic bytes of a full scan of the list. */
int ds2i_gpu_index_list_bytes(const ds2i_gpu_index*, const uint32_t* terms, size_t nterms, uint64_t* out_bytes);

/* ---- wand_data: replaces succinct::mapper::map(wdata, md) (queries.cpp:90-95; wand_data.hpp:55-78). */
int ds2i_gpu_wand_open(const void* file_bytes, size_t nbytes, int device, ds2i_gpu_wand** out);
int ds2i_gpu_wand_open_file(const char* path, int device, ds2i_gpu_wand** out);
void ds2i_gpu_wand_close(ds2i_gpu_wand*);

/* ---- Query operators: replaces the hot loop `query_op(index, query)` of op_perftest
 * (queries.cpp:25-35) for a whole batch.  terms[query_offsets[q] .. query_offsets[q+1]) are the raw
 * term ids of query q exactly as read_query (queries.hpp:15-27) returns them (duplicates allowed).
 *   out_counts[q]         the operator's return value (match count for and/or, topk().size() else)
 *   out_scores[q*k .. +k) topk() of the ranked operators, descending, zero padded (may be NULL)
 *   out_elapsed_ms        CUDA-event time of the device work of this call (may be NULL)
 * wand may be NULL for and/or.  Host buffers in, host buffers out (H2D and D2H inside the call). */
int ds2i_gpu_query_batch(ds2i_gpu_index*, ds2i_gpu_wand*, int op, uint32_t k,
                         const uint32_t* terms, const uint64_t* query_offsets, size_t nq,
                         uint64_t* out_counts, float* out_scores, float* out_elapsed_ms);
/* the same with the docid of every score (nq * k, 0xffffffff padding; out_docids may be NULL) */
int ds2i_gpu_query_batch_docids(ds2i_gpu_index*, ds2i_gpu_wand*, int op, uint32_t k,
                                const uint32_t* terms, const uint64_t* query_offsets, size_t nq,
                                uint64_t* out_counts, float* out_scores, uint32_t* out_docids, float* out_elapsed_ms);

/* The same in three steps, for callers that keep a batch resident in HBM and run several
 * operators over it (what op_perftest does with its `queries` vector, queries.cpp:97-121). */
int ds2i_gpu_batch_prepare(ds2i_gpu_index*, ds2i_gpu_wand*, const uint32_t* terms,
                           const uint64_t* query_offsets, size_t nq, ds2i_gpu_batch** out);
int ds2i_gpu_batch_run(ds2i_gpu_batch*, int op, uint32_t k, float* out_elapsed_ms);   /* device only, synchronises */
/* flags: DS2I_RUN_FAITHFUL evaluates with the reference's own control flow, one warp per query, one
 * candidate at a time (same blocks decoded as the reference; the slow, literal kernels; scores
 * bit-identical to a -ffp-contract=off build of the reference).  Without it and / ranked_and use the
 * block-at-a-time kernels (same results, bit for bit) and wand / maxscore use the block-parallel
 * dynamic-pruning kernel (same top-k; a score may differ from the reference in the last bit because
 * the BM25 summation order follows the pruning state — within 1e-5 relative). */
#define DS2I_RUN_FAITHFUL 1u
/* DS2I_RUN_ASYNC: launch and return without waiting; the work is ordered on the device's default stream (stream 0), so a
 * caller may enqueue a collective or a copy behind it, or call ds2i_gpu_batch_wait (which also reports the CUDA-event time).
 * The fetch / stats calls copy on the same stream and therefore see the finished results either way. */
#define DS2I_RUN_ASYNC 2u
/* DS2I_RUN_NO_STATS: run the kernel instance without the algorithmic-work counters of ds2i_gpu_batch_stats (they then read 0).
 * The counters cost registers in register-bound kernels; a benchmark times launches without them and collects them from one
 * more launch of the same batch.  Instances exist for block_optpfor and the Elias-Fano family; other types ignore the flag. */
#define DS2I_RUN_NO_STATS 4u
int ds2i_gpu_batch_run_ex(ds2i_gpu_batch*, int op, uint32_t k, uint32_t flags, float* out_elapsed_ms);
int ds2i_gpu_batch_wait(ds2i_gpu_batch*, float* out_elapsed_ms);
int ds2i_gpu_batch_fetch(ds2i_gpu_batch*, uint64_t* out_counts, float* out_scores);   /* D2H of the last run */
/* docids of the scores of the last (ranked) run, nq * k, same layout as out_scores, 0xffffffff where a query has
 * fewer than k results.  The reference's topk_queue keeps scores only (queries.hpp:157-172); a caller that has to
 * fetch the documents — or merge the results of document-partitioned shards — needs the ids. */
int ds2i_gpu_batch_fetch_docids(ds2i_gpu_batch*, uint32_t* out_docids);
/* Device-side counters of the last run: [0] docs blocks decoded, [1] freqs blocks decoded,
 * [2] compressed docs bytes, [3] compressed freqs bytes, [4] block_max entries read,
 * [5] documents scored, [6] kernel launches.  These are the algorithmic bytes of SURVEY.md §8(d). */
int ds2i_gpu_batch_stats(ds2i_gpu_batch*, uint64_t out_stats[8]);
/* Device addresses of the last run's results (counts: nq u64; scores: nq*k f32), for callers that
 * hand them to a collective (NCCL gather of per-shard top-k) without a host round trip. */
int ds2i_gpu_batch_device_results(ds2i_gpu_batch*, void** d_counts, void** d_scores);
int ds2i_gpu_batch_device_docids(ds2i_gpu_batch*, void** d_docids);      /* device pointer to the nq * k docids of the last ranked run */
/* The three result arrays of the last run are ONE contiguous device buffer, [counts: nq u64][scores: nq*k f32][docids: nq*k u32]
 * (unranked operators: counts only), so that a single collective or copy moves a shard's results. */
int ds2i_gpu_batch_device_fused(ds2i_gpu_batch*, void** d_fused, size_t* bytes);

/* Document-partitioned shards.  Every shard evaluated the same nq queries; row (s * nq + q) of the
 * DEVICE arrays d_counts / d_scores / d_docids (docids already global) holds shard s's result for query q — the layout an
 * NCCL all_gather of the per-shard results produces.  Shards hold disjoint documents: unranked counts add up, the
 * global top-k is the top-k of the shards' lists.  Outputs are device arrays (nq, nq * k, nq * k). */
int ds2i_gpu_merge_shards(const uint64_t* d_counts, const float* d_scores, const uint32_t* d_docids, uint32_t nshards, size_t nq,
                          uint32_t k, int ranked, uint64_t* d_out_counts, float* d_out_scores, uint32_t* d_out_docids);
void ds2i_gpu_batch_free(ds2i_gpu_batch*);

/* ---- Several GPUs of one process -------------------------------------------------------------------
 * The reference's only parallel driver deals query i to thread i % n, one operator copy per thread, index shared
 * (profile_queries.cpp:21-39).  Here a group holds one replica of the index (and wand data) per device; a batch is cut into
 * cost-balanced shards (queries sorted by the postings of their lists, dealt round-robin), every GPU evaluates its shard
 * concurrently (one host thread per device), the fused per-shard results are gathered on the first device over NCCL
 * (ncclCommInitAll communicators, ncclSend / ncclRecv in one group; libnccl is dlopen'ed on first use) and reach the caller's
 * host buffers in one D2H copy, in the caller's query order.  devices == NULL means 0 .. ndevices-1; wand_path may be NULL
 * for and / or.  out_elapsed_ms is the largest per-GPU CUDA-event time. */
typedef struct ds2i_gpu_group ds2i_gpu_group;
int ds2i_gpu_group_open(const char* index_path, const char* index_type, const char* wand_path, const int* devices, int ndevices,
                        ds2i_gpu_group** out);
void ds2i_gpu_group_close(ds2i_gpu_group*);
int ds2i_gpu_group_size(const ds2i_gpu_group*);
ds2i_gpu_index* ds2i_gpu_group_index(ds2i_gpu_group*, int i);            /* the replica on the i-th device (owned by the group) */
int ds2i_gpu_group_query_batch(ds2i_gpu_group*, int op, uint32_t k, const uint32_t* terms, const uint64_t* query_offsets, size_t nq,
                               uint64_t* out_counts, float* out_scores, uint32_t* out_docids, float* out_elapsed_ms);

/* ---- Enumerator-level entry points (document_enumerator, block_posting_list.hpp:105-186) -------
 * Full sequential decode of whole lists: for each term, docid()/freq() of every posting as
 * next() would deliver them.  out_offsets[i] (nterms+1 entries, in postings) says where list i
 * goes in out_docs/out_freqs; it must equal the prefix sums of ds2i_gpu_index_list_sizes. */
int ds2i_gpu_decode_lists(ds2i_gpu_index*, const uint32_t* terms, size_t nterms,
                          const uint64_t* out_offsets, uint32_t* out_docs, uint32_t* out_freqs,
                          float* out_elapsed_ms);
/* The same decode with the outputs left in HBM and reduced there: sum of every docid() and of every freq() over all postings
 * of the listed terms — the checksum a full-scale decode (hundreds of millions of postings) is compared on, next to a
 * bit-exact comparison of sampled lists through ds2i_gpu_decode_lists. */
int ds2i_gpu_decode_lists_checksum(ds2i_gpu_index*, const uint32_t* terms, size_t nterms, const uint64_t* out_offsets,
                                   uint64_t* out_sum_docids, uint64_t* out_sum_freqs, float* out_elapsed_ms);
/* next_geq sweeps: list i = index[terms[i]] is opened (positioned on its first posting), then
 * next_geq(bounds[j]) is applied for j in bound_offsets[i]..bound_offsets[i+1] (non-decreasing
 * bounds); out_docids[j] = docid() after the call, out_freqs[j] = freq() (0 past the end). */
int ds2i_gpu_next_geq_batch(ds2i_gpu_index*, const uint32_t* terms, size_t nlists,
                            const uint64_t* bounds, const uint64_t* bound_offsets,
                            uint64_t* out_docids, uint64_t* out_freqs, float* out_elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* DS2I_GPU_H_ */
