// Partitioned Elias-Fano (`opt` index) path: host-side loader that flattens the per-list partition
// headers into a directory, the device enumerator with ds2i's next / next_geq / docid / freq
// semantics (freq_index.hpp:116-190, partitioned_sequence.hpp:122-347, positive_sequence.hpp:33-78),
// and the kernels behind decode_lists / next_geq_batch / the query operators for this index type.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "format.hpp"
#include "pef.cuh"
#include "query_kernels.cuh"

namespace ds2i_gpu {

// ------------------------------------------------------------------------------------------------
// device enumerator

struct PefState {
    // docs sequence
    uint64_t d_first_part;
    uint32_t d_nparts, n;
    uint32_t cur_part;          // index of the current docs partition inside the list
    uint32_t chunk_i0;          // local index (inside the partition) of docs[0]
    uint32_t chunk_size;
    uint32_t pos;               // position inside the chunk
    uint32_t cur_docid;
    uint32_t part_begin;        // list position of the partition's first element
    // freqs sequence
    uint64_t f_first_part;
    uint32_t f_nparts;
    uint32_t f_cur_part;
    uint32_t f_g0;              // list position of fvals[0]
    uint32_t f_cnt;             // valid entries in fvals (0 = nothing cached)
    uint32_t f_prev0;           // prefix-sum value just before fvals[0] (valid when f_has_prev)
    uint32_t f_has_prev;
    uint32_t pad[2];
    uint32_t docs[128];         // docids of the current chunk (0xffffffff beyond chunk_size)
    uint32_t fvals[128];        // prefix sums of freqs for list positions f_g0 ..
};
static_assert(sizeof(PefState) % 16 == 0, "PefState must keep 16-byte alignment of its arrays");

struct PefEnum {
    typedef PefState State;
    typedef PefIndexDev Index;

    static __device__ __forceinline__ PefPart load_part(const PefPart* parts, uint64_t i) { return pef_load_part(parts, i); }

    // decode the chunk of partition `part` that starts at local index i0
    static __device__ __forceinline__ void load_chunk(Index const& idx, State* st, uint32_t part, uint32_t i0) {
        PefPart p = load_part(idx.docs.parts, st->d_first_part + part);
        PefBody b = pef_open_body(idx.docs, p, false);
        uint32_t cnt = min(128u, p.size - i0);
        pef_decode_range(idx.docs, p, b, i0, cnt, st->docs);
        const unsigned lane = lane_id();
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) { uint32_t e = lane + 32u * j; if (e >= cnt) st->docs[e] = 0xffffffffu; }
        __syncwarp();
        if (lane == 0) {
            st->cur_part = part; st->chunk_i0 = i0; st->chunk_size = cnt; st->pos = 0; st->part_begin = p.begin;
            st->cur_docid = st->docs[0];
        }
        __syncwarp();
    }

    static __device__ __forceinline__ void open(WarpCtx&, Index const& idx, State* st, uint32_t term) {
        if (lane_id() == 0) {
            PefListDir d = idx.docs.lists[term], f = idx.freqs.lists[term];
            st->d_first_part = d.first_part; st->d_nparts = d.nparts; st->n = d.n;
            st->f_first_part = f.first_part; st->f_nparts = f.nparts; st->f_cur_part = 0; st->f_cnt = 0; st->f_g0 = 0;
            st->f_has_prev = 0; st->f_prev0 = 0;
        }
        __syncwarp();
        load_chunk(idx, st, 0, 0);
    }

    static __device__ __forceinline__ uint32_t docid(const State* st) { return st->cur_docid; }
    static __device__ __forceinline__ uint32_t size(const State* st) { return st->n; }
    static __device__ __forceinline__ uint64_t position(const State* st) { return uint64_t(st->part_begin) + st->chunk_i0 + st->pos; }

    static __device__ __forceinline__ uint32_t set_end(Index const& idx, State* st) {
        __syncwarp();
        if (lane_id() == 0) { st->cur_docid = idx.num_docs; st->pos = st->chunk_size; }
        __syncwarp();
        return idx.num_docs;
    }

    static __device__ __forceinline__ uint32_t next(WarpCtx&, Index const& idx, State* st) {
        uint32_t pos = st->pos + 1;
        if (pos < st->chunk_size) {
            uint32_t d = st->docs[pos];
            __syncwarp();
            if (lane_id() == 0) { st->pos = pos; st->cur_docid = d; }
            __syncwarp();
            return d;
        }
        // next chunk of the partition, or the next partition, or the end (partitioned_sequence.hpp:236-250)
        PefPart p = load_part(idx.docs.parts, st->d_first_part + st->cur_part);
        uint32_t ni = st->chunk_i0 + st->chunk_size;
        if (ni < p.size) { load_chunk(idx, st, st->cur_part, ni); return st->cur_docid; }
        if (st->cur_part + 1 < st->d_nparts) { load_chunk(idx, st, st->cur_part + 1, 0); return st->cur_docid; }
        return set_end(idx, st);
    }

    // first posting with docid >= lower_bound at or after the cursor (the block-list semantics; the
    // reference's EF enumerators agree with it for non-decreasing bounds)
    static __device__ __forceinline__ uint32_t next_geq(WarpCtx&, Index const& idx, State* st, uint32_t lower_bound) {
        const unsigned lane = lane_id();
        const uint32_t cur = st->cur_docid;
        if (cur == idx.num_docs || cur >= lower_bound) return cur;
        uint32_t chunk_last = st->docs[st->chunk_size - 1];
        if (lower_bound > chunk_last) {
            uint32_t part = st->cur_part;
            PefPart p = load_part(idx.docs.parts, st->d_first_part + part);
            if (lower_bound > p.ub) {
                // partition search on the flattened upper bounds (the reference: m_upper_bounds.next_geq, :286)
                const uint32_t nparts = st->d_nparts;
                uint32_t lo = part + 1;
                bool found = false;
                while (lo < nparts) {
                    uint32_t pi = lo + lane;
                    uint32_t ub = 0;
                    if (pi < nparts) ub = uint32_t(__ldg(reinterpret_cast<const uint64_t*>(idx.docs.parts + st->d_first_part + pi) + 2) >> 32);
                    unsigned hit = __ballot_sync(FULL, pi < nparts && ub >= lower_bound);
                    if (hit) { lo += __ffs(hit) - 1; found = true; break; }
                    lo += 32;
                }
                if (!found) return set_end(idx, st);
                part = lo;
                p = load_part(idx.docs.parts, st->d_first_part + part);
                if (lower_bound <= p.base) { load_chunk(idx, st, part, 0); return st->cur_docid; }
            }
            // inside partition `part`: locate the chunk holding the first value >= lower_bound
            PefBody b = pef_open_body(idx.docs, p, false);
            uint32_t from = (part == st->cur_part) ? st->chunk_i0 + st->chunk_size : 0u;
            uint32_t i0 = from;
            if (i0 >= p.size) {
                // only single / ef indexes get here (their one "partition" is bounded by the universe, not by its last
                // value): the cursor's chunk was the last one and the bound lies behind it
                if (part + 1 < st->d_nparts) { load_chunk(idx, st, part + 1, 0); return st->cur_docid; }
                return set_end(idx, st);
            }
            if (p.size - from > 128u) {
                uint32_t hint = pef_rank_hint(idx.docs, b, lower_bound - p.base);
                if (hint > i0) i0 = hint;
                if (i0 >= p.size) i0 = p.size - 1;      // lower_bound <= ub: the last element qualifies
            }
            while (true) {
                load_chunk(idx, st, part, i0);
                if (st->docs[st->chunk_size - 1] >= lower_bound) break;
                i0 += st->chunk_size;                   // (EF: elements sharing the high part of the bound may span chunks)
                if (i0 >= p.size) {
                    // single / ef indexes: the one "partition" is bounded by the universe, not by its last value, so
                    // the bound may lie behind every element
                    if (part + 1 < st->d_nparts) { load_chunk(idx, st, part + 1, 0); return st->cur_docid; }
                    return set_end(idx, st);
                }
            }
        }
        uint4 v = reinterpret_cast<const uint4*>(st->docs)[lane];
        uint32_t cnt = (v.x < lower_bound) + (v.y < lower_bound) + (v.z < lower_bound) + (v.w < lower_bound);
        uint32_t pos = __reduce_add_sync(FULL, cnt);
        uint32_t cp = st->pos;
        if (pos < cp) pos = cp;
        uint32_t d = st->docs[pos & 127u];
        __syncwarp();
        if (lane == 0) { st->pos = pos; st->cur_docid = d; }
        __syncwarp();
        return d;
    }

    // prefix sums of the freqs sequence for list positions starting at max(g - 1, partition begin)
    static __device__ __forceinline__ void load_freq_chunk(Index const& idx, State* st, uint32_t g) {
        const unsigned lane = lane_id();
        // partition of the freqs sequence that holds position g
        uint32_t fp = st->f_cur_part;
        const uint32_t nparts = st->f_nparts;
        PefPart p = load_part(idx.freqs.parts, st->f_first_part + fp);
        if (g < p.begin) { fp = 0; p = load_part(idx.freqs.parts, st->f_first_part); }
        while (g >= p.begin + p.size) {
            // forward scan, 32 partitions per step
            uint32_t pi = fp + 1 + lane;
            uint32_t end = 0;
            if (pi < nparts) {
                uint64_t a = __ldg(reinterpret_cast<const uint64_t*>(idx.freqs.parts + st->f_first_part + pi) + 1);
                end = uint32_t(a) + uint32_t(a >> 32);
            }
            unsigned hit = __ballot_sync(FULL, pi < nparts && g < end);
            if (hit) fp = fp + 1 + (__ffs(hit) - 1); else fp += 32;
            if (fp >= nparts) fp = nparts - 1;
            p = load_part(idx.freqs.parts, st->f_first_part + fp);
            if (hit) break;
        }
        PefBody b = pef_open_body(idx.freqs, p, true);
        uint32_t local = g - p.begin;
        uint32_t ls = local ? local - 1 : 0;
        uint32_t cnt = min(128u, p.size - ls);
        pef_decode_range(idx.freqs, p, b, ls, cnt, st->fvals);
        if (lane == 0) {
            st->f_cur_part = fp; st->f_g0 = p.begin + ls; st->f_cnt = cnt;
            st->f_has_prev = (ls == 0) ? 1u : 0u;
            // the value before a partition's first element is the previous partition's last value = base - 1
            // (partition 0: base is the first value itself and the sum before it is 0)
            st->f_prev0 = (ls == 0 && fp) ? p.base - 1u : 0u;
        }
        __syncwarp();
    }

    // positive_sequence::enumerator::move(position).second (positive_sequence.hpp:48-66)
    static __device__ __forceinline__ uint32_t freq(WarpCtx&, Index const& idx, State* st) {
        const uint32_t g = st->part_begin + st->chunk_i0 + st->pos;
        bool ok = st->f_cnt && g >= st->f_g0 && g < st->f_g0 + st->f_cnt && (g > st->f_g0 || st->f_has_prev);
        if (!ok) load_freq_chunk(idx, st, g);
        uint32_t j = g - st->f_g0;
        uint32_t cur = st->fvals[j];
        uint32_t prev = j ? st->fvals[j - 1] : st->f_prev0;
        return cur - prev;
    }
};

// ------------------------------------------------------------------------------------------------
// kernels

// the reference's operators, literally, over the PEF enumerator (same control flow as query_kernel)
template <int OP>
__global__ void __launch_bounds__(128) pef_query_kernel(PefIndexDev idx, DevWand wand, DevBatch batch, uint32_t k, int slots) {
    run_queries<PefEnum, OP>(idx, wand, batch, k, slots, 0);
}

struct PefDecodeItem { uint32_t list_pos, part /* bit 31: freqs sequence */, first, count; };   // elements [first, first+count) of one partition

struct PefDecodeJob {
    const uint32_t* terms;
    const PefDecodeItem* items;
    const uint64_t* out_offsets;
    uint32_t* out_docs;
    uint32_t* out_freqs;
    uint32_t nitems;
};

// full decode: one warp per item = up to PEF_ITEM_ELEMS consecutive elements of one partition (long
// partitions — e.g. the single-partition freqs sequence of a long list — are cut so that no warp walks
// thousands of chunks alone)
constexpr uint32_t PEF_ITEM_ELEMS = 2048;

static __global__ void __launch_bounds__(256) pef_decode_kernel(PefIndexDev idx, PefDecodeJob job) {
    uint32_t* buf = smem_words((threadIdx.x >> 5) * 1024);
    const unsigned lane = lane_id();
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < job.nitems; g += nwarps) {
        const PefDecodeItem it = job.items[g];
        const uint32_t term = job.terms[it.list_pos];
        const uint64_t o = job.out_offsets[it.list_pos];
        const bool is_freq = it.part >> 31;
        const uint32_t pi = it.part & 0x7fffffffu;
        if (!is_freq) {
            const PefListDir d = idx.docs.lists[term];
            PefPart p = PefEnum::load_part(idx.docs.parts, d.first_part + pi);
            PefBody b = pef_open_body(idx.docs, p, false);
            for (uint32_t i0 = it.first; i0 < it.first + it.count; i0 += 128) {
                uint32_t cnt = min(128u, it.first + it.count - i0);
                pef_decode_range(idx.docs, p, b, i0, cnt, buf);
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) { uint32_t e = lane + 32u * j; if (e < cnt) job.out_docs[o + p.begin + i0 + e] = buf[e]; }
                __syncwarp();
            }
        } else {
            const PefListDir f = idx.freqs.lists[term];
            PefPart p = PefEnum::load_part(idx.freqs.parts, f.first_part + pi);
            PefBody b = pef_open_body(idx.freqs, p, true);
            // prefix sums c[i]; freq[i] = c[i] - c[i-1]: every chunk is decoded from one element earlier
            for (uint32_t i0 = it.first; i0 < it.first + it.count; i0 += 127) {
                uint32_t cnt = min(127u, it.first + it.count - i0);
                uint32_t ls = i0 ? i0 - 1 : 0;                 // first decoded local index
                uint32_t n = cnt + (i0 - ls);
                pef_decode_range(idx.freqs, p, b, ls, n, buf);
                uint32_t prev0 = pi ? p.base - 1u : 0u;         // prefix sum before the partition's first element
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) {
                    uint32_t e = lane + 32u * j;
                    if (e < cnt) {
                        uint32_t k = e + (i0 - ls);
                        job.out_freqs[o + p.begin + i0 + e] = buf[k] - (k ? buf[k - 1] : prev0);
                    }
                }
                __syncwarp();
            }
        }
    }
}

struct PefGeqJob {
    const uint32_t* terms;
    const uint64_t* bounds;
    const uint64_t* bound_offsets;
    uint64_t* out_docids;
    uint64_t* out_freqs;
    uint32_t* work_counter;
    uint32_t nlists;
};

static __global__ void __launch_bounds__(128) pef_next_geq_kernel(PefIndexDev idx, PefGeqJob job) {
    PefState* st = reinterpret_cast<PefState*>(g_smem + (threadIdx.x >> 5) * sizeof(PefState));
    const unsigned lane = lane_id();
    WarpCtx c;
    c.c_docs_blocks = c.c_freqs_blocks = c.c_docs_bytes = c.c_freqs_bytes = c.c_maxs = c.c_scored = 0;
    while (true) {
        uint32_t li = 0;
        if (lane == 0) li = atomicAdd(job.work_counter, 1u);
        li = __shfl_sync(FULL, li, 0);
        if (li >= job.nlists) break;
        PefEnum::open(c, idx, st, job.terms[li]);
        for (uint64_t j = job.bound_offsets[li]; j < job.bound_offsets[li + 1]; ++j) {
            uint64_t lb = job.bounds[j];
            uint32_t d = lb >= idx.num_docs ? PefEnum::next_geq(c, idx, st, idx.num_docs) : PefEnum::next_geq(c, idx, st, uint32_t(lb));
            uint32_t f = d < idx.num_docs ? PefEnum::freq(c, idx, st) : 0u;
            if (lane == 0) { job.out_docids[j] = d; job.out_freqs[j] = f; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host: loader

struct PefSeqHost {
    std::vector<PefListDir> lists;
    std::vector<PefPart> parts;
    uint64_t* d_bits = nullptr;
    PefListDir* d_lists = nullptr;
    PefPart* d_parts = nullptr;
    uint64_t device_bytes = 0;
    ~PefSeqHost() { if (d_bits) cudaFree(d_bits); if (d_lists) cudaFree(d_lists); if (d_parts) cudaFree(d_parts); }
};

struct bit_cursor {          // succinct::bit_vector::enumerator subset
    bitvec_view const& bv;
    uint64_t pos;
    uint64_t take(uint32_t len) { uint64_t v = get_bits(bv, pos, len); pos += len; return v; }
    uint32_t skip_zeros() {
        uint32_t z = 0;
        while (true) {
            if (pos >= bv.bits) throw format_error("gamma code runs past the bit vector");
            if (get_bits(bv, pos, 1)) { ++pos; return z; }
            ++pos; ++z;
        }
    }
    uint64_t gamma() { uint32_t l = skip_zeros(); return (take(l) | (uint64_t(1) << l)) - 1; }       // integer_codes.hpp:21-25
    uint64_t gamma_nonzero() { return gamma() + 1; }
    uint64_t delta() { uint64_t l = gamma(); return (take(uint32_t(l)) | (uint64_t(1) << l)) - 1; }   // :41-45
};

// partitioned_sequence header at `offset` (partitioned_sequence.hpp:130-178) -> flat partitions
static inline void pef_parse_sequence(bitvec_view const& bv, uint64_t offset, uint64_t universe, uint64_t n, global_params const& gp,
                                      std::vector<PefPart>& out) {
    bit_cursor it{bv, offset};
    uint64_t partitions = it.gamma_nonzero();
    if (partitions == 0 || partitions > n) throw format_error("bad partition count");
    if (universe > 0x100000000ull) throw format_error("sequence universe beyond 32 bits");
    if (partitions == 1) {
        uint32_t universe_bits = uint32_t(ceil_log2_u64(universe));
        uint64_t base = it.take(universe_bits);
        uint64_t ub = 0;
        if (n > 1) {
            uint64_t universe_delta = it.delta();
            ub = universe_delta ? universe_delta : (universe - base - 1);
        }
        out.push_back(PefPart{it.pos, 0u, uint32_t(n), uint32_t(base), uint32_t(base + ub)});
        return;
    }
    uint64_t endpoint_bits = it.gamma();
    uint64_t cur = it.pos;
    std::vector<uint64_t> sizes = ef_decode_all(bv, cur, n, partitions - 1, gp);
    cur += ef_offsets(0, n, partitions - 1, gp.ef_log_sampling0, gp.ef_log_sampling1).end;
    std::vector<uint64_t> ubs = ef_decode_all(bv, cur, universe, partitions + 1, gp);
    cur += ef_offsets(0, universe, partitions + 1, gp.ef_log_sampling0, gp.ef_log_sampling1).end;
    uint64_t endpoints_offset = cur;
    uint64_t sequences_offset = cur + endpoint_bits * (partitions - 1);
    for (uint64_t p = 0; p < partitions; ++p) {
        uint64_t endpoint = p ? get_bits(bv, endpoints_offset + (p - 1) * endpoint_bits, uint32_t(endpoint_bits)) : 0;
        uint64_t begin = p ? sizes[p - 1] : 0, end = p + 1 < partitions ? sizes[p] : n;
        if (end <= begin || end > n) throw format_error("partition sizes out of order");
        uint64_t base = ubs[p] + (p ? 1 : 0), ub = ubs[p + 1];
        if (ub < base || ub >= universe) throw format_error("partition bounds out of order");
        out.push_back(PefPart{sequences_offset + endpoint, uint32_t(begin), uint32_t(end - begin), uint32_t(base), uint32_t(ub)});
    }
}

// uniform_partitioned_sequence header (uniform_partitioned_sequence.hpp:19-103,121-160): as above without the sizes
// sequence — every partition but the last holds 2^log_partition_size elements
static inline void pef_parse_uniform_sequence(bitvec_view const& bv, uint64_t offset, uint64_t universe, uint64_t n, global_params const& gp,
                                              std::vector<PefPart>& out) {
    bit_cursor it{bv, offset};
    uint64_t partitions = it.gamma_nonzero();
    const uint64_t psize = uint64_t(1) << gp.log_partition_size;
    if (partitions != (n + psize - 1) / psize) throw format_error("bad uniform partition count");
    if (universe > 0x100000000ull) throw format_error("sequence universe beyond 32 bits");
    if (partitions == 1) {
        uint32_t universe_bits = uint32_t(ceil_log2_u64(universe));
        uint64_t base = it.take(universe_bits);
        uint64_t ub = 0;
        if (n > 1) {
            uint64_t universe_delta = it.delta();
            ub = universe_delta ? universe_delta : (universe - base - 1);
        }
        out.push_back(PefPart{it.pos, 0u, uint32_t(n), uint32_t(base), uint32_t(base + ub)});
        return;
    }
    uint64_t endpoint_bits = it.gamma();
    uint64_t cur = it.pos;
    std::vector<uint64_t> ubs = ef_decode_all(bv, cur, universe, partitions + 1, gp);
    cur += ef_offsets(0, universe, partitions + 1, gp.ef_log_sampling0, gp.ef_log_sampling1).end;
    uint64_t endpoints_offset = cur;
    uint64_t sequences_offset = cur + endpoint_bits * (partitions - 1);
    for (uint64_t p = 0; p < partitions; ++p) {
        uint64_t endpoint = p ? get_bits(bv, endpoints_offset + (p - 1) * endpoint_bits, uint32_t(endpoint_bits)) : 0;
        uint64_t begin = p * psize, end = std::min<uint64_t>(n, (p + 1) * psize);
        uint64_t base = ubs[p] + (p ? 1 : 0), ub = ubs[p + 1];
        if (ub < base || ub >= universe) throw format_error("partition bounds out of order");
        out.push_back(PefPart{sequences_offset + endpoint, uint32_t(begin), uint32_t(end - begin), uint32_t(base), uint32_t(ub)});
    }
}

// which freq_index instantiation the file holds (index_types.hpp:18-35)
enum : int { PEF_VARIANT_OPT = 0, PEF_VARIANT_UNIFORM = 1, PEF_VARIANT_SINGLE = 2, PEF_VARIANT_EF = 3 };

static inline void pef_parse_any(int variant, bitvec_view const& bv, uint64_t offset, uint64_t universe, uint64_t n, global_params const& gp,
                                 std::vector<PefPart>& out) {
    if (universe > 0x100000000ull) throw format_error("sequence universe beyond 32 bits");
    switch (variant) {
        case PEF_VARIANT_OPT: pef_parse_sequence(bv, offset, universe, n, gp, out); break;
        case PEF_VARIANT_UNIFORM: pef_parse_uniform_sequence(bv, offset, universe, n, gp, out); break;
        default:
            // single_index: one indexed_sequence / strict_sequence over the whole list; ef_index: one (strict_)elias_fano.
            // Both are "one partition with base 0" in the flattened directory.
            out.push_back(PefPart{offset, 0u, uint32_t(n), 0u, uint32_t(universe - 1)});
    }
}

// block directory of the docs sequences: entry j of a list = (last docid of its j-th 128-element window, index of the window's
// partition inside the list) — the analogue of the block indexes' (block_max, endpoint) directory, so the block-parallel
// query kernels find a window of an Elias-Fano list with the same ballot search they use for block_max
static __global__ void __launch_bounds__(256) pef_block_dir_kernel(const uint32_t* docs, const uint64_t* last_elem, const uint32_t* part_rel,
                                                                   uint64_t nblocks, uint2* out) {
    for (uint64_t j = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < nblocks; j += uint64_t(gridDim.x) * blockDim.x)
        out[j] = make_uint2(__ldg(docs + last_elem[j]), part_rel[j]);
}

struct PefIndexHost {
    uint64_t size = 0, num_docs = 0, device_bytes = 0;
    uint2* d_bdir = nullptr;            // block directory (device), total_blocks + 1 entries
    uint32_t* d_bfirst = nullptr;       // size + 1 entries: first block of every list
    ListDir* d_dir = nullptr;           // per list: n (the other ListDir fields are unused by this index family)
    uint64_t total_blocks = 0;
    ~PefIndexHost() { if (d_bdir) cudaFree(d_bdir); if (d_bfirst) cudaFree(d_bfirst); if (d_dir) cudaFree(d_dir); }
    struct host_list { uint64_t n; uint64_t bits; uint64_t blocks; };      // postings; bits of the list in the docs + freqs bit vectors; 128-element windows of the docs sequence
    std::vector<host_list> host_dir;
    PefSeqHost docs, freqs;
    PefIndexDev dev{};

    static int upload(PefSeqHost& s, bitvec_view const& bv, std::string& err) {
        size_t words = size_t(bv.nwords) + 64;       // zero tail: scans read 32 words per step
        if (cudaMalloc(reinterpret_cast<void**>(&s.d_bits), words * 8) != cudaSuccess) { err = "cudaMalloc failed (bit vector)"; return -3; }
        if (cudaMemset(s.d_bits, 0, words * 8) != cudaSuccess) { err = "cudaMemset failed (bit vector)"; return -3; }
        if (bv.nwords && cudaMemcpy(s.d_bits, bv.raw, size_t(bv.nwords) * 8, cudaMemcpyHostToDevice) != cudaSuccess) { err = "H2D copy failed"; return -3; }
        if (cudaMalloc(reinterpret_cast<void**>(&s.d_lists), std::max<size_t>(1, s.lists.size()) * sizeof(PefListDir)) != cudaSuccess ||
            cudaMalloc(reinterpret_cast<void**>(&s.d_parts), (s.parts.size() + 64) * sizeof(PefPart)) != cudaSuccess) { err = "cudaMalloc failed (directory)"; return -3; }
        if (cudaMemset(s.d_parts, 0, (s.parts.size() + 64) * sizeof(PefPart)) != cudaSuccess ||
            cudaMemcpy(s.d_lists, s.lists.data(), s.lists.size() * sizeof(PefListDir), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(s.d_parts, s.parts.data(), s.parts.size() * sizeof(PefPart), cudaMemcpyHostToDevice) != cudaSuccess) { err = "H2D copy failed (directory)"; return -3; }
        s.device_bytes = words * 8 + s.lists.size() * sizeof(PefListDir) + s.parts.size() * sizeof(PefPart);
        return 0;
    }

    // Decodes every docs sequence once on the device (chunks of lists, so the scratch stays bounded) and keeps the last
    // docid of every 128-element window.  Load-time work, like the (block_max, endpoint) directory of the block indexes.
    int build_block_directory(std::string& err);

    // freq_index::map (freq_index.hpp:234-243) + per-list headers (freq_index.hpp:192-214)
    int load(const uint8_t* p, size_t nbytes, int variant, std::string& err) {
        try {
            byte_reader r(p, nbytes);
            (void)r.get<uint64_t>();
            global_params gp = read_params(r);
            num_docs = r.get<uint64_t>();
            if (num_docs == 0 || num_docs > 0xFFFFFFFFull) throw format_error("num_docs out of range");
            uint64_t dsize = r.get<uint64_t>();
            bitvec_view dend = read_bitvec(r), dbits = read_bitvec(r);
            uint64_t fsize = r.get<uint64_t>();
            bitvec_view fend = read_bitvec(r), fbits = read_bitvec(r);
            if (dsize != fsize) throw format_error("docs and freqs collections differ in size");
            size = dsize;
            std::vector<uint64_t> dstart = ef_decode_all(dend, 0, dbits.bits, size, gp);
            std::vector<uint64_t> fstart = ef_decode_all(fend, 0, fbits.bits, size, gp);
            host_dir.resize(size);
            docs.lists.resize(size); freqs.lists.resize(size);
            for (uint64_t i = 0; i < size; ++i) {
                bit_cursor it{dbits, dstart[i]};
                uint64_t occurrences = it.gamma_nonzero();
                uint64_t n = 1;
                if (occurrences > 1) n = it.take(uint32_t(ceil_log2_u64(occurrences + 1)));
                if (n == 0 || n > num_docs) throw format_error("bad list length");
                host_dir[i].n = n;
                host_dir[i].bits = ((i + 1 < size ? dstart[i + 1] : dbits.bits) - dstart[i]) + ((i + 1 < size ? fstart[i + 1] : fbits.bits) - fstart[i]);
                docs.lists[i] = PefListDir{docs.parts.size(), 0u, uint32_t(n)};
                pef_parse_any(variant, dbits, it.pos, num_docs, n, gp, docs.parts);
                docs.lists[i].nparts = uint32_t(docs.parts.size() - docs.lists[i].first_part);
                freqs.lists[i] = PefListDir{freqs.parts.size(), 0u, uint32_t(n)};
                pef_parse_any(variant, fbits, fstart[i], occurrences + 1, n, gp, freqs.parts);
                freqs.lists[i].nparts = uint32_t(freqs.parts.size() - freqs.lists[i].first_part);
                // windows before each docs partition, and how many bits each body spans (bodies are back to back)
                auto finish = [&](PefSeqHost& sq, uint64_t list_end_bit) {
                    uint64_t fb = 0;
                    for (uint64_t k = sq.lists[i].first_part; k < sq.parts.size(); ++k) {
                        PefPart& pp = sq.parts[k];
                        pp.first_block = uint32_t(fb);
                        fb += (uint64_t(pp.size) + 127) / 128;
                        const uint64_t end = k + 1 < sq.parts.size() ? sq.parts[k + 1].bit_off : list_end_bit;
                        pp.body_bits = end > pp.bit_off && end - pp.bit_off < 0xffffffffull ? uint32_t(end - pp.bit_off) : 0u;
                    }
                    return fb;
                };
                const uint64_t nb = finish(docs, i + 1 < size ? dstart[i + 1] : dbits.bits);
                finish(freqs, i + 1 < size ? fstart[i + 1] : fbits.bits);
                if (total_blocks + nb > 0xfffffff0ull) throw format_error("more than 2^32 windows in the index");
                total_blocks += nb;
                host_dir[i].blocks = nb;
            }
            int rc = upload(docs, dbits, err);
            if (rc) return rc;
            rc = upload(freqs, fbits, err);
            if (rc) return rc;
            auto mk = [&](PefSeqHost& s) {
                PefSeq d;
                d.bits = s.d_bits; d.lists = s.d_lists; d.parts = s.d_parts;
                d.log_sampling0 = gp.ef_log_sampling0; d.log_sampling1 = gp.ef_log_sampling1;
                d.rb_log_rank1_sampling = gp.rb_log_rank1_sampling; d.rb_log_sampling1 = gp.rb_log_sampling1;
                d.raw_ef = variant == PEF_VARIANT_EF ? 1u : 0u;
                return d;
            };
            dev.docs = mk(docs); dev.freqs = mk(freqs); dev.num_lists = size; dev.num_docs = uint32_t(num_docs);
            device_bytes = docs.device_bytes + freqs.device_bytes;
            rc = build_block_directory(err);
            if (rc) return rc;
            if (getenv("DS2I_GPU_TRACE")) {
                uint64_t staged = 0, big_windows = 0;
                for (auto const& pp : docs.parts) {
                    const uint64_t bytes = ((pp.bit_off + pp.body_bits + 7) >> 3) - (pp.bit_off >> 3) + 30;
                    if (pp.body_bits && bytes <= 1280) ++staged; else big_windows += (uint64_t(pp.size) + 127) / 128;
                }
                fprintf(stderr, "[ds2i_gpu] opt index: %llu lists, %zu docs partitions (%llu small enough to stage; %llu windows in the others), %zu freqs partitions, %llu windows\n",
                        (unsigned long long)size, docs.parts.size(), (unsigned long long)staged, (unsigned long long)big_windows, freqs.parts.size(), (unsigned long long)total_blocks);
            }
        } catch (std::exception const& e) {
            err = e.what();
            return -2;
        }
        return 0;
    }
};

// ------------------------------------------------------------------------------------------------
// host: launchers

template <int OP>
static int pef_launch_op(PefIndexHost& ix, DevWand wand, DevBatch const& db, uint32_t k, int max_terms, int sm_count, std::string& err) {
    const int warps = 4;
    size_t smem = S16_TAB_BYTES + warps * warp_smem_bytes_t<PefState>(max_terms);
    auto kern = pef_query_kernel<OP>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess) { err = "cudaFuncSetAttribute failed"; return -3; }
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem);
    if (per_sm < 1) { err = "opt query kernel does not fit on an SM"; return -3; }
    int grid = std::max(1, std::min(per_sm * sm_count, int((db.nq + warps - 1) / warps)));
    kern<<<grid, warps * 32, smem>>>(ix.dev, wand, db, k, max_terms);
    return 0;
}

inline int pef_launch_query(PefIndexHost& ix, DevWand wand, DevBatch const& db, int op, uint32_t k, int max_terms, int sm_count, std::string& err) {
    switch (op) {
        case OP_AND: return pef_launch_op<OP_AND>(ix, wand, db, k, max_terms, sm_count, err);
#ifndef DS2I_DEV_FAST_BUILD      // kernel-experiment builds instantiate one operator only (seconds instead of minutes)
        case OP_AND_FREQ: return pef_launch_op<OP_AND_FREQ>(ix, wand, db, k, max_terms, sm_count, err);
        case OP_OR: return pef_launch_op<OP_OR>(ix, wand, db, k, max_terms, sm_count, err);
        case OP_OR_FREQ: return pef_launch_op<OP_OR_FREQ>(ix, wand, db, k, max_terms, sm_count, err);
        case OP_RANKED_AND: return pef_launch_op<OP_RANKED_AND>(ix, wand, db, k, max_terms, sm_count, err);
        case OP_WAND: return pef_launch_op<OP_WAND>(ix, wand, db, k, max_terms, sm_count, err);
        case OP_MAXSCORE: return pef_launch_op<OP_MAXSCORE>(ix, wand, db, k, max_terms, sm_count, err);
        case OP_RANKED_OR: return pef_launch_op<OP_RANKED_OR>(ix, wand, db, k, max_terms, sm_count, err);
#endif
    }
    err = "unknown operator";
    return -1;
}

// work items of a full decode, cut on the host from the partition directory and uploaded (outside the timed region)
inline int pef_decode_prepare(PefIndexHost& ix, const uint32_t* h_terms, uint32_t nterms, PefDecodeItem** d_items, uint32_t* nitems, std::string& err) {
    std::vector<PefDecodeItem> items;
    for (uint32_t i = 0; i < nterms; ++i) {
        for (int fq = 0; fq < 2; ++fq) {
            PefSeqHost const& sq = fq ? ix.freqs : ix.docs;
            PefListDir const& d = sq.lists[h_terms[i]];
            for (uint32_t pi = 0; pi < d.nparts; ++pi) {
                uint32_t size = sq.parts[d.first_part + pi].size;
                for (uint32_t first = 0; first < size; first += PEF_ITEM_ELEMS)
                    items.push_back(PefDecodeItem{i, pi | (fq ? 0x80000000u : 0u), first, std::min(PEF_ITEM_ELEMS, size - first)});
            }
        }
    }
    *nitems = uint32_t(items.size());
    *d_items = nullptr;
    if (items.empty()) return 0;
    if (cudaMalloc(reinterpret_cast<void**>(d_items), items.size() * sizeof(PefDecodeItem)) != cudaSuccess) { err = "cudaMalloc failed"; return -3; }
    if (cudaMemcpy(*d_items, items.data(), items.size() * sizeof(PefDecodeItem), cudaMemcpyHostToDevice) != cudaSuccess) { err = "H2D copy failed"; return -3; }
    return 0;
}

inline void pef_decode_launch(PefIndexHost& ix, const PefDecodeItem* d_items, uint32_t nitems, const uint32_t* d_terms, const uint64_t* d_offsets,
                              uint32_t* d_docs, uint32_t* d_freqs, int sm_count) {
    if (!nitems) return;
    PefDecodeJob job{d_terms, d_items, d_offsets, d_docs, d_freqs, nitems};
    int grid = int(std::max<uint64_t>(1, std::min<uint64_t>((uint64_t(nitems) + 7) / 8, uint64_t(sm_count) * 8)));
    pef_decode_kernel<<<grid, 256, 8 * 4096>>>(ix.dev, job);
}

inline int pef_next_geq(PefIndexHost& ix, const uint32_t* d_terms, uint32_t nlists, const uint64_t* d_bounds, const uint64_t* d_offsets,
                        uint64_t* d_docids, uint64_t* d_freqs, uint32_t* d_counter, int sm_count, std::string&) {
    PefGeqJob job{d_terms, d_bounds, d_offsets, d_docids, d_freqs, d_counter, nlists};
    int grid = int(std::max<uint64_t>(1, std::min<uint64_t>((uint64_t(nlists) + 3) / 4, uint64_t(sm_count) * 8)));
    pef_next_geq_kernel<<<grid, 128, 4 * sizeof(PefState)>>>(ix.dev, job);
    return 0;
}

inline int PefIndexHost::build_block_directory(std::string& err) {
    auto fail_cuda = [&](const char* what) { err = std::string(what) + ": " + cudaGetErrorString(cudaGetLastError()); return -3; };
    std::vector<uint32_t> bfirst(size + 1, 0);
    std::vector<ListDir> dir(size);
    for (uint64_t i = 0; i < size; ++i) {
        uint64_t nb = 0;
        for (uint32_t k = 0; k < docs.lists[i].nparts; ++k) nb += (uint64_t(docs.parts[docs.lists[i].first_part + k].size) + 127) / 128;
        bfirst[i + 1] = bfirst[i] + uint32_t(nb);
        dir[i] = ListDir{0, uint32_t(host_dir[i].n), 0};
    }
    if (cudaMalloc(reinterpret_cast<void**>(&d_bdir), (total_blocks + 1) * sizeof(uint2)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&d_bfirst), (size + 1) * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&d_dir), std::max<uint64_t>(size, 1) * sizeof(ListDir)) != cudaSuccess) return fail_cuda("cudaMalloc (block directory)");
    const uint2 sentinel = make_uint2(0xffffffffu, 0u);            // probes may read one entry past the last list
    if (cudaMemcpy(d_bdir + total_blocks, &sentinel, sizeof(uint2), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(d_bfirst, bfirst.data(), (size + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(d_dir, dir.data(), size * sizeof(ListDir), cudaMemcpyHostToDevice) != cudaSuccess) return fail_cuda("H2D copy (block directory)");
    device_bytes += (total_blocks + 1) * sizeof(uint2) + (size + 1) * 4 + size * sizeof(ListDir);

    const uint64_t chunk_postings = uint64_t(1) << 27;              // 512 MB of decoded docids at a time
    std::vector<PefDecodeItem> items;
    std::vector<uint32_t> terms, part_rel;
    std::vector<uint64_t> offsets, last_elem;
    for (uint64_t l0 = 0; l0 < size;) {
        uint64_t l1 = l0, postings = 0;
        while (l1 < size && (l1 == l0 || postings + host_dir[l1].n <= chunk_postings)) postings += host_dir[l1++].n;
        items.clear(); terms.clear(); part_rel.clear(); offsets.clear(); last_elem.clear();
        uint64_t at = 0;
        for (uint64_t l = l0; l < l1; ++l) {
            terms.push_back(uint32_t(l));
            offsets.push_back(at);
            PefListDir const& d = docs.lists[l];
            for (uint32_t pi = 0; pi < d.nparts; ++pi) {
                PefPart const& pp = docs.parts[d.first_part + pi];
                for (uint32_t first = 0; first < pp.size; first += PEF_ITEM_ELEMS)
                    items.push_back(PefDecodeItem{uint32_t(l - l0), pi, first, std::min(PEF_ITEM_ELEMS, pp.size - first)});
                for (uint32_t w = 0; w * 128u < pp.size; ++w) {
                    last_elem.push_back(at + pp.begin + std::min<uint64_t>(pp.size, 128ull * (w + 1)) - 1);
                    part_rel.push_back(pi);
                }
            }
            at += host_dir[l].n;
        }
        offsets.push_back(at);
        const uint64_t nb = last_elem.size();
        uint32_t *d_terms = nullptr, *d_docs = nullptr, *d_prel = nullptr;
        uint64_t *d_offs = nullptr, *d_last = nullptr;
        PefDecodeItem* d_items = nullptr;
        auto cleanup = [&]() { cudaFree(d_terms); cudaFree(d_docs); cudaFree(d_prel); cudaFree(d_offs); cudaFree(d_last); cudaFree(d_items); };
        if (cudaMalloc(reinterpret_cast<void**>(&d_terms), terms.size() * 4) != cudaSuccess || cudaMalloc(reinterpret_cast<void**>(&d_docs), std::max<uint64_t>(at, 1) * 4) != cudaSuccess ||
            cudaMalloc(reinterpret_cast<void**>(&d_prel), std::max<uint64_t>(nb, 1) * 4) != cudaSuccess || cudaMalloc(reinterpret_cast<void**>(&d_offs), offsets.size() * 8) != cudaSuccess ||
            cudaMalloc(reinterpret_cast<void**>(&d_last), std::max<uint64_t>(nb, 1) * 8) != cudaSuccess ||
            cudaMalloc(reinterpret_cast<void**>(&d_items), std::max<size_t>(items.size(), 1) * sizeof(PefDecodeItem)) != cudaSuccess) { cleanup(); return fail_cuda("cudaMalloc (block directory scratch)"); }
        bool ok = cudaMemcpy(d_terms, terms.data(), terms.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
                  cudaMemcpy(d_offs, offsets.data(), offsets.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
                  cudaMemcpy(d_prel, part_rel.data(), nb * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
                  cudaMemcpy(d_last, last_elem.data(), nb * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
                  cudaMemcpy(d_items, items.data(), items.size() * sizeof(PefDecodeItem), cudaMemcpyHostToDevice) == cudaSuccess;
        if (ok && !items.empty()) {
            PefDecodeJob job{d_terms, d_items, d_offs, d_docs, nullptr, uint32_t(items.size())};
            int grid = int(std::max<uint64_t>(1, std::min<uint64_t>((items.size() + 7) / 8, 148ull * 8)));
            pef_decode_kernel<<<grid, 256, 8 * 4096>>>(dev, job);
            pef_block_dir_kernel<<<int(std::max<uint64_t>(1, std::min<uint64_t>((nb + 255) / 256, 148ull * 8))), 256>>>(d_docs, d_last, d_prel, nb, d_bdir + bfirst[l0]);
            ok = cudaDeviceSynchronize() == cudaSuccess;
        }
        cleanup();
        if (!ok) return fail_cuda("block directory build");
        l0 = l1;
    }
    return 0;
}

}  // namespace ds2i_gpu
