// Batched decode of whole posting lists (BASELINE config 2; ds2i_gpu_decode_lists): docids and
// freqs of every posting materialised in HBM.  The reference decodes a list block by block through
// document_enumerator (block_posting_list.hpp:292-331); here every 128-posting block is an
// independent unit: a warp takes 32 consecutive blocks of the job, reads their (block_max,
// endpoint) directory entries with one load per lane, stages each [docs | freqs] block pair with one
// TMA bulk copy and decodes both halves from shared memory.  Blocks coded with the bit-serial
// interpolative codec (list tails; every block of block_interpolative) go to
// decode_serial_blocks_kernel, one lane per block.
#pragma once
#include "and_kernels.cuh"

namespace ds2i_gpu {

struct DecodeJob {
    const uint32_t* terms;        // nterms
    const uint64_t* blk_prefix;   // nterms+1: blocks before list i
    const uint64_t* out_offsets;  // nterms+1: postings before list i
    uint32_t* out_docs;
    uint32_t* out_freqs;
    uint64_t total_blocks;
    uint32_t nterms;
    const uint32_t* tail_order;   // nterms: positions of the job's lists ordered by the size of their partial last block (nullable)
    uint32_t tail_first, tail_count;   // the slice of tail_order a launch of the tail kernel covers
};

constexpr size_t DECODE_WARP_BYTES = 2 * BLOCK * 4 + STAGE_WORDS * 4 + SCRATCH_WORDS * 4 + 16;

// last list l in [0, n) with prefix[l] <= g (prefix is non-decreasing, prefix[0] == 0): 32 probes per step
__device__ __forceinline__ uint32_t warp_upper_list(const uint64_t* prefix, uint32_t n, uint64_t g) {
    const unsigned lane = lane_id();
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const uint32_t span = hi - lo;
        const uint32_t step = (span + 31u) / 32u;
        const uint32_t p = lo + lane * step;
        const bool le = p < hi && __ldg(prefix + p) <= g;
        const unsigned m = __ballot_sync(FULL, le);
        const uint32_t f = 31u - __clz(m);
        lo = lo + f * step;
        hi = min(hi, lo + step);
    }
    return lo;
}

template <int CODEC>
__global__ void __launch_bounds__(128, 8) decode_full_blocks_kernel(DevIndex idx, DecodeJob job) {
    s16_table_init(smem_words(0));
    __syncthreads();
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    uint8_t* base = g_smem + S16_TAB_BYTES + warp * DECODE_WARP_BYTES;
    uint32_t* docs = reinterpret_cast<uint32_t*>(base);
    uint32_t* freqs = docs + BLOCK;
    uint32_t* stage = freqs + BLOCK;
    uint32_t* stack = stage + STAGE_WORDS;
    uint64_t* bar = reinterpret_cast<uint64_t*>(stack + SCRATCH_WORDS);
    const uint32_t stage_off = smem_offset(stage), stack_off = smem_offset(stack), docs_off = smem_offset(docs), freqs_off = smem_offset(freqs);
    uint32_t phase = 0;
    if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncwarp();

    constexpr uint64_t CHUNK = 32;
    const uint64_t nwarps = uint64_t(gridDim.x) * (blockDim.x >> 5);
    const uint64_t nchunks = (job.total_blocks + CHUNK - 1) / CHUNK;
    for (uint64_t ch = uint64_t(blockIdx.x) * (blockDim.x >> 5) + warp; ch < nchunks; ch += nwarps) {
        const uint64_t g0 = ch * CHUNK, g1 = min(job.total_blocks, g0 + CHUNK);
        uint32_t li = warp_upper_list(job.blk_prefix, job.nterms, g0);
        uint64_t g = g0;
        while (g < g1) {
            const uint64_t lf = __ldg(job.blk_prefix + li), le = __ldg(job.blk_prefix + li + 1);
            if (g >= le) { ++li; continue; }
            const uint32_t term = __ldg(job.terms + li);
            const ListDir d = idx.dir[term];
            const uint32_t nblocks = uint32_t(le - lf);
            const uint2* bd = idx.bdir + idx.bfirst[term];
            const uint64_t data_off = d.maxs_off + 4ull * nblocks + 4ull * (nblocks - 1);
            const uint32_t b_lo = uint32_t(g - lf), cnt = uint32_t(min(g1, le) - g);
            const uint64_t out_base = __ldg(job.out_offsets + li);
            // directory entries of the run, one block per lane
            uint2 en = make_uint2(0u, 0u), prev = make_uint2(0xffffffffu, 0u);
            if (lane < cnt) en = __ldg(bd + b_lo + lane);
            if (b_lo) prev = __ldg(bd + b_lo - 1);
            for (uint32_t r = 0; r < cnt; ++r) {
                const uint32_t b = b_lo + r;
                const uint32_t pm_s = __shfl_sync(FULL, en.x, (r + 31) & 31), pe_s = __shfl_sync(FULL, en.y, (r + 31) & 31);
                const uint32_t cur_max = __shfl_sync(FULL, en.x, r), e1 = __shfl_sync(FULL, en.y, r);
                const uint32_t prev_max = r ? pm_s : prev.x, e0 = r ? pe_s : prev.y;
                if (CODEC == CODEC_INTERPOLATIVE || (b + 1u) * BLOCK > d.n) continue;      // bit-serial blocks: the other kernel
                const uint32_t off = and_stage(idx.lists, data_off + e0, data_off + e1, stage, bar, phase);
                bool dprefix, fprefix;
                const uint32_t consumed = and_decode_values<CODEC, true>(stage_off, off, BLOCK, cur_max - prev_max - BLOCK, docs_off, stack_off, dprefix);
                and_decode_values<CODEC, true>(stage_off, off + consumed, BLOCK, 0xffffffffu, freqs_off, stack_off, fprefix);
                uint32_t* od = job.out_docs + out_base + uint64_t(b) * BLOCK + lane;
                uint32_t* of = job.out_freqs + out_base + uint64_t(b) * BLOCK + lane;
                if (!dprefix) {
                    uint4 v = reinterpret_cast<uint4*>(docs)[lane];
                    v.y += v.x; v.z += v.y; v.w += v.z;
                    const uint32_t incl = warp_inclusive_scan(v.w);
                    const uint32_t add = prev_max + 1u + (incl - v.w) + 4u * lane;     // docid_i = base + sum_{k<=i} gap_k + i
                    v.x += add; v.y += add + 1u; v.z += add + 2u; v.w += add + 3u;
                    reinterpret_cast<uint4*>(docs)[lane] = v;
                    __syncwarp();
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) od[32 * j] = docs[32 * j + lane];
                } else {          // interpolative leaves prefix sums (mixed index only)
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) od[32 * j] = prev_max + 1u + docs[32 * j + lane] + 32 * j + lane;
                }
                if (!fprefix) {
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) of[32 * j] = freqs[32 * j + lane] + 1u;
                } else {
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) {
                        const uint32_t i = 32 * j + lane;
                        of[32 * j] = freqs[i] - (i ? freqs[i - 1] : 0u) + 1u;
                    }
                }
                (void)cur_max;
            }
            g += cnt;
        }
    }
}

}  // namespace ds2i_gpu
