// libds2i_gpu.so — C ABI (include/ds2i_gpu.h) over the sm_100a kernels.  Host side: parse ds2i's
// on-disk format, build the flat list directory, copy the compressed index once into HBM, turn
// each query into the device descriptor (term order, host-computed BM25 query weights), launch.
// There is no CPU fallback: without a CUDA device every entry point fails with DS2I_E_CUDA.
#include "../../include/ds2i_gpu.h"

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <cstdio>
#include <cstring>
#include <limits>
#include <cstdlib>
#include <ctime>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>
#include <fcntl.h>
#include <nccl.h>          // types only: the library is dlopen'ed by the multi-GPU entry points, never linked
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "format.hpp"
#include "query_kernels.cuh"
#include "and_kernels.cuh"
#include "union_kernels.cuh"
#include "decode_kernels.cuh"
#include "pef_kernels.cuh"

using namespace ds2i_gpu;

static_assert(MAX_TERMS == DS2I_GPU_MAX_TERMS, "header/kernels disagree");
static_assert(MAX_K == DS2I_GPU_MAX_K, "header/kernels disagree");

// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static int fail(int code, std::string const& msg) { g_last_error = msg; return code; }

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fail(DS2I_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));       \
    } while (0)

// Device buffers come from the stream-ordered allocator with a pool that never trims: after the
// first batch a query_batch call allocates and frees without touching the driver's slow paths
// (cudaFree used to cost up to 290 ms per call in the end-to-end path).
static void tune_mem_pool(int device) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t threshold = ~uint64_t(0);
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
}

// makes `device` current for the scope and restores the caller's device afterwards: handles may live on different GPUs
// of one process (ds2i_gpu_query_batch_multi), and stream 0 / cudaFreeAsync act on the CURRENT device
struct device_scope {
    int prev = -1;
    cudaError_t status = cudaSuccess;           // of the switch: CUDA_TRY(scope.status) where a wrong device must not go unnoticed
    explicit device_scope(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) status = cudaSetDevice(device); else prev = -1;
    }
    ~device_scope() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename T>
struct dev_buf {
    T* p = nullptr;
    size_t n = 0;
    int device = -1;          // the device the buffer was allocated on
    void release() {
        if (!p) return;
        device_scope ds(device);
        cudaFreeAsync(p, 0);
        p = nullptr;
    }
    ~dev_buf() { release(); }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        cudaError_t e = cudaGetDevice(&device);
        if (e != cudaSuccess) return e;
        return cudaMallocAsync(reinterpret_cast<void**>(&p), std::max<size_t>(count, 1) * sizeof(T), 0);
    }
    cudaError_t upload(std::vector<T> const& v) {
        cudaError_t e = alloc(v.size());
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, 0);
    }
};

struct cuda_event {            // RAII: error paths must not leak events
    cudaEvent_t e = nullptr;
    ~cuda_event() { if (e) cudaEventDestroy(e); }
    cudaError_t create() { return cudaEventCreate(&e); }
};
struct cuda_free_guard {       // RAII for a plain cudaMalloc'ed pointer
    void* p = nullptr;
    ~cuda_free_guard() { if (p) cudaFree(p); }
};

struct mapped_file {
    const uint8_t* p = nullptr; size_t n = 0;
    ~mapped_file() { if (p) munmap(const_cast<uint8_t*>(p), n); }
    bool open(const char* path) {
        int fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st; fstat(fd, &st); n = size_t(st.st_size);
        void* m = n ? mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
        ::close(fd);
        if (n && m == MAP_FAILED) { p = nullptr; return false; }
        p = static_cast<const uint8_t*>(m);
        return true;
    }
};

// pinned host staging memory, grown on demand and kept for the life of the index handle (one batch call in flight per
// handle): every per-batch upload is ONE cudaMemcpyAsync from here instead of a dozen pageable copies
struct pinned_arena {
    uint8_t* p = nullptr;
    size_t cap = 0;
    ~pinned_arena() { if (p) cudaFreeHost(p); }
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = std::max<size_t>(bytes + bytes / 4, 1 << 20);
        cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&p), want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
};

enum : int { KIND_BLOCK = 0, KIND_PEF = 1 };

struct ds2i_gpu_index {
    int device = 0;
    int kind = KIND_BLOCK;
    int codec = CODEC_OPTPFOR;
    uint64_t size = 0, num_docs = 0;
    uint64_t device_bytes = 0;
    // collection-wide statistics of a document-partitioned shard (empty / 0: the index is the whole collection)
    std::vector<uint64_t> g_df;
    uint64_t g_num_docs = 0;
    // block indexes
    std::vector<ListDir> host_dir;
    dev_buf<uint8_t> d_lists;
    dev_buf<ListDir> d_dir;
    dev_buf<uint2> d_bdir;
    dev_buf<uint32_t> d_bfirst;
    DevIndex dev{};
    // opt (partitioned Elias-Fano) index
    std::unique_ptr<PefIndexHost> pef;
    int sm_count = 148;
    pinned_arena staging;     // host staging of the per-batch uploads / downloads
    std::mutex staging_mu;
};

struct ds2i_gpu_wand {
    int device = 0;
    uint64_t num_docs = 0, num_terms = 0;
    std::vector<float> h_max_term_weight;
    dev_buf<float> d_norm_lens, d_max_term_weight;
    DevWand dev{};
};

template <typename T> struct dev_view { T* p = nullptr; };     // a typed window of the batch's device arena

struct ds2i_gpu_batch {
    ds2i_gpu_index* index = nullptr;
    ds2i_gpu_wand* wand = nullptr;
    uint32_t nq = 0;
    int max_terms = 1;
    uint32_t last_k = 0;
    bool last_ranked = false;
    bool pending = false;          // an asynchronous run has been launched and not waited for yet
    size_t counters_bytes = 0;     // work counters + statistics: one contiguous stretch of the arena, zeroed before every run
    // ONE device allocation per batch: the uploaded descriptors first (one H2D copy), then the device-only buffers
    dev_buf<uint8_t> arena;
    dev_view<uint32_t> q_begin, term, sched, work_counter;
    dev_view<float> q_weight, max_weight;
    dev_view<uint8_t> ord_size, ord_maxw;
    // results of the last run, fused so that one collective / one D2H copy moves them: [counts: nq u64][scores: nq*k f32][docids: nq*k u32]
    dev_view<uint8_t> out_fused;
    dev_view<uint64_t> out_counts;
    dev_view<float> out_scores;
    dev_view<uint32_t> out_docids;
    dev_view<unsigned long long> stats;
    // block-at-a-time conjunctive path: work items = (query, chunk of blocks of its shortest list)
    dev_view<uint32_t> and_gstart, and_item_begin, and_item_counts, and_item_sizes;
    dev_view<uint8_t> and_qchunk;
    dev_buf<float> and_item_scores;
    size_t and_item_scores_k = 0;
    uint32_t n_and_items = 0;
    // block-parallel union path (wand / maxscore): work items = (query, driving list, run of its blocks), implicit
    dev_view<uint32_t> un_gstart, un_gterm, un_gquery, un_gbase, un_gblocks, un_item_begin, un_item_sizes, un_threshold;
    dev_view<float> un_ub;
    dev_buf<float> un_item_scores;
    size_t un_item_scores_k = 0;      // k the partial top-k buffer was sized for
    uint32_t n_un_items = 0, n_un_groups = 0, un_item_blocks = 64;
    unsigned items_built = 3;      // which work-item lists exist (bit 0 conjunctive, bit 1 union)
    uint64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ~ds2i_gpu_batch() { if (ev0) cudaEventDestroy(ev0); if (ev1) cudaEventDestroy(ev1); }
    // the fused result buffer is laid out for the k of the run
    void layout_outputs(uint32_t k) {
        out_counts.p = reinterpret_cast<uint64_t*>(out_fused.p);
        out_scores.p = reinterpret_cast<float*>(out_fused.p + size_t(nq) * 8);
        out_docids.p = reinterpret_cast<uint32_t*>(out_fused.p + size_t(nq) * 8 + size_t(nq) * k * 4);
    }
    size_t fused_bytes(uint32_t k) const { return size_t(nq) * (8 + 8 * size_t(k)); }
};

static int codec_from_type(const char* t, int* kind) {
    std::string s(t ? t : "");
    *kind = KIND_BLOCK;
    if (s == "block_optpfor") return CODEC_OPTPFOR;
    if (s == "block_varint") return CODEC_VARINT;
    if (s == "block_interpolative") return CODEC_INTERPOLATIVE;
    if (s == "block_qmx") return CODEC_QMX;
    if (s == "block_mixed") return CODEC_MIXED;
    if (s == "opt") { *kind = KIND_PEF; return PEF_VARIANT_OPT; }
    if (s == "uniform") { *kind = KIND_PEF; return PEF_VARIANT_UNIFORM; }
    if (s == "single") { *kind = KIND_PEF; return PEF_VARIANT_SINGLE; }
    if (s == "ef") { *kind = KIND_PEF; return PEF_VARIANT_EF; }
    return -1;
}

// ------------------------------------------------------------------------------------------------
extern "C" const char* ds2i_gpu_last_error(void) { return g_last_error.c_str(); }

extern "C" int ds2i_gpu_op_from_name(const char* name) {
    static const char* names[] = {"and", "and_freq", "or", "or_freq", "ranked_and", "wand", "maxscore", "ranked_or"};
    if (!name) return DS2I_E_ARG;
    for (int i = 0; i < 8; ++i) if (!strcmp(name, names[i])) return i;
    return DS2I_E_ARG;
}

extern "C" int ds2i_gpu_index_type_known(const char* index_type) {
    int kind;
    return index_type && codec_from_type(index_type, &kind) >= 0 ? 1 : 0;
}

extern "C" int ds2i_gpu_index_open(const void* file_bytes, size_t nbytes, const char* index_type, int device,
                                   ds2i_gpu_index** out) {
    if (!file_bytes || !out || !index_type) return fail(DS2I_E_ARG, "null argument");
    int kind;
    int codec = codec_from_type(index_type, &kind);
    if (codec < 0) return fail(DS2I_E_UNSUPPORTED, std::string("unsupported index type ") + index_type);
    device_scope on_device(device);
    CUDA_TRY(on_device.status);
    std::unique_ptr<ds2i_gpu_index> ix(new ds2i_gpu_index);
    ix->device = device; ix->kind = kind; ix->codec = codec;
    cudaDeviceGetAttribute(&ix->sm_count, cudaDevAttrMultiProcessorCount, device);
    tune_mem_pool(device);
    try {
        const uint8_t* p = static_cast<const uint8_t*>(file_bytes);
        if (kind == KIND_PEF) {
            ix->pef.reset(new PefIndexHost);
            std::string err;
            int rc = ix->pef->load(p, nbytes, codec, err);
            if (rc != 0) return fail(rc, err);
            ix->size = ix->pef->size; ix->num_docs = ix->pef->num_docs; ix->device_bytes = ix->pef->device_bytes;
            // what the block-parallel query kernels see: the window directory in place of (block_max, endpoint)
            ix->dev.lists = nullptr; ix->dev.dir = ix->pef->d_dir; ix->dev.bdir = ix->pef->d_bdir; ix->dev.bfirst = ix->pef->d_bfirst;
            ix->dev.num_lists = ix->size; ix->dev.num_docs = uint32_t(ix->num_docs); ix->dev.codec = CODEC_PEF;
            ix->dev.pdocs = ix->pef->dev.docs; ix->dev.pfreqs = ix->pef->dev.freqs;
            *out = ix.release();
            return DS2I_OK;
        }
        block_index_file f = parse_block_index(p, nbytes);
        ix->size = f.size; ix->num_docs = f.num_docs;
        // list directory: start offsets come from the Elias-Fano endpoints (block_freq_index.hpp:85-94)
        std::vector<uint64_t> starts = ef_decode_all(f.endpoints, 0, f.lists_bytes, f.size, f.params);
        ix->host_dir.resize(f.size);
        const uint8_t* lists_end = f.lists + f.lists_bytes;
        std::vector<uint2> bdir;
        std::vector<uint32_t> bfirst(f.size);
        for (uint64_t i = 0; i < f.size; ++i) {
            uint64_t begin = starts[i], end = (i + 1 < f.size) ? starts[i + 1] : f.lists_bytes;
            if (begin >= end || end > f.lists_bytes) throw format_error("list endpoints out of order");
            uint32_t n;
            const uint8_t* q = tight_vbyte_decode(f.lists + begin, lists_end, &n);
            if (!n) throw format_error("empty posting list");
            uint64_t blocks = (uint64_t(n) + BLOCK - 1) / BLOCK;
            uint64_t maxs_off = uint64_t(q - f.lists);
            uint64_t data_off = maxs_off + 4 * blocks + 4 * (blocks - 1);
            if (data_off > end) throw format_error("posting list header exceeds the list");
            if (end - data_off > 0xffffffffull) return fail(DS2I_E_LIMIT, "posting list " + std::to_string(i) + " holds 4 GB or more of block data");
            ix->host_dir[i] = ListDir{maxs_off, n, uint32_t(end - data_off)};
            if (bdir.size() + blocks > 0xfffffff0ull) return fail(DS2I_E_LIMIT, "more than 2^32 blocks in the index");
            bfirst[i] = uint32_t(bdir.size());
            const uint8_t* maxs = f.lists + maxs_off;
            const uint8_t* ends = maxs + 4 * blocks;
            uint32_t e_prev = 0;
            for (uint64_t b = 0; b < blocks; ++b) {
                uint32_t m, e = uint32_t(end - data_off);
                memcpy(&m, maxs + 4 * b, 4);
                if (b + 1 < blocks) memcpy(&e, ends + 4 * b, 4);
                // the kernels stage every [docs | freqs] block pair as one 16-B aligned window of at most STAGE_BYTES
                // (and_stage, stage_range): endpoints that run backwards, past the list, or a pair that cannot fit the
                // window would be decoded from stale shared memory, so they are refused here
                if (e < e_prev || e > uint32_t(end - data_off)) throw format_error("block endpoints of list " + std::to_string(i) + " are not monotone");
                if (uint64_t(e - e_prev) + 30 > STAGE_BYTES)
                    return fail(DS2I_E_LIMIT, "block pair of " + std::to_string(e - e_prev) + " bytes in list " + std::to_string(i) + " exceeds the staging window");
                e_prev = e;
                bdir.push_back(make_uint2(m, e));
            }
        }
        bdir.push_back(make_uint2(0xffffffffu, 0u));   // probes may read one entry past the last list
        const size_t pad = 4096;   // staged windows and unaligned reads may run past the last list
        CUDA_TRY(ix->d_lists.alloc(f.lists_bytes + pad));
        CUDA_TRY(cudaMemcpy(ix->d_lists.p, f.lists, f.lists_bytes, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemset(ix->d_lists.p + f.lists_bytes, 0, pad));
        CUDA_TRY(ix->d_dir.upload(ix->host_dir));
        CUDA_TRY(ix->d_bdir.upload(bdir)); CUDA_TRY(ix->d_bfirst.upload(bfirst));
        CUDA_TRY(cudaStreamSynchronize(0));     // the host vectors go out of scope
        ix->dev.lists = ix->d_lists.p; ix->dev.dir = ix->d_dir.p; ix->dev.bdir = ix->d_bdir.p; ix->dev.bfirst = ix->d_bfirst.p;
        ix->dev.num_lists = f.size; ix->dev.num_docs = uint32_t(f.num_docs); ix->dev.codec = codec;
        ix->device_bytes = f.lists_bytes + pad + f.size * (sizeof(ListDir) + 4) + bdir.size() * sizeof(uint2);
    } catch (std::exception const& e) {
        return fail(DS2I_E_FORMAT, e.what());
    }
    *out = ix.release();
    return DS2I_OK;
}

extern "C" int ds2i_gpu_index_open_file(const char* path, const char* index_type, int device, ds2i_gpu_index** out) {
    if (!path) return fail(DS2I_E_ARG, "null path");
    mapped_file m;
    if (!m.open(path)) return fail(DS2I_E_ARG, std::string("cannot open ") + path);
    return ds2i_gpu_index_open(m.p, m.n, index_type, device, out);
}

extern "C" void ds2i_gpu_index_close(ds2i_gpu_index* ix) { if (ix) { device_scope ds(ix->device); delete ix; } }
extern "C" uint64_t ds2i_gpu_index_size(const ds2i_gpu_index* ix) { return ix ? ix->size : 0; }
extern "C" uint64_t ds2i_gpu_index_num_docs(const ds2i_gpu_index* ix) { return ix ? ix->num_docs : 0; }
extern "C" uint64_t ds2i_gpu_index_device_bytes(const ds2i_gpu_index* ix) { return ix ? ix->device_bytes : 0; }

static inline uint64_t list_size_of(const ds2i_gpu_index* ix, uint32_t term) {
    return ix->kind == KIND_PEF ? ix->pef->host_dir[term].n : ix->host_dir[term].n;
}
// 128-posting blocks of a list as the block-parallel kernels count them (Elias-Fano lists: windows never straddle partitions)
static inline uint64_t list_blocks_of(const ds2i_gpu_index* ix, uint32_t term) {
    return ix->kind == KIND_PEF ? ix->pef->host_dir[term].blocks : (uint64_t(ix->host_dir[term].n) + BLOCK - 1) / BLOCK;
}

extern "C" int ds2i_gpu_index_set_global_stats(ds2i_gpu_index* ix, const uint64_t* df, size_t nterms, uint64_t num_docs_total) {
    if (!ix) return fail(DS2I_E_ARG, "null index");
    if (!df) { ix->g_df.clear(); ix->g_num_docs = 0; return DS2I_OK; }          // back to the index's own statistics
    if (nterms != ix->size) return fail(DS2I_E_ARG, "one document frequency per list of the index expected");
    if (num_docs_total < ix->num_docs) return fail(DS2I_E_ARG, "the collection cannot be smaller than its shard");
    for (size_t i = 0; i < nterms; ++i)
        if (df[i] < list_size_of(ix, uint32_t(i)) || df[i] > num_docs_total) return fail(DS2I_E_ARG, "document frequency out of range for list " + std::to_string(i));
    ix->g_df.assign(df, df + nterms);
    ix->g_num_docs = num_docs_total;
    return DS2I_OK;
}

extern "C" int ds2i_gpu_index_list_sizes(const ds2i_gpu_index* ix, const uint32_t* terms, size_t nterms, uint64_t* out_sizes) {
    if (!ix || (!terms && nterms) || (!out_sizes && nterms)) return fail(DS2I_E_ARG, "null argument");
    for (size_t i = 0; i < nterms; ++i) {
        if (terms[i] >= ix->size) return fail(DS2I_E_ARG, "term id out of range");
        out_sizes[i] = list_size_of(ix, terms[i]);
    }
    return DS2I_OK;
}

extern "C" int ds2i_gpu_index_list_bytes(const ds2i_gpu_index* ix, const uint32_t* terms, size_t nterms, uint64_t* out_bytes) {
    if (!ix || (!terms && nterms) || (!out_bytes && nterms)) return fail(DS2I_E_ARG, "null argument");
    for (size_t i = 0; i < nterms; ++i) {
        if (terms[i] >= ix->size) return fail(DS2I_E_ARG, "term id out of range");
        if (ix->kind == KIND_PEF) out_bytes[i] = (ix->pef->host_dir[terms[i]].bits + 7) / 8;
        else {
            ListDir const& d = ix->host_dir[terms[i]];
            const uint64_t blocks = (uint64_t(d.n) + BLOCK - 1) / BLOCK;
            out_bytes[i] = 4 * blocks + 4 * (blocks - 1) + d.data_bytes;          // block_maxs + block_endpoints + block data
        }
    }
    return DS2I_OK;
}

extern "C" int ds2i_gpu_wand_open(const void* file_bytes, size_t nbytes, int device, ds2i_gpu_wand** out) {
    if (!file_bytes || !out) return fail(DS2I_E_ARG, "null argument");
    device_scope on_device(device);
    CUDA_TRY(on_device.status);
    std::unique_ptr<ds2i_gpu_wand> w(new ds2i_gpu_wand);
    w->device = device;
    try {
        wand_file f = parse_wand(static_cast<const uint8_t*>(file_bytes), nbytes);
        w->num_docs = f.num_docs; w->num_terms = f.num_terms;
        w->h_max_term_weight.resize(f.num_terms);
        if (f.num_terms) memcpy(w->h_max_term_weight.data(), f.max_term_weight, f.num_terms * 4);
        CUDA_TRY(w->d_norm_lens.alloc(f.num_docs));
        if (f.num_docs) CUDA_TRY(cudaMemcpy(w->d_norm_lens.p, f.norm_lens, f.num_docs * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(w->d_max_term_weight.upload(w->h_max_term_weight));
        w->dev.norm_lens = w->d_norm_lens.p; w->dev.max_term_weight = w->d_max_term_weight.p;
        float mn = f.num_docs ? std::numeric_limits<float>::max() : 0.f;
        for (uint64_t i = 0; i < f.num_docs; ++i) { float v; memcpy(&v, f.norm_lens + 4 * i, 4); mn = std::min(mn, v); }
        w->dev.min_norm_len = mn;
    } catch (std::exception const& e) {
        return fail(DS2I_E_FORMAT, e.what());
    }
    *out = w.release();
    return DS2I_OK;
}

extern "C" int ds2i_gpu_wand_open_file(const char* path, int device, ds2i_gpu_wand** out) {
    if (!path) return fail(DS2I_E_ARG, "null path");
    mapped_file m;
    if (!m.open(path)) return fail(DS2I_E_ARG, std::string("cannot open ") + path);
    return ds2i_gpu_wand_open(m.p, m.n, device, out);
}
extern "C" void ds2i_gpu_wand_close(ds2i_gpu_wand* w) { if (w) { device_scope ds(w->device); delete w; } }

// ------------------------------------------------------------------------------------------------
// bm25::query_term_weight (bm25.hpp:17-24), evaluated on the host with the same libm as the reference
static float query_term_weight(uint64_t freq, uint64_t df, uint64_t num_docs) {
    float f = float(freq);
    float fdf = float(df);
    float idf = std::log((float(num_docs) - fdf + 0.5f) / (fdf + 0.5f));
    static const float epsilon_score = 1.0E-6f;
    return f * std::max(epsilon_score, idf) * (1.0f + 1.2f);
}

static double now_ms() {
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
static bool trace_on() { static const bool t = getenv("DS2I_GPU_TRACE") != nullptr; return t; }

// cudaFuncSetAttribute + the occupancy query are not free: do them once per (kernel, shared-memory size)
// (and per DEVICE: the opt-in shared-memory limit and the occupancy are per-device state of the current context)
static int cached_blocks_per_sm(const void* kern, int threads, size_t smem, int* out) {
    struct key { const void* k; size_t s; int dev; int v; };
    static std::vector<key> cache;
    static std::mutex mu;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    for (auto const& c : cache) if (c.k == kern && c.s == smem && c.dev == dev) { *out = c.v; return DS2I_OK; }
    // the opt-in limit is per-function state and the last call wins: only ever raise it, or a batch with few terms
    // would lower it under a later batch that needs the larger window again
    size_t raised = 0;
    for (auto const& c : cache) if (c.k == kern && c.dev == dev) raised = std::max(raised, c.s);
    if (smem > raised) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
#ifdef DS2I_CARVEOUT_MAX
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
#endif
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    cache.push_back(key{kern, smem, dev, per_sm});
    *out = per_sm;
    return DS2I_OK;
}

// which: bit 0 = build the work items of the conjunctive path, bit 1 = of the union path
static int batch_prepare_impl(ds2i_gpu_index* ix, ds2i_gpu_wand* wand, const uint32_t* terms,
                              const uint64_t* query_offsets, size_t nq, unsigned which, ds2i_gpu_batch** out);

extern "C" int ds2i_gpu_batch_prepare(ds2i_gpu_index* ix, ds2i_gpu_wand* wand, const uint32_t* terms,
                                      const uint64_t* query_offsets, size_t nq, ds2i_gpu_batch** out) {
    return batch_prepare_impl(ix, wand, terms, query_offsets, nq, 3u, out);
}

// Persistent host threads for the per-query preparation: spawning std::threads per batch cost more than the work they did
// (1.1 ms for a 10k-query batch on 16 threads, against 2.2 ms single-threaded).  run(n, fn) calls fn(0) .. fn(n-1), fn(0) on
// the caller; it returns when all are done.  One batch is prepared at a time per process (the pool is a shared resource).
class host_pool {
public:
    static host_pool& get() { static host_pool p; return p; }
    unsigned size() const { return unsigned(m_threads.size()) + 1; }
    template <typename F>
    void run(unsigned n, F&& fn) {
        if (n <= 1) { if (n) fn(0u); return; }
        std::unique_lock<std::mutex> user(m_user);              // one parallel region at a time
        {
            std::lock_guard<std::mutex> lk(m_mu);
            m_fn = [&fn](unsigned i) { fn(i); };
            m_n = n; m_next = 1; m_done = 1; ++m_epoch;          // index 0 runs on the caller
        }
        m_cv.notify_all();
        fn(0u);
        work();
        std::unique_lock<std::mutex> lk(m_mu);
        m_cv_done.wait(lk, [&] { return m_done >= m_n + 0u && m_active == 0; });
        m_fn = nullptr;
    }
private:
    host_pool() {
        unsigned n = std::thread::hardware_concurrency();
        if (const char* ev = getenv("DS2I_GPU_HOST_THREADS")) n = unsigned(std::max(1, atoi(ev)));
        n = std::max(1u, std::min(n, 16u));
        for (unsigned t = 1; t < n; ++t) m_threads.emplace_back([this] { loop(); });
    }
    ~host_pool() {
        { std::lock_guard<std::mutex> lk(m_mu); m_stop = true; }
        m_cv.notify_all();
        for (auto& t : m_threads) t.join();
    }
    void work() {
        while (true) {
            unsigned i;
            std::function<void(unsigned)> fn;
            {
                std::lock_guard<std::mutex> lk(m_mu);
                if (m_next >= m_n) return;
                i = m_next++; fn = m_fn;
            }
            fn(i);
            std::lock_guard<std::mutex> lk(m_mu);
            if (++m_done >= m_n) m_cv_done.notify_all();
        }
    }
    void loop() {
        uint64_t seen = 0;
        while (true) {
            {
                std::unique_lock<std::mutex> lk(m_mu);
                m_cv.wait(lk, [&] { return m_stop || m_epoch != seen; });
                if (m_stop) return;
                seen = m_epoch; ++m_active;
            }
            work();
            std::lock_guard<std::mutex> lk(m_mu);
            --m_active;
            m_cv_done.notify_all();
        }
    }
    std::vector<std::thread> m_threads;
    std::mutex m_mu, m_user;
    std::condition_variable m_cv, m_cv_done;
    std::function<void(unsigned)> m_fn;
    unsigned m_n = 0, m_next = 0, m_done = 0, m_active = 0;
    uint64_t m_epoch = 0;
    bool m_stop = false;
};

// number of host threads for the per-query preparation (DS2I_GPU_HOST_THREADS overrides; small batches stay single-threaded)
static unsigned prepare_threads(size_t nq) {
    return unsigned(std::max<size_t>(1, std::min<size_t>(host_pool::get().size(), nq / 512)));
}

#ifndef DS2I_AND_ITEM_PROBES
#define DS2I_AND_ITEM_PROBES 128
#endif
#ifndef DS2I_UNION_ITEM_PROBES
#define DS2I_UNION_ITEM_PROBES 2048
#endif
constexpr uint64_t AND_ITEM_PROBES = DS2I_AND_ITEM_PROBES;
constexpr uint64_t UNION_ITEM_PROBES = DS2I_UNION_ITEM_PROBES;
static uint32_t and_chunk_max() {        // DS2I_GPU_AND_CHUNK_BLOCKS caps the blocks per item (<= 32)
    static const uint32_t v = [] {
        uint32_t c = AND_CHUNK_BLOCKS;
        if (const char* ev = getenv("DS2I_GPU_AND_CHUNK_BLOCKS")) c = std::min<uint32_t>(32, std::max<uint32_t>(1, uint32_t(atoi(ev))));
        return c;
    }();
    return v;
}

// what one host thread produces for its contiguous range of queries
struct prep_part {
    std::vector<uint32_t> term, nt;          // distinct terms (query_freqs order); distinct terms per query
    std::vector<float> q_weight, max_weight;
    std::vector<uint8_t> ord_size, ord_maxw;
    std::vector<uint64_t> cost, shortest;
    std::vector<uint8_t> chunk;               // blocks of the driving list per conjunctive work item
    int max_terms = 1;
    int rc = DS2I_OK;
    std::string err;
};

static void prepare_range(const ds2i_gpu_index* ix, const ds2i_gpu_wand* wand, const uint32_t* terms, const uint64_t* query_offsets,
                          size_t q0, size_t q1, prep_part& out) {
    struct ent { uint64_t n; float mw; uint8_t pos; uint64_t local_n; uint32_t term; };   // n: the list size the reference would see (collection-wide df for a shard)
    std::vector<uint32_t> tmp;
    ent ents[MAX_TERMS], by_size[MAX_TERMS], by_mw[MAX_TERMS];
    out.nt.reserve(q1 - q0); out.cost.reserve(q1 - q0); out.shortest.reserve(q1 - q0); out.chunk.reserve(q1 - q0);
    {
        const size_t nt_max = q1 > q0 && query_offsets[q1] >= query_offsets[q0] ? size_t(query_offsets[q1] - query_offsets[q0]) : 0;
        out.term.reserve(nt_max); out.q_weight.reserve(nt_max); out.max_weight.reserve(nt_max); out.ord_size.reserve(nt_max); out.ord_maxw.reserve(nt_max);
    }
    for (size_t q = q0; q < q1; ++q) {
        if (query_offsets[q + 1] < query_offsets[q]) { out.rc = DS2I_E_ARG; out.err = "query_offsets not monotone"; return; }
        tmp.assign(terms + query_offsets[q], terms + query_offsets[q + 1]);
        std::sort(tmp.begin(), tmp.end());               // query_freqs (queries.hpp:136-150)
        uint32_t ne = 0;
        uint64_t cost = 0;
        for (size_t i = 0; i < tmp.size();) {
            size_t j = i;
            while (j < tmp.size() && tmp[j] == tmp[i]) ++j;
            uint32_t t = tmp[i];
            if (t >= ix->size) { out.rc = DS2I_E_ARG; out.err = "term id out of range in query " + std::to_string(q); return; }
            if (wand && t >= wand->num_terms) { out.rc = DS2I_E_ARG; out.err = "term id beyond wand data"; return; }
            const uint64_t local_n = list_size_of(ix, t);
            // a document-partitioned shard scores with the statistics of the whole collection (ds2i_gpu_index_set_global_stats)
            const uint64_t n = ix->g_df.empty() ? local_n : ix->g_df[t];
            float qw = query_term_weight(j - i, n, ix->g_num_docs ? ix->g_num_docs : ix->num_docs);
            float mw = wand ? qw * wand->h_max_term_weight[t] : 0.f;
            if (ne >= uint32_t(MAX_TERMS)) {
                out.rc = DS2I_E_LIMIT;
                out.err = "query " + std::to_string(q) + " has more than " + std::to_string(MAX_TERMS) + " distinct terms";
                return;
            }
            ents[ne] = ent{n, mw, uint8_t(ne), local_n, t};
            ++ne;
            out.term.push_back(t); out.q_weight.push_back(qw); out.max_weight.push_back(mw);
            cost += local_n;
            i = j;
        }
        out.max_terms = std::max<int>(out.max_terms, int(ne));
        out.nt.push_back(ne);
        // the reference's own std::sort calls, on the same keys in the same initial order (queries.hpp:357-360, 521-524)
        std::copy(ents, ents + ne, by_size); std::copy(ents, ents + ne, by_mw);
        std::sort(by_size, by_size + ne, [](ent const& l, ent const& r) { return l.n < r.n; });
        std::sort(by_mw, by_mw + ne, [](ent const& l, ent const& r) { return l.mw < r.mw; });
        for (uint32_t i = 0; i < ne; ++i) { out.ord_size.push_back(by_size[i].pos); out.ord_maxw.push_back(by_mw[i].pos); }
        out.cost.push_back(cost);
        out.shortest.push_back(ne ? list_blocks_of(ix, by_size[0].term) : 0);            // blocks of the driving list
        // Work per block of the driving list ~ 1 + the blocks of every other list its 128 candidates can fall into.  An item
        // is cut so that it holds about AND_ITEM_PROBES block probes: 32 blocks when the lists are of similar length, a few
        // when a mid-sized list is intersected with huge ones (such items ran for milliseconds and were the tail of a batch).
        {
            uint64_t per_block = 1;
            const uint64_t nb0 = ne ? std::max<uint64_t>(1, list_blocks_of(ix, by_size[0].term)) : 1;
            for (uint32_t i = 1; i < ne; ++i) per_block += std::min<uint64_t>(BLOCK, std::max<uint64_t>(1, list_blocks_of(ix, by_size[i].term) / nb0));
            out.chunk.push_back(uint8_t(std::min<uint64_t>(and_chunk_max(), std::max<uint64_t>(1, AND_ITEM_PROBES / per_block))));
        }
    }
}

// bump allocator over the staging / device arena: every array starts 256-B aligned
struct arena_layout {
    size_t bytes = 0;
    size_t take(size_t n) { size_t o = bytes; bytes = (bytes + n + 255) & ~size_t(255); return o; }
};

static int batch_prepare_impl(ds2i_gpu_index* ix, ds2i_gpu_wand* wand, const uint32_t* terms,
                              const uint64_t* query_offsets, size_t nq, unsigned which, ds2i_gpu_batch** out) {
    if (!ix || !out || !query_offsets || (!terms && nq && query_offsets[nq])) return fail(DS2I_E_ARG, "null argument");
    if (nq > 0x7fffffffull) return fail(DS2I_E_LIMIT, "too many queries in one batch");
    if (wand && wand->num_docs < ix->num_docs) return fail(DS2I_E_ARG, "wand data has fewer documents than the index");
    device_scope on_device(ix->device);
    CUDA_TRY(on_device.status);
    std::unique_ptr<ds2i_gpu_batch> b(new ds2i_gpu_batch);
    b->index = ix; b->wand = wand; b->nq = uint32_t(nq); b->items_built = which;
    const double tp0 = now_ms();

    // ---- per-query host work, on several threads (contiguous ranges, so the parts concatenate in query order)
    const unsigned nthreads = prepare_threads(nq);
    std::vector<prep_part> parts(nthreads);
    host_pool::get().run(nthreads, [&](unsigned t) {
        prepare_range(ix, wand, terms, query_offsets, nq * t / nthreads, nq * (t + 1) / nthreads, parts[t]);
    });
    size_t nterms_total = 0;
    int max_terms = 1;
    for (auto const& pt : parts) {
        if (pt.rc != DS2I_OK) return fail(pt.rc, pt.err);
        nterms_total += pt.term.size();
        max_terms = std::max(max_terms, pt.max_terms);
    }
    if (nterms_total > 0x7fffffffull) return fail(DS2I_E_LIMIT, "too many query terms in one batch");
    b->max_terms = max_terms;
    const size_t T = nterms_total;
    const size_t G = (which & 2u) ? T : 0;            // union path: one group per query term

    // ---- layout of the uploaded part, then of the device-only part
    arena_layout lay;
    const size_t o_q_begin = lay.take((nq + 1) * 4), o_term = lay.take(T * 4), o_sched = lay.take(nq * 4), o_qw = lay.take(T * 4),
                 o_mw = lay.take(T * 4), o_os = lay.take(T), o_om = lay.take(T), o_and_gstart = lay.take((nq + 1) * 4),
                 o_and_begin = lay.take((nq + 1) * 4), o_un_gstart = lay.take((G + 1) * 4), o_un_gterm = lay.take(G * 4),
                 o_un_gquery = lay.take(G * 4), o_un_gbase = lay.take(G * 4), o_un_begin = lay.take((nq + 1) * 4), o_un_ub = lay.take(T * 4),
                 o_qchunk = lay.take(nq), o_un_gblocks = lay.take(G * 4);
    const size_t upload_bytes = lay.bytes;

    std::lock_guard<std::mutex> staging_lock(ix->staging_mu);
    CUDA_TRY(ix->staging.reserve(upload_bytes));
    uint8_t* h = ix->staging.p;
    uint32_t* q_begin = reinterpret_cast<uint32_t*>(h + o_q_begin);
    uint32_t* term = reinterpret_cast<uint32_t*>(h + o_term);
    uint32_t* sched = reinterpret_cast<uint32_t*>(h + o_sched);
    float* q_weight = reinterpret_cast<float*>(h + o_qw);
    float* max_weight = reinterpret_cast<float*>(h + o_mw);
    uint8_t* ord_size = h + o_os;
    uint8_t* ord_maxw = h + o_om;
    std::vector<uint64_t> cost(nq), shortest(nq);
    uint8_t* qchunk = h + o_qchunk;
    {
        size_t q = 0, t = 0;
        q_begin[0] = 0;
        for (auto const& pt : parts) {
            if (!pt.term.empty()) {
                memcpy(term + t, pt.term.data(), pt.term.size() * 4);
                memcpy(q_weight + t, pt.q_weight.data(), pt.term.size() * 4);
                memcpy(max_weight + t, pt.max_weight.data(), pt.term.size() * 4);
                memcpy(ord_size + t, pt.ord_size.data(), pt.term.size());
                memcpy(ord_maxw + t, pt.ord_maxw.data(), pt.term.size());
            }
            for (size_t i = 0; i < pt.nt.size(); ++i, ++q) {
                t += pt.nt[i];
                q_begin[q + 1] = uint32_t(t);
                cost[q] = pt.cost[i]; shortest[q] = pt.shortest[i]; qchunk[q] = pt.chunk[i];
            }
        }
    }
    // processing order: costliest queries first.  The order only steers the scheduler (longest work first), so a stable
    // counting sort on a 7-bit logarithmic key (exponent + two mantissa bits of the cost) replaces the comparison sort
    // that used to be a third of the host time of a 10k-query batch.
    {
        auto key_of = [](uint64_t c) -> uint32_t {
            if (c < 4) return uint32_t(c);
            const uint32_t e = 63u - uint32_t(__builtin_clzll(c));
            return 4u * (e - 1u) + uint32_t((c >> (e - 2u)) & 3u);          // monotone in c, < 256
        };
        uint32_t count[257] = {0};
        for (size_t q = 0; q < nq; ++q) count[256u - key_of(cost[q])] += 1;      // descending
        uint32_t pos = 0;
        for (uint32_t kx = 0; kx <= 256; ++kx) { const uint32_t c = count[kx]; count[kx] = pos; pos += c; }
        for (size_t q = 0; q < nq; ++q) sched[count[256u - key_of(cost[q])]++] = uint32_t(q);
    }

    const double tp1 = now_ms();
    // work items of the conjunctive path: query q owns ceil(blocks of its shortest list / qchunk[q]) items
    uint32_t* item_begin = reinterpret_cast<uint32_t*>(h + o_and_begin);
    uint32_t* and_gstart = reinterpret_cast<uint32_t*>(h + o_and_gstart);
    {
        uint64_t nitems = 0;
        item_begin[0] = 0; and_gstart[0] = 0;
        for (size_t q = 0; q < nq; ++q) {
            if (which & 1u) nitems += (shortest[q] + qchunk[q] - 1) / qchunk[q];
            if (nitems > 0x7fffffffull) return fail(DS2I_E_LIMIT, "too many work items in one batch");
            item_begin[q + 1] = uint32_t(nitems);
        }
        // the items of the costliest queries come first
        for (size_t p = 0; p < nq; ++p) and_gstart[p + 1] = and_gstart[p] + (item_begin[sched[p] + 1] - item_begin[sched[p]]);
        b->n_and_items = uint32_t(nitems);
    }

    // work items of the union path: (query, driving list, run of its blocks), highest-weight lists first.  Only the
    // groups (one per query term) are materialised; the kernel derives the items from the prefix array.  (A batch prepared
    // for the conjunctive operators alone, which == 1, is only ever run with those: it needs neither the items nor the
    // upper-bound prefix sums of wand / maxscore.)
    if (which != 1u) {
        float* ub = reinterpret_cast<float*>(h + o_un_ub);
        uint32_t* ubegin = reinterpret_cast<uint32_t*>(h + o_un_begin);
        uint32_t* gstart = reinterpret_cast<uint32_t*>(h + o_un_gstart);
        uint32_t* gterm = reinterpret_cast<uint32_t*>(h + o_un_gterm);
        uint32_t* gquery = reinterpret_cast<uint32_t*>(h + o_un_gquery);
        uint32_t* gres = reinterpret_cast<uint32_t*>(h + o_un_gbase);
        uint32_t* gblocks_out = reinterpret_cast<uint32_t*>(h + o_un_gblocks);
        std::vector<uint32_t> gitem(T, 0);
        std::vector<uint32_t> gbase(T, 0), gchunks(T, 0);
        // postings per work item; DS2I_GPU_UNION_ITEM_POSTINGS overrides it (tests use a tiny value to
        // exercise the splitting path on small collections)
        uint64_t per_item = 8192;
        if (const char* ev = getenv("DS2I_GPU_UNION_ITEM_POSTINGS")) per_item = std::max<uint64_t>(1, strtoull(ev, nullptr, 10));
        const uint32_t item_blocks = uint32_t(std::min<uint64_t>(1u << 20, std::max<uint64_t>(1, per_item / BLOCK)));
        b->un_item_blocks = item_blocks;
        std::vector<uint32_t> level_count(MAX_TERMS + 1, 0);
        // The size of an item follows the size of the batch: about UNION_ITEM_PROBES block probes per item when there is
        // plenty of work (fewer, fatter items: most items of a pruned list return at once and each costs a warp a visit),
        // down to an eighth of that when the whole batch would otherwise be a few thousand items (a strong-scaling shard:
        // there the longest item is the run time).  Measured on B200: 10k queries 30.7 ms at 2048 vs 33.4 at 512; 1250
        // queries 7.0 ms at 2048 vs 5.2 at 512.
        uint64_t union_probes = UNION_ITEM_PROBES;
        if (which & 2u) {
            uint64_t total = 0;
            for (size_t q = 0; q < nq; ++q) {
                const uint32_t t0 = q_begin[q], nt = q_begin[q + 1] - t0;
                for (uint32_t i = 0; i < nt; ++i) total += list_blocks_of(ix, term[t0 + i]) * uint64_t(nt);
            }
            const uint64_t want_items = uint64_t(ix->sm_count) * 28 * 16;                     // ~16 items per resident warp
            union_probes = std::min<uint64_t>(UNION_ITEM_PROBES, std::max<uint64_t>(UNION_ITEM_PROBES / 8, total / want_items));
        }
        uint64_t nitems = 0;
        ubegin[0] = 0;
        for (size_t q = 0; q < nq; ++q) {
            const uint32_t t0 = q_begin[q], nt = q_begin[q + 1] - t0;
            float acc = 0.f;
            for (uint32_t i = 0; i < nt; ++i) {          // queries.hpp:526-530, same sequential fp32 sum; slot i = i-th list by max_weight
                const float mw = max_weight[t0 + ord_maxw[t0 + i]];
                acc = i ? acc + mw : mw;
                ub[t0 + i] = acc;
                if (!(which & 2u)) continue;
                const uint64_t nb = list_blocks_of(ix, term[t0 + ord_maxw[t0 + i]]);
                // blocks per item of this (query, list) group: ~UNION_ITEM_PROBES block probes (ownership probes of the lists
                // above, completion probes of the lists below), at most item_blocks
                uint64_t per_block = 1;
                for (uint32_t j = 0; j < nt; ++j)
                    if (j != i) per_block += std::min<uint64_t>(BLOCK, std::max<uint64_t>(1, list_blocks_of(ix, term[t0 + ord_maxw[t0 + j]]) / std::max<uint64_t>(nb, 1)));
                const uint32_t ib = uint32_t(std::min<uint64_t>(item_blocks, std::max<uint64_t>(1, union_probes / per_block)));
                gitem[t0 + i] = ib;
                gchunks[t0 + i] = uint32_t((nb + ib - 1) / ib);
                gbase[t0 + i] = uint32_t(nitems);
                nitems += gchunks[t0 + i];
                level_count[nt - 1 - i] += 1;
            }
            if (nitems > 0x7fffffffull) return fail(DS2I_E_LIMIT, "too many work items in one batch");
            ubegin[q + 1] = uint32_t(nitems);
        }
        // processing order: level by level (level 0 = each query's highest-weight list), costliest queries first inside a level
        gstart[0] = 0;
        if (G) {
            std::vector<uint32_t> level_pos(MAX_TERMS + 1, 0);
            for (int l = 1; l <= MAX_TERMS; ++l) level_pos[l] = level_pos[l - 1] + level_count[l - 1];
            for (size_t p = 0; p < nq; ++p) {
                const uint32_t qi = sched[p];
                const uint32_t t0 = q_begin[qi], nt = q_begin[qi + 1] - t0;
                for (uint32_t i = 0; i < nt; ++i) {
                    const uint32_t pos = level_pos[nt - 1 - i]++;
                    gterm[pos] = t0 + i; gquery[pos] = qi; gres[pos] = gbase[t0 + i]; gstart[pos + 1] = gchunks[t0 + i]; gblocks_out[pos] = gitem[t0 + i];
                }
            }
            for (size_t g = 0; g < G; ++g) gstart[g + 1] += gstart[g];
        }
        b->n_un_items = uint32_t(nitems); b->n_un_groups = uint32_t(G);
    }
    const double tp2 = now_ms();

    // ---- device arena: the uploaded prefix + the device-only buffers, one allocation, one H2D copy
    const size_t o_counter = lay.take(16), o_stats = lay.take(16 * 8), o_and_counts = lay.take(size_t(b->n_and_items) * 4),
                 o_and_sizes = lay.take(size_t(b->n_and_items) * 4), o_un_sizes = lay.take(size_t(b->n_un_items) * 4),
                 o_un_thr = lay.take(std::max<size_t>(nq, 1) * 4), o_fused = lay.take(nq * (8 + 8 * size_t(MAX_K)));
    CUDA_TRY(b->arena.alloc(lay.bytes));
    uint8_t* d = b->arena.p;
    if (upload_bytes) CUDA_TRY(cudaMemcpyAsync(d, h, upload_bytes, cudaMemcpyHostToDevice, 0));
    b->counters_bytes = o_and_counts - o_counter;
    CUDA_TRY(cudaMemsetAsync(d + o_counter, 0, b->counters_bytes, 0));          // work counters + stats
    b->q_begin.p = reinterpret_cast<uint32_t*>(d + o_q_begin); b->term.p = reinterpret_cast<uint32_t*>(d + o_term);
    b->sched.p = reinterpret_cast<uint32_t*>(d + o_sched); b->q_weight.p = reinterpret_cast<float*>(d + o_qw);
    b->max_weight.p = reinterpret_cast<float*>(d + o_mw); b->ord_size.p = d + o_os; b->ord_maxw.p = d + o_om;
    b->and_gstart.p = reinterpret_cast<uint32_t*>(d + o_and_gstart); b->and_item_begin.p = reinterpret_cast<uint32_t*>(d + o_and_begin);
    b->un_gstart.p = reinterpret_cast<uint32_t*>(d + o_un_gstart); b->un_gterm.p = reinterpret_cast<uint32_t*>(d + o_un_gterm);
    b->un_gquery.p = reinterpret_cast<uint32_t*>(d + o_un_gquery); b->un_gbase.p = reinterpret_cast<uint32_t*>(d + o_un_gbase);
    b->un_item_begin.p = reinterpret_cast<uint32_t*>(d + o_un_begin); b->un_ub.p = reinterpret_cast<float*>(d + o_un_ub);
    b->and_qchunk.p = d + o_qchunk; b->un_gblocks.p = reinterpret_cast<uint32_t*>(d + o_un_gblocks);
    b->work_counter.p = reinterpret_cast<uint32_t*>(d + o_counter); b->stats.p = reinterpret_cast<unsigned long long*>(d + o_stats);
    b->and_item_counts.p = reinterpret_cast<uint32_t*>(d + o_and_counts); b->and_item_sizes.p = reinterpret_cast<uint32_t*>(d + o_and_sizes);
    b->un_item_sizes.p = reinterpret_cast<uint32_t*>(d + o_un_sizes); b->un_threshold.p = reinterpret_cast<uint32_t*>(d + o_un_thr);
    b->out_fused.p = d + o_fused;
    b->layout_outputs(MAX_K);
    CUDA_TRY(cudaEventCreate(&b->ev0)); CUDA_TRY(cudaEventCreate(&b->ev1));
    // the staging memory is reused by the next batch of this index: the copy must have left it
    CUDA_TRY(cudaStreamSynchronize(0));
    if (trace_on()) fprintf(stderr, "[ds2i_gpu] prepare: per-query host work %.2f ms (%u threads), work items %.2f ms, allocation + upload of %zu bytes %.2f ms\n",
                            tp1 - tp0, nthreads, tp2 - tp1, upload_bytes, now_ms() - tp2);
    *out = b.release();
    return DS2I_OK;
}

template <int CODEC, int OP>
static int launch_query(ds2i_gpu_batch* b, DevBatch const& db, uint32_t k) {
    ds2i_gpu_index* ix = b->index;
    const int warps = 4;
    size_t smem = S16_TAB_BYTES + warps * warp_smem_bytes(b->max_terms);
    auto kern = query_kernel<CODEC, OP>;
    int per_sm = 0;
    int orc = cached_blocks_per_sm(reinterpret_cast<const void*>(kern), warps * 32, smem, &per_sm);
    if (orc != DS2I_OK) return orc;
    if (per_sm < 1) return fail(DS2I_E_CUDA, "query kernel does not fit on an SM");
    int grid = per_sm * ix->sm_count;
    int needed = int((b->nq + warps - 1) / warps);
    if (grid > needed) grid = std::max(needed, 1);
    DevWand dw = b->wand ? b->wand->dev : DevWand{nullptr, nullptr, 0.f};
    kern<<<grid, warps * 32, smem>>>(ix->dev, dw, db, k, b->max_terms);
    return DS2I_OK;
}

#ifndef DS2I_AND_MIN_CTAS
#define DS2I_AND_MIN_CTAS 7
#endif
constexpr int AND_MIN_CTAS = DS2I_AND_MIN_CTAS;     // 7: 28 resident warps per SM, <= 72 registers per thread (measured: 5 / 6 / 7 / 8 CTAs -> ranked_and 16.98 / 16.14 / 15.77 / 16.9 ms)
// the Elias-Fano window decoders carry a wide partition descriptor (PefBody): fewer, fatter warps beat spilling it
#ifndef DS2I_PEF_MIN_CTAS
#define DS2I_PEF_MIN_CTAS 4
#endif
constexpr int and_min_ctas(int codec) { return codec == CODEC_PEF ? DS2I_PEF_MIN_CTAS : AND_MIN_CTAS; }

template <int CODEC, bool RANKED, bool STATS = true, int MIN_CTAS = and_min_ctas(CODEC)>
static int launch_and_block(ds2i_gpu_batch* b, DevBatch const& db, uint32_t k) {
    ds2i_gpu_index* ix = b->index;
    const int warps = 4;
    auto kern = and_block_kernel<CODEC, RANKED, MIN_CTAS, STATS>;
    DevWand dw = b->wand ? b->wand->dev : DevWand{nullptr, nullptr, 0.f};
    if (b->n_and_items) {
        int slots = b->max_terms;
        if (const char* ev = getenv("DS2I_GPU_SLOTS_OVERRIDE")) slots = std::max(slots, atoi(ev));     // occupancy experiments
        size_t smem = S16_TAB_BYTES + warps * and_warp_smem_bytes(slots, CODEC == CODEC_PEF);
        int per_sm = 0;
        int orc = cached_blocks_per_sm(reinterpret_cast<const void*>(kern), warps * 32, smem, &per_sm);
        if (orc != DS2I_OK) return orc;
        if (per_sm < 1) return fail(DS2I_E_CUDA, "conjunctive kernel does not fit on an SM");
        int grid = per_sm * ix->sm_count;
        int needed = int((b->n_and_items + warps - 1) / warps);
        if (grid > needed) grid = std::max(needed, 1);
        if (RANKED && (b->and_item_scores_k < k || !b->and_item_scores.p)) {      // partial top-k lists: k floats per item
            CUDA_TRY(b->and_item_scores.alloc(size_t(b->n_and_items) * 2 * k));
            b->and_item_scores_k = k;
        }
        AndJob job{b->and_gstart.p, b->and_item_begin.p, b->n_and_items, b->and_qchunk.p, b->work_counter.p + 1, b->and_item_counts.p, b->and_item_sizes.p, b->and_item_scores.p};
        kern<<<grid, warps * 32, smem>>>(ix->dev, dw, db, job, k, slots);
        b->launches += 1;
    }
    merge_items_kernel<<<(b->nq + 3) / 4, 128>>>(b->and_item_begin.p, b->nq, b->and_item_counts.p, b->and_item_sizes.p, b->and_item_scores.p,
                                                 k, RANKED, b->out_counts.p, b->out_scores.p, b->out_docids.p);
    return DS2I_OK;   // the caller counts the merge launch
}

// the conjunctive kernel is instantiated per codec: the path is issue-bound and a codec-specific
// instance is a third smaller than one that dispatches at run time
// (no_stats: the instance without the work counters, built for the benchmarked index types only — DS2I_RUN_NO_STATS)
template <bool RANKED>
static int launch_and_block_codec(ds2i_gpu_batch* b, DevBatch const& db, uint32_t k, bool no_stats) {
    switch (b->index->codec) {
        case CODEC_OPTPFOR:
            if (no_stats) return launch_and_block<CODEC_OPTPFOR, RANKED, false>(b, db, k);
            return launch_and_block<CODEC_OPTPFOR, RANKED>(b, db, k);
#ifndef DS2I_DEV_FAST_BUILD
        case CODEC_VARINT: return launch_and_block<CODEC_VARINT, RANKED>(b, db, k);
        case CODEC_INTERPOLATIVE: return launch_and_block<CODEC_INTERPOLATIVE, RANKED>(b, db, k);
        case CODEC_QMX: return launch_and_block<CODEC_QMX, RANKED>(b, db, k);
        case CODEC_MIXED: return launch_and_block<CODEC_MIXED, RANKED>(b, db, k);
#endif
    }
    return fail(DS2I_E_UNSUPPORTED, "unknown codec");
}

#ifndef DS2I_UNION_MIN_CTAS
#define DS2I_UNION_MIN_CTAS 7
#endif
constexpr int UNION_MIN_CTAS = DS2I_UNION_MIN_CTAS;

template <int CODEC, bool STATS = true, int MODE = UNION_TOPK>
static int launch_union_block(ds2i_gpu_batch* b, DevBatch const& db, uint32_t k) {
    ds2i_gpu_index* ix = b->index;
    const int warps = 4;
    auto kern = union_drive_kernel<CODEC, CODEC == CODEC_PEF ? DS2I_PEF_MIN_CTAS : UNION_MIN_CTAS, STATS, MODE>;
    size_t smem = S16_TAB_BYTES + warps * union_warp_smem_bytes(b->max_terms, CODEC == CODEC_PEF);
    int per_sm = 0;
    int orc = cached_blocks_per_sm(reinterpret_cast<const void*>(kern), warps * 32, smem, &per_sm);
    if (orc != DS2I_OK) return orc;
    if (per_sm < 1) return fail(DS2I_E_CUDA, "union kernel does not fit on an SM");
    int grid = per_sm * ix->sm_count;
    int needed = int((b->n_un_items + warps - 1) / warps);
    if (grid > needed) grid = std::max(needed, 1);
    if (b->un_item_scores_k < k || !b->un_item_scores.p) {      // partial top-k lists: k floats per item
        CUDA_TRY(b->un_item_scores.alloc(size_t(b->n_un_items) * 2 * k));
        b->un_item_scores_k = k;
    }
    CUDA_TRY(cudaMemsetAsync(b->un_threshold.p, 0, std::max<size_t>(b->nq, 1) * sizeof(uint32_t)));
    UnionJob job{b->un_gstart.p, b->un_gterm.p, b->un_gquery.p, b->un_gbase.p, b->un_ub.p, b->un_gblocks.p, b->n_un_groups, b->n_un_items, b->work_counter.p + 3, b->un_threshold.p, b->un_item_sizes.p, b->un_item_scores.p};
    const DevWand dw = b->wand ? b->wand->dev : DevWand{nullptr, nullptr, 0.f};
    if (b->n_un_items) { kern<<<grid, warps * 32, smem>>>(ix->dev, dw, db, job, k, b->max_terms); b->launches += 1; }
    if (MODE == UNION_COUNT) merge_union_counts_kernel<<<(b->nq + 3) / 4, 128>>>(b->un_item_begin.p, b->nq, b->un_item_sizes.p, b->out_counts.p);
    else merge_union_items_kernel<<<(b->nq + 3) / 4, 128>>>(b->un_item_begin.p, b->nq, b->un_item_sizes.p, b->un_item_scores.p, k, b->out_counts.p, b->out_scores.p, b->out_docids.p);
    return DS2I_OK;
}

// or / ranked_or on the block-parallel union machinery (every index type)
template <int MODE>
static int launch_union_mode(ds2i_gpu_batch* b, DevBatch const& db, uint32_t k) {
    switch (b->index->kind == KIND_PEF ? int(CODEC_PEF) : b->index->codec) {      // (an Elias-Fano index keeps its variant in `codec`)
        case CODEC_OPTPFOR: return launch_union_block<CODEC_OPTPFOR, true, MODE>(b, db, k);
#ifndef DS2I_DEV_FAST_BUILD
        case CODEC_VARINT: return launch_union_block<CODEC_VARINT, true, MODE>(b, db, k);
        case CODEC_INTERPOLATIVE: return launch_union_block<CODEC_INTERPOLATIVE, true, MODE>(b, db, k);
        case CODEC_QMX: return launch_union_block<CODEC_QMX, true, MODE>(b, db, k);
        case CODEC_MIXED: return launch_union_block<CODEC_MIXED, true, MODE>(b, db, k);
#endif
        case CODEC_PEF: return launch_union_block<CODEC_PEF, true, MODE>(b, db, k);
    }
    return fail(DS2I_E_UNSUPPORTED, "unknown codec");
}
static int launch_union_exhaustive(ds2i_gpu_batch* b, DevBatch const& db, int op, uint32_t k) {
    return op == OP_OR ? launch_union_mode<UNION_COUNT>(b, db, k) : launch_union_mode<UNION_EXHAUSTIVE>(b, db, k);
}

template <int CODEC>
static int launch_query_op(ds2i_gpu_batch* b, DevBatch const& db, int op, uint32_t k) {
    switch (op) {
        case OP_AND: return launch_query<CODEC, OP_AND>(b, db, k);
#ifndef DS2I_DEV_FAST_BUILD
        case OP_AND_FREQ: return launch_query<CODEC, OP_AND_FREQ>(b, db, k);
        case OP_OR: return launch_query<CODEC, OP_OR>(b, db, k);
        case OP_OR_FREQ: return launch_query<CODEC, OP_OR_FREQ>(b, db, k);
        case OP_RANKED_AND: return launch_query<CODEC, OP_RANKED_AND>(b, db, k);
        case OP_WAND: return launch_query<CODEC, OP_WAND>(b, db, k);
        case OP_MAXSCORE: return launch_query<CODEC, OP_MAXSCORE>(b, db, k);
        case OP_RANKED_OR: return launch_query<CODEC, OP_RANKED_OR>(b, db, k);
#endif
    }
    return fail(DS2I_E_ARG, "unknown operator");
}

extern "C" int ds2i_gpu_batch_run(ds2i_gpu_batch* b, int op, uint32_t k, float* out_elapsed_ms) {
    return ds2i_gpu_batch_run_ex(b, op, k, 0u, out_elapsed_ms);
}

extern "C" int ds2i_gpu_batch_run_ex(ds2i_gpu_batch* b, int op, uint32_t k, uint32_t flags, float* out_elapsed_ms) {
    if (!b) return fail(DS2I_E_ARG, "null batch");
    if (op < 0 || op > OP_RANKED_OR) return fail(DS2I_E_ARG, "unknown operator");
    const bool ranked = op >= OP_RANKED_AND;
    if (ranked && !b->wand) return fail(DS2I_E_ARG, "ranked operators need wand data");
    if (ranked && (k < 1 || k > MAX_K)) return fail(DS2I_E_LIMIT, "k must be in 1.." + std::to_string(MAX_K));
    if (b->items_built == 1u && op != OP_AND && op != OP_RANKED_AND) return fail(DS2I_E_ARG, "batch was prepared for the conjunctive operators only");
    ds2i_gpu_index* ix = b->index;
    device_scope on_device(ix->device);
    CUDA_TRY(on_device.status);
    if (!ranked) k = 1;
    b->layout_outputs(k);
    DevBatch db{};
    db.nq = b->nq; db.q_begin = b->q_begin.p; db.term = b->term.p; db.q_weight = b->q_weight.p;
    db.max_weight = b->max_weight.p; db.ord_size = b->ord_size.p; db.ord_maxw = b->ord_maxw.p;
    db.sched = b->sched.p; db.work_counter = b->work_counter.p; db.out_counts = b->out_counts.p;
    db.out_scores = b->out_scores.p; db.out_docids = b->out_docids.p; db.stats = b->stats.p;
    CUDA_TRY(cudaMemsetAsync(b->work_counter.p, 0, b->counters_bytes));          // work counters and, behind them in the arena, the statistics
    CUDA_TRY(cudaEventRecord(b->ev0));
    int rc = DS2I_OK;
    if (b->nq) {
        const bool fast = !(flags & DS2I_RUN_FAITHFUL), no_stats = (flags & DS2I_RUN_NO_STATS) != 0;
        if (ix->kind == KIND_PEF && fast && (op == OP_AND || op == OP_RANKED_AND) && (b->items_built & 1u)) {
            if (op == OP_AND) rc = launch_and_block<CODEC_PEF, false>(b, db, k);
            else rc = no_stats ? launch_and_block<CODEC_PEF, true, false>(b, db, k) : launch_and_block<CODEC_PEF, true>(b, db, k);
        }
        else if (ix->kind == KIND_PEF && fast && (op == OP_WAND || op == OP_MAXSCORE) && (b->items_built & 2u))
            rc = no_stats ? launch_union_block<CODEC_PEF, false>(b, db, k) : launch_union_block<CODEC_PEF>(b, db, k);
        else if (fast && (op == OP_OR || op == OP_RANKED_OR) && (b->items_built & 2u)) rc = launch_union_exhaustive(b, db, op, k);
        else if (ix->kind == KIND_PEF) rc = pef_launch_query(*ix->pef, b->wand ? b->wand->dev : DevWand{nullptr, nullptr, 0.f}, db, op, k, b->max_terms, ix->sm_count, g_last_error);
        else if (!(flags & DS2I_RUN_FAITHFUL) && (op == OP_AND || op == OP_RANKED_AND) && (b->items_built & 1u)) {
            rc = op == OP_AND ? launch_and_block_codec<false>(b, db, k, no_stats) : launch_and_block_codec<true>(b, db, k, no_stats);
        }
        else if (!(flags & DS2I_RUN_FAITHFUL) && (op == OP_WAND || op == OP_MAXSCORE) && (b->items_built & 2u)) {
            switch (ix->codec) {       // per-codec instances: smaller kernels, fewer instruction-cache misses
                case CODEC_OPTPFOR: rc = no_stats ? launch_union_block<CODEC_OPTPFOR, false>(b, db, k) : launch_union_block<CODEC_OPTPFOR>(b, db, k); break;
#ifndef DS2I_DEV_FAST_BUILD
                case CODEC_VARINT: rc = launch_union_block<CODEC_VARINT>(b, db, k); break;
                case CODEC_INTERPOLATIVE: rc = launch_union_block<CODEC_INTERPOLATIVE>(b, db, k); break;
                case CODEC_QMX: rc = launch_union_block<CODEC_QMX>(b, db, k); break;
                case CODEC_MIXED: rc = launch_union_block<CODEC_MIXED>(b, db, k); break;
#endif
                default: rc = fail(DS2I_E_UNSUPPORTED, "unknown codec");
            }
        }
        else rc = launch_query_op<CODEC_ANY>(b, db, op, k);
        if (rc != DS2I_OK) return rc;
        b->launches += 1;
    }
    CUDA_TRY(cudaEventRecord(b->ev1));
    b->last_k = ranked ? k : 0; b->last_ranked = ranked;
    b->pending = true;
    if (flags & DS2I_RUN_ASYNC) {                      // the caller orders later work behind stream 0 or calls ds2i_gpu_batch_wait
        CUDA_TRY(cudaGetLastError());
        if (out_elapsed_ms) *out_elapsed_ms = 0.f;
        return DS2I_OK;
    }
    return ds2i_gpu_batch_wait(b, out_elapsed_ms);
}

extern "C" int ds2i_gpu_batch_wait(ds2i_gpu_batch* b, float* out_elapsed_ms) {
    if (!b) return fail(DS2I_E_ARG, "null batch");
    if (!b->pending) { if (out_elapsed_ms) *out_elapsed_ms = 0.f; return DS2I_OK; }
    device_scope on_device(b->index->device);
    CUDA_TRY(on_device.status);
    CUDA_TRY(cudaEventSynchronize(b->ev1));
    CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
    if (out_elapsed_ms) *out_elapsed_ms = ms;
    b->pending = false;
#ifdef DS2I_NIN_HIST
    {
        unsigned long long h[32];
        cudaMemcpyFromSymbol(h, g_hist, sizeof(h));
        fprintf(stderr, "[hist]");
        for (int i = 0; i < 32; ++i) fprintf(stderr, " %llu", h[i]);
        fprintf(stderr, "\n");
        memset(h, 0, sizeof(h));
        cudaMemcpyToSymbol(g_hist, h, sizeof(h));
    }
#endif
    return DS2I_OK;
}

extern "C" int ds2i_gpu_batch_fetch(ds2i_gpu_batch* b, uint64_t* out_counts, float* out_scores) {
    if (!b) return fail(DS2I_E_ARG, "null batch");
    device_scope on_device(b->index->device);
    CUDA_TRY(on_device.status);
    if (out_counts && b->nq) CUDA_TRY(cudaMemcpy(out_counts, b->out_counts.p, size_t(b->nq) * 8, cudaMemcpyDeviceToHost));
    if (out_scores && b->nq && b->last_ranked)
        CUDA_TRY(cudaMemcpy(out_scores, b->out_scores.p, size_t(b->nq) * b->last_k * 4, cudaMemcpyDeviceToHost));
    return DS2I_OK;
}

extern "C" int ds2i_gpu_batch_fetch_docids(ds2i_gpu_batch* b, uint32_t* out_docids) {
    if (!b || !out_docids) return fail(DS2I_E_ARG, "null argument");
    if (!b->last_ranked) return fail(DS2I_E_ARG, "the last operator run on this batch was not a ranked one");
    device_scope on_device(b->index->device);
    CUDA_TRY(on_device.status);
    if (b->nq) CUDA_TRY(cudaMemcpy(out_docids, b->out_docids.p, size_t(b->nq) * b->last_k * 4, cudaMemcpyDeviceToHost));
    return DS2I_OK;
}

extern "C" int ds2i_gpu_batch_stats(ds2i_gpu_batch* b, uint64_t out_stats[8]) {
    if (!b || !out_stats) return fail(DS2I_E_ARG, "null argument");
    device_scope on_device(b->index->device);
    CUDA_TRY(on_device.status);
    unsigned long long s[8];
    CUDA_TRY(cudaMemcpy(s, b->stats.p, sizeof(s), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 8; ++i) out_stats[i] = s[i];
    out_stats[6] = b->launches;
    return DS2I_OK;
}

extern "C" int ds2i_gpu_batch_device_results(ds2i_gpu_batch* b, void** d_counts, void** d_scores) {
    if (!b) return fail(DS2I_E_ARG, "null batch");
    if (d_counts) *d_counts = b->out_counts.p;
    if (d_scores) *d_scores = b->out_scores.p;
    return DS2I_OK;
}

extern "C" int ds2i_gpu_batch_device_fused(ds2i_gpu_batch* b, void** d_fused, size_t* bytes) {
    if (!b || !d_fused || !bytes) return fail(DS2I_E_ARG, "null argument");
    *d_fused = b->out_fused.p;
    *bytes = b->fused_bytes(b->last_ranked ? b->last_k : 0);
    return DS2I_OK;
}

extern "C" int ds2i_gpu_batch_device_docids(ds2i_gpu_batch* b, void** d_docids) {
    if (!b || !d_docids) return fail(DS2I_E_ARG, "null argument");
    *d_docids = b->out_docids.p;
    return DS2I_OK;
}

extern "C" void ds2i_gpu_batch_free(ds2i_gpu_batch* b) { if (b) { device_scope ds(b->index->device); delete b; } }

// ------------------------------------------------------------------------------------------------
// Document-partitioned shards: fold the per-shard results of one query batch.  Shards hold disjoint
// documents, so match counts add up and the global top-k is the top-k of the shards' top-k lists.
__global__ void __launch_bounds__(128) merge_shards_kernel(const uint64_t* counts, const float* scores, const uint32_t* docids, uint32_t nshards,
                                                           uint32_t nq, uint32_t k, bool ranked, uint64_t* out_counts, float* out_scores,
                                                           uint32_t* out_docids) {
    const unsigned lane = lane_id();
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    uint64_t total = 0;
    TopK topk;
    topk.init(ranked ? k : 1u);
    for (uint32_t s = 0; s < nshards; ++s) {
        const size_t row = size_t(s) * nq + q;
        const uint64_t c = counts[row];
        total += c;
        if (!ranked) continue;
        const uint32_t n = uint32_t(c < k ? c : k);
        const float v = lane < n ? scores[row * k + lane] : 0.f;
        const uint32_t id = lane < n ? docids[row * k + lane] : 0xffffffffu;
        for (uint32_t j = 0; j < n; ++j) {
            const float sc = __shfl_sync(FULL, v, j);
            if (!topk.would_enter(sc)) break;          // per-shard lists are sorted descending
            topk.insert(sc, __shfl_sync(FULL, id, j));
        }
    }
    if (lane == 0) out_counts[q] = ranked ? uint64_t(topk.size) : total;
    if (ranked && lane < k) {
        out_scores[size_t(q) * k + lane] = lane < topk.size ? topk.v : 0.f;
        out_docids[size_t(q) * k + lane] = lane < topk.size ? topk.id : 0xffffffffu;
    }
}

extern "C" int ds2i_gpu_merge_shards(const uint64_t* d_counts, const float* d_scores, const uint32_t* d_docids, uint32_t nshards, size_t nq,
                                     uint32_t k, int ranked, uint64_t* d_out_counts, float* d_out_scores, uint32_t* d_out_docids) {
    if (!d_counts || !d_out_counts || (ranked && (!d_scores || !d_docids || !d_out_scores || !d_out_docids))) return fail(DS2I_E_ARG, "null argument");
    if (ranked && (k < 1 || k > MAX_K)) return fail(DS2I_E_LIMIT, "k must be in 1.." + std::to_string(MAX_K));
    if (nq > 0x7fffffffull) return fail(DS2I_E_LIMIT, "too many queries in one batch");
    if (nq && nshards)
        merge_shards_kernel<<<unsigned((nq + 3) / 4), 128>>>(d_counts, d_scores, d_docids, nshards, uint32_t(nq), k, ranked != 0, d_out_counts,
                                                             d_out_scores, d_out_docids);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(0));
    return DS2I_OK;
}

extern "C" int ds2i_gpu_query_batch(ds2i_gpu_index* ix, ds2i_gpu_wand* wand, int op, uint32_t k,
                                    const uint32_t* terms, const uint64_t* query_offsets, size_t nq,
                                    uint64_t* out_counts, float* out_scores, float* out_elapsed_ms) {
    return ds2i_gpu_query_batch_docids(ix, wand, op, k, terms, query_offsets, nq, out_counts, out_scores, nullptr, out_elapsed_ms);
}

extern "C" int ds2i_gpu_query_batch_docids(ds2i_gpu_index* ix, ds2i_gpu_wand* wand, int op, uint32_t k,
                                           const uint32_t* terms, const uint64_t* query_offsets, size_t nq,
                                           uint64_t* out_counts, float* out_scores, uint32_t* out_docids, float* out_elapsed_ms) {
    const bool trace = trace_on();
    double t0 = now_ms();
    ds2i_gpu_batch* b = nullptr;
    unsigned which = (op == OP_AND || op == OP_RANKED_AND) ? 1u : (op == OP_WAND || op == OP_MAXSCORE || op == OP_OR || op == OP_RANKED_OR) ? 2u : 0u;
    int rc = batch_prepare_impl(ix, wand, terms, query_offsets, nq, which, &b);
    if (rc != DS2I_OK) return rc;
    std::unique_ptr<ds2i_gpu_batch> guard(b);
    double t1 = now_ms();
    rc = ds2i_gpu_batch_run_ex(b, op, k, DS2I_RUN_NO_STATS, out_elapsed_ms);     // the batch dies with the call: nobody could read its counters
    if (rc != DS2I_OK) return rc;
    double t2 = now_ms();
    rc = ds2i_gpu_batch_fetch(b, out_counts, out_scores);
    if (rc == DS2I_OK && out_docids && b->last_ranked) rc = ds2i_gpu_batch_fetch_docids(b, out_docids);
    double t3 = now_ms();
    guard.reset();
    if (trace) fprintf(stderr, "[ds2i_gpu] query_batch nq=%zu prepare %.2f ms, run %.2f ms, fetch %.2f ms, free %.2f ms\n", nq, t1 - t0, t2 - t1, t3 - t2, now_ms() - t3);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// Batched block decode (BASELINE config 2): one warp per 128-posting block, any list, any order.
constexpr size_t SINGLE_LIST_WARP_BYTES = sizeof(ListState) + STAGE_WORDS * 4 + SCRATCH_WORDS * 4 + 16;

// one THREAD per interpolative-coded block, reading the bit stream straight from global memory.  TAILS_ONLY: the job's
// index codes full blocks with another codec, so only the last, partial block of a list is bit-serial (thread i <-> list i);
// otherwise (block_interpolative) thread g <-> the g-th block of the job.
constexpr int SERIAL_WARPS = 2;
__host__ __device__ constexpr size_t serial_warp_bytes(uint32_t rows) { return size_t(rows) * SERIAL_STRIDE * 4; }

// ROWS: values per lane column of the transposition buffer (tails of up to 64 values run with half the shared memory, twice the warps)
template <bool TAILS_ONLY, uint32_t ROWS>
__global__ void __launch_bounds__(SERIAL_WARPS * 32) decode_serial_blocks_kernel(DevIndex idx, DecodeJob job) {
    const unsigned lane = lane_id();
    uint32_t* buf = reinterpret_cast<uint32_t*>(g_smem) + (threadIdx.x >> 5) * (serial_warp_bytes(ROWS) / 4);
    uint32_t* col = buf + lane;
    const uint64_t g = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    uint32_t lo = 0, b = 0, size = 0;
    ListDir d{};
    bool valid;
    if (TAILS_ONLY) {
        valid = g < job.tail_count;
        if (valid) {
            // neighbouring lanes take tails of similar size (the loop below runs as long as the warp's longest block)
            lo = job.tail_order ? job.tail_order[job.tail_first + g] : uint32_t(job.tail_first + g);
            d = idx.dir[job.terms[lo]];
            valid = d.n % BLOCK != 0;
            b = (d.n + BLOCK - 1) / BLOCK - 1;
        }
    } else {
        valid = g < job.total_blocks;
        if (valid) {
            uint32_t hi = job.nterms;
            while (hi - lo > 1) {
                uint32_t mid = (lo + hi) >> 1;
                if (job.blk_prefix[mid] <= g) lo = mid; else hi = mid;
            }
            b = uint32_t(g - job.blk_prefix[lo]);
            d = idx.dir[job.terms[lo]];
        }
    }
    const uint8_t* in = nullptr;
    uint32_t cur_base = 0;
    uint64_t o = 0;
    if (valid) {
        const uint32_t nblocks = (d.n + BLOCK - 1) / BLOCK;
        size = ((uint64_t(b) + 1) * BLOCK > d.n) ? d.n % BLOCK : BLOCK;
        const uint2* bd = idx.bdir + idx.bfirst[job.terms[lo]];
        const uint2 prev = b ? __ldg(bd + b - 1) : make_uint2(0xffffffffu, 0u);
        cur_base = prev.x + 1u;
        const uint32_t cur_max = __ldg(bd + b).x;
        in = idx.lists + d.maxs_off + 4ull * nblocks + 4ull * (nblocks - 1) + prev.y;
        o = job.out_offsets[lo] + uint64_t(b) * BLOCK;
        in += decode_interpolative_lane(in, size, cur_max - cur_base - (size - 1u), col);
    }
    __syncwarp();
    // the warp writes the 32 blocks out one after the other: docid_i = base + P[i] + i (block_posting_list.hpp:313-317)
    for (uint32_t l = 0; l < 32; ++l) {
        const uint32_t n_l = __shfl_sync(FULL, size, l);
        if (!n_l) continue;
        const uint32_t base_l = __shfl_sync(FULL, cur_base, l);
        const uint64_t o_l = (uint64_t(__shfl_sync(FULL, uint32_t(o >> 32), l)) << 32) | __shfl_sync(FULL, uint32_t(o), l);
        for (uint32_t i = lane; i < n_l; i += 32) job.out_docs[o_l + i] = base_l + buf[i * SERIAL_STRIDE + l] + i;
    }
    __syncwarp();
    if (valid) decode_interpolative_lane(in, size, 0xffffffffu, col);
    __syncwarp();
    for (uint32_t l = 0; l < 32; ++l) {
        const uint32_t n_l = __shfl_sync(FULL, size, l);
        if (!n_l) continue;
        const uint64_t o_l = (uint64_t(__shfl_sync(FULL, uint32_t(o >> 32), l)) << 32) | __shfl_sync(FULL, uint32_t(o), l);
        for (uint32_t i = lane; i < n_l; i += 32)
            job.out_freqs[o_l + i] = buf[i * SERIAL_STRIDE + l] - (i ? buf[(i - 1) * SERIAL_STRIDE + l] : 0u) + 1u;
    }
}

// sum of n u32 values as u64 (parity of a full-scale decode without moving 3.4 GB to the host: "a checksum of checksums")
__global__ void __launch_bounds__(256) sum_u32_kernel(const uint32_t* v, uint64_t n, unsigned long long* out) {
    unsigned long long acc = 0;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) acc += v[i];
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(FULL, acc, d);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

static int decode_lists_impl(ds2i_gpu_index* ix, const uint32_t* terms, size_t nterms, const uint64_t* out_offsets, uint32_t* out_docs,
                             uint32_t* out_freqs, uint64_t* out_sums /* [2] or null */, float* out_elapsed_ms);

extern "C" int ds2i_gpu_decode_lists(ds2i_gpu_index* ix, const uint32_t* terms, size_t nterms,
                                     const uint64_t* out_offsets, uint32_t* out_docs, uint32_t* out_freqs,
                                     float* out_elapsed_ms) {
    return decode_lists_impl(ix, terms, nterms, out_offsets, out_docs, out_freqs, nullptr, out_elapsed_ms);
}

extern "C" int ds2i_gpu_decode_lists_checksum(ds2i_gpu_index* ix, const uint32_t* terms, size_t nterms, const uint64_t* out_offsets,
                                              uint64_t* out_sum_docids, uint64_t* out_sum_freqs, float* out_elapsed_ms) {
    if (!out_sum_docids || !out_sum_freqs) return fail(DS2I_E_ARG, "null argument");
    uint64_t sums[2] = {0, 0};
    int rc = decode_lists_impl(ix, terms, nterms, out_offsets, nullptr, nullptr, sums, out_elapsed_ms);
    *out_sum_docids = sums[0]; *out_sum_freqs = sums[1];
    return rc;
}

static int decode_lists_impl(ds2i_gpu_index* ix, const uint32_t* terms, size_t nterms, const uint64_t* out_offsets, uint32_t* out_docs,
                             uint32_t* out_freqs, uint64_t* out_sums, float* out_elapsed_ms) {
    if (!ix || (!terms && nterms) || !out_offsets) return fail(DS2I_E_ARG, "null argument");
    if (nterms > 0x7fffffffull) return fail(DS2I_E_LIMIT, "too many lists");
    device_scope on_device(ix->device);
    CUDA_TRY(on_device.status);
    std::vector<uint64_t> blk(nterms + 1, 0), offs(out_offsets, out_offsets + nterms + 1);
    for (size_t i = 0; i < nterms; ++i) {
        if (terms[i] >= ix->size) return fail(DS2I_E_ARG, "term id out of range");
        uint64_t n = list_size_of(ix, terms[i]);
        if (offs[i + 1] - offs[i] != n) return fail(DS2I_E_ARG, "out_offsets do not match the list sizes");
        blk[i + 1] = blk[i] + (n + BLOCK - 1) / BLOCK;
    }
    const uint64_t total = offs[nterms];
    dev_buf<uint32_t> d_terms, d_docs, d_freqs;
    dev_buf<uint64_t> d_blk, d_offs;
    std::vector<uint32_t> tv(terms, terms + nterms);
    CUDA_TRY(d_terms.upload(tv)); CUDA_TRY(d_blk.upload(blk)); CUDA_TRY(d_offs.upload(offs));
    CUDA_TRY(d_docs.alloc(total)); CUDA_TRY(d_freqs.alloc(total));
    PefDecodeItem* pef_items = nullptr;
    uint32_t pef_nitems = 0;
    cuda_free_guard pef_items_guard;
    if (ix->kind == KIND_PEF && nterms && total) {
        int prc = pef_decode_prepare(*ix->pef, terms, uint32_t(nterms), &pef_items, &pef_nitems, g_last_error);
        pef_items_guard.p = pef_items;
        if (prc) return prc;
    }
    // the bit-serial tails of a block index, ordered by size (stable counting sort): a warp then decodes blocks of one length
    dev_buf<uint32_t> d_order;
    uint32_t tail_bounds[2] = {0, 0};        // first list with a tail; first list with a tail of more than 64 values
    if (ix->kind != KIND_PEF && ix->codec != CODEC_INTERPOLATIVE && nterms) {
        std::vector<uint32_t> order(nterms), start(BLOCK + 1, 0);
        for (size_t i = 0; i < nterms; ++i) start[(list_size_of(ix, terms[i]) % BLOCK) + 1] += 1;
        for (uint32_t t = 0; t < BLOCK; ++t) start[t + 1] += start[t];
        tail_bounds[0] = start[1]; tail_bounds[1] = start[65];
        for (size_t i = 0; i < nterms; ++i) order[start[list_size_of(ix, terms[i]) % BLOCK]++] = uint32_t(i);
        CUDA_TRY(d_order.upload(order));
        CUDA_TRY(cudaStreamSynchronize(0));      // `order` dies with this scope
    }
    cuda_event ev0, ev1;
    CUDA_TRY(ev0.create()); CUDA_TRY(ev1.create());
    cudaEvent_t e0 = ev0.e, e1 = ev1.e;
    CUDA_TRY(cudaEventRecord(e0));
    int rc = DS2I_OK;
    if (nterms && total) {
        if (ix->kind == KIND_PEF) {
            pef_decode_launch(*ix->pef, pef_items, pef_nitems, d_terms.p, d_offs.p, d_docs.p, d_freqs.p, ix->sm_count);
        } else {
            DecodeJob job{d_terms.p, d_blk.p, d_offs.p, d_docs.p, d_freqs.p, blk[nterms], uint32_t(nterms), d_order.p, 0u, uint32_t(nterms)};
            uint64_t want = (blk[nterms] / 32 + 4) / 4;
            int grid = int(std::min<uint64_t>(want, uint64_t(ix->sm_count) * 8));
            const size_t dsmem = S16_TAB_BYTES + 4 * DECODE_WARP_BYTES;
            switch (ix->codec) {
                case CODEC_OPTPFOR: decode_full_blocks_kernel<CODEC_OPTPFOR><<<grid, 128, dsmem>>>(ix->dev, job); break;
                case CODEC_VARINT: decode_full_blocks_kernel<CODEC_VARINT><<<grid, 128, dsmem>>>(ix->dev, job); break;
                case CODEC_QMX: decode_full_blocks_kernel<CODEC_QMX><<<grid, 128, dsmem>>>(ix->dev, job); break;
                case CODEC_MIXED: decode_full_blocks_kernel<CODEC_MIXED><<<grid, 128, dsmem>>>(ix->dev, job); break;
                default: break;      // block_interpolative: every block is bit-serial
            }
            const int st = SERIAL_WARPS * 32;
            if (ix->codec == CODEC_INTERPOLATIVE) {
                decode_serial_blocks_kernel<false, BLOCK><<<unsigned((blk[nterms] + st - 1) / st), st, SERIAL_WARPS * serial_warp_bytes(BLOCK)>>>(ix->dev, job);
            } else {
                // tails in order of size: [no tail][1 .. 64 values][65 .. 127 values]
                auto slice = [&](uint32_t first, uint32_t count, bool small) {
                    if (!count) return;
                    DecodeJob j2 = job;
                    j2.tail_first = first; j2.tail_count = count;
                    const unsigned grid = unsigned((count + st - 1) / st);
                    if (small) decode_serial_blocks_kernel<true, 64><<<grid, st, SERIAL_WARPS * serial_warp_bytes(64)>>>(ix->dev, j2);
                    else decode_serial_blocks_kernel<true, BLOCK><<<grid, st, SERIAL_WARPS * serial_warp_bytes(BLOCK)>>>(ix->dev, j2);
                };
                slice(tail_bounds[0], tail_bounds[1] - tail_bounds[0], true);
                slice(tail_bounds[1], uint32_t(nterms) - tail_bounds[1], false);
            }
        }
    }
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (rc != DS2I_OK) return rc;
    if (out_elapsed_ms) *out_elapsed_ms = ms;
    if (total && out_docs) CUDA_TRY(cudaMemcpy(out_docs, d_docs.p, total * 4, cudaMemcpyDeviceToHost));
    if (total && out_freqs) CUDA_TRY(cudaMemcpy(out_freqs, d_freqs.p, total * 4, cudaMemcpyDeviceToHost));
    if (out_sums) {
        dev_buf<unsigned long long> d_sums;
        CUDA_TRY(d_sums.alloc(2));
        CUDA_TRY(cudaMemsetAsync(d_sums.p, 0, 16));
        if (total) {
            sum_u32_kernel<<<ix->sm_count * 8, 256>>>(d_docs.p, total, d_sums.p);
            sum_u32_kernel<<<ix->sm_count * 8, 256>>>(d_freqs.p, total, d_sums.p + 1);
        }
        unsigned long long h[2];
        CUDA_TRY(cudaMemcpy(h, d_sums.p, 16, cudaMemcpyDeviceToHost));
        out_sums[0] = h[0]; out_sums[1] = h[1];
    }
    return DS2I_OK;
}

// ------------------------------------------------------------------------------------------------
// next_geq sweeps (BASELINE config 3 harness, also the enumerator parity test): warp per list.
struct GeqJob {
    const uint32_t* terms;
    const uint64_t* bounds;
    const uint64_t* bound_offsets;
    uint64_t* out_docids;
    uint64_t* out_freqs;
    uint32_t* work_counter;
    uint32_t nlists;
};

template <int CODEC>
__global__ void __launch_bounds__(128) next_geq_kernel(DevIndex idx, GeqJob job) {
    s16_table_init(smem_words(0));
    __syncthreads();
    typedef BlockEnum<CODEC> E;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    uint8_t* base = g_smem + S16_TAB_BYTES + warp * SINGLE_LIST_WARP_BYTES;
    ListState* st = reinterpret_cast<ListState*>(base);
    uint32_t* stage = reinterpret_cast<uint32_t*>(base + sizeof(ListState));
    uint32_t* scratch = stage + STAGE_WORDS;
    uint64_t* bar = reinterpret_cast<uint64_t*>(scratch + SCRATCH_WORDS);
    WarpCtx c;
    ctx_init(c, stage, scratch, bar, idx.codec);
    while (true) {
        uint32_t li = 0;
        if (lane == 0) li = atomicAdd(job.work_counter, 1u);
        li = __shfl_sync(FULL, li, 0);
        if (li >= job.nlists) break;
        E::open(c, idx, st, job.terms[li]);
        for (uint64_t j = job.bound_offsets[li]; j < job.bound_offsets[li + 1]; ++j) {
            uint64_t lb = job.bounds[j];
            uint32_t d = lb >= idx.num_docs ? E::next_geq(c, idx, st, idx.num_docs) : E::next_geq(c, idx, st, uint32_t(lb));
            uint32_t f = d < idx.num_docs ? E::freq(c, idx, st) : 0u;
            if (lane == 0) { job.out_docids[j] = d; job.out_freqs[j] = f; }
        }
    }
}

extern "C" int ds2i_gpu_next_geq_batch(ds2i_gpu_index* ix, const uint32_t* terms, size_t nlists,
                                       const uint64_t* bounds, const uint64_t* bound_offsets,
                                       uint64_t* out_docids, uint64_t* out_freqs, float* out_elapsed_ms) {
    if (!ix || (!terms && nlists) || !bound_offsets) return fail(DS2I_E_ARG, "null argument");
    if (nlists > 0x7fffffffull) return fail(DS2I_E_LIMIT, "too many lists");
    device_scope on_device(ix->device);
    CUDA_TRY(on_device.status);
    for (size_t i = 0; i < nlists; ++i)
        if (terms[i] >= ix->size) return fail(DS2I_E_ARG, "term id out of range");
    const uint64_t total = nlists ? bound_offsets[nlists] : 0;
    std::vector<uint32_t> tv(terms, terms + nlists);
    std::vector<uint64_t> bv(bounds, bounds + total), ov(bound_offsets, bound_offsets + nlists + 1);
    dev_buf<uint32_t> d_terms, d_counter;
    dev_buf<uint64_t> d_bounds, d_offs, d_docids, d_freqs;
    CUDA_TRY(d_terms.upload(tv)); CUDA_TRY(d_bounds.upload(bv)); CUDA_TRY(d_offs.upload(ov));
    CUDA_TRY(d_docids.alloc(total)); CUDA_TRY(d_freqs.alloc(total)); CUDA_TRY(d_counter.alloc(1));
    CUDA_TRY(cudaMemset(d_counter.p, 0, 4));
    cuda_event ev0, ev1;
    CUDA_TRY(ev0.create()); CUDA_TRY(ev1.create());
    cudaEvent_t e0 = ev0.e, e1 = ev1.e;
    CUDA_TRY(cudaEventRecord(e0));
    int rc = DS2I_OK;
    if (nlists) {
        if (ix->kind == KIND_PEF) {
            rc = pef_next_geq(*ix->pef, d_terms.p, uint32_t(nlists), d_bounds.p, d_offs.p, d_docids.p, d_freqs.p, d_counter.p, ix->sm_count, g_last_error);
        } else {
            GeqJob job{d_terms.p, d_bounds.p, d_offs.p, d_docids.p, d_freqs.p, d_counter.p, uint32_t(nlists)};
            int grid = int(std::min<uint64_t>((nlists + 3) / 4, uint64_t(ix->sm_count) * 8));
            next_geq_kernel<CODEC_ANY><<<grid, 128, S16_TAB_BYTES + 4 * SINGLE_LIST_WARP_BYTES>>>(ix->dev, job);
        }
    }
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (rc != DS2I_OK) return rc;
    if (out_elapsed_ms) *out_elapsed_ms = ms;
    if (total && out_docids) CUDA_TRY(cudaMemcpy(out_docids, d_docids.p, total * 8, cudaMemcpyDeviceToHost));
    if (total && out_freqs) CUDA_TRY(cudaMemcpy(out_freqs, d_freqs.p, total * 8, cudaMemcpyDeviceToHost));
    return DS2I_OK;
}

// ------------------------------------------------------------------------------------------------
// Several GPUs behind the C ABI: one process, the index (and wand data) replicated on every device, the
// query batch cut into cost-balanced shards, one host thread per GPU (the reference's own parallel driver deals query i to
// thread i % n the same way, profile_queries.cpp:21-39), per-shard results gathered to the first device over NCCL
// (ncclCommInitAll; NVLink / NVSwitch) and copied to the caller's host buffers in one D2H transfer.
struct nccl_api {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        // a process that already holds an NCCL (e.g. PyTorch's bundled one) gets that copy: same soname
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) { lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) { err = std::string("NCCL not found: ") + dlerror(); return false; }
        auto sym = [&](const char* n) { void* p = dlsym(lib, n); if (!p) err = std::string("NCCL symbol missing: ") + n; return p; };
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(sym("ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
        Send = reinterpret_cast<decltype(Send)>(sym("ncclSend"));
        Recv = reinterpret_cast<decltype(Recv)>(sym("ncclRecv"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
        return CommInitAll && CommDestroy && GroupStart && GroupEnd && Send && Recv && GetErrorString;
    }
};
static nccl_api g_nccl;
static std::mutex g_nccl_mu;

struct ds2i_gpu_group {
    std::vector<int> devices;
    std::vector<ds2i_gpu_index*> indexes;
    std::vector<ds2i_gpu_wand*> wands;          // empty when the group was opened without wand data
    std::vector<ncclComm_t> comms;              // empty for a single device
    dev_buf<uint8_t> gathered;                  // on devices[0]: the fused results of every shard, back to back
    pinned_arena host;                          // D2H staging of the gathered results
    std::mutex mu;                              // one batch at a time per group (gathered / host are per group)
    ~ds2i_gpu_group() {
        for (ncclComm_t c : comms) if (c) g_nccl.CommDestroy(c);
        for (auto* w : wands) ds2i_gpu_wand_close(w);
        for (auto* ix : indexes) ds2i_gpu_index_close(ix);
    }
};

extern "C" int ds2i_gpu_group_open(const char* index_path, const char* index_type, const char* wand_path, const int* devices, int ndevices,
                                   ds2i_gpu_group** out) {
    if (!index_path || !index_type || !out || ndevices < 1) return fail(DS2I_E_ARG, "bad argument");
    int have = 0;
    CUDA_TRY(cudaGetDeviceCount(&have));
    std::unique_ptr<ds2i_gpu_group> g(new ds2i_gpu_group);
    for (int i = 0; i < ndevices; ++i) {
        const int dev = devices ? devices[i] : i;
        if (dev < 0 || dev >= have) return fail(DS2I_E_ARG, "device " + std::to_string(dev) + " does not exist (" + std::to_string(have) + " visible)");
        if (std::find(g->devices.begin(), g->devices.end(), dev) != g->devices.end()) return fail(DS2I_E_ARG, "device listed twice");
        g->devices.push_back(dev);
    }
    // replicas are loaded concurrently, one thread per device (the file is read through the page cache once)
    g->indexes.assign(ndevices, nullptr);
    if (wand_path) g->wands.assign(ndevices, nullptr);
    std::vector<int> rcs(ndevices, DS2I_OK);
    std::vector<std::string> errs(ndevices);
    {
        std::vector<std::thread> pool;
        for (int i = 0; i < ndevices; ++i)
            pool.emplace_back([&, i] {
                rcs[i] = ds2i_gpu_index_open_file(index_path, index_type, g->devices[i], &g->indexes[i]);
                if (rcs[i] == DS2I_OK && wand_path) rcs[i] = ds2i_gpu_wand_open_file(wand_path, g->devices[i], &g->wands[i]);
                if (rcs[i] != DS2I_OK) errs[i] = ds2i_gpu_last_error();
            });
        for (auto& t : pool) t.join();
    }
    for (int i = 0; i < ndevices; ++i) if (rcs[i] != DS2I_OK) return fail(rcs[i], errs[i]);
    if (ndevices > 1) {
        std::lock_guard<std::mutex> lock(g_nccl_mu);
        std::string err;
        if (!g_nccl.load(err)) return fail(DS2I_E_UNSUPPORTED, err);
        g->comms.assign(ndevices, nullptr);
        ncclResult_t r = g_nccl.CommInitAll(g->comms.data(), ndevices, g->devices.data());
        if (r != ncclSuccess) { g->comms.clear(); return fail(DS2I_E_CUDA, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r)); }
    }
    *out = g.release();
    return DS2I_OK;
}

extern "C" void ds2i_gpu_group_close(ds2i_gpu_group* g) { delete g; }
extern "C" int ds2i_gpu_group_size(const ds2i_gpu_group* g) { return g ? int(g->devices.size()) : 0; }
extern "C" ds2i_gpu_index* ds2i_gpu_group_index(ds2i_gpu_group* g, int i) { return g && i >= 0 && size_t(i) < g->indexes.size() ? g->indexes[i] : nullptr; }

extern "C" int ds2i_gpu_group_query_batch(ds2i_gpu_group* g, int op, uint32_t k, const uint32_t* terms, const uint64_t* query_offsets, size_t nq,
                                          uint64_t* out_counts, float* out_scores, uint32_t* out_docids, float* out_elapsed_ms) {
    if (!g || !query_offsets || (!terms && nq && query_offsets[nq])) return fail(DS2I_E_ARG, "null argument");
    if (op < 0 || op > OP_RANKED_OR) return fail(DS2I_E_ARG, "unknown operator");
    const bool ranked = op >= OP_RANKED_AND;
    if (ranked && g->wands.empty()) return fail(DS2I_E_ARG, "ranked operators need wand data");
    if (ranked && (k < 1 || k > MAX_K)) return fail(DS2I_E_LIMIT, "k must be in 1.." + std::to_string(MAX_K));
    const size_t G = g->devices.size();
    const uint32_t kk = ranked ? k : 0;

    std::lock_guard<std::mutex> lock(g->mu);
    device_scope scope(g->devices[0]);          // the caller's current device is restored on every return path ...
    struct first_device { int d; ~first_device() { cudaSetDevice(d); } } back_to_first{g->devices[0]};     // ... also after the per-device waits below

    // cost-balanced shards: queries sorted by estimated cost, dealt round-robin (shard sizes differ by <= 1).  Cost of a
    // conjunctive query = block decodes of its evaluation: two per block of the shortest list (docs + freqs) plus, for every
    // other list, the blocks it can be probed in, min(its blocks, 128 candidates x blocks of the shortest list); of any other
    // query = the postings of its lists.  (ds2i_b200/parallel.py query_costs is the same model for torch.distributed hosts.)
    std::vector<uint32_t> order(nq);
    {
        ds2i_gpu_index* ix0 = g->indexes[0];
        const bool conjunctive = op == OP_AND || op == OP_RANKED_AND || op == OP_AND_FREQ;
        std::vector<uint64_t> cost(nq, 0), nb;
        for (size_t q = 0; q < nq; ++q) {
            if (query_offsets[q + 1] < query_offsets[q]) return fail(DS2I_E_ARG, "query_offsets not monotone");
            nb.clear();
            for (uint64_t j = query_offsets[q]; j < query_offsets[q + 1]; ++j) {
                if (terms[j] >= ix0->size) return fail(DS2I_E_ARG, "term id out of range in query " + std::to_string(q));
                const uint64_t n = list_size_of(ix0, terms[j]);
                if (!conjunctive) { cost[q] += n; continue; }
                if (std::find(terms + query_offsets[q], terms + j, terms[j]) == terms + j) nb.push_back((n + 127) / 128);    // distinct terms
            }
            if (conjunctive && !nb.empty()) {
                const uint64_t shortest = *std::min_element(nb.begin(), nb.end());
                uint64_t c = 0;
                for (uint64_t b : nb) c += std::min<uint64_t>(b, 128 * shortest);
                cost[q] = c - std::min<uint64_t>(shortest, 128 * shortest) + 2 * shortest;
            }
        }
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t c) { return cost[a] > cost[c]; });
    }
    std::vector<std::vector<uint32_t>> shard_q(G), shard_terms(G);
    std::vector<std::vector<uint64_t>> shard_offs(G);
    for (size_t s = 0; s < G; ++s) {
        shard_offs[s].push_back(0);
        for (size_t p = s; p < nq; p += G) {
            const uint32_t q = order[p];
            shard_q[s].push_back(q);
            shard_terms[s].insert(shard_terms[s].end(), terms + query_offsets[q], terms + query_offsets[q + 1]);
            shard_offs[s].push_back(shard_terms[s].size());
        }
    }

    // one host thread per GPU: prepare + run (results stay in the shard's fused device buffer)
    std::vector<ds2i_gpu_batch*> batches(G, nullptr);
    std::vector<int> rcs(G, DS2I_OK);
    std::vector<std::string> errs(G);
    std::vector<float> kernel_ms(G, 0.f);
    const unsigned which = (op == OP_AND || op == OP_RANKED_AND) ? 1u : (op == OP_WAND || op == OP_MAXSCORE || op == OP_OR || op == OP_RANKED_OR) ? 2u : 0u;
    auto work = [&](size_t s) {
        static const uint32_t no_terms = 0;
        const uint32_t* tp = shard_terms[s].empty() ? &no_terms : shard_terms[s].data();
        rcs[s] = batch_prepare_impl(g->indexes[s], g->wands.empty() ? nullptr : g->wands[s], tp, shard_offs[s].data(), shard_q[s].size(), which, &batches[s]);
        if (rcs[s] == DS2I_OK) rcs[s] = ds2i_gpu_batch_run_ex(batches[s], op, k, DS2I_RUN_NO_STATS, &kernel_ms[s]);     // the batch dies with the call
        if (rcs[s] != DS2I_OK) errs[s] = ds2i_gpu_last_error();
    };
    {
        std::vector<std::thread> pool;
        for (size_t s = 1; s < G; ++s) pool.emplace_back(work, s);
        work(0);
        for (auto& t : pool) t.join();
    }
    struct batch_guard { std::vector<ds2i_gpu_batch*>& v; ~batch_guard() { for (auto* b : v) ds2i_gpu_batch_free(b); } } guard{batches};
    for (size_t s = 0; s < G; ++s) if (rcs[s] != DS2I_OK) return fail(rcs[s], errs[s]);

    // gather the fused per-shard results on the first device: ncclSend from every other GPU, ncclRecv on the first, one group
    std::vector<size_t> off(G + 1, 0);
    for (size_t s = 0; s < G; ++s) off[s + 1] = off[s] + ((batches[s]->fused_bytes(kk) + 15) & ~size_t(15));
    const uint8_t* d_all = nullptr;
    CUDA_TRY(cudaSetDevice(g->devices[0]));
    if (G == 1) {
        d_all = batches[0]->out_fused.p;
    } else {
        CUDA_TRY(g->gathered.alloc(off[G]));
        CUDA_TRY(cudaMemcpyAsync(g->gathered.p, batches[0]->out_fused.p, batches[0]->fused_bytes(kk), cudaMemcpyDeviceToDevice, 0));
        ncclResult_t r = g_nccl.GroupStart();
        for (size_t s = 1; s < G && r == ncclSuccess; ++s) {
            const size_t bytes = batches[s]->fused_bytes(kk);
            if (!bytes) continue;
            r = g_nccl.Recv(g->gathered.p + off[s], bytes, ncclUint8, int(s), g->comms[0], 0);
            if (r == ncclSuccess) r = g_nccl.Send(batches[s]->out_fused.p, bytes, ncclUint8, 0, g->comms[s], 0);
        }
        ncclResult_t r2 = g_nccl.GroupEnd();
        if (r == ncclSuccess) r = r2;
        if (r != ncclSuccess) return fail(DS2I_E_CUDA, std::string("NCCL gather: ") + g_nccl.GetErrorString(r));
        CUDA_TRY(cudaSetDevice(g->devices[0]));
        d_all = g->gathered.p;
    }
    CUDA_TRY(g->host.reserve(std::max<size_t>(off[G], 16)));
    if (off[G]) CUDA_TRY(cudaMemcpyAsync(g->host.p, d_all, G == 1 ? batches[0]->fused_bytes(kk) : off[G], cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaStreamSynchronize(0));
    for (size_t s = 1; s < G; ++s) { CUDA_TRY(cudaSetDevice(g->devices[s])); CUDA_TRY(cudaStreamSynchronize(0)); }     // the sends have left the shard buffers

    // rows back into the caller's query order
    for (size_t s = 0; s < G; ++s) {
        const size_t n = shard_q[s].size();
        const uint8_t* base = g->host.p + off[s];
        const uint64_t* c = reinterpret_cast<const uint64_t*>(base);
        const float* sc = reinterpret_cast<const float*>(base + n * 8);
        const uint32_t* di = reinterpret_cast<const uint32_t*>(base + n * 8 + n * kk * 4);
        for (size_t i = 0; i < n; ++i) {
            const uint32_t q = shard_q[s][i];
            if (out_counts) out_counts[q] = c[i];
            if (ranked && out_scores) memcpy(out_scores + size_t(q) * k, sc + i * k, size_t(k) * 4);
            if (ranked && out_docids) memcpy(out_docids + size_t(q) * k, di + i * k, size_t(k) * 4);
        }
    }
    if (out_elapsed_ms) *out_elapsed_ms = *std::max_element(kernel_ms.begin(), kernel_ms.end());
    return DS2I_OK;
}
