// ds2i_gpu.hpp — header-only C++ adapters over the C ABI (ds2i_gpu.h) that model ds2i's own
// concepts, so code written against queries.hpp keeps its shape:
//
//   ds2i:  Index index; mapper::map(index, m);            here:  ds2i_gpu::gpu_index index(path, "block_optpfor");
//          wand_data<> wdata; mapper::map(wdata, md);             ds2i_gpu::gpu_wand_data wdata(path);
//          ranked_and_query op(wdata, k);                         ds2i_gpu::gpu_ranked_and_query op(wdata, k);
//          uint64_t n = op(index, terms); op.topk();              same call, same return value, same scores
//
// A single-query call is a batch of one; run_batch() is the product path (one launch per operator
// for the whole query log, what op_perftest's loop becomes — queries.cpp:25-35).
// Errors surface as exceptions like in the reference (std::invalid_argument / std::runtime_error).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ds2i_gpu.h"

namespace ds2i_gpu {

typedef uint32_t term_id_type;                      // queries.hpp:12-13
typedef std::vector<term_id_type> term_id_vec;

inline void check(int rc) {
    if (rc == DS2I_OK) return;
    std::string msg = ds2i_gpu_last_error();
    if (rc == DS2I_E_ARG) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}

// document_enumerator concept (block_posting_list.hpp:105-186, freq_index.hpp:116-190): the list is decoded ONCE on the
// device (ds2i_gpu_decode_lists) and the cursor then walks the decoded postings on the host.  This is for code that steps
// through single lists (verification, debugging, ad-hoc scans); the query operators never go through it — they evaluate
// whole batches on the device.  Semantics kept from the reference: positioned on the first posting after construction;
// past the end docid() == num_docs(); next_geq(lb) is "first posting >= lb at or after the cursor" (the block lists'
// definition) and keeps returning num_docs() after the end; freq() is valid on a posting only.
class gpu_document_enumerator {
public:
    gpu_document_enumerator() : m_pos(0), m_universe(0) {}
    gpu_document_enumerator(std::vector<uint32_t> docs, std::vector<uint32_t> freqs, uint64_t universe)
        : m_docs(std::move(docs)), m_freqs(std::move(freqs)), m_pos(0), m_universe(universe) {}
    void reset() { m_pos = 0; }
    void next() { if (m_pos < m_docs.size()) ++m_pos; }
    void next_geq(uint64_t lower_bound) {
        if (m_pos >= m_docs.size() || m_docs[m_pos] >= lower_bound) return;
        size_t lo = m_pos + 1, hi = m_docs.size();        // first position in (m_pos, size) with docs >= lower_bound
        while (lo < hi) { size_t mid = lo + (hi - lo) / 2; if (m_docs[mid] < lower_bound) lo = mid + 1; else hi = mid; }
        m_pos = lo;
    }
    void move(uint64_t position) { m_pos = position < m_docs.size() ? size_t(position) : m_docs.size(); }
    uint64_t docid() const { return m_pos < m_docs.size() ? m_docs[m_pos] : m_universe; }
    uint64_t freq() const { return m_freqs[m_pos]; }
    uint64_t position() const { return m_pos; }
    uint64_t size() const { return m_docs.size(); }
private:
    std::vector<uint32_t> m_docs, m_freqs;
    size_t m_pos;
    uint64_t m_universe;
};

class gpu_index {                                    // models the Index concept (block_freq_index.hpp:73-134)
public:
    gpu_index(const char* path, const char* index_type, int device = 0) : m_h(nullptr) {
        check(ds2i_gpu_index_open_file(path, index_type, device, &m_h));
    }
    gpu_index(const void* bytes, size_t n, const char* index_type, int device = 0) : m_h(nullptr) {
        check(ds2i_gpu_index_open(bytes, n, index_type, device, &m_h));
    }
    explicit gpu_index(ds2i_gpu_index* adopted) : m_h(adopted) {}      // takes ownership of a handle opened through the C ABI
    ~gpu_index() { ds2i_gpu_index_close(m_h); }
    gpu_index(gpu_index const&) = delete;
    gpu_index& operator=(gpu_index const&) = delete;
    size_t size() const { return size_t(ds2i_gpu_index_size(m_h)); }
    uint64_t num_docs() const { return ds2i_gpu_index_num_docs(m_h); }
    uint64_t list_size(term_id_type term) const { uint64_t n; check(ds2i_gpu_index_list_sizes(m_h, &term, 1, &n)); return n; }
    typedef gpu_document_enumerator document_enumerator;
    document_enumerator operator[](size_t term) const {       // Index::operator[] (block_freq_index.hpp:85-94, freq_index.hpp:192-214)
        const uint32_t t = uint32_t(term);
        uint64_t offsets[2] = {0, list_size(t)};
        const size_t n = size_t(offsets[1]);
        std::vector<uint32_t> docs(n), freqs(n);
        check(ds2i_gpu_decode_lists(m_h, &t, 1, offsets, docs.data(), freqs.data(), nullptr));
        return document_enumerator(std::move(docs), std::move(freqs), num_docs());
    }
    void warmup(size_t) const {}                     // the index is resident in HBM
    ds2i_gpu_index* handle() const { return m_h; }
private:
    ds2i_gpu_index* m_h;
};

class gpu_wand_data {                                // wand_data<bm25> (wand_data.hpp)
public:
    gpu_wand_data() : m_h(nullptr) {}
    explicit gpu_wand_data(const char* path, int device = 0) : m_h(nullptr) { open(path, device); }
    ~gpu_wand_data() { ds2i_gpu_wand_close(m_h); }
    gpu_wand_data(gpu_wand_data const&) = delete;
    gpu_wand_data& operator=(gpu_wand_data const&) = delete;
    void open(const char* path, int device = 0) { ds2i_gpu_wand_close(m_h); m_h = nullptr; check(ds2i_gpu_wand_open_file(path, device, &m_h)); }
    ds2i_gpu_wand* handle() const { return m_h; }
private:
    ds2i_gpu_wand* m_h;
};

struct query_batch_result {
    std::vector<uint64_t> counts;                    // the operator's return value per query
    std::vector<float> scores;                       // nq * k, descending, zero padded
    std::vector<uint32_t> docids;                    // nq * k: the document of every score (ranked operators)
    uint32_t k = 0;
    float elapsed_ms = 0;                            // CUDA-event time of the device work
    std::vector<float> topk(size_t q) const {
        size_t n = size_t(counts[q] < k ? counts[q] : k);
        return std::vector<float>(scores.begin() + q * k, scores.begin() + q * k + n);
    }
    std::vector<uint32_t> topk_docids(size_t q) const {
        size_t n = size_t(counts[q] < k ? counts[q] : k);
        return std::vector<uint32_t>(docids.begin() + q * k, docids.begin() + q * k + n);
    }
    double qps() const { return elapsed_ms > 0 ? counts.size() / (elapsed_ms * 1e-3) : 0; }
};

inline query_batch_result run_batch(gpu_index const& index, gpu_wand_data const* wdata, int op,
                                    std::vector<term_id_vec> const& queries, uint32_t k = 10) {
    std::vector<uint32_t> terms;
    std::vector<uint64_t> offsets(queries.size() + 1, 0);
    for (size_t i = 0; i < queries.size(); ++i) {
        terms.insert(terms.end(), queries[i].begin(), queries[i].end());
        offsets[i + 1] = terms.size();
    }
    query_batch_result r;
    r.k = k;
    r.counts.resize(queries.size());
    r.scores.assign(queries.size() * k, 0.f);
    r.docids.assign(queries.size() * k, 0xffffffffu);
    check(ds2i_gpu_query_batch_docids(index.handle(), wdata ? wdata->handle() : nullptr, op, k, terms.data(), offsets.data(), queries.size(),
                                      r.counts.data(), r.scores.data(), r.docids.data(), &r.elapsed_ms));
    return r;
}

// the index (and wand data) replicated over several GPUs of this process; a batch is sharded over them (ds2i_gpu_group_*)
class gpu_group {
public:
    gpu_group(const char* index_path, const char* index_type, const char* wand_path, int ngpus) : m_h(nullptr) {
        check(ds2i_gpu_group_open(index_path, index_type, wand_path, nullptr, ngpus, &m_h));
    }
    explicit gpu_group(ds2i_gpu_group* adopted) : m_h(adopted) {}
    ~gpu_group() { ds2i_gpu_group_close(m_h); }
    gpu_group(gpu_group const&) = delete;
    gpu_group& operator=(gpu_group const&) = delete;
    int size() const { return ds2i_gpu_group_size(m_h); }
    ds2i_gpu_group* handle() const { return m_h; }
private:
    ds2i_gpu_group* m_h;
};

inline query_batch_result run_batch(gpu_group const& group, int op, std::vector<term_id_vec> const& queries, uint32_t k = 10) {
    std::vector<uint32_t> terms;
    std::vector<uint64_t> offsets(queries.size() + 1, 0);
    for (size_t i = 0; i < queries.size(); ++i) {
        terms.insert(terms.end(), queries[i].begin(), queries[i].end());
        offsets[i + 1] = terms.size();
    }
    query_batch_result r;
    r.k = k;
    r.counts.resize(queries.size());
    r.scores.assign(queries.size() * k, 0.f);
    r.docids.assign(queries.size() * k, 0xffffffffu);
    check(ds2i_gpu_group_query_batch(group.handle(), op, k, terms.data(), offsets.data(), queries.size(), r.counts.data(), r.scores.data(),
                                     r.docids.data(), &r.elapsed_ms));
    return r;
}

inline query_batch_result run_batch(gpu_index const& index, gpu_wand_data const& wdata, std::string const& op_name,
                                    std::vector<term_id_vec> const& queries, uint32_t k = 10) {
    int op = ds2i_gpu_op_from_name(op_name.c_str());
    if (op < 0) throw std::invalid_argument("Unsupported query type: " + op_name);
    return run_batch(index, wdata.handle() ? &wdata : nullptr, op, queries, k);
}

// ---- QueryOperator concept: uint64_t operator()(Index const&, term_id_vec); topk() -------------
template <int OP, bool RANKED>
class gpu_query_operator {
public:
    gpu_query_operator() : m_wdata(nullptr), m_k(10) {}                              // and_query / or_query (queries.hpp:35,88)
    gpu_query_operator(gpu_wand_data const& wdata, uint64_t k) : m_wdata(&wdata), m_k(uint32_t(k)) {}   // ranked ones (:204-207)
    uint64_t operator()(gpu_index const& index, term_id_vec const& terms) {
        std::vector<term_id_vec> one(1, terms);
        query_batch_result r = run_batch(index, m_wdata, OP, one, m_k);
        if (RANKED) { m_topk = r.topk(0); m_topk_docids = r.topk_docids(0); }
        return r.counts[0];
    }
    std::vector<float> const& topk() const { return m_topk; }
    std::vector<uint32_t> const& topk_docids() const { return m_topk_docids; }     // extension: the reference keeps scores only
private:
    gpu_wand_data const* m_wdata;
    uint32_t m_k;
    std::vector<float> m_topk;
    std::vector<uint32_t> m_topk_docids;
};

typedef gpu_query_operator<DS2I_OP_AND, false> gpu_and_query;
typedef gpu_query_operator<DS2I_OP_AND_FREQ, false> gpu_and_freq_query;
typedef gpu_query_operator<DS2I_OP_OR, false> gpu_or_query;
typedef gpu_query_operator<DS2I_OP_OR_FREQ, false> gpu_or_freq_query;
typedef gpu_query_operator<DS2I_OP_RANKED_AND, true> gpu_ranked_and_query;
typedef gpu_query_operator<DS2I_OP_WAND, true> gpu_wand_query;
typedef gpu_query_operator<DS2I_OP_MAXSCORE, true> gpu_maxscore_query;
typedef gpu_query_operator<DS2I_OP_RANKED_OR, true> gpu_ranked_or_query;

}  // namespace ds2i_gpu
