// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into or called by the product path.
//
// Driver over the UNMODIFIED ds2i reference headers (/root/reference, read-only), compiled by
// oracle/Makefile into oracle/_ref/.  The reference `queries` binary prints timings only and
// discards results (queries.cpp:28-29), so parity needs this results-dumping driver; it calls the
// reference's own index classes, enumerators and query operators and adds no algorithm of its own.
//
//   ref_tool dump    <type> <index> <wand> <queries> <out.bin> <op[:op..]> [k] [max_queries]
//   ref_tool lists   <type> <index> <terms.txt> <out.bin>
//   ref_tool nextgeq <type> <index> <spec.bin> <out.bin>
//   ref_tool bench   <type> <index> <wand> <queries> <op> <threads> [max_queries] [passes]
//   ref_tool profile <type> <index> <wand> <queries> <op> <out.bin> [max_queries]   (block types only)
//   ref_tool scan    <type> <index> <terms.txt | all | first:N> <threads> [passes]   timed next()/docid()/freq() scan of whole lists
//                    (list i -> thread i % n, profile_queries.cpp:21-39); prints postings/s and sum(docid) / sum(freq) checksums
//   ref_tool geqbench <type> <index> <spec.bin> <threads> [passes]   timed next_geq sweeps (same spec as nextgeq)
//   ref_tool info    <type> <index>
//   ref_tool mkmixed block_optpfor <index> <out block_mixed index> [seed]   block types drawn per block; the reference's
//                    own transformation path (optimal_hybrid_index.cpp:266-292) writes the block_mixed index
//
// ops: and, or, ranked_and, wand, maxscore, ranked_or  (queries.hpp:35-591)

#include <iostream>
#include <fstream>
#include <thread>
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <sstream>

#include <succinct/mapper.hpp>

#include "index_types.hpp"
#include "wand_data.hpp"
#include "queries.hpp"
#include "util.hpp"

using namespace ds2i;

static std::vector<term_id_vec> load_queries(const char* path, size_t max_q)
{
    std::vector<term_id_vec> queries;
    std::ifstream in(path);
    term_id_vec q;
    while (queries.size() < max_q && read_query(q, in)) queries.push_back(q);
    return queries;
}

static void put_u64(FILE* f, uint64_t v) { fwrite(&v, 8, 1, f); }

template <typename Index>
struct op_runner {
    Index const& index;
    wand_data<> const& wdata;
    uint64_t k;
    ranked_and_query ra;
    wand_query wq;
    maxscore_query ms;
    ranked_or_query ro;

    op_runner(Index const& idx, wand_data<> const& wd, uint64_t k_)
        : index(idx), wdata(wd), k(k_), ra(wd, k_), wq(wd, k_), ms(wd, k_), ro(wd, k_) {}

    // returns count; fills scores (size <= k) for ranked operators
    uint64_t run(std::string const& op, term_id_vec const& q, std::vector<float>& scores)
    {
        scores.clear();
        if (op == "and") return and_query<false>()(index, q);
        if (op == "and_freq") return and_query<true>()(index, q);
        if (op == "or") return or_query<false>()(index, q);
        if (op == "or_freq") return or_query<true>()(index, q);
        uint64_t r;
        if (op == "ranked_and") { r = ra(index, q); scores = ra.topk(); return r; }
        if (op == "wand") { r = wq(index, q); scores = wq.topk(); return r; }
        if (op == "maxscore") { r = ms(index, q); scores = ms.topk(); return r; }
        if (op == "ranked_or") { r = ro(index, q); scores = ro.topk(); return r; }
        throw std::invalid_argument("unknown op " + op);
    }
};

static std::vector<std::string> split_ops(std::string const& s)
{
    std::vector<std::string> out;
    std::string cur;
    for (char c : s) {
        if (c == ':') { out.push_back(cur); cur.clear(); } else cur.push_back(c);
    }
    out.push_back(cur);
    return out;
}

template <typename Index>
int cmd_dump(int argc, char** argv)
{
    // argv: dump type index wand queries out ops [k] [max]
    Index index;
    boost::iostreams::mapped_file_source m(argv[3]);
    succinct::mapper::map(index, m);
    wand_data<> wdata;
    boost::iostreams::mapped_file_source md(argv[4]);
    succinct::mapper::map(wdata, md);
    auto ops = split_ops(argv[7]);
    uint64_t k = argc > 8 ? std::stoull(argv[8]) : 10;
    size_t max_q = argc > 9 ? std::stoull(argv[9]) : size_t(-1);
    auto queries = load_queries(argv[5], max_q);

    FILE* f = fopen(argv[6], "wb");
    if (!f) { perror("out"); return 1; }
    put_u64(f, queries.size());
    put_u64(f, k);
    put_u64(f, ops.size());
    // layout: for op in ops: for q: u64 count, float[k] (zero padded)
    size_t n_threads = std::max<size_t>(1, std::thread::hardware_concurrency());
    for (auto const& op : ops) {
        std::vector<uint64_t> counts(queries.size());
        std::vector<float> scores(queries.size() * k, 0.f);
        std::vector<std::thread> threads;
        for (size_t t = 0; t < n_threads; ++t) {
            threads.emplace_back([&, t]() {
                op_runner<Index> runner(index, wdata, k);
                std::vector<float> s;
                for (size_t i = t; i < queries.size(); i += n_threads) {
                    counts[i] = runner.run(op, queries[i], s);
                    std::copy(s.begin(), s.end(), scores.begin() + i * k);
                }
            });
        }
        for (auto& th : threads) th.join();
        for (size_t i = 0; i < queries.size(); ++i) {
            put_u64(f, counts[i]);
            fwrite(&scores[i * k], 4, k, f);
        }
    }
    fclose(f);
    return 0;
}

template <typename Index>
int cmd_lists(int, char** argv)
{
    // argv: lists type index terms out
    Index index;
    boost::iostreams::mapped_file_source m(argv[3]);
    succinct::mapper::map(index, m);
    std::ifstream tin(argv[4]);
    FILE* f = fopen(argv[5], "wb");
    if (!f) { perror("out"); return 1; }
    uint64_t term;
    std::vector<uint32_t> docs, freqs;
    while (tin >> term) {
        auto e = index[term];
        uint64_t n = e.size();
        docs.resize(n); freqs.resize(n);
        for (uint64_t i = 0; i < n; ++i) {
            docs[i] = uint32_t(e.docid());
            freqs[i] = uint32_t(e.freq());
            e.next();
        }
        if (e.docid() != index.num_docs()) { std::cerr << "sentinel mismatch\n"; return 2; }
        put_u64(f, n);
        fwrite(docs.data(), 4, n, f);
        fwrite(freqs.data(), 4, n, f);
    }
    fclose(f);
    return 0;
}

template <typename Index>
int cmd_nextgeq(int, char** argv)
{
    // spec.bin: u64 nlists; per list: u64 term, u64 nbounds, u64 bounds[nbounds] (non-decreasing)
    // out.bin : per list: per bound: u64 docid, u64 freq (0 when docid == num_docs)
    Index index;
    boost::iostreams::mapped_file_source m(argv[3]);
    succinct::mapper::map(index, m);
    FILE* in = fopen(argv[4], "rb");
    FILE* f = fopen(argv[5], "wb");
    if (!in || !f) { perror("file"); return 1; }
    uint64_t nl;
    if (fread(&nl, 8, 1, in) != 1) return 1;
    std::vector<uint64_t> bounds, out;
    for (uint64_t l = 0; l < nl; ++l) {
        uint64_t term, nb;
        if (fread(&term, 8, 1, in) != 1 || fread(&nb, 8, 1, in) != 1) return 1;
        bounds.resize(nb);
        if (nb && fread(bounds.data(), 8, nb, in) != nb) return 1;
        auto e = index[term];
        out.resize(2 * nb);
        for (uint64_t i = 0; i < nb; ++i) {
            e.next_geq(bounds[i]);
            out[2 * i] = e.docid();
            out[2 * i + 1] = e.docid() < index.num_docs() ? e.freq() : 0;
        }
        fwrite(out.data(), 8, 2 * nb, f);
    }
    fclose(in); fclose(f);
    return 0;
}

template <typename Index>
int cmd_bench(int argc, char** argv)
{
    // argv: bench type index wand queries op threads [max] [passes]
    Index index;
    boost::iostreams::mapped_file_source m(argv[3]);
    succinct::mapper::map(index, m);
    wand_data<> wdata;
    boost::iostreams::mapped_file_source md(argv[4]);
    succinct::mapper::map(wdata, md, succinct::mapper::map_flags::warmup);
    std::string op = argv[6];
    size_t n_threads = std::stoull(argv[7]);
    if (!n_threads) n_threads = std::max<size_t>(1, std::thread::hardware_concurrency());
    size_t max_q = argc > 8 ? std::stoull(argv[8]) : size_t(-1);
    size_t passes = argc > 9 ? std::stoull(argv[9]) : 1;
    auto queries = load_queries(argv[5], max_q);
    // warm-up as queries.cpp:79-88 does
    for (auto const& q : queries) for (auto t : q) index.warmup(t);

    // scheme of profile_queries.cpp:21-39: thread t takes queries t, t+n, ...; one operator per thread
    double best = 1e300;
    uint64_t checksum = 0;
    std::string pass_list;
    for (size_t pass = 0; pass < passes; ++pass) {
        std::vector<uint64_t> sums(n_threads, 0);
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> threads;
        for (size_t t = 0; t < n_threads; ++t) {
            threads.emplace_back([&, t]() {
                op_runner<Index> runner(index, wdata, 10);
                std::vector<float> s;
                uint64_t acc = 0;
                for (size_t i = t; i < queries.size(); i += n_threads) acc += runner.run(op, queries[i], s);
                sums[t] = acc;
            });
        }
        for (auto& th : threads) th.join();
        double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        best = std::min(best, secs);
        char tmp[64]; snprintf(tmp, sizeof tmp, "%s%.6f", pass ? ", " : "", secs); pass_list += tmp;
        checksum = 0;
        for (auto s : sums) checksum += s;
    }
    printf("{\"type\": \"%s\", \"query\": \"%s\", \"threads\": %zu, \"queries\": %zu, \"seconds\": %.6f, "
           "\"qps\": %.3f, \"checksum\": %llu, \"pass_seconds\": [%s]}\n",
           argv[2], op.c_str(), n_threads, queries.size(), best, queries.size() / best,
           (unsigned long long)checksum, pass_list.c_str());
    return 0;
}

// per-query algorithmic-byte accounting with the reference's own block_profiler
// (block_profiler.hpp:9-62; what profile_queries.cpp does, but per query and sized in bytes)
template <typename Codec>
int cmd_profile(int argc, char** argv)
{
    // argv: profile type index wand queries op out [max]
    typedef block_freq_index<Codec, true> Index;
    Index index;
    boost::iostreams::mapped_file_source m(argv[3]);
    succinct::mapper::map(index, m);
    wand_data<> wdata;
    boost::iostreams::mapped_file_source md(argv[4]);
    succinct::mapper::map(wdata, md);
    std::string op = argv[6];
    size_t max_q = argc > 8 ? std::stoull(argv[8]) : size_t(-1);
    auto queries = load_queries(argv[5], max_q);
    FILE* f = fopen(argv[7], "wb");
    if (!f) { perror("out"); return 1; }

    struct list_info {
        block_profiler::counter_type* counters;
        std::vector<uint32_t> docs_bytes, freqs_bytes;
    };
    std::map<uint32_t, list_info> cache;
    op_runner<Index> runner(index, wdata, 10);
    std::vector<float> s;
    put_u64(f, queries.size());
    for (auto const& q : queries) {
        term_id_vec terms = q;
        remove_duplicate_terms(terms);
        uint64_t list_bytes = 0, postings = 0;
        for (auto t : terms) {
            auto e = index[t];
            postings += e.size();
            if (!cache.count(t)) {
                list_info li;
                li.counters = block_profiler::open_list(t, e.num_blocks());
                auto blocks = e.get_blocks();
                std::vector<uint8_t> tmp;
                for (auto const& b : blocks) {
                    tmp.clear(); b.append_docs_block(tmp); li.docs_bytes.push_back(tmp.size());
                    tmp.clear(); b.append_freqs_block(tmp); li.freqs_bytes.push_back(tmp.size());
                }
                cache[t] = li;
            }
            auto& li = cache[t];
            for (size_t b = 0; b < li.docs_bytes.size(); ++b) {
                list_bytes += li.docs_bytes[b] + li.freqs_bytes[b] + 8;
                li.counters[2 * b] = 0; li.counters[2 * b + 1] = 0;
            }
        }
        runner.run(op, q, s);
        uint64_t db = 0, fb = 0, dbytes = 0, fbytes = 0;
        for (auto t : terms) {
            auto& li = cache[t];
            for (size_t b = 0; b < li.docs_bytes.size(); ++b) {
                if (li.counters[2 * b]) { db++; dbytes += li.docs_bytes[b]; }
                if (li.counters[2 * b + 1]) { fb++; fbytes += li.freqs_bytes[b]; }
            }
        }
        put_u64(f, db); put_u64(f, fb); put_u64(f, dbytes); put_u64(f, fbytes);
        put_u64(f, list_bytes); put_u64(f, postings);
    }
    fclose(f);
    return 0;
}

// "decode baseline ... 1 thread and all cores": the reference's own sequential enumeration of whole lists
template <typename Index>
int cmd_scan(int argc, char** argv)
{
    // argv: scan type index spec threads [passes]
    Index index;
    boost::iostreams::mapped_file_source m(argv[3]);
    succinct::mapper::map(index, m);
    std::string spec = argv[4];
    size_t n_threads = std::stoull(argv[5]);
    if (!n_threads) n_threads = std::max<size_t>(1, std::thread::hardware_concurrency());
    size_t passes = argc > 6 ? std::stoull(argv[6]) : 1;
    std::vector<uint64_t> terms;
    if (spec == "all") { for (uint64_t t = 0; t < index.size(); ++t) terms.push_back(t); }
    else if (spec.compare(0, 6, "first:") == 0) { uint64_t n = std::min<uint64_t>(index.size(), std::stoull(spec.substr(6))); for (uint64_t t = 0; t < n; ++t) terms.push_back(t); }
    else { std::ifstream tin(spec); uint64_t t; while (tin >> t) terms.push_back(t); }
    for (auto t : terms) index.warmup(t);
    double best = 1e300;
    uint64_t postings = 0, sum_docs = 0, sum_freqs = 0;
    std::string pass_list;
    for (size_t pass = 0; pass < passes; ++pass) {
        std::vector<uint64_t> n(n_threads, 0), sd(n_threads, 0), sf(n_threads, 0);
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> threads;
        for (size_t t = 0; t < n_threads; ++t) {
            threads.emplace_back([&, t]() {
                uint64_t cnt = 0, a = 0, b = 0;
                for (size_t i = t; i < terms.size(); i += n_threads) {
                    auto e = index[terms[i]];
                    uint64_t len = e.size();
                    for (uint64_t j = 0; j < len; ++j) { a += e.docid(); b += e.freq(); e.next(); }
                    cnt += len;
                }
                n[t] = cnt; sd[t] = a; sf[t] = b;
            });
        }
        for (auto& th : threads) th.join();
        double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        best = std::min(best, secs);
        char tmp[64]; snprintf(tmp, sizeof tmp, "%s%.6f", pass ? ", " : "", secs); pass_list += tmp;
        postings = sum_docs = sum_freqs = 0;
        for (size_t t = 0; t < n_threads; ++t) { postings += n[t]; sum_docs += sd[t]; sum_freqs += sf[t]; }
    }
    printf("{\"type\": \"%s\", \"lists\": %zu, \"threads\": %zu, \"postings\": %llu, \"seconds\": %.6f, \"postings_per_s\": %.1f, "
           "\"sum_docids\": %llu, \"sum_freqs\": %llu, \"pass_seconds\": [%s]}\n", argv[2], terms.size(), n_threads,
           (unsigned long long)postings, best, postings / best, (unsigned long long)sum_docs, (unsigned long long)sum_freqs, pass_list.c_str());
    return 0;
}

// timed next_geq sweeps over the lists of a spec file (the format of `nextgeq`); list i -> thread i % n
template <typename Index>
int cmd_geqbench(int argc, char** argv)
{
    Index index;
    boost::iostreams::mapped_file_source m(argv[3]);
    succinct::mapper::map(index, m);
    FILE* in = fopen(argv[4], "rb");
    if (!in) { perror("spec"); return 1; }
    size_t n_threads = std::stoull(argv[5]);
    if (!n_threads) n_threads = std::max<size_t>(1, std::thread::hardware_concurrency());
    size_t passes = argc > 6 ? std::stoull(argv[6]) : 1;
    uint64_t nl;
    if (fread(&nl, 8, 1, in) != 1) return 1;
    std::vector<uint64_t> terms(nl);
    std::vector<std::vector<uint64_t>> bounds(nl);
    uint64_t calls = 0;
    for (uint64_t l = 0; l < nl; ++l) {
        uint64_t nb;
        if (fread(&terms[l], 8, 1, in) != 1 || fread(&nb, 8, 1, in) != 1) return 1;
        bounds[l].resize(nb);
        if (nb && fread(bounds[l].data(), 8, nb, in) != nb) return 1;
        calls += nb;
    }
    fclose(in);
    double best = 1e300;
    uint64_t checksum = 0;
    for (size_t pass = 0; pass < passes; ++pass) {
        std::vector<uint64_t> acc(n_threads, 0);
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> threads;
        for (size_t t = 0; t < n_threads; ++t) {
            threads.emplace_back([&, t]() {
                uint64_t a = 0;
                for (size_t l = t; l < nl; l += n_threads) {
                    auto e = index[terms[l]];
                    for (uint64_t b : bounds[l]) { e.next_geq(b); a += e.docid(); if (e.docid() < index.num_docs()) a += e.freq(); }
                }
                acc[t] = a;
            });
        }
        for (auto& th : threads) th.join();
        best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        checksum = 0;
        for (auto a : acc) checksum += a;
    }
    printf("{\"type\": \"%s\", \"lists\": %llu, \"threads\": %zu, \"calls\": %llu, \"seconds\": %.6f, \"calls_per_s\": %.1f, \"checksum\": %llu}\n",
           argv[2], (unsigned long long)nl, n_threads, (unsigned long long)calls, best, calls / best, (unsigned long long)checksum);
    return 0;
}

template <typename Index>
int cmd_info(int, char** argv)
{
    Index index;
    boost::iostreams::mapped_file_source m(argv[3]);
    succinct::mapper::map(index, m);
    printf("{\"type\": \"%s\", \"lists\": %zu, \"num_docs\": %llu}\n", argv[2], size_t(index.size()),
           (unsigned long long)index.num_docs());
    return 0;
}

// A block_mixed index can only be created by transformation (mixed_block.hpp:32-36).  The reference chooses the block
// types with its space/time optimiser; for a decode fixture any assignment will do, so the types are drawn per block
// from a seeded LCG and everything else is the reference's code: get_blocks(), mixed_block::block_transformer,
// block_posting_list<mixed_block>::write_blocks, block_freq_index::builder.
static int cmd_mkmixed(int argc, char** argv)
{
    block_optpfor_index in;
    boost::iostreams::mapped_file_source m(argv[3]);
    succinct::mapper::map(in, m);
    uint64_t rng = argc > 5 ? std::stoull(argv[5]) : 20261017ull;
    auto next = [&]() { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return uint32_t(rng >> 33); };
    global_parameters params;
    block_mixed_index::builder builder(in.num_docs(), params);
    typedef block_optpfor_index::document_enumerator::block_data input_block_type;
    typedef mixed_block::block_transformer<input_block_type> output_block_type;
    auto const& possLogs = optpfor_block::codec_type::possLogs;
    uint64_t counts[3] = {0, 0, 0};
    std::vector<uint32_t> vals;
    auto pick = [&](std::vector<uint32_t> const& v, mixed_block::block_type& type, mixed_block::compr_param_type& param) {
        type = mixed_block::block_type(next() % 3);
        param = 0;
        if (type == mixed_block::block_type::pfor) {
            uint32_t mb = FastPFor::maxbits(v.data(), v.data() + v.size());
            uint32_t want = mb > next() % 4 ? mb - next() % 4 : 0;             // a few exceptions now and then
            if (mb > 28 + want) want = mb - 28;                                  // Simple16 codes at most 28 bits
            uint8_t i = 0;
            while (i + 1 < possLogs.size() && possLogs[i] < want) ++i;
            param = i;
        }
        counts[size_t(type)] += 1;
    };
    for (size_t l = 0; l < in.size(); ++l) {
        auto e = in[l];
        auto blocks = e.get_blocks();
        std::vector<output_block_type> out_blocks;
        for (auto const& ib : blocks) {
            mixed_block::block_type dt = mixed_block::block_type::interpolative, ft = dt;
            mixed_block::compr_param_type dp = 0, fp = 0;
            if (ib.size == mixed_block::block_size) {
                ib.decode_doc_gaps(vals); pick(vals, dt, dp);
                ib.decode_freqs(vals); pick(vals, ft, fp);
            }
            out_blocks.emplace_back(ib, dt, ft, dp, fp);
        }
        std::vector<uint8_t> buf;
        block_posting_list<mixed_block>::write_blocks(buf, e.size(), out_blocks);
        builder.add_posting_list(buf);
    }
    block_mixed_index out;
    builder.build(out);
    succinct::mapper::freeze(out, argv[4]);
    printf("{\"lists\": %zu, \"pfor_blocks\": %llu, \"varint_blocks\": %llu, \"interpolative_blocks\": %llu}\n", size_t(in.size()),
           (unsigned long long)counts[0], (unsigned long long)counts[1], (unsigned long long)counts[2]);
    return 0;
}

template <typename Index>
int dispatch(std::string const& cmd, int argc, char** argv)
{
    if (cmd == "dump") return cmd_dump<Index>(argc, argv);
    if (cmd == "lists") return cmd_lists<Index>(argc, argv);
    if (cmd == "nextgeq") return cmd_nextgeq<Index>(argc, argv);
    if (cmd == "bench") return cmd_bench<Index>(argc, argv);
    if (cmd == "info") return cmd_info<Index>(argc, argv);
    if (cmd == "scan") return cmd_scan<Index>(argc, argv);
    if (cmd == "geqbench") return cmd_geqbench<Index>(argc, argv);
    std::cerr << "unknown command " << cmd << std::endl;
    return 1;
}

int main(int argc, char** argv)
{
    if (argc < 4) { std::cerr << "usage: ref_tool <cmd> <type> <index> ..." << std::endl; return 1; }
    std::string cmd = argv[1], type = argv[2];
    try {
        if (cmd == "profile") {
            if (type == "block_optpfor") return cmd_profile<optpfor_block>(argc, argv);
            if (type == "block_varint") return cmd_profile<varint_G8IU_block>(argc, argv);
            if (type == "block_interpolative") return cmd_profile<interpolative_block>(argc, argv);
            if (type == "block_qmx") return cmd_profile<qmx_block>(argc, argv);
            std::cerr << "profile: block index types only" << std::endl;
            return 1;
        }
        if (cmd == "mkmixed") return cmd_mkmixed(argc, argv);
        if (type == "block_mixed") return dispatch<block_mixed_index>(cmd, argc, argv);
        if (type == "block_optpfor") return dispatch<block_optpfor_index>(cmd, argc, argv);
        if (type == "block_varint") return dispatch<block_varint_index>(cmd, argc, argv);
        if (type == "block_interpolative") return dispatch<block_interpolative_index>(cmd, argc, argv);
        if (type == "block_qmx") return dispatch<block_qmx_index>(cmd, argc, argv);
        if (type == "opt") return dispatch<opt_index>(cmd, argc, argv);
        if (type == "uniform") return dispatch<uniform_index>(cmd, argc, argv);
        if (type == "ef") return dispatch<ef_index>(cmd, argc, argv);
        if (type == "single") return dispatch<single_index>(cmd, argc, argv);
    } catch (std::exception const& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 3;
    }
    std::cerr << "unknown index type " << type << std::endl;
    return 1;
}
