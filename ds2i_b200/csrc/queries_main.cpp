// queries_gpu — the `queries` front end of ds2i (queries.cpp:124-153) over the B200 path.
// Same contract: queries_gpu <index_type> <op[:op...]> <index_file> [<wand_file>] < queries.txt
//   * one query per stdin line, whitespace separated term ids (read_query, queries.hpp:15-27);
//   * ops: and, and_freq, or, or_freq, ranked_and, wand, maxscore (+ ranked_or); ranked ops need the wand file;
//   * unknown type / op -> a log line on stderr, exit code 0 (queries.cpp:119-121,148-150);
//   * stdout: one stats_line JSON object per op with the reference's keys type/query/avg/q50/q90/q95 (µs per
//     query; the batch is evaluated in one launch, so every quantile equals the mean) plus qps and batch_ms;
//   * --dump <file> additionally writes per-query counts and top-k scores (the reference discards them).
// The timing protocol follows op_perftest (queries.cpp:13-62): 3 passes over the query log, the first discarded.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "ds2i_gpu.hpp"

using namespace ds2i_gpu;

static bool read_query(term_id_vec& ret, std::istream& is) {
    ret.clear();
    std::string line;
    if (!std::getline(is, line)) return false;
    std::istringstream iline(line);
    term_id_type t;
    while (iline >> t) ret.push_back(t);
    return true;
}

int main(int argc, const char** argv) {
    std::string dump_path;
    std::vector<const char*> args;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--dump") && i + 1 < argc) { dump_path = argv[++i]; continue; }
        args.push_back(argv[i]);
    }
    if (args.size() < 3) {
        std::cerr << "Usage: " << argv[0] << " <index type> <query type> <index filename> [<wand data filename>] [--dump <file>]" << std::endl;
        return 0;
    }
    std::string type = args[0], ops = args[1];
    const char* index_filename = args[2];
    const char* wand_filename = args.size() > 3 ? args[3] : nullptr;

    std::vector<term_id_vec> queries;
    term_id_vec q;
    while (read_query(q, std::cin)) queries.push_back(q);

    try {
        std::unique_ptr<gpu_index> index;
        try {
            index.reset(new gpu_index(index_filename, type.c_str()));
        } catch (std::exception const& e) {
            std::cerr << "ERROR: Unknown type " << type << " (" << e.what() << ")" << std::endl;
            return 0;
        }
        std::cerr << "Performing " << type << " queries" << std::endl;
        gpu_wand_data wdata;
        if (wand_filename) wdata.open(wand_filename);
        FILE* dump = dump_path.empty() ? nullptr : fopen(dump_path.c_str(), "wb");

        size_t start = 0;
        while (start <= ops.size()) {
            size_t end = ops.find(':', start);
            if (end == std::string::npos) end = ops.size();
            std::string t = ops.substr(start, end - start);
            start = end + 1;
            std::cerr << "Query type: " << t << std::endl;
            int op = ds2i_gpu_op_from_name(t.c_str());
            bool ranked = op >= DS2I_OP_RANKED_AND;
            if (op < 0 || (ranked && !wand_filename)) {
                std::cerr << "Unsupported query type: " << t << std::endl;
                continue;
            }
            double ms_sum = 0;
            query_batch_result r;
            const int runs = 2;
            for (int run = 0; run <= runs; ++run) {                 // first pass is a warm-up (queries.cpp:25-35)
                r = run_batch(*index, wand_filename ? &wdata : nullptr, op, queries, 10);
                if (run) ms_sum += r.elapsed_ms;
            }
            double batch_ms = ms_sum / runs;
            double avg_us = queries.empty() ? 0 : batch_ms * 1000.0 / queries.size();
            std::cerr << "---- " << type << " " << t << "\nMean: " << avg_us << "\n";
            printf("{\"type\": \"%s\", \"query\": \"%s\", \"avg\": %g, \"q50\": %g, \"q90\": %g, \"q95\": %g, \"qps\": %g, \"batch_ms\": %g}\n",
                   type.c_str(), t.c_str(), avg_us, avg_us, avg_us, avg_us, batch_ms > 0 ? queries.size() / (batch_ms * 1e-3) : 0.0, batch_ms);
            if (dump) {
                for (size_t i = 0; i < queries.size(); ++i) {
                    fwrite(&r.counts[i], 8, 1, dump);
                    fwrite(&r.scores[i * r.k], 4, r.k, dump);
                }
            }
        }
        if (dump) fclose(dump);
    } catch (std::exception const& e) {
        std::cerr << "ERROR: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
