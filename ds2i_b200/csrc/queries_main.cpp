// queries_gpu — the `queries` front end of ds2i (queries.cpp:124-153) over the B200 path.
// Same contract: queries_gpu <index_type> <op[:op...]> <index_file> [<wand_file>] < queries.txt
//   * one query per stdin line, whitespace separated term ids (read_query, queries.hpp:15-27);
//   * ops: and, and_freq, or, or_freq, ranked_and, wand, maxscore (+ ranked_or); ranked ops need the wand file;
//   * unknown type / op -> a log line on stderr, exit code 0 (queries.cpp:119-121,148-150);
//   * stdout: one stats_line JSON object per op with the reference's keys type/query/avg/q50/q90/q95: WALL-CLOCK µs per
//     query like the reference (gettimeofday around query_op, queries.cpp:26-34) — here the wall time of the whole
//     batch call (host preparation + H2D + kernels + D2H) divided by the number of queries; a batch is one launch, so
//     the quantiles are taken over the timed passes, not over queries.  Extra keys: qps (wall), batch_ms (wall),
//     kernel_ms (CUDA-event time of the device work only), passes;
//   * --gpus N shards the query log over N GPUs of this process (ds2i_gpu_query_batch_multi; index replicated);
//   * --dump <file> additionally writes per-query counts and top-k scores (the reference discards them).
// The timing protocol follows op_perftest (queries.cpp:13-62): 3 passes over the query log, the first discarded.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
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
    int gpus = 1;
    std::vector<const char*> args;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--dump") && i + 1 < argc) { dump_path = argv[++i]; continue; }
        if (!strcmp(argv[i], "--gpus") && i + 1 < argc) { gpus = std::max(1, atoi(argv[++i])); continue; }
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
        // an unknown index type is a log line and exit code 0 like in the reference (queries.cpp:148-150); any other
        // failure to open the index (missing file, format error, CUDA error) is an error and exits non-zero
        std::unique_ptr<gpu_index> index;
        std::unique_ptr<gpu_wand_data> wdata;
        std::unique_ptr<gpu_group> group;
        {
            ds2i_gpu_index* h = nullptr;
            ds2i_gpu_group* gh = nullptr;
            int rc = gpus == 1 ? ds2i_gpu_index_open_file(index_filename, type.c_str(), 0, &h)
                               : ds2i_gpu_group_open(index_filename, type.c_str(), wand_filename, nullptr, gpus, &gh);
            if (rc == DS2I_E_UNSUPPORTED && ds2i_gpu_index_type_known(type.c_str()) == 0) {
                std::cerr << "ERROR: Unknown type " << type << std::endl;
                return 0;
            }
            check(rc);
            if (h) index.reset(new gpu_index(h));
            if (gh) group.reset(new gpu_group(gh));
        }
        std::cerr << "Performing " << type << " queries" << std::endl;
        if (wand_filename && gpus == 1) wdata.reset(new gpu_wand_data(wand_filename));
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
            query_batch_result r;
            const int runs = 2;
            std::vector<double> wall_ms;                            // per timed pass: host clock around the whole call
            double kernel_ms = 0;
            for (int run = 0; run <= runs; ++run) {                 // first pass is a warm-up (queries.cpp:25-35)
                auto t0 = std::chrono::steady_clock::now();
                if (gpus == 1) r = run_batch(*index, wdata.get(), op, queries, 10);
                else r = run_batch(*group, op, queries, 10);
                double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                if (run) { wall_ms.push_back(ms); kernel_ms += r.elapsed_ms / runs; }
            }
            std::sort(wall_ms.begin(), wall_ms.end());
            const double nq = double(std::max<size_t>(queries.size(), 1));
            double batch_ms = 0;
            for (double v : wall_ms) batch_ms += v / wall_ms.size();
            auto quant = [&](double q) { return wall_ms[std::min(wall_ms.size() - 1, size_t(q * wall_ms.size()))] * 1000.0 / nq; };
            double avg_us = batch_ms * 1000.0 / nq;
            std::cerr << "---- " << type << " " << t << "\nMean: " << avg_us << "\n50% quantile: " << quant(0.5) << "\n90% quantile: " << quant(0.9)
                      << "\n95% quantile: " << quant(0.95) << "\n";
            printf("{\"type\": \"%s\", \"query\": \"%s\", \"avg\": %g, \"q50\": %g, \"q90\": %g, \"q95\": %g, \"qps\": %g, \"batch_ms\": %g, \"kernel_ms\": %g, "
                   "\"passes\": %d, \"gpus\": %d}\n",
                   type.c_str(), t.c_str(), avg_us, quant(0.5), quant(0.9), quant(0.95), batch_ms > 0 ? queries.size() / (batch_ms * 1e-3) : 0.0, batch_ms,
                   kernel_ms, runs, gpus);
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
