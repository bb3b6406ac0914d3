// Walks every posting list of an index through the C++ adapters exactly as code written against ds2i's Index /
// document_enumerator concepts would (verify_collection.hpp:9-54 style): next() scan, then a next_geq pass.
// Prints: <lists> <postings> <sum of docid + freq over all postings> <next_geq checksum>
#include <cstdio>
#include "ds2i_gpu.hpp"

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: enumerator_walk <index_type> <index_file>\n"); return 1; }
    ds2i_gpu::gpu_index index(argv[2], argv[1]);
    unsigned long long postings = 0, sum = 0, geq = 0;
    for (size_t t = 0; t < index.size(); ++t) {
        ds2i_gpu::gpu_index::document_enumerator e = index[t];
        if (e.size() != index.list_size(ds2i_gpu::term_id_type(t)) || e.position() != 0) return 2;
        for (; e.docid() < index.num_docs(); e.next()) { sum += e.docid() + e.freq(); ++postings; }
        if (e.position() != e.size()) return 3;
        e.next_geq(index.num_docs());                       // stays past the end
        if (e.docid() != index.num_docs()) return 4;
        e.reset();
        // every 5th docid + 1 as the bound: lands on the following posting (or past the end)
        uint64_t prev = e.docid();
        for (uint64_t i = 0; i < e.size(); i += 5) {
            e.move(i);
            uint64_t d = e.docid();
            e.next_geq(d + 1);
            if (e.docid() <= d || e.position() != i + 1) return 5;
            e.next_geq(prev);                               // a smaller bound never moves the cursor back
            if (e.position() != i + 1) return 6;
            geq += e.docid();
            prev = d;
        }
    }
    std::printf("%zu %llu %llu %llu\n", index.size(), postings, sum, geq);
    return 0;
}
