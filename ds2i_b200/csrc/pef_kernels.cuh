// Partitioned Elias-Fano (opt_index) device path — placeholder until the PEF enumerator lands.
#pragma once
#include <string>
#include <vector>
#include "query_kernels.cuh"

namespace ds2i_gpu {

struct PefListDir { uint64_t n; };

struct PefIndexHost {
    uint64_t size = 0, num_docs = 0, device_bytes = 0;
    std::vector<PefListDir> host_dir;
    int load(const uint8_t*, size_t, std::string& err) { err = "opt index: not built yet"; return -4; }
};

inline int pef_launch_query(PefIndexHost&, DevWand, DevBatch const&, int, uint32_t, int, int, std::string& err) { err = "opt index: not built yet"; return -4; }
inline int pef_decode_lists(PefIndexHost&, const uint32_t*, uint32_t, const uint64_t*, uint32_t*, uint32_t*, int, std::string& err) { err = "opt index: not built yet"; return -4; }
inline int pef_next_geq(PefIndexHost&, const uint32_t*, uint32_t, const uint64_t*, const uint64_t*, uint64_t*, uint64_t*, uint32_t*, int, std::string& err) { err = "opt index: not built yet"; return -4; }

}  // namespace ds2i_gpu
