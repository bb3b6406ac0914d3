// Warp-level building blocks for the sm_100a query path: TMA bulk staging of compressed bytes into
// shared memory, unaligned bit/byte reads from the staged window, warp scans.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ds2i_gpu {

constexpr unsigned FULL = 0xffffffffu;

// The one dynamic shared-memory window of every kernel in this library.  De-inlined device
// functions address it by byte offset, so the compiler keeps emitting LDS/STS (a generic pointer
// parameter would turn every access into a generic LD/ST).
extern __shared__ __align__(16) uint8_t g_smem[];
__device__ __forceinline__ uint32_t* smem_words(uint32_t byte_off) { return reinterpret_cast<uint32_t*>(g_smem + byte_off); }
__device__ __forceinline__ uint32_t smem_offset(const void* p) {
    return uint32_t(reinterpret_cast<const uint8_t*>(p) - g_smem);
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D TMA (cp.async.bulk global -> shared::cta) --------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    // src, dst and size must be multiples of 16 bytes
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_addr(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// ---- unaligned reads ---------------------------------------------------------------------------
// 32 bits starting at byte offset `off` of a word-aligned shared-memory window
__device__ __forceinline__ uint32_t lds_u32(const uint32_t* win, uint32_t off) {
    uint32_t w = off >> 2;
    return __funnelshift_r(win[w], win[w + 1], (off & 3u) * 8u);
}
__device__ __forceinline__ uint32_t lds_u8(const uint32_t* win, uint32_t off) {
    return (win[off >> 2] >> ((off & 3u) * 8u)) & 0xffu;
}
// `len` (0..32) bits starting at absolute bit position `bit` of the window, LSB-first
__device__ __forceinline__ uint32_t lds_bits(const uint32_t* win, uint32_t bit, uint32_t len) {
    uint32_t w = bit >> 5;
    uint32_t v = __funnelshift_r(win[w], win[w + 1], bit & 31u);
    return len >= 32 ? v : (v & ((1u << len) - 1u));
}
// unaligned little-endian u32 in global memory (block_maxs / block_endpoints are byte-aligned only,
// block_posting_list.hpp:43,49,289,296)
__device__ __forceinline__ uint32_t ldg_u32_unaligned(const uint8_t* p) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* base = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    uint32_t sh = uint32_t(a & 3u) * 8u;
    uint32_t lo = __ldg(base);
    if (sh == 0) return lo;
    uint32_t hi = __ldg(base + 1);
    return __funnelshift_r(lo, hi, sh);
}

// ---- warp scans ----------------------------------------------------------------------------------
// shfl.up's predicate output says whether the source lane exists, so a scan step is two instructions
// (SHFL.UP + predicated IADD) instead of shuffle + compare + select + add.
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            ".reg .u32 t;\n"
            "shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n"
            "@p add.u32 %0, %0, t;\n"
            "}\n"
            : "+r"(v)
            : "r"(d));
    }
    return v;
}

}  // namespace ds2i_gpu
