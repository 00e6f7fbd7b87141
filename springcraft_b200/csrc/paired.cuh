// Row-paired operator records shared by the FP64 and FP32 SpMM kernels.
#pragma once
#include "subspace.cuh"

namespace scb {

constexpr int kPairWarps = 16;     // FP64: warps per CTA (128 registers per thread; 24 warps x 80 registers spills)
constexpr int kPairChunk = 16;     // merged contacts staged per warp and round (4 per slot)

template <int D>
struct PairEntry {                 // one merged contact
    double blk[2 * D * D];         // rows of residue 2t (D x D), then of residue 2t+1
    int32_t col;                   // node index of the contact
    int32_t pad[3];
};
static_assert(sizeof(PairEntry<3>) == 160 && sizeof(PairEntry<1>) == 32, "record layout");

template <int D>
struct PairEntry32 {               // single-precision copy used by the early filter iterations
    float blk[2 * D * D];
    int32_t col;
    int32_t pad;
};
static_assert(sizeof(PairEntry32<3>) == 80 && sizeof(PairEntry32<1>) == 16, "record layout");

// TMA bulk copy + mbarrier helpers (SASS: UBLKCP / SYNCS)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

}  // namespace scb
