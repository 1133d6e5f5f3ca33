// tma_stage.cuh -- staging of the profiled sequences' symbol codes into shared memory with one TMA bulk copy
// (cp.async.bulk global -> shared, completion signalled on an mbarrier; SASS: UBLKCP + SYNCS), sm_90+/sm_100a.
//
// The profiled side (zoe: the sequences a StripedProfile is built from, src/alignment/profile.rs:270-306) is read by
// every step of every sweep, so it lives in shared memory; it is loaded once per CTA.  One elected thread arms the
// barrier with the byte count and issues the copy, everybody waits on the barrier's phase 0.  Size and both addresses
// must be multiples of 16 bytes (the host pads the device buffer; the shared carve-up is 16-byte aligned).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace zoe_cuda {

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ZOE_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra ZOE_MBAR_DONE;\n"
        "bra ZOE_MBAR_WAIT;\n"
        "ZOE_MBAR_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// Copies `bytes` (rounded up to 16) from `src` to `dst` for the whole CTA and returns when the data is visible to
// every thread.  Must be called by all threads of the CTA, once (the barrier is single-use: phase 0).
__device__ __forceinline__ void stage_with_tma(void *dst, const void *src, uint32_t bytes) {
    __shared__ __align__(8) uint64_t zoe_stage_bar;
    const uint32_t padded = (bytes + 15u) & ~15u;
    if (threadIdx.x == 0) mbar_init(&zoe_stage_bar, 1);
    __syncthreads();
    if (padded) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(&zoe_stage_bar, padded);
            // a single bulk copy moves at most 2^20 - 16 bytes; longer inputs go in slices on the same barrier
            uint32_t done = 0;
            while (done < padded) {
                const uint32_t n = min(padded - done, 1u << 19);
                tma_load_1d(reinterpret_cast<uint8_t *>(dst) + done, reinterpret_cast<const uint8_t *>(src) + done, n,
                            &zoe_stage_bar);
                done += n;
            }
        }
        mbar_wait(&zoe_stage_bar, 0);
    }
}

}  // namespace zoe_cuda
