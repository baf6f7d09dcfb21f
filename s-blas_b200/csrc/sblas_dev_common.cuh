/* sblas_dev_common.cuh -- device helpers shared by the SpMV kernels (sm_100a). */
#pragma once
#include <stdint.h>
#include "sblas_device.h"

namespace sblas {

constexpr unsigned kFull = 0xffffffffu;

/* write one finished row: edge rows (split between segments) keep their raw sum
 * for the ordered merge (reference merge: dspmv_mgpu_v1.cu:235-248,
 * dspmv_mgpu_v2.cu:385-441); beta == 0 does not read y (csrmv convention). */
__device__ __forceinline__ void emit_row(const sblas_seg_args &a, int r, double s)
{
    if (r == a.skip_first) {
        a.edge[0] = s;
    } else if (r == a.skip_last) {
        a.edge[1] = s;
    } else {
        double out = a.alpha * s;
        if (a.beta != 0.0) out += a.beta * a.y[r];
        a.y[r] = out;
    }
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
    return v;
}

/* ---- mbarrier / bulk-copy (TMA) PTX wrappers */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

/* all wrappers take 32-bit shared-window addresses (smem_u32 once, integer math after) */
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    /* try_wait suspends the warp in hardware up to the time hint, so the loop rarely spins */
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
/* 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP).
 * L2 evict-first: the streamed matrix must not push x out of L2. */
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
/* Hand a stage back to the producer only once the values read from it have ARRIVED in registers.
 * `witness` is the result of arithmetic on every value the warp loaded from the stage; making the
 * arrive conditional on it gives nvcc and ptxas a true dependency, so the arrive cannot be
 * scheduled while one of those shared-memory loads is still in flight.  (It was, in SASS: arrive
 * issued between the last LDS and the FMA that consumes it.  With scattered x gathers the LSU queue
 * gets deep enough for such a load to return AFTER the next bulk copy has rewritten the stage --
 * seen as one wrong part in ~1e-3 of the long rows of a test matrix.)  The condition is always
 * true: a computed double is never a signalling NaN, and 0x7ff0dead is one. */
__device__ __forceinline__ void release_after(uint32_t empty_bar, int lane, double witness)
{
    __syncwarp();
#ifdef SBLAS_UNSAFE_EARLY_RELEASE
    /* the pre-fix behaviour, built ONLY into lib/libsblas_spmv_unsafe.so for the regression test
     * tests/test_spmv_gpu.py::test_stage_release_hazard_regression (the arrive does not depend on the loads) */
    (void)witness;
    if (lane == 0) mbar_arrive(empty_bar);
#else
    if (lane == 0 && __double2hiint(witness) != 0x7ff0dead) mbar_arrive(empty_bar);
#endif
}
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace sblas
