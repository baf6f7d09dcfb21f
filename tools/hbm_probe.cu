// tools/hbm_probe.cu -- read-only HBM bandwidth ceilings on this B200 (diagnostic, not product):
//   (a) plain LDG.128 streaming sum, (b) 1-D bulk-copy (TMA) ring with NO consumer work.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/hbm_probe tools/hbm_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(256) ldg_sum(const int4* __restrict__ p, size_t n16, double* out, int unroll_dummy)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    int acc = 0;
    for (; i + 7 * stride < n16; i += 8 * stride) {
        int4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w) : "l"(p + i + k * stride));
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
    }
    if (acc == 0x7fffffff) out[0] = acc;
}

template <int STAGES, int BYTES>
__global__ void __launch_bounds__(64) tma_ring(const char* __restrict__ p, size_t ntile, double* out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * BYTES);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            uint32_t a = (uint32_t)__cvta_generic_to_shared(&full[s]);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(a));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (tid != 0) return;
    // single thread: issue STAGES copies ahead, wait in order (no consumer work at all)
    size_t j = blockIdx.x; int issued = 0, done = 0; int acc = 0;
    size_t cnt = 0; for (size_t q = j; q < ntile; q += gridDim.x) ++cnt;
    while (done < (int)cnt) {
        while (issued < (int)cnt && issued - done < STAGES) {
            int s = issued % STAGES;
            uint32_t bar = (uint32_t)__cvta_generic_to_shared(&full[s]);
            uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem + (size_t)s * BYTES);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(BYTES));
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst), "l"(p + (j + (size_t)issued * gridDim.x) * BYTES), "r"(BYTES), "r"(bar) : "memory");
            ++issued;
        }
        int s = done % STAGES; uint32_t par = (done / STAGES) & 1;
        uint32_t bar = (uint32_t)__cvta_generic_to_shared(&full[s]);
        uint32_t ok = 0;
        while (!ok) asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0,1,0,q; }" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
        acc += smem[(size_t)s * BYTES];
        ++done;
    }
    if (acc == 0x7fffffff) out[0] = acc;
}

int main()
{
    const size_t bytes = (size_t)14 << 30;
    char* d; double* out;
    CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&out, 8));
    CK(cudaMemset(d, 1, bytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int blocks_per_sm : {4, 8, 16, 32}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            ldg_sum<<<148 * blocks_per_sm, 256>>>((const int4*)d, bytes / 16, out, 0);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("ldg128 x8 unroll, %2d CTAs/SM: %.3f ms  %.1f GB/s\n", blocks_per_sm, ms, bytes / ms / 1e6);
    }
    {
        constexpr int B = 24576;
        auto run = [&](auto kern, int stages, int ctas, const char* name) {
            int smem = stages * B + 64;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                kern<<<148 * ctas, 64, smem>>>(d, bytes / B, out);
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            }
            printf("%s: %.3f ms  %.1f GB/s (%s)\n", name, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        };
        run(tma_ring<3, B>, 3, 2, "tma ring 24KB x3 stages x2 CTA/SM");
        run(tma_ring<4, B>, 4, 2, "tma ring 24KB x4 stages x2 CTA/SM");
        run(tma_ring<8, B>, 8, 1, "tma ring 24KB x8 stages x1 CTA/SM");
        run(tma_ring<2, B>, 2, 4, "tma ring 24KB x2 stages x4 CTA/SM");
        run(tma_ring<2, B>, 2, 3, "tma ring 24KB x2 stages x3 CTA/SM");
    }
    {
        char* d2; CK(cudaMalloc(&d2, bytes / 2));
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0); cudaMemcpyAsync(d2, d, bytes / 2, cudaMemcpyDeviceToDevice); cudaEventRecord(e1);
            cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("cudaMemcpy D2D %.1f GB: %.3f ms  %.1f GB/s (read+write)\n", bytes / 2 / 1e9, ms, bytes / ms / 1e6);
    }
    return 0;
}
