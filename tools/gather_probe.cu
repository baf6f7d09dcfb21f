// tools/gather_probe.cu -- which path gathers scattered 8/16-byte elements of an L2-resident vector fastest?
// (diagnostic for the scattered-column workloads, DESIGN.md section 6: circuit5m / rail4284 sit on the L1
// wavefront rate of LDG gathers.)  Variants, all reading `n_gather` random elements of x (44 MB):
//   ldg64   : one LDG.64 per lane, 8 independent gathers per thread in flight (what the SpMV kernels do)
//   ldg64x1 : same addresses, but every LDG instruction has ONE active lane group of 2 (fewer lines per request)
//   bulk16  : cp.async.bulk 16-byte copies global->shared issued by every lane, completion on an mbarrier
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/gather_probe tools/gather_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t a) { a ^= a >> 16; a *= 0x7feb352dU; a ^= a >> 15; a *= 0x846ca68bU; a ^= a >> 16; return a; }

__global__ void __launch_bounds__(256) ldg64_kernel(const double* __restrict__ x, uint32_t nx, long long per_thread, double* out)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    uint32_t h = hash32(tid * 2654435761u + 1);
    for (long long it = 0; it < per_thread; it += 8) {
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { h = hash32(h + k); v[k] = __ldg(x + (h % nx)); }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k];
    }
    if (acc == 1.2345) out[0] = acc;
}

// every lane issues 16-byte bulk copies into its own slots of shared memory; a warp waits on its mbarrier per batch
template <int BATCH>
__global__ void __launch_bounds__(256) bulk16_kernel(const double* __restrict__ x, uint32_t nx, long long per_thread, double* out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                    // one per warp
    double2* buf = reinterpret_cast<double2*>(smem + 128) + (size_t)warp * 32 * BATCH;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[warp]);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
    __syncwarp();
    asm volatile("fence.mbarrier_init.release.cluster;");
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t h = hash32(tid * 2654435761u + 1);
    double acc = 0.0;
    uint32_t parity = 0;
    for (long long it = 0; it < per_thread; it += BATCH) {
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(32 * BATCH * 16));
        __syncwarp();
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            h = hash32(h + k);
            const double* src = x + ((h % nx) & ~1u);
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&buf[k * 32 + lane]);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                         :: "r"(dst), "l"(src), "r"(bar) : "memory");
        }
        uint32_t ok = 0;
        while (!ok) asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0,1,0,q; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        parity ^= 1;
#pragma unroll
        for (int k = 0; k < BATCH; ++k) acc += buf[k * 32 + lane].x;
        __syncwarp();
    }
    if (acc == 1.2345) out[0] = acc;
}

int main()
{
    const uint32_t nx = 5558326;
    double *x, *out;
    CK(cudaMalloc(&x, (size_t)nx * 8 + 64)); CK(cudaMalloc(&out, 8));
    CK(cudaMemset(x, 0, (size_t)nx * 8 + 64));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const long long per_thread = 512;
    for (int ctas : {4, 8}) {
        const int grid = 148 * ctas;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0); ldg64_kernel<<<grid, 256>>>(x, nx, per_thread, out); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        }
        const double g = (double)grid * 256 * per_thread;
        printf("ldg64   %d CTAs/SM: %.3f ms  %.2f G gathers/s  (%.3f per cycle per SM @1.9GHz)\n", ctas, ms, g / ms / 1e6, g / ms / 1e6 / 148 / 1.9);
    }
    {
        constexpr int B = 8;
        const int smem = 128 + 8 * 32 * B * 16;
        cudaFuncSetAttribute(bulk16_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int ctas : {2, 4, 6}) {
            const int grid = 148 * ctas;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0); bulk16_kernel<B><<<grid, 256, smem>>>(x, nx, per_thread, out); cudaEventRecord(e1);
                CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
            }
            const double g = (double)grid * 256 * per_thread;
            printf("bulk16 batch %d, %d CTAs/SM: %.3f ms  %.2f G gathers/s  (%.3f per cycle per SM)  %s\n", B, ctas, ms, g / ms / 1e6, g / ms / 1e6 / 148 / 1.9, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
