/* sblas_sptrans.cu -- hand-written sm_100a kernels for CSR -> CSC (sparse transposition), SURVEY.md section 8f-4.
 * Replaces the cusparseCsr2cscEx2 call of the reference's kernal_sptrans
 * (sptrans/sptrans_v1/src/sptrans_kernal.h:228-262) and its two composition kernels (:12-78).
 *
 * The result is THE csc of the host reference (sptrans/sptrans_v1/src/tranpose.h:3-40): inside a column the
 * entries keep their CSR order (rows ascending, duplicates in input order), so the output is unique and the
 * comparison is bit-for-bit.  That order is a STABLE sort of the entries by column, done here as a
 * least-significant-digit radix sort of (column, CSR position) pairs, 8 bits per pass, every pass
 *     digit histogram per 4096-entry tile  ->  exclusive scan over (digit, tile)  ->  stable scatter
 * with the rank of an entry inside its tile taken from warp match masks in tile order: integer work only,
 * deterministic, no atomics on the data path.  All HBM-bound streaming (8 B in + 8 B out per entry and pass,
 * ceil(log2(n)/8) passes), then one gather of (row, value) through the sorted positions.
 */
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int kT = 256;              /* threads per CTA */
constexpr int kItems = 16;           /* entries per thread and tile */
constexpr int kTile = kT * kItems;   /* 4096 */
constexpr unsigned kFull = 0xffffffffu;

/* hist[digit * ntile + tile] = number of keys of the tile with that digit */
__global__ void __launch_bounds__(kT) rs_hist_kernel(const int *__restrict__ keys, long long n, int shift,
                                                     int *__restrict__ hist, int ntile)
{
    __shared__ int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * kTile;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const long long p = base + i * kT + threadIdx.x;
        if (p < n) atomicAdd(&h[(keys[p] >> shift) & 255], 1);          /* integer counts: order does not matter */
    }
    __syncthreads();
    hist[(long long)threadIdx.x * ntile + blockIdx.x] = h[threadIdx.x];
}

/* stable scatter of one pass: entry p of the tile goes to offs[digit][tile] + (entries of the tile with the same
 * digit that come before p).  The tile is walked in rounds of 256 consecutive entries; inside a round the rank is
 * (same-digit entries in earlier warps) + (same-digit lanes below mine), from __match_any_sync masks. */
__global__ void __launch_bounds__(kT) rs_scatter_kernel(const int *__restrict__ keys, const int *__restrict__ vals,
                                                        int *__restrict__ keys_out, int *__restrict__ vals_out,
                                                        long long n, int shift, const int *__restrict__ offs, int ntile)
{
    __shared__ int run[256];                 /* next free slot per digit */
    __shared__ int cnt[kT / 32][256];        /* per warp: entries of this round per digit */
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    run[threadIdx.x] = offs[(long long)threadIdx.x * ntile + blockIdx.x];
    const long long base = (long long)blockIdx.x * kTile;
    for (int i = 0; i < kItems; ++i) {
        for (int w = 0; w < kT / 32; ++w) cnt[w][threadIdx.x] = 0;
        __syncthreads();
        const long long p = base + i * kT + threadIdx.x;
        const bool live = p < n;
        int key = 0, val = 0, d = 256 + lane;                 /* dead lanes: a digit nobody shares */
        if (live) { key = keys[p]; val = vals[p]; d = (key >> shift) & 255; }
        const unsigned same = __match_any_sync(kFull, d);
        const int below = __popc(same & ((1u << lane) - 1u));
        if (live && below == 0) cnt[warp][d] = __popc(same);  /* the lowest lane of every digit group */
        __syncthreads();
        if (live) {
            int before = 0;
            for (int w = 0; w < warp; ++w) before += cnt[w][d];
            const int dst = run[d] + before + below;
            keys_out[dst] = key;
            vals_out[dst] = val;
        }
        __syncthreads();
        int tot = 0;
        for (int w = 0; w < kT / 32; ++w) tot += cnt[w][threadIdx.x];
        run[threadIdx.x] += tot;
        __syncthreads();
    }
}

/* exclusive scan of an int array, three phases: tiles of 2048, the tile sums (recursively), add back */
constexpr int kScanTile = 2048;
__global__ void __launch_bounds__(256) scan_tiles_kernel(int *__restrict__ data, long long n, int *__restrict__ sums)
{
    __shared__ int s[256];
    const long long base = (long long)blockIdx.x * kScanTile + threadIdx.x * 8;
    int v[8], tot = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = (base + i < n) ? data[base + i] : 0; tot += v[i]; }
    s[threadIdx.x] = tot;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        const int t = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    int run = s[threadIdx.x] - tot;
#pragma unroll
    for (int i = 0; i < 8; ++i) { if (base + i < n) data[base + i] = run; run += v[i]; }
    if (threadIdx.x == 255 && sums) sums[blockIdx.x] = s[255];
}
__global__ void __launch_bounds__(256) scan_add_kernel(int *__restrict__ data, long long n, const int *__restrict__ sums)
{
    const long long base = (long long)blockIdx.x * kScanTile + threadIdx.x * 8;
    const int add = sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < 8; ++i) if (base + i < n) data[base + i] += add;
}

/* rowidx[p] = row_base + (row holding CSR position p): a CTA per 256 rows, row pointers staged in shared memory */
__global__ void __launch_bounds__(256) expand_rows_kernel(const int *__restrict__ rowptr, int m, int row_base,
                                                          int *__restrict__ rowidx, int *__restrict__ pos)
{
    __shared__ int rp[257];
    const int r0 = blockIdx.x * 256;
    const int nr = min(256, m - r0);
    for (int i = threadIdx.x; i <= nr; i += 256) rp[i] = rowptr[r0 + i];
    __syncthreads();
    for (int p = rp[0] + threadIdx.x; p < rp[nr]; p += 256) {
        int lo = 0, hi = nr;                       /* last i with rp[i] <= p */
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (rp[mid] <= p) lo = mid; else hi = mid; }
        rowidx[p] = row_base + r0 + lo;
        pos[p] = p;
    }
}

/* column pointer from the sorted keys: colptr[c] = first sorted position whose key is >= c */
__global__ void __launch_bounds__(256) colptr_kernel(const int *__restrict__ keys, long long nnz, int n, int *__restrict__ colptr)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (nnz == 0) { if (q <= n) colptr[q] = 0; return; }
    if (q >= nnz) return;
    const int k = keys[q], kp = q > 0 ? keys[q - 1] : -1;
    for (int c = kp + 1; c <= k; ++c) colptr[c] = (int)q;
    if (q == nnz - 1) for (int c = k + 1; c <= n; ++c) colptr[c] = (int)nnz;
}

/* out_row[dst] = rowidx[pos[q]], out_val[dst] = val[pos[q]], dst = q (+ base[key[q]] when composing row blocks) */
__global__ void __launch_bounds__(256) gather_kernel(const int *__restrict__ keys, const int *__restrict__ pos,
                                                     const int *__restrict__ rowidx, const double *__restrict__ val,
                                                     long long nnz, const int *__restrict__ base, int *__restrict__ out_row,
                                                     double *__restrict__ out_val)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nnz) return;
    const int p = pos[q];
    const long long dst = base ? (long long)base[keys[q]] + q : q;
    out_row[dst] = rowidx[p];
    out_val[dst] = val[p];
}

/* composition of the row blocks of several GPUs (sptrans_kernal.h:12-78 computes the same two things):
 *   gcolptr[c] = sum_d colptr_d[c];   base_d[c] = sum_{d' < d} colptr_d'[c+1] + sum_{d' > d} colptr_d'[c]
 * so that entry q of block d (column c) lands at base_d[c] + q.  ptrs = [ndev][n+1]. */
__global__ void __launch_bounds__(256) compose_kernel(const int *__restrict__ ptrs, int ndev, int n, int *__restrict__ gcolptr,
                                                      int *__restrict__ bases)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n) return;
    int g = 0;
    for (int d = 0; d < ndev; ++d) g += ptrs[(long long)d * (n + 1) + c];
    gcolptr[c] = g;
    if (c == n) return;
    int lower = 0;                                  /* sum over d' < d of colptr_d'[c+1] */
    int upper = g;                                  /* sum over d' >= d of colptr_d'[c]  */
    for (int d = 0; d < ndev; ++d) {
        upper -= ptrs[(long long)d * (n + 1) + c];
        bases[(long long)d * n + c] = lower + upper;
        lower += ptrs[(long long)d * (n + 1) + c + 1];
    }
}

cudaError_t exclusive_scan(int *data, long long n, int *scratch, cudaStream_t s)
{
    /* scratch: room for the tile sums of every level (n/2048 + n/2048^2 + ... + 8 ints) */
    if (n <= 0) return cudaSuccess;
    const long long nt = (n + kScanTile - 1) / kScanTile;
    if (nt == 1) {
        scan_tiles_kernel<<<1, 256, 0, s>>>(data, n, nullptr);
        return cudaGetLastError();
    }
    scan_tiles_kernel<<<(unsigned)nt, 256, 0, s>>>(data, n, scratch);
    cudaError_t e = exclusive_scan(scratch, nt, scratch + nt + 8, s);
    if (e != cudaSuccess) return e;
    scan_add_kernel<<<(unsigned)nt, 256, 0, s>>>(data, n, scratch);
    return cudaGetLastError();
}

}  // namespace

/* ints of scratch the conversion of nnz entries needs: two (key, position) buffer pairs, the expanded row index,
 * the (digit, tile) histogram and the scan levels */
extern "C" long long sblas_csr2csc_scratch_ints(long long nnz)
{
    const long long ntile = (nnz + kTile - 1) / kTile + 1;
    const long long hist = 256 * ntile;
    return 5 * (nnz + 8) + hist + hist / 1024 + 4096;
}

/* CSR (m x n, nnz entries; row r of the block is global row row_base + r) -> CSC on one GPU.
 * Outputs: colptr[n+1]; sorted_keys (the column of every CSC position, needed by the multi-GPU composition;
 * may be NULL); rows / vals written at out_row[base[col] + q] when base != NULL (composition), else at q. */
extern "C" cudaError_t sblas_launch_csr2csc(int m, int n, long long nnz, const int *d_rowptr, const int *d_col,
                                            const double *d_val, int row_base, int *d_colptr, int **sorted_keys,
                                            int **sorted_pos, int **rowidx, int *scratch, cudaStream_t s)
{
    const long long ntile = (nnz + kTile - 1) / kTile;
    int *kA = scratch, *kB = kA + (nnz + 8), *pA = kB + (nnz + 8), *pB = pA + (nnz + 8), *ridx = pB + (nnz + 8);
    int *hist = ridx + (nnz + 8), *scan_scratch = hist + 256 * (ntile + 1);
    if (nnz > 0) {
        cudaError_t e = cudaMemcpyAsync(kA, d_col, (size_t)nnz * sizeof(int), cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return e;
        expand_rows_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(d_rowptr, m, row_base, ridx, pA);
        int bits = 1;
        while (bits < 31 && (1LL << bits) < (long long)n) ++bits;
        for (int shift = 0; shift < bits; shift += 8) {
            rs_hist_kernel<<<(unsigned)ntile, kT, 0, s>>>(kA, nnz, shift, hist, (int)ntile);
            e = exclusive_scan(hist, 256 * ntile, scan_scratch, s);
            if (e != cudaSuccess) return e;
            rs_scatter_kernel<<<(unsigned)ntile, kT, 0, s>>>(kA, pA, kB, pB, nnz, shift, hist, (int)ntile);
            int *t = kA; kA = kB; kB = t;
            t = pA; pA = pB; pB = t;
        }
    }
    const long long nq = nnz > 0 ? nnz : (long long)n + 1;
    colptr_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, s>>>(kA, nnz, n, d_colptr);
    if (sorted_keys) *sorted_keys = kA;
    if (sorted_pos) *sorted_pos = pA;
    if (rowidx) *rowidx = ridx;
    return cudaGetLastError();
}

extern "C" cudaError_t sblas_launch_csc_gather(const int *keys, const int *pos, const int *rowidx, const double *val,
                                               long long nnz, const int *base, int *out_row, double *out_val, cudaStream_t s)
{
    if (nnz <= 0) return cudaSuccess;
    gather_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, s>>>(keys, pos, rowidx, val, nnz, base, out_row, out_val);
    return cudaGetLastError();
}

extern "C" cudaError_t sblas_launch_csc_compose(const int *ptrs, int ndev, int n, int *gcolptr, int *bases, cudaStream_t s)
{
    compose_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, s>>>(ptrs, ndev, n, gcolptr, bases);
    return cudaGetLastError();
}
