/* sblas_spmm.cu -- hand-written sm_100a kernels for C = alpha*A*B + beta*C, A CSR (int32), B and C dense
 * column-major, double: what cusparseDcsrmm does inside the reference's cusparse_mgpu_csrmm[_omp]
 * (spmm/src/dspmm_mgpu_baseline.cu:193-208 and :436-451).  SURVEY.md section 8f-2.
 *
 * What bounds it: every entry a_ij needs row j of B for all nd columns of the GPU's slice -- nd*8 bytes
 * out of L1/L2 per 12 streamed bytes of A -- so the work sits on the on-chip gather bandwidth, not on HBM
 * and not on the FP64 pipe; tensor cores do not apply (no dense operand tile is reused).  The layout is
 * chosen for that gather:
 *   - B is transposed once per product into ROW-major Bt (k x ldbt) on the device, so row j of B is one
 *     contiguous run (column-major B would cost one cache line per column and entry);
 *   - a CTA owns 32 consecutive rows of A and a chunk of up to 128 columns; a warp takes one row at a
 *     time, lane l holds CPL adjacent columns (one 8/16-byte load per lane: 32 lanes read 256/512/1024
 *     contiguous bytes of Bt[j]); the row's (col, val) pairs are staged 32 at a time through shared
 *     memory and broadcast (one 16-byte LDS per entry), CPL independent FMA chains per lane;
 *   - the 32 x chunk results go through shared memory and are written to column-major C with lanes
 *     across ROWS (256 contiguous bytes per column), alpha / beta applied there (beta == 0 does not read C);
 *   - a warp takes G = 4 consecutive rows as ONE entry stream (batches of 32, the next batch in flight), so
 *     short rows cost one load latency per group, not per row;
 *   - rows longer than kLongRow entries are cut into segments of kSegLen entries at plan time: a CTA per
 *     (segment, chunk) writes partial sums, spmm_segreduce_kernel adds a row's segments in ascending order
 *     (deterministic, no atomics) -- a 1.3 M-entry hub row becomes 158 CTAs instead of one.
 * Column chunks are the slow grid dimension, so one chunk of Bt (k x chunk x 8 bytes) stays L2-resident
 * while A streams past it once per chunk.
 */
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int kRowsPerCta = 32;
constexpr int kWarps = 8;
constexpr int kLongRow = 2048;      /* rows above the plan's threshold (this, or kLongRowFew for matrices with few rows) */
constexpr int kLongRowFew = 256;    /* are cut into segments of kSegLen entries */
constexpr int kSegLen = 2048;

struct __align__(16) Entry { double v; int c; int pad; };

template <int CPL>
__device__ __forceinline__ void load_b(const double *p, double (&b)[CPL])
{
    if (CPL == 1) b[0] = __ldg(p);
    else if (CPL == 2) { const double2 t = __ldg(reinterpret_cast<const double2 *>(p)); b[0] = t.x; b[1] = t.y; }
    else {
        const double2 t = __ldg(reinterpret_cast<const double2 *>(p));
        const double2 u = __ldg(reinterpret_cast<const double2 *>(p) + 1);
        b[0] = t.x; b[1] = t.y; b[CPL > 2 ? 2 : 0] = u.x; b[CPL > 3 ? 3 : 0] = u.y;
    }
}

/* One warp streams the entries [lo, hi) of its rows in batches of 32 (coalesced (col, val) loads staged through
 * shared memory, the next batch's loads issued before the current one is consumed) and accumulates
 * acc[q] += val * Bt[col][c0 + lane*CPL + q].  `bnd` (shared, warp-private) holds the row boundaries of the warp's
 * G rows: whenever the stream crosses bnd[cur + 1] the finished row's sums go to tile[(row0 + cur) * pitch + ...]
 * and the accumulators restart -- short rows cost one entry-load latency per GROUP of rows, not per row. */
template <int CPL>
__device__ __forceinline__ void stream_rows(const int *__restrict__ col, const double *__restrict__ val,
                                            const double *__restrict__ bt, long long ldbt, int lo, int hi,
                                            const int *bnd, int G, Entry *stage, int lane, bool live, double *tile,
                                            int pitch, int row0)
{
    double acc[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) acc[q] = 0.0;
    int cur = 0;
    int nb = bnd[1];
    double nv = 0.0;
    int nc = 0;
    if (lo + lane < hi) { nv = __ldg(val + lo + lane); nc = __ldg(col + lo + lane); }
    for (int base = lo; base < hi; base += 32) {
        stage[lane].v = nv; stage[lane].c = nc;
        const int nidx = base + 32 + lane;
        if (nidx < hi) { nv = __ldg(val + nidx); nc = __ldg(col + nidx); }      /* in flight during this batch */
        __syncwarp();
        const int cnt = min(32, hi - base);
#pragma unroll 4
        for (int e = 0; e < cnt; ++e) {
            while (base + e >= nb) {                            /* warp-uniform: the stream enters the next row */
#pragma unroll
                for (int q = 0; q < CPL; ++q) { tile[(row0 + cur) * pitch + lane * CPL + q] = acc[q]; acc[q] = 0.0; }
                ++cur;
                nb = bnd[cur + 1];
            }
            const Entry en = stage[e];                          /* broadcast: one wavefront */
            if (live) {
                double b[CPL];
                load_b<CPL>(bt + (long long)en.c * ldbt, b);
#pragma unroll
                for (int q = 0; q < CPL; ++q) acc[q] = fma(en.v, b[q], acc[q]);
            }
        }
        __syncwarp();
    }
    for (; cur < G; ++cur) {                                     /* the last row with entries, then empty rows */
#pragma unroll
        for (int q = 0; q < CPL; ++q) { tile[(row0 + cur) * pitch + lane * CPL + q] = acc[q]; acc[q] = 0.0; }
    }
}

template <int CPL>
__global__ void __launch_bounds__(kWarps * 32) spmm_rows_kernel(int m, const int *__restrict__ rowptr,
                                                                const int *__restrict__ col,
                                                                const double *__restrict__ val,
                                                                const double *__restrict__ Bt, long long ldbt,
                                                                double *__restrict__ C, long long ldc, int nd,
                                                                double alpha, double beta, int long_thr)
{
    constexpr int CH = 32 * CPL;                       /* columns per chunk */
    constexpr int PITCH = CH + 1;
    constexpr int G = kRowsPerCta / kWarps;            /* consecutive rows per warp */
    extern __shared__ __align__(16) unsigned char smem[];
    double *tile = reinterpret_cast<double *>(smem);                                   /* [32][PITCH] */
    Entry *stage_all = reinterpret_cast<Entry *>(smem + ((sizeof(double) * kRowsPerCta * PITCH + 15) & ~(size_t)15));
    __shared__ int bnd_all[kWarps][G + 2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Entry *stage = stage_all + warp * 32;
    int *bnd = bnd_all[warp];
    const int r0 = blockIdx.x * kRowsPerCta;
    const int c0 = blockIdx.y * CH;
    const int ncols = min(CH, nd - c0);
    const bool live = lane * CPL < ncols;              /* chunks are padded to CPL columns: ldbt covers them */
    const double *bt = Bt + c0 + lane * CPL;
    /* my G rows: boundaries, with rows that are long (left to the segment kernels) or past m emptied */
    const int g0 = r0 + warp * G;
    if (lane <= G) bnd[lane] = __ldg(rowptr + min(g0 + lane, m));
    __syncwarp();
    bool any_long = false;
#pragma unroll
    for (int q = 0; q < G; ++q) any_long |= (bnd[q + 1] - bnd[q]) > long_thr;
    if (!any_long) {
        stream_rows<CPL>(col, val, bt, ldbt, bnd[0], bnd[G], bnd, G, stage, lane, live, tile, PITCH, warp * G);
    } else {
        for (int q = 0; q < G; ++q) {                  /* rare: a long row sits in the group -> row by row */
            const int lo = bnd[q], hi = bnd[q + 1];
            __shared__ int one_all[kWarps][3];
            int *one = one_all[warp];
            if (lane == 0) { one[0] = lo; one[1] = (hi - lo) > long_thr ? lo : hi; one[2] = one[1]; }
            __syncwarp();
            stream_rows<CPL>(col, val, bt, ldbt, one[0], one[1], one, 1, stage, lane, live, tile, PITCH, warp * G + q);
            __syncwarp();
        }
    }
    __syncthreads();
    /* 32 x chunk results -> column-major C with lanes across ROWS; long rows are written by spmm_segreduce_kernel */
    const int row = r0 + lane;
    bool mine = row < m;
    if (mine) mine = (__ldg(rowptr + row + 1) - __ldg(rowptr + row)) <= long_thr;
    for (int c = warp; c < ncols; c += kWarps) {
        if (mine) {
            double *dst = C + (long long)(c0 + c) * ldc + row;
            double out = alpha * tile[lane * PITCH + c];
            if (beta != 0.0) out += beta * *dst;
            *dst = out;
        }
    }
}

/* Long rows (more than kLongRow entries) are cut into segments of kSegLen entries at plan time.  A CTA per
 * (segment, column chunk): its eight warps take interleaved batches of the segment, meet in shared memory in warp
 * order and write the segment's partial sums to part[seg][0..ldp); spmm_segreduce_kernel then adds the segments of
 * a row in ascending order (deterministic, no atomics) and applies alpha / beta. */
template <int CPL>
__global__ void __launch_bounds__(kWarps * 32) spmm_segment_kernel(const int *__restrict__ seg_lo,
                                                                   const int *__restrict__ seg_hi,
                                                                   const int *__restrict__ col,
                                                                   const double *__restrict__ val,
                                                                   const double *__restrict__ Bt, long long ldbt,
                                                                   double *__restrict__ part, long long ldp, int nd)
{
    constexpr int CH = 32 * CPL;
    __shared__ double psum[kWarps][CH + 1];
    __shared__ Entry stage_all[kWarps * 32];
    __shared__ int bnd_all[kWarps][3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.y * CH;
    const int ncols = min(CH, nd - c0);
    const bool live = lane * CPL < ncols;
    const int lo = seg_lo[blockIdx.x], hi = seg_hi[blockIdx.x];
    /* warp w takes the contiguous eighth [lo + w*len8, ...) of the segment (rounded to 32 entries) */
    const int len8 = ((hi - lo + kWarps - 1) / kWarps + 31) & ~31;
    const int wlo = min(lo + warp * len8, hi), whi = min(wlo + len8, hi);
    int *bnd = bnd_all[warp];
    if (lane == 0) { bnd[0] = wlo; bnd[1] = whi; bnd[2] = whi; }
    __syncwarp();
    stream_rows<CPL>(col, val, Bt + c0 + lane * CPL, ldbt, wlo, whi, bnd, 1, stage_all + warp * 32, lane, live,
                     &psum[0][0], CH + 1, warp);
    __syncthreads();
    for (int c = threadIdx.x; c < ncols; c += kWarps * 32) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += psum[w][c];
        part[(long long)blockIdx.x * ldp + c0 + c] = s;
    }
}

/* thread per (long row, column): C = alpha * (sum of the row's segment partials, ascending) + beta * C */
__global__ void __launch_bounds__(256) spmm_segreduce_kernel(const int *__restrict__ long_rows,
                                                             const int *__restrict__ row_seg, int nlong,
                                                             const double *__restrict__ part, long long ldp,
                                                             double *__restrict__ C, long long ldc, int nd, double alpha,
                                                             double beta)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nlong * nd) return;
    const int i = (int)(t / nd), c = (int)(t - (long long)i * nd);
    double s = 0.0;
    for (int g = row_seg[i]; g < row_seg[i + 1]; ++g) s += part[(long long)g * ldp + c];
    double *dst = C + (long long)c * ldc + long_rows[i];
    double out = alpha * s;
    if (beta != 0.0) out += beta * *dst;
    *dst = out;
}

/* Bt[j][c] = B[c*ldb + j], 32 x 32 tiles through shared memory; columns [nd, ldbt) of Bt are zero-filled */
__global__ void __launch_bounds__(256) transpose_b_kernel(const double *__restrict__ B, long long ldb, int k, int nd,
                                                          double *__restrict__ Bt, long long ldbt)
{
    __shared__ double t[32][33];
    const int j0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int cc = ty; cc < 32; cc += 8) {
        const int c = c0 + cc, j = j0 + tx;
        t[cc][tx] = (c < nd && j < k) ? __ldg(B + (long long)c * ldb + j) : 0.0;
    }
    __syncthreads();
    for (int jj = ty; jj < 32; jj += 8) {
        const int j = j0 + jj, c = c0 + tx;
        if (j < k && c < ldbt) Bt[(long long)j * ldbt + c] = t[tx][jj];
    }
}

template <int CPL>
cudaError_t launch_cpl(int m, int nd, const int *rowptr, const int *col, const double *val, const double *Bt,
                       long long ldbt, double *C, long long ldc, double alpha, double beta, const int *long_rows,
                       const int *row_seg, int nlong, const int *seg_lo, const int *seg_hi, int nseg, double *part,
                       int long_thr, cudaStream_t s)
{
    constexpr int CH = 32 * CPL;
    const int smem = (int)(((sizeof(double) * kRowsPerCta * (CH + 1) + 15) & ~(size_t)15) + sizeof(Entry) * kWarps * 32);
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(spmm_rows_kernel<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    const unsigned chunks = (unsigned)((nd + CH - 1) / CH);
    const dim3 grid((unsigned)((m + kRowsPerCta - 1) / kRowsPerCta), chunks);
    spmm_rows_kernel<CPL><<<grid, kWarps * 32, smem, s>>>(m, rowptr, col, val, Bt, ldbt, C, ldc, nd, alpha, beta, long_thr);
    if (nlong > 0 && nseg > 0) {
        const long long ldp = ((long long)nd + 3) & ~3LL;
        spmm_segment_kernel<CPL><<<dim3((unsigned)nseg, chunks), kWarps * 32, 0, s>>>(seg_lo, seg_hi, col, val, Bt, ldbt, part, ldp, nd);
        const long long total = (long long)nlong * nd;
        spmm_segreduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(long_rows, row_seg, nlong, part, ldp, C, ldc, nd, alpha, beta);
    }
    return cudaGetLastError();
}

}  // namespace

/* rows above the threshold go to the segment kernels.  With few rows (less than ~8 CTAs of 32 rows per SM) the row
 * kernel alone cannot fill the GPU, so more of the matrix is cut into segments (a short-wide matrix like
 * rail4284: 4,284 rows of ~2,600 entries). */
extern "C" int sblas_spmm_long_row_threshold(int m) { return m < 148 * 8 * kRowsPerCta ? kLongRowFew : kLongRow; }
extern "C" int sblas_spmm_segment_length(void) { return kSegLen; }

/* row pitch of Bt for nd columns: whole 4-column groups, so that every lane's 16/32-byte load is aligned
 * and the padded columns of the last chunk exist (they are zero) */
extern "C" long long sblas_spmm_bt_pitch(int nd) { return ((long long)nd + 3) & ~3LL; }

extern "C" cudaError_t sblas_launch_transpose_b(const double *d_B, long long ldb, int k, int nd, double *d_Bt,
                                                cudaStream_t s)
{
    if (k <= 0 || nd <= 0) return cudaSuccess;
    const long long ldbt = sblas_spmm_bt_pitch(nd);
    const dim3 grid((unsigned)((k + 31) / 32), (unsigned)((ldbt + 31) / 32));
    transpose_b_kernel<<<grid, 256, 0, s>>>(d_B, ldb, k, nd, d_Bt, ldbt);
    return cudaGetLastError();
}

/* C (m x nd, column-major, ld ldc) = alpha * A * B + beta * C with B given as row-major Bt (pitch
 * sblas_spmm_bt_pitch(nd)).  long_rows[nlong] = the rows of A holding more than the long-row threshold,
 * cut into the segments [seg_lo[g], seg_hi[g]) of at most sblas_spmm_segment_length() entries, row i owning
 * segments [row_seg[i], row_seg[i+1]); part = nseg x sblas_spmm_bt_pitch(nd) doubles of scratch; long_thr = the
 * threshold the lists were built with (sblas_spmm_long_row_threshold(m)). */
extern "C" cudaError_t sblas_launch_spmm(int m, int nd, const int *rowptr, const int *col, const double *val,
                                         const double *d_Bt, double *d_C, long long ldc, double alpha, double beta,
                                         const int *long_rows, const int *row_seg, int nlong, const int *seg_lo,
                                         const int *seg_hi, int nseg, double *part, int long_thr, cudaStream_t s)
{
    if (m <= 0 || nd <= 0) return cudaSuccess;
    const long long ldbt = sblas_spmm_bt_pitch(nd);
    if (nd > 64) return launch_cpl<4>(m, nd, rowptr, col, val, d_Bt, ldbt, d_C, ldc, alpha, beta, long_rows, row_seg, nlong, seg_lo, seg_hi, nseg, part, long_thr, s);
    if (nd > 32) return launch_cpl<2>(m, nd, rowptr, col, val, d_Bt, ldbt, d_C, ldc, alpha, beta, long_rows, row_seg, nlong, seg_lo, seg_hi, nseg, part, long_thr, s);
    return launch_cpl<1>(m, nd, rowptr, col, val, d_Bt, ldbt, d_C, ldc, alpha, beta, long_rows, row_seg, nlong, seg_lo, seg_hi, nseg, part, long_thr, s);
}
