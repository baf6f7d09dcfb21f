/* sblas_kernels.cu -- hand-written sm_100a kernels for double-precision CSR SpMV,
 * y = alpha*A*x + beta*y, replacing the cusparseDcsrmv / cusparseDcsrmv_mp calls
 * of the reference (spmv/src/dspmv_mgpu_baseline.cu:163, dspmv_mgpu_v1.cu:200,206,
 * dspmv_mgpu_v2.cu:351,357) and the disabled CSR5 back-end
 * (spmv/include/detail/cuda/csr5_spmv_cuda.h).
 *
 * The work is HBM-bound (12 B per nnz streamed once, 2 flop per nnz): no tensor
 * cores.  What matters is bytes in flight per SM, fully coalesced 128/256-bit
 * streaming loads that bypass L1, x served from L1/L2, and an nnz-balanced grid.
 *
 *  spmv_tile_kernel   nnz-balanced tiles (256 threads x IPT nnz).  val is streamed
 *                     with 256-bit and col with 128-bit ld.global.nc.L1::no_allocate
 *                     loads, always aligned because tiles sit on absolute multiples
 *                     of the tile size.  Per tile the reduction strategy adapts to the
 *                     rows inside it: whole tile inside <=2 rows -> block reduction
 *                     from registers (the "block per long row" case); otherwise
 *                     products go to shared memory and G = 1..32 lanes reduce each
 *                     row (thread-per-row for short rows ... warp-per-row).
 *                     Rows that leave a tile are finished by spmv_tile_fixup in a
 *                     fixed order (deterministic; no floating-point atomics).
 *  spmv_vec_kernel    LANES (2..32) lanes per row, for small inputs and as the
 *                     simple cross-check path.
 */
#include <stdint.h>
#include "sblas_dev_common.cuh"
#include "sblas_device.h"

cudaError_t sblas_launch_tma(const sblas_seg_args *a, cudaStream_t s);      /* sblas_spmv_tma.cu */
int sblas_tma_tile_size(void);
cudaError_t sblas_launch_rowtile(const sblas_seg_args *a, int R, int window, cudaStream_t s);
cudaError_t sblas_launch_rowsplit(const sblas_seg_args *a, int G, cudaStream_t s);  /* sblas_spmv_rowtile.cu */

namespace {

using sblas::emit_row;
using sblas::kFull;
constexpr int kThreads = 256;

__device__ __forceinline__ void ldg_nc_i4(const int *p, int &a, int &b, int &c, int &d)
{
    asm("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
        : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
}
/* 256-bit load (LDG.E.256 on sm_100a) */
__device__ __forceinline__ void ldg_nc_d4(const double *p, double &a, double &b, double &c, double &d)
{
    asm("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
        : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

/* ------------------------------------------------------------------ vector kernel */
template <int LANES>
__global__ void __launch_bounds__(kThreads) spmv_vec_kernel(const sblas_seg_args a)
{
    const int lane = threadIdx.x & (LANES - 1);
    const long long gid = ((long long)blockIdx.x * kThreads + threadIdx.x) / LANES;
    const long long nrows = (long long)a.row_hi - a.row_lo + 1;
    const bool live = gid < nrows;
    const int r = a.row_lo + (int)(live ? gid : 0);
    int lo = 0, hi = 0;
    if (live) {
        lo = max(__ldg(a.rowptr + r), a.nz0);
        hi = min(__ldg(a.rowptr + r + 1), a.nz1);
    }
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int k = lo + lane;
    for (; k + 3 * LANES < hi; k += 4 * LANES) {
        const int c0 = __ldg(a.col + k), c1 = __ldg(a.col + k + LANES);
        const int c2 = __ldg(a.col + k + 2 * LANES), c3 = __ldg(a.col + k + 3 * LANES);
        const double v0 = __ldg(a.val + k), v1 = __ldg(a.val + k + LANES);
        const double v2 = __ldg(a.val + k + 2 * LANES), v3 = __ldg(a.val + k + 3 * LANES);
        s0 += v0 * __ldg(a.x + c0);
        s1 += v1 * __ldg(a.x + c1);
        s2 += v2 * __ldg(a.x + c2);
        s3 += v3 * __ldg(a.x + c3);
    }
    for (; k < hi; k += LANES) s0 += __ldg(a.val + k) * __ldg(a.x + __ldg(a.col + k));
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int off = LANES >> 1; off > 0; off >>= 1) s += __shfl_xor_sync(kFull, s, off);
    if (live && lane == 0) emit_row(a, r, s);
}

/* ------------------------------------------------------------------ short-row kernel
 * Thread per row for panels whose rows all hold at most a handful of entries (the plan bins row
 * blocks by their longest row).  Nothing to reduce across lanes and nothing to stage: every load
 * of a thread is independent of its neighbours', 2048 threads per SM keep ~70 KB in flight, and
 * the row pointer / y accesses of a warp are 128 / 256 contiguous bytes.  Rows of exactly two
 * entries on an even offset take val and col as one 16-byte and one 8-byte load. */
constexpr int kShortRows = 2;         /* rows per thread: two independent load chains in flight */
__global__ void __launch_bounds__(kThreads, 6) spmv_short_kernel(const sblas_seg_args a)
{
    const long long nrows = (long long)a.row_hi - a.row_lo + 1;
    const long long g0 = (long long)blockIdx.x * (kThreads * kShortRows) + threadIdx.x;
    int r[kShortRows], lo[kShortRows], hi[kShortRows];
    bool live[kShortRows];
    double yv[kShortRows], acc[kShortRows];
#pragma unroll
    for (int u = 0; u < kShortRows; ++u) {          /* row u of this thread: consecutive threads, consecutive rows */
        const long long g = g0 + (long long)u * kThreads;
        live[u] = g < nrows;
        r[u] = a.row_lo + (int)(live[u] ? g : 0);
        lo[u] = 0; hi[u] = 0;
        if (live[u]) {
            lo[u] = max(__ldg(a.rowptr + r[u]), a.nz0);
            hi[u] = min(__ldg(a.rowptr + r[u] + 1), a.nz1);
        }
    }
#pragma unroll
    for (int u = 0; u < kShortRows; ++u) {
        yv[u] = 0.0;
        if (live[u] && a.beta != 0.0 && r[u] != a.skip_first && r[u] != a.skip_last) yv[u] = a.y[r[u]];
    }
    /* rows of exactly two entries on an even offset: one 16-byte and one 8-byte load each */
    bool pair[kShortRows];
    double2 v2[kShortRows];
    int2 c2[kShortRows];
#pragma unroll
    for (int u = 0; u < kShortRows; ++u) {
        pair[u] = (hi[u] - lo[u] == 2) && ((lo[u] & 1) == 0);
        if (pair[u]) {
            v2[u] = __ldg(reinterpret_cast<const double2 *>(a.val + lo[u]));
            c2[u] = __ldg(reinterpret_cast<const int2 *>(a.col + lo[u]));
        }
    }
#pragma unroll
    for (int u = 0; u < kShortRows; ++u) {
        acc[u] = 0.0;
        if (pair[u]) {
            acc[u] = fma(v2[u].y, __ldg(a.x + c2[u].y), v2[u].x * __ldg(a.x + c2[u].x));
        } else {
            for (int k = lo[u]; k < hi[u]; ++k) acc[u] = fma(__ldg(a.val + k), __ldg(a.x + __ldg(a.col + k)), acc[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < kShortRows; ++u) {
        if (!live[u]) continue;
        if (r[u] == a.skip_first) a.edge[0] = acc[u];
        else if (r[u] == a.skip_last) a.edge[1] = acc[u];
        else a.y[r[u]] = a.alpha * acc[u] + a.beta * yv[u];
    }
}

/* ------------------------------------------------------------------ pipelined vector kernel
 * Warp per row for MEDIUM rows (tens to a few thousand nnz), persistent and software-pipelined:
 * a warp walks its rows (row_lo + w, + nwarps, ...) as a stream of chunks of 32*EPL entries and
 * always has the NEXT chunk's val/col loads (and the next row's bounds and y) in flight while it
 * gathers x and accumulates the current one.  The row structure is used directly, so there is no
 * transpose / flag / scan work at all: ~60 warp instructions per 256 nnz. */
template <int EPL>
__global__ void __launch_bounds__(kThreads, 3) spmv_vecp_kernel(const sblas_seg_args a)
{
    constexpr int CH = 32 * EPL;
    const int lane = threadIdx.x & 31;
    const int nw = (gridDim.x * kThreads) >> 5;
    const int gw = (blockIdx.x * kThreads + threadIdx.x) >> 5;
    const int nrows = a.row_hi - a.row_lo + 1;
    if (gw >= nrows) return;
    const double *__restrict__ xp = a.x;

    auto bounds = [&](int r, int &lo, int &hi) {
        lo = max(__ldg(a.rowptr + r), a.nz0);
        hi = min(__ldg(a.rowptr + r + 1), a.nz1);
    };
    auto load_chunk = [&](int k, int hi, double (&v)[EPL], unsigned (&c)[EPL]) {
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            const int idx = k + lane + 32 * i;
            v[i] = 0.0; c[i] = 0u;
            if (idx < hi) {
                v[i] = __ldg(a.val + idx);            /* streamed once: coalesced 256 B per request */
                c[i] = (unsigned)__ldg(a.col + idx);
            }
        }
    };

    int r = a.row_lo + gw, lo, hi;
    bounds(r, lo, hi);
    int rn = r + nw, lon = 0, hin = 0;
    if (rn <= a.row_hi) bounds(rn, lon, hin);
    double v[EPL];
    unsigned c[EPL];
    load_chunk(lo, hi, v, c);
    const bool has_y = a.beta != 0.0;

    while (true) {
        double yv = 0.0;
        if (has_y && lane == 0 && r != a.skip_first && r != a.skip_last) yv = a.y[r];
        double acc = 0.0;
        const bool more_rows = rn <= a.row_hi;
        int k = lo;
        do {
            /* gather x for the chunk in registers */
            double xv[EPL];
#pragma unroll
            for (int i = 0; i < EPL; ++i) xv[i] = __ldg(xp + c[i]);      /* c == 0 for padding: harmless */
            /* next chunk: rest of this row, else the first chunk of my next row */
            const int kn = k + CH;
            double v2[EPL];
            unsigned c2[EPL];
            if (kn < hi) load_chunk(kn, hi, v2, c2);
            else if (more_rows) load_chunk(lon, hin, v2, c2);
            else {
#pragma unroll
                for (int i = 0; i < EPL; ++i) { v2[i] = 0.0; c2[i] = 0u; }
            }
#pragma unroll
            for (int i = 0; i < EPL; ++i) acc = fma(v[i], xv[i], acc);   /* padding has v == 0 */
#pragma unroll
            for (int i = 0; i < EPL; ++i) { v[i] = v2[i]; c[i] = c2[i]; }
            k = kn;
        } while (k < hi);
        acc = sblas::warp_sum(acc);
        if (lane == 0) {
            if (r == a.skip_first) a.edge[0] = acc;
            else if (r == a.skip_last) a.edge[1] = acc;
            else a.y[r] = a.alpha * acc + a.beta * yv;
        }
        if (!more_rows) break;
        r = rn; lo = lon; hi = hin;
        rn += nw;
        if (rn <= a.row_hi) bounds(rn, lon, hin);
    }
}

/* ------------------------------------------------------------------ tile kernel */
template <int IPT>
struct TileCfg {
    static constexpr int kTile = kThreads * IPT;
    static constexpr int kGroups = IPT / 4;          /* 4 consecutive nnz per thread per group */
    static constexpr int kBBCap = kTile + 8;         /* row boundaries staged per pass        */
    static constexpr int kSmem = kTile * 8 + kBBCap * 4;
};

template <int IPT>
__global__ void __launch_bounds__(kThreads, (IPT <= 8 ? 4 : 2)) spmv_tile_kernel(const sblas_seg_args a)
{
    using Cfg = TileCfg<IPT>;
    constexpr int TILE = Cfg::kTile;
    extern __shared__ __align__(32) unsigned char smem_raw[];
    double *P = reinterpret_cast<double *>(smem_raw);
    int *BB = reinterpret_cast<int *>(smem_raw + (size_t)TILE * 8);
    __shared__ double red[2][kThreads / 32];

    const int j = blockIdx.x;
    const int t = threadIdx.x;
    const long long base = (long long)(a.tile0 + j) * TILE;           /* GPU-local nnz index */
    const int T0 = (int)max((long long)a.nz0, base);
    const int T1 = (int)min((long long)a.nz1, base + TILE);
    const bool full = (T0 == base) && ((long long)T1 == base + TILE);

    /* rows that start in this tile: [rs, re) */
    const int rs = __ldg(a.tstart + j), re = __ldg(a.tstart + j + 1);
    const int nown = re - rs;

    double p[IPT];
    const double *vbase = a.val + base;
    const int *cbase = a.col + base;
    if (full) {
        int c[IPT];
        double v[IPT];
#pragma unroll
        for (int g = 0; g < Cfg::kGroups; ++g) {
            const int e = g * (kThreads * 4) + 4 * t;
            ldg_nc_i4(cbase + e, c[4 * g], c[4 * g + 1], c[4 * g + 2], c[4 * g + 3]);
            ldg_nc_d4(vbase + e, v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
#pragma unroll
        for (int i = 0; i < IPT; ++i) p[i] = v[i] * __ldg(a.x + c[i]);
    } else {
#pragma unroll
        for (int g = 0; g < Cfg::kGroups; ++g) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long pos = base + g * (kThreads * 4) + 4 * t + i;
                double prod = 0.0;
                if (pos >= T0 && pos < T1) prod = __ldg(a.val + pos) * __ldg(a.x + __ldg(a.col + pos));
                p[4 * g + i] = prod;
            }
        }
    }

    /* does the last row that starts here continue into the next tile? */
    bool ext = false;
    if (nown > 0) ext = min(__ldg(a.rowptr + re), a.nz1) > T1;

    if (nown <= 1) {
        /* ---- mode A: at most one row boundary inside the tile: reduce from registers */
        int split = T1;
        if (nown == 1) split = min(max(__ldg(a.rowptr + rs), T0), T1);
        const int lsplit = (int)(split - base);                     /* tile-local */
        double sc = 0.0, so = 0.0;
#pragma unroll
        for (int g = 0; g < Cfg::kGroups; ++g) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = g * (kThreads * 4) + 4 * t + i;
                if (e < lsplit) sc += p[4 * g + i]; else so += p[4 * g + i];
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sc += __shfl_xor_sync(kFull, sc, off);
            so += __shfl_xor_sync(kFull, so, off);
        }
        if ((t & 31) == 0) { red[0][t >> 5] = sc; red[1][t >> 5] = so; }
        __syncthreads();
        if (t == 0) {
            double c = 0.0, o = 0.0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) { c += red[0][w]; o += red[1][w]; }
            a.carry[j] = c;
            if (nown == 1) {
                if (ext) a.tail[j] = o; else emit_row(a, rs, o);
            }
        }
        return;
    }

    /* ---- mode S: products to shared memory, G lanes per row */
#pragma unroll
    for (int g = 0; g < Cfg::kGroups; ++g) {
        const int e = g * (kThreads * 4) + 4 * t;
        *reinterpret_cast<double2 *>(P + e) = make_double2(p[4 * g], p[4 * g + 1]);
        *reinterpret_cast<double2 *>(P + e + 2) = make_double2(p[4 * g + 2], p[4 * g + 3]);
    }
    const int nseg = nown + 1;                 /* segment 0 = the row left open by tile j-1 */
    const int avg = (T1 - T0) / nseg;
    int G = 1;
    while (G < 32 && G * 8 <= avg) G <<= 1;
    const int ngroups = kThreads / G;
    const int grp = t / G, lane = t & (G - 1);

    for (int c0 = 0; c0 < nseg; c0 += Cfg::kBBCap - 1) {
        const int cn = min(nseg - c0, Cfg::kBBCap - 1);
        __syncthreads();                        /* P written / previous pass done with BB */
        for (int idx = t; idx <= cn; idx += kThreads) {
            const int s = c0 + idx;             /* boundary s: start of segment s */
            int b = T0;
            if (s > 0) b = min(max(__ldg(a.rowptr + rs + s - 1), T0), T1);
            BB[idx] = (int)(b - base);
        }
        __syncthreads();
        for (int s0 = 0; s0 < cn; s0 += ngroups) {
            const int s = s0 + grp;
            double acc = 0.0;
            if (s < cn) {
                const int b = BB[s], e = BB[s + 1];
                double acc1 = 0.0;
                int k = b + lane;
                for (; k + G < e; k += 2 * G) { acc += P[k]; acc1 += P[k + G]; }
                if (k < e) acc += P[k];
                acc += acc1;
            }
            for (int off = G >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(kFull, acc, off);
            if (s < cn && lane == 0) {
                const int gs = c0 + s;
                if (gs == 0) {
                    a.carry[j] = acc;
                } else if (gs == nown && ext) {
                    a.tail[j] = acc;
                } else {
                    emit_row(a, rs + gs - 1, acc);
                }
            }
        }
    }
}

/* rows that leave their tile: y[r] = alpha*(tail[j] + carry[j+1] + ... ) + beta*y[r],
 * one thread per tile (the per-tile metadata says at once whether there is anything to do),
 * fixed left-to-right summation order (the CSR5 "calibrator" step,
 * csr5_spmv_cuda.h:313-382, without atomics). */
__global__ void __launch_bounds__(kThreads) spmv_tile_fixup(const sblas_seg_args a, int tile)
{
    const int j = blockIdx.x * kThreads + threadIdx.x;
    if (j >= a.ntile) return;
    const int4 m = __ldg(reinterpret_cast<const int4 *>(a.tmeta) + 2 * j);
    if ((m.w & 1) == 0) return;                       /* no row leaves this tile */
    const int r = m.y - 1;
    const int e = min(__ldg(a.rowptr + r + 1), a.nz1);
    const int jend = (e - 1) / tile - a.tile0;
    double acc = a.tail[j];
    for (int i = j + 1; i <= jend; ++i) acc += a.carry[i];
    emit_row(a, r, acc);
}

/* ------------------------------------------------------------------ plan helpers */
__global__ void tile_rows_kernel(const sblas_seg_args a, int tile, int *tstart)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > a.ntile) return;
    int out;
    if (j == 0) {
        out = a.row_lo;
    } else if (j == a.ntile) {
        out = a.row_hi + 1;
    } else {
        const long long T0 = (long long)(a.tile0 + j) * tile;
        int lo = a.row_lo, hi = a.row_hi + 1;            /* first r in [lo,hi] with rowptr[r] >= T0 */
        while (lo < hi) {
            const int mid = lo + ((hi - lo) >> 1);
            if ((long long)__ldg(a.rowptr + mid) < T0) lo = mid + 1; else hi = mid;
        }
        out = lo;
    }
    tstart[j] = out;
}

/* tmeta[2j]   = {rs, re, start of row rs clamped to the tile (tile end if no row starts),
 *               flags: bit0 = the last row that starts here leaves the tile,
 *                      bit1 = one of the rows that start here is empty,
 *                      bit2 = no chunk holds more than 7 row starts (warp-piece path)}
 * tmeta[2j+1] = 8 x uint16: q_w = how many of the tile's rows start before chunk w
 *               (chunk = tile/8 consecutive entries, one per consumer warp)            */
__global__ void tile_meta_kernel(const sblas_seg_args a, int tile, int4 *tmeta)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.ntile) return;
    const long long base = (long long)(a.tile0 + j) * tile;
    const int T0 = (int)max((long long)a.nz0, base);
    const int T1 = (int)min((long long)a.nz1, base + tile);
    const int rs = a.tstart[j], re = a.tstart[j + 1];
    int start0 = T1, flags = 0;
    unsigned q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (re > rs) {
        start0 = min(max(__ldg(a.rowptr + rs), T0), T1);
        if (min(__ldg(a.rowptr + re), a.nz1) > T1) flags |= 1;
        int prev = __ldg(a.rowptr + rs);
        for (int r = rs + 1; r <= re; ++r) {               /* empty rows among [rs, re) */
            const int cur = __ldg(a.rowptr + r);
            if (cur == prev) { flags |= 2; break; }
            prev = cur;
        }
        const int chunk = tile / 8;
        for (int w = 0; w < 8; ++w) {
            const long long cpos = base + (long long)w * chunk;      /* first r in [rs,re) with start >= cpos */
            int lo = rs, hi = re;
            while (lo < hi) {
                const int mid = lo + ((hi - lo) >> 1);
                if ((long long)max(__ldg(a.rowptr + mid), T0) < cpos) lo = mid + 1; else hi = mid;
            }
            q[w] = (unsigned)min(lo - rs, 65535);
        }
        int most = re - rs - (int)q[7];                    /* row starts of the fullest chunk */
        for (int w = 0; w < 7; ++w) most = max(most, (int)q[w + 1] - (int)q[w]);
        if (most <= 7 && re - rs < 65535) flags |= 4;
    }
    tmeta[2 * j] = make_int4(rs, re, start0, flags);
    tmeta[2 * j + 1] = make_int4((int)(q[0] | (q[1] << 16)), (int)(q[2] | (q[3] << 16)), (int)(q[4] | (q[5] << 16)),
                                 (int)(q[6] | (q[7] << 16)));
}

/* per block of `rb` consecutive rows: the longest row and the (clamped) row pointer at the block
 * start -- what the plan needs to bin row panels by length (adaptive kernel choice per panel) */
__global__ void row_block_stats_kernel(const int *__restrict__ rowptr, int row_lo, int nrows, int rb, int nz0, int nz1,
                                       int *__restrict__ out_max, int *__restrict__ out_ptr)
{
    __shared__ int red[kThreads / 32];
    const int b = blockIdx.x;
    const int r0 = b * rb, r1 = min(r0 + rb, nrows);
    int mx = 0;
    for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x)
        mx = max(mx, __ldg(rowptr + row_lo + r + 1) - __ldg(rowptr + row_lo + r));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(kFull, mx, off));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; ++w) mx = max(mx, red[w]);
        out_max[b] = mx;
        out_ptr[b] = min(max(__ldg(rowptr + row_lo + r0), nz0), nz1);
    }
}

/* smallest and largest column index of a shard: the only part of x its products can read */
__global__ void col_range_kernel(const int *__restrict__ col, long long count, int *__restrict__ out)
{
    int lo = 0x7fffffff, hi = -1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        const int c = __ldg(col + i);
        lo = min(lo, c);
        hi = max(hi, c);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        lo = min(lo, __shfl_xor_sync(kFull, lo, off));
        hi = max(hi, __shfl_xor_sync(kFull, hi, off));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out, lo);
        atomicMax(out + 1, hi);
    }
}

__global__ void rebase_rowptr_kernel(const long long *__restrict__ rp64, long long first_idx, int total,
                                     long long count, int *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    long long v = rp64[i] - first_idx;
    v = v < 0 ? 0 : (v > total ? total : v);
    out[i] = (int)v;
}

__global__ void edge_merge_kernel(const int *__restrict__ mrow, const int *__restrict__ mbeg,
                                  const double *const *__restrict__ msrc, int nmerge, double *y, double alpha,
                                  double beta)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nmerge) return;
    double s = 0.0;
    for (int k = mbeg[i]; k < mbeg[i + 1]; ++k) s += *reinterpret_cast<const volatile double *>(msrc[k]);
    const int r = mrow[i];
    double out = alpha * s;
    if (beta != 0.0) out += beta * y[r];
    y[r] = out;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

/* One warp.  The product counter (epoch) lives in device memory and is advanced HERE, so that a
 * product -- segments, publish, merge -- is the same launch sequence every time (CUDA-graph replay).
 *   1. epoch = ++*epoch_ctr; back-pressure: every owner must have consumed what was written two products ago
 *   2. this rank's own partials (local_edge, nlocal = edge slots) go into its half `epoch & 1` of its OWN table
 *   3. the outgoing ones go straight into the OWNER ranks' tables over NVLink (P2P stores)
 *   4. fence, then raise the owners' arrive flags to `epoch` */
__global__ void edge_publish_kernel(const double *local_edge, int nlocal, const int *out_slot, const int *out_owner,
                                    const long long *out_off, int nout, const int *owners, int nowners,
                                    void *const *peer_bases, long long table_words, int world, int my_rank,
                                    unsigned long long *epoch_ctr)
{
    if (blockIdx.x != 0) return;
    __shared__ unsigned long long s_epoch;
    unsigned long long *mine = reinterpret_cast<unsigned long long *>(peer_bases[my_rank]);
    if (threadIdx.x == 0) {
        const unsigned long long e = *epoch_ctr + 1ull;
        *epoch_ctr = e;
        s_epoch = e;
        for (int k = 0; k < nowners; ++k) {
            const unsigned long long *ack = mine + 2 * table_words + world + owners[k];
            while (ld_acquire_sys(ack) + 2ull < e) { }
        }
    }
    __syncthreads();
    const unsigned long long epoch = s_epoch;
    const int parity = (int)(epoch & 1ull);
    double *own = reinterpret_cast<double *>(mine) + parity * table_words + (long long)my_rank * nlocal;
    for (int i = threadIdx.x; i < nlocal; i += blockDim.x) own[i] = local_edge[i];
    for (int i = threadIdx.x; i < nout; i += blockDim.x) {
        double *dst = reinterpret_cast<double *>(peer_bases[out_owner[i]]) + parity * table_words + out_off[i];
        *reinterpret_cast<volatile double *>(dst) = local_edge[out_slot[i]];
    }
    __threadfence_system();
    __syncthreads();
    for (int k = threadIdx.x; k < nowners; k += blockDim.x) {
        unsigned long long *flag = reinterpret_cast<unsigned long long *>(peer_bases[owners[k]]) + 2 * table_words + my_rank;
        st_release_sys(flag, epoch);
    }
}

__global__ void edge_merge_wait_kernel(const int *__restrict__ mrow, const int *__restrict__ mbeg,
                                       const long long *__restrict__ msrc_off, int nmerge, double *y, double alpha,
                                       double beta, const int *contrib, int ncontrib, void *const *peer_bases,
                                       long long table_words, int world, int my_rank, const unsigned long long *epoch_ctr)
{
    unsigned long long *mine = reinterpret_cast<unsigned long long *>(peer_bases[my_rank]);
    const unsigned long long epoch = *epoch_ctr;              /* advanced by this product's publish kernel */
    if (threadIdx.x == 0) {
        for (int k = 0; k < ncontrib; ++k) {
            const unsigned long long *flag = mine + 2 * table_words + contrib[k];
            while (ld_acquire_sys(flag) < epoch) { }
        }
    }
    __syncthreads();
    const double *table = reinterpret_cast<const double *>(mine) + (long long)(epoch & 1ull) * table_words;
    for (int i = threadIdx.x; i < nmerge; i += blockDim.x) {
        double s = 0.0;
        for (int k = mbeg[i]; k < mbeg[i + 1]; ++k) s += *reinterpret_cast<const volatile double *>(table + msrc_off[k]);
        const int r = mrow[i];
        double out = alpha * s;
        if (beta != 0.0) out += beta * y[r];
        y[r] = out;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        for (int k = 0; k < ncontrib; ++k) {
            unsigned long long *ack = reinterpret_cast<unsigned long long *>(peer_bases[contrib[k]]) + 2 * table_words + world + my_rank;
            st_release_sys(ack, epoch);
        }
    }
}

/* ---- x <- y across the ranks of a one-process-per-GPU job, over peer-mapped memory:
 * an all-gather of the y rows each rank owns into EVERY rank's replica of x, with two rounds of flags
 *   ready:   "I have finished reading x for this product" -- nobody's x may be overwritten before all are ready
 *   written: "my rows are in your x"                       -- the next product starts when all have written
 * flags buffer of every rank (8-byte words): ready[world], written[world]; the counter lives on the device so
 * that the sequence is the same every product. */
__global__ void chain_ready_kernel(void *const *peer_flags, int world, int my_rank, unsigned long long *chain_ctr)
{
    if (blockIdx.x != 0) return;
    __shared__ unsigned long long s_e;
    if (threadIdx.x == 0) { const unsigned long long e = *chain_ctr + 1ull; *chain_ctr = e; s_e = e; }
    __syncthreads();
    const unsigned long long e = s_e;
    __threadfence_system();
    for (int p = threadIdx.x; p < world; p += blockDim.x)
        st_release_sys(reinterpret_cast<unsigned long long *>(peer_flags[p]) + my_rank, e);
    const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(peer_flags[my_rank]);
    for (int p = threadIdx.x; p < world; p += blockDim.x)
        while (ld_acquire_sys(mine + p) < e) { }
}

__global__ void __launch_bounds__(256) chain_copy_kernel(const double *__restrict__ y_src, long long count, long long dst_off,
                                                         void *const *peer_x, int world)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int p = 0; p < world; ++p) {
        double *dst = reinterpret_cast<double *>(peer_x[p]) + dst_off;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) dst[i] = y_src[i];
    }
    __threadfence_system();
}

__global__ void chain_done_kernel(void *const *peer_flags, int world, int my_rank, const unsigned long long *chain_ctr)
{
    if (blockIdx.x != 0) return;
    const unsigned long long e = *chain_ctr;
    __threadfence_system();
    for (int p = threadIdx.x; p < world; p += blockDim.x)
        st_release_sys(reinterpret_cast<unsigned long long *>(peer_flags[p]) + world + my_rank, e);
    const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(peer_flags[my_rank]);
    for (int p = threadIdx.x; p < world; p += blockDim.x)
        while (ld_acquire_sys(mine + world + p) < e) { }
}

__global__ void fill_f64_kernel(double *p, long long n, double v)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

template <int IPT>
cudaError_t launch_tile(const sblas_seg_args *a, cudaStream_t s)
{
    using Cfg = TileCfg<IPT>;
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(spmv_tile_kernel<IPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::kSmem);
        if (e != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    spmv_tile_kernel<IPT><<<a->ntile, kThreads, Cfg::kSmem, s>>>(*a);
    spmv_tile_fixup<<<(a->ntile + kThreads - 1) / kThreads, kThreads, 0, s>>>(*a, Cfg::kTile);
    return cudaGetLastError();
}

template <int LANES>
cudaError_t launch_vec(const sblas_seg_args *a, cudaStream_t s)
{
    const long long nrows = (long long)a->row_hi - a->row_lo + 1;
    const long long blocks = (nrows * LANES + kThreads - 1) / kThreads;
    spmv_vec_kernel<LANES><<<(unsigned)blocks, kThreads, 0, s>>>(*a);
    return cudaGetLastError();
}

template <int EPL>
cudaError_t launch_vecp(const sblas_seg_args *a, cudaStream_t s)
{
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && sms[dev] == 0) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    const long long nrows = (long long)a->row_hi - a->row_lo + 1;
    long long blocks = (nrows * 32 + kThreads - 1) / kThreads;
    const long long cap = (long long)(dev < 64 && sms[dev] ? sms[dev] : 148) * 3;      /* persistent: 3 CTAs per SM */
    if (blocks > cap) blocks = cap;
    spmv_vecp_kernel<EPL><<<(unsigned)blocks, kThreads, 0, s>>>(*a);
    return cudaGetLastError();
}

}  // namespace

extern "C" int sblas_tile_size(int ipt) { return kThreads * ipt; }

extern "C" cudaError_t sblas_launch_rebase_rowptr(const long long *rp64, long long first_idx, int total_nnz,
                                                  long long count, int *out, cudaStream_t s)
{
    if (count <= 0) return cudaSuccess;
    const long long blocks = (count + 255) / 256;
    rebase_rowptr_kernel<<<(unsigned)blocks, 256, 0, s>>>(rp64, first_idx, total_nnz, count, out);
    return cudaGetLastError();
}

extern "C" cudaError_t sblas_launch_tile_rows(const sblas_seg_args *a, int tile, int *tstart_out, cudaStream_t s)
{
    const int n = a->ntile + 1;
    tile_rows_kernel<<<(n + 255) / 256, 256, 0, s>>>(*a, tile, tstart_out);
    return cudaGetLastError();
}

extern "C" cudaError_t sblas_launch_tile_meta(const sblas_seg_args *a, int tile, int *tmeta_out, cudaStream_t s)
{
    if (a->ntile <= 0) return cudaSuccess;
    tile_meta_kernel<<<(a->ntile + 255) / 256, 256, 0, s>>>(*a, tile, reinterpret_cast<int4 *>(tmeta_out));
    return cudaGetLastError();
}

extern "C" int sblas_tile_size_kind(int kind, int ipt)
{
    return kind == SBLAS_K_TMA ? sblas_tma_tile_size() : kThreads * ipt;
}

extern "C" cudaError_t sblas_launch_row_block_stats(const int *rowptr, int row_lo, int nrows, int rb, int nz0, int nz1,
                                                    int *out_max, int *out_ptr, cudaStream_t s)
{
    const int nb = (nrows + rb - 1) / rb;
    if (nb <= 0) return cudaSuccess;
    row_block_stats_kernel<<<nb, kThreads, 0, s>>>(rowptr, row_lo, nrows, rb, nz0, nz1, out_max, out_ptr);
    return cudaGetLastError();
}

extern "C" cudaError_t sblas_launch_col_range(const int *col, long long count, int *out_min_max, cudaStream_t s)
{
    if (count <= 0) return cudaSuccess;
    long long blocks = (count + 4095) / 4096;
    if (blocks > 148 * 8) blocks = 148 * 8;
    col_range_kernel<<<(unsigned)blocks, 256, 0, s>>>(col, count, out_min_max);
    return cudaGetLastError();
}

extern "C" cudaError_t sblas_launch_edge_merge(const int *mrow, const int *mbeg, const double *const *msrc,
                                               int nmerge, double *y, double alpha, double beta, cudaStream_t s)
{
    if (nmerge <= 0) return cudaSuccess;
    edge_merge_kernel<<<(nmerge + 127) / 128, 128, 0, s>>>(mrow, mbeg, msrc, nmerge, y, alpha, beta);
    return cudaGetLastError();
}

extern "C" cudaError_t sblas_launch_edge_publish(const double *local_edge, int nlocal, const int *out_slot,
                                                 const int *out_owner, const long long *out_off, int nout,
                                                 const int *owners, int nowners, void *const *peer_bases,
                                                 long long table_words, int world, int my_rank,
                                                 unsigned long long *epoch_ctr, cudaStream_t s)
{
    edge_publish_kernel<<<1, 32, 0, s>>>(local_edge, nlocal, out_slot, out_owner, out_off, nout, owners, nowners,
                                         peer_bases, table_words, world, my_rank, epoch_ctr);
    return cudaGetLastError();
}

extern "C" cudaError_t sblas_launch_edge_merge_wait(const int *mrow, const int *mbeg, const long long *msrc_off,
                                                    int nmerge, double *y, double alpha, double beta,
                                                    const int *contrib, int ncontrib, void *const *peer_bases,
                                                    long long table_words, int world, int my_rank,
                                                    const unsigned long long *epoch_ctr, cudaStream_t s)
{
    if (nmerge <= 0 && ncontrib <= 0) return cudaSuccess;
    edge_merge_wait_kernel<<<1, 128, 0, s>>>(mrow, mbeg, msrc_off, nmerge, y, alpha, beta, contrib, ncontrib,
                                             peer_bases, table_words, world, my_rank, epoch_ctr);
    return cudaGetLastError();
}

/* The flag protocols above have kernels that SPIN until another kernel has run.  With CUDA's lazy module loading
 * the FIRST launch of a kernel may synchronise the device: launching, say, chain_copy_kernel for the first time while
 * chain_ready_kernel spins on the same device (several rank plans sharing one GPU, as in the tests) would block the
 * host behind a kernel that waits for a kernel the host has not launched yet.  Loading them up front (at bind time)
 * removes that window. */
extern "C" cudaError_t sblas_preload_exchange_kernels(void)
{
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, edge_publish_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, edge_merge_wait_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, chain_ready_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, chain_copy_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, chain_done_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, edge_merge_kernel);
    return e;
}

extern "C" cudaError_t sblas_launch_chain_gather(const double *y_src, long long count, long long dst_off,
                                                 void *const *peer_x, void *const *peer_flags, int world, int my_rank,
                                                 unsigned long long *chain_ctr, cudaStream_t s)
{
    chain_ready_kernel<<<1, 32, 0, s>>>(peer_flags, world, my_rank, chain_ctr);
    if (count > 0) {
        long long blocks = (count + 255) / 256;
        if (blocks > 148 * 4) blocks = 148 * 4;
        chain_copy_kernel<<<(unsigned)blocks, 256, 0, s>>>(y_src, count, dst_off, peer_x, world);
    }
    chain_done_kernel<<<1, 32, 0, s>>>(peer_flags, world, my_rank, chain_ctr);
    return cudaGetLastError();
}

extern "C" cudaError_t sblas_launch_fill_f64(double *p, long long n, double v, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    fill_f64_kernel<<<(unsigned)blocks, 256, 0, s>>>(p, n, v);
    return cudaGetLastError();
}

extern "C" cudaError_t sblas_launch_spmv_segment(const sblas_seg_args *a, int kind, int ipt, int lanes,
                                                 cudaStream_t s)
{
    const long long nrows = (long long)a->row_hi - a->row_lo + 1;
    if (nrows <= 0) return cudaSuccess;
    const long long nnz = (long long)a->nz1 - a->nz0;
    if (kind == SBLAS_K_VECP) return ipt == 4 ? launch_vecp<4>(a, s) : launch_vecp<8>(a, s);
    if (kind == SBLAS_K_ROWTILE) return sblas_launch_rowtile(a, ipt & 0xff, ipt >> 8, s);
    if (kind == SBLAS_K_ROWSPLIT) return sblas_launch_rowsplit(a, ipt, s);
    if (kind == SBLAS_K_SHORT) {
        spmv_short_kernel<<<(unsigned)((nrows + kThreads * kShortRows - 1) / (kThreads * kShortRows)), kThreads, 0, s>>>(*a);
        return cudaGetLastError();
    }
    if (kind == SBLAS_K_TMA && nnz > 0 && a->ntile > 0) {
        cudaError_t e = sblas_launch_tma(a, s);
        if (e != cudaSuccess) return e;
        spmv_tile_fixup<<<(a->ntile + kThreads - 1) / kThreads, kThreads, 0, s>>>(*a, sblas_tma_tile_size());
        return cudaGetLastError();
    }
    if (kind == SBLAS_K_TILE && nnz > 0 && a->ntile > 0) {
        switch (ipt) {
        case 4: return launch_tile<4>(a, s);
        case 8: return launch_tile<8>(a, s);
        case 16: return launch_tile<16>(a, s);
        default: return cudaErrorInvalidValue;
        }
    }
    if (lanes <= 0) {
        const long long mean = nnz / nrows;
        lanes = 2;
        while (lanes < 32 && lanes * 2 <= mean) lanes <<= 1;
    }
    switch (lanes) {
    case 2: return launch_vec<2>(a, s);
    case 4: return launch_vec<4>(a, s);
    case 8: return launch_vec<8>(a, s);
    case 16: return launch_vec<16>(a, s);
    case 32: return launch_vec<32>(a, s);
    default: return cudaErrorInvalidValue;
    }
}
