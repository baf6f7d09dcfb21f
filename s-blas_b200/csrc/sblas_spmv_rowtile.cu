/* sblas_spmv_rowtile.cu -- SpMV for panels of MEDIUM rows (every row of the panel holds at most
 * 256 entries, most of them at least 32): warp per R whole rows, TMA-fed.
 *
 * Replaces cusparseDcsrmv / cusparseDcsrmv_mp (spmv/src/dspmv_mgpu_v1.cu:200,206,
 * dspmv_mgpu_v2.cu:351,357, dspmv_mgpu_baseline.cu:163) for the row panels the plan bins as
 * "medium" (sblas_plan.c); the north star's "warp-per-row" bin.
 *
 * The nnz-balanced tile kernel (sblas_spmv_tma.cu) pays for rows that cross chunk and tile
 * borders: flags or piece bookkeeping, a block barrier per tile, a fix-up pass.  When every row
 * fits a warp's 256-entry window none of that is needed: a tile is 8*R consecutive WHOLE rows
 * (R = floor(256 / longest row of the panel), 1..8), warp w owns rows [w*R, (w+1)*R) of it, and
 * a row never leaves its warp.
 *
 *   - grid = 2 CTAs per SM, persistent; tile j goes to CTA j mod grid
 *   - one producer warp: a single lane issues 1-D bulk copies (cp.async.bulk, SASS UBLKCP) of the
 *     tile's val and col range (from the 16-byte-aligned entry at or before its first entry) and
 *     of its 8R+1 row pointers into a 3-stage shared-memory ring; completion on mbarriers, L2
 *     evict-first; the tile's entry range is looked up one round ahead
 *   - eight consumer warps; slot i of lane l = entry 32*i + l of the warp's rows (stride-1 across
 *     lanes: neighbouring columns coalesce in the x gather).  Per tile and warp:
 *       (1) products of the current tile: val (shared memory) times the x values gathered one
 *           tile earlier; the stage goes back to the producer
 *       (2) col reads + x gathers of the NEXT tile into the same registers
 *       (3) R == 1: tree + one warp reduction, lane 0 writes the row
 *           R  > 1: lanes add to the running row and park their partial sum where the next row
 *                   starts; one transposed pass finishes all rows (lanes 4c..4c+3 sum row c)
 *   - no block barrier, no carry between tiles, no fix-up kernel: every row is written exactly
 *     once by the warp that owns it (deterministic, no atomics).
 */
#include <cuda_runtime.h>
#include "sblas_dev_common.cuh"

namespace {

using namespace sblas;

constexpr int kWin = 256;                      /* entries a warp handles per tile */
constexpr int kRtWarps = 8;                    /* consumer warps */
constexpr int kRtThreads = kRtWarps * 32 + 32; /* + one producer warp */
constexpr int kRtCap = kRtWarps * kWin + 8;    /* entries staged per tile (+ alignment slack) */
constexpr int kRtRp = 80;                      /* row pointers staged per tile: 8*8+1, + alignment slack */
constexpr int kRtStages = 3;

struct __align__(128) RtStage {
    double val[kRtCap];
    int col[kRtCap];
    int rp[kRtRp];
    int s0;          /* GPU-local index of val[0] / col[0] */
    int rp_off;      /* rp[rp_off + q] == rowptr[first row of the tile + q] */
    int nrows;       /* rows of the tile (8R except in the last tile) */
    int pad;
};

constexpr int kRtBarBytes = 2 * kRtStages * 8;
constexpr int kRtScratch = 8 * 32;             /* doubles per warp: P[row][lane] */
constexpr int kRtSmem = kRtStages * (int)sizeof(RtStage) + kRtBarBytes + kRtWarps * kRtScratch * 8;

/* The producer lane of both kernels: tile j = rows [row_lo + j*rows_per_tile, +rows_per_tile); one bulk
 * copy each for its val range, col range (from the 16-byte-aligned entry at or before the first
 * entry) and row pointers into the stage ring; the tile's entry range is looked up a round ahead. */
__device__ __forceinline__ void rt_produce(const sblas_seg_args &a, RtStage *st, uint32_t smem0, uint32_t full0,
                                           uint32_t empty0, int rows_per_tile, int ntile, int cta, int ncta)
{
    const uint64_t pol = policy_evict_first();
    int j = cta;
    /* entry range of tile j: [rowptr[r0], rowptr[r1]) clamped to the segment */
    auto bounds = [&](int jj, int &r0, int &r1, int &e0, int &e1) {
        r0 = a.row_lo + jj * rows_per_tile;
        r1 = min(r0 + rows_per_tile, a.row_hi + 1);
        e0 = __ldg(a.rowptr + r0);
        e1 = __ldg(a.rowptr + r1);
    };
    int r0n = 0, r1n = 0, e0n = 0, e1n = 0;
    if (j < ntile) bounds(j, r0n, r1n, e0n, e1n);
    int s = 0;
    uint32_t ph = 0;
    for (; j < ntile; j += ncta) {
        const int r0 = r0n, r1 = r1n;
        const int e0 = min(max(e0n, a.nz0), a.nz1), e1 = min(max(e1n, a.nz0), a.nz1);
        if (j + ncta < ntile) bounds(j + ncta, r0n, r1n, e0n, e1n);
        mbar_wait(empty0 + 8u * s, ph ^ 1u);
        const int s0 = e0 & ~3;                                  /* 32-byte aligned in val, 16 in col */
        const int cnt = min(e1, a.nz_total) - s0;                /* <= 2048 + 3 */
        const uint32_t vb = cnt > 0 ? (((uint32_t)cnt * 8u + 15u) & ~15u) : 0u;
        const uint32_t cb = cnt > 0 ? (((uint32_t)cnt * 4u + 15u) & ~15u) : 0u;
        const int rp0 = r0 & ~3;
        const uint32_t rb = (uint32_t)(((r1 - rp0 + 1) + 3) & ~3) * 4u;
        st[s].s0 = s0;
        st[s].rp_off = r0 - rp0;
        st[s].nrows = r1 - r0;
        const uint32_t sbase = smem0 + (uint32_t)(s * sizeof(RtStage));
        const uint32_t fb = full0 + 8u * s;
        mbar_arrive_expect_tx(fb, vb + cb + rb);
        if (cnt > 0) {
            bulk_g2s(sbase, a.val + s0, vb, fb, pol);
            bulk_g2s(sbase + kRtCap * 8, a.col + s0, cb, fb, pol);
        }
        bulk_g2s(sbase + kRtCap * 12, a.rowptr + rp0, rb, fb, pol);
        if (++s == kRtStages) { s = 0; ph ^= 1u; }
    }
}

/* NS = slots per lane: the panel's R rows hold at most 32*NS entries.  RT = R when it is a power of two (1, 2, 4, 8:
 * rows of 129..256, 65..128, 33..64, <= 32 entries), else 0 = R from the argument.  With RT the warp is cut into RT
 * LANE GROUPS of L = 32/RT lanes, one row each: slot i of lane l is entry (l % L) + L*i of row l / L, every group
 * reads its two row boundaries with broadcast loads, and ONE butterfly over log2(L) levels finishes all RT rows at
 * once -- no row-split predicates, no scratch, no transposed pass (round 2; the general RT = 0 path keeps them). */
template <int NS, int RT>
__global__ void __launch_bounds__(kRtThreads, 2) spmv_rowtile_kernel(const sblas_seg_args a, const int Rarg)
{
    const int R = RT ? RT : Rarg;
    extern __shared__ __align__(128) unsigned char smem[];
    RtStage *st = reinterpret_cast<RtStage *>(smem);
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t full0 = smem0 + (uint32_t)(kRtStages * sizeof(RtStage));
    const uint32_t empty0 = full0 + 8u * kRtStages;
    double *scratch = reinterpret_cast<double *>(smem + kRtStages * sizeof(RtStage) + kRtBarBytes);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ncta = gridDim.x, cta = blockIdx.x;
    const int rows_per_tile = kRtWarps * R;
    const int nrows_all = a.row_hi - a.row_lo + 1;
    const int ntile = (nrows_all + rows_per_tile - 1) / rows_per_tile;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kRtStages; ++s) {
            mbar_init(full0 + 8u * s, 1);
            mbar_init(empty0 + 8u * s, kRtWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kRtWarps) {
        /* ------------------------------------------------------------ producer */
        if (lane == 0) rt_produce(a, st, smem0, full0, empty0, rows_per_tile, ntile, cta, ncta);
        return;
    }

    /* ---------------------------------------------------------------- consumers */
    int j = cta;
    if (j >= ntile) return;
    const double *__restrict__ xp = a.x;
    const bool has_y = a.beta != 0.0;
    double *P = scratch + warp * kRtScratch;

    /* xv: the x values of the tile being multiplied; xvn: those of the NEXT tile, whose gathers are issued at the top
     * of the iteration -- a whole tile's products and row sums ahead of their use (with the gathers issued only after
     * the products, 19 % of the stall samples sat on the multiply waiting for x: profiles/r01_ncu_hotspots_*rowtile*) */
    double xv[NS], xvn[NS];
    int cb = 0, ce = 0;       /* my rows' entries, stage-local [cb, ce) */
    int first = 0, nr = 0;    /* my rows: first (GPU-local row id) and how many */
    int v = kWin;             /* lane q < nr-1: where row first+q+1 starts, relative to cb; else kWin */
    double yv = 0.0;          /* lane 4c, c < nr: y of row first+c (loaded a whole tile ahead of its use) */

    auto gather = [&](const RtStage &S, int jj) {
        const int r0 = a.row_lo + jj * rows_per_tile;
        const int mine0 = min(warp * R, S.nrows);
        nr = min(R, S.nrows - mine0);
        first = r0 + mine0;
        const int *rp = S.rp + S.rp_off + mine0;
        if (RT > 0) {
            /* lane groups: my group's row, both boundaries with (group-)broadcast loads, no shuffles */
            constexpr int L = 32 / (RT > 0 ? RT : 1);
            const int g = lane / L;
            const bool has = g < nr;
            cb = has ? min(max(rp[g], a.nz0), a.nz1) - S.s0 : 0;
            ce = has ? min(max(rp[g + 1], a.nz0), a.nz1) - S.s0 : cb;
            const int row = first + g;
            yv = 0.0;
            if (has_y && has && (lane % L) == 0 && row != a.skip_first && row != a.skip_last) yv = a.y[row];
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                const int pos = cb + L * i + (lane % L);
                xvn[i] = 0.0;
                if (pos < ce) xvn[i] = __ldg(xp + (unsigned)S.col[pos]);
            }
            return;
        }
        {
            /* lane q <= nr reads the q-th boundary of my rows */
            int b = 0;
            if (lane <= nr) b = min(max(rp[lane], a.nz0), a.nz1) - S.s0;
            cb = __shfl_sync(kFull, b, 0);
            ce = __shfl_sync(kFull, b, nr);
            const int nxt = __shfl_down_sync(kFull, b, 1);
            v = (lane + 1 < nr) ? nxt - cb : kWin;
            const int row = first + (lane >> 2);
            yv = 0.0;
            if (has_y && (lane & 3) == 0 && (lane >> 2) < nr && row != a.skip_first && row != a.skip_last) yv = a.y[row];
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int pos = cb + 32 * i + lane;
            xvn[i] = 0.0;
            if (pos < ce) xvn[i] = __ldg(xp + (unsigned)S.col[pos]);
        }
    };

    int s = 0;
    uint32_t ph = 0;
    mbar_wait(full0, 0u);
    gather(st[0], j);

    for (; j < ntile; j += ncta) {
        RtStage &S = st[s];
        int sn = s + 1;
        uint32_t phn = ph;
        if (sn == kRtStages) { sn = 0; phn ^= 1u; }
        const bool has_next = j + ncta < ntile;
        const int ccb = cb, cce = ce, cfirst = first, cnr = nr, cv = v;
        const double cyv = yv;
#pragma unroll
        for (int i = 0; i < NS; ++i) xv[i] = xvn[i];

        /* (0) the next tile's col reads + x gathers go out first */
        if (has_next) {
            mbar_wait(full0 + 8u * sn, phn);
            gather(st[sn], j + ncta);
        }

        /* (1) products; slots past the end of my rows are zero */
        constexpr int L = RT > 0 ? 32 / (RT > 0 ? RT : 1) : 32;      /* lanes per row (RT > 0) / stride of the slots */
        const int q0 = RT > 0 ? (lane % L) : lane;
        double p[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int pos = ccb + L * i + q0;
            p[i] = (pos < cce) ? S.val[pos] * xv[i] : 0.0;
        }
        double psum = p[0], psum1 = p[1];                  /* every product: the witness of the release, and */
#pragma unroll
        for (int i = 2; i < NS; i += 2) psum += p[i];      /* my lane's share of its row's sum when RT > 0   */
#pragma unroll
        for (int i = 3; i < NS; i += 2) psum1 += p[i];
        psum += psum1;
        release_after(empty0 + 8u * s, lane, psum);        /* the stage goes back once its values are in registers */

        /* (3) row sums */
        double mine;
        int myrow;
        bool owner;
        if (RT > 0) {
            /* one butterfly inside every lane group finishes all RT rows at once */
            mine = psum;
#pragma unroll
            for (int off = L >> 1; off > 0; off >>= 1) mine += __shfl_xor_sync(kFull, mine, off);
            myrow = cfirst + lane / L;
            owner = q0 == 0 && lane / L < cnr;
        } else {
            /* lane 4c owns row cfirst + c */
            const int pc = lane >> 2;
            myrow = cfirst + pc;
            owner = (lane & 3) == 0 && pc < cnr;
            double acc = 0.0;
            int cur = 0;
            int nb = __shfl_sync(kFull, cv, 0);                  /* next row start relative to ccb, 256 = none */
            const unsigned hasb = __reduce_or_sync(kFull, (lane + 1 < cnr && cv < kWin) ? 1u << (cv >> 5) : 0u);
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                if ((hasb & (1u << i)) == 0) { acc += p[i]; continue; }
                int lo_lane = 0;
                while (nb < 32 * (i + 1)) {                      /* a row starts inside slice i */
                    const int o = nb - 32 * i;
                    if (lane >= lo_lane && lane < o) acc += p[i];
                    P[cur * 32 + lane] = acc;
                    acc = 0.0;
                    lo_lane = o;
                    ++cur;
                    nb = __shfl_sync(kFull, cv, cur);
                }
                if (lane >= lo_lane) acc += p[i];
            }
            P[cur * 32 + lane] = acc;                            /* my last row with entries */
            for (++cur; cur < cnr; ++cur) P[cur * 32 + lane] = 0.0;   /* empty rows at the end of my window */
            __syncwarp();
            /* lane (c, q) takes elements 16h + 4q + ((c + e) & 3), e = 0..7, h = e >> 2: every load is
             * bank-conflict-free and the four lanes of a row cover its 32 partial sums */
            const double *Q = P + pc * 32 + 4 * (lane & 3);
            double t[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) t[e] = Q[16 * (e >> 2) + ((pc + e) & 3)];
            mine = ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
            mine += __shfl_xor_sync(kFull, mine, 1);
            mine += __shfl_xor_sync(kFull, mine, 2);             /* rows beyond cnr: unused garbage */
            __syncwarp();                                        /* P is rewritten by the next tile */
        }
        if (owner) {
            if (myrow == a.skip_first) a.edge[0] = mine;
            else if (myrow == a.skip_last) a.edge[1] = mine;
            else a.y[myrow] = a.alpha * mine + a.beta * cyv;
        }
        s = sn; ph = phn;
    }
}

/* On by default since round 2 (SBLAS_MEDIUM bit 1; SBLAS_MEDIUM=1 turns these panels off).
 * Rows of 257 .. 2048 entries: G = 2, 4 or 8 warps per row (every row of the panel holds at most 256*G
 * entries), a tile is 8/G whole rows.  Warp w takes part w % G of row w / G (the row cut into G equal
 * pieces of at most 256 entries), reduces it like the R == 1 case above and posts one partial sum; after
 * the tile's barrier lane l of warp 0 finishes row l of the tile from its G partials in part order
 * (row-aligned: no carry, no fix-up).  y of a tile's rows is loaded a tile ahead by warp 0.
 * 7.1 TB/s on 1.05 M rows of 1,000 entries (general kernel: 5.8).  Its first version returned one wrong
 * part in ~1e-3 of the long rows of a scattered-column test matrix: the stage was handed back while a
 * shared-memory load of it was still in flight (see release_after in sblas_dev_common.cuh, which fixed
 * it here and closed the same window in the other kernels). */
constexpr int kSplitRing = 4;
__global__ void __launch_bounds__(kRtThreads, 2) spmv_rowsplit_kernel(const sblas_seg_args a, const int G)
{
    extern __shared__ __align__(128) unsigned char smem[];
    RtStage *st = reinterpret_cast<RtStage *>(smem);
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t full0 = smem0 + (uint32_t)(kRtStages * sizeof(RtStage));
    const uint32_t empty0 = full0 + 8u * kRtStages;
    double *red = reinterpret_cast<double *>(smem + kRtStages * sizeof(RtStage) + kRtBarBytes);   /* [ring][8] */

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ncta = gridDim.x, cta = blockIdx.x;
    const int rows_per_tile = kRtWarps / G;
    const int nrows_all = a.row_hi - a.row_lo + 1;
    const int ntile = (nrows_all + rows_per_tile - 1) / rows_per_tile;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kRtStages; ++s) {
            mbar_init(full0 + 8u * s, 1);
            mbar_init(empty0 + 8u * s, kRtWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kRtWarps) {
        if (lane == 0) rt_produce(a, st, smem0, full0, empty0, rows_per_tile, ntile, cta, ncta);
        return;
    }

    int j = cta;
    if (j >= ntile) return;
    const double *__restrict__ xp = a.x;
    const bool has_y = a.beta != 0.0;
    const int rr = warp / G, part = warp - rr * G;      /* my row of the tile, my piece of that row */

    constexpr int NS = kWin / 32;
    double xv[NS], xvn[NS];   /* current tile's x values / the next tile's, gathered a whole tile ahead (see spmv_rowtile_kernel) */
    int cb = 0, ce = 0;       /* my piece, stage-local [cb, ce) */
    int trows = 0;            /* rows of the tile */
    double yv = 0.0;          /* warp 0, lane l < trows: y of the tile's row l */

    auto gather = [&](const RtStage &S, int jj) {
        trows = S.nrows;
        cb = 0; ce = 0;
        if (rr < trows) {
            const int *rp = S.rp + S.rp_off + rr;
            const int b = min(max(rp[0], a.nz0), a.nz1) - S.s0;
            const int e = min(max(rp[1], a.nz0), a.nz1) - S.s0;
            const int ps = (e - b + G - 1) / G;                       /* <= 256 */
            cb = min(b + part * ps, e);
            ce = min(cb + ps, e);
        }
        if (warp == 0) {
            const int row = a.row_lo + jj * rows_per_tile + lane;
            yv = 0.0;
            if (has_y && lane < trows && row != a.skip_first && row != a.skip_last) yv = a.y[row];
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int pos = cb + 32 * i + lane;
            xvn[i] = 0.0;
            if (pos < ce) xvn[i] = __ldg(xp + (unsigned)S.col[pos]);
        }
    };

    int s = 0;
    uint32_t ph = 0;
    unsigned it = 0;
    mbar_wait(full0, 0u);
    gather(st[0], j);

    for (; j < ntile; j += ncta) {
        RtStage &S = st[s];
        int sn = s + 1;
        uint32_t phn = ph;
        if (sn == kRtStages) { sn = 0; phn ^= 1u; }
        const bool has_next = j + ncta < ntile;
        const int ccb = cb, cce = ce, ctrows = trows;
        const double cyv = yv;
        const int ring = (int)(it & (kSplitRing - 1));
#pragma unroll
        for (int i = 0; i < NS; ++i) xv[i] = xvn[i];
        if (has_next) {                                           /* the next tile's gathers go out first */
            mbar_wait(full0 + 8u * sn, phn);
            gather(st[sn], j + ncta);
        }

        double t0 = 0.0, t1 = 0.0;
#pragma unroll
        for (int i = 0; i < NS; i += 2) {
            const int pos = ccb + 32 * i + lane;
            if (pos < cce) t0 = fma(S.val[pos], xv[i], t0);
            if (pos + 32 < cce) t1 = fma(S.val[pos + 32], xv[i + 1], t1);
        }
        release_after(empty0 + 8u * s, lane, t0 + t1);              /* the stage goes back once its values are in registers */

        const double mine = warp_sum(t0 + t1);
        double *R = red + ring * kRtWarps;
        if (lane == 0) R[warp] = mine;
        named_bar_sync(1 + ring, kRtWarps * 32);            /* every warp waits: lock-step, as in path W */
        if (warp == 0) {
            if (lane < ctrows) {
                double tot = 0.0;
                for (int q = 0; q < G; ++q) tot += R[lane * G + q];
                const int row = a.row_lo + j * rows_per_tile + lane;
                if (row == a.skip_first) a.edge[0] = tot;
                else if (row == a.skip_last) a.edge[1] = tot;
                else a.y[row] = a.alpha * tot + a.beta * cyv;
            }
        }
        s = sn; ph = phn; ++it;
    }
}

int g_rt_sm_count[64] = {0};

}  // namespace

/* y[rows] = alpha*A*x + beta*y for a panel whose rows all hold at most 256/R entries;
 * window = the most entries R consecutive rows of the panel can hold (<= 256) */
cudaError_t sblas_launch_rowtile(const sblas_seg_args *a, int R, int window, cudaStream_t s)
{
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64 || R < 1 || R > 8) return cudaErrorInvalidValue;
    if (!attr_done[dev]) {
        cudaError_t e = cudaSuccess;
#define RT_ATTR1(NS, RT) if (e == cudaSuccess) e = cudaFuncSetAttribute(spmv_rowtile_kernel<NS, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRtSmem)
#define RT_ATTR(NS) RT_ATTR1(NS, 0); RT_ATTR1(NS, 1); RT_ATTR1(NS, 2); RT_ATTR1(NS, 4); RT_ATTR1(NS, 8)
        RT_ATTR(4); RT_ATTR(5); RT_ATTR(6); RT_ATTR(7); RT_ATTR(8);
#undef RT_ATTR
#undef RT_ATTR1
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&g_rt_sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    const int nrows = a->row_hi - a->row_lo + 1;
    if (nrows <= 0) return cudaSuccess;
    const int ntile = (nrows + 8 * R - 1) / (8 * R);
    int grid = 2 * g_rt_sm_count[dev];
    if (grid > ntile) grid = ntile;
    int ns = (window + 31) / 32;
    if (window <= 0 || ns > 8) ns = 8;
#define RT_LAUNCH(NS) do { if (R == 1) spmv_rowtile_kernel<NS, 1><<<grid, kRtThreads, kRtSmem, s>>>(*a, R); \
                           else if (R == 2) spmv_rowtile_kernel<NS, 2><<<grid, kRtThreads, kRtSmem, s>>>(*a, R); \
                           else if (R == 4) spmv_rowtile_kernel<NS, 4><<<grid, kRtThreads, kRtSmem, s>>>(*a, R); \
                           else if (R == 8) spmv_rowtile_kernel<NS, 8><<<grid, kRtThreads, kRtSmem, s>>>(*a, R); \
                           else spmv_rowtile_kernel<NS, 0><<<grid, kRtThreads, kRtSmem, s>>>(*a, R); } while (0)
    switch (ns) {
    case 1: case 2: case 3: case 4: RT_LAUNCH(4); break;
    case 5: RT_LAUNCH(5); break;
    case 6: RT_LAUNCH(6); break;
    case 7: RT_LAUNCH(7); break;
    default: RT_LAUNCH(8); break;
    }
#undef RT_LAUNCH
    return cudaGetLastError();
}

/* y[rows] = alpha*A*x + beta*y for a panel whose rows all hold at most 256*G entries, G = 2, 4 or 8 */
cudaError_t sblas_launch_rowsplit(const sblas_seg_args *a, int G, cudaStream_t s)
{
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64 || (G != 2 && G != 4 && G != 8)) return cudaErrorInvalidValue;
    if (!attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(spmv_rowsplit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRtSmem);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&g_rt_sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    const int nrows = a->row_hi - a->row_lo + 1;
    if (nrows <= 0) return cudaSuccess;
    const int rpt = 8 / G;
    const int ntile = (nrows + rpt - 1) / rpt;
    int grid = 2 * g_rt_sm_count[dev];
    if (grid > ntile) grid = ntile;
    spmv_rowsplit_kernel<<<grid, kRtThreads, kRtSmem, s>>>(*a, G);
    return cudaGetLastError();
}
