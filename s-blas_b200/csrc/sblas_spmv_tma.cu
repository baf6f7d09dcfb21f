/* sblas_spmv_tma.cu -- the GENERAL SpMV kernel: persistent, warp-specialised, TMA-fed,
 * nnz-balanced tiles; whatever the rows look like.
 *
 * Replaces cusparseDcsrmv_mp / cusparseDcsrmv (spmv/src/dspmv_mgpu_v1.cu:200,206,
 * dspmv_mgpu_v2.cu:351,357, dspmv_mgpu_baseline.cu:163) and the CSR5 tile kernels
 * (spmv/include/detail/cuda/csr5_spmv_cuda.h:275-311).  The plan sends row panels that are
 * uniformly short or medium to leaner kernels (spmv_short_kernel, sblas_spmv_rowtile.cu);
 * everything else -- long rows, mixed and power-law rows -- runs here.
 *
 * SpMV is HBM-bound (12 B streamed per nnz, 2 flop): the kernel is built around keeping
 * as many bytes in flight per SM as possible with no register or LSU cost, and around a
 * short instruction stream per nnz (the part runs into its power cap otherwise):
 *
 *   - grid = 2 CTAs per SM, persistent; tile j of 2048 nnz goes to CTA j mod grid
 *     (nnz-balanced: every CTA streams the same number of bytes, whatever the rows are)
 *   - one producer warp: a single lane issues 1-D bulk copies (cp.async.bulk, SASS UBLKCP)
 *     of the tile's val (16 KB), col (8 KB) and -- when the tile holds several rows -- its
 *     slice of the row pointer into a 3-stage shared-memory ring, completion counted on
 *     mbarriers, L2 evict-first so the stream does not push x out of L2; the per-tile
 *     metadata (two int4) is prefetched one round ahead
 *   - eight consumer warps; every warp owns a contiguous 256-entry chunk of the tile, element
 *     i of lane l = chunk entry 32*i + l (stride-1 across lanes, so the x gather of
 *     neighbouring columns coalesces).  Per tile:
 *       (1) multiply val (shared memory) by the x values gathered one iteration earlier,
 *       (2) issue the col reads + x gathers of the NEXT tile into the same registers,
 *       (3) reduce the current tile while those gathers are in flight; the path is picked per
 *           tile from its metadata (adaptive binning at tile granularity):
 *             A  <= 1 row starts in the tile (rows >= 2048): FMA into registers, warp shuffle,
 *                8 partials through shared memory; only warp 0 waits on the named barrier
 *             W  <= 7 row starts per chunk, none empty (rows ~ 32 .. 2048): the row starts cut
 *                a chunk into pieces; lanes add to the running piece and park partial sums
 *                where a row starts, one transposed pass finishes all pieces
 *             M  several short rows, none empty: byte flags + lane walk + one segmented scan
 *             S  tiles with empty rows: products in place, G = 1..32 lanes per row
 *   - rows that leave their tile are finished by spmv_tile_fixup in a fixed order
 *     (deterministic; no floating-point atomics).
 */
#include <cuda_runtime.h>
#include "sblas_dev_common.cuh"

namespace {

using namespace sblas;

constexpr int kTile = 2048;                    /* nnz per tile */
constexpr int kConsumers = 256;                /* consumer threads of one group (one tile at a time) */
constexpr int kCWarps = kConsumers / 32;
constexpr int kGroups = 1;                     /* consumer groups per CTA, each on its own tile (3 groups on a
                                                  7-stage ring, one CTA per SM, measured slower: every group holds
                                                  two stages, which starves the ring) */
#ifndef SBLAS_TMA_CTAS
#define SBLAS_TMA_CTAS 2
#endif
#ifndef SBLAS_TMA_STAGES
#define SBLAS_TMA_STAGES 3
#endif
constexpr int kCtasPerSm = SBLAS_TMA_CTAS;
constexpr int kThreadsTma = kGroups * kConsumers + 32;   /* + one producer warp */
constexpr int kIPT = kTile / kConsumers;       /* 8 products per consumer thread */
constexpr int kRpCap = 1032;                   /* row-pointer ints staged per tile (x4) */
constexpr int kStages = SBLAS_TMA_STAGES;
constexpr int kChunk = kTile / kCWarps;             /* entries owned by one consumer warp */

struct __align__(128) Stage {
    double val[kTile];
    int col[kTile];
    int rp[kRpCap];
    int4 meta;       /* {rs, re, start of row rs clamped to the tile, flags (1: last row leaves, 2: empty rows)} */
    int4 meta2;      /* 8 x uint16: rows of the tile that start before chunk w */
    int rp_off;      /* rp[rp_off + q] == rowptr[rs + q] */
    int rp_ok;       /* the row pointer slice was staged */
};

constexpr int kRing = 4;                       /* partial-sum buffers / named-barrier ids: a warp is at most
                                                  kStages tiles ahead of the slowest one, so 4 slots suffice */
static_assert(kGroups * kRing < 16, "named barrier ids");
static_assert(kStages < kRing, "a warp may run kStages tiles ahead: the partial-sum ring / barrier ids need one more slot");
static_assert((kRing & (kRing - 1)) == 0, "ring index is masked");
constexpr int kRedDoubles = kRing * (2 + 3) * kCWarps;          /* per group */
constexpr int kBarBytes = 2 * kStages * 8;                      /* full[], empty[] mbarriers */
constexpr int kSmemBytes = kStages * (int)sizeof(Stage) + kBarBytes + kGroups * kRedDoubles * 8;

__device__ __forceinline__ void release_stage(uint32_t empty_bar, int lane)
{
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar);
}

/* finish one row in the several-rows path */
__device__ __forceinline__ void emit_seg(const sblas_seg_args &a, int j, int sg, int nown, bool ext, int rs,
                                         double acc, double yv)
{
    const int row = rs + sg - 1;
    if (sg == 0) {
        a.carry[j] = acc;
    } else if (sg == nown && ext) {
        a.tail[j] = acc;
    } else if (row == a.skip_first) {
        a.edge[0] = acc;
    } else if (row == a.skip_last) {
        a.edge[1] = acc;
    } else {
        a.y[row] = a.alpha * acc + a.beta * yv;      /* yv == 0 when beta == 0 */
    }
}

/* Several rows in a tile: products already sit in S.val; G lanes reduce each segment.
 * Segment 0 is the row left open by the previous tile, segment q >= 1 is row rs+q-1;
 * segment sg = [bound(sg-1), bound(sg)), bound(-1) = clo, bound(q) = clamp(rowptr[rs+q]). */
template <int G>
__device__ __forceinline__ void reduce_rows(const sblas_seg_args &a, const Stage &S, int j, int t, int base,
                                            int clo, int chi, int rs, int nown, bool ext, const double *yin,
                                            int ypre)
{
    constexpr int ngroups = kConsumers / G;
    const int grp = t / G, gl = t & (G - 1);
    const int nseg = nown + 1;
    const int T0 = base + clo, T1 = base + chi;
    const bool staged = S.rp_ok != 0;
    const int *rpl = S.rp + S.rp_off;
    const int *rpg = a.rowptr + rs;
    int rr = 0;
    for (int s0 = 0; s0 < nseg; s0 += ngroups, ++rr) {
        const int sg = s0 + grp;
        double acc = 0.0;
        if (sg < nseg) {
            int b = clo;
            if (sg > 0) b = min(max(staged ? rpl[sg - 1] : __ldg(rpg + sg - 1), T0), T1) - base;
            const int e = min(max(staged ? rpl[sg] : __ldg(rpg + sg), T0), T1) - base;
            double acc1 = 0.0;
            int k = b + gl;
            for (; k + G < e; k += 2 * G) { acc += S.val[k]; acc1 += S.val[k + G]; }
            if (k < e) acc += S.val[k];
            acc += acc1;
        }
#pragma unroll
        for (int off = G >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(kFull, acc, off);
        if (sg < nseg && gl == 0) {
            double yv = 0.0;
            if (a.beta != 0.0) {
                if (rr < ypre) {
                    yv = yin[0];
#pragma unroll
                    for (int q = 1; q < 4; ++q) if (rr == q) yv = yin[q];
                } else if (sg >= 1) {
                    yv = a.y[rs + sg - 1];
                }
            }
            emit_seg(a, j, sg, nown, ext, rs, acc, yv);
        }
    }
}

__global__ void __launch_bounds__(kThreadsTma, kCtasPerSm) spmv_tma_kernel(const sblas_seg_args a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    Stage *st = reinterpret_cast<Stage *>(smem);
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t full0 = smem0 + (uint32_t)(kStages * sizeof(Stage));    /* full[s]  = full0 + 8 s  */
    const uint32_t empty0 = full0 + 8u * kStages;                           /* empty[s] = empty0 + 8 s */
    const int tid = threadIdx.x, pw = tid >> 5, lane = tid & 31;
    const int grp = pw / kCWarps, warp = pw % kCWarps;      /* consumer group, warp inside the group */
    double *red = reinterpret_cast<double *>(smem + kStages * sizeof(Stage) + kBarBytes) + grp * kRedDoubles;
    const int ncta = gridDim.x, cta = blockIdx.x;
    const int ntile = a.ntile;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full0 + 8u * s, 1);
            mbar_init(empty0 + 8u * s, kCWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (pw == kGroups * kCWarps) {
        /* ------------------------------------------------------------ producer */
        if (lane == 0) {
            const uint64_t pol = policy_evict_first();
            const int4 *tm = reinterpret_cast<const int4 *>(a.tmeta);
            int j = cta;
            int4 mnext = make_int4(0, 0, 0, 0), m2next = make_int4(0, 0, 0, 0);
            if (j < ntile) { mnext = __ldg(tm + 2 * j); m2next = __ldg(tm + 2 * j + 1); }
            int s = 0;
            uint32_t ph = 0;
            for (; j < ntile; j += ncta) {
                const int4 m = mnext, m2 = m2next;
                if (j + ncta < ntile) { mnext = __ldg(tm + 2 * (j + ncta)); m2next = __ldg(tm + 2 * (j + ncta) + 1); }
                mbar_wait(empty0 + 8u * s, ph ^ 1u);
                const int base = (a.tile0 + j) * kTile;
                const int cnt = min(kTile, a.nz_total - base);
                const uint32_t vb = ((uint32_t)cnt * 8u + 15u) & ~15u;
                const uint32_t cb = ((uint32_t)cnt * 4u + 15u) & ~15u;
                const int rp0 = m.x & ~3;
                const int rpn = ((m.y - rp0 + 1) + 3) & ~3;
                const bool rp_ok = (m.y - m.x >= 2) && (rpn <= kRpCap);
                const uint32_t rb = rp_ok ? (uint32_t)rpn * 4u : 0u;
                st[s].meta = m;
                st[s].meta2 = m2;
                st[s].rp_off = m.x - rp0;
                st[s].rp_ok = rp_ok ? 1 : 0;
                const uint32_t sbase = smem0 + (uint32_t)(s * sizeof(Stage));
                const uint32_t fb = full0 + 8u * s;
                mbar_arrive_expect_tx(fb, vb + cb + rb);
                bulk_g2s(sbase, a.val + base, vb, fb, pol);
                bulk_g2s(sbase + kTile * 8, a.col + base, cb, fb, pol);
                if (rp_ok) bulk_g2s(sbase + kTile * 12, a.rowptr + rp0, rb, fb, pol);
                if (++s == kStages) { s = 0; ph ^= 1u; }
            }
        }
        return;
    }

    /* ---------------------------------------------------------------- consumers
     * Element i of a thread is tile-local index  warp*256 + i*32 + lane : every warp owns a
     * contiguous 256-entry chunk of the tile, lanes are stride-1 inside it. */
    const int t = tid - grp * kConsumers;         /* thread inside the group */
    int j = cta + grp * ncta;                     /* the CTA's tiles cta, cta+ncta, ... go round the groups */
    if (j >= ntile) return;
    const double *__restrict__ xp = a.x;
    const int nz0 = a.nz0, nz1 = a.nz1;
    int base = (a.tile0 + j) * kTile;            /* GPU-local nnz index of the tile (fits int32) */
    const int jstep = kGroups * ncta;
    const int step = jstep * kTile;
    const int c0 = warp * kChunk;                /* my warp's chunk [c0, c0 + 256) */
    const int e0 = c0 + lane;                    /* my element i is e0 + 32 i */

    int4 m;                 /* metadata of the current tile */
    int lo, hi;             /* its valid tile-local range   */
    double xv[kIPT];        /* its gathered x values        */
    int s = grp;            /* the CTA's n-th tile sits in stage n mod kStages */
    uint32_t ph = 0;
    unsigned it = 0;        /* tiles done by this group */

    auto gather = [&](const Stage &S, int b) {
        m = S.meta;
        lo = max(nz0, b) - b;
        hi = min(nz1 - b, kTile);
        unsigned c[kIPT];
#pragma unroll
        for (int i = 0; i < kIPT; ++i) c[i] = (unsigned)S.col[e0 + 32 * i];
        if (lo == 0 && hi == kTile) {
#pragma unroll
            for (int i = 0; i < kIPT; ++i) xv[i] = __ldg(xp + c[i]);
        } else {
#pragma unroll
            for (int i = 0; i < kIPT; ++i) {
                const int e = e0 + 32 * i;
                xv[i] = (e >= lo && e < hi) ? __ldg(xp + c[i]) : 0.0;
            }
        }
    };

    /* path W, after the tile's barrier: lane 4k of a warp finishes the row that left its chunk, warp 0
     * the row left open by the previous tile.  (Settling one tile later on an mbarrier instead of a
     * block barrier was measured slower: warps drift apart and the stage ring turns over unevenly.) */
    auto settle = [&](const double *WC, int p_j, int p_k, bool p_ext, int p_row, double p_mine, double p_yv) {
        if (p_k > 0 && lane == 4 * p_k) {                       /* the row that left my chunk */
            double tot = p_mine;
            bool closed_in_tile = false;
            for (int w = warp + 1; w < kCWarps; ++w) {
                tot += WC[w];
                if (WC[kCWarps + w] != 0.0) { closed_in_tile = true; break; }
            }
            if (!closed_in_tile && !p_ext) closed_in_tile = true;            /* ends exactly at the tile end */
            if (!closed_in_tile) a.tail[p_j] = tot;
            else if (p_row == a.skip_first) a.edge[0] = tot;
            else if (p_row == a.skip_last) a.edge[1] = tot;
            else a.y[p_row] = a.alpha * tot + a.beta * p_yv;
        }
        if (warp == 0 && lane == 0) {
            double tot = 0.0;                                   /* the row left open by the previous tile */
            for (int w = 0; w < kCWarps; ++w) {
                tot += WC[w];
                if (WC[kCWarps + w] != 0.0) break;
            }
            a.carry[p_j] = tot;
        }
    };

    mbar_wait(full0 + 8u * s, 0u);
    gather(st[s], base);

    for (; j < ntile; j += jstep, base += step) {
        Stage &S = st[s];
        int sn = s + kGroups;
        uint32_t phn = ph;
        if (sn >= kStages) { sn -= kStages; phn ^= 1u; }
        const int rs = m.x, nown = m.y - m.x;
        const bool ext = (m.w & 1) != 0;
        const int clo = lo, chi = hi;              /* current tile's range (gather overwrites lo/hi) */
        const int lsplit = m.z - base;
        const int ring = (int)(it & (kRing - 1));
        const int bar_id = 1 + grp * kRing + ring;
        const bool has_next = j + jstep < ntile;
        const bool whole = (clo == 0 && chi == kTile);
        const uint32_t eb = empty0 + 8u * s;
        const int nseg = nown + 1;                 /* segment 0 = the row left open by the previous tile */


        if (nown <= 1) {
            /* ---- A: at most one row starts here: block reduction straight from registers */
            double sc = 0.0, so = 0.0;
            if (!whole) {
                /* partial tile (first / last of a segment): mask explicitly, the slot may hold
                 * other segments' entries or stale data outside [lo,hi) */
#pragma unroll
                for (int i = 0; i < kIPT; ++i) {
                    const int e = e0 + 32 * i;
                    if (e >= clo && e < chi) {
                        const double pr = S.val[e] * xv[i];
                        if (e < lsplit) sc += pr; else so += pr;
                    }
                }
            } else if (nown == 0) {
                double s1 = 0.0;
#pragma unroll
                for (int i = 0; i < kIPT; i += 2) {
                    sc = fma(S.val[e0 + 32 * i], xv[i], sc);
                    s1 = fma(S.val[e0 + 32 * (i + 1)], xv[i + 1], s1);
                }
                sc += s1;
            } else {
#pragma unroll
                for (int i = 0; i < kIPT; ++i) {
                    const int e = e0 + 32 * i;
                    const double pr = S.val[e] * xv[i];
                    if (e < lsplit) sc += pr; else so += pr;
                }
            }
            release_after(eb, lane, sc + so);                  /* slot consumed (values in registers): hand it back */
            if (has_next) {
                mbar_wait(full0 + 8u * sn, phn);
                gather(st[sn], base + step);                   /* in flight during the reduction */
            }
            sc = warp_sum(sc);
            if (nown != 0) so = warp_sum(so);
            double *R = red + ring * (2 * kCWarps);
            if (lane == 0) { R[warp] = sc; R[kCWarps + warp] = so; }
            if (warp != 0) {
                asm volatile("bar.arrive %0, %1;" ::"r"(bar_id), "r"(kConsumers) : "memory");
            } else {
                named_bar_sync(bar_id, kConsumers);
                if (lane == 0) {
                    double c = 0.0, o = 0.0;
#pragma unroll
                    for (int w = 0; w < kCWarps; ++w) { c += R[w]; o += R[kCWarps + w]; }
                    a.carry[j] = c;
                    if (nown == 1) {
                        if (ext) a.tail[j] = o; else emit_row(a, rs, o);
                    }
                }
            }
            s = sn; ph = phn; ++it;
            continue;
        }

        /* ---- products of my chunk (registers) */
        double p[kIPT];
        if (whole) {
#pragma unroll
            for (int i = 0; i < kIPT; ++i) p[i] = S.val[e0 + 32 * i] * xv[i];
        } else {
#pragma unroll
            for (int i = 0; i < kIPT; ++i) {
                const int e = e0 + 32 * i;
                p[i] = (e >= clo && e < chi) ? S.val[e] * xv[i] : 0.0;
            }
        }

        /* ---- W: medium rows (no chunk holds more than 7 row starts, none empty).  The row starts cut a
         * warp's 256-entry chunk into at most 8 pieces.  The products never leave the registers: the
         * warp walks its 8 slices of 32 consecutive entries, every lane adds its product to the
         * running piece and parks its partial sum in the warp's scratch (P[piece][lane]) wherever a
         * row starts; ONE transposed pass then finishes all pieces at once (lanes 4c..4c+3 sum piece
         * c: 8 loads, a 3-level tree, 2 shuffles) -- no flags, no scan, no per-piece shuffle chain.
         * The scratch is the warp's own, already consumed slice of val.  Piece 0 belongs to the row open
         * at the chunk start, the last piece to the row that leaves the chunk; they meet the
         * neighbours' pieces in shared memory after the tile's one barrier (settle()). */
        if ((m.w & 6) == 4 && (a.mode & 1) == 0) {
            const unsigned short *qw = reinterpret_cast<const unsigned short *>(&S.meta2);
            const int qa = qw[warp];
            const int qb = warp == kCWarps - 1 ? nown : (int)qw[warp + 1];
            const int k = qb - qa;                               /* row starts in my chunk (<= 7) */
            /* lane l < k: chunk-local start of row rs+qa+l; other lanes: the chunk end */
            int v = kChunk;
            if (lane < k) {
                const int r = S.rp[S.rp_off + qa + lane];         /* always staged: 2 <= nown <= 56 rows */
                v = min(max(r, base + clo), base + chi) - base - c0;
            }
            double *pieces = S.val + c0;                         /* scratch: my own (consumed) slice of val */
            /* piece c is finished by lanes 4c..4c+3; lane 4c, c in [1,k], owns row rs+qa+c-1 */
            const int pc = lane >> 2;
            const int myrow = rs + qa + pc - 1;
            const bool owner = (lane & 3) == 0 && pc >= 1 && pc <= k;
            double yv = 0.0;
            if (a.beta != 0.0 && owner && myrow != a.skip_first && myrow != a.skip_last) yv = a.y[myrow];
            if (has_next) {
                mbar_wait(full0 + 8u * sn, phn);
                gather(st[sn], base + step);
            }
            double acc = 0.0, mine = 0.0;
            int cur = 0;
            int nb = __shfl_sync(kFull, v, 0);                   /* next row start (chunk-local), 256 = none */
            /* bit i: a row starts inside slice i (one warp-wide OR, the result is warp-uniform) */
            const unsigned hasb = __reduce_or_sync(kFull, lane < k ? 1u << (v >> 5) : 0u);
#pragma unroll
            for (int i = 0; i < kIPT; ++i) {
                if ((hasb & (1u << i)) == 0) { acc += p[i]; continue; }
                int lo_lane = 0;
                while (nb < 32 * (i + 1)) {                      /* a row starts inside slice i */
                    const int o = nb - 32 * i;
                    if (lane >= lo_lane && lane < o) acc += p[i];
                    pieces[cur * 32 + lane] = acc;
                    acc = 0.0;
                    lo_lane = o;
                    ++cur;
                    nb = __shfl_sync(kFull, v, cur);
                }
                if (lane >= lo_lane) acc += p[i];
            }
            pieces[cur * 32 + lane] = acc;                       /* the piece that leaves the chunk: cur == k */
            __syncwarp();
            {
                /* lane (c, q) takes elements 16h + 4q + ((c + e) & 3), e = 0..7, h = e >> 2: every load is
                 * bank-conflict-free and the four lanes of a piece cover its 32 partial sums */
                const double *Q = pieces + pc * 32 + 4 * (lane & 3);
                double t[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) t[e] = Q[16 * (e >> 2) + ((pc + e) & 3)];
                mine = ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
                mine += __shfl_xor_sync(kFull, mine, 1);
                mine += __shfl_xor_sync(kFull, mine, 2);         /* pieces beyond k: unused garbage */
            }
            fence_proxy_async_smem();          /* generic writes to the slot before the next bulk copy */
            release_after(eb, lane, mine);
            if (owner && pc < k) {             /* rows that start and end inside my chunk */
                if (myrow == a.skip_first) a.edge[0] = mine;
                else if (myrow == a.skip_last) a.edge[1] = mine;
                else a.y[myrow] = a.alpha * mine + a.beta * yv;
            }
            double *WC = red + (kRing * 2 * kCWarps) + ring * (3 * kCWarps);   /* [WC | WCend | WT] x 8 */
            if (lane == 0) { WC[warp] = mine; WC[kCWarps + warp] = k > 0 ? 1.0 : 0.0; }
            if (lane == 4 * k) WC[2 * kCWarps + warp] = mine;
            named_bar_sync(bar_id, kConsumers);
            settle(WC, j, k, ext, myrow, mine, yv);
            s = sn; ph = phn; ++it;
            continue;
        }

        /* ---- M: several rows, none empty (the common case).  Merge-style segmented sum inside
         * every warp's own 256-entry chunk, no block-wide product buffer:
         *   1. the warp's products go through its OWN slice of the stage (in place over val, pair-
         *      swizzled) so that lane l ends up with entries 8l..8l+7 of the chunk;
         *   2. row starts inside the chunk are scattered as byte flags (in place over the col slice,
         *      already consumed); the metadata gives the chunk's first/last starting row directly;
         *   3. every lane walks its 8 entries (complete rows inside a lane are written at once),
         *      one segmented scan over the lanes' open sums closes rows that span lanes;
         *   4. pieces of rows that cross chunk borders meet in shared memory (WC = the row open at
         *      the chunk start, WT = the row that leaves the chunk) and are summed in ascending
         *      chunk order after the tile's one barrier. */
        if ((m.w & 2) == 0 && (a.mode & 2) == 0) {
            const unsigned short *qw = reinterpret_cast<const unsigned short *>(&S.meta2);
            const unsigned qa = qw[warp];
            const unsigned qb = warp == kCWarps - 1 ? (unsigned)nown : (unsigned)qw[warp + 1];
            const int T0 = base + clo, T1 = base + chi;
            double *cv = S.val + c0;
            unsigned char *F = reinterpret_cast<unsigned char *>(S.col + c0);
            /* 1. transposed, swizzled store */
#pragma unroll
            for (int i = 0; i < kIPT; ++i) {
                const int L = 4 * i + (lane >> 3), k = (lane >> 1) & 3;
                cv[2 * (4 * L + (k ^ ((L >> 1) & 3))) + (lane & 1)] = p[i];
            }
            /* 2. flags */
            reinterpret_cast<unsigned long long *>(F)[lane] = 0ull;
            __syncwarp();
            {
                const bool staged = S.rp_ok != 0;
                const int *rpl = S.rp + S.rp_off;
                for (unsigned q = qa + lane; q < qb; q += 32) {
                    const int v = staged ? rpl[q] : __ldg(a.rowptr + rs + q);
                    F[min(max(v, T0), T1) - base - c0] = 1;
                }
            }
            /* y of the row that leaves my chunk and of the first rows my lane will close */
            const int wt_row = rs + (int)qb - 1;
            const bool has_y = a.beta != 0.0;
            double ywt = 0.0;
            if (has_y && lane == 0 && qb > qa && wt_row != a.skip_first && wt_row != a.skip_last) ywt = a.y[wt_row];
            __syncwarp();
            double v[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double2 d = *reinterpret_cast<const double2 *>(cv + 2 * (4 * lane + (k ^ ((lane >> 1) & 3))));
                v[2 * k] = d.x; v[2 * k + 1] = d.y;
            }
            const unsigned long long fl = reinterpret_cast<const unsigned long long *>(F)[lane];
            const int cnt = __popcll(fl);
            int excl = cnt;                                     /* rows starting in lanes before mine */
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int tt = __shfl_up_sync(kFull, excl, d);
                if (lane >= d) excl += tt;
            }
            excl -= cnt;
            const int orow = rs + (int)qa - 1 + excl;           /* row open at the start of my 8 entries */
            double yp0 = 0.0, yp1 = 0.0, yp2 = 0.0, yp3 = 0.0;
            if (has_y && cnt > 0) {
                if (orow >= rs && orow != a.skip_first && orow != a.skip_last) yp0 = a.y[orow];
                if (cnt > 1 && orow + 1 != a.skip_first && orow + 1 != a.skip_last) yp1 = a.y[orow + 1];
                if (cnt > 2 && orow + 2 != a.skip_first && orow + 2 != a.skip_last) yp2 = a.y[orow + 2];
                if (cnt > 3 && orow + 3 != a.skip_first && orow + 3 != a.skip_last) yp3 = a.y[orow + 3];
            }
            if (has_next) {
                mbar_wait(full0 + 8u * sn, phn);
                gather(st[sn], base + step);
            }
            /* 3. lane walk */
            double sum = 0.0, head = 0.0;
            bool seen = false;
            int row = orow;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if ((fl >> (8 * k)) & 1ull) {
                    if (!seen) {
                        head = sum; seen = true;
                    } else {                                    /* a whole row inside my 8 entries */
                        double yv = 0.0;
                        if (has_y) yv = (row == orow + 1) ? yp1 : (row == orow + 2) ? yp2 : (row == orow + 3) ? yp3 : a.y[row];
                        if (row == a.skip_first) a.edge[0] = sum;
                        else if (row == a.skip_last) a.edge[1] = sum;
                        else a.y[row] = a.alpha * sum + a.beta * yv;
                    }
                    ++row;
                    sum = 0.0;
                }
                sum += v[k];
            }
            const unsigned bmask = __ballot_sync(kFull, seen);
            const unsigned lower = bmask & ((2u << lane) - 1u);
            const int segs = lower ? 31 - __clz((int)lower) : 0;       /* lane where my open run starts */
            double I = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double tt = __shfl_up_sync(kFull, I, d);
                if (lane - d >= segs) I += tt;
            }
            double cin = __shfl_up_sync(kFull, I, 1);
            if (lane == 0) cin = 0.0;
            const double closed = cin + head;                    /* total of the row open at my start */
            const bool started_here = (bmask & ((1u << lane) - 1u)) != 0u;
            if (seen && started_here) {
                if (orow == a.skip_first) a.edge[0] = closed;
                else if (orow == a.skip_last) a.edge[1] = closed;
                else a.y[orow] = a.alpha * closed + a.beta * yp0;
            }
            /* 4. chunk borders */
            const double I31 = __shfl_sync(kFull, I, 31);
            const int f = bmask ? __ffs((int)bmask) - 1 : 0;
            const double wcv = __shfl_sync(kFull, closed, f);
            double *WC = red + (kRing * 2 * kCWarps) + ring * (3 * kCWarps);   /* [WC | WCend | WT] x 8 */
            if (lane == 0) {
                WC[warp] = bmask ? wcv : I31;
                WC[kCWarps + warp] = bmask ? 1.0 : 0.0;
                WC[2 * kCWarps + warp] = I31;
            }
            named_bar_sync(bar_id, kConsumers);
            if (lane == 0) {
                if (bmask) {                                     /* the row that left my chunk */
                    double tot = I31;
                    bool closed_in_tile = false;
                    for (int w = warp + 1; w < kCWarps; ++w) {
                        tot += WC[w];
                        if (WC[kCWarps + w] != 0.0) { closed_in_tile = true; break; }
                    }
                    if (!closed_in_tile && !ext) closed_in_tile = true;      /* ends exactly at the tile end */
                    if (!closed_in_tile) a.tail[j] = tot;
                    else if (wt_row == a.skip_first) a.edge[0] = tot;
                    else if (wt_row == a.skip_last) a.edge[1] = tot;
                    else a.y[wt_row] = a.alpha * tot + a.beta * ywt;
                }
                if (warp == 0) {
                    double tot = 0.0;                            /* the row left open by the previous tile */
                    for (int w = 0; w < kCWarps; ++w) {
                        tot += WC[w];
                        if (WC[kCWarps + w] != 0.0) break;
                    }
                    a.carry[j] = tot;
                }
            }
            fence_proxy_async_smem();          /* generic writes to the slot before the next bulk copy */
            /* released only now: the WC ring slot of this stage is reused when the stage is */
            release_stage(eb, lane);
            s = sn; ph = phn; ++it;
            continue;
        }

        /* ---- S: many short rows (or empty rows): products to shared memory in place over val,
         * then G lanes per row */
#pragma unroll
        for (int i = 0; i < kIPT; ++i) S.val[e0 + 32 * i] = p[i];
        const int len = chi - clo;                 /* mean segment length picks the lanes per row */
        const int lg = len >= 128 * nseg ? 5 : len >= 64 * nseg ? 4 : len >= 32 * nseg ? 3
                     : len >= 16 * nseg ? 2 : len >= 8 * nseg ? 1 : 0;
        /* beta*y of the rows this lane will write: loaded now, used after the reduction,
         * so the global-load latency hides behind the barrier and the row sums */
        double yin[4] = {0.0, 0.0, 0.0, 0.0};
        int ypre = 0;
        if (a.beta != 0.0) {
            const int G = 1 << lg, ngroups = kConsumers >> lg;
            const int grp = t >> lg;
            const bool lead = (t & (G - 1)) == 0;
            ypre = min(4, (nseg + ngroups - 1) >> (8 - lg));
            for (int rr = 0; rr < ypre; ++rr) {
                const int sg = rr * ngroups + grp;
                const int row = rs + sg - 1;
                double v = 0.0;
                if (lead && sg >= 1 && sg < nseg && row != a.skip_first && row != a.skip_last) v = a.y[row];
                if (rr == 0) yin[0] = v; else if (rr == 1) yin[1] = v; else if (rr == 2) yin[2] = v; else yin[3] = v;
            }
        }
        if (has_next) {
            mbar_wait(full0 + 8u * sn, phn);
            gather(st[sn], base + step);
        }
        named_bar_sync(bar_id, kConsumers);
        switch (lg) {
        case 5: reduce_rows<32>(a, S, j, t, base, clo, chi, rs, nown, ext, yin, ypre); break;
        case 4: reduce_rows<16>(a, S, j, t, base, clo, chi, rs, nown, ext, yin, ypre); break;
        case 3: reduce_rows<8>(a, S, j, t, base, clo, chi, rs, nown, ext, yin, ypre); break;
        case 2: reduce_rows<4>(a, S, j, t, base, clo, chi, rs, nown, ext, yin, ypre); break;
        case 1: reduce_rows<2>(a, S, j, t, base, clo, chi, rs, nown, ext, yin, ypre); break;
        default: reduce_rows<1>(a, S, j, t, base, clo, chi, rs, nown, ext, yin, ypre); break;
        }
        fence_proxy_async_smem();              /* generic writes to the slot before the next bulk copy */
        release_stage(eb, lane);
        s = sn; ph = phn; ++it;
    }
}

int g_sm_count[64] = {0};

}  // namespace

int sblas_tma_tile_size(void) { return kTile; }

cudaError_t sblas_launch_tma(const sblas_seg_args *a, cudaStream_t s)
{
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64) return cudaErrorInvalidDevice;
    if (!attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(spmv_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&g_sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    int grid = kCtasPerSm * g_sm_count[dev];
    if (grid > a->ntile) grid = a->ntile;
    spmv_tma_kernel<<<grid, kThreadsTma, kSmemBytes, s>>>(*a);
    return cudaGetLastError();
}
