/* sblas_plan.c -- host side (C) of the multi-GPU CSR SpMV path.
 *
 * Replaces the bodies of spMV_mgpu_baseline / _v1 / _v2
 * (spmv/src/dspmv_mgpu_baseline.cu:14-214, dspmv_mgpu_v1.cu:16-280,
 * dspmv_mgpu_v2.cu:33-207) with a plan/execute split:
 *
 *   plan   = reference partition (sblas_partition.c) -> segments -> one contiguous
 *            resident nnz range per GPU (val, col uploaded once), int32 row pointer
 *            rebased on the GPU, tile metadata, merge lists for split rows.
 *   execute= x replicated (sliced H2D + NVLink all-gather when several GPUs are
 *            driven from this process), one launch per segment, then ONE merge
 *            kernel per GPU that finishes rows split between segments in ascending
 *            segment order, reading the other GPUs' partial sums over NVLink P2P
 *            (reference: host loops dspmv_mgpu_v1.cu:235-248, dspmv_mgpu_v2.cu:385-441).
 *
 * There is no CPU arithmetic on this path and no CPU fallback.
 */
#include <cuda_runtime_api.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include "sblas_device.h"
#include "sblas_internal.h"
#include "sblas_spmv.h"

static __thread char g_err[512];
const char *sblas_last_error(void) { return g_err; }
void sblas_set_error(const char *fmt, const char *a, const char *b, int line)
{
    snprintf(g_err, sizeof g_err, fmt, a, b, line);
}

#define CU(call)                                                                      \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            sblas_set_error("%s failed: %s (sblas_plan.c:%d)", #call, cudaGetErrorString(e_), __LINE__); \
            rc = 1;                                                                   \
            goto fail;                                                                \
        }                                                                             \
    } while (0)

/* ------------------------------------------------------------------ helpers */
double sblas_get_time(void)          /* spmv/src/spmv_helper.cu:41-48: wall clock in seconds */
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return (double)tv.tv_sec + (double)tv.tv_usec * 1e-6;
}

double sblas_get_gpu_availble_mem(int ngpu)   /* spmv_helper.cu:51-76: min free GB over GPUs */
{
    double best = 1e300;
    for (int d = 0; d < ngpu; ++d) {
        size_t fr = 0, tot = 0;
        if (cudaSetDevice(d) != cudaSuccess || cudaMemGetInfo(&fr, &tot) != cudaSuccess) return 0.0;
        const double gb = (double)fr / 1e9;
        if (gb < best) best = gb;
    }
    return best;
}

/* last r in [0,m) with rp[r] <= idx: the row that really holds entry idx */
static int true_row_of(int m, const long long *rp, long long idx)
{
    int lo = 0, hi = m;                       /* first r in [0,m] with rp[r] > idx */
    while (lo < hi) {
        const int mid = lo + (hi - lo) / 2;
        if (rp[mid] <= idx) lo = mid + 1; else hi = mid;
    }
    return lo - 1 < 0 ? 0 : (lo - 1 >= m ? m - 1 : lo - 1);
}

static int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

/* `kernel` of the reference API -> kernel family + tile shape.
 *   1: adaptive -- the persistent TMA-fed tile kernel (its per-tile reduction adapts to
 *      the rows inside the tile) for anything big enough to fill the GPU, the
 *      lanes-per-row kernel otherwise
 *   2: nnz-balanced TMA tile kernel everywhere (the reference's "merge-path" choice)
 *   3: CSR5-style small register tiles (1024 nnz, one tile per CTA) */
static void pick_kernel(int kernel, long long dev_nnz, int *kind, int *ipt)
{
    *ipt = 16;
    *kind = SBLAS_K_TMA;
    if (kernel == 1 && dev_nnz < (long long)env_int("SBLAS_VEC_BELOW", 1 << 16)) *kind = SBLAS_K_VECTOR;
    if (kernel == 3) { *kind = SBLAS_K_TILE; *ipt = 4; }
    const char *k = getenv("SBLAS_KIND");
    if (k && !strcmp(k, "vec")) *kind = SBLAS_K_VECTOR;
    if (k && !strcmp(k, "tile")) *kind = SBLAS_K_TILE;
    if (k && !strcmp(k, "tma")) *kind = SBLAS_K_TMA;
    if (k && !strcmp(k, "vecp")) *kind = SBLAS_K_VECP;
    const int e = env_int("SBLAS_IPT", 0);
    if (e == 4 || e == 8 || e == 16) *ipt = e;
}

/* Opt-in (SBLAS_X_PERSIST=1): an L2 access-policy window over the part of x the shard reads, persisting hits /
 * streaming misses, on every stream of the GPU (the north star's "keep x in L2 via access-policy windows").
 * The default instead marks the STREAMED arrays evict-first inside the kernels (sblas_dev_common.cuh), which needs no
 * carve-out of the L2; DESIGN.md section 6 has the A/B.  Stream attributes are not captured into graph kernel
 * nodes, so sblas_spmv_plan_step falls back to plain launches while the window is on. */
static int apply_x_window_policy(sblas_spmv_plan *P, sblas_dev *D)
{
    if (!env_int("SBLAS_X_PERSIST", 0) || !D->d_x || D->col_hi < D->col_lo) return 0;
    int max_win = 0, max_persist = 0;
    if (cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, D->device) != cudaSuccess ||
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, D->device) != cudaSuccess ||
        max_win <= 0 || max_persist <= 0) { cudaGetLastError(); return 0; }
    size_t bytes = ((size_t)D->col_hi - D->col_lo + 1) * sizeof(double);
    if (bytes > (size_t)max_win) bytes = (size_t)max_win;
    size_t carve = bytes < (size_t)max_persist ? bytes : (size_t)max_persist;
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof av);
    av.accessPolicyWindow.base_ptr = (void *)(D->d_x + D->col_lo);
    av.accessPolicyWindow.num_bytes = bytes;
    av.accessPolicyWindow.hitRatio = (float)((double)carve / (double)bytes);
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    for (int c = 0; c < D->nstreams; ++c)
        if (cudaStreamSetAttribute(D->streams[c], cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) { cudaGetLastError(); return 0; }
    P->x_policy = 1;
    return 1;
}

/* ------------------------------------------------------------------ per-GPU pool of main allocations */
#define SBLAS_POOL_DEVS 64
static struct { char *ptr; size_t bytes; } g_pool[SBLAS_POOL_DEVS];
static pthread_mutex_t g_pool_mu = PTHREAD_MUTEX_INITIALIZER;

static size_t pool_bytes(int device)
{
    size_t b = 0;
    if (device >= 0 && device < SBLAS_POOL_DEVS) {
        pthread_mutex_lock(&g_pool_mu);
        b = g_pool[device].ptr ? g_pool[device].bytes : 0;
        pthread_mutex_unlock(&g_pool_mu);
    }
    return b;
}

/* the current device is `device`; *got = the size of the block handed out (>= need) */
static cudaError_t slab_get(int device, size_t need, int pooled, char **out, size_t *got)
{
    if (pooled && device >= 0 && device < SBLAS_POOL_DEVS) {
        char *old = NULL;
        pthread_mutex_lock(&g_pool_mu);
        if (g_pool[device].ptr && g_pool[device].bytes >= need) {
            *out = g_pool[device].ptr; *got = g_pool[device].bytes;
            g_pool[device].ptr = NULL; g_pool[device].bytes = 0;
            pthread_mutex_unlock(&g_pool_mu);
            return cudaSuccess;
        }
        old = g_pool[device].ptr;                              /* too small: make room before asking for more */
        g_pool[device].ptr = NULL; g_pool[device].bytes = 0;
        pthread_mutex_unlock(&g_pool_mu);
        if (old) cudaFree(old);
    }
    *got = need;
    return cudaMalloc((void **)out, need);
}

/* blocks above SBLAS_POOL_MAX_GB (default 32) are not retained: a caller that pushes one giant matrix through the
 * one-shot entry point should get the memory back like the reference gives it back */
static size_t pool_cap(void)
{
    const int gb = env_int("SBLAS_POOL_MAX_GB", 32);
    return (size_t)(gb > 0 ? gb : 0) << 30;
}

/* the current device is `device` and nothing is in flight on the block */
static void slab_put(int device, char *ptr, size_t bytes, int pooled)
{
    if (!ptr) return;
    if (pooled && device >= 0 && device < SBLAS_POOL_DEVS && bytes <= pool_cap()) {
        pthread_mutex_lock(&g_pool_mu);
        if (!g_pool[device].ptr || g_pool[device].bytes < bytes) {   /* keep the larger of the two */
            char *t = g_pool[device].ptr;
            g_pool[device].ptr = ptr; g_pool[device].bytes = bytes;
            ptr = t;
        }
        pthread_mutex_unlock(&g_pool_mu);
    }
    if (ptr) cudaFree(ptr);
}

void sblas_pool_release(void)
{
    int cur = 0;
    const int have = cudaGetDevice(&cur) == cudaSuccess;
    for (int d = 0; d < SBLAS_POOL_DEVS; ++d) {
        pthread_mutex_lock(&g_pool_mu);
        char *p = g_pool[d].ptr;
        g_pool[d].ptr = NULL; g_pool[d].bytes = 0;
        pthread_mutex_unlock(&g_pool_mu);
        if (p && cudaSetDevice(d) == cudaSuccess) cudaFree(p);
    }
    if (have) cudaSetDevice(cur); else cudaGetLastError();
}

/* ------------------------------------------------------------------ build */
static void free_dev(sblas_dev *D, int dry, int pooled)
{
    if (dry) {
        free(D->h_mrow); free(D->h_mbeg); free(D->h_msrc); free(D->h_msrc_off);
        return;
    }
    if (D->device >= 0) cudaSetDevice(D->device);
    /* every device array lives in these three */
    if (pooled && D->slab_main) cudaDeviceSynchronize();      /* the next plan writes into the block straight away */
    slab_put(D->device, D->slab_main, D->slab_main_bytes, pooled);
    D->slab_main = NULL;
    cudaFree(D->slab_tiles); cudaFree(D->slab_merge);
    if (D->streams) {
        for (int c = 0; c < D->nstreams; ++c) {
            if (D->streams[c]) cudaStreamDestroy(D->streams[c]);
            if (D->ev_seg && D->ev_seg[c]) cudaEventDestroy(D->ev_seg[c]);
        }
    }
    if (D->ev_in) cudaEventDestroy(D->ev_in);
    if (D->ev_done) cudaEventDestroy(D->ev_done);
    if (D->copy_stream) cudaStreamDestroy(D->copy_stream);
    for (int i = 0; i < D->nev_unit; ++i) if (D->ev_unit[i]) cudaEventDestroy(D->ev_unit[i]);
    free(D->ev_unit);
    if (D->ev_y) cudaEventDestroy(D->ev_y);
    if (D->ev_chain) cudaEventDestroy(D->ev_chain);
    if (D->ev_merge) cudaEventDestroy(D->ev_merge);
    if (D->ev_xpull) cudaEventDestroy(D->ev_xpull);
    free(D->streams); free(D->ev_seg);
    free(D->h_mrow); free(D->h_mbeg); free(D->h_msrc); free(D->h_msrc_off);
}

void sblas_spmv_plan_destroy(sblas_spmv_plan *P)
{
    if (!P) return;
    const double t_destroy = getenv("SBLAS_TIMING") ? sblas_get_time() : 0.0;
    if (P->x_alloc) { P->devs[0].d_x = P->x_alloc; P->x_alloc = NULL; }   /* the bound peer x is the caller's: free our own */
    for (int d = 0; d < P->ndev; ++d) free_dev(&P->devs[d], P->dry, P->pooled);
    if (P->peer_bound) {
        cudaFree(P->d_peer_bases); cudaFree(P->d_out_slot); cudaFree(P->d_out_owner); cudaFree(P->d_out_off);
        cudaFree(P->d_owners); cudaFree(P->d_contrib); cudaFree(P->d_msrc_off); cudaFree(P->d_epoch);
    }
    if (P->graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)P->graph_exec);
    if (P->d_peer_x) { cudaFree(P->d_peer_x); cudaFree(P->d_peer_xflags); cudaFree(P->d_chain_ctr); }

    free(P->devs); free(P->segs); free(P->units); free(P->parts); free(P->g_owner); free(P->g_local);
    free(P->g_lo); free(P->g_hi); free(P->g_sf); free(P->g_sl);
    free(P->piece_lo); free(P->piece_hi); free(P->piece_unit);
    free(P);
    if (t_destroy > 0.0) fprintf(stderr, "sblas plan destroy: %.3f ms\n", (sblas_get_time() - t_destroy) * 1e3);
}

/* Global segment table (every GPU / rank computes the same one): reference
 * records + the rows each segment really covers + who shares what. */
static int build_global(sblas_spmv_plan *P, const long long *rp)
{
    const int m = P->m;
    int G = 0;
    if (P->version == SBLAS_V2) {
        G = sblas_v2_num_tasks(P->nnz, P->nb);
    } else {
        G = P->world;
    }
    if (G <= 0) return -1;
    P->nparts = G;
    P->parts = (sblas_part *)calloc((size_t)G, sizeof(sblas_part));
    P->g_owner = (int *)calloc((size_t)G, sizeof(int));
    P->g_local = (int *)calloc((size_t)G, sizeof(int));
    P->g_lo = (int *)calloc((size_t)G, sizeof(int));
    P->g_hi = (int *)calloc((size_t)G, sizeof(int));
    P->g_sf = (int *)calloc((size_t)G, sizeof(int));
    P->g_sl = (int *)calloc((size_t)G, sizeof(int));
    if (!P->parts || !P->g_owner || !P->g_local || !P->g_lo || !P->g_hi || !P->g_sf || !P->g_sl) return -1;

    if (P->version == SBLAS_BASELINE) sblas_partition_baseline(m, rp, G, P->parts);
    else if (P->version == SBLAS_V1) sblas_partition_v1(m, P->nnz, rp, G, P->parts);
    else if (P->version == SBLAS_V1_BYTES) sblas_partition_bytes(m, P->nnz, rp, G, env_int("SBLAS_ROW_BYTES", SBLAS_ROW_BYTES), P->parts);
    else sblas_generate_tasks_v2(m, P->nnz, rp, P->nb, P->parts);

    int prev = -1;                 /* previous non-empty segment */
    int prev_hi = -1;
    for (int t = 0; t < G; ++t) {
        sblas_part *p = &P->parts[t];
        P->g_owner[t] = (P->version == SBLAS_V2) ? sblas_v2_task_owner(G, P->world, t) : t;
        P->g_lo[t] = 0; P->g_hi[t] = -1; P->g_sf[t] = 0; P->g_sl[t] = 0;
        if (P->version == SBLAS_BASELINE) {
            P->g_lo[t] = p->start_row; P->g_hi[t] = p->end_row;       /* may be empty (dev_m == 0) */
            continue;
        }
        if (p->end_idx < p->start_idx) continue;                       /* no nnz: nothing to run */
        const int tsr = true_row_of(m, rp, p->start_idx);
        const int ter = true_row_of(m, rp, p->end_idx);
        const int split = p->start_idx > rp[tsr];
        P->g_sf[t] = (prev >= 0) && split;
        P->g_lo[t] = P->g_sf[t] ? tsr : prev_hi + 1;
        P->g_hi[t] = ter;
        if (prev >= 0) P->g_sl[prev] = P->g_sf[t];
        prev = t; prev_hi = ter;
    }
    if (P->version != SBLAS_BASELINE) {
        if (prev >= 0) {
            P->g_hi[prev] = m - 1;                                     /* trailing empty rows */
        } else {
            /* matrix without entries: the first segment scales y */
            P->g_lo[0] = 0; P->g_hi[0] = m - 1;
            P->parts[0].start_idx = 0; P->parts[0].end_idx = -1;
        }
    }
    return 0;
}

static int seg_is_live(const sblas_spmv_plan *P, int t) { return P->g_hi[t] >= P->g_lo[t]; }

/* owner (global segment) of the row shared at the START of segment t: walk back */
static int row_owner_seg(const sblas_spmv_plan *P, int t)
{
    int o = t;
    while (P->g_sf[o]) {
        int k = o - 1;
        while (k >= 0 && !seg_is_live(P, k)) --k;
        o = k;
        /* o's last row is the shared row; if o is a single-row segment that is
         * itself split at its start, keep walking */
        if (!(P->g_sf[o] && P->g_lo[o] == P->g_hi[o])) break;
    }
    return o;
}

static sblas_unit *new_unit(sblas_spmv_plan *P)
{
    if (P->nunits == P->cap_units) {
        const int cap = P->cap_units ? 2 * P->cap_units : 16;
        sblas_unit *u = (sblas_unit *)realloc(P->units, (size_t)cap * sizeof(sblas_unit));
        if (!u) return NULL;
        P->units = u; P->cap_units = cap;
    }
    sblas_unit *U = &P->units[P->nunits++];
    memset(U, 0, sizeof *U);
    return U;
}

#define SBLAS_PANEL_ROWS 4096          /* rows per statistics block */
#define SBLAS_PANEL_MAX 32             /* most panels per segment (else: one panel) */

/* Adaptive row binning at plan level: cut a segment into panels of consecutive rows by the
 * longest row of every 4096-row block (computed on the GPU):
 *   class 1 "short"   every row holds at most `short_max` entries -> thread-per-row kernel
 *   class 2 "medium"  longest row in [16, 256] and the rows fill at least half of a warp's
 *                     256-entry window -> warp per R = floor(256/longest) whole rows (row-tile kernel)
 *   class 3 "long-medium"  longest row in (256, 2048] and the mean row at least half of the 256*G
 *                     entries G = 2, 4 or 8 warps cover -> G warps per row (row-split kernel); off
 *                     unless `medium_on` has bit 1 set
 *   class 0 "general" everything else -> the GPU's general kernel (the nnz-balanced TMA tile
 *                     kernel, whose per-tile reduction adapts further)
 * Runs below `min_nnz` entries are not worth a launch of their own and turn general; equal
 * neighbours merge (medium runs keep the smallest R).  run_class/run_R/run_begin: outputs (run i
 * covers blocks [run_begin[i], run_begin[i+1])); returns the number of runs. */
/* (R | longest << 8) of two medium runs that merge: the smaller R, the longer row */
static int merge_R(int x, int y)
{
    const int R = (x & 0xff) < (y & 0xff) ? (x & 0xff) : (y & 0xff);
    const int mx = (x >> 8) > (y >> 8) ? (x >> 8) : (y >> 8);
    return R | mx << 8;
}

static int bin_blocks(const int *bmax, const int *bptr, int nblk, int nrows, int nz1, int short_max, int medium_on,
                      long long min_nnz, int *run_class, int *run_R, int *run_begin)
{
    int nrun = 0;
    for (int b = 0; b < nblk; ++b) {
        const long long bn = (b + 1 < nblk ? bptr[b + 1] : nz1) - (long long)bptr[b];
        const long long br = (b + 1 < nblk) ? SBLAS_PANEL_ROWS : nrows - (long long)b * SBLAS_PANEL_ROWS;
        int cls = 0, R = 0;
        if (bmax[b] <= short_max) cls = 1;
        else if (medium_on && bmax[b] >= 16 && bmax[b] <= 256) {
            R = 256 / bmax[b];
            if (R > 8) R = 8;
            if (bn * R >= 128 * br) cls = 2;
        } else if ((medium_on & 2) && bmax[b] > 256 && bmax[b] <= 2048) {
            R = bmax[b] <= 512 ? 2 : bmax[b] <= 1024 ? 4 : 8;           /* G warps per row */
            if (bn >= 128LL * R * br) cls = 3;
        }
        if (cls == 2) R |= bmax[b] << 8;                     /* R | longest row << 8 */
        if (nrun > 0 && run_class[nrun - 1] == cls) {
            if (cls == 2) run_R[nrun - 1] = merge_R(run_R[nrun - 1], R);
            if (cls == 3 && R > run_R[nrun - 1]) run_R[nrun - 1] = R;
            continue;
        }
        run_class[nrun] = cls; run_R[nrun] = R; run_begin[nrun] = b; ++nrun;
    }
    run_begin[nrun] = nblk;
    for (int i = 0; i < nrun; ++i) {
        const long long e = run_begin[i + 1] < nblk ? bptr[run_begin[i + 1]] : nz1;
        if (run_class[i] != 0 && e - bptr[run_begin[i]] < min_nnz) run_class[i] = 0;
    }
    int w = 0;
    for (int i = 0; i < nrun; ++i) {
        if (w > 0 && run_class[w - 1] == run_class[i]) {
            if (run_class[i] == 2) run_R[w - 1] = merge_R(run_R[w - 1], run_R[i]);
            if (run_class[i] == 3 && run_R[i] > run_R[w - 1]) run_R[w - 1] = run_R[i];
            continue;
        }
        run_class[w] = run_class[i]; run_R[w] = run_R[i]; run_begin[w] = run_begin[i]; ++w;
    }
    run_begin[w] = nblk;
    return w;
}

/* the binning rule alone, on host arrays (tests and tools; no CUDA): see bin_blocks */
int sblas_bin_row_blocks(const int *block_longest, const int *block_first_entry, int nblocks, int nrows, int nz_end,
                         int short_max, int medium_on, long long min_nnz, int *run_class, int *run_R, int *run_begin)
{
    const int n = bin_blocks(block_longest, block_first_entry, nblocks, nrows, nz_end, short_max, medium_on, min_nnz,
                             run_class, run_R, run_begin);
    for (int i = 0; i < n; ++i) if (run_class[i] == 2) run_R[i] &= 0xff; else if (run_class[i] != 3) run_R[i] = 0;
    return n;
}

static int plan_build(sblas_spmv_plan *P, const double *val, const long long *rp, const int *col,
                      const int *devices, int src_flags)
{
    int rc = 0;
    int *d_stats = NULL, *h_stats = NULL, stats_cap = 0;      /* row-block statistics scratch */
    const int dry = (src_flags & SBLAS_LAYOUT_ONLY) != 0;     /* host layout only: no CUDA call at all */
    P->dry = dry;
    /* SBLAS_TIMING: where a plan build spends its wall time (stderr, one line per plan) */
    const int timing = getenv("SBLAS_TIMING") != NULL;
    double tm[8] = {0};
    int ntm = 0;
#define STAMP() do { if (timing && ntm < 8) tm[ntm++] = sblas_get_time(); } while (0)
    STAMP();
    if (build_global(P, rp) != 0) { sblas_set_error("%s%s (line %d)", "partition failed", "", __LINE__); return -1; }

    /* ---- local segments per GPU */
    const int ndev = P->ndev;
    int nloc = 0;
    for (int t = 0; t < P->nparts; ++t) {
        P->g_local[t] = -1;
        if (!seg_is_live(P, t)) continue;
        const int o = P->g_owner[t];
        const int d = P->rank_mode ? (o == P->rank ? 0 : -1) : o;
        if (d >= 0) ++nloc;
    }
    P->nseg = nloc;
    P->segs = (sblas_seg *)calloc((size_t)(nloc > 0 ? nloc : 1), sizeof(sblas_seg));
    for (int d = 0; d < ndev; ++d) { P->devs[d].seg_begin = -1; P->devs[d].seg_end = -1; }
    int k = 0;
    for (int t = 0; t < P->nparts; ++t) {
        if (!seg_is_live(P, t)) continue;
        const int o = P->g_owner[t];
        const int d = P->rank_mode ? (o == P->rank ? 0 : -1) : o;
        if (d < 0) continue;
        sblas_seg *S = &P->segs[k];
        S->gidx = t; S->dev = d;
        sblas_dev *D = &P->devs[d];
        if (D->seg_begin < 0) D->seg_begin = k;
        D->seg_end = k + 1;
        S->lidx = k - D->seg_begin;
        P->g_local[t] = S->lidx;
        ++k;
    }
    /* edge slot of a global segment = 2 * (index among its owner's live segments) */
    {
        int *cnt = (int *)calloc((size_t)P->world, sizeof(int));
        P->max_local = 0;
        for (int t = 0; t < P->nparts; ++t) {
            if (!seg_is_live(P, t)) { P->g_local[t] = -1; continue; }
            P->g_local[t] = cnt[P->g_owner[t]]++;
            if (cnt[P->g_owner[t]] > P->max_local) P->max_local = cnt[P->g_owner[t]];
        }
        free(cnt);
    }

    /* ---- per GPU: resident range, buffers, upload */
    for (int d = 0; d < ndev; ++d) {
        sblas_dev *D = &P->devs[d];
        D->device = devices[d];
        if (D->seg_begin < 0) { D->rows = 0; D->nnz = 0; continue; }
        const sblas_seg *S0 = &P->segs[D->seg_begin], *S1 = &P->segs[D->seg_end - 1];
        D->first_idx = P->parts[S0->gidx].start_idx;
        D->last_idx = P->parts[S1->gidx].end_idx;
        for (int s = D->seg_begin; s < D->seg_end; ++s)           /* baseline blocks may be empty of nnz */
            if (P->parts[P->segs[s].gidx].end_idx > D->last_idx) D->last_idx = P->parts[P->segs[s].gidx].end_idx;
        D->first_row = P->g_lo[S0->gidx];
        D->last_row = P->g_hi[S1->gidx];
        D->rows = D->last_row - D->first_row + 1;
        const long long dn = D->last_idx - D->first_idx + 1;
        if (dn > 2147483647LL - 65536) {
            sblas_set_error("%s%s (line %d)", "per-GPU nnz exceeds int32 (reference casts to int too)", "", __LINE__);
            return -1;
        }
        D->nnz = (int)(dn < 0 ? 0 : dn);
        D->nstreams = P->q > 1 ? P->q : 1;

        if (!dry) {
        CU(cudaSetDevice(D->device));
        /* memory guard of the reference: shard > 0.8 x free -> -1
         * (dspmv_mgpu_baseline.cu:70-79, dspmv_mgpu_v1.cu:106-116) */
        {
            size_t fr = 0, tot = 0;
            CU(cudaMemGetInfo(&fr, &tot));
            fr += pool_bytes(D->device);                       /* the retained block is ours to reuse */
            const double need = 12.0 * D->nnz * ((src_flags & SBLAS_SRC_DEVICE_SHARD) ? 0.0 : 1.0) +
                                4.0 * (D->rows + 1) + 8.0 * P->n + 8.0 * D->rows;
            if (need / 1e9 > 0.8 * ((double)fr / 1e9)) {
                sblas_set_error("%s%s (line %d)", "shard exceeds 0.8 x free device memory", "", __LINE__);
                rc = -1; goto fail;
            }
        }
        D->streams = (cudaStream_t *)calloc((size_t)D->nstreams, sizeof(cudaStream_t));
        D->ev_seg = (cudaEvent_t *)calloc((size_t)D->nstreams, sizeof(cudaEvent_t));
        for (int c = 0; c < D->nstreams; ++c) {
            CU(cudaStreamCreateWithFlags(&D->streams[c], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&D->ev_seg[c], cudaEventDisableTiming));
        }
        CU(cudaEventCreateWithFlags(&D->ev_in, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&D->ev_done, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&D->ev_y, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&D->ev_chain, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&D->ev_merge, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&D->ev_xpull, cudaEventDisableTiming));
        cudaStream_t st = D->streams[0];
        (void)st;

        /* one allocation for everything whose size is known now, each array on a 256-byte boundary */
        D->own_matrix = !(src_flags & SBLAS_SRC_DEVICE_SHARD);
        const size_t nedge = (size_t)(2 * (P->rank_mode ? P->max_local : D->seg_end - D->seg_begin) + 2);
        size_t off = 0;
#define SLOT(bytes) (off = (off + 255) & ~(size_t)255, off += (size_t)(bytes), off - (size_t)(bytes))
        /* +16: bulk copies round the last tile up to 16 bytes */
        const size_t o_val = D->own_matrix ? SLOT(((size_t)D->nnz + 16) * sizeof(double)) : 0;
        const size_t o_col = D->own_matrix ? SLOT(((size_t)D->nnz + 16) * sizeof(int)) : 0;
        const size_t o_rp = SLOT(((size_t)D->rows + 1 + 8) * sizeof(int));
        const size_t o_x = SLOT((size_t)(P->n > 0 ? P->n : 1) * sizeof(double));
        const size_t o_y = SLOT(((size_t)D->rows + 2) * sizeof(double));      /* doubles as the int64 row pointer stage */
        const size_t o_edge = SLOT(nedge * sizeof(double));
        const size_t o_mm = SLOT(2 * sizeof(int));
#undef SLOT
        CU(slab_get(D->device, off, P->pooled, &D->slab_main, &D->slab_main_bytes));
        if (D->own_matrix) {
            D->d_val = (double *)(D->slab_main + o_val); D->d_col = (int *)(D->slab_main + o_col);
            if (D->nnz > 0) {
                CU(cudaMemcpyAsync(D->d_val, val + D->first_idx, (size_t)D->nnz * sizeof(double), cudaMemcpyHostToDevice, st));
                CU(cudaMemcpyAsync(D->d_col, col + D->first_idx, (size_t)D->nnz * sizeof(int), cudaMemcpyHostToDevice, st));
            }
        } else {
            D->d_val = (double *)val; D->d_col = (int *)col;
        }
        D->d_rowptr = (int *)(D->slab_main + o_rp);
        D->d_x = (double *)(D->slab_main + o_x);
        D->d_y = (double *)(D->slab_main + o_y);
        D->d_edge = (double *)(D->slab_main + o_edge);      /* 2 doubles per local segment (peers read it over NVLink) */
        D->d_mm = (int *)(D->slab_main + o_mm);
        CU(cudaMemsetAsync(D->d_rowptr, 0, ((size_t)D->rows + 1 + 8) * sizeof(int), st));
        CU(cudaMemsetAsync(D->d_edge, 0, nedge * sizeof(double), st));
        /* int64 host row pointer slice -> int32 rebased/clamped, on the GPU; staged in y's storage, which nothing
         * touches before the first product */
        D->stage64 = (long long *)D->d_y;
        CU(cudaMemcpyAsync(D->stage64, rp + D->first_row, (size_t)(D->rows + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
        CU(sblas_launch_rebase_rowptr(D->stage64, D->first_idx, D->nnz, (long long)D->rows + 1, D->d_rowptr, st));
        }   /* !dry */
    }
    STAMP();                                                   /* [1] partition, allocations, uploads enqueued */

    /* ---- second pass per GPU (the uploads of ALL GPUs are in flight by now, each over its own PCIe link):
     * wait for this GPU's shard, then column window, panels, tile metadata */
    for (int d = 0; d < ndev; ++d) {
        sblas_dev *D = &P->devs[d];
        if (D->seg_begin < 0) continue;
        if (!dry) {
        CU(cudaSetDevice(D->device));
        cudaStream_t st = D->streams[0];
        CU(cudaStreamSynchronize(st));
        D->stage64 = NULL;

        /* the window of x this shard reads: [col_lo, col_hi] (one reduction over col at plan time) */
        D->col_lo = 0; D->col_hi = P->n - 1;
        if (D->nnz > 0 && env_int("SBLAS_X_WINDOW", 1)) {
            int h_mm[2] = {0x7fffffff, -1};
            CU(cudaMemcpyAsync(D->d_mm, h_mm, sizeof h_mm, cudaMemcpyHostToDevice, st));
            CU(sblas_launch_col_range(D->d_col, D->nnz, D->d_mm, st));
            CU(cudaMemcpyAsync(h_mm, D->d_mm, sizeof h_mm, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            if (h_mm[0] >= 0 && h_mm[1] < P->n && h_mm[0] <= h_mm[1]) { D->col_lo = h_mm[0]; D->col_hi = h_mm[1]; }
        }

        apply_x_window_policy(P, D);

        }   /* !dry */

        /* ---- segments: kernel choice, panels, tiles */
        pick_kernel(P->kernel, D->nnz, &D->kind, &D->ipt);
        const int panels_on = !dry && P->kernel == 1 && D->kind == SBLAS_K_TMA && !getenv("SBLAS_KIND") &&
                              env_int("SBLAS_PANELS", 1) != 0;
        const int short_max = env_int("SBLAS_SHORT_MAX", 4);
        const int medium_on = env_int("SBLAS_MEDIUM", 3);       /* bit 0: row-tile panels, bit 1: row-split panels */
        const long long panel_min_nnz = env_int("SBLAS_PANEL_MIN_NNZ", 1 << 20);
        long long tiles_total = 0;
        for (int s = D->seg_begin; s < D->seg_end; ++s) {
            sblas_seg *S = &P->segs[s];
            const sblas_part *p = &P->parts[S->gidx];
            sblas_seg_args *a = &S->args;
            memset(a, 0, sizeof *a);
            a->row_lo = P->g_lo[S->gidx] - D->first_row;
            a->row_hi = P->g_hi[S->gidx] - D->first_row;
            a->nz0 = (int)(p->start_idx - D->first_idx);
            a->nz1 = (int)(p->end_idx + 1 - D->first_idx);
            if (a->nz1 < a->nz0) a->nz1 = a->nz0;
            a->skip_first = P->g_sf[S->gidx] ? a->row_lo : -1;
            a->skip_last = P->g_sl[S->gidx] ? a->row_hi : -1;
            a->nz_total = D->nnz;
            a->mode = env_int("SBLAS_TMA_MODE", 0);
            if (!dry) {
                a->val = D->d_val; a->col = D->d_col; a->rowptr = D->d_rowptr;
                a->x = D->d_x; a->y = D->d_y;
                a->edge = D->d_edge + 2 * S->lidx;
            }
            S->stream = S->lidx % D->nstreams;
            S->unit_begin = P->nunits;

            /* panels: runs of row blocks of one class */
            const int nrows = a->row_hi - a->row_lo + 1;
            const int nblk = (nrows + SBLAS_PANEL_ROWS - 1) / SBLAS_PANEL_ROWS;
            int nrun = 1;
            int run_class[SBLAS_PANEL_MAX + 2] = {0}, run_R[SBLAS_PANEL_MAX + 2] = {0}, run_begin[SBLAS_PANEL_MAX + 2] = {0};
            const int *bptr = NULL;
            if (panels_on && nblk >= 1 && (long long)a->nz1 - a->nz0 >= 2 * panel_min_nnz) {
                if (nblk > stats_cap) {
                    if (d_stats) cudaFree(d_stats);
                    free(h_stats);
                    d_stats = NULL; h_stats = NULL;
                    stats_cap = nblk;
                    CU(cudaMalloc((void **)&d_stats, (size_t)2 * stats_cap * sizeof(int)));
                    h_stats = (int *)malloc((size_t)2 * stats_cap * sizeof(int));
                    if (!h_stats) { rc = 1; goto fail; }
                }
                cudaStream_t st0 = D->streams[0];
                CU(sblas_launch_row_block_stats(D->d_rowptr, a->row_lo, nrows, SBLAS_PANEL_ROWS, a->nz0, a->nz1,
                                                d_stats, d_stats + nblk, st0));
                CU(cudaMemcpyAsync(h_stats, d_stats, (size_t)2 * nblk * sizeof(int), cudaMemcpyDeviceToHost, st0));
                CU(cudaStreamSynchronize(st0));
                bptr = h_stats + nblk;
                int *rc_all = (int *)malloc((size_t)(nblk + 2) * sizeof(int));
                int *rr_all = (int *)malloc((size_t)(nblk + 2) * sizeof(int));
                int *rb_all = (int *)malloc((size_t)(nblk + 2) * sizeof(int));
                if (!rc_all || !rr_all || !rb_all) { free(rc_all); free(rr_all); free(rb_all); rc = 1; goto fail; }
                const int nr = bin_blocks(h_stats, bptr, nblk, nrows, a->nz1, short_max, medium_on, panel_min_nnz,
                                          rc_all, rr_all, rb_all);
                if (nr >= 2 && nr <= SBLAS_PANEL_MAX) {
                    nrun = nr;
                    memcpy(run_class, rc_all, (size_t)nr * sizeof(int));
                    memcpy(run_R, rr_all, (size_t)nr * sizeof(int));
                    memcpy(run_begin, rb_all, (size_t)(nr + 1) * sizeof(int));
                } else if (nr == 1) {
                    run_class[0] = rc_all[0]; run_R[0] = rr_all[0];
                }
                free(rc_all); free(rr_all); free(rb_all);
            }
            if (nrun == 1) { run_begin[0] = 0; run_begin[1] = nblk; }
            for (int i = 0; i < nrun; ++i) {
                sblas_unit *U = new_unit(P);
                if (!U) { rc = 1; goto fail; }
                U->args = *a;
                U->kind = run_class[i] == 1 ? SBLAS_K_SHORT : run_class[i] == 2 ? SBLAS_K_ROWTILE :
                          run_class[i] == 3 ? SBLAS_K_ROWSPLIT : D->kind;
                U->ipt = D->ipt;
                if (run_class[i] == 3) U->ipt = run_R[i];          /* G warps per row */
                if (run_class[i] == 2) {           /* R rows per warp, window = R x longest row (<= 256) */
                    const int R = run_R[i] & 0xff, win = R * (run_R[i] >> 8);
                    U->ipt = R | (win < 256 ? win : 256) << 8;
                }
                if (nrun > 1) {
                    sblas_seg_args *u = &U->args;
                    u->row_lo = a->row_lo + run_begin[i] * SBLAS_PANEL_ROWS;
                    u->row_hi = i + 1 < nrun ? a->row_lo + run_begin[i + 1] * SBLAS_PANEL_ROWS - 1 : a->row_hi;
                    if (i > 0) { u->nz0 = bptr[run_begin[i]]; u->skip_first = -1; }
                    if (i + 1 < nrun) { u->nz1 = bptr[run_begin[i + 1]]; u->skip_last = -1; }
                }
                sblas_seg_args *u = &U->args;
                const int tiled = (U->kind == SBLAS_K_TMA || U->kind == SBLAS_K_TILE);
                const int TILE = tiled ? sblas_tile_size_kind(U->kind, U->ipt) : 1;
                u->tile0 = tiled ? u->nz0 / TILE : 0;
                u->ntile = (tiled && u->nz1 > u->nz0) ? (int)(((long long)u->nz1 - 1) / TILE - u->tile0 + 1) : 0;
                U->tile_off = tiles_total;
                if (tiled) tiles_total += u->ntile + 1;
            }
            S->unit_end = P->nunits;
            S->args.tile0 = P->units[S->unit_begin].args.tile0;          /* informational */
            S->args.ntile = P->units[S->unit_begin].args.ntile;
        }
        if (d_stats) cudaFree(d_stats);
        free(h_stats);
        d_stats = NULL; h_stats = NULL; stats_cap = 0;
        if (dry) continue;
        cudaStream_t st = D->streams[0];
        if (tiles_total > 0) {
            const size_t nt = (((size_t)tiles_total + 1) + 31) & ~(size_t)31;      /* keeps every array 256-byte aligned */
            CU(cudaMalloc((void **)&D->slab_tiles, nt * (8 * sizeof(int) + sizeof(int) + 2 * sizeof(double))));
            D->d_tmeta = (int *)D->slab_tiles;
            D->d_carry = (double *)(D->slab_tiles + nt * 8 * sizeof(int));
            D->d_tail = D->d_carry + nt;
            D->d_tstart = (int *)(D->d_tail + nt);
        }
        for (int s = D->seg_begin; s < D->seg_end; ++s) {
            for (int ui = P->segs[s].unit_begin; ui < P->segs[s].unit_end; ++ui) {
                sblas_unit *U = &P->units[ui];
                sblas_seg_args *u = &U->args;
                if (U->kind != SBLAS_K_TMA && U->kind != SBLAS_K_TILE) continue;
                const int TILE = sblas_tile_size_kind(U->kind, U->ipt);
                u->tstart = D->d_tstart + U->tile_off;
                u->tmeta = D->d_tmeta + 8 * U->tile_off;
                u->carry = D->d_carry + U->tile_off;
                u->tail = D->d_tail + U->tile_off;
                if (u->ntile > 0) {
                    CU(sblas_launch_tile_rows(u, TILE, D->d_tstart + U->tile_off, st));
                    CU(sblas_launch_tile_meta(u, TILE, D->d_tmeta + 8 * U->tile_off, st));
                }
            }
        }
        CU(cudaStreamSynchronize(st));
    }

    STAMP();                                                   /* [2] uploads landed, panels, tile metadata */
    /* ---- peer access between the GPUs of an in-process plan */
    P->p2p = 0;
    if (!dry && !P->rank_mode && ndev > 1) {
        P->p2p = 1;
        for (int a = 0; a < ndev && P->p2p; ++a)
            for (int b = 0; b < ndev; ++b) {
                if (a == b) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, P->devs[a].device, P->devs[b].device);
                if (!can) { P->p2p = 0; break; }
            }
        if (P->p2p) {
            for (int a = 0; a < ndev; ++a) {
                cudaSetDevice(P->devs[a].device);
                for (int b = 0; b < ndev; ++b) {
                    if (a == b) continue;
                    cudaError_t e = cudaDeviceEnablePeerAccess(P->devs[b].device, 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                    else if (e != cudaSuccess) { cudaGetLastError(); P->p2p = 0; }
                }
            }
        }
        if (!P->p2p) {
            sblas_set_error("%s%s (line %d)", "multi-GPU plan needs peer access between all GPUs (NVLink/NVSwitch)", "", __LINE__);
            rc = 1; goto fail;
        }
    }

    STAMP();                                                   /* [3] peer access */
    /* ---- merge lists: one entry per split row, on the GPU that owns the row's start.
     * Sources are listed in ascending global segment order (deterministic sum). */
    for (int d = 0; d < ndev; ++d) {
        sblas_dev *D = &P->devs[d];
        if (D->seg_begin < 0) continue;
        int nrow = 0, nsrc = 0;
        /* first pass counts, second pass fills */
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1) {
                D->h_mrow = (int *)calloc((size_t)nrow + 1, sizeof(int));
                D->h_mbeg = (int *)calloc((size_t)nrow + 2, sizeof(int));
                D->h_msrc = (const double **)calloc((size_t)nsrc + 1, sizeof(double *));
                D->h_msrc_off = (long long *)calloc((size_t)nsrc + 1, sizeof(long long));
                nrow = 0; nsrc = 0;
            }
            for (int t = 0; t < P->nparts; ++t) {
                if (!seg_is_live(P, t) || !P->g_sl[t]) continue;
                /* t's last row is split with t+1...; only start a list where the row STARTS */
                if (P->g_sf[t] && P->g_lo[t] == P->g_hi[t]) continue;          /* middle of a longer chain */
                const int od = P->rank_mode ? (P->g_owner[t] == P->rank ? 0 : -1) : P->g_owner[t];
                if (od != d) continue;
                if (pass == 1) { D->h_mrow[nrow] = P->g_hi[t] - D->first_row; D->h_mbeg[nrow] = nsrc; }
                /* chain: t (edge[1]), then following live segments while they are split at their start */
                int u = t, which = 1;
                for (;;) {
                    if (pass == 1) {
                        const long long off = 2LL * P->g_local[u] + which;
                        D->h_msrc_off[nsrc] = (long long)P->g_owner[u] * (2LL * P->max_local) + off;
                        if (!P->rank_mode) D->h_msrc[nsrc] = P->devs[P->g_owner[u]].d_edge + off;
                    }
                    ++nsrc;
                    /* next contributor */
                    int v = u + 1;
                    while (v < P->nparts && !seg_is_live(P, v)) ++v;
                    if (v >= P->nparts || !P->g_sf[v]) break;
                    const int cont = (which == 1) ? P->g_sl[u] : (P->g_lo[u] == P->g_hi[u] && P->g_sl[u]);
                    if (!cont) break;
                    u = v; which = 0;
                }
                ++nrow;
            }
        }
        D->nmerge = nrow; D->nmsrc = nsrc;
        D->h_mbeg[nrow] = nsrc;
        if (nrow > 0 && !dry) {
            CU(cudaSetDevice(D->device));
            const size_t o_beg = (((size_t)nsrc * sizeof(double *)) + 255) & ~(size_t)255;
            const size_t o_row = (o_beg + ((size_t)nrow + 1) * sizeof(int) + 255) & ~(size_t)255;
            CU(cudaMalloc((void **)&D->slab_merge, o_row + (size_t)nrow * sizeof(int)));
            D->d_msrc = (const double **)D->slab_merge;
            D->d_mbeg = (int *)(D->slab_merge + o_beg);
            D->d_mrow = (int *)(D->slab_merge + o_row);
            CU(cudaMemcpy(D->d_mrow, D->h_mrow, (size_t)nrow * sizeof(int), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(D->d_mbeg, D->h_mbeg, (size_t)(nrow + 1) * sizeof(int), cudaMemcpyHostToDevice));
            if (!P->rank_mode)
                CU(cudaMemcpy(D->d_msrc, D->h_msrc, (size_t)nsrc * sizeof(double *), cudaMemcpyHostToDevice));
        }
    }
    STAMP();                                                   /* [4] merge lists */
    if (timing && ntm == 5)
        fprintf(stderr, "sblas plan build (%d GPU%s): alloc+enqueue %.3f ms, upload wait+panels+tiles %.3f ms, peer access %.3f ms, "
                "merge lists %.3f ms\n", ndev, ndev > 1 ? "s" : "", (tm[1] - tm[0]) * 1e3, (tm[2] - tm[1]) * 1e3,
                (tm[3] - tm[2]) * 1e3, (tm[4] - tm[3]) * 1e3);
#undef STAMP
    return 0;
fail:
    if (d_stats) cudaFree(d_stats);
    free(h_stats);
    return rc;
}

static int plan_alloc(sblas_spmv_plan **out, int version, int m, int n, long long nnz, int world, int ndev,
                      int kernel, long long nb, int q)
{
    if (!out || m <= 0 || n <= 0 || nnz < 0 || world <= 0) {
        sblas_set_error("%s%s (line %d)", "invalid argument", "", __LINE__);
        return -1;
    }
    if (kernel != 1 && kernel != 2 && kernel != 3) {
        sblas_set_error("%s%s (line %d)", "kernel must be 1, 2 or 3", "", __LINE__);
        return -1;
    }
    sblas_spmv_plan *P = (sblas_spmv_plan *)calloc(1, sizeof *P);
    if (!P) return 1;
    P->version = version; P->m = m; P->n = n; P->nnz = nnz; P->world = world; P->ndev = ndev;
    P->kernel = kernel; P->nb = nb; P->q = q > 0 ? q : 1;
    P->devs = (sblas_dev *)calloc((size_t)ndev, sizeof(sblas_dev));
    for (int d = 0; d < ndev; ++d) P->devs[d].device = -1;
    *out = P;
    return 0;
}

/* v2's nb clamp and argument check, dspmv_mgpu_v2.cu:43-46 */
static int v2_clamp(long long *nb, int ngpu, int q)
{
    if (ngpu <= 0 || q <= 0) return -1;
    const double free_gb = sblas_get_gpu_availble_mem(ngpu);
    const long long cap = (long long)(0.8 * free_gb * 1e9 / 16.0) / q;
    if (*nb > cap) *nb = cap;
    return (*nb <= 0) ? -1 : 0;
}

int sblas_spmv_plan_create(sblas_spmv_plan **plan, int version, int m, int n, long long nnz,
                           const double *csrVal, const long long *csrRowPtr, const int *csrColIndex,
                           int ngpu, int kernel, long long nb, int q)
{
    return sblas_spmv_plan_create_flags(plan, version, m, n, nnz, csrVal, csrRowPtr, csrColIndex, ngpu, kernel, nb, q, 0);
}

int sblas_spmv_plan_create_flags(sblas_spmv_plan **plan, int version, int m, int n, long long nnz,
                                 const double *csrVal, const long long *csrRowPtr, const int *csrColIndex,
                                 int ngpu, int kernel, long long nb, int q, int flags)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < ngpu || ngpu <= 0) {
        cudaGetLastError();
        sblas_set_error("%s%s (line %d)", "not enough CUDA devices (no CPU fallback)", "", __LINE__);
        return (version == SBLAS_V2) ? -1 : 1;
    }
    if (version == SBLAS_V2 && v2_clamp(&nb, ngpu, q) != 0) return -1;
    int rc = plan_alloc(plan, version, m, n, nnz, ngpu, ngpu, kernel, nb, version == SBLAS_V2 ? q : 1);
    if (rc) return rc;
    int devices[64];
    if (ngpu > 64) { sblas_spmv_plan_destroy(*plan); *plan = NULL; return -1; }
    for (int d = 0; d < ngpu; ++d) devices[d] = d;
    (*plan)->rank_mode = 0;
    (*plan)->pooled = (flags & SBLAS_CREATE_POOLED) != 0;
    rc = plan_build(*plan, csrVal, csrRowPtr, csrColIndex, devices, SBLAS_SRC_HOST);
    if (rc) { sblas_spmv_plan_destroy(*plan); *plan = NULL; }
    return rc;
}

int sblas_spmv_plan_create_rank(sblas_spmv_plan **plan, int version, int m, int n, long long nnz,
                                const double *csrVal, const long long *csrRowPtr, const int *csrColIndex,
                                int world, int rank, int device, int kernel, long long nb, int q, int flags)
{
    if (rank < 0 || rank >= world) return -1;
    if (!(flags & SBLAS_LAYOUT_ONLY) && cudaSetDevice(device) != cudaSuccess) {
        cudaGetLastError();
        sblas_set_error("%s%s (line %d)", "cudaSetDevice failed (no CPU fallback)", "", __LINE__);
        return 1;
    }
    if (version == SBLAS_V2 && (nb <= 0 || q <= 0)) return -1;
    int rc = plan_alloc(plan, version, m, n, nnz, world, 1, kernel, nb, version == SBLAS_V2 ? q : 1);
    if (rc) return rc;
    (*plan)->rank_mode = 1; (*plan)->rank = rank;
    rc = plan_build(*plan, csrVal, csrRowPtr, csrColIndex, &device, flags);
    if (rc) { sblas_spmv_plan_destroy(*plan); *plan = NULL; }
    return rc;
}

/* ------------------------------------------------------------------ execute */
static int enqueue_segments(sblas_spmv_plan *P, int d, double alpha, double beta)
{
    int rc = 0;
    sblas_dev *D = &P->devs[d];
    if (D->seg_begin < 0) return 0;
    CU(cudaSetDevice(D->device));
    /* in-process multi-GPU: the segments below overwrite this GPU's edge table, which the other GPUs'
     * merge kernels of the PREVIOUS product read over NVLink: wait for those merges first */
    if (!P->rank_mode && P->ndev > 1)
        for (int o = 0; o < P->ndev; ++o)
            if (o != d && P->devs[o].merge_recorded) CU(cudaStreamWaitEvent(D->streams[0], P->devs[o].ev_merge, 0));
    if (D->nstreams > 1) {
        CU(cudaEventRecord(D->ev_in, D->streams[0]));
        for (int c = 1; c < D->nstreams; ++c) CU(cudaStreamWaitEvent(D->streams[c], D->ev_in, 0));
    }
    for (int s = D->seg_begin; s < D->seg_end; ++s) {
        sblas_seg *S = &P->segs[s];
        S->args.alpha = alpha; S->args.beta = beta;
        for (int ui = S->unit_begin; ui < S->unit_end; ++ui) {
            sblas_unit *U = &P->units[ui];
            U->args.alpha = alpha; U->args.beta = beta;
            U->args.edge = S->args.edge;
            CU(sblas_launch_spmv_segment(&U->args, U->kind, U->ipt, 0, D->streams[S->stream]));
        }
    }
    for (int c = 1; c < D->nstreams; ++c) {
        CU(cudaEventRecord(D->ev_seg[c], D->streams[c]));
        CU(cudaStreamWaitEvent(D->streams[0], D->ev_seg[c], 0));
    }
    CU(cudaEventRecord(D->ev_done, D->streams[0]));
fail:
    return rc;
}

static int enqueue_merge(sblas_spmv_plan *P, int d, double alpha, double beta)
{
    int rc = 0;
    sblas_dev *D = &P->devs[d];
    if (D->seg_begin < 0 || D->nmerge == 0) return 0;
    CU(cudaSetDevice(D->device));
    /* wait for every GPU that contributes a partial sum (cross-device event wait) */
    if (!P->rank_mode)
        for (int o = 0; o < P->ndev; ++o)
            if (o != d && P->devs[o].seg_begin >= 0) CU(cudaStreamWaitEvent(D->streams[0], P->devs[o].ev_done, 0));
    CU(sblas_launch_edge_merge(D->d_mrow, D->d_mbeg, (const double *const *)D->d_msrc, D->nmerge, D->d_y,
                               alpha, beta, D->streams[0]));
    if (!P->rank_mode && P->ndev > 1) {
        CU(cudaEventRecord(D->ev_merge, D->streams[0]));
        D->merge_recorded = 1;
    }
fail:
    return rc;
}

int sblas_spmv_plan_execute_device(sblas_spmv_plan *P, double alpha, double beta, int sync)
{
    int rc = 0;
    for (int d = 0; d < P->ndev; ++d) if ((rc = enqueue_segments(P, d, alpha, beta)) != 0) return rc;
    if (!P->rank_mode)
        for (int d = 0; d < P->ndev; ++d) if ((rc = enqueue_merge(P, d, alpha, beta)) != 0) return rc;
    if (sync) {
        for (int d = 0; d < P->ndev; ++d) {
            if (P->devs[d].seg_begin < 0) continue;
            CU(cudaSetDevice(P->devs[d].device));
            CU(cudaStreamSynchronize(P->devs[d].streams[0]));
        }
    }
fail:
    return rc;
}

/* rank plans: finish the split rows this rank owns from a table that holds every
 * rank's edge partials (world x 2*max_local doubles, rank-major; e.g. the output
 * of an NCCL all-gather of each rank's edge table, or a symmetric-memory buffer) */
int sblas_spmv_plan_merge_gathered(sblas_spmv_plan *P, const double *gathered, double alpha, double beta)
{
    int rc = 0;
    sblas_dev *D = &P->devs[0];
    if (!P->rank_mode || D->seg_begin < 0 || D->nmerge == 0) return 0;
    CU(cudaSetDevice(D->device));
    if (gathered != P->gather_base) {
        for (int i = 0; i < D->nmsrc; ++i) D->h_msrc[i] = gathered + D->h_msrc_off[i];
        CU(cudaMemcpyAsync(D->d_msrc, D->h_msrc, (size_t)D->nmsrc * sizeof(double *), cudaMemcpyHostToDevice, D->streams[0]));
        CU(cudaStreamSynchronize(D->streams[0]));
        P->gather_base = gathered;
    }
    CU(sblas_launch_edge_merge(D->d_mrow, D->d_mbeg, (const double *const *)D->d_msrc, D->nmerge, D->d_y,
                               alpha, beta, D->streams[0]));
fail:
    return rc;
}

/* x (and y when beta != 0) from HOST memory to every GPU of the plan, asynchronously on
 * the plan's streams.  In-process multi-GPU: every GPU uploads its 1/nd slice of x over its
 * own PCIe link and the slices are exchanged over NVLink (replaces nd full-size H2D copies,
 * dspmv_mgpu_v1.cu:183). */
/* The part [*lo, *hi) of a peer's slice [slice_lo, slice_hi) of x that a GPU whose shard reads the columns
 * [win_lo, win_hi] has to pull (empty when *hi <= *lo).  Pure host arithmetic, tested on the CPU. */
void sblas_x_pull_range(long long slice_lo, long long slice_hi, long long win_lo, long long win_hi,
                        long long *lo, long long *hi)
{
    *lo = slice_lo > win_lo ? slice_lo : win_lo;
    *hi = slice_hi < win_hi + 1 ? slice_hi : win_hi + 1;
}

/* slice of x GPU number li of `live` uploads over its own PCIe link: [n*li/live, n*(li+1)/live) */
void sblas_x_slice(long long n, int li, int live, long long *lo, long long *hi)
{
    *lo = n * li / live;
    *hi = n * (li + 1) / live;
}

int sblas_spmv_plan_upload(sblas_spmv_plan *P, const double *x, const double *y)
{
    int rc = 0;
    const int nd = P->ndev;
    int live = 0;
    for (int d = 0; d < nd; ++d) if (P->devs[d].seg_begin >= 0) ++live;
    int li = 0;
    for (int d = 0; d < nd; ++d) {
        sblas_dev *D = &P->devs[d];
        if (D->seg_begin < 0) continue;
        CU(cudaSetDevice(D->device));
        cudaStream_t st = D->streams[0];
        long long lo, hi;
        sblas_x_slice(P->n, li, live, &lo, &hi);
        if (live == 1) { lo = D->col_lo; hi = (long long)D->col_hi + 1; }     /* only what the shard reads */
        D->xs_lo = lo; D->xs_hi = hi;
        if (x && live > 1)              /* peers may still be pulling the previous x out of this replica */
            for (int o = 0; o < nd; ++o)
                if (o != d && P->devs[o].xpull_recorded) CU(cudaStreamWaitEvent(st, P->devs[o].ev_xpull, 0));
        if (x && hi > lo)
            CU(cudaMemcpyAsync(D->d_x + lo, x + lo, (size_t)(hi - lo) * sizeof(double), cudaMemcpyHostToDevice, st));
        if (y)
            CU(cudaMemcpyAsync(D->d_y, y + D->first_row, (size_t)D->rows * sizeof(double), cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(D->ev_in, st));
        ++li;
    }
    if (x && live > 1) {
        for (int d = 0; d < nd; ++d) {
            sblas_dev *D = &P->devs[d];
            if (D->seg_begin < 0) continue;
            CU(cudaSetDevice(D->device));
            for (int o = 0; o < nd; ++o) {
                sblas_dev *O = &P->devs[o];
                if (o == d || O->seg_begin < 0 || O->xs_hi <= O->xs_lo) continue;
                /* pull O's slice once it has landed there -- only the part inside the window of columns D's
                 * shard reads (a banded matrix on 8 GPUs needs about an eighth of x per GPU) */
                long long lo, hi;
                sblas_x_pull_range(O->xs_lo, O->xs_hi, D->col_lo, D->col_hi, &lo, &hi);
                if (hi <= lo) continue;
                CU(cudaStreamWaitEvent(D->streams[0], O->ev_in, 0));
                CU(cudaMemcpyPeerAsync(D->d_x + lo, D->device, O->d_x + lo, O->device,
                                       (size_t)(hi - lo) * sizeof(double), D->streams[0]));
            }
            CU(cudaEventRecord(D->ev_xpull, D->streams[0]));     /* the next upload on every peer waits for this */
            D->xpull_recorded = 1;
        }
    }
fail:
    return rc;
}

/* the rows each GPU owns back to HOST y (disjoint ranges), then wait for every GPU */
int sblas_spmv_plan_download(sblas_spmv_plan *P, double *y)
{
    int rc = 0;
    const int nd = P->ndev;
    for (int d = 0; d < nd; ++d) {
        sblas_dev *D = &P->devs[d];
        if (D->seg_begin < 0) continue;
        CU(cudaSetDevice(D->device));
        const sblas_seg *S0 = &P->segs[D->seg_begin];
        int skip = 0;
        if (P->g_sf[S0->gidx]) {
            const int og = row_owner_seg(P, S0->gidx);
            const int od = P->rank_mode ? (P->g_owner[og] == P->rank ? 0 : -1) : P->g_owner[og];
            if (od != d) skip = 1;
        }
        if (y && D->rows - skip > 0)
            CU(cudaMemcpyAsync(y + D->first_row + skip, D->d_y + skip, (size_t)(D->rows - skip) * sizeof(double),
                               cudaMemcpyDeviceToHost, D->streams[0]));
    }
    for (int d = 0; d < nd; ++d) {
        sblas_dev *D = &P->devs[d];
        if (D->seg_begin < 0) continue;
        CU(cudaSetDevice(D->device));
        CU(cudaStreamSynchronize(D->streams[0]));
    }
fail:
    return rc;
}

/* rows of y a GPU owns (a split first row belongs to the GPU where the row starts): skip = 0 or 1 */
static int owned_skip(const sblas_spmv_plan *P, int d)
{
    const sblas_dev *D = &P->devs[d];
    const sblas_seg *S0 = &P->segs[D->seg_begin];
    if (!P->g_sf[S0->gidx]) return 0;
    const int og = row_owner_seg(P, S0->gidx);
    const int od = P->rank_mode ? (P->g_owner[og] == P->rank ? 0 : -1) : P->g_owner[og];
    return od != d;
}

/* Iterative use: x <- y on every GPU of the plan, entirely on the devices.
 * Every GPU pulls the rows each GPU owns out of that GPU's y slice into its own replica of x --
 * an all-gather over NVLink (cudaMemcpyPeerAsync), ordered by events only; the host does not
 * wait.  Needs a square matrix; plans that hold one shard of a multi-process job (world > 1)
 * gather through the caller's collective instead. */
int sblas_spmv_plan_chain(sblas_spmv_plan *P)
{
    int rc = 0;
    if (!P->dry && P->m == P->n && P->rank_mode && P->world > 1 && P->peer_x_bound) {
        /* one process per GPU: all-gather of the owned y rows into every rank's x with P2P stores + flags */
        sblas_dev *D = &P->devs[0];
        CU(cudaSetDevice(D->device));
        int skip = 0;
        long long cnt = 0;
        if (D->seg_begin >= 0) { skip = owned_skip(P, 0); cnt = (long long)D->rows - skip; }
        CU(sblas_launch_chain_gather(D->seg_begin >= 0 ? D->d_y + skip : NULL, cnt > 0 ? cnt : 0,
                                     D->seg_begin >= 0 ? (long long)D->first_row + skip : 0, (void *const *)P->d_peer_x,
                                     (void *const *)P->d_peer_xflags, P->world, P->rank, P->d_chain_ctr,
                                     D->streams ? D->streams[0] : 0));
        return 0;
    }
    if (P->dry || P->m != P->n || (P->rank_mode && P->world > 1)) {
        sblas_set_error("%s%s (line %d)", "chain needs a square matrix and every shard in this process (or bound peer x buffers)", "", __LINE__);
        return -1;
    }
    const int nd = P->ndev;
    for (int d = 0; d < nd; ++d) {
        sblas_dev *D = &P->devs[d];
        if (D->seg_begin < 0) continue;
        CU(cudaSetDevice(D->device));
        CU(cudaEventRecord(D->ev_y, D->streams[0]));             /* everything enqueued so far: y is final */
    }
    for (int d = 0; d < nd; ++d) {
        sblas_dev *D = &P->devs[d];
        if (D->seg_begin < 0) continue;
        CU(cudaSetDevice(D->device));
        for (int o = 0; o < nd; ++o) {
            sblas_dev *O = &P->devs[o];
            if (O->seg_begin < 0) continue;
            const int skip = owned_skip(P, o);
            if (O->rows - skip <= 0) continue;
            if (o != d) CU(cudaStreamWaitEvent(D->streams[0], O->ev_y, 0));
            CU(cudaMemcpyPeerAsync(D->d_x + O->first_row + skip, D->device, O->d_y + skip, O->device,
                                   (size_t)(O->rows - skip) * sizeof(double), D->streams[0]));
        }
        CU(cudaEventRecord(D->ev_chain, D->streams[0]));
    }
    /* nobody overwrites its y (next product) before every GPU has pulled it */
    for (int o = 0; o < nd; ++o) {
        sblas_dev *O = &P->devs[o];
        if (O->seg_begin < 0) continue;
        CU(cudaSetDevice(O->device));
        for (int d = 0; d < nd; ++d)
            if (d != o && P->devs[d].seg_begin >= 0) CU(cudaStreamWaitEvent(O->streams[0], P->devs[d].ev_chain, 0));
    }
fail:
    return rc;
}

/* One GPU, one segment (a one-GPU plan, or one rank of a multi-process job): the product is cut into
 * PIECES -- its row panels, the row-aligned ones (short / medium / long-medium kernels) cut further
 * into runs of ~128 Ki rows -- and the y slice of piece p+1 goes up and the finished slice of piece p-1
 * comes down on a second stream while piece p computes; only the first piece's upload and the last
 * piece's download are exposed.  (At N GPUs the reference's nnz-balanced split leaves most ROWS on
 * the last shard -- 875 k of g1m's 1 M rows at N = 8 -- so without this the last rank's 7 MB of y each
 * way over PCIe would double the product time.)  The split-row exchange of a rank plan runs after the
 * last piece; the row it finishes is that piece's last row. */
#define SBLAS_PIECE_ROWS (128 * 1024)
#define SBLAS_PIECE_MAX 16
static int execute_pipelined(sblas_spmv_plan *P, double alpha, const double *x, double beta, double *y)
{
    int rc = 0;
    sblas_dev *D = &P->devs[0];
    sblas_seg *S = &P->segs[D->seg_begin];
    const int nu = S->unit_end - S->unit_begin;
    CU(cudaSetDevice(D->device));
    if (!D->copy_stream) CU(cudaStreamCreateWithFlags(&D->copy_stream, cudaStreamNonBlocking));
    /* count the pieces */
    int npiece = 0;
    for (int u = 0; u < nu; ++u) {
        const sblas_unit *U = &P->units[S->unit_begin + u];
        const long long rows = (long long)U->args.row_hi - U->args.row_lo + 1;
        int k = 1;
        if ((U->kind == SBLAS_K_SHORT || U->kind == SBLAS_K_ROWTILE || U->kind == SBLAS_K_ROWSPLIT) &&
            rows >= 2 * SBLAS_PIECE_ROWS) {
            k = (int)(rows / SBLAS_PIECE_ROWS);
            if (k > SBLAS_PIECE_MAX) k = SBLAS_PIECE_MAX;
        }
        npiece += k;
    }
    if (D->nev_unit < 2 * npiece) {
        cudaEvent_t *ev = (cudaEvent_t *)calloc((size_t)2 * npiece, sizeof(cudaEvent_t));
        if (!ev) return 1;
        for (int i = 0; i < D->nev_unit; ++i) ev[i] = D->ev_unit[i];
        free(D->ev_unit);
        D->ev_unit = ev;
        for (int i = D->nev_unit; i < 2 * npiece; ++i) CU(cudaEventCreateWithFlags(&D->ev_unit[i], cudaEventDisableTiming));
        D->nev_unit = 2 * npiece;
    }
    if (P->cap_pieces < npiece) {
        free(P->piece_lo); free(P->piece_hi); free(P->piece_unit);
        P->piece_lo = (int *)malloc((size_t)npiece * sizeof(int));
        P->piece_hi = (int *)malloc((size_t)npiece * sizeof(int));
        P->piece_unit = (int *)malloc((size_t)npiece * sizeof(int));
        if (!P->piece_lo || !P->piece_hi || !P->piece_unit) { P->cap_pieces = 0; return 1; }
        P->cap_pieces = npiece;
    }
    int np = 0;
    for (int u = 0; u < nu; ++u) {
        const sblas_unit *U = &P->units[S->unit_begin + u];
        const long long rows = (long long)U->args.row_hi - U->args.row_lo + 1;
        int k = 1;
        if ((U->kind == SBLAS_K_SHORT || U->kind == SBLAS_K_ROWTILE || U->kind == SBLAS_K_ROWSPLIT) &&
            rows >= 2 * SBLAS_PIECE_ROWS) {
            k = (int)(rows / SBLAS_PIECE_ROWS);
            if (k > SBLAS_PIECE_MAX) k = SBLAS_PIECE_MAX;
        }
        for (int i = 0; i < k; ++i) {
            /* cuts on multiples of 64 rows: whole tiles of the row-tile kernel for every R */
            long long lo = U->args.row_lo + (rows * i / k) / 64 * 64, hi = U->args.row_lo + (rows * (i + 1) / k) / 64 * 64 - 1;
            if (i == 0) lo = U->args.row_lo;
            if (i == k - 1) hi = U->args.row_hi;
            P->piece_lo[np] = (int)lo; P->piece_hi[np] = (int)hi; P->piece_unit[np] = S->unit_begin + u;
            ++np;
        }
    }
    cudaStream_t st = D->streams[0], cs = D->copy_stream;
    const int skip = owned_skip(P, 0);                  /* a split first row belongs to the previous rank */
    const int exchange = P->rank_mode && P->world > 1 && P->peer_bound;
    if (x && D->col_hi >= D->col_lo)
        CU(cudaMemcpyAsync(D->d_x + D->col_lo, x + D->col_lo, (size_t)(D->col_hi - D->col_lo + 1) * sizeof(double),
                           cudaMemcpyHostToDevice, st));
    /* uploads: piece 0 on the compute stream, the others on the copy stream */
    for (int p = 0; p < np && beta != 0.0; ++p) {
        const long long rows = (long long)P->piece_hi[p] - P->piece_lo[p] + 1;
        if (rows <= 0) continue;
        CU(cudaMemcpyAsync(D->d_y + P->piece_lo[p], y + D->first_row + P->piece_lo[p], (size_t)rows * sizeof(double),
                           cudaMemcpyHostToDevice, p == 0 ? st : cs));
        if (p > 0) CU(cudaEventRecord(D->ev_unit[2 * p], cs));
    }
    for (int p = 0; p < np; ++p) {
        sblas_unit *U = &P->units[P->piece_unit[p]];
        const long long rows = (long long)P->piece_hi[p] - P->piece_lo[p] + 1;
        if (p > 0 && beta != 0.0 && rows > 0) CU(cudaStreamWaitEvent(st, D->ev_unit[2 * p], 0));
        U->args.alpha = alpha; U->args.beta = beta;
        U->args.edge = S->args.edge;
        sblas_seg_args a = U->args;
        a.row_lo = P->piece_lo[p]; a.row_hi = P->piece_hi[p];
        CU(sblas_launch_spmv_segment(&a, U->kind, U->ipt, 0, st));
        if (rows <= 0) continue;
        long long lo = P->piece_lo[p], cnt = rows;
        if (lo == 0 && skip) { lo = 1; cnt -= 1; }
        if (p + 1 < np) {                          /* comes down while the next piece computes */
            CU(cudaEventRecord(D->ev_unit[2 * p + 1], st));
            CU(cudaStreamWaitEvent(cs, D->ev_unit[2 * p + 1], 0));
            if (cnt > 0)
                CU(cudaMemcpyAsync(y + D->first_row + lo, D->d_y + lo, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, cs));
        } else {
            if (exchange) CU(sblas_spmv_plan_exchange_merge(P, alpha, beta));
            else if (D->nmerge > 0) CU(enqueue_merge(P, 0, alpha, beta));
            if (cnt > 0)
                CU(cudaMemcpyAsync(y + D->first_row + lo, D->d_y + lo, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, st));
        }
    }
    CU(cudaEventRecord(D->ev_done, st));
    CU(cudaStreamSynchronize(cs));
    CU(cudaStreamSynchronize(st));
fail:
    return rc;
}

int sblas_spmv_plan_execute(sblas_spmv_plan *P, const double *alpha, const double *x, const double *beta, double *y)
{
    int rc;
    if (!P->dry && P->ndev == 1 && P->nseg == 1 && P->devs[0].seg_begin >= 0 && P->devs[0].nstreams == 1 && x && y &&
        (!P->rank_mode || P->world == 1 || P->peer_bound) && (P->rank_mode || P->devs[0].nmerge == 0) &&
        env_int("SBLAS_PIPELINE", 1))
        return execute_pipelined(P, *alpha, x, *beta, y);
    if ((rc = sblas_spmv_plan_upload(P, x, *beta != 0.0 ? y : NULL)) != 0) return rc;
    if ((rc = sblas_spmv_plan_execute_device(P, *alpha, *beta, 0)) != 0) return rc;
    /* one rank of a multi-process job with bound peer tables: the split-row exchange is part of the
     * product (P2P publish + merge kernels on the same stream); every rank must make this call */
    if (P->rank_mode && P->world > 1 && P->peer_bound &&
        (rc = sblas_spmv_plan_exchange_merge(P, *alpha, *beta)) != 0) return rc;
    return sblas_spmv_plan_download(P, y);
}

/* ------------------------------------------------------------------ accessors */
int sblas_spmv_plan_local_segments(const sblas_spmv_plan *P) { return P->nseg; }

int sblas_spmv_plan_local_segment(const sblas_spmv_plan *P, int i, long long out[10])
{
    if (i < 0 || i >= P->nseg) return -1;
    const sblas_seg *S = &P->segs[i];
    const sblas_dev *D = &P->devs[S->dev];
    out[0] = S->gidx;
    out[1] = P->g_lo[S->gidx];
    out[2] = P->g_hi[S->gidx];
    out[3] = D->first_idx + S->args.nz0;
    out[4] = D->first_idx + S->args.nz1;
    out[5] = P->g_sf[S->gidx];
    out[6] = P->g_sl[S->gidx];
    out[7] = 2LL * S->lidx;
    out[8] = S->dev;
    out[9] = D->first_row;
    return 0;
}

int sblas_spmv_plan_merge_list(const sblas_spmv_plan *P, int dev, int *nmerge, const int **mrow, const int **mbeg,
                               const long long **msrc_off)
{
    if (dev < 0 || dev >= P->ndev) return -1;
    const sblas_dev *D = &P->devs[dev];
    *nmerge = D->nmerge;
    *mrow = D->h_mrow; *mbeg = D->h_mbeg; *msrc_off = D->h_msrc_off;
    return 0;
}

int sblas_spmv_plan_num_devices(const sblas_spmv_plan *P) { return P->ndev; }
int sblas_spmv_plan_num_segments(const sblas_spmv_plan *P) { return P->nparts; }
int sblas_spmv_plan_segment(const sblas_spmv_plan *P, int seg, sblas_part *out, int *device)
{
    if (seg < 0 || seg >= P->nparts) return -1;
    if (out) *out = P->parts[seg];
    if (device) *device = P->g_owner[seg];
    return 0;
}
double *sblas_spmv_plan_x(sblas_spmv_plan *P, int dev) { return P->devs[dev].d_x; }
int sblas_spmv_plan_x_window(const sblas_spmv_plan *P, int dev, long long *first_col, long long *last_col)
{
    if (dev < 0 || dev >= P->ndev) return -1;
    if (first_col) *first_col = P->devs[dev].col_lo;
    if (last_col) *last_col = P->devs[dev].col_hi;
    return 0;
}
double *sblas_spmv_plan_y(sblas_spmv_plan *P, int dev, int *first_row, int *rows)
{
    if (first_row) *first_row = P->devs[dev].first_row;
    if (rows) *rows = P->devs[dev].rows;
    return P->devs[dev].d_y;
}
const int *sblas_spmv_plan_rowptr(sblas_spmv_plan *P, int dev, int *count)
{
    if (count) *count = P->devs[dev].rows + 1;
    return P->devs[dev].d_rowptr;
}
void *sblas_spmv_plan_stream(sblas_spmv_plan *P, int dev) { return P->devs[dev].streams ? (void *)P->devs[dev].streams[0] : NULL; }
double *sblas_spmv_plan_edge_ptr(sblas_spmv_plan *P, int dev) { return P->devs[dev].d_edge; }
int sblas_spmv_plan_edge_slots(const sblas_spmv_plan *P) { return 2 * P->max_local; }

/* ---- fused exchange over peer-mapped memory (one process per GPU) */
int sblas_spmv_plan_bind_peer_tables(sblas_spmv_plan *P, void *const *peer_bases, long long table_words)
{
    int rc = 0;
    if (!P->rank_mode || P->dry || !peer_bases) return -1;
    sblas_dev *D = &P->devs[0];
    const int W = P->world, slots = 2 * P->max_local;
    if (table_words < (long long)W * slots) return -1;
    int *out_slot = (int *)calloc((size_t)P->nseg + 1, sizeof(int));
    int *out_owner = (int *)calloc((size_t)P->nseg + 1, sizeof(int));
    long long *out_off = (long long *)calloc((size_t)P->nseg + 1, sizeof(long long));
    int *owners = (int *)calloc((size_t)W + 1, sizeof(int));
    int *contrib = (int *)calloc((size_t)W + 1, sizeof(int));
    char *seen = (char *)calloc((size_t)W + 1, 1);
    int nout = 0, nown = 0, ncon = 0;
    for (int s = 0; s < P->nseg; ++s) {
        const sblas_seg *S = &P->segs[s];
        if (!P->g_sf[S->gidx]) continue;
        const int orank = P->g_owner[row_owner_seg(P, S->gidx)];
        if (orank == P->rank) continue;
        out_slot[nout] = 2 * S->lidx;
        out_owner[nout] = orank;
        out_off[nout] = (long long)P->rank * slots + 2 * S->lidx;
        ++nout;
        if (!seen[orank]) { seen[orank] = 1; owners[nown++] = orank; }
    }
    memset(seen, 0, (size_t)W + 1);
    for (int i = 0; i < D->nmsrc; ++i) {
        const int r = (int)(D->h_msrc_off[i] / slots);
        if (r != P->rank && !seen[r]) { seen[r] = 1; contrib[ncon++] = r; }
    }
    if (D->seg_begin >= 0 || 1) {
        CU(cudaSetDevice(D->device >= 0 ? D->device : 0));
        CU(sblas_preload_exchange_kernels());
        CU(cudaMalloc((void **)&P->d_peer_bases, (size_t)W * sizeof(void *)));
        CU(cudaMemcpy(P->d_peer_bases, peer_bases, (size_t)W * sizeof(void *), cudaMemcpyHostToDevice));
        CU(cudaMalloc((void **)&P->d_out_slot, (size_t)(nout + 1) * sizeof(int)));
        CU(cudaMalloc((void **)&P->d_out_owner, (size_t)(nout + 1) * sizeof(int)));
        CU(cudaMalloc((void **)&P->d_out_off, (size_t)(nout + 1) * sizeof(long long)));
        CU(cudaMalloc((void **)&P->d_owners, (size_t)(nown + 1) * sizeof(int)));
        CU(cudaMalloc((void **)&P->d_contrib, (size_t)(ncon + 1) * sizeof(int)));
        CU(cudaMalloc((void **)&P->d_msrc_off, (size_t)(D->nmsrc + 1) * sizeof(long long)));
        CU(cudaMalloc((void **)&P->d_epoch, sizeof(unsigned long long)));
        CU(cudaMemset(P->d_epoch, 0, sizeof(unsigned long long)));
        CU(cudaMemcpy(P->d_out_slot, out_slot, (size_t)nout * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(P->d_out_owner, out_owner, (size_t)nout * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(P->d_out_off, out_off, (size_t)nout * sizeof(long long), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(P->d_owners, owners, (size_t)nown * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(P->d_contrib, contrib, (size_t)ncon * sizeof(int), cudaMemcpyHostToDevice));
        if (D->nmsrc > 0)
            CU(cudaMemcpy(P->d_msrc_off, D->h_msrc_off, (size_t)D->nmsrc * sizeof(long long), cudaMemcpyHostToDevice));
    }
    P->nout = nout; P->nowners = nown; P->ncontrib = ncon;
    P->my_base = (double *)peer_bases[P->rank];
    P->table_words = table_words;
    P->peer_bound = 1;
    P->graph_valid = 0;
fail:
    free(out_slot); free(out_owner); free(out_off); free(owners); free(contrib); free(seen);
    return rc;
}

/* publish this rank's split-row partials into their owners' tables and finish the rows this
 * rank owns; both enqueued on the plan's stream, nothing waits on the host */
int sblas_spmv_plan_exchange_merge(sblas_spmv_plan *P, double alpha, double beta)
{
    return sblas_spmv_plan_exchange_merge_phase(P, alpha, beta, 0);
}

/* phase 0: publish + merge (normal use); 1: publish only; 2: merge only.  The split form lets
 * several rank plans that share ONE GPU (tests) be stepped without kernels waiting on each other. */
int sblas_spmv_plan_exchange_merge_phase(sblas_spmv_plan *P, double alpha, double beta, int phase)
{
    int rc = 0;
    if (!P->peer_bound) return -1;
    sblas_dev *D = &P->devs[0];
    CU(cudaSetDevice(D->device));
    const int slots = 2 * P->max_local;
    cudaStream_t st = D->streams ? D->streams[0] : 0;
    /* publish always runs: it advances the product counter the merge reads */
    if (phase != 2)
        CU(sblas_launch_edge_publish(D->d_edge, slots, P->d_out_slot, P->d_out_owner, P->d_out_off, P->nout, P->d_owners,
                                     P->nowners, (void *const *)P->d_peer_bases, P->table_words, P->world, P->rank,
                                     P->d_epoch, st));
    if (D->seg_begin >= 0 && phase != 1)
        CU(sblas_launch_edge_merge_wait(D->d_mrow, D->d_mbeg, P->d_msrc_off, D->nmerge, D->d_y, alpha, beta,
                                        P->d_contrib, P->ncontrib, (void *const *)P->d_peer_bases, P->table_words,
                                        P->world, P->rank, P->d_epoch, st));
fail:
    return rc;
}

/* One product on a resident plan, x and y on the device: the segments' kernels and, for a rank plan
 * with bound peer tables, the split-row exchange.  On single-GPU plans (one rank of a multi-process job
 * or a one-GPU plan) the launch sequence is captured into a CUDA graph the second time it runs with the
 * same alpha / beta and REPLAYED from then on: one cudaGraphLaunch per product.
 * SBLAS_GRAPH=0 keeps plain stream launches. */
int sblas_spmv_plan_step(sblas_spmv_plan *P, double alpha, double beta)
{
    int rc = 0;
    const int exchange = P->rank_mode && P->world > 1 && P->peer_bound;
    const int graphable = !P->dry && !P->x_policy && P->ndev == 1 && P->devs[0].seg_begin >= 0 && P->devs[0].nstreams == 1 &&
                          (P->rank_mode ? (P->world == 1 || exchange) : P->devs[0].nmerge >= 0) && env_int("SBLAS_GRAPH", 1);
    if (!graphable) {
        if ((rc = sblas_spmv_plan_execute_device(P, alpha, beta, 0)) != 0) return rc;
        if (exchange) rc = sblas_spmv_plan_exchange_merge(P, alpha, beta);
        return rc;
    }
    sblas_dev *D = &P->devs[0];
    cudaStream_t st = D->streams[0];
    CU(cudaSetDevice(D->device));
    if (P->graph_valid && P->graph_alpha == alpha && P->graph_beta == beta) {
        CU(cudaGraphLaunch((cudaGraphExec_t)P->graph_exec, st));
        return 0;
    }
    if (!P->graph_warm) {               /* first product: plain launches (one-time attribute set-up happens here) */
        P->graph_warm = 1;
        if ((rc = sblas_spmv_plan_execute_device(P, alpha, beta, 0)) != 0) return rc;
        if (exchange) rc = sblas_spmv_plan_exchange_merge(P, alpha, beta);
        return rc;
    }
    if (P->graph_exec) { cudaGraphExecDestroy((cudaGraphExec_t)P->graph_exec); P->graph_exec = NULL; P->graph_valid = 0; }
    {
        cudaGraph_t g = NULL;
        CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        P->capturing = 1;
        rc = sblas_spmv_plan_execute_device(P, alpha, beta, 0);
        if (rc == 0 && exchange) rc = sblas_spmv_plan_exchange_merge(P, alpha, beta);
        P->capturing = 0;
        cudaError_t e = cudaStreamEndCapture(st, &g);
        if (rc != 0 || e != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            if (rc == 0) { sblas_set_error("%s%s (line %d)", "graph capture failed: ", cudaGetErrorString(e), __LINE__); rc = 1; }
            return rc;
        }
        cudaGraphExec_t ge = NULL;
        e = cudaGraphInstantiate(&ge, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { sblas_set_error("%s%s (line %d)", "cudaGraphInstantiate: ", cudaGetErrorString(e), __LINE__); return 1; }
        P->graph_exec = (void *)ge;
        P->graph_alpha = alpha; P->graph_beta = beta; P->graph_valid = 1;
        CU(cudaGraphLaunch(ge, st));
    }
fail:
    return rc;
}

/* Rank plans, iterative use: make every rank's x a peer-mapped buffer (n doubles; e.g. symmetric memory) so that
 * sblas_spmv_plan_chain can all-gather y into it over NVLink.  peer_x[r] / peer_flags[r] (2*world 8-byte words,
 * zeroed) = rank r's buffers as mapped in THIS process.  The plan computes on peer_x[rank] from now on. */
int sblas_spmv_plan_bind_peer_x(sblas_spmv_plan *P, void *const *peer_x, void *const *peer_flags)
{
    int rc = 0;
    if (!P->rank_mode || P->dry || !peer_x || !peer_flags) return -1;
    sblas_dev *D = &P->devs[0];
    CU(cudaSetDevice(D->device >= 0 ? D->device : 0));
    CU(sblas_preload_exchange_kernels());
    if (!P->d_peer_x) {
        CU(cudaMalloc((void **)&P->d_peer_x, (size_t)P->world * sizeof(void *)));
        CU(cudaMalloc((void **)&P->d_peer_xflags, (size_t)P->world * sizeof(void *)));
        CU(cudaMalloc((void **)&P->d_chain_ctr, sizeof(unsigned long long)));
        CU(cudaMemset(P->d_chain_ctr, 0, sizeof(unsigned long long)));
    }
    CU(cudaMemcpy(P->d_peer_x, peer_x, (size_t)P->world * sizeof(void *), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(P->d_peer_xflags, peer_flags, (size_t)P->world * sizeof(void *), cudaMemcpyHostToDevice));
    if (D->seg_begin >= 0) {
        if (!P->x_alloc) P->x_alloc = D->d_x;                 /* still freed with the plan */
        D->d_x = (double *)peer_x[P->rank];
        for (int s = 0; s < P->nseg; ++s) P->segs[s].args.x = D->d_x;
        for (int u = 0; u < P->nunits; ++u) P->units[u].args.x = D->d_x;
        apply_x_window_policy(P, D);
    }
    P->peer_x_bound = 1;
    P->graph_valid = 0;
fail:
    return rc;
}

int sblas_spmv_plan_bind_edge_table(sblas_spmv_plan *P, double *device_block)
{
    if (!P->rank_mode || !device_block) return -1;
    sblas_dev *D = &P->devs[0];
    if (D->seg_begin < 0) return 0;
    cudaSetDevice(D->device);
    D->d_edge = device_block;                 /* the plan's own table stays inside its allocation, unused */
    D->edge_bound = 1;
    for (int s = D->seg_begin; s < D->seg_end; ++s) P->segs[s].args.edge = D->d_edge + 2 * P->segs[s].lidx;
    return 0;
}

int sblas_memcpy(void *dst, const void *src, unsigned long long bytes, int kind)
{
    return (int)cudaMemcpy(dst, src, (size_t)bytes, (enum cudaMemcpyKind)kind);
}
int sblas_device_synchronize(void) { return (int)cudaDeviceSynchronize(); }

int sblas_spmv_plan_edges(sblas_spmv_plan *P, double *out)
{
    int rc = 0;
    for (int t = 0; t < P->nparts; ++t) { out[2 * t] = 0.0; out[2 * t + 1] = 0.0; }
    for (int s = 0; s < P->nseg; ++s) {
        sblas_seg *S = &P->segs[s];
        sblas_dev *D = &P->devs[S->dev];
        CU(cudaSetDevice(D->device));
        CU(cudaMemcpy(out + 2 * S->gidx, D->d_edge + 2 * S->lidx, 2 * sizeof(double), cudaMemcpyDeviceToHost));
    }
fail:
    return rc;
}

double sblas_spmv_plan_alg_bytes(const sblas_spmv_plan *P, int beta_nonzero, long long x_touched_per_gpu)
{
    double b = 0.0;
    for (int d = 0; d < P->ndev; ++d) {
        const sblas_dev *D = &P->devs[d];
        if (D->seg_begin < 0) continue;
        const double xt = x_touched_per_gpu >= 0 ? (double)x_touched_per_gpu : (double)P->n;
        b += 12.0 * D->nnz + 4.0 * (D->rows + 1) + 8.0 * xt + 8.0 * D->rows * (beta_nonzero ? 2.0 : 1.0);
    }
    return b;
}

int sblas_spmv_plan_num_units(const sblas_spmv_plan *P) { return P->nunits; }

static const sblas_seg *unit_segment(const sblas_spmv_plan *P, int i, int *sidx)
{
    for (int s = 0; s < P->nseg; ++s)
        if (i >= P->segs[s].unit_begin && i < P->segs[s].unit_end) { if (sidx) *sidx = s; return &P->segs[s]; }
    return NULL;
}

int sblas_spmv_plan_unit(const sblas_spmv_plan *P, int i, long long out[8])
{
    int sidx = 0;
    if (i < 0 || i >= P->nunits) return -1;
    const sblas_seg *S = unit_segment(P, i, &sidx);
    if (!S) return -1;
    const sblas_unit *U = &P->units[i];
    const sblas_dev *D = &P->devs[S->dev];
    out[0] = sidx; out[1] = U->kind; out[2] = U->ipt;
    out[3] = (long long)D->first_row + U->args.row_lo;
    out[4] = (long long)D->first_row + U->args.row_hi;
    out[5] = D->first_idx + U->args.nz0;
    out[6] = D->first_idx + U->args.nz1;
    out[7] = ((U->kind == SBLAS_K_TMA || U->kind == SBLAS_K_TILE) && U->args.ntile > 0) ? 2 : 1;
    return 0;
}

int sblas_spmv_plan_execute_unit(sblas_spmv_plan *P, int i, double alpha, double beta)
{
    int rc = 0;
    if (i < 0 || i >= P->nunits || P->dry) return -1;
    const sblas_seg *S = unit_segment(P, i, NULL);
    if (!S) return -1;
    sblas_unit *U = &P->units[i];
    sblas_dev *D = &P->devs[S->dev];
    CU(cudaSetDevice(D->device));
    U->args.alpha = alpha; U->args.beta = beta;
    U->args.edge = S->args.edge;
    CU(sblas_launch_spmv_segment(&U->args, U->kind, U->ipt, 0, D->streams[S->stream]));
fail:
    return rc;
}

int sblas_spmv_plan_launches(const sblas_spmv_plan *P)
{
    int n = 0;
    for (int ui = 0; ui < P->nunits; ++ui) {
        const sblas_unit *U = &P->units[ui];
        if (U->args.row_hi < U->args.row_lo) continue;
        n += ((U->kind == SBLAS_K_TMA || U->kind == SBLAS_K_TILE) && U->args.ntile > 0) ? 2 : 1;
    }
    for (int d = 0; d < P->ndev; ++d) if (P->devs[d].nmerge > 0) ++n;
    return n;
}
