/* sblas_api.c -- the one-shot entry points: host pointers in, host y out.
 * plan_create + execute + destroy, with the reference's return codes
 * (see include/sblas_spmv.h).  Also the plain-C definitions of the reference's own
 * names declared in include/spmv_kernel.h. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sblas_internal.h"
#include "sblas_spmv.h"
#include "spmv_kernel.h"

/* ---- optional plan cache (SBLAS_PLAN_CACHE=1): the harness calls the one-shot entry points
 * again and again with the same host arrays; with the cache the matrix is uploaded once and
 * later calls only move x and y.  Opt-in, because the reference semantics re-read the host
 * arrays on every call: a caller that edits csrVal in place must not enable it (or must call
 * sblas_spmv_cache_clear).  Keyed on the array addresses, sizes and parameters, and guarded by
 * a fingerprint of the first/last entries of every array. */
#define SBLAS_CACHE_SLOTS 8
typedef struct cache_ent {
    sblas_spmv_plan *plan;
    int version, m, n, ngpu, kernel, q;
    long long nnz, nb;
    const void *val, *rp, *col;
    unsigned long long finger;
    unsigned long long stamp;
} cache_ent;
static cache_ent g_cache[SBLAS_CACHE_SLOTS];
static unsigned long long g_stamp;

static unsigned long long fingerprint(int m, long long nnz, const double *val, const long long *rp, const int *col)
{
    unsigned long long h = 1469598103934665603ull;
    const long long k = nnz < 64 ? nnz : 64, kr = m + 1 < 64 ? m + 1 : 64;
#define MIX(p, bytes) do { const unsigned char *b_ = (const unsigned char *)(p); \
        for (long long i_ = 0; i_ < (long long)(bytes); ++i_) { h ^= b_[i_]; h *= 1099511628211ull; } } while (0)
    MIX(val, k * 8); MIX(val + (nnz - k), k * 8);
    MIX(col, k * 4); MIX(col + (nnz - k), k * 4);
    MIX(rp, kr * 8); MIX(rp + (m + 1 - kr), kr * 8);
    if (nnz > 128) { MIX(val + nnz / 2, 64); MIX(col + nnz / 2, 32); }
#undef MIX
    return h;
}

void sblas_spmv_cache_clear(void)
{
    for (int i = 0; i < SBLAS_CACHE_SLOTS; ++i) {
        if (g_cache[i].plan) sblas_spmv_plan_destroy(g_cache[i].plan);
        memset(&g_cache[i], 0, sizeof g_cache[i]);
    }
    sblas_pool_release();
}

static int one_shot(int version, int m, int n, long long nnz, double *alpha, double *val, long long *rp,
                    int *col, double *x, double *beta, double *y, int ngpu, int kernel, long long nb, int q)
{
    const char *ce = getenv("SBLAS_PLAN_CACHE");
    if (ce && *ce && *ce != '0' && m > 0 && nnz > 0) {
        const unsigned long long fp = fingerprint(m, nnz, val, rp, col);
        int slot = -1, victim = 0;
        for (int i = 0; i < SBLAS_CACHE_SLOTS; ++i) {
            const cache_ent *e = &g_cache[i];
            if (e->plan && e->version == version && e->m == m && e->n == n && e->nnz == nnz && e->ngpu == ngpu &&
                e->kernel == kernel && e->nb == nb && e->q == q && e->val == val && e->rp == rp && e->col == col &&
                e->finger == fp) { slot = i; break; }
            if (g_cache[i].stamp < g_cache[victim].stamp) victim = i;
        }
        if (slot < 0) {
            cache_ent *e = &g_cache[victim];
            if (e->plan) sblas_spmv_plan_destroy(e->plan);
            memset(e, 0, sizeof *e);
            int rc = sblas_spmv_plan_create(&e->plan, version, m, n, nnz, val, rp, col, ngpu, kernel, nb, q);
            if (rc != 0) { e->plan = NULL; return rc; }
            e->version = version; e->m = m; e->n = n; e->nnz = nnz; e->ngpu = ngpu; e->kernel = kernel;
            e->nb = nb; e->q = q; e->val = val; e->rp = rp; e->col = col; e->finger = fp;
            slot = victim;
        }
        g_cache[slot].stamp = ++g_stamp;
        return sblas_spmv_plan_execute(g_cache[slot].plan, alpha, x, beta, y) == 0 ? 0 : -1;
    }
    sblas_spmv_plan *plan = NULL;
    const int timing = getenv("SBLAS_TIMING") != NULL;
    const double t0 = timing ? sblas_get_time() : 0.0;
    const char *pe = getenv("SBLAS_POOL");
    const int pooled = (pe && *pe == '0') ? 0 : SBLAS_CREATE_POOLED;
    int rc = sblas_spmv_plan_create_flags(&plan, version, m, n, nnz, val, rp, col, ngpu, kernel, nb, q, pooled);
    if (rc != 0) {
        if (getenv("SBLAS_VERBOSE")) fprintf(stderr, "sblas: %s\n", sblas_last_error());
        return rc;
    }
    const double t1 = timing ? sblas_get_time() : 0.0;
    rc = sblas_spmv_plan_execute(plan, alpha, x, beta, y);
    if (timing) {
        const double t2 = sblas_get_time();
        fprintf(stderr, "sblas one-shot version %d ngpu %d: plan %.3f ms (%d segments, %d launches), execute %.3f ms\n", version,
                ngpu, (t1 - t0) * 1e3, sblas_spmv_plan_num_segments(plan), sblas_spmv_plan_launches(plan), (t2 - t1) * 1e3);
    }
    if (rc != 0) {
        if (getenv("SBLAS_VERBOSE")) fprintf(stderr, "sblas: %s\n", sblas_last_error());
        rc = -1;                                   /* kernel / copy failure: dspmv_mgpu_v1.cu:226-228 */
    }
    sblas_spmv_plan_destroy(plan);
    return rc;
}

int sblas_spmv_mgpu_baseline(int m, int n, long long nnz, double *alpha, double *csrVal, long long *csrRowPtr,
                             int *csrColIndex, double *x, double *beta, double *y, int ngpu)
{
    return one_shot(SBLAS_BASELINE, m, n, nnz, alpha, csrVal, csrRowPtr, csrColIndex, x, beta, y, ngpu, 2, 0, 1);
}

int sblas_spmv_mgpu_v1(int m, int n, long long nnz, double *alpha, double *csrVal, long long *csrRowPtr,
                       int *csrColIndex, double *x, double *beta, double *y, int ngpu, int kernel)
{
    return one_shot(SBLAS_V1, m, n, nnz, alpha, csrVal, csrRowPtr, csrColIndex, x, beta, y, ngpu, kernel, 0, 1);
}

int sblas_spmv_mgpu_v2(int m, int n, long long nnz, double *alpha, double *csrVal, long long *csrRowPtr,
                       int *csrColIndex, double *x, double *beta, double *y, int ngpu, int kernel, long long nb,
                       int copy_of_workspace)
{
    if (nb <= 0 || ngpu == 0 || copy_of_workspace == 0) return -1;      /* dspmv_mgpu_v2.cu:44-46 */
    return one_shot(SBLAS_V2, m, n, nnz, alpha, csrVal, csrRowPtr, csrColIndex, x, beta, y, ngpu, kernel, nb,
                    copy_of_workspace);
}

/* ---- the reference's names, C linkage (include/spmv_kernel.h) */
int spMV_mgpu_baseline(int m, int n, long long nnz, double *alpha, double *csrVal, long long *csrRowPtr,
                       int *csrColIndex, double *x, double *beta, double *y, int ngpu)
{
    return sblas_spmv_mgpu_baseline(m, n, nnz, alpha, csrVal, csrRowPtr, csrColIndex, x, beta, y, ngpu);
}
int spMV_mgpu_v1(int m, int n, long long nnz, double *alpha, double *csrVal, long long *csrRowPtr,
                 int *csrColIndex, double *x, double *beta, double *y, int ngpu, int kernel)
{
    return sblas_spmv_mgpu_v1(m, n, nnz, alpha, csrVal, csrRowPtr, csrColIndex, x, beta, y, ngpu, kernel);
}
int spMV_mgpu_v2(int m, int n, long long nnz, double *alpha, double *csrVal, long long *csrRowPtr,
                 int *csrColIndex, double *x, double *beta, double *y, int ngpu, int kernel, long long nb,
                 int copy_of_workspace)
{
    return sblas_spmv_mgpu_v2(m, n, nnz, alpha, csrVal, csrRowPtr, csrColIndex, x, beta, y, ngpu, kernel, nb,
                              copy_of_workspace);
}
int get_row_from_index(int n, long long *a, long long idx) { return sblas_get_row_from_index(n, a, idx); }
double get_time() { return sblas_get_time(); }
double get_gpu_availble_mem(int ngpu) { return sblas_get_gpu_availble_mem(ngpu); }
