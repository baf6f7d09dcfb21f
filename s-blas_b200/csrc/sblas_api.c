/* sblas_api.c -- the one-shot entry points: host pointers in, host y out.
 * plan_create + execute + destroy, with the reference's return codes
 * (see include/sblas_spmv.h).  Also the plain-C definitions of the reference's own
 * names declared in include/spmv_kernel.h. */
#include <stdio.h>
#include <stdlib.h>
#include "sblas_spmv.h"
#include "spmv_kernel.h"

static int one_shot(int version, int m, int n, long long nnz, double *alpha, double *val, long long *rp,
                    int *col, double *x, double *beta, double *y, int ngpu, int kernel, long long nb, int q)
{
    sblas_spmv_plan *plan = NULL;
    int rc = sblas_spmv_plan_create(&plan, version, m, n, nnz, val, rp, col, ngpu, kernel, nb, q);
    if (rc != 0) {
        if (getenv("SBLAS_VERBOSE")) fprintf(stderr, "sblas: %s\n", sblas_last_error());
        return rc;
    }
    rc = sblas_spmv_plan_execute(plan, alpha, x, beta, y);
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
