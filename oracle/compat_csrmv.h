/* oracle/compat_csrmv.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The reference's three entry points (spmv/src/dspmv_mgpu_baseline.cu:163,
 * dspmv_mgpu_v1.cu:200,206, dspmv_mgpu_v2.cu:351,357) call the legacy cuSPARSE routines
 * cusparseDcsrmv / cusparseDcsrmv_mp, which were removed in CUDA 11.  This header, force-included
 * with `nvcc -include oracle/compat_csrmv.h`, supplies those two names (and cusparseDcsrmm, which the
 * reference's SpMM calls, spmm/src/dspmm_mgpu_baseline.cu:225,450) on top of the generic API
 * that replaced them (cusparseSpMV, CUSPARSE_SPMV_CSR_ALG1 for csrmv and _ALG2 -- the
 * load-balanced "merge path" algorithm -- for csrmv_mp), so that the UNMODIFIED reference sources
 * compile where they lie into oracle/_ref/libref_spmv.so (oracle/Makefile, target refspmv).
 * That library is the reference arithmetic + partition + upload + host merge run on the GPU box:
 * the -m gpu tests compare this repo's library with it on the same host arrays, and
 * tests/golden/make_golden_y.py commits its y vectors as fixtures.
 *
 * One limit of the mapping: cusparseCreateCsr refuses nnz > rows*cols (a shard of a matrix whose rows repeat a
 * column can exceed it); the legacy routines did not check.  The tests leave such v2 tasks out.
 *
 * Same argument conventions as the legacy calls: alpha/beta are host pointers (pointer mode host),
 * int32 base-0 CSR, y = alpha*op(A)*x + beta*y, work enqueued on the handle's stream.
 */
#ifndef SBLAS_ORACLE_COMPAT_CSRMV_H
#define SBLAS_ORACLE_COMPAT_CSRMV_H
#include <cuda_runtime.h>
#include <cusparse.h>

static inline cusparseStatus_t sblas_compat_csrmv(cusparseHandle_t handle, cusparseOperation_t transA, int m, int n,
                                                  int nnz, const double *alpha, const cusparseMatDescr_t descrA,
                                                  const double *csrVal, const int *csrRowPtr, const int *csrColInd,
                                                  const double *x, const double *beta, double *y,
                                                  cusparseSpMVAlg_t alg)
{
    (void)descrA;                                    /* general, base 0: the only descriptor the reference builds */
    cusparseSpMatDescr_t A = NULL;
    cusparseDnVecDescr_t vx = NULL, vy = NULL;
    cusparseStatus_t st;
    cudaStream_t s = 0;
    void *buf = NULL;
    size_t need = 0;
    const int xr = transA == CUSPARSE_OPERATION_NON_TRANSPOSE ? n : m;
    const int yr = transA == CUSPARSE_OPERATION_NON_TRANSPOSE ? m : n;
    if (m == 0) return CUSPARSE_STATUS_SUCCESS;      /* nothing to write */
    st = cusparseGetStream(handle, &s);
    if (st != CUSPARSE_STATUS_SUCCESS) return st;
    st = cusparseCreateCsr(&A, m, n, nnz, (void *)csrRowPtr, (void *)csrColInd, (void *)csrVal, CUSPARSE_INDEX_32I,
                           CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F);
    if (st != CUSPARSE_STATUS_SUCCESS) return st;
    st = cusparseCreateDnVec(&vx, xr, (void *)x, CUDA_R_64F);
    if (st == CUSPARSE_STATUS_SUCCESS) st = cusparseCreateDnVec(&vy, yr, (void *)y, CUDA_R_64F);
    if (st == CUSPARSE_STATUS_SUCCESS)
        st = cusparseSpMV_bufferSize(handle, transA, alpha, A, vx, beta, vy, CUDA_R_64F, alg, &need);
    if (st == CUSPARSE_STATUS_SUCCESS && need > 0 && cudaMalloc(&buf, need) != cudaSuccess)
        st = CUSPARSE_STATUS_ALLOC_FAILED;
    if (st == CUSPARSE_STATUS_SUCCESS)
        st = cusparseSpMV(handle, transA, alpha, A, vx, beta, vy, CUDA_R_64F, alg, buf);
    /* the legacy call was asynchronous; the scratch buffer forces a wait here (the reference
     * synchronises every device right after the calls anyway, dspmv_mgpu_v1.cu:222-229) */
    cudaStreamSynchronize(s);
    if (buf) cudaFree(buf);
    if (vy) cusparseDestroyDnVec(vy);
    if (vx) cusparseDestroyDnVec(vx);
    if (A) cusparseDestroySpMat(A);
    return st;
}

static inline cusparseStatus_t cusparseDcsrmv(cusparseHandle_t handle, cusparseOperation_t transA, int m, int n, int nnz,
                                              const double *alpha, const cusparseMatDescr_t descrA,
                                              const double *csrVal, const int *csrRowPtr, const int *csrColInd,
                                              const double *x, const double *beta, double *y)
{
    return sblas_compat_csrmv(handle, transA, m, n, nnz, alpha, descrA, csrVal, csrRowPtr, csrColInd, x, beta, y,
                              CUSPARSE_SPMV_CSR_ALG1);
}

static inline cusparseStatus_t cusparseDcsrmv_mp(cusparseHandle_t handle, cusparseOperation_t transA, int m, int n,
                                                 int nnz, const double *alpha, const cusparseMatDescr_t descrA,
                                                 const double *csrVal, const int *csrRowPtr, const int *csrColInd,
                                                 const double *x, const double *beta, double *y)
{
    return sblas_compat_csrmv(handle, transA, m, n, nnz, alpha, descrA, csrVal, csrRowPtr, csrColInd, x, beta, y,
                              CUSPARSE_SPMV_CSR_ALG2);
}

/* cusparseDcsrmm (spmm/src/dspmm_mgpu_baseline.cu:225-241, :450-466): C = alpha*op(A)*B + beta*C, A CSR,
 * B (k x n, ld ldb) and C (m x n, ld ldc) column-major -> cusparseSpMM, default algorithm. */
static inline cusparseStatus_t cusparseDcsrmm(cusparseHandle_t handle, cusparseOperation_t transA, int m, int n, int k,
                                              int nnz, const double *alpha, const cusparseMatDescr_t descrA,
                                              const double *csrVal, const int *csrRowPtr, const int *csrColInd,
                                              const double *B, int ldb, const double *beta, double *C, int ldc)
{
    (void)descrA;
    cusparseSpMatDescr_t A = NULL;
    cusparseDnMatDescr_t mB = NULL, mC = NULL;
    cusparseStatus_t st;
    cudaStream_t s = 0;
    void *buf = NULL;
    size_t need = 0;
    if (m == 0 || n == 0) return CUSPARSE_STATUS_SUCCESS;
    st = cusparseGetStream(handle, &s);
    if (st != CUSPARSE_STATUS_SUCCESS) return st;
    st = cusparseCreateCsr(&A, m, k, nnz, (void *)csrRowPtr, (void *)csrColInd, (void *)csrVal, CUSPARSE_INDEX_32I,
                           CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F);
    if (st != CUSPARSE_STATUS_SUCCESS) return st;
    st = cusparseCreateDnMat(&mB, k, n, ldb, (void *)B, CUDA_R_64F, CUSPARSE_ORDER_COL);
    if (st == CUSPARSE_STATUS_SUCCESS) st = cusparseCreateDnMat(&mC, m, n, ldc, (void *)C, CUDA_R_64F, CUSPARSE_ORDER_COL);
    if (st == CUSPARSE_STATUS_SUCCESS)
        st = cusparseSpMM_bufferSize(handle, transA, CUSPARSE_OPERATION_NON_TRANSPOSE, alpha, A, mB, beta, mC, CUDA_R_64F,
                                     CUSPARSE_SPMM_ALG_DEFAULT, &need);
    if (st == CUSPARSE_STATUS_SUCCESS && need > 0 && cudaMalloc(&buf, need) != cudaSuccess) st = CUSPARSE_STATUS_ALLOC_FAILED;
    if (st == CUSPARSE_STATUS_SUCCESS)
        st = cusparseSpMM(handle, transA, CUSPARSE_OPERATION_NON_TRANSPOSE, alpha, A, mB, beta, mC, CUDA_R_64F,
                          CUSPARSE_SPMM_ALG_DEFAULT, buf);
    cudaStreamSynchronize(s);
    if (buf) cudaFree(buf);
    if (mC) cusparseDestroyDnMat(mC);
    if (mB) cusparseDestroyDnMat(mB);
    if (A) cusparseDestroySpMat(A);
    return st;
}
#endif
