// sblas_shim.cpp -- the reference's entry points with the reference's C++ linkage
// (spmv/include/spmv_kernel.h:11-36 has no extern "C"), so that the
// UNMODIFIED spmv/test/dspmv_test.cu links against libsblas_spmv.so.  Thin forwards
// to the extern "C" layer; no logic here.
#include "sblas_spmv.h"

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

// spmm/include/spmm_kernel.h:6-31 (C++ linkage in the reference, like the SpMV entry points)
#include "sblas_spmm.h"
int cusparse_mgpu_csrmm(const int m, const int n, const int k, const double *alpha, const int nnz_A, int *csrRowPtr_A,
                        int *csrColIndex_A, double *csrVal_A, const double *beta, double *B_dense, double *C_dense,
                        const int ngpu)
{
    return sblas_spmm_mgpu(m, n, k, alpha, nnz_A, csrRowPtr_A, csrColIndex_A, csrVal_A, beta, B_dense, C_dense, ngpu);
}
int cusparse_mgpu_csrmm_omp(const int m, const int n, const int k, const double *alpha, const int nnz_A, int *csrRowPtr_A,
                            int *csrColIndex_A, double *csrVal_A, const double *beta, double *B_dense, double *C_dense,
                            const int ngpu)
{
    return sblas_spmm_mgpu(m, n, k, alpha, nnz_A, csrRowPtr_A, csrColIndex_A, csrVal_A, beta, B_dense, C_dense, ngpu);
}
