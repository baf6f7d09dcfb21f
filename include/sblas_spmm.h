/*
 * sblas_spmm.h -- C-ABI of the multi-GPU CSR SpMM of the library:
 *     C = alpha * A * B + beta * C,   A sparse m x k (CSR, int32 row pointer, base 0),
 *                                     B dense k x n and C dense m x n, COLUMN-major, double.
 * Replaces cusparse_mgpu_csrmm / cusparse_mgpu_csrmm_omp of the reference
 * (spmm/include/spmm_kernel.h:6-31; bodies spmm/src/dspmm_mgpu_baseline.cu:83-280 and :282-524) and the
 * cusparseDcsrmm call inside them (:436-451): A is resident on every GPU, B and C are split by COLUMNS,
 * GPU d owning columns [floor(d*n/ngpu), floor((d+1)*n/ngpu)) (dspmm_mgpu_baseline.cu:338-342).
 * Hand-written sm_100a kernels (s-blas_b200/csrc/sblas_spmm.cu); no cuSPARSE, no CPU fallback.
 *
 * Return convention of the reference: 0 success; -1 the matrices do not fit the free device memory
 * (dspmm_mgpu_baseline.cu:328-336) or a kernel failed (:455-461); 1 set-up failure.
 */
#ifndef SBLAS_SPMM_H
#define SBLAS_SPMM_H
#ifdef __cplusplus
extern "C" {
#endif

/* one-shot, host pointers in / host C out: argument for argument the reference's entry point */
int sblas_spmm_mgpu(int m, int n, int k, const double *alpha, int nnz_A, const int *csrRowPtr_A,
                    const int *csrColIndex_A, const double *csrVal_A, const double *beta,
                    const double *B_dense, double *C_dense, int ngpu);

/* Plan API: A uploaded once (every GPU takes a 1/ngpu slice over its own PCIe link, the slices are
 * exchanged over NVLink; the reference does ngpu full uploads), products re-use it. */
typedef struct sblas_spmm_plan sblas_spmm_plan;
int sblas_spmm_plan_create(sblas_spmm_plan **plan, int m, int k, int nnz_A, const int *csrRowPtr_A,
                           const int *csrColIndex_A, const double *csrVal_A, int ngpu);
/* C = alpha*A*B + beta*C with HOST B (k x n) and HOST C (m x n), column-major */
int sblas_spmm_plan_execute(sblas_spmm_plan *plan, int n, const double *alpha, const double *B_dense,
                            const double *beta, double *C_dense);
/* device-resident product on GPU `dev` of the plan: d_B (k x nd, column-major, ld k) and d_C (m x nd,
 * column-major, ld m) are DEVICE pointers on that GPU; enqueued on the plan's stream, no copies.
 * sync != 0 waits for it.  Used by bench.py to time the kernels alone. */
int sblas_spmm_plan_execute_device(sblas_spmm_plan *plan, int dev, int nd, double alpha, const double *d_B,
                                   double beta, double *d_C, int sync);
/* columns GPU dev owns for a given n: [*first, *first + *count)  (dspmm_mgpu_baseline.cu:338-342) */
int sblas_spmm_plan_columns(const sblas_spmm_plan *plan, int n, int dev, int *first, int *count);
void *sblas_spmm_plan_stream(sblas_spmm_plan *plan, int dev);
int sblas_spmm_plan_num_devices(const sblas_spmm_plan *plan);
void sblas_spmm_plan_destroy(sblas_spmm_plan *plan);

#ifdef __cplusplus
}
#endif
#endif
