/* spmm_kernel.h -- the reference's SpMM interface (spmm/include/spmm_kernel.h:6-31), kept signature for
 * signature so that the reference's spmm/test/dspmm_baseline_test.cu compiles against this header
 * unchanged.  Implemented by libsblas_spmv.so (s-blas_b200/csrc/sblas_spmm_plan.c): hand-written sm_100a
 * kernels behind the same two entry points.  Like spmv_kernel.h there is no extern "C": a C caller sees C
 * functions (sblas_spmm_plan.c), a C++ caller the reference's C++ linkage (mangled exports in sblas_shim.cpp).  C = alpha*A*B + beta*C, A CSR (int32 row
 * pointer), B (k x n) and C (m x n) dense column-major, columns split over the GPUs. */
#ifndef SPMM_KERNEL
#define SPMM_KERNEL

int cusparse_mgpu_csrmm(const int m,
			const int n,
			const int k,
                        const double * alpha,
			const int nnz_A,
			int * csrRowPtr_A,
			int * csrColIndex_A,
			double * csrVal_A,
			const double * beta,
			double * B_dense,
			double * C_dense,
			const int ngpu);


int cusparse_mgpu_csrmm_omp(const int m,
			const int n,
			const int k,
                        const double * alpha,
			const int nnz_A,
			int * csrRowPtr_A,
			int * csrColIndex_A,
			double * csrVal_A,
			const double * beta,
			double * B_dense,
			double * C_dense,
			const int ngpu);

#endif /* SPMM_KERNEL */
