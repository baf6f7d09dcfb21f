/*
 * spmv_kernel.h -- the reference's public SpMV header, kept signature-for-signature
 * (pnnl/s-blas spmv/include/spmv_kernel.h:11-36) so that code written against
 * s-BLAS compiles and links unchanged against libsblas_spmv.so.
 *
 * The reference declares these with C++ linkage (everything is compiled by nvcc).
 * Both worlds are served:
 *   - a C++ translation unit (e.g. the unmodified spmv/test/dspmv_test.cu) sees
 *     C++-linkage declarations; the library exports the matching mangled symbols
 *     (_Z18spMV_mgpu_baselineiixPdS_PxPiS_S_S_i, _Z12spMV_mgpu_v1iixPdS_PxPiS_S_S_ii,
 *     _Z12spMV_mgpu_v2iixPdS_PxPiS_S_S_iixi, _Z18get_row_from_indexiPxx, _Z8get_timev,
 *     _Z20get_gpu_availble_memi) from s-blas_b200/csrc/sblas_shim.cpp;
 *   - a C translation unit (this repo's host code, test_spmv.c) sees plain C
 *     functions of the same names, defined in s-blas_b200/csrc/sblas_api.c.
 * Both forward to the extern "C" sblas_* entry points of include/sblas_spmv.h.
 */
#ifndef SPMV_KERNEL
#define SPMV_KERNEL

int spMV_mgpu_baseline(int m, int n, long long nnz, double * alpha,
				 double * csrVal, long long * csrRowPtr, int * csrColIndex,
				 double * x, double * beta,
				 double * y,
				 int ngpu);

int spMV_mgpu_v1(int m, int n, long long nnz, double * alpha,
				  double * csrVal, long long * csrRowPtr, int * csrColIndex,
				  double * x, double * beta,
				  double * y,
				  int ngpu,
				  int kernel);

int spMV_mgpu_v2(int m, int n, long long nnz, double * alpha,
				  double * csrVal, long long * csrRowPtr, int * csrColIndex,
				  double * x, double * beta,
				  double * y,
				  int ngpu,
				  int kernel,
				  long long nb,
				  int copy_of_workspace);

int get_row_from_index(int n, long long * a, long long idx);

double get_time();

double get_gpu_availble_mem(int ngpu);

#endif /* SPMV_KERNEL */
