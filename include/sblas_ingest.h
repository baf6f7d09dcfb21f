/* sblas_ingest.h -- SURVEY.md section 8(f) row 1: a CORRECT Matrix-Market -> CSR ingest, opt-in.
 *
 * The SpMV harness of the reference uses the COO arrays of the file, in file order, AS the CSR
 * arrays (spmv/test/dspmv_test.cu:111-136,228-251): the result is only a row-major
 * CSR when the file happens to be sorted by row, and the symmetric flag is ignored.  That
 * behaviour stays the default of test_spmv (parity).  These two entry points restate the loader
 * the reference uses everywhere else (sptrsv/sptrsv_v1/src/mmio_highlevel.h:8-136 mmio_info,
 * :139-298 mmio_data): entries bucketed by row in file order (stable), off-diagonal entries of
 * symmetric / hermitian files mirrored with the same value, `pattern` files get 1.0, `integer`
 * values are converted, `complex` keeps the real part -- with a 64-bit row pointer and nnz.
 * Plain C, host only, no CUDA.
 */
#ifndef SBLAS_INGEST_H
#define SBLAS_INGEST_H
#ifdef __cplusplus
extern "C" {
#endif

/* sizes of the CSR the file expands to.  returns 0; -1 cannot open; -2 bad banner; -4 bad size
 * line; -5 short or malformed entry list (mmio_highlevel.h:21,26,39 use the same codes); -6 out of memory;
 * -7 a symmetric / hermitian file that is not square (the mirrored entries would fall outside the matrix) */
int sblas_mtx_info(const char *path, int *m, int *n, long long *nnz, int *is_symmetric);

/* fill csrRowPtr[m+1], csrColIndex[nnz], csrVal[nnz] (caller-allocated from sblas_mtx_info) */
int sblas_mtx_read_csr(const char *path, long long *csrRowPtr, int *csrColIndex, double *csrVal);

#ifdef __cplusplus
}
#endif
#endif
