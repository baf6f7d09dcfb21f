/*
 * sblas_synth.h -- deterministic synthetic CSR content generated ON the GPU, for
 * bench.py and the full-size tests (the named configs are
 * synthetic shapes; there is no network for SuiteSparse files).  Not part of the
 * reference path: the reference's only generator is the CLI's `g` mode
 * (spmv/test/dspmv_test.cu:137-208), whose STRUCTURE (col = 0..k-1 per row) is
 * pattern SBLAS_COLS_PREFIX here; its glibc rand() values are reproduced on the
 * host by test_spmv.c, not by these kernels.
 */
#ifndef SBLAS_SYNTH_H
#define SBLAS_SYNTH_H
#ifdef __cplusplus
extern "C" {
#endif

enum {
    SBLAS_COLS_PREFIX = 0,   /* col = j                      (the `g` generator)                 */
    SBLAS_COLS_BANDED = 1,   /* sorted-unique inside a window of +-band around the diagonal      */
    SBLAS_COLS_UNIFORM = 2,  /* sorted-unique, stratified over [0,n)                             */
    SBLAS_COLS_CIRCUIT = 3,  /* 80 % of the rows banded, 20 % uniform (hub rows); long rows uniform */
    SBLAS_COLS_BANDRUN = 4   /* banded, in runs of 16 consecutive columns (block-structured / multi-DOF
                                meshes): run starts are sorted-unique inside the +-band window */
};

/* Fill val/col for the global nnz range [k0,k1) (written at out[k-k0]) of a matrix
 * whose int64 row pointer entries for rows [row_first, row_first+nrows] are at
 * d_rowptr (DEVICE pointer).  value_mode: 0 = uniform(0,1) hashed from (seed,k),
 * 1 = the constant `value_const`.  Runs on `stream` (a cudaStream_t, may be NULL). */
int sblas_synth_fill_csr(const long long *d_rowptr, int row_first, int nrows, long long k0, long long k1,
                         int n, int cols_mode, long long band, unsigned long long seed, int value_mode,
                         double value_const, double *d_val, int *d_col, void *stream);

/* p[i] = lo + (hi-lo)*u(seed,i) */
int sblas_synth_fill_uniform(double *d_p, long long count, unsigned long long seed, double lo, double hi, void *stream);

/* Read-only HBM bandwidth of this GPU in GB/s (measurement support for bench.py's roofline: the
 * driver's measured peak is a read+write copy, which a read-dominated kernel can exceed): best of
 * `reps` passes of a plain 128-bit streaming-load kernel over `bytes` of device memory at d_buf,
 * timed with CUDA events on `stream`.  < 0 on failure. */
double sblas_synth_read_probe(const void *d_buf, unsigned long long bytes, int reps, void *stream);

#ifdef __cplusplus
}
#endif
#endif
