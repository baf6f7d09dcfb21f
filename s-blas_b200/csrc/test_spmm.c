/* test_spmm.c -- the `test_spmm` command line of the reference, in C against libsblas_spmv.so
 * (SURVEY.md section 8f-2).  Same argv and the stdout lines run_test.py scrapes
 * (run_test.py:146-175: `Matrix A --`, `Matrix B --`, `SPMM:`):
 *
 *   ./test_spmm <matrix A .mtx> <n: columns of B and C> <ngpu> <repeats>
 *
 * Follows spmm/test/dspmm_baseline_test.cu:381-559: argument checks (:382-404), the loader (one
 * "%d %d %lg" per entry, :441-458) followed by a sort by (row, column) (:459, :41-55) -- unlike the SpMV
 * harness the SpMM harness DOES build a proper CSR --, COO -> int32 row pointer (:474-494), B then C filled
 * with glibc rand()/RAND_MAX in that order (:499-512), alpha = -0.7, beta = 0.8 (:516-517), one single-GPU
 * product as the truth (:521-528; the reference uses cuSPARSE on one GPU, here the library on one GPU), the
 * multi-GPU product through cusparse_mgpu_csrmm_omp (:531-537) and the abs-1e-3 comparison (:540-545).
 * SBLAS_REPORT=1 adds GFLOP/s lines after the check (the scraped lines do not move).
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sblas_spmm.h"
#include "sblas_spmv.h"
#include "spmm_kernel.h"

typedef struct { int r, c; double v; } rcv;

static int cmp_rcv(const void *aa, const void *bb)          /* dspmm_baseline_test.cu:27-38 */
{
    const rcv *a = (const rcv *)aa, *b = (const rcv *)bb;
    if (a->r != b->r) return a->r > b->r ? 1 : -1;
    if (a->c != b->c) return a->c > b->c ? 1 : -1;
    return 0;
}

static void *pinned(size_t bytes)
{
    void *p = NULL;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        fprintf(stderr, "cudaMallocHost(%zu) failed\n", bytes);
        exit(1);
    }
    return p;
}

static int read_banner_and_size(FILE *f, int *m, int *n, int *nnz)
{
    char line[1025], t0[64], t1[64], t2[64], t3[64], t4[64];
    if (!fgets(line, sizeof line, f)) return 1;
    if (sscanf(line, "%63s %63s %63s %63s %63s", t0, t1, t2, t3, t4) != 5) return 1;
    if (strncmp(t0, "%%MatrixMarket", 14) != 0) return 1;
    do {
        if (!fgets(line, sizeof line, f)) return 2;
    } while (line[0] == '%');
    for (;;) {
        if (sscanf(line, "%d %d %d", m, n, nnz) == 3) return 0;
        if (!fgets(line, sizeof line, f)) return 2;
    }
}

int main(int argc, char *argv[])
{
    if (argc < 5) {
        printf("Usage: ./spmm [input sparse matrix A file] [output row number] [number of GPU(s)] [number of test(s)]\n");
        return -1;
    }
    const char *filename_A = argv[1];
    const int n = atoi(argv[2]);
    const int ngpu = atoi(argv[3]);
    const int repeat_test = atoi(argv[4]);
    int deviceCount = 0;
    cudaGetDeviceCount(&deviceCount);
    if (deviceCount < ngpu) {
        printf("Error: Not enough number of GPUs. Only %davailable.\n", deviceCount);
        return -1;
    }
    if (ngpu <= 0) {
        printf("Error: Number of GPU(s) needs to be greater than 0.\n");
        return -1;
    }
    if (n <= 0) {
        printf("Error: the number of columns of B needs to be greater than 0.\n");
        return -1;
    }
    printf("Using %d GPU(s).\n", ngpu);

    int m = 0, k = 0, nnz = 0;
    FILE *f = fopen(filename_A, "r");
    if (!f) { printf("Could not open matrix A file.\n"); exit(1); }
    const int rb = read_banner_and_size(f, &m, &k, &nnz);
    if (rb == 1) { printf("Could not process Matrix Market banner for matrix A.\n"); exit(1); }
    if (rb == 2) { printf("Could not read Matrix Market format for matrix A.\n"); exit(1); }
    printf("Matrix A -- #row: %d #col: %d nnz: %d\n", m, k, nnz);

    int *cooCol = (int *)pinned((size_t)nnz * sizeof(int));
    double *cooVal = (double *)pinned((size_t)nnz * sizeof(double));
    rcv *ent = (rcv *)malloc((size_t)(nnz ? nnz : 1) * sizeof(rcv));
    printf("Loading input matrix A from %s\n", filename_A);
    for (int i = 0; i < nnz; ++i) {
        int r = 0, c = 0;
        double v = 0.0;
        if (fscanf(f, "%d %d %lg\n", &r, &c, &v) != 3) { printf("Could not read Matrix Market format for matrix A.\n"); exit(1); }
        ent[i].r = r - 1; ent[i].c = c - 1; ent[i].v = v;
        if (ent[i].r < 0 || ent[i].c < 0 || ent[i].r >= m || ent[i].c >= k) {
            printf("i = %d [%d, %d] = %g\n", i, ent[i].r, ent[i].c, v);
            exit(1);
        }
    }
    fclose(f);
    qsort(ent, (size_t)nnz, sizeof(rcv), cmp_rcv);
    int *csrRowPtr = (int *)pinned((size_t)(m + 1) * sizeof(int));
    memset(csrRowPtr, 0, (size_t)(m + 1) * sizeof(int));
    for (int i = 0; i < nnz; ++i) { csrRowPtr[ent[i].r + 1]++; cooCol[i] = ent[i].c; cooVal[i] = ent[i].v; }
    for (int i = 1; i <= m; ++i) csrRowPtr[i] += csrRowPtr[i - 1];
    free(ent);

    const double space = ((double)nnz * 12.0 + (double)(m + 1) * 4.0 + ((double)k * n + (double)m * n) * 8.0) / 1e9;
    printf("Matrix space size(total): %g GB.\n", space);
    printf("Matrix B -- #row: %d #col: %d (dense)\n", k, n);
    printf("Start generating data for Matrix B\n");
    fflush(stdout);
    double *B = (double *)pinned((size_t)k * n * sizeof(double));
    double *C1 = (double *)pinned((size_t)m * n * sizeof(double));
    double *CN = (double *)pinned((size_t)m * n * sizeof(double));
    for (long long i = 0; i < (long long)k * n; ++i) B[i] = (double)rand() / RAND_MAX;
    for (long long i = 0; i < (long long)m * n; ++i) C1[i] = (double)rand() / RAND_MAX;
    memcpy(CN, C1, (size_t)m * n * sizeof(double));
    double alpha = -0.7, beta = 0.8;

    printf("Start computing SpMM on a single GPU (CuSPARSE).\n");
    fflush(stdout);
    double t0 = sblas_get_time();
    int rc = cusparse_mgpu_csrmm(m, n, k, &alpha, nnz, csrRowPtr, cooCol, cooVal, &beta, B, C1, 1);
    const double single = sblas_get_time() - t0;
    if (rc != 0) { printf("single GPU SpMM failed (%d): %s\n", rc, sblas_last_error()); return 1; }
    printf("CuSPARSE single gpu processing time(s): %g\n", single);
    printf("Matrix C -- #row: %d #col: %d (dense)\n", m, n);

    double mgpu = 0.0;
    double *C0 = NULL;
    if (repeat_test > 1) { C0 = (double *)malloc((size_t)m * n * sizeof(double)); memcpy(C0, CN, (size_t)m * n * sizeof(double)); }
    for (int rep = 0; rep < (repeat_test > 0 ? repeat_test : 1); ++rep) {
        if (rep > 0) memcpy(CN, C0, (size_t)m * n * sizeof(double));
        t0 = sblas_get_time();
        rc = cusparse_mgpu_csrmm_omp(m, n, k, &alpha, nnz, csrRowPtr, cooCol, cooVal, &beta, B, CN, ngpu);
        const double t = sblas_get_time() - t0;
        if (rc != 0) { printf("SpMM on %d GPUs failed (%d): %s\n", ngpu, rc, sblas_last_error()); return 1; }
        if (rep == 0 || t < mgpu) mgpu = t;
    }
    printf("SPMM: %d GPUs processing time(s): %g\n", ngpu, mgpu);
    int ok = 1;
    for (long long i = 0; i < (long long)m * n && ok; ++i) ok = fabs(CN[i] - C1[i]) < 0.001;
    printf("mgpu check: %s\n", ok ? "PASS" : "FAILED");

    if (getenv("SBLAS_REPORT")) {
        sblas_spmm_plan *P = NULL;
        if (sblas_spmm_plan_create(&P, m, k, nnz, csrRowPtr, cooCol, cooVal, ngpu) == 0) {
            sblas_spmm_plan_execute(P, n, &alpha, B, &beta, CN);
            t0 = sblas_get_time();
            sblas_spmm_plan_execute(P, n, &alpha, B, &beta, CN);
            const double t = sblas_get_time() - t0;
            printf("resident plan, host B and C: %g s, %.1f GFLOP/s\n", t, 2.0 * nnz * (double)n / t / 1e9);
            sblas_spmm_plan_destroy(P);
        }
        printf("whole call: %.1f GFLOP/s on %d GPU(s)\n", 2.0 * nnz * (double)n / mgpu / 1e9, ngpu);
    }
    free(C0);
    cudaFreeHost(cooCol); cudaFreeHost(cooVal); cudaFreeHost(csrRowPtr);
    cudaFreeHost(B); cudaFreeHost(C1); cudaFreeHost(CN);
    return 0;
}
