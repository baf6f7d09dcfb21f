/* test_spmv.c -- the `test_spmv` command line of the reference, rewritten in C against
 * libsblas_spmv.so.  Same argv, same stdout lines (run_test.py scrapes the `m: ` and
 * `Average` lines), same data preparation:
 *
 *   ./test_spmv f <matrix.mtx> <ngpu> <repeats> <kernel 1-3> <f|b>
 *   ./test_spmv g <n>          <ngpu> <repeats> <kernel 1-3>
 *
 * Follows spmv/test/dspmv_test.cu:38-472 of the reference step by step: argument checks
 * (:41-83), the Matrix-Market loader that keeps the file's entry order (:101-136, the COO
 * arrays are later used AS the CSR arrays: SURVEY.md F3), the `g` generator with glibc
 * rand() (:137-208), COO -> row pointer (:217-251), x = 1, y = 0, ALPHA/BETA = rand()
 * (:253-282), one warm-up v1 call (:304-311), the v2 (GPU count, copies) sweep (:314-332),
 * the timed loop with the abs-1e-3 comparison against the baseline (:346-440) and the
 * Average row (:443-465).
 *
 * SBLAS_INGEST=csr (opt-in): `f` mode reads the file with the correct
 * loader of include/sblas_ingest.h (rows bucketed, symmetric files expanded) instead of using the
 * COO arrays in file order as CSR; everything after the load is unchanged.
 *
 * Extras (only with SBLAS_REPORT=1, printed AFTER the Average row so the scraped lines do
 * not move): GFLOP/s and algorithmic GB/s of the whole calls and of a resident plan.
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sblas_ingest.h"
#include "sblas_spmv.h"
#include "spmv_kernel.h"

static int read_banner_and_size(FILE *f, int *m, int *n, int *nnz)
{
    /* what mm_read_banner + mm_read_mtx_crd_size (spmv/include/mmio.h:254,339) consume:
     * one banner line of five tokens starting with %%MatrixMarket, then comment lines, then
     * the first line that parses as three integers */
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

static void *pinned(size_t bytes)
{
    void *p = NULL;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        fprintf(stderr, "cudaMallocHost(%zu) failed\n", bytes);
        exit(1);
    }
    return p;
}

int main(int argc, char *argv[])
{
    if (argc < 6) {
        printf("Incorrect number of arguments!\n");
        printf("Usage ./spmv [input matrix file] [number of GPU(s)] [number of test(s)] [kernel version (1-3)] [data type ('f' or 'b')]\n");
        return -1;
    }
    const char input_type = argv[1][0];
    const char *filename = argv[2];
    const int ngpu = atoi(argv[3]);
    const int repeat_test = atoi(argv[4]);
    const int kernel_version = atoi(argv[5]);

    int m = 0, n = 0;
    long long nnz = 0;
    int *cooRowIndex = NULL, *cooColIndex = NULL;
    double *cooVal = NULL;
    long long *csrRowPtr = NULL;
    const char *ingest = getenv("SBLAS_INGEST");
    const int ingest_csr = ingest && !strcmp(ingest, "csr");

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
    if (kernel_version != 1 && kernel_version != 2 && kernel_version != 3) {
        printf("Error: The kernel version can only be: 1, 2, or 3.\n");
        return -1;
    }
    printf("Using %d GPU(s).\n", ngpu);
    printf("Kernel #%d is selected.\n", kernel_version);

    if (input_type == 'f') {
        if (argc < 7) {
            printf("Incorrect number of arguments!\n");
            return -1;
        }
        printf("Loading input matrix from %s\n", filename);
        if (ingest_csr) {
            int sym = 0;
            int rc = sblas_mtx_info(filename, &m, &n, &nnz, &sym);
            if (rc == -2) printf("Could not process Matrix Market banner.\n");
            if (rc != 0) exit(1);
            printf("m: %d n: %d nnz: %lld\n", m, n, nnz);
            if (nnz >= 2147483647LL) {
                printf("Error: nnz does not fit the harness's int loops.\n");
                return -1;
            }
            csrRowPtr = (long long *)pinned((size_t)(m + 1) * sizeof(long long));
            cooColIndex = (int *)pinned((size_t)nnz * sizeof(int));
            cooVal = (double *)pinned((size_t)nnz * sizeof(double));
            if (sblas_mtx_read_csr(filename, csrRowPtr, cooColIndex, cooVal) != 0) exit(1);
            if (argv[6][0] == 'b')
                for (long long i = 0; i < nnz; i++) cooVal[i] = 0.00001;
            goto loaded;
        }
        FILE *f = fopen(filename, "r");
        if (!f) exit(1);
        int nnz_int = 0;
        const int rc = read_banner_and_size(f, &m, &n, &nnz_int);
        if (rc == 1) {
            printf("Could not process Matrix Market banner.\n");
            exit(1);
        } else if (rc != 0) {
            exit(1);
        }
        nnz = nnz_int;
        printf("m: %d n: %d nnz: %lld\n", m, n, nnz);
        if (m <= 0 || n <= 0 || nnz < 0) {
            printf("Error: the size line must hold a positive row and column count and a non-negative entry count.\n");
            exit(1);
        }
        cooRowIndex = (int *)pinned((size_t)nnz * sizeof(int));
        cooColIndex = (int *)pinned((size_t)nnz * sizeof(int));
        cooVal = (double *)pinned((size_t)nnz * sizeof(double));
        const char data_type = argv[6][0];
        long long nread = 0;
        for (int i = 0; i < nnz; i++) {
            if (data_type == 'b') {
                if (fscanf(f, "%d %d\n", &cooRowIndex[i], &cooColIndex[i]) < 2) break;
                cooVal[i] = 0.00001;
            } else if (data_type == 'f') {
                if (fscanf(f, "%d %d %lg\n", &cooRowIndex[i], &cooColIndex[i], &cooVal[i]) < 3) break;
            } else {
                break;
            }
            cooRowIndex[i]--;
            cooColIndex[i]--;
            if (cooRowIndex[i] < 0 || cooColIndex[i] < 0)
                printf("i = %d [%d, %d] = %g\n", i, cooRowIndex[i], cooColIndex[i], cooVal[i]);
            nread++;
        }
        fclose(f);
        /* The reference carries on with whatever it read (dspmv_test.cu:150-166) and then counts rows through the
         * indices unchecked; the file is untrusted input, so a short entry list or an index outside the matrix
         * stops here instead of writing past the row counters. */
        if (nread < nnz) {
            printf("Error: the file holds %lld of the %lld entries its size line announces.\n", nread, nnz);
            exit(1);
        }
        for (long long i = 0; i < nnz; i++)
            if (cooRowIndex[i] < 0 || cooRowIndex[i] >= m || cooColIndex[i] < 0 || cooColIndex[i] >= n) {
                printf("Error: entry %lld [%d, %d] lies outside the %d x %d matrix.\n", i, cooRowIndex[i] + 1,
                       cooColIndex[i] + 1, m, n);
                exit(1);
            }
    } else if (input_type == 'g') {
        n = atoi(filename);
        m = n;
        const int nb = m / 8;
        if (nb <= 0 || m % 8 != 0) {
            /* the reference loops forever (nb == 0) or writes rows >= m here */
            printf("Error: in g mode n must be a positive multiple of 8.\n");
            return -1;
        }
        const double r1 = 0.9, r2 = 0.01;
        double r;
        long long p = 0;
        for (int i = 0; i < m; i += nb) {
            r = (i == 0) ? r1 : r2;
            long long per_row = 0;
            for (int j = 0; j < n * r; j++) per_row++;
            p += per_row * nb;
        }
        nnz = p;
        printf("m: %d n: %d nnz: %lld\n", m, n, nnz);
        if (nnz >= 2147483647LL) {
            printf("Error: nnz does not fit the harness's int loops.\n");
            return -1;
        }
        cooRowIndex = (int *)pinned((size_t)nnz * sizeof(int));
        cooColIndex = (int *)pinned((size_t)nnz * sizeof(int));
        cooVal = (double *)pinned((size_t)nnz * sizeof(double));
        p = 0;
        printf("Start generating data ");
        fflush(stdout);
        for (int i = 0; i < m; i += nb) {
            printf(".");
            fflush(stdout);
            r = (i == 0) ? r1 : r2;
            for (int ii = i; ii < i + nb; ii++) {
                for (int j = 0; j < n * r; j++) {
                    cooRowIndex[p] = ii;
                    cooColIndex[p] = j;
                    cooVal[p] = (double)rand() / (RAND_MAX);
                    p++;
                }
            }
        }
        printf("\n");
        printf("Done generating data.\n");
    } else {
        printf("Incorrect number of arguments!\n");
        return -1;
    }

loaded:;
    /* COO -> row pointer; the COO col/val arrays are used as they are */
    const int have_csr = csrRowPtr != NULL;
    if (!have_csr) csrRowPtr = (long long *)pinned((size_t)(m + 1) * sizeof(long long));
    const long long matrix_data_space =
        nnz * (long long)sizeof(double) + nnz * (long long)sizeof(int) + (long long)(m + 1) * (long long)sizeof(int);
    printf("Matrix space size: %g GB.\n", (double)matrix_data_space / 1e9);
    if (!have_csr) {
        int *counter = (int *)calloc((size_t)(m > 0 ? m : 1), sizeof(int));
        for (long long i = 0; i < nnz; i++) counter[cooRowIndex[i]]++;
        csrRowPtr[0] = 0;
        for (int i = 1; i <= m; i++) csrRowPtr[i] = csrRowPtr[i - 1] + counter[i - 1];
        free(counter);
    }

    double *x = (double *)pinned((size_t)n * sizeof(double));
    double *y1 = (double *)pinned((size_t)m * sizeof(double));
    double *y2 = (double *)malloc((size_t)m * sizeof(double));
    double *y3 = (double *)pinned((size_t)m * sizeof(double));
    for (int i = 0; i < n; i++) x[i] = 1.0;
    for (int i = 0; i < m; i++) { y1[i] = 0.0; y2[i] = 0.0; y3[i] = 0.0; }

    double ALPHA = (double)rand() / (RAND_MAX);
    double BETA = (double)rand() / (RAND_MAX);

    double time_baseline = 0.0, time_v1 = 0.0, time_v2 = 0.0;
    double avg_time_baseline = 0.0, avg_time_v1 = 0.0, avg_time_v2 = 0.0;
    double curr_time = 0.0, profile_time = 0.0, min_profile_time = 1e300;
    double best_dev_count = 0.0, best_copy = 0.0;

    printf("Warming up GPU(s)...\n");
    spMV_mgpu_v1(m, n, nnz, &ALPHA, cooVal, csrRowPtr, cooColIndex, x, &BETA, y2, ngpu, kernel_version);

    /* v2 "auto GPU count": sweep devices x workspace copies, keep the fastest */
    for (int d = 1; d <= ngpu; d *= 2) {
        for (int c = 1; c <= 8; c *= 2) {
            curr_time = get_time();
            spMV_mgpu_v2(m, n, nnz, &ALPHA, cooVal, csrRowPtr, cooColIndex, x, &BETA, y3, d, kernel_version,
                         nnz / (d * c), c);
            profile_time = get_time() - curr_time;
            if (profile_time < min_profile_time) {
                min_profile_time = profile_time;
                best_dev_count = d;
                best_copy = c;
            }
        }
    }

    int ret1 = 0, ret2 = 0, ret3 = 0;
    printf("Starting tests...\n");
    printf("  Test No.   Baseline    Version 1     Pass     Version 2     Pass\n");
    printf("              Time(s)      Time(s)                Time(s)         \n");
    printf("=======================================================================\n");

    for (int i = 0; i < repeat_test; i++) {
        for (int k = 0; k < m; k++) { y1[k] = 0.0; y2[k] = 0.0; y3[k] = 0.0; }

        curr_time = get_time();
        ret1 = spMV_mgpu_baseline(m, n, nnz, &ALPHA, cooVal, csrRowPtr, cooColIndex, x, &BETA, y1, ngpu);
        time_baseline = get_time() - curr_time;

        curr_time = get_time();
        ret2 = spMV_mgpu_v1(m, n, nnz, &ALPHA, cooVal, csrRowPtr, cooColIndex, x, &BETA, y2, ngpu, kernel_version);
        time_v1 = get_time() - curr_time;

        curr_time = get_time();
        ret3 = spMV_mgpu_v2(m, n, nnz, &ALPHA, cooVal, csrRowPtr, cooColIndex, x, &BETA, y3, (int)best_dev_count,
                            kernel_version, (long long)(nnz / (best_dev_count * best_copy)), (int)best_copy);
        time_v2 = get_time() - curr_time;

        avg_time_baseline += time_baseline;
        avg_time_v1 += time_v1;
        avg_time_v2 += time_v2;

        int correct1 = 1, correct2 = 1;
        for (int k = 0; k < m; k++) {
            if (fabs(y1[k] - y2[k]) > 1e-3) correct1 = 0;
            if (fabs(y1[k] - y3[k]) > 1e-3) correct2 = 0;
        }

        printf("%10d", i + 1);
        if (ret1 == 0) printf("%11g", time_baseline); else printf("%11s", "Failed");
        if (ret2 == 0) printf("%13g", time_v1); else printf("%13s", "Failed");
        if (ret1 == 0) printf("%9s", correct1 ? "Y" : "N"); else printf("%9s", "N/A");
        if (ret3 == 0) printf("%14g", time_v2); else printf("%14s", "Failed.");
        if (ret1 == 0) printf("%9s", correct2 ? "Y" : "N"); else printf("%9s", "N/A");
        printf("\n");
    }

    avg_time_baseline /= repeat_test;
    avg_time_v1 /= repeat_test;
    avg_time_v2 /= repeat_test;

    printf(".......................................................................\n");
    printf("%10s ", "Average");
    if (ret1 == 0) printf("%11g", avg_time_baseline); else printf("%11s", "Failed");
    if (ret2 == 0) printf("%13g", avg_time_v1); else printf("%13s", "Failed");
    if (ret3 == 0) printf("%23g", avg_time_v2); else printf("%23s", "Failed");
    printf("\n");

    if (getenv("SBLAS_REPORT")) {
        /* not part of the reference output: throughput of the whole calls and of a resident plan */
        const double flop = 2.0 * (double)nnz;
        printf("[sblas] v2 sweep picked %d GPU(s) x %d copies\n", (int)best_dev_count, (int)best_copy);
        printf("[sblas] whole-call GFLOP/s (incl. upload): baseline %.3f  v1 %.3f  v2 %.3f\n",
               flop / avg_time_baseline / 1e9, flop / avg_time_v1 / 1e9, flop / avg_time_v2 / 1e9);
        sblas_spmv_plan *plan = NULL;
        if (sblas_spmv_plan_create(&plan, SBLAS_V1, m, n, nnz, cooVal, csrRowPtr, cooColIndex, ngpu, kernel_version, 0, 1) == 0) {
            const int reps = 20;
            sblas_spmv_plan_upload(plan, x, y1);
            sblas_spmv_plan_execute_device(plan, ALPHA, BETA, 1);
            const double t0 = get_time();
            for (int i = 0; i < reps; i++) sblas_spmv_plan_execute_device(plan, ALPHA, BETA, 0);
            sblas_spmv_plan_execute_device(plan, ALPHA, BETA, 1);
            const double t = (get_time() - t0) / (reps + 1);
            const double bytes = sblas_spmv_plan_alg_bytes(plan, BETA != 0.0, -1);
            printf("[sblas] resident plan (v1, %d GPU): %.6f s/SpMV  %.1f GFLOP/s  %.1f GB/s algorithmic (%.1f%% of %d x 8000 GB/s)\n",
                   ngpu, t, flop / t / 1e9, bytes / t / 1e9, 100.0 * bytes / t / 1e9 / (8000.0 * ngpu), ngpu);
            if (m == n) {
                /* chained products x <- y (device-side NVLink all-gather between products, beta = 0 so
                 * that the iterates stay bounded only by the matrix; timing only) */
                const double t1 = get_time();
                for (int i = 0; i < reps; i++) {
                    sblas_spmv_plan_execute_device(plan, ALPHA, 0.0, 0);
                    sblas_spmv_plan_chain(plan);
                }
                sblas_spmv_plan_execute_device(plan, ALPHA, 0.0, 1);
                const double tc = (get_time() - t1) / (reps + 1);
                printf("[sblas] chained (x <- y on the GPUs, %d GPU): %.6f s/SpMV  %.1f GFLOP/s  (all-gather of %.1f MB per GPU per product)\n",
                       ngpu, tc, flop / tc / 1e9, 8.0 * m / 1e6);
            }
            sblas_spmv_plan_destroy(plan);
        }
    }

    cudaFreeHost(cooRowIndex);
    cudaFreeHost(cooColIndex);
    cudaFreeHost(cooVal);
    cudaFreeHost(csrRowPtr);
    cudaFreeHost(x);
    cudaFreeHost(y1);
    cudaFreeHost(y3);
    free(y2);
    return 0;
}
