/*
 * oracle/spmv_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the s-BLAS multi-GPU CSR SpMV path (y = alpha*A*x + beta*y,
 * double precision) used ONLY as the checker by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  Nothing under
 * s-blas_b200/ may include, link or call this file.
 *
 * Parity status
 *   - partitioners / get_row_from_index / local row pointers / loader / generator:
 *     PINNED.  They are checked bit-for-bit against (a) the reference's own
 *     spmv_helper.cu compiled into oracle/_ref/libref_helper.so (see
 *     oracle/Makefile) and (b) golden vectors in tests/golden/ produced from that
 *     object and from the reference's sample matrix by tests/golden/make_golden.py.
 *   - the arithmetic itself (the csrmv call): PINNED since round 2.  The reference delegates it to
 *     legacy cuSPARSE (cusparseDcsrmv / cusparseDcsrmv_mp, removed in CUDA 11; call sites
 *     spmv/src/dspmv_mgpu_baseline.cu:163, dspmv_mgpu_v1.cu:200,206, dspmv_mgpu_v2.cu:351,357) and
 *     ships no golden y vectors; oracle/compat_csrmv.h maps the two removed names onto cusparseSpMV, so
 *     the reference's own unmodified entry points compile into oracle/_ref/libref_spmv.so and run on the
 *     GPU box.  tests/golden/ref_y.npz holds the y vectors they produced on a B200
 *     (tests/golden/make_golden_y.py); tests/test_reference_y.py checks this file's csrmv and its
 *     restatements of the whole entry points (partition, per-shard csrmv, the host merges of
 *     dspmv_mgpu_v1.cu:235-248 and dspmv_mgpu_v2.cu:385-441) against them on the CPU, and the library
 *     against them and against the reference code run live on the GPU.
 *   - SpMM (oracle_csrmm) and the transposition (oracle_csr2csc): see the notes at those functions.
 *
 * Every function cites the reference file:line it follows.  Paths are relative
 * to the reference checkout.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef long long ll;

/* ------------------------------------------------------------------------- */
/* a9: the csrmv semantics.  y = alpha*A*x + beta*y, CSR base 0, left-to-right
 * accumulation per row.  rowptr is the harness's 64-bit row pointer
 * (spmv/test/dspmv_test.cu:219,247-251). */
void oracle_csr_spmv(int m, const ll *rowptr, const int *col, const double *val,
                     const double *x, double alpha, double beta, double *y)
{
    for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (ll k = rowptr[i]; k < rowptr[i + 1]; ++k) s += val[k] * x[col[k]];
        y[i] = alpha * s + beta * y[i];
    }
}

/* Same, rows in parallel over all host cores (the timed CPU baseline,
 * BASELINE.md section 4).  Returns the number of threads used. */
int oracle_csr_spmv_omp(int m, const ll *rowptr, const int *col, const double *val,
                        const double *x, double alpha, double beta, double *y)
{
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 256)
#endif
    for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (ll k = rowptr[i]; k < rowptr[i + 1]; ++k) s += val[k] * x[col[k]];
        y[i] = alpha * s + beta * y[i];
    }
    return nt;
}

/* nnz-balanced variant for the timed CPU baseline on skewed matrices: thread t
 * takes the rows whose first entry lies in its equal-nnz slice.  Same result as
 * oracle_csr_spmv (each row is still summed left to right by one thread). */
int oracle_csr_spmv_omp_balanced(int m, const ll *rowptr, const int *col, const double *val,
                                 const double *x, double alpha, double beta, double *y)
{
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    ll nnz = rowptr[m];
    int chunks = nt * 16;
    if (chunks > m) chunks = m > 0 ? m : 1;
    int *bound = (int *)malloc((size_t)(chunks + 1) * sizeof(int));
    bound[0] = 0;
    for (int c = 1; c < chunks; ++c) {
        ll target = (ll)((double)nnz * c / chunks);
        int lo = bound[c - 1], hi = m;
        while (lo < hi) { int mid = lo + (hi - lo) / 2; if (rowptr[mid] < target) lo = mid + 1; else hi = mid; }
        bound[c] = lo;
    }
    bound[chunks] = m;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int c = 0; c < chunks; ++c) {
        for (int i = bound[c]; i < bound[c + 1]; ++i) {
            double s = 0.0;
            for (ll k = rowptr[i]; k < rowptr[i + 1]; ++k) s += val[k] * x[col[k]];
            y[i] = alpha * s + beta * y[i];
        }
    }
    free(bound);
    return nt;
}

/* Error bound companion: bound[i] = |alpha| * sum_j |a_ij||x_j| + |beta||y_i|
 * (the denominator of the BASELINE.json tolerance). */
void oracle_csr_spmv_bound(int m, const ll *rowptr, const int *col, const double *val,
                           const double *x, double alpha, double beta, const double *y_in,
                           double *bound)
{
    for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (ll k = rowptr[i]; k < rowptr[i + 1]; ++k) s += fabs(val[k]) * fabs(x[col[k]]);
        bound[i] = fabs(alpha) * s + fabs(beta) * fabs(y_in[i]);
    }
}

/* ------------------------------------------------------------------------- */
/* a3: get_row_from_index, spmv/src/spmv_helper.cu:16-39.  Bisection over a[0..n]
 * that stops early on an exact hit; with equal neighbours (empty rows) it
 * returns whichever probe lands on the value -- reproduced. */
int oracle_get_row_from_index(int n, const ll *a, ll idx)
{
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        int probe = lo + (hi - lo) / 2;
        ll v = a[probe];
        if (v == idx) return probe;
        if (idx < v) hi = probe; else lo = probe;
    }
    if (a[lo] == idx) return lo;
    if (a[hi] == idx) return hi;
    return lo;
}

/* ------------------------------------------------------------------------- */
/* a1: baseline row-count split, spmv/src/dspmv_mgpu_baseline.cu:60-87. */
void oracle_partition_baseline(int m, const ll *rowptr, int ngpu,
                               int *start_row, int *end_row, int *dev_m, int *dev_nnz)
{
    for (int d = 0; d < ngpu; ++d) {
        start_row[d] = (d * m) / ngpu;                 /* :64, int arithmetic */
        end_row[d] = ((d + 1) * m) / ngpu - 1;         /* :65 */
        dev_m[d] = end_row[d] - start_row[d] + 1;      /* :67 */
        dev_nnz[d] = (int)(rowptr[end_row[d] + 1] - rowptr[start_row[d]]); /* :81 */
    }
}

/* local row pointer of a baseline shard, dspmv_mgpu_baseline.cu:82-85 */
void oracle_local_rowptr_baseline(const ll *rowptr, int start_row, int dev_m, int *local)
{
    for (int i = 0; i < dev_m + 1; ++i)
        local[i] = (int)(rowptr[start_row + i] - rowptr[start_row]);
}

/* a2: v1 nnz-balanced split, spmv/src/dspmv_mgpu_v1.cu:59-100,119.
 * flags are written as 0/1 ints. */
void oracle_partition_v1(int m, ll nnz, const ll *rowptr, int ngpu,
                         ll *start_idx, ll *end_idx, int *start_row, int *end_row,
                         int *start_flag, int *end_flag, int *dev_m, int *dev_nnz)
{
    for (int i = 0; i < ngpu; ++i) {
        ll t1 = (ll)i * nnz, t2 = (ll)(i + 1) * nnz;           /* :62-63 */
        start_idx[i] = (ll)floor((double)t1 / ngpu);             /* :68 */
        end_idx[i] = (ll)floor((double)t2 / ngpu) - 1;           /* :69 */
    }
    for (int i = 0; i < ngpu; ++i) {
        start_row[i] = oracle_get_row_from_index(m, rowptr, start_idx[i]); /* :74 */
        start_flag[i] = start_idx[i] > rowptr[start_row[i]];              /* :77 */
        end_row[i] = oracle_get_row_from_index(m, rowptr, end_idx[i]);     /* :86 */
        end_flag[i] = end_idx[i] < rowptr[end_row[i] + 1] - 1;            /* :89 */
        dev_m[i] = end_row[i] - start_row[i] + 1;                         /* :98 */
        dev_nnz[i] = (int)(end_idx[i] - start_idx[i] + 1);                /* :119 */
    }
}

/* local row pointer of a v1 shard / v2 task, dspmv_mgpu_v1.cu:125-133 and
 * dspmv_mgpu_v2.cu:279-289: [0]=0, [dev_m]=dev_nnz, middle rebased by start_idx. */
void oracle_local_rowptr_v1(const ll *rowptr, ll start_idx, int start_row, int dev_m,
                            int dev_nnz, int *local)
{
    local[0] = 0;
    local[dev_m] = dev_nnz;
    for (int j = 1; j < dev_m; ++j) local[j] = (int)(rowptr[start_row + j] - start_idx);
}

/* a4/a5: v2 task count after the memory clamp is applied by the caller,
 * dspmv_mgpu_v2.cu:218. */
int oracle_v2_num_tasks(ll nnz, ll nb) { return (int)((nnz + nb - 1) / nb); }

/* a5: generate_tasks, spmv/src/dspmv_mgpu_v2.cu:211-275.  Integer division happens
 * before the conversion to double (:235-236). */
void oracle_generate_tasks_v2(int m, ll nnz, const ll *rowptr, ll nb,
                              ll *start_idx, ll *end_idx, int *start_row, int *end_row,
                              int *start_flag, int *end_flag, int *dev_m, int *dev_nnz)
{
    int T = oracle_v2_num_tasks(nnz, nb);
    for (int t = 0; t < T; ++t) {
        ll t1 = (ll)t * nnz, t2 = (ll)(t + 1) * nnz;             /* :229-230 */
        start_idx[t] = (ll)floor((double)(t1 / T));                /* :235 */
        end_idx[t] = (ll)floor((double)(t2 / T)) - 1;              /* :236 */
        dev_nnz[t] = (int)(end_idx[t] - start_idx[t] + 1);         /* :237 */
    }
    for (int t = 0; t < T; ++t) {
        start_row[t] = oracle_get_row_from_index(m, rowptr, start_idx[t]); /* :244 */
        start_flag[t] = start_idx[t] > rowptr[start_row[t]];              /* :247 */
        end_row[t] = oracle_get_row_from_index(m, rowptr, end_idx[t]);     /* :257 */
        end_flag[t] = end_idx[t] < rowptr[end_row[t] + 1] - 1;            /* :261 */
        dev_m[t] = end_row[t] - start_row[t] + 1;                         /* :271 */
    }
}

/* a4: per-device task quota, dspmv_mgpu_v2.cu:125-126 */
int oracle_v2_quota(int T, int dev_id, int ngpu) { return T * (dev_id + 1) / ngpu - T * dev_id / ngpu; }

/* ------------------------------------------------------------------------- */
/* One shard's csrmv on the reference's own local arrays: y_local = alpha*A_d*x +
 * beta*y_slice (what each GPU computes, dspmv_mgpu_v1.cu:199-211). */
static void shard_csrmv(int dev_m, const int *local_rowptr, const int *col, const double *val,
                        const double *x, double alpha, double beta, double *y_local)
{
    for (int i = 0; i < dev_m; ++i) {
        double s = 0.0;
        for (int k = local_rowptr[i]; k < local_rowptr[i + 1]; ++k) s += val[k] * x[col[k]];
        y_local[i] = alpha * s + beta * y_local[i];
    }
}

/* The whole of spMV_mgpu_v1 on the CPU, including the ordered host merge of
 * split boundary rows (dspmv_mgpu_v1.cu:59-133 partition, :199-211 per-shard
 * csrmv, :235-248 merge: tmp = y[start_row]; overwrite; y += tmp; y -= y2*beta). */
int oracle_spmv_mgpu_v1(int m, int n, ll nnz, double alpha, const double *val, const ll *rowptr,
                        const int *col, const double *x, double beta, double *y, int ngpu)
{
    (void)n;
    ll *si = malloc(sizeof(ll) * ngpu), *ei = malloc(sizeof(ll) * ngpu);
    int *sr = malloc(sizeof(int) * ngpu), *er = malloc(sizeof(int) * ngpu);
    int *sf = malloc(sizeof(int) * ngpu), *ef = malloc(sizeof(int) * ngpu);
    int *dm = malloc(sizeof(int) * ngpu), *dz = malloc(sizeof(int) * ngpu);
    double *y2 = malloc(sizeof(double) * ngpu);
    double **yl = malloc(sizeof(double *) * ngpu);
    oracle_partition_v1(m, nnz, rowptr, ngpu, si, ei, sr, er, sf, ef, dm, dz);
    for (int d = 0; d < ngpu; ++d) if (sf[d]) y2[d] = y[sr[d]];           /* :79 */
    for (int d = 0; d < ngpu; ++d) {
        int *lp = malloc(sizeof(int) * (dm[d] + 1));
        oracle_local_rowptr_v1(rowptr, si[d], sr[d], dm[d], dz[d], lp);
        yl[d] = malloc(sizeof(double) * dm[d]);
        memcpy(yl[d], y + sr[d], sizeof(double) * dm[d]);                 /* :182, y uploaded before any merge */
        shard_csrmv(dm[d], lp, col + si[d], val + si[d], x, alpha, beta, yl[d]);
        free(lp);
    }
    for (int d = 0; d < ngpu; ++d) {                                      /* :235-248 */
        double tmp = 0.0;
        if (sf[d]) tmp = y[sr[d]];
        memcpy(y + sr[d], yl[d], sizeof(double) * dm[d]);
        if (sf[d]) { y[sr[d]] += tmp; y[sr[d]] -= y2[d] * beta; }
        free(yl[d]);
    }
    free(si); free(ei); free(sr); free(er); free(sf); free(ef); free(dm); free(dz); free(y2); free(yl);
    return 0;
}

/* spMV_mgpu_baseline on the CPU (dspmv_mgpu_baseline.cu:60-87,163-187): row
 * blocks, no shared rows, no merge. */
int oracle_spmv_mgpu_baseline(int m, int n, ll nnz, double alpha, const double *val, const ll *rowptr,
                              const int *col, const double *x, double beta, double *y, int ngpu)
{
    (void)n; (void)nnz;
    for (int d = 0; d < ngpu; ++d) {
        int sr = (d * m) / ngpu, er = ((d + 1) * m) / ngpu - 1, dm = er - sr + 1;
        int *lp = malloc(sizeof(int) * (dm + 1));
        oracle_local_rowptr_baseline(rowptr, sr, dm, lp);
        shard_csrmv(dm, lp, col + rowptr[sr], val + rowptr[sr], x, alpha, beta, y + sr);
        free(lp);
    }
    return 0;
}

/* spMV_mgpu_v2 on the CPU: generate_tasks (dspmv_mgpu_v2.cu:211-322), every task's
 * csrmv on the ORIGINAL y slice (assign_task uploads host_y before any merge,
 * :337-338), then gather_results (:385-441) in ascending task order (the
 * reference merges in completion order, which is timing dependent; ascending is
 * one legal order and the deterministic one).  y2 follows :249,263: end row's
 * original y if end_flag, else start row's if start_flag. */
int oracle_spmv_mgpu_v2_ex(int m, int n, ll nnz, double alpha, const double *val, const ll *rowptr,
                           const int *col, const double *x, double beta, double *y, ll nb, int faithful_y2);
int oracle_spmv_mgpu_v2(int m, int n, ll nnz, double alpha, const double *val, const ll *rowptr,
                        const int *col, const double *x, double beta, double *y, ll nb)
{
    return oracle_spmv_mgpu_v2_ex(m, n, nnz, alpha, val, rowptr, col, x, beta, y, nb, 0);
}

/* faithful_y2 != 0 reproduces a defect of the reference: struct spmv_task has ONE y2
 * field, written for the start row (:249) and then overwritten for the end row (:263),
 * so a task split at both ends subtracts beta*y[end_row] from its START row.  It is
 * invisible in the reference harness (y == 0 there, dspmv_test.cu:347-352) and wrong for
 * y != 0, beta != 0.  faithful_y2 == 0 keeps both originals (the intended arithmetic);
 * with y == 0 the two are bit-identical. */
int oracle_spmv_mgpu_v2_ex(int m, int n, ll nnz, double alpha, const double *val, const ll *rowptr,
                           const int *col, const double *x, double beta, double *y, ll nb, int faithful_y2)
{
    (void)n;
    if (nb <= 0) return -1;
    int T = oracle_v2_num_tasks(nnz, nb);
    ll *si = malloc(sizeof(ll) * T), *ei = malloc(sizeof(ll) * T);
    int *sr = malloc(sizeof(int) * T), *er = malloc(sizeof(int) * T);
    int *sf = malloc(sizeof(int) * T), *ef = malloc(sizeof(int) * T);
    int *dm = malloc(sizeof(int) * T), *dz = malloc(sizeof(int) * T);
    double *y2 = calloc(T, sizeof(double)), *y2e = calloc(T, sizeof(double));
    double **yl = malloc(sizeof(double *) * T);
    oracle_generate_tasks_v2(m, nnz, rowptr, nb, si, ei, sr, er, sf, ef, dm, dz);
    for (int t = 0; t < T; ++t) { if (sf[t]) y2[t] = y[sr[t]]; }
    for (int t = 0; t < T; ++t) { if (ef[t]) { y2e[t] = y[er[t]]; if (faithful_y2) y2[t] = y[er[t]]; } }
    for (int t = 0; t < T; ++t) {
        int *lp = malloc(sizeof(int) * (dm[t] + 1));
        oracle_local_rowptr_v1(rowptr, si[t], sr[t], dm[t], dz[t], lp);
        yl[t] = malloc(sizeof(double) * dm[t]);
        memcpy(yl[t], y + sr[t], sizeof(double) * dm[t]);
        shard_csrmv(dm[t], lp, col + si[t], val + si[t], x, alpha, beta, yl[t]);
        free(lp);
    }
    char *seen = calloc(m > 0 ? m : 1, 1);                                 /* :387-391 */
    for (int t = 0; t < T; ++t) {
        if (dm[t] == 1 && sf[t] && ef[t]) {                                /* :402-412 */
            if (!seen[sr[t]]) seen[sr[t]] = 1;
            else { double tmp = y[sr[t]]; yl[t][0] += tmp; yl[t][0] -= beta * y2[t]; }
        } else {
            if (sf[t]) {                                                   /* :415-423 */
                if (!seen[sr[t]]) seen[sr[t]] = 1;
                else { double tmp = y[sr[t]]; yl[t][0] += tmp; yl[t][0] -= beta * y2[t]; }
            }
            if (ef[t]) {                                                   /* :425-433 */
                if (!seen[er[t]]) seen[er[t]] = 1;
                else { double tmp = y[er[t]]; yl[t][dm[t] - 1] += tmp; yl[t][dm[t] - 1] -= beta * y2e[t]; }
            }
        }
        memcpy(y + sr[t], yl[t], sizeof(double) * dm[t]);                  /* :436-438 */
        free(yl[t]);
    }
    free(seen);
    free(si); free(ei); free(sr); free(er); free(sf); free(ef); free(dm); free(dz); free(y2); free(y2e); free(yl);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* a11: the `g` generator, spmv/test/dspmv_test.cu:137-208.  Per-row count is the
 * number of ints j>=0 with (double)j < n*r evaluated in double; r = 0.9 for the
 * first m/8 block, 0.01 after; values (double)rand()/RAND_MAX in generation
 * order; columns 0..k-1.  r1/r2 are parameters so the scaled shape (config 2b)
 * uses the same code.  Call once with val==NULL to get nnz. Uses glibc rand()
 * WITHOUT reseeding, like the harness. Returns nnz, or -1 if m is not a
 * positive multiple of 8 (the reference would write out of bounds). */
ll oracle_gen_g(int n, double r1, double r2, int *coo_row, int *coo_col, double *val)
{
    int m = n, blk = m / 8;
    if (blk <= 0 || m % 8 != 0) return -1;
    ll p = 0;
    for (int i = 0; i < m; i += blk) {
        double r = (i == 0) ? r1 : r2;
        for (int ii = i; ii < i + blk; ++ii) {
            for (int j = 0; j < n * r; ++j) {
                if (val) { coo_row[p] = ii; coo_col[p] = j; val[p] = (double)rand() / RAND_MAX; }
                ++p;
            }
        }
    }
    return p;
}

void oracle_srand(unsigned seed) { srand(seed); }
double oracle_rand_unit(void) { return (double)rand() / RAND_MAX; }   /* ALPHA/BETA, dspmv_test.cu:281-282 */

/* a11: COO -> "CSR" exactly like the harness (dspmv_test.cu:228-251): count per
 * row, prefix-sum; the COO col/val arrays are then used AS IS (unsorted, F3). */
void oracle_coo_to_rowptr(int m, ll nnz, const int *coo_row, ll *rowptr)
{
    int *cnt = calloc(m > 0 ? m : 1, sizeof(int));
    for (ll i = 0; i < nnz; ++i) cnt[coo_row[i]]++;
    rowptr[0] = 0;
    for (int i = 1; i <= m; ++i) rowptr[i] = rowptr[i - 1] + cnt[i - 1];
    free(cnt);
}

/* a11: the .mtx loader, dspmv_test.cu:101-136 (mm_read_banner + mm_read_mtx_crd_size
 * from spmv/include/mmio.h:254,339, then one fscanf per entry, indices made
 * 0-based, symmetric flag ignored).  mode 'f' reads "%d %d %lg", mode 'b' reads
 * "%d %d" and sets the value to 0.00001.  Two-pass: call with coo_row==NULL to
 * get m, n, nnz. Returns 0 on success. */
int oracle_load_mtx(const char *path, char mode, int *m, int *n, int *nnz,
                    int *coo_row, int *coo_col, double *val)
{
    FILE *f = fopen(path, "r");
    if (!f) return 1;
    char line[1025], a[64], b[64], c[64], d[64], e[64];
    if (!fgets(line, sizeof line, f)) { fclose(f); return 2; }
    if (sscanf(line, "%63s %63s %63s %63s %63s", a, b, c, d, e) != 5) { fclose(f); return 2; }
    if (strncmp(a, "%%MatrixMarket", 14) != 0) { fclose(f); return 3; }
    do { if (!fgets(line, sizeof line, f)) { fclose(f); return 2; } } while (line[0] == '%');
    while (sscanf(line, "%d %d %d", m, n, nnz) != 3) {
        if (!fgets(line, sizeof line, f)) { fclose(f); return 2; }
    }
    if (coo_row) {
        for (int i = 0; i < *nnz; ++i) {
            if (mode == 'b') { if (fscanf(f, "%d %d\n", &coo_row[i], &coo_col[i]) < 2) break; val[i] = 0.00001; }
            else { if (fscanf(f, "%d %d %lg\n", &coo_row[i], &coo_col[i], &val[i]) < 3) break; }
            coo_row[i]--; coo_col[i]--;
        }
    }
    fclose(f);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Helpers of bench.py's checker / CPU legs (not restatements of reference code). */

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU legs set their thread count explicitly
 * (all cores of the affinity mask) and report it.  Returns the value now in force. */
int oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* Full-vector check of a shard, all rows in parallel: want = alpha*A*x + beta*y_in (rows summed left
 * to right, as oracle_csr_spmv), err_i = |y_got[i] - want_i| / (|alpha| sum|a||x| + |beta||y_in[i]|).
 * rowptr is the shard's LOCAL row pointer (rowptr[0] = 0).  Rows skip_first / skip_last (shard-local
 * index, or -1) are rows split with another shard: left out of the maximum; their raw partial sums and
 * partial bounds are returned in edge[0..3] = {sum_first, bound_first, sum_last, bound_last} so that the
 * caller can finish them across shards.  Returns the worst err; *worst_row = its row. */
double oracle_csr_check(int m, const ll *rowptr, const int *col, const double *val, const double *x,
                        double alpha, double beta, const double *y_in, const double *y_got,
                        int skip_first, int skip_last, double *edge, int *worst_row)
{
    double worst = 0.0;
    int wrow = -1;
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        double lw = 0.0;
        int lr = -1;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4096) nowait
#endif
        for (int i = 0; i < m; ++i) {
            double s = 0.0, b = 0.0;
            for (ll k = rowptr[i]; k < rowptr[i + 1]; ++k) {
                const double p = val[k] * x[col[k]];
                s += p;
                b += fabs(p);
            }
            if (i == skip_first) { edge[0] = s; edge[1] = b; continue; }
            if (i == skip_last) { edge[2] = s; edge[3] = b; continue; }
            const double want = alpha * s + beta * y_in[i];
            const double bound = fabs(alpha) * b + fabs(beta) * fabs(y_in[i]);
            const double d = fabs(y_got[i] - want);
            const double e = bound > 0.0 ? d / bound : (d == 0.0 ? 0.0 : INFINITY);
            if (!(e <= lw)) { lw = e; lr = i; }         /* NaN counts as worst */
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        {
            if (!(lw <= worst)) { worst = lw; wrow = lr; }
            else if (wrow < 0) wrow = lr;
        }
    }
    if (worst_row) *worst_row = wrow;
    return worst;
}

/* The synthetic matrices of bench.py on the HOST: the same hash generator as the GPU one
 * (s-blas_b200/csrc/sblas_synth.cu: splitmix64 of (seed, entry index); column patterns PREFIX 0,
 * BANDED 1, UNIFORM 2, CIRCUIT 3, BANDRUN 4), restated here so that the reference arm can build the
 * workload without a GPU.  tests/ checks the two generators agree entry for entry. */
static inline unsigned long long o_mix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double o_u01(unsigned long long h) { return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

void oracle_synth_fill_csr(const ll *rp, int row_first, int nrows, ll k0, ll k1, int n, int mode, ll band,
                           unsigned long long seed, int vmode, double vconst, double *val, int *col)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1024)
#endif
    for (int i = 0; i < nrows; ++i) {
        const ll b = rp[i], e = rp[i + 1];
        if (e <= k0 || b >= k1) continue;
        const ll len = e - b, row = (ll)row_first + i;
        int m = mode;
        if (m == 3) {
            const int hub = (o_mix64(seed ^ (unsigned long long)row * 0x51ull) % 5u) == 0u;
            m = (hub || len > 2 * band) ? 2 : 1;
        }
        ll W = n, start = 0;
        const int runs = (m == 4);
        if (runs) m = 1;
        if (m == 1) {
            W = 2 * band < n ? 2 * band : n;
            if (W < len) W = len < n ? len : n;
            start = row - W / 2;
            if (start < 0) start = 0;
            if (start + W > n) start = n - W;
        }
        const ll jb = (b > k0 ? b : k0) - b, je = (e < k1 ? e : k1) - b;
        for (ll j = jb; j < je; ++j) {
            const ll k = b + j;
            const unsigned long long h = o_mix64(seed ^ (unsigned long long)k);
            int c;
            if (m == 0) c = (int)(j < n ? j : n - 1);
            else if (runs && len <= W / 16) {
                const ll nrun = (len + 15) / 16, r = j / 16;
                const unsigned long long hr = o_mix64(seed ^ (unsigned long long)(b + r * 16) * 0x9E37ull);
                const ll cell = (W / 16) / nrun;
                const ll slot = r * cell + (ll)(hr % (unsigned long long)cell);
                const ll cc = start + slot * 16 + (j - r * 16);
                c = (int)(cc < n ? cc : n - 1);
            } else if (len <= W) {
                const ll lo = (ll)(((__int128)j * W) / len), hi = (ll)(((__int128)(j + 1) * W) / len);
                const ll span = hi - lo > 0 ? hi - lo : 1;
                c = (int)(start + lo + (ll)(h % (unsigned long long)span));
            } else c = (int)(j % n);
            col[k - k0] = c;
            val[k - k0] = vmode ? vconst : o_u01(o_mix64(h));
        }
    }
}

void oracle_synth_fill_uniform(double *p, ll count, unsigned long long seed, double lo, double hi)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (ll i = 0; i < count; ++i)
        p[i] = lo + (hi - lo) * o_u01(o_mix64(seed ^ (unsigned long long)i * 0x2545F4914F6CDD1Dull));
}

/* ------------------------------------------------------------------------- */
/* SpMM.  The arithmetic of cusparseDcsrmm as the reference calls it
 * (spmm/src/dspmm_mgpu_baseline.cu:225-241, :450-466): C = alpha*A*B + beta*C, A CSR with an int32 row
 * pointer (base 0), B (k x n, leading dimension ldb) and C (m x n, ldc) dense COLUMN-major; every
 * C(i,c) summed left to right over row i.  Columns in parallel over the host cores. */
void oracle_csrmm(int m, int n, const int *rowptr, const int *col, const double *val, const double *B, ll ldb,
                  double *C, ll ldc, double alpha, double beta)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int c = 0; c < n; ++c) {
        const double *b = B + (ll)c * ldb;
        double *out = C + (ll)c * ldc;
        for (int i = 0; i < m; ++i) {
            double s = 0.0;
            for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) s += val[k] * b[col[k]];
            out[i] = alpha * s + beta * out[i];
        }
    }
}

/* bound(i,c) = |alpha| sum_j |a_ij||b_jc| + |beta||c_ic|: the denominator of the per-entry tolerance */
void oracle_csrmm_bound(int m, int n, const int *rowptr, const int *col, const double *val, const double *B, ll ldb,
                        const double *C_in, ll ldc, double alpha, double beta, double *bound)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int c = 0; c < n; ++c) {
        const double *b = B + (ll)c * ldb;
        for (int i = 0; i < m; ++i) {
            double s = 0.0;
            for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) s += fabs(val[k]) * fabs(b[col[k]]);
            bound[(ll)c * m + i] = fabs(alpha) * s + fabs(beta) * fabs(C_in[(ll)c * ldc + i]);
        }
    }
}

/* the whole entry point, cusparse_mgpu_csrmm[_omp] (dspmm_mgpu_baseline.cu:83-280, :282-524): columns of B
 * and C split over the GPUs, dev_n[d] = floor((d+1)n/ngpu) - floor(dn/ngpu), offsets floor(dn/ngpu)*k and *m
 * (:338-342; C int arithmetic, the floor() of an int is a no-op), one csrmm per GPU on its slice */
int oracle_spmm_mgpu(int m, int n, int k, double alpha, const int *rowptr, const int *col, const double *val,
                     double beta, const double *B, double *C, int ngpu)
{
    if (ngpu <= 0) return -1;
    for (int d = 0; d < ngpu; ++d) {
        const int c0 = (int)((ll)d * n / ngpu), c1 = (int)((ll)(d + 1) * n / ngpu);
        oracle_csrmm(m, c1 - c0, rowptr, col, val, B + (ll)c0 * k, k, C + (ll)c0 * m, m, alpha, beta);
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* sparse transposition CSR -> CSC, restating the reference's host routine
 * sptrans/sptrans_v1/src/tranpose.h:3-40 (matrix_transposition): histogram of the column indices,
 * exclusive scan (utils.h:312-329), then the rows in order, every entry appended to its column -- so a
 * column keeps the CSR order of its entries (rows ascending, duplicates in input order). */
void oracle_csr2csc(int m, int n, int nnz, const int *csrRowPtr, const int *csrColIdx, const double *csrVal,
                    int *cscRowIdx, int *cscColPtr, double *cscVal)
{
    memset(cscColPtr, 0, sizeof(int) * ((size_t)n + 1));
    for (int i = 0; i < nnz; ++i) cscColPtr[csrColIdx[i]]++;
    int run = 0;                                        /* in-place exclusive scan over n + 1 entries */
    for (int c = 0; c <= n; ++c) { const int t = cscColPtr[c]; cscColPtr[c] = run; run += t; }
    int *incr = (int *)malloc(sizeof(int) * ((size_t)n + 1));
    memcpy(incr, cscColPtr, sizeof(int) * ((size_t)n + 1));
    for (int row = 0; row < m; ++row)
        for (int j = csrRowPtr[row]; j < csrRowPtr[row + 1]; ++j) {
            const int c = csrColIdx[j];
            cscRowIdx[incr[c]] = row;
            cscVal[incr[c]] = csrVal[j];
            incr[c]++;
        }
    free(incr);
}
