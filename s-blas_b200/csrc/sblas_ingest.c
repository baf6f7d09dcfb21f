/* sblas_ingest.c -- correct Matrix-Market -> CSR ingest (opt-in; see include/sblas_ingest.h).
 * Semantics of sptrsv/sptrsv_v1/src/mmio_highlevel.h:8-298 of the reference, written from its
 * description: one pass over the file into COO, per-row counts (+ mirrored counts for symmetric
 * files), exclusive scan, stable scatter in file order.  64-bit nnz / row pointer. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include "sblas_ingest.h"

typedef struct mtx_head {
    int m, n;
    long long listed;          /* entries in the file */
    int pattern, integer, complex_, symmetric;
} mtx_head;

static void lower(char *s) { for (; *s; ++s) *s = (char)tolower((unsigned char)*s); }

/* banner "%%MatrixMarket matrix coordinate <field> <symmetry>", comments, then "m n nnz" */
static int read_head(FILE *f, mtx_head *h)
{
    char line[1025], t0[64], t1[64], t2[64], t3[64], t4[64];
    memset(h, 0, sizeof *h);
    if (!fgets(line, sizeof line, f)) return -2;
    if (sscanf(line, "%63s %63s %63s %63s %63s", t0, t1, t2, t3, t4) != 5) return -2;
    if (strncmp(t0, "%%MatrixMarket", 14) != 0) return -2;
    lower(t1); lower(t2); lower(t3); lower(t4);
    if (strcmp(t1, "matrix") != 0 || strcmp(t2, "coordinate") != 0) return -2;
    h->pattern = !strcmp(t3, "pattern");
    h->integer = !strcmp(t3, "integer");
    h->complex_ = !strcmp(t3, "complex");
    if (!h->pattern && !h->integer && !h->complex_ && strcmp(t3, "real") != 0) return -2;
    h->symmetric = !strcmp(t4, "symmetric") || !strcmp(t4, "hermitian");
    do {
        if (!fgets(line, sizeof line, f)) return -4;
    } while (line[0] == '%');
    for (;;) {
        if (sscanf(line, "%d %d %lld", &h->m, &h->n, &h->listed) == 3) break;
        if (!fgets(line, sizeof line, f)) return -4;
    }
    if (h->m <= 0 || h->n <= 0 || h->listed < 0) return -4;
    /* a symmetric / hermitian file must be square: the mirrored entry (j, i) is written into row j,
     * which does not exist when n > m (untrusted input: reject instead of writing past the row pointer) */
    if (h->symmetric && h->m != h->n) return -7;
    return 0;
}

/* one entry; 0-based on return */
static int read_entry(FILE *f, const mtx_head *h, int *i, int *j, double *v)
{
    double im;
    int got;
    if (h->pattern) { got = fscanf(f, "%d %d", i, j); *v = 1.0; if (got != 2) return -5; }
    else if (h->complex_) { got = fscanf(f, "%d %d %lg %lg", i, j, v, &im); if (got != 4) return -5; }
    else { got = fscanf(f, "%d %d %lg", i, j, v); if (got != 3) return -5; }     /* real and integer */
    --*i; --*j;
    if (*i < 0 || *i >= h->m || *j < 0 || *j >= h->n) return -5;
    return 0;
}

int sblas_mtx_info(const char *path, int *m, int *n, long long *nnz, int *is_symmetric)
{
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    mtx_head h;
    int rc = read_head(f, &h);
    long long total = 0;
    for (long long e = 0; rc == 0 && e < h.listed; ++e) {
        int i, j;
        double v;
        rc = read_entry(f, &h, &i, &j, &v);
        if (rc == 0) total += (h.symmetric && i != j) ? 2 : 1;
    }
    fclose(f);
    if (rc != 0) return rc;
    *m = h.m; *n = h.n; *nnz = total;
    if (is_symmetric) *is_symmetric = h.symmetric;
    return 0;
}

int sblas_mtx_read_csr(const char *path, long long *rp, int *col, double *val)
{
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    mtx_head h;
    int rc = read_head(f, &h);
    if (rc != 0) { fclose(f); return rc; }
    /* untrusted size line: an entry count whose byte size wraps size_t would turn into a short allocation */
    if ((unsigned long long)h.listed > (unsigned long long)((size_t)-1) / (4 * sizeof(double))) { fclose(f); return -6; }
    int *ci = (int *)malloc((size_t)(h.listed ? h.listed : 1) * sizeof(int));
    int *cj = (int *)malloc((size_t)(h.listed ? h.listed : 1) * sizeof(int));
    double *cv = (double *)malloc((size_t)(h.listed ? h.listed : 1) * sizeof(double));
    long long *fill = (long long *)calloc((size_t)h.m + 1, sizeof(long long));
    if (!ci || !cj || !cv || !fill) { rc = -6; goto done; }
    for (int r = 0; r <= h.m; ++r) rp[r] = 0;
    for (long long e = 0; e < h.listed; ++e) {
        rc = read_entry(f, &h, &ci[e], &cj[e], &cv[e]);
        if (rc != 0) goto done;
        rp[ci[e] + 1]++;
        if (h.symmetric && ci[e] != cj[e]) rp[cj[e] + 1]++;        /* the mirrored entry */
    }
    for (int r = 0; r < h.m; ++r) rp[r + 1] += rp[r];
    /* stable scatter in file order; a mirrored entry follows its original (mmio_highlevel.h:254-266) */
    for (long long e = 0; e < h.listed; ++e) {
        long long o = rp[ci[e]] + fill[ci[e]]++;
        col[o] = cj[e]; val[o] = cv[e];
        if (h.symmetric && ci[e] != cj[e]) {
            o = rp[cj[e]] + fill[cj[e]]++;
            col[o] = ci[e]; val[o] = cv[e];
        }
    }
done:
    fclose(f);
    free(ci); free(cj); free(cv); free(fill);
    return rc;
}
