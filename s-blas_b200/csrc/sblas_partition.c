/* sblas_partition.c -- the three partitioners of the reference SpMV path, host C,
 * integer-exact.  No CUDA here.
 *
 *   baseline  equal row counts          spmv/src/dspmv_mgpu_baseline.cu:60-87
 *   v1        equal nnz, rows may split spmv/src/dspmv_mgpu_v1.cu:59-133
 *   v2        nb-sized task tiles       spmv/src/dspmv_mgpu_v2.cu:211-289, quota :125-126
 *   row lookup                          spmv/src/spmv_helper.cu:16-39
 */
#include <math.h>
#include <stddef.h>
#include "sblas_spmv.h"

/* Row lookup used by v1/v2.  The reference bisects rowptr[0..n] and returns as
 * soon as a probe equals idx; next to empty rows that can name a neighbouring
 * (empty) row instead of the row that holds idx.  The partition
 * records stay bit-exact with that behaviour; the plan separately resolves the
 * row that really holds the entry (sblas_plan.c: true_row_of). */
int sblas_get_row_from_index(int n, const long long *a, long long idx)
{
    int left = 0, right = n;
    while (right - left > 1) {
        const int mid = left + (right - left) / 2;
        if (a[mid] == idx) return mid;
        if (a[mid] > idx) right = mid; else left = mid;
    }
    if (a[left] == idx) return left;
    if (a[right] == idx) return right;
    return left;
}

int sblas_partition_baseline(int m, const long long *rp, int ngpu, sblas_part *out)
{
    if (ngpu <= 0 || m < 0) return -1;
    for (int d = 0; d < ngpu; ++d) {
        sblas_part *p = &out[d];
        p->start_row = (int)(((long long)d * m) / ngpu);
        p->end_row = (int)(((long long)(d + 1) * m) / ngpu) - 1;
        p->dev_m = p->end_row - p->start_row + 1;
        p->start_idx = rp[p->start_row];
        p->end_idx = rp[p->end_row + 1] - 1;
        p->dev_nnz = (int)(rp[p->end_row + 1] - rp[p->start_row]);
        p->start_flag = 0;
        p->end_flag = 0;
    }
    return 0;
}

static void rows_and_flags(int m, const long long *rp, sblas_part *p)
{
    p->start_row = sblas_get_row_from_index(m, rp, p->start_idx);
    p->start_flag = p->start_idx > rp[p->start_row];
    p->end_row = sblas_get_row_from_index(m, rp, p->end_idx);
    p->end_flag = p->end_idx < rp[p->end_row + 1] - 1;
    p->dev_m = p->end_row - p->start_row + 1;
    p->dev_nnz = (int)(p->end_idx - p->start_idx + 1);
}

int sblas_partition_v1(int m, long long nnz, const long long *rp, int ngpu, sblas_part *out)
{
    if (ngpu <= 0 || m <= 0) return -1;
    for (int i = 0; i < ngpu; ++i) {
        /* the reference divides in double precision and floors */
        out[i].start_idx = (long long)floor((double)((long long)i * nnz) / ngpu);
        out[i].end_idx = (long long)floor((double)((long long)(i + 1) * nnz) / ngpu) - 1;
        rows_and_flags(m, rp, &out[i]);
    }
    return 0;
}

/* Opt-in fourth version (NOT in the reference): the same contiguous nnz ranges with split rows as v1, but cut so
 * that every GPU gets the same share of the BYTES one product streams, 12 per entry + row_bytes per row (row
 * pointer, y read and write, x amortised), instead of the same share of the entries.  v1 leaves the GPU that gets
 * the short rows with far more rows -- and bytes -- than the others (the 50M-row config at 8 GPUs: 2.7 GB on the
 * last GPU against 1.8 GB elsewhere).  Weight of the entries before index idx: W(idx) = 12*idx + row_bytes*row(idx);
 * boundary i = the smallest idx with W(idx) >= i * W(nnz) / ngpu. */
static int row_of_idx(int m, const long long *rp, long long idx)
{
    int lo = 0, hi = m;                       /* first r in [0,m] with rp[r] > idx, minus one */
    while (lo < hi) {
        const int mid = lo + (hi - lo) / 2;
        if (rp[mid] <= idx) lo = mid + 1; else hi = mid;
    }
    return lo - 1 < 0 ? 0 : lo - 1;
}

int sblas_partition_bytes(int m, long long nnz, const long long *rp, int ngpu, int row_bytes, sblas_part *out)
{
    if (ngpu <= 0 || m <= 0 || row_bytes < 0) return -1;
    const double total = 12.0 * (double)nnz + (double)row_bytes * m;
    long long prev = 0;
    for (int i = 0; i < ngpu; ++i) {
        long long next = nnz;
        if (i + 1 < ngpu) {
            const double target = total * (i + 1) / ngpu;
            long long lo = prev, hi = nnz;                      /* smallest idx with W(idx) >= target */
            while (lo < hi) {
                const long long mid = lo + (hi - lo) / 2;
                const double w = 12.0 * (double)mid + (double)row_bytes * row_of_idx(m, rp, mid);
                if (w >= target) hi = mid; else lo = mid + 1;
            }
            next = lo;
        }
        out[i].start_idx = prev;
        out[i].end_idx = next - 1;
        if (next > prev) rows_and_flags(m, rp, &out[i]);
        else { out[i].start_row = out[i].end_row = 0; out[i].start_flag = out[i].end_flag = 0; out[i].dev_m = 0; out[i].dev_nnz = 0; }
        prev = next;
    }
    return 0;
}

int sblas_v2_num_tasks(long long nnz, long long nb)
{
    if (nb <= 0) return 0;
    return (int)((nnz + nb - 1) / nb);
}

int sblas_generate_tasks_v2(int m, long long nnz, const long long *rp, long long nb, sblas_part *out)
{
    const int T = sblas_v2_num_tasks(nnz, nb);
    if (T <= 0 || m <= 0) return -1;
    for (int t = 0; t < T; ++t) {
        /* integer division first, then the (no-op) floor of the reference */
        out[t].start_idx = ((long long)t * nnz) / T;
        out[t].end_idx = ((long long)(t + 1) * nnz) / T - 1;
        rows_and_flags(m, rp, &out[t]);
    }
    return T;
}

int sblas_v2_task_owner(int T, int ngpu, int task)
{
    /* GPU d owns [T*d/ngpu, T*(d+1)/ngpu): size == the reference quota */
    int lo = 0, hi = ngpu - 1;
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if ((long long)T * (mid + 1) / ngpu > task) hi = mid; else lo = mid + 1;
    }
    return lo;
}

void sblas_local_rowptr(const long long *rp, const sblas_part *p, int baseline, int *out)
{
    if (baseline) {
        for (int i = 0; i <= p->dev_m; ++i) out[i] = (int)(rp[p->start_row + i] - rp[p->start_row]);
        return;
    }
    out[0] = 0;
    out[p->dev_m] = p->dev_nnz;
    for (int j = 1; j < p->dev_m; ++j) out[j] = (int)(rp[p->start_row + j] - p->start_idx);
}
