/* sblas_spmm_plan.c -- host side (C) of the multi-GPU CSR SpMM, C = alpha*A*B + beta*C
 * (SURVEY.md section 8f-2).  Replaces the bodies of cusparse_mgpu_csrmm / cusparse_mgpu_csrmm_omp
 * (spmm/src/dspmm_mgpu_baseline.cu:282-524 and :83-280) with a plan / execute split:
 *
 *   plan    = A resident on every GPU of the plan (the reference uploads all of A to every GPU on every
 *             call, :391-409): each GPU takes a 1/ngpu slice of col / val over its own PCIe link and the
 *             slices are exchanged over NVLink; the list of rows too long for a warp is built once.
 *   execute = columns of B and C split over the GPUs exactly like the reference,
 *             dev_n[d] = floor((d+1)n/ngpu) - floor(dn/ngpu)  (:338-342): per GPU one H2D copy of its
 *             (contiguous, column-major) slice of B and, when beta != 0, of C; B is transposed on the
 *             device (sblas_spmm.cu), the SpMM kernels run, the slice of C comes back.  All GPUs work
 *             concurrently on their own streams; one host thread drives them (the _omp variant of the
 *             reference spawns a thread per GPU for the same work).
 * No cuSPARSE, no CPU arithmetic, no CPU fallback.
 */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sblas_internal.h"
#include "sblas_spmm.h"
#include "spmm_kernel.h"

cudaError_t sblas_launch_transpose_b(const double *d_B, long long ldb, int k, int nd, double *d_Bt, cudaStream_t s);
cudaError_t sblas_launch_spmm(int m, int nd, const int *rowptr, const int *col, const double *val, const double *d_Bt,
                              double *d_C, long long ldc, double alpha, double beta, const int *long_rows,
                              const int *row_seg, int nlong, const int *seg_lo, const int *seg_hi, int nseg, double *part,
                              int long_thr, cudaStream_t s);
int sblas_spmm_segment_length(void);
long long sblas_spmm_bt_pitch(int nd);
int sblas_spmm_long_row_threshold(int m);

typedef struct spmm_dev {
    int device;
    int *d_rowptr, *d_col, *d_long, *d_row_seg, *d_seg_lo, *d_seg_hi;
    double *d_val;
    double *d_B, *d_Bt, *d_C, *d_part; /* work buffers, grown on demand (B, C: host-pointer execute only) */
    size_t cap_B, cap_Bt, cap_C, cap_part;
    cudaStream_t stream;
    cudaEvent_t ev_slice;
} spmm_dev;

struct sblas_spmm_plan {
    int m, k, nnz, ndev, nlong, nseg, long_thr;
    spmm_dev *devs;
};

#define CU(call)                                                                      \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            sblas_set_error("%s failed: %s (sblas_spmm_plan.c:%d)", #call, cudaGetErrorString(e_), __LINE__); \
            rc = 1;                                                                   \
            goto fail;                                                                \
        }                                                                             \
    } while (0)

int sblas_spmm_plan_num_devices(const sblas_spmm_plan *P) { return P->ndev; }
void *sblas_spmm_plan_stream(sblas_spmm_plan *P, int dev) { return (void *)P->devs[dev].stream; }

int sblas_spmm_plan_columns(const sblas_spmm_plan *P, int n, int dev, int *first, int *count)
{
    if (dev < 0 || dev >= P->ndev || n < 0) return -1;
    /* dspmm_mgpu_baseline.cu:338-342: integer arithmetic, floor() of an int is a no-op */
    const int lo = (int)((long long)dev * n / P->ndev), hi = (int)((long long)(dev + 1) * n / P->ndev);
    if (first) *first = lo;
    if (count) *count = hi - lo;
    return 0;
}

void sblas_spmm_plan_destroy(sblas_spmm_plan *P)
{
    if (!P) return;
    for (int d = 0; d < P->ndev; ++d) {
        spmm_dev *D = &P->devs[d];
        if (D->device < 0) continue;
        cudaSetDevice(D->device);
        cudaFree(D->d_rowptr); cudaFree(D->d_col); cudaFree(D->d_val); cudaFree(D->d_long);
        cudaFree(D->d_row_seg); cudaFree(D->d_seg_lo); cudaFree(D->d_seg_hi);
        cudaFree(D->d_B); cudaFree(D->d_Bt); cudaFree(D->d_C); cudaFree(D->d_part);
        if (D->stream) cudaStreamDestroy(D->stream);
        if (D->ev_slice) cudaEventDestroy(D->ev_slice);
    }
    free(P->devs);
    free(P);
}

int sblas_spmm_plan_create(sblas_spmm_plan **out, int m, int k, int nnz, const int *rp, const int *col,
                           const double *val, int ngpu)
{
    int rc = 0, count = 0;
    int *h_long = NULL;
    if (!out || m <= 0 || k <= 0 || nnz < 0 || ngpu <= 0) {
        sblas_set_error("%s%s (line %d)", "invalid argument", "", __LINE__);
        return -1;
    }
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < ngpu) {
        cudaGetLastError();
        sblas_set_error("%s%s (line %d)", "not enough CUDA devices (no CPU fallback)", "", __LINE__);
        return 1;
    }
    sblas_spmm_plan *P = (sblas_spmm_plan *)calloc(1, sizeof *P);
    if (!P) return 1;
    P->m = m; P->k = k; P->nnz = nnz; P->ndev = ngpu;
    P->devs = (spmm_dev *)calloc((size_t)ngpu, sizeof(spmm_dev));
    if (!P->devs) { free(P); return 1; }
    for (int d = 0; d < ngpu; ++d) P->devs[d].device = -1;
    *out = P;

    /* rows a single warp should not take alone, cut into segments of at most `seglen` entries */
    const int thr = sblas_spmm_long_row_threshold(m), seglen = sblas_spmm_segment_length();
    P->long_thr = thr;
    int nlong = 0, nseg = 0;
    for (int r = 0; r < m; ++r) {
        const int len = rp[r + 1] - rp[r];
        if (len > thr) { ++nlong; nseg += (len + seglen - 1) / seglen; }
    }
    P->nlong = nlong; P->nseg = nseg;
    if (nlong > 0) {
        h_long = (int *)malloc(((size_t)2 * nlong + 1 + 2 * (size_t)nseg) * sizeof(int));
        if (!h_long) { rc = 1; goto fail; }
        int *h_row_seg = h_long + nlong, *h_lo = h_row_seg + nlong + 1, *h_hi = h_lo + nseg;
        int w = 0, g = 0;
        for (int r = 0; r < m; ++r) {
            const int len = rp[r + 1] - rp[r];
            if (len <= thr) continue;
            h_long[w] = r; h_row_seg[w] = g; ++w;
            for (int b0 = rp[r]; b0 < rp[r + 1]; b0 += seglen) { h_lo[g] = b0; h_hi[g] = b0 + seglen < rp[r + 1] ? b0 + seglen : rp[r + 1]; ++g; }
        }
        h_row_seg[nlong] = g;
    }

    /* memory guard of the reference (dspmm_mgpu_baseline.cu:328-336) is applied per product, where n is known */
    int p2p = ngpu > 1;
    for (int a = 0; a < ngpu && p2p; ++a)
        for (int b = 0; b < ngpu; ++b) {
            int can = 0;
            if (a != b) { cudaDeviceCanAccessPeer(&can, a, b); if (!can) { p2p = 0; break; } }
        }
    for (int d = 0; d < ngpu; ++d) {
        spmm_dev *D = &P->devs[d];
        CU(cudaSetDevice(d));
        D->device = d;
        if (p2p)
            for (int b = 0; b < ngpu; ++b)
                if (b != d) { cudaError_t e = cudaDeviceEnablePeerAccess(b, 0); if (e != cudaSuccess) cudaGetLastError(); }
        CU(cudaStreamCreateWithFlags(&D->stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&D->ev_slice, cudaEventDisableTiming));
        CU(cudaMalloc((void **)&D->d_rowptr, ((size_t)m + 1) * sizeof(int)));
        CU(cudaMalloc((void **)&D->d_col, ((size_t)nnz + 4) * sizeof(int)));
        CU(cudaMalloc((void **)&D->d_val, ((size_t)nnz + 4) * sizeof(double)));
        CU(cudaMemcpyAsync(D->d_rowptr, rp, ((size_t)m + 1) * sizeof(int), cudaMemcpyHostToDevice, D->stream));
        if (nlong > 0) {
            CU(cudaMalloc((void **)&D->d_long, (size_t)nlong * sizeof(int)));
            CU(cudaMalloc((void **)&D->d_row_seg, ((size_t)nlong + 1) * sizeof(int)));
            CU(cudaMalloc((void **)&D->d_seg_lo, (size_t)nseg * sizeof(int)));
            CU(cudaMalloc((void **)&D->d_seg_hi, (size_t)nseg * sizeof(int)));
            CU(cudaMemcpyAsync(D->d_long, h_long, (size_t)nlong * sizeof(int), cudaMemcpyHostToDevice, D->stream));
            CU(cudaMemcpyAsync(D->d_row_seg, h_long + nlong, ((size_t)nlong + 1) * sizeof(int), cudaMemcpyHostToDevice, D->stream));
            CU(cudaMemcpyAsync(D->d_seg_lo, h_long + 2 * nlong + 1, (size_t)nseg * sizeof(int), cudaMemcpyHostToDevice, D->stream));
            CU(cudaMemcpyAsync(D->d_seg_hi, h_long + 2 * nlong + 1 + nseg, (size_t)nseg * sizeof(int), cudaMemcpyHostToDevice, D->stream));
        }
        /* this GPU's slice of A over its own PCIe link (everything, when peers cannot be reached) */
        const long long lo = p2p ? (long long)nnz * d / ngpu : 0, hi = p2p ? (long long)nnz * (d + 1) / ngpu : nnz;
        if (hi > lo) {
            CU(cudaMemcpyAsync(D->d_col + lo, col + lo, (size_t)(hi - lo) * sizeof(int), cudaMemcpyHostToDevice, D->stream));
            CU(cudaMemcpyAsync(D->d_val + lo, val + lo, (size_t)(hi - lo) * sizeof(double), cudaMemcpyHostToDevice, D->stream));
        }
        CU(cudaEventRecord(D->ev_slice, D->stream));
    }
    if (p2p) {              /* all-gather of the slices over NVLink */
        for (int d = 0; d < ngpu; ++d) {
            spmm_dev *D = &P->devs[d];
            CU(cudaSetDevice(d));
            for (int o = 0; o < ngpu; ++o) {
                if (o == d) continue;
                const long long lo = (long long)nnz * o / ngpu, hi = (long long)nnz * (o + 1) / ngpu;
                if (hi <= lo) continue;
                CU(cudaStreamWaitEvent(D->stream, P->devs[o].ev_slice, 0));
                CU(cudaMemcpyPeerAsync(D->d_col + lo, d, P->devs[o].d_col + lo, o, (size_t)(hi - lo) * sizeof(int), D->stream));
                CU(cudaMemcpyPeerAsync(D->d_val + lo, d, P->devs[o].d_val + lo, o, (size_t)(hi - lo) * sizeof(double), D->stream));
            }
        }
    }
    for (int d = 0; d < ngpu; ++d) {
        CU(cudaSetDevice(d));
        CU(cudaStreamSynchronize(P->devs[d].stream));
    }
    free(h_long);
    return 0;
fail:
    free(h_long);
    sblas_spmm_plan_destroy(P);
    *out = NULL;
    return rc;
}

static int grow(void **p, size_t *cap, size_t need)
{
    if (*cap >= need) return 0;
    if (*p) cudaFree(*p);
    *p = NULL; *cap = 0;
    if (cudaMalloc(p, need) != cudaSuccess) { cudaGetLastError(); return 1; }
    *cap = need;
    return 0;
}

int sblas_spmm_plan_execute_device(sblas_spmm_plan *P, int dev, int nd, double alpha, const double *d_B, double beta,
                                   double *d_C, int sync)
{
    int rc = 0;
    if (dev < 0 || dev >= P->ndev || nd < 0) return -1;
    if (nd == 0) return 0;
    spmm_dev *D = &P->devs[dev];
    CU(cudaSetDevice(D->device));
    const size_t bt_bytes = (size_t)P->k * (size_t)sblas_spmm_bt_pitch(nd) * sizeof(double);
    if (grow((void **)&D->d_Bt, &D->cap_Bt, bt_bytes)) { sblas_set_error("%s%s (line %d)", "cudaMalloc(Bt) failed", "", __LINE__); return 1; }
    if (P->nseg > 0 && grow((void **)&D->d_part, &D->cap_part, (size_t)P->nseg * (size_t)sblas_spmm_bt_pitch(nd) * sizeof(double))) {
        sblas_set_error("%s%s (line %d)", "cudaMalloc(segment partials) failed", "", __LINE__);
        return 1;
    }
    CU(sblas_launch_transpose_b(d_B, P->k, P->k, nd, D->d_Bt, D->stream));
    CU(sblas_launch_spmm(P->m, nd, D->d_rowptr, D->d_col, D->d_val, D->d_Bt, d_C, P->m, alpha, beta, D->d_long, D->d_row_seg,
                         P->nlong, D->d_seg_lo, D->d_seg_hi, P->nseg, D->d_part, P->long_thr, D->stream));
    if (sync) CU(cudaStreamSynchronize(D->stream));
fail:
    return rc;
}

int sblas_spmm_plan_execute(sblas_spmm_plan *P, int n, const double *alpha, const double *B, const double *beta, double *C)
{
    int rc = 0;
    if (n < 0) return -1;
    /* the reference's guard (dspmm_mgpu_baseline.cu:328-336: 1.2 x (A + B/ngpu + C/ngpu) must fit the free
     * device memory) for what this product still has to allocate: the B slice twice (as given and transposed)
     * and the C slice; A is resident already */
    {
        const spmm_dev *D0 = &P->devs[0];
        const double need = 1.2e-9 * 8.0 * (2.0 * (double)P->k * n + (double)P->m * n) / P->ndev;
        const double have = sblas_get_gpu_availble_mem(P->ndev) + 1e-9 * (double)(D0->cap_B + D0->cap_Bt + D0->cap_C);
        if (need > have) {
            sblas_set_error("%s%s (line %d)", "No available device memory for the product", "", __LINE__);
            return -1;
        }
    }
    for (int d = 0; d < P->ndev; ++d) {
        spmm_dev *D = &P->devs[d];
        int c0 = 0, nd = 0;
        sblas_spmm_plan_columns(P, n, d, &c0, &nd);
        if (nd <= 0) continue;
        CU(cudaSetDevice(D->device));
        if (grow((void **)&D->d_B, &D->cap_B, (size_t)P->k * nd * sizeof(double)) ||
            grow((void **)&D->d_C, &D->cap_C, (size_t)P->m * nd * sizeof(double))) {
            sblas_set_error("%s%s (line %d)", "cudaMalloc(B / C slice) failed", "", __LINE__);
            return 1;
        }
        CU(cudaMemcpyAsync(D->d_B, B + (size_t)c0 * P->k, (size_t)P->k * nd * sizeof(double), cudaMemcpyHostToDevice, D->stream));
        if (*beta != 0.0)
            CU(cudaMemcpyAsync(D->d_C, C + (size_t)c0 * P->m, (size_t)P->m * nd * sizeof(double), cudaMemcpyHostToDevice, D->stream));
        if ((rc = sblas_spmm_plan_execute_device(P, d, nd, *alpha, D->d_B, *beta, D->d_C, 0)) != 0) return rc;
        CU(cudaMemcpyAsync(C + (size_t)c0 * P->m, D->d_C, (size_t)P->m * nd * sizeof(double), cudaMemcpyDeviceToHost, D->stream));
    }
    for (int d = 0; d < P->ndev; ++d) {
        CU(cudaSetDevice(P->devs[d].device));
        CU(cudaStreamSynchronize(P->devs[d].stream));
    }
fail:
    return rc;
}

int sblas_spmm_mgpu(int m, int n, int k, const double *alpha, int nnz, const int *rp, const int *col, const double *val,
                    const double *beta, const double *B, double *C, int ngpu)
{
    sblas_spmm_plan *P = NULL;
    int rc = sblas_spmm_plan_create(&P, m, k, nnz, rp, col, val, ngpu);
    if (rc != 0) return rc;
    rc = sblas_spmm_plan_execute(P, n, alpha, B, beta, C);
    sblas_spmm_plan_destroy(P);
    return rc == 1 ? -1 : rc;                 /* kernel / copy failure: dspmm_mgpu_baseline.cu:455-461 */
}

/* ---- the reference's names, C linkage (include/spmm_kernel.h) */
int cusparse_mgpu_csrmm(const int m, const int n, const int k, const double *alpha, const int nnz_A, int *csrRowPtr_A,
                        int *csrColIndex_A, double *csrVal_A, const double *beta, double *B_dense, double *C_dense,
                        const int ngpu)
{
    return sblas_spmm_mgpu(m, n, k, alpha, nnz_A, csrRowPtr_A, csrColIndex_A, csrVal_A, beta, B_dense, C_dense, ngpu);
}
int cusparse_mgpu_csrmm_omp(const int m, const int n, const int k, const double *alpha, const int nnz_A, int *csrRowPtr_A,
                            int *csrColIndex_A, double *csrVal_A, const double *beta, double *B_dense, double *C_dense,
                            const int ngpu)
{
    return sblas_spmm_mgpu(m, n, k, alpha, nnz_A, csrRowPtr_A, csrColIndex_A, csrVal_A, beta, B_dense, C_dense, ngpu);
}
