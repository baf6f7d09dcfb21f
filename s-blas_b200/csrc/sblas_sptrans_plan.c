/* sblas_sptrans_plan.c -- host side (C) of the multi-GPU CSR -> CSC transposition.
 * Follows the structure of the reference's kernal_sptrans (sptrans/sptrans_v1/src/sptrans_kernal.h:80-530):
 *   rows split in equal blocks, start_row = floor(d*m/ngpu), end_row = floor((d+1)*m/ngpu) - 1 (:131-133),
 *   block-local row pointers (:143-147), one conversion per GPU (:228-262), composition (:12-78).
 * Here the conversion is the radix-sort path of sblas_sptrans.cu, the blocks' column pointers meet on GPU 0
 * (peer copies over NVLink), GPU 0 computes the global column pointer and every block's per-column base, and
 * every GPU then writes its (row, value) pairs STRAIGHT into the final arrays on GPU 0 with P2P stores --
 * no staging copy of the blocks.  No cuSPARSE, no CPU arithmetic on the data path.
 */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sblas_internal.h"
#include "sblas_sptrans.h"

long long sblas_csr2csc_scratch_ints(long long nnz);
cudaError_t sblas_launch_csr2csc(int m, int n, long long nnz, const int *d_rowptr, const int *d_col, const double *d_val,
                                 int row_base, int *d_colptr, int **sorted_keys, int **sorted_pos, int **rowidx,
                                 int *scratch, cudaStream_t s);
cudaError_t sblas_launch_csc_gather(const int *keys, const int *pos, const int *rowidx, const double *val, long long nnz,
                                    const int *base, int *out_row, double *out_val, cudaStream_t s);
cudaError_t sblas_launch_csc_compose(const int *ptrs, int ndev, int n, int *gcolptr, int *bases, cudaStream_t s);

#define CU(call)                                                                      \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            sblas_set_error("%s failed: %s (sblas_sptrans_plan.c:%d)", #call, cudaGetErrorString(e_), __LINE__); \
            rc = 1;                                                                   \
            goto fail;                                                                \
        }                                                                             \
    } while (0)

typedef struct tr_dev {
    int start_row, rows, nnz;
    int *h_rowptr;
    int *d_rowptr, *d_col, *d_colptr, *d_scratch, *d_base;
    double *d_val;
    int *keys, *pos, *rowidx;          /* inside d_scratch */
    cudaStream_t stream;
    cudaEvent_t ev;
} tr_dev;

static double g_last_ms;
double sblas_sptrans_last_device_ms(void) { return g_last_ms; }

int sblas_sptrans_mgpu(int m, int n, int nnz, int ngpu, const int *rp, const int *col, const double *val, int *cscRow,
                       int *cscColPtr, double *cscVal)
{
    int rc = 0, count = 0;
    tr_dev *T = NULL;
    int *d_ptrs = NULL, *d_gcolptr = NULL, *d_bases = NULL, *d_out_row = NULL;
    double *d_out_val = NULL;
    cudaEvent_t e0 = NULL, e1 = NULL;
    if (m <= 0 || n <= 0 || nnz < 0 || ngpu <= 0 || !rp || !cscColPtr) {
        sblas_set_error("%s%s (line %d)", "invalid argument", "", __LINE__);
        return 1;
    }
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < ngpu) {
        cudaGetLastError();
        sblas_set_error("%s%s (line %d)", "not enough CUDA devices (no CPU fallback)", "", __LINE__);
        return 1;
    }
    T = (tr_dev *)calloc((size_t)ngpu, sizeof(tr_dev));
    if (!T) return 1;
    if (ngpu > 1) {
        for (int a = 0; a < ngpu; ++a) {
            CU(cudaSetDevice(a));
            for (int b = 0; b < ngpu; ++b) {
                if (a == b) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, a, b);
                if (!can) {
                    sblas_set_error("%s%s (line %d)", "multi-GPU transposition needs peer access between the GPUs (NVLink/NVSwitch)", "", __LINE__);
                    rc = 1; goto fail;
                }
                cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
                if (e != cudaSuccess) cudaGetLastError();           /* already enabled */
            }
        }
    }
    /* ---- blocks, uploads, conversions: every GPU on its own stream */
    for (int d = 0; d < ngpu; ++d) {
        tr_dev *D = &T[d];
        D->start_row = (int)((long long)d * m / ngpu);                           /* sptrans_kernal.h:131 */
        const int end_row = (int)((long long)(d + 1) * m / ngpu) - 1;            /* :132 */
        D->rows = end_row - D->start_row + 1;
        D->nnz = rp[end_row + 1] - rp[D->start_row];
        D->h_rowptr = (int *)malloc(((size_t)D->rows + 1) * sizeof(int));
        if (!D->h_rowptr) { rc = 1; goto fail; }
        for (int i = 0; i <= D->rows; ++i) D->h_rowptr[i] = rp[D->start_row + i] - rp[D->start_row];   /* :143-147 */
        CU(cudaSetDevice(d));
        CU(cudaStreamCreateWithFlags(&D->stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&D->ev, cudaEventDisableTiming));
        CU(cudaMalloc((void **)&D->d_rowptr, ((size_t)D->rows + 1) * sizeof(int)));
        CU(cudaMalloc((void **)&D->d_col, ((size_t)D->nnz + 8) * sizeof(int)));
        CU(cudaMalloc((void **)&D->d_val, ((size_t)D->nnz + 8) * sizeof(double)));
        CU(cudaMalloc((void **)&D->d_colptr, ((size_t)n + 1) * sizeof(int)));
        CU(cudaMalloc((void **)&D->d_scratch, (size_t)sblas_csr2csc_scratch_ints(D->nnz) * sizeof(int)));
        if (ngpu > 1) CU(cudaMalloc((void **)&D->d_base, (size_t)n * sizeof(int)));
        CU(cudaMemcpyAsync(D->d_rowptr, D->h_rowptr, ((size_t)D->rows + 1) * sizeof(int), cudaMemcpyHostToDevice, D->stream));
        if (D->nnz > 0) {
            CU(cudaMemcpyAsync(D->d_col, col + rp[D->start_row], (size_t)D->nnz * sizeof(int), cudaMemcpyHostToDevice, D->stream));
            CU(cudaMemcpyAsync(D->d_val, val + rp[D->start_row], (size_t)D->nnz * sizeof(double), cudaMemcpyHostToDevice, D->stream));
        }
    }
    CU(cudaSetDevice(0));
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaMalloc((void **)&d_out_row, ((size_t)nnz + 8) * sizeof(int)));
    CU(cudaMalloc((void **)&d_out_val, ((size_t)nnz + 8) * sizeof(double)));
    if (ngpu > 1) {
        CU(cudaMalloc((void **)&d_ptrs, (size_t)ngpu * ((size_t)n + 1) * sizeof(int)));
        CU(cudaMalloc((void **)&d_gcolptr, ((size_t)n + 1) * sizeof(int)));
        CU(cudaMalloc((void **)&d_bases, (size_t)ngpu * (size_t)n * sizeof(int)));
    }
    for (int d = 0; d < ngpu; ++d) { CU(cudaSetDevice(d)); CU(cudaStreamSynchronize(T[d].stream)); }   /* uploads done */
    CU(cudaSetDevice(0));
    CU(cudaEventRecord(e0, T[0].stream));
    for (int d = 0; d < ngpu; ++d) {
        tr_dev *D = &T[d];
        CU(cudaSetDevice(d));
        if (d > 0) CU(cudaStreamWaitEvent(D->stream, e0, 0));
        CU(sblas_launch_csr2csc(D->rows, n, D->nnz, D->d_rowptr, D->d_col, D->d_val, D->start_row, D->d_colptr, &D->keys,
                                &D->pos, &D->rowidx, D->d_scratch, D->stream));
        if (ngpu > 1) {
            CU(cudaMemcpyPeerAsync(d_ptrs + (size_t)d * ((size_t)n + 1), 0, D->d_colptr, d, ((size_t)n + 1) * sizeof(int), D->stream));
            CU(cudaEventRecord(D->ev, D->stream));
        }
    }
    if (ngpu == 1) {
        CU(sblas_launch_csc_gather(T[0].keys, T[0].pos, T[0].rowidx, T[0].d_val, T[0].nnz, NULL, d_out_row, d_out_val, T[0].stream));
        CU(cudaEventRecord(e1, T[0].stream));
        CU(cudaMemcpyAsync(cscColPtr, T[0].d_colptr, ((size_t)n + 1) * sizeof(int), cudaMemcpyDeviceToHost, T[0].stream));
    } else {
        CU(cudaSetDevice(0));
        for (int d = 1; d < ngpu; ++d) CU(cudaStreamWaitEvent(T[0].stream, T[d].ev, 0));
        CU(sblas_launch_csc_compose(d_ptrs, ngpu, n, d_gcolptr, d_bases, T[0].stream));
        CU(cudaEventRecord(T[0].ev, T[0].stream));
        for (int d = 0; d < ngpu; ++d) {
            tr_dev *D = &T[d];
            CU(cudaSetDevice(d));
            if (d > 0) CU(cudaStreamWaitEvent(D->stream, T[0].ev, 0));
            CU(cudaMemcpyPeerAsync(D->d_base, d, d_bases + (size_t)d * (size_t)n, 0, (size_t)n * sizeof(int), D->stream));
            /* every block writes its entries straight into the final arrays on GPU 0 (P2P stores over NVLink) */
            CU(sblas_launch_csc_gather(D->keys, D->pos, D->rowidx, D->d_val, D->nnz, D->d_base, d_out_row, d_out_val, D->stream));
            if (d > 0) CU(cudaEventRecord(D->ev, D->stream));
        }
        CU(cudaSetDevice(0));
        for (int d = 1; d < ngpu; ++d) CU(cudaStreamWaitEvent(T[0].stream, T[d].ev, 0));
        CU(cudaEventRecord(e1, T[0].stream));
        CU(cudaMemcpyAsync(cscColPtr, d_gcolptr, ((size_t)n + 1) * sizeof(int), cudaMemcpyDeviceToHost, T[0].stream));
    }
    CU(cudaSetDevice(0));
    if (nnz > 0) {
        CU(cudaMemcpyAsync(cscRow, d_out_row, (size_t)nnz * sizeof(int), cudaMemcpyDeviceToHost, T[0].stream));
        CU(cudaMemcpyAsync(cscVal, d_out_val, (size_t)nnz * sizeof(double), cudaMemcpyDeviceToHost, T[0].stream));
    }
    CU(cudaStreamSynchronize(T[0].stream));
    {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) g_last_ms = ms;
    }
fail:
    if (T) {
        for (int d = 0; d < ngpu; ++d) {
            tr_dev *D = &T[d];
            if (!D->stream && !D->h_rowptr) continue;
            cudaSetDevice(d);
            if (D->stream) cudaStreamSynchronize(D->stream);
            cudaFree(D->d_rowptr); cudaFree(D->d_col); cudaFree(D->d_val); cudaFree(D->d_colptr); cudaFree(D->d_scratch);
            cudaFree(D->d_base);
            if (D->stream) cudaStreamDestroy(D->stream);
            if (D->ev) cudaEventDestroy(D->ev);
            free(D->h_rowptr);
        }
        cudaSetDevice(0);
        cudaFree(d_ptrs); cudaFree(d_gcolptr); cudaFree(d_bases); cudaFree(d_out_row); cudaFree(d_out_val);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        free(T);
    }
    return rc;
}

int kernal_sptrans(const int m, const int n, const int nnz, int ngpu, const int *csrRowPtr, const int *csrColIdx,
                   const double *csrVal, int *cscRowIdx, int *cscColPtr, double *cscVal, const int *cscRowIdx_ref,
                   const int *cscColPtr_ref, const double *cscVal_ref)
{
    int rc = sblas_sptrans_mgpu(m, n, nnz, ngpu, csrRowPtr, csrColIdx, csrVal, cscRowIdx, cscColPtr, cscVal);
    if (rc != 0) return rc;
    if (cscColPtr_ref && memcmp(cscColPtr, cscColPtr_ref, ((size_t)n + 1) * sizeof(int)) != 0) return 2;
    if (cscRowIdx_ref && memcmp(cscRowIdx, cscRowIdx_ref, (size_t)nnz * sizeof(int)) != 0) return 2;
    if (cscVal_ref && memcmp(cscVal, cscVal_ref, (size_t)nnz * sizeof(double)) != 0) return 2;
    return 0;
}
