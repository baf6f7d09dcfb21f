/*
 * sblas_sptrans.h -- C-ABI of the multi-GPU sparse transposition CSR -> CSC.
 * Replaces kernal_sptrans of the reference (sptrans/sptrans_v1/src/sptrans_kernal.h:80-530: rows split in
 * equal blocks over the GPUs :131-148, cusparseCsr2cscEx2 per block :228-262, composition of the blocks'
 * column pointers and entries :12-78) with hand-written sm_100a kernels (s-blas_b200/csrc/sblas_sptrans.cu).
 * The output equals, entry for entry, the host transposition the reference checks itself against
 * (sptrans/sptrans_v1/src/tranpose.h:3-40): inside a column the entries keep their CSR order.
 * A is m x n with int32 row pointer / indices (base 0), double values; all pointers are HOST pointers.
 * Returns 0 on success, 1 on a set-up / copy / kernel failure (the reference's convention, :165-217).
 */
#ifndef SBLAS_SPTRANS_H
#define SBLAS_SPTRANS_H
#ifdef __cplusplus
extern "C" {
#endif

int sblas_sptrans_mgpu(int m, int n, int nnz, int ngpu, const int *csrRowPtr, const int *csrColIdx,
                       const double *csrVal, int *cscRowIdx, int *cscColPtr, double *cscVal);

/* the reference's own entry point, argument for argument (the three *_ref arrays are the host result it
 * prints a comparison against; here they are compared when not NULL and a mismatch returns 2) */
int kernal_sptrans(const int m, const int n, const int nnz, int ngpu, const int *csrRowPtr, const int *csrColIdx,
                   const double *csrVal, int *cscRowIdx, int *cscColPtr, double *cscVal, const int *cscRowIdx_ref,
                   const int *cscColPtr_ref, const double *cscVal_ref);

/* last call: milliseconds spent on the devices between the end of the uploads and the end of the composition
 * (the reference prints this phase as "cuSparse trans", sptrans_kernal.h:219-275) */
double sblas_sptrans_last_device_ms(void);

#ifdef __cplusplus
}
#endif
#endif
