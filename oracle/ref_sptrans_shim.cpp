// oracle/ref_sptrans_shim.cpp -- TEST INFRASTRUCTURE.  A C-linkage door onto the reference's OWN host
// transposition, sptrans/sptrans_v1/src/tranpose.h:3-40 (matrix_transposition) with the exclusive_scan of
// sptrans/sptrans_v1/src/utils.h:312-329, both compiled where they lie (oracle/Makefile, target refsptrans ->
// oracle/_ref/libref_sptrans.so).  It is the result the reference's multi-GPU kernal_sptrans is checked against
// (sptrans/sptrans_v1/src/main.cu:152-160,254-258), so it pins the oracle and the library bit for bit.
#include "utils.h"
#include "tranpose.h"

extern "C" void ref_matrix_transposition(int m, int n, int nnz, const int *csrRowPtr, const int *csrColIdx,
                                         const double *csrVal, int *cscRowIdx, int *cscColPtr, double *cscVal)
{
    matrix_transposition(m, n, nnz, csrRowPtr, csrColIdx, csrVal, cscRowIdx, cscColPtr, cscVal);
}
