// oracle/ref_ingest_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// A three-line C-linkage door onto the reference's OWN Matrix-Market loader
// (sptrsv/sptrsv_v1/src/mmio_highlevel.h: mmio_info :8-136, mmio_data :139-298), compiled where the
// reference lies (-I into /root/reference; no reference source enters this repo).  Used to pin the
// oracle's load_mtx_csr and the product's sblas_mtx_read_csr, and to generate tests/golden/ingest_*.npz.
#include "mmio_highlevel.h"

extern "C" int ref_mmio_info(int *m, int *n, int *nnz, int *is_symmetric, char *filename)
{
    return mmio_info(m, n, nnz, is_symmetric, filename);
}
extern "C" int ref_mmio_data(int *csrRowPtr, int *csrColIdx, double *csrVal, char *filename)
{
    return mmio_data(csrRowPtr, csrColIdx, csrVal, filename);
}
