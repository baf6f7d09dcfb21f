/* sblas_device.h -- the boundary between the C host code and the sm_100a kernels.
 * Plain C: a POD argument block per work unit and extern "C" launchers that take
 * device pointers and a stream.  No torch types, no C++ in the signatures.
 *
 * A "segment" is one unit of work of the reference path: a v1 shard
 * (spmv/src/dspmv_mgpu_v1.cu:59-133), a v2 task (dspmv_mgpu_v2.cu:211-289) or a
 * baseline row block (dspmv_mgpu_baseline.cu:60-87): a contiguous nnz range
 * [nz0, nz1) of the arrays resident on one GPU plus the rows it touches.  The
 * launch replaces the per-GPU / per-task cusparseDcsrmv[_mp] call
 * (dspmv_mgpu_baseline.cu:163, dspmv_mgpu_v1.cu:200,206, dspmv_mgpu_v2.cu:351,357).
 */
#ifndef SBLAS_DEVICE_H
#define SBLAS_DEVICE_H
#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sblas_seg_args {
    const double *val;    /* values of this GPU's resident nnz range                      */
    const int *col;       /* column indices, same range                                   */
    const int *rowptr;    /* int32 row pointer, rebased to the GPU's range and clamped     */
    const double *x;      /* full x (replicated on every GPU)                              */
    double *y;            /* this GPU's y slice (index 0 == the GPU's first row)           */
    double *edge;         /* [2] raw partial sums of a split first / last row              */
    double *carry;        /* [ntile] tile kernel: sum of the row left open by tile j-1     */
    double *tail;         /* [ntile] tile kernel: partial of a row that leaves tile j      */
    const int *tstart;    /* [ntile+1] first row that STARTS inside tile j                 */
    const int *tmeta;     /* [8*ntile] per tile {rs, re, start of rs, flags} + 8 x u16 q_w */
    double alpha, beta;
    int row_lo, row_hi;   /* rows of the segment, inclusive, GPU-local numbering           */
    int nz0, nz1;         /* nnz range [nz0, nz1), GPU-local numbering                     */
    int skip_first;       /* row whose raw sum goes to edge[0] instead of y (or -1)        */
    int skip_last;        /* row whose raw sum goes to edge[1] instead of y (or -1)        */
    int tile0;            /* absolute index of the segment's first tile (nz0 / tile)       */
    int ntile;
    int nz_total;         /* entries resident on this GPU (bound for bulk copies)          */
    int mode;             /* diagnostics: bit0 = no warp-piece path (W), bit1 = several-rows tiles use the
                             block path (S) instead of the merge path (M); 0 in production        */
} sblas_seg_args;

/* kernel families (the `kernel` argument of the reference API maps onto these,
 * see sblas_plan.c) */
enum { SBLAS_K_VECTOR = 1, SBLAS_K_TILE = 2, SBLAS_K_TMA = 3, SBLAS_K_VECP = 4, SBLAS_K_SHORT = 5,
       SBLAS_K_ROWTILE = 6 /* ipt = R | window << 8: R rows per warp, every R consecutive rows hold <= window <= 256 entries */,
       SBLAS_K_ROWSPLIT = 7 /* ipt = G (2, 4, 8) warps per row: every row of the panel holds <= 256*G entries */ };

/* nnz per tile of the tile kernel for a given items-per-thread choice (kind TILE),
 * or of the TMA-pipelined kernel (kind TMA, ipt ignored) */
int sblas_tile_size(int ipt);
int sblas_tile_size_kind(int kind, int ipt);

/* int64 harness row pointer slice -> int32 GPU-local row pointer:
 * out[i] = clamp(rp64[i] - first_idx, 0, total_nnz), i in [0,count).  Reproduces
 * the local row pointer of dspmv_mgpu_v1.cu:125-133 / dspmv_mgpu_baseline.cu:82-85. */
cudaError_t sblas_launch_rebase_rowptr(const long long *rp64, long long first_idx, int total_nnz,
                                       long long count, int *out, cudaStream_t s);

/* tstart[j] for j in [0,ntile]: first row whose first entry lies at or after the
 * start of tile j (binary search per tile; replaces CSR5's tile pointer
 * generation, spmv/include/detail/cuda/format_cuda.h:21-42). */
cudaError_t sblas_launch_tile_rows(const sblas_seg_args *a, int tile, int *tstart_out, cudaStream_t s);
/* tmeta[8*j..] = {rs, re, clamp(rowptr[rs]) or tile end, flags (1: last row leaves the tile,
 * 2: an empty row starts here, 4: no warp chunk holds more than 7 row starts)} followed by 8 x uint16 "rows starting before chunk w":
 * everything a tile needs in two 16-byte loads, computed once per plan. */
cudaError_t sblas_launch_tile_meta(const sblas_seg_args *a, int tile, int *tmeta_out, cudaStream_t s);

/* Row-length statistics per block of `rb` rows of a segment (rows row_lo .. row_lo+nrows-1):
 * out_max[b] = longest row, out_ptr[b] = clamp(rowptr[first row of block], nz0, nz1).  The plan
 * bins consecutive blocks into panels and picks a kernel per panel (adaptive row binning). */
cudaError_t sblas_launch_row_block_stats(const int *rowptr, int row_lo, int nrows, int rb, int nz0, int nz1,
                                         int *out_max, int *out_ptr, cudaStream_t s);

/* out_min_max[0] = min(out[0], min col), out_min_max[1] = max(out[1], max col) over col[0..count):
 * the window of x a shard can read (initialise to {INT_MAX, -1}). */
cudaError_t sblas_launch_col_range(const int *col, long long count, int *out_min_max, cudaStream_t s);

/* y[row_lo..row_hi] = alpha*A_seg*x + beta*y for one segment. kind: SBLAS_K_*;
 * ipt: items per thread of the tile kernel (4, 8 or 16); lanes: lanes per row of
 * the vector kernel (0 = choose from the mean row length). */
cudaError_t sblas_launch_spmv_segment(const sblas_seg_args *a, int kind, int ipt, int lanes, cudaStream_t s);

/* finish rows split between segments: y[mrow[i]] = alpha * sum_k *msrc[k] + beta*y[..],
 * k in [mbeg[i], mbeg[i+1]) in list order (ascending segment order).  Sources may be
 * peer-GPU addresses (NVLink P2P).  Replaces the host merges of
 * dspmv_mgpu_v1.cu:235-248 and dspmv_mgpu_v2.cu:385-441. */
cudaError_t sblas_launch_edge_merge(const int *mrow, const int *mbeg, const double *const *msrc, int nmerge,
                                    double *y, double alpha, double beta, cudaStream_t s);

/* Fused split-row exchange for one-process-per-GPU plans over peer-mapped (symmetric)
 * memory, no host involvement and no NCCL call per product:
 *   publish: advance the product counter *epoch_ctr (device memory, so that a product is the same
 *            launch sequence every time: CUDA-graph replay); copy this rank's partials (local_edge,
 *            nlocal = edge slots) into its own table half `epoch & 1` and the outgoing ones straight into
 *            the OWNER ranks' tables over NVLink (P2P stores), fence, then raise the owners' arrive flags
 *            (epoch counters); waits for the owners' ack of epoch-2 first (tables are double buffered)
 *   merge:   spin until every contributing rank's arrive flag reached *epoch_ctr, finish the
 *            split rows from the local table in ascending segment order, ack the contributors.
 * Buffer layout of every rank (8-byte words): table[2][table_words], arrive[world], ack[world]. */
cudaError_t sblas_launch_edge_publish(const double *local_edge, int nlocal, const int *out_slot, const int *out_owner,
                                      const long long *out_off, int nout, const int *owners, int nowners,
                                      void *const *peer_bases, long long table_words, int world, int my_rank,
                                      unsigned long long *epoch_ctr, cudaStream_t s);
cudaError_t sblas_launch_edge_merge_wait(const int *mrow, const int *mbeg, const long long *msrc_off, int nmerge,
                                         double *y, double alpha, double beta, const int *contrib, int ncontrib,
                                         void *const *peer_bases, long long table_words, int world, int my_rank,
                                         const unsigned long long *epoch_ctr, cudaStream_t s);

/* x <- y across the ranks of a one-process-per-GPU job over peer-mapped memory: every rank
 * stores the `count` rows of y it owns (y_src) into every rank's x at [dst_off, dst_off + count) -- an all-gather
 * with P2P stores over NVLink --, fenced by two rounds of epoch flags (ready: all ranks have finished reading
 * x; written: all slices have landed).  peer_x[r] / peer_flags[r] (2*world 8-byte words, zeroed) are rank r's
 * buffers as mapped in this process; chain_ctr is a device counter owned by the plan. */
cudaError_t sblas_launch_chain_gather(const double *y_src, long long count, long long dst_off, void *const *peer_x,
                                      void *const *peer_flags, int world, int my_rank, unsigned long long *chain_ctr,
                                      cudaStream_t s);

/* load the kernels of the flag protocols (publish / merge-wait / chain) now: their first launch must not happen
 * while one of them is already spinning on the same device (lazy module loading may synchronise the device) */
cudaError_t sblas_preload_exchange_kernels(void);

/* device fill helpers used by the plan */
cudaError_t sblas_launch_fill_f64(double *p, long long n, double v, cudaStream_t s);

#ifdef __cplusplus
}
#endif
#endif
