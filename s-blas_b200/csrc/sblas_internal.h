/* sblas_internal.h -- plan data structures (host C). Not part of the C-ABI. */
#ifndef SBLAS_INTERNAL_H
#define SBLAS_INTERNAL_H
#include <cuda_runtime_api.h>
#include "sblas_device.h"
#include "sblas_spmv.h"

/* one kernel launch: a panel of a segment (consecutive rows binned by their longest row) */
typedef struct sblas_unit {
    sblas_seg_args args;   /* the panel's rows and nnz range; edge rows only at the segment's ends */
    int kind, ipt;         /* SBLAS_K_* chosen for the panel                     */
    long long tile_off;    /* offset of its tile metadata in the GPU's arrays    */
} sblas_unit;

/* one live segment held by this process (a v1 shard, a v2 task, a baseline block) */
typedef struct sblas_seg {
    int gidx;              /* index in the global partition (plan->parts)       */
    int dev;               /* index of the GPU inside the plan                  */
    int lidx;              /* index among that GPU's segments                   */
    int stream;            /* which of the GPU's q streams runs it              */
    sblas_seg_args args;   /* the whole segment (rows, nnz range, edge rows)    */
    int unit_begin, unit_end;   /* plan->units[unit_begin, unit_end): its panels in row order */
} sblas_seg;

/* one GPU of the plan: a contiguous resident nnz range and the rows it touches */
typedef struct sblas_dev {
    int device;                       /* CUDA ordinal */
    int seg_begin, seg_end;           /* plan->segs[seg_begin, seg_end) live here (-1: none) */
    long long first_idx, last_idx;    /* global nnz range, inclusive */
    int first_row, last_row, rows, nnz;
    double *d_val; int *d_col; int own_matrix;
    int *d_rowptr; double *d_x; double *d_y;
    long long *stage64;               /* plan build: the int64 row pointer slice on its way to d_rowptr (aliases d_y) */
    double *d_edge; int edge_bound;
    double *d_carry, *d_tail; int *d_tstart, *d_tmeta;
    /* merge lists of the split rows this GPU owns */
    int nmerge, nmsrc;
    int *d_mrow, *d_mbeg; const double **d_msrc;
    /* the arrays above are carved out of three allocations (cudaMalloc/cudaFree with peer access enabled map into
     * every GPU's address space: their count, not their size, is what a one-shot call pays for) */
    char *slab_main;                  /* val, col (when owned), rowptr, x, y, edge table, column-range scratch */
    size_t slab_main_bytes;
    char *slab_tiles;                 /* tmeta, tstart, carry, tail */
    char *slab_merge;                 /* mrow, mbeg, msrc */
    int *d_mm;                        /* plan build: {min, max} column of the shard */
    int *h_mrow, *h_mbeg; const double **h_msrc; long long *h_msrc_off;
    cudaStream_t *streams; int nstreams;
    cudaEvent_t *ev_seg, ev_in, ev_done;
    cudaStream_t copy_stream;         /* second stream: y slices move while other panels compute */
    cudaEvent_t *ev_unit; int nev_unit; /* per panel: its y slice is on the GPU / its kernel is done */
    cudaEvent_t ev_y, ev_chain;       /* y complete on this GPU / this GPU has pulled every y slice (chain) */
    cudaEvent_t ev_merge, ev_xpull;   /* this GPU's merge has read its peers' edge tables / has pulled its peers' x slices */
    int merge_recorded, xpull_recorded;
    int kind, ipt;
    long long xs_lo, xs_hi;           /* slice of x this GPU uploads itself */
    int col_lo, col_hi;               /* smallest / largest column of the resident shard: the x it reads */
} sblas_dev;

struct sblas_spmv_plan {
    int version, m, n, kernel, q, world, rank, rank_mode, ndev, p2p, dry;
    int x_policy;                     /* an L2 access-policy window over x is set on the streams (SBLAS_X_PERSIST=1) */
    int pooled;                       /* main allocations come from / go back to the per-GPU pool (one-shot calls) */
    long long nnz, nb;
    /* global partition, identical on every rank */
    int nparts; sblas_part *parts;
    int *g_owner, *g_local, *g_lo, *g_hi, *g_sf, *g_sl;
    int max_local;
    int nseg; sblas_seg *segs;
    int nunits, cap_units; sblas_unit *units;
    sblas_dev *devs;
    const double *gather_base;
    int cap_pieces; int *piece_lo, *piece_hi, *piece_unit;    /* pieces of the pipelined host execute */
    /* fused exchange over peer-mapped memory */
    int peer_bound, nout, nowners, ncontrib;
    unsigned long long *d_epoch;     /* product counter of the fused exchange, advanced by the publish kernel */
    /* x <- y across ranks over peer-mapped memory (sblas_spmv_plan_bind_peer_x) */
    int peer_x_bound; void **d_peer_x, **d_peer_xflags; unsigned long long *d_chain_ctr; double *x_alloc;
    /* CUDA-graph replay of one product (sblas_spmv_plan_step) */
    void *graph_exec; int graph_valid, graph_warm, capturing; double graph_alpha, graph_beta;
    long long table_words;
    double *my_base;
    void **d_peer_bases; int *d_out_slot, *d_out_owner, *d_owners, *d_contrib; long long *d_out_off, *d_msrc_off;
};

void sblas_set_error(const char *fmt, const char *a, const char *b, int line);

/* One retained main allocation per GPU for the one-shot entry points (sblas_api.c): with peer access enabled a
 * cudaMalloc + cudaFree pair of a GB-sized shard costs tens of ms, far more than the product, and the reference's
 * harness calls the one-shot entry points again and again.  SBLAS_POOL=0 turns it off;
 * sblas_spmv_cache_clear() gives the memory back. */
#define SBLAS_CREATE_POOLED 0x100
int sblas_spmv_plan_create_flags(sblas_spmv_plan **plan, int version, int m, int n, long long nnz,
                                 const double *csrVal, const long long *csrRowPtr, const int *csrColIndex,
                                 int ngpu, int kernel, long long nb, int q, int flags);
void sblas_pool_release(void);

/* host arithmetic of the in-process x upload (sblas_plan.c), exposed for the CPU tests */
void sblas_x_slice(long long n, int li, int live, long long *lo, long long *hi);
void sblas_x_pull_range(long long slice_lo, long long slice_hi, long long win_lo, long long win_hi,
                        long long *lo, long long *hi);
#endif
