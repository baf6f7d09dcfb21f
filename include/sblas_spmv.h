/*
 * sblas_spmv.h -- C-ABI of the B200-native multi-GPU CSR SpMV library
 * (libsblas_spmv.so).  y = alpha*A*x + beta*y, double precision, CSR with the
 * harness's 64-bit row pointer.
 *
 * This is the drop-in boundary for ONE path of pnnl/s-blas: the three host entry
 * points of spmv/include/spmv_kernel.h:11-36 and what they call.  Every symbol is
 * extern "C", takes plain pointers and sizes, and cites the reference interface it
 * replaces.  The same three functions are additionally exported under the
 * reference's own names (C linkage for C callers, and the reference's C++-mangled
 * names for an unmodified spmv/test/dspmv_test.cu) by include/spmv_kernel.h.
 *
 * Return convention (spmv/src/dspmv_mgpu_baseline.cu:77-79,99-153,173-175;
 * dspmv_mgpu_v1.cu:114-116,141-157,226-228; dspmv_mgpu_v2.cu:44-46):
 *    0  success
 *   -1  shard larger than 0.8 x the smallest free device memory, kernel failure,
 *       or (v2) nb <= 0 || ngpu == 0 || q == 0
 *    1  set-up failure (device, stream, allocation, copy)
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef SBLAS_SPMV_H
#define SBLAS_SPMV_H

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ one-shot API
 * Host pointers in, host y out; replaces, argument for argument,
 *   spMV_mgpu_baseline  spmv/include/spmv_kernel.h:11-15  (spmv/src/dspmv_mgpu_baseline.cu:14)
 *   spMV_mgpu_v1        spmv/include/spmv_kernel.h:16-21  (spmv/src/dspmv_mgpu_v1.cu:16)
 *   spMV_mgpu_v2        spmv/include/spmv_kernel.h:23-30  (spmv/src/dspmv_mgpu_v2.cu:33)
 * kernel: 1 = adaptive (binned per tile), 2 = nnz-balanced tile kernel everywhere,
 *         3 = CSR5-style small tiles (the reference's kernel 3 is a no-op).
 */
int sblas_spmv_mgpu_baseline(int m, int n, long long nnz, double *alpha, double *csrVal,
                             long long *csrRowPtr, int *csrColIndex, double *x, double *beta,
                             double *y, int ngpu);
int sblas_spmv_mgpu_v1(int m, int n, long long nnz, double *alpha, double *csrVal,
                       long long *csrRowPtr, int *csrColIndex, double *x, double *beta,
                       double *y, int ngpu, int kernel);
int sblas_spmv_mgpu_v2(int m, int n, long long nnz, double *alpha, double *csrVal,
                       long long *csrRowPtr, int *csrColIndex, double *x, double *beta,
                       double *y, int ngpu, int kernel, long long nb, int copy_of_workspace);

/* Optional plan cache of the one-shot entry points (environment SBLAS_PLAN_CACHE=1, off by
 * default = reference semantics): repeated calls with the same host arrays reuse the resident
 * shards and only move x and y.  Drop every cached plan (needed after editing csrVal in place). */
void sblas_spmv_cache_clear(void);

/* helpers of spmv/include/spmv_kernel.h:32-36 (spmv/src/spmv_helper.cu:16-39,41-48,51-76) */
int sblas_get_row_from_index(int n, const long long *a, long long idx);
double sblas_get_time(void);
double sblas_get_gpu_availble_mem(int ngpu);

/* ------------------------------------------------------------------ partitioners
 * Host-side, pure integer work; bit-exact restatements of the reference formulas.
 * One record per GPU (baseline, v1) or per task (v2); field names follow
 * struct spmv_task (spmv/include/spmv_task.h:4-36). */
typedef struct sblas_part {
    long long start_idx, end_idx;   /* inclusive nnz range                       */
    int start_row, end_row;         /* rows containing start_idx / end_idx       */
    int start_flag, end_flag;       /* 1 = the row is shared with a neighbour    */
    int dev_m, dev_nnz;
} sblas_part;

/* spmv/src/dspmv_mgpu_baseline.cu:60-87 (flags are always 0) */
int sblas_partition_baseline(int m, const long long *csrRowPtr, int ngpu, sblas_part *out);
/* spmv/src/dspmv_mgpu_v1.cu:59-100,119 */
int sblas_partition_v1(int m, long long nnz, const long long *csrRowPtr, int ngpu, sblas_part *out);
/* NOT in the reference, opt-in (version SBLAS_V1_BYTES): v1's contiguous nnz ranges with split rows, cut at equal
 * shares of the streamed BYTES (12 per entry + row_bytes per row) instead of equal entry counts */
int sblas_partition_bytes(int m, long long nnz, const long long *csrRowPtr, int ngpu, int row_bytes, sblas_part *out);
/* spmv/src/dspmv_mgpu_v2.cu:218 (task count) and :211-275 (generate_tasks) */
int sblas_v2_num_tasks(long long nnz, long long nb);
int sblas_generate_tasks_v2(int m, long long nnz, const long long *csrRowPtr, long long nb, sblas_part *out);
/* task -> GPU map: GPU d owns tasks [T*d/ngpu, T*(d+1)/ngpu): the quota of
 * dspmv_mgpu_v2.cu:125-126 with a fixed (deterministic) assignment */
int sblas_v2_task_owner(int T, int ngpu, int task);
/* local int32 row pointer of one shard/task: dspmv_mgpu_v1.cu:125-133 and
 * dspmv_mgpu_v2.cu:279-289 (baseline != 0: dspmv_mgpu_baseline.cu:82-85).
 * out has dev_m + 1 entries. */
void sblas_local_rowptr(const long long *csrRowPtr, const sblas_part *p, int baseline, int *out);

/* ------------------------------------------------------------------ plan API
 * The reference re-uploads the matrix on every call.  A plan keeps
 * the partition, the device-resident shards, the int32 row pointers and the tile
 * metadata, so that repeated products touch only HBM (and NVLink for the
 * boundary rows).  The one-shot functions above are create + execute + destroy.
 */
typedef struct sblas_spmv_plan sblas_spmv_plan;

enum { SBLAS_BASELINE = 0, SBLAS_V1 = 1, SBLAS_V2 = 2, SBLAS_V1_BYTES = 3 /* byte-balanced v1, opt-in, not in the reference */ };
#define SBLAS_ROW_BYTES 28      /* per-row weight of SBLAS_V1_BYTES: row pointer 4 + y read and write 16 + x 8 */

/* In-process multi-GPU plan on devices 0..ngpu-1 from HOST arrays (pinned or
 * pageable).  version: SBLAS_*.  nb / q are only used by SBLAS_V2. */
int sblas_spmv_plan_create(sblas_spmv_plan **plan, int version, int m, int n, long long nnz,
                           const double *csrVal, const long long *csrRowPtr, const int *csrColIndex,
                           int ngpu, int kernel, long long nb, int q);

/* One-process-per-GPU plan: this process holds only shard `rank` of `world`
 * (same partition formulas) on CUDA device `device`.  The CSR arrays are HOST
 * pointers to the WHOLE matrix unless SBLAS_SRC_DEVICE_SHARD is set in `flags`,
 * in which case csrVal/csrColIndex are DEVICE pointers to exactly this rank's
 * nnz range [start_idx, end_idx] (adopted, not copied) and csrRowPtr is still
 * the whole host row pointer.  Adopted arrays must start on a 32-byte boundary and
 * be readable for 16 bytes past their last entry (the kernels fetch them with
 * 16-byte-granular bulk copies; any cudaMalloc / framework allocation qualifies). */
enum { SBLAS_SRC_HOST = 0, SBLAS_SRC_DEVICE_SHARD = 1, SBLAS_LAYOUT_ONLY = 2 };
int sblas_spmv_plan_create_rank(sblas_spmv_plan **plan, int version, int m, int n, long long nnz,
                                const double *csrVal, const long long *csrRowPtr, const int *csrColIndex,
                                int world, int rank, int device, int kernel, long long nb, int q, int flags);

/* y = alpha*A*x + beta*y with HOST x (length n) and HOST y (length m, in/out):
 * uploads x (and y when beta != 0), runs every segment, downloads y and merges
 * the split boundary rows in ascending segment order. In a rank plan only this
 * rank's rows of y are written; rows shared with other ranks are finished by the fused
 * exchange when peer tables are bound (sblas_spmv_plan_bind_peer_tables: then every rank
 * must make this call once per product), else left to sblas_spmv_plan_edges / the caller's
 * exchange. */
int sblas_spmv_plan_execute(sblas_spmv_plan *plan, const double *alpha, const double *x,
                            const double *beta, double *y);

/* The two halves of sblas_spmv_plan_execute, for callers that put their own exchange
 * between them (rank plans): upload enqueues the H2D copies of x (n doubles) and, when y is
 * not NULL, of this plan's rows of y; download copies the rows this plan owns back to the
 * host y and waits for the plan's GPUs. */
int sblas_spmv_plan_upload(sblas_spmv_plan *plan, const double *x, const double *y);
int sblas_spmv_plan_download(sblas_spmv_plan *plan, double *y);
/* Iterative use (y -> x chaining): x <- y on every GPU of the plan, on the
 * devices only (NVLink all-gather of the row slices each GPU owns, event-ordered, no host wait).
 * Square matrices; in-process plans (or a single-rank plan).  Follow with execute_device. */
int sblas_spmv_plan_chain(sblas_spmv_plan *plan);

/* One-process-per-GPU form of the chain: bind peer-mapped x buffers (n doubles each; peer_x[r] / peer_flags[r] =
 * rank r's x and flag buffer -- 2*world zeroed 8-byte words -- as mapped in this process, e.g. from
 * torch.distributed._symmetric_memory).  The plan then computes on peer_x[rank], and sblas_spmv_plan_chain
 * all-gathers the y rows every rank owns into EVERY rank's x with P2P stores over NVLink, fenced by epoch flags
 * (no NCCL call, no host synchronisation); every rank must call it once per product. */
int sblas_spmv_plan_bind_peer_x(sblas_spmv_plan *plan, void *const *peer_x, void *const *peer_flags);

/* Device-resident execute: x and y already sit in the plan's device buffers
 * (see sblas_spmv_plan_x / _y); nothing crosses PCIe.  Enqueues on the plan's
 * streams and returns without synchronising unless sync != 0. */
int sblas_spmv_plan_execute_device(sblas_spmv_plan *plan, double alpha, double beta, int sync);

/* One product with x and y resident, the iterative-use entry point: the segments'
 * kernels plus, for a rank plan with bound peer tables, the fused split-row exchange -- captured into a
 * CUDA graph on single-GPU plans and replayed (one cudaGraphLaunch per product; re-captured when alpha or
 * beta change; SBLAS_GRAPH=0 disables).  Asynchronous on the plan's stream. */
int sblas_spmv_plan_step(sblas_spmv_plan *plan, double alpha, double beta);

/* accessors (dev = index of the GPU inside the plan, 0 for a rank plan) */
int sblas_spmv_plan_num_devices(const sblas_spmv_plan *plan);
int sblas_spmv_plan_num_segments(const sblas_spmv_plan *plan);
int sblas_spmv_plan_segment(const sblas_spmv_plan *plan, int seg, sblas_part *out, int *device);
double *sblas_spmv_plan_x(sblas_spmv_plan *plan, int dev);
/* the columns the GPU's resident shard references, [first_col, last_col]: a plan with one GPU in
 * this process uploads only that window of x (the reference uploads all of x to every GPU,
 * dspmv_mgpu_v1.cu:183) */
int sblas_spmv_plan_x_window(const sblas_spmv_plan *plan, int dev, long long *first_col, long long *last_col);            /* device pointer, n doubles      */
double *sblas_spmv_plan_y(sblas_spmv_plan *plan, int dev, int *first_row, int *rows); /* device y slice */
const int *sblas_spmv_plan_rowptr(sblas_spmv_plan *plan, int dev, int *count);        /* device int32   */
void *sblas_spmv_plan_stream(sblas_spmv_plan *plan, int dev);         /* cudaStream_t of the GPU        */
/* raw partial sums of split rows after an execute: out[2*seg] (first row) and
 * out[2*seg+1] (last row); valid where the segment's start_flag / end_flag is set */
int sblas_spmv_plan_edges(sblas_spmv_plan *plan, double *out);
/* device pointer of the edge table of one GPU (2 doubles per local segment) */
double *sblas_spmv_plan_edge_ptr(sblas_spmv_plan *plan, int dev);
/* Host-side layout of a plan, without touching a GPU (flags |= SBLAS_LAYOUT_ONLY in
 * sblas_spmv_plan_create_rank builds only this): the segments this process runs and the
 * merge lists of the split rows it owns.  Used by the CPU tests of the multi-rank path.
 *   local_segment: out = {global segment, first row, last row, first nnz, end nnz (exclusive),
 *                         first row shared, last row shared, edge slot, GPU index, GPU's first row}
 *   merge_list:    row i (GPU-local index mrow[i]) = alpha * sum of table[msrc_off[k]],
 *                  k in [mbeg[i], mbeg[i+1]), + beta*y; table = rank-major gathered edges */
int sblas_spmv_plan_local_segments(const sblas_spmv_plan *plan);
int sblas_spmv_plan_local_segment(const sblas_spmv_plan *plan, int i, long long out[10]);
int sblas_spmv_plan_merge_list(const sblas_spmv_plan *plan, int dev, int *nmerge, const int **mrow,
                               const int **mbeg, const long long **msrc_off);

/* Rank plans: number of doubles each rank contributes to the exchange of split-row
 * partials (2 per local segment, padded to the largest rank), and the merge that
 * finishes the rows this rank owns from a table holding every rank's block
 * (world x edge_slots doubles, rank-major: the output of an all-gather of each
 * rank's sblas_spmv_plan_edge_ptr block, or a peer-mapped symmetric buffer).
 * Enqueued on the plan's stream; ascending segment order (deterministic). */
int sblas_spmv_plan_edge_slots(const sblas_spmv_plan *plan);
int sblas_spmv_plan_merge_gathered(sblas_spmv_plan *plan, const double *gathered, double alpha, double beta);
/* Rank plans: make the local segments write their edge partials into caller-owned
 * device memory (edge_slots doubles), e.g. a torch tensor that is all-gathered or a
 * symmetric-memory buffer peers can read over NVLink. */
int sblas_spmv_plan_bind_edge_table(sblas_spmv_plan *plan, double *device_block);
/* blocking copy helper for callers without a CUDA binding of their own
 * (kind: 1 host->device, 2 device->host, 3 device->device); returns the cudaError_t */
int sblas_memcpy(void *dst, const void *src, unsigned long long bytes, int kind);
int sblas_device_synchronize(void);
/* Rank plans, fused exchange: every rank owns a peer-mapped buffer of
 * 2*table_words + 2*world 8-byte words (zero-initialised; table_words >= world*edge_slots),
 * e.g. from torch.distributed._symmetric_memory; peer_bases[r] is rank r's buffer as mapped in
 * THIS process.  After binding, sblas_spmv_plan_execute_device writes split-row partials into
 * the table and sblas_spmv_plan_exchange_merge does the whole exchange on the GPU (P2P stores
 * + flags over NVLink, see include/sblas_device.h): no NCCL call, no host synchronisation. */
int sblas_spmv_plan_bind_peer_tables(sblas_spmv_plan *plan, void *const *peer_bases, long long table_words);
int sblas_spmv_plan_exchange_merge(sblas_spmv_plan *plan, double alpha, double beta);
/* phase 1 = publish only, 2 = merge only (0 = both): lets tests step several rank plans that
 * share one GPU without any kernel waiting on a kernel that has not been launched */
int sblas_spmv_plan_exchange_merge_phase(sblas_spmv_plan *plan, double alpha, double beta, int phase);
/* algorithmic bytes of one execute: 12*nnz + 4*(rows+1)
 * + 8*x_touched + 8*rows*(1 + [beta != 0]), summed over the plan's GPUs */
double sblas_spmv_plan_alg_bytes(const sblas_spmv_plan *plan, int beta_nonzero, long long x_touched_per_gpu);
/* number of kernel launches one execute enqueues (all GPUs of the plan) */
int sblas_spmv_plan_launches(const sblas_spmv_plan *plan);
/* Row panels (one kernel launch each; adaptive row binning, sblas_plan.c).  unit i ->
 * out = {local segment, kernel kind (SBLAS_K_*), ipt / R | window << 8, first row, last row
 *        (global, inclusive), first entry, one-past-last entry (global), launches}. */
int sblas_spmv_plan_num_units(const sblas_spmv_plan *plan);
/* The row-binning rule on host arrays (no CUDA): block b of 4096 rows has longest row
 * block_longest[b] and starts at entry block_first_entry[b]; outputs one (class, R, first block)
 * per run, class 0 general / 1 short / 2 medium, run_begin[nruns] = nblocks; returns nruns.
 * All output arrays need nblocks + 2 ints. */
int sblas_bin_row_blocks(const int *block_longest, const int *block_first_entry, int nblocks, int nrows, int nz_end,
                         int short_max, int medium_on, long long min_nnz, int *run_class, int *run_R, int *run_begin);
int sblas_spmv_plan_unit(const sblas_spmv_plan *plan, int i, long long out[8]);
/* launch ONE panel's kernel(s) on its stream (measurement: time a kernel alone) */
int sblas_spmv_plan_execute_unit(sblas_spmv_plan *plan, int i, double alpha, double beta);
void sblas_spmv_plan_destroy(sblas_spmv_plan *plan);

const char *sblas_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* SBLAS_SPMV_H */
