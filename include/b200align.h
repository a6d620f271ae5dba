/*
 * b200align.h -- thin C ABI of libb200align.so, the Blackwell (sm_100a) replacement for the device side of
 * MASA-CUDAlign's aligner extension.
 *
 * Path shorthand in the citations:  R/ = masa-cudalign-4.0.2.1028/,  C/ = R/libs/masa-core/src/.
 *
 * Two groups of entry points:
 *
 *  (1) "diag" primitives.  One call per protected virtual that reference CUDAligner implements for
 *      AbstractDiagonalAligner (R/src/CUDAligner.hpp:216-232): the reference-side C++ adapter
 *      (masa-cudalign_b200/host/B200Aligner.cpp) forwards each virtual to one of these, so MASA-Core's own
 *      host policy (C/libmasa/aligners/AbstractDiagonalAligner.cpp:59-501) keeps deciding which rows, columns
 *      and scores leave the aligner.  Used for stages 2/3 (goal matching, early stop) and as the
 *      compatibility path of stage 1.
 *
 *  (2) b200_align_partition: the B200-first path.  The whole partition is aligned by one persistent kernel
 *      (warp-per-strip chained wavefront, flag-gated borders in L2, on-device special-row area, exact
 *      best-cell tracking); the results are the same artefacts the reference dispatches through IManager
 *      (C/libmasa/IManager.hpp:98-313), delivered through the callbacks below after the kernel finishes.
 *
 * Conventions: all functions return 0 on success, non-zero on error (b200_last_error() describes it); no
 * exceptions, no torch types, plain pointers and sizes.  Cells are the reference's cell_t
 * (C/libmasa/libmasaTypes.hpp:35-41): a cell travelling in a ROW carries (H,F), in a COLUMN (H,E).
 * Coordinates are 0-based indices relative to the pointers given to b200_set_sequences, exactly like the
 * Partition the reference hands to IAligner::alignPartition (C/common/AlignerManager.cpp:165,177).
 * Scores: match +1, mismatch -3, gap open 3, gap extend 2 (R/src/CUDAligner.hpp:77-98) -- compile-time, as
 * in the reference (variable_penalties = NOT_SUPPORTED, R/src/CUDAligner.cpp:102).
 */
#ifndef B200ALIGN_H
#define B200ALIGN_H

#ifdef __cplusplus
extern "C" {
#endif

#define B200_INF 999999999            /* C/libmasa/libmasaTypes.hpp:46 */

#define B200_NEEDLEMAN_WUNSCH 0       /* == NEEDLEMAN_WUNSCH, C/libmasa/IManager.hpp:31 */
#define B200_SMITH_WATERMAN 1         /* == SMITH_WATERMAN,   C/libmasa/IManager.hpp:33 */

/* first row / first column sources: values of C/libmasa/IManager.hpp:38-47, cells of C/common/io/InitialCellsReader.cpp:84-108 */
#define B200_INIT_ZEROES 0            /* h = 0,                   e/f = -INF */
#define B200_INIT_GAPS 1              /* h = -ext*pos - open,     e/f = -INF; pos 0 -> h = 0 */
#define B200_INIT_CUSTOM 2            /* == INIT_WITH_CUSTOM_DATA (:47): cells supplied by the caller */
#define B200_INIT_GAPS_OPENED 3       /* == INIT_WITH_GAPS_OPENED (:44): h = -ext*pos, e/f = -INF */

/* kernel selection */
#define B200_KERNEL_AUTO 0
#define B200_KERNEL_S32 1             /* exact int32 lanes, byte compare: any alphabet, any border values */
#define B200_KERNEL_S16X2 2           /* packed s16x2 DPX lanes with per-block rebasing; ACGT only */

typedef struct { int h; int x; } b200_cell;          /* == cell_t: x is F in rows, E in columns */
typedef struct { int i; int j; int score; } b200_score;   /* == score_t field for field, 0-based (C/libmasa/libmasaTypes.hpp:88-95) */
typedef struct { int found; int k; int score; int type; } b200_match;   /* == match_result_t (:51-60) */

typedef struct {
	int device;            /* CUDA device ordinal (reference: --gpu, R/src/CUDAlignerParameters.cpp:33-54) */
	int kernel;            /* B200_KERNEL_* */
	int warps_per_sm;      /* resident strip-warps per SM for the persistent kernel; 0 = default */
	int reserved[5];
} b200_config;

typedef struct {
	int i0, j0, i1, j1;          /* rows [i0,i1) of seq0, columns [j0,j1) of seq1 (Partition.hpp) */
	int recurrence;              /* B200_SMITH_WATERMAN | B200_NEEDLEMAN_WUNSCH (IManager::getRecurrenceType) */
	int first_row_init;          /* B200_INIT_* (IManager::getFirstRowInitType)    */
	int first_col_init;          /* B200_INIT_* (IManager::getFirstColumnInitType) */
	int special_row_interval;    /* IManager::getSpecialRowInterval; <=0 disables special rows */
	int block_height;            /* row granularity of the special-row policy: 4*min(128,width) in the reference
	                                (R/src/CUDAligner.cpp:295-297); 0 = that default */
	int want_special_rows;       /* IManager::mustDispatchSpecialRows */
	int want_last_row;           /* IManager::mustDispatchLastRow     */
	int want_last_column;        /* IManager::mustDispatchLastColumn  */
	int want_best_score;         /* IManager::mustDispatchScores: exact best cell of the partition */
	int prune;                   /* IManager::mustPruneBlocks */
	int super_i1, super_j1;      /* IManager::getSuperPartition: bounds used by the pruning test */
	int reserved[4];             /* [0] flags: B200_MGPU_CHAIN | B200_CONT_CHUNK; [1] chunk width of a chained call;
	                                [2] rows of the partition already aligned by earlier chunk calls; [3] total rows of
	                                the partition (0 = i1-i0): the special-row policy is applied to the WHOLE partition */
} b200_partition;

/* B200_CONT_CHUNK: this call continues the previous b200_align_partition call of the same handle one chunk of rows
 * further down (same columns): the top border is taken from the device (it is the previous chunk's last row), the
 * corner cells are not read again, receive_first_column continues where it stopped, and the first cell of the last
 * column is not dispatched again.  Used by the adapter to run stage-2/3 partitions (goal matching on the last column,
 * early stop) as a few persistent launches instead of one launch per external diagonal. */
#define B200_CONT_CHUNK 2

/* Callbacks == the IManager methods the reference aligner calls (C/libmasa/IManager.hpp:150-313).
 * Buffers passed to dispatch_* are borrowed for the duration of the call. May be NULL when not needed. */
typedef struct {
	void* ctx;
	void (*receive_first_row)(void* ctx, b200_cell* buffer, int len);
	void (*receive_first_column)(void* ctx, b200_cell* buffer, int len);
	void (*dispatch_row)(void* ctx, int i, const b200_cell* buffer, int len);
	void (*dispatch_column)(void* ctx, int j, const b200_cell* buffer, int len);
	void (*dispatch_score)(void* ctx, b200_score score);
	int  (*must_continue)(void* ctx);
} b200_callbacks;

typedef struct {
	b200_score best;             /* lexicographically smallest (i,j) among maximal cells, or score=-INF */
	long long cells;             /* DP cells actually computed (pruned blocks excluded) */
	long long cells_total;       /* (i1-i0)*(j1-j0) */
	double device_ms;            /* CUDA-event time of the alignment kernels */
	int strips;                  /* strip jobs executed */
	int kernel_launches;         /* kernels launched by this call */
	int kernel_used;             /* B200_KERNEL_S32 | B200_KERNEL_S16X2 */
	int reserved[5];             /* [0] column chunks and [1] widest chunk of a chained call; [2] share of the resident warps'
	                                time spent inside compute segments, per mille (packed kernel); [3] resident warps */
} b200_result;

typedef struct b200_handle b200_handle;

/* lifetime (IAligner::initialize / finalize, R/src/CUDAligner.cpp:137-164,579-586) */
int b200_create(const b200_config* cfg, b200_handle** out);
void b200_destroy(b200_handle* h);
const char* b200_last_error(const b200_handle* h);      /* h may be NULL: error of the last failed b200_create */
int b200_device_count(void);                             /* --list-gpus (R/src/CUDAlignerParameters.cpp) */

/* IAligner::setSequences / unsetSequences (R/src/CUDAligner.cpp:229-283): host pointers, uploaded here */
int b200_set_sequences(b200_handle* h, const char* seq0, int seq0_len, const char* seq1, int seq1_len);
int b200_unset_sequences(b200_handle* h);

/* (2) B200-first whole-partition path (replaces the loop of AbstractDiagonalAligner.cpp:59-159 plus
 *     R/src/CUDAligner.cu:745-1156 for partitions that need no early stop). */
int b200_align_partition(b200_handle* h, const b200_partition* p, const b200_callbacks* cb, b200_result* out);

/* (1) diag primitives: one per CUDAligner virtual (R/src/CUDAligner.hpp:216-232). */
int b200_diag_begin(b200_handle* h, const b200_partition* p, int grid_width, const int* split /* grid_width+1 */,
                    int block_height);                                             /* initializeDiagonals  */
int b200_diag_set_first_row(b200_handle* h, const b200_cell* cells, int j, int len);    /* setFirstRow     */
int b200_diag_set_first_column(b200_handle* h, const b200_cell* cells, int i, int len); /* setFirstColumn: cells[0]=diag, [1..block_height] */
int b200_diag_process(b200_handle* h, int diagonal, int window_left, int window_right); /* processDiagonal */
int b200_diag_get_row(b200_handle* h, int j, int len, b200_cell* out);     /* getSpecialRow / getLastRow   */
int b200_diag_get_last_column(b200_handle* h, int i, int len, b200_cell* out);          /* getLastColumn   */
int b200_diag_get_block_scores(b200_handle* h, b200_score* out /* grid_width */);       /* getBlockScores  */
int b200_diag_clear_pruned(b200_handle* h, int j0, int j1);                             /* clearPrunedBlocks (busH <- -INF) */
int b200_diag_end(b200_handle* h);                                                      /* finalizeDiagonals */

/* IAligner::matchLastColumn (C/libmasa/utils/AlignerUtils.cpp:50-107): goal matching of a last-column chunk
 * against a reversed special-row chunk; first k wins, match before gap. Runs on the device. */
int b200_match_last_column(b200_handle* h, const b200_cell* buffer, const b200_cell* base, int len, int goal,
                           b200_match* out);

/* Multi-GPU chained wavefront on one NVLink/NVSwitch box.  The reference splits seq1 into one contiguous column
 * slice per process (--fork/--split, C/libmasa/libmasa.cpp:540-642) and streams the slice border through TCP sockets
 * (SocketCellsWriter/Reader + Buffer2, C/stage1/sw_stage1.cpp:168-186); block pruning is switched off in that mode
 * (libmasa.cpp:1318-1321).  Here the columns are cut into CHUNKS dealt round-robin to the GPUs (chunk c belongs to GPU
 * c mod world): every strip of rows travels GPU 0 -> 1 -> ... -> world-1 -> 0 -> ... once per chunk.  The strip kernel
 * stores the right border of a (strip, chunk) job straight into the exchange block of the next GPU (peer memory, P2P
 * stores over NVLink) and then counts a "left event" on that strip's event word over there; the job on the right is
 * pushed into that GPU's work queue as soon as its left border AND the first columns of the strip above exist, and
 * resident warps pop jobs from the queue (dataflow scheduling, no host, no socket, no collective on the data path).
 * Round-robin chunks spread the cells that survive block pruning evenly over the GPUs (the reference's static slices
 * cannot: README "MultiBP"), and a pipeline hop costs one chunk sweep instead of one slice sweep.  The running best
 * score is pushed to every peer (peer atomicMax), replacing AlignerPool's file+signal polling
 * (C/common/AlignerPool.cpp:46-68), so pruning stays ON.
 *
 * One process per GPU (bench.py, tests/mgpu_check.py under torchrun):
 *   1. every rank: b200_chain_plan(partition, world, &info); b200_mgpu_export(h, rows, info.max_jobs, &mine)
 *   2. exchange the 64-byte handles out of band (torch.distributed all_gather)
 *   3. every rank: b200_mgpu_connect(h, rank, world, all_handles)
 *   4. every rank: b200_align_partition with the WHOLE partition and b200_partition.reserved[0] = B200_MGPU_CHAIN;
 *      reserved[1] = chunk width in columns (0 = automatic, < 0 = one contiguous slice per GPU with the reference's
 *      --split arithmetic, libmasa.cpp:632-635).  Each rank gets the artefacts of ITS chunks: dispatch_row delivers
 *      [first-column cell, rank 0 only] and then one call per owned chunk in column order; dispatch_column / the last
 *      row's tail come from the owner of the last chunk; b200_result.best is the best cell of the rank's own chunks.
 *      Consecutive chained calls must be separated by a barrier over all ranks.
 * One process for all GPUs (build/cudalign --gpus=N): b200_group_* below; same kernels, peers mapped with
 * cudaDeviceEnablePeerAccess, artefacts delivered exactly like the single-GPU call (whole rows, merged best). */
typedef struct { unsigned char bytes[64]; } b200_ipc_handle;
typedef struct {
	int chunks;                  /* column chunks of the partition */
	int chunk_cols;              /* width of a chunk (the last one may be narrower) */
	int chunks_per_gpu;          /* most chunks owned by one GPU */
	int reserved0;
	long long max_strips;        /* upper bound of the strips of the partition */
	long long max_jobs;          /* capacity to pass to b200_mgpu_export / b200_group_create */
} b200_chain_info;
#define B200_MGPU_CHAIN 1
int b200_chain_plan(const b200_partition* p, int world, b200_chain_info* out);      /* host only, no GPU needed */
int b200_mgpu_export(b200_handle* h, long long max_rows, long long max_jobs, b200_ipc_handle* out);
int b200_mgpu_connect(b200_handle* h, int rank, int world, const b200_ipc_handle* all_handles);
int b200_mgpu_disconnect(b200_handle* h);
int b200_last_chain_result(const b200_handle* h, b200_result* out);   /* this GPU's share of the last chained call */

typedef struct b200_group b200_group;
int b200_group_create(const int* devices, int n, const b200_config* cfg, long long max_rows, long long max_jobs, b200_group** out);
void b200_group_destroy(b200_group* g);
const char* b200_group_last_error(const b200_group* g);
int b200_group_size(const b200_group* g);
b200_handle* b200_group_handle(b200_group* g, int rank);            /* rank 0 serves the single-GPU calls (stages 2-4) */
int b200_group_set_sequences(b200_group* g, const char* seq0, int seq0_len, const char* seq1, int seq1_len);
int b200_group_align_partition(b200_group* g, const b200_partition* p, const b200_callbacks* cb, b200_result* out);
int b200_group_rank_result(const b200_group* g, int rank, b200_result* out);        /* cells / device time per GPU */

/* Stage 4: batched Myers-Miller partition split on the GPU (replaces reduce_partitions/split_thread/ort_split_2 of
 * C/stage4/sw_stage4.cpp:87-380,806-852, 4 pthreads in the reference).  Crosspoints are the reference's
 * crosspoint_t (C/common/Crosspoint.hpp:30-40): 0-based prefix lengths (i, j), type 0 MATCH / 1 GAP_1 / 2 GAP_2.
 * The sequences are the ones given to b200_set_sequences (whole trimmed sequences, like stage 4's
 * seq->getData()).  Only the default OPTIMIZED strategy (ort_split_2) is implemented.
 *   b200_stage4_round : one round; out[k] (k = 1..n-1) = midpoint of partition (k-1, k) or type = -1 (not split)
 *   b200_stage4       : rounds + merge_partitions (:785-804) until the largest partition <= max_partition
 *                       (:926-945); writes at most cap crosspoints, *n_out = count.
 * Return 6 = the reference's fatal conditions ("Error Match" / "NOT FOUND"). */
typedef struct { int i; int j; int type; int score; } b200_xpoint;
int b200_stage4_round(b200_handle* h, const b200_xpoint* in, int n, int max_partition, b200_xpoint* out);
int b200_stage4(b200_handle* h, const b200_xpoint* in, int n, int max_partition, b200_xpoint* out, int cap, int* n_out);

/* Stage 5: batched traceback on the GPU (replaces the serial partition loop of C/stage5/sw_stage5.cpp:86-319,404-424:
 * one CPU thread, static 1024 x 1024 tables).  pts = the stage-4 crosspoints (crosspoint_04.NN); partition k (k = 1..n-1)
 * lies between pts[k-1] and pts[k].  Every partition is walked back from its bottom-right corner with the reference's
 * precedence (diagonal, then vertical, then horizontal; :222-257) and leaves one byte per step:
 *   0 diagonal (match / mismatch), 1 vertical = the reference's dot(..., 1): gap in seq1, 2 horizontal = dot(..., 2): gap in seq0.
 * Partition k owns the slots ops[off_k .. off_k + op_len[k]) with off_k = (pts[k-1].i - pts[0].i) + (pts[k-1].j - pts[0].j),
 * so ops needs (pts[n-1].i - pts[0].i) + (pts[n-1].j - pts[0].j) bytes; op_len has n entries (op_len[0] = 0).
 * Replaying the bytes of partitions 1, 2, ... in order into Alignment::addGapInSeq0/1 (dot(), :70-84) reproduces the
 * reference's alignment.NN.bin byte for byte (host/stage5_gpu.cpp does exactly that).  *total == total_score_t (:51-67)
 * summed over all partitions.  Sequences: the ones given to b200_set_sequences (whole sequences, like stage 4). */
typedef struct { int score, matches, mismatches, gap_open, gap_ext; } b200_s5_stats;
int b200_stage5(b200_handle* h, const b200_xpoint* pts, int n, unsigned char* ops, long long ops_cap, int* op_len,
                b200_s5_stats* total);

/* Host-side policy helper (no GPU needed): the special-row ids (rows above, relative to i0) that the reference
 * flushes for a partition of `height` rows: AbstractDiagonalAligner::isSpecialRow,
 * C/libmasa/aligners/AbstractDiagonalAligner.cpp:466-478 (8192-row floor, top and bottom rows excluded).
 * Writes at most `cap` ids to `out`, returns the total count. */
int b200_special_row_ids(int height, int block_height, int interval, int* out, int cap);

/* statistics (IAligner::getProcessedCells etc.) */
long long b200_processed_cells(const b200_handle* h);
long long b200_kernel_launches(const b200_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* B200ALIGN_H */
