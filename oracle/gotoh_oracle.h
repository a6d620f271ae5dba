/* oracle/gotoh_oracle.h -- TEST INFRASTRUCTURE ONLY (see gotoh_oracle.c). */
#ifndef GOTOH_ORACLE_H
#define GOTOH_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int h; int x; } go_cell;          /* x = F for row cells, E for column cells (libmasaTypes.hpp:35-41) */
typedef struct { int score, i, j; } go_score;      /* 0-based cell indices (libmasaTypes.hpp:88-95) */
typedef struct { int found, k, score, type; } go_match;   /* libmasaTypes.hpp:51-60 */
typedef struct { int i, j, type, score; } go_xpoint;      /* == crosspoint_t, common/Crosspoint.hpp:30-40 */

#define GO_INF 999999999
#define GO_NW 0   /* NEEDLEMAN_WUNSCH (libmasa/IManager.hpp:31) */
#define GO_SW 1   /* SMITH_WATERMAN   (libmasa/IManager.hpp:33) */

/* init types (common/io/InitialCellsReader.cpp:84-108) */
#define GO_INIT_ZEROES 0
#define GO_INIT_GAPS 1          /* h = -ext*pos - open  */
#define GO_INIT_CUSTOM 2        /* INIT_WITH_CUSTOM_DATA (libmasa/IManager.hpp:47) */
#define GO_INIT_GAPS_OPENED 3   /* h = -ext*pos         (libmasa/IManager.hpp:44) */

int go_full_matrix(const unsigned char* s0, int m, const unsigned char* s1, int n, int recurrence,
                   const go_cell* first_row /*n+1 or NULL*/, int first_row_type,
                   const go_cell* first_col /*m+1 or NULL*/, int first_col_type,
                   const int* row_ids, int n_rows, go_cell* rows_out /* n_rows*(n+1) */,
                   go_cell* last_col_out /* m+1 or NULL */, go_score* best_out);

void go_init_cells(go_cell* buf, int len, int type, int start_pos);

go_match go_match_column(const go_cell* buffer, const go_cell* base, int len, int goal, int gap_open);

int go_block_prunable(int score, int best, int i0, int j0, int i1, int j1, int max_i, int max_j, int recurrence);

/* Stage 4 (C/stage4/sw_stage4.cpp): Myers-Miller midpoint of one partition with the default OPTIMIZED strategy
 * (ort_split_2, :297-380) and the per-partition driver logic of split_thread (:87-217).  Crosspoints carry
 * 0-based prefix lengths (i, j), type 0 = MATCH, 1 = GAP_1, 2 = GAP_2.  Returns 0, or <0 on the reference's
 * fatal conditions ("Error Match", "NOT FOUND").  out->type = -1 when the partition is not split. */
int go_stage4_split_one(const unsigned char* seq0, const unsigned char* seq1, go_xpoint a, go_xpoint b, int max_part, go_xpoint* out);
/* One reduce_partitions round (:806-852) + merge_partitions (:785-804): returns the new number of crosspoints
 * written to out (capacity >= 2*n), or <0 on error; *changed tells whether any midpoint was inserted. */
int go_stage4_round(const unsigned char* seq0, const unsigned char* seq1, const go_xpoint* in, int n, int max_part, go_xpoint* out, int* changed);
int go_largest_partition(const go_xpoint* pts, int n);

/* Stage 5 (C/stage5/sw_stage5.cpp:86-319): traceback of the partitions between consecutive stage-4 crosspoints.
 * go_s5_stats == total_score_t (:51-67).  ops: one byte per step in walk order (bottom-right -> top-left of each partition),
 * 0 = diagonal, 1 = vertical (the reference's dot(...,1): gap in seq1), 2 = horizontal (dot(...,2): gap in seq0). */
typedef struct { int score, matches, mismatches, gap_open, gap_ext; } go_s5_stats;
int go_stage5_partition(const unsigned char* seq0, const unsigned char* seq1, go_xpoint a, go_xpoint b,
                        unsigned char* ops, go_s5_stats* st);
int go_stage5(const unsigned char* seq0, const unsigned char* seq1, const go_xpoint* pts, int n, unsigned char* ops,
              int* op_len, go_s5_stats* st);

#ifdef __cplusplus
}
#endif
#endif
