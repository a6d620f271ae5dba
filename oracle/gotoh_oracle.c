/* oracle/gotoh_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product
 * path (masa-cudalign_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may use it.
 *
 * Plain-C restatement of the reference's CPU algorithm for the hot path (C/ = masa-core/src):
 *   - Gotoh cell            C/libmasa/processors/CPUBlockProcessor.cpp:66-93  (sw()/nw())
 *   - border conventions    C/libmasa/processors/CPUBlockProcessor.cpp:96-174 (row cells carry (H,F), column cells (H,E))
 *   - first row/col init    C/common/io/InitialCellsReader.cpp:84-108
 *   - best-cell rule        CPUBlockProcessor.cpp:154-158 (strict <, row-major) + C/common/BestScoreList.hpp:30-38
 *                           => lexicographically smallest (i,j) among maximal cells
 *   - goal matching         C/libmasa/utils/AlignerUtils.cpp:50-107
 *   - pruning bound         C/libmasa/pruning/AbstractBlockPruning.cpp:70-113
 * Parity pinning: the reference ships no golden vectors; this file is pinned against
 * the reference's own code compiled from /root/reference (oracle/_ref/oracle_cpu, oracle_cpu_block) by
 * tests/test_oracle_pinning.py and the committed fixtures in tests/golden/.
 */
#include "gotoh_oracle.h"
#include <stdlib.h>
#include <string.h>

#define MATCH 1
#define MISMATCH (-3)
#define GAP_OPEN 3
#define GAP_EXT 2
#define MAX2(a, b) ((a) > (b) ? (a) : (b))

void go_init_cells(go_cell* buf, int len, int type, int start_pos) {
    for (int k = 0; k < len; k++) {
        int pos = start_pos + k;
        if (type == GO_INIT_ZEROES) { buf[k].h = 0; buf[k].x = -GO_INF; }
        else if (pos == 0)          { buf[k].h = 0; buf[k].x = -GO_INF; }
        else { buf[k].h = -GAP_EXT * pos - (type == GO_INIT_GAPS ? GAP_OPEN : 0); buf[k].x = -GO_INF; }
    }
}

/* Full matrix, row by row, O(n) memory.  first_row[k] = (H,F) of cell (-1, k-1) for k = 0..n (k = 0 is the
 * corner); first_col[k] = (H,E) of cell (k-1, -1) for k = 0..m.  When NULL they are generated from *_type.
 * rows_out[r*(n+1) + k] receives (H,F) of cell (row_ids[r], k-1); element 0 is the first-column cell with
 * F = -INF, as AbstractDiagonalAligner::flushSpecialRows dispatches it (AbstractDiagonalAligner.cpp:290-298).
 * last_col_out[k] = (H,E) of cell (k-1, n-1), element 0 being the first-row tail with x = -INF (:419-422). */
int go_full_matrix(const unsigned char* s0, int m, const unsigned char* s1, int n, int recurrence,
                   const go_cell* first_row, int first_row_type, const go_cell* first_col, int first_col_type,
                   const int* row_ids, int n_rows, go_cell* rows_out, go_cell* last_col_out, go_score* best_out) {
    go_cell* row = (go_cell*)malloc(sizeof(go_cell) * (size_t)(n + 1));
    go_cell* col = (go_cell*)malloc(sizeof(go_cell) * (size_t)(m + 1));
    if (!row || !col) return -1;
    if (first_row) memcpy(row, first_row, sizeof(go_cell) * (size_t)(n + 1)); else go_init_cells(row, n + 1, first_row_type, 0);
    if (first_col) memcpy(col, first_col, sizeof(go_cell) * (size_t)(m + 1)); else go_init_cells(col, m + 1, first_col_type, 0);
    go_score best; best.score = -GO_INF; best.i = -1; best.j = -1;
    if (last_col_out) { last_col_out[0].h = row[n].h; last_col_out[0].x = -GO_INF; }
    int next_row = 0;
    for (int i = 0; i < m; i++) {
        int h11 = (i == 0) ? row[0].h : col[i].h;    /* diagonal H[i-1][-1] */
        int h01 = col[i + 1].h, e00 = col[i + 1].x;  /* H[i][-1], E[i][-1]  */
        const unsigned char c = s0[i];
        for (int j = 0; j < n; j++) {
            int h10 = row[j + 1].h, f10 = row[j + 1].x;
            e00 = MAX2(h01 - GAP_OPEN, e00) - GAP_EXT;
            f10 = MAX2(h10 - GAP_OPEN, f10) - GAP_EXT;
            int v1 = h11 + ((s1[j] != c) ? MISMATCH : MATCH);
            int h00 = MAX2(MAX2(v1, e00), f10);
            if (recurrence == GO_SW) h00 = MAX2(h00, 0);
            h11 = h10; h01 = h00;
            row[j + 1].h = h00; row[j + 1].x = f10;
            if (best.score < h00) { best.score = h00; best.i = i; best.j = j; }
        }
        if (last_col_out) { last_col_out[i + 1].h = h01; last_col_out[i + 1].x = e00; }
        row[0].h = col[i + 1].h; row[0].x = -GO_INF;
        while (next_row < n_rows && row_ids[next_row] == i) {
            memcpy(rows_out + (size_t)next_row * (n + 1), row, sizeof(go_cell) * (size_t)(n + 1));
            next_row++;
        }
    }
    if (best_out) *best_out = best;
    free(row); free(col);
    return 0;
}

go_match go_match_column(const go_cell* buffer, const go_cell* base, int len, int goal, int gap_open) {
    go_match r; r.found = 0; r.k = -1; r.score = 0; r.type = 0;
    for (int k = 0; k < len; k++) {
        int sum_match = base[k].h + buffer[k].h;
        int sum_gap = base[k].x + buffer[k].x + gap_open;
        if (sum_match == goal)      { r.found = 1; r.k = k; r.score = base[k].h; r.type = 0; return r; }
        else if (sum_gap == goal)   { r.found = 1; r.k = k; r.score = base[k].x; r.type = 1; return r; }
        else if (sum_match > goal || sum_gap > goal) { r.k = k; r.type = sum_match > goal ? -1 : -2; return r; }
    }
    return r;
}

int go_block_prunable(int score, int best, int i0, int j0, int i1, int j1, int max_i, int max_j, int recurrence) {
    int distI = max_i - i0, distJ = max_j - j0;
    int distMin = distI < distJ ? distI : distJ;
    int inc = distMin * MATCH;
    if (recurrence == GO_NW) {
        int d = distJ - distI; if (d < 0) d = -d;
        int mx = (j1 - j0) > (i1 - i0) ? (j1 - j0) : (i1 - i0);
        int gaps = d - mx;
        if (gaps > 0) inc -= GAP_OPEN + gaps * GAP_EXT;
    }
    return (score + inc) <= best;
}
