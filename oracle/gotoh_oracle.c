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
 *   - stage-4 split         C/stage4/sw_stage4.cpp:87-380,785-852 (split_thread, ort_split_2, merge_partitions)
 *   - stage-5 traceback     C/stage5/sw_stage5.cpp:86-319,404-424
 * Parity pinning: the reference ships no golden vectors; this file is pinned against
 * the reference's own code compiled from /root/reference (oracle/_ref/oracle_cpu, oracle_cpu_block) by
 * tests/test_oracle_cpu.py and the committed fixtures in tests/golden/ (reference_runs.json: stage-1 best cells, special-row
 * files, crosspoints of stages 1-4; stage5_runs.json: what the reference's stage 5 stored in alignment.00.bin, read back with
 * the reference's own reader; cfg1_stage1.json: BASELINE cfg1 at full size).
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

/* ------------------------------------------------------------------------------------------------------------
 * Stage 4 restatement (C/stage4/sw_stage4.cpp)
 * ---------------------------------------------------------------------------------------------------------- */
#define GAP_FIRST (GAP_OPEN + GAP_EXT)
#define T_MATCH 0
#define T_GAP_1 1
#define T_GAP_2 2
#define MAX3(a, b, c) (MAX2(MAX2((a), (b)), (c)))

/* one column of the half-matrix: sw_stage4.cpp:254-273 (processCol); s0 is indexed with stride (reverse = -1) */
static go_cell s4_process_col(const unsigned char* s0, int stride, unsigned char c, int h11, int h10, go_cell* col, int len) {
    int f0 = -GO_INF;
    for (int j = 0; j < len; j++) {
        col[j].x = MAX2(col[j].h - GAP_FIRST, col[j].x - GAP_EXT);
        f0 = MAX2(h10 - GAP_FIRST, f0 - GAP_EXT);
        h10 = MAX3(h11 + ((c == s0[(long)j * stride]) ? MATCH : MISMATCH), col[j].x, f0);
        h11 = col[j].h;
        col[j].h = h10;
    }
    go_cell r; r.x = f0; r.h = h10;
    return r;
}

/* sw_stage4.cpp:275-294 (match): 1 = found, 0 = not yet, -1 = "Error Match" */
static int s4_match(go_cell a, go_cell b, int diff, go_xpoint* pt) {
    int sum_match = a.h + b.h, sum_gap = a.x + b.x + GAP_OPEN;
    if (sum_match == diff) { pt->type = T_MATCH; pt->score = a.h; return 1; }
    if (sum_gap == diff) { pt->type = T_GAP_2; pt->score = a.x; return 1; }
    if (sum_match > diff || sum_gap > diff) return -1;
    return 0;
}

/* ort_split_2, sw_stage4.cpp:297-380.  s0/s1 are accessed through (base, stride) so that the transposed call of
 * split_thread (seq1 as rows) and the reversed halves need no copies. */
static int s4_ort_split_2(const unsigned char* q0, const unsigned char* q1, int i0, int j0, int i1, int j1,
                          int type_s, int type_e, int score_s, int score_e, go_xpoint* cross) {
    int len0 = i1 - i0, len1 = j1 - j0;
    int diff = score_e - score_s;
    int imid0 = len0 / 2, imid1 = len0 - imid0;
    int jmid1 = len1 - len1 / 2;
    go_cell* c0 = (go_cell*)malloc(sizeof(go_cell) * (size_t)(imid0 + 1));
    go_cell* c1 = (go_cell*)malloc(sizeof(go_cell) * (size_t)(imid1 + 1));
    go_cell* r0 = (go_cell*)malloc(sizeof(go_cell) * (size_t)(jmid1 + 2));
    go_cell* r1 = (go_cell*)malloc(sizeof(go_cell) * (size_t)(jmid1 + 2));
    for (int i = 0; i < imid0; i++) { c0[i].h = -(i + 1) * GAP_EXT - GAP_OPEN * (type_s != T_GAP_2); c0[i].x = -GO_INF; }
    for (int i = 0; i < imid1; i++) { c1[i].h = -(i + 1) * GAP_EXT - GAP_OPEN; c1[i].x = -GO_INF; }
    r0[0].h = r0[0].x = c0[imid0 - 1].h;
    r1[0].h = r1[0].x = c1[imid1 - 1].h;
    int d0 = (type_s != T_MATCH) ? -GO_INF : 0, d1 = (type_e != T_MATCH) ? -GO_INF : 0;
    int rc = -2;   /* NOT FOUND */
    for (int j = 0; j < len1; j++) {
        int h0 = -(j + 1) * GAP_EXT - GAP_OPEN * (type_s != T_GAP_1);
        go_cell rr0 = s4_process_col(q0 + i0, 1, q1[j0 + j], d0, h0, c0, imid0);
        d0 = h0;
        int h1 = -(j + 1) * GAP_EXT - GAP_OPEN;
        go_cell rr1 = s4_process_col(q0 + i1 - 1, -1, q1[j1 - 1 - j], d1, h1, c1, imid1);
        d1 = h1;
        if (j + 1 <= jmid1) { r0[j + 1] = rr0; r1[j + 1] = rr1; }
        if (j + 1 >= jmid1) {
            int m = s4_match(rr0, r1[len1 - (j + 1)], diff, cross);
            if (m < 0) { rc = -1; break; }
            if (m) { cross->j = j0 + (j + 1); cross->i = imid0 + i0; cross->score += score_s; rc = 0; break; }
            m = s4_match(r0[len1 - (j + 1)], rr1, diff, cross);
            if (m < 0) { rc = -1; break; }
            if (m) { cross->j = j0 + (len1 - (j + 1)); cross->i = imid0 + i0; cross->score += score_s; rc = 0; break; }
        }
    }
    free(c0); free(c1); free(r0); free(r1);
    return rc;
}

int go_stage4_split_one(const unsigned char* seq0, const unsigned char* seq1, go_xpoint a, go_xpoint b, int max_part, go_xpoint* out) {
    static const int inv_type[3] = {0, 2, 1};
    int di = b.i - a.i, dj = b.j - a.j;
    out->type = -1; out->i = out->j = out->score = 0;
    if (di == 0 || dj == 0) return 0;
    if (di < dj) {                                   /* transposed: sw_stage4.cpp:135-171 */
        if (!(a.j < b.j - max_part)) return 0;
        go_xpoint t;
        int rc = s4_ort_split_2(seq1, seq0, a.j, a.i, b.j, b.i, inv_type[a.type], inv_type[b.type], a.score, b.score, &t);
        if (rc) return rc;
        out->i = t.j; out->j = t.i; out->type = inv_type[t.type]; out->score = t.score;
    } else {
        if (!(a.i < b.i - max_part)) return 0;
        go_xpoint t;
        int rc = s4_ort_split_2(seq0, seq1, a.i, a.j, b.i, b.j, a.type, b.type, a.score, b.score, &t);
        if (rc) return rc;
        *out = t;
    }
    return 0;
}

int go_stage4_round(const unsigned char* seq0, const unsigned char* seq1, const go_xpoint* in, int n, int max_part, go_xpoint* out, int* changed) {
    int cnt = 0;
    *changed = 0;
    out[cnt++] = in[0];
    for (int k = 1; k < n; k++) {
        go_xpoint np;
        int rc = go_stage4_split_one(seq0, seq1, in[k - 1], in[k], max_part, &np);
        if (rc) return rc;
        int diff_pos = (np.i != in[k - 1].i || np.j != in[k - 1].j);     /* merge_partitions, :785-804 */
        if (np.type != -1 && diff_pos) { *changed = 1; out[cnt++] = np; }
        out[cnt++] = in[k];
    }
    return cnt;
}

int go_largest_partition(const go_xpoint* pts, int n) {                 /* CrosspointsFile.cpp:71-92 */
    int mi = 0, mj = 0;
    for (int k = 1; k < n; k++) {
        int di = abs(pts[k - 1].i - pts[k].i), dj = abs(pts[k - 1].j - pts[k].j);
        if (di != 0 && dj != 0) { if (mi < di) mi = di; if (mj < dj) mj = dj; }
    }
    return mi > mj ? mi : mj;
}

/* ---------------------------------------------------------------------------------------------------------
 * Stage 5 (C/stage5/sw_stage5.cpp:86-319): traceback of one stage-4 partition.  Full (H, E, F) tables of the
 * partition [a, b) with the border rules of :134-143 (NB: stage 5 names the VERTICAL gap state e and the
 * horizontal one f, :154-155), then the walk of :209-305 from the bottom-right corner:
 *   state MATCH: diagonal first, then vertical (e), then horizontal (f)  (:222-241)
 *   state GAP_2: forced vertical step, GAP_1: forced horizontal step     (:242-257)
 *   a gap step returns to MATCH when the gap was opened in that cell (e == h_above - first / f == h_left - first)
 *   leftovers: vertical steps while i > 0, then horizontal steps while j > 0 (:291-308)
 * An end crosspoint of type MATCH makes the reference add one extra row and column (:117-120) that the walk
 * leaves immediately (:197-200) without looking at them; they are not computed here.
 * ops[k] = direction of step k (0 diagonal, 1 vertical = gap in seq1, 2 horizontal = gap in seq0), in walk order;
 * at most (b.i-a.i)+(b.j-a.j) steps.  st accumulates exactly like total_score_t (:51-67). */
int go_stage5_partition(const unsigned char* seq0, const unsigned char* seq1, go_xpoint a, go_xpoint b,
                        unsigned char* ops, go_s5_stats* st) {
    const int first = GAP_OPEN + GAP_EXT;
    int di = b.i - a.i, dj = b.j - a.j, n = 0;
    if (di < 0 || dj < 0) return -1;
    if (di == 0) {                                           /* :88-99: pure horizontal run */
        int sum = -dj * GAP_EXT;
        if (a.type != T_GAP_1) { st->gap_open++; sum -= GAP_OPEN; }
        for (int j = 0; j < dj; j++) { ops[n++] = 2; st->gap_ext++; }
        st->score += sum;
        return n;
    }
    if (dj == 0) {                                           /* :100-112: pure vertical run */
        int sum = -di * GAP_EXT;
        if (a.type != T_GAP_2) { st->gap_open++; sum -= GAP_OPEN; }
        for (int i = 0; i < di; i++) { ops[n++] = 1; st->gap_ext++; }
        st->score += sum;
        return n;
    }
    const size_t W = (size_t)dj + 1;
    int* H = (int*)malloc(sizeof(int) * W * ((size_t)di + 1));
    int* E = (int*)malloc(sizeof(int) * W * ((size_t)di + 1));
    int* F = (int*)malloc(sizeof(int) * W * ((size_t)di + 1));
    const unsigned char* s0 = seq0 + a.i;
    const unsigned char* s1 = seq1 + a.j;
    for (int j = 1; j <= dj; j++) { H[j] = -j * GAP_EXT - GAP_OPEN * (a.type != T_GAP_1); E[j] = -GO_INF; F[j] = -GO_INF; }
    H[0] = a.type != T_MATCH ? -GO_INF : 0;
    E[0] = -GO_INF; F[0] = -GO_INF;
    for (int i = 1; i <= di; i++) {
        int* h0 = H + W * i; int* h1 = h0 - W; int* e0 = E + W * i; int* e1 = e0 - W; int* f0 = F + W * i;
        h0[0] = -i * GAP_EXT - GAP_OPEN * (a.type != T_GAP_2);
        f0[0] = -GO_INF; e0[0] = -GO_INF;
        for (int j = 1; j <= dj; j++) {
            e0[j] = MAX2(h1[j] - first, e1[j] - GAP_EXT);
            f0[j] = MAX2(h0[j - 1] - first, f0[j - 1] - GAP_EXT);
            int d = h1[j - 1] + (s0[i - 1] == s1[j - 1] ? MATCH : MISMATCH);
            h0[j] = MAX2(d, MAX2(e0[j], f0[j]));
        }
    }
    int i = di, j = dj, c = b.type, sum = 0;
    while (i > 0 && j > 0) {
        const int hh = H[W * i + j], ee = E[W * i + j], ff = F[W * i + j];
        const int same = s0[i - 1] == s1[j - 1];
        int dir;
        if (c == T_MATCH) {
            if (hh == H[W * (i - 1) + j - 1] + (same ? MATCH : MISMATCH)) dir = 0;
            else if (hh == ee) dir = 1;
            else dir = 2;                                   /* hh == ff: H is the maximum of the three */
        } else dir = (c == T_GAP_2) ? 1 : 2;
        if (dir == 1) c = (ee == H[W * (i - 1) + j] - first) ? T_MATCH : T_GAP_2;
        else if (dir == 2) c = (ff == H[W * i + j - 1] - first) ? T_MATCH : T_GAP_1;
        else c = T_MATCH;
        ops[n++] = (unsigned char)dir;
        if (dir == 0) {
            if (same) { st->matches++; sum += MATCH; } else { st->mismatches++; sum += MISMATCH; }
            i--; j--;
        } else {
            st->gap_ext++;
            if (c == T_MATCH) { st->gap_open++; sum -= first; } else sum -= GAP_EXT;
            if (dir == 1) i--; else j--;
        }
    }
    for (; i > 0; i--) { ops[n++] = 1; st->gap_ext++; c = T_GAP_2; sum -= GAP_EXT; }
    for (; j > 0; j--) { ops[n++] = 2; st->gap_ext++; c = T_GAP_1; sum -= GAP_EXT; }
    if (a.type == T_MATCH && c != T_MATCH) sum -= GAP_OPEN;  /* :309-311: the opening is charged, not counted */
    st->score += sum;
    free(H); free(E); free(F);
    return n;
}

/* stage5() driver loop, :404-424: all partitions in order.  op_len[k] (k = 1..n-1) = steps of partition (k-1, k), written
 * at ops + (pts[k-1].i - pts[0].i) + (pts[k-1].j - pts[0].j).  Returns 0, or -1 on bad input. */
int go_stage5(const unsigned char* seq0, const unsigned char* seq1, const go_xpoint* pts, int n, unsigned char* ops,
              int* op_len, go_s5_stats* st) {
    memset(st, 0, sizeof(*st));
    if (n > 0) op_len[0] = 0;
    for (int k = 1; k < n; k++) {
        long long off = (long long)(pts[k - 1].i - pts[0].i) + (pts[k - 1].j - pts[0].j);
        int r = go_stage5_partition(seq0, seq1, pts[k - 1], pts[k], ops + off, st);
        if (r < 0) return -1;
        op_len[k] = r;
    }
    return 0;
}
