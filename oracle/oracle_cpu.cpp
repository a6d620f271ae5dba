// oracle/oracle_cpu.cpp -- TEST INFRASTRUCTURE (probe 2): THE PARITY ORACLE.
//
// Fills exactly the protected virtuals that the reference's CUDAligner fills (R/src/CUDAligner.hpp:216-232)
// but with host arrays and the reference's scalar block processor (CPUBlockProcessor::processBlock,
// C/libmasa/processors/CPUBlockProcessor.cpp:113-174).  Everything that decides WHICH rows/columns/scores
// leave the aligner and in what order is the reference's own AbstractDiagonalAligner
// (C/libmasa/aligners/AbstractDiagonalAligner.cpp:59-501), linked unmodified from _ref/libmasa.a, so this
// binary produces what reference cudalign would produce, without needing a GPU or the (uncompilable
// under CUDA 12) texture-reference kernels.
//
//   block height  = 4*min(T,width), T=128       (R/src/CUDAligner.cpp:295-297,682-684)
//   grid width    = heuristic of R/src/CUDAligner.cpp:307-347 with 148 SMs, no --blocks override
//   block (bx,by) is evaluated inside processDiagonal(bx+by+1) (R/src/CUDAligner.cu:1272-1273 finishes the
//   long phase of step-1 inside launch `step`), which is what flushSpecialRows/flushLastColumn/pruneBlocks assume.
//
// Environment overrides for invariance tests: ORACLE_T (threads), ORACLE_SMS, ORACLE_B (force grid width).
#include "libmasa/libmasa.hpp"
#include <vector>
#include <algorithm>
#include <cstdlib>

static int envi(const char* k, int d) { const char* v = getenv(k); return v ? atoi(v) : d; }

class DiagCPUAligner : public AbstractDiagonalAligner {
	CPUBlockProcessor proc;
	BlockAlignerParameters* params;
	score_params_t sp;
	long long cells;
	int T, SMS, MAXB, forceB;
	std::vector<cell_t> busH, lastCol, col0next, col0cur;
	std::vector<std::vector<cell_t> > colIn;
	std::vector<score_t> scores;
public:
	DiagCPUAligner() {
		sp.match = 1; sp.mismatch = -3; sp.gap_open = 3; sp.gap_ext = 2;   // R/src/CUDAligner.hpp:77-98
		params = new BlockAlignerParameters();
		cells = 0;
		T = envi("ORACLE_T", 128); SMS = envi("ORACLE_SMS", 148); MAXB = 512; forceB = envi("ORACLE_B", 0);
	}
	aligner_capabilities_t getCapabilities() {           // R/src/CUDAligner.cpp:87-111 minus fork, no length cap
		aligner_capabilities_t c;
		c.smith_waterman = c.needleman_wunsch = c.block_pruning = SUPPORTED;
		c.customize_first_column = c.customize_first_row = SUPPORTED;
		c.dispatch_last_cell = c.dispatch_last_column = c.dispatch_last_row = SUPPORTED;
		c.dispatch_special_row = c.dispatch_block_scores = c.dispatch_scores = SUPPORTED;
		c.process_partition = SUPPORTED;
		return c;
	}
	IAlignerParameters* getParameters() { return params; }
	const score_params_t* getScoreParameters() { return &sp; }
	void initialize() {}
	void finalize() {}
	void unsetSequences() {}
	void setSequences(const char* s0, const char* s1, int l0, int l1) {
		proc.setSequences(s0, s1, l0, l1);
		busH.assign(l1 + 2, cell_t());
	}
	long long getProcessedCells() { return cells; }
protected:
	int getBlockHeight() { int w = getPartition().getWidth(); return (w <= T ? w : T) * 4; }
	int getGridWidth(int width) {
		if (forceB > 0) return std::max(1, std::min(forceB, width));
		int blocks = width / 2 / T, rec = 4 * SMS, maximum = mustPruneBlocks() ? 1000 * rec : rec;
		if (blocks <= 1) return 1;
		if (blocks > MAXB) blocks = MAXB;
		if (blocks <= SMS) {} else if (blocks <= maximum) blocks = (blocks / SMS) * SMS; else blocks = maximum;
		return blocks;
	}
	const cell_t* getSpecialRow(int j, int) { return &busH[j]; }
	const cell_t* getLastRow(int j, int) { return &busH[j]; }
	const cell_t* getLastColumn(int, int) { return &lastCol[1]; }
	const score_t* getBlockScores() { return &scores[0]; }
	void setFirstRow(const cell_t* c, int j, int len) { for (int k = 0; k < len; k++) busH[j + k] = c[k]; }
	void setFirstColumn(const cell_t* c, int, int) { col0next.assign(c, c + getBlockHeight() + 1); }  // [0]=diag, [1..]=(H,E)
	void clearPrunedBlocks(int b0, int b1) {
		int p0, p1;
		getGrid()->getBlockPosition(b0, 0, NULL, &p0, NULL, NULL);
		getGrid()->getBlockPosition(b1, 0, NULL, &p1, NULL, NULL);
		if (p1 < 0) p1 = getPartition().getJ1();
		for (int j = p0; j < p1; j++) { busH[j].h = -INF; busH[j].f = -INF; }
	}
	void initializeDiagonals() {
		int B = getGrid()->getGridWidth(), bh = getBlockHeight();
		colIn.assign(B + 1, std::vector<cell_t>(bh + 1));
		lastCol.assign(bh + 1, cell_t());
		score_t z; z.i = z.j = -1; z.score = -INF;
		scores.assign(B, z);
	}
	void finalizeDiagonals() {}
	void processDiagonal(int d, int wl, int wr) {
		int B = getGrid()->getGridWidth(), bh = getBlockHeight();
		Partition p = getPartition();
		for (int bx = B - 1; bx >= 0; bx--) {       // right-to-left: colIn[bx+1] is consumed before it is overwritten
			scores[bx].score = -INF; scores[bx].i = scores[bx].j = -1;
			int by = d - 1 - bx;
			if (by < 0) continue;
			int i0 = p.getI0() + by * bh;
			if (i0 >= p.getI1()) continue;
			int i1 = std::min(i0 + bh, p.getI1());
			int j0, j1;
			getGrid()->getBlockPosition(bx, 0, NULL, &j0, NULL, &j1);
			if (bx == 0 && getFirstColumnInitType() != INIT_WITH_ZEROES) colIn[0] = col0cur;
			std::vector<cell_t>& col = colIn[bx];
			if (bx == 0 && getFirstColumnInitType() == INIT_WITH_ZEROES)
				for (int k = 0; k <= bh; k++) { col[k].h = 0; col[k].e = -INF; }
			if (bx < wl || bx > wr) {                 // pruned: -INF to the right, busH left stale (CUDAligner.cu:950-960)
				for (int k = 0; k <= bh; k++) { col[k].h = -INF; col[k].e = -INF; }
			} else {
				scores[bx] = proc.processBlock(&busH[j0], &col[0], i0, j0, i1, j1, getRecurrenceType());
				cells += (long long)(i1 - i0) * (j1 - j0);
			}
			if (bx == B - 1) lastCol = col; else std::swap(colIn[bx], colIn[bx + 1]);
		}
		col0cur = col0next;
	}
};

int main(int argc, char** argv) {
	return libmasa_entry_point(argc, argv, new DiagCPUAligner(), (char*)"oracle-cpu (reference CPUBlockProcessor, Diagonal policy)");
}
