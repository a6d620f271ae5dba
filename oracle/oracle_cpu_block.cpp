// oracle/oracle_cpu_block.cpp -- TEST INFRASTRUCTURE (probe 1).
//
// A concrete AbstractBlockAligner so that the reference's own CPU path (CPUBlockProcessor::processBlock,
// C/libmasa/processors/CPUBlockProcessor.cpp:113-174, driven by C/libmasa/aligners/AbstractBlockAligner.cpp)
// becomes a runnable `cudalign`-compatible binary.  The Block family is the one that supports --fork
// (AbstractBlockAligner.cpp:104-110) so this binary is the CPU timing baseline (bench.py --impl reference)
// and an independent cross-check of scores/coordinates.  It is NOT used for crosspoint/transcript parity
// (its special-row ids follow the Block policy; see SURVEY.md 8c caveat).
#include "libmasa/libmasa.hpp"

class SerialBlockAligner : public AbstractBlockAligner {
public:
	SerialBlockAligner() : AbstractBlockAligner(NULL, NULL) {}
protected:
	// row-major schedule: rows and columns leave the aligner in order
	void scheduleBlocks(int grid_width, int grid_height) {
		for (int by = 0; by < grid_height; by++)
			for (int bx = 0; bx < grid_width; bx++)
				AbstractBlockAligner::alignBlock(bx, by);
	}
	void alignBlock(int bx, int by, int i0, int j0, int i1, int j1) {
		if (by == 0) {
			receiveFirstRow(row[bx], j1 - j0);
			if (isSpecialColumn(bx)) {
				cell_t c = getFirstRowTail(); c.f = -INF;
				dispatchColumn(j1, &c, 1);
			}
		}
		if (bx == 0) {
			col[by][0] = getFirstColumnTail();
			receiveFirstColumn(col[by] + 1, i1 - i0);
		}
		processBlock(bx, by, i0, j0, i1, j1);
		if (isSpecialRow(by)) {
			if (bx == 0) {
				cell_t c = getFirstColumnTail(); c.f = -INF;
				dispatchRow(i1, &c, 1);
			}
			dispatchRow(i1, row[bx], j1 - j0);
		}
		if (isSpecialColumn(bx)) dispatchColumn(j1, col[by] + 1, i1 - i0);
	}
};

int main(int argc, char** argv) {
	return libmasa_entry_point(argc, argv, new SerialBlockAligner(), (char*)"oracle-cpu-block (reference CPUBlockProcessor, Block policy)");
}
