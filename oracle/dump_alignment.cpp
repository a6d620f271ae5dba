// oracle/dump_alignment.cpp -- TEST INFRASTRUCTURE ONLY (built into oracle/_ref/, never linked by the product).
//
// Prints what the reference's stage 5 stored in an alignment.NN.bin, read back with the reference's OWN reader
// (C/common/biology/AlignmentBinaryFile.cpp:86-102, linked from oracle/_ref/libmasa.a), as one JSON object:
// raw score, the four counters of total_score_t (C/stage5/sw_stage5.cpp:51-67), start/end and the two gap lists
// [[pos, len], ...] in file order (sorted by position, Alignment.cpp finalize()).  tests/golden/make_stage5_golden.py
// uses it to pin oracle/gotoh_oracle.c's stage-5 restatement to the reference.
#include <stdio.h>
#include <string>
#include <vector>
#include "common/biology/biology.hpp"

int main(int argc, char** argv) {
	if (argc != 2) { fprintf(stderr, "usage: dump_alignment alignment.NN.bin\n"); return 2; }
	Alignment* al = AlignmentBinaryFile::read(std::string(argv[1]));
	if (al == NULL) return 1;
	printf("{\"raw_score\": %d, \"matches\": %d, \"mismatches\": %d, \"gap_open\": %d, \"gap_ext\": %d,\n",
	       al->getRawScore(), al->getMatches(), al->getMismatches(), al->getGapOpen(), al->getGapExtensions());
	printf(" \"start\": [%d, %d], \"end\": [%d, %d],\n", al->getStart(0), al->getStart(1), al->getEnd(0), al->getEnd(1));
	for (int s = 0; s < 2; s++) {
		std::vector<gap_t>* g = al->getGaps(s);
		printf(" \"gaps%d\": [", s);
		for (size_t k = 0; k < g->size(); k++) printf("%s[%d, %d]", k ? ", " : "", (*g)[k].pos, (*g)[k].len);
		printf("]%s\n", s == 0 ? "," : "}");
	}
	return 0;
}
