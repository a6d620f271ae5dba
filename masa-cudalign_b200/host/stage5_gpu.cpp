// stage5_gpu.cpp -- GPU stage 5 for the cudalign binary.
//
// Like stage 4, MASA-Core offers no plugin hook for stage 5: stage5(Job*, int) (C/stage5/sw_stage5.cpp:322-498) is a
// free function of libmasa.a that aligns and walks back every partition of crosspoint_04.NN on one CPU thread.
// build/cudalign substitutes it AT LINK TIME (this object precedes libmasa.a, see the Makefile); MASA-Core's sources
// stay untouched.  A maintainer who prefers an explicit hook replaces the partition loop (:404-424) by the
// b200_stage5 call + replay below.
//
// Same inputs and outputs as the reference driver: reads crosspoint_04.NN, fills an Alignment through the same
// addGapInSeq0/addGapInSeq1 calls in the same order (dot(), :70-84), checks the score against the crosspoints
// (:452-456), writes alignment.NN.bin and statistics_05.NN.  The tracebacks themselves run on the GPU, one thread
// per partition (csrc/stage5.cuh).
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "libmasa/libmasa.hpp"
#include "common/Common.hpp"
#include "B200Aligner.hpp"

static_assert(sizeof(crosspoint_t) == sizeof(b200_xpoint), "crosspoint_t layout (C/common/Crosspoint.hpp:30-40)");

int stage5(Job* job, int id) {
	FILE* stats = job->fopenStatistics(STAGE_5, id);
	AlignmentParams* ap = job->getAlignmentParams();
	Sequence* seq0 = ap->getSequence(0);
	Sequence* seq1 = ap->getSequence(1);
	ap->printParams(stats);
	fflush(stats);
	// the device code has the aligner's compile-time scores (variable_penalties = NOT_SUPPORTED, like the reference GPU aligner)
	// (AlignmentParams keeps the penalties negated: setAffineGapPenalties(-gap_open, -gap_ext), libmasa.cpp:775; :329-330)
	if (ap->getMatch() != 1 || ap->getMismatch() != -3 || -ap->getGapOpen() != 3 || -ap->getGapExtension() != 2) {
		fprintf(stderr, "cudalign-b200: stage 5 on the GPU supports the scores +1/-3/3/2 only.\n");
		exit(1);
	}
	b200_handle* h = B200Aligner::activeHandle();
	if (h == NULL) {
		fprintf(stderr, "cudalign-b200: stage 5 needs an initialised GPU aligner.\n");
		exit(1);
	}

	Timer timer;
	int ev_load = timer.createEvent("LOAD");
	int ev_gpu = timer.createEvent("GPU_TRACEBACK");
	int ev_replay = timer.createEvent("REPLAY");
	int ev_finalize = timer.createEvent("FINALIZE");
	int ev_write = timer.createEvent("WRITE_BINARY");
	timer.init();

	CrosspointsFile* xp = new CrosspointsFile(job->getCrosspointFile(STAGE_4, id));
	xp->loadCrosspoints();
	const int n = (int)xp->size();
	int largest = xp->getLargestPartitionSize();
	if (largest > 1024) {                       // the reference's table limit (H_MAX / W_MAX, :31-32,396-399): same refusal
		fprintf(stderr, "ERROR: MAX SIZE: %d\n", largest);
		exit(1);
	}
	fprintf(stats, "Largest Block: %d\n", largest);
	std::vector<b200_xpoint> pts(n);
	for (int k = 0; k < n; k++) {
		const crosspoint_t& c = xp->at(k);
		pts[k].i = c.i; pts[k].j = c.j; pts[k].type = c.type; pts[k].score = c.score;
	}
	timer.eventRecord(ev_load);

	const long long cap = n > 0 ? ((long long)pts[n - 1].i - pts[0].i) + ((long long)pts[n - 1].j - pts[0].j) : 0;
	std::vector<unsigned char> ops((size_t)cap + 1);
	std::vector<int> op_len(n > 0 ? n : 1, 0);
	b200_s5_stats total = {0, 0, 0, 0, 0};
	if (n > 1) {
		if (b200_set_sequences(h, seq0->getData(false), seq0->getInfo()->getSize(), seq1->getData(false), seq1->getInfo()->getSize()) != 0 ||
		    b200_stage5(h, pts.data(), n, ops.data(), cap, op_len.data(), &total) != 0) {
			fprintf(stderr, "cudalign-b200: stage 5 failed: %s\n", b200_last_error(h));
			exit(1);
		}
	}
	timer.eventRecord(ev_gpu);

	// replay: partitions top-down, each one from its bottom-right corner, exactly the reference's order of calls
	Alignment* alignment = new Alignment(ap);
	const int adjust0 = seq0->isReversed() ? 0 : 1;
	const int adjust1 = seq1->isReversed() ? 0 : 1;
	for (int k = 1; k < n; k++) {
		int i = pts[k].i - pts[k - 1].i, j = pts[k].j - pts[k - 1].j;
		const int i0 = pts[k - 1].i, j0 = pts[k - 1].j;
		const unsigned char* op = ops.data() + (((long long)i0 - pts[0].i) + ((long long)j0 - pts[0].j));
		for (int q = 0; q < op_len[k]; q++) {
			switch (op[q]) {
			case 0: i--; j--; break;
			case 1: alignment->addGapInSeq1(seq1->getAbsolutePos(j0 + j + adjust1)); i--; break;
			case 2: alignment->addGapInSeq0(seq0->getAbsolutePos(i0 + i + adjust0)); j--; break;
			default:
				fprintf(stderr, "cudalign-b200: stage 5: corrupt traceback of partition %d\n", k);
				exit(1);
			}
		}
		if (i != 0 || j != 0) {
			fprintf(stderr, "cudalign-b200: stage 5: the traceback of partition %d stopped at (%d,%d)\n", k, i, j);
			exit(1);
		}
	}
	timer.eventRecord(ev_replay);

	crosspoint_t start = xp->front();
	crosspoint_t end = xp->back();
	fprintf(stats, "(%d,%d)\n", start.type, end.type);
	fprintf(stats, "(%d,%d)->(%d,%d)\n", start.i, start.j, end.i, end.j);
	if (n != 1) {
		alignment->setStart(0, seq0->getAbsolutePos(start.i + 1));
		alignment->setStart(1, seq1->getAbsolutePos(start.j + 1));
		alignment->setEnd(0, seq0->getAbsolutePos(end.i));
		alignment->setEnd(1, seq1->getAbsolutePos(end.j));
	} else {
		for (int s = 0; s < 2; s++) { alignment->setStart(s, -1); alignment->setEnd(s, -1); }
	}
	const int expected = end.score - start.score;
	if (expected != total.score) {
		fprintf(stderr, "stage5: Wrong Alignment Score: %d != %d*.\n", total.score, expected);
		exit(1);
	}
	alignment->setRawScore(total.score);
	alignment->setMatches(total.matches);
	alignment->setMismatches(total.mismatches);
	alignment->setGapOpen(total.gap_open);
	alignment->setGapExtensions(total.gap_ext);
	if (job->dump_blocks) alignment->setPruningFile(job->dump_pruning_text_filename.c_str());
	alignment->finalize();
	job->setAlignment(alignment);
	timer.eventRecord(ev_finalize);

	AlignmentBinaryFile::write(job->getAlignmentBinaryFile(id), alignment);
	timer.eventRecord(ev_write);

	fprintf(stats, "Goal Diff: %d\n", expected);
	fprintf(stats, "Partitions: %d (B200 batched GPU traceback)\n", n > 0 ? n - 1 : 0);
	fprintf(stats, "Stage5 times:\n");
	float diff = timer.printStatistics(stats);
	fprintf(stats, "        total: %.4f\n", diff);
	fclose(stats);
	delete xp;
	return 0;
}
