// stage4_gpu.cpp -- GPU stage 4 for the cudalign binary.
//
// MASA-Core has no plugin hook for stage 4: stage4(Job*, int) (C/stage4/sw_stage4.cpp:883-975) is a free function
// of libmasa.a that runs the Myers-Miller split on 4 CPU threads.  build/cudalign substitutes it AT LINK TIME:
// this object defines stage4() itself and is placed before libmasa.a (-Wl,--allow-multiple-definition keeps the
// archive's other symbol of that object, stage4_pool_wait).  MASA-Core's sources stay untouched; a maintainer who
// prefers an explicit hook replaces the body of reduce_partitions() (:806-852) by the same b200_stage4 call.
//
// Same inputs and outputs as the reference driver: reads crosspoint_03.NN, writes crosspoint_04.NN and
// statistics_04.NN; only the default --stage-4-strategy (OPTIMIZED, ort_split_2) is implemented on the GPU.
// Set B200_STAGE4=off in the environment to refuse (for A/B timing use the oracle binary).
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "libmasa/libmasa.hpp"
#include "common/Job.hpp"
#include "common/CrosspointsFile.hpp"
#include "common/Timer.hpp"
#include "B200Aligner.hpp"

static_assert(sizeof(crosspoint_t) == sizeof(b200_xpoint), "crosspoint_t layout (C/common/Crosspoint.hpp:30-40)");

void stage4(Job* job, int id) {
	FILE* stats = job->fopenStatistics(STAGE_4, id);
	Sequence* seq0 = job->getAlignmentParams()->getSequence(0);
	Sequence* seq1 = job->getAlignmentParams()->getSequence(1);
	job->getAlignmentParams()->printParams(stats);
	fprintf(stats, "MAXIMUM PARTITION SIZE: %d\n", job->stage4_maximum_partition_size);
	fprintf(stats, "STAGE4 STRATEGY: #%d (B200 batched GPU split)\n", job->stage4_strategy);
	if (job->stage4_strategy != STAGE_4_STRATEGY_OPTIMIZED) {
		fprintf(stderr, "cudalign-b200: only --stage-4-strategy=%d (optimized, the default) runs on the GPU.\n", STAGE_4_STRATEGY_OPTIMIZED);
		exit(1);
	}
	b200_handle* h = B200Aligner::activeHandle();
	if (h == NULL) {
		fprintf(stderr, "cudalign-b200: stage 4 needs an initialised GPU aligner.\n");
		exit(1);
	}
	Timer timer;
	int ev_start = timer.createEvent("START");
	int ev_split = timer.createEvent("GPU_SPLIT");
	int ev_write = timer.createEvent("WRITE");
	timer.eventRecord(ev_start);

	CrosspointsFile* stage3 = new CrosspointsFile(job->getCrosspointFile(STAGE_3, id));
	stage3->loadCrosspoints();
	std::vector<b200_xpoint> in(stage3->size());
	for (size_t k = 0; k < stage3->size(); k++) {
		const crosspoint_t& c = stage3->at(k);
		in[k].i = c.i; in[k].j = c.j; in[k].type = c.type; in[k].score = c.score;
	}
	delete stage3;

	int max_i = 0, max_j = 0;
	if (!in.empty()) { max_i = in.back().i - in.front().i; max_j = in.back().j - in.front().j; }
	const int cap = 4 * (max_i + max_j) + 4 * (int)in.size() + 64;
	std::vector<b200_xpoint> out(cap);
	int n_out = 0;
	if (b200_set_sequences(h, seq0->getData(false), seq0->getInfo()->getSize(), seq1->getData(false), seq1->getInfo()->getSize()) != 0 ||
	    b200_stage4(h, in.data(), (int)in.size(), job->stage4_maximum_partition_size, out.data(), cap, &n_out) != 0) {
		fprintf(stderr, "cudalign-b200: stage 4 failed: %s\n", b200_last_error(h));
		exit(1);
	}
	float t_split = timer.eventRecord(ev_split);

	CrosspointsFile* crosspoints = new CrosspointsFile(job->getCrosspointFile(STAGE_4, id));
	crosspoints->clear();
	for (int k = 0; k < n_out; k++) {
		crosspoint_t c;
		c.i = out[k].i; c.j = out[k].j; c.type = out[k].type; c.score = out[k].score;
		crosspoints->push_back(c);
	}
	crosspoints->save();
	fprintf(stats, " crosspoints: %8d -> %8d   gpu time: %.4f\n", (int)in.size(), n_out, t_split);
	delete crosspoints;
	timer.eventRecord(ev_write);
	fprintf(stats, "Stage4 times:\n");
	float diff = timer.printStatistics(stats);
	fprintf(stats, "        Total: %.4f\n", diff);
	fclose(stats);
}
