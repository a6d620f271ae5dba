#include "B200Aligner.hpp"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stddef.h>
#include <algorithm>

static_assert(sizeof(cell_t) == sizeof(b200_cell), "cell_t layout (C/libmasa/libmasaTypes.hpp:35-41)");
static_assert(sizeof(score_t) == sizeof(b200_score), "score_t layout (C/libmasa/libmasaTypes.hpp:88-95)");
static_assert(offsetof(score_t, i) == offsetof(b200_score, i) && offsetof(score_t, j) == offsetof(b200_score, j) &&
              offsetof(score_t, score) == offsetof(b200_score, score), "score_t field order");
static_assert(offsetof(cell_t, h) == offsetof(b200_cell, h) && offsetof(cell_t, f) == offsetof(b200_cell, x), "cell_t field order");

static b200_handle* g_active_handle = NULL;
b200_handle* B200Aligner::activeHandle() { return g_active_handle; }

B200Aligner::B200Aligner() {
	params = new B200AlignerParameters();
	score_params.match = 1;          /* R/src/CUDAligner.hpp:77-98 */
	score_params.mismatch = -3;
	score_params.gap_open = 3;
	score_params.gap_ext = 2;
	handle = NULL;
	group = NULL; groupRows = groupJobs = 0; groupSeqValid = false; seq0_ptr = seq1_ptr = NULL; groupPartitions = 0;
	multiprocessors = 148;
	seq0_len = seq1_len = 0;
	fastActive = false;
	fastCells = 0;
	fastDeviceMs = 0;
	fastPartitions = diagPartitions = chunkPartitions = chunkLaunches = 0;
	bufferingTail = false; tailFirstCellSeen = false;
}

B200Aligner::~B200Aligner() {}

/* print-and-exit, like cutilSafeCall (R/src/cuda_util.h:34-61) */
void B200Aligner::check(int rc, const char* what) {
	if (rc != 0) {
		fprintf(stderr, "B200Aligner: %s failed: %s\n", what, b200_last_error(handle));
		exit(-1);
	}
}

aligner_capabilities_t B200Aligner::getCapabilities() {
	aligner_capabilities_t c;                       /* R/src/CUDAligner.cpp:87-111 */
	c.smith_waterman = SUPPORTED;
	c.needleman_wunsch = SUPPORTED;
	c.block_pruning = SUPPORTED;
	c.customize_first_column = SUPPORTED;
	c.customize_first_row = SUPPORTED;
	c.dispatch_last_cell = SUPPORTED;
	c.dispatch_last_column = SUPPORTED;
	c.dispatch_last_row = SUPPORTED;
	c.dispatch_special_column = NOT_SUPPORTED;
	c.dispatch_special_row = SUPPORTED;
	c.dispatch_block_scores = SUPPORTED;
	c.dispatch_scores = SUPPORTED;
	c.process_partition = SUPPORTED;
	c.variable_penalties = NOT_SUPPORTED;
	c.fork_processes = NOT_SUPPORTED;               /* multi-GPU = --gpus=N: in-kernel NVLink chain (b200_group_*), not fork()+sockets */
	c.maximum_seq0_len = 0;                         /* no 2^27 texture limit (R/src/CUDAligner.cpp:41-44) */
	c.maximum_seq1_len = 0;
	return c;
}

IAlignerParameters* B200Aligner::getParameters() { return params; }
const score_params_t* B200Aligner::getScoreParameters() { return &score_params; }

void B200Aligner::initialize() {
	if (params->getGpuList().size() > 1) return;   /* --gpus: the group is created by the first setSequences (it is sized by the sequences) */
	b200_config cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.device = params->getGpuList().size() == 1 ? params->getGpuList()[0] : (params->getGPU() < 0 ? 0 : params->getGPU());
	cfg.kernel = params->getKernel();
	int rc = b200_create(&cfg, &handle);
	if (rc != 0) {
		fprintf(stderr, "B200Aligner: cannot initialise the GPU: %s\n", b200_last_error(NULL));
		exit(-1);
	}
	g_active_handle = handle;
}

void B200Aligner::finalize() {
	if (group != NULL) {
		if (g_active_handle == handle) g_active_handle = NULL;
		b200_group_destroy(group);          /* owns every handle, including `handle` */
		group = NULL; handle = NULL;
	}
	if (handle != NULL) {
		if (g_active_handle == handle) g_active_handle = NULL;
		b200_destroy(handle);
		handle = NULL;
	}
}

/* --gpus=N: (re)create the group when the sequences need a larger exchange block than the current one has */
void B200Aligner::ensureGroup() {
	b200_partition whole;
	memset(&whole, 0, sizeof(whole));
	/* MASA-Core hands over empty sequences for degenerate (pure-gap) stages of semi-global alignments: plan for one cell */
	whole.i1 = seq0_len > 0 ? seq0_len : 1; whole.j1 = seq1_len > 0 ? seq1_len : 1;
	if (const char* e = getenv("B200_CHAIN_CHUNK")) whole.reserved[1] = atoi(e);   /* the chunk width override of alignPartitionFast (tests): size the exchange blocks for it */
	b200_chain_info info;
	const std::vector<int>& devs = params->getGpuList();
	if (b200_chain_plan(&whole, (int)devs.size(), &info) != 0) { fprintf(stderr, "B200Aligner: b200_chain_plan failed\n"); exit(-1); }
	if (group != NULL && seq0_len <= groupRows && info.max_jobs <= groupJobs) return;
	if (group != NULL) { if (g_active_handle == handle) g_active_handle = NULL; b200_group_destroy(group); group = NULL; handle = NULL; }
	b200_config cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.kernel = params->getKernel();
	groupRows = seq0_len; groupJobs = info.max_jobs;
	int rc = b200_group_create(&devs[0], (int)devs.size(), &cfg, groupRows, groupJobs, &group);
	if (rc != 0) {
		fprintf(stderr, "B200Aligner: cannot initialise the GPU group: %s\n", b200_group_last_error(NULL));
		exit(-1);
	}
	handle = b200_group_handle(group, 0);
	g_active_handle = handle;
}

void B200Aligner::setSequences(const char* seq0, const char* seq1, int seq0_len, int seq1_len) {
	this->seq0_len = seq0_len;
	this->seq1_len = seq1_len;
	seq0_ptr = seq0; seq1_ptr = seq1;
	if (params->getGpuList().size() > 1) ensureGroup();
	/* the other GPUs of a group receive the sequences only when a partition is large enough to use them (useGroupFor) */
	groupSeqValid = false;
	check(b200_set_sequences(handle, seq0, seq0_len, seq1, seq1_len), "b200_set_sequences");
	rowBuffer.resize((size_t)seq1_len + 2);
}

void B200Aligner::unsetSequences() {
	check(b200_unset_sequences(handle), "b200_unset_sequences");
}

match_result_t B200Aligner::matchLastColumn(const cell_t* buffer, const cell_t* base, int len, int goalScore) {
	/* The chunks are <= 1024 cells (C/common/AlignerManager.cpp:643-652): a device round trip would cost more
	 * than the scan, so the host matcher of AbstractAligner is used; b200_match_last_column is the device
	 * variant (parity: tests/test_match_gpu.py). */
	return AbstractAligner::matchLastColumn(buffer, base, len, goalScore);
}

/* ------------------------------------------------------------------------------------------------------------
 * B200-first path for stage 1
 * ---------------------------------------------------------------------------------------------------------- */
/* --dump-blocks (C/libmasa/libmasa.cpp:1082, C/stage1/sw_stage1.cpp:310-314) makes the manager file the score of EVERY block of
 * the reference's grid (AlignerManager.cpp:418-422) -- a per-block artefact only the per-diagonal path produces.  The flag is
 * MASA-Core's own (the aligner is never told: IManager has no query for it), so the adapter looks at the command line. */
static bool dumpBlocksRequested() {
	static int cached = -1;
	if (cached < 0) {
		cached = 0;
		FILE* f = fopen("/proc/self/cmdline", "rb");
		if (f != NULL) {
			std::string all;
			char buf[4096];
			size_t got;
			while ((got = fread(buf, 1, sizeof(buf), f)) > 0) all.append(buf, got);
			fclose(f);
			for (size_t pos = 0; pos < all.size(); pos += strlen(all.c_str() + pos) + 1)
				if (strcmp(all.c_str() + pos, "--dump-blocks") == 0) cached = 1;
		}
	}
	return cached == 1;
}

bool B200Aligner::canUseFastPath() {
	if (!params->useFastPath()) return false;
	if (dumpBlocksRequested()) return false;          /* per-block scores: the reference's per-diagonal contract */
	if (mustDispatchLastColumn()) return false;      /* stages 2/3 (goal matching, early stop), split partitions */
	if (mustDispatchSpecialColumns()) return false;
	return true;
}

/* Stage 2/3 partitions of the common kind: the goal is matched on the LAST COLUMN (AT_SEQUENCE_1_OR_2,
 * C/stage2/sw_stage2.cpp:80-88), nothing depends on per-block scores.  The last column is a sequential stream
 * (first hit wins, C/common/AlignerManager.cpp:334-369), so the partition can be aligned in chunks of rows, each one
 * persistent launch, dispatching the column chunks in order and stopping after the chunk in which the manager found
 * its crosspoint: identical crosspoints, ~100x fewer launches than one per external diagonal. */
bool B200Aligner::canUseChunkPath() {
	if (!params->useFastPath()) return false;
	if (!mustDispatchLastColumn()) return false;
	if (mustDispatchScores() || mustDispatchLastCell() || mustPruneBlocks()) return false;
	if (mustDispatchSpecialColumns()) return false;
	return true;
}

void B200Aligner::alignPartition(Partition partition) {
	static const bool dbg = getenv("B200_DEBUG") != NULL;
	if (dbg) fprintf(stderr, "[B200Aligner] partition %dx%d lastcol=%d scores=%d lastrow=%d lastcell=%d prune=%d srows=%d rec=%d\n",
			partition.getHeight(), partition.getWidth(), (int)mustDispatchLastColumn(), (int)mustDispatchScores(), (int)mustDispatchLastRow(),
			(int)mustDispatchLastCell(), (int)mustPruneBlocks(), (int)mustDispatchSpecialRows(), getRecurrenceType());
	if (canUseFastPath()) {
		alignPartitionFast(partition);
	} else if (canUseChunkPath()) {
		alignPartitionChunked(partition);
	} else {
		diagPartitions++;
		AbstractDiagonalAligner::alignPartition(partition);
	}
}

void B200Aligner::fillPartition(b200_partition& p, Partition partition) {
	memset(&p, 0, sizeof(p));
	p.i0 = partition.getI0(); p.j0 = partition.getJ0(); p.i1 = partition.getI1(); p.j1 = partition.getJ1();
	p.recurrence = getRecurrenceType();
	p.first_row_init = getFirstRowInitType();
	p.first_col_init = getFirstColumnInitType();
	p.special_row_interval = getSpecialRowInterval();
	int width = partition.getWidth();
	p.block_height = (width <= B200_THREADS_COUNT ? width : B200_THREADS_COUNT) * B200_ALPHA;
	p.want_special_rows = mustDispatchSpecialRows() && getSpecialRowInterval() > 0;
	p.want_last_row = mustDispatchLastRow() || mustDispatchLastCell();
	p.want_last_column = mustDispatchLastColumn();
	p.want_best_score = mustDispatchScores();
	p.prune = mustPruneBlocks();
	Partition sp = getSuperPartition();
	p.super_i1 = sp.getI1(); p.super_j1 = sp.getJ1();
}

void B200Aligner::alignPartitionChunked(Partition partition) {
	chunkPartitions++;
	fastPartition = partition;
	fastActive = true;
	b200_callbacks cb;
	memset(&cb, 0, sizeof(cb));
	cb.ctx = this;
	cb.receive_first_row = cbReceiveFirstRow;
	cb.receive_first_column = cbReceiveFirstColumn;
	cb.dispatch_row = cbDispatchRow;
	cb.dispatch_column = cbDispatchColumn;
	cb.dispatch_score = cbDispatchScore;
	cb.must_continue = cbMustContinue;
	const int height = partition.getHeight(), width = partition.getWidth();
	const int bh = (width <= B200_THREADS_COUNT ? width : B200_THREADS_COUNT) * B200_ALPHA;
	const int B = getGridWidth(width);
	const bool wantLastRow = mustDispatchLastRow();
	/* The path the manager is looking for is roughly diagonal: it leaves a partition of this width after about
	 * `width` rows.  Chunks are whole blocks, at least B block-rows (2*width rows) so that, in the reference's
	 * external-diagonal order, every last-column chunk of an earlier launch precedes the first last-row piece
	 * (column chunk `by` leaves at diagonal by+B, row piece p at diagonal gridHeight+p). */
	long long target = 2LL * width;
	if (target < 8192) target = 8192;
	if (target > 131072) target = 131072;
	int chunk = (int)(((target + bh - 1) / bh) * bh);
	if (chunk < B * bh) chunk = B * bh;
	for (int off = 0; off < height && mustContinue();) {
		int rows = chunk;
		if (height - (off + rows) < B * bh) rows = height - off;      /* never leave a short tail chunk */
		if (off + rows > height) rows = height - off;
		const bool finalChunk = off + rows >= height;
		b200_partition p;
		fillPartition(p, partition);
		p.i0 = partition.getI0() + off;
		p.i1 = p.i0 + rows;
		p.want_last_row = (finalChunk && wantLastRow) ? 1 : 0;
		p.want_best_score = 0;
		p.prune = 0;
		p.reserved[0] = off > 0 ? B200_CONT_CHUNK : 0;
		p.reserved[2] = off;
		p.reserved[3] = height;
		bufferingTail = finalChunk && wantLastRow;
		tailFirstCellSeen = off > 0;                /* the corner cell of the last column is dispatched by the first chunk only */
		tailCol.clear(); tailRow.clear();
		b200_result res;
		check(b200_align_partition(handle, &p, &cb, &res), "b200_align_partition (chunk)");
		fastCells += res.cells;
		fastDeviceMs += res.device_ms;
		chunkLaunches++;
		if (bufferingTail) {
			bufferingTail = false;
			/* replay in the order of AbstractDiagonalAligner::processNextIteration (flushLastRow before
			 * flushLastColumn inside one external diagonal, AbstractDiagonalAligner.cpp:120-130,325-369) */
			Grid* grid = createGrid(partition);
			grid->setBlockHeight(bh);
			grid->splitGridHorizontally(B);
			const int gh = height / bh + 1;
			const int by0 = off / bh;
			for (int d = 0; d <= gh + B && mustContinue(); d++) {
				const int piece = d - gh;
				if (piece >= 0 && piece < B && !tailRow.empty()) {
					int x0, x1;
					grid->getBlockPosition(piece, 0, NULL, &x0, NULL, &x1);
					if (piece == 0) dispatchRow(partition.getI1(), &tailRowFirst, 1);
					if (mustContinue()) dispatchRow(partition.getI1(), &tailRow[x0 - partition.getJ0()], x1 - x0);
				}
				const int by = d - B;
				if (mustContinue() && by >= by0 && (long long)by * bh < height) {
					const int r0 = by * bh - off;
					const int len = (by * bh + bh <= height) ? bh : height - by * bh;
					if (r0 >= 0 && r0 + len <= (int)tailCol.size()) dispatchColumn(partition.getJ1(), &tailCol[r0], len);
				}
			}
		}
		off += rows;
	}
	fastActive = false;
}

/* Partitions worth the chain: stage 1 of a big comparison.  Small ones (stage 3, short sequences) would spend longer
 * filling the multi-GPU pipeline than computing. */
bool B200Aligner::useGroupFor(Partition partition) {
	if (group == NULL) return false;
	long long min_cells = 20000000000LL;
	if (const char* e = getenv("B200_GROUP_MIN_CELLS")) min_cells = atoll(e);      /* tests: small partitions through the chain */
	return (long long)partition.getHeight() * (long long)partition.getWidth() >= min_cells;
}

void B200Aligner::alignPartitionFast(Partition partition) {
	fastPartitions++;
	fastPartition = partition;
	fastActive = true;
	b200_partition p;
	fillPartition(p, partition);
	p.want_last_column = 0;
	const bool chain = useGroupFor(partition);
	if (chain) { if (const char* e = getenv("B200_CHAIN_CHUNK")) p.reserved[1] = atoi(e); }     /* chunk width override (tests) */
	if (chain && !groupSeqValid) {
		for (int r = 1; r < b200_group_size(group); r++)
			check(b200_set_sequences(b200_group_handle(group, r), seq0_ptr, seq0_len, seq1_ptr, seq1_len), "b200_set_sequences (group)");
		groupSeqValid = true;
	}

	b200_callbacks cb;
	memset(&cb, 0, sizeof(cb));
	cb.ctx = this;
	cb.receive_first_row = cbReceiveFirstRow;
	cb.receive_first_column = cbReceiveFirstColumn;
	cb.dispatch_row = cbDispatchRow;
	cb.dispatch_column = cbDispatchColumn;
	cb.dispatch_score = cbDispatchScore;
	cb.must_continue = cbMustContinue;
	b200_result res;
	if (chain) {
		groupPartitions++;
		if (b200_group_align_partition(group, &p, &cb, &res) != 0) {
			fprintf(stderr, "B200Aligner: b200_group_align_partition failed: %s\n", b200_group_last_error(group));
			exit(-1);
		}
	} else {
		check(b200_align_partition(handle, &p, &cb, &res), "b200_align_partition");
	}
	fastCells += res.cells;
	fastDeviceMs += res.device_ms;
	fastActive = false;
}

void B200Aligner::cbReceiveFirstRow(void* ctx, b200_cell* buffer, int len) {
	((B200Aligner*)ctx)->receiveFirstRow((cell_t*)buffer, len);
}
void B200Aligner::cbReceiveFirstColumn(void* ctx, b200_cell* buffer, int len) {
	((B200Aligner*)ctx)->receiveFirstColumn((cell_t*)buffer, len);
}
void B200Aligner::cbDispatchRow(void* ctx, int i, const b200_cell* buffer, int len) {
	B200Aligner* a = (B200Aligner*)ctx;
	const bool last = (i == a->fastPartition.getI1());
	if (last && a->bufferingTail) {
		if (len == 1 && a->tailRow.empty()) { a->tailRowFirst = *(const cell_t*)buffer; return; }
		a->tailRow.insert(a->tailRow.end(), (const cell_t*)buffer, (const cell_t*)buffer + len);
		return;
	}
	if (last && !a->mustDispatchLastRow()) {
		/* only the last CELL was asked for (bestScoreLocation == AT_SEQUENCE_1_AND_2): AbstractDiagonalAligner::flushLastCell */
		if (len > 1 && a->mustDispatchLastCell()) {
			score_t s;
			s.i = a->fastPartition.getI1() - 1;
			s.j = a->fastPartition.getJ1() - 1;
			s.score = buffer[len - 1].h;
			a->dispatchScore(s);
		}
		return;
	}
	a->dispatchRow(i, (const cell_t*)buffer, len);
	if (last && len > 1 && a->mustDispatchLastCell()) {
		score_t s;
		s.i = a->fastPartition.getI1() - 1;
		s.j = a->fastPartition.getJ1() - 1;
		s.score = buffer[len - 1].h;
		a->dispatchScore(s);
	}
}
void B200Aligner::cbDispatchColumn(void* ctx, int j, const b200_cell* buffer, int len) {
	B200Aligner* a = (B200Aligner*)ctx;
	if (a->bufferingTail) {
		if (!a->tailFirstCellSeen) { a->tailFirstCellSeen = true; a->dispatchColumn(j, (const cell_t*)buffer, len); return; }   /* corner cell */
		a->tailCol.insert(a->tailCol.end(), (const cell_t*)buffer, (const cell_t*)buffer + len);
		return;
	}
	a->dispatchColumn(j, (const cell_t*)buffer, len);
}

void B200Aligner::cbDispatchScore(void* ctx, b200_score score) {
	score_t s;
	s.i = score.i; s.j = score.j; s.score = score.score;
	((B200Aligner*)ctx)->dispatchScore(s);
}
int B200Aligner::cbMustContinue(void* ctx) {
	B200Aligner* a = (B200Aligner*)ctx;
	return (a->bufferingTail || a->mustContinue()) ? 1 : 0;      /* while buffering, nothing has reached the manager yet */
}

/* ------------------------------------------------------------------------------------------------------------
 * Compatibility path: one C-ABI call per CUDAligner virtual
 * ---------------------------------------------------------------------------------------------------------- */
int B200Aligner::getBlockHeight() {                  /* R/src/CUDAligner.cpp:295-297,682-684 */
	int w = getPartition().getWidth();
	return (w <= B200_THREADS_COUNT ? w : B200_THREADS_COUNT) * B200_ALPHA;
}

int B200Aligner::getGridWidth(int width) {           /* the reference heuristic, R/src/CUDAligner.cpp:307-347 */
	int blocks = params->getBlocks();
	const int recommended = 4 * multiprocessors;
	int maximum = mustPruneBlocks() ? 1000 * recommended : recommended;
	if (blocks == 0 || width < (2 * blocks * B200_THREADS_COUNT)) {
		blocks = width / 2 / B200_THREADS_COUNT;
		if (blocks <= 1) {
			blocks = 1;
		} else {
			if (blocks > B200_MAX_BLOCKS_COUNT) blocks = B200_MAX_BLOCKS_COUNT;
			if (blocks <= multiprocessors) {
			} else if (blocks <= maximum) {
				blocks = (blocks / multiprocessors) * multiprocessors;
			} else {
				blocks = maximum;
			}
		}
	}
	return blocks;
}

void B200Aligner::initializeDiagonals() {
	Partition part = getPartition();
	const Grid* grid = getGrid();
	const int B = grid->getGridWidth();
	std::vector<int> split(B + 1);
	for (int bx = 0; bx < B; bx++) {
		int j0, j1;
		grid->getBlockPosition(bx, 0, NULL, &j0, NULL, &j1);
		split[bx] = j0;
		split[bx + 1] = j1;
	}
	b200_partition p;
	memset(&p, 0, sizeof(p));
	p.i0 = part.getI0(); p.j0 = part.getJ0(); p.i1 = part.getI1(); p.j1 = part.getJ1();
	p.recurrence = getRecurrenceType();
	p.first_row_init = getFirstRowInitType();
	p.first_col_init = getFirstColumnInitType();
	p.prune = mustPruneBlocks();
	check(b200_diag_begin(handle, &p, B, &split[0], getBlockHeight()), "b200_diag_begin");
	scoreBuffer.resize(B);
	colBuffer.resize((size_t)getBlockHeight() + 2);
	rowBuffer.resize((size_t)seq1_len + 2);
}

void B200Aligner::finalizeDiagonals() {
	check(b200_diag_end(handle), "b200_diag_end");
}

void B200Aligner::processDiagonal(int diagonal, int windowLeft, int windowRight) {
	check(b200_diag_process(handle, diagonal, windowLeft, windowRight), "b200_diag_process");
}

const cell_t* B200Aligner::getSpecialRow(int j, int len) {
	check(b200_diag_get_row(handle, j, len, (b200_cell*)&rowBuffer[0]), "b200_diag_get_row");
	return &rowBuffer[0];
}

const cell_t* B200Aligner::getLastRow(int j, int len) {
	check(b200_diag_get_row(handle, j, len, (b200_cell*)&rowBuffer[0]), "b200_diag_get_row");
	return &rowBuffer[0];
}

const cell_t* B200Aligner::getLastColumn(int i, int len) {
	check(b200_diag_get_last_column(handle, i, len, (b200_cell*)&colBuffer[0]), "b200_diag_get_last_column");
	return &colBuffer[0];
}

const score_t* B200Aligner::getBlockScores() {
	check(b200_diag_get_block_scores(handle, (b200_score*)&scoreBuffer[0]), "b200_diag_get_block_scores");
	return &scoreBuffer[0];
}

void B200Aligner::setFirstRow(const cell_t* cells, int j, int len) {
	check(b200_diag_set_first_row(handle, (const b200_cell*)cells, j, len), "b200_diag_set_first_row");
}

void B200Aligner::setFirstColumn(const cell_t* cells, int i, int len) {
	check(b200_diag_set_first_column(handle, (const b200_cell*)cells, i, len), "b200_diag_set_first_column");
}

void B200Aligner::clearPrunedBlocks(int b0, int b1) {   /* R/src/CUDAligner.cpp:511-517 */
	int p0, p1;
	getGrid()->getBlockPosition(b0, 0, NULL, &p0, NULL, NULL);
	getGrid()->getBlockPosition(b1, 0, NULL, &p1, NULL, NULL);
	if (p1 < 0) p1 = getPartition().getJ1();
	check(b200_diag_clear_pruned(handle, p0, p1), "b200_diag_clear_pruned");
}

/* ------------------------------------------------------------------------------------------------------------
 * statistics
 * ---------------------------------------------------------------------------------------------------------- */
void B200Aligner::clearStatistics() {
	AbstractDiagonalAligner::clearStatistics();
	fastCells = 0;
	fastDeviceMs = 0;
}

void B200Aligner::printInitialStatistics(FILE* file) {
	if (params->getGpuList().size() > 1) {
		fprintf(file, "B200 aligner extension: %d CUDA device(s), stage-1 chain over %d GPUs (first:", b200_device_count(), (int)params->getGpuList().size());
		fprintf(file, " %d)\n", params->getGpuList()[0]);
	} else
		fprintf(file, "B200 aligner extension: %d CUDA device(s), using GPU %d\n", b200_device_count(), params->getGPU() < 0 ? 0 : params->getGPU());
}

void B200Aligner::printStageStatistics(FILE* file) {
	fprintf(file, "Sequences on device: %d x %d\n", seq0_len, seq1_len);
}

void B200Aligner::printFinalStatistics(FILE* file) {
	fprintf(file, "B200 partitions: %lld whole-partition (persistent kernel; %lld of them on the multi-GPU chain), %lld chunked (%lld launches), %lld per-diagonal; kernel launches: %lld\n",
			fastPartitions, groupPartitions, chunkPartitions, chunkLaunches, diagPartitions, handle ? b200_kernel_launches(handle) : 0LL);
}

void B200Aligner::printStatistics(FILE* file) {
	AbstractDiagonalAligner::printStatistics(file);
	if (fastDeviceMs > 0) {
		fprintf(file, "B200 persistent kernel: %.3f ms device time, %lld cells, %.1f GCUPS\n", fastDeviceMs, fastCells,
				fastCells / fastDeviceMs / 1e6);
	}
}

long long B200Aligner::getProcessedCells() {
	return AbstractDiagonalAligner::getProcessedCells() + fastCells;
}

const char* B200Aligner::getProgressString() const {
	if (fastActive) return "PROGRESS: persistent strip kernel running";
	return AbstractDiagonalAligner::getProgressString();
}
