// B200Aligner -- the reference-side adapter: a MASA-Core aligner extension (IAligner) whose device work is done
// by libb200align.so through the C ABI of include/b200align.h.  Drop-in for class CUDAligner
// (R/src/CUDAligner.hpp:192): same base class, same protected virtuals, same capabilities, same error
// convention (print to stderr and exit, R/src/cuda_util.h:34-61).
//
//  * Stages 2/3 (goal matching, early stop) and any partition that must stream its last column run through the
//    reference's own AbstractDiagonalAligner loop (C/libmasa/aligners/AbstractDiagonalAligner.cpp:59-159); each
//    protected virtual below forwards to one b200_diag_* call.
//  * Stage 1 (no last column wanted) takes the B200-first path: alignPartition() is overridden and hands the whole
//    partition to b200_align_partition (one persistent kernel); the special rows, last row/cell and best score
//    come back through callbacks that forward to the IManager delegates of AbstractAligner.
#ifndef B200ALIGNER_HPP_
#define B200ALIGNER_HPP_

#include <vector>

#include "libmasa/libmasa.hpp"
#include "B200AlignerParameters.hpp"
#include "../../include/b200align.h"

#define B200_THREADS_COUNT 128      /* THREADS_COUNT as configured by the reference build (R/configure.ac:79-83) */
#define B200_ALPHA 4                /* rows per thread, R/src/CUDAligner.hpp:62 */

class B200Aligner : public AbstractDiagonalAligner {
public:
	B200Aligner();
	virtual ~B200Aligner();

	/* IAligner (C/libmasa/IAligner.hpp:159-377) */
	virtual aligner_capabilities_t getCapabilities();
	virtual IAlignerParameters* getParameters();
	virtual const score_params_t* getScoreParameters();
	virtual void initialize();
	virtual void finalize();
	virtual void setSequences(const char* seq0, const char* seq1, int seq0_len, int seq1_len);
	virtual void unsetSequences();
	virtual void alignPartition(Partition partition);
	virtual match_result_t matchLastColumn(const cell_t* buffer, const cell_t* base, int len, int goalScore);
	virtual void clearStatistics();
	virtual void printInitialStatistics(FILE* file);
	virtual void printStageStatistics(FILE* file);
	virtual void printFinalStatistics(FILE* file);
	virtual void printStatistics(FILE* file);
	virtual long long getProcessedCells();
	virtual const char* getProgressString() const;

	/* the handle of the (single) initialised aligner of this process: used by the GPU stage 4 (host/stage4_gpu.cpp) */
	static b200_handle* activeHandle();

protected:
	/* AbstractDiagonalAligner virtuals == the ones CUDAligner fills (R/src/CUDAligner.hpp:216-232) */
	virtual int getGridWidth(int width);
	virtual int getBlockHeight();
	virtual const cell_t* getSpecialRow(int j, int len);
	virtual const cell_t* getLastRow(int j, int len);
	virtual const cell_t* getLastColumn(int i, int len);
	virtual const score_t* getBlockScores();
	virtual void setFirstRow(const cell_t* cells, int j, int len);
	virtual void setFirstColumn(const cell_t* cells, int i, int len);
	virtual void clearPrunedBlocks(int b0, int b1);
	virtual void initializeDiagonals();
	virtual void processDiagonal(int diagonal, int windowLeft, int windowRight);
	virtual void finalizeDiagonals();

private:
	B200AlignerParameters* params;
	score_params_t score_params;
	b200_handle* handle;            /* the GPU of stages 2-6 and of small partitions (rank 0 of the group with --gpus) */
	b200_group* group;              /* --gpus=N: the GPUs of the stage-1 chain (created when the sequence sizes are known) */
	long long groupRows, groupJobs; /* capacity the group was created for */
	bool groupSeqValid;             /* the current sequences are resident on every GPU of the group */
	const char* seq0_ptr; const char* seq1_ptr;
	long long groupPartitions;
	int multiprocessors;
	int seq0_len, seq1_len;
	bool fastActive;
	long long fastCells;
	double fastDeviceMs;
	long long fastPartitions, diagPartitions, chunkPartitions, chunkLaunches;
	Partition fastPartition;
	/* final chunk of a chunked partition: last row and last-column chunks are buffered and replayed in the
	 * reference's external-diagonal order (see alignPartitionChunked) */
	bool bufferingTail;
	bool tailFirstCellSeen;
	std::vector<cell_t> tailCol, tailRow;
	cell_t tailRowFirst;
	std::vector<cell_t> rowBuffer, colBuffer;
	std::vector<score_t> scoreBuffer;

	void check(int rc, const char* what);
	bool canUseFastPath();
	bool canUseChunkPath();
	void alignPartitionFast(Partition partition);
	bool useGroupFor(Partition partition);
	void ensureGroup();
	void alignPartitionChunked(Partition partition);
	void fillPartition(b200_partition& p, Partition partition);

	/* C callbacks of b200_align_partition -> IManager delegates */
	static void cbReceiveFirstRow(void* ctx, b200_cell* buffer, int len);
	static void cbReceiveFirstColumn(void* ctx, b200_cell* buffer, int len);
	static void cbDispatchRow(void* ctx, int i, const b200_cell* buffer, int len);
	static void cbDispatchColumn(void* ctx, int j, const b200_cell* buffer, int len);
	static void cbDispatchScore(void* ctx, b200_score score);
	static int cbMustContinue(void* ctx);
};

#endif
