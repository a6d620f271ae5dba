// B200AlignerParameters -- command-line flags of the B200 aligner extension.
// Keeps the reference's flag names (R/src/CUDAlignerParameters.cpp:33-54,81-110): --gpu, --list-gpus, --blocks;
// adds --kernel, --no-fast-path and --gpus (multi-GPU stage 1 on one box; the reference uses --fork, libmasa.cpp:540-642).  Unknown MASA-Core flags never reach this class (libmasa.cpp:1203-1213).
#ifndef B200ALIGNERPARAMETERS_HPP_
#define B200ALIGNERPARAMETERS_HPP_

#include <vector>
#include "libmasa/libmasa.hpp"

#define B200_MAX_BLOCKS_COUNT 512        /* MAX_BLOCKS_COUNT, R/src/CUDAligner.hpp:55 */

class B200AlignerParameters : public AbstractAlignerParameters {
public:
	B200AlignerParameters();
	virtual ~B200AlignerParameters();
	virtual void printUsage() const;
	virtual int processArgument(int argc, char** argv);

	int getGPU() const { return gpu; }
	int getBlocks() const { return blocks; }
	int getKernel() const { return kernel; }
	bool useFastPath() const { return fastPath; }
	const std::vector<int>& getGpuList() const { return gpuList; }     // --gpus: devices of the stage-1 chain (empty or 1 entry: single GPU)

private:
	int gpu;        // -1: device 0 (all B200s of a box are identical; the reference picks "the fastest")
	int blocks;     // forced grid width of the diag path, 0 = heuristic
	int kernel;     // B200_KERNEL_*
	bool fastPath;  // whole-partition persistent kernel for stage 1
	std::vector<int> gpuList;
};

#endif
