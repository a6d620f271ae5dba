#include "B200AlignerParameters.hpp"

#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sstream>

#include "../../include/b200align.h"

#define USAGE "\
--gpu=GPU               Selects the index of the GPU used for the computation.  \n\
                           Default: GPU 0. See --list-gpus. \n\
--list-gpus             Lists all available GPUs. \n\
--gpus=N | --gpus=A,B,..  Stage 1 on several GPUs of this box (NVLink chain: column  \n\
                           chunks dealt round-robin, borders through peer memory,    \n\
                           block pruning stays on). Stages 2-6 use the first one.    \n\
                           Replaces the reference's --fork/--split for one box.      \n\
--blocks=B              Run B blocks per external diagonal (compatibility path). \n\
--kernel=auto|s32|s16x2 DP kernel: packed s16x2 DPX lanes (ACGT inputs) or exact \n\
                           int32 lanes (any alphabet). Default: auto. \n\
--no-fast-path          Stage 1 through the per-diagonal compatibility path too. \n\
"

#define ARG_GPU        0x1001
#define ARG_LIST_GPUS  0x1002
#define ARG_BLOCKS     0x1003
#define ARG_KERNEL     0x1004
#define ARG_NO_FAST    0x1005
#define ARG_GPUS       0x1006

static struct option long_options[] = {
	{"gpu",          required_argument, 0, ARG_GPU},
	{"list-gpus",    no_argument,       0, ARG_LIST_GPUS},
	{"blocks",       required_argument, 0, ARG_BLOCKS},
	{"kernel",       required_argument, 0, ARG_KERNEL},
	{"no-fast-path", no_argument,       0, ARG_NO_FAST},
	{"gpus",         required_argument, 0, ARG_GPUS},
	{0, 0, 0, 0}
};

B200AlignerParameters::B200AlignerParameters() : gpu(-1), blocks(0), kernel(B200_KERNEL_AUTO), fastPath(true) {}
B200AlignerParameters::~B200AlignerParameters() {}

void B200AlignerParameters::printUsage() const {
	AbstractAlignerParameters::printFormattedUsage("B200 Specific Options", USAGE);
}

int B200AlignerParameters::processArgument(int argc, char** argv) {
	int ret = AbstractAlignerParameters::callGetOpt(argc, argv, long_options);
	switch (ret) {
	case ARG_GPU:
		if (optarg != NULL) sscanf(optarg, "%d", &gpu);
		break;
	case ARG_LIST_GPUS: {
		int n = b200_device_count();
		printf("%d CUDA device(s) visible\n", n);
		for (int i = 0; i < n; i++) printf("  GPU %d\n", i);
		exit(1);
	}
	case ARG_BLOCKS:
		if (optarg != NULL) {
			sscanf(optarg, "%d", &blocks);
			if (blocks > B200_MAX_BLOCKS_COUNT) {
				std::stringstream out;
				out << "Blocks count cannot be greater than " << B200_MAX_BLOCKS_COUNT << ".";
				setLastError(out.str().c_str());
				return -1;
			}
		}
		break;
	case ARG_KERNEL:
		if (optarg != NULL) {
			if (!strcmp(optarg, "auto")) kernel = B200_KERNEL_AUTO;
			else if (!strcmp(optarg, "s32")) kernel = B200_KERNEL_S32;
			else if (!strcmp(optarg, "s16x2")) kernel = B200_KERNEL_S16X2;
			else { setLastError("--kernel must be auto, s32 or s16x2."); return -1; }
		}
		break;
	case ARG_NO_FAST:
		fastPath = false;
		break;
	case ARG_GPUS:
		if (optarg != NULL) {
			gpuList.clear();
			if (strchr(optarg, ',') != NULL) {
				std::stringstream in(optarg);
				std::string tok;
				while (std::getline(in, tok, ',')) gpuList.push_back(atoi(tok.c_str()));
			} else {
				int n = atoi(optarg);
				for (int k = 0; k < n; k++) gpuList.push_back((gpu < 0 ? 0 : gpu) + k);
			}
			if (gpuList.empty() || gpuList.size() > 8) { setLastError("--gpus takes 1..8 GPUs."); return -1; }
		}
		break;
	default:
		return ret;
	}
	return 0;
}
