// cudalign entry point: MASA-Core owns main() and the CLI (C/libmasa/libmasa.cpp:762); the extension only
// supplies the aligner object, exactly like R/src/main.cpp:40.
#include "B200Aligner.hpp"

#define HEADER "cudalign-b200  -  MASA-CUDAlign aligner extension for NVIDIA B200 (sm_100a)\033[0m\n"

int main(int argc, char** argv) {
	return libmasa_entry_point(argc, argv, new B200Aligner(), (char*)HEADER);
}
