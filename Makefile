# Top-level build: the product library (CUDA, sm_100a only) and, when the reference mount is present,
# the reference-linked artefacts (oracle/_ref/*, build/cudalign).
NVCC      ?= nvcc
PKG       := masa-cudalign_b200
NVFLAGS   := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
LIB       := $(PKG)/libb200align.so
CSRC      := $(wildcard $(PKG)/csrc/*.cu) $(wildcard $(PKG)/csrc/*.cuh) include/b200align.h

.PHONY: all lib oracle cudalign clean
all: lib oracle

lib: $(LIB)
$(LIB): $(CSRC)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(PKG)/csrc/engine.cu

oracle:
	$(MAKE) -C oracle all

clean:
	rm -f $(LIB)
	$(MAKE) -C oracle clean
