# Top-level build: the product library (CUDA, sm_100a only) and, when the reference mount is present,
# the reference-linked artefacts (oracle/_ref/*, build/cudalign).
NVCC      ?= nvcc
PKG       := masa-cudalign_b200
NVFLAGS   := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
LIB       := $(PKG)/libb200align.so
CSRC      := $(wildcard $(PKG)/csrc/*.cu) $(wildcard $(PKG)/csrc/*.cuh) $(wildcard $(PKG)/csrc/*.inl) include/b200align.h

.PHONY: all lib oracle cudalign refgpu clean
all: lib oracle

lib: $(LIB)
$(LIB): $(CSRC)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(PKG)/csrc/engine.cu

oracle:
	$(MAKE) -C oracle all

# the drop-in binary (needs the read-only reference mount: MASA-Core is compiled from there into build/ref/)
cudalign: lib
	$(MAKE) -C $(PKG)/host

# optional measurement baseline: the reference's own CUDA aligner, texture references patched out (oracle/_ref/cudalign_ref_gpu)
refgpu: oracle
	bash oracle/build_ref_gpu.sh

clean:
	rm -f $(LIB)
	$(MAKE) -C oracle clean
