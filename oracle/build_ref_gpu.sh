#!/bin/bash
# oracle/build_ref_gpu.sh -- TEST / MEASUREMENT INFRASTRUCTURE, not product code.
#
# Optional baseline of SURVEY.md section 8(d) "Reference GPU baseline": the reference's OWN CUDA aligner
# (R/src/CUDAligner.cu + CUDAligner.cpp + cuda_util.cpp, MASA-CUDAlign 4.0.2.1028) built for sm_100 so that the strip
# kernels can be compared with the reference's kernels on the same B200 instead of with a CPU.  The reference does not
# compile with CUDA 12 (texture references were removed from the toolkit, R/src/CUDAligner.cu:74-94), so this recipe
# applies a COMPATIBILITY PATCH to a scratch copy under /tmp -- nothing of the reference is copied into the repo:
#   * texture<T,1> t_seq0 / t_seq1 / t_busH      ->  __device__ const T* pointers
#   * tex1Dfetch(t, i)                          ->  __ldg(t + i)     (same read-only data path on sm_100)
#   * cudaBindTexture / cudaUnbindTexture       ->  cudaMemcpyToSymbol of the pointer (only when it changes) / nothing
# Kernels, grid policy (getGridWidth: B = 444 blocks of T = 128 threads on 148 SMs), host loop and MASA-Core are the
# reference's, unmodified.  Output: oracle/_ref/cudalign_ref_gpu (git-ignored; travels to the GPU box).  Results it
# produces are labelled "reference kernel, texture refs removed" wherever they are quoted (profiles/).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${REF:-/root/reference/masa-cudalign-4.0.2.1028}
OUT=$HERE/_ref
TMP=$(mktemp -d /tmp/refgpu.XXXXXX)
[ -f "$OUT/libmasa.a" ] || { echo "build oracle/_ref/libmasa.a first (make -C oracle ref)"; exit 1; }
cp "$REF"/src/*.cu "$REF"/src/*.cpp "$REF"/src/*.hpp "$REF"/src/*.h "$TMP"/
cat > "$TMP/config.h" <<'CFG'
#define COMPILED_CUDA_ARCH "sm_100"
#define PACKAGE_STRING "MASA-CUDAlign 4.0.2.1028"
#define PACKAGE_VERSION "4.0.2.1028"
#define VERSION "4.0.2.1028"
CFG
cat > "$TMP/compat_tex.h" <<'CPT'
// compatibility shim of oracle/build_ref_gpu.sh: linear textures -> global pointers read with __ldg
#include <cuda_runtime.h>
template <class T> static inline cudaError_t compat_bind(const T* const& symbol, const void* ptr) {
	return cudaMemcpyToSymbol(symbol, &ptr, sizeof(ptr));
}
CPT
sed -i \
  -e 's|^texture<unsigned char, 1, cudaReadModeElementType> t_seq0;|#include "compat_tex.h"\n__device__ const unsigned char* t_seq0;|' \
  -e 's|^texture<unsigned char, 1, cudaReadModeElementType> t_seq1;|__device__ const unsigned char* t_seq1;|' \
  -e 's|^texture<         int2, 1, cudaReadModeElementType> t_busH;|__device__ const int2* t_busH;|' \
  -e 's|tex1Dfetch(t_seq0,\([^)]*\))|__ldg(t_seq0+(\1))|g' \
  -e 's|tex1Dfetch(t_seq1, *\([^)]*\))|__ldg(t_seq1+(\1))|g' \
  -e 's|tex1Dfetch(t_busH, *\([^)]*\))|__ldg(t_busH+(\1))|g' \
  -e 's|cudaBindTexture(0, t_seq0, seq0, seq0_len)|compat_bind(t_seq0, seq0)|' \
  -e 's|cudaBindTexture(0, t_seq1, seq1, seq1_len)|compat_bind(t_seq1, seq1)|' \
  -e 's|cutilSafeCall(cudaBindTexture(0, t_busH, cuda->d_busH, cuda->busH_size));|{ static const void* bound = 0; if (bound != (const void*)cuda->d_busH) { cutilSafeCall(compat_bind(t_busH, cuda->d_busH)); bound = cuda->d_busH; } }|' \
  -e 's|cutilSafeCall(cudaUnbindTexture(t_[a-zA-Z0-9]*));|;|' \
  "$TMP/CUDAligner.cu"
if grep -n "tex1Dfetch\|cudaBindTexture\|cudaUnbindTexture\|^texture<" "$TMP/CUDAligner.cu"; then echo "compat patch incomplete"; exit 1; fi
CORE=$REF/libs/masa-core/src
INC="-I$TMP -I$OUT/gen -I$OUT/gen/sub -I$CORE -I$REF/libs/masa-core"
NVCC=${NVCC:-nvcc}
FLAGS="-Wno-deprecated-gpu-targets -O3 -gencode arch=compute_100,code=sm_100 -DTHREADS_COUNT=128 -ftz=true -prec-sqrt=false -prec-div=false -w -Xcompiler -fno-strict-aliasing,-fpermissive,-w"
for f in CUDAligner.cu cuda_util.cpp CUDAligner.cpp CUDAlignerParameters.cpp main.cpp; do
  $NVCC $FLAGS $INC -x cu -c "$TMP/$f" -o "$TMP/$f.o"
done
# -lcuda like the reference's Makefile.am (cuda_util.cpp calls cuMemGetInfo); the stub library serves the link where no driver is installed
$NVCC -Wno-deprecated-gpu-targets -gencode arch=compute_100,code=sm_100 -o "$OUT/cudalign_ref_gpu" "$TMP"/CUDAligner.cu.o "$TMP"/CUDAligner.cpp.o "$TMP"/cuda_util.cpp.o \
      "$TMP"/CUDAlignerParameters.cpp.o "$TMP"/main.cpp.o "$OUT/libmasa.a" -L/usr/local/cuda/lib64/stubs -lcuda -lpthread
rm -rf "$TMP"
ls -la "$OUT/cudalign_ref_gpu"
