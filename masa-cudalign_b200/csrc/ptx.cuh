// ptx.cuh -- every inline-PTX statement of the library in one place: scoped loads / stores of the strip-chain protocol,
// the raw PRMT, shared-memory loads by 32-bit address, %globaltimer / %smid, and the kernel-launch macro.
//
// B200_EMU is defined only by the test suite's SIMT emulation build (tests/emu/: the same kernel sources compiled for the
// host CPU so that `-m "not gpu"` tests can run the device code against the oracle where no GPU exists); that build takes
// CPU restatements of exactly these primitives from tests/emu/emu_ptx.h.  The product library never defines it.
#pragma once
#ifdef B200_EMU
#include "emu_ptx.h"
#else
#include <cuda_runtime.h>

// kernel<<<grid, block, 0, stream>>>(args...)
#define B200_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)

namespace b200 {

__device__ __forceinline__ int ld_acquire(const int* p) {
	int v;
	asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
	asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed(const int* p) {
	int v;
	asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
	int v;
	asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ int ld_relaxed_sys(const int* p) {
	int v;
	asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
	asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
__device__ __forceinline__ unsigned sm_id() {
	unsigned v;
	asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
	return v;
}
// raw PRMT: unlike __byte_perm (which masks the selector with 0x7777) bit 3 of a selector nibble replicates the
// sign bit of the selected byte, which is how one instruction yields two sign-extended s16 scores
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel) {
	unsigned d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
	return d;
}
// shared-memory loads by 32-bit shared address (a generic pointer would have its 64-bit address materialised in the hot
// loop): `_pure` may be scheduled freely (read-only table), the other one is ordered with the surrounding memory operations
__device__ __forceinline__ unsigned lds32_pure(unsigned addr) {
	unsigned v;
	asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ unsigned lds32(unsigned addr) {
	unsigned v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
	return v;
}
// opaque copy: keeps a value in its register instead of letting the compiler re-derive it at every use
__device__ __forceinline__ void keep_in_register(unsigned& v) { asm volatile("mov.u32 %0, %0;" : "+r"(v)); }

}  // namespace b200
#endif
