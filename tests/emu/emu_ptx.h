// tests/emu/emu_ptx.h -- TEST INFRASTRUCTURE: CPU restatements of the inline-PTX wrappers of masa-cudalign_b200/csrc/ptx.cuh
// for the SIMT emulation build (see tests/emu/cuda_runtime.h).  Same names, same meaning, no PTX.
#pragma once
#include <cuda_runtime.h>      // tests/emu/cuda_runtime.h (found through -I tests/emu)

#define B200_LAUNCH(kernel, grid, block, stream, ...) emu::launch_kernel((stream), dim3(grid), dim3(block), kernel, __VA_ARGS__)

namespace b200 {

// scoped loads / stores of the strip-chain protocol: C++ acquire / release on host memory.  The polling loads also
// tick the fiber scheduler, so that a spin loop without __nanosleep cannot starve the fiber it is waiting for.
static inline int ld_acquire(const int* p) { emu::poll_tick(); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void st_release(int* p, int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline int ld_relaxed(const int* p) { emu::poll_tick(); return __atomic_load_n(p, __ATOMIC_RELAXED); }
static inline int ld_acquire_sys(const int* p) { emu::poll_tick(); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline int ld_relaxed_sys(const int* p) { emu::poll_tick(); return __atomic_load_n(p, __ATOMIC_RELAXED); }
static inline void st_release_sys(int* p, int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }

static inline unsigned long long global_ns() {
	timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}
static inline unsigned sm_id() { return emu::sm_id_of_block(); }

// prmt.b32 d, a, b, sel (default mode): result byte k = byte (sel nibble k & 7) of {b:a}; bit 3 of the nibble replicates
// the sign bit of that byte instead
static inline unsigned prmt(unsigned a, unsigned b, unsigned sel) {
	const unsigned long long src = ((unsigned long long)b << 32) | a;
	unsigned d = 0;
	for (int k = 0; k < 4; k++) {
		const unsigned nib = (sel >> (4 * k)) & 15u;
		unsigned byte = (unsigned)(src >> (8 * (nib & 7u))) & 0xffu;
		if (nib & 8u) byte = (byte & 0x80u) ? 0xffu : 0x00u;
		d |= byte << (8 * k);
	}
	return d;
}
// 32-bit "shared addresses" are offsets from an anchor inside the CTA's thread-local storage (__cvta_generic_to_shared)
static inline unsigned lds32_pure(unsigned addr) { return *reinterpret_cast<const unsigned*>(emu::smem_anchor() + (int)addr); }
static inline unsigned lds32(unsigned addr) {
	asm volatile("" ::: "memory");
	return *reinterpret_cast<const volatile unsigned*>(emu::smem_anchor() + (int)addr);
}
static inline void keep_in_register(unsigned&) {}

}  // namespace b200
