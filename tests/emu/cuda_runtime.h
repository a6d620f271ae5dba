// tests/emu/cuda_runtime.h -- TEST INFRASTRUCTURE: a CPU SIMT emulation of the CUDA subset that masa-cudalign_b200/csrc uses.
//
// `g++ -DB200_EMU -I tests/emu -x c++ masa-cudalign_b200/csrc/engine.cu tests/emu/emu_runtime.cpp` compiles the
// library's own kernel and engine sources, unchanged, into tests/emu/_build/libb200align_emu.so: same C ABI, no GPU.
// A kernel launch becomes one OS thread per co-resident CTA and one fiber per CUDA thread; warp collectives
// (__shfl*_sync, __ballot_sync, __reduce_max_sync, __syncwarp) and __syncthreads are rendezvous points between
// fibers; __shared__ is per-CTA storage; global memory is host memory; streams are worker threads that run their
// queue in order, so copies, host-mapped completion flags and several "devices" overlap as they do on the GPU.
// What this checks: the C++ semantics of the device code and of the strip-chain / multi-GPU protocols (it also flags
// s16 overflow of the packed arithmetic, which the hardware would silently wrap).  What it cannot check: anything
// that depends on the hardware (memory-ordering strength, convergence, timing).  Only tests load this library
// (tests/test_emu_cpu.py through B200_LIB); the product library libb200align.so is built by nvcc and has no CPU path.
#pragma once
// the standard headers first: __noinline__ below is also the spelling libstdc++ uses inside __attribute__((...))
#include <algorithm>
#include <atomic>
#include <climits>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <random>
#include <thread>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <functional>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static thread_local

// ---------------------------------------------------------------------------------------------- vector types
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
	unsigned x, y, z;
	dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline int2 make_int2(int x, int y) { int2 v; v.x = x; v.y = y; return v; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }

// ---------------------------------------------------------------------------------------------- runtime API subset
typedef int cudaError_t;
enum : int {
	cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600,
	cudaErrorPeerAccessAlreadyEnabled = 704, cudaErrorNotSupported = 801, cudaErrorLaunchFailure = 719,
};
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum : unsigned { cudaStreamNonBlocking = 1, cudaHostAllocMapped = 2, cudaHostAllocPortable = 1, cudaIpcMemLazyEnablePeerAccess = 1 };

namespace emu { struct Stream; struct Event; }
typedef emu::Stream* cudaStream_t;
typedef emu::Event* cudaEvent_t;
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaDeviceProp {
	char name[256];
	int major, minor, multiProcessorCount;
	size_t totalGlobalMem;
};

const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int* d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d);
cudaError_t cudaMemGetInfo(size_t* free_b, size_t* total_b);
cudaError_t cudaDeviceCanAccessPeer(int* can, int d, int peer);
cudaError_t cudaDeviceEnablePeerAccess(int peer, unsigned flags);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaMalloc(void** p, size_t bytes);
cudaError_t cudaFree(void* p);
cudaError_t cudaMallocHost(void** p, size_t bytes);
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned flags);
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s);
cudaError_t cudaMemset(void* dst, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void* dst, int v, size_t bytes, cudaStream_t s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamQuery(cudaStream_t s);
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p);
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void* p);
template <class K>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) { *n = 4; return cudaSuccess; }   // __launch_bounds__(128, 4)

// ---------------------------------------------------------------------------------------------- SIMT execution
namespace emu {

uint3& cur_thread_idx();
extern thread_local uint3 t_blockIdx, t_blockDim, t_gridDim;

// rendezvous of the calling fiber's warp: deposits `v`, returns the 32 deposited values once every live lane has arrived
const unsigned long long* warp_exchange(unsigned long long v);
void block_barrier();
void sleep_yield();                    // __nanosleep: let the other fibers of the CTA (and the other CTAs) run
void poll_tick();                      // called by the polling loads: a spin loop without __nanosleep still lets the others run
int lane_id();
char* smem_anchor();                   // any address inside this CTA's thread-local storage
unsigned sm_id_of_block();
void note_s16_overflow(int a, int b);

cudaError_t launch(cudaStream_t s, dim3 grid, dim3 block, std::function<void()> body);
template <class K, class... A>
cudaError_t launch_kernel(cudaStream_t s, dim3 grid, dim3 block, K kernel, A... args) {
	std::tuple<A...> t(args...);          // arguments are evaluated and copied at launch time, like kernel parameters
	return launch(s, grid, block, [kernel, t]() { std::apply(kernel, t); });
}

}  // namespace emu

#define threadIdx (emu::cur_thread_idx())
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)

// ---------------------------------------------------------------------------------------------- warp collectives
template <class T>
static inline unsigned long long emu_bits(T v) { static_assert(sizeof(T) <= 8, ""); unsigned long long b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T>
static inline T emu_unbits(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_exchange(0); }
static inline void __syncthreads() { emu::block_barrier(); }
template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int = 32) { const unsigned long long* b = emu::warp_exchange(emu_bits(v)); return emu_unbits<T>(b[src & 31]); }
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
	const int lane = emu::lane_id();
	const unsigned long long* b = emu::warp_exchange(emu_bits(v));
	return lane >= (int)d ? emu_unbits<T>(b[lane - d]) : v;
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
	const int lane = emu::lane_id();
	const unsigned long long* b = emu::warp_exchange(emu_bits(v));
	return emu_unbits<T>(b[(lane ^ m) & 31]);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
	const unsigned long long* b = emu::warp_exchange(pred ? 1ull : 0ull);
	unsigned m = 0;
	for (int k = 0; k < 32; k++) m |= (unsigned)(b[k] & 1ull) << k;
	return m;
}
static inline int __any_sync(unsigned mk, int pred) { return __ballot_sync(mk, pred) != 0; }
static inline int __reduce_max_sync(unsigned, int v) {
	const unsigned long long* b = emu::warp_exchange(emu_bits(v));
	int m = INT_MIN;
	for (int k = 0; k < 32; k++) m = max(m, emu_unbits<int>(b[k]));
	return m;
}

// ---------------------------------------------------------------------------------------------- memory
template <class T>
static inline T __ldcg(const T* p) {
	T v;
	if constexpr (sizeof(T) == 4) { unsigned b = __atomic_load_n(reinterpret_cast<const unsigned*>(p), __ATOMIC_RELAXED); memcpy(&v, &b, 4); }
	else if constexpr (sizeof(T) == 8) { unsigned long long b = __atomic_load_n(reinterpret_cast<const unsigned long long*>(p), __ATOMIC_RELAXED); memcpy(&v, &b, 8); }
	else { asm volatile("" ::: "memory"); memcpy(&v, p, sizeof(T)); asm volatile("" ::: "memory"); }
	return v;
}
template <class T>
static inline T __ldg(const T* p) { return *p; }
template <class T>
static inline void __stcg(T* p, T v) {
	if constexpr (sizeof(T) == 4) { unsigned b; memcpy(&b, &v, 4); __atomic_store_n(reinterpret_cast<unsigned*>(p), b, __ATOMIC_RELAXED); }
	else if constexpr (sizeof(T) == 8) { unsigned long long b; memcpy(&b, &v, 8); __atomic_store_n(reinterpret_cast<unsigned long long*>(p), b, __ATOMIC_RELAXED); }
	else { asm volatile("" ::: "memory"); memcpy(p, &v, sizeof(T)); asm volatile("" ::: "memory"); }
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) { emu::sleep_yield(); }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)(reinterpret_cast<const char*>(p) - emu::smem_anchor()); }

template <class T>
static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T>
static inline T atomicAdd_system(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T>
static inline T atomicSub(T* p, T v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }
template <class T>
static inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T>
static inline T atomicCAS(T* p, T expect, T v) { __atomic_compare_exchange_n(p, &expect, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return expect; }
static inline int atomicMax(int* p, int v) {
	int o = __atomic_load_n(p, __ATOMIC_RELAXED);
	while (o < v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {}
	return o;
}
static inline int atomicMin(int* p, int v) {
	int o = __atomic_load_n(p, __ATOMIC_RELAXED);
	while (o > v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {}
	return o;
}
static inline int atomicMax_system(int* p, int v) { return atomicMax(p, v); }

// ---------------------------------------------------------------------------------------------- integer intrinsics
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int emu_s16(unsigned v) { return (int)(short)(v & 0xffffu); }
// every packed s16 operation that ADDS reports a result outside s16 (the hardware wraps silently; the kernel's frame
// logic must never let that happen for a value that matters -- B200_EMU_OVERFLOW counts the events, see emu_runtime.cpp)
static inline unsigned emu_add16(int a, int b) {
	const int s = a + b;
	if (s > 32767 || s < -32768) emu::note_s16_overflow(a, b);
	return (unsigned)s & 0xffffu;
}
static inline unsigned __vadd2(unsigned a, unsigned b) { return emu_add16(emu_s16(a), emu_s16(b)) | (emu_add16(emu_s16(a >> 16), emu_s16(b >> 16)) << 16); }
static inline unsigned __vmaxs2(unsigned a, unsigned b) {
	const int lo = max(emu_s16(a), emu_s16(b)), hi = max(emu_s16(a >> 16), emu_s16(b >> 16));
	return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16);
}
static inline unsigned __viaddmax_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vadd2(a, b), c); }
static inline unsigned __vimax3_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vmaxs2(a, b), c); }
static inline unsigned __vibmax_s16x2(unsigned a, unsigned b, bool* pred_hi, bool* pred_lo) {
	*pred_lo = emu_s16(a) >= emu_s16(b);
	*pred_hi = emu_s16(a >> 16) >= emu_s16(b >> 16);
	return __vmaxs2(a, b);
}
static inline int __viaddmax_s32(int a, int b, int c) { return max((int)((unsigned)a + (unsigned)b), c); }
static inline int __viaddmax_s32_relu(int a, int b, int c) { return max(__viaddmax_s32(a, b, c), 0); }
