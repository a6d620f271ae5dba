// Pure step rate of the packed strip kernel: runs StripS16<16,SW,TRACK,LUT>::step<false,false> (the steady-state
// wavefront step of csrc/strip_s16.cuh, unchanged) in a loop with NO strip chaining, NO border traffic and NO
// publication, on every SM at a chosen number of resident warps.  The difference between this rate and the rate of
// the full kernel is the cost of the orchestration (progress waits, border loads, fences); the difference to the
// pipe roofline is the cost of the dependency chains inside a step.  Development aid, not part of the product path.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../masa-cudalign_b200/csrc -o step_rate step_rate.cu
#include <cstdio>
#include <cstdlib>
#include <climits>
#include "strip_s16.cuh"

using namespace b200;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int R, bool TRACK, bool LUT, int MINB>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, MINB) step_rate_kernel(StripParams p, int blocks32, unsigned* out, const unsigned* seed) {
	using K = StripS16<R, true, TRACK, LUT>;
	__shared__ typename K::Smem smu[kWarpsPerBlock];
	extern __shared__ unsigned lut[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	typename K::Smem& sm = smu[warp];
	unsigned lut_addr = 0;
	if (LUT) {
		for (int idx = threadIdx.x; idx < 256 * 32; idx += blockDim.x) {
			const int e = idx >> 5, cq = e & 15, dq = e >> 4;
			const int lo = ((cq & 3) == (dq & 3)) ? kMatch + kGapFirst : kMismatch + kGapFirst;
			const int hi = ((cq >> 2) == (dq >> 2)) ? kMatch + kGapFirst : kMismatch + kGapFirst;
			lut[idx] = (unsigned)lo | ((unsigned)hi << 16);
		}
		__syncthreads();
		lut_addr = (unsigned)__cvta_generic_to_shared(lut);
	}
	StripJob jb; jb.i0 = 0; jb.rows = K::SH; jb.j0 = 0; jb.cols = INT_MAX; jb.dep = -1; jb.flags = 0; jb.left_off = 0; jb.right_off = -1;
	jb.sra_off = -1; jb.sra_index = 0;
	typename K::State s;
	s.lut = lut_addr + 4u * (unsigned)lane;
	unsigned x = seed[threadIdx.x & 31] + blockIdx.x * 977u;
#pragma unroll
	for (int r = 0; r < R; r++) {
		x = x * 1664525u + 1013904223u;
		const unsigned cl = (x >> 8) & 3u, ch = (x >> 12) & 3u;
		s.T[r] = dup2(-kGapFirst); s.E[r] = dup2(kNeg);
		if (LUT) s.sel[r] = (cl | (ch << 2)) << 7;
		else s.sel[r] = cl | ((8u | cl) << 4) | ((4u + ch) << 8) | ((12u + ch) << 12);
	}
	s.tprev = dup2(-kGapFirst); s.botH = 0; s.botF = 0; s.pa = LUT ? 0u : 0x02020202u; s.pb = s.pa;
	s.Zp = seed[31] >> 31;   // zero, but not a compile-time constant: a literal 0 makes ptxas materialise it once per row (17 PRMT RZ per step)
	s.base = 0; s.bs = INT_MIN; s.bi = -1; s.bj = -1; s.thr = 30000; s.pub = INT_MIN; s.thrp = dup2(30000); s.ncand = 0;
	s.blk = 0x80008000u;
	const int vo = (K::SH - 1) / R, ro = (K::SH - 1) % R;
	for (int b = 0; b < blocks32; b++) {
		x = x * 1664525u + 1013904223u;
		sm.top[lane] = make_uint4(0u, (unsigned)kNeg << 16, LUT ? (((x >> 10) & 3u) << 11) : profile_word("ACGT"[(x >> 10) & 3u]), 0u);
		__syncwarp();
#pragma unroll kStepUnroll
		for (int u = 0; u < 32; u++)
			K::template step<false, false>(p, jb, s, sm, warp, lane, 64 + b * 32 + u, u, R, R, vo, ro, 0, INT_MAX);
		// keep the frame bounded like the rebase of the real kernel (cheap, every 32 steps)
		if ((b & 15) == 15) {
#pragma unroll
			for (int r = 0; r < R; r++) { s.T[r] = __vmaxs2(__vmins2(s.T[r], dup2(2000)), dup2(-2000)); }
		}
		__syncwarp();
	}
	unsigned acc = s.botH ^ s.botF ^ s.blk ^ (unsigned)s.ncand;
#pragma unroll
	for (int r = 0; r < R; r++) acc ^= s.T[r] ^ s.E[r];
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int R, bool TRACK, bool LUT, int MINB = 4>
void run(const char* name, int sms, int warps_per_sm, int blocks32) {
	StripParams p; memset(&p, 0, sizeof(p));
	const int grid = sms * warps_per_sm / kWarpsPerBlock;
	unsigned* out; unsigned* seed;
	CK(cudaMalloc(&out, sizeof(unsigned) * grid * kWarpsPerBlock * 32));
	CK(cudaMalloc(&seed, 32 * sizeof(unsigned)));
	unsigned hs[32]; for (int i = 0; i < 32; i++) hs[i] = 12345u * (i + 1);
	CK(cudaMemcpy(seed, hs, sizeof(hs), cudaMemcpyHostToDevice));
	auto k = step_rate_kernel<R, TRACK, LUT, MINB>;
	int occ = 0; CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, LUT ? kLutBytes : 4));
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, kWarpsPerBlock * 32, LUT ? kLutBytes : 4));
	if (occ * kWarpsPerBlock < warps_per_sm) { printf("%-34s warps/SM=%2d  not resident (max %d)\n", name, warps_per_sm, occ * kWarpsPerBlock); return; }
	CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, LUT ? kLutBytes : 4));
	const int dyn = LUT ? kLutBytes : 4;
	k<<<grid, kWarpsPerBlock * 32, dyn>>>(p, 8, out, seed);
	CK(cudaDeviceSynchronize());
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	CK(cudaEventRecord(e0));
	k<<<grid, kWarpsPerBlock * 32, dyn>>>(p, blocks32, out, seed);
	CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
	float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
	const double cells = (double)grid * kWarpsPerBlock * 64.0 * R * 32.0 * blocks32;
	printf("%-34s warps/SM=%2d  %8.1f GCUPS  (%.2f ms)  %.2f cells/clk/SM at 1965 MHz\n", name, warps_per_sm, cells / ms / 1e6, ms,
	       cells / (ms * 1e-3) / sms / 1.965e9);
	CK(cudaFree(out)); CK(cudaFree(seed));
}

int main() {
	cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
	const int sms = pr.multiProcessorCount;
	printf("# %s SMs=%d\n", pr.name, sms);
	for (int w : {8, 12, 16}) {
		run<16, true, true>("R=16 SW track LUT (production)", sms, w, 20000);
		run<20, true, true>("R=20 SW track LUT", sms, w, 16000);
		run<24, true, true>("R=24 SW track LUT", sms, w, 13000);
		run<24, true, true, 3>("R=24 SW track LUT 168reg", sms, w, 13000);
		run<32, true, true, 3>("R=32 SW track LUT 168reg", sms, w, 10000);
		run<16, false, true>("R=16 SW no-track LUT", sms, w, 20000);
	}
	return 0;
}
