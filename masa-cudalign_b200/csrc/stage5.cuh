// stage5.cuh -- device side of the batched traceback (stage 5).
//
// Replaces the serial loop of C/stage5/sw_stage5.cpp:404-424 (one static 1024 x 1024 (H,E,F) table, every partition
// of crosspoint_04 aligned and walked back one after the other on one CPU thread) by ONE THREAD PER PARTITION: after
// stage 4 the partitions are at most --maximum-partition (default 16) cells on a side, millions of them on a
// chromosome-sized alignment, and they are independent.
//
// A thread sweeps its partition row by row with one rolling (H, E) row and records, per cell, the five equalities the
// reference's walk tests on its full tables (:213-257):
//     bit 0  H == H(i-1,j-1) + s       bit 1  H == E        bit 2  H == F
//     bit 3  E == H(i-1,j) - first     bit 4  F == H(i,j-1) - first          (E vertical, F horizontal, as in stage 5)
// in the same int32 arithmetic (the -INF borders of :134-143 drift exactly like the reference's), then walks back from
// the bottom-right corner with the reference's precedence (diagonal, vertical, horizontal; a gap step returns to the
// MATCH state when bit 3 / bit 4 says the gap was opened in that cell) and emits one byte per step.  The host replays
// the bytes into Alignment::addGapInSeq0/1 in the reference's call order, so alignment.NN.bin is byte-identical.
//
// Two instances of the same code: partitions up to 32 x 32 keep row and flags in thread-local memory (interleaved per
// thread by the hardware = coalesced), larger ones (--maximum-partition up to 1024) in a global scratch area.
#pragma once
#include "strip_common.cuh"

namespace b200 {

struct S5Part {
	int i0, j0;                  // start crosspoint (0-based prefix lengths)
	int di, dj;                  // rows / columns, both > 0 (pure-gap partitions never reach the device)
	int ts, te;                  // crosspoint types at the start / end: 0 MATCH, 1 GAP_1, 2 GAP_2
	long long op_off;            // first slot of the partition in the ops buffer (di + dj slots)
	long long row_off;           // global variant: first int of the 2 * (dj + 1) row scratch
	long long flag_off;          // global variant: first byte of the di * dj flag scratch
	int out_index;               // partition number (index of the end crosspoint)
	int pad;
};

struct S5Out { int n_ops, matches, mismatches, gap_open, gap_ext, score; };      // counters == total_score_t (:51-67)

constexpr int kS5Local = 32;     // largest side kept in thread-local memory

// __host__ too: tests/test_stage5_host_cpu.py compiles this very function for the CPU and runs it against the reference's
// golden vectors where no GPU exists (the -m gpu tests run the kernels themselves).

__host__ __device__ __forceinline__ void s5_partition(const unsigned char* __restrict__ seq0, const unsigned char* __restrict__ seq1,
                                             const S5Part& p, int* hrow, int* erow, unsigned char* fl, int fstride,
                                             unsigned char* __restrict__ ops, S5Out& o) {
	const unsigned char* s0 = seq0 + p.i0;
	const unsigned char* s1 = seq1 + p.j0;
	const int di = p.di, dj = p.dj;
	// first row (:134-139) and the rolling first-column cell (:141-142)
	hrow[0] = p.ts != 0 ? -kInf : 0;
	for (int j = 1; j <= dj; j++) { hrow[j] = -j * kGapExt - (p.ts != 1 ? kGapOpen : 0); erow[j] = -kInf; }
	for (int i = 1; i <= di; i++) {
		int diag = hrow[0];
		int hleft = -i * kGapExt - (p.ts != 2 ? kGapOpen : 0);
		hrow[0] = hleft;
		int f = -kInf;
		const unsigned char c0 = s0[i - 1];
		unsigned char* frow = fl + (size_t)(i - 1) * fstride;
		for (int j = 1; j <= dj; j++) {
			const int up = hrow[j];
			const int eo = up - kGapFirst, fo = hleft - kGapFirst;
			const int e = max(eo, erow[j] - kGapExt);
			const int fv = max(fo, f - kGapExt);
			const int d = diag + (c0 == s1[j - 1] ? kMatch : kMismatch);
			const int hv = max(d, max(e, fv));
			frow[j - 1] = (unsigned char)((hv == d) | ((hv == e) << 1) | ((hv == fv) << 2) | ((e == eo) << 3) | ((fv == fo) << 4));
			diag = up; hrow[j] = hv; erow[j] = e; f = fv; hleft = hv;
		}
	}
	// the walk (:209-308); an end crosspoint of type MATCH starts in the MATCH state at the same corner (:197-200)
	int i = di, j = dj, c = p.te, n = 0, sum = 0;
	int matches = 0, mismatches = 0, gap_open = 0, gap_ext = 0;
	while (i > 0 && j > 0) {
		const int fg = fl[(size_t)(i - 1) * fstride + (j - 1)];
		int dir;
		if (c == 0) dir = (fg & 1) ? 0 : ((fg & 2) ? 1 : 2);
		else dir = c == 2 ? 1 : 2;
		if (dir == 0) {
			c = 0;
			if (s0[i - 1] == s1[j - 1]) { matches++; sum += kMatch; } else { mismatches++; sum += kMismatch; }
			i--; j--;
		} else {
			const bool opened = dir == 1 ? (fg & 8) != 0 : (fg & 16) != 0;
			c = opened ? 0 : (dir == 1 ? 2 : 1);
			gap_ext++;
			if (opened) { gap_open++; sum -= kGapFirst; } else sum -= kGapExt;
			if (dir == 1) i--; else j--;
		}
		ops[n++] = (unsigned char)dir;
	}
	for (; i > 0; i--) { ops[n++] = 1; gap_ext++; c = 2; sum -= kGapExt; }
	for (; j > 0; j--) { ops[n++] = 2; gap_ext++; c = 1; sum -= kGapExt; }
	if (p.ts == 0 && c != 0) sum -= kGapOpen;          // :309-311: charged to the score, not counted as an opening
	o.n_ops = n; o.matches = matches; o.mismatches = mismatches; o.gap_open = gap_open; o.gap_ext = gap_ext; o.score = sum;
}

__global__ void __launch_bounds__(64) s5_local_kernel(const unsigned char* __restrict__ seq0, const unsigned char* __restrict__ seq1,
                                                       const S5Part* __restrict__ parts, int nparts, unsigned char* __restrict__ ops,
                                                       S5Out* __restrict__ out) {
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= nparts) return;
	const S5Part p = parts[k];
	int hrow[kS5Local + 1], erow[kS5Local + 1];
	unsigned char fl[kS5Local * kS5Local];
	S5Out o;
	s5_partition(seq0, seq1, p, hrow, erow, fl, kS5Local, ops + p.op_off, o);
	out[k] = o;
}

__global__ void __launch_bounds__(64) s5_global_kernel(const unsigned char* __restrict__ seq0, const unsigned char* __restrict__ seq1,
                                                        const S5Part* __restrict__ parts, int nparts, unsigned char* __restrict__ ops,
                                                        int* __restrict__ rows, unsigned char* __restrict__ flags, S5Out* __restrict__ out) {
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= nparts) return;
	const S5Part p = parts[k];
	S5Out o;
	s5_partition(seq0, seq1, p, rows + p.row_off, rows + p.row_off + p.dj + 1, flags + p.flag_off, p.dj, ops + p.op_off, o);
	out[k] = o;
}

}  // namespace b200
