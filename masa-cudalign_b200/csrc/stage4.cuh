// stage4.cuh -- device side of the batched Myers-Miller partition split (stage 4).
//
// Replaces ort_split_2 / processCol / match of C/stage4/sw_stage4.cpp:254-380 (4 pthreads, scalar) by
//   1. s4_fill_kernel     gap-initialised first rows / first columns of every half-partition,
//   2. the strip kernels  (NW, no tracking) over ALL forward and reverse half-matrices of a round as independent
//                         jobs: the last row (H,F) of a half is exactly the r0[] / r1[] vector of ort_split_2,
//   3. s4_match_kernel    one warp per partition scans the middle row outward from the middle column in the
//                         reference's order (forward candidate, then backward candidate, :347-375); first hit wins,
//                         an overshooting sum is the reference's fatal "Error Match".
// The halves are computed completely (the reference stops at the first hit), which costs at most 2x the cells
// but keeps every partition of a round inside four launches.
#pragma once
#include "strip_common.cuh"

namespace b200 {

struct XPoint { int i, j, type, score; };          // == crosspoint_t (C/common/Crosspoint.hpp:30-40)

struct S4Half {                  // one half-matrix (forward or reverse) of one partition
	int bus_off;                 // first column in the group's busH (index into the column sequence)
	int cols;
	int row_open;                // 1: first row h = -(j+1)*ext - open, 0: without open
	long long left_off;          // slot 0 (corner) of the half's left border
	int rows;
	int col_open;                // first column with / without gap-open
	int corner;                  // H of the corner cell: 0 or -INF
	int pad;
};

struct S4Part {                  // geometry of one partition for the matcher
	int fwd_bus, rev_bus;        // first column of the forward / reverse half in their busH arrays
	long long fwd_left, rev_left;   // left-border slots of the halves (slot rows = first-column cell of the last row)
	int len1;                    // columns (after a possible transposition)
	int imid0, imid1;
	int diff;                    // score_e - score_s
	int i0, j0, score_s;
	int transposed;
	int grp;                     // 0: normal, 1: transposed (selects the busH pair)
	int out_index;
};

__global__ void s4_fill_kernel(const S4Half* halves, int nhalves, Cell* busH, Cell* left) {
	const S4Half hf = halves[blockIdx.x];
	for (int j = threadIdx.x; j < hf.cols; j += blockDim.x) {
		Cell c; c.h = -(j + 1) * kGapExt - (hf.row_open ? kGapOpen : 0); c.x = -kInf;
		busH[hf.bus_off + j] = c;
	}
	for (int r = threadIdx.x; r <= hf.rows; r += blockDim.x) {
		Cell c;
		if (r == 0) { c.h = hf.corner; c.x = -kInf; }
		else { c.h = -r * kGapExt - (hf.col_open ? kGapOpen : 0); c.x = -kInf; }
		left[hf.left_off + r] = c;
	}
}

__global__ void s4_reverse_kernel(const unsigned char* src, unsigned char* dst, int n) {
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < n) dst[k] = src[n - 1 - k];
}

// event code of one candidate: 0 none, 1 match (H+H), 2 gapped match (F+F+open), 3 error
__device__ __forceinline__ int s4_event(int ah, int af, int bh, int bf, int diff) {
	const int sm = ah + bh, sg = af + bf + kGapOpen;
	if (sm == diff) return 1;
	if (sg == diff) return 2;
	if (sm > diff || sg > diff) return 3;
	return 0;
}

// One warp per partition.  r0[k] / r1[k] (k = 1..len1) are busH cells of the forward / reverse half, k = 0 is the
// first-column cell of the half's last row (sw_stage4.cpp:342-343).
__global__ void s4_match_kernel(const S4Part* parts, int nparts, const Cell* busF0, const Cell* busR0, const Cell* busF1,
                                const Cell* busR1, const Cell* left0, const Cell* left1, XPoint* out, int* error) {
	const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (w >= nparts) return;
	const S4Part pt = parts[w];
	const Cell* bf = pt.grp ? busF1 : busF0;
	const Cell* br = pt.grp ? busR1 : busR0;
	const Cell* lf = pt.grp ? left1 : left0;
	const int len1 = pt.len1, jmid1 = len1 - len1 / 2;
	auto r0 = [&](int k, int& h, int& f) {
		if (k == 0) { h = lf[pt.fwd_left + pt.imid0].h; f = h; }
		else { const Cell c = bf[pt.fwd_bus + k - 1]; h = c.h; f = c.x; }
	};
	auto r1 = [&](int k, int& h, int& f) {
		if (k == 0) { h = lf[pt.rev_left + pt.imid1].h; f = h; }
		else { const Cell c = br[pt.rev_bus + k - 1]; h = c.h; f = c.x; }
	};
	// candidates in the reference's order: for j = jmid1-1 .. len1-1: A(j) then B(j); index q = 2*(j-(jmid1-1)) + {0,1}
	const int nq = 2 * (len1 - (jmid1 - 1));
	for (int q0 = 0; q0 < nq; q0 += 32) {
		const int q = q0 + lane;
		int ev = 0, sc = 0;
		if (q < nq) {
			const int j = (jmid1 - 1) + (q >> 1);
			int ah, af, bh, bfv;
			if ((q & 1) == 0) { r0(j + 1, ah, af); r1(len1 - (j + 1), bh, bfv); }
			else { r0(len1 - (j + 1), ah, af); r1(j + 1, bh, bfv); }
			ev = s4_event(ah, af, bh, bfv, pt.diff);
			sc = ev == 1 ? ah : af;
		}
		const unsigned m = __ballot_sync(0xffffffffu, ev != 0);
		if (m) {
			const int first = __ffs(m) - 1;
			const int fev = __shfl_sync(0xffffffffu, ev, first);
			const int fsc = __shfl_sync(0xffffffffu, sc, first);
			if (lane == 0) {
				XPoint o;
				if (fev == 3) { atomicExch(error, 1 + pt.out_index); o.i = o.j = o.score = 0; o.type = -1; }
				else {
					const int qq = q0 + first, j = (jmid1 - 1) + (qq >> 1);
					const int cj = (qq & 1) == 0 ? (j + 1) : (len1 - (j + 1));
					const int ci = pt.imid0;
					const int type = fev == 1 ? 0 : 2;                     // TYPE_MATCH : TYPE_GAP_2
					if (!pt.transposed) { o.i = pt.i0 + ci; o.j = pt.j0 + cj; o.type = type; }
					else { o.i = pt.i0 + cj; o.j = pt.j0 + ci; o.type = type == 2 ? 1 : type; }   // inv_type
					o.score = fsc + pt.score_s;
				}
				out[pt.out_index] = o;
			}
			return;
		}
	}
	if (lane == 0) { atomicExch(error, -(1 + pt.out_index)); XPoint o; o.i = o.j = o.score = 0; o.type = -1; out[pt.out_index] = o; }   // "NOT FOUND"
}

}  // namespace b200
