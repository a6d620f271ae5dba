// strip_s32.cuh -- exact int32 strip kernel (any alphabet, any border values, SW and NW).
//
// Replaces kernel_sw/kernel_sw4/kernel_check_max4/kernel_flush and the short/long/single phase kernels of
// R/src/CUDAligner.cu:276-1156.  Arithmetic is the reference recurrence restated for the DPX instructions:
//     E(i,j) = max(E(i,j-1) - 2, H(i,j-1) - 5)      == max(H-3, E) - 2   (CPUBlockProcessor.cpp:68,89)
//     F(i,j) = max(F(i-1,j) - 2, H(i-1,j) - 5)
//     H(i,j) = max(H(i-1,j-1) + s, E, F [, 0])
// in plain wrapping int32, so -INF (= -999999999) drifts exactly like it does in the reference.
// Register state per lane: T[r] = H(row r, previous column) - 5 and E[r] for its R rows; one VIADDMNMX each
// for E, F and the diagonal term, one VIMNMX for H, one add for T, ISETP+SEL for the substitution score.
#pragma once
#include "strip_common.cuh"

namespace b200 {

constexpr int kWarpsPerBlock = 4;   // strip-warps per CTA: they only share the 32 KB substitution table of the packed kernel

template <int R, bool SW, bool TRACK>
struct StripS32 {
	static constexpr int V = 32;          // virtual lanes per warp
	static constexpr int SH = V * R;      // strip height

	struct Smem {
		Cell top[32];
		Cell bot[64];
		unsigned char seq[32];
	};

	__device__ static void run_job(const StripParams& p, int job, Smem& sm, int warp, int lane) {
		const JobCtx cx = fetch_job<-1>(p, job);
		const StripJob& jb = cx.jb;
		const int rows = jb.rows, cols = jb.cols, i0 = jb.i0, j0 = jb.j0;

		if (jb.flags & JOB_PRUNED) {             // diag path only (never in chain mode)
			if (jb.right_off >= 0)
				for (int k = lane; k <= rows; k += 32) stcg_cell(right_border(p, cx) + k, -kInf, -kInf);
			if (TRACK && lane == 0) { Score3 s; s.score = -kInf; s.i = -1; s.j = -1; s.pad = 0; p.results[cx.pidx] = s; }
			__threadfence(); __syncwarp();
			if (lane == 0) st_release(p.progress + cx.pidx, cx.prog_base + cols);
			return;
		}

		const int row_base = lane * R;                       // first row of this lane inside the strip
		int nvalid = rows - row_base; nvalid = nvalid < 0 ? 0 : (nvalid > R ? R : nvalid);
		const int vo = (rows - 1) / R, ro = (rows - 1) % R;  // lane / register that hold the strip's bottom row
		const bool partial = rows < SH;

		int c0[R], T[R], E[R];
#pragma unroll
		for (int r = 0; r < R; r++) c0[r] = (r < nvalid) ? (int)p.s0[i0 + row_base + r] : 0x100;   // 0x100 never equals a byte

		int tprev;   // H(row above this lane, previous column) - 5: diagonal term of row 0
		if (jb.flags & JOB_LEFT_ZERO) {
#pragma unroll
			for (int r = 0; r < R; r++) { T[r] = 0 - kGapFirst; E[r] = -kInf; }
			tprev = 0 - kGapFirst;
		} else {
			const Cell* lb = left_border(p, cx);
#pragma unroll
			for (int r = 0; r < R; r++) {
				if (r < nvalid) { Cell c = ldcg_cell(lb + 1 + row_base + r); T[r] = c.h - kGapFirst; E[r] = c.x; }
				else { T[r] = -kInf; E[r] = -kInf; }
			}
			tprev = (row_base < rows) ? __ldcg(&lb[row_base].h) - kGapFirst : -kInf;
		}

		int botH = 0, botF = 0, ccur = 0x200;
		int bs = INT_MIN, bi = -1, bj = -1;      // exact best candidate of this lane
		int thr = INT_MIN, pub = INT_MIN;         // warp-uniform: enter the rare path when a cell reaches thr
		int flushed = 0;                          // columns of the bottom row already published
		const int total = cols + V - 1;
		const bool top_minf = (jb.flags & JOB_TOP_MINF) != 0;

#pragma unroll 1
		for (int tb = 0; tb < total; tb += 32) {
			// ---- stage the next 32 columns of top border and seq1 (coalesced), gated on the strip above
			if (tb < cols) {
				int need = tb + 32 < cols ? tb + 32 : cols;
				if (jb.dep >= 0 && wait_progress_v(p, jb.dep, cx.prog_base + need, lane) < cx.prog_base + need) return;   // stopping (watchdog)
				int c = tb + lane;
				Cell tv; tv.h = -kInf; tv.x = -kInf; unsigned char ch = 0;
				if (c < cols) {
					if (!top_minf) tv = ldcg_cell(p.busH + j0 + c);
					ch = p.s1[j0 + c];
				}
				sm.top[lane] = tv;
				sm.seq[lane] = ch;
				if (TRACK && p.track == 2) {
					if (thr > pub) { if (lane == 0) push_best(p, thr); pub = thr; }
					const int g = ld_uniform(p.global_best);
					if (g > thr) { thr = g; pub = g; }
				}
				__syncwarp();
			}

#pragma unroll 1
			for (int u = 0; u < 32; u++) {
				const int t = tb + u;
				int upH = __shfl_up_sync(0xffffffffu, botH, 1);
				int upF = __shfl_up_sync(0xffffffffu, botF, 1);
				int cc = __shfl_up_sync(0xffffffffu, ccur, 1);
				const Cell tv = sm.top[u];
				const int tc = sm.seq[u];
				if (lane == 0) { upH = tv.h; upF = tv.x; cc = tc; }
				ccur = cc;
				const int col = t - lane;
				bool trig = false;
				if ((unsigned)col < (unsigned)cols) {
					int dT = tprev;
					int tup = upH - kGapFirst;
					tprev = tup;
					int f = upF;
					int h = 0, smax = INT_MIN, oh = 0, of = 0;
#pragma unroll
					for (int r = 0; r < R; r++) {
						const int s5 = (c0[r] == cc) ? (kMatch + kGapFirst) : (kMismatch + kGapFirst);
						E[r] = __viaddmax_s32(E[r], -kGapExt, T[r]);
						const int a = SW ? __viaddmax_s32_relu(dT, s5, E[r]) : __viaddmax_s32(dT, s5, E[r]);
						f = __viaddmax_s32(f, -kGapExt, tup);
						h = max(a, f);
						dT = T[r];
						tup = h - kGapFirst;
						T[r] = tup;
						if (TRACK) { if (!partial || r < nvalid) smax = max(smax, h); }
						if (partial && r == ro) { oh = h; of = f; }
					}
					if (!partial) { oh = h; of = f; }
					botH = h; botF = f;
					if (lane == vo) { Cell o; o.h = oh; o.x = of; sm.bot[col & 63] = o; }
					if (TRACK) trig = smax >= thr;
					if (col == cols - 1 && jb.right_off >= 0) {
						Cell* rb = right_border(p, cx);
#pragma unroll
						for (int r = 0; r < R; r++)
							if (r < nvalid) stcg_cell(rb + 1 + row_base + r, T[r] + kGapFirst, E[r]);
						if (lane == 0) __stcg(&rb[0].h, upH);     // corner for the block on our right
					}
				}
				if (TRACK && __any_sync(0xffffffffu, trig)) {
					// rare path (warp-uniform): some cell of this anti-diagonal ties or beats the best known so far
					if (trig) {
						const int j = j0 + col;
#pragma unroll
						for (int r = 0; r < R; r++) {
							if (r < nvalid) {
								const int hv = T[r] + kGapFirst, i = i0 + row_base + r;
								if (better(hv, i, j, bs, bi, bj)) { bs = hv; bi = i; bj = j; }
							}
						}
					}
					const int wb = __reduce_max_sync(0xffffffffu, bs);
					thr = thr > wb ? thr : wb;
				}
			}

			// ---- publish the columns of the bottom row completed during these 32 steps
			int cdone = tb + 31 - vo; cdone = cdone < cols - 1 ? cdone : cols - 1;
			if (cdone >= flushed) {
				__syncwarp();
				for (int c = flushed + lane; c <= cdone; c += 32) {
					const Cell v = sm.bot[c & 63];
					stcg_cell(p.busH + j0 + c, v.h, v.x);
					if (jb.sra_off >= 0) stcg_cell(p.sra + jb.sra_off + c, v.h, v.x);
				}
				const bool first = flushed == 0 && p.chain.enabled;      // chain mode: our first publication lets the strip below start
				flushed = cdone + 1;
				__threadfence();
				__syncwarp();
				if (lane == 0) {
					st_release(p.progress + cx.pidx, cx.prog_base + flushed);
					if (first) chain_notify_below(p, job);
				}
			}
		}

		if (TRACK) {
#pragma unroll
			for (int d = 16; d >= 1; d >>= 1) {
				const int os = __shfl_xor_sync(0xffffffffu, bs, d);
				const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
				const int oj = __shfl_xor_sync(0xffffffffu, bj, d);
				if (better(os, oi, oj, bs, bi, bj)) { bs = os; bi = oi; bj = oj; }
			}
			if (lane == 0) {
				store_result<-1>(p, cx, bs, bi, bj);
				if (bs != INT_MIN) push_best(p, bs);
			}
		}
		if (p.chain.enabled) {
			// the right border is in the next GPU's memory and our result is folded into the strip's: hand the strip over
			__syncwarp();
			if (lane == 0) chain_notify_right(p, job);
		}
		signal_special_row(p, jb, lane);
		if (lane == 0) atomicAdd(p.cells_done, (unsigned long long)rows * (unsigned long long)cols);
	}
};

template <int R, bool SW, bool TRACK>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) strip_kernel_s32(const StripParams p) {
	using K = StripS32<R, SW, TRACK>;
	__shared__ typename K::Smem smw[kWarpsPerBlock];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	typename K::Smem& sm = smw[warp];
	for (;;) {
		const int job = claim_job<-1>(p, lane);
		if (job < 0) break;
		if (ld_uniform(p.stop_flag)) break;
		K::run_job(p, job, sm, warp, lane);
	}
}

}  // namespace b200
