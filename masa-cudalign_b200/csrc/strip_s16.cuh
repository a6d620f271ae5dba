// strip_s16.cuh -- packed s16x2 DPX strip kernel: the production path for A/C/G/T inputs.
//
// Two DP cells per 32-bit lane-register.  A warp carries 64 "virtual lanes": virtual lane v = 2*lane + half
// owns R consecutive rows of the strip and works on column (t - v) at step t, so the low half of a register
// is always one column ahead of the high half and both halves advance along the same anti-diagonal.  The
// (H,F) border of virtual lane v-1 reaches v through one warp shuffle plus one PRMT that rotates the halves.
//
// Per row PAIR (two cells) the inner loop issues (SW):
//     PRMT                 substitution scores of both cells: byte-select from the two column profile words
//     VIADDMNMX.S16x2      E  = max(E - 2, T_left)                    T = H - 5 is the stored form of H
//     VIADDMNMX.S16x2      x  = max(T_diag + (s+5), zero)             SW clamp folded into the diagonal term
//     VIADDMNMX.S16x2      F  = max(F - 2, T_up)
//     VIMNMX3.S16x2        H  = max(x, E, F)
//     VIADD.16x2           T  = H - 5                                 (issues down the other integer pipe)
//     VIMNMX.S16x2         running maximum for best-score tracking
// i.e. ~3.5 ALU-pipe slots per cell against ~7-8 for the int32 kernel (PRMT counts double: measured 32 vs
// 64 lanes/clk/SM, profiles/r01_pipe_rates.txt).
//
// Range: scores are kept relative to a per-warp base that is re-centred every 32 columns when the reference
// cell drifts by more than kRebase, so all live values stay inside s16.  Neighbouring DP cells differ by a
// bounded amount, so the spread over a warp's parallelogram (64*R rows x 64 columns) is < 13k for R <= 32.
// -INF border inputs (E/F of first row/column, pruned neighbours in SW) are clamped to kNeg and vanish after
// one cell exactly as -INF does in the reference; NW partitions whose H inputs are -INF take the int32 kernel.
// Results (bottom row, right column, best cell) are converted back to the reference's exact int32 values.
//
// Block pruning (SW stage 1; replaces isBlockPrunable/updatePruningWindow/clearPrunedBlocks,
// C/libmasa/pruning/AbstractBlockPruning.cpp:70-113, BlockPruningDiagonal.cpp:112-152, on the device):
// a strip alternates between COMPUTE segments and a SKIP mode, decided every 32 columns.  A 32-column block is
// skipped when  max(H entering it) + slack + min(rows left, columns left) < best score known so far  (strict,
// so not even a tie can be lost and the best cell stays exact); skipped cells are published as H = 0 (a valid SW
// lower bound), which costs one coalesced load, one warp reduction and one store instead of 32 wavefront steps.
#pragma once
#include "strip_common.cuh"
#include "strip_s32.cuh"

namespace b200 {

// unroll factor of the check-free wavefront loop: the loop-carried registers of one step do not map onto themselves,
// so every trip around the loop ends in a block of register moves; unrolling amortises them (2: 175, 4: 163.5,
// 8: 157 instructions per step, tools/sass_step_count.py)
#ifndef B200_STEP_UNROLL
#define B200_STEP_UNROLL 4
#endif
constexpr int kStepUnroll = B200_STEP_UNROLL;

constexpr int kR16 = 8;                 // diag-compat instance: 64*8  = 512-row strips (reference block height)
constexpr int kSH16 = 64 * kR16;
constexpr int kR16F = 16;               // whole-partition instance: 64*16 = 1024-row strips
constexpr int kSH16F = 64 * kR16F;

constexpr int kNeg = -30000;            // local-frame stand-in for -INF
constexpr int kRebase = 4096;           // re-centre when |reference cell| exceeds this
constexpr int kCand = 32;               // candidate-ring entries per warp (>= 32: one step can trigger every lane)
constexpr int kFiltMin = 128;            // FILT steps only above this best score: random DNA reaches H ~ 30-40, so the F-chain bound (+33) stays below it
constexpr int kPruneSlack = 66;         // growth possible while the 64-lane pipeline drains, plus rounding

__device__ __forceinline__ unsigned pack2(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }
__device__ __forceinline__ int lo16(unsigned v) { return (int)(short)(v & 0xffffu); }
__device__ __forceinline__ int hi16(unsigned v) { return (int)(short)(v >> 16); }
__device__ __forceinline__ int clamp16(int v) { return v < kNeg ? kNeg : (v > 32767 ? 32767 : v); }
__device__ __forceinline__ unsigned dup2(int v) { return pack2(v, v); }
__device__ __forceinline__ unsigned thr_pack(int thr, int base) {
	if (thr == INT_MIN) return 0x80008000u;
	const long long d = (long long)thr - base;
	const int v = d > 32767 ? 32767 : (d < -32768 ? -32768 : (int)d);
	return ((unsigned)v & 0xffffu) | ((unsigned)v << 16);
}
__device__ __forceinline__ int code_of(int c) { return (c >> 1) & 3; }     // A->0 C->1 T->2 G->3
__device__ __forceinline__ bool is_acgt(int c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
// profile word of a column: byte k = score(column, row code k) + 5.  A non-ACGT column byte (N, IUPAC codes, ...)
// equals no A/C/G/T row byte, so all four entries are the mismatch score -- the reference's raw byte compare
// (R/src/CUDAligner.cu:282) restricted to strips whose ROWS are pure A/C/G/T (JOB_S32 handles the others).
__device__ __forceinline__ unsigned profile_word(int c) {
	return is_acgt(c) ? (0x02020202u ^ (0x04u << (8 * code_of(c)))) : 0x02020202u;
}

// LUT: fetch the packed substitution scores of a row pair with one conflict-free LDS from a table replicated per
// lane (lut[dq*16+cq][lane], dq = column codes of the two halves, cq = row codes) instead of the PRMT byte-select:
// PRMT issues at half the VIADDMNMX rate (profiles/r01_pipe_rates.txt) and the kernel is bound by that pipe, the LSU
// is idle.  Needs both sequences to be pure A/C/G/T (the PRMT variant serves launches with N / IUPAC columns).
constexpr int kLutBytes = 256 * 32 * 4;

template <int R, bool SW, bool TRACK, bool LUT = false>
struct StripS16 {
	static constexpr int V = 64;
	static constexpr int SH = V * R;

	struct Smem {
		// per column of the current 32-column block: {H<<16, F<<16 of the top border in the local frame (lane 0 injects
		// them into its low half), column profile word (byte k = s(k, column base) + 5) or LUT row offset (code << 11)}:
		// one LDS.128 per step
		uint4 top[32];
		unsigned botH[32];   // packed H / F registers of the lane that owns the bottom row, indexed by the step inside the
		unsigned botF[32];   // block: step u completes column (block start + u - vo) of the bottom row
		unsigned cand[kCand][R + 2];   // deferred best-cell candidates: T[0..R) of a lane + {step, lane|halves<<8}
	};

	// all per-warp state of one job
	struct State {
		unsigned T[R], E[R], sel[R];
		unsigned tprev, botH, botF, pa, pb;
		unsigned Zp, thrp, blk;
		unsigned lut;                        // shared-memory address of lut[0][lane] (LUT variant)
		int base;
		int bs, bi, bj, thr, pub, ncand;
	};

	// FILT (check-free steps of a tracking kernel, once the best known score is >= kFiltMin): the per-row running
	// maximum (R/2 VIMNMX3 per step) is replaced by ONE conservative bound taken from the F chain.  Inside a lane-step
	// F(r) >= H(k) - 5 - 2(r-1-k) for every row k < r, so  max_k H(k) <= max(F(R-1) + 2R + 1, H(R-1)).  Only when that
	// bound reaches the threshold (rows just below a high-scoring cell) is the exact maximum taken from the T registers,
	// on the rare path.  The bound also feeds the pruning maximum, where an upper bound is all that is needed.
	template <bool PARTIAL, bool CHECK, bool FILT = false>
	__device__ __forceinline__ static void step(const StripParams& p, const JobCtx& cx, State& s, Smem& sm, int warp, int lane,
	                                            int t, int u, int nv_lo, int nv_hi, int vo, int ro, int c0, int c1) {
		const StripJob& jb = cx.jb;
		const unsigned M2 = dup2(-kGapExt), M5 = dup2(-kGapFirst);
		unsigned shH = __shfl_up_sync(0xffffffffu, s.botH, 1);
		unsigned shF = __shfl_up_sync(0xffffffffu, s.botF, 1);
		unsigned shP = __shfl_up_sync(0xffffffffu, s.pb, 1);
		if (lane == 0) { const uint4 tp = sm.top[u]; shH = tp.x; shF = tp.y; shP = tp.z; }   // one predicated LDS.128
		const unsigned upH = prmt(shH, s.botH, 0x5432);      // lo <- neighbour's hi, hi <- own lo
		const unsigned upF = prmt(shF, s.botF, 0x5432);
		s.pb = s.pa; s.pa = shP;
		const int col_lo = t - 2 * lane, col_hi = col_lo - 1;
		bool act = true;
		if (CHECK) act = (col_lo >= c0) && (col_hi < c1);           // at least one half inside the segment [c0, c1)
		bool trig = false;
		unsigned smax_out = 0x80008000u;
		unsigned keep = 0;                                            // halves to freeze (CHECK): 0xffff lo, 0xffff0000 hi
		if (act) {
			unsigned dT = s.tprev;
			unsigned tup = __vadd2(upH, M5);
			unsigned f = upF, h = 0, smax = dup2(-32768), oh = 0, of = 0;
			// during fill/drain one half may be outside [0, cols): its state must not move
			if (CHECK) {
				if (col_lo >= c1) keep |= 0x0000ffffu;
				if (col_hi < c0) keep |= 0xffff0000u;
			}
			s.tprev = CHECK ? ((tup & ~keep) | (s.tprev & keep)) : tup;
			unsigned lrow = 0;
			if (LUT) lrow = s.lut + s.pa + (s.pb << 2);                // table rows of this step's two column codes (pa, pb = code << 11)
#pragma unroll
			for (int r = 0; r < R; r++) {
				unsigned sc;
				if (LUT) sc = lds32_pure(lrow + s.sel[r]);
				else sc = prmt(s.pa, s.pb, s.sel[r]);
				const unsigned e = __viaddmax_s16x2(s.E[r], M2, s.T[r]);
				const unsigned x = SW ? __viaddmax_s16x2(dT, sc, s.Zp) : __vadd2(dT, sc);
				f = __viaddmax_s16x2(f, M2, tup);
				h = __vimax3_s16x2(x, e, f);
				dT = s.T[r];
				tup = __vadd2(h, M5);
				if (CHECK) { s.E[r] = (e & ~keep) | (s.E[r] & keep); s.T[r] = (tup & ~keep) | (s.T[r] & keep); }
				else { s.E[r] = e; s.T[r] = tup; }
				if (TRACK && !FILT) smax = __vmaxs2(smax, h);
				if (PARTIAL && r == ro) { oh = h; of = f; }
			}
			if (TRACK && FILT) smax = __viaddmax_s16x2(f, dup2(2 * R + 1), h);
			if (!PARTIAL) { oh = h; of = f; }
			if (CHECK) { s.botH = (h & ~keep) | (s.botH & keep); s.botF = (f & ~keep) | (s.botF & keep); }
			else { s.botH = h; s.botF = f; }

			if (lane == (vo >> 1)) {
				const int oc = (vo & 1) ? col_hi : col_lo;
				if (!CHECK || (oc >= c0 && oc < c1)) { sm.botH[u] = oh; sm.botF[u] = of; }   // halves are picked at flush time
			}
			if (TRACK) {
				if (CHECK) {
					// frozen halves carry stale values: they feed neither the trigger nor the pruning maximum
					if (keep & 0x0000ffffu) smax = (smax & 0xffff0000u) | 0x8000u;
					if (keep & 0xffff0000u) smax = (smax & 0x0000ffffu) | 0x80000000u;
				}
				bool plo, phi;
				(void)__vibmax_s16x2(smax, s.thrp, &phi, &plo);
				if (CHECK) { if (keep & 0x0000ffffu) plo = false; if (keep & 0xffff0000u) phi = false; }
				trig = plo || phi;
				smax_out = smax;
				s.blk = __vmaxs2(s.blk, smax);
			}
			// (the right border is stored after the last step of the strip, from the frozen registers: run_job)
		}
		if (TRACK) {
			// Rare path, deferred: a lane whose column pair ties/beats the best known so far parks its R packed
			// H registers in the warp's candidate ring (a handful of STS); the exact (score,i,j) bookkeeping is done
			// by all 32 lanes together in drain(), every 32 columns, so the strip that carries the alignment path
			// does not throttle the strips chained below it.
			unsigned mask = __ballot_sync(0xffffffffu, trig);
			if (FILT && mask) {
				// some lane's bound reached the threshold: exact maximum of this step's H values from the T registers
				unsigned tm = s.T[0];
#pragma unroll
				for (int r = 1; r + 1 < R; r += 2) tm = __vimax3_s16x2(tm, s.T[r], s.T[r + 1]);
				if ((R & 1) == 0) tm = __vmaxs2(tm, s.T[R - 1]);
				smax_out = __vadd2(tm, dup2(kGapFirst));
				bool plo, phi;
				(void)__vibmax_s16x2(smax_out, s.thrp, &phi, &plo);
				trig = trig && (plo || phi);
				mask = __ballot_sync(0xffffffffu, trig);
			}
			if (mask) {
				const int n = __popc(mask);
				if (s.ncand + n > kCand) drain(jb, s, sm, warp, lane, c0, c1);
				if (trig) {
					// which halves reached the threshold is recomputed here (opaque copy: no flag registers on the hot path)
					unsigned sx = smax_out;
					asm volatile("" : "+r"(sx));
					bool plo, phi;
					(void)__vibmax_s16x2(sx, s.thrp, &phi, &plo);
					if (CHECK) { if (keep & 0x0000ffffu) plo = false; if (keep & 0xffff0000u) phi = false; }
					const int slot = s.ncand + __popc(mask & ((1u << lane) - 1u));
					unsigned* e = sm.cand[slot];
#pragma unroll
					for (int r = 0; r < R; r++) e[r] = s.T[r];
					e[R] = (unsigned)t;
					e[R + 1] = (unsigned)lane | (plo ? 0x100u : 0u) | (phi ? 0x200u : 0u);
				}
				s.ncand += n;
			}
		}
	}

	struct Best { int bs, bi, bj; };

	// Cooperative scan of the candidate ring: lane l examines cell (half = l / R, row = l % R) of every entry.
	// Takes and returns scalars only, so the register-resident State never has its address taken.
	// `cand` is the 32-bit shared-memory address of the ring: a generic pointer argument would have its 64-bit address
	// materialised in the hot loop of the caller
	__device__ __noinline__ static Best drain_scan(unsigned cand, int ncand, int rows, int c0, int c1, int i0, int j0,
	                                               int base, int lane, Best b) {
		__syncwarp();
		const int half = lane / R, r = lane % R;
		for (int e = 0; e < ncand; e++) {
			const unsigned en = cand + (unsigned)e * (R + 2) * 4u;
			const unsigned meta = lds32(en + (R + 1) * 4u);
			const int te = (int)lds32(en + R * 4u), src = (int)(meta & 31u);
			if (half < 2 && ((meta >> (8 + half)) & 1u)) {
				const unsigned w = lds32(en + (unsigned)r * 4u);
				const int v = 2 * src + half, col = te - v, row = v * R + r;
				if (row < rows && col >= c0 && col < c1) {
					const int hv = (half ? hi16(w) : lo16(w)) + kGapFirst + base;
					const int i = i0 + row, j = j0 + col;
					if (better(hv, i, j, b.bs, b.bi, b.bj)) { b.bs = hv; b.bi = i; b.bj = j; }
				}
			}
		}
		__syncwarp();
		return b;
	}

	__device__ __forceinline__ static void drain(const StripJob& jb, State& s, Smem& sm, int warp, int lane, int c0, int c1) {
		Best b; b.bs = s.bs; b.bi = s.bi; b.bj = s.bj;
		b = drain_scan((unsigned)__cvta_generic_to_shared(&sm.cand[0][0]), s.ncand, jb.rows, c0, c1, jb.i0, jb.j0, s.base, lane, b);
		s.bs = b.bs; s.bi = b.bi; s.bj = b.bj;
		s.ncand = 0;
		const int wb = __reduce_max_sync(0xffffffffu, s.bs);
		if (wb > s.thr) { s.thr = wb; s.thrp = thr_pack(s.thr, s.base); }
	}

	// (re)start a compute segment whose left neighbour is all zeros (SW): H = 0, E = -INF.
	// The frame is taken from the TOP border the segment starts under (tmax = its largest H): in the lower part of a
	// large matrix the surviving band starts at cells worth millions (reached from the alignment path through long
	// gaps), and a frame anchored at zero would saturate the s16 lanes for tmax/32767 blocks while re-centring catches
	// up -- under-estimated cells that the strips below inherit, a front that creeps towards the alignment path by
	// (tmax/1024 - strip shift) columns per strip and finally erases it (seen on the 23M x 25M pair: best lost after row
	// 20.9M).  With the frame at the top border the zeros on the left become the frame's floor, like every SW zero once
	// base > 30000; they are dead cells by the pruning bound, far below every live value.
	__device__ __forceinline__ static void start_zero_segment(State& s, int nv_lo, int nv_hi, int tmax) {
		const int b = tmax > 16384 ? tmax - 8192 : 0;
		const int z = clamp16(-b), t0 = clamp16(-b - kGapFirst);
		s.base = b;
#pragma unroll
		for (int r = 0; r < R; r++) {
			s.T[r] = pack2(r < nv_lo ? t0 : kNeg, r < nv_hi ? t0 : kNeg);
			s.E[r] = dup2(kNeg);
		}
		s.tprev = dup2(t0);
		s.botH = dup2(z); s.botF = dup2(z); s.pa = LUT ? 0u : 0x02020202u; s.pb = s.pa;
		s.Zp = dup2(z);
		s.thrp = thr_pack(s.thr, b);
		s.blk = 0x80008000u;
	}

	template <bool PARTIAL, bool CHAIN>
	__device__ static void run_job(const StripParams& p, int job, Smem& sm, int warp, int lane, unsigned lut_addr = 0) {
		const JobCtx cx = fetch_job<CHAIN ? 1 : 0>(p, job);
		const StripJob& jb = cx.jb;
		const int rows = jb.rows, cols = jb.cols, i0 = jb.i0, j0 = jb.j0;
		const int rb_lo = (2 * lane) * R, rb_hi = (2 * lane + 1) * R;     // first row of each half inside the strip
		int nv_lo = rows - rb_lo; nv_lo = nv_lo < 0 ? 0 : (nv_lo > R ? R : nv_lo);
		int nv_hi = rows - rb_hi; nv_hi = nv_hi < 0 ? 0 : (nv_hi > R ? R : nv_hi);
		const int vo = (rows - 1) / R, ro = (rows - 1) % R;
		const bool lz = (jb.flags & JOB_LEFT_ZERO) != 0;
		const bool top_minf = (jb.flags & JOB_TOP_MINF) != 0;
		const bool prune = TRACK && SW && p.prune != 0;
		const int rows_left = p.prune_i1 - i0;                            // rows from the top of this strip to the end

		State s;
		s.lut = lut_addr + 4u * (unsigned)lane;
		keep_in_register(s.lut);                               // instead of re-deriving it every step
		// ---- left border; the frame starts at the H of the corner
		const Cell* lb = left_border(p, cx);
		int base = 0;
		int lmaxv = INT_MIN;                   // largest H of this lane's left-border cells (true scores)
		if (!lz) {
			int v0 = 0;
			if (lane == 0) { v0 = __ldcg(&lb[0].h); if (v0 < -kInf / 2) v0 = __ldcg(&lb[1].h); if (v0 < -kInf / 2) v0 = 0; }
			base = __shfl_sync(0xffffffffu, v0, 0);
		}
		s.base = base;
#pragma unroll
		for (int r = 0; r < R; r++) {
			int hl = -kGapFirst - base, el = kNeg, hh = -kGapFirst - base, eh = kNeg, cl = 0, chh = 0;
			if (r < nv_lo) {
				cl = (LUT && p.s0p) ? packed_code(p.s0p, i0 + rb_lo + r) : code_of(p.s0[i0 + rb_lo + r]);
				if (!lz) { Cell c = ldcg_cell(lb + 1 + rb_lo + r); lmaxv = max(lmaxv, c.h); hl = clamp16(c.h < -kInf / 2 ? kNeg : c.h - kGapFirst - base); el = clamp16(c.x < -kInf / 2 ? kNeg : c.x - base); }
			} else { hl = kNeg; }
			if (r < nv_hi) {
				chh = (LUT && p.s0p) ? packed_code(p.s0p, i0 + rb_hi + r) : code_of(p.s0[i0 + rb_hi + r]);
				if (!lz) { Cell c = ldcg_cell(lb + 1 + rb_hi + r); lmaxv = max(lmaxv, c.h); hh = clamp16(c.h < -kInf / 2 ? kNeg : c.h - kGapFirst - base); eh = clamp16(c.x < -kInf / 2 ? kNeg : c.x - base); }
			} else { hh = kNeg; }
			s.T[r] = pack2(clamp16(hl), clamp16(hh));
			s.E[r] = pack2(el, eh);
			if (LUT) s.sel[r] = (unsigned)(cl | (chh << 2)) << 7;       // byte offset of table row cq inside a dq block (32 lanes x 4 B)
			else s.sel[r] = (unsigned)cl | ((unsigned)(8 | cl) << 4) | ((unsigned)(4 + chh) << 8) | ((unsigned)(12 + chh) << 12);
		}
		{
			// diagonal term of row 0 of each half at its first column: H(row above, column -1) - 5
			int dl = -kGapFirst - base, dh = -kGapFirst - base;
			if (!lz) {
				if (rb_lo < rows) { int v = __ldcg(&lb[rb_lo].h); dl = clamp16(v < -kInf / 2 ? kNeg : v - kGapFirst - base); }
				if (rb_hi < rows) { int v = __ldcg(&lb[rb_hi].h); dh = clamp16(v < -kInf / 2 ? kNeg : v - kGapFirst - base); }
			}
			s.tprev = pack2(dl, dh);
		}
		s.botH = 0; s.botF = 0; s.pa = LUT ? 0u : 0x02020202u; s.pb = s.pa;
		s.Zp = dup2(clamp16(-base));
		s.bs = INT_MIN; s.bi = -1; s.bj = -1; s.thr = INT_MIN; s.pub = INT_MIN; s.thrp = 0x80008000u; s.ncand = 0;
		s.blk = 0x80008000u;

		int flushed = 0;                       // columns of the bottom row published so far (computed or skipped)
		int seen = jb.dep < 0 ? INT_MAX : 0;   // progress (in our columns) of the strip above observed by our last acquire (OPT_SEEN_CACHE)
		const int opt = p.opt;
		int pos = 0;                           // next column to decide in skip mode
		constexpr bool chained = CHAIN;              // chain mode: our first publication lets the strip below start
		// A job that starts from a real left border (custom first column, or the border delivered by the chunk on our left
		// in chain mode) must carry that border in the pruning test until every virtual lane has consumed its cells: the
		// running block maxima only know the lanes that have started (the alignment path may enter through the lower rows).
		int lpend = INT_MIN;
		bool left_dead = false;
		if (CHAIN && prune && !lz) {                // (the single-GPU instances prune only behind a zero first column: engine.cu)
			lpend = __reduce_max_sync(0xffffffffu, lmaxv);
			if (lpend < 0) lpend = 0;
			const int g = ld_uniform(p.global_best);
			if (g > s.thr) { s.thr = g; s.pub = g; s.thrp = thr_pack(s.thr, s.base); }
			const int cols_left0 = p.prune_j1 - j0;
			left_dead = s.thr != INT_MIN && (long long)lpend + kPruneSlack + (rows_left < cols_left0 ? rows_left : cols_left0) < (long long)s.thr;
		}
		bool computing = !(prune && (lz || left_dead));   // a zero (or dead) left border lets the strip start in skip mode
		long long computed_cols = 0;
		int bm1 = INT_MIN, bm2 = INT_MIN;      // maxima (true scores) of the last two computed 32-step blocks

		for (;;) {
			if (!computing) {
				// =========================== SKIP mode: one 32-column block per iteration ===========================
				if (pos >= cols) break;
				const int need = pos + 32 < cols ? pos + 32 : cols;
				if (!(opt & OPT_SEEN_CACHE)) wait_progress(p, jb.dep, cx.prog_base + need, lane);
				else if (seen < need) {
					seen = wait_progress_v(p, jb.dep, cx.prog_base + need, lane) - cx.prog_base;
					if (seen < need) return;           // the kernel is stopping (watchdog): do not sweep the rest of the strip
				}
				if (p.track == 2 && (!(opt & OPT_BEST_EVERY_4) || (pos & 96) == 0)) {
					const int g = ld_uniform(p.global_best);
					if (g > s.thr) { s.thr = g; s.pub = g; }
				}
				if ((opt & OPT_SKIP_128) && s.thr != INT_MIN && pos + 128 <= cols && seen >= pos + 128) {
					// the strip above is at least 128 columns ahead: decide four blocks with one reduction, one burst of
					// stores and one release (bound of the first block: the largest distance term, so never less strict)
					int th = 0;
					if (!top_minf) {
						const Cell* tp = p.busH + j0 + pos + lane;
						const int h0 = __ldcg(&tp[0].h), h1 = __ldcg(&tp[32].h), h2 = __ldcg(&tp[64].h), h3 = __ldcg(&tp[96].h);
						th = max(max(h0, h1), max(h2, h3));
					}
					int tmax = __reduce_max_sync(0xffffffffu, th);
					if (tmax < 0) tmax = 0;
					const int cols_left = p.prune_j1 - (j0 + pos);
					const long long bound = (long long)tmax + kPruneSlack + (rows_left < cols_left ? rows_left : cols_left);
					if (bound < (long long)s.thr) {
#pragma unroll
						for (int k = 0; k < 4; k++) {
							stcg_cell(p.busH + j0 + pos + lane + 32 * k, 0, -kInf);
							if (jb.sra_off >= 0) stcg_cell(p.sra + jb.sra_off + pos + lane + 32 * k, 0, -kInf);
						}
						const bool first = chained && flushed == 0;
						pos += 128; flushed = pos;
						__syncwarp();
						if (lane == 0) {
							if (!(opt & OPT_NO_SC_FENCE)) __threadfence();
							st_release(p.progress + cx.pidx, cx.prog_base + flushed);
							if (first) chain_notify_below(p, job);
						}
						continue;
					}
				}
				const int c = pos + lane;
				int th = 0;
				if (c < need && !top_minf) th = __ldcg(&p.busH[j0 + c].h);
				int tmax = __reduce_max_sync(0xffffffffu, th);
				if (tmax < 0) tmax = 0;
				const int cols_left = p.prune_j1 - (j0 + pos);
				const long long bound = (long long)tmax + kPruneSlack + (rows_left < cols_left ? rows_left : cols_left);
				if (s.thr != INT_MIN && bound < (long long)s.thr) {
					if (c < need) {
						stcg_cell(p.busH + j0 + c, 0, -kInf);
						if (jb.sra_off >= 0) stcg_cell(p.sra + jb.sra_off + c, 0, -kInf);
					}
					const bool first = chained && flushed == 0;
					pos = need; flushed = need;
					__syncwarp();
					if (lane == 0) {
						if (!(opt & OPT_NO_SC_FENCE)) __threadfence();
						st_release(p.progress + cx.pidx, cx.prog_base + flushed);
						if (first) chain_notify_below(p, job);
					}
					continue;
				}
				start_zero_segment(s, nv_lo, nv_hi, tmax);
				computing = true;
				bm1 = bm2 = INT_MIN;
				lpend = INT_MIN;                   // the left neighbour of a restarted segment is the zero border
			}

			// =========================== COMPUTE mode: one segment [c0, c1) ===========================
			const int c0 = pos;
			int c1 = cols;
			unsigned long long seg_t0 = 0;                     // statistics (chain instances): time this warp spends in compute segments
			if (CHAIN) seg_t0 = global_ns();
			if (CHAIN && p.sm_load != nullptr && lane == 0) { atomicAdd(p.sm_load, 1); atomicAdd(p.sm_load + 1 + sched_slot(), 1); }
#pragma unroll 1
			for (int tb = c0; tb < c1 + V - 1; tb += 32) {
				// ---- re-centre the frame on H(row 0 of the strip, last column done by virtual lane 0)
				if (tb > c0 && tb <= c1) {
					int ref = lo16(s.T[0]) + kGapFirst;
					ref = __shfl_sync(0xffffffffu, ref, 0);
					if (ref > kRebase || ref < -kRebase) {
						const unsigned d = dup2(-ref), fl = dup2(kNeg + (ref > 0 ? ref : 0));
#pragma unroll
						for (int r = 0; r < R; r++) {
							s.T[r] = __vadd2(__vmaxs2(s.T[r], fl), d);
							s.E[r] = __vadd2(__vmaxs2(s.E[r], fl), d);
						}
						s.tprev = __vadd2(__vmaxs2(s.tprev, fl), d);
						s.botH = __vadd2(__vmaxs2(s.botH, fl), d);
						s.botF = __vadd2(__vmaxs2(s.botF, fl), d);
						s.base += ref;
						s.Zp = dup2(clamp16(-s.base));
						s.thrp = thr_pack(s.thr, s.base);
					}
				}
				// ---- stage the next 32 columns of top border and seq1 (coalesced), gated on the strip above
				if (tb < c1) {
					const int need = tb + 32 < cols ? tb + 32 : cols;
					if (!(opt & OPT_SEEN_CACHE)) wait_progress(p, jb.dep, cx.prog_base + need, lane);
					else if (seen < need) {
						// Look-ahead (OPT_LOOKAHEAD): a strip that has caught up with the one above would find every block
						// "just not ready" and pay the detection latency of a spin-wait per block; waiting for two more blocks
						// instead lets it run the next ones without touching the counter (same pace, a third of the waits).
						int want = need;
						if ((opt & OPT_LOOKAHEAD) && tb >= c0 + 256) want = need + 64 < cols ? need + 64 : cols;
						seen = wait_progress_v(p, jb.dep, cx.prog_base + want, lane) - cx.prog_base;
						if (seen < need) return;       // the kernel is stopping (watchdog)
					}
					if (TRACK && p.track == 2 && (!(opt & OPT_BEST_EVERY_4) || (tb & 96) == 0 || tb == c0)) {
						// share the running best: publish ours, adopt a higher one (monotone, staleness is harmless)
						if (s.thr > s.pub) { if (lane == 0) push_best(p, s.thr); s.pub = s.thr; }
						const int g = ld_uniform(p.global_best);
						if (g > s.thr) { s.thr = g; s.pub = g; s.thrp = thr_pack(s.thr, s.base); }
					}
					const int c = tb + lane;
					Cell tv; tv.h = -kInf; tv.x = -kInf;
					if (c < cols && !top_minf) tv = ldcg_cell(p.busH + j0 + c);
					if (c == cols - 1 && jb.right_off >= 0) __stcg(&right_border(p, cx)[0].h, tv.h);   // corner cell of the block on our right
					if (prune && tb > c0) {
						// stop the segment here if nothing entering [tb, ...) can still reach the best known score
						int tmax = __reduce_max_sync(0xffffffffu, c < cols ? tv.h : 0);
						if (tmax < 0) tmax = 0;
						int in_max = tmax > bm1 ? tmax : bm1;
						in_max = in_max > bm2 ? in_max : bm2;
						const int cols_left = p.prune_j1 - (j0 + tb);
						const long long bound = (long long)in_max + kPruneSlack + (rows_left < cols_left ? rows_left : cols_left);
						bool stop = s.thr != INT_MIN && bm1 != INT_MIN && bound < (long long)s.thr;
						if (stop && lpend != INT_MIN && tb < c0 + 2 * V) {
							// lanes that have not started yet still hold left-border cells: those enter at column c0
							const int cl0 = p.prune_j1 - (j0 + c0);
							stop = (long long)lpend + kPruneSlack + (rows_left < cl0 ? rows_left : cl0) < (long long)s.thr;
						}
						if (stop) c1 = tb;
					}
					if (tb < c1) {
						int th = kNeg, tf = kNeg; unsigned pw = LUT ? 0u : 0x02020202u;
						if (c < cols) {
							th = tv.h < -kInf / 2 ? kNeg : clamp16(tv.h - s.base);
							tf = tv.x < -kInf / 2 ? kNeg : clamp16(tv.x - s.base);
							if (LUT) pw = (unsigned)(p.s1p ? packed_code(p.s1p, j0 + c) : code_of(p.s1[j0 + c])) << 11;
							else pw = profile_word(p.s1[j0 + c]);   // LUT row offset of the column code, or profile word: byte k = 6 (match+5) / 2
						}
						sm.top[lane] = make_uint4((unsigned)th << 16, (unsigned)tf << 16, pw, 0u);
						__syncwarp();
					}
				}

				// ---- 32 steps; the check-free body runs whenever every virtual lane is inside [c0, c1)
				const bool steady = (tb >= c0 + V) && (tb + 32 < c1);
				if (steady && TRACK && s.thr >= kFiltMin) {
#pragma unroll kStepUnroll
					for (int u = 0; u < 32; u++) step<PARTIAL, false, true>(p, cx, s, sm, warp, lane, tb + u, u, nv_lo, nv_hi, vo, ro, c0, c1);
				} else if (steady) {
#pragma unroll kStepUnroll
					for (int u = 0; u < 32; u++) step<PARTIAL, false>(p, cx, s, sm, warp, lane, tb + u, u, nv_lo, nv_hi, vo, ro, c0, c1);
				} else {
#pragma unroll 1
					for (int u = 0; u < 32; u++) step<PARTIAL, true>(p, cx, s, sm, warp, lane, tb + u, u, nv_lo, nv_hi, vo, ro, c0, c1);
				}
				if (TRACK) {
					if (s.ncand > 0) drain(jb, s, sm, warp, lane, c0, c1);
					if (prune) {
						const int lm = lo16(s.blk) > hi16(s.blk) ? lo16(s.blk) : hi16(s.blk);
						const int bm = __reduce_max_sync(0xffffffffu, lm);
						bm2 = bm1;
						bm1 = bm <= -32768 ? 0 : bm + s.base;
						s.blk = 0x80008000u;
					}
				}

				// ---- publish the columns of the bottom row completed during these 32 steps (exact int32 values)
				int cdone = tb + 31 - vo; cdone = cdone < c1 - 1 ? cdone : c1 - 1;
				if (cdone >= flushed) {
					__syncwarp();
					// Deferred release (OPT_DEFER_RELEASE): in the steady part of a segment the counter published here is the one
					// of the PREVIOUS block -- its stores landed 32 steps ago, so the fence inside st.release has nothing to wait
					// for (a release right behind the stores stalls the warp for an L2 round trip, every block, on the critical
					// path of every strip chained below).  The ramp-up (first 256 columns) and the end of a segment publish at once.
					const bool now = !(opt & OPT_DEFER_RELEASE) || cdone < 256 || cdone + 1 == cols || tb + 32 >= c1;
					if (!now && lane == 0) st_release(p.progress + cx.pidx, cx.prog_base + flushed);
					for (int c = flushed + lane; c <= cdone; c += 32) {
						const int k = c - tb + vo;                 // step of this block that completed column c of the bottom row
						const unsigned px = sm.botH[k], py = sm.botF[k];
						const int vx = (vo & 1) ? hi16(px) : lo16(px), vy = (vo & 1) ? hi16(py) : lo16(py);
						const int hv = vx <= kNeg ? -kInf : vx + s.base, fv = vy <= kNeg ? -kInf : vy + s.base;
						stcg_cell(p.busH + j0 + c, hv, fv);
						if (jb.sra_off >= 0) stcg_cell(p.sra + jb.sra_off + c, hv, fv);
					}
					// the strips below start as soon as the first columns are out; once the chain is filled the counter is
					// released every 128 columns (a release costs a fence; the consumers run several blocks behind anyway)
					const bool rel = now && (!(opt & OPT_RELEASE_128) || cdone + 1 == cols || cdone < 2048 || tb + 32 >= c1 ||
					                         ((cdone + 1) >> 7) != (flushed >> 7));
					const bool first = chained && flushed == 0;      // (the first release is never deferred: cdone < 256)
					flushed = cdone + 1;
					__syncwarp();
					if (rel && lane == 0) {
						if (!(opt & OPT_NO_SC_FENCE)) __threadfence();
						st_release(p.progress + cx.pidx, cx.prog_base + flushed);
						if (first) chain_notify_below(p, job);
					}
				}
			}
			computed_cols += c1 - c0;
			if (CHAIN && lane == 0) {
				atomicAdd(p.cells_done + 1, global_ns() - seg_t0);
				if (p.sm_load != nullptr) { atomicSub(p.sm_load, 1); atomicSub(p.sm_load + 1 + sched_slot(), 1); }
			}
			if (c1 >= cols) {
				// the segment ran to the last column: every half froze when it passed it, so the registers hold column cols-1
				if (jb.right_off >= 0) {
					Cell* rb = right_border(p, cx);
#pragma unroll
					for (int r = 0; r < R; r++) {
						if (r < nv_lo) stcg_cell(rb + 1 + rb_lo + r, lo16(s.T[r]) + kGapFirst + s.base, lo16(s.E[r]) + s.base);
						if (r < nv_hi) stcg_cell(rb + 1 + rb_hi + r, hi16(s.T[r]) + kGapFirst + s.base, hi16(s.E[r]) + s.base);
					}
				}
				break;
			}
			pos = c1;                          // the segment was cut short: continue in skip mode
			computing = false;
			lpend = INT_MIN;
		}

		if (prune && !computing && jb.right_off >= 0) {
			// the strip ended in skip mode: its right border is the zero border
			Cell* rb = right_border(p, cx);
			for (int k = lane; k < rows; k += 32) stcg_cell(rb + 1 + k, 0, -kInf);
			if (lane == 0) __stcg(&rb[0].h, 0);
		}

		if (TRACK) {
			int bs = s.bs, bi = s.bi, bj = s.bj;
#pragma unroll
			for (int d = 16; d >= 1; d >>= 1) {
				const int os = __shfl_xor_sync(0xffffffffu, bs, d);
				const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
				const int oj = __shfl_xor_sync(0xffffffffu, bj, d);
				if (better(os, oi, oj, bs, bi, bj)) { bs = os; bi = oi; bj = oj; }
			}
			if (lane == 0) {
				store_result<CHAIN ? 1 : 0>(p, cx, bs, bi, bj);
				if (bs != INT_MIN) push_best(p, bs);
			}
		}
		if (CHAIN) {
			// the right border is in the next GPU's memory and our best is folded into the strip's result: hand the strip over
			__syncwarp();
			if (lane == 0) chain_notify_right(p, job);
		}
		signal_special_row(p, jb, lane);
		if (lane == 0) atomicAdd(p.cells_done, (unsigned long long)rows * (unsigned long long)computed_cols);
	}
};

// MIXED: the launch also contains JOB_S32 strips (rows with N / IUPAC bytes) and columns may hold such bytes: PRMT
// variant plus the int32 path.  Pure A/C/G/T launches of the whole-partition instance (R = kR16F) use the LUT variant.
template <int R, bool SW, bool TRACK, bool MIXED = false, bool CHAIN = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4) strip_kernel_s16(const StripParams p) {
	constexpr bool LUT = !MIXED;
	using K = StripS16<R, SW, TRACK, LUT>;
	using K32 = StripS32<16, SW, TRACK>;
	__shared__ union { typename K::Smem s16; typename K32::Smem s32; } smu[kWarpsPerBlock];   // per warp: warps run different job kinds
	__shared__ unsigned lut[LUT ? 256 * 32 : 1];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	typename K::Smem& sm = smu[warp].s16;
	unsigned lut_addr = 0;
	if (LUT) {
		// lut[dq*16 + cq][lane]: scores (+5) of row codes cq = c_lo | c_hi<<2 against column codes dq = d_lo | d_hi<<2
		for (int idx = threadIdx.x; idx < 256 * 32; idx += blockDim.x) {
			const int e = idx >> 5, cq = e & 15, dq = e >> 4;
			const int lo = ((cq & 3) == (dq & 3)) ? kMatch + kGapFirst : kMismatch + kGapFirst;
			const int hi = ((cq >> 2) == (dq >> 2)) ? kMatch + kGapFirst : kMismatch + kGapFirst;
			lut[idx] = (unsigned)lo | ((unsigned)hi << 16);
		}
		__syncthreads();
		lut_addr = (unsigned)__cvta_generic_to_shared(lut);
	}
	for (;;) {
		const int job = claim_job<CHAIN ? 1 : 0>(p, lane);
		if (job < 0) break;
		if (ld_uniform(p.stop_flag)) break;
		int flags, rows;
		if (CHAIN) { const StripRow& sr = p.chain.strips[job % p.chain.nstrips]; flags = sr.flags; rows = sr.rows; }
		else { flags = p.jobs[job].flags; rows = p.jobs[job].rows; }
		if (flags & JOB_PRUNED) {
			// diag path only. Same semantics as the int32 kernel: -INF to the right border, no score (CUDAligner.cu:950-960)
			const StripJob jb = p.jobs[job];
			if (jb.right_off >= 0)
				for (int k = lane; k <= jb.rows; k += 32) stcg_cell(p.right + jb.right_off + k, -kInf, -kInf);
			if (TRACK && lane == 0) { Score3 o; o.score = -kInf; o.i = -1; o.j = -1; o.pad = 0; p.results[job] = o; }
			__syncwarp();
			if (lane == 0) { __threadfence(); st_release(p.progress + job, jb.cols); }
			continue;
		}
		if (MIXED && (flags & JOB_S32)) K32::run_job(p, job, smu[warp].s32, warp, lane);      // rows with N / IUPAC bytes: exact int32 path
		else if (rows < K::SH) K::template run_job<true, CHAIN>(p, job, sm, warp, lane, lut_addr);
		else K::template run_job<false, CHAIN>(p, job, sm, warp, lane, lut_addr);
	}
}

}  // namespace b200
