// strip_common.cuh -- shared definitions of the strip ("chained wavefront") kernels.
//
// Work decomposition (B200-first; replaces the grid/external-diagonal scheme of R/src/CUDAligner.cu:745-1156):
//   * the partition is cut into horizontal STRIPS of at most SH rows; one WARP owns one strip at a time and
//     sweeps it left to right, lanes skewed by one column (anti-diagonal wavefront inside the warp);
//   * every lane keeps R rows of the strip in registers; neighbouring lanes exchange the (H,F) border with
//     warp shuffles; no shared-memory traffic or barrier in the cell loop;
//   * the strip's bottom row goes to the horizontal bus busH (int32 (H,F) per column, L2 resident) in
//     coalesced 32-column bursts, followed by a release-store of the strip's progress counter; the warp of
//     the strip below spins on that counter (acquire) before it loads the same 32 columns as its top border:
//     strips are chained through L2, never through the host;
//   * strips are claimed from an atomic counter in row order, so a waiting warp always waits on a strip that
//     is owned by a resident, running warp: no deadlock as long as the grid fits on the GPU (the launcher
//     sizes it from the occupancy API).
// The same kernels run the "diag" compatibility path: there a job is one block of the reference's grid and
// jobs of one launch do not depend on each other.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr int kInf = 999999999;                // C/libmasa/libmasaTypes.hpp:46
constexpr int kMatch = 1, kMismatch = -3;      // R/src/CUDAligner.hpp:77-98
constexpr int kGapOpen = 3, kGapExt = 2, kGapFirst = kGapOpen + kGapExt;

struct Cell { int h; int x; };                 // == cell_t (x = F in rows, E in columns)
struct Score3 { int score; int i; int j; int pad; };

enum JobFlags : int {
	JOB_LEFT_ZERO = 1,      // left border is the constant (H=0, E=-INF): SW stage 1 first column
	JOB_PRUNED    = 2,      // do not compute: write -INF to the right border (CUDAligner.cu:950-960 semantics)
	JOB_TOP_MINF  = 4,      // treat the top border as -INF instead of reading busH (strip above was pruned)
	JOB_S32       = 8,      // rows of this strip contain a non-ACGT byte: run it with the exact int32 code path
};

struct StripJob {
	int i0, rows;           // first row (index into seq0) and number of rows (1..SH)
	int j0, cols;           // first column (index into seq1) and number of columns (>= 1)
	int dep;                // job whose progress gates our top border, or -1
	int flags;              // JobFlags
	int left_off;           // cell index of slot 0 (corner) of our left border in StripParams::left; slots 1..rows follow
	int right_off;          // cell index of slot 0 of our right border in StripParams::right, or -1
	long long sra_off;      // cell index in StripParams::sra where column j0 of our bottom row goes, or -1
	long long sra_index;    // which special row this strip's bottom row is (index into StripParams::sra_done)
};

struct StripParams {
	const unsigned char* s0;
	const unsigned char* s1;
	Cell* busH;             // [seq1_len] (H,F) of the row above the strip being read / bottom row being written
	const Cell* left;       // left borders  (H,E)
	Cell* right;            // right borders (H,E)
	Cell* sra;              // on-device special-rows area
	const StripJob* jobs;
	int njobs;
	int* job_counter;       // atomic claim counter (zeroed before launch)
	int* progress;          // [njobs] columns of the bottom row published so far
	Score3* results;        // [njobs] best cell of each job (track != 0)
	int* global_best;       // running best score of the whole partition (atomicMax)
	unsigned long long* cells_done;   // statistics
	int* stop_flag;         // non-zero asks the kernel to stop (set by a spin-wait watchdog: no hung GPU on a protocol bug)
	const int* left_ready;  // multi-GPU: rows of our left border published by the previous GPU (system scope), or NULL
	int* right_ready;       // multi-GPU: row counter in the NEXT GPU's exchange block (peer memory), or NULL
	int* peer_best[8];      // multi-GPU: running-best words of the other GPUs (peer memory)
	int n_peer_best;
	int* sra_done;          // host-mapped flags, one per special row: set when the row is complete in the device SRA, or NULL
	int recurrence;         // B200_SMITH_WATERMAN | B200_NEEDLEMAN_WUNSCH
	int track;              // 0: no best tracking; 1: exact best cell per job; 2: per job, thresholded by global_best
	int prune;              // SW block pruning inside the strips (needs track == 2)
	int prune_i1, prune_j1; // end of the (super) partition: bounds of the distance term of the pruning test
	int opt;                // StripOpt bits: protocol variants of the strip chain (engine default, B200_OPT overrides)
};

// Variants of the strip-chaining protocol (all exact; they only change how often the chain synchronises).
enum StripOpt : int {
	OPT_NO_SC_FENCE   = 1,   // publish with st.release.gpu alone (no extra fence.sc in front of it)
	OPT_SEEN_CACHE    = 2,   // remember the last progress value observed: no acquire while it covers the next block
	OPT_RELEASE_128   = 4,   // after the first 2048 columns release the progress counter every 128 columns, not every 32
	OPT_BEST_EVERY_4  = 8,   // exchange the running best with global_best every 4th block
	OPT_SKIP_128      = 16,  // skip mode decides 128 columns at a time when the strip above is that far ahead
};

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
__device__ __forceinline__ void st_release_sys(int* p, int v) {
	asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// wait until the previous GPU has delivered rows [.., upto) of our left border (lane 0 spins, warp follows)
constexpr unsigned kSpinLimit = 1u << 22;      // each spin costs an L2 round trip (~1 us): a dependency stuck for seconds is a bug
// spin (lane 0) until the strip above has published `need` columns; the watchdog turns a protocol bug into an error
__device__ __forceinline__ void wait_progress(const StripParams& p, int dep, int need, int lane) {
	if (dep < 0) return;
	if (lane == 0) {
		unsigned spins = 0;
		while (ld_acquire(p.progress + dep) < need) {
			if (ld_relaxed(p.stop_flag)) break;
			if (++spins > kSpinLimit) { atomicExch(p.stop_flag, 2); break; }
			__nanosleep(64);
		}
	}
	__syncwarp();
}
// same, returning the progress value observed (warp-uniform) so that the caller can skip later waits it already covers
__device__ __forceinline__ int wait_progress_v(const StripParams& p, int dep, int need, int lane) {
	int v = 0;
	if (lane == 0) {
		unsigned spins = 0;
		while ((v = ld_acquire(p.progress + dep)) < need) {
			if (ld_relaxed(p.stop_flag)) break;
			if (++spins > kSpinLimit) { atomicExch(p.stop_flag, 2); break; }
			__nanosleep(64);
		}
	}
	return __shfl_sync(0xffffffffu, v, 0);
}
__device__ __forceinline__ void wait_left(const StripParams& p, int upto, int lane) {
	if (p.left_ready == nullptr) return;
	if (lane == 0) {
		unsigned spins = 0;
		while (ld_acquire_sys(p.left_ready) < upto) {
			if (ld_relaxed(p.stop_flag)) break;
			if (++spins > kSpinLimit) { atomicExch(p.stop_flag, 3); break; }
			__nanosleep(256);
		}
	}
	__syncwarp();
}
// publish our finished right-border rows to the next GPU; called before the strip's final progress release so
// that publications of consecutive strips are ordered
__device__ __forceinline__ void publish_right(const StripParams& p, int upto, int lane) {
	if (p.right_ready == nullptr) return;
	__syncwarp();
	if (lane == 0) { __threadfence_system(); st_release_sys(p.right_ready, upto); }
}
// tell the host that special row `jb.sra_index` is complete in the device-resident special-rows area, so that it can be
// copied out and dispatched while the kernel keeps running (replaces the blocking per-block D2H of
// R/src/CUDAligner.cpp:393-399)
__device__ __forceinline__ void signal_special_row(const StripParams& p, const StripJob& jb, int lane) {
	if (p.sra_done == nullptr || jb.sra_off < 0) return;
	__syncwarp();
	if (lane == 0) { __threadfence_system(); st_release_sys(p.sra_done + jb.sra_index, 1); }
}
__device__ __forceinline__ void push_best(const StripParams& p, int v) {
	atomicMax(p.global_best, v);
	for (int k = 0; k < p.n_peer_best; k++) atomicMax_system(p.peer_best[k], v);
}
// Warp-uniform read of a word that other warps/GPUs update concurrently.  After lane-divergent code the lanes
// of a warp need not be converged, so 32 independent loads could return different values; anything that feeds
// control flow must be read once and broadcast.
__device__ __forceinline__ int ld_uniform(const int* p) {
	__syncwarp();
	const int v = ld_relaxed(p);
	return __shfl_sync(0xffffffffu, v, 0);
}
__device__ __forceinline__ Cell ldcg_cell(const Cell* p) {
	int2 v = __ldcg(reinterpret_cast<const int2*>(p));
	Cell c; c.h = v.x; c.x = v.y; return c;
}
__device__ __forceinline__ void stcg_cell(Cell* p, int h, int x) {
	__stcg(reinterpret_cast<int2*>(p), make_int2(h, x));
}

// lexicographic "better" for best cells: higher score, then smaller i, then smaller j
// (CPUBlockProcessor.cpp:154-158 row-major strict '<' + BestScoreList.hpp:30-38)
__device__ __forceinline__ bool better(int s, int i, int j, int bs, int bi, int bj) {
	return s > bs || (s == bs && (i < bi || (i == bi && j < bj)));
}

}  // namespace b200
