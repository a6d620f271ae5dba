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
#include <stdint.h>
#include "ptx.cuh"

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
	JOB_RIGHT_PEER = 16,    // chain mode: right_off indexes the exchange cells of the next GPU, not StripParams::right
	JOB_LEFT_XCHG = 32,     // chain mode: left_off indexes our exchange cells (border delivered by the previous GPU), not StripParams::left
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

// ---- block-cyclic multi-GPU chain ---------------------------------------------------------
// seq1 is cut into column CHUNKS; chunk c belongs to GPU (c mod world).  A job is (strip r, local chunk k); it may
// start when (r, c-1) has delivered its right border (left event, from the previous GPU over NVLink) and (r-1, c) has
// published its first columns (top event, same GPU).  Each strip has one 64-bit event word {left events : 32 | top
// events : 32} on the GPU that owns the job; whoever completes the pair pushes the job into that GPU's work queue.
// Warps pop jobs from the queue, so a resident warp never waits for a job that cannot start yet.
struct StripRow {            // one horizontal strip of the partition (identical on every GPU)
	int i0, rows;            // first row (index into seq0), number of rows
	int flags;               // JOB_S32
	int left_off;            // rows above this strip (offset of its slot 0 in a first/last-column array)
	int sra_row;             // index of the special row this strip's bottom row is, or -1
	int pad[3];
};
struct ChunkCol {            // one column chunk owned by this GPU
	int j0, cols;            // first column (index into seq1), number of columns
	int cum;                 // columns of this GPU's earlier chunks (offset in the cumulative progress counters / local SRA rows)
	int gidx;                // global chunk index c
};
struct ChainParams {
	int enabled;
	int world;               // GPUs in the chain
	int nstrips;             // S
	int nchunks_local;       // K: chunks owned by this GPU
	int nchunks_total;       // C
	int left_zero;           // the partition's first column is the constant (H=0, E=-INF)
	long long local_cols;    // sum of this GPU's chunk widths (row pitch of the local special-rows area)
	const StripRow* strips;  // [S]
	const ChunkCol* chunks;  // [K]
	// this GPU's exchange block (peer-visible memory)
	int* queue;              // [K*S] job ids in push order, -1 = not pushed yet
	int* q_tail;             // push counter
	unsigned long long* events;   // [S] {left events << 32 | top events}
	const Cell* my_cells;    // left borders delivered by the previous GPU: strip r at [left_off + r, +rows+1)
	// the exchange block of the GPU that owns the chunks on our right (peer memory; our own block when world == 1)
	int* nx_queue;
	int* nx_tail;
	unsigned long long* nx_events;
	Cell* nx_cells;
};

struct StripParams {
	const unsigned char* s0;
	const unsigned char* s1;
	const unsigned* s0p;    // the same sequences packed 2 bits per base (16 bases per word, code = (byte >> 1) & 3: A 0, C 1, T 2, G 3),
	const unsigned* s1p;    // or NULL (inputs with N / IUPAC bytes, stage-4 reversed copies): the packed kernel then reads the bytes
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
	unsigned long long* cells_done;   // statistics: [0] cells computed, [1] nanoseconds spent by warps inside compute segments
	int* stop_flag;         // non-zero asks the kernel to stop (set by a spin-wait watchdog: no hung GPU on a protocol bug)
	int* peer_best[8];      // multi-GPU: running-best words of the other GPUs (peer memory)
	int n_peer_best;
	int* sra_done;          // host-mapped flags, one per special row: set when the row is complete in the device SRA, or NULL
	int recurrence;         // B200_SMITH_WATERMAN | B200_NEEDLEMAN_WUNSCH
	int track;              // 0: no best tracking; 1: exact best cell per job; 2: per job, thresholded by global_best
	int prune;              // SW block pruning inside the strips (needs track == 2)
	int prune_i1, prune_j1; // end of the (super) partition: bounds of the distance term of the pruning test
	int opt;                // StripOpt bits: protocol variants of the strip chain (engine default, B200_OPT overrides)
	long long watchdog_ns;  // a dependency that shows no progress for this long stops the kernel with an error (0 = never)
	int* sm_load;           // chain mode: [0] warps inside compute segments on this GPU, [1 + 4 * smid + scheduler] the same per warp scheduler
	int nsm;                // SMs of this GPU
	unsigned long long* trace;   // development: per job {pushed, popped, first publication, finished} in globaltimer ns, or NULL
	unsigned long long* nx_trace;   // the same array on the GPU that owns the chunks on our right
	int test_delay_ms;      // test hook: the first job sleeps this long before it starts (tests/test_watchdog_gpu.py)
	ChainParams chain;
};

// A job as the kernels see it: geometry plus where its progress counter and result live.  Kept small on purpose: the
// packed kernel runs at the 128-register limit, so everything the rare paths need is re-derived from `job` there.
struct JobCtx {
	StripJob jb;            // jb.dep = index of the progress counter of the strip above, or -1
	int job;                // job id (chain mode: k * nstrips + r)
	int pidx;               // index of our progress counter and of our result (chain mode: the strip, otherwise the job)
	int prog_base;          // value of the counters that corresponds to column 0 of this job (chain: cumulative over chunks)
};
__device__ __forceinline__ const Cell* left_border(const StripParams& p, const JobCtx& c) {
	return ((c.jb.flags & JOB_LEFT_XCHG) ? p.chain.my_cells : p.left) + c.jb.left_off;
}
__device__ __forceinline__ Cell* right_border(const StripParams& p, const JobCtx& c) {      // only when jb.right_off >= 0
	return ((c.jb.flags & JOB_RIGHT_PEER) ? p.chain.nx_cells : p.right) + c.jb.right_off;
}

// Variants of the strip-chaining protocol (all exact; they only change how often the chain synchronises).
enum StripOpt : int {
	OPT_NO_SC_FENCE   = 1,   // publish with st.release.gpu alone (no extra fence.sc in front of it)
	OPT_SEEN_CACHE    = 2,   // remember the last progress value observed: no acquire while it covers the next block
	OPT_RELEASE_128   = 4,   // after the first 2048 columns release the progress counter every 128 columns, not every 32
	OPT_BEST_EVERY_4  = 8,   // exchange the running best with global_best every 4th block
	OPT_SKIP_128      = 16,  // skip mode decides 128 columns at a time when the strip above is that far ahead
	OPT_LOOKAHEAD     = 32,  // a strip that has to wait for the one above waits for two extra blocks (fewer spin-waits at the same pace)
	OPT_DEFER_RELEASE = 64,  // steady state: publish a block's counter one block later, when its stores have long landed
};

// lexicographic "better" for best cells: higher score, then smaller i, then smaller j
// (CPUBlockProcessor.cpp:154-158 row-major strict '<' + BestScoreList.hpp:30-38)
__device__ __forceinline__ bool better(int s, int i, int j, int bs, int bi, int bj) {
	return s > bs || (s == bs && (i < bi || (i == bi && j < bj)));
}
// Spin-wait watchdog.  Waits in the strip chain are legitimately as long as a whole sweep of a column chunk on another
// GPU, so the limit is a TIME without any observable progress (StripParams::watchdog_ns, sized by the engine from the
// chunk width and the number of GPUs), not a spin count: a protocol bug becomes an error code, a slow neighbour does not.
struct Watchdog {
	unsigned long long t0;
	unsigned spins;
	int last;
	__device__ __forceinline__ Watchdog() : t0(0), spins(0), last(INT_MIN) {}
	// returns true when the wait must be abandoned (stop requested by another warp, or the deadline passed)
	__device__ __forceinline__ bool expired(const StripParams& p, int observed, int code) {
		if (ld_relaxed(p.stop_flag)) return true;
		if ((++spins & 255u) == 0 && p.watchdog_ns > 0) {
			const unsigned long long now = global_ns();
			if (t0 == 0 || observed != last) { t0 = now; last = observed; }
			else if (now - t0 > (unsigned long long)p.watchdog_ns) { atomicExch(p.stop_flag, code); return true; }
		}
		return false;
	}
};
// spin (lane 0) until the strip above has published `need` columns
__device__ __forceinline__ void wait_progress(const StripParams& p, int dep, int need, int lane) {
	if (dep < 0) return;
	if (lane == 0) {
		Watchdog wd;
		int v;
		while ((v = ld_acquire(p.progress + dep)) < need) {
			if (wd.expired(p, v, 2)) break;
			__nanosleep(64);
		}
	}
	__syncwarp();
}
// same, returning the progress value observed (warp-uniform) so that the caller can skip later waits it already covers
__device__ __forceinline__ int wait_progress_v(const StripParams& p, int dep, int need, int lane) {
	int v = 0;
	if (lane == 0) {
		Watchdog wd;
		while ((v = ld_acquire(p.progress + dep)) < need) {
			if (wd.expired(p, v, 2)) break;
			__nanosleep(64);
		}
	}
	return __shfl_sync(0xffffffffu, v, 0);
}

// ---- chain mode: work queue and readiness events -------------------------------------------------------------------
__device__ __forceinline__ void chain_push(int* queue, int* tail, int job) {
	__threadfence_system();
	const int slot = atomicAdd_system(tail, 1);
	st_release_sys(queue + slot, job);
}
// index of this warp's scheduler in StripParams::sm_load (a CTA's four warps sit on the four sub-partitions of its SM)
__device__ __forceinline__ unsigned sched_slot() { return 4u * sm_id() + ((threadIdx.x >> 5) & 3u); }
// Next job of this GPU in push order, or -1 when all of them have been handed out (or the kernel is stopping).
// A job is taken only once it is actually in the queue, and preferably by a warp whose scheduler (SM sub-partition)
// carries no more computing warps than the average one: the strips of a wavefront advance in lockstep (each one waits
// for the strip above), so the whole front moves at the pace of the most crowded scheduler -- with the GPU half empty
// (multi-GPU runs, narrow pruning bands) random placement costs far more than the few microseconds an idle warp waits
// for a better-placed taker.
__device__ __forceinline__ int chain_pop(const StripParams& p, int lane) {
	int job = -1;
	if (lane == 0) {
		Watchdog wd;
		const unsigned slot_id = sched_slot();
		int refused = 0;
		// Idle back-off: 0.5 us doubling to 2 us.  Pick-up latency is part of the start-to-start distance of consecutive
		// strips (a strip may start once the one above has published its first columns), which is paid once per strip while
		// the front ramps up: measured with tools/chain_perf.py, a 64 us ceiling is 7 % slower than 8 us.
		unsigned nap = 500;
		for (;;) {
			const int head = ld_relaxed(p.job_counter);
			if (head >= p.njobs) break;
			const int tail = ld_relaxed_sys(p.chain.q_tail);
			if (head < tail) {
				bool take = true;
				if (p.sm_load != nullptr && refused < 8)
					take = (long long)ld_relaxed(p.sm_load + 1 + slot_id) * (4 * p.nsm) <= (long long)ld_relaxed(p.sm_load);
				if (!take) { refused++; __nanosleep(1000); continue; }
				if (atomicCAS(p.job_counter, head, head + 1) != head) continue;          // another warp took it
				while ((job = ld_acquire_sys(p.chain.queue + head)) < 0) {                // the tail moves before the entry is stored
					if (wd.expired(p, -1, 3)) { job = -1; break; }
				}
				if (p.trace && job >= 0) p.trace[4 * (size_t)job + 1] = global_ns();
				break;
			}
			if (wd.expired(p, tail, 3)) break;
			__nanosleep(nap);
			if (nap < 2000) nap *= 2;
		}
	}
	return __shfl_sync(0xffffffffu, job, 0);
}
// top event: job (strip, k) has published its first columns, so (strip+1, k) may follow it (lane 0 only)
// (out of line and fed with scalars only: a reference to the kernel parameters would force a local copy of them)
__device__ __noinline__ void chain_notify_below_(unsigned long long* events, int* queue, int* tail, int nstrips, int job, unsigned long long* trace) {
	const int k = job / nstrips, r = job - k * nstrips;
	if (trace) trace[4 * (size_t)job + 2] = global_ns();
	if (r + 1 >= nstrips) return;
	const unsigned long long old = atomicAdd_system(events + r + 1, 1ULL);
	if ((unsigned)(old & 0xffffffffu) == (unsigned)k && (unsigned)(old >> 32) >= (unsigned)k + 1u) {
		if (trace) trace[4 * (size_t)(job + 1)] = global_ns();
		chain_push(queue, tail, job + 1);
	}
}
__device__ __forceinline__ void chain_notify_below(const StripParams& p, int job) {
	chain_notify_below_(p.chain.events, p.chain.queue, p.chain.q_tail, p.chain.nstrips, job, p.trace);
}
// left event: job (strip, chunk c) has stored its right border into the next GPU's exchange block, so (strip, c+1) may
// start over there (lane 0 only; the border stores of the other lanes are ordered by the __syncwarp of the caller)
__device__ __noinline__ void chain_notify_right_(const ChunkCol* chunks, unsigned long long* nx_events, int* nx_queue, int* nx_tail,
                                                 int nstrips, int nchunks_total, int world, int job, unsigned long long* trace, unsigned long long* nx_trace) {
	const int k = job / nstrips, r = job - k * nstrips;
	const int gidx = chunks[k].gidx;
	if (trace) trace[4 * (size_t)job + 3] = global_ns();
	if (gidx + 1 >= nchunks_total) return;
	const int kn = (gidx + 1) / world;                 // local index of chunk c+1 on its owner
	__threadfence_system();
	const unsigned long long old = atomicAdd_system(nx_events + r, 1ULL << 32);
	if ((unsigned)(old >> 32) == (unsigned)kn && (unsigned)(old & 0xffffffffu) >= (unsigned)kn + 1u) {
		if (nx_trace) nx_trace[4 * (size_t)(kn * nstrips + r)] = global_ns();
		chain_push(nx_queue, nx_tail, kn * nstrips + r);
	}
}
__device__ __forceinline__ void chain_notify_right(const StripParams& p, int job) {
	chain_notify_right_(p.chain.chunks, p.chain.nx_events, p.chain.nx_queue, p.chain.nx_tail, p.chain.nstrips, p.chain.nchunks_total, p.chain.world, job, p.trace, p.nx_trace);
}

// CHAIN is a compile-time property of the packed kernel: the single-GPU instances carry none of the chain's code or
// registers (measured on a 4.6M x 5M pair, profiles/r02_single_gpu_ab.txt: the run-time switch alone cost 10 % with
// pruning and 1 % without, the per-segment statistics and the left-border term of the pruning test 2 % each -- the
// kernel sits at the 128-register limit).  The int32 kernel takes the switch at run time (CHAIN < 0).
template <int CHAIN>
__device__ __forceinline__ bool chain_on(const StripParams& p) { return CHAIN < 0 ? p.chain.enabled != 0 : CHAIN != 0; }
// next job of this launch, or -1 (all lanes return the same value)
template <int CHAIN>
__device__ __forceinline__ int claim_job(const StripParams& p, int lane) {
	int job;
	if (chain_on<CHAIN>(p)) job = chain_pop(p, lane);
	else {
		job = 0;
		if (lane == 0) job = atomicAdd(p.job_counter, 1);
		job = __shfl_sync(0xffffffffu, job, 0);
		if (job >= p.njobs) job = -1;
	}
	if (job == 0 && p.test_delay_ms > 0) {
		// test hook: everything chained behind the first strip must sit through this delay without tripping the watchdog
		const unsigned long long t0 = global_ns();
		while (global_ns() - t0 < (unsigned long long)p.test_delay_ms * 1000000ULL) __nanosleep(100000);
	}
	return job;
}

// Resolve job id -> JobCtx.  Outside chain mode the job table is explicit; in chain mode job = k * S + r and the
// geometry comes from the strip and chunk tables.
template <int CHAIN>
__device__ __forceinline__ JobCtx fetch_job(const StripParams& p, int job) {
	JobCtx c;
	c.job = job;
	if (!chain_on<CHAIN>(p)) {
		c.jb = p.jobs[job];
		c.pidx = job; c.prog_base = 0;
		return c;
	}
	const ChainParams& ch = p.chain;
	const int k = job / ch.nstrips, r = job - k * ch.nstrips;
	const StripRow sr = ch.strips[r];
	const ChunkCol cc = ch.chunks[k];
	c.jb.i0 = sr.i0; c.jb.rows = sr.rows; c.jb.j0 = cc.j0; c.jb.cols = cc.cols;
	c.jb.dep = r - 1;
	c.jb.flags = sr.flags;
	c.jb.sra_off = sr.sra_row >= 0 ? (long long)sr.sra_row * ch.local_cols + cc.cum : -1;
	c.jb.sra_index = (sr.sra_row >= 0 && k == ch.nchunks_local - 1) ? sr.sra_row : -1;   // the row is complete on this GPU after its last chunk
	// borders inside the chain live in the exchange blocks, one private range of rows+1 slots per strip (offset
	// left_off + r); the partition's own first / last column keep the layout of the single-GPU path
	if (cc.gidx == 0) { c.jb.left_off = sr.left_off; if (ch.left_zero) c.jb.flags |= JOB_LEFT_ZERO; }
	else { c.jb.left_off = sr.left_off + r; c.jb.flags |= JOB_LEFT_XCHG; }
	if (cc.gidx == ch.nchunks_total - 1) c.jb.right_off = p.right ? sr.left_off : -1;
	else { c.jb.right_off = sr.left_off + r; c.jb.flags |= JOB_RIGHT_PEER; }
	c.pidx = r; c.prog_base = cc.cum;
	return c;
}
// Chain mode keeps ONE result per strip: the jobs of a strip run strictly one after the other (chunk c+1 starts from
// the border that chunk c delivers at its very end), so each job folds the strip's previous best into its own.
template <int CHAIN>
__device__ __forceinline__ void store_result(const StripParams& p, const JobCtx& c, int bs, int bi, int bj) {
	Score3 o; o.score = bs == INT_MIN ? -kInf : bs; o.i = bi; o.j = bj; o.pad = 0;
	if (chain_on<CHAIN>(p) && c.job >= p.chain.nstrips) {
		// written by another SM: read through L2 (the causality chain is (r,k-1) result -> fence -> left event -> ... -> our pop)
		const int4 q = __ldcg(reinterpret_cast<const int4*>(p.results + c.pidx));
		if (q.y >= 0 && (o.i < 0 || better(q.x, q.y, q.z, o.score, o.i, o.j))) { o.score = q.x; o.i = q.y; o.j = q.z; }
	}
	__stcg(reinterpret_cast<int4*>(p.results + c.pidx), make_int4(o.score, o.i, o.j, 0));
}
// tell the host that special row `jb.sra_index` is complete in the device-resident special-rows area, so that it can be
// copied out and dispatched while the kernel keeps running (replaces the blocking per-block D2H of
// R/src/CUDAligner.cpp:393-399)
__device__ __forceinline__ void signal_special_row(const StripParams& p, const StripJob& jb, int lane) {
	if (p.sra_done == nullptr || jb.sra_index < 0) return;
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
// base `idx` of a 2-bit packed sequence
__device__ __forceinline__ int packed_code(const unsigned* sp, int idx) { return (int)((__ldg(sp + (idx >> 4)) >> ((idx & 15) * 2)) & 3u); }
__device__ __forceinline__ Cell ldcg_cell(const Cell* p) {
	int2 v = __ldcg(reinterpret_cast<const int2*>(p));
	Cell c; c.h = v.x; c.x = v.y; return c;
}
__device__ __forceinline__ void stcg_cell(Cell* p, int h, int x) {
	__stcg(reinterpret_cast<int2*>(p), make_int2(h, x));
}


}  // namespace b200
