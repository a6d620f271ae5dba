// engine.cu -- host side of libb200align.so: device buffers, strip-job construction, kernel launches and the
// C ABI declared in include/b200align.h.  No CPU fallback: every entry point fails loudly without a GPU.
// One translation unit: this file holds the handle, the launch helpers and the lifetime calls; the entry points live in
// engine_sequences.inl, engine_partition.inl, engine_diag.inl, engine_chain.inl, engine_stage4.inl, engine_stage5.inl,
// included at the end (the kernels are instantiated once, by launch_strips below).
#include "ptx.cuh"       // <cuda_runtime.h> + the inline-PTX wrappers + B200_LAUNCH
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/b200align.h"
#include "strip_common.cuh"
#include "strip_s32.cuh"
#include "strip_s16.cuh"
#include "stage4.cuh"
#include "stage5.cuh"

using namespace b200;

static_assert(sizeof(b200_cell) == sizeof(Cell), "cell layout");

namespace {

std::string g_create_error;

template <class T>
struct DevBuf {
	T* p = nullptr;
	size_t cap = 0;
	cudaError_t reserve(size_t n) {
		if (n <= cap) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		size_t want = n + n / 8 + 64;
		cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
		if (e == cudaSuccess) cap = want;
		return e;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <class T>
struct PinBuf {
	T* p = nullptr;
	size_t cap = 0;
	cudaError_t reserve(size_t n) {
		if (n <= cap) return cudaSuccess;
		if (p) cudaFreeHost(p);
		p = nullptr; cap = 0;
		size_t want = n + n / 8 + 64;
		cudaError_t e = cudaMallocHost((void**)&p, want * sizeof(T));
		if (e == cudaSuccess) cap = want;
		return e;
	}
	void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace

struct b200_handle {
	b200_config cfg;
	int sm_count = 0;
	cudaStream_t stream = nullptr, copy_stream = nullptr;
	int* sra_flags = nullptr; size_t sra_flags_cap = 0;     // host-mapped "special row k is complete" flags
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	std::string err;

	// sequences on the device: 2 bits per base for pure A/C/G/T inputs (what the packed kernel reads; only these words
	// cross PCIe) plus the byte view of the reference (R/src/cuda_util.cpp:50-56) for the int32 kernel, rebuilt on the device
	DevBuf<unsigned char> s0, s1;
	DevBuf<unsigned> s0p, s1p;
	PinBuf<unsigned> hpack;
	bool packed = false;
	int n0 = 0, n1 = 0;
	bool acgt_only = false;
	std::vector<unsigned char> bad0;   // per 64 rows of seq0: 1 when the block holds a non-ACGT byte (forces the int32 strip path)

	// strip machinery
	DevBuf<Cell> busH, left, right, sra;
	DevBuf<StripJob> jobs;
	DevBuf<int> progress;
	DevBuf<Score3> results;
	DevBuf<int> scalars;            // [0] job counter, [1] global best, [2] stop flag, [4..5] cells (u64), [6..7] busy ns (u64)
	DevBuf<int> smload;             // chain mode: computing warps of the GPU [0] and per SM [1 + smid]
	DevBuf<Cell> matchbuf;          // scratch of b200_match_last_column (its own: the call may come between chunked launches)
	DevBuf<int> matchflag;
	PinBuf<int> hmatchflag;
	PinBuf<Cell> hcells;            // pinned staging for rows / columns
	PinBuf<Score3> hresults;
	PinBuf<int> hscalars;
	std::vector<StripJob> hjobs;

	// diag-mode state (R/src/CUDAligner.hpp:216-232 contract)
	struct {
		bool active = false;
		b200_partition part;
		int B = 0, bh = 0;
		std::vector<int> split;
		DevBuf<Cell> vbuf;          // [2][B+1][bh+1] vertical borders, ping-pong by diagonal parity
		DevBuf<Cell> col0;          // [2][bh+1] first-column chunks (next / current)
		int col0_cur = 0;
		bool col0_valid[2] = {false, false};
		int last_diag = -1;
		std::vector<b200_score> scores;
		PinBuf<Cell> hlastcol; int hlastcol_diag = -2;
	} dg;

	// multi-GPU chain: one peer-visible exchange block per GPU (ExLayout below): control words (running best, queue tail),
	// per-strip event words, the work queue, and the left-border cells delivered by the GPU on our left
	struct {
		int* block = nullptr;
		long long cap_rows = 0, cap_strips = 0, cap_jobs = 0;
		int rank = -1, world = 0;
		int* peers[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
		bool connected = false;
		bool ipc = false;              // peers were mapped with cudaIpcOpenMemHandle (one process per GPU) rather than peer access
		unsigned epoch = 0;            // chained calls made so far: selects the running-best word
		DevBuf<StripRow> strips;
		DevBuf<ChunkCol> chunks;
		PinBuf<Cell> hrow;             // pinned staging for rows assembled from several GPUs
	} mg;
	b200_result last_chain;            // per-GPU figures of the last chained call (b200_group_rank_result)
	// per-launch overrides of the border / sharing pointers (diag mode and chain mode)
	struct {
		const Cell* left = nullptr; Cell* right = nullptr; bool no_right = false;
		ChainParams chain;
		int* gbest = nullptr; int* peer_best[8]; int npeer = 0;
		int prune = 0, prune_i1 = 0, prune_j1 = 0;
		const unsigned char* s0 = nullptr; const unsigned char* s1 = nullptr; Cell* busH = nullptr;
		int job_off = 0; int* counter = nullptr;
		long long chunk_cols_max = 0;
		unsigned long long* trace = nullptr; unsigned long long* nx_trace = nullptr;
		int* sra_done = nullptr;
		bool mixed = false;
	} ov;
	// stage 4
	struct {
		DevBuf<unsigned char> s0r, s1r;
		bool rev_valid = false;
		DevBuf<Cell> bus[4], left;
		DevBuf<S4Half> halves; DevBuf<S4Part> parts; DevBuf<XPoint> out;
	} s4;
	// stage 5
	struct {
		DevBuf<S5Part> parts; DevBuf<S5Out> out; DevBuf<unsigned char> ops, flags; DevBuf<int> rows;
	} s5;

	Cell cont_corner;              // first-column cell of the last row of the previous chunk (B200_CONT_CHUNK)
	bool cont_force32 = false;     // the chunked partition in progress runs the int32 kernel (-INF in an NW border)
	int last_grid_warps = 0;
	long long stat_cells = 0;
	long long stat_launches = 0;
};

#define CU(h, call)                                                                                   \
	do {                                                                                              \
		cudaError_t e_ = (call);                                                                      \
		if (e_ != cudaSuccess) {                                                                      \
			(h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
			return 1;                                                                                 \
		}                                                                                             \
	} while (0)

namespace {

constexpr int kR32 = 16;             // rows per lane, s32 kernel  -> 512-row strips (== reference block height 4*128)
constexpr int kSH32 = 32 * kR32;

__global__ void fill_cells_kernel(Cell* dst, long long n, int type, int start_pos, int h_const) {
	long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (k >= n) return;
	Cell c;
	c.x = -kInf;
	if (type == B200_INIT_ZEROES) c.h = h_const;
	else {
		long long pos = start_pos + k;
		c.h = (pos == 0) ? 0 : (int)(-kGapExt * pos - (type == B200_INIT_GAPS ? kGapOpen : 0));   // InitialCellsReader.cpp:84-108
	}
	dst[k] = c;
}

__global__ void fill_const_kernel(Cell* dst, long long n, int hv, int xv) {
	long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (k < n) { dst[k].h = hv; dst[k].x = xv; }
}

__global__ void match_column_kernel(const Cell* buffer, const Cell* base, int len, int goal, int gap_open, int* out) {
	// first k with H+H == goal, else E+E+open == goal, error if a sum exceeds the goal (AlignerUtils.cpp:59-84).
	// out[0] = smallest k with any event, encoded as k*4 + kind (0 match, 1 gap, 2 err1, 3 err2); INT_MAX if none.
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	int code = INT_MAX;
	if (k < len) {
		int sm = base[k].h + buffer[k].h;
		int sg = base[k].x + buffer[k].x + gap_open;
		if (sm == goal) code = k * 4 + 0;
		else if (sg == goal) code = k * 4 + 1;
		else if (sm > goal) code = k * 4 + 2;
		else if (sg > goal) code = k * 4 + 3;
	}
	for (int d = 16; d >= 1; d >>= 1) code = min(code, __shfl_xor_sync(0xffffffffu, code, d));
	if ((threadIdx.x & 31) == 0 && code != INT_MAX) atomicMin(out, code);
}

// Grid of the persistent strip kernel.  All CTAs must be co-resident (a strip spins on the progress of the
// strip above it), so the grid never exceeds SMs x occupancy.  Below that limit the number of resident warps
// per SM is chosen so that the strips fill whole waves: every strip sweeps the full width at the pace of one
// warp, so a last, nearly empty wave would cost as much as a full one.  kSatWarps: resident strip-warps per SM
// beyond which the measured throughput no longer grows.  With the LUT kernel the rate still grows up to the 16
// warps that fit (5M x 5M without pruning: 4083 GCUPS at 11 warps/SM, 4426 at 16; profiles/r01_protocol_options.txt).
constexpr int kSatWarps = 16;
// protocol variant of the strip chain (StripOpt bits); B200_OPT overrides the default for experiments
// OPT_LOOKAHEAD / OPT_DEFER_RELEASE (round 2) are exact too but bought nothing: 2813 / 2844 / 2797 / 2831 ms for the four
// combinations on a pruned 3.45M x 3.75M run, 567 / 590 / 568 / 599 ms on a 4-rank chain (profiles/r02_chain_starvation.txt)
constexpr int kDefaultStripOpt = OPT_NO_SC_FENCE | OPT_SEEN_CACHE | OPT_BEST_EVERY_4 | OPT_SKIP_128;   // measured: profiles/r01_protocol_options.txt; OPT_RELEASE_128 is within noise at 5M x 5M and lengthens the pipeline fill of narrow partitions
int strip_opt() {
	const char* e = getenv("B200_OPT");      // read per launch: experiments switch it between runs of one process
	return e ? atoi(e) : kDefaultStripOpt;
}

// Time without observable progress after which a spin-wait gives up (strip_common.cuh Watchdog).  A wait in the chain
// is legitimately as long as the sweep of one column chunk on every other GPU; B200_WATCHDOG_S overrides (0 = never).
long long watchdog_ns(b200_handle* h) {
	const char* e = getenv("B200_WATCHDOG_S");
	if (e) return (long long)(atof(e) * 1e9);
	long long ns = 30LL * 1000000000LL;
	if (h->ov.chain.enabled)      // 2 us per column and hop: four times the sweep time of a chunk at full occupancy
		ns += h->ov.chunk_cols_max * 2000LL * std::max(1, h->ov.chain.world);
	return ns;
}

int grid_for(b200_handle* h, const void* kernel, int njobs, bool chained) {
	int per_sm = 0;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kWarpsPerBlock * 32, 0);
	if (per_sm < 1) per_sm = 1;
	int lim = per_sm * kWarpsPerBlock;
	if (h->cfg.warps_per_sm > 0) lim = std::min(lim, h->cfg.warps_per_sm);
	int best_w = lim;
	// with pruning about half of the resident strips are skipping (and mostly sleeping): keep every slot occupied
	if (h->cfg.warps_per_sm <= 0 && chained && !h->ov.prune && !h->ov.chain.enabled && njobs > h->sm_count * kSatWarps) {
		double best_cost = 1e300;
		for (int w = std::min(kSatWarps, lim); w <= lim; w++) {
			long long cap = (long long)h->sm_count * w;
			long long waves = (njobs + cap - 1) / cap;
			double cost = (double)waves * std::max(kSatWarps, w);
			if (cost < best_cost - 1e-9) { best_cost = cost; best_w = w; }
		}
	}
	int cap = best_w * h->sm_count / kWarpsPerBlock;
	int need = (njobs + kWarpsPerBlock - 1) / kWarpsPerBlock;
	return std::max(1, std::min(cap, need));
}

// Launch the strip kernel over h->hjobs (already uploaded to h->jobs).
int launch_strips(b200_handle* h, int njobs, int recurrence, int track, int kernel_kind, int SH, bool chained) {
	StripParams sp;
	sp.s0 = h->ov.s0 ? h->ov.s0 : h->s0.p; sp.s1 = h->ov.s1 ? h->ov.s1 : h->s1.p;
	sp.s0p = (h->packed && !h->ov.s0) ? h->s0p.p : nullptr; sp.s1p = (h->packed && !h->ov.s1) ? h->s1p.p : nullptr;
	sp.busH = h->ov.busH ? h->ov.busH : h->busH.p; sp.sra = h->sra.p;
	sp.left = h->ov.left ? h->ov.left : h->left.p;
	sp.right = h->ov.no_right ? nullptr : (h->ov.right ? h->ov.right : h->right.p);
	sp.chain = h->ov.chain;
	sp.sm_load = nullptr; sp.nsm = h->sm_count;
	sp.trace = h->ov.trace; sp.nx_trace = h->ov.nx_trace;
	// scheduler-balanced job placement (strip_common.cuh chain_pop): implemented, measured, and OFF by default -- on an
	// under-filled GPU it was 10 % slower than first-come placement (profiles/r02_chain_starvation.txt)
	if (sp.chain.enabled && getenv("B200_SM_BALANCE")) {
		CU(h, h->smload.reserve(1024));
		CU(h, cudaMemsetAsync(h->smload.p, 0, 1024 * sizeof(int), h->stream));
		sp.sm_load = h->smload.p;
	}
	sp.watchdog_ns = watchdog_ns(h);
	{ const char* e = getenv("B200_TEST_DELAY_MS"); sp.test_delay_ms = e ? atoi(e) : 0; }
	sp.n_peer_best = h->ov.npeer;
	for (int k = 0; k < 8; k++) sp.peer_best[k] = k < h->ov.npeer ? h->ov.peer_best[k] : nullptr;
	sp.jobs = h->jobs.p + h->ov.job_off; sp.njobs = njobs;
	sp.job_counter = h->ov.counter ? h->ov.counter : h->scalars.p + (h->ov.chain.enabled ? 32 : 0);   // chain: polled by idle warps, own 128-byte line
	sp.global_best = h->ov.gbest ? h->ov.gbest : h->scalars.p + 1;
	sp.stop_flag = h->scalars.p + 2;
	sp.cells_done = reinterpret_cast<unsigned long long*>(h->scalars.p + 4);
	sp.progress = h->progress.p + h->ov.job_off;
	sp.results = h->results.p + h->ov.job_off;
	sp.sra_done = h->ov.sra_done;
	sp.recurrence = recurrence;
	sp.track = track;
	sp.prune = h->ov.prune; sp.prune_i1 = h->ov.prune_i1; sp.prune_j1 = h->ov.prune_j1;
	sp.opt = strip_opt();
	const bool sw = recurrence == B200_SMITH_WATERMAN;
	const void* fn = nullptr;
	const bool chain = sp.chain.enabled != 0;
	// packed kernel <rows per virtual lane, SW, TRACK, MIXED, CHAIN>: every combination in use is its own instance
#define S16_PICK(R, MIXED)                                                                                                   \
	(chain ? (sw ? (track ? (const void*)strip_kernel_s16<R, true, true, MIXED, true> : (const void*)strip_kernel_s16<R, true, false, MIXED, true>)    \
	             : (track ? (const void*)strip_kernel_s16<R, false, true, MIXED, true> : (const void*)strip_kernel_s16<R, false, false, MIXED, true>)) \
	       : (sw ? (track ? (const void*)strip_kernel_s16<R, true, true, MIXED, false> : (const void*)strip_kernel_s16<R, true, false, MIXED, false>)  \
	             : (track ? (const void*)strip_kernel_s16<R, false, true, MIXED, false> : (const void*)strip_kernel_s16<R, false, false, MIXED, false>)))
	if (kernel_kind == B200_KERNEL_S16X2 && SH == kSH16F && h->ov.mixed) fn = S16_PICK(kR16F, true);
	else if (kernel_kind == B200_KERNEL_S16X2 && SH == kSH16F) fn = S16_PICK(kR16F, false);
	else if (kernel_kind == B200_KERNEL_S16X2) fn = S16_PICK(kR16, false);
	else {
		if (sw) fn = track ? (const void*)strip_kernel_s32<kR32, true, true> : (const void*)strip_kernel_s32<kR32, true, false>;
		else    fn = track ? (const void*)strip_kernel_s32<kR32, false, true> : (const void*)strip_kernel_s32<kR32, false, false>;
	}
#undef S16_PICK
	int grid = grid_for(h, fn, njobs, chained);
	h->last_grid_warps = grid * kWarpsPerBlock;
#ifdef B200_EMU      // SIMT emulation build of the test suite (tests/emu): the same kernel, run as a function per emulated thread
	CU(h, emu::launch_kernel(h->stream, dim3(grid), dim3(kWarpsPerBlock * 32), reinterpret_cast<void (*)(const StripParams)>(const_cast<void*>(fn)), sp));
#else
	void* args[] = {(void*)&sp};
	CU(h, cudaLaunchKernel(fn, dim3(grid), dim3(kWarpsPerBlock * 32), args, 0, h->stream));
#endif
	h->stat_launches++;
	return 0;
}

int reset_scalars(b200_handle* h, int global_best) {
	CU(h, h->scalars.reserve(64));
	CU(h, h->hscalars.reserve(64));
	int* s = h->hscalars.p;
	memset(s, 0, 64 * sizeof(int));
	s[1] = global_best;
	CU(h, cudaMemcpyAsync(h->scalars.p, s, 64 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
	return 0;
}

int strip_height(int kernel_kind, bool fast) { return kernel_kind == B200_KERNEL_S16X2 ? (fast ? kSH16F : kSH16) : kSH32; }

int pick_kernel(b200_handle* h, int requested) {
	int k = requested ? requested : h->cfg.kernel;
	if (k == B200_KERNEL_AUTO) k = h->acgt_only ? B200_KERNEL_S16X2 : B200_KERNEL_S32;
	if (k == B200_KERNEL_S16X2 && !h->acgt_only) k = B200_KERNEL_S32;   // packed kernel needs 2-bit codes
	return k;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------------------------
extern "C" int b200_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

extern "C" const char* b200_last_error(const b200_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int b200_create(const b200_config* cfg, b200_handle** out) {
	if (!out) return 1;
	*out = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) {
		g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); libb200align has no CPU fallback";
		return 2;
	}
	b200_handle* h = new b200_handle();
	memset(&h->cfg, 0, sizeof(h->cfg));
	memset(&h->ov.chain, 0, sizeof(h->ov.chain));
	memset(&h->last_chain, 0, sizeof(h->last_chain));
	if (cfg) h->cfg = *cfg;
	h->cont_corner.h = 0; h->cont_corner.x = -kInf;
	if (h->cfg.device < 0 || h->cfg.device >= ndev) h->cfg.device = h->cfg.device < 0 ? 0 : h->cfg.device % ndev;   // wrap like R/src/CUDAligner.cpp:142-148
	if ((e = cudaSetDevice(h->cfg.device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete h; return 3; }
	cudaDeviceProp prop;
	if ((e = cudaGetDeviceProperties(&prop, h->cfg.device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete h; return 3; }
	if (prop.major < 10) {
		g_create_error = std::string("device ") + prop.name + " is not sm_100-class; this library ships sm_100a code only";
		delete h; return 4;
	}
	h->sm_count = prop.multiProcessorCount;
	if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess ||
	    (e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
	    (e = cudaEventCreate(&h->ev0)) != cudaSuccess || (e = cudaEventCreate(&h->ev1)) != cudaSuccess) {
		g_create_error = cudaGetErrorString(e); delete h; return 3;
	}
	*out = h;
	return 0;
}

extern "C" int b200_mgpu_disconnect(b200_handle* h);
extern "C" void b200_destroy(b200_handle* h) {
	if (!h) return;
	cudaSetDevice(h->cfg.device);
	cudaStreamSynchronize(h->stream);
	h->s0.release(); h->s1.release(); h->s0p.release(); h->s1p.release(); h->hpack.release(); h->busH.release(); h->left.release(); h->right.release(); h->sra.release();
	h->jobs.release(); h->progress.release(); h->results.release(); h->scalars.release();
	h->hcells.release(); h->hresults.release(); h->hscalars.release();
	h->matchbuf.release(); h->matchflag.release(); h->hmatchflag.release(); h->smload.release();
	h->dg.vbuf.release(); h->dg.col0.release(); h->dg.hlastcol.release();
	h->s4.s0r.release(); h->s4.s1r.release(); h->s4.left.release(); h->s4.halves.release(); h->s4.parts.release(); h->s4.out.release();
	for (int k = 0; k < 4; k++) h->s4.bus[k].release();
	h->s5.parts.release(); h->s5.out.release(); h->s5.ops.release(); h->s5.flags.release(); h->s5.rows.release();
	b200_mgpu_disconnect(h);
	cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1);
	cudaStreamDestroy(h->stream);
	cudaStreamDestroy(h->copy_stream);
	if (h->sra_flags) cudaFreeHost(h->sra_flags);
	delete h;
}

extern "C" long long b200_processed_cells(const b200_handle* h) { return h ? h->stat_cells : 0; }
extern "C" long long b200_kernel_launches(const b200_handle* h) { return h ? h->stat_launches : 0; }

#include "engine_sequences.inl"
#include "engine_partition.inl"
#include "engine_diag.inl"
#include "engine_chain.inl"
#include "engine_stage4.inl"
#include "engine_stage5.inl"
