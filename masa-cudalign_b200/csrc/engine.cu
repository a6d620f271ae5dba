// engine.cu -- host side of libb200align.so: device buffers, strip-job construction, kernel launches and the
// C ABI declared in include/b200align.h.  No CPU fallback: every entry point fails loudly without a GPU.
#include <cuda_runtime.h>
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
	void* args[] = {(void*)&sp};
	CU(h, cudaLaunchKernel(fn, dim3(grid), dim3(kWarpsPerBlock * 32), args, 0, h->stream));
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

// ---------------------------------------------------------------------------------------------------------
// sequences
// ---------------------------------------------------------------------------------------------------------
namespace {
__global__ void unpack2_kernel(const unsigned* src, unsigned char* dst, int n) {
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < n) dst[k] = (unsigned char)("ACTG"[(src[k >> 4] >> ((k & 15) * 2)) & 3u]);      // code = (byte >> 1) & 3
}
// One pass over a host sequence: alphabet check + 2-bit packing (16 bases per word).  Returns false at the first
// non-A/C/G/T byte (the words written so far are then meaningless).
bool pack2(const char* s, int n, unsigned* out) {
	int k = 0;
	for (int w = 0; k < n; w++) {
		unsigned v = 0;
		const int lim = n - k < 16 ? n - k : 16;
		for (int q = 0; q < lim; q++) {
			const unsigned char c = (unsigned char)s[k + q];
			if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T')) return false;
			v |= (unsigned)((c >> 1) & 3) << (2 * q);
		}
		out[w] = v;
		k += lim;
	}
	return true;
}
}  // namespace

extern "C" int b200_set_sequences(b200_handle* h, const char* seq0, int seq0_len, const char* seq1, int seq1_len) {
	if (!h) return 1;
	if (!seq0 || !seq1 || seq0_len < 0 || seq1_len < 0) { h->err = "b200_set_sequences: bad arguments"; return 1; }
	CU(h, cudaSetDevice(h->cfg.device));
	CU(h, h->s0.reserve((size_t)seq0_len + 64));
	CU(h, h->s1.reserve((size_t)seq1_len + 64));
	const size_t w0 = ((size_t)seq0_len + 15) / 16, w1 = ((size_t)seq1_len + 15) / 16;
	CU(h, h->hpack.reserve(w0 + w1 + 2));
	// FASTA bytes -> 2-bit words on the host (the reference keeps one byte per base, C/common/biology/SequenceData.cpp:67-114):
	// pure A/C/G/T inputs cross PCIe packed and stay packed in HBM for the DPX kernel
	h->acgt_only = pack2(seq0, seq0_len, h->hpack.p) && pack2(seq1, seq1_len, h->hpack.p + w0);
	h->packed = h->acgt_only && !getenv("B200_NO_PACK");
	h->bad0.assign((size_t)seq0_len / 64 + 2, 0);
	if (h->packed) {
		CU(h, h->s0p.reserve(w0 + 1));
		CU(h, h->s1p.reserve(w1 + 1));
		CU(h, cudaMemcpyAsync(h->s0p.p, h->hpack.p, w0 * sizeof(unsigned), cudaMemcpyHostToDevice, h->stream));
		CU(h, cudaMemcpyAsync(h->s1p.p, h->hpack.p + w0, w1 * sizeof(unsigned), cudaMemcpyHostToDevice, h->stream));
		if (seq0_len) unpack2_kernel<<<(seq0_len + 255) / 256, 256, 0, h->stream>>>(h->s0p.p, h->s0.p, seq0_len);
		if (seq1_len) unpack2_kernel<<<(seq1_len + 255) / 256, 256, 0, h->stream>>>(h->s1p.p, h->s1.p, seq1_len);
		h->stat_launches += 2;
	} else {
		CU(h, cudaMemcpyAsync(h->s0.p, seq0, (size_t)seq0_len, cudaMemcpyHostToDevice, h->stream));
		CU(h, cudaMemcpyAsync(h->s1.p, seq1, (size_t)seq1_len, cudaMemcpyHostToDevice, h->stream));
		if (!h->acgt_only)
			for (int k = 0; k < seq0_len; k++) { unsigned char c = (unsigned char)seq0[k]; if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T')) h->bad0[k >> 6] = 1; }
	}
	h->n0 = seq0_len; h->n1 = seq1_len;
	h->s4.rev_valid = false;
	CU(h, h->busH.reserve((size_t)seq1_len + 64));
	CU(h, cudaStreamSynchronize(h->stream));
	CU(h, cudaGetLastError());
	return 0;
}

extern "C" int b200_unset_sequences(b200_handle* h) {
	if (!h) return 1;
	h->n0 = h->n1 = 0;
	return 0;
}

// ---------------------------------------------------------------------------------------------------------
// (2) whole-partition path
// ---------------------------------------------------------------------------------------------------------
namespace {

// Row ids (number of rows above the special row, relative to i0) at which the reference flushes special rows:
// AbstractDiagonalAligner::isSpecialRow (AbstractDiagonalAligner.cpp:466-478) with block height bh.
void special_row_ids(int height, int bh, int interval, std::vector<int>& ids) {
	ids.clear();
	if (interval <= 0 || bh <= 0) return;
	int fbi = (interval + bh - 1) / bh;
	if (fbi <= 0) fbi = 1;
	if (fbi <= 8192 / bh) fbi = 8192 / bh;
	if (fbi <= 0) fbi = 1;
	for (long long by = fbi; by * bh < height; by += fbi) ids.push_back((int)(by * bh));
}

}  // namespace

extern "C" int b200_special_row_ids(int height, int block_height, int interval, int* out, int cap) {
	std::vector<int> ids;
	special_row_ids(height, block_height, interval, ids);
	for (size_t k = 0; k < ids.size() && (int)k < cap && out; k++) out[k] = ids[k];
	return (int)ids.size();
}

// Strips of a partition: cut every kSH16F (packed) / kSH32 (int32) rows and additionally at the reference's special-row
// ids, so that every special row is the bottom row of a strip.  Identical on every GPU of a chain.
static void build_strips(b200_handle* h, const b200_partition* p, int m, const std::vector<int>& sr_ids,
                         std::vector<StripRow>& rows, bool& any_s16, int sh16 = kSH16F, bool force32 = false) {
	rows.clear();
	any_s16 = false;
	// rows [a, b) of the partition free of non-ACGT bytes?  (64-row granularity, conservative)
	auto rows_clean = [&](int a, int b) {
		if (h->acgt_only) return true;
		for (int k = (p->i0 + a) >> 6; k <= (p->i0 + b - 1) >> 6; k++) if (h->bad0[k]) return false;
		return true;
	};
	// the packed kernel may be used for a strip iff the caller allows it and the strip's ROWS are pure A/C/G/T
	// (non-ACGT COLUMN bytes are exact in the packed kernel: they mismatch every A/C/G/T row)
	const bool allow16 = !force32 && (h->cfg.kernel == B200_KERNEL_AUTO || h->cfg.kernel == B200_KERNEL_S16X2);
	size_t next_sr = 0;
	int r = 0;
	while (r < m) {
		int lim = m;
		if (next_sr < sr_ids.size()) lim = std::min(lim, sr_ids[next_sr]);
		int end;
		bool s16;
		if (allow16 && rows_clean(r, std::min(lim, r + kSH32))) {
			s16 = true;
			end = std::min(lim, r + kSH32);
			if (sh16 > kSH32 && end == r + kSH32 && end < lim && rows_clean(end, std::min(lim, r + sh16))) end = std::min(lim, r + sh16);
		} else {
			s16 = false;
			end = std::min(lim, r + kSH32);
		}
		StripRow sr;
		memset(&sr, 0, sizeof(sr));
		sr.i0 = p->i0 + r; sr.rows = end - r; sr.left_off = r;
		sr.flags = s16 ? 0 : JOB_S32;
		sr.sra_row = -1;
		if (next_sr < sr_ids.size() && sr_ids[next_sr] == end) sr.sra_row = (int)next_sr++;
		if (s16) any_s16 = true;
		rows.push_back(sr);
		r = end;
	}
}

static int chain_align(b200_handle* const* hs, int nlocal, const b200_partition* p, const b200_callbacks* cb, b200_result* out);

// The on-device special-rows area holds every special row of the partition until the host has copied it out (rows are
// streamed while the kernel runs, but their slots are not recycled).  The reference bounds the NUMBER of rows by
// --ram-size + --disk-size (C/common/Job.cpp:231-257: interval = rows * 8 * n / budget), so the area is at most that
// budget; a budget beyond the free HBM is refused here with the numbers instead of a bare cudaMalloc error.
static int reserve_sra(b200_handle* h, size_t rows, size_t cols) {
	if (rows == 0) return 0;
	if (h->sra.reserve(rows * cols) != cudaSuccess) {
		cudaGetLastError();
		size_t fr = 0, tot = 0;
		cudaMemGetInfo(&fr, &tot);
		char msg[320];
		snprintf(msg, sizeof(msg), "device special-rows area: %zu rows x %zu columns x 8 B = %.1f GB do not fit into the %.1f GB of free HBM; "
		         "lower --ram-size/--disk-size (fewer special rows) or split seq1 over more GPUs (--gpus)", rows, cols, rows * cols * 8e-9, fr * 1e-9);
		h->err = msg;
		return 1;
	}
	return 0;
}

extern "C" int b200_align_partition(b200_handle* h, const b200_partition* p, const b200_callbacks* cb, b200_result* out) {
	if (!h) return 1;
	if (!p || !out) { h->err = "b200_align_partition: bad arguments"; return 1; }
	if (p->reserved[0] & B200_MGPU_CHAIN) return chain_align(&h, 1, p, cb, out);
	memset(out, 0, sizeof(*out));
	struct timespec ts_entry; clock_gettime(CLOCK_MONOTONIC, &ts_entry);
	const int m = p->i1 - p->i0, n = p->j1 - p->j0;
	if (m <= 0 || n <= 0 || p->i0 < 0 || p->j0 < 0 || p->i1 > h->n0 || p->j1 > h->n1) { h->err = "b200_align_partition: partition outside the sequences"; return 1; }
	CU(h, cudaSetDevice(h->cfg.device));
	int kind = B200_KERNEL_S16X2;                  // decided per strip below; all-int32 partitions use the int32 kernel
	const int SH = kSH16F;
	const bool sw = p->recurrence == B200_SMITH_WATERMAN;
	const int track = p->want_best_score ? 2 : 0;
	const bool cont = (p->reserved[0] & B200_CONT_CHUNK) != 0;
	const int row_offset = p->reserved[2];
	const int total_rows = p->reserved[3] > 0 ? p->reserved[3] : m;

	// ---- special rows and strips
	int bh = p->block_height > 0 ? p->block_height : 4 * std::min(128, n);
	std::vector<int> sr_ids;
	if (p->want_special_rows) {
		std::vector<int> all_ids;
		special_row_ids(total_rows, bh, p->special_row_interval, all_ids);
		for (int g : all_ids) if (g > row_offset && g <= row_offset + m) sr_ids.push_back(g - row_offset);
	}
	// ---- buffers
	if (reserve_sra(h, sr_ids.size(), (size_t)n)) return 1;
	if (p->want_last_column) CU(h, h->right.reserve((size_t)m + 1));
	const bool have_cb = cb != nullptr;
	if (have_cb) {
		// pinned staging for rows / columns handed to the callbacks (nothing to stage without callbacks)
		size_t stage_cells = std::max<size_t>((size_t)std::max(m, n) + 1, 1024);
		CU(h, h->hcells.reserve(stage_cells));
	}
	if (reset_scalars(h, sw ? 0 : -kInf)) return 1;

	// The packed kernel keeps scores in a 16-bit frame: -INF E/F inputs vanish after one cell exactly as in the reference,
	// but an NW partition whose border carries -INF in H (it can, when the border comes from a pruned neighbour) must
	// drift like the reference's plain int32 arithmetic does -> such partitions run the int32 kernel.
	bool force32 = false;
	auto has_minf_h = [](const Cell* c, size_t len) { for (size_t k = 0; k < len; k++) if (c[k].h <= -kInf / 2) return true; return false; };

	// ---- first row -> busH[j0..j1), first column -> left[0..m]   (AbstractDiagonalAligner.cpp:83-89,409-456)
	Cell corner_col; corner_col.h = 0; corner_col.x = -kInf;
	Cell corner_row = corner_col;
	if (cont) corner_col = h->cont_corner;
	if (!cont && have_cb && cb->receive_first_column) cb->receive_first_column(cb->ctx, reinterpret_cast<b200_cell*>(&corner_col), 1);
	if (!cont && have_cb && cb->receive_first_row) cb->receive_first_row(cb->ctx, reinterpret_cast<b200_cell*>(&corner_row), 1);
	Cell first_row_tail = corner_row;
	if (cont) {
		// top border = last row of the previous chunk, already in busH
	} else if (p->first_row_init == B200_INIT_ZEROES || !(have_cb && cb->receive_first_row)) {
		int type = p->first_row_init == B200_INIT_CUSTOM ? B200_INIT_ZEROES : p->first_row_init;
		fill_cells_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->busH.p + p->j0, n, type, 1, 0);
		h->stat_launches++;
		first_row_tail.h = type == B200_INIT_ZEROES ? 0 : -kGapExt * n - (type == B200_INIT_GAPS ? kGapOpen : 0);
	} else {
		cb->receive_first_row(cb->ctx, reinterpret_cast<b200_cell*>(h->hcells.p), n);
		first_row_tail = h->hcells.p[n - 1];
		if (!sw && has_minf_h(h->hcells.p, (size_t)n)) force32 = true;
		CU(h, cudaMemcpyAsync(h->busH.p + p->j0, h->hcells.p, (size_t)n * sizeof(Cell), cudaMemcpyHostToDevice, h->stream));
		CU(h, cudaStreamSynchronize(h->stream));
	}
	if (p->first_col_init != B200_INIT_ZEROES) {
		CU(h, h->left.reserve((size_t)m + 1));
		if (have_cb && cb->receive_first_column) {
			h->hcells.p[0] = corner_col;
			cb->receive_first_column(cb->ctx, reinterpret_cast<b200_cell*>(h->hcells.p + 1), m);
			h->cont_corner = h->hcells.p[m];
			if (!sw && has_minf_h(h->hcells.p, (size_t)m + 1)) force32 = true;
			CU(h, cudaMemcpyAsync(h->left.p, h->hcells.p, ((size_t)m + 1) * sizeof(Cell), cudaMemcpyHostToDevice, h->stream));
			CU(h, cudaStreamSynchronize(h->stream));
		} else {
			int type = p->first_col_init == B200_INIT_CUSTOM ? B200_INIT_ZEROES : p->first_col_init;
			fill_cells_kernel<<<(m + 1 + 255) / 256, 256, 0, h->stream>>>(h->left.p, (long long)m + 1, type, row_offset, 0);
			h->stat_launches++;
		}
	}

	// ---- strips (after the borders: their content can force the int32 kernel, whose strips are 512 rows)
	if (cont && h->cont_force32) force32 = true;                 // a chunked partition keeps the kernel of its first chunk
	h->cont_force32 = force32;
	std::vector<StripRow> srows;
	bool any_s16 = false;
	build_strips(h, p, m, sr_ids, srows, any_s16, kSH16F, force32);
	h->hjobs.clear();
	for (const StripRow& sr : srows) {
		StripJob j;
		memset(&j, 0, sizeof(j));
		j.i0 = sr.i0; j.rows = sr.rows; j.j0 = p->j0; j.cols = n;
		j.dep = (int)h->hjobs.size() - 1;
		j.flags = sr.flags | (p->first_col_init == B200_INIT_ZEROES ? JOB_LEFT_ZERO : 0);
		j.left_off = sr.left_off;
		j.right_off = p->want_last_column ? sr.left_off : -1;
		j.sra_off = sr.sra_row >= 0 ? (long long)sr.sra_row * n : -1;
		j.sra_index = sr.sra_row;
		h->hjobs.push_back(j);
	}
	const int njobs = (int)h->hjobs.size();
	if (!any_s16) kind = B200_KERNEL_S32;
	CU(h, h->jobs.reserve(njobs));
	CU(h, h->progress.reserve(njobs));
	CU(h, h->results.reserve(njobs));
	CU(h, h->hresults.reserve(njobs));
	CU(h, cudaMemcpyAsync(h->jobs.p, h->hjobs.data(), njobs * sizeof(StripJob), cudaMemcpyHostToDevice, h->stream));
	CU(h, cudaMemsetAsync(h->progress.p, 0, njobs * sizeof(int), h->stream));

	static const bool dbg = getenv("B200_DEBUG") != nullptr;
	auto now_ms = []() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
	const double t_launch = now_ms();
	if (dbg) fprintf(stderr, "[b200] launch: %d strips, prune=%d track=%d kind=%d, %zu special rows; %.1f ms of setup (buffers, borders, jobs)\n", njobs, (int)(p->prune && sw), track, kind, sr_ids.size(),
	                 t_launch - (ts_entry.tv_sec * 1e3 + ts_entry.tv_nsec * 1e-6));
	// ---- the alignment itself: one persistent launch
	// block pruning: SW stage 1 behind a zero first column, as in the reference (sw_stage1.cpp:219-225); a partition that
	// starts from a real left border is pruned only by the chain instances, which carry that border in the pruning test
	h->ov.prune = (p->prune && sw && track == 2 && kind == B200_KERNEL_S16X2 && p->first_col_init == B200_INIT_ZEROES) ? 1 : 0;
	h->ov.prune_i1 = p->super_i1 > 0 ? p->super_i1 : p->i1;
	h->ov.prune_j1 = p->super_j1 > 0 ? p->super_j1 : p->j1;
	// special rows are streamed out while the kernel runs: host-mapped completion flags, one per row
	const bool stream_rows = have_cb && cb->dispatch_row && !sr_ids.empty();
	if (stream_rows) {
		if (h->sra_flags_cap < sr_ids.size()) {
			if (h->sra_flags) cudaFreeHost(h->sra_flags);
			h->sra_flags = nullptr; h->sra_flags_cap = 0;
			CU(h, cudaHostAlloc((void**)&h->sra_flags, (sr_ids.size() + 64) * sizeof(int), cudaHostAllocMapped));
			h->sra_flags_cap = sr_ids.size() + 64;
		}
		memset(h->sra_flags, 0, sr_ids.size() * sizeof(int));
		h->ov.sra_done = h->sra_flags;
	}
	h->ov.mixed = !h->acgt_only;      // N / IUPAC bytes anywhere: PRMT variant (+ int32 strips); pure A/C/G/T: LUT variant
	h->ov.no_right = !p->want_last_column;
	CU(h, cudaEventRecord(h->ev0, h->stream));
	int lrc = launch_strips(h, njobs, p->recurrence, track, kind, SH, true);
	h->ov.sra_done = nullptr;
	h->ov.mixed = false;
	h->ov.prune = 0;
	h->ov.no_right = false;
	if (lrc) return 1;
	CU(h, cudaEventRecord(h->ev1, h->stream));
	size_t rows_streamed = 0;
	std::vector<int> sr_first_h(sr_ids.size(), 0);
	if (stream_rows) {
		// first-column H of every special row (its first dispatched cell), read before the kernel can finish
		if (p->first_col_init != B200_INIT_ZEROES)
			for (size_t k = 0; k < sr_ids.size(); k++)
				CU(h, cudaMemcpyAsync(&sr_first_h[k], &h->left.p[sr_ids[k]].h, sizeof(int), cudaMemcpyDeviceToHost, h->copy_stream));
		CU(h, cudaStreamSynchronize(h->copy_stream));
		volatile int* flags = h->sra_flags;
		while (rows_streamed < sr_ids.size()) {
			if (!flags[rows_streamed]) {
				if (cudaStreamQuery(h->stream) != cudaErrorNotReady) { if (!flags[rows_streamed]) break; }   // kernel over (or failed): fall through
				else { struct timespec ts = {0, 20000}; nanosleep(&ts, nullptr); continue; }
			}
			const size_t k = rows_streamed;
			CU(h, cudaMemcpyAsync(h->hcells.p, h->sra.p + k * (size_t)n, (size_t)n * sizeof(Cell), cudaMemcpyDeviceToHost, h->copy_stream));
			CU(h, cudaStreamSynchronize(h->copy_stream));
			b200_cell fc; fc.h = sr_first_h[k]; fc.x = -kInf;
			cb->dispatch_row(cb->ctx, p->i0 + sr_ids[k], &fc, 1);
			cb->dispatch_row(cb->ctx, p->i0 + sr_ids[k], reinterpret_cast<b200_cell*>(h->hcells.p), n);
			rows_streamed++;
		}
	}
	if (track) CU(h, cudaMemcpyAsync(h->hresults.p, h->results.p, njobs * sizeof(Score3), cudaMemcpyDeviceToHost, h->stream));
	CU(h, cudaMemcpyAsync(h->hscalars.p, h->scalars.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
	CU(h, cudaStreamSynchronize(h->stream));
	CU(h, cudaGetLastError());
	if (dbg) fprintf(stderr, "[b200] kernel done, stop=%d; %zu of %zu special rows streamed while it ran; %.1f ms since launch\n", h->hscalars.p[2], rows_streamed, sr_ids.size(), now_ms() - t_launch);
	if (dbg && h->hscalars.p[2] != 0) {
		std::vector<int> prog(njobs);
		cudaMemcpy(prog.data(), h->progress.p, njobs * sizeof(int), cudaMemcpyDeviceToHost);
		int shown = 0;
		for (int k = 0; k < njobs && shown < 12; k++)
			if (prog[k] < n) { fprintf(stderr, "[b200]   strip %d progress %d / %d (dep progress %d)\n", k, prog[k], n, k ? prog[k - 1] : -1); shown++; }
	}
	if (h->hscalars.p[2] != 0) { h->err = "strip kernel watchdog: a border dependency did not advance (code " + std::to_string(h->hscalars.p[2]) + ")"; return 5; }
	float ms = 0;
	CU(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));

	out->device_ms = ms;
	out->strips = njobs;
	out->kernel_launches = 1;
	out->kernel_used = kind;
	out->cells_total = (long long)m * n;
	out->cells = (long long)*reinterpret_cast<unsigned long long*>(h->hscalars.p + 4);
	h->stat_cells += out->cells;
	{
		const double busy_ns = (double)*reinterpret_cast<unsigned long long*>(h->hscalars.p + 6);
		const double cap_ns = (double)ms * 1e6 * h->last_grid_warps;
		out->reserved[2] = cap_ns > 0 ? (int)(1000.0 * busy_ns / cap_ns) : 0;     // warp-time spent in compute segments, per mille
		out->reserved[3] = h->last_grid_warps;
	}

	b200_score best; best.score = -kInf; best.i = -1; best.j = -1;
	if (track) {
		for (int k = 0; k < njobs; k++) {
			const Score3& s = h->hresults.p[k];
			if (s.i >= 0 && (s.score > best.score || (s.score == best.score && (s.i < best.i || (s.i == best.i && s.j < best.j))))) {
				best.score = s.score; best.i = s.i; best.j = s.j;
			}
		}
	}
	out->best = best;
	if (dbg) fprintf(stderr, "[b200] best %d (%d,%d); dispatching\n", best.score, best.i, best.j);

	// ---- hand the artefacts to the caller in the reference's dispatch format
	if (have_cb) {
		// first-column H values for the first cell of each dispatched row
		int last_first_h = 0;
		if (p->first_col_init != B200_INIT_ZEROES) {
			for (size_t k = rows_streamed; k < sr_ids.size(); k++)
				CU(h, cudaMemcpy(&sr_first_h[k], &h->left.p[sr_ids[k]].h, sizeof(int), cudaMemcpyDeviceToHost));
			CU(h, cudaMemcpy(&last_first_h, &h->left.p[m].h, sizeof(int), cudaMemcpyDeviceToHost));
		}
		if (cb->dispatch_row) {
			for (size_t k = rows_streamed; k < sr_ids.size(); k++) {
				CU(h, cudaMemcpy(h->hcells.p, h->sra.p + k * (size_t)n, (size_t)n * sizeof(Cell), cudaMemcpyDeviceToHost));
				b200_cell fc; fc.h = sr_first_h[k]; fc.x = -kInf;
				cb->dispatch_row(cb->ctx, p->i0 + sr_ids[k], &fc, 1);
				cb->dispatch_row(cb->ctx, p->i0 + sr_ids[k], reinterpret_cast<b200_cell*>(h->hcells.p), n);
			}
			if (p->want_last_row) {
				CU(h, cudaMemcpy(h->hcells.p, h->busH.p + p->j0, (size_t)n * sizeof(Cell), cudaMemcpyDeviceToHost));
				b200_cell fc; fc.h = last_first_h; fc.x = -kInf;
				cb->dispatch_row(cb->ctx, p->i1, &fc, 1);
				cb->dispatch_row(cb->ctx, p->i1, reinterpret_cast<b200_cell*>(h->hcells.p), n);
			}
		}
		if (cb->dispatch_column && p->want_last_column) {
			CU(h, cudaMemcpy(h->hcells.p, h->right.p, ((size_t)m + 1) * sizeof(Cell), cudaMemcpyDeviceToHost));
			b200_cell fc; fc.h = first_row_tail.h; fc.x = -kInf;
			if (!cont) cb->dispatch_column(cb->ctx, p->j1, &fc, 1);
			for (int r = 0; r < m; r += bh) {
				int len = std::min(bh, m - r);
				cb->dispatch_column(cb->ctx, p->j1, reinterpret_cast<b200_cell*>(h->hcells.p + 1 + r), len);
				if (cb->must_continue && !cb->must_continue(cb->ctx)) break;
			}
		}
		if (cb->dispatch_score && track && best.i >= 0) cb->dispatch_score(cb->ctx, best);
		if (dbg) fprintf(stderr, "[b200] artefacts dispatched; %.1f ms since launch\n", now_ms() - t_launch);
	}
	return 0;
}

// ---------------------------------------------------------------------------------------------------------
// (1) diag primitives
// ---------------------------------------------------------------------------------------------------------
extern "C" int b200_diag_begin(b200_handle* h, const b200_partition* p, int grid_width, const int* split, int block_height) {
	if (!h) return 1;
	if (!p || !split || grid_width < 1 || block_height < 1) { h->err = "b200_diag_begin: bad arguments"; return 1; }
	if (p->i0 < 0 || p->j0 < 0 || p->i1 > h->n0 || p->j1 > h->n1 || p->i1 <= p->i0 || p->j1 <= p->j0) { h->err = "b200_diag_begin: partition outside the sequences"; return 1; }
	CU(h, cudaSetDevice(h->cfg.device));
	auto& d = h->dg;
	d.part = *p; d.B = grid_width; d.bh = block_height;
	d.split.assign(split, split + grid_width + 1);
	for (int b = 0; b < grid_width; b++)
		if (d.split[b + 1] <= d.split[b] || d.split[b] < p->j0 || d.split[b + 1] > p->j1) { h->err = "b200_diag_begin: bad column split"; return 1; }
	const size_t slot = (size_t)block_height + 1;
	CU(h, d.vbuf.reserve(2 * (size_t)(grid_width + 1) * slot));
	CU(h, d.col0.reserve(2 * slot));
	CU(h, h->jobs.reserve(grid_width));
	CU(h, h->progress.reserve(grid_width));
	CU(h, h->results.reserve(grid_width));
	CU(h, h->hresults.reserve(grid_width));
	CU(h, h->hcells.reserve(std::max<size_t>((size_t)(p->j1 - p->j0) + 1, slot + 1)));
	CU(h, h->scalars.reserve(8));
	CU(h, h->hscalars.reserve(8));
	d.col0_cur = 0; d.col0_valid[0] = d.col0_valid[1] = false;
	d.last_diag = -1; d.hlastcol_diag = -2;
	b200_score z; z.score = -kInf; z.i = -1; z.j = -1;
	d.scores.assign(grid_width, z);
	d.active = true;
	return 0;
}

extern "C" int b200_diag_set_first_row(b200_handle* h, const b200_cell* cells, int j, int len) {
	// AbstractDiagonalAligner::prepareIterations loads the first row BEFORE initializeDiagonals
	// (AbstractDiagonalAligner.cpp:89,103), so this call only needs the sequences (busH), not an open diag session.
	if (!h) return 1;
	CU(h, cudaSetDevice(h->cfg.device));
	if (!cells || j < 0 || len < 0 || j + len > h->n1) { h->err = "b200_diag_set_first_row: bad range"; return 1; }
	CU(h, cudaMemcpyAsync(h->busH.p + j, cells, (size_t)len * sizeof(Cell), cudaMemcpyHostToDevice, h->stream));
	CU(h, cudaStreamSynchronize(h->stream));
	return 0;
}

extern "C" int b200_diag_set_first_column(b200_handle* h, const b200_cell* cells, int i, int len) {
	if (!h) return 1;
	if (!h->dg.active) { h->err = "diag primitive called outside b200_diag_begin/b200_diag_end"; return 1; }
	auto& d = h->dg;
	(void)i; (void)len;
	// cells[0] = diagonal cell, cells[1..bh] = (H,E) of the chunk; consumed by block (0, by) one call later
	const size_t slot = (size_t)d.bh + 1;
	int nxt = d.col0_cur ^ 1;
	CU(h, cudaMemcpyAsync(d.col0.p + nxt * slot, cells, slot * sizeof(Cell), cudaMemcpyHostToDevice, h->stream));
	CU(h, cudaStreamSynchronize(h->stream));
	d.col0_valid[nxt] = true;
	return 0;
}

extern "C" int b200_diag_process(b200_handle* h, int diagonal, int window_left, int window_right) {
	if (!h) return 1;
	if (!h->dg.active) { h->err = "diag primitive called outside b200_diag_begin/b200_diag_end"; return 1; }
	CU(h, cudaSetDevice(h->cfg.device));
	auto& d = h->dg;
	const b200_partition& p = d.part;
	const size_t slot = (size_t)d.bh + 1;
	const int kind = pick_kernel(h, 0);
	const int SH = strip_height(kind, false);
	if (d.bh > SH) { h->err = "b200_diag_process: block height larger than a strip"; return 1; }
	const int par = diagonal & 1;
	// Lay the left/right border regions out in one address space: [0, 2*(B+1)*slot) = vbuf, then col0.
	// StripParams::left and ::right both point at vbuf; col0 is addressed through a second launch-free trick:
	// block 0 reads its border from col0 copied into vbuf slot [par][0] below.
	if (p.first_col_init != B200_INIT_ZEROES && d.col0_valid[d.col0_cur]) {
		CU(h, cudaMemcpyAsync(d.vbuf.p + ((size_t)par * (d.B + 1) + 0) * slot, d.col0.p + d.col0_cur * slot, slot * sizeof(Cell), cudaMemcpyDeviceToDevice, h->stream));
	}
	h->hjobs.clear();
	std::vector<int> job_bx;
	for (int bx = d.B - 1; bx >= 0; bx--) {
		int by = diagonal - 1 - bx;
		d.scores[bx].score = -kInf; d.scores[bx].i = d.scores[bx].j = -1;
		if (by < 0) continue;
		long long i0 = (long long)p.i0 + (long long)by * d.bh;
		if (i0 >= p.i1) continue;
		int i1 = (int)std::min<long long>(i0 + d.bh, p.i1);
		StripJob j;
		memset(&j, 0, sizeof(j));
		j.i0 = (int)i0; j.rows = i1 - (int)i0; j.j0 = d.split[bx]; j.cols = d.split[bx + 1] - d.split[bx];
		j.dep = -1;
		j.flags = 0;
		if (bx == 0 && p.first_col_init == B200_INIT_ZEROES) j.flags |= JOB_LEFT_ZERO;
		if (bx < window_left || bx > window_right) j.flags |= JOB_PRUNED;
		j.left_off = (int)(((size_t)par * (d.B + 1) + bx) * slot);
		j.right_off = (int)(((size_t)(par ^ 1) * (d.B + 1) + bx + 1) * slot);
		j.sra_off = -1;
		h->hjobs.push_back(j);
		job_bx.push_back(bx);
	}
	d.col0_cur ^= 1;                      // col0cur = col0next (oracle_cpu.cpp / CUDAligner.cpp:474-504)
	d.last_diag = diagonal;
	const int njobs = (int)h->hjobs.size();
	if (njobs == 0) return 0;
	if (reset_scalars(h, -kInf)) return 1;
	CU(h, cudaMemcpyAsync(h->jobs.p, h->hjobs.data(), njobs * sizeof(StripJob), cudaMemcpyHostToDevice, h->stream));
	h->ov.left = d.vbuf.p; h->ov.right = d.vbuf.p;
	int rc = launch_strips(h, njobs, p.recurrence, 1, kind, SH, false);
	h->ov.left = nullptr; h->ov.right = nullptr;
	if (rc) return 1;
	CU(h, cudaMemcpyAsync(h->hresults.p, h->results.p, njobs * sizeof(Score3), cudaMemcpyDeviceToHost, h->stream));
	CU(h, cudaMemcpyAsync(h->hscalars.p, h->scalars.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
	{
		// the last block column's right border of this diagonal travels with the results (one sync per diagonal);
		// b200_diag_get_last_column then serves it from pinned host memory
		const int parn = (diagonal + 1) & 1;
		CU(h, d.hlastcol.reserve(slot));
		CU(h, cudaMemcpyAsync(d.hlastcol.p, d.vbuf.p + ((size_t)parn * (d.B + 1) + d.B) * slot, slot * sizeof(Cell), cudaMemcpyDeviceToHost, h->stream));
	}
	CU(h, cudaStreamSynchronize(h->stream));
	CU(h, cudaGetLastError());
	if (h->hscalars.p[2] != 0) { h->err = "strip kernel watchdog: a border dependency did not advance"; return 5; }
	d.hlastcol_diag = diagonal;
	for (int k = 0; k < njobs; k++) {
		const Score3& s = h->hresults.p[k];
		b200_score& o = d.scores[job_bx[k]];
		o.score = s.score; o.i = s.i; o.j = s.j;
	}
	h->stat_cells += (long long)*reinterpret_cast<unsigned long long*>(h->hscalars.p + 4);
	return 0;
}

extern "C" int b200_diag_get_row(b200_handle* h, int j, int len, b200_cell* out) {
	if (!h) return 1;
	if (!h->dg.active) { h->err = "diag primitive called outside b200_diag_begin/b200_diag_end"; return 1; }
	if (!out || j < 0 || len < 0 || j + len > h->n1) { h->err = "b200_diag_get_row: bad range"; return 1; }
	CU(h, cudaMemcpyAsync(out, h->busH.p + j, (size_t)len * sizeof(Cell), cudaMemcpyDeviceToHost, h->stream));
	CU(h, cudaStreamSynchronize(h->stream));
	return 0;
}

extern "C" int b200_diag_get_last_column(b200_handle* h, int i, int len, b200_cell* out) {
	if (!h) return 1;
	if (!h->dg.active) { h->err = "diag primitive called outside b200_diag_begin/b200_diag_end"; return 1; }
	auto& d = h->dg;
	(void)i;
	if (!out || len < 0 || len > d.bh) { h->err = "b200_diag_get_last_column: bad range"; return 1; }
	// the last block column wrote its right border for the diagonal just processed into parity (last_diag+1)&1, slot B
	const size_t slot = (size_t)d.bh + 1;
	if (d.hlastcol_diag == d.last_diag && d.hlastcol.p) { memcpy(out, d.hlastcol.p + 1, (size_t)len * sizeof(Cell)); return 0; }
	const int par = (d.last_diag + 1) & 1;
	CU(h, cudaMemcpyAsync(out, d.vbuf.p + ((size_t)par * (d.B + 1) + d.B) * slot + 1, (size_t)len * sizeof(Cell), cudaMemcpyDeviceToHost, h->stream));
	CU(h, cudaStreamSynchronize(h->stream));
	return 0;
}

extern "C" int b200_diag_get_block_scores(b200_handle* h, b200_score* out) {
	if (!h) return 1;
	if (!h->dg.active || !out) { h->err = "b200_diag_get_block_scores: no open diag session"; return 1; }
	memcpy(out, h->dg.scores.data(), h->dg.scores.size() * sizeof(b200_score));
	return 0;
}

extern "C" int b200_diag_clear_pruned(b200_handle* h, int j0, int j1) {
	if (!h) return 1;
	if (!h->dg.active) { h->err = "diag primitive called outside b200_diag_begin/b200_diag_end"; return 1; }
	if (j0 < 0 || j1 > h->n1) { h->err = "b200_diag_clear_pruned: bad range"; return 1; }
	if (j1 <= j0) return 0;
	long long n = (long long)j1 - j0;
	fill_const_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->busH.p + j0, n, -kInf, -kInf);
	h->stat_launches++;
	CU(h, cudaGetLastError());
	return 0;
}

extern "C" int b200_diag_end(b200_handle* h) {
	if (!h) return 1;
	h->dg.active = false;
	return 0;
}

extern "C" int b200_match_last_column(b200_handle* h, const b200_cell* buffer, const b200_cell* base, int len, int goal, b200_match* out) {
	if (!h) return 1;
	if (!buffer || !base || !out || len < 0) { h->err = "b200_match_last_column: bad arguments"; return 1; }
	out->found = 0; out->k = -1; out->score = 0; out->type = 0;
	if (len == 0) return 0;
	CU(h, cudaSetDevice(h->cfg.device));
	CU(h, h->matchbuf.reserve(2 * (size_t)len + 2));
	CU(h, h->matchflag.reserve(4));
	CU(h, h->hmatchflag.reserve(4));
	Cell* dbuf = h->matchbuf.p; Cell* dbase = h->matchbuf.p + len;
	CU(h, cudaMemcpyAsync(dbuf, buffer, (size_t)len * sizeof(Cell), cudaMemcpyHostToDevice, h->stream));
	CU(h, cudaMemcpyAsync(dbase, base, (size_t)len * sizeof(Cell), cudaMemcpyHostToDevice, h->stream));
	h->hmatchflag.p[0] = INT_MAX;
	CU(h, cudaMemcpyAsync(h->matchflag.p, h->hmatchflag.p, sizeof(int), cudaMemcpyHostToDevice, h->stream));
	match_column_kernel<<<(len + 255) / 256, 256, 0, h->stream>>>(dbuf, dbase, len, goal, kGapOpen, h->matchflag.p);
	h->stat_launches++;
	CU(h, cudaMemcpyAsync(h->hmatchflag.p, h->matchflag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
	CU(h, cudaStreamSynchronize(h->stream));
	int code = h->hmatchflag.p[0];
	if (code != INT_MAX) {
		int k = code >> 2, kindc = code & 3;
		out->k = k;
		if (kindc == 0) { out->found = 1; out->score = base[k].h; out->type = 0; }
		else if (kindc == 1) { out->found = 1; out->score = base[k].x; out->type = 1; }
		else { out->found = 0; out->type = kindc == 2 ? -1 : -2; }
	}
	return 0;
}

// ---------------------------------------------------------------------------------------------------------
// multi-GPU chain (block-cyclic column chunks, dataflow work queues)
// ---------------------------------------------------------------------------------------------------------
static_assert(sizeof(b200_ipc_handle) >= sizeof(cudaIpcMemHandle_t), "ipc handle size");

namespace {

// Exchange block of one GPU (peer-visible): [64 control ints][events u64 x cap_strips][queue int x cap_jobs][cells].
//   ctrl[1], ctrl[2]  running best score shared by all GPUs; chained call e uses word 1 + (e & 1)
//   ctrl[32]          queue tail (jobs pushed so far)
constexpr int kCtlBest = 1, kCtlTail = 32, kCtlInts = 64;      // the tail is polled by every idle warp: its own 128-byte line
struct ExLayout { size_t off_events, off_queue, off_cells, off_trace, bytes; };
bool trace_enabled() { return getenv("B200_TRACE_DIR") != nullptr; }       // development: per-job timestamps (tools/trace_report.py)
ExLayout ex_layout(long long cap_rows, long long cap_strips, long long cap_jobs) {
	ExLayout l;
	l.off_events = kCtlInts * sizeof(int);
	l.off_queue = l.off_events + (size_t)cap_strips * sizeof(unsigned long long);
	l.off_cells = (l.off_queue + (size_t)cap_jobs * sizeof(int) + 15) & ~(size_t)15;
	l.off_trace = l.off_cells + ((size_t)cap_rows + (size_t)cap_strips + 8) * sizeof(Cell);
	l.bytes = l.off_trace + (trace_enabled() ? (size_t)cap_jobs * 32 : 0);
	return l;
}
long long strips_cap_for(long long max_rows) { return max_rows / 256 + 4096; }

// (Re-)arm an exchange block: empty queue, event words at their start values.  GPU 0 owns chunk 0, whose jobs have no
// left neighbour (one left event pre-counted); strip 0 has no strip above (top events pre-counted); job 0 = (strip 0,
// chunk 0) is therefore ready from the start and pre-pushed.
__global__ void chain_arm_kernel(int* block, size_t off_events, size_t off_queue, long long nstrips, long long njobs, int rank, int best_word) {
	unsigned long long* ev = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(block) + off_events);
	int* q = reinterpret_cast<int*>(reinterpret_cast<char*>(block) + off_queue);
	const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
	for (long long k = tid; k < nstrips; k += nth) ev[k] = (rank == 0 ? (1ULL << 32) : 0ULL) | (k == 0 ? 0x40000000ULL : 0ULL);
	for (long long k = tid; k < njobs; k += nth) q[k] = (rank == 0 && k == 0) ? 0 : -1;
	if (tid == 0) {
		block[kCtlTail] = rank == 0 ? 1 : 0;
		if (best_word < 0) { block[kCtlBest] = INT_MIN; block[kCtlBest + 1] = INT_MIN; }
		else block[kCtlBest + best_word] = INT_MIN;
	}
}

int alloc_exchange(b200_handle* h, long long max_rows, long long max_jobs) {
	CU(h, cudaSetDevice(h->cfg.device));
	if (h->mg.block) { cudaFree(h->mg.block); h->mg.block = nullptr; }
	h->mg.cap_rows = max_rows; h->mg.cap_strips = strips_cap_for(max_rows); h->mg.cap_jobs = std::max<long long>(max_jobs, 1);
	const ExLayout l = ex_layout(h->mg.cap_rows, h->mg.cap_strips, h->mg.cap_jobs);
	CU(h, cudaMalloc((void**)&h->mg.block, l.bytes));
	CU(h, cudaMemset(h->mg.block, 0, l.off_cells));
	if (trace_enabled()) CU(h, cudaMemset(reinterpret_cast<char*>(h->mg.block) + l.off_trace, 0, l.bytes - l.off_trace));
	return 0;
}

int arm_exchange(b200_handle* h, long long nstrips, long long njobs, int best_word) {
	const ExLayout l = ex_layout(h->mg.cap_rows, h->mg.cap_strips, h->mg.cap_jobs);
	chain_arm_kernel<<<512, 256, 0, h->stream>>>(h->mg.block, l.off_events, l.off_queue, nstrips, njobs, h->mg.rank, best_word);
	h->stat_launches++;
	CU(h, cudaGetLastError());
	return 0;
}

// Column chunks of a chained partition.  chunk_cols > 0: block-cyclic chunks of that width (the last one may be
// narrower).  chunk_cols == 0: a width that gives every GPU about 16 chunks (load balance under pruning, short pipeline
// fill) within [32 Ki, 1 Mi] columns.  chunk_cols < 0: one contiguous slice per GPU with the integer arithmetic of the
// reference's --split (C/libmasa/libmasa.cpp:632-635).
void chain_bounds(int n, int world, int chunk_cols, std::vector<int>& b) {
	b.clear();
	if (chunk_cols < 0) {
		for (int r = 0; r <= world; r++) b.push_back((int)((long long)n * r / world));
		// drop empty slices (n < world)
		std::vector<int> u; u.push_back(0);
		for (size_t k = 1; k < b.size(); k++) if (b[k] > u.back()) u.push_back(b[k]);
		b.swap(u);
		return;
	}
	long long w = chunk_cols;
	if (w == 0) {
		w = (long long)n / ((long long)world * 16);
		w = std::max<long long>(32768, std::min<long long>(w, 1 << 20));
		w = (w + 1023) / 1024 * 1024;
	}
	for (long long j = 0; j < n; j += w) b.push_back((int)j);
	b.push_back(n);
}

}  // namespace

extern "C" int b200_chain_plan(const b200_partition* p, int world, b200_chain_info* out) {
	if (!p || !out || world < 1 || world > 8) return 1;
	const int m = p->i1 - p->i0, n = p->j1 - p->j0;
	if (m <= 0 || n <= 0) return 1;
	std::vector<int> b;
	chain_bounds(n, world, p->reserved[1], b);
	memset(out, 0, sizeof(*out));
	out->chunks = (int)b.size() - 1;
	out->chunk_cols = b.size() > 1 ? b[1] - b[0] : n;
	out->chunks_per_gpu = (out->chunks + world - 1) / world;
	out->max_strips = strips_cap_for(m);
	out->max_jobs = (long long)out->chunks_per_gpu * out->max_strips;
	return 0;
}

extern "C" int b200_mgpu_export(b200_handle* h, long long max_rows, long long max_jobs, b200_ipc_handle* out) {
	if (!h) return 1;
	if (!out || max_rows <= 0 || max_jobs <= 0) { h->err = "b200_mgpu_export: bad arguments"; return 1; }
	if (alloc_exchange(h, max_rows, max_jobs)) return 1;
	cudaIpcMemHandle_t ih;
	CU(h, cudaIpcGetMemHandle(&ih, h->mg.block));
	memset(out, 0, sizeof(*out));
	memcpy(out->bytes, &ih, sizeof(ih));
	return 0;
}

extern "C" int b200_mgpu_connect(b200_handle* h, int rank, int world, const b200_ipc_handle* all_handles) {
	if (!h) return 1;
	if (!all_handles || world < 1 || world > 8 || rank < 0 || rank >= world || !h->mg.block) { h->err = "b200_mgpu_connect: bad arguments (world <= 8, export first)"; return 1; }
	CU(h, cudaSetDevice(h->cfg.device));
	for (int r = 0; r < world; r++) {
		h->mg.peers[r] = nullptr;
		if (r == rank) { h->mg.peers[r] = h->mg.block; continue; }
		cudaIpcMemHandle_t ih;
		memcpy(&ih, all_handles[r].bytes, sizeof(ih));
		void* ptr = nullptr;
		CU(h, cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess));
		h->mg.peers[r] = reinterpret_cast<int*>(ptr);
	}
	h->mg.rank = rank; h->mg.world = world; h->mg.connected = true; h->mg.ipc = true; h->mg.epoch = 0;
	if (arm_exchange(h, h->mg.cap_strips, h->mg.cap_jobs, -1)) return 1;
	CU(h, cudaStreamSynchronize(h->stream));
	return 0;
}

extern "C" int b200_mgpu_disconnect(b200_handle* h) {
	if (!h) return 1;
	if (h->mg.connected && h->mg.ipc)
		for (int r = 0; r < h->mg.world; r++)
			if (r != h->mg.rank && h->mg.peers[r]) cudaIpcCloseMemHandle(h->mg.peers[r]);
	h->mg.connected = false;
	if (h->mg.block) { cudaSetDevice(h->cfg.device); cudaFree(h->mg.block); h->mg.block = nullptr; }
	h->mg.strips.release(); h->mg.chunks.release(); h->mg.hrow.release();
	return 0;
}

// One chained alignment over the `nlocal` handles of THIS process (1 with one process per GPU: the other GPUs run the
// same call in their own processes; all of them with b200_group).  Every handle is connected to the same chain and
// holds both sequences.  Callers separate consecutive chained calls by a barrier over all ranks.
static int chain_align(b200_handle* const* hs, int nlocal, const b200_partition* p, const b200_callbacks* cb, b200_result* out) {
	b200_handle* h0 = hs[0];
	memset(out, 0, sizeof(*out));
	const int m = p->i1 - p->i0, n = p->j1 - p->j0;
	const int world = h0->mg.world;
#define FAIL(msg) do { h0->err = (msg); return 1; } while (0)
	for (int q = 0; q < nlocal; q++) {
		b200_handle* h = hs[q];
		if (!h->mg.connected || h->mg.world != world) FAIL("chained alignment: b200_mgpu_connect / b200_group_create was not called on every handle");
		if (m <= 0 || n <= 0 || p->i0 < 0 || p->j0 < 0 || p->i1 > h->n0 || p->j1 > h->n1) FAIL("chained alignment: partition outside the sequences");
		if ((long long)m > h->mg.cap_rows) FAIL("chained alignment: more rows than the exchange block was exported for");
	}
	if (p->reserved[0] & B200_CONT_CHUNK) FAIL("chained alignment: B200_CONT_CHUNK is not supported");
	const bool all_local = nlocal == world;
	const bool sw = p->recurrence == B200_SMITH_WATERMAN;
	const int track = p->want_best_score ? 2 : 0;
	const bool have_cb = cb != nullptr;
	static const bool dbg = getenv("B200_DEBUG") != nullptr;

	// ---- plan: strips (identical everywhere), chunks (owner = index mod world)
	int bh = p->block_height > 0 ? p->block_height : 4 * std::min(128, n);
	std::vector<int> sr_ids;
	if (p->want_special_rows) special_row_ids(m, bh, p->special_row_interval, sr_ids);
	// Strip height of the packed kernel: 1024 rows (16 per virtual lane) is the cheapest per cell, but a front needs about
	// 2400 resident strips per GPU to fill it (148 SMs x 16 warps); when the rows cannot provide that many per GPU, 512-row strips (8 per
	// virtual lane) double the number of strips and halve the dependent chain of a step.  B200_CHAIN_SH overrides.
	int sh16 = kSH16F;
	if (h0->acgt_only && (long long)m / kSH16F < 2400LL * world) sh16 = kSH16;      // fewer 1024-row strips than resident warps
	if (const char* e = getenv("B200_CHAIN_SH")) sh16 = atoi(e) == kSH16 ? kSH16 : kSH16F;
	if (!h0->acgt_only) sh16 = kSH16F;
	std::vector<StripRow> srows;
	bool any_s16 = false;
	build_strips(h0, p, m, sr_ids, srows, any_s16, sh16);
	const int S = (int)srows.size();
	const int kind = any_s16 ? B200_KERNEL_S16X2 : B200_KERNEL_S32;
	std::vector<int> bounds;
	chain_bounds(n, world, p->reserved[1], bounds);
	const int C = (int)bounds.size() - 1;
	if ((long long)S > h0->mg.cap_strips) FAIL("chained alignment: more strips than the exchange block was exported for");
	const int last_owner = (C - 1) % world;
	int chunk_max = 0;
	for (int c = 0; c < C; c++) chunk_max = std::max(chunk_max, bounds[c + 1] - bounds[c]);

	struct Local { std::vector<ChunkCol> chunks; long long cols = 0; long long njobs = 0; std::vector<int> first_h; };
	std::vector<Local> loc(nlocal);
	for (int q = 0; q < nlocal; q++) {
		b200_handle* h = hs[q];
		Local& L = loc[q];
		for (int c = h->mg.rank; c < C; c += world) {
			ChunkCol cc; cc.j0 = p->j0 + bounds[c]; cc.cols = bounds[c + 1] - bounds[c]; cc.cum = (int)L.cols; cc.gidx = c;
			L.chunks.push_back(cc);
			L.cols += cc.cols;
		}
		L.njobs = (long long)L.chunks.size() * S;
		if (L.njobs > h->mg.cap_jobs) FAIL("chained alignment: more jobs than the exchange block was exported for (b200_chain_plan gives the size)");
		if (L.njobs > 0x7fffffffLL) FAIL("chained alignment: too many jobs; use wider chunks");
	}

	// ---- first row / first column from the caller (host side, once)
	Cell corner_col; corner_col.h = 0; corner_col.x = -kInf;
	Cell corner_row = corner_col;
	b200_handle* hr0 = nullptr;                    // the local handle that is rank 0 (owner of chunk 0), if any
	b200_handle* hlast = nullptr;                  // the local handle that owns the last chunk, if any
	for (int q = 0; q < nlocal; q++) { if (hs[q]->mg.rank == 0) hr0 = hs[q]; if (hs[q]->mg.rank == last_owner) hlast = hs[q]; }
	if (have_cb && cb->receive_first_column && hr0) cb->receive_first_column(cb->ctx, reinterpret_cast<b200_cell*>(&corner_col), 1);
	if (have_cb && cb->receive_first_row) cb->receive_first_row(cb->ctx, reinterpret_cast<b200_cell*>(&corner_row), 1);
	Cell first_row_tail = corner_row;
	const bool custom_row = !(p->first_row_init == B200_INIT_ZEROES || !(have_cb && cb->receive_first_row));
	const bool need_rows = have_cb && cb->dispatch_row && (!sr_ids.empty() || p->want_last_row);
	if (custom_row || need_rows) {
		CU(h0, cudaSetDevice(h0->cfg.device));
		CU(h0, h0->mg.hrow.reserve((size_t)n + 8));
	}
	if (custom_row) {
		cb->receive_first_row(cb->ctx, reinterpret_cast<b200_cell*>(h0->mg.hrow.p), n);
		first_row_tail = h0->mg.hrow.p[n - 1];
		if (!sw && kind == B200_KERNEL_S16X2)
			for (int k = 0; k < n; k++)
				if (h0->mg.hrow.p[k].h <= -kInf / 2) FAIL("chained alignment: NW border with -INF in H needs the int32 kernel (create the handles with B200_KERNEL_S32)");
	} else {
		const int type = p->first_row_init == B200_INIT_CUSTOM ? B200_INIT_ZEROES : p->first_row_init;
		first_row_tail.h = type == B200_INIT_ZEROES ? 0 : -kGapExt * n - (type == B200_INIT_GAPS ? kGapOpen : 0);
	}

	// ---- per GPU: buffers, tables, borders, launch
	const bool stream_rows = have_cb && cb->dispatch_row && !sr_ids.empty();
	for (int q = 0; q < nlocal; q++) {
		b200_handle* h = hs[q];
		Local& L = loc[q];
		const int K = (int)L.chunks.size();
		CU(h, cudaSetDevice(h->cfg.device));
		CU(h, h->mg.strips.reserve(S));
		CU(h, h->mg.chunks.reserve(std::max(K, 1)));
		CU(h, h->progress.reserve(S));
		CU(h, h->results.reserve(S));
		CU(h, h->hresults.reserve(S));
		if (reserve_sra(h, sr_ids.size(), (size_t)std::max<long long>(L.cols, 1))) { h0->err = h->err; return 1; }
		if (p->want_last_column && h == hlast) CU(h, h->right.reserve((size_t)m + 1));
		if (reset_scalars(h, INT_MIN)) { h0->err = h->err; return 1; }
		CU(h, cudaMemcpyAsync(h->mg.strips.p, srows.data(), S * sizeof(StripRow), cudaMemcpyHostToDevice, h->stream));
		if (K) CU(h, cudaMemcpyAsync(h->mg.chunks.p, L.chunks.data(), K * sizeof(ChunkCol), cudaMemcpyHostToDevice, h->stream));
		CU(h, cudaMemsetAsync(h->progress.p, 0, S * sizeof(int), h->stream));
		{
			// results start as "none": strips whose jobs all live on other GPUs keep this value
			fill_const_kernel<<<(2 * S + 255) / 256, 256, 0, h->stream>>>(reinterpret_cast<Cell*>(h->results.p), 2LL * S, -kInf, -1);
			h->stat_launches++;
		}
		if (custom_row) {
			for (const ChunkCol& cc : L.chunks)
				CU(h, cudaMemcpyAsync(h->busH.p + cc.j0, h0->mg.hrow.p + (cc.j0 - p->j0), (size_t)cc.cols * sizeof(Cell), cudaMemcpyHostToDevice, h->stream));
		} else {
			const int type = p->first_row_init == B200_INIT_CUSTOM ? B200_INIT_ZEROES : p->first_row_init;
			fill_cells_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->busH.p + p->j0, n, type, 1, 0);
			h->stat_launches++;
		}
		if (h == hr0 && p->first_col_init != B200_INIT_ZEROES) {
			CU(h, h->left.reserve((size_t)m + 1));
			if (have_cb && cb->receive_first_column) {
				CU(h, h->hcells.reserve((size_t)m + 2));
				h->hcells.p[0] = corner_col;
				cb->receive_first_column(cb->ctx, reinterpret_cast<b200_cell*>(h->hcells.p + 1), m);
				CU(h, cudaMemcpyAsync(h->left.p, h->hcells.p, ((size_t)m + 1) * sizeof(Cell), cudaMemcpyHostToDevice, h->stream));
			} else {
				const int type = p->first_col_init == B200_INIT_CUSTOM ? B200_INIT_ZEROES : p->first_col_init;
				fill_cells_kernel<<<(m + 1 + 255) / 256, 256, 0, h->stream>>>(h->left.p, (long long)m + 1, type, 0, 0);
				h->stat_launches++;
			}
		}
		// every allocation happens before the first launch: cudaHostAlloc / cudaMalloc may wait for running kernels, and a
		// persistent kernel that waits for a neighbour which has not been launched yet would never finish
		if (stream_rows && h->sra_flags_cap < sr_ids.size()) {
			if (h->sra_flags) cudaFreeHost(h->sra_flags);
			h->sra_flags = nullptr; h->sra_flags_cap = 0;
			CU(h, cudaHostAlloc((void**)&h->sra_flags, (sr_ids.size() + 64) * sizeof(int), cudaHostAllocMapped));
			h->sra_flags_cap = sr_ids.size() + 64;
		}
		if (h == hlast && have_cb && cb->dispatch_column && p->want_last_column) CU(h, h->hcells.reserve((size_t)m + 2));
		CU(h, cudaStreamSynchronize(h->stream));          // pinned staging is reused below
	}
	for (int q = 0; q < nlocal; q++) {
		b200_handle* h = hs[q];
		Local& L = loc[q];
		CU(h, cudaSetDevice(h->cfg.device));
		if (stream_rows) memset(h->sra_flags, 0, sr_ids.size() * sizeof(int));
		const ExLayout l = ex_layout(h->mg.cap_rows, h->mg.cap_strips, h->mg.cap_jobs);
		const int nxr = (h->mg.rank + 1) % world;
		char* mine = reinterpret_cast<char*>(h->mg.block);
		char* next = reinterpret_cast<char*>(h->mg.peers[nxr]);
		ChainParams& ch = h->ov.chain;
		memset(&ch, 0, sizeof(ch));
		ch.enabled = 1; ch.world = world; ch.nstrips = S; ch.nchunks_local = (int)L.chunks.size(); ch.nchunks_total = C;
		ch.left_zero = p->first_col_init == B200_INIT_ZEROES ? 1 : 0;
		ch.local_cols = L.cols;
		ch.strips = h->mg.strips.p; ch.chunks = h->mg.chunks.p;
		ch.queue = reinterpret_cast<int*>(mine + l.off_queue); ch.q_tail = h->mg.block + kCtlTail;
		ch.events = reinterpret_cast<unsigned long long*>(mine + l.off_events);
		ch.my_cells = reinterpret_cast<const Cell*>(mine + l.off_cells);
		ch.nx_queue = reinterpret_cast<int*>(next + l.off_queue); ch.nx_tail = h->mg.peers[nxr] + kCtlTail;
		ch.nx_events = reinterpret_cast<unsigned long long*>(next + l.off_events);
		ch.nx_cells = reinterpret_cast<Cell*>(next + l.off_cells);
		h->ov.trace = trace_enabled() ? reinterpret_cast<unsigned long long*>(mine + l.off_trace) : nullptr;
		h->ov.nx_trace = trace_enabled() ? reinterpret_cast<unsigned long long*>(next + l.off_trace) : nullptr;
		const int word = kCtlBest + (int)(h->mg.epoch & 1u);
		h->ov.gbest = h->mg.block + word;
		h->ov.npeer = 0;
		for (int r = 0; r < world; r++) if (r != h->mg.rank) h->ov.peer_best[h->ov.npeer++] = h->mg.peers[r] + word;
		h->ov.prune = (p->prune && sw && track == 2 && kind == B200_KERNEL_S16X2) ? 1 : 0;
		h->ov.prune_i1 = p->super_i1 > 0 ? p->super_i1 : p->i1;
		h->ov.prune_j1 = p->super_j1 > 0 ? p->super_j1 : p->j1;
		h->ov.sra_done = stream_rows ? h->sra_flags : nullptr;
		h->ov.mixed = !h->acgt_only;
		h->ov.no_right = !(p->want_last_column && h == hlast);
		h->ov.chunk_cols_max = chunk_max;
		if (dbg) fprintf(stderr, "[b200] chain rank %d/%d: %d strips x %d chunks (of %d, <= %d columns), prune=%d kind=%d\n", h->mg.rank, world, S, (int)L.chunks.size(), C, chunk_max, h->ov.prune, kind);
		int lrc = 0;
		CU(h, cudaEventRecord(h->ev0, h->stream));
		if (L.njobs > 0) lrc = launch_strips(h, (int)L.njobs, p->recurrence, track, kind, sh16, true);
		CU(h, cudaEventRecord(h->ev1, h->stream));
		memset(&ch, 0, sizeof(ch));
		h->ov.trace = nullptr; h->ov.nx_trace = nullptr;
		h->ov.gbest = nullptr; h->ov.npeer = 0; h->ov.prune = 0; h->ov.sra_done = nullptr; h->ov.mixed = false; h->ov.no_right = false; h->ov.chunk_cols_max = 0;
		if (lrc) { h0->err = h->err; return 1; }
	}

	// ---- while the kernels run: stream the special rows out (a row is complete once every local GPU has flagged it)
	// Rows are handed over as: [first-column cell, when rank 0 is local] then the chunks owned by local GPUs in column
	// order -- i.e. the whole row in one piece when all GPUs are local, exactly like the single-GPU path.
	std::vector<int> sr_first_h(sr_ids.size() + 1, 0);           // + the last row
	if (need_rows && hr0 && p->first_col_init != B200_INIT_ZEROES) {
		CU(hr0, cudaSetDevice(hr0->cfg.device));
		for (size_t k = 0; k < sr_ids.size(); k++)
			CU(hr0, cudaMemcpyAsync(&sr_first_h[k], &hr0->left.p[sr_ids[k]].h, sizeof(int), cudaMemcpyDeviceToHost, hr0->copy_stream));
		CU(hr0, cudaMemcpyAsync(&sr_first_h[sr_ids.size()], &hr0->left.p[m].h, sizeof(int), cudaMemcpyDeviceToHost, hr0->copy_stream));
		CU(hr0, cudaStreamSynchronize(hr0->copy_stream));
	}
	// copy row `k` of the local special-rows areas (k < 0: the last row, from busH) into the staging row and dispatch it
	auto dispatch_row = [&](long long k, int row_id, int first_h) -> int {
		for (int q = 0; q < nlocal; q++) {
			b200_handle* h = hs[q];
			CU(h, cudaSetDevice(h->cfg.device));
			for (const ChunkCol& cc : loc[q].chunks) {
				const Cell* src = k >= 0 ? h->sra.p + (size_t)k * (size_t)loc[q].cols + cc.cum : h->busH.p + cc.j0;
				CU(h, cudaMemcpyAsync(h0->mg.hrow.p + (cc.j0 - p->j0), src, (size_t)cc.cols * sizeof(Cell), cudaMemcpyDeviceToHost, h->copy_stream));
			}
		}
		for (int q = 0; q < nlocal; q++) { CU(hs[q], cudaSetDevice(hs[q]->cfg.device)); CU(hs[q], cudaStreamSynchronize(hs[q]->copy_stream)); }
		if (hr0) { b200_cell fc; fc.h = first_h; fc.x = -kInf; cb->dispatch_row(cb->ctx, row_id, &fc, 1); }
		if (all_local) cb->dispatch_row(cb->ctx, row_id, reinterpret_cast<b200_cell*>(h0->mg.hrow.p), n);
		else
			for (int c = 0; c < C; c++)
				for (int q = 0; q < nlocal; q++)
					if (c % world == hs[q]->mg.rank)
						cb->dispatch_row(cb->ctx, row_id, reinterpret_cast<b200_cell*>(h0->mg.hrow.p + bounds[c]), bounds[c + 1] - bounds[c]);
		return 0;
	};
	size_t rows_streamed = 0;
	if (stream_rows) {
		while (rows_streamed < sr_ids.size()) {
			bool ready = true, running = false;
			for (int q = 0; q < nlocal; q++) {
				if (loc[q].chunks.empty()) continue;
				if (!((volatile int*)hs[q]->sra_flags)[rows_streamed]) ready = false;
				cudaSetDevice(hs[q]->cfg.device);
				if (cudaStreamQuery(hs[q]->stream) == cudaErrorNotReady) running = true;
			}
			if (!ready) {
				if (!running) break;                               // kernels over (or failed): the rest is handled below
				struct timespec ts = {0, 20000}; nanosleep(&ts, nullptr);
				continue;
			}
			if (dispatch_row((long long)rows_streamed, p->i0 + sr_ids[rows_streamed], sr_first_h[rows_streamed])) return 1;
			rows_streamed++;
		}
	}

	// ---- completion
	int stop = 0;
	b200_score best; best.score = -kInf; best.i = -1; best.j = -1;
	for (int q = 0; q < nlocal; q++) {
		b200_handle* h = hs[q];
		CU(h, cudaSetDevice(h->cfg.device));
		if (track) CU(h, cudaMemcpyAsync(h->hresults.p, h->results.p, S * sizeof(Score3), cudaMemcpyDeviceToHost, h->stream));
		CU(h, cudaMemcpyAsync(h->hscalars.p, h->scalars.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
	}
	for (int q = 0; q < nlocal; q++) {
		b200_handle* h = hs[q];
		CU(h, cudaSetDevice(h->cfg.device));
		cudaError_t e = cudaStreamSynchronize(h->stream);
		if (e != cudaSuccess) { h0->err = std::string("chained alignment, GPU ") + std::to_string(h->mg.rank) + ": " + cudaGetErrorString(e); return 1; }
		if (h->hscalars.p[2] != 0 && stop == 0) stop = h->hscalars.p[2];
		float ms = 0;
		CU(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
		b200_result& r = h->last_chain;
		memset(&r, 0, sizeof(r));
		r.device_ms = ms; r.strips = S; r.kernel_launches = loc[q].njobs > 0 ? 1 : 0; r.kernel_used = kind;
		r.cells = (long long)*reinterpret_cast<unsigned long long*>(h->hscalars.p + 4);
		r.cells_total = (long long)m * loc[q].cols;
		{
			// share of the resident warps' time spent computing (per mille): the rest is waiting for a neighbour / the queue
			const double busy_ns = (double)*reinterpret_cast<unsigned long long*>(h->hscalars.p + 6);
			const double cap_ns = (double)ms * 1e6 * h->last_grid_warps;
			r.reserved[2] = cap_ns > 0 ? (int)(1000.0 * busy_ns / cap_ns) : 0;
			r.reserved[3] = h->last_grid_warps;
		}
		r.best.score = -kInf; r.best.i = r.best.j = -1;
		if (track)
			for (int k = 0; k < S; k++) {
				const Score3& s = h->hresults.p[k];
				if (s.i >= 0 && (s.score > r.best.score || (s.score == r.best.score && (s.i < r.best.i || (s.i == r.best.i && s.j < r.best.j))))) {
					r.best.score = s.score; r.best.i = s.i; r.best.j = s.j;
				}
			}
		h->stat_cells += r.cells;
		out->cells += r.cells;
		out->device_ms = std::max(out->device_ms, (double)ms);
		out->kernel_launches += r.kernel_launches;
		if (r.best.i >= 0 && (r.best.score > best.score || (r.best.score == best.score && (r.best.i < best.i || (r.best.i == best.i && r.best.j < best.j))))) best = r.best;
		if (trace_enabled()) {
			// development: dump {pushed, popped, first publication, finished} of every job of this GPU (last call wins)
			const ExLayout l = ex_layout(h->mg.cap_rows, h->mg.cap_strips, h->mg.cap_jobs);
			std::vector<unsigned long long> tr((size_t)loc[q].njobs * 4 + 8);
			tr[0] = (unsigned long long)S; tr[1] = loc[q].chunks.size(); tr[2] = (unsigned long long)world; tr[3] = (unsigned long long)h->mg.rank;
			tr[4] = (unsigned long long)C; tr[5] = (unsigned long long)chunk_max; tr[6] = (unsigned long long)(ms * 1e6); tr[7] = 0;
			CU(h, cudaMemcpy(tr.data() + 8, reinterpret_cast<char*>(h->mg.block) + l.off_trace, (size_t)loc[q].njobs * 32, cudaMemcpyDeviceToHost));
			CU(h, cudaMemset(reinterpret_cast<char*>(h->mg.block) + l.off_trace, 0, (size_t)loc[q].njobs * 32));
			std::string fn = std::string(getenv("B200_TRACE_DIR")) + "/trace_rank" + std::to_string(h->mg.rank) + ".bin";
			if (FILE* f = fopen(fn.c_str(), "wb")) { fwrite(tr.data(), 8, tr.size(), f); fclose(f); }
		}
		// re-arm this GPU's exchange block for the next chained call (everything that writes into it has finished: its
		// only producers are the jobs on its left, all consumed; running-best pushes of slower peers go to this call's
		// word, the NEXT call's word is reset here)
		h->mg.epoch++;
		if (arm_exchange(h, S, loc[q].njobs, (int)(h->mg.epoch & 1u))) { h0->err = h->err; return 1; }
		CU(h, cudaStreamSynchronize(h->stream));
	}
	if (stop != 0) { h0->err = "strip kernel watchdog: a border dependency did not advance (code " + std::to_string(stop) + ")"; return 5; }
	out->strips = S; out->kernel_used = kind; out->cells_total = (long long)m * n; out->best = best;
	out->reserved[0] = C; out->reserved[1] = chunk_max; out->reserved[4] = sh16;
	{
		long long busy = 0, warps = 0;
		for (int q = 0; q < nlocal; q++) { busy += (long long)hs[q]->last_chain.reserved[2] * hs[q]->last_chain.reserved[3]; warps += hs[q]->last_chain.reserved[3]; }
		out->reserved[2] = warps ? (int)(busy / warps) : 0; out->reserved[3] = (int)warps;
	}

	// ---- remaining artefacts
	if (have_cb) {
		if (cb->dispatch_row) {
			for (size_t k = rows_streamed; k < sr_ids.size(); k++)
				if (dispatch_row((long long)k, p->i0 + sr_ids[k], sr_first_h[k])) return 1;
			if (p->want_last_row && dispatch_row(-1, p->i1, sr_first_h[sr_ids.size()])) return 1;
		}
		if (cb->dispatch_column && p->want_last_column && hlast) {
			b200_handle* h = hlast;
			CU(h, cudaSetDevice(h->cfg.device));
			CU(h, cudaMemcpy(h->hcells.p, h->right.p, ((size_t)m + 1) * sizeof(Cell), cudaMemcpyDeviceToHost));
			b200_cell fc; fc.h = first_row_tail.h; fc.x = -kInf;
			cb->dispatch_column(cb->ctx, p->j1, &fc, 1);
			for (int r = 0; r < m; r += bh) {
				int len = std::min(bh, m - r);
				cb->dispatch_column(cb->ctx, p->j1, reinterpret_cast<b200_cell*>(h->hcells.p + 1 + r), len);
				if (cb->must_continue && !cb->must_continue(cb->ctx)) break;
			}
		}
		if (cb->dispatch_score && track && best.i >= 0) cb->dispatch_score(cb->ctx, best);
	}
#undef FAIL
	return 0;
}

// ---------------------------------------------------------------------------------------------------------
// in-process group: one host thread drives several GPUs (the multi-GPU mode of build/cudalign)
// ---------------------------------------------------------------------------------------------------------
struct b200_group {
	std::vector<b200_handle*> hs;
	std::string err;
};

extern "C" const char* b200_group_last_error(const b200_group* g) {
	if (!g) return g_create_error.c_str();
	if (!g->err.empty()) return g->err.c_str();
	return g->hs.empty() ? "" : g->hs[0]->err.c_str();
}

extern "C" void b200_group_destroy(b200_group* g) {
	if (!g) return;
	for (b200_handle* h : g->hs) b200_destroy(h);
	delete g;
}

extern "C" int b200_group_create(const int* devices, int n, const b200_config* cfg, long long max_rows, long long max_jobs, b200_group** out) {
	if (!out) return 1;
	*out = nullptr;
	if (!devices || n < 1 || n > 8 || max_rows <= 0 || max_jobs <= 0) { g_create_error = "b200_group_create: bad arguments (1..8 devices)"; return 1; }
	b200_group* g = new b200_group();
	for (int r = 0; r < n; r++) {
		b200_config c;
		memset(&c, 0, sizeof(c));
		if (cfg) c = *cfg;
		c.device = devices[r];
		// test hook: several ranks on ONE device must share its warp slots to be co-resident (tests/test_chain_gpu.py)
		if (const char* e = getenv("B200_GROUP_WARPS_PER_SM")) c.warps_per_sm = atoi(e);
		b200_handle* h = nullptr;
		int rc = b200_create(&c, &h);
		if (rc) { b200_group_destroy(g); return rc; }
		g->hs.push_back(h);
	}
	// peer access between every pair (NVLink / NVSwitch), exchange blocks, chain wiring
	for (int r = 0; r < n; r++) {
		b200_handle* h = g->hs[r];
		cudaSetDevice(h->cfg.device);
		for (int q = 0; q < n; q++) {
			if (q == r || g->hs[q]->cfg.device == h->cfg.device) continue;
			int can = 0;
			cudaDeviceCanAccessPeer(&can, h->cfg.device, g->hs[q]->cfg.device);
			if (!can) { g_create_error = "b200_group_create: no peer access between GPU " + std::to_string(h->cfg.device) + " and GPU " + std::to_string(g->hs[q]->cfg.device); b200_group_destroy(g); return 3; }
			cudaError_t e = cudaDeviceEnablePeerAccess(g->hs[q]->cfg.device, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { g_create_error = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); b200_group_destroy(g); return 3; }
			cudaGetLastError();
		}
		if (alloc_exchange(h, max_rows, max_jobs)) { g_create_error = h->err; b200_group_destroy(g); return 3; }
	}
	for (int r = 0; r < n; r++) {
		b200_handle* h = g->hs[r];
		for (int q = 0; q < n; q++) h->mg.peers[q] = g->hs[q]->mg.block;
		h->mg.rank = r; h->mg.world = n; h->mg.connected = true; h->mg.ipc = false; h->mg.epoch = 0;
		cudaSetDevice(h->cfg.device);
		if (arm_exchange(h, h->mg.cap_strips, h->mg.cap_jobs, -1) || cudaStreamSynchronize(h->stream) != cudaSuccess) { g_create_error = "b200_group_create: cannot initialise the exchange block: " + h->err; b200_group_destroy(g); return 3; }
	}
	*out = g;
	return 0;
}

extern "C" int b200_group_size(const b200_group* g) { return g ? (int)g->hs.size() : 0; }
extern "C" b200_handle* b200_group_handle(b200_group* g, int rank) { return (g && rank >= 0 && rank < (int)g->hs.size()) ? g->hs[rank] : nullptr; }

extern "C" int b200_group_set_sequences(b200_group* g, const char* seq0, int seq0_len, const char* seq1, int seq1_len) {
	if (!g) return 1;
	g->err.clear();
	for (b200_handle* h : g->hs) {
		int rc = b200_set_sequences(h, seq0, seq0_len, seq1, seq1_len);
		if (rc) { g->err = h->err; return rc; }
	}
	return 0;
}

extern "C" int b200_group_align_partition(b200_group* g, const b200_partition* p, const b200_callbacks* cb, b200_result* out) {
	if (!g) return 1;
	g->err.clear();
	if (!p || !out) { g->err = "b200_group_align_partition: bad arguments"; return 1; }
	int rc = chain_align(g->hs.data(), (int)g->hs.size(), p, cb, out);
	if (rc) g->err = g->hs[0]->err;
	return rc;
}

extern "C" int b200_group_rank_result(const b200_group* g, int rank, b200_result* out) {
	if (!g || !out || rank < 0 || rank >= (int)g->hs.size()) return 1;
	*out = g->hs[rank]->last_chain;
	return 0;
}

extern "C" int b200_last_chain_result(const b200_handle* h, b200_result* out) {
	if (!h || !out) return 1;
	*out = h->last_chain;
	return 0;
}

// ---------------------------------------------------------------------------------------------------------
// stage 4: batched Myers-Miller split
// ---------------------------------------------------------------------------------------------------------
namespace {

const int kInvType[3] = {0, 2, 1};      // sw_stage4.cpp:88

struct S4Plan {
	std::vector<S4Half> halves[4];       // [grp*2 + rev]
	std::vector<StripJob> jobs[4];
	std::vector<S4Part> parts;
	long long left_cells = 0;
};

// one half-matrix -> strip jobs (chained when taller than a strip)
void s4_add_half(S4Plan& pl, int g, int SH, int row0, int rows, int col0, int cols, int row_open, int col_open, int corner, long long& left_off_out) {
	S4Half hf;
	memset(&hf, 0, sizeof(hf));
	hf.bus_off = col0; hf.cols = cols; hf.row_open = row_open; hf.left_off = pl.left_cells; hf.rows = rows; hf.col_open = col_open; hf.corner = corner;
	left_off_out = pl.left_cells;
	pl.halves[g].push_back(hf);
	int prev = -1;
	for (int r = 0; r < rows; r += SH) {
		StripJob j;
		memset(&j, 0, sizeof(j));
		j.i0 = row0 + r; j.rows = std::min(SH, rows - r); j.j0 = col0; j.cols = cols;
		j.dep = prev;
		j.flags = 0;
		j.left_off = (int)(pl.left_cells + r);
		j.right_off = -1; j.sra_off = -1;
		prev = (int)pl.jobs[g].size();
		pl.jobs[g].push_back(j);
	}
	pl.left_cells += rows + 1;
}

}  // namespace

extern "C" int b200_stage4_round(b200_handle* h, const b200_xpoint* in, int n, int max_partition, b200_xpoint* out) {
	if (!h) return 1;
	if (!in || !out || n < 1 || max_partition < 1) { h->err = "b200_stage4_round: bad arguments"; return 1; }
	if (h->n0 <= 0 || h->n1 <= 0) { h->err = "b200_stage4_round: call b200_set_sequences first"; return 1; }
	CU(h, cudaSetDevice(h->cfg.device));
	const int L0 = h->n0, L1 = h->n1;
	const int kind = pick_kernel(h, 0);
	const int SH = strip_height(kind, false);
	for (int k = 0; k < n; k++) { out[k].i = out[k].j = out[k].score = 0; out[k].type = -1; }

	// ---- plan (split_thread, sw_stage4.cpp:87-217)
	S4Plan pl;
	for (int k = 1; k < n; k++) {
		const b200_xpoint a = in[k - 1], b = in[k];
		if (a.i < 0 || a.j < 0 || b.i > L0 || b.j > L1 || b.i < a.i || b.j < a.j || a.type < 0 || a.type > 2 || b.type < 0 || b.type > 2) { h->err = "b200_stage4_round: crosspoints outside the sequences or not monotone"; return 1; }
		const int di = b.i - a.i, dj = b.j - a.j;
		if (di == 0 || dj == 0) continue;
		const bool inverse = di < dj;
		S4Part pt;
		memset(&pt, 0, sizeof(pt));
		pt.out_index = k; pt.i0 = a.i; pt.j0 = a.j; pt.score_s = a.score; pt.diff = b.score - a.score;
		if (!inverse) {
			if (!(a.i < b.i - max_partition)) continue;
			const int ts = a.type, te = b.type;
			const int imid0 = di / 2, imid1 = di - imid0;
			pt.transposed = 0; pt.grp = 0; pt.len1 = dj; pt.imid0 = imid0; pt.imid1 = imid1;
			pt.fwd_bus = a.j; pt.rev_bus = L1 - b.j;
			s4_add_half(pl, 0, SH, a.i, imid0, a.j, dj, ts != 1, ts != 2, ts != 0 ? -kInf : 0, pt.fwd_left);
			s4_add_half(pl, 1, SH, L0 - b.i, imid1, L1 - b.j, dj, 1, 1, te != 0 ? -kInf : 0, pt.rev_left);
		} else {
			if (!(a.j < b.j - max_partition)) continue;
			const int ts = kInvType[a.type], te = kInvType[b.type];
			const int imid0 = dj / 2, imid1 = dj - imid0;       // rows of the transposed call = seq1
			pt.transposed = 1; pt.grp = 1; pt.len1 = di; pt.imid0 = imid0; pt.imid1 = imid1;
			pt.fwd_bus = a.i; pt.rev_bus = L0 - b.i;
			s4_add_half(pl, 2, SH, a.j, imid0, a.i, di, ts != 1, ts != 2, ts != 0 ? -kInf : 0, pt.fwd_left);
			s4_add_half(pl, 3, SH, L1 - b.j, imid1, L0 - b.i, di, 1, 1, te != 0 ? -kInf : 0, pt.rev_left);
		}
		pl.parts.push_back(pt);
	}
	const int nparts = (int)pl.parts.size();
	if (nparts == 0) return 0;

	// ---- device state: reversed sequences, four bus arrays, left borders
	if (!h->s4.rev_valid) {
		CU(h, h->s4.s0r.reserve((size_t)L0 + 64));
		CU(h, h->s4.s1r.reserve((size_t)L1 + 64));
		s4_reverse_kernel<<<(L0 + 255) / 256, 256, 0, h->stream>>>(h->s0.p, h->s4.s0r.p, L0);
		s4_reverse_kernel<<<(L1 + 255) / 256, 256, 0, h->stream>>>(h->s1.p, h->s4.s1r.p, L1);
		h->stat_launches += 2;
		h->s4.rev_valid = true;
	}
	CU(h, h->s4.bus[0].reserve((size_t)L1 + 64)); CU(h, h->s4.bus[1].reserve((size_t)L1 + 64));
	CU(h, h->s4.bus[2].reserve((size_t)L0 + 64)); CU(h, h->s4.bus[3].reserve((size_t)L0 + 64));
	CU(h, h->s4.left.reserve((size_t)pl.left_cells + 64));
	size_t nh = 0, nj = 0;
	for (int g = 0; g < 4; g++) { nh += pl.halves[g].size(); nj += pl.jobs[g].size(); }
	CU(h, h->s4.halves.reserve(nh)); CU(h, h->s4.parts.reserve(nparts)); CU(h, h->s4.out.reserve(n));
	CU(h, h->jobs.reserve(nj)); CU(h, h->progress.reserve(nj)); CU(h, h->results.reserve(nj));
	if (reset_scalars(h, -kInf)) return 1;
	CU(h, cudaMemsetAsync(h->scalars.p + 8, 0, 8 * sizeof(int), h->stream));
	CU(h, cudaMemsetAsync(h->progress.p, 0, nj * sizeof(int), h->stream));
	CU(h, cudaMemcpyAsync(h->s4.parts.p, pl.parts.data(), nparts * sizeof(S4Part), cudaMemcpyHostToDevice, h->stream));
	const unsigned char* rows_seq[4] = {h->s0.p, h->s4.s0r.p, h->s1.p, h->s4.s1r.p};
	const unsigned char* cols_seq[4] = {h->s1.p, h->s4.s1r.p, h->s0.p, h->s4.s0r.p};
	size_t hoff = 0, joff = 0;
	for (int g = 0; g < 4; g++) {
		const int ng = (int)pl.halves[g].size(), njg = (int)pl.jobs[g].size();
		if (ng == 0) continue;
		CU(h, cudaMemcpyAsync(h->s4.halves.p + hoff, pl.halves[g].data(), ng * sizeof(S4Half), cudaMemcpyHostToDevice, h->stream));
		CU(h, cudaMemcpyAsync(h->jobs.p + joff, pl.jobs[g].data(), njg * sizeof(StripJob), cudaMemcpyHostToDevice, h->stream));
		s4_fill_kernel<<<ng, 128, 0, h->stream>>>(h->s4.halves.p + hoff, ng, h->s4.bus[g].p, h->s4.left.p);
		h->stat_launches++;
		h->ov.s0 = rows_seq[g]; h->ov.s1 = cols_seq[g]; h->ov.busH = h->s4.bus[g].p;
		h->ov.left = h->s4.left.p; h->ov.job_off = (int)joff; h->ov.counter = h->scalars.p + 8 + g;
		int rc = launch_strips(h, njg, B200_NEEDLEMAN_WUNSCH, 0, kind, SH, false);
		h->ov.s0 = nullptr; h->ov.s1 = nullptr; h->ov.busH = nullptr; h->ov.left = nullptr; h->ov.job_off = 0; h->ov.counter = nullptr;
		if (rc) return 1;
		hoff += ng; joff += njg;
	}
	CU(h, cudaMemsetAsync(h->scalars.p + 3, 0, sizeof(int), h->stream));
	s4_match_kernel<<<(nparts * 32 + 127) / 128, 128, 0, h->stream>>>(h->s4.parts.p, nparts, h->s4.bus[0].p, h->s4.bus[1].p, h->s4.bus[2].p,
	                                                                   h->s4.bus[3].p, h->s4.left.p, h->s4.left.p, h->s4.out.p, h->scalars.p + 3);
	h->stat_launches++;
	std::vector<XPoint> tmp(n);
	CU(h, cudaMemcpyAsync(tmp.data(), h->s4.out.p, n * sizeof(XPoint), cudaMemcpyDeviceToHost, h->stream));
	CU(h, cudaMemcpyAsync(h->hscalars.p, h->scalars.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
	CU(h, cudaStreamSynchronize(h->stream));
	CU(h, cudaGetLastError());
	if (h->hscalars.p[2] != 0) { h->err = "stage 4: strip kernel watchdog"; return 5; }
	if (h->hscalars.p[3] != 0) {
		int e = h->hscalars.p[3];
		h->err = std::string(e > 0 ? "stage 4: Error Match" : "stage 4: NOT FOUND") + " at partition " + std::to_string(e > 0 ? e - 1 : -e - 1);
		return 6;
	}
	for (const S4Part& pt : pl.parts) {
		const XPoint& o = tmp[pt.out_index];
		out[pt.out_index].i = o.i; out[pt.out_index].j = o.j; out[pt.out_index].type = o.type; out[pt.out_index].score = o.score;
	}
	h->stat_cells += (long long)*reinterpret_cast<unsigned long long*>(h->hscalars.p + 4);
	return 0;
}

extern "C" int b200_stage4(b200_handle* h, const b200_xpoint* in, int n, int max_partition, b200_xpoint* out, int cap, int* n_out) {
	if (!h) return 1;
	if (!in || !out || !n_out || n < 1 || cap < n) { h->err = "b200_stage4: bad arguments"; return 1; }
	std::vector<b200_xpoint> cur(in, in + n), mid, merged;
	auto largest = [](const std::vector<b200_xpoint>& v) {      // CrosspointsFile::getLargestPartitionSize (:71-92)
		int mi = 0, mj = 0;
		for (size_t k = 1; k < v.size(); k++) {
			int di = abs(v[k - 1].i - v[k].i), dj = abs(v[k - 1].j - v[k].j);
			if (di != 0 && dj != 0) { mi = std::max(mi, di); mj = std::max(mj, dj); }
		}
		return std::max(mi, mj);
	};
	while (largest(cur) > max_partition) {
		mid.assign(cur.size(), b200_xpoint());
		int rc = b200_stage4_round(h, cur.data(), (int)cur.size(), max_partition, mid.data());
		if (rc) return rc;
		merged.clear();
		merged.push_back(cur[0]);
		bool changed = false;
		for (size_t k = 1; k < cur.size(); k++) {                // merge_partitions (:785-804)
			const bool diff_pos = mid[k].i != cur[k - 1].i || mid[k].j != cur[k - 1].j;
			if (mid[k].type != -1 && diff_pos) { changed = true; merged.push_back(mid[k]); }
			merged.push_back(cur[k]);
		}
		if (!changed) break;                                      // "Didn't reduce partition." (:930-934)
		cur.swap(merged);
	}
	if ((int)cur.size() > cap) { h->err = "b200_stage4: output capacity too small"; return 1; }
	memcpy(out, cur.data(), cur.size() * sizeof(b200_xpoint));
	*n_out = (int)cur.size();
	return 0;
}

// ---------------------------------------------------------------------------------------------------------
// stage 5: batched traceback
// ---------------------------------------------------------------------------------------------------------
extern "C" int b200_stage5(b200_handle* h, const b200_xpoint* pts, int n, unsigned char* ops, long long ops_cap, int* op_len,
                           b200_s5_stats* total) {
	if (!h) return 1;
	if (!pts || !ops || !op_len || !total || n < 1) { h->err = "b200_stage5: bad arguments"; return 1; }
	if (h->n0 <= 0 || h->n1 <= 0) { h->err = "b200_stage5: call b200_set_sequences first"; return 1; }
	CU(h, cudaSetDevice(h->cfg.device));
	const long long need = ((long long)pts[n - 1].i - pts[0].i) + ((long long)pts[n - 1].j - pts[0].j);
	if (ops_cap < need) { h->err = "b200_stage5: ops buffer smaller than (i_end - i_start) + (j_end - j_start)"; return 1; }
	memset(total, 0, sizeof(*total));
	op_len[0] = 0;

	// ---- plan: pure-gap partitions are walked here (sw_stage5.cpp:88-112), the others go to the device in two classes
	std::vector<S5Part> small, big;
	constexpr long long kFlagBudget = 1ll << 30;            // bytes of flag scratch per launch of the global variant
	for (int k = 1; k < n; k++) {
		const b200_xpoint a = pts[k - 1], b = pts[k];
		if (a.i < 0 || a.j < 0 || b.i > h->n0 || b.j > h->n1 || b.i < a.i || b.j < a.j || a.type < 0 || a.type > 2 || b.type < 0 || b.type > 2) {
			h->err = "b200_stage5: crosspoints outside the sequences or not monotone"; return 1;
		}
		const int di = b.i - a.i, dj = b.j - a.j;
		const long long off = ((long long)a.i - pts[0].i) + ((long long)a.j - pts[0].j);
		if (di == 0 || dj == 0) {
			const int len = di + dj;
			memset(ops + off, di == 0 ? 2 : 1, (size_t)len);
			op_len[k] = len;
			int sum = -len * kGapExt;                      // an empty partition still pays the opening, as in the reference
			if (a.type != (di == 0 ? 1 : 2)) { total->gap_open++; sum -= kGapOpen; }
			total->gap_ext += len;
			total->score += sum;
			continue;
		}
		if ((long long)di * dj > kFlagBudget) { h->err = "b200_stage5: partition " + std::to_string(k) + " is too large for a traceback (run stage 4 first)"; return 6; }
		S5Part p;
		memset(&p, 0, sizeof(p));
		p.i0 = a.i; p.j0 = a.j; p.di = di; p.dj = dj; p.ts = a.type; p.te = b.type; p.op_off = off; p.out_index = k;
		(di <= kS5Local && dj <= kS5Local ? small : big).push_back(p);
	}
	const size_t ns = small.size(), nb = big.size();
	if (ns + nb == 0) return 0;

	CU(h, h->s5.ops.reserve((size_t)need + 64));
	CU(h, h->s5.parts.reserve(ns + nb));
	CU(h, h->s5.out.reserve(ns + nb));
	std::vector<S5Out> outs(ns + nb);
	if (ns) {
		CU(h, cudaMemcpyAsync(h->s5.parts.p, small.data(), ns * sizeof(S5Part), cudaMemcpyHostToDevice, h->stream));
		s5_local_kernel<<<(unsigned)((ns + 63) / 64), 64, 0, h->stream>>>(h->s0.p, h->s1.p, h->s5.parts.p, (int)ns, h->s5.ops.p, h->s5.out.p);
		h->stat_launches++;
	}
	// the global variant in batches that fit the flag budget (a batch always takes at least one partition)
	size_t done = 0;
	while (done < nb) {
		long long rows = 0, flags = 0;
		size_t end = done;
		while (end < nb) {
			const long long fb = (long long)big[end].di * big[end].dj;
			if (end > done && flags + fb > kFlagBudget) break;
			big[end].row_off = rows; big[end].flag_off = flags;
			rows += 2ll * (big[end].dj + 1); flags += fb;
			end++;
		}
		CU(h, h->s5.rows.reserve((size_t)rows)); CU(h, h->s5.flags.reserve((size_t)flags));
		CU(h, cudaMemcpyAsync(h->s5.parts.p + ns + done, big.data() + done, (end - done) * sizeof(S5Part), cudaMemcpyHostToDevice, h->stream));
		s5_global_kernel<<<(unsigned)((end - done + 63) / 64), 64, 0, h->stream>>>(h->s0.p, h->s1.p, h->s5.parts.p + ns + done, (int)(end - done), h->s5.ops.p,
		                                                                           h->s5.rows.p, h->s5.flags.p, h->s5.out.p + ns + done);
		h->stat_launches++;
		done = end;
	}
	CU(h, cudaMemcpyAsync(outs.data(), h->s5.out.p, (ns + nb) * sizeof(S5Out), cudaMemcpyDeviceToHost, h->stream));
	CU(h, cudaStreamSynchronize(h->stream));
	CU(h, cudaGetLastError());
	// the steps: one D2H per run of consecutive device partitions (pure-gap partitions in between were written above)
	auto fetch = [&](const std::vector<S5Part>& v, size_t base) -> int {
		for (size_t q = 0; q < v.size(); q++) {
			const S5Out& o = outs[base + q];
			if (o.n_ops < 0 || o.n_ops > v[q].di + v[q].dj) { h->err = "b200_stage5: corrupt walk length"; return 1; }
			op_len[v[q].out_index] = o.n_ops;
			total->matches += o.matches; total->mismatches += o.mismatches; total->gap_open += o.gap_open; total->gap_ext += o.gap_ext;
			total->score += o.score;
			h->stat_cells += (long long)v[q].di * v[q].dj;
		}
		return 0;
	};
	if (fetch(small, 0) || fetch(big, ns)) return 1;
	std::vector<std::pair<long long, long long>> runs;         // [offset, bytes) of device-written slots, merged
	{
		std::vector<const S5Part*> all;
		all.reserve(ns + nb);
		for (const S5Part& p : small) all.push_back(&p);
		for (const S5Part& p : big) all.push_back(&p);
		std::sort(all.begin(), all.end(), [](const S5Part* x, const S5Part* y) { return x->op_off < y->op_off; });
		for (const S5Part* p : all) {
			const long long len = (long long)p->di + p->dj;
			if (!runs.empty() && runs.back().first + runs.back().second == p->op_off) runs.back().second += len;
			else runs.push_back({p->op_off, len});
		}
	}
	for (const auto& r : runs)
		CU(h, cudaMemcpyAsync(ops + r.first, h->s5.ops.p + r.first, (size_t)r.second, cudaMemcpyDeviceToHost, h->stream));
	CU(h, cudaStreamSynchronize(h->stream));
	return 0;
}
