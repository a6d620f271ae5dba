// tests/emu/emu_runtime.cpp -- TEST INFRASTRUCTURE: runtime of the CPU SIMT emulation (see tests/emu/cuda_runtime.h).
//
//   * fibers: one per CUDA thread, hand-switched (x86-64), scheduled round-robin inside the OS thread of their CTA;
//     a fiber runs until its next rendezvous (warp collective, __syncthreads), sleep or exit
//   * CTAs: one OS thread per co-resident CTA (up to B200_EMU_MAX_CTAS, default 16), so spin-waits between CTAs and
//     between "devices" make real progress; larger grids are drained block by block by those threads
//   * streams: one worker thread each, tasks in order (copies, memsets, event records, kernel launches)
//   * devices: B200_EMU_DEVICES (default 4) identical "sm_100" devices of B200_EMU_SMS (default 2) SMs sharing the host's
//     memory; peer access always possible; CUDA IPC inside one process, and between processes with B200_EMU_SHM=1
//   * B200_EMU_GUARD=1: guard pages around every device / pinned allocation (out-of-bounds accesses fault)
//   * B200_EMU_POISON=<byte>: fill pattern of fresh allocations (default 0xA5 = a large negative int; 0x3F = a large positive one):
//     results may not depend on it
//   * B200_EMU_SHUFFLE=<seed>: fibers of a CTA are visited in a random order that changes every pass (protocol fuzzing)
#include <cuda_runtime.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <random>
#include <thread>

#include <execinfo.h>
#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/asan_interface.h>
#endif
#include <fcntl.h>
#include <sched.h>
#include <unistd.h>
#include <sys/mman.h>

#if !defined(__x86_64__)
#error "the emulation's fiber switch is written for x86-64"
#endif

extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.hidden emu_switch
.type emu_switch,@function
emu_switch:
	pushq %rbp
	pushq %rbx
	pushq %r12
	pushq %r13
	pushq %r14
	pushq %r15
	movq %rsp, (%rdi)
	movq %rsi, %rsp
	popq %r15
	popq %r14
	popq %r13
	popq %r12
	popq %rbx
	popq %rbp
	ret
.size emu_switch,.-emu_switch
)");

namespace emu {

namespace {

int env_int(const char* name, int dflt) {
	const char* e = getenv(name);
	return e && *e ? atoi(e) : dflt;
}

constexpr size_t kStackBytes = 128 * 1024;
enum : int { RUNNABLE = 0, BLOCKED = 1, DONE = 2, BLOCKED_CTA = 3 };   // BLOCKED: at a warp rendezvous, BLOCKED_CTA: at __syncthreads

struct Warp;
struct Fiber {
	void* sp = nullptr;
	int state = RUNNABLE;
	uint3 tid{0, 0, 0};
	Warp* warp = nullptr;
	int lane = 0;
	unsigned polls = 0;
};
struct Warp {
	unsigned long long buf[2][32];
	int arrived = 0, alive = 0;
	unsigned gen = 0;
	Fiber* lanes[32];
};
struct Block {
	std::vector<Fiber> fibers;
	std::vector<Warp> warps;
	int alive = 0;
	int bar_arrived = 0;
	unsigned bar_gen = 0;
	void* sched_sp = nullptr;
	const std::function<void()>* body = nullptr;
	bool progress = false;
};

thread_local Block* t_blk = nullptr;
thread_local Fiber* t_cur = nullptr;
thread_local char t_anchor;
thread_local int t_device = 0;
thread_local cudaError_t t_last_error = cudaSuccess;

std::atomic<long long> g_s16_overflow{0};

void to_scheduler() { emu_switch(&t_cur->sp, t_blk->sched_sp); }

// A lane that blocks at a warp rendezvous hands the CPU straight to the next runnable lane of its warp (one switch instead of
// two through the scheduler); every 256th time, and always under B200_EMU_SHUFFLE, it goes through the scheduler so that the
// other warps of the CTA get their turn.
thread_local unsigned t_handoffs = 0;
thread_local bool t_no_handoff = false;
void block_in_warp(Fiber* f);

void release_warp(Warp* w) {
	w->arrived = 0;
	w->gen++;
	for (int k = 0; k < 32; k++)
		if (w->lanes[k] && w->lanes[k]->state == BLOCKED) w->lanes[k]->state = RUNNABLE;
}
void release_block(Block* b) {
	b->bar_arrived = 0;
	b->bar_gen++;
	for (auto& f : b->fibers)
		if (f.state == BLOCKED_CTA) f.state = RUNNABLE;
}

void block_in_warp(Fiber* f) {
	if (!t_no_handoff && (++t_handoffs & 255u) != 0) {
		Warp* w = f->warp;
		for (int d = 1; d < 32; d++) {
			Fiber* n = w->lanes[(f->lane + d) & 31];
			if (n && n->state == RUNNABLE) { t_cur = n; emu_switch(&f->sp, n->sp); return; }
		}
	}
	to_scheduler();
}

void fiber_main() {
	Block* b = t_blk;
	(*b->body)();
	Fiber* f = t_cur;
	f->state = DONE;
	b->alive--;
	b->progress = true;
	Warp* w = f->warp;
	w->alive--;
	// a lane that has exited no longer takes part in rendezvous (CUDA semantics of *_sync with exited threads)
	if (w->alive > 0 && w->arrived >= w->alive) release_warp(w);
	if (b->alive > 0 && b->bar_arrived >= b->alive) release_block(b);
	to_scheduler();
	abort();   // never resumed
}

void run_block(unsigned bidx, dim3 grid, dim3 block, const std::function<void()>& body, char* stacks, std::mt19937* rng) {
	const int n = (int)block.x;
#if defined(__SANITIZE_ADDRESS__)
	// fibers never unwind (they end in a switch), so the redzones of their last frames would stay poisoned on a reused stack
	__asan_unpoison_memory_region(stacks, (size_t)n * kStackBytes);
#endif
	Block b;
	b.fibers.resize(n);
	b.warps.resize((n + 31) / 32);
	b.alive = n;
	b.body = &body;
	for (auto& w : b.warps) for (int k = 0; k < 32; k++) w.lanes[k] = nullptr;
	for (int i = 0; i < n; i++) {
		Fiber& f = b.fibers[i];
		f.tid.x = (unsigned)i; f.tid.y = 0; f.tid.z = 0;
		f.warp = &b.warps[i / 32]; f.lane = i & 31;
		f.warp->lanes[f.lane] = &f;
		f.warp->alive++;
		char* top = stacks + (size_t)(i + 1) * kStackBytes;      // 16-byte aligned
		void** sp = reinterpret_cast<void**>(top) - 8;
		for (int k = 0; k < 6; k++) sp[k] = nullptr;             // r15, r14, r13, r12, rbx, rbp
		sp[6] = reinterpret_cast<void*>(&fiber_main);            // "return address" of the first switch
		sp[7] = nullptr;                                         // fake return address of fiber_main (keeps rsp = 8 mod 16 at entry)
		f.sp = sp;
	}
	t_blk = &b;
	t_no_handoff = rng != nullptr;
	t_blockIdx.x = bidx; t_blockIdx.y = 0; t_blockIdx.z = 0;
	t_blockDim.x = block.x; t_blockDim.y = block.y; t_blockDim.z = block.z;
	t_gridDim.x = grid.x; t_gridDim.y = grid.y; t_gridDim.z = grid.z;
	std::vector<int> order(n);
	for (int i = 0; i < n; i++) order[i] = i;
	while (b.alive > 0) {
		b.progress = false;
		if (rng) std::shuffle(order.begin(), order.end(), *rng);
		for (int k = 0; k < n; k++) {
			Fiber* f = &b.fibers[order[k]];
			if (f->state != RUNNABLE) continue;
			t_cur = f;
			emu_switch(&b.sched_sp, f->sp);
		}
		if (!b.progress) sched_yield();      // every runnable fiber is spinning on something another OS thread will provide
	}
	t_blk = nullptr; t_cur = nullptr;
}

void run_grid(dim3 grid, dim3 block, const std::function<void()>& body) {
	const unsigned nblocks = grid.x * grid.y * grid.z;
	if (nblocks == 0 || block.x == 0) return;
	const unsigned nthreads = std::min<unsigned>(nblocks, (unsigned)std::max(1, env_int("B200_EMU_MAX_CTAS", 16)));
	const int seed = env_int("B200_EMU_SHUFFLE", 0);
	std::atomic<unsigned> next{0};
	auto worker = [&](unsigned wid) {
		const size_t bytes = (size_t)block.x * kStackBytes;
		char* stacks = static_cast<char*>(mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
		if (stacks == MAP_FAILED) { perror("emu: mmap of fiber stacks"); abort(); }
		std::mt19937 rng((unsigned)seed * 7919u + wid);
		for (;;) {
			const unsigned bidx = next.fetch_add(1);
			if (bidx >= nblocks) break;
			run_block(bidx, grid, block, body, stacks, seed ? &rng : nullptr);
		}
		munmap(stacks, bytes);
	};
	std::vector<std::thread> th;
	for (unsigned w = 1; w < nthreads; w++) th.emplace_back(worker, w);
	worker(0);
	for (auto& t : th) t.join();
}

double now_ms() {
	timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

}  // namespace

thread_local uint3 t_blockIdx{0, 0, 0}, t_blockDim{1, 1, 1}, t_gridDim{1, 1, 1};

uint3& cur_thread_idx() { return t_cur->tid; }
int lane_id() { return t_cur->lane; }
char* smem_anchor() { return &t_anchor; }
unsigned sm_id_of_block() { return t_blockIdx.x % (unsigned)std::max(1, env_int("B200_EMU_SMS", 2)); }
void note_s16_overflow(int a, int b) {
	const long long k = g_s16_overflow.fetch_add(1, std::memory_order_relaxed);
	if (k < 40 && env_int("B200_EMU_OVERFLOW_TRACE", 0)) {      // where: resolve the addresses with addr2line -e libb200align_emu.so
		void* bt[16];
		const int nbt = backtrace(bt, 16);
		fprintf(stderr, "[emu] s16 overflow #%lld in block %u thread %u: %d + %d\n", k, t_blockIdx.x, t_cur ? t_cur->tid.x : 0u, a, b);
		backtrace_symbols_fd(bt, nbt, 2);
	}
}

const unsigned long long* warp_exchange(unsigned long long v) {
	Fiber* f = t_cur;
	Warp* w = f->warp;
	const unsigned g = w->gen;
	unsigned long long* b = w->buf[g & 1u];
	b[f->lane] = v;
	f->polls = 0;
	t_blk->progress = true;
	if (++w->arrived >= w->alive) { release_warp(w); return b; }
	f->state = BLOCKED;
	block_in_warp(f);
	return b;
}
void block_barrier() {
	Fiber* f = t_cur;
	Block* b = t_blk;
	f->polls = 0;
	b->progress = true;
	if (++b->bar_arrived >= b->alive) { release_block(b); return; }
	f->state = BLOCKED_CTA;
	to_scheduler();
}
void sleep_yield() { to_scheduler(); }
void poll_tick() {
	Fiber* f = t_cur;
	if (f && (++f->polls & 31u) == 0) to_scheduler();
}

// -------------------------------------------------------------------------------------------------- streams
struct Stream {
	std::thread worker;
	std::mutex m;
	std::condition_variable cv, idle_cv;
	std::deque<std::function<void()>> q;
	bool busy = false, stop = false;
	Stream() {
		worker = std::thread([this] {
			std::unique_lock<std::mutex> lk(m);
			for (;;) {
				cv.wait(lk, [this] { return stop || !q.empty(); });
				if (q.empty()) { if (stop) return; continue; }
				std::function<void()> fn = std::move(q.front());
				q.pop_front();
				busy = true;
				lk.unlock();
				fn();
				lk.lock();
				busy = false;
				if (q.empty()) idle_cv.notify_all();
			}
		});
	}
	void push(std::function<void()> fn) {
		{ std::lock_guard<std::mutex> lk(m); q.push_back(std::move(fn)); }
		cv.notify_one();
	}
	void sync() {
		std::unique_lock<std::mutex> lk(m);
		idle_cv.wait(lk, [this] { return q.empty() && !busy; });
	}
	bool idle() {
		std::lock_guard<std::mutex> lk(m);
		return q.empty() && !busy;
	}
	~Stream() {
		{ std::lock_guard<std::mutex> lk(m); stop = true; }
		cv.notify_one();
		worker.join();
	}
};
struct Event { std::atomic<double> t_ms{0.0}; std::atomic<int> pending{0}; };

cudaError_t launch(cudaStream_t s, dim3 grid, dim3 block, std::function<void()> body) {
	auto task = [grid, block, body]() { run_grid(grid, block, body); };
	if (s) s->push(task);
	else task();
	return cudaSuccess;
}

}  // namespace emu

// ------------------------------------------------------------------------------------------------------ API
using emu::env_int;

const char* cudaGetErrorString(cudaError_t e) {
	switch (e) {
	case cudaSuccess: return "no error";
	case cudaErrorInvalidValue: return "invalid argument";
	case cudaErrorMemoryAllocation: return "out of memory";
	case cudaErrorNotReady: return "device not ready";
	case cudaErrorPeerAccessAlreadyEnabled: return "peer access is already enabled";
	case cudaErrorNotSupported: return "operation not supported (SIMT emulation)";
	default: return "unknown error";
	}
}
cudaError_t cudaGetLastError() { cudaError_t e = emu::t_last_error; emu::t_last_error = cudaSuccess; return e; }
cudaError_t cudaGetDeviceCount(int* n) { *n = std::max(1, env_int("B200_EMU_DEVICES", 4)); return cudaSuccess; }
cudaError_t cudaSetDevice(int d) {
	int n; cudaGetDeviceCount(&n);
	if (d < 0 || d >= n) return cudaErrorInvalidValue;
	emu::t_device = d;
	return cudaSuccess;
}
cudaError_t cudaGetDevice(int* d) { *d = emu::t_device; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
	memset(p, 0, sizeof(*p));
	snprintf(p->name, sizeof(p->name), "SIMT emulation of an sm_100 device (tests/emu)");
	p->major = 10; p->minor = 0;
	p->multiProcessorCount = std::max(1, env_int("B200_EMU_SMS", 2));
	p->totalGlobalMem = (size_t)16 << 30;
	return cudaSuccess;
}
cudaError_t cudaMemGetInfo(size_t* free_b, size_t* total_b) { *free_b = (size_t)8 << 30; *total_b = (size_t)16 << 30; return cudaSuccess; }
cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 1; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }

// B200_EMU_GUARD=1: every allocation ends right in front of an inaccessible page (and starts behind one), so a kernel or a
// copy that reads or writes past a buffer dies with SIGSEGV at the faulting instruction instead of touching a neighbour
namespace {
std::mutex g_guard_m;
std::vector<std::tuple<void*, void*, size_t>> g_guarded;     // user pointer, mapping base, mapping length
}
// B200_EMU_SHM=1 (one process per emulated GPU, tests/mgpu_check.py under torchrun with gloo): device allocations live in
// memfd-backed shared mappings, so that cudaIpcOpenMemHandle in ANOTHER process can map them (through /proc/<pid>/fd/<fd>)
// and the chain's peer stores and system-scope atomics cross process borders as they cross NVLink
namespace {
struct ShmAlloc { void* p; size_t len; int fd; bool imported; };
std::vector<ShmAlloc> g_shm;
}
cudaError_t cudaMalloc(void** p, size_t bytes) {
	if (env_int("B200_EMU_SHM", 0)) {
		const size_t len = (std::max<size_t>(bytes, 1) + 4095) & ~(size_t)4095;
		const int fd = memfd_create("b200emu", 0);
		if (fd < 0 || ftruncate(fd, (off_t)len) != 0) return cudaErrorMemoryAllocation;
		void* q = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
		if (q == MAP_FAILED) { close(fd); return cudaErrorMemoryAllocation; }
		memset(q, env_int("B200_EMU_POISON", 0xA5), std::min<size_t>(len, (size_t)64 << 20));
		std::lock_guard<std::mutex> lk(g_guard_m);
		g_shm.push_back({q, len, fd, false});
		*p = q;
		return cudaSuccess;
	}
	if (env_int("B200_EMU_GUARD", 0)) {
		const size_t page = 4096, body = (std::max<size_t>(bytes, 1) + page - 1) & ~(page - 1), len = body + 2 * page;
		char* base = static_cast<char*>(mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
		if (base == MAP_FAILED) return cudaErrorMemoryAllocation;
		mprotect(base, page, PROT_NONE);
		mprotect(base + page + body, page, PROT_NONE);
		char* user = base + page + body - ((bytes + 15) & ~(size_t)15);        // 16-byte aligned, ends at the guard page
		memset(base + page, env_int("B200_EMU_POISON", 0xA5), body);
		std::lock_guard<std::mutex> lk(g_guard_m);
		g_guarded.emplace_back(user, base, len);
		*p = user;
		return cudaSuccess;
	}
	const size_t sz = (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255;
	*p = aligned_alloc(256, sz);
	if (!*p) return cudaErrorMemoryAllocation;
	memset(*p, env_int("B200_EMU_POISON", 0xA5), std::min<size_t>(sz, (size_t)64 << 20));     // device memory is not zero-initialised: make reliance on it visible
	return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
	if (!p) return cudaSuccess;
	{
		std::lock_guard<std::mutex> lk(g_guard_m);
		for (size_t k = 0; k < g_shm.size(); k++)
			if (g_shm[k].p == p) {
				munmap(g_shm[k].p, g_shm[k].len);
				if (g_shm[k].fd >= 0) close(g_shm[k].fd);
				g_shm.erase(g_shm.begin() + k);
				return cudaSuccess;
			}
		for (size_t k = 0; k < g_guarded.size(); k++)
			if (std::get<0>(g_guarded[k]) == p) {
				munmap(std::get<1>(g_guarded[k]), std::get<2>(g_guarded[k]));
				g_guarded.erase(g_guarded.begin() + k);
				return cudaSuccess;
			}
	}
	free(p);
	return cudaSuccess;
}
cudaError_t cudaMallocHost(void** p, size_t bytes) { return cudaMalloc(p, bytes); }
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { return cudaMalloc(p, bytes); }
cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }

cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) { memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t s) {
	if (!s) { memmove(dst, src, bytes); return cudaSuccess; }
	s->push([dst, src, bytes] { memmove(dst, src, bytes); });
	return cudaSuccess;
}
cudaError_t cudaMemset(void* dst, int v, size_t bytes) { memset(dst, v, bytes); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* dst, int v, size_t bytes, cudaStream_t s) {
	if (!s) { memset(dst, v, bytes); return cudaSuccess; }
	s->push([dst, v, bytes] { memset(dst, v, bytes); });
	return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new emu::Stream(); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { if (s) { s->sync(); delete s; } return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t s) { if (s) s->sync(); return cudaSuccess; }
cudaError_t cudaStreamQuery(cudaStream_t s) { return (!s || s->idle()) ? cudaSuccess : cudaErrorNotReady; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emu::Event(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
	e->pending.fetch_add(1);
	auto stamp = [e] { e->t_ms.store(emu::now_ms()); e->pending.fetch_sub(1); };
	if (s) s->push(stamp); else stamp();
	return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t e) { while (e->pending.load() > 0) sched_yield(); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
	if (a->pending.load() > 0 || b->pending.load() > 0) return cudaErrorNotReady;
	*ms = (float)(b->t_ms.load() - a->t_ms.load());
	return cudaSuccess;
}
// CUDA IPC: a handle is {pointer, pid, memfd, length}.  Inside the exporting process it resolves to the pointer itself (the
// self-chain tests, where a GPU is its own neighbour); from another process it maps the exporter's memfd (B200_EMU_SHM=1).
namespace {
struct IpcHandle { void* p; long long pid; long long len; int fd; };
static_assert(sizeof(IpcHandle) <= sizeof(cudaIpcMemHandle_t), "ipc handle");
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
	memset(h, 0, sizeof(*h));
	IpcHandle ih; ih.p = p; ih.pid = (long long)getpid(); ih.len = 0; ih.fd = -1;
	{
		std::lock_guard<std::mutex> lk(g_guard_m);
		for (auto& a : g_shm) if (a.p == p && !a.imported) { ih.len = (long long)a.len; ih.fd = a.fd; }
	}
	memcpy(h->reserved, &ih, sizeof(ih));
	return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
	IpcHandle ih;
	memcpy(&ih, h.reserved, sizeof(ih));
	if (ih.pid == (long long)getpid()) { *p = ih.p; return cudaSuccess; }
	if (ih.fd < 0) return cudaErrorNotSupported;
	char path[64];
	snprintf(path, sizeof(path), "/proc/%lld/fd/%d", ih.pid, ih.fd);
	const int fd = open(path, O_RDWR);
	if (fd < 0) return cudaErrorInvalidValue;
	void* q = mmap(nullptr, (size_t)ih.len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	close(fd);
	if (q == MAP_FAILED) return cudaErrorMemoryAllocation;
	std::lock_guard<std::mutex> lk(g_guard_m);
	g_shm.push_back({q, (size_t)ih.len, -1, true});
	*p = q;
	return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void* p) {
	std::lock_guard<std::mutex> lk(g_guard_m);
	for (size_t k = 0; k < g_shm.size(); k++)
		if (g_shm[k].p == p && g_shm[k].imported) { munmap(p, g_shm[k].len); g_shm.erase(g_shm.begin() + k); break; }
	return cudaSuccess;
}

// test hooks (not part of the C ABI of the product library)
extern "C" long long b200_emu_s16_overflows(void) { return emu::g_s16_overflow.load(); }
extern "C" int b200_emu_is_emulation(void) { return 1; }
