// sra_prefault.cpp -- RAM special rows without page-fault stalls on the dispatch path (build/cudalign only).
//
// MASA-Core stores a dispatched special row in a buffer it malloc()s when the row's first cell arrives
// (SpecialRowRAM::initialize, C/common/sra/SpecialRowRAM.cpp:68-86) and fills with one memcpy
// (SpecialRowRAM::write, :88-100).  A row of a 5M-column comparison is 40 MB of untouched memory: the memcpy runs at
// page-fault speed (measured on the cfg2 pair with --ram-size=8G: 26 ms per row, 212 rows = 5.6 s, while the stage-1
// kernel that produces them needs 3.6 s -- the single host thread that files the rows was the bottleneck of stage 1).
// This object replaces ONLY the allocation, at link time like stage4_gpu.cpp / stage5_gpu.cpp (MASA-Core's sources stay
// untouched; the row object, its write/read/free and every byte stored are the reference's): a helper thread keeps a few
// buffers of the current row length allocated, huge-page advised and already touched, so that initialize() just takes
// one.  Buffers come from posix_memalign and are released by the reference's own free() in ~SpecialRowRAM.
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#include "common/sra/SpecialRowRAM.hpp"

namespace {

const size_t kHuge = 2u << 20;
const size_t kPoolBytes = 1u << 30;          // at most this much memory prepared ahead
const int kPoolMax = 8;

pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
pthread_cond_t g_cv = PTHREAD_COND_INITIALIZER;
pthread_t g_thread;
bool g_started = false;
size_t g_bytes = 0;                          // buffer size the pool currently prepares
int g_want = 0;                              // buffers to keep ready
void* g_ready[kPoolMax];
int g_nready = 0;
long g_hits = 0, g_misses = 0;               // B200_DEBUG=1 prints them at exit

void report() { fprintf(stderr, "[sra_prefault] %ld special-row buffers taken ready-made, %ld allocated on the spot\n", g_hits, g_misses); }

void* fresh_buffer(size_t bytes) {
	void* p = NULL;
	if (bytes >= 2 * kHuge) {
		const size_t rounded = (bytes + kHuge - 1) / kHuge * kHuge;
		if (posix_memalign(&p, kHuge, rounded) != 0) return NULL;
#ifdef MADV_HUGEPAGE
		madvise(p, rounded, MADV_HUGEPAGE);                 // no-op where transparent huge pages are off
#endif
		return p;
	}
	return malloc(bytes);
}

void touch(void* p, size_t bytes) {
#ifdef MADV_POPULATE_WRITE
	if (bytes >= 2 * kHuge && madvise(p, bytes, MADV_POPULATE_WRITE) == 0) return;
#endif
	volatile char* c = (volatile char*)p;
	for (size_t k = 0; k < bytes; k += 4096) c[k] = 0;
}

void* pool_main(void*) {
	pthread_mutex_lock(&g_mu);
	for (;;) {
		while (g_nready >= g_want) pthread_cond_wait(&g_cv, &g_mu);
		const size_t bytes = g_bytes;
		pthread_mutex_unlock(&g_mu);
		void* p = fresh_buffer(bytes);
		if (p != NULL) touch(p, bytes);
		pthread_mutex_lock(&g_mu);
		if (p == NULL) { g_want = 0; continue; }               // out of memory: stop preparing, initialize() reports it
		if (bytes == g_bytes && g_nready < kPoolMax) g_ready[g_nready++] = p;
		else free(p);                                          // the row length changed meanwhile (next partition)
		pthread_cond_broadcast(&g_cv);
	}
	return NULL;
}

void* take_buffer(size_t bytes) {
	if (bytes < 2 * kHuge) return malloc(bytes);               // small rows (stage 2/3 partitions): nothing to gain
	static const bool off = getenv("B200_SRA_PREFAULT") != NULL && atoi(getenv("B200_SRA_PREFAULT")) == 0;
	if (off) return malloc(bytes);                             // A/B switch: the reference's plain allocation
	void* p = NULL;
	pthread_mutex_lock(&g_mu);
	if (bytes != g_bytes) {                                    // new row length: drop what was prepared for the old one
		for (int k = 0; k < g_nready; k++) free(g_ready[k]);
		g_nready = 0;
		g_bytes = bytes;
		g_want = (int)(kPoolBytes / bytes);
		if (g_want < 1) g_want = 1;
		if (g_want > kPoolMax) g_want = kPoolMax;
	}
	if (!g_started) {
		g_started = pthread_create(&g_thread, NULL, pool_main, NULL) == 0;
		if (g_started) pthread_detach(g_thread);
		if (getenv("B200_DEBUG") != NULL) atexit(report);
	}
	if (g_nready > 0) { p = g_ready[--g_nready]; g_hits++; } else g_misses++;
	pthread_cond_broadcast(&g_cv);
	pthread_mutex_unlock(&g_mu);
	if (p == NULL) p = fresh_buffer(bytes);                    // pool not ready yet: allocate here, untouched, like the reference
	return p;
}

}  // namespace

void SpecialRowRAM::initialize(bool readOnly, int length) {
	if (row == NULL) {
		this->length = length == 0 ? 1048 * 1048 : length;     // INITIAL_LENGTH (:27)
		row = (cell_t*)take_buffer((size_t)this->length * sizeof(cell_t));
		if (row == NULL) {
			fprintf(stderr, "Out of memory (special row of %d cells)\n", this->length);
			exit(1);
		}
	}
}
