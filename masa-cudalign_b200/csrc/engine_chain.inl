// engine_chain.inl -- part of engine.cu (included there, same translation unit; not compiled on its own).
// multi-GPU chain: exchange blocks, chunk planning, chained alignment, the in-process group (b200_mgpu_*, b200_chain_plan, b200_group_*).
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
	B200_LAUNCH(chain_arm_kernel, 512, 256, h->stream, h->mg.block, l.off_events, l.off_queue, nstrips, njobs, h->mg.rank, best_word);
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
			B200_LAUNCH(fill_const_kernel, (2 * S + 255) / 256, 256, h->stream, reinterpret_cast<Cell*>(h->results.p), 2LL * S, -kInf, -1);
			h->stat_launches++;
		}
		if (custom_row) {
			for (const ChunkCol& cc : L.chunks)
				CU(h, cudaMemcpyAsync(h->busH.p + cc.j0, h0->mg.hrow.p + (cc.j0 - p->j0), (size_t)cc.cols * sizeof(Cell), cudaMemcpyHostToDevice, h->stream));
		} else {
			const int type = p->first_row_init == B200_INIT_CUSTOM ? B200_INIT_ZEROES : p->first_row_init;
			B200_LAUNCH(fill_cells_kernel, (n + 255) / 256, 256, h->stream, h->busH.p + p->j0, n, type, 1, 0);
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
				B200_LAUNCH(fill_cells_kernel, (m + 1 + 255) / 256, 256, h->stream, h->left.p, (long long)m + 1, type, 0, 0);
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
