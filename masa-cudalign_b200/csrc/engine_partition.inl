// engine_partition.inl -- part of engine.cu (included there, same translation unit; not compiled on its own).
// b200_align_partition: the whole partition in one persistent launch; strips, borders, special rows streamed while the kernel runs.
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
		B200_LAUNCH(fill_cells_kernel, (n + 255) / 256, 256, h->stream, h->busH.p + p->j0, n, type, 1, 0);
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
			B200_LAUNCH(fill_cells_kernel, (m + 1 + 255) / 256, 256, h->stream, h->left.p, (long long)m + 1, type, row_offset, 0);
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
