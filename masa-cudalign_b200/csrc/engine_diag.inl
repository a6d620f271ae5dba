// engine_diag.inl -- part of engine.cu (included there, same translation unit; not compiled on its own).
// b200_diag_*: one call per CUDAligner virtual (R/src/CUDAligner.hpp:216-232), and the device goal matcher.
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
	B200_LAUNCH(fill_const_kernel, (unsigned)((n + 255) / 256), 256, h->stream, h->busH.p + j0, n, -kInf, -kInf);
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
	B200_LAUNCH(match_column_kernel, (len + 255) / 256, 256, h->stream, dbuf, dbase, len, goal, kGapOpen, h->matchflag.p);
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
