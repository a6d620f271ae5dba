// engine_stage4.inl -- part of engine.cu (included there, same translation unit; not compiled on its own).
// b200_stage4_round / b200_stage4: batched Myers-Miller split.
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
		B200_LAUNCH(s4_reverse_kernel, (L0 + 255) / 256, 256, h->stream, h->s0.p, h->s4.s0r.p, L0);
		B200_LAUNCH(s4_reverse_kernel, (L1 + 255) / 256, 256, h->stream, h->s1.p, h->s4.s1r.p, L1);
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
		B200_LAUNCH(s4_fill_kernel, ng, 128, h->stream, h->s4.halves.p + hoff, ng, h->s4.bus[g].p, h->s4.left.p);
		h->stat_launches++;
		h->ov.s0 = rows_seq[g]; h->ov.s1 = cols_seq[g]; h->ov.busH = h->s4.bus[g].p;
		h->ov.left = h->s4.left.p; h->ov.job_off = (int)joff; h->ov.counter = h->scalars.p + 8 + g;
		int rc = launch_strips(h, njg, B200_NEEDLEMAN_WUNSCH, 0, kind, SH, false);
		h->ov.s0 = nullptr; h->ov.s1 = nullptr; h->ov.busH = nullptr; h->ov.left = nullptr; h->ov.job_off = 0; h->ov.counter = nullptr;
		if (rc) return 1;
		hoff += ng; joff += njg;
	}
	CU(h, cudaMemsetAsync(h->scalars.p + 3, 0, sizeof(int), h->stream));
	B200_LAUNCH(s4_match_kernel, (nparts * 32 + 127) / 128, 128, h->stream, h->s4.parts.p, nparts, h->s4.bus[0].p, h->s4.bus[1].p, h->s4.bus[2].p,
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
