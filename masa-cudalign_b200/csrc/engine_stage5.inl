// engine_stage5.inl -- part of engine.cu (included there, same translation unit; not compiled on its own).
// b200_stage5: batched traceback.
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
		B200_LAUNCH(s5_local_kernel, (unsigned)((ns + 63) / 64), 64, h->stream, h->s0.p, h->s1.p, h->s5.parts.p, (int)ns, h->s5.ops.p, h->s5.out.p);
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
		B200_LAUNCH(s5_global_kernel, (unsigned)((end - done + 63) / 64), 64, h->stream, h->s0.p, h->s1.p, h->s5.parts.p + ns + done, (int)(end - done), h->s5.ops.p,
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
