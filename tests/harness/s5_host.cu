// tests/harness/s5_host.cu -- TEST INFRASTRUCTURE.  Host build of the stage-5 device function (csrc/stage5.cuh,
// s5_partition is __host__ __device__): tests/test_stage5_host_cpu.py compiles this file with nvcc into a temporary
// shared library and walks the reference's golden crosspoints with it on the CPU, so that the traceback logic the GPU
// threads execute is checked in the container that has no GPU.  Mirrors the planning of b200_stage5 (engine.cu) for
// the non-degenerate partitions only.
#include <vector>
#include "../../masa-cudalign_b200/csrc/stage4.cuh"
#include "../../masa-cudalign_b200/csrc/stage5.cuh"

using namespace b200;

extern "C" int s5_host_walk(const unsigned char* seq0, const unsigned char* seq1, const XPoint* pts, int n, unsigned char* ops,
                            int* op_len, int* totals /* score, matches, mismatches, gap_open, gap_ext */, int force_global) {
	for (int k = 1; k < n; k++) {
		const XPoint a = pts[k - 1], b = pts[k];
		const int di = b.i - a.i, dj = b.j - a.j;
		if (di <= 0 || dj <= 0) { op_len[k] = -1; continue; }          // pure-gap partitions never reach the device
		S5Part p = {};
		p.i0 = a.i; p.j0 = a.j; p.di = di; p.dj = dj; p.ts = a.type; p.te = b.type;
		p.op_off = ((long long)a.i - pts[0].i) + ((long long)a.j - pts[0].j);
		S5Out o;
		if (!force_global && di <= kS5Local && dj <= kS5Local) {
			int hrow[kS5Local + 1], erow[kS5Local + 1];
			unsigned char fl[kS5Local * kS5Local];
			s5_partition(seq0, seq1, p, hrow, erow, fl, kS5Local, ops + p.op_off, o);
		} else {
			std::vector<int> rows(2 * (dj + 1));
			std::vector<unsigned char> fl((size_t)di * dj);
			s5_partition(seq0, seq1, p, rows.data(), rows.data() + dj + 1, fl.data(), dj, ops + p.op_off, o);
		}
		op_len[k] = o.n_ops;
		totals[0] += o.score; totals[1] += o.matches; totals[2] += o.mismatches; totals[3] += o.gap_open; totals[4] += o.gap_ext;
	}
	return 0;
}
