// engine_sequences.inl -- part of engine.cu (included there, same translation unit; not compiled on its own).
// b200_set_sequences / b200_unset_sequences: 2-bit packing on the host, packed upload, byte view rebuilt on the device.
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
		if (seq0_len) B200_LAUNCH(unpack2_kernel, (seq0_len + 255) / 256, 256, h->stream, h->s0p.p, h->s0.p, seq0_len);
		if (seq1_len) B200_LAUNCH(unpack2_kernel, (seq1_len + 255) / 256, 256, h->stream, h->s1p.p, h->s1.p, seq1_len);
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
