// fft_core.h -- index math and butterflies of the in-shared-memory FFT used by the fused MFCC kernel.
//
// Written as host/device inline functions parameterised by (worker, n_workers) so that exactly the
// same schedule the CUDA threads execute can be replayed serially on the CPU by the unit tests
// (tests/csrc/fft_emul.cc; the product library itself contains no host compute path).
//
// Transform: the reference takes numpy.fft.fft(frame, FFT_SIZE)[:FFT_SIZE/2+1] of a REAL frame
// (src/feature/MFCC.py:66).  With N = FFT_SIZE and M = N/2 we pack z[n] = y[2n] + i*y[2n+1], run one
// complex M-point forward FFT (Stockham autosort, radix-4 stages plus a radix-2 tail when log2(M) is
// odd; natural-order output), and unpack Y[k], k = 0..M, with the split-radix identity
//     Y[k] = (Z[k] + conj(Z[M-k]))/2  -  (i/2) * e^{-2 pi i k/N} * (Z[k] - conj(Z[M-k])),   Z[M] := Z[0].
#pragma once

#if defined(__CUDACC__)
#define SRB_HD __host__ __device__ __forceinline__
#else
#define SRB_HD inline
#endif

namespace srbfft {

struct alignas(8) cpx {
	float x, y;
};

SRB_HD cpx cadd(cpx a, cpx b) { return cpx{a.x + b.x, a.y + b.y}; }
SRB_HD cpx csub(cpx a, cpx b) { return cpx{a.x - b.x, a.y - b.y}; }
SRB_HD cpx cmul(cpx a, cpx b) { return cpx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
SRB_HD cpx cmulj_neg(cpx a) { return cpx{a.y, -a.x}; }   // a * (-i)

// One Stockham radix-4 pass over `batch` independent M-point transforms laid out [batch][M].
//   n = current sub-transform length (M, M/4, ...), s = stride (1, 4, ...), tw[j] = exp(-2 pi i j / M), j < M.
// Work item t in [0, batch*M/4): transform b = t / (M/4), butterfly bt = t % (M/4), p = bt / s, q = bt % s.
// M, n, s are powers of two (lgM = log2 M, lgs = log2 s), so the index split is shifts and masks.
SRB_HD void radix4_pass(const cpx *in, cpx *out, int lgM, int batch, int n, int lgs, const cpx *tw, int worker, int n_workers) {
	const int M = 1 << lgM, s = 1 << lgs;
	const int quarter = M >> 2;
	const int n1 = n >> 2;
	const int total = batch * quarter;
	for (int t = worker; t < total; t += n_workers) {
		const int b = t >> (lgM - 2);
		const int bt = t & (quarter - 1);
		const int p = bt >> lgs;
		const int q = bt & (s - 1);
		const cpx *x = in + (long)b * M;
		cpx *y = out + (long)b * M;
		const cpx a = x[q + s * (p)];
		const cpx bb = x[q + s * (p + n1)];
		const cpx c = x[q + s * (p + 2 * n1)];
		const cpx d = x[q + s * (p + 3 * n1)];
		const cpx apc = cadd(a, c), amc = csub(a, c);
		const cpx bpd = cadd(bb, d);
		const cpx jbmd = cmulj_neg(csub(bb, d));          // -i (b - d)
		const int tj = p * s;                               // twiddle index: exp(-2 pi i p / n) = tw[p * (M/n)] and M/n == s
		const cpx w1 = tw[tj];
		const cpx w2 = tw[(2 * tj) & (M - 1)];
		const cpx w3 = tw[(3 * tj) & (M - 1)];
		y[q + s * (4 * p + 0)] = cadd(apc, bpd);
		y[q + s * (4 * p + 1)] = cmul(w1, cadd(amc, jbmd));
		y[q + s * (4 * p + 2)] = cmul(w2, csub(apc, bpd));
		y[q + s * (4 * p + 3)] = cmul(w3, csub(amc, jbmd));
	}
}

// Final radix-2 pass (n == 2, s == M/2): no twiddles.
SRB_HD void radix2_tail(const cpx *in, cpx *out, int lgM, int batch, int worker, int n_workers) {
	const int M = 1 << lgM;
	const int half = M >> 1;
	const int total = batch * half;
	for (int t = worker; t < total; t += n_workers) {
		const int b = t >> (lgM - 1);
		const int q = t & (half - 1);
		const cpx *x = in + (long)b * M;
		cpx *y = out + (long)b * M;
		const cpx a = x[q], bb = x[q + half];
		y[q] = cadd(a, bb);
		y[q + half] = csub(a, bb);
	}
}

// number of passes for an M-point transform (M a power of two >= 2)
SRB_HD int num_passes(int M) {
	int lg = 0;
	while ((1 << lg) < M) ++lg;
	return (lg >> 1) + (lg & 1);
}

// Unpack the power spectrum of the real N = 2M point transform from the packed complex result Z:
// P[k] = |Y[k]|^2, k = 0..M, for `batch` transforms; tw2[k] = exp(-2 pi i k / N), k = 0..M.
SRB_HD void power_unpack(const cpx *Z, float *P, int M, int batch, int p_stride, const cpx *tw2, int worker, int n_workers) {
	for (int b = 0; b < batch; ++b)
	for (int k = worker; k <= M; k += n_workers) {
		const cpx *z = Z + (long)b * M;
		const cpx zk = z[k & (M - 1)];
		cpx zc = z[(M - k) & (M - 1)];
		zc.y = -zc.y;
		const cpx e = cadd(zk, zc);          // Z[k] + conj(Z[M-k])
		const cpx o = csub(zk, zc);          // Z[k] - conj(Z[M-k])
		const cpx wo = cmul(tw2[k], o);
		// Y = e/2 - (i/2) * wo = (e.x + wo.y, e.y - wo.x) / 2
		const float yr = 0.5f * (e.x + wo.y);
		const float yi = 0.5f * (e.y - wo.x);
		P[(long)b * p_stride + k] = yr * yr + yi * yi;
	}
}

}  // namespace srbfft
