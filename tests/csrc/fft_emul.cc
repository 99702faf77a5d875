// tests/csrc/fft_emul.cc -- TEST ONLY: replays the CUDA MFCC kernel's FFT schedule (csrc/fft_core.h)
// serially on the CPU so its index math can be checked against numpy without a GPU.
#include <cmath>
#include <vector>
#include "fft_core.h"
using namespace srbfft;

extern "C" int emul_power_spectrum(const float *y, int N, int batch, int n_workers, float *P) {
	const int M = N / 2;
	std::vector<cpx> tw(M), tw2(M + 1), A((size_t)batch * M), B((size_t)batch * M);
	for (int j = 0; j < M; ++j) tw[j] = cpx{(float)std::cos(-2.0 * M_PI * j / M), (float)std::sin(-2.0 * M_PI * j / M)};
	for (int k = 0; k <= M; ++k) tw2[k] = cpx{(float)std::cos(-2.0 * M_PI * k / N), (float)std::sin(-2.0 * M_PI * k / N)};
	for (int b = 0; b < batch; ++b)
		for (int n = 0; n < M; ++n) A[(size_t)b * M + n] = cpx{y[(size_t)b * N + 2 * n], y[(size_t)b * N + 2 * n + 1]};
	cpx *in = A.data(), *out = B.data();
	int lgM = 0;
	while ((1 << lgM) < M) ++lgM;
	int n = M, lgs = 0;
	while (n >= 4) {
		for (int w = 0; w < n_workers; ++w) radix4_pass(in, out, lgM, batch, n, lgs, tw.data(), w, n_workers);
		std::swap(in, out);
		n >>= 2;
		lgs += 2;
	}
	if (n == 2) {
		for (int w = 0; w < n_workers; ++w) radix2_tail(in, out, lgM, batch, w, n_workers);
		std::swap(in, out);
	}
	for (int w = 0; w < n_workers; ++w) power_unpack(in, P, M, batch, M + 1, tw2.data(), w, n_workers);
	return num_passes(M);
}

// replay of the fast path (M = 1024, <= 256 non-zero packed inputs): y = real frames [batch][2048] with zeros
// beyond 512; P = power spectrum [batch][FAST_PSTRIDE], natural order
extern "C" void emul_fast_power_spectrum(const float *y, int batch, float *P) {
	const int N = 2048, M = 1024;
	std::vector<cpx> tw1(3 * 256), twA(256), tw2(M + 1);
	for (int r = 1; r < 4; ++r)
		for (int n = 0; n < 256; ++n) tw1[(r - 1) * 256 + n] = cpx{(float)std::cos(-2.0 * M_PI * n * r / 1024), (float)std::sin(-2.0 * M_PI * n * r / 1024)};
	for (int q2 = 0; q2 < 16; ++q2)
		for (int u = 0; u < 16; ++u) twA[q2 * 16 + u] = cpx{(float)std::cos(-2.0 * M_PI * u * q2 / 256), (float)std::sin(-2.0 * M_PI * u * q2 / 256)};
	for (int k = 0; k <= M; ++k) tw2[k] = cpx{(float)std::cos(-2.0 * M_PI * k / N), (float)std::sin(-2.0 * M_PI * k / N)};
	for (int b = 0; b < batch; ++b) {
		std::vector<cpx> z(256), T(4 * 16 * FAST_TSTRIDE), Zs(4 * FAST_ZSTRIDE);
		for (int n = 0; n < 256; ++n) z[n] = cpx{y[(size_t)b * N + 2 * n], y[(size_t)b * N + 2 * n + 1]};
		for (int r = 0; r < 4; ++r)
			for (int u = 0; u < 16; ++u) fast_step_a(z.data(), tw1.data(), twA.data(), T.data() + r * 16 * FAST_TSTRIDE, r, u);
		for (int r = 0; r < 4; ++r)
			for (int u = 0; u < 16; ++u) fast_step_b(T.data() + r * 16 * FAST_TSTRIDE, Zs.data() + r * FAST_ZSTRIDE, u);
		for (int w = 0; w < 64; ++w) fast_power_unpack(Zs.data(), P + (size_t)b * FAST_PSTRIDE, tw2.data(), w);
	}
}
