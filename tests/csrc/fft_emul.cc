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
