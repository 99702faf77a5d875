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

// replay of the warp-per-frame path (mfcc_frames_warp_kernel): 32 lanes = 2 halves x 16, two rounds of sub-transforms, swizzled
// transpose tiles, in-place step B, paired power unpack; P = power spectrum [batch][1025] natural order (de-skewed here)
extern "C" void emul_warp_power_spectrum(const float *y, int batch, float *P) {
	const int N = 2048, M = 1024;
	std::vector<cpx> wc(4 * 16), wt(64 * 16), tw2(513);
	for (int r = 0; r < 4; ++r)
		for (int n2 = 0; n2 < 16; ++n2) wc[r * 16 + n2] = cpx{(float)std::cos(-2.0 * M_PI * n2 * r / 64), (float)std::sin(-2.0 * M_PI * n2 * r / 64)};
	for (int j = 0; j < 64; ++j)
		for (int u = 0; u < 16; ++u) wt[j * 16 + u] = cpx{(float)std::cos(-2.0 * M_PI * u * j / 1024), (float)std::sin(-2.0 * M_PI * u * j / 1024)};
	for (int k = 0; k <= 512; ++k) tw2[k] = cpx{(float)std::cos(-2.0 * M_PI * k / N), (float)std::sin(-2.0 * M_PI * k / N)};
	(void)M;
	for (int b = 0; b < batch; ++b) {
		std::vector<cpx> ybuf(256), TZ(4 * WF_TILE);
		std::vector<float> Ps(WF_PLEN, -1.f);
		for (int n = 0; n < 256; ++n) ybuf[n] = cpx{y[(size_t)b * N + 2 * n], y[(size_t)b * N + 2 * n + 1]};
		for (int round = 0; round < 2; ++round)
			for (int lane = 0; lane < 32; ++lane) {
				const int h = lane >> 4, u = lane & 15, r = h + 2 * round;
				cpx z[16];
				for (int n2 = 0; n2 < 16; ++n2) z[n2] = ybuf[u + 16 * n2];
				wf_step_a(z, wc.data() + r * 16, wt.data() + r * 16, TZ.data() + r * WF_TILE, u);
			}
		for (int round = 0; round < 2; ++round) {
			cpx regs[32][16];
			for (int lane = 0; lane < 32; ++lane) wf_step_b_load(TZ.data() + ((lane >> 4) + 2 * round) * WF_TILE, lane & 15, regs[lane]);
			for (int lane = 0; lane < 32; ++lane) wf_step_b_store(regs[lane], TZ.data(), (lane >> 4) + 2 * round, lane & 15);
		}
		for (int lane = 0; lane < 32; ++lane) wf_power_lane(TZ.data(), Ps.data(), tw2.data(), lane);
		for (int k = 0; k <= 1024; ++k) P[(size_t)b * 1025 + k] = Ps[wf_p_index(k)];
	}
}

// the layout functions themselves, for the shared-memory wavefront count of tests/test_fft_schedule_cpu.py
extern "C" int emul_wf_t_index(int row, int col) { return wf_t_index(row, col); }
extern "C" int emul_wf_z_index(int r, int q) { return wf_z_index(r, q); }
extern "C" int emul_wf_tile() { return WF_TILE; }
