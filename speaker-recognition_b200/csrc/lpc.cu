// lpc.cu -- per-frame linear-prediction coefficients (the LPC half of the reference's mix_feature).
//
// Replaces src/feature/LPC.py:40-57: for every frame (same window and pre-emphasis as the MFCC front-end, :49-53)
// the reference calls scikits.talkbox.linpred.levinson_lpc.lpc(frame, n_lpc) and keeps coefficients 1..n_lpc (:41-42).
// talkbox (third-party, not in the repo, unpinned) implements the autocorrelation method: biased autocorrelation of
// the frame followed by the Levinson-Durbin recursion; the in-repo prototype src/lpcc/lpc_jp.py:13-66 (autocorr +
// LevinsonDurbin) is the same algorithm and is what the fixtures are generated from.
//
// One warp per frame, float64 throughout (Levinson-Durbin amplifies input error by the condition number of the
// autocorrelation matrix, ~1e3..1e5 for voiced speech, so fp32 would not reach the 1e-4 parity bound):
//   lanes build the windowed, pre-emphasised frame in shared memory, lane l accumulates lag l, lane 0 runs the recursion.
#include "kernels.h"

namespace srb {

constexpr int LPC_WARPS = 4;
constexpr int LPC_MAX_ORDER = 31;

struct LpcArgs {
	const void *pcm;
	int is_f32;
	const long long *smp_off;
	const int *frame_off;
	int n_utt, frames_total;
	int L, shift, order;
	const double *win, *win2;
	float *out;
	int ld, col_off;
};

__device__ __forceinline__ int lpc_find_segment(const int *__restrict__ start, int n, int x) {
	int lo = 0, hi = n - 1;
	while (lo < hi) {
		const int mid = (lo + hi + 1) >> 1;
		if (__ldg(start + mid) <= x) lo = mid;
		else hi = mid - 1;
	}
	return lo;
}

__global__ void __launch_bounds__(LPC_WARPS * 32) lpc_frames_kernel(const LpcArgs A) {
	extern __shared__ __align__(16) unsigned char lpc_smem[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	double *y = reinterpret_cast<double *>(lpc_smem) + (size_t)warp * (A.L + 2 * (LPC_MAX_ORDER + 1));
	double *r = y + A.L;                       // [order+1] autocorrelation
	double *a = r + (LPC_MAX_ORDER + 1);       // [order+1] coefficients
	const int g = blockIdx.x * LPC_WARPS + warp;
	if (g >= A.frames_total) return;           // whole warp leaves together
	const int u = lpc_find_segment(A.frame_off, A.n_utt, g);
	const int t = g - __ldg(A.frame_off + u);
	const long long s0 = A.smp_off[u] + (long long)t * A.shift;
	const int L = A.L, order = A.order;

	for (int i = lane; i < L; i += 32) {
		double sc, sp = 0.0;
		if (A.is_f32) {
			const float *p = reinterpret_cast<const float *>(A.pcm) + s0;
			sc = (double)__ldg(p + i);
			if (i > 0) sp = (double)__ldg(p + i - 1);
		} else {
			const int16_t *p = reinterpret_cast<const int16_t *>(A.pcm) + s0;
			sc = (double)__ldg(p + i);
			if (i > 0) sp = (double)__ldg(p + i - 1);
		}
		// window, then pre-emphasis on the windowed frame (LPC.py:49-53, same as MFCC.py:61-64)
		y[i] = __ldg(A.win + i) * sc - __ldg(A.win2 + i) * sp;
	}
	__syncwarp();
	if (lane <= order) {
		double acc = 0.0;
		const int n1 = L - lane;
		for (int n = 0; n < n1; ++n) acc = fma(y[n], y[n + lane], acc);
		r[lane] = acc;
	}
	__syncwarp();
	if (lane == 0) {
		// Levinson-Durbin (talkbox levinson_1d / lpc_jp.py:27-66): a[0] = 1, e = r[0]; for i = 1..order:
		//   k = -(r[i] + sum_{j<i} a[j] r[i-j]) / e;  a[i] = k;  a[j] += k * a_old[i-j];  e *= 1 - k^2
		double tmp[LPC_MAX_ORDER + 1];
		double e = r[0];
		a[0] = 1.0;
		bool ok = (e != 0.0) && (e == e);
		for (int i = 1; i <= order && ok; ++i) {
			double acc = r[i];
			for (int j = 1; j < i; ++j) acc = fma(a[j], r[i - j], acc);
			const double k = -acc / e;
			for (int j = 1; j < i; ++j) tmp[j] = a[j];
			for (int j = 1; j < i; ++j) a[j] = fma(k, tmp[i - j], a[j]);
			a[i] = k;
			e *= 1.0 - k * k;
		}
		float *dst = A.out + (size_t)g * A.ld + A.col_off;
		for (int i = 1; i <= order; ++i) {
			const double v = ok ? a[i] : 0.0;
			dst[i - 1] = (v == v) ? (float)v : 0.f;      // feature[isnan(feature)] = 0, LPC.py:56
		}
	}
}

int launch_lpc(const double *d_win, const double *d_win2, int L, int shift, const void *d_pcm, int is_f32, const long long *d_smp_off,
               const int *d_frame_off, int n_utt, int frames_total, int order, float *d_out, int ld, int col_off, cudaStream_t st) {
	SRB_REQUIRE(order >= 1 && order <= LPC_MAX_ORDER, "lpc: order %d outside [1,%d]", order, LPC_MAX_ORDER);
	SRB_REQUIRE(order < L, "lpc: order %d >= frame length %d", order, L);
	SRB_REQUIRE(ld >= col_off + order, "lpc: ld %d too small for %d coefficients at column %d", ld, order, col_off);
	if (frames_total <= 0) return 0;
	LpcArgs a;
	a.pcm = d_pcm;
	a.is_f32 = is_f32;
	a.smp_off = d_smp_off;
	a.frame_off = d_frame_off;
	a.n_utt = n_utt;
	a.frames_total = frames_total;
	a.L = L;
	a.shift = shift;
	a.order = order;
	a.win = d_win;
	a.win2 = d_win2;
	a.out = d_out;
	a.ld = ld;
	a.col_off = col_off;
	const size_t smem = (size_t)LPC_WARPS * (L + 2 * (LPC_MAX_ORDER + 1)) * sizeof(double);
	static bool attr = false;
	if (!attr) {
		SRB_CUDA(cudaFuncSetAttribute(lpc_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
		attr = true;
	}
	SRB_REQUIRE(smem <= 160 * 1024, "lpc: frame length %d too long", L);
	lpc_frames_kernel<<<ceil_div(frames_total, LPC_WARPS), LPC_WARPS * 32, smem, st>>>(a);
	SRB_CUDA(cudaGetLastError());
	return 0;
}

}  // namespace srb
