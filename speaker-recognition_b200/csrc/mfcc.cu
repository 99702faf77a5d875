// mfcc.cu -- fused MFCC front-end for sm_100a.
//
// Replaces the reference's hot loop #1, a per-frame Python loop:
//   MFCCExtractor.extract     src/feature/MFCC.py:49-79   window -> pre-emphasis -> fft(2048) -> |.|^2 -> floor
//                                                          -> mel (dense 50x1025 matvec) -> log -> DCT -> CMVN
//   diff_feature              src/feature/utils.py:24-31   delta / delta-delta by first differences
// with three kernels over a batch of utterances:
//   mfcc_frames_kernel  one CTA = NF frames of one utterance: samples staged once in shared memory (frames overlap
//                       50 %), window*pre-emphasis folded into a 2-tap form, real FFT as a packed complex M = N/2
//                       point Stockham radix-4 transform in shared memory (fft_core.h), power spectrum, sparse mel
//                       (CSR, 1989 non-zeros instead of 51 250), log, DCT (float64 accumulate), raw cepstra to HBM
//                       plus per-CTA float64 partial sums for CMVN.
//   mfcc_stats_kernel   per utterance: deterministic reduction of the partials -> mean, 1/std (population, ddof = 0).
//   mfcc_norm_kernel    CMVN + delta/delta-delta epilogue -> feature rows [rows][ld] fp32.
// The tables (window, mel, DCT, twiddles) are built on the host in float64 exactly as the reference builds them
// (MFCC.py:14-16, 81-105, 107-113) and shipped as fp32 (DCT stays float64).
#include <math.h>
#include <algorithm>
#include "fft_core.h"
#include "kernels.h"

namespace srb {

using srbfft::cpx;

constexpr int MF_THREADS = 256;
constexpr double POWER_SPECTRUM_FLOOR = 1e-100;   // MFCC.py:7

struct MfccShape {                 // cached per batch shape
	std::vector<long long> key;    // smp_off + nd
	int n_utt = 0, n_cta = 0, frames_total = 0, rows_total = 0;
	std::vector<int> feat_off;     // host copy [n_utt+1]
	long long *d_smp_off = nullptr;
	int *d_cta_start = nullptr, *d_frame_off = nullptr, *d_feat_off = nullptr;
	float *d_ceps = nullptr;
	double *d_partials = nullptr, *d_stats = nullptr;
	int4 *d_cta_rec = nullptr;      // fast path, per CTA: {first sample lo, hi, first output frame row, frames (1..NF)}
	uint64_t stamp = 0;
};

struct MfccPlan {
	int fs = 0, L = 0, shift = 0, N = 0, M = 0, lgM = 0, nb = 0, nc = 0, CL = 0, NF = 0;
	double pre = 0;
	size_t smem = 0;
	float *d_win = nullptr, *d_win2 = nullptr, *d_mel_val = nullptr, *d_logfloor = nullptr;
	cpx *d_tw = nullptr, *d_tw2 = nullptr;
	int *d_mel_ptr = nullptr, *d_mel_col = nullptr;
	double *d_dct = nullptr, *d_dct_t = nullptr;
	double *d_win_d = nullptr, *d_win2_d = nullptr;   // float64 window / pre-emphasised window (LPC path)
	// fast path (M == 1024, frame length <= 512): register-resident 4 x 256-point transforms
	bool fast = false;
	cpx *d_tw1 = nullptr, *d_twA = nullptr;
	int2 *d_mel_seg = nullptr;       // 16-bin segments of the bands' runs: {first bin, band}; weights d_mel_w16[seg][16]
	float *d_mel_w16 = nullptr;
	int *d_mel_segptr = nullptr;     // [nb + 1] first segment of every band
	int n_seg = 0;
	size_t smem_fast = 0;
	std::vector<MfccShape> shapes;
	uint64_t stamp = 0;
};

struct MfccArgs {
	const void *pcm;
	int is_f32;
	const long long *smp_off;
	const int *cta_start, *frame_off;
	int n_utt;
	int L, shift, M, lgM, nb, nc, CL, NF;
	const float *win, *win2;
	const cpx *tw, *tw2;
	const int *mel_ptr, *mel_col;
	const float *mel_val, *logfloor;
	const double *dct;
	const double *dct_t;       // fast path: transposed [nb][16] (coefficients beyond nc are zero)
	float *ceps;
	double *partials;
	const cpx *tw1, *twA;      // fast path tables
	const int4 *cta_rec;       // fast path: per CTA {first sample (lo, hi), first frame row in ceps, frame count}
	const int2 *mel_seg;       // fast path: 16-bin segments {first bin, band}, weights mel_w16[seg][16] (zero outside the band's run)
	const float *mel_w16;
	const int *mel_segptr;
	int n_seg;
};

__device__ __forceinline__ int find_segment(const int *__restrict__ start, int n, int x) {
	// largest u in [0, n) with start[u] <= x   (start ascending, start[0] == 0)
	int lo = 0, hi = n - 1;
	while (lo < hi) {
		const int mid = (lo + hi + 1) >> 1;
		if (__ldg(start + mid) <= x) lo = mid;
		else hi = mid - 1;
	}
	return lo;
}

__global__ void __launch_bounds__(MF_THREADS) mfcc_frames_kernel(const MfccArgs A) {
	extern __shared__ __align__(16) unsigned char mf_smem[];
	const int tid = threadIdx.x;
	const int M = A.M, NF = A.NF, L = A.L, nb = A.nb, nc = A.nc;
	cpx *bufA = reinterpret_cast<cpx *>(mf_smem);
	cpx *bufB = bufA + (size_t)NF * M;
	float *sbuf = reinterpret_cast<float *>(bufB + (size_t)NF * M);

	const int u = find_segment(A.cta_start, A.n_utt, blockIdx.x);
	const int fbase = __ldg(A.frame_off + u);
	const int T_u = __ldg(A.frame_off + u + 1) - fbase;
	const int fr0 = (blockIdx.x - __ldg(A.cta_start + u)) * NF;
	const int nfr = min(NF, T_u - fr0);
	const long long s0 = A.smp_off[u] + (long long)fr0 * A.shift;
	const int span = (nfr - 1) * A.shift + L;

	// ---- stage the CTA's sample span once (frames overlap) ----
	if (A.is_f32) {
		const float *p = reinterpret_cast<const float *>(A.pcm) + s0;
		for (int i = tid; i < span; i += MF_THREADS) sbuf[i] = __ldg(p + i);
	} else {
		const int16_t *p = reinterpret_cast<const int16_t *>(A.pcm) + s0;
		for (int i = tid; i < span; i += MF_THREADS) sbuf[i] = (float)__ldg(p + i);
	}
	__syncthreads();

	// ---- window, then pre-emphasis (MFCC.py:61-64): y[i] = w[i] s[i] - pre * w[i-1] s[i-1]; pack z[n] = y[2n] + i y[2n+1]
	for (int t = tid; t < NF * M; t += MF_THREADS) {
		const int f = t >> A.lgM, n = t & (M - 1);
		float re = 0.f, im = 0.f;
		if (f < nfr) {
			const float *s = sbuf + f * A.shift;
			const int i0 = 2 * n, i1 = 2 * n + 1;
			if (i0 < L) {
				re = __ldg(A.win + i0) * s[i0];
				if (i0 > 0) re = fmaf(-__ldg(A.win2 + i0), s[i0 - 1], re);
			}
			if (i1 < L) im = fmaf(-__ldg(A.win2 + i1), s[i1 - 1], __ldg(A.win + i1) * s[i1]);
		}
		bufA[t] = cpx{re, im};
	}
	__syncthreads();

	// ---- packed complex M-point FFT, Stockham radix-4 (+ radix-2 tail), ping-pong between bufA/bufB ----
	cpx *in = bufA, *out = bufB;
	{
		int n = M, lgs = 0;
		while (n >= 4) {
			srbfft::radix4_pass(in, out, A.lgM, NF, n, lgs, A.tw, tid, MF_THREADS);
			__syncthreads();
			cpx *tmp = in;
			in = out;
			out = tmp;
			n >>= 2;
			lgs += 2;
		}
		if (n == 2) {
			srbfft::radix2_tail(in, out, A.lgM, NF, tid, MF_THREADS);
			__syncthreads();
			cpx *tmp = in;
			in = out;
			out = tmp;
		}
	}
	// ---- power spectrum P[f][0..M] into the free buffer ----
	float *P = reinterpret_cast<float *>(out);
	const int pstride = M + 1;
	srbfft::power_unpack(in, P, M, NF, pstride, A.tw2, tid, MF_THREADS);
	__syncthreads();

	// ---- sparse mel filterbank + log (MFCC.py:67-69).  Floor semantics: the reference clamps each bin to 1e-100,
	// which fp32 cannot hold; a band whose fp32 energy is exactly 0 takes log(1e-100 * sum of its weights) instead.
	float *logE = reinterpret_cast<float *>(in);       // [NF][nb]
	for (int t = tid; t < NF * nb; t += MF_THREADS) {
		const int f = t / nb, b = t - f * nb;
		const float *Pf = P + f * pstride;
		float e = 0.f;
		const int j1 = __ldg(A.mel_ptr + b + 1);
		for (int j = __ldg(A.mel_ptr + b); j < j1; ++j) e = fmaf(__ldg(A.mel_val + j), Pf[__ldg(A.mel_col + j)], e);
		logE[t] = e > 0.f ? logf(e) : __ldg(A.logfloor + b);
	}
	__syncthreads();

	// ---- DCT rows 1..n_ceps (MFCC.py:36-37,69), float64 accumulate ----
	float *cep = logE + NF * nb;                        // [NF][nc]
	for (int t = tid; t < NF * nc; t += MF_THREADS) {
		const int f = t / nc, i = t - f * nc;
		const double *dr = A.dct + (size_t)i * nb;
		const float *le = logE + f * nb;
		double acc = 0.0;
		for (int b = 0; b < nb; ++b) acc = fma(__ldg(dr + b), (double)le[b], acc);
		const float c = (float)acc;
		cep[t] = c;
		if (f < nfr) A.ceps[(size_t)(fbase + fr0 + f) * A.CL + i] = c;
	}
	for (int t = tid; t < NF * (A.CL - nc); t += MF_THREADS) {   // zero the pad columns
		const int f = t / (A.CL - nc), i = nc + t - f * (A.CL - nc);
		if (f < nfr) A.ceps[(size_t)(fbase + fr0 + f) * A.CL + i] = 0.f;
	}
	__syncthreads();
	if (tid < nc) {
		double s = 0.0, ss = 0.0;
		for (int f = 0; f < nfr; ++f) {
			const double c = (double)cep[f * nc + tid];
			s += c;
			ss = fma(c, c, ss);
		}
		A.partials[(size_t)blockIdx.x * 2 * nc + tid] = s;
		A.partials[(size_t)blockIdx.x * 2 * nc + nc + tid] = ss;
	}
}


// ------------------------------------------------------------------------------------------------
// Fast path: FFT_SIZE = 2048 (M = 1024) and frame length <= 512, i.e. at most 256 non-zero packed inputs -- the
// reference's default geometry at 16 kHz and 8 kHz.  The zero-padded transform is four 256-point transforms of
// pre-twiddled inputs, each done by 16 threads x 16 points in registers (fft_core.h: fast_step_a/b), so a frame costs
// two shared-memory round trips instead of five radix-4 passes, the pruned first pass disappears, and the mel sums are
// computed by 8-lane groups.  CTA = 256 threads = 4 frames x (4 sub-transforms x 16 lanes).
// ------------------------------------------------------------------------------------------------
constexpr int MFF_NF = 4;
constexpr int MFF_YLEN = 512;

__global__ void __launch_bounds__(MF_THREADS, 4) mfcc_frames_fast_kernel(const MfccArgs A) {
	extern __shared__ __align__(16) unsigned char mf_smem[];
	const int tid = threadIdx.x;
	const int L = A.L, nb = A.nb, nc = A.nc;
	// Two regions, 51 KB together (4 CTAs per SM at 64 registers).  Region A: sample span -> transpose tiles T -> sub-transform
	// outputs Zs, written IN PLACE over T after a barrier -> later the mel / log / cepstrum scratch.  Region B: windowed
	// frames ybuf -> later the power spectrum P.
	cpx *T = reinterpret_cast<cpx *>(mf_smem);                                   // [NF][4][16*17]
	cpx *Zs = T;                                                                 // [NF][4][260]  (aliases T)
	float *ybuf = reinterpret_cast<float *>(T + MFF_NF * 4 * 16 * srbfft::FAST_TSTRIDE);   // [NF][512], then P [NF][1028]

	// one 16-byte record per CTA (built on the host with the shape; a binary search over the utterance table cost six
	// dependent global loads before the first sample could be requested)
	const int4 rec = __ldg(A.cta_rec + blockIdx.x);
	const long long s0 = ((long long)(unsigned)rec.x) | ((long long)rec.y << 32);
	const int row0 = rec.z, nfr = rec.w;

	// window then pre-emphasis (MFCC.py:61-64) as a 2-tap form y[i] = w[i] x[i] - w2[i] x[i-1] (w2[0] = 0).  The CTA's frames
	// overlap, so its whole sample span (plus the sample before it) is staged once in shared memory as floats: one coalesced
	// pass, (NF-1) shift + L + 1 conversions instead of 2 NF L scattered 2-byte loads.
	{
		float *span = reinterpret_cast<float *>(T);                              // region A is free until the first FFT step
		const int n_span = (nfr - 1) * A.shift + L + 1;                          // span[j] = sample s0 - 1 + j
		for (int j = tid; j < n_span; j += MF_THREADS) {
			const long long g = s0 - 1 + j;
			float v = 0.f;
			if (g >= 0) v = A.is_f32 ? __ldg(reinterpret_cast<const float *>(A.pcm) + g) : (float)__ldg(reinterpret_cast<const int16_t *>(A.pcm) + g);
			span[j] = v;
		}
		__syncthreads();
		const float w0 = tid < L ? __ldg(A.win + tid) : 0.f, w1 = tid + 256 < L ? __ldg(A.win + tid + 256) : 0.f;
		const float v0 = tid < L ? __ldg(A.win2 + tid) : 0.f, v1 = tid + 256 < L ? __ldg(A.win2 + tid + 256) : 0.f;
#pragma unroll
		for (int it = 0; it < 8; ++it) {
			const int ff = it >> 1, i = tid + ((it & 1) << 8);
			float y = 0.f;
			if (ff < nfr && i < L) {
				const float *sp = span + ff * A.shift + i;
				y = fmaf(-((it & 1) ? v1 : v0), sp[0], ((it & 1) ? w1 : w0) * sp[1]);
			}
			ybuf[ff * MFF_YLEN + i] = y;
		}
	}
	__syncthreads();
	const int f = tid >> 6, r = (tid >> 4) & 3, u = tid & 15;
	cpx *Tfr = T + (f * 4 + r) * 16 * srbfft::FAST_TSTRIDE;
	srbfft::fast_step_a(reinterpret_cast<const cpx *>(ybuf + f * MFF_YLEN), A.tw1, A.twA, Tfr, r, u);
	__syncthreads();
	{
		cpx in[16];
		srbfft::fast_step_b_load(Tfr, u, in);
		__syncthreads();                                                         // every tile has been read: Zs may overwrite them
		srbfft::fast_step_b_store(in, Zs + (f * 4 + r) * srbfft::FAST_ZSTRIDE, u);
	}
	__syncthreads();
	float *P = ybuf;                                                             // [NF][1028], natural bin order (the frames are consumed)
	constexpr int PSTR = srbfft::FAST_PSTRIDE;
	srbfft::fast_power_unpack(Zs + f * 4 * srbfft::FAST_ZSTRIDE, P + f * PSTR, A.tw2, tid & 63);
	__syncthreads();

	// mel + log.  A band's non-zeros are one contiguous run of bins (triangular filters; checked on the host).  The runs are cut
	// into 16-bin segments (the last one end-aligned, weights zero where it overlaps its predecessor): one thread per
	// (frame, segment) does 16 multiply-adds on consecutive P values with its weights in four 128-bit loads, all items carry
	// the same work; then one thread per (frame, band) adds the band's segment sums in segment order (deterministic) and
	// takes the log.  (The first version walked (frame, band) items with 8-lane groups: 25 instructions per multiply-add.)
	float *logE = reinterpret_cast<float *>(T);                                  // [NF][nb]   (region A: Zs is consumed)
	float *segsum = logE + 512;                                                  // [NF][n_seg]  (n_seg <= 384, host-checked)
	{
		const int n_seg = A.n_seg, items = MFF_NF * n_seg;
		for (int item = tid; item < items; item += MF_THREADS) {
			const int ff = item / n_seg, sg = item - ff * n_seg;
			const int2 d = __ldg(A.mel_seg + sg);
			const float4 *w4 = reinterpret_cast<const float4 *>(A.mel_w16 + (size_t)sg * 16);
			const float *Pp = P + ff * PSTR + d.x;
			const float4 w0 = __ldg(w4), w1 = __ldg(w4 + 1), w2 = __ldg(w4 + 2), w3 = __ldg(w4 + 3);
			float e0 = w0.x * Pp[0], e1 = w0.y * Pp[1], e2 = w0.z * Pp[2], e3 = w0.w * Pp[3];
			e0 = fmaf(w1.x, Pp[4], e0);
			e1 = fmaf(w1.y, Pp[5], e1);
			e2 = fmaf(w1.z, Pp[6], e2);
			e3 = fmaf(w1.w, Pp[7], e3);
			e0 = fmaf(w2.x, Pp[8], e0);
			e1 = fmaf(w2.y, Pp[9], e1);
			e2 = fmaf(w2.z, Pp[10], e2);
			e3 = fmaf(w2.w, Pp[11], e3);
			e0 = fmaf(w3.x, Pp[12], e0);
			e1 = fmaf(w3.y, Pp[13], e1);
			e2 = fmaf(w3.z, Pp[14], e2);
			e3 = fmaf(w3.w, Pp[15], e3);
			segsum[item] = (e0 + e1) + (e2 + e3);
		}
		__syncthreads();
		for (int t = tid; t < MFF_NF * nb; t += MF_THREADS) {
			const int ff = t / nb, b = t - ff * nb;
			const int s0 = __ldg(A.mel_segptr + b), s1 = __ldg(A.mel_segptr + b + 1);
			float e = 0.f;
			for (int sg = s0; sg < s1; ++sg) e += segsum[ff * n_seg + sg];
			logE[t] = e > 0.f ? logf(e) : __ldg(A.logfloor + b);
		}
	}
	__syncthreads();
	// DCT rows 1..n_ceps, float64 accumulate.  One warp per frame: lane = (band parity, coefficient), the table is stored
	// transposed [band][16 coefficients] so that a request reads two contiguous 128-byte rows (the row-major table cost
	// eight 32-byte sectors per request: a fifth of the kernel's L1 wavefronts); the two parities are added by one shuffle
	// (fixed order -> deterministic).
	float *cep = logE + MFF_NF * nb;                                             // [NF][nc]
	if (tid < MFF_NF * 32) {
		const int ff = tid >> 5, h = (tid >> 4) & 1, i = tid & 15;
		const float *le = logE + ff * nb;
		double acc = 0.0;
		for (int b2 = h; b2 < nb; b2 += 2) acc = fma(__ldg(A.dct_t + b2 * 16 + i), (double)le[b2], acc);
		acc += __shfl_xor_sync(0xffffffffu, acc, 16);
		if (h == 0 && i < nc) {
			const float c = (float)acc;
			cep[ff * nc + i] = c;
			if (ff < nfr) A.ceps[(size_t)(row0 + ff) * A.CL + i] = c;
		}
	}
	for (int t = tid; t < MFF_NF * (A.CL - nc); t += MF_THREADS) {
		const int ff = t / (A.CL - nc), i = nc + t - ff * (A.CL - nc);
		if (ff < nfr) A.ceps[(size_t)(row0 + ff) * A.CL + i] = 0.f;
	}
	__syncthreads();
	if (tid < nc) {
		double s = 0.0, ss = 0.0;
		for (int ff = 0; ff < nfr; ++ff) {
			const double c = (double)cep[ff * nc + tid];
			s += c;
			ss = fma(c, c, ss);
		}
		A.partials[(size_t)blockIdx.x * 2 * nc + tid] = s;
		A.partials[(size_t)blockIdx.x * 2 * nc + nc + tid] = ss;
	}
}

// one block (32 warps) per utterance; lane = coefficient (nc <= 32), warps stride over the utterance's CTAs
__global__ void __launch_bounds__(1024) mfcc_stats_kernel(const double *__restrict__ partials, const int *__restrict__ cta_start,
                                                         const int *__restrict__ frame_off, int nc, double *__restrict__ stats) {
	__shared__ double sh[2][32][33];
	const int u = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int c0 = cta_start[u], c1 = cta_start[u + 1];
	double s = 0.0, ss = 0.0;
	if (lane < nc)
		for (int c = c0 + warp; c < c1; c += 32) {
			s += partials[(size_t)c * 2 * nc + lane];
			ss += partials[(size_t)c * 2 * nc + nc + lane];
		}
	sh[0][warp][lane] = s;
	sh[1][warp][lane] = ss;
	__syncthreads();
	if (warp == 0 && lane < nc) {
		double a = 0.0, b = 0.0;
		for (int w = 0; w < 32; ++w) {
			a += sh[0][w][lane];
			b += sh[1][w][lane];
		}
		const double T = (double)(frame_off[u + 1] - frame_off[u]);
		const double mean = a / T;
		double var = b / T - mean * mean;
		if (var < 0.0) var = 0.0;
		stats[(size_t)u * 2 * nc + lane] = mean;
		stats[(size_t)u * 2 * nc + nc + lane] = 1.0 / sqrt(var);   // std == 0 -> inf, as the reference's division does (MFCC.py:77)
	}
}

// CMVN (MFCC.py:74-77) + diff_feature (utils.py:24-31).  Output row r of utterance u uses raw frames r+nd, r+nd-1, r+nd-2.
__global__ void mfcc_norm_kernel(const float *__restrict__ ceps, int CL, const double *__restrict__ stats,
                                 const int *__restrict__ frame_off, const int *__restrict__ feat_off, int n_utt, int nc,
                                 int nd, int cmvn, float *__restrict__ feat, int ld, long long total) {
	const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= total) return;
	const int row = (int)(idx / ld), col = (int)(idx - (long long)row * ld);
	const int u = find_segment(feat_off, n_utt, row);
	const int T_u = frame_off[u + 1] - frame_off[u];
	const int t = frame_off[u] + (row - feat_off[u]) + nd;   // raw frame index (global)
	const int k = col / nc, i = col - k * nc;
	float v = 0.f;
	if (k <= nd) {
		double mean = 0.0, inv = 1.0;
		if (cmvn && T_u > 1) {
			mean = stats[(size_t)u * 2 * nc + i];
			inv = stats[(size_t)u * 2 * nc + nc + i];
		}
		const double c0 = (double)ceps[(size_t)t * CL + i];
		if (k == 0) v = (float)((c0 - mean) * inv);
		else {
			const double c1 = (double)ceps[(size_t)(t - 1) * CL + i];
			if (k == 1) v = (float)((c0 - c1) * inv);
			else {
				const double c2 = (double)ceps[(size_t)(t - 2) * CL + i];
				v = (float)((c0 - 2.0 * c1 + c2) * inv);
			}
		}
	}
	feat[idx] = v;
}

// ------------------------------------------------------------------------------------------------
// host: tables (float64, as the reference builds them), plan, shape cache, launch
// ------------------------------------------------------------------------------------------------
template <class T> static int to_device(T **dst, const std::vector<T> &src) {
	SRB_CUDA(cudaMalloc((void **)dst, std::max<size_t>(1, src.size()) * sizeof(T)));
	if (!src.empty()) SRB_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
	return 0;
}

static int build_tables(MfccPlan *p) {
	const int L = p->L, N = p->N, M = p->M, nb = p->nb, nc = p->nc;
	// hamming, MFCC.py:14-16 (half-sample phase, period L); win2[i] = pre * w[i-1]
	std::vector<float> win(L), win2(L);
	std::vector<double> w(L);
	for (int i = 0; i < L; ++i) w[i] = 0.54 - 0.46 * cos(2 * M_PI / L * (i + 0.5));
	std::vector<double> w2(L);
	for (int i = 0; i < L; ++i) {
		win[i] = (float)w[i];
		win2[i] = i > 0 ? (float)(p->pre * w[i - 1]) : 0.f;
		w2[i] = i > 0 ? p->pre * w[i - 1] : 0.0;
	}
	if (to_device(&p->d_win_d, w) || to_device(&p->d_win2_d, w2)) return 1;
	// mel filterbank, MFCC.py:81-105
	const double f0 = 700.0 / p->fs;
	const int fn2 = N / 2;
	const double lr = log(1 + 0.5 / f0) / (nb + 1);
	const double cs[4] = {0.0, 1.0, (double)nb, (double)nb + 1};
	double bl[4];
	for (int i = 0; i < 4; ++i) bl[i] = N * f0 * (exp(cs[i] * lr) - 1);
	const int b1 = (int)floor(bl[0]) + 1, b2 = (int)ceil(bl[1]), b3 = (int)floor(bl[2]);
	const int b4 = std::min(fn2, (int)ceil(bl[3])) - 1;
	const int npf = b4 - b1 + 1;
	if (npf <= 0) {
		set_error("mfcc: degenerate mel filterbank (fs=%d FFT_SIZE=%d n_filters=%d)", p->fs, N, nb);
		return 1;
	}
	std::vector<double> fp(npf), pm(npf);
	for (int c = 0; c < npf; ++c) {
		const double pf = log(1 + (double)(b1 + c) / f0 / N) / lr;
		fp[c] = floor(pf);
		pm[c] = pf - fp[c];
	}
	std::vector<double> dense((size_t)nb * (fn2 + 1), 0.0);
	auto add = [&](int r, int col, double v) -> int {
		if (r < 0) r += nb;   // numpy negative-index wrap
		if (r < 0 || r >= nb || col < 0 || col > fn2) {
			set_error("mfcc: mel filterbank index out of range (the reference raises IndexError here)");
			return 1;
		}
		dense[(size_t)r * (fn2 + 1) + col] += v;
		return 0;
	};
	for (int c = b2 - 1; c < b4; ++c) {
		if (c < 0 || c >= npf) continue;
		if (add((int)fp[c] - 1, c + 1, 2 * (1 - pm[c]))) return 1;
	}
	for (int c = 0; c < b3; ++c) {
		if (c >= npf) break;
		if (add((int)fp[c], c + 1, 2 * pm[c])) return 1;
	}
	std::vector<int> ptr(nb + 1, 0), col;
	std::vector<float> val, logfloor(nb);
	for (int b = 0; b < nb; ++b) {
		double rowsum = 0;
		for (int j = 0; j <= fn2; ++j) {
			const double v = dense[(size_t)b * (fn2 + 1) + j];
			if (v != 0.0) {
				col.push_back(j);
				val.push_back((float)v);
				rowsum += v;
			}
		}
		ptr[b + 1] = (int)col.size();
		logfloor[b] = (float)log(rowsum * POWER_SPECTRUM_FLOOR);   // all bins at the floor (digital silence); -inf for an empty band
	}
	// DCT-II rows 1..nc, MFCC.py:107-113,36-37
	std::vector<double> dct((size_t)nc * nb);
	for (int i = 0; i < nc; ++i)
		for (int b = 0; b < nb; ++b) dct[(size_t)i * nb + b] = sqrt(2.0 / nb) * cos(M_PI * (2 * b + 1) * (i + 1) / (2.0 * nb));
	std::vector<double> dct_t((size_t)nb * 16, 0.0);
	if (nc <= 16)
		for (int i = 0; i < nc; ++i)
			for (int b = 0; b < nb; ++b) dct_t[(size_t)b * 16 + i] = dct[(size_t)i * nb + b];
	else p->fast = false;
	if (to_device(&p->d_dct_t, dct_t)) return 1;
	// twiddles
	std::vector<cpx> tw(M), tw2(M + 1);
	for (int j = 0; j < M; ++j) tw[j] = cpx{(float)cos(-2.0 * M_PI * j / M), (float)sin(-2.0 * M_PI * j / M)};
	for (int k = 0; k <= M; ++k) tw2[k] = cpx{(float)cos(-2.0 * M_PI * k / N), (float)sin(-2.0 * M_PI * k / N)};
	// the fast path reads each band as one contiguous run of bins; true for triangular filters, verified here
	for (int b = 0; b < nb && p->fast; ++b)
		for (int j = ptr[b] + 1; j < ptr[b + 1]; ++j)
			if (col[j] != col[j - 1] + 1) p->fast = false;
	if (p->fast) {
		std::vector<cpx> tw1(3 * 256), twA(256);
		for (int r = 1; r < 4; ++r)
			for (int n = 0; n < 256; ++n)
				tw1[(r - 1) * 256 + n] = cpx{(float)cos(-2.0 * M_PI * n * r / 1024), (float)sin(-2.0 * M_PI * n * r / 1024)};
		for (int q2 = 0; q2 < 16; ++q2)
			for (int u = 0; u < 16; ++u) twA[q2 * 16 + u] = cpx{(float)cos(-2.0 * M_PI * u * q2 / 256), (float)sin(-2.0 * M_PI * u * q2 / 256)};
		// 16-bin segments of the runs (see the mel stage of mfcc_frames_fast_kernel)
		std::vector<int2> seg;
		std::vector<float> w16;
		std::vector<int> segptr(nb + 1, 0);
		for (int b = 0; b < nb; ++b) {
			const int n = ptr[b + 1] - ptr[b], c0 = n > 0 ? col[ptr[b]] : 0;
			int covered = c0;                                    // bins of the run below this index are already in a segment
			for (int st = c0; covered < c0 + n;) {
				int first = st;
				if (first + 16 > c0 + n) first = std::max(0, c0 + n - 16);   // end-align the last segment
				seg.push_back(make_int2(first, b));
				for (int i = 0; i < 16; ++i) {
					const int bin = first + i;
					const bool in = bin >= covered && bin < c0 + n;
					w16.push_back(in ? val[ptr[b] + (bin - c0)] : 0.f);
				}
				covered = first + 16;
				st = covered;
			}
			segptr[b + 1] = (int)seg.size();
		}
		p->n_seg = (int)seg.size();
		if (p->n_seg > 384 || p->n_seg == 0) p->fast = false;   // the segment sums share a 1.5 K-float shared-memory region
		if (to_device(&p->d_tw1, tw1) || to_device(&p->d_twA, twA) || to_device(&p->d_mel_seg, seg) ||
		    to_device(&p->d_mel_w16, w16) || to_device(&p->d_mel_segptr, segptr))
			return 1;
	}
	if (to_device(&p->d_win, win) || to_device(&p->d_win2, win2) || to_device(&p->d_mel_ptr, ptr) ||
	    to_device(&p->d_mel_col, col) || to_device(&p->d_mel_val, val) || to_device(&p->d_logfloor, logfloor) ||
	    to_device(&p->d_dct, dct) || to_device(&p->d_tw, tw) || to_device(&p->d_tw2, tw2))
		return 1;
	return 0;
}

MfccPlan *mfcc_plan_create(int fs, double win_ms, double shift_ms, int nfft, int n_filters, int n_ceps, double pre_emph) {
	if (ctx().init()) return nullptr;
	MfccPlan *p = new MfccPlan();
	p->fs = fs;
	p->L = (int)(win_ms / 1000 * fs);       // MFCC.py:28 int(float(ms) / 1000 * fs)
	p->shift = (int)(shift_ms / 1000 * fs); // MFCC.py:29
	p->N = nfft;
	p->M = nfft / 2;
	p->nb = n_filters;
	p->nc = n_ceps;
	p->pre = pre_emph;
	p->CL = (n_ceps + 3) & ~3;
	int lg = 0;
	while ((1 << lg) < p->M) ++lg;
	p->lgM = lg;
	bool ok = fs > 0 && p->L >= 2 && p->shift >= 1 && nfft >= 64 && nfft <= 4096 && (nfft & (nfft - 1)) == 0 && p->L <= nfft &&
	          n_filters >= 2 && n_filters <= 128 && n_ceps >= 1 && n_ceps <= 32 && n_ceps < n_filters;
	if (!ok) {
		set_error("mfcc: unsupported configuration fs=%d win=%gms shift=%gms FFT_SIZE=%d n_filters=%d n_ceps=%d "
		          "(need power-of-two FFT_SIZE in [64,4096] >= frame length %d, n_filters<=128, n_ceps<=32)",
		          fs, win_ms, shift_ms, nfft, n_filters, n_ceps, p->L);
		delete p;
		return nullptr;
	}
	p->fast = (p->M == 1024 && p->L <= MFF_YLEN && n_filters + n_ceps <= MFF_YLEN);
	p->NF = p->fast ? MFF_NF : std::max(1, std::min(16, 4096 / p->M));
	// the log/cepstra scratch reuses one FFT buffer: NF*(nb+nc) floats must fit in NF*M*2 floats
	if ((p->nb + p->nc) > 2 * p->M) {
		set_error("mfcc: FFT_SIZE %d too small for %d filters", nfft, n_filters);
		delete p;
		return nullptr;
	}
	p->smem = (size_t)2 * p->NF * p->M * sizeof(cpx) + ((size_t)(p->NF - 1) * p->shift + p->L) * sizeof(float);
	p->smem_fast = (size_t)MFF_NF * 4 * 16 * srbfft::FAST_TSTRIDE * sizeof(cpx) + (size_t)MFF_NF * srbfft::FAST_PSTRIDE * sizeof(float);
	if (p->smem > 200 * 1024 || (p->fast && p->smem_fast > 110 * 1024)) {
		set_error("mfcc: frame shift %d too large for the shared-memory staging", p->shift);
		delete p;
		return nullptr;
	}
	if (build_tables(p)) {
		mfcc_plan_free(p);
		return nullptr;
	}
	if (cudaFuncSetAttribute(mfcc_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
	    cudaFuncSetAttribute(mfcc_frames_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
		set_error("mfcc: cudaFuncSetAttribute failed");
		mfcc_plan_free(p);
		return nullptr;
	}
	return p;
}

static void free_shape(MfccShape &s) {
	cudaFree(s.d_smp_off);
	cudaFree(s.d_cta_start);
	cudaFree(s.d_frame_off);
	cudaFree(s.d_feat_off);
	cudaFree(s.d_ceps);
	cudaFree(s.d_partials);
	cudaFree(s.d_stats);
	cudaFree(s.d_cta_rec);
	s = MfccShape();
}

void mfcc_plan_free(MfccPlan *p) {
	if (!p) return;
	for (auto &s : p->shapes) free_shape(s);
	cudaFree(p->d_win);
	cudaFree(p->d_win2);
	cudaFree(p->d_mel_val);
	cudaFree(p->d_logfloor);
	cudaFree(p->d_tw);
	cudaFree(p->d_tw2);
	cudaFree(p->d_mel_ptr);
	cudaFree(p->d_mel_col);
	cudaFree(p->d_dct);
	cudaFree(p->d_dct_t);
	cudaFree(p->d_win_d);
	cudaFree(p->d_win2_d);
	cudaFree(p->d_tw1);
	cudaFree(p->d_twA);
	cudaFree(p->d_mel_seg);
	cudaFree(p->d_mel_w16);
	cudaFree(p->d_mel_segptr);
	delete p;
}

int mfcc_frame_len(const MfccPlan *p) { return p->L; }
int mfcc_frame_shift(const MfccPlan *p) { return p->shift; }
int mfcc_n_ceps(const MfccPlan *p) { return p->nc; }

static int get_shape(MfccPlan *p, const long long *smp_off, int n_utt, int nd, cudaStream_t st, MfccShape **out) {
	std::vector<long long> key(smp_off, smp_off + n_utt + 1);
	key.push_back(nd);
	for (auto &s : p->shapes)
		if (s.key == key) {
			s.stamp = ++p->stamp;
			*out = &s;
			return 0;
		}
	std::vector<int> cta_start(n_utt + 1), frame_off(n_utt + 1), feat_off(n_utt + 1);
	cta_start[0] = frame_off[0] = feat_off[0] = 0;
	for (int u = 0; u < n_utt; ++u) {
		const long long n = smp_off[u + 1] - smp_off[u];
		SRB_REQUIRE(n > 5LL * p->L, "mfcc: Signal too short! (utterance %d has %lld samples, need > 5*%d; MFCC.py:56)", u, n, p->L);
		const long long T = (n - p->L) / p->shift + 1;   // MFCC.py:57
		SRB_REQUIRE(T - nd >= 1, "mfcc: utterance %d has too few frames for nd=%d", u, nd);
		SRB_REQUIRE((long long)frame_off[u] + T < (1LL << 31) - 64, "mfcc: batch too large (frame index overflows int32)");
		frame_off[u + 1] = frame_off[u] + (int)T;
		feat_off[u + 1] = feat_off[u] + (int)(T - nd);
		cta_start[u + 1] = cta_start[u] + (int)((T + p->NF - 1) / p->NF);
	}
	if (p->shapes.size() >= 16) {
		size_t lru = 0;
		for (size_t i = 1; i < p->shapes.size(); ++i)
			if (p->shapes[i].stamp < p->shapes[lru].stamp) lru = i;
		SRB_CUDA(cudaStreamSynchronize(st));
		free_shape(p->shapes[lru]);
		p->shapes.erase(p->shapes.begin() + lru);
	}
	MfccShape s;
	s.key = key;
	s.n_utt = n_utt;
	s.n_cta = cta_start[n_utt];
	s.frames_total = frame_off[n_utt];
	s.rows_total = feat_off[n_utt];
	s.feat_off = feat_off;
	s.stamp = ++p->stamp;
	const size_t nu = (size_t)n_utt + 1;
	SRB_CUDA(cudaMalloc(&s.d_smp_off, nu * sizeof(long long)));
	SRB_CUDA(cudaMalloc(&s.d_cta_start, nu * sizeof(int)));
	SRB_CUDA(cudaMalloc(&s.d_frame_off, nu * sizeof(int)));
	SRB_CUDA(cudaMalloc(&s.d_feat_off, nu * sizeof(int)));
	SRB_CUDA(cudaMalloc(&s.d_ceps, (size_t)s.frames_total * p->CL * sizeof(float)));
	SRB_CUDA(cudaMalloc(&s.d_partials, (size_t)s.n_cta * 2 * p->nc * sizeof(double)));
	SRB_CUDA(cudaMalloc(&s.d_stats, (size_t)n_utt * 2 * p->nc * sizeof(double)));
	{
		std::vector<int4> rec((size_t)s.n_cta);
		for (int u = 0; u < n_utt; ++u) {
			const int T = frame_off[u + 1] - frame_off[u];
			for (int c = cta_start[u]; c < cta_start[u + 1]; ++c) {
				const int fr0 = (c - cta_start[u]) * p->NF;
				const long long s0 = smp_off[u] + (long long)fr0 * p->shift;
				rec[c] = make_int4((int)(unsigned)(s0 & 0xFFFFFFFFll), (int)(s0 >> 32), frame_off[u] + fr0, std::min(p->NF, T - fr0));
			}
		}
		SRB_CUDA(cudaMalloc(&s.d_cta_rec, rec.size() * sizeof(int4)));
		SRB_CUDA(cudaMemcpy(s.d_cta_rec, rec.data(), rec.size() * sizeof(int4), cudaMemcpyHostToDevice));
	}
	SRB_CUDA(cudaMemcpyAsync(s.d_smp_off, smp_off, nu * sizeof(long long), cudaMemcpyHostToDevice, st));
	SRB_CUDA(cudaMemcpyAsync(s.d_cta_start, cta_start.data(), nu * sizeof(int), cudaMemcpyHostToDevice, st));
	SRB_CUDA(cudaMemcpyAsync(s.d_frame_off, frame_off.data(), nu * sizeof(int), cudaMemcpyHostToDevice, st));
	SRB_CUDA(cudaMemcpyAsync(s.d_feat_off, feat_off.data(), nu * sizeof(int), cudaMemcpyHostToDevice, st));
	SRB_CUDA(cudaStreamSynchronize(st));
	p->shapes.push_back(s);
	*out = &p->shapes.back();
	return 0;
}

int launch_mfcc(MfccPlan *p, const void *d_pcm, int is_f32, const long long *smp_off, int n_utt, int nd, int cmvn, float *d_feat,
                int ld, int *feat_off, cudaStream_t st) {
	SRB_REQUIRE(p != nullptr, "mfcc: null plan");
	SRB_REQUIRE(n_utt >= 1, "mfcc: no utterances");
	SRB_REQUIRE(nd >= 0 && nd <= 2, "mfcc: nd must be 0, 1 or 2");
	SRB_REQUIRE(cmvn || nd == 0, "mfcc: raw (cmvn=0) output supports nd=0 only");
	const int out_dim = p->nc * (nd + 1);
	SRB_REQUIRE(ld >= out_dim, "mfcc: ld %d < feature dim %d", ld, out_dim);
	MfccShape *s = nullptr;
	if (get_shape(p, smp_off, n_utt, nd, st, &s)) return 1;
	MfccArgs a;
	a.pcm = d_pcm;
	a.is_f32 = is_f32;
	a.smp_off = s->d_smp_off;
	a.cta_start = s->d_cta_start;
	a.frame_off = s->d_frame_off;
	a.n_utt = n_utt;
	a.L = p->L;
	a.shift = p->shift;
	a.M = p->M;
	a.lgM = p->lgM;
	a.nb = p->nb;
	a.nc = p->nc;
	a.CL = p->CL;
	a.NF = p->NF;
	a.win = p->d_win;
	a.win2 = p->d_win2;
	a.tw = p->d_tw;
	a.tw2 = p->d_tw2;
	a.mel_ptr = p->d_mel_ptr;
	a.mel_col = p->d_mel_col;
	a.mel_val = p->d_mel_val;
	a.logfloor = p->d_logfloor;
	a.dct = p->d_dct;
	a.dct_t = p->d_dct_t;
	a.ceps = s->d_ceps;
	a.partials = s->d_partials;
	a.tw1 = p->d_tw1;
	a.twA = p->d_twA;
	a.cta_rec = s->d_cta_rec;
	a.mel_seg = p->d_mel_seg;
	a.mel_w16 = p->d_mel_w16;
	a.mel_segptr = p->d_mel_segptr;
	a.n_seg = p->n_seg;
	static const bool force_generic = [] { const char *e = getenv("SRB200_MFCC_GENERIC"); return e && e[0] == '1'; }();
	if (p->fast && !force_generic) mfcc_frames_fast_kernel<<<s->n_cta, MF_THREADS, p->smem_fast, st>>>(a);
	else mfcc_frames_kernel<<<s->n_cta, MF_THREADS, p->smem, st>>>(a);
	SRB_CUDA(cudaGetLastError());
	if (cmvn) {
		mfcc_stats_kernel<<<n_utt, 1024, 0, st>>>(s->d_partials, s->d_cta_start, s->d_frame_off, p->nc, s->d_stats);
		SRB_CUDA(cudaGetLastError());
	}
	const long long total = (long long)s->rows_total * ld;
	const int blocks = (int)((total + 255) / 256);
	mfcc_norm_kernel<<<blocks, 256, 0, st>>>(s->d_ceps, p->CL, s->d_stats, s->d_frame_off, s->d_feat_off, n_utt, p->nc, nd, cmvn,
	                                        d_feat, ld, total);
	SRB_CUDA(cudaGetLastError());
	if (feat_off) memcpy(feat_off, s->feat_off.data(), sizeof(int) * (n_utt + 1));
	return 0;
}

int launch_lpc_for_plan(MfccPlan *p, const void *d_pcm, int is_f32, const long long *smp_off, int n_utt, int order, float *d_out,
                        int ld, int col_off, int *feat_off, cudaStream_t st) {
	SRB_REQUIRE(p != nullptr, "lpc: null plan");
	SRB_REQUIRE(n_utt >= 1, "lpc: no utterances");
	MfccShape *s = nullptr;
	if (get_shape(p, smp_off, n_utt, 0, st, &s)) return 1;      // nd = 0: one output row per frame
	if (launch_lpc(p->d_win_d, p->d_win2_d, p->L, p->shift, d_pcm, is_f32, s->d_smp_off, s->d_frame_off, n_utt, s->frames_total, order,
	               d_out, ld, col_off, st))
		return 1;
	if (feat_off) memcpy(feat_off, s->feat_off.data(), sizeof(int) * (n_utt + 1));
	return 0;
}

}  // namespace srb
