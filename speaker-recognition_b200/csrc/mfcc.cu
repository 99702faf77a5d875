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

struct MfccShape {                 // cached per batch shape (slots are REUSED on eviction: their buffers only grow)
	std::vector<long long> key;    // smp_off + nd
	int n_utt = 0, n_cta = 0, frames_total = 0, rows_total = 0, NF = 0;
	std::vector<int> feat_off;     // host copy [n_utt+1]
	// one device allocation + its pinned host image: smp_off[n_utt+1] (int64) | cta_start | frame_off | feat_off [n_utt+1] (int32) |
	// chunk records [n_cta] (int4: first sample lo, hi, first output frame row, frames)
	unsigned char *d_meta = nullptr, *h_meta = nullptr;
	size_t meta_cap = 0;
	cudaEvent_t copied = nullptr;  // the upload of h_meta has executed (h_meta may be rewritten)
	long long *d_smp_off = nullptr;
	int *d_cta_start = nullptr, *d_frame_off = nullptr, *d_feat_off = nullptr;
	int4 *d_cta_rec = nullptr;
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
	// fast path (M == 1024, frame length <= 512): one warp per frame, register-resident 4 x 256-point transforms
	bool fast = false;
	cpx *d_tw1 = nullptr, *d_twA = nullptr;
	int2 *d_mel_seg = nullptr;       // 16-bin segments of the bands' runs, starting on 4-bin boundaries: {first bin, band}
	float *d_mel_w16 = nullptr;      // their weights, [4 quarters][n_seg] float4 (zero outside the band's run)
	int *d_mel_segptr = nullptr;     // [nb + 1] first segment of every band
	int n_seg = 0;
	// work buffers shared by all shapes, grow-only: raw cepstra [frames][CL], CMVN partials [chunks][2 nc], stats [utt][2 nc]
	DevBuf ceps, partials, stats;
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
// Fast path: FFT_SIZE = 2048 (M = 1024) and frame length <= 512, i.e. at most 256 non-zero packed inputs -- the reference's
// default geometry at 16 kHz and 8 kHz.  ONE WARP OWNS ONE FRAME (and walks a chunk of consecutive frames of one utterance):
//   * persistent CTAs of 12 warps, one per SM; a warp's buffers are private, so the only synchronisation is __syncwarp -- round 1
//     ran 4 frames per CTA of 256 threads with six CTA barriers per frame group and 160 064 CTAs per step
//   * a frame's samples come in as 16-byte asynchronous copies (LDGSTS: global -> shared memory without registers), issued a whole
//     frame ahead into the second of two staging buffers (round 1: scalar 2-byte loads with the I2F stalled on them); the 50 %
//     overlap of consecutive frames is served by L2; every table of the kernel lives in shared memory (L1 is 3 KB here)
//   * window and pre-emphasis as the 2-tap form y[i] = w[i] x[i] - w2[i] x[i-1] with the lane's 32 coefficients in registers
//   * the zero-padded 1024-point packed transform = four 256-point transforms of pre-twiddled inputs, each two radix-16 steps in
//     registers (16 lanes x 16 points); the half-warps take r = h and r = h + 2, the packed frame is loaded once for both;
//     XOR-swizzled transpose tiles, second step in place (fft_core.h wf_*; replayed on the CPU by tests/test_fft_schedule_cpu.py)
//   * bins k and 1024 - k of the power spectrum come from ONE pair of packed results (half the unpack work)
//   * sparse mel: every band is one contiguous run of bins, cut into 16-bin segments that start on 4-bin boundaries: a lane does
//     one segment with four 128-bit loads of the (bank-skewed) spectrum and four of the weights; band sums in fixed order, log
//   * DCT rows 1..n_ceps with float64 accumulation, raw cepstra to HBM, float64 CMVN partial sums per chunk
// ------------------------------------------------------------------------------------------------
constexpr int MFW_WARPS = 12;
constexpr int MFW_THREADS = MFW_WARPS * 32;
constexpr int MFW_YLEN = 512;                     // longest frame of the fast path
constexpr int MFW_TW2 = 520;                      // unpack twiddles k = 0..512 (padded)
constexpr int MFW_STAGE_SAMPLES = 544;            // raw samples of one frame incl. the one before it, rounded out to 16-byte blocks (<= 67 x 8)
constexpr size_t MFW_STAGE_BYTES = 2 * MFW_STAGE_SAMPLES * sizeof(int16_t);      // two int16 buffers (double-buffered), or one float buffer
constexpr size_t MFW_CTAB_BYTES = (size_t)(4 * 16 + 64 * 16 + MFW_TW2) * sizeof(cpx);      // W_64^{n2 r} | W_1024^{u j} | unpack twiddles
constexpr size_t MFW_WARP_BYTES = MFW_STAGE_BYTES + (size_t)srbfft::WF_PLEN * sizeof(float) + (size_t)4 * srbfft::WF_TILE * sizeof(cpx);
// mel / DCT tables in shared memory: first bins int[n_segp] | weights float4[4][n_segp] | segptr int[nbp] | logfloor float[nbp] | dct double[nb][16]
static inline size_t mfw_mel_bytes(int n_seg, int nb) {
	const size_t n_segp = (size_t)(n_seg + 3) & ~(size_t)3, nbp = (size_t)(nb + 1 + 3) & ~(size_t)3;
	return n_segp * 4 + 4 * n_segp * 16 + nbp * 4 + nbp * 4 + (size_t)nb * 16 * 8;
}
static inline size_t mfw_smem_bytes(int n_seg, int nb) { return MFW_CTAB_BYTES + mfw_mel_bytes(n_seg, nb) + MFW_WARPS * MFW_WARP_BYTES; }

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <bool F32>
__global__ void __launch_bounds__(MFW_THREADS, 1) mfcc_frames_warp_kernel(const MfccArgs A, int n_chunks, long long n_total) {
	extern __shared__ __align__(16) unsigned char mf_smem[];
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int L = A.L, nb = A.nb, nc = A.nc, shift = A.shift, n_seg = A.n_seg;
	const int n_segp = (n_seg + 3) & ~3, nbp = (nb + 1 + 3) & ~3;
	cpx *s_tw1 = reinterpret_cast<cpx *>(mf_smem);          // [4][16]  W_64^{n2 r}: the half-warp-uniform part of the pre-twiddle
	cpx *s_twA = s_tw1 + 4 * 16;                            // [64][16] W_1024^{u j}, j = r + 4 q2: the rest of it merged with the inter-step twiddle
	cpx *s_tw2 = s_twA + 64 * 16;                           // [0..512] exp(-2 pi i k / 2048)
	int *s_segfirst = reinterpret_cast<int *>(s_tw2 + MFW_TW2);             // [n_segp] first bin of every 16-bin segment (multiple of 4)
	float4 *s_w16 = reinterpret_cast<float4 *>(s_segfirst + n_segp);        // [4][n_segp] segment weights, quarter-major
	int *s_segptr = reinterpret_cast<int *>(s_w16 + 4 * n_segp);            // [nb + 1]
	float *s_logfloor = reinterpret_cast<float *>(s_segptr + nbp);          // [nb]
	double *s_dct = reinterpret_cast<double *>(s_logfloor + nbp);           // [nb][16] transposed DCT rows 1..n_ceps
	unsigned char *wb = reinterpret_cast<unsigned char *>(s_dct + (size_t)nb * 16) + (size_t)warp * MFW_WARP_BYTES;
	unsigned char *stage = wb;                              // raw samples of the current / next frame, exactly as they lie in HBM
	float *Pb = reinterpret_cast<float *>(wb + MFW_STAGE_BYTES);            // [WF_PLEN] windowed frame (first 512 floats), then the skewed power spectrum
	cpx *TZ = reinterpret_cast<cpx *>(Pb + srbfft::WF_PLEN);                // [4][256] transpose tiles / sub-transform outputs, later the mel scratch

	for (int i = tid; i < 4 * 16; i += MFW_THREADS) s_tw1[i] = A.tw1[i];
	for (int i = tid; i < 64 * 16; i += MFW_THREADS) s_twA[i] = A.twA[i];
	for (int i = tid; i <= 512; i += MFW_THREADS) s_tw2[i] = A.tw2[i];
	for (int i = tid; i < n_seg; i += MFW_THREADS) s_segfirst[i] = A.mel_seg[i].x;
	for (int i = tid; i < 4 * n_seg; i += MFW_THREADS) s_w16[(i / n_seg) * n_segp + (i % n_seg)] = reinterpret_cast<const float4 *>(A.mel_w16)[i];
	for (int i = tid; i <= nb; i += MFW_THREADS) s_segptr[i] = A.mel_segptr[i];
	for (int i = tid; i < nb; i += MFW_THREADS) s_logfloor[i] = A.logfloor[i];
	for (int i = tid; i < nb * 16; i += MFW_THREADS) s_dct[i] = A.dct_t[i];
	for (int i = lane; i < srbfft::WF_PLEN; i += 32) Pb[i] = 0.f;          // the bins beyond 1024 a last mel segment may touch stay 0 (weights 0)
	for (int i = lane; i < (int)(MFW_STAGE_BYTES / 4); i += 32) reinterpret_cast<uint32_t *>(stage)[i] = 0u;
	// this lane's window coefficients for the frame positions i = lane + 32 j (zero beyond the frame length)
	float wr[16], w2r[16];
#pragma unroll
	for (int j = 0; j < 16; ++j) {
		const int i = lane + 32 * j;
		wr[j] = i < L ? __ldg(A.win + i) : 0.f;
		w2r[j] = i < L ? __ldg(A.win2 + i) : 0.f;
	}
	__syncthreads();

	const int h = lane >> 4, u = lane & 15;
	constexpr int G = F32 ? 4 : 8;                            // elements per 16 bytes
	constexpr int ES = F32 ? 4 : 2;                           // bytes per sample
	const long long ptr_elems = (long long)(reinterpret_cast<uintptr_t>(A.pcm) / ES);
	const unsigned char *pcm_b = reinterpret_cast<const unsigned char *>(A.pcm);

	// Stage the samples [b0, b0 + 8 nblk) of the frame that starts at sample a (b0 = the last 16-byte-aligned index <= a - 1) into
	// `dst`, asynchronously (LDGSTS: global -> shared without registers); pieces that stick out of [0, n_total) are filled by hand.
	auto stage_frame = [&](long long a, unsigned char *dst) -> int {
		long long m = (ptr_elems + a - 1) % G;
		if (m < 0) m += G;
		const long long b0 = a - 1 - m;
		const int npiece = (int)((a + L - b0 + G - 1) / G);           // 16-byte pieces
		for (int t = lane; t < npiece; t += 32) {
			const long long e0 = b0 + (long long)t * G;
			unsigned char *d = dst + (size_t)t * 16;
			if (e0 >= 0 && e0 + G <= n_total) cp_async16(d, pcm_b + e0 * ES);
			else
				for (int q = 0; q < G; ++q) {
					const long long g = e0 + q;
					if (F32) reinterpret_cast<float *>(d)[q] = (g >= 0 && g < n_total) ? reinterpret_cast<const float *>(A.pcm)[g] : 0.f;
					else reinterpret_cast<int16_t *>(d)[q] = (g >= 0 && g < n_total) ? reinterpret_cast<const int16_t *>(A.pcm)[g] : (int16_t)0;
				}
		}
		cp_async_commit();
		return (int)(a - b0);                                        // position of sample a in the buffer (1..G)
	};

	for (int c = blockIdx.x * MFW_WARPS + warp; c < n_chunks; c += gridDim.x * MFW_WARPS) {
		const int4 rec = __ldg(A.cta_rec + c);
		const long long a0 = ((long long)(unsigned)rec.x) | ((long long)rec.y << 32);
		const int row0 = rec.z, nfr = rec.w;
		double cs = 0.0, css = 0.0;                            // CMVN partial sums of coefficient `lane`
		int cur = 0;
		__syncwarp();                                          // the previous chunk's last frame has left the staging buffers
		int off = stage_frame(a0, stage);
		for (int f = 0; f < nfr; ++f) {
			cp_async_wait_all();
			__syncwarp();
			const unsigned char *sb = stage + (F32 ? 0 : (size_t)cur * MFW_STAGE_SAMPLES * 2);
			int off_next = 0;
			if (!F32 && f + 1 < nfr) {                         // the next frame's samples start their way now, into the other buffer
				off_next = stage_frame(a0 + (long long)(f + 1) * shift, stage + (size_t)(cur ^ 1) * MFW_STAGE_SAMPLES * 2);
			}
			// ---- window, then pre-emphasis (MFCC.py:61-64) as the 2-tap form y[i] = w[i] x[i] - w2[i] x[i-1]
#pragma unroll
			for (int j = 0; j < 16; ++j) {
				const int i = lane + 32 * j;
				float x1, x0;
				if (F32) {
					x1 = reinterpret_cast<const float *>(sb)[off + i];
					x0 = reinterpret_cast<const float *>(sb)[off + i - 1];
				} else {
					x1 = (float)reinterpret_cast<const int16_t *>(sb)[off + i];
					x0 = (float)reinterpret_cast<const int16_t *>(sb)[off + i - 1];
				}
				Pb[i] = fmaf(-w2r[j], x0, wr[j] * x1);
			}
			__syncwarp();
			if (F32 && f + 1 < nfr) off_next = stage_frame(a0 + (long long)(f + 1) * shift, stage);     // single float buffer: refill once it is consumed
			// ---- first radix-16 step of the two sub-transforms of this half-warp (r = h, h + 2)
			{
				cpx z[16];
				const cpx *yb = reinterpret_cast<const cpx *>(Pb);
#pragma unroll
				for (int n2 = 0; n2 < 16; ++n2) z[n2] = yb[u + 16 * n2];
				srbfft::wf_step_a(z, s_tw1 + h * 16, s_twA + h * 16, TZ + h * srbfft::WF_TILE, u);
				srbfft::wf_step_a(z, s_tw1 + (h + 2) * 16, s_twA + (h + 2) * 16, TZ + (h + 2) * srbfft::WF_TILE, u);
			}
			__syncwarp();
			// ---- second radix-16 step, in place tile by tile
#pragma unroll
			for (int round = 0; round < 2; ++round) {
				const int r = h + 2 * round;
				cpx in[16];
				srbfft::wf_step_b_load(TZ + r * srbfft::WF_TILE, u, in);
				__syncwarp();
				srbfft::wf_step_b_store(in, TZ, r, u);
			}
			__syncwarp();
			// ---- power spectrum, bins k and 1024 - k from one pair
			srbfft::wf_power_lane(TZ, Pb, s_tw2, lane);
			__syncwarp();
			// ---- sparse mel + log.  Floor semantics: the reference clamps each bin to 1e-100, which fp32 cannot hold; a band whose
			// fp32 energy is exactly 0 takes log(1e-100 * sum of its weights) instead.
			float *segsum = reinterpret_cast<float *>(TZ);         // [n_seg]   (the sub-transform outputs are consumed)
			float *logE = segsum + 512;                            // [nb]
			for (int sg = lane; sg < n_seg; sg += 32) {
				const int g0 = s_segfirst[sg] >> 2;                // first bin / 4
				float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
#pragma unroll
				for (int q = 0; q < 4; ++q) {
					const int g = g0 + q;
					const float4 pv = *reinterpret_cast<const float4 *>(Pb + 4 * (g + (g >> 2)));
					const float4 wv = s_w16[q * n_segp + sg];
					e0 = fmaf(wv.x, pv.x, e0);
					e1 = fmaf(wv.y, pv.y, e1);
					e2 = fmaf(wv.z, pv.z, e2);
					e3 = fmaf(wv.w, pv.w, e3);
				}
				segsum[sg] = (e0 + e1) + (e2 + e3);
			}
			__syncwarp();
			for (int b = lane; b < nb; b += 32) {
				const int s0 = s_segptr[b], s1 = s_segptr[b + 1];
				float e = 0.f;
				for (int sg = s0; sg < s1; ++sg) e += segsum[sg];
				logE[b] = e > 0.f ? logf(e) : s_logfloor[b];
			}
			__syncwarp();
			// ---- DCT rows 1..n_ceps, float64 accumulate: lane = (band parity, coefficient), four independent partial sums per lane
			// (a fixed order, so the result is deterministic); the table is transposed [band][16]
			{
				const int i = lane & 15;
				double acc[4] = {0.0, 0.0, 0.0, 0.0};
				for (int b2 = h; b2 < nb; b2 += 8) {
#pragma unroll
					for (int t = 0; t < 4; ++t) {
						const int bb = b2 + 2 * t;
						if (bb < nb) acc[t] = fma(s_dct[bb * 16 + i], (double)logE[bb], acc[t]);
					}
				}
				double tot = (acc[0] + acc[1]) + (acc[2] + acc[3]);
				tot += __shfl_xor_sync(0xffffffffu, tot, 16);
				if (lane < A.CL) {
					const float cf = lane < nc ? (float)tot : 0.f;
					A.ceps[(size_t)(row0 + f) * A.CL + lane] = cf;
					if (lane < nc) {
						cs += (double)cf;
						css = fma((double)cf, (double)cf, css);
					}
				}
			}
			off = off_next;
			cur ^= 1;
		}
		if (lane < nc) {
			A.partials[(size_t)c * 2 * nc + lane] = cs;
			A.partials[(size_t)c * 2 * nc + nc + lane] = css;
		}
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
// thread = (output row, cepstral coefficient i): one read of c[t], c[t-1], c[t-2] serves the row's static, delta and delta-delta
// columns (the same float64 arithmetic as one thread per output element, which this kernel replaced: 2.5x fewer threads, no
// per-element divisions); columns beyond (nd + 1) nc are zero-filled up to ld
__global__ void mfcc_norm_kernel(const float *__restrict__ ceps, int CL, const double *__restrict__ stats,
                                 const int *__restrict__ frame_off, const int *__restrict__ feat_off, int n_utt, int nc,
                                 int nd, int cmvn, float *__restrict__ feat, int ld, long long total_rows) {
	const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long row_ll = idx / CL;
	if (row_ll >= total_rows) return;
	const int row = (int)row_ll, i = (int)(idx - row_ll * CL);
	float *out = feat + (size_t)row * ld;
	for (int c = (nd + 1) * nc + i; c < ld; c += CL) out[c] = 0.f;
	if (i >= nc) return;
	const int u = find_segment(feat_off, n_utt, row);
	const int T_u = frame_off[u + 1] - frame_off[u];
	const int t = frame_off[u] + (row - feat_off[u]) + nd;   // raw frame index (global)
	double mean = 0.0, inv = 1.0;
	if (cmvn && T_u > 1) {
		mean = stats[(size_t)u * 2 * nc + i];
		inv = stats[(size_t)u * 2 * nc + nc + i];
	}
	const double c0 = (double)ceps[(size_t)t * CL + i];
	out[i] = (float)((c0 - mean) * inv);
	if (nd >= 1) {
		const double c1 = (double)ceps[(size_t)(t - 1) * CL + i];
		out[nc + i] = (float)((c0 - c1) * inv);
		if (nd >= 2) {
			const double c2 = (double)ceps[(size_t)(t - 2) * CL + i];
			out[2 * nc + i] = (float)((c0 - 2.0 * c1 + c2) * inv);
		}
	}
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
		std::vector<cpx> tw1(4 * 16), twA(64 * 16);      // W_64^{n2 r} and W_1024^{u j} (fft_core.h wf_step_a)
		for (int r = 0; r < 4; ++r)
			for (int n2 = 0; n2 < 16; ++n2) tw1[r * 16 + n2] = cpx{(float)cos(-2.0 * M_PI * n2 * r / 64), (float)sin(-2.0 * M_PI * n2 * r / 64)};
		for (int j = 0; j < 64; ++j)
			for (int u = 0; u < 16; ++u) twA[j * 16 + u] = cpx{(float)cos(-2.0 * M_PI * u * j / 1024), (float)sin(-2.0 * M_PI * u * j / 1024)};
		// 16-bin segments of the runs, each starting on a 4-bin boundary (see the mel stage of mfcc_frames_warp_kernel)
		std::vector<int2> seg;
		std::vector<float> w16;
		std::vector<int> segptr(nb + 1, 0);
		for (int b = 0; b < nb; ++b) {
			const int n = ptr[b + 1] - ptr[b], c0 = n > 0 ? col[ptr[b]] : 0;
			for (int first = c0 & ~3; first < c0 + n; first += 16) {
				seg.push_back(make_int2(first, b));
				for (int i = 0; i < 16; ++i) {
					const int bin = first + i;
					w16.push_back(bin >= c0 && bin < c0 + n ? val[ptr[b] + (bin - c0)] : 0.f);
				}
			}
			segptr[b + 1] = (int)seg.size();
		}
		{   // [seg][16] -> [4 quarters][seg] float4: the lanes of a warp read consecutive 16-byte words
			const size_t ns = seg.size();
			std::vector<float> t(w16.size());
			for (size_t sg = 0; sg < ns; ++sg)
				for (int i = 0; i < 16; ++i) t[((size_t)(i >> 2) * ns + sg) * 4 + (i & 3)] = w16[sg * 16 + i];
			w16.swap(t);
		}
		p->n_seg = (int)seg.size();
		if (p->n_seg > 512 || p->n_seg == 0 || mfw_smem_bytes(p->n_seg, nb) > 227 * 1024) p->fast = false;   // segment sums: 512 floats; tables: shared memory
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
	p->fast = (p->M == 1024 && p->L <= MFW_YLEN && n_ceps <= 16);
	p->NF = std::max(1, std::min(16, 4096 / p->M));      // generic kernel: frames per CTA (the fast path sizes its chunks per batch)
	// the log/cepstra scratch reuses one FFT buffer: NF*(nb+nc) floats must fit in NF*M*2 floats
	if ((p->nb + p->nc) > 2 * p->M) {
		set_error("mfcc: FFT_SIZE %d too small for %d filters", nfft, n_filters);
		delete p;
		return nullptr;
	}
	p->smem = (size_t)2 * p->NF * p->M * sizeof(cpx) + ((size_t)(p->NF - 1) * p->shift + p->L) * sizeof(float);
	if (p->smem > 200 * 1024) {
		set_error("mfcc: frame shift %d too large for the shared-memory staging", p->shift);
		delete p;
		return nullptr;
	}
	if (build_tables(p)) {
		mfcc_plan_free(p);
		return nullptr;
	}
	if (cudaFuncSetAttribute(mfcc_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
	    cudaFuncSetAttribute(mfcc_frames_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
	    cudaFuncSetAttribute(mfcc_frames_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
		set_error("mfcc: cudaFuncSetAttribute failed");
		mfcc_plan_free(p);
		return nullptr;
	}
	return p;
}

static void free_shape(MfccShape &s) {
	if (s.d_meta) cudaFree(s.d_meta);
	if (s.h_meta) cudaFreeHost(s.h_meta);
	if (s.copied) cudaEventDestroy(s.copied);
	s = MfccShape();
}

void mfcc_plan_free(MfccPlan *p) {
	if (!p) return;
	for (auto &s : p->shapes) free_shape(s);
	p->ceps.release();
	p->partials.release();
	p->stats.release();
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

static bool mfcc_force_generic() {      // A/B switch: run the generic shared-memory-FFT kernel on fast-path geometries too
	static const bool f = [] { const char *e = getenv("SRB200_MFCC_GENERIC"); return e && e[0] == '1'; }();
	return f;
}

// frames per chunk of the warp-per-frame kernel: chunks are dealt round-robin to the (SMs x 12) resident warps, so pick the
// chunk length in [32, 96] for which the number of chunks is closest below a whole number of rounds
static int pick_chunk_frames(const std::vector<long long> &T, int n_warps) {
	int best = 64;
	double best_eff = -1.0;
	for (int nf = 96; nf >= 32; --nf) {
		long long chunks = 0;
		for (long long t : T) chunks += (t + nf - 1) / nf;
		const long long rounds = (chunks + n_warps - 1) / n_warps;
		const double eff = (double)chunks / (double)(rounds * n_warps) - 0.002 * (96 - nf) / 64.0;      // slight preference for long chunks
		if (eff > best_eff) {
			best_eff = eff;
			best = nf;
		}
	}
	return best;
}

// Shape bookkeeping WITHOUT steady-state allocations: up to 16 shapes are cached; on a miss the least recently used slot is
// rewritten in place (pinned host image -> one async copy on the caller's stream); every buffer only ever grows.
static int get_shape(MfccPlan *p, const long long *smp_off, int n_utt, int nd, cudaStream_t st, MfccShape **out) {
	std::vector<long long> key(smp_off, smp_off + n_utt + 1);
	key.push_back(nd);
	for (auto &s : p->shapes)
		if (s.key == key) {
			s.stamp = ++p->stamp;
			*out = &s;
			return 0;
		}
	std::vector<long long> T(n_utt);
	for (int u = 0; u < n_utt; ++u) {
		const long long n = smp_off[u + 1] - smp_off[u];
		SRB_REQUIRE(n > 5LL * p->L, "mfcc: Signal too short! (utterance %d has %lld samples, need > 5*%d; MFCC.py:56)", u, n, p->L);
		T[u] = (n - p->L) / p->shift + 1;   // MFCC.py:57
		SRB_REQUIRE(T[u] - nd >= 1, "mfcc: utterance %d has too few frames for nd=%d", u, nd);
	}
	const int NF = (p->fast && !mfcc_force_generic()) ? pick_chunk_frames(T, ctx().sm_count * MFW_WARPS) : p->NF;
	std::vector<int> cta_start(n_utt + 1), frame_off(n_utt + 1), feat_off(n_utt + 1);
	cta_start[0] = frame_off[0] = feat_off[0] = 0;
	for (int u = 0; u < n_utt; ++u) {
		SRB_REQUIRE((long long)frame_off[u] + T[u] < (1LL << 31) - 64, "mfcc: batch too large (frame index overflows int32)");
		frame_off[u + 1] = frame_off[u] + (int)T[u];
		feat_off[u + 1] = feat_off[u] + (int)(T[u] - nd);
		cta_start[u + 1] = cta_start[u] + (int)((T[u] + NF - 1) / NF);
	}
	MfccShape *s = nullptr;
	if (p->shapes.size() < 16) {
		p->shapes.reserve(16);                 // pointers handed out earlier stay valid
		p->shapes.emplace_back();
		s = &p->shapes.back();
		SRB_CUDA(cudaEventCreateWithFlags(&s->copied, cudaEventDisableTiming));
	} else {
		s = &p->shapes[0];
		for (auto &c : p->shapes)
			if (c.stamp < s->stamp) s = &c;
		SRB_CUDA(cudaEventSynchronize(s->copied));      // its previous upload has executed: the pinned image may be rewritten
	}
	s->key = key;
	s->n_utt = n_utt;
	s->n_cta = cta_start[n_utt];
	s->frames_total = frame_off[n_utt];
	s->rows_total = feat_off[n_utt];
	s->NF = NF;
	s->feat_off = feat_off;
	s->stamp = ++p->stamp;
	const size_t nu = (size_t)n_utt + 1;
	const size_t off_cta = nu * sizeof(long long), off_frame = off_cta + nu * sizeof(int), off_feat = off_frame + nu * sizeof(int);
	const size_t off_rec = (off_feat + nu * sizeof(int) + 15) & ~(size_t)15, bytes = off_rec + (size_t)s->n_cta * sizeof(int4);
	if (bytes > s->meta_cap) {
		if (s->d_meta) {
			SRB_CUDA(cudaStreamSynchronize(st));
			cudaFree(s->d_meta);
			cudaFreeHost(s->h_meta);
			s->d_meta = s->h_meta = nullptr;
		}
		const size_t want = bytes + bytes / 2 + 4096;
		SRB_CUDA(cudaMalloc(&s->d_meta, want));
		SRB_CUDA(cudaMallocHost(&s->h_meta, want));
		s->meta_cap = want;
	}
	memcpy(s->h_meta, smp_off, nu * sizeof(long long));
	memcpy(s->h_meta + off_cta, cta_start.data(), nu * sizeof(int));
	memcpy(s->h_meta + off_frame, frame_off.data(), nu * sizeof(int));
	memcpy(s->h_meta + off_feat, feat_off.data(), nu * sizeof(int));
	int4 *rec = reinterpret_cast<int4 *>(s->h_meta + off_rec);
	for (int u = 0; u < n_utt; ++u)
		for (int c = cta_start[u]; c < cta_start[u + 1]; ++c) {
			const int fr0 = (c - cta_start[u]) * NF;
			const long long s0 = smp_off[u] + (long long)fr0 * p->shift;
			rec[c] = make_int4((int)(unsigned)(s0 & 0xFFFFFFFFll), (int)(s0 >> 32), frame_off[u] + fr0, (int)std::min<long long>(NF, T[u] - fr0));
		}
	s->d_smp_off = reinterpret_cast<long long *>(s->d_meta);
	s->d_cta_start = reinterpret_cast<int *>(s->d_meta + off_cta);
	s->d_frame_off = reinterpret_cast<int *>(s->d_meta + off_frame);
	s->d_feat_off = reinterpret_cast<int *>(s->d_meta + off_feat);
	s->d_cta_rec = reinterpret_cast<int4 *>(s->d_meta + off_rec);
	SRB_CUDA(cudaMemcpyAsync(s->d_meta, s->h_meta, bytes, cudaMemcpyHostToDevice, st));
	SRB_CUDA(cudaEventRecord(s->copied, st));
	*out = s;
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
	if (p->ceps.ensure((size_t)s->frames_total * p->CL * sizeof(float)) || p->partials.ensure((size_t)s->n_cta * 2 * p->nc * sizeof(double)) ||
	    p->stats.ensure((size_t)n_utt * 2 * p->nc * sizeof(double)))
		return 1;
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
	a.NF = s->NF;
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
	a.ceps = p->ceps.as<float>();
	a.partials = p->partials.as<double>();
	a.tw1 = p->d_tw1;
	a.twA = p->d_twA;
	a.cta_rec = s->d_cta_rec;
	a.mel_seg = p->d_mel_seg;
	a.mel_w16 = p->d_mel_w16;
	a.mel_segptr = p->d_mel_segptr;
	a.n_seg = p->n_seg;
	if (p->fast && !mfcc_force_generic()) {
		const int grid = std::min(ctx().sm_count, (s->n_cta + MFW_WARPS - 1) / MFW_WARPS);
		const long long n_total = smp_off[n_utt];
		if (is_f32) mfcc_frames_warp_kernel<true><<<grid, MFW_THREADS, mfw_smem_bytes(p->n_seg, p->nb), st>>>(a, s->n_cta, n_total);
		else mfcc_frames_warp_kernel<false><<<grid, MFW_THREADS, mfw_smem_bytes(p->n_seg, p->nb), st>>>(a, s->n_cta, n_total);
	} else mfcc_frames_kernel<<<s->n_cta, MF_THREADS, p->smem, st>>>(a);
	SRB_CUDA(cudaGetLastError());
	if (cmvn) {
		mfcc_stats_kernel<<<n_utt, 1024, 0, st>>>(p->partials.as<double>(), s->d_cta_start, s->d_frame_off, p->nc, p->stats.as<double>());
		SRB_CUDA(cudaGetLastError());
	}
	const long long total = (long long)s->rows_total * p->CL;
	const int blocks = (int)((total + 255) / 256);
	mfcc_norm_kernel<<<blocks, 256, 0, st>>>(p->ceps.as<float>(), p->CL, p->stats.as<double>(), s->d_frame_off, s->d_feat_off, n_utt, p->nc, nd, cmvn,
	                                        d_feat, ld, (long long)s->rows_total);
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
