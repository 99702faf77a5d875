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
// complex add / subtract as ONE packed fp32x2 instruction on the device (sm_100 FADD2: both halves round to nearest exactly like two
// scalar adds, so host replay and device agree bit for bit); a complex value lives in an aligned register pair, the moves are renames
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ cpx cadd2(cpx a, cpx b) {
	unsigned long long ra, rb, rc;
	cpx c;
	asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
	asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
	asm("mov.b64 {%0, %1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(rc));
	return c;
}
__device__ __forceinline__ cpx csub2(cpx a, cpx b) {
	unsigned long long ra, rb, rc;
	cpx c;
	asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
	asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
	asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
	asm("mov.b64 {%0, %1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(rc));
	return c;
}
#else
SRB_HD cpx cadd2(cpx a, cpx b) { return cadd(a, b); }
SRB_HD cpx csub2(cpx a, cpx b) { return csub(a, b); }
#endif
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

// ------------------------------------------------------------------------------------------------
// Register-resident building blocks of the fast MFCC path (M = 1024, at most 256 non-zero packed inputs):
// the zero-padded 1024-point transform splits exactly into four 256-point transforms of pre-twiddled inputs,
//     Z[4q + r] = sum_{n<256} (z[n] W_1024^{n r}) W_256^{n q},        r = 0..3, q = 0..255,
// tables: wf_c[r][n2] = W_64^{n2 r} (4 x 16) and wf_t[j][u] = W_1024^{u j} (64 x 16, j = r + 4 q2), see wf_step_a,
// and each 256-point transform is two radix-16 steps held in registers (16 threads x 16 points) with one
// transpose through shared memory in between.
// ------------------------------------------------------------------------------------------------

// forward 4-point DFT, outputs in natural order
SRB_HD void dft4(cpx &a, cpx &b, cpx &c, cpx &d) {
	const cpx apc = cadd2(a, c), amc = csub2(a, c), bpd = cadd2(b, d);
	const cpx jbmd = cmulj_neg(csub2(b, d));     // -i (b - d): the swap of the halves keeps these two scalar
	a = cadd2(apc, bpd);
	b = cadd(amc, jbmd);
	c = csub2(apc, bpd);
	d = csub(amc, jbmd);
}

// forward 16-point DFT of v[0..15] into o[0..15] (natural order).  n = 4 n1 + n0, k = 4 k1 + k0:
//   X[4k1 + k0] = sum_{n0} W4^{n0 k1} [ W16^{n0 k0} sum_{n1} x[4 n1 + n0] W4^{n1 k0} ]
SRB_HD void dft16(const cpx *v, cpx *o) {
	const float C1 = 0.92387953251128675613f, S1 = 0.38268343236508977173f, R2 = 0.70710678118654752440f;
	cpx t[4][4];   // t[n0][k0]
#pragma unroll
	for (int n0 = 0; n0 < 4; ++n0) {
		cpx a = v[n0], b = v[n0 + 4], c = v[n0 + 8], d = v[n0 + 12];
		dft4(a, b, c, d);
		t[n0][0] = a;
		t[n0][1] = b;
		t[n0][2] = c;
		t[n0][3] = d;
	}
	// twiddles W16^{n0 k0}: exponents 1,2,3 | 2,4,6 | 3,6,9
	t[1][1] = cmul(t[1][1], cpx{C1, -S1});
	t[1][2] = cmul(t[1][2], cpx{R2, -R2});
	t[1][3] = cmul(t[1][3], cpx{S1, -C1});
	t[2][1] = cmul(t[2][1], cpx{R2, -R2});
	t[2][2] = cmulj_neg(t[2][2]);
	t[2][3] = cmul(t[2][3], cpx{-R2, -R2});
	t[3][1] = cmul(t[3][1], cpx{S1, -C1});
	t[3][2] = cmul(t[3][2], cpx{-R2, -R2});
	t[3][3] = cmul(t[3][3], cpx{-C1, S1});
#pragma unroll
	for (int k0 = 0; k0 < 4; ++k0) {
		cpx a = t[0][k0], b = t[1][k0], c = t[2][k0], d = t[3][k0];
		dft4(a, b, c, d);        // over n0 -> k1 = 0..3
		o[0 + k0] = a;
		o[4 + k0] = b;
		o[8 + k0] = c;
		o[12 + k0] = d;
	}
}

// ---- warp-per-frame layout (mfcc.cu mfcc_frames_warp_kernel): one warp owns one frame; lane = (h, u), h = lane >> 4 the half,
// u = lane & 15; the half-warps run the sub-transforms r = h (round 0) and r = h + 2 (round 1).  All buffers are per warp:
//   T / Zs   4 tiles, tile r at r * WF_TILE complex elements.  Transpose tile: element (row, col) at row * 17 + col -- a PADDED row
//            (17 elements) instead of an XOR swizzle: the column writes of step A (lanes = rows) and the column reads of step B
//            (lanes = columns) still touch 16 distinct bank pairs, and every address is a per-lane base plus a compile-time
//            constant (the swizzle cost a LOP3 + LEA per access: ~350 integer instructions per frame);
//            sub-transform output Y_r[q] IN PLACE over the same tile, at (q >> 4) * 17 + (q & 15); the tile stride 276 = 4 (mod 16)
//            makes the four r of neighbouring bins fall into different bank pairs when the power stage reads Z[k] = Y_{k & 3}[k >> 2]
//            with consecutive k per lane (tests/test_fft_schedule_cpu.py counts the wavefronts of every access pattern)
//   P        power spectrum, bins 0..1024, bin k at k + 4 (k >> 4): every run of 16 bins is skewed by 16 bytes, so that the
//            128-bit reads of the mel stage (lanes on segments that start 16 bins apart) spread over the banks
constexpr int WF_TROW = 17;
constexpr int WF_TILE = 276;
constexpr int WF_PLEN = 1296;           // floats: skewed bins 0..1024 (1024 + 4 * 64 + 1) plus the 16-bin reach of a mel segment that starts on the last 4-bin group
SRB_HD int wf_t_index(int row, int col) { return row * WF_TROW + col; }
SRB_HD int wf_z_index(int r, int q) { return r * WF_TILE + q + (q >> 4); }
SRB_HD int wf_p_index(int k) { return k + 4 * (k >> 4); }

// step A for (r, u): z[n2] = packed frame element u + 16 n2 (loaded once by the caller and used for both rounds).  The pre-twiddle
// W_1024^{(u + 16 n2) r} is split into W_64^{n2 r} -- cr[n2], the same for all 16 lanes of a half-warp: one broadcast wavefront per
// load -- and W_1024^{u r}, which is linear in the transform and merges with the inter-step twiddle W_256^{u q2} into ONE table
// tTr[q2 * 64 + u] = W_1024^{u (r + 4 q2)} (the caller passes row r of the [64][16] table W_1024^{u j}); result into tile r of T
SRB_HD void wf_step_a(const cpx (&z)[16], const cpx *cr, const cpx *tTr, cpx *Ttile, int u) {
	cpx in[16], out[16];
#pragma unroll
	for (int n2 = 0; n2 < 16; ++n2) in[n2] = cmul(z[n2], cr[n2]);
	dft16(in, out);
#pragma unroll
	for (int q2 = 0; q2 < 16; ++q2) Ttile[wf_t_index(u, q2)] = cmul(out[q2], tTr[q2 * 64 + u]);
}
// step B for (r, u), in place: all 16 lanes of the half-warp load, (warp barrier), then transform and store Y_r[16 q1 + u]
SRB_HD void wf_step_b_load(const cpx *Ttile, int u, cpx (&in)[16]) {
#pragma unroll
	for (int n1 = 0; n1 < 16; ++n1) in[n1] = Ttile[wf_t_index(n1, u)];
}
SRB_HD void wf_step_b_store(const cpx (&in)[16], cpx *Z, int r, int u) {
	cpx out[16];
	dft16(in, out);
#pragma unroll
	for (int q1 = 0; q1 < 16; ++q1) Z[r * WF_TILE + q1 * WF_TROW + u] = out[q1];      // = wf_z_index(r, 16 q1 + u), u < 16
}
// bins k and 1024 - k of the real 2048-point power spectrum from ONE pair of packed results: with e = Z[k] + conj(Z[1024-k]),
// o = Z[k] - conj(Z[1024-k]), w = exp(-2 pi i k / 2048), (a, b) = w o:
//     Y[k] = (e.x + b, e.y - a) / 2,        Y[1024 - k] = (e.x - b, -e.y - a) / 2          (Z[1024] := Z[0])
SRB_HD void wf_power_pair(const cpx *Z, int k, cpx w, float &pk, float &pmk) {
	const int km = (1024 - k) & 1023;
	const cpx zk = Z[wf_z_index(k & 3, k >> 2)];
	cpx zc = Z[wf_z_index(km & 3, km >> 2)];
	zc.y = -zc.y;
	const cpx e = cadd2(zk, zc), o = csub2(zk, zc);
	const cpx wo = cmul(w, o);
	const float yr = 0.5f * (e.x + wo.y), yi = 0.5f * (e.y - wo.x);
	const float mr = 0.5f * (e.x - wo.y), mi = 0.5f * (e.y + wo.x);
	pk = yr * yr + yi * yi;
	pmk = mr * mr + mi * mi;
}
// the whole power stage of one frame as the 32 lanes do it: lane l owns the pairs k = l + 32 i, i = 0..15 (k = 0..511), lane 0 also
// the self-paired bin 512; tw2[k] = exp(-2 pi i k / 2048), k = 0..512
SRB_HD void wf_power_lane(const cpx *Z, float *P, const cpx *tw2, int lane) {
#pragma unroll
	for (int i = 0; i < 16; ++i) {
		const int k = lane + 32 * i;
		float pk, pmk;
		wf_power_pair(Z, k, tw2[k], pk, pmk);
		P[wf_p_index(k)] = pk;
		P[wf_p_index(1024 - k)] = pmk;
	}
	if (lane == 0) {
		float pk, pmk;
		wf_power_pair(Z, 512, tw2[512], pk, pmk);
		P[wf_p_index(512)] = pk;
	}
}

}  // namespace srbfft
