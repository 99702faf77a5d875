// gmm_em.cu -- EM sufficient statistics and M-step for diagonal GMMs on sm_100a.
//
// Replaces the reference's hot loop #3:
//   GMMTrainerBaseline::iteration       src/gmm/src/gmm.cc:439-531  (materialises a K x N float64 responsibility
//                                        matrix -- 4.1 GB at K=512, N=1e6 -- and sweeps it three times)
//   update_weights / means / variance   src/gmm/src/gmm.cc:388-437
//   GMMUBMTrainerBaseline::update_*     src/gmm/src/gmmubm.cc:40-81  (MAP, means only, relevance 16)
//
// The responsibility matrix is never materialised.  An E-step is two passes over the frames:
//   (A) gmm_score_kernel (gmm_score.cu) with frame_log2 = 1: per-frame normaliser log2 p(x_t) and the total LL;
//   (B) em_accumulate_kernel (here): one CTA owns one 32-mixture parameter tile (loaded once by cp.async.bulk) and
//       a contiguous frame range.  Per 256-frame step each warp (64 frames)
//         phase 1  recomputes the exponents of its 64 x 32 (frame, mixture) pairs with the same 8x8 register tile as
//                  the scorer, turns them into responsibilities gamma = 2^(l - lse) and parks them in its private
//                  shared-memory tile;
//         phase 2  accumulates  N_k += gamma,  Sx_kd += gamma x,  Sxx_kd += gamma x^2  with a 4-mixture x 10-dim x
//                  2-moment register tile per lane (lanes = 8 mixture groups x 4 dim groups, 4 frames per LDS.128).
//       fp32 register accumulators are flushed into float64 per-warp partials every 16 steps (1024 frames per
//       warp) so the error of a sum over 1e6 frames stays at the 1e-6 relative level; a second kernel reduces the
//       partials deterministically into the float64 statistics vector the C ABI exposes.
//   Multi-GPU: each rank runs (A)+(B) on its frame shard, the host all-reduces the statistics (NCCL), and every
//   rank runs the same float64 M-step kernel -- no broadcast needed.
//
// Guards reproduced from the reference: frames whose likelihood underflows float64 contribute nothing
// (prob_sum <= 0 -> 1e-15, gmm.cc:482-490: gamma = 0/1e-15 = 0); N_k == 0 -> 1e-6 (gmm.cc:502-509).
#include "kernels.h"

namespace srb {

constexpr int EM_THREADS = 128;
constexpr int EM_FT = 256;             // frames per CTA step (64 per warp)
constexpr int EM_XS = 264;             // padded row stride of the x tile (floats): rows 4j+dg land on distinct banks
constexpr int EM_DP = 40;              // dims per pass of the phase-2 tile (4 dim groups x 10)
constexpr int EM_GS = 36;              // padded row stride of a warp's gamma tile [64][32]
constexpr int EM_FLUSH = 16;           // steps between float64 flushes
constexpr float LN2F_EM = 0.69314718055994531f;
constexpr float DENORM_LOG_MIN_EM = -744.4400719213812f;

__device__ __forceinline__ unsigned long long em_pack2(float lo, float hi) {
	unsigned long long r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ void em_unpack2(unsigned long long v, float &lo, float &hi) {
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long em_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
	unsigned long long d;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));   // SASS FFMA2: two fp32 FMAs per issue slot
	return d;
}

struct EmArgs {
	const float *X;
	int ld, D, n;
	const float *tiles;        // packed tiles of ONE model
	int n_ktiles;              // Kpad / 32
	int n_split;               // CTAs along the frame axis
	int dpass;                 // which block of 40 dims this launch accumulates (D > 40 -> several launches)
	const float *lse2;         // [n] log2 p(x_t)
	double *partials;          // [n_split][4 warps][n_ktiles][32][84]
	int clamp;
	int x_vec_ok;
};

template <bool P1, bool P2>
__global__ void __launch_bounds__(EM_THREADS, 2) em_accumulate_kernel(const EmArgs A) {
	extern __shared__ __align__(128) unsigned char em_smem[];
	const int D = A.D;
	const int tileF = 64 * D + 32;
	const int xrows = ((D + EM_DP - 1) / EM_DP) * EM_DP;                // rows of the x tile: whole 40-dim passes, rows >= D are zero
	float *ps = reinterpret_cast<float *>(em_smem);                     // [tileF]
	float *xs = ps + tileF;                                             // [xrows][EM_XS]
	float *gs = xs + (size_t)xrows * EM_XS;                             // [4 warps][64][EM_GS]
	float *ls = gs + 4 * 64 * EM_GS;                                    // [256] lse2 of the step's frames (+inf marks "no contribution")
	uint64_t *bar = reinterpret_cast<uint64_t *>(ls + EM_FT);

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int g = lane >> 2, q = lane & 3;       // phase-1 lane roles
	const int kg = lane & 7, dg = lane >> 3;     // phase-2 lane roles: mixtures 4kg..4kg+3, dims {4j + dg}
	const int ktile = blockIdx.x, split = blockIdx.y;

	if (tid == 0) {
		mbar_init(bar, 1);
		fence_mbar_init();
	}
	__syncthreads();
	if (tid == 0) {
		mbar_expect_tx(bar, (uint32_t)tileF * 4u);
		bulk_g2s(ps, A.tiles + (size_t)ktile * tileF, (uint32_t)tileF * 4u, bar);
	}
	// zero the x rows beyond D once (phase 2 reads 40 rows per pass)
	for (int i = tid; i < (xrows - D) * EM_XS; i += EM_THREADS) xs[(size_t)D * EM_XS + i] = 0.f;

	// frame range of this CTA: steps of 256 frames, split evenly
	const int n_steps_total = (A.n + EM_FT - 1) / EM_FT;
	const int per = (n_steps_total + A.n_split - 1) / A.n_split;
	const int step0 = split * per;
	const int step1 = min(n_steps_total, step0 + per);

	// phase-2 accumulators: acc2[kk][j] = {sum gamma x, sum gamma x^2} for mixture 4kg+kk, dim 4j+dg; accn[kk] = sum gamma
	unsigned long long acc2[4][10];
	float accn[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
	for (int kk = 0; kk < 4; ++kk)
#pragma unroll
		for (int j = 0; j < 10; ++j) acc2[kk][j] = 0ull;
	double *my_part = A.partials + ((((size_t)split * 4 + warp) * A.n_ktiles + ktile) * 32) * 84;

	mbar_wait(bar, 0);
	float cneg[8];
	{
		const float *cp = ps + 64 * D;
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const float2 c = *reinterpret_cast<const float2 *>(cp + 2 * (4 * j + q));
			cneg[2 * j] = -c.x;
			cneg[2 * j + 1] = -c.y;
		}
	}
	const float *pq = ps + 4 * q;
	float *gw = gs + (size_t)warp * 64 * EM_GS;
	const int d_lo = A.dpass * EM_DP;

	int since_flush = 0;
	for (int step = step0; step < step1; ++step) {
		const int f0 = step * EM_FT;
		__syncthreads();   // everyone is done reading xs / ls of the previous step
		for (int f = tid; f < EM_FT; f += EM_THREADS) {
			const int fr = f0 + f;
			if (fr < A.n) {
				const float *row = A.X + (size_t)fr * A.ld;
				if (A.x_vec_ok) {
					const float4 *row4 = reinterpret_cast<const float4 *>(row);
					const int n4 = D >> 2;
					for (int d4 = 0; d4 < n4; ++d4) {
						const float4 v = __ldg(row4 + d4);
						xs[(4 * d4 + 0) * EM_XS + f] = v.x;
						xs[(4 * d4 + 1) * EM_XS + f] = v.y;
						xs[(4 * d4 + 2) * EM_XS + f] = v.z;
						xs[(4 * d4 + 3) * EM_XS + f] = v.w;
					}
					for (int d = n4 * 4; d < D; ++d) xs[d * EM_XS + f] = __ldg(row + d);
				} else {
					for (int d = 0; d < D; ++d) xs[d * EM_XS + f] = __ldg(row + d);
				}
				const float l2 = __ldg(A.lse2 + fr);
				// reference guard: a frame whose total probability underflows float64 gives gamma = 0 for every mixture
				const bool dead = A.clamp && !(l2 * LN2F_EM >= DENORM_LOG_MIN_EM);
				ls[f] = dead ? __int_as_float(0x7f800000) : l2;
			} else {
				for (int d = 0; d < D; ++d) xs[d * EM_XS + f] = 0.f;
				ls[f] = __int_as_float(0x7f800000);   // +inf -> gamma = 2^(-inf) = 0
			}
		}
		__syncthreads();

		// ---------------- phase 1: exponents of (64 frames of this warp) x (32 mixtures) ----------------
		{
			float a8[8][8];
			if (P1) {
				unsigned long long a2[8][4];
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const unsigned long long c2 = em_pack2(cneg[2 * j], cneg[2 * j + 1]);
#pragma unroll
					for (int f = 0; f < 8; ++f) a2[f][j] = c2;
				}
				const float *xbase = xs + warp * 64 + 4 * g;
#pragma unroll 2
				for (int d = 0; d < D; ++d) {
					const float4 xa = *reinterpret_cast<const float4 *>(xbase + d * EM_XS);
					const float4 xb = *reinterpret_cast<const float4 *>(xbase + d * EM_XS + 32);
					const unsigned long long x2[8] = {em_pack2(xa.x, xa.x), em_pack2(xa.y, xa.y), em_pack2(xa.z, xa.z), em_pack2(xa.w, xa.w),
					                                  em_pack2(xb.x, xb.x), em_pack2(xb.y, xb.y), em_pack2(xb.z, xb.z), em_pack2(xb.w, xb.w)};
#pragma unroll
					for (int j = 0; j < 4; ++j) {
						const ulonglong2 p = *reinterpret_cast<const ulonglong2 *>(pq + d * 64 + 16 * j);   // {a0,a1}, {nb0,nb1}
#pragma unroll
						for (int f = 0; f < 8; ++f) {
							const unsigned long long t2 = em_fma2(x2[f], p.x, p.y);
							a2[f][j] = em_fma2(t2, t2, a2[f][j]);
						}
					}
				}
#pragma unroll
				for (int f = 0; f < 8; ++f)
#pragma unroll
					for (int j = 0; j < 4; ++j) em_unpack2(a2[f][j], a8[f][2 * j], a8[f][2 * j + 1]);
			} else {
#pragma unroll
				for (int f = 0; f < 8; ++f)
#pragma unroll
					for (int m = 0; m < 8; ++m) a8[f][m] = cneg[m];
				const float *xbase = xs + warp * 64 + 4 * g;
#pragma unroll 2
				for (int d = 0; d < D; ++d) {
					const float4 xa = *reinterpret_cast<const float4 *>(xbase + d * EM_XS);
					const float4 xb = *reinterpret_cast<const float4 *>(xbase + d * EM_XS + 32);
					const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
					for (int j = 0; j < 4; ++j) {
						const float4 p = *reinterpret_cast<const float4 *>(pq + d * 64 + 16 * j);   // a0 a1 nb0 nb1
#pragma unroll
						for (int f = 0; f < 8; ++f) {
							const float t0 = fmaf(xv[f], p.x, p.z);
							a8[f][2 * j] = fmaf(t0, t0, a8[f][2 * j]);
							const float t1 = fmaf(xv[f], p.y, p.w);
							a8[f][2 * j + 1] = fmaf(t1, t1, a8[f][2 * j + 1]);
						}
					}
				}
			}
			// gamma = 2^(-acc - lse2); park in the warp's tile gw[frame][mixture]
#pragma unroll
			for (int f = 0; f < 8; ++f) {
				const int fl = 4 * g + (f & 3) + ((f >> 2) << 5);       // frame within the warp's 64
				const float l = ls[warp * 64 + fl];
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					float2 gm;
					gm.x = ex2_approx(-a8[f][2 * j] - l);
					gm.y = ex2_approx(-a8[f][2 * j + 1] - l);
					*reinterpret_cast<float2 *>(gw + fl * EM_GS + 2 * (4 * j + q)) = gm;
				}
			}
		}
		__syncwarp();

		// ---------------- phase 2: accumulate the warp's 64 frames into the 4x10x2 (+4) register tile ----------------
		{
			const float *xw = xs + (size_t)d_lo * EM_XS + warp * 64;
#pragma unroll 1
			for (int f4 = 0; f4 < 16; ++f4) {
				float4 gmm[4];
#pragma unroll
				for (int i = 0; i < 4; ++i) gmm[i] = *reinterpret_cast<const float4 *>(gw + (4 * f4 + i) * EM_GS + 4 * kg);
				if (dg == 0) {
#pragma unroll
					for (int i = 0; i < 4; ++i) {
						accn[0] += gmm[i].x;
						accn[1] += gmm[i].y;
						accn[2] += gmm[i].z;
						accn[3] += gmm[i].w;
					}
				}
#pragma unroll
				for (int j = 0; j < 10; ++j) {
					const float4 xq = *reinterpret_cast<const float4 *>(xw + (size_t)(4 * j + dg) * EM_XS + 4 * f4);
					const float xv[4] = {xq.x, xq.y, xq.z, xq.w};
#pragma unroll
					for (int i = 0; i < 4; ++i) {
						if (P2) {
							const unsigned long long xx = em_pack2(xv[i], xv[i] * xv[i]);          // {x, x^2}
							acc2[0][j] = em_fma2(em_pack2(gmm[i].x, gmm[i].x), xx, acc2[0][j]);
							acc2[1][j] = em_fma2(em_pack2(gmm[i].y, gmm[i].y), xx, acc2[1][j]);
							acc2[2][j] = em_fma2(em_pack2(gmm[i].z, gmm[i].z), xx, acc2[2][j]);
							acc2[3][j] = em_fma2(em_pack2(gmm[i].w, gmm[i].w), xx, acc2[3][j]);
						} else {
							const float x1 = xv[i], x2 = xv[i] * xv[i];
							const float gk[4] = {gmm[i].x, gmm[i].y, gmm[i].z, gmm[i].w};
#pragma unroll
							for (int kk = 0; kk < 4; ++kk) {
								float sx, sxx;
								em_unpack2(acc2[kk][j], sx, sxx);
								sx = fmaf(gk[kk], x1, sx);
								sxx = fmaf(gk[kk], x2, sxx);
								acc2[kk][j] = em_pack2(sx, sxx);
							}
						}
					}
				}
			}
		}
		__syncwarp();

		if (++since_flush == EM_FLUSH || step + 1 == step1) {
			// float64 flush into this warp's private partial slot: row (4kg + kk), columns: dims {4j+dg} as (x, x^2), then N_k
#pragma unroll
			for (int kk = 0; kk < 4; ++kk) {
				double *rowp = my_part + (size_t)(4 * kg + kk) * 84;
#pragma unroll
				for (int j = 0; j < 10; ++j) {
					const int dd = 4 * j + dg;
					float sx, sxx;
					em_unpack2(acc2[kk][j], sx, sxx);
					rowp[2 * dd] += (double)sx;
					rowp[2 * dd + 1] += (double)sxx;
					acc2[kk][j] = 0ull;
				}
				if (dg == 0) {
					rowp[80] += (double)accn[kk];
					accn[kk] = 0.f;
				}
			}
			since_flush = 0;
		}
	}
}

// stats[0..K) = N_k ; stats[K + k*D + d] = Sx ; stats[K + K*D + k*D + d] = Sxx  (for the dims of pass `dpass`)
__global__ void em_reduce_kernel(const double *__restrict__ partials, int n_slots, int n_ktiles, int K, int D, int dpass,
                                 double *__restrict__ stats) {
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over Kpad * 84
	const int Kpad = n_ktiles * 32;
	if (idx >= Kpad * 84) return;
	const int row = idx / 84, col = idx - row * 84;
	if (row >= K) return;
	double s = 0.0;
	for (int sl = 0; sl < n_slots; ++sl) s += partials[((size_t)sl * Kpad + row) * 84 + col];
	if (col == 80) {
		if (dpass == 0) stats[row] = s;
	} else if (col < 80) {
		const int d = dpass * EM_DP + (col >> 1);
		if (d < D) stats[K + (size_t)(col & 1) * K * D + (size_t)row * D + d] = s;
	}
}

__global__ void em_tail_kernel(const double *__restrict__ ll_sum, const float *__restrict__ lse2, int n, int clamp,
                               double *__restrict__ stats, int K, int D) {
	// stats tail: [sum log p] [frames used].  The count is a sum of small integers: exact (and so order-independent) in float64.
	__shared__ int sh[256];
	int cnt = 0;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const float l2 = lse2[i];
		const bool dead = clamp && !(l2 * LN2F_EM >= DENORM_LOG_MIN_EM);
		cnt += dead ? 0 : 1;
	}
	sh[threadIdx.x] = cnt;
	__syncthreads();
	for (int w = blockDim.x / 2; w > 0; w >>= 1) {
		if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		if (gridDim.x == 1) stats[K + 2 * (size_t)K * D + 1] = (double)sh[0];
		else atomicAdd(&stats[K + 2 * (size_t)K * D + 1], (double)sh[0]);      // zeroed by the launcher
		if (blockIdx.x == 0) stats[K + 2 * (size_t)K * D] = ll_sum[0];
	}
}

int launch_em_tail(const double *d_ll_sum, const float *d_lse2, int n, int clamp, double *d_stats, int K, int D, cudaStream_t st) {
	const int blocks = n <= 32768 ? 1 : std::min(ctx().sm_count * 4, ceil_div(n, 2048));      // short utterances (batched enrolment): one block, no memset
	if (blocks > 1) SRB_CUDA(cudaMemsetAsync(d_stats + K + 2 * (size_t)K * D + 1, 0, sizeof(double), st));
	em_tail_kernel<<<blocks, 256, 0, st>>>(d_ll_sum, d_lse2, n, clamp, d_stats, K, D);
	SRB_CUDA(cudaGetLastError());
	return 0;
}

// ---- M-step (float64), one thread per (k, d); params = w[K] | mean[K*D] | sigma[K*D] ----
// mode 0: gmm.cc:388-437   w_k = N_k/n (renormalised), mean = Sx/N_k, sigma = max(sqrt(min_covar), sqrt(E[(x-mean)^2]))
// mode 1: gmmubm.cc:40-81  mean = alpha*Sx/N_k + (1-alpha)*mean_ubm, alpha = N_k/(N_k+16); w, sigma untouched
__global__ void em_mstep_kernel(double *__restrict__ params, const double *__restrict__ ubm, int K, int D,
                                const double *__restrict__ stats, double n_total, double min_sigma, int mode) {
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= K * D) return;
	const int k = idx / D;
	double Nk = stats[k];
	if (Nk == 0.0) Nk = 1e-6;                                   // gmm.cc:508-509
	const double sx = stats[K + idx], sxx = stats[K + (size_t)K * D + idx];
	double *mean = params + K, *sigma = params + K + (size_t)K * D;
	if (mode == 0) {
		const double mu = sx * (1.0 / Nk);                      // gmm.cc:408
		// sum g (x - mu)^2 = Sxx - 2 mu Sx + mu^2 N   (gmm.cc:423-429 with the NEW mean)
		double v = (sxx - 2.0 * mu * sx + mu * mu * stats[k]) * (1.0 / Nk);
		if (v < 0.0) v = 0.0;
		const double s = sqrt(v);
		mean[idx] = mu;
		sigma[idx] = s > min_sigma ? s : min_sigma;             // gmm.cc:430-433
	} else {
		const double alpha = Nk / (Nk + 16.0);                  // gmm.hh:110-112, gmmubm.cc:57
		mean[idx] = sx * (1.0 / Nk * alpha) + ubm[K + idx] * (1.0 - alpha);   // gmmubm.cc:66-69
	}
}

__global__ void em_weights_kernel(double *__restrict__ params, int K, const double *__restrict__ stats, double n_total) {
	// gmm.cc:388-394 + normalize_weights :363-369, single block, deterministic
	__shared__ double sh[256];
	double s = 0.0;
	for (int k = threadIdx.x; k < K; k += blockDim.x) {
		double Nk = stats[k];
		if (Nk == 0.0) Nk = 1e-6;
		s += Nk / n_total;
	}
	sh[threadIdx.x] = s;
	__syncthreads();
	for (int w = blockDim.x / 2; w > 0; w >>= 1) {
		if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
		__syncthreads();
	}
	const double wsum = sh[0];
	for (int k = threadIdx.x; k < K; k += blockDim.x) {
		double Nk = stats[k];
		if (Nk == 0.0) Nk = 1e-6;
		params[k] = (Nk / n_total) / wsum;
	}
}

// ---- repack the float64 master parameters into the fp32 tile layout (gmm_model.h), one thread per (row, d) ----
__global__ void em_pack_kernel(const double *__restrict__ params, int K, int D, int Kpad, float *__restrict__ tiles) {
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	const int tileF = 64 * D + 32;
	if (idx < Kpad * D) {
		const int r = idx / D, d = idx - r * D;
		const int tile = r >> 5, m = r & 31;
		float a = 0.f, nb = 0.f;
		if (r < K) {
			const double sg = params[K + (size_t)K * D + (size_t)r * D + d];
			const double mu = params[K + (size_t)r * D + d];
			const double ad = 0.84932180028801904272 / sg;       // sqrt(log2(e) / 2)
			a = (float)ad;
			nb = (float)(-mu * ad);
		}
		float *t = tiles + (size_t)tile * tileF + (size_t)d * 64 + (m >> 1) * 4 + (m & 1);   // chunk = [a0, a1, nb0, nb1]
		t[0] = a;
		t[2] = nb;
	}
	if (idx < Kpad) {
		const int r = idx, tile = r >> 5, m = r & 31;
		float c2 = PAD_C2;
		if (r < K) {
			const double w = params[r];
			double c = w > 0.0 ? log2(w) : -1.0e30;
			for (int d = 0; d < D; ++d) c -= log2(2.5066282746310002 * params[K + (size_t)K * D + (size_t)r * D + d]);
			c2 = c < -1.0e30 ? PAD_C2 : (float)c;
		}
		tiles[(size_t)tile * tileF + 64 * D + m] = c2;
	}
}

// ------------------------------------------------------------------------------------------------
static DevBuf g_em_partials;
static bool g_em_attr_set = false;
static int g_em_variant = 2;   // default: FFMA2 in phase 1, scalar FFMA in phase 2 (measured fastest: 7.86 vs 8.07 / 8.64 / 8.66 ms at C3)

void em_release() {
	g_em_partials.release();
	em_tc_release();
}

static size_t em_smem_bytes(int D) {
	const int xrows = ((D + EM_DP - 1) / EM_DP) * EM_DP;
	return (size_t)(64 * D + 32) * 4 + (size_t)xrows * EM_XS * 4 + (size_t)4 * 64 * EM_GS * 4 + EM_FT * 4 + 16;
}

int launch_estep_accumulate(const PackedModels &pm, const float *dX, int n, int ld, int D, int K, const float *d_lse2,
                            const double *d_ll_sum, double *d_stats, int clamp, cudaStream_t st) {
	SRB_REQUIRE(pm.S == 1 && pm.d_tiles != nullptr, "em: model not packed");
	SRB_REQUIRE(D == pm.D && D >= 1 && D <= MAX_DIM, "em: bad dim %d", D);
	SRB_REQUIRE(n >= 1, "em: no frames");
	Context &c = ctx();
	const size_t smem = em_smem_bytes(D);
	SRB_REQUIRE(smem <= 227 * 1024, "em: dim %d needs %zu B of shared memory", D, smem);
	if (!g_em_attr_set) {
		SRB_CUDA(cudaFuncSetAttribute(em_accumulate_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		SRB_CUDA(cudaFuncSetAttribute(em_accumulate_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		SRB_CUDA(cudaFuncSetAttribute(em_accumulate_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		SRB_CUDA(cudaFuncSetAttribute(em_accumulate_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		const char *ev = getenv("SRB200_EM_VARIANT");     // A/B switch: two digits, phase 1 / phase 2 packed (FFMA2) = 1
		if (ev && (ev[0] == '0' || ev[0] == '1') && (ev[1] == '0' || ev[1] == '1')) g_em_variant = (ev[0] - '0') * 2 + (ev[1] - '0');
		g_em_attr_set = true;
	}
	const int n_ktiles = pm.tiles_per_model;
	const int n_steps = ceil_div(n, EM_FT);
	int per_sm = (int)((227 * 1024) / (smem + 1024));
	if (per_sm > 2) per_sm = 2;
	if (per_sm < 1) per_sm = 1;
	int n_split = (c.sm_count * per_sm) / n_ktiles;
	if (n_split < 1) n_split = 1;
	if (n_split > n_steps) n_split = n_steps;
	// each split should own a whole number of steps; avoid empty trailing splits
	{
		const int per = ceil_div(n_steps, n_split);
		n_split = ceil_div(n_steps, per);
	}
	const int n_slots = n_split * 4;
	const size_t part_bytes = (size_t)n_slots * n_ktiles * 32 * 84 * sizeof(double);
	if (g_em_partials.ensure(part_bytes)) return 1;
	const int n_dpass = ceil_div(D, EM_DP);
	SRB_CUDA(cudaMemsetAsync(d_stats, 0, (size_t)em_stats_len(K, D) * sizeof(double), st));
	for (int dp = 0; dp < n_dpass; ++dp) {
		SRB_CUDA(cudaMemsetAsync(g_em_partials.p, 0, part_bytes, st));
		EmArgs a;
		a.X = dX;
		a.ld = ld;
		a.D = D;
		a.n = n;
		a.tiles = pm.d_tiles;
		a.n_ktiles = n_ktiles;
		a.n_split = n_split;
		a.dpass = dp;
		a.lse2 = d_lse2;
		a.partials = g_em_partials.as<double>();
		a.clamp = clamp;
		a.x_vec_ok = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(dX) & 15) == 0);
		dim3 grid(n_ktiles, n_split);
		switch (g_em_variant) {
			case 3: em_accumulate_kernel<true, true><<<grid, EM_THREADS, smem, st>>>(a); break;
			case 2: em_accumulate_kernel<true, false><<<grid, EM_THREADS, smem, st>>>(a); break;
			case 1: em_accumulate_kernel<false, true><<<grid, EM_THREADS, smem, st>>>(a); break;
			default: em_accumulate_kernel<false, false><<<grid, EM_THREADS, smem, st>>>(a); break;
		}
		SRB_CUDA(cudaGetLastError());
		const int total = n_ktiles * 32 * 84;
		em_reduce_kernel<<<ceil_div(total, 256), 256, 0, st>>>(g_em_partials.as<double>(), n_slots, n_ktiles, K, D, dp, d_stats);
		SRB_CUDA(cudaGetLastError());
	}
	return launch_em_tail(d_ll_sum, d_lse2, n, clamp, d_stats, K, D, st);
}

int launch_mstep(double *d_params, const double *d_ubm_params, int K, int D, const double *d_stats, long long n_total,
                 double min_covar, int mode, cudaStream_t st) {
	SRB_REQUIRE(mode == 0 || (mode == 1 && d_ubm_params != nullptr), "mstep: MAP mode needs the UBM parameters");
	if (mode == 0) {
		em_weights_kernel<<<1, 256, 0, st>>>(d_params, K, d_stats, (double)n_total);
		SRB_CUDA(cudaGetLastError());
	}
	em_mstep_kernel<<<ceil_div(K * D, 256), 256, 0, st>>>(d_params, d_ubm_params, K, D, d_stats, (double)n_total, sqrt(min_covar), mode);
	SRB_CUDA(cudaGetLastError());
	return 0;
}

int launch_pack_from_master(const double *d_params, int K, int D, PackedModels &pm, cudaStream_t st) {
	const int Kpad = ceil_div(K, MT) * MT;
	const size_t bytes = (size_t)(Kpad / MT) * tile_floats(D) * sizeof(float);
	if (pm.d_tiles == nullptr || pm.bytes < bytes) {
		pm.release();
		SRB_CUDA(cudaMalloc(&pm.d_tiles, bytes));
		pm.bytes = bytes;
	}
	// the tensor-core images (built on the host from the PREVIOUS parameters) no longer describe this model: drop them, so that
	// scoring falls back to the FMA tiles written below until gmm_ensure_packed rebuilds the images from the pulled master
	if (pm.tc_built || pm.d_tc2) {
		if (pm.d_tc2) SRB_CUDA(cudaStreamSynchronize(st));      // a scoring launch may still be reading the device images
		pm.release_tc();
	}
	pm.S = 1;
	pm.K = K;
	pm.Kpad = Kpad;
	pm.D = D;
	pm.tiles_per_model = Kpad / MT;
	const int total = Kpad * D;
	em_pack_kernel<<<ceil_div(total, 256), 256, 0, st>>>(d_params, K, D, Kpad, pm.d_tiles);
	SRB_CUDA(cudaGetLastError());
	return 0;
}

}  // namespace srb
