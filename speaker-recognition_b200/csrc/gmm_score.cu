// gmm_score.cu -- diagonal-GMM log-likelihood scoring for sm_100a.
//
// Replaces the reference's hot loop #2:
//   Gaussian::probability_of_fast_exp      src/gmm/src/gmm.cc:176-202   (per-dim exp, linear-domain product)
//   GMM::log_probability_of_fast_exp       src/gmm/src/gmm.cc:237-244   (weighted sum, safe_log)
//   threaded_log_probability_of            src/gmm/src/gmm.cc:533-569   (thread pool over frames, serial sum)
// and the Python loop over speakers of GMMSet.predict_one_scores (src/testbench/gmmset.py:59-64).
//
// Formulation (log2 domain, no cancellation -- SURVEY.md s7 hard part 2a):
//   t = x_d * a_kd + nb_kd            a = sqrt(log2e/2)/sigma, nb = -mean*a       (1 FFMA)
//   acc_k = -c2_k + sum_d t*t                                                     (1 FFMA)
//   log2 p(x) = log2 sum_k 2^(-acc_k)    via an online max/sum over 32-row mixture tiles
// One CTA = 128 threads = 4 warps owns a tile of 256 frames (x tile transposed in shared memory for the whole
// CTA lifetime) and walks the packed parameter tiles of its speaker chunk, which a single elected thread streams
// with cp.async.bulk (TMA engine) into a 3-stage full/empty mbarrier ring (no CTA-wide barrier in the tile loop: warps
// drift by up to two tiles, which keeps the FMA pipe of every SM sub-partition busy while a sibling warp is in its
// log-sum-exp epilogue).  Each thread keeps an 8-frame x 8-mixture
// register tile: per dimension 2 LDS.128 (frames) + 4 LDS.128 (broadcast parameters) feed 128 FFMA.
// The kernel is bound by the fp32 FMA pipe (2*D FFMA per frame-mixture pair); HBM traffic is the x tile once
// and the parameters from L2 (DESIGN.md "roofline").
#include "gmm_model.h"
#include <algorithm>
#include "kernels.h"
#include "gmm_score_tc_layout.h"

namespace srb {

constexpr int SC_MAX_WARPS = 8;     // CTA = 4 warps x 64 frames (3 CTAs/SM) or 8 warps x 64 frames (2 CTAs/SM, <=128 registers)
constexpr int SC_STAGES = 3;
constexpr float LN2F = 0.69314718055994531f;
constexpr float DENORM_LOG_MIN_F = -744.4400719213812f;   // log(4.94e-324): the reference's linear-domain sum is 0 below this
constexpr float SAFE_LOG_FLOOR_F = -34.538776394910684f;  // log(1e-15), gmm.cc:34-38

struct ScoreArgs {
	const float *X;
	int ld, D;
	const float *tiles;        // packed parameter tiles
	int tiles_per_model, S, spk_per_chunk, n_chunks;
	const int4 *ftiles;        // per frame tile: {first frame, frame count, utterance, 0}
	float *frame_ll;           // optional [S][ldf]
	int ldf;
	double *partial;           // [n_frame_tiles][S]
	int clamp;
	int frame_log2;            // frame_ll receives the unclamped log2-domain value (E-step normaliser) instead of ln
	int stagger_cycles;        // first-wave CTAs of the same SM start this many cycles apart (see kernel)
	int n_sm;
	int x_vec_ok;              // X rows are 16-B aligned and ld % 4 == 0
};

// packed fp32x2 FMA (SASS FFMA2, sm_100+): two lanes per instruction.  The scorer is issue-bound with scalar FFMA
// (ncu r1: issue slots 88 % busy, FMA pipe 76 %); FFMA2 halves the FMA instruction count so that the LDS / MUFU /
// ALU instructions of the loop and of the log-sum-exp epilogue no longer take issue slots away from the FMA pipe.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
	unsigned long long r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) {
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
	unsigned long long d;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
	return d;
}

template <bool PACKED, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 4 ? 3 : 2) gmm_score_kernel(const ScoreArgs A) {
	constexpr int SC_THREADS = WARPS * 32;
	constexpr int SC_FT = WARPS * 64;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int D = A.D;
	const int tileF = 64 * D + 32;
	float *xs = reinterpret_cast<float *>(smem_raw);                  // [D][256]
	float *ps = xs + (size_t)D * SC_FT;                               // [STAGES][tileF]
	uint64_t *bars = reinterpret_cast<uint64_t *>(ps + (size_t)SC_STAGES * tileF);   // full[STAGES] then empty[STAGES]
	uint64_t *empty = bars + SC_STAGES;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int g = lane >> 2, q = lane & 3;
	// 1-D grid, speaker chunk fastest: CTAs that share a frame tile are launched back to back, so the tile's features
	// are read from HBM once and served from L2 to the other chunks (ncu r1 with the opposite order: 474 MB of DRAM
	// reads per launch against 104 MB algorithmic)
	const int ftile = blockIdx.x / A.n_chunks;
	const int4 ft = A.ftiles[ftile];
	const int f0 = ft.x, nf = ft.y;
	const int s_begin = (blockIdx.x - ftile * A.n_chunks) * A.spk_per_chunk;
	const int s_end = min(A.S, s_begin + A.spk_per_chunk);
	const int n_tiles = (s_end - s_begin) * A.tiles_per_model;
	const float *gt0 = A.tiles + (size_t)s_begin * A.tiles_per_model * tileF;
	const uint32_t tile_bytes = (uint32_t)tileF * 4u;

	if (tid == 0) {
		for (int s = 0; s < SC_STAGES; ++s) {
			mbar_init(&bars[s], 1);
			mbar_init(&empty[s], WARPS);
		}
		fence_mbar_init();
	}
	__syncthreads();
	if (tid == 0) {
		const int pre = n_tiles < SC_STAGES ? n_tiles : SC_STAGES;
		for (int i = 0; i < pre; ++i) {
			mbar_expect_tx(&bars[i], tile_bytes);
			bulk_g2s(ps + (size_t)i * tileF, gt0 + (size_t)i * tileF, tile_bytes, &bars[i]);
		}
	}

	// ---- x tile: global [frame][ld] -> shared [d][frame]; each thread owns frames tid and tid+128 ----
	for (int f = tid; f < SC_FT; f += SC_THREADS) {
		if (f < nf) {
			const float *row = A.X + (size_t)(f0 + f) * A.ld;
			if (A.x_vec_ok) {
				const float4 *row4 = reinterpret_cast<const float4 *>(row);
				const int n4 = D >> 2;
				for (int d4 = 0; d4 < n4; ++d4) {
					const float4 v = __ldg(row4 + d4);
					xs[(4 * d4 + 0) * SC_FT + f] = v.x;
					xs[(4 * d4 + 1) * SC_FT + f] = v.y;
					xs[(4 * d4 + 2) * SC_FT + f] = v.z;
					xs[(4 * d4 + 3) * SC_FT + f] = v.w;
				}
				for (int d = n4 * 4; d < D; ++d) xs[d * SC_FT + f] = __ldg(row + d);
			} else {
				for (int d = 0; d < D; ++d) xs[d * SC_FT + f] = __ldg(row + d);
			}
		} else {
			for (int d = 0; d < D; ++d) xs[d * SC_FT + f] = 0.f;
		}
	}
	__syncthreads();

	// De-phase the CTAs that share an SM.  Every CTA does identical work at identical speed, so without this the warps
	// co-resident on an SM sub-partition (warp w of each of the 3 CTAs) march in lockstep: all of them sit in the
	// log-sum-exp epilogue at the same time and the FMA pipe idles for that quarter of every tile (ncu r1: pipe 77 %
	// busy, top stall math-pipe throttle).  Block b of the first wave lands on SM b % n_sm in slot b / n_sm; slot s
	// waits s thirds of a tile period once, later CTAs inherit the stagger from the CTAs they replace.
	if (A.stagger_cycles > 0) {
		const int lin = blockIdx.x;
		if (lin < 3 * A.n_sm) {
			const long long wait = (long long)(lin / A.n_sm) * A.stagger_cycles;
			const long long t0 = clock64();
			while (clock64() - t0 < wait) {
			}
		}
	}

	// thread's 8 frames: {warp*64 + 4g + i} and {warp*64 + 32 + 4g + i}, i = 0..3
	const float *xbase = xs + warp * 64 + 4 * g;
	float m_run[8], s_run[8];
#pragma unroll
	for (int f = 0; f < 8; ++f) {
		m_run[f] = -3.0e38f;
		s_run[f] = 0.f;
	}
	int spk = s_begin, tile_in_spk = 0;

	for (int it = 0; it < n_tiles; ++it) {
		const int stage = it % SC_STAGES;
		mbar_wait(&bars[stage], (uint32_t)((it / SC_STAGES) & 1));
		const float *pt = ps + (size_t)stage * tileF;

		float acc[8][8];
		const float *cp = pt + 64 * D;
		const float *pq = pt + 4 * q;
		if (PACKED) {
			// acc2[f][j] = {acc of mixture 2(4j+q), acc of mixture 2(4j+q)+1}
			unsigned long long acc2[8][4];
			{   // d = 0 peeled: the accumulators start as fma(t, t, -c2) instead of being initialised by 64 moves
				const float4 xa = *reinterpret_cast<const float4 *>(xbase);
				const float4 xb = *reinterpret_cast<const float4 *>(xbase + 32);
				const unsigned long long x2[8] = {pack2(xa.x, xa.x), pack2(xa.y, xa.y), pack2(xa.z, xa.z), pack2(xa.w, xa.w),
				                                  pack2(xb.x, xb.x), pack2(xb.y, xb.y), pack2(xb.z, xb.z), pack2(xb.w, xb.w)};
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const float2 c = *reinterpret_cast<const float2 *>(cp + 2 * (4 * j + q));
					const unsigned long long c2 = pack2(-c.x, -c.y);
					const ulonglong2 p = *reinterpret_cast<const ulonglong2 *>(pq + 16 * j);
#pragma unroll
					for (int f = 0; f < 8; ++f) {
						const unsigned long long t2 = fma2(x2[f], p.x, p.y);
						acc2[f][j] = fma2(t2, t2, c2);
					}
				}
			}
#pragma unroll 2
			for (int d = 1; d < D; ++d) {
				const float4 xa = *reinterpret_cast<const float4 *>(xbase + d * SC_FT);
				const float4 xb = *reinterpret_cast<const float4 *>(xbase + d * SC_FT + 32);
				const unsigned long long x2[8] = {pack2(xa.x, xa.x), pack2(xa.y, xa.y), pack2(xa.z, xa.z), pack2(xa.w, xa.w),
				                                  pack2(xb.x, xb.x), pack2(xb.y, xb.y), pack2(xb.z, xb.z), pack2(xb.w, xb.w)};
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					// {a_2c, a_2c+1} and {nb_2c, nb_2c+1} arrive as two aligned 64-bit halves of one LDS.128
					const ulonglong2 p = *reinterpret_cast<const ulonglong2 *>(pq + d * 64 + 16 * j);
#pragma unroll
					for (int f = 0; f < 8; ++f) {
						const unsigned long long t2 = fma2(x2[f], p.x, p.y);
						acc2[f][j] = fma2(t2, t2, acc2[f][j]);
					}
				}
			}
#pragma unroll
			for (int f = 0; f < 8; ++f)
#pragma unroll
				for (int j = 0; j < 4; ++j) unpack2(acc2[f][j], acc[f][2 * j], acc[f][2 * j + 1]);
		} else {
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const float2 c = *reinterpret_cast<const float2 *>(cp + 2 * (4 * j + q));
#pragma unroll
				for (int f = 0; f < 8; ++f) {
					acc[f][2 * j] = -c.x;
					acc[f][2 * j + 1] = -c.y;
				}
			}
#pragma unroll 2
			for (int d = 0; d < D; ++d) {
				const float4 xa = *reinterpret_cast<const float4 *>(xbase + d * SC_FT);
				const float4 xb = *reinterpret_cast<const float4 *>(xbase + d * SC_FT + 32);
				const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const float4 p = *reinterpret_cast<const float4 *>(pq + d * 64 + 16 * j);   // a0 a1 nb0 nb1
#pragma unroll
					for (int f = 0; f < 8; ++f) {
						const float t0 = fmaf(xv[f], p.x, p.z);
						acc[f][2 * j] = fmaf(t0, t0, acc[f][2 * j]);
						const float t1 = fmaf(xv[f], p.y, p.w);
						acc[f][2 * j + 1] = fmaf(t1, t1, acc[f][2 * j + 1]);
					}
				}
			}
		}
		// ---- online log2-sum-exp2 over this thread's 8 mixtures, per frame ----
#pragma unroll
		for (int f = 0; f < 8; ++f) {
			float mn = acc[f][0];
#pragma unroll
			for (int m = 1; m < 8; ++m) mn = fminf(mn, acc[f][m]);
			const float m_new = fmaxf(m_run[f], -mn);
			float sum = 0.f;
#pragma unroll
			for (int m = 0; m < 8; ++m) sum += ex2_approx(-acc[f][m] - m_new);
			s_run[f] = fmaf(s_run[f], ex2_approx(m_run[f] - m_new), sum);
			m_run[f] = m_new;
		}

		++tile_in_spk;
		const bool spk_done = (tile_in_spk == A.tiles_per_model);
		if (spk_done) {
			// merge the four mixture lanes (q) of each frame group, then finish the speaker
			double local = 0.0;
#pragma unroll
			for (int f = 0; f < 8; ++f) {
				float m = m_run[f], s = s_run[f];
#pragma unroll
				for (int off = 1; off <= 2; off <<= 1) {
					const float mo = __shfl_xor_sync(0xffffffffu, m, off);
					const float so = __shfl_xor_sync(0xffffffffu, s, off);
					const float M = fmaxf(m, mo);
					s = s * ex2_approx(m - M) + so * ex2_approx(mo - M);
					m = M;
				}
				const float l2 = m + lg2_approx(s);    // s in [1, K]: absolute error of lg2.approx < 3e-7, |log-likelihood| ~ 50
				float ll = l2 * LN2F;
				if (A.clamp && !(ll >= DENORM_LOG_MIN_F)) ll = SAFE_LOG_FLOOR_F;
				const int fr = warp * 64 + 4 * g + (f & 3) + ((f >> 2) << 5);
				if (fr < nf) {
					local += (double)ll;
					if (A.frame_ll != nullptr && q == 0) A.frame_ll[(size_t)spk * A.ldf + f0 + fr] = A.frame_log2 ? l2 : ll;
				}
				m_run[f] = -3.0e38f;
				s_run[f] = 0.f;
			}
			// sum over the 8 frame groups of the warp (lanes with equal q hold identical values); one partial per warp
#pragma unroll
			for (int off = 4; off <= 16; off <<= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
			if (lane == 0) A.partial[((size_t)ftile * A.S + spk) * WARPS + warp] = local;
		}
		// this warp is done with the stage: release it; the elected producer refills the stage of the PREVIOUS tile
		// (which every warp has left by now, so the wait below almost never blocks) with the tile two ahead
		__syncwarp();
		if (lane == 0) mbar_arrive(&empty[stage]);
		if (tid == 0 && it >= 1) {
			const int prev = it - 1, nxt = prev + SC_STAGES;
			if (nxt < n_tiles) {
				const int pstg = prev % SC_STAGES;
				mbar_wait(&empty[pstg], (uint32_t)((prev / SC_STAGES) & 1));
				mbar_expect_tx(&bars[pstg], tile_bytes);
				bulk_g2s(ps + (size_t)pstg * tileF, gt0 + (size_t)nxt * tileF, tile_bytes, &bars[pstg]);
			}
		}
		if (spk_done) {
			++spk;
			tile_in_spk = 0;
		}
	}
}

// sums[u][s] = sum over the frame tiles of utterance u (deterministic order); best[u] = first argmax_s
// (over the first S_arg models: a speaker set may carry its UBM as an extra, unranked model at the end)
constexpr int SCORE_REDUCE_THREADS = 1024;      // one block per utterance: the more threads share a speaker's tiles, the shorter the serial sums
__global__ void __launch_bounds__(SCORE_REDUCE_THREADS) score_reduce_kernel(const double *__restrict__ partial, const int *__restrict__ utt_tile_start, int S, int S_arg, int parts,
                                    double *__restrict__ sums, int *__restrict__ best) {
	const int u = blockIdx.x;
	const int t0 = utt_tile_start[u], t1 = utt_tile_start[u + 1];
	__shared__ double sv[SCORE_REDUCE_THREADS];
	__shared__ int si[SCORE_REDUCE_THREADS];
	double bestv = -1.0e300;
	int besti = 0x7fffffff;
	// tps threads per speaker (a power of two) split the utterance's tiles and are tree-summed: a long utterance scored against
	// few models (the E-step's pass A: one model, thousands of tiles) is not one thread's serial sum
	int tps = 1;
	while (tps * 2 * S <= (int)blockDim.x) tps *= 2;
	const int spk_per_pass = blockDim.x / tps, j = threadIdx.x % tps, sl = threadIdx.x / tps;
	for (int s0 = 0; s0 < S; s0 += spk_per_pass) {
		const int s = s0 + sl;
		double acc = 0.0;
		if (s < S)
			for (int t = t0 + j; t < t1; t += tps)
				for (int w = 0; w < parts; ++w) acc += partial[((size_t)t * S + s) * parts + w];
		if (tps > 1) {
			sv[threadIdx.x] = acc;
			__syncthreads();
			for (int w = tps / 2; w > 0; w >>= 1) {
				if (j < w) sv[threadIdx.x] += sv[threadIdx.x + w];
				__syncthreads();
			}
			acc = sv[threadIdx.x];
			__syncthreads();
		}
		if (j == 0 && s < S) {
			sums[(size_t)u * S + s] = acc;
			if (s < S_arg && acc > bestv) {   // strict: keeps the first maximum within a thread's ascending s sequence
				bestv = acc;
				besti = s;
			}
		}
	}
	if (best == nullptr) return;
	sv[threadIdx.x] = bestv;
	si[threadIdx.x] = besti;
	__syncthreads();
	for (int w = blockDim.x / 2; w > 0; w >>= 1) {
		if ((int)threadIdx.x < w) {
			const double ov = sv[threadIdx.x + w];
			const int oi = si[threadIdx.x + w];
			if (ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) {
				sv[threadIdx.x] = ov;
				si[threadIdx.x] = oi;
			}
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) best[u] = si[0] == 0x7fffffff ? 0 : si[0];
}

// ------------------------------------------------------------------------------------------------
// host side: frame-tile tables (cached per shape), chunking, launch
// ------------------------------------------------------------------------------------------------
struct TileTable {                   // one cached shape; slots are REUSED on eviction and their buffers only grow
	std::vector<int> key;            // utt_off + S + FT
	int n_ftiles = 0;
	unsigned char *d_meta = nullptr, *h_meta = nullptr;      // ftiles int4[n_ftiles] | utt_tile_start int[n_utt + 1]: device copy and pinned image
	size_t meta_cap = 0, partial_cap = 0;
	cudaEvent_t copied = nullptr;    // the upload of h_meta has executed (the image may be rewritten)
	int4 *d_ftiles = nullptr;
	int *d_utt_tile_start = nullptr;
	double *d_partial = nullptr;
	uint64_t stamp = 0;
};
constexpr size_t MAX_TABLES = 256;   // batched enrolment scores every speaker's own frame range: one shape per speaker
static std::vector<TileTable> g_tables;
static uint64_t g_stamp = 0;

static void free_table(TileTable &t) {
	if (t.d_meta) cudaFree(t.d_meta);
	if (t.h_meta) cudaFreeHost(t.h_meta);
	if (t.d_partial) cudaFree(t.d_partial);
	if (t.copied) cudaEventDestroy(t.copied);
	t = TileTable();
}

void score_tables_release() {
	for (auto &t : g_tables) free_table(t);
	g_tables.clear();
}

// Frame-tile table of a batch shape WITHOUT steady-state allocations: a miss rewrites the least recently used slot in place
// (pinned image -> one async copy on the caller's stream, no synchronisation); buffers are reallocated only to grow.
static int get_table(const int *utt_off, int n_utt, int S, int FT, cudaStream_t st, TileTable **out) {
	std::vector<int> key(utt_off, utt_off + n_utt + 1);
	key.push_back(S);
	key.push_back(FT);
	for (auto &t : g_tables)
		if (t.key == key) {
			t.stamp = ++g_stamp;
			*out = &t;
			return 0;
		}
	TileTable *t = nullptr;
	if (g_tables.size() < MAX_TABLES) {
		g_tables.reserve(MAX_TABLES);            // pointers handed out earlier stay valid
		g_tables.emplace_back();
		t = &g_tables.back();
		SRB_CUDA(cudaEventCreateWithFlags(&t->copied, cudaEventDisableTiming));
	} else {
		t = &g_tables[0];
		for (auto &c : g_tables)
			if (c.stamp < t->stamp) t = &c;
		SRB_CUDA(cudaEventSynchronize(t->copied));
	}
	long long nft = 0;
	for (int u = 0; u < n_utt; ++u) nft += (utt_off[u + 1] - utt_off[u] + FT - 1) / FT;
	const size_t off_uts = (size_t)std::max<long long>(nft, 1) * sizeof(int4), bytes = off_uts + (size_t)(n_utt + 1) * sizeof(int);
	const size_t pbytes = (size_t)std::max<long long>(nft, 1) * (size_t)S * SC_MAX_WARPS * sizeof(double);
	if (bytes > t->meta_cap) {
		if (t->d_meta) {
			SRB_CUDA(cudaStreamSynchronize(st));
			cudaFree(t->d_meta);
			cudaFreeHost(t->h_meta);
			t->d_meta = t->h_meta = nullptr;
		}
		const size_t want = bytes + bytes / 2 + 1024;
		SRB_CUDA(cudaMalloc(&t->d_meta, want));
		SRB_CUDA(cudaMallocHost(&t->h_meta, want));
		t->meta_cap = want;
	}
	if (pbytes > t->partial_cap) {
		if (t->d_partial) {
			SRB_CUDA(cudaStreamSynchronize(st));
			cudaFree(t->d_partial);
			t->d_partial = nullptr;
		}
		const size_t want = pbytes + pbytes / 2 + 1024;
		SRB_CUDA(cudaMalloc(&t->d_partial, want));
		t->partial_cap = want;
	}
	int4 *ft = reinterpret_cast<int4 *>(t->h_meta);
	int *uts = reinterpret_cast<int *>(t->h_meta + off_uts);
	int n = 0;
	for (int u = 0; u < n_utt; ++u) {
		uts[u] = n;
		const int b = utt_off[u], e = utt_off[u + 1];
		for (int f = b; f < e; f += FT) ft[n++] = make_int4(f, (e - f < FT ? e - f : FT), u, 0);
	}
	uts[n_utt] = n;
	t->key = key;
	t->n_ftiles = n;
	t->stamp = ++g_stamp;
	t->d_ftiles = reinterpret_cast<int4 *>(t->d_meta);
	t->d_utt_tile_start = reinterpret_cast<int *>(t->d_meta + off_uts);
	SRB_CUDA(cudaMemcpyAsync(t->d_meta, t->h_meta, bytes, cudaMemcpyHostToDevice, st));
	SRB_CUDA(cudaEventRecord(t->copied, st));
	*out = t;
	return 0;
}

// choose how many speaker chunks to cut the S axis into so that (frame tiles x chunks) fills whole waves
static int choose_spk_per_chunk(int n_ftiles, int S, int ctas_per_wave) {
	double best_eff = -1.0;
	int best_spc = S;
	for (int nchunks = 1; nchunks <= S; ++nchunks) {
		const int spc = ceil_div(S, nchunks);
		const int real_chunks = ceil_div(S, spc);
		if (real_chunks != nchunks) continue;
		const long long ctas = (long long)n_ftiles * real_chunks;
		const long long waves = ceil_div_ll(ctas, ctas_per_wave);
		// useful work / capacity; the x-tile load is amortised over spc speakers (small penalty for tiny chunks)
		double eff = ((double)n_ftiles * S) / ((double)waves * ctas_per_wave * spc);
		eff *= (double)spc / ((double)spc + 0.05);
		if (eff > best_eff + 1e-9) {
			best_eff = eff;
			best_spc = spc;
		}
	}
	return best_spc;
}

static size_t score_smem_bytes(int D, int FT) {
	return (size_t)D * FT * 4 + (size_t)SC_STAGES * tile_floats(D) * 4 + 2 * SC_STAGES * 8;
}

static bool g_score_attr_set = false;
static bool g_score_scalar = false;
static int g_score_warps = 4;
static bool g_score_stagger = false;   // measured: no effect (fair-share scheduling pulls co-resident warps back into lockstep)

int launch_score_tc_images(const Tc2Images &im, const float *dX, int ld, int D, const int *utt_off, int n_utt, double *d_sums, float *d_frame_ll,
                           int ldf, int *d_best, int clamp, int frame_log2, cudaStream_t st) {
	Context &c = ctx();
	c.last_score_route = 1;
	++c.score_launches;
	TileTable *tt = nullptr;
	if (get_table(utt_off, n_utt, im.S, 128, st, &tt)) return 1;
	if (launch_score_tc2_images(im, dX, ld, D, tt->d_ftiles, tt->n_ftiles, tt->d_partial, d_frame_ll, ldf, clamp, frame_log2, st)) return 1;
	score_reduce_kernel<<<n_utt, SCORE_REDUCE_THREADS, 0, st>>>(tt->d_partial, tt->d_utt_tile_start, im.S, im.S, T2_PARTS, d_sums, d_best);
	SRB_CUDA(cudaGetLastError());
	return 0;
}

int launch_score(PackedModels &pm, const float *dX, int ld, int D, const int *utt_off, int n_utt, double *d_sums,
                 float *d_frame_ll, int ldf, int *d_best, int clamp, int frame_log2, cudaStream_t st) {
	SRB_REQUIRE(pm.S > 0 && pm.d_tiles != nullptr, "score: empty model set");
	SRB_REQUIRE(D == pm.D, "score: feature dim %d != model dim %d", D, pm.D);
	SRB_REQUIRE(D >= 1 && D <= MAX_DIM, "score: dim %d outside [1,%d]", D, MAX_DIM);
	SRB_REQUIRE(ld >= D, "score: ld %d < dim %d", ld, D);
	SRB_REQUIRE(n_utt >= 1, "score: no utterances");
	for (int u = 0; u < n_utt; ++u) SRB_REQUIRE(utt_off[u + 1] >= utt_off[u], "score: utterance offsets not ascending");
	Context &c = ctx();
	if (!g_score_attr_set) {
		SRB_CUDA(cudaFuncSetAttribute(gmm_score_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		SRB_CUDA(cudaFuncSetAttribute(gmm_score_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		SRB_CUDA(cudaFuncSetAttribute(gmm_score_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		const char *ev = getenv("SRB200_SCORE_SCALAR");     // A/B switches for profiling
		g_score_scalar = ev && ev[0] == '1';
		const char *ew = getenv("SRB200_SCORE_WARPS");
		if (ew && ew[0] == '8') g_score_warps = 8;
		if (ew && ew[0] == '4') g_score_warps = 4;
		const char *es = getenv("SRB200_SCORE_STAGGER");
		if (es && es[0] == '1') g_score_stagger = true;
		g_score_attr_set = true;
	}
	// ---- arithmetic route: tcgen05 3xTF32 (default where the model set supports it) or fp32 FMA (per-frame log2 output for
	// the EM E-step, D > 55, too many tiny-sigma entries, or on request) ----
	int path = c.score_path;
	if (path < 0) {
		const char *et = getenv("SRB200_SCORE_TC");
		path = (et && et[0] == '0') ? 0 : 1;
	}
	c.last_score_route = 0;
	++c.score_launches;
	if (path == 1 && !frame_log2 && tc2_supported(pm, D)) {
		c.last_score_route = 1;
		TileTable *tt = nullptr;
		if (get_table(utt_off, n_utt, pm.S, 128, st, &tt)) return 1;
		if (launch_score_tc2(pm, dX, ld, D, tt->d_ftiles, tt->n_ftiles, tt->d_partial, d_frame_ll, ldf, clamp, frame_log2, st)) return 1;
		score_reduce_kernel<<<n_utt, SCORE_REDUCE_THREADS, 0, st>>>(tt->d_partial, tt->d_utt_tile_start, pm.S, pm.S_arg > 0 ? pm.S_arg : pm.S, T2_PARTS, d_sums, d_best);
		SRB_CUDA(cudaGetLastError());
		return 0;
	}
	int warps = g_score_scalar ? 4 : g_score_warps;
	if (score_smem_bytes(D, warps * 64) > 113 * 1024) warps = 4;   // 8-warp CTAs need two per SM
	const int FT = warps * 64;
	TileTable *tt = nullptr;
	if (get_table(utt_off, n_utt, pm.S, FT, st, &tt)) return 1;
	const size_t smem = score_smem_bytes(D, FT);
	if (tt->n_ftiles > 0) {
		int per_sm = (int)((227 * 1024) / (smem + 1024));
		if (per_sm > (warps == 4 ? 3 : 2)) per_sm = (warps == 4 ? 3 : 2);
		if (per_sm < 1) per_sm = 1;
		const int spc = choose_spk_per_chunk(tt->n_ftiles, pm.S, c.sm_count * per_sm);
		ScoreArgs a;
		a.X = dX;
		a.ld = ld;
		a.D = D;
		a.tiles = pm.d_tiles;
		a.tiles_per_model = pm.tiles_per_model;
		a.S = pm.S;
		a.spk_per_chunk = spc;
		a.ftiles = tt->d_ftiles;
		a.frame_ll = d_frame_ll;
		a.ldf = ldf;
		a.partial = tt->d_partial;
		a.clamp = clamp;
		a.frame_log2 = frame_log2;
		a.stagger_cycles = g_score_stagger ? 140 * D : 0;     // ~ one third of a tile period (3 warps x D x 64 FFMA2 x 2 cycles)
		a.n_sm = c.sm_count;
		a.x_vec_ok = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(dX) & 15) == 0);
		a.n_chunks = ceil_div(pm.S, spc);
		SRB_REQUIRE((long long)a.n_chunks * tt->n_ftiles < (1LL << 31), "score: grid too large");
		dim3 grid((unsigned)(a.n_chunks * tt->n_ftiles));
		if (g_score_scalar) gmm_score_kernel<false, 4><<<grid, 128, smem, st>>>(a);
		else if (warps == 8) gmm_score_kernel<true, 8><<<grid, 256, smem, st>>>(a);
		else gmm_score_kernel<true, 4><<<grid, 128, smem, st>>>(a);
		SRB_CUDA(cudaGetLastError());
	}
	score_reduce_kernel<<<n_utt, SCORE_REDUCE_THREADS, 0, st>>>(tt->d_partial, tt->d_utt_tile_start, pm.S, pm.S_arg > 0 ? pm.S_arg : pm.S, warps, d_sums, d_best);
	SRB_CUDA(cudaGetLastError());
	return 0;
}

}  // namespace srb
