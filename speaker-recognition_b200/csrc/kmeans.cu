// kmeans.cu -- K-MeansII (k-means||) initialisation of the GMM means, the reference's `init_with_kmeans = 1` path:
//   GMMTrainerBaseline::init_gaussians      src/gmm/src/gmm.cc:331-344
//   KMeansIISolver::cluster                 src/gmm/src/kmeansII.cc:80-172   oversampling rounds -> candidate set -> weights
//   KMeansppSolver::cluster_weighted        src/gmm/src/kmeans++.cc:157-213  weighted k-means++ seeding over the candidates
//   KMeansSolver::Lloyds_iteration(_weighted)  src/gmm/src/kmeans.cc:150-246, 249-340
//
// Every O(points x centroids x dim) step runs here in float64 (the reference's real_t), one thread per point, dimensions
// summed in index order and centroids visited in index order with a strict `<`, exactly like distancesqr / calc_belonging
// (kmeans.cc:62-107), so that assignments agree with a CPU restatement bit for bit.  Sums over points (costs, centroid
// accumulators) are two-level reductions in a fixed order: deterministic, but not the reference's thread-block order.
// The reference draws from libc rand() and std::default_random_engine; here every draw is a counter-based hash
// u(seed, stream, index) (splitmix64 finaliser), so runs are reproducible from `seed` and the test-side checker can restate
// the same draws.  Deliberate deviation: a cluster that loses all its points keeps its previous centroid (the
// reference divides 0/0 and carries a NaN centroid from then on).
#include <float.h>
#include <math.h>
#include <algorithm>
#include "common.cuh"
#include "kernels.h"

namespace srb {

static __host__ __device__ inline double km_uniform(uint64_t seed, uint32_t stream, uint64_t idx) {
	uint64_t z = seed + 0x9E3779B97F4A7C15ull * ((((uint64_t)stream) << 40) + idx + 1);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z ^= z >> 31;
	return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// squared distance in the reference's order (kmeans.cc:62-69): no FMA contraction, so that a CPU restatement agrees exactly
__device__ __forceinline__ double km_dist(const float *__restrict__ x, const double *__restrict__ c, int D) {
	double s = 0.0;
	for (int d = 0; d < D; ++d) {
		const double t = __dsub_rn((double)x[d], c[d]);
		s = __dadd_rn(s, __dmul_rn(t, t));
	}
	return s;
}

// dist[i] = min(dist[i], w_i * |x_i - c_j|^2) over j in [c0, c1), strict <, belong updated (kmeansII.cc:52-75, kmeans++.cc:60-96)
__global__ void km_min_dist_kernel(const float *__restrict__ X, int n, int ld, int D, const double *__restrict__ cent, int c0, int c1,
                                   const double *__restrict__ w, double *__restrict__ dist, int *__restrict__ belong) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float *x = X + (size_t)i * ld;
	double best = dist[i];
	int bj = belong ? belong[i] : 0;
	const double wi = w ? w[i] : 1.0;
	for (int j = c0; j < c1; ++j) {
		double dj = km_dist(x, cent + (size_t)j * D, D);
		if (w) dj = __dmul_rn(dj, wi);
		if (dj < best) {
			best = dj;
			bj = j;
		}
	}
	dist[i] = best;
	if (belong) belong[i] = bj;
}

// Lloyd assignment (calc_belonging, kmeans.cc:72-107 / 109-147): nearest of all K centroids from DBL_MAX, strict <
__global__ void km_assign_kernel(const float *__restrict__ X, int n, int ld, int D, const double *__restrict__ cent, int K,
                                 const double *__restrict__ w, double *__restrict__ dist, int *__restrict__ belong) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float *x = X + (size_t)i * ld;
	double best = DBL_MAX;
	int bj = 0;
	const double wi = w ? w[i] : 1.0;
	for (int j = 0; j < K; ++j) {
		double dj = km_dist(x, cent + (size_t)j * D, D);
		if (w) dj = __dmul_rn(dj, wi);
		if (dj < best) {
			best = dj;
			bj = j;
		}
	}
	dist[i] = best;
	belong[i] = bj;
}

// Tiled form of the two kernels above for D <= DT: the point lives in registers as float64 (the per-use float -> double
// conversion of the simple kernels runs on the quarter-rate conversion path and dominated them), centroids are staged 32 at a
// time in shared memory and read as broadcasts.  Same operations in the same order -> bit-identical results.
// fresh != 0: Lloyd assignment from DBL_MAX (belong always written); fresh == 0: running minimum over centroids [c0, c1).
template <int DT>
__global__ void __launch_bounds__(128) km_nearest_kernel(const float *__restrict__ X, int n, int ld, int D, const double *__restrict__ cent, int c0,
                                                        int c1, const double *__restrict__ w, double *__restrict__ dist, int *__restrict__ belong,
                                                        int fresh) {
	constexpr int TJ = 32;
	__shared__ double sc[TJ * DT];
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const bool live = i < n;
	double xr[DT];
#pragma unroll
	for (int d = 0; d < DT; ++d) xr[d] = (live && d < D) ? (double)X[(size_t)i * ld + d] : 0.0;
	double best = fresh ? DBL_MAX : (live ? dist[i] : 0.0);
	int bj = (fresh || belong == nullptr || !live) ? 0 : belong[i];
	const double wi = (w && live) ? w[i] : 1.0;
	for (int j0 = c0; j0 < c1; j0 += TJ) {
		const int nj = min(TJ, c1 - j0);
		__syncthreads();
		for (int t = threadIdx.x; t < nj * D; t += blockDim.x) {
			const int j = t / D, d = t - j * D;
			sc[j * DT + d] = cent[(size_t)(j0 + j) * D + d];
		}
		__syncthreads();
		for (int j = 0; j < nj; ++j) {
			const double *c = sc + j * DT;
			double sacc = 0.0;
#pragma unroll
			for (int d = 0; d < DT; ++d) {
				if (d < D) {
					const double t = __dsub_rn(xr[d], c[d]);
					sacc = __dadd_rn(sacc, __dmul_rn(t, t));
				}
			}
			if (w) sacc = __dmul_rn(sacc, wi);
			if (sacc < best) {
				best = sacc;
				bj = j0 + j;
			}
		}
	}
	if (live) {
		dist[i] = best;
		if (belong) belong[i] = bj;
	}
}

static int launch_km_nearest(const float *X, int n, int ld, int D, const double *cent, int c0, int c1, const double *w, double *dist, int *belong,
                             int fresh, cudaStream_t st) {
	const int grid = (n + 127) / 128;
	if (D <= 16) km_nearest_kernel<16><<<grid, 128, 0, st>>>(X, n, ld, D, cent, c0, c1, w, dist, belong, fresh);
	else if (D <= 28) km_nearest_kernel<28><<<grid, 128, 0, st>>>(X, n, ld, D, cent, c0, c1, w, dist, belong, fresh);
	else if (D <= 40) km_nearest_kernel<40><<<grid, 128, 0, st>>>(X, n, ld, D, cent, c0, c1, w, dist, belong, fresh);
	else if (fresh) km_assign_kernel<<<grid, 128, 0, st>>>(X, n, ld, D, cent, c1, w, dist, belong);
	else km_min_dist_kernel<<<grid, 128, 0, st>>>(X, n, ld, D, cent, c0, c1, w, dist, belong);
	SRB_CUDA(cudaGetLastError());
	return 0;
}

// deterministic sum of n doubles: fixed chunk per CTA, fixed tree inside; second launch (1 CTA) sums the partials in order
__global__ void km_sum_kernel(const double *__restrict__ v, long long n, double *__restrict__ out) {
	__shared__ double sh[256];
	const long long per = (n + gridDim.x - 1) / gridDim.x;
	const long long b = per * blockIdx.x, e = min(n, b + per);
	double s = 0.0;
	for (long long i = b + threadIdx.x; i < e; i += 256) s += v[i];
	sh[threadIdx.x] = s;
	__syncthreads();
	for (int o = 128; o >= 1; o >>= 1) {
		if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
		__syncthreads();
	}
	if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

// oversampling draw of one round (kmeansII.cc:135-141): sel[i] = u_i * S < dist[i] * l * K
__global__ void km_sample_kernel(const double *__restrict__ dist, int n, double S, double lK, uint64_t seed, uint32_t stream,
                                 unsigned char *__restrict__ sel) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	sel[i] = (km_uniform(seed, stream, (uint64_t)i) * S < dist[i] * lK) ? 1 : 0;
}

__global__ void km_gather_kernel(const float *__restrict__ X, int ld, int D, const int *__restrict__ idx, int m, float *__restrict__ out, int ldo) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= m * D) return;
	const int r = t / D, d = t - r * D;
	out[(size_t)r * ldo + d] = X[(size_t)idx[r] * ld + d];
}

__global__ void km_rows_to_double_kernel(const float *__restrict__ X, int ld, int D, const int *__restrict__ idx, int m, double *__restrict__ out) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= m * D) return;
	const int r = t / D, d = t - r * D;
	out[(size_t)r * D + d] = (double)X[(size_t)idx[r] * ld + d];
}

__global__ void km_count_kernel(const int *__restrict__ belong, int n, int m, double *__restrict__ weight) {
	// weight[j] = #points with belong == j (kmeansII.cc:162-164); integer-valued doubles: atomic order cannot change the result
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) atomicAdd(&weight[belong[i]], 1.0);
}

// centroid accumulators of one CTA's contiguous chunk of points, cells (k - k0) * (D + 1) + d for k in [k0, k1): sum of w x_d,
// and in column D the sum of w.  Cell c is owned by thread c % 1024 for the whole kernel and points are visited in index
// order, so every cell is a sequential sum in point order: deterministic without atomics.
__global__ void __launch_bounds__(1024) km_accumulate_kernel(const float *__restrict__ X, int n, int ld, int D, const int *__restrict__ belong,
                                                            const double *__restrict__ w, int k0, int k1, double *__restrict__ partial) {
	extern __shared__ double cells[];
	const int DP = D + 1, ncell = (k1 - k0) * DP;
	for (int c = threadIdx.x; c < ncell; c += 1024) cells[c] = 0.0;
	__syncthreads();
	const int per = (n + gridDim.x - 1) / gridDim.x;
	const int b = per * blockIdx.x, e = min(n, b + per);
	const int t = threadIdx.x;
	for (int i = b; i < e; ++i) {
		const int k = belong[i];
		if (k < k0 || k >= k1) continue;              // uniform across the CTA
		const int base = (k - k0) * DP;
		const int off = (t - base) & 1023;            // the unique cell of this cluster's row owned by thread t, if any
		if (off < DP) {
			const double wi = w ? w[i] : 1.0;
			const double v = off < D ? (double)X[(size_t)i * ld + off] * wi : wi;
			cells[base + off] += v;
		}
	}
	__syncthreads();
	double *out = partial + (size_t)blockIdx.x * ncell;
	for (int c = threadIdx.x; c < ncell; c += 1024) out[c] = cells[c];
}

// new centroid = sum(w x) / sum(w) from the CTA partials summed in CTA order (kmeans.cc:206-232); empty cluster keeps its centroid
__global__ void km_update_kernel(const double *__restrict__ partial, int n_cta, int D, int k0, int k1, double *__restrict__ cent) {
	const int DP = D + 1, ncell = (k1 - k0) * DP;
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (k1 - k0) * D) return;
	const int k = t / D, d = t - k * D;
	double s = 0.0, cnt = 0.0;
	for (int c = 0; c < n_cta; ++c) {
		s += partial[(size_t)c * ncell + k * DP + d];
		cnt += partial[(size_t)c * ncell + k * DP + D];
	}
	if (cnt > 0.0) cent[(size_t)(k0 + k) * D + d] = s / cnt;
}

namespace {

struct KmWork {
	DevBuf dist, belong, sel, partial, sums, cent, cent_best, cand, cand_d, idx, weight, wdist, one;
	~KmWork() {
		dist.release();
		belong.release();
		sel.release();
		partial.release();
		sums.release();
		cent.release();
		cent_best.release();
		cand.release();
		cand_d.release();
		idx.release();
		weight.release();
		wdist.release();
		one.release();
	}
};

int km_sum(const double *d_v, long long n, KmWork &W, cudaStream_t st, double *out) {
	const int ctas = 128;
	if (W.sums.ensure((ctas + 1) * sizeof(double))) return 1;
	km_sum_kernel<<<ctas, 256, 0, st>>>(d_v, n, W.sums.as<double>());
	km_sum_kernel<<<1, 256, 0, st>>>(W.sums.as<double>(), ctas, W.sums.as<double>() + ctas);
	SRB_CUDA(cudaGetLastError());
	SRB_CUDA(cudaMemcpyAsync(out, W.sums.as<double>() + ctas, sizeof(double), cudaMemcpyDeviceToHost, st));
	SRB_CUDA(cudaStreamSynchronize(st));
	return 0;
}

// Lloyd iterations (kmeans.cc:150-246 unweighted, 249-340 weighted) on device data; d_cent [K][D] in/out (best centroids on return)
int km_lloyd(const float *dX, int n, int ld, int D, int K, const double *d_w, KmWork &W, double *d_cent, cudaStream_t st, double *best_out,
             int *iters_out, int verbosity) {
	Context &c = ctx();
	const int DP = D + 1;
	const int max_cells = (200 * 1024) / (int)sizeof(double);
	const int kpass = std::max(1, std::min(K, max_cells / DP));
	const int n_cta = std::max(1, std::min(c.sm_count, (n + 255) / 256));
	if (W.partial.ensure((size_t)n_cta * kpass * DP * sizeof(double))) return 1;
	if (W.cent_best.ensure((size_t)K * D * sizeof(double))) return 1;
	static bool attr = false;
	if (!attr) {
		SRB_CUDA(cudaFuncSetAttribute(km_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
		attr = true;
	}
	double best = DBL_MAX, last = DBL_MAX;
	int it = 0;
	for (; it < 200; ++it) {
		if (launch_km_nearest(dX, n, ld, D, d_cent, 0, K, d_w, W.dist.as<double>(), W.belong.as<int>(), 1, st)) return 1;
		double cost = 0.0;
		if (km_sum(W.dist.as<double>(), n, W, st, &cost)) return 1;
		if (cost < best) {
			best = cost;
			SRB_CUDA(cudaMemcpyAsync(W.cent_best.p, d_cent, (size_t)K * D * sizeof(double), cudaMemcpyDeviceToDevice, st));
		}
		if (verbosity >= 2) printf("kmeans iteration %3d: %f\n", it, cost);
		if (fabs(last - cost) < 1e-6) break;                               // kmeans.cc:194 / 299
		if (d_w == nullptr && cost > best * 1.5) break;                    // kmeans.cc:199-203 (unweighted only)
		for (int k0 = 0; k0 < K; k0 += kpass) {
			const int k1 = std::min(K, k0 + kpass);
			km_accumulate_kernel<<<n_cta, 1024, (size_t)(k1 - k0) * DP * sizeof(double), st>>>(dX, n, ld, D, W.belong.as<int>(), d_w, k0, k1,
			                                                                                 W.partial.as<double>());
			km_update_kernel<<<((k1 - k0) * D + 127) / 128, 128, 0, st>>>(W.partial.as<double>(), n_cta, D, k0, k1, d_cent);
		}
		SRB_CUDA(cudaGetLastError());
		last = cost;
	}
	SRB_CUDA(cudaMemcpyAsync(d_cent, W.cent_best.p, (size_t)K * D * sizeof(double), cudaMemcpyDeviceToDevice, st));
	SRB_CUDA(cudaStreamSynchronize(st));
	if (best_out) *best_out = best;
	if (iters_out) *iters_out = it;
	return 0;
}

}  // namespace

// Lloyd iterations from given centroids (float64 [K][D], in/out) on staged device features; w: optional host weights [n]
int kmeans_lloyd_device(const float *dX, int n, int ld, int D, int K, const double *h_w, double *h_cent, double *cost_out, int *iters_out) {
	Context &c = ctx();
	cudaStream_t st = c.stream;
	KmWork W;
	if (W.dist.ensure((size_t)n * sizeof(double)) || W.belong.ensure((size_t)n * sizeof(int)) || W.cent.ensure((size_t)K * D * sizeof(double))) return 1;
	const double *d_w = nullptr;
	if (h_w) {
		if (W.weight.ensure((size_t)n * sizeof(double))) return 1;
		SRB_CUDA(cudaMemcpyAsync(W.weight.p, h_w, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
		d_w = W.weight.as<double>();
	}
	SRB_CUDA(cudaMemcpyAsync(W.cent.p, h_cent, (size_t)K * D * sizeof(double), cudaMemcpyHostToDevice, st));
	if (km_lloyd(dX, n, ld, D, K, d_w, W, W.cent.as<double>(), st, cost_out, iters_out, 0)) return 1;
	SRB_CUDA(cudaMemcpy(h_cent, W.cent.p, (size_t)K * D * sizeof(double), cudaMemcpyDeviceToHost));
	return 0;
}

// K-MeansII on staged device features dX [n][ld] (fp32): centroids_out [K][D] float64.  Streams of draws: 0 first point,
// 1 + r oversampling round r, 100 fill, 200 weighted k-means++ (index 0 = first candidate, k = k-th seeding step).
int kmeans2_init_device(const float *dX, int n, int ld, int D, int K, uint64_t seed, double *centroids_out, double *cost_out, int verbosity) {
	SRB_REQUIRE(n >= 1 && K >= 1 && D >= 1, "kmeans: bad shape n=%d K=%d D=%d", n, K, D);
	Context &c = ctx();
	cudaStream_t st = c.stream;
	KmWork W;
	const double l = 2.0, size_factor = 2.0;                 // KMeansIISolver defaults, kmeansII.hh:29
	if (W.dist.ensure((size_t)n * sizeof(double)) || W.belong.ensure((size_t)n * sizeof(int)) || W.sel.ensure((size_t)n)) return 1;
	std::vector<int> cand;                                    // candidate = index of a data row
	cand.push_back(std::min(n - 1, (int)(km_uniform(seed, 0, 0) * n)));          // kmeansII.cc:97
	{
		std::vector<double> init((size_t)n, DBL_MAX);
		SRB_CUDA(cudaMemcpyAsync(W.dist.p, init.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
		SRB_CUDA(cudaMemsetAsync(W.belong.p, 0, (size_t)n * sizeof(int), st));
		SRB_CUDA(cudaStreamSynchronize(st));
	}
	std::vector<unsigned char> h_sel((size_t)n);
	int last = 0;
	auto upload_cand_rows = [&](int from, int to) -> int {   // candidate rows [from, to) as float64 into W.cand_d (whole array re-ensured)
		if (W.idx.ensure((size_t)cand.size() * sizeof(int))) return 1;
		SRB_CUDA(cudaMemcpyAsync(W.idx.p, cand.data(), cand.size() * sizeof(int), cudaMemcpyHostToDevice, st));
		if (W.cand_d.ensure((size_t)cand.size() * D * sizeof(double))) return 1;
		const int m = (int)cand.size();
		km_rows_to_double_kernel<<<(m * D + 255) / 256, 256, 0, st>>>(dX, ld, D, W.idx.as<int>(), m, W.cand_d.as<double>());
		SRB_CUDA(cudaGetLastError());
		(void)from;
		(void)to;
		return 0;
	};
	for (int round = 0;; ++round) {
		if (upload_cand_rows(last, (int)cand.size())) return 1;
		if (launch_km_nearest(dX, n, ld, D, W.cand_d.as<double>(), last, (int)cand.size(), nullptr, W.dist.as<double>(), W.belong.as<int>(), 0, st)) return 1;
		if ((double)cand.size() > size_factor * K) break;                          // kmeansII.cc:122
		double S = 0.0;
		if (km_sum(W.dist.as<double>(), n, W, st, &S)) return 1;
		last = (int)cand.size();
		km_sample_kernel<<<(n + 255) / 256, 256, 0, st>>>(W.dist.as<double>(), n, S, l * K, seed, 1u + (uint32_t)round, W.sel.as<unsigned char>());
		SRB_CUDA(cudaGetLastError());
		SRB_CUDA(cudaMemcpyAsync(h_sel.data(), W.sel.p, (size_t)n, cudaMemcpyDeviceToHost, st));
		SRB_CUDA(cudaStreamSynchronize(st));
		for (int i = 0; i < n; ++i)
			if (h_sel[i]) cand.push_back(i);
		if (verbosity >= 2) printf("kmeansII iteration %3d: %f, #new: %d, #all: %d\n", round, S, (int)cand.size() - last, (int)cand.size());
		if ((int)cand.size() == last) break;                                       // kmeansII.cc:150-151
		SRB_REQUIRE(round < 1000, "kmeansII: sampling did not terminate");
	}
	for (uint64_t k = 0; (double)cand.size() <= size_factor * K; ++k)            // kmeansII.cc:157-160 (no distance update for these)
		cand.push_back(std::min(n - 1, (int)(km_uniform(seed, 100, k) * n)));
	const int m = (int)cand.size();
	// weights = cluster sizes of the candidates (kmeansII.cc:162-164)
	if (W.weight.ensure((size_t)m * sizeof(double))) return 1;
	SRB_CUDA(cudaMemsetAsync(W.weight.p, 0, (size_t)m * sizeof(double), st));
	km_count_kernel<<<(n + 255) / 256, 256, 0, st>>>(W.belong.as<int>(), n, m, W.weight.as<double>());
	SRB_CUDA(cudaGetLastError());
	// candidate rows as an fp32 matrix [m][D] (the "dataset" of the weighted stage)
	if (W.idx.ensure((size_t)m * sizeof(int)) || W.cand.ensure((size_t)m * D * sizeof(float))) return 1;
	SRB_CUDA(cudaMemcpyAsync(W.idx.p, cand.data(), (size_t)m * sizeof(int), cudaMemcpyHostToDevice, st));
	km_gather_kernel<<<(m * D + 255) / 256, 256, 0, st>>>(dX, ld, D, W.idx.as<int>(), m, W.cand.as<float>(), D);
	SRB_CUDA(cudaGetLastError());
	// ---- weighted k-means++ over the candidates (kmeans++.cc:157-213) ----
	if (W.cent.ensure((size_t)K * D * sizeof(double)) || W.wdist.ensure((size_t)m * sizeof(double))) return 1;
	std::vector<double> h_wd((size_t)m, DBL_MAX);
	SRB_CUDA(cudaMemcpyAsync(W.wdist.p, h_wd.data(), (size_t)m * sizeof(double), cudaMemcpyHostToDevice, st));
	SRB_CUDA(cudaMemsetAsync(W.cent.p, 0, (size_t)K * D * sizeof(double), st));
	std::vector<int> pick(1, std::min(m - 1, (int)(km_uniform(seed, 200, 0) * m)));
	auto set_centroid = [&](int k, int cand_row) -> int {
		int h = cand_row;
		if (W.one.ensure(sizeof(int))) return 1;
		int *d_one = W.one.as<int>();
		SRB_CUDA(cudaMemcpyAsync(d_one, &h, sizeof(int), cudaMemcpyHostToDevice, st));
		km_rows_to_double_kernel<<<(D + 255) / 256, 256, 0, st>>>(W.cand.as<float>(), D, D, d_one, 1, W.cent.as<double>() + (size_t)k * D);
		SRB_CUDA(cudaGetLastError());
		SRB_CUDA(cudaStreamSynchronize(st));                    // h leaves scope
		return 0;
	};
	if (set_centroid(0, pick[0])) return 1;
	for (int k = 1; k < K; ++k) {
		if (launch_km_nearest(W.cand.as<float>(), m, D, D, W.cent.as<double>(), k - 1, k, W.weight.as<double>(), W.wdist.as<double>(), nullptr, 0, st)) return 1;
		SRB_CUDA(cudaMemcpyAsync(h_wd.data(), W.wdist.p, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
		SRB_CUDA(cudaStreamSynchronize(st));
		double S = 0.0;
		for (int i = 0; i < m; ++i) S += h_wd[i];              // m is a few thousand: summed in index order on the host
		double r = km_uniform(seed, 200, (uint64_t)k) * S;     // kmeans++.cc:198
		int chosen = -1, last_pos = -1;
		for (int i = 0; i < m; ++i) {
			if (h_wd[i] > 0.0) last_pos = i;
			r -= h_wd[i];
			if (r <= 0.0) {
				chosen = i;
				break;
			}
		}
		if (chosen < 0) chosen = last_pos >= 0 ? last_pos : 0; // rounding left r > 0 (the reference would keep an all-zero centroid)
		if (set_centroid(k, chosen)) return 1;
	}
	int iters = 0;
	double wcost = 0.0;
	// the weighted stage needs per-point buffers of size m: dist/belong have n >= ... entries only if m <= n
	KmWork Wc;
	if (Wc.dist.ensure((size_t)m * sizeof(double)) || Wc.belong.ensure((size_t)m * sizeof(int))) return 1;
	if (km_lloyd(W.cand.as<float>(), m, D, D, K, W.weight.as<double>(), Wc, W.cent.as<double>(), st, &wcost, &iters, verbosity)) return 1;
	// ---- Lloyd on the full data set (kmeansII.cc:171) ----
	double cost = 0.0;
	if (km_lloyd(dX, n, ld, D, K, nullptr, W, W.cent.as<double>(), st, &cost, &iters, verbosity)) return 1;
	SRB_CUDA(cudaMemcpy(centroids_out, W.cent.p, (size_t)K * D * sizeof(double), cudaMemcpyDeviceToHost));
	if (cost_out) *cost_out = cost;
	if (verbosity >= 1) printf("minimum distance squared sum(Lloyds_iteration): %f\n", cost);
	return 0;
}

}  // namespace srb
