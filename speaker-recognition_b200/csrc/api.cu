// api.cu -- the extern "C" surface of libsrb200.so (include/srb200.h): handle management, model text I/O,
// host<->device marshalling, the training driver.  All arithmetic of the hot path happens in the kernels of
// gmm_score.cu / gmm_em.cu / mfcc.cu; this file has no CPU compute path for scoring, EM statistics or MFCC.
#include <float.h>
#include <math.h>
#include <stdarg.h>
#include <algorithm>
#include "../../include/srb200.h"
#include "kernels.h"
#include <cuda_fp16.h>
#include "gmm_score_tc_layout.h"

namespace srb {

// ---------------------------------------------------------------------------------------------
// errors, context, buffers
// ---------------------------------------------------------------------------------------------
static thread_local std::string t_err;
void set_error(const char *fmt, ...) {
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	t_err = buf;
}
void clear_error() { t_err.clear(); }

int DevBuf::ensure(size_t bytes) {
	if (bytes <= cap && p) return 0;
	if (p) {
		cudaDeviceSynchronize();
		cudaFree(p);
		p = nullptr;
		cap = 0;
	}
	size_t want = bytes + bytes / 4 + 256;
	SRB_CUDA(cudaMalloc(&p, want));
	cap = want;
	return 0;
}
void DevBuf::release() {
	if (p) cudaFree(p);
	p = nullptr;
	cap = 0;
}
int PinBuf::ensure(size_t bytes) {
	if (bytes <= cap && p) return 0;
	if (p) {
		cudaDeviceSynchronize();
		cudaFreeHost(p);
		p = nullptr;
		cap = 0;
	}
	size_t want = bytes + bytes / 4 + 256;
	SRB_CUDA(cudaMallocHost(&p, want));
	cap = want;
	return 0;
}
void PinBuf::release() {
	if (p) cudaFreeHost(p);
	p = nullptr;
	cap = 0;
}

Context &ctx() {
	static Context c;
	return c;
}

int Context::init() {
	if (inited) return 0;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0) {
		set_error("libsrb200: no usable CUDA device (%s); this library has no CPU path", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
		return 1;
	}
	SRB_REQUIRE(device >= 0 && device < n, "libsrb200: device %d out of range (have %d)", device, n);
	SRB_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop;
	SRB_CUDA(cudaGetDeviceProperties(&prop, device));
	SRB_REQUIRE(prop.major >= 10, "libsrb200: device %d is sm_%d%d; this build contains sm_100a code only", device, prop.major, prop.minor);
	sm_count = prop.multiProcessorCount;
	SRB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
	inited = true;
	return 0;
}

void PackedModels::release_tc() {
	if (d_tc_hard) cudaFree(d_tc_hard);
	if (d_tc_cs) cudaFree(d_tc_cs);
	if (d_tc2) cudaFree(d_tc2);
	d_tc2 = nullptr;
	h_tc2.clear();
	d_tc_hard = nullptr;
	d_tc_cs = nullptr;
	h_tc_hard.clear();
	h_tc_cs.clear();
	tc_built = false;
}

void PackedModels::release() {
	if (d_tiles) cudaFree(d_tiles);
	d_tiles = nullptr;
	release_tc();
	bytes = 0;
	S = 0;
}

// Tensor-core weight images (gmm_score_tc2.cuh).  Features are transformed per dimension, x'' = (x - c_d) alpha_d, with c_d the
// weighted mean of the set's means and alpha_d the power of two nearest to 1 / (the set's overall standard deviation in d), so
// that the expanded square works on unit-scale, centred numbers whatever the feature's own scale is (MFCC after CMVN ~ 1, LPC
// coefficients ~ 0.05: the shipped 34-dim UBMs).  Row r of speaker s holds W = [-a''^2 (D), -2 a'' nb'' (D), C, 0-pad] with
// a'' = sqrt(log2e/2) / (sigma alpha), nb'' = -(mean - c) alpha a'', C = log2 w - sum log2(sqrt(2pi) sigma) - sum nb''^2; each
// value is split by round-to-nearest into an fp16 head and an fp16 tail.  The expanded form cancels where |nb''| is large (error
// ~ eps * nb''^2 per term, and the fp32 accumulator works at magnitude sum nb''^2): per row, the dimensions with the largest
// nb''^2 are taken out of W and C until the remaining sum is below TC_HARD_BUDGET, and are added back exactly by the epilogue
// (t = a'' x'' + nb'', l -= t^2).  Speakers are padded to 32 / 64 / multiples of 128 rows; inside every 32-column group the rows
// that carry such entries are placed FIRST (a log-sum-exp does not care about the order of its terms), so the epilogue's
// correction is a prefix of the group's columns.
void pack_models_tc_host(PackedModels &pm, int S, int K, int D, const double *const *weights, const double *const *means,
                         const double *const *sigmas, const int *Ks) {
	const int NB = TC_NB;
	const int cols = K <= 32 ? 32 : (K <= 64 ? 64 : ((K + NB - 1) / NB) * NB);
	const long long rows_total = (long long)S * cols;
	const int blocks = (int)((rows_total + NB - 1) / NB);
	pm.release_tc();
	pm.tc_built = true;
	const int Kp2 = t2_kp(D);
	pm.tc2_Kp = Kp2;
	pm.tc_cols_per_spk = cols;
	pm.tc_blocks_total = blocks;
	if (Kp2 > T2_KP_MAX) return;                   // 2 D + 1 > 127: PHI does not fit beside the accumulator slots -> FMA route
	const size_t bh = t2_block_halfs(Kp2);
	pm.h_tc2.assign((size_t)blocks * bh, 0);
	const double A0 = 0.84932180028801904272;
	double BUDGET = TC_HARD_BUDGET;
	if (const char *eh = getenv("SRB200_TC_HARD_BUDGET")) BUDGET = atof(eh);   // experiments: a huge value disables the add-back list
	// ---- the feature transform of this set: centre and power-of-two scale per dimension ----
	std::vector<double> cen(D, 0.0), spread(D, 0.0), alpha(D, 1.0);
	{
		double wsum = 0.0;
		for (int s = 0; s < S; ++s)
			for (int k = 0; k < Ks[s]; ++k) {
				const double w = weights[s][k] > 0.0 ? weights[s][k] : 0.0;
				wsum += w;
				for (int d = 0; d < D; ++d) cen[d] += w * means[s][(size_t)k * D + d];
			}
		if (!(wsum > 0.0)) wsum = 1.0;
		for (int d = 0; d < D; ++d) cen[d] /= wsum;
		for (int s = 0; s < S; ++s)
			for (int k = 0; k < Ks[s]; ++k) {
				const double w = weights[s][k] > 0.0 ? weights[s][k] : 0.0;
				for (int d = 0; d < D; ++d) {
					const double dm = means[s][(size_t)k * D + d] - cen[d], sg = sigmas[s][(size_t)k * D + d];
					spread[d] += w * (dm * dm + sg * sg);
				}
			}
		pm.h_tc_cs.assign((size_t)2 * D, 0.f);
		for (int d = 0; d < D; ++d) {
			const double sd = sqrt(spread[d] / wsum);
			int e = (sd > 0.0 && sd == sd) ? (int)lround(-log2(sd)) : 0;
			e = std::max(-60, std::min(60, e));
			alpha[d] = ldexp(1.0, e);
			cen[d] = (double)(float)cen[d];            // the kernel subtracts the fp32 value: use exactly that one here
			pm.h_tc_cs[d] = (float)cen[d];
			pm.h_tc_cs[D + d] = (float)alpha[d];
		}
	}
	// image writer: block blk, column col (0..127), k -> fp16 head + fp16 tail, both round-to-nearest (22 significant bits, unbiased);
	// the CTA of the pair that holds the row is col >> 6
	auto put = [&](int blk, int col, int k, double w) {
		const double wc = std::max(-T2_W_MAX, std::min(T2_W_MAX, w));
		const __half h = __float2half_rn((float)wc);
		const __half l = __float2half_rn((float)(wc - (double)__half2float(h)));
		const size_t base = (size_t)blk * bh + (size_t)(col >> 6) * 2 * T2_NH * Kp2 + t2_img_off(col & 63, k);
		pm.h_tc2[base] = __half_as_ushort(h);
		pm.h_tc2[base + (size_t)T2_NH * Kp2] = __half_as_ushort(l);
	};
	// hard-entry lists (layout: gmm_score_tc_layout.h)
	pm.h_tc_hard.assign((size_t)blocks * 4 * TC_HSTRIDE, 0);
	struct HardEntry { int x; float a, nb; int w; };
	struct RowPlan { bool real; double C; std::vector<double> av, nbv; std::vector<char> hard; int n_hard; };
	bool overflow = false;
	std::vector<int> order(D);
	const int seg_cols = std::min(cols, NB), n_seg = NB / seg_cols, gps = seg_cols / 32;      // a segment = one speaker's columns inside a block
	std::vector<RowPlan> rp(seg_cols);
	for (int blk = 0; blk < blocks && !overflow; ++blk) {
		int *hdr = pm.h_tc_hard.data() + (size_t)blk * TC_HSTRIDE * 4;
		std::vector<HardEntry> glist[4];
		for (int sg_i = 0; sg_i < n_seg && !overflow; ++sg_i) {
			for (int j = 0; j < seg_cols; ++j) {
				const long long row = (long long)blk * NB + sg_i * seg_cols + j;
				const int s = (int)(row / cols), k = (int)(row % cols);
				RowPlan &r = rp[j];
				r.real = s < S && k < Ks[s];
				r.n_hard = 0;
				if (!r.real) continue;
				const double *mu = means[s] + (size_t)k * D, *sg = sigmas[s] + (size_t)k * D;
				r.av.resize(D), r.nbv.resize(D), r.hard.assign(D, 0);
				double C = weights[s][k] > 0.0 ? log2(weights[s][k]) : -1.0e30, tot = 0.0;
				for (int d = 0; d < D; ++d) {
					r.av[d] = A0 / (sg[d] * alpha[d]);
					r.nbv[d] = -(mu[d] - cen[d]) * alpha[d] * r.av[d];
					C -= log2(2.5066282746310002 * sg[d]);
					order[d] = d;
					if (r.av[d] * r.av[d] > T2_W_MAX || fabs(2.0 * r.av[d] * r.nbv[d]) > T2_W_MAX) r.hard[d] = 1;   // outside the fp16 image's range
					else tot += r.nbv[d] * r.nbv[d];
				}
				if (tot > BUDGET) {
					std::sort(order.begin(), order.end(), [&](int x, int y) { return r.nbv[x] * r.nbv[x] > r.nbv[y] * r.nbv[y]; });
					for (int i = 0; i < D && tot > BUDGET; ++i) {
						if (r.hard[order[i]]) continue;
						r.hard[order[i]] = 1;
						tot -= r.nbv[order[i]] * r.nbv[order[i]];
					}
				}
				for (int d = 0; d < D; ++d) {
					if (r.hard[d]) ++r.n_hard;
					else C -= r.nbv[d] * r.nbv[d];
				}
				r.C = (C < -1.0e30 || C != C) ? -1.0e30 : C;
			}
			// Column order inside the segment (a log-sum-exp does not care): the rows that carry exact entries go FIRST into one
			// 32-column group -- the group rotates with the block index, so that the epilogue warp that owns it is a different one
			// from block to block -- up to T2_MAX_FLAGGED per group, spilling into the next group; plain rows fill up, padding last.
			std::vector<int> flagged, plain, pad;
			for (int j = 0; j < seg_cols; ++j) (!rp[j].real ? pad : (rp[j].n_hard > 0 ? flagged : plain)).push_back(j);
			if ((int)flagged.size() > T2_MAX_FLAGGED * gps) {
				overflow = true;
				break;
			}
			std::vector<int> members[4];
			int nflag_g[4] = {0, 0, 0, 0};
			size_t fi = 0;
			for (int t = 0; t < gps; ++t) {      // (spreading them evenly over the groups instead measured 5 % slower: the fixed cost per group counts)
				const int g = (t + blk) % gps;
				while (fi < flagged.size() && nflag_g[g] < T2_MAX_FLAGGED) {
					members[g].push_back(flagged[fi++]);
					++nflag_g[g];
				}
			}
			size_t pi = 0, di = 0;
			for (int g = 0; g < gps; ++g)
				while ((int)members[g].size() < 32) members[g].push_back(pi < plain.size() ? plain[pi++] : pad[di++]);
			for (int g = 0; g < gps; ++g) {
				const int gi = sg_i * gps + g;                  // group index inside the block
				std::vector<HardEntry> &gl = glist[gi];
				for (int p = 0; p < 32; ++p) {
					const RowPlan &r = rp[members[g][p]];
					const int col = gi * 32 + p;
					if (!r.real) {
						put(blk, col, 2 * D, -T2_W_MAX);        // padded row; its column is masked out by the valid mask anyway
						continue;
					}
					hdr[gi * 4 + 3] |= (int)(1u << p);          // header word gi, .w: valid columns
					for (int d = 0; d < D; ++d) {
						if (r.hard[d]) {
							gl.push_back({4 * d, (float)r.av[d], (float)r.nbv[d], 0});
						} else {
							put(blk, col, d, -r.av[d] * r.av[d]);
							put(blk, col, D + d, -2.0 * r.av[d] * r.nbv[d]);
						}
					}
					if (r.n_hard > 0) gl.back().w = 1 + 4 * p;   // last entry of this column: flush the sum to scratch slot p
					put(blk, col, 2 * D, r.C);
				}
				// the epilogue corrects whole groups of four columns: null entries flush a zero into the unused slots of the last four
				const int nflag4 = (nflag_g[g] + 3) & ~3;
				for (int p = nflag_g[g]; p < nflag4; ++p) gl.push_back({0, 0.f, 0.f, 1 + 4 * p});
				while (gl.size() % 4) gl.push_back({0, 0.f, 0.f, 0});          // null entries: t = 0, no flush
				hdr[gi * 4 + 2] = nflag4;                       // .z: flagged columns of the group (its first nflag4)
			}
		}
		int pos = 0;
		for (int gi = 0; gi < 4 && !overflow; ++gi) {
			if (pos + (int)glist[gi].size() > TC_HSTRIDE - TC_HHEAD) {
				overflow = true;
				break;
			}
			hdr[gi * 4 + 0] = pos;                              // .x, .y: the group's entries [e0, e1)
			for (const HardEntry &he : glist[gi]) {
				int *e = hdr + (size_t)(TC_HHEAD + pos) * 4;
				e[0] = he.x;
				memcpy(&e[1], &he.a, 4);
				memcpy(&e[2], &he.nb, 4);
				e[3] = he.w;
				++pos;
			}
			hdr[gi * 4 + 1] = pos;
		}
	}
	if (overflow) {      // too many cancelling entries for the sparse list: this model set stays on the FMA route
		pm.h_tc2.clear();
		pm.h_tc_hard.clear();
		pm.h_tc_cs.clear();
	}
}

// ---------------------------------------------------------------------------------------------
// packing: host float64 parameters -> fp32 tiles (layout documented in gmm_model.h)
// ---------------------------------------------------------------------------------------------
void pack_models_host(int S, int K, int D, const double *const *weights, const double *const *means, const double *const *sigmas,
                      const int *Ks, std::vector<float> &out, int *Kpad_out) {
	const int Kpad = ceil_div(K, MT) * MT;
	const int tiles = Kpad / MT;
	const size_t tf = tile_floats(D);
	out.assign((size_t)S * tiles * tf, 0.f);
	const double A0 = 0.84932180028801904272;   // sqrt(log2(e)/2)
	for (int s = 0; s < S; ++s) {
		for (int r = 0; r < Kpad; ++r) {
			float *tile = out.data() + ((size_t)s * tiles + (r >> 5)) * tf;
			const int m = r & 31;
			float c2 = PAD_C2;
			if (r < Ks[s]) {
				const double *mu = means[s] + (size_t)r * D, *sg = sigmas[s] + (size_t)r * D;
				double c = weights[s][r] > 0.0 ? log2(weights[s][r]) : -1.0e30;
				for (int d = 0; d < D; ++d) {
					const double a = A0 / sg[d];
					float *t = tile + (size_t)d * 64 + (m >> 1) * 4 + (m & 1);   // chunk = [a0, a1, nb0, nb1]
					t[0] = (float)a;
					t[2] = (float)(-mu[d] * a);
					c -= log2(2.5066282746310002 * sg[d]);   // SQRT_2_PI, gmm.cc:22
				}
				c2 = (c < -1.0e30 || c != c) ? PAD_C2 : (float)c;
			}
			tile[64 * D + m] = c2;
		}
	}
	*Kpad_out = Kpad;
}

int upload_packed(PackedModels &pm, int S, int K, int D, const std::vector<float> &host, int Kpad, cudaStream_t st) {
	const size_t bytes = host.size() * sizeof(float);
	if (pm.d_tiles == nullptr || pm.bytes < bytes) {
		pm.release();
		SRB_CUDA(cudaMalloc(&pm.d_tiles, bytes));
		pm.bytes = bytes;
	}
	SRB_CUDA(cudaMemcpyAsync(pm.d_tiles, host.data(), bytes, cudaMemcpyHostToDevice, st));
	SRB_CUDA(cudaStreamSynchronize(st));
	pm.S = S;
	pm.K = K;
	pm.Kpad = Kpad;
	pm.D = D;
	pm.tiles_per_model = Kpad / MT;
	return 0;
}

int gmm_pull_device_master(GMM *g) {
	if (!g->host_stale) return 0;
	const int K = g->nr_mixtures, D = g->dim;
	std::vector<double> buf((size_t)K + 2 * (size_t)K * D);
	SRB_CUDA(cudaDeviceSynchronize());      // the M-step that wrote the master may have run on any (non-blocking) caller stream
	SRB_CUDA(cudaMemcpy(buf.data(), g->d_params, buf.size() * sizeof(double), cudaMemcpyDeviceToHost));
	g->weights.assign(buf.begin(), buf.begin() + K);
	g->means.assign(buf.begin() + K, buf.begin() + K + (size_t)K * D);
	g->sigmas.assign(buf.begin() + K + (size_t)K * D, buf.end());
	g->host_stale = false;
	return 0;
}

int gmm_ensure_device_master(GMM *g) {
	const int K = g->nr_mixtures, D = g->dim;
	SRB_REQUIRE(K > 0 && D > 0 && (int)g->weights.size() == K, "gmm: model has no parameters");
	if (g->master_valid && g->d_params && g->d_params_K == K && g->d_params_D == D) return 0;
	if (g->d_params && (g->d_params_K != K || g->d_params_D != D)) {
		cudaFree(g->d_params);
		g->d_params = nullptr;
	}
	const size_t n = (size_t)K + 2 * (size_t)K * D;
	if (!g->d_params) SRB_CUDA(cudaMalloc(&g->d_params, n * sizeof(double)));
	g->d_params_K = K;
	g->d_params_D = D;
	std::vector<double> buf;
	buf.reserve(n);
	buf.insert(buf.end(), g->weights.begin(), g->weights.end());
	buf.insert(buf.end(), g->means.begin(), g->means.end());
	buf.insert(buf.end(), g->sigmas.begin(), g->sigmas.end());
	SRB_CUDA(cudaMemcpy(g->d_params, buf.data(), n * sizeof(double), cudaMemcpyHostToDevice));
	g->master_valid = true;
	return 0;
}

int gmm_ensure_packed(GMM *g) {
	if (g->packed_valid && g->packed.tc_built) return 0;
	if (gmm_pull_device_master(g)) return 1;
	const int K = g->nr_mixtures, D = g->dim;
	SRB_REQUIRE(K > 0 && D > 0 && (int)g->weights.size() == K && (int)g->means.size() == K * D, "gmm: model has no parameters (train, load or set_params first)");
	SRB_REQUIRE(D <= MAX_DIM, "gmm: dim %d > %d unsupported", D, MAX_DIM);
	const double *w = g->weights.data(), *m = g->means.data(), *s = g->sigmas.data();
	if (!g->packed_valid) {
		std::vector<float> host;
		int Kpad = 0;
		pack_models_host(1, K, D, &w, &m, &s, &K, host, &Kpad);
		if (upload_packed(g->packed, 1, K, D, host, Kpad, ctx().stream)) return 1;
	}
	// the tensor-core images are (re)built whenever the FMA tiles were repacked -- from the host here, or from the device
	// master by the training loop (launch_pack_from_master drops them: a model scored, then trained, must not be scored with
	// the images of its old parameters)
	pack_models_tc_host(g->packed, 1, K, D, &w, &m, &s, &K);
	g->packed_valid = true;
	return 0;
}

}  // namespace srb

GMM::~GMM() {
	packed.release();
	if (d_params) cudaFree(d_params);
}

struct srb200_gmmset {
	srb::PackedModels packed;
	int S = 0, D = 0, K = 0;
};

struct srb200_mfcc {
	srb::MfccPlan *plan = nullptr;
};

using namespace srb;

#define API_ENTER()                                    \
	std::lock_guard<std::recursive_mutex> lk_(ctx().mu); \
	clear_error()

// ---------------------------------------------------------------------------------------------
// model text I/O -- format of GMM::dump/load (gmm.cc:655-682) and Gaussian::dump/load (gmm.cc:101-150)
// ---------------------------------------------------------------------------------------------
static int read_model_file(const char *path, GMM *g) {
	FILE *f = fopen(path, "rb");
	SRB_REQUIRE(f != nullptr, "load: cannot open '%s'", path);
	std::string txt;
	char buf[1 << 16];
	size_t n;
	while ((n = fread(buf, 1, sizeof buf, f)) > 0) txt.append(buf, n);
	fclose(f);
	const char *p = txt.c_str();
	auto next_num = [&](double *out) -> bool {
		char *end = nullptr;
		while (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r') ++p;
		if (!*p) return false;
		*out = strtod(p, &end);
		if (end == p) return false;
		p = end;
		return true;
	};
	double v;
	SRB_REQUIRE(next_num(&v) && v >= 1 && v < 1e7, "load: '%s' is not a GMM model file", path);
	const int K = (int)v;
	g->weights.resize(K);
	for (int k = 0; k < K; ++k) SRB_REQUIRE(next_num(&g->weights[k]), "load: '%s' truncated in the weights", path);
	int D = 0;
	for (int k = 0; k < K; ++k) {
		double dd, cov;
		SRB_REQUIRE(next_num(&dd) && next_num(&cov), "load: '%s' truncated at gaussian %d", path, k);
		SRB_REQUIRE((int)cov == 1, "load: only diagonal covariance supported (gmm.cc:211-215)");
		if (k == 0) {
			D = (int)dd;
			SRB_REQUIRE(D >= 1 && D <= 100000, "load: bad dimension in '%s'", path);
			g->means.resize((size_t)K * D);
			g->sigmas.resize((size_t)K * D);
		}
		SRB_REQUIRE((int)dd == D, "load: inconsistent dimensions in '%s'", path);
		for (int d = 0; d < D; ++d) SRB_REQUIRE(next_num(&g->means[(size_t)k * D + d]), "load: '%s' truncated", path);
		for (int d = 0; d < D; ++d) SRB_REQUIRE(next_num(&g->sigmas[(size_t)k * D + d]), "load: '%s' truncated", path);
	}
	g->nr_mixtures = K;
	g->dim = D;
	g->covariance_type = 1;
	g->invalidate();
	return 0;
}

static int write_model_file(GMM *g, const char *path, int digits) {
	if (gmm_pull_device_master(g)) return 1;
	SRB_REQUIRE((int)g->weights.size() == g->nr_mixtures && g->nr_mixtures > 0, "dump: model has no parameters");
	FILE *f = fopen(path, "wb");
	SRB_REQUIRE(f != nullptr, "dump: cannot open '%s' for writing", path);
	char fmt[16];
	snprintf(fmt, sizeof fmt, "%%.%dg ", digits);
	const int K = g->nr_mixtures, D = g->dim;
	fprintf(f, "%d\n", K);
	for (int k = 0; k < K; ++k) fprintf(f, fmt, g->weights[k]);
	fprintf(f, "\n");
	for (int k = 0; k < K; ++k) {
		fprintf(f, "%d %d\n", D, 1);
		for (int d = 0; d < D; ++d) fprintf(f, fmt, g->means[(size_t)k * D + d]);
		fprintf(f, "\n");
		for (int d = 0; d < D; ++d) fprintf(f, fmt, g->sigmas[(size_t)k * D + d]);
		fprintf(f, "\n");
	}
	const bool ok = !ferror(f);
	fclose(f);
	SRB_REQUIRE(ok, "dump: write error on '%s'", path);
	return 0;
}

// ---------------------------------------------------------------------------------------------
// host scoring helpers
// ---------------------------------------------------------------------------------------------
struct RowSrc {      // where the host frames come from
	const double *const *pp = nullptr;   // double** (reference ABI)
	const double *c64 = nullptr;         // contiguous [n][dim]
	const float *c32 = nullptr;
};

static int stage_frames(const RowSrc &src, int n, int dim, int *ld_out) {
	Context &c = ctx();
	const int ld = (dim + 3) & ~3;
	if (c.h_x.ensure((size_t)n * ld * sizeof(float)) || c.d_x.ensure((size_t)n * ld * sizeof(float))) return 1;
	float *h = c.h_x.as<float>();
	for (int i = 0; i < n; ++i) {
		float *row = h + (size_t)i * ld;
		if (src.pp) {
			const double *r = src.pp[i];
			for (int d = 0; d < dim; ++d) row[d] = (float)r[d];
		} else if (src.c64) {
			const double *r = src.c64 + (size_t)i * dim;
			for (int d = 0; d < dim; ++d) row[d] = (float)r[d];
		} else {
			const float *r = src.c32 + (size_t)i * dim;
			for (int d = 0; d < dim; ++d) row[d] = r[d];
		}
		for (int d = dim; d < ld; ++d) row[d] = 0.f;
	}
	SRB_CUDA(cudaMemcpyAsync(c.d_x.p, h, (size_t)n * ld * sizeof(float), cudaMemcpyHostToDevice, c.stream));
	*ld_out = ld;
	return 0;
}

// scores n staged frames against pm; returns sums[S] and optionally per-frame values for model 0
static int score_staged(PackedModels &pm, int n, int ld, int dim, double *sums, float *frame_out_host) {
	Context &c = ctx();
	const int S = pm.S;
	if (c.d_sums.ensure((size_t)S * sizeof(double)) || c.h_out.ensure(std::max((size_t)S * sizeof(double), (size_t)n * sizeof(float)))) return 1;
	float *d_frame = nullptr;
	if (frame_out_host) {
		if (c.d_frame.ensure((size_t)n * sizeof(float))) return 1;
		d_frame = c.d_frame.as<float>();
	}
	const int utt_off[2] = {0, n};
	if (launch_score(pm, c.d_x.as<float>(), ld, dim, utt_off, 1, c.d_sums.as<double>(), d_frame, n, nullptr, c.compat_clamp, 0, c.stream)) return 1;
	SRB_CUDA(cudaMemcpyAsync(c.h_out.p, c.d_sums.p, (size_t)S * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
	SRB_CUDA(cudaStreamSynchronize(c.stream));
	memcpy(sums, c.h_out.p, (size_t)S * sizeof(double));
	if (frame_out_host) {
		SRB_CUDA(cudaMemcpyAsync(c.h_out.p, d_frame, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
		SRB_CUDA(cudaStreamSynchronize(c.stream));
		memcpy(frame_out_host, c.h_out.p, (size_t)n * sizeof(float));
	}
	return 0;
}

static int gmm_score_host(GMM *g, const RowSrc &src, int n, int dim, double *frame_out64, float *frame_out32, double *sum_out) {
	SRB_REQUIRE(g != nullptr, "score: null model");
	if (ctx().init()) return 1;
	if (gmm_ensure_packed(g)) return 1;
	SRB_REQUIRE(dim == g->dim, "score: nr_dim %d != model dim %d (the reference asserts here, gmm.cc:177)", dim, g->dim);
	if (n <= 0) {
		if (sum_out) *sum_out = 0.0;
		return 0;
	}
	int ld = 0;
	if (stage_frames(src, n, dim, &ld)) return 1;
	std::vector<float> tmp;
	float *fo = frame_out32;
	if (frame_out64 && !fo) {
		tmp.resize(n);
		fo = tmp.data();
	}
	double sum = 0.0;
	if (score_staged(g->packed, n, ld, dim, &sum, fo)) return 1;
	if (frame_out64)
		for (int i = 0; i < n; ++i) frame_out64[i] = (double)fo[i];
	if (sum_out) *sum_out = sum;
	return 0;
}

// ---------------------------------------------------------------------------------------------
// training driver (GMMTrainerBaseline::train, gmm.cc:581-653, on device-resident frames)
// ---------------------------------------------------------------------------------------------
static int em_estep(GMM *g, const float *dX, int n, int ld, int D, double *d_stats, cudaStream_t st) {
	Context &c = ctx();
	if (c.d_frame.ensure((size_t)n * sizeof(float)) || c.d_sums.ensure(sizeof(double))) return 1;      // no-ops once the workspaces are large enough
	const int utt_off[2] = {0, n};
	// large models: both passes on the tensor cores, operand images built on the device from the float64 master (gmm_em_tc.cu);
	// parameters the uncorrected expanded form is not accurate for (and small jobs) take the FMA kernels below
	c.last_em_route = 0;
	if (em_tc_eligible(g->nr_mixtures, D, n)) {
		int used = 0;
		if (gmm_ensure_device_master(g)) return 1;
		if (launch_estep_tc(g->packed, g->d_params, dX, n, ld, D, g->nr_mixtures, c.d_frame.as<float>(), c.d_sums.as<double>(), d_stats, c.compat_clamp, st,
		                    &used))
			return 1;
		if (used) {
			c.last_em_route = 1;
			return 0;
		}
	}
	// pass A: per-frame log2 p(x) (unclamped) + total log-likelihood (clamped like the reference's monitor, gmm.cc:635)
	if (launch_score(g->packed, dX, ld, D, utt_off, 1, c.d_sums.as<double>(), c.d_frame.as<float>(), n, nullptr, c.compat_clamp, 1, st)) return 1;
	// pass B: responsibilities + sufficient statistics
	return launch_estep_accumulate(g->packed, dX, n, ld, D, g->nr_mixtures, c.d_frame.as<float>(), c.d_sums.as<double>(), d_stats,
	                               c.compat_clamp, st);
}

static int train_driver(GMM *g, const float *dX, int n, int ld, int D, const Parameter *p, int mode, GMM *ubm, int *iters_out, double *ll_out) {
	Context &c = ctx();
	const int K = g->nr_mixtures;
	g->master_valid = false;
	if (gmm_ensure_device_master(g)) return 1;
	if (launch_pack_from_master(g->d_params, K, D, g->packed, c.stream)) return 1;
	g->packed_valid = true;
	const double *d_ubm = nullptr;
	if (mode == 1) {
		if (gmm_pull_device_master(ubm)) return 1;
		ubm->master_valid = ubm->master_valid && !ubm->host_stale;
		if (gmm_ensure_device_master(ubm)) return 1;
		d_ubm = ubm->d_params;
	}
	if (c.d_stats.ensure((size_t)em_stats_len(K, D) * sizeof(double))) return 1;
	double *d_stats = c.d_stats.as<double>();
	double last_ll = -DBL_MAX, ll = 0.0;
	int iters = 0;
	for (int i = 0; i < p->nr_iteration; ++i) {
		if (em_estep(g, dX, n, ld, D, d_stats, c.stream)) return 1;
		if (launch_mstep(g->d_params, d_ubm, K, D, d_stats, (long long)n, p->min_covar, mode, c.stream)) return 1;
		if (launch_pack_from_master(g->d_params, K, D, g->packed, c.stream)) return 1;
		++iters;
		if (i % 2 == 0) continue;                       // gmm.cc:622
		// monitor the log-likelihood of the UPDATED model (gmm.cc:631-651)
		const int utt_off[2] = {0, n};
		if (launch_score(g->packed, dX, ld, D, utt_off, 1, c.d_sums.as<double>(), nullptr, 0, nullptr, c.compat_clamp, 0, c.stream)) return 1;
		SRB_CUDA(cudaMemcpyAsync(&ll, c.d_sums.p, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
		SRB_CUDA(cudaStreamSynchronize(c.stream));
		if (p->verbosity >= 1) {
			printf("iter %d: ll %lf\n", i, ll);
			fflush(stdout);
		}
		const double ll_diff = ll - last_ll;
		if (fabs(ll_diff) / fabs(ll) < p->threshold && ll_diff < p->threshold) {
			if (p->verbosity >= 1) printf("too small log likelihood increment, abort iteration.\n");
			break;
		}
		last_ll = ll;
	}
	SRB_CUDA(cudaStreamSynchronize(c.stream));
	g->host_stale = true;
	g->master_valid = true;
	if (gmm_pull_device_master(g)) return 1;
	if (iters_out) *iters_out = iters;
	if (ll_out) *ll_out = ll;
	return 0;
}

static void print_param(const Parameter *p);

// Batched MAP enrolment (what ModelInterface.train does with a UBM, src/gui/interface.py:77-83 -> gmmset.py:30-34 ->
// train_model_from_ubm per speaker): every speaker's model starts as a copy of the UBM and runs GMMUBMTrainerBaseline's loop
// (gmmubm.cc:40-81 inside GMMTrainerBaseline::train, gmm.cc:613-652) on its own frames.  Here ALL speakers' frames are staged by ONE
// host -> device copy and the speakers advance in lockstep: per iteration the E-step / M-step / repack kernels of every still-active
// speaker are enqueued back to back on one stream, and the log-likelihood monitor of the odd iterations is ONE device -> host copy
// of all speakers' values -- the host synchronises nr_iteration / 2 times in total instead of once per speaker and check.
static int train_many_from_ubm(GMM *const *models, int n_models, GMM *ubm, const RowSrc &src, const int *frame_off, int D, const Parameter *p,
                               int *iters_out, double *ll_out) {
	Context &c = ctx();
	SRB_REQUIRE(models && n_models >= 1 && ubm && frame_off && p, "train_many: null argument");
	if (c.init()) return 1;
	if (gmm_pull_device_master(ubm)) return 1;
	SRB_REQUIRE(ubm->dim == D && (int)ubm->weights.size() == ubm->nr_mixtures && ubm->nr_mixtures >= 1, "train_many: UBM dim %d != data dim %d (or empty UBM)", ubm->dim, D);
	SRB_REQUIRE(D >= 1 && D <= MAX_DIM, "train_many: nr_dim %d outside [1,%d]", D, MAX_DIM);
	const int n_total = frame_off[n_models];
	int n_max = 0;
	for (int s = 0; s < n_models; ++s) {
		SRB_REQUIRE(models[s] != nullptr && models[s] != ubm, "train_many: model %d is null or the UBM itself", s);
		SRB_REQUIRE(frame_off[s + 1] > frame_off[s], "X.size() == 0");       // gmm.cc:582-586, per speaker
		n_max = std::max(n_max, frame_off[s + 1] - frame_off[s]);
	}
	if (p->verbosity >= 1) print_param(p);
	int ld = 0;
	if (stage_frames(src, n_total, D, &ld)) return 1;
	const float *dX = c.d_x.as<float>();
	const int K = ubm->nr_mixtures;
	ubm->master_valid = ubm->master_valid && !ubm->host_stale;
	if (gmm_ensure_device_master(ubm)) return 1;
	// workspaces sized once for the largest speaker: nothing below reallocates, so nothing below synchronises
	if (c.d_frame.ensure((size_t)n_max * sizeof(float)) || c.d_sums.ensure(sizeof(double)) || c.d_stats.ensure((size_t)em_stats_len(K, D) * sizeof(double)) ||
	    c.d_misc.ensure((size_t)n_models * sizeof(double)) || c.h_misc.ensure((size_t)n_models * sizeof(double)))
		return 1;
	double *d_stats = c.d_stats.as<double>(), *d_ll = c.d_misc.as<double>(), *h_ll = c.h_misc.as<double>();
	for (int s = 0; s < n_models; ++s) {                              // gmm_replace_with (gmmubm.cc:29-38): deep copy of the UBM
		GMM *g = models[s];
		g->nr_mixtures = K;
		g->dim = D;
		g->weights = ubm->weights;
		g->means = ubm->means;
		g->sigmas = ubm->sigmas;
		g->invalidate();
		if (gmm_ensure_device_master(g)) return 1;
		if (launch_pack_from_master(g->d_params, K, D, g->packed, c.stream)) return 1;
		g->packed_valid = true;
	}
	std::vector<char> active(n_models, 1);
	std::vector<double> last_ll(n_models, -DBL_MAX), ll(n_models, 0.0);
	std::vector<int> iters(n_models, 0);
	int n_active = n_models;
	for (int i = 0; i < p->nr_iteration && n_active > 0; ++i) {
		for (int s = 0; s < n_models; ++s) {
			if (!active[s]) continue;
			GMM *g = models[s];
			const int n = frame_off[s + 1] - frame_off[s];
			const float *xs = dX + (size_t)frame_off[s] * ld;
			if (em_estep(g, xs, n, ld, D, d_stats, c.stream)) return 1;
			if (launch_mstep(g->d_params, ubm->d_params, K, D, d_stats, (long long)n, p->min_covar, 1, c.stream)) return 1;
			if (launch_pack_from_master(g->d_params, K, D, g->packed, c.stream)) return 1;
			++iters[s];
		}
		if (i % 2 == 0) continue;                                     // gmm.cc:622
		for (int s = 0; s < n_models; ++s) {                         // log-likelihood of the UPDATED models (gmm.cc:631-651)
			if (!active[s]) continue;
			const int n = frame_off[s + 1] - frame_off[s];
			const int utt_off[2] = {0, n};
			if (launch_score(models[s]->packed, dX + (size_t)frame_off[s] * ld, ld, D, utt_off, 1, d_ll + s, nullptr, 0, nullptr, c.compat_clamp, 0, c.stream))
				return 1;
		}
		SRB_CUDA(cudaMemcpyAsync(h_ll, d_ll, (size_t)n_models * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
		SRB_CUDA(cudaStreamSynchronize(c.stream));                   // the ONE synchronisation of this monitoring round
		for (int s = 0; s < n_models; ++s) {
			if (!active[s]) continue;
			ll[s] = h_ll[s];
			if (p->verbosity >= 1) printf("model %d iter %d: ll %lf\n", s, i, ll[s]);
			const double ll_diff = ll[s] - last_ll[s];
			if (fabs(ll_diff) / fabs(ll[s]) < p->threshold && ll_diff < p->threshold) {
				active[s] = 0;
				--n_active;
			}
			last_ll[s] = ll[s];
		}
	}
	SRB_CUDA(cudaStreamSynchronize(c.stream));
	for (int s = 0; s < n_models; ++s) {
		GMM *g = models[s];
		g->host_stale = true;
		g->master_valid = true;
		if (gmm_pull_device_master(g)) return 1;
		if (iters_out) iters_out[s] = iters[s];
		if (ll_out) ll_out[s] = ll[s];
	}
	return 0;
}

static uint64_t g_train_calls = 0;
static inline uint64_t splitmix64(uint64_t &s) {
	uint64_t z = (s += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

// init_gaussians (gmm.cc:306-361): sigma = global unbiased std, w = 1/K; means = random data rows, or (init_with_kmeans) the
// K-MeansII centroids computed on the staged device features (kmeans.cu)
static int init_like_reference(GMM *g, const RowSrc &src, int n, int D, uint64_t seed, const float *d_staged, int ld, int with_kmeans,
                               int verbosity) {
	const int K = g->nr_mixtures;
	auto row = [&](int i, int d) -> double { return src.pp ? src.pp[i][d] : src.c64[(size_t)i * D + d]; };
	std::vector<double> mean(D, 0.0), var(D, 0.0);
	for (int i = 0; i < n; ++i)
		for (int d = 0; d < D; ++d) mean[d] += row(i, d);
	for (int d = 0; d < D; ++d) mean[d] /= n;
	for (int i = 0; i < n; ++i)
		for (int d = 0; d < D; ++d) {
			const double t = row(i, d) - mean[d];
			var[d] += t * t;
		}
	g->dim = D;
	g->weights.assign(K, 1.0 / K);
	g->means.resize((size_t)K * D);
	g->sigmas.resize((size_t)K * D);
	if (with_kmeans) {
		if (kmeans2_init_device(d_staged, n, ld, D, K, seed, g->means.data(), nullptr, verbosity)) return 1;   // gmm.cc:331-344
	}
	uint64_t s = seed;
	for (int k = 0; k < K; ++k) {
		const double u = (double)(splitmix64(s) >> 11) * (1.0 / 9007199254740992.0);
		int idx = (int)(u * n);                          // rand_int: rand_real() * max_val, random.hh:46-48
		if (idx >= n) idx = n - 1;
		for (int d = 0; d < D; ++d) {
			if (!with_kmeans) g->means[(size_t)k * D + d] = row(idx, d);
			g->sigmas[(size_t)k * D + d] = sqrt(var[d] * (1.0 / (n - 1)));   // gmm.cc:323-325
		}
	}
	g->invalidate();
	return 0;
}

static void print_param(const Parameter *p) {   // pygmm.cc:31-41 (the reference prints on every train; here only when verbose)
	printf("nr_instance   :   %d\nnr_dim        :   %d\nnr_mixture    :   %d\nmin_covar     :   %f\nthreshold     :   %f\n"
	       "nr_iteration  :   %d\ninit_with_kmeans: %d\nconcurrency   :   %d\nverbosity     :   %d\n",
	       p->nr_instance, p->nr_dim, p->nr_mixture, p->min_covar, p->threshold, p->nr_iteration, p->init_with_kmeans, p->concurrency,
	       p->verbosity);
}

static int train_host(GMM *g, GMM *ubm, const RowSrc &src, const Parameter *p, uint64_t seed, int *iters_out, double *ll_out) {
	SRB_REQUIRE(g != nullptr && p != nullptr, "train: null argument");
	if (ctx().init()) return 1;
	const int n = p->nr_instance, D = p->nr_dim;
	SRB_REQUIRE(n > 0, "X.size() == 0");                 // gmm.cc:582-586 (thrown as const char* in the reference)
	SRB_REQUIRE(D >= 1 && D <= MAX_DIM, "train: nr_dim %d outside [1,%d]", D, MAX_DIM);
	if (p->verbosity >= 1) print_param(p);
	int ld = 0;
	if (stage_frames(src, n, D, &ld)) return 1;
	if (ubm == nullptr) {
		SRB_REQUIRE(g->nr_mixtures >= 1, "train: model has no mixtures");
		SRB_REQUIRE(n >= 2, "train: need at least 2 instances");
		if (init_like_reference(g, src, n, D, seed, ctx().d_x.as<float>(), ld, p->init_with_kmeans != 0, p->verbosity)) return 1;
	} else {
		if (gmm_pull_device_master(ubm)) return 1;
		SRB_REQUIRE(ubm->dim == D, "train_from_ubm: UBM dim %d != data dim %d", ubm->dim, D);
		// gmm_replace_with (gmmubm.cc:29-38): the model becomes a deep copy of the UBM
		g->nr_mixtures = ubm->nr_mixtures;
		g->dim = ubm->dim;
		g->weights = ubm->weights;
		g->means = ubm->means;
		g->sigmas = ubm->sigmas;
		g->invalidate();
	}
	return train_driver(g, ctx().d_x.as<float>(), n, ld, D, p, ubm ? 1 : 0, ubm, iters_out, ll_out);
}

// =============================================================================================
// extern "C"
// =============================================================================================
extern "C" {

const char *srb200_last_error(void) { return t_err.c_str(); }
const char *srb200_version(void) { return "srb200 0.1.0 (sm_100a)"; }

int srb200_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}
int srb200_set_device(int device) {
	API_ENTER();
	Context &c = ctx();
	if (c.inited) {
		SRB_REQUIRE(device == c.device, "set_device: the library is already bound to device %d", c.device);
		return 0;
	}
	c.device = device;
	return c.init();
}
int srb200_get_device(void) { return ctx().device; }
int srb200_set_compat_clamp(int mode) {
	API_ENTER();
	SRB_REQUIRE(mode == 0 || mode == 1, "compat_clamp must be 0 or 1");
	ctx().compat_clamp = mode;
	return 0;
}
int srb200_get_compat_clamp(void) { return ctx().compat_clamp; }
int srb200_set_score_path(int path) {
	API_ENTER();
	SRB_REQUIRE(path >= -1 && path <= 1, "score path must be -1 (default), 0 (fp32 FMA) or 1 (tcgen05 3xFP16 on CTA pairs)");
	ctx().score_path = path;
	return 0;
}
int srb200_get_score_path(void) { return ctx().score_path; }
int srb200_last_score_route(void) { return ctx().last_score_route; }
int srb200_last_em_route(void) { return ctx().last_em_route; }
int srb200_set_em_path(int path) {
	API_ENTER();
	SRB_REQUIRE(path >= -1 && path <= 1, "em path must be -1 (default), 0 (fp32 FMA kernels) or 1 (tcgen05 wherever the shape allows)");
	ctx().em_path = path;
	return 0;
}
long long srb200_score_launch_count(void) { return ctx().score_launches; }

GMM *new_gmm(int nr_mixture, int covariance_type) {
	API_ENTER();
	if (covariance_type != 1) {
		set_error("only diagonal matrix supported.");   // gmm.cc:211-215
		return nullptr;
	}
	if (nr_mixture < 1) {
		set_error("new_gmm: nr_mixture must be >= 1");
		return nullptr;
	}
	GMM *g = new GMM();
	g->nr_mixtures = nr_mixture;
	return g;
}

GMM *load(const char *model_file) {
	API_ENTER();
	GMM *g = new GMM();
	if (read_model_file(model_file, g)) {
		delete g;
		return nullptr;
	}
	return g;
}

void dump(GMM *gmm, const char *model_file) {
	API_ENTER();
	if (!gmm) {
		set_error("dump: null model");
		return;
	}
	write_model_file(gmm, model_file, 6);   // ostream default precision, gmm.cc:655-662
}

int srb200_gmm_dump_precise(GMM *gmm, const char *model_file, int digits) {
	API_ENTER();
	SRB_REQUIRE(gmm != nullptr, "dump: null model");
	SRB_REQUIRE(digits >= 1 && digits <= 17, "dump: digits must be in [1,17]");
	return write_model_file(gmm, model_file, digits);
}

void free_gmm(GMM *gmm) {
	API_ENTER();
	delete gmm;
}

int get_dim(GMM *gmm) { return gmm ? gmm->dim : 0; }
int get_nr_mixtures(GMM *gmm) { return gmm ? gmm->nr_mixtures : 0; }

int srb200_gmm_set_params(GMM *g, int K, int D, const double *w, const double *m, const double *s) {
	API_ENTER();
	SRB_REQUIRE(g && w && m && s, "set_params: null argument");
	SRB_REQUIRE(K >= 1 && D >= 1 && D <= MAX_DIM, "set_params: bad shape K=%d D=%d", K, D);
	for (size_t i = 0; i < (size_t)K * D; ++i) SRB_REQUIRE(s[i] > 0.0, "set_params: sigma must be positive");
	g->nr_mixtures = K;
	g->dim = D;
	g->weights.assign(w, w + K);
	g->means.assign(m, m + (size_t)K * D);
	g->sigmas.assign(s, s + (size_t)K * D);
	g->invalidate();
	return 0;
}

int srb200_gmm_get_params(GMM *g, double *w, double *m, double *s) {
	API_ENTER();
	SRB_REQUIRE(g != nullptr, "get_params: null model");
	if (gmm_pull_device_master(g)) return 1;
	SRB_REQUIRE((int)g->weights.size() == g->nr_mixtures && g->dim > 0, "get_params: model has no parameters");
	if (w) memcpy(w, g->weights.data(), g->weights.size() * sizeof(double));
	if (m) memcpy(m, g->means.data(), g->means.size() * sizeof(double));
	if (s) memcpy(s, g->sigmas.data(), g->sigmas.size() * sizeof(double));
	return 0;
}

int srb200_gmm_sync_from_device(GMM *g) {
	API_ENTER();
	SRB_REQUIRE(g != nullptr, "null model");
	SRB_CUDA(cudaDeviceSynchronize());
	return gmm_pull_device_master(g);
}

// ---- scoring: the reference ABI ----
double score_all(GMM *gmm, double **X_in, int nr_instance, int nr_dim, int /*concurrency*/) {
	API_ENTER();
	RowSrc src;
	src.pp = X_in;
	double sum = 0.0;
	if (gmm_score_host(gmm, src, nr_instance, nr_dim, nullptr, nullptr, &sum)) return NAN;
	return sum;
}

void score_batch(GMM *gmm, double **X_in, double *prob_out, int nr_instance, int nr_dim, int /*concurrency*/) {
	API_ENTER();
	RowSrc src;
	src.pp = X_in;
	if (gmm_score_host(gmm, src, nr_instance, nr_dim, prob_out, nullptr, nullptr))
		for (int i = 0; i < nr_instance; ++i) prob_out[i] = NAN;
}

double score_instance(GMM *gmm, double *x_in, int nr_dim) {
	API_ENTER();
	RowSrc src;
	src.c64 = x_in;
	double sum = 0.0;
	if (gmm_score_host(gmm, src, 1, nr_dim, nullptr, nullptr, &sum)) return NAN;
	return sum;
}

int srb200_score_batch_f64(GMM *gmm, const double *X, int n, int dim, double *prob_out) {
	API_ENTER();
	RowSrc src;
	src.c64 = X;
	return gmm_score_host(gmm, src, n, dim, prob_out, nullptr, nullptr);
}
int srb200_score_all_f64(GMM *gmm, const double *X, int n, int dim, double *sum_out) {
	API_ENTER();
	RowSrc src;
	src.c64 = X;
	return gmm_score_host(gmm, src, n, dim, nullptr, nullptr, sum_out);
}
int srb200_score_batch_f32(GMM *gmm, const float *X, int n, int dim, float *prob_out) {
	API_ENTER();
	RowSrc src;
	src.c32 = X;
	return gmm_score_host(gmm, src, n, dim, nullptr, prob_out, nullptr);
}

// ---- training ----
void train_model(GMM *gmm, double **X_in, Parameter *param) {
	API_ENTER();
	RowSrc src;
	src.pp = X_in;
	train_host(gmm, nullptr, src, param, 0x5EEDull + (++g_train_calls), nullptr, nullptr);
}
void train_model_from_ubm(GMM *gmm, GMM *ubm, double **X_in, Parameter *param) {
	API_ENTER();
	if (!ubm) {
		set_error("train_model_from_ubm: null ubm");
		return;
	}
	RowSrc src;
	src.pp = X_in;
	train_host(gmm, ubm, src, param, 0, nullptr, nullptr);
}
int srb200_kmeans_init_f64(const double *X, int n, int dim, int K, uint64_t seed, double *centroids, double *cost_out) {
	API_ENTER();
	SRB_REQUIRE(X != nullptr && centroids != nullptr, "kmeans_init: null argument");
	SRB_REQUIRE(n >= 1 && K >= 1 && dim >= 1 && dim <= MAX_DIM, "kmeans_init: bad shape n=%d K=%d dim=%d", n, K, dim);
	if (ctx().init()) return 1;
	RowSrc src;
	src.c64 = X;
	int ld = 0;
	if (stage_frames(src, n, dim, &ld)) return 1;
	return kmeans2_init_device(ctx().d_x.as<float>(), n, ld, dim, K, seed, centroids, cost_out, 0);
}

int srb200_kmeans_lloyd_f64(const double *X, int n, int dim, int K, const double *weight, double *centroids, double *cost_out, int *iters_out) {
	API_ENTER();
	SRB_REQUIRE(X != nullptr && centroids != nullptr, "kmeans_lloyd: null argument");
	SRB_REQUIRE(n >= 1 && K >= 1 && dim >= 1 && dim <= MAX_DIM, "kmeans_lloyd: bad shape n=%d K=%d dim=%d", n, K, dim);
	if (ctx().init()) return 1;
	RowSrc src;
	src.c64 = X;
	int ld = 0;
	if (stage_frames(src, n, dim, &ld)) return 1;
	return kmeans_lloyd_device(ctx().d_x.as<float>(), n, ld, dim, K, weight, centroids, cost_out, iters_out);
}

int srb200_train_f64(GMM *gmm, const double *X, const Parameter *param, uint64_t seed, int *iters_out, double *ll_out) {
	API_ENTER();
	RowSrc src;
	src.c64 = X;
	return train_host(gmm, nullptr, src, param, seed, iters_out, ll_out);
}
int srb200_train_from_ubm_f64(GMM *gmm, GMM *ubm, const double *X, const Parameter *param, int *iters_out, double *ll_out) {
	API_ENTER();
	SRB_REQUIRE(ubm != nullptr, "train_from_ubm: null ubm");
	RowSrc src;
	src.c64 = X;
	return train_host(gmm, ubm, src, param, 0, iters_out, ll_out);
}

int srb200_train_many_from_ubm_f64(GMM *const *models, int n_models, GMM *ubm, const double *X, const int *frame_off, int dim, const Parameter *param,
                                   int *iters_out, double *ll_out) {
	API_ENTER();
	RowSrc src;
	src.c64 = X;
	return train_many_from_ubm(models, n_models, ubm, src, frame_off, dim, param, iters_out, ll_out);
}

// ---- EM building blocks on device data ----
int srb200_em_stats_len(int K, int D) { return em_stats_len(K, D); }

int srb200_em_estep_dev(GMM *g, const float *dX, int n, int ld, int dim, double *d_stats, void *stream) {
	API_ENTER();
	SRB_REQUIRE(g && dX && d_stats, "estep: null argument");
	if (ctx().init()) return 1;
	SRB_REQUIRE(dim == g->dim, "estep: dim %d != model dim %d", dim, g->dim);
	if (!g->packed_valid) {
		// pack from the device master when it is the newest copy, else from the host arrays
		if (g->host_stale) {
			if (launch_pack_from_master(g->d_params, g->nr_mixtures, g->dim, g->packed, (cudaStream_t)stream)) return 1;
			g->packed_valid = true;
		} else if (gmm_ensure_packed(g)) return 1;
	}
	return em_estep(g, dX, n, ld, dim, d_stats, (cudaStream_t)stream);
}

int srb200_em_mstep_dev(GMM *g, const double *d_stats, long long n_total, double min_covar, int mode, GMM *ubm, void *stream) {
	API_ENTER();
	SRB_REQUIRE(g && d_stats, "mstep: null argument");
	if (ctx().init()) return 1;
	cudaStream_t st = (cudaStream_t)stream;
	if (!g->host_stale) g->master_valid = g->master_valid && g->d_params != nullptr;
	if (gmm_ensure_device_master(g)) return 1;
	const double *d_ubm = nullptr;
	if (mode == 1) {
		SRB_REQUIRE(ubm != nullptr, "mstep: MAP mode needs a UBM");
		if (gmm_ensure_device_master(ubm)) return 1;
		d_ubm = ubm->d_params;
	}
	if (launch_mstep(g->d_params, d_ubm, g->nr_mixtures, g->dim, d_stats, n_total, min_covar, mode, st)) return 1;
	if (launch_pack_from_master(g->d_params, g->nr_mixtures, g->dim, g->packed, st)) return 1;
	g->packed_valid = true;
	g->host_stale = true;
	g->master_valid = true;
	return 0;
}

// ---- speaker sets ----
srb200_gmmset *srb200_gmmset_create(GMM *const *models, int n_models) {
	API_ENTER();
	if (!models || n_models < 1) {
		set_error("gmmset_create: no models");
		return nullptr;
	}
	if (ctx().init()) return nullptr;
	int K = 0, D = 0;
	std::vector<const double *> w(n_models), m(n_models), s(n_models);
	std::vector<int> Ks(n_models);
	for (int i = 0; i < n_models; ++i) {
		GMM *g = models[i];
		if (!g || gmm_pull_device_master(g)) return nullptr;
		if ((int)g->weights.size() != g->nr_mixtures || g->dim < 1) {
			set_error("gmmset_create: model %d has no parameters", i);
			return nullptr;
		}
		if (i == 0) D = g->dim;
		if (g->dim != D) {
			set_error("gmmset_create: model %d has dim %d, expected %d", i, g->dim, D);
			return nullptr;
		}
		K = std::max(K, g->nr_mixtures);
		Ks[i] = g->nr_mixtures;
		w[i] = g->weights.data();
		m[i] = g->means.data();
		s[i] = g->sigmas.data();
	}
	if (D > MAX_DIM) {
		set_error("gmmset_create: dim %d > %d unsupported", D, MAX_DIM);
		return nullptr;
	}
	std::vector<float> host;
	int Kpad = 0;
	pack_models_host(n_models, K, D, w.data(), m.data(), s.data(), Ks.data(), host, &Kpad);
	srb200_gmmset *set = new srb200_gmmset();
	if (upload_packed(set->packed, n_models, K, D, host, Kpad, ctx().stream)) {
		delete set;
		return nullptr;
	}
	pack_models_tc_host(set->packed, n_models, K, D, w.data(), m.data(), s.data(), Ks.data());
	set->S = n_models;
	set->D = D;
	set->K = K;
	return set;
}

// host-only diagnostics of the tensor-core route (no device work, usable without a GPU): run the packer on temporary host
// buffers.  srb200_tc_pack_stats: is the set eligible, and with how many exact add-back entries?  srb200_tc_pack_export: the
// packed operands themselves (what the kernel will read), for the CPU emulation in tests/test_tc_pack_cpu.py.
static int tc_pack_tmp(GMM *const *models, int n_models, PackedModels &pm, int *D_out) {
	if (!models || n_models < 1) {
		set_error("tc_pack: no models");
		return -1;
	}
	int K = 0, D = 0;
	std::vector<const double *> w(n_models), m(n_models), s(n_models);
	std::vector<int> Ks(n_models);
	for (int i = 0; i < n_models; ++i) {
		GMM *g = models[i];
		if (!g || g->host_stale || (int)g->weights.size() != g->nr_mixtures || g->dim < 1 || (i > 0 && g->dim != D)) {
			set_error("tc_pack: model %d has no host parameters or a different dimension", i);
			return -1;
		}
		D = g->dim;
		K = std::max(K, g->nr_mixtures);
		Ks[i] = g->nr_mixtures;
		w[i] = g->weights.data();
		m[i] = g->means.data();
		s[i] = g->sigmas.data();
	}
	pack_models_tc_host(pm, n_models, K, D, w.data(), m.data(), s.data(), Ks.data());
	pm.S = n_models;
	*D_out = D;
	return tc2_supported(pm, D) ? 1 : 0;
}

int srb200_tc_pack_stats(GMM *const *models, int n_models, int *blocks_out, int *hard_out, int *max_hard_per_block_out) {
	API_ENTER();
	PackedModels pm;
	int D = 0;
	const int ok = tc_pack_tmp(models, n_models, pm, &D);
	if (ok < 0) return ok;
	int hard = 0, mx = 0;
	if (ok)
		for (int b = 0; b < pm.tc_blocks_total; ++b) {
			int n = 0;
			for (int e = 0; e < pm.h_tc_hard[(size_t)b * TC_HSTRIDE * 4 + 13]; ++e)
				n += pm.h_tc_hard[((size_t)b * TC_HSTRIDE + TC_HHEAD + e) * 4 + 1] != 0;        // null padding entries have a = 0
			hard += n;
			mx = std::max(mx, n);
		}
	if (blocks_out) *blocks_out = pm.tc_blocks_total;
	if (hard_out) *hard_out = hard;
	if (max_hard_per_block_out) *max_hard_per_block_out = mx;
	return ok;
}

// info_out[8] = {Kp, blocks, columns per speaker, halfs per block, ints per block list, header words, 0, 0}; the arrays are
// copied when non-NULL (sizes: blocks * info[3] uint16, blocks * info[4] int32, 2 * D float)
int srb200_tc_pack_export(GMM *const *models, int n_models, int *info_out, uint16_t *images_out, int *hard_out, float *cs_out) {
	API_ENTER();
	PackedModels pm;
	int D = 0;
	const int ok = tc_pack_tmp(models, n_models, pm, &D);
	if (ok <= 0) return ok;
	if (info_out) {
		const int info[8] = {pm.tc2_Kp, pm.tc_blocks_total, pm.tc_cols_per_spk, (int)t2_block_halfs(pm.tc2_Kp), TC_HSTRIDE * 4, TC_HHEAD, 0, 0};
		memcpy(info_out, info, sizeof info);
	}
	if (images_out) memcpy(images_out, pm.h_tc2.data(), pm.h_tc2.size() * sizeof(uint16_t));
	if (hard_out) memcpy(hard_out, pm.h_tc_hard.data(), pm.h_tc_hard.size() * sizeof(int));
	if (cs_out) memcpy(cs_out, pm.h_tc_cs.data(), pm.h_tc_cs.size() * sizeof(float));
	return 1;
}

void srb200_gmmset_free(srb200_gmmset *set) {
	API_ENTER();
	if (set) {
		set->packed.release();
		delete set;
	}
}
int srb200_gmmset_size(srb200_gmmset *set) { return set ? set->S : 0; }
int srb200_gmmset_set_ranked(srb200_gmmset *set, int n_ranked) {
	API_ENTER();
	SRB_REQUIRE(set != nullptr && n_ranked >= 1 && n_ranked <= set->S, "gmmset_set_ranked: need 1 <= n_ranked <= %d", set ? set->S : 0);
	set->packed.S_arg = n_ranked;
	return 0;
}
int srb200_gmmset_dim(srb200_gmmset *set) { return set ? set->D : 0; }

static int gmmset_score_host(srb200_gmmset *set, const RowSrc &src, int n, int dim, double *sums) {
	SRB_REQUIRE(set && sums, "gmmset_score: null argument");
	SRB_REQUIRE(dim == set->D, "gmmset_score: dim %d != set dim %d", dim, set->D);
	if (n <= 0) {
		for (int s = 0; s < set->S; ++s) sums[s] = 0.0;
		return 0;
	}
	int ld = 0;
	if (stage_frames(src, n, dim, &ld)) return 1;
	return score_staged(set->packed, n, ld, dim, sums, nullptr);
}
int srb200_gmmset_score_f64(srb200_gmmset *set, const double *X, int n, int dim, double *sums) {
	API_ENTER();
	RowSrc src;
	src.c64 = X;
	return gmmset_score_host(set, src, n, dim, sums);
}
int srb200_gmmset_score_f32(srb200_gmmset *set, const float *X, int n, int dim, double *sums) {
	API_ENTER();
	RowSrc src;
	src.c32 = X;
	return gmmset_score_host(set, src, n, dim, sums);
}
int srb200_gmmset_score_dev(srb200_gmmset *set, const float *dX, int ld, int dim, const int *utt_off, int n_utt, double *d_sums,
                            float *d_frame_ll, int ldf, int *d_best, void *stream) {
	API_ENTER();
	SRB_REQUIRE(set && dX && utt_off && d_sums, "gmmset_score_dev: null argument");
	return launch_score(set->packed, dX, ld, dim, utt_off, n_utt, d_sums, d_frame_ll, ldf, d_best, ctx().compat_clamp, 0, (cudaStream_t)stream);
}

// ---- MFCC ----
srb200_mfcc *srb200_mfcc_create(int fs, double win_length_ms, double win_shift_ms, int fft_size, int n_filters, int n_ceps,
                                double pre_emphasis_coef) {
	API_ENTER();
	MfccPlan *p = mfcc_plan_create(fs, win_length_ms, win_shift_ms, fft_size, n_filters, n_ceps, pre_emphasis_coef);
	if (!p) return nullptr;
	srb200_mfcc *h = new srb200_mfcc();
	h->plan = p;
	return h;
}
void srb200_mfcc_free(srb200_mfcc *h) {
	API_ENTER();
	if (h) {
		mfcc_plan_free(h->plan);
		delete h;
	}
}
int srb200_mfcc_frame_len(srb200_mfcc *h) { return h ? mfcc_frame_len(h->plan) : 0; }
int srb200_mfcc_frame_shift(srb200_mfcc *h) { return h ? mfcc_frame_shift(h->plan) : 0; }
int srb200_mfcc_num_frames(srb200_mfcc *h, long long n_samples) {
	if (!h) return 0;
	const long long L = mfcc_frame_len(h->plan), S = mfcc_frame_shift(h->plan);
	if (n_samples < L) return 0;
	return (int)((n_samples - L) / S + 1);
}
int srb200_mfcc_out_dim(srb200_mfcc *h, int nd) { return h ? mfcc_n_ceps(h->plan) * (nd + 1) : 0; }

long long srb200_mfcc_rows_total(srb200_mfcc *h, const long long *smp_off, int n_utt, int nd) {
	if (!h || !smp_off) return -1;
	long long rows = 0;
	for (int u = 0; u < n_utt; ++u) {
		const int T = srb200_mfcc_num_frames(h, smp_off[u + 1] - smp_off[u]);
		if (T - nd < 1) return -1;
		rows += T - nd;
	}
	return rows;
}

static int mfcc_extract_host(srb200_mfcc *h, const void *pcm, int kind /*0 i16, 1 f32, 2 f64*/, long long n, int nd, int cmvn, double *out,
                             int *rows_out) {
	SRB_REQUIRE(h && pcm && out, "mfcc_extract: null argument");
	SRB_REQUIRE(n > 0 && n < (1LL << 31), "mfcc_extract: bad sample count %lld", n);
	Context &c = ctx();
	if (c.init()) return 1;
	const size_t esz = kind == 0 ? 2 : 4;
	if (c.h_pcm.ensure((size_t)n * esz) || c.d_pcm.ensure((size_t)n * esz)) return 1;
	if (kind == 2) {
		float *dst = c.h_pcm.as<float>();
		const double *srcp = (const double *)pcm;
		for (long long i = 0; i < n; ++i) dst[i] = (float)srcp[i];
	} else
		memcpy(c.h_pcm.p, pcm, (size_t)n * esz);
	SRB_CUDA(cudaMemcpyAsync(c.d_pcm.p, c.h_pcm.p, (size_t)n * esz, cudaMemcpyHostToDevice, c.stream));
	const int T = srb200_mfcc_num_frames(h, n);
	const int rows = T - nd;
	const int cols = mfcc_n_ceps(h->plan) * (nd + 1);
	const int ld = (cols + 3) & ~3;
	const long long smp_off[2] = {0, n};
	int feat_off[2];
	if (c.d_feat.ensure((size_t)std::max(rows, 1) * ld * sizeof(float))) return 1;
	if (launch_mfcc(h->plan, c.d_pcm.p, kind != 0, smp_off, 1, nd, cmvn, c.d_feat.as<float>(), ld, feat_off, c.stream)) return 1;
	if (c.h_out.ensure((size_t)rows * ld * sizeof(float))) return 1;
	SRB_CUDA(cudaMemcpyAsync(c.h_out.p, c.d_feat.p, (size_t)rows * ld * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
	SRB_CUDA(cudaStreamSynchronize(c.stream));
	const float *src = c.h_out.as<float>();
	for (int r = 0; r < rows; ++r)
		for (int j = 0; j < cols; ++j) out[(size_t)r * cols + j] = (double)src[(size_t)r * ld + j];
	if (rows_out) *rows_out = rows;
	return 0;
}
int srb200_mfcc_extract_i16(srb200_mfcc *h, const int16_t *pcm, long long n, int nd, int cmvn, double *out, int *rows_out) {
	API_ENTER();
	return mfcc_extract_host(h, pcm, 0, n, nd, cmvn, out, rows_out);
}
int srb200_mfcc_extract_f32(srb200_mfcc *h, const float *pcm, long long n, int nd, int cmvn, double *out, int *rows_out) {
	API_ENTER();
	return mfcc_extract_host(h, pcm, 1, n, nd, cmvn, out, rows_out);
}
int srb200_mfcc_extract_f64(srb200_mfcc *h, const double *pcm, long long n, int nd, int cmvn, double *out, int *rows_out) {
	API_ENTER();
	return mfcc_extract_host(h, pcm, 2, n, nd, cmvn, out, rows_out);
}
int srb200_mfcc_extract_dev(srb200_mfcc *h, const int16_t *d_pcm, const long long *smp_off, int n_utt, int nd, float *d_feat, int ld,
                            int *feat_off, void *stream) {
	API_ENTER();
	SRB_REQUIRE(h && d_pcm && smp_off && d_feat, "mfcc_extract_dev: null argument");
	return launch_mfcc(h->plan, d_pcm, 0, smp_off, n_utt, nd, 1, d_feat, ld, feat_off, (cudaStream_t)stream);
}


// ---- LPC ----
static int lpc_extract_host(srb200_mfcc *h, const void *pcm, int kind, long long n, int n_lpc, double *out, int *rows_out) {
	SRB_REQUIRE(h && pcm && out, "lpc_extract: null argument");
	SRB_REQUIRE(n > 0 && n < (1LL << 31), "lpc_extract: bad sample count %lld", n);
	Context &c = ctx();
	if (c.init()) return 1;
	const size_t esz = kind == 0 ? 2 : 4;
	if (c.h_pcm.ensure((size_t)n * esz) || c.d_pcm.ensure((size_t)n * esz)) return 1;
	if (kind == 2) {
		float *dst = c.h_pcm.as<float>();
		const double *srcp = (const double *)pcm;
		for (long long i = 0; i < n; ++i) dst[i] = (float)srcp[i];
	} else
		memcpy(c.h_pcm.p, pcm, (size_t)n * esz);
	SRB_CUDA(cudaMemcpyAsync(c.d_pcm.p, c.h_pcm.p, (size_t)n * esz, cudaMemcpyHostToDevice, c.stream));
	const int rows = srb200_mfcc_num_frames(h, n);
	SRB_REQUIRE(rows >= 1, "lpc_extract: signal shorter than one frame");
	const int ld = (n_lpc + 3) & ~3;
	const long long smp_off[2] = {0, n};
	int feat_off[2];
	if (c.d_feat.ensure((size_t)rows * ld * sizeof(float))) return 1;
	SRB_CUDA(cudaMemsetAsync(c.d_feat.p, 0, (size_t)rows * ld * sizeof(float), c.stream));
	if (launch_lpc_for_plan(h->plan, c.d_pcm.p, kind != 0, smp_off, 1, n_lpc, c.d_feat.as<float>(), ld, 0, feat_off, c.stream)) return 1;
	if (c.h_out.ensure((size_t)rows * ld * sizeof(float))) return 1;
	SRB_CUDA(cudaMemcpyAsync(c.h_out.p, c.d_feat.p, (size_t)rows * ld * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
	SRB_CUDA(cudaStreamSynchronize(c.stream));
	const float *src = c.h_out.as<float>();
	for (int r = 0; r < rows; ++r)
		for (int j = 0; j < n_lpc; ++j) out[(size_t)r * n_lpc + j] = (double)src[(size_t)r * ld + j];
	if (rows_out) *rows_out = rows;
	return 0;
}
int srb200_lpc_extract_i16(srb200_mfcc *h, const int16_t *pcm, long long n, int n_lpc, double *out, int *rows_out) {
	API_ENTER();
	return lpc_extract_host(h, pcm, 0, n, n_lpc, out, rows_out);
}
int srb200_lpc_extract_f32(srb200_mfcc *h, const float *pcm, long long n, int n_lpc, double *out, int *rows_out) {
	API_ENTER();
	return lpc_extract_host(h, pcm, 1, n, n_lpc, out, rows_out);
}
int srb200_lpc_extract_f64(srb200_mfcc *h, const double *pcm, long long n, int n_lpc, double *out, int *rows_out) {
	API_ENTER();
	return lpc_extract_host(h, pcm, 2, n, n_lpc, out, rows_out);
}
int srb200_lpc_extract_dev(srb200_mfcc *h, const int16_t *d_pcm, const long long *smp_off, int n_utt, int n_lpc, float *d_feat, int ld,
                           int col_off, int *feat_off, void *stream) {
	API_ENTER();
	SRB_REQUIRE(h && d_pcm && smp_off && d_feat, "lpc_extract_dev: null argument");
	return launch_lpc_for_plan(h->plan, d_pcm, 0, smp_off, n_utt, n_lpc, d_feat, ld, col_off, feat_off, (cudaStream_t)stream);
}


// ---- silence removal ----
static int remove_silence_host(int fs, const void *pcm, int is_u8, long long n, double frame_duration, double frame_shift, double perc,
                               void *out, long long *n_out) {
	SRB_REQUIRE(pcm && out && n_out, "remove_silence: null argument");
	SRB_REQUIRE(n >= 1 && n < (1LL << 40), "remove_silence: bad sample count %lld", n);
	Context &c = ctx();
	if (c.init()) return 1;
	const int frame_len = (int)(frame_duration * fs), shift_len = (int)(frame_shift * fs);     // silence.py:23-24
	SRB_REQUIRE(frame_len >= 1 && shift_len >= 1, "remove_silence: frame length/shift below one sample at fs=%d", fs);
	const size_t esz = is_u8 ? 1 : 2;
	const size_t nb = (size_t)((n + 1023) / 1024), nseg = (size_t)(n / shift_len + 2);
	if (c.h_pcm.ensure((size_t)n * esz) || c.d_pcm.ensure((size_t)n * esz) || c.d_out.ensure((size_t)n * esz) ||
	    c.d_misc.ensure(((size_t)n + 1 + nb + 2) * sizeof(long long) + nseg * sizeof(longlong3) + 64))
		return 1;
	memcpy(c.h_pcm.p, pcm, (size_t)n * esz);
	SRB_CUDA(cudaMemcpyAsync(c.d_pcm.p, c.h_pcm.p, (size_t)n * esz, cudaMemcpyHostToDevice, c.stream));
	long long *d_ps = c.d_misc.as<long long>();
	long long *d_block = d_ps + n + 1;
	long long *d_counts = d_block + nb;
	longlong3 *d_segs = reinterpret_cast<longlong3 *>(d_counts + 2);
	if (launch_remove_silence(c.d_pcm.p, is_u8, n, frame_len, shift_len, perc, d_ps, d_block, d_segs, d_counts, c.d_out.p, c.stream)) return 1;
	long long counts[2] = {0, 0};
	SRB_CUDA(cudaMemcpyAsync(counts, d_counts, sizeof counts, cudaMemcpyDeviceToHost, c.stream));
	SRB_CUDA(cudaStreamSynchronize(c.stream));
	if (counts[1] > 0) SRB_CUDA(cudaMemcpy(out, c.d_out.p, (size_t)counts[1] * esz, cudaMemcpyDeviceToHost));
	*n_out = counts[1];
	return 0;
}
int srb200_remove_silence_i16(int fs, const int16_t *pcm, long long n, double frame_duration, double frame_shift, double perc, int16_t *out,
                              long long *n_out) {
	API_ENTER();
	return remove_silence_host(fs, pcm, 0, n, frame_duration, frame_shift, perc, out, n_out);
}
int srb200_remove_silence_u8(int fs, const uint8_t *pcm, long long n, double frame_duration, double frame_shift, double perc, uint8_t *out,
                             long long *n_out) {
	API_ENTER();
	return remove_silence_host(fs, pcm, 1, n, frame_duration, frame_shift, perc, out, n_out);
}

// ---- streaming: slide B device windows by `hop` samples and append the fresh samples (one launch) ----
int srb200_stream_push_i16(const int16_t *d_window_in, const int16_t *d_fresh, int16_t *d_window_out, int n_streams, int window, int hop, void *stream) {
	API_ENTER();
	if (ctx().init()) return 1;
	return launch_ring_shift_append(d_window_in, d_fresh, d_window_out, n_streams, window, hop, (cudaStream_t)stream);
}

}  // extern "C"
