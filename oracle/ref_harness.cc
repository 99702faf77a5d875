/*
 * oracle/ref_harness.cc -- TEST INFRASTRUCTURE ONLY.  Never linked into the product.
 *
 * A thin extern "C" shim of OUR OWN that is compiled together with the unmodified reference
 * sources (see oracle/Makefile) so that the parity tests can drive the reference's C++ GMM
 * deterministically:
 *   - inject / read back parameters (the reference's C ABI, src/gmm/src/pygmm.hh:28-41, offers
 *     neither; its training init draws from an RNG seeded by rand(), random.hh:21-25),
 *   - run exactly ONE GMMTrainerBaseline::iteration (gmm.cc:439-531) or one MAP iteration
 *     (GMMUBMTrainerBaseline, gmmubm.cc:40-81) from an injected start,
 *   - score contiguous double matrices without the per-row ctypes marshalling,
 *   - run the reference's Lloyd iterations from injected centroids (KMeansSolver::Lloyds_iteration[_weighted],
 *     kmeans.cc:150-340) and its whole K-MeansII initialisation after srand(seed) (kmeansII.cc:80-172).
 * It includes the reference headers from where they lie; it copies no reference code.
 */
#include "gmm.hh"
#include "kmeansII.hh"
#include "kmeans++.hh"

#include <vector>
#include <cstring>

typedef std::vector<std::vector<real_t>> Dense;

static void to_dense(const double *X, int n, int dim, Dense &out) {
	out.resize(n);
	for (int i = 0; i < n; i++)
		out[i].assign(X + (size_t)i * dim, X + (size_t)(i + 1) * dim);
}

/* Lay the trainer's work buffers out the way GMMTrainerBaseline::train does (gmm.cc:588-600)
 * but WITHOUT clear_gaussians/init_gaussians, so the injected parameters are the start point. */
static void prime_trainer(GMMTrainerBaseline &t, GMM *gmm, Dense &X) {
	t.gmm = gmm;
	gmm->dim = t.dim = (int)X[0].size();
	t.prob_of_y_given_x.resize(gmm->nr_mixtures);
	for (auto &v : t.prob_of_y_given_x)
		v.resize(X.size());
	t.prob_sum.resize(X.size());
	t.N_k.resize(gmm->nr_mixtures);
}

extern "C" {

GMM *ref_new(int K, int D, const double *w, const double *mean, const double *sigma) {
	GMM *g = new GMM(K, COVTYPE_DIAGONAL);
	g->dim = D;
	g->weights.assign(w, w + K);
	g->gaussians.resize(K);
	for (int k = 0; k < K; k++) {
		Gaussian *ga = new Gaussian(D, COVTYPE_DIAGONAL);
		ga->mean.assign(mean + (size_t)k * D, mean + (size_t)(k + 1) * D);
		ga->sigma.assign(sigma + (size_t)k * D, sigma + (size_t)(k + 1) * D);
		g->gaussians[k] = ga;
	}
	return g;
}

void ref_get(GMM *g, double *w, double *mean, double *sigma) {
	int K = g->nr_mixtures, D = g->dim;
	for (int k = 0; k < K; k++) {
		w[k] = g->weights[k];
		for (int d = 0; d < D; d++) {
			mean[(size_t)k * D + d] = g->gaussians[k]->mean[d];
			sigma[(size_t)k * D + d] = g->gaussians[k]->sigma[d];
		}
	}
}

void ref_free(GMM *g) { delete g; }

/* one full EM iteration from the current parameters (gmm.cc:439-531) */
void ref_em_iteration(GMM *g, const double *X, int n, int dim, double min_covar, int concurrency) {
	Dense Xd;
	to_dense(X, n, dim, Xd);
	GMMTrainerBaseline t(1, min_covar, 0.0, 0, concurrency, 0);
	prime_trainer(t, g, Xd);
	t.iteration(Xd);
}

/* one MAP (means-only) iteration; g must already hold a copy of the UBM parameters, as
 * GMMUBMTrainerBaseline::init_gaussians leaves it (gmmubm.cc:25-38). */
void ref_map_iteration(GMM *g, GMM *ubm, const double *X, int n, int dim, int concurrency) {
	Dense Xd;
	to_dense(X, n, dim, Xd);
	GMMUBMTrainerBaseline t(ubm, 1, 1e-3, 0.0, concurrency, 0);
	prime_trainer(t, g, Xd);
	t.iteration(Xd);
}

/* the reference's production scorer (fast-exp, threaded; gmm.cc:533-578) on a contiguous matrix */
void ref_score_batch(GMM *g, const double *X, double *out, int n, int dim, int concurrency) {
	Dense Xd;
	to_dense(X, n, dim, Xd);
	std::vector<real_t> prob;
	g->log_probability_of_fast_exp_threaded(Xd, prob, concurrency);
	memcpy(out, prob.data(), sizeof(double) * n);
}

/* the reference's libm-exp scorer (gmm.cc:229-235, 153-174), single thread: brackets the
 * fast-exp error of the production scorer */
void ref_score_batch_exactexp(GMM *g, const double *X, double *out, int n, int dim) {
	std::vector<real_t> x(dim);
	for (int i = 0; i < n; i++) {
		x.assign(X + (size_t)i * dim, X + (size_t)(i + 1) * dim);
		out[i] = g->log_probability_of(x);
	}
}

int ref_dim(GMM *g) { return g->dim; }
int ref_nr_mixtures(GMM *g) { return g->nr_mixtures; }


static void to_dataset(const double *X, int n, int dim, Dataset &ds) {
	ds.resize(n);
	for (int i = 0; i < n; i++) {
		ds[i].resize(dim);
		for (int d = 0; d < dim; d++) ds[i][d] = std::make_pair(d, X[(size_t)i * dim + d]);
	}
}

/* Lloyd iterations of the reference from the given centroids (in/out); weight may be NULL.  Returns the best cost. */
double ref_lloyd(const double *X, int n, int dim, int K, const double *weight, double *cent, int concurrency) {
	Dataset ds;
	to_dataset(X, n, dim, ds);
	std::vector<Vector> c(K, Vector(dim));
	for (int k = 0; k < K; k++) c[k].assign(cent + (size_t)k * dim, cent + (size_t)(k + 1) * dim);
	KMeansSolver solver(concurrency);
	double cost;
	if (weight) {
		std::vector<real_t> w(weight, weight + n);
		cost = solver.Lloyds_iteration_weighted(ds, w, c);
	} else {
		cost = solver.Lloyds_iteration(ds, c);
	}
	for (int k = 0; k < K; k++) memcpy(cent + (size_t)k * dim, c[k].data(), sizeof(double) * dim);
	return cost;
}

/* the reference's whole K-MeansII initialisation (what init_gaussians runs, gmm.cc:331-337) after srand(seed) */
double ref_kmeansII(const double *X, int n, int dim, int K, unsigned seed, double *cent, int concurrency) {
	srand(seed);
	Dataset ds;
	to_dataset(X, n, dim, ds);
	std::vector<Vector> c;
	KMeansIISolver solver(concurrency);
	const double cost = solver.cluster(ds, c, K);
	for (int k = 0; k < K; k++) memcpy(cent + (size_t)k * dim, c[k].data(), sizeof(double) * dim);
	return cost;
}


/* the reference's weighted k-means++ stage (KMeansppSolver::cluster_weighted, kmeans++.cc:157-213) on given candidates, after srand(seed) */
double ref_kmeanspp_weighted(const double *Cand, int m, int dim, const double *weight, int K, unsigned seed, double *cent, int concurrency) {
	srand(seed);
	std::vector<Vector> ds(m, Vector(dim));
	for (int i = 0; i < m; i++) ds[i].assign(Cand + (size_t)i * dim, Cand + (size_t)(i + 1) * dim);
	std::vector<real_t> w(weight, weight + m);
	std::vector<Vector> c;
	KMeansppSolver solver(concurrency);
	const double cost = solver.cluster_weighted(ds, w, c, K);
	for (int k = 0; k < K; k++) memcpy(cent + (size_t)k * dim, c[k].data(), sizeof(double) * dim);
	return cost;
}

}
