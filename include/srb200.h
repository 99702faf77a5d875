/*
 * include/srb200.h -- the drop-in C ABI of libsrb200.so (speaker-recognition_b200/csrc).
 *
 * Boundary being replaced: the reference's `lib/pygmm.so` C ABI, src/gmm/src/pygmm.hh:11-43, which
 * src/gmm/python/pygmm.py binds with ctypes (pygmm.py:16,53,64,69,114,117,124,131,136,139).  The ten
 * reference symbols are exported with identical names, argument order and struct layout, so a
 * pygmm.py-style caller works unchanged; every computation behind them runs in sm_100a CUDA kernels
 * (there is no CPU path in this library).  The reference has no native boundary for MFCC
 * (src/feature/MFCC.py is pure Python); the srb200_mfcc_* entry points are what a ctypes stub in
 * MFCC.py would bind instead of its per-frame loop (MFCC.py:59-70).  See INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; device pointers are raw CUDA device addresses (e.g. torch.Tensor.data_ptr()),
 *     streams are cudaStream_t passed as void* (NULL = the library's own stream).
 *   - int-returning functions return 0 on success, non-zero on failure with the message available
 *     from srb200_last_error() (thread-local).  Nothing throws or aborts across the boundary
 *     (the reference throws `const char*` through extern "C", gmm.cc:42-46,582-586).
 *   - host entry points are synchronous; *_dev entry points enqueue on the given stream and return.
 *   - `concurrency` arguments are accepted for signature compatibility and ignored.
 */
#ifndef SRB200_H
#define SRB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct GMM GMM; /* opaque; the reference exposes its C++ class under the same name (pygmm.hh:28) */

/* src/gmm/src/pygmm.hh:12-26 (mirrored by GMMParameter, src/gmm/python/pygmm.py:18-27) */
typedef struct Parameter {
	int nr_instance;
	int nr_dim;
	int nr_mixture;
	double min_covar;
	double threshold;
	int nr_iteration;
	int init_with_kmeans;
	int concurrency;
	int verbosity;
} Parameter;

/* ------------------------------------------------------------------------------------------------
 * The reference's ten symbols -- src/gmm/src/pygmm.hh:28-41, bodies pygmm.cc:50-120
 * ---------------------------------------------------------------------------------------------- */
GMM *new_gmm(int nr_mixture, int covariance_type);                 /* pygmm.hh:28; covariance_type must be 1 (diagonal), gmm.cc:211-215; NULL on error */
GMM *load(const char *model_file);                                 /* pygmm.hh:29; text format gmm.cc:655-682,101-150; NULL on error (the reference silently returns an empty model) */
void dump(GMM *gmm, const char *model_file);                       /* pygmm.hh:31; 6 significant digits like the reference's ostream default */
void train_model(GMM *gmm, double **X_in, Parameter *param);       /* pygmm.hh:33; GMMTrainerBaseline::train gmm.cc:581-653 */
void train_model_from_ubm(GMM *gmm, GMM *ubm, double **X_in, Parameter *param); /* pygmm.hh:34; GMMUBMTrainerBaseline gmmubm.cc:17-81 */
double score_all(GMM *gmm, double **X_in, int nr_instance, int nr_dim, int concurrency);                    /* pygmm.hh:36; gmm.cc:562-573 */
void score_batch(GMM *gmm, double **X_in, double *prob_out, int nr_instance, int nr_dim, int concurrency);  /* pygmm.hh:37; gmm.cc:533-560,575-578 */
double score_instance(GMM *gmm, double *x_in, int nr_dim);         /* pygmm.hh:38; (asserts in the reference, pygmm.cc:116 -> gmm.cc:185; works here) */
int get_dim(GMM *gmm);                                             /* pygmm.hh:40 */
int get_nr_mixtures(GMM *gmm);                                     /* pygmm.hh:41 */

/* ------------------------------------------------------------------------------------------------
 * Additions the reference lacks: ownership, errors, device selection, parameters
 * ---------------------------------------------------------------------------------------------- */
void free_gmm(GMM *gmm);                       /* the reference exports no destructor: every GMM* leaks */
const char *srb200_last_error(void);           /* "" when the last call on this thread succeeded */
const char *srb200_version(void);
int srb200_device_count(void);                 /* 0 when no CUDA device is usable (calls then fail loudly) */
int srb200_set_device(int device);             /* per process: the device all handles of this process live on */
int srb200_get_device(void);

/* score clamp emulation.  0: exact log-domain values.  1 (default): a frame whose log-likelihood is below
 * log(DBL_TRUE_MIN) = -744.44 -- where the reference's linear-domain sum underflows to 0 -- returns
 * log(1e-15) = -34.5388 exactly as safe_log does (gmm.cc:34-38). */
int srb200_set_compat_clamp(int mode);
int srb200_get_compat_clamp(void);

/* scoring arithmetic route.  1: tensor-core kernel (tcgen05, fp32 accumulation in tensor memory) on the expanded-square GEMM
 * form over centred, power-of-two-scaled features, with the few (mixture, dim) entries whose |mean - centre| / sigma would
 * make that form cancel added back exactly in fp32: <= 5e-5 relative per frame against float64 (tests/test_gmm_tc_gpu.py),
 * inside the 1e-4 parity bound, identical argmax.  0: fp32 FMA kernel, 2e-7 relative per frame, about 3x slower.
 * -1 (default): route 1 wherever the model set supports it (D <= 55, add-back list within its capacity) unless the
 * environment variable SRB200_SCORE_TC=0; the per-frame log2 output of the EM E-step always uses route 0.
 * srb200_last_score_route() tells which route the last launch took; srb200_tc_pack_stats() (host only, no GPU needed) tells
 * whether a model set is eligible for route 1 and how many exact add-back entries it needs (returns 1 / 0; < 0 on error). */
int srb200_set_score_path(int path);
int srb200_get_score_path(void);
int srb200_last_score_route(void);
/* route of the most recent EM E-step: 1 = tensor cores (gmm_em_tc.cu), 0 = fp32 FMA kernels (gmm_em.cu), -1 = none yet */
int srb200_last_em_route(void);
/* -1 (default): tensor cores for large jobs (K > 64, n K >= 2^24, 2 D + 1 <= 128), 0: FMA kernels only, 1: tensor cores wherever the shape allows */
int srb200_set_em_path(int path);
long long srb200_score_launch_count(void);          /* scoring launches issued by this process so far */
int srb200_tc_pack_stats(GMM *const *models, int n_models, int *blocks_out, int *hard_entries_out, int *max_hard_per_block_out);
/* the packed operands of route 1 (host only): info_out[8] = {Kp, blocks, columns per speaker, halfs per block, ints per block
 * list, header words, 0, 0}; arrays copied when non-NULL (blocks * info[3] uint16, blocks * info[4] int32, 2 * dim float).
 * Layouts: speaker-recognition_b200/csrc/gmm_score_tc_layout.h.  Used by the CPU emulation test of the packer. */
int srb200_tc_pack_export(GMM *const *models, int n_models, int *info_out, uint16_t *images_out, int *hard_out, float *cs_out);

int srb200_gmm_set_params(GMM *gmm, int nr_mixture, int dim, const double *weights, const double *means, const double *sigmas); /* means,sigmas [K*D]; sigma = std-dev (gmm.hh:31) */
int srb200_gmm_get_params(GMM *gmm, double *weights, double *means, double *sigmas);
int srb200_gmm_dump_precise(GMM *gmm, const char *model_file, int digits);   /* same text format, `digits` significant digits (17 = lossless) */

/* contiguous-matrix variants of score_batch/score_all: X is [n][dim] row-major HOST memory */
int srb200_score_batch_f64(GMM *gmm, const double *X, int n, int dim, double *prob_out);
int srb200_score_all_f64(GMM *gmm, const double *X, int n, int dim, double *sum_out);
int srb200_score_batch_f32(GMM *gmm, const float *X, int n, int dim, float *prob_out);

/* training on a contiguous HOST matrix, with optional injected start (init_* may be NULL -> reference-style
 * init: random data rows, global std, uniform weights, gmm.cc:306-361).  seed drives the row choice.
 * Returns the number of EM iterations run in *iters_out and the last monitored log-likelihood in *ll_out
 * (either may be NULL). */
int srb200_train_f64(GMM *gmm, const double *X, const Parameter *param, uint64_t seed, int *iters_out, double *ll_out);
int srb200_train_from_ubm_f64(GMM *gmm, GMM *ubm, const double *X, const Parameter *param, int *iters_out, double *ll_out);
/* BATCHED MAP enrolment: models[s] becomes the UBM adapted to the frames X[frame_off[s] .. frame_off[s+1]) (rows of `dim` float64),
 * exactly as n_models calls of train_model_from_ubm would (src/gmm/src/gmmubm.cc:17-81 inside the loop of gmm.cc:613-652; what
 * ModelInterface.train does per name, src/gui/interface.py:77-83) -- but with ONE host->device copy of all frames, the speakers'
 * kernels enqueued back to back, and one synchronisation per log-likelihood check for all speakers together.
 * param->nr_instance / nr_mixture are ignored (taken from frame_off / the UBM); iters_out, ll_out: [n_models], may be NULL. */
int srb200_train_many_from_ubm_f64(GMM *const *models, int n_models, GMM *ubm, const double *X, const int *frame_off, int dim, const Parameter *param,
                                   int *iters_out, double *ll_out);

/* K-MeansII initialisation (the reference's init_with_kmeans = 1: gmm.cc:331-344 -> kmeansII.cc:80-172, kmeans++.cc:157-213,
 * kmeans.cc:150-340), also what train_model / srb200_train_f64 run when param->init_with_kmeans != 0.  X is [n][dim] HOST memory,
 * centroids [K][dim].  All draws derive from `seed` (counter-based; the reference uses libc rand()), distances are float64 in the
 * reference's operation order.  srb200_kmeans_lloyd_f64 runs the Lloyd iterations alone from the given centroids (in/out),
 * optionally weighted (weight[n] or NULL), and returns the best cost and the number of iterations. */
int srb200_kmeans_init_f64(const double *X, int n, int dim, int K, uint64_t seed, double *centroids, double *cost_out);
int srb200_kmeans_lloyd_f64(const double *X, int n, int dim, int K, const double *weight, double *centroids, double *cost_out, int *iters_out);

/* ------------------------------------------------------------------------------------------------
 * Speaker set: S models scored in ONE launch (replaces the Python loop of GMMSet.predict_one_scores,
 * src/testbench/gmmset.py:59-64,95-99)
 * ---------------------------------------------------------------------------------------------- */
typedef struct srb200_gmmset srb200_gmmset;
srb200_gmmset *srb200_gmmset_create(GMM *const *models, int n_models);   /* copies parameters; models may be freed afterwards; all must share dim */
void srb200_gmmset_free(srb200_gmmset *set);
int srb200_gmmset_size(srb200_gmmset *set);
int srb200_gmmset_dim(srb200_gmmset *set);
/* only the first n_ranked models take part in the per-utterance argmax (d_best); the others are scored all the same.  Used to carry a
 * UBM as model S + 1 of a speaker set: predict_one_with_rejection (src/testbench/gmmset.py:69-84) becomes one launch. */
int srb200_gmmset_set_ranked(srb200_gmmset *set, int n_ranked);
/* host matrix, one utterance: sums[s] = sum_t log p_s(x_t) */
int srb200_gmmset_score_f64(srb200_gmmset *set, const double *X, int n, int dim, double *sums);
int srb200_gmmset_score_f32(srb200_gmmset *set, const float *X, int n, int dim, double *sums);
/* device matrix dX [n_total][ld] fp32, B utterances delimited by HOST frame offsets utt_off[B+1].
 * d_sums [B][S] float64 (device).  d_frame_ll (optional, may be NULL) [S][ldf] float32 per-frame values,
 * ldf >= n_total.  d_best (optional) [B] int32 argmax_s of d_sums (first maximum, like Python's max). */
int srb200_gmmset_score_dev(srb200_gmmset *set, const float *dX, int ld, int dim, const int *utt_off, int n_utt,
                            double *d_sums, float *d_frame_ll, int ldf, int *d_best, void *stream);

/* ------------------------------------------------------------------------------------------------
 * EM building blocks on DEVICE data (frame-sharded multi-GPU training: each rank calls estep on its
 * shard, all-reduces d_stats with NCCL, then calls mstep)
 *   stats layout (float64): [K] N_k | [K*D] sum gamma x | [K*D] sum gamma x^2 | [1] sum_t log p(x_t) | [1] frames used
 * ---------------------------------------------------------------------------------------------- */
int srb200_em_stats_len(int nr_mixture, int dim);
int srb200_em_estep_dev(GMM *gmm, const float *dX, int n, int ld, int dim, double *d_stats, void *stream);
/* mode 0: full M-step (gmm.cc:388-437) with sigma >= sqrt(min_covar); mode 1: MAP means-only against `ubm`
 * (gmmubm.cc:40-81, relevance 16).  n_total = global frame count (after all-reduce). */
int srb200_em_mstep_dev(GMM *gmm, const double *d_stats, long long n_total, double min_covar, int mode, GMM *ubm, void *stream);
int srb200_gmm_sync_from_device(GMM *gmm);      /* pull the device-resident parameters back into the host copy */

/* ------------------------------------------------------------------------------------------------
 * MFCC front-end (replaces MFCCExtractor.extract, src/feature/MFCC.py:49-79, and diff_feature,
 * src/feature/utils.py:24-31)
 * ---------------------------------------------------------------------------------------------- */
typedef struct srb200_mfcc srb200_mfcc;
/* arguments = get_mfcc_extractor's (MFCC.py:115-121): fs, win_length_ms=32, win_shift_ms=16, FFT_SIZE=2048,
 * n_filters=50, n_ceps=13, pre_emphasis_coef=0.95.  FFT_SIZE must be a power of two in [64, 4096] and
 * >= frame length; n_filters <= 128; n_ceps <= 32, n_ceps < n_filters. */
srb200_mfcc *srb200_mfcc_create(int fs, double win_length_ms, double win_shift_ms, int fft_size, int n_filters, int n_ceps, double pre_emphasis_coef);
void srb200_mfcc_free(srb200_mfcc *plan);
int srb200_mfcc_frame_len(srb200_mfcc *plan);
int srb200_mfcc_frame_shift(srb200_mfcc *plan);
int srb200_mfcc_num_frames(srb200_mfcc *plan, long long n_samples);      /* (n - FRAME_LEN)//FRAME_SHIFT + 1, MFCC.py:57; <=0 if too short */
int srb200_mfcc_out_dim(srb200_mfcc *plan, int nd);                     /* n_ceps * (nd+1) */
/* host PCM, one utterance.  nd = 0: CMVN cepstra [T][n_ceps]; nd = 1|2: diff_feature(.., nd) -> [T-nd][n_ceps*(nd+1)].
 * cmvn = 0 returns the raw (un-normalised) cepstra (nd must be 0).  out is float64 [rows][cols] like the reference.
 * Fails (like the reference's assert, MFCC.py:56) unless n_samples > 5*FRAME_LEN. */
int srb200_mfcc_extract_i16(srb200_mfcc *plan, const int16_t *pcm, long long n_samples, int nd, int cmvn, double *out, int *rows_out);
int srb200_mfcc_extract_f32(srb200_mfcc *plan, const float *pcm, long long n_samples, int nd, int cmvn, double *out, int *rows_out);
int srb200_mfcc_extract_f64(srb200_mfcc *plan, const double *pcm, long long n_samples, int nd, int cmvn, double *out, int *rows_out);
/* device PCM (int16), B utterances delimited by HOST sample offsets smp_off[B+1]; output feature rows are packed
 * utterance after utterance into d_feat [rows_total][ld] fp32 (ld >= out_dim, multiple of 4); feat_off[B+1] (HOST,
 * written by the call) receives the row offsets.  CMVN is per utterance. */
int srb200_mfcc_extract_dev(srb200_mfcc *plan, const int16_t *d_pcm, const long long *smp_off, int n_utt, int nd,
                            float *d_feat, int ld, int *feat_off, void *stream);
long long srb200_mfcc_rows_total(srb200_mfcc *plan, const long long *smp_off, int n_utt, int nd);

/* ------------------------------------------------------------------------------------------------
 * LPC features (replaces LPCExtractor.extract, src/feature/LPC.py:40-57: per frame, coefficients 1..n_lpc of the
 * autocorrelation-method linear predictor, i.e. scikits.talkbox lpc(frame, n_lpc)[0][1:]; same window and
 * pre-emphasis as the MFCC front-end).  The frame geometry (fs, win_length_ms, win_shift_ms, pre-emphasis) is taken
 * from the plan; n_lpc <= 31.  One output row per frame, no normalisation (the reference applies none).
 * ---------------------------------------------------------------------------------------------- */
int srb200_lpc_extract_i16(srb200_mfcc *plan, const int16_t *pcm, long long n_samples, int n_lpc, double *out, int *rows_out);
int srb200_lpc_extract_f32(srb200_mfcc *plan, const float *pcm, long long n_samples, int n_lpc, double *out, int *rows_out);
int srb200_lpc_extract_f64(srb200_mfcc *plan, const double *pcm, long long n_samples, int n_lpc, double *out, int *rows_out);
/* device PCM for B utterances; writes columns [col_off, col_off + n_lpc) of d_feat rows (row = frame, packed utterance
 * after utterance): call after srb200_mfcc_extract_dev(nd = 0) on the same buffer to build mix_feature = MFCC || LPC
 * (src/feature/__init__.py:25-30) without leaving the device. */
int srb200_lpc_extract_dev(srb200_mfcc *plan, const int16_t *d_pcm, const long long *smp_off, int n_utt, int n_lpc,
                           float *d_feat, int ld, int col_off, int *feat_off, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Energy-based silence removal (replaces remove_silence, src/filters/silence.py:11-50).  Integer arithmetic:
 * results are bit-exact with the reference.  frame_duration = 0.02, frame_shift = 0.01, perc = 0.15 are the
 * reference defaults.  out must hold n_samples elements; *n_out receives the kept length.
 * ---------------------------------------------------------------------------------------------- */
int srb200_remove_silence_i16(int fs, const int16_t *pcm, long long n_samples, double frame_duration, double frame_shift, double perc,
                              int16_t *out, long long *n_out);
int srb200_remove_silence_u8(int fs, const uint8_t *pcm, long long n_samples, double frame_duration, double frame_shift, double perc,
                             uint8_t *out, long long *n_out);

/* streaming (BASELINE config 5): d_window_out[b] = [d_window_in[b][hop:], d_fresh[b]] for n_streams windows of `window` int16 samples,
 * 0 < hop < window, one launch (the reference GUI re-slices a Python list per tick, src/gui/gui.py:179-190).  In and out must differ. */
int srb200_stream_push_i16(const int16_t *d_window_in, const int16_t *d_fresh, int16_t *d_window_out, int n_streams, int window, int hop, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SRB200_H */
