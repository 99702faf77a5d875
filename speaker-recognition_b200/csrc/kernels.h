// kernels.h -- launch entry points of the CUDA translation units (internal; the public surface is include/srb200.h)
#pragma once
#include "gmm_model.h"

namespace srb {

// ---- gmm_score.cu ----
// dX [n_total][ld] fp32 on device; utt_off[n_utt+1] HOST frame offsets; d_sums [n_utt][S] float64;
// d_frame_ll optional [S][ldf]; d_best optional [n_utt].
int launch_score(PackedModels &pm, const float *dX, int ld, int D, const int *utt_off, int n_utt, double *d_sums,
                 float *d_frame_ll, int ldf, int *d_best, int clamp, int frame_log2, cudaStream_t st);
void score_tables_release();

// ---- gmm_score_tc2.cu: tensor-core route (tcgen05 cta_group::2, 3xFP16) ----
bool tc2_supported(const PackedModels &pm, int D);
int launch_score_tc2(PackedModels &pm, const float *dX, int ld, int D, const int4 *d_ftiles, int n_ftiles, double *d_partial, float *d_frame_ll,
                     int ldf, int clamp, int frame_log2, cudaStream_t st);
// the same kernel on operand images that already live on the device (built there by gmm_em_tc.cu from a float64 master)
struct Tc2Images {
	const void *B;             // [blocks_total][2 ranks][head | tail][64 * Kp] fp16
	const int4 *hard;          // [blocks_total][TC_HSTRIDE]
	const float *cs;           // [2][D]
	int Kp, S, cols_per_spk, blocks_total;
};
int launch_score_tc2_images(const Tc2Images &im, const float *dX, int ld, int D, const int4 *d_ftiles, int n_ftiles, double *d_partial,
                            float *d_frame_ll, int ldf, int clamp, int frame_log2, cudaStream_t st);
// frame tiles + launch + per-utterance reduction (gmm_score.cu)
int launch_score_tc_images(const Tc2Images &im, const float *dX, int ld, int D, const int *utt_off, int n_utt, double *d_sums, float *d_frame_ll,
                           int ldf, int *d_best, int clamp, int frame_log2, cudaStream_t st);

// ---- gmm_em.cu ----
// stats (float64): [K] N_k | [K*D] sum g x | [K*D] sum g x^2 | [1] sum log p | [1] frames used
static inline int em_stats_len(int K, int D) { return K + 2 * K * D + 2; }
// E-step = launch_score(frame_log2 = 1) for the per-frame normaliser + this accumulation pass.
// d_lse2 [n] per-frame log2 p(x); d_ll_sum [1] the (clamped) total log-likelihood from the scoring pass.
int launch_estep_accumulate(const PackedModels &pm, const float *dX, int n, int ld, int D, int K, const float *d_lse2,
                            const double *d_ll_sum, double *d_stats, int clamp, cudaStream_t st);
int launch_mstep(double *d_params, const double *d_ubm_params, int K, int D, const double *d_stats, long long n_total,
                 double min_covar, int mode, cudaStream_t st);
int launch_pack_from_master(const double *d_params, int K, int D, PackedModels &pm, cudaStream_t st);
int launch_em_tail(const double *d_ll_sum, const float *d_lse2, int n, int clamp, double *d_stats, int K, int D, cudaStream_t st);

// ---- gmm_em_tc.cu: the E-step of large models on the tensor cores ----
bool em_tc_eligible(int K, int D, int n);
// *used = 1: d_stats holds the E-step's statistics; 0: these parameters are not safe for the uncorrected expanded form, nothing written
int launch_estep_tc(PackedModels &pm, const double *d_params, const float *dX, int n, int ld, int D, int K, float *d_lse2, double *d_ll_sum,
                    double *d_stats, int clamp, cudaStream_t st, int *used);
void em_tc_release();
void em_release();

// ---- mfcc.cu ----
struct MfccPlan;
MfccPlan *mfcc_plan_create(int fs, double win_ms, double shift_ms, int nfft, int n_filters, int n_ceps, double pre_emph);
void mfcc_plan_free(MfccPlan *p);
int mfcc_frame_len(const MfccPlan *p);
int mfcc_frame_shift(const MfccPlan *p);
int mfcc_n_ceps(const MfccPlan *p);
// d_pcm int16 or float (is_f32) on device, smp_off[n_utt+1] HOST sample offsets.  Writes features
// [rows_total][ld] (nd in 0..2, cmvn 0/1) and the HOST row offsets feat_off[n_utt+1].
int launch_mfcc(MfccPlan *p, const void *d_pcm, int is_f32, const long long *smp_off, int n_utt, int nd, int cmvn,
                float *d_feat, int ld, int *feat_off, cudaStream_t st);

// ---- lpc.cu (driven through the MFCC plan, which owns the frame geometry and the float64 window) ----
int launch_lpc(const double *d_win, const double *d_win2, int L, int shift, const void *d_pcm, int is_f32, const long long *d_smp_off,
               const int *d_frame_off, int n_utt, int frames_total, int order, float *d_out, int ld, int col_off, cudaStream_t st);
// features = LPC coefficients 1..order of every frame, written to columns [col_off, col_off+order) of d_out rows
int launch_lpc_for_plan(MfccPlan *p, const void *d_pcm, int is_f32, const long long *smp_off, int n_utt, int order, float *d_out,
                        int ld, int col_off, int *feat_off, cudaStream_t st);

// ---- kmeans.cu: K-MeansII initialisation (gmm.cc:331-344) on staged device features ----
int kmeans2_init_device(const float *dX, int n, int ld, int D, int K, uint64_t seed, double *centroids_out, double *cost_out, int verbosity);
int kmeans_lloyd_device(const float *dX, int n, int ld, int D, int K, const double *h_w, double *h_cent, double *cost_out, int *iters_out);

// ---- silence.cu ----
int launch_remove_silence(const void *d_pcm, int is_u8, long long n, int frame_len, int shift_len, double perc, long long *d_ps,
                          long long *d_block, longlong3 *d_segs, long long *d_counts, void *d_out, cudaStream_t st);

int launch_ring_shift_append(const int16_t *d_src, const int16_t *d_fresh, int16_t *d_dst, int n_streams, int W, int hop, cudaStream_t st);

}  // namespace srb
