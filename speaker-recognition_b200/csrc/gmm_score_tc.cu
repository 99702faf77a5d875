// gmm_score_tc.cu -- host side of the tensor-core scorer (kernel and layouts: gmm_score_tc.cuh)
#include "gmm_score_tc.cuh"
#include "kernels.h"

namespace srb {

// feature images: per 128-frame tile, PHI = [x^2 (D), x (D), 1, 0-pad] split into TF32 hi / fp32 remainder lo, written in
// the shared-memory order of the MMA's A operand so the scorer loads a tile with one bulk copy.
// One thread per (tile, row, 4-wide k-group): both parts as one float4 each.
__global__ void tc_prep_features_kernel(const float *__restrict__ X, int ld, int D, int Kp, const int4 *__restrict__ ftiles, int n_ftiles,
                                        float *__restrict__ Aimg) {
	const int kgs = Kp >> 2;
	const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long total = (long long)n_ftiles * kgs * TC_M;
	if (idx >= total) return;
	const int r = (int)(idx % TC_M);
	const int kg = (int)((idx / TC_M) % kgs);
	const int tile = (int)(idx / ((long long)TC_M * kgs));
	const int4 ft = ftiles[tile];
	float hi[4] = {0.f, 0.f, 0.f, 0.f}, lo[4] = {0.f, 0.f, 0.f, 0.f};
	if (r < ft.y) {
		const float *row = X + (size_t)(ft.x + r) * ld;
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const int k = 4 * kg + j;
			float v = 0.f;
			if (k < D) {
				const float x = __ldg(row + k);
				v = x * x;
			} else if (k < 2 * D) v = __ldg(row + k - D);
			else if (k == 2 * D) v = 1.f;
			const float h = __uint_as_float(__float_as_uint(v) & ~0x1FFFu);
			hi[j] = h;
			lo[j] = v - h;
		}
	}
	float *base = Aimg + (size_t)tile * tc_a_tile_floats(Kp);
	const size_t off = (size_t)kg * (TC_M / 8) * 32 + (size_t)(r >> 3) * 32 + (r & 7) * 4;
	*reinterpret_cast<float4 *>(base + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
	*reinterpret_cast<float4 *>(base + (size_t)TC_M * Kp + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

static bool g_tc_attr = false;

size_t tc_smem_bytes(int Kp) {
	// operand tiles + mbarriers / TMEM slot word (128 B) + staged hard entries [8][TC_HHALF + 1] int4 + pair-merge buffer [128] float2
	return (tc_a_tile_floats(Kp) + (size_t)TC_BSTAGES * TC_NB * Kp) * sizeof(float) + 128 + 8 * (TC_HHALF + 1) * 16 + TC_M * 8 + 64;
}

bool tc_supported(const PackedModels &pm, int D) {
	return !pm.h_tc.empty() && pm.tc_Kp == ((2 * D + 1 + 7) / 8) * 8 && tc_smem_bytes(pm.tc_Kp) <= 227 * 1024 - 1024;
}

// d_ftiles: frame tiles of TC_M frames (device), d_Aimg: workspace of n_ftiles * tc_a_tile_floats(Kp) floats,
// d_partial: [n_ftiles][S][4] doubles
int launch_score_tc(PackedModels &pm, const float *dX, int ld, int D, const int4 *d_ftiles, int n_ftiles, float *d_Aimg, double *d_partial,
                    float *d_frame_ll, int ldf, int clamp, int frame_log2, cudaStream_t st) {
	Context &c = ctx();
	const int Kp = pm.tc_Kp;
	if (pm.d_tc == nullptr) {
		SRB_CUDA(cudaMalloc(&pm.d_tc, pm.h_tc.size() * sizeof(float)));
		SRB_CUDA(cudaMemcpyAsync(pm.d_tc, pm.h_tc.data(), pm.h_tc.size() * sizeof(float), cudaMemcpyHostToDevice, st));
		SRB_CUDA(cudaMalloc(&pm.d_tc_hard, pm.h_tc_hard.size() * sizeof(int)));
		SRB_CUDA(cudaMemcpyAsync(pm.d_tc_hard, pm.h_tc_hard.data(), pm.h_tc_hard.size() * sizeof(int), cudaMemcpyHostToDevice, st));
		SRB_CUDA(cudaStreamSynchronize(st));
	}
	if (!g_tc_attr) {
		SRB_CUDA(cudaFuncSetAttribute(gmm_score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		g_tc_attr = true;
	}
	if (n_ftiles <= 0) return 0;
	{
		const long long total = (long long)n_ftiles * (Kp / 4) * TC_M;
		tc_prep_features_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dX, ld, D, Kp, d_ftiles, n_ftiles, d_Aimg);
		SRB_CUDA(cudaGetLastError());
	}
	TcArgs a;
	a.A = d_Aimg;
	a.B = pm.d_tc;
	a.ftiles = d_ftiles;
	a.n_ftiles = n_ftiles;
	a.Kp = Kp;
	a.S = pm.S;
	a.seg = pm.tc_cols_per_spk;
	a.blocks_per_spk = pm.tc_cols_per_spk >= TC_NB ? pm.tc_cols_per_spk / TC_NB : 1;
	a.n_blocks_total = pm.tc_blocks_total;
	// chunking of the speaker axis: one chunk when there are enough frame tiles to fill the SMs, else split whole speakers
	int n_chunks = 1;
	if (n_ftiles < c.sm_count) n_chunks = std::min(pm.S, ceil_div(c.sm_count, n_ftiles));
	int spc = ceil_div(pm.S, n_chunks);
	if (pm.tc_cols_per_spk < TC_NB) {                        // small speakers: chunks start on block boundaries
		const int per_block = TC_NB / pm.tc_cols_per_spk;
		spc = ceil_div(spc, per_block) * per_block;
	}
	n_chunks = ceil_div(pm.S, spc);
	a.n_chunks = n_chunks;
	a.spk_per_chunk = spc;
	a.frame_ll = d_frame_ll;
	a.ldf = ldf;
	a.partial = d_partial;
	a.clamp = clamp;
	a.frame_log2 = frame_log2;
	a.D = D;
	a.hard = reinterpret_cast<const int4 *>(pm.d_tc_hard);
	const long long units = (long long)n_ftiles * n_chunks;
	const int grid = (int)std::min<long long>(units, c.sm_count);
	gmm_score_tc_kernel<<<grid, TC_THREADS, tc_smem_bytes(Kp), st>>>(a);
	SRB_CUDA(cudaGetLastError());
	return 0;
}

}  // namespace srb
