// gmm_score_tc.cu -- host side of the tensor-core scorer (kernel and layouts: gmm_score_tc.cuh)
#include "gmm_score_tc.cuh"
#include "kernels.h"

namespace srb {

static bool g_tc_attr = false;

static size_t tc_smem_bytes_n(int Kp, int D, int nst) {
	// B part stages + hard-entry ring + add-back scratch [16][32][32] + merge buffers [2][3][128] float2 + mbarriers / TMEM slot word
	// (256 B) + raw feature tile [128][D|1]
	return (size_t)nst * TC_NB * Kp * sizeof(float) + (size_t)TC_HSTAGES * TC_HSTRIDE * 16 + 16 * 32 * 32 * sizeof(float) + 2 * 3 * TC_M * 8 + 256 +
	       (size_t)TC_M * tc_x_stride(D) * sizeof(float);
}

// B ring depth that fits: 3 stages up to Kp = 80 (D <= 39), 2 stages up to Kp = 120 (D <= 59); 0 = the shape stays on the FMA route
static int tc_stages(int Kp, int D) {
	for (int nst = TC_BSTAGES; nst >= 2; --nst)
		if (tc_smem_bytes_n(Kp, D, nst) <= 227 * 1024) return nst;
	return 0;
}

bool tc_supported(const PackedModels &pm, int D) {
	return !pm.h_tc.empty() && pm.tc_Kp == ((2 * D + 1 + 7) / 8) * 8 && pm.tc_Kp <= TC_KP_MAX && tc_stages(pm.tc_Kp, D) > 0;
}

// d_ftiles: frame tiles of TC_M frames (device), d_partial: [n_ftiles][S][4] doubles
int launch_score_tc(PackedModels &pm, const float *dX, int ld, int D, const int4 *d_ftiles, int n_ftiles, double *d_partial, float *d_frame_ll,
                    int ldf, int clamp, int frame_log2, cudaStream_t st) {
	Context &c = ctx();
	const int Kp = pm.tc_Kp;
	if (pm.d_tc == nullptr) {
		SRB_CUDA(cudaMalloc(&pm.d_tc, pm.h_tc.size() * sizeof(float)));
		SRB_CUDA(cudaMemcpyAsync(pm.d_tc, pm.h_tc.data(), pm.h_tc.size() * sizeof(float), cudaMemcpyHostToDevice, st));
		SRB_CUDA(cudaMalloc(&pm.d_tc_hard, pm.h_tc_hard.size() * sizeof(int)));
		SRB_CUDA(cudaMemcpyAsync(pm.d_tc_hard, pm.h_tc_hard.data(), pm.h_tc_hard.size() * sizeof(int), cudaMemcpyHostToDevice, st));
		SRB_CUDA(cudaMalloc(&pm.d_tc_cs, pm.h_tc_cs.size() * sizeof(float)));
		SRB_CUDA(cudaMemcpyAsync(pm.d_tc_cs, pm.h_tc_cs.data(), pm.h_tc_cs.size() * sizeof(float), cudaMemcpyHostToDevice, st));
		SRB_CUDA(cudaStreamSynchronize(st));
	}
	if (!g_tc_attr) {
		SRB_CUDA(cudaFuncSetAttribute(gmm_score_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		SRB_CUDA(cudaFuncSetAttribute(gmm_score_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		SRB_CUDA(cudaFuncSetAttribute(gmm_score_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		SRB_CUDA(cudaFuncSetAttribute(gmm_score_tc_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		g_tc_attr = true;
	}
	if (n_ftiles <= 0) return 0;
	TcArgs a;
	a.X = dX;
	a.ld = ld;
	a.B = pm.d_tc;
	a.cs = pm.d_tc_cs;
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
	a.bstages = tc_stages(Kp, D);
	a.hard = reinterpret_cast<const int4 *>(pm.d_tc_hard);
	const long long units = (long long)n_ftiles * n_chunks;
	const int grid = (int)std::min<long long>(units, c.sm_count);
	const size_t smem = tc_smem_bytes_n(Kp, D, a.bstages);
	switch (Kp / 8) {      // D = 13 -> 4, D = 28 (MFCC + LPC) -> 8, D = 39 -> 10
		case 4: gmm_score_tc_kernel<4><<<grid, TC_THREADS, smem, st>>>(a); break;
		case 8: gmm_score_tc_kernel<8><<<grid, TC_THREADS, smem, st>>>(a); break;
		case 10: gmm_score_tc_kernel<10><<<grid, TC_THREADS, smem, st>>>(a); break;
		default: gmm_score_tc_kernel<0><<<grid, TC_THREADS, smem, st>>>(a); break;
	}
	SRB_CUDA(cudaGetLastError());
	return 0;
}

}  // namespace srb
