// gmm_score_tc2.cu -- host side of the second-generation tensor-core scorer (kernel and layouts: gmm_score_tc2.cuh)
#include "gmm_score_tc2.cuh"
#include "kernels.h"

namespace srb {

static bool g_tc2_attr = false;

static size_t tc2_smem_bytes_n(int Kp, int D, int nst) {
	// half-block stages (head | tail: 256 Kp bytes) + hard-entry ring + add-back scratch [16][32][T2_SCR_STRIDE] + finisher ring
	// [T2_FSLOTS][4][128] float2 + mbarriers / TMEM slot word (512 B) + x'' tile [128][D|1]
	return (size_t)nst * 4 * T2_NH * Kp + (size_t)T2_HSTAGES * TC_HSTRIDE * 16 + 16 * 32 * T2_SCR_STRIDE * sizeof(float) +
	       (size_t)T2_FSLOTS * 4 * TC_M * 8 + 512 + (size_t)TC_M * tc_x_stride(D) * sizeof(float) +
	       (size_t)2 * T2_RED_SPK * T2_RED_STRIDE * sizeof(float);      // + the finishers' row-sum buffers
}

// deepest B ring that fits (a stage is 960 tensor cycles of work at Kp = 80; the ring has to cover the L2 latency); 0 = unsupported
static int tc2_stages(int Kp, int D) {
	for (int nst = T2_BSTAGES_MAX; nst >= 2; --nst)
		if (tc2_smem_bytes_n(Kp, D, nst) <= 227 * 1024) return nst;
	return 0;
}

bool tc2_supported(const PackedModels &pm, int D) {
	return !pm.h_tc2.empty() && pm.tc2_Kp == t2_kp(D) && pm.tc2_Kp <= T2_KP_MAX && tc2_stages(pm.tc2_Kp, D) > 0;
}

// d_ftiles: frame tiles of TC_M frames (device), d_partial: [n_ftiles][S][T2_PARTS] doubles (the caller sizes it for 4 parts: the FMA route)
int launch_score_tc2(PackedModels &pm, const float *dX, int ld, int D, const int4 *d_ftiles, int n_ftiles, double *d_partial, float *d_frame_ll,
                     int ldf, int clamp, int frame_log2, cudaStream_t st) {
	Context &c = ctx();
	const int Kp = pm.tc2_Kp;
	if (pm.d_tc2 == nullptr) {
		SRB_CUDA(cudaMalloc(&pm.d_tc2, pm.h_tc2.size() * sizeof(uint16_t)));
		SRB_CUDA(cudaMemcpyAsync(pm.d_tc2, pm.h_tc2.data(), pm.h_tc2.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
		if (pm.d_tc_hard == nullptr) {
			SRB_CUDA(cudaMalloc(&pm.d_tc_hard, pm.h_tc_hard.size() * sizeof(int)));
			SRB_CUDA(cudaMemcpyAsync(pm.d_tc_hard, pm.h_tc_hard.data(), pm.h_tc_hard.size() * sizeof(int), cudaMemcpyHostToDevice, st));
			SRB_CUDA(cudaMalloc(&pm.d_tc_cs, pm.h_tc_cs.size() * sizeof(float)));
			SRB_CUDA(cudaMemcpyAsync(pm.d_tc_cs, pm.h_tc_cs.data(), pm.h_tc_cs.size() * sizeof(float), cudaMemcpyHostToDevice, st));
		}
		SRB_CUDA(cudaStreamSynchronize(st));
	}
	Tc2Images im;
	im.B = pm.d_tc2;
	im.hard = reinterpret_cast<const int4 *>(pm.d_tc_hard);
	im.cs = pm.d_tc_cs;
	im.Kp = Kp;
	im.S = pm.S;
	im.cols_per_spk = pm.tc_cols_per_spk;
	im.blocks_total = pm.tc_blocks_total;
	return launch_score_tc2_images(im, dX, ld, D, d_ftiles, n_ftiles, d_partial, d_frame_ll, ldf, clamp, frame_log2, st);
}

int launch_score_tc2_images(const Tc2Images &pm, const float *dX, int ld, int D, const int4 *d_ftiles, int n_ftiles, double *d_partial,
                            float *d_frame_ll, int ldf, int clamp, int frame_log2, cudaStream_t st) {
	Context &c = ctx();
	const int Kp = pm.Kp;
	if (!g_tc2_attr) {
		SRB_CUDA(cudaFuncSetAttribute(gmm_score_tc2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		SRB_CUDA(cudaFuncSetAttribute(gmm_score_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		SRB_CUDA(cudaFuncSetAttribute(gmm_score_tc2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		SRB_CUDA(cudaFuncSetAttribute(gmm_score_tc2_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		g_tc2_attr = true;
	}
	if (n_ftiles <= 0) return 0;
	Tc2Args a;
	a.X = dX;
	a.ld = ld;
	a.cs = pm.cs;
	a.B = reinterpret_cast<const __half *>(pm.B);
	a.ftiles = d_ftiles;
	a.n_ftiles = n_ftiles;
	a.Kp = Kp;
	a.S = pm.S;
	a.seg = pm.cols_per_spk;
	a.blocks_per_spk = pm.cols_per_spk >= TC_NB ? pm.cols_per_spk / TC_NB : 1;
	a.n_blocks_total = pm.blocks_total;
	// a work unit is a PAIR of frame tiles x a chunk of speakers; one chunk when the pairs alone fill the CTA pairs of the chip
	const int n_pairs = (n_ftiles + 1) / 2, n_clusters_max = std::max(1, c.sm_count / 2);
	int n_chunks = 1;
	if (n_pairs < n_clusters_max) n_chunks = std::min(pm.S, ceil_div(n_clusters_max, n_pairs));
	int spc = ceil_div(pm.S, n_chunks);
	if (pm.cols_per_spk < TC_NB) {                        // small speakers: chunks start on block boundaries
		const int per_block = TC_NB / pm.cols_per_spk;
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
	a.bstages = tc2_stages(Kp, D);
	a.hard = pm.hard;
	const long long units = (long long)n_pairs * n_chunks;
	const int grid = 2 * (int)std::min<long long>(units, n_clusters_max);      // cluster dims (2,1,1) are compiled into the kernel
	const size_t smem = tc2_smem_bytes_n(Kp, D, a.bstages);
	switch (Kp / 16) {      // D = 13 -> 2, D = 28 (MFCC + LPC) -> 4, D = 34 / 39 -> 5
		case 2: gmm_score_tc2_kernel<2><<<grid, T2_THREADS, smem, st>>>(a); break;
		case 4: gmm_score_tc2_kernel<4><<<grid, T2_THREADS, smem, st>>>(a); break;
		case 5: gmm_score_tc2_kernel<5><<<grid, T2_THREADS, smem, st>>>(a); break;
		default: gmm_score_tc2_kernel<0><<<grid, T2_THREADS, smem, st>>>(a); break;
	}
	SRB_CUDA(cudaGetLastError());
	return 0;
}

}  // namespace srb

#if SRB_T2_PROF
extern "C" int srb200_debug_tc2_prof(unsigned long long *out) {      // experiments only: read and clear the cycle counters
	if (cudaDeviceSynchronize() != cudaSuccess) return 1;
	if (cudaMemcpyFromSymbol(out, srb::g_tc2_prof, sizeof(srb::g_tc2_prof)) != cudaSuccess) return 1;
	static unsigned long long zero[2 * 20 * 8];
	return cudaMemcpyToSymbol(srb::g_tc2_prof, zero, sizeof(zero)) != cudaSuccess;
}
#endif
