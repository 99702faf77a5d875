// fma_peak.cu -- measurement helper for bench.py only (NOT part of libsrb200.so): the fp32 FMA-pipe roofline
// denominator.  MEASURED_PEAKS.json carries HBM and bf16-tensor peaks but no fp32 figure, and the GMM kernels are
// bound by the FMA pipe, so the peak is measured on the box with register-only FMA loops:
//   variant 0: independent FFMA chains (16 accumulators per thread)
//   variant 1: the scorer's dependent pair  t = fma(x, a, b); acc = fma(t, t, acc)   (16 pairs per thread)
//   variant 2: packed fma.rn.f32x2 (FFMA2, sm_100+) independent chains (8 x 2 accumulators per thread)
#include <cuda_runtime.h>
#include <stdint.h>

template <int VARIANT>
__global__ void __launch_bounds__(256) fma_loop(float *out, int iters, float a, float b) {
	float acc[16];
#pragma unroll
	for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i) * 1e-3f;
	if (VARIANT == 0) {
		for (int it = 0; it < iters; ++it) {
#pragma unroll
			for (int r = 0; r < 8; ++r)
#pragma unroll
				for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
		}
	} else if (VARIANT == 1) {
		// x changes every repetition (it depends on the accumulators), so t cannot be hoisted out of the loop
		for (int it = 0; it < iters; ++it) {
#pragma unroll
			for (int r = 0; r < 4; ++r)
#pragma unroll
				for (int i = 0; i < 16; ++i) {
					const float t = fmaf(acc[(i + 5) & 15], a, b);
					acc[i] = fmaf(t, t, acc[i]);
				}
		}
	} else {
		unsigned long long p[8], pa, pb;
		asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a), "f"(a));
		asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
#pragma unroll
		for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(acc[2 * i]), "f"(acc[2 * i + 1]));
		for (int it = 0; it < iters; ++it) {
#pragma unroll
			for (int r = 0; r < 8; ++r)
#pragma unroll
				for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
		}
#pragma unroll
		for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * i]), "=f"(acc[2 * i + 1]) : "l"(p[i]));
	}
	float s = 0.f;
#pragma unroll
	for (int i = 0; i < 16; ++i) s += acc[i];
	if (s == 123.456f) out[0] = s;   // keep the loop alive
}

// returns achieved TFLOP/s (2 flops per FMA lane-op) for the variant, best of `reps`; <0 on error
extern "C" double srb200_bench_fma_tflops(int variant, int iters, int reps, double *ms_out) {
	int dev = 0, sms = 0;
	if (cudaGetDevice(&dev) != cudaSuccess) return -1.0;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	float *d = nullptr;
	if (cudaMalloc(&d, 64) != cudaSuccess) return -1.0;
	const int blocks = sms * 8, threads = 256;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	double best_ms = 1e30;
	for (int r = 0; r < reps + 1; ++r) {
		cudaEventRecord(e0);
		if (variant == 0) fma_loop<0><<<blocks, threads>>>(d, iters, 1.0001f, 1e-7f);
		else if (variant == 1) fma_loop<1><<<blocks, threads>>>(d, iters, 1.0001f, 1e-7f);
		else fma_loop<2><<<blocks, threads>>>(d, iters, 1.0001f, 1e-7f);
		cudaEventRecord(e1);
		if (cudaEventSynchronize(e1) != cudaSuccess) return -1.0;
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		if (r > 0 && ms < best_ms) best_ms = ms;
	}
	cudaFree(d);
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	// lane-FMAs per thread per iteration: v0 8*16 = 128, v1 4*16*2 = 128, v2 8*8*2 = 128
	const double fmas = (double)blocks * threads * (double)iters * 128.0;
	if (ms_out) *ms_out = best_ms;
	return 2.0 * fmas / (best_ms * 1e-3) / 1e12;
}
