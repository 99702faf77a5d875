// silence.cu -- energy-based silence removal on the device.
//
// Replaces remove_silence, src/filters/silence.py:11-50: integer arithmetic on int64 sample squares, so the result is
// BIT-EXACT with the reference:
//   average_energy = sum(signal^2) / float(siglen)                                        (:34)
//   walk i over the signal: frame = signal[i : i+frame_length]; if mean(frame^2) < average_energy * perc the frame is
//   skipped (i += frame_length) else its first frame_shift samples are kept (i += frame_shift)     (:37-46)
// Kernels: (1) exclusive prefix sums of the squared samples in int64 (three-phase block scan) -> any window sum is two
// loads; (2) the walk itself is a data-dependent chain (the next position depends on the decision), run by one thread
// per utterance, emitting (src, dst, len) segments; (3) segment gather.  uint8 input is re-centred like the reference
// (:18-19, :48-49, including its (max+1)/2 vs max/2 asymmetry).
#include <algorithm>
#include "kernels.h"

namespace srb {

constexpr int SCAN_BLOCK = 1024;

// phase 1: per-block inclusive scan of squares, block totals out
__global__ void __launch_bounds__(SCAN_BLOCK) sq_scan_block_kernel(const void *pcm, int is_u8, long long n, long long *ps /*[n+1]*/,
                                                                  long long *block_tot) {
	__shared__ long long sh[SCAN_BLOCK];
	const long long i = (long long)blockIdx.x * SCAN_BLOCK + threadIdx.x;
	long long v = 0;
	if (i < n) {
		const long long s = is_u8 ? (long long)reinterpret_cast<const uint8_t *>(pcm)[i] - 128 : (long long)reinterpret_cast<const int16_t *>(pcm)[i];
		v = s * s;
	}
	sh[threadIdx.x] = v;
	__syncthreads();
	for (int off = 1; off < SCAN_BLOCK; off <<= 1) {
		long long t = 0;
		if ((int)threadIdx.x >= off) t = sh[threadIdx.x - off];
		__syncthreads();
		sh[threadIdx.x] += t;
		__syncthreads();
	}
	if (i < n) ps[i + 1] = sh[threadIdx.x];
	if (threadIdx.x == SCAN_BLOCK - 1) block_tot[blockIdx.x] = sh[SCAN_BLOCK - 1];
	if (i == 0) ps[0] = 0;
}

// phase 2: exclusive scan of the block totals (single block, sequential over chunks; n_blocks is small)
__global__ void __launch_bounds__(SCAN_BLOCK) block_tot_scan_kernel(long long *block_tot, int n_blocks) {
	__shared__ long long sh[SCAN_BLOCK];
	__shared__ long long carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int base = 0; base < n_blocks; base += SCAN_BLOCK) {
		const int i = base + threadIdx.x;
		const long long v = i < n_blocks ? block_tot[i] : 0;
		sh[threadIdx.x] = v;
		__syncthreads();
		for (int off = 1; off < SCAN_BLOCK; off <<= 1) {
			long long t = 0;
			if ((int)threadIdx.x >= off) t = sh[threadIdx.x - off];
			__syncthreads();
			sh[threadIdx.x] += t;
			__syncthreads();
		}
		if (i < n_blocks) block_tot[i] = carry + sh[threadIdx.x] - v;   // exclusive
		__syncthreads();
		if (threadIdx.x == 0) carry += sh[SCAN_BLOCK - 1];
		__syncthreads();
	}
}

// phase 3: add the block offsets
__global__ void __launch_bounds__(SCAN_BLOCK) sq_scan_add_kernel(long long n, long long *ps, const long long *block_tot) {
	const long long i = (long long)blockIdx.x * SCAN_BLOCK + threadIdx.x;
	if (i < n) ps[i + 1] += block_tot[blockIdx.x];
}

// the walk (silence.py:37-46), one thread; segments = {src, dst, len}
__global__ void silence_walk_kernel(const long long *__restrict__ ps, long long n, int frame_len, int shift_len, double perc,
                                    longlong3 *__restrict__ segs, long long *__restrict__ counts /* [0]=n_segs, [1]=n_out */) {
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	const double average_energy = (double)ps[n] / (double)n;      // np.sum(signal ** 2) / float(siglen)
	const double thr = average_energy * perc;
	long long i = 0, out = 0, ns = 0;
	while (i < n) {
		const long long e = (i + frame_len < n) ? i + frame_len : n;
		const long long len = e - i;
		const double ave = (double)(ps[e] - ps[i]) / (double)len;
		if (ave < thr) {
			i += frame_len;
		} else {
			const long long add = len < shift_len ? len : shift_len;
			segs[ns] = make_longlong3(i, out, add);
			++ns;
			out += add;
			i += shift_len;
		}
	}
	counts[0] = ns;
	counts[1] = out;
}

__global__ void silence_gather_kernel(const void *pcm, int is_u8, const longlong3 *__restrict__ segs, const long long *__restrict__ counts,
                                      void *out) {
	const long long ns = counts[0];
	for (long long s = blockIdx.x; s < ns; s += gridDim.x) {
		const longlong3 sg = segs[s];
		for (long long t = threadIdx.x; t < sg.z; t += blockDim.x) {
			if (is_u8) {
				// signal - 128, then + 255/2 = 127 (Python 2 integer division), cast back to uint8 (wraps like astype)
				const long long v = (long long)reinterpret_cast<const uint8_t *>(pcm)[sg.x + t] - 128 + 127;
				reinterpret_cast<uint8_t *>(out)[sg.y + t] = (uint8_t)v;
			} else {
				reinterpret_cast<int16_t *>(out)[sg.y + t] = reinterpret_cast<const int16_t *>(pcm)[sg.x + t];
			}
		}
	}
}

// d_ps [n+1], d_block [ceil(n/1024)], d_segs [n / shift_len + 2], d_counts [2] are caller-provided workspaces
int launch_remove_silence(const void *d_pcm, int is_u8, long long n, int frame_len, int shift_len, double perc, long long *d_ps,
                          long long *d_block, longlong3 *d_segs, long long *d_counts, void *d_out, cudaStream_t st) {
	SRB_REQUIRE(n >= 1, "remove_silence: empty signal");
	SRB_REQUIRE(frame_len >= 1 && shift_len >= 1, "remove_silence: frame length/shift must be >= 1 sample");
	const int nb = (int)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
	sq_scan_block_kernel<<<nb, SCAN_BLOCK, 0, st>>>(d_pcm, is_u8, n, d_ps, d_block);
	SRB_CUDA(cudaGetLastError());
	block_tot_scan_kernel<<<1, SCAN_BLOCK, 0, st>>>(d_block, nb);
	SRB_CUDA(cudaGetLastError());
	sq_scan_add_kernel<<<nb, SCAN_BLOCK, 0, st>>>(n, d_ps, d_block);
	SRB_CUDA(cudaGetLastError());
	silence_walk_kernel<<<1, 32, 0, st>>>(d_ps, n, frame_len, shift_len, perc, d_segs, d_counts);
	SRB_CUDA(cudaGetLastError());
	silence_gather_kernel<<<ctx().sm_count * 4, 128, 0, st>>>(d_pcm, is_u8, d_segs, d_counts, d_out);
	SRB_CUDA(cudaGetLastError());
	return 0;
}

// ---- streaming windows: dst[b] = [src[b][hop:], fresh[b]] for B streams of W samples (one launch per tick; the reference GUI
// re-slices a Python list per tick, gui.py:179-190)
__global__ void ring_shift_append_kernel(const int16_t *__restrict__ src, const int16_t *__restrict__ fresh, int16_t *__restrict__ dst, int W, int hop) {
	const int b = blockIdx.y;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < W; i += gridDim.x * blockDim.x)
		dst[(size_t)b * W + i] = i < W - hop ? src[(size_t)b * W + i + hop] : fresh[(size_t)b * hop + (i - (W - hop))];
}

int launch_ring_shift_append(const int16_t *d_src, const int16_t *d_fresh, int16_t *d_dst, int n_streams, int W, int hop, cudaStream_t st) {
	SRB_REQUIRE(d_src && d_fresh && d_dst && d_src != d_dst, "stream_push: null or aliased buffers");
	SRB_REQUIRE(n_streams >= 1 && W >= 2 && hop >= 1 && hop < W, "stream_push: need 0 < hop < window (hop %d, window %d)", hop, W);
	dim3 grid((unsigned)std::min(64, (W + 255) / 256), (unsigned)n_streams);
	ring_shift_append_kernel<<<grid, 256, 0, st>>>(d_src, d_fresh, d_dst, W, hop);
	SRB_CUDA(cudaGetLastError());
	return 0;
}

}  // namespace srb
