// gmm_score_tc_layout.h -- tile sizes and operand layouts of the tensor-core scorer (gmm_score_tc.cuh), shared with the
// host-side model packer (api.cu).
#pragma once
#include <stddef.h>

namespace srb {

constexpr int TC_M = 128;            // frames per tile (TMEM lanes)
constexpr int TC_NB = 128;           // mixture rows per B block / accumulator slot (N = 128: measured, experiments/tc_mma_rate.cu --
                                     // an M=128 tf32 MMA costs >= 60 cycles whatever N <= 64 is, N = 128 runs at the 64-cycle peak)
constexpr int TC_SLOTS = 2;          // TMEM accumulator slots: columns [0, 256)
constexpr int TC_A_COL = 256;        // TMEM columns [256, 256 + Kp): PHI hi (tf32), [256 + Kp, 256 + 2 Kp): PHI lo
constexpr int TC_KP_MAX = 128;       // so that both halves of PHI fit in the 512 TMEM columns; shared memory is the tighter limit
                                     // (gmm_score_tc.cu tc_stages: 3 B stages up to Kp = 80 / D = 39, 2 stages up to Kp = 112 / D = 55)
constexpr int TC_BSTAGES = 3;        // ring of B PART stages: a stage holds the hi or the lo image of one block (128 x Kp floats)
constexpr int TC_HSTAGES = 4;        // ring of per-block hard-entry lists
constexpr int TC_THREADS = 576;      // producer warp, MMA warp, 16 epilogue warps (four per TMEM lane quarter: 32-column groups)
constexpr float TC_LN2 = 0.69314718055994531f;
constexpr int TC_HSTRIDE = 2 + 204;  // int4 per block: two header words + up to 204 hard entries (incl. the x4 padding of the four groups)
constexpr double TC_HARD_BUDGET = 1000.0;   // per mixture row: sum over the dims left in the GEMM of nb''^2 (the magnitude the expanded square cancels at);
                                            // the dims above it are added back exactly by the epilogue (api.cu pack_models_tc_host)

__host__ __device__ static inline size_t tc_b_block_floats(int Kp) { return (size_t)2 * TC_NB * Kp; }  // hi | lo
// element (row r, k) of one part (hi or lo) of a B image: no-swizzle K-major canonical layout, K-groups outermost
__host__ __device__ static inline size_t tc_img_off(int rows, int r, int k) {
	return (size_t)(k >> 2) * (rows / 8) * 32 + (size_t)(r >> 3) * 32 + (r & 7) * 4 + (k & 3);
}
// row stride (floats) of the raw feature tile kept in shared memory for the add-back: odd, so that the 32 rows of a warp
// hit 32 different banks
__host__ __device__ static inline int tc_x_stride(int D) { return D | 1; }

}  // namespace srb
