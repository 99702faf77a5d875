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
constexpr int TC_HHEAD = 3;          // header words of a block's hard-entry list: [0] bit masks of the columns with entries in the four 32-column groups,
                                     // [1] {first entry of group 1, of group 2, of group 3, total}, [2] bit masks of the VALID columns (rows that are real
                                     // mixtures, not speaker padding) of the four groups
constexpr int TC_HSTRIDE = TC_HHEAD + 125;  // int4 per block (2 KB): header + up to 125 hard entries (incl. the x4 padding of the four groups); entries sorted by
                                     // column, every group padded to a multiple of 4 with null entries: {4 * dim, float bits of a'', of nb'', last entry of
                                     // its column ? 1 + 128 * (column within the group) : 0}
constexpr double TC_HARD_BUDGET = 1000.0;   // per mixture row: sum over the dims left in the GEMM of nb''^2 (the magnitude the expanded square cancels at);
                                            // the dims above it are added back exactly by the epilogue (api.cu pack_models_tc_host)


// ---- second generation (gmm_score_tc2.cuh): fp16 head/tail images, CTA pairs (tcgen05 cta_group::2) ----
constexpr int T2_NH = TC_NB / 2;     // rows of a block each CTA of the pair holds
constexpr int T2_SLOTS = 3;          // TMEM accumulator slots: columns [0, 384)
constexpr int T2_A_COL = 384;        // TMEM columns [384, 384 + Kp/2): PHI head (two fp16 per column), [384 + Kp/2, 384 + Kp): PHI tail
constexpr int T2_KP_MAX = 128;       // 2 D + 1 <= 127
constexpr int T2_BSTAGES_MAX = 6;    // ring of half-block stages (head | tail image of 64 rows: 256 Kp bytes)
constexpr int T2_HSTAGES = 8;        // ring of per-block hard-entry lists: the producer runs this many blocks ahead of the SLOWEST epilogue warp, and the
                                     // epilogue may lag the MMA warp by the three accumulator slots -- the B loads need the difference to cover the L2 latency
constexpr int T2_THREADS = 576;      // producer warp, MMA / relay warp, 16 epilogue warps
constexpr float T2_X_CLAMP = 240.f;  // |x''| bound before squaring (fp16 max 65504)
constexpr double T2_W_MAX = 60000.0; // |W| bound of an fp16 image entry; larger terms are added back exactly instead
__host__ __device__ static inline int t2_kp(int D) { return ((2 * D + 1 + 15) / 16) * 16; }
// halfs per block in HBM: [2 ranks][head | tail][64 * Kp]
__host__ __device__ static inline size_t t2_block_halfs(int Kp) { return (size_t)4 * T2_NH * Kp; }
// element (row r of the 64, k) of one image part: no-swizzle K-major canonical layout (core matrix = 8 rows x 8 halfs), K-groups outermost
__host__ __device__ static inline size_t t2_img_off(int r, int k) {
	return ((size_t)(k >> 3) * (T2_NH / 8) + (size_t)(r >> 3)) * 64 + (r & 7) * 8 + (k & 7);
}

__host__ __device__ static inline size_t tc_b_block_floats(int Kp) { return (size_t)2 * TC_NB * Kp; }  // hi | lo
// element (row r, k) of one part (hi or lo) of a B image: no-swizzle K-major canonical layout, K-groups outermost
__host__ __device__ static inline size_t tc_img_off(int rows, int r, int k) {
	return (size_t)(k >> 2) * (rows / 8) * 32 + (size_t)(r >> 3) * 32 + (r & 7) * 4 + (k & 3);
}
// row stride (floats) of the raw feature tile kept in shared memory for the add-back: odd, so that the 32 rows of a warp
// hit 32 different banks
__host__ __device__ static inline int tc_x_stride(int D) { return D | 1; }

}  // namespace srb
