// gmm_score_tc_layout.h -- tile sizes and operand layouts of the tensor-core scorer (gmm_score_tc2.cuh), shared with the
// host-side model packer (api.cu pack_models_tc_host).
#pragma once
#include <stddef.h>

namespace srb {

constexpr int TC_M = 128;            // frames per tile and per CTA (TMEM lanes); a CTA pair works on 256
constexpr int TC_NB = 128;           // mixture rows per weight block = columns of one accumulator slot (N = 128 runs at the 64-cycle MMA peak,
                                     // experiments/tc_mma_rate.cu and tc2_gemm_test.cu)
constexpr float TC_LN2 = 0.69314718055994531f;

// ---- the per-block list of exactly corrected ("hard") entries: TC_HSTRIDE int4 per 128-row block ----
//   header word g (g = 0..3, one per 32-column group): {first entry, end entry, flagged columns n (the group's FIRST n columns, a
//       multiple of 4), bit mask of the VALID columns (rows that are real mixtures, not speaker padding)}
//   entries, sorted by column, every group padded to a multiple of 4 with null entries:
//       {4 * dim, float bits of a'', float bits of nb'', last entry of its column ? 1 + 4 * (scratch slot = column within the group) : 0}
constexpr int TC_HHEAD = 4;
constexpr int TC_HSTRIDE = TC_HHEAD + 124;  // 2 KB per block
constexpr double TC_HARD_BUDGET = 2000.0;   // per mixture row: sum over the dims left in the GEMM of nb''^2 (the magnitude the expanded square cancels at);
                                            // the dims above it are added back exactly by the epilogue.  Emulated per-frame error vs float64
                                            // (experiments/f16x3_numerics.py): 7e-6 at 800, 1.8e-5 at 2000, 4.2e-5 with no add-back at all; bound 1e-4

// ---- fp16 head/tail images, CTA pairs (tcgen05 cta_group::2) ----
constexpr int T2_NH = TC_NB / 2;     // rows of a block each CTA of the pair holds
constexpr int T2_SLOTS = 3;          // TMEM accumulator slots: columns [0, 384)
constexpr int T2_A_COL = 384;        // TMEM columns [384, 384 + Kp/2): PHI head (two fp16 per column), [384 + Kp/2, 384 + Kp): PHI tail
constexpr int T2_KP_MAX = 128;       // 2 D + 1 <= 127
constexpr int T2_BSTAGES_MAX = 6;    // ring of half-block stages (head | tail image of 64 rows: 256 Kp bytes)
constexpr int T2_HSTAGES = 8;        // ring of per-block hard-entry lists: the producer runs this many blocks ahead of the SLOWEST epilogue warp, and the
                                     // epilogue may lag the MMA warp by the three accumulator slots -- the B loads need the difference to cover the L2 latency
constexpr int T2_THREADS = 640;      // 16 epilogue warps (0-15: TMEM lane quarter = warp % 4), 2 finisher warps, producer warp, MMA / relay warp
constexpr int T2_WARP_FIN0 = 16, T2_WARP_PRODUCER = 18, T2_WARP_MMA = 19;
constexpr int T2_FSLOTS = 8;         // ring of (max, sum) hand-overs from the epilogue warps to the finishers: [slot][4 groups][128 rows] float2
constexpr int T2_PARTS = 2;          // row-sum partials per (frame tile, speaker): one per finisher warp (64 rows each)
constexpr int T2_RED_SPK = 32;       // speakers whose per-row values a finisher parks before it takes their row sums (lane = speaker)
constexpr int T2_RED_STRIDE = 65;    // floats per parked speaker: 64 rows + 1 (the 32 lanes of the row sum hit 32 different banks)
constexpr int T2_MAX_FLAGGED = 16;   // flagged (exactly corrected) columns per 32-column group; the packer orders them first
constexpr int T2_SCR_STRIDE = 20;    // floats per thread of add-back scratch: 16 slots, stride 20 keeps the 128-bit loads conflict-free
#ifndef SRB_T2_POLY_EXP
#define SRB_T2_POLY_EXP 0
#endif
constexpr int T2_POLY_EXP = SRB_T2_POLY_EXP;   // of every 32 exponentials of the epilogue, this many run as a polynomial on the FMA pipe instead of MUFU.EX2.
                                               // Measured on B200 (C2 x 64 utterances): 0 -> 3.28 ms, 8 -> 3.34 ms, 16 -> 3.65 ms: the epilogue is bound by
                                               // issue slots (0.58 IPC, 63 % MUFU), not by the MUFU unit, so the extra FMA instructions lose; kept switchable.
#ifndef SRB_T2_PROF
#define SRB_T2_PROF 0
#endif
#ifndef SRB_T2_FAKE
#define SRB_T2_FAKE 0
#endif
constexpr float T2_X_CLAMP = 240.f;  // |x''| bound before squaring (fp16 max 65504)
constexpr double T2_W_MAX = 60000.0; // |W| bound of an fp16 image entry; larger terms are added back exactly instead
__host__ __device__ static inline int t2_kp(int D) { return ((2 * D + 1 + 15) / 16) * 16; }
// halfs per block in HBM: [2 ranks][head | tail][64 * Kp]
__host__ __device__ static inline size_t t2_block_halfs(int Kp) { return (size_t)4 * T2_NH * Kp; }
// element (row r of the 64, k) of one image part: no-swizzle K-major canonical layout (core matrix = 8 rows x 8 halfs), K-groups outermost
__host__ __device__ static inline size_t t2_img_off(int r, int k) {
	return ((size_t)(k >> 3) * (T2_NH / 8) + (size_t)(r >> 3)) * 64 + (r & 7) * 8 + (k & 7);
}
// row stride (floats) of the x'' tile kept in shared memory for the add-back: odd, so that the 32 rows of a warp hit 32 different banks
__host__ __device__ static inline int tc_x_stride(int D) { return D | 1; }

}  // namespace srb
