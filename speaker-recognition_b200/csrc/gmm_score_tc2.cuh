// gmm_score_tc2.cuh -- tensor-core diagonal-GMM scorer for sm_100a: tcgen05 kind::f16 as "3xFP16" on CTA PAIRS.
//
// Same quantity as gmm_score.cu (log2 p(x) per frame and speaker, per-utterance sums).  The Mahalanobis term is expanded over
// centred, power-of-two-scaled features x'' = (x - c_d) alpha_d (api.cu pack_models_tc_host),
//     l2_k(x) = C_k + sum_d [ (-a''_kd^2) x''_d^2 + (-2 a''_kd nb''_kd) x''_d ],     C_k = c2_k - sum_d nb''_kd^2,
// i.e. a GEMM  L[frames x rows] = PHI[frames x Kp] . W[rows x Kp]^T,  PHI = [x''^2, x'', 1, 0-pad],  Kp = 16 ceil((2D+1)/16).
// Every fp32 value of PHI and W is split by ROUND-TO-NEAREST into an fp16 head and an fp16 tail (22 significant bits, unbiased;
// round 1 truncated to tf32 heads: 2^-21 one-sided) and the product runs as three fp16 passes, hi*hi + lo*hi + hi*lo, fp32
// accumulation in tensor memory: the same accuracy class as 3xTF32 at HALF the tensor time (a kind::f16 MMA covers K = 16 in the
// 64 cycles a kind::tf32 MMA needs for K = 8; measured, experiments/tc2_gemm_test.cu) and half the operand bytes.
// The few (row, dim) entries whose nb'' is so large that the expanded form would cancel are left out of W and C and added back
// EXACTLY in fp32 by the epilogue (t = a'' x'' + nb'', l2 -= t*t), from a per-block sparse list.
//
// Structure: a CLUSTER OF TWO CTAs (one per SM of a TPC) works as one tcgen05 cta_group::2 unit: 256 frames (128 per CTA) against
// the stream of 128-row weight blocks.  Each CTA bulk-copies only ITS HALF of every block (64 rows: rank r supplies rows
// [64 r, 64 r + 64) of the block -- verified by the probe) into its own shared memory, so the L2 -> SM traffic per frame is half
// that of a single-CTA kernel (round 1 ran at the chip's ~6.3 KB/clk L2 limit: every SM streamed every block).
// Warp roles (the SM's issue arbiter favours the HIGHER warp id of a sub-partition: the light, latency-critical roles get the
// highest ids -- with the MMA issuer as warp 1 below sixteen busy epilogue warps, its ~250 instructions per block were starved
// to ~1400 cycles per block against 960 of tensor time, ncu):
//   warp 18 producer (both CTAs): cp.async.bulk of the block's hard-entry list and of this CTA's half block (fp16 head image and
//           tail image, one 256 Kp-byte copy) into a ring of up to 6 stages
//   warp 19 leader CTA: MMA issuer -- waits for its own stage and for the peer's (relayed), then ONE elected lane issues the block's
//           3 x Kp/16 tcgen05.mma.cta_group::2.kind::f16 (M 256 = 2 x 128, N 128, K 16) with the A operand (PHI head | tail) in TENSOR
//           MEMORY and both commits (multicast to both CTAs: free the stage, publish the accumulator slot -- three 128-column slots);
//           with one election per block the operands stay in uniform registers (15 UTCHMMA back to back in SASS)
//           peer CTA: relay -- forwards "my half of stage s has landed" to the leader with a remote mbarrier arrive
//   warps 0-15 epilogue, thread = frame row = TMEM lane, the four warps of a lane quarter own the four 32-column groups of every
//           block: per unit they build PHI (global -> registers -> tcgen05.st) and keep the raw x'' in shared memory; per block
//           they tcgen05.ld the accumulator, release the slot (remote arrive on the leader's barrier), subtract the exact terms of
//           the hard entries (t^2 summed per flagged column through per-thread scratch), and fold the 32 values into an online
//           log2-sum-exp2; at a speaker's end every warp hands its (max, sum) to the finishers through a shared-memory ring -- no
//           barrier between the epilogue warps.  All shared-memory and barrier accesses of the block loop use 32-bit shared-window
//           addresses computed once per thread (a generic pointer costs an S2R CgaCtaId + LEA conversion per use in a cluster kernel)
//   warps 16-17 finishers (64 rows each): merge the four column groups of a speaker, log2, clamp, per-frame output; the row's value
//           is parked in shared memory and the per-(tile, speaker) row sums are taken 32 speakers at a time with lane = speaker
// What bounds the kernel (measured: DESIGN.md 4.1a, profiles/r2g_scorer_ab.txt): the serial per-block chain of an epilogue warp, not
// the tensor pipe, the MUFU unit or the instruction count.  SRB_T2_PROF builds count cycles per warp role, SRB_T2_FAKE builds leave
// out parts of the work (timing only, wrong results) -- experiments/ab_build.py, experiments/tc2_prof.py.
// B images are K-major in the no-swizzle canonical layout (8-row x 16-byte core matrices = 8 rows x 8 halfs; LBO = 1024 B between
// K-adjacent core matrices of a 64-row half block, SBO = 128 B between 8-row groups), stored in HBM in shared-memory order.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "gmm_score_tc_layout.h"

namespace srb {

struct Tc2Args {
	const float *X;            // [n][ld] features
	int ld;
	const float *cs;           // [2][D] feature transform of the model set: centre c_d, power-of-two scale alpha_d
	const __half *B;           // [n_blocks_total][2 ranks][head | tail][64 * Kp] weight images
	const int4 *ftiles;        // per frame tile of 128 frames: {first frame, frame count, utterance, 0}
	int n_ftiles;
	int Kp;
	int S;                     // speakers
	int blocks_per_spk;        // 128-row blocks per speaker (rows padded to a multiple of 128) -- or, when seg < 128:
	int seg;                   // columns per speaker inside a block (32 or 64) when a speaker is smaller than a block
	int n_blocks_total;
	int n_chunks;              // the block range is cut into n_chunks contiguous chunks (whole speakers each)
	int spk_per_chunk;
	float *frame_ll;           // optional [S][ldf]
	int ldf;
	double *partial;           // [n_ftiles][S][T2_PARTS]: one row sum per finisher warp
	int clamp, frame_log2;
	int D;
	int bstages;               // B ring depth
	const int4 *hard;          // [n_blocks_total][TC_HSTRIDE], layout in gmm_score_tc_layout.h
};

#if SRB_T2_PROF          /* experiments only: where do the warps of CTA pair 0 spend their cycles? [20 warps][8 counters] per CTA rank */
__device__ unsigned long long g_tc2_prof[2][20][8];
#define T2P_DECL unsigned long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pt = clock64(); const long long pt00 = pt
#define T2P(k)                              \
	do {                                    \
		const long long now_ = clock64();   \
		pf[k] += (unsigned long long)(now_ - pt); \
		pt = now_;                          \
	} while (0)
#define T2P_FLUSH                                                                                      \
	do {                                                                                               \
		pf[6] = (unsigned long long)(clock64() - pt00);                                                \
		if (lane == 0 && blockIdx.x < 2)                                                               \
			for (int k_ = 0; k_ < 8; ++k_) atomicAdd(&g_tc2_prof[blockIdx.x][warp][k_], pf[k_]);       \
	} while (0)
#else
#define T2P_DECL
#define T2P(k)
#define T2P_FLUSH
#endif
#if SRB_T2_PROF == 2      /* finer split of the epilogue's block loop */
#define T2E(coarse, fine) T2P(fine)
#else
#define T2E(coarse, fine) T2P(coarse)
#endif

namespace tc2 {

__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_rank() {
	uint32_t r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}
__device__ __forceinline__ void cluster_sync_all() {
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t rank) {
	uint32_t r;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
	return r;
}
// arrive on a barrier of the LEADER CTA (works from either CTA of the pair: `addr` is a shared::cluster address).  Default
// semantics (release at CTA scope), as CUTLASS's ClusterBarrier::arrive(cta_id) uses for the same hand-overs: what is handed over
// lives in tensor memory / was written by the TMA engine and is ordered by the tcgen05 fences and the mbarrier itself -- a
// `.release.cluster` here compiles to MEMBAR.ALL.GPU in front of EVERY arrive (measured: the kernel ran at 4300 cycles per block).
__device__ __forceinline__ void arrive_cluster(uint32_t addr) {
	asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
// bounded waits: a broken pipeline traps instead of hanging the GPU
__device__ __forceinline__ void wait(uint64_t *bar, uint32_t parity) {
	if (mbar_try_wait(bar, parity)) return;
	int spins = 0;
	while (!mbar_try_wait(bar, parity)) {
		if (++spins > 400000000) __trap();
	}
}
// The same on 32-bit shared-window addresses computed ONCE per thread: the per-block loops of the epilogue add ring offsets to them
// (a generic pointer costs a generic -> shared conversion -- S2UR CgaCtaId, LEA, subtract -- at every use inside a cluster kernel).
__device__ __forceinline__ bool try_wait_a(uint32_t bar, uint32_t parity) {
	uint32_t ok;
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
	    "selp.u32 %0, 1, 0, p;\n\t}"
	    : "=r"(ok)
	    : "r"(bar), "r"(parity)
	    : "memory");
	return ok != 0;
}
__device__ __forceinline__ void wait_a(uint32_t bar, uint32_t parity) {
	if (try_wait_a(bar, parity)) return;
	int spins = 0;
	while (!try_wait_a(bar, parity)) {
		if (++spins > 400000000) __trap();
	}
}
__device__ __forceinline__ void arrive_a(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ int4 lds128(uint32_t a) {
	int4 v;
	asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
	return v;
}
__device__ __forceinline__ uint32_t keep(uint32_t x) {      // a value ptxas must hold in a register (it cannot re-execute a volatile statement)
	uint32_t y;
	asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
	return y;
}
__device__ __forceinline__ float4 lds128f(uint32_t a) {
	float4 v;
	asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
	return v;
}
__device__ __forceinline__ float lds32f(uint32_t a) {
	float v;
	asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
	return v;
}
__device__ __forceinline__ void sts32(uint32_t a, float x) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, float x, float y) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory"); }
// 32 accumulator columns of this thread's lane; the registers are valid after ld_wait()
__device__ __forceinline__ void ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
	    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
	      "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
	      "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
	      "=r"(r[31])
	    : "r"(taddr));
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t (&v)[8]) {
	asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
	             "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
	             : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T over the CTA pair: A (this CTA's 128 rows) is read from tensor memory, B (64 rows from each
// CTA) from shared memory.  For a region that ONE elected lane of the leader's MMA warp executes: with one election per BLOCK
// (not per instruction) every operand stays in uniform registers -- SASS: 15 UTCHMMA back to back with one UIADD3 between them,
// where the per-instruction form cost ELECT + 2 VOTEU + 4 R2UR + 2 UMOV each (ten instructions per MMA).
template <int ACC>
__device__ __forceinline__ void mma_f16_ts_1t(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
	asm volatile(
	    "{\n\t.reg .b64 db;\n\t.reg .pred p;\n\t"
	    "mov.b64 db, {%2, %3};\n\t"
	    "setp.ne.b32 p, %5, 0;\n\t"
	    "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem_d),
	    "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(ACC)
	    : "memory");
}
__device__ __forceinline__ void commit_pair_1t(uint64_t *bar) {
	asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
	             "h"((uint16_t)3)
	             : "memory");
}
__device__ __forceinline__ bool elect_one() {
	uint32_t ok;
	asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\tselp.u32 %0, 1, 0, pe;\n\t}" : "=r"(ok));
	return ok != 0;
}
// fp32 -> fp16 head + fp16 tail, both round-to-nearest; two values per call, packed (first value in the low half)
__device__ __forceinline__ void split_pack(float v0, float v1, uint32_t &hi, uint32_t &lo) {
	const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
	const __half l0 = __float2half_rn(v0 - __half2float(h0)), l1 = __float2half_rn(v1 - __half2float(h1));
	hi = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
	lo = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
}

}  // namespace tc2

// 2^x for two x <= 0 on the FMA pipe (the MUFU unit does 4 ex2 per clock and sub-partition and is the epilogue's busiest pipe):
// round-to-nearest split x = n + f by the magic-number add, degree-4 minimax polynomial for 2^f on [-0.5, 0.5] (max relative error
// 2.7e-6, mean 4e-8: fitted in experiments/f16x3_numerics.py's sibling script, checked in tests), 2^n spliced into the exponent.
__device__ __forceinline__ unsigned long long tc2_exp2_poly2(unsigned long long x01) {
	float a, b;
	asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x01));
	a = fmaxf(a, -126.f);                             // below: the exponent splice would wrap; such terms are 0 against a sum >= 1
	b = fmaxf(b, -126.f);
	unsigned long long x, t, n, f, p;
	const float MAGIC = 12582912.f;                   // 1.5 * 2^23: the integer part lands in the low mantissa bits
	unsigned long long mg, nmg, neg1, c0, c1, c2, c3, c4;
	asm("mov.b64 %0, {%1, %1};" : "=l"(mg) : "f"(MAGIC));
	asm("mov.b64 %0, {%1, %1};" : "=l"(nmg) : "f"(-MAGIC));
	asm("mov.b64 %0, {%1, %1};" : "=l"(neg1) : "f"(-1.f));
	asm("mov.b64 %0, {%1, %1};" : "=l"(c0) : "f"(0.9999992624f));
	asm("mov.b64 %0, {%1, %1};" : "=l"(c1) : "f"(0.693121851f));
	asm("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(0.2402474484f));
	asm("mov.b64 %0, {%1, %1};" : "=l"(c3) : "f"(0.0559175783f));
	asm("mov.b64 %0, {%1, %1};" : "=l"(c4) : "f"(0.009569989201f));
	asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a), "f"(b));
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(x), "l"(mg));
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(n) : "l"(t), "l"(nmg));
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(f) : "l"(n), "l"(neg1), "l"(x));
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p) : "l"(c4), "l"(f), "l"(c3));
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p) : "l"(p), "l"(f), "l"(c2));
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p) : "l"(p), "l"(f), "l"(c1));
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p) : "l"(p), "l"(f), "l"(c0));
	uint32_t p0, p1, t0, t1;
	asm("mov.b64 {%0, %1}, %2;" : "=r"(p0), "=r"(p1) : "l"(p));
	asm("mov.b64 {%0, %1}, %2;" : "=r"(t0), "=r"(t1) : "l"(t));
	p0 += t0 << 23;                                   // (0x4B400000 + n) << 23 == n << 23 (mod 2^32)
	p1 += t1 << 23;
	unsigned long long r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(p0), "r"(p1));
	return r;
}

// online log2-sum-exp2 update with 32 new values (one chunk of accumulator columns) for one frame; the first NP of the 32
// exponentials are evaluated on the FMA pipe, the others by MUFU.EX2
// FRESH: the running pair is known to be empty (m_run = -3e38, s_run = 0: the first chunk of a speaker) -- no rescale of the old sum
template <int NP, bool FRESH>
__device__ __forceinline__ void tc2_lse_update(const float *v, float &m_run, float &s_run) {
	// three-input maxima throughout (SASS FMNMX3): 10 + 4 + 2 instructions for the 32 values and the running maximum
	auto max3 = [](float a, float b, float c) { return fmaxf(fmaxf(a, b), c); };
	float t[10];
#pragma unroll
	for (int i = 0; i < 10; ++i) t[i] = max3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
	const float u0 = max3(t[0], t[1], t[2]), u1 = max3(t[3], t[4], t[5]), u2 = max3(t[6], t[7], t[8]), u3 = max3(t[9], v[30], v[31]);
	const float m_new = max3(max3(u0, u1, u2), u3, m_run);
	// packed fp32x2 adds (SASS FADD2): the 32 subtractions and the 32 sum updates cost 16 instructions each
	const float nm = -m_new;
	unsigned long long nm2, acc01 = 0ull, acc23 = 0ull;
	asm("mov.b64 %0, {%1, %1};" : "=l"(nm2) : "f"(nm));
#pragma unroll
	for (int i = 0; i < 32; i += 4) {
		unsigned long long p01, p23, d01, d23, e01, e23;
		asm("mov.b64 %0, {%1, %2};" : "=l"(p01) : "f"(v[i]), "f"(v[i + 1]));
		asm("mov.b64 %0, {%1, %2};" : "=l"(p23) : "f"(v[i + 2]), "f"(v[i + 3]));
		asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d01) : "l"(p01), "l"(nm2));
		asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d23) : "l"(p23), "l"(nm2));
		if (i < NP) {
			e01 = tc2_exp2_poly2(d01);
			e23 = tc2_exp2_poly2(d23);
		} else {
			float a, b, c, d;
			asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(d01));
			asm("mov.b64 {%0, %1}, %2;" : "=f"(c), "=f"(d) : "l"(d23));
#if (SRB_T2_FAKE & 1)
			asm("mov.b64 %0, {%1, %2};" : "=l"(e01) : "f"(a), "f"(b));
			asm("mov.b64 %0, {%1, %2};" : "=l"(e23) : "f"(c), "f"(d));
#else
			asm("mov.b64 %0, {%1, %2};" : "=l"(e01) : "f"(ex2_approx(a)), "f"(ex2_approx(b)));
			asm("mov.b64 %0, {%1, %2};" : "=l"(e23) : "f"(ex2_approx(c)), "f"(ex2_approx(d)));
#endif
		}
		asm("add.rn.f32x2 %0, %1, %2;" : "=l"(acc01) : "l"(acc01), "l"(e01));
		asm("add.rn.f32x2 %0, %1, %2;" : "=l"(acc23) : "l"(acc23), "l"(e23));
	}
	float s0, s1, s2, s3;
	asm("mov.b64 {%0, %1}, %2;" : "=f"(s0), "=f"(s1) : "l"(acc01));
	asm("mov.b64 {%0, %1}, %2;" : "=f"(s2), "=f"(s3) : "l"(acc23));
	if (FRESH) s_run = (s0 + s1) + (s2 + s3);
	else s_run = fmaf(s_run, ex2_approx(m_run - m_new), (s0 + s1) + (s2 + s3));
	m_run = m_new;
}

// KSTEPS = Kp / 16 as a compile-time constant (fully unrolled MMA issue), or 0 for the generic loop
template <int KSTEPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS, 1) gmm_score_tc2_kernel(const Tc2Args A) {
	extern __shared__ __align__(1024) unsigned char tc2_smem[];
	const int Kp = A.Kp, D = A.D, XS = tc_x_stride(D);
	const uint32_t part_bytes = (uint32_t)T2_NH * Kp * 2;              // one image part (head or tail) of a half block
	const uint32_t stage_bytes = 2 * part_bytes;
	const int NST = A.bstages;
	unsigned char *sB = tc2_smem;                                                               // [NST] stages: head | tail of this CTA's half block
	int4 *hring = reinterpret_cast<int4 *>(sB + (size_t)NST * stage_bytes);                    // [T2_HSTAGES][TC_HSTRIDE] hard-entry lists
	float *scr = reinterpret_cast<float *>(hring + T2_HSTAGES * TC_HSTRIDE);                    // [16 warps][32 lanes][T2_SCR_STRIDE] add-back sums
	float2 *fin = reinterpret_cast<float2 *>(scr + 16 * 32 * T2_SCR_STRIDE);                    // [T2_FSLOTS][4 groups][128 rows] (m, s) for the finishers
	uint64_t *bars = reinterpret_cast<uint64_t *>(fin + T2_FSLOTS * 4 * TC_M);
	uint64_t *full_b = bars, *peer_full = full_b + T2_BSTAGES_MAX, *empty_b = peer_full + T2_BSTAGES_MAX;
	uint64_t *empty_h = empty_b + T2_BSTAGES_MAX;
	uint64_t *tm_full = empty_h + T2_HSTAGES, *tm_empty = tm_full + T2_SLOTS;
	uint64_t *a_ready = tm_empty + T2_SLOTS;
	uint64_t *fin_full = a_ready + 1, *fin_empty = fin_full + T2_FSLOTS;
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(fin_empty + T2_FSLOTS);
	float *sX = reinterpret_cast<float *>(bars + 64);                                           // [128][XS] x'' of the unit's tile
	float *red = sX + (size_t)TC_M * XS;                                                        // [2 finishers][T2_RED_SPK][T2_RED_STRIDE] per-row log-likelihoods awaiting their row sums

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t rank = tc2::cluster_rank();
	const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
	if (tid == 0) {
		for (int s = 0; s < T2_BSTAGES_MAX; ++s) {
			mbar_init(&full_b[s], 1);
			mbar_init(&peer_full[s], 1);
			mbar_init(&empty_b[s], 1);
		}
		for (int s = 0; s < T2_HSTAGES; ++s) mbar_init(&empty_h[s], 16);      // one arrival per epilogue warp of this CTA
		for (int s = 0; s < T2_SLOTS; ++s) {
			mbar_init(&tm_full[s], 1);
			mbar_init(&tm_empty[s], 32);     // one arrival per epilogue warp of BOTH CTAs (used in the leader only)
		}
		mbar_init(a_ready, 32);
		for (int s = 0; s < T2_FSLOTS; ++s) {
			mbar_init(&fin_full[s], 16);     // every epilogue warp has stored its (m, s)
			mbar_init(&fin_empty[s], 2);     // both finisher warps have read them
		}
		fence_mbar_init();
	}
	if (warp == T2_WARP_MMA) {
		asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
	}
	tc2::fence_before();
	__syncthreads();
	tc2::cluster_sync_all();           // both CTAs' barriers are initialised and both allocations are done before any cross-CTA traffic
	tc2::fence_after();
	const uint32_t tmem_base = *tmem_slot;

	// work units: (pair of frame tiles, chunk of whole speakers); CTA rank r of the cluster takes tile 2 p + r
	const int n_pairs = (A.n_ftiles + 1) >> 1;
	const int n_units = n_pairs * A.n_chunks;
	const int blocks_per_chunk_full = (A.seg >= TC_NB) ? A.spk_per_chunk * A.blocks_per_spk : (A.spk_per_chunk * A.seg + TC_NB - 1) / TC_NB;

	if (warp == T2_WARP_PRODUCER) {
		// ======================= producer =======================
		if (lane == 0) {
			uint32_t it_h = 0, ph_b = 0;
			int st_b = 0;
			T2P_DECL;
			for (int u = cluster_id; u < n_units; u += n_clusters) {
				const int chunk = u % A.n_chunks;
				const int blk0 = chunk * blocks_per_chunk_full;
				const int blk1 = min(A.n_blocks_total, blk0 + blocks_per_chunk_full);
				for (int b = blk0; b < blk1; ++b, ++it_h) {
					// the block's hard-entry list and this CTA's half of the block (head image then tail image, contiguous in HBM) complete
					// on ONE barrier: an epilogue warp that sees the block's accumulator (tm_full) knows the list has landed too -- the MMA
					// warp waited for this barrier before it issued -- so the epilogue needs no wait of its own for the list
					const int hst = it_h % T2_HSTAGES;
					T2P(2);
					tc2::wait(&empty_h[hst], ((it_h / T2_HSTAGES) & 1) ^ 1);
					T2P(0);
					tc2::wait(&empty_b[st_b], ph_b ^ 1);
					T2P(1);
					const uint32_t copy_bytes = (SRB_T2_FAKE & 32) ? 1024u : stage_bytes;      /* timing experiment: a token copy */
					mbar_expect_tx(&full_b[st_b], copy_bytes + TC_HSTRIDE * 16);
					bulk_g2s(hring + (size_t)hst * TC_HSTRIDE, A.hard + (size_t)b * TC_HSTRIDE, TC_HSTRIDE * 16, &full_b[st_b]);
					const unsigned char *src = reinterpret_cast<const unsigned char *>(A.B) + ((size_t)b * 2 + rank) * stage_bytes;
					bulk_g2s(sB + (size_t)st_b * stage_bytes, src, copy_bytes, &full_b[st_b]);
					if (++st_b == NST) {
						st_b = 0;
						ph_b ^= 1;
					}
				}
			}
			T2P_FLUSH;
		}
	} else if (warp == T2_WARP_MMA && rank != 0) {
		// ======================= peer CTA: relay "my half of the stage has landed" to the leader =======================
		if (lane == 0) {
			uint32_t ph_b = 0;
			int st_b = 0;
			for (int u = cluster_id; u < n_units; u += n_clusters) {
				const int chunk = u % A.n_chunks;
				const int blk0 = chunk * blocks_per_chunk_full;
				const int blk1 = min(A.n_blocks_total, blk0 + blocks_per_chunk_full);
				for (int b = blk0; b < blk1; ++b) {
					tc2::wait(&full_b[st_b], ph_b);
					tc2::arrive_cluster(tc2::mapa(smem_u32(&peer_full[st_b]), 0));
					if (++st_b == NST) {
						st_b = 0;
						ph_b ^= 1;
					}
				}
			}
		}
	} else if (warp == T2_WARP_MMA) {
		// ======================= leader CTA: MMA issuer (whole warp converged; one lane elected per instruction) =======================
		// instruction descriptor: D = F32 (bits 4-5 = 1), A/B = F16 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24 (M = 256: the pair)
		const uint32_t idesc = (1u << 4) | ((uint32_t)(TC_NB >> 3) << 17) | ((uint32_t)((2 * TC_M) >> 4) << 24);
		constexpr uint32_t lbo_b = (T2_NH / 8) * 128, sbo = 128;
		constexpr uint32_t d_hi = (uint32_t)(sbo >> 4) | (1u << 14);            // SBO | descriptor version bit 46
		constexpr uint32_t stepB = (2 * lbo_b) >> 4;                            // one k-step = 16 halfs = two K-adjacent core matrices
		const int ksteps = KSTEPS > 0 ? KSTEPS : Kp / 16;
		const uint32_t a_hi = tmem_base + T2_A_COL, a_lo = a_hi + (uint32_t)(Kp >> 1);
		const uint32_t tail_off = part_bytes >> 4;
		uint32_t ph_b = 0, it_a = 0, it_s = 0;
		int st_b = 0;
		T2P_DECL;
		for (int u = cluster_id; u < n_units; u += n_clusters, ++it_a) {
			const int chunk = u % A.n_chunks;
			const int blk0 = chunk * blocks_per_chunk_full;
			const int blk1 = min(A.n_blocks_total, blk0 + blocks_per_chunk_full);
			T2P(4);
			tc2::wait(a_ready, it_a & 1);            // PHI of both CTAs' tiles is in tensor memory
			tc2::fence_after();
			T2P(0);
			for (int b = blk0; b < blk1; ++b, ++it_s) {
				const int slot = it_s % T2_SLOTS;
				tc2::wait(&tm_empty[slot], ((it_s / T2_SLOTS) & 1) ^ 1);
				T2P(1);
				const int st = st_b;
				tc2::wait(&full_b[st], ph_b);
				T2P(2);
				tc2::wait(&peer_full[st], ph_b);
				tc2::fence_after();
				T2P(5);
				const uint32_t d_tmem = tmem_base + (uint32_t)(slot * TC_NB);
				const uint32_t dB0 = ((smem_u32(sB + (size_t)st * stage_bytes) >> 4) & 0x3FFF) | ((lbo_b >> 4) << 16);
				// 3xFP16: (A_hi + A_lo) * B_hi, then A_hi * B_lo
				if (tc2::elect_one()) {                 // ONE election per block: the lane issues the block's MMAs and both commits
					tc2::mma_f16_ts_1t<0>(d_tmem, a_hi, dB0, d_hi, idesc);
#pragma unroll
					for (int k = 1; k < ((SRB_T2_FAKE & 16) ? 1 : ksteps); ++k) tc2::mma_f16_ts_1t<1>(d_tmem, a_hi + 8 * k, dB0 + stepB * k, d_hi, idesc);
#pragma unroll
					for (int k = 0; k < ((SRB_T2_FAKE & 16) ? 2 : ksteps); ++k) tc2::mma_f16_ts_1t<1>(d_tmem, a_lo + 8 * k, dB0 + stepB * k, d_hi, idesc);
#pragma unroll
					for (int k = 0; k < ((SRB_T2_FAKE & 16) ? 2 : ksteps); ++k) tc2::mma_f16_ts_1t<1>(d_tmem, a_hi + 8 * k, dB0 + tail_off + stepB * k, d_hi, idesc);
					tc2::commit_pair_1t(&empty_b[st]);
					tc2::commit_pair_1t(&tm_full[slot]);
				}
				__syncwarp();
				T2P(3);
				if (++st_b == NST) {
					st_b = 0;
					ph_b ^= 1;
				}
			}
		}
		T2P_FLUSH;
	} else if (warp >= T2_WARP_FIN0) {
		// ======================= finishers: warp 16 rows 0-63, warp 17 rows 64-127 =======================
		// Per speaker: merge the four column groups' (max, sum), log2, clamp, optional per-frame output.  The per-(tile, speaker) ROW
		// SUMS are not taken speaker by speaker (a 5-step shuffle tree per speaker and batch is a ~300-cycle dependent chain: measured, it
		// paced the whole kernel at 1 900 cycles per block against 1 370 without it): the per-row values of up to T2_RED_SPK speakers
		// are parked in shared memory, [speaker][row], and summed with lane = speaker -- 64 independent loads and adds per 32 speakers,
		// in a fixed order (deterministic), one store of partial[tile][speaker][finisher] per lane.
		const int f = warp - T2_WARP_FIN0;
		const int seg = A.seg;
		float *tb = red + (size_t)f * (T2_RED_SPK * T2_RED_STRIDE);
		uint32_t it_f = 0;
		T2P_DECL;
		for (int u = cluster_id; u < n_units; u += n_clusters) {
			const int pair = u / A.n_chunks, chunk = u - pair * A.n_chunks;
			const int tile = 2 * pair + (int)rank;
			const bool tile_ok = tile < A.n_ftiles;
			const int blk0 = chunk * blocks_per_chunk_full;
			const int blk1 = min(A.n_blocks_total, blk0 + blocks_per_chunk_full);
			const int4 ft = tile_ok ? A.ftiles[tile] : make_int4(0, 0, 0, 0);
			int spk = chunk * A.spk_per_chunk;
			int base = spk, cnt = 0;                  // speakers [base, base + cnt) are parked
			const int n_entries = seg >= TC_NB ? (blk1 - blk0) / A.blocks_per_spk : (blk1 - blk0);
			auto flush = [&]() {
				__syncwarp();
				if (cnt > 0) {
					// g = 32 / (cnt rounded up to a power of two) lanes share a speaker: each sums every g-th row, then a shuffle tree
					// over the g lanes (a full buffer: g = 1, lane = speaker, no shuffles; one speaker -- the E-step's pass A -- g = 32)
					int cp = 1;
					while (cp < cnt) cp <<= 1;
					const int g = 32 / cp, sl = lane / g, j = lane - sl * g;
					float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
					if (sl < cnt) {
						const float *r = tb + sl * T2_RED_STRIDE;
						for (int i = j; i < 64; i += 4 * g) {
							a0 += r[i];
							if (i + g < 64) a1 += r[i + g];
							if (i + 2 * g < 64) a2 += r[i + 2 * g];
							if (i + 3 * g < 64) a3 += r[i + 3 * g];
						}
					}
					float tot = (a0 + a1) + (a2 + a3);
					for (int off = g >> 1; off >= 1; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
					if (tile_ok && j == 0 && sl < cnt && base + sl < A.S) A.partial[((size_t)tile * A.S + base + sl) * T2_PARTS + f] = (double)tot;
				}
				__syncwarp();
				base += cnt;
				cnt = 0;
			};
			for (int e = 0; e < n_entries; ++e, ++it_f) {
				const int slot = it_f % T2_FSLOTS;
				T2P(2);
				tc2::wait(&fin_full[slot], (it_f / T2_FSLOTS) & 1);
				T2P(0);
				const float2 *fs = fin + (size_t)slot * 4 * TC_M;
#pragma unroll
				for (int batch = 0; batch < ((SRB_T2_FAKE & 8) ? 0 : 2); ++batch) {
					const int q = 2 * f + batch, row = q * 32 + lane;
					const bool valid = row < ft.y;
					float2 g[4];
#pragma unroll
					for (int c = 0; c < 4; ++c) g[c] = fs[c * TC_M + row];
					// close one speaker from its (m, s): clamp, per-frame output, park the row's value for the row sum
					auto out = [&](int my_spk, float m, float sacc) {
						if (my_spk < A.S && tile_ok) {
							const float l2 = m + lg2_approx(sacc);
							float ll = l2 * TC_LN2;
							if (A.clamp && !(ll >= -744.4400719213812f)) ll = -34.538776394910684f;
							if (valid && A.frame_ll != nullptr) A.frame_ll[(size_t)my_spk * A.ldf + ft.x + row] = A.frame_log2 ? l2 : ll;
							tb[(my_spk - base) * T2_RED_STRIDE + batch * 32 + lane] = valid ? ll : 0.f;
						}
					};
					auto merge2 = [](float2 a, float2 b) {
						const float M = fmaxf(a.x, b.x);
						return make_float2(M, a.y * ex2_approx(a.x - M) + b.y * ex2_approx(b.x - M));
					};
					if (seg >= TC_NB) {
						// four groups against their common maximum: 4 exponentials (a pairwise tree costs 6)
						const float M = fmaxf(fmaxf(g[0].x, g[1].x), fmaxf(g[2].x, g[3].x));
						const float sacc = (g[0].y * ex2_approx(g[0].x - M) + g[1].y * ex2_approx(g[1].x - M)) +
						                   (g[2].y * ex2_approx(g[2].x - M) + g[3].y * ex2_approx(g[3].x - M));
						out(spk, M, sacc);
					} else if (seg == 64) {
						const float2 r0 = merge2(g[0], g[1]), r1 = merge2(g[2], g[3]);
						out(spk, r0.x, r0.y);
						out(spk + 1, r1.x, r1.y);
					} else {
#pragma unroll
						for (int c = 0; c < 4; ++c) out(spk + c, g[c].x, g[c].y);
					}
				}
				__syncwarp();
				if (lane == 0) mbar_arrive(&fin_empty[slot]);
				T2P(1);
				const int step = seg >= TC_NB ? 1 : (seg == 64 ? 2 : 4);
				spk += step;
				cnt += step;
				if (cnt == T2_RED_SPK) flush();
			}
			flush();
		}
		T2P_FLUSH;
	} else {
		// ======================= epilogue (16 warps; thread = frame row; the 4 warps of a lane quarter take the 4 column groups) ====
		const int ew = warp;                          // 0..15
		const int q = warp & 3;                       // TMEM lane quarter this warp may access (hardware rule: warp id % 4)
		const int cg = ew >> 2;                       // 32-column group of every block this warp owns
		const int row = q * 32 + lane;                // frame within the tile
		const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
		float *my_scr = scr + ((size_t)ew * 32 + lane) * T2_SCR_STRIDE;      // this thread's flagged-column sums (slot p = p-th flagged column of the group)
		float *xrow = sX + (size_t)row * XS;
		// entry.w = 1 + byte offset of the scratch column; shared-window addresses for the scratch and the x'' row (a generic pointer
		// store costs an S2R CgaCtaId + LEA + IADD3 conversion per access inside the add-back loop)
		// (through an opaque move: at the register cap ptxas otherwise REMATERIALISES the conversion inside the loop)
		const uint32_t a_scr = tc2::keep(smem_u32(my_scr)), a_scr_m1 = a_scr - 1u, a_xrow = tc2::keep(smem_u32(xrow));
		const int seg = A.seg;
		const int cols_per_spk = (seg >= TC_NB) ? A.blocks_per_spk * TC_NB : seg;
		const int nch = Kp >> 4;                      // chunks of 16 k-values = 8 packed columns
		const uint32_t a_ready_leader = tc2::mapa(smem_u32(a_ready), 0);
		const uint32_t tm_empty_leader = tc2::mapa(smem_u32(tm_empty), 0);
		// shared-window addresses of everything the block loop touches, computed once (ring offsets are added per block)
		const uint32_t a_tm_full = smem_u32(tm_full), a_empty_h = smem_u32(empty_h), a_hdr = smem_u32(hring) + 16u * (uint32_t)cg;
		const uint32_t a_fin = smem_u32(fin) + 8u * (uint32_t)(cg * TC_M + row), a_fin_full = smem_u32(fin_full), a_fin_empty = smem_u32(fin_empty);
		const uint32_t t_ld = lane_base + (uint32_t)(cg * 32);
		const bool single = seg < TC_NB || A.blocks_per_spk == 1;      // every block closes a speaker: no running pair to rescale
		uint32_t ph_s = 0, ph_f = 0;                  // phase of the accumulator ring / of the finisher ring
		uint32_t slot8 = 0;                           // 8 x accumulator slot of the next block (barrier offset; x 16 = TMEM column offset)
		uint32_t hoff = 0;                            // byte offset of the next block's list in the hard-entry ring (its barrier: hoff >> 8)
		uint32_t fslot8 = 0;                          // 8 x finisher-ring slot of the next hand-over (rings never reset)
		static_assert(TC_NB == 128 && TC_HSTRIDE * 16 == 2048 && (T2_HSTAGES & (T2_HSTAGES - 1)) == 0 && T2_FSLOTS == 8, "ring offsets below");
#define SRB_TC2_QBAR() asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory")      /* the 4 warps of this lane quarter */
		T2P_DECL;
		for (int u = cluster_id; u < n_units; u += n_clusters) {
			const int pair = u / A.n_chunks, chunk = u - pair * A.n_chunks;
			const int tile = 2 * pair + (int)rank;
			const bool tile_ok = tile < A.n_ftiles;
			const int blk0 = chunk * blocks_per_chunk_full;
			const int blk1 = min(A.n_blocks_total, blk0 + blocks_per_chunk_full);
			const int4 ft = tile_ok ? A.ftiles[tile] : make_int4(0, 0, 0, 0);
			const bool valid = row < ft.y;

			// ---- PHI = [x''^2, x'', 1, 0-pad] of this thread's frame -> tensor memory (fp16 head | fp16 tail, two k per column); the
			// quarter's four warps split the 16-value chunks.  All MMAs of the previous unit have completed (this warp consumed its
			// last tm_full); the quarter barrier makes sure no sibling still reads the previous tile's x''.
			SRB_TC2_QBAR();
			{
				const float *xg = A.X + (size_t)(ft.x + (valid ? row : 0)) * A.ld;
				const float *cen = A.cs, *alp = A.cs + D;
				auto xpp = [&](int d) -> float {
					if (!valid) return 0.f;
					const float x = (__ldg(xg + d) - __ldg(cen + d)) * __ldg(alp + d);
					return fminf(fmaxf(x, -T2_X_CLAMP), T2_X_CLAMP);       // keeps x''^2 inside fp16 (frames out there underflow every mixture anyway)
				};
				auto val = [&](int k) -> float {
					if (k < D) {
						const float x = xpp(k);
						return x * x;
					}
					if (k < 2 * D) {
						const float x = xpp(k - D);
						xrow[k - D] = x;
						return x;
					}
					return (k == 2 * D && valid) ? 1.f : 0.f;
				};
				for (int c = cg; c < nch; c += 4) {
					uint32_t hi[8], lo[8];
#pragma unroll
					for (int j = 0; j < 8; ++j) tc2::split_pack(val(16 * c + 2 * j), val(16 * c + 2 * j + 1), hi[j], lo[j]);
					tc2::st8(lane_base + T2_A_COL + 8 * c, hi);
					tc2::st8(lane_base + T2_A_COL + (Kp >> 1) + 8 * c, lo);
				}
				asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
				tc2::fence_before();
				SRB_TC2_QBAR();                                // the quarter's x'' are in shared memory
				if (lane == 0) tc2::arrive_cluster(a_ready_leader);
				// the NEXT unit's frames towards L2 now: its PHI build sits between two units with the tensor pipe drained
				const int un = u + n_clusters;
				if (cg == 0 && un < n_units) {
					const int tn = 2 * (un / A.n_chunks) + (int)rank;
					if (tn != tile && tn < A.n_ftiles) {
						const int4 fn = A.ftiles[tn];
						if (row < fn.y) {
							const float *pn = A.X + (size_t)(fn.x + row) * A.ld;
							asm volatile("prefetch.global.L2 [%0];" ::"l"(pn));
							if (A.ld > 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(pn + 32));
						}
					}
				}
			}

			T2E(5, 6);
			float m_run = -3.0e38f, s_run = 0.f;
			int col_in_spk = 0;                       // columns of the current speaker consumed so far (by the quarter's warps together)

			// end of a speaker segment: this warp's running (max, sum) of its 32-column group goes to the finishers.  The ring is
			// checked once per FOUR hand-overs (the finishers release in order: the last slot of the next four being free means all are)
			auto handoff = [&]() {
				// (waiting for the slot about to be written at EVERY hand-over instead -- a full ring of slack -- measured 0.5 % slower)
				if ((fslot8 & 24u) == 0) tc2::wait_a(a_fin_empty + fslot8 + 24u, ph_f ^ 1);
				T2E(4, 5);
				tc2::sts64(a_fin + fslot8 * (4 * TC_M), m_run, s_run);
				__syncwarp();
				if (lane == 0) tc2::arrive_a(a_fin_full + fslot8);
				fslot8 = (fslot8 + 8u) & (8u * T2_FSLOTS - 1);
				if (fslot8 == 0) ph_f ^= 1;
				m_run = -3.0e38f;
				s_run = 0.f;
			};

			for (int b = blk0; b < blk1; ++b) {
				// ---- accumulator -> registers.  The only wait of the block: tm_full also covers the block's hard-entry list (producer).
				T2E(4, 7);
				tc2::wait_a(a_tm_full + slot8, ph_s);
				tc2::fence_after();
				T2E(0, 0);
				uint32_t vr[32];
#if (SRB_T2_FAKE & 4)         /* timing experiment: the accumulator is never read */
#pragma unroll
				for (int i = 0; i < 32; ++i) vr[i] = 0x3f800000u + (uint32_t)b;
				const int4 hd = make_int4(0, 0, 0, -1);
#else
				tc2::ld32_issue(t_ld + 16u * slot8, vr);
				const int4 hd = tc2::lds128(a_hdr + hoff);   // this warp's group: {first entry, end entry, flagged columns, valid-column mask}
				tc2::ld_wait();
#endif
				T2E(1, 0);
				tc2::fence_before();
				__syncwarp();
				if (lane == 0) tc2::arrive_cluster(tm_empty_leader + slot8);      // the leader's MMA warp overwrites both CTAs' slots
				T2E(2, 1);
				float v[32];
#pragma unroll
				for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(vr[i]);
				// ---- exact add-back (in one block only the warps of ONE column group have entries; the group rotates with the block):
				// sum t^2 of the hard entries per column -- the list is uniform across the warp, sorted by column, `last` flag on a
				// column's final entry -- park each column's sum in this thread's scratch slots (a dynamic column index cannot address
				// registers), then subtract: the flagged columns are the group's FIRST hd.z columns, a multiple of four
				if (hd.y > hd.x) {
					const uint32_t a_e = a_hdr - 16u * (uint32_t)cg + hoff + 16u * (uint32_t)(TC_HHEAD + hd.x);      // the group's first entry
					const int n = hd.y - hd.x;
					float csum = 0.f;
#pragma unroll 1
					for (int i = 0; i < n; i += 4) {      // four entries per step, all loads of a step ahead of its first scratch store
						const int4 a0 = tc2::lds128(a_e + 16u * i), a1 = tc2::lds128(a_e + 16u * i + 16u), a2 = tc2::lds128(a_e + 16u * i + 32u), a3 = tc2::lds128(a_e + 16u * i + 48u);
						const float x0 = tc2::lds32f(a_xrow + (uint32_t)a0.x), x1 = tc2::lds32f(a_xrow + (uint32_t)a1.x);
						const float x2 = tc2::lds32f(a_xrow + (uint32_t)a2.x), x3 = tc2::lds32f(a_xrow + (uint32_t)a3.x);
						const float t0 = fmaf(x0, __int_as_float(a0.y), __int_as_float(a0.z)), t1 = fmaf(x1, __int_as_float(a1.y), __int_as_float(a1.z));
						const float t2 = fmaf(x2, __int_as_float(a2.y), __int_as_float(a2.z)), t3 = fmaf(x3, __int_as_float(a3.y), __int_as_float(a3.z));
						csum = fmaf(t0, t0, csum);
						if (a0.w) {
							tc2::sts32(a_scr_m1 + (uint32_t)a0.w, csum);
							csum = 0.f;
						}
						csum = fmaf(t1, t1, csum);
						if (a1.w) {
							tc2::sts32(a_scr_m1 + (uint32_t)a1.w, csum);
							csum = 0.f;
						}
						csum = fmaf(t2, t2, csum);
						if (a2.w) {
							tc2::sts32(a_scr_m1 + (uint32_t)a2.w, csum);
							csum = 0.f;
						}
						csum = fmaf(t3, t3, csum);
						if (a3.w) {
							tc2::sts32(a_scr_m1 + (uint32_t)a3.w, csum);
							csum = 0.f;
						}
					}
#pragma unroll
					for (int c = 0; c < T2_MAX_FLAGGED / 4; ++c) {
						if (4 * c < hd.z) {
							const float4 t = tc2::lds128f(a_scr + 16u * c);
							v[4 * c] -= t.x;
							v[4 * c + 1] -= t.y;
							v[4 * c + 2] -= t.z;
							v[4 * c + 3] -= t.w;
						}
					}
				}
				const uint32_t vmask = (uint32_t)hd.w;
				T2E(2, 2);
				__syncwarp();
				if (lane == 0) tc2::arrive_a(a_empty_h + (hoff >> 8));
				slot8 += 8u;
				if (slot8 == 8u * T2_SLOTS) {
					slot8 = 0;
					ph_s ^= 1;
				}
				hoff = (hoff + TC_HSTRIDE * 16) & (T2_HSTAGES * TC_HSTRIDE * 16 - 1);
				// ---- columns beyond a speaker's mixtures (row padding) take no part in the sum
				if (vmask != 0xffffffffu) {
#pragma unroll
					for (int j = 0; j < 32; ++j)
						if (!(vmask & (1u << j))) v[j] = -INFINITY;
				}
				// column segmentation: seg >= 128 -> the block belongs to one speaker (quarter-wide merge at its end); seg == 64 ->
				// two speakers per block, two warps each; seg == 32 -> every warp's group is its own speaker
				T2E(2, 3);
#if (SRB_T2_FAKE & 6)         /* timing experiments only: results are wrong */
				m_run = v[0] + v[31];
				s_run = 1.f;
#else
				if (single) tc2_lse_update<T2_POLY_EXP, true>(v, m_run, s_run);
				else tc2_lse_update<T2_POLY_EXP, false>(v, m_run, s_run);
#endif
				T2E(3, 4);
#if SRB_T2_PROF == 1
				pf[7] += 1;
#endif
				if (seg < TC_NB) {
					handoff();
				} else {
					col_in_spk += TC_NB;
					if (col_in_spk == cols_per_spk) {
						handoff();
						col_in_spk = 0;
					}
				}
			}
		}
		T2P_FLUSH;
#undef SRB_TC2_QBAR
	}
	tc2::fence_before();
	__syncthreads();
	tc2::cluster_sync_all();           // the peer's tensor memory and barriers stay alive until the leader's last MMA / commit has landed
	if (warp == T2_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

}  // namespace srb
