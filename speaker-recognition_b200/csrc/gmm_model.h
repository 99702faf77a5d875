// gmm_model.h -- host-side model objects and the packed device layout shared by the scoring and EM kernels.
//
// Reference types being replaced: Gaussian (gmm.hh:24-46: dim, mean[], sigma[] = std-dev) and GMM
// (gmm.hh:123-168: nr_mixtures, dim, weights[], Gaussian*[]), an AoS of heap vectors walked per mixture.
// Here a model is three flat float64 arrays on the host (the truth for I/O) and, on the device, a
// tile-packed fp32 form the kernels stream with the TMA engine:
//
//   mixture row r = s * Kpad + k            (speaker-major, K padded to a multiple of 32 rows)
//   tile  = 32 consecutive rows             (never straddles a speaker)
//   per tile, contiguous in HBM:  [D][16 chunks][a_{2c}, a_{2c+1}, nb_{2c}, nb_{2c+1}]  then  c2[32]
//       (one LDS.128 per chunk = the two 64-bit operand pairs of a packed fma.rn.f32x2)
//       a   = sqrt(log2(e)/2) / sigma              (so that t = x*a + nb, t*t is the exponent in log2 units)
//       nb  = -mean * a
//       c2  = log2(w) - sum_d log2(sqrt(2 pi) sigma_d)      (padded rows: a = nb = 0, c2 = -1e30)
//   tile bytes = 256*D + 128  (16-B multiple -> one cp.async.bulk per tile)
#pragma once
#include "common.cuh"

namespace srb {

constexpr int MT = 32;            // mixture rows per packed tile
constexpr int MAX_DIM = 128;      // scoring smem budget: D*1024 + 3*(256*D+128) <= 227 KB
constexpr float PAD_C2 = -1.0e30f;

static inline size_t tile_floats(int D) { return (size_t)64 * D + 32; }

struct PackedModels {             // device-resident packed parameters of S models sharing D and Kpad
	int S = 0, K = 0, Kpad = 0, D = 0, tiles_per_model = 0;
	int S_arg = 0;                // models that take part in the per-utterance argmax (0 = all): a set may carry its UBM as an extra model
	float *d_tiles = nullptr;     // S * tiles_per_model * tile_floats(D)
	size_t bytes = 0;
	// tensor-core path (gmm_score_tc2.cuh), built on the host with the FMA tiles, uploaded on first use
	std::vector<uint16_t> h_tc2;  // weight images: fp16 head/tail, [blocks][2 ranks][head | tail][64 * Kp]
	std::vector<int> h_tc_hard;   // per block: 4*TC_HSTRIDE ints (header + hard entries), see gmm_score_tc_layout.h
	void *d_tc2 = nullptr;
	int tc2_Kp = 0;
	std::vector<float> h_tc_cs;   // [2][D]: per-dimension centre c_d and power-of-two scale alpha_d of the feature transform x'' = (x - c) alpha
	int *d_tc_hard = nullptr;
	float *d_tc_cs = nullptr;
	int tc_cols_per_spk = 0, tc_blocks_total = 0;
	bool tc_built = false;        // the tensor-core images describe the CURRENT parameters (possibly empty: set not supported)
	void release();
	void release_tc();            // drop the tensor-core images (host and device); the FMA tiles stay
};

// host-side packing of S models (each K_s <= K rows used; others padded) into `out` (size S*tiles*tile_floats)
void pack_models_host(int S, int K, int D, const double *const *weights, const double *const *means,
                      const double *const *sigmas, const int *Ks, std::vector<float> &out, int *Kpad_out);
int upload_packed(PackedModels &pm, int S, int K, int D, const std::vector<float> &host, int Kpad, cudaStream_t st);
// host-side build of the tensor-core weight images into pm.h_tc (no device work)
void pack_models_tc_host(PackedModels &pm, int S, int K, int D, const double *const *weights, const double *const *means,
                         const double *const *sigmas, const int *Ks);

}  // namespace srb

// The opaque handle of the C ABI (include/srb200.h).  Host arrays are the truth; device copies are caches.
struct GMM {
	int nr_mixtures = 0;
	int covariance_type = 1;
	int dim = 0;
	std::vector<double> weights, means, sigmas;   // [K], [K*D], [K*D]
	// device state
	srb::PackedModels packed;        // S = 1 scoring/EM layout; valid iff packed_valid
	bool packed_valid = false;
	double *d_params = nullptr;      // float64 master copy on device during training: w[K] | mean[K*D] | sigma[K*D]
	int d_params_K = 0, d_params_D = 0;
	bool master_valid = false;       // d_params mirrors (or is newer than) the host arrays
	bool host_stale = false;         // device master newer than host arrays (after *_dev M-steps)
	void invalidate() { packed_valid = false; master_valid = false; host_stale = false; }
	~GMM();
};

namespace srb {
int gmm_ensure_packed(GMM *g);       // (re)build + upload the packed tiles from the host arrays
int gmm_ensure_device_master(GMM *g);
int gmm_pull_device_master(GMM *g);
}
