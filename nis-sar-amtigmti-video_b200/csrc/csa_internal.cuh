// csa_internal.cuh -- declarations shared by the power-of-two (csa.cu) and general-size
// (csa_generic.cu) Chirp Scaling paths.
#pragma once

#include <cuda.h>
#include <math.h>

#include <vector>

#include "common.cuh"

namespace nis {
namespace csa {

struct RowCoef {          // one per Doppler row
    uint64_t a1, b1, c1;  // Phi1: a1 n^2 + b1 n + c1
    uint64_t a2, b2, bn2; // Phi2: a2 k'^2 + b2 k', bn2 = b2 * n_rg (subtracted for negative bins)
    uint64_t a3, b3, c3;  // Phi3
    uint64_t pad;
};
static_assert(sizeof(RowCoef) == 80, "RowCoef layout");

__device__ __forceinline__ uint64_t quad_phase(uint64_t a, uint64_t b, uint64_t c, uint32_t n) {
    return a * (uint64_t)(n * n) + b * (uint64_t)n + c;  // n < 65536: n*n fits 32 bits
}
// Phi2 argument for FFT bin k of an n-point range transform (fftfreq sign convention, any n)
__device__ __forceinline__ uint64_t phi2_phase(const RowCoef& rc, uint32_t k, uint32_t n) {
    const bool neg = k >= (n + 1) / 2;
    const uint32_t ka = neg ? n - k : k;
    return rc.a2 * (uint64_t)(ka * ka) + rc.b2 * (uint64_t)k - (neg ? rc.bn2 : 0ull);
}

// Compact per-row form of a quadratic phase a n^2 + b n + c for the kernels that evaluate it at scattered (row, n) pairs
// (the azimuth kernels apply Phi1 on their store and Phi3 on their load): the 64-bit a and b split into words, c reduced to its
// top word with the half-ulp rounding offset of cis_u32_pre() folded in.  Only the top 32 bits of the sum are needed; dropping
// the carries out of the low words costs <= 3 * 2^-32 turns.
struct RowPhase {
    uint32_t a_lo, a_hi, b_lo, b_hi;
};
static_assert(sizeof(RowPhase) == 16, "RowPhase layout");

__device__ __forceinline__ uint32_t row_phase_hi(uint4 ab, uint32_t c_hi, uint32_t n, uint32_t n2) {
    uint32_t t = c_hi + ab.y * n2 + ab.w * n;
    t += __umulhi(ab.x, n2);
    t += __umulhi(ab.z, n);
    return t;
}

inline uint64_t to_fix(long double turns) {
    long double f = turns - floorl(turns);
    long double s = f * 18446744073709551616.0L;
    if (s >= 18446744073709551615.0L) return 0xFFFFFFFFFFFFFFFFull;
    return (uint64_t)s;
}

// Per-row phase coefficients (reference arithmetic in fp64 for D, Cs, tau_ref -- :244-262 --, the
// polynomial expansion in long double).  Row rho holds Doppler bin kk = rho/A2 + A1*(rho%A2);
// A1 = 1 gives natural bin order.
std::vector<RowCoef> build_row_coefs(int n_az, int n_rg, const nis_csa_params& prm, int A1, int A2);
void build_axes(int n_az, int n_rg, const nis_csa_params& prm, std::vector<double>& range_axis,
                std::vector<double>& cross_range);

// four-step azimuth engine as a stand-alone object (csa.cu): workspace + launchers, for other focusing algorithms
bool az_engine_supported(int n_az, int n_rg);
int az_engine_create(nis_ctx* ctx, int n_az, int n_rg, nis_csa_plan** out);

// one-length row DFT engine of the general-size path (csa_generic.cu): mixed radix or Bluestein
struct RowDft;
bool rowdft_supported(int n);
int rowdft_create(int n, RowDft** out);
void rowdft_destroy(RowDft* d);
// in-place DFT of n_rows rows (pitch elements apart); inverse = conjugate transform times `scale`
int rowdft_run(nis_ctx* ctx, const RowDft* d, float2* data, int64_t pitch, int n_rows, bool inverse, float scale,
               cudaStream_t st);

struct GenericState;   // csa_generic.cu
int generic_supported(int n_az, int n_rg);
int generic_create(nis_csa_plan* pl);
void generic_destroy(nis_csa_plan* pl);
int generic_focus(nis_csa_plan* pl, const float2* phist, int64_t pitch, float2* slc, double* max_sq, cudaStream_t st);

}  // namespace csa
}  // namespace nis

// every stage takes the window it works on: columns [col0, col0 + ncols), row blocks [k10, k10 + nk1), rows
typedef int (*nis_az_outer_fwd_fn)(nis_csa_plan*, const float2*, int64_t, int col0, int ncols, cudaStream_t);
typedef int (*nis_az_outer_inv_fn)(nis_csa_plan*, float2*, double*, int col0, int ncols, cudaStream_t);
typedef int (*nis_az_inner_fn)(nis_csa_plan*, bool, int col0, int ncols, int k10, int nk1, cudaStream_t);
typedef int (*nis_range_fn)(nis_csa_plan*, int row0, int nrows, cudaStream_t);

struct nis_csa_plan {
    nis_ctx* ctx = nullptr;
    int n_az = 0, n_rg = 0, A1 = 0, A2 = 0;
    int size_class = 0;   // 1: fused power-of-two path, 2: general-size path
    nis_csa_params prm{};
    float2* work = nullptr;
    float2* tw_inner = nullptr;
    float2* tw_full = nullptr;
    float2* tw_rg = nullptr;
    nis::csa::RowCoef* coef = nullptr;
    // Phi1 / Phi3 in the compact form the azimuth kernels read (null: the range kernel applies all three phase functions)
    uint4* ph1_ab = nullptr;
    uint32_t* ph1_c = nullptr;
    uint4* ph3_ab = nullptr;
    uint32_t* ph3_c = nullptr;
    bool phase_in_az = false;
    CUtensorMap tile_map{};   // [n_az][n_rg] workspace, box = min(A2,256) rows x inner_w columns
    int inner_w = 0;
    std::vector<double> range_axis, cross_range;
    // optional per-stage timing: a ring of event sets, one per nis_csa_focus call
    static constexpr int kProfRing = 64;
    bool profiling = false;
    uint64_t prof_calls = 0;
    cudaEvent_t prof_ev[kProfRing][6] = {};
    nis_az_outer_fwd_fn outer_fwd = nullptr;
    nis_az_outer_inv_fn outer_inv = nullptr;
    nis_az_inner_fn inner = nullptr;
    nis_range_fn range = nullptr;
    int (*outer_inv_mag)(nis_csa_plan*, float* mag, float scale, cudaStream_t) = nullptr;   // inverse outer stage -> |.|, no corner turn
    // whole-column azimuth transforms by thread-block clusters (one HBM pass each)
    bool az_cluster = false;
    int (*az_fwd)(nis_csa_plan*, const float2*, int64_t, cudaStream_t) = nullptr;
    int (*az_inv)(nis_csa_plan*, float2*, double*, cudaStream_t) = nullptr;
    CUtensorMap work_map3{}, in_map{};
    const void* in_map_base = nullptr;
    int64_t in_map_pitch = 0;
    nis::csa::GenericState* generic = nullptr;
};
