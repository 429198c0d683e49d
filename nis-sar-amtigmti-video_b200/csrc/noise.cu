// noise.cu -- thermal noise + K-distributed sea clutter injected into a raw echo on the device (replaces add_ocean_noise,
// sar_satellite_sim.py:331-344 = sar_vehicle_sim.py:152-165 = sar_satellite_moving_sim.py:188-206, and
// generate_noise_tensor, sar_batch_sim.py:65-81).
//
//   signal_power  = mean |raw|^2                                       k_power_sum   (8 B/sample read)
//   raw[i]       += sqrt(Pn/2) (N1 + j N2)                             k_noise_add   (8 B read + 8 B write per sample)
//                 + sqrt(Pc * G * E) exp(j U),  G ~ Gamma(nu, 1/nu), E ~ Exp(1), U ~ U(0, 2 pi)
// with Pn = signal_power / 10^(snr/10), Pc = signal_power / 10^(scr/10).  Random numbers: Philox4x32-10, counter = sample
// index (+ a draw number for the Gamma rejection loop), key = seed -- every sample's draws are independent of the launch
// shape, so a seed reproduces the array exactly.  Parity with the reference is statistical (its generator is numpy's
// unseeded Mersenne Twister): same distributions and powers, tested by moments and Kolmogorov-Smirnov distances.
// The powers stay on the device (the mean is read from the reduction's output), so the echo never leaves HBM between
// synthesis and focusing and there is no host synchronisation.
#include <math.h>

#include "common.cuh"

using namespace nis;

namespace {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
// uniform in (0, 1): 24 random bits, never 0 or 1
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * 5.9604644775390625e-8f; }

__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    const float r = sqrtf(-2.0f * logf(u01(a)));
    float s, c;
    sincospif(2.0f * u01(b), &s, &c);
    return make_float2(r * c, r * s);
}

// Gamma(shape a, scale 1): Marsaglia & Tsang (2000); a < 1 through Gamma(a + 1) U^(1/a).  Draws come from Philox blocks
// (index, draw = 2, 3, ...) of the same sample.
__device__ float gamma_mt(float a, uint64_t index, uint2 key) {
    const float a1 = a < 1.0f ? a + 1.0f : a;
    const float d = a1 - (1.0f / 3.0f), c = rsqrtf(9.0f * d);
    float g = d;
    uint32_t draw = 2;
    for (int it = 0; it < 64; ++it, ++draw) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)index, (uint32_t)(index >> 32), draw, 0u), key);
        const float2 n = box_muller(r.x, r.y);
        const float v0 = 1.0f + c * n.x;
        if (v0 > 0.0f) {
            const float v = v0 * v0 * v0;
            if (logf(u01(r.z)) < 0.5f * n.x * n.x + d - d * v + d * logf(v)) {
                g = d * v;
                if (a < 1.0f) g *= powf(u01(r.w), 1.0f / a);
                break;
            }
        }
    }
    return g;
}

__global__ void __launch_bounds__(256) k_power_sum(const float2* __restrict__ x, uint64_t n, double* __restrict__ sum) {
    double acc = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        acc += sq_mag_f64(x[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += part[w];
        atomicAdd(sum, t);
    }
}

// |x|^2 is non-negative, so its fp64 bit pattern orders like an unsigned integer
__global__ void __launch_bounds__(256) k_power_max(const float2* __restrict__ x, uint64_t n, double* __restrict__ out) {
    double m = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        m = fmax(m, sq_mag_f64(x[i]));
    atomic_max_f64(out, warp_max_f64(m));
}

__global__ void __launch_bounds__(256) k_noise_add(float2* __restrict__ x, uint64_t n, const double* __restrict__ power_dev,
                                                   double power_value, double inv_snr, double inv_scr, float k_nu,
                                                   uint2 key, int accumulate) {
    const double p = power_dev != nullptr ? *power_dev * power_value : power_value;
    const float sig_n = (float)sqrt(0.5 * p * inv_snr);   // sqrt(noise_power / 2)
    const float pc = (float)(p * inv_scr);                 // clutter_power
    const float inv_nu = 1.0f / k_nu;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 r0 = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0u, 0u), key);
        const uint4 r1 = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 1u, 0u), key);
        const float2 th = box_muller(r0.x, r0.y);
        const float speckle = -logf(u01(r0.z));
        const float texture = (k_nu == 1.0f) ? -logf(u01(r1.x)) : gamma_mt(k_nu, i, key) * inv_nu;
        const float amp = sqrtf(pc * texture * speckle);
        float s, c;
        sincospif(2.0f * u01(r0.w), &s, &c);
        float2 v = make_float2(fmaf(sig_n, th.x, amp * c), fmaf(sig_n, th.y, amp * s));
        if (accumulate) {
            const float2 o = x[i];
            v.x += o.x;
            v.y += o.y;
        }
        x[i] = v;
    }
}

}  // namespace

extern "C" int nis_power_sum(nis_ctx* ctx, const nis_c32* x, uint64_t n, double* sum_dev, nis_stream stream) {
    NIS_REQUIRE(ctx && x && sum_dev, "nis_power_sum: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    NIS_CUDA_TRY(cudaMemsetAsync(sum_dev, 0, sizeof(double), st));
    if (n == 0) return NIS_OK;
    uint64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
    const uint64_t cap = (uint64_t)ctx->num_sms * 8;
    if (blocks > cap) blocks = cap;
    k_power_sum<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float2*>(x), n, sum_dev);
    NIS_LAUNCH_CHECK(ctx);
    return NIS_OK;
}

extern "C" int nis_power_max(nis_ctx* ctx, const nis_c32* x, uint64_t n, double* max_dev, nis_stream stream) {
    NIS_REQUIRE(ctx && x && max_dev, "nis_power_max: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    NIS_CUDA_TRY(cudaMemsetAsync(max_dev, 0, sizeof(double), st));
    if (n == 0) return NIS_OK;
    uint64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
    const uint64_t cap = (uint64_t)ctx->num_sms * 8;
    if (blocks > cap) blocks = cap;
    k_power_max<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float2*>(x), n, max_dev);
    NIS_LAUNCH_CHECK(ctx);
    return NIS_OK;
}

extern "C" int nis_noise_add(nis_ctx* ctx, nis_c32* x, uint64_t n, const double* power_sum_dev, double ref_power,
                             double snr_db, double scr_db, double k_nu, uint64_t seed, int32_t accumulate,
                             nis_stream stream) {
    NIS_REQUIRE(ctx && x, "nis_noise_add: null argument");
    NIS_REQUIRE(k_nu > 0, "nis_noise_add: the K-distribution shape must be positive (got %g)", k_nu);
    NIS_REQUIRE(ref_power >= 0, "nis_noise_add: negative reference power / power scale");
    if (n == 0) return NIS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    uint64_t blocks = (n + 256 * 4 - 1) / (256 * 4);
    const uint64_t cap = (uint64_t)ctx->num_sms * 8;
    if (blocks > cap) blocks = cap;
    k_noise_add<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<float2*>(x), n, power_sum_dev, ref_power,
                                                  1.0 / pow(10.0, snr_db / 10.0), 1.0 / pow(10.0, scr_db / 10.0), (float)k_nu,
                                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), accumulate);
    NIS_LAUNCH_CHECK(ctx);
    return NIS_OK;
}
