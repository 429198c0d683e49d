// common.cuh -- shared host/device helpers for libnis_sar (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/nis_sar.h"

namespace nis {

constexpr int kNumSMsB200 = 148;

void set_error(const char* fmt, ...);

// device the calling thread is bound to (function attributes and occupancy answers are cached per device)
inline int current_device() {
    int d = 0;
    cudaGetDevice(&d);
    return d;
}

#define NIS_CUDA_TRY(expr)                                                              \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            ::nis::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                             __FILE__, __LINE__);                                      \
            return NIS_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

#define NIS_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            ::nis::set_error(__VA_ARGS__);     \
            return NIS_ERR_INVALID;            \
        }                                      \
    } while (0)

// after a kernel launch
#define NIS_LAUNCH_CHECK(ctx)                  \
    do {                                       \
        (ctx)->launches++;                     \
        NIS_CUDA_TRY(cudaGetLastError());      \
    } while (0)

}  // namespace nis

struct nis_ctx {
    int device = 0;
    int num_sms = nis::kNumSMsB200;
    uint64_t launches = 0;
    // Small library-owned workspaces (pulse tables of the backprojector, histograms of the order-statistics select) are
    // kept PER STREAM: two calls in flight on different streams of one device never share bytes.  A buffer that has to
    // grow is retired, not freed (a captured CUDA graph may still hold its address), and growth is refused while the
    // stream is capturing.  Large workspaces (nis_gmti_fused) belong to the caller.
    struct Scratch { void* ptr = nullptr; size_t bytes = 0; };
    std::mutex mu;
    std::map<cudaStream_t, Scratch> scratch;
    std::vector<void*> retired;
    int stream_scratch(cudaStream_t st, size_t bytes, void** out);
};

namespace nis {
// cudaSetDevice for the lifetime of a C entry point; the caller's current device is restored on return
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
}  // namespace nis

namespace nis {

// out[c][r] = in[r][c]; `in` rows are in_pitch elements apart (api.cu)
int launch_transpose(nis_ctx* ctx, const float2* in, int64_t in_pitch, float2* out, int rows, int cols,
                     cudaStream_t st);

// ---------------------------------------------------------------- device helpers
#ifdef __CUDA_ARCH__
#define NIS_LDG(p) __ldg(p)
#else
#define NIS_LDG(p) (*(p))
#endif

__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__host__ __device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
// Packed fp32x2 complex arithmetic (sm_100: FADD2 / FMUL2 / FFMA2).  A complex number already sits in an aligned
// register pair, and the SASS forms take per-operand half swaps (.LO_HI), lane sign patterns (.NP) and scalar
// broadcasts (.F32), so ptxas folds the mov.b64 shuffles below into the arithmetic: a complex add is ONE instruction and
// a complex multiply TWO (scalar code: 2 and 4).  Same FP32 pipe time (measured, csrc/pipebench.cu), half the issue slots --
// which is what the FFT kernels are short of.
__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#ifdef __CUDA_ARCH__
__device__ __forceinline__ unsigned long long pk2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float2 upk2(unsigned long long r) {
    float2 c;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(r));
    return c;
}
#endif
__host__ __device__ __forceinline__ float2 cadd_pk(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
    return upk2(r);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__host__ __device__ __forceinline__ float2 csub_pk(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
    return upk2(r);
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
__host__ __device__ __forceinline__ float2 cmul_pk(float2 a, float2 b) {   // a * b
#ifdef __CUDA_ARCH__
    unsigned long long t, r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.x)));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a.y, a.x)), "l"(pk2(-b.y, b.y)), "l"(t));
    return upk2(r);
#else
    return cmul(a, b);
#endif
}
__host__ __device__ __forceinline__ float2 cmul_conj_pk(float2 a, float2 b) {   // a * conj(b)
#ifdef __CUDA_ARCH__
    unsigned long long t, r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.x)));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a.y, a.x)), "l"(pk2(b.y, -b.y)), "l"(t));
    return upk2(r);
#else
    return cmul_conj(a, b);
#endif
}
// s * a - b with a real s: one step of a three-term recurrence on a complex sequence (one FFMA2)
__host__ __device__ __forceinline__ float2 cfms_pk(float s, float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(s, s)), "l"(pk2(a.x, a.y)), "l"(pk2(-b.x, -b.y)));
    return upk2(r);
#else
    return make_float2(fmaf(s, a.x, -b.x), fmaf(s, a.y, -b.y));
#endif
}
__host__ __device__ __forceinline__ float2 cscale_pk(float2 a, float s) {   // a * real s
#ifdef __CUDA_ARCH__
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(s, s)));
    return upk2(r);
#else
    return make_float2(a.x * s, a.y * s);
#endif
}

// Phase arithmetic in "turns" held as unsigned fixed point: value = u / 2^64 (or / 2^32).
// Integer wrap-around IS the mod-1 reduction, so quadratic phases of 1e4..1e8 rad keep full
// precision without fp64 in the inner loops.
__host__ __device__ __forceinline__ uint64_t turns_to_u64(double t) {
    t -= floor(t);                       // [0,1)
    double s = t * 18446744073709551616.0;  // 2^64
    if (s >= 18446744073709551615.0) return 0xFFFFFFFFFFFFFFFFull;
    return (uint64_t)s;
}

// exp(j 2 pi u/2^32): top 32 bits of the phase -> signed turns in [-0.5,0.5) -> MUFU sin/cos.
// Absolute error ~5e-7 (MUFU.SIN/COS on a reduced argument).
__device__ __forceinline__ float2 cis_u32(uint32_t u) {
    // top 23 phase bits become the mantissa of a float in [1, 2): no int->float conversion (which
    // would share the quarter-rate pipe with the two MUFU ops).  a = 2 pi (f - 1.5) in [-pi, pi);
    // the half-turn offset is undone by negating both outputs.
    const float f = __uint_as_float(((u + 0x100u) >> 9) | 0x3f800000u);   // rounded, wraps mod 1
    const float a = fmaf(f, 6.283185307179586f, -9.42477796076938f);
    float s, c;
    __sincosf(a, &s, &c);
    return make_float2(-c, -s);
}
// same, for a phase that already carries the +0x100 rounding offset
__device__ __forceinline__ float2 cis_u32_pre(uint32_t u) {
    const float f = __uint_as_float((u >> 9) | 0x3f800000u);
    const float a = fmaf(f, 6.283185307179586f, -9.42477796076938f);
    float s, c;
    __sincosf(a, &s, &c);
    return make_float2(-c, -s);
}
__device__ __forceinline__ float2 cis_u64(uint64_t u) { return cis_u32((uint32_t)(u >> 32)); }

// |s|^2 in fp64: both products are exact for fp32 inputs, one rounding in the sum
__device__ __forceinline__ double sq_mag_f64(float2 s) {
    const double re = (double)s.x, im = (double)s.y;
    return fma(re, re, im * im);
}
__device__ __forceinline__ double warp_max_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// lane 0 of each warp publishes; non-negative doubles order like their bit patterns
__device__ __forceinline__ void atomic_max_f64(double* dst, double v) {
    if ((threadIdx.x & 31) == 0 && v > 0.0)
        atomicMax(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)__double_as_longlong(v));
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace nis
