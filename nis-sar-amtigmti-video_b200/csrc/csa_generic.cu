// csa_generic.cu -- Chirp Scaling focusing for sizes that are not powers of two (the reference's
// default scene is 7199 x 13200 after the DPCA pulse shift, sar_ati_dcpa_sim_csa.py:402-403).
//
// Every transform is a row transform; the two azimuth transforms run on corner-turned data:
//   transpose  raw[n_az][n_rg] -> T[n_rg][n_az]      (T is the caller's slc buffer)
//   k_row<AZ_FWD>   T   rows: azimuth DFT
//   transpose  T -> W[n_az][n_rg]
//   k_row<RANGE>    W   rows: x Phi1, range DFT, x Phi2, inverse range DFT, x Phi3
//   transpose  W -> T
//   k_row<AZ_INV>   T   rows: inverse azimuth DFT, 1/(n_az n_rg)  => slc[n_rg][n_az]
// A length is handled by one of two engines:
//   mixed radix  (all prime factors <= 13, length <= 14000): generalised Stockham passes between two
//                shared-memory row buffers, radix-2/4/8/16 butterflies from fft.cuh, odd radices as a
//                small DFT with twiddles from the exp(-2 pi i m / N) table;
//   Bluestein    (any length <= 8192): x[n] a[n] -> FFT_M -> x FFT_M(b)/M -> IFFT_M -> x a[k], a[n] =
//                exp(-i pi n^2 / N), M in {64, 256, 1024, 4096, 16384} >= 2N-1, on the register-resident
//                power-of-two transforms of fft.cuh.
#include <stdlib.h>

#include <complex>

#include "csa_internal.cuh"
#include "fft.cuh"
#include "mixed_ct.cuh"

using namespace nis;
using namespace nis::fft;
using namespace nis::csa;

namespace {

constexpr int kMaxPass = 12;
constexpr int kMaxMixedLen = 14000;   // 2 row buffers of float2 must fit 227 KB of shared memory
constexpr int kMaxBluesteinLen = 8192;

enum Mode { AZ_FWD = 0, RANGE = 1, AZ_INV = 2 };

struct GenDev {   // by-value kernel argument
    int N, npass, M;
    int radix[kMaxPass];
    uint32_t inv_ns[kMaxPass];   // ceil(2^32 / Ns) of every pass (Ns = product of the earlier radices)
    const float2* twN;
    const float2* chirp;
    const float2* bfft;
    const float2* tw_pow2;
    const float2* bfe;      // pruned Bluestein (N <= M/2, M = 16384): even / odd bins of bfft, w_M^n, 8192-point twiddles
    const float2* bfo;
    const float2* twm;
    const float2* tw_half;
    uint64_t chirp_k;       // round(2^64 / 2N): n^2 * chirp_k >> 32 = (n^2 / 2N mod 1) in 2^-32 turns (integer wrap-around)
    int twm_shift;          // 32 - log2(M): n << twm_shift = n / M in 2^-32 turns
};

// Bluestein chirp a[n] = exp(-i pi n^2 / N) and modulation w_M^n = exp(-2 pi i n / M) evaluated on the fly (two MUFU
// operations each) instead of read from tables: the pruned core touches them four times per element, 256 KB per row that
// no L1 could hold beside 196 KB of shared memory -- ncu showed the stage waiting on L2 for them (long_scoreboard 37 %).
__device__ __forceinline__ float2 blue_chirp(const GenDev& g, uint32_t n) {
    const uint32_t u = (uint32_t)(((uint64_t)(n * n) * g.chirp_k) >> 32);
    return cis_u32(0u - u);
}
__device__ __forceinline__ float2 blue_twm(const GenDev& g, uint32_t n) { return cis_u32(0u - (n << g.twm_shift)); }

struct GenLen {
    int N = 0, kind = 0 /* 0 mixed, 1 Bluestein */, npass = 0, M = 0;
    int radix[kMaxPass] = {};
    float2 *twN = nullptr, *chirp = nullptr, *bfft = nullptr, *tw_pow2 = nullptr;
    float2 *bfe = nullptr, *bfo = nullptr, *twm = nullptr, *tw_half = nullptr, *tw_half32 = nullptr;
    float2* ct_tw = nullptr;   // per-pass tables of the compile-time plan (mixed_ct.cuh), when the length has one
    GenDev dev() const {
        GenDev d{};
        d.N = N; d.npass = npass; d.M = M;
        uint64_t ns = 1;
        for (int i = 0; i < kMaxPass; ++i) {
            d.radix[i] = radix[i];
            d.inv_ns[i] = (uint32_t)((0x100000000ull + ns - 1) / ns);
            if (i < npass) ns *= (uint64_t)radix[i];
        }
        d.twN = twN; d.chirp = chirp; d.bfft = bfft; d.tw_pow2 = tw_pow2;
        d.bfe = bfe; d.bfo = bfo; d.twm = twm; d.tw_half = tw_half;
        d.chirp_k = N > 0 ? (uint64_t)((((unsigned __int128)1 << 64) + (unsigned)N) / (2u * (unsigned)N)) : 0;
        d.twm_shift = 32;
        for (int m = M; m > 1; m >>= 1) --d.twm_shift;
        return d;
    }
    void release() {
        cudaFree(twN); cudaFree(chirp); cudaFree(bfft); cudaFree(tw_pow2);
        cudaFree(bfe); cudaFree(bfo); cudaFree(twm); cudaFree(tw_half); cudaFree(tw_half32); cudaFree(ct_tw);
    }
};

bool factorize(int n, int* radix, int* npass) {
    int k = 0;
    const int pow2[] = {16, 8, 4, 2};
    for (int p : pow2)
        while (n % p == 0 && k < kMaxPass) { radix[k++] = p; n /= p; }
    const int odd[] = {13, 11, 7, 5, 3};
    for (int p : odd)
        while (n % p == 0 && k < kMaxPass) { radix[k++] = p; n /= p; }
    *npass = k;
    return n == 1 && k > 0;
}
bool smooth(int n) {
    int r[kMaxPass], k;
    return n >= 2 && n <= kMaxMixedLen && factorize(n, r, &k);
}
int bluestein_m(int n) {
    const int ms[] = {64, 256, 1024, 4096, 16384};
    for (int m : ms)
        if (m >= 2 * n - 1) return m;
    return 0;
}
bool length_supported(int n) { return smooth(n) || (n >= 2 && n <= kMaxBluesteinLen && bluestein_m(n) > 0); }

// ----------------------------------------------------------------------------- mixed radix engine
// Odd radices: the R-point DFT is evaluated through the conjugate-symmetric pairs s_r = v_r + v_(R-r), d_r = v_r - v_(R-r):
//   X[q], X[R-q] = (v_0 + sum_r cos(2 pi r q / R) s_r) +- (i sum_r w_(rq).y d_r),   r, q = 1 .. (R-1)/2
// i.e. real-by-complex products only: (R-1)^2 FMAs instead of the (R-1)^2 complex multiplies of the direct sum.
// inv_ns = ceil(2^32 / Ns): j / Ns == umulhi(j, inv_ns) for every j * Ns < 2^32 (no integer division in the loop).
template <int R, bool INV>
__device__ __forceinline__ void mixed_pass(const float2* __restrict__ src, float2* __restrict__ dst, int N, int Ns,
                                           uint32_t inv_ns, const float2* __restrict__ twN) {
    const int nb = N / R, stride = N / (Ns * R);
    constexpr bool kPow2 = (R & (R - 1)) == 0;
    constexpr int H = (R - 1) / 2;
    float2 wr[R];
    if constexpr (!kPow2) {
#pragma unroll
        for (int m = 0; m < R; ++m) wr[m] = __ldg(twN + m * nb);
    }
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
        const int jq = (Ns == 1) ? j : (int)__umulhi((uint32_t)j, inv_ns);
        const int k = j - jq * Ns, base = jq * Ns * R + k;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float2 x = src[j + r * nb];
            if (r > 0 && Ns > 1) {
                const float2 w = __ldg(twN + k * r * stride);
                x = INV ? cmul_conj(x, w) : cmul(x, w);
            }
            v[r] = x;
        }
        if constexpr (kPow2) {
            fft_dif<R, INV, 1>(v);
            constexpr int L = ilog2(R);
#pragma unroll
            for (int q = 0; q < R; ++q) dst[base + q * Ns] = v[brev(q, L)];
        } else {
            float2 x0 = v[0];
#pragma unroll
            for (int r = 1; r <= H; ++r) {
                const float2 a = v[r], b = v[R - r];
                v[r] = make_float2(a.x + b.x, a.y + b.y);          // s_r
                v[R - r] = make_float2(a.x - b.x, a.y - b.y);      // d_r
                x0.x += v[r].x;
                x0.y += v[r].y;
            }
            dst[base] = x0;
#pragma unroll
            for (int q = 1; q <= H; ++q) {
                float2 A = v[0], B = make_float2(0.f, 0.f);
#pragma unroll
                for (int r = 1; r <= H; ++r) {
                    const float2 w = wr[(r * q) % R];              // (cos, -sin) of 2 pi r q / R
                    A.x = fmaf(w.x, v[r].x, A.x);
                    A.y = fmaf(w.x, v[r].y, A.y);
                    B.x = fmaf(-w.y, v[R - r].y, B.x);             // i w.y d = w.y (-d.y, d.x)
                    B.y = fmaf(w.y, v[R - r].x, B.y);
                }
                const float2 p = make_float2(A.x + B.x, A.y + B.y), m = make_float2(A.x - B.x, A.y - B.y);
                dst[base + q * Ns] = INV ? m : p;
                dst[base + (R - q) * Ns] = INV ? p : m;
            }
        }
    }
}

// DFT of the row in `a` using `b` as the other buffer; returns the buffer that holds the result.
// Ends with a __syncthreads().
template <bool INV>
__device__ float2* mixed_dft(const GenDev& g, float2* a, float2* b) {
    int Ns = 1;
    for (int p = 0; p < g.npass; ++p) {
        const int R = g.radix[p];
        switch (R) {
            case 2: mixed_pass<2, INV>(a, b, g.N, Ns, g.inv_ns[p], g.twN); break;
            case 3: mixed_pass<3, INV>(a, b, g.N, Ns, g.inv_ns[p], g.twN); break;
            case 4: mixed_pass<4, INV>(a, b, g.N, Ns, g.inv_ns[p], g.twN); break;
            case 5: mixed_pass<5, INV>(a, b, g.N, Ns, g.inv_ns[p], g.twN); break;
            case 7: mixed_pass<7, INV>(a, b, g.N, Ns, g.inv_ns[p], g.twN); break;
            case 8: mixed_pass<8, INV>(a, b, g.N, Ns, g.inv_ns[p], g.twN); break;
            case 11: mixed_pass<11, INV>(a, b, g.N, Ns, g.inv_ns[p], g.twN); break;
            case 13: mixed_pass<13, INV>(a, b, g.N, Ns, g.inv_ns[p], g.twN); break;
            default: mixed_pass<16, INV>(a, b, g.N, Ns, g.inv_ns[p], g.twN); break;
        }
        __syncthreads();
        Ns *= R;
        float2* t = a; a = b; b = t;
    }
    return a;
}

template <int MODE>
__global__ void k_row_mixed(GenDev g, float2* __restrict__ data, int64_t pitch, int n_rows,
                            const RowCoef* __restrict__ coef, float scale, double* __restrict__ max_sq) {
    extern __shared__ float2 sm[];
    float2* b0 = sm;
    float2* b1 = sm + g.N;
    double mx = 0.0;
    for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
        float2* p = data + (int64_t)row * pitch;
        RowCoef rc{};
        if (MODE == RANGE) rc = coef[row];
        for (int n = threadIdx.x; n < g.N; n += blockDim.x) {
            float2 x = p[n];
            if (MODE == RANGE) x = cmul(x, cis_u64(quad_phase(rc.a1, rc.b1, rc.c1, (uint32_t)n)));
            b0[n] = x;
        }
        __syncthreads();
        float2* cur = (MODE == AZ_INV) ? mixed_dft<true>(g, b0, b1) : mixed_dft<false>(g, b0, b1);
        if (MODE == RANGE) {
            float2* other = (cur == b0) ? b1 : b0;
            for (int k = threadIdx.x; k < g.N; k += blockDim.x)
                cur[k] = cmul(cur[k], cis_u64(phi2_phase(rc, (uint32_t)k, (uint32_t)g.N)));
            __syncthreads();
            cur = mixed_dft<true>(g, cur, other);
            for (int n = threadIdx.x; n < g.N; n += blockDim.x)
                p[n] = cmul(cur[n], cis_u64(quad_phase(rc.a3, rc.b3, rc.c3, (uint32_t)n)));
        } else {
            for (int n = threadIdx.x; n < g.N; n += blockDim.x) {
                float2 x = cur[n];
                if (MODE == AZ_INV) {
                    x.x *= scale; x.y *= scale;
                    if (max_sq != nullptr) mx = fmax(mx, sq_mag_f64(x));
                }
                p[n] = x;
            }
        }
        __syncthreads();
    }
    if (MODE == AZ_INV && max_sq != nullptr) atomic_max_f64(max_sq, warp_max_f64(mx));
}

// ------------------------------------------------------------------------------- Bluestein engine
// On entry v[s] = x[idx] a[idx] (idx = t + NT s < N, else 0) for the forward transform (conj(a) for
// the inverse); on exit v[s] = X[idx] for idx < N.
template <class P, int PAD, bool INV>
__device__ __forceinline__ void bluestein_core(float2* v, int t, float2* sm, const GenDev& g) {
    constexpr int E = P::E, NT = P::NT;
    transform<P, false, 1, PAD>(v, t, sm, g.tw_pow2);
#pragma unroll
    for (int s = 0; s < E; ++s) {
        const float2 h = __ldg(g.bfft + t + NT * s);
        v[s] = INV ? cmul_conj(v[s], h) : cmul(v[s], h);
    }
    __syncthreads();
    transform<P, true, 1, PAD>(v, t, sm, g.tw_pow2);
#pragma unroll
    for (int s = 0; s < E; ++s) {
        const int idx = t + NT * s;
        if (idx < g.N) {
            const float2 a = __ldg(g.chirp + idx);
            v[s] = INV ? cmul_conj(v[s], a) : cmul(v[s], a);
        }
    }
    __syncthreads();
}

// The same core with ONE copy of the transform in the instruction stream (the inverse transform runs as
// conj(FFT(conj(.))), the direction of the DFT is a run-time sign): for the 32-samples-per-thread 16384-point plan, whose
// kernels inlined two to four transform bodies and spilled 460-1900 bytes per thread at the plan's 128-register cap.
template <class P, int PAD>
__device__ __forceinline__ void bluestein_core_rolled(float2* v, int t, float2* sm, const GenDev& g, bool inv) {
    constexpr int E = P::E, NT = P::NT;
    const float sg = inv ? -1.0f : 1.0f;
#pragma unroll 1
    for (int step = 0; step < 2; ++step) {
        transform<P, false, 1, PAD>(v, t, sm, g.tw_pow2);
        if (step == 0) {   // v <- conj(v h), h conjugated for the inverse DFT
#pragma unroll
            for (int s = 0; s < E; ++s) {
                float2 h = __ldg(g.bfft + t + NT * s);
                h.y *= sg;
                const float2 w = cmul(v[s], h);
                v[s] = make_float2(w.x, -w.y);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int s = 0; s < E; ++s) {
        const int idx = t + NT * s;
        if (idx < g.N) {
            float2 a = __ldg(g.chirp + idx);
            a.y *= sg;
            v[s] = cmul(make_float2(v[s].x, -v[s].y), a);
        }
    }
}

template <int MODE, class P, int PAD>
__global__ void __launch_bounds__(P::NT) k_row_blue(GenDev g, float2* __restrict__ data, int64_t pitch, int n_rows,
                                                    const RowCoef* __restrict__ coef, float scale,
                                                    double* __restrict__ max_sq) {
    extern __shared__ float2 sm[];
    constexpr int E = P::E, NT = P::NT;
    constexpr bool ROLLED = P::E >= 32;
    const int t = threadIdx.x;
    double mx = 0.0;
    for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
        float2* p = data + (int64_t)row * pitch;
        RowCoef rc{};
        if (MODE == RANGE) rc = coef[row];
        float2 v[E];
#pragma unroll
        for (int s = 0; s < E; ++s) {
            const int idx = t + NT * s;
            float2 x = make_float2(0.f, 0.f);
            if (idx < g.N) {
                x = p[idx];
                if (MODE == RANGE) x = cmul(x, cis_u64(quad_phase(rc.a1, rc.b1, rc.c1, (uint32_t)idx)));
                const float2 a = __ldg(g.chirp + idx);
                x = (MODE == AZ_INV) ? cmul_conj(x, a) : cmul(x, a);
            }
            v[s] = x;
        }
        if constexpr (ROLLED) {
            // forward DFT (or the inverse one for AZ_INV); in RANGE mode a second trip: x Phi2, inverse DFT
#pragma unroll 1
            for (int pass = 0; pass < (MODE == RANGE ? 2 : 1); ++pass) {
                if (MODE == RANGE && pass == 1) {
#pragma unroll
                    for (int s = 0; s < E; ++s) {
                        const int idx = t + NT * s;
                        float2 x = make_float2(0.f, 0.f);
                        if (idx < g.N) {
                            x = cmul(v[s], cis_u64(phi2_phase(rc, (uint32_t)idx, (uint32_t)g.N)));
                            x = cmul_conj(x, __ldg(g.chirp + idx));
                        }
                        v[s] = x;
                    }
                }
                bluestein_core_rolled<P, PAD>(v, t, sm, g, MODE == AZ_INV || pass == 1);
            }
        } else {
            if (MODE == AZ_INV) bluestein_core<P, PAD, true>(v, t, sm, g);
            else bluestein_core<P, PAD, false>(v, t, sm, g);
            if (MODE == RANGE) {
#pragma unroll
                for (int s = 0; s < E; ++s) {
                    const int idx = t + NT * s;
                    float2 x = make_float2(0.f, 0.f);
                    if (idx < g.N) {
                        x = cmul(v[s], cis_u64(phi2_phase(rc, (uint32_t)idx, (uint32_t)g.N)));
                        x = cmul_conj(x, __ldg(g.chirp + idx));
                    }
                    v[s] = x;
                }
                bluestein_core<P, PAD, true>(v, t, sm, g);
            }
        }
#pragma unroll
        for (int s = 0; s < E; ++s) {
            const int idx = t + NT * s;
            if (idx < g.N) {
                float2 x = v[s];
                if (MODE == RANGE) x = cmul(x, cis_u64(quad_phase(rc.a3, rc.b3, rc.c3, (uint32_t)idx)));
                if (MODE == AZ_INV) {
                    x.x *= scale; x.y *= scale;
                    if (max_sq != nullptr) mx = fmax(mx, sq_mag_f64(x));
                }
                p[idx] = x;
            }
        }
    }
    if (MODE == AZ_INV && max_sq != nullptr) atomic_max_f64(max_sq, warp_max_f64(mx));
}

// Pruned Bluestein for N <= H = M/2: the product sequence x a is zero on [N, M) and only outputs [0, N) are wanted, so the
// M-point circular convolution splits into even / odd spectral bins -- two H-point transforms of u and u w_M^n, two H-point
// inverses A, B, y[n] = A[n] + w_M^-n B[n] -- on the 16-elements-per-thread plan instead of the 32-element one.  u and A
// are parked in shared memory (each thread re-reads only its own elements).
template <class P, int PAD, bool INV, bool FLY>
__device__ __forceinline__ void bluestein_core_pruned(float2* v, int t, float2* sm, float2* park_u, float2* park_a,
                                                      const GenDev& g) {
    constexpr int E = P::E, NT = P::NT;
#pragma unroll
    for (int s = 0; s < E; ++s) park_u[t + NT * s] = v[s];
    // even bins, then odd bins, as a rolled loop: one copy of the two transforms in the instruction stream
#pragma unroll 1
    for (int br = 0; br < 2; ++br) {
        const float2* __restrict__ bf = br ? g.bfo : g.bfe;
        transform<P, false, 1, PAD>(v, t, sm, g.tw_half);
#pragma unroll
        for (int s = 0; s < E; ++s) {
            const float2 h = __ldg(bf + t + NT * s);
            v[s] = INV ? cmul_conj(v[s], h) : cmul(v[s], h);
        }
        __syncthreads();
        transform<P, true, 1, PAD>(v, t, sm, g.tw_half);
        if (br == 0) {
#pragma unroll
            for (int s = 0; s < E; ++s) {
                park_a[t + NT * s] = v[s];
                v[s] = cmul(park_u[t + NT * s], FLY ? blue_twm(g, (uint32_t)(t + NT * s)) : __ldg(g.twm + t + NT * s));
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int s = 0; s < E; ++s) {
        const int idx = t + NT * s;
        float2 y = make_float2(0.f, 0.f);
        if (idx < g.N) {
            const float2 c = cmul_conj(v[s], FLY ? blue_twm(g, (uint32_t)idx) : __ldg(g.twm + idx));
            const float2 a = park_a[idx];
            y = make_float2(a.x + c.x, a.y + c.y);
            const float2 ch = FLY ? blue_chirp(g, (uint32_t)idx) : __ldg(g.chirp + idx);
            y = INV ? cmul_conj(y, ch) : cmul(y, ch);
        }
        v[s] = y;
    }
    __syncthreads();
}

// azimuth transforms (AZ_FWD / AZ_INV) of rows whose length takes the pruned Bluestein core
template <int MODE, class P, int PAD, bool FLY>
__global__ void __launch_bounds__(P::NT) k_row_blue_pruned(GenDev g, float2* __restrict__ data, int64_t pitch, int n_rows,
                                                           float scale, double* __restrict__ max_sq) {
    extern __shared__ float2 sm[];
    constexpr int E = P::E, NT = P::NT, H = P::N;
    constexpr int SMROW = H + (PAD ? (H >> PAD) : 0);
    float2* const park_u = sm + SMROW;
    float2* const park_a = park_u + H;
    const int t = threadIdx.x;
    double mx = 0.0;
    for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
        float2* p = data + (int64_t)row * pitch;
        float2 v[E];
#pragma unroll
        for (int s = 0; s < E; ++s) {
            const int idx = t + NT * s;
            float2 x = make_float2(0.f, 0.f);
            if (idx < g.N) {
                const float2 a = FLY ? blue_chirp(g, (uint32_t)idx) : __ldg(g.chirp + idx);
                x = (MODE == AZ_INV) ? cmul_conj(p[idx], a) : cmul(p[idx], a);
            }
            v[s] = x;
        }
        if (MODE == AZ_INV) bluestein_core_pruned<P, PAD, true, FLY>(v, t, sm, park_u, park_a, g);
        else bluestein_core_pruned<P, PAD, false, FLY>(v, t, sm, park_u, park_a, g);
#pragma unroll
        for (int s = 0; s < E; ++s) {
            const int idx = t + NT * s;
            if (idx < g.N) {
                float2 x = v[s];
                if (MODE == AZ_INV) {
                    x.x *= scale; x.y *= scale;
                    if (max_sq != nullptr) mx = fmax(mx, sq_mag_f64(x));
                }
                p[idx] = x;
            }
        }
    }
    if (MODE == AZ_INV && max_sq != nullptr) atomic_max_f64(max_sq, warp_max_f64(mx));
}

// --------------------------------------------------------------------------------------- host side
using P64 = Plan<64, 8, 8, 8, 1>;
using P256 = Plan<256, 16, 16, 16, 1>;
using P1024 = Plan<1024, 16, 16, 8, 8>;
using P4096 = Plan<4096, 16, 16, 16, 16>;
using P16384 = Plan<16384, 32, 32, 32, 16>;
// half-length transforms of the pruned Bluestein core: 16 samples per thread, four passes; NIS_BLUE_PLAN=e32 selects the
// 32-sample three-pass plan of the 8192-sample range kernel (development knob)
using P8192H = Plan<8192, 16, 16, 8, 8, 8>;
using P8192H32 = Plan<8192, 32, 32, 16, 16>;

void host_fft(std::vector<std::complex<double>>& a) {   // in-place radix-2, forward
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    const double two_pi = 6.283185307179586476925286766559;
    for (size_t len = 2; len <= n; len <<= 1) {
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const double ang = -two_pi * (double)k / (double)len;
                const std::complex<double> w(cos(ang), sin(ang));
                const std::complex<double> u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
}

int upload(const std::vector<float2>& h, float2** dev) {
    NIS_CUDA_TRY(cudaMalloc(dev, (h.size() + 1) * sizeof(float2)));
    NIS_CUDA_TRY(cudaMemcpy(*dev, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
    return NIS_OK;
}

template <class P>
int upload_pow2_twiddles(float2** dev) {
    std::vector<float2> h(P::tw_len + 1);
    build_twiddles<P>(h.data());
    return upload(h, dev);
}

template <class MP>
int upload_ct_tables(float2** dev) {
    std::vector<float2> h(MP::TW_LEN);
    mixedct::build_tables<MP>(h.data());
    return upload(h, dev);
}

template <int MODE, class MP>
int launch_mixed_ct(nis_ctx* ctx, const GenLen& g, float2* data, int64_t pitch, int n_rows, const RowCoef* coef,
                    float scale, double* max_sq, cudaStream_t st) {
    static bool attr_done_dev[64] = {};
    bool& attr_done = attr_done_dev[nis::current_device() & 63];
    if (!attr_done) {
        NIS_CUDA_TRY(cudaFuncSetAttribute(mixedct::k_row_mixed_ct<MODE, MP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)MP::smem_bytes));
        attr_done = true;
    }
    int grid = ctx->num_sms;
    if (grid > n_rows) grid = n_rows;
    mixedct::k_row_mixed_ct<MODE, MP><<<grid, MP::NT, MP::smem_bytes, st>>>(data, pitch, n_rows, coef, g.ct_tw, scale, max_sq);
    NIS_LAUNCH_CHECK(ctx);
    return NIS_OK;
}

int build_length(int n, GenLen* g) {
    g->N = n;
    const double pi = 3.14159265358979323846264338327950288;
    if (smooth(n)) {
        g->kind = 0;
        factorize(n, g->radix, &g->npass);
        std::vector<float2> tw(n);
        for (int m = 0; m < n; ++m) {
            const double a = -2.0 * pi * (double)m / (double)n;
            tw[m] = make_float2((float)cos(a), (float)sin(a));
        }
        int rc = NIS_OK;
        if (n == mixedct::MP13200::N) rc = upload_ct_tables<mixedct::MP13200>(&g->ct_tw);
        if (n == mixedct::MP7200::N) rc = upload_ct_tables<mixedct::MP7200>(&g->ct_tw);
        if (rc != NIS_OK) return rc;
        return upload(tw, &g->twN);
    }
    g->kind = 1;
    g->M = bluestein_m(n);
    const int M = g->M;
    std::vector<float2> chirp(n);
    std::vector<std::complex<double>> b(M, 0.0);
    for (int i = 0; i < n; ++i) {
        const long long q = ((long long)i * i) % (2LL * n);   // n^2 mod 2N keeps the argument exact
        const double a = pi * (double)q / (double)n;
        chirp[i] = make_float2((float)cos(a), (float)-sin(a));        // a[n] = exp(-i pi n^2 / N)
        const std::complex<double> bv(cos(a), sin(a));                // b[n] = conj(a[n])
        b[i] = bv;
        if (i > 0) b[M - i] = bv;
    }
    host_fft(b);
    std::vector<float2> bf(M);
    for (int i = 0; i < M; ++i) bf[i] = make_float2((float)(b[i].real() / M), (float)(b[i].imag() / M));
    int rc;
    if ((rc = upload(chirp, &g->chirp)) != NIS_OK) return rc;
    if ((rc = upload(bf, &g->bfft)) != NIS_OK) return rc;
    if (M == 16384 && n <= 8192) {
        const int H = M / 2;
        std::vector<float2> be(H), bo(H), tm(H);
        for (int k = 0; k < H; ++k) {
            be[k] = bf[2 * k];
            bo[k] = bf[2 * k + 1];
            const double a = -2.0 * pi * (double)k / (double)M;
            tm[k] = make_float2((float)cos(a), (float)sin(a));
        }
        if ((rc = upload(be, &g->bfe)) != NIS_OK) return rc;
        if ((rc = upload(bo, &g->bfo)) != NIS_OK) return rc;
        if ((rc = upload(tm, &g->twm)) != NIS_OK) return rc;
        if ((rc = upload_pow2_twiddles<P8192H>(&g->tw_half)) != NIS_OK) return rc;
        if ((rc = upload_pow2_twiddles<P8192H32>(&g->tw_half32)) != NIS_OK) return rc;
    }
    switch (M) {
        case 64: return upload_pow2_twiddles<P64>(&g->tw_pow2);
        case 256: return upload_pow2_twiddles<P256>(&g->tw_pow2);
        case 1024: return upload_pow2_twiddles<P1024>(&g->tw_pow2);
        case 4096: return upload_pow2_twiddles<P4096>(&g->tw_pow2);
        default: return upload_pow2_twiddles<P16384>(&g->tw_pow2);
    }
}

template <int MODE, class P, int PAD>
int launch_blue(nis_ctx* ctx, const GenLen& g, float2* data, int64_t pitch, int n_rows, const RowCoef* coef,
                float scale, double* max_sq, cudaStream_t st) {
    constexpr int SMROW = P::N + (PAD ? (P::N >> PAD) : 0);
    const size_t smem = (size_t)SMROW * sizeof(float2);
    static bool attr_done_dev[64] = {};
    bool& attr_done = attr_done_dev[nis::current_device() & 63];
    if (!attr_done) {
        NIS_CUDA_TRY(cudaFuncSetAttribute(k_row_blue<MODE, P, PAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    int per_sm = 1;
    NIS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_row_blue<MODE, P, PAD>, P::NT, smem));
    int grid = ctx->num_sms * (per_sm < 1 ? 1 : per_sm);
    if (grid > n_rows) grid = n_rows;
    k_row_blue<MODE, P, PAD><<<grid, P::NT, smem, st>>>(g.dev(), data, pitch, n_rows, coef, scale, max_sq);
    NIS_LAUNCH_CHECK(ctx);
    return NIS_OK;
}

template <int MODE, class P, int PAD, bool FLY>
int launch_blue_pruned(nis_ctx* ctx, const GenLen& g, float2* data, int64_t pitch, int n_rows, float scale, double* max_sq,
                       cudaStream_t st) {
    auto kern = k_row_blue_pruned<MODE, P, PAD, FLY>;
    constexpr int SMROW = P::N + (P::N >> PAD);
    const size_t smem = (size_t)(SMROW + 2 * P::N) * sizeof(float2);
    static bool attr_done_dev[64] = {};
    bool& attr_done = attr_done_dev[nis::current_device() & 63];
    if (!attr_done) {
        NIS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    int grid = ctx->num_sms;
    if (grid > n_rows) grid = n_rows;
    GenDev d = g.dev();
    if (P::E == 32) d.tw_half = g.tw_half32;
    kern<<<grid, P::NT, smem, st>>>(d, data, pitch, n_rows, scale, max_sq);
    NIS_LAUNCH_CHECK(ctx);
    return NIS_OK;
}

template <int MODE>
int launch_row(nis_ctx* ctx, const GenLen& g, float2* data, int64_t pitch, int n_rows, const RowCoef* coef,
               float scale, double* max_sq, cudaStream_t st) {
    if (g.kind == 1 && g.bfe != nullptr && MODE != RANGE && !getenv("NIS_BLUE_NOPRUNE")) {
        const char* bp = getenv("NIS_BLUE_PLAN");
        if (bp && bp[0] == 'e' && bp[1] == '3') return launch_blue_pruned<MODE, P8192H32, 5, true>(ctx, g, data, pitch, n_rows, scale, max_sq, st);
        if (getenv("NIS_BLUE_TABLES")) return launch_blue_pruned<MODE, P8192H, 4, false>(ctx, g, data, pitch, n_rows, scale, max_sq, st);
        return launch_blue_pruned<MODE, P8192H, 4, true>(ctx, g, data, pitch, n_rows, scale, max_sq, st);
    }
    if (g.kind == 1) {
        switch (g.M) {
            case 64: return launch_blue<MODE, P64, 3>(ctx, g, data, pitch, n_rows, coef, scale, max_sq, st);
            case 256: return launch_blue<MODE, P256, 4>(ctx, g, data, pitch, n_rows, coef, scale, max_sq, st);
            case 1024: return launch_blue<MODE, P1024, 4>(ctx, g, data, pitch, n_rows, coef, scale, max_sq, st);
            case 4096: return launch_blue<MODE, P4096, 4>(ctx, g, data, pitch, n_rows, coef, scale, max_sq, st);
            default: return launch_blue<MODE, P16384, 5>(ctx, g, data, pitch, n_rows, coef, scale, max_sq, st);
        }
    }
    if (g.ct_tw != nullptr && !getenv("NIS_MIXED_RUNTIME")) {
        if (g.N == mixedct::MP13200::N)
            return launch_mixed_ct<MODE, mixedct::MP13200>(ctx, g, data, pitch, n_rows, coef, scale, max_sq, st);
        if (g.N == mixedct::MP7200::N)
            return launch_mixed_ct<MODE, mixedct::MP7200>(ctx, g, data, pitch, n_rows, coef, scale, max_sq, st);
    }
    const size_t smem = 2 * (size_t)g.N * sizeof(float2);
    static bool attr_done_dev[64] = {};
    bool& attr_done = attr_done_dev[nis::current_device() & 63];
    if (!attr_done) {
        NIS_CUDA_TRY(cudaFuncSetAttribute(k_row_mixed<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = true;
    }
    int threads = (g.N / 8 + 31) / 32 * 32;
    if (threads < 64) threads = 64;
    if (threads > 512) threads = 512;
    int per_sm = 1;
    NIS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_row_mixed<MODE>, threads, smem));
    int grid = ctx->num_sms * (per_sm < 1 ? 1 : per_sm);
    if (grid > n_rows) grid = n_rows;
    k_row_mixed<MODE><<<grid, threads, smem, st>>>(g.dev(), data, pitch, n_rows, coef, scale, max_sq);
    NIS_LAUNCH_CHECK(ctx);
    return NIS_OK;
}

}  // namespace

namespace nis {
namespace csa {

struct GenericState {
    GenLen az, rg;
};

struct RowDft {
    GenLen g;
};
bool rowdft_supported(int n) { return length_supported(n); }
int rowdft_create(int n, RowDft** out) {
    RowDft* d = new RowDft();
    int rc = build_length(n, &d->g);
    if (rc != NIS_OK) { d->g.release(); delete d; return rc; }
    *out = d;
    return NIS_OK;
}
void rowdft_destroy(RowDft* d) {
    if (!d) return;
    d->g.release();
    delete d;
}
int rowdft_run(nis_ctx* ctx, const RowDft* d, float2* data, int64_t pitch, int n_rows, bool inverse, float scale,
               cudaStream_t st) {
    if (inverse) return launch_row<AZ_INV>(ctx, d->g, data, pitch, n_rows, nullptr, scale, nullptr, st);
    return launch_row<AZ_FWD>(ctx, d->g, data, pitch, n_rows, nullptr, 1.f, nullptr, st);
}

int generic_supported(int n_az, int n_rg) { return length_supported(n_az) && length_supported(n_rg); }

int generic_create(nis_csa_plan* pl) {
    pl->generic = new GenericState();
    int rc;
    if ((rc = build_length(pl->n_az, &pl->generic->az)) != NIS_OK) return rc;
    if ((rc = build_length(pl->n_rg, &pl->generic->rg)) != NIS_OK) return rc;
    return NIS_OK;
}

void generic_destroy(nis_csa_plan* pl) {
    if (!pl->generic) return;
    pl->generic->az.release();
    pl->generic->rg.release();
    delete pl->generic;
    pl->generic = nullptr;
}

int generic_focus(nis_csa_plan* pl, const float2* phist, int64_t pitch, float2* slc, double* max_sq, cudaStream_t st) {
    nis_ctx* ctx = pl->ctx;
    const int n_az = pl->n_az, n_rg = pl->n_rg;
    const float scale = (float)(1.0 / ((double)n_az * (double)n_rg));
    int rc;
    if ((rc = launch_transpose(ctx, phist, pitch, slc, n_az, n_rg, st)) != NIS_OK) return rc;
    if ((rc = launch_row<AZ_FWD>(ctx, pl->generic->az, slc, n_az, n_rg, nullptr, 1.f, nullptr, st)) != NIS_OK) return rc;
    if ((rc = launch_transpose(ctx, slc, n_az, pl->work, n_rg, n_az, st)) != NIS_OK) return rc;
    if ((rc = launch_row<RANGE>(ctx, pl->generic->rg, pl->work, n_rg, n_az, pl->coef, 1.f, nullptr, st)) != NIS_OK) return rc;
    if ((rc = launch_transpose(ctx, pl->work, n_rg, slc, n_az, n_rg, st)) != NIS_OK) return rc;
    if ((rc = launch_row<AZ_INV>(ctx, pl->generic->az, slc, n_az, n_rg, nullptr, scale, max_sq, st)) != NIS_OK) return rc;
    return NIS_OK;
}

}  // namespace csa
}  // namespace nis
