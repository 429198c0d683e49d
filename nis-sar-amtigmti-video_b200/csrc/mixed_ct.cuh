// mixed_ct.cuh -- compile-time mixed-radix row transforms for the two non-power-of-two lengths the reference's default scenes
// produce: 13200 range samples (int(22e-6 * 600e6), sar_ati_dcpa_sim_csa.py:111) and 7200 pulses (ceil(1.2 * 6000),
// sar_satellite_sim.py:84-85).  Other smooth lengths stay on the run-time engine of csa_generic.cu.
//
// A length N = R0 R1 R2 R3 is transformed by four Stockham passes.  Pass p gathers x[j + r N/Rp] (r < Rp) for butterfly
// j < N/Rp, multiplies by w^(k r) (k = j mod Ns, Ns = R0..R(p-1), w = exp(-2 pi i / (Ns Rp))), takes an Rp-point DFT and
// scatters to (j - k) Rp + k + q Ns.  The inverse runs the TRANSPOSED passes in the order 3, 2, 1, 0 with conjugated
// constants (the DFT matrix is symmetric), so that
//   * both directions use the same per-pass twiddle tables, laid out [r - 1][k] (consecutive k: conflict-free LDS.64);
//   * the first forward pass reads, and the last inverse pass writes, x[j + r N/R0]: straight from / to global memory,
//     coalesced, with Phi1 / Phi3 applied on the way;
//   * forward pass 3 leaves X[j + q N/R3] in the registers of thread j, which is exactly what transposed pass 3 gathers:
//     range FFT -> x Phi2 -> inverse range FFT joins in registers (one shared-memory round trip saved, no barrier).
// Two row buffers (2 N float2) + the tables live in shared memory: a pass reads one buffer and writes the other, one barrier
// per pass and nothing held in registers across it.  For 13200 the pass-3 table keeps only w^k; w^(k r) is formed by
// squarings / multiplications (depth <= 5), which is cheaper than eleven more shared-memory loads and lets both buffers fit.
// Small DFTs: radix 2/4/8/16 from fft.cuh; odd radices through the conjugate-symmetric pairs with compile-time roots (FFMA
// immediates); 10 = 2 x 5, 12 = 4 x 3 ... by the prime-factor (Good-Thomas) map, no twiddles.
#pragma once
#include <type_traits>

#include "csa_internal.cuh"
#include "fft.cuh"

namespace nis {
namespace mixedct {

using namespace nis::fft;
using nis::csa::RowCoef;

// ---- compile-time roots of unity
constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double cx_sin(double x) {   // |x| <= pi
    double term = x, sum = x;
    for (int k = 1; k < 24; ++k) {
        term *= -x * x / (double)((2 * k) * (2 * k + 1));
        sum += term;
    }
    return sum;
}
constexpr double cx_cos(double x) {
    double term = 1.0, sum = 1.0;
    for (int k = 1; k < 24; ++k) {
        term *= -x * x / (double)((2 * k - 1) * (2 * k));
        sum += term;
    }
    return sum;
}
template <int R, int M>
struct Root {   // exp(-2 pi i M / R), argument folded into (-pi, pi]
    static constexpr int m = ((M % R) + R) % R;
    static constexpr double ang = 2.0 * kPi * (double)(m > R / 2 ? m - R : m) / (double)R;
    static constexpr float c = (float)cx_cos(ang);
    static constexpr float s = (float)cx_sin(ang);   // sin(2 pi m / R); the root is (c, -s)
};

template <int B, int E, class F>
__host__ __device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

// ---- small DFTs, in place, natural order.  STRIDE lets the prime-factor map run sub-transforms on strided registers.
template <int R, bool INV, int STRIDE>
__host__ __device__ __forceinline__ void dft_odd(float2* v) {
    constexpr int H = (R - 1) / 2;
    float2 s[H + 1], d[H + 1];
    float2 x0 = v[0];
#pragma unroll
    for (int r = 1; r <= H; ++r) {
        const float2 a = v[r * STRIDE], b = v[(R - r) * STRIDE];
        s[r] = make_float2(a.x + b.x, a.y + b.y);
        d[r] = make_float2(a.x - b.x, a.y - b.y);
        x0.x += s[r].x;
        x0.y += s[r].y;
    }
    const float2 v0 = v[0];
    v[0] = x0;
    static_for<1, H + 1>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        float2 A = v0, B = make_float2(0.f, 0.f);
        static_for<1, H + 1>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            constexpr float c = Root<R, r * q>::c, sn = Root<R, r * q>::s;
            A.x = fmaf(c, s[r].x, A.x);
            A.y = fmaf(c, s[r].y, A.y);
            B.x = fmaf(sn, d[r].y, B.x);      // -i sin (d) = sin (d.y, -d.x)
            B.y = fmaf(-sn, d[r].x, B.y);
        });
        const float2 p = make_float2(A.x + B.x, A.y + B.y), m = make_float2(A.x - B.x, A.y - B.y);
        v[q * STRIDE] = INV ? m : p;
        v[(R - q) * STRIDE] = INV ? p : m;
    });
}

template <int R, bool INV, int STRIDE>
__host__ __device__ __forceinline__ void dft_pow2(float2* v) {
    if constexpr (R == 2) {
        const float2 a = v[0], b = v[STRIDE];
        v[0] = make_float2(a.x + b.x, a.y + b.y);
        v[STRIDE] = make_float2(a.x - b.x, a.y - b.y);
    } else {
        fft_dif<R, INV, STRIDE>(v);
        constexpr int L = ilog2(R);
        float2 t[R];
#pragma unroll
        for (int q = 0; q < R; ++q) t[q] = v[brev(q, L) * STRIDE];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q * STRIDE] = t[q];
    }
}

constexpr bool is_pow2(int r) { return (r & (r - 1)) == 0; }
constexpr int pow2_part(int r) { return r & (-r); }

template <int R, bool INV, int STRIDE = 1>
__host__ __device__ __forceinline__ void dft_small(float2* v) {
    if constexpr (is_pow2(R)) {
        dft_pow2<R, INV, STRIDE>(v);
    } else if constexpr (R % 2 == 1) {
        dft_odd<R, INV, STRIDE>(v);
    } else {
        // R = R1 R2, R1 = 2^a, R2 odd (coprime): n = (R2 n1 + R1 n2) mod R in, k = k1 (mod R1) = k2 (mod R2) out
        constexpr int R1 = pow2_part(R), R2 = R / R1;
        float2 a[R];
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1)
#pragma unroll
            for (int n2 = 0; n2 < R2; ++n2) a[n1 * R2 + n2] = v[((R2 * n1 + R1 * n2) % R) * STRIDE];
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) dft_odd<R2, INV, 1>(a + n1 * R2);
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) dft_pow2<R1, INV, R2>(a + k2);
#pragma unroll
        for (int k = 0; k < R; ++k) v[k * STRIDE] = a[(k % R1) * R2 + (k % R2)];
    }
}

// ---- plan.  POW3: pass 3 keeps only w^k (k < NS3) and forms w^(k r) by squaring / multiplying (the full [r-1][k] table of
// pass 3 is about N entries -- with it a 13200-point row could not have two row buffers in 227 KB).
template <int N_, int NT_, int R0_, int R1_, int R2_, int R3_, bool POW3_>
struct MPlan {
    static constexpr int N = N_, NT = NT_, R0 = R0_, R1 = R1_, R2 = R2_, R3 = R3_;
    static constexpr bool POW3 = POW3_;
    static_assert(R0_ * R1_ * R2_ * R3_ == N_, "radices must multiply to N");
    static_assert(R0_ % 2 == 1, "an odd first radix keeps the stride-R0 shared-memory accesses conflict-free");
    static constexpr int NS1 = R0_, NS2 = R0_ * R1_, NS3 = R0_ * R1_ * R2_;
    static constexpr int TW1 = 0, TW2 = TW1 + (R1_ - 1) * NS1, TW3 = TW2 + (R2_ - 1) * NS2,
                         TW_LEN = TW3 + (POW3_ ? 1 : (R3_ - 1)) * NS3;
    static constexpr size_t smem_bytes = (size_t)(2 * N_ + TW_LEN) * sizeof(float2);
    static_assert(smem_bytes <= 227 * 1024, "two row buffers + tables must fit the CTA's shared memory");
};
#ifndef MP13200_NT
#define MP13200_NT 672
#endif
#ifndef MP7200_NT
#define MP7200_NT 480
#endif
using MP13200 = MPlan<13200, MP13200_NT, 11, 10, 10, 12, true>;
using MP7200 = MPlan<7200, MP7200_NT, 9, 10, 10, 8, false>;

template <class MP>
void build_tables(float2* h) {   // host: [pass 1 | pass 2 | pass 3], each [r - 1][k]
    const int R[4] = {MP::R0, MP::R1, MP::R2, MP::R3};
    int ns = R[0], o = 0;
    for (int p = 1; p < 4; ++p) {
        const int rmax = (p == 3 && MP::POW3) ? 2 : R[p];
        for (int r = 1; r < rmax; ++r)
            for (int k = 0; k < ns; ++k) {
                const double a = -2.0 * kPi * (double)((long long)k * r % ((long long)ns * R[p])) / (double)((long long)ns * R[p]);
                h[o++] = make_float2((float)cos(a), (float)sin(a));
            }
        ns *= R[p];
    }
}

// ---- passes.  R = radix, NS = product of the earlier radices, NB = N / R butterflies, IT per thread.
template <class MP, int R, int NS, bool POW>
struct Pass {
    static constexpr int R_ = R, NS_ = NS;
    static constexpr int NB = MP::N / R, IT = (NB + MP::NT - 1) / MP::NT;
    static constexpr bool FULL = (NB % MP::NT) == 0;
    __host__ __device__ static __forceinline__ bool active(int j) { return FULL || j < NB; }

    // v[r] *= w^(k r) (CONJ: conjugated)
    template <bool CONJ>
    __host__ __device__ static __forceinline__ void twiddle(float2* v, const float2* __restrict__ tw, int k) {
        if constexpr (NS > 1 && !POW) {
#pragma unroll
            for (int r = 1; r < R; ++r) {
                const float2 w = tw[(r - 1) * NS + k];
                v[r] = CONJ ? cmul_conj(v[r], w) : cmul(v[r], w);
            }
        } else if constexpr (NS > 1) {
            float2 w[R];
            w[1] = tw[k];
#pragma unroll
            for (int r = 2; r < R; ++r) {
                if (r % 2 == 0) {
                    const float2 h = w[r / 2];
                    w[r] = make_float2(fmaf(h.x, h.x, -h.y * h.y), 2.f * h.x * h.y);
                } else {
                    w[r] = cmul(w[r - 1], w[1]);
                }
            }
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = CONJ ? cmul_conj(v[r], w[r]) : cmul(v[r], w[r]);
        }
    }
    __host__ __device__ static __forceinline__ void gather(float2* v, const float2* sm, int j) {
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = sm[j + r * NB];
    }
    __host__ __device__ static __forceinline__ void spread(const float2* v, float2* sm, int j) {
#pragma unroll
        for (int r = 0; r < R; ++r) sm[j + r * NB] = v[r];
    }
    __host__ __device__ static __forceinline__ void gather_t(float2* v, const float2* sm, int j, int k) {
        const int base = (j - k) * R + k;
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = sm[base + q * NS];
    }
    __host__ __device__ static __forceinline__ void spread_t(const float2* v, float2* sm, int j, int k) {
        const int base = (j - k) * R + k;
#pragma unroll
        for (int q = 0; q < R; ++q) sm[base + q * NS] = v[q];
    }

    // shared -> the other shared buffer, forward: gather, twiddle, DFT, scatter | barrier
    __device__ static __forceinline__ void fwd_mid(const float2* src, float2* dst, const float2* __restrict__ tw, int t) {
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            const int j = t + MP::NT * it;
            if (active(j)) {
                const int k = j % NS;
                float2 v[R];
                gather(v, src, j);
                twiddle<false>(v, tw, k);
                dft_small<R, false>(v);
                spread_t(v, dst, j, k);
            }
        }
        __syncthreads();
    }
    // transposed inverse: gather_t, DFT*, twiddle*, spread | barrier
    __device__ static __forceinline__ void inv_mid(const float2* src, float2* dst, const float2* __restrict__ tw, int t) {
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            const int j = t + MP::NT * it;
            if (active(j)) {
                const int k = j % NS;
                float2 v[R];
                gather_t(v, src, j, k);
                dft_small<R, true>(v);
                twiddle<true>(v, tw, k);
                spread(v, dst, j);
            }
        }
        __syncthreads();
    }
};

enum { M_AZ_FWD = 0, M_RANGE = 1, M_AZ_INV = 2 };

// Buffers: every row makes an even number of buffer changes, so the row's first pass may write the buffer the previous
// row's last pass is NOT reading -- rows alternate (a, b) and no barrier separates them.
template <int MODE, class MP>
__global__ void __launch_bounds__(MP::NT, 1)
k_row_mixed_ct(float2* __restrict__ data, int64_t pitch, int n_rows, const RowCoef* __restrict__ coef,
               const float2* __restrict__ tables, float scale, double* __restrict__ max_sq) {
    extern __shared__ float2 smx[];
    __shared__ RowCoef rcs[2];   // the row's phase coefficients stay in shared memory: 20 registers less to carry
    float2* tw = smx + 2 * MP::N;
    using P0 = Pass<MP, MP::R0, 1, false>;
    using P1 = Pass<MP, MP::R1, MP::NS1, false>;
    using P2 = Pass<MP, MP::R2, MP::NS2, false>;
    using P3 = Pass<MP, MP::R3, MP::NS3, MP::POW3>;
    const int t = threadIdx.x;
    for (int i = t; i < MP::TW_LEN; i += MP::NT) tw[i] = tables[i];
    __syncthreads();
    double mx = 0.0;
    int flip = 0;
    for (int row = blockIdx.x; row < n_rows; row += gridDim.x, flip ^= 1) {
        float2* p = data + (int64_t)row * pitch;
        float2* a = smx + (flip ? MP::N : 0);
        float2* b = smx + (flip ? 0 : MP::N);
        if (MODE == M_RANGE && t < (int)(sizeof(RowCoef) / sizeof(uint64_t)))
            reinterpret_cast<uint64_t*>(&rcs[flip])[t] = reinterpret_cast<const uint64_t*>(coef + row)[t];
        const RowCoef& rc = rcs[flip];
        if (MODE != M_AZ_INV) {
            // ---- forward pass 0: global -> DFT -> a[j R0 + q]   (Phi1 is applied in pass 1: the coefficients written above
            //      become visible at the barrier that ends this pass)
#pragma unroll
            for (int it = 0; it < P0::IT; ++it) {
                const int j = t + MP::NT * it;
                if (P0::active(j)) {
                    float2 v[MP::R0];
#pragma unroll
                    for (int r = 0; r < MP::R0; ++r) v[r] = p[j + r * P0::NB];
                    if (MODE == M_RANGE) {
                        RowCoef c0;   // this pass runs before the barrier: read the row's coefficients from global (L1 broadcast)
                        c0.a1 = coef[row].a1; c0.b1 = coef[row].b1; c0.c1 = coef[row].c1;
#pragma unroll
                        for (int r = 0; r < MP::R0; ++r)
                            v[r] = cmul(v[r], cis_u64(csa::quad_phase(c0.a1, c0.b1, c0.c1, (uint32_t)(j + r * P0::NB))));
                    }
                    dft_small<MP::R0, false>(v);
                    P0::spread_t(v, a, j, 0);
                }
            }
            __syncthreads();
            P1::fwd_mid(a, b, tw + MP::TW1, t);
            P2::fwd_mid(b, a, tw + MP::TW2, t);
            // ---- forward pass 3 (k = j): a -> DFT -> registers X[j + q NS3]
#pragma unroll
            for (int it = 0; it < P3::IT; ++it) {
                const int j = t + MP::NT * it;
                if (P3::active(j)) {
                    float2 w[MP::R3];
                    P3::gather(w, a, j);
                    P3::template twiddle<false>(w, tw + MP::TW3, j);
                    dft_small<MP::R3, false>(w);
                    if (MODE == M_RANGE) {
#pragma unroll
                        for (int q = 0; q < MP::R3; ++q)
                            w[q] = cmul(w[q], cis_u64(csa::phi2_phase(rc, (uint32_t)(j + q * MP::NS3), (uint32_t)MP::N)));
                        dft_small<MP::R3, true>(w);
                        P3::template twiddle<true>(w, tw + MP::TW3, j);
                        P3::spread(w, a, j);          // the locations this thread gathered: no barrier in between
                    } else {
#pragma unroll
                        for (int q = 0; q < MP::R3; ++q) p[j + q * MP::NS3] = w[q];
                    }
                }
            }
            if (MODE == M_RANGE) __syncthreads();
        } else {
            // ---- transposed pass 3 from global: X[j + q NS3] -> DFT* -> x w* -> a[j + r NB3]
#pragma unroll
            for (int it = 0; it < P3::IT; ++it) {
                const int j = t + MP::NT * it;
                if (P3::active(j)) {
                    float2 v[MP::R3];
#pragma unroll
                    for (int q = 0; q < MP::R3; ++q) v[q] = p[j + q * MP::NS3];
                    dft_small<MP::R3, true>(v);
                    P3::template twiddle<true>(v, tw + MP::TW3, j);
                    P3::spread(v, a, j);
                }
            }
            __syncthreads();
        }
        if (MODE != M_AZ_FWD) {
            P2::inv_mid(a, b, tw + MP::TW2, t);
            P1::inv_mid(b, a, tw + MP::TW1, t);
            // ---- transposed pass 0: a[j R0 + q] -> DFT* -> (x Phi3 | x scale) -> global [j + r NB0]
#pragma unroll
            for (int it = 0; it < P0::IT; ++it) {
                const int j = t + MP::NT * it;
                if (P0::active(j)) {
                    float2 w[MP::R0];
                    P0::gather_t(w, a, j, 0);
                    dft_small<MP::R0, true>(w);
#pragma unroll
                    for (int r = 0; r < MP::R0; ++r) {
                        float2 x = w[r];
                        if (MODE == M_RANGE) {
                            x = cmul(x, cis_u64(csa::quad_phase(rc.a3, rc.b3, rc.c3, (uint32_t)(j + r * P0::NB))));
                        } else {
                            x.x *= scale;
                            x.y *= scale;
                            if (max_sq != nullptr) mx = fmax(mx, sq_mag_f64(x));
                        }
                        p[j + r * P0::NB] = x;
                    }
                }
            }
        }
    }
    if (MODE == M_AZ_INV && max_sq != nullptr) atomic_max_f64(max_sq, warp_max_f64(mx));
}

}  // namespace mixedct
}  // namespace nis
