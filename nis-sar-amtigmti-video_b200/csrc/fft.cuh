// fft.cuh -- register-resident Stockham FFT building blocks (power-of-two lengths, radix 2..32).
//
// Layout convention (all kernels): a transform of length N is owned by NT = N/E cooperating
// threads; thread t holds E elements in registers, register slot s <-> element index t + NT*s.
// That mapping is independent of the radix, so the output of one transform (natural order) is
// directly the input of the next one -- the range kernel chains FFT -> x Phi2 -> IFFT without
// touching shared memory in between.
//
// A radix-R Stockham pass (Govindaraju et al. formulation):
//     v[r]  = x[j + r N/R] * w^(r k),  k = j mod Ns,  w = exp(-+2 pi i / (Ns R))
//     FFT_R(v)
//     y[(j div Ns) Ns R + k + q Ns] = V[q]
// Between passes the data crosses shared memory once (write scattered, read strided).
#pragma once

#include <cuda_runtime.h>
#include <utility>

#include "common.cuh"

namespace nis {
namespace fft {

__host__ __device__ constexpr float cos32(int i) {
    constexpr float t[32] = {
        1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654757f,
        0.55557023301960229f, 0.38268343236508984f, 0.19509032201612833f, 0.0f, -0.19509032201612833f,
        -0.38268343236508984f, -0.55557023301960229f, -0.70710678118654757f, -0.83146961230254524f,
        -0.92387953251128674f, -0.98078528040323043f, -1.0f, -0.98078528040323043f, -0.92387953251128674f,
        -0.83146961230254524f, -0.70710678118654757f, -0.55557023301960229f, -0.38268343236508984f,
        -0.19509032201612833f, 0.0f, 0.19509032201612833f, 0.38268343236508984f, 0.55557023301960229f,
        0.70710678118654757f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f};
    return t[i & 31];
}
__host__ __device__ constexpr float sin32(int i) { return cos32(i + 24); }  // sin(x) = cos(x - pi/2)

__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }
__host__ __device__ constexpr int brev(int k, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((k >> i) & 1) << (bits - 1 - i);
    return r;
}

// d * exp(-+ 2 pi i I / N)  (forward: minus), with the trivial cases folded at compile time.  PK: packed fp32x2 forms.
template <int N, bool INV, int I, bool PK = false>
__host__ __device__ __forceinline__ float2 twmul(float2 d) {
    constexpr int q = I * (32 / N);  // 0 <= q < 16
    constexpr float h = 0.70710678118654757f;
    if constexpr (q == 0) {
        return d;
    } else if constexpr (q == 8) {
        return INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
    } else if constexpr (PK && q == 4) {     // h (1 -+ i) d = h (d + (-+i) d)
        return cscale_pk(cadd_pk(d, INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x)), h);
    } else if constexpr (PK && q == 12) {    // h (-1 -+ i) d = h ((-+i) d - d)
        return cscale_pk(csub_pk(INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x), d), h);
    } else if constexpr (PK) {
        constexpr float c = cos32(q), s = sin32(q);   // w = c - i s (forward)
        return INV ? cmul_pk(d, make_float2(c, s)) : cmul_conj_pk(d, make_float2(c, s));
    } else if constexpr (q == 4) {
        return INV ? make_float2((d.x - d.y) * h, (d.x + d.y) * h) : make_float2((d.x + d.y) * h, (d.y - d.x) * h);
    } else if constexpr (q == 12) {
        return INV ? make_float2(-(d.x + d.y) * h, (d.x - d.y) * h) : make_float2((d.y - d.x) * h, -(d.x + d.y) * h);
    } else {
        constexpr float c = cos32(q), s = sin32(q);
        return INV ? make_float2(fmaf(d.x, c, -d.y * s), fmaf(d.y, c, d.x * s))
                   : make_float2(fmaf(d.x, c, d.y * s), fmaf(d.y, c, -d.x * s));
    }
}

// PK selects the packed fp32x2 add/sub (one FADD2 per complex add): fewer issue slots, same pipe time.
template <int N, bool INV, int STRIDE, int I, bool PK>
__host__ __device__ __forceinline__ void dif_bfly(float2* v) {
    constexpr int H = N / 2;
    float2 a = v[I * STRIDE], b = v[(I + H) * STRIDE];
    v[I * STRIDE] = PK ? cadd_pk(a, b) : cadd(a, b);
    v[(I + H) * STRIDE] = twmul<N, INV, I, PK>(PK ? csub_pk(a, b) : csub(a, b));
}
template <int N, bool INV, int STRIDE, bool PK, int... I>
__host__ __device__ __forceinline__ void dif_level(float2* v, std::integer_sequence<int, I...>) {
    (dif_bfly<N, INV, STRIDE, I, PK>(v), ...);
}
// In-register decimation-in-frequency FFT over v[0], v[STRIDE], ..., v[(N-1) STRIDE].
// Result X[k] is left at position brev(k).
template <int N, bool INV, int STRIDE, bool PK = false>
__host__ __device__ __forceinline__ void fft_dif(float2* v) {
    if constexpr (N > 1) {
        dif_level<N, INV, STRIDE, PK>(v, std::make_integer_sequence<int, N / 2>{});
        fft_dif<N / 2, INV, STRIDE, PK>(v);
        fft_dif<N / 2, INV, STRIDE, PK>(v + (N / 2) * STRIDE);
    }
}

// ---------------------------------------------------------------------------------------------
// Plan of up to four radix passes R0*R1*R2*R3 = N (trailing radices equal to 1 are skipped).
template <int N_, int E_, int R0_, int R1_, int R2_, int R3_ = 1>
struct Plan {
    static constexpr int N = N_, E = E_, NT = N_ / E_, R0 = R0_, R1 = R1_, R2 = R2_, R3 = R3_;
    static_assert(R0_ * R1_ * R2_ * R3_ == N_, "radices must multiply to N");
    static_assert(E_ >= R0_ && E_ >= R1_ && E_ >= R2_ && E_ >= R3_, "E must hold the largest butterfly");
    static_assert(R3_ == 1 || R2_ > 1, "radices must be packed to the front");
    static_assert(R2_ == 1 || R1_ > 1, "radices must be packed to the front");
    static constexpr int passes = (R3_ > 1) ? 4 : ((R2_ > 1) ? 3 : ((R1_ > 1) ? 2 : 1));
    // twiddle table: pass p >= 1 has (Rp - 1) * (R0..Rp-1) entries
    static constexpr int tw_off1 = 0;
    static constexpr int tw_off2 = (R1_ > 1) ? (R1_ - 1) * R0_ : 0;
    static constexpr int tw_off3 = tw_off2 + ((R2_ > 1) ? (R2_ - 1) * R0_ * R1_ : 0);
    static constexpr int tw_len = tw_off3 + ((R3_ > 1) ? (R3_ - 1) * R0_ * R1_ * R2_ : 0);
};

template <int PADSHIFT>
__host__ __device__ __forceinline__ int padidx(int i) {
    if constexpr (PADSHIFT > 0) return i + (i >> PADSHIFT);
    else return i;
}

// One pass: twiddle, butterflies.  v slots: element t + NT*s.  Ns = product of earlier radices.
template <int E, int NT, int R, int Ns, bool INV, bool PK = false>
__host__ __device__ __forceinline__ void pass_compute(float2* v, int t, const float2* __restrict__ tw) {
    constexpr int B = E / R;  // butterflies per thread
#pragma unroll
    for (int b = 0; b < B; ++b) {
        if constexpr (Ns > 1) {
            const int k = (t + b * NT) & (Ns - 1);
#pragma unroll
            for (int r = 1; r < R; ++r) {
                float2 w = NIS_LDG(tw + (r - 1) * Ns + k);
                if constexpr (PK) v[b + r * B] = INV ? cmul_conj_pk(v[b + r * B], w) : cmul_pk(v[b + r * B], w);
                else v[b + r * B] = INV ? cmul_conj(v[b + r * B], w) : cmul(v[b + r * B], w);
            }
        }
        fft_dif<R, INV, B, PK>(v + b);
    }
}

// Shared-memory index padding i -> i + (i >> PADSHIFT).  All accesses below are "runtime base + compile-time
// offset": padidx(base + c) == padidx(base) + c + (c >> PADSHIFT) whenever c is a multiple of 2^PADSHIFT, and for the
// first pass (Ns == 1, R == 2^PADSHIFT) padidx(j R + q) == j (R + 1) + q.  The static_asserts pin the plans to those cases.
template <int PADSHIFT>
__host__ __device__ constexpr int padoff(int c) {
    return PADSHIFT > 0 ? c + (c >> PADSHIFT) : c;
}

// scatter pass output to shared memory: y[(j div Ns) Ns R + k + q Ns] = V[q]
template <int E, int NT, int R, int Ns, int SMS, int PADSHIFT>
__host__ __device__ __forceinline__ void pass_scatter(const float2* v, int t, float2* sm) {
    constexpr int B = E / R;
    constexpr int LR = ilog2(R);
    constexpr int G = 1 << PADSHIFT;
    static_assert(PADSHIFT == 0 || (Ns == 1 && R <= G) || (Ns % G == 0), "padding rule needs Ns == 1 or Ns % 2^PAD == 0");
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const int j = t + b * NT;
        const int base = ((j & ~(Ns - 1)) * R) + (j & (Ns - 1));
        float2* p = sm + padidx<PADSHIFT>(base) * SMS;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            // Ns == 1: base = j R is a multiple of R, q < R <= 2^PAD never carries into the padded digit
            const int off = (Ns == 1) ? q : padoff<PADSHIFT>(q * Ns);
            p[off * SMS] = v[b + brev(q, LR) * B];
        }
    }
}
template <int E, int NT, int SMS, int PADSHIFT>
__host__ __device__ __forceinline__ void gather_slots(float2* v, int t, const float2* sm) {
    static_assert(PADSHIFT == 0 || NT % (1 << PADSHIFT) == 0 || E == 1, "padding rule needs NT % 2^PAD == 0");
    const float2* p = sm + padidx<PADSHIFT>(t) * SMS;
#pragma unroll
    for (int s = 0; s < E; ++s) v[s] = p[padoff<PADSHIFT>(NT * s) * SMS];
}
// last pass: natural-order result back into canonical slots (pure register renaming)
template <int E, int R>
__host__ __device__ __forceinline__ void pass_unpermute(float2* v) {
    constexpr int B = E / R;
    constexpr int LR = ilog2(R);
    float2 o[E];
#pragma unroll
    for (int b = 0; b < B; ++b)
#pragma unroll
        for (int q = 0; q < R; ++q) o[b + q * B] = v[b + brev(q, LR) * B];
#pragma unroll
    for (int s = 0; s < E; ++s) v[s] = o[s];
}

// Whole transform.  On entry v[s] = x[t + NT s]; on exit v[s] = X[t + NT s] (unnormalised).
// INV is the same pass sequence with conjugated twiddles and inverse butterflies, so forward and
// inverse share one table.
// sm: base pointer of this transform's shared buffer (already offset to this column when SMS > 1).
// All threads of the CTA must call it (it uses __syncthreads()).
struct CtaBarrier {   // default: the whole CTA takes part in the transform
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
struct NamedBarrier {  // a sub-group of the CTA (e.g. one row of a multi-row block): bar.sync id, count
    int id, count;
    __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
};

template <class P, bool INV, int SMS, int PADSHIFT, class Bar = CtaBarrier, bool PK = false>
__device__ __forceinline__ void transform(float2* v, int t, float2* sm, const float2* __restrict__ tw, Bar bar = Bar()) {
    constexpr int E = P::E, NT = P::NT;
    pass_compute<E, NT, P::R0, 1, INV, PK>(v, t, tw);
    if constexpr (P::passes == 1) {
        pass_unpermute<E, P::R0>(v);
    } else {
        pass_scatter<E, NT, P::R0, 1, SMS, PADSHIFT>(v, t, sm);
        bar();
        gather_slots<E, NT, SMS, PADSHIFT>(v, t, sm);
        pass_compute<E, NT, P::R1, P::R0, INV, PK>(v, t, tw + P::tw_off1);
        if constexpr (P::passes == 2) {
            pass_unpermute<E, P::R1>(v);
        } else {
            bar();
            pass_scatter<E, NT, P::R1, P::R0, SMS, PADSHIFT>(v, t, sm);
            bar();
            gather_slots<E, NT, SMS, PADSHIFT>(v, t, sm);
            pass_compute<E, NT, P::R2, P::R0 * P::R1, INV, PK>(v, t, tw + P::tw_off2);
            if constexpr (P::passes == 3) {
                pass_unpermute<E, P::R2>(v);
            } else {
                bar();
                pass_scatter<E, NT, P::R2, P::R0 * P::R1, SMS, PADSHIFT>(v, t, sm);
                bar();
                gather_slots<E, NT, SMS, PADSHIFT>(v, t, sm);
                pass_compute<E, NT, P::R3, P::R0 * P::R1 * P::R2, INV, PK>(v, t, tw + P::tw_off3);
                pass_unpermute<E, P::R3>(v);
            }
        }
    }
}

// Host: forward twiddle table of a plan (fp64 -> fp32): pass p entry (r-1)*Ns + k = exp(-2 pi i r k/(Ns R)).
template <class P>
inline void build_twiddles(float2* out) {
    const double two_pi = 6.283185307179586476925286766559;
    if (P::R1 > 1) {
        const int Ns = P::R0, R = P::R1;
        for (int r = 1; r < R; ++r)
            for (int k = 0; k < Ns; ++k) {
                double a = -two_pi * (double)r * (double)k / ((double)Ns * R);
                out[P::tw_off1 + (r - 1) * Ns + k] = make_float2((float)cos(a), (float)sin(a));
            }
    }
    if (P::R2 > 1) {
        const int Ns = P::R0 * P::R1, R = P::R2;
        for (int r = 1; r < R; ++r)
            for (int k = 0; k < Ns; ++k) {
                double a = -two_pi * (double)r * (double)k / ((double)Ns * R);
                out[P::tw_off2 + (r - 1) * Ns + k] = make_float2((float)cos(a), (float)sin(a));
            }
    }
    if (P::R3 > 1) {
        const int Ns = P::R0 * P::R1 * P::R2, R = P::R3;
        for (int r = 1; r < R; ++r)
            for (int k = 0; k < Ns; ++k) {
                double a = -two_pi * (double)r * (double)k / ((double)Ns * R);
                out[P::tw_off3 + (r - 1) * Ns + k] = make_float2((float)cos(a), (float)sin(a));
            }
    }
}

}  // namespace fft
}  // namespace nis
