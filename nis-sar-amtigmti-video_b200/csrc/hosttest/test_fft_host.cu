// Host emulation of the register/shared-memory Stockham transform in fft.cuh: every pass function
// is __host__ __device__, so the exact index arithmetic the kernels use is exercised on the CPU
// (threads run one after another between the points where the device code has __syncthreads()).
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../fft.cuh"

using namespace nis::fft;
namespace nis { void set_error(const char*, ...) {} }

// the transform of one row, threads emulated one after another; result in natural order
template <class P, bool INV, int SMS, int PADSHIFT>
std::vector<float2> transform_host(const std::vector<float2>& x) {
    constexpr int N = P::N, E = P::E, NT = P::NT;
    std::vector<float2> tw(P::tw_len + 1);
    build_twiddles<P>(tw.data());
    std::vector<std::vector<float2>> v(NT, std::vector<float2>(E));
    std::vector<float2> sm((size_t)(N + (PADSHIFT ? (N >> PADSHIFT) : 0) + 8) * SMS);
    for (int t = 0; t < NT; ++t)
        for (int s = 0; s < E; ++s) v[t][s] = x[t + NT * s];
    for (int t = 0; t < NT; ++t) pass_compute<E, NT, P::R0, 1, INV>(v[t].data(), t, tw.data());
    if (P::passes == 1) {
        for (int t = 0; t < NT; ++t) pass_unpermute<E, P::R0>(v[t].data());
    } else {
        for (int t = 0; t < NT; ++t) pass_scatter<E, NT, P::R0, 1, SMS, PADSHIFT>(v[t].data(), t, sm.data());
        for (int t = 0; t < NT; ++t) gather_slots<E, NT, SMS, PADSHIFT>(v[t].data(), t, sm.data());
        for (int t = 0; t < NT; ++t) pass_compute<E, NT, P::R1, P::R0, INV>(v[t].data(), t, tw.data() + P::tw_off1);
        if (P::passes == 2) {
            for (int t = 0; t < NT; ++t) pass_unpermute<E, P::R1>(v[t].data());
        } else {
            for (int t = 0; t < NT; ++t) pass_scatter<E, NT, P::R1, P::R0, SMS, PADSHIFT>(v[t].data(), t, sm.data());
            for (int t = 0; t < NT; ++t) gather_slots<E, NT, SMS, PADSHIFT>(v[t].data(), t, sm.data());
            for (int t = 0; t < NT; ++t)
                pass_compute<E, NT, P::R2, P::R0 * P::R1, INV>(v[t].data(), t, tw.data() + P::tw_off2);
            if (P::passes == 3) {
                for (int t = 0; t < NT; ++t) pass_unpermute<E, P::R2>(v[t].data());
            } else {
                for (int t = 0; t < NT; ++t)
                    pass_scatter<E, NT, P::R2, P::R0 * P::R1, SMS, PADSHIFT>(v[t].data(), t, sm.data());
                for (int t = 0; t < NT; ++t) gather_slots<E, NT, SMS, PADSHIFT>(v[t].data(), t, sm.data());
                for (int t = 0; t < NT; ++t)
                    pass_compute<E, NT, P::R3, P::R0 * P::R1 * P::R2, INV>(v[t].data(), t, tw.data() + P::tw_off3);
                for (int t = 0; t < NT; ++t) pass_unpermute<E, P::R3>(v[t].data());
            }
        }
    }
    std::vector<float2> out(N);
    for (int k = 0; k < N; ++k) out[k] = v[k % NT][k / NT];
    return out;
}

template <class P, bool INV, int SMS, int PADSHIFT>
double run_plan() {
    constexpr int N = P::N;
    std::vector<float2> x(N);
    srand(N * 7 + INV);
    for (auto& e : x) e = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
    const std::vector<float2> v = transform_host<P, INV, SMS, PADSHIFT>(x);
    // reference DFT in double
    double err = 0, nrm = 0;
    const double two_pi = 6.283185307179586476925286766559;
    for (int k = 0; k < N; k += (N > 2048 ? 37 : 1)) {
        std::complex<double> acc = 0;
        for (int n = 0; n < N; ++n) {
            double a = (INV ? 1.0 : -1.0) * two_pi * (double)(((long long)k * n) % N) / N;
            acc += std::complex<double>(x[n].x, x[n].y) * std::complex<double>(cos(a), sin(a));
        }
        float2 got = v[k];
        err += std::norm(acc - std::complex<double>(got.x, got.y));
        nrm += std::norm(acc);
    }
    return sqrt(err / nrm);
}

// The rolled kernels (k_range_rolled, the 16384-point range compressions, ...) run the inverse transform as
// conj(FFT(conj(x))) through the FORWARD code path.  The butterflies of the conjugated-twiddle inverse are mirrored
// exactly (sign flips only); the inter-pass twiddle multiplies are not -- fmaf(a.x, b.y, -(a.y b.x)) rounds the other
// product first than fmaf(a.y, b.x, -(a.x b.y)) -- so the two agree to rounding, not bit for bit.  This pins how closely.
template <class P, int SMS, int PADSHIFT>
double conj_identity_error() {
    constexpr int N = P::N;
    std::vector<float2> x(N), xc(N);
    srand(N * 13 + 5);
    for (int i = 0; i < N; ++i) {
        x[i] = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
        xc[i] = make_float2(x[i].x, -x[i].y);
    }
    const std::vector<float2> inv = transform_host<P, true, SMS, PADSHIFT>(x);
    const std::vector<float2> fwd = transform_host<P, false, SMS, PADSHIFT>(xc);
    double err = 0, nrm = 0;
    for (int i = 0; i < N; ++i) {
        const double dr = (double)inv[i].x - fwd[i].x, di = (double)inv[i].y + fwd[i].y;
        err += dr * dr + di * di;
        nrm += (double)inv[i].x * inv[i].x + (double)inv[i].y * inv[i].y;
    }
    return sqrt(err / nrm);
}

#define CHECK_CONJ(P, SMS, PAD)                                                                  \
    do {                                                                                         \
        const double e = conj_identity_error<P, SMS, PAD>();                                      \
        printf("%-28s sms=%d pad=%d |inverse - conj(forward(conj))| / |inverse| = %.2e\n", #P, SMS, PAD, e); \
        if (!(e < 3e-7)) fails++;                                                                \
    } while (0)

#define CHECK(P, SMS, PAD)                                                                       \
    do {                                                                                         \
        double ef = run_plan<P, false, SMS, PAD>(), ei = run_plan<P, true, SMS, PAD>();          \
        printf("%-28s sms=%d pad=%d fwd %.2e inv %.2e\n", #P, SMS, PAD, ef, ei);                 \
        if (!(ef < 2e-6 && ei < 2e-6)) fails++;                                                  \
    } while (0)

int main() {
    int fails = 0;
    using P16 = Plan<16, 16, 16, 1, 1>;
    using P64 = Plan<64, 8, 8, 8, 1>;
    using P256 = Plan<256, 16, 16, 16, 1>;
    using P512 = Plan<512, 16, 8, 8, 8>;
    using P1024 = Plan<1024, 16, 16, 8, 8>;
    using P2048 = Plan<2048, 16, 16, 16, 8>;
    using P4096 = Plan<4096, 16, 16, 16, 16>;
    using P8192 = Plan<8192, 32, 32, 16, 16>;
    using P128 = Plan<128, 16, 16, 8, 1>;
    using P32 = Plan<32, 32, 32, 1, 1>;
    using P2 = Plan<2, 2, 2, 1, 1>;
    using P4 = Plan<4, 4, 4, 1, 1>;
    using P8 = Plan<8, 8, 8, 1, 1>;
    CHECK(P2, 1, 0); CHECK(P4, 1, 0); CHECK(P8, 1, 0);
    CHECK(P16, 1, 0); CHECK(P32, 1, 0);
    CHECK(P64, 1, 3); CHECK(P64, 32, 0);
    CHECK(P128, 1, 0);
    CHECK(P256, 1, 4); CHECK(P256, 16, 0);
    CHECK(P512, 1, 3);
    CHECK(P1024, 1, 4);
    CHECK(P2048, 1, 4);
    CHECK(P4096, 1, 4);
    CHECK(P8192, 1, 5);
    using Q8192 = Plan<8192, 16, 16, 8, 8, 8>;
    using Q16384 = Plan<16384, 16, 16, 16, 8, 8>;
    using Q512 = Plan<512, 8, 8, 8, 8>;
    using Q4096 = Plan<4096, 8, 8, 8, 8, 8>;
    CHECK(Q8192, 1, 4); CHECK(Q16384, 1, 4); CHECK(Q512, 1, 3); CHECK(Q4096, 1, 3);
    // round 2: two-pass plans of the azimuth tiles, the 32-sample 16384-point plan, and the conjugation identity
    using P512E32 = Plan<512, 32, 32, 16, 1>;
    using P1024E32 = Plan<1024, 32, 32, 32, 1>;
    using P16384E32 = Plan<16384, 32, 32, 32, 16>;
    CHECK(P512E32, 1, 0); CHECK(P512E32, 16, 0); CHECK(P1024E32, 8, 0); CHECK(P16384E32, 1, 5);
    CHECK_CONJ(P8192, 1, 5); CHECK_CONJ(P16384E32, 1, 5); CHECK_CONJ(Q8192, 1, 4); CHECK_CONJ(P512E32, 16, 0);
    CHECK_CONJ(P4096, 1, 4);
    printf(fails ? "FAILED %d\n" : "all ok\n", fails);
    return fails != 0;
}
