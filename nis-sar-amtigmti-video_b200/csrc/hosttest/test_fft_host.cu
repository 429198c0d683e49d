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

template <class P, bool INV, int SMS, int PADSHIFT>
double run_plan() {
    constexpr int N = P::N, E = P::E, NT = P::NT;
    std::vector<float2> tw(P::tw_len + 1);
    build_twiddles<P>(tw.data());
    std::vector<float2> x(N);
    srand(N * 7 + INV);
    for (auto& e : x) e = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
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
    // reference DFT in double
    double err = 0, nrm = 0;
    const double two_pi = 6.283185307179586476925286766559;
    for (int k = 0; k < N; k += (N > 2048 ? 37 : 1)) {
        std::complex<double> acc = 0;
        for (int n = 0; n < N; ++n) {
            double a = (INV ? 1.0 : -1.0) * two_pi * (double)(((long long)k * n) % N) / N;
            acc += std::complex<double>(x[n].x, x[n].y) * std::complex<double>(cos(a), sin(a));
        }
        float2 got = v[k % NT][k / NT];
        err += std::norm(acc - std::complex<double>(got.x, got.y));
        nrm += std::norm(acc);
    }
    return sqrt(err / nrm);
}

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
    printf(fails ? "FAILED %d\n" : "all ok\n", fails);
    return fails != 0;
}
