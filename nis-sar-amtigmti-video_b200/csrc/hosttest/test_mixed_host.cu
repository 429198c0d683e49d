// Host emulation of the compile-time mixed-radix engine (mixed_ct.cuh): the small DFTs (odd radices through conjugate-symmetric
// pairs, 10 = 2 x 5 and 12 = 4 x 3 by the prime-factor map), the per-pass twiddle tables (full and powers-of-w^k forms), and the
// pass index arithmetic -- forward Stockham passes 0..3 and the TRANSPOSED inverse passes 3..0 -- are __host__ __device__, so
// the sequence k_row_mixed_ct runs is replayed here on the CPU (threads one after another between the barriers) against a
// double-precision DFT.
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../mixed_ct.cuh"

using namespace nis::mixedct;
namespace nis { void set_error(const char*, ...) {} }

template <int R>
double check_small() {
    double worst = 0;
    for (int inv = 0; inv < 2; ++inv) {
        float2 v[R];
        std::complex<double> x[R];
        for (int i = 0; i < R; ++i) {
            v[i] = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
            x[i] = {v[i].x, v[i].y};
        }
        if (inv) dft_small<R, true>(v); else dft_small<R, false>(v);
        double err = 0, nrm = 0;
        for (int q = 0; q < R; ++q) {
            std::complex<double> acc = 0;
            for (int n = 0; n < R; ++n) {
                const double a = (inv ? 2.0 : -2.0) * kPi * (double)((q * n) % R) / R;
                acc += x[n] * std::complex<double>(cos(a), sin(a));
            }
            err += std::norm(acc - std::complex<double>(v[q].x, v[q].y));
            nrm += std::norm(acc);
        }
        worst = fmax(worst, sqrt(err / nrm));
    }
    return worst;
}

template <class MP>
int check_plan(const char* name) {
    constexpr int N = MP::N, NT = MP::NT;
    using P0 = Pass<MP, MP::R0, 1, false>;
    using P1 = Pass<MP, MP::R1, MP::NS1, false>;
    using P2 = Pass<MP, MP::R2, MP::NS2, false>;
    using P3 = Pass<MP, MP::R3, MP::NS3, MP::POW3>;
    std::vector<float2> tw(MP::TW_LEN), x(N), a(N), b(N), X(N), y(N);
    build_tables<MP>(tw.data());
    srand(N);
    for (auto& e : x) e = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
    // ---- forward: passes 0..3
    for (int t = 0; t < NT; ++t)
        for (int it = 0; it < P0::IT; ++it) {
            const int j = t + NT * it;
            if (!P0::active(j)) continue;
            float2 v[MP::R0];
            for (int r = 0; r < MP::R0; ++r) v[r] = x[j + r * P0::NB];
            dft_small<MP::R0, false>(v);
            P0::spread_t(v, a.data(), j, 0);
        }
    auto fwd_mid = [&](auto pass, const std::vector<float2>& src, std::vector<float2>& dst, const float2* tab) {
        using PP = decltype(pass);
        for (int t = 0; t < NT; ++t)
            for (int it = 0; it < PP::IT; ++it) {
                const int j = t + NT * it;
                if (!PP::active(j)) continue;
                const int k = j % PP::NS_;
                float2 v[PP::R_];
                PP::gather(v, src.data(), j);
                PP::template twiddle<false>(v, tab, k);
                dft_small<PP::R_, false>(v);
                PP::spread_t(v, dst.data(), j, k);
            }
    };
    fwd_mid(P1{}, a, b, tw.data() + MP::TW1);
    fwd_mid(P2{}, b, a, tw.data() + MP::TW2);
    for (int t = 0; t < NT; ++t)
        for (int it = 0; it < P3::IT; ++it) {
            const int j = t + NT * it;
            if (!P3::active(j)) continue;
            float2 w[MP::R3];
            P3::gather(w, a.data(), j);
            P3::template twiddle<false>(w, tw.data() + MP::TW3, j);
            dft_small<MP::R3, false>(w);
            for (int q = 0; q < MP::R3; ++q) X[j + q * MP::NS3] = w[q];
        }
    // ---- inverse: transposed passes 3..0
    for (int t = 0; t < NT; ++t)
        for (int it = 0; it < P3::IT; ++it) {
            const int j = t + NT * it;
            if (!P3::active(j)) continue;
            float2 v[MP::R3];
            for (int q = 0; q < MP::R3; ++q) v[q] = X[j + q * MP::NS3];
            dft_small<MP::R3, true>(v);
            P3::template twiddle<true>(v, tw.data() + MP::TW3, j);
            P3::spread(v, a.data(), j);
        }
    auto inv_mid = [&](auto pass, const std::vector<float2>& src, std::vector<float2>& dst, const float2* tab) {
        using PP = decltype(pass);
        for (int t = 0; t < NT; ++t)
            for (int it = 0; it < PP::IT; ++it) {
                const int j = t + NT * it;
                if (!PP::active(j)) continue;
                const int k = j % PP::NS_;
                float2 v[PP::R_];
                PP::gather_t(v, src.data(), j, k);
                dft_small<PP::R_, true>(v);
                PP::template twiddle<true>(v, tab, k);
                PP::spread(v, dst.data(), j);
            }
    };
    inv_mid(P2{}, a, b, tw.data() + MP::TW2);
    inv_mid(P1{}, b, a, tw.data() + MP::TW1);
    for (int t = 0; t < NT; ++t)
        for (int it = 0; it < P0::IT; ++it) {
            const int j = t + NT * it;
            if (!P0::active(j)) continue;
            float2 w[MP::R0];
            P0::gather_t(w, a.data(), j, 0);
            dft_small<MP::R0, true>(w);
            for (int r = 0; r < MP::R0; ++r) y[j + r * P0::NB] = w[r];
        }
    // ---- checks: sampled bins against a double DFT; round trip y == N x
    double err = 0, nrm = 0;
    for (int k = 0; k < N; k += 53) {
        std::complex<double> acc = 0;
        for (int n = 0; n < N; ++n) {
            const double ang = -2.0 * kPi * (double)(((long long)k * n) % N) / N;
            acc += std::complex<double>(x[n].x, x[n].y) * std::complex<double>(cos(ang), sin(ang));
        }
        err += std::norm(acc - std::complex<double>(X[k].x, X[k].y));
        nrm += std::norm(acc);
    }
    const double e_fwd = sqrt(err / nrm);
    err = nrm = 0;
    for (int n = 0; n < N; ++n) {
        const std::complex<double> want((double)N * x[n].x, (double)N * x[n].y);
        err += std::norm(want - std::complex<double>(y[n].x, y[n].y));
        nrm += std::norm(want);
    }
    const double e_rt = sqrt(err / nrm);
    printf("%-10s N=%5d NT=%4d radices %d.%d.%d.%d  fwd %.2e  inverse(forward) %.2e\n", name, N, NT, MP::R0, MP::R1, MP::R2,
           MP::R3, e_fwd, e_rt);
    return (e_fwd < 2e-6 && e_rt < 2e-6) ? 0 : 1;
}

#define SMALL(R)                                             \
    do {                                                     \
        const double e = check_small<R>();                   \
        printf("dft_small<%2d>  %.2e\n", R, e);              \
        if (!(e < 1e-6)) fails++;                            \
    } while (0)

int main() {
    int fails = 0;
    srand(1);
    SMALL(2); SMALL(3); SMALL(4); SMALL(5); SMALL(6); SMALL(7); SMALL(8); SMALL(9); SMALL(10); SMALL(11); SMALL(12);
    SMALL(13); SMALL(15); SMALL(16); SMALL(20);
    fails += check_plan<MP13200>("MP13200");
    fails += check_plan<MP7200>("MP7200");
    using MPsmallA = MPlan<1320, 64, 11, 10, 3, 4, true>;     // partial last iterations, powers-of-w^k pass 3
    using MPsmallB = MPlan<2520, 96, 7, 9, 5, 8, false>;
    fails += check_plan<MPsmallA>("small A");
    fails += check_plan<MPsmallB>("small B");
    printf(fails ? "FAILED %d\n" : "all ok\n", fails);
    return fails != 0;
}
