// tdbp.cu -- time-domain backprojection of a spotlight CPI (replaces tdbp_gpu, sar_batch_sim.py:171-238; SURVEY.md section
// 8f row N4: the reference's VideoSAR path -- 50 frames of a sliding 2500-pulse CPI, each backprojected on a 512 x 512 grid).
//
//   k_tdbp_range   circular matched filtering of every pulse (:179-185).  The reference correlates through FFTs of length
//                  num_samples (22004 = 4 x 5501, not a friendly length); the same numbers come from overlap-save blocks on
//                  the register-resident power-of-two FFTs: rc[i] = sum_m raw[(i + m) mod N] conj(ref_s[m]), m < L taps,
//                  block b produces M - L + 1 outputs from M (cyclically indexed) inputs.
//   k_tdbp         one thread per pixel, loop over pulses (the pulse's position / velocity / time are warp-uniform loads):
//                  fp64 geometry (:207-223) -> sample index -> float32 coordinate EXACTLY as torch's grid_sample sees it
//                  (idx_norm cast to fp32, ix = fma(x + 1, W, -1) / 2 with one rounding, pinned against torch in
//                  oracle/make_golden.py) -> two-tap linear interpolation of the fp32 range-compressed pulse (zero outside)
//                  -> x exp(j 2 pi FC tau) with the carrier phase reduced mod 1 in fp64 -> complex128 accumulation.
//                  The float32 coordinate is part of the reference's result (it moves the sample point by up to 1e-3
//                  samples of a nearly critically sampled signal), so it is reproduced rather than "improved".
#include <math.h>

#include <complex>
#include <vector>

#include "common.cuh"
#include "fft.cuh"

#ifndef TDBP_UNROLL
#define TDBP_UNROLL 4
#endif

using namespace nis;
using namespace nis::fft;

namespace {

// ------------------------------------------------------------------------------ range compression (overlap-save)
template <class P, int PAD>
__global__ void __launch_bounds__(P::NT) k_tdbp_range(const float2* __restrict__ raw, int64_t pitch, float2* __restrict__ rc,
                                                      int n_pulses, int N, int n_blocks, int B,
                                                      const float2* __restrict__ Hc, const float2* __restrict__ tw) {
    extern __shared__ float2 sm[];
    constexpr int E = P::E, NT = P::NT, M = P::N;
    const int t = threadIdx.x;
    const int total = n_pulses * n_blocks;
    for (int job = blockIdx.x; job < total; job += gridDim.x) {
        const int pulse = job / n_blocks, i0 = (job % n_blocks) * B;
        const float2* p = raw + (int64_t)pulse * pitch;
        float2 v[E];
#pragma unroll
        for (int s = 0; s < E; ++s) {
            int idx = i0 + t + NT * s;          // cyclic: the circular correlation wraps around the window
            if (idx >= N) idx -= N;
            if (idx >= N) idx %= N;
            v[s] = p[idx];
        }
        // forward and inverse transform share ONE copy of the code (inverse = conj FFT conj, equal to the
        // conjugated-twiddle form up to rounding: 1e-7, csrc/hosttest): two inlined bodies of the 32-element plan spilled 340 bytes per thread at 128 registers
#pragma unroll 1
        for (int step = 0; step < 2; ++step) {
            transform<P, false, 1, PAD>(v, t, sm, tw);
            if (step == 0) {   // v <- conj(v Hc)
#pragma unroll
                for (int s = 0; s < E; ++s) {
                    const float2 h = __ldg(Hc + t + NT * s), x = v[s];
                    v[s] = make_float2(fmaf(x.x, h.x, -x.y * h.y), fmaf(-x.x, h.y, -x.y * h.x));
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int s = 0; s < E; ++s) {
            const int j = t + NT * s;
            if (j < B && i0 + j < N) rc[(int64_t)pulse * N + i0 + j] = make_float2(v[s].x, -v[s].y);
        }
    }
    (void)M;
}

// ------------------------------------------------------------------------------ backprojection
struct TdbpConst {
    double c, inv_c, fc, k_rate, fs, t_start, t_centre, inv_w;
    double two_inv_c, xn_tau, xn_vrad, xn_0, fc_2p32;   // folded constants of the sample coordinate and the carrier phase
    double vfx, vfy, vfz;
    int nx, ny, n_pulses, W;
};

// 1/sqrt(a) and sqrt(a) to within ~1 ulp without the DSQRT / DRCP sequences (~55 instructions) and -- after the first
// pulse -- without touching the conversion / MUFU pipe at all: the seed y0 is the reciprocal distance of the previous
// evaluation (pulse to pulse and transmit to receive the distance changes by ~1e-6 relative), refined by one third-order
// step (error h^3: 1e-6 -> 1e-18) and one Heron correction of the root.  A float rsqrt seed (2^-22 -> 2^-66) serves the
// first pulse and any caller whose consecutive pulses are not close (|h| test).
__device__ __forceinline__ void rsqrt_sqrt(double a, double y0, double& rinv, double& root) {
    double y = y0;
    double h = fma(-a * y, y, 1.0);
    if (fabs(h) > 1e-4) {
        y = (double)rsqrtf((float)a);
        h = fma(-a * y, y, 1.0);
    }
    y = fma(y * h, fma(0.375, h, 0.5), y);
    const double r = a * y;
    root = fma(0.5 * y, fma(-r, r, a), r);
    rinv = y;
}

// per-pulse quantities that every pixel shares (6 doubles): the pixel-independent part of (pixel - platform) with the pixel
// moved by the focusing velocity about the CPI centre (:207-210), and the platform velocity relative to the focused frame
__global__ void k_tdbp_pulse_table(TdbpConst k, const double* __restrict__ pos, const double* __restrict__ vel,
                                   const double* __restrict__ t_pulses, int n_pulses, double* __restrict__ tab) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_pulses) return;
    const double dt = t_pulses[q] - k.t_centre;
    tab[6 * q + 0] = k.vfx * dt - pos[3 * q];
    tab[6 * q + 1] = k.vfy * dt - pos[3 * q + 1];
    tab[6 * q + 2] = k.vfz * dt - pos[3 * q + 2];
    tab[6 * q + 3] = vel[3 * q] - k.vfx;
    tab[6 * q + 4] = vel[3 * q + 1] - k.vfy;
    tab[6 * q + 5] = vel[3 * q + 2] - k.vfz;
}

__global__ void __launch_bounds__(64) k_tdbp(TdbpConst k, const float2* __restrict__ rc, const double* __restrict__ tab,
                                             const double* __restrict__ xs, const double* __restrict__ ys, int p_begin,
                                             int p_end, double2* __restrict__ img, int accumulate) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= k.nx * k.ny) return;
    const double gx0 = xs[pix % k.nx], gy0 = ys[pix / k.nx];
    double ar = 0.0, ai = 0.0;
    float pr = 0.f, pi = 0.f;        // fp32 partial sums, flushed into the fp64 accumulators every 8 pulses
    double inv = 0.0;                // reciprocal distance carried from pulse to pulse (seed of the next ones)
    const float wf = (float)k.W;
    // one pulse: returns the rotated sample; `seed` ~ 1 / distance
    auto pulse = [&](int q, double seed, double& inv_out, float& re, float& im) {
        const double* tq = tab + 6 * q;                       // warp-uniform loads
        const double dx = gx0 + tq[0], dy = gy0 + tq[1], dz = tq[2];
        const double rvx = tq[3], rvy = tq[4], rvz = tq[5];
        double d_tx, iv;
        rsqrt_sqrt(fma(dx, dx, fma(dy, dy, dz * dz)), seed, iv, d_tx);
        const double v_rad = fma(rvx, dx, fma(rvy, dy, rvz * dz)) * iv;           // (v_plat - v_focus) . r_unit (:211-213)
        const double ta = d_tx * k.two_inv_c;                                      // 2 d_tx / c
        // both ends advanced by the flight time (:219-222): (g + v_f ta) - (p + v ta) = d - (v - v_f) ta
        const double ex = fma(-rvx, ta, dx), ey = fma(-rvy, ta, dy), ez = fma(-rvz, ta, dz);
        double d_rx, inv_rx;
        rsqrt_sqrt(fma(ex, ex, fma(ey, ey, ez * ez)), iv, inv_rx, d_rx);
        const double tau = (d_tx + d_rx) * k.inv_c;
        // idx_norm = 2 ((tau - t_start + t_shift) FS / W) - 1 with t_shift = -FC (2 v_rad / C) / K_RATE (:214-226)
        const float xn = (float)fma(tau, k.xn_tau, fma(v_rad, k.xn_vrad, k.xn_0));   // grid.float() (:228)
        // grid_sample, bilinear, zeros padding, align_corners=False, as the CPU kernel rounds it
        const float ix = __fmul_rn(__fmaf_rn(__fadd_rn(xn, 1.0f), wf, -1.0f), 0.5f);
        const float f0 = floorf(ix);
        const int i0 = (int)f0;
        const float w1 = __fsub_rn(ix, f0), w0 = __fsub_rn(__fadd_rn(f0, 1.0f), ix);
        const float2* row = rc + (int64_t)q * k.W;
        float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
        if (i0 >= 0 && i0 < k.W) a = __ldg(row + i0);
        if (i0 + 1 >= 0 && i0 + 1 < k.W) b = __ldg(row + i0 + 1);
        const float sr = __fadd_rn(__fmul_rn(a.x, w0), __fmul_rn(b.x, w1));
        const float si = __fadd_rn(__fmul_rn(a.y, w0), __fmul_rn(b.y, w1));
        // exp(j 2 pi FC tau): ~4e7 turns.  tau * (FC 2^32) < 2^63 keeps the fraction in the low 32 bits of the integer
        // (7e-9 turns of resolution); two MUFU ops on the fixed-point fraction.  The rotated sample is formed in fp32
        // (1e-7 relative, incoherent over the pulses), summed in fp32 over 8 pulses and then accumulated in fp64.
        const float2 e = cis_u32((uint32_t)(unsigned long long)(long long)(tau * k.fc_2p32));
        re = fmaf(sr, e.x, -si * e.y);
        im = fmaf(sr, e.y, si * e.x);
        inv_out = iv;
    };
    int q = p_begin;
    // U pulses per iteration: U independent fp64 dependency chains in flight per thread (measured: 4.45 ms at U = 1,
    // 3.67 ms at U = 2, 3.44 ms at U = 4 for the full-size frame)
    constexpr int U = TDBP_UNROLL;
    for (int it = 0; q + U <= p_end; q += U, ++it) {
        float re[U], im[U];
        double iv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) pulse(q + u, inv, iv[u], re[u], im[u]);
        inv = iv[U - 1];
#pragma unroll
        for (int u = 0; u < U; ++u) { pr += re[u]; pi += im[u]; }
        if ((it & (8 / U - 1)) == 8 / U - 1) {
            ar += (double)pr;
            ai += (double)pi;
            pr = pi = 0.f;
        }
    }
    for (; q < p_end; ++q) {
        float r0, i0;
        double inv0;
        pulse(q, inv, inv0, r0, i0);
        inv = inv0;
        pr += r0;
        pi += i0;
    }
    ar += (double)pr;
    ai += (double)pi;
    if (accumulate) {
        img[pix].x += ar;
        img[pix].y += ai;
    } else {
        img[pix] = make_double2(ar, ai);
    }
}

using P256 = Plan<256, 16, 16, 16, 1>;
using P1024 = Plan<1024, 16, 16, 8, 8>;
using P4096 = Plan<4096, 16, 16, 16, 16>;
using P16384 = Plan<16384, 32, 32, 32, 16>;

void host_fft_pow2(std::vector<std::complex<double>>& a) {   // in-place radix-2, forward
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    const double two_pi = 6.283185307179586476925286766559;
    for (size_t len = 2; len <= n; len <<= 1)
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const double ang = -two_pi * (double)k / (double)len;
                const std::complex<double> w(cos(ang), sin(ang));
                const std::complex<double> u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
}

}  // namespace

struct nis_tdbp_plan {
    nis_ctx* ctx = nullptr;
    nis_tdbp_params prm{};
    int L = 0, M = 0, B = 0, n_blocks = 0;
    float2 *Hc = nullptr, *tw = nullptr;
    double *xs = nullptr, *ys = nullptr;
    int (*range_fn)(nis_tdbp_plan*, const float2*, int64_t, float2*, int, cudaStream_t) = nullptr;
};

namespace {

template <class P, int PAD>
int launch_tdbp_range(nis_tdbp_plan* pl, const float2* raw, int64_t pitch, float2* rc, int n_pulses, cudaStream_t st) {
    constexpr int SMROW = P::N + (PAD ? (P::N >> PAD) : 0);
    const size_t smem = (size_t)SMROW * sizeof(float2);
    static bool attr_done_dev[64] = {};
    bool& attr_done = attr_done_dev[nis::current_device() & 63];
    if (!attr_done) {
        NIS_CUDA_TRY(cudaFuncSetAttribute(k_tdbp_range<P, PAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    int per_sm = 1;
    NIS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tdbp_range<P, PAD>, P::NT, smem));
    const int jobs = n_pulses * pl->n_blocks;
    int grid = pl->ctx->num_sms * (per_sm < 1 ? 1 : per_sm);
    if (grid > jobs) grid = jobs;
    k_tdbp_range<P, PAD><<<grid, P::NT, smem, st>>>(raw, pitch, rc, n_pulses, pl->prm.n_samples, pl->n_blocks, pl->B, pl->Hc,
                                                     pl->tw);
    NIS_LAUNCH_CHECK(pl->ctx);
    return NIS_OK;
}

template <class P>
int upload_tw(float2** dev) {
    std::vector<float2> h(P::tw_len + 1);
    build_twiddles<P>(h.data());
    NIS_CUDA_TRY(cudaMalloc(dev, h.size() * sizeof(float2)));
    NIS_CUDA_TRY(cudaMemcpy(*dev, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
    return NIS_OK;
}

// numpy.linspace(-s/2, s/2, n)
std::vector<double> linspace(double start, double stop, int n) {
    std::vector<double> v(n);
    const double step = n > 1 ? (stop - start) / (double)(n - 1) : 0.0;
    for (int i = 0; i < n; ++i) v[i] = (double)i * step + start;
    if (n > 1) v[n - 1] = stop;
    return v;
}

}  // namespace

extern "C" int nis_tdbp_plan_destroy(nis_tdbp_plan* pl) {
    if (!pl) return NIS_OK;
    cudaFree(pl->Hc);
    cudaFree(pl->tw);
    cudaFree(pl->xs);
    cudaFree(pl->ys);
    delete pl;
    return NIS_OK;
}

extern "C" int nis_tdbp_plan_create(nis_ctx* ctx, const nis_tdbp_params* prm, nis_tdbp_plan** out) {
    NIS_REQUIRE(ctx && prm && out, "nis_tdbp_plan_create: null argument");
    NIS_REQUIRE(prm->c > 0 && prm->fs > 0 && prm->t_p > 0 && prm->k_rate != 0 && prm->n_samples >= 2 && prm->nx >= 1 &&
                    prm->ny >= 1,
                "nis_tdbp_plan_create: non-physical parameters");
    const int L = (int)(prm->t_p * prm->fs);   // int(T_P * FS) taps (:177)
    NIS_REQUIRE(L >= 1 && L <= 12288, "nis_tdbp_plan_create: %d reference-chirp taps (supported: 1..12288)", L);
    DeviceGuard device_guard(ctx->device);   // the caller's current device is restored on return
    nis_tdbp_plan* pl = new nis_tdbp_plan();
    pl->ctx = ctx;
    pl->prm = *prm;
    pl->L = L;
    const int N = prm->n_samples;
    // block FFT length: the whole window in one block when it fits, else 4 L rounded up (at most 16384)
    const int ms[] = {256, 1024, 4096, 16384};
    int M = 0;
    for (int m : ms)
        if (m >= N + L - 1) { M = m; break; }
    if (!M)
        for (int m : ms)
            if (m >= 4 * L) { M = m; break; }
    if (!M) M = 16384;
    pl->M = M;
    pl->B = M - L + 1;
    pl->n_blocks = (N + pl->B - 1) / pl->B;
    int rc = NIS_OK;
#define FAIL_IF(x) do { rc = (x); if (rc != NIS_OK) { nis_tdbp_plan_destroy(pl); return rc; } } while (0)
#define CUDA_FAIL_IF(x) FAIL_IF((x) == cudaSuccess ? NIS_OK : (set_error("%s failed", #x), NIS_ERR_CUDA))
    {
        // reference chirp exp(j pi K t^2) on linspace(-T_P/2, T_P/2, L), fftshifted (:177-179); block spectrum
        // conj(FFT_M(ref_s)) / M turns the block product into the correlation with ref_s
        const std::vector<double> tr = linspace(-prm->t_p / 2, prm->t_p / 2, L);
        std::vector<std::complex<double>> h(M, 0.0);
        const int sh = L / 2;   // fftshift moves element i to (i + L // 2) mod L
        for (int i = 0; i < L; ++i) {
            const double ph = M_PI * prm->k_rate * (tr[i] * tr[i]);
            h[(i + sh) % L] = std::complex<double>(cos(ph), sin(ph));
        }
        host_fft_pow2(h);
        std::vector<float2> hc(M);
        for (int i = 0; i < M; ++i) hc[i] = make_float2((float)(h[i].real() / M), (float)(-h[i].imag() / M));
        CUDA_FAIL_IF(cudaMalloc(&pl->Hc, M * sizeof(float2)));
        CUDA_FAIL_IF(cudaMemcpy(pl->Hc, hc.data(), M * sizeof(float2), cudaMemcpyHostToDevice));
    }
    switch (M) {
        case 256: pl->range_fn = launch_tdbp_range<P256, 4>; FAIL_IF(upload_tw<P256>(&pl->tw)); break;
        case 1024: pl->range_fn = launch_tdbp_range<P1024, 4>; FAIL_IF(upload_tw<P1024>(&pl->tw)); break;
        case 4096: pl->range_fn = launch_tdbp_range<P4096, 4>; FAIL_IF(upload_tw<P4096>(&pl->tw)); break;
        default: pl->range_fn = launch_tdbp_range<P16384, 5>; FAIL_IF(upload_tw<P16384>(&pl->tw)); break;
    }
    {
        const std::vector<double> xs = linspace(-prm->scene_size / 2, prm->scene_size / 2, prm->nx);
        const std::vector<double> ys = linspace(-prm->scene_size / 2, prm->scene_size / 2, prm->ny);
        CUDA_FAIL_IF(cudaMalloc(&pl->xs, xs.size() * sizeof(double)));
        CUDA_FAIL_IF(cudaMalloc(&pl->ys, ys.size() * sizeof(double)));
        CUDA_FAIL_IF(cudaMemcpy(pl->xs, xs.data(), xs.size() * sizeof(double), cudaMemcpyHostToDevice));
        CUDA_FAIL_IF(cudaMemcpy(pl->ys, ys.data(), ys.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
#undef CUDA_FAIL_IF
#undef FAIL_IF
    *out = pl;
    return NIS_OK;
}

extern "C" int nis_tdbp_range_compress(nis_tdbp_plan* pl, const nis_c32* raw, int64_t pitch, int32_t n_pulses, nis_c32* rc,
                                       nis_stream stream) {
    NIS_REQUIRE(pl && raw && rc, "nis_tdbp_range_compress: null argument");
    NIS_REQUIRE(n_pulses >= 0 && pitch >= pl->prm.n_samples, "nis_tdbp_range_compress: bad sizes");
    if (n_pulses == 0) return NIS_OK;
    return pl->range_fn(pl, reinterpret_cast<const float2*>(raw), pitch, reinterpret_cast<float2*>(rc), n_pulses,
                        (cudaStream_t)stream);
}

extern "C" int nis_tdbp_backproject(nis_tdbp_plan* pl, const nis_c32* rc, const double* pos_plat, const double* vel_plat,
                                    const double* t_pulses, int32_t n_pulses, int32_t p_begin, int32_t p_end,
                                    double t_centre, const double* vel_focus_host, double* image, int32_t accumulate,
                                    nis_stream stream) {
    NIS_REQUIRE(pl && rc && pos_plat && vel_plat && t_pulses && vel_focus_host && image, "nis_tdbp_backproject: null argument");
    NIS_REQUIRE(0 <= p_begin && p_begin <= p_end && p_end <= n_pulses, "nis_tdbp_backproject: pulse range [%d, %d) of %d",
                p_begin, p_end, n_pulses);
    const nis_tdbp_params& p = pl->prm;
    TdbpConst k{};
    {
        const double shift_scale = (-p.fc * (2.0 / p.c)) / p.k_rate;      // t_shift per unit radial velocity
        const double a = 2.0 * p.fs / (double)p.n_samples;               // d idx_norm / d time
        k.two_inv_c = 2.0 / p.c;
        k.xn_tau = a;
        k.xn_vrad = a * shift_scale;
        k.xn_0 = -p.t_start * a - 1.0;
        k.fc_2p32 = p.fc * 4294967296.0;
    }
    k.c = p.c; k.inv_c = 1.0 / p.c; k.inv_w = 1.0 / (double)p.n_samples; k.fc = p.fc; k.k_rate = p.k_rate; k.fs = p.fs; k.t_start = p.t_start; k.t_centre = t_centre;
    k.vfx = vel_focus_host[0]; k.vfy = vel_focus_host[1]; k.vfz = vel_focus_host[2];
    k.nx = p.nx; k.ny = p.ny; k.n_pulses = n_pulses; k.W = p.n_samples;
    const int n_pix = p.nx * p.ny;
    cudaStream_t st = (cudaStream_t)stream;
    void* tab_raw = nullptr;
    int rc2 = pl->ctx->stream_scratch(st, (size_t)n_pulses * 6 * sizeof(double), &tab_raw);   // per stream, never freed under a graph
    if (rc2 != NIS_OK) return rc2;
    double* tab = reinterpret_cast<double*>(tab_raw);
    k_tdbp_pulse_table<<<(n_pulses + 127) / 128, 128, 0, st>>>(k, pos_plat, vel_plat, t_pulses, n_pulses, tab);
    NIS_LAUNCH_CHECK(pl->ctx);
    k_tdbp<<<(n_pix + 63) / 64, 64, 0, st>>>(k, reinterpret_cast<const float2*>(rc), tab, pl->xs, pl->ys, p_begin, p_end,
                                             reinterpret_cast<double2*>(image), accumulate);
    NIS_LAUNCH_CHECK(pl->ctx);
    return NIS_OK;
}
