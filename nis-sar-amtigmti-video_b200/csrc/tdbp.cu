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
        transform<P, false, 1, PAD>(v, t, sm, tw);
#pragma unroll
        for (int s = 0; s < E; ++s) v[s] = cmul_pk(v[s], __ldg(Hc + t + NT * s));
        __syncthreads();
        transform<P, true, 1, PAD>(v, t, sm, tw);
#pragma unroll
        for (int s = 0; s < E; ++s) {
            const int j = t + NT * s;
            if (j < B && i0 + j < N) rc[(int64_t)pulse * N + i0 + j] = v[s];
        }
        __syncthreads();
    }
    (void)M;
}

// ------------------------------------------------------------------------------ backprojection
struct TdbpConst {
    double c, inv_c, fc, k_rate, fs, t_start, t_centre, inv_w;
    double x0, dx, y0, dy;     // pixel (i, j) sits at (x0 + i dx, y0 + j dy, 0); the last pixel of a row is pinned to -x0
    double vfx, vfy, vfz;
    int nx, ny, n_pulses, W;
};

__global__ void __launch_bounds__(128) k_tdbp(TdbpConst k, const float2* __restrict__ rc, const double* __restrict__ pos,
                                              const double* __restrict__ vel, const double* __restrict__ t_pulses,
                                              const double* __restrict__ xs, const double* __restrict__ ys,
                                              int p_begin, int p_end, double2* __restrict__ img, int accumulate) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= k.nx * k.ny) return;
    const double gx0 = xs[pix % k.nx], gy0 = ys[pix / k.nx];
    double ar = 0.0, ai = 0.0;
    const float wf = (float)k.W;
    for (int q = p_begin; q < p_end; ++q) {
        const double dt = t_pulses[q] - k.t_centre;
        const double px = pos[3 * q], py = pos[3 * q + 1], pz = pos[3 * q + 2];
        const double vx = vel[3 * q], vy = vel[3 * q + 1], vz = vel[3 * q + 2];
        // pixel moved with the focusing velocity about the CPI centre (:207-209)
        const double gx = gx0 + k.vfx * dt, gy = gy0 + k.vfy * dt, gz = k.vfz * dt;
        const double dx = gx - px, dy = gy - py, dz = gz - pz;
        const double d_tx = sqrt(dx * dx + dy * dy + dz * dz);
        const double inv = 1.0 / d_tx;
        const double v_rad = ((vx - k.vfx) * dx + (vy - k.vfy) * dy + (vz - k.vfz) * dz) * inv;
        const double t_shift = (-k.fc * (2.0 * v_rad * k.inv_c)) * (1.0 / k.k_rate);               // (:214-217)
        const double ta = 2.0 * d_tx * k.inv_c;
        // both ends advanced by the one-way-and-back flight time (:219-222)
        const double ex = (gx + k.vfx * ta) - (px + vx * ta), ey = (gy + k.vfy * ta) - (py + vy * ta),
                     ez = (gz + k.vfz * ta) - (pz + vz * ta);
        const double tau = (d_tx + sqrt(ex * ex + ey * ey + ez * ez)) * k.inv_c;
        const double idx_f = (tau - k.t_start + t_shift) * k.fs;
        const float xn = (float)(2.0 * (idx_f * k.inv_w) - 1.0);                           // grid.float() (:228)
        // grid_sample, bilinear, zeros padding, align_corners=False, as the CPU kernel rounds it
        const float ix = __fmul_rn(__fmaf_rn(__fadd_rn(xn, 1.0f), wf, -1.0f), 0.5f);
        const float f0 = floorf(ix);
        const int i0 = (int)f0;
        const float w1 = __fsub_rn(ix, f0), w0 = __fsub_rn(__fadd_rn(f0, 1.0f), ix);
        const float2* row = rc + (int64_t)q * k.W;
        float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
        if (i0 >= 0 && i0 < k.W) a = __ldg(row + i0);
        if (i0 + 1 >= 0 && i0 + 1 < k.W) b = __ldg(row + i0 + 1);
        const double sr = (double)__fadd_rn(__fmul_rn(a.x, w0), __fmul_rn(b.x, w1));
        const double si = (double)__fadd_rn(__fmul_rn(a.y, w0), __fmul_rn(b.y, w1));
        // exp(j 2 pi FC tau): the carrier phase is ~4e7 turns -- reduced mod 1 in fp64, then a float sincos
        double turns = k.fc * tau;
        turns -= floor(turns);
        float sn, cs;
        sincospif(2.0f * (float)turns, &sn, &cs);
        ar += sr * (double)cs - si * (double)sn;
        ai += sr * (double)sn + si * (double)cs;
    }
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
    static bool attr_done = false;
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
    NIS_CUDA_TRY(cudaSetDevice(ctx->device));
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
    k.c = p.c; k.inv_c = 1.0 / p.c; k.inv_w = 1.0 / (double)p.n_samples; k.fc = p.fc; k.k_rate = p.k_rate; k.fs = p.fs; k.t_start = p.t_start; k.t_centre = t_centre;
    k.vfx = vel_focus_host[0]; k.vfy = vel_focus_host[1]; k.vfz = vel_focus_host[2];
    k.nx = p.nx; k.ny = p.ny; k.n_pulses = n_pulses; k.W = p.n_samples;
    const int n_pix = p.nx * p.ny;
    k_tdbp<<<(n_pix + 127) / 128, 128, 0, (cudaStream_t)stream>>>(k, reinterpret_cast<const float2*>(rc), pos_plat, vel_plat,
                                                                    t_pulses, pl->xs, pl->ys, p_begin, p_end,
                                                                    reinterpret_cast<double2*>(image), accumulate);
    NIS_LAUNCH_CHECK(pl->ctx);
    return NIS_OK;
}
