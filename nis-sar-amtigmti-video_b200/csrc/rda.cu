// rda.cu -- Range-Doppler focusing (replaces sar_focus_rda: sar_satellite_sim.py:356-448, sar_vehicle_sim.py:182-274,
// sar_satellite_moving_sim.py:208-285).
//
// Everything stays in the pulse-major layout raw[P][S] of the echo engines (the reference's `phist` is the .T view of it,
// sar_satellite_sim.py:453-454), and the image |.|.T the reference returns is again [P][S], so no corner turn is needed:
//   k_rda_range     per pulse: zero-padded FFT_M -> x FFT_M(matched filter)/M -> IFFT_M -> the N samples of the linear
//                   convolution that scipy's convolve(mode='same') keeps (:388-392); writes the compressed pulse
//                   (optional export) and, times the azimuth Hamming weight of the pulse (:396-397), the workspace
//   azimuth DFT     four-step engine of csa.cu (power-of-two P) or transpose + row-DFT engine (any other P)
//   k_rda_rcmc      per Doppler bin: range-cell-migration correction = linear interpolation from the axis
//                   r (1 - fd^2 lam^2 / 8 Vr^2) back onto r, zero outside it (:408-427), then x exp(-j pi fd^2 / Ka(r))
//                   (:431-435); optional exports of the three Range-Doppler maps the viewers read
//   inverse DFT     -> |.| / P  (:438-439)
// The fftshift / ifftshift pairs around both transforms (:398, :438) cancel for the image (any P, even or odd): the
// image is IFFT_k( G[r, fd(k)] . FFT_n(w . rc) ) in natural order.  Only the exported Range-Doppler maps carry them:
// row j = (k + floor(P/2)) mod P and the factor exp(-2 pi i floor(P/2) k / P) ((-1)^k for even P).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <complex>
#include <vector>

#include "csa_internal.cuh"
#include "fft.cuh"
#include "tma.cuh"

using namespace nis;
using namespace nis::fft;
using namespace nis::csa;

namespace {

struct RdaRow {          // one per stored Doppler row
    double beta;         // alpha / (1 - alpha), alpha = fd^2 lam^2 / (8 Vr^2): source position p = i + (i + q0) beta
    uint64_t ph_a, ph_b; // azimuth-compression phase in fixed-point turns: ph_b + ph_a * i
    float2 factor;       // exported maps: stored value times this
    int32_t export_row;  // exported maps: row in fftshift order
    int32_t pad;
};
static_assert(sizeof(RdaRow) == 40, "RdaRow layout");

// ------------------------------------------------------------------------------ range compression
// (8192 points on the 32-samples-per-thread three-pass plan: 256 threads at <= 128 registers, two CTAs per SM -- the
// arrangement of k_range_rolled in csa.cu)
template <class P, int PAD>
__global__ void __launch_bounds__(P::NT, (P::E == 32 && P::N == 8192) ? 2 : ((P::E == 32 && P::N == 4096) ? 4 : 1)) k_rda_range(const float2* __restrict__ in, int64_t in_pitch,
                                                     float2* __restrict__ work, int64_t work_pitch,
                                                     float2* __restrict__ rc_out, int n_rows, int N, int s0,
                                                     const float2* __restrict__ Hf, const float* __restrict__ win,
                                                     const float2* __restrict__ tw) {
    extern __shared__ float2 sm[];
    constexpr int E = P::E, NT = P::NT;
    const int t = threadIdx.x;
    for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
        const float2* p = in + (int64_t)row * in_pitch;
        float2 v[E];
#pragma unroll
        for (int s = 0; s < E; ++s) {
            const int idx = t + NT * s;
            v[s] = idx < N ? p[idx] : make_float2(0.f, 0.f);
        }
        // forward and inverse transform share ONE copy of the code: the inverse runs as conj(FFT(conj(.))), equal to
        // the conjugated-twiddle form up to rounding (1e-7, csrc/hosttest), the conjugations folded into the filter multiply and the store (two inlined bodies of
        // the 32-element 16384-point plan spilled 680 bytes per thread at its 128-register cap)
#pragma unroll 1
        for (int step = 0; step < 2; ++step) {
            transform<P, false, 1, PAD>(v, t, sm, tw);
            if (step == 0) {   // v <- conj(v Hf)
#pragma unroll
                for (int s = 0; s < E; ++s) {
                    const float2 h = __ldg(Hf + t + NT * s), x = v[s];
                    v[s] = make_float2(fmaf(x.x, h.x, -x.y * h.y), fmaf(-x.x, h.y, -x.y * h.x));
                }
            }
            __syncthreads();
        }
        const float w = win[row];
#pragma unroll
        for (int s = 0; s < E; ++s) {
            const int o = t + NT * s - s0;
            if (o >= 0 && o < N) {
                const float2 y = make_float2(v[s].x, -v[s].y);
                if (rc_out != nullptr) rc_out[(int64_t)row * N + o] = y;
                work[(int64_t)row * work_pitch + o] = make_float2(y.x * w, y.y * w);
            }
        }
    }
}

// Overlap-save variant for filters too long for one block (the satellite scripts compress 13200-sample pulses with a
// 12001-tap filter: 19200 > 16384): block b produces outputs [b B, (b+1) B), B = M - L + 1, from the M inputs that end at its
// last output (zero outside the pulse); the first L - 1 samples of every block's circular convolution are discarded.
template <class P, int PAD>
__global__ void __launch_bounds__(P::NT) k_rda_range_blocked(const float2* __restrict__ in, int64_t in_pitch,
                                                             float2* __restrict__ work, int64_t work_pitch,
                                                             float2* __restrict__ rc_out, int n_rows, int N, int s0, int L,
                                                             int n_blocks, const float2* __restrict__ Hf,
                                                             const float* __restrict__ win, const float2* __restrict__ tw) {
    extern __shared__ float2 sm[];
    constexpr int E = P::E, NT = P::NT, M = P::N;
    const int B = M - L + 1;
    const int t = threadIdx.x;
    const int total = n_rows * n_blocks;
    for (int job = blockIdx.x; job < total; job += gridDim.x) {
        const int row = job / n_blocks, o0 = (job % n_blocks) * B;
        const int in0 = o0 + s0 - (L - 1);
        const float2* p = in + (int64_t)row * in_pitch;
        float2 v[E];
#pragma unroll
        for (int s = 0; s < E; ++s) {
            const int xi = in0 + t + NT * s;
            v[s] = (xi >= 0 && xi < N) ? p[xi] : make_float2(0.f, 0.f);
        }
#pragma unroll 1
        for (int step = 0; step < 2; ++step) {   // one copy of the transform: inverse = conj FFT conj (see k_rda_range)
            transform<P, false, 1, PAD>(v, t, sm, tw);
            if (step == 0) {
#pragma unroll
                for (int s = 0; s < E; ++s) {
                    const float2 h = __ldg(Hf + t + NT * s), x = v[s];
                    v[s] = make_float2(fmaf(x.x, h.x, -x.y * h.y), fmaf(-x.x, h.y, -x.y * h.x));
                }
            }
            __syncthreads();
        }
        const float w = win[row];
#pragma unroll
        for (int s = 0; s < E; ++s) {
            const int j = t + NT * s, o = o0 + j - (L - 1);
            if (j >= L - 1 && o < N) {
                const float2 y = make_float2(v[s].x, -v[s].y);
                if (rc_out != nullptr) rc_out[(int64_t)row * N + o] = y;
                work[(int64_t)row * work_pitch + o] = make_float2(y.x * w, y.y * w);
            }
        }
    }
}

// Pruned variant for N <= M/2 (the common case: the pulse is at most half of the padded length).  The upper half of the
// input is zero, so the M-point spectrum splits into two H = M/2-point transforms of x[n] and x[n] w_M^n (even / odd
// bins), and the M-point inverse into y[n] = A[n mod H] +- w_M^-(n mod H) B[n mod H] with A, B the H-point inverses of the
// even / odd products: four H-point transforms on the 16-elements-per-thread plan instead of two M-point ones on the
// 32-elements-per-thread plan (8192^2 frame, 6001 taps: 1.30 -> 1.09 ms).  HfE / HfO: even / odd bins of the filter
// spectrum / M; twM[n] = w_M^n.  A is parked in shared memory (each thread re-reads only its own elements).
// GPARK (M = 32768, H = 16384: the satellite scripts' 13200-sample pulses and 12001-tap filter need 19200 points, which
// overlap-save on 16384-point blocks pays with four blocks = eight 16384-point transforms per row; this form takes four):
// the row buffer leaves no shared memory for A, which is parked in a per-CTA global scratch line instead -- 128 KB per CTA,
// written and re-read by the same thread, resident in L2.
template <class P, int PAD, bool GPARK>
__global__ void __launch_bounds__(P::NT, (P::E == 32 && P::N == 8192) ? 2 : 1) k_rda_range_pruned(const float2* __restrict__ in, int64_t in_pitch,
                                                            float2* __restrict__ work, int64_t work_pitch,
                                                            float2* __restrict__ rc_out, int n_rows, int N, int s0,
                                                            const float2* __restrict__ HfE, const float2* __restrict__ HfO,
                                                            const float2* __restrict__ twM, const float* __restrict__ win,
                                                            const float2* __restrict__ tw, float2* __restrict__ gpark) {
    extern __shared__ float2 sm[];
    constexpr int E = P::E, NT = P::NT, H = P::N;
    constexpr int SMROW = H + (PAD ? (H >> PAD) : 0);
    float2* const park = GPARK ? gpark + (size_t)blockIdx.x * H : sm + SMROW;
    const int t = threadIdx.x;
    for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
        const float2* p = in + (int64_t)row * in_pitch;
        float2 v[E];
        // even bins, then odd bins, forward and inverse: ONE copy of the transform in the instruction stream.  The inverse is
        // run as conj(FFT(conj(.))) -- the conjugated-twiddle form up to rounding -- with the conjugations folded into the
        // filter multiply and into the uses of the result.  Round 2: with four inlined transform bodies (two per branch) the
        // 32-element plan needed ~700 bytes of local memory per thread at its 128-register cap (ncu: 4.8 GB of DRAM writes
        // for a 0.76 GB pass); one body fits the registers.
        if constexpr (P::E < 32) {
            // (the 16-element plans fit their registers either way, and two bodies schedule a little better: 8192^2 frame
            // 1.86 ms against 1.93 ms rolled)
#pragma unroll 1
            for (int br = 0; br < 2; ++br) {
                const float2* __restrict__ hf = br ? HfO : HfE;
#pragma unroll
                for (int s = 0; s < E; ++s) {
                    const int idx = t + NT * s;
                    float2 x = idx < N ? p[idx] : make_float2(0.f, 0.f);
                    if (br) x = cmul_pk(x, __ldg(twM + idx));
                    v[s] = x;
                }
                transform<P, false, 1, PAD>(v, t, sm, tw);
#pragma unroll
                for (int s = 0; s < E; ++s) {   // v <- conj(v Hf) ...
                    const float2 h = __ldg(hf + t + NT * s), x = v[s];
                    v[s] = make_float2(fmaf(x.x, h.x, -x.y * h.y), fmaf(-x.x, h.y, -x.y * h.x));
                }
                __syncthreads();
                transform<P, false, 1, PAD>(v, t, sm, tw);   // ... so that conj of this forward transform is the inverse
                if (br == 0) {
#pragma unroll
                    for (int s = 0; s < E; ++s) park[t + NT * s] = make_float2(v[s].x, -v[s].y);
                }
                __syncthreads();
            }
        } else
#pragma unroll 1
        for (int step = 0; step < 4; ++step) {
            const int br = step >> 1;
            if ((step & 1) == 0) {
#pragma unroll
                for (int s = 0; s < E; ++s) {
                    const int idx = t + NT * s;
                    float2 x = idx < N ? p[idx] : make_float2(0.f, 0.f);
                    if (br) x = cmul_pk(x, __ldg(twM + idx));
                    v[s] = x;
                }
            }
            transform<P, false, 1, PAD>(v, t, sm, tw);
            if ((step & 1) == 0) {          // v <- conj(v Hf)
                const float2* __restrict__ hf = br ? HfO : HfE;
#pragma unroll
                for (int s = 0; s < E; ++s) {
                    const float2 h = __ldg(hf + t + NT * s), x = v[s];
                    v[s] = make_float2(fmaf(x.x, h.x, -x.y * h.y), fmaf(-x.x, h.y, -x.y * h.x));
                }
            } else if (br == 0) {           // A = conj(v), parked
#pragma unroll
                for (int s = 0; s < E; ++s) park[t + NT * s] = make_float2(v[s].x, -v[s].y);
            }
            __syncthreads();
        }
        const float w = win[row];
#pragma unroll
        for (int s = 0; s < E; ++s) {
            const int idx = t + NT * s;
            const float2 a = park[idx];
            const float2 c = cmul_conj_pk(make_float2(v[s].x, -v[s].y), __ldg(twM + idx));
            const float2 y1 = cadd_pk(a, c), y2 = csub_pk(a, c);
            const int o1 = idx - s0, o2 = idx + H - s0;
            if (o1 >= 0 && o1 < N) {
                if (rc_out != nullptr) rc_out[(int64_t)row * N + o1] = y1;
                work[(int64_t)row * work_pitch + o1] = cscale_pk(y1, w);
            }
            if (o2 >= 0 && o2 < N) {
                if (rc_out != nullptr) rc_out[(int64_t)row * N + o2] = y2;
                work[(int64_t)row * work_pitch + o2] = cscale_pk(y2, w);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------ RCMC + azimuth compression
// One Doppler row per CTA iteration.  The row is pulled into shared memory by ONE 1-D bulk copy (TMA) when its address and
// length allow (16-byte multiples), so no thread waits on a chain of dependent global loads; several CTAs share an SM and
// cover each other's copy latency.
__global__ void __launch_bounds__(256) k_rda_rcmc(float2* __restrict__ work, int64_t pitch, int n_rows, int N, double q0,
                                                  const RdaRow* __restrict__ rows, float2* __restrict__ rd_out,
                                                  float2* __restrict__ rcmc_out, float2* __restrict__ filt_out) {
    extern __shared__ __align__(128) float2 line[];
    __shared__ uint64_t full;
    const bool bulk = ((N & 1) == 0) && ((pitch & 1) == 0) && ((reinterpret_cast<uintptr_t>(work) & 15) == 0);
    if (threadIdx.x == 0) {
        tma::mbar_init(&full, 1);
        tma::fence_barrier_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
        float2* p = work + (int64_t)row * pitch;
        const RdaRow rr = rows[row];
        if (bulk) {
            if (threadIdx.x == 0) {
                tma::mbar_arrive_expect_tx(&full, (uint32_t)N * sizeof(float2));
                tma::bulk_load_1d(line, p, (uint32_t)N * sizeof(float2), &full);
            }
            tma::mbar_wait(&full, phase);
            phase ^= 1;
        } else {
#pragma unroll 8
            for (int i = threadIdx.x; i < N; i += blockDim.x) line[i] = p[i];
            __syncthreads();
        }
        const int64_t eo = (int64_t)rr.export_row * N;
#pragma unroll 4
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const double pos = fma((double)i + q0, rr.beta, (double)i);   // where sample i comes from on the shifted axis
            float2 o = make_float2(0.f, 0.f);
            if (pos >= 0.0 && pos <= (double)(N - 1)) {
                int lo = (int)pos;
                if (lo > N - 2) lo = N - 2;
                const float f = (float)(pos - (double)lo);
                const float2 y0 = line[lo], y1 = line[lo + 1];
                o = make_float2(fmaf(f, y1.x - y0.x, y0.x), fmaf(f, y1.y - y0.y, y0.y));
            }
            const float2 h = cis_u64(rr.ph_b + rr.ph_a * (uint64_t)i);
            const float2 g = cmul(o, h);
            if (rd_out != nullptr) rd_out[eo + i] = cmul(line[i], rr.factor);
            if (rcmc_out != nullptr) rcmc_out[eo + i] = cmul(o, rr.factor);
            if (filt_out != nullptr) filt_out[eo + i] = cmul(g, rr.factor);
            p[i] = g;
        }
        tma::fence_proxy_async();   // this thread's reads of the line are ordered before the next bulk copy into it
        __syncthreads();
    }
}

// mag[c][r] = scale * |in[r][c]|  (in: rows x cols)
__global__ void __launch_bounds__(256) k_transpose_mag(const float2* __restrict__ in, int rows, int cols,
                                                       float* __restrict__ mag, float scale) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8) {
        const int r = r0 + k, c = c0 + tx;
        if (r < rows && c < cols) {
            const float2 x = in[(int64_t)r * cols + c];
            tile[k][tx] = scale * sqrtf(fmaf(x.x, x.x, x.y * x.y));
        }
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int c = c0 + k, r = r0 + tx;
        if (r < rows && c < cols) mag[(int64_t)c * rows + r] = tile[tx][k];
    }
}

// ------------------------------------------------------------------------------ host
using P256 = Plan<256, 16, 16, 16, 1>;
using P1024 = Plan<1024, 16, 16, 8, 8>;
using P2048 = Plan<2048, 16, 16, 16, 8>;
using P4096 = Plan<4096, 16, 16, 16, 16>;
using P8192 = Plan<8192, 16, 16, 8, 8, 8>;
using P16384 = Plan<16384, 32, 32, 32, 16>;
using P8192E32 = Plan<8192, 32, 32, 16, 16>;
using P4096E32 = Plan<4096, 32, 32, 16, 8>;

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

// scipy.signal.windows.hamming(m) (symmetric)
std::vector<double> hamming_sym(int m) {
    std::vector<double> w(m, 1.0);
    const double two_pi = 6.283185307179586476925286766559;
    if (m > 1)
        for (int n = 0; n < m; ++n) w[n] = 0.54 - 0.46 * cos(two_pi * (double)n / (double)(m - 1));
    return w;
}

// number of matched-filter taps, with the reference's own fp64 expression (:380-381)
int mf_taps(const nis_rda_params& prm) {
    const double step = 1.0 / prm.fs;
    return (int)floor(prm.t_p / step) + 1;
}

// One zero-padded 32768-point block, evaluated as four 16384-point transforms: possible when the pulse fits the lower half
// and the 'same' window stays free of wrap-around; worth it when overlap-save on 16384-point blocks would need > 2 blocks.
bool wide_pruned_ok(int n, int taps) {
    if (getenv("NIS_RDA_NOWIDE")) return false;
    if (taps < 1 || n > 16384 || taps > 32768 || n + taps - 1 - (taps - 1) / 2 > 32768) return false;
    if (taps > 14337) return true;
    const int B = 16384 - taps + 1;
    return (n + B - 1) / B > 2;
}

int conv_fft_len(int n, int taps) {   // smallest supported M that keeps the 'same' window free of wrap-around
    const int need = n + taps - 1 - (taps - 1) / 2;
    const int ms[] = {256, 1024, 2048, 4096, 8192, 16384};
    for (int m : ms)
        if (m >= need && m >= taps) return m;
    return 0;
}

}  // namespace

struct nis_rda_plan {
    nis_ctx* ctx = nullptr;
    int P = 0, S = 0, taps = 0, M = 0, s0 = 0;
    nis_rda_params prm{};
    nis_csa_plan* az = nullptr;       // four-step azimuth engine + workspace [P][S] (power-of-two P)
    RowDft* dft = nullptr;            // row-DFT engine (other P): needs the transposed buffer
    float2 *work = nullptr, *tbuf = nullptr, *Hf = nullptr, *tw = nullptr;
    float2 *HfE = nullptr, *HfO = nullptr, *twM = nullptr;   // pruned range compression (N <= M/2)
    float2* gpark = nullptr;                                  // M = 32768: one 16384-point scratch line per CTA
    int gpark_lines = 0;
    float* win = nullptr;
    RdaRow* rows = nullptr;
    double q0 = 0;
    std::vector<double> range_axis_centered, cross_range, doppler;
    int (*range_fn)(nis_rda_plan*, const float2*, int64_t, float2*, cudaStream_t) = nullptr;
};

namespace {

template <class P, int PAD>
int launch_rda_range(nis_rda_plan* pl, const float2* in, int64_t pitch, float2* rc_out, cudaStream_t st) {
    constexpr int SMROW = P::N + (PAD ? (P::N >> PAD) : 0);
    const size_t smem = (size_t)SMROW * sizeof(float2);
    static bool attr_done_dev[64] = {};
    bool& attr_done = attr_done_dev[nis::current_device() & 63];
    if (!attr_done) {
        NIS_CUDA_TRY(cudaFuncSetAttribute(k_rda_range<P, PAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    int per_sm = 1;
    NIS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rda_range<P, PAD>, P::NT, smem));
    int grid = pl->ctx->num_sms * (per_sm < 1 ? 1 : per_sm);
    if (grid > pl->P) grid = pl->P;
    k_rda_range<P, PAD><<<grid, P::NT, smem, st>>>(in, pitch, pl->work, pl->S, rc_out, pl->P, pl->S, pl->s0, pl->Hf,
                                                    pl->win, pl->tw);
    NIS_LAUNCH_CHECK(pl->ctx);
    return NIS_OK;
}

template <class P, int PAD>
int launch_rda_range_blocked(nis_rda_plan* pl, const float2* in, int64_t pitch, float2* rc_out, cudaStream_t st) {
    constexpr int SMROW = P::N + (PAD ? (P::N >> PAD) : 0);
    const size_t smem = (size_t)SMROW * sizeof(float2);
    static bool attr_done_dev[64] = {};
    bool& attr_done = attr_done_dev[nis::current_device() & 63];
    if (!attr_done) {
        NIS_CUDA_TRY(cudaFuncSetAttribute(k_rda_range_blocked<P, PAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    const int B = P::N - pl->taps + 1, n_blocks = (pl->S + B - 1) / B;
    int per_sm = 1;
    NIS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rda_range_blocked<P, PAD>, P::NT, smem));
    const int jobs = pl->P * n_blocks;
    int grid = pl->ctx->num_sms * (per_sm < 1 ? 1 : per_sm);
    if (grid > jobs) grid = jobs;
    k_rda_range_blocked<P, PAD><<<grid, P::NT, smem, st>>>(in, pitch, pl->work, pl->S, rc_out, pl->P, pl->S, pl->s0, pl->taps,
                                                            n_blocks, pl->Hf, pl->win, pl->tw);
    NIS_LAUNCH_CHECK(pl->ctx);
    return NIS_OK;
}

template <class P, int PAD, bool GPARK>
int launch_rda_range_pruned(nis_rda_plan* pl, const float2* in, int64_t pitch, float2* rc_out, cudaStream_t st) {
    constexpr int SMROW = P::N + (PAD ? (P::N >> PAD) : 0);
    const size_t smem = (size_t)(SMROW + (GPARK ? 0 : P::N)) * sizeof(float2);
    static bool attr_done_dev[64] = {};
    bool& attr_done = attr_done_dev[nis::current_device() & 63];
    if (!attr_done) {
        NIS_CUDA_TRY(cudaFuncSetAttribute(k_rda_range_pruned<P, PAD, GPARK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    int per_sm = 1;
    NIS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rda_range_pruned<P, PAD, GPARK>, P::NT, smem));
    int grid = pl->ctx->num_sms * (per_sm < 1 ? 1 : per_sm);
    if (grid > pl->P) grid = pl->P;
    if (GPARK && grid > pl->gpark_lines) grid = pl->gpark_lines;
    k_rda_range_pruned<P, PAD, GPARK><<<grid, P::NT, smem, st>>>(in, pitch, pl->work, pl->S, rc_out, pl->P, pl->S, pl->s0,
                                                                  pl->HfE, pl->HfO, pl->twM, pl->win, pl->tw, pl->gpark);
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

bool rda_sizes_ok(int P, int S, const nis_rda_params& prm, int* M_out) {
    if (P < 2 || S < 2 || S > 24576) return false;
    const int taps = mf_taps(prm);
    int M = conv_fft_len(S, taps);
    if (M == 0 && wide_pruned_ok(S, taps)) M = 32768;       // one 32768-point block in the pruned form (S <= 16384)
    if (M == 0 && taps >= 1 && taps <= 14337) M = 16384;   // overlap-save blocks of >= 2048 outputs
    if (taps < 1 || M == 0) return false;
    if (M_out) *M_out = M;
    return (az_engine_supported(P, S)) || (rowdft_supported(P));
}

}  // namespace

extern "C" int nis_rda_supported(int32_t n_pulses, int32_t n_ranges, const nis_rda_params* prm) {
    if (!prm || !(prm->fs > 0) || !(prm->t_p > 0)) return 0;
    return rda_sizes_ok(n_pulses, n_ranges, *prm, nullptr) ? 1 : 0;
}

extern "C" int nis_rda_plan_destroy(nis_rda_plan* pl) {
    if (!pl) return NIS_OK;
    if (pl->az) {
        if (pl->work == pl->az->work) pl->work = nullptr;
        nis_csa_plan_destroy(pl->az);
    }
    rowdft_destroy(pl->dft);
    cudaFree(pl->work);
    cudaFree(pl->tbuf);
    cudaFree(pl->Hf);
    cudaFree(pl->HfE);
    cudaFree(pl->gpark);
    cudaFree(pl->HfO);
    cudaFree(pl->twM);
    cudaFree(pl->tw);
    cudaFree(pl->win);
    cudaFree(pl->rows);
    delete pl;
    return NIS_OK;
}

extern "C" int nis_rda_plan_create(nis_ctx* ctx, int32_t P, int32_t S, const nis_rda_params* prm, nis_rda_plan** out) {
    NIS_REQUIRE(ctx && prm && out, "nis_rda_plan_create: null argument");
    NIS_REQUIRE(prm->fs > 0 && prm->prf > 0 && prm->vr > 0 && prm->lambda > 0 && prm->c > 0 && prm->t_p > 0,
                "nis_rda_plan_create: non-physical parameters");
    int M = 0;
    if (!rda_sizes_ok(P, S, *prm, &M)) {
        set_error("nis_rda_plan_create: %d pulses x %d samples with a %d-tap matched filter is not supported (taps <= "
                  "14337, or <= 32768 with samples <= 16384 and samples + taps / 2 <= 32768; samples <= 24576; pulses: power of two 64..32768 with samples %% 32 == 0, or any length the row-DFT "
                  "engine takes)", P, S, mf_taps(*prm));
        return NIS_ERR_UNSUPPORTED;
    }
    DeviceGuard device_guard(ctx->device);   // the caller's current device is restored on return
    nis_rda_plan* pl = new nis_rda_plan();
    pl->ctx = ctx;
    pl->P = P;
    pl->S = S;
    pl->prm = *prm;
    pl->taps = mf_taps(*prm);
    pl->M = M;
    pl->s0 = (pl->taps - 1) / 2;
    int rc = NIS_OK;
#define FAIL_IF(x) do { rc = (x); if (rc != NIS_OK) { nis_rda_plan_destroy(pl); return rc; } } while (0)
#define CUDA_FAIL_IF(x) FAIL_IF((x) == cudaSuccess ? NIS_OK : (set_error("%s failed", #x), NIS_ERR_CUDA))
    const double c = prm->c, lam = prm->lambda, vr = prm->vr;
    const int h_az = P / 2;   // fftshift offset along pulses (n // 2 for even and odd n)

    // ---- axes (:363-375, :400-406, :441-443)
    pl->cross_range.resize(P);
    pl->doppler.resize(P);
    for (int i = 0; i < P; ++i) {
        const double centre = (P % 2 == 0) ? (double)P / 2 : (double)(P - 1) / 2;
        pl->cross_range[i] = vr * (((double)i - centre) / prm->prf);
        pl->doppler[i] = ((double)i - centre) * (prm->prf / (double)P);
    }
    const double t_grp = 2 * prm->range_grp / c;
    const double centre_r = (S % 2 == 0) ? (double)S / 2 : (double)(S - 1) / 2;
    std::vector<double> rax(S);
    long double acc = 0;
    for (int i = 0; i < S; ++i) {
        rax[i] = (((double)i - centre_r) / prm->fs + t_grp) * c / 2;
        acc += rax[i];
    }
    const double mean = (double)(acc / S);
    pl->range_axis_centered.resize(S);
    for (int i = 0; i < S; ++i) pl->range_axis_centered[i] = rax[i] - mean;
    pl->q0 = (double)((long double)t_grp * (long double)prm->fs - (long double)centre_r);

    // ---- matched filter (:379-386) -> FFT_M / M
    {
        const int L = pl->taps;
        std::vector<std::complex<double>> hpad(M, 0.0);
        const std::vector<double> wh = hamming_sym(L);
        const double start = -prm->t_p / 2, stop = prm->t_p / 2;
        const double step = (L > 1) ? (stop - start) / (double)(L - 1) : 0.0;
        double norm = 0;
        for (int i = 0; i < L; ++i) {
            double t = (double)i * step + start;    // numpy.linspace
            if (i == L - 1 && L > 1) t = stop;
            const double ph = M_PI * prm->kr * (t * t);
            hpad[i] = std::conj(std::complex<double>(cos(ph), sin(ph))) * wh[i];
            norm += std::norm(hpad[i]);
        }
        norm = sqrt(norm);
        for (int i = 0; i < L; ++i) hpad[i] /= norm;
        host_fft_pow2(hpad);
        std::vector<float2> hf(M);
        for (int i = 0; i < M; ++i) hf[i] = make_float2((float)(hpad[i].real() / M), (float)(hpad[i].imag() / M));
        CUDA_FAIL_IF(cudaMalloc(&pl->Hf, M * sizeof(float2)));
        CUDA_FAIL_IF(cudaMemcpy(pl->Hf, hf.data(), M * sizeof(float2), cudaMemcpyHostToDevice));
    }
    // measured: the pruned form wins only where the unpruned one needs the 32-elements-per-thread plan (M = 16384: 1.30 ->
    // 1.09 ms at 8192 rows); at M = 8192 its 256-thread CTAs run four dependent transforms per row and lose (0.22 -> 0.45 ms)
    const bool wide = M == 32768;
    const bool blocked = !wide && conv_fft_len(S, pl->taps) == 0;   // the filter does not fit one block with the pulse
    const bool prune = wide || (!blocked && (2 * S <= M) && M == 16384 && !getenv("NIS_RDA_NOPRUNE"));
    if (prune) {
        const int H = M / 2;
        std::vector<float2> hf(M), he(H), ho(H), twm(H);
        NIS_CUDA_TRY(cudaMemcpy(hf.data(), pl->Hf, M * sizeof(float2), cudaMemcpyDeviceToHost));
        const double two_pi = 6.283185307179586476925286766559;
        for (int k = 0; k < H; ++k) {
            he[k] = hf[2 * k];
            ho[k] = hf[2 * k + 1];
            const double a = -two_pi * (double)k / (double)M;
            twm[k] = make_float2((float)cos(a), (float)sin(a));
        }
        CUDA_FAIL_IF(cudaMalloc(&pl->HfE, H * sizeof(float2)));
        CUDA_FAIL_IF(cudaMalloc(&pl->HfO, H * sizeof(float2)));
        CUDA_FAIL_IF(cudaMalloc(&pl->twM, H * sizeof(float2)));
        CUDA_FAIL_IF(cudaMemcpy(pl->HfE, he.data(), H * sizeof(float2), cudaMemcpyHostToDevice));
        CUDA_FAIL_IF(cudaMemcpy(pl->HfO, ho.data(), H * sizeof(float2), cudaMemcpyHostToDevice));
        CUDA_FAIL_IF(cudaMemcpy(pl->twM, twm.data(), H * sizeof(float2), cudaMemcpyHostToDevice));
    }
    if (blocked) {
        pl->range_fn = launch_rda_range_blocked<P16384, 5>;
        FAIL_IF(upload_tw<P16384>(&pl->tw));
    } else if (wide) {
        pl->gpark_lines = ctx->num_sms;
        CUDA_FAIL_IF(cudaMalloc(&pl->gpark, (size_t)pl->gpark_lines * 16384 * sizeof(float2)));
        pl->range_fn = launch_rda_range_pruned<P16384, 5, true>;
        FAIL_IF(upload_tw<P16384>(&pl->tw));
    } else if (prune && !getenv("NIS_RDA_E16")) {
        // two CTAs per SM: 32 samples per thread on the three-pass plan, one transform body at <= 128 registers, A parked in an
        // L2-resident scratch line per CTA instead of shared memory (which would hold one CTA only)
        pl->gpark_lines = 2 * ctx->num_sms;
        CUDA_FAIL_IF(cudaMalloc(&pl->gpark, (size_t)pl->gpark_lines * 8192 * sizeof(float2)));
        pl->range_fn = launch_rda_range_pruned<P8192E32, 5, true>;
        FAIL_IF(upload_tw<P8192E32>(&pl->tw));
    } else if (prune) {
        pl->range_fn = launch_rda_range_pruned<P8192, 4, false>;
        FAIL_IF(upload_tw<P8192>(&pl->tw));
    } else
    switch (M) {
        case 256: pl->range_fn = launch_rda_range<P256, 4>; FAIL_IF(upload_tw<P256>(&pl->tw)); break;
        case 1024: pl->range_fn = launch_rda_range<P1024, 4>; FAIL_IF(upload_tw<P1024>(&pl->tw)); break;
        case 2048: pl->range_fn = launch_rda_range<P2048, 4>; FAIL_IF(upload_tw<P2048>(&pl->tw)); break;
        case 4096:   // 32 samples per thread (32 x 16 x 8), 128 threads, four CTAs per SM -- as k_range_rolled at 4096 samples
            if (getenv("NIS_RDA_E16")) { pl->range_fn = launch_rda_range<P4096, 4>; FAIL_IF(upload_tw<P4096>(&pl->tw)); }
            else { pl->range_fn = launch_rda_range<P4096E32, 5>; FAIL_IF(upload_tw<P4096E32>(&pl->tw)); }
            break;
        case 8192:
            if (getenv("NIS_RDA_E16")) { pl->range_fn = launch_rda_range<P8192, 4>; FAIL_IF(upload_tw<P8192>(&pl->tw)); }
            else { pl->range_fn = launch_rda_range<P8192E32, 5>; FAIL_IF(upload_tw<P8192E32>(&pl->tw)); }
            break;
        default: pl->range_fn = launch_rda_range<P16384, 5>; FAIL_IF(upload_tw<P16384>(&pl->tw)); break;
    }
    // ---- azimuth Hamming weights (:396)
    {
        const std::vector<double> wa = hamming_sym(P);
        std::vector<float> wf(P);
        for (int i = 0; i < P; ++i) wf[i] = (float)wa[i];
        CUDA_FAIL_IF(cudaMalloc(&pl->win, P * sizeof(float)));
        CUDA_FAIL_IF(cudaMemcpy(pl->win, wf.data(), P * sizeof(float), cudaMemcpyHostToDevice));
    }
    // ---- azimuth engine
    int A1 = 1, A2 = P;
    if (az_engine_supported(P, S)) {
        FAIL_IF(az_engine_create(ctx, P, S, &pl->az));
        pl->work = pl->az->work;
        A1 = pl->az->A1;
        A2 = pl->az->A2;
    } else {
        FAIL_IF(rowdft_create(P, &pl->dft));
        CUDA_FAIL_IF(cudaMalloc(&pl->work, (size_t)P * S * sizeof(float2)));
        CUDA_FAIL_IF(cudaMalloc(&pl->tbuf, (size_t)P * S * sizeof(float2)));
    }
    // ---- per Doppler row: RCMC scale, azimuth-compression phase ramp, export placement (:408-435)
    {
        std::vector<RdaRow> h(P);
        const long double dr = (long double)c / (2.0L * (long double)prm->fs);
        const double two_pi = 6.283185307179586476925286766559;
        for (int rho = 0; rho < P; ++rho) {
            const int k1 = rho / A2, k2 = rho % A2;
            const int kk = k1 + A1 * k2;                 // FFT bin held by this row
            const int j = (kk + h_az) % P;               // its row in fftshift order
            const double fd = pl->doppler[j];
            const long double alpha = (long double)fd * fd * (long double)lam * lam / (8.0L * (long double)vr * vr);
            RdaRow r{};
            r.beta = (double)(alpha / (1.0L - alpha));
            const long double a = -((long double)fd * fd * (long double)lam / (4.0L * (long double)vr * vr)) * dr;
            r.ph_a = to_fix(a);
            r.ph_b = to_fix(a * (long double)pl->q0);
            const double ang = -two_pi * (double)(((int64_t)h_az * kk) % P) / (double)P;
            r.factor = make_float2((float)cos(ang), (float)sin(ang));
            r.export_row = j;
            h[rho] = r;
        }
        CUDA_FAIL_IF(cudaMalloc(&pl->rows, P * sizeof(RdaRow)));
        CUDA_FAIL_IF(cudaMemcpy(pl->rows, h.data(), P * sizeof(RdaRow), cudaMemcpyHostToDevice));
    }
#undef CUDA_FAIL_IF
#undef FAIL_IF
    *out = pl;
    return NIS_OK;
}

extern "C" int nis_rda_axes(const nis_rda_plan* pl, double* range_axis_centered, double* cross_range, double* doppler) {
    NIS_REQUIRE(pl, "nis_rda_axes: null plan");
    if (range_axis_centered) memcpy(range_axis_centered, pl->range_axis_centered.data(), pl->S * sizeof(double));
    if (cross_range) memcpy(cross_range, pl->cross_range.data(), pl->P * sizeof(double));
    if (doppler) memcpy(doppler, pl->doppler.data(), pl->P * sizeof(double));
    return NIS_OK;
}

extern "C" int nis_rda_focus(nis_rda_plan* pl, const nis_c32* phist, int64_t pitch, float* image_mag, nis_c32* rc_out,
                             nis_c32* rd_out, nis_c32* rcmc_out, nis_c32* filt_out, nis_stream stream) {
    NIS_REQUIRE(pl && phist && image_mag, "nis_rda_focus: null argument");
    NIS_REQUIRE(pitch >= pl->S, "nis_rda_focus: pitch %lld < samples per pulse %d", (long long)pitch, pl->S);
    cudaStream_t st = (cudaStream_t)stream;
    nis_ctx* ctx = pl->ctx;
    const int P = pl->P, S = pl->S;
    int rc;
#define RUN(x) do { if ((rc = (x)) != NIS_OK) return rc; } while (0)
    RUN(pl->range_fn(pl, reinterpret_cast<const float2*>(phist), pitch, reinterpret_cast<float2*>(rc_out), st));
    if (pl->az) {
        RUN(pl->az->outer_fwd(pl->az, pl->work, S, 0, S, st));   // in place: a thread reads and writes the same 16 cells
        RUN(pl->az->inner(pl->az, false, 0, S, 0, pl->az->A1, st));
    } else {
        RUN(launch_transpose(ctx, pl->work, S, pl->tbuf, P, S, st));
        RUN(rowdft_run(ctx, pl->dft, pl->tbuf, P, S, false, 1.f, st));
        RUN(launch_transpose(ctx, pl->tbuf, P, pl->work, S, P, st));
    }
    {
        static bool attr_done_dev[64] = {};
    bool& attr_done = attr_done_dev[nis::current_device() & 63];
        if (!attr_done) {
            NIS_CUDA_TRY(cudaFuncSetAttribute(k_rda_rcmc, cudaFuncAttributeMaxDynamicSharedMemorySize, 24576 * 8));
            attr_done = true;
        }
        const size_t smem = (size_t)S * sizeof(float2);
        int per_sm = 1;
        NIS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rda_rcmc, 256, smem));
        int grid = ctx->num_sms * (per_sm < 1 ? 1 : per_sm);
        if (grid > P) grid = P;
        k_rda_rcmc<<<grid, 256, smem, st>>>(pl->work, S, P, S, pl->q0, pl->rows, reinterpret_cast<float2*>(rd_out),
                                            reinterpret_cast<float2*>(rcmc_out), reinterpret_cast<float2*>(filt_out));
        NIS_LAUNCH_CHECK(ctx);
    }
    const float scale = (float)(1.0 / (double)P);
    if (pl->az) {
        RUN(pl->az->inner(pl->az, true, 0, S, 0, pl->az->A1, st));
        RUN(pl->az->outer_inv_mag(pl->az, image_mag, scale, st));
    } else {
        RUN(launch_transpose(ctx, pl->work, S, pl->tbuf, P, S, st));
        RUN(rowdft_run(ctx, pl->dft, pl->tbuf, P, S, true, 1.f, st));
        dim3 grid((P + 31) / 32, (S + 31) / 32);
        k_transpose_mag<<<grid, 256, 0, st>>>(pl->tbuf, S, P, image_mag, scale);
        NIS_LAUNCH_CHECK(ctx);
    }
#undef RUN
    return NIS_OK;
}
