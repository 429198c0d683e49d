// csa.cu -- K2: Chirp Scaling Algorithm focusing (replaces sar_focus_csa,
// sar_ati_dcpa_sim_csa.py:202-396) for power-of-two scenes.
//
// Data flow (complex64, one workspace W[n_az][n_rg]):
//   k_az_outer_fwd   raw  -> W   radix-A1 butterflies down the columns (registers only) x w_N^(a2 k1)
//   k_az_inner_tma<fwd> W -> W  A2-point column FFTs on TMA-prefetched [A2 rows x 8..32 cols] tiles in shared memory
//   k_range          W   -> W   per Doppler row: x Phi1, range FFT, x Phi2, range IFFT, x Phi3
//   k_az_inner_tma<inv> W -> W  A2-point inverse column FFTs
//   k_az_outer_inv   W   -> slc x w_N^-(k1 a2), radix-A1 inverse butterflies, 1/(n_az n_rg), corner turn to [n_rg][n_az]
// Azimuth length n_az = A1*A2 (four-step split).  Between the two azimuth transforms rows live in
// the permuted order rho = k1*A2 + k2 <-> Doppler bin kk = k1 + A1*k2; every per-row quantity is
// tabulated in that order, and the reference's fftshift / ifftshift pairs (:234, :280, :331, :385)
// reduce to index relabelling (bin kk has frequency fftfreq[kk]).
//
// Phase functions are quadratic in the sample (or frequency) index with per-row coefficients:
//   Phi1: -(Kr Cs/2) (tau_n - tau_ref)^2                      tau_n = t0 + n/fs        (:272)
//   Phi2:  fr^2 / (2 Kr (1+Cs)) + 2 R_ref Cs fr / c            fr = k' fs/n_rg          (:318-324)
//   Phi3:  c tau_n D / lam - (Kr Cs (1+Cs)/2) (tau_n - 2R_ref/c)^2                      (:359-380)
// (in turns).  The coefficients are derived on the host in extended precision from the reference's
// own fp64 D, Cs, tau_ref and converted to 64-bit fixed-point turns, so that 1e4..1e8 rad phases are
// reduced mod 1 exactly by integer wrap-around on the device.
#include <math.h>
#include <string.h>

#include <vector>

#include <stdlib.h>

#include <type_traits>

#include "csa_internal.cuh"
#include "fft.cuh"
#include "tma.cuh"

using namespace nis;
using namespace nis::fft;
using namespace nis::csa;

namespace {

// ------------------------------------------------------------------------------ azimuth, outer
// Forward: Y[k1*A2 + a2][n] = w_N^(a2 k1) * sum_a1 x[a1*A2 + a2][n] w_A1^(a1 k1)
template <int A1>
__global__ void __launch_bounds__(256) k_az_outer_fwd(const float2* __restrict__ in, int64_t in_pitch,
                                                      float2* __restrict__ out, int64_t out_pitch, int n_rg,
                                                      int A2, int n_az, const float2* __restrict__ twN) {
    const int n = blockIdx.x * 32 + (threadIdx.x & 31);
    const int a2 = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (n >= n_rg || a2 >= A2) return;
    float2 v[A1];
#pragma unroll
    for (int a1 = 0; a1 < A1; ++a1) v[a1] = in[(int64_t)(a1 * A2 + a2) * in_pitch + n];
    fft_dif<A1, false, 1, true>(v);
    constexpr int L = ilog2(A1);
#pragma unroll
    for (int k1 = 0; k1 < A1; ++k1) {
        float2 x = v[brev(k1, L)];
        if (k1 > 0) x = cmul_pk(x, __ldg(twN + ((a2 * k1) & (n_az - 1))));
        out[(int64_t)(k1 * A2 + a2) * out_pitch + n] = x;
    }
}

// Inverse + corner turn: slc[n][a1*A2 + a2] = scale * sum_k1 Z[k1*A2 + a2][n] w_A1^-(k1 a1)
// Tile: TN range columns x TA azimuth offsets; transposed through shared memory so that both the reads
// (TN*8-byte row pieces) and the writes (TA*8-byte pieces of an slc row) are coalesced.
template <int A1, int TN, int TA>
__global__ void __launch_bounds__(256) k_az_outer_inv(const float2* __restrict__ in, int64_t pitch,
                                                      float2* __restrict__ slc, int n_rg, int A2, int n_az,
                                                      float scale, double* __restrict__ max_sq,
                                                      const float2* __restrict__ twN) {
    extern __shared__ float2 tile[];  // [A1][TN][TA+1]
    constexpr int AROWS = 256 / TN;   // azimuth offsets covered per iteration
    static_assert(TA % AROWS == 0, "tile height must be a multiple of 256 / TN");
    const int n0 = blockIdx.x * TN, a20 = blockIdx.y * TA;
    const int n_off = threadIdx.x % TN, a_off = threadIdx.x / TN;
    constexpr int L = ilog2(A1);
#pragma unroll
    for (int it = 0; it < TA / AROWS; ++it) {
        const int a2 = a20 + a_off + AROWS * it;
        float2 v[A1];
#pragma unroll
        for (int k1 = 0; k1 < A1; ++k1) {
            float2 x = in[(int64_t)(k1 * A2 + a2) * pitch + n0 + n_off];
            if (k1 > 0) x = cmul_conj_pk(x, __ldg(twN + ((k1 * a2) & (n_az - 1))));   // four-step twiddle w_N^-(k1 a2)
            v[k1] = x;
        }
        fft_dif<A1, true, 1, true>(v);
#pragma unroll
        for (int a1 = 0; a1 < A1; ++a1) {
            float2 x = v[brev(a1, L)];
            tile[(a1 * TN + n_off) * (TA + 1) + a_off + AROWS * it] = make_float2(x.x * scale, x.y * scale);
        }
    }
    __syncthreads();
    double m = 0.0;
    constexpr int LPR = TA;                 // lanes per (a1, n) row piece
    constexpr int RPI = 256 / LPR;          // row pieces per iteration
    const int l = threadIdx.x % LPR, g = threadIdx.x / LPR;
    for (int p = g; p < A1 * TN; p += RPI) {
        const int a1 = p / TN, nn = p % TN;
        float2 x = tile[(a1 * TN + nn) * (TA + 1) + l];
        slc[(int64_t)(n0 + nn) * n_az + a1 * A2 + a20 + l] = x;
        if (max_sq != nullptr) m = fmax(m, sq_mag_f64(x));
    }
    if (max_sq != nullptr) {   // one atomic per CTA (16384 CTAs at 8192^2), not one per warp
        __shared__ double wmax[8];
        m = warp_max_f64(m);
        if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x < 32) {
            m = threadIdx.x < 8 ? wmax[threadIdx.x] : 0.0;
            atomic_max_f64(max_sq, warp_max_f64(m));
        }
    }
}


// Inverse radix-A1 stage WITHOUT the corner turn, magnitude out (RDA image, sar_satellite_sim.py:438-439):
// mag[a1*A2 + a2][n] = scale * | sum_k1 w_N^-(k1 a2) Z[k1*A2 + a2][n] w_A1^-(k1 a1) |
template <int A1>
__global__ void __launch_bounds__(256) k_az_outer_inv_mag(const float2* __restrict__ in, int64_t pitch,
                                                          float* __restrict__ mag, int64_t mag_pitch, int n_rg, int A2,
                                                          int n_az, float scale, const float2* __restrict__ twN) {
    const int n = blockIdx.x * 32 + (threadIdx.x & 31);
    const int a2 = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (n >= n_rg || a2 >= A2) return;
    float2 v[A1];
#pragma unroll
    for (int k1 = 0; k1 < A1; ++k1) {
        float2 x = in[(int64_t)(k1 * A2 + a2) * pitch + n];
        if (k1 > 0) x = cmul_conj_pk(x, __ldg(twN + ((k1 * a2) & (n_az - 1))));
        v[k1] = x;
    }
    fft_dif<A1, true, 1, true>(v);
    constexpr int L = ilog2(A1);
#pragma unroll
    for (int a1 = 0; a1 < A1; ++a1) {
        const float2 x = v[brev(a1, L)];
        mag[(int64_t)(a1 * A2 + a2) * mag_pitch + n] = scale * sqrtf(fmaf(x.x, x.x, x.y * x.y));
    }
}

// ------------------------------------------------------------------------------ azimuth, inner
// A2-point transforms down W adjacent columns of the row block [k1*A2, (k1+1)*A2); lanes <-> columns, so every
// shared-memory access of the transform is a contiguous W*8-byte row piece (bank-conflict free).
// Persistent and TMA-fed: each CTA walks tiles tile = x + n_col_tiles * k1; while tile i is transformed the
// [A2 x W] box of tile i+1 is already in flight into the other shared buffer (cp.async.bulk.tensor.2d, completion
// on an mbarrier).  The tile buffer doubles as the exchange buffer of the transform, so global loads never stall a
// warp and stores leave straight from registers.
// PH: the chirp-scaling multiply Phi1 (:272-274) rides on the forward transform's store and the azimuth-compression /
// residual-phase multiply Phi3 (:380-382) on the inverse transform's load: element (row rho, column n) is multiplied by
// cis(a_rho n^2 + b_rho n + c_rho) from the per-row tables (L1-resident: all W threads of a row read the same 20 bytes).
// (Development arrangement NIS_CSA_PHASE=az; the default keeps both multiplies in k_range -- measured faster.)
template <class P, bool INV, int W, bool PH>
__global__ void __launch_bounds__(P::NT* W) k_az_inner_tma(const __grid_constant__ CUtensorMap map,
                                                           float2* __restrict__ data, int64_t pitch,
                                                           int n_col_tiles, int n_tiles, int x0, int k10,
                                                           const float2* __restrict__ tw,
                                                           const uint4* __restrict__ ph_ab,
                                                           const uint32_t* __restrict__ ph_c) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full[2];
    constexpr int E = P::E, NT = P::NT, A2 = P::N;
    constexpr int TILE_ELEMS = A2 * W;
    constexpr uint32_t TILE_BYTES = TILE_ELEMS * sizeof(float2);
    constexpr int BOX_ROWS = A2 < 256 ? A2 : 256, NBOX = A2 / BOX_ROWS;
    float2* const buf0 = reinterpret_cast<float2*>(smem_raw);
    const int tid = threadIdx.x, c = tid % W, t = tid / W;
    if (tid == 0) {
        tma::mbar_init(&full[0], 1);
        tma::mbar_init(&full[1], 1);
        tma::fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int tile, int b) {
        const int x = x0 + tile % n_col_tiles, k1 = k10 + tile / n_col_tiles;
        tma::mbar_arrive_expect_tx(&full[b], TILE_BYTES);
#pragma unroll
        for (int i = 0; i < NBOX; ++i)
            tma::tile_load_2d(buf0 + b * TILE_ELEMS + i * BOX_ROWS * W, &map, x * W, k1 * A2 + i * BOX_ROWS, &full[b]);
    };
    int tile = blockIdx.x;
    if (tid == 0 && tile < n_tiles) issue(tile, 0);
    for (int i = 0; tile < n_tiles; ++i, tile += gridDim.x) {
        const int b = i & 1;
        float2* buf = buf0 + b * TILE_ELEMS;
        if (tid == 0 && tile + (int)gridDim.x < n_tiles) issue(tile + gridDim.x, b ^ 1);
        const int x = x0 + tile % n_col_tiles, k1 = k10 + tile / n_col_tiles;
        const uint32_t n = (uint32_t)(x * W + c), n2 = n * n;
        tma::mbar_wait(&full[b], (i >> 1) & 1);
        float2 v[E];
#pragma unroll
        for (int s = 0; s < E; ++s) v[s] = buf[(t + NT * s) * W + c];
        if constexpr (PH && INV) {
#pragma unroll
            for (int s = 0; s < E; ++s) {
                const int row = k1 * A2 + t + NT * s;
                v[s] = cmul_pk(v[s], cis_u32_pre(row_phase_hi(__ldg(ph_ab + row), __ldg(ph_c + row), n, n2)));
            }
        }
        __syncthreads();   // every element is in registers before the exchange overwrites the tile
        transform<P, INV, W, 0, CtaBarrier, true>(v, t, buf + c, tw);
        float2* base = data + (int64_t)k1 * A2 * pitch + x * W + c;
#pragma unroll
        for (int s = 0; s < E; ++s) {
            if constexpr (PH && !INV) {
                const int row = k1 * A2 + t + NT * s;
                v[s] = cmul_pk(v[s], cis_u32_pre(row_phase_hi(__ldg(ph_ab + row), __ldg(ph_c + row), n, n2)));
            }
            base[(int64_t)(t + NT * s) * pitch] = v[s];
        }
        tma::fence_proxy_async();   // this thread's exchange writes are ordered before the TMA refill of buf
        __syncthreads();
    }
}


// ------------------------------------------------------------------------------ azimuth, whole columns per cluster
// One cluster of C CTAs transforms all n_az = C*M samples of W adjacent columns in ONE pass over HBM (the two-kernel
// four-step above moves every sample twice).  Decimation in time across the cluster:
//     X[k2 + M k1] = sum_n1 w_C^(n1 k1) * ( w_N^(n1 k2) * sum_n2 x[n1 + C n2] w_M^(n2 k2) )
// CTA n1 pulls rows n1, n1 + C, ... of the column tile with one strided TMA box sequence (3-D tensor map), runs the
// M-point transforms in registers / its own shared memory, applies the four-step twiddle and leaves Z'_n1[k2] in shared
// memory.  After a cluster barrier CTA r gathers, for its slice k2 in [r M/C, (r+1) M/C), the C partial spectra through
// distributed shared memory (each value crosses the SM-to-SM network once), finishes with radix-C butterflies in
// registers and stores rows k2 + M k1 (natural Doppler order) straight to HBM -- or, for the inverse transform, the
// scaled, corner-turned image slc[column][azimuth] with azimuth-contiguous 256-byte stores, plus max |slc|^2.
// Several CTAs of different clusters share an SM, so one cluster's load / barrier latency hides behind another's math.
template <class P, int C, int W, bool INV, bool TOUT, bool PH>
__global__ void __launch_bounds__(P::NT* W, (P::E >= 32 ? 512 : 1024) / (P::NT * W)) k_az_cluster(const __grid_constant__ CUtensorMap map,
                                                         float2* __restrict__ out, int64_t out_pitch, int n_col_tiles,
                                                         float scale, double* __restrict__ max_sq,
                                                         const float2* __restrict__ tw, const float2* __restrict__ twN,
                                                         const uint4* __restrict__ ph_ab, const uint32_t* __restrict__ ph_c) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full;
    constexpr int E = P::E, NT = P::NT, M = P::N, NTH = NT * W;
    constexpr int SLICE = M / C;              // k2 values finished by one CTA
    constexpr int JN = SLICE * W / NTH;       // (k2, column) pairs per thread
    static_assert(SLICE * W % NTH == 0 && JN >= 1, "slice must tile the CTA");
    constexpr int ZP = M + 2;                 // padded column length of the corner-turned layout
    constexpr uint32_t TILE_BYTES = M * W * sizeof(float2);
    constexpr int BOX_ROWS = M < 256 ? M : 256, NBOX = M / BOX_ROWS;
    float2* const buf = reinterpret_cast<float2*>(smem_raw);
    const int tid = threadIdx.x, c = tid % W, t = tid / W;
    const uint32_t rank = tma::cluster_ctarank();
    const int cluster_id = blockIdx.x / C, n_clusters = gridDim.x / C;
    if (tid == 0) {
        tma::mbar_init(&full, 1);
        tma::fence_barrier_init();
    }
    __syncthreads();
    double mx = 0.0;
    uint32_t phase = 0;
    bool pending = false;   // a "my reads of the other CTAs' tiles are done" arrival not yet waited for
    for (int x = cluster_id; x < n_col_tiles; x += n_clusters) {
        if (pending) tma::cluster_wait();   // every CTA of the cluster has finished reading this CTA's buffer
        if (tid == 0) {
            tma::mbar_arrive_expect_tx(&full, TILE_BYTES);
#pragma unroll
            for (int i = 0; i < NBOX; ++i)
                tma::tile_load_3d(buf + i * BOX_ROWS * W, &map, x * W, (int)rank, i * BOX_ROWS, &full);
        }
        tma::mbar_wait(&full, phase);
        phase ^= 1;
        float2 v[E];
#pragma unroll
        for (int s = 0; s < E; ++s) v[s] = buf[(t + NT * s) * W + c];
        if constexpr (PH && INV) {   // Phi3 on the load: this CTA holds rows rank + C (t + NT s) of column x W + c
            const uint32_t n = (uint32_t)(x * W + c), n2 = n * n;
#pragma unroll
            for (int s = 0; s < E; ++s) {
                const int row = (int)rank + C * (t + NT * s);
                v[s] = cmul_pk(v[s], cis_u32_pre(row_phase_hi(__ldg(ph_ab + row), __ldg(ph_c + row), n, n2)));
            }
        }
        __syncthreads();   // the tile is in registers: the buffer becomes the exchange buffer
        transform<P, INV, W, 0, CtaBarrier, true>(v, t, buf + c, tw);
        if (rank != 0) {
#pragma unroll
            for (int s = 0; s < E; ++s) {
                const float2 w = __ldg(twN + rank * (uint32_t)(t + NT * s));   // n1 k2 < N
                v[s] = INV ? cmul_conj_pk(v[s], w) : cmul_pk(v[s], w);
            }
        }
        __syncthreads();   // last exchange read by everyone
#pragma unroll
        for (int s = 0; s < E; ++s) {
            const int k2 = t + NT * s;
            buf[TOUT ? c * ZP + k2 : k2 * W + c] = v[s];
        }
        tma::cluster_arrive();   // release: this CTA's partial spectrum is visible to the cluster
        tma::cluster_wait();     // all C partial spectra are in place
        constexpr int L = ilog2(C);
#pragma unroll
        for (int j = 0; j < JN; ++j) {
            const int q = tid + NTH * j;
            int k2, cc;
            if constexpr (TOUT) { k2 = rank * SLICE + q % SLICE; cc = q / SLICE; }
            else { k2 = rank * SLICE + q / W; cc = q % W; }
            const uint32_t local = tma::smem_u32(buf + (TOUT ? cc * ZP + k2 : k2 * W + cc));
            float2 u[C];
#pragma unroll
            for (int n1 = 0; n1 < C; ++n1) u[n1] = tma::ld_cluster_f2(tma::map_to_rank(local, (uint32_t)n1));
            fft_dif<C, INV, 1, true>(u);
            if constexpr (TOUT) {
                float2* o = out + (int64_t)(x * W + cc) * out_pitch + k2;
#pragma unroll
                for (int a1 = 0; a1 < C; ++a1) {
                    const float2 r = make_float2(u[brev(a1, L)].x * scale, u[brev(a1, L)].y * scale);
                    o[a1 * M] = r;
                    if (max_sq != nullptr) mx = fmax(mx, sq_mag_f64(r));
                }
            } else {
                float2* o = out + (int64_t)k2 * out_pitch + x * W + cc;
                const uint32_t n = (uint32_t)(x * W + cc), n2 = n * n;
#pragma unroll
                for (int k1 = 0; k1 < C; ++k1) {
                    float2 r = u[brev(k1, L)];
                    if constexpr (PH) {   // Phi1 on the store: Doppler row k2 + M k1 (natural order)
                        const int row = k2 + M * k1;
                        r = cmul_pk(r, cis_u32_pre(row_phase_hi(__ldg(ph_ab + row), __ldg(ph_c + row), n, n2)));
                    }
                    o[(int64_t)(k1 * M) * out_pitch] = r;
                }
            }
        }
        tma::fence_proxy_async();   // generic-proxy traffic on the buffers is ordered before the next TMA refill
        // "my reads of your buffers are done": nothing to publish, so no release fence -- a releasing arrive would also
        // wait for the global stores just issued (MEMBAR.GPU), 7 % of the kernel when measured
        tma::cluster_arrive_relaxed();
        pending = true;
    }
    if (pending) tma::cluster_wait();   // nobody may exit while a neighbour still reads its shared memory
    if (TOUT && max_sq != nullptr) atomic_max_f64(max_sq, warp_max_f64(mx));
}

// ------------------------------------------------------------------------------ range
// Quadratic phase a n^2 + b n + c (64-bit fixed-point turns) stepped along n = n0, n0+step, ... by
// second differences: two 64-bit adds per sample, exact (integer wrap-around == mod 1).
struct PhaseStepper {
    uint32_t ph, d, dd;   // top 32 bits of the exact 64-bit values: (E^2/2) 2^-32 turns of drift at most
    static constexpr uint32_t kRound = 0x100u;   // cis_u32_pre() truncates to 23 bits: pre-add half an ulp once
    __device__ __forceinline__ void init(uint64_t a, uint64_t b, uint64_t c, uint32_t n0, uint32_t step) {
        ph = (uint32_t)((a * (uint64_t)(n0 * n0) + b * (uint64_t)n0 + c) >> 32) + kRound;
        d = (uint32_t)((a * (uint64_t)(2u * n0 * step + step * step) + b * (uint64_t)step) >> 32);
        dd = (uint32_t)((a * (uint64_t)(2u * step * step)) >> 32);
    }
    // descending argument m0, m0-step, ... of the quadratic term with an ascending linear term
    // (negative range-frequency bins of Phi2: a (N-k)^2 + b k - b N)
    __device__ __forceinline__ void init_mirror(uint64_t a, uint64_t b, uint64_t bn, uint32_t m0, uint32_t k0,
                                                uint32_t step) {
        ph = (uint32_t)((a * (uint64_t)(m0 * m0) + b * (uint64_t)k0 - bn) >> 32) + kRound;
        d = (uint32_t)((a * (uint64_t)(step * step) - a * (uint64_t)(2u * m0 * step) + b * (uint64_t)step) >> 32);
        dd = (uint32_t)((a * (uint64_t)(2u * step * step)) >> 32);
    }
    __device__ __forceinline__ float2 next() {
        const float2 w = cis_u32_pre(ph);
        ph += d;
        d += dd;
        return w;
    }
};

// One Doppler row per group of NT threads: x Phi1 -> FFT -> x Phi2 -> IFFT -> x Phi3, one HBM round
// trip.  RPB independent row groups share a CTA (named barriers, so groups drift apart and overlap
// each other's load / exchange / store phases).
// PHI13: the kernel also applies Phi1 on its load and Phi3 on its store (default).  With NIS_CSA_PHASE=az only Phi2 is applied
// here and Phi1 / Phi3 ride on the azimuth kernels either side (measured slower overall, see nis_csa_plan_create).
template <class P, int PAD, int RPB, int MINB, bool PK, bool PHI13>
__global__ void __launch_bounds__(P::NT* RPB, MINB) k_range(float2* __restrict__ data, int64_t pitch, int n_rows,
                                                            const RowCoef* __restrict__ coef,
                                                            const float2* __restrict__ tw) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full[RPB];
    constexpr int E = P::E, NT = P::NT, N = P::N;
    constexpr int SMROW = N + (PAD ? (N >> PAD) : 0);
    static_assert(E % 2 == 0, "positive / negative frequency halves split the register slots");
    // bar.sync counts whole warps: row groups narrower than a warp (or a single group) use the CTA barrier
    constexpr bool kNamed = (NT >= 32) && (RPB > 1);
    // the next row is prefetched into shared memory by a 1-D bulk copy (TMA) while this one is transformed
    constexpr bool kPrefetch = (NT >= 32) && (N <= 8192);   // a 16384-sample row leaves no room for the prefetch buffer
    constexpr int GROUP_ELEMS = SMROW + (kPrefetch ? N : 0);
    const int t = threadIdx.x;
    float2* sm = reinterpret_cast<float2*>(smem_raw) + threadIdx.y * GROUP_ELEMS;
    float2* pf = sm + SMROW;
    uint64_t* mb = &full[threadIdx.y];
    using Bar = typename std::conditional<kNamed, NamedBarrier, CtaBarrier>::type;
    Bar bar;
    if constexpr (kNamed) bar = NamedBarrier{1 + (int)threadIdx.y, NT};
    // with the CTA barrier every group must run the same number of iterations: the last ones may idle
    const int row_stride = gridDim.x * RPB;
    const int n_iter = (n_rows + row_stride - 1) / row_stride;
    const int row0 = blockIdx.x * RPB + threadIdx.y;
    if constexpr (kPrefetch) {
        if (t == 0) {
            tma::mbar_init(mb, 1);
            tma::fence_barrier_init();
            if (row0 < n_rows) {
                tma::mbar_arrive_expect_tx(mb, N * sizeof(float2));
                tma::bulk_load_1d(pf, data + (int64_t)row0 * pitch, N * sizeof(float2), mb);
            }
        }
        __syncthreads();
    }
    for (int it = 0; it < n_iter; ++it) {
        int row = row0 + it * row_stride;
        const bool live = row < n_rows;
        if (!live) {
            if constexpr (kNamed || (kPrefetch && RPB == 1)) break;
            row = n_rows - 1;
        }
        const RowCoef* rc = coef + row;
        float2* p = data + (int64_t)row * pitch;
        float2 v[E];
        if constexpr (kPrefetch) {
            tma::mbar_wait(mb, it & 1);
#pragma unroll
            for (int s = 0; s < E; ++s) v[s] = pf[t + NT * s];
            bar();   // the whole row is in registers: the prefetch buffer may be refilled
            if (t == 0 && row + row_stride < n_rows) {
                tma::mbar_arrive_expect_tx(mb, N * sizeof(float2));
                tma::bulk_load_1d(pf, data + (int64_t)(row + row_stride) * pitch, N * sizeof(float2), mb);
            }
        } else {
#pragma unroll
            for (int s = 0; s < E; ++s) v[s] = p[t + NT * s];
        }
        if constexpr (PHI13) {
            PhaseStepper ps;
            ps.init(__ldg(&rc->a1), __ldg(&rc->b1), __ldg(&rc->c1), (uint32_t)t, NT);
#pragma unroll
            for (int s = 0; s < E; ++s) v[s] = PK ? cmul_pk(v[s], ps.next()) : cmul(v[s], ps.next());
        }
        transform<P, false, 1, PAD, Bar, PK>(v, t, sm, tw, bar);
        {
            const uint64_t a2 = __ldg(&rc->a2), b2 = __ldg(&rc->b2);
            PhaseStepper ps;
            ps.init(a2, b2, 0ull, (uint32_t)t, NT);
#pragma unroll
            for (int s = 0; s < E / 2; ++s) v[s] = PK ? cmul_pk(v[s], ps.next()) : cmul(v[s], ps.next());
            ps.init_mirror(a2, b2, __ldg(&rc->bn2), (uint32_t)(N / 2 - t), (uint32_t)(N / 2 + t), NT);
#pragma unroll
            for (int s = E / 2; s < E; ++s) v[s] = PK ? cmul_pk(v[s], ps.next()) : cmul(v[s], ps.next());
        }
        bar();
        transform<P, true, 1, PAD, Bar, PK>(v, t, sm, tw, bar);
        if constexpr (PHI13) {
            PhaseStepper ps;
            ps.init(__ldg(&rc->a3), __ldg(&rc->b3), __ldg(&rc->c3), (uint32_t)t, NT);
#pragma unroll
            for (int s = 0; s < E; ++s) {
                const float2 x = PK ? cmul_pk(v[s], ps.next()) : cmul(v[s], ps.next());
                if (live) p[t + NT * s] = x;
            }
        } else {
#pragma unroll
            for (int s = 0; s < E; ++s)
                if (live) p[t + NT * s] = v[s];
        }
        bar();
    }
}

// The same row pipeline with ONE copy of the transform in the instruction stream: the inverse transform is run as
// conj(FFT(conj(.))) (equal to the conjugated-twiddle form up to rounding -- the fused multiply-adds of the inter-pass
// twiddles round the other product first: 1e-7 relative, csrc/hosttest/test_fft_host.cu), so the loop
// body executes twice per row and the conjugations ride on the phase multiplies.  The straight-line form above is
// 4096 SASS instructions (64 KB) at 8192 samples -- twice the 32 KB instruction cache, ncu: 5 % of the stalls
// `no_instruction`; this one is a little more than half of that.  One row per CTA, next row prefetched by TMA.
template <class P, int PAD, int PF, int MINB>
__global__ void __launch_bounds__(P::NT, MINB) k_range_rolled(float2* __restrict__ data, int64_t pitch, int n_rows,
                                                           const RowCoef* __restrict__ coef,
                                                           const float2* __restrict__ tw) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full;
    constexpr int E = P::E, NT = P::NT, N = P::N;
    constexpr int SMROW = N + (PAD ? (N >> PAD) : 0);
    const int t = threadIdx.x;
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    // PF = 1: the next row is prefetched into a buffer of its own while this one is transformed; PF = 2: into the exchange
    // buffer itself, as soon as the last gather of the row has left it (the copy then overlaps the last pass, Phi3 and the
    // stores) -- no second buffer, so two CTAs fit an SM; PF = 0: plain loads
    float2* pf = PF == 1 ? sm + SMROW : sm;
    CtaBarrier bar;
    const int row_stride = gridDim.x;
    if constexpr (PF != 0) {
        if (t == 0) {
            tma::mbar_init(&full, 1);
            tma::fence_barrier_init();
            if ((int)blockIdx.x < n_rows) {
                tma::mbar_arrive_expect_tx(&full, N * sizeof(float2));
                tma::bulk_load_1d(pf, data + (int64_t)blockIdx.x * pitch, N * sizeof(float2), &full);
            }
        }
        __syncthreads();
    }
    int it = 0;
    for (int row = blockIdx.x; row < n_rows; row += row_stride, ++it) {
        const RowCoef* rc = coef + row;
        float2* p = data + (int64_t)row * pitch;
        float2 v[E];
        if constexpr (PF != 0) {
            tma::mbar_wait(&full, it & 1);
#pragma unroll
            for (int s = 0; s < E; ++s) v[s] = pf[t + NT * s];
            bar();   // the whole row is in registers: the prefetch buffer may be refilled
            if (PF == 1 && t == 0 && row + row_stride < n_rows) {
                tma::mbar_arrive_expect_tx(&full, N * sizeof(float2));
                tma::bulk_load_1d(pf, data + (int64_t)(row + row_stride) * pitch, N * sizeof(float2), &full);
            }
        } else {
#pragma unroll
            for (int s = 0; s < E; ++s) v[s] = __ldcs(p + t + NT * s);
        }
        {
            PhaseStepper ps;
            ps.init(__ldg(&rc->a1), __ldg(&rc->b1), __ldg(&rc->c1), (uint32_t)t, NT);
#pragma unroll
            for (int s = 0; s < E; ++s) v[s] = cmul(v[s], ps.next());
        }
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            transform<P, false, 1, PAD, CtaBarrier, false>(v, t, sm, tw, bar);
            if (half == 0) {   // v <- conj(v Phi2)
                const uint64_t a2 = __ldg(&rc->a2), b2 = __ldg(&rc->b2);
                PhaseStepper ps;
                ps.init(a2, b2, 0ull, (uint32_t)t, NT);
#pragma unroll
                for (int s = 0; s < E / 2; ++s) {
                    const float2 w = ps.next(), x = v[s];
                    v[s] = make_float2(fmaf(x.x, w.x, -x.y * w.y), fmaf(-x.x, w.y, -x.y * w.x));
                }
                ps.init_mirror(a2, b2, __ldg(&rc->bn2), (uint32_t)(N / 2 - t), (uint32_t)(N / 2 + t), NT);
#pragma unroll
                for (int s = E / 2; s < E; ++s) {
                    const float2 w = ps.next(), x = v[s];
                    v[s] = make_float2(fmaf(x.x, w.x, -x.y * w.y), fmaf(-x.x, w.y, -x.y * w.x));
                }
            }
            bar();
        }
        if constexpr (PF == 2) {
            if (t == 0 && row + row_stride < n_rows) {
                tma::fence_proxy_async();   // the generic-proxy reads of the buffer (ordered by the barrier) precede the refill
                tma::mbar_arrive_expect_tx(&full, N * sizeof(float2));
                tma::bulk_load_1d(pf, data + (int64_t)(row + row_stride) * pitch, N * sizeof(float2), &full);
            }
        }
        {   // conj(v) Phi3
            PhaseStepper ps;
            ps.init(__ldg(&rc->a3), __ldg(&rc->b3), __ldg(&rc->c3), (uint32_t)t, NT);
#pragma unroll
            for (int s = 0; s < E; ++s) p[t + NT * s] = cmul_conj(ps.next(), v[s]);
        }
    }
}

// ------------------------------------------------------------------------------ host: plan
}  // namespace

namespace {

template <int A1>
int launch_outer_fwd(nis_csa_plan* pl, const float2* in, int64_t in_pitch, int col0, int ncols, cudaStream_t st) {
    dim3 grid((ncols + 31) / 32, (pl->A2 + 7) / 8);
    k_az_outer_fwd<A1><<<grid, 256, 0, st>>>(in + col0, in_pitch, pl->work + col0, pl->n_rg, ncols, pl->A2, pl->n_az,
                                              pl->tw_full);
    NIS_LAUNCH_CHECK(pl->ctx);
    return NIS_OK;
}

template <int A1, int TN, int TA>
int launch_outer_inv(nis_csa_plan* pl, float2* slc, double* max_sq, int col0, int ncols, cudaStream_t st) {
    const size_t smem = (size_t)A1 * TN * (TA + 1) * sizeof(float2);
    static bool attr_done_dev[64] = {};
    bool& attr_done = attr_done_dev[nis::current_device() & 63];
    if (!attr_done) {
        NIS_CUDA_TRY(cudaFuncSetAttribute(k_az_outer_inv<A1, TN, TA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
        attr_done = true;
    }
    dim3 grid(ncols / TN, pl->A2 / TA);
    const float scale = (float)(1.0 / ((double)pl->n_az * (double)pl->n_rg));
    k_az_outer_inv<A1, TN, TA><<<grid, 256, smem, st>>>(pl->work + col0, pl->n_rg, slc + (int64_t)col0 * pl->n_az,
                                                        pl->n_rg, pl->A2, pl->n_az, scale, max_sq, pl->tw_full);
    NIS_LAUNCH_CHECK(pl->ctx);
    return NIS_OK;
}

template <class P, int W, bool INV, bool PH>
int launch_inner_one(nis_csa_plan* pl, int col0, int ncols, int k10, int nk1, cudaStream_t st) {
    auto kern = k_az_inner_tma<P, INV, W, PH>;
    const size_t smem = 2 * (size_t)P::N * W * sizeof(float2);   // double-buffered tile
    static int per_sm_dev[64] = {};
    int& per_sm = per_sm_dev[nis::current_device() & 63];
    if (!per_sm) {
        NIS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int n = 1;
        NIS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, P::NT * W, smem));
        per_sm = n < 1 ? 1 : n;
    }
    const int n_col_tiles = ncols / W, n_tiles = n_col_tiles * nk1;
    int grid = pl->ctx->num_sms * per_sm;
    if (grid > n_tiles) grid = n_tiles;
    kern<<<grid, P::NT * W, smem, st>>>(pl->tile_map, pl->work, pl->n_rg, n_col_tiles, n_tiles, col0 / W, k10, pl->tw_inner,
                                        INV ? pl->ph3_ab : pl->ph1_ab, INV ? pl->ph3_c : pl->ph1_c);
    NIS_LAUNCH_CHECK(pl->ctx);
    return NIS_OK;
}
template <class P, int W>
int launch_inner(nis_csa_plan* pl, bool inv, int col0, int ncols, int k10, int nk1, cudaStream_t st) {
    if (pl->phase_in_az)
        return inv ? launch_inner_one<P, W, true, true>(pl, col0, ncols, k10, nk1, st)
                   : launch_inner_one<P, W, false, true>(pl, col0, ncols, k10, nk1, st);
    return inv ? launch_inner_one<P, W, true, false>(pl, col0, ncols, k10, nk1, st)
               : launch_inner_one<P, W, false, false>(pl, col0, ncols, k10, nk1, st);
}

// PK = false: the packed forms cut this kernel's instruction count by a third but not its time (it is bound by the
// FMA pipe and the MIO/shared-memory pipe back to back, not by issue slots; measured 0.466 vs 0.456 ms at 8192^2)
template <class P, int PAD, int RPB, int MINB, bool PK, bool PHI13>
int launch_range_one(nis_csa_plan* pl, int row0, int nrows, cudaStream_t st) {
    auto kern = k_range<P, PAD, RPB, MINB, PK, PHI13>;
    constexpr int SMROW = P::N + (PAD ? (P::N >> PAD) : 0);
    constexpr int GROUP_ELEMS = SMROW + ((P::NT >= 32 && P::N <= 8192) ? P::N : 0);   // exchange buffer + prefetch buffer
    const size_t smem = (size_t)GROUP_ELEMS * RPB * sizeof(float2);
    static int per_sm_dev[64] = {};
    int& per_sm = per_sm_dev[nis::current_device() & 63];
    if (!per_sm) {
        NIS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int n = 1;
        NIS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, P::NT * RPB, smem));
        per_sm = n < 1 ? 1 : n;
    }
    const int blocks_needed = (nrows + RPB - 1) / RPB;
    int grid = pl->ctx->num_sms * per_sm;
    if (grid > blocks_needed) grid = blocks_needed;
    kern<<<grid, dim3(P::NT, RPB), smem, st>>>(pl->work + (int64_t)row0 * pl->n_rg, pl->n_rg, nrows, pl->coef + row0, pl->tw_rg);
    NIS_LAUNCH_CHECK(pl->ctx);
    return NIS_OK;
}
template <class P, int PAD, int RPB, int MINB, bool PK = false>
int launch_range(nis_csa_plan* pl, int row0, int nrows, cudaStream_t st) {
    if (pl->phase_in_az) return launch_range_one<P, PAD, RPB, MINB, PK, false>(pl, row0, nrows, st);
    return launch_range_one<P, PAD, RPB, MINB, PK, true>(pl, row0, nrows, st);
}

template <class P, int PAD, int PF, int MINB>
int launch_range_rolled(nis_csa_plan* pl, int row0, int nrows, cudaStream_t st) {
    auto kern = k_range_rolled<P, PAD, PF, MINB>;
    constexpr int SMROW = P::N + (PAD ? (P::N >> PAD) : 0);
    const size_t smem = (size_t)(SMROW + (PF == 1 ? P::N : 0)) * sizeof(float2);
    static bool attr_done_dev[64] = {};
    bool& attr_done = attr_done_dev[nis::current_device() & 63];
    if (!attr_done) {
        NIS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    int grid = pl->ctx->num_sms * MINB;
    if (grid > nrows) grid = nrows;
    kern<<<grid, P::NT, smem, st>>>(pl->work + (int64_t)row0 * pl->n_rg, pl->n_rg, nrows, pl->coef + row0, pl->tw_rg);
    NIS_LAUNCH_CHECK(pl->ctx);
    return NIS_OK;
}

// plans: <N, E, R0, R1, R2>
using P16 = Plan<16, 16, 16, 1, 1>;
using P64 = Plan<64, 8, 8, 8, 1>;
using P128 = Plan<128, 16, 16, 8, 1>;
using P256 = Plan<256, 16, 16, 16, 1>;
using P512 = Plan<512, 16, 8, 8, 8>;
using P1024 = Plan<1024, 16, 16, 8, 8>;
using P2048 = Plan<2048, 16, 16, 16, 8>;
using P4096 = Plan<4096, 16, 16, 16, 16>;
using P8192 = Plan<8192, 16, 16, 8, 8, 8>;
using P16384 = Plan<16384, 16, 16, 16, 8, 8>;
// 8192 samples, 32 per thread: three passes (two shared-memory exchanges per transform instead of three) on 256 threads that
// may use the whole register file (one CTA per SM either way) -- the default at 8192 samples
using P8192E32 = Plan<8192, 32, 32, 16, 16>;
using P512E32 = Plan<512, 32, 32, 16, 1>;
using P4096E32 = Plan<4096, 32, 32, 16, 8>;
using P16384E32 = Plan<16384, 32, 32, 32, 16>;
using P1024E32 = Plan<1024, 32, 32, 32, 1>;

template <class P>
int upload_twiddles(float2** dev) {
    std::vector<float2> h(P::tw_len + 1);
    build_twiddles<P>(h.data());
    NIS_CUDA_TRY(cudaMalloc(dev, h.size() * sizeof(float2)));
    NIS_CUDA_TRY(cudaMemcpy(*dev, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
    return NIS_OK;
}


// cluster azimuth transforms: forward (raw -> W, natural Doppler row order) and inverse (W -> slc, corner-turned)
template <class P, int C, int W, bool INV, bool PH>
int launch_az_cluster_ph(nis_csa_plan* pl, const CUtensorMap& map, float2* out, int64_t out_pitch, double* max_sq,
                         cudaStream_t st);
template <class P, int C, int W, bool INV>
int launch_az_cluster(nis_csa_plan* pl, const CUtensorMap& map, float2* out, int64_t out_pitch, double* max_sq,
                      cudaStream_t st) {
    if (!pl->phase_in_az) return launch_az_cluster_ph<P, C, W, INV, false>(pl, map, out, out_pitch, max_sq, st);
    return launch_az_cluster_ph<P, C, W, INV, true>(pl, map, out, out_pitch, max_sq, st);
}
template <class P, int C, int W, bool INV, bool PH>
int launch_az_cluster_ph(nis_csa_plan* pl, const CUtensorMap& map, float2* out, int64_t out_pitch, double* max_sq,
                         cudaStream_t st) {
    auto kern = k_az_cluster<P, C, W, INV, INV, PH>;
    const size_t smem = INV ? (size_t)W * (P::N + 2) * sizeof(float2) : (size_t)P::N * W * sizeof(float2);
    static int n_clusters_dev[64] = {};
    int& n_clusters = n_clusters_dev[nis::current_device() & 63];
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(P::NT * W);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (!n_clusters) {
        NIS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (C > 8) NIS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cfg.gridDim = dim3(C * pl->ctx->num_sms);
        int n = 0;
        NIS_CUDA_TRY(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
        if (n < 1) {
            set_error("azimuth cluster kernel: no cluster of %d CTAs x %zu bytes fits this device", C, smem);
            return NIS_ERR_UNSUPPORTED;
        }
        if (const char* v = getenv("NIS_AZ_CLUSTERS")) { const int lim = atoi(v); if (lim > 0 && lim < n) n = lim; }
        n_clusters = n;
        if (getenv("NIS_DEBUG")) fprintf(stderr, "[nis] az cluster C=%d M=%d W=%d inv=%d: %d clusters resident, %zu B smem\n", C, P::N, W, (int)INV, n, smem);
    }
    const int n_col_tiles = pl->n_rg / W;
    const int nc = n_clusters < n_col_tiles ? n_clusters : n_col_tiles;
    cfg.gridDim = dim3(C * nc);
    const float scale = (float)(1.0 / ((double)pl->n_az * (double)pl->n_rg));
    NIS_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, map, out, out_pitch, n_col_tiles, scale, max_sq,
                                    (const float2*)pl->tw_inner, (const float2*)pl->tw_full,
                                    (const uint4*)(INV ? pl->ph3_ab : pl->ph1_ab), (const uint32_t*)(INV ? pl->ph3_c : pl->ph1_c)));
    NIS_LAUNCH_CHECK(pl->ctx);
    return NIS_OK;
}
template <class P, int C, int W>
int az_cluster_fwd(nis_csa_plan* pl, const float2* in, int64_t pitch, cudaStream_t st) {
    if (pl->in_map_base != in || pl->in_map_pitch != pitch) {
        int rc = tma::make_tile_map_3d(&pl->in_map, in, pl->n_rg, C, P::N, pitch, P::N < 256 ? P::N : 256, W);
        if (rc != NIS_OK) return rc;
        pl->in_map_base = in;
        pl->in_map_pitch = pitch;
    }
    return launch_az_cluster<P, C, W, false>(pl, pl->in_map, pl->work, pl->n_rg, nullptr, st);
}
template <class P, int C, int W>
int az_cluster_inv(nis_csa_plan* pl, float2* slc, double* max_sq, cudaStream_t st) {
    return launch_az_cluster<P, C, W, true>(pl, pl->work_map3, slc, pl->n_az, max_sq, st);
}
template <class P, int C, int W>
int az_cluster_setup(nis_csa_plan* pl) {
    pl->az_cluster = true;
    pl->A1 = 1;            // rows of W hold Doppler bins in natural order
    pl->A2 = pl->n_az;
    pl->az_fwd = az_cluster_fwd<P, C, W>;
    pl->az_inv = az_cluster_inv<P, C, W>;
    int rc = upload_twiddles<P>(&pl->tw_inner);
    if (rc != NIS_OK) return rc;
    return tma::make_tile_map_3d(&pl->work_map3, pl->work, pl->n_rg, C, P::N, pl->n_rg, P::N < 256 ? P::N : 256, W);
}

struct AzSplit { int n, a1, a2; };
const AzSplit kAzSplits[] = {{64, 4, 16},     {128, 8, 16},    {256, 16, 16},   {512, 8, 64},    {1024, 16, 64},
                             {2048, 8, 256},  {4096, 16, 256}, {8192, 16, 512}, {16384, 16, 1024},
                             {32768, 32, 1024}};   // sar_vehicle_sim.py focuses 32768 pulses (:43)

struct AzClusterCfg { int n_az, id; int (*setup)(nis_csa_plan*); };
const AzClusterCfg kAzCluster[] = {
    {16384, 1, az_cluster_setup<P1024, 16, 8>},
    {8192, 1, az_cluster_setup<P1024, 8, 8>}, {8192, 2, az_cluster_setup<P512, 16, 8>},
    {8192, 3, az_cluster_setup<P512, 16, 16>},
    {4096, 1, az_cluster_setup<P512, 8, 8>},  {4096, 2, az_cluster_setup<P1024, 4, 8>},
    {4096, 3, az_cluster_setup<P256, 16, 8>}, {4096, 4, az_cluster_setup<P256, 16, 16>},
    {4096, 5, az_cluster_setup<P512E32, 8, 16>}, {4096, 6, az_cluster_setup<P512E32, 8, 8>},
    {8192, 5, az_cluster_setup<P1024E32, 8, 8>}, {8192, 6, az_cluster_setup<P512E32, 16, 16>},
    {8192, 7, az_cluster_setup<P512E32, 16, 8>},
    {2048, 1, az_cluster_setup<P256, 8, 8>},  {2048, 2, az_cluster_setup<P512, 4, 8>},
    {1024, 1, az_cluster_setup<P256, 4, 8>},
};

bool range_supported(int n) { return n >= 64 && n <= 16384 && (n & (n - 1)) == 0; }
bool az_supported(int n) {
    for (const auto& s : kAzSplits)
        if (s.n == n) return true;
    return false;
}


template <int A1>
int launch_outer_inv_mag(nis_csa_plan* pl, float* mag, float scale, cudaStream_t st) {
    dim3 grid((pl->n_rg + 31) / 32, (pl->A2 + 7) / 8);
    k_az_outer_inv_mag<A1><<<grid, 256, 0, st>>>(pl->work, pl->n_rg, mag, pl->n_rg, pl->n_rg, pl->A2, pl->n_az, scale,
                                                 pl->tw_full);
    NIS_LAUNCH_CHECK(pl->ctx);
    return NIS_OK;
}

// two-kernel four-step azimuth engine on pl->work: launchers, inner-transform twiddles, TMA tile map
int setup_az_four_step(nis_csa_plan* pl) {
    const int n_az = pl->n_az, n_rg = pl->n_rg;
    int rc;
#define TRY_RC(x) do { rc = (x); if (rc != NIS_OK) return rc; } while (0)
    switch (pl->A1) {
        case 4: pl->outer_fwd = launch_outer_fwd<4>; pl->outer_inv = launch_outer_inv<4, 16, 16>;
                pl->outer_inv_mag = launch_outer_inv_mag<4>; break;
        case 8: pl->outer_fwd = launch_outer_fwd<8>; pl->outer_inv = launch_outer_inv<8, 16, 16>;
                pl->outer_inv_mag = launch_outer_inv_mag<8>; break;
        case 32: pl->outer_fwd = launch_outer_fwd<32>; pl->outer_inv = launch_outer_inv<32, 16, 16>;
                 pl->outer_inv_mag = launch_outer_inv_mag<32>; break;
        default: pl->outer_fwd = launch_outer_fwd<16>; pl->outer_inv = launch_outer_inv<16, 16, 16>;
                 pl->outer_inv_mag = launch_outer_inv_mag<16>;
                 if (const char* v = getenv("NIS_OUTER_INV")) {   // tuning knob (development only)
                     if (pl->A2 >= 32 && n_rg >= 32) {
                         if (v[0] == '1') pl->outer_inv = launch_outer_inv<16, 32, 8>;
                         if (v[0] == '2') pl->outer_inv = launch_outer_inv<16, 32, 16>;
                         if (v[0] == '3') pl->outer_inv = launch_outer_inv<16, 16, 32>;
                         if (v[0] == '4') pl->outer_inv = launch_outer_inv<16, 8, 32>;
                     }
                 }
                 break;
    }
    switch (pl->A2) {
        case 16: pl->inner_w = 32; pl->inner = launch_inner<P16, 32>; TRY_RC(upload_twiddles<P16>(&pl->tw_inner)); break;
        case 64: pl->inner_w = 32; pl->inner = launch_inner<P64, 32>; TRY_RC(upload_twiddles<P64>(&pl->tw_inner)); break;
        case 256: pl->inner_w = 16; pl->inner = launch_inner<P256, 16>; TRY_RC(upload_twiddles<P256>(&pl->tw_inner)); break;
        // 512-point tiles: 16 columns (128-byte row pieces, one 512-thread CTA per SM) measured 12 % faster than 8 columns
        // (three 256-thread CTAs per SM): 0.194 vs 0.220 ms per transform at 8192^2; 4 columns: 0.245 ms
        // 512- and 1024-point tiles: 32 samples per thread, TWO passes (32 x 16, 32 x 32) = one exchange through the tile
        // instead of two -- three shared-memory accesses per sample instead of five, 256 threads.  Measured at 8192^2:
        // 0.197 -> 0.184 ms per stage (5.8 TB/s).  NIS_AZ_INNER_PLAN=e16 restores the three-pass plans (development knob).
        // (Single-buffered 64 KB tiles with THREE such CTAs per SM, the next tile pulled into the same buffer after the last
        // gather: 0.183 / 0.185 ms -- no change, occupancy is not what limits this stage; not kept.)
        case 512:
            pl->inner_w = 16;
            if (const char* v = getenv("NIS_AZ_INNER_PLAN"); v && v[0] == 'e' && v[1] == '1') {
                pl->inner = launch_inner<P512, 16>;
                TRY_RC(upload_twiddles<P512>(&pl->tw_inner));
            } else {
                pl->inner = launch_inner<P512E32, 16>;
                TRY_RC(upload_twiddles<P512E32>(&pl->tw_inner));
            }
            break;
        default:
            pl->inner_w = 8;
            if (const char* v = getenv("NIS_AZ_INNER_PLAN"); v && v[0] == 'e' && v[1] == '1') {
                pl->inner = launch_inner<P1024, 8>;
                TRY_RC(upload_twiddles<P1024>(&pl->tw_inner));
            } else {
                pl->inner = launch_inner<P1024E32, 8>;
                TRY_RC(upload_twiddles<P1024E32>(&pl->tw_inner));
            }
            break;
    }
#undef TRY_RC
    if (n_rg % pl->inner_w) {
        set_error("azimuth engine: the range length %d must be a multiple of %d", n_rg, pl->inner_w);
        return NIS_ERR_UNSUPPORTED;
    }
    return tma::make_tile_map(&pl->tile_map, pl->work, n_az, n_rg, n_rg, pl->A2 < 256 ? pl->A2 : 256, pl->inner_w);
}

int upload_full_twiddles(nis_csa_plan* pl) {
    const int n_az = pl->n_az;
    std::vector<float2> h(n_az);
    const double two_pi = 6.283185307179586476925286766559;
    for (int m = 0; m < n_az; ++m) {
        double a = -two_pi * (double)m / (double)n_az;
        h[m] = make_float2((float)cos(a), (float)sin(a));
    }
    if (cudaMalloc(&pl->tw_full, n_az * sizeof(float2)) != cudaSuccess) return NIS_ERR_NOMEM;
    NIS_CUDA_TRY(cudaMemcpy(pl->tw_full, h.data(), n_az * sizeof(float2), cudaMemcpyHostToDevice));
    return NIS_OK;
}

}  // namespace

// ---- shared builders (also used by csa_generic.cu)
namespace nis {
namespace csa {

void build_axes(int n_az, int n_rg, const nis_csa_params& prm, std::vector<double>& range_axis,
                std::vector<double>& cross_range) {
    const double dt = 1.0 / prm.fs;
    range_axis.resize(n_rg);
    for (int n = 0; n < n_rg; ++n) range_axis[n] = prm.c * (prm.t_start + n * dt) / 2.0;  // :219, :346
    cross_range.resize(n_az);
    // t_slow = arange(N)/prf; t_slow -= mean(t_slow)  (:392-394); numpy's pairwise mean is reproduced to
    // the last few ulps by an extended-precision sum
    long double acc = 0;
    for (int i = 0; i < n_az; ++i) acc += (long double)((double)i / prm.prf);
    const double mean = (double)(acc / n_az);
    for (int i = 0; i < n_az; ++i) cross_range[i] = ((double)i / prm.prf - mean) * prm.vr;
}

std::vector<RowCoef> build_row_coefs(int n_az, int n_rg, const nis_csa_params& prm, int A1, int A2) {
    const double c = prm.c, lam = prm.lambda, Kr = prm.kr, Vr = prm.vr, R_ref = prm.r_ref;
    const double dt = 1.0 / prm.fs;
    std::vector<RowCoef> h(n_az);
    const long double dtl = dt, t0 = prm.t_start;
    const long double df = (long double)prm.fs / n_rg;
    const double fa_unit = 1.0 / ((double)n_az * (1.0 / prm.prf));  // np.fft.fftfreq(n, d): k * (1/(n d))
    for (int rho = 0; rho < n_az; ++rho) {
        const int k1 = rho / A2, k2 = rho % A2;
        const int kk = k1 + A1 * k2;                            // Doppler bin held by this row
        const int ks = (kk < (n_az + 1) / 2) ? kk : kk - n_az;  // fftfreq sign convention
        const double fa = (double)ks * fa_unit;
        double arg = 1.0 - (lam * fa / (2.0 * Vr)) * (lam * fa / (2.0 * Vr));
        if (arg < 0) arg = 1e-9;                                 // :246
        const double D = sqrt(arg);
        const double Cs = (1.0 / D) - 1.0;
        const double tau_ref = 2.0 * R_ref / (c * D);
        const long double cs = Cs, d1 = t0 - (long double)tau_ref;
        RowCoef r{};
        r.a1 = to_fix(-(Kr * cs / 2) * dtl * dtl);
        r.b1 = to_fix(-(Kr * cs) * dtl * d1);
        r.c1 = to_fix(-(Kr * cs / 2) * d1 * d1);
        r.a2 = to_fix(df * df / (2.0L * Kr * (1.0L + cs)));
        r.b2 = to_fix(2.0L * R_ref * cs * df / c);
        r.bn2 = r.b2 * (uint64_t)n_rg;
        const long double e3 = t0 - 2.0L * R_ref / c;
        const long double q3 = (long double)Kr * cs * (1.0L + cs);
        r.a3 = to_fix(-(q3 / 2) * dtl * dtl);
        r.b3 = to_fix((long double)c * D * dtl / lam - q3 * dtl * e3);
        r.c3 = to_fix((long double)c * D * t0 / lam - (q3 / 2) * e3 * e3);
        h[rho] = r;
    }
    return h;
}

}  // namespace csa
}  // namespace nis

namespace nis {
namespace csa {

bool az_engine_supported(int n_az, int n_rg) {
    for (const auto& s : kAzSplits)
        if (s.n == n_az) {
            const int w = (s.a2 <= 64) ? 32 : (s.a2 <= 512 ? 16 : 8);   // tile width of the inner transform (setup_az_four_step)
            return n_rg % w == 0;
        }
    return false;
}

// A plan that owns only a workspace and the four-step azimuth engine on it (used by the RDA path, rda.cu)
int az_engine_create(nis_ctx* ctx, int n_az, int n_rg, nis_csa_plan** out) {
    nis_csa_plan* pl = new nis_csa_plan();
    pl->ctx = ctx;
    pl->n_az = n_az;
    pl->n_rg = n_rg;
    pl->size_class = 1;
    for (const auto& s : kAzSplits)
        if (s.n == n_az) { pl->A1 = s.a1; pl->A2 = s.a2; }
    int rc = NIS_OK;
    if (cudaMalloc(&pl->work, (size_t)n_az * n_rg * sizeof(float2)) != cudaSuccess) {
        set_error("az_engine_create: cannot allocate %zu-byte workspace", (size_t)n_az * n_rg * sizeof(float2));
        rc = NIS_ERR_NOMEM;
    }
    if (rc == NIS_OK) rc = setup_az_four_step(pl);
    if (rc == NIS_OK) rc = upload_full_twiddles(pl);
    if (rc != NIS_OK) { nis_csa_plan_destroy(pl); return rc; }
    *out = pl;
    return NIS_OK;
}

}  // namespace csa
}  // namespace nis

extern "C" int nis_csa_size_class(int32_t n_az, int32_t n_rg) {
    if (az_supported(n_az) && range_supported(n_rg)) return 1;
    return generic_supported(n_az, n_rg) ? 2 : 0;
}

extern "C" int nis_csa_plan_destroy(nis_csa_plan* pl) {
    if (!pl) return NIS_OK;
    generic_destroy(pl);
    cudaFree(pl->work);
    cudaFree(pl->tw_inner);
    cudaFree(pl->tw_full);
    cudaFree(pl->tw_rg);
    cudaFree(pl->coef);
    cudaFree(pl->ph1_ab);
    cudaFree(pl->ph1_c);
    cudaFree(pl->ph3_ab);
    cudaFree(pl->ph3_c);
    for (auto& set : pl->prof_ev)
        for (auto& e : set)
            if (e) cudaEventDestroy(e);
    delete pl;
    return NIS_OK;
}

extern "C" int nis_csa_plan_create(nis_ctx* ctx, int32_t n_az, int32_t n_rg, const nis_csa_params* prm,
                                   nis_csa_plan** out) {
    NIS_REQUIRE(ctx && prm && out, "nis_csa_plan_create: null argument");
    const int cls = nis_csa_size_class(n_az, n_rg);
    if (!cls) {
        set_error("nis_csa_plan_create: size %d x %d not supported (azimuth: power of two 64..32768, range: power of two 64..16384; or per axis any length "
                  "<= 14000 whose prime factors are <= 13, or any length <= 8192)", n_az, n_rg);
        return NIS_ERR_UNSUPPORTED;
    }
    NIS_REQUIRE(prm->fs > 0 && prm->prf > 0 && prm->vr > 0 && prm->kr != 0 && prm->lambda > 0 && prm->c > 0,
                "nis_csa_plan_create: non-physical parameters");
    DeviceGuard device_guard(ctx->device);   // the caller's current device is restored on return
    nis_csa_plan* pl = new nis_csa_plan();
    pl->ctx = ctx;
    pl->n_az = n_az;
    pl->n_rg = n_rg;
    pl->prm = *prm;
    pl->size_class = cls;
    int rc = NIS_OK;
#define FAIL_IF(x) do { rc = (x); if (rc != NIS_OK) { nis_csa_plan_destroy(pl); return rc; } } while (0)
    build_axes(n_az, n_rg, *prm, pl->range_axis, pl->cross_range);
    if (cudaMalloc(&pl->work, (size_t)n_az * n_rg * sizeof(float2)) != cudaSuccess) {
        set_error("nis_csa_plan_create: cannot allocate %zu-byte workspace", (size_t)n_az * n_rg * sizeof(float2));
        nis_csa_plan_destroy(pl);
        return NIS_ERR_NOMEM;
    }
    if (cls == 2) {
        pl->A1 = 1;
        pl->A2 = n_az;
    } else {
        for (const auto& s : kAzSplits)
            if (s.n == n_az) { pl->A1 = s.a1; pl->A2 = s.a2; }
        // (8192 = 32 x 256 instead of 16 x 512 measured slower: 1.216 vs 1.096 ms per frame, the radix-32 corner-turning
        // outer stage takes 0.271 ms)
        // default: whole-column cluster transforms where they measured faster than the two-kernel four-step
        // (n_az = 4096: 0.187 vs 0.204 ms per frame on a B200); NIS_CSA_AZ = 0 / k overrides (development knob).
        // Configuration 6 (round 2): 512-point per-CTA transforms on the TWO-pass plan 32 x 16 (32 samples per thread, one
        // exchange through the tile instead of two, 128-thread CTAs): 0.0856 / 0.0863 ms per transform at 4096^2 against
        // 0.0936 / 0.0941 for configuration 1 (three passes, 256 threads); 16-column tiles (5): 0.097 / 0.103.  At 8192 the
        // same plans (5: 8 x 1024, 6: 16 x 512) reach 0.38-0.39 / 0.45-0.47 ms per transform: the four-step engine (0.354 /
        // 0.384 for its two kernels) stays the default there.
        int az_id = (n_az == 4096) ? 6 : 0;
        if (const char* v = getenv("NIS_CSA_AZ")) az_id = atoi(v);
        for (const auto& c : kAzCluster)
            if (c.n_az == n_az && c.id == az_id && n_rg % 16 == 0) FAIL_IF(c.setup(pl));
    }
    {
        std::vector<RowCoef> h = build_row_coefs(n_az, n_rg, *prm, pl->A1, pl->A2);
        FAIL_IF(cudaMalloc(&pl->coef, n_az * sizeof(RowCoef)) == cudaSuccess ? NIS_OK : NIS_ERR_NOMEM);
        if (cudaMemcpy(pl->coef, h.data(), n_az * sizeof(RowCoef), cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("nis_csa_plan_create: upload of the phase coefficients failed: %s", cudaGetErrorString(cudaGetLastError()));
            nis_csa_plan_destroy(pl);
            return NIS_ERR_CUDA;
        }
    }
    if (cls == 2) {
        FAIL_IF(generic_create(pl));
        *out = pl;
        return NIS_OK;
    }
    {
        // Where Phi1 / Phi3 are applied.  Default: in k_range with Phi2 (one HBM round trip for all three).  NIS_CSA_PHASE=az
        // moves Phi1 onto the store of the forward azimuth kernel and Phi3 onto the load of the inverse one.  Measured on a
        // B200 (profiles/kbench_r2b_*.jsonl): k_range gains little (8192^2: 0.453 -> 0.429 ms; it is bound by its shared-memory
        // exchanges, not by the phase arithmetic) while the azimuth kernels, which run one 512-thread CTA per SM, cannot hide
        // the extra dependent work (0.196 -> 0.287 and 0.197 -> 0.243 ms): 1.20 -> 1.32 ms per frame, and 0.279 -> 0.291 ms at
        // 4096^2 with the cluster engine.  Kept as a tested development knob, not the default.
        const char* v = getenv("NIS_CSA_PHASE");
        pl->phase_in_az = (v && v[0] == 'a');
        if (pl->phase_in_az) {
            std::vector<RowCoef> h = build_row_coefs(n_az, n_rg, *prm, pl->A1, pl->A2);
            std::vector<uint4> ab1(n_az), ab3(n_az);
            std::vector<uint32_t> c1(n_az), c3(n_az);
            for (int r = 0; r < n_az; ++r) {
                ab1[r] = make_uint4((uint32_t)h[r].a1, (uint32_t)(h[r].a1 >> 32), (uint32_t)h[r].b1, (uint32_t)(h[r].b1 >> 32));
                ab3[r] = make_uint4((uint32_t)h[r].a3, (uint32_t)(h[r].a3 >> 32), (uint32_t)h[r].b3, (uint32_t)(h[r].b3 >> 32));
                c1[r] = (uint32_t)(h[r].c1 >> 32) + 0x100u;   // + half an ulp of the 23-bit sincos argument (cis_u32_pre)
                c3[r] = (uint32_t)(h[r].c3 >> 32) + 0x100u;
            }
            auto up = [&](void** dst, const void* src, size_t bytes) -> int {
                if (cudaMalloc(dst, bytes) != cudaSuccess) return NIS_ERR_NOMEM;
                return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess ? NIS_OK : NIS_ERR_CUDA;
            };
            FAIL_IF(up((void**)&pl->ph1_ab, ab1.data(), n_az * sizeof(uint4)));
            FAIL_IF(up((void**)&pl->ph1_c, c1.data(), n_az * sizeof(uint32_t)));
            FAIL_IF(up((void**)&pl->ph3_ab, ab3.data(), n_az * sizeof(uint4)));
            FAIL_IF(up((void**)&pl->ph3_c, c3.data(), n_az * sizeof(uint32_t)));
        }
    }

    // ---- kernel selection (power-of-two path)
    if (!pl->az_cluster) FAIL_IF(setup_az_four_step(pl));
    switch (n_rg) {
        case 64: pl->range = launch_range<P64, 3, 8, 1>; FAIL_IF(upload_twiddles<P64>(&pl->tw_rg)); break;
        case 128: pl->range = launch_range<P128, 0, 8, 1>; FAIL_IF(upload_twiddles<P128>(&pl->tw_rg)); break;
        case 256: pl->range = launch_range<P256, 4, 8, 1>; FAIL_IF(upload_twiddles<P256>(&pl->tw_rg)); break;
        case 512: pl->range = launch_range<P512, 3, 4, 2>; FAIL_IF(upload_twiddles<P512>(&pl->tw_rg)); break;
        case 1024: pl->range = launch_range<P1024, 4, 4, 2>; FAIL_IF(upload_twiddles<P1024>(&pl->tw_rg)); break;
        case 2048: pl->range = launch_range<P2048, 4, 1, 4>; FAIL_IF(upload_twiddles<P2048>(&pl->tw_rg)); break;
        // one row per CTA, several CTAs per SM: independent CTAs drift out of phase, so one CTA's shared-memory exchange
        // overlaps another's butterflies (measured: 0.096 vs 0.104 ms at 4096^2 against two row groups inside one CTA)
        case 4096:
            // the rolled form with 32 samples per thread (32 x 16 x 8, 128 threads at 128 registers, FOUR CTAs per SM, next row
            // prefetched into the exchange buffer): 0.0955 -> 0.0884 ms at 4096^2.  (The rolled form on the 16-sample plan had
            // gained nothing: 0.098 ms.)  NIS_RANGE_PLAN=e16 keeps the straight-line 16-sample kernel reachable.
            if (const char* v = getenv("NIS_RANGE_PLAN"); !(v && v[0] == 'e') && !pl->phase_in_az) {
                pl->range = launch_range_rolled<P4096E32, 5, 2, 4>;
                FAIL_IF(upload_twiddles<P4096E32>(&pl->tw_rg));
            } else {
                pl->range = launch_range<P4096, 4, 1, 2>;
                FAIL_IF(upload_twiddles<P4096>(&pl->tw_rg));
            }
            break;
        case 8192: {
            // 32 samples per thread, three passes 32 x 16 x 16.  Default (round 2): k_range_rolled -- ONE copy of the transform
            // in the instruction stream (inverse = conj FFT conj), which ptxas fits into 128 registers without spills, so TWO
            // 256-thread CTAs share an SM and drift out of phase (one CTA's exchange overlaps the other's butterflies); the
            // next row is prefetched by TMA into the exchange buffer itself once the row's last gather has left it.
            // Measured at 8192^2 (profiles/kbench_r2f.jsonl): 0.370 ms; without the prefetch 0.399; one CTA per SM with a
            // prefetch buffer of its own (153 registers) 0.468; packed fp32x2 butterflies (216 B of spills) 0.430.
            // NIS_RANGE_PLAN=e32flat: the straight-line kernel (213 registers, one CTA per SM, 4096 SASS instructions =
            // twice the instruction cache): 0.418 ms; NIS_RANGE_PLAN=e16: 16 samples per thread, four passes: 0.457 ms.
            // All three are parity-tested (tests/test_gpu_parity_r2.py).
            const char* v = getenv("NIS_RANGE_PLAN");
            if (v && v[0] == 'e' && v[1] == '1') {
                pl->range = launch_range<P8192, 4, 1, 1>;
                FAIL_IF(upload_twiddles<P8192>(&pl->tw_rg));
            } else {
                if (pl->phase_in_az || (v && strcmp(v, "e32flat") == 0)) pl->range = launch_range<P8192E32, 5, 1, 1>;
                else pl->range = launch_range_rolled<P8192E32, 5, 2, 2>;
                FAIL_IF(upload_twiddles<P8192E32>(&pl->tw_rg));
            }
            break;
        }
        default:   // 16384 samples: the rolled form on the three-pass plan 32 x 32 x 16 (512 threads, one CTA per SM)
            if (pl->phase_in_az || getenv("NIS_RANGE_PLAN")) {
                pl->range = launch_range<P16384, 4, 1, 1>;
                FAIL_IF(upload_twiddles<P16384>(&pl->tw_rg));
            } else {
                pl->range = launch_range_rolled<P16384E32, 5, 2, 1>;
                FAIL_IF(upload_twiddles<P16384E32>(&pl->tw_rg));
            }
            break;
    }
    FAIL_IF(upload_full_twiddles(pl));
#undef FAIL_IF
    *out = pl;
    return NIS_OK;
}

extern "C" int nis_csa_axes(const nis_csa_plan* pl, double* range_axis, double* cross_range) {
    NIS_REQUIRE(pl, "nis_csa_axes: null plan");
    if (range_axis) memcpy(range_axis, pl->range_axis.data(), pl->range_axis.size() * sizeof(double));
    if (cross_range) memcpy(cross_range, pl->cross_range.data(), pl->cross_range.size() * sizeof(double));
    return NIS_OK;
}

extern "C" int nis_csa_focus(nis_csa_plan* pl, const nis_c32* phist, int64_t pitch, nis_c32* slc, double* max_sq,
                             nis_stream stream) {
    NIS_REQUIRE(pl && phist && slc, "nis_csa_focus: null argument");
    NIS_REQUIRE(pitch >= pl->n_rg, "nis_csa_focus: pitch %lld < n_rg %d", (long long)pitch, pl->n_rg);
    cudaStream_t st = (cudaStream_t)stream;
    if (max_sq != nullptr) NIS_CUDA_TRY(cudaMemsetAsync(max_sq, 0, sizeof(double), st));   // the last kernel accumulates with max
    if (pl->size_class == 2)
        return generic_focus(pl, reinterpret_cast<const float2*>(phist), pitch, reinterpret_cast<float2*>(slc), max_sq, st);
    cudaEvent_t* ev = pl->profiling ? pl->prof_ev[pl->prof_calls % nis_csa_plan::kProfRing] : nullptr;
#define STAGE_MARK(i) do { if (ev) NIS_CUDA_TRY(cudaEventRecord(ev[i], st)); } while (0)
    int rc;
    const int n_az = pl->n_az, n_rg = pl->n_rg, A1 = pl->A1;
    const float2* in = reinterpret_cast<const float2*>(phist);
    float2* out = reinterpret_cast<float2*>(slc);
#define RUN(x) do { if ((rc = (x)) != NIS_OK) return rc; } while (0)
    if (pl->az_cluster) {
        STAGE_MARK(0);
        RUN(pl->az_fwd(pl, in, pitch, st));
        STAGE_MARK(1);
        STAGE_MARK(2);
        RUN(pl->range(pl, 0, n_az, st));
        STAGE_MARK(3);
        STAGE_MARK(4);
        RUN(pl->az_inv(pl, out, max_sq, st));
        STAGE_MARK(5);
    } else {
        STAGE_MARK(0);
        RUN(pl->outer_fwd(pl, in, pitch, 0, n_rg, st));
        STAGE_MARK(1);
        RUN(pl->inner(pl, false, 0, n_rg, 0, A1, st));
        STAGE_MARK(2);
        RUN(pl->range(pl, 0, n_az, st));
        STAGE_MARK(3);
        RUN(pl->inner(pl, true, 0, n_rg, 0, A1, st));
        STAGE_MARK(4);
        RUN(pl->outer_inv(pl, out, max_sq, 0, n_rg, st));
        STAGE_MARK(5);
    }
#undef RUN
#undef STAGE_MARK
    if (ev) pl->prof_calls++;
    return NIS_OK;
}

extern "C" int nis_csa_plan_set_profiling(nis_csa_plan* pl, int32_t enable) {
    NIS_REQUIRE(pl, "nis_csa_plan_set_profiling: null plan");
    if (enable && !pl->prof_ev[0][0]) {
        for (auto& set : pl->prof_ev)
            for (auto& e : set) NIS_CUDA_TRY(cudaEventCreate(&e));
    }
    pl->profiling = enable != 0;
    pl->prof_calls = 0;
    return NIS_OK;
}

extern "C" int nis_csa_stage_times(nis_csa_plan* pl, int32_t calls_back, float* ms5) {
    NIS_REQUIRE(pl && ms5, "nis_csa_stage_times: null argument");
    NIS_REQUIRE(pl->profiling && calls_back >= 0 && (uint64_t)calls_back < pl->prof_calls &&
                    calls_back < nis_csa_plan::kProfRing,
                "nis_csa_stage_times: no profiled call %d back (have %llu)", calls_back,
                (unsigned long long)pl->prof_calls);
    cudaEvent_t* ev = pl->prof_ev[(pl->prof_calls - 1 - calls_back) % nis_csa_plan::kProfRing];
    NIS_CUDA_TRY(cudaEventSynchronize(ev[5]));
    for (int i = 0; i < 5; ++i) NIS_CUDA_TRY(cudaEventElapsedTime(&ms5[i], ev[i], ev[i + 1]));
    return NIS_OK;
}
