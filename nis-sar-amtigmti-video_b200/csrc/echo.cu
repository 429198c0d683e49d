// echo.cu -- K1: scatterer x pulse x range-bin raw-echo accumulator.
//
// Replaces the per-pulse loops of run_bistatic_physics_gpu (sar_ati_dcpa_sim_csa.py:137-178),
// run_physics_engine (sar_satellite_sim.py:264-302), run_moving_physics
// (sar_satellite_moving_sim.py:129-156) and run_custom_physics (sar_vehicle_sim.py:102-123):
//   raw[i][n] = sum_b amp_b exp(j 2 pi (-fc tau_bi + (k/2) (t_n - tau_bi - T_p/2)^2)) [|t_n - tau_bi - T_p/2| <= T_p/2]
//
// Work split: one CTA = one pulse x one chunk of CH = 256*SPT consecutive samples; each thread owns
// SPT consecutive samples in registers and loops over every scatterer (no cross-thread reduction).
//  * Prologue (fp64, once per scatterer per CTA, 256 scatterers at a time into shared memory): exact
//    two-way range -> delay tau, the phase polynomial about the chunk centre reduced mod 1 and stored
//    as 32-bit fixed-point turns, and the closed-interval chirp support [lo, hi) evaluated with the
//    reference's own fp64 expression on the caller's sample-time table (boundary samples bit-exact).
//    Scatterers whose support misses the chunk are dropped by an ordered (deterministic) compaction.
//  * Inner loop (fp32): the phase is quadratic in the sample index, phase(m_t + jj) =
//    phi_t + jj*delta_t + A jj^2.  A is the same for every scatterer, so exp(j 2 pi A jj^2) factors
//    out of the scatterer sum and is applied once at the end; per scatterer a thread evaluates two
//    fast sincos (u = amp cis(phi_t), v = cis(delta_t)) and then runs the phasor recurrence
//    p <- p v outward from the centre sample: 1 complex multiply + 1 complex add per sample.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

using namespace nis;

namespace {

struct EchoConst {
    double c, fc, k_rate, t_p, t_start, dt_fast;
    double a_turns;  // (k/2) dt^2
    uint64_t a64;    // the same in 64-bit fixed-point turns (exact A m^2 mod 1 by integer wrap-around)
    int T, P0, S, per_target_velocity, accumulate, bistatic;
    int spotlight;   // 1: run_physics_spotlight model (sar_batch_sim.py:83-169): pos_rx carries the platform velocity
    double ant;      // pi l_ant / lambda of the one-way sinc^2 pattern (spotlight)
};

template <int SPT>
struct EchoTail {
    float2 e[SPT];  // exp(j 2 pi A jj^2), jj = j - SPT/2
};

__device__ __forceinline__ uint32_t frac32(double turns) {
    turns -= floor(turns);
    return (uint32_t)(unsigned long long)(turns * 4294967296.0);  // < 2^32 except turns==1-eps -> saturates below
}

// closed-interval gate of the reference, evaluated exactly as numpy does in fp64 (:164-166): |t_n - tau - off| <= half,
// off = T_p/2 for the engines whose chirp STARTS at the delay, 0 for the spotlight engine whose chirp is centred on it
__device__ __forceinline__ bool gate(const double* __restrict__ t_fast, int n, double tau, double off, double half) {
    return fabs(__dsub_rn(__dsub_rn(t_fast[n], tau), off)) <= half;
}

// Inner loop of both kernels: add the samples [t_lo, t_hi) (relative to the chunk start) of every kept scatterer of the
// chunk to the thread's accumulators.
// Control flow is decided per WARP, not per thread (late round 2; ncu on the bench scene had shown the lane that
// straddles an end of a chirp running a 78-instruction predicated copy of the sample loop on its own, after the other 31
// lanes had run theirs -- 15 % of all instructions -- and on short chirps, sar_vehicle_sim.py's 360 samples, every warp a
// chirp touches has such a lane): a chirp either misses the warp's 32 windows (skip), covers all of them (plain loop, no
// per-thread tests at all), or has an end among them -- then EVERY lane runs one predicated loop on a bit mask of its live
// samples (one LOP3-to-predicate and one predicated packed add per sample).
template <int SPT>
__device__ __forceinline__ void accumulate_kept(float2 (&acc)[SPT], const uint4* __restrict__ rec,
                                                const float2* __restrict__ recv, int total, int t_lo, int t_hi, int mt,
                                                uint32_t k1, float2 vk2) {
    constexpr int HALF = SPT / 2;
    const int wlo = t_lo - (int)(threadIdx.x & 31) * SPT, whi = wlo + 32 * SPT;   // the warp's samples
#pragma unroll 1
    for (int q = 0; q < total; ++q) {
        const uint4 s = rec[q];
        const int lo = (int)(s.w & 0xffffu), hi = (int)(s.w >> 16);
        if (wlo >= hi || whi <= lo) continue;          // (warp-uniform)
        const bool whole = lo <= wlo && hi >= whi;     // (warp-uniform) every window of the warp is covered completely
        const uint32_t phi = s.x + s.y * (uint32_t)mt + k1;
        const float2 u = cscale_pk(cis_u32(phi), __uint_as_float(s.z));
        const float2 v = cmul_pk(recv[q], vk2);
        // the samples of one scatterer are a geometric sequence u v^k, and any such sequence obeys the three-term
        // recurrence p[k+1] = 2 cos(delta) p[k] - p[k-1] (v + 1/v = 2 cos delta): TWO FMAs per new complex term instead
        // of the four of a complex multiply.  Two chains run outward from the centre sample (forward f, backward b),
        // real and imaginary parts independently: four independent FMA chains per scatterer.  The recurrence amplifies
        // an error of the coefficient by at most k^2 / 2 over k steps; k <= SPT/2 here (<= 3e-5 relative, reached only
        // where the chirp's instantaneous frequency is near zero).
        const float c2 = 2.0f * v.x;
        float2 f0 = u, f1 = cmul_pk(u, v);
        float2 b0 = cmul_conj_pk(u, v), b1 = cfms_pk(c2, b0, u);
        if (whole) {
#pragma unroll
            for (int i = 0; i < HALF; i += 2) {
                acc[HALF + i] = cadd_pk(acc[HALF + i], f0);
                acc[HALF + i + 1] = cadd_pk(acc[HALF + i + 1], f1);
                acc[HALF - 1 - i] = cadd_pk(acc[HALF - 1 - i], b0);
                acc[HALF - 2 - i] = cadd_pk(acc[HALF - 2 - i], b1);
                if (i + 2 < HALF) {
                    f0 = cfms_pk(c2, f1, f0);
                    f1 = cfms_pk(c2, f0, f1);
                    b0 = cfms_pk(c2, b1, b0);
                    b1 = cfms_pk(c2, b0, b1);
                }
            }
        } else {
            // bit j of `live` <-> sample t_lo + j lies in [lo, hi); lanes the chirp does not reach carry 0
            const int ea = max(lo - t_lo, 0), eb = min(hi - t_lo, SPT);
            const uint32_t live = ea < eb ? ((0xffffffffu >> (32 - eb)) & (0xffffffffu << ea)) : 0u;
#pragma unroll
            for (int i = 0; i < HALF; i += 2) {
                if (live & (1u << (HALF + i))) acc[HALF + i] = cadd_pk(acc[HALF + i], f0);
                if (live & (1u << (HALF + i + 1))) acc[HALF + i + 1] = cadd_pk(acc[HALF + i + 1], f1);
                if (live & (1u << (HALF - 1 - i))) acc[HALF - 1 - i] = cadd_pk(acc[HALF - 1 - i], b0);
                if (live & (1u << (HALF - 2 - i))) acc[HALF - 2 - i] = cadd_pk(acc[HALF - 2 - i], b1);
                if (i + 2 < HALF) {
                    f0 = cfms_pk(c2, f1, f0);
                    f1 = cfms_pk(c2, f0, f1);
                    b0 = cfms_pk(c2, b1, b0);
                    b1 = cfms_pk(c2, b0, b1);
                }
            }
        }
    }
}

// exact chirp support [lo, hi) of a delay on the caller's sample-time table: the closed-form estimate corrected by the
// reference's own fp64 gate expression (boundary samples bit-exact)
__device__ __forceinline__ void chirp_support(const EchoConst& k, const double* __restrict__ t_fast, double tau, double off,
                                              double half, int& lo, int& hi) {
    lo = (int)ceil((tau + off - half - k.t_start) / k.dt_fast);
    hi = (int)floor((tau + off + half - k.t_start) / k.dt_fast) + 1;
    lo = max(0, min(lo, k.S));
    hi = max(0, min(hi, k.S));
#pragma unroll 1
    for (int it = 0; it < 3 && lo > 0 && gate(t_fast, lo - 1, tau, off, half); ++it) --lo;
#pragma unroll 1
    for (int it = 0; it < 3 && lo < k.S && !gate(t_fast, lo, tau, off, half); ++it) ++lo;
#pragma unroll 1
    for (int it = 0; it < 3 && hi < k.S && gate(t_fast, hi, tau, off, half); ++it) ++hi;
#pragma unroll 1
    for (int it = 0; it < 3 && hi > 0 && !gate(t_fast, hi - 1, tau, off, half); ++it) --hi;
}

// ---------------------------------------------------------------------------------------------------------------------
// Sparse scenes (T <= 256 scatterers: point-target grids, a single ship): ONE CTA per pulse walks all chunks of the window.
// The fp64 geometry -- position, two norms, delay, exact chirp support -- is evaluated once per (scatterer, pulse) and kept
// in shared memory; per chunk only the phase polynomial about the chunk centre is re-expanded (a dozen fp64 operations).
// Measured on the bench scene (81 scatterers, 8192 x 8192): 263 M instead of 299 M warp instructions, 0.42 instead of
// 0.49 ms, no spills at 72 registers (profiles/prof_echo_sparse_r2.txt); later, at 64 registers (still no spills) and four
// CTAs per SM: 0.418 -> 0.396 ms (ncu had shown 24 of 64 warp slots in use and the schedulers idle 44 % of the time).  What remains is the inner loop itself: per
// (scatterer, warp) 13 instructions of loop control, 21 of per-scatterer set-up, 29 for the sixteen samples.  Tried and
// rejected (measured slower): 32 samples per thread (113 registers, two CTAs per SM: 0.64 ms); per-warp work lists that
// skip non-intersecting (scatterer, warp) pairs (dense kernel: +4 % on the vehicle and the default clutter scene -- the
// ballot-compacted list and its indirection cost more than the 13-instruction skips they remove); record tiles double
// buffered with one barrier per tile instead of two (no change: the barrier stalls ncu shows are load imbalance).
template <int SPT>
__global__ void __launch_bounds__(256, 4) k_echo_sparse(EchoConst k, EchoTail<SPT> tail, const double* __restrict__ pos0,
                                                        const double* __restrict__ vel, const double* __restrict__ amp,
                                                        const double* __restrict__ pos_tx, const double* __restrict__ pos_rx,
                                                        const double* __restrict__ t_slow, const double* __restrict__ t_fast,
                                                        float2* __restrict__ raw, int n_chunks) {
    constexpr int NTH = 256, CH = NTH * SPT, HALF = SPT / 2;
    __shared__ uint4 rec[256];
    __shared__ float2 recv[256];
    __shared__ double s_tau[256];
    __shared__ int2 s_sup[256];      // absolute support [lo, hi)
    __shared__ float s_amp[256];
    __shared__ int warp_cnt[8];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int pulse = k.P0 + blockIdx.x;
    const int mt = tid * SPT + HALF - CH / 2;
    const int t_lo = tid * SPT, t_hi = t_lo + SPT;
    const uint32_t k1 = frac32(k.a_turns * (double)mt * (double)mt);
    const uint32_t k2 = frac32(2.0 * k.a_turns * (double)mt);
    const float2 vk2 = cis_u32(k2);
    const double half = k.t_p / 2, off = half;

    // ---- once per pulse: geometry of scatterer `tid`
    if (tid < k.T) {
        const double ti = t_slow[pulse];
        const double tx0 = pos_tx[3 * pulse], tx1 = pos_tx[3 * pulse + 1], tx2 = pos_tx[3 * pulse + 2];
        const double* vb = k.per_target_velocity ? vel + 3 * tid : vel;
        const double px = __dadd_rn(pos0[3 * tid], __dmul_rn(vb[0], ti)),
                     py = __dadd_rn(pos0[3 * tid + 1], __dmul_rn(vb[1], ti)),
                     pz = __dadd_rn(pos0[3 * tid + 2], __dmul_rn(vb[2], ti));
        double dx = px - tx0, dy = py - tx1, dz = pz - tx2;
        const double d_tx = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        double tau;
        if (k.bistatic) {
            dx = px - pos_rx[3 * pulse]; dy = py - pos_rx[3 * pulse + 1]; dz = pz - pos_rx[3 * pulse + 2];
            const double d_rx = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
            tau = __ddiv_rn(__dadd_rn(d_tx, d_rx), k.c);
        } else {
            tau = __ddiv_rn(__dmul_rn(2.0, d_tx), k.c);
        }
        int lo, hi;
        chirp_support(k, t_fast, tau, off, half, lo, hi);
        s_tau[tid] = tau;
        s_sup[tid] = make_int2(lo, hi);
        s_amp[tid] = (float)amp[tid];
    }
    __syncthreads();

    for (int chunk = 0; chunk < n_chunks; ++chunk) {
        const int n0 = chunk * CH, nc = n0 + CH / 2;
        // ---- per chunk: phase polynomial about the chunk centre, ordered compaction of the scatterers that reach it
        bool keep = false;
        uint4 r = make_uint4(0, 0, 0, 0);
        if (tid < k.T) {
            const int2 sup = s_sup[tid];
            const int rlo = max(sup.x - n0, 0), rhi = min(sup.y - n0, CH);
            if (rlo < rhi) {
                keep = true;
                const double tau = s_tau[tid];
                const double uu = (k.t_start + (double)nc * k.dt_fast) - tau - off;
                r.x = frac32(fma(0.5 * k.k_rate * uu, uu, -k.fc * tau));
                r.y = frac32(k.k_rate * uu * k.dt_fast);
                r.z = __float_as_uint(s_amp[tid]);
                r.w = (uint32_t)rlo | ((uint32_t)rhi << 16);
            }
        }
        const unsigned ball = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[wid] = __popc(ball);
        __syncthreads();   // also: the previous chunk's consumers are done with rec[]
        int slot = __popc(ball & ((1u << lane) - 1u));
        int total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int cw = warp_cnt[w];
            if (w < wid) slot += cw;
            total += cw;
        }
        if (keep) {
            rec[slot] = r;
            recv[slot] = cis_u32(r.y);
        }
        __syncthreads();
        if (n0 + t_lo >= k.S) continue;       // threads past the end of the window (uniform barriers above)
        float2 acc[SPT];
#pragma unroll
        for (int j = 0; j < SPT; ++j) acc[j] = make_float2(0.f, 0.f);
        accumulate_kept<SPT>(acc, rec, recv, total, t_lo, t_hi, mt, k1, vk2);
        float2* out = raw + (int64_t)pulse * k.S + n0 + t_lo;
        if (k.accumulate == 0 && n0 + t_hi <= k.S) {   // whole thread inside the window: 16-byte stores
#pragma unroll
            for (int j = 0; j < SPT; j += 2) {
                const float2 x0 = cmul(acc[j], tail.e[j]), x1 = cmul(acc[j + 1], tail.e[j + 1]);
                *reinterpret_cast<float4*>(out + j) = make_float4(x0.x, x0.y, x1.x, x1.y);
            }
        } else {
#pragma unroll
            for (int j = 0; j < SPT; ++j) {
                if (n0 + t_lo + j < k.S) {
                    float2 x = cmul(acc[j], tail.e[j]);
                    if (k.accumulate == 2) {
                        atomicAdd_system(&out[j].x, x.x);
                        atomicAdd_system(&out[j].y, x.y);
                    } else {
                        if (k.accumulate) { const float2 o = out[j]; x.x += o.x; x.y += o.y; }
                        out[j] = x;
                    }
                }
            }
        }
    }
}

// Sparse scenes, second form (late round 2): the windows of a pulse are handed to the warps DYNAMICALLY.
// In k_echo_sparse a warp owns one fixed window (32 x SPT samples) of every chunk and the CTA meets at a barrier per chunk:
// with 81 chirps that cover a fifth of the sample window most (scatterer, window) pairs miss, the warps whose windows no
// chirp reaches idle at the barrier, and the ones that work spend a fifth of their instructions skipping records (ncu:
// issue slots 56 % busy, 24-32 resident warps of which about half own work).  Here a CTA is four independent warps; the
// geometry of all scatterers is evaluated once per pulse, with the phase polynomial expanded about the PULSE centre in
// 64-bit fixed point (exact over +-4096 samples, so it need not be re-expanded per chunk); then every warp repeatedly
// takes the next window of the pulse from a shared counter, builds the list of the chirps that reach this window by
// ballot (3 warp-wide tests for 81 scatterers) and runs the recurrence loop over that list only.  No barrier after the
// prologue, no skipped records except at the two ends of a chirp.
template <int SPT>
__global__ void __launch_bounds__(128, 7) k_echo_sparse_dyn(EchoConst k, EchoTail<SPT> tail, const double* __restrict__ pos0,
                                                            const double* __restrict__ vel, const double* __restrict__ amp,
                                                            const double* __restrict__ pos_tx, const double* __restrict__ pos_rx,
                                                            const double* __restrict__ t_slow, const double* __restrict__ t_fast,
                                                            float2* __restrict__ raw, int n_windows) {
    constexpr int NTH = 128, NW = NTH / 32, HALF = SPT / 2, WIN = 32 * SPT;
    __shared__ uint64_t s_cq[256], s_bq[256];   // phase at the pulse centre and per sample, 2^-64 turns
    __shared__ int2 s_sup[256];                 // absolute support [lo, hi)
    __shared__ float s_amp[256];
    __shared__ uint4 w_rec[NW][256];            // per warp: the records of the window in hand (format of accumulate_kept)
    __shared__ float2 w_recv[NW][256];
    __shared__ int s_next;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int pulse = k.P0 + blockIdx.x;
    const int nc = k.S / 2;
    const double half = k.t_p / 2, off = half;
    // this thread's samples relative to the centre of whatever window its warp holds: constants of the kernel
    const int t_lo = lane * SPT, t_hi = t_lo + SPT;
    const int mt = t_lo + HALF - WIN / 2;
    const uint32_t k1 = frac32(k.a_turns * (double)mt * (double)mt);
    const float2 vk2 = cis_u32(frac32(2.0 * k.a_turns * (double)mt));

    // ---- once per pulse: geometry and phase polynomial of every scatterer
    {
        const double ti = t_slow[pulse];
        const double tx0 = pos_tx[3 * pulse], tx1 = pos_tx[3 * pulse + 1], tx2 = pos_tx[3 * pulse + 2];
        const double t_center = k.t_start + (double)nc * k.dt_fast;
        for (int b = tid; b < k.T; b += NTH) {
            const double* vb = k.per_target_velocity ? vel + 3 * b : vel;
            const double px = __dadd_rn(pos0[3 * b], __dmul_rn(vb[0], ti)),
                         py = __dadd_rn(pos0[3 * b + 1], __dmul_rn(vb[1], ti)),
                         pz = __dadd_rn(pos0[3 * b + 2], __dmul_rn(vb[2], ti));
            double dx = px - tx0, dy = py - tx1, dz = pz - tx2;
            const double d_tx = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
            double tau;
            if (k.bistatic) {
                dx = px - pos_rx[3 * pulse]; dy = py - pos_rx[3 * pulse + 1]; dz = pz - pos_rx[3 * pulse + 2];
                const double d_rx = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
                tau = __ddiv_rn(__dadd_rn(d_tx, d_rx), k.c);
            } else {
                tau = __ddiv_rn(__dmul_rn(2.0, d_tx), k.c);
            }
            int lo, hi;
            chirp_support(k, t_fast, tau, off, half, lo, hi);
            const double uu = t_center - tau - off;
            s_cq[b] = turns_to_u64(fma(0.5 * k.k_rate * uu, uu, -k.fc * tau));
            s_bq[b] = turns_to_u64(k.k_rate * uu * k.dt_fast);
            s_sup[b] = make_int2(lo, hi);
            s_amp[b] = (float)amp[b];
        }
        if (tid == 0) s_next = 0;
    }
    __syncthreads();

    uint4* const rec = w_rec[wid];
    float2* const recv = w_recv[wid];
    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(&s_next, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n_windows) break;
        const int w_lo = w * WIN, w_hi = min(w_lo + WIN, k.S);
        // ---- the chirps that reach this window, in scatterer order (deterministic summation order): each lane tests
        // one scatterer and re-expands its phase polynomial about the WINDOW centre (64-bit integer arithmetic, exact)
        const uint64_t wc = (uint64_t)(int64_t)(w_lo + WIN / 2 - nc);
        int cnt = 0;
        for (int b0 = 0; b0 < k.T; b0 += 32) {
            const int b = b0 + lane;
            bool ov = false;
            int2 sup = make_int2(0, 0);
            if (b < k.T) {
                sup = s_sup[b];
                ov = sup.x < w_hi && sup.y > w_lo;
            }
            const unsigned ball = __ballot_sync(0xffffffffu, ov);
            if (ov) {
                const uint64_t bq = s_bq[b];
                const uint64_t c64 = s_cq[b] + bq * wc + k.a64 * (wc * wc);
                const uint64_t b64 = bq + 2ull * k.a64 * wc;
                const int slot = cnt + __popc(ball & ((1u << lane) - 1u));
                uint4 r;
                r.x = (uint32_t)(c64 >> 32);
                r.y = (uint32_t)(b64 >> 32);
                r.z = __float_as_uint(s_amp[b]);
                r.w = (uint32_t)max(sup.x - w_lo, 0) | ((uint32_t)min(sup.y - w_lo, WIN) << 16);
                rec[slot] = r;
                recv[slot] = cis_u32(r.y);
            }
            cnt += __popc(ball);
        }
        __syncwarp();
        float2 acc[SPT];
#pragma unroll
        for (int j = 0; j < SPT; ++j) acc[j] = make_float2(0.f, 0.f);
        accumulate_kept<SPT>(acc, rec, recv, cnt, t_lo, t_hi, mt, k1, vk2);
        __syncwarp();   // the list is rebuilt for the next window
        if (w_lo + t_lo >= k.S) continue;
        float2* out = raw + (int64_t)pulse * k.S + w_lo + t_lo;
        if (k.accumulate == 0 && w_lo + t_hi <= k.S) {   // whole thread inside the sample window: 16-byte stores
#pragma unroll
            for (int j = 0; j < SPT; j += 2) {
                const float2 x0 = cmul(acc[j], tail.e[j]), x1 = cmul(acc[j + 1], tail.e[j + 1]);
                *reinterpret_cast<float4*>(out + j) = make_float4(x0.x, x0.y, x1.x, x1.y);
            }
        } else {
#pragma unroll
            for (int j = 0; j < SPT; ++j) {
                if (w_lo + t_lo + j < k.S) {
                    float2 x = cmul(acc[j], tail.e[j]);
                    if (k.accumulate == 2) {
                        atomicAdd_system(&out[j].x, x.x);
                        atomicAdd_system(&out[j].y, x.y);
                    } else {
                        if (k.accumulate) { const float2 o = out[j]; x.x += o.x; x.y += o.y; }
                        out[j] = x;
                    }
                }
            }
        }
    }
}

// SPOT: the spotlight model is a separate instantiation, so that the stripmap engines keep their register budget
// W256: the CTA has all 256 threads (chunk width and tile size are compile-time constants)
template <int SPT, bool SPOT, bool W256>
// four resident CTAs per SM (<= 64 registers; 24 bytes of spills at 16 samples per thread): measured -- 3 CTAs at 75
// registers are 7 % slower on sparse scenes and equal on dense ones, 2 CTAs at 94 registers 12-25 % slower
__global__ void __launch_bounds__(256, SPOT ? 1 : 4) k_echo(EchoConst k, EchoTail<SPT> tail, const double* __restrict__ pos0,
                                              const double* __restrict__ vel, const double* __restrict__ amp,
                                              const double* __restrict__ pos_tx, const double* __restrict__ pos_rx,
                                              const double* __restrict__ t_slow, const double* __restrict__ t_fast,
                                              float2* __restrict__ raw) {
    const int NTH = W256 ? 256 : blockDim.x;   // 32 .. 256 threads: the launcher balances the chunks over the sample window
    const int CH = NTH * SPT;
    constexpr int HALF = SPT / 2;
    __shared__ uint4 rec[256];
    __shared__ float2 recv[256];   // cis(per-sample phase step) of each kept scatterer at the chunk centre
    __shared__ int warp_cnt[8];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int pulse = k.P0 + blockIdx.y;
    const int n0 = blockIdx.x * CH;                 // first sample of the chunk
    const int nc = n0 + CH / 2;                     // chunk centre: phase polynomial is expanded about it
    const int mt = tid * SPT + HALF - CH / 2;       // this thread's centre sample relative to nc (signed)
    const int t_lo = tid * SPT, t_hi = t_lo + SPT;  // this thread's samples relative to n0

    // per-thread constants of the common quadratic term: A mt^2 and 2 A mt (mod 1)
    const uint32_t k1 = frac32(k.a_turns * (double)mt * (double)mt);
    const uint32_t k2 = frac32(2.0 * k.a_turns * (double)mt);
    const float2 vk2 = cis_u32(k2);   // this thread's share of the phase step: cis(delta_t) = cis(s.y) cis(k2)

    const double ti = t_slow[pulse];
    const double tx0 = pos_tx[3 * pulse], tx1 = pos_tx[3 * pulse + 1], tx2 = pos_tx[3 * pulse + 2];
    double rx0 = tx0, rx1 = tx1, rx2 = tx2;
    if (k.bistatic || SPOT) { rx0 = pos_rx[3 * pulse]; rx1 = pos_rx[3 * pulse + 1]; rx2 = pos_rx[3 * pulse + 2]; }
    const double half = k.t_p / 2;
    const double off = SPOT ? 0.0 : half;   // chirp centre relative to the delay
    const double t_center = k.t_start + (double)nc * k.dt_fast;

    float2 acc[SPT];
#pragma unroll
    for (int j = 0; j < SPT; ++j) acc[j] = make_float2(0.f, 0.f);

    if (!W256) {
        if (tid < 8) warp_cnt[tid] = 0;    // warps a narrower CTA does not have
        __syncthreads();
    }
    for (int b0 = 0; b0 < k.T; b0 += NTH) {
        // ------------------------------------------------ prologue: one scatterer per thread, fp64
        const int b = b0 + tid;
        bool keep = false;
        uint4 r = make_uint4(0, 0, 0, 0);
        if (b < k.T) {
            const double* vb = k.per_target_velocity ? vel + 3 * b : vel;
            // p = p0 + v t with separate roundings, as numpy / torch evaluate it (:151)
            const double px = __dadd_rn(pos0[3 * b], __dmul_rn(vb[0], ti)),
                         py = __dadd_rn(pos0[3 * b + 1], __dmul_rn(vb[1], ti)),
                         pz = __dadd_rn(pos0[3 * b + 2], __dmul_rn(vb[2], ti));
            double dx = px - tx0, dy = py - tx1, dz = pz - tx2;
            const double d_tx = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
            double tau, gain = 1.0;
            if constexpr (SPOT) {
                // start-stop correction: the receive position is the platform advanced by v * 2 d_tx / c (:133-137)
                const double ta = __ddiv_rn(__dmul_rn(2.0, d_tx), k.c);
                const double ex = px - (tx0 + rx0 * ta), ey = py - (tx1 + rx1 * ta), ez = pz - (tx2 + rx2 * ta);
                const double d_rx = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez)));
                tau = __ddiv_rn(__dadd_rn(d_tx, d_rx), k.c);
                // one-way sinc^2 pattern of the aperture steered at the scene centre (:139-149)
                const double bn = sqrt(tx0 * tx0 + tx1 * tx1 + tx2 * tx2);
                double co = -(tx0 * dx + tx1 * dy + tx2 * dz) / (bn * d_tx);
                co = fmin(1.0, fmax(-1.0, co));
                const double xv = k.ant * sin(acos(co));
                if (fabs(xv) > 1e-6) { const double sc = sin(xv) / xv; gain = sc * sc; }
            } else if (k.bistatic) {   // (not reached in the spotlight instantiation)
                dx = px - rx0; dy = py - rx1; dz = pz - rx2;
                const double d_rx = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
                tau = __ddiv_rn(__dadd_rn(d_tx, d_rx), k.c);
            } else {
                tau = __ddiv_rn(__dmul_rn(2.0, d_tx), k.c);
            }
            // support [lo, hi) in absolute sample indices
            int lo = (int)ceil((tau + off - half - k.t_start) / k.dt_fast);
            int hi = (int)floor((tau + off + half - k.t_start) / k.dt_fast) + 1;
            lo = max(0, min(lo, k.S));
            hi = max(0, min(hi, k.S));
#pragma unroll 1
            for (int it = 0; it < 3 && lo > 0 && gate(t_fast, lo - 1, tau, off, half); ++it) --lo;
#pragma unroll 1
            for (int it = 0; it < 3 && lo < k.S && !gate(t_fast, lo, tau, off, half); ++it) ++lo;
#pragma unroll 1
            for (int it = 0; it < 3 && hi < k.S && gate(t_fast, hi, tau, off, half); ++it) ++hi;
#pragma unroll 1
            for (int it = 0; it < 3 && hi > 0 && !gate(t_fast, hi - 1, tau, off, half); ++it) --hi;
            const int rlo = max(lo - n0, 0), rhi = min(hi - n0, CH);
            if (rlo < rhi) {
                keep = true;
                const double uu = t_center - tau - off;
                const double cq = fma(0.5 * k.k_rate * uu, uu, -k.fc * tau);  // turns at the chunk centre
                const double bq = k.k_rate * uu * k.dt_fast;                  // turns per sample
                r.x = frac32(cq);
                r.y = frac32(bq);
                r.z = __float_as_uint((float)(amp[b] * gain));
                r.w = (uint32_t)rlo | ((uint32_t)rhi << 16);
            }
        }
        // ordered compaction (deterministic summation order)
        const unsigned ball = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[wid] = __popc(ball);
        __syncthreads();  // also: previous tile's consumers are done with rec[]
        int off = __popc(ball & ((1u << lane) - 1u));
        int total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int cw = warp_cnt[w];
            if (w < wid) off += cw;
            total += cw;
        }
        if (keep) {
            rec[off] = r;
            recv[off] = cis_u32(r.y);
        }
        __syncthreads();

        // ------------------------------------------------ inner loop: every thread, every kept scatterer
        accumulate_kept<SPT>(acc, rec, recv, total, t_lo, t_hi, mt, k1, vk2);
    }

    // ------------------------------------------------ epilogue: common quadratic factor, store
    float2* out = raw + (int64_t)pulse * k.S + n0 + t_lo;
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
        if (n0 + t_lo + j < k.S) {
            float2 x = cmul(acc[j], tail.e[j]);
            if (k.accumulate == 2) {
                // several GPUs add their scatterer shards into one owner's rows, possibly over NVLink: fire-and-forget
                // reductions (RED.E.ADD.F32 ... .SYS), no read-modify-write round trip.  SYSTEM scope: concurrent updates of one
                // address by several devices are only defined for .sys atomics, and only where the link performs them natively
                // (nis_peer_native_atomics; the Python layer falls back to the NCCL reduce otherwise)
                atomicAdd_system(&out[j].x, x.x);
                atomicAdd_system(&out[j].y, x.y);
            } else {
                if (k.accumulate) { const float2 o = out[j]; x.x += o.x; x.y += o.y; }
                out[j] = x;
            }
        }
    }
}

// Chunks of equal width over the window: as few CTAs per pulse as 256 threads x SPT samples allow, each with just the
// warps it needs (13200 samples at 16 per thread: 4 CTAs of 224 threads instead of 4 x 256 with a quarter of them idle).
struct EchoShape { int chunks, threads; };
inline EchoShape echo_shape(int S, int spt) {
    EchoShape e;
    e.chunks = (S + 256 * spt - 1) / (256 * spt);
    const int per = (S + e.chunks - 1) / e.chunks;
    e.threads = ((per + spt - 1) / spt + 31) / 32 * 32;
    return e;
}

template <int SPT>
int launch_echo(nis_ctx* ctx, const EchoConst& k, const double* pos0, const double* vel, const double* amp,
                const double* pos_tx, const double* pos_rx, const double* t_slow, const double* t_fast, float2* raw,
                int n_pulses, cudaStream_t st) {
    EchoTail<SPT> tail;
    const double two_pi = 6.283185307179586476925286766559;
    for (int j = 0; j < SPT; ++j) {
        const double jj = (double)(j - SPT / 2);
        double ph = k.a_turns * jj * jj;
        ph -= floor(ph);
        tail.e[j] = make_float2((float)cos(two_pi * ph), (float)sin(two_pi * ph));
    }
    const EchoShape sh = echo_shape(k.S, SPT);
    dim3 grid(sh.chunks, n_pulses);
    // sparse scenes: one CTA per pulse, geometry once per (scatterer, pulse).  Needs enough pulses to fill the GPU and rows
    // whose 16-byte vector stores are aligned (S even); NIS_ECHO_SPARSE=0 disables it (development knob)
    const char* sparse_env = getenv("NIS_ECHO_SPARSE");
    const bool sparse_ok = !(sparse_env && sparse_env[0] == '0');
    if (sparse_ok && !k.spotlight && k.T > 0 && k.T <= 256 && n_pulses >= 2 * ctx->num_sms && (k.S % 2) == 0 &&
        (((uintptr_t)raw) & 15) == 0) {
        // NIS_ECHO_SPARSE=chunk: the first form (fixed windows, one barrier per chunk); default: windows handed out dynamically
        if (!(sparse_env && sparse_env[0] == 'c')) {
            const int n_windows = (k.S + 32 * SPT - 1) / (32 * SPT);
            k_echo_sparse_dyn<SPT><<<n_pulses, 128, 0, st>>>(k, tail, pos0, vel, amp, pos_tx, pos_rx, t_slow, t_fast, raw, n_windows);
            NIS_LAUNCH_CHECK(ctx);
            return NIS_OK;
        }
        const int n_chunks = (k.S + 256 * SPT - 1) / (256 * SPT);
        k_echo_sparse<SPT><<<n_pulses, 256, 0, st>>>(k, tail, pos0, vel, amp, pos_tx, pos_rx, t_slow, t_fast, raw, n_chunks);
        NIS_LAUNCH_CHECK(ctx);
        return NIS_OK;
    }
    if (k.spotlight) k_echo<SPT, true, false><<<grid, sh.threads, 0, st>>>(k, tail, pos0, vel, amp, pos_tx, pos_rx, t_slow, t_fast, raw);
    else if (sh.threads == 256) k_echo<SPT, false, true><<<grid, 256, 0, st>>>(k, tail, pos0, vel, amp, pos_tx, pos_rx, t_slow, t_fast, raw);
    else k_echo<SPT, false, false><<<grid, sh.threads, 0, st>>>(k, tail, pos0, vel, amp, pos_tx, pos_rx, t_slow, t_fast, raw);
    NIS_LAUNCH_CHECK(ctx);
    return NIS_OK;
}

}  // namespace

static int echo_common(nis_ctx* ctx, const nis_echo_params* prm, const double* tgt_pos0, const double* tgt_vel,
                       const double* tgt_amp, const double* pos_tx, const double* pos_rx, const double* t_slow,
                       const double* t_fast, int32_t T, int32_t P0, int32_t P1, int32_t S, nis_c32* raw, int32_t accumulate,
                       int spotlight, double ant, nis_stream stream) {
    NIS_REQUIRE(ctx && prm && tgt_pos0 && tgt_vel && tgt_amp && pos_tx && t_slow && t_fast && raw,
                "nis_echo_accumulate: null argument");
    NIS_REQUIRE(T >= 0 && S > 0 && P0 >= 0 && P1 >= P0, "nis_echo_accumulate: bad sizes T=%d S=%d P0=%d P1=%d", T, S, P0, P1);
    NIS_REQUIRE(prm->c > 0 && prm->dt_fast > 0 && prm->t_p > 0, "nis_echo_accumulate: non-physical parameters");
    if (P1 == P0) return NIS_OK;
    NIS_REQUIRE(P1 - P0 <= 65535, "nis_echo_accumulate: at most 65535 pulses per call (got %d)", P1 - P0);
    cudaStream_t st = (cudaStream_t)stream;
    EchoConst k;
    k.c = prm->c; k.fc = prm->fc; k.k_rate = prm->k_rate; k.t_p = prm->t_p;
    k.t_start = prm->t_start; k.dt_fast = prm->dt_fast;
    k.a_turns = 0.5 * prm->k_rate * prm->dt_fast * prm->dt_fast;
    k.a64 = turns_to_u64(k.a_turns);
    k.T = T; k.P0 = P0; k.S = S; k.per_target_velocity = prm->per_target_velocity;
    k.accumulate = accumulate; k.bistatic = (pos_rx != nullptr) && !spotlight;
    k.spotlight = spotlight; k.ant = ant;
    float2* r = reinterpret_cast<float2*>(raw);
    // chunk = 256*SPT samples: take the wider chunk unless it wastes > 12 % of its threads past S, or the
    // caller knows that the chirps cover only part of the window (samples_per_thread hint)
    const EchoShape s16 = echo_shape(S, 16);
    const int waste16 = s16.chunks * s16.threads * 16 - S;
    int want = prm->samples_per_thread;
    if (const char* e = getenv("NIS_ECHO_SPT")) want = atoi(e);
    // ... and unless the wider chunk leaves the GPU short of threads (few pulses x few chunks: 1e5 scatterers on a 256-pulse
    // block of 2048-sample rows ran 18 -> 26 ms with 128-thread CTAs)
    const bool fills = (int64_t)(P1 - P0) * s16.chunks * s16.threads >= (int64_t)ctx->num_sms * 1024;
    const bool wide = want == 16 || (want != 8 && waste16 * 8 <= S && fills);
    if (wide)
        return launch_echo<16>(ctx, k, tgt_pos0, tgt_vel, tgt_amp, pos_tx, pos_rx, t_slow, t_fast, r, P1 - P0, st);
    return launch_echo<8>(ctx, k, tgt_pos0, tgt_vel, tgt_amp, pos_tx, pos_rx, t_slow, t_fast, r, P1 - P0, st);
}

extern "C" int nis_echo_accumulate(nis_ctx* ctx, const nis_echo_params* prm, const double* tgt_pos0,
                                   const double* tgt_vel, const double* tgt_amp, const double* pos_tx,
                                   const double* pos_rx, const double* t_slow, const double* t_fast, int32_t T,
                                   int32_t P0, int32_t P1, int32_t S, nis_c32* raw, int32_t accumulate,
                                   nis_stream stream) {
    return echo_common(ctx, prm, tgt_pos0, tgt_vel, tgt_amp, pos_tx, pos_rx, t_slow, t_fast, T, P0, P1, S, raw, accumulate,
                       0, 0.0, stream);
}

extern "C" int nis_echo_spotlight(nis_ctx* ctx, const nis_echo_params* prm, const double* tgt_pos0, const double* tgt_vel,
                                  const double* tgt_rcs, const double* pos_sat, const double* vel_sat,
                                  const double* t_slow, const double* t_fast, int32_t T, int32_t P0, int32_t P1, int32_t S,
                                  double ant_pi_l_over_lambda, nis_c32* raw, int32_t accumulate, nis_stream stream) {
    NIS_REQUIRE(vel_sat != nullptr, "nis_echo_spotlight: the platform velocity is required (start-stop correction)");
    return echo_common(ctx, prm, tgt_pos0, tgt_vel, tgt_rcs, pos_sat, vel_sat, t_slow, t_fast, T, P0, P1, S, raw, accumulate,
                       1, ant_pi_l_over_lambda, stream);
}
