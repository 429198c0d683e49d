// gmti.cu -- K3: fused DPCA subtraction + ATI conjugate multiply + phase + threshold + compaction.
//
// Replaces the seven numpy passes of sar_ati_dcpa_sim_csa.py:414-419, :447-449 (and
// SARData.compute_all, sar_ati_dcpa_viewer_csa.py:42-52) by ONE pass over the channel pair:
//   k_gmti_init   zeroes the tile-status words / ticket of the caller's workspace, seeds the result record
//   k_gmti_max    read slc1 -> max |slc1|^2 (fp64, exact on fp32 samples); SKIPPED when nis_csa_focus already
//                 produced it while writing slc1 (the normal case on the focusing path)
//   k_gmti_fused  read slc1, slc2 (16 B), write every requested product (<= 33 B), flag detections, and compact
//                 them in the same kernel: a CTA owns a 2048-pixel tile (ticket order), counts its detections with
//                 warp ballots, obtains the number of detections in all earlier tiles by a decoupled look-back over
//                 per-tile status words (aggregate / inclusive prefix, one 64-bit word each) and writes its indices
//                 straight to their final positions -- ascending flat indices == np.flatnonzero(mag_mask).
// The detection test |slc1| > frac * max|slc1| is the reference's strict '>' (:447) evaluated in fp64
// on the fp32 samples, so the index list is bit-exact against numpy given the same SLC.
#include <math.h>

#include "common.cuh"

using namespace nis;

namespace {

constexpr int kTile = 2048;              // pixels per CTA (256 threads x 8)
constexpr int kIters = kTile / 512;      // a warp covers 64 consecutive pixels per iteration
static_assert(kIters * 8 == 32, "one warp scans the per-(iteration, warp) detection counts");
constexpr size_t kWsHeader = 16;         // workspace: u32 ticket (+ pad), then one u64 status word per tile

constexpr unsigned long long kFlagAggregate = 1ull << 62, kFlagPrefix = 2ull << 62, kFlagMask = 3ull << 62;

struct GmtiOut {
    float2* interf; float* phase; float2* diff; float* dpca_mag; float* slc1_mag;
    uint8_t* mask; float* phase_masked;
};

__global__ void __launch_bounds__(1024) k_gmti_init(nis_gmti_result* res, const double* __restrict__ max_sq_in,
                                                    uint32_t* __restrict__ ticket, unsigned long long* __restrict__ status,
                                                    int n_tiles) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tiles) status[i] = 0ull;
    if (i == 0) {
        *ticket = 0u;
        res->det_count = 0;
        res->peak_idx = 0xFFFFFFFFu;
        res->max_mag_sq = max_sq_in ? *max_sq_in : 0.0;
    }
}

__device__ __forceinline__ double sq_mag(float2 s) {
    const double re = (double)s.x, im = (double)s.y;
    return fma(re, re, im * im);  // both products are exact in fp64; one rounding
}

__global__ void __launch_bounds__(256) k_gmti_max(const float2* __restrict__ slc1, uint64_t n, nis_gmti_result* res) {
    double m = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        m = fmax(m, sq_mag(__ldg(slc1 + i)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ double wm[8];
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmax(m, wm[w]);
        // non-negative doubles order like their bit patterns
        atomicMax(reinterpret_cast<unsigned long long*>(&res->max_mag_sq), (unsigned long long)__double_as_longlong(m));
    }
}

struct PixelOut {
    float2 itf, df;
    float ph, dmag, mag1, phm;
    bool det;
};

__device__ __forceinline__ PixelOut gmti_pixel(float2 s1, float2 s2, float2 cal, int use_cal, double max_sq, double thr,
                                               double lo_sq, double hi_sq, uint32_t idx, nis_gmti_result* res) {
    PixelOut o;
    if (use_cal) s2 = cmul(s2, cal);
    const double sq = sq_mag(s1);
    o.det = sq > hi_sq ? true : (sq < lo_sq ? false : (sqrt(sq) > thr));
    if (sq == max_sq) atomicMin(&res->peak_idx, idx);
    o.itf = cmul_conj(s1, s2);
    o.ph = atan2f(o.itf.y, o.itf.x);
    o.df = csub(s1, s2);
    o.dmag = sqrtf(fmaf(o.df.x, o.df.x, o.df.y * o.df.y));
    o.mag1 = sqrtf(fmaf(s1.x, s1.x, s1.y * s1.y));
    o.phm = o.det ? o.ph : 0.f;
    return o;
}

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Two adjacent pixels per thread.  VEC: 16-byte loads and 16- / 8- / 2-byte stores (every pointer suitably aligned);
// otherwise element-wise accesses (views at odd element offsets).
template <bool VEC>
__global__ void __launch_bounds__(256, 4) k_gmti_fused(const float2* __restrict__ slc1, const float2* __restrict__ slc2,
                                                    uint64_t n, double thresh_frac, float2 cal, int use_cal, GmtiOut o,
                                                    uint32_t* __restrict__ ticket, unsigned long long* __restrict__ status,
                                                    int n_tiles, uint32_t* __restrict__ det_idx, uint32_t det_cap,
                                                    nis_gmti_result* res) {
    __shared__ uint32_t s_tile, s_base;
    __shared__ uint32_t s_cnt[32];           // detections of group (iteration, warp) -> exclusive offset inside the tile
    __shared__ uint2 s_bal[32];              // the group's detection ballots (even pixels, odd pixels)
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);   // tiles are numbered in the order CTAs start: a tile only
    __syncthreads();                                        // ever waits for tiles that are already running
    const uint32_t tile = s_tile;
    const double max_sq = res->max_mag_sq;
    // np.max(np.abs(slc1)) * frac, compared against np.abs(slc1): both sides are sqrt of the fp64 |.|^2
    const double thr = sqrt(max_sq) * thresh_frac;
    const double thr_sq = thr * thr;
    const double lo_sq = thr_sq * (1.0 - 1e-12), hi_sq = thr_sq * (1.0 + 1e-12);
    const uint64_t tile_base = (uint64_t)tile * kTile;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll 2
    for (int it = 0; it < kIters; ++it) {
        const uint64_t i = tile_base + (uint64_t)it * 512 + wid * 64 + 2 * lane;
        bool d0 = false, d1 = false;
        if (i + 1 < n) {
            float2 a0, a1, b0, b1;
            if constexpr (VEC) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(slc1 + i));
                const float4 b = __ldg(reinterpret_cast<const float4*>(slc2 + i));
                a0 = make_float2(a.x, a.y); a1 = make_float2(a.z, a.w);
                b0 = make_float2(b.x, b.y); b1 = make_float2(b.z, b.w);
            } else {
                a0 = __ldg(slc1 + i); a1 = __ldg(slc1 + i + 1);
                b0 = __ldg(slc2 + i); b1 = __ldg(slc2 + i + 1);
            }
            const PixelOut p0 = gmti_pixel(a0, b0, cal, use_cal, max_sq, thr, lo_sq, hi_sq, (uint32_t)i, res);
            const PixelOut p1 = gmti_pixel(a1, b1, cal, use_cal, max_sq, thr, lo_sq, hi_sq, (uint32_t)i + 1u, res);
            d0 = p0.det; d1 = p1.det;
            if constexpr (VEC) {
                if (o.interf) *reinterpret_cast<float4*>(o.interf + i) = make_float4(p0.itf.x, p0.itf.y, p1.itf.x, p1.itf.y);
                if (o.phase) *reinterpret_cast<float2*>(o.phase + i) = make_float2(p0.ph, p1.ph);
                if (o.diff) *reinterpret_cast<float4*>(o.diff + i) = make_float4(p0.df.x, p0.df.y, p1.df.x, p1.df.y);
                if (o.dpca_mag) *reinterpret_cast<float2*>(o.dpca_mag + i) = make_float2(p0.dmag, p1.dmag);
                if (o.slc1_mag) *reinterpret_cast<float2*>(o.slc1_mag + i) = make_float2(p0.mag1, p1.mag1);
                if (o.mask) *reinterpret_cast<uchar2*>(o.mask + i) = make_uchar2(d0 ? 1 : 0, d1 ? 1 : 0);
                if (o.phase_masked) *reinterpret_cast<float2*>(o.phase_masked + i) = make_float2(p0.phm, p1.phm);
            } else {
                if (o.interf) { o.interf[i] = p0.itf; o.interf[i + 1] = p1.itf; }
                if (o.phase) { o.phase[i] = p0.ph; o.phase[i + 1] = p1.ph; }
                if (o.diff) { o.diff[i] = p0.df; o.diff[i + 1] = p1.df; }
                if (o.dpca_mag) { o.dpca_mag[i] = p0.dmag; o.dpca_mag[i + 1] = p1.dmag; }
                if (o.slc1_mag) { o.slc1_mag[i] = p0.mag1; o.slc1_mag[i + 1] = p1.mag1; }
                if (o.mask) { o.mask[i] = d0 ? 1 : 0; o.mask[i + 1] = d1 ? 1 : 0; }
                if (o.phase_masked) { o.phase_masked[i] = p0.phm; o.phase_masked[i + 1] = p1.phm; }
            }
        } else if (i < n) {   // odd tail pixel
            const PixelOut p0 = gmti_pixel(__ldg(slc1 + i), __ldg(slc2 + i), cal, use_cal, max_sq, thr, lo_sq, hi_sq,
                                           (uint32_t)i, res);
            d0 = p0.det;
            if (o.interf) o.interf[i] = p0.itf;
            if (o.phase) o.phase[i] = p0.ph;
            if (o.diff) o.diff[i] = p0.df;
            if (o.dpca_mag) o.dpca_mag[i] = p0.dmag;
            if (o.slc1_mag) o.slc1_mag[i] = p0.mag1;
            if (o.mask) o.mask[i] = d0 ? 1 : 0;
            if (o.phase_masked) o.phase_masked[i] = p0.phm;
        }
        const uint32_t b0 = __ballot_sync(0xffffffffu, d0), b1 = __ballot_sync(0xffffffffu, d1);
        if (lane == 0) {
            s_cnt[it * 8 + wid] = __popc(b0) + __popc(b1);
            s_bal[it * 8 + wid] = make_uint2(b0, b1);
        }
    }
    __syncthreads();
    if (wid == 0) {
        // exclusive scan of the 32 group counts (pixel order: iteration, warp)
        const uint32_t c = s_cnt[lane];
        uint32_t x = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, x, 31);
        s_cnt[lane] = x - c;
        // decoupled look-back: detections in all earlier tiles
        uint32_t before = 0;
        if (tile > 0) {
            if (lane == 0) st_status(status + tile, kFlagAggregate | total);
            int base = (int)tile - 1;
            while (true) {
                const int idx = base - lane;       // lane 0 looks at the nearest predecessor
                unsigned long long w = kFlagPrefix;   // before tile 0: an inclusive prefix of zero
                if (idx >= 0) {
                    do { w = ld_status(status + idx); } while ((w & kFlagMask) == 0ull);
                }
                const unsigned pref = __ballot_sync(0xffffffffu, (w & kFlagMask) == kFlagPrefix);
                const uint32_t v = (uint32_t)(w & 0xFFFFFFFFull);
                // sum the aggregates of the lanes nearer than the first inclusive prefix, plus that prefix
                const int first = pref ? __ffs(pref) - 1 : 32;
                uint32_t part = (lane <= first) ? v : 0u;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                before += part;
                if (pref) break;
                base -= 32;
            }
        }
        if (lane == 0) {
            st_status(status + tile, kFlagPrefix | (unsigned long long)(before + total));
            s_base = before;
            if (tile == (uint32_t)n_tiles - 1u) res->det_count = before + total;
        }
    }
    __syncthreads();
    if (det_idx != nullptr) {
        const uint32_t below = (1u << lane) - 1u;
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            const uint2 bb = s_bal[it * 8 + wid];
            const uint32_t b0 = bb.x, b1 = bb.y;
            if (((b0 | b1) >> lane) & 1u) {
                const uint64_t i = tile_base + (uint64_t)it * 512 + wid * 64 + 2 * lane;
                uint32_t pos = s_base + s_cnt[it * 8 + wid] + __popc(b0 & below) + __popc(b1 & below);
                if ((b0 >> lane) & 1u) {
                    if (pos < det_cap) det_idx[pos] = (uint32_t)i;
                    ++pos;
                }
                if (((b1 >> lane) & 1u) && pos < det_cap) det_idx[pos] = (uint32_t)i + 1u;
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_balance_sum(const float2* __restrict__ s1, const float2* __restrict__ s2,
                                                     uint64_t n, double* __restrict__ out) {
    double re = 0, im = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float2 a = __ldg(s1 + i), b = __ldg(s2 + i);
        re += (double)a.x * b.x + (double)a.y * b.y;
        im += (double)a.y * b.x - (double)a.x * b.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, o);
        im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    __shared__ double wr[8], wi[8];
    if ((threadIdx.x & 31) == 0) { wr[threadIdx.x >> 5] = re; wi[threadIdx.x >> 5] = im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { re += wr[w]; im += wi[w]; }
        atomicAdd(out, re);
        atomicAdd(out + 1, im);
    }
}

}  // namespace

extern "C" uint64_t nis_gmti_workspace_bytes(uint64_t n_pix) {
    return kWsHeader + ((n_pix + kTile - 1) / kTile) * sizeof(unsigned long long);
}

extern "C" int nis_gmti_fused(nis_ctx* ctx, const nis_c32* slc1, const nis_c32* slc2, uint64_t n_pix,
                              double thresh_frac, double cal_phase, nis_c32* ati_interf, float* ati_phase,
                              nis_c32* dpca_diff, float* dpca_mag, float* slc1_mag, uint8_t* mag_mask,
                              float* ati_phase_masked, uint32_t* det_idx, uint32_t det_cap,
                              const double* max_mag_sq_in, void* workspace, uint64_t workspace_bytes,
                              nis_gmti_result* result, nis_stream stream) {
    NIS_REQUIRE(ctx && slc1 && slc2 && result && workspace, "nis_gmti_fused: null argument");
    NIS_REQUIRE(n_pix > 0 && n_pix < 0xFFFFFFFFull, "nis_gmti_fused: n_pix %llu out of range (flat indices are 32-bit)",
                (unsigned long long)n_pix);
    NIS_REQUIRE(workspace_bytes >= nis_gmti_workspace_bytes(n_pix) && ((uintptr_t)workspace & 7) == 0,
                "nis_gmti_fused: workspace of %llu bytes (8-byte aligned) required, got %llu",
                (unsigned long long)nis_gmti_workspace_bytes(n_pix), (unsigned long long)workspace_bytes);
    NIS_REQUIRE(((uintptr_t)slc1 & 7) == 0 && ((uintptr_t)slc2 & 7) == 0 && ((uintptr_t)ati_interf & 7) == 0 &&
                    ((uintptr_t)dpca_diff & 7) == 0 && ((uintptr_t)ati_phase & 3) == 0 && ((uintptr_t)dpca_mag & 3) == 0 &&
                    ((uintptr_t)slc1_mag & 3) == 0 && ((uintptr_t)ati_phase_masked & 3) == 0 && ((uintptr_t)det_idx & 3) == 0 &&
                    ((uintptr_t)result & 7) == 0,
                "nis_gmti_fused: a buffer is not aligned to its element size");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_tiles = (int)((n_pix + kTile - 1) / kTile);
    uint32_t* ticket = reinterpret_cast<uint32_t*>(workspace);
    unsigned long long* status = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(workspace) + kWsHeader);

    k_gmti_init<<<(n_tiles + 1023) / 1024, 1024, 0, st>>>(result, max_mag_sq_in, ticket, status, n_tiles);
    NIS_LAUNCH_CHECK(ctx);
    if (max_mag_sq_in == nullptr) {   // otherwise the producer of slc1 (nis_csa_focus) already reduced it
        const int max_grid = ctx->num_sms * 8;
        const int g1 = (int)((n_pix + 255) / 256) < max_grid ? (int)((n_pix + 255) / 256) : max_grid;
        k_gmti_max<<<g1, 256, 0, st>>>(reinterpret_cast<const float2*>(slc1), n_pix, result);
        NIS_LAUNCH_CHECK(ctx);
    }
    GmtiOut o{reinterpret_cast<float2*>(ati_interf), ati_phase, reinterpret_cast<float2*>(dpca_diff), dpca_mag,
              slc1_mag, mag_mask, ati_phase_masked};
    const float2 cal = make_float2((float)cos(cal_phase), (float)sin(cal_phase));
    // two pixels per thread move as 16-byte (complex), 8-byte (float) and 2-byte (mask) vectors when every buffer allows it
    const bool vec = (((uintptr_t)slc1 | (uintptr_t)slc2 | (uintptr_t)ati_interf | (uintptr_t)dpca_diff) & 15) == 0 &&
                     (((uintptr_t)ati_phase | (uintptr_t)dpca_mag | (uintptr_t)slc1_mag | (uintptr_t)ati_phase_masked) & 7) == 0 &&
                     ((uintptr_t)mag_mask & 1) == 0;
    if (det_cap == 0) det_idx = nullptr;
    if (vec)
        k_gmti_fused<true><<<n_tiles, 256, 0, st>>>(reinterpret_cast<const float2*>(slc1), reinterpret_cast<const float2*>(slc2),
                                                    n_pix, thresh_frac, cal, cal_phase != 0.0 ? 1 : 0, o, ticket, status,
                                                    n_tiles, det_idx, det_cap, result);
    else
        k_gmti_fused<false><<<n_tiles, 256, 0, st>>>(reinterpret_cast<const float2*>(slc1), reinterpret_cast<const float2*>(slc2),
                                                     n_pix, thresh_frac, cal, cal_phase != 0.0 ? 1 : 0, o, ticket, status,
                                                     n_tiles, det_idx, det_cap, result);
    NIS_LAUNCH_CHECK(ctx);
    return NIS_OK;
}

extern "C" int nis_gmti_balance_sum(nis_ctx* ctx, const nis_c32* slc1, const nis_c32* slc2, uint64_t n_pix,
                                    double* out_sum, nis_stream stream) {
    NIS_REQUIRE(ctx && slc1 && slc2 && out_sum, "nis_gmti_balance_sum: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    NIS_CUDA_TRY(cudaMemsetAsync(out_sum, 0, 2 * sizeof(double), st));
    const int max_grid = ctx->num_sms * 8;
    const int g = (int)((n_pix + 255) / 256) < max_grid ? (int)((n_pix + 255) / 256) : max_grid;
    k_balance_sum<<<g, 256, 0, st>>>(reinterpret_cast<const float2*>(slc1), reinterpret_cast<const float2*>(slc2), n_pix,
                                     out_sum);
    NIS_LAUNCH_CHECK(ctx);
    return NIS_OK;
}
