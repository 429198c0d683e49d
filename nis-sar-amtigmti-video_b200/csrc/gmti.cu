// gmti.cu -- K3: fused DPCA subtraction + ATI conjugate multiply + phase + threshold + compaction.
//
// Replaces the seven numpy passes of sar_ati_dcpa_sim_csa.py:414-419, :447-449 (and
// SARData.compute_all, sar_ati_dcpa_viewer_csa.py:42-52) by ONE pass over the channel pair plus a small tail kernel:
//   [k_gmti_max     read slc1 -> max |slc1|^2 (fp64, exact on fp32 samples); only when the caller does not hand in the value
//                   nis_csa_focus produced while writing slc1 -- the normal case on the focusing path skips it]
//   k_gmti_products read slc1, slc2 (16 B), write every requested product (<= 33 B); per 2048-pixel tile: detection
//                   bitmap (256 B), detection count, first arg-max candidate -- plain stores into the caller's workspace
//   k_gmti_compact  reads only the workspace (264 B per tile): every CTA sums the counts of the tiles before its group
//                   (L2-resident, a few KB), scans its own tiles and writes their set bits as ascending flat indices
//                   (== np.flatnonzero(mag_mask)); the last CTA publishes det_count / peak_idx / max.
// No atomics, no inter-CTA waiting, nothing to zero beforehand: the list is deterministic and the kernels never spin.
// (A single-kernel variant with a decoupled look-back over per-tile status words was built first and measured SLOWER --
// 0.248 vs 0.146 ms for the products of a 4096^2 pair: 2048-pixel CTAs live ~11 us, and the ticket + look-back round trips
// at their tail leave each one idle for a fifth of that.)
// The detection test |slc1| > frac * max|slc1| is the reference's strict '>' (:447) evaluated in fp64
// on the fp32 samples, so the index list is bit-exact against numpy given the same SLC.
#include <math.h>

#include "common.cuh"

using namespace nis;

namespace {

constexpr int kTile = 2048;              // pixels per CTA of the products kernel (256 threads x 8)
constexpr int kIters = kTile / 512;      // a warp covers 64 consecutive pixels per iteration
constexpr int kWordsPerTile = kTile / 32;
constexpr size_t kWsHeader = 16;
constexpr int kCompactCtas = 592;        // tail kernel: about this many CTAs, each compacting a group of consecutive tiles

struct GmtiOut {
    float2* interf; float* phase; float2* diff; float* dpca_mag; float* slc1_mag;
    uint8_t* mask; float* phase_masked;
};

// workspace: [bitmap: n_tiles x 64 words][count: n_tiles][peak: n_tiles]
struct GmtiWs {
    uint32_t* bitmap;
    uint32_t* count;
    uint32_t* peak;
};
__host__ __device__ inline GmtiWs ws_layout(void* base, int n_tiles) {
    GmtiWs w;
    w.bitmap = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(base) + kWsHeader);
    w.count = w.bitmap + (size_t)n_tiles * kWordsPerTile;
    w.peak = w.count + n_tiles;
    return w;
}

__device__ __forceinline__ double sq_mag(float2 s) {
    const double re = (double)s.x, im = (double)s.y;
    return fma(re, re, im * im);  // both products are exact in fp64; one rounding
}

__global__ void __launch_bounds__(256) k_gmti_max(const float2* __restrict__ slc1, uint64_t n, double* __restrict__ max_sq) {
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
        atomicMax(reinterpret_cast<unsigned long long*>(max_sq), (unsigned long long)__double_as_longlong(m));
    }
}

__device__ __forceinline__ uint32_t interleave16(uint32_t even, uint32_t odd) {
    // bit i of `even` -> bit 2i, bit i of `odd` -> bit 2i+1 (16-bit inputs)
    auto spread = [](uint32_t x) {
        x &= 0xFFFFu;
        x = (x | (x << 8)) & 0x00FF00FFu;
        x = (x | (x << 4)) & 0x0F0F0F0Fu;
        x = (x | (x << 2)) & 0x33333333u;
        x = (x | (x << 1)) & 0x55555555u;
        return x;
    };
    return spread(even) | (spread(odd) << 1);
}

struct PixelOut {
    float2 itf, df;
    float ph, dmag, mag1, phm;
    bool det;
};

__device__ __forceinline__ PixelOut gmti_pixel(float2 s1, float2 s2, float2 cal, int use_cal, double max_sq, double thr,
                                               double lo_sq, double hi_sq, uint32_t idx, uint32_t& peak) {
    PixelOut o;
    if (use_cal) s2 = cmul(s2, cal);
    const double sq = sq_mag(s1);
    o.det = sq > hi_sq ? true : (sq < lo_sq ? false : (sqrt(sq) > thr));
    if (sq == max_sq) peak = min(peak, idx);
    o.itf = cmul_conj(s1, s2);
    o.ph = atan2f(o.itf.y, o.itf.x);
    o.df = csub(s1, s2);
    o.dmag = sqrtf(fmaf(o.df.x, o.df.x, o.df.y * o.df.y));
    o.mag1 = sqrtf(fmaf(s1.x, s1.x, s1.y * s1.y));
    o.phm = o.det ? o.ph : 0.f;
    return o;
}

// Two adjacent pixels per thread; a warp covers 64 consecutive pixels per iteration and emits two detection-bitmap words.
// VEC: 16-byte loads and 16- / 8- / 2-byte stores (every pointer suitably aligned); otherwise element-wise accesses
// (views at odd element offsets).
template <bool VEC>
__global__ void __launch_bounds__(256) k_gmti_products(const float2* __restrict__ slc1, const float2* __restrict__ slc2,
                                                       uint64_t n, double thresh_frac, float2 cal, int use_cal, GmtiOut o,
                                                       const double* __restrict__ max_sq_ptr, GmtiWs ws) {
    const double max_sq = *max_sq_ptr;
    // np.max(np.abs(slc1)) * frac, compared against np.abs(slc1): both sides are sqrt of the fp64 |.|^2
    const double thr = sqrt(max_sq) * thresh_frac;
    const double thr_sq = thr * thr;
    const double lo_sq = thr_sq * (1.0 - 1e-12), hi_sq = thr_sq * (1.0 + 1e-12);
    const uint64_t tile_base = (uint64_t)blockIdx.x * kTile;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int cnt = 0;
    uint32_t peak = 0xFFFFFFFFu;
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
            const PixelOut p0 = gmti_pixel(a0, b0, cal, use_cal, max_sq, thr, lo_sq, hi_sq, (uint32_t)i, peak);
            const PixelOut p1 = gmti_pixel(a1, b1, cal, use_cal, max_sq, thr, lo_sq, hi_sq, (uint32_t)i + 1u, peak);
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
                                           (uint32_t)i, peak);
            d0 = p0.det;
            if (o.interf) o.interf[i] = p0.itf;
            if (o.phase) o.phase[i] = p0.ph;
            if (o.diff) o.diff[i] = p0.df;
            if (o.dpca_mag) o.dpca_mag[i] = p0.dmag;
            if (o.slc1_mag) o.slc1_mag[i] = p0.mag1;
            if (o.mask) o.mask[i] = d0 ? 1 : 0;
            if (o.phase_masked) o.phase_masked[i] = p0.phm;
        }
        const unsigned b0 = __ballot_sync(0xffffffffu, d0), b1 = __ballot_sync(0xffffffffu, d1);
        if (lane == 0) {
            uint32_t* w = ws.bitmap + (size_t)blockIdx.x * kWordsPerTile + it * 16 + wid * 2;
            w[0] = interleave16(b0, b1);
            w[1] = interleave16(b0 >> 16, b1 >> 16);
            cnt += __popc(b0) + __popc(b1);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) peak = min(peak, __shfl_xor_sync(0xffffffffu, peak, d));
    __shared__ int wc[8];
    __shared__ uint32_t wp[8];
    if (lane == 0) { wc[wid] = cnt; wp[wid] = peak; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        uint32_t p = 0xFFFFFFFFu;
        for (int w = 0; w < 8; ++w) { t += wc[w]; p = min(p, wp[w]); }
        ws.count[blockIdx.x] = (uint32_t)t;
        ws.peak[blockIdx.x] = p;
    }
}

// Tail: CTA g compacts tiles [g G, (g+1) G).  Its base offset is the sum of the counts of all earlier tiles, which every CTA
// adds up for itself (n_tiles words, L2-resident) -- no ordering between CTAs.
__global__ void __launch_bounds__(256) k_gmti_compact(GmtiWs ws, int n_tiles, int G, uint32_t* __restrict__ det_idx,
                                                      uint32_t det_cap, const double* __restrict__ max_sq_ptr,
                                                      nis_gmti_result* __restrict__ res) {
    __shared__ uint32_t red[8];
    __shared__ uint32_t word_off[kWordsPerTile];
    __shared__ uint32_t s_run;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int t0 = blockIdx.x * G, t1 = min(t0 + G, n_tiles);
    uint32_t sum = 0;
    for (int i = threadIdx.x; i < t0; i += 256) sum += __ldg(ws.count + i);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (lane == 0) red[wid] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t b = 0;
        for (int w = 0; w < 8; ++w) b += red[w];
        s_run = b;
    }
    __syncthreads();
    for (int tile = t0; tile < t1; ++tile) {
        const uint32_t tile_cnt = __ldg(ws.count + tile);
        const uint32_t tile_off = s_run;
        if (tile_cnt != 0 && det_idx != nullptr && tile_off < det_cap) {     // uniform per CTA
            const uint32_t* words = ws.bitmap + (size_t)tile * kWordsPerTile;
            if (threadIdx.x < 32) {
                // 64 words: two per lane, exclusive prefix of popcounts
                const uint32_t c0 = __popc(words[2 * threadIdx.x]), c1 = __popc(words[2 * threadIdx.x + 1]);
                uint32_t x = c0 + c1;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                    if (lane >= d) x += y;
                }
                const uint32_t excl = x - (c0 + c1);
                word_off[2 * threadIdx.x] = excl;
                word_off[2 * threadIdx.x + 1] = excl + c0;
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < kTile / 256; ++it) {
                const int w = it * 8 + wid;
                const uint32_t bits = words[w];
                if ((bits >> lane) & 1u) {
                    const uint32_t pos = tile_off + word_off[w] + __popc(bits & ((1u << lane) - 1u));
                    if (pos < det_cap) det_idx[pos] = (uint32_t)((uint64_t)tile * kTile + it * 256 + threadIdx.x);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run = tile_off + tile_cnt;
        __syncthreads();
    }
    if (t1 == n_tiles && t0 < n_tiles) {   // the CTA that owns the last tile publishes the record
        uint32_t p = 0xFFFFFFFFu;
        for (int i = threadIdx.x; i < n_tiles; i += 256) p = min(p, __ldg(ws.peak + i));
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) p = min(p, __shfl_xor_sync(0xffffffffu, p, d));
        if (lane == 0) red[wid] = p;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) p = min(p, red[w]);
            res->det_count = s_run;
            res->peak_idx = p;
            res->max_mag_sq = *max_sq_ptr;
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
    const uint64_t n_tiles = (n_pix + kTile - 1) / kTile;
    return kWsHeader + n_tiles * (kWordsPerTile + 2) * sizeof(uint32_t);
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
    const GmtiWs ws = ws_layout(workspace, n_tiles);
    const double* max_ptr = max_mag_sq_in;
    if (max_ptr == nullptr) {   // otherwise the producer of slc1 (nis_csa_focus) already reduced it
        double* own = reinterpret_cast<double*>(workspace);   // header of the workspace
        NIS_CUDA_TRY(cudaMemsetAsync(own, 0, sizeof(double), st));
        const int max_grid = ctx->num_sms * 8;
        const int g1 = (int)((n_pix + 255) / 256) < max_grid ? (int)((n_pix + 255) / 256) : max_grid;
        k_gmti_max<<<g1, 256, 0, st>>>(reinterpret_cast<const float2*>(slc1), n_pix, own);
        NIS_LAUNCH_CHECK(ctx);
        max_ptr = own;
    }
    GmtiOut o{reinterpret_cast<float2*>(ati_interf), ati_phase, reinterpret_cast<float2*>(dpca_diff), dpca_mag,
              slc1_mag, mag_mask, ati_phase_masked};
    const float2 cal = make_float2((float)cos(cal_phase), (float)sin(cal_phase));
    // two pixels per thread move as 16-byte (complex), 8-byte (float) and 2-byte (mask) vectors when every buffer allows it
    const bool vec = (((uintptr_t)slc1 | (uintptr_t)slc2 | (uintptr_t)ati_interf | (uintptr_t)dpca_diff) & 15) == 0 &&
                     (((uintptr_t)ati_phase | (uintptr_t)dpca_mag | (uintptr_t)slc1_mag | (uintptr_t)ati_phase_masked) & 7) == 0 &&
                     ((uintptr_t)mag_mask & 1) == 0;
    if (vec)
        k_gmti_products<true><<<n_tiles, 256, 0, st>>>(reinterpret_cast<const float2*>(slc1), reinterpret_cast<const float2*>(slc2),
                                                       n_pix, thresh_frac, cal, cal_phase != 0.0 ? 1 : 0, o, max_ptr, ws);
    else
        k_gmti_products<false><<<n_tiles, 256, 0, st>>>(reinterpret_cast<const float2*>(slc1), reinterpret_cast<const float2*>(slc2),
                                                        n_pix, thresh_frac, cal, cal_phase != 0.0 ? 1 : 0, o, max_ptr, ws);
    NIS_LAUNCH_CHECK(ctx);
    if (det_cap == 0) det_idx = nullptr;
    const int G = (n_tiles + kCompactCtas - 1) / kCompactCtas;
    k_gmti_compact<<<(n_tiles + G - 1) / G, 256, 0, st>>>(ws, n_tiles, G, det_idx, det_cap, max_ptr, result);
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
