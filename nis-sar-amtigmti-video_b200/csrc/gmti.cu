// gmti.cu -- K3: fused DPCA subtraction + ATI conjugate multiply + phase + threshold + compaction.
//
// Replaces the seven numpy passes of sar_ati_dcpa_sim_csa.py:414-419, :447-449 (and
// SARData.compute_all, sar_ati_dcpa_viewer_csa.py:42-52) by
//   pass 1  k_gmti_max      read slc1                      -> max |slc1|^2 (fp64, exact on fp32 samples);
//                           skipped when nis_csa_focus already produced it while writing slc1
//   pass 2  k_gmti_products read slc1, slc2, write products -> + detection bitmap, per-tile counts, peak
//   pass 3  k_gmti_scan     per-tile counts -> exclusive offsets, total
//   pass 4  k_gmti_compact  bitmap -> ascending flat indices (== np.flatnonzero(mag_mask))
// The detection test |slc1| > frac * max|slc1| is the reference's strict '>' (:447) evaluated in fp64
// on the fp32 samples, so the index list is bit-exact against numpy given the same SLC.
#include <math.h>

#include "common.cuh"

using namespace nis;

namespace {

constexpr int kTile = 2048;     // pixels per CTA (256 threads x 8)
constexpr int kWordsPerTile = kTile / 32;

struct GmtiOut {
    float2* interf; float* phase; float2* diff; float* dpca_mag; float* slc1_mag;
    uint8_t* mask; float* phase_masked;
};

__global__ void k_gmti_init(nis_gmti_result* res, const double* __restrict__ max_sq_in) {
    res->det_count = 0;
    res->peak_idx = 0xFFFFFFFFu;
    res->max_mag_sq = max_sq_in ? *max_sq_in : 0.0;
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

// Two adjacent pixels per thread (16-byte loads and stores); a warp covers 64 consecutive pixels per iteration
// and emits two detection-bitmap words.
__global__ void __launch_bounds__(256) k_gmti_products(const float2* __restrict__ slc1, const float2* __restrict__ slc2,
                                                       uint64_t n, double thresh_frac, float2 cal, int use_cal, GmtiOut o,
                                                       uint32_t* __restrict__ bitmap, uint32_t* __restrict__ tile_count,
                                                       nis_gmti_result* res) {
    const double max_sq = res->max_mag_sq;
    // np.max(np.abs(slc1)) * frac, compared against np.abs(slc1): both sides are sqrt of the fp64 |.|^2
    const double thr = sqrt(max_sq) * thresh_frac;
    const double thr_sq = thr * thr;
    const double lo_sq = thr_sq * (1.0 - 1e-12), hi_sq = thr_sq * (1.0 + 1e-12);
    const uint64_t tile_base = (uint64_t)blockIdx.x * kTile;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int cnt = 0;
#pragma unroll 2
    for (int it = 0; it < kTile / 512; ++it) {
        const uint64_t i = tile_base + (uint64_t)it * 512 + wid * 64 + 2 * lane;
        bool d0 = false, d1 = false;
        if (i + 1 < n) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(slc1 + i));
            const float4 b = __ldg(reinterpret_cast<const float4*>(slc2 + i));
            const PixelOut p0 = gmti_pixel(make_float2(a.x, a.y), make_float2(b.x, b.y), cal, use_cal, max_sq, thr, lo_sq,
                                           hi_sq, (uint32_t)i, res);
            const PixelOut p1 = gmti_pixel(make_float2(a.z, a.w), make_float2(b.z, b.w), cal, use_cal, max_sq, thr, lo_sq,
                                           hi_sq, (uint32_t)i + 1u, res);
            d0 = p0.det; d1 = p1.det;
            if (o.interf) *reinterpret_cast<float4*>(o.interf + i) = make_float4(p0.itf.x, p0.itf.y, p1.itf.x, p1.itf.y);
            if (o.phase) *reinterpret_cast<float2*>(o.phase + i) = make_float2(p0.ph, p1.ph);
            if (o.diff) *reinterpret_cast<float4*>(o.diff + i) = make_float4(p0.df.x, p0.df.y, p1.df.x, p1.df.y);
            if (o.dpca_mag) *reinterpret_cast<float2*>(o.dpca_mag + i) = make_float2(p0.dmag, p1.dmag);
            if (o.slc1_mag) *reinterpret_cast<float2*>(o.slc1_mag + i) = make_float2(p0.mag1, p1.mag1);
            if (o.mask) *reinterpret_cast<uchar2*>(o.mask + i) = make_uchar2(d0 ? 1 : 0, d1 ? 1 : 0);
            if (o.phase_masked) *reinterpret_cast<float2*>(o.phase_masked + i) = make_float2(p0.phm, p1.phm);
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
        const unsigned b0 = __ballot_sync(0xffffffffu, d0), b1 = __ballot_sync(0xffffffffu, d1);
        if (lane == 0) {
            uint32_t* w = bitmap + (size_t)blockIdx.x * kWordsPerTile + it * 16 + wid * 2;
            w[0] = interleave16(b0, b1);
            w[1] = interleave16(b0 >> 16, b1 >> 16);
            cnt += __popc(b0) + __popc(b1);
        }
    }
    __shared__ int wc[8];
    if (lane == 0) wc[wid] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += wc[w];
        tile_count[blockIdx.x] = (uint32_t)t;
    }
}

// exclusive scan of n_tiles counts by a single CTA (n_tiles <= a few 10^4)
__global__ void __launch_bounds__(1024) k_gmti_scan(const uint32_t* __restrict__ cnt, uint32_t* __restrict__ off,
                                                    int n_tiles, nis_gmti_result* res) {
    constexpr int PER = 8;   // counts per thread per round: 8192 tiles (a 4096^2 frame) in one round
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < n_tiles; base += 1024 * PER) {
        const int i0 = base + threadIdx.x * PER;
        uint32_t v[PER], local = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            v[j] = (i0 + j < n_tiles) ? cnt[i0 + j] : 0u;
            local += v[j];
        }
        uint32_t x = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sum[lane] = w;  // inclusive
        }
        __syncthreads();
        uint32_t run = carry + (wid > 0 ? warp_sum[wid - 1] : 0u) + (x - local);
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (i0 + j < n_tiles) off[i0 + j] = run;
            run += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = run;
        __syncthreads();
    }
    if (threadIdx.x == 0) res->det_count = carry;
}

__global__ void __launch_bounds__(256) k_gmti_compact(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ off,
                                                      uint32_t* __restrict__ det_idx, uint32_t det_cap) {
    __shared__ uint32_t word_off[kWordsPerTile];
    const uint32_t* words = bitmap + (size_t)blockIdx.x * kWordsPerTile;
    if (threadIdx.x < 32) {
        // 64 words: two per lane, exclusive prefix of popcounts
        const uint32_t c0 = __popc(words[2 * threadIdx.x]), c1 = __popc(words[2 * threadIdx.x + 1]);
        uint32_t x = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((int)threadIdx.x >= o) x += y;
        }
        const uint32_t excl = x - (c0 + c1);
        word_off[2 * threadIdx.x] = excl;
        word_off[2 * threadIdx.x + 1] = excl + c0;
    }
    __syncthreads();
    const uint32_t tile_off = off[blockIdx.x];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int it = 0; it < kTile / 256; ++it) {
        const int w = it * 8 + wid;
        const uint32_t bits = words[w];
        if ((bits >> lane) & 1u) {
            const uint32_t pos = tile_off + word_off[w] + __popc(bits & ((1u << lane) - 1u));
            if (pos < det_cap) det_idx[pos] = (uint32_t)((uint64_t)blockIdx.x * kTile + it * 256 + threadIdx.x);
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

extern "C" int nis_gmti_fused(nis_ctx* ctx, const nis_c32* slc1, const nis_c32* slc2, uint64_t n_pix,
                              double thresh_frac, double cal_phase, nis_c32* ati_interf, float* ati_phase,
                              nis_c32* dpca_diff, float* dpca_mag, float* slc1_mag, uint8_t* mag_mask,
                              float* ati_phase_masked, uint32_t* det_idx, uint32_t det_cap,
                              const double* max_mag_sq_in, nis_gmti_result* result, nis_stream stream) {
    NIS_REQUIRE(ctx && slc1 && slc2 && result, "nis_gmti_fused: null argument");
    NIS_REQUIRE(n_pix > 0 && n_pix < 0xFFFFFFFFull, "nis_gmti_fused: n_pix %llu out of range", (unsigned long long)n_pix);
    cudaStream_t st = (cudaStream_t)stream;
    const int n_tiles = (int)((n_pix + kTile - 1) / kTile);
    const size_t bitmap_bytes = (size_t)n_tiles * kWordsPerTile * sizeof(uint32_t);
    const size_t need = bitmap_bytes + 2 * (size_t)n_tiles * sizeof(uint32_t);
    int rc = ctx->ensure_scratch(need);
    if (rc != NIS_OK) return rc;
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(ctx->scratch);
    uint32_t* tile_count = bitmap + (size_t)n_tiles * kWordsPerTile;
    uint32_t* tile_off = tile_count + n_tiles;

    k_gmti_init<<<1, 1, 0, st>>>(result, max_mag_sq_in);
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
    k_gmti_products<<<n_tiles, 256, 0, st>>>(reinterpret_cast<const float2*>(slc1), reinterpret_cast<const float2*>(slc2),
                                             n_pix, thresh_frac, cal, cal_phase != 0.0 ? 1 : 0, o, bitmap, tile_count,
                                             result);
    NIS_LAUNCH_CHECK(ctx);
    k_gmti_scan<<<1, 1024, 0, st>>>(tile_count, tile_off, n_tiles, result);
    NIS_LAUNCH_CHECK(ctx);
    if (det_idx != nullptr && det_cap > 0) {
        k_gmti_compact<<<n_tiles, 256, 0, st>>>(bitmap, tile_off, det_idx, det_cap);
        NIS_LAUNCH_CHECK(ctx);
    }
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
