// viewer.cu -- the data layer of the interactive ATI/DPCA viewer on the device (sar_ati_dcpa_viewer_csa.py; SURVEY.md
// section 8f row N3): SARData.compute_all's seven product maps (:42-52) in one pass, and the per-view statistics the
// viewer recomputes on every zoom / mode change -- mean, std, min, max (:117-139), median (:118, :136) and the 99.9th
// percentile that sets the colour limits (:147, :178, :184) -- on a rectangular region of a map, without sorting:
//   k_viewer_products  read 16 B, write up to 28 B per pixel (HBM bound)
//   k_region_moments   two passes: (sum, min, max), then sum of squared deviations about the mean read from the device
//   k_region_hist x 3  exact order statistics by radix select on the order-preserving 32-bit key of a float
//                      (11 + 11 + 10 bits): each pass histograms the digits of the values that match the prefix found so
//                      far, k_select_digit picks the digit that holds the wanted rank -- no host round trip in between.
// dB display (20 log10(x + 1e-12), :128, :176) is monotonic, so order statistics are selected on the linear values and
// transformed afterwards; the moments transform every sample (fp64 log10).
#include <math.h>

#include "common.cuh"

using namespace nis;

namespace {

__global__ void __launch_bounds__(256) k_viewer_products(const float2* __restrict__ s1, const float2* __restrict__ s2,
                                                         uint64_t n, float2 cal, int use_cal, float* __restrict__ ch1_mag,
                                                         float* __restrict__ ch1_phase, float* __restrict__ ch2_mag,
                                                         float* __restrict__ ch2_phase, float* __restrict__ dpca_mag,
                                                         float* __restrict__ dpca_phase, float* __restrict__ ati_phase) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float2 a = s1[i];
        float2 b = s2[i];
        if (use_cal) b = cmul(b, cal);                       // s2 * exp(j cal_phase)  (:43)
        const float2 d = make_float2(a.x - b.x, a.y - b.y);
        const float2 x = cmul_conj(a, b);                    // s1 conj(s2_cal)         (:51)
        if (ch1_mag) ch1_mag[i] = hypotf(a.x, a.y);
        if (ch1_phase) ch1_phase[i] = atan2f(a.y, a.x);
        if (ch2_mag) ch2_mag[i] = hypotf(b.x, b.y);
        if (ch2_phase) ch2_phase[i] = atan2f(b.y, b.x);
        if (dpca_mag) dpca_mag[i] = hypotf(d.x, d.y);
        if (dpca_phase) dpca_phase[i] = atan2f(d.y, d.x);
        if (ati_phase) ati_phase[i] = atan2f(x.y, x.x);
    }
}

__device__ __forceinline__ double disp_value(float v, int db) {
    return db ? 20.0 * log10((double)v + 1e-12) : (double)v;
}
__device__ __forceinline__ uint32_t float_key(float v) {     // order-preserving: a < b  <=>  key(a) < key(b)
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ void atomic_min_f64(double* dst, double v) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(dst);
    unsigned long long old = *p;
    while (__longlong_as_double((long long)old) > v) {
        const unsigned long long prev = atomicCAS(p, old, (unsigned long long)__double_as_longlong(v));
        if (prev == old) break;
        old = prev;
    }
}
__device__ __forceinline__ void atomic_max_f64_any(double* dst, double v) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(dst);
    unsigned long long old = *p;
    while (__longlong_as_double((long long)old) < v) {
        const unsigned long long prev = atomicCAS(p, old, (unsigned long long)__double_as_longlong(v));
        if (prev == old) break;
        old = prev;
    }
}

// acc: [0] sum, [1] min, [2] max, [3] sum of squared deviations.  PASS 0 fills 0..2, PASS 1 fills 3 using mean = acc[0]/n.
template <int PASS>
__global__ void __launch_bounds__(256) k_region_moments(const float* __restrict__ map, int64_t pitch, int rows, int cols,
                                                        int db, double* __restrict__ acc) {
    const uint64_t n = (uint64_t)rows * cols;
    const double mean = PASS ? acc[0] / (double)n : 0.0;
    double s = 0.0, mn = INFINITY, mx = -INFINITY;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const double v = disp_value(map[(i / cols) * pitch + (i % cols)], db);
        if (PASS == 0) {
            s += v;
            mn = fmin(mn, v);
            mx = fmax(mx, v);
        } else {
            s += (v - mean) * (v - mean);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        if (PASS == 0) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&acc[PASS ? 3 : 0], s);
        if (PASS == 0) {
            atomic_min_f64(&acc[1], mn);
            atomic_max_f64_any(&acc[2], mx);
        }
    }
}

constexpr int kMaxRanks = 4;
struct SelectState {                // one per wanted rank
    uint32_t prefix, mask;          // key & mask == prefix for every candidate still in play
    unsigned long long remaining;   // rank among those candidates
};

// digit = (key >> shift) & (2^bits - 1), counted for every rank whose prefix the key matches
__global__ void __launch_bounds__(256) k_region_hist(const float* __restrict__ map, int64_t pitch, int rows, int cols,
                                                     int shift, int bits, int n_ranks,
                                                     const SelectState* __restrict__ st,
                                                     unsigned long long* __restrict__ hist /* [n_ranks][2048] */) {
    __shared__ unsigned int h[kMaxRanks][2048];
    for (int i = threadIdx.x; i < kMaxRanks * 2048; i += blockDim.x) (&h[0][0])[i] = 0;
    SelectState s[kMaxRanks];
    for (int r = 0; r < n_ranks; ++r) s[r] = st[r];
    __syncthreads();
    const uint64_t n = (uint64_t)rows * cols;
    const uint32_t dmask = (1u << bits) - 1u;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t k = float_key(map[(i / cols) * pitch + (i % cols)]);
        for (int r = 0; r < n_ranks; ++r)
            if ((k & s[r].mask) == s[r].prefix) atomicAdd(&h[r][(k >> shift) & dmask], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_ranks * 2048; i += blockDim.x) {
        const unsigned int c = (&h[0][0])[i];
        if (c) atomicAdd(&hist[i], (unsigned long long)c);
    }
}

// one CTA: for each rank, find the digit whose cumulative count passes `remaining`; extend the prefix; clear the histogram.
// After the last pass (shift == 0) the prefix is the key of the wanted order statistic: write its value.
__global__ void __launch_bounds__(256) k_select_digit(int shift, int bits, int n_ranks, SelectState* __restrict__ st,
                                                      unsigned long long* __restrict__ hist, float* __restrict__ out) {
    const int nd = 1 << bits;
    for (int r = 0; r < n_ranks; ++r) {
        if (threadIdx.x == 0) {
            unsigned long long rem = st[r].remaining, run = 0;
            int d = 0;
            for (; d < nd - 1; ++d) {
                const unsigned long long c = hist[r * 2048 + d];
                if (run + c > rem) break;
                run += c;
            }
            st[r].prefix |= (uint32_t)d << shift;
            st[r].mask |= ((1u << bits) - 1u) << shift;
            st[r].remaining = rem - run;
            if (shift == 0) out[r] = key_float(st[r].prefix);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2048; i += blockDim.x) hist[r * 2048 + i] = 0;
        __syncthreads();
    }
}

unsigned grid_for(const nis_ctx* ctx, uint64_t n, int per_thread) {
    uint64_t blocks = (n + 256ull * per_thread - 1) / (256ull * per_thread);
    const uint64_t cap = (uint64_t)ctx->num_sms * 8;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks ? blocks : 1);
}

}  // namespace

extern "C" int nis_viewer_products(nis_ctx* ctx, const nis_c32* slc1, const nis_c32* slc2, uint64_t n_pix, double cal_phase,
                                   float* ch1_mag, float* ch1_phase, float* ch2_mag, float* ch2_phase, float* dpca_mag,
                                   float* dpca_phase, float* ati_phase, nis_stream stream) {
    NIS_REQUIRE(ctx && slc1 && slc2, "nis_viewer_products: null argument");
    if (n_pix == 0) return NIS_OK;
    const float2 cal = make_float2((float)cos(cal_phase), (float)sin(cal_phase));
    k_viewer_products<<<grid_for(ctx, n_pix, 4), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(slc1), reinterpret_cast<const float2*>(slc2), n_pix, cal, cal_phase != 0.0 ? 1 : 0,
        ch1_mag, ch1_phase, ch2_mag, ch2_phase, dpca_mag, dpca_phase, ati_phase);
    NIS_LAUNCH_CHECK(ctx);
    return NIS_OK;
}

extern "C" int nis_region_stats(nis_ctx* ctx, const float* map, int64_t pitch, int32_t rows, int32_t cols, int32_t db_scale,
                                double* out_dev, nis_stream stream) {
    NIS_REQUIRE(ctx && map && out_dev, "nis_region_stats: null argument");
    NIS_REQUIRE(rows > 0 && cols > 0 && pitch >= cols, "nis_region_stats: empty region or pitch < cols");
    cudaStream_t st = (cudaStream_t)stream;
    const double init[4] = {0.0, INFINITY, -INFINITY, 0.0};
    NIS_CUDA_TRY(cudaMemcpyAsync(out_dev, init, sizeof(init), cudaMemcpyHostToDevice, st));
    const uint64_t n = (uint64_t)rows * cols;
    k_region_moments<0><<<grid_for(ctx, n, 8), 256, 0, st>>>(map, pitch, rows, cols, db_scale, out_dev);
    NIS_LAUNCH_CHECK(ctx);
    k_region_moments<1><<<grid_for(ctx, n, 8), 256, 0, st>>>(map, pitch, rows, cols, db_scale, out_dev);
    NIS_LAUNCH_CHECK(ctx);
    return NIS_OK;
}

extern "C" int nis_region_select(nis_ctx* ctx, const float* map, int64_t pitch, int32_t rows, int32_t cols,
                                 const uint64_t* ranks, int32_t n_ranks, float* out_dev, nis_stream stream) {
    NIS_REQUIRE(ctx && map && ranks && out_dev, "nis_region_select: null argument");
    NIS_REQUIRE(rows > 0 && cols > 0 && pitch >= cols, "nis_region_select: empty region or pitch < cols");
    NIS_REQUIRE(n_ranks >= 1 && n_ranks <= kMaxRanks, "nis_region_select: 1..%d ranks per call", kMaxRanks);
    const uint64_t n = (uint64_t)rows * cols;
    SelectState init[kMaxRanks] = {};
    for (int r = 0; r < n_ranks; ++r) {
        NIS_REQUIRE(ranks[r] < n, "nis_region_select: rank %llu outside a region of %llu values",
                    (unsigned long long)ranks[r], (unsigned long long)n);
        init[r].remaining = ranks[r];
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t hist_bytes = (size_t)kMaxRanks * 2048 * sizeof(unsigned long long);
    void* ws = nullptr;
    int rc = ctx->stream_scratch(st, hist_bytes + sizeof(init), &ws);   // per stream, never freed under a graph
    if (rc != NIS_OK) return rc;
    unsigned long long* hist = reinterpret_cast<unsigned long long*>(ws);
    SelectState* state = reinterpret_cast<SelectState*>(reinterpret_cast<char*>(ws) + hist_bytes);
    NIS_CUDA_TRY(cudaMemsetAsync(hist, 0, hist_bytes, st));
    NIS_CUDA_TRY(cudaMemcpyAsync(state, init, sizeof(init), cudaMemcpyHostToDevice, st));
    const int shifts[3] = {21, 10, 0}, bits[3] = {11, 11, 10};
    for (int p = 0; p < 3; ++p) {
        k_region_hist<<<grid_for(ctx, n, 16), 256, 0, st>>>(map, pitch, rows, cols, shifts[p], bits[p], n_ranks, state, hist);
        NIS_LAUNCH_CHECK(ctx);
        k_select_digit<<<1, 256, 0, st>>>(shifts[p], bits[p], n_ranks, state, hist, out_dev);
        NIS_LAUNCH_CHECK(ctx);
    }
    return NIS_OK;
}
