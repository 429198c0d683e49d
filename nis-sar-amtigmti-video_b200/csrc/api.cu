// api.cu -- context, error reporting and buffer-format helpers of libnis_sar.
#include <stdarg.h>
#include <string.h>

#include <exception>

#include "common.cuh"

namespace nis {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace nis

using namespace nis;

int nis_ctx::stream_scratch(cudaStream_t st, size_t bytes, void** out) {
    std::lock_guard<std::mutex> lock(mu);
    Scratch& s = scratch[st];
    if (bytes > s.bytes) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) {
            set_error("workspace of %zu bytes needed while the stream is capturing: run the call once outside the capture "
                      "first (allocation is not captured)", bytes);
            return NIS_ERR_INVALID;
        }
        const size_t want = bytes + (bytes >> 2);
        void* p = nullptr;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            set_error("scratch allocation of %zu bytes failed: %s", want, cudaGetErrorString(cudaGetLastError()));
            return NIS_ERR_NOMEM;
        }
        if (s.ptr) retired.push_back(s.ptr);   // a captured graph may still replay with the old address
        s.ptr = p;
        s.bytes = want;
    }
    *out = s.ptr;
    return NIS_OK;
}

extern "C" int nis_version(void) { return NIS_SAR_ABI_VERSION; }

extern "C" size_t nis_last_error(char* buf, size_t cap) {
    const size_t len = strlen(g_err);
    if (buf && cap > 0) {
        const size_t n = len < cap - 1 ? len : cap - 1;
        memcpy(buf, g_err, n);
        buf[n] = 0;
    }
    return len;
}

extern "C" int nis_ctx_create(int device, nis_ctx** out) {
    NIS_REQUIRE(out != nullptr, "nis_ctx_create: null out pointer");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("nis_ctx_create: no CUDA device (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return NIS_ERR_CUDA;
    }
    NIS_REQUIRE(device >= 0 && device < count, "nis_ctx_create: device %d out of range (0..%d)", device, count - 1);
    cudaDeviceProp prop;
    NIS_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("nis_ctx_create: device %d is sm_%d%d; this build carries sm_100a code only", device, prop.major,
                  prop.minor);
        return NIS_ERR_UNSUPPORTED;
    }
    nis_ctx* ctx = new nis_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    *out = ctx;
    return NIS_OK;
}

extern "C" int nis_ctx_destroy(nis_ctx* ctx) {
    if (!ctx) return NIS_OK;
    {
        DeviceGuard guard(ctx->device);
        for (auto& kv : ctx->scratch)
            if (kv.second.ptr) cudaFree(kv.second.ptr);
        for (void* p : ctx->retired) cudaFree(p);
    }
    delete ctx;
    return NIS_OK;
}

extern "C" uint64_t nis_ctx_launch_count(const nis_ctx* ctx) { return ctx ? ctx->launches : 0; }

// --------------------------------------------------------------------------- format helpers
namespace {

__global__ void __launch_bounds__(256) k_narrow(const double2* __restrict__ src, float2* __restrict__ dst, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const double2 v = src[i];
        dst[i] = make_float2((float)v.x, (float)v.y);
    }
}
__global__ void __launch_bounds__(256) k_widen(const float2* __restrict__ src, double2* __restrict__ dst, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float2 v = src[i];
        dst[i] = make_double2((double)v.x, (double)v.y);
    }
}
// 32 x 32 tiles, 256 threads, padded shared tile: both sides move 256 B row pieces
__global__ void __launch_bounds__(256) k_transpose(const float2* __restrict__ in, int64_t in_pitch,
                                                   float2* __restrict__ out, int rows, int cols) {
    __shared__ float2 tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + 8 * i, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + 8 * i][tx] = in[(int64_t)r * in_pitch + c];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, r = r0 + tx;
        if (r < rows && c < cols) out[(int64_t)c * rows + r] = tile[tx][ty + 8 * i];
    }
}

int grid_for(const nis_ctx* ctx, uint64_t n) {
    const uint64_t want = (n + 255) / 256, cap = (uint64_t)ctx->num_sms * 16;
    return (int)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace

extern "C" int nis_narrow_c128_to_c32(nis_ctx* ctx, const double* src, nis_c32* dst, uint64_t n, nis_stream stream) {
    NIS_REQUIRE(ctx && src && dst, "nis_narrow_c128_to_c32: null argument");
    if (n == 0) return NIS_OK;
    k_narrow<<<grid_for(ctx, n), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double2*>(src),
                                                                reinterpret_cast<float2*>(dst), n);
    NIS_LAUNCH_CHECK(ctx);
    return NIS_OK;
}

extern "C" int nis_widen_c32_to_c128(nis_ctx* ctx, const nis_c32* src, double* dst, uint64_t n, nis_stream stream) {
    NIS_REQUIRE(ctx && src && dst, "nis_widen_c32_to_c128: null argument");
    if (n == 0) return NIS_OK;
    k_widen<<<grid_for(ctx, n), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(src),
                                                               reinterpret_cast<double2*>(dst), n);
    NIS_LAUNCH_CHECK(ctx);
    return NIS_OK;
}

// host-converted transfers of the complex128 contract (hostcopy.cpp)
namespace nis {
int hostcopy_d2h_widen(int device, const float* dev_src, double* dst, size_t n_floats, int threads, cudaStream_t st);
int hostcopy_h2d_narrow(int device, const double* src, float* dev_dst, size_t n_floats, int threads, cudaStream_t st);
}
extern "C" int nis_d2h_widen(nis_ctx* ctx, const nis_c32* dev_src, double* host_dst, uint64_t n, int32_t threads,
                             nis_stream stream) {
    NIS_REQUIRE(ctx && dev_src && host_dst, "nis_d2h_widen: null argument");
    NIS_REQUIRE(threads >= 1 && threads <= 32, "nis_d2h_widen: threads = %d outside 1..32", threads);
    DeviceGuard guard(ctx->device);
    try {   // (std::thread / std::vector may throw; nothing crosses the C ABI)
        return hostcopy_d2h_widen(ctx->device, reinterpret_cast<const float*>(dev_src), host_dst, (size_t)n * 2, threads,
                                  (cudaStream_t)stream);
    } catch (const std::exception& e) {
        set_error("nis_d2h_widen: %s", e.what());
        return NIS_ERR_NOMEM;
    }
}
extern "C" int nis_h2d_narrow(nis_ctx* ctx, const double* host_src, nis_c32* dev_dst, uint64_t n, int32_t threads,
                              nis_stream stream) {
    NIS_REQUIRE(ctx && host_src && dev_dst, "nis_h2d_narrow: null argument");
    NIS_REQUIRE(threads >= 1 && threads <= 32, "nis_h2d_narrow: threads = %d outside 1..32", threads);
    DeviceGuard guard(ctx->device);
    try {
        return hostcopy_h2d_narrow(ctx->device, host_src, reinterpret_cast<float*>(dev_dst), (size_t)n * 2, threads,
                                   (cudaStream_t)stream);
    } catch (const std::exception& e) {
        set_error("nis_h2d_narrow: %s", e.what());
        return NIS_ERR_NOMEM;
    }
}

namespace nis {
int launch_transpose(nis_ctx* ctx, const float2* in, int64_t in_pitch, float2* out, int rows, int cols,
                     cudaStream_t st) {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32);
    k_transpose<<<grid, 256, 0, st>>>(in, in_pitch, out, rows, cols);
    NIS_LAUNCH_CHECK(ctx);
    return NIS_OK;
}
}  // namespace nis

extern "C" int nis_transpose_c32(nis_ctx* ctx, const nis_c32* in, nis_c32* out, int32_t rows, int32_t cols,
                                 nis_stream stream) {
    NIS_REQUIRE(ctx && in && out && rows > 0 && cols > 0, "nis_transpose_c32: bad argument");
    return launch_transpose(ctx, reinterpret_cast<const float2*>(in), cols, reinterpret_cast<float2*>(out), rows, cols,
                            (cudaStream_t)stream);
}

// ---- peer memory: buffers that other processes (one per GPU) map into their own address space over CUDA IPC, so that
// their kernels load from / reduce into this GPU's HBM directly over NVLink.
extern "C" int nis_peer_alloc(uint64_t bytes, void** ptr, uint8_t* handle64) {
    NIS_REQUIRE(ptr && handle64 && bytes > 0, "nis_peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    NIS_CUDA_TRY(cudaMalloc(ptr, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *ptr);
    if (e != cudaSuccess) {
        cudaFree(*ptr);
        *ptr = nullptr;
        nis::set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return NIS_ERR_CUDA;
    }
    memcpy(handle64, &h, 64);
    return NIS_OK;
}
extern "C" int nis_peer_free(void* ptr) {
    if (ptr) NIS_CUDA_TRY(cudaFree(ptr));
    return NIS_OK;
}
// maps a buffer exported by another process into the CURRENT device's context; peer access to the exporting GPU is
// enabled as part of the mapping (cudaIpcMemLazyEnablePeerAccess)
extern "C" int nis_peer_open(const uint8_t* handle64, void** ptr) {
    NIS_REQUIRE(ptr && handle64, "nis_peer_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    NIS_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return NIS_OK;
}
extern "C" int nis_peer_close(void* ptr) {
    if (ptr) NIS_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
    return NIS_OK;
}
