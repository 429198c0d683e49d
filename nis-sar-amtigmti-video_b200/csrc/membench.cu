// membench.cu -- access-pattern microbenchmarks that decide the CSA data-flow design on B200.
// Every pattern moves an N x N complex64 matrix once in and once out (16 B / element); the number
// printed is (2 * N*N*8 B) / time, directly comparable with MEASURED_PEAKS.json's copy bandwidth.
//   copy            float4 grid-stride copy (ceiling)
//   transpose32     32x32 shared-memory tile transpose (256 B pieces both sides)
//   coltile RxW     in-place read+write of [R contiguous rows x W columns] tiles (k_az_inner's pattern)
//   colstride RxW   in-place read+write of R rows at stride N/R x W columns (k_az_outer's pattern)
//   rowT G          read G full rows, write them transposed as G*8-byte pieces (a row-FFT kernel that
//                   corner-turns on the way out: the 3-pass CSA design)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void k_copy(const float4* __restrict__ a, float4* __restrict__ b, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

__global__ void __launch_bounds__(256) k_transpose32(const float2* __restrict__ in, float2* __restrict__ out, int n) {
    __shared__ float2 tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) tile[ty + 8 * i][tx] = in[(size_t)(r0 + ty + 8 * i) * n + c0 + tx];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) out[(size_t)(c0 + ty + 8 * i) * n + r0 + tx] = tile[tx][ty + 8 * i];
}

// tile of R rows (row r of the tile is matrix row rbase + r*rstride) x W columns, E rows per thread
template <int W, int E>
__global__ void k_tile(float2* __restrict__ data, int n, int R, int rstride_is_strided) {
    const int c = threadIdx.x % W, t = threadIdx.x / W, NT = blockDim.x / W;
    const int tiles_per_col = n / R;
    const int tb = blockIdx.y;  // which row-group
    size_t rbase, rstride;
    if (rstride_is_strided) { rbase = tb; rstride = tiles_per_col; } else { rbase = (size_t)tb * R; rstride = 1; }
    float2* base = data + rbase * n + blockIdx.x * W + c;
    float2 v[E];
#pragma unroll
    for (int s = 0; s < E; ++s) v[s] = base[(size_t)(t + NT * s) * rstride * n];
#pragma unroll
    for (int s = 0; s < E; ++s) { v[s].x += 1.f; base[(size_t)(t + NT * s) * rstride * n] = v[s]; }
}

// G rows per CTA, chunks of 256 columns staged through shared memory, written transposed
template <int G>
__global__ void __launch_bounds__(256) k_rowT(const float2* __restrict__ in, float2* __restrict__ out, int n) {
    __shared__ float2 sm[G][257];
    const int r0 = blockIdx.x * G;
    for (int c0 = 0; c0 < n; c0 += 256) {
#pragma unroll
        for (int g = 0; g < G; ++g) sm[g][threadIdx.x] = in[(size_t)(r0 + g) * n + c0 + threadIdx.x];
        __syncthreads();
        // 256 columns x G rows -> each output piece is G contiguous elements
        for (int i = threadIdx.x; i < 256 * G; i += 256) {
            const int g = i % G, c = i / G;
            out[(size_t)(c0 + c) * n + r0 + g] = sm[g][c];
        }
        __syncthreads();
    }
}

template <class F>
float time_it(F f, int iters = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < iters; ++i) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / iters;
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 8192;
    const size_t bytes = (size_t)n * n * sizeof(float2);
    float2 *a, *b;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 0, bytes)); CK(cudaMemset(b, 0, bytes));
    auto report = [&](const char* name, float ms) {
        printf("{\"n\": %d, \"pattern\": \"%s\", \"ms\": %.4f, \"GBps\": %.1f}\n", n, name, ms, 2.0 * bytes / ms * 1e-6);
        fflush(stdout);
    };
    report("copy", time_it([&] { k_copy<<<148 * 16, 256>>>((const float4*)a, (float4*)b, bytes / 16); }));
    report("transpose32", time_it([&] { k_transpose32<<<dim3(n / 32, n / 32), 256>>>(a, b, n); }));
    report("coltile 256x32", time_it([&] { k_tile<32, 16><<<dim3(n / 32, n / 256), 512>>>(a, n, 256, 0); }));
    report("coltile 256x16", time_it([&] { k_tile<16, 16><<<dim3(n / 16, n / 256), 256>>>(a, n, 256, 0); }));
    report("coltile 512x16", time_it([&] { k_tile<16, 16><<<dim3(n / 16, n / 512), 512>>>(a, n, 512, 0); }));
    report("coltile 256x8", time_it([&] { k_tile<8, 16><<<dim3(n / 8, n / 256), 128>>>(a, n, 256, 0); }));
    report("coltile 256x4", time_it([&] { k_tile<4, 16><<<dim3(n / 4, n / 256), 64>>>(a, n, 256, 0); }));
    report("colstride 16x32", time_it([&] { k_tile<32, 16><<<dim3(n / 32, n / 16), 32>>>(a, n, 16, 1); }));
    report("colstride 16x32 b256", time_it([&] { k_tile<32, 2><<<dim3(n / 32, n / 16), 256>>>(a, n, 16, 1); }));
    report("rowT 2", time_it([&] { k_rowT<2><<<n / 2, 256>>>(a, b, n); }));
    report("rowT 4", time_it([&] { k_rowT<4><<<n / 4, 256>>>(a, b, n); }));
    report("rowT 8", time_it([&] { k_rowT<8><<<n / 8, 256>>>(a, b, n); }));
    report("rowT 16", time_it([&] { k_rowT<16><<<n / 16, 256>>>(a, b, n); }));
    CK(cudaFree(a)); CK(cudaFree(b));
    return 0;
}
