// pipebench.cu -- does the packed fp32x2 arithmetic of sm_100 (FADD2 / FFMA2) double complex add/mul throughput?
// Measures thread-level ops per clock per SM for scalar FADD pairs vs FADD2, and FFMA pairs vs FFMA2.
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 upk(unsigned long long r) { float2 c; asm("mov.b64 {%0, %1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(r)); return c; }

template <int MODE>
__global__ void k(float2* out, int iters, float2 seed) {
    constexpr int U = 8;
    float2 a[U];
    unsigned long long p[U], s = pk(seed.x, seed.y);
    for (int i = 0; i < U; ++i) { a[i] = make_float2(threadIdx.x + i, i); p[i] = pk(a[i].x, a[i].y); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            if (MODE == 0) { a[i].x += seed.x; a[i].y += seed.y; }
            if (MODE == 1) asm("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(s));
            if (MODE == 2) { a[i].x = fmaf(a[i].x, seed.x, seed.y); a[i].y = fmaf(a[i].y, seed.x, seed.y); }
            if (MODE == 3) asm("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(s));
        }
    }
    float2 acc = make_float2(0, 0);
    for (int i = 0; i < U; ++i) { float2 v = (MODE & 1) ? upk(p[i]) : a[i]; acc.x += v.x; acc.y += v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name) {
    float2* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2));
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 100, make_float2(1.0001f, 0.5f));
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, iters, make_float2(1.0001f, 0.5f));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops_pairs = 148.0 * 8 * 256 * (double)iters * 8;   // complex (2-float) ops
    printf("{\"mode\": \"%s\", \"ms\": %.3f, \"G_complex_ops_per_s\": %.1f}\n", name, ms, flops_pairs / ms * 1e-6);
    cudaFree(out);
}
int main() { run<0>("FADD x2 scalar"); run<1>("FADD2 packed"); run<2>("FFMA x2 scalar"); run<3>("FFMA2 packed"); return 0; }
