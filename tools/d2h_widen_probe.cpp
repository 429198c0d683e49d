// Probe behind nis_d2h_widen: is "complex64 over PCIe + float -> double on the host cores"
// faster than "complex128 over PCIe" for the reference's complex128 result contract?
//   g++ -O3 -pthread -I/usr/local/cuda/include tools/d2h_widen_probe.cpp -o lib/d2h_widen_probe \
//       -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt
// Prints one JSON line per measurement: plain D2H of n complex128, host widening alone at T threads, and the
// pipelined transfer (chunked D2H of complex64 into a pinned ring, T threads widen chunk i while chunk i+1.. arrive).
#include <cuda_runtime.h>
#include <immintrin.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

__attribute__((target("avx2"))) static void widen_avx2(const float* src, double* dst, size_t n) {
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31)) { dst[i] = src[i]; ++i; }
    for (; i + 8 <= n; i += 8) {
        const __m256 v = _mm256_loadu_ps(src + i);
        _mm256_stream_pd(dst + i, _mm256_cvtps_pd(_mm256_castps256_ps128(v)));
        _mm256_stream_pd(dst + i + 4, _mm256_cvtps_pd(_mm256_extractf128_ps(v, 1)));
    }
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}
static void widen_scalar(const float* src, double* dst, size_t n) { for (size_t i = 0; i < n; ++i) dst[i] = src[i]; }
static void widen(const float* s, double* d, size_t n) {
    if (__builtin_cpu_supports("avx2")) widen_avx2(s, d, n); else widen_scalar(s, d, n);
}

int main(int argc, char** argv) {
    const size_t n = (argc > 1 ? atoll(argv[1]) : 8192ll * 8192ll) * 2;   // floats
    const size_t chunk = (argc > 2 ? atoll(argv[2]) : 4) << 20;            // floats per chunk (x4 bytes)
    const bool pinned_dst = argc > 3 ? atoi(argv[3]) : 1;
    CK(cudaSetDevice(0));
    float* dsrc; double* dwide;
    CK(cudaMalloc(&dsrc, n * 4)); CK(cudaMalloc(&dwide, n * 8));
    CK(cudaMemset(dsrc, 0x3c, n * 4)); CK(cudaMemset(dwide, 0, n * 8));
    double* hdst;
    if (pinned_dst) CK(cudaHostAlloc(&hdst, n * 8, cudaHostAllocDefault)); else { hdst = (double*)aligned_alloc(4096, n * 8); memset(hdst, 0, n * 8); }
    cudaStream_t st; CK(cudaStreamCreate(&st));
    printf("{\"hw_threads\": %u, \"floats\": %zu, \"chunk_floats\": %zu, \"pinned_dst\": %d}\n", std::thread::hardware_concurrency(), n, chunk, (int)pinned_dst);
    // (1) the present path: D2H of complex128
    if (pinned_dst) for (int r = 0; r < 3; ++r) {
        const double t0 = now();
        CK(cudaMemcpyAsync(hdst, dwide, n * 8, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
        const double t = now() - t0;
        printf("{\"what\": \"d2h_c128\", \"ms\": %.3f, \"GBps\": %.1f}\n", t * 1e3, n * 8 / t * 1e-9);
    }
    const int NS = 4;
    float* ring[NS]; cudaEvent_t ev[NS];
    for (int i = 0; i < NS; ++i) { CK(cudaHostAlloc(&ring[i], chunk * 4, cudaHostAllocDefault)); CK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)); }
    float* hsrc; CK(cudaHostAlloc(&hsrc, n * 4, cudaHostAllocDefault));
    CK(cudaMemcpy(hsrc, dsrc, n * 4, cudaMemcpyDeviceToHost));
    {
        const double t0 = now();
        CK(cudaMemcpyAsync(hsrc, dsrc, n * 4, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
        const double t = now() - t0;
        printf("{\"what\": \"d2h_c64\", \"ms\": %.3f, \"GBps\": %.1f}\n", t * 1e3, n * 4 / t * 1e-9);
    }
    const int tlist[] = {1, 2, 4, 8, 12, 16, 24, 32};
    for (int T : tlist) {
        if (T > (int)std::thread::hardware_concurrency()) break;
        // (2) widening alone, whole array, T threads
        double best = 1e9;
        for (int r = 0; r < 3; ++r) {
            const double t0 = now();
            std::vector<std::thread> th;
            for (int k = 0; k < T; ++k) th.emplace_back([&, k] { const size_t a = n * k / T / 8 * 8, b = (k == T - 1) ? n : n * (k + 1) / T / 8 * 8; widen(hsrc + a, hdst + a, b - a); });
            for (auto& x : th) x.join();
            best = std::min(best, now() - t0);
        }
        printf("{\"what\": \"widen_only\", \"threads\": %d, \"ms\": %.3f, \"GBps_written\": %.1f}\n", T, best * 1e3, n * 8 / best * 1e-9);
        // (3) pipelined: ring of NS pinned chunks; the main thread is worker 0
        best = 1e9;
        for (int r = 0; r < 3; ++r) {
            const size_t nchunks = (n + chunk - 1) / chunk;
            std::atomic<long> ready{0};
            std::vector<std::atomic<int>> done(nchunks);
            for (auto& d : done) d.store(0);
            auto work = [&](int k, size_t c) {
                const size_t len = std::min(chunk, n - c * chunk);
                const size_t a = len * k / T / 8 * 8, b = (k == T - 1) ? len : len * (k + 1) / T / 8 * 8;
                widen(ring[c % NS] + a, hdst + c * chunk + a, b - a);
                done[c].fetch_add(1, std::memory_order_release);
            };
            const double t0 = now();
            std::vector<std::thread> th;
            for (int k = 1; k < T; ++k) th.emplace_back([&, k] {
                for (size_t c = 0; c < nchunks; ++c) {
                    while (ready.load(std::memory_order_acquire) <= (long)c) _mm_pause();
                    work(k, c);
                }
            });
            auto enqueue = [&](size_t c) {
                const size_t len = std::min(chunk, n - c * chunk);
                CK(cudaMemcpyAsync(ring[c % NS], dsrc + c * chunk, len * 4, cudaMemcpyDeviceToHost, st));
                CK(cudaEventRecord(ev[c % NS], st));
            };
            for (size_t c = 0; c < std::min<size_t>(NS, nchunks); ++c) enqueue(c);
            for (size_t c = 0; c < nchunks; ++c) {
                CK(cudaEventSynchronize(ev[c % NS]));
                ready.store((long)c + 1, std::memory_order_release);
                work(0, c);
                while (done[c].load(std::memory_order_acquire) < T) _mm_pause();
                if (c + NS < nchunks) enqueue(c + NS);
            }
            for (auto& x : th) x.join();
            best = std::min(best, now() - t0);
        }
        printf("{\"what\": \"pipelined_d2h_c64_widen\", \"threads\": %d, \"ms\": %.3f, \"GBps_of_c128\": %.1f}\n", T, best * 1e3, n * 8 / best * 1e-9);
    }
    // correctness of the last pipelined run
    size_t bad = 0;
    for (size_t i = 0; i < n; i += 4097) bad += (hdst[i] != (double)hsrc[i]);
    printf("{\"mismatches\": %zu}\n", bad);
    return 0;
}
