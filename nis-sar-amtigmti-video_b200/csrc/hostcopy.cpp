// hostcopy.cpp -- the host half of the complex128 array contract (plain C++, no device code).
//
// The reference's functions take and return complex128 numpy arrays; the device holds complex64.
// Moving the WIDE form over PCIe costs twice the bytes of the data that exists: 1.07 GB for one 8192 x 8192 image,
// 19 ms at the 56 GB/s the link delivers -- 90 % of the end-to-end step.  These two transfers move the NARROW form and
// convert on the host cores while the DMA engine runs:
//   d2h_widen   device complex64 -> chunked DMA into a ring of page-locked slots -> T threads widen float -> double
//               (AVX2, non-temporal stores) straight into the caller's array, one whole chunk per thread, no barrier
//   h2d_narrow  caller's complex128 array (pageable is fine) -> T threads narrow into the ring -> chunked DMA
// Measured (tools/d2h_widen_probe.cpp, tools/e2e_route_probe.py): 8192^2 image 19 ms -> 12-15 ms
// with 8-12 threads on a rank that has the host to itself; with several ranks the host memory path is the limit either way.
// float -> double is exact and double -> float rounds to nearest even, exactly as the device kernels k_widen / k_narrow
// (api.cu) and numpy's astype do, so the bytes the caller sees do not depend on the route.
#include <cuda_runtime.h>
#include <immintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

namespace nis {
void set_error(const char* fmt, ...);   // api.cu

namespace {

constexpr int kMaxSlots = 64;
constexpr size_t kSlotFloats = size_t(1) << 20;    // capacity of a ring slot: 4 MiB of complex64 = 8 MiB of complex128

// chunk length (floats) and ring depth in use; NIS_HOST_CHUNK_KB / NIS_HOST_SLOTS override them (development knobs)
size_t chunk_floats() {
    static const size_t v = [] {
        size_t kb = 1024;   // measured: 1 MiB chunks beat 4 MiB by 10-15 % alone and with two ranks (staging reads hit the LLC)
        if (const char* e = getenv("NIS_HOST_CHUNK_KB")) kb = (size_t)atol(e);
        if (kb < 64) kb = 64;
        if (kb > 4096) kb = 4096;
        return kb * 1024 / sizeof(float);
    }();
    return v;
}
int ring_slots(int threads) {
    static const int forced = [] {
        const char* e = getenv("NIS_HOST_SLOTS");
        return e ? atoi(e) : 0;
    }();
    int n = forced > 0 ? forced : 2 * threads + 4;
    if (n < threads + 2) n = threads + 2;
    return n > kMaxSlots ? kMaxSlots : n;
}

struct Ring {
    std::mutex mu;                 // one transfer at a time per device
    float* slot[kMaxSlots] = {};
    cudaEvent_t ev[kMaxSlots] = {};
    int n = 0;
};
Ring g_ring[64];

bool ring_reserve(Ring& r, int want, cudaError_t* err) {
    while (r.n < want) {
        float* p = nullptr;
        cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&p), kSlotFloats * sizeof(float), cudaHostAllocDefault);
        if (e == cudaSuccess) {
            e = cudaEventCreateWithFlags(&r.ev[r.n], cudaEventDisableTiming);
            if (e != cudaSuccess) cudaFreeHost(p);
        }
        if (e != cudaSuccess) { *err = e; return false; }
        r.slot[r.n++] = p;
    }
    return true;
}

__attribute__((target("avx2"))) void widen_avx2(const float* __restrict__ src, double* __restrict__ dst, size_t n) {
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31)) { dst[i] = src[i]; ++i; }
    for (; i + 8 <= n; i += 8) {
        const __m256 v = _mm256_loadu_ps(src + i);
        _mm256_stream_pd(dst + i, _mm256_cvtps_pd(_mm256_castps256_ps128(v)));
        _mm256_stream_pd(dst + i + 4, _mm256_cvtps_pd(_mm256_extractf128_ps(v, 1)));
    }
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}
__attribute__((target("avx2"))) void narrow_avx2(const double* __restrict__ src, float* __restrict__ dst, size_t n) {
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        const __m128 lo = _mm256_cvtpd_ps(_mm256_loadu_pd(src + i)), hi = _mm256_cvtpd_ps(_mm256_loadu_pd(src + i + 4));
        _mm256_storeu_ps(dst + i, _mm256_set_m128(hi, lo));
    }
    for (; i < n; ++i) dst[i] = (float)src[i];
}
void widen(const float* s, double* d, size_t n) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) return widen_avx2(s, d, n);
    for (size_t i = 0; i < n; ++i) d[i] = s[i];
}
void narrow(const double* s, float* d, size_t n) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) return narrow_avx2(s, d, n);
    for (size_t i = 0; i < n; ++i) d[i] = (float)s[i];
}

// spin politely: a burst of pauses, then give the core away (the workers may share cores with the threads of other ranks)
int spin_limit() {   // NIS_HOST_SPIN: pauses before the first yield (0: never yield; development knob)
    static const int v = [] { const char* e = getenv("NIS_HOST_SPIN"); return e ? atoi(e) : 2048; }();
    return v;
}
struct Backoff {
    int n = 0;
    void operator()() {
        const int lim = spin_limit();
        if (lim == 0 || ++n < lim) _mm_pause();
        else { std::this_thread::yield(); n = lim / 2; }
    }
};

int fail_cuda(const char* what, cudaError_t e) {
    set_error("%s failed: %s", what, cudaGetErrorString(e));
    return -2;   // NIS_ERR_CUDA
}

}  // namespace

// dst[0 .. n_floats) = (double) dev_src[0 .. n_floats); returns when dst is complete.
int hostcopy_d2h_widen(int device, const float* dev_src, double* dst, size_t n_floats, int threads, cudaStream_t st) {
    if (n_floats == 0) return 0;
    if (threads < 1) threads = 1;
    if (threads > 32) threads = 32;
    Ring& r = g_ring[device & 63];
    std::lock_guard<std::mutex> lock(r.mu);
    const size_t kChunkFloats = chunk_floats();
    const size_t nchunks = (n_floats + kChunkFloats - 1) / kChunkFloats;
    const int NS = (int)std::min<size_t>(nchunks, (size_t)ring_slots(threads));
    cudaError_t err = cudaSuccess;
    if (!ring_reserve(r, NS, &err)) return fail_cuda("cudaHostAlloc (transfer ring)", err);

    std::atomic<long> arrived{0};            // chunks [0, arrived) are in their slots
    std::atomic<int> abort_flag{0};
    std::vector<std::atomic<int>> done(nchunks);
    for (auto& d : done) d.store(0, std::memory_order_relaxed);
    auto len_of = [&](size_t c) { return std::min(kChunkFloats, n_floats - c * kChunkFloats); };
    auto worker = [&](int k) {
        for (size_t c = (size_t)k; c < nchunks; c += (size_t)threads) {
            Backoff relax;
            while (arrived.load(std::memory_order_acquire) <= (long)c) {
                if (abort_flag.load(std::memory_order_relaxed)) return;
                relax();
            }
            widen(r.slot[c % NS], dst + c * kChunkFloats, len_of(c));
            done[c].store(1, std::memory_order_release);
        }
    };
    std::vector<std::thread> pool;
    pool.reserve(threads);
    try {
        for (int k = 0; k < threads; ++k) pool.emplace_back(worker, k);
    } catch (...) {   // thread creation failed: release the workers that did start, then report
        abort_flag.store(1);
        for (auto& t : pool) t.join();
        throw;
    }

    size_t next = 0;                          // next chunk to enqueue
    auto enqueue_ready = [&]() -> cudaError_t {
        // chunk `next` may take its slot once chunk next - NS has been consumed
        while (next < nchunks && (next < (size_t)NS || done[next - NS].load(std::memory_order_acquire))) {
            cudaError_t e = cudaMemcpyAsync(r.slot[next % NS], dev_src + next * kChunkFloats, len_of(next) * sizeof(float),
                                            cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaEventRecord(r.ev[next % NS], st);
            if (e != cudaSuccess) return e;
            ++next;
        }
        return cudaSuccess;
    };
    for (size_t c = 0; c < nchunks && err == cudaSuccess; ++c) {
        Backoff relax;
        while (err == cudaSuccess && next <= c) {    // (cannot stall: chunk c - NS was published long ago)
            err = enqueue_ready();
            if (next <= c) relax();
        }
        if (err != cudaSuccess) break;
        // wait for chunk c without sleeping in the driver, so that freed slots are refilled at once
        for (;;) {
            const cudaError_t q = cudaEventQuery(r.ev[c % NS]);
            if (q == cudaSuccess) break;
            if (q != cudaErrorNotReady) { err = q; break; }
            err = enqueue_ready();
            if (err != cudaSuccess) break;
            relax();
        }
        if (err != cudaSuccess) break;
        arrived.store((long)c + 1, std::memory_order_release);
        err = enqueue_ready();
    }
    if (err != cudaSuccess) abort_flag.store(1);
    for (auto& t : pool) t.join();
    if (err != cudaSuccess) {
        cudaStreamSynchronize(st);
        return fail_cuda("device-to-host transfer", err);
    }
    return 0;
}

// dev_dst[0 .. n_floats) = (float) src[0 .. n_floats); returns when the device buffer is complete.
int hostcopy_h2d_narrow(int device, const double* src, float* dev_dst, size_t n_floats, int threads, cudaStream_t st) {
    if (n_floats == 0) return 0;
    if (threads < 1) threads = 1;
    if (threads > 32) threads = 32;
    Ring& r = g_ring[device & 63];
    std::lock_guard<std::mutex> lock(r.mu);
    const size_t kChunkFloats = chunk_floats();
    const size_t nchunks = (n_floats + kChunkFloats - 1) / kChunkFloats;
    const int NS = (int)std::min<size_t>(nchunks, (size_t)ring_slots(threads));
    cudaError_t err = cudaSuccess;
    if (!ring_reserve(r, NS, &err)) return fail_cuda("cudaHostAlloc (transfer ring)", err);

    std::atomic<long> freed{0};              // the DMA reads of chunks [0, freed) have completed
    std::atomic<int> abort_flag{0};
    std::vector<std::atomic<int>> filled(nchunks);
    for (auto& f : filled) f.store(0, std::memory_order_relaxed);
    auto len_of = [&](size_t c) { return std::min(kChunkFloats, n_floats - c * kChunkFloats); };
    auto worker = [&](int k) {
        for (size_t c = (size_t)k; c < nchunks; c += (size_t)threads) {
            Backoff relax;
            while ((long)c >= freed.load(std::memory_order_acquire) + NS) {
                if (abort_flag.load(std::memory_order_relaxed)) return;
                relax();
            }
            narrow(src + c * kChunkFloats, r.slot[c % NS], len_of(c));
            filled[c].store(1, std::memory_order_release);
        }
    };
    std::vector<std::thread> pool;
    pool.reserve(threads);
    try {
        for (int k = 0; k < threads; ++k) pool.emplace_back(worker, k);
    } catch (...) {   // thread creation failed: release the workers that did start, then report
        abort_flag.store(1);
        for (auto& t : pool) t.join();
        throw;
    }

    size_t sent = 0;
    long fr = 0;
    auto poll_freed = [&]() {
        while ((size_t)fr < sent) {
            const cudaError_t q = cudaEventQuery(r.ev[fr % NS]);
            if (q == cudaErrorNotReady) break;
            if (q != cudaSuccess) { err = q; break; }
            freed.store(++fr, std::memory_order_release);
        }
    };
    Backoff relax;
    while (sent < nchunks && err == cudaSuccess) {
        if (filled[sent].load(std::memory_order_acquire)) {
            err = cudaMemcpyAsync(dev_dst + sent * kChunkFloats, r.slot[sent % NS], len_of(sent) * sizeof(float),
                                  cudaMemcpyHostToDevice, st);
            if (err == cudaSuccess) err = cudaEventRecord(r.ev[sent % NS], st);
            ++sent;
        } else {
            relax();
        }
        if (err == cudaSuccess) poll_freed();
    }
    if (err != cudaSuccess) abort_flag.store(1);
    for (auto& t : pool) t.join();
    const cudaError_t e2 = cudaStreamSynchronize(st);   // the ring is reusable, the device buffer complete
    if (err == cudaSuccess) err = e2;
    if (err != cudaSuccess) return fail_cuda("host-to-device transfer", err);
    return 0;
}

}  // namespace nis
