// Shared internals of libpss.so: the context object, error plumbing, small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pss.h"

#define PSS_VERSION 100

struct pss_fft_tables {
    void* twiddle = nullptr;   // device: per-pass base twiddles, complex<T>
    void* window[3] = {nullptr, nullptr, nullptr};  // device: T[N] for HAMMING / HANN (index = enum)
};

struct pss_demod_plan;  // pss_demod.cu

struct pss_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;       // own_stream or an adopted one
    int64_t launches = 0;
    std::string last_error;
    // FFT tables keyed by (log2N, precision)
    std::map<int, pss_fft_tables> fft_tables;
    // scratch for host-pointer variants (grown on demand)
    void* d_in = nullptr;   size_t d_in_bytes = 0;
    void* d_out = nullptr;  size_t d_out_bytes = 0;
    void* d_aux = nullptr;  size_t d_aux_bytes = 0;
    void* d_aux2 = nullptr; size_t d_aux2_bytes = 0;
    // pipeline scratch (pss_pipeline.cu)
    void* p_buf[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t p_bytes[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    // demod plans keyed by (mode, fs, N)
    std::map<std::string, pss_demod_plan*> demod_plans;
    // display rings (pss_display.cu)
    std::map<int, void*> displays;
};

int pss_fail_cuda(pss_ctx* ctx, cudaError_t e, const char* what, const char* file, int line);

#define PSS_CUDA(ctx, call)                                                          \
    do {                                                                             \
        cudaError_t _e = (call);                                                     \
        if (_e != cudaSuccess) return pss_fail_cuda((ctx), _e, #call, __FILE__, __LINE__); \
    } while (0)

#define PSS_LAUNCH_CHECK(ctx)                                                        \
    do {                                                                             \
        cudaError_t _e = cudaGetLastError();                                         \
        if (_e != cudaSuccess) return pss_fail_cuda((ctx), _e, "kernel launch", __FILE__, __LINE__); \
        (ctx)->launches++;                                                           \
    } while (0)

// grow-only device scratch
int pss_reserve(pss_ctx* ctx, void** p, size_t* have, size_t need);

// ---------------------------------------------------------------- device helpers
template <typename T>
struct cx {
    T x, y;
};

template <typename T>
__device__ __forceinline__ cx<T> cmul(const cx<T> a, const cx<T> b) {
    cx<T> r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
template <typename T>
__device__ __forceinline__ cx<T> csqr(const cx<T> a) {
    cx<T> r;
    r.x = a.x * a.x - a.y * a.y;
    r.y = (a.x + a.x) * a.y;
    return r;
}
template <typename T>
__device__ __forceinline__ cx<T> cadd(const cx<T> a, const cx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T>
__device__ __forceinline__ cx<T> csub(const cx<T> a, const cx<T> b) { return {a.x - b.x, a.y - b.y}; }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 10*log10(p) for p >= 1e-10 from an fp64 power, evaluated in fp32 (an fp32
// power/log tail is within 1e-5 dB; the fp64 part is everything up to |X|^2 + 1e-10).
// lg2.approx = one MUFU.LG2: exponent + log2(mantissa) with ~2^-22 absolute error on the mantissa
// part, i.e. < 1e-5 dB after the 3.0103 scale.  p is a normal float here (>= 1e-10), so the
// denormal pre-scaling of log2f() is not needed.
__device__ __forceinline__ float db_from_power(double p) {
    float l;
    const float pf = (float)p;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(pf));
    return 3.01029995663981195f * l;
}
