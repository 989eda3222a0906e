// Shared internals of libpss.so: the context object, error plumbing, small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <set>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pss.h"

#define PSS_VERSION 200

struct pss_fft_tables {
    void* twiddle = nullptr;   // device: per-pass base twiddles, complex<T>
    void* window[3] = {nullptr, nullptr, nullptr};  // device: T[N] for HAMMING / HANN (index = enum)
};

struct pss_demod_plan;  // pss_demod.cu

// tables of the large transforms (pss_psd.cu), keyed by log2 N
struct pss_large_tables {
    void *tw1 = nullptr, *thi = nullptr, *tlo = nullptr;   // big path: W_N1^k, W_N^(1024 j), W_N^j
    void* twN = nullptr;                                   // cx<double>[N2]: W_N^n2
    void* window[3] = {nullptr, nullptr, nullptr};
    bool ready = false;                                    // set only after every allocation succeeded
};

// copy / compute overlap of pss_pipeline_c64 (pss_pipeline.cu)
struct pss_pipe_streams {
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaEvent_t in_ready[2] = {nullptr, nullptr};     // H2D of slot s finished
    cudaEvent_t in_free[2] = {nullptr, nullptr};      // kernels reading slot s finished
    cudaEvent_t done = nullptr;                       // kernels of the chunk finished
    cudaEvent_t db_free = nullptr;                    // D2H of the db scratch finished
    bool ready = false;
};

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
    void* p_buf[16] = {};
    size_t p_bytes[16] = {};
    // demod plans keyed by (mode, fs, N)
    std::map<std::string, pss_demod_plan*> demod_plans;
    // display rings (pss_display.cu)
    std::map<int, void*> displays;
    // large-transform tables (pss_psd.cu), classifier window, pipeline streams: all owned by the context
    std::map<int, pss_large_tables> large_tables;
    void* hann_periodic = nullptr;
    pss_pipe_streams pipe;
    std::set<const void*> configured;    // kernels whose dynamic shared-memory limit was raised on this device
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

// Programmatic dependent launch (PSS_PDL=1): the kernels of the bench step's chain -- PSD, display render, iq-correction
// coefficients, forcing, scan -- are launched with the programmatic-stream-serialisation attribute.  Each of them runs
// its prologue (tables into shared memory, barrier set-up), then `griddepcontrol.wait` (the previous kernel of the
// stream has completed and its writes are visible), then `griddepcontrol.launch_dependents`, so that the next
// kernel's CTAs are scheduled while this one drains and meet their own wait with the prologue already done.  Without
// the attribute both instructions do nothing.
bool pss_use_pdl();
template <typename... KArgs, typename... Args>
static inline cudaError_t pss_launch(void (*kern)(KArgs...), const unsigned grid, const unsigned block, const size_t smem,
                                     cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pss_use_pdl() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ void pss_grid_dependency_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// geometry of an open display stream (pss_display.cu); PSS_ERR_ARG if it does not exist
int pss_display_geom(const pss_ctx* ctx, int stream, int* W, int* rows_max);

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

// Order-preserving float <-> unsigned key (total order: -inf < ... < -0 < +0 < ... < +inf < NaN).
__device__ __forceinline__ unsigned f2key(float f) {
    unsigned u = __float_as_uint(f);
    return u ^ ((unsigned)((int)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
    unsigned u = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
    return __uint_as_float(u);
}

// Exact order statistics of a row streamed from global/L2 memory by one CTA of 512 threads:
// returns the keys of the elements of (0-based) rank `rank` and rank+1 (rank+1 clamped to n-1).
// 4 x 8-bit radix select on keys normalised to the occupied range, then one counting pass.
// `hist` [256] and `us` [8] are shared scratch; all threads of the CTA must call it.
__device__ __forceinline__ void row_select2_512(const float* __restrict__ row, const int n, const unsigned rank_in,
                                                unsigned* hist, unsigned* us, unsigned& key_a, unsigned& key_b) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 8) us[tid] = (tid == 0 || tid == 5) ? 0xffffffffu : 0u;
    __syncthreads();
    unsigned kmin = 0xffffffffu, kmax = 0u;
    for (int i = tid; i < n; i += 512) {
        const unsigned k = f2key(row[i]);
        kmin = min(kmin, k);
        kmax = max(kmax, k);
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0) {
        atomicMin(&us[0], kmin);
        atomicMax(&us[1], kmax);
    }
    __syncthreads();
    kmin = us[0];
    kmax = us[1];
    const int common = min(__clz((int)(kmin ^ kmax)), 31);
    unsigned rank = rank_in, prefix = 0u;
    for (int ps = 0; ps < 4; ++ps) {
        const int shift = 24 - 8 * ps;
        if (tid < 256) hist[tid] = 0u;
        __syncthreads();
        for (int i = tid; i < n; i += 512) {
            const unsigned k = (f2key(row[i]) - kmin) << common;
            if (ps == 0 || (k >> (shift + 8)) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (warp == 0) {
            unsigned c[8], sum = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                c[q] = hist[8 * lane + q];
                sum += c[q];
            }
            unsigned incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            const unsigned hit = __ballot_sync(0xffffffffu, incl > rank);
            const int Ln = __ffs(hit) - 1;
            if (lane == Ln) {
                unsigned r = rank - (incl - sum);
                int dg = 0;
                bool found = false;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (!found) {
                        if (r < c[q]) { dg = q; found = true; }
                        else r -= c[q];
                    }
                }
                us[2] = (unsigned)(8 * lane + dg);
                us[3] = r;
            }
        }
        __syncthreads();
        prefix = (prefix << 8) | us[2];
        rank = us[3];
    }
    unsigned cnt_le = 0, min_gt = 0xffffffffu;
    for (int i = tid; i < n; i += 512) {
        const unsigned k = (f2key(row[i]) - kmin) << common;
        cnt_le += k <= prefix;
        if (k > prefix) min_gt = min(min_gt, k);
    }
    cnt_le = __reduce_add_sync(0xffffffffu, cnt_le);
    min_gt = __reduce_min_sync(0xffffffffu, min_gt);
    if (lane == 0) {
        atomicAdd(&us[4], cnt_le);
        atomicMin(&us[5], min_gt);
    }
    __syncthreads();
    key_a = (prefix >> common) + kmin;
    const unsigned nb = (us[4] > rank_in + 1u || us[5] == 0xffffffffu) ? prefix : us[5];
    key_b = (nb >> common) + kmin;
    __syncthreads();
}
