// Small reductions / conversions on either side of the hot path:
//   measure_signal_power (signal_processing.py:325-328) and the WAV/pipe int16 pack
//   (audio_processing.py:36-38, io_manager.py:25-26).
#include <math.h>

#include "pss_common.cuh"

__global__ void __launch_bounds__(256)
power_kernel(const float2* __restrict__ iq, const int N, const long long n_frames, float* __restrict__ out_db) {
    __shared__ double red[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (long long f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const float2* x = iq + f * N;
        double acc = 0.0;
        for (int i = tid; i < N; i += 256) {
            const float2 v = __ldg(x + i);
            const float a = hypotf(v.x, v.y);              // np.abs(complex64) -> float32
            acc += (double)__fmul_rn(a, a);                 // ** 2 in float32, mean accumulated wider
        }
        acc = warp_sum(acc);
        __syncthreads();
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += red[w];
            out_db[f] = (float)(10.0 * log10(t / (double)N + 1e-10));
        }
    }
}

__global__ void int16_kernel(const float* __restrict__ a, const long long n, int16_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int16_t)(int)((double)a[i] * 32767.0);     // C cast: truncation toward zero
}

// np.int16(samples * 32767) on the float64 array the reference holds (audio_processing.py:36-38): one fp64
// product, then the C cast (truncation toward zero; wrap-around of out-of-range values like the x86 cast).
__global__ void int16_f64_kernel(const double* __restrict__ a, const long long n, int16_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int16_t)(int)__dmul_rn(a[i], 32767.0);
}

extern "C" {

int pss_power_c64(pss_ctx* ctx, const float* iq, int N, int64_t n_frames, float* power_db) {
    if (!ctx || !iq || !power_db || N < 1 || n_frames < 0) return PSS_ERR_ARG;
    if (n_frames == 0) return PSS_OK;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t in_b = (size_t)n_frames * N * 8;
    int rc;
    if ((rc = pss_reserve(ctx, &ctx->d_in, &ctx->d_in_bytes, in_b))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_aux, &ctx->d_aux_bytes, (size_t)n_frames * 4))) return rc;
    PSS_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, iq, in_b, cudaMemcpyHostToDevice, ctx->stream));
    long long grid = n_frames < 4LL * ctx->sm_count ? n_frames : 4LL * ctx->sm_count;
    power_kernel<<<(unsigned)grid, 256, 0, ctx->stream>>>((const float2*)ctx->d_in, N, n_frames, (float*)ctx->d_aux);
    PSS_LAUNCH_CHECK(ctx);
    PSS_CUDA(ctx, cudaMemcpyAsync(power_db, ctx->d_aux, (size_t)n_frames * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PSS_OK;
}

int pss_audio_to_int16(pss_ctx* ctx, const float* audio, int64_t n, int16_t* pcm) {
    if (!ctx || !audio || !pcm || n < 0) return PSS_ERR_ARG;
    if (n == 0) return PSS_OK;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = pss_reserve(ctx, &ctx->d_in, &ctx->d_in_bytes, (size_t)n * 4))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_out, &ctx->d_out_bytes, (size_t)n * 2))) return rc;
    PSS_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, audio, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    int16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const float*)ctx->d_in, n, (int16_t*)ctx->d_out);
    PSS_LAUNCH_CHECK(ctx);
    PSS_CUDA(ctx, cudaMemcpyAsync(pcm, ctx->d_out, (size_t)n * 2, cudaMemcpyDeviceToHost, ctx->stream));
    PSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PSS_OK;
}

int pss_audio_to_int16_f64(pss_ctx* ctx, const double* audio, int64_t n, int16_t* pcm) {
    if (!ctx || !audio || !pcm || n < 0) return PSS_ERR_ARG;
    if (n == 0) return PSS_OK;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = pss_reserve(ctx, &ctx->d_in, &ctx->d_in_bytes, (size_t)n * 8))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_out, &ctx->d_out_bytes, (size_t)n * 2))) return rc;
    PSS_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, audio, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    int16_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const double*)ctx->d_in, n, (int16_t*)ctx->d_out);
    PSS_LAUNCH_CHECK(ctx);
    PSS_CUDA(ctx, cudaMemcpyAsync(pcm, ctx->d_out, (size_t)n * 2, cudaMemcpyDeviceToHost, ctx->stream));
    PSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PSS_OK;
}

}  // extern "C"
