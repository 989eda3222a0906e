// Display accumulate: history stack range + normalisation of the W-column rows the PSD kernel
// emitted (pyspecsdr.py:1351-1358, 1373-1398 waterfall; :1649-1696 gradient; :1521-1556 persistence).
#include <math.h>

#include "pss_common.cuh"

__global__ void __launch_bounds__(256)
display_render_kernel(const float* __restrict__ cols, const float* __restrict__ stats, const int W,
                      const long long n_frames, const int rows_max, const long long first,
                      const long long step, const int guard, float* __restrict__ norm,
                      float* __restrict__ minmax) {
    const long long r = blockIdx.x;
    const long long t = first + r * step;          // newest frame of this render
    if (t < 0 || t >= n_frames) return;
    // stack range over the history rows (each row's finite min/max came with the PSD row)
    float lo = INFINITY, hi = -INFINITY;
    for (int y = 0; y < rows_max; ++y) {
        const long long f = t - y;
        if (f < 0) break;
        const float4 st = __ldg(reinterpret_cast<const float4*>(stats) + f);
        lo = fminf(lo, st.z);
        hi = fmaxf(hi, st.w);
    }
    float range = hi - lo;
    if (guard && range == 0.f) range = 1.f;
    if (threadIdx.x == 0) {
        minmax[2 * r] = lo;
        minmax[2 * r + 1] = hi;
    }
    float* dst = norm + r * (long long)rows_max * W;
    const int total = rows_max * W;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int y = e / W, c = e - y * W;
        const long long f = t - y;
        float v = __int_as_float(0x7fc00000);
        if (f >= 0) v = (__ldg(cols + f * W + c) - lo) / range;
        dst[e] = v;
    }
}

// draw_spectrogram's numeric part (pyspecsdr.py:418-452): 20th-percentile noise floor (np.percentile,
// linear interpolation between order statistics), display range, clip, ** 0.7, W-column resample.
__global__ void __launch_bounds__(512)
spectrum_normalise_kernel(const float* __restrict__ db, const int n, const long long n_frames, const int W,
                          float* __restrict__ cols, float* __restrict__ range_out) {
    __shared__ unsigned hist[256];
    __shared__ unsigned us[8];
    __shared__ float fmx[16];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (long long f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const float* row = db + f * n;
        // noise floor: np.percentile(row, 20) = x[k] + frac * (x[k+1] - x[k]) on the sorted row
        const double pos = 0.2 * (double)(n - 1);
        const unsigned k = (unsigned)pos;
        const double frac = pos - (double)k;
        unsigned ka, kb;
        row_select2_512(row, n, k, hist, us, ka, kb);
        const double xa = key2f(ka), xb = key2f(kb);
        const double floor_db = xa + frac * (xb - xa);
        float mx = -INFINITY;
        for (int i = tid; i < n; i += 512) mx = fmaxf(mx, row[i]);
        mx = warp_max(mx);
        if (lane == 0) fmx[warp] = mx;
        __syncthreads();
        mx = fmx[0];
        for (int w = 1; w < 16; ++w) mx = fmaxf(mx, fmx[w]);
        const double span = (double)mx - floor_db;                  // :425
        const double dmin = floor_db - span * 0.1;                  // :426
        const double dmax = (double)mx + span * 0.05;               // :427
        if (tid == 0 && range_out) {
            range_out[2 * f] = (float)dmin;
            range_out[2 * f + 1] = (float)dmax;
        }
        const double inv = 1.0 / (dmax - dmin);
        const double step = W > 1 ? (double)(n - 1) / (double)(W - 1) : 0.0;
        for (int c = tid; c < W; c += 512) {
            const double x = (c == W - 1 && W > 1) ? (double)(n - 1) : c * step;
            const int j = min((int)x, n - 1);
            const int j1 = min(j + 1, n - 1);
            const double v0 = pow(fmin(fmax(((double)row[j] - dmin) * inv, 0.0), 1.0), 0.7);    // :442, :445
            const double v1 = pow(fmin(fmax(((double)row[j1] - dmin) * inv, 0.0), 1.0), 0.7);
            cols[f * W + c] = (float)((v1 - v0) * (x - (double)j) + v0);
        }
        __syncthreads();
    }
}

// Glyph / colour index planes of the draw_* functions from normalised values.
// int() in the reference truncates toward zero; NaN (rows older than the history) -> 255.
__global__ void display_quantise_kernel(const float* __restrict__ norm, const long long n, const int kind,
                                        const int H, uint8_t* __restrict__ a, uint8_t* __restrict__ b) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = norm[i];
    if (v != v) {
        a[i] = 255;
        if (b) b[i] = 255;
        return;
    }
    int pa, pb = (int)(v * 5.f);                                      // colour_index, :1388 / :1695
    if (kind == PSS_QUANT_WATERFALL) pa = (v > 0.25f) + (v > 0.5f) + (v > 0.75f);          // '.', '-', '=', '#' :1390-1397
    else if (kind == PSS_QUANT_GRADIENT) pa = (int)(v * 8.f);                              // ' ._-=+*#@' :1691
    else if (kind == PSS_QUANT_PERSISTENCE) pa = (int)((1.f - v) * (float)(H - 1));        // screen row :1556
    else pa = (int)(v * 20.f);                                                             // surface magnitude :1593
    a[i] = (uint8_t)min(max(pa, 0), 254);
    if (b) b[i] = (uint8_t)min(max(pb, 0), 254);
}

void pss_display_release(pss_ctx*) {}

extern "C" {

int pss_display_render_dev(pss_ctx* ctx, const float* cols, const float* stats, int W, int64_t n_frames,
                           int rows_max, int64_t first, int64_t step, int64_t n_renders,
                           int guard_zero_range, float* norm, float* minmax) {
    if (!ctx || !cols || !stats || !norm || !minmax || W < 1 || rows_max < 1 || n_frames < 0 || n_renders < 0)
        return PSS_ERR_ARG;
    if (n_renders == 0) return PSS_OK;
    if (first < 0 || step < 0 || first + (n_renders - 1) * step >= n_frames) return PSS_ERR_ARG;
    if (n_renders > 0x7fffffffLL) return PSS_ERR_ARG;
    display_render_kernel<<<(unsigned)n_renders, 256, 0, ctx->stream>>>(cols, stats, W, n_frames, rows_max, first,
                                                                        step, guard_zero_range, norm, minmax);
    PSS_LAUNCH_CHECK(ctx);
    return PSS_OK;
}

int pss_display_render(pss_ctx* ctx, const float* cols, const float* stats, int W, int64_t n_frames,
                       int rows_max, int64_t first, int64_t step, int64_t n_renders,
                       int guard_zero_range, float* norm, float* minmax) {
    if (!ctx || !cols || !stats || !norm || !minmax || W < 1 || rows_max < 1 || n_frames < 0 || n_renders < 0)
        return PSS_ERR_ARG;
    if (n_renders == 0) return PSS_OK;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t cols_b = (size_t)n_frames * W * 4, st_b = (size_t)n_frames * 16;
    const size_t norm_b = (size_t)n_renders * rows_max * W * 4, mm_b = (size_t)n_renders * 8;
    int rc;
    if ((rc = pss_reserve(ctx, &ctx->d_in, &ctx->d_in_bytes, cols_b))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_aux, &ctx->d_aux_bytes, st_b))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_out, &ctx->d_out_bytes, norm_b))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_aux2, &ctx->d_aux2_bytes, mm_b))) return rc;
    PSS_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, cols, cols_b, cudaMemcpyHostToDevice, ctx->stream));
    PSS_CUDA(ctx, cudaMemcpyAsync(ctx->d_aux, stats, st_b, cudaMemcpyHostToDevice, ctx->stream));
    rc = pss_display_render_dev(ctx, (const float*)ctx->d_in, (const float*)ctx->d_aux, W, n_frames, rows_max, first,
                                step, n_renders, guard_zero_range, (float*)ctx->d_out, (float*)ctx->d_aux2);
    if (rc) return rc;
    PSS_CUDA(ctx, cudaMemcpyAsync(norm, ctx->d_out, norm_b, cudaMemcpyDeviceToHost, ctx->stream));
    PSS_CUDA(ctx, cudaMemcpyAsync(minmax, ctx->d_aux2, mm_b, cudaMemcpyDeviceToHost, ctx->stream));
    PSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PSS_OK;
}


int pss_display_quantise(pss_ctx* ctx, const float* norm, int64_t n, int kind, int H, uint8_t* plane_a,
                         uint8_t* plane_b) {
    if (!ctx || !norm || !plane_a || n < 0 || kind < 0 || kind > 3) return PSS_ERR_ARG;
    if (kind == PSS_QUANT_PERSISTENCE && H < 2) return PSS_ERR_ARG;
    if (n == 0) return PSS_OK;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = pss_reserve(ctx, &ctx->d_in, &ctx->d_in_bytes, (size_t)n * 4))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_out, &ctx->d_out_bytes, (size_t)n * 2))) return rc;
    uint8_t* da = (uint8_t*)ctx->d_out;
    uint8_t* db = plane_b ? da + n : nullptr;
    PSS_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, norm, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    display_quantise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const float*)ctx->d_in, n, kind, H, da, db);
    PSS_LAUNCH_CHECK(ctx);
    PSS_CUDA(ctx, cudaMemcpyAsync(plane_a, da, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (plane_b) PSS_CUDA(ctx, cudaMemcpyAsync(plane_b, db, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    PSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PSS_OK;
}

int pss_spectrum_normalise(pss_ctx* ctx, const float* db, int n_bins, int64_t n_frames, int W, float* cols,
                           float* range) {
    if (!ctx || !db || !cols || n_bins < 2 || W < 1 || n_frames < 0) return PSS_ERR_ARG;
    if (n_frames == 0) return PSS_OK;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t in_b = (size_t)n_frames * n_bins * 4, out_b = (size_t)n_frames * W * 4;
    int rc;
    if ((rc = pss_reserve(ctx, &ctx->d_in, &ctx->d_in_bytes, in_b))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_out, &ctx->d_out_bytes, out_b))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_aux2, &ctx->d_aux2_bytes, (size_t)n_frames * 8))) return rc;
    PSS_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, db, in_b, cudaMemcpyHostToDevice, ctx->stream));
    const long long grid = n_frames < 4LL * ctx->sm_count ? n_frames : 4LL * ctx->sm_count;
    spectrum_normalise_kernel<<<(unsigned)grid, 512, 0, ctx->stream>>>((const float*)ctx->d_in, n_bins, n_frames, W,
                                                                      (float*)ctx->d_out, (float*)ctx->d_aux2);
    PSS_LAUNCH_CHECK(ctx);
    PSS_CUDA(ctx, cudaMemcpyAsync(cols, ctx->d_out, out_b, cudaMemcpyDeviceToHost, ctx->stream));
    if (range) PSS_CUDA(ctx, cudaMemcpyAsync(range, ctx->d_aux2, (size_t)n_frames * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PSS_OK;
}

}  // extern "C"
