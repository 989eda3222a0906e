// Context lifetime, error plumbing and the host-pointer wrappers of the C ABI (include/pss.h).
#include "pss_common.cuh"

int pss_fail_cuda(pss_ctx* ctx, cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    if (ctx) ctx->last_error = buf;
    (void)cudaGetLastError();   // clear the sticky non-fatal error state
    return PSS_ERR_CUDA;
}

int pss_reserve(pss_ctx* ctx, void** p, size_t* have, size_t need) {
    if (*have >= need) return PSS_OK;
    if (*p) {
        PSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        PSS_CUDA(ctx, cudaFree(*p));
        *p = nullptr;
        *have = 0;
    }
    size_t want = need + need / 4;
    if (cudaMalloc(p, want) != cudaSuccess) {
        (void)cudaGetLastError();
        *p = nullptr;
        ctx->last_error = "cudaMalloc failed for scratch";
        return PSS_ERR_NOMEM;
    }
    *have = want;
    return PSS_OK;
}

bool pss_use_pdl() {
    static const bool on = getenv("PSS_PDL") && atoi(getenv("PSS_PDL")) != 0;
    return on;
}

void pss_demod_release(pss_ctx* ctx);     // pss_demod.cu
void pss_display_release(pss_ctx* ctx);   // pss_display.cu

extern "C" {

int pss_version(void) { return PSS_VERSION; }

const char* pss_strerror(int status) {
    switch (status) {
        case PSS_OK: return "ok";
        case PSS_ERR_ARG: return "invalid argument";
        case PSS_ERR_CUDA: return "CUDA error (see pss_last_error)";
        case PSS_ERR_NOMEM: return "out of device memory";
        case PSS_ERR_UNSUPPORTED: return "unsupported configuration";
        case PSS_ERR_NODEVICE: return "no usable CUDA device";
        default: return "unknown status";
    }
}

int pss_init(int device, pss_ctx** out) {
    if (!out) return PSS_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        (void)cudaGetLastError();
        return PSS_ERR_NODEVICE;
    }
    if (device < 0 || device >= count) return PSS_ERR_ARG;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PSS_ERR_NODEVICE;
    if (prop.major != 10) return PSS_ERR_NODEVICE;   // the cubin is sm_100a only
    if (cudaSetDevice(device) != cudaSuccess) return PSS_ERR_NODEVICE;
    pss_ctx* ctx = new (std::nothrow) pss_ctx();
    if (!ctx) return PSS_ERR_NOMEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return PSS_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return PSS_OK;
}

void pss_destroy(pss_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    pss_demod_release(ctx);
    pss_display_release(ctx);
    for (auto& kv : ctx->fft_tables) {
        cudaFree(kv.second.twiddle);
        for (void* w : kv.second.window) cudaFree(w);
    }
    cudaFree(ctx->d_in);
    cudaFree(ctx->d_out);
    cudaFree(ctx->d_aux);
    cudaFree(ctx->d_aux2);
    for (void* b : ctx->p_buf) cudaFree(b);
    for (auto& kv : ctx->large_tables) {
        pss_large_tables& lt = kv.second;
        cudaFree(lt.tw1); cudaFree(lt.thi); cudaFree(lt.tlo); cudaFree(lt.twN);
        for (void* w : lt.window) cudaFree(w);
    }
    cudaFree(ctx->hann_periodic);
    pss_pipe_streams& ps = ctx->pipe;
    if (ps.h2d) cudaStreamDestroy(ps.h2d);
    if (ps.d2h) cudaStreamDestroy(ps.d2h);
    for (int i = 0; i < 2; ++i) {
        if (ps.in_ready[i]) cudaEventDestroy(ps.in_ready[i]);
        if (ps.in_free[i]) cudaEventDestroy(ps.in_free[i]);
    }
    if (ps.done) cudaEventDestroy(ps.done);
    if (ps.db_free) cudaEventDestroy(ps.db_free);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* pss_last_error(const pss_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int pss_set_stream(pss_ctx* ctx, void* cuda_stream) {
    if (!ctx) return PSS_ERR_ARG;
    ctx->stream = (cudaStream_t)cuda_stream;
    return PSS_OK;
}

int pss_use_own_stream(pss_ctx* ctx) {
    if (!ctx) return PSS_ERR_ARG;
    ctx->stream = ctx->own_stream;
    return PSS_OK;
}

int pss_sync(pss_ctx* ctx) {
    if (!ctx) return PSS_ERR_ARG;
    PSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PSS_OK;
}

int64_t pss_kernel_launches(const pss_ctx* ctx) { return ctx ? ctx->launches : 0; }

void* pss_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return p;
}

void pss_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------ host-pointer PSD / scanner
int pss_psd_c64(pss_ctx* ctx, const float* iq, int N, int64_t n_frames, int window, int epilogue,
                int precision, const pss_psd_out* out) {
    if (!ctx || !iq || !out || n_frames < 0 || N <= 0) return PSS_ERR_ARG;
    if (out->struct_size != sizeof(pss_psd_out)) return PSS_ERR_ARG;
    if (n_frames == 0) return PSS_OK;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n_out = epilogue == PSS_EPI_SMOOTH_CLAMP ? (size_t)N - 4 : (size_t)N;
    const size_t in_b = (size_t)n_frames * N * 8;
    const size_t db_b = out->db ? (size_t)n_frames * n_out * 4 : 0;
    const size_t cols_b = out->cols ? (size_t)n_frames * out->W * 4 : 0;
    const size_t st_b = out->stats ? (size_t)n_frames * 16 : 0;
    int rc;
    if ((rc = pss_reserve(ctx, &ctx->d_in, &ctx->d_in_bytes, in_b))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_out, &ctx->d_out_bytes, db_b + 16))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_aux, &ctx->d_aux_bytes, cols_b + 16))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_aux2, &ctx->d_aux2_bytes, st_b + 16))) return rc;
    PSS_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, iq, in_b, cudaMemcpyHostToDevice, ctx->stream));
    pss_psd_out dev = *out;
    dev.db = out->db ? (float*)ctx->d_out : nullptr;
    dev.cols = out->cols ? (float*)ctx->d_aux : nullptr;
    dev.stats = out->stats ? (float*)ctx->d_aux2 : nullptr;
    dev.moments = nullptr;                 // a device-side by-product; not copied back by the host variant
    rc = pss_psd_c64_dev(ctx, (const float*)ctx->d_in, N, n_frames, window, epilogue, precision, &dev);
    if (rc) return rc;
    if (db_b) PSS_CUDA(ctx, cudaMemcpyAsync(out->db, dev.db, db_b, cudaMemcpyDeviceToHost, ctx->stream));
    if (cols_b) PSS_CUDA(ctx, cudaMemcpyAsync(out->cols, dev.cols, cols_b, cudaMemcpyDeviceToHost, ctx->stream));
    if (st_b) PSS_CUDA(ctx, cudaMemcpyAsync(out->stats, dev.stats, st_b, cudaMemcpyDeviceToHost, ctx->stream));
    PSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PSS_OK;
}

int pss_scan_c64(pss_ctx* ctx, const float* iq, int N, int64_t n_steps, int use_abs, float thr_db,
                 float* peak_db, int32_t* count_above, float* db_rows) {
    if (!ctx || !iq || !peak_db || !count_above || n_steps < 0 || N <= 0) return PSS_ERR_ARG;
    if (n_steps == 0) return PSS_OK;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t in_b = (size_t)n_steps * N * 8;
    const size_t db_b = db_rows ? (size_t)n_steps * N * 4 : 0;
    int rc;
    if ((rc = pss_reserve(ctx, &ctx->d_in, &ctx->d_in_bytes, in_b))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_out, &ctx->d_out_bytes, db_b + 16))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_aux, &ctx->d_aux_bytes, (size_t)n_steps * 4))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_aux2, &ctx->d_aux2_bytes, (size_t)n_steps * 4))) return rc;
    PSS_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, iq, in_b, cudaMemcpyHostToDevice, ctx->stream));
    rc = pss_scan_c64_dev(ctx, (const float*)ctx->d_in, N, n_steps, use_abs, thr_db, (float*)ctx->d_aux,
                          (int32_t*)ctx->d_aux2, db_rows ? (float*)ctx->d_out : nullptr);
    if (rc) return rc;
    PSS_CUDA(ctx, cudaMemcpyAsync(peak_db, ctx->d_aux, (size_t)n_steps * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PSS_CUDA(ctx, cudaMemcpyAsync(count_above, ctx->d_aux2, (size_t)n_steps * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (db_b) PSS_CUDA(ctx, cudaMemcpyAsync(db_rows, ctx->d_out, db_b, cudaMemcpyDeviceToHost, ctx->stream));
    PSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PSS_OK;
}

int pss_classify_c64(pss_ctx* ctx, const float* iq, int N, int64_t n_blocks, double fs, double* features,
                     int32_t* label) {
    if (!ctx || !iq || !features || !label || n_blocks < 0 || N <= 0) return PSS_ERR_ARG;
    if (n_blocks == 0) return PSS_OK;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t in_b = (size_t)n_blocks * N * 8;
    int rc;
    if ((rc = pss_reserve(ctx, &ctx->d_in, &ctx->d_in_bytes, in_b))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_aux, &ctx->d_aux_bytes, (size_t)n_blocks * 32))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_aux2, &ctx->d_aux2_bytes, (size_t)n_blocks * 4))) return rc;
    PSS_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, iq, in_b, cudaMemcpyHostToDevice, ctx->stream));
    rc = pss_classify_c64_dev(ctx, (const float*)ctx->d_in, N, n_blocks, fs, (double*)ctx->d_aux, (int32_t*)ctx->d_aux2);
    if (rc) return rc;
    PSS_CUDA(ctx, cudaMemcpyAsync(features, ctx->d_aux, (size_t)n_blocks * 32, cudaMemcpyDeviceToHost, ctx->stream));
    PSS_CUDA(ctx, cudaMemcpyAsync(label, ctx->d_aux2, (size_t)n_blocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PSS_OK;
}

}  // extern "C"

