// pss_pipeline_c64: one batched main-loop iteration with host buffers (include/pss.h).
#include "pss_common.cuh"

extern "C" int pss_pipeline_c64(pss_ctx* ctx, const float* iq_host, int64_t n_blocks, const pss_pipeline_io* io) {
    if (!ctx || !iq_host || !io || n_blocks < 0) return PSS_ERR_ARG;
    if (io->N_block < 64 || io->N_fft < 512 || io->N_block % io->N_fft || io->W < 1 || io->rows_max < 1)
        return PSS_ERR_ARG;
    if (n_blocks == 0) return PSS_OK;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int fpb = io->N_block / io->N_fft;
    const int64_t n_frames = n_blocks * fpb;
    const size_t n_bins = (size_t)io->N_fft - 4;
    const int out_len = io->plan ? pss_demod_plan_out_len(io->plan) : 0;
    const int ch = io->plan ? pss_demod_plan_channels(io->plan) : 0;
    const size_t sz[7] = {
        (size_t)n_blocks * io->N_block * 8,                 // 0 iq
        (size_t)n_frames * n_bins * 4,                      // 1 db
        (size_t)n_frames * io->W * 4,                       // 2 cols
        (size_t)n_frames * 16,                              // 3 stats
        (size_t)n_blocks * io->rows_max * io->W * 4,        // 4 norm
        (size_t)n_blocks * 8,                               // 5 minmax
        (size_t)n_blocks * out_len * ch * 4 + 16,           // 6 audio
    };
    int rc;
    for (int i = 0; i < 7; ++i)
        if ((rc = pss_reserve(ctx, &ctx->p_buf[i], &ctx->p_bytes[i], sz[i]))) return rc;
    cudaStream_t st = ctx->stream;
    PSS_CUDA(ctx, cudaMemcpyAsync(ctx->p_buf[0], iq_host, sz[0], cudaMemcpyHostToDevice, st));
    pss_psd_out po;
    po.db = (float*)ctx->p_buf[1];
    po.cols = (float*)ctx->p_buf[2];
    po.W = io->W;
    po.stats = (float*)ctx->p_buf[3];
    if ((rc = pss_psd_c64_dev(ctx, (const float*)ctx->p_buf[0], io->N_fft, n_frames, PSS_WINDOW_HAMMING,
                              PSS_EPI_SMOOTH_CLAMP, PSS_PREC_FP64, &po)))
        return rc;
    if ((rc = pss_display_render_dev(ctx, po.cols, po.stats, io->W, n_frames, io->rows_max, fpb - 1, fpb, n_blocks, 0,
                                     (float*)ctx->p_buf[4], (float*)ctx->p_buf[5])))
        return rc;
    if (io->plan)
        if ((rc = pss_demod_c64_dev(ctx, io->plan, (const float*)ctx->p_buf[0], n_blocks, (float*)ctx->p_buf[6])))
            return rc;
    if (io->db) PSS_CUDA(ctx, cudaMemcpyAsync(io->db, ctx->p_buf[1], sz[1], cudaMemcpyDeviceToHost, st));
    if (io->cols) PSS_CUDA(ctx, cudaMemcpyAsync(io->cols, ctx->p_buf[2], sz[2], cudaMemcpyDeviceToHost, st));
    if (io->stats) PSS_CUDA(ctx, cudaMemcpyAsync(io->stats, ctx->p_buf[3], sz[3], cudaMemcpyDeviceToHost, st));
    if (io->norm) PSS_CUDA(ctx, cudaMemcpyAsync(io->norm, ctx->p_buf[4], sz[4], cudaMemcpyDeviceToHost, st));
    if (io->minmax) PSS_CUDA(ctx, cudaMemcpyAsync(io->minmax, ctx->p_buf[5], sz[5], cudaMemcpyDeviceToHost, st));
    if (io->audio && io->plan)
        PSS_CUDA(ctx, cudaMemcpyAsync(io->audio, ctx->p_buf[6], sz[6] - 16, cudaMemcpyDeviceToHost, st));
    PSS_CUDA(ctx, cudaStreamSynchronize(st));
    return PSS_OK;
}
