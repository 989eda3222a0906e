// pss_pipeline_c64: one batched main-loop iteration with host buffers (include/pss.h).
//
// The batch is cut into chunks of blocks; chunk i+1's host->device copy, chunk i's kernels and chunk
// i-1's device->host copies run concurrently on three streams (PCIe is full duplex, the copy engines
// are independent of the SMs), so the call is bound by the larger of the two PCIe directions rather
// than by their sum plus the compute.  The copy streams and events belong to the context.
#include "pss_common.cuh"

#define PIPE_CHUNK_BLOCKS 256     // 256 x 32768 complex64 = 64 MiB per host->device copy

static int pipe_streams(pss_ctx* ctx, pss_pipe_streams** out) {
    pss_pipe_streams& p = ctx->pipe;
    if (!p.ready) {
        if (!p.h2d) PSS_CUDA(ctx, cudaStreamCreateWithFlags(&p.h2d, cudaStreamNonBlocking));
        if (!p.d2h) PSS_CUDA(ctx, cudaStreamCreateWithFlags(&p.d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            if (!p.in_ready[i]) PSS_CUDA(ctx, cudaEventCreateWithFlags(&p.in_ready[i], cudaEventDisableTiming));
            if (!p.in_free[i]) PSS_CUDA(ctx, cudaEventCreateWithFlags(&p.in_free[i], cudaEventDisableTiming));
        }
        if (!p.done) PSS_CUDA(ctx, cudaEventCreateWithFlags(&p.done, cudaEventDisableTiming));
        if (!p.db_free) PSS_CUDA(ctx, cudaEventCreateWithFlags(&p.db_free, cudaEventDisableTiming));
        p.ready = true;
    }
    *out = &p;
    return PSS_OK;
}

extern "C" int pss_pipeline_c64(pss_ctx* ctx, const float* iq_host, int64_t n_blocks, const pss_pipeline_io* io) {
    if (!ctx || !iq_host || !io || n_blocks < 0) return PSS_ERR_ARG;
    if (io->struct_size != sizeof(pss_pipeline_io)) return PSS_ERR_ARG;
    if (io->N_block < 64 || io->N_fft < 512 || io->N_block % io->N_fft || io->W < 1 || io->rows_max < 1)
        return PSS_ERR_ARG;
    // a plan built for another block length would index the staging slots with its own stride
    if (io->plan && pss_demod_plan_block_len(io->plan) != io->N_block) return PSS_ERR_ARG;
    const bool ring = io->display_stream >= 0;
    if (!ring && (io->plane_a || io->plane_b)) return PSS_ERR_ARG;
    if (ring) {
        int rw = 0, rr = 0;
        if (pss_display_geom(ctx, io->display_stream, &rw, &rr) || rw != io->W || rr != io->rows_max) return PSS_ERR_ARG;
    }
    if (n_blocks == 0) return PSS_OK;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int fpb = io->N_block / io->N_fft;
    const int64_t n_frames = n_blocks * fpb;
    const size_t n_bins = (size_t)io->N_fft - 4;
    const int out_len = io->plan ? pss_demod_plan_out_len(io->plan) : 0;
    const int ch = io->plan ? pss_demod_plan_channels(io->plan) : 0;
    int rc0;
    int64_t chunk = PIPE_CHUNK_BLOCKS;
    if ((size_t)chunk * io->N_block * 8 > (256u << 20)) chunk = (256 << 20) / ((int64_t)io->N_block * 8);
    if (chunk < 1) chunk = 1;
    if (chunk > n_blocks) chunk = n_blocks;
    const size_t blk_in = (size_t)io->N_block * 8;
    const size_t cells = (size_t)io->rows_max * io->W;
    // frame moments: a by-product of the PSD pass that WFM plans consume (others ignore them); the fused PSD
    // kernels emit them up to 65536 points.  Own slot: 7..9 are the large-transform scratch.
    const bool use_mom = io->plan && io->N_fft <= 65536;
    if ((rc0 = pss_reserve(ctx, &ctx->p_buf[10], &ctx->p_bytes[10], (size_t)chunk * fpb * 32 + 32))) return rc0;
    const size_t sz[7] = {
        2 * (size_t)chunk * blk_in,                         // 0 iq, two slots
        (size_t)chunk * fpb * n_bins * 4,                   // 1 db of one chunk
        (size_t)n_frames * io->W * 4,                       // 2 cols, whole batch (display history)
        (size_t)n_frames * 16,                              // 3 stats, whole batch
        (size_t)n_blocks * cells * 4,                       // 4 norm
        (size_t)n_blocks * 8,                               // 5 minmax
        (size_t)n_blocks * out_len * ch * 4 + 16,           // 6 audio
    };
    int rc;
    for (int i = 0; i < 7; ++i)
        if ((rc = pss_reserve(ctx, &ctx->p_buf[i], &ctx->p_bytes[i], sz[i]))) return rc;
    const bool planes = io->plane_a || io->plane_b;
    if (planes && (rc = pss_reserve(ctx, &ctx->p_buf[11], &ctx->p_bytes[11], 2 * (size_t)n_blocks * cells))) return rc;
    pss_pipe_streams* ps;
    if ((rc = pipe_streams(ctx, &ps))) return rc;
    cudaStream_t st = ctx->stream;
    char* d_iq = (char*)ctx->p_buf[0];
    float* d_db = (float*)ctx->p_buf[1];
    float* d_cols = (float*)ctx->p_buf[2];
    float* d_stats = (float*)ctx->p_buf[3];
    float* d_norm = (float*)ctx->p_buf[4];
    float* d_mm = (float*)ctx->p_buf[5];
    float* d_audio = (float*)ctx->p_buf[6];
    uint8_t* d_pa = planes ? (uint8_t*)ctx->p_buf[11] : nullptr;
    uint8_t* d_pb = planes ? d_pa + (size_t)n_blocks * cells : nullptr;
    // order the side streams after whatever the caller queued on the context stream
    PSS_CUDA(ctx, cudaEventRecord(ps->done, st));
    PSS_CUDA(ctx, cudaStreamWaitEvent(ps->h2d, ps->done, 0));
    PSS_CUDA(ctx, cudaStreamWaitEvent(ps->d2h, ps->done, 0));
    int slot = 0;
    bool slot_used[2] = {false, false};
    bool db_copy_pending = false;
    for (int64_t b0 = 0; b0 < n_blocks; b0 += chunk, slot ^= 1) {
        const int64_t nb = n_blocks - b0 < chunk ? n_blocks - b0 : chunk;
        const int64_t f0 = b0 * fpb, nf = nb * fpb;
        float* iq_slot = (float*)(d_iq + (size_t)slot * chunk * blk_in);
        if (slot_used[slot]) PSS_CUDA(ctx, cudaStreamWaitEvent(ps->h2d, ps->in_free[slot], 0));
        PSS_CUDA(ctx, cudaMemcpyAsync(iq_slot, (const char*)iq_host + (size_t)b0 * blk_in, (size_t)nb * blk_in,
                                      cudaMemcpyHostToDevice, ps->h2d));
        PSS_CUDA(ctx, cudaEventRecord(ps->in_ready[slot], ps->h2d));
        slot_used[slot] = true;
        PSS_CUDA(ctx, cudaStreamWaitEvent(st, ps->in_ready[slot], 0));
        if (db_copy_pending) PSS_CUDA(ctx, cudaStreamWaitEvent(st, ps->db_free, 0));
        pss_psd_out po{};
        po.struct_size = sizeof po;
        po.db = d_db;
        po.cols = d_cols + f0 * io->W;
        po.W = io->W;
        po.stats = d_stats + f0 * 4;
        po.moments = use_mom ? (double*)ctx->p_buf[10] : nullptr;
        if ((rc = pss_psd_c64_dev(ctx, iq_slot, io->N_fft, nf, PSS_WINDOW_HAMMING, PSS_EPI_SMOOTH_CLAMP,
                                  PSS_PREC_FP64, &po)))
            return rc;
        if (ring) {
            // carried history: the stream's ring holds the rows of earlier chunks and earlier calls
            pss_display_out dout{};
            dout.struct_size = sizeof dout;
            dout.norm = d_norm + (size_t)b0 * cells;
            dout.minmax = d_mm + b0 * 2;
            dout.plane_a = io->plane_a ? d_pa + (size_t)b0 * cells : nullptr;
            dout.plane_b = io->plane_b ? d_pb + (size_t)b0 * cells : nullptr;
            if ((rc = pss_display_accumulate_dev(ctx, io->display_stream, d_cols + f0 * io->W, d_stats + f0 * 4, nf,
                                                 fpb - 1, fpb, nb, &dout)))
                return rc;
        } else {
            // history of a block reaches back into earlier chunks: cols/stats are whole-batch arrays
            if ((rc = pss_display_render_dev(ctx, d_cols, d_stats, io->W, f0 + nf, io->rows_max, f0 + fpb - 1, fpb, nb, 0,
                                             d_norm + (size_t)b0 * cells, d_mm + b0 * 2)))
                return rc;
        }
        if (io->plan)
            if ((rc = pss_demod_c64_dev_moments(ctx, io->plan, iq_slot, nb, d_audio + (size_t)b0 * out_len * ch,
                                                use_mom ? (const double*)ctx->p_buf[10] : nullptr, fpb, io->N_fft)))
                return rc;
        PSS_CUDA(ctx, cudaEventRecord(ps->in_free[slot], st));
        PSS_CUDA(ctx, cudaEventRecord(ps->done, st));
        // results of this chunk go home while the next chunk computes
        PSS_CUDA(ctx, cudaStreamWaitEvent(ps->d2h, ps->done, 0));
        if (io->db) {
            PSS_CUDA(ctx, cudaMemcpyAsync(io->db + (size_t)f0 * n_bins, d_db, (size_t)nf * n_bins * 4,
                                          cudaMemcpyDeviceToHost, ps->d2h));
            PSS_CUDA(ctx, cudaEventRecord(ps->db_free, ps->d2h));
            db_copy_pending = true;
        }
        if (io->cols)
            PSS_CUDA(ctx, cudaMemcpyAsync(io->cols + (size_t)f0 * io->W, d_cols + f0 * io->W, (size_t)nf * io->W * 4,
                                          cudaMemcpyDeviceToHost, ps->d2h));
        if (io->stats)
            PSS_CUDA(ctx, cudaMemcpyAsync(io->stats + (size_t)f0 * 4, d_stats + f0 * 4, (size_t)nf * 16,
                                          cudaMemcpyDeviceToHost, ps->d2h));
        if (io->norm)
            PSS_CUDA(ctx, cudaMemcpyAsync(io->norm + (size_t)b0 * cells, d_norm + (size_t)b0 * cells,
                                          (size_t)nb * cells * 4, cudaMemcpyDeviceToHost, ps->d2h));
        if (io->minmax)
            PSS_CUDA(ctx, cudaMemcpyAsync(io->minmax + b0 * 2, d_mm + b0 * 2, (size_t)nb * 8, cudaMemcpyDeviceToHost,
                                          ps->d2h));
        if (io->plane_a)
            PSS_CUDA(ctx, cudaMemcpyAsync(io->plane_a + (size_t)b0 * cells, d_pa + (size_t)b0 * cells, (size_t)nb * cells,
                                          cudaMemcpyDeviceToHost, ps->d2h));
        if (io->plane_b)
            PSS_CUDA(ctx, cudaMemcpyAsync(io->plane_b + (size_t)b0 * cells, d_pb + (size_t)b0 * cells, (size_t)nb * cells,
                                          cudaMemcpyDeviceToHost, ps->d2h));
        if (io->audio && io->plan)
            PSS_CUDA(ctx, cudaMemcpyAsync(io->audio + (size_t)b0 * out_len * ch, d_audio + (size_t)b0 * out_len * ch,
                                          (size_t)nb * out_len * ch * 4, cudaMemcpyDeviceToHost, ps->d2h));
    }
    PSS_CUDA(ctx, cudaStreamSynchronize(ps->d2h));
    PSS_CUDA(ctx, cudaStreamSynchronize(st));
    return PSS_OK;
}
