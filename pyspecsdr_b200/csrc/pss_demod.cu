// Demodulation kernels (NFM / WFM decimating chain; AM / SSB / RAW frame kernels).
//
// Decimating modes replace demodulate_nfm (signal_processing.py:91-116) and demodulate_wfm
// (:119-176, with iq_correction :46-80).  One CTA owns one block ("frame") at a time:
//   0. (WFM) second moments of I/Q over the block -> the 2x2 correction of iq_correction
//   1. fp32 phase-difference discriminator, computed exactly the way numpy evaluates
//      angle(s[1:] * conj(s[:-1])) on complex64 (fused multiply-add form of the SIMD complex product)
//   2. per chunk of q discriminator samples, fp64 tensor-core products (mma.sync m8n8k4 f64 = DMMA)
//      of the sample window with the response tables built by pyspecsdr_b200/filters.py:
//      forcing of the forward state (pre-filter + Chebyshev forward pass), of the backward state
//      (Chebyshev reversed pass) and of the forward output at the chunk's last sample
//   3. blocked linear scans of the 8/16-dimensional states over the chunk sequence
//   4. y[k] = CB . t + DB * yf, per-block peak normalisation, stereo store.
// No sample-rate recurrence is ever run: only every q-th output of the zero-phase filter exists.
#include <math.h>

#include "pss_common.cuh"

#define DEMOD_THREADS 256
#define TILE_CHUNKS 32
#define EDGE 27

struct DecimDev {
    int mode, N, L, q, n_out, lead, SF, SB, n_body, m_tail, tail_start, tail_len;
    int Bf, Bb;
    int Kp, KS, NT, rows, stride;
    float scale, norm;
    const double *tabF, *AF, *AFB, *AB, *ABB, *MB, *CR, *CB, *head, *tailT, *tailM;
    double DB;
    int tab_in_smem, U_in_smem;
    int off_tile, off_misc, off_tab, off_U;     // byte offsets into dynamic shared memory
    size_t smem_bytes, U_bytes;
};

struct pss_demod_plan {
    int kind = 0, mode = 0, N = 0, out_len = 0, channels = 1;
    DecimDev dec{};
    std::vector<void*> dev_allocs;
    void* U_scratch = nullptr;
    size_t U_scratch_bytes = 0;
    // FIR / SOS plans (pss_demod_frame section)
    float* d_taps_f32 = nullptr;
    int n_taps = 0;
    double* d_sos = nullptr;
    int n_sections = 0;
};

struct IqCorr {
    float inv_q, inv_a, g, inv_c;
};

__device__ __forceinline__ float2 iq_apply(const float2 s, const IqCorr k) {
    // iq_correction (signal_processing.py:55-71) in the reference's float32 op order; the final
    // positive power rescale (:80) does not change a phase difference and is skipped here
    const float zr = __fmul_rn(s.x, k.inv_q), zi = __fmul_rn(s.y, k.inv_q);
    const float i2 = __fmul_rn(k.inv_a, zr);
    const float q2 = __fadd_rn(__fmul_rn(k.g, zr), zi);
    return make_float2(__fmul_rn(i2, k.inv_c), __fmul_rn(q2, k.inv_c));
}

template <bool WFM>
__device__ __forceinline__ float discriminator(const float2* __restrict__ x, const int g, const int L,
                                               const IqCorr k, const float scale) {
    if (g < 0 || g >= L) return 0.f;
    float2 b = __ldg(x + g), a = __ldg(x + g + 1);
    if (WFM) {
        a = iq_apply(a, k);
        b = iq_apply(b, k);
    }
    // numpy complex64 product a * conj(b): re = fma(ar, br, ai*bi), im = fma(ai, br, -(ar*bi))
    const float re = __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y));
    const float im = __fmaf_rn(a.y, b.x, -__fmul_rn(a.x, b.y));
    const float d = atan2f(im, re);
    return WFM ? d : __fmul_rn(d, scale);
}

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// x_{i+1} = A x_i + u_i, i = 0..n-1, over the slot field U[slot*rows + foff + r]; forward walks
// slots 1..n, backward walks slots n..1.  Blocked: every group of S lanes owns one block of B steps.
template <int S>
__device__ void blocked_scan(double* U, const int rows, const int foff, const int n, const bool fwd,
                             const double* __restrict__ A, const double* __restrict__ APow, const int B,
                             const double* x0, double* XS, const int tid) {
    const int r = tid % S, grp = tid / S;
    const int lane_base = (tid & 31) & ~(S - 1);
    const int n_units = (n + B - 1) / B;
    double a[S];
#pragma unroll
    for (int c = 0; c < S; ++c) a[c] = A[r * S + c];
    const int i0 = grp * B;
    // level 1: block-local prefixes from a zero state
    {
        double x = 0.0;
        for (int s = 0; s < B; ++s) {
            const int i = i0 + s;
            const bool act = grp < n_units && i < n;
            const int slot = fwd ? 1 + i : n - i;
            double acc = act ? U[slot * rows + foff + r] : 0.0;
#pragma unroll
            for (int c = 0; c < S; ++c) acc = fma(a[c], __shfl_sync(0xffffffffu, x, lane_base + c), acc);
            x = acc;
            if (act) U[slot * rows + foff + r] = x;
        }
    }
    __syncthreads();
    // level 2: true state at the start of every block
    if (tid < 32) {
        double ap[S];
#pragma unroll
        for (int c = 0; c < S; ++c) ap[c] = APow[r * S + c];
        double X = x0[r];
        for (int b = 0; b < n_units; ++b) {
            if (grp == 0) XS[b * S + r] = X;
            const int ilast = b * B + B - 1;
            const bool more = ilast < n;          // a full block follows
            const int slot = fwd ? 1 + ilast : n - ilast;
            double acc = (more && grp == 0) ? U[slot * rows + foff + r] : 0.0;
#pragma unroll
            for (int c = 0; c < S; ++c) acc = fma(ap[c], __shfl_sync(0xffffffffu, X, lane_base + c), acc);
            X = acc;
        }
    }
    __syncthreads();
    // level 3: add the free response of the block's true start state
    {
        double z = grp < n_units ? XS[grp * S + r] : 0.0;
        for (int s = 0; s < B; ++s) {
            const int i = i0 + s;
            const bool act = grp < n_units && i < n;
            const int slot = fwd ? 1 + i : n - i;
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < S; ++c) acc = fma(a[c], __shfl_sync(0xffffffffu, z, lane_base + c), acc);
            z = acc;
            if (act) U[slot * rows + foff + r] += z;
        }
    }
    __syncthreads();
}

template <int SF>
__global__ void __launch_bounds__(DEMOD_THREADS, 2)
demod_decim_kernel(const DecimDev D, const float2* __restrict__ iq, float* __restrict__ audio,
                   const long long n_frames, double* __restrict__ U_global) {
    constexpr bool WFM = SF == 16;
    constexpr int SB = 8, ROWS = SF + SB + 1, NT = (ROWS + 7) / 8;
    extern __shared__ __align__(16) unsigned char smem[];
    float* tile = reinterpret_cast<float*>(smem + D.off_tile);
    double* misc = reinterpret_cast<double*>(smem + D.off_misc);
    double* XS = misc;                 // [32][16]
    double* dh = misc + 512;           // [28] head discriminator samples
    double* red = dh + 32;             // [32] reduction scratch
    double* tres = red + 32;           // [SB + m_tail] tail result (<= 64)
    double* yout = tres + 64;          // [n_out]
    const double* tab = D.tab_in_smem ? reinterpret_cast<const double*>(smem + D.off_tab) : D.tabF;
    double* U = D.U_in_smem ? reinterpret_cast<double*>(smem + D.off_U)
                            : U_global + (size_t)blockIdx.x * (D.U_bytes / 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = D.q, lead = D.lead, L = D.L, n_body = D.n_body, Kp = D.Kp, stride = D.stride;

    if (D.tab_in_smem) {
        double* ts = reinterpret_cast<double*>(smem + D.off_tab);
        for (int i = tid; i < D.KS * NT * 32; i += DEMOD_THREADS) ts[i] = D.tabF[i];
    }
    __syncthreads();

    for (long long frame = blockIdx.x; frame < n_frames; frame += gridDim.x) {
        const float2* x = iq + frame * D.N;
        IqCorr kc = {1.f, 1.f, 0.f, 1.f};
        if (WFM) {
            // second moments over the block (fp64 accumulation), then iq_correction's estimates
            double sii = 0.0, sqq = 0.0, siq = 0.0;
            for (int i = tid; i < D.N; i += DEMOD_THREADS) {
                const float2 s = __ldg(x + i);
                sii += (double)s.x * (double)s.x;
                sqq += (double)s.y * (double)s.y;
                siq += (double)s.x * (double)s.y;
            }
            sii = warp_sum(sii);
            sqq = warp_sum(sqq);
            siq = warp_sum(siq);
            if (lane == 0) {
                red[warp] = sii;
                red[8 + warp] = sqq;
                red[16 + warp] = siq;
            }
            __syncthreads();
            double a = 0, b = 0, c = 0;
            for (int w = 0; w < DEMOD_THREADS / 32; ++w) {
                a += red[w];
                b += red[8 + w];
                c += red[16 + w];
            }
            const double n = (double)D.N;
            const float q_amp = (float)sqrt(2.0 * b / n);                        // :52
            const double qa = (double)q_amp;
            const float alpha = (float)sqrt(2.0 * a / n / (qa * qa));           // :60
            const float sin_phi = (float)((2.0 / (double)alpha) * (c / n / (qa * qa)));   // :61
            const float cos_phi = sqrtf(1.f - sin_phi * sin_phi);               // :64
            kc.inv_q = 1.f / q_amp;
            kc.inv_a = 1.f / alpha;
            kc.g = -sin_phi / alpha;
            kc.inv_c = 1.f / cos_phi;
            __syncthreads();
        }

        // ---- head: ext[0..27] depends on d[0..27] only
        if (tid <= EDGE) dh[tid] = (double)discriminator<WFM>(x, tid, L, kc, D.scale);
        __syncthreads();
        if (tid <= SF) {
            double acc = 0.0;
            for (int i = 0; i <= EDGE; ++i) acc = fma(D.head[tid * (EDGE + 1) + i], dh[i], acc);
            if (tid < SF) U[tid] = acc;          // slot 0 .F = s_1
            else red[24] = acc;                  // yf at ext index 27
        }

        // ---- body chunks: discriminator tile -> DMMA against the response tables
        const int n_tiles = (n_body + TILE_CHUNKS - 1) / TILE_CHUNKS;
        for (int tl = 0; tl < n_tiles; ++tl) {
            const int j0 = 1 + tl * TILE_CHUNKS;
            for (int c = warp; c < TILE_CHUNKS; c += DEMOD_THREADS / 32) {
                const int j = j0 + c;
                const int gbase = (j - 1) * q + 1 - lead;
                float* rowp = tile + c * stride;
                for (int i = lane; i < Kp; i += 32) {
                    float v = 0.f;
                    if (j <= n_body && i < q + lead) v = discriminator<WFM>(x, gbase + i, L, kc, D.scale);
                    rowp[i] = v;
                }
            }
            __syncthreads();
            {
                const int mt = warp & 3, kg = warp >> 2;
                const int ks0 = kg ? D.KS / 2 : 0, ks1 = kg ? D.KS : D.KS / 2;
                double acc[NT][2];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
                const float* arow = tile + (mt * 8 + (lane >> 2)) * stride + (lane & 3);
                const double* bp = tab + lane;
#pragma unroll 2
                for (int ks = ks0; ks < ks1; ++ks) {
                    const double a = (double)arow[4 * ks];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        dmma_m8n8k4(acc[nt][0], acc[nt][1], a, bp[(ks * NT + nt) * 32]);
                }
                const int j = j0 + mt * 8 + (lane >> 2);
                double* us = U + (size_t)j * ROWS;
                if (kg == 0 && j <= n_body) {
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const int col = nt * 8 + 2 * (lane & 3);
                        if (col < ROWS) us[col] = acc[nt][0];
                        if (col + 1 < ROWS) us[col + 1] = acc[nt][1];
                    }
                }
                __syncthreads();
                if (kg == 1 && j <= n_body) {
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const int col = nt * 8 + 2 * (lane & 3);
                        if (col < ROWS) us[col] += acc[nt][0];
                        if (col + 1 < ROWS) us[col + 1] += acc[nt][1];
                    }
                }
            }
            __syncthreads();
        }

        // ---- forward state scan: slot j .F becomes s_{j+1} (state after chunk j)
        blocked_scan<SF>(U, ROWS, 0, n_body, true, D.AF, D.AFB, D.Bf, U, XS, tid);

        // ---- tail block: reversed-pass state entering chunk n_body, and the last m_tail outputs
        for (int i = tid; i < D.tail_len; i += DEMOD_THREADS)
            tile[i] = discriminator<WFM>(x, D.tail_start + i, L, kc, D.scale);
        __syncthreads();
        {
            const double* s_end = U + (size_t)n_body * ROWS;       // s_{n_body+1}
            for (int rr = warp; rr < SB + D.m_tail; rr += DEMOD_THREADS / 32) {
                double acc = 0.0;
                const double* tr = D.tailT + (size_t)rr * D.tail_len;
                for (int i = lane; i < D.tail_len; i += 32) acc = fma(tr[i], (double)tile[i], acc);
                if (lane < SF) acc = fma(D.tailM[rr * SF + lane], s_end[lane], acc);
                acc = warp_sum(acc);
                if (lane == 0) tres[rr] = acc;
            }
        }
        __syncthreads();
        if (tid < SB) U[(size_t)(n_body + 1) * ROWS + SF + tid] = tres[tid];

        // ---- w_j = MB s_j + vB_j and yf_last_j = CR s_j + r_j  (s_j = slot (j-1) .F)
        for (int j = 1 + tid; j <= n_body; j += DEMOD_THREADS) {
            const double* sj = U + (size_t)(j - 1) * ROWS;
            double* uj = U + (size_t)j * ROWS;
            double s[SF];
#pragma unroll
            for (int c = 0; c < SF; ++c) s[c] = sj[c];
            double yl = uj[SF + SB];
#pragma unroll
            for (int c = 0; c < SF; ++c) yl = fma(D.CR[c], s[c], yl);
#pragma unroll
            for (int rr = 0; rr < SB; ++rr) {
                double w = uj[SF + rr];
#pragma unroll
                for (int c = 0; c < SF; ++c) w = fma(D.MB[rr * SF + c], s[c], w);
                uj[SF + rr] = w;
            }
            uj[SF + SB] = yl;
        }
        __syncthreads();

        // ---- backward state scan: slot j .Bk becomes t_j (reversed-pass state after chunk j)
        blocked_scan<SB>(U, ROWS, SF, n_body, false, D.AB, D.ABB, D.Bb, tres, XS, tid);

        // ---- outputs
        for (int j = tid; j <= n_body; j += DEMOD_THREADS) {
            const double* tn = U + (size_t)(j + 1) * ROWS + SF;      // t_{j+1}
            const double yf = j == 0 ? red[24] : U[(size_t)j * ROWS + SF + SB];
            double y = D.DB * yf;
#pragma unroll
            for (int c = 0; c < SB; ++c) y = fma(D.CB[c], tn[c], y);
            yout[j] = y;
        }
        for (int i = tid; i < D.m_tail; i += DEMOD_THREADS) yout[n_body + 1 + i] = tres[SB + i];
        __syncthreads();
        double mx = 0.0;
        for (int k = tid; k < D.n_out; k += DEMOD_THREADS) mx = fmax(mx, fabs(yout[k]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) red[warp] = mx;
        __syncthreads();
        mx = red[0];
        for (int w = 1; w < DEMOD_THREADS / 32; ++w) mx = fmax(mx, red[w]);
        float2* dst = reinterpret_cast<float2*>(audio) + frame * D.n_out;
        for (int k = tid; k < D.n_out; k += DEMOD_THREADS) {
            const float v = (float)(yout[k] / mx * (double)D.norm);   // audio / max|audio| * 0.95 (:115)
            dst[k] = make_float2(v, v);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------- host side
static int upload(pss_ctx* ctx, pss_demod_plan* pl, const void* src, size_t bytes, const void** dst) {
    void* d = nullptr;
    PSS_CUDA(ctx, cudaMalloc(&d, bytes));
    pl->dev_allocs.push_back(d);
    PSS_CUDA(ctx, cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
    *dst = d;
    return PSS_OK;
}

static int create_decim(pss_ctx* ctx, const pss_demod_desc* d, pss_demod_plan* pl) {
    if (d->SB != 8 || (d->SF != 8 && d->SF != 16)) return PSS_ERR_UNSUPPORTED;
    if (!d->body || !d->AF || !d->AFB || !d->AB || !d->ABB || !d->MB || !d->CR || !d->CB || !d->head ||
        !d->tail_T || !d->tail_M)
        return PSS_ERR_ARG;
    if (d->q < 2 || d->n_body < 0 || d->m_tail < 1 || d->m_tail > 48 || d->tail_len < EDGE + 1) return PSS_ERR_ARG;
    DecimDev& D = pl->dec;
    D.mode = d->mode; D.N = d->N; D.L = d->N - 1; D.q = d->q; D.n_out = d->n_out; D.lead = d->lead;
    D.SF = d->SF; D.SB = d->SB; D.n_body = d->n_body; D.m_tail = d->m_tail;
    D.tail_start = d->tail_start; D.tail_len = d->tail_len;
    D.Bf = d->scan_block_f; D.Bb = d->scan_block_b;
    D.scale = d->scale; D.norm = d->norm; D.DB = d->DB;
    D.rows = D.SF + D.SB + 1;
    D.NT = (D.rows + 7) / 8;
    const int win = D.q + D.lead;
    D.Kp = (win + 3) & ~3;
    D.KS = D.Kp / 4;
    D.stride = D.Kp;
    while (D.stride % 32 != 4) ++D.stride;
    // every block of the scans must fit one lane group
    if (D.Bf < 1 || D.Bb < 1) return PSS_ERR_ARG;
    if ((D.n_body + D.Bf - 1) / D.Bf > DEMOD_THREADS / D.SF) return PSS_ERR_ARG;
    if ((D.n_body + D.Bb - 1) / D.Bb > DEMOD_THREADS / D.SB) return PSS_ERR_ARG;
    if (D.n_out != D.n_body + 1 + D.m_tail) return PSS_ERR_ARG;
    // fragment-ordered body table: [(ks*NT + nt)*32 + lane] = T[row nt*8 + lane/4][i 4ks + lane%4]
    std::vector<double> frag((size_t)D.KS * D.NT * 32, 0.0);
    for (int ks = 0; ks < D.KS; ++ks)
        for (int nt = 0; nt < D.NT; ++nt)
            for (int l = 0; l < 32; ++l) {
                const int row = nt * 8 + l / 4, i = 4 * ks + l % 4;
                if (row < D.rows && i < win) frag[((size_t)ks * D.NT + nt) * 32 + l] = d->body[(size_t)row * win + i];
            }
    int rc;
    const void* p;
    if ((rc = upload(ctx, pl, frag.data(), frag.size() * 8, &p))) return rc; D.tabF = (const double*)p;
    if ((rc = upload(ctx, pl, d->AF, (size_t)D.SF * D.SF * 8, &p))) return rc; D.AF = (const double*)p;
    if ((rc = upload(ctx, pl, d->AFB, (size_t)D.SF * D.SF * 8, &p))) return rc; D.AFB = (const double*)p;
    if ((rc = upload(ctx, pl, d->AB, (size_t)D.SB * D.SB * 8, &p))) return rc; D.AB = (const double*)p;
    if ((rc = upload(ctx, pl, d->ABB, (size_t)D.SB * D.SB * 8, &p))) return rc; D.ABB = (const double*)p;
    if ((rc = upload(ctx, pl, d->MB, (size_t)D.SB * D.SF * 8, &p))) return rc; D.MB = (const double*)p;
    if ((rc = upload(ctx, pl, d->CR, (size_t)D.SF * 8, &p))) return rc; D.CR = (const double*)p;
    if ((rc = upload(ctx, pl, d->CB, (size_t)D.SB * 8, &p))) return rc; D.CB = (const double*)p;
    if ((rc = upload(ctx, pl, d->head, (size_t)(D.SF + 1) * (EDGE + 1) * 8, &p))) return rc; D.head = (const double*)p;
    if ((rc = upload(ctx, pl, d->tail_T, (size_t)(D.SB + D.m_tail) * D.tail_len * 8, &p))) return rc; D.tailT = (const double*)p;
    if ((rc = upload(ctx, pl, d->tail_M, (size_t)(D.SB + D.m_tail) * D.SF * 8, &p))) return rc; D.tailM = (const double*)p;

    // shared-memory layout: tile + misc always; table and state slots when they fit in 113 KB
    const size_t budget = 113 * 1024;
    size_t tile_b = (size_t)TILE_CHUNKS * D.stride * 4;
    if (tile_b < (size_t)D.tail_len * 4) tile_b = (size_t)D.tail_len * 4;
    tile_b = (tile_b + 15) & ~(size_t)15;
    const size_t misc_b = ((size_t)(512 + 32 + 32 + 64 + D.n_out) * 8 + 15) & ~(size_t)15;
    const size_t tab_b = frag.size() * 8;
    D.U_bytes = (((size_t)(D.n_body + 2) * D.rows * 8) + 15) & ~(size_t)15;
    size_t used = 0;
    D.off_tile = (int)used; used += tile_b;
    D.off_misc = (int)used; used += misc_b;
    if (used > budget) return PSS_ERR_UNSUPPORTED;
    D.tab_in_smem = used + tab_b <= budget;
    if (D.tab_in_smem) { D.off_tab = (int)used; used += tab_b; }
    D.U_in_smem = used + D.U_bytes <= budget;
    if (D.U_in_smem) { D.off_U = (int)used; used += D.U_bytes; }
    D.smem_bytes = used;
    pl->out_len = D.n_out;
    pl->channels = 2;
    return PSS_OK;
}

static int launch_decim(pss_ctx* ctx, pss_demod_plan* pl, const float* iq, int64_t n_frames, float* audio) {
    DecimDev& D = pl->dec;
    long long grid = 2LL * ctx->sm_count;
    if (grid > n_frames) grid = n_frames;
    if (!D.U_in_smem) {
        const size_t need = (size_t)grid * D.U_bytes;
        if (pl->U_scratch_bytes < need) {
            if (pl->U_scratch) {
                PSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                PSS_CUDA(ctx, cudaFree(pl->U_scratch));
                pl->U_scratch = nullptr;
            }
            PSS_CUDA(ctx, cudaMalloc(&pl->U_scratch, need));
            pl->U_scratch_bytes = need;
        }
    }
    if (D.SF == 8) {
        auto k = demod_decim_kernel<8>;
        PSS_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D.smem_bytes));
        k<<<(unsigned)grid, DEMOD_THREADS, D.smem_bytes, ctx->stream>>>(D, (const float2*)iq, audio, n_frames,
                                                                         (double*)pl->U_scratch);
    } else {
        auto k = demod_decim_kernel<16>;
        PSS_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D.smem_bytes));
        k<<<(unsigned)grid, DEMOD_THREADS, D.smem_bytes, ctx->stream>>>(D, (const float2*)iq, audio, n_frames,
                                                                          (double*)pl->U_scratch);
    }
    PSS_LAUNCH_CHECK(ctx);
    return PSS_OK;
}

void pss_demod_release(pss_ctx*) {}

extern "C" {

int pss_demod_plan_create(pss_ctx* ctx, const pss_demod_desc* desc, pss_demod_plan** out) {
    if (!ctx || !desc || !out || desc->N < 2) return PSS_ERR_ARG;
    *out = nullptr;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    pss_demod_plan* pl = new (std::nothrow) pss_demod_plan();
    if (!pl) return PSS_ERR_NOMEM;
    pl->kind = desc->kind;
    pl->mode = desc->mode;
    pl->N = desc->N;
    int rc = PSS_ERR_UNSUPPORTED;
    if (desc->kind == PSS_PLAN_DECIM) rc = create_decim(ctx, desc, pl);
    if (rc != PSS_OK) {
        pss_demod_plan_destroy(ctx, pl);
        return rc;
    }
    *out = pl;
    return PSS_OK;
}

void pss_demod_plan_destroy(pss_ctx* ctx, pss_demod_plan* pl) {
    if (!pl) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    for (void* p : pl->dev_allocs) cudaFree(p);
    cudaFree(pl->U_scratch);
    cudaFree(pl->d_taps_f32);
    cudaFree(pl->d_sos);
    delete pl;
}

int pss_demod_plan_out_len(const pss_demod_plan* pl) { return pl ? pl->out_len : 0; }
int pss_demod_plan_channels(const pss_demod_plan* pl) { return pl ? pl->channels : 0; }

int pss_demod_c64_dev(pss_ctx* ctx, pss_demod_plan* pl, const float* iq, int64_t n_frames, float* audio) {
    if (!ctx || !pl || !iq || !audio || n_frames < 0) return PSS_ERR_ARG;
    if (n_frames == 0) return PSS_OK;
    if (pl->kind == PSS_PLAN_DECIM) return launch_decim(ctx, pl, iq, n_frames, audio);
    return PSS_ERR_UNSUPPORTED;
}

int pss_demod_c64(pss_ctx* ctx, pss_demod_plan* pl, const float* iq, int64_t n_frames, float* audio) {
    if (!ctx || !pl || !iq || !audio || n_frames < 0) return PSS_ERR_ARG;
    if (n_frames == 0) return PSS_OK;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t in_b = (size_t)n_frames * pl->N * 8;
    const size_t out_b = (size_t)n_frames * pl->out_len * pl->channels * 4;
    int rc;
    if ((rc = pss_reserve(ctx, &ctx->d_in, &ctx->d_in_bytes, in_b))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_out, &ctx->d_out_bytes, out_b))) return rc;
    PSS_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, iq, in_b, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = pss_demod_c64_dev(ctx, pl, (const float*)ctx->d_in, n_frames, (float*)ctx->d_out))) return rc;
    PSS_CUDA(ctx, cudaMemcpyAsync(audio, ctx->d_out, out_b, cudaMemcpyDeviceToHost, ctx->stream));
    PSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PSS_OK;
}

}  // extern "C"
