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
#define EDGE 27

struct DecimDev {
    int mode, N, L, q, n_out, lead, SF, SB, n_body, m_tail, tail_start, tail_len;
    int Bf, Bb;
    int Kp, KS, NT, rows, T, nbuf, tile_floats;
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

// atan2f replacement: branch-free, |error| < 1.5e-7 rad (minimax degree-8 polynomial in t^2 for
// atan(t)/t on [0,1], max fp32 evaluation error 9.3e-8, plus a 2-ulp fast division).  The reference's
// np.angle is numpy/SVML arctan2 in float32, itself 1-4 ulp; parity is a tolerance (1e-5 RMS on the
// normalised audio), not bit equality.  Signs follow atan2: result carries the sign of `im`
// (including -0.0), and is pi-mirrored when `re` is negative; atan2(0, 0) = 0.
__device__ __forceinline__ float fast_atan2f(const float im, const float re) {
    const float ax = fabsf(re), ay = fabsf(im);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float t = __fdividef(mn, mx);
    t = mx == 0.f ? 0.f : t;
    const float z = t * t;
    float p = 2.456712816e-03f;
    p = fmaf(p, z, -1.440130838e-02f);
    p = fmaf(p, z, 3.978113781e-02f);
    p = fmaf(p, z, -7.234849502e-02f);
    p = fmaf(p, z, 1.049894197e-01f);
    p = fmaf(p, z, -1.416122798e-01f);
    p = fmaf(p, z, 1.998590658e-01f);
    p = fmaf(p, z, -3.333259701e-01f);
    p = fmaf(p, z, 9.999998864e-01f);
    float r = p * t;
    r = ay > ax ? 1.57079632679489662f - r : r;
    r = re < 0.f ? 3.14159265358979324f - r : r;
    return copysignf(r, im);
}

// d = angle(a * conj(b)) the way numpy evaluates it on complex64:
// re = fma(ar, br, ai*bi), im = fma(ai, br, -(ar*bi))  (SIMD fused multiply-add/sub complex product)
template <bool WFM>
__device__ __forceinline__ float disc_core(const float2 a, const float2 b, const float scale) {
    const float re = __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y));
    const float im = __fmaf_rn(a.y, b.x, -__fmul_rn(a.x, b.y));
    const float d = fast_atan2f(im, re);
    return WFM ? d : __fmul_rn(d, scale);
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
    return disc_core<WFM>(a, b, scale);
}

// One contiguous tile of discriminator samples d[g0 .. g0+E) (zero outside [0, L)) is produced in two
// halves so that the global loads of tile t+1 are in flight while tile t feeds the tensor pipe:
//   tile_load : every warp-iteration loads 32 consecutive IQ samples (one 8-byte load per lane) into
//               registers;
//   tile_store: IQ-correct (WFM), take the neighbour from the previous lane, discriminate, store the
//               31 outputs of the iteration.  Each sample is loaded and corrected once.
template <int PF>
__device__ __forceinline__ void tile_load(float2 (&pf)[PF], const float2* __restrict__ x, const int g0,
                                          const int n_wi, const int N, const int warp, const int lane) {
#pragma unroll
    for (int it = 0; it < PF; ++it) {
        const int wi = warp + it * (DEMOD_THREADS / 32);
        const int gi = g0 + 31 * wi + lane;
        pf[it] = make_float2(0.f, 0.f);
        if (wi < n_wi && gi >= 0 && gi < N) pf[it] = __ldg(x + gi);
    }
}

template <bool WFM, int PF>
__device__ __forceinline__ void tile_store(float* __restrict__ buf, const float2 (&pf)[PF],
                                           const float2* __restrict__ x, const int g0, const int E,
                                           const int n_wi, const int N, const IqCorr k, const float scale,
                                           const int warp, const int lane) {
    const int L = N - 1;
#pragma unroll
    for (int it = 0; it < PF; ++it) {
        const int wi = warp + it * (DEMOD_THREADS / 32);
        float2 cur = pf[it];
        if (WFM) cur = iq_apply(cur, k);
        float2 prev;
        prev.x = __shfl_up_sync(0xffffffffu, cur.x, 1);
        prev.y = __shfl_up_sync(0xffffffffu, cur.y, 1);
        const int e = 31 * wi + lane - 1;
        const int g = g0 + e;
        float d = disc_core<WFM>(cur, prev, scale);
        if (g < 0 || g >= L) d = 0.f;
        if (wi < n_wi && lane > 0 && e < E) buf[e] = d;
    }
    // tiles larger than PF iterations per warp (very large q): finish without the register prefetch
    for (int wi = warp + PF * (DEMOD_THREADS / 32); wi < n_wi; wi += DEMOD_THREADS / 32) {
        const int gi = g0 + 31 * wi + lane;
        float2 cur = make_float2(0.f, 0.f);
        if (gi >= 0 && gi < N) cur = __ldg(x + gi);
        if (WFM) cur = iq_apply(cur, k);
        float2 prev;
        prev.x = __shfl_up_sync(0xffffffffu, cur.x, 1);
        prev.y = __shfl_up_sync(0xffffffffu, cur.y, 1);
        const int e = 31 * wi + lane - 1;
        const int g = g0 + e;
        float d = disc_core<WFM>(cur, prev, scale);
        if (g < 0 || g >= L) d = 0.f;
        if (lane > 0 && e < E) buf[e] = d;
    }
}

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// One row of y = A x with x spread over the S lanes of a group (4 independent partial sums keep the
// dependent fp64 chain short).
template <int S>
__device__ __forceinline__ double matvec_row(const double (&a)[S], const double x, const int lane_base) {
    double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
    for (int c = 0; c < S; c += 4) {
        p0 = fma(a[c], __shfl_sync(0xffffffffu, x, lane_base + c), p0);
        p1 = fma(a[c + 1], __shfl_sync(0xffffffffu, x, lane_base + c + 1), p1);
        p2 = fma(a[c + 2], __shfl_sync(0xffffffffu, x, lane_base + c + 2), p2);
        p3 = fma(a[c + 3], __shfl_sync(0xffffffffu, x, lane_base + c + 3), p3);
    }
    return (p0 + p1) + (p2 + p3);
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
            const double u = act ? U[slot * rows + foff + r] : 0.0;
            x = u + matvec_row<S>(a, x, lane_base);
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
            const double u = (more && grp == 0) ? U[slot * rows + foff + r] : 0.0;
            X = u + matvec_row<S>(ap, X, lane_base);
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
            z = matvec_row<S>(a, z, lane_base);
            if (act) U[slot * rows + foff + r] += z;
        }
    }
    __syncthreads();
}

template <int SF, int T, int NBUF>
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
    const int q = D.q, lead = D.lead, L = D.L, n_body = D.n_body, Kp = D.Kp;

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
            // (per-thread float partials over N/256 samples, combined in fp64: the same order of
            // rounding error as numpy's own float32 pairwise means at :52, :60, :61)
            float fii[4] = {0.f, 0.f, 0.f, 0.f}, fqq[4] = {0.f, 0.f, 0.f, 0.f}, fiq[4] = {0.f, 0.f, 0.f, 0.f};
            int i = tid;
            for (; i + 15 * DEMOD_THREADS < D.N; i += 16 * DEMOD_THREADS) {
                float2 v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) v[u] = __ldg(x + i + u * DEMOD_THREADS);   // 16 loads in flight
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    fii[u & 3] = fmaf(v[u].x, v[u].x, fii[u & 3]);
                    fqq[u & 3] = fmaf(v[u].y, v[u].y, fqq[u & 3]);
                    fiq[u & 3] = fmaf(v[u].x, v[u].y, fiq[u & 3]);
                }
            }
            for (; i < D.N; i += DEMOD_THREADS) {
                const float2 s = __ldg(x + i);
                fii[0] = fmaf(s.x, s.x, fii[0]);
                fqq[0] = fmaf(s.y, s.y, fqq[0]);
                fiq[0] = fmaf(s.x, s.y, fiq[0]);
            }
            double sii = ((double)fii[0] + (double)fii[1]) + ((double)fii[2] + (double)fii[3]);
            double sqq = ((double)fqq[0] + (double)fqq[1]) + ((double)fqq[2] + (double)fqq[3]);
            double siq = ((double)fiq[0] + (double)fiq[1]) + ((double)fiq[2] + (double)fiq[3]);
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

        // ---- clear the state slots (the tile products are accumulated with atomics)
        for (int i = tid; i < (n_body + 2) * ROWS; i += DEMOD_THREADS) U[i] = 0.0;
        // ---- head: ext[0..27] depends on d[0..27] only
        if (tid <= EDGE) dh[tid] = (double)discriminator<WFM>(x, tid, L, kc, D.scale);
        const int n_tiles = (n_body + T - 1) / T;
        const int E = (T - 1) * q + Kp;                          // samples one tile's windows touch
        const int n_wi = (E + 30) / 31;                          // warp-iterations of 31 outputs each
        constexpr int PF = T == 32 ? 16 : 8;
        float2 pf[PF];
        if (n_tiles > 0) {
            tile_load<PF>(pf, x, 1 - lead, n_wi, D.N, warp, lane);
            tile_store<WFM, PF>(tile, pf, x, 1 - lead, E, n_wi, D.N, kc, D.scale, warp, lane);
        }
        __syncthreads();
        if (tid <= SF) {
            double acc = 0.0;
            for (int i = 0; i <= EDGE; ++i) acc = fma(D.head[tid * (EDGE + 1) + i], dh[i], acc);
            if (tid < SF) U[tid] = acc;          // slot 0 .F = s_1
            else red[24] = acc;                  // yf at ext index 27
        }

        // ---- body chunks: discriminator tile -> DMMA against the response tables.  The IQ loads of
        // tile t+1 are issued before tile t is consumed by the tensor pipe; one barrier per tile.
        for (int tl = 0; tl < n_tiles; ++tl) {
            const int j0 = 1 + tl * T;
            const float* cur = tile + (NBUF == 2 ? (tl & 1) * D.tile_floats : 0);
            const int g_next = (j0 + T - 1) * q + 1 - lead;
            const bool more = tl + 1 < n_tiles;
            if (more) tile_load<PF>(pf, x, g_next, n_wi, D.N, warp, lane);
            {
                // work split over the 8 warps: a unit = (m-tile of 8 chunks, n-tile of 8 table rows)
                // over the whole window.  MT*NT is 8 or 16 -> whole units per warp, plain stores.
                // MT=4, NT=3 (NFM, 32-chunk tiles): n-tiles 0/1 whole, the last n-tile (it only
                // carries the yf-forcing row) split in two K halves -> atomics on that one column.
                constexpr int MT = T / 8, UNITS = MT * NT, NW = DEMOD_THREADS / 32;
                constexpr bool SPLIT_LAST = (UNITS % NW) != 0 && MT == 4 && NT == 3;
                constexpr int UPW = SPLIT_LAST ? 1 : (UNITS + NW - 1) / NW;      // whole units per warp
#pragma unroll
                for (int uu = 0; uu < UPW + (SPLIT_LAST ? 1 : 0); ++uu) {
                    int mt, nt, ks0 = 0, ks1 = D.KS;
                    bool atomic = false;
                    if (SPLIT_LAST) {
                        mt = warp % MT;
                        if (uu == 0) nt = warp / MT;
                        else {
                            nt = 2;
                            atomic = true;
                            ks0 = (warp / MT) ? D.KS / 2 : 0;
                            ks1 = (warp / MT) ? D.KS : D.KS / 2;
                        }
                    } else {
                        const int unit = warp * UPW + uu;
                        if (unit >= UNITS) break;
                        mt = unit % MT;
                        nt = unit / MT;
                    }
                    double c0 = 0.0, c1 = 0.0;
                    const float* arow = cur + (mt * 8 + (lane >> 2)) * q + (lane & 3);
                    const double* bp = tab + nt * 32 + lane;
#pragma unroll 4
                    for (int ks = ks0; ks < ks1; ++ks)
                        dmma_m8n8k4(c0, c1, (double)arow[4 * ks], bp[ks * NT * 32]);
                    const int j = j0 + mt * 8 + (lane >> 2);
                    const int col = nt * 8 + 2 * (lane & 3);
                    if (j <= n_body) {
                        double* us = U + (size_t)j * ROWS;
                        if (atomic) {
                            if (col < ROWS) atomicAdd(us + col, c0);
                            if (col + 1 < ROWS) atomicAdd(us + col + 1, c1);
                        } else {
                            if (col < ROWS) us[col] = c0;
                            if (col + 1 < ROWS) us[col + 1] = c1;
                        }
                    }
                }
            }
            if (NBUF == 1) __syncthreads();
            if (more)
                tile_store<WFM, PF>(tile + (NBUF == 2 ? ((tl + 1) & 1) * D.tile_floats : 0), pf, x, g_next, E, n_wi,
                                    D.N, kc, D.scale, warp, lane);
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

    // shared-memory layout: tile(s) + misc always; table and state slots when they fit in 113 KB.
    // Preference: 32-chunk double-buffered tiles, then smaller / single-buffered ones.
    const size_t budget = 113 * 1024;
    const size_t misc_b = ((size_t)(512 + 32 + 32 + 64 + D.n_out) * 8 + 15) & ~(size_t)15;
    const size_t tab_b = frag.size() * 8;
    D.U_bytes = (((size_t)(D.n_body + 2) * D.rows * 8) + 15) & ~(size_t)15;
    const int cand[4][2] = {{32, 2}, {16, 2}, {32, 1}, {16, 1}};
    int pick = -1;
    for (int c = 0; c < 4 && pick < 0; ++c) {
        size_t tf = (size_t)(cand[c][0] - 1) * D.q + D.Kp;
        if (tf * cand[c][1] < (size_t)D.tail_len) tf = ((size_t)D.tail_len + cand[c][1] - 1) / cand[c][1];
        tf = (tf + 3) & ~(size_t)3;
        const size_t tot = tf * 4 * cand[c][1] + misc_b + tab_b + D.U_bytes;
        if (tot <= budget || c == 3) {
            pick = c;
            D.T = cand[c][0];
            D.nbuf = cand[c][1];
            D.tile_floats = (int)tf;
        }
    }
    size_t used = 0;
    D.off_tile = (int)used; used += (size_t)D.tile_floats * 4 * D.nbuf;
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
#define DECIM_LAUNCH(SFv, Tv, NBv)                                                                         \
    do {                                                                                                   \
        auto k = demod_decim_kernel<SFv, Tv, NBv>;                                                         \
        PSS_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D.smem_bytes)); \
        k<<<(unsigned)grid, DEMOD_THREADS, D.smem_bytes, ctx->stream>>>(D, (const float2*)iq, audio, n_frames, \
                                                                         (double*)pl->U_scratch);          \
    } while (0)
    if (D.SF == 8) {
        if (D.T == 32 && D.nbuf == 2) DECIM_LAUNCH(8, 32, 2);
        else if (D.T == 16 && D.nbuf == 2) DECIM_LAUNCH(8, 16, 2);
        else if (D.T == 32) DECIM_LAUNCH(8, 32, 1);
        else DECIM_LAUNCH(8, 16, 1);
    } else {
        if (D.T == 32 && D.nbuf == 2) DECIM_LAUNCH(16, 32, 2);
        else if (D.T == 16 && D.nbuf == 2) DECIM_LAUNCH(16, 16, 2);
        else if (D.T == 32) DECIM_LAUNCH(16, 32, 1);
        else DECIM_LAUNCH(16, 16, 1);
    }
#undef DECIM_LAUNCH
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
