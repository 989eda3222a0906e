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
    int iq_correct;                             // WFM: 1 = iq_correction fused in front of the discriminator
    float scale, norm;
    const double *tabF, *AF, *AFB, *AB, *ABB, *MB, *CR, *CB, *head, *tailT, *tailM;
    double DB;
    int tab_in_smem, U_in_smem, yout_in_smem;
    long long yout_off;                         // doubles from the start of the CTA's global U slice (yout_in_smem == 0)
    int off_tile, off_misc, off_tab, off_U;     // byte offsets into dynamic shared memory
    size_t smem_bytes, U_bytes;
};

struct FrameDevFwd {
    int N = 0, n_taps = 0, n_sections = 0, C = 0, B = 0;
    int T = 0, n_tiles = 1;     // long blocks run as n_tiles tiles of T samples (T == N when the block fits one CTA)
    const float* taps = nullptr;
    const double* sos = nullptr;
    const double* AC = nullptr;
    const double* ACB = nullptr;
    size_t smem_bytes = 0;
    double coef[5][5] = {};     // b0 b1 b2 a1 a2 (a0-normalised) by value: kernel-parameter constants
};

struct pss_demod_plan {
    int kind = 0, mode = 0, N = 0, out_len = 0, channels = 1, plain = 0;
    DecimDev dec{};
    FrameDevFwd frm{};
    std::vector<void*> dev_allocs;
    void* U_scratch = nullptr;
    size_t U_scratch_bytes = 0;
    void* tile_scratch = nullptr;       // per-block max|y| of tiled FIR plans
    size_t tile_scratch_bytes = 0;
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
    float rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(mx));      // 1 MUFU; t is within 1.5 ulp
    float t = mn * rc;
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
// EDGE_CHECK = false is the interior fast path: every sample index of the tile is inside the block,
// so the loads and stores carry no range predicates.
template <int PF, bool EDGE_CHECK>
__device__ __forceinline__ void tile_load(float2 (&pf)[PF], const float2* __restrict__ x, const int g0,
                                          const int n_wi, const int N, const int warp, const int lane) {
    const float2* xp = x + g0 + 31 * warp + lane;
#pragma unroll
    for (int it = 0; it < PF; ++it) {
        const int wi = warp + it * (DEMOD_THREADS / 32);
        bool ok = wi < n_wi;
        if (EDGE_CHECK) {
            const int gi = g0 + 31 * wi + lane;
            ok = ok && gi >= 0 && gi < N;
        }
        pf[it] = make_float2(0.f, 0.f);
        if (ok) pf[it] = __ldg(xp + it * (31 * (DEMOD_THREADS / 32)));
    }
}

template <bool WFM, int PF, bool EDGE_CHECK>
__device__ __forceinline__ void tile_store(float* __restrict__ buf, const float2 (&pf)[PF],
                                           const float2* __restrict__ x, const int g0, const int E,
                                           const int n_wi, const int N, const IqCorr k, const float scale,
                                           const int warp, const int lane) {
    const int L = N - 1;
    float* bp = buf + 31 * warp + lane - 1;
#pragma unroll
    for (int it = 0; it < PF; ++it) {
        const int wi = warp + it * (DEMOD_THREADS / 32);
        float2 cur = pf[it];
        if (WFM) cur = iq_apply(cur, k);
        float2 prev;
        prev.x = __shfl_up_sync(0xffffffffu, cur.x, 1);
        prev.y = __shfl_up_sync(0xffffffffu, cur.y, 1);
        float d = disc_core<WFM>(cur, prev, scale);
        const int e = 31 * wi + lane - 1;
        if (EDGE_CHECK) {
            const int g = g0 + e;
            if (g < 0 || g >= L) d = 0.f;
        }
        if (wi < n_wi && lane > 0 && e < E) bp[it * (31 * (DEMOD_THREADS / 32))] = d;
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

// One row of y = A x with x spread over the S lanes of a group.  The state is broadcast through a
// per-group shared-memory line (one 8-byte store, S/2 16-byte broadcast loads) instead of 2*S
// shuffles; `xb` is the group's double-buffered line [2][S], `ph` flips every call.  Four independent
// partial sums keep the dependent fp64 chain short.
template <int S>
__device__ __forceinline__ double matvec_row(const double (&a)[S], const double x, double* xb, int& ph,
                                             const int r) {
    double* line = xb + ph * S;
    ph ^= 1;
    line[r] = x;
    __syncwarp();
    const double2* l2 = reinterpret_cast<const double2*>(line);
    double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
    for (int c = 0; c < S; c += 4) {
        const double2 u = l2[c / 2], v = l2[c / 2 + 1];
        p0 = fma(a[c], u.x, p0);
        p1 = fma(a[c + 1], u.y, p1);
        p2 = fma(a[c + 2], v.x, p2);
        p3 = fma(a[c + 3], v.y, p3);
    }
    return (p0 + p1) + (p2 + p3);
}

// x_{i+1} = A x_i + u_i, i = 0..n-1, over the slot field U[slot*rows + foff + r]; forward walks
// slots 1..n, backward walks slots n..1.  Blocked: every group of S lanes owns one block of B steps.
template <int S>
__device__ void blocked_scan(double* U, const int rows, const int foff, const int n, const bool fwd,
                             const double* __restrict__ A, const double* __restrict__ APow, const int B,
                             const double* x0, double* XS, double* XB, const int tid) {
    const int r = tid % S, grp = tid / S;
    const int n_units = (n + B - 1) / B;
    double* xb = XB + (size_t)grp * 2 * S;      // this group's broadcast line (double-buffered)
    int ph = 0;
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
            x = u + matvec_row<S>(a, x, xb, ph, r);
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
            X = u + matvec_row<S>(ap, X, xb, ph, r);
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
            z = matvec_row<S>(a, z, xb, ph, r);
            if (act) U[slot * rows + foff + r] += z;
        }
    }
    __syncthreads();
}

template <int SF, int T, int NBUF>
__global__ void __launch_bounds__(DEMOD_THREADS, 2)
demod_decim_kernel(const DecimDev D, const float2* __restrict__ iq, float* __restrict__ audio,
                   const long long n_frames, double* __restrict__ U_global, const double* __restrict__ moments,
                   const int mom_fpb) {
    constexpr bool WFM = SF == 16;
    constexpr int SB = 8, ROWS = SF + SB + 1, NT = (ROWS + 7) / 8;
    extern __shared__ __align__(16) unsigned char smem[];
    float* tile = reinterpret_cast<float*>(smem + D.off_tile);
    double* misc = reinterpret_cast<double*>(smem + D.off_misc);
    double* XS = misc;                 // [32][16]
    double* XB = misc + 512;           // [groups][2][S] scan broadcast lines
    double* dh = misc + 1024;          // [28] head discriminator samples
    double* red = dh + 32;             // [32] reduction scratch
    double* tres = red + 32;           // [SB + m_tail] tail result (<= 64)
    const double* tab = D.tab_in_smem ? reinterpret_cast<const double*>(smem + D.off_tab) : D.tabF;
    double* U = D.U_in_smem ? reinterpret_cast<double*>(smem + D.off_U)
                            : U_global + (size_t)blockIdx.x * (D.U_bytes / 8);
    double* yout = D.yout_in_smem ? tres + 64 : U + D.yout_off;      // [n_out]; long low-rate blocks keep it in L2
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
        if (WFM && D.iq_correct) {
            // second moments over the block (fp64 accumulation), then iq_correction's estimates
            // (per-thread float partials over N/256 samples, combined in fp64: the same order of
            // rounding error as numpy's own float32 pairwise means at :52, :60, :61)
            double a = 0, b = 0, c = 0;
            if (moments) {
                // the PSD kernel already summed I^2, Q^2, IQ per FFT frame while it read this block
                const double* m = moments + (size_t)frame * mom_fpb * 4;
                for (int k = 0; k < mom_fpb; ++k) {
                    a += m[4 * k];
                    b += m[4 * k + 1];
                    c += m[4 * k + 2];
                }
            } else {
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
                for (int w = 0; w < DEMOD_THREADS / 32; ++w) {
                    a += red[w];
                    b += red[8 + w];
                    c += red[16 + w];
                }
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
        // tile ti covers chunks 1 + ti*T ..; its first discriminator index:
        auto tile_g0 = [&](int ti) { return ti * T * q + 1 - lead; };
        auto load_t = [&](int ti) {
            const int g0 = tile_g0(ti);
            if (g0 >= 0 && g0 + 31 * n_wi + 1 < D.N) tile_load<PF, false>(pf, x, g0, n_wi, D.N, warp, lane);
            else tile_load<PF, true>(pf, x, g0, n_wi, D.N, warp, lane);
        };
        auto store_t = [&](int ti) {
            const int g0 = tile_g0(ti);
            float* dst = tile + (NBUF == 2 ? (ti & 1) * D.tile_floats : 0);
            if (g0 >= 0 && g0 + 31 * n_wi + 1 < D.N)
                tile_store<WFM, PF, false>(dst, pf, x, g0, E, n_wi, D.N, kc, D.scale, warp, lane);
            else
                tile_store<WFM, PF, true>(dst, pf, x, g0, E, n_wi, D.N, kc, D.scale, warp, lane);
        };
        // tensor half of a tile: a unit = (m-tile of 8 chunks, n-tile of 8 table rows) over the whole
        // window.  MT*NT is 8 or 16 -> whole units per warp, plain stores.  MT=4, NT=3 (NFM, 32-chunk
        // tiles): n-tiles 0/1 whole, the last n-tile (it only carries the yf-forcing row) split in two
        // K halves -> atomics on that one column.
        auto dmma_t = [&](int ti) {
            const int j0 = 1 + ti * T;
            const float* cur = tile + (NBUF == 2 ? (ti & 1) * D.tile_floats : 0);
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
        };
        // Odd warps run the producer half (discriminator, ALU/LSU) of an iteration before the tensor
        // half, even warps after it, so the two pipes overlap inside a CTA; an odd warp therefore keeps
        // its IQ loads two tiles ahead.
        const bool late = NBUF == 2 && SF == 8 && (warp & 1);      // measured: +4 % NFM, -2 % WFM
        if (n_tiles > 0) {
            load_t(0);
            store_t(0);
            if (late && n_tiles > 1) load_t(1);
        }
        __syncthreads();
        if (tid <= SF) {
            double acc = 0.0;
            for (int i = 0; i <= EDGE; ++i) acc = fma(D.head[tid * (EDGE + 1) + i], dh[i], acc);
            if (tid < SF) U[tid] = acc;          // slot 0 .F = s_1
            else red[24] = acc;                  // yf at ext index 27
        }

        // ---- body chunks: discriminator tile -> DMMA against the response tables; one barrier per tile
        for (int tl = 0; tl < n_tiles; ++tl) {
            const bool more = tl + 1 < n_tiles;
            if (!late) {
                if (more) load_t(tl + 1);
                dmma_t(tl);
                if (NBUF == 1) __syncthreads();
                if (more) store_t(tl + 1);
            } else {
                if (more) store_t(tl + 1);
                if (tl + 2 < n_tiles) load_t(tl + 2);
                dmma_t(tl);
            }
            __syncthreads();
        }

        // ---- forward state scan: slot j .F becomes s_{j+1} (state after chunk j)
        blocked_scan<SF>(U, ROWS, 0, n_body, true, D.AF, D.AFB, D.Bf, U, XS, XB, tid);

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
        blocked_scan<SB>(U, ROWS, SF, n_body, false, D.AB, D.ABB, D.Bb, tres, XS, XB, tid);

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
    D.iq_correct = d->iq_correct ? 1 : 0;
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
    // the un-normalised outputs stay in shared memory unless the block has more than 4096 of them (low
    // sample rate x long read); then they live behind the state slots in the CTA's global slice
    D.yout_in_smem = D.n_out <= 4096;
    const size_t misc_b = ((size_t)(512 + 512 + 32 + 32 + 64 + (D.yout_in_smem ? D.n_out : 0)) * 8 + 15) & ~(size_t)15;
    const size_t tab_b = frag.size() * 8;
    D.U_bytes = (((size_t)(D.n_body + 2) * D.rows * 8) + 15) & ~(size_t)15;
    if (!D.yout_in_smem) {
        D.yout_off = (long long)(D.U_bytes / 8);
        D.U_bytes += ((size_t)D.n_out * 8 + 15) & ~(size_t)15;
    }
    const int cand[5][2] = {{32, 2}, {16, 2}, {32, 1}, {16, 1}, {8, 1}};      // {8, 1}: q > ~1600 (fs > 36 MS/s)
    int pick = -1;
    for (int c = 0; c < 5 && pick < 0; ++c) {
        size_t tf = (size_t)(cand[c][0] - 1) * D.q + D.Kp;
        if (tf * cand[c][1] < (size_t)D.tail_len) tf = ((size_t)D.tail_len + cand[c][1] - 1) / cand[c][1];
        tf = (tf + 3) & ~(size_t)3;
        const size_t tot = tf * 4 * cand[c][1] + misc_b + tab_b + D.U_bytes;
        if (tot <= budget || c == 4) {
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
    D.U_in_smem = D.yout_in_smem && used + D.U_bytes <= budget;
    if (D.U_in_smem) { D.off_U = (int)used; used += D.U_bytes; }
    D.smem_bytes = used;
    pl->out_len = D.n_out;
    pl->channels = 2;
    return PSS_OK;
}

static int launch_decim(pss_ctx* ctx, pss_demod_plan* pl, const float* iq, int64_t n_frames, float* audio,
                        const double* moments = nullptr, int mom_fpb = 0) {
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
                                                                         (double*)pl->U_scratch, moments, mom_fpb); \
    } while (0)
    if (D.SF == 8) {
        if (D.T == 32 && D.nbuf == 2) DECIM_LAUNCH(8, 32, 2);
        else if (D.T == 16 && D.nbuf == 2) DECIM_LAUNCH(8, 16, 2);
        else if (D.T == 32) DECIM_LAUNCH(8, 32, 1);
        else if (D.T == 16) DECIM_LAUNCH(8, 16, 1);
        else DECIM_LAUNCH(8, 8, 1);
    } else {
        if (D.T == 32 && D.nbuf == 2) DECIM_LAUNCH(16, 32, 2);
        else if (D.T == 16 && D.nbuf == 2) DECIM_LAUNCH(16, 16, 2);
        else if (D.T == 32) DECIM_LAUNCH(16, 32, 1);
        else if (D.T == 16) DECIM_LAUNCH(16, 16, 1);
        else DECIM_LAUNCH(16, 8, 1);
    }
#undef DECIM_LAUNCH
    PSS_LAUNCH_CHECK(ctx);
    return PSS_OK;
}


// =================================================================================================
// Frame kernels: AM (signal_processing.py:179-195), USB/LSB (:198-217), RAW (:237-238 + :46-80).
// One CTA of 512 threads owns one block; the fp32 working row lives in shared memory, the block is
// read from HBM once and the peak-normalised mono result written once.
// =================================================================================================
#define FRAME_THREADS 512
#define FIR_MAX_TAPS 65

typedef FrameDevFwd FrameDev;   // C = samples per thread-chunk (AM), B = scan block, AC/ACB [16][16] padded

// ---- USB / LSB: y = lfilter(taps, 1, x).real (the hilbert() round trip is the identity on the real
// part and both side-band branches are identical), / max|y| * 0.95.
__global__ void __launch_bounds__(FRAME_THREADS, 1)
demod_fir_kernel(const FrameDev D, const float2* __restrict__ iq, float* __restrict__ audio, const long long n_frames,
                 unsigned* __restrict__ blkmax /* tiled plans: per-block max|y| (float bits), zeroed by the host */) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* row = reinterpret_cast<float*>(smem) + 64;     // [-64 .. n_rounds*ROUND): history in front
    __shared__ float taps_s[FIR_MAX_TAPS + 3];
    __shared__ float redf[FRAME_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = D.N, T = D.T, n_tiles = D.n_tiles;
    if (tid < FIR_MAX_TAPS) taps_s[tid] = tid < D.n_taps ? D.taps[tid] : 0.f;
    constexpr int OUT_PER = 8, ROUND = FRAME_THREADS * OUT_PER;
    const int n_rounds = (T + ROUND - 1) / ROUND;
    // work item = (block, tile); a block longer than one CTA's shared memory is cut into tiles of T
    // samples, each tile re-reading the 64 samples in front of it; the peak normalisation of a tiled
    // block is finished by frame_scale_kernel
    for (long long item = blockIdx.x; item < n_frames * n_tiles; item += gridDim.x) {
        const long long frame = item / n_tiles;
        const int base = (int)(item - frame * n_tiles) * T;
        const int len = min(T, N - base);
        const float2* x = iq + frame * N + base;
        __syncthreads();
        if (tid < 64) row[tid - 64] = base > 0 ? __ldg(x + tid - 64).x : 0.f;      // lfilter starts from zero state
        for (int ib = tid; ib < n_rounds * ROUND; ib += FRAME_THREADS * 8) {      // ROUND = 8 * FRAME_THREADS
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = ib + u * FRAME_THREADS;
                v[u] = i < len ? __ldcs(x + i).x : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) row[ib + u * FRAME_THREADS] = v[u];
        }
        {   // next work item of this CTA -> L2 while the taps run
            const long long nx_item = item + gridDim.x;
            if (nx_item < n_frames * n_tiles) {
                const long long nf = nx_item / n_tiles;
                const char* nx = reinterpret_cast<const char*>(iq + nf * N + (nx_item - nf * n_tiles) * T);
                const int nl = min(T, N - (int)(nx_item - nf * n_tiles) * T);
                for (int l = tid; l < nl * 8 / 128; l += FRAME_THREADS) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (size_t)l * 128));
            }
        }
        __syncthreads();
        float mx = 0.f;
        // rounds walk the tile from its end so the in-place overwrite never touches unread input
        for (int r = n_rounds - 1; r >= 0; --r) {
            const int n0 = r * ROUND + tid * OUT_PER;
            float win[FIR_MAX_TAPS - 1 + 8];              // x[n0-64 .. n0+7]
            const float4* w4 = reinterpret_cast<const float4*>(row + n0 - (FIR_MAX_TAPS - 1));
#pragma unroll
            for (int k = 0; k < (FIR_MAX_TAPS - 1 + 8) / 4; ++k) {
                const float4 v = w4[k];
                win[4 * k] = v.x; win[4 * k + 1] = v.y; win[4 * k + 2] = v.z; win[4 * k + 3] = v.w;
            }
            float acc[8];
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[o] = 0.f;
#pragma unroll
            for (int k = 0; k < FIR_MAX_TAPS; ++k) {
                const float h = taps_s[k];
#pragma unroll
                for (int o = 0; o < 8; ++o) acc[o] = fmaf(h, win[o + (FIR_MAX_TAPS - 1) - k], acc[o]);
            }
            __syncthreads();
            *reinterpret_cast<float4*>(row + n0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            *reinterpret_cast<float4*>(row + n0 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
#pragma unroll
            for (int o = 0; o < 8; ++o)
                if (n0 + o < len) mx = fmaxf(mx, fabsf(acc[o]));
            __syncthreads();
        }
        mx = warp_max(mx);
        if (lane == 0) redf[warp] = mx;
        __syncthreads();
        mx = redf[0];
        for (int w = 1; w < FRAME_THREADS / 32; ++w) mx = fmaxf(mx, redf[w]);
        float* dst = audio + frame * N + base;
        if (n_tiles == 1) {
            const double g = 0.95 / (double)mx;               // y / max|y| * 0.95 (signal_processing.py:216)
            for (int i = tid; i < len; i += FRAME_THREADS) __stcs(dst + i, (float)((double)row[i] * g));
        } else {
            for (int i = tid; i < len; i += FRAME_THREADS) dst[i] = row[i];
            if (tid == 0) atomicMax(blkmax + frame, __float_as_uint(mx));      // mx >= 0 (NaN never wins, like fmaxf)
        }
    }
}

// Second half of the peak normalisation of tiled FIR plans: y * (0.95 / max|y|) per block, in place.
__global__ void __launch_bounds__(256)
frame_scale_kernel(float* __restrict__ audio, const int N, const long long n_frames, const unsigned* __restrict__ blkmax) {
    const long long total = n_frames * N;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const double g = 0.95 / (double)__uint_as_float(blkmax[i / N]);
        audio[i] = (float)((double)audio[i] * g);
    }
}

// ---- AM: |x| - mean -> cascaded biquads (fp64 state, scipy sosfilt DF2T order, zero initial state)
// -> / max|y| * 0.95.  The recurrence over the block is a chunked scan: every thread owns C
// consecutive samples; pass A gives each chunk's zero-state response end state, a blocked scan of
// the 2*n_sections-dimensional state over the chunks gives every chunk's true initial state, pass B
// re-runs the chunk from it and emits the outputs.
template <int NS>
__device__ __forceinline__ double sos_step(const double (&c)[5][5], double (&z)[NS][2], double v) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const double y = fma(c[s][0], v, z[s][0]);                    // b0*x + z0
        z[s][0] = fma(-c[s][3], y, fma(c[s][1], v, z[s][1]));         // b1*x - a1*y + z1
        z[s][1] = fma(-c[s][4], y, c[s][2] * v);                      // b2*x - a2*y
        v = y;
    }
    return v;
}

template <int NS>
__global__ void __launch_bounds__(FRAME_THREADS, 1)
demod_sos_kernel(const FrameDev D, const float2* __restrict__ iq, float* __restrict__ audio, const long long n_frames,
                 const int plain /* 1: input is a real float32 row, no envelope / mean / normalisation */) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int N = D.N, C = D.C, LC = 31 - __clz(C);       // C is a power of two
    const int T = D.T, n_tiles = D.n_tiles;               // blocks longer than one CTA's shared memory: tiles of T = C * FRAME_THREADS
    double* US = reinterpret_cast<double*>(smem);         // [(n_chunks + 2)][16] scan slots
    double* XS = US + (size_t)(FRAME_THREADS + 2) * 16;   // [32][16]
    double* XB = XS + 512;                                // [32 groups][2][16] scan broadcast lines
    double* redd = XB + 1024;                             // [32]
    float* row = reinterpret_cast<float*>(redd + 32);     // [T + T/C] skewed: idx + idx / C
    __shared__ double carry[16];                          // filter state at the start of the current tile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // the coefficients are read straight from the kernel-parameter constant bank (no registers)
    const double (&c)[5][5] = D.coef;
    auto envelope = [](const float2 v) {
        const float q = fmaf(v.x, v.x, v.y * v.y);
        // sqrt of the float sum is within 1 ulp of hypotf in the normal range; the scaled library
        // routine only where the squares over/underflow
        return (q > 1e-30f && q < 1e30f) ? __fsqrt_rn(q) : hypotf(v.x, v.y);
    };
    for (long long frame = blockIdx.x; frame < n_frames; frame += gridDim.x) {
        const float2* xb = iq + frame * N;
        const float* xrb = reinterpret_cast<const float*>(iq) + frame * N;
        float* dst = audio + frame * N;
        float mean = 0.f;
        __syncthreads();
        if (n_tiles > 1 && !plain) {
            // tiled block: the mean of the whole envelope is needed before the first tile is filtered
            double sum = 0.0;
            for (int ib = tid; ib < N; ib += FRAME_THREADS * 8) {
                float2 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = ib + u * FRAME_THREADS;
                    v[u] = i < N ? __ldg(xb + i) : make_float2(0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (ib + u * FRAME_THREADS < N) sum += (double)envelope(v[u]);
            }
            sum = warp_sum(sum);
            if (lane == 0) redd[warp] = sum;
            __syncthreads();
            double tot = 0.0;
            for (int w = 0; w < FRAME_THREADS / 32; ++w) tot += redd[w];
            mean = (float)(tot / (double)N);
        }
        if (tid < 16) carry[tid] = 0.0;                   // sosfilt starts from zero state
        float mx = 0.f;
        for (int tile = 0; tile < n_tiles; ++tile) {
            const int base = tile * T, len = min(T, N - base);
            const int n_chunks = (len + C - 1) / C;       // <= FRAME_THREADS
            const float2* x = xb + base;
            const float* xr = xrb + base;
            __syncthreads();
            // envelope (float32 hypot like np.abs on complex64) and its mean; 8 loads in flight per thread
            double sum = 0.0;
            for (int ib = tid; ib < len; ib += FRAME_THREADS * 8) {
                float e[8];
                if (plain) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = ib + u * FRAME_THREADS;
                        e[u] = i < len ? __ldg(xr + i) : 0.f;
                    }
                } else {
                    float2 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = ib + u * FRAME_THREADS;
                        v[u] = i < len ? __ldcs(x + i) : make_float2(0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) e[u] = envelope(v[u]);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = ib + u * FRAME_THREADS;
                    if (i < len) {
                        row[i + (i >> LC)] = e[u];
                        sum += (double)e[u];
                    }
                }
            }
            if (n_tiles == 1) {   // next block of this CTA -> L2 while the recurrences run
                const long long nf = frame + gridDim.x;
                if (nf < n_frames && !plain) {
                    const char* nx = reinterpret_cast<const char*>(iq + nf * N);
                    for (int l = tid; l < N * 8 / 128; l += FRAME_THREADS) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (size_t)l * 128));
                }
            }
            for (int i = tid; i < (FRAME_THREADS + 2) * 16; i += FRAME_THREADS) US[i] = i < 16 ? carry[i] : 0.0;   // slot 0 = start state
            if (n_tiles == 1) {
                sum = warp_sum(sum);
                if (lane == 0) redd[warp] = sum;
            }
            __syncthreads();
            if (n_tiles == 1 && !plain) {
                double tot = 0.0;
                for (int w = 0; w < FRAME_THREADS / 32; ++w) tot += redd[w];
                mean = (float)(tot / (double)N);                   // np.mean(envelope), float32
            }
            // pass A: zero-state response end state of every chunk
            const int i0 = tid * C, i1 = min(len, i0 + C);
            const float* rp = row + i0 + tid;                              // skew: i0 / C == tid
            if (tid < n_chunks) {
                double z[NS][2];
#pragma unroll
                for (int s = 0; s < NS; ++s) z[s][0] = z[s][1] = 0.0;
                for (int i = 0; i < i1 - i0; ++i) sos_step<NS>(c, z, (double)__fsub_rn(rp[i], mean));
                double* u = US + (size_t)(tid + 1) * 16;
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    u[2 * s] = z[s][0];
                    u[2 * s + 1] = z[s][1];
                }
            }
            __syncthreads();
            // true state after every chunk: x_{c+1} = AC x_c + u_c, x_0 = the tile's start state (slot 0)
            blocked_scan<16>(US, 16, 0, n_chunks, true, D.AC, D.ACB, D.B, US, XS, XB, tid);
            // pass B: re-run from the true initial state (slot tid = state after chunk tid-1), in place
            if (tid < n_chunks) {
                double z[NS][2];
                const double* u = US + (size_t)tid * 16;
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    z[s][0] = u[2 * s];
                    z[s][1] = u[2 * s + 1];
                }
                float* wp = row + i0 + tid;
                for (int i = 0; i < i1 - i0; ++i) {
                    const float y = (float)sos_step<NS>(c, z, (double)__fsub_rn(wp[i], mean));
                    wp[i] = y;
                    mx = fmaxf(mx, fabsf(y));
                }
            }
            if (n_tiles > 1) {
                __syncthreads();
                if (tid < 16) carry[tid] = US[(size_t)n_chunks * 16 + tid];      // state after the tile's last chunk
                for (int i = tid; i < len; i += FRAME_THREADS) dst[base + i] = row[i + (i >> LC)];
            }
        }
        mx = warp_max(mx);
        __syncthreads();
        if (lane == 0) redd[warp] = (double)mx;
        __syncthreads();
        double m = redd[0];
        for (int w = 1; w < FRAME_THREADS / 32; ++w) m = fmax(m, redd[w]);
        if (n_tiles == 1) {
            if (plain) {
                for (int i = tid; i < N; i += FRAME_THREADS) dst[i] = row[i + (i >> LC)];
            } else {
                const double g = 0.95 / m;                    // y / max|y| * 0.95 (signal_processing.py:194)
                for (int i = tid; i < N; i += FRAME_THREADS) __stcs(dst + i, (float)((double)row[i + (i >> LC)] * g));
            }
        } else if (!plain) {
            const double g = 0.95 / m;                        // the unscaled tiles were written by this CTA
            for (int i = tid; i < N; i += FRAME_THREADS) dst[i] = (float)((double)dst[i] * g);
        }
    }
}

// ---- RAW: real(iq_correction(x)) (signal_processing.py:46-80, :237-238), float32 like the reference
__global__ void __launch_bounds__(256)
demod_raw_kernel(const int N, const float2* __restrict__ iq, float* __restrict__ audio, const long long n_frames,
                 const int complex_out /* 1: write the whole corrected complex64 block (iq_correction) */) {
    __shared__ double red[5][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (long long frame = blockIdx.x; frame < n_frames; frame += gridDim.x) {
        const float2* x = iq + frame * N;
        double m[5] = {0, 0, 0, 0, 0};     // sum I, Q, I^2, Q^2, IQ
        for (int i = tid; i < N; i += 256) {
            const float2 v = __ldg(x + i);
            const double a = v.x, b = v.y;
            m[0] += a; m[1] += b; m[2] = fma(a, a, m[2]); m[3] = fma(b, b, m[3]); m[4] = fma(a, b, m[4]);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            m[k] = warp_sum(m[k]);
            if (lane == 0) red[k][warp] = m[k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += red[k][w];
            m[k] = t / (double)N;          // means
        }
        const double p_in = (m[2] + m[3]) - (m[0] * m[0] + m[1] * m[1]);          // var(x - mean)  :48-49
        const float q_amp = (float)sqrt(2.0 * m[3]);                              // :52
        const double qa = q_amp;
        const float alpha = (float)sqrt(2.0 * m[2] / (qa * qa));                  // :60
        const float sin_phi = (float)((2.0 / (double)alpha) * (m[4] / (qa * qa)));   // :61
        const float cos_phi = sqrtf(1.f - sin_phi * sin_phi);                     // :64
        const float inv_q = 1.f / q_amp, inv_a = 1.f / alpha, g = -sin_phi / alpha, inv_c = 1.f / cos_phi;
        // corrected = (a1*I, a2*I + a3*Q); its variance from the input moments
        const double a1 = (double)inv_q * inv_a * inv_c, a2 = (double)g * inv_q * inv_c, a3 = (double)inv_q * inv_c;
        const double e2 = a1 * a1 * m[2] + a2 * a2 * m[2] + 2.0 * a2 * a3 * m[4] + a3 * a3 * m[3];
        const double mr = a1 * m[0], mi = a2 * m[0] + a3 * m[1];
        const float scale = (float)sqrt(p_in / (e2 - (mr * mr + mi * mi)));       // :80
        if (complex_out) {
            float2* dst = reinterpret_cast<float2*>(audio) + frame * N;
            for (int i = tid; i < N; i += 256) {
                const float2 v = __ldg(x + i);
                const float zr = __fmul_rn(v.x, inv_q), zi = __fmul_rn(v.y, inv_q);
                const float i2 = __fmul_rn(inv_a, zr);
                const float q2 = __fadd_rn(__fmul_rn(g, zr), zi);
                dst[i] = make_float2(__fmul_rn(__fmul_rn(i2, inv_c), scale), __fmul_rn(__fmul_rn(q2, inv_c), scale));
            }
        } else {
            float* dst = audio + frame * N;
            for (int i = tid; i < N; i += 256) {
                const float I = __ldg(x + i).x;
                const float zr = __fmul_rn(I, inv_q);
                dst[i] = __fmul_rn(__fmul_rn(__fmul_rn(inv_a, zr), inv_c), scale);
            }
        }
    }
}

static int create_frame(pss_ctx* ctx, const pss_demod_desc* d, pss_demod_plan* pl) {
    FrameDev& F = pl->frm;
    F.N = d->N;
    pl->out_len = d->N;
    pl->channels = 1;
    int rc;
    const void* p;
    if (d->kind == PSS_PLAN_RAW) return PSS_OK;
    if (d->kind == PSS_PLAN_FIR) {
        if (!d->taps || d->n_taps < 1 || d->n_taps > FIR_MAX_TAPS) return PSS_ERR_UNSUPPORTED;
        std::vector<float> t(d->n_taps);
        for (int i = 0; i < d->n_taps; ++i) t[i] = (float)d->taps[i];
        if ((rc = upload(ctx, pl, t.data(), t.size() * 4, &p))) return rc;
        F.taps = (const float*)p;
        F.n_taps = d->n_taps;
        const size_t round = (size_t)FRAME_THREADS * 8;
        F.T = d->N;
        F.n_tiles = 1;
        F.smem_bytes = (64 + ((size_t)F.T + round - 1) / round * round) * 4;
        if (F.smem_bytes > 220 * 1024) {      // block too long for one CTA: tiles of 32768 samples
            F.T = 32768;
            F.n_tiles = (d->N + F.T - 1) / F.T;
            F.smem_bytes = (64 + ((size_t)F.T + round - 1) / round * round) * 4;
        }
        return PSS_OK;
    }
    // SOS
    if (!d->sos || d->n_sections < 1 || d->n_sections > 5) return PSS_ERR_UNSUPPORTED;
    const int ns = d->n_sections;
    F.n_sections = ns;
    F.T = d->N;
    F.n_tiles = 1;
    if (d->N > 32768) {                                    // tiles of 64 samples x FRAME_THREADS chunks
        F.T = 32768;
        F.n_tiles = (d->N + F.T - 1) / F.T;
    }
    F.C = 1;                                               // samples per thread-chunk: power of two
    while ((long long)F.C * FRAME_THREADS < F.T) F.C *= 2;
    const int n_chunks = (F.T + F.C - 1) / F.C;
    F.B = (n_chunks + 31) / 32;
    if (F.B < 1) F.B = 1;
    // zero-input transition of one chunk, by stepping the cascade on unit states (padded to 16x16)
    std::vector<double> AC(256, 0.0), ACB(256, 0.0), tmp(256, 0.0);
    for (int col = 0; col < 2 * ns; ++col) {
        std::vector<double> z(2 * ns, 0.0);
        z[col] = 1.0;
        for (int step = 0; step < F.C; ++step) {
            double v = 0.0;
            for (int s = 0; s < ns; ++s) {
                const double* c = d->sos + s * 6;
                const double y = (c[0] * v + z[2 * s] * c[3]) / c[3];
                const double z0 = (c[1] * v - c[4] * y) / c[3] + z[2 * s + 1];
                const double z1 = (c[2] * v - c[5] * y) / c[3];
                z[2 * s] = z0;
                z[2 * s + 1] = z1;
                v = y;
            }
        }
        for (int r = 0; r < 2 * ns; ++r) AC[r * 16 + col] = z[r];
    }
    ACB = AC;
    for (int k = 1; k < F.B; ++k) {
        for (int r = 0; r < 16; ++r)
            for (int c2 = 0; c2 < 16; ++c2) {
                double acc = 0.0;
                for (int m = 0; m < 16; ++m) acc += AC[r * 16 + m] * ACB[m * 16 + c2];
                tmp[r * 16 + c2] = acc;
            }
        ACB = tmp;
    }
    if ((rc = upload(ctx, pl, d->sos, (size_t)ns * 6 * 8, &p))) return rc; F.sos = (const double*)p;
    for (int s2 = 0; s2 < ns; ++s2) {
        const double* cc = d->sos + s2 * 6;
        F.coef[s2][0] = cc[0] / cc[3];
        F.coef[s2][1] = cc[1] / cc[3];
        F.coef[s2][2] = cc[2] / cc[3];
        F.coef[s2][3] = cc[4] / cc[3];
        F.coef[s2][4] = cc[5] / cc[3];
    }
    if ((rc = upload(ctx, pl, AC.data(), 256 * 8, &p))) return rc; F.AC = (const double*)p;
    if ((rc = upload(ctx, pl, ACB.data(), 256 * 8, &p))) return rc; F.ACB = (const double*)p;
    F.smem_bytes = ((size_t)(FRAME_THREADS + 2) * 16 * 8 + 512 * 8 + 1024 * 8 + 32 * 8 + ((size_t)F.T + F.T / F.C + 8) * 4 + 15) & ~(size_t)15;
    if (F.smem_bytes > 220 * 1024) return PSS_ERR_UNSUPPORTED;
    return PSS_OK;
}

static int launch_frame(pss_ctx* ctx, pss_demod_plan* pl, const float* iq, int64_t n_frames, float* audio) {
    FrameDev& F = pl->frm;
    long long grid = ctx->sm_count;
    if (pl->kind == PSS_PLAN_RAW) grid *= 4;
    if (grid > n_frames) grid = n_frames;
    if (pl->kind == PSS_PLAN_RAW) {
        demod_raw_kernel<<<(unsigned)grid, 256, 0, ctx->stream>>>(F.N, (const float2*)iq, audio, n_frames,
                                                                  pl->channels == 2 ? 1 : 0);
    } else if (pl->kind == PSS_PLAN_FIR) {
        PSS_CUDA(ctx, cudaFuncSetAttribute(demod_fir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F.smem_bytes));
        unsigned* blkmax = nullptr;
        if (F.n_tiles > 1) {
            int rc = pss_reserve(ctx, &pl->tile_scratch, &pl->tile_scratch_bytes, (size_t)n_frames * 4);
            if (rc) return rc;
            blkmax = (unsigned*)pl->tile_scratch;
            PSS_CUDA(ctx, cudaMemsetAsync(blkmax, 0, (size_t)n_frames * 4, ctx->stream));
            grid = ctx->sm_count;
            if (grid > n_frames * F.n_tiles) grid = n_frames * F.n_tiles;
        }
        demod_fir_kernel<<<(unsigned)grid, FRAME_THREADS, F.smem_bytes, ctx->stream>>>(F, (const float2*)iq, audio, n_frames, blkmax);
        if (F.n_tiles > 1) {
            PSS_LAUNCH_CHECK(ctx);
            frame_scale_kernel<<<(unsigned)(4 * ctx->sm_count), 256, 0, ctx->stream>>>(audio, F.N, n_frames, blkmax);
        }
    } else {
#define SOS_LAUNCH(NSv)                                                                                         \
    do {                                                                                                        \
        auto k = demod_sos_kernel<NSv>;                                                                         \
        PSS_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F.smem_bytes)); \
        k<<<(unsigned)grid, FRAME_THREADS, F.smem_bytes, ctx->stream>>>(F, (const float2*)iq, audio, n_frames, \
                                                                        pl->plain);                             \
    } while (0)
        switch (F.n_sections) {
            case 1: SOS_LAUNCH(1); break;
            case 2: SOS_LAUNCH(2); break;
            case 3: SOS_LAUNCH(3); break;
            case 4: SOS_LAUNCH(4); break;
            default: SOS_LAUNCH(5); break;
        }
#undef SOS_LAUNCH
    }
    PSS_LAUNCH_CHECK(ctx);
    return PSS_OK;
}

void pss_demod_release(pss_ctx*) {}

extern "C" {

int pss_demod_plan_create(pss_ctx* ctx, const pss_demod_desc* desc, pss_demod_plan** out) {
    if (!ctx || !desc || !out || desc->N < 2) return PSS_ERR_ARG;
    if (desc->struct_size != sizeof(pss_demod_desc)) return PSS_ERR_ARG;
    *out = nullptr;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    pss_demod_plan* pl = new (std::nothrow) pss_demod_plan();
    if (!pl) return PSS_ERR_NOMEM;
    pl->kind = desc->kind;
    pl->mode = desc->mode;
    pl->N = desc->N;
    int rc = PSS_ERR_UNSUPPORTED;
    if (desc->kind == PSS_PLAN_DECIM) rc = create_decim(ctx, desc, pl);
    else if (desc->kind == PSS_PLAN_FIR || desc->kind == PSS_PLAN_SOS || desc->kind == PSS_PLAN_RAW)
        rc = create_frame(ctx, desc, pl);
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
    cudaFree(pl->tile_scratch);
    cudaFree(pl->d_taps_f32);
    cudaFree(pl->d_sos);
    delete pl;
}

int pss_demod_plan_out_len(const pss_demod_plan* pl) { return pl ? pl->out_len : 0; }
int pss_demod_plan_block_len(const pss_demod_plan* pl) { return pl ? pl->N : 0; }
int pss_demod_plan_channels(const pss_demod_plan* pl) { return pl ? pl->channels : 0; }

int pss_demod_c64_dev(pss_ctx* ctx, pss_demod_plan* pl, const float* iq, int64_t n_frames, float* audio) {
    if (!ctx || !pl || !iq || !audio || n_frames < 0) return PSS_ERR_ARG;
    if (n_frames == 0) return PSS_OK;
    if (pl->kind == PSS_PLAN_DECIM) return launch_decim(ctx, pl, iq, n_frames, audio);
    return launch_frame(ctx, pl, iq, n_frames, audio);
}

int pss_demod_c64_dev_moments(pss_ctx* ctx, pss_demod_plan* pl, const float* iq, int64_t n_frames, float* audio,
                              const double* moments, int frames_per_block, int frame_len) {
    if (!ctx || !pl || !iq || !audio || n_frames < 0) return PSS_ERR_ARG;
    // the moment rows must tile the plan's block exactly, or the kernel would sum rows of another block
    if (moments && (frames_per_block < 1 || (long long)frames_per_block * frame_len != pl->N)) return PSS_ERR_ARG;
    if (n_frames == 0) return PSS_OK;
    if (pl->kind == PSS_PLAN_DECIM && pl->dec.SF == 16) return launch_decim(ctx, pl, iq, n_frames, audio, moments, frames_per_block);
    return pss_demod_c64_dev(ctx, pl, iq, n_frames, audio);
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


// ---- helpers outside the main loop's hot path, same kernels -------------------------------------
int pss_iq_correct_c64(pss_ctx* ctx, const float* iq, int N, int64_t n_frames, float* out) {
    if (!ctx || !iq || !out || N < 2 || n_frames < 0) return PSS_ERR_ARG;
    pss_demod_desc d{};
    d.struct_size = sizeof d;
    d.kind = PSS_PLAN_RAW;
    d.mode = PSS_MODE_RAW;
    d.N = N;
    pss_demod_plan* pl = nullptr;
    int rc = pss_demod_plan_create(ctx, &d, &pl);
    if (rc) return rc;
    pl->channels = 2;
    rc = pss_demod_c64(ctx, pl, iq, n_frames, out);
    pss_demod_plan_destroy(ctx, pl);
    return rc;
}

int pss_sosfilt_f32(pss_ctx* ctx, const float* x, int N, int64_t n_frames, const double* sos, int n_sections,
                    float* y) {
    if (!ctx || !x || !y || !sos || N < 1 || n_frames < 0) return PSS_ERR_ARG;
    if (n_frames == 0) return PSS_OK;
    pss_demod_desc d{};
    d.struct_size = sizeof d;
    d.kind = PSS_PLAN_SOS;
    d.mode = PSS_MODE_AM;
    d.N = N;
    d.sos = sos;
    d.n_sections = n_sections;
    pss_demod_plan* pl = nullptr;
    int rc = pss_demod_plan_create(ctx, &d, &pl);
    if (rc) return rc;
    pl->plain = 1;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t b = (size_t)n_frames * N * 4;
    if (!(rc = pss_reserve(ctx, &ctx->d_in, &ctx->d_in_bytes, b)) && !(rc = pss_reserve(ctx, &ctx->d_out, &ctx->d_out_bytes, b))) {
        cudaMemcpyAsync(ctx->d_in, x, b, cudaMemcpyHostToDevice, ctx->stream);
        rc = pss_demod_c64_dev(ctx, pl, (const float*)ctx->d_in, n_frames, (float*)ctx->d_out);
        if (!rc) {
            cudaMemcpyAsync(y, ctx->d_out, b, cudaMemcpyDeviceToHost, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = PSS_ERR_CUDA;
        }
    }
    pss_demod_plan_destroy(ctx, pl);
    return rc;
}

}  // extern "C"
