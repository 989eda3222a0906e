// Decimating demodulators: demodulate_nfm (signal_processing.py:91-116) and demodulate_wfm (:119-176,
// with iq_correction :46-80), i.e. float32 discriminator -> [65-tap FIR | Butterworth low-pass + /2 +
// de-emphasis] -> scipy.signal.decimate(q) (8th-order Chebyshev sosfiltfilt, every q-th sample) -> peak
// normalisation.
//
// The chain is linear after the discriminator, and only every q-th output exists.  Cut the block into
// chunks of q samples: the filter states after a chunk are (state before) x (chunk transition) + (the
// chunk's samples) x (response table).  pyspecsdr_b200/filters.py probes scipy's own procedure for those
// tables and block-diagonalises the transitions into independent 2x2 real blocks (modal coordinates).
// Two kernels, nothing sequential in either:
//
//   demod_force_kernel  streaming, one warp per group of 8 chunks, warps never synchronise with each
//       other: load IQ (coalesced 8-byte loads), iq-correct (WFM), float32 discriminator exactly the way
//       numpy evaluates angle(s[1:] * conj(s[:-1])), window into the warp's shared-memory slice, fp64
//       tensor-core products (mma.sync m8n8k4.f64 = DMMA) against the fragment-ordered table, one plain
//       fp64 dot product for the single forward-output row; the 8 x (SF+SB+1) forcing values go to a
//       scratch slot per chunk.  The scratch of one sub-batch of blocks is sized to stay in L2.
//   demod_scan_kernel   one small CTA per block of audio: head, forward block scan (states before every
//       chunk), coupling into the backward forcing, tail, backward block scan, outputs, peak, store.
//       A block scan = every thread folds SCAN_CPL chunks, a Kogge-Stone scan over the 32 lanes with the
//       precomputed powers M^(2^s), warp carries with M^32, per-lane M^l for the carry-in.
//   demod_corr_kernel   per-block iq_correction coefficients from the I/Q second moments (which the PSD
//       kernel emits as a by-product of its own pass, or demod_moments_kernel computes).
#include "pss_demod.cuh"

// ------------------------------------------------------------------------------------------------
// iq_correction's estimates (signal_processing.py:52-64) from the block's second moments.
__global__ void demod_corr_kernel(const double* __restrict__ moments, const int mom_fpb, const int N,
                                  const long long n_frames, float4* __restrict__ corr) {
    pss_grid_dependency_sync();       // PSS_PDL: the moments come from the previous kernel of the stream
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    double a = 0, b = 0, c = 0;
    const double* m = moments + (size_t)f * mom_fpb * 4;
    for (int k = 0; k < mom_fpb; ++k) {
        a += m[4 * k];
        b += m[4 * k + 1];
        c += m[4 * k + 2];
    }
    const double n = (double)N;
    const float q_amp = (float)sqrt(2.0 * b / n);                        // :52
    const double qa = (double)q_amp;
    const float alpha = (float)sqrt(2.0 * a / n / (qa * qa));           // :60
    const float sin_phi = (float)((2.0 / (double)alpha) * (c / n / (qa * qa)));   // :61
    const float cos_phi = sqrtf(1.f - sin_phi * sin_phi);               // :64
    corr[f] = make_float4(1.f / q_amp, 1.f / alpha, -sin_phi / alpha, 1.f / cos_phi);
}

// Second moments of a block when no PSD pass supplied them: sum I^2, Q^2, IQ (float partials per thread,
// fp64 across the block: the same order of rounding error as numpy's float32 pairwise means).
__global__ void __launch_bounds__(256)
demod_moments_kernel(const float2* __restrict__ iq, const int N, const long long n_frames, double* __restrict__ mom) {
    __shared__ double red[3][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (long long f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const float2* x = iq + f * N;
        float fii[4] = {0.f, 0.f, 0.f, 0.f}, fqq[4] = {0.f, 0.f, 0.f, 0.f}, fiq[4] = {0.f, 0.f, 0.f, 0.f};
        int i = tid;
        for (; i + 7 * 256 < N; i += 8 * 256) {
            float2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(x + i + u * 256);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                fii[u & 3] = fmaf(v[u].x, v[u].x, fii[u & 3]);
                fqq[u & 3] = fmaf(v[u].y, v[u].y, fqq[u & 3]);
                fiq[u & 3] = fmaf(v[u].x, v[u].y, fiq[u & 3]);
            }
        }
        for (; i < N; i += 256) {
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
        __syncthreads();
        if (lane == 0) {
            red[0][warp] = sii;
            red[1][warp] = sqq;
            red[2][warp] = siq;
        }
        __syncthreads();
        if (tid == 0) {
            double a = 0, b = 0, c = 0;
            for (int w = 0; w < 8; ++w) {
                a += red[0][w];
                b += red[1][w];
                c += red[2][w];
            }
            double* m = mom + f * 4;
            m[0] = a; m[1] = b; m[2] = c; m[3] = 0.0;
        }
    }
}

// ------------------------------------------------------------------------------------------------ kernel 1
// Discriminator samples d[g0 .. g0 + len) of block x into dst[0 .. len) (zero outside [0, L)); one warp.
// Every iteration loads 32 consecutive IQ samples (one 8-byte load per lane), corrects each once, takes the
// neighbour from the previous lane and produces 31 outputs.  Batches of UNR iterations, the next batch's loads
// issued before the current one is consumed.  Nothing but the final store is predicated, so the shuffles sit
// in straight-line code.  EDGE_CHECK = false: every sample index is inside the block.
template <bool WFM, bool EDGE_CHECK>
__device__ __forceinline__ void warp_discriminate(float* __restrict__ dst, const float2* __restrict__ x, const int g0,
                                                  const int len, const int N, const IqCorr kc, const float scale,
                                                  const int lane) {
    constexpr int UNR = 4;
    const int L = N - 1;
    const int n_it = (len + 30) / 31;
    const float2* xp = x + g0 + lane;
    float2 nx[UNR];
    auto fetch = [&](int it0) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int it = it0 + u;
            bool ok = it < n_it;
            if (EDGE_CHECK) {
                const int gi = g0 + 31 * it + lane;
                ok = ok && gi >= 0 && gi < N;
            }
            nx[u] = make_float2(0.f, 0.f);
            if (ok) nx[u] = __ldg(xp + 31 * it);
        }
    };
    fetch(0);
    float* dp = dst + lane - 1;
    for (int it0 = 0; it0 < n_it; it0 += UNR) {
        float2 pf[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) pf[u] = nx[u];
        fetch(it0 + UNR);
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            float2 cur = pf[u];
            if (WFM) cur = iq_apply(cur, kc);
            float2 prev;
            prev.x = __shfl_up_sync(0xffffffffu, cur.x, 1);
            prev.y = __shfl_up_sync(0xffffffffu, cur.y, 1);
            float d = disc_core<WFM>(cur, prev, scale);
            const int e = 31 * (it0 + u) + lane - 1;
            if (EDGE_CHECK) {
                const int g = g0 + e;
                if (g < 0 || g >= L) d = 0.f;
            }
            if (lane > 0 && e < len) dp[31 * (it0 + u)] = d;
        }
    }
}

template <bool WFM>
__device__ __forceinline__ void warp_disc(float* __restrict__ dst, const float2* __restrict__ x, const int g0, const int len,
                                          const int N, const IqCorr kc, const float scale, const int lane) {
    const int n_it = (len + 30) / 31;
    if (g0 >= 0 && g0 + 31 * n_it + 1 < N) warp_discriminate<WFM, false>(dst, x, g0, len, N, kc, scale, lane);
    else warp_discriminate<WFM, true>(dst, x, g0, len, N, kc, scale, lane);
}

// Scratch of one block of audio: [rows][CS] doubles, row-major with the chunk index contiguous (chunk
// c = j - 1 of body chunk j); rows 0..SF-1 forward forcing -> forward state, SF..SF+7 backward forcing,
// row SF+8 the forward output at the kept sample -> the un-normalised audio sample.
template <int SF>
__global__ void __launch_bounds__(FORCE_THREADS, 3)
demod_force_kernel(const DecimDev D, const float2* __restrict__ iq, const int n_frames, double* __restrict__ F,
                   const float4* __restrict__ corr) {
    constexpr bool WFM = SF == 16;
    constexpr int SB = 8, NTD = (SF + SB) / 8, RROW = SF + SB;       // r row = the forward output at the kept sample
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* wbuf = reinterpret_cast<float*>(smem) + warp * WBUF_FLOATS;
    const double* tab = D.tabF;
    if (D.tab_in_smem) {
        double* ts = reinterpret_cast<double*>(smem + (size_t)(FORCE_THREADS / 32) * WBUF_FLOATS * 4);
        const int n = D.KS * (NTD * 32 + 4);
        for (int i = tid; i < n; i += FORCE_THREADS) ts[i] = D.tabF[i];
        tab = ts;
        __syncthreads();
    }
    const double* trow = tab + (size_t)D.KS * NTD * 32;               // [KS][4]: r-row taps
    const int q = D.q, lead = D.lead, Kp = D.Kp, KS = D.KS, groups = D.groups, CS = D.CS;
    const int n_units = n_frames * groups;
    const int wstride = gridDim.x * (FORCE_THREADS / 32);
    for (int unit = blockIdx.x * (FORCE_THREADS / 32) + warp; unit < n_units; unit += wstride) {
        const int frame = unit / groups;
        const int g = unit - frame * groups;
        const float2* x = iq + (long long)frame * D.N;
        IqCorr kc = {1.f, 1.f, 0.f, 1.f};
        if (WFM && D.iq_correct) {
            const float4 c = __ldg(corr + frame);
            kc = {c.x, c.y, c.z, c.w};
        }
        const int g0 = 8 * g * q + 1 - lead;                          // first discriminator index of the group
        double c0[NTD], c1[NTD], racc = 0.0;
#pragma unroll
        for (int nt = 0; nt < NTD; ++nt) c0[nt] = c1[nt] = 0.0;
        if (D.contiguous) {
            // one contiguous window: chunk m's samples start at m*q
            warp_disc<WFM>(wbuf, x, g0, 7 * q + Kp, D.N, kc, D.scale, lane);
            __syncwarp();
            const float* arow = wbuf + (lane >> 2) * q + (lane & 3);
            const double* bp = tab + lane;
            const double* tp = trow + (lane & 3);
#pragma unroll 3
            for (int ks = 0; ks < KS; ++ks) {
                const double a = (double)arow[4 * ks];
#pragma unroll
                for (int nt = 0; nt < NTD; ++nt) dmma_m8n8k4(c0[nt], c1[nt], a, bp[(ks * NTD + nt) * 32]);
                racc = fma(a, tp[4 * ks], racc);
            }
        } else {
            // long windows: k-slabs of SLAB_K samples, the 8 chunk rows side by side (row stride SLAB_ROW)
            for (int k0 = 0; k0 < Kp; k0 += SLAB_K) {
                const int len = min(SLAB_K, Kp - k0);
                __syncwarp();
                for (int r = 0; r < 8; ++r)
                    warp_disc<WFM>(wbuf + r * SLAB_ROW, x, g0 + r * q + k0, len, D.N, kc, D.scale, lane);
                __syncwarp();
                const float* arow = wbuf + (lane >> 2) * SLAB_ROW + (lane & 3);
                const double* bp = tab + lane + (size_t)(k0 / 4) * NTD * 32;
                const double* tp = trow + (k0 / 4) * 4 + (lane & 3);
                for (int ks = 0; ks < len / 4; ++ks) {
                    const double a = (double)arow[4 * ks];
#pragma unroll
                    for (int nt = 0; nt < NTD; ++nt) dmma_m8n8k4(c0[nt], c1[nt], a, bp[(ks * NTD + nt) * 32]);
                    racc = fma(a, tp[4 * ks], racc);
                }
            }
        }
        // the r row: the four lanes of a quad hold the partial sums of one chunk
        racc += __shfl_xor_sync(0xffffffffu, racc, 1);
        racc += __shfl_xor_sync(0xffffffffu, racc, 2);
        const int c = 8 * g + (lane >> 2);                            // chunk index (body chunk j = c + 1)
        if (c < D.n_body) {
            double* col = F + (long long)frame * D.slot_doubles + c;
#pragma unroll
            for (int nt = 0; nt < NTD; ++nt) {
                const int row = nt * 8 + 2 * (lane & 3);
                col[(long long)row * CS] = c0[nt];
                col[(long long)(row + 1) * CS] = c1[nt];
            }
            if ((lane & 3) == 0) col[(long long)RROW * CS] = racc;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ kernel 1, fused
// The same products with the discriminator computed where the tensor-core fragment needs it: lane
// (r = lane / 4, c = lane % 4) of k-step ks owns window sample r*q + 4*ks + c of chunk r, which is exactly
// the A-fragment element of mma.m8n8k4.  Each lane loads ONE IQ sample per k-step and chunk group (a depth-PD
// register pipeline keeps the loads in flight), corrects it once, and takes the later neighbour from lane c+1
// or, for c == 3, from the next k-step's sample of lane (r, 0).  No shared-memory window, no warp barrier,
// any q; shared memory holds the response table only.  A warp works on U consecutive groups of 8 chunks at a
// time and feeds every table fragment it loads to U tensor-core products: the kernel is bound by the
// shared-memory / shuffle data path (ncu: 69 % of the LSU wavefront peak at U = 1), not by issue or the DMMA pipe.
template <int SF, bool EDGE_CHECK, bool TAB_SMEM, int DIAG, int UNR, int U, int PDX, bool PACK>
__device__ __forceinline__ void force_unit_fused(const DecimDev& D, const float2* __restrict__ x, const int g0, const IqCorr kc,
                                                 const double* __restrict__ tab, const double* __restrict__ trow,
                                                 double (&c0)[U][(SF + 8) / 8], double (&c1)[U][(SF + 8) / 8],
                                                 double (&racc)[U], const int lane) {
    constexpr bool WFM = SF == 16;
    constexpr int NTD = (SF + 8) / 8, PD = PDX ? PDX : (U == 1 ? 6 : 4);
    const int r = lane >> 2, c = lane & 3, KS = D.KS, N = D.N, L = D.N - 1;
    const int e0 = r * D.q + c, ustep = 8 * D.q;                      // group u starts 8 chunks further
    const float2* xp = x + g0 + e0;
    auto load = [&](int ks, int u) {
        if (DIAG == 3) return make_float2(1.f + ks, 0.5f * lane);                   // DIAG 3: timing without the IQ loads
        // (measured: clamping the index of the never-consumed k-steps beyond KS instead of predicating the load is 4 %
        // slower, 0.633 vs 0.609 ms per GiB; widening the unchecked path's condition to the pipeline's whole reach
        // and dropping the predicate 1.5 % slower -- more units take the checked path)
        bool ok = ks <= KS;
        if (EDGE_CHECK) {
            const int gi = g0 + u * ustep + e0 + 4 * ks;
            ok = ok && gi >= 0 && gi < N;
        }
        float2 v = make_float2(0.f, 0.f);
        if (ok) v = __ldg(xp + u * ustep + 4 * ks);
        return v;
    };
    // iq_correction (signal_processing.py:55-71) up to its positive common factor 1 / (q_amp * cos_phi), which a
    // phase difference does not see: (I', Q') = (I / alpha, Q - sin_phi / alpha * I) -- two operations per sample
    auto correct = [&](const float2 v) { return make_float2(__fmul_rn(kc.inv_a, v.x), __fmaf_rn(kc.g, v.x, v.y)); };
    float2 ring[U][PD], S0[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int i = 0; i < PD; ++i) ring[u][i] = load(i + 1, u);
        S0[u] = load(0, u);
        if (WFM) S0[u] = correct(S0[u]);
    }
    const double* bp = tab + lane;
    const double* tp = trow + c;
    const int src0 = lane & ~3;
    double racc1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) racc1[u] = 0.0;
#pragma unroll UNR
    for (int ks = 0; ks < KS; ++ks) {
        double av[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float2 S1 = ring[u][0];
#pragma unroll
            for (int i = 0; i + 1 < PD; ++i) ring[u][i] = ring[u][i + 1];
            ring[u][PD - 1] = load(ks + PD + 1, u);
            if (WFM) S1 = correct(S1);
            float2 a, n;
            a.x = __shfl_down_sync(0xffffffffu, S0[u].x, 1);
            a.y = __shfl_down_sync(0xffffffffu, S0[u].y, 1);
            n.x = __shfl_sync(0xffffffffu, S1.x, src0);
            n.y = __shfl_sync(0xffffffffu, S1.y, src0);
            if (c == 3) a = n;
            float d = DIAG == 2 ? a.x + S0[u].y : disc_core<WFM>(a, S0[u], D.scale);   // DIAG 2: timing without the discriminator math
            if (EDGE_CHECK) {
                const int g = g0 + u * ustep + e0 + 4 * ks;
                if (g < 0 || g >= L) d = 0.f;
            }
            av[u] = (double)d;
            S0[u] = S1;
        }
        double bv[4];
        if (PACK) {                  // two 16-byte loads per k-step: (b0, b1), (b2 | r tap, r tap | -)
            const double2* pp = reinterpret_cast<const double2*>(tab) + (size_t)ks * 64 + lane;
            const double2 p0 = TAB_SMEM ? pp[0] : __ldg(pp), p1 = TAB_SMEM ? pp[32] : __ldg(pp + 32);
            bv[0] = p0.x, bv[1] = p0.y, bv[2] = p1.x, bv[3] = p1.y;
        }
#pragma unroll
        for (int nt = 0; nt < NTD; ++nt) {
            double b;
            if (PACK) b = bv[nt];
            else if (TAB_SMEM) b = bp[(ks * NTD + nt) * 32];
            else b = __ldg(bp + (ks * NTD + nt) * 32);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (DIAG == 1) c0[u][nt] = fma(av[u], b, c0[u][nt]);                 // DIAG 1: timing without the tensor pipe
                else dmma_m8n8k4(c0[u][nt], c1[u][nt], av[u], b);
            }
        }
        const double tr = PACK ? bv[NTD] : (TAB_SMEM ? tp[4 * ks] : __ldg(tp + 4 * ks));
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (ks & 1) racc1[u] = fma(av[u], tr, racc1[u]);
            else racc[u] = fma(av[u], tr, racc[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) racc[u] += racc1[u];
}

template <int SF, bool TAB_SMEM, int DIAG = 0, int UNR = 6, int MINB = 3, int U = 1, int PDX = 0, bool PACK = false>
__global__ void __launch_bounds__(FORCE_THREADS, MINB)
demod_force_fused_kernel(const DecimDev D, const float2* __restrict__ iq, const int n_frames, double* __restrict__ F,
                         const float4* __restrict__ corr) {
    constexpr bool WFM = SF == 16;
    constexpr int SB = 8, NTD = (SF + SB) / 8, RROW = SF + SB;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* tab = PACK ? D.tabP : D.tabF;
    if (TAB_SMEM) {
        double* ts = reinterpret_cast<double*>(smem);
        const int n = PACK ? D.KS * 128 : D.KS * (NTD * 32 + 4);
        const double* src = tab;
        for (int i = tid; i < n; i += FORCE_THREADS) ts[i] = src[i];
        tab = ts;
        __syncthreads();
    }
    pss_grid_dependency_sync();       // PSS_PDL: the table is in shared memory; corr / the scratch belong to earlier kernels
    const double* trow = tab + (size_t)D.KS * NTD * 32;               // [KS][4]: r-row taps
    const int q = D.q, CS = D.CS;
    const int sgroups = (D.groups + U - 1) / U;                       // super-groups of U * 8 chunks per block
    const int n_units = n_frames * sgroups;
    const int wstride = gridDim.x * (FORCE_THREADS / 32);
    for (int unit = blockIdx.x * (FORCE_THREADS / 32) + warp; unit < n_units; unit += wstride) {
        const int frame = unit / sgroups;
        const int g = (unit - frame * sgroups) * U;
        const float2* x = iq + (long long)frame * D.N;
        IqCorr kc = {1.f, 1.f, 0.f, 1.f};
        if (WFM && D.iq_correct) {
            const float4 cf = __ldg(corr + frame);
            kc = {cf.x, cf.y, cf.z, cf.w};
        }
        const int g0 = 8 * g * q + 1 - D.lead;                        // first discriminator index of the first group
        double c0[U][NTD], c1[U][NTD], racc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            racc[u] = 0.0;
#pragma unroll
            for (int nt = 0; nt < NTD; ++nt) c0[u][nt] = c1[u][nt] = 0.0;
        }
        if (g0 >= 0 && g0 + (8 * U - 1) * q + 4 * D.KS + 4 < D.N)
            force_unit_fused<SF, false, TAB_SMEM, DIAG, UNR, U, PDX, PACK>(D, x, g0, kc, tab, trow, c0, c1, racc, lane);
        else
            force_unit_fused<SF, true, TAB_SMEM, DIAG, UNR, U, PDX, PACK>(D, x, g0, kc, tab, trow, c0, c1, racc, lane);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            double ra = racc[u];
            ra += __shfl_xor_sync(0xffffffffu, ra, 1);
            ra += __shfl_xor_sync(0xffffffffu, ra, 2);
            const int c = 8 * (g + u) + (lane >> 2);                  // chunk index (body chunk j = c + 1)
            if (c < D.n_body) {
                double* col = F + (long long)frame * D.slot_doubles + c;
#pragma unroll
                for (int nt = 0; nt < NTD; ++nt) {
                    const int row = nt * 8 + 2 * (lane & 3);
                    col[(long long)row * CS] = c0[u][nt];
                    col[(long long)(row + 1) * CS] = c1[u][nt];
                }
                if ((lane & 3) == 0) col[(long long)RROW * CS] = ra;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ kernel 2
__device__ __forceinline__ void mv2(const double* __restrict__ B, const double x0, const double x1, double& y0,
                                    double& y1) {
    y0 = fma(B[0], x0, B[1] * x1);
    y1 = fma(B[2], x0, B[3] * x1);
}

// One 2x2 recurrence x_{i+1} = B x_i + f_i over the warp's 32 * SCAN_CPL steps: every lane folds its SCAN_CPL
// consecutive steps, a Kogge-Stone scan over the lanes combines the lane aggregates with the powers
// PW[s] = (B^SCAN_CPL)^(2^s).  FWD: lane 0 holds the earliest steps, otherwise lane 31 does.
// In: (f0, f1)[i] the lane's forcing, (s0, s1) the state before the segment's first step (warp-uniform).
// Out: (X0, X1) = state before the lane's first step; (s0, s1) = state after the segment's last step.
template <bool FWD>
__device__ __forceinline__ void lane_scan(const double (&f0)[SCAN_CPL], const double (&f1)[SCAN_CPL], double& s0, double& s1,
                                          double& X0, double& X1, const double* __restrict__ B,
                                          const double* __restrict__ PW, const int pw_stride, const int lane) {
    const int first = FWD ? 0 : 31;
    double e0 = lane == first ? s0 : 0.0, e1 = lane == first ? s1 : 0.0;
#pragma unroll
    for (int i = 0; i < SCAN_CPL; ++i) {
        const int k = FWD ? i : SCAN_CPL - 1 - i;
        double y0, y1;
        mv2(B, e0, e1, y0, y1);
        e0 = y0 + f0[k];
        e1 = y1 + f1[k];
    }
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const double o0 = FWD ? __shfl_up_sync(0xffffffffu, e0, 1 << s) : __shfl_down_sync(0xffffffffu, e0, 1 << s);
        const double o1 = FWD ? __shfl_up_sync(0xffffffffu, e1, 1 << s) : __shfl_down_sync(0xffffffffu, e1, 1 << s);
        const bool has = FWD ? lane >= (1 << s) : lane + (1 << s) < 32;
        double y0, y1;
        mv2(PW + s * pw_stride, o0, o1, y0, y1);
        if (has) {
            e0 += y0;
            e1 += y1;
        }
    }
    const double p0 = FWD ? __shfl_up_sync(0xffffffffu, e0, 1) : __shfl_down_sync(0xffffffffu, e0, 1);
    const double p1 = FWD ? __shfl_up_sync(0xffffffffu, e1, 1) : __shfl_down_sync(0xffffffffu, e1, 1);
    X0 = lane == first ? s0 : p0;
    X1 = lane == first ? s1 : p1;
    s0 = __shfl_sync(0xffffffffu, e0, FWD ? 31 : 0);
    s1 = __shfl_sync(0xffffffffu, e1, FWD ? 31 : 0);
}

// One CTA of 4 warps per block of audio.  The 2x2 recurrences are independent, so every warp scans its own
// ones without talking to the others; CTA barriers only separate the phases.
// TMA bulk copy (cp.async.bulk, SASS UBLKCP) of one block's scratch into shared memory, completion on an mbarrier.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, const unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, const unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, const unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, const unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// SLOT_SMEM: the block's whole scratch (rows x CS doubles, ~60 KB at the bench geometry) is pulled into shared
// memory by one TMA bulk copy while the head / tail discriminator samples are computed, every phase then
// runs at shared-memory latency and nothing is written back (only the audio leaves).  Blocks whose scratch
// does not fit keep it in global memory (L2).
template <int SF, bool SLOT_SMEM>
__global__ void __launch_bounds__(SCAN_THREADS, 3)
demod_scan_kernel(const DecimDev D, const float2* __restrict__ iq, float* __restrict__ audio, const int n_frames,
                  double* __restrict__ F, const float4* __restrict__ corr) {
    constexpr bool WFM = SF == 16;
    constexpr int SB = 8, NF = SF / 2, NBK = SB / 2, RROW = SF + SB, CPL = SCAN_CPL, SEG = 32 * CPL, NW = SCAN_THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem[];
    double* tabs = reinterpret_cast<double*>(smem);                  // packed scan tables
    double* zs = tabs + D.scan_tab_doubles;                           // [SF]  modal state entering chunk 1
    double* zend = zs + SF;                                           // [SF]  forward state after the last body chunk
    double* tres = zend + SF;                                         // [64]  tail result: backward start state, last outputs
    double* dh = tres + 64;                                           // [28]  head discriminator samples
    double* t1 = dh + 28;                                             // [SB]  backward state after chunk 0
    double* red = t1 + SB;                                            // [16]  [0..7] warp maxima, [8] yf at ext index 27
    float* tile = reinterpret_cast<float*>(red + 16);                 // [tail_pad] tail discriminator samples
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(tile + D.tail_pad);
    double* Fs = reinterpret_cast<double*>(bar + 2);                  // [rows][CS] (SLOT_SMEM)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const ScanTabLayout T{NF, NBK};
    auto issue_slot_copy = [&](const int frame) {       // one thread: the block's forcing rows -> shared memory (TMA bulk copies)
        const unsigned bytes = (unsigned)(D.slot_doubles * 8);
        mbar_expect_tx(bar, bytes);
        const char* src = reinterpret_cast<const char*>(F + (long long)frame * D.slot_doubles);
        for (unsigned o = 0; o < bytes; o += 32768u)              // bulk copies of <= 32 KB
            bulk_g2s(reinterpret_cast<char*>(Fs) + o, src + o, min(32768u, bytes - o), bar);
    };
    pss_grid_dependency_sync();       // PSS_PDL: the forcing rows belong to the previous kernel
    if (SLOT_SMEM && tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if ((int)blockIdx.x < n_frames) issue_slot_copy(blockIdx.x);   // the first block's copy runs beside the table copy
    }
    unsigned phase = 0;
    const double *Gm = tabs + T.G(), *CR = tabs + T.CR(), *CB = tabs + T.CB();
    const double *BF = tabs + T.BF(), *PWF = tabs + T.PWF(), *BBk = tabs + T.BB(), *PWB = tabs + T.PWB();
    const int nbd = D.n_body, L = D.L, CS = D.CS;

    for (int frame = blockIdx.x; frame < n_frames; frame += gridDim.x) {
        const float2* x = iq + (long long)frame * D.N;
        double* Fb = SLOT_SMEM ? Fs : F + (long long)frame * D.slot_doubles;
        __syncthreads();                                              // the previous block is done with Fs / bar
        if (SLOT_SMEM && tid == 0 && frame != (int)blockIdx.x) issue_slot_copy(frame);
        IqCorr kc = {1.f, 1.f, 0.f, 1.f};
        if (WFM && D.iq_correct) {
            const float4 c = __ldg(corr + frame);
            kc = {c.x, c.y, c.z, c.w};
        }
        // ---- head: ext[0..27] depends on d[0..27] only; tail window
        if (tid <= EDGE) dh[tid] = (double)discriminator<WFM>(x, tid, L, kc, D.scale);
        for (int i = tid; i < D.tail_len; i += SCAN_THREADS) tile[i] = discriminator<WFM>(x, D.tail_start + i, L, kc, D.scale);
        if (frame == (int)blockIdx.x)             // the scan tables: first needed after this barrier
            for (int i = tid; i < D.scan_tab_doubles; i += SCAN_THREADS) tabs[i] = D.scanTab[i];
        __syncthreads();
        if (tid <= SF) {
            double acc = 0.0;
            for (int i = 0; i <= EDGE; ++i) acc = fma(D.head[tid * (EDGE + 1) + i], dh[i], acc);
            if (tid < SF) zs[tid] = acc;             // modal state entering chunk 1
            else red[8] = acc;                       // yf at ext index 27
        }
        if (SLOT_SMEM) {
            mbar_wait(bar, phase);                   // the forcing rows have landed
            phase ^= 1;
        }
        __syncthreads();

        // ---- forward scans: warp w owns blocks w, w + 4, ...; forcing rows -> states BEFORE every chunk, in place
        for (int b = warp; b < NF; b += NW) {
            double s0 = zs[2 * b], s1 = zs[2 * b + 1];
            for (int base = 0; base < nbd; base += SEG) {
                const int c0 = base + lane * CPL;
                double* r0 = Fb + (long long)(2 * b) * CS + c0;
                double* r1 = r0 + CS;
                double f0[CPL], f1[CPL];
#pragma unroll
                for (int i = 0; i < CPL; i += 2) {                    // c0 and CS are even: 16-byte loads
                    const double2 u = c0 + i < nbd ? *reinterpret_cast<const double2*>(r0 + i) : make_double2(0.0, 0.0);
                    const double2 v = c0 + i < nbd ? *reinterpret_cast<const double2*>(r1 + i) : make_double2(0.0, 0.0);
                    f0[i] = u.x;
                    f0[i + 1] = c0 + i + 1 < nbd ? u.y : 0.0;
                    f1[i] = v.x;
                    f1[i + 1] = c0 + i + 1 < nbd ? v.y : 0.0;
                }
                double X0, X1;
                lane_scan<true>(f0, f1, s0, s1, X0, X1, BF + 4 * b, PWF + 4 * b, NF * 4, lane);
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                    if (c0 + i < nbd) {
                        r0[i] = X0;
                        r1[i] = X1;
                    }
                    double y0, y1;
                    mv2(BF + 4 * b, X0, X1, y0, y1);
                    X0 = y0 + f0[i];
                    X1 = y1 + f1[i];
                    if (c0 + i == nbd - 1) {
                        zend[2 * b] = X0;
                        zend[2 * b + 1] = X1;
                    }
                }
            }
            if (nbd == 0 && lane == 0) {
                zend[2 * b] = s0;
                zend[2 * b + 1] = s1;
            }
        }
        __syncthreads();

        // ---- tail block: backward state entering chunk n_body, and the last m_tail outputs
        for (int rr = warp; rr < SB + D.m_tail; rr += NW) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            const double* tr = D.tailT + (size_t)rr * D.tail_len;
            int i = lane;
            for (; i + 96 < D.tail_len; i += 128) {                  // four loads in flight
                const double t0 = tr[i], t1v = tr[i + 32], t2 = tr[i + 64], t3 = tr[i + 96];
                a0 = fma(t0, (double)tile[i], a0);
                a1 = fma(t1v, (double)tile[i + 32], a1);
                a2 = fma(t2, (double)tile[i + 64], a2);
                a3 = fma(t3, (double)tile[i + 96], a3);
            }
            for (; i < D.tail_len; i += 32) a0 = fma(tr[i], (double)tile[i], a0);
            double acc = (a0 + a1) + (a2 + a3);
            if (lane < SF) acc = fma(D.tailM[rr * SF + lane], zend[lane], acc);
            acc = warp_sum(acc);
            if (lane == 0) tres[rr] = acc;
        }
        // ---- coupling: backward forcing fb = wB + G z and forward output yl = r + CR z of every chunk
        for (int c = tid; c < nbd; c += SCAN_THREADS) {
            double* col = Fb + c;
            double z[SF];
#pragma unroll
            for (int k = 0; k < SF; ++k) z[k] = col[(long long)k * CS];
            double w[SB + 1];
#pragma unroll
            for (int r = 0; r <= SB; ++r) w[r] = col[(long long)(SF + r) * CS];
#pragma unroll
            for (int k = 0; k < SF; ++k) w[SB] = fma(CR[k], z[k], w[SB]);
#pragma unroll 2
            for (int r = 0; r < SB; ++r) {
                double acc = w[r];
#pragma unroll
                for (int k = 0; k < SF; ++k) acc = fma(Gm[r * SF + k], z[k], acc);
                col[(long long)(SF + r) * CS] = acc;
            }
            col[(long long)RROW * CS] = w[SB];
        }
        __syncthreads();

        // ---- backward scans (top-aligned segments, lane 31 holds the latest chunks): warp w owns block w;
        // backward forcing rows -> backward states BEFORE every chunk is processed, in place
        for (int b = warp; b < NBK; b += NW) {
            double s0 = tres[2 * b], s1 = tres[2 * b + 1];
            for (int top = nbd; top > 0; top -= SEG) {
                const int c0 = top - SEG + lane * CPL;                // may be negative at the bottom
                double* r0 = Fb + (long long)(SF + 2 * b) * CS + c0;
                double* r1 = r0 + CS;
                double f0[CPL], f1[CPL];
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                    const bool ok = c0 + i >= 0;
                    f0[i] = ok ? r0[i] : 0.0;
                    f1[i] = ok ? r1[i] : 0.0;
                }
                double X0, X1;
                lane_scan<false>(f0, f1, s0, s1, X0, X1, BBk + 4 * b, PWB + 4 * b, NBK * 4, lane);
#pragma unroll
                for (int i = CPL - 1; i >= 0; --i) {
                    if (c0 + i >= 0) {
                        r0[i] = X0;
                        r1[i] = X1;
                    }
                    double y0, y1;
                    mv2(BBk + 4 * b, X0, X1, y0, y1);
                    X0 = y0 + f0[i];
                    X1 = y1 + f1[i];
                    if (c0 + i == 0) {                                // state after chunk 0 gives output 0
                        t1[2 * b] = X0;
                        t1[2 * b + 1] = X1;
                    }
                }
            }
            if (nbd == 0 && lane == 0) {
                t1[2 * b] = s0;
                t1[2 * b + 1] = s1;
            }
        }
        __syncthreads();

        // ---- outputs: y_j = CB . t_{j+1} + DB * yl_j (chunk c = j - 1), y_0 from t_1 and the head
        double mx = 0.0;
        for (int c = tid; c < nbd; c += SCAN_THREADS) {
            double* col = Fb + c;
            double y = D.DB * col[(long long)RROW * CS];
#pragma unroll
            for (int k = 0; k < SB; ++k) y = fma(CB[k], col[(long long)(SF + k) * CS], y);
            col[(long long)RROW * CS] = y;
            mx = fmax(mx, fabs(y));
        }
        double y_first = D.DB * red[8];
#pragma unroll
        for (int k = 0; k < SB; ++k) y_first = fma(CB[k], t1[k], y_first);
        mx = fmax(mx, fabs(y_first));
        for (int i = tid; i < D.m_tail; i += SCAN_THREADS) mx = fmax(mx, fabs(tres[SB + i]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) red[warp] = mx;
        __syncthreads();                                              // also orders the y stores above
        mx = red[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) mx = fmax(mx, red[w]);
        const double gain = (double)D.norm / mx;                     // audio / max|audio| * 0.95 (:115)
        float2* dst = reinterpret_cast<float2*>(audio) + (long long)frame * D.n_out;
        for (int k = tid; k < D.n_out; k += SCAN_THREADS) {
            const double y = k == 0 ? y_first : (k <= nbd ? Fb[(long long)RROW * CS + k - 1] : tres[SB + (k - nbd - 1)]);
            const float v = (float)(y * gain);
            dst[k] = make_float2(v, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------ host side
static void mat2_mul(const double* a, const double* b, double* c) {
    const double r[4] = {a[0] * b[0] + a[1] * b[2], a[0] * b[1] + a[1] * b[3], a[2] * b[0] + a[3] * b[2],
                         a[2] * b[1] + a[3] * b[3]};
    memcpy(c, r, sizeof r);
}

static void fill_scan_tables(const double* blocks, int nb, double* B, double* PW) {
    for (int b = 0; b < nb; ++b) {
        const double* blk = blocks + 4 * b;
        memcpy(B + 4 * b, blk, 32);
        double M[4] = {1, 0, 0, 1};
        for (int c = 0; c < SCAN_CPL; ++c) mat2_mul(blk, M, M);       // M = B^CPL
        double P[4];
        memcpy(P, M, 32);
        for (int s = 0; s < 6; ++s) {                                 // M^(2^s)
            memcpy(PW + (s * nb + b) * 4, P, 32);
            mat2_mul(P, P, P);
        }
    }
}

int pss_decim_create(pss_ctx* ctx, const pss_demod_desc* d, pss_demod_plan* pl) {
    if (d->SB != 8 || (d->SF != 8 && d->SF != 16)) return PSS_ERR_UNSUPPORTED;
    if (!d->body || !d->BF || !d->BB || !d->G || !d->CR || !d->CB || !d->head || !d->tail_T || !d->tail_M)
        return PSS_ERR_ARG;
    if (d->q < 2 || d->n_body < 0 || d->m_tail < 1 || d->m_tail > 48 || d->tail_len < EDGE + 1) return PSS_ERR_ARG;
    if (d->n_out != d->n_body + 1 + d->m_tail) return PSS_ERR_ARG;
    DecimDev& D = pl->dec;
    D.mode = d->mode; D.N = d->N; D.L = d->N - 1; D.q = d->q; D.n_out = d->n_out; D.lead = d->lead;
    D.SF = d->SF; D.SB = d->SB; D.n_body = d->n_body; D.m_tail = d->m_tail;
    D.tail_start = d->tail_start; D.tail_len = d->tail_len;
    D.scale = d->scale; D.norm = d->norm; D.DB = d->DB;
    D.iq_correct = d->iq_correct ? 1 : 0;
    D.rows = D.SF + D.SB + 1;
    D.NTD = (D.SF + D.SB) / 8;
    D.CS = (D.n_body + 15) & ~15;
    if (D.CS == 0) D.CS = 16;
    D.groups = (D.n_body + 7) / 8;
    D.slot_doubles = (long long)D.rows * D.CS;
    D.tail_pad = (D.tail_len + 3) & ~3;
    const int win = D.q + D.lead;
    D.Kp = (win + 3) & ~3;
    D.KS = D.Kp / 4;
    D.contiguous = 7 * D.q + D.Kp <= WBUF_FLOATS;
    // fragment-ordered body table: [(ks*NTD + nt)*32 + lane] = T[row nt*8 + lane/4][i 4ks + lane%4], then the r row
    std::vector<double> frag((size_t)D.KS * (D.NTD * 32 + 4), 0.0);
    for (int ks = 0; ks < D.KS; ++ks) {
        for (int nt = 0; nt < D.NTD; ++nt)
            for (int l = 0; l < 32; ++l) {
                const int row = nt * 8 + l / 4, i = 4 * ks + l % 4;
                if (i < win) frag[((size_t)ks * D.NTD + nt) * 32 + l] = d->body[(size_t)row * win + i];
            }
        for (int c = 0; c < 4; ++c)
            if (4 * ks + c < win) frag[(size_t)D.KS * D.NTD * 32 + ks * 4 + c] = d->body[(size_t)(D.SF + D.SB) * win + 4 * ks + c];
    }
    const ScanTabLayout T{D.SF / 2, D.SB / 2};
    std::vector<double> st(T.total(), 0.0);
    fill_scan_tables(d->BF, T.nf, &st[T.BF()], &st[T.PWF()]);
    fill_scan_tables(d->BB, T.nb, &st[T.BB()], &st[T.PWB()]);
    memcpy(&st[T.G()], d->G, (size_t)D.SB * D.SF * 8);
    memcpy(&st[T.CR()], d->CR, (size_t)D.SF * 8);
    memcpy(&st[T.CB()], d->CB, (size_t)D.SB * 8);
    D.scan_tab_doubles = T.total();
    int rc;
    const void* p;
    if ((rc = pss_demod_upload(ctx, pl, frag.data(), frag.size() * 8, &p))) return rc; D.tabF = (const double*)p;
    if ((rc = pss_demod_upload(ctx, pl, st.data(), st.size() * 8, &p))) return rc; D.scanTab = (const double*)p;
    {   // the same fragments packed for 16-byte loads: per k-step two double2 per lane, (b0, b1) and (b2, r tap) [NTD = 3]
        // or (r tap, 0) [NTD = 2]
        std::vector<double> pk((size_t)D.KS * 2 * 32 * 2, 0.0);
        for (int ks = 0; ks < D.KS; ++ks)
            for (int l = 0; l < 32; ++l) {
                double v[4] = {0.0, 0.0, 0.0, 0.0};
                for (int nt = 0; nt < D.NTD; ++nt) v[nt] = frag[((size_t)ks * D.NTD + nt) * 32 + l];
                v[D.NTD] = frag[(size_t)D.KS * D.NTD * 32 + ks * 4 + (l & 3)];
                for (int e = 0; e < 4; ++e) pk[(((size_t)ks * 2 + e / 2) * 32 + l) * 2 + (e & 1)] = v[e];
            }
        if ((rc = pss_demod_upload(ctx, pl, pk.data(), pk.size() * 8, &p))) return rc;
        D.tabP = (const double*)p;
        D.packed_tab_smem = pk.size() * 8 <= 72 * 1024 ? (int)(pk.size() * 8) : 0;
    }
    if ((rc = pss_demod_upload(ctx, pl, d->head, (size_t)(D.SF + 1) * (EDGE + 1) * 8, &p))) return rc; D.head = (const double*)p;
    if ((rc = pss_demod_upload(ctx, pl, d->tail_T, (size_t)(D.SB + D.m_tail) * D.tail_len * 8, &p))) return rc; D.tailT = (const double*)p;
    if ((rc = pss_demod_upload(ctx, pl, d->tail_M, (size_t)(D.SB + D.m_tail) * D.SF * 8, &p))) return rc; D.tailM = (const double*)p;
    // forcing kernel: 8 warp windows + the table when it fits beside them in 72 KB (3 CTAs / SM)
    const size_t wb = (size_t)(FORCE_THREADS / 32) * WBUF_FLOATS * 4, tb = frag.size() * 8;
    D.tab_in_smem = wb + tb <= 72 * 1024;
    D.force_smem = (int)(wb + (D.tab_in_smem ? tb : 0));
    D.fused_tab_smem = tb <= 72 * 1024 ? (int)tb : 0;              // fused kernel: the table alone
    // scan kernel: one CTA per block of audio; the block's scratch in shared memory when 3 CTAs / SM still fit
    const size_t ss = (size_t)(D.scan_tab_doubles + 2 * D.SF + 64 + 28 + 8 + 16) * 8 + (size_t)D.tail_pad * 4 + 16 + 16;
    if (ss > 200 * 1024) return PSS_ERR_UNSUPPORTED;
    const size_t slot_b = (size_t)D.slot_doubles * 8;
    D.scan_slot_smem = ss + slot_b <= 74 * 1024;
    D.scan_smem = (int)(ss + (D.scan_slot_smem ? slot_b : 0));
    D.scan_warps = SCAN_THREADS / 32;
    pl->out_len = D.n_out;
    pl->channels = 2;
    return PSS_OK;
}

// PSS_DEMOD_FORCE=window selects the shared-memory-window variant of the forcing kernel (kept for comparison,
// DESIGN.md 4); the default is the fused register-pipelined one.
static bool use_window_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PSS_DEMOD_FORCE");
        v = (e && !strcmp(e, "window")) ? 1 : 0;
    }
    return v == 1;
}

template <int SF>
static int launch_force(pss_ctx* ctx, pss_demod_plan* pl, const float2* iq, long long nf, double* F, const float4* corr,
                        bool beside_scan) {
    DecimDev& D = pl->dec;
    if (nf * D.groups > 0x7fffffffLL) return PSS_ERR_UNSUPPORTED;
    const long long units = nf * D.groups;
    long long g1 = (units + FORCE_THREADS / 32 - 1) / (FORCE_THREADS / 32);
    if (g1 < 1) return PSS_OK;
    if (use_window_variant()) {
        auto kf = demod_force_kernel<SF>;
        PSS_CUDA(ctx, cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, D.force_smem));
        if (g1 > 3LL * ctx->sm_count) g1 = 3LL * ctx->sm_count;
        kf<<<(unsigned)g1, FORCE_THREADS, D.force_smem, ctx->stream>>>(D, iq, (int)nf, F, corr);
        PSS_LAUNCH_CHECK(ctx);
        return PSS_OK;
    }
    // 2 CTAs / SM when a scan kernel shares the SMs (its CTA needs the third slot's registers), else 3
    const long long per_sm = beside_scan ? 2 : 3;
    if (g1 > per_sm * ctx->sm_count) g1 = per_sm * ctx->sm_count;
    const int tb = D.fused_tab_smem;
    static const int diag = getenv("PSS_DIAG") ? atoi(getenv("PSS_DIAG")) : 0;   // timing-only builds, wrong results
    static const int var = getenv("PSS_FORCE_VARIANT") ? atoi(getenv("PSS_FORCE_VARIANT")) : 0;   // tuning experiments
    static const int pack = getenv("PSS_FORCE_PACK") ? atoi(getenv("PSS_FORCE_PACK")) : 0;       // packed table, 16-byte loads
    static const int unr = getenv("PSS_FORCE_UNR") ? atoi(getenv("PSS_FORCE_UNR")) : 0;           // 3, 4, 5, 7, 8, 9, 12
    if (tb && diag) {
        auto k = diag == 1 ? demod_force_fused_kernel<SF, true, 1> : diag == 2 ? demod_force_fused_kernel<SF, true, 2>
                 : diag == 3 ? demod_force_fused_kernel<SF, true, 3> : demod_force_fused_kernel<SF, true, 4>;
        PSS_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, tb));
        k<<<(unsigned)g1, FORCE_THREADS, tb, ctx->stream>>>(D, iq, (int)nf, F, corr);
    } else if (tb && var) {
        // 1: two groups per warp, 2 CTAs / SM;  2: two groups per warp, 3 CTAs / SM (80 registers)
        // 3 / 4: the same with the loop unrolled by their pipeline depth (4)
        auto k = var == 1 ? demod_force_fused_kernel<SF, true, 0, 3, 2, 2>
                 : var == 3 ? demod_force_fused_kernel<SF, true, 0, 4, 2, 2>
                 : var == 4 ? demod_force_fused_kernel<SF, true, 0, 4, 3, 2>
                            : demod_force_fused_kernel<SF, true, 0, 3, 3, 2>;
        PSS_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, tb));
        long long gv = ((var == 1 || var == 3) ? 2LL : 3LL) * ctx->sm_count;
        k<<<(unsigned)gv, FORCE_THREADS, tb, ctx->stream>>>(D, iq, (int)nf, F, corr);
    } else if (tb && unr) {
        // Measured (same box, WFM demodulator per GiB): the k-step loop unrolled by the depth of the load pipeline
        // (PD = 6: the register ring rotates by renaming, no moves) 0.608 ms -- the default; unrolled by 4 (the
        // previous default, 13 register moves per 4 steps) 0.654; by 3: 0.661; by 12: 0.631; pipeline depth =
        // unroll = 5 / 7 / 8 / 9: 0.651 / 0.643 / 0.623 / 0.639 (spills beyond 6 at 80 registers).
        auto k = unr == 62 ? demod_force_fused_kernel<SF, true, 0, 6, 2, 1>          // 2 CTAs / SM at up to 128 registers
                 : unr == 82 ? demod_force_fused_kernel<SF, true, 0, 8, 2, 1, 8>
                 : unr == 102 ? demod_force_fused_kernel<SF, true, 0, 10, 2, 1, 10>
                 : unr == 12 ? demod_force_fused_kernel<SF, true, 0, 12, 3, 1>
                 : unr == 5 ? demod_force_fused_kernel<SF, true, 0, 5, 3, 1, 5>
                 : unr == 7 ? demod_force_fused_kernel<SF, true, 0, 7, 3, 1, 7>
                 : unr == 8 ? demod_force_fused_kernel<SF, true, 0, 8, 3, 1, 8>
                 : unr == 9 ? demod_force_fused_kernel<SF, true, 0, 9, 3, 1, 9>
                 : unr == 3 ? demod_force_fused_kernel<SF, true, 0, 3, 3, 1>
                            : demod_force_fused_kernel<SF, true, 0, 4, 3, 1>;
        PSS_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, tb));
        if (unr > 50 && g1 > 2LL * ctx->sm_count) g1 = 2LL * ctx->sm_count;
        PSS_CUDA(ctx, pss_launch(k, (unsigned)g1, (unsigned)FORCE_THREADS, (size_t)tb, ctx->stream, D, iq, (int)nf, F, corr));
    } else if (tb && pack && D.packed_tab_smem) {
        auto k = demod_force_fused_kernel<SF, true, 0, 6, 3, 1, 0, true>;
        PSS_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, D.packed_tab_smem));
        PSS_CUDA(ctx, pss_launch(k, (unsigned)g1, (unsigned)FORCE_THREADS, (size_t)D.packed_tab_smem, ctx->stream, D, iq, (int)nf, F, corr));
    } else if (tb) {
        auto k = demod_force_fused_kernel<SF, true>;
        PSS_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, tb));
        PSS_CUDA(ctx, pss_launch(k, (unsigned)g1, (unsigned)FORCE_THREADS, (size_t)tb, ctx->stream, D, iq, (int)nf, F, corr));
    } else {
        demod_force_fused_kernel<SF, false><<<(unsigned)g1, FORCE_THREADS, 0, ctx->stream>>>(D, iq, (int)nf, F, corr);
    }
    PSS_LAUNCH_CHECK(ctx);
    return PSS_OK;
}

template <int SF>
static int launch_scan(pss_ctx* ctx, pss_demod_plan* pl, const float2* iq, long long nf, float* audio, double* F,
                       const float4* corr, cudaStream_t st) {
    DecimDev& D = pl->dec;
    auto ks = D.scan_slot_smem ? demod_scan_kernel<SF, true> : demod_scan_kernel<SF, false>;
    PSS_CUDA(ctx, cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, D.scan_smem));
    long long g2 = nf < 16LL * ctx->sm_count ? nf : 16LL * ctx->sm_count;
    PSS_CUDA(ctx, pss_launch(ks, (unsigned)g2, (unsigned)SCAN_THREADS, (size_t)D.scan_smem, st, D, iq, audio, (int)nf, F, corr));
    PSS_LAUNCH_CHECK(ctx);
    return PSS_OK;
}

int pss_decim_launch(pss_ctx* ctx, pss_demod_plan* pl, const float* iq, int64_t n_frames, float* audio,
                     const double* moments, int mom_fpb) {
    DecimDev& D = pl->dec;
    int rc;
    // Sub-batches: the forcing scratch of a sub-batch stays in L2 (<= 56 MB, a whole number of scan-kernel waves:
    // measured 0.88 / 0.79 / 0.76 / 0.72 / 0.73 ms per GiB NFM at 296 / 592 / 689 / 888 / 1184 blocks) between the
    // forcing kernel and the scan kernel, both on the context's stream.
    // PSS_DEMOD_OVERLAP=1 (measured, not the default: 0.95 / 1.12 ms per GiB NFM / WFM against 0.75 / 0.86
    // serial): two half-size scratch buffers, the scan kernel of sub-batch k on a side stream beside the
    // forcing kernel of sub-batch k+1 at 2 CTAs / SM.
    const size_t slot_b = (size_t)D.slot_doubles * 8;
    static const bool want_overlap = getenv("PSS_DEMOD_OVERLAP") != nullptr;
    long long sub = (long long)(((want_overlap ? 20u : 56u) << 20) / slot_b);
    if (!want_overlap && sub >= 3LL * ctx->sm_count) sub -= sub % (3LL * ctx->sm_count);   // whole waves of scan CTAs (3 / SM)
    static const long long sub_env = getenv("PSS_DEMOD_SUB") ? atoll(getenv("PSS_DEMOD_SUB")) : 0;      // tuning
    if (sub_env > 0) sub = sub_env;
    if (sub < 1) sub = 1;
    if (sub > n_frames) sub = n_frames;
    const bool overlap = want_overlap && n_frames > sub && !use_window_variant();
    const size_t buf_b = (size_t)sub * slot_b;
    if ((rc = pss_reserve(ctx, &pl->F_scratch, &pl->F_scratch_bytes, (overlap ? 2 : 1) * buf_b))) return rc;
    if (overlap && !pl->side) {
        PSS_CUDA(ctx, cudaStreamCreateWithFlags(&pl->side, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            PSS_CUDA(ctx, cudaEventCreateWithFlags(&pl->ev_force[i], cudaEventDisableTiming));
            PSS_CUDA(ctx, cudaEventCreateWithFlags(&pl->ev_scan[i], cudaEventDisableTiming));
        }
    }
    const float4* corr = nullptr;
    if (D.SF == 16 && D.iq_correct) {
        if ((rc = pss_reserve(ctx, &pl->corr, &pl->corr_bytes, (size_t)n_frames * 16))) return rc;
        if (!moments) {
            if ((rc = pss_reserve(ctx, &pl->mom_scratch, &pl->mom_scratch_bytes, (size_t)n_frames * 32))) return rc;
            const long long g = n_frames < 8LL * ctx->sm_count ? n_frames : 8LL * ctx->sm_count;
            demod_moments_kernel<<<(unsigned)g, 256, 0, ctx->stream>>>((const float2*)iq, D.N, n_frames, (double*)pl->mom_scratch);
            PSS_LAUNCH_CHECK(ctx);
            moments = (const double*)pl->mom_scratch;
            mom_fpb = 1;
        }
        PSS_CUDA(ctx, pss_launch(demod_corr_kernel, (unsigned)((n_frames + 127) / 128), 128u, 0, ctx->stream, moments, mom_fpb, D.N,
                                 (long long)n_frames, (float4*)pl->corr));
        PSS_LAUNCH_CHECK(ctx);
        corr = (const float4*)pl->corr;
    }
    int k = 0;
    for (int64_t f0 = 0; f0 < n_frames; f0 += sub, ++k) {
        const long long nf = n_frames - f0 < sub ? n_frames - f0 : sub;
        const float2* x = (const float2*)iq + f0 * D.N;
        float* a = audio + f0 * D.n_out * 2;
        const int b = overlap ? (k & 1) : 0;
        double* Fbuf = (double*)((char*)pl->F_scratch + (size_t)b * buf_b);
        const float4* cr = corr ? corr + f0 : nullptr;
        if (overlap && k >= 2) PSS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, pl->ev_scan[b], 0));     // buffer b is free again
        rc = D.SF == 8 ? launch_force<8>(ctx, pl, x, nf, Fbuf, cr, overlap) : launch_force<16>(ctx, pl, x, nf, Fbuf, cr, overlap);
        if (rc) return rc;
        cudaStream_t ss = ctx->stream;
        if (overlap) {
            PSS_CUDA(ctx, cudaEventRecord(pl->ev_force[b], ctx->stream));
            PSS_CUDA(ctx, cudaStreamWaitEvent(pl->side, pl->ev_force[b], 0));
            ss = pl->side;
        }
        rc = D.SF == 8 ? launch_scan<8>(ctx, pl, x, nf, a, Fbuf, cr, ss) : launch_scan<16>(ctx, pl, x, nf, a, Fbuf, cr, ss);
        if (rc) return rc;
        if (overlap) PSS_CUDA(ctx, cudaEventRecord(pl->ev_scan[b], pl->side));
    }
    if (overlap) {                                   // join: the caller's stream sees every scan finished
        PSS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, pl->ev_scan[0], 0));
        if (k >= 2) PSS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, pl->ev_scan[1], 0));
    }
    return PSS_OK;
}
