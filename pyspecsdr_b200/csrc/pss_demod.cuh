// Shared by the demodulation translation units: plan object, device-side descriptors, the float32
// discriminator (numpy's evaluation order) and the blocked linear scan used by the AM kernel.
#pragma once
#include <math.h>

#include "pss_common.cuh"

#define EDGE 27
#define FORCE_THREADS 256        // forcing kernel: 8 independent warps per CTA
#ifndef SCAN_THREADS
#define SCAN_THREADS 256         // scan kernel: one CTA of 8 warps per block of audio (-DSCAN_THREADS=128: 4 warps, same 3 CTAs/SM -- measured +3 % on the demodulator)
#endif
#define SCAN_CPL 10              // chunks per lane of the scan kernel (32 * SCAN_CPL chunks per segment)
#define WBUF_FLOATS 1088         // per-warp discriminator window of the forcing kernel
#define SLAB_ROW 132             // row stride of the 8-row slab layout (== 4 mod 8: conflict-free A fragments)
#define SLAB_K 128               // samples per row and slab (32 k-steps)

// Decimating plans (NFM / WFM), pss_demod_decim.cu.
struct DecimDev {
    int mode, N, L, q, n_out, lead, SF, SB, n_body, m_tail, tail_start, tail_len;
    int Kp, KS, NTD, rows, CS;                  // NTD = (SF + SB) / 8 tensor n-tiles; CS = chunk stride of a scratch row
    int tail_pad, scan_warps, scan_slot_smem;
    int groups;                                 // chunk groups of 8 per block
    int contiguous;                             // 1: one contiguous window per group; 0: 8 row slabs per k-slab
    int iq_correct;                             // WFM: 1 = iq_correction fused in front of the discriminator
    float scale, norm;
    int tab_in_smem;
    int force_smem, scan_smem, fused_tab_smem;  // dynamic shared memory of the kernels (fused: 0 = table stays global)
    const double* tabP;                         // packed fused-kernel table: [ks][pair][lane] double2 = (b0, b1), (b2 | r tap, r tap | 0)
    int packed_tab_smem;
    long long slot_doubles;                     // doubles of scratch per block: rows * CS
    const double *tabF;                         // fragment-ordered body table [KS][NTD][32] then the r row [KS][4]
    const double *head, *tailT, *tailM;
    const double *scanTab;                      // packed scan tables (ScanTabLayout)
    int scan_tab_doubles;
    double DB;
};

// Packed scan tables (doubles), see pss_demod_decim.cu: for the forward (nf = SF/2 blocks) and backward
// (nb = SB/2) recurrences: the 2x2 blocks B and M^(2^s) for s = 0..5 with M = B^SCAN_CPL; then G [SB][SF],
// CR [SF], CB [SB].
struct ScanTabLayout {
    int nf, nb;
    __host__ __device__ int BF() const { return 0; }
    __host__ __device__ int PWF() const { return BF() + nf * 4; }
    __host__ __device__ int BB() const { return PWF() + 6 * nf * 4; }
    __host__ __device__ int PWB() const { return BB() + nb * 4; }
    __host__ __device__ int G() const { return PWB() + 6 * nb * 4; }
    __host__ __device__ int CR() const { return G() + 2 * nb * 2 * nf; }
    __host__ __device__ int CB() const { return CR() + 2 * nf; }
    __host__ __device__ int total() const { return CB() + 2 * nb; }
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
    void* F_scratch = nullptr;          // forcing / state slots of one sub-batch of blocks (stays in L2)
    size_t F_scratch_bytes = 0;
    void* corr = nullptr;               // per-block iq_correction coefficients (float4)
    size_t corr_bytes = 0;
    void* mom_scratch = nullptr;        // per-block I/Q second moments when the caller supplies none
    size_t mom_scratch_bytes = 0;
    cudaStream_t side = nullptr;        // scan kernels of sub-batch k overlap the forcing kernel of k+1
    cudaEvent_t ev_force[2] = {nullptr, nullptr}, ev_scan[2] = {nullptr, nullptr};
    void* tile_scratch = nullptr;       // per-block max|y| of tiled FIR plans
    size_t tile_scratch_bytes = 0;
};

int pss_demod_upload(pss_ctx* ctx, pss_demod_plan* pl, const void* src, size_t bytes, const void** dst);
int pss_decim_create(pss_ctx* ctx, const pss_demod_desc* d, pss_demod_plan* pl);
int pss_decim_launch(pss_ctx* ctx, pss_demod_plan* pl, const float* iq, int64_t n_frames, float* audio,
                     const double* moments, int mom_fpb);

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
