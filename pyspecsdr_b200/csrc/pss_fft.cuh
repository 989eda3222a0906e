// In-register radix-2/4/8/16 DFTs and the Stockham pass schedule used by the fused PSD kernel.
// Forward transform, exp(-2*pi*i*nk/N), same sign convention as np.fft.fft
// (reference call site: signal_processing.py:250).
#pragma once
#include "pss_common.cuh"

// Pass schedule for N = 2^LOG2N: radix-16 passes first, the remainder (2, 4 or 8) last.
__host__ __device__ constexpr int pss_pass_bits(int log2n, int p) {
    return p < log2n / 4 ? 4 : log2n - 4 * (log2n / 4);
}
__host__ __device__ constexpr int pss_num_passes(int log2n) { return (log2n + 3) / 4; }

// After fft<R>(v), register position p holds output bin fft_perm<R>(p).
template <int R>
__host__ __device__ constexpr int fft_perm(int p) {
    return R == 16 ? (p >> 2) + 4 * (p & 3) : R == 8 ? (p >> 2) + 2 * (p & 3) : p;
}

// Shared-memory swizzle for 16-byte complex elements: keeps every quarter-warp of a 128-bit
// access on 8 distinct 16-byte bank groups for all passes (checked exhaustively on the host,
// see DESIGN.md "FFT exchange layout").
__device__ __forceinline__ int fft_swz(int i) { return i ^ ((i >> 4) & 7); }

// Alternative layout used by the large-transform kernel: one 16-byte pad element after every 16 complex
// elements (element i at i + i/16).  Every access of a pass is then a per-thread base plus a compile-time
// offset (the XOR swizzle costs ~3 integer instructions per access), also conflict-free; measured -4 ... -7 %
// on the fused large transforms, +-0 ... +4 % on the one-frame-per-CTA kernels, which keep the XOR form.
__device__ __forceinline__ int fft_pad(int i) { return i + (i >> 4); }
__host__ __device__ constexpr int fft_padded(int n) { return n + n / 16; }

template <typename T>
__device__ __forceinline__ void bfly2(cx<T>& a, cx<T>& b) {
    cx<T> t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

// 4-point DFT, natural order in and out.
template <typename T>
__device__ __forceinline__ void bfly4(cx<T>& a, cx<T>& b, cx<T>& c, cx<T>& d) {
    cx<T> t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = csub(b, d);
    a = cadd(t0, t2);
    c = csub(t0, t2);
    b = {t1.x + t3.y, t1.y - t3.x};   // t1 - i*t3
    d = {t1.x - t3.y, t1.y + t3.x};   // t1 + i*t3
}

template <typename T>
__device__ __forceinline__ cx<T> mul_w8_1(cx<T> a) {   // * (1 - i)/sqrt(2)
    const T h = (T)0.70710678118654752440;
    return {(a.x + a.y) * h, (a.y - a.x) * h};
}
template <typename T>
__device__ __forceinline__ cx<T> mul_w8_3(cx<T> a) {   // * (-1 - i)/sqrt(2)
    const T h = (T)0.70710678118654752440;
    return {(a.y - a.x) * h, -(a.x + a.y) * h};
}
template <typename T>
__device__ __forceinline__ cx<T> mul_mi(cx<T> a) { return {a.y, -a.x}; }   // * (-i)

template <int R, typename T>
struct fft_regs;

template <typename T>
struct fft_regs<2, T> {
    static __device__ __forceinline__ void run(cx<T>* v) { bfly2(v[0], v[1]); }
};
template <typename T>
struct fft_regs<4, T> {
    static __device__ __forceinline__ void run(cx<T>* v) { bfly4(v[0], v[1], v[2], v[3]); }
};
template <typename T>
struct fft_regs<8, T> {
    static __device__ __forceinline__ void run(cx<T>* v) {
        // k = m + 2n: radix-2 over (i, i+4), twiddle W8^(i*m), radix-4 over i
#pragma unroll
        for (int i = 0; i < 4; ++i) bfly2(v[i], v[i + 4]);
        v[5] = mul_w8_1(v[5]);
        v[6] = mul_mi(v[6]);
        v[7] = mul_w8_3(v[7]);
        bfly4(v[0], v[1], v[2], v[3]);
        bfly4(v[4], v[5], v[6], v[7]);
    }
};
template <typename T>
struct fft_regs<16, T> {
    static __device__ __forceinline__ void run(cx<T>* v) {
        // k = m + 4n: radix-4 over (i, i+4, i+8, i+12) -> y[i][m] at v[i+4m];
        // twiddle W16^(i*m); radix-4 over i for each m -> X[m+4n] at v[4m+n].
        const T C = (T)0.92387953251128675613, S = (T)0.38268343236508977173;
#pragma unroll
        for (int i = 0; i < 4; ++i) bfly4(v[i], v[i + 4], v[i + 8], v[i + 12]);
        // m = 1: i = 1,2,3 -> W16^1, W16^2, W16^3
        v[5] = cmul(v[5], cx<T>{C, -S});
        v[6] = mul_w8_1(v[6]);
        v[7] = cmul(v[7], cx<T>{S, -C});
        // m = 2: W16^2, W16^4, W16^6
        v[9] = mul_w8_1(v[9]);
        v[10] = mul_mi(v[10]);
        v[11] = mul_w8_3(v[11]);
        // m = 3: W16^3, W16^6, W16^9
        v[13] = cmul(v[13], cx<T>{S, -C});
        v[14] = mul_w8_3(v[14]);
        v[15] = cmul(v[15], cx<T>{-C, S});
#pragma unroll
        for (int m = 0; m < 4; ++m) bfly4(v[4 * m], v[4 * m + 1], v[4 * m + 2], v[4 * m + 3]);
    }
};

// v[r] *= w^r for r = 1..R-1, powers built by a depth-4 product tree from the table value w.
template <int R, typename T>
__device__ __forceinline__ void twiddle_apply(cx<T>* v, const cx<T> w1) {
    v[1] = cmul(v[1], w1);
    if constexpr (R >= 4) {
        const cx<T> w2 = csqr(w1);
        const cx<T> w3 = cmul(w2, w1);
        v[2] = cmul(v[2], w2);
        v[3] = cmul(v[3], w3);
        if constexpr (R >= 8) {
            const cx<T> w4 = csqr(w2);
            const cx<T> w5 = cmul(w4, w1), w6 = csqr(w3), w7 = cmul(w4, w3);
            v[4] = cmul(v[4], w4);
            v[5] = cmul(v[5], w5);
            v[6] = cmul(v[6], w6);
            v[7] = cmul(v[7], w7);
            if constexpr (R >= 16) {
                const cx<T> w8 = csqr(w4);
                v[8] = cmul(v[8], w8);
                v[9] = cmul(v[9], cmul(w8, w1));
                v[10] = cmul(v[10], csqr(w5));
                v[11] = cmul(v[11], cmul(w8, w3));
                v[12] = cmul(v[12], csqr(w6));
                v[13] = cmul(v[13], cmul(w8, w5));
                v[14] = cmul(v[14], csqr(w7));
                v[15] = cmul(v[15], cmul(w8, w7));
            }
        }
    }
}
