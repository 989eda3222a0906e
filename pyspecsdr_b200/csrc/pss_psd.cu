// Fused PSD kernel: window -> Stockham FFT (radix-16 passes through shared memory) -> fftshift ->
// 10*log10(|X|^2 + 1e-10) [-> 5-bin smoothing, median-10 dB clamp, row statistics, W-column resample]
// in ONE launch, one frame per group of N/16 threads.
//
// Replaces compute_fft (signal_processing.py:243-264), the main-loop epilogue
// (pyspecsdr.py:2278-2283), the header read-outs (pyspecsdr.py:388-389), the np.interp column
// resample of the draw_* functions (pyspecsdr.py:1378-1382 etc.) and the scanner's per-step
// spectrum + peak + above-threshold count (pyspecsdr.py:2542-2552, 1050-1057).
//
// Precision: the window multiply and every butterfly are fp64 (the reference runs c64*f64 -> c128
// pocketfft); SURVEY.md §7.2 shows an fp32 transform misses the 1e-4 dB bar by 10-4000x whenever a
// strong tone is present.  Only the |X|^2 -> dB tail is fp32.
#include <math.h>

#include "pss_fft.cuh"

enum { EPI_RAW = 0, EPI_SMOOTH = 1, EPI_SCAN = 2 };
// Threads per CTA of the small transforms: one frame per CTA down to one warp (N = 512), several frames per
// warp-CTA below.  Until late in round 2 every CTA had 256 threads (4 frames of 1024 points, 8 of 512): the same
// warps per SM, but every barrier then waits for 8 warps instead of 1, 2 or 4.  Same box, ms per GiB, raw /
// smoothing / scanner: 512: 0.528 / 1.012 / 0.616 -> 0.456 / 0.866 / 0.517; 1024: 0.519 / 0.891 / 0.632 ->
// 0.440 / 0.818 / 0.527; 2048: 0.537 / 0.874 / 0.641 -> 0.479 / 0.785 / 0.558 (-DPSS_PSD_MIN_THREADS=256|128|64
// rebuilds the other points of the comparison).
#ifndef PSS_PSD_MIN_THREADS
#define PSS_PSD_MIN_THREADS 32
#endif
// CTAs per SM the 4096-point smoothing variant is compiled for (-DPSS_SMOOTH_MINB=n overrides)
#ifdef PSS_SMOOTH_MINB
constexpr int pss_smooth_minb(int) { return PSS_SMOOTH_MINB; }
#else
constexpr int pss_smooth_minb(int) { return 2; }
#endif

struct PsdParams {
    const float2* iq;
    const void* window;   // T[N] or nullptr
    const void* tw;       // cx<T> per-pass base twiddles
    long long n_frames;
    float* db;
    float* cols;
    int W;
    float* stats;
    float* peak;
    int* count;
    int use_abs;
    double thr_pow;       // scanner threshold in the power domain: 10^(thr/10) (absolute) or 10^(-rel/10) (x peak)
    const void* ystage;   // second stage of a large transform: cx<T> rows [n_frames][N] (LOG2N1 > 0)
    double* moments;      // [n_frames][4] sum I^2, Q^2, IQ of each frame (by-product for the WFM demod), or null
    int ahead;            // CTAs resident on the device: CTA b prefetches the frames of CTA b + ahead into L2
    double col_step;      // (n - 1) / (W - 1) of the column resample
};

template <int LOG2N, typename T>
struct PsdCfg {
    static constexpr int N = 1 << LOG2N;
    static constexpr int TPF = N / 16;                       // threads per frame
    static constexpr int THREADS = TPF > PSS_PSD_MIN_THREADS ? TPF : PSS_PSD_MIN_THREADS;
    static constexpr int FPC = THREADS / TPF;                // frames per CTA
    // ~120 registers per thread: 512 threads per SM whatever the CTA size; shared memory scales with the threads
    static constexpr int MINB = (sizeof(T) * 2 * N * FPC > 64 * 1024) ? 1 : (512 / THREADS < 1 ? 1 : 512 / THREADS);
    static constexpr size_t SMEM = (size_t)FPC * N * sizeof(cx<T>);
    static constexpr int NP = pss_num_passes(LOG2N);
    static constexpr int CAP = TPF < 64 ? TPF : 64;          // median candidates ranked directly (EPI_SMOOTH)
};


// Every warp scans a 256-bin shared histogram by itself (no broadcast barrier): finds the bin holding
// the element of 0-based rank `rank`, returns the bin, its count and the rank within the bin.
__device__ __forceinline__ void hist_pick(const unsigned* h, unsigned& rank, unsigned& digit, unsigned& count,
                                          const int lane) {
    const uint4 a = reinterpret_cast<const uint4*>(h)[2 * lane];
    const uint4 b = reinterpret_cast<const uint4*>(h)[2 * lane + 1];
    const unsigned c[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    unsigned sum = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) sum += c[q];
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, incl > rank);
    const int L = hit ? __ffs(hit) - 1 : 31;
    unsigned r = rank - (incl - sum), dg = 7, cc = c[7];
    bool found = false;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (!found) {
            if (r < c[q]) { dg = q; cc = c[q]; found = true; }
            else if (q < 7) r -= c[q];
        }
    }
    digit = __shfl_sync(0xffffffffu, 8u * lane + dg, L);
    rank = __shfl_sync(0xffffffffu, r, L);
    count = __shfl_sync(0xffffffffu, cc, L);
}


// The same for a BINS-bin histogram (512 or 1024), run by ONE warp: stage 1 picks the segment of BINS/32
// consecutive bins (lane L sums segment L, reading it in an order rotated by L so that the 32 lanes hit 32
// banks), stage 2 the bin inside it (one bin per lane).
template <int BINS>
__device__ __forceinline__ void hist_pick_wide(const unsigned* h, unsigned& rank, unsigned& digit, unsigned& count,
                                               const int lane) {
    constexpr int SEG = BINS / 32;
    const unsigned* seg = h + SEG * lane;
    const int rot = SEG == 32 ? lane : (lane >> 1);
    unsigned sum = 0;
#pragma unroll
    for (int j = 0; j < SEG; ++j) sum += seg[(j + rot) & (SEG - 1)];
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, incl > rank);
    const int L = hit ? __ffs(hit) - 1 : 31;
    const unsigned r = rank - __shfl_sync(0xffffffffu, incl - sum, L);
    const unsigned c = lane < SEG ? h[SEG * L + lane] : 0u;
    unsigned inc2 = c;
#pragma unroll
    for (int o = 1; o < SEG; o <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, inc2, o);
        if (lane >= o) inc2 += up;
    }
    const unsigned hit2 = __ballot_sync(0xffffffffu, lane < SEG && inc2 > r);
    const int l2 = hit2 ? __ffs(hit2) - 1 : SEG - 1;
    digit = (unsigned)(SEG * L + l2);
    count = __shfl_sync(0xffffffffu, c, l2);
    rank = r - __shfl_sync(0xffffffffu, inc2 - c, l2);
}


// Rare path of the epilogue's median (crowded bucket: flat rows): 4 x 8-bit radix select on order-preserving
// keys normalised to the occupied key range.  A real call, so that its ~1000 instructions stay out of the hot
// code (the fused kernel is ~60 KB of SASS and its speed moves by 3 % with its placement in the instruction
// cache); it re-reads the thread's 16 smoothed values from shared memory instead of taking them in registers.
// Called by every thread of the CTA (all frames); hist [1024], uf [8] ([0] = [3] = 0xffffffff, rest 0 on entry).
template <int N, int TPF>
__device__ __noinline__ void rare_median(const float* __restrict__ srow, unsigned* hist, unsigned* uf, const int t,
                                         float& v1, float& v2) {
    constexpr int n = N - 4;
    const int lane = t & 31;
    unsigned key[16];
    bool gvalid[4];
    unsigned kmin = 0xffffffffu, kmax = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i0 = 4 * t + 4 * TPF * q;
        gvalid[q] = i0 < n;
        const float4 v = *reinterpret_cast<const float4*>(srow + i0);
        key[4 * q] = f2key(v.x);
        key[4 * q + 1] = f2key(v.y);
        key[4 * q + 2] = f2key(v.z);
        key[4 * q + 3] = f2key(v.w);
        if (gvalid[q]) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                kmin = min(kmin, key[4 * q + e]);
                kmax = max(kmax, key[4 * q + e]);
            }
        }
    }
    for (int b = t; b < 1024; b += TPF) hist[b] = 0u;
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0) {
        atomicMin(&uf[0], kmin);
        atomicMax(&uf[1], kmax);
    }
    __syncthreads();
    kmin = uf[0];
    kmax = uf[1];
    const int common = min(__clz((int)(kmin ^ kmax)), 31);
#pragma unroll
    for (int m2 = 0; m2 < 16; ++m2) key[m2] = (key[m2] - kmin) << common;
    unsigned rk = (unsigned)((n - 1) / 2), prefix = 0u, dg, cnt;
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
        const int shift = 24 - 8 * ps;
        unsigned* h = hist + ps * 256;
#pragma unroll
        for (int m2 = 0; m2 < 16; ++m2) {
            const bool match = ps == 0 ? true : (key[m2] >> (shift + 8)) == prefix;
            if (gvalid[m2 >> 2] && match) atomicAdd(&h[(key[m2] >> shift) & 255u], 1u);
        }
        __syncthreads();
        hist_pick(h, rk, dg, cnt, lane);
        prefix = (prefix << 8) | dg;
    }
    const unsigned key1n = prefix;                   // normalised key of the lower median
    unsigned cnt_le = 0, min_gt = 0xffffffffu;
#pragma unroll
    for (int m2 = 0; m2 < 16; ++m2) {
        if (gvalid[m2 >> 2]) {
            cnt_le += key[m2] <= key1n;
            if (key[m2] > key1n) min_gt = min(min_gt, key[m2]);
        }
    }
    cnt_le = __reduce_add_sync(0xffffffffu, cnt_le);
    min_gt = __reduce_min_sync(0xffffffffu, min_gt);
    if (lane == 0) {
        atomicAdd(&uf[2], cnt_le);
        atomicMin(&uf[3], min_gt);
    }
    __syncthreads();
    const unsigned key2n = ((n & 1) || uf[2] > (unsigned)(n / 2)) ? key1n : uf[3];
    v1 = key2f((key1n >> common) + kmin);
    v2 = key2f((key2n >> common) + kmin);
}

// The main-loop epilogue on one dB row held in shared memory (pyspecsdr.py:2278-2283, 388-389, and the
// W-column np.interp resample of the draw_* functions): 5-bin 'valid' mean, exact median - 10 dB clamp, row
// statistics, stores.  Called by every thread of the frame's group of TPF = N/16 threads; `row` [N] raw dB
// (fft-shifted), `srow` [N] and `hist` [1024] frame-local shared scratch; (rsum, rsq, rnan) = sum, sum of squares and NaN
// flag of the raw values this thread produced or loaded (together the threads cover the whole row).
template <int N, int TPF>
__device__ __forceinline__ void smooth_epilogue(const float* __restrict__ row, float* __restrict__ srow, unsigned* hist,
                                                unsigned* us, unsigned* uf, float* cand, double* dscr, float* fscr,
                                                const float rsum, const float rsq, const bool rnan, const int t,
                                                const bool live, const long long frame, const PsdParams& p) {
        // Thread t owns four groups of 4 consecutive bins, group q at 4*t + 4*TPF*q: float4 shared and
        // global accesses with a 16-byte lane stride (conflict-free, fully coalesced).
        constexpr int n = N - 4;
        constexpr int CAP = TPF < 64 ? TPF : 64;
        constexpr int BINS = N >= 1024 ? 1024 : 512;
        const int wf = t >> 5, lane = t & 31, nw = TPF / 32;
        {   // mean and spread of the raw row (warp partials; folded by every thread after the row barrier), NaN flag
            // (measured: fixed-point totals through one REDUX and one shared atomic per warp are 0.3 % slower)
            float ws = rsum, wq = rsq;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ws += __shfl_xor_sync(0xffffffffu, ws, o);
                wq += __shfl_xor_sync(0xffffffffu, wq, o);
            }
            if (lane == 0) {
                fscr[wf] = ws;
                fscr[16 + wf] = wq;
            }
            if (__any_sync(0xffffffffu, rnan) && lane == 0) atomicOr(&us[6], 1u);
        }
        if (t < CAP) cand[t] = INFINITY;                         // unclaimed candidate slots never count in the ranking
        for (int b = t; b < BINS / 4; b += TPF)                  // the exchange buffer is dead after the last pass
            reinterpret_cast<uint4*>(hist)[b] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();                                     // B1: row complete, histogram clear, partials visible
        float s[16];
        unsigned b16[16];
        bool gvalid[4];
        {
            // The median of any data lies within one standard deviation of its mean, and the 5-bin means have
            // (up to edge terms) the raw row's mean and no more spread: BINS buckets over mean +- 1.25 sigma of
            // the raw row put a handful of elements in the median's bucket.  The bucket function only has to
            // be monotone -- whatever falls outside lands in the end buckets, and a crowded bucket takes the
            // exact fallback below -- so the estimate affects speed, never the result.
            float S = 0.f, Q = 0.f;
            if constexpr (TPF >= 128) {
#pragma unroll
                for (int w = 0; w < TPF / 32; w += 4) {
                    const float4 a = *reinterpret_cast<const float4*>(fscr + w);
                    const float4 b = *reinterpret_cast<const float4*>(fscr + 16 + w);
                    S += (a.x + a.y) + (a.z + a.w);
                    Q += (b.x + b.y) + (b.z + b.w);
                }
            } else {
#pragma unroll
                for (int w = 0; w < TPF / 32; ++w) {
                    S += fscr[w];
                    Q += fscr[16 + w];
                }
            }
            // (measured, same box: folding in the last warp to arrive and broadcasting through shared memory is
            // 4 % slower -- it lengthens the path to the barrier; a min-of-xor quick reject before the match mask
            // below 2 % slower; this fused multiply-add form of the bucket 0.4 % faster than (v - lo) * scale)
            const float mean = S * (1.0f / N);
            const float var = Q * (1.0f / N) - mean * mean;
#ifdef PSS_V_SQRT
            const float sd = sqrtf(fmaxf(var, 0.f));
            const float scale = sd > 0.f ? (BINS / 2.5f) / sd : 0.f;
            const float off = scale > 0.f ? (1.25f * sd - mean) * scale : 0.f;      // bucket = v * scale + off
#else
            const float rs = rsqrtf(var);                                          // approximate: speed only
            const float scale = var > 0.f ? (BINS / 2.5f) * rs : 0.f;
            const float off = var > 0.f ? fmaf(-mean, scale, BINS / 2.0f) : 0.f;   // bucket = v * scale + off
#endif
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i0 = 4 * t + 4 * TPF * q;
                gvalid[q] = i0 < n;
                const float4 a = *reinterpret_cast<const float4*>(row + i0);
                const float4 c = i0 + 4 < N ? *reinterpret_cast<const float4*>(row + i0 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float d[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float v = ((d[e] + d[e + 1]) + (d[e + 2] + d[e + 3]) + d[e + 4]) * 0.2f;
                    s[4 * q + e] = v;
                    // monotone bucket (NaN and values below the range convert to 0)
                                        b16[4 * q + e] = min((unsigned)(BINS - 1), __float2uint_rz(fmaf(v, scale, off)));
                }
                // unclamped smoothed row: the candidate ranking and the column resample read it (the clamp is
                // applied on the fly there), this thread's own copy stays in registers
                *reinterpret_cast<float4*>(srow + i0) = make_float4(s[4 * q], s[4 * q + 1], s[4 * q + 2], s[4 * q + 3]);
            }
        }
        // exact lower/upper median: one histogram over the buckets, then the few elements of the selected
        // bucket are ranked directly (crowded buckets -- flat rows -- fall back to a key radix select)
        unsigned rank = (unsigned)((n - 1) / 2), sel, m;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (gvalid[q]) {
#pragma unroll
                for (int e = 0; e < 4; ++e) atomicAdd(&hist[b16[4 * q + e]], 1u);
            }
        __syncthreads();                                     // B2
#ifdef PSS_PICK_ALL_WARPS
        hist_pick_wide<BINS>(hist, rank, sel, m, lane);      // measured variant: every warp scans, no barrier
#else
        if (wf == 0) {                                       // one warp scans the histogram for the frame
            hist_pick_wide<BINS>(hist, rank, sel, m, lane);
            if (lane == 0) {
                us[8] = rank;
                us[9] = sel;
                us[10] = m;
            }
        }
        __syncthreads();                                     // B3
        rank = us[8];
        sel = us[9];
        m = us[10];
#endif
        const bool flat = m > (unsigned)CAP;
        {
            // the elements of the selected bucket: one bit per element, one slot claim per thread that has any
            unsigned eqm = 0u;
#pragma unroll
            for (int i = 0; i < 16; ++i) eqm |= (gvalid[i >> 2] && b16[i] == sel ? 1u : 0u) << i;
            if (!flat && eqm) {
                unsigned slot = atomicAdd(&us[2], (unsigned)__popc(eqm));
                while (eqm) {
                    const int i = __ffs(eqm) - 1;
                    eqm &= eqm - 1u;
                    if (slot < (unsigned)CAP) cand[slot] = srow[4 * t + 4 * TPF * (i >> 2) + (i & 3)];
                    ++slot;
                }
            }
            if (rank + 1u >= m) {                            // frame-uniform, rare: upper median lies above the bucket
                float fgt = INFINITY;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (gvalid[i >> 2] && b16[i] > sel) fgt = fminf(fgt, s[i]);
                const unsigned kgt = __reduce_min_sync(0xffffffffu, f2key(fgt));
                if (lane == 0) atomicMin(&us[3], kgt);
            }
        }
        float v1, v2;
        if (!__syncthreads_or(flat)) {                       // B4
#ifndef PSS_RANK_WARPS
            if ((unsigned)t < m) {                           // thread t ranks candidate t against all (the other warps wait)
                const float c = cand[t];
                unsigned rk = 0;
#ifdef PSS_V_RANK1
                for (unsigned j = 0; j < m; ++j) {
                    const float o = cand[j];
                    rk += (o < c) || (o == c && j < (unsigned)t);
                }
#else
                for (unsigned j = 0; j < m; j += 4) {        // slots >= m hold +inf: they never count
                    const float4 o = *reinterpret_cast<const float4*>(cand + j);
                    rk += (o.x < c) || (o.x == c && j < (unsigned)t);
                    rk += (o.y < c) || (o.y == c && j + 1 < (unsigned)t);
                    rk += (o.z < c) || (o.z == c && j + 2 < (unsigned)t);
                    rk += (o.w < c) || (o.w == c && j + 3 < (unsigned)t);
                }
#endif
                if (rk == rank) us[4] = __float_as_uint(c);
                if (rk == rank + 1u) us[5] = __float_as_uint(c);
            }
#else
            {   // measured variant (slower: +2 % kernel time, the extra instructions of 7 more warps cost more than
                // the one warp's serial chain): warp w ranks candidates w, w + nw, ... with ballots
                const float o0 = (unsigned)lane < m ? cand[lane] : INFINITY;
                const float o1 = (unsigned)lane + 32u < m ? cand[lane + 32] : INFINITY;
                for (unsigned i = (unsigned)wf; i < m; i += (unsigned)nw) {
                    const float c = cand[i];
                    const bool p0 = (unsigned)lane < m && ((o0 < c) || (o0 == c && (unsigned)lane < i));
                    const bool p1 = (unsigned)lane + 32u < m && ((o1 < c) || (o1 == c && (unsigned)lane + 32u < i));
                    const unsigned rk = __popc(__ballot_sync(0xffffffffu, p0)) + (CAP > 32 ? __popc(__ballot_sync(0xffffffffu, p1)) : 0);
                    if (lane == 0) {
                        if (rk == rank) us[4] = __float_as_uint(c);
                        if (rk == rank + 1u) us[5] = __float_as_uint(c);
                    }
                }
            }
#endif
            __syncthreads();                                 // B5
            v1 = __uint_as_float(us[4]);
            v2 = (n & 1) ? v1 : (rank + 1u < m ? __uint_as_float(us[5]) : key2f(us[3]));
        } else {
            // ---- rare path (a frame of this CTA has > CAP equal-bucket elements); every frame of the CTA runs it
            rare_median<N, TPF>(srow, hist, uf, t, v1, v2);
        }
        float thr = (float)(0.5 * ((double)v1 + (double)v2) - 10.0);
        const bool any_nan = us[6] != 0u;
        if (any_nan) thr = __int_as_float(0x7fc00000);       // np.median propagates NaN
        // clamp, row statistics, store (measured: taking the statistics of the unclamped values early and
        // redoing them only in threads that hold a value below the threshold is 0.5 % slower, not faster)
        float mx = -INFINITY, mn = INFINITY, fsum = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float v = s[4 * q + e];
                if (v < thr) v = thr;
                s[4 * q + e] = v;
                if (gvalid[q]) {
                    mx = fmaxf(mx, v);
                    mn = fminf(mn, v);
                    fsum += v;
                }
            }
            if (live && p.db && gvalid[q])
                *reinterpret_cast<float4*>(p.db + frame * n + 4 * t + 4 * TPF * q) =
                    make_float4(s[4 * q], s[4 * q + 1], s[4 * q + 2], s[4 * q + 3]);
        }
        {   // warp partials of the row statistics; after the barrier the frame's last warp (which has the least
            // column-resample work) folds them while the others resample.  (Measured, same box: folding in
            // whichever warp arrives last on a shared ticket counter, with fences instead of the barrier, is
            // 1.8 % faster -- and is reported by compute-sanitizer's racecheck, which only models barriers;
            // accumulating with shared atomics instead is 3.5 % slower.)
            const unsigned kx = __reduce_max_sync(0xffffffffu, f2key(mx));
            const unsigned kn = __reduce_min_sync(0xffffffffu, f2key(mn));
            const double dsum = warp_sum((double)fsum);
            if (lane == 0) {
                fscr[wf] = key2f(kx);
                fscr[16 + wf] = key2f(kn);
                dscr[wf] = dsum;
            }
        }
        __syncthreads();                                     // B6: warp partials visible
        if (live && p.stats && wf == nw - 1) {               // lane w takes warp w's partials
            float a = lane < nw ? fscr[lane] : -INFINITY, b = lane < nw ? fscr[16 + lane] : INFINITY;
            double sm = lane < nw ? dscr[lane] : 0.0;
            a = key2f(__reduce_max_sync(0xffffffffu, f2key(a)));
            b = key2f(__reduce_min_sync(0xffffffffu, f2key(b)));
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);     // nw <= 16
            if (lane == 0) {
                const float nanv = __int_as_float(0x7fc00000);
                float4 st;
                st.x = any_nan ? nanv : a;                        // np.max
                st.y = any_nan ? nanv : (float)(sm * (1.0 / n));  // np.mean
                st.z = b;                                         // finite min
                st.w = a;                                         // finite max
                reinterpret_cast<float4*>(p.stats)[frame] = st;
            }
        }
        if (live && p.cols) {
            const int W = p.W;
            const double step = p.col_step;                  // (n - 1) / (W - 1), 0 for W == 1
            for (int c = t; c < W; c += TPF) {
                const double x = (c == W - 1 && W > 1) ? (double)(n - 1) : c * step;
                const int j = (int)x;
                float o;
                if (j >= n - 1) {
                    o = srow[n - 1];
                    if (o < thr) o = thr;
                } else {
                    float f0 = srow[j], f1 = srow[j + 1];
                    if (f0 < thr) f0 = thr;
                    if (f1 < thr) f1 = thr;
                    const double y0 = f0, y1 = f1;
                    o = (float)((y1 - y0) * (x - (double)j) + y0);
                }
                p.cols[frame * W + c] = o;
            }
        }
}

// One Stockham pass P >= 1 (shared -> registers -> shared, or -> `out` on the last pass).
struct CtaBarrier {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
// Named barrier over one 256-thread half of a 512-thread CTA (ids 1 and 2), so the two halves run
// their row transforms out of phase like two independent CTAs would.
struct HalfBarrier {
    int id;
    __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }
};

template <int LOG2N, int P, typename T, bool PAD = false, typename OutF, typename Bar = CtaBarrier>
__device__ __forceinline__ void stockham_pass(cx<T>* __restrict__ buf, const cx<T>* __restrict__ wpre,
                                              const int t, OutF&& out, Bar bar = Bar()) {
    constexpr int N = 1 << LOG2N, TPF = N / 16;
    constexpr int BITS = pss_pass_bits(LOG2N, P), R = 1 << BITS, NS = 1 << (4 * P);
    constexpr int TI = N / R, ITEMS = 16 / R, NP = pss_num_passes(LOG2N);
    cx<T> v[16];
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const cx<T>* rb = buf + fft_pad(t + it * TPF);           // PAD: per-thread base + compile-time offsets
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if constexpr (PAD && TI % 16 == 0) v[it * R + r] = rb[fft_padded(r * TI)];
            else v[it * R + r] = buf[fft_swz(t + it * TPF + r * TI)];
        }
    }
    bar();
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int j = t + it * TPF;
        const int k = j & (NS - 1);
        twiddle_apply<R, T>(v + it * R, wpre[it]);
        fft_regs<R, T>::run(v + it * R);
        const int base = ((j - k) << BITS) + k;
        cx<T>* wb = buf + fft_pad(base);
#pragma unroll
        for (int p = 0; p < R; ++p) {
            const int idx = base + fft_perm<R>(p) * NS;
            if constexpr (P == NP - 1)
                out(idx, v[it * R + p]);
            else if constexpr (PAD)
                wb[fft_padded(fft_perm<R>(p) * NS)] = v[it * R + p];      // NS = 16^P: a multiple of 16
            else
                buf[fft_swz(idx)] = v[it * R + p];
        }
    }
    if constexpr (P != NP - 1) bar();
}

// LOG2N1 > 0: this launch is the second stage of a length N*2^LOG2N1 transform (four-step FFT): frame
// index = big_frame * N1 + k1, input = column-transformed, twiddled fp64 rows, output bin = k1 + N1*k2.
template <int LOG2N, typename T, int EPI, int LOG2N1 = 0>
// Occupancy: ~120 registers per thread, i.e. 512 threads per SM whatever the CTA size (PsdCfg::MINB).  The
// 4096-point smoothing variant (the only 2-CTA configuration left since the small transforms run one frame per
// CTA) went back and forth with its median: 3 CTAs/SM at 80 registers with the 13-barrier key radix select,
// 2 CTAs/SM (+6 %) with the 6-barrier two-level histogram, 3 again (+2.5 %) with the first single-histogram
// version, and 2 (+2.7 %) once one warp scans the histogram; `tools/build_variant.sh minb3 -DPSS_SMOOTH_MINB=3`
// rebuilds the other side of the comparison.
__global__ void __launch_bounds__(PsdCfg<LOG2N, T>::THREADS,
                                  (EPI == EPI_SMOOTH && PsdCfg<LOG2N, T>::MINB == 2) ? pss_smooth_minb(LOG2N) : PsdCfg<LOG2N, T>::MINB)
psd_kernel(const PsdParams p) {
    using C = PsdCfg<LOG2N, T>;
    constexpr int N = C::N, TPF = C::TPF, NP = C::NP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int f = tid / TPF, t = tid % TPF;
    const long long frame = (long long)blockIdx.x * C::FPC + f;
    const bool live = frame < p.n_frames;
    cx<T>* buf = reinterpret_cast<cx<T>*>(smem_raw) + (size_t)f * N;
    const cx<T>* tw = reinterpret_cast<const cx<T>*>(p.tw);
    // base twiddles of every pass depend only on the thread: fetch them before anything else so the
    // loads are never exposed behind a barrier
    cx<T> wpre[3][8];
#pragma unroll
    for (int ps = 1; ps < NP; ++ps) {
        const int bits = pss_pass_bits(LOG2N, ps), ns = 1 << (4 * ps), items = 16 >> bits;
#pragma unroll
        for (int it = 0; it < 8; ++it)
            if (it < items) wpre[ps - 1][it] = tw[(ns - 16) / 15 + ((t + it * TPF) & (ns - 1))];
    }

    // frame-local scratch that aliases the exchange buffer once the last pass has read it
    float* row = reinterpret_cast<float*>(buf);           // [N] dB, fft-shifted
    float* srow = row + N;                                // [N] smoothed / clamped
    unsigned* hist = reinterpret_cast<unsigned*>(srow + N);   // [1024] bins (rare path: [4][256])
    static_assert(EPI == EPI_RAW || N >= 512, "epilogues need N >= 512");
    constexpr bool SM = EPI == EPI_SMOOTH;
    __shared__ unsigned us_s[SM ? C::FPC : 1][16];        // [2]=#cand [3]=min key above [4]=v1 [5]=v2 [6]=nan [8..10]=scan result
    __shared__ unsigned uf_s[SM ? C::FPC : 1][8];         // rare path: [0]=kmin [1]=kmax [2]=cnt_le [3]=min_gt
    __shared__ __align__(16) float cand_s[SM ? C::FPC : 1][SM ? C::CAP : 4];
    constexpr int FS = EPI == EPI_RAW ? 1 : C::FPC;
    __shared__ double dscr_s[FS][16];                     // per-warp sums
    __shared__ __align__(16) float fscr_s[FS][32];        // per-warp max / min (and the raw row's sum / sum of squares)
    __shared__ unsigned uscr_s[FS][16];                   // scanner partial counts
    double* dscr = dscr_s[EPI == EPI_RAW ? 0 : f];
    float* fscr = fscr_s[EPI == EPI_RAW ? 0 : f];
    unsigned* uscr = uscr_s[EPI == EPI_RAW ? 0 : f];
    if constexpr (SM) {
        if (t < 16) us_s[f][t] = (t == 0 || t == 3) ? 0xffffffffu : 0u;
        if (t < 8) uf_s[f][t] = (t == 0 || t == 3) ? 0xffffffffu : 0u;
    }
    float rsum = 0.f, rsq = 0.f;                          // EPI_SMOOTH: sum, sum of squares, NaN flag of this thread's raw dB values
    bool rnan = false;
    __shared__ double mom_s[C::THREADS / 32][3];

    pss_grid_dependency_sync();       // PSS_PDL: everything above touched only tables and this CTA's shared memory
    // ---- pass 0: global (coalesced 8-byte loads) * window -> radix-16 -> shared
    {
        cx<T> v[16];
        if constexpr (LOG2N1 > 0) {
            const cx<T>* src = reinterpret_cast<const cx<T>*>(p.ystage) + frame * N;
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r] = live ? src[t + r * TPF] : cx<T>{(T)0, (T)0};
        } else {
            const float2* src = p.iq + frame * N;
            const T* win = reinterpret_cast<const T*>(p.window);
            if (p.ahead > 0) {
                // pull the frames of the CTA that takes this SM slot next into L2 (CTAs start in index order;
                // measured -2 % on the smoothing variant, -4 % at 8192 points)
                const long long total = p.n_frames * N;
                const long long e0 = (long long)(blockIdx.x + p.ahead) * C::FPC * N + (long long)tid * 16;
                constexpr int LINES = C::FPC * N / 16;             // 16 float2 = one 128-byte line per thread
                for (int l = 0; l < LINES; l += C::THREADS)
                    if (e0 + (long long)l * 16 < total) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.iq + e0 + (long long)l * 16));
            }
            float mii = 0.f, mqq = 0.f, miq = 0.f;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int idx = t + r * TPF;
                const float2 s = live ? __ldg(src + idx) : make_float2(0.f, 0.f);
                if constexpr (EPI != EPI_SCAN && N >= 512) {
                    mii = fmaf(s.x, s.x, mii);
                    mqq = fmaf(s.y, s.y, mqq);
                    miq = fmaf(s.x, s.y, miq);
                }
                if (win) {
                    const T w = __ldg(win + idx);
                    v[r] = {(T)s.x * w, (T)s.y * w};
                } else {
                    v[r] = {(T)s.x, (T)s.y};
                }
            }
            if constexpr (EPI != EPI_SCAN && N >= 512) {
                if (p.moments) {          // frame moments: float partials per warp, fp64 across the frame
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        mii += __shfl_xor_sync(0xffffffffu, mii, o);
                        mqq += __shfl_xor_sync(0xffffffffu, mqq, o);
                        miq += __shfl_xor_sync(0xffffffffu, miq, o);
                    }
                    if ((tid & 31) == 0) {
                        mom_s[tid >> 5][0] = (double)mii;
                        mom_s[tid >> 5][1] = (double)mqq;
                        mom_s[tid >> 5][2] = (double)miq;
                    }
                }
            }
        }
        fft_regs<16, T>::run(v);
        const int base = t << 4;
#pragma unroll
        for (int q = 0; q < 16; ++q) buf[fft_swz(base + fft_perm<16>(q))] = v[q];
    }
    __syncthreads();
    if constexpr (EPI != EPI_SCAN && N >= 512 && LOG2N1 == 0) {
        if (p.moments && live && t == 0) {
            double a = 0.0, b = 0.0, c = 0.0;
            for (int w = 0; w < TPF / 32; ++w) {
                a += mom_s[f * (TPF / 32) + w][0];
                b += mom_s[f * (TPF / 32) + w][1];
                c += mom_s[f * (TPF / 32) + w][2];
            }
            double* m = p.moments + frame * 4;
            m[0] = a; m[1] = b; m[2] = c; m[3] = 0.0;
        }
    }

    // ---- |X|^2 -> dB at the fft-shifted position
    auto emit = [&](int k, const cx<T> X) {
        float d = 0.f;
        [[maybe_unused]] double pw = 0.0;
        if constexpr (EPI == EPI_SCAN) {
            pw = (double)X.x * (double)X.x + ((double)X.y * (double)X.y + 1e-10);
        } else if constexpr (sizeof(T) == 8) {
            d = db_from_power((double)X.x * (double)X.x + ((double)X.y * (double)X.y + 1e-10));
        } else {            // PSS_PREC_FP32 fast mode: the whole chain in float (does NOT meet 1e-4 dB)
            const float pf = (float)X.x * (float)X.x + ((float)X.y * (float)X.y + 1e-10f);
            float l;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(pf));
            d = 3.01029995663981195f * l;
        }
        const int pos = k ^ (N >> 1);
        if constexpr (LOG2N1 > 0) {
            const long long fb = frame >> LOG2N1;
            const int bin = (int)(frame & ((1 << LOG2N1) - 1)) + (k << LOG2N1);
            const int ntot = N << LOG2N1;
            if (live) p.db[fb * ntot + (bin ^ (ntot >> 1))] = d;
        } else if constexpr (EPI == EPI_RAW) {
            if (live) p.db[frame * N + pos] = d;
        } else if constexpr (EPI == EPI_SCAN) {
            reinterpret_cast<double*>(buf)[pos] = pw;        // fp64 power row (the exchange buffer is dead)
        } else {
            row[pos] = d;
            rsum += d;
            rsq = fmaf(d, d, rsq);
            rnan |= d != d;
        }
    };
    if constexpr (NP == 2) {
        stockham_pass<LOG2N, 1, T>(buf, wpre[0], t, emit);
    } else if constexpr (NP == 3) {
        stockham_pass<LOG2N, 1, T>(buf, wpre[0], t, [](int, cx<T>) {});
        stockham_pass<LOG2N, 2, T>(buf, wpre[1], t, emit);
    } else {
        stockham_pass<LOG2N, 1, T>(buf, wpre[0], t, [](int, cx<T>) {});
        stockham_pass<LOG2N, 2, T>(buf, wpre[1], t, [](int, cx<T>) {});
        stockham_pass<LOG2N, 3, T>(buf, wpre[2], t, emit);
    }

    if constexpr (EPI == EPI_SCAN) {
        // scanner: peak and number of bins above (peak - rel) or above an absolute threshold.  The integer
        // count is taken in the fp64 POWER domain (|X|^2 + 1e-10 against p_max * 10^(-rel/10) or 10^(thr/10)):
        // the comparison then differs from the reference's dB comparison only on ties at the 1e-15 level,
        // where a float32 dB row would flip bins within ~1e-5 dB of the threshold.
        const int wf = t >> 5, lane = t & 31, nw = TPF / 32;
        const double* prow = reinterpret_cast<const double*>(buf);
        __syncthreads();
        double vals[16];
        double mx = 0.0;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            vals[m] = prow[t + m * TPF];
            mx = fmax(mx, vals[m]);                      // fmax drops NaN like the float row maximum did
            if (live && p.db) p.db[frame * N + t + m * TPF] = db_from_power(vals[m]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) dscr[wf] = mx;
        __syncthreads();
        double pmax = dscr[0];
        for (int w = 1; w < nw; ++w) pmax = fmax(pmax, dscr[w]);
        const double thr = p.use_abs ? p.thr_pow : pmax * p.thr_pow;
        int cnt = 0;
#pragma unroll
        for (int m = 0; m < 16; ++m) cnt += vals[m] > thr;
        cnt = warp_sum(cnt);
        if (lane == 0) uscr[wf] = (unsigned)cnt;
        __syncthreads();
        if (t == 0 && live) {
            unsigned tot = 0;
            for (int w = 0; w < nw; ++w) tot += uscr[w];
            p.peak[frame] = (float)(10.0 * log10(pmax));
            p.count[frame] = (int)tot;
        }
    }

    if constexpr (EPI == EPI_SMOOTH) {
        smooth_epilogue<N, TPF>(row, srow, hist, us_s[f], uf_s[f], cand_s[f], dscr, fscr, rsum, rsq, rnan, t, live, frame, p);
    }
}


// The same epilogue as a kernel of its own, one CTA of N/16 threads per raw dB row (N = 4096, 8192).  The
// fused kernel keeps a 64 KB fp64 exchange buffer per frame, so only 2 of its CTAs fit an SM and the
// epilogue's barriers and shared atomics are latency the SM cannot hide; here a CTA needs 36 KB and
// ~64 registers, 4 are resident, and the rows come out of L2 (the caller runs transform and epilogue
// over L2-sized chunks of frames, so the raw rows never reach DRAM).
struct PsdEpiParams {
    const float* raw;     // [n_frames][N] raw dB rows
    long long n_frames;
    PsdParams out;        // db / cols / W / stats / col_step
};

template <int LOG2N>
__global__ void __launch_bounds__((1 << LOG2N) / 16, LOG2N == 12 ? 4 : 2)
psd_epilogue_kernel(const PsdEpiParams q) {
    constexpr int N = 1 << LOG2N, TPF = N / 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* row = reinterpret_cast<float*>(smem_raw);
    float* srow = row + N;
    unsigned* hist = reinterpret_cast<unsigned*>(srow + N);          // [1024]
    __shared__ unsigned us[16], uf[8];
    __shared__ __align__(16) float cand[64];
    __shared__ double dscr[16];
    __shared__ __align__(16) float fscr[32];
    const int t = threadIdx.x;
    const long long frame = blockIdx.x;
    if (t < 16) us[t] = (t == 0 || t == 3) ? 0xffffffffu : 0u;
    if (t < 8) uf[t] = (t == 0 || t == 3) ? 0xffffffffu : 0u;
    const float4* src = reinterpret_cast<const float4*>(q.raw + frame * N);
    float rsum = 0.f, rsq = 0.f;
    bool rnan = false;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const float4 v = __ldcs(src + t + g * TPF);
        reinterpret_cast<float4*>(row)[t + g * TPF] = v;
        rsum += (v.x + v.y) + (v.z + v.w);
        rsq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, rsq))));
        rnan |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
    }
    smooth_epilogue<N, TPF>(row, srow, hist, us, uf, cand, dscr, fscr, rsum, rsq, rnan, t, true, frame, q.out);
}

// ---------------------------------------------------------------------------------- large transforms
// N = N1 * N2 (N1 = 4, 8, 16; N2 = 4096 or 8192), four-step FFT:
//   stage A (this kernel): per column n2, window, N1-point DFT over the stride-N2 samples held in
//   registers, twiddle W_N^(n2*k1), fp64 rows Y[frame][k1][n2] to a scratch that is sized to stay in L2;
//   stage B: psd_kernel<LOG2N2, T, EPI_RAW, LOG2N1> on every row, bins k1 + N1*k2.
template <int LOG2N1, typename T>
__global__ void __launch_bounds__(256)
psd_colfft_kernel(const float2* __restrict__ iq, const T* __restrict__ window, const cx<T>* __restrict__ twN,
                  const int N2, const long long n_frames, cx<T>* __restrict__ Y) {
    constexpr int N1 = 1 << LOG2N1;
    const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long frame = gid / N2;
    const int n2 = (int)(gid - frame * N2);
    if (frame >= n_frames) return;
    const long long N = (long long)N1 * N2;
    const float2* src = iq + frame * N + n2;
    cx<T> v[N1];
#pragma unroll
    for (int n1 = 0; n1 < N1; ++n1) {
        const float2 s = __ldg(src + (long long)n1 * N2);
        if (window) {
            const T w = __ldg(window + (long long)n1 * N2 + n2);
            v[n1] = {(T)s.x * w, (T)s.y * w};
        } else {
            v[n1] = {(T)s.x, (T)s.y};
        }
    }
    fft_regs<N1, T>::run(v);
    cx<T> o[N1];
#pragma unroll
    for (int q = 0; q < N1; ++q) o[fft_perm<N1>(q)] = v[q];
    twiddle_apply<N1, T>(o, twN[n2]);                      // o[k1] *= W_N^(n2*k1)
    cx<T>* dst = Y + frame * N + n2;
#pragma unroll
    for (int k1 = 0; k1 < N1; ++k1) dst[(long long)k1 * N2] = o[k1];
}

// Column transforms of the largest reads (N = N1 * 4096 with N1 = 64, 128, 256: 2^18 ... 2^20 points, the
// app's SAMPLES = 10 ... 12, pyspecsdr.py:2236,2420-2422).  One CTA = 16 adjacent columns: window, N1-point
// radix-2 DIF over the stride-4096 samples in shared memory, twiddle W_N^(n2*k1) from a two-level table,
// fp64 rows Y[frame][k1][n2] for psd_kernel<12, double, EPI_RAW, LOG2N1>.  Availability path, not tuned.
template <int LOG2N1>
__global__ void __launch_bounds__(256)
psd_colfft_big_kernel(const float2* __restrict__ iq, const double* __restrict__ window, const double2* __restrict__ tw1,
                      const double2* __restrict__ thi, const double2* __restrict__ tlo, const long long n_frames,
                      cx<double>* __restrict__ Y) {
    constexpr int N1 = 1 << LOG2N1, N2 = 4096, COLS = 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<double>* buf = reinterpret_cast<cx<double>*>(smem_raw);           // [N1][COLS]
    __shared__ double2 w1[N1 / 2];                                        // W_N1^k
    const int tid = threadIdx.x;
    const long long frame = blockIdx.x / (N2 / COLS);
    const int c0 = (int)(blockIdx.x % (N2 / COLS)) * COLS;
    if (frame >= n_frames) return;
    const long long N = (long long)N1 * N2;
    for (int k = tid; k < N1 / 2; k += 256) w1[k] = tw1[k];
    const float2* src = iq + frame * N + c0;
    for (int idx = tid; idx < N1 * COLS; idx += 256) {
        const int n1 = idx / COLS, col = idx % COLS;
        const float2 v = __ldg(src + (long long)n1 * N2 + col);
        const double w = window ? __ldg(window + (long long)n1 * N2 + c0 + col) : 1.0;
        buf[idx] = {(double)v.x * w, (double)v.y * w};
    }
    __syncthreads();
#pragma unroll 1
    for (int half = N1 / 2; half >= 1; half >>= 1) {
        const int tstep = (N1 / 2) / half;
        for (int b = tid; b < (N1 / 2) * COLS; b += 256) {
            const int j = b / COLS, col = b % COLS;
            const int pos = j % half, i0 = (j / half) * 2 * half + pos, i1 = i0 + half;
            const cx<double> a = buf[i0 * COLS + col], c = buf[i1 * COLS + col];
            const double2 w = w1[pos * tstep];
            buf[i0 * COLS + col] = cadd(a, c);
            buf[i1 * COLS + col] = cmul(csub(a, c), cx<double>{w.x, w.y});
        }
        __syncthreads();
    }
    cx<double>* dst = Y + frame * N + c0;
    for (int idx = tid; idx < N1 * COLS; idx += 256) {
        const int pos = idx / COLS, col = idx % COLS;
        const int k1 = (int)(__brev((unsigned)pos) >> (32 - LOG2N1));     // DIF leaves bit-reversed order
        const unsigned e = (unsigned)(c0 + col) * (unsigned)k1;           // < N
        const double2 a = __ldg(thi + (e >> 10)), b = __ldg(tlo + (e & 1023u));
        const cx<double> w = cmul(cx<double>{a.x, a.y}, cx<double>{b.x, b.y});
        dst[(long long)k1 * N2 + col] = cmul(buf[idx], w);
    }
}

// Row epilogue for rows that do not fit one CTA's shared memory: the same 5-bin smoothing, exact
// median clamp, statistics and W-column resample as EPI_SMOOTH, streaming the row from L2.
__device__ __forceinline__ bool row_median1_512(const float* __restrict__ srow, const int n, const float* pst,
                                                unsigned* hist, unsigned* us, float* cand, float& v1, float& v2);
__device__ __forceinline__ void row_stats_512(float rsum, float rsq, float* pst);

__global__ void __launch_bounds__(512, 2)
row_epilogue_kernel(const float* __restrict__ raw, const int N, const long long n_frames, float* __restrict__ db,
                    float* __restrict__ cols, const int W, float* __restrict__ stats) {
    __shared__ unsigned hist[1024];
    __shared__ unsigned us[16];
    __shared__ __align__(16) float cand[64];
    __shared__ __align__(16) float pst[32];
    __shared__ double dsum[16];
    __shared__ float fmx[16], fmn[16];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = N - 4;
    for (long long f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const float* d = raw + f * N;
        float* s = db + f * n;
        if (tid < 16) us[tid] = tid == 3 ? 0xffffffffu : 0u;
        __syncthreads();
        // 5-bin means; their mean and spread place the buckets of the median's histogram
        float rsum = 0.f, rsq = 0.f;
        bool has_nan = false;
        for (int i = tid; i < n; i += 512) {
            const float v = ((d[i] + d[i + 1]) + (d[i + 2] + d[i + 3]) + d[i + 4]) * 0.2f;
            s[i] = v;
            rsum += v;
            rsq = fmaf(v, v, rsq);
            has_nan |= (v != v);
        }
        row_stats_512(rsum, rsq, pst);
        const bool any_nan = __syncthreads_or(has_nan);
        float v1, v2;
        if (!row_median1_512(s, n, pst, hist, us, cand, v1, v2)) {
            unsigned ka, kb;
            row_select2_512(s, n, (unsigned)((n - 1) / 2), hist, us, ka, kb);
            v1 = key2f(ka);
            v2 = key2f((n & 1) ? ka : kb);
        }
        float thr = (float)(0.5 * ((double)v1 + (double)v2) - 10.0);
        if (any_nan) thr = __int_as_float(0x7fc00000);
        float mx = -INFINITY, mn = INFINITY;
        double sm = 0.0;
        for (int i = tid; i < n; i += 512) {
            float v = s[i];
            if (v < thr) v = thr;
            s[i] = v;
            mx = fmaxf(mx, v);
            mn = fminf(mn, v);
            sm += (double)v;
        }
        mx = warp_max(mx);
        mn = warp_min(mn);
        sm = warp_sum(sm);
        if (lane == 0) {
            fmx[warp] = mx;
            fmn[warp] = mn;
            dsum[warp] = sm;
        }
        __syncthreads();                       // also orders the clamped row for the resample below
        if (stats && tid == 0) {
            float a = fmx[0], b = fmn[0];
            double t = dsum[0];
            for (int w = 1; w < 16; ++w) {
                a = fmaxf(a, fmx[w]);
                b = fminf(b, fmn[w]);
                t += dsum[w];
            }
            const float nanv = __int_as_float(0x7fc00000);
            float4 st;
            st.x = any_nan ? nanv : a;
            st.y = any_nan ? nanv : (float)(t / n);
            st.z = b;
            st.w = a;
            reinterpret_cast<float4*>(stats)[f] = st;
        }
        if (cols) {
            const double step = W > 1 ? (double)(n - 1) / (double)(W - 1) : 0.0;
            for (int c = tid; c < W; c += 512) {
                const double x = (c == W - 1 && W > 1) ? (double)(n - 1) : c * step;
                const int j = (int)x;
                float o;
                if (j >= n - 1) o = s[n - 1];
                else {
                    const double y0 = s[j], y1 = s[j + 1];
                    o = (float)((y1 - y0) * (x - (double)j) + y0);
                }
                cols[f * W + c] = o;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------- fused large transforms
// N = N1 * 4096 (N1 = 4, 8, 16: 16384 ... 65536 points; the app's default read is 32768 samples,
// pyspecsdr.py:105,2236) in ONE persistent launch, one frame per CTA at a time (512 threads, 1 CTA/SM):
//   stage A: per column n2, window, N1-point DFT over the stride-4096 samples in registers, twiddle
//            W_N^(n2*k1), fp64 rows Y[k1][n2] to this CTA's private scratch (reused every frame, so it
//            lives in L2; HBM sees the 8-byte samples once and the 4-byte dB once);
//   stage B: the two 256-thread halves of the CTA each run 4096-point row transforms (radix-16
//            Stockham through their own 64 KB of shared memory, named barriers, out of phase like
//            two CTAs) and scatter bins k1 + N1*k2 as dB;
//   epilogue (EPI_SMOOTH): the raw dB row comes back from L2, the smoothed row is kept in the
//            128 KB of shared memory the transforms no longer need (N <= 32768) and the exact
//            median / clamp / statistics / resample run on it.
// Per-warp partials of a row's sum and sum of squares (pst[0..15], pst[16..31]); the caller's next CTA
// barrier publishes them.
__device__ __forceinline__ void row_stats_512(float rsum, float rsq, float* pst) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
        rsq += __shfl_xor_sync(0xffffffffu, rsq, o);
    }
    if ((threadIdx.x & 31) == 0) {
        pst[threadIdx.x >> 5] = rsum;
        pst[16 + (threadIdx.x >> 5)] = rsq;
    }
}

// Exact lower/upper median of a row of n floats (shared or L2-resident), all 512 threads of the CTA.  Every
// value maps monotonically to a 20-bit key over mean +- 1.1 sigma of the row (pst: the per-warp partials
// above, taken over the smoothed values themselves, so the median is inside; values outside land in the end
// keys).  One 1024-bin histogram of the upper 10 bits, scanned by one warp; if the selected bucket holds more
// than 64 elements (rows beyond ~64 k bins) a second histogram of the lower 10 bits inside it; then the few
// elements left are ranked directly.  Returns false (CTA-uniform) when even that bucket holds more than 64
// elements (flat rows); the caller then runs the key radix select.
// hist [1024], us [16] ([2] = 0, [3] = 0xffffffff on entry), cand [64] are shared scratch.
__device__ __forceinline__ bool row_median1_512(const float* __restrict__ srow, const int n, const float* pst,
                                                unsigned* hist, unsigned* us, float* cand, float& v1, float& v2) {
    const int tid = threadIdx.x, lane = tid & 31;
    float S = 0.f, Q = 0.f;
#pragma unroll
    for (int w = 0; w < 16; w += 4) {
        const float4 a = *reinterpret_cast<const float4*>(pst + w);
        const float4 b = *reinterpret_cast<const float4*>(pst + 16 + w);
        S += (a.x + a.y) + (a.z + a.w);
        Q += (b.x + b.y) + (b.z + b.w);
    }
    const float inv = 1.0f / (float)n;
    const float mean = S * inv, var = Q * inv - mean * mean;
    const float scale = var > 0.f ? (1048576 / 2.2f) * rsqrtf(var) : 0.f;
    const float off = var > 0.f ? fmaf(-mean, scale, 524288.0f) : 0.f;
    auto key = [&](const float v) { return min(1048575u, __float2uint_rz(fmaf(v, scale, off))); };
    hist[tid] = 0u;
    hist[tid + 512] = 0u;
    if (tid < 64) cand[tid] = INFINITY;
    __syncthreads();
    // (measured: counting the out-of-range quarter of a Gaussian row in registers instead of in the two end
    // buckets is slower -- the branches cost more than the contended atomics)
    for (int i0 = 4 * tid; i0 < n; i0 += 2048) {
        const float4 v = *reinterpret_cast<const float4*>(srow + i0);
        atomicAdd(&hist[key(v.x) >> 10], 1u);
        atomicAdd(&hist[key(v.y) >> 10], 1u);
        atomicAdd(&hist[key(v.z) >> 10], 1u);
        atomicAdd(&hist[key(v.w) >> 10], 1u);
    }
    __syncthreads();
    auto pick = [&](unsigned r_in) {                         // one warp scans, everyone reads (rank, bin, count)
        if (tid < 32) {
            unsigned r = r_in, d, c;
            hist_pick_wide<1024>(hist, r, d, c, lane);
            if (lane == 0) {
                us[8] = r;
                us[9] = d;
                us[10] = c;
            }
        }
        __syncthreads();
    };
    pick((unsigned)((n - 1) / 2));
    unsigned rank = us[8], sel = us[9], m = us[10];
    int shift = 10;
    if (m > 64u) {                                           // CTA-uniform
        __syncthreads();                                     // everyone has read us[8..10]
        hist[tid] = 0u;
        hist[tid + 512] = 0u;
        __syncthreads();
        for (int i0 = 4 * tid; i0 < n; i0 += 2048) {
            const float4 v = *reinterpret_cast<const float4*>(srow + i0);
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned k = key(e[q]);
                if ((k >> 10) == sel) atomicAdd(&hist[k & 1023u], 1u);
            }
        }
        __syncthreads();
        pick(rank);
        rank = us[8];
        sel = (sel << 10) | us[9];
        m = us[10];
        shift = 0;
        if (m > 64u) return false;
    }
    unsigned kgt = 0xffffffffu;
    for (int i0 = 4 * tid; i0 < n; i0 += 2048) {
        const float4 v = *reinterpret_cast<const float4*>(srow + i0);
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const unsigned b = key(e[q]) >> shift;
            if (b == sel) {
                const unsigned slot = atomicAdd(&us[2], 1u);
                if (slot < 64u) cand[slot] = e[q];
            }
            if (b > sel) kgt = min(kgt, f2key(e[q]));
        }
    }
    kgt = __reduce_min_sync(0xffffffffu, kgt);
    if (lane == 0 && kgt != 0xffffffffu) atomicMin(&us[3], kgt);
    __syncthreads();
    if ((unsigned)tid < m) {
        const float c = cand[tid];
        unsigned rk = 0;
        for (unsigned j = 0; j < m; j += 4) {                // slots >= m hold +inf: they never count
            const float4 o = *reinterpret_cast<const float4*>(cand + j);
            rk += (o.x < c) || (o.x == c && j < (unsigned)tid);
            rk += (o.y < c) || (o.y == c && j + 1 < (unsigned)tid);
            rk += (o.z < c) || (o.z == c && j + 2 < (unsigned)tid);
            rk += (o.w < c) || (o.w == c && j + 3 < (unsigned)tid);
        }
        if (rk == rank) us[4] = __float_as_uint(c);
        if (rk == rank + 1u) us[5] = __float_as_uint(c);
    }
    __syncthreads();
    v1 = __uint_as_float(us[4]);
    v2 = (n & 1) ? v1 : (rank + 1u < m ? __uint_as_float(us[5]) : key2f(us[3]));
    return true;
}

// The same as a real call: the 32768-point fused kernel sits at its 128-register limit and spills in the
// transform loops when the median is inlined into it (measured 1.98 inlined vs 1.80 ms per GiB called; the
// other sizes are faster inlined).
__device__ __noinline__ bool row_median1_512_call(const float* __restrict__ srow, const int n, const float* pst,
                                                  unsigned* hist, unsigned* us, float* cand, float& v1, float& v2) {
    return row_median1_512(srow, n, pst, hist, us, cand, v1, v2);
}

// L2 residency control for the fused large transforms: the per-CTA scratch (fp64 rows, raw dB row) is
// written and read back within one frame period and must survive the 1 GB/ms input stream flowing
// through the same L2, so scratch accesses carry an evict_last policy and the stream is read evict-first.
__device__ __forceinline__ unsigned long long l2_evict_last_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st_keep(double2* p, const double x, const double y, const unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(x), "d"(y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_keep(float* p, const float v, const unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ double2 ld_keep(const double2* p, const unsigned long long pol) {
    double2 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                 : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_keep(const float4* p, const unsigned long long pol) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol) : "memory");
    return v;
}

struct PsdLargeParams {
    const float2* iq;
    const double* window;          // [N] or nullptr
    const cx<double>* twN;         // [4096]: W_N^n2
    const cx<double>* tw;          // row-transform base twiddles (4096-point tables)
    long long n_frames;
    float* db;
    float* cols;
    int W;
    float* stats;
    cx<double>* Y;                 // [grid][N] scratch
    float* rawdb;                  // [grid][N] scratch (EPI_SMOOTH)
    float* srow2;                  // [grid][N] scratch (EPI_SMOOTH, rows that do not fit shared memory)
    double* moments;               // [n_frames][4] sum I^2, Q^2, IQ per frame (by-product for the WFM demod), or null
};

// Cluster helpers (thread-block clusters, sm_90+): rank of this CTA in its cluster and a full cluster barrier
// whose release / acquire makes the global-memory rows written before it visible to the other CTAs.
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// CL = CTAs per frame.  One CTA per frame (CL = 1) keeps grid x N x 16 bytes of fp64 rows alive: 38 MB at 16384
// points, 76 MB at 32768 - more than the L2 holds beside the input stream, so half of the scratch went to DRAM
// and back (ncu, round 1: 2.5x the algorithmic traffic).  A cluster of CL CTAs on CL SMs shares ONE frame: every
// CTA transforms 1/CL of the columns and, after a cluster barrier, 1/CL of the rows, so only grid / CL frames
// are in flight (38 MB again at 32768 with CL = 2, at 65536 with CL = 4) and the rows stay in L2.  With the
// epilogue the cluster takes CL frames at a time and every CTA finishes one of them.
template <int LOG2N, int EPI, int CL>
__global__ void __launch_bounds__(512, 1) psd_large_kernel(const PsdLargeParams p) {
    constexpr int N = 1 << LOG2N, LOG2N1 = LOG2N - 12, N1 = 1 << LOG2N1, N2 = 4096, n = N - 4;
    constexpr bool SROW_SMEM = (size_t)N * 4 <= 2 * fft_padded(N2) * sizeof(cx<double>);
    static_assert((N2 / 512) % CL == 0 && (N1 / 2) % CL == 0, "columns and row pairs must split evenly over the cluster");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned hist[1024];
    __shared__ unsigned us[16];
    __shared__ __align__(16) float cand[64];
    __shared__ __align__(16) float pst[32];
    __shared__ double dsum_s[16];
    __shared__ float fmx_s[16], fmn_s[16];
    __shared__ double mom_s[16][3];
    const int tid = threadIdx.x, g = tid >> 8, t = tid & 255, lane = tid & 31, warp = tid >> 5;
    const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
    const long long cid = blockIdx.x / CL, ncl = gridDim.x / CL;
    cx<double>* buf = reinterpret_cast<cx<double>*>(smem_raw) + (size_t)g * fft_padded(N2);
    const HalfBarrier hbar{1 + g};
    cx<double> wpre[2];
    wpre[0] = p.tw[t & 15];                       // pass 1: ns = 16, offset 0
    wpre[1] = p.tw[16 + t];                       // pass 2: ns = 256, offset (256-16)/15 = 16
    cx<double>* Y = p.Y + (size_t)cid * N;
    float* rawbase = EPI == EPI_SMOOTH ? p.rawdb + (size_t)cid * CL * N : nullptr;
    const unsigned long long keep = l2_evict_last_policy();

    // with the epilogue the cluster works on groups of CL consecutive frames, otherwise frame by frame
    constexpr int FPG = EPI == EPI_SMOOTH ? CL : 1;
    for (long long grp = cid; grp * FPG < p.n_frames; grp += ncl) {
#pragma unroll 1
      for (int j = 0; j < FPG; ++j) {
        const long long frame = FPG == 1 ? grp : grp * FPG + j;
        if (FPG > 1 && frame >= p.n_frames) {     // ragged last group: keep the barrier count uniform
            cluster_barrier();
            cluster_barrier();
            continue;
        }
        float* raw = EPI == EPI_SMOOTH ? (FPG == 1 ? rawbase : rawbase + (size_t)j * N) : nullptr;
        if constexpr (EPI == EPI_SMOOTH && CL == 1) {    // one CTA per frame: reset the median scratch ahead of the barriers below
            if (tid < 16) us[tid] = tid == 3 ? 0xffffffffu : 0u;
        }
        // ---- stage A: column DFTs of this CTA's share of the columns
        const float2* src = p.iq + frame * N;
        float mii = 0.f, mqq = 0.f, miq = 0.f;
#pragma unroll(N1 == 4 ? 4 : N1 == 8 ? 2 : 1)
        for (int c = rank * (N2 / 512 / CL); c < (rank + 1) * (N2 / 512 / CL); ++c) {
            const int n2 = tid + 512 * c;
            cx<double> v[N1];
#pragma unroll
            for (int n1 = 0; n1 < N1; ++n1) {
                const float2 s = __ldcs(src + n2 + N2 * n1);
                if constexpr (EPI == EPI_SMOOTH) {
                    mii = fmaf(s.x, s.x, mii);
                    mqq = fmaf(s.y, s.y, mqq);
                    miq = fmaf(s.x, s.y, miq);
                }
                if (p.window) {
                    const double w = __ldg(p.window + n2 + N2 * n1);
                    v[n1] = {(double)s.x * w, (double)s.y * w};
                } else {
                    v[n1] = {(double)s.x, (double)s.y};
                }
            }
            fft_regs<N1, double>::run(v);
            cx<double> o[N1];
#pragma unroll
            for (int q = 0; q < N1; ++q) o[fft_perm<N1>(q)] = v[q];
            const double2 wn = __ldg(reinterpret_cast<const double2*>(p.twN + n2));
            twiddle_apply<N1, double>(o, cx<double>{wn.x, wn.y});     // o[k1] *= W_N^(n2*k1)
#pragma unroll
            for (int k1 = 0; k1 < N1; ++k1)
                st_keep(reinterpret_cast<double2*>(Y + k1 * N2 + n2), o[k1].x, o[k1].y, keep);
        }
        if constexpr (EPI == EPI_SMOOTH) {
            if (p.moments) {              // float partials per thread / warp, fp64 across the frame
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mii += __shfl_xor_sync(0xffffffffu, mii, o);
                    mqq += __shfl_xor_sync(0xffffffffu, mqq, o);
                    miq += __shfl_xor_sync(0xffffffffu, miq, o);
                }
                if (lane == 0) {
                    mom_s[warp][0] = (double)mii;
                    mom_s[warp][1] = (double)mqq;
                    mom_s[warp][2] = (double)miq;
                }
            }
        }
        __syncthreads();
        if constexpr (EPI == EPI_SMOOTH) {
            if (p.moments && tid == 0) {
                double a = 0.0, b = 0.0, c = 0.0;
                for (int w = 0; w < 16; ++w) {
                    a += mom_s[w][0];
                    b += mom_s[w][1];
                    c += mom_s[w][2];
                }
                double* m = p.moments + frame * 4;
                if (CL == 1) {
                    m[0] = a; m[1] = b; m[2] = c; m[3] = 0.0;
                } else {                      // every CTA of the cluster adds its columns' share (zeroed by the host)
                    atomicAdd(m, a);
                    atomicAdd(m + 1, b);
                    atomicAdd(m + 2, c);
                }
            }
        }
        if constexpr (LOG2N == 14) {
            // pull the next frame of this CTA into L2 while the row transforms keep the fp64 pipe busy
            // (measured: -4 % at 16384 points; at 32768+ the scratch already fills the L2 and it hurts)
            const long long nf = frame + gridDim.x;                       // (CL == 1 at 16384 points)
            if (nf < p.n_frames) {
                const char* nx = reinterpret_cast<const char*>(p.iq + nf * N);
                for (int l = tid; l < N * 8 / 128; l += 512) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (size_t)l * 128));
            }
        }
        if (CL > 1) cluster_barrier();            // every column of the frame is in the scratch
        // ---- stage B: row transforms, two rows at a time
        for (int pair = rank * (N1 / 2 / CL); pair < (rank + 1) * (N1 / 2 / CL); ++pair) {
            const int k1 = 2 * pair + g;
            {
                const double2* rowY = reinterpret_cast<const double2*>(Y + k1 * N2);
                cx<double> v[16];
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const double2 y = ld_keep(rowY + t + r * 256, keep);
                    v[r] = {y.x, y.y};
                }
                fft_regs<16, double>::run(v);
                const int base = t << 4;
#pragma unroll
                for (int q = 0; q < 16; ++q) buf[fft_pad(base) + fft_perm<16>(q)] = v[q];      // base = 16 t
            }
            hbar();
            auto emit = [&](int k, const cx<double> X) {
                const float d = db_from_power(X.x * X.x + (X.y * X.y + 1e-10));
                const int pos = (k1 + (k << LOG2N1)) ^ (N >> 1);
                if constexpr (EPI == EPI_RAW) p.db[frame * N + pos] = d;
                else st_keep(raw + pos, d, keep);
            };
            stockham_pass<12, 1, double, true>(buf, &wpre[0], t, [](int, cx<double>) {}, hbar);
            stockham_pass<12, 2, double, true>(buf, &wpre[1], t, emit, hbar);
        }
        if (CL > 1) cluster_barrier();            // every row is transformed: the scratch may be overwritten
        else __syncthreads();
      }
      // ---- epilogue: CTA `rank` finishes frame grp * CL + rank
      if constexpr (EPI == EPI_SMOOTH) {
        const long long frame = CL == 1 ? grp : grp * CL + rank;
        if (CL == 1 || frame < p.n_frames) {
            float* raw = CL == 1 ? rawbase : rawbase + (size_t)rank * N;
            if (CL > 1) {
                if (tid < 16) us[tid] = tid == 3 ? 0xffffffffu : 0u;
                __syncthreads();
            }
        {
            float* srow = SROW_SMEM ? reinterpret_cast<float*>(smem_raw) : p.srow2 + (size_t)blockIdx.x * N;
            bool has_nan = false;
            float rsum = 0.f, rsq = 0.f;
            for (int i0 = 4 * tid; i0 < n; i0 += 2048) {
                const float4 a = ld_keep(reinterpret_cast<const float4*>(raw + i0), keep);
                const float4 c = ld_keep(reinterpret_cast<const float4*>(raw + i0 + 4), keep);     // i0 + 4 <= N - 4
                const float d[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
                float sv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    sv[e] = ((d[e] + d[e + 1]) + (d[e + 2] + d[e + 3]) + d[e + 4]) * 0.2f;
                    has_nan |= sv[e] != sv[e];
                    rsum += sv[e];
                    rsq = fmaf(sv[e], sv[e], rsq);
                }
                *reinterpret_cast<float4*>(srow + i0) = make_float4(sv[0], sv[1], sv[2], sv[3]);
            }
            row_stats_512(rsum, rsq, pst);
            const bool any_nan = __syncthreads_or(has_nan);
            float v1, v2;
            if (!(LOG2N == 15 ? row_median1_512_call(srow, n, pst, hist, us, cand, v1, v2)
                              : row_median1_512(srow, n, pst, hist, us, cand, v1, v2))) {
                unsigned ka, kb;
                row_select2_512(srow, n, (unsigned)((n - 1) / 2), hist, us, ka, kb);
                v1 = key2f(ka);
                v2 = key2f((n & 1) ? ka : kb);
            }
            float thr = (float)(0.5 * ((double)v1 + (double)v2) - 10.0);
            if (any_nan) thr = __int_as_float(0x7fc00000);
            float mx = -INFINITY, mn = INFINITY;
            double sm = 0.0;
            float* dst = p.db ? p.db + frame * n : nullptr;
            for (int i0 = 4 * tid; i0 < n; i0 += 2048) {
                float4 v = *reinterpret_cast<const float4*>(srow + i0);
                v.x = v.x < thr ? thr : v.x;
                v.y = v.y < thr ? thr : v.y;
                v.z = v.z < thr ? thr : v.z;
                v.w = v.w < thr ? thr : v.w;
                mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
                mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
                sm += (double)((v.x + v.y) + (v.z + v.w));
                *reinterpret_cast<float4*>(srow + i0) = v;
                if (dst) __stcs(reinterpret_cast<float4*>(dst + i0), v);
            }
            mx = warp_max(mx);
            mn = warp_min(mn);
            sm = warp_sum(sm);
            if (lane == 0) {
                fmx_s[warp] = mx;
                fmn_s[warp] = mn;
                dsum_s[warp] = sm;
            }
            __syncthreads();
            if (p.stats && tid == 0) {
                float a = fmx_s[0], b = fmn_s[0];
                double tt = dsum_s[0];
                for (int w = 1; w < 16; ++w) {
                    a = fmaxf(a, fmx_s[w]);
                    b = fminf(b, fmn_s[w]);
                    tt += dsum_s[w];
                }
                const float nanv = __int_as_float(0x7fc00000);
                float4 st;
                st.x = any_nan ? nanv : a;
                st.y = any_nan ? nanv : (float)(tt / n);
                st.z = b;
                st.w = a;
                reinterpret_cast<float4*>(p.stats)[frame] = st;
            }
            if (p.cols) {
                const int W = p.W;
                const double step = W > 1 ? (double)(n - 1) / (double)(W - 1) : 0.0;
                for (int c = tid; c < W; c += 512) {
                    const double x = (c == W - 1 && W > 1) ? (double)(n - 1) : c * step;
                    const int j = (int)x;
                    float o;
                    if (j >= n - 1) {
                        o = srow[n - 1];
                    } else {
                        const double y0 = srow[j], y1 = srow[j + 1];
                        o = (float)((y1 - y0) * (x - (double)j) + y0);
                    }
                    p.cols[frame * W + c] = o;
                }
            }
            __syncthreads();          // srow (shared) is the next frame's exchange buffer
        }
        }
        if (CL > 1) cluster_barrier();            // the raw rows of this group are consumed
      }
    }
}

// ---------------------------------------------------------------------------------- host side
template <typename T>
static int build_tables(pss_ctx* ctx, int log2n, pss_fft_tables& tab) {
    const int N = 1 << log2n;
    const int np = pss_num_passes(log2n);
    std::vector<cx<T>> tw;
    for (int ps = 1; ps < np; ++ps) {
        const long ns = 1L << (4 * ps);
        const long r = 1L << pss_pass_bits(log2n, ps);
        for (long k = 0; k < ns; ++k) {
            const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k /
                                  (long double)(ns * r);
            tw.push_back({(T)cosl(a), (T)sinl(a)});
        }
    }
    if (tw.empty()) tw.push_back({(T)1, (T)0});
    PSS_CUDA(ctx, cudaMalloc(&tab.twiddle, tw.size() * sizeof(cx<T>)));
    PSS_CUDA(ctx, cudaMemcpy(tab.twiddle, tw.data(), tw.size() * sizeof(cx<T>), cudaMemcpyHostToDevice));
    std::vector<T> w(N);
    for (int kind = PSS_WINDOW_HAMMING; kind <= PSS_WINDOW_HANN; ++kind) {
        // np.hamming / np.hanning: a - b*cos(2*pi*n/(N-1)), symmetric (signal_processing.py:246)
        const long double a = kind == PSS_WINDOW_HAMMING ? 0.54L : 0.5L;
        const long double b = kind == PSS_WINDOW_HAMMING ? 0.46L : 0.5L;
        for (int i = 0; i < N; ++i)
            w[i] = (T)(a - b * cosl(2.0L * 3.14159265358979323846264338327950288L * (long double)i /
                                    (long double)(N - 1)));
        PSS_CUDA(ctx, cudaMalloc(&tab.window[kind], N * sizeof(T)));
        PSS_CUDA(ctx, cudaMemcpy(tab.window[kind], w.data(), N * sizeof(T), cudaMemcpyHostToDevice));
    }
    return PSS_OK;
}

template <int LOG2N, typename T, int EPI>
static int launch_one(pss_ctx* ctx, const PsdParams& p) {
    using C = PsdCfg<LOG2N, T>;
    auto kern = psd_kernel<LOG2N, T, EPI>;
    if (ctx->configured.insert((const void*)kern).second)      // once per context (the attribute is per device)
        PSS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    const long long grid = (p.n_frames + C::FPC - 1) / C::FPC;
    PsdParams q = p;
    q.ahead = ctx->sm_count * ((EPI == EPI_SMOOTH && C::MINB == 2) ? pss_smooth_minb(LOG2N) : C::MINB);
    {   // PSS_PSD_AHEAD=<percent of the resident CTAs> (tuning): 0 switches the L2 prefetch-ahead off
        static const int pct = getenv("PSS_PSD_AHEAD") ? atoi(getenv("PSS_PSD_AHEAD")) : 100;
        q.ahead = (int)((long long)q.ahead * pct / 100);
    }
    if (q.W > 1) q.col_step = (double)(C::N - 4 - 1) / (double)(q.W - 1);
    PSS_CUDA(ctx, pss_launch(kern, (unsigned)grid, (unsigned)C::THREADS, (size_t)C::SMEM, ctx->stream, q));
    PSS_LAUNCH_CHECK(ctx);
    return PSS_OK;
}

template <typename T, int EPI>
static int launch_by_n(pss_ctx* ctx, int log2n, const PsdParams& p) {
    switch (log2n) {
        case 9: return launch_one<9, T, EPI>(ctx, p);
        case 10: return launch_one<10, T, EPI>(ctx, p);
        case 11: return launch_one<11, T, EPI>(ctx, p);
        case 12: return launch_one<12, T, EPI>(ctx, p);
        case 13: return launch_one<13, T, EPI>(ctx, p);
        default: break;
    }
    if constexpr (EPI == EPI_RAW) {
        switch (log2n) {
            case 6: return launch_one<6, T, EPI>(ctx, p);
            case 7: return launch_one<7, T, EPI>(ctx, p);
            case 8: return launch_one<8, T, EPI>(ctx, p);
            default: break;
        }
    }
    return PSS_ERR_UNSUPPORTED;
}

static int ilog2_exact(int n) {
    if (n <= 0 || (n & (n - 1))) return -1;
    int l = 0;
    while ((1 << l) < n) ++l;
    return l;
}

static int get_tables(pss_ctx* ctx, int log2n, pss_fft_tables** out, bool fp32 = false) {
    const int key = log2n + (fp32 ? 1000 : 0);
    auto it = ctx->fft_tables.find(key);
    if (it == ctx->fft_tables.end()) {
        pss_fft_tables tab;
        int rc = fp32 ? build_tables<float>(ctx, log2n, tab) : build_tables<double>(ctx, log2n, tab);
        if (rc != PSS_OK) return rc;
        it = ctx->fft_tables.emplace(key, tab).first;
    }
    *out = &it->second;
    return PSS_OK;
}


// Large transforms: N = 2^log2n with 14 <= log2n <= 20; their tables (pss_large_tables) live in the context.
typedef pss_large_tables LargeTables;

template <int LOG2N2, int LOG2N1>
static int launch_stage_b(pss_ctx* ctx, const PsdParams& p) {
    using C = PsdCfg<LOG2N2, double>;
    auto kern = psd_kernel<LOG2N2, double, EPI_RAW, LOG2N1>;
    PSS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    const long long grid = (p.n_frames + C::FPC - 1) / C::FPC;
    kern<<<(unsigned)grid, C::THREADS, C::SMEM, ctx->stream>>>(p);
    PSS_LAUNCH_CHECK(ctx);
    return PSS_OK;
}


template <int LOG2N, int EPI, int CL>
static int launch_large_fused(pss_ctx* ctx, const PsdLargeParams& p, unsigned grid) {
    auto kern = psd_large_kernel<LOG2N, EPI, CL>;
    constexpr int SMEM = 2 * fft_padded(4096) * (int)sizeof(cx<double>);
    PSS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    if (CL == 1) {
        kern<<<grid, 512, SMEM, ctx->stream>>>(p);
    } else {
        // thread-block cluster of CL CTAs per frame (cudaLaunchKernelEx, cluster dimension attribute)
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid, 1, 1);
        cfg.blockDim = dim3(512, 1, 1);
        cfg.dynamicSmemBytes = SMEM;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        PSS_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, p));
    }
    PSS_LAUNCH_CHECK(ctx);
    return PSS_OK;
}

template <int LOG2N, int CL>
static int launch_large_epi(pss_ctx* ctx, const PsdLargeParams& p, unsigned grid, bool smooth) {
    return smooth ? launch_large_fused<LOG2N, EPI_SMOOTH, CL>(ctx, p, grid) : launch_large_fused<LOG2N, EPI_RAW, CL>(ctx, p, grid);
}

static int psd_large(pss_ctx* ctx, const float* iq, int log2n, int64_t n_frames, int window, int epilogue,
                     const pss_psd_out* out) {
    const int log2n2 = log2n == 17 ? 13 : 12;
    const int log2n1 = log2n - log2n2;
    if (log2n1 < 2 || log2n1 > 8 || log2n1 == 5) return PSS_ERR_UNSUPPORTED;
    const bool big = log2n1 >= 6;          // 2^18 ... 2^20 points: shared-memory column transforms
    const long long N = 1LL << log2n, N2 = 1LL << log2n2;
    int rc;
    pss_fft_tables* tab2;
    if ((rc = get_tables(ctx, log2n2, &tab2))) return rc;
    LargeTables& lt = ctx->large_tables[log2n];
    if (!lt.ready) {
        std::vector<cx<double>> tw(N2);
        std::vector<double> w(N);
        const long double PI = 3.14159265358979323846264338327950288L;
        for (long long k = 0; k < N2; ++k) {
            const long double a = -2.0L * PI * (long double)k / (long double)N;
            tw[k] = {(double)cosl(a), (double)sinl(a)};
        }
        if (!lt.twN) PSS_CUDA(ctx, cudaMalloc(&lt.twN, N2 * sizeof(cx<double>)));
        PSS_CUDA(ctx, cudaMemcpy(lt.twN, tw.data(), N2 * sizeof(cx<double>), cudaMemcpyHostToDevice));
        if (big) {
            const long long N1 = 1LL << log2n1;
            std::vector<cx<double>> t(1024);
            auto put = [&](void** dst, long long count, long double num, long double den) -> int {
                for (long long k = 0; k < count; ++k) {
                    const long double a = -2.0L * PI * num * (long double)k / den;
                    t[k] = {(double)cosl(a), (double)sinl(a)};
                }
                if (!*dst) PSS_CUDA(ctx, cudaMalloc(dst, count * sizeof(cx<double>)));
                PSS_CUDA(ctx, cudaMemcpy(*dst, t.data(), count * sizeof(cx<double>), cudaMemcpyHostToDevice));
                return PSS_OK;
            };
            if ((rc = put(&lt.tw1, N1 / 2, 1.0L, (long double)N1))) return rc;          // W_N1^k
            if ((rc = put(&lt.thi, N / 1024, 1024.0L, (long double)N))) return rc;       // W_N^(1024 j)
            if ((rc = put(&lt.tlo, 1024, 1.0L, (long double)N))) return rc;              // W_N^j
        }
        for (int kind = PSS_WINDOW_HAMMING; kind <= PSS_WINDOW_HANN; ++kind) {
            const long double a = kind == PSS_WINDOW_HAMMING ? 0.54L : 0.5L, b = kind == PSS_WINDOW_HAMMING ? 0.46L : 0.5L;
            for (long long i = 0; i < N; ++i) w[i] = (double)(a - b * cosl(2.0L * PI * (long double)i / (long double)(N - 1)));
            if (!lt.window[kind]) PSS_CUDA(ctx, cudaMalloc(&lt.window[kind], N * sizeof(double)));
            PSS_CUDA(ctx, cudaMemcpy(lt.window[kind], w.data(), N * sizeof(double), cudaMemcpyHostToDevice));
        }
        lt.ready = true;           // only now: a failed allocation above leaves the tables to be rebuilt
    }
    if (log2n <= 16 && !big) {
        // fused persistent kernel: per-CTA scratch only (grid * N * 16 bytes of fp64 rows + the raw dB row)
        const bool smooth = epilogue == PSS_EPI_SMOOTH_CLAMP;
        // CTAs per frame.  PSS_LARGE_CL=2|4 runs a thread-block cluster per frame (32768 / 65536 points), which
        // halves / quarters the frames in flight so that their fp64 scratch (grid / CL x N x 16 bytes) fits the L2
        // again.  Measured on B200 (ms per GiB, raw / smoothing): 32768 points 1.41 / 2.02 at CL = 1, 1.64 / 2.16 at
        // CL = 2; 65536 points 1.49 / 2.26 at CL = 1, 1.72 / 2.81 at CL = 2, 3.23 / 4.61 at CL = 4 - every CTA's
        // phases are latency-bound (column loads, L2 row loads), so half the work per phase does not take half
        // the time and the cluster barriers add their own.  One CTA per frame stays the default.
        static const int cl_env = getenv("PSS_LARGE_CL") ? atoi(getenv("PSS_LARGE_CL")) : 0;
        int CL = 1;
        if (cl_env == 2 || cl_env == 4) CL = cl_env;
        if (log2n == 14 && CL > 2) CL = 2;               // 4 column passes / 2 row pairs bound the split
        if (log2n == 15 && CL > 4) CL = 4;
        unsigned grid = (unsigned)(ctx->sm_count - ctx->sm_count % CL);
        const long long groups = smooth ? (n_frames + CL - 1) / CL : n_frames;       // units of cluster work
        if ((long long)grid / CL > groups) grid = (unsigned)(groups * CL);
        if ((rc = pss_reserve(ctx, &ctx->p_buf[7], &ctx->p_bytes[7], (size_t)grid / CL * N * 16))) return rc;
        PsdLargeParams lp{};
        lp.iq = reinterpret_cast<const float2*>(iq);
        lp.window = window == PSS_WINDOW_NONE ? nullptr : (const double*)lt.window[window];
        lp.twN = (const cx<double>*)lt.twN;
        lp.tw = (const cx<double>*)tab2->twiddle;
        lp.n_frames = n_frames;
        lp.db = out->db;
        lp.cols = out->cols;
        lp.W = out->W;
        lp.stats = out->stats;
        lp.moments = smooth ? out->moments : nullptr;
        lp.Y = (cx<double>*)ctx->p_buf[7];
        if (smooth) {
            if ((rc = pss_reserve(ctx, &ctx->p_buf[9], &ctx->p_bytes[9], (size_t)grid * N * 4))) return rc;
            lp.rawdb = (float*)ctx->p_buf[9];
            if (log2n == 16) {
                if ((rc = pss_reserve(ctx, &ctx->p_buf[8], &ctx->p_bytes[8], (size_t)grid * N * 4))) return rc;
                lp.srow2 = (float*)ctx->p_buf[8];
            }
            if (CL > 1 && lp.moments) PSS_CUDA(ctx, cudaMemsetAsync(lp.moments, 0, (size_t)n_frames * 32, ctx->stream));
        }
        if (log2n == 14) return CL == 1 ? launch_large_epi<14, 1>(ctx, lp, grid, smooth) : launch_large_epi<14, 2>(ctx, lp, grid, smooth);
        if (log2n == 15)
            return CL == 1 ? launch_large_epi<15, 1>(ctx, lp, grid, smooth)
                           : CL == 2 ? launch_large_epi<15, 2>(ctx, lp, grid, smooth) : launch_large_epi<15, 4>(ctx, lp, grid, smooth);
        return CL == 1 ? launch_large_epi<16, 1>(ctx, lp, grid, smooth)
                       : CL == 2 ? launch_large_epi<16, 2>(ctx, lp, grid, smooth) : launch_large_epi<16, 4>(ctx, lp, grid, smooth);
    }
    // 131072 points: three launches per L2-sized sub-batch.
    // scratch: fp64 rows of a sub-batch (kept <= 64 MB so it lives in L2) and, with an epilogue, the raw rows
    long long sub = (64LL << 20) / (N * 16);
    if (sub < 1) sub = 1;
    if (sub > n_frames) sub = n_frames;
    if ((rc = pss_reserve(ctx, &ctx->p_buf[7], &ctx->p_bytes[7], (size_t)sub * N * 16))) return rc;
    // with an epilogue the raw rows of up to 2*SMs frames (<= 512 MB) are collected first, so the row
    // epilogue runs one CTA per frame over a full grid instead of once per small sub-batch
    const bool smooth = epilogue == PSS_EPI_SMOOTH_CLAMP;
    long long eb = n_frames;
    float* raw = out->db;
    if (smooth) {
        eb = 2LL * ctx->sm_count;
        const long long cap = (512LL << 20) / (N * 4);
        if (eb > cap) eb = cap;
        if (eb < sub) eb = sub;
        if (eb > n_frames) eb = n_frames;
        eb = (eb / sub) * sub > 0 ? (eb / sub) * sub : sub;           // whole sub-batches
        if ((rc = pss_reserve(ctx, &ctx->p_buf[9], &ctx->p_bytes[9], (size_t)eb * N * 4))) return rc;
        raw = (float*)ctx->p_buf[9];
    }
    const long long n_out = smooth ? N - 4 : N;
    for (int64_t e0 = 0; e0 < n_frames; e0 += eb) {
    const long long ne = n_frames - e0 < eb ? n_frames - e0 : eb;
    for (int64_t f0 = e0; f0 < e0 + ne; f0 += sub) {
        const long long nf = e0 + ne - f0 < sub ? e0 + ne - f0 : sub;
        const float2* src = reinterpret_cast<const float2*>(iq) + f0 * N;
        const double* win = window == PSS_WINDOW_NONE ? nullptr : (const double*)lt.window[window];
        const long long threads = nf * N2;
        const unsigned grid = (unsigned)((threads + 255) / 256);
        cx<double>* Y = (cx<double>*)ctx->p_buf[7];
        if (big) {
            const unsigned gb = (unsigned)(nf * (N2 / 16));
            const size_t sm = (size_t)(1 << log2n1) * 16 * sizeof(cx<double>);
#define PSS_BIG(L)                                                                                                   \
    do {                                                                                                             \
        PSS_CUDA(ctx, cudaFuncSetAttribute(psd_colfft_big_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        psd_colfft_big_kernel<L><<<gb, 256, sm, ctx->stream>>>(src, win, (const double2*)lt.tw1, (const double2*)lt.thi, \
                                                               (const double2*)lt.tlo, nf, Y);                       \
    } while (0)
            if (log2n1 == 6) PSS_BIG(6);
            else if (log2n1 == 7) PSS_BIG(7);
            else PSS_BIG(8);
#undef PSS_BIG
        } else if (log2n1 == 2) psd_colfft_kernel<2, double><<<grid, 256, 0, ctx->stream>>>(src, win, (const cx<double>*)lt.twN, (int)N2, nf, Y);
        else if (log2n1 == 3) psd_colfft_kernel<3, double><<<grid, 256, 0, ctx->stream>>>(src, win, (const cx<double>*)lt.twN, (int)N2, nf, Y);
        else psd_colfft_kernel<4, double><<<grid, 256, 0, ctx->stream>>>(src, win, (const cx<double>*)lt.twN, (int)N2, nf, Y);
        PSS_LAUNCH_CHECK(ctx);
        PsdParams p{};
        p.tw = tab2->twiddle;
        p.ystage = Y;
        p.n_frames = nf << log2n1;
        p.db = smooth ? raw + (f0 - e0) * N : out->db + f0 * N;
        if (log2n1 == 6) rc = launch_stage_b<12, 6>(ctx, p);
        else if (log2n1 == 7) rc = launch_stage_b<12, 7>(ctx, p);
        else if (log2n1 == 8) rc = launch_stage_b<12, 8>(ctx, p);
        else if (log2n2 == 13) rc = launch_stage_b<13, 4>(ctx, p);
        else if (log2n1 == 2) rc = launch_stage_b<12, 2>(ctx, p);
        else if (log2n1 == 3) rc = launch_stage_b<12, 3>(ctx, p);
        else rc = launch_stage_b<12, 4>(ctx, p);
        if (rc) return rc;
    }
    if (smooth) {
        // the smoothed row is always produced (median/clamp work on it); use scratch when the caller
        // does not want it
        float* dbo = out->db ? out->db + e0 * n_out : nullptr;
        if (!dbo) {
            if ((rc = pss_reserve(ctx, &ctx->p_buf[8], &ctx->p_bytes[8], (size_t)eb * n_out * 4))) return rc;
            dbo = (float*)ctx->p_buf[8];
        }
        const unsigned g2 = (unsigned)(ne < 2LL * ctx->sm_count ? ne : 2LL * ctx->sm_count);
        row_epilogue_kernel<<<g2, 512, 0, ctx->stream>>>(raw, (int)N, ne, dbo,
                                                        out->cols ? out->cols + e0 * out->W : nullptr, out->W,
                                                        out->stats ? out->stats + e0 * 4 : nullptr);
        PSS_LAUNCH_CHECK(ctx);
    }
    }
    return PSS_OK;
}

// Device-pointer PSD; see pss.h.
extern "C" int pss_psd_c64_dev(pss_ctx* ctx, const float* iq, int N, int64_t n_frames, int window,
                               int epilogue, int precision, const pss_psd_out* out) {
    if (!ctx || !iq || !out || n_frames < 0) return PSS_ERR_ARG;
    if (out->struct_size != sizeof(pss_psd_out)) return PSS_ERR_ARG;
    if (window < 0 || window > 2 || (epilogue != PSS_EPI_RAW && epilogue != PSS_EPI_SMOOTH_CLAMP))
        return PSS_ERR_ARG;
    if (precision != PSS_PREC_FP64 && precision != PSS_PREC_FP32) return PSS_ERR_ARG;
    if (epilogue == PSS_EPI_RAW && !out->db) return PSS_ERR_ARG;
    if (epilogue == PSS_EPI_RAW && (out->cols || out->stats)) return PSS_ERR_UNSUPPORTED;
    if (epilogue == PSS_EPI_RAW && out->moments && (N < 512 || N > 8192)) return PSS_ERR_UNSUPPORTED;
    if (out->moments && N > 65536) return PSS_ERR_UNSUPPORTED;
    if (out->cols && out->W < 1) return PSS_ERR_ARG;
    const int log2n = ilog2_exact(N);
    if (log2n < 0) return PSS_ERR_UNSUPPORTED;
    if (n_frames == 0) return PSS_OK;
    if (n_frames > 0x7fffffffLL) return PSS_ERR_ARG;
    const bool fp32 = precision == PSS_PREC_FP32;
    // the float fast mode exists for the plain spectrum only (it cannot meet the 1e-4 dB bar, see pss.h)
    if (fp32 && (epilogue != PSS_EPI_RAW || log2n > 13)) return PSS_ERR_UNSUPPORTED;
    if (log2n > 13) return psd_large(ctx, iq, log2n, n_frames, window, epilogue, out);
    pss_fft_tables* tab;
    int rc = get_tables(ctx, log2n, &tab, fp32);
    if (rc != PSS_OK) return rc;
    PsdParams p{};
    p.iq = reinterpret_cast<const float2*>(iq);
    p.window = window == PSS_WINDOW_NONE ? nullptr : tab->window[window];
    p.tw = tab->twiddle;
    p.n_frames = n_frames;
    p.db = out->db;
    p.cols = out->cols;
    p.W = out->W;
    p.stats = out->stats;
    p.moments = out->moments;
    if (fp32) return launch_by_n<float, EPI_RAW>(ctx, log2n, p);
    if (epilogue == PSS_EPI_RAW) return launch_by_n<double, EPI_RAW>(ctx, log2n, p);
    // PSS_PSD_SPLIT=1 (measured, not the default: 4096-point 1.11 / 1.03 / 0.98 ms per GiB at 32 / 64 / 128 MB chunks
    // against 0.93 fused): transform and epilogue as two kernels over chunks of frames whose raw dB rows stay in
    // L2, the epilogue at 4 CTAs / SM instead of 2.  The epilogue is bound by its ~64 instructions per bin, not by
    // the barrier latency the extra occupancy hides, so the fused kernel stays the product path.
    static const bool split = getenv("PSS_PSD_SPLIT") != nullptr;
    if (!split || log2n < 12) return launch_by_n<double, EPI_SMOOTH>(ctx, log2n, p);
    static const long long chunk_mb = getenv("PSS_PSD_CHUNK_MB") ? atoll(getenv("PSS_PSD_CHUNK_MB")) : 128;
    long long chunk = (chunk_mb << 20) / ((long long)N * 4);
    if (chunk < 1) chunk = 1;
    if (chunk > n_frames) chunk = n_frames;
    if ((rc = pss_reserve(ctx, &ctx->p_buf[12], &ctx->p_bytes[12], (size_t)chunk * N * 4))) return rc;
    float* raw = (float*)ctx->p_buf[12];
    const size_t n_out = (size_t)N - 4;
    for (int64_t f0 = 0; f0 < n_frames; f0 += chunk) {
        const long long nf = n_frames - f0 < chunk ? n_frames - f0 : chunk;
        PsdParams pr = p;
        pr.iq = p.iq + f0 * N;
        pr.n_frames = nf;
        pr.db = raw;
        pr.cols = nullptr;
        pr.stats = nullptr;
        pr.moments = p.moments ? p.moments + f0 * 4 : nullptr;
        if ((rc = launch_by_n<double, EPI_RAW>(ctx, log2n, pr))) return rc;
        PsdEpiParams q{};
        q.raw = raw;
        q.n_frames = nf;
        q.out = PsdParams{};
        q.out.db = p.db ? p.db + f0 * n_out : nullptr;
        q.out.cols = p.cols ? p.cols + f0 * p.W : nullptr;
        q.out.W = p.W;
        q.out.stats = p.stats ? p.stats + f0 * 4 : nullptr;
        if (p.W > 1) q.out.col_step = (double)(N - 4 - 1) / (double)(p.W - 1);
        const size_t sm = (size_t)N * 8 + 4096;
        if (log2n == 12) {
            PSS_CUDA(ctx, cudaFuncSetAttribute(psd_epilogue_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            psd_epilogue_kernel<12><<<(unsigned)nf, 256, sm, ctx->stream>>>(q);
        } else {
            PSS_CUDA(ctx, cudaFuncSetAttribute(psd_epilogue_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            psd_epilogue_kernel<13><<<(unsigned)nf, 512, sm, ctx->stream>>>(q);
        }
        PSS_LAUNCH_CHECK(ctx);
    }
    return PSS_OK;
}

extern "C" int pss_scan_c64_dev(pss_ctx* ctx, const float* iq, int N, int64_t n_steps, int use_abs,
                                float thr_db, float* peak_db, int32_t* count_above, float* db_rows) {
    if (!ctx || !iq || !peak_db || !count_above || n_steps < 0) return PSS_ERR_ARG;
    const int log2n = ilog2_exact(N);
    if (log2n < 0) return PSS_ERR_UNSUPPORTED;
    if (n_steps == 0) return PSS_OK;
    if (n_steps > 0x7fffffffLL) return PSS_ERR_ARG;
    pss_fft_tables* tab;
    int rc = get_tables(ctx, log2n, &tab);
    if (rc != PSS_OK) return rc;
    PsdParams p{};
    p.iq = reinterpret_cast<const float2*>(iq);
    p.window = nullptr;
    p.tw = tab->twiddle;
    p.n_frames = n_steps;
    p.db = db_rows;
    p.peak = peak_db;
    p.count = count_above;
    p.use_abs = use_abs;
    p.thr_pow = use_abs ? pow(10.0, (double)thr_db / 10.0) : pow(10.0, -(double)thr_db / 10.0);
    return launch_by_n<double, EPI_SCAN>(ctx, log2n, p);
}

// ---------------------------------------------------------------------------------- signal classifier
// classify_signal (signal_processing.py:296-322) = Welch PSD (nperseg 1024, periodic
// Hann, 50 % overlap, per-segment mean removal, density scaling: scipy.signal.welch defaults) ->
// estimate_bandwidth (:267-280), estimate_modulation_index (:283-293), spectral flatness (:304) ->
// decision tree (:306-322).  The reference raises NameError at :299 (`welch` is never imported); this
// is the computation that line intends, checked against the reference run with that one name supplied
// (tests/golden/classifier.npz).  Opt-in from the Python shim; the default keeps raising.
//
// welch_kernel: 4 segments of 1024 per CTA, same radix-16 Stockham passes as the PSD kernel, power
// accumulated per block with fp64 atomics (63 segments per 32768-sample block).
__global__ void __launch_bounds__(256, 2)
welch_kernel(const float2* __restrict__ iq, const long long N_block, const long long n_seg, const long long total_segs,
             const double* __restrict__ hann, const cx<double>* __restrict__ tw, double* __restrict__ acc) {
    constexpr int LOG2N = 10, N = 1024, TPF = 64, FPC = 4, NP = 3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double msum[8][2];
    const int tid = threadIdx.x, f = tid / TPF, t = tid % TPF;
    const long long seg = (long long)blockIdx.x * FPC + f;
    const bool live = seg < total_segs;
    const long long b = live ? seg / n_seg : 0, sg = live ? seg - b * n_seg : 0;
    cx<double>* buf = reinterpret_cast<cx<double>*>(smem_raw) + (size_t)f * N;
    cx<double> wpre[2][8];
#pragma unroll
    for (int ps = 1; ps < NP; ++ps) {
        const int bits = pss_pass_bits(LOG2N, ps), ns = 1 << (4 * ps), items = 16 >> bits;
#pragma unroll
        for (int it = 0; it < 8; ++it)
            if (it < items) wpre[ps - 1][it] = tw[(ns - 16) / 15 + ((t + it * TPF) & (ns - 1))];
    }
    const float2* src = iq + b * N_block + sg * (N / 2);
    cx<double> v[16];
    double sx = 0.0, sy = 0.0;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const float2 s = live ? __ldg(src + t + r * TPF) : make_float2(0.f, 0.f);
        v[r] = {(double)s.x, (double)s.y};
        sx += v[r].x;
        sy += v[r].y;
    }
    sx = warp_sum(sx);
    sy = warp_sum(sy);
    if ((tid & 31) == 0) {
        msum[tid >> 5][0] = sx;
        msum[tid >> 5][1] = sy;
    }
    __syncthreads();
    const double mx = (msum[2 * f][0] + msum[2 * f + 1][0]) * (1.0 / N);      // detrend='constant'
    const double my = (msum[2 * f][1] + msum[2 * f + 1][1]) * (1.0 / N);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const double w = __ldg(hann + t + r * TPF);
        v[r] = {(v[r].x - mx) * w, (v[r].y - my) * w};
    }
    fft_regs<16, double>::run(v);
    const int base = t << 4;
#pragma unroll
    for (int q = 0; q < 16; ++q) buf[fft_swz(base + fft_perm<16>(q))] = v[q];
    __syncthreads();
    double* dst = acc + b * N;
    auto emit = [&](int k, const cx<double> X) {
        if (live) atomicAdd(dst + k, X.x * X.x + X.y * X.y);          // FFT order, like scipy's two-sided output
    };
    stockham_pass<LOG2N, 1, double>(buf, wpre[0], t, [](int, cx<double>) {});
    stockham_pass<LOG2N, 2, double>(buf, wpre[1], t, emit);
}

// One CTA per block: spectral features from the accumulated Welch power, amplitude / phase-step
// variances from the samples, then the decision tree.  feat [n_blocks][4] = signal_bw, modulation
// index, spectral flatness, peak dB; label: 0 UNKNOWN 1 FM_BROADCAST 2 NARROW_FM 3 AM_BROADCAST 4 SSB 5 DIGITAL.
__global__ void __launch_bounds__(256)
classify_kernel(const float2* __restrict__ iq, const int N_block, const long long n_blocks, const double* __restrict__ acc,
                const double psd_scale, const double fs, double* __restrict__ feat, int* __restrict__ label) {
    __shared__ double red[8][6];
    __shared__ float fred[8];
    __shared__ int ired[8][2];
    __shared__ float thr_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (long long b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        // ---- spectrum: float32 like scipy's output for complex64 input
        float p[4], db[4];
        float mx = -INFINITY;
        double slog = 0.0, sp = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            p[q] = (float)(acc[b * 1024 + tid + 256 * q] * psd_scale);
            db[q] = 10.f * log10f(p[q] + 1e-10f);                    // :270
            mx = fmaxf(mx, db[q]);
            slog += (double)logf(p[q] + 1e-10f);                     // :304
            sp += (double)p[q];
        }
        mx = warp_max(mx);
        if (lane == 0) fred[warp] = mx;
        __syncthreads();
        if (tid == 0) {
            float m = fred[0];
            for (int w = 1; w < 8; ++w) m = fmaxf(m, fred[w]);
            fred[0] = m;
            thr_s = m + (-20.f);                                     // :271-274
        }
        __syncthreads();
        const float peak_db = fred[0], thr = thr_s;
        int first = 1 << 20, last = -1;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (db[q] > thr) {
                first = min(first, tid + 256 * q);
                last = max(last, tid + 256 * q);
            }
        // ---- samples: np.abs / np.angle in float32, wrapped phase steps (= diff(unwrap(angle)))
        const float2* x = iq + b * N_block;
        const float PI_F = 3.14159274101257324f, TWO_PI_F = 6.28318548202514648f;
        double sa = 0.0, saa = 0.0, sd = 0.0, sdd = 0.0;
#pragma unroll 4
        for (int ib = 0; ib < N_block; ib += 256) {          // uniform trip count: the shuffle below needs whole warps
            const int i = ib + tid;
            const bool in = i < N_block;
            const float2 s0 = in ? __ldg(x + i) : make_float2(1.f, 0.f);
            const float a = hypotf(s0.x, s0.y);
            const float ang0 = atan2f(s0.y, s0.x);
            // the next sample's angle comes from the next lane; lane 31 computes it itself
            float ang1 = __shfl_down_sync(0xffffffffu, ang0, 1);
            if (lane == 31 && i + 1 < N_block) {
                const float2 s1 = __ldg(x + i + 1);
                ang1 = atan2f(s1.y, s1.x);
            }
            if (in) {
                sa += (double)a;
                saa += (double)a * (double)a;
            }
            if (i + 1 < N_block) {
                const float dd = ang1 - ang0;
                float d = dd;
                if (fabsf(dd) >= PI_F) {                             // np.unwrap: fold the step into (-pi, pi]
                    float m = fmodf(dd + PI_F, TWO_PI_F);
                    if (m < 0.f) m += TWO_PI_F;
                    d = m - PI_F;
                    if (d == -PI_F && dd > 0.f) d = PI_F;
                }
                sd += (double)d;
                sdd += (double)d * (double)d;
            }
        }
        sa = warp_sum(sa); saa = warp_sum(saa); sd = warp_sum(sd); sdd = warp_sum(sdd);
        slog = warp_sum(slog); sp = warp_sum(sp);
        first = __reduce_min_sync(0xffffffffu, first);
        last = __reduce_max_sync(0xffffffffu, last);
        if (lane == 0) {
            red[warp][0] = sa; red[warp][1] = saa; red[warp][2] = sd; red[warp][3] = sdd;
            red[warp][4] = slog; red[warp][5] = sp;
            ired[warp][0] = first; ired[warp][1] = last;
        }
        __syncthreads();
        if (tid == 0) {
            double r[6] = {0, 0, 0, 0, 0, 0};
            int fi = 1 << 20, la = -1;
            for (int w = 0; w < 8; ++w) {
                for (int k = 0; k < 6; ++k) r[k] += red[w][k];
                fi = min(fi, ired[w][0]);
                la = max(la, ired[w][1]);
            }
            const double n = (double)N_block, nd = (double)(N_block - 1);
            const float amp_var = (float)(r[1] / n - (r[0] / n) * (r[0] / n));              // np.var, :290
            const float phase_var = nd > 0 ? (float)(r[3] / nd - (r[2] / nd) * (r[2] / nd)) : 0.f;   // :291
            const float mi = phase_var / (amp_var + 1e-10f);                                // :293
            const float flat = expf((float)(r[4] / 1024.0)) / (float)(r[5] / 1024.0);       // :304
            const double df = fs / 1024.0;
            const double f_first = (fi < 512 ? fi : fi - 1024) * df, f_last = (la < 512 ? la : la - 1024) * df;
            const double bw = la >= 0 ? f_last - f_first : 0.0;                             // :275-280
            int lab = 0;                                                                    // :306-322
            if (bw > 150e3) {
                if (mi > 0.8f) lab = 1;
            } else if (bw >= 8e3 && bw <= 16e3) {
                if (mi < 0.3f) lab = 2;
            } else if (bw >= 8e3 && bw <= 10e3) {
                if (mi < 0.2f && flat < 0.3f) lab = 3;
            } else if (bw >= 2e3 && bw <= 3e3) {
                if (flat < 0.2f) lab = 4;
            } else if (flat > 0.7f) {
                lab = 5;
            }
            feat[b * 4 + 0] = bw;
            feat[b * 4 + 1] = (double)mi;
            feat[b * 4 + 2] = (double)flat;
            feat[b * 4 + 3] = (double)peak_db;
            label[b] = lab;
        }
        __syncthreads();
    }
}

extern "C" int pss_classify_c64_dev(pss_ctx* ctx, const float* iq, int N, int64_t n_blocks, double fs,
                                    double* features, int32_t* label) {
    if (!ctx || !iq || !features || !label || n_blocks < 0 || !(fs > 0)) return PSS_ERR_ARG;
    if (N < 1024) return PSS_ERR_UNSUPPORTED;        // scipy shrinks nperseg below 1024 samples; not mirrored
    if (n_blocks == 0) return PSS_OK;
    int rc;
    pss_fft_tables* tab;
    if ((rc = get_tables(ctx, 10, &tab))) return rc;
    void*& hann = ctx->hann_periodic;
    if (!hann) {
        std::vector<double> w(1024);
        for (int i = 0; i < 1024; ++i)      // scipy get_window('hann', 1024): periodic (fftbins=True)
            w[i] = (double)(0.5L - 0.5L * cosl(2.0L * 3.14159265358979323846264338327950288L * i / 1024.0L));
        PSS_CUDA(ctx, cudaMalloc(&hann, 1024 * sizeof(double)));
        PSS_CUDA(ctx, cudaMemcpy(hann, w.data(), 1024 * sizeof(double), cudaMemcpyHostToDevice));
    }
    const long long n_seg = (N - 512) / 512;          // (N - noverlap) // (nperseg - noverlap)
    const long long total = n_seg * n_blocks;
    if ((rc = pss_reserve(ctx, &ctx->p_buf[8], &ctx->p_bytes[8], (size_t)n_blocks * 1024 * 8))) return rc;
    double* acc = (double*)ctx->p_buf[8];
    PSS_CUDA(ctx, cudaMemsetAsync(acc, 0, (size_t)n_blocks * 1024 * 8, ctx->stream));
    if (ctx->configured.insert((const void*)welch_kernel).second)
        PSS_CUDA(ctx, cudaFuncSetAttribute(welch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    welch_kernel<<<(unsigned)((total + 3) / 4), 256, 65536, ctx->stream>>>(
        reinterpret_cast<const float2*>(iq), N, n_seg, total, (const double*)hann, (const cx<double>*)tab->twiddle, acc);
    PSS_LAUNCH_CHECK(ctx);
    const double scale = 1.0 / ((double)n_seg * fs * 384.0);          // mean over segments / (fs * sum(w^2))
    const long long grid = n_blocks < 4LL * ctx->sm_count ? n_blocks : 4LL * ctx->sm_count;
    classify_kernel<<<(unsigned)grid, 256, 0, ctx->stream>>>(reinterpret_cast<const float2*>(iq), N, n_blocks, acc, scale,
                                                            fs, features, label);
    PSS_LAUNCH_CHECK(ctx);
    return PSS_OK;
}
