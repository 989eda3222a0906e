// Demodulation: the frame kernels (AM / SSB / RAW) and the C ABI of every demodulation plan.
// The decimating NFM / WFM chain lives in pss_demod_decim.cu.
#include "pss_demod.cuh"

int pss_demod_upload(pss_ctx* ctx, pss_demod_plan* pl, const void* src, size_t bytes, const void** dst) {
    void* d = nullptr;
    PSS_CUDA(ctx, cudaMalloc(&d, bytes));
    pl->dev_allocs.push_back(d);
    PSS_CUDA(ctx, cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
    *dst = d;
    return PSS_OK;
}
#define upload pss_demod_upload

// =================================================================================================
// Frame kernels: AM (signal_processing.py:179-195), USB/LSB (:198-217), RAW (:237-238 + :46-80).
// One CTA of 512 threads owns one block; the fp32 working row lives in shared memory, the block is
// read from HBM once and the peak-normalised mono result written once.
// =================================================================================================
#define DEMOD_THREADS 256
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
    if (desc->kind == PSS_PLAN_DECIM) rc = pss_decim_create(ctx, desc, pl);
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
    cudaFree(pl->F_scratch);
    if (pl->side) {
        cudaStreamSynchronize(pl->side);
        cudaStreamDestroy(pl->side);
        for (int i = 0; i < 2; ++i) {
            cudaEventDestroy(pl->ev_force[i]);
            cudaEventDestroy(pl->ev_scan[i]);
        }
    }
    cudaFree(pl->corr);
    cudaFree(pl->mom_scratch);
    cudaFree(pl->tile_scratch);
    delete pl;
}

int pss_demod_plan_out_len(const pss_demod_plan* pl) { return pl ? pl->out_len : 0; }
int pss_demod_plan_block_len(const pss_demod_plan* pl) { return pl ? pl->N : 0; }
int pss_demod_plan_channels(const pss_demod_plan* pl) { return pl ? pl->channels : 0; }

int pss_demod_c64_dev(pss_ctx* ctx, pss_demod_plan* pl, const float* iq, int64_t n_frames, float* audio) {
    if (!ctx || !pl || !iq || !audio || n_frames < 0) return PSS_ERR_ARG;
    if (n_frames == 0) return PSS_OK;
    if (pl->kind == PSS_PLAN_DECIM) return pss_decim_launch(ctx, pl, iq, n_frames, audio, nullptr, 0);
    return launch_frame(ctx, pl, iq, n_frames, audio);
}

int pss_demod_c64_dev_moments(pss_ctx* ctx, pss_demod_plan* pl, const float* iq, int64_t n_frames, float* audio,
                              const double* moments, int frames_per_block, int frame_len) {
    if (!ctx || !pl || !iq || !audio || n_frames < 0) return PSS_ERR_ARG;
    // the moment rows must tile the plan's block exactly, or the kernel would sum rows of another block
    if (moments && (frames_per_block < 1 || (long long)frames_per_block * frame_len != pl->N)) return PSS_ERR_ARG;
    if (n_frames == 0) return PSS_OK;
    if (pl->kind == PSS_PLAN_DECIM && pl->dec.SF == 16) return pss_decim_launch(ctx, pl, iq, n_frames, audio, moments, frames_per_block);
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
