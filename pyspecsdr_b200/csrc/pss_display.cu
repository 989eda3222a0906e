// Display accumulate: per-stream history ring + stack range + normalisation + glyph / colour planes of the
// draw_* functions (pyspecsdr.py:1351-1358, 1373-1398 waterfall; :1649-1696 gradient; :1521-1556
// persistence; :1575-1596 surface; :418-452 spectrum).
//
// Arithmetic contract: every value between the dB rows and the integer planes is computed in fp64 with
// numpy's operation order and NO fused multiply-add (np.interp = slope * (x - xp[j]) + fp[j] with separately
// rounded product and sum, linspace = k * step, (v - min) / range, int() truncation), so that the planes
// are bit-identical to what the reference draws when it is given the same dB rows
// (tests/test_display_gpu.py compares them with the cells the unmodified draw_* functions drew).
#include <math.h>

#include "pss_common.cuh"

// ---------------------------------------------------------------------------------------- exact helpers
__device__ __forceinline__ double nofma_lerp(const double y0, const double y1, const double dx) {
    // numpy arr_interp: slope = (fp[j+1] - fp[j]) / (xp[j+1] - xp[j]) with xp = arange -> / 1.0 (exact)
    return __dadd_rn(__dmul_rn(__dsub_rn(y1, y0), dx), y0);
}

// np.interp(np.linspace(0, n - 1, W), np.arange(n), row)[c]; `step` = (n - 1) / (W - 1) from the host.
template <typename F>
__device__ __forceinline__ double interp_col(F&& row_at, const int n, const int W, const int c, const double step) {
    const double x = (c == W - 1 && W > 1) ? (double)(n - 1) : __dmul_rn((double)c, step);
    const int j = (int)x;
    if (j >= n - 1) return row_at(n - 1);
    const double y0 = row_at(j), y1 = row_at(j + 1);
    double r = nofma_lerp(y0, y1, __dsub_rn(x, (double)j));
    if (r != r) {                                  // numpy retries from the right neighbour (inf/nan rows)
        r = __dadd_rn(__dmul_rn(__dsub_rn(y1, y0), __dsub_rn(x, (double)(j + 1))), y1);
        if (r != r && y0 == y1) r = y0;
    }
    return r;
}

__device__ __forceinline__ bool finite_d(const double v) { return fabs(v) <= 1.7976931348623157e308; }

__device__ __forceinline__ uint8_t clamp_u8(const int v) { return (uint8_t)min(max(v, 0), 254); }

// glyph / colour planes from a normalised value (fp64), int() = truncation toward zero
__device__ __forceinline__ void quantise(const double v, const int kind, const int H, const int row_colour,
                                         uint8_t& a, uint8_t& b) {
    if (!finite_d(v)) {
        a = b = 255;
        return;
    }
    const int colour = (int)__dmul_rn(v, 5.0);                                             // :1388 / :1695
    if (kind == PSS_QUANT_WATERFALL) {
        a = (uint8_t)((v > 0.25) + (v > 0.5) + (v > 0.75));                                // '.', '-', '=', '#' :1390-1397
        b = clamp_u8(colour);
    } else if (kind == PSS_QUANT_GRADIENT) {
        a = clamp_u8((int)__dmul_rn(v, 8.0));                                              // ' ._-=+*#@' :1691
        b = clamp_u8(colour);
    } else if (kind == PSS_QUANT_PERSISTENCE) {
        const int y = (int)__dmul_rn(__dsub_rn(1.0, v), (double)(H - 1));                  // :1556
        a = (y >= 0 && y < H && y < 255) ? (uint8_t)y : 255;                               // drawn only if 0 <= y < H :1557
        b = clamp_u8(row_colour);                                                          // :1544-1545, per trace
    } else {
        a = clamp_u8((int)__dmul_rn(v, 20.0));                                             // surface magnitude :1593
        b = clamp_u8(colour);
    }
}

// ---------------------------------------------------------------------------------------- history render
// Rows live in two places: `prev` = the ring carried from earlier calls (the last n_prev of its rows_max-1
// slots are valid, oldest first), `cur` = the rows of this call in time order.  Row index f of the call's
// time axis: f >= 0 -> cur[f], f < 0 -> prev[(rows_max - 1) + f].
template <typename T>
struct RenderArgs {
    const T* cur_cols;     // [n_frames][W]
    const T* cur_mm;       // min at [f * mm_stride + mm_off], max at +1
    int mm_stride, mm_off;
    const T* prev_cols;    // [(rows_max - 1)][W]
    const T* prev_mm;      // [(rows_max - 1)][2]
    int n_prev;
    int W, rows_max, kind, H, guard;
    long long n_frames, first, step;
    float* norm;
    float* minmax;
    double* norm64;
    double* minmax64;
    uint8_t* plane_a;
    uint8_t* plane_b;
    int* n_rows;
    int colours[32];       // persistence: colour pair of the trace that is k-th newest in a history of full length
};

// PLANES = false: the lean instantiation for callers that only want the normalised values (the bench step).
template <typename T, bool PLANES>
__global__ void __launch_bounds__(256) display_render_kernel(const RenderArgs<T> A) {
    pss_grid_dependency_sync();                        // PSS_PDL: the rows come from the previous kernel of the stream
    const long long r = blockIdx.x;
    const long long t = A.first + r * A.step;          // newest frame of this render
    if (t < 0 || t >= A.n_frames) return;
    const int R = A.rows_max, W = A.W;
    const long long avail = t + 1 + A.n_prev;
    const int len = (int)(avail < R ? avail : R);      // rows in the history at this render
    // stack range over the history rows (each row's finite min/max came with the row): lane y of the first warp
    // takes row y, a warp reduction, one broadcast through shared memory (rows_max <= 31)
    __shared__ double s_lohi[2];
    if (threadIdx.x < 32) {
        const int y = threadIdx.x;
        T a = (T)INFINITY, b = (T)-INFINITY;              // rows without a finite value carry (+inf, -inf)
        if (y < len) {
            const long long f = t - y;
            if (f >= 0) {
                a = A.cur_mm[f * A.mm_stride + A.mm_off];
                b = A.cur_mm[f * A.mm_stride + A.mm_off + 1];
            } else {
                a = A.prev_mm[((R - 1) + f) * 2];
                b = A.prev_mm[((R - 1) + f) * 2 + 1];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const T oa = __shfl_xor_sync(0xffffffffu, a, o), ob = __shfl_xor_sync(0xffffffffu, b, o);
            a = oa < a ? oa : a;                            // fmin / fmax semantics for the finite values stored here
            b = ob > b ? ob : b;
        }
        if (threadIdx.x == 0) {
            s_lohi[0] = (double)a;
            s_lohi[1] = (double)b;
        }
    }
    __syncthreads();
    const double lo = s_lohi[0], hi = s_lohi[1];
    double range = __dsub_rn(hi, lo);
    if (A.guard && range == 0.0) range = 1.0;
    if (threadIdx.x == 0) {
        if (A.minmax) {
            A.minmax[2 * r] = (float)lo;
            A.minmax[2 * r + 1] = (float)hi;
        }
        if (A.minmax64) {
            A.minmax64[2 * r] = lo;
            A.minmax64[2 * r + 1] = hi;
        }
        if (A.n_rows) A.n_rows[r] = len;
    }
    const long long base = r * (long long)R * W;
    const float lo_f = (float)lo, range_f = (float)range;
    if constexpr (!PLANES && sizeof(T) == 4) {
        // float32 rows, values only (the pipeline / bench step): four columns per thread, 16-byte loads and stores
        const bool vec = A.norm && (W & 3) == 0 &&
                         ((((uintptr_t)A.cur_cols | (uintptr_t)A.prev_cols | (uintptr_t)A.norm) & 15) == 0);
        if (vec) {
            constexpr int UN = 3;
            const int W4 = W >> 2, items = R * W4, nt = blockDim.x;
            const float4* cur4 = reinterpret_cast<const float4*>(A.cur_cols);
            const float4* prev4 = reinterpret_cast<const float4*>(A.prev_cols);
            float4* out4 = reinterpret_cast<float4*>(A.norm + base);
            const float qnan = __int_as_float(0x7fc00000);
            const int dy = nt / W4, dc = nt - dy * W4;
            int y = (int)threadIdx.x / W4, c = (int)threadIdx.x - y * W4;
            for (int i0 = threadIdx.x; i0 < items; i0 += UN * nt) {
                float4 v[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    v[u] = make_float4(qnan, qnan, qnan, qnan);
                    if (i0 + u * nt < items && y < len) {
                        const long long f = t - y;
                        const float4* src = f >= 0 ? cur4 + f * W4 : prev4 + ((R - 1) + f) * W4;
                        v[u] = __ldg(src + c);
                    }
                    c += dc;
                    y += dy;
                    if (c >= W4) {
                        c -= W4;
                        ++y;
                    }
                }
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    if (i0 + u * nt >= items) break;
                    float4 o;
                    o.x = fabsf(v[u].x) <= 3.402823466e38f ? __fdiv_rn(__fsub_rn(v[u].x, lo_f), range_f) : qnan;
                    o.y = fabsf(v[u].y) <= 3.402823466e38f ? __fdiv_rn(__fsub_rn(v[u].y, lo_f), range_f) : qnan;
                    o.z = fabsf(v[u].z) <= 3.402823466e38f ? __fdiv_rn(__fsub_rn(v[u].z, lo_f), range_f) : qnan;
                    o.w = fabsf(v[u].w) <= 3.402823466e38f ? __fdiv_rn(__fsub_rn(v[u].w, lo_f), range_f) : qnan;
                    out4[i0 + u * nt] = o;
                }
            }
            return;
        }
    }
    // thread = column, loop over the history rows (UN rows in flight): no index arithmetic per element
    constexpr int UN = 5;
    const T nanT = sizeof(T) == 8 ? (T)__longlong_as_double(0x7ff8000000000000LL) : (T)__int_as_float(0x7fc00000);
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
        for (int y0 = 0; y0 < R; y0 += UN) {
            T col[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int y = y0 + u;
                col[u] = nanT;
                if (y < len) {
                    const long long f = t - y;
                    col[u] = f >= 0 ? __ldg(A.cur_cols + f * W + c) : __ldg(A.prev_cols + ((R - 1) + f) * W + c);
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int y = y0 + u;
                if (y >= R) break;
                const long long e = base + (long long)y * W + c;
                double v;
                float vf;
                if constexpr (sizeof(T) == 8) {                       // numpy's fp64 arithmetic
                    v = finite_d(col[u]) ? __ddiv_rn(__dsub_rn(col[u], lo), range) : __longlong_as_double(0x7ff8000000000000LL);
                    vf = (float)v;
                } else {                                              // float32 rows: float32 value
                    vf = fabsf(col[u]) <= 3.402823466e38f ? __fdiv_rn(__fsub_rn(col[u], lo_f), range_f) : __int_as_float(0x7fc00000);
                    v = (double)vf;
                }
                if (A.norm) A.norm[e] = vf;
                if constexpr (PLANES || sizeof(T) == 8) {
                    if (A.norm64) A.norm64[e] = v;
                }
                if constexpr (PLANES) if (A.plane_a) {
                    uint8_t pa, pb;
                    // persistence trace i (oldest = 0) of a history of `len` rows: alpha = 0.7 ** (rows_max - i), i = len-1-y
                    quantise(v, A.kind, A.H, A.colours[min(31, max(0, R - (len - 1 - y)))], pa, pb);
                    A.plane_a[e] = pa;
                    if (A.plane_b) A.plane_b[e] = pb;
                }
            }
        }
    }
}

// new ring = the last rows_max-1 rows of (old ring ++ this call's rows)
template <typename T>
__global__ void display_carry_kernel(const RenderArgs<T> A, T* __restrict__ next_cols, T* __restrict__ next_mm) {
    const int R1 = A.rows_max - 1, W = A.W;
    const int slot = blockIdx.x;                        // 0 .. R1-1, oldest first
    const long long f = A.n_frames - R1 + slot;         // position on this call's time axis
    const bool from_cur = f >= 0;
    const long long p = R1 + f;                         // index in the old ring
    const bool valid = from_cur || p >= R1 - A.n_prev;
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
        T v = (T)0;
        if (valid) v = from_cur ? A.cur_cols[f * W + c] : A.prev_cols[p * W + c];
        next_cols[(long long)slot * W + c] = v;
    }
    if (threadIdx.x < 2) {
        T v = (T)0;
        if (valid) v = from_cur ? A.cur_mm[f * A.mm_stride + A.mm_off + threadIdx.x] : A.prev_mm[p * 2 + threadIdx.x];
        next_mm[slot * 2 + threadIdx.x] = v;
    }
}

// fp64 dB rows -> W-column np.interp resample + finite min / max of the row (one CTA per row).
// SURFACE rows are normalised by their own range BEFORE the resample, like draw_surface_plot (:1580-1587).
__global__ void __launch_bounds__(256)
display_rows_kernel(const double* __restrict__ rows, const int n, const long long n_rows, const int W, const double step,
                    const int surface, double* __restrict__ cols, double* __restrict__ mm) {
    __shared__ double smn[8], smx[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (long long f = blockIdx.x; f < n_rows; f += gridDim.x) {
        const double* row = rows + f * n;
        double mn = INFINITY, mx = -INFINITY;
        for (int i = tid; i < n; i += 256) {
            const double v = row[i];
            if (finite_d(v)) {
                mn = fmin(mn, v);
                mx = fmax(mx, v);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        __syncthreads();
        if (lane == 0) {
            smn[warp] = mn;
            smx[warp] = mx;
        }
        __syncthreads();
        mn = smn[0];
        mx = smx[0];
        for (int w = 1; w < 8; ++w) {
            mn = fmin(mn, smn[w]);
            mx = fmax(mx, smx[w]);
        }
        if (tid == 0) {
            mm[2 * f] = mn;
            mm[2 * f + 1] = mx;
        }
        double range = __dsub_rn(mx, mn);
        if (range == 0.0) range = 1.0;
        for (int c = tid; c < W; c += 256) {
            double v;
            if (surface) v = interp_col([&](int j) { return __ddiv_rn(__dsub_rn(row[j], mn), range); }, n, W, c, step);
            else v = interp_col([&](int j) { return row[j]; }, n, W, c, step);
            cols[f * W + c] = v;
        }
    }
}

// surface rows: the resampled values are already normalised; magnitude = int(v * 20)
__global__ void surface_planes_kernel(const double* __restrict__ cols, const long long n, float* __restrict__ norm,
                                      double* __restrict__ norm64, uint8_t* __restrict__ a, uint8_t* __restrict__ b) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = cols[i];
    if (norm) norm[i] = (float)v;
    if (norm64) norm64[i] = v;
    if (a) {
        uint8_t pa, pb;
        quantise(v, PSS_QUANT_SURFACE, 0, 0, pa, pb);
        a[i] = pa;
        if (b) b[i] = pb;
    }
}


// Exact order statistics of a row for the percentile: values of the elements of 0-based rank `rank` and
// rank + 1 (clamped to n - 1), as doubles.  float rows use the 32-bit key select of pss_common.cuh; double
// rows run an 8 x 8-bit radix select on 64-bit order-preserving keys.  All 512 threads must call it.
__device__ __forceinline__ unsigned long long d2key(const double d) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return u ^ ((u >> 63) ? 0xffffffffffffffffULL : 0x8000000000000000ULL);
}
__device__ __forceinline__ double key2d(const unsigned long long k) {
    const unsigned long long u = (k >> 63) ? (k ^ 0x8000000000000000ULL) : ~k;
    return __longlong_as_double((long long)u);
}

template <typename T>
__device__ __forceinline__ void row_select2(const T* __restrict__ row, const int n, const unsigned rank_in,
                                            unsigned* hist, unsigned* us, double& xa, double& xb);

template <>
__device__ __forceinline__ void row_select2<float>(const float* __restrict__ row, const int n, const unsigned rank_in,
                                                   unsigned* hist, unsigned* us, double& xa, double& xb) {
    unsigned ka, kb;
    row_select2_512(row, n, rank_in, hist, us, ka, kb);
    xa = (double)key2f(ka);
    xb = (double)key2f(kb);
}

template <>
__device__ __forceinline__ void row_select2<double>(const double* __restrict__ row, const int n, const unsigned rank_in,
                                                    unsigned* hist, unsigned* us, double& xa, double& xb) {
    __shared__ unsigned long long s64[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned rank = rank_in;
    unsigned long long prefix = 0ULL;
    for (int ps = 0; ps < 8; ++ps) {
        const int shift = 56 - 8 * ps;
        if (tid < 256) hist[tid] = 0u;
        if (tid == 0) {
            us[4] = 0u;
            s64[0] = 0xffffffffffffffffULL;
        }
        __syncthreads();
        for (int i = tid; i < n; i += 512) {
            const unsigned long long k = d2key(row[i]);
            if (ps == 0 || (k >> (shift + 8)) == prefix) atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (warp == 0) {
            unsigned c[8], sum = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                c[q] = hist[8 * lane + q];
                sum += c[q];
            }
            unsigned incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            const unsigned hit = __ballot_sync(0xffffffffu, incl > rank);
            const int Ln = __ffs(hit) - 1;
            if (lane == Ln) {
                unsigned r = rank - (incl - sum);
                int dg = 0;
                bool found = false;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (!found) {
                        if (r < c[q]) { dg = q; found = true; }
                        else r -= c[q];
                    }
                }
                us[2] = (unsigned)(8 * lane + dg);
                us[3] = r;
            }
        }
        __syncthreads();
        prefix = (prefix << 8) | us[2];
        rank = us[3];
    }
    unsigned cnt_le = 0;
    unsigned long long min_gt = 0xffffffffffffffffULL;
    for (int i = tid; i < n; i += 512) {
        const unsigned long long k = d2key(row[i]);
        cnt_le += k <= prefix;
        if (k > prefix && k < min_gt) min_gt = k;
    }
    cnt_le = __reduce_add_sync(0xffffffffu, cnt_le);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, min_gt, o);
        min_gt = other < min_gt ? other : min_gt;
    }
    if (lane == 0) {
        atomicAdd(&us[4], cnt_le);
        atomicMin(&s64[0], min_gt);
    }
    __syncthreads();
    xa = key2d(prefix);
    xb = (us[4] > rank_in + 1u || s64[0] == 0xffffffffffffffffULL) ? xa : key2d(s64[0]);
    __syncthreads();
}

// ---------------------------------------------------------------------------------------- spectrum view
// draw_spectrogram's numeric part (pyspecsdr.py:418-452): 20th-percentile noise floor (np.percentile,
// linear interpolation between order statistics), display range, clip, ** 0.7, W-column resample.
// `k`, `frac`: numpy's virtual index of the 20th percentile, computed on the host with numpy's formula.
template <typename T>
__global__ void __launch_bounds__(512)
spectrum_normalise_kernel(const T* __restrict__ db, const int n, const long long n_frames, const int W,
                          const unsigned k, const double frac, const double col_step, float* __restrict__ cols,
                          float* __restrict__ range_out, double* __restrict__ cols64, double* __restrict__ range64) {
    __shared__ unsigned hist[256];
    __shared__ unsigned us[8];
    __shared__ double fmx[16];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (long long f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const T* row = db + f * n;
        // noise floor: np.percentile(row, 20) = x[k] + (x[k+1] - x[k]) * frac on the sorted row
        double xa, xb;
        row_select2<T>(row, n, k, hist, us, xa, xb);
        const double floor_db = frac < 0.5 ? __dadd_rn(xa, __dmul_rn(__dsub_rn(xb, xa), frac))
                                           : __dsub_rn(xb, __dmul_rn(__dsub_rn(xb, xa), __dsub_rn(1.0, frac)));
        double mx = -INFINITY;
        for (int i = tid; i < n; i += 512) mx = fmax(mx, (double)row[i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) fmx[warp] = mx;
        __syncthreads();
        mx = fmx[0];
        for (int w = 1; w < 16; ++w) mx = fmax(mx, fmx[w]);
        const double span = __dsub_rn(mx, floor_db);                                    // :425
        const double dmin = __dsub_rn(floor_db, __dmul_rn(span, 0.1));                  // :426
        const double dmax = __dadd_rn(mx, __dmul_rn(span, 0.05));                       // :427
        if (tid == 0) {
            if (range_out) {
                range_out[2 * f] = (float)dmin;
                range_out[2 * f + 1] = (float)dmax;
            }
            if (range64) {
                range64[2 * f] = dmin;
                range64[2 * f + 1] = dmax;
            }
        }
        const double den = __dsub_rn(dmax, dmin);
        auto value = [&](int j) {
            const double u = __ddiv_rn(__dsub_rn((double)row[j], dmin), den);           // :442
            return pow(fmin(fmax(u, 0.0), 1.0), 0.7);                                   // np.clip, np.power :442, :445
        };
        for (int c = tid; c < W; c += 512) {
            const double v = interp_col(value, n, W, c, col_step);
            if (cols) cols[f * W + c] = (float)v;
            if (cols64) cols64[f * W + c] = v;
        }
        __syncthreads();
    }
}

// Glyph / colour index planes from already-normalised float32 values.
__global__ void display_quantise_kernel(const float* __restrict__ norm, const long long n, const int kind,
                                        const int H, uint8_t* __restrict__ a, uint8_t* __restrict__ b) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t pa, pb;
    quantise((double)norm[i], kind, H, 0, pa, pb);
    a[i] = pa;
    if (b) b[i] = kind == PSS_QUANT_PERSISTENCE ? clamp_u8((int)__dmul_rn((double)norm[i], 5.0)) : pb;
}

// ---------------------------------------------------------------------------------------- host side
struct DisplayRing {
    int kind = 0, W = 0, rows_max = 0, H = 0;
    int elem = 0;                   // 0 = not fixed yet, 4 = float rows (PSD kernel output), 8 = double rows
    int n_prev = 0, cur = 0;
    void* cols[2] = {nullptr, nullptr};
    void* mm[2] = {nullptr, nullptr};
    int colours[32] = {};
};

static void ring_free(DisplayRing* r) {
    if (!r) return;
    for (int i = 0; i < 2; ++i) {
        cudaFree(r->cols[i]);
        cudaFree(r->mm[i]);
    }
    delete r;
}

void pss_display_release(pss_ctx* ctx) {
    for (auto& kv : ctx->displays) ring_free((DisplayRing*)kv.second);
    ctx->displays.clear();
}

int pss_display_geom(const pss_ctx* ctx, int stream, int* W, int* rows_max) {
    auto it = ctx->displays.find(stream);
    if (it == ctx->displays.end()) return PSS_ERR_ARG;
    const DisplayRing* r = (const DisplayRing*)it->second;
    *W = r->W;
    *rows_max = r->rows_max;
    return PSS_OK;
}

static int guard_of(int kind) { return kind == PSS_QUANT_WATERFALL ? 0 : 1; }      // :1528-1530, :1657-1659

static void persistence_colours(int rows_max, int* out) {
    // trace i of the history list: alpha = PERSISTENCE_ALPHA ** (PERSISTENCE_LENGTH - i),
    // color_pair = int(1 + 5 * (1 - alpha))  (pyspecsdr.py:151-152, 1544-1545); index here = LENGTH - i
    for (int k = 0; k < 32; ++k) {
        const double alpha = pow(0.7, (double)k);
        out[k] = (int)(1 + (5 * (1 - alpha)));
    }
    (void)rows_max;
}

template <typename T>
static int run_render(pss_ctx* ctx, DisplayRing* ring, RenderArgs<T>& A, int64_t n_renders, bool carry) {
    if (n_renders > 0) {
        if (A.plane_a || A.norm64) PSS_CUDA(ctx, pss_launch(display_render_kernel<T, true>, (unsigned)n_renders, 256u, 0, ctx->stream, A));
        else PSS_CUDA(ctx, pss_launch(display_render_kernel<T, false>, (unsigned)n_renders, 256u, 0, ctx->stream, A));
        PSS_LAUNCH_CHECK(ctx);
    }
    if (carry && ring && ring->rows_max > 1) {
        const int nx = ring->cur ^ 1;
        display_carry_kernel<T><<<ring->rows_max - 1, 128, 0, ctx->stream>>>(A, (T*)ring->cols[nx], (T*)ring->mm[nx]);
        PSS_LAUNCH_CHECK(ctx);
        ring->cur = nx;
        const long long np = (long long)ring->n_prev + A.n_frames;
        ring->n_prev = (int)(np < ring->rows_max - 1 ? np : ring->rows_max - 1);
    }
    return PSS_OK;
}

static int ring_fix_elem(pss_ctx* ctx, DisplayRing* ring, int elem) {
    if (ring->elem == elem) return PSS_OK;
    if (ring->elem != 0 && ring->n_prev > 0) return PSS_ERR_ARG;      // a stream is fed float OR double rows
    for (int i = 0; i < 2; ++i) {
        cudaFree(ring->cols[i]);
        cudaFree(ring->mm[i]);
        ring->cols[i] = ring->mm[i] = nullptr;
    }
    const size_t rows = (size_t)(ring->rows_max > 1 ? ring->rows_max - 1 : 1);
    for (int i = 0; i < 2; ++i) {
        PSS_CUDA(ctx, cudaMalloc(&ring->cols[i], rows * ring->W * elem));
        PSS_CUDA(ctx, cudaMalloc(&ring->mm[i], rows * 2 * elem));
    }
    ring->elem = elem;
    ring->n_prev = 0;
    ring->cur = 0;
    return PSS_OK;
}

static DisplayRing* find_ring(pss_ctx* ctx, int stream) {
    auto it = ctx->displays.find(stream);
    return it == ctx->displays.end() ? nullptr : (DisplayRing*)it->second;
}

template <typename T>
static void fill_args(RenderArgs<T>& A, const DisplayRing* ring, const pss_display_out* out) {
    A.W = ring->W;
    A.rows_max = ring->rows_max;
    A.kind = ring->kind;
    A.H = ring->H;
    A.guard = guard_of(ring->kind);
    A.prev_cols = (const T*)ring->cols[ring->cur];
    A.prev_mm = (const T*)ring->mm[ring->cur];
    A.n_prev = ring->n_prev;
    memcpy(A.colours, ring->colours, sizeof A.colours);
    A.norm = out->norm;
    A.minmax = out->minmax;
    A.norm64 = out->norm64;
    A.minmax64 = out->minmax64;
    A.plane_a = out->plane_a;
    A.plane_b = out->plane_b;
    A.n_rows = out->n_rows;
}

extern "C" {

int pss_display_open(pss_ctx* ctx, int stream, int kind, int W, int rows_max, int H) {
    if (!ctx || W < 1 || rows_max < 1 || rows_max > 31 || kind < 0 || kind > 3) return PSS_ERR_ARG;
    if (kind == PSS_QUANT_PERSISTENCE && H < 2) return PSS_ERR_ARG;
    if (kind == PSS_QUANT_SURFACE && rows_max != 1) return PSS_ERR_ARG;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    pss_display_close(ctx, stream);
    DisplayRing* r = new (std::nothrow) DisplayRing();
    if (!r) return PSS_ERR_NOMEM;
    r->kind = kind;
    r->W = W;
    r->rows_max = rows_max;
    r->H = H;
    persistence_colours(rows_max, r->colours);
    ctx->displays[stream] = r;
    return PSS_OK;
}

int pss_display_close(pss_ctx* ctx, int stream) {
    if (!ctx) return PSS_ERR_ARG;
    auto it = ctx->displays.find(stream);
    if (it == ctx->displays.end()) return PSS_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ring_free((DisplayRing*)it->second);
    ctx->displays.erase(it);
    return PSS_OK;
}

int pss_display_rows(const pss_ctx* ctx, int stream) {
    if (!ctx) return PSS_ERR_ARG;
    auto it = ctx->displays.find(stream);
    if (it == ctx->displays.end()) return PSS_ERR_ARG;
    const DisplayRing* r = (const DisplayRing*)it->second;
    return r->rows_max > 1 ? r->n_prev : 0;
}

// float32 rows from the PSD kernel (device pointers, asynchronous on the context's stream)
int pss_display_accumulate_dev(pss_ctx* ctx, int stream, const float* cols, const float* stats, int64_t n_frames,
                               int64_t first, int64_t step, int64_t n_renders, const pss_display_out* out) {
    if (!ctx || !cols || !stats || !out || n_frames < 0 || n_renders < 0) return PSS_ERR_ARG;
    if (out->struct_size != sizeof(pss_display_out)) return PSS_ERR_ARG;
    DisplayRing* ring = find_ring(ctx, stream);
    if (!ring) return PSS_ERR_ARG;
    if (n_frames == 0) return PSS_OK;
    if (n_renders > 0 && (first < 0 || step < 0 || first + (n_renders - 1) * step >= n_frames)) return PSS_ERR_ARG;
    if (n_renders > 0x7fffffffLL) return PSS_ERR_ARG;
    int rc;
    if ((rc = ring_fix_elem(ctx, ring, 4))) return rc;
    RenderArgs<float> A{};
    fill_args(A, ring, out);
    A.cur_cols = cols;
    A.cur_mm = stats;
    A.mm_stride = 4;            // stats rows: max, mean, finite-min, finite-max
    A.mm_off = 2;
    A.n_frames = n_frames;
    A.first = first;
    A.step = step;
    return run_render<float>(ctx, ring, A, n_renders, true);
}

// fp64 dB rows in HOST memory: one render after every row (the call pattern of the draw_* functions).
int pss_display_accumulate_f64(pss_ctx* ctx, int stream, const double* rows, int n_bins, int64_t n_rows,
                               const pss_display_out* out) {
    if (!ctx || !rows || !out || n_bins < 2 || n_rows < 0) return PSS_ERR_ARG;
    if (out->struct_size != sizeof(pss_display_out)) return PSS_ERR_ARG;
    DisplayRing* ring = find_ring(ctx, stream);
    if (!ring) return PSS_ERR_ARG;
    if (n_rows == 0) return PSS_OK;
    if (n_rows > 0x7fffffffLL) return PSS_ERR_ARG;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ring_fix_elem(ctx, ring, 8))) return rc;
    const int W = ring->W, R = ring->rows_max;
    const size_t in_b = (size_t)n_rows * n_bins * 8, cols_b = (size_t)n_rows * W * 8, mm_b = (size_t)n_rows * 16;
    const size_t cells = (size_t)n_rows * R * W;
    if ((rc = pss_reserve(ctx, &ctx->d_in, &ctx->d_in_bytes, in_b))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_aux, &ctx->d_aux_bytes, cols_b))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_aux2, &ctx->d_aux2_bytes, mm_b + (size_t)n_rows * 16 + (size_t)n_rows * 4 + 64))) return rc;
    // outputs: norm64 | norm | plane_a | plane_b
    if ((rc = pss_reserve(ctx, &ctx->d_out, &ctx->d_out_bytes, cells * (8 + 4 + 2) + 64))) return rc;
    double* d_cols = (double*)ctx->d_aux;
    double* d_mm = (double*)ctx->d_aux2;
    double* d_mm64 = d_mm + 2 * n_rows;
    int* d_nrows = (int*)(d_mm64 + 2 * n_rows);
    double* d_norm64 = (double*)ctx->d_out;
    float* d_norm = (float*)(d_norm64 + cells);
    uint8_t* d_pa = (uint8_t*)(d_norm + cells);
    uint8_t* d_pb = d_pa + cells;
    PSS_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, rows, in_b, cudaMemcpyHostToDevice, ctx->stream));
    const double step = W > 1 ? (double)(n_bins - 1) / (double)(W - 1) : 0.0;
    const bool surface = ring->kind == PSS_QUANT_SURFACE;
    const unsigned grid = (unsigned)(n_rows < 8LL * ctx->sm_count ? n_rows : 8LL * ctx->sm_count);
    display_rows_kernel<<<grid, 256, 0, ctx->stream>>>((const double*)ctx->d_in, n_bins, n_rows, W, step, surface ? 1 : 0,
                                                       d_cols, d_mm);
    PSS_LAUNCH_CHECK(ctx);
    if (surface) {
        surface_planes_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, ctx->stream>>>(
            d_cols, (long long)cells, out->norm ? d_norm : nullptr, out->norm64 ? d_norm64 : nullptr,
            out->plane_a ? d_pa : nullptr, out->plane_b ? d_pb : nullptr);
        PSS_LAUNCH_CHECK(ctx);
        if (out->minmax64) PSS_CUDA(ctx, cudaMemcpyAsync(out->minmax64, d_mm, mm_b, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        pss_display_out dev = *out;
        dev.norm = out->norm ? d_norm : nullptr;
        dev.norm64 = out->norm64 ? d_norm64 : nullptr;
        dev.plane_a = out->plane_a ? d_pa : nullptr;
        dev.plane_b = out->plane_b ? d_pb : nullptr;
        dev.minmax = nullptr;
        dev.minmax64 = d_mm64;
        dev.n_rows = d_nrows;
        RenderArgs<double> A{};
        fill_args(A, ring, &dev);
        A.cur_cols = d_cols;
        A.cur_mm = d_mm;
        A.mm_stride = 2;
        A.mm_off = 0;
        A.n_frames = n_rows;
        A.first = 0;
        A.step = 1;
        if ((rc = run_render<double>(ctx, ring, A, n_rows, true))) return rc;
        if (out->minmax64) PSS_CUDA(ctx, cudaMemcpyAsync(out->minmax64, d_mm64, mm_b, cudaMemcpyDeviceToHost, ctx->stream));
        if (out->n_rows) PSS_CUDA(ctx, cudaMemcpyAsync(out->n_rows, d_nrows, (size_t)n_rows * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (out->norm64) PSS_CUDA(ctx, cudaMemcpyAsync(out->norm64, d_norm64, cells * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (out->norm) PSS_CUDA(ctx, cudaMemcpyAsync(out->norm, d_norm, cells * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out->plane_a) PSS_CUDA(ctx, cudaMemcpyAsync(out->plane_a, d_pa, cells, cudaMemcpyDeviceToHost, ctx->stream));
    if (out->plane_b) PSS_CUDA(ctx, cudaMemcpyAsync(out->plane_b, d_pb, cells, cudaMemcpyDeviceToHost, ctx->stream));
    PSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PSS_OK;
}

// Stateless render over rows that are all part of this call (no carried history).
int pss_display_render_dev(pss_ctx* ctx, const float* cols, const float* stats, int W, int64_t n_frames,
                           int rows_max, int64_t first, int64_t step, int64_t n_renders,
                           int guard_zero_range, float* norm, float* minmax) {
    if (!ctx || !cols || !stats || !norm || !minmax || W < 1 || rows_max < 1 || n_frames < 0 || n_renders < 0)
        return PSS_ERR_ARG;
    if (n_renders == 0) return PSS_OK;
    if (first < 0 || step < 0 || first + (n_renders - 1) * step >= n_frames) return PSS_ERR_ARG;
    if (n_renders > 0x7fffffffLL) return PSS_ERR_ARG;
    RenderArgs<float> A{};
    A.cur_cols = cols;
    A.cur_mm = stats;
    A.mm_stride = 4;
    A.mm_off = 2;
    A.W = W;
    A.rows_max = rows_max;
    A.kind = guard_zero_range ? PSS_QUANT_GRADIENT : PSS_QUANT_WATERFALL;
    A.guard = guard_zero_range;
    A.n_frames = n_frames;
    A.first = first;
    A.step = step;
    A.norm = norm;
    A.minmax = minmax;
    return run_render<float>(ctx, nullptr, A, n_renders, false);
}

int pss_display_render(pss_ctx* ctx, const float* cols, const float* stats, int W, int64_t n_frames,
                       int rows_max, int64_t first, int64_t step, int64_t n_renders,
                       int guard_zero_range, float* norm, float* minmax) {
    if (!ctx || !cols || !stats || !norm || !minmax || W < 1 || rows_max < 1 || n_frames < 0 || n_renders < 0)
        return PSS_ERR_ARG;
    if (n_renders == 0) return PSS_OK;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t cols_b = (size_t)n_frames * W * 4, st_b = (size_t)n_frames * 16;
    const size_t norm_b = (size_t)n_renders * rows_max * W * 4, mm_b = (size_t)n_renders * 8;
    int rc;
    if ((rc = pss_reserve(ctx, &ctx->d_in, &ctx->d_in_bytes, cols_b))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_aux, &ctx->d_aux_bytes, st_b))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_out, &ctx->d_out_bytes, norm_b))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_aux2, &ctx->d_aux2_bytes, mm_b))) return rc;
    PSS_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, cols, cols_b, cudaMemcpyHostToDevice, ctx->stream));
    PSS_CUDA(ctx, cudaMemcpyAsync(ctx->d_aux, stats, st_b, cudaMemcpyHostToDevice, ctx->stream));
    rc = pss_display_render_dev(ctx, (const float*)ctx->d_in, (const float*)ctx->d_aux, W, n_frames, rows_max, first,
                                step, n_renders, guard_zero_range, (float*)ctx->d_out, (float*)ctx->d_aux2);
    if (rc) return rc;
    PSS_CUDA(ctx, cudaMemcpyAsync(norm, ctx->d_out, norm_b, cudaMemcpyDeviceToHost, ctx->stream));
    PSS_CUDA(ctx, cudaMemcpyAsync(minmax, ctx->d_aux2, mm_b, cudaMemcpyDeviceToHost, ctx->stream));
    PSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PSS_OK;
}

int pss_display_quantise(pss_ctx* ctx, const float* norm, int64_t n, int kind, int H, uint8_t* plane_a,
                         uint8_t* plane_b) {
    if (!ctx || !norm || !plane_a || n < 0 || kind < 0 || kind > 3) return PSS_ERR_ARG;
    if (kind == PSS_QUANT_PERSISTENCE && H < 2) return PSS_ERR_ARG;
    if (n == 0) return PSS_OK;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = pss_reserve(ctx, &ctx->d_in, &ctx->d_in_bytes, (size_t)n * 4))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_out, &ctx->d_out_bytes, (size_t)n * 2))) return rc;
    uint8_t* da = (uint8_t*)ctx->d_out;
    uint8_t* db = plane_b ? da + n : nullptr;
    PSS_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, norm, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    display_quantise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const float*)ctx->d_in, n, kind, H, da, db);
    PSS_LAUNCH_CHECK(ctx);
    PSS_CUDA(ctx, cudaMemcpyAsync(plane_a, da, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (plane_b) PSS_CUDA(ctx, cudaMemcpyAsync(plane_b, db, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    PSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PSS_OK;
}

}  // extern "C"

// numpy's virtual index of the q-th percentile, method 'linear': (n - 1) * quantile with quantile = q / 100,
// previous index = floor, gamma = the remainder (numpy/lib/_function_base_impl.py: _QuantileMethods['linear']).
static void percentile_index(int n, double quantile, unsigned* k, double* frac) {
    const double vi = (double)(n - 1) * quantile;
    double fl = floor(vi);
    if (fl < 0) fl = 0;
    if (fl > n - 1) fl = n - 1;
    *k = (unsigned)fl;
    *frac = vi - fl;
}

template <typename T>
static int spectrum_normalise_impl(pss_ctx* ctx, const T* db, int n_bins, int64_t n_frames, int W, void* cols,
                                   void* range) {
    if (!ctx || !db || !cols || n_bins < 2 || W < 1 || n_frames < 0) return PSS_ERR_ARG;
    if (n_frames == 0) return PSS_OK;
    PSS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t in_b = (size_t)n_frames * n_bins * sizeof(T), out_b = (size_t)n_frames * W * sizeof(T);
    int rc;
    if ((rc = pss_reserve(ctx, &ctx->d_in, &ctx->d_in_bytes, in_b))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_out, &ctx->d_out_bytes, out_b))) return rc;
    if ((rc = pss_reserve(ctx, &ctx->d_aux2, &ctx->d_aux2_bytes, (size_t)n_frames * 2 * sizeof(T)))) return rc;
    PSS_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, db, in_b, cudaMemcpyHostToDevice, ctx->stream));
    const long long grid = n_frames < 4LL * ctx->sm_count ? n_frames : 4LL * ctx->sm_count;
    unsigned k;
    double frac;
    percentile_index(n_bins, 20.0 / 100.0, &k, &frac);
    const double step = W > 1 ? (double)(n_bins - 1) / (double)(W - 1) : 0.0;
    constexpr bool F64 = sizeof(T) == 8;
    spectrum_normalise_kernel<T><<<(unsigned)grid, 512, 0, ctx->stream>>>(
        (const T*)ctx->d_in, n_bins, n_frames, W, k, frac, step, F64 ? nullptr : (float*)ctx->d_out,
        F64 ? nullptr : (float*)ctx->d_aux2, F64 ? (double*)ctx->d_out : nullptr, F64 ? (double*)ctx->d_aux2 : nullptr);
    PSS_LAUNCH_CHECK(ctx);
    PSS_CUDA(ctx, cudaMemcpyAsync(cols, ctx->d_out, out_b, cudaMemcpyDeviceToHost, ctx->stream));
    if (range) PSS_CUDA(ctx, cudaMemcpyAsync(range, ctx->d_aux2, (size_t)n_frames * 2 * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    PSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PSS_OK;
}

extern "C" {

int pss_spectrum_normalise(pss_ctx* ctx, const float* db, int n_bins, int64_t n_frames, int W, float* cols,
                           float* range) {
    return spectrum_normalise_impl<float>(ctx, db, n_bins, n_frames, W, cols, range);
}

int pss_spectrum_normalise_f64(pss_ctx* ctx, const double* db, int n_bins, int64_t n_frames, int W, double* cols,
                               double* range) {
    return spectrum_normalise_impl<double>(ctx, db, n_bins, n_frames, W, cols, range);
}

}  // extern "C"
