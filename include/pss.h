/*
 * pss.h — C ABI of the B200-native IQ-processing core (libpss.so).
 *
 * This is the drop-in boundary for PySpecSDR's DSP hot path.  The reference has no FFI of
 * its own (it is pure Python over numpy/scipy); the functions below are what a binding of
 * that path binds, one entry point per reference function / per inlined numeric block.
 * Each declaration cites the reference code it replaces (paths are into the PySpecSDR tree).
 * The ctypes stub a maintainer adds on the reference side is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C types only; IQ is interleaved float32 (I,Q) = numpy complex64 = SoapySDR CF32
 *    (pyspecsdr.py:1870,1885-1891).  "frames" are independent blocks, contiguous, n_frames*N.
 *  - every call returns PSS_OK (0) or a negative pss_status; nothing throws, nothing exits.
 *  - `*_dev` variants take DEVICE pointers and only enqueue work on the context's stream
 *    (asynchronous); the un-suffixed variants take HOST pointers, copy in, run, copy out and
 *    return when the result is in the caller's buffer.
 *  - the caller owns every buffer; the library owns only the context (stream, window/twiddle
 *    tables, filter tables, scratch, display history rings).
 *  - one context per thread / per GPU.  Calls on one context are stream-ordered.
 */
#ifndef PSS_H
#define PSS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pss_ctx pss_ctx;

typedef enum {
    PSS_OK = 0,
    PSS_ERR_ARG = -1,          /* bad argument (NULL, size, enum) */
    PSS_ERR_CUDA = -2,         /* a CUDA runtime call failed; see pss_last_error */
    PSS_ERR_NOMEM = -3,
    PSS_ERR_UNSUPPORTED = -4,  /* valid request this build does not implement (e.g. N not 2^k) */
    PSS_ERR_NODEVICE = -5      /* no CUDA device / not sm_100 */
} pss_status;

enum { PSS_WINDOW_NONE = 0, PSS_WINDOW_HAMMING = 1, PSS_WINDOW_HANN = 2 };
enum { PSS_EPI_RAW = 0, PSS_EPI_SMOOTH_CLAMP = 1 };
enum { PSS_PREC_FP64 = 0, PSS_PREC_FP32 = 1 };
enum { PSS_MODE_NFM = 0, PSS_MODE_WFM = 1, PSS_MODE_AM = 2, PSS_MODE_USB = 3, PSS_MODE_LSB = 4,
       PSS_MODE_RAW = 5 };
enum { PSS_DISPLAY_WATERFALL = 0, PSS_DISPLAY_PERSISTENCE = 1 };

/* ------------------------------------------------------------------ lifetime / plumbing */
int         pss_init(int device, pss_ctx** out);
void        pss_destroy(pss_ctx* ctx);
const char* pss_strerror(int status);
const char* pss_last_error(const pss_ctx* ctx);      /* text of the last CUDA failure */
int         pss_version(void);
/* Adopt a caller-owned cudaStream_t (e.g. torch's current stream) for all *_dev calls; NULL is
 * the CUDA default stream.  pss_use_own_stream() goes back to the context's private stream. */
int         pss_set_stream(pss_ctx* ctx, void* cuda_stream);
int         pss_use_own_stream(pss_ctx* ctx);
int         pss_sync(pss_ctx* ctx);
/* Number of kernels this library has launched on this context (bench.py's gpu_launches). */
int64_t     pss_kernel_launches(const pss_ctx* ctx);
/* Page-locked host memory for the host-pointer variants (any host pointer works; pinned is
 * what makes the copies run at PCIe speed). */
void*       pss_host_alloc(size_t bytes);
void        pss_host_free(void* p);

/* ------------------------------------------------------------------ PSD
 * Replaces compute_fft(samples)              signal_processing.py:243-264  (window = HAMMING)
 *      and the scanner's un-windowed spectrum  pyspecsdr.py:2542-2543, 1050-1051 (window = NONE)
 *      + optionally the main-loop epilogue      pyspecsdr.py:2278-2283  (epilogue = SMOOTH_CLAMP:
 *        5-bin 'valid' moving average -> N-4 bins, clamp below median-10 dB)
 *      + header peak/avg                        pyspecsdr.py:388-389
 *      + the W-column np.interp resample every draw_* does (pyspecsdr.py:448-452, 1378-1382, ...)
 * One fused kernel: window -> radix-16 Stockham FFT (fp64) -> fftshift -> 10*log10(|X|^2+1e-10).
 * N must be a power of two, 64 <= N <= 2^20.  n_out = N (RAW) or N-4 (SMOOTH_CLAMP).
 */
typedef struct {
    float* db;      /* [n_frames][n_out] dB rows, or NULL */
    float* cols;    /* [n_frames][W] rows resampled to W columns (np.interp semantics), or NULL */
    int    W;
    float* stats;   /* [n_frames][4] = max, mean, finite-min, finite-max of the row, or NULL */
} pss_psd_out;

int pss_psd_c64(pss_ctx* ctx, const float* iq, int N, int64_t n_frames, int window, int epilogue,
                int precision, const pss_psd_out* out);
int pss_psd_c64_dev(pss_ctx* ctx, const float* iq, int N, int64_t n_frames, int window,
                    int epilogue, int precision, const pss_psd_out* out);

/* ------------------------------------------------------------------ scanner
 * Replaces the per-step numerics of the 'c'-key scan loop pyspecsdr.py:2542-2552 and of
 * scan_frequencies pyspecsdr.py:1050-1057: un-windowed FFT -> dB -> peak -> number of bins above
 * (peak - rel_db) [use_abs = 0, the inline loop: rel_db = 20] or above abs threshold `thr_db`
 * [use_abs = 1, scan_frequencies].  bandwidth = count * fs / N is left to the caller.
 * db_rows (optional) receives the N-bin dB row of every step (the wide-band "stitch").
 */
int pss_scan_c64(pss_ctx* ctx, const float* iq, int N, int64_t n_steps, int use_abs, float thr_db,
                 float* peak_db, int32_t* count_above, float* db_rows);
int pss_scan_c64_dev(pss_ctx* ctx, const float* iq, int N, int64_t n_steps, int use_abs,
                     float thr_db, float* peak_db, int32_t* count_above, float* db_rows);

#ifdef __cplusplus
}
#endif
#endif /* PSS_H */
