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
 *  - one context per thread / per GPU.  Calls on one context are stream-ordered.  Everything the library
 *    allocates (tables, scratch, pipeline streams and events, display rings) belongs to a context and is
 *    released by pss_destroy; two contexts on one GPU share nothing.
 *  - every struct that crosses the ABI starts with `size_t struct_size`, which the caller sets to
 *    sizeof(the struct) as compiled; a mismatch (a binding built against another header revision) is
 *    rejected with PSS_ERR_ARG instead of reading past the caller's struct.
 *  - stream contract of the *_dev variants: inputs must be complete on, and outputs are produced on, the
 *    context's stream (pss_set_stream adopts the caller's; with the context's own stream the caller orders
 *    its producers before the call and calls pss_sync before consuming).
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
 * N must be a power of two, 64 <= N <= 1048576 (every read size of the app, pyspecsdr.py:2236,2415-2422).
 * n_out = N (RAW) or N-4 (SMOOTH_CLAMP).
 * precision = PSS_PREC_FP64 is the parity path.  PSS_PREC_FP32 is an explicit fast mode (RAW epilogue,
 * N <= 8192) that runs window, butterflies and power in float: it does NOT meet the 1e-4 dB bar when
 * a strong carrier is present (errors up to ~0.1-0.5 dB in bins 60 dB under a tone) and is never
 * selected implicitly.
 */
typedef struct {
    size_t struct_size; /* = sizeof(pss_psd_out) */
    float* db;      /* [n_frames][n_out] dB rows, or NULL */
    float* cols;    /* [n_frames][W] rows resampled to W columns (np.interp semantics), or NULL */
    int    W;
    float* stats;   /* [n_frames][4] = max, mean, finite-min, finite-max of the row, or NULL */
    double* moments; /* [n_frames][4] = sum I^2, sum Q^2, sum I*Q, 0 of the frame's samples, or NULL (512 <= N <=
                        65536).  A by-product of the pass that already reads the IQ: feeding it to
                        pss_demod_c64_dev_moments saves the WFM demodulator its own iq_correction pass over
                        the block (signal_processing.py:52-61 need exactly these sums). */
} pss_psd_out;

int pss_psd_c64(pss_ctx* ctx, const float* iq, int N, int64_t n_frames, int window, int epilogue,
                int precision, const pss_psd_out* out);
int pss_psd_c64_dev(pss_ctx* ctx, const float* iq, int N, int64_t n_frames, int window,
                    int epilogue, int precision, const pss_psd_out* out);

/* ------------------------------------------------------------------ scanner
 * Replaces the per-step numerics of the 'c'-key scan loop pyspecsdr.py:2542-2552: un-windowed FFT -> dB ->
 * peak -> number of bins above (peak - rel_db) [use_abs = 0, the inline loop: rel_db = 20], or above an
 * absolute threshold `thr_db` [use_abs = 1: the mask of scan_frequencies pyspecsdr.py:1050-1057; that
 * function itself reads 0.1 s * fs samples per step, which is not a power of two, so only its mask rule
 * is offered here, on power-of-two reads].  512 <= N <= 8192, a power of two.
 * bandwidth = count * fs / N is left to the caller.
 * The count is INTEGER output and is taken in the fp64 power domain (|X|^2 + 1e-10 > p_max * 10^(-rel/10)),
 * which equals the reference's dB comparison except on ties at the 1e-15 level.
 * db_rows (optional) receives the N-bin dB row of every step (the wide-band "stitch").
 */
int pss_scan_c64(pss_ctx* ctx, const float* iq, int N, int64_t n_steps, int use_abs, float thr_db,
                 float* peak_db, int32_t* count_above, float* db_rows);
int pss_scan_c64_dev(pss_ctx* ctx, const float* iq, int N, int64_t n_steps, int use_abs,
                     float thr_db, float* peak_db, int32_t* count_above, float* db_rows);

/* Replaces the numeric part of draw_spectrogram (pyspecsdr.py:418-452): noise floor = 20th percentile
 * of the row (np.percentile, linear interpolation), display range [floor - 0.1*span, max + 0.05*span],
 * clip to [0,1], ** 0.7, W-column np.interp resample.  db [n_frames][n_bins] (host), cols [n_frames][W],
 * range [n_frames][2] = display_min, display_max (may be NULL).
 * (draw_surface_plot :1575-1596 is a display stream of kind PSS_QUANT_SURFACE, see "display accumulate".) */
int pss_spectrum_normalise(pss_ctx* ctx, const float* db, int n_bins, int64_t n_frames, int W, float* cols,
                           float* range);
/* The same on HOST fp64 rows, fp64 throughout in numpy's operation order (percentile virtual index, lerp,
 * clip, pow, np.interp without fused multiply-add): the bar heights int(value * display_height) match the
 * reference's wherever CUDA's and the host libm's pow() agree to the last bit. */
int pss_spectrum_normalise_f64(pss_ctx* ctx, const double* db, int n_bins, int64_t n_frames, int W, double* cols,
                               double* range);

/* ------------------------------------------------------------------ demodulation
 * Replaces demodulate_signal(samples, sample_rate, mode)  signal_processing.py:220-240 and the
 * per-mode chains demodulate_nfm :91-116, demodulate_wfm :119-176 (+ iq_correction :46-80),
 * demodulate_am :179-195, demodulate_ssb :198-217, RAW :237-238.
 *
 * Filter DESIGN stays with the caller (the Python shim designs with scipy exactly as the reference
 * does and hands the coefficients / response tables over once per (mode, sample_rate, block
 * length)); a plan owns their device copies.  Every block is independent (zero initial filter
 * state, per-block peak normalisation), exactly like the reference.
 *
 * PSS_PLAN_DECIM (NFM, WFM): fp32 discriminator -> [65-tap FIR | Butterworth low-pass + de-emphasis]
 *   -> scipy.signal.decimate(q) = 8th-order Chebyshev-I sosfiltfilt (odd extension 27, sosfilt_zi
 *   initial conditions) -> [::q] -> / max|.| * norm.  Evaluated in "chunk-table" form: fp64 tensor
 *   -core (DMMA) products of the discriminator stream with precomputed response tables (a streaming
 *   kernel of independent warps), then parallel prefix scans of the filter states over the chunk
 *   sequence in modal coordinates (independent 2x2 recurrences).
 *   Output: audio[n_frames][n_out][2] float32 (L == R, as in the reference).
 * PSS_PLAN_FIR (USB, LSB): 65-tap FIR on the I channel (the reference's hilbert() is an identity on
 *   the real part, USB == LSB) -> / max|.| * 0.95.          Output: audio[n_frames][N] float32 mono.
 * PSS_PLAN_SOS (AM): |x| - mean -> 5-section Butterworth band-pass (fp64) -> / max|.| * 0.95.
 *                                                            Output: audio[n_frames][N] float32 mono.
 * PSS_PLAN_RAW: real(iq_correction(x)).                      Output: audio[n_frames][N] float32.
 */
enum { PSS_PLAN_DECIM = 0, PSS_PLAN_FIR = 1, PSS_PLAN_SOS = 2, PSS_PLAN_RAW = 3 };

typedef struct {
    size_t struct_size;    /* = sizeof(pss_demod_desc) */
    int kind;              /* PSS_PLAN_* */
    int mode;              /* PSS_MODE_* */
    int N;                 /* IQ samples per block */
    /* --- PSS_PLAN_DECIM (see pyspecsdr_b200/filters.py: build_decim_plan + build_modal_plan).  All tables are
       in MODAL coordinates: the chunk-to-chunk transitions of the forward (pre-filter + Chebyshev forward
       pass, SF states) and backward (Chebyshev reversed pass, SB = 8 states) recurrences are block-diagonal
       with 2x2 real blocks. */
    int q, n_out, lead, SF, SB, n_body, m_tail, tail_start, tail_len;
    float scale, norm;
    int iq_correct;        /* WFM plans: 1 = iq_correction (signal_processing.py:46-80) fused in front of the
                              discriminator, which is what demodulate_signal does (:222-225); 0 = none, which
                              is demodulate_wfm called directly (:119-176) */
    const double* body;    /* [(SF+SB+1)][q+lead] response tables of one body chunk */
    const double* BF;      /* [SF/2][2][2] forward transition blocks */
    const double* BB;      /* [SB/2][2][2] backward transition blocks */
    const double* G;       /* [SB][SF]  backward forcing from the forward state */
    const double* CR;      /* [SF] */
    const double* CB;      /* [SB] */
    double DB;
    const double* head;    /* [(SF+1)][28] */
    const double* tail_T;  /* [(SB+m_tail)][tail_len] */
    const double* tail_M;  /* [(SB+m_tail)][SF] */
    /* --- PSS_PLAN_FIR */
    const double* taps;    /* [n_taps] */
    int n_taps;
    /* --- PSS_PLAN_SOS */
    const double* sos;     /* [n_sections][6] scipy layout b0 b1 b2 a0 a1 a2 */
    int n_sections;
} pss_demod_desc;

typedef struct pss_demod_plan pss_demod_plan;

int  pss_demod_plan_create(pss_ctx* ctx, const pss_demod_desc* desc, pss_demod_plan** out);
void pss_demod_plan_destroy(pss_ctx* ctx, pss_demod_plan* plan);
/* Samples written per block and channel count of a plan's output; IQ samples per block it was built for. */
int  pss_demod_plan_out_len(const pss_demod_plan* plan);
int  pss_demod_plan_block_len(const pss_demod_plan* plan);
int  pss_demod_plan_channels(const pss_demod_plan* plan);
int  pss_demod_c64(pss_ctx* ctx, pss_demod_plan* plan, const float* iq, int64_t n_frames, float* audio);
int  pss_demod_c64_dev(pss_ctx* ctx, pss_demod_plan* plan, const float* iq, int64_t n_frames,
                       float* audio);
/* As pss_demod_c64_dev, with the I/Q second moments of every block supplied as `frames_per_block`
 * consecutive rows of a PSD call's `moments` output over the same IQ (device pointer), each row covering
 * `frame_len` samples; frames_per_block * frame_len must equal the plan's block length (PSS_ERR_ARG
 * otherwise).  Used by WFM plans (iq_correction); ignored by the others. */
int  pss_demod_c64_dev_moments(pss_ctx* ctx, pss_demod_plan* plan, const float* iq, int64_t n_frames,
                               float* audio, const double* moments, int frames_per_block, int frame_len);

/* ------------------------------------------------------------------ display accumulate
 * Replaces the numeric part of draw_waterfall (pyspecsdr.py:1351-1358, 1373-1398),
 * draw_gradient_waterfall (:1649-1696), draw_persistence (:1521-1556) and draw_surface_plot (:1575-1596):
 * a history of the last `rows_max` dB rows (WATERFALL_MAX_LINES = 30, PERSISTENCE_LENGTH = 10, :131, :152),
 * finite min/max over the whole stack, every row resampled to W columns with np.interp semantics,
 * normalised (v - min) / (max - min), then the per-cell int() quantisation.
 *
 * STATEFUL form (what the reference's global WATERFALL_HISTORY / PERSISTENCE_HISTORY lists are): a context
 * owns any number of display streams, each a ring of the last rows_max - 1 rows (their W-column resample and
 * finite min/max) carried on the device across calls, so N calls of one row are bitwise one call of N rows.
 *   pss_display_open(ctx, stream, kind, W, rows_max, H)   create / reset stream `stream` (any int id)
 *        kind = PSS_QUANT_*: WATERFALL divides by max-min as is (a constant stack gives NaN values and plane
 *        cells 255 = nothing drawn: the reference's int(nan) raises there and the frame is not drawn);
 *        GRADIENT / PERSISTENCE / SURFACE replace a zero range by 1 (:1528-1530, :1657-1659, :1577-1579).  H = display_height (PERSISTENCE only).
 *        SURFACE has no history (rows_max must be 1).
 *   pss_display_accumulate_f64   HOST fp64 dB rows [n_rows][n_bins], one render after every row (the call
 *        pattern of the draw_* functions).  Everything is fp64 in numpy's operation order with no fused
 *        multiply-add, so the planes are bit-identical to what the reference draws from the same rows.
 *   pss_display_accumulate_dev   DEVICE float32 rows as the PSD kernel leaves them (`cols`, `stats` of
 *        pss_psd_out, frame order = time order); render r shows the display as it stands after frame
 *        first + r*step of this call.  Asynchronous on the context's stream.
 * Outputs (any may be NULL), newest row first (the order draw_waterfall walks reversed(WATERFALL_HISTORY));
 * rows beyond the history are NaN / 255:
 *   norm / norm64 [n_renders][rows_max][W]   minmax / minmax64 [n_renders][2] stack min, max
 *   n_rows [n_renders]                       rows in the history at each render
 *   plane_a, plane_b [n_renders][rows_max][W] uint8:
 *     WATERFALL    a = glyph level 0..3 ('.', '-', '=', '#'; :1390-1397)   b = int(v*5) colour index (:1388)
 *     GRADIENT     a = int(v*8) index into ' ._-=+*#@' (:1691)             b = int(v*5) (:1695)
 *     PERSISTENCE  a = screen row int((1-v)*(H-1)), 255 if outside [0,H) (:1556-1557)
 *                  b = colour pair of the trace int(1 + 5*(1 - 0.7**(rows_max - i))) (:1544-1545)
 *     SURFACE      a = magnitude int(v*20) (:1593)                          b = int(v*5)
 *   (accumulate_f64 fills norm, norm64, minmax64, n_rows and the planes; accumulate_dev fills norm, minmax,
 *    n_rows and the planes.)
 */
enum { PSS_QUANT_WATERFALL = 0, PSS_QUANT_GRADIENT = 1, PSS_QUANT_PERSISTENCE = 2, PSS_QUANT_SURFACE = 3 };

typedef struct {
    size_t   struct_size;   /* = sizeof(pss_display_out) */
    float*   norm;
    float*   minmax;
    double*  norm64;
    double*  minmax64;
    uint8_t* plane_a;
    uint8_t* plane_b;
    int32_t* n_rows;
} pss_display_out;

int pss_display_open(pss_ctx* ctx, int stream, int kind, int W, int rows_max, int H);
int pss_display_close(pss_ctx* ctx, int stream);
int pss_display_rows(const pss_ctx* ctx, int stream);     /* rows carried in the ring (< rows_max), or < 0 */
int pss_display_accumulate_f64(pss_ctx* ctx, int stream, const double* rows, int n_bins, int64_t n_rows,
                               const pss_display_out* out);
int pss_display_accumulate_dev(pss_ctx* ctx, int stream, const float* cols, const float* stats, int64_t n_frames,
                               int64_t first, int64_t step, int64_t n_renders, const pss_display_out* out);

/* STATELESS form: the history of frame t is frames t, t-1, ..., t-rows_max+1 of one call's `cols` / `stats`
 * arrays.  guard_zero_range: 0 = waterfall, 1 = gradient / persistence. */
int pss_display_render_dev(pss_ctx* ctx, const float* cols, const float* stats, int W, int64_t n_frames,
                           int rows_max, int64_t first, int64_t step, int64_t n_renders,
                           int guard_zero_range, float* norm, float* minmax);
int pss_display_render(pss_ctx* ctx, const float* cols, const float* stats, int W, int64_t n_frames,
                       int rows_max, int64_t first, int64_t step, int64_t n_renders,
                       int guard_zero_range, float* norm, float* minmax);

/* Glyph / colour planes from already-normalised float32 values (same quantisation rules as above). */
int pss_display_quantise(pss_ctx* ctx, const float* norm, int64_t n, int kind, int H, uint8_t* plane_a,
                         uint8_t* plane_b);

/* ------------------------------------------------------------------ whole main-loop iteration, batched
 * One call = what pyspecsdr.py's main loop does per SDR read (pyspecsdr.py:2236-2283 plus the
 * draw_waterfall accumulate), for a batch of reads ("blocks") that are already in HOST memory:
 *   audio  = demodulate_signal(block, fs, mode)            -> plan (PSS_PLAN_*)
 *   rows   = compute_fft(frame) + smoothing + clamp        -> every N_fft-sample frame of the block
 *   header peak/avg, W-column resample, waterfall history normalisation after each block.
 * Host pointers in, host pointers out; the copies are inside the call (pinned memory from
 * pss_host_alloc makes them run at PCIe speed).  Any output pointer may be NULL.
 *   audio  [n_blocks][out_len][channels]          cols   [n_blocks*fpb][W]     (fpb = N_block / N_fft)
 *   stats  [n_blocks*fpb][4]                      db     [n_blocks*fpb][N_fft-4]
 *   norm   [n_blocks][rows_max][W]                minmax [n_blocks][2]
 * The call uses its own copy/compute streams and events (owned by the context) and returns when every
 * output is in the caller's buffers.
 */
typedef struct {
    size_t struct_size;            /* = sizeof(pss_pipeline_io) */
    int N_block, N_fft, W, rows_max;
    pss_demod_plan* plan;          /* NULL = no demodulation; must have been created for N = N_block */
    float *audio, *cols, *stats, *db, *norm, *minmax;
    int display_stream;            /* < 0: the waterfall history starts empty at the first block of the call;
                                      >= 0: a stream from pss_display_open(ctx, id, kind, W, rows_max, H): the
                                      history is carried across calls (one call of N blocks == N calls of one) */
    uint8_t *plane_a, *plane_b;    /* [n_blocks][rows_max][W] glyph / colour planes of the stream's kind
                                      (display_stream >= 0 only), or NULL */
} pss_pipeline_io;

int pss_pipeline_c64(pss_ctx* ctx, const float* iq_host, int64_t n_blocks, const pss_pipeline_io* io);

/* ------------------------------------------------------------------ helpers either side of the path
 * pss_iq_correct_c64   iq_correction(samples)            signal_processing.py:46-80  -> complex64 out[n_frames][N]
 * pss_sosfilt_f32      scipy sosfilt(sos, data), zero state, as used by bandpass_filter
 *                                                        signal_processing.py:34-42  (sos designed by caller)
 * pss_power_c64        measure_signal_power(samples)     signal_processing.py:325-328 -> dB per block
 * pss_audio_to_int16   np.int16(samples * 32767)         audio_processing.py:36-38, io_manager.py:25-26
 *                      on float32 samples (the library's wire format)
 * pss_audio_to_int16_f64  the same on the float64 array the reference holds at that line: one fp64 product
 *                      and the truncating cast, bit-identical to numpy's
 * Host pointers in and out.
 */
int pss_iq_correct_c64(pss_ctx* ctx, const float* iq, int N, int64_t n_frames, float* out);
int pss_sosfilt_f32(pss_ctx* ctx, const float* x, int N, int64_t n_frames, const double* sos, int n_sections,
                    float* y);
int pss_power_c64(pss_ctx* ctx, const float* iq, int N, int64_t n_frames, float* power_db);
int pss_audio_to_int16(pss_ctx* ctx, const float* audio, int64_t n, int16_t* pcm);
int pss_audio_to_int16_f64(pss_ctx* ctx, const double* audio, int64_t n, int16_t* pcm);

/* ------------------------------------------------------------------ signal classifier
 * classify_signal(samples, sample_rate, bandwidth), signal_processing.py:296-322: Welch PSD (scipy.signal.welch
 * defaults at nperseg = 1024: periodic Hann, 50 % overlap, per-segment mean removal, density scaling, FFT-order
 * two-sided output) -> estimate_bandwidth :267-280, estimate_modulation_index :283-293, spectral flatness :304
 * -> decision tree :306-322.  The reference raises NameError at :299 (welch is never imported); this is the
 * computation that line intends (opt-in in the Python shim, the default keeps raising like the reference).
 * features [n_blocks][4] = signal_bw (Hz, FFT-order difference, may be negative like the reference's),
 * modulation_index, spectral_flatness, peak dB of the Welch PSD;  label [n_blocks] = PSS_CLASS_*.
 * N >= 1024 (scipy shrinks nperseg for shorter blocks; not mirrored: PSS_ERR_UNSUPPORTED).
 */
enum { PSS_CLASS_UNKNOWN = 0, PSS_CLASS_FM_BROADCAST = 1, PSS_CLASS_NARROW_FM = 2, PSS_CLASS_AM_BROADCAST = 3,
       PSS_CLASS_SSB = 4, PSS_CLASS_DIGITAL = 5 };
int pss_classify_c64(pss_ctx* ctx, const float* iq, int N, int64_t n_blocks, double fs, double* features,
                     int32_t* label);
int pss_classify_c64_dev(pss_ctx* ctx, const float* iq, int N, int64_t n_blocks, double fs, double* features,
                         int32_t* label);

#ifdef __cplusplus
}
#endif
#endif /* PSS_H */
