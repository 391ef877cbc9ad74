/* libspecgpu -- C ABI of the B200-native spectrogram / denoise / cross-spectrum hot path.
 *
 * Drop-in boundary for the one data-parallel path of PlasmaControl/spectrogram-enhancement
 * (file:line below are into that tree).  The reference has no FFI of its own: its boundary is a set
 * of module-level Python functions; each entry point here is what a ctypes binding of one of
 * those functions calls (see INTEGRATION.md for the stub a maintainer would add).
 *
 * Conventions
 *   - extern "C", POD arguments only, 64-bit sizes, no C++ exception crosses the ABI.
 *   - Every data pointer is a DEVICE pointer owned by the caller (e.g. torch.Tensor.data_ptr());
 *     the library never frees or retains it past the stream work it enqueues.  Host pointers are
 *     named *_host.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls only enqueue
 *     work; they do not synchronise unless documented.
 *   - Return value: SPECGPU_OK (0) or a negative specgpu_status; text via specgpu_last_error().
 *   - A ctx belongs to one device and owns ONE scratch workspace: use it from one host thread and one
 *     stream at a time (calls on different streams through the same ctx are not ordered against each
 *     other and would share the scratch).  Distinct ctxs are independent; create one per (thread, stream).
 *   - Entry points run on the ctx's device and restore the caller's current device before returning.
 *   - Matrices are row-major; `ld*` arguments are leading dimensions in ELEMENTS.
 *   - There is no CPU fallback anywhere behind this header.
 */
#ifndef SPECGPU_H_
#define SPECGPU_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPECGPU_VERSION_MAJOR 0
#define SPECGPU_VERSION_MINOR 1

typedef struct specgpu_ctx specgpu_ctx;
typedef struct specgpu_plan specgpu_plan;

typedef enum specgpu_status {
  SPECGPU_OK = 0,
  SPECGPU_ERR_INVALID_ARG = -1,
  SPECGPU_ERR_UNSUPPORTED_NPERSEG = -2, /* nperseg must be a power of two in [8, 8192] */
  SPECGPU_ERR_CUDA = -3,
  SPECGPU_ERR_NCCL = -4,
  SPECGPU_ERR_WORKSPACE = -5,
  SPECGPU_ERR_UNSUPPORTED_SHAPE = -6
} specgpu_status;

enum { SPECGPU_DETREND_NONE = 0, SPECGPU_DETREND_CONSTANT = 1, SPECGPU_DETREND_LINEAR = 2 };
enum { SPECGPU_SCALING_DENSITY = 0, SPECGPU_SCALING_SPECTRUM = 1 };
enum { SPECGPU_WINDOW_CUSTOM = 0, SPECGPU_WINDOW_HANN = 1, SPECGPU_WINDOW_HAMMING = 2, SPECGPU_WINDOW_BOXCAR = 3 };

/* The reference's `spec_params` dict (spec_denoising/pipeline_data.py:77-84) as a POD.
 * Semantics are scipy.signal.spectrogram's (periodic window, density/spectrum scaling,
 * per-segment detrend, one-sided). */
typedef struct specgpu_stft_params {
  int32_t nperseg;   /* power of two, 8..8192 */
  int32_t noverlap;  /* 0 <= noverlap < nperseg */
  int32_t detrend;   /* SPECGPU_DETREND_* */
  int32_t scaling;   /* SPECGPU_SCALING_* */
  int32_t window;    /* SPECGPU_WINDOW_*; CUSTOM takes window_host in specgpu_plan_create */
  int32_t reserved;
  double fs;         /* sample rate */
  double eps;        /* added before the log (spec_params['eps']) */
} specgpu_stft_params;

/* ---- lifetime ----------------------------------------------------------------------------- */
int specgpu_version(void);                                   /* major*1000 + minor */
int specgpu_init(int device, specgpu_ctx** ctx);
int specgpu_destroy(specgpu_ctx* ctx);
const char* specgpu_last_error(const specgpu_ctx* ctx);      /* "" if none; valid until next call */
/* Grow the ctx-owned device workspace up front (it otherwise grows on demand, which synchronises). */
int specgpu_workspace_reserve(specgpu_ctx* ctx, int64_t bytes);

int specgpu_plan_create(specgpu_ctx* ctx, const specgpu_stft_params* params,
                        const double* window_host /* nperseg values or NULL */, specgpu_plan** plan);
int specgpu_plan_destroy(specgpu_plan* plan);
/* Number of full segments for an n-sample record: (n - noverlap) / (nperseg - noverlap), 0 if n < nperseg
 * (scipy/signal/_spectral_py.py `_fft_helper`; bit-exact integer contract). */
int64_t specgpu_plan_num_segments(const specgpu_plan* plan, int64_t n);
int32_t specgpu_plan_num_freqs(const specgpu_plan* plan);    /* nperseg/2 + 1 */
/* f[k] = k*fs/nperseg (k < nfreq), t[j] = (j*hop + nperseg/2)/fs (j < nseg), in double, on the host. */
int specgpu_plan_axes(const specgpu_plan* plan, int64_t n, double* f_host, double* t_host);

/* ---- spectrogram front-end ---------------------------------------------------------------- */
/* scipy.signal.spectrogram(mode='psd') of B signals: x[B][ldx] (first n valid) -> Sxx[B][nfreq][ldt]
 * (first nseg columns valid).  Replaces the call at spec_denoising/pipeline_data.py:32. */
int specgpu_spectrogram(specgpu_ctx* ctx, const specgpu_plan* plan, const float* x, int64_t B, int64_t n,
                        int64_t ldx, float* Sxx, int64_t ldt, void* stream);

/* Body of `specgr` after the pickle slice (spec_denoising/pipeline_data.py:32-35; BES twin
 * denoising_by_svd.ipynb:52-62): spectrogram -> log(Sxx+eps) -> global min-max over all nfreq rows
 * (per signal) -> drop the Nyquist row.  S[B][nfreq-1][ldt]; minmax[B][2] (optional, may be NULL)
 * receives the per-signal (min, max) of the log image. */
int specgpu_specgr(specgpu_ctx* ctx, const specgpu_plan* plan, const float* x, int64_t B, int64_t n,
                   int64_t ldx, float* S, int64_t ldt, float* minmax, void* stream);

/* scipy.signal.stft (one-sided, real input): Z[B][nfreq][ldt] interleaved complex64, scaled by
 * sqrt(scale).  boundary_zeros: pad nperseg/2 zeros both sides; padded: zero-extend to a whole number
 * of hops.  The number of columns is returned by specgpu_stft_num_segments. */
int64_t specgpu_stft_num_segments(const specgpu_plan* plan, int64_t n, int boundary_zeros, int padded);
int specgpu_stft(specgpu_ctx* ctx, const specgpu_plan* plan, const float* x, int64_t B, int64_t n, int64_t ldx,
                 int boundary_zeros, int padded, float* Z /* float2 */, int64_t ldt, void* stream);

/* ---- array helpers of the reference ------------------------------------------------------- */
/* rescale(data) = (data-min)/(max-min) over the whole array (pipeline_data.py:43-44), per batch item:
 * src/dst [B][rows][ld], cols valid.  In-place allowed. */
int specgpu_rescale(specgpu_ctx* ctx, const float* src, int64_t B, int64_t rows, int64_t cols, int64_t ld,
                    float* dst, void* stream);
/* norm(data) = (data-mean)/std, population std (pipeline_data.py:38-41). */
int specgpu_norm(specgpu_ctx* ctx, const float* src, int64_t B, int64_t rows, int64_t cols, int64_t ld,
                 float* dst, void* stream);

/* clip: dst = src with negatives set to 0 (denoising_by_svd.ipynb:280-281), n contiguous elements; in place allowed.
 * (specgpu_svd_denoise / specgpu_pipeline fuse this through their `clip` argument.) */
int specgpu_clip(specgpu_ctx* ctx, const float* src, int64_t n, float* dst, void* stream);

/* quantfilt (pipeline_data.py:46-49): per column, q = np.quantile(src[:,j], thr) over the `rows`
 * axis with numpy's float32 'linear' arithmetic (bit-exact), dst = src < q ? 0 : src.
 * dst (optional) has the row pitch ld of src; thr_out[B][cols] and mask[B][rows][ld] (uint8, 1 = kept) are optional
 * (NULL to skip). rows <= 1024. */
int specgpu_quantfilt(specgpu_ctx* ctx, const float* src, int64_t B, int64_t rows, int64_t cols, int64_t ld,
                      float thr, float* dst, float* thr_out, uint8_t* mask, void* stream);

/* ---- cv2 image chain (spec_denoising/pipeline_data.py:52-72) -------------------------------------------------------- */
/* gaussblr (:52-55): (rescale(src)*255).astype('uint8') -> cv2.GaussianBlur(ksize = (kw, kh), sigma 0; kw taps along the
 * contiguous/time axis, kh along rows; OpenCV's CV_8U fixed-point path, BORDER_REFLECT_101) -> rescale -> float64.
 * src is float32 (in_f64 = 0) or float64; u8_out (optional) receives the blurred uint8 image [B][rows][cols] (bit-exact). */
int specgpu_gaussblr(specgpu_ctx* ctx, const void* src, int32_t in_f64, int64_t B, int64_t rows, int64_t cols, int64_t ld,
                     int32_t kw, int32_t kh, double* dst, int64_t ldo, uint8_t* u8_out, void* stream);
/* meansub (:58-61): rescale(|src - mean over the contiguous axis of each row|), float64 in and out; the row means are
 * np.mean's (pairwise summation reproduced), so the result is bit-identical to numpy's.  cols <= 120000. */
int specgpu_meansub(specgpu_ctx* ctx, const double* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, double* dst,
                    int64_t ldo, void* stream);
/* morph (:64-72): uint8-quantise -> MORPH_CLOSE rect(4,4) -> MORPH_OPEN rect(3,1) -> rescale -> float64; u8_out (optional)
 * receives the uint8 mask before the final rescale (bit-exact). */
int specgpu_morph(specgpu_ctx* ctx, const void* src, int32_t in_f64, int64_t B, int64_t rows, int64_t cols, int64_t ld,
                  double* dst, int64_t ldo, uint8_t* u8_out, void* stream);
/* The whole denoising body of the reference's main loop (pipeline_data.py:101-110) on a [B][rows][cols] float32 stack:
 * quantfilt(thr) -> gaussblr((kw, kh)) -> meansub -> morph -> meansub, float64 out.  Identical, bit for bit, to chaining
 * the five calls above; between the stages only uint8 planes (and their min / max) exist on the device.  src rows have
 * pitch ld, dst rows pitch ldo; rows <= 1024 (quantfilt's limit), cols <= 120000 (meansub's). */
int specgpu_filter_chain(specgpu_ctx* ctx, const float* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, float thr,
                         int32_t kw, int32_t kh, double* dst, int64_t ldo, void* stream);

/* ---- SVD denoise (denoising_by_svd.ipynb:155-229, 280-281) --------------------------------- */
/* out = U[:,a:b] diag(s[a:b]) Vh[a:b,:] of each S[b] (rows x cols, rows <= cols, rows <= 512), where
 *   use_optimal == 0: a = start, b = stop (pass start = 1, stop = rows for the reference defaults)
 *   use_optimal != 0: tau = omega(rows/cols)*median(s); num_sing = #(s > tau); a = 0, b = num_sing-1
 * then the reference's clamps and Python slice semantics (negative b counts from the end).
 * clip != 0 additionally applies out[out<0] = 0 (notebook lines 280-281).
 * s_out[B][rows] (descending singular values, optional) and info[B][4] = {a, b, num_sing or -1, status}
 * (optional; status 0 = ok, 1 = eigen-iteration hit its cap) are written on the device.
 * mode: 0 = auto (power iteration when only the leading component is removed and nothing else is asked,
 * full Jacobi otherwise), 1 = force the full eigen-decomposition. */
int specgpu_svd_denoise(specgpu_ctx* ctx, const float* S, int64_t B, int64_t rows, int64_t cols, int64_t ld,
                        int32_t start, int32_t stop, int32_t use_optimal, int32_t clip, int32_t mode,
                        float* out, int64_t ldo, float* s_out, int32_t* info, void* stream);
/* computeSignal (denoising_by_svd.ipynb:161-186): sum_{idx=1}^{2*num_sing-1} s u v^T, float64 output
 * [B][rows][ldo]; info as above with {1, 2*num_sing, num_sing, status}. */
int specgpu_compute_signal(specgpu_ctx* ctx, const float* S, int64_t B, int64_t rows, int64_t cols, int64_t ld,
                           double* out, int64_t ldo, float* s_out, int32_t* info, void* stream);

/* ---- VAE tile export (VAE/manual_scan.py:28-54) ---------------------------------------------- */
/* patch: out[(i*ntiles + x)][r][c] = src[i][r][x*tile_w + c]; out_f64 selects float64 (the reference's
 * np.empty default) or float32 output.  unpatch is the inverse into dst[i][r][ld]. */
int specgpu_patch(specgpu_ctx* ctx, const float* src, int64_t n, int64_t rows, int64_t ld, int32_t tile_w,
                  int32_t ntiles, void* out, int32_t out_f64, void* stream);
int specgpu_unpatch(specgpu_ctx* ctx, const void* tiles, int32_t in_f64, int64_t n, int64_t rows, int32_t tile_w,
                    int32_t ntiles, void* dst, int32_t out_f64, int64_t ld, void* stream);

/* ---- cross-power spectrum (interferometer/crosspowerspec.py:39; scipy.signal.csd, mean) ------ */
/* Stage 1: unscaled one-sided spectra of every full segment, X[C][nseg][ldf] interleaved complex64
 * (ldf >= nfreq).  Stage 2: P[i][j][f] = mean_t conj(X_i) X_j * scale, one-sided doubled, for
 * i in [i0, i0+ni), all j < C; P is [ni][C][nfreq] complex64.  specgpu_csd_allpairs runs both on one
 * device (P[C][C][nfreq]); the two stages are exported separately so that a channel-block-sharded
 * caller can all-gather X between them. */
int specgpu_csd_spectra(specgpu_ctx* ctx, const specgpu_plan* plan, const float* x, int64_t C, int64_t n,
                        int64_t ldx, float* X, int64_t ldf, void* stream);
int specgpu_csd_pairs(specgpu_ctx* ctx, const specgpu_plan* plan, const float* X, int64_t C, int64_t nseg,
                      int64_t ldf, int64_t i0, int64_t ni, float* P, void* stream);
/* One block of `nseg` consecutive segments of an average over `nseg_total`: P (+)= sum_t conj(X_i) X_j * scale /
 * nseg_total.  Lets a channel-sharded caller exchange and consume the spectra block by block (the all-gather of
 * block n+1 overlaps the pair products of block n); accumulate = 0 on the first block. */
int specgpu_csd_pairs_block(specgpu_ctx* ctx, const specgpu_plan* plan, const float* X, int64_t C, int64_t nseg,
                            int64_t nseg_total, int64_t ldf, int64_t i0, int64_t ni, int32_t accumulate, float* P,
                            void* stream);
/* specgpu_csd_spectra with the output cut into frequency blocks of block_w bins, BLOCK-MAJOR:
 * X[nblocks][C][nseg][block_ld] (nblocks = ceil(nfreq / block_w), block_ld >= block_w), bin k of a segment at block
 * k / block_w, column k % block_w.  Plane h is exactly what a frequency-block all-to-all sends to rank h. */
int specgpu_csd_spectra_blocked(specgpu_ctx* ctx, const specgpu_plan* plan, const float* x, int64_t C, int64_t n,
                                int64_t ldx, float* X, int32_t block_w, int32_t block_ld, void* stream);
/* All pairs over a BLOCK OF FREQUENCY BINS: X[C][nseg][ldf] holds bins f0 .. f0 + nf - 1 of every channel's one-sided
 * spectra in columns 0 .. nf - 1 (the layout a frequency-block all-to-all of specgpu_csd_spectra outputs produces on
 * each rank; pair products are independent per bin); P[C][C][nf] (+)= sum_t conj(X_i) X_j * scale / nseg_total with the
 * one-sided doubling decided by the global bin index f0 + f. */
int specgpu_csd_pairs_bins(specgpu_ctx* ctx, const specgpu_plan* plan, const float* X, int64_t C, int64_t nseg,
                           int64_t nseg_total, int64_t ldf, int64_t f0, int64_t nf, int32_t accumulate, float* P,
                           void* stream);
/* Time-resolved cross-power amplitude for the `ampsp[n_time, n_freq]` image of interferometer/crosspowerspec.py:39-50:
 * amp[k][f] = | mean over segments [k*seg_stride, k*seg_stride + navg) of conj(X_i) X_j | * scale (one-sided doubled),
 * from the spectra of specgpu_csd_spectra.  amp is [nframes][nfreq] float32. */
int specgpu_csd_frames(specgpu_ctx* ctx, const specgpu_plan* plan, const float* X, int64_t C, int64_t nseg, int64_t ldf,
                       int64_t i, int64_t j, int64_t seg_stride, int32_t navg, int64_t nframes, float* amp, void* stream);
int specgpu_csd_allpairs(specgpu_ctx* ctx, const specgpu_plan* plan, const float* x, int64_t C, int64_t n,
                         int64_t ldx, float* P, void* stream);

/* ---- the whole path for one batch of channels (bench / production entry) -------------------- */
/* specgr -> denoiseSignal(default: drop the leading component) -> clip, for x[B][ldx]
 * (the loop body of pipeline_data.py:92-97 + denoising_by_svd.ipynb:263, 280-281):
 * S[B][nfreq-1][ldt] (normalised spectrogram) and D[B][nfreq-1][ldt] (denoised).
 * flags: SPECGPU_PIPE_CLIP      clip D at 0 (hacked[hacked<0] = 0);
 *        SPECGPU_PIPE_FALLBACK  channels whose leading singular pair did not converge in the power iteration
 *                               (a (nearly) degenerate pair, an iterate in the null space) are redone IN THE STREAM by
 *                               the full float64 eigensolver -- three launches that return at once when every channel
 *                               converged.  Always on when `info` is NULL (the caller could not see the status).
 * Optional tiles (NULL to skip): float32 [B*ntiles][nfreq-1][tile_w] cut from D.
 *        SPECGPU_PIPE_STATIC_TILES  scheduling hint, results do not depend on it: the STFT's persistent CTAs walk their
 *                               tiles with a fixed stride instead of taking them from a counter.  The counter balances
 *                               the CTAs of a call that runs alone (-2 % per shot); with several shots in flight on
 *                               several streams the staggered finish of the fixed walk overlaps better (api.ShotStreams
 *                               sets it).
 * info[B][4] (optional, device) = {1, nfreq-1, -1, status}; status 1 (only possible without SPECGPU_PIPE_FALLBACK) =
 * the leading pair of that channel did not converge: D of that channel must be recomputed with
 * specgpu_svd_denoise(S, mode = 1). */
enum { SPECGPU_PIPE_CLIP = 1, SPECGPU_PIPE_FALLBACK = 2, SPECGPU_PIPE_STATIC_TILES = 4 };
int specgpu_pipeline(specgpu_ctx* ctx, const specgpu_plan* plan, const float* x, int64_t B, int64_t n, int64_t ldx,
                     float* S, float* D, int64_t ldt, int32_t flags, float* tiles, int32_t tile_w, int32_t ntiles,
                     int32_t* info, void* stream);

/* Strided copy of `nrows` rows of `row_bytes` bytes between any two of {pinned host, device} buffers with independent row
 * pitches (cudaMemcpy2DAsync, enqueued on `stream`): moves the row-pitched images the pipeline works on to and from the
 * dense host arrays of the reference's interface without a staging pass. */
int specgpu_copy_rows(specgpu_ctx* ctx, void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t row_bytes,
                      int64_t nrows, void* stream);

/* specgpu_pipeline can process the batch in groups of `channels` channels, alternating between two library-owned
 * streams that are forked from and joined to the caller's stream.  0 (default) or a value >= B runs the batch as one
 * group on the caller's stream, which is what measures fastest on B200; the knob exists for batches far
 * larger than the L2 cache and for experiments.  Results do not depend on it. */
int specgpu_set_pipeline_group(specgpu_ctx* ctx, int32_t channels);

/* Interlock between two contexts that run specgpu_pipeline on two streams (shots in flight, see api.ShotStreams).
 * `wait_event` (a cudaEvent_t, may be NULL): the call makes its stream wait for it before its first kernel (the STFT).
 * `record_event` (may be NULL): the call records it on its stream right before the projection (its last, HBM-bound
 * kernel).  With context A recording what context B waits for and vice versa, the instruction-bound STFT of one shot
 * starts exactly when the memory-bound projection of the other does, instead of whenever the streams happen to drift.
 * The events stay owned by the caller and must outlive the calls; (NULL, NULL) turns the interlock off. */
int specgpu_set_pipeline_interlock(specgpu_ctx* ctx, void* wait_event, void* record_event);

/* Cap of the power iteration that finds the leading singular pair on the default denoise route (0 restores the
 * default, 200).  A channel that does not converge within the cap is flagged (info[b][3] = 1) and, where the fallback is
 * on, redone by the full float64 eigensolver; a cap of 1 therefore sends every channel down the fallback route, which
 * is how the tests exercise it on spectrogram images (their leading pair is never degenerate by itself). */
int specgpu_set_power_iterations(specgpu_ctx* ctx, int32_t max_iter);

/* Number of kernel launches the library has enqueued on this ctx since creation (for bench.py's
 * gpu_launches claim). */
int64_t specgpu_launch_count(const specgpu_ctx* ctx);

/* Optional per-kernel timing for bench.py's roofline line: while enabled, every launch group the
 * library enqueues is bracketed by CUDA events on the caller's stream.  specgpu_profile_count
 * synchronises on the recorded events and returns the number of distinct kernel names seen since
 * the last enable; name / total milliseconds / number of timed launches are then read per index. */
int specgpu_profile_enable(specgpu_ctx* ctx, int enable);
int specgpu_profile_count(specgpu_ctx* ctx);
const char* specgpu_profile_name(const specgpu_ctx* ctx, int i);
double specgpu_profile_ms(const specgpu_ctx* ctx, int i);
int64_t specgpu_profile_calls(const specgpu_ctx* ctx, int i);

#ifdef __cplusplus
}
#endif
#endif /* SPECGPU_H_ */
