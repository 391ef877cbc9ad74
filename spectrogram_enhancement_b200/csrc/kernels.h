// Internal launcher interface between the C ABI (specgpu.cu) and the kernel translation units.
#pragma once
#include "common.cuh"
#include "tensor_map.h"

namespace specgpu {

enum { STFT_MODE_PSD = 0, STFT_MODE_LOGPSD = 1, STFT_MODE_COMPLEX = 2, STFT_MODE_SPECTRA = 3 };

struct StftArgs {
  const float* x;        // [B][ldx]
  int64_t n;             // valid samples per signal
  int64_t ldx;
  int64_t first_start;   // sample index of segment 0 (negative with boundary='zeros')
  int64_t nseg;
  int hop;
  int detrend;
  int vec_ok;            // float2 loads are aligned
  float scale;           // psd scale (sqrt(scale) for STFT_MODE_COMPLEX)
  float eps;
  const float* window;   // [nperseg]
  const float2* twM;     // per-pass twiddle tables of the M-point transform (fft.cuh: fft_twiddle_count entries)
  const float2* twN;     // exp(-2 pi i k / N), k <= M/2
  void* out;
  int64_t ld_out;
  MinMaxWord* minmax;    // [B][2] generation-tagged min / max words (STFT_MODE_LOGPSD), see common.cuh
  unsigned minmax_gen;   // generation of this call
  int64_t tiles_per_signal, ntiles;   // filled by the launcher (persistent tile loop)
  // filled by the launcher:
  int stage_in;          // the tile's sample span goes through shared memory (bulk copy for interior tiles)
  int bulk_ok;           // base pointer / hop / first_start allow 16-byte aligned bulk copies
  int span;              // samples in a tile's span: (TT-1)*hop + nperseg
  int tma_out;           // whole boxes of the output tile leave through a TMA tensor store
  int tma_rows, tma_nbox;   // rows per box, boxes per tile (rows beyond tma_rows*tma_nbox use plain stores)
  int flow;              // barrier-free tile loop (stft.cu): staged bulk input + whole tile through the tensor store
  unsigned* dyn;         // set by the caller (nullptr by default): {next tile counter, CTAs that have left}, both zero between
                         // launches; the barrier-free path then takes its tiles from the counter (one launch at a time per pair)
  // set by the caller (0 by default): fraction of the output's cache lines to keep in L2 (evict_last) because a later
  // kernel of the same call reads the output back; the rest streams out (evict_first).  0: no preference.
  float l2_pin;
  // STFT_MODE_SPECTRA only: write bin k to element (k / fblock_w) * fblock_stride + k % fblock_w of the segment's row
  // (frequency blocks for the all-to-all of a frequency-sharded CSD; fblock_stride = the block pitch inside a row, or
  // the size of a whole [C][nseg][ld_out] plane for block-major output); fblock_w == 0: bin k to column k.
  int fblock_w;
  int64_t fblock_stride;
};

// stft.cu
int launch_stft(int log2n, int mode, const StftArgs& a, int64_t B, cudaStream_t stream);
int launch_lognorm(float* S, int64_t B, int64_t rows, int64_t cols, int64_t ld, const MinMaxWord* mm, float* mm_out,
                   cudaStream_t stream);

// stft_gram.cu: the nperseg-512 log-PSD STFT that also accumulates the Gram partials of the raw image (device build only)
bool stft_gram_supported(int log2n);
int launch_stft_gram(const StftArgs& a, int64_t B, float* partial_ws, int num_sms, int64_t* nchunk, int64_t* per,
                     cudaStream_t stream);

// elementwise.cu
int launch_rescale(const float* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, float* dst, unsigned* mm_ws,
                   cudaStream_t stream);
int launch_clip_neg(const float* src, int64_t n, float* dst, cudaStream_t stream);
int norm_num_parts(int64_t rows, int64_t cols);
int launch_norm(const float* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, float* dst, double* sums_ws,
                cudaStream_t stream);
int launch_patch(const float* src, int64_t n, int64_t rows, int64_t ld, int tile_w, int ntiles, void* out, int out_f64,
                 cudaStream_t stream);
int launch_unpatch(const void* tiles, int in_f64, int64_t n, int64_t rows, int tile_w, int ntiles, void* dst,
                   int out_f64, int64_t ld, cudaStream_t stream);

// quantile.cu
int launch_quantfilt(const float* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, int lo, float g, float* dst,
                     float* thr_out, uint8_t* mask, cudaStream_t stream);

// svd.cu
int launch_gram_simt(const float* S, int64_t B, int64_t rows, int64_t cols, int64_t ld, void* G, int g_f64, cudaStream_t stream,
                     const MinMaxWord* minmax = nullptr, const int32_t* only_flagged = nullptr);
int launch_eig_power(const float* G, int64_t B, int n, int max_iter /* <= 0: default */, float* U, float* lam, int32_t* plan, cudaStream_t stream);
// split-K reduce of launch_gram_tc's partials fused with the leading-pair power iteration (n in {128, 256}); G itself is
// never written.  (nchunk, per) = gram_tc_geometry of the launch that produced the partials.
// raw_minmax != nullptr: the partials are of the raw log image (launch_gram_tma) and G of the normalised image is formed as
// G_raw - m (r 1^T + 1 r^T) + m^2 cols 1 1^T with m = min of the matrix and r its row sums (scale 1 / (max - min)^2 dropped:
// eigenvectors do not care); raw_minmax == nullptr: the partials already are of the operands to decompose.
int launch_gram_eig(const float* partial, int64_t nchunk, int64_t per, int64_t B, int n, int max_iter, float* U, float* lam,
                    int32_t* plan, cudaStream_t stream, const MinMaxWord* raw_minmax = nullptr, int64_t cols = 0,
                    int per_matrix = 0 /* partials written by launch_gram_tma */, int32_t* flagged = nullptr);
size_t jacobi_workspace_bytes(int64_t B, int n);
bool eig_jacobi_f64_supported(int n);
int launch_eig_jacobi(const void* G, int g_f64, int64_t B, int n, int skip_converged, float* U, float* lam, int32_t* plan, void* ws,
                      cudaStream_t stream);
// kind 0: explicit (start, stop); 1: use_optimal; 2: computeSignal.  plan[b] = {a, e, num_sing, status}
int launch_svd_plan(const float* lam, int64_t B, int n, int kind, int start, int stop, double omega, int32_t* plan,
                    float* s_out, cudaStream_t stream);
// Leading-component removal fused with the min-max normalisation: L (log image) -> S = (L-min)/(max-min) and
// D = S - u0 (u0^T S) [clipped]; S may alias L.  minmax == nullptr: L is already normalised (S not written if null).
// Optional second destination of the rank-1 projection: the VAE tiles of the denoised image (VAE/manual_scan.py:28-36,
// float32 [B * ntiles][rows][tile_w]; column c of matrix b goes to tile c / tile_w when that is < ntiles) written in the
// same pass, instead of a separate tile cut that reads D back.
struct R1Tiles {
  float* ptr = nullptr;
  int tile_w = 0, ntiles = 0;
  // optional: the kernel also copies plan[b][0..3] -> info[b][0..3] (a cudaMemcpyAsync after the projection would break the
  // programmatic launch chain into the next call's STFT: 4.6 us per shot, and most of what two shots in flight can overlap)
  const int32_t* plan = nullptr;
  int32_t* info = nullptr;
};
int launch_svd_rank1(const float* L, int64_t B, int rows, int64_t cols, int64_t ld, const MinMaxWord* minmax, const float* U,
                     int clip, float* S, float* D, int64_t ldo, cudaStream_t stream, int stream_out = 0,
                     const int32_t* only_flagged = nullptr /* [B]: skip matrices whose entry is 0 */, R1Tiles tiles = R1Tiles{});
int launch_svd_project(const float* S, int64_t B, int rows, int64_t cols, int64_t ld, const float* U, const int32_t* plan,
                       int clip, void* out, int out_f64, int64_t ldo, cudaStream_t stream);

// eig_tridiag.cu: values-first float64 eigensolver (Householder tridiagonalisation + bisection + inverse iteration) for the
// use_optimal / computeSignal modes; matrices it cannot serve are flagged in plan[b][3] for the Jacobi solver.
bool eig_tridiag_supported(int n);
size_t eig_tridiag_workspace_bytes(int64_t B, int n);
int launch_eig_tridiag_values(const double* G, double* W /* [B][n][n]: receives the reflectors */, int64_t B, int n, float* lam, void* ws,
                              cudaStream_t stream);
int launch_eig_tridiag_vectors(const double* W, const double* G, int64_t B, int n, int32_t* plan, float* U, void* ws,
                               cudaStream_t stream);

// gram_tc.cu
bool gram_tc_supported(int64_t rows);
size_t gram_tc_workspace_bytes(int64_t B, int64_t rows);
// G == nullptr: leave the split-K partials in partial_ws (launch_gram_eig consumes them) and skip the reduce launch.
int launch_gram_tc(const float* S, int64_t B, int64_t rows, int64_t cols, int64_t ld, const MinMaxWord* minmax,
                   float* partial_ws, float* G, int num_sms, cudaStream_t stream, float l2_pin = 0.f);
void gram_tc_geometry(int64_t B, int64_t cols, int num_sms, int tma /* launch_gram_tma wrote the partials */, int64_t* nchunk, int64_t* per);
// TMA-fed Gram partials of the RAW (un-normalised) operands plus their row sums (row-pitched S only): 0 = launched,
// 1 = layout not describable by a tensor map (use launch_gram_tc), else a CUDA error.
int launch_gram_tma(const float* S, int64_t B, int64_t rows, int64_t cols, int64_t ld, float* partial_ws, int num_sms,
                    cudaStream_t stream, float l2_pin = 0.f);

// imgchain.cu (the cv2 chain: gaussblr / meansub / morph)
size_t imgchain_workspace_bytes(int64_t B, int64_t rows, int64_t cols);
int launch_gaussblr(const void* src, int in_f64, int64_t B, int64_t rows, int64_t cols, int64_t ld, const uint16_t* taps_host,
                    int kw, int kh, void* ws, double* dst, int64_t ldo, uint8_t* u8_out, cudaStream_t st);
int launch_filter_tail(const float* q, int64_t B, int64_t rows, int64_t cols, int64_t ld, const uint16_t* taps_host, int kw,
                       int kh, void* ws, double* dst, int64_t ldo, cudaStream_t st);
int launch_meansub(const double* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, void* ws, double* dst, int64_t ldo,
                   cudaStream_t st);
int launch_morph(const void* src, int in_f64, int64_t B, int64_t rows, int64_t cols, int64_t ld, void* ws, double* dst,
                 int64_t ldo, uint8_t* u8_out, cudaStream_t st);

// csd.cu
// X holds bins f0 .. f0 + nfreq - 1 of a one-sided spectrum of nfreq_total bins (0: nfreq itself) in columns 0 .. nfreq - 1.
int launch_csd_pairs(const float* X, int64_t C, int64_t nseg, int64_t nseg_total, int64_t ldf, int nfreq, int64_t i0,
                     int64_t ni, float scale, int accumulate, float* partial_ws, float* P, cudaStream_t stream, int f0 = 0,
                     int nfreq_total = 0);
int launch_csd_frames(const float* X, int64_t nseg, int64_t ldf, int nfreq, int ci, int cj, int64_t seg_stride, int navg,
                      int64_t nframes, float scale, float* amp, cudaStream_t stream);
size_t csd_pairs_workspace_bytes(int64_t C, int64_t ni, int nfreq, int64_t nseg);

}  // namespace specgpu
