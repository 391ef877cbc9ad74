// libspecgpu C ABI: context / plan plumbing and the entry points declared in include/specgpu.h.
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "fft.cuh"
#include "kernels.h"

using namespace specgpu;

struct specgpu_ctx {
  int device = 0;
  std::string err;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  int64_t launches = 0;
  int num_sms = 148;
  size_t ws_csd_off = 0;   // offset of the pair partials inside ws (set by specgpu_csd_allpairs)
  int power_max_iter = 0;  // cap of the leading-pair power iteration (0: the kernel default; specgpu_set_power_iterations)
  // specgpu_pipeline: channel groups alternate between two library-owned side streams (forked from / joined to the
  // caller's stream with events) so that a group's log image is consumed while it is still in L2
  // per-signal (min, max) words of the log image (common.cuh): persistent, generation-tagged, never reset between calls
  MinMaxWord* mm64 = nullptr;
  unsigned mm_gen = 0;
  int pipe_group = 0;      // channels per group (0: automatic, see specgpu_set_pipeline_group)
  cudaEvent_t ilk_wait = nullptr, ilk_record = nullptr;   // caller-owned events (specgpu_set_pipeline_interlock)
  bool ilk_armed = false;                                 // true inside specgpu_pipeline only
  cudaStream_t repair_stream = nullptr;      // side stream of the non-converged-channel repair (see svd_run)
  cudaEvent_t ev_repair_fork = nullptr, ev_repair_join = nullptr;
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  // optional per-kernel timing (specgpu_profile_*): CUDA events recorded around every launch group
  bool prof_on = false;
  std::vector<std::string> prof_names;
  std::vector<double> prof_ms;
  std::vector<int64_t> prof_calls;
#ifndef SPECGPU_EMULATE
  struct ProfRec { int name; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof_open;
  std::vector<cudaEvent_t> prof_pool;
#endif
};

struct specgpu_plan {
  specgpu_ctx* ctx = nullptr;
  specgpu_stft_params p{};
  int log2n = 0;
  int hop = 0;
  double scale = 0.0;  // psd scale in double: 1/(fs*sum w^2) or 1/(sum w)^2
  float* d_window = nullptr;
  float2* d_twM = nullptr;
  float2* d_twN = nullptr;
};

namespace {

// Entry points run on the context's device and leave the caller's current device as they found it (a process that
// drives several GPUs, e.g. through torch, must not have its device changed behind its back).
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

int fail(specgpu_ctx* ctx, int code, const char* fmt, ...) {
  if (ctx) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    ctx->err = buf;
  }
  return code;
}

int cuda_fail(specgpu_ctx* ctx, int e, const char* what) {
  return fail(ctx, SPECGPU_ERR_CUDA, "%s: CUDA error %d (%s)", what, e, cudaGetErrorString((cudaError_t)e));
}

#ifndef SPECGPU_EMULATE
int prof_name_id(specgpu_ctx* ctx, const char* what) {
  for (size_t i = 0; i < ctx->prof_names.size(); ++i)
    if (ctx->prof_names[i] == what) return (int)i;
  ctx->prof_names.push_back(what);
  ctx->prof_ms.push_back(0.0);
  ctx->prof_calls.push_back(0);
  return (int)ctx->prof_names.size() - 1;
}
cudaEvent_t prof_event(specgpu_ctx* ctx) {
  if (!ctx->prof_pool.empty()) {
    cudaEvent_t e = ctx->prof_pool.back();
    ctx->prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
// Fold finished event pairs into the per-name totals (synchronises on the last recorded event).
void prof_collect(specgpu_ctx* ctx) {
  for (auto& r : ctx->prof_open) {
    cudaEventSynchronize(r.e1);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      ctx->prof_ms[r.name] += ms;
      ctx->prof_calls[r.name] += 1;
    }
    ctx->prof_pool.push_back(r.e0);
    ctx->prof_pool.push_back(r.e1);
  }
  ctx->prof_open.clear();
}
struct ProfScope {
  specgpu_ctx* ctx;
  cudaStream_t st;
  cudaEvent_t e1 = nullptr;
  ProfScope(specgpu_ctx* c, cudaStream_t s, const char* what) : ctx(c), st(s) {
    if (!ctx->prof_on) return;
    if (ctx->prof_open.size() >= 8192) prof_collect(ctx);
    cudaEvent_t e0 = prof_event(ctx);
    e1 = prof_event(ctx);
    cudaEventRecord(e0, st);
    ctx->prof_open.push_back({prof_name_id(ctx, what), e0, e1});
  }
  ~ProfScope() {
    if (e1) cudaEventRecord(e1, st);
  }
};
#else
struct ProfScope {
  ProfScope(specgpu_ctx*, cudaStream_t, const char*) {}
};
#endif

// `stream` (void*) must be in scope at every use.
#define CHECK_LAUNCH(ctx, expr, what, nlaunch)                       \
  do {                                                               \
    int e__;                                                         \
    {                                                                \
      ProfScope prof__((ctx), (cudaStream_t)stream, what);           \
      e__ = (expr);                                                  \
    }                                                                \
    if (e__ != 0) return cuda_fail(ctx, e__, what);                  \
    (ctx)->launches += (nlaunch);                                    \
  } while (0)

// Grow-on-demand device workspace.  Growing synchronises the device (cudaFree), so production callers
// reserve once with specgpu_workspace_reserve().
int ensure_ws(specgpu_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->ws_bytes) return SPECGPU_OK;
  if (ctx->ws) {
    cudaDeviceSynchronize();
    cudaFree(ctx->ws);
    ctx->ws = nullptr;
    ctx->ws_bytes = 0;
  }
  size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) return fail(ctx, SPECGPU_ERR_WORKSPACE, "workspace of %zu bytes: %s", want, cudaGetErrorString(e));
  ctx->ws = p;
  ctx->ws_bytes = want;
  return SPECGPU_OK;
}

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base(static_cast<char*>(b)) {}
  template <class T>
  T* take(size_t count) {
    // 256-byte aligned ABSOLUTE addresses (vector loads/stores in the kernels rely on it)
    const uintptr_t a = (reinterpret_cast<uintptr_t>(base) + off + 255) & ~(uintptr_t)255;
    off = a - reinterpret_cast<uintptr_t>(base);
    T* p = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return p;
  }
};
// Upper bound of what a Carver consumes for these parts from any base address; a multiple of 256.
inline size_t carve_size(std::initializer_list<size_t> parts) {
  size_t off = 0;
  for (size_t b : parts) off = ((off + 255) & ~(size_t)255) + b;
  return ((off + 255) & ~(size_t)255) + 256;
}

// The persistent (min, max) buffer for up to 65535 signals and a fresh generation for this call.
constexpr int64_t kMaxBatch = 65535;
int minmax_begin(specgpu_ctx* ctx, MinMaxWord** mm, unsigned* gen) {
  if (!ctx->mm64 || ctx->mm_gen == 0xffffffffu) {
    if (!ctx->mm64) {
      // (+ 2 words behind the pairs: the STFT's dynamic tile counter and its exit count, zero between launches)
      cudaError_t e = cudaMalloc(&ctx->mm64, (size_t)(kMaxBatch * 2 + 2) * sizeof(MinMaxWord));
      if (e != cudaSuccess) return fail(ctx, SPECGPU_ERR_WORKSPACE, "min/max buffer: %s", cudaGetErrorString(e));
    } else {
      cudaDeviceSynchronize();     // generation counter wrapped (once per 4 billion calls)
    }
    cudaMemset(ctx->mm64, 0, (size_t)(kMaxBatch * 2 + 2) * sizeof(MinMaxWord));
    ctx->mm_gen = 0;
  }
  *mm = ctx->mm64;
  *gen = ++ctx->mm_gen;
  return SPECGPU_OK;
}

int64_t num_segments(int64_t n, int nperseg, int noverlap) {
  if (n < nperseg) return 0;
  return (n - noverlap) / (nperseg - noverlap);
}

StftArgs make_args(const specgpu_plan* plan, const float* x, int64_t n, int64_t ldx, int64_t first_start, int64_t nseg,
                   float scale, void* out, int64_t ld_out, MinMaxWord* mm, unsigned mm_gen = 0) {
  StftArgs a{};
  a.x = x;
  a.n = n;
  a.ldx = ldx;
  a.first_start = first_start;
  a.nseg = nseg;
  a.hop = plan->hop;
  a.detrend = plan->p.detrend;
  a.vec_ok = ((reinterpret_cast<uintptr_t>(x) & 7) == 0) && (ldx % 2 == 0) && (plan->hop % 2 == 0) &&
             (first_start % 2 == 0);
  a.scale = scale;
  a.eps = (float)plan->p.eps;
  a.window = plan->d_window;
  a.twM = plan->d_twM;
  a.twN = plan->d_twN;
  a.out = out;
  a.ld_out = ld_out;
  a.minmax = mm;
  a.minmax_gen = mm_gen;
  return a;
}

int check_signal_args(specgpu_ctx* ctx, const specgpu_plan* plan, const void* x, int64_t B, int64_t n, int64_t ldx) {
  if (!ctx || !plan) return SPECGPU_ERR_INVALID_ARG;
  if (B < 0 || n < 0 || ldx < n) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "bad signal shape B=%lld n=%lld ldx=%lld", (long long)B, (long long)n, (long long)ldx);
  if (B > 65535) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "batch %lld exceeds 65535 per call", (long long)B);
  if (B > 0 && n > 0 && x == nullptr) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "null signal pointer");
  return SPECGPU_OK;
}

}  // namespace

extern "C" {

int specgpu_version(void) { return SPECGPU_VERSION_MAJOR * 1000 + SPECGPU_VERSION_MINOR; }

int specgpu_init(int device, specgpu_ctx** out) {
  if (!out) return SPECGPU_ERR_INVALID_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SPECGPU_ERR_CUDA;
  DeviceGuard dev_guard(device);
  specgpu_ctx* ctx = new (std::nothrow) specgpu_ctx();
  if (!ctx) return SPECGPU_ERR_WORKSPACE;
  ctx->device = device;
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) ctx->num_sms = v;
#ifndef SPECGPU_EMULATE
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (major != 10) {
    delete ctx;
    return SPECGPU_ERR_CUDA;  // sm_100a cubin only: no other architecture, no fallback
  }
#endif
  *out = ctx;
  return SPECGPU_OK;
}

int specgpu_destroy(specgpu_ctx* ctx) {
  if (!ctx) return SPECGPU_OK;
  DeviceGuard dev_guard(ctx->device);
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->mm64) cudaFree(ctx->mm64);
#ifndef SPECGPU_EMULATE
  for (int i = 0; i < 2; ++i) {
    if (ctx->side[i]) cudaStreamDestroy(ctx->side[i]);
    if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->repair_stream) cudaStreamDestroy(ctx->repair_stream);
  if (ctx->ev_repair_fork) cudaEventDestroy(ctx->ev_repair_fork);
  if (ctx->ev_repair_join) cudaEventDestroy(ctx->ev_repair_join);
  prof_collect(ctx);
  for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
#endif
  delete ctx;
  return SPECGPU_OK;
}

const char* specgpu_last_error(const specgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int specgpu_workspace_reserve(specgpu_ctx* ctx, int64_t bytes) {
  if (!ctx || bytes < 0) return SPECGPU_ERR_INVALID_ARG;
  DeviceGuard dev_guard(ctx->device);
  return ensure_ws(ctx, (size_t)bytes);
}

int64_t specgpu_launch_count(const specgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

int specgpu_copy_rows(specgpu_ctx* ctx, void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t row_bytes,
                      int64_t nrows, void* stream) {
  if (!ctx || row_bytes < 0 || nrows < 0 || dst_pitch < row_bytes || src_pitch < row_bytes) return SPECGPU_ERR_INVALID_ARG;
  if (row_bytes == 0 || nrows == 0) return SPECGPU_OK;
  if (!dst || !src) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "null pointer");
#ifdef SPECGPU_EMULATE
  for (int64_t r = 0; r < nrows; ++r)
    std::memcpy(static_cast<char*>(dst) + r * dst_pitch, static_cast<const char*>(src) + r * src_pitch, (size_t)row_bytes);
  (void)stream;
#else
  DeviceGuard dev_guard(ctx->device);
  cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)row_bytes, (size_t)nrows, cudaMemcpyDefault,
                                    (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(ctx, (int)e, "copy_rows");
#endif
  return SPECGPU_OK;
}

int specgpu_set_pipeline_interlock(specgpu_ctx* ctx, void* wait_event, void* record_event) {
  if (!ctx) return SPECGPU_ERR_INVALID_ARG;
  ctx->ilk_wait = (cudaEvent_t)wait_event;
  ctx->ilk_record = (cudaEvent_t)record_event;
  return SPECGPU_OK;
}

int specgpu_set_pipeline_group(specgpu_ctx* ctx, int32_t channels) {
  if (!ctx || channels < 0) return SPECGPU_ERR_INVALID_ARG;
  ctx->pipe_group = channels;
  return SPECGPU_OK;
}

int specgpu_set_power_iterations(specgpu_ctx* ctx, int32_t max_iter) {
  if (!ctx || max_iter < 0) return SPECGPU_ERR_INVALID_ARG;
  ctx->power_max_iter = max_iter;
  return SPECGPU_OK;
}

int specgpu_profile_enable(specgpu_ctx* ctx, int enable) {
  if (!ctx) return SPECGPU_ERR_INVALID_ARG;
#ifndef SPECGPU_EMULATE
  DeviceGuard dev_guard(ctx->device);
  prof_collect(ctx);
#endif
  ctx->prof_on = enable != 0;
  if (enable) {
    ctx->prof_names.clear();
    ctx->prof_ms.clear();
    ctx->prof_calls.clear();
  }
  return SPECGPU_OK;
}

int specgpu_profile_count(specgpu_ctx* ctx) {
  if (!ctx) return 0;
#ifndef SPECGPU_EMULATE
  DeviceGuard dev_guard(ctx->device);
  prof_collect(ctx);
#endif
  return (int)ctx->prof_names.size();
}

const char* specgpu_profile_name(const specgpu_ctx* ctx, int i) {
  return (ctx && i >= 0 && i < (int)ctx->prof_names.size()) ? ctx->prof_names[i].c_str() : "";
}
double specgpu_profile_ms(const specgpu_ctx* ctx, int i) {
  return (ctx && i >= 0 && i < (int)ctx->prof_ms.size()) ? ctx->prof_ms[i] : 0.0;
}
int64_t specgpu_profile_calls(const specgpu_ctx* ctx, int i) {
  return (ctx && i >= 0 && i < (int)ctx->prof_calls.size()) ? ctx->prof_calls[i] : 0;
}

int specgpu_plan_create(specgpu_ctx* ctx, const specgpu_stft_params* p, const double* window_host, specgpu_plan** out) {
  if (!ctx || !p || !out) return SPECGPU_ERR_INVALID_ARG;
  *out = nullptr;
  const int N = p->nperseg;
  int log2n = 0;
  while ((1 << log2n) < N) ++log2n;
  if (N < 8 || N > 8192 || (1 << log2n) != N)
    return fail(ctx, SPECGPU_ERR_UNSUPPORTED_NPERSEG, "nperseg=%d: must be a power of two in [8, 8192]", N);
  if (p->noverlap < 0 || p->noverlap >= N) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "noverlap must be less than nperseg.");
  if (p->detrend < 0 || p->detrend > 2) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "unknown detrend %d", p->detrend);
  if (p->scaling != SPECGPU_SCALING_DENSITY && p->scaling != SPECGPU_SCALING_SPECTRUM)
    return fail(ctx, SPECGPU_ERR_INVALID_ARG, "Unknown scaling: %d", p->scaling);
  if (p->window == SPECGPU_WINDOW_CUSTOM && !window_host) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "custom window needs window_host");
  if (p->window < 0 || p->window > 3) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "unknown window %d", p->window);
  if (!(p->fs > 0.0)) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "fs must be positive");

  std::vector<double> w(N);
  const double two_pi = 6.283185307179586476925286766559;
  for (int i = 0; i < N; ++i) {
    switch (p->window) {
      case SPECGPU_WINDOW_CUSTOM: w[i] = window_host[i]; break;
      case SPECGPU_WINDOW_HANN: w[i] = 0.5 - 0.5 * std::cos(two_pi * i / N); break;      // periodic (fftbins=True)
      case SPECGPU_WINDOW_HAMMING: w[i] = 0.54 - 0.46 * std::cos(two_pi * i / N); break;
      default: w[i] = 1.0;
    }
  }
  double sw = 0.0, sw2 = 0.0;
  for (double v : w) {
    sw += v;
    sw2 += v * v;
  }
  specgpu_plan* plan = new (std::nothrow) specgpu_plan();
  if (!plan) return SPECGPU_ERR_WORKSPACE;
  plan->ctx = ctx;
  plan->p = *p;
  plan->log2n = log2n;
  plan->hop = N - p->noverlap;
  plan->scale = (p->scaling == SPECGPU_SCALING_DENSITY) ? 1.0 / (p->fs * sw2) : 1.0 / (sw * sw);

  const int M = N / 2;
  std::vector<float> wf(N);
  for (int i = 0; i < N; ++i) wf[i] = (float)w[i];
  // per-pass twiddle tables of the M-point complex transform (fft.cuh): pass p holds exp(-2 pi i k r / (P R)) at
  // [r * P + k], P = product of the earlier radices
  const int log2m = log2n - 1;
  const int ntw = std::max(fft_twiddle_count(log2m), 1);
  std::vector<float2> twM(ntw), twN(M / 2 + 1);
  {
    int off = 0, P = fft_radix_at(log2m, 0);
    for (int pass = 1; pass <= 2; ++pass) {
      const int R = fft_radix_at(log2m, pass);
      if (R <= 1) break;
      for (int r = 0; r < R; ++r)
        for (int k = 0; k < P; ++k) {
          const int j = (int)(((int64_t)k * r) % ((int64_t)P * R));
          twM[off + r * P + k].x = (float)std::cos(two_pi * j / ((double)P * R));
          twM[off + r * P + k].y = (float)(-std::sin(two_pi * j / ((double)P * R)));
        }
      off += R * P;
      P *= R;
    }
  }
  for (int k = 0; k <= M / 2; ++k) {
    twN[k].x = (float)std::cos(two_pi * k / N);
    twN[k].y = (float)(-std::sin(two_pi * k / N));
  }
  DeviceGuard dev_guard(ctx->device);
  cudaError_t e;
  if ((e = cudaMalloc(&plan->d_window, N * sizeof(float))) != cudaSuccess ||
      (e = cudaMalloc(&plan->d_twM, ntw * sizeof(float2))) != cudaSuccess ||
      (e = cudaMalloc(&plan->d_twN, (M / 2 + 1) * sizeof(float2))) != cudaSuccess) {
    specgpu_plan_destroy(plan);
    return cuda_fail(ctx, (int)e, "plan tables");
  }
  cudaMemcpy(plan->d_window, wf.data(), N * sizeof(float), cudaMemcpyHostToDevice);
  cudaMemcpy(plan->d_twM, twM.data(), ntw * sizeof(float2), cudaMemcpyHostToDevice);
  cudaMemcpy(plan->d_twN, twN.data(), (M / 2 + 1) * sizeof(float2), cudaMemcpyHostToDevice);
  *out = plan;
  return SPECGPU_OK;
}

int specgpu_plan_destroy(specgpu_plan* plan) {
  if (!plan) return SPECGPU_OK;
  DeviceGuard dev_guard(plan->ctx ? plan->ctx->device : 0);
  if (plan->d_window) cudaFree(plan->d_window);
  if (plan->d_twM) cudaFree(plan->d_twM);
  if (plan->d_twN) cudaFree(plan->d_twN);
  delete plan;
  return SPECGPU_OK;
}

int64_t specgpu_plan_num_segments(const specgpu_plan* plan, int64_t n) {
  if (!plan) return 0;
  return num_segments(n, plan->p.nperseg, plan->p.noverlap);
}

int32_t specgpu_plan_num_freqs(const specgpu_plan* plan) { return plan ? plan->p.nperseg / 2 + 1 : 0; }

int specgpu_plan_axes(const specgpu_plan* plan, int64_t n, double* f_host, double* t_host) {
  if (!plan) return SPECGPU_ERR_INVALID_ARG;
  const int N = plan->p.nperseg;
  // np.fft.rfftfreq(N, 1/fs): arange(N/2+1) / (N * (1/fs)) ; segment centres (j*hop + N/2) / fs
  if (f_host) {
    const double val = 1.0 / (N * (1.0 / plan->p.fs));
    for (int k = 0; k <= N / 2; ++k) f_host[k] = k * val;
  }
  if (t_host) {
    const int64_t nseg = num_segments(n, N, plan->p.noverlap);
    for (int64_t j = 0; j < nseg; ++j) t_host[j] = (double)(j * plan->hop + N / 2) / plan->p.fs;
  }
  return SPECGPU_OK;
}

int specgpu_spectrogram(specgpu_ctx* ctx, const specgpu_plan* plan, const float* x, int64_t B, int64_t n, int64_t ldx,
                        float* Sxx, int64_t ldt, void* stream) {
  int rc = check_signal_args(ctx, plan, x, B, n, ldx);
  if (rc) return rc;
  const int64_t nseg = num_segments(n, plan->p.nperseg, plan->p.noverlap);
  if (nseg == 0 || B == 0) return SPECGPU_OK;
  if (!Sxx || ldt < nseg) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "bad output (ldt=%lld < nseg=%lld)", (long long)ldt, (long long)nseg);
  StftArgs a = make_args(plan, x, n, ldx, 0, nseg, (float)plan->scale, Sxx, ldt, nullptr);
  CHECK_LAUNCH(ctx, launch_stft(plan->log2n, STFT_MODE_PSD, a, B, (cudaStream_t)stream), "stft_kernel", 1);
  return SPECGPU_OK;
}

int specgpu_specgr(specgpu_ctx* ctx, const specgpu_plan* plan, const float* x, int64_t B, int64_t n, int64_t ldx,
                   float* S, int64_t ldt, float* minmax, void* stream) {
  int rc = check_signal_args(ctx, plan, x, B, n, ldx);
  if (rc) return rc;
  const int64_t nseg = num_segments(n, plan->p.nperseg, plan->p.noverlap);
  if (nseg == 0 || B == 0) return SPECGPU_OK;
  if (!S || ldt < nseg) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "bad output (ldt=%lld < nseg=%lld)", (long long)ldt, (long long)nseg);
  DeviceGuard dev_guard(ctx->device);
  MinMaxWord* mm = nullptr;
  unsigned gen = 0;
  if ((rc = minmax_begin(ctx, &mm, &gen))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  StftArgs a = make_args(plan, x, n, ldx, 0, nseg, (float)plan->scale, S, ldt, mm, gen);
  if (!(std::getenv("SPECGPU_STFT_DYNAMIC") && std::getenv("SPECGPU_STFT_DYNAMIC")[0] == '0'))
    a.dyn = reinterpret_cast<unsigned*>(ctx->mm64 + kMaxBatch * 2);     // dynamic tile walk (see specgpu_pipeline)
  CHECK_LAUNCH(ctx, launch_stft(plan->log2n, STFT_MODE_LOGPSD, a, B, st), "stft_kernel", 1);
  CHECK_LAUNCH(ctx, launch_lognorm(S, B, plan->p.nperseg / 2, nseg, ldt, mm, minmax, st), "lognorm", 1);
  return SPECGPU_OK;
}

int64_t specgpu_stft_num_segments(const specgpu_plan* plan, int64_t n, int boundary_zeros, int padded) {
  if (!plan) return 0;
  const int N = plan->p.nperseg;
  const int hop = plan->hop;
  int64_t len = n + (boundary_zeros ? 2 * (int64_t)(N / 2) : 0);
  if (padded) {
    // scipy: nadd = (-(len - nperseg) % nstep) % nperseg
    int64_t r = (len - N) % hop;
    if (r < 0) r += hop;
    int64_t nadd = ((hop - r) % hop) % N;
    len += nadd;
  }
  return num_segments(len, N, plan->p.noverlap);
}

int specgpu_stft(specgpu_ctx* ctx, const specgpu_plan* plan, const float* x, int64_t B, int64_t n, int64_t ldx,
                 int boundary_zeros, int padded, float* Z, int64_t ldt, void* stream) {
  int rc = check_signal_args(ctx, plan, x, B, n, ldx);
  if (rc) return rc;
  const int64_t nseg = specgpu_stft_num_segments(plan, n, boundary_zeros, padded);
  if (nseg == 0 || B == 0) return SPECGPU_OK;
  if (!Z || ldt < nseg) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "bad output (ldt=%lld < nseg=%lld)", (long long)ldt, (long long)nseg);
  const int64_t first = boundary_zeros ? -(int64_t)(plan->p.nperseg / 2) : 0;
  StftArgs a = make_args(plan, x, n, ldx, first, nseg, (float)std::sqrt(plan->scale), Z, ldt, nullptr);
  CHECK_LAUNCH(ctx, launch_stft(plan->log2n, STFT_MODE_COMPLEX, a, B, (cudaStream_t)stream), "stft_kernel", 1);
  return SPECGPU_OK;
}


// ---- array helpers -------------------------------------------------------------------------------
static int check_matrix_args(specgpu_ctx* ctx, const void* src, int64_t B, int64_t rows, int64_t cols, int64_t ld) {
  if (!ctx) return SPECGPU_ERR_INVALID_ARG;
  if (B < 0 || rows < 0 || cols < 0 || ld < cols)
    return fail(ctx, SPECGPU_ERR_INVALID_ARG, "bad matrix shape B=%lld rows=%lld cols=%lld ld=%lld", (long long)B,
                (long long)rows, (long long)cols, (long long)ld);
  if (B > 65535) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "batch %lld exceeds 65535 per call", (long long)B);
  if (B * rows * cols > 0 && src == nullptr) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "null matrix pointer");
  return SPECGPU_OK;
}

int specgpu_rescale(specgpu_ctx* ctx, const float* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, float* dst,
                    void* stream) {
  int rc = check_matrix_args(ctx, src, B, rows, cols, ld);
  if (rc) return rc;
  if (B * rows * cols == 0) return SPECGPU_OK;
  if (!dst) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "null output pointer");
  DeviceGuard dev_guard(ctx->device);
  if ((rc = ensure_ws(ctx, carve_size({(size_t)B * 2 * sizeof(unsigned)})))) return rc;
  Carver cv(ctx->ws);
  unsigned* mm = cv.take<unsigned>(B * 2);
  CHECK_LAUNCH(ctx, launch_rescale(src, B, rows, cols, ld, dst, mm, (cudaStream_t)stream), "rescale", 3);
  return SPECGPU_OK;
}

int specgpu_norm(specgpu_ctx* ctx, const float* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, float* dst,
                 void* stream) {
  int rc = check_matrix_args(ctx, src, B, rows, cols, ld);
  if (rc) return rc;
  if (B * rows * cols == 0) return SPECGPU_OK;
  if (!dst) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "null output pointer");
  DeviceGuard dev_guard(ctx->device);
  const int parts = norm_num_parts(rows, cols);
  if ((rc = ensure_ws(ctx, carve_size({(size_t)B * parts * 2 * sizeof(double)})))) return rc;
  Carver cv(ctx->ws);
  double* sums = cv.take<double>(B * parts * 2);
  CHECK_LAUNCH(ctx, launch_norm(src, B, rows, cols, ld, dst, sums, (cudaStream_t)stream), "norm", 2);
  return SPECGPU_OK;
}

int specgpu_clip(specgpu_ctx* ctx, const float* src, int64_t n, float* dst, void* stream) {
  if (!ctx || n < 0) return SPECGPU_ERR_INVALID_ARG;
  if (n == 0) return SPECGPU_OK;
  if (!src || !dst) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "null pointer");
  CHECK_LAUNCH(ctx, launch_clip_neg(src, n, dst, (cudaStream_t)stream), "clip", 1);
  return SPECGPU_OK;
}

int specgpu_quantfilt(specgpu_ctx* ctx, const float* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, float thr,
                      float* dst, float* thr_out, uint8_t* mask, void* stream) {
  int rc = check_matrix_args(ctx, src, B, rows, cols, ld);
  if (rc) return rc;
  if (!(thr >= 0.0f && thr <= 1.0f)) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "Quantiles must be in the range [0, 1]");
  if (rows > 1024) return fail(ctx, SPECGPU_ERR_UNSUPPORTED_SHAPE, "quantfilt: rows=%lld > 1024", (long long)rows);
  if (B * rows * cols == 0) return SPECGPU_OK;
  // numpy (float32 array): pos = float32(rows-1) * float32(thr); lo = floor(pos); g = pos - lo
  volatile float pos = (float)(rows - 1) * thr;
  int lo = (int)std::floor((float)pos);
  float g = (float)pos - (float)lo;
  if (lo >= rows - 1) {
    lo = (int)rows - 1;
    g = 0.f;
  }
  CHECK_LAUNCH(ctx, launch_quantfilt(src, B, rows, cols, ld, lo, g, dst, thr_out, mask, (cudaStream_t)stream), "quantfilt", 1);
  return SPECGPU_OK;
}

// ---- SVD denoise ---------------------------------------------------------------------------------
namespace {

// omega(beta) of the notebook (denoising_by_svd.ipynb:155-159).  beta = np.min(shape) / np.max(shape) is an np.float64
// (a strong scalar), so omega(beta) is float64 and t* = omega * np.median(s) is evaluated and compared in float64 even
// when s is float32: the device multiplies the double omega by the float32 median in double.
double omega_of(double beta) { return 0.56 * std::pow(beta, 3.0) - 0.95 * std::pow(beta, 2.0) + 1.82 * beta + 1.43; }   // as Python: beta ** 3

struct SvdWs {
  void* tri = nullptr;          // scratch of the values-first eigensolver (eig_tridiag.cu)
  int32_t* flagged = nullptr;   // [B] status snapshot for the side-stream repair (pipeline, tiled scratch image)
  float* G;
  float* U;
  float* lam;
  int32_t* plan;
  float* gram_partial;
  void* jacobi;
};

size_t svd_ws_bytes(int64_t B, int64_t rows, bool tc, bool full) {
  return carve_size({(size_t)B * rows * rows * 8, (size_t)B * rows * rows * 4, (size_t)B * rows * 4, (size_t)B * 16,
                     tc ? gram_tc_workspace_bytes(B, rows) : 0, full ? jacobi_workspace_bytes(B, (int)rows) : 0,
                     full && eig_tridiag_supported((int)rows) ? eig_tridiag_workspace_bytes(B, (int)rows) : 0});
}

SvdWs svd_carve(void* ws, int64_t B, int64_t rows, bool tc, bool full) {
  Carver cv(ws);
  SvdWs w{};
  w.G = reinterpret_cast<float*>(cv.take<double>(B * rows * rows));   // float or double Gram (see svd_run)
  w.U = cv.take<float>(B * rows * rows);
  w.lam = cv.take<float>(B * rows);
  w.plan = cv.take<int32_t>(B * 4);
  w.gram_partial = tc ? cv.take<float>(gram_tc_workspace_bytes(B, rows) / 4) : nullptr;
  w.jacobi = full ? (void*)cv.take<char>(jacobi_workspace_bytes(B, (int)rows)) : nullptr;
  w.tri = (full && eig_tridiag_supported((int)rows)) ? (void*)cv.take<char>(eig_tridiag_workspace_bytes(B, (int)rows)) : nullptr;
  return w;
}

// Shared body of svd_denoise / compute_signal / pipeline.  kind: 0 explicit range, 1 use_optimal,
// 2 computeSignal.  `ws_base` lets the pipeline place the SVD scratch behind its own.
//
// `raw_mm` != nullptr (pipeline only): S holds the un-normalised log image and raw_mm its per-matrix
// (min, max); the normalisation is then fused into the Gram producer and the rank-1 projection, which
// also writes the normalised image back over S.
// `fallback`: after the power iteration also enqueue the full float64 solver for the matrices whose iteration did not
// converge (info[3] would be 1); three launches that exit at once when every matrix converged.
// `Limg` / `ldL` (pipeline only, with raw_mm): the raw log image lives in a separate scratch buffer (tiled when ldL < 0,
// see img_off) and S is only written -- by the rank-1 projection, normalised, with pitch ldo.
int svd_run(specgpu_ctx* ctx, const SvdWs& w, float* S, const MinMaxWord* raw_mm, int64_t B, int64_t rows, int64_t cols,
            int64_t ld, int kind, int start, int stop, int clip, bool power_ok, bool fallback, void* out, int out_f64,
            int64_t ldo, float* s_out, int32_t* info, cudaStream_t st, float l2_pin = 0.f, const float* Limg = nullptr,
            int64_t ldL = 0, const int64_t* pre_gram = nullptr /* {nchunk, per}: partials already written by stft_gram */,
            R1Tiles r1_tiles = R1Tiles{} /* float32 tiles of the projection written in the same pass (rank-1 route only) */) {
  void* stream = (void*)st;
  const float* Lsrc = Limg ? Limg : S;       // what the Gram and projection kernels read
  const int64_t ldsrc = Limg ? ldL : ld;
  const bool tc = power_ok && gram_tc_supported(rows);   // TF32 Gram only feeds the leading-pair route
  // full decomposition: Gram and Jacobi in double (float would square the condition number into the noise floor)
  const int g_f64 = (!power_ok && eig_jacobi_f64_supported((int)rows)) ? 1 : 0;
  bool side_repair = false;
  bool gram_raw = false;     // the Gram partials are of the un-normalised image (TMA-fed kernel) and carry row sums
  bool gram_tma = false;
  if (tc && pre_gram != nullptr) {
    gram_raw = true;            // the fused STFT kernel left raw-operand partials (with row sums) behind
  } else if (tc) {
    // tensor-core Gram partials, then ONE cluster kernel per matrix that sums them and runs the power iteration
    int e__ = 1;
    if (!std::getenv("SPECGPU_NO_GRAM_TMA")) {
      ProfScope prof__(ctx, st, "gram_tc");
      e__ = launch_gram_tma(Lsrc, B, rows, cols, ldsrc, w.gram_partial, ctx->num_sms, st, l2_pin);
    }
    if (e__ == 0) {
      gram_tma = true;
      gram_raw = raw_mm != nullptr;
      ctx->launches += 1;
    } else if (e__ == 1) {   // rows not 16-byte aligned: the cp.async-fed kernel normalises the operands itself
      CHECK_LAUNCH(ctx, launch_gram_tc(Lsrc, B, rows, cols, ldsrc, raw_mm, w.gram_partial, nullptr, ctx->num_sms, st, l2_pin), "gram_tc", 1);
    } else {
      return cuda_fail(ctx, e__, "gram_tc");
    }
  } else {
    if (raw_mm) {   // no fused route for this shape: normalise in place first
      CHECK_LAUNCH(ctx, launch_lognorm(S, B, rows, cols, ld, raw_mm, nullptr, st), "lognorm", 1);
      raw_mm = nullptr;
    }
    CHECK_LAUNCH(ctx, launch_gram_simt(S, B, rows, cols, ld, w.G, g_f64, st), "gram_simt", 1);
  }
  if (power_ok) {
    if (tc) {
      int64_t nchunk = 0, per = 0;
      if (pre_gram != nullptr) {
        nchunk = pre_gram[0];
        per = pre_gram[1];
      } else {
        gram_tc_geometry(B, cols, ctx->num_sms, gram_tma ? 1 : 0, &nchunk, &per);
      }
      CHECK_LAUNCH(ctx, launch_gram_eig(w.gram_partial, nchunk, per, B, (int)rows, ctx->power_max_iter, w.U, w.lam, w.plan, st,
                                        gram_raw ? raw_mm : nullptr, cols, gram_tma ? 1 : 0, w.flagged),
                   "gram_eig", 1);
    } else {
      CHECK_LAUNCH(ctx, launch_eig_power(w.G, B, (int)rows, ctx->power_max_iter, w.U, w.lam, w.plan, st), "eig_power", 1);
    }
    // Matrices whose iteration did not converge (degenerate leading pair, iterate fallen into the null space) are
    // redone by the full solver in float64: Gram matrix of the flagged matrices only, straight from the image (the
    // TF32 Gram is dead after the power iteration and is overwritten), then the cluster Jacobi, which skips the
    // converged ones.  All three launches return at once when nothing is flagged.
    // When the image is in the separate (tiled) scratch buffer, the repair runs on a side stream UNDER the projection
    // and a second, flagged-only projection pass fixes up the (rare) flagged matrices afterwards: with nothing flagged
    // the critical path only sees one launch that exits at once.  Otherwise it runs in line before the projection.
#ifndef SPECGPU_EMULATE
    side_repair = fallback && Limg != nullptr && w.flagged != nullptr && !out_f64 && ctx->repair_stream != nullptr;
#endif
    if (fallback && !side_repair) {
      CHECK_LAUNCH(ctx, launch_gram_simt(Lsrc, B, rows, cols, ldsrc, w.G, 1, st, raw_mm, w.plan), "gram_simt_flagged", 1);
      CHECK_LAUNCH(ctx, launch_eig_jacobi(w.G, 1, B, (int)rows, 1, w.U, w.lam, w.plan, w.jacobi, st), "eig_jacobi", 2);
    }
  } else {
    // Full-spectrum modes.  use_optimal / computeSignal need every singular VALUE but only a few leading vectors: the
    // values-first solver (tridiagonalisation + bisection, then inverse iteration for the vectors the plan asks for) serves
    // them; explicit ranges go there too when the host can see that they only need a few leading vectors.  Whatever it
    // cannot serve is flagged and redone by the Jacobi solver (which skips the rest).
    bool tri = g_f64 && w.tri != nullptr && eig_tridiag_supported((int)rows) && !std::getenv("SPECGPU_NO_TRIDIAG");
    if (tri && kind == 0) {
      int a0 = start < 0 ? 0 : start, e0 = stop > (int)rows ? (int)rows : stop;
      if (e0 < 0) e0 = (e0 + (int)rows > 0) ? e0 + (int)rows : 0;
      if (a0 > (int)rows) a0 = (int)rows;
      if (e0 < a0) e0 = a0;
      const int nk = e0 - a0;
      const bool complement = ((int)rows - nk) < nk;
      const int lead = complement ? a0 : (nk > 0 ? e0 : 0), trailing = complement ? (int)rows - e0 : 0;
      tri = trailing == 0 && lead <= 16;
    }
    if (tri) {
      const double beta0 = (double)std::min(rows, cols) / (double)std::max(rows, cols);
      double* Wrefl = static_cast<double*>(w.jacobi);       // the Jacobi scratch is free until the fallback below
      CHECK_LAUNCH(ctx, launch_eig_tridiag_values(reinterpret_cast<const double*>(w.G), Wrefl, B, (int)rows, w.lam, w.tri, st),
                   "eig_tridiag", 2);
      CHECK_LAUNCH(ctx, launch_svd_plan(w.lam, B, (int)rows, kind, start, stop, omega_of(beta0), w.plan, nullptr, st), "svd_plan", 1);
      CHECK_LAUNCH(ctx, launch_eig_tridiag_vectors(Wrefl, reinterpret_cast<const double*>(w.G), B, (int)rows, w.plan, w.U, w.tri, st),
                   "eig_trivec", 2);
      if (!std::getenv("SPECGPU_TRIDIAG_STRICT"))   // (tests: without the fallback a matrix this route gave up on fails parity)
        CHECK_LAUNCH(ctx, launch_eig_jacobi(w.G, 1, B, (int)rows, 1, w.U, w.lam, w.plan, w.jacobi, st), "eig_jacobi", 2);
    } else {
      CHECK_LAUNCH(ctx, launch_eig_jacobi(w.G, g_f64, B, (int)rows, 0, w.U, w.lam, w.plan, w.jacobi, st), "eig_jacobi", 2);
    }
  }
  const double beta = (double)std::min(rows, cols) / (double)std::max(rows, cols);
  if (!power_ok)   // the power kernel writes the (fixed) default plan itself
    CHECK_LAUNCH(ctx, launch_svd_plan(w.lam, B, (int)rows, kind, start, stop, omega_of(beta), w.plan, s_out, st),
                 "svd_plan", 1);
  if (power_ok && !out_f64) {
    // power_ok implies the range [1, rows): only the leading component is removed
    if (info) {                 // the projection copies the plan rows itself (no memcpy node behind it)
      r1_tiles.plan = w.plan;
      r1_tiles.info = info;
    }
#ifndef SPECGPU_EMULATE
    if (ctx->ilk_armed && ctx->ilk_record) cudaEventRecord(ctx->ilk_record, st);    // specgpu_set_pipeline_interlock
    if (side_repair) {
      cudaStream_t rs = ctx->repair_stream;
      cudaEventRecord(ctx->ev_repair_fork, st);
      cudaStreamWaitEvent(rs, ctx->ev_repair_fork, 0);
      {
        void* stream = (void*)rs;     // CHECK_LAUNCH takes the stream from this name
        CHECK_LAUNCH(ctx, launch_gram_simt(Lsrc, B, rows, cols, ldsrc, w.G, 1, rs, raw_mm, w.plan), "gram_simt_flagged", 1);
        CHECK_LAUNCH(ctx, launch_eig_jacobi(w.G, 1, B, (int)rows, 1, w.U, w.lam, w.plan, w.jacobi, rs), "eig_jacobi", 2);
      }
      cudaEventRecord(ctx->ev_repair_join, rs);
    }
#endif
    CHECK_LAUNCH(ctx, launch_svd_rank1(Lsrc, B, (int)rows, cols, ldsrc, raw_mm, w.U, clip, raw_mm ? S : nullptr, (float*)out, ldo, st, 0, nullptr, r1_tiles),
                 "svd_rank1", 1);
#ifndef SPECGPU_EMULATE
    if (side_repair) {
      cudaStreamWaitEvent(st, ctx->ev_repair_join, 0);
      CHECK_LAUNCH(ctx, launch_svd_rank1(Lsrc, B, (int)rows, cols, ldsrc, raw_mm, w.U, clip, raw_mm ? S : nullptr, (float*)out, ldo, st, 0, w.flagged, r1_tiles),
                   "svd_rank1_flagged", 1);
    }
#endif
  } else {
    CHECK_LAUNCH(ctx, launch_svd_project(S, B, (int)rows, cols, ld, w.U, w.plan, clip, out, out_f64, ldo, st), "svd_project", 1);
  }
  if (info && !(power_ok && !out_f64)) {
    cudaError_t e = cudaMemcpyAsync(info, w.plan, (size_t)B * 16, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_fail(ctx, (int)e, "info copy");
  }
  return SPECGPU_OK;
}

int check_svd_args(specgpu_ctx* ctx, const float* S, int64_t B, int64_t rows, int64_t cols, int64_t ld, const void* out,
                   int64_t ldo) {
  int rc = check_matrix_args(ctx, S, B, rows, cols, ld);
  if (rc) return rc;
  if (rows > cols) return fail(ctx, SPECGPU_ERR_UNSUPPORTED_SHAPE, "svd: rows=%lld > cols=%lld (pass the transpose)", (long long)rows, (long long)cols);
  if (rows > 512) return fail(ctx, SPECGPU_ERR_UNSUPPORTED_SHAPE, "svd: rows=%lld > 512", (long long)rows);
  if (B * rows * cols > 0 && (!out || ldo < cols)) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "bad output (ldo=%lld)", (long long)ldo);
  return SPECGPU_OK;
}

}  // namespace

int specgpu_svd_denoise(specgpu_ctx* ctx, const float* S, int64_t B, int64_t rows, int64_t cols, int64_t ld,
                        int32_t start, int32_t stop, int32_t use_optimal, int32_t clip, int32_t mode, float* out,
                        int64_t ldo, float* s_out, int32_t* info, void* stream) {
  int rc = check_svd_args(ctx, S, B, rows, cols, ld, out, ldo);
  if (rc) return rc;
  if (B * rows * cols == 0) return SPECGPU_OK;
  DeviceGuard dev_guard(ctx->device);
  const bool power_ok = (mode == 0) && !use_optimal && start == 1 && stop >= rows && s_out == nullptr && rows <= 256;
  if ((rc = ensure_ws(ctx, svd_ws_bytes(B, rows, power_ok && gram_tc_supported(rows), true)))) return rc;
  return svd_run(ctx, svd_carve(ctx->ws, B, rows, power_ok && gram_tc_supported(rows), true), const_cast<float*>(S), nullptr, B, rows, cols, ld, use_optimal ? 1 : 0, start, stop, clip,
                 power_ok, true, out, 0, ldo, s_out, info, (cudaStream_t)stream);
}

int specgpu_compute_signal(specgpu_ctx* ctx, const float* S, int64_t B, int64_t rows, int64_t cols, int64_t ld, double* out,
                           int64_t ldo, float* s_out, int32_t* info, void* stream) {
  int rc = check_svd_args(ctx, S, B, rows, cols, ld, out, ldo);
  if (rc) return rc;
  if (B * rows * cols == 0) return SPECGPU_OK;
  DeviceGuard dev_guard(ctx->device);
  if ((rc = ensure_ws(ctx, svd_ws_bytes(B, rows, false, true)))) return rc;
  return svd_run(ctx, svd_carve(ctx->ws, B, rows, false, true), const_cast<float*>(S), nullptr, B, rows, cols, ld, 2, 0, 0, 0, false, false, out, 1, ldo,
                 s_out, info, (cudaStream_t)stream);
}

// ---- tiles -----------------------------------------------------------------------------------------
int specgpu_patch(specgpu_ctx* ctx, const float* src, int64_t n, int64_t rows, int64_t ld, int32_t tile_w, int32_t ntiles,
                  void* out, int32_t out_f64, void* stream) {
  if (!ctx) return SPECGPU_ERR_INVALID_ARG;
  if (n < 0 || rows < 0 || tile_w <= 0 || ntiles < 0 || ld < (int64_t)tile_w * ntiles)
    return fail(ctx, SPECGPU_ERR_INVALID_ARG, "patch: need ld >= tile_w*ntiles (ld=%lld, %d x %d)", (long long)ld, tile_w, ntiles);
  if (n > 65535 || rows > 65535) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "patch: n and rows must be <= 65535");
  if (n * rows * ntiles == 0) return SPECGPU_OK;
  if (!src || !out) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "null pointer");
  CHECK_LAUNCH(ctx, launch_patch(src, n, rows, ld, tile_w, ntiles, out, out_f64, (cudaStream_t)stream), "patch", 1);
  return SPECGPU_OK;
}

int specgpu_unpatch(specgpu_ctx* ctx, const void* tiles, int32_t in_f64, int64_t n, int64_t rows, int32_t tile_w,
                    int32_t ntiles, void* dst, int32_t out_f64, int64_t ld, void* stream) {
  if (!ctx) return SPECGPU_ERR_INVALID_ARG;
  if (n < 0 || rows < 0 || tile_w <= 0 || ntiles < 0 || ld < (int64_t)tile_w * ntiles)
    return fail(ctx, SPECGPU_ERR_INVALID_ARG, "unpatch: need ld >= tile_w*ntiles (ld=%lld, %d x %d)", (long long)ld, tile_w, ntiles);
  if (n > 65535 || rows > 65535) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "unpatch: n and rows must be <= 65535");
  if (n * rows * ntiles == 0) return SPECGPU_OK;
  if (!tiles || !dst) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "null pointer");
  CHECK_LAUNCH(ctx, launch_unpatch(tiles, in_f64, n, rows, tile_w, ntiles, dst, out_f64, ld, (cudaStream_t)stream), "unpatch", 1);
  return SPECGPU_OK;
}

// ---- cv2 image chain (pipeline_data.py:52-72) --------------------------------------------------------------------
namespace {
// OpenCV's Q8.8 Gaussian taps for CV_8U (getGaussianKernel + fixed-point error-diffusion rounding): the double kernel
// (small_gaussian_tab for n <= 7 when sigma <= 0, else exp(-x^2 / (2 sigma^2)) normalised, sigma = 0.3((n-1)/2 - 1) + 0.8)
// is rounded tap by tap from the outside in, carrying the rounding error; the centre tap takes what is left of 256.
void gaussian_taps_q8(int n, double sigma, uint16_t* out) {
  std::vector<double> k(n);
  static const double small[4][7] = {{1.0}, {0.25, 0.5, 0.25}, {0.0625, 0.25, 0.375, 0.25, 0.0625},
                                     {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125}};
  if (n <= 7 && sigma <= 0) {
    for (int i = 0; i < n; ++i) k[i] = small[n / 2][i];
  } else {
    const double s = sigma > 0 ? sigma : ((n - 1) * 0.5 - 1) * 0.3 + 0.8;
    double sum = 0;
    for (int i = 0; i < n; ++i) {
      const double x = i - (n - 1) * 0.5;
      k[i] = std::exp(-0.5 * x * x / (s * s));
      sum += k[i];
    }
    for (int i = 0; i < n; ++i) k[i] /= sum;
  }
  double err = 0.0;
  long total = 0;
  for (int i = 0; i < n / 2; ++i) {
    const double adj = k[i] * 256.0 + err;
    const long v = std::lrint(adj);
    err = adj - (double)v;
    out[i] = out[n - 1 - i] = (uint16_t)v;
    total += 2 * v;
  }
  out[n / 2] = (uint16_t)(256 - total);
}
}  // namespace

int specgpu_gaussblr(specgpu_ctx* ctx, const void* src, int32_t in_f64, int64_t B, int64_t rows, int64_t cols, int64_t ld,
                     int32_t kw, int32_t kh, double* dst, int64_t ldo, uint8_t* u8_out, void* stream) {
  int rc = check_matrix_args(ctx, src, B, rows, cols, ld);
  if (rc) return rc;
  if (kw < 1 || kh < 1 || !(kw & 1) || !(kh & 1) || kw > 255 || kh > 255)
    return fail(ctx, SPECGPU_ERR_INVALID_ARG, "ksize (%d, %d): both must be odd and in [1, 255]", kw, kh);
  if (B * rows * cols == 0) return SPECGPU_OK;
  if (cols > (1 << 30) || rows > 65535) return fail(ctx, SPECGPU_ERR_UNSUPPORTED_SHAPE, "image too large");
  if (!dst || ldo < cols) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "bad output");
  DeviceGuard dev_guard(ctx->device);
  if ((rc = ensure_ws(ctx, imgchain_workspace_bytes(B, rows, cols)))) return rc;
  uint16_t taps[512];
  gaussian_taps_q8(kw, 0.0, taps);
  gaussian_taps_q8(kh, 0.0, taps + kw);
  CHECK_LAUNCH(ctx, launch_gaussblr(src, in_f64, B, rows, cols, ld, taps, kw, kh, ctx->ws, dst, ldo, u8_out, (cudaStream_t)stream),
               "gaussblr", 4 + (u8_out ? 1 : 0));
  return SPECGPU_OK;
}

int specgpu_meansub(specgpu_ctx* ctx, const double* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, double* dst,
                    int64_t ldo, void* stream) {
  int rc = check_matrix_args(ctx, src, B, rows, cols, ld);
  if (rc) return rc;
  if (B * rows * cols == 0) return SPECGPU_OK;
  if (rows > 65535 || cols > 120000)      // the pairwise-sum leaves of a row live in shared memory
    return fail(ctx, SPECGPU_ERR_UNSUPPORTED_SHAPE, "meansub: rows=%lld > 65535 or cols=%lld > 120000", (long long)rows, (long long)cols);
  if (!dst || ldo < cols) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "bad output");
  DeviceGuard dev_guard(ctx->device);
  if ((rc = ensure_ws(ctx, imgchain_workspace_bytes(B, rows, cols)))) return rc;
  CHECK_LAUNCH(ctx, launch_meansub(src, B, rows, cols, ld, ctx->ws, dst, ldo, (cudaStream_t)stream), "meansub", 3);
  return SPECGPU_OK;
}

int specgpu_morph(specgpu_ctx* ctx, const void* src, int32_t in_f64, int64_t B, int64_t rows, int64_t cols, int64_t ld,
                  double* dst, int64_t ldo, uint8_t* u8_out, void* stream) {
  int rc = check_matrix_args(ctx, src, B, rows, cols, ld);
  if (rc) return rc;
  if (B * rows * cols == 0) return SPECGPU_OK;
  if (cols > (1 << 30) || rows > 65535) return fail(ctx, SPECGPU_ERR_UNSUPPORTED_SHAPE, "image too large");
  if (!dst || ldo < cols) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "bad output");
  DeviceGuard dev_guard(ctx->device);
  if ((rc = ensure_ws(ctx, imgchain_workspace_bytes(B, rows, cols)))) return rc;
  CHECK_LAUNCH(ctx, launch_morph(src, in_f64, B, rows, cols, ld, ctx->ws, dst, ldo, u8_out, (cudaStream_t)stream), "morph", 4 + (u8_out ? 1 : 0));
  return SPECGPU_OK;
}

int specgpu_filter_chain(specgpu_ctx* ctx, const float* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, float thr,
                         int32_t kw, int32_t kh, double* dst, int64_t ldo, void* stream) {
  int rc = check_matrix_args(ctx, src, B, rows, cols, ld);
  if (rc) return rc;
  if (kw < 1 || kh < 1 || !(kw & 1) || !(kh & 1) || kw > 255 || kh > 255)
    return fail(ctx, SPECGPU_ERR_INVALID_ARG, "ksize (%d, %d): both must be odd and in [1, 255]", kw, kh);
  if (!(thr >= 0.0f && thr <= 1.0f)) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "Quantiles must be in the range [0, 1]");
  if (B * rows * cols == 0) return SPECGPU_OK;
  if (cols > 120000 || rows > 1024)
    return fail(ctx, SPECGPU_ERR_UNSUPPORTED_SHAPE, "filter_chain: rows=%lld > 1024 or cols=%lld > 120000", (long long)rows, (long long)cols);
  if (!dst || ldo < cols) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "bad output");
  DeviceGuard dev_guard(ctx->device);
  // workspace: the image-chain planes, then the thresholded float32 image (quantfilt writes it with the source's pitch)
  const size_t img_bytes = (imgchain_workspace_bytes(B, rows, cols) + 255) & ~(size_t)255;
  if ((rc = ensure_ws(ctx, img_bytes + (size_t)B * rows * ld * sizeof(float) + 256))) return rc;
  float* q = reinterpret_cast<float*>(static_cast<char*>(ctx->ws) + img_bytes);
  if ((rc = specgpu_quantfilt(ctx, src, B, rows, cols, ld, thr, q, nullptr, nullptr, stream))) return rc;
  uint16_t taps[512];
  gaussian_taps_q8(kw, 0.0, taps);
  gaussian_taps_q8(kh, 0.0, taps + kw);
  CHECK_LAUNCH(ctx, launch_filter_tail(q, B, rows, cols, ld, taps, kw, kh, ctx->ws, dst, ldo, (cudaStream_t)stream),
               "filter_tail", 12);
  return SPECGPU_OK;
}

// ---- cross-power spectrum ----------------------------------------------------------------------------
int specgpu_csd_spectra(specgpu_ctx* ctx, const specgpu_plan* plan, const float* x, int64_t C, int64_t n, int64_t ldx,
                        float* X, int64_t ldf, void* stream) {
  int rc = check_signal_args(ctx, plan, x, C, n, ldx);
  if (rc) return rc;
  const int64_t nseg = num_segments(n, plan->p.nperseg, plan->p.noverlap);
  if (nseg == 0 || C == 0) return SPECGPU_OK;
  const int nfreq = plan->p.nperseg / 2 + 1;
  if (!X || ldf < nfreq) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "bad spectra buffer (ldf=%lld < nfreq=%d)", (long long)ldf, nfreq);
  StftArgs a = make_args(plan, x, n, ldx, 0, nseg, 1.0f, X, ldf, nullptr);
  CHECK_LAUNCH(ctx, launch_stft(plan->log2n, STFT_MODE_SPECTRA, a, C, (cudaStream_t)stream), "stft_kernel", 1);
  return SPECGPU_OK;
}

int specgpu_csd_pairs_block(specgpu_ctx* ctx, const specgpu_plan* plan, const float* X, int64_t C, int64_t nseg,
                            int64_t nseg_total, int64_t ldf, int64_t i0, int64_t ni, int32_t accumulate, float* P,
                            void* stream) {
  if (!ctx || !plan) return SPECGPU_ERR_INVALID_ARG;
  const int nfreq = plan->p.nperseg / 2 + 1;
  if (C < 0 || nseg < 0 || nseg_total < nseg || ldf < nfreq || i0 < 0 || ni < 0 || i0 + ni > C)
    return fail(ctx, SPECGPU_ERR_INVALID_ARG, "csd_pairs: bad shape C=%lld nseg=%lld/%lld ldf=%lld i0=%lld ni=%lld", (long long)C,
                (long long)nseg, (long long)nseg_total, (long long)ldf, (long long)i0, (long long)ni);
  if (C > 64) return fail(ctx, SPECGPU_ERR_UNSUPPORTED_SHAPE, "csd_pairs: C=%lld > 64 channels", (long long)C);
  if (C == 0 || ni == 0) return SPECGPU_OK;
  if (nseg == 0) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "csd_pairs: no segments to average");
  if (!X || !P) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "null pointer");
  DeviceGuard dev_guard(ctx->device);
  // the pair partials live behind whatever specgpu_csd_allpairs put in front (its spectra)
  const size_t need = ctx->ws_csd_off + csd_pairs_workspace_bytes(C, ni, nfreq, nseg) + 256;
  int rc = ensure_ws(ctx, need);
  if (rc) return rc;
  float* partial = reinterpret_cast<float*>(static_cast<char*>(ctx->ws) + ctx->ws_csd_off);
  CHECK_LAUNCH(ctx, launch_csd_pairs(X, C, nseg, nseg_total, ldf, nfreq, i0, ni, (float)plan->scale, accumulate ? 1 : 0, partial,
                                     P, (cudaStream_t)stream),
               "csd_pairs", 2);
  return SPECGPU_OK;
}

int specgpu_csd_spectra_blocked(specgpu_ctx* ctx, const specgpu_plan* plan, const float* x, int64_t C, int64_t n, int64_t ldx,
                                float* X, int32_t block_w, int32_t block_ld, void* stream) {
  int rc = check_signal_args(ctx, plan, x, C, n, ldx);
  if (rc) return rc;
  const int64_t nseg = num_segments(n, plan->p.nperseg, plan->p.noverlap);
  if (nseg == 0 || C == 0) return SPECGPU_OK;
  if (block_w < 1 || block_ld < block_w) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "bad frequency block (%d bins, pitch %d)", block_w, block_ld);
  if (!X) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "null spectra buffer");
  // block-major: X[nblocks][C][nseg][block_ld]; a row of one block is block_ld wide, blocks are whole planes apart
  StftArgs a = make_args(plan, x, n, ldx, 0, nseg, 1.0f, X, block_ld, nullptr);
  a.fblock_w = block_w;
  a.fblock_stride = C * nseg * (int64_t)block_ld;
  CHECK_LAUNCH(ctx, launch_stft(plan->log2n, STFT_MODE_SPECTRA, a, C, (cudaStream_t)stream), "stft_kernel", 1);
  return SPECGPU_OK;
}

int specgpu_csd_pairs_bins(specgpu_ctx* ctx, const specgpu_plan* plan, const float* X, int64_t C, int64_t nseg,
                           int64_t nseg_total, int64_t ldf, int64_t f0, int64_t nf, int32_t accumulate, float* P, void* stream) {
  if (!ctx || !plan) return SPECGPU_ERR_INVALID_ARG;
  const int nfreq = plan->p.nperseg / 2 + 1;
  if (C < 0 || nseg < 0 || nseg_total < nseg || f0 < 0 || nf < 0 || f0 + nf > nfreq || ldf < nf)
    return fail(ctx, SPECGPU_ERR_INVALID_ARG, "csd_pairs_bins: bad shape C=%lld nseg=%lld/%lld ldf=%lld bins [%lld, %lld) of %d", (long long)C,
                (long long)nseg, (long long)nseg_total, (long long)ldf, (long long)f0, (long long)(f0 + nf), nfreq);
  if (C > 64) return fail(ctx, SPECGPU_ERR_UNSUPPORTED_SHAPE, "csd_pairs_bins: C=%lld > 64 channels", (long long)C);
  if (C == 0 || nf == 0) return SPECGPU_OK;
  if (nseg == 0) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "csd_pairs_bins: no segments to average");
  if (!X || !P) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "null pointer");
  DeviceGuard dev_guard(ctx->device);
  int rc = ensure_ws(ctx, csd_pairs_workspace_bytes(C, C, (int)nf, nseg) + 256);
  if (rc) return rc;
  CHECK_LAUNCH(ctx, launch_csd_pairs(X, C, nseg, nseg_total, ldf, (int)nf, 0, C, (float)plan->scale, accumulate ? 1 : 0,
                                     static_cast<float*>(ctx->ws), P, (cudaStream_t)stream, (int)f0, nfreq),
               "csd_pairs", 2);
  return SPECGPU_OK;
}

int specgpu_csd_pairs(specgpu_ctx* ctx, const specgpu_plan* plan, const float* X, int64_t C, int64_t nseg, int64_t ldf,
                      int64_t i0, int64_t ni, float* P, void* stream) {
  return specgpu_csd_pairs_block(ctx, plan, X, C, nseg, nseg, ldf, i0, ni, 0, P, stream);
}

int specgpu_csd_frames(specgpu_ctx* ctx, const specgpu_plan* plan, const float* X, int64_t C, int64_t nseg, int64_t ldf,
                       int64_t i, int64_t j, int64_t seg_stride, int32_t navg, int64_t nframes, float* amp, void* stream) {
  if (!ctx || !plan) return SPECGPU_ERR_INVALID_ARG;
  const int nfreq = plan->p.nperseg / 2 + 1;
  if (C <= 0 || i < 0 || j < 0 || i >= C || j >= C || ldf < nfreq || navg < 1 || seg_stride < 1 || nframes < 0 ||
      (nframes > 0 && (nframes - 1) * seg_stride + navg > nseg))
    return fail(ctx, SPECGPU_ERR_INVALID_ARG, "csd_frames: frames (%lld x stride %lld + %d) exceed the %lld segments", (long long)nframes,
                (long long)seg_stride, navg, (long long)nseg);
  if (nframes == 0) return SPECGPU_OK;
  if (nframes > 65535) return fail(ctx, SPECGPU_ERR_UNSUPPORTED_SHAPE, "csd_frames: more than 65535 frames");
  if (!X || !amp) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "null pointer");
  CHECK_LAUNCH(ctx, launch_csd_frames(X, nseg, ldf, nfreq, (int)i, (int)j, seg_stride, navg, nframes, (float)plan->scale, amp,
                                      (cudaStream_t)stream), "csd_frames", 1);
  return SPECGPU_OK;
}

int specgpu_csd_allpairs(specgpu_ctx* ctx, const specgpu_plan* plan, const float* x, int64_t C, int64_t n, int64_t ldx,
                         float* P, void* stream) {
  int rc = check_signal_args(ctx, plan, x, C, n, ldx);
  if (rc) return rc;
  if (C > 64) return fail(ctx, SPECGPU_ERR_UNSUPPORTED_SHAPE, "csd_allpairs: C=%lld > 64 channels", (long long)C);
  const int64_t nseg = num_segments(n, plan->p.nperseg, plan->p.noverlap);
  if (C == 0) return SPECGPU_OK;
  if (nseg == 0) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "csd_allpairs: record shorter than nperseg");
  if (!P) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "null output pointer");
  const int nfreq = plan->p.nperseg / 2 + 1;
  const int64_t ldf = (nfreq + 15) & ~(int64_t)15;  // 128-byte aligned rows of float2
  DeviceGuard dev_guard(ctx->device);
  const size_t xbytes = (((size_t)C * nseg * ldf * 8) + 255) & ~(size_t)255;
  if ((rc = ensure_ws(ctx, xbytes + csd_pairs_workspace_bytes(C, C, nfreq, nseg) + 512))) return rc;
  float* X = static_cast<float*>(ctx->ws);
  if ((rc = specgpu_csd_spectra(ctx, plan, x, C, n, ldx, X, ldf, stream))) return rc;
  ctx->ws_csd_off = xbytes;
  rc = specgpu_csd_pairs(ctx, plan, X, C, nseg, ldf, 0, C, P, stream);
  ctx->ws_csd_off = 0;
  return rc;
}

// ---- whole path --------------------------------------------------------------------------------------
namespace {
// Channels per group of specgpu_pipeline: the log image of a group (and the two groups in flight) should sit in the
// 126 MB L2 between the STFT that writes it and the Gram / projection kernels that read it back.
int64_t pipeline_group_size(const specgpu_ctx* ctx, int64_t B, int64_t rows, int64_t ldt) {
  int64_t g = ctx->pipe_group;
  if (const char* env = std::getenv("SPECGPU_PIPELINE_GROUP")) g = std::atoll(env);
  (void)rows;
  (void)ldt;
  // Measured on B200 (profiles/r02_summary.md): groups on two streams are SLOWER than one batch-wide pass (40 channels:
  // 287 us as one group, 290 / 337 / 384 / 429 us in groups of 20 / 10 / 8 / 5) -- every kernel already fills the GPU,
  // the persistent kernels do not co-reside, and each group pays the ramp / tail and the latency-bound eigen step again.
  // So the default is ONE group; L2 reuse is obtained with eviction-priority hints instead (see l2_pin below).
  if (g <= 0) g = B;
  return std::min(g, B);
}
}  // namespace

int specgpu_pipeline(specgpu_ctx* ctx, const specgpu_plan* plan, const float* x, int64_t B, int64_t n, int64_t ldx, float* S,
                     float* D, int64_t ldt, int32_t flags, float* tiles, int32_t tile_w, int32_t ntiles, int32_t* info,
                     void* stream) {
  const int clip = (flags & SPECGPU_PIPE_CLIP) ? 1 : 0;
  // without `info` the caller cannot see a non-converged leading pair: the repair runs in the stream regardless
  const bool fallback = (flags & SPECGPU_PIPE_FALLBACK) != 0 || info == nullptr;
  int rc = check_signal_args(ctx, plan, x, B, n, ldx);
  if (rc) return rc;
  const int64_t nseg = num_segments(n, plan->p.nperseg, plan->p.noverlap);
  if (nseg == 0 || B == 0) return SPECGPU_OK;
  const int64_t rows = plan->p.nperseg / 2;
  if (!S || !D || ldt < nseg) return fail(ctx, SPECGPU_ERR_INVALID_ARG, "bad output (ldt=%lld < nseg=%lld)", (long long)ldt, (long long)nseg);
  if (rows > nseg) return fail(ctx, SPECGPU_ERR_UNSUPPORTED_SHAPE, "pipeline: %lld frequency rows > %lld segments", (long long)rows, (long long)nseg);
  if (rows > 512) return fail(ctx, SPECGPU_ERR_UNSUPPORTED_SHAPE, "pipeline: nperseg/2=%lld > 512 rows for the SVD stage", (long long)rows);
  if (tiles && (tile_w <= 0 || ntiles < 0 || (int64_t)tile_w * ntiles > nseg))
    return fail(ctx, SPECGPU_ERR_INVALID_ARG, "pipeline: tile_w*ntiles exceeds the %lld columns", (long long)nseg);
  DeviceGuard dev_guard(ctx->device);
  const bool power_ok = rows <= 256;
  const bool tc = power_ok && gram_tc_supported(rows);
  const int64_t gsz = pipeline_group_size(ctx, B, rows, ldt);
  const int64_t ngroups = ceil_div(B, gsz);
  // groups alternate between two side streams, or (SPECGPU_PIPELINE_LANES=1) follow each other in the caller's stream,
  // where the programmatic dependent launches chain them without a gap
  static const bool one_lane = std::getenv("SPECGPU_PIPELINE_LANES") && std::getenv("SPECGPU_PIPELINE_LANES")[0] == '1';
  const int nlanes = (ngroups > 1 && !one_lane) ? 2 : 1;        // streams in use
  // ---- workspace: min/max pairs and per-channel SVD arrays for the whole batch; Gram partials and Jacobi scratch per lane ----
  const size_t part_bytes = tc ? gram_tc_workspace_bytes(gsz, rows) : 0;
  const size_t jac_bytes = jacobi_workspace_bytes(gsz, (int)rows);
  // On the tensor-core route the raw log image goes to a TILED scratch buffer ([rows x 32] column tiles stored
  // contiguously, see img_off): the STFT's tile stores, the Gram kernel's box loads and the projection's tile reads are
  // then contiguous 16-32 KB blocks instead of 64-128 byte pieces 15.7 KB apart.  S is written once, by the projection.
  const int64_t ntile = ceil_div(nseg, kTileCols);
  const bool tiled = tc && plan->log2n <= 9 && !std::getenv("SPECGPU_NO_TILED");
  const size_t scratch_bytes = tiled ? (size_t)B * ntile * rows * kTileCols * sizeof(float) : 0;
  if ((rc = ensure_ws(ctx, carve_size({(size_t)B * rows * rows * 8, (size_t)B * rows * rows * 4,
                                       (size_t)B * rows * 4, (size_t)B * 16, part_bytes, part_bytes, jac_bytes, jac_bytes,
                                       scratch_bytes, (size_t)B * 4}))))
    return rc;
  MinMaxWord* mm = nullptr;
  unsigned gen = 0;
  if ((rc = minmax_begin(ctx, &mm, &gen))) return rc;
  Carver cv(ctx->ws);
  double* Gall = cv.take<double>(B * rows * rows);
  float* Uall = cv.take<float>(B * rows * rows);
  float* lam_all = cv.take<float>(B * rows);
  int32_t* plan_all = cv.take<int32_t>(B * 4);
  float* part[2];
  void* jac[2];
  for (int i = 0; i < 2; ++i) part[i] = tc ? cv.take<float>(part_bytes / 4) : nullptr;
  for (int i = 0; i < 2; ++i) jac[i] = cv.take<char>(jac_bytes);
  float* Lt = tiled ? cv.take<float>(scratch_bytes / sizeof(float)) : nullptr;
  int32_t* flagged_all = cv.take<int32_t>(B);
#ifndef SPECGPU_EMULATE
  // (Measured on B200: the side-stream variant is no faster than the in-line repair -- 288.1 vs 287.9 us per 40-channel shot,
  // 278.6 without any repair: the projection's grid leaves the repair kernels no SMs to overlap on -- so it is opt-in.)
  if (tiled && fallback && !ctx->repair_stream && std::getenv("SPECGPU_SIDE_REPAIR")) {
    // highest priority: the repair kernels (which return at once unless a channel is flagged) take SMs as the projection's
    // short CTAs retire instead of queueing behind its whole grid
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&ctx->repair_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_repair_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_repair_join, cudaEventDisableTiming) != cudaSuccess)
      return cuda_fail(ctx, (int)cudaGetLastError(), "repair stream");
  }
#endif

  cudaStream_t user = (cudaStream_t)stream;
  cudaStream_t lane[2] = {user, user};
#ifndef SPECGPU_EMULATE
  if (nlanes > 1) {
    for (int i = 0; i < 2; ++i) {
      if (!ctx->side[i] && cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking) != cudaSuccess)
        return cuda_fail(ctx, (int)cudaGetLastError(), "side stream");
      if (!ctx->ev_join[i] && cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming) != cudaSuccess)
        return cuda_fail(ctx, (int)cudaGetLastError(), "join event");
      lane[i] = ctx->side[i];
    }
    if (!ctx->ev_fork && cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess)
      return cuda_fail(ctx, (int)cudaGetLastError(), "fork event");
  }
#endif
#ifndef SPECGPU_EMULATE
  if (nlanes > 1) {
    cudaEventRecord(ctx->ev_fork, user);
    cudaStreamWaitEvent(lane[0], ctx->ev_fork, 0);
    cudaStreamWaitEvent(lane[1], ctx->ev_fork, 0);
  }
#endif
  // The log image of a group is written by the STFT and read back twice (Gram, projection).  When it is larger than what
  // the L2 can hold next to the streams passing through, keep an address-hashed fraction of its lines (evict_last) and
  // let the rest stream, instead of letting an LRU sweep evict every line just before it is reused.
  float l2_pin = 1.0f;
  {
    double pin_mb = 64.0;
    if (const char* env = std::getenv("SPECGPU_L2_PIN_MB")) pin_mb = std::atof(env);
    const double img_mb = (double)gsz * rows * ldt * 4.0 / (1024.0 * 1024.0) * nlanes;
    l2_pin = pin_mb <= 0.0 ? 0.f : (float)std::min(1.0, pin_mb / std::max(img_mb, 1e-9));
  }
  for (int64_t g = 0; g < ngroups; ++g) {
    const int64_t b0 = g * gsz, nb = std::min(gsz, B - b0);
    const int li = (int)(g % nlanes);
    void* stream = (void*)lane[li];            // CHECK_LAUNCH / ProfScope take the stream from this name
    cudaStream_t st = lane[li];
    float* Sg = S + b0 * rows * ldt;
    float* Dg = D + b0 * rows * ldt;
    MinMaxWord* mmg = mm + 2 * b0;
    float* Lg = tiled ? Lt + (size_t)b0 * ntile * rows * kTileCols : nullptr;
    StftArgs a = make_args(plan, x + b0 * ldx, n, ldx, 0, nseg, (float)plan->scale, tiled ? Lg : Sg, tiled ? -ntile : ldt, mmg, gen);
    a.l2_pin = l2_pin;
    // dynamic tile walk of the STFT (one launch at a time per counter: not with channel groups on two lanes)
    static const bool dyn_env = !(std::getenv("SPECGPU_STFT_DYNAMIC") && std::getenv("SPECGPU_STFT_DYNAMIC")[0] == '0');
    a.dyn = (ngroups == 1 && dyn_env && !(flags & SPECGPU_PIPE_STATIC_TILES)) ? reinterpret_cast<unsigned*>(ctx->mm64 + kMaxBatch * 2) : nullptr;
    int64_t pre_gram[2] = {0, 0};
    bool fused = false;
    if (tiled && stft_gram_supported(plan->log2n) && std::getenv("SPECGPU_FUSED_GRAM")) {
      int e__;
      {
        ProfScope prof__(ctx, st, "stft_gram");
        e__ = launch_stft_gram(a, nb, part[li], ctx->num_sms, &pre_gram[0], &pre_gram[1], st);
      }
      if (e__ == 0) {
        fused = true;
        ctx->launches += 1;
      } else if (e__ != 1) {
        return cuda_fail(ctx, e__, "stft_gram");
      }
    }
#ifndef SPECGPU_EMULATE
    if (g == 0 && ctx->ilk_wait) cudaStreamWaitEvent(st, ctx->ilk_wait, 0);      // specgpu_set_pipeline_interlock
#endif
    if (!fused) CHECK_LAUNCH(ctx, launch_stft(plan->log2n, STFT_MODE_LOGPSD, a, nb, st), "stft_kernel", 1);
    SvdWs w{};
    w.G = reinterpret_cast<float*>(Gall + b0 * rows * rows);
    w.U = Uall + b0 * rows * rows;
    w.lam = lam_all + b0 * rows;
    w.plan = plan_all + b0 * 4;
    w.gram_partial = part[li];
    w.jacobi = jac[li];
    w.flagged = flagged_all + b0;
    // Sg holds the raw log image until the rank-1 projection normalises it in place (see svd_run)
    // the rank-1 projection (rows <= 256) writes the tiles itself; the general projection leaves them to the tile cut
    const bool r1_writes_tiles = tiles && ntiles > 0 && power_ok && !std::getenv("SPECGPU_NO_FUSED_TILES");
    R1Tiles r1t;
    if (r1_writes_tiles) {
      r1t.ptr = tiles + (size_t)b0 * ntiles * rows * tile_w;
      r1t.tile_w = tile_w;
      r1t.ntiles = ntiles;
    }
    ctx->ilk_armed = (g == ngroups - 1);
    rc = svd_run(ctx, w, Sg, mmg, nb, rows, nseg, ldt, 0, 1, (int)rows, clip, power_ok, fallback, Dg, 0, ldt, nullptr,
                 info ? info + b0 * 4 : nullptr, st, l2_pin, Lg, tiled ? -ntile : 0, fused ? pre_gram : nullptr, r1t);
    ctx->ilk_armed = false;
    if (rc) return rc;
    if (tiles && ntiles > 0 && !r1_writes_tiles)
      CHECK_LAUNCH(ctx, launch_patch(Dg, nb, rows, ldt, tile_w, ntiles, tiles + (size_t)b0 * ntiles * rows * tile_w, 0, st), "patch", 1);
  }
#ifndef SPECGPU_EMULATE
  if (nlanes > 1) {
    for (int i = 0; i < 2; ++i) {
      cudaEventRecord(ctx->ev_join[i], lane[i]);
      cudaStreamWaitEvent(user, ctx->ev_join[i], 0);
    }
  }
#endif
  return SPECGPU_OK;
}

}  // extern "C"
