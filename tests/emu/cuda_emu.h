// Minimal CPU emulation of the CUDA execution model -- TEST INFRASTRUCTURE ONLY.
//
// The build container has nvcc but no GPU.  To debug kernel index arithmetic before spending GPU
// minutes, the product kernels (spectrogram_enhancement_b200/csrc/*.cu) can be compiled as plain
// C++ with -DSPECGPU_EMULATE: every CUDA thread of a block becomes an OS thread, __syncthreads()
// is a std::barrier, warp shuffles go through a per-warp exchange slot, thread-block clusters run
// their blocks concurrently with DSMEM mapped to the peer block's buffer.  The result is
// libspecgpu_emu.so, used only by the CPU tests (tests/test_emulated.py); the product package
// never loads it and has no CPU path.  tcgen05/TMA inline PTX is not emulated: those kernels
// carry a scalar stand-in under #ifdef SPECGPU_EMULATE that reproduces their numerics (TF32
// operand rounding) so that the surrounding pipeline can still be exercised.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __constant__ static

struct dim3 {
  unsigned x, y, z;
  constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) double2 { double x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributeNonPortableClusterSizeAllowed = 10 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97, cudaDevAttrComputeCapabilityMajor = 75, cudaDevAttrComputeCapabilityMinor = 76 };

namespace emu {

struct WarpState {
  std::unique_ptr<std::barrier<>> bar;
  uint64_t slot[32];
  int nlanes = 32;
};

struct BlockState {
  std::unique_ptr<std::barrier<>> bar;
  std::vector<WarpState> warps;
  unsigned char* dyn_smem = nullptr;
  size_t dyn_bytes = 0;
  unsigned cluster_rank = 0;
};

struct ClusterState {
  std::unique_ptr<std::barrier<>> bar;
  std::vector<BlockState*> blocks;
};

struct ThreadCtx {
  BlockState* blk = nullptr;
  ClusterState* cl = nullptr;
  int warp = 0, lane = 0;
};

extern thread_local ThreadCtx tctx;
extern thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;

inline void init_block(BlockState& bs, dim3 block, size_t smem) {
  unsigned nthreads = block.x * block.y * block.z;
  bs.bar = std::make_unique<std::barrier<>>(nthreads);
  unsigned nw = (nthreads + 31) / 32;
  bs.warps.resize(nw);
  for (unsigned w = 0; w < nw; ++w) {
    int n = std::min(32u, nthreads - w * 32);
    bs.warps[w].nlanes = n;
    bs.warps[w].bar = std::make_unique<std::barrier<>>(n);
  }
  bs.dyn_bytes = smem;
  bs.dyn_smem = smem ? static_cast<unsigned char*>(std::aligned_alloc(1024, (smem + 1023) / 1024 * 1024)) : nullptr;
  if (bs.dyn_smem) std::memset(bs.dyn_smem, 0xCD, smem);  // poison: uninitialised smem reads show up
}

// Launch: one OS thread per CUDA thread of a block (x cluster size); the same threads walk the
// blocks (or clusters of `cluster` consecutive blocks) of the grid one after another.
template <class K, class... Args>
void launch(K kernel, dim3 grid, dim3 block, size_t smem, unsigned cluster, Args... args) {
  if (cluster == 0) cluster = 1;
  const unsigned nblocks = grid.x * grid.y * grid.z;
  const unsigned nthreads = block.x * block.y * block.z;
  if (nblocks == 0 || nthreads == 0) return;
  ClusterState cl;
  cl.bar = std::make_unique<std::barrier<>>(nthreads * cluster);
  std::barrier<> end_bar(nthreads * cluster);
  std::vector<BlockState> bss(cluster);
  for (unsigned c = 0; c < cluster; ++c) {
    init_block(bss[c], block, smem);
    bss[c].cluster_rank = c;
    cl.blocks.push_back(&bss[c]);
  }
  std::vector<std::thread> pool;
  pool.reserve(nthreads * cluster);
  for (unsigned w = 0; w < nthreads * cluster; ++w) {
    pool.emplace_back([&, w]() {
      const unsigned c = w / nthreads, t = w % nthreads;
      t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
      t_blockDim = block;
      t_gridDim = grid;
      tctx.blk = &bss[c];
      tctx.cl = &cl;
      tctx.warp = t / 32;
      tctx.lane = t % 32;
      for (unsigned b0 = 0; b0 < nblocks; b0 += cluster) {
        const unsigned b = b0 + c;
        t_blockIdx = dim3(b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y));
        kernel(args...);
        end_bar.arrive_and_wait();
      }
    });
  }
  for (auto& th : pool) th.join();
  for (auto& bs : bss) std::free(bs.dyn_smem);
}

template <class T>
inline uint64_t to_bits(T v) {
  static_assert(sizeof(T) <= 8, "shuffle payload too large");
  uint64_t b = 0;
  std::memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T>
inline T from_bits(uint64_t b) {
  T v;
  std::memcpy(&v, &b, sizeof(T));
  return v;
}
template <class T>
inline T shfl_generic(T v, int src_lane) {
  WarpState& w = tctx.blk->warps[tctx.warp];
  w.slot[tctx.lane] = to_bits(v);
  w.bar->arrive_and_wait();
  T r = (src_lane >= 0 && src_lane < w.nlanes) ? from_bits<T>(w.slot[src_lane]) : v;
  w.bar->arrive_and_wait();
  return r;
}
}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)
#define warpSize 32

static inline void __syncthreads() { emu::tctx.blk->bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::tctx.blk->warps[emu::tctx.warp].bar->arrive_and_wait(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  int lane = emu::tctx.lane;
  int base = lane & ~(width - 1);
  return emu::shfl_generic(v, base + (src & (width - 1)));
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int mask, int width = 32) {
  (void)width;
  return emu::shfl_generic(v, emu::tctx.lane ^ mask);
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32) {
  int lane = emu::tctx.lane;
  int src = lane + (int)delta;
  if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
  return emu::shfl_generic(v, src);
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32) {
  int lane = emu::tctx.lane;
  int src = lane - (int)delta;
  if (src < (lane & ~(width - 1))) src = lane;
  return emu::shfl_generic(v, src);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned bit = pred ? (1u << emu::tctx.lane) : 0u;
  unsigned r = 0;
  emu::WarpState& w = emu::tctx.blk->warps[emu::tctx.warp];
  w.slot[emu::tctx.lane] = bit;
  w.bar->arrive_and_wait();
  for (int i = 0; i < w.nlanes; ++i) r |= (unsigned)w.slot[i];
  w.bar->arrive_and_wait();
  return r;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
  emu::WarpState& w = emu::tctx.blk->warps[emu::tctx.warp];
  w.slot[emu::tctx.lane] = v;
  w.bar->arrive_and_wait();
  unsigned r = 0;
  for (int i = 0; i < w.nlanes; ++i) r = std::max(r, (unsigned)w.slot[i]);
  w.bar->arrive_and_wait();
  return r;
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v) {
  emu::WarpState& w = emu::tctx.blk->warps[emu::tctx.warp];
  w.slot[emu::tctx.lane] = v;
  w.bar->arrive_and_wait();
  unsigned r = 0xffffffffu;
  for (int i = 0; i < w.nlanes; ++i) r = std::min(r, (unsigned)w.slot[i]);
  w.bar->arrive_and_wait();
  return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) {
  unsigned full = emu::tctx.blk->warps[emu::tctx.warp].nlanes == 32 ? 0xffffffffu
                                                                    : ((1u << emu::tctx.blk->warps[emu::tctx.warp].nlanes) - 1);
  return __ballot_sync(m, pred) == full;
}

// ---- atomics -------------------------------------------------------------------------------------
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline float atomicAdd(float* p, float v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(p);
  uint32_t old = __atomic_load_n(u, __ATOMIC_SEQ_CST), nw;
  float f;
  do {
    std::memcpy(&f, &old, 4);
    f += v;
    std::memcpy(&nw, &f, 4);
  } while (!__atomic_compare_exchange_n(u, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  std::memcpy(&f, &old, 4);
  return f;
}
static inline double atomicAdd(double* p, double v) {
  uint64_t* u = reinterpret_cast<uint64_t*>(p);
  uint64_t old = __atomic_load_n(u, __ATOMIC_SEQ_CST), nw;
  double f;
  do {
    std::memcpy(&f, &old, 8);
    f += v;
    std::memcpy(&nw, &f, 8);
  } while (!__atomic_compare_exchange_n(u, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  std::memcpy(&f, &old, 8);
  return f;
}
static inline int atomicMax(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline int atomicMin(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline unsigned atomicMax(unsigned* p, unsigned v) {
  unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline unsigned atomicMin(unsigned* p, unsigned v) {
  unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned v) {
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;
}
static inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }

// ---- intrinsics ----------------------------------------------------------------------------------
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
  const unsigned long long v = ((unsigned long long)hi << 32) | lo;
  return (unsigned)(v >> (shift & 31));
}
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
  for (int s = 0; s < 4; ++s) c += ((a >> (8 * s)) & 255u) * ((b >> (8 * s)) & 255u);
  return c;
}
static inline unsigned __vmaxu4(unsigned a, unsigned b) {
  unsigned r = 0;
  for (int s = 0; s < 4; ++s) {
    const unsigned x = (a >> (8 * s)) & 255u, y = (b >> (8 * s)) & 255u;
    r |= (x > y ? x : y) << (8 * s);
  }
  return r;
}
static inline unsigned __vmaxs2(unsigned a, unsigned b) {
  const int16_t a0 = (int16_t)a, a1 = (int16_t)(a >> 16), b0 = (int16_t)b, b1 = (int16_t)(b >> 16);
  return (unsigned)(uint16_t)(a0 > b0 ? a0 : b0) | ((unsigned)(uint16_t)(a1 > b1 ? a1 : b1) << 16);
}
static inline unsigned __vmins2(unsigned a, unsigned b) {
  const int16_t a0 = (int16_t)a, a1 = (int16_t)(a >> 16), b0 = (int16_t)b, b1 = (int16_t)(b >> 16);
  return (unsigned)(uint16_t)(a0 < b0 ? a0 : b0) | ((unsigned)(uint16_t)(a1 < b1 ? a1 : b1) << 16);
}
static inline unsigned __vimax3_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vmaxs2(a, b), c); }
static inline unsigned __vimin3_s16x2(unsigned a, unsigned b, unsigned c) { return __vmins2(__vmins2(a, b), c); }
static inline unsigned __vminu4(unsigned a, unsigned b) {
  unsigned r = 0;
  for (int s = 0; s < 4; ++s) {
    const unsigned x = (a >> (8 * s)) & 255u, y = (b >> (8 * s)) & 255u;
    r |= (x < y ? x : y) << (8 * s);
  }
  return r;
}
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
// packed FP32 pairs (FADD2 / FMUL2 / FFMA2 on sm_100a): lane-wise, one rounding per lane
static inline float2 __fadd2_rn(float2 a, float2 b) { volatile float x = a.x + b.x, y = a.y + b.y; return make_float2(x, y); }
static inline float2 __fmul2_rn(float2 a, float2 b) { volatile float x = a.x * b.x, y = a.y * b.y; return make_float2(x, y); }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)); }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline long long __double_as_longlong(double d) { long long i; std::memcpy(&i, &d, 8); return i; }
static inline double __longlong_as_double(long long i) { double d; std::memcpy(&d, &i, 8); return d; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float emu_fast_logf(float a) { return std::log(a); }
static inline float emu_fast_expf(float a) { return std::exp(a); }
static inline float emu_fast_log2f(float a) { return std::log2(a); }
#define __log2f emu_fast_log2f
#define __logf emu_fast_logf
#define __expf emu_fast_expf
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
static inline double rsqrt(double a) { return 1.0 / std::sqrt(a); }
static inline float __saturatef(float a) { return a < 0.f ? 0.f : (a > 1.f ? 1.f : a); }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __brev(unsigned v) {
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
  v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
  return (v >> 16) | (v << 16);
}
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline long min(long a, long b) { return a < b ? a : b; }
static inline long max(long a, long b) { return a > b ? a : b; }
static inline float min(float a, float b) { return a < b ? a : b; }
static inline float max(float a, float b) { return a > b ? a : b; }

// ---- clusters / DSMEM ------------------------------------------------------------------------------
namespace emu {
inline unsigned cluster_ctarank() { return tctx.blk->cluster_rank; }
inline unsigned cluster_nctarank() { return (unsigned)tctx.cl->blocks.size(); }
inline void cluster_sync() { tctx.cl->bar->arrive_and_wait(); }
// pointer into this block's dynamic smem -> same offset in block `rank` of the cluster
template <class T>
inline T* map_shared_rank(T* p, unsigned rank) {
  unsigned char* me = tctx.blk->dyn_smem;
  size_t off = reinterpret_cast<unsigned char*>(p) - me;
  return reinterpret_cast<T*>(tctx.cl->blocks[rank]->dyn_smem + off);
}
inline unsigned char* dyn_smem() { return tctx.blk->dyn_smem; }
}  // namespace emu

// ---- runtime API stubs ---------------------------------------------------------------------------
static inline cudaError_t cudaMalloc(void** p, size_t n) {
  *p = std::aligned_alloc(256, (n + 255) / 256 * 256 + 256);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
  switch (a) {
    case cudaDevAttrMultiProcessorCount: *v = 148; break;
    case cudaDevAttrMaxSharedMemoryPerBlockOptin: *v = 232448; break;
    case cudaDevAttrComputeCapabilityMajor: *v = 10; break;
    case cudaDevAttrComputeCapabilityMinor: *v = 0; break;
    default: *v = 0;
  }
  return cudaSuccess;
}
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
template <class K> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) { *n = 2; return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
