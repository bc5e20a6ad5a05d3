// Shared helpers for libb200gan.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/b200gan.h"

namespace b200gan {

// thread-local error message returned by b200gan_last_error_string()
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);

#define B200_CHECK_ARG(cond, ...)                       \
  do {                                                  \
    if (!(cond)) {                                      \
      ::b200gan::set_error(__VA_ARGS__);                \
      return B200GAN_ERR_BAD_ARG;                       \
    }                                                   \
  } while (0)

#define B200_UNSUPPORTED(...)                           \
  do {                                                  \
    ::b200gan::set_error(__VA_ARGS__);                  \
    return B200GAN_ERR_UNSUPPORTED;                     \
  } while (0)

#define B200_CUDA(expr)                                                     \
  do {                                                                      \
    cudaError_t e__ = (expr);                                               \
    if (e__ != cudaSuccess) return ::b200gan::cuda_fail(e__, #expr);        \
  } while (0)

#define B200_LAUNCH_CHECK(name)                                             \
  do {                                                                      \
    cudaError_t e__ = cudaGetLastError();                                   \
    if (e__ != cudaSuccess) return ::b200gan::cuda_fail(e__, name);         \
  } while (0)

// device-side mirror of b200gan_view with typed access
struct View {
  void*   ptr;
  int32_t dtype;
  int32_t n, h, w, c;
  int64_t sn, sh, sw, sc;
};

static inline View to_view(const b200gan_view* v) {
  View r;
  r.ptr = v->ptr; r.dtype = v->dtype; r.n = v->n; r.h = v->h; r.w = v->w; r.c = v->c;
  r.sn = v->sn; r.sh = v->sh; r.sw = v->sw; r.sc = v->sc;
  return r;
}

static inline int check_view(const b200gan_view* v, const char* name) {
  if (!v || !v->ptr) { set_error("%s: null view/pointer", name); return B200GAN_ERR_BAD_ARG; }
  if (v->dtype != B200GAN_F32 && v->dtype != B200GAN_BF16) { set_error("%s: bad dtype %d", name, v->dtype); return B200GAN_ERR_BAD_ARG; }
  if (v->n <= 0 || v->h <= 0 || v->w <= 0 || v->c <= 0) { set_error("%s: empty view (%d,%d,%d,%d)", name, v->n, v->h, v->w, v->c); return B200GAN_ERR_BAD_ARG; }
  return 0;
}

template <typename T> __device__ __forceinline__ float ld_as_float(const T* p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T> __device__ __forceinline__ void st_from_float(T* p, float v);
template <> __device__ __forceinline__ void st_from_float<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_from_float<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// SM count of the CURRENT device (B200: 148), read from the driver once per device
int num_sms();
#define kNumSMs (::b200gan::num_sms())

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device, per-function attribute: raise it at most once per (device, kernel,
// size), from any host thread (the C ABI promises re-entrancy per device; a plain `static int` guard was neither).
template <auto Kernel>                                       // the kernel FUNCTION is the template argument: one guard array per kernel
static inline cudaError_t ensure_dynamic_smem(int smem) {
  constexpr int kMaxDevices = 64;
  static std::atomic<int> configured[kMaxDevices];          // zero-initialised
  const auto kernel = Kernel;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (configured[dev].load(std::memory_order_acquire) >= smem) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);      // idempotent: a race only repeats the call
  if (e != cudaSuccess) return e;
  int cur = configured[dev].load(std::memory_order_relaxed);
  while (cur < smem && !configured[dev].compare_exchange_weak(cur, smem, std::memory_order_release)) {}
  return cudaSuccess;
}

// 8 bf16 <-> 8 floats through one 16-byte vector
__device__ __forceinline__ void unpack8(const uint4& t, float* v) {
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&b); }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// act'(.) expressed through the SAVED OUTPUT a of the activation (what the forward pass keeps): LeakyReLU and ReLU
// preserve the sign of their argument, tanh' = 1 - a^2, sigmoid' = a (1 - a).
__device__ __forceinline__ float act_grad_from_output(float a, int act, float slope) {
  switch (act) {
    case B200GAN_ACT_RELU: return a > 0.f ? 1.f : 0.f;
    case B200GAN_ACT_LRELU: return a > 0.f ? 1.f : slope;
    case B200GAN_ACT_TANH: return 1.f - a * a;
    case B200GAN_ACT_SIGMOID: return a * (1.f - a);
    default: return 1.f;
  }
}
__device__ __forceinline__ float act_apply(float z, int act, float slope) {
  switch (act) {
    case B200GAN_ACT_RELU: return z > 0.f ? z : 0.f;
    case B200GAN_ACT_LRELU: return z > 0.f ? z : z * slope;
    case B200GAN_ACT_TANH: return tanhf(z);
    case B200GAN_ACT_SIGMOID: return 1.f / (1.f + expf(-z));
    default: return z;
  }
}
// one element of a view whose dtype is only known at run time
__device__ __forceinline__ float ld_rt(const void* base, int dtype, int64_t off) {
  return dtype == B200GAN_F32 ? reinterpret_cast<const float*>(base)[off] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[off]);
}

struct TcEpi {                       // host-side description of the requested epilogue fusion
  int mode = 0;
  double* sums = nullptr;
  const b200gan_view* prev_y = nullptr;
  const float *scale = nullptr, *shift = nullptr, *mean = nullptr, *invstd = nullptr;
  int act = 0; float slope = 0.f;
};

// Fusions requested around one convolution call, resolved from b200gan_fuse by the dispatcher (api.cu).
struct ConvFuse {
  int out_act = B200GAN_ACT_NONE; float out_slope = 0.f;        // activation on the result
  const b200gan_view* g_ref = nullptr; int g_act = B200GAN_ACT_NONE; float g_slope = 0.f;   // gradient operand *= act'(g_ref)
};

}  // namespace b200gan
