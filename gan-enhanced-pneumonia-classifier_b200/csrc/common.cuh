// Shared helpers for libb200gan.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

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

constexpr int kNumSMs = 148;   // B200

}  // namespace b200gan
