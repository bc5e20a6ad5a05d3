// BatchNorm2d (train/eval) statistics + apply, activations, their backward, Sigmoid+BCE, Adam and the
// layout copy: the HBM-bound kernels of the DCGAN step.  fp32 math throughout, f32 or bf16 storage.
// Reference semantics: torch native_batch_norm(+backward), relu/leaky_relu/tanh/sigmoid,
// binary_cross_entropy and optim.Adam as called from dcgan.py:27-47,66-85 and train_gan.py:90-95,128-150.
#include <math.h>

#include <cstdlib>
#include "common.cuh"

namespace b200gan {

__device__ __forceinline__ float act_fwd(float z, int act, float slope) {
  switch (act) {
    case B200GAN_ACT_RELU: return z > 0.f ? z : 0.f;
    case B200GAN_ACT_LRELU: return z > 0.f ? z : z * slope;
    case B200GAN_ACT_TANH: return tanhf(z);
    case B200GAN_ACT_SIGMOID: return 1.f / (1.f + expf(-z));
    default: return z;
  }
}
// derivative of the activation given pre-activation z and (for tanh / sigmoid) the saved output a
__device__ __forceinline__ float act_grad(float z, float a, int act, float slope) {
  switch (act) {
    case B200GAN_ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case B200GAN_ACT_LRELU: return z > 0.f ? 1.f : slope;
    case B200GAN_ACT_TANH: return 1.f - a * a;
    case B200GAN_ACT_SIGMOID: return (1.f - a) * a;
    default: return 1.f;
  }
}

__device__ __forceinline__ int64_t pix_offset(const View& v, int64_t pix) {
  const int w = (int)(pix % v.w);
  const int64_t t = pix / v.w;
  const int h = (int)(t % v.h);
  const int64_t n = t / v.h;
  return n * v.sn + (int64_t)h * v.sh + (int64_t)w * v.sw;
}


// ================================================================================================
// Fast paths: dense NHWC tensors ([P][C] contiguous, C % VEC == 0) processed as 16-byte vectors with the
// per-channel coefficients hoisted into registers.  These are the kernels the training step actually runs;
// the strided scalar kernels below remain for the reference's NCHW edge tensors and odd channel counts.
// ================================================================================================
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  using Raw = float4;
  __device__ static Raw load_raw(const float* p) { return *reinterpret_cast<const float4*>(p); }
  __device__ static void unpack(const Raw& t, float* v) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  __device__ static void load(const float* p, float* v) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  __device__ static void store(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  using Raw = uint4;
  __device__ static Raw load_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  __device__ static void unpack(const Raw& t, float* v) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
  __device__ static void load(const __nv_bfloat16* p, float* v) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
  __device__ static void store(__nv_bfloat16* p, const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&b); }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

static bool dense_nhwc(const b200gan_view* v) {
  return v->sc == 1 && v->sw == v->c && v->sh == (int64_t)v->w * v->c && v->sn == (int64_t)v->h * v->w * v->c &&
         (reinterpret_cast<uintptr_t>(v->ptr) & 15) == 0;
}

template <typename T, int U = 4>
__global__ void __launch_bounds__(256) bn_act_fwd_dense_kernel(const T* __restrict__ y, T* __restrict__ a, int64_t nvec, int C,
                                                               const float* __restrict__ scale, const float* __restrict__ shift,
                                                               int act, float slope) {
  constexpr int V = Vec<T>::N;
  // chunks of 256*U consecutive vectors per CTA and iteration (see bn_act_bwd_apply_dense_kernel)
  const bool hoist = (256 * V) % C == 0;
  float sc[V], sh[V];
  int c0 = (int)(((int64_t)threadIdx.x * V) % C);
#pragma unroll
  for (int j = 0; j < V; ++j) { sc[j] = scale ? scale[c0 + j] : 1.f; sh[j] = scale ? shift[c0 + j] : 0.f; }
  using Raw = typename Vec<T>::Raw;
  constexpr int64_t CH = 256 * U;
  for (int64_t base = (int64_t)blockIdx.x * CH + threadIdx.x; base < nvec; base += (int64_t)gridDim.x * CH) {
    Raw r[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (base + u * 256 < nvec) r[u] = Vec<T>::load_raw(y + (base + u * 256) * V);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t k = base + u * 256;
      if (k >= nvec) break;
      if (!hoist && scale) {
        c0 = (int)((k * V) % C);
#pragma unroll
        for (int j = 0; j < V; ++j) { sc[j] = scale[c0 + j]; sh[j] = shift[c0 + j]; }
      }
      float v[V];
      Vec<T>::unpack(r[u], v);
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = act_fwd(fmaf(v[j], sc[j], sh[j]), act, slope);
      Vec<T>::store(a + k * V, v);
    }
  }
}

struct BwdDenseArgs {
  const void *da, *y, *a;
  void* dy;
  int64_t nvec;
  int C;
  const float *scale, *shift, *mean, *invstd, *gamma;
  const double* sums;
  double count;
  int act; float slope;
  float *dgamma, *dbeta;
};

// dy = k1*dz + p*y + q  with k1 = gamma*invstd, p = -k1*m2*invstd, q = k1*(m2*invstd*mean - m1)
// (algebraically gamma*invstd*(dz - m1 - xhat*m2)); no BatchNorm: dy = dz.
// HAS_ACT = false: da already is dz (the activation backward was applied by the producing convolution's epilogue,
// b200gan_fuse.prev_*): no activation coefficients are held, which keeps the kernel at <= 64 registers (4 CTAs per SM).
template <typename T, bool HAS_ACT>
__global__ void __launch_bounds__(256, 3) bn_act_bwd_apply_dense_kernel(BwdDenseArgs g) {
  constexpr int V = Vec<T>::N;
  const int C = g.C;
  if (g.scale && blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (g.dbeta) g.dbeta[c] += (float)g.sums[c];
      if (g.dgamma) g.dgamma[c] += (float)g.sums[C + c];
    }
  }
  // a CTA walks chunks of 256*U consecutive vectors (grid-stride over chunks): a warp's U loads of one tensor cover U consecutive
  // 512-byte segments, and a thread keeps the same V channels as long as 256*V is a multiple of C
  const bool hoist = (256 * V) % C == 0;
  float sc[V], sh[V], k1[V], pp[V], qq[V];
  auto coeffs = [&](int c0) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int c = c0 + j;
      if (g.scale) {
        const float is = g.invstd[c], mu = g.mean[c];
        const float m1 = (float)(g.sums[c] / g.count), m2 = (float)(g.sums[C + c] / g.count);
        if (HAS_ACT) { sc[j] = g.scale[c]; sh[j] = g.shift[c]; }
        k1[j] = g.gamma[c] * is;
        pp[j] = -k1[j] * m2 * is;
        qq[j] = k1[j] * (m2 * is * mu - m1);
      } else {
        if (HAS_ACT) { sc[j] = 1.f; sh[j] = 0.f; }
        k1[j] = 1.f; pp[j] = 0.f; qq[j] = 0.f;
      }
    }
  };
  coeffs((int)(((int64_t)threadIdx.x * V) % C));
  const T* da = reinterpret_cast<const T*>(g.da);
  const T* y = reinterpret_cast<const T*>(g.y);
  const T* a = reinterpret_cast<const T*>(g.a);
  T* dy = reinterpret_cast<T*>(g.dy);
  // all loads of a chunk are issued before its first store and kept RAW (packed) until they are used.  Measured on D1's tensor
  // (tools/ew_bench.py, one resident wave): U = 1 / 2 / 3 / 4 -> 126 / 110 / 114 / 123 us; U = 2 is 5.6 TB/s = 86 % of the copy peak
  constexpr int U = 2;
  using Raw = typename Vec<T>::Raw;
  constexpr int64_t CH = 256 * U;
  for (int64_t base = (int64_t)blockIdx.x * CH + threadIdx.x; base < g.nvec; base += (int64_t)gridDim.x * CH) {
    Raw rd[U], ry[U], ra[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t k = base + u * 256;
      if (k < g.nvec) {
        rd[u] = Vec<T>::load_raw(da + k * V);
        ry[u] = Vec<T>::load_raw(y + k * V);
        if (HAS_ACT && a) ra[u] = Vec<T>::load_raw(a + k * V);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t k = base + u * 256;
      if (k >= g.nvec) break;
      if (!hoist) coeffs((int)((k * V) % C));
      float d[V], yv[V], av[V];
      Vec<T>::unpack(rd[u], d);
      Vec<T>::unpack(ry[u], yv);
      if (HAS_ACT && a) Vec<T>::unpack(ra[u], av);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float dz = d[j];
        if (HAS_ACT) dz *= act_grad(fmaf(yv[j], sc[j], sh[j]), a ? av[j] : 0.f, g.act, g.slope);
        d[j] = fmaf(k1[j], dz, fmaf(pp[j], yv[j], qq[j]));
      }
      Vec<T>::store(dy + k * V, d);
    }
  }
}

struct ReduceDenseArgs {
  const void *y, *da, *a;
  int64_t P;
  int C;
  const float *scale, *shift, *mean, *invstd;
  int act; float slope;
  double* sums;
  void* dz_out;        // MODE 1 only: when non-NULL, dz is also stored here (may alias da) and the sums are those of the stored dz
};

// thread = (vector lane vl over C/V, pixel row pr); block-level smem reduction, one double atomic per channel per block
template <typename T, int MODE>
__global__ void __launch_bounds__(256) channel_reduce_dense_kernel(ReduceDenseArgs g) {
  constexpr int V = Vec<T>::N;
  __shared__ float red[2][256 * V / 8 + 1][8];          // [which][row-major scratch]; sized for V = 8 worst case below
  const int C = g.C, lanes = C / V, rows = 256 / lanes;
  const int vl = threadIdx.x % lanes, pr = threadIdx.x / lanes;
  const int c0 = vl * V;
  float s0[V], s1[V], sc[V], sh[V], mu[V], is[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    s0[j] = 0.f; s1[j] = 0.f;
    sc[j] = (MODE == 1 && g.scale) ? g.scale[c0 + j] : 1.f;
    sh[j] = (MODE == 1 && g.scale) ? g.shift[c0 + j] : 0.f;
    mu[j] = (MODE == 1 && g.mean) ? g.mean[c0 + j] : 0.f;
    is[j] = (MODE == 1 && g.invstd) ? g.invstd[c0 + j] : 1.f;
  }
  const T* y = reinterpret_cast<const T*>(g.y);
  const T* da = reinterpret_cast<const T*>(g.da);
  const T* a = reinterpret_cast<const T*>(g.a);
  if (pr < rows) {
    for (int64_t p = (int64_t)blockIdx.x * rows + pr; p < g.P; p += (int64_t)gridDim.x * rows) {
      float yv[V];
      Vec<T>::load(y + p * C + c0, yv);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < V; ++j) { s0[j] += yv[j]; s1[j] = fmaf(yv[j], yv[j], s1[j]); }
      } else {
        float d[V], av[V];
        Vec<T>::load(da + p * C + c0, d);
        if (a) Vec<T>::load(a + p * C + c0, av);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float z = fmaf(yv[j], sc[j], sh[j]);
          d[j] *= act_grad(z, a ? av[j] : 0.f, g.act, g.slope);
        }
        if (g.dz_out) {
          T* o = reinterpret_cast<T*>(g.dz_out) + p * C + c0;
          Vec<T>::store(o, d);
          Vec<T>::load(o, d);          // the value as stored (bf16 rounding)
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float dz = d[j];
          s0[j] += dz;
          s1[j] = fmaf(dz, (yv[j] - mu[j]) * is[j], s1[j]);
        }
      }
    }
  }
  // reduce over pixel rows: first fold rows with shuffles where a warp holds several rows of the same lane set
  float* r0 = &red[0][0][0];
  float* r1 = &red[1][0][0];
  // scratch layout [row][C]; rows*C = 256*V floats at most
  if (pr < rows) {
#pragma unroll
    for (int j = 0; j < V; ++j) { r0[pr * C + c0 + j] = s0[j]; r1[pr * C + c0 + j] = s1[j]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    const int which = c >= C, cc = which ? c - C : c;
    const float* r = which ? r1 : r0;
    float t = 0.f;
    for (int rr = 0; rr < rows; ++rr) t += r[rr * C + cc];
    atomicAdd(g.sums + c, (double)t);
  }
}

template <typename T>
static bool reduce_dense_ok(const b200gan_view* y) {
  constexpr int V = sizeof(T) == 2 ? 8 : 4;
  const int C = y->c;
  return dense_nhwc(y) && C % V == 0 && C / V <= 256 && 256 % (C / V) == 0;
}

// ------------------------------------------------------------------------------------------------
// per-channel reductions.  Thread layout: cl = tid % lanes_c is the channel lane (consecutive threads
// read consecutive channels of one pixel: coalesced for NHWC), pr = tid / lanes_c the pixel row of
// the block; a thread owns channels cl + j*lanes_c, j < 4 (so C <= 4*256).
// MODE 0: sums (y, y^2).  MODE 1: sums (dz, dz*xhat) for the BatchNorm backward.
// ------------------------------------------------------------------------------------------------
struct ReduceArgs {
  View y, da, a;
  const float *scale, *shift, *mean, *invstd;
  int act; float slope;
  double* sums;
  int lanes_c, rows;
};

template <typename T, int MODE>
__global__ void __launch_bounds__(256) channel_reduce_kernel(ReduceArgs g) {
  __shared__ float red[256];
  const int cl = threadIdx.x % g.lanes_c, pr = threadIdx.x / g.lanes_c;
  const int C = g.y.c;
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
  float sc[4], sh[4], mu[4], is[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cl + j * g.lanes_c;
    const bool ok = c < C;
    sc[j] = (MODE == 1 && g.scale && ok) ? g.scale[c] : 1.f;
    sh[j] = (MODE == 1 && g.scale && ok) ? g.shift[c] : 0.f;
    mu[j] = (MODE == 1 && g.mean && ok) ? g.mean[c] : 0.f;
    is[j] = (MODE == 1 && g.invstd && ok) ? g.invstd[c] : 1.f;
  }
  const int64_t P = (int64_t)g.y.n * g.y.h * g.y.w;
  if (pr < g.rows) {
    for (int64_t p = (int64_t)blockIdx.x * g.rows + pr; p < P; p += (int64_t)gridDim.x * g.rows) {
      const T* yb = reinterpret_cast<const T*>(g.y.ptr) + pix_offset(g.y, p);
      const T* db = MODE == 1 ? reinterpret_cast<const T*>(g.da.ptr) + pix_offset(g.da, p) : nullptr;
      const T* ab = (MODE == 1 && g.a.ptr) ? reinterpret_cast<const T*>(g.a.ptr) + pix_offset(g.a, p) : nullptr;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = cl + j * g.lanes_c;
        if (c < C) {
          const float y = ld_as_float(yb + (int64_t)c * g.y.sc);
          if (MODE == 0) {
            s0[j] += y;
            s1[j] = fmaf(y, y, s1[j]);
          } else {
            const float d = ld_as_float(db + (int64_t)c * g.da.sc);
            const float a = ab ? ld_as_float(ab + (int64_t)c * g.a.sc) : 0.f;
            const float z = fmaf(y, sc[j], sh[j]);
            const float dz = d * act_grad(z, a, g.act, g.slope);
            s0[j] += dz;
            s1[j] = fmaf(dz, (y - mu[j]) * is[j], s1[j]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cl + j * g.lanes_c;
    if (j * g.lanes_c >= C) break;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      __syncthreads();
      red[threadIdx.x] = which ? s1[j] : s0[j];
      __syncthreads();
      if (pr == 0 && c < C) {
        float t = 0.f;
        for (int r = 0; r < g.rows; ++r) t += red[r * g.lanes_c + cl];
        atomicAdd(g.sums + which * C + c, (double)t);
      }
    }
  }
}

static void reduce_geometry(int C, int* lanes_c, int* rows) {
  *lanes_c = C < 256 ? C : 256;
  *rows = 256 / *lanes_c;
}

int ew_bn_stats(const b200gan_view* y, double* sums, cudaStream_t st) {
  B200_CHECK_ARG(y->c <= 1024, "bn_stats: at most 1024 channels (got %d)", y->c);
  ReduceArgs g{};
  g.y = to_view(y); g.sums = sums;
  reduce_geometry(y->c, &g.lanes_c, &g.rows);
  B200_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * y->c, st));
  const int64_t P = (int64_t)y->n * y->h * y->w;
  if ((y->dtype == B200GAN_F32 && reduce_dense_ok<float>(y)) || (y->dtype == B200GAN_BF16 && reduce_dense_ok<__nv_bfloat16>(y))) {
    ReduceDenseArgs d{};
    d.y = y->ptr; d.P = P; d.C = y->c; d.sums = sums;
    const int V = y->dtype == B200GAN_F32 ? 4 : 8;
    const int rows = 256 / (y->c / V);
    int64_t nb = (P + (int64_t)rows * 32 - 1) / ((int64_t)rows * 32);
    if (nb > 8 * kNumSMs) nb = 8 * kNumSMs;
    if (nb < 1) nb = 1;
    if (y->dtype == B200GAN_F32) channel_reduce_dense_kernel<float, 0><<<(unsigned)nb, 256, 0, st>>>(d);
    else channel_reduce_dense_kernel<__nv_bfloat16, 0><<<(unsigned)nb, 256, 0, st>>>(d);
    B200_LAUNCH_CHECK("bn_stats(dense)");
    return 0;
  }
  int64_t blocks = (P + (int64_t)g.rows * 64 - 1) / ((int64_t)g.rows * 64);   // ~64 pixels per thread
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  if (blocks < 1) blocks = 1;
  if (y->dtype == B200GAN_F32) channel_reduce_kernel<float, 0><<<(unsigned)blocks, 256, 0, st>>>(g);
  else channel_reduce_kernel<__nv_bfloat16, 0><<<(unsigned)blocks, 256, 0, st>>>(g);
  B200_LAUNCH_CHECK("bn_stats");
  return 0;
}

// dz_out (optional): dense fast path only -- dz is stored there (may alias da); returns 2 instead of 0 when it was NOT written
// (generic strided path), so that the caller runs the in-place pass.
int ew_bn_bwd_reduce(const b200gan_view* da, const b200gan_view* y, const b200gan_view* a, const float* scale,
                     const float* shift, const float* mean, const float* invstd, int act, float slope, double* sums,
                     const b200gan_view* dz_out, cudaStream_t st) {
  B200_CHECK_ARG(y->c <= 1024, "bn_act_bwd_reduce: at most 1024 channels (got %d)", y->c);
  B200_CHECK_ARG(da->dtype == y->dtype && (!a || a->dtype == y->dtype), "bn_act_bwd_reduce: mixed dtypes");
  ReduceArgs g{};
  g.y = to_view(y); g.da = to_view(da);
  if (a) g.a = to_view(a); else g.a.ptr = nullptr;
  g.scale = scale; g.shift = shift; g.mean = mean; g.invstd = invstd; g.act = act; g.slope = slope; g.sums = sums;
  reduce_geometry(y->c, &g.lanes_c, &g.rows);
  B200_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * y->c, st));
  const int64_t P = (int64_t)y->n * y->h * y->w;
  void* dzp = (dz_out && dense_nhwc(dz_out) && dz_out->dtype == y->dtype) ? dz_out->ptr : nullptr;
  if (dense_nhwc(da) && (!a || dense_nhwc(a)) && (!dz_out || dzp) &&
      ((y->dtype == B200GAN_F32 && reduce_dense_ok<float>(y)) || (y->dtype == B200GAN_BF16 && reduce_dense_ok<__nv_bfloat16>(y)))) {
    ReduceDenseArgs d{};
    d.y = y->ptr; d.da = da->ptr; d.a = a ? a->ptr : nullptr; d.P = P; d.C = y->c; d.sums = sums;
    d.scale = scale; d.shift = shift; d.mean = mean; d.invstd = invstd; d.act = act; d.slope = slope;
    d.dz_out = dzp;
    const int V = y->dtype == B200GAN_F32 ? 4 : 8;
    const int rows = 256 / (y->c / V);
    int64_t nb = (P + (int64_t)rows * 32 - 1) / ((int64_t)rows * 32);
    if (nb > 8 * kNumSMs) nb = 8 * kNumSMs;
    if (nb < 1) nb = 1;
    if (y->dtype == B200GAN_F32) channel_reduce_dense_kernel<float, 1><<<(unsigned)nb, 256, 0, st>>>(d);
    else channel_reduce_dense_kernel<__nv_bfloat16, 1><<<(unsigned)nb, 256, 0, st>>>(d);
    B200_LAUNCH_CHECK("bn_act_bwd_reduce(dense)");
    return 0;
  }
  int64_t blocks = (P + (int64_t)g.rows * 64 - 1) / ((int64_t)g.rows * 64);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  if (blocks < 1) blocks = 1;
  if (y->dtype == B200GAN_F32) channel_reduce_kernel<float, 1><<<(unsigned)blocks, 256, 0, st>>>(g);
  else channel_reduce_kernel<__nv_bfloat16, 1><<<(unsigned)blocks, 256, 0, st>>>(g);
  B200_LAUNCH_CHECK("bn_act_bwd_reduce");
  return dz_out ? 2 : 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const double* sums, int C, double count, const float* gamma, const float* beta,
                                   float* rmean, float* rvar, int64_t* nbt, float momentum, float eps, float* scale,
                                   float* shift, float* smean, float* sinvstd) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double mean = sums[c] / count;
    double var = sums[C + c] / count - mean * mean;     // biased variance
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
    smean[c] = (float)mean;
    sinvstd[c] = invstd;
    if (rmean) {
      const double unbiased = count > 1.0 ? var * (count / (count - 1.0)) : var;
      rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
      rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unbiased;
    }
  }
  if (threadIdx.x == 0 && nbt) *nbt += 1;
}

int ew_bn_finalize(double* sums, int C, int64_t count, const float* gamma, const float* beta, float* rmean, float* rvar,
                   int64_t* nbt, float momentum, float eps, float* scale, float* shift, float* smean, float* sinvstd,
                   cudaStream_t st) {
  bn_finalize_kernel<<<1, 256, 0, st>>>(sums, C, (double)count, gamma, beta, rmean, rvar, nbt, momentum, eps, scale, shift,
                                        smean, sinvstd);
  B200_LAUNCH_CHECK("bn_finalize");
  return 0;
}

__global__ void bn_eval_coeffs_kernel(int C, const float* gamma, const float* beta, const float* rmean, const float* rvar,
                                      float eps, float* scale, float* shift) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float sc = gamma[c] / sqrtf(rvar[c] + eps);
    scale[c] = sc;
    shift[c] = beta[c] - rmean[c] * sc;
  }
}

int ew_bn_eval_coeffs(int C, const float* gamma, const float* beta, const float* rmean, const float* rvar, float eps,
                      float* scale, float* shift, cudaStream_t st) {
  bn_eval_coeffs_kernel<<<1, 256, 0, st>>>(C, gamma, beta, rmean, rvar, eps, scale, shift);
  B200_LAUNCH_CHECK("bn_eval_coeffs");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// bn_finalize + bn_act_fwd in ONE launch (training forward, local statistics): every CTA derives the layer's scale / shift from the
// fp64 sums with the arithmetic of bn_finalize_kernel (bit-identical coefficients) into shared memory, the first CTA also writes the
// saved coefficients and the running statistics; then the normalise + activation pass of bn_act_fwd_dense_kernel.  Saves one
// single-CTA launch (5 us + a dependent-launch gap) per BatchNorm layer and forward pass: 17 per DCGAN iteration.
// ------------------------------------------------------------------------------------------------
struct FinalizeArgs {
  const double* sums; int C; double count;
  const float *gamma, *beta;
  float *rmean, *rvar; int64_t* nbt; float momentum, eps;
  float *scale, *shift, *smean, *sinvstd;
};
constexpr int kFinMaxC = 1024;

template <typename T, int U = 4>
__global__ void __launch_bounds__(256) bn_finalize_act_fwd_dense_kernel(const T* __restrict__ y, T* __restrict__ a, int64_t nvec, FinalizeArgs f,
                                                                        int act, float slope) {
  __shared__ float s_sc[kFinMaxC], s_sh[kFinMaxC];
  const int C = f.C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double mean = f.sums[c] / f.count;
    double var = f.sums[C + c] / f.count - mean * mean;     // biased variance
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)f.eps));
    const float sc = f.gamma[c] * invstd;
    const float sh = f.beta[c] - (float)mean * sc;
    s_sc[c] = sc; s_sh[c] = sh;
    if (blockIdx.x == 0) {
      f.scale[c] = sc; f.shift[c] = sh; f.smean[c] = (float)mean; f.sinvstd[c] = invstd;
      if (f.rmean) {
        const double unbiased = f.count > 1.0 ? var * (f.count / (f.count - 1.0)) : var;
        f.rmean[c] = (1.f - f.momentum) * f.rmean[c] + f.momentum * (float)mean;
        f.rvar[c] = (1.f - f.momentum) * f.rvar[c] + f.momentum * (float)unbiased;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && f.nbt) *f.nbt += 1;
  __syncthreads();
  constexpr int V = Vec<T>::N;
  const bool hoist = (256 * V) % C == 0;
  float sc[V], sh[V];
  int c0 = (int)(((int64_t)threadIdx.x * V) % C);
#pragma unroll
  for (int j = 0; j < V; ++j) { sc[j] = s_sc[c0 + j]; sh[j] = s_sh[c0 + j]; }
  using Raw = typename Vec<T>::Raw;
  constexpr int64_t CH = 256 * U;
  for (int64_t base = (int64_t)blockIdx.x * CH + threadIdx.x; base < nvec; base += (int64_t)gridDim.x * CH) {
    Raw r[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (base + u * 256 < nvec) r[u] = Vec<T>::load_raw(y + (base + u * 256) * V);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t k = base + u * 256;
      if (k >= nvec) break;
      if (!hoist) {
        c0 = (int)((k * V) % C);
#pragma unroll
        for (int j = 0; j < V; ++j) { sc[j] = s_sc[c0 + j]; sh[j] = s_sh[c0 + j]; }
      }
      float v[V];
      Vec<T>::unpack(r[u], v);
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = act_fwd(fmaf(v[j], sc[j], sh[j]), act, slope);
      Vec<T>::store(a + k * V, v);
    }
  }
}

// returns 1 when the tensors do not qualify for the fused kernel (the caller then runs the two passes)
int ew_bn_finalize_act_fwd(double* sums, int C, int64_t count, const float* gamma, const float* beta, float* rmean, float* rvar, int64_t* nbt,
                           float momentum, float eps, float* scale, float* shift, float* smean, float* sinvstd, const b200gan_view* y, int act,
                           float slope, const b200gan_view* a, cudaStream_t st) {
  const int V = y->dtype == B200GAN_F32 ? 4 : 8;
  if (C > kFinMaxC || y->c != C || y->dtype != a->dtype || !dense_nhwc(y) || !dense_nhwc(a) || y->c % V != 0) return 1;
  if (y->n != a->n || y->h != a->h || y->w != a->w || y->c != a->c) return 1;
  const int64_t nvec = (int64_t)y->n * y->h * y->w * y->c / V;
  int64_t nbk = (nvec + 256 * 4 - 1) / (256 * 4);
  if (nbk > 4 * kNumSMs) nbk = 4 * kNumSMs;
  if (nbk < 1) nbk = 1;
  FinalizeArgs f{sums, C, (double)count, gamma, beta, rmean, rvar, nbt, momentum, eps, scale, shift, smean, sinvstd};
  if (y->dtype == B200GAN_F32)
    bn_finalize_act_fwd_dense_kernel<float><<<(unsigned)nbk, 256, 0, st>>>((const float*)y->ptr, (float*)a->ptr, nvec, f, act, slope);
  else
    bn_finalize_act_fwd_dense_kernel<__nv_bfloat16><<<(unsigned)nbk, 256, 0, st>>>((const __nv_bfloat16*)y->ptr, (__nv_bfloat16*)a->ptr, nvec, f, act, slope);
  B200_LAUNCH_CHECK("bn_finalize_act_fwd");
  return 0;
}

// ------------------------------------------------------------------------------------------------
template <typename T, typename TD>
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(View y, View a, const float* scale, const float* shift, int act,
                                                         float slope) {
  const int64_t total = (int64_t)y.n * y.h * y.w * y.c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % y.c);
    const int64_t pix = i / y.c;
    float z = ld_as_float(reinterpret_cast<const T*>(y.ptr) + pix_offset(y, pix) + (int64_t)c * y.sc);
    if (scale) z = fmaf(z, scale[c], shift[c]);
    st_from_float(reinterpret_cast<TD*>(a.ptr) + pix_offset(a, pix) + (int64_t)c * a.sc, act_fwd(z, act, slope));
  }
}

static unsigned ew_blocks(int64_t total) {
  int64_t b = (total + 256 * 4 - 1) / (256 * 4);
  if (b > 16 * kNumSMs) b = 16 * kNumSMs;
  if (b < 1) b = 1;
  return (unsigned)b;
}

int ew_bn_act_fwd(const b200gan_view* y, const float* scale, const float* shift, int act, float slope, const b200gan_view* a,
                  cudaStream_t st) {
  B200_CHECK_ARG(y->n == a->n && y->h == a->h && y->w == a->w && y->c == a->c, "bn_act_fwd: extent mismatch");
  const int64_t total = (int64_t)y->n * y->h * y->w * y->c;
  {
    const int V = y->dtype == B200GAN_F32 ? 4 : 8;
    // without BatchNorm nothing is per-channel: any dense tensor whose size is a multiple of the vector width qualifies
    const bool chan_free = scale == nullptr;
    if (y->dtype == a->dtype && dense_nhwc(y) && dense_nhwc(a) && (y->c % V == 0 || (chan_free && total % V == 0))) {
      const int64_t nvec = total / V;
      const int Cv = (y->c % V == 0) ? y->c : V;
      int64_t nbk = (nvec + 256 * 4 - 1) / (256 * 4);
      if (nbk > 4 * kNumSMs) nbk = 4 * kNumSMs;       // one resident wave (64 registers: 4 CTAs per SM), grid-stride inside
      if (nbk < 1) nbk = 1;
      if (y->dtype == B200GAN_F32)
        bn_act_fwd_dense_kernel<float><<<(unsigned)nbk, 256, 0, st>>>((const float*)y->ptr, (float*)a->ptr, nvec, Cv, scale, shift, act, slope);
      else
        bn_act_fwd_dense_kernel<__nv_bfloat16><<<(unsigned)nbk, 256, 0, st>>>((const __nv_bfloat16*)y->ptr, (__nv_bfloat16*)a->ptr, nvec, Cv, scale, shift, act, slope);
      B200_LAUNCH_CHECK("bn_act_fwd(dense)");
      return 0;
    }
  }
  const unsigned nb = ew_blocks(total);
  const View vy = to_view(y), va = to_view(a);
  if (y->dtype == B200GAN_F32 && a->dtype == B200GAN_F32) bn_act_fwd_kernel<float, float><<<nb, 256, 0, st>>>(vy, va, scale, shift, act, slope);
  else if (y->dtype == B200GAN_F32) bn_act_fwd_kernel<float, __nv_bfloat16><<<nb, 256, 0, st>>>(vy, va, scale, shift, act, slope);
  else if (a->dtype == B200GAN_F32) bn_act_fwd_kernel<__nv_bfloat16, float><<<nb, 256, 0, st>>>(vy, va, scale, shift, act, slope);
  else bn_act_fwd_kernel<__nv_bfloat16, __nv_bfloat16><<<nb, 256, 0, st>>>(vy, va, scale, shift, act, slope);
  B200_LAUNCH_CHECK("bn_act_fwd");
  return 0;
}

struct BwdApplyArgs {
  View da, y, a, dy;
  const float *scale, *shift, *mean, *invstd, *gamma;
  const double* sums;
  double count;
  int act; float slope;
  float *dgamma, *dbeta;
};

template <typename T, typename TD>
__global__ void __launch_bounds__(256) bn_act_bwd_apply_kernel(BwdApplyArgs g) {
  const int C = g.y.c;
  if (g.scale && blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (g.dbeta) g.dbeta[c] += (float)g.sums[c];
      if (g.dgamma) g.dgamma[c] += (float)g.sums[C + c];
    }
  }
  const int64_t total = (int64_t)g.y.n * g.y.h * g.y.w * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t pix = i / C;
    const float y = ld_as_float(reinterpret_cast<const T*>(g.y.ptr) + pix_offset(g.y, pix) + (int64_t)c * g.y.sc);
    const float d = ld_as_float(reinterpret_cast<const T*>(g.da.ptr) + pix_offset(g.da, pix) + (int64_t)c * g.da.sc);
    const float a = g.a.ptr ? ld_as_float(reinterpret_cast<const T*>(g.a.ptr) + pix_offset(g.a, pix) + (int64_t)c * g.a.sc) : 0.f;
    float r;
    if (g.scale) {
      const float z = fmaf(y, g.scale[c], g.shift[c]);
      const float dz = d * act_grad(z, a, g.act, g.slope);
      const float xhat = (y - g.mean[c]) * g.invstd[c];
      const float m1 = (float)(g.sums[c] / g.count), m2 = (float)(g.sums[C + c] / g.count);
      r = g.gamma[c] * g.invstd[c] * (dz - m1 - xhat * m2);
    } else {
      r = d * act_grad(y, a, g.act, g.slope);
    }
    st_from_float(reinterpret_cast<TD*>(g.dy.ptr) + pix_offset(g.dy, pix) + (int64_t)c * g.dy.sc, r);
  }
}

int ew_bn_act_bwd_apply(const b200gan_view* da, const b200gan_view* y, const b200gan_view* a, const float* scale,
                        const float* shift, const float* mean, const float* invstd, const float* gamma, double* sums,
                        int64_t count, int act, float slope, const b200gan_view* dy, float* dgamma, float* dbeta,
                        cudaStream_t st) {
  B200_CHECK_ARG(da->dtype == y->dtype && (!a || a->dtype == y->dtype), "bn_act_bwd_apply: da, y and a must share a dtype");
  BwdApplyArgs g{};
  g.da = to_view(da); g.y = to_view(y); g.dy = to_view(dy);
  if (a) g.a = to_view(a); else g.a.ptr = nullptr;
  g.scale = scale; g.shift = shift; g.mean = mean; g.invstd = invstd; g.gamma = gamma; g.sums = sums; g.count = (double)count;
  g.act = act; g.slope = slope; g.dgamma = dgamma; g.dbeta = dbeta;
  const int64_t total = (int64_t)y->n * y->h * y->w * y->c;
  {
    const int V = y->dtype == B200GAN_F32 ? 4 : 8;
    const bool chan_free = scale == nullptr;
    if (y->dtype == dy->dtype && dense_nhwc(y) && dense_nhwc(da) && dense_nhwc(dy) && (!a || dense_nhwc(a)) &&
        (y->c % V == 0 || (chan_free && total % V == 0))) {
      BwdDenseArgs d{};
      d.da = da->ptr; d.y = y->ptr; d.a = a ? a->ptr : nullptr; d.dy = dy->ptr; d.nvec = total / V; d.C = (y->c % V == 0) ? y->c : V;
      d.scale = scale; d.shift = shift; d.mean = mean; d.invstd = invstd; d.gamma = gamma; d.sums = sums; d.count = (double)count;
      d.act = act; d.slope = slope; d.dgamma = dgamma; d.dbeta = dbeta;
      int64_t nbk = (d.nvec + 256 * 4 - 1) / (256 * 4);
      // one resident wave (3 CTAs per SM at 80 registers): every thread derives its channel coefficients (fp64 divisions of the
      // sums, five loads per channel) once; with 16 CTAs per SM queued that prologue was a third of the kernel's load traffic
      const int64_t cap = 3 * kNumSMs;
      if (nbk > cap) nbk = cap;
      if (nbk < 1) nbk = 1;
      const bool has_act = act != B200GAN_ACT_NONE;
      if (y->dtype == B200GAN_F32) {
        if (has_act) bn_act_bwd_apply_dense_kernel<float, true><<<(unsigned)nbk, 256, 0, st>>>(d);
        else bn_act_bwd_apply_dense_kernel<float, false><<<(unsigned)nbk, 256, 0, st>>>(d);
      } else {
        if (has_act) bn_act_bwd_apply_dense_kernel<__nv_bfloat16, true><<<(unsigned)nbk, 256, 0, st>>>(d);
        else bn_act_bwd_apply_dense_kernel<__nv_bfloat16, false><<<(unsigned)nbk, 256, 0, st>>>(d);
      }
      B200_LAUNCH_CHECK("bn_act_bwd_apply(dense)");
      return 0;
    }
  }
  const unsigned nb = ew_blocks(total);
  if (y->dtype == B200GAN_F32 && dy->dtype == B200GAN_F32) bn_act_bwd_apply_kernel<float, float><<<nb, 256, 0, st>>>(g);
  else if (y->dtype == B200GAN_F32) bn_act_bwd_apply_kernel<float, __nv_bfloat16><<<nb, 256, 0, st>>>(g);
  else if (dy->dtype == B200GAN_F32) bn_act_bwd_apply_kernel<__nv_bfloat16, float><<<nb, 256, 0, st>>>(g);
  else bn_act_bwd_apply_kernel<__nv_bfloat16, __nv_bfloat16><<<nb, 256, 0, st>>>(g);
  B200_LAUNCH_CHECK("bn_act_bwd_apply");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// d *= act'(scale[c]*y + shift[c]) in place (the unfused form of the `prev_*` input-gradient epilogue, b200gan_fuse)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) act_bwd_inplace_kernel(View d, View y, const float* scale, const float* shift, int act, float slope) {
  const int64_t total = (int64_t)y.n * y.h * y.w * y.c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % y.c);
    const int64_t pix = i / y.c;
    float z = ld_as_float(reinterpret_cast<const T*>(y.ptr) + pix_offset(y, pix) + (int64_t)c * y.sc);
    if (scale) z = fmaf(z, scale[c], shift[c]);
    T* dp = reinterpret_cast<T*>(d.ptr) + pix_offset(d, pix) + (int64_t)c * d.sc;
    st_from_float(dp, ld_as_float(dp) * act_grad(z, 0.f, act, slope));
  }
}

int ew_act_bwd_inplace(const b200gan_view* d, const b200gan_view* y, const float* scale, const float* shift, int act, float slope,
                       cudaStream_t st) {
  B200_CHECK_ARG(d->dtype == y->dtype && d->n == y->n && d->h == y->h && d->w == y->w && d->c == y->c, "act_bwd_inplace: d and y differ");
  const int64_t total = (int64_t)y->n * y->h * y->w * y->c;
  if (y->dtype == B200GAN_F32) act_bwd_inplace_kernel<float><<<ew_blocks(total), 256, 0, st>>>(to_view(d), to_view(y), scale, shift, act, slope);
  else act_bwd_inplace_kernel<__nv_bfloat16><<<ew_blocks(total), 256, 0, st>>>(to_view(d), to_view(y), scale, shift, act, slope);
  B200_LAUNCH_CHECK("act_bwd_inplace");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Sigmoid + BCE(mean) against a constant target, forward and backward, one CTA (B <= a few thousand).
// torch: loss_i = (t-1)*max(log1p(-p),-100) - t*max(log p,-100);  dL/dp = (p-t)/max((1-p)p,1e-12)/B
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bce_sigmoid_kernel(const float* logit, int B, float t, float gscale, float* prob,
                                                          float* out2, float* dlogit) {
  __shared__ float red[2][8];
  float ls = 0.f, ps = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float p = 1.f / (1.f + expf(-logit[i]));
    if (prob) prob[i] = p;
    const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(log1pf(-p), -100.f);
    ls += (t - 1.f) * l1p - t * lp;
    ps += p;
    if (dlogit) {
      const float pq = (1.f - p) * p;
      const float dp = (p - t) / fmaxf(pq, 1e-12f) / (float)B;
      dlogit[i] = gscale * dp * pq;
    }
  }
  ls = warp_sum(ls);
  ps = warp_sum(ps);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ls; red[1][threadIdx.x >> 5] = ps; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
    out2[0] = a / (float)B;
    out2[1] = b / (float)B;
  }
}

int ew_bce_sigmoid(const float* logit, int B, float target, float gscale, float* prob, float* out2, float* dlogit,
                   cudaStream_t st) {
  bce_sigmoid_kernel<<<1, 256, 0, st>>>(logit, B, target, gscale, prob, out2, dlogit);
  B200_LAUNCH_CHECK("bce_sigmoid");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// torch.optim.Adam, one flat arena.  exp_avg.lerp_(g, 1-b1); exp_avg_sq = b2*v + (1-b2) g^2;
// denom = sqrt(v)/sqrt(1-b2^t) + eps; p -= lr/(1-b1^t) * m/denom
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float one_m_b1, float b2, float one_m_b2,
                                         float step_size, float bc2_sqrt, float eps) {
  m = m + (g - m) * one_m_b1;
  v = v * b2 + one_m_b2 * g * g;
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float one_m_b1, float b2, float one_m_b2,
                                                   float step_size, float bc2_sqrt, float eps, float gscale, int vec_ok,
                                                   const int64_t* __restrict__ step_dev, double lr, double b1, double b2d) {
  if (step_dev) {
    // step count held on the device (CUDA-graph replay: nothing about the step may be baked into the launch parameters);
    // same double-precision bias corrections as the host path below
    __shared__ float sh[2];
    if (threadIdx.x == 0) {
      const double t = (double)*step_dev;
      sh[0] = (float)(lr / (1.0 - pow(b1, t)));
      sh[1] = (float)sqrt(1.0 - pow(b2d, t));
    }
    __syncthreads();
    step_size = sh[0];
    bc2_sqrt = sh[1];
  }
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = vec_ok ? n / 4 : 0;
  for (int64_t i = tid; i < n4; i += nth) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    adam_one(pp.x, gg.x * gscale, mm.x, vv.x, one_m_b1, b2, one_m_b2, step_size, bc2_sqrt, eps);
    adam_one(pp.y, gg.y * gscale, mm.y, vv.y, one_m_b1, b2, one_m_b2, step_size, bc2_sqrt, eps);
    adam_one(pp.z, gg.z * gscale, mm.z, vv.z, one_m_b1, b2, one_m_b2, step_size, bc2_sqrt, eps);
    adam_one(pp.w, gg.w * gscale, mm.w, vv.w, one_m_b1, b2, one_m_b2, step_size, bc2_sqrt, eps);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (int64_t i = n4 * 4 + tid; i < n; i += nth) {
    float pp = p[i], mm = m[i], vv = v[i];
    adam_one(pp, g[i] * gscale, mm, vv, one_m_b1, b2, one_m_b2, step_size, bc2_sqrt, eps);
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

int ew_adam(float* p, const float* g, float* m, float* v, int64_t n, double lr, double b1, double b2, double eps, int step,
            const int64_t* step_dev, float gscale, cudaStream_t st) {
  const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
  const float step_size = (float)(lr / bc1), bc2_sqrt = (float)sqrt(bc2);
  const int vec_ok = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0) ? 1 : 0;
  adam_kernel<<<ew_blocks(n / 4 + 1), 256, 0, st>>>(p, g, m, v, n, (float)(1.0 - b1), (float)b2, (float)(1.0 - b2), step_size, bc2_sqrt,
                                                    (float)eps, gscale, vec_ok, step_dev, lr, b1, b2);
  B200_LAUNCH_CHECK("adam");
  return 0;
}

// ------------------------------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) copy_view_kernel(View s, View d) {
  const int64_t total = (int64_t)s.n * s.h * s.w * s.c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % s.c);
    const int64_t pix = i / s.c;
    const float v = ld_as_float(reinterpret_cast<const TS*>(s.ptr) + pix_offset(s, pix) + (int64_t)c * s.sc);
    st_from_float(reinterpret_cast<TD*>(d.ptr) + pix_offset(d, pix) + (int64_t)c * d.sc, v);
  }
}

int ew_copy_view(const b200gan_view* s, const b200gan_view* d, cudaStream_t st) {
  B200_CHECK_ARG(s->n == d->n && s->h == d->h && s->w == d->w && s->c == d->c, "copy_view: extent mismatch");
  const int64_t total = (int64_t)s->n * s->h * s->w * s->c;
  const unsigned b = ew_blocks(total);
  if (s->dtype == B200GAN_F32 && d->dtype == B200GAN_F32) copy_view_kernel<float, float><<<b, 256, 0, st>>>(to_view(s), to_view(d));
  else if (s->dtype == B200GAN_F32) copy_view_kernel<float, __nv_bfloat16><<<b, 256, 0, st>>>(to_view(s), to_view(d));
  else if (d->dtype == B200GAN_F32) copy_view_kernel<__nv_bfloat16, float><<<b, 256, 0, st>>>(to_view(s), to_view(d));
  else copy_view_kernel<__nv_bfloat16, __nv_bfloat16><<<b, 256, 0, st>>>(to_view(s), to_view(d));
  B200_LAUNCH_CHECK("copy_view");
  return 0;
}

__global__ void fill_kernel(float* p, int64_t n, float v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

int ew_fill(float* p, int64_t n, float v, cudaStream_t st) {
  if (n <= 0) return 0;
  fill_kernel<<<ew_blocks(n), 256, 0, st>>>(p, n, v);
  B200_LAUNCH_CHECK("fill");
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Input pipeline (reference: src/data_loader.py:17-23 'train' transform after Resize, consumed at src/train_gan.py:123):
// out[b] = Normalize(ToTensor(hflip?(cache[index[b]]))) gathered from a device-resident uint8 cache (N, C, H, W).
// One thread = four consecutive output pixels of one (b, c, h) row: a 4-byte read (mirrored rows read the mirrored group),
// (v / 255 - mean) / std in fp32 with true divisions, exactly torchvision's ToTensor + Normalize.
// ---------------------------------------------------------------------------------------------------
struct AugCoef { float mean[4], stdv[4]; };

template <typename TO>
__global__ void __launch_bounds__(256) gather_augment_kernel(const uint8_t* __restrict__ cache, const int64_t* __restrict__ index,
                                                             const uint8_t* __restrict__ flip, AugCoef cf, View out, int64_t num_images) {
  const int W4 = (out.w + 3) >> 2;
  const int64_t total = (int64_t)out.n * out.c * out.h * W4;
  const bool vec_in = (out.w & 3) == 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w4 = (int)(i % W4);
    int64_t r = i / W4;
    const int h = (int)(r % out.h); r /= out.h;
    const int c = (int)(r % out.c);
    const int b = (int)(r / out.c);
    int64_t src = index[b];
    src = src < 0 ? 0 : (src >= num_images ? num_images - 1 : src);
    const bool mirror = flip && flip[b] != 0;
    const uint8_t* row = cache + ((src * out.c + c) * out.h + h) * (int64_t)out.w;
    const int w0 = 4 * w4;
    uint8_t px[4];
    if (vec_in) {
      const uchar4 v = *reinterpret_cast<const uchar4*>(row + (mirror ? out.w - 4 - w0 : w0));
      if (mirror) { px[0] = v.w; px[1] = v.z; px[2] = v.y; px[3] = v.x; }
      else { px[0] = v.x; px[1] = v.y; px[2] = v.z; px[3] = v.w; }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) px[e] = w0 + e < out.w ? row[mirror ? out.w - 1 - (w0 + e) : w0 + e] : 0;
    }
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = __fdiv_rn(__fdiv_rn((float)px[e], 255.f) - cf.mean[c], cf.stdv[c]);
    TO* dst = reinterpret_cast<TO*>(out.ptr) + (int64_t)b * out.sn + (int64_t)h * out.sh + (int64_t)w0 * out.sw + (int64_t)c * out.sc;
    if (sizeof(TO) == 4 && out.sw == 1 && vec_in && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (w0 + e < out.w) st_from_float(dst + (int64_t)e * out.sw, o[e]);
    }
  }
}

int ew_gather_augment(const uint8_t* cache, int64_t num_images, const int64_t* index, const uint8_t* flip, const float* mean, const float* stdv,
                      const b200gan_view* out, cudaStream_t st) {
  B200_CHECK_ARG(out->c >= 1 && out->c <= 4, "gather_augment: 1..4 channels");
  AugCoef cf;
  for (int c = 0; c < 4; ++c) { cf.mean[c] = c < out->c && mean ? mean[c] : 0.f; cf.stdv[c] = c < out->c && stdv ? stdv[c] : 1.f; }
  const int64_t total = (int64_t)out->n * out->c * out->h * ((out->w + 3) / 4);
  if (total <= 0) return 0;
  if (out->dtype == B200GAN_F32) gather_augment_kernel<float><<<ew_blocks(total), 256, 0, st>>>(cache, index, flip, cf, to_view(out), num_images);
  else gather_augment_kernel<__nv_bfloat16><<<ew_blocks(total), 256, 0, st>>>(cache, index, flip, cf, to_view(out), num_images);
  B200_LAUNCH_CHECK("gather_augment");
  return 0;
}

}  // namespace b200gan
