// Kernels specific to the WGAN-GP critic update (reference src/wggan.py:72-89 `gradient_penalty`, src/train_wggan.py:70-85): the pieces
// of the gradient penalty's DOUBLE backward that are not convolutions -- the second-order terms of training-mode BatchNorm -- plus the
// per-sample norm / interpolation arithmetic of the penalty itself.  The convolutions of the double backward are ordinary
// b200gan_conv2d_{fprop,dgrad,wgrad} calls (the reverse of an input-gradient convolution is a forward convolution of the adjoint).
//
// Notation (per channel, n = N*H*W samples): xhat = (y - mean) invstd, P(v) = v - mean(v) - xhat mean(v xhat).  BatchNorm backward is
// dy = gamma invstd P(dz).  Given r = adjoint of dy (from the layer below in the reverse sweep):
//     adjoint of dz      = gamma invstd P(r)                       (times the activation derivative: adjoint of the activation gradient)
//     adjoint of gamma  += invstd sum r P(dz)
//     adjoint of y (inj) = -gamma invstd^2 [ xhat mean(r P(dz)) + mean(dz xhat) P(r) + mean(r xhat) P(dz) ]
// (derivation and numpy restatement: oracle/wgan_oracle.py, pinned against torch.autograd's double backward of the reference).
// Storage f32 or bf16 through strided views; all arithmetic fp32, channel sums in fp64.
#include <math.h>

#include "common.cuh"

namespace b200gan {

namespace {

__device__ __forceinline__ void st_rt(void* base, int dtype, int64_t off, float x) {
  if (dtype == B200GAN_F32) reinterpret_cast<float*>(base)[off] = x;
  else reinterpret_cast<__nv_bfloat16*>(base)[off] = __float2bfloat16_rn(x);
}

// Row-based indexing for the strided views: a CTA row (blockIdx.y) is one (n, h) line of W*C elements, a thread walks elements e = tid, tid + 256,
// ... of the line, so the per-element index arithmetic is one 32-bit division by C (no 64-bit div / mod chains).
struct RowBase { int64_t r, y, dz, u, inj; };
__device__ __forceinline__ int64_t row_off(const View& v, int n, int h) { return (int64_t)n * v.sn + (int64_t)h * v.sh; }

// sums[0..C) = sum r, [C..2C) = sum r xhat, [2C..3C) = sum r dz.  256 % C == 0 or C % 256 == 0 (every channel count of the critic): a thread
// only ever meets channels tid % C (+ 256, ...), so it accumulates in registers (up to four channel slots) and touches shared memory once.
__global__ void __launch_bounds__(256) bn_bwd_bwd_reduce_kernel(View r, View y, View dz, const float* __restrict__ mean, const float* __restrict__ invstd,
                                                               double* __restrict__ sums, int rows_per_cta) {
  extern __shared__ float acc[];                 // [3][C]
  const int C = r.c, WC = r.w * C;
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int nslots = C > 256 ? C / 256 : 1;
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  float mu[4], is[4];
  for (int k = 0; k < 4; ++k) { const int c = (threadIdx.x + 256 * k) % C; mu[k] = mean[c]; is[k] = invstd[c]; }
  const int row0 = blockIdx.x * rows_per_cta, nrows = r.n * r.h;
  for (int row = row0; row < row0 + rows_per_cta && row < nrows; ++row) {
    const int n = row / r.h, h = row - n * r.h;
    const int64_t br = row_off(r, n, h), by = row_off(y, n, h), bd = row_off(dz, n, h);
    int k = 0;
    for (int e = threadIdx.x; e < WC; e += 256) {
      const int w = e / C, c = e - w * C;
      const float rv = ld_rt(r.ptr, r.dtype, br + (int64_t)w * r.sw + (int64_t)c * r.sc);
      const float yv = ld_rt(y.ptr, y.dtype, by + (int64_t)w * y.sw + (int64_t)c * y.sc);
      const float dv = ld_rt(dz.ptr, dz.dtype, bd + (int64_t)w * dz.sw + (int64_t)c * dz.sc);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q == k) { s0[q] += rv; s1[q] = fmaf(rv, (yv - mu[q]) * is[q], s1[q]); s2[q] = fmaf(rv, dv, s2[q]); }
      if (++k == nslots) k = 0;
    }
  }
  for (int k = 0; k < nslots; ++k) {
    const int c = (threadIdx.x + 256 * k) % C;
    atomicAdd(&acc[c], s0[k]); atomicAdd(&acc[C + c], s1[k]); atomicAdd(&acc[2 * C + c], s2[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x)
    if (acc[i] != 0.f) atomicAdd(sums + i, (double)acc[i]);
}

// per channel: {mean r, mean r xhat, mean r P(dz), mean dz, mean dz xhat} -> u = gamma invstd P(r) act'(z), inj as above; block 0 adds the
// gamma adjoint.  `dzs` are the sums of the first backward ([0..C) = sum dz, [C..2C) = sum dz xhat: b200gan_fuse.prev_sums' contract).
__global__ void __launch_bounds__(256) bn_bwd_bwd_apply_kernel(View r, View y, View dz, const float* __restrict__ scale, const float* __restrict__ shift,
                                                              const float* __restrict__ mean, const float* __restrict__ invstd,
                                                              const float* __restrict__ gamma, const double* __restrict__ dzs,
                                                              const double* __restrict__ sums, double count, int act, float slope, View u, View inj,
                                                              float* __restrict__ dgamma, int rows_per_cta) {
  extern __shared__ float coef[];                // [10][C]: mr, mrx, mrp, m1, m2, gamma*invstd, invstd, mean, scale, shift
  const int C = r.c, WC = r.w * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double sr = sums[c], srx = sums[C + c], srd = sums[2 * C + c];
    const double m1 = dzs[c] / count, m2 = dzs[C + c] / count;
    const double srp = srd - m1 * sr - m2 * srx;             // sum r P(dz)
    coef[c] = (float)(sr / count); coef[C + c] = (float)(srx / count); coef[2 * C + c] = (float)(srp / count);
    coef[3 * C + c] = (float)m1; coef[4 * C + c] = (float)m2;
    coef[5 * C + c] = gamma[c] * invstd[c]; coef[6 * C + c] = invstd[c]; coef[7 * C + c] = mean[c];
    coef[8 * C + c] = scale[c]; coef[9 * C + c] = shift[c];
    if (blockIdx.x == 0 && dgamma) dgamma[c] += (float)((double)invstd[c] * srp);
  }
  __syncthreads();
  const int row0 = blockIdx.x * rows_per_cta, nrows = r.n * r.h;
  for (int row = row0; row < row0 + rows_per_cta && row < nrows; ++row) {
    const int n = row / r.h, h = row - n * r.h;
    const int64_t br = row_off(r, n, h), by = row_off(y, n, h), bd = row_off(dz, n, h), bu = row_off(u, n, h), bi = row_off(inj, n, h);
    for (int e = threadIdx.x; e < WC; e += 256) {
      const int w = e / C, c = e - w * C;
      const float rv = ld_rt(r.ptr, r.dtype, br + (int64_t)w * r.sw + (int64_t)c * r.sc);
      const float yv = ld_rt(y.ptr, y.dtype, by + (int64_t)w * y.sw + (int64_t)c * y.sc);
      const float dv = ld_rt(dz.ptr, dz.dtype, bd + (int64_t)w * dz.sw + (int64_t)c * dz.sc);
      const float is = coef[6 * C + c], gs = coef[5 * C + c];
      const float xh = (yv - coef[7 * C + c]) * is;
      const float pr = rv - coef[c] - xh * coef[C + c];
      const float pdz = dv - coef[3 * C + c] - xh * coef[4 * C + c];
      const float z = fmaf(yv, coef[8 * C + c], coef[9 * C + c]);
      const float d = act == B200GAN_ACT_RELU ? (z > 0.f ? 1.f : 0.f) : (act == B200GAN_ACT_LRELU ? (z > 0.f ? 1.f : slope) : 1.f);
      st_rt(u.ptr, u.dtype, bu + (int64_t)w * u.sw + (int64_t)c * u.sc, gs * pr * d);
      st_rt(inj.ptr, inj.dtype, bi + (int64_t)w * inj.sw + (int64_t)c * inj.sc, -(gs * is) * (xh * coef[2 * C + c] + coef[4 * C + c] * pr + coef[C + c] * pdz));
    }
  }
}

// out[n] += sum over (h,w,c) of x^2 (fp64; zeroed by the host wrapper): gridDim.y = samples
__global__ void __launch_bounds__(256) sample_sumsq_kernel(View x, double* __restrict__ out) {
  const int n = blockIdx.y;
  const int64_t per = (int64_t)x.h * x.w * x.c;
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % x.c);
    int64_t pix = i / x.c;
    const int w = (int)(pix % x.w);
    const int h = (int)(pix / x.w);
    const float v = ld_rt(x.ptr, x.dtype, (int64_t)n * x.sn + (int64_t)h * x.sh + (int64_t)w * x.sw + (int64_t)c * x.sc);
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += (double)ws[i];
    atomicAdd(out + n, t);
  }
}

// gp = lambda mean_n (||g_n|| - 1)^2 and coeff[n] = d gp / d g_n / g_n = lambda (2/N) (||g_n|| - 1) / ||g_n||   (wggan.py:87-88)
__global__ void gp_from_norms_kernel(const double* __restrict__ sumsq, int n, float lambda, float* __restrict__ gp, float* __restrict__ coeff) {
  __shared__ double part[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double nrm = sqrt(sumsq[i]);
    s += (nrm - 1.0) * (nrm - 1.0);
    coeff[i] = nrm > 0.0 ? (float)((double)lambda * 2.0 / n * (nrm - 1.0) / nrm) : 0.f;
  }
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) gp[0] = (float)((double)lambda * part[0] / n);
}

// out(n,.) = a[n] x(n,.) + b[n] y(n,.)   (y == nullptr: no second term).  a / b may be nullptr (= 1).  One (n, h) line per blockIdx.y step.
__global__ void __launch_bounds__(256) sample_axpby_kernel(View x, const float* __restrict__ a, View y, bool has_y, const float* __restrict__ b, View out,
                                                          int rows_per_cta) {
  const int C = x.c, WC = x.w * C, nrows = x.n * x.h;
  const int row0 = blockIdx.x * rows_per_cta;
  for (int row = row0; row < row0 + rows_per_cta && row < nrows; ++row) {
    const int n = row / x.h, h = row - n * x.h;
    const float an = a ? a[n] : 1.f, bn = b ? b[n] : 1.f;
    const int64_t bx = row_off(x, n, h), by = row_off(y, n, h), bo = row_off(out, n, h);
    for (int e = threadIdx.x; e < WC; e += 256) {
      const int w = e / C, c = e - w * C;
      float v = an * ld_rt(x.ptr, x.dtype, bx + (int64_t)w * x.sw + (int64_t)c * x.sc);
      if (has_y) v = fmaf(bn, ld_rt(y.ptr, y.dtype, by + (int64_t)w * y.sw + (int64_t)c * y.sc), v);
      st_rt(out.ptr, out.dtype, bo + (int64_t)w * out.sw + (int64_t)c * out.sc, v);
    }
  }
}

// out[0] = scale * sum x[0..n)
__global__ void mean_f32_kernel(const float* __restrict__ x, int64_t n, float scale, float* __restrict__ out) {
  __shared__ double part[256];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)x[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(part[0] * (double)scale);
}


// ---- dense bf16 fast paths: every operand a dense NHWC bf16 tensor of one shape, C a power of two in [8, 1024].  A thread owns ONE group of eight
//      consecutive channels for its whole life (the grid stride is a multiple of C / 8 vectors), so the per-channel coefficients / partial sums live in
//      registers and every access is 16 bytes.  Same arithmetic as the generic kernels above.
__device__ __forceinline__ void unpack8v(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) { const float2 t = __bfloat1622float2(h[k]); f[2 * k] = t.x; f[2 * k + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8v(const float* f) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
  return v;
}

__global__ void __launch_bounds__(256) bn_bwd_bwd_reduce_dense_kernel(const uint4* __restrict__ r, const uint4* __restrict__ y, const uint4* __restrict__ dz,
                                                                     const float* __restrict__ mean, const float* __restrict__ invstd, double* __restrict__ sums,
                                                                     int C, int64_t vecs) {
  extern __shared__ float acc[];                 // [3][C]
  for (int i = threadIdx.x; i < 3 * C; i += 256) acc[i] = 0.f;
  __syncthreads();
  const int64_t first = blockIdx.x * 256ll + threadIdx.x;
  const int c0 = (int)(first % (C / 8)) * 8;
  float mu[8], is[8], s0[8], s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { mu[k] = mean[c0 + k]; is[k] = invstd[c0 + k]; s0[k] = s1[k] = s2[k] = 0.f; }
  for (int64_t i = first; i < vecs; i += (int64_t)gridDim.x * 256) {
    float rv[8], yv[8], dv[8];
    unpack8v(r[i], rv); unpack8v(y[i], yv); unpack8v(dz[i], dv);
#pragma unroll
    for (int k = 0; k < 8; ++k) { s0[k] += rv[k]; s1[k] = fmaf(rv[k], (yv[k] - mu[k]) * is[k], s1[k]); s2[k] = fmaf(rv[k], dv[k], s2[k]); }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) { atomicAdd(&acc[c0 + k], s0[k]); atomicAdd(&acc[C + c0 + k], s1[k]); atomicAdd(&acc[2 * C + c0 + k], s2[k]); }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C; i += 256)
    if (acc[i] != 0.f) atomicAdd(sums + i, (double)acc[i]);
}

__global__ void __launch_bounds__(256) bn_bwd_bwd_apply_dense_kernel(const uint4* __restrict__ r, const uint4* __restrict__ y, const uint4* __restrict__ dz,
                                                                    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                                                                    const float* __restrict__ invstd, const float* __restrict__ gamma, const double* __restrict__ dzs,
                                                                    const double* __restrict__ sums, double count, int act, float slope, uint4* __restrict__ u,
                                                                    uint4* __restrict__ inj, float* __restrict__ dgamma, int C, int64_t vecs) {
  const int64_t first = blockIdx.x * 256ll + threadIdx.x;
  const int c0 = (int)(first % (C / 8)) * 8;
  float mr[8], mrx[8], mrp[8], m1[8], m2[8], gs[8], is[8], mu[8], sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = c0 + k;
    const double sr = sums[c], srx = sums[C + c], srd = sums[2 * C + c];
    const double a1 = dzs[c] / count, a2 = dzs[C + c] / count;
    const double srp = srd - a1 * sr - a2 * srx;             // sum r P(dz)
    mr[k] = (float)(sr / count); mrx[k] = (float)(srx / count); mrp[k] = (float)(srp / count);
    m1[k] = (float)a1; m2[k] = (float)a2;
    gs[k] = gamma[c] * invstd[c]; is[k] = invstd[c]; mu[k] = mean[c]; sc[k] = scale[c]; sh[k] = shift[c];
    if (first < C / 8 && dgamma) dgamma[c] += (float)((double)invstd[c] * srp);         // the first C/8 threads of the grid: one per channel group
  }
  for (int64_t i = first; i < vecs; i += (int64_t)gridDim.x * 256) {
    float rv[8], yv[8], dv[8], uo[8], io[8];
    unpack8v(r[i], rv); unpack8v(y[i], yv); unpack8v(dz[i], dv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (yv[k] - mu[k]) * is[k];
      const float pr = rv[k] - mr[k] - xh * mrx[k];
      const float pdz = dv[k] - m1[k] - xh * m2[k];
      const float z = fmaf(yv[k], sc[k], sh[k]);
      const float d = act == B200GAN_ACT_RELU ? (z > 0.f ? 1.f : 0.f) : (act == B200GAN_ACT_LRELU ? (z > 0.f ? 1.f : slope) : 1.f);
      uo[k] = gs[k] * pr * d;
      io[k] = -(gs[k] * is[k]) * (xh * mrp[k] + m2[k] * pr + mrx[k] * pdz);
    }
    u[i] = pack8v(uo);
    inj[i] = pack8v(io);
  }
}

// out = a[n] x + b[n] y, dense tensors of one dtype (T = float: 4 per access, bf16: 8), `per` accesses per sample
template <typename T>
__global__ void __launch_bounds__(256) sample_axpby_dense_kernel(const uint4* __restrict__ x, const float* __restrict__ a, const uint4* __restrict__ y,
                                                                const float* __restrict__ b, uint4* __restrict__ out, int64_t vecs, int64_t per) {
  constexpr int V = sizeof(T) == 4 ? 4 : 8;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < vecs; i += (int64_t)gridDim.x * 256) {
    const int n = (int)(i / per);
    const float an = a ? a[n] : 1.f, bn = b ? b[n] : 1.f;
    float xv[V], yv[V];
    const uint4 xr = x[i];
    if (sizeof(T) == 4) { const float* p = reinterpret_cast<const float*>(&xr); for (int k = 0; k < V; ++k) xv[k] = p[k]; }
    else unpack8v(xr, xv);
    if (y) {
      const uint4 yr = y[i];
      if (sizeof(T) == 4) { const float* p = reinterpret_cast<const float*>(&yr); for (int k = 0; k < V; ++k) yv[k] = p[k]; }
      else unpack8v(yr, yv);
#pragma unroll
      for (int k = 0; k < V; ++k) xv[k] = fmaf(bn, yv[k], an * xv[k]);
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) xv[k] *= an;
    }
    uint4 o;
    if (sizeof(T) == 4) { float* p = reinterpret_cast<float*>(&o); for (int k = 0; k < V; ++k) p[k] = xv[k]; }
    else o = pack8v(xv);
    out[i] = o;
  }
}

bool dense_view(const b200gan_view* v) {
  return v->sc == 1 && v->sw == v->c && v->sh == (int64_t)v->w * v->c && v->sn == (int64_t)v->h * v->w * v->c && (reinterpret_cast<uintptr_t>(v->ptr) & 15) == 0;
}

bool same_extent(const b200gan_view* a, const b200gan_view* b) { return a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c; }

}  // namespace

int gp_bn_bwd_bwd(const b200gan_view* r, const b200gan_view* y, const b200gan_view* dz, const float* scale, const float* shift, const float* mean,
                  const float* invstd, const float* gamma, const double* dz_sums, int64_t count, int act, float slope, const b200gan_view* u,
                  const b200gan_view* inj, float* dgamma, double* sums3, cudaStream_t st) {
  B200_CHECK_ARG(same_extent(r, y) && same_extent(r, dz) && same_extent(r, u) && same_extent(r, inj), "bn_bwd_bwd: views differ in extent");
  const int C = r->c;
  B200_CHECK_ARG(C <= 1024 && (256 % C == 0 || C % 256 == 0), "bn_bwd_bwd: channel count must divide 256 or be a multiple of it, at most 1024 (got %d)", C);
  B200_CUDA(cudaMemsetAsync(sums3, 0, sizeof(double) * 3 * C, st));
  const b200gan_view* all[5] = {r, y, dz, u, inj};
  bool fast = C >= 8 && (C & (C - 1)) == 0;
  for (const b200gan_view* v : all) fast = fast && v->dtype == B200GAN_BF16 && dense_view(v);
  if (fast) {
    const int64_t vecs = (int64_t)r->n * r->h * r->w * C / 8;
    int64_t blocks = (vecs + 255) / 256;
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;              // any grid keeps a thread on one channel group: 256 % (C / 8) == 0
    bn_bwd_bwd_reduce_dense_kernel<<<(unsigned)blocks, 256, 3 * C * sizeof(float), st>>>((const uint4*)r->ptr, (const uint4*)y->ptr, (const uint4*)dz->ptr, mean, invstd,
                                                                                        sums3, C, vecs);
    B200_LAUNCH_CHECK("bn_bwd_bwd_reduce_dense_kernel");
    bn_bwd_bwd_apply_dense_kernel<<<(unsigned)blocks, 256, 0, st>>>((const uint4*)r->ptr, (const uint4*)y->ptr, (const uint4*)dz->ptr, scale, shift, mean, invstd, gamma,
                                                                   dz_sums, sums3, (double)count, act, slope, (uint4*)u->ptr, (uint4*)inj->ptr, dgamma, C, vecs);
    B200_LAUNCH_CHECK("bn_bwd_bwd_apply_dense_kernel");
    return 0;
  }
  const int nrows = r->n * r->h;
  int rows_per_cta = (nrows + 8 * kNumSMs - 1) / (8 * kNumSMs);
  if (rows_per_cta < 1) rows_per_cta = 1;
  const unsigned blocks = (unsigned)((nrows + rows_per_cta - 1) / rows_per_cta);
  bn_bwd_bwd_reduce_kernel<<<blocks, 256, 3 * C * sizeof(float), st>>>(to_view(r), to_view(y), to_view(dz), mean, invstd, sums3, rows_per_cta);
  B200_LAUNCH_CHECK("bn_bwd_bwd_reduce_kernel");
  bn_bwd_bwd_apply_kernel<<<blocks, 256, 10 * C * sizeof(float), st>>>(to_view(r), to_view(y), to_view(dz), scale, shift, mean, invstd, gamma, dz_sums, sums3,
                                                                     (double)count, act, slope, to_view(u), to_view(inj), dgamma, rows_per_cta);
  B200_LAUNCH_CHECK("bn_bwd_bwd_apply_kernel");
  return 0;
}

int gp_sample_sumsq(const b200gan_view* x, double* out, cudaStream_t st) {
  B200_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * x->n, st));
  const int64_t per = (int64_t)x->h * x->w * x->c;
  int bx = (int)((per + 256 * 16 - 1) / (256 * 16));
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  sample_sumsq_kernel<<<dim3((unsigned)bx, (unsigned)x->n), 256, 0, st>>>(to_view(x), out);
  B200_LAUNCH_CHECK("sample_sumsq_kernel");
  return 0;
}

int gp_from_norms(const double* sumsq, int n, float lambda, float* gp, float* coeff, cudaStream_t st) {
  gp_from_norms_kernel<<<1, 256, 0, st>>>(sumsq, n, lambda, gp, coeff);
  B200_LAUNCH_CHECK("gp_from_norms_kernel");
  return 0;
}

int gp_sample_axpby(const b200gan_view* x, const float* a, const b200gan_view* y, const float* b, const b200gan_view* out, cudaStream_t st) {
  B200_CHECK_ARG(same_extent(x, out) && (!y || same_extent(x, y)), "sample_axpby: views differ in extent");
  {
    const int per_vec = x->dtype == B200GAN_F32 ? 4 : 8;
    const int64_t per = (int64_t)x->h * x->w * x->c;
    bool fast = per % per_vec == 0 && dense_view(x) && dense_view(out) && out->dtype == x->dtype && (!y || (dense_view(y) && y->dtype == x->dtype));
    if (fast) {
      const int64_t vecs = (int64_t)x->n * per / per_vec;
      int64_t blocks = (vecs + 255) / 256;
      if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
      if (x->dtype == B200GAN_F32)
        sample_axpby_dense_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const uint4*)x->ptr, a, y ? (const uint4*)y->ptr : nullptr, b, (uint4*)out->ptr, vecs, per / per_vec);
      else
        sample_axpby_dense_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((const uint4*)x->ptr, a, y ? (const uint4*)y->ptr : nullptr, b, (uint4*)out->ptr, vecs,
                                                                                  per / per_vec);
      B200_LAUNCH_CHECK("sample_axpby_dense_kernel");
      return 0;
    }
  }
  const int nrows = x->n * x->h;
  int rows_per_cta = (nrows + 16 * kNumSMs - 1) / (16 * kNumSMs);
  if (rows_per_cta < 1) rows_per_cta = 1;
  const unsigned blocks = (unsigned)((nrows + rows_per_cta - 1) / rows_per_cta);
  sample_axpby_kernel<<<blocks, 256, 0, st>>>(to_view(x), a, y ? to_view(y) : to_view(x), y != nullptr, b, to_view(out), rows_per_cta);
  B200_LAUNCH_CHECK("sample_axpby_kernel");
  return 0;
}

int gp_mean_f32(const float* x, int64_t n, float scale, float* out, cudaStream_t st) {
  mean_f32_kernel<<<1, 256, 0, st>>>(x, n, scale, out);
  B200_LAUNCH_CHECK("mean_f32_kernel");
  return 0;
}

}  // namespace b200gan
