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

__device__ __forceinline__ void decode(const View& v, int64_t idx, int& c, int64_t& off) {
  c = (int)(idx % v.c);
  int64_t pix = idx / v.c;
  const int w = (int)(pix % v.w); pix /= v.w;
  const int h = (int)(pix % v.h);
  const int64_t n = pix / v.h;
  off = n * v.sn + (int64_t)h * v.sh + (int64_t)w * v.sw + (int64_t)c * v.sc;
}
__device__ __forceinline__ int64_t offset_like(const View& v, int64_t idx) {
  int c; int64_t off;
  decode(v, idx, c, off);
  return off;
}
__device__ __forceinline__ void st_rt(void* base, int dtype, int64_t off, float x) {
  if (dtype == B200GAN_F32) reinterpret_cast<float*>(base)[off] = x;
  else reinterpret_cast<__nv_bfloat16*>(base)[off] = __float2bfloat16_rn(x);
}

// sums[0..C) = sum r, [C..2C) = sum r xhat, [2C..3C) = sum r dz.  One CTA = a contiguous chunk of elements; per-channel partial sums in
// shared memory (fp32), one fp64 atomic per channel, quantity and CTA.
__global__ void __launch_bounds__(256) bn_bwd_bwd_reduce_kernel(View r, View y, View dz, const float* __restrict__ mean, const float* __restrict__ invstd,
                                                               double* __restrict__ sums, int64_t total, int64_t chunk) {
  extern __shared__ float acc[];                 // [3][C]
  const int C = r.c;
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int64_t lo = (int64_t)blockIdx.x * chunk, hi = min(total, lo + chunk);
  for (int64_t idx = lo + threadIdx.x; idx < hi; idx += blockDim.x) {
    int c; int64_t off;
    decode(r, idx, c, off);
    const float rv = ld_rt(r.ptr, r.dtype, off);
    const float yv = ld_rt(y.ptr, y.dtype, offset_like(y, idx));
    const float dv = ld_rt(dz.ptr, dz.dtype, offset_like(dz, idx));
    const float xh = (yv - mean[c]) * invstd[c];
    atomicAdd(&acc[c], rv);
    atomicAdd(&acc[C + c], rv * xh);
    atomicAdd(&acc[2 * C + c], rv * dv);
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
                                                              float* __restrict__ dgamma, int64_t total) {
  extern __shared__ float coef[];                // [8][C]: mr, mrx, mrp, m1, m2, gamma*invstd, invstd, mean
  const int C = r.c;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double sr = sums[c], srx = sums[C + c], srd = sums[2 * C + c];
    const double m1 = dzs[c] / count, m2 = dzs[C + c] / count;
    const double srp = srd - m1 * sr - m2 * srx;             // sum r P(dz)
    coef[c] = (float)(sr / count); coef[C + c] = (float)(srx / count); coef[2 * C + c] = (float)(srp / count);
    coef[3 * C + c] = (float)m1; coef[4 * C + c] = (float)m2;
    coef[5 * C + c] = gamma[c] * invstd[c]; coef[6 * C + c] = invstd[c]; coef[7 * C + c] = mean[c];
    if (blockIdx.x == 0 && dgamma) dgamma[c] += (float)((double)invstd[c] * srp);
  }
  __syncthreads();
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int c; int64_t off;
    decode(r, idx, c, off);
    const float rv = ld_rt(r.ptr, r.dtype, off);
    const float yv = ld_rt(y.ptr, y.dtype, offset_like(y, idx));
    const float dv = ld_rt(dz.ptr, dz.dtype, offset_like(dz, idx));
    const float is = coef[6 * C + c], gs = coef[5 * C + c];
    const float xh = (yv - coef[7 * C + c]) * is;
    const float pr = rv - coef[c] - xh * coef[C + c];
    const float pdz = dv - coef[3 * C + c] - xh * coef[4 * C + c];
    const float z = fmaf(yv, scale[c], shift[c]);
    const float d = act == B200GAN_ACT_RELU ? (z > 0.f ? 1.f : 0.f) : (act == B200GAN_ACT_LRELU ? (z > 0.f ? 1.f : slope) : 1.f);
    st_rt(u.ptr, u.dtype, offset_like(u, idx), gs * pr * d);
    st_rt(inj.ptr, inj.dtype, offset_like(inj, idx), -(gs * is) * (xh * coef[2 * C + c] + coef[4 * C + c] * pr + coef[C + c] * pdz));
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

// out(n,.) = a[n] x(n,.) + b[n] y(n,.)   (y == nullptr: no second term).  a / b may be nullptr (= 1).
__global__ void __launch_bounds__(256) sample_axpby_kernel(View x, const float* __restrict__ a, View y, bool has_y, const float* __restrict__ b, View out,
                                                          int64_t total) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int c; int64_t off;
    decode(x, idx, c, off);
    const int64_t n = idx / ((int64_t)x.h * x.w * x.c);
    float v = (a ? a[n] : 1.f) * ld_rt(x.ptr, x.dtype, off);
    if (has_y) v = fmaf(b ? b[n] : 1.f, ld_rt(y.ptr, y.dtype, offset_like(y, idx)), v);
    st_rt(out.ptr, out.dtype, offset_like(out, idx), v);
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

bool same_extent(const b200gan_view* a, const b200gan_view* b) { return a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c; }

}  // namespace

int gp_bn_bwd_bwd(const b200gan_view* r, const b200gan_view* y, const b200gan_view* dz, const float* scale, const float* shift, const float* mean,
                  const float* invstd, const float* gamma, const double* dz_sums, int64_t count, int act, float slope, const b200gan_view* u,
                  const b200gan_view* inj, float* dgamma, double* sums3, cudaStream_t st) {
  B200_CHECK_ARG(same_extent(r, y) && same_extent(r, dz) && same_extent(r, u) && same_extent(r, inj), "bn_bwd_bwd: views differ in extent");
  const int C = r->c;
  B200_CHECK_ARG(C <= 1024, "bn_bwd_bwd: at most 1024 channels (got %d)", C);
  const int64_t total = (int64_t)r->n * r->h * r->w * C;
  B200_CUDA(cudaMemsetAsync(sums3, 0, sizeof(double) * 3 * C, st));
  int64_t blocks = (total + 256 * 64 - 1) / (256 * 64);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  if (blocks < 1) blocks = 1;
  const int64_t chunk = (total + blocks - 1) / blocks;
  bn_bwd_bwd_reduce_kernel<<<(unsigned)blocks, 256, 3 * C * sizeof(float), st>>>(to_view(r), to_view(y), to_view(dz), mean, invstd, sums3, total, chunk);
  B200_LAUNCH_CHECK("bn_bwd_bwd_reduce_kernel");
  bn_bwd_bwd_apply_kernel<<<(unsigned)blocks, 256, 8 * C * sizeof(float), st>>>(to_view(r), to_view(y), to_view(dz), scale, shift, mean, invstd, gamma, dz_sums,
                                                                               sums3, (double)count, act, slope, to_view(u), to_view(inj), dgamma, total);
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
  const int64_t total = (int64_t)x->n * x->h * x->w * x->c;
  int64_t blocks = (total + 256 * 8 - 1) / (256 * 8);
  if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
  if (blocks < 1) blocks = 1;
  sample_axpby_kernel<<<(unsigned)blocks, 256, 0, st>>>(to_view(x), a, y ? to_view(y) : to_view(x), y != nullptr, b, to_view(out), total);
  B200_LAUNCH_CHECK("sample_axpby_kernel");
  return 0;
}

int gp_mean_f32(const float* x, int64_t n, float scale, float* out, cudaStream_t st) {
  mean_f32_kernel<<<1, 256, 0, st>>>(x, n, scale, out);
  B200_LAUNCH_CHECK("mean_f32_kernel");
  return 0;
}

}  // namespace b200gan
