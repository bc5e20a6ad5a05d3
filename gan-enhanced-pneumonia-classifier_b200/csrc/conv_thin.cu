// Dedicated CUDA-core kernels for the layers of the DCGAN that are NOT tensor-core shaped and are bound by
// HBM traffic / fp32 FMA issue rather than by MMA throughput:
//   * the image-side layers D0 = Conv2d(nc->32) (dcgan.py:65) and G5 = ConvTranspose2d(32->nc) (dcgan.py:46),
//     nc in {1,3}: "thin" k4 s2 p1 convolutions between a 32-channel bf16 NHWC tensor (coarse side) and the
//     nc-channel image (fine side, any dtype / strides -- the reference's NCHW fp32 tensors are read in place);
//   * the latent projection G0 = ConvTranspose2d(nz->C, k7 s1 p0) on a 1x1 input (dcgan.py:26): a plain GEMM;
//   * the final D5 = Conv2d(C->1, k7 s1 p0) on a 7x7 input (dcgan.py:84): a GEMV per image.
// Each wrapper returns 1 when the problem is not of its shape (the dispatcher then uses the generic kernels).
#include "common.cuh"

namespace b200gan {

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

static bool dense_bf16_c(const b200gan_view* v, int c) {
  return v->dtype == B200GAN_BF16 && v->c == c && v->sc == 1 && v->sw == c && v->sh == (int64_t)v->w * c &&
         v->sn == (int64_t)v->h * v->w * c && (reinterpret_cast<uintptr_t>(v->ptr) & 15) == 0;
}

// ---------------------------------------------------------------------------------------------------
// thin DOWN: coarse[n,oh,ow,co] = act( sum_{kh,kw,ci} fine[n,2oh-1+kh,2ow-1+kw,ci] * w[co,ci,kh,kw] ), co < 32
// one thread = one coarse pixel x 32 channels; weights broadcast from shared memory.
// ---------------------------------------------------------------------------------------------------
template <int NC, typename TF>
__global__ void __launch_bounds__(256, 2) thin_down_kernel(View fine, const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                                        int N, int H, int W, int act, float slope) {
  __shared__ __align__(16) float ws[16 * NC][32];
  for (int i = threadIdx.x; i < 512 * NC; i += blockDim.x) {
    const int co = i & 31, r = i >> 5, tap = r / NC, ci = r - tap * NC;
    ws[r][co] = w[(co * NC + ci) * 16 + tap];
  }
  __syncthreads();
  const int64_t M = (int64_t)N * H * W;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    const int ow = (int)(m % W);
    const int64_t t = m / W;
    const int oh = (int)(t % H), n = (int)(t / H);
    const TF* base = reinterpret_cast<const TF*>(fine.ptr) + (int64_t)n * fine.sn;
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const int ih = 2 * oh - 1 + kh;
      if ((unsigned)ih >= (unsigned)fine.h) continue;
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) {
        const int iw = 2 * ow - 1 + kw;
        if ((unsigned)iw >= (unsigned)fine.w) continue;
#pragma unroll
        for (int ci = 0; ci < NC; ++ci) {
          const float x = ld_as_float(base + (int64_t)ih * fine.sh + (int64_t)iw * fine.sw + (int64_t)ci * fine.sc);
          const float4* wr = reinterpret_cast<const float4*>(ws[(kh * 4 + kw) * NC + ci]);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 wv = wr[j];
            acc[4 * j + 0] = fmaf(x, wv.x, acc[4 * j + 0]);
            acc[4 * j + 1] = fmaf(x, wv.y, acc[4 * j + 1]);
            acc[4 * j + 2] = fmaf(x, wv.z, acc[4 * j + 2]);
            acc[4 * j + 3] = fmaf(x, wv.w, acc[4 * j + 3]);
          }
        }
      }
    }
    if (act == B200GAN_ACT_LRELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = acc[j] > 0.f ? acc[j] : acc[j] * slope;
    }
    uint4* o = reinterpret_cast<uint4*>(out + m * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = pack8(acc + 8 * j);
  }
}

// fine: (N,2H,2W,nc) any layout; coarse: (N,H,W,32) dense bf16.  act: NONE or LRELU (fused, dcgan.py:66)
int thin_down(const b200gan_view* fine, const float* w, const b200gan_view* coarse, int act, float slope, cudaStream_t st) {
  if (!dense_bf16_c(coarse, 32) || (fine->c != 1 && fine->c != 3)) return 1;
  const int64_t M = (int64_t)coarse->n * coarse->h * coarse->w;
  int64_t nb = (M + 255) / 256;
  if (nb > 32 * kNumSMs) nb = 32 * kNumSMs;
  const View f = to_view(fine);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(coarse->ptr);
#define LAUNCH(NC, T) thin_down_kernel<NC, T><<<(unsigned)nb, 256, 0, st>>>(f, w, o, coarse->n, coarse->h, coarse->w, act, slope)
  if (fine->c == 1) { if (fine->dtype == B200GAN_F32) LAUNCH(1, float); else LAUNCH(1, __nv_bfloat16); }
  else { if (fine->dtype == B200GAN_F32) LAUNCH(3, float); else LAUNCH(3, __nv_bfloat16); }
#undef LAUNCH
  B200_LAUNCH_CHECK("thin_down_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// thin UP: fine[n,2q+py,2r+px,ci] = act( sum_{co} sum_{taps} coarse[n,q+di,r+dj,co] * w[co,ci,kh,kw] ), kh = py+1-2di
// one thread = one coarse position -> the 2x2 block of fine pixels (3x3 coarse neighbourhood, 16 taps).
// ---------------------------------------------------------------------------------------------------
template <int NC, typename TF>
__global__ void __launch_bounds__(256) thin_up_kernel(const __nv_bfloat16* __restrict__ coarse, const float* __restrict__ w, View fine,
                                                      int N, int H, int W, int act) {
  __shared__ __align__(16) float ws[16 * NC][32];     // [tap*NC+ci][co]
  for (int i = threadIdx.x; i < 512 * NC; i += blockDim.x) {
    const int co = i & 31, r = i >> 5, tap = r / NC, ci = r - tap * NC;
    ws[r][co] = w[(co * NC + ci) * 16 + tap];
  }
  __syncthreads();
  const int64_t M = (int64_t)N * H * W;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(m % W);
    const int64_t t = m / W;
    const int q = (int)(t % H), n = (int)(t / H);
    float acc[2][2][NC];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[a][b][c] = 0.f;
#pragma unroll
    for (int di = -1; di <= 1; ++di) {
      const int iy = q + di;
      if ((unsigned)iy >= (unsigned)H) continue;
#pragma unroll
      for (int dj = -1; dj <= 1; ++dj) {
        const int ix = r + dj;
        if ((unsigned)ix >= (unsigned)W) continue;
        const uint4* src = reinterpret_cast<const uint4*>(coarse + (((int64_t)n * H + iy) * W + ix) * 32);
        float v[32];
#pragma unroll
        for (int j = 0; j < 4; ++j) unpack8(__ldg(src + j), v + 8 * j);
#pragma unroll
        for (int py = 0; py < 2; ++py) {
          const int kh = py + 1 - 2 * di;
          if (kh < 0 || kh > 3) continue;
#pragma unroll
          for (int px = 0; px < 2; ++px) {
            const int kw = px + 1 - 2 * dj;
            if (kw < 0 || kw > 3) continue;
#pragma unroll
            for (int ci = 0; ci < NC; ++ci) {
              const float4* wr = reinterpret_cast<const float4*>(ws[(kh * 4 + kw) * NC + ci]);
              float s = 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 wv = wr[j];
                s = fmaf(v[4 * j + 0], wv.x, s);
                s = fmaf(v[4 * j + 1], wv.y, s);
                s = fmaf(v[4 * j + 2], wv.z, s);
                s = fmaf(v[4 * j + 3], wv.w, s);
              }
              acc[py][px][ci] += s;
            }
          }
        }
      }
    }
    TF* base = reinterpret_cast<TF*>(fine.ptr) + (int64_t)n * fine.sn;
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
      for (int px = 0; px < 2; ++px)
#pragma unroll
        for (int ci = 0; ci < NC; ++ci) {
          float o = acc[py][px][ci];
          if (act == B200GAN_ACT_TANH) o = tanhf(o);
          st_from_float(base + (int64_t)(2 * q + py) * fine.sh + (int64_t)(2 * r + px) * fine.sw + (int64_t)ci * fine.sc, o);
        }
  }
}

// coarse: (N,H,W,32) dense bf16; fine: (N,2H,2W,nc) any layout.  act: NONE or TANH (fused, dcgan.py:47)
int thin_up(const b200gan_view* coarse, const float* w, const b200gan_view* fine, int act, cudaStream_t st) {
  if (!dense_bf16_c(coarse, 32) || (fine->c != 1 && fine->c != 3)) return 1;
  const int64_t M = (int64_t)coarse->n * coarse->h * coarse->w;
  int64_t nb = (M + 255) / 256;
  if (nb > 32 * kNumSMs) nb = 32 * kNumSMs;
  const View f = to_view(fine);
  const __nv_bfloat16* c = reinterpret_cast<const __nv_bfloat16*>(coarse->ptr);
#define LAUNCH(NC, T) thin_up_kernel<NC, T><<<(unsigned)nb, 256, 0, st>>>(c, w, f, coarse->n, coarse->h, coarse->w, act)
  if (fine->c == 1) { if (fine->dtype == B200GAN_F32) LAUNCH(1, float); else LAUNCH(1, __nv_bfloat16); }
  else { if (fine->dtype == B200GAN_F32) LAUNCH(3, float); else LAUNCH(3, __nv_bfloat16); }
#undef LAUNCH
  B200_LAUNCH_CHECK("thin_up_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// thin WGRAD: dw[co,ci,kh,kw] += sum_{n,oh,ow} coarse[n,oh,ow,co] * fine[n,2oh-1+kh,2ow-1+kw,ci]
// block = 8 pixel slices x 32 lanes; lane = (4 output channels) x (one kernel row kh, 4 kw) x NC.
// ---------------------------------------------------------------------------------------------------
template <int NC, typename TF>
__global__ void __launch_bounds__(256) thin_wgrad_kernel(const __nv_bfloat16* __restrict__ coarse, View fine, float* __restrict__ dw,
                                                         int N, int H, int W, int64_t pix_per_block) {
  __shared__ float red[512 * NC];
  for (int i = threadIdx.x; i < 512 * NC; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int slice = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cog = lane & 7, kh = lane >> 3;
  const int64_t P = (int64_t)N * H * W;
  const int64_t pbeg = (int64_t)blockIdx.x * pix_per_block;
  const int64_t pend = pbeg + pix_per_block < P ? pbeg + pix_per_block : P;
  float acc[4][4][NC];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[a][b][c] = 0.f;
  for (int64_t p = pbeg + slice; p < pend; p += 8) {
    const int ow = (int)(p % W);
    const int64_t t = p / W;
    const int oh = (int)(t % H), n = (int)(t / H);
    const int ih = 2 * oh - 1 + kh;
    if ((unsigned)ih >= (unsigned)fine.h) continue;
    const uint2 cr = __ldg(reinterpret_cast<const uint2*>(coarse + p * 32 + cog * 4));
    const float c4[4] = {__uint_as_float(cr.x << 16), __uint_as_float(cr.x & 0xffff0000u), __uint_as_float(cr.y << 16),
                         __uint_as_float(cr.y & 0xffff0000u)};
    const TF* row = reinterpret_cast<const TF*>(fine.ptr) + (int64_t)n * fine.sn + (int64_t)ih * fine.sh;
#pragma unroll
    for (int kw = 0; kw < 4; ++kw) {
      const int iw = 2 * ow - 1 + kw;
      if ((unsigned)iw >= (unsigned)fine.w) continue;
#pragma unroll
      for (int ci = 0; ci < NC; ++ci) {
        const float x = ld_as_float(row + (int64_t)iw * fine.sw + (int64_t)ci * fine.sc);
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[a][kw][ci] = fmaf(c4[a], x, acc[a][kw][ci]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int kw = 0; kw < 4; ++kw)
#pragma unroll
      for (int ci = 0; ci < NC; ++ci) atomicAdd(&red[((cog * 4 + a) * NC + ci) * 16 + kh * 4 + kw], acc[a][kw][ci]);
  __syncthreads();
  for (int i = threadIdx.x; i < 512 * NC; i += blockDim.x) atomicAdd(dw + i, red[i]);
}

// coarse: (N,H,W,32) dense bf16; fine: (N,2H,2W,nc) any layout; dw: (32,nc,4,4) fp32 accumulated
int thin_wgrad(const b200gan_view* fine, const b200gan_view* coarse, float* dw, cudaStream_t st) {
  if (!dense_bf16_c(coarse, 32) || (fine->c != 1 && fine->c != 3)) return 1;
  const int64_t P = (int64_t)coarse->n * coarse->h * coarse->w;
  int64_t nb = 8 * kNumSMs;
  if (nb > (P + 63) / 64) nb = (P + 63) / 64;
  if (nb < 1) nb = 1;
  const int64_t ppb = (P + nb - 1) / nb;
  nb = (P + ppb - 1) / ppb;
  const View f = to_view(fine);
  const __nv_bfloat16* c = reinterpret_cast<const __nv_bfloat16*>(coarse->ptr);
#define LAUNCH(NC, T) thin_wgrad_kernel<NC, T><<<(unsigned)nb, 256, 0, st>>>(c, f, dw, coarse->n, coarse->h, coarse->w, ppb)
  if (fine->c == 1) { if (fine->dtype == B200GAN_F32) LAUNCH(1, float); else LAUNCH(1, __nv_bfloat16); }
  else { if (fine->dtype == B200GAN_F32) LAUNCH(3, float); else LAUNCH(3, __nv_bfloat16); }
#undef LAUNCH
  B200_LAUNCH_CHECK("thin_wgrad_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// full-window convolution to one channel (D5: Conv2d(C->1, k=H=W, p0) + its gradients) as GEMV / outer product.
//   x: (N,K,K,C) dense bf16, w: (1,C,K,K) fp32.  J = K*K*C, NHWC index j = hw*C + c  <->  weight index c*K*K + hw
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) window_dot_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                             float* __restrict__ out, int J, int C, int HW) {
  __shared__ float red[8];
  const int n = blockIdx.x;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (int64_t)n * J);
  float s = 0.f;
  for (int v = threadIdx.x; v < J / 8; v += blockDim.x) {
    float xv[8];
    unpack8(__ldg(xr + v), xv);
    const int j = v * 8, hw = j / C, c = j - hw * C;
#pragma unroll
    for (int e = 0; e < 8; ++e) s = fmaf(xv[e], __ldg(w + (int64_t)(c + e) * HW + hw), s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    out[n] = t;
  }
}

// dx[n][j] = dl[n] * w[j]  (optional);  dw[j] += sum_n dl[n] * x[n][j]  (optional)
__global__ void __launch_bounds__(256) window_dot_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ dl, __nv_bfloat16* __restrict__ dx,
                                                             float* __restrict__ dw, int N, int J, int C, int HW, int n_per_block) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= J / 8) return;
  const int j = v * 8, hw = j / C, c = j - hw * C;
  float wv[8], acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { wv[e] = w ? __ldg(w + (int64_t)(c + e) * HW + hw) : 0.f; acc[e] = 0.f; }
  const int n0 = blockIdx.y * n_per_block, n1 = min(n0 + n_per_block, N);
  for (int n = n0; n < n1; ++n) {
    const float d = __ldg(dl + n);
    if (dw) {
      float xv[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + (int64_t)n * J) + v), xv);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(d, xv[e], acc[e]);
    }
    if (dx) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = d * wv[e];
      reinterpret_cast<uint4*>(dx + (int64_t)n * J)[v] = pack8(o);
    }
  }
  if (dw) {
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(dw + (int64_t)(c + e) * HW + hw, acc[e]);
  }
}

static bool window_shape(const b200gan_conv* cv, const b200gan_view* fine, const b200gan_view* coarse) {
  return cv->stride == 1 && cv->pad == 0 && fine->h == cv->k && fine->w == cv->k && coarse->h == 1 && coarse->w == 1 &&
         coarse->c == 1 && dense_bf16_c(fine, fine->c) && fine->c % 8 == 0 && coarse->dtype == B200GAN_F32 && coarse->sn == 1;
}

int window_fprop(const b200gan_conv* cv, const b200gan_view* x, const float* w, const b200gan_view* y, cudaStream_t st) {
  if (!window_shape(cv, x, y)) return 1;
  window_dot_fwd_kernel<<<x->n, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x->ptr), w, reinterpret_cast<float*>(y->ptr),
                                              x->h * x->w * x->c, x->c, x->h * x->w);
  B200_LAUNCH_CHECK("window_dot_fwd_kernel");
  return 0;
}

static int window_bwd(const b200gan_conv* cv, const b200gan_view* x, const float* w, const b200gan_view* dy, __nv_bfloat16* dx,
                      float* dw, cudaStream_t st) {
  const int J = x->h * x->w * x->c;
  const int npb = 16;
  dim3 grid((J / 8 + 255) / 256, (x->n + npb - 1) / npb);
  window_dot_bwd_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x->ptr), w, reinterpret_cast<const float*>(dy->ptr),
                                              dx, dw, x->n, J, x->c, x->h * x->w, npb);
  B200_LAUNCH_CHECK("window_dot_bwd_kernel");
  return 0;
}
int window_dgrad(const b200gan_conv* cv, const b200gan_view* dy, const float* w, const b200gan_view* dx, cudaStream_t st) {
  if (!window_shape(cv, dx, dy)) return 1;
  return window_bwd(cv, dx, w, dy, reinterpret_cast<__nv_bfloat16*>(dx->ptr), nullptr, st);
}
int window_wgrad(const b200gan_conv* cv, const b200gan_view* x, const b200gan_view* dy, float* dw, cudaStream_t st) {
  if (!window_shape(cv, x, dy)) return 1;
  return window_bwd(cv, x, nullptr, dy, nullptr, dw, st);
}

// ---------------------------------------------------------------------------------------------------
// latent projection (G0: ConvTranspose2d(nz->C, k, s1, p0) on a 1x1 input) = GEMM  y[n][(hw,co)] = z[n][:] . w[:, co, hw]
// and its weight gradient dw[ci][co][hw] += sum_n z[n][ci] * dy[n][(hw,co)].  64x64x16 smem-tiled SIMT GEMM.
// z: (N,1,1,nz) any layout/dtype; y, dy: (N,k,k,C) dense bf16; w: (nz, C, k, k) fp32.
// ---------------------------------------------------------------------------------------------------
template <typename TZ>
__global__ void __launch_bounds__(256) latent_fprop_kernel(View z, const float* __restrict__ w, __nv_bfloat16* __restrict__ y, int N,
                                                           int NZ, int C, int HW) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x, m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int NC = HW * C;
  const int a_row = tid >> 2, a_k = (tid & 3) * 4;
  const int b_col = tid & 63, b_k = (tid >> 6) * 4;
  const int col = n0 + b_col, hw = col / C, co = col - hw * C;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < NZ; k0 += BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int kk = k0 + a_k + e, m = m0 + a_row;
      As[a_k + e][a_row] = (m < N && kk < NZ) ? ld_as_float(reinterpret_cast<const TZ*>(z.ptr) + (int64_t)m * z.sn + (int64_t)kk * z.sc) : 0.f;
      const int kb = k0 + b_k + e;
      Bs[b_k + e][b_col] = (col < NC && kb < NZ) ? __ldg(w + ((int64_t)kb * C + co) * HW + hw) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c < NC) y[(int64_t)m * NC + c] = __float2bfloat16_rn(acc[i][j]);
    }
  }
}

template <typename TZ>
__global__ void __launch_bounds__(256) latent_wgrad_kernel(View z, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw, int N,
                                                           int NZ, int C, int HW) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ __align__(16) float As[BK][BM + 4];     // [n][ci]
  __shared__ __align__(16) float Bs[BK][BN + 4];     // [n][col]
  const int tid = threadIdx.x, m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int NC = HW * C;
  const int lk = tid >> 4, l4 = (tid & 15) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < N; k0 += BK) {
    const int n = k0 + lk;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ci = m0 + l4 + e, col = n0 + l4 + e;
      As[lk][l4 + e] = (n < N && ci < NZ) ? ld_as_float(reinterpret_cast<const TZ*>(z.ptr) + (int64_t)n * z.sn + (int64_t)ci * z.sc) : 0.f;
      Bs[lk][l4 + e] = (n < N && col < NC) ? __bfloat162float(dy[(int64_t)n * NC + col]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = m0 + ty * 4 + i;
    if (ci >= NZ) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= NC) continue;
      const int hw = col / C, co = col - hw * C;
      dw[((int64_t)ci * C + co) * HW + hw] += acc[i][j];      // each output owned by exactly one thread: no atomics
    }
  }
}

static bool latent_shape(const b200gan_conv* cv, const b200gan_view* fine, const b200gan_view* coarse) {
  return cv->stride == 1 && cv->pad == 0 && coarse->h == 1 && coarse->w == 1 && fine->h == cv->k && fine->w == cv->k &&
         dense_bf16_c(fine, fine->c);
}

// ConvTranspose2d forward seen from the conv geometry: coarse = z (N,1,1,nz), fine = y (N,k,k,C)
int latent_fprop(const b200gan_conv* cv, const b200gan_view* z, const float* w, const b200gan_view* y, cudaStream_t st) {
  if (!latent_shape(cv, y, z)) return 1;
  const int HW = y->h * y->w, C = y->c;
  dim3 grid((z->n + 63) / 64, (HW * C + 63) / 64);
  const View zv = to_view(z);
  if (z->dtype == B200GAN_F32)
    latent_fprop_kernel<float><<<grid, 256, 0, st>>>(zv, w, reinterpret_cast<__nv_bfloat16*>(y->ptr), z->n, z->c, C, HW);
  else
    latent_fprop_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(zv, w, reinterpret_cast<__nv_bfloat16*>(y->ptr), z->n, z->c, C, HW);
  B200_LAUNCH_CHECK("latent_fprop_kernel");
  return 0;
}

int latent_wgrad(const b200gan_conv* cv, const b200gan_view* dy_fine, const b200gan_view* z, float* dw, cudaStream_t st) {
  if (!latent_shape(cv, dy_fine, z)) return 1;
  const int HW = dy_fine->h * dy_fine->w, C = dy_fine->c;
  dim3 grid((z->c + 63) / 64, (HW * C + 63) / 64);
  const View zv = to_view(z);
  if (z->dtype == B200GAN_F32)
    latent_wgrad_kernel<float><<<grid, 256, 0, st>>>(zv, reinterpret_cast<const __nv_bfloat16*>(dy_fine->ptr), dw, z->n, z->c, C, HW);
  else
    latent_wgrad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(zv, reinterpret_cast<const __nv_bfloat16*>(dy_fine->ptr), dw, z->n, z->c, C, HW);
  B200_LAUNCH_CHECK("latent_wgrad_kernel");
  return 0;
}

}  // namespace b200gan
