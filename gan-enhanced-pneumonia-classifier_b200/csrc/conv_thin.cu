// Dedicated CUDA-core kernels for two layers of the DCGAN that are NOT tensor-core shaped and are bound by
// HBM traffic rather than by MMA throughput (the image-side layers live in conv_thin_mma.cu):
//   * the latent projection G0 = ConvTranspose2d(nz->C, k7 s1 p0) on a 1x1 input (dcgan.py:26): a plain GEMM;
//   * the final D5 = Conv2d(C->1, k7 s1 p0) on a 7x7 input (dcgan.py:84): a GEMV per image.
// Each wrapper returns 1 when the problem is not of its shape (the dispatcher then uses the generic kernels).
#include "common.cuh"

namespace b200gan {

static bool dense_bf16_c(const b200gan_view* v, int c) {
  return v->dtype == B200GAN_BF16 && v->c == c && v->sc == 1 && v->sw == c && v->sh == (int64_t)v->w * c &&
         v->sn == (int64_t)v->h * v->w * c && (reinterpret_cast<uintptr_t>(v->ptr) & 15) == 0;
}

// ---------------------------------------------------------------------------------------------------
// "score map": valid convolution of a SMALL map to ONE channel, Conv2d(C -> 1, k, stride 1, pad 0) with an output map of up to 8 x 8
// (the WGAN-GP critic's last layer, wggan.py:63: 512 x 14 x 14 -> 1 x 8 x 8) and its gradients.  As an implicit GEMM this has N = 1: the
// generic kernels waste 63/64 of every tile on it (measured 3.6 ms per call at batch 128).  Here one CTA owns one image: the weights sit
// transposed ([tap][C], fp32) in shared memory, a lane owns C/32 consecutive channels, warps split output rows / input pixels / taps.
// x: (N,H,W,C) dense bf16;  y, dy: (N,OH,OW,1) dense fp32;  w, dw: (1,C,k,k) fp32.
// ---------------------------------------------------------------------------------------------------
template <int CPL>
__device__ __forceinline__ void ld_channels(const __nv_bfloat16* p, float (&v)[CPL]) {
  if constexpr (CPL % 8 == 0) {
#pragma unroll
    for (int j = 0; j < CPL / 8; ++j) unpack8(*reinterpret_cast<const uint4*>(p + 8 * j), &v[8 * j]);
  } else {
#pragma unroll
    for (int j = 0; j < CPL; ++j) v[j] = __bfloat162float(p[j]);
  }
}

template <int CPL>
__global__ void __launch_bounds__(256) score_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int H, int W,
                                                       int k, int OH, int OW) {
  extern __shared__ float wt[];                                  // [k*k][C]
  constexpr int C = 32 * CPL;
  const int kk = k * k, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kk * C; i += blockDim.x) { const int c = i / kk, t = i - c * kk; wt[t * C + c] = w[i]; }
  __syncthreads();
  const __nv_bfloat16* xn = x + (int64_t)blockIdx.x * H * W * C + lane * CPL;
  for (int oh = warp; oh < OH; oh += 8) {
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = 0.f;
    for (int kh = 0; kh < k; ++kh)
      for (int px = 0; px < W; ++px) {
        float xv[CPL];
        ld_channels<CPL>(xn + ((int64_t)(oh + kh) * W + px) * C, xv);
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          const int kw = px - o;
          if (o < OW && kw >= 0 && kw < k) {
            const float* wr = wt + (kh * k + kw) * C + lane * CPL;
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) s = fmaf(xv[j], wr[j], s);
            acc[o] += s;
          }
        }
      }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const float s = warp_sum(acc[o]);
      if (lane == 0 && o < OW) y[((int64_t)blockIdx.x * OH + oh) * OW + o] = s;
    }
  }
}

template <int CPL>
__global__ void __launch_bounds__(256) score_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, __nv_bfloat16* __restrict__ dx, int H,
                                                         int W, int k, int OH, int OW) {
  extern __shared__ float wt[];                                  // [k*k][C] then dy[OH*OW]
  constexpr int C = 32 * CPL;
  const int kk = k * k, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* dys = wt + kk * C;
  for (int i = threadIdx.x; i < kk * C; i += blockDim.x) { const int c = i / kk, t = i - c * kk; wt[t * C + c] = w[i]; }
  for (int i = threadIdx.x; i < OH * OW; i += blockDim.x) dys[i] = dy[(int64_t)blockIdx.x * OH * OW + i];
  __syncthreads();
  __nv_bfloat16* dxn = dx + (int64_t)blockIdx.x * H * W * C + lane * CPL;
  for (int p = warp; p < H * W; p += 8) {
    const int ih = p / W, iw = p - ih * W;
    float acc[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) acc[j] = 0.f;
    for (int kh = 0; kh < k; ++kh) {
      const int oh = ih - kh;
      if (oh < 0 || oh >= OH) continue;
      for (int kw = 0; kw < k; ++kw) {
        const int ow = iw - kw;
        if (ow < 0 || ow >= OW) continue;
        const float g = dys[oh * OW + ow];
        const float* wr = wt + (kh * k + kw) * C + lane * CPL;
#pragma unroll
        for (int j = 0; j < CPL; ++j) acc[j] = fmaf(g, wr[j], acc[j]);
      }
    }
    __nv_bfloat16* o = dxn + (int64_t)p * C;
#pragma unroll
    for (int j = 0; j < CPL; ++j) o[j] = __float2bfloat16_rn(acc[j]);
  }
}

// dw[c][kh][kw] += sum_n sum_{oh,ow} dy[n][oh][ow] x[n][oh+kh][ow+kw][c]: a CTA walks images n = blockIdx.x, + gridDim.x, ...; warp w owns taps
// t = w, w + 8, ...; the per-CTA partial sums live in shared memory ([tap][C], one owner thread per element) and are added once at the end
template <int CPL>
__global__ void __launch_bounds__(256) score_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, int N, int H,
                                                         int W, int k, int OH, int OW) {
  extern __shared__ float accs[];                                // [k*k][C] then dy[OH*OW]
  constexpr int C = 32 * CPL;
  const int kk = k * k, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* dys = accs + kk * C;
  for (int i = threadIdx.x; i < kk * C; i += blockDim.x) accs[i] = 0.f;
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < OH * OW; i += blockDim.x) dys[i] = dy[(int64_t)n * OH * OW + i];
    __syncthreads();
    const __nv_bfloat16* xn = x + (int64_t)n * H * W * C + lane * CPL;
    for (int t = warp; t < kk; t += 8) {
      const int kh = t / k, kw = t - kh * k;
      float acc[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[j] = 0.f;
      for (int oh = 0; oh < OH; ++oh)
        for (int ow = 0; ow < OW; ++ow) {
          float xv[CPL];
          ld_channels<CPL>(xn + ((int64_t)(oh + kh) * W + ow + kw) * C, xv);
          const float g = dys[oh * OW + ow];
#pragma unroll
          for (int j = 0; j < CPL; ++j) acc[j] = fmaf(g, xv[j], acc[j]);
        }
      float* a = accs + t * C + lane * CPL;
#pragma unroll
      for (int j = 0; j < CPL; ++j) a[j] += acc[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kk * C; i += blockDim.x) {       // i = c * kk + t in the reference layout
    const int c = i / kk, t = i - c * kk;
    const float v = accs[t * C + c];
    if (v != 0.f) atomicAdd(dw + i, v);
  }
}

static bool score_shape(const b200gan_conv* cv, const b200gan_view* fine, const b200gan_view* coarse) {
  if (cv->stride != 1 || cv->pad != 0 || coarse->c != 1 || fine->c % 32 != 0 || fine->c > 512) return false;
  const int cpl = fine->c / 32;
  if (cpl != 1 && cpl != 2 && cpl != 4 && cpl != 8 && cpl != 16) return false;
  if (coarse->h != fine->h - cv->k + 1 || coarse->w != fine->w - cv->k + 1 || coarse->w > 8 || coarse->h < 1) return false;
  if (coarse->h == 1 && coarse->w == 1) return false;                      // the full-window kernels below are better at that
  return dense_bf16_c(fine, fine->c) && coarse->dtype == B200GAN_F32 && coarse->sw == 1 && coarse->sh == coarse->w &&
         coarse->sn == (int64_t)coarse->h * coarse->w;
}

#define SCORE_DISPATCH(KERNEL, GRID, SMEM, ...)                                                                  \
  do {                                                                                                           \
    const int cpl__ = fine__->c / 32;                                                                            \
    if (cpl__ == 16) { B200_CUDA((ensure_dynamic_smem<KERNEL<16>>((int)(SMEM)))); KERNEL<16><<<GRID, 256, SMEM, st>>>(__VA_ARGS__); } \
    else if (cpl__ == 8) { B200_CUDA((ensure_dynamic_smem<KERNEL<8>>((int)(SMEM)))); KERNEL<8><<<GRID, 256, SMEM, st>>>(__VA_ARGS__); }  \
    else if (cpl__ == 4) { B200_CUDA((ensure_dynamic_smem<KERNEL<4>>((int)(SMEM)))); KERNEL<4><<<GRID, 256, SMEM, st>>>(__VA_ARGS__); }  \
    else if (cpl__ == 2) { B200_CUDA((ensure_dynamic_smem<KERNEL<2>>((int)(SMEM)))); KERNEL<2><<<GRID, 256, SMEM, st>>>(__VA_ARGS__); }  \
    else { B200_CUDA((ensure_dynamic_smem<KERNEL<1>>((int)(SMEM)))); KERNEL<1><<<GRID, 256, SMEM, st>>>(__VA_ARGS__); }                    \
  } while (0)

int score_fprop(const b200gan_conv* cv, const b200gan_view* x, const float* w, const b200gan_view* y, cudaStream_t st) {
  if (!score_shape(cv, x, y)) return 1;
  const b200gan_view* fine__ = x;
  const size_t smem = (size_t)cv->k * cv->k * x->c * sizeof(float);
  SCORE_DISPATCH(score_fwd_kernel, x->n, smem, reinterpret_cast<const __nv_bfloat16*>(x->ptr), w, reinterpret_cast<float*>(y->ptr), x->h, x->w, cv->k, y->h, y->w);
  B200_LAUNCH_CHECK("score_fwd_kernel");
  return 0;
}
int score_dgrad(const b200gan_conv* cv, const b200gan_view* dy, const float* w, const b200gan_view* dx, cudaStream_t st) {
  if (!score_shape(cv, dx, dy)) return 1;
  const b200gan_view* fine__ = dx;
  const size_t smem = ((size_t)cv->k * cv->k * dx->c + dy->h * dy->w) * sizeof(float);
  SCORE_DISPATCH(score_dgrad_kernel, dx->n, smem, reinterpret_cast<const float*>(dy->ptr), w, reinterpret_cast<__nv_bfloat16*>(dx->ptr), dx->h, dx->w, cv->k, dy->h,
                 dy->w);
  B200_LAUNCH_CHECK("score_dgrad_kernel");
  return 0;
}
int score_wgrad(const b200gan_conv* cv, const b200gan_view* x, const b200gan_view* dy, float* dw, cudaStream_t st) {
  if (!score_shape(cv, x, dy)) return 1;
  const b200gan_view* fine__ = x;
  const size_t smem = ((size_t)cv->k * cv->k * x->c + dy->h * dy->w) * sizeof(float);
  const int grid = x->n < 2 * kNumSMs ? x->n : 2 * kNumSMs;
  SCORE_DISPATCH(score_wgrad_kernel, grid, smem, reinterpret_cast<const __nv_bfloat16*>(x->ptr), reinterpret_cast<const float*>(dy->ptr), dw, x->n, x->h, x->w,
                 cv->k, dy->h, dy->w);
  B200_LAUNCH_CHECK("score_wgrad_kernel");
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

// dx[n][j] = dl[n] * w[j] followed by the activation backward and the BatchNorm-backward sums of the layer below (b200gan_fuse.prev_*
// with BatchNorm): dz = dx * act'(scale*y + shift) stored, sums[c] += dz, sums[C+c] += dz * (y - mean) * invstd, from the stored
// (bf16-rounded) dz.  A thread owns 8 channels of one position for 16 samples; four saved-output loads in flight.
__global__ void __launch_bounds__(256) window_dgrad_bn_kernel(const float* __restrict__ w, const float* __restrict__ dl,
                                                              const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ dx,
                                                              const float* __restrict__ scale, const float* __restrict__ shift,
                                                              const float* __restrict__ mean, const float* __restrict__ invstd, float neg,
                                                              double* __restrict__ sums, int N, int J, int C, int HW, int n_per_block) {
  extern __shared__ float wred[];                          // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) wred[i] = 0.f;
  __syncthreads();
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < J / 8) {
    const int j = v * 8, hw = j / C, c = j - hw * C;
    float wv[8], sc[8], sh[8], mu[8], s0[8], s1[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      wv[e] = __ldg(w + (int64_t)(c + e) * HW + hw);
      sc[e] = scale[c + e]; sh[e] = shift[c + e]; mu[e] = mean[c + e];
      s0[e] = 0.f; s1[e] = 0.f;
    }
    const int n0 = blockIdx.y * n_per_block, n1 = min(n0 + n_per_block, N);
    for (int nb = n0; nb < n1; nb += 4) {
      uint4 ry[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (nb + u < n1) ry[u] = __ldg(reinterpret_cast<const uint4*>(y + (int64_t)(nb + u) * J) + v);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int n = nb + u;
        if (n >= n1) break;
        const float d = __ldg(dl + n);
        float yv[8], o[8];
        unpack8(ry[u], yv);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = d * wv[e] * (fmaf(yv[e], sc[e], sh[e]) > 0.f ? 1.f : neg);
        const uint4 packed = pack8(o);
        reinterpret_cast<uint4*>(dx + (int64_t)n * J)[v] = packed;
        unpack8(packed, o);                                // the sums are those of the stored values
#pragma unroll
        for (int e = 0; e < 8; ++e) { s0[e] += o[e]; s1[e] = fmaf(o[e], yv[e] - mu[e], s1[e]); }
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) { atomicAdd(&wred[c + e], s0[e]); atomicAdd(&wred[C + c + e], s1[e]); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const float t = wred[i];
    if (t != 0.f) atomicAdd(sums + i, (double)t * (i >= C ? (double)invstd[i - C] : 1.0));
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
// epi: mode 0 (none) or 2 (activation backward + BatchNorm-backward sums of the layer below); other modes return 1
int window_dgrad(const b200gan_conv* cv, const b200gan_view* dy, const float* w, const b200gan_view* dx, const TcEpi& epi, cudaStream_t st) {
  if (!window_shape(cv, dx, dy)) return 1;
  if (epi.mode == 0) return window_bwd(cv, dx, w, dy, reinterpret_cast<__nv_bfloat16*>(dx->ptr), nullptr, st);
  if (epi.mode != 2 || !dense_bf16_c(epi.prev_y, dx->c) || epi.prev_y->n != dx->n || epi.prev_y->h != dx->h || epi.prev_y->w != dx->w || dx->c > 2048) return 1;
  const int J = dx->h * dx->w * dx->c, npb = 16;
  B200_CUDA(cudaMemsetAsync(epi.sums, 0, sizeof(double) * 2 * dx->c, st));
  dim3 grid((J / 8 + 255) / 256, (dx->n + npb - 1) / npb);
  const float neg = epi.act == B200GAN_ACT_RELU ? 0.f : (epi.act == B200GAN_ACT_LRELU ? epi.slope : 1.f);
  window_dgrad_bn_kernel<<<grid, 256, 2 * dx->c * sizeof(float), st>>>(w, reinterpret_cast<const float*>(dy->ptr),
      reinterpret_cast<const __nv_bfloat16*>(epi.prev_y->ptr), reinterpret_cast<__nv_bfloat16*>(dx->ptr), epi.scale, epi.shift, epi.mean, epi.invstd,
      neg, epi.sums, dx->n, J, dx->c, dx->h * dx->w, npb);
  B200_LAUNCH_CHECK("window_dgrad_bn_kernel");
  return 0;
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
