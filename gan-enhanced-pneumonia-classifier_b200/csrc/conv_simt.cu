// CUDA-core (SIMT) implicit-GEMM convolution kernels: the any-shape, fp32-accumulating path of
// libb200gan.so.  They serve (a) the fp32 parity mode (rtol 1e-4 against the reference's CPU fp32 path),
// (b) the layers that are not tensor-core shaped (nc-channel image layers, the latent GEMM, the final
// 7x7 GEMV) in bf16 mode.  One gather-GEMM kernel covers Conv2d fprop, Conv2d dgrad (one launch per
// output-parity class, so no multiply-by-zero work for stride 2) and, through the role swap
// ConvTranspose2d == conv-dgrad, all of ConvTranspose2d; one split-K kernel covers both weight gradients.
//
// Reference semantics: torch conv2d / conv_transpose2d and their autograd, called from dcgan.py:26-46,65-84.
#include "common.cuh"

namespace b200gan {

struct GatherParams {
  int N, QH, QW;                       // GEMM rows: (n, qh, qw)
  const void* in;                      // gathered operand
  int64_t i_sn, i_sh, i_sw, i_sc;
  int IH, IW;
  int i_mul, i_base_h, i_base_w, i_tstep;   // ih = qh*i_mul + i_base_h + jh*i_tstep
  int TH, TW, CK, CN;                  // taps, reduction channels, output channels
  const float* w;                      // fp32 master weight (Co, Ci, k, k)
  int64_t w_scn, w_sck;                // strides of the output-channel / reduction-channel index
  int w_base_h, w_base_w, w_step, ksize;    // kh = w_base_h + jh*w_step
  void* out;
  int64_t o_sn, o_sh, o_sw, o_sc;
  int o_mul, o_off_h, o_off_w;         // output pixel = (qh*o_mul + o_off_h, qw*o_mul + o_off_w)
  // fusions (ConvFuse): gathered operand *= act'(ref) with ref addressed like `in`; activation on the result
  const void* ref; int ref_dtype; int64_t r_sn, r_sh, r_sw, r_sc;
  int g_act; float g_slope;
  int out_act; float out_slope;
};

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) gather_gemm_kernel(GatherParams p) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t M = (int64_t)p.N * p.QH * p.QW;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int K = p.TH * p.TW * p.CK;

  const int a_row = tid >> 2, a_k = (tid & 3) * 4;
  const int64_t am = m0 + a_row;
  const bool a_valid = am < M;
  int an = 0, aqh = 0, aqw = 0;
  if (a_valid) {
    aqw = (int)(am % p.QW);
    int64_t t = am / p.QW;
    aqh = (int)(t % p.QH);
    an = (int)(t / p.QH);
  }
  const int a_h0 = aqh * p.i_mul + p.i_base_h, a_w0 = aqw * p.i_mul + p.i_base_w;
  const TIn* in = reinterpret_cast<const TIn*>(p.in) + (int64_t)an * p.i_sn;
  const int64_t ref_n = (int64_t)an * p.r_sn;

  const int b_col = tid & 63, b_k = (tid >> 6) * 4;
  const int bcn = n0 + b_col;
  const bool b_valid = bcn < p.CN;
  const float* wcol = p.w + (int64_t)bcn * p.w_scn;

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int kk = k0 + a_k + e;
      float v = 0.f;
      if (a_valid && kk < K) {
        const int tap = kk / p.CK, ck = kk - tap * p.CK;
        const int jh = tap / p.TW, jw = tap - jh * p.TW;
        const int ih = a_h0 + jh * p.i_tstep, iw = a_w0 + jw * p.i_tstep;
        if ((unsigned)ih < (unsigned)p.IH && (unsigned)iw < (unsigned)p.IW) {
          v = ld_as_float(in + (int64_t)ih * p.i_sh + (int64_t)iw * p.i_sw + (int64_t)ck * p.i_sc);
          if (p.ref)
            v *= act_grad_from_output(ld_rt(p.ref, p.ref_dtype, ref_n + (int64_t)ih * p.r_sh + (int64_t)iw * p.r_sw + (int64_t)ck * p.r_sc),
                                      p.g_act, p.g_slope);
        }
      }
      As[a_k + e][a_row] = v;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int kk = k0 + b_k + e;
      float v = 0.f;
      if (b_valid && kk < K) {
        const int tap = kk / p.CK, ck = kk - tap * p.CK;
        const int jh = tap / p.TW, jw = tap - jh * p.TW;
        const int kh = p.w_base_h + jh * p.w_step, kw = p.w_base_w + jw * p.w_step;
        v = __ldg(wcol + (int64_t)ck * p.w_sck + kh * p.ksize + kw);
      }
      Bs[b_k + e][b_col] = v;
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
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int qw = (int)(m % p.QW);
    const int64_t t = m / p.QW;
    const int qh = (int)(t % p.QH), n = (int)(t / p.QH);
    TOut* o = reinterpret_cast<TOut*>(p.out) + (int64_t)n * p.o_sn + (int64_t)(qh * p.o_mul + p.o_off_h) * p.o_sh +
              (int64_t)(qw * p.o_mul + p.o_off_w) * p.o_sw;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cn = n0 + tx * 4 + j;
      if (cn < p.CN) st_from_float(o + (int64_t)cn * p.o_sc, act_apply(acc[i][j], p.out_act, p.out_slope));
    }
  }
}

static void set_fuse(GatherParams& p, const ConvFuse& f) {
  p.ref = nullptr; p.g_act = B200GAN_ACT_NONE; p.g_slope = 0.f;
  if (f.g_ref) {
    p.ref = f.g_ref->ptr; p.ref_dtype = f.g_ref->dtype;
    p.r_sn = f.g_ref->sn; p.r_sh = f.g_ref->sh; p.r_sw = f.g_ref->sw; p.r_sc = f.g_ref->sc;
    p.g_act = f.g_act; p.g_slope = f.g_slope;
  }
  p.out_act = f.out_act; p.out_slope = f.out_slope;
}

static int launch_gather(const GatherParams& p, int in_dtype, int out_dtype, cudaStream_t st) {
  const int64_t M = (int64_t)p.N * p.QH * p.QW;
  if (M == 0 || p.CN == 0) return 0;
  dim3 grid((unsigned)((M + 63) / 64), (unsigned)((p.CN + 63) / 64));
  if (in_dtype == B200GAN_F32 && out_dtype == B200GAN_F32)
    gather_gemm_kernel<float, float><<<grid, 256, 0, st>>>(p);
  else if (in_dtype == B200GAN_BF16 && out_dtype == B200GAN_BF16)
    gather_gemm_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(p);
  else if (in_dtype == B200GAN_F32 && out_dtype == B200GAN_BF16)
    gather_gemm_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(p);
  else
    gather_gemm_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>(p);
  B200_LAUNCH_CHECK("gather_gemm_kernel");
  return 0;
}

// y = conv(x, w): conv geometry Co = y.c, Ci = x.c
int simt_conv_fprop(const b200gan_conv* cv, const b200gan_view* x, const float* w, const b200gan_view* y, const ConvFuse& fz,
                    cudaStream_t st) {
  GatherParams p{};
  set_fuse(p, fz);
  p.N = x->n; p.QH = y->h; p.QW = y->w;
  p.in = x->ptr; p.i_sn = x->sn; p.i_sh = x->sh; p.i_sw = x->sw; p.i_sc = x->sc; p.IH = x->h; p.IW = x->w;
  p.i_mul = cv->stride; p.i_base_h = p.i_base_w = -cv->pad; p.i_tstep = 1;
  p.TH = p.TW = cv->k; p.CK = x->c; p.CN = y->c;
  p.w = w; p.w_scn = (int64_t)x->c * cv->k * cv->k; p.w_sck = (int64_t)cv->k * cv->k;
  p.w_base_h = p.w_base_w = 0; p.w_step = 1; p.ksize = cv->k;
  p.out = y->ptr; p.o_sn = y->sn; p.o_sh = y->sh; p.o_sw = y->sw; p.o_sc = y->sc; p.o_mul = 1; p.o_off_h = p.o_off_w = 0;
  return launch_gather(p, x->dtype, y->dtype, st);
}

// dx = conv_dgrad(dy, w): conv geometry Co = dy.c, Ci = dx.c; one launch per output parity class
int simt_conv_dgrad(const b200gan_conv* cv, const b200gan_view* dy, const float* w, const b200gan_view* dx, const ConvFuse& fz,
                    cudaStream_t st) {
  const int s = cv->stride, k = cv->k, pad = cv->pad;
  for (int ph = 0; ph < s; ++ph)
    for (int pw = 0; pw < s; ++pw) {
      GatherParams p{};
      set_fuse(p, fz);
      p.N = dx->n; p.QH = (dx->h - ph + s - 1) / s; p.QW = (dx->w - pw + s - 1) / s;
      if (p.QH <= 0 || p.QW <= 0) continue;
      p.in = dy->ptr; p.i_sn = dy->sn; p.i_sh = dy->sh; p.i_sw = dy->sw; p.i_sc = dy->sc; p.IH = dy->h; p.IW = dy->w;
      const int rh = (ph + pad) % s, rw = (pw + pad) % s;
      p.i_mul = 1; p.i_base_h = (ph + pad) / s; p.i_base_w = (pw + pad) / s; p.i_tstep = -1;
      p.TH = rh < k ? (k - rh + s - 1) / s : 0; p.TW = rw < k ? (k - rw + s - 1) / s : 0;
      p.CK = dy->c; p.CN = dx->c;
      p.w = w; p.w_scn = (int64_t)k * k; p.w_sck = (int64_t)dx->c * k * k;
      p.w_base_h = rh; p.w_base_w = rw; p.w_step = s; p.ksize = k;
      p.out = dx->ptr; p.o_sn = dx->sn; p.o_sh = dx->sh; p.o_sw = dx->sw; p.o_sc = dx->sc; p.o_mul = s; p.o_off_h = ph; p.o_off_w = pw;
      int rc = launch_gather(p, dy->dtype, dx->dtype, st);
      if (rc) return rc;
    }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// weight gradient: dw[co,ci,kh,kw] += sum_{n,oh,ow} dy[n,oh,ow,co] * x[n,oh*s-p+kh,ow*s-p+kw,ci]
// GEMM rows = co, columns = (tap, ci), reduction = pixels, split over blockIdx.z, fp32 atomics.
// ---------------------------------------------------------------------------------------------------
struct WgradParams {
  int N, OH, OW, CO;
  const void* dy; int64_t d_sn, d_sh, d_sw, d_sc;
  const void* x;  int64_t x_sn, x_sh, x_sw, x_sc;
  int IH, IW, CI, k, stride, pad;
  float* dw;
  int64_t pix_per_split;
  // fused activation backward on the gradient operand: which = 0 none, 1 dy (coarse side), 2 x (fine side)
  const void* ref; int ref_dtype; int64_t r_sn, r_sh, r_sw, r_sc; int which, g_act; float g_slope;
};

template <typename TX, typename TD>
__global__ void __launch_bounds__(256) wgrad_kernel(WgradParams p) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int co0 = blockIdx.x * BM, col0 = blockIdx.y * BN;
  const int NC = p.k * p.k * p.CI;
  const int64_t P = (int64_t)p.N * p.OH * p.OW;
  const int64_t pbeg = (int64_t)blockIdx.z * p.pix_per_split;
  const int64_t pend = pbeg + p.pix_per_split < P ? pbeg + p.pix_per_split : P;

  const int lk = tid >> 4;            // pixel within the K-step handled by this thread's loads
  const int l4 = (tid & 15) * 4;      // 4 consecutive rows (A) / columns (B)
  // decode this thread's 4 B columns once
  int b_ci[4], b_kh[4], b_kw[4];
  bool b_ok[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int col = col0 + l4 + e;
    b_ok[e] = col < NC;
    const int tap = b_ok[e] ? col / p.CI : 0;
    b_ci[e] = b_ok[e] ? col - tap * p.CI : 0;
    b_kh[e] = tap / p.k;
    b_kw[e] = tap - b_kh[e] * p.k;
  }
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = pbeg; k0 < pend; k0 += BK) {
    const int64_t pix = k0 + lk;
    const bool pv = pix < pend;
    int n = 0, oh = 0, ow = 0;
    if (pv) {
      ow = (int)(pix % p.OW);
      const int64_t t = pix / p.OW;
      oh = (int)(t % p.OH);
      n = (int)(t / p.OH);
    }
    const TD* dyp = reinterpret_cast<const TD*>(p.dy) + (int64_t)n * p.d_sn + (int64_t)oh * p.d_sh + (int64_t)ow * p.d_sw;
    const TX* xp = reinterpret_cast<const TX*>(p.x) + (int64_t)n * p.x_sn;
    const int ih0 = oh * p.stride - p.pad, iw0 = ow * p.stride - p.pad;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int co = co0 + l4 + e;
      float av = (pv && co < p.CO) ? ld_as_float(dyp + (int64_t)co * p.d_sc) : 0.f;
      if (p.which == 1 && pv && co < p.CO)
        av *= act_grad_from_output(ld_rt(p.ref, p.ref_dtype, (int64_t)n * p.r_sn + (int64_t)oh * p.r_sh + (int64_t)ow * p.r_sw + (int64_t)co * p.r_sc),
                                   p.g_act, p.g_slope);
      As[lk][l4 + e] = av;
      float v = 0.f;
      if (pv && b_ok[e]) {
        const int ih = ih0 + b_kh[e], iw = iw0 + b_kw[e];
        if ((unsigned)ih < (unsigned)p.IH && (unsigned)iw < (unsigned)p.IW) {
          v = ld_as_float(xp + (int64_t)ih * p.x_sh + (int64_t)iw * p.x_sw + (int64_t)b_ci[e] * p.x_sc);
          if (p.which == 2)
            v *= act_grad_from_output(ld_rt(p.ref, p.ref_dtype, (int64_t)n * p.r_sn + (int64_t)ih * p.r_sh + (int64_t)iw * p.r_sw + (int64_t)b_ci[e] * p.r_sc),
                                      p.g_act, p.g_slope);
        }
      }
      Bs[lk][l4 + e] = v;
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
  const int kk2 = p.k * p.k;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= p.CO) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + tx * 4 + j;
      if (col >= NC) continue;
      const int tap = col / p.CI, ci = col - tap * p.CI;
      atomicAdd(p.dw + ((int64_t)co * p.CI + ci) * kk2 + tap, acc[i][j]);
    }
  }
}

// conv geometry: x (N,IH,IW,Ci) fine side, dy (N,OH,OW,Co) coarse side, dw (Co,Ci,k,k) accumulated
// fz.g_ref applies to dy when grad_is_coarse, to x otherwise (ConvTranspose2d: the gradient is the fine side)
int simt_conv_wgrad(const b200gan_conv* cv, const b200gan_view* x, const b200gan_view* dy, float* dw, const ConvFuse& fz, bool grad_is_coarse,
                    cudaStream_t st) {
  WgradParams p{};
  p.which = 0; p.ref = nullptr;
  if (fz.g_ref) {
    p.which = grad_is_coarse ? 1 : 2;
    p.ref = fz.g_ref->ptr; p.ref_dtype = fz.g_ref->dtype;
    p.r_sn = fz.g_ref->sn; p.r_sh = fz.g_ref->sh; p.r_sw = fz.g_ref->sw; p.r_sc = fz.g_ref->sc;
    p.g_act = fz.g_act; p.g_slope = fz.g_slope;
  }
  p.N = dy->n; p.OH = dy->h; p.OW = dy->w; p.CO = dy->c;
  p.dy = dy->ptr; p.d_sn = dy->sn; p.d_sh = dy->sh; p.d_sw = dy->sw; p.d_sc = dy->sc;
  p.x = x->ptr; p.x_sn = x->sn; p.x_sh = x->sh; p.x_sw = x->sw; p.x_sc = x->sc;
  p.IH = x->h; p.IW = x->w; p.CI = x->c; p.k = cv->k; p.stride = cv->stride; p.pad = cv->pad;
  p.dw = dw;
  const int64_t P = (int64_t)p.N * p.OH * p.OW;
  const int tiles_m = (p.CO + 63) / 64, tiles_n = (p.k * p.k * p.CI + 63) / 64;
  int64_t splits = (4 * kNumSMs + (int64_t)tiles_m * tiles_n - 1) / ((int64_t)tiles_m * tiles_n);
  const int64_t max_splits = (P + 127) / 128;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  p.pix_per_split = ((P + splits - 1) / splits + 15) / 16 * 16;
  splits = (P + p.pix_per_split - 1) / p.pix_per_split;
  dim3 grid(tiles_m, tiles_n, (unsigned)splits);
  if (x->dtype == B200GAN_F32 && dy->dtype == B200GAN_F32) wgrad_kernel<float, float><<<grid, 256, 0, st>>>(p);
  else if (x->dtype == B200GAN_BF16 && dy->dtype == B200GAN_BF16) wgrad_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(p);
  else if (x->dtype == B200GAN_F32 && dy->dtype == B200GAN_BF16) wgrad_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(p);
  else wgrad_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>(p);
  B200_LAUNCH_CHECK("wgrad_kernel");
  return 0;
}

}  // namespace b200gan
