// The pieces of the VGG16 perceptual loss of the conditional GAN (reference src/train_cgan.py:57-73,186: MSE between the feature maps of
// torchvision's vgg16.features[:4], [4:9], [9:16] of the fake and the real batch) that are not convolutions.
//
// Its stride-1 3x3 convolutions run on the library's stride-2 4x4 kernels (tcgen05 for 64..256 channels): output pixel (2i+a, 2j+b) of a
// Conv2d(3,1,1) reads input rows 2i+a-1 .. 2i+a+1, all inside the four rows 2i-1 .. 2i+2 a Conv2d(4,2,1) output pixel (i, j) reads, so
//     T[n, i, j, (a, b, co)] = sum_{ci,kh,kw} X[n, 2i-1+kh, 2j-1+kw, ci] * W4[(a, b, co), ci, kh, kw],   W4[(a,b,co),ci,kh,kw] = w3[co,ci,kh-a,kw-b]
// (zero where kh-a or kw-b leaves 0..2; zero padding identical in both forms) computes all four output parities as 4*Co channels of ONE k4 s2 p1
// convolution: 16/9 of the multiply-adds, on kernels that exist and run at tensor-core speed, input gradient included (b200gan_conv2d_dgrad with
// the same folded weight).  What remains is data movement, fused with the elementwise work here:
//   conv3x3_fold     the weight fold above (optionally zero-padding Ci: the 3-channel image is stored 32 channels wide for the tensor cores)
//   bias_relu_d2s    A[n, 2i+a, 2j+b, co] = relu(T[n, i, j, (a,b,co)] + bias[co])      (bias + ReLU + depth-to-space in one pass)
//   relu_bwd_s2d     dT[n, i, j, (a,b,co)] = dA[n, 2i+a, 2j+b, co] * (A[n, 2i+a, 2j+b, co] > 0)     (ReLU backward + space-to-depth)
//   maxpool2_fwd/bwd nn.MaxPool2d(2, 2); the backward routes each gradient to the FIRST maximum of its window in (row, column) order, as ATen does
// Dense NHWC tensors, f32 or bf16 (all operands of a call the same dtype); channels a multiple of 8 (bf16) / 4 (f32): 16-byte accesses throughout.
#include "common.cuh"

namespace b200gan {

namespace {

template <typename T> struct Vec;                       // one 16-byte access
template <> struct Vec<float> { static constexpr int N = 4; };
template <> struct Vec<__nv_bfloat16> { static constexpr int N = 8; };

template <typename T> __device__ __forceinline__ void load16(const T* p, float* out);
template <> __device__ __forceinline__ void load16<float>(const float* p, float* out) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
template <> __device__ __forceinline__ void load16<__nv_bfloat16>(const __nv_bfloat16* p, float* out) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h[k]); out[2 * k] = f.x; out[2 * k + 1] = f.y; }
}
template <typename T> __device__ __forceinline__ void store16(T* p, const float* in);
template <> __device__ __forceinline__ void store16<float>(float* p, const float* in) {
  *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[2], in[3]);
}
template <> __device__ __forceinline__ void store16<__nv_bfloat16>(__nv_bfloat16* p, const float* in) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(in[2 * k], in[2 * k + 1]);
  *reinterpret_cast<uint4*>(p) = v;
}

__global__ void conv3x3_fold_kernel(const float* __restrict__ w3, int co, int ci, int ci_pad, float* __restrict__ w4) {
  const int64_t total = 4ll * co * ci_pad * 16;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int kw = (int)(i & 3), kh = (int)((i >> 2) & 3);
    const int64_t q = i >> 4;
    const int c = (int)(q % ci_pad);
    const int oc = (int)(q / ci_pad);                    // (a, b, co)
    const int o = oc % co, ab = oc / co, a = ab >> 1, b = ab & 1;
    const int u = kh - a, v = kw - b;
    w4[i] = (c < ci && u >= 0 && u <= 2 && v >= 0 && v <= 2) ? w3[(((int64_t)o * ci + c) * 3 + u) * 3 + v] : 0.f;
  }
}

// one thread per 16-byte group of channels of one FINE pixel
template <typename T, bool BACKWARD>
__global__ void __launch_bounds__(256) d2s_kernel(const T* __restrict__ src, const T* __restrict__ ref, const float* __restrict__ bias, T* __restrict__ dst,
                                                  int n, int h, int w, int c) {
  constexpr int V = Vec<T>::N;
  const int groups = c / V;
  const int64_t total = (int64_t)n * h * w * groups;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int g = (int)(i % groups);
    int64_t q = i / groups;
    const int x = (int)(q % w); q /= w;
    const int y = (int)(q % h);
    const int img = (int)(q / h);
    const int64_t fine = (((int64_t)img * h + y) * w + x) * c + g * V;
    const int64_t coarse = ((((int64_t)img * (h >> 1) + (y >> 1)) * (w >> 1) + (x >> 1)) * 4 + ((y & 1) * 2 + (x & 1))) * c + g * V;
    float v[V];
    if (!BACKWARD) {                                     // A = relu(T + bias)
      load16<T>(src + coarse, v);
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] = fmaxf(v[k] + bias[g * V + k], 0.f);
      store16<T>(dst + fine, v);
    } else {                                             // dT = dA * (A > 0)
      float a[V];
      load16<T>(src + fine, v);
      load16<T>(ref + fine, a);
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] = a[k] > 0.f ? v[k] : 0.f;
      store16<T>(dst + coarse, v);
    }
  }
}

// one thread per 16-byte group of channels of one POOLED pixel
template <typename T, bool BACKWARD>
__global__ void __launch_bounds__(256) maxpool2_kernel(const T* __restrict__ a, const T* __restrict__ dp, T* __restrict__ out, int n, int ph, int pw, int c,
                                                       int add) {
  constexpr int V = Vec<T>::N;
  const int groups = c / V;
  const int64_t total = (int64_t)n * ph * pw * groups;
  const int64_t row = (int64_t)2 * pw * c;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int g = (int)(i % groups);
    int64_t q = i / groups;
    const int x = (int)(q % pw); q /= pw;
    const int y = (int)(q % ph);
    const int img = (int)(q / ph);
    const int64_t base = (((int64_t)img * 2 * ph + 2 * y) * 2 * pw + 2 * x) * c + g * V;
    const int64_t off[4] = {base, base + c, base + row, base + row + c};
    float v[4][V];
#pragma unroll
    for (int s = 0; s < 4; ++s) load16<T>(a + off[s], v[s]);
    const int64_t pooled = (((int64_t)img * ph + y) * pw + x) * c + g * V;
    if (!BACKWARD) {
      float m[V];
#pragma unroll
      for (int k = 0; k < V; ++k) m[k] = fmaxf(fmaxf(v[0][k], v[1][k]), fmaxf(v[2][k], v[3][k]));
      store16<T>(out + pooled, m);
    } else {
      float d[V], o[4][V];
      load16<T>(dp + pooled, d);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        int best = 0;                                    // first maximum in window order (strict > keeps the earliest)
#pragma unroll
        for (int s = 1; s < 4; ++s)
          if (v[s][k] > v[best][k]) best = s;
#pragma unroll
        for (int s = 0; s < 4; ++s) o[s][k] = s == best ? d[k] : 0.f;
      }
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (add) {
          float cur[V];
          load16<T>(out + off[s], cur);
#pragma unroll
          for (int k = 0; k < V; ++k) o[s][k] += cur[k];
        }
        store16<T>(out + off[s], o[s]);
      }
    }
  }
}

bool dense(const b200gan_view* v) {
  return v->sc == 1 && v->sw == v->c && v->sh == (int64_t)v->w * v->c && v->sn == (int64_t)v->h * v->w * v->c && (reinterpret_cast<uintptr_t>(v->ptr) & 15) == 0;
}

int grid_for(int64_t total) { return (int)((total + 255) / 256 < 16 * kNumSMs ? (total + 255) / 256 : 16 * kNumSMs); }

}  // namespace

int vgg_conv3x3_fold(const float* w3, int co, int ci, int ci_pad, float* w4, cudaStream_t st) {
  conv3x3_fold_kernel<<<grid_for(4ll * co * ci_pad * 16), 256, 0, st>>>(w3, co, ci, ci_pad, w4);
  B200_LAUNCH_CHECK("conv3x3_fold_kernel");
  return 0;
}

// fine: (N, H, W, C); coarse: (N, H/2, W/2, 4C)
static int check_pair(const b200gan_view* fine, const b200gan_view* coarse, const char* what) {
  B200_CHECK_ARG(dense(fine) && dense(coarse) && fine->dtype == coarse->dtype, "%s: operands must be dense NHWC tensors of one dtype", what);
  B200_CHECK_ARG(fine->h % 2 == 0 && fine->w % 2 == 0 && coarse->n == fine->n && coarse->h * 2 == fine->h && coarse->w * 2 == fine->w && coarse->c == 4 * fine->c,
                 "%s: extents must be (N,H,W,C) and (N,H/2,W/2,4C)", what);
  B200_CHECK_ARG(fine->c % (fine->dtype == B200GAN_F32 ? 4 : 8) == 0, "%s: channel count must be a multiple of %d", what, fine->dtype == B200GAN_F32 ? 4 : 8);
  return 0;
}

int vgg_bias_relu_d2s(const b200gan_view* t, const float* bias, const b200gan_view* a, cudaStream_t st) {
  int rc;
  if ((rc = check_pair(a, t, "bias_relu_d2s"))) return rc;
  const int64_t total = (int64_t)a->n * a->h * a->w * a->c;
  if (a->dtype == B200GAN_F32)
    d2s_kernel<float, false><<<grid_for(total / 4), 256, 0, st>>>((const float*)t->ptr, nullptr, bias, (float*)a->ptr, a->n, a->h, a->w, a->c);
  else
    d2s_kernel<__nv_bfloat16, false><<<grid_for(total / 8), 256, 0, st>>>((const __nv_bfloat16*)t->ptr, nullptr, bias, (__nv_bfloat16*)a->ptr, a->n, a->h, a->w, a->c);
  B200_LAUNCH_CHECK("d2s_kernel");
  return 0;
}

int vgg_relu_bwd_s2d(const b200gan_view* da, const b200gan_view* a, const b200gan_view* dt, cudaStream_t st) {
  int rc;
  if ((rc = check_pair(da, dt, "relu_bwd_s2d"))) return rc;
  B200_CHECK_ARG(dense(a) && a->dtype == da->dtype && a->n == da->n && a->h == da->h && a->w == da->w && a->c == da->c, "relu_bwd_s2d: the saved output must match the gradient");
  const int64_t total = (int64_t)da->n * da->h * da->w * da->c;
  if (da->dtype == B200GAN_F32)
    d2s_kernel<float, true><<<grid_for(total / 4), 256, 0, st>>>((const float*)da->ptr, (const float*)a->ptr, nullptr, (float*)dt->ptr, da->n, da->h, da->w, da->c);
  else
    d2s_kernel<__nv_bfloat16, true><<<grid_for(total / 8), 256, 0, st>>>((const __nv_bfloat16*)da->ptr, (const __nv_bfloat16*)a->ptr, nullptr, (__nv_bfloat16*)dt->ptr,
                                                                         da->n, da->h, da->w, da->c);
  B200_LAUNCH_CHECK("d2s_kernel");
  return 0;
}

static int check_pool(const b200gan_view* a, const b200gan_view* p, const char* what) {
  B200_CHECK_ARG(dense(a) && dense(p) && a->dtype == p->dtype, "%s: operands must be dense NHWC tensors of one dtype", what);
  B200_CHECK_ARG(a->h % 2 == 0 && a->w % 2 == 0 && p->n == a->n && p->h * 2 == a->h && p->w * 2 == a->w && p->c == a->c, "%s: extents must be (N,H,W,C) and (N,H/2,W/2,C)", what);
  B200_CHECK_ARG(a->c % (a->dtype == B200GAN_F32 ? 4 : 8) == 0, "%s: channel count must be a multiple of %d", what, a->dtype == B200GAN_F32 ? 4 : 8);
  return 0;
}

int vgg_maxpool2_fwd(const b200gan_view* a, const b200gan_view* p, cudaStream_t st) {
  int rc;
  if ((rc = check_pool(a, p, "maxpool2_fwd"))) return rc;
  const int64_t total = (int64_t)p->n * p->h * p->w * p->c;
  if (a->dtype == B200GAN_F32)
    maxpool2_kernel<float, false><<<grid_for(total / 4), 256, 0, st>>>((const float*)a->ptr, nullptr, (float*)p->ptr, p->n, p->h, p->w, p->c, 0);
  else
    maxpool2_kernel<__nv_bfloat16, false><<<grid_for(total / 8), 256, 0, st>>>((const __nv_bfloat16*)a->ptr, nullptr, (__nv_bfloat16*)p->ptr, p->n, p->h, p->w, p->c, 0);
  B200_LAUNCH_CHECK("maxpool2_kernel");
  return 0;
}

int vgg_maxpool2_bwd(const b200gan_view* a, const b200gan_view* dp, const b200gan_view* da, int add, cudaStream_t st) {
  int rc;
  if ((rc = check_pool(a, dp, "maxpool2_bwd"))) return rc;
  if ((rc = check_pool(da, dp, "maxpool2_bwd"))) return rc;
  const int64_t total = (int64_t)dp->n * dp->h * dp->w * dp->c;
  if (a->dtype == B200GAN_F32)
    maxpool2_kernel<float, true><<<grid_for(total / 4), 256, 0, st>>>((const float*)a->ptr, (const float*)dp->ptr, (float*)da->ptr, dp->n, dp->h, dp->w, dp->c, add);
  else
    maxpool2_kernel<__nv_bfloat16, true><<<grid_for(total / 8), 256, 0, st>>>((const __nv_bfloat16*)a->ptr, (const __nv_bfloat16*)dp->ptr, (__nv_bfloat16*)da->ptr, dp->n,
                                                                              dp->h, dp->w, dp->c, add);
  B200_LAUNCH_CHECK("maxpool2_kernel");
  return 0;
}

}  // namespace b200gan
