// Kernels specific to the conditional GAN (reference src/cgan.py) -- the pieces around its convolutions:
//   * label conditioning of the Generator (cgan.py:22,55-56): x = z + label_emb[labels], written with a trailing column of ones so that the
//     Linear layer's bias (cgan.py:24,57) rides as one more input feature of the latent GEMM (and its gradient falls out of that GEMM's
//     weight gradient);
//   * nearest Upsample(2) + Conv2d(3, 1, 1) (cgan.py:28-29,33-34,38-39,43-44,48-49) as ONE stride-2 transposed convolution: an output pixel
//     2i+a reads upsampled pixels 2i+a-1 .. 2i+a+1, i.e. source pixels {i-1, i, i} (a = 0) or {i, i, i+1} (a = 1), so per dimension the three
//     taps fold into the two taps a ConvTranspose2d(4, 2, 1) applies to the same sources (zero padding of the upsampled border = source index
//     out of range in both forms):   W4[kh] = sum_a A[kh][a] w3[a],  A = [[0,0,1],[0,1,1],[1,1,0],[1,0,0]]  (kh = 0..3), both dimensions,
//     and (Cout,Cin) -> (Cin,Cout).  The folded weight feeds the ConvTranspose2d kernels (tcgen05 for the wide layers); the weight
//     gradient folds back with the adjoint map.  4/9 of the multiply-adds of the upsample-then-convolve form, no upsampled tensor in HBM;
//   * the projection term of the Discriminator (cgan.py:103): out[n] += <label_emb[labels[n]], features[n] in (C,H,W) order>, and its backward.
// All fp32 arithmetic; activations f32 or bf16 through strided views.  Batch-order loops instead of atomics: results are run-to-run identical.
#include "common.cuh"

namespace b200gan {

namespace {

__device__ __forceinline__ void st_any(void* base, int dtype, int64_t off, float x) {
  if (dtype == B200GAN_F32) reinterpret_cast<float*>(base)[off] = x;
  else reinterpret_cast<__nv_bfloat16*>(base)[off] = __float2bfloat16_rn(x);
}

__global__ void embed_add_kernel(const float* __restrict__ table, const int64_t* __restrict__ labels, const float* __restrict__ z, int batch,
                                 int dim, int tail, float* __restrict__ out) {
  const int row = dim + tail;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)batch * row; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / row), d = (int)(i - (int64_t)n * row);
    out[i] = d < dim ? table[labels[n] * dim + d] + (z ? z[(int64_t)n * dim + d] : 0.f) : 1.f;
  }
}

// dtable[cls][d] += sum over the samples of class cls of dx[n][d]   (rows of dx are `stride` floats apart)
__global__ void embed_bwd_kernel(const float* __restrict__ dx, const int64_t* __restrict__ labels, int batch, int dim, int stride, int classes,
                                 float* __restrict__ dtable) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < classes * dim; i += gridDim.x * blockDim.x) {
    const int cls = i / dim, d = i - cls * dim;
    float s = 0.f;
    for (int n = 0; n < batch; ++n)
      if (labels[n] == cls) s += dx[(int64_t)n * stride + d];
    dtable[i] += s;
  }
}

__device__ __forceinline__ float fold_coeff(int k4, int a) {      // A[k4][a]
  return (k4 == 0 && a == 2) || (k4 == 1 && a >= 1) || (k4 == 2 && a <= 1) || (k4 == 3 && a == 0) ? 1.f : 0.f;
}

// w4[(ci, co, kh, kw)] = sum_{a,b} A[kh][a] A[kw][b] w3[(co, ci, a, b)]
__global__ void upconv3_fold_kernel(const float* __restrict__ w3, int co, int ci, float* __restrict__ w4) {
  const int total = ci * co * 16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kw = i & 3, kh = (i >> 2) & 3, o = (i >> 4) % co, c = (i >> 4) / co;
    const float* src = w3 + ((int64_t)o * ci + c) * 9;
    float s = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) s += fold_coeff(kh, a) * fold_coeff(kw, b) * src[a * 3 + b];
    w4[i] = s;
  }
}

// dw3[(co, ci, a, b)] += sum_{kh,kw} A[kh][a] A[kw][b] dw4[(ci, co, kh, kw)]
__global__ void upconv3_unfold_kernel(const float* __restrict__ dw4, int co, int ci, float* __restrict__ dw3) {
  const int total = co * ci * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i % 3, a = (i / 3) % 3, c = (i / 9) % ci, o = (i / 9) / ci;
    const float* src = dw4 + ((int64_t)c * co + o) * 16;
    float s = 0.f;
#pragma unroll
    for (int kh = 0; kh < 4; ++kh)
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) s += fold_coeff(kh, a) * fold_coeff(kw, b) * src[kh * 4 + kw];
    dw3[i] += s;
  }
}

__device__ __forceinline__ int64_t feat_off(const View& v, int n, int j) {     // j = (c, h, w) flattened, the order of x.view(N, -1) on NCHW
  const int hw = v.h * v.w, c = j / hw, r = j - c * hw, h = r / v.w, w = r - h * v.w;
  return (int64_t)n * v.sn + (int64_t)h * v.sh + (int64_t)w * v.sw + (int64_t)c * v.sc;
}

// one CTA per sample: out[n] += <table[labels[n]], x[n]>
__global__ void __launch_bounds__(256) class_proj_fwd_kernel(View x, const float* __restrict__ table, const int64_t* __restrict__ labels,
                                                            float* __restrict__ out) {
  __shared__ float part[256];
  const int n = blockIdx.x, F = x.h * x.w * x.c;
  const float* row = table + labels[n] * F;
  float s = 0.f;
  for (int j = threadIdx.x; j < F; j += 256) s = fmaf(row[j], ld_rt(x.ptr, x.dtype, feat_off(x, n, j)), s);
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] += part[0];
}

// dx[n] += dout[n] * table[labels[n]]
__global__ void class_proj_bwd_x_kernel(View dx, const float* __restrict__ table, const int64_t* __restrict__ labels, const float* __restrict__ dout) {
  const int n = blockIdx.y, F = dx.h * dx.w * dx.c;
  const float g = dout[n];
  const float* row = table + labels[n] * F;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < F; j += gridDim.x * blockDim.x) {
    const int64_t off = feat_off(dx, n, j);
    st_any(dx.ptr, dx.dtype, off, fmaf(g, row[j], ld_rt(dx.ptr, dx.dtype, off)));
  }
}

// dtable[cls][j] += sum over the samples of class cls of dout[n] * x[n][j]
__global__ void class_proj_bwd_table_kernel(View x, const int64_t* __restrict__ labels, const float* __restrict__ dout, int classes,
                                            float* __restrict__ dtable) {
  const int F = x.h * x.w * x.c;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < classes * F; i += gridDim.x * blockDim.x) {
    const int cls = i / F, j = i - cls * F;
    float s = 0.f;
    for (int n = 0; n < x.n; ++n)
      if (labels[n] == cls) s = fmaf(dout[n], ld_rt(x.ptr, x.dtype, feat_off(x, n, j)), s);
    dtable[i] += s;
  }
}

// BCEWithLogitsLoss(mean) against per-sample targets, its gradient and the mean probability, one CTA:
//   out2[0] = mean_i max(x,0) - x t + log1p(exp(-|x|)),  out2[1] = mean_i sigmoid(x),  dlogit[i] = grad_scale (sigmoid(x) - t) / B
__global__ void __launch_bounds__(256) bce_logits_kernel(const float* __restrict__ x, const float* __restrict__ t, int batch, float grad_scale,
                                                        float* __restrict__ out2, float* __restrict__ dlogit) {
  __shared__ double part[2][256];
  double loss = 0.0, prob = 0.0;
  for (int i = threadIdx.x; i < batch; i += 256) {
    const float xi = x[i], ti = t[i];
    const float p = 1.f / (1.f + expf(-xi));
    loss += (double)(fmaxf(xi, 0.f) - xi * ti + log1pf(expf(-fabsf(xi))));
    prob += (double)p;
    if (dlogit) dlogit[i] = grad_scale * (p - ti) / (float)batch;
  }
  part[0][threadIdx.x] = loss; part[1][threadIdx.x] = prob;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { part[0][threadIdx.x] += part[0][threadIdx.x + o]; part[1][threadIdx.x] += part[1][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out2[0] = (float)(part[0][0] / batch); out2[1] = (float)(part[1][0] / batch); }
}

// Feature matching on one pair of equally shaped tensors (train_cgan.py:75-76): sum += sum (r - f)^2 (fp64), dfake = coeff (r - f)
// (coeff = -2 weight multiplicity / numel; overwrites dfake or adds to it).  Element order follows the fake view (dense NHWC in practice).
__global__ void __launch_bounds__(256) fm_pair_kernel(View r, View f, View d, float coeff, int add, double* __restrict__ sum) {
  __shared__ double part[256];
  const int64_t total = (int64_t)f.n * f.h * f.w * f.c;
  double s = 0.0;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int c = (int)(i % f.c);
    int64_t q = i / f.c;
    const int w = (int)(q % f.w); q /= f.w;
    const int h = (int)(q % f.h);
    const int n = (int)(q / f.h);
    const float rv = ld_rt(r.ptr, r.dtype, (int64_t)n * r.sn + (int64_t)h * r.sh + (int64_t)w * r.sw + (int64_t)c * r.sc);
    const float fv = ld_rt(f.ptr, f.dtype, (int64_t)n * f.sn + (int64_t)h * f.sh + (int64_t)w * f.sw + (int64_t)c * f.sc);
    const float diff = rv - fv;
    s += (double)diff * diff;
    if (d.ptr) {
      const int64_t off = (int64_t)n * d.sn + (int64_t)h * d.sh + (int64_t)w * d.sw + (int64_t)c * d.sc;
      st_any(d.ptr, d.dtype, off, add ? fmaf(coeff, diff, ld_rt(d.ptr, d.dtype, off)) : coeff * diff);
    }
  }
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(sum, part[0]);
}

// the same for three dense bf16 tensors of one layout (every pair but the padded one): 16-byte accesses, eight elements per thread and trip
__global__ void __launch_bounds__(256) fm_pair_dense_kernel(const uint4* __restrict__ r, const uint4* __restrict__ f, uint4* __restrict__ d, int64_t vecs,
                                                           float coeff, int add, double* __restrict__ sum) {
  __shared__ double part[256];
  float s = 0.f;
  double acc = 0.0;
  int trips = 0;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < vecs; i += (int64_t)gridDim.x * 256) {
    const uint4 rv = r[i], fv = f[i];
    uint4 dv = (d && add) ? d[i] : make_uint4(0u, 0u, 0u, 0u);
    const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&rv);
    const __nv_bfloat162* fp = reinterpret_cast<const __nv_bfloat162*>(&fv);
    __nv_bfloat162* dp = reinterpret_cast<__nv_bfloat162*>(&dv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 a = __bfloat1622float2(rp[k]), b = __bfloat1622float2(fp[k]);
      const float d0 = a.x - b.x, d1 = a.y - b.y;
      s = fmaf(d0, d0, fmaf(d1, d1, s));
      if (d) {
        const float2 o = __bfloat1622float2(dp[k]);
        dp[k] = __floats2bfloat162_rn(fmaf(coeff, d0, o.x), fmaf(coeff, d1, o.y));
      }
    }
    if (d) d[i] = dv;
    if (++trips == 16) { acc += (double)s; s = 0.f; trips = 0; }          // bound the fp32 partial sums
  }
  part[threadIdx.x] = acc + (double)s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(sum, part[0]);
}

bool dense_bf16_same(const b200gan_view* a, const b200gan_view* b) {
  return a->dtype == B200GAN_BF16 && b->dtype == B200GAN_BF16 && a->sc == 1 && a->sw == a->c && a->sh == (int64_t)a->w * a->c &&
         a->sn == (int64_t)a->h * a->w * a->c && b->sc == 1 && b->sw == a->sw && b->sh == a->sh && b->sn == a->sn &&
         ((reinterpret_cast<uintptr_t>(a->ptr) | reinterpret_cast<uintptr_t>(b->ptr)) & 15) == 0;
}

// dst[r][c] += src[r * srs + c * scs]   (src float or double): bias gradients out of fp64 channel sums, the Linear's gradient out of the
// latent GEMM's transposed layout
__global__ void accumulate_2d_kernel(float* __restrict__ dst, const void* __restrict__ src, int f64, int rows, int cols, int64_t srs, int64_t scs) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)rows * cols; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols, j = r * srs + c * scs;
    dst[i] += f64 ? (float)reinterpret_cast<const double*>(src)[j] : reinterpret_cast<const float*>(src)[j];
  }
}

int grid_for(int64_t total) { return (int)((total + 255) / 256 < 8 * kNumSMs ? (total + 255) / 256 : 8 * kNumSMs); }

}  // namespace

int cgan_embed_add(const float* table, const int64_t* labels, const float* z, int batch, int dim, int tail, float* out, cudaStream_t st) {
  embed_add_kernel<<<grid_for((int64_t)batch * (dim + tail)), 256, 0, st>>>(table, labels, z, batch, dim, tail, out);
  B200_LAUNCH_CHECK("embed_add_kernel");
  return 0;
}

int cgan_embed_bwd(const float* dx, const int64_t* labels, int batch, int dim, int stride, int classes, float* dtable, cudaStream_t st) {
  embed_bwd_kernel<<<grid_for((int64_t)classes * dim), 256, 0, st>>>(dx, labels, batch, dim, stride, classes, dtable);
  B200_LAUNCH_CHECK("embed_bwd_kernel");
  return 0;
}

int cgan_upconv3_fold(const float* w3, int co, int ci, float* w4, cudaStream_t st) {
  upconv3_fold_kernel<<<grid_for((int64_t)co * ci * 16), 256, 0, st>>>(w3, co, ci, w4);
  B200_LAUNCH_CHECK("upconv3_fold_kernel");
  return 0;
}

int cgan_upconv3_unfold(const float* dw4, int co, int ci, float* dw3, cudaStream_t st) {
  upconv3_unfold_kernel<<<grid_for((int64_t)co * ci * 9), 256, 0, st>>>(dw4, co, ci, dw3);
  B200_LAUNCH_CHECK("upconv3_unfold_kernel");
  return 0;
}

int cgan_class_proj_fwd(const b200gan_view* x, const float* table, const int64_t* labels, float* out, cudaStream_t st) {
  class_proj_fwd_kernel<<<x->n, 256, 0, st>>>(to_view(x), table, labels, out);
  B200_LAUNCH_CHECK("class_proj_fwd_kernel");
  return 0;
}

int cgan_class_proj_bwd(const b200gan_view* x, const float* table, const int64_t* labels, const float* dout, const b200gan_view* dx, int classes,
                        float* dtable, cudaStream_t st) {
  const int F = x->h * x->w * x->c;
  if (dx) {
    class_proj_bwd_x_kernel<<<dim3((F + 255) / 256, x->n), 256, 0, st>>>(to_view(dx), table, labels, dout);
    B200_LAUNCH_CHECK("class_proj_bwd_x_kernel");
  }
  if (dtable) {
    class_proj_bwd_table_kernel<<<grid_for((int64_t)classes * F), 256, 0, st>>>(to_view(x), labels, dout, classes, dtable);
    B200_LAUNCH_CHECK("class_proj_bwd_table_kernel");
  }
  return 0;
}

int cgan_bce_logits(const float* x, const float* t, int batch, float grad_scale, float* out2, float* dlogit, cudaStream_t st) {
  bce_logits_kernel<<<1, 256, 0, st>>>(x, t, batch, grad_scale, out2, dlogit);
  B200_LAUNCH_CHECK("bce_logits_kernel");
  return 0;
}

int cgan_fm_pair(const b200gan_view* r, const b200gan_view* f, const b200gan_view* d, float coeff, int add, double* sum, cudaStream_t st) {
  View dv;
  if (d) dv = to_view(d); else dv.ptr = nullptr;
  const int64_t total = (int64_t)f->n * f->h * f->w * f->c;
  if (total % 8 == 0 && dense_bf16_same(f, r) && (!d || dense_bf16_same(f, d))) {
    fm_pair_dense_kernel<<<grid_for(total / 8), 256, 0, st>>>(reinterpret_cast<const uint4*>(r->ptr), reinterpret_cast<const uint4*>(f->ptr),
                                                             d ? reinterpret_cast<uint4*>(d->ptr) : nullptr, total / 8, coeff, add, sum);
    B200_LAUNCH_CHECK("fm_pair_dense_kernel");
    return 0;
  }
  fm_pair_kernel<<<grid_for(total), 256, 0, st>>>(to_view(r), to_view(f), dv, coeff, add, sum);
  B200_LAUNCH_CHECK("fm_pair_kernel");
  return 0;
}

int cgan_accumulate_2d(float* dst, const void* src, int f64, int rows, int cols, int64_t srs, int64_t scs, cudaStream_t st) {
  accumulate_2d_kernel<<<grid_for((int64_t)rows * cols), 256, 0, st>>>(dst, src, f64, rows, cols, srs, scs);
  B200_LAUNCH_CHECK("accumulate_2d_kernel");
  return 0;
}

}  // namespace b200gan
