// Image-side k4 s2 p1 layers whose feature side is NOT 32 channels wide: the WGAN-GP critic's first layer Conv2d(nc -> 64) and the WGAN-GP
// generator's last layer ConvTranspose2d(64 -> nc) (reference src/wggan.py:52,41), forward, input gradient and weight gradient, with the same
// fused activation semantics as the 32-channel kernels of conv_thin_mma.cu (activation on the result; activation backward on the gathered /
// gradient operand taken from a saved output).  nc <= 4 makes these layers pure bandwidth + fp32-FMA work (K = 16 nc per output): a thread owns
// one pixel, the weights sit in shared memory, everything accumulates in fp32 on the CUDA cores.  They replace the generic SIMT implicit GEMM
// (64 x 64 x 16 tiles with per-element index arithmetic: 2.7 ms per call at batch 128) for these shapes.
//   fine   : the image side, (N, 2H, 2W, nc) through ANY strided view (the reference's NCHW fp32 tensors included), f32 or bf16
//   coarse : the feature side, (N, H, W, C) dense NHWC bf16, C % 16 == 0, C <= 128
//   w / dw : fp32 in conv geometry (C, nc, 4, 4)
#include "common.cuh"

namespace b200gan {

namespace {

constexpr int kMaxNc = 4;

struct EdgeArgs {
  View fine, fine_ref;                 // fine_ref.ptr == nullptr: none
  const __nv_bfloat16* coarse;
  const __nv_bfloat16* coarse_ref;     // nullptr: none
  __nv_bfloat16* coarse_out;
  const float* w;
  float* dw;
  int N, H, W, C, nc;                  // coarse extents
  int fine_act, coarse_act, out_act;
  float slope;
};

__device__ __forceinline__ void st_rt(void* base, int dtype, int64_t off, float x) {
  if (dtype == B200GAN_F32) reinterpret_cast<float*>(base)[off] = x;
  else reinterpret_cast<__nv_bfloat16*>(base)[off] = __float2bfloat16_rn(x);
}

// one image value (n, ih, iw, ci) with zero padding and the optional activation backward from the saved output
__device__ __forceinline__ float fine_at(const EdgeArgs& a, int n, int ih, int iw, int ci) {
  if ((unsigned)ih >= (unsigned)a.fine.h || (unsigned)iw >= (unsigned)a.fine.w) return 0.f;
  float v = ld_rt(a.fine.ptr, a.fine.dtype, (int64_t)n * a.fine.sn + (int64_t)ih * a.fine.sh + (int64_t)iw * a.fine.sw + (int64_t)ci * a.fine.sc);
  if (a.fine_ref.ptr)
    v *= act_grad_from_output(ld_rt(a.fine_ref.ptr, a.fine_ref.dtype, (int64_t)n * a.fine_ref.sn + (int64_t)ih * a.fine_ref.sh +
                                                                       (int64_t)iw * a.fine_ref.sw + (int64_t)ci * a.fine_ref.sc),
                              a.fine_act, a.slope);
  return v;
}

// ---- "down": coarse[n,oh,ow,:] = out_act( sum_{ci,kh,kw} fine[n, 2oh-1+kh, 2ow-1+kw, ci] w[:, ci, kh, kw] ) ------------------------------
template <int NC>
__global__ void __launch_bounds__(256) edge_down_kernel(const EdgeArgs a) {
  extern __shared__ float ws[];                                  // [(ci,kh,kw)][C]
  const int C = a.C, K = 16 * NC;
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) { const int co = i / K, k = i - co * K; ws[k * C + co] = a.w[i]; }
  __syncthreads();
  const int64_t total = (int64_t)a.N * a.H * a.W;
  for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (int64_t)gridDim.x * blockDim.x) {
    const int ow = (int)(pix % a.W);
    const int64_t t = pix / a.W;
    const int oh = (int)(t % a.H), n = (int)(t / a.H);
    float x[16 * NC];
#pragma unroll
    for (int ci = 0; ci < NC; ++ci)
#pragma unroll
      for (int kh = 0; kh < 4; ++kh)
#pragma unroll
        for (int kw = 0; kw < 4; ++kw) x[ci * 16 + kh * 4 + kw] = fine_at(a, n, 2 * oh - 1 + kh, 2 * ow - 1 + kw, ci);
    __nv_bfloat16* o = a.coarse_out + pix * C;
    for (int c0 = 0; c0 < C; c0 += 16) {
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float4* wr = reinterpret_cast<const float4*>(ws + k * C + c0);          // the same address in every lane: a broadcast
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wv = wr[q];
          acc[4 * q] = fmaf(x[k], wv.x, acc[4 * q]); acc[4 * q + 1] = fmaf(x[k], wv.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(x[k], wv.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(x[k], wv.w, acc[4 * q + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = act_apply(acc[j], a.out_act, a.slope);
      *reinterpret_cast<uint4*>(o + c0) = pack8(acc);
      *reinterpret_cast<uint4*>(o + c0 + 8) = pack8(acc + 8);
    }
  }
}

// ---- "up": fine[n,ih,iw,ci] = out_act( sum_{co} sum_{(kh,kw) of matching parity} g[n,(ih+1-kh)/2,(iw+1-kw)/2,co] w[co,ci,kh,kw] ),
//      g = coarse (* act'(coarse_ref)) -------------------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(256) edge_up_kernel(const EdgeArgs a) {
  extern __shared__ float ws[];                                  // [(ci,kh,kw)][C]
  const int C = a.C, K = 16 * NC;
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) { const int co = i / K, k = i - co * K; ws[k * C + co] = a.w[i]; }
  __syncthreads();
  const int FH = 2 * a.H, FW = 2 * a.W;
  const int64_t total = (int64_t)a.N * FH * FW;
  for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (int64_t)gridDim.x * blockDim.x) {
    const int iw = (int)(pix % FW);
    const int64_t t = pix / FW;
    const int ih = (int)(t % FH), n = (int)(t / FH);
    float acc[NC];
#pragma unroll
    for (int ci = 0; ci < NC; ++ci) acc[ci] = 0.f;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int kh = ((ih + 1) & 1) + 2 * jh, oh = (ih + 1 - kh) >> 1;
      if ((unsigned)oh >= (unsigned)a.H) continue;
#pragma unroll
      for (int jw = 0; jw < 2; ++jw) {
        const int kw = ((iw + 1) & 1) + 2 * jw, ow = (iw + 1 - kw) >> 1;
        if ((unsigned)ow >= (unsigned)a.W) continue;
        const int64_t off = (((int64_t)n * a.H + oh) * a.W + ow) * C;
        for (int c0 = 0; c0 < C; c0 += 8) {
          float g[8];
          unpack8(*reinterpret_cast<const uint4*>(a.coarse + off + c0), g);
          if (a.coarse_ref) {
            float r[8];
            unpack8(*reinterpret_cast<const uint4*>(a.coarse_ref + off + c0), r);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] *= act_grad_from_output(r[j], a.coarse_act, a.slope);
          }
#pragma unroll
          for (int ci = 0; ci < NC; ++ci) {
            const float* wr = ws + (ci * 16 + kh * 4 + kw) * C + c0;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[ci] = fmaf(g[j], wr[j], acc[ci]);
          }
        }
      }
    }
#pragma unroll
    for (int ci = 0; ci < NC; ++ci)
      st_rt(a.fine.ptr, a.fine.dtype, (int64_t)n * a.fine.sn + (int64_t)ih * a.fine.sh + (int64_t)iw * a.fine.sw + (int64_t)ci * a.fine.sc,
            act_apply(acc[ci], a.out_act, a.slope));
  }
}

// ---- weight gradient: dw[co,ci,kh,kw] += sum_{n,oh,ow} g[n,oh,ow,co] f[n, 2oh-1+kh, 2ow-1+kw, ci].  A warp walks segments of 32 consecutive output
//      pixels of one row: it stages the 4 x 66 image band under the segment in its own shared-memory slice (coalesced, bounds and the optional
//      activation backward applied once per image value), then every lane -- owner of CPL consecutive feature channels -- reads its gradient
//      values (one coalesced 2 CPL-byte load per pixel) and the 16 nc taps as shared-memory broadcasts.  Per-CTA partial sums go through shared
//      memory, one atomic per element and CTA at the end. ------------------------------------------------------------------------------------
constexpr int kBandCols = 66;

template <int NC, int CPL>
__global__ void __launch_bounds__(256) edge_wgrad_kernel(const EdgeArgs a) {
  extern __shared__ float sm[];                                  // [C][16*NC] partial sums of the CTA, then 8 x band[NC][4][66]
  const int C = a.C, K = 16 * NC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* red = sm;
  float* band = sm + C * K + warp * (NC * 4 * kBandCols);
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int c_lo = lane * CPL;
  const bool active = c_lo < C;
  float acc[CPL][16 * NC];
#pragma unroll
  for (int j = 0; j < CPL; ++j)
#pragma unroll
    for (int k = 0; k < K; ++k) acc[j][k] = 0.f;
  const int segs_w = (a.W + 31) / 32;
  const int64_t total = (int64_t)a.N * a.H * segs_w;
  const int64_t wid = (int64_t)blockIdx.x * 8 + warp, nw = (int64_t)gridDim.x * 8;
  for (int64_t sg = wid; sg < total; sg += nw) {
    const int sw = (int)(sg % segs_w);
    const int64_t t = sg / segs_w;
    const int oh = (int)(t % a.H), n = (int)(t / a.H);
    const int ow0 = sw * 32;
    __syncwarp();
    for (int i = lane; i < NC * 4 * kBandCols; i += 32) {
      const int col = i % kBandCols, r = (i / kBandCols) & 3, ci = i / (4 * kBandCols);
      band[i] = fine_at(a, n, 2 * oh - 1 + r, 2 * ow0 - 1 + col, ci);
    }
    __syncwarp();
    const int npx = a.W - ow0 < 32 ? a.W - ow0 : 32;
    const __nv_bfloat16* gp = a.coarse + (((int64_t)n * a.H + oh) * a.W + ow0) * C + c_lo;
    const __nv_bfloat16* rp = a.coarse_ref ? a.coarse_ref + (((int64_t)n * a.H + oh) * a.W + ow0) * C + c_lo : nullptr;
    for (int px = 0; px < npx; ++px) {
      float g[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        g[j] = 0.f;
        if (active) {
          g[j] = __bfloat162float(gp[(int64_t)px * C + j]);
          if (rp) g[j] *= act_grad_from_output(__bfloat162float(rp[(int64_t)px * C + j]), a.coarse_act, a.slope);
        }
      }
#pragma unroll
      for (int ci = 0; ci < NC; ++ci)
#pragma unroll
        for (int kh = 0; kh < 4; ++kh)
#pragma unroll
          for (int kw = 0; kw < 4; ++kw) {
            const float f = band[(ci * 4 + kh) * kBandCols + 2 * px + kw];          // the same address in every lane: a broadcast
#pragma unroll
            for (int j = 0; j < CPL; ++j) acc[j][ci * 16 + kh * 4 + kw] = fmaf(g[j], f, acc[j][ci * 16 + kh * 4 + kw]);
          }
    }
  }
  if (active) {
#pragma unroll
    for (int j = 0; j < CPL; ++j)
#pragma unroll
      for (int k = 0; k < K; ++k) atomicAdd(&red[(c_lo + j) * K + k], acc[j][k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * K; i += blockDim.x)
    if (red[i] != 0.f) atomicAdd(a.dw + i, red[i]);
}

bool dense_bf16(const b200gan_view* v) {
  return v->dtype == B200GAN_BF16 && v->sc == 1 && v->sw == v->c && v->sh == (int64_t)v->w * v->c && v->sn == (int64_t)v->h * v->w * v->c &&
         (reinterpret_cast<uintptr_t>(v->ptr) & 15) == 0;
}

// returns false when the problem is not of the edge shape
bool edge_setup(EdgeArgs* a, const b200gan_view* fine, const b200gan_view* fine_ref, int fine_act, const b200gan_view* coarse,
                const b200gan_view* coarse_ref, int coarse_act, float slope) {
  if (!dense_bf16(coarse) || coarse->c % 16 != 0 || coarse->c > 128 || fine->c < 1 || fine->c > kMaxNc) return false;
  if (fine->h != 2 * coarse->h || fine->w != 2 * coarse->w || fine->n != coarse->n) return false;
  if (coarse_ref && (!dense_bf16(coarse_ref) || coarse_ref->n != coarse->n || coarse_ref->h != coarse->h || coarse_ref->w != coarse->w || coarse_ref->c != coarse->c))
    return false;
  if (fine_ref && (fine_ref->n != fine->n || fine_ref->h != fine->h || fine_ref->w != fine->w || fine_ref->c != fine->c)) return false;
  a->fine = to_view(fine);
  if (fine_ref) a->fine_ref = to_view(fine_ref); else a->fine_ref.ptr = nullptr;
  a->coarse = reinterpret_cast<const __nv_bfloat16*>(coarse->ptr);
  a->coarse_out = reinterpret_cast<__nv_bfloat16*>(coarse->ptr);
  a->coarse_ref = coarse_ref ? reinterpret_cast<const __nv_bfloat16*>(coarse_ref->ptr) : nullptr;
  a->N = coarse->n; a->H = coarse->h; a->W = coarse->w; a->C = coarse->c; a->nc = fine->c;
  a->fine_act = fine_act; a->coarse_act = coarse_act; a->out_act = B200GAN_ACT_NONE; a->slope = slope;
  return true;
}

int grid_for(int64_t items, int per_block) {
  int64_t b = (items + per_block - 1) / per_block;
  const int64_t cap = 8 * (int64_t)kNumSMs;
  if (b > cap) b = cap;
  return b < 1 ? 1 : (int)b;
}

}  // namespace

#define EDGE_NC(KERNEL, GRID, SMEM, A)                                                  \
  do {                                                                                  \
    switch ((A).nc) {                                                                   \
      case 1: KERNEL<1><<<GRID, 256, SMEM, st>>>(A); break;                             \
      case 2: KERNEL<2><<<GRID, 256, SMEM, st>>>(A); break;                             \
      case 3: KERNEL<3><<<GRID, 256, SMEM, st>>>(A); break;                             \
      default: KERNEL<4><<<GRID, 256, SMEM, st>>>(A); break;                            \
    }                                                                                   \
  } while (0)

// fine (optionally * act'(fine_ref)) -> coarse = out_act(conv)
int edge_down(const b200gan_view* fine, const b200gan_view* fine_ref, int fine_act, const float* w, const b200gan_view* coarse, int out_act, float slope,
              cudaStream_t st) {
  EdgeArgs a{};
  if (!edge_setup(&a, fine, fine_ref, fine_act, coarse, nullptr, B200GAN_ACT_NONE, slope)) return 1;
  a.w = w; a.out_act = out_act;
  const size_t smem = (size_t)16 * a.nc * a.C * sizeof(float);
  if (smem > 48 * 1024) return 1;
  const int grid = grid_for((int64_t)a.N * a.H * a.W, 256);
  EDGE_NC(edge_down_kernel, grid, smem, a);
  B200_LAUNCH_CHECK("edge_down_kernel");
  return 0;
}

// coarse (optionally * act'(coarse_ref)) -> fine = out_act(transposed conv)
int edge_up(const b200gan_view* coarse, const b200gan_view* coarse_ref, int coarse_act, float slope, const float* w, const b200gan_view* fine, int out_act,
            cudaStream_t st) {
  EdgeArgs a{};
  if (!edge_setup(&a, fine, nullptr, B200GAN_ACT_NONE, coarse, coarse_ref, coarse_act, slope)) return 1;
  a.w = w; a.out_act = out_act;
  const size_t smem = (size_t)16 * a.nc * a.C * sizeof(float);
  if (smem > 48 * 1024) return 1;
  const int grid = grid_for((int64_t)a.N * 4 * a.H * a.W, 256);
  EDGE_NC(edge_up_kernel, grid, smem, a);
  B200_LAUNCH_CHECK("edge_up_kernel");
  return 0;
}

// dw (C, nc, 4, 4) += coarse^T x im2col(fine); either operand may carry the fused activation backward
int edge_wgrad(const b200gan_view* fine, const b200gan_view* fine_ref, int fine_act, const b200gan_view* coarse, const b200gan_view* coarse_ref, int coarse_act,
               float slope, float* dw, cudaStream_t st) {
  EdgeArgs a{};
  if (!edge_setup(&a, fine, fine_ref, fine_act, coarse, coarse_ref, coarse_act, slope)) return 1;
  a.dw = dw;
  const size_t smem = ((size_t)16 * a.nc * a.C + (size_t)8 * a.nc * 4 * kBandCols) * sizeof(float);
  const int64_t segs = (int64_t)a.N * a.H * ((a.W + 31) / 32);
  const int grid = grid_for(segs, 8 * 4);                               // a few segments per warp at least
  const int cpl = (a.C + 31) / 32;
#define EDGE_WG(NC, CPL)                                                                    \
  do {                                                                                      \
    B200_CUDA((ensure_dynamic_smem<edge_wgrad_kernel<NC, CPL>>((int)smem)));                \
    edge_wgrad_kernel<NC, CPL><<<grid, 256, smem, st>>>(a);                                 \
  } while (0)
  if (cpl == 1) { if (a.nc == 1) EDGE_WG(1, 1); else if (a.nc == 2) EDGE_WG(2, 1); else if (a.nc == 3) EDGE_WG(3, 1); else EDGE_WG(4, 1); }
  else if (cpl == 2) { if (a.nc == 1) EDGE_WG(1, 2); else if (a.nc == 2) EDGE_WG(2, 2); else if (a.nc == 3) EDGE_WG(3, 2); else EDGE_WG(4, 2); }
  else if (cpl <= 4 && a.nc <= 2) { if (a.nc == 1) EDGE_WG(1, 4); else EDGE_WG(2, 4); }
  else return 1;                                                          // register budget: 4 channels x 48 taps per lane is too much
#undef EDGE_WG
  B200_LAUNCH_CHECK("edge_wgrad_kernel");
  return 0;
}

}  // namespace b200gan
