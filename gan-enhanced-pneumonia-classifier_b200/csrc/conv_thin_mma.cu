// The image-side layers of the DCGAN: D0 = Conv2d(nc->32, k4 s2 p1) + LeakyReLU (dcgan.py:65-66) and
// G5 = ConvTranspose2d(32->nc, k4 s2 p1) + Tanh (dcgan.py:46-47), nc in {1,3}, forward, input gradient and weight
// gradient.  They move the largest activation of each network (N x 112 x 112 x 32) against a 1- or 3-channel image:
// ~14 FLOP per byte, far below the B200 ridge, so the design goal is one pass over HBM per operand:
//   * a CTA stages a band of image rows (any dtype / strides: the reference's NCHW fp32 tensors are read in place)
//     and/or of 32-channel rows in shared memory once, zero padding included, and every tap re-reads shared memory;
//   * the arithmetic runs on warp-level tensor-core MMAs (mma.sync m16n8k16, bf16 x bf16 -> fp32): with K = 16 taps
//     (down / wgrad) or N = 4 parity classes (up) these GEMMs are far too thin for a 128-row tcgen05 tile, and the
//     legacy MMA path is already >10x faster than the HBM floor here;
//   * the activation around the convolution is fused: LeakyReLU / Tanh on the way out, and the activation BACKWARD
//     (gradient x act'(saved output)) while the gradient operand is staged, so no elementwise pass touches these tensors.
// Conv geometry names: fine = (N, 2H, 2W, nc) image side, coarse = (N, H, W, 32) dense bf16 NHWC, w = (32, nc, 4, 4) fp32.
#include "common.cuh"
#include "ptx.cuh"

namespace b200gan {

namespace {

struct ThinArgs {
  View fine, fine_ref;                 // fine_ref.ptr == nullptr: no transform on the fine operand
  const __nv_bfloat16* coarse;         // gathered coarse operand (up, wgrad)
  const __nv_bfloat16* coarse_ref;     // nullptr: no transform on the coarse operand
  __nv_bfloat16* coarse_out;           // result of down
  const float* w;
  float* dw;
  int N, H, W;                         // coarse extents
  int R, tiles_per_img, num_tiles;     // coarse rows per tile
  int fine_act, coarse_act, out_act;   // b200gan_act: derivative applied to the staged operand / activation of the result
  float slope;
  int fine_vec, ref_vec;               // 4-wide vector loads along W are legal for fine / fine_ref
  int RT;                              // coarse rows per tile of the TMA-staged kernels
  int cpitch;                          // elements between consecutive coarse pixels: 32 (dense) or the width of the tensor a 32-channel slice is cut
                                       // from (only the TMA-staged kernels take slices: their tensor map carries the strides)
  // down only: BatchNorm fusions on the 32-channel result (b200gan_fuse): epi 1 = statistics of the result (bn_sums),
  // epi 2 = the result is the gradient w.r.t. act(BN(prev_y)): dz = dx * act'(scale*prev_y + shift), sums of dz and dz*xhat
  int epi;
  double* sums;
  const __nv_bfloat16* prev_y;
  const float *prev_scale, *prev_shift, *prev_mean, *prev_invstd;
  float prev_neg;
};

__device__ __forceinline__ float tanh_fast(float x) {       // MUFU.TANH: 2^-11 relative error, below the bf16 storage rounding
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// byte offset of 16-byte chunk `chunk` of pixel `pix` inside a TMA tile written with CU_TENSOR_MAP_SWIZZLE_64B (64-byte pixel
// rows, tile base 512-byte aligned): address bits [4:5] are XORed with bits [7:8]
__device__ __forceinline__ uint32_t sw64(int pix, int chunk) { return (uint32_t)pix * 64u + (uint32_t)((chunk ^ ((pix >> 1) & 3)) << 4); }

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&b);
}

// four elements consecutive along W of a strided view, runtime dtype
__device__ __forceinline__ void load4_rt(const View& v, int64_t off, int vec, float (&o)[4]) {
  if (v.dtype == B200GAN_F32) {
    const float* p = reinterpret_cast<const float*>(v.ptr) + off;
    if (vec) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p));
      o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = __ldg(p + e * v.sw);
    }
  } else {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(v.ptr) + off;
    if (vec) {
      const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
      o[0] = __uint_as_float(t.x << 16); o[1] = __uint_as_float(t.x & 0xffff0000u);
      o[2] = __uint_as_float(t.y << 16); o[3] = __uint_as_float(t.y & 0xffff0000u);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = __bfloat162float(p[e * v.sw]);
    }
  }
}

// ---- row-per-warp staging (same shared-memory layout as stage_fine below) --------------------------------------------------------------------------
// ncu on the per-element loop of stage_fine (round 2, nc = 3): 154 M warp instructions per launch of the first Discriminator layer, ~2/3 of them the
// staging loop's index arithmetic (two divisions per four pixels) and its three odd-aligned 2+4+2-byte shared-memory stores; issue slots 57 % busy at 24 %
// occupancy.  Here a warp owns a whole image row: a lane loads EIGHT consecutive pixels with 16-byte accesses, packs them to bf16 pairs, takes its left
// neighbour's last pixel by one shuffle (the layout puts pixel iw at element iw+1, so the aligned 4-byte words hold the pairs (iw odd, iw+1)) and writes
// four aligned words.  No division per element, one shuffle and four 4-byte stores per eight pixels.  Two layouts qualify: channel-planar rows with unit
// W stride (the reference's NCHW tensors, f32 or bf16; any tensor at nc = 1) and the dense interleaved 3-channel bf16 image the fused trainer keeps
// between the networks; a reference tensor for the fused activation backward must have the same layout.  IW % 8 == 0, IW <= 256.
__device__ __forceinline__ void row_words_store(uint32_t* Srow32, int lane, int groups, uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3) {
  // q_i = bf16 pair (pixel 2i low, pixel 2i+1 high) of this lane's eight pixels; lanes >= groups hold zeros
  uint32_t prev = __shfl_up_sync(0xffffffffu, q3, 1);
  if (lane == 0) prev = 0u;                                        // iw = -1: the left zero padding
  if (lane < groups) {
    uint32_t* d = Srow32 + 4 * lane;
    d[0] = __funnelshift_l(prev, q0, 16);                          // (iw = 8 lane - 1, 8 lane)
    d[1] = __funnelshift_l(q0, q1, 16);
    d[2] = __funnelshift_l(q1, q2, 16);
    d[3] = __funnelshift_l(q2, q3, 16);
    if (lane == groups - 1) d[4] = q3 >> 16;                       // (iw = IW - 1, the right zero padding)
  }
}

__device__ __forceinline__ void load8_planar(const View& v, int64_t off, float (&o)[8]) {       // eight pixels consecutive along W, unit stride
  if (v.dtype == B200GAN_F32) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(v.ptr) + off));
    const float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(v.ptr) + off + 4));
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  } else {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(v.ptr) + off));
    o[0] = __uint_as_float(t.x << 16); o[1] = __uint_as_float(t.x & 0xffff0000u); o[2] = __uint_as_float(t.y << 16); o[3] = __uint_as_float(t.y & 0xffff0000u);
    o[4] = __uint_as_float(t.z << 16); o[5] = __uint_as_float(t.z & 0xffff0000u); o[6] = __uint_as_float(t.w << 16); o[7] = __uint_as_float(t.w & 0xffff0000u);
  }
}

__device__ __forceinline__ bool rows_planar_ok(const View& v, int IW) {
  const int per = v.dtype == B200GAN_F32 ? 4 : 8;                  // elements per 16 bytes
  return v.sw == 1 && v.sn % per == 0 && v.sh % per == 0 && (v.c == 1 || v.sc % per == 0) &&
         (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0 && IW % 8 == 0 && IW <= 256;
}
__device__ __forceinline__ bool rows_rgb_ok(const View& v, int IW) {                           // dense interleaved (N,H,W,3) bf16
  return v.dtype == B200GAN_BF16 && v.c == 3 && v.sc == 1 && v.sw == 3 && v.sh == (int64_t)3 * v.w && v.sn == (int64_t)3 * v.w * v.h &&
         (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0 && IW % 8 == 0 && IW <= 256;
}

// channel c of eight interleaved RGB pixels held as twelve 32-bit words (24 bf16): the four pixel pairs
template <int CH>
__device__ __forceinline__ void rgb_pairs(const uint32_t (&w)[12], uint32_t (&q)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ka = 3 * (2 * i) + CH, kb = ka + 3;                  // element indices of pixels 2i and 2i+1
    const uint32_t sel = ((ka & 1) ? 0x32u : 0x10u) | (((kb & 1) ? 0x76u : 0x54u) << 8);
    q[i] = __byte_perm(w[ka >> 1], w[kb >> 1], sel);
  }
}

template <int NC>
__device__ __forceinline__ bool stage_fine_rows(__nv_bfloat16* S, int pitch, int rows, const ThinArgs& a, int n, int ih0) {
  const int IH = 2 * a.H, IW = 2 * a.W, groups = IW >> 3;
  const bool has_ref = a.fine_ref.ptr != nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint32_t* S32 = reinterpret_cast<uint32_t*>(S);
  if (rows_planar_ok(a.fine, IW) && (!has_ref || rows_planar_ok(a.fine_ref, IW))) {
    // (one row per trip: issuing two rows' loads before converting either was measured -- it pushes these kernels past 128 registers, one resident CTA
    //  per SM, and every consumer of this routine lost 25-60 %)
    for (int r = warp; r < NC * rows; r += nwarps) {
      const int ci = r / rows, row = r - ci * rows, ih = ih0 + row;
      uint32_t q[4] = {0u, 0u, 0u, 0u};
      if (lane < groups && (unsigned)ih < (unsigned)IH) {
        float v[8];
        load8_planar(a.fine, (int64_t)n * a.fine.sn + (int64_t)ih * a.fine.sh + (int64_t)ci * a.fine.sc + 8 * lane, v);
        if (has_ref) {
          float rf[8];
          load8_planar(a.fine_ref, (int64_t)n * a.fine_ref.sn + (int64_t)ih * a.fine_ref.sh + (int64_t)ci * a.fine_ref.sc + 8 * lane, rf);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] *= act_grad_from_output(rf[e], a.fine_act, a.slope);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) q[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
      }
      row_words_store(S32 + ((ci * rows + row) * pitch >> 1), lane, groups, q[0], q[1], q[2], q[3]);
    }
    return true;
  }
  if (NC == 3 && rows_rgb_ok(a.fine, IW) && (!has_ref || rows_rgb_ok(a.fine_ref, IW))) {
    for (int row = warp; row < rows; row += nwarps) {
      const int ih = ih0 + row;
      uint32_t w[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) w[k] = 0u;
      if (lane < groups && (unsigned)ih < (unsigned)IH) {
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.fine.ptr) + (int64_t)n * a.fine.sn + (int64_t)ih * a.fine.sh + 24 * lane);
        const uint4 t0 = __ldg(src), t1 = __ldg(src + 1), t2 = __ldg(src + 2);
        w[0] = t0.x; w[1] = t0.y; w[2] = t0.z; w[3] = t0.w; w[4] = t1.x; w[5] = t1.y; w[6] = t1.z; w[7] = t1.w; w[8] = t2.x; w[9] = t2.y; w[10] = t2.z; w[11] = t2.w;
        if (has_ref) {
          const uint4* rs = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.fine_ref.ptr) + (int64_t)n * a.fine_ref.sn + (int64_t)ih * a.fine_ref.sh + 24 * lane);
          const uint4 r0 = __ldg(rs), r1 = __ldg(rs + 1), r2 = __ldg(rs + 2);
          const uint32_t rw[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
#pragma unroll
          for (int k = 0; k < 12; ++k) {
            const float lo = __uint_as_float(w[k] << 16) * act_grad_from_output(__uint_as_float(rw[k] << 16), a.fine_act, a.slope);
            const float hi = __uint_as_float(w[k] & 0xffff0000u) * act_grad_from_output(__uint_as_float(rw[k] & 0xffff0000u), a.fine_act, a.slope);
            w[k] = pack_bf16x2(lo, hi);
          }
        }
      }
      uint32_t q[4];
      rgb_pairs<0>(w, q);
      row_words_store(S32 + ((0 * rows + row) * pitch >> 1), lane, groups, q[0], q[1], q[2], q[3]);
      rgb_pairs<1>(w, q);
      row_words_store(S32 + ((1 * rows + row) * pitch >> 1), lane, groups, q[0], q[1], q[2], q[3]);
      rgb_pairs<2>(w, q);
      row_words_store(S32 + ((2 * rows + row) * pitch >> 1), lane, groups, q[0], q[1], q[2], q[3]);
    }
    return true;
  }
  return false;
}

// Stage fine rows [ih0, ih0+rows) of image n as bf16: S[(ci*rows + row)*pitch + s], s = iw + 1 in [0, IW+1]; s = 0 and
// s = IW+1 (iw = -1, IW) and rows outside the image are the zero padding of the convolution.  IW % 4 == 0.
// Loads are issued four deep per thread before anything is stored: the staging loops are what keeps HBM busy here
// (each SM needs ~40 KB in flight), a one-load-at-a-time loop runs at a fifth of the bandwidth.
template <int NC>
__device__ __forceinline__ void stage_fine(__nv_bfloat16* S, int pitch, int rows, const ThinArgs& a, int n, int ih0) {
  if (stage_fine_rows<NC>(S, pitch, rows, a, n, ih0)) return;      // (uniform across the CTA: it depends on the views only)
  const int IH = 2 * a.H, IW = 2 * a.W, IW4 = IW >> 2;
  const int total = NC * rows * IW4;
  constexpr int U = 4;
  for (int base = threadIdx.x; base < total; base += U * blockDim.x) {
    float v[U][4], r[U][4];
    int dst[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * blockDim.x;
      dst[u] = -1;
#pragma unroll
      for (int e = 0; e < 4; ++e) { v[u][e] = 0.f; r[u][e] = 0.f; }
      if (idx < total) {
        const int j = idx % IW4;
        const int t = idx / IW4;
        const int row = t % rows, ci = t / rows;
        const int ih = ih0 + row;
        dst[u] = (ci * rows + row) * pitch + 4 * j + 1;
        if ((unsigned)ih < (unsigned)IH) {
          load4_rt(a.fine, (int64_t)n * a.fine.sn + (int64_t)ih * a.fine.sh + (int64_t)(4 * j) * a.fine.sw + (int64_t)ci * a.fine.sc, a.fine_vec, v[u]);
          if (a.fine_ref.ptr)
            load4_rt(a.fine_ref, (int64_t)n * a.fine_ref.sn + (int64_t)ih * a.fine_ref.sh + (int64_t)(4 * j) * a.fine_ref.sw + (int64_t)ci * a.fine_ref.sc,
                     a.ref_vec, r[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (dst[u] < 0) continue;
      if (a.fine_ref.ptr) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[u][e] *= act_grad_from_output(r[u][e], a.fine_act, a.slope);
      }
      __nv_bfloat16* d = S + dst[u];                                   // odd index: [1] [2,3] [4]
      d[0] = __float2bfloat16_rn(v[u][0]);
      *reinterpret_cast<uint32_t*>(d + 1) = pack_bf16x2(v[u][1], v[u][2]);
      d[3] = __float2bfloat16_rn(v[u][3]);
    }
  }
  for (int idx = threadIdx.x; idx < NC * rows * 2; idx += blockDim.x)
    S[(idx >> 1) * pitch + ((idx & 1) ? IW + 1 : 0)] = __float2bfloat16_rn(0.f);
}

constexpr int CP = 40;   // bf16 elements per staged coarse pixel: 32 channels + 8 pad (80 B pitch: conflict-free ldmatrix)

__device__ __forceinline__ void cp_async_16_zfill(void* smem_dst, const void* gmem_src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;                                     // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Stage coarse rows [q0, q0+rows) x cols [c0, c0+cols) of image n (zeros outside the image) as S[(row*cols + col)*CP + ch].
// Plain operand: 16-byte cp.async straight into shared memory (every copy of the tile in flight at once); with the fused
// activation backward: four-deep register staging.  The caller must cp_async_wait_all() + __syncthreads() before reading.
__device__ __forceinline__ void stage_coarse(__nv_bfloat16* S, int rows, int cols, const ThinArgs& a, int n, int q0, int c0) {
  const int total = rows * cols * 4;
  if (!a.coarse_ref) {
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int chunk = idx & 3, pix = idx >> 2;
      const int col = pix % cols, row = pix / cols;
      const int q = q0 + row, r = c0 + col;
      const bool ok = (unsigned)q < (unsigned)a.H && (unsigned)r < (unsigned)a.W;
      const int64_t off = ok ? (((int64_t)n * a.H + q) * a.W + r) * 32 + chunk * 8 : 0;
      cp_async_16_zfill(S + pix * CP + chunk * 8, a.coarse + off, ok);
    }
    return;
  }
  constexpr int U = 4;
  for (int base = threadIdx.x; base < total; base += U * blockDim.x) {
    uint4 val[U], rv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * blockDim.x;
      val[u] = make_uint4(0u, 0u, 0u, 0u);
      rv[u] = make_uint4(0u, 0u, 0u, 0u);
      if (idx < total) {
        const int chunk = idx & 3, pix = idx >> 2;
        const int col = pix % cols, row = pix / cols;
        const int q = q0 + row, r = c0 + col;
        if ((unsigned)q < (unsigned)a.H && (unsigned)r < (unsigned)a.W) {
          const int64_t off = (((int64_t)n * a.H + q) * a.W + r) * 32 + chunk * 8;
          val[u] = __ldg(reinterpret_cast<const uint4*>(a.coarse + off));
          rv[u] = __ldg(reinterpret_cast<const uint4*>(a.coarse_ref + off));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * blockDim.x;
      if (idx >= total) continue;
      float d[8], f[8];
      unpack8(val[u], d);
      unpack8(rv[u], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) d[e] *= act_grad_from_output(f[e], a.coarse_act, a.slope);
      *reinterpret_cast<uint4*>(S + (idx >> 2) * CP + (idx & 3) * 8) = pack8(d);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// DOWN: coarse[n,oh,ow,:] = act( sum_{kh,kw,ci} fine[n,2oh-1+kh,2ow-1+kw,ci] * w[:,ci,kh,kw] )
// GEMM per 16 consecutive ow: M = 16 pixels, K = 16 taps (one k-step per input channel), N = 32 channels.
// MMA column n = 8j + g of n-tile j is mapped to channel 8*(g>>1) + 2j + (g&1), so that lane (g,t) ends up with the
// eight consecutive channels 8t..8t+7 of its pixel: one 16-byte store per pixel row, fully coalesced across the warp.
// ---------------------------------------------------------------------------------------------------
// (Measured, round 2: at NC = 3 this kernel holds 106-128 registers = two resident CTAs per SM.  Capping it at 80 registers for three CTAs is SLOWER, before
//  and after the row-per-warp staging (same-box A/B: 247 -> 261 us on the fp32 batch, 171 -> 193 us on the bf16 one): what it spends its time on is
//  instruction issue, not latency -- ncu put the issue slots at 57 % busy with 24 % of the warp slots filled, most of it the old staging loop.)
template <int NC, int EPI>
__global__ void __launch_bounds__(256) thin_down_mma_kernel(const ThinArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* S = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int rows = 2 * a.R + 2, pitch = 2 * a.W + 2;
  uint8_t* Sy = smem_raw + ((NC * rows * pitch * 2 + 127) & ~127);      // EPI 2: the saved-output tile behind the image band
  uint32_t bw[NC][4][2];
#pragma unroll
  for (int ci = 0; ci < NC; ++ci)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ch = 8 * (g >> 1) + 2 * j + (g & 1);
      const float* wp = a.w + (ch * NC + ci) * 16;
      bw[ci][j][0] = pack_bf16x2(__ldg(wp + 2 * t), __ldg(wp + 2 * t + 1));
      bw[ci][j][1] = pack_bf16x2(__ldg(wp + 2 * t + 8), __ldg(wp + 2 * t + 9));
    }
  const int WB = a.W >> 4, mtiles = a.R * WB;
  const int kh_lo = t >> 1, kw0 = (t & 1) * 2;
  float ssum[8], ssq[8], cf_scale[8], cf_shift[8], cf_mean[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    ssum[c] = 0.f; ssq[c] = 0.f;
    cf_scale[c] = EPI == 2 ? a.prev_scale[8 * t + c] : 1.f;
    cf_shift[c] = EPI == 2 ? a.prev_shift[8 * t + c] : 0.f;
    cf_mean[c] = EPI == 2 ? a.prev_mean[8 * t + c] : 0.f;
  }
  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int n = tile / a.tiles_per_img, oh0 = (tile - n * a.tiles_per_img) * a.R;
    __syncthreads();                       // previous tile fully consumed
    if (EPI == 2) {
      // the saved forward output under this tile (R rows x W pixels x 64 bytes) goes to shared memory by 16-byte cp.async, all of it
      // in flight while the image band is staged: fetched from registers one m-tile ahead, each warp had 1 KB in flight and
      // the stream ran at 1.6 TB/s
      const int total = a.R * a.W * 4;
      const __nv_bfloat16* src = a.prev_y + ((int64_t)n * a.H + oh0) * a.W * 32;
      for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int pl = idx >> 2, chunk = idx & 3;
        cp_async_16_zfill(Sy + sw64(pl, chunk), src + (int64_t)pl * 32 + chunk * 8, oh0 + pl / a.W < a.H);
      }
    }
    stage_fine<NC>(S, pitch, rows, a, n, 2 * oh0 - 1);
    if (EPI == 2) cp_async_wait_all();
    __syncthreads();
    for (int mt = warp; mt < mtiles; mt += (blockDim.x >> 5)) {
      const int rr = mt / WB, c = mt - rr * WB;
      const int oh = oh0 + rr;
      if (oh >= a.H) break;
      uint4 yc0 = make_uint4(0u, 0u, 0u, 0u), yc1 = yc0;
      if (EPI == 2) {
        const int pl0 = rr * a.W + 16 * c + g;
        yc0 = *reinterpret_cast<const uint4*>(Sy + sw64(pl0, t));
        yc1 = *reinterpret_cast<const uint4*>(Sy + sw64(pl0 + 8, t));
      }
      float acc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
      const int ow = 16 * c + g;
#pragma unroll
      for (int ci = 0; ci < NC; ++ci) {
        const __nv_bfloat16* base = S + (ci * rows + 2 * rr + kh_lo) * pitch + 2 * ow + kw0;
        uint32_t af[4];
        af[0] = *reinterpret_cast<const uint32_t*>(base);
        af[1] = *reinterpret_cast<const uint32_t*>(base + 16);
        af[2] = *reinterpret_cast<const uint32_t*>(base + 2 * pitch);
        af[3] = *reinterpret_cast<const uint32_t*>(base + 2 * pitch + 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_bf16_16816(acc[j], af, bw[ci][j][0], bw[ci][j][1]);
      }
      if (a.out_act == B200GAN_ACT_LRELU) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[j][e] = acc[j][e] > 0.f ? acc[j][e] : acc[j][e] * a.slope;
      } else if (a.out_act == B200GAN_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[j][e] = fmaxf(acc[j][e], 0.f);
      }
      const int64_t ooff = (((int64_t)n * a.H + oh) * a.W + ow) * a.cpitch + 8 * t;
      // lane (g,t) holds channels 8t..8t+7 of pixels ow and ow+8: value index c = 2j+e <-> acc[j][e] (row g), acc[j][2+e] (row g+8)
      float yv[2][8];
      if (EPI == 2) {
        unpack8(yc0, yv[0]);
        unpack8(yc1, yv[1]);
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float z = fmaf(yv[hrow][c], cf_scale[c], cf_shift[c]);
            acc[c >> 1][2 * hrow + (c & 1)] *= z > 0.f ? 1.f : a.prev_neg;
          }
      }
      uint4 lo, hi;
      lo.x = pack_bf16x2(acc[0][0], acc[0][1]); lo.y = pack_bf16x2(acc[1][0], acc[1][1]);
      lo.z = pack_bf16x2(acc[2][0], acc[2][1]); lo.w = pack_bf16x2(acc[3][0], acc[3][1]);
      hi.x = pack_bf16x2(acc[0][2], acc[0][3]); hi.y = pack_bf16x2(acc[1][2], acc[1][3]);
      hi.z = pack_bf16x2(acc[2][2], acc[2][3]); hi.w = pack_bf16x2(acc[3][2], acc[3][3]);
      __nv_bfloat16* o = a.coarse_out + ooff;
      *reinterpret_cast<uint4*>(o) = lo;
      *reinterpret_cast<uint4*>(o + 8 * a.cpitch) = hi;
      if (EPI != 0) {
        // per-thread channel sums over every pixel this thread produces (the thread's eight channels never change);
        // the statistics are those of the stored (bf16-rounded) values
        float r0[8], r1[8];
        unpack8(lo, r0);
        unpack8(hi, r1);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          ssum[c] += r0[c] + r1[c];
          if (EPI == 1) ssq[c] = fmaf(r0[c], r0[c], fmaf(r1[c], r1[c], ssq[c]));
          else ssq[c] = fmaf(r0[c], yv[0][c] - cf_mean[c], fmaf(r1[c], yv[1][c] - cf_mean[c], ssq[c]));
        }
      }
    }
  }
  if (EPI != 0) {
    // reduce over the eight row lanes g (lane = 4g + t), then over the warps through shared memory, one fp64 atomic per channel
    float* red = reinterpret_cast<float*>(smem_raw);           // [2][32], the staging tile is no longer needed
    __syncthreads();
    if (threadIdx.x < 64) red[threadIdx.x] = 0.f;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float v0 = ssum[c], v1 = ssq[c];
#pragma unroll
      for (int off = 4; off < 32; off <<= 1) { v0 += __shfl_xor_sync(0xffffffffu, v0, off); v1 += __shfl_xor_sync(0xffffffffu, v1, off); }
      if (g == 0) { atomicAdd(&red[8 * t + c], v0); atomicAdd(&red[32 + 8 * t + c], v1); }
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      const int c = threadIdx.x & 31;
      const double scale = (EPI == 2 && threadIdx.x >= 32) ? (double)a.prev_invstd[c] : 1.0;
      atomicAdd(a.sums + threadIdx.x, (double)red[threadIdx.x] * scale);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// UP: fine[n,2q+py,2r+px,ci] = act( sum_{co} sum_{di,dj} coarse[n,q+di,r+dj,co] * w[co,ci,py+1-2di,px+1-2dj] )
// GEMM per 16 consecutive r: M = 16 coarse positions, K = 9 neighbours x 32 channels (18 k-steps), N = 4 parity classes x nc
// (zero weights where a neighbour does not reach a class: the MMA work is free, the 32-channel tensor is read once).
// ---------------------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(256) thin_up_mma_kernel(const ThinArgs a) {
  constexpr int NT = (4 * NC + 7) / 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint2* Bf = reinterpret_cast<uint2*>(smem_raw);                                 // [18*NT][32] fragment-ordered weights
  __nv_bfloat16* S = reinterpret_cast<__nv_bfloat16*>(smem_raw + 18 * NT * 32 * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int idx = threadIdx.x; idx < 18 * NT * 32; idx += blockDim.x) {
    const int ln = idx & 31, f = idx >> 5, nt = f % NT, ks = f / NT, nbr = ks >> 1, h = ks & 1;
    const int gg = ln >> 2, tt = ln & 3, nn = 8 * nt + gg;
    float wv[4] = {0.f, 0.f, 0.f, 0.f};
    if (nn < 4 * NC) {
      const int cls = nn / NC, ci = nn - cls * NC, py = cls >> 1, px = cls & 1;
      const int di = nbr / 3 - 1, dj = nbr % 3 - 1, kh = py + 1 - 2 * di, kw = px + 1 - 2 * dj;
      if (kh >= 0 && kh < 4 && kw >= 0 && kw < 4) {
        const int kk[4] = {2 * tt, 2 * tt + 1, 2 * tt + 8, 2 * tt + 9};
#pragma unroll
        for (int e = 0; e < 4; ++e) wv[e] = __ldg(a.w + ((16 * h + kk[e]) * NC + ci) * 16 + kh * 4 + kw);
      }
    }
    Bf[idx] = make_uint2(pack_bf16x2(wv[0], wv[1]), pack_bf16x2(wv[2], wv[3]));
  }
  const int cols = a.W + 2, WB = a.W >> 4, mtiles = a.R * WB;
  const int prow = (lane & 7) + 8 * ((lane >> 3) & 1), koff = 8 * (lane >> 4);
  const bool pair_store = NC == 1 && a.fine_vec;
  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int n = tile / a.tiles_per_img, q0 = (tile - n * a.tiles_per_img) * a.R;
    __syncthreads();
    stage_coarse(S, a.R + 2, cols, a, n, q0 - 1, -1);
    cp_async_wait_all();
    __syncthreads();
    for (int mt = warp; mt < mtiles; mt += 8) {
      const int rr = mt / WB, c = mt - rr * WB;
      const int q = q0 + rr;
      if (q >= a.H) break;
      float acc[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
#pragma unroll
      for (int nbr = 0; nbr < 9; ++nbr) {
        const int di = nbr / 3 - 1, dj = nbr % 3 - 1;
        const __nv_bfloat16* base = S + ((rr + 1 + di) * cols + 16 * c + prow + 1 + dj) * CP + koff;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t af[4];
          ldmatrix_x4(af, base + 16 * h);
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            const uint2 b = Bf[((nbr * 2 + h) * NT + j) * 32 + lane];
            mma_bf16_16816(acc[j], af, b.x, b.y);
          }
        }
      }
      // lane (g,t): rows (coarse positions) r = 16c+g and +8; columns n = 8j + 2t + e -> (class, ci)
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int r = 16 * c + g + 8 * half;
          float v0 = acc[j][2 * half], v1 = acc[j][2 * half + 1];
          if (a.out_act == B200GAN_ACT_TANH) { v0 = tanh_fast(v0); v1 = tanh_fast(v1); }
          const int n0 = 8 * j + 2 * t;
          if (pair_store) {
            if (n0 < 4) {                                                  // n0 = 2*py, columns px = 0,1 are adjacent pixels
              const int64_t off = (int64_t)n * a.fine.sn + (int64_t)(2 * q + (n0 >> 1)) * a.fine.sh + 2 * r;
              if (a.fine.dtype == B200GAN_F32) *reinterpret_cast<float2*>(reinterpret_cast<float*>(a.fine.ptr) + off) = make_float2(v0, v1);
              else *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(a.fine.ptr) + off) = pack_bf16x2(v0, v1);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int nn = n0 + e;
              if (nn < 4 * NC) {
                const int cls = nn / NC, ci = nn - cls * NC;
                const int64_t off = (int64_t)n * a.fine.sn + (int64_t)(2 * q + (cls >> 1)) * a.fine.sh + (int64_t)(2 * r + (cls & 1)) * a.fine.sw +
                                    (int64_t)ci * a.fine.sc;
                const float v = e ? v1 : v0;
                if (a.fine.dtype == B200GAN_F32) reinterpret_cast<float*>(a.fine.ptr)[off] = v;
                else reinterpret_cast<__nv_bfloat16*>(a.fine.ptr)[off] = __float2bfloat16_rn(v);
              }
            }
          }
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// WGRAD: dw[co,ci,kh,kw] += sum_{n,oh,ow} coarse[n,oh,ow,co] * fine[n,2oh-1+kh,2ow-1+kw,ci]
// GEMM: M = 32 channels (two m-tiles, A = coarse^T through ldmatrix.trans), N = 16 taps per input channel (two n-tiles),
// K = pixels, 16 consecutive ow per k-step.  Each warp keeps its accumulators over the whole kernel; one shared-memory and
// one global fp32 atomic reduction per CTA at the end.
// ---------------------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(256) thin_wgrad_mma_kernel(const ThinArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* red = reinterpret_cast<float*>(smem_raw);                                  // [512*NC]
  __nv_bfloat16* Sc = reinterpret_cast<__nv_bfloat16*>(smem_raw + 512 * NC * 4);    // [R][W][CP]
  const int rows = 2 * a.R + 2, pitch = 2 * a.W + 2;
  __nv_bfloat16* Sf = Sc + a.R * a.W * CP;                                          // [NC][rows][pitch]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int i = threadIdx.x; i < 512 * NC; i += blockDim.x) red[i] = 0.f;
  float acc[2][2 * NC][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < 2 * NC; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[m][j][e] = 0.f;
  const int WB = a.W >> 4, ksteps = a.R * WB;
  const int px_l = (lane & 7) + 8 * (lane >> 4), co_l = 8 * ((lane >> 3) & 1);
  const int kw = g & 3, khg = g >> 2;
  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int n = tile / a.tiles_per_img, oh0 = (tile - n * a.tiles_per_img) * a.R;
    __syncthreads();
    stage_coarse(Sc, a.R, a.W, a, n, oh0, 0);
    stage_fine<NC>(Sf, pitch, rows, a, n, 2 * oh0 - 1);
    cp_async_wait_all();
    __syncthreads();
    for (int ks = warp; ks < ksteps; ks += 8) {
      const int rr = ks / WB, c = ks - rr * WB;
      uint32_t af[2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m) ldmatrix_x4_trans(af[m], Sc + (rr * a.W + 16 * c + px_l) * CP + 16 * m + co_l);
#pragma unroll
      for (int ci = 0; ci < NC; ++ci)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const unsigned short* p = reinterpret_cast<const unsigned short*>(Sf) + (ci * rows + 2 * rr + 2 * hf + khg) * pitch + 2 * (16 * c + 2 * t) + kw;
          const uint32_t b0 = (uint32_t)p[0] | ((uint32_t)p[2] << 16);
          const uint32_t b1 = (uint32_t)p[16] | ((uint32_t)p[18] << 16);
#pragma unroll
          for (int m = 0; m < 2; ++m) mma_bf16_16816(acc[m][ci * 2 + hf], af[m], b0, b1);
        }
    }
  }
  // acc[m][ci*2+hf][e]: co = 16m + g + 8*(e>>1), tap = 8hf + 2t + (e&1)
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < 2 * NC; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int co = 16 * m + g + 8 * (e >> 1), ci = j >> 1, tap = 8 * (j & 1) + 2 * t + (e & 1);
        atomicAdd(&red[(co * NC + ci) * 16 + tap], acc[m][j][e]);
      }
  __syncthreads();
  for (int i = threadIdx.x; i < 512 * NC; i += blockDim.x) atomicAdd(a.dw + i, red[i]);
}

// ---------------------------------------------------------------------------------------------------
// TMA-staged variants for a PLAIN coarse operand (no fused activation backward on it): the 32-channel tensor -- 411 MB per
// launch at the benchmark size, the whole cost of these kernels -- is brought in by one cp.async.bulk.tensor per tile
// (zero halo = TMA out-of-bounds fill, 64B-swizzled so that ldmatrix is conflict free without padding), double buffered
// behind mbarriers by a producer warp, so no thread spends instructions on staging it.  Seven consumer warps own the seven
// 16-pixel column blocks of a 112-wide row band (any W % 16 == 0 with W/16 <= 7 consumer warps... see thin_tma_ok).
// ---------------------------------------------------------------------------------------------------
constexpr int kThinStages = 2;

template <int NC>
__global__ void __launch_bounds__(256) thin_up_tma_kernel(const __grid_constant__ CUtensorMap map_c, const ThinArgs a) {
  constexpr int NT = (4 * NC + 7) / 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  const int cols = a.W + 2, rows = a.RT + 2;
  const uint32_t tile_bytes = (uint32_t)rows * cols * 64u;
  const uint32_t stage_bytes = (tile_bytes + 1023u) & ~1023u;
  uint2* Bf = reinterpret_cast<uint2*>(smem + kThinStages * stage_bytes);                 // [18*NT][32]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kThinStages * stage_bytes + 18 * NT * 32 * 8);
  uint64_t* empty_bar = full_bar + kThinStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int WB = a.W >> 4;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kThinStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], WB); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
  }
  for (int idx = threadIdx.x; idx < 18 * NT * 32; idx += blockDim.x) {
    const int ln = idx & 31, f = idx >> 5, nt = f % NT, ks = f / NT, nbr = ks >> 1, h = ks & 1;
    const int gg = ln >> 2, tt = ln & 3, nn = 8 * nt + gg;
    float wv[4] = {0.f, 0.f, 0.f, 0.f};
    if (nn < 4 * NC) {
      const int cls = nn / NC, ci = nn - cls * NC, py = cls >> 1, px = cls & 1;
      const int di = nbr / 3 - 1, dj = nbr % 3 - 1, kh = py + 1 - 2 * di, kw = px + 1 - 2 * dj;
      if (kh >= 0 && kh < 4 && kw >= 0 && kw < 4) {
        const int kk[4] = {2 * tt, 2 * tt + 1, 2 * tt + 8, 2 * tt + 9};
#pragma unroll
        for (int e = 0; e < 4; ++e) wv[e] = __ldg(a.w + ((16 * h + kk[e]) * NC + ci) * 16 + kh * 4 + kw);
      }
    }
    Bf[idx] = make_uint2(pack_bf16x2(wv[0], wv[1]), pack_bf16x2(wv[2], wv[3]));
  }
  __syncthreads();

  if (warp == 7) {
    // ===== producer: one elected lane issues one bulk tensor copy per tile =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int n = tile / a.tiles_per_img, q0 = (tile - n * a.tiles_per_img) * a.RT;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], tile_bytes);
        tma_load_4d(smem + s * stage_bytes, &map_c, &full_bar[s], 0, -1, q0 - 1, n);
        if (++s == kThinStages) { s = 0; ph ^= 1; }
      }
    }
    return;
  }
  if (warp >= WB) return;                     // one consumer warp per 16-pixel column block
  const int c = warp;
  const int prow = (lane & 7) + 8 * ((lane >> 3) & 1), kchunk = lane >> 4;
  const bool pair_store = NC == 1 && a.fine_vec;
  // Dense interleaved RGB bf16 result (the image the fused trainer keeps between the networks): the warp's 2 x 32 pixels x 3 channels are gathered in a
  // 384-byte shared-memory scratch and leave as 24 16-byte stores.  (At nc = 1 the paired 4-byte stores are already one 128-byte line per instruction:
  // the staged form measured 168 us against 154 us.)  (ncu, round 2: the per-element path issued 3.2 M two-byte store requests per launch
  // for 154 MB; the kernel sat at 0.28 of the copy bandwidth with the LSU as its busiest unit.)
  const bool rgb_store = NC == 3 && a.fine.dtype == B200GAN_BF16 && a.fine.sc == 1 && a.fine.sw == NC && a.fine.sh == (int64_t)2 * NC * a.W &&
                         a.fine.sn % 8 == 0 && (reinterpret_cast<uintptr_t>(a.fine.ptr) & 15) == 0;
  __nv_bfloat16* scratch = reinterpret_cast<__nv_bfloat16*>(smem + kThinStages * stage_bytes + 18 * NT * 32 * 8 + 64) + warp * 192;
  int s = 0;
  uint32_t ph = 0;
  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int n = tile / a.tiles_per_img, q0 = (tile - n * a.tiles_per_img) * a.RT;
    mbar_wait(&full_bar[s], ph);
    const uint32_t S = smem_u32(smem + s * stage_bytes);
    // two rows at a time: two independent accumulator chains per warp hide the MMA latency
    for (int rr = 0; rr < a.RT; rr += 2) {
      float acc[2][NT][4];
#pragma unroll
      for (int r2 = 0; r2 < 2; ++r2)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[r2][j][e] = 0.f;
#pragma unroll
      for (int nbr = 0; nbr < 9; ++nbr) {
        const int di = nbr / 3 - 1, dj = nbr % 3 - 1;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t af[2][4];
#pragma unroll
          for (int r2 = 0; r2 < 2; ++r2) {
            const int pix = (rr + r2 + 1 + di) * cols + 16 * c + prow + 1 + dj;
            const uint32_t addr = S + sw64(pix, 2 * h + kchunk);
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                         : "=r"(af[r2][0]), "=r"(af[r2][1]), "=r"(af[r2][2]), "=r"(af[r2][3]) : "r"(addr));
          }
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            const uint2 b = Bf[((nbr * 2 + h) * NT + j) * 32 + lane];
#pragma unroll
            for (int r2 = 0; r2 < 2; ++r2) mma_bf16_16816(acc[r2][j], af[r2], b.x, b.y);
          }
        }
      }
#pragma unroll
      for (int r2 = 0; r2 < 2; ++r2) {
        const int q = q0 + rr + r2;
        if (rr + r2 >= a.RT || q >= a.H) continue;
        if (rgb_store) {
#pragma unroll
          for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int half = 0; half < 2; ++half)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int nn = 8 * j + 2 * t + e;
                if (nn < 4 * NC) {
                  float v = acc[r2][j][2 * half + e];
                  if (a.out_act == B200GAN_ACT_TANH) v = tanh_fast(v);
                  const int cls = nn / NC, ci = nn - cls * NC;
                  scratch[(cls >> 1) * 32 * NC + (2 * (g + 8 * half) + (cls & 1)) * NC + ci] = __float2bfloat16_rn(v);
                }
              }
          __syncwarp();
          if (lane < 8 * NC) {                                     // two fine rows of 32 pixels x NC channels: 4 NC 16-byte chunks each
            const int row = lane / (4 * NC), chunk = lane - row * 4 * NC;
            const uint4 v = *reinterpret_cast<const uint4*>(scratch + row * 32 * NC + chunk * 8);
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(a.fine.ptr) + (int64_t)n * a.fine.sn + (int64_t)(2 * q + row) * a.fine.sh + 32 * NC * c + chunk * 8;
            *reinterpret_cast<uint4*>(dst) = v;
          }
          __syncwarp();
          continue;
        }
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int r = 16 * c + g + 8 * half;
            float v0 = acc[r2][j][2 * half], v1 = acc[r2][j][2 * half + 1];
            if (a.out_act == B200GAN_ACT_TANH) { v0 = tanh_fast(v0); v1 = tanh_fast(v1); }
            const int n0 = 8 * j + 2 * t;
            if (pair_store) {
              if (n0 < 4) {                                                  // n0 = 2*py, columns px = 0,1 are adjacent pixels
                const int64_t off = (int64_t)n * a.fine.sn + (int64_t)(2 * q + (n0 >> 1)) * a.fine.sh + 2 * r;
                if (a.fine.dtype == B200GAN_F32) *reinterpret_cast<float2*>(reinterpret_cast<float*>(a.fine.ptr) + off) = make_float2(v0, v1);
                else *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(a.fine.ptr) + off) = pack_bf16x2(v0, v1);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int nn = n0 + e;
                if (nn < 4 * NC) {
                  const int cls = nn / NC, ci = nn - cls * NC;
                  const int64_t off = (int64_t)n * a.fine.sn + (int64_t)(2 * q + (cls >> 1)) * a.fine.sh + (int64_t)(2 * r + (cls & 1)) * a.fine.sw +
                                      (int64_t)ci * a.fine.sc;
                  const float v = e ? v1 : v0;
                  if (a.fine.dtype == B200GAN_F32) reinterpret_cast<float*>(a.fine.ptr)[off] = v;
                  else reinterpret_cast<__nv_bfloat16*>(a.fine.ptr)[off] = __float2bfloat16_rn(v);
                }
              }
            }
          }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);     // this warp no longer reads the stage
    if (++s == kThinStages) { s = 0; ph ^= 1; }
  }
}

// WGRAD with the coarse operand staged by TMA and the image band staged RAW (its own dtype TF, optionally together with the
// reference tensor of the fused tanh backward) by 16-byte cp.async of the producer warp; the bf16 conversion and the
// dy * (1 - a^2) product happen when the consumers build their B fragments.  Requirements (checked on the host): unit W stride
// and 16-byte aligned rows of the image tensors.
template <typename TF> __device__ __forceinline__ float band_ld(const TF* p);
template <> __device__ __forceinline__ float band_ld<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float band_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <int NC, typename TF, bool HAS_REF>
__global__ void __launch_bounds__(256) thin_wgrad_tma_kernel(const __grid_constant__ CUtensorMap map_c, const ThinArgs a) {
  constexpr int PADL = 16 / (int)sizeof(TF);                 // image column iw lives at element iw + PADL (16-byte aligned rows)
  constexpr int NB = HAS_REF ? 2 : 1;                        // bands per stage: gradient (+ reference)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  const int frows = 2 * a.RT + 2, IH = 2 * a.H, IW = 2 * a.W, pitch = IW + 2 * PADL;
  const uint32_t tile_bytes = (uint32_t)a.RT * a.W * 64u;
  const uint32_t band_bytes = (uint32_t)NC * frows * pitch * (uint32_t)sizeof(TF);
  const uint32_t stage_bytes = (tile_bytes + NB * band_bytes + 1023u) & ~1023u;
  float* red = reinterpret_cast<float*>(smem + kThinStages * stage_bytes);                // [512*NC]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(red + 512 * NC);
  uint64_t* empty_bar = full_bar + kThinStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int WB = a.W >> 4;
  if (threadIdx.x == 0) {
    // full: expect_tx arrival of the TMA issue + one cp.async-completion arrival per producer lane
    for (int s = 0; s < kThinStages; ++s) { mbar_init(&full_bar[s], 33); mbar_init(&empty_bar[s], WB); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
  }
  for (int i = threadIdx.x; i < 512 * NC; i += blockDim.x) red[i] = 0.f;
  // the zero padding columns (iw = -1 and iw = IW) are never written by the copies: clear the bands once
  for (int s = 0; s < kThinStages; ++s) {
    uint32_t* z = reinterpret_cast<uint32_t*>(smem + s * stage_bytes + tile_bytes);
    for (int i = threadIdx.x; i < (int)(NB * band_bytes / 4); i += blockDim.x) z[i] = 0u;
  }
  __syncthreads();

  if (warp == 7) {
    int s = 0;
    uint32_t ph = 0;
    const int chunks = IW / PADL;                            // 16-byte chunks per image row
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const int n = tile / a.tiles_per_img, oh0 = (tile - n * a.tiles_per_img) * a.RT;
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* st = smem + s * stage_bytes;
      if (lane == 0) {
        mbar_expect_tx(&full_bar[s], tile_bytes);
        tma_load_4d(st, &map_c, &full_bar[s], 0, 0, oh0, n);
      }
      const int ih0 = 2 * oh0 - 1, total = NC * frows * chunks;
      for (int idx = lane; idx < total; idx += 32) {
        const int j = idx % chunks, tt = idx / chunks, row = tt % frows, ci = tt / frows, ih = ih0 + row;
        const bool ok = (unsigned)ih < (unsigned)IH;
        const uint32_t dst = (uint32_t)((ci * frows + row) * pitch + PADL + j * PADL) * (uint32_t)sizeof(TF);
        const int64_t off = ok ? (int64_t)n * a.fine.sn + (int64_t)ih * a.fine.sh + (int64_t)ci * a.fine.sc + j * PADL : 0;
        cp_async_16_zfill(st + tile_bytes + dst, reinterpret_cast<const TF*>(a.fine.ptr) + off, ok);
        if (HAS_REF) {
          const int64_t roff = ok ? (int64_t)n * a.fine_ref.sn + (int64_t)ih * a.fine_ref.sh + (int64_t)ci * a.fine_ref.sc + j * PADL : 0;
          cp_async_16_zfill(st + tile_bytes + band_bytes + dst, reinterpret_cast<const TF*>(a.fine_ref.ptr) + roff, ok);
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full_bar[s])) : "memory");
      if (++s == kThinStages) { s = 0; ph ^= 1; }
    }
    return;
  }
  float acc[2][2 * NC][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < 2 * NC; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[m][j][e] = 0.f;
  if (warp < WB) {
    const int c = warp;
    const int px_l = (lane & 7) + 8 * (lane >> 4), cchunk = (lane >> 3) & 1;
    const int kw = g & 3, khg = g >> 2;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      mbar_wait(&full_bar[s], ph);
      const uint8_t* st = smem + s * stage_bytes;
      const uint32_t Sc = smem_u32(st);
      const TF* Sf = reinterpret_cast<const TF*>(st + tile_bytes);
      const TF* Sr = reinterpret_cast<const TF*>(st + tile_bytes + band_bytes);
      for (int rr = 0; rr < a.RT; ++rr) {
        uint32_t af[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          const uint32_t addr = Sc + sw64(rr * a.W + 16 * c + px_l, 2 * m + cchunk);
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(af[m][0]), "=r"(af[m][1]), "=r"(af[m][2]), "=r"(af[m][3]) : "r"(addr));
        }
#pragma unroll
        for (int ci = 0; ci < NC; ++ci)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            // pixels ow = 16c + 2t (+1, +8, +9) of this k-step at tap (kh = 2hf + khg, kw): iw = 2*ow - 1 + kw
            const int e0 = (ci * frows + 2 * rr + 2 * hf + khg) * pitch + 2 * (16 * c + 2 * t) + kw - 1 + PADL;
            float x[4] = {band_ld(Sf + e0), band_ld(Sf + e0 + 2), band_ld(Sf + e0 + 16), band_ld(Sf + e0 + 18)};
            if (HAS_REF) {
              const float r[4] = {band_ld(Sr + e0), band_ld(Sr + e0 + 2), band_ld(Sr + e0 + 16), band_ld(Sr + e0 + 18)};
#pragma unroll
              for (int e = 0; e < 4; ++e) x[e] *= act_grad_from_output(r[e], a.fine_act, a.slope);
            }
            const uint32_t b0 = pack_bf16x2(x[0], x[1]), b1 = pack_bf16x2(x[2], x[3]);
#pragma unroll
            for (int m = 0; m < 2; ++m) mma_bf16_16816(acc[m][ci * 2 + hf], af[m], b0, b1);
          }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
      if (++s == kThinStages) { s = 0; ph ^= 1; }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int j = 0; j < 2 * NC; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int co = 16 * m + g + 8 * (e >> 1), ci = j >> 1, tap = 8 * (j & 1) + 2 * t + (e & 1);
          atomicAdd(&red[(co * NC + ci) * 16 + tap], acc[m][j][e]);
        }
  }
  asm volatile("bar.sync 1, 224;" ::: "memory");          // the seven consumer warps (the producer warp has left)
  for (int i = threadIdx.x; i < 512 * NC; i += 224) atomicAdd(a.dw + i, red[i]);
}

bool dense_bf16_32(const b200gan_view* v) {
  return v->dtype == B200GAN_BF16 && v->c == 32 && v->sc == 1 && v->sw == 32 && v->sh == (int64_t)v->w * 32 &&
         v->sn == (int64_t)v->h * v->w * 32 && (reinterpret_cast<uintptr_t>(v->ptr) & 15) == 0;
}

// 4-wide loads along W: unit W stride, every row start a multiple of 4 elements, base pointer aligned to 4 elements
int vec4_ok(const b200gan_view* v) {
  const uintptr_t bytes = v->dtype == B200GAN_F32 ? 16 : 8;
  return v->sw == 1 && v->sn % 4 == 0 && v->sh % 4 == 0 && (v->c == 1 || v->sc % 4 == 0) && (reinterpret_cast<uintptr_t>(v->ptr) % bytes) == 0;
}

// 16-byte cp.async along W: unit W stride, every row start and the base pointer 16-byte aligned
bool vec16_ok(const b200gan_view* v) {
  const int per = v->dtype == B200GAN_F32 ? 4 : 8;
  return v->sw == 1 && v->sn % per == 0 && v->sh % per == 0 && (v->c == 1 || v->sc % per == 0) && v->w % per == 0 &&
         (reinterpret_cast<uintptr_t>(v->ptr) & 15) == 0;
}

// a 32-channel slice of a dense wider NHWC bf16 tensor (pixel pitch a multiple of 8 elements = 16 bytes), or the dense tensor itself
bool slice_bf16_32(const b200gan_view* v) {
  return v->dtype == B200GAN_BF16 && v->c == 32 && v->sc == 1 && v->sw >= 32 && v->sw % 8 == 0 && v->sh == (int64_t)v->w * v->sw &&
         v->sn == (int64_t)v->h * v->w * v->sw && (reinterpret_cast<uintptr_t>(v->ptr) & 15) == 0;
}

// returns false when the problem is not of the thin shape
bool thin_setup(ThinArgs* a, const b200gan_view* fine, const b200gan_view* fine_ref, int fine_act, const b200gan_view* coarse,
                const b200gan_view* coarse_ref, int coarse_act, float slope) {
  if (!slice_bf16_32(coarse) || (fine->c != 1 && fine->c != 3)) return false;
  if (coarse->sw != 32 && coarse_ref) return false;
  if (coarse->w % 16 != 0 || coarse->w > 128 || fine->h != 2 * coarse->h || fine->w != 2 * coarse->w || fine->n != coarse->n) return false;
  if (coarse_ref && (!dense_bf16_32(coarse_ref) || coarse_ref->n != coarse->n || coarse_ref->h != coarse->h || coarse_ref->w != coarse->w)) return false;
  if (fine_ref && (fine_ref->n != fine->n || fine_ref->h != fine->h || fine_ref->w != fine->w || fine_ref->c != fine->c)) return false;
  a->fine = to_view(fine);
  if (fine_ref) a->fine_ref = to_view(fine_ref); else a->fine_ref.ptr = nullptr;
  a->coarse = reinterpret_cast<const __nv_bfloat16*>(coarse->ptr);
  a->coarse_out = reinterpret_cast<__nv_bfloat16*>(coarse->ptr);
  a->coarse_ref = coarse_ref ? reinterpret_cast<const __nv_bfloat16*>(coarse_ref->ptr) : nullptr;
  a->N = coarse->n; a->H = coarse->h; a->W = coarse->w;
  a->cpitch = (int)coarse->sw;
  a->R = coarse->h < 8 ? coarse->h : 8;
  a->tiles_per_img = (a->H + a->R - 1) / a->R;
  a->num_tiles = a->N * a->tiles_per_img;
  a->fine_act = fine_act; a->coarse_act = coarse_act; a->out_act = B200GAN_ACT_NONE; a->slope = slope;
  a->fine_vec = vec4_ok(fine);
  a->ref_vec = fine_ref ? vec4_ok(fine_ref) : 0;
  return true;
}

// 4-d tensor map over the (N,H,W,32) bf16 coarse tensor (dense, or a channel slice: pixel pitch a.cpitch): box {32 ch, box_w, box_h, 1}, 64B swizzle,
// zero OOB fill
int coarse_tensor_map(CUtensorMap* m, const ThinArgs& a, int box_w, int box_h) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return B200GAN_ERR_CUDA; }
  cuuint64_t gdim[4] = {32, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.N};
  const cuuint64_t pix = (cuuint64_t)a.cpitch * 2;              // bytes between pixels
  cuuint64_t gstr[3] = {pix, (cuuint64_t)a.W * pix, (cuuint64_t)a.H * a.W * pix};
  cuuint32_t box[4] = {32, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(a.coarse), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(coarse) failed: %d", (int)r); return B200GAN_ERR_CUDA; }
  return 0;
}

bool thin_tma_ok(const ThinArgs& a) { return a.coarse_ref == nullptr && a.W % 16 == 0 && a.W / 16 <= 7 && a.H >= 2; }

template <void (*Kernel)(const CUtensorMap, const ThinArgs)>
int launch_thin_tma(const CUtensorMap& m, const ThinArgs& a, size_t smem, int ctas_per_sm, cudaStream_t st, const char* name) {
  B200_CUDA((ensure_dynamic_smem<Kernel>((int)smem)));
  int grid = ctas_per_sm * kNumSMs;
  if (grid > a.num_tiles) grid = a.num_tiles;
  Kernel<<<grid, 256, smem, st>>>(m, a);
  B200_LAUNCH_CHECK(name);
  return 0;
}

template <void (*Kernel)(const ThinArgs)>
int launch_thin(const ThinArgs& a, size_t smem, int ctas_per_sm, cudaStream_t st, const char* name, int threads = 256) {
  B200_CUDA((ensure_dynamic_smem<Kernel>((int)smem)));
  int grid = ctas_per_sm * kNumSMs;
  if (grid > a.num_tiles) grid = a.num_tiles;
  Kernel<<<grid, threads, smem, st>>>(a);
  B200_LAUNCH_CHECK(name);
  return 0;
}

}  // namespace

// Each wrapper returns 1 when the problem is not of its shape (the dispatcher then uses the generic kernels).

// fine (gathered; optionally multiplied by act'(fine_ref)) -> coarse = out_act(conv)
// epi: optional BatchNorm fusion on the result (TcEpi modes 0 none, 1 statistics, 2 activation backward + BN-backward sums)
int thin_down(const b200gan_view* fine, const b200gan_view* fine_ref, int fine_act, const float* w, const b200gan_view* coarse, int out_act,
              float slope, const TcEpi& epi, cudaStream_t st) {
  ThinArgs a{};
  if (epi.mode == 3) return 1;
  if (epi.mode == 2 && (!dense_bf16_32(epi.prev_y) || epi.prev_y->n != coarse->n || epi.prev_y->h != coarse->h || epi.prev_y->w != coarse->w)) return 1;
  if (out_act != B200GAN_ACT_NONE && out_act != B200GAN_ACT_LRELU && out_act != B200GAN_ACT_RELU) return 1;
  if (!thin_setup(&a, fine, fine_ref, fine_act, coarse, nullptr, B200GAN_ACT_NONE, slope)) return 1;
  if (a.cpitch != 32 && epi.mode != 0) return 1;                // a 32-channel slice as the result: plain / activation epilogue only
  a.w = w; a.out_act = out_act;
  a.epi = epi.mode;
  if (epi.mode != 0) {
    a.sums = epi.sums;
    B200_CUDA(cudaMemsetAsync(epi.sums, 0, sizeof(double) * 64, st));
    if (epi.mode == 2) {
      a.prev_y = reinterpret_cast<const __nv_bfloat16*>(epi.prev_y->ptr);
      a.prev_scale = epi.scale; a.prev_shift = epi.shift; a.prev_mean = epi.mean; a.prev_invstd = epi.invstd;
      a.prev_neg = epi.act == B200GAN_ACT_RELU ? 0.f : (epi.act == B200GAN_ACT_LRELU ? epi.slope : 1.f);
    }
  }
  // Four-warp CTAs on four-row tiles, twice as many of them: the kernel is a load -> barrier -> compute -> store loop per CTA and its 106-128 registers
  // allow 16 warps per SM either way; four small CTAs are more often in different phases than two large ones (same-box A/B at batch 512: first-layer
  // forward 174 -> 165 us, RGB input gradient with the BatchNorm-backward epilogue 445 -> 405 us, the nc = 1 variants -2 %).
  int threads = 256, mult = 1;
  if (a.H % 4 == 0) {
    a.R = 4; a.tiles_per_img = (a.H + a.R - 1) / a.R; a.num_tiles = a.N * a.tiles_per_img;
    threads = 128; mult = 2;
  }
  size_t smem = (size_t)fine->c * (2 * a.R + 2) * (2 * a.W + 2) * 2;
  if (epi.mode == 2) smem = ((smem + 127) & ~(size_t)127) + (size_t)a.R * a.W * 64;
  const char* nm = "thin_down_mma_kernel";
  if (fine->c == 1) {
    if (epi.mode == 0) return launch_thin<thin_down_mma_kernel<1, 0>>(a, smem, mult * 6, st, nm, threads);
    if (epi.mode == 1) return launch_thin<thin_down_mma_kernel<1, 1>>(a, smem, mult * 4, st, nm, threads);
    return launch_thin<thin_down_mma_kernel<1, 2>>(a, smem, mult * 4, st, nm, threads);
  }
  if (epi.mode == 0) return launch_thin<thin_down_mma_kernel<3, 0>>(a, smem, mult * 4, st, nm, threads);
  if (epi.mode == 1) return launch_thin<thin_down_mma_kernel<3, 1>>(a, smem, mult * 3, st, nm, threads);
  return launch_thin<thin_down_mma_kernel<3, 2>>(a, smem, mult * 3, st, nm, threads);
}

// coarse (gathered; optionally multiplied by act'(coarse_ref)) -> fine = out_act(transposed conv)
int thin_up(const b200gan_view* coarse, const b200gan_view* coarse_ref, int coarse_act, float slope, const float* w, const b200gan_view* fine,
            int out_act, cudaStream_t st) {
  ThinArgs a{};
  if (out_act != B200GAN_ACT_NONE && out_act != B200GAN_ACT_TANH) return 1;
  if (!thin_setup(&a, fine, nullptr, B200GAN_ACT_NONE, coarse, coarse_ref, coarse_act, slope)) return 1;
  a.w = w; a.out_act = out_act;
  const int NT = (4 * fine->c + 7) / 8;
  if (thin_tma_ok(a)) {
    a.RT = 4;
    a.tiles_per_img = (a.H + a.RT - 1) / a.RT;
    a.num_tiles = a.N * a.tiles_per_img;
    CUtensorMap m;
    int rc = coarse_tensor_map(&m, a, a.W + 2, a.RT + 2);
    if (rc) return rc;
    const size_t stage = ((size_t)(a.RT + 2) * (a.W + 2) * 64 + 1023) & ~(size_t)1023;
    const size_t smem_tma = kThinStages * stage + (size_t)18 * NT * 32 * 8 + 64 + 8 * 384 + 1024;      // + the per-warp RGB store scratch
    if (fine->c == 1) return launch_thin_tma<thin_up_tma_kernel<1>>(m, a, smem_tma, 2, st, "thin_up_tma_kernel");
    return launch_thin_tma<thin_up_tma_kernel<3>>(m, a, smem_tma, 2, st, "thin_up_tma_kernel");
  }
  if (a.cpitch != 32) return 1;                                 // the cp.async-staged kernel indexes the coarse tensor densely
  const size_t smem = (size_t)18 * NT * 32 * 8 + (size_t)(a.R + 2) * (a.W + 2) * CP * 2;
  if (fine->c == 1) return launch_thin<thin_up_mma_kernel<1>>(a, smem, 2, st, "thin_up_mma_kernel");
  return launch_thin<thin_up_mma_kernel<3>>(a, smem, 2, st, "thin_up_mma_kernel");
}

// dw (32, nc, 4, 4) fp32 += coarse^T x im2col(fine); either operand may carry the fused activation backward
int thin_wgrad(const b200gan_view* fine, const b200gan_view* fine_ref, int fine_act, const b200gan_view* coarse, const b200gan_view* coarse_ref,
               int coarse_act, float slope, float* dw, cudaStream_t st) {
  ThinArgs a{};
  if (!thin_setup(&a, fine, fine_ref, fine_act, coarse, coarse_ref, coarse_act, slope)) return 1;
  a.dw = dw;
  // raw cp.async staging of the image band: unit W stride, 16-byte aligned rows, reference (if any) of the same dtype
  const bool band_ok = a.fine_vec && vec16_ok(fine) && (!fine_ref || (fine_ref->dtype == fine->dtype && vec16_ok(fine_ref)));
  // nc = 3: the cp.async-staged kernel below with the row-per-warp image staging beats the TMA-staged one (fp32 real batch: 433 -> 268 us); its raw
  // image band is 3-6x as large and left one or two resident CTAs per SM
  if (thin_tma_ok(a) && band_ok && fine->c != 3) {
    // rows of the coarse tensor per tile: the raw image band beside it is 3x as large at nc = 3 (6x in fp32), and at 4 rows one stage pair took
    // 113-170 KB = ONE resident CTA per SM (ncu: 12 % of the warp slots, 0.19 of the copy bandwidth); 2 rows keep two CTAs resident
    a.RT = fine->c == 3 ? 2 : 4;
    a.tiles_per_img = (a.H + a.RT - 1) / a.RT;
    a.num_tiles = a.N * a.tiles_per_img;
    CUtensorMap m;
    int rc = coarse_tensor_map(&m, a, a.W, a.RT);
    if (rc) return rc;
    const int esz = fine->dtype == B200GAN_F32 ? 4 : 2, padl = 16 / esz, nb = fine_ref ? 2 : 1;
    const size_t band = (size_t)fine->c * (2 * a.RT + 2) * (2 * a.W + 2 * padl) * esz;
    const size_t stage = ((size_t)a.RT * a.W * 64 + nb * band + 1023) & ~(size_t)1023;
    const size_t smem_tma = kThinStages * stage + (size_t)512 * fine->c * 4 + 64 + 1024;
    const char* nm = "thin_wgrad_tma_kernel";
#define WG(NC, T, R) launch_thin_tma<thin_wgrad_tma_kernel<NC, T, R>>(m, a, smem_tma, 2, st, nm)
    if (fine->c == 1) {
      if (fine->dtype == B200GAN_F32) return fine_ref ? WG(1, float, true) : WG(1, float, false);
      return fine_ref ? WG(1, __nv_bfloat16, true) : WG(1, __nv_bfloat16, false);
    }
    if (fine->dtype == B200GAN_F32) return fine_ref ? WG(3, float, true) : WG(3, float, false);
    return fine_ref ? WG(3, __nv_bfloat16, true) : WG(3, __nv_bfloat16, false);
#undef WG
  }
  if (a.cpitch != 32) return 1;
  const size_t smem = (size_t)512 * fine->c * 4 + (size_t)a.R * a.W * CP * 2 + (size_t)fine->c * (2 * a.R + 2) * (2 * a.W + 2) * 2;
  if (fine->c == 1) return launch_thin<thin_wgrad_mma_kernel<1>>(a, smem, 2, st, "thin_wgrad_mma_kernel");
  return launch_thin<thin_wgrad_mma_kernel<3>>(a, smem, 2, st, "thin_wgrad_mma_kernel");
}

}  // namespace b200gan
