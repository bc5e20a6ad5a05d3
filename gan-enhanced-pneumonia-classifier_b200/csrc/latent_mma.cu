// Latent projection G0 = ConvTranspose2d(nz -> C, k7 s1 p0) on a 1x1 input (reference: src/dcgan.py:26) as warp-level
// tensor-core GEMMs (mma.sync m16n8k16, bf16 x bf16 -> fp32):
//   forward          y[n][hw][co]  = sum_k z[n][k] * w[k][co][hw]            (M = batch, K = nz, N = C*49)
//   weight gradient  dw[k][co][hw] += sum_n z[n][k] * dy[n][hw][co]          (M = nz,    K = batch)
// 2.6 GFLOP against 36 MB of operands: the work is data movement, and the only awkward part is that the weight is (co, hw)-major
// while the activation is (hw, co)-major.  Both kernels therefore give a CTA the columns {8 channels} x {49 positions}: in the
// WEIGHT these are 392 contiguous floats per latent index (coalesced float4 reads / read-modify-writes), in the ACTIVATION 49
// pieces of 16 bytes per sample; the (co, hw) <-> (hw, co) transposition happens in shared memory.  z (any dtype / strides) and
// the fp32 weight are rounded to bf16 on the way into shared memory like every other tensor-core operand of the step.
#include "common.cuh"
#include "ptx.cuh"

namespace b200gan {

namespace {

constexpr int kHW = 49;                 // 7 x 7 output positions
constexpr int kCB = 8;                  // channels per CTA
constexpr int kCols = kHW * kCB;        // 392 GEMM columns per CTA = 49 n8 tiles = 7 warps x 7 tiles
constexpr int kLatThreads = 224;        // 7 warps
constexpr int kZP = 8;                  // pad of the staged z rows (bf16 elements): conflict-free ldmatrix for KP in {16..128}

__device__ __forceinline__ void lat_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void lat_ldsm4(uint32_t (&r)[4], const void* p) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void lat_ldsm4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void lat_ldsm2_t(uint32_t (&r)[2], const void* p) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}

// B fragments (k16 x n8, ".col") of this warp's seven n8 tiles from a tile stored [k][col] (col contiguous), rows k0..k0+15
__device__ __forceinline__ void load_b7(uint32_t (&b)[7][2], const __nv_bfloat16* Bs, int pitch, int k0, int col0, int lane) {
  const __nv_bfloat16* base = Bs + (k0 + (lane & 7) + 8 * ((lane >> 3) & 1)) * pitch + col0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    uint32_t r[4];
    lat_ldsm4_t(r, base + 16 * j + 8 * (lane >> 4));
    b[2 * j][0] = r[0]; b[2 * j][1] = r[1]; b[2 * j + 1][0] = r[2]; b[2 * j + 1][1] = r[3];
  }
  uint32_t r2[2];
  lat_ldsm2_t(r2, base + 48);
  b[6][0] = r2[0]; b[6][1] = r2[1];
}

// Stage z rows [n0, n0+64) x latent indices [kbase, kbase+KP) as bf16 Zs[row][KP + kZP], zero outside (n >= N, k >= NZ);
// eight loads in flight per thread (one at a time this loop took longer than the GEMM)
template <typename TZ>
__device__ __forceinline__ void stage_z(__nv_bfloat16* Zs, const View& z, int n0, int N, int NZ, int kbase, int KP) {
  const int zp = KP + kZP, total = 64 * KP;
  constexpr int U = 8;
  for (int base = threadIdx.x; base < total; base += U * kLatThreads) {
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * kLatThreads, r = i / KP, k = kbase + (i - r * KP), n = n0 + r;
      v[u] = (i < total && n < N && k < NZ) ? ld_as_float(reinterpret_cast<const TZ*>(z.ptr) + (int64_t)n * z.sn + (int64_t)k * z.sc) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * kLatThreads, r = i / KP;
      if (i < total) Zs[r * zp + (i - r * KP)] = __float2bfloat16_rn(v[u]);
    }
  }
}

// ---- forward: CTA = 64 samples x (8 channels x 49 positions), all of K resident -----------------------------------------
template <typename TZ>
__global__ void __launch_bounds__(kLatThreads, 1)
latent_fprop_mma_kernel(View z, const float* __restrict__ w, __nv_bfloat16* __restrict__ y, int N, int NZ, int KP, int C) {
  extern __shared__ __align__(16) uint8_t lat_smem[];
  __nv_bfloat16* Bs = reinterpret_cast<__nv_bfloat16*>(lat_smem);          // [KP][392]   (co_l * 49 + hw)
  __nv_bfloat16* Zs = Bs + KP * kCols;                                      // [64][KP + 8]
  const int n0 = blockIdx.x * 64, co0 = blockIdx.y * kCB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int zp = KP + kZP;
  // weights: rows k of 392 contiguous floats; eight float4 loads in flight per thread (the load phase is the kernel)
  {
    constexpr int U = 8, R4 = kCols / 4;
    const int total = KP * R4;
    for (int base = threadIdx.x; base < total; base += U * kLatThreads) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * kLatThreads, k = i / R4, c4 = i - k * R4;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < total && k < NZ) v[u] = __ldg(reinterpret_cast<const float4*>(w + ((int64_t)k * C + co0) * kHW) + c4);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * kLatThreads;
        if (i < total) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(v[u].x, v[u].y), hi = __floats2bfloat162_rn(v[u].z, v[u].w);
          *reinterpret_cast<uint2*>(Bs + 4 * i) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        }
      }
    }
  }
  stage_z<TZ>(Zs, z, n0, N, NZ, 0, KP);
  __syncthreads();
  float acc[4][7][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int j = 0; j < 7; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[m][j][e] = 0.f;
  const int col0 = warp * 56;
  for (int k0 = 0; k0 < KP; k0 += 16) {
    uint32_t b[7][2];
    load_b7(b, Bs, kCols, k0, col0, lane);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      uint32_t a[4];
      lat_ldsm4(a, Zs + (16 * m + (lane & 7) + 8 * ((lane >> 3) & 1)) * zp + k0 + 8 * (lane >> 4));
#pragma unroll
      for (int j = 0; j < 7; ++j) lat_mma(acc[m][j], a, b[j][0], b[j][1]);
    }
  }
  __syncthreads();
  // transpose through shared memory: Os[row][hw * 8 + co_l] so that every (sample, position) is one 16-byte store
  __nv_bfloat16* Os = Bs;                                                   // 64 x 392 bf16 = 50 KB <= KP x 392 x 2 for KP >= 64
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int j = 0; j < 7; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = 16 * m + g + 8 * (e >> 1), col = col0 + 8 * j + 2 * t + (e & 1);
        const int co_l = col / kHW, hw = col - co_l * kHW;
        Os[row * kCols + hw * kCB + co_l] = __float2bfloat16_rn(acc[m][j][e]);
      }
  __syncthreads();
  const int64_t NC = (int64_t)kHW * C;
  for (int i = threadIdx.x; i < 64 * kHW; i += blockDim.x) {
    const int row = i / kHW, hw = i - row * kHW;
    if (n0 + row < N)
      *reinterpret_cast<uint4*>(y + (int64_t)(n0 + row) * NC + (int64_t)hw * C + co0) = *reinterpret_cast<const uint4*>(Os + row * kCols + hw * kCB);
  }
}

// ---- weight gradient: CTA = 32 latent indices x (8 channels x 49 positions), batch walked in chunks of 64 samples ----------
template <typename TZ>
__global__ void __launch_bounds__(kLatThreads, 2)
latent_wgrad_mma_kernel(View z, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw, int N, int NZ, int C) {
  extern __shared__ __align__(16) uint8_t lat_smem[];
  constexpr int KP = 32, zp = KP + kZP;
  __nv_bfloat16* Ds = reinterpret_cast<__nv_bfloat16*>(lat_smem);          // [64 samples][392]   (hw * 8 + co_l)
  __nv_bfloat16* Zs = Ds + 64 * kCols;                                      // [64 samples][32 + 8] latent indices k0..k0+31
  const int co0 = blockIdx.x * kCB, kbase = blockIdx.y * KP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t NC = (int64_t)kHW * C;
  float acc[2][7][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < 7; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[m][j][e] = 0.f;
  const int col0 = warp * 56;
  for (int n0 = 0; n0 < N; n0 += 64) {
    // gradient chunk: 64 x 49 pieces of 16 bytes (8 channels) = 14 per thread, seven in flight
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint4 v[7];
#pragma unroll
      for (int u = 0; u < 7; ++u) {
        const int i = threadIdx.x + (half * 7 + u) * kLatThreads, r = i / kHW, hw = i - r * kHW, n = n0 + r;
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        if (n < N) v[u] = __ldg(reinterpret_cast<const uint4*>(dy + (int64_t)n * NC + (int64_t)hw * C + co0));
      }
#pragma unroll
      for (int u = 0; u < 7; ++u) {
        const int i = threadIdx.x + (half * 7 + u) * kLatThreads;
        *reinterpret_cast<uint4*>(Ds + i * kCB) = v[u];
      }
    }
    stage_z<TZ>(Zs, z, n0, N, NZ, kbase, KP);
    __syncthreads();
#pragma unroll
    for (int s0 = 0; s0 < 64; s0 += 16) {
      uint32_t b[7][2];
      load_b7(b, Ds, kCols, s0, col0, lane);
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        // A = z^T: stored [sample][k]; transposed 8x8 loads give the row-major (k x sample) fragments
        uint32_t a[4];
        lat_ldsm4_t(a, Zs + (s0 + (lane & 7) + 8 * (lane >> 4)) * zp + 16 * m + 8 * ((lane >> 3) & 1));
#pragma unroll
        for (int j = 0; j < 7; ++j) lat_mma(acc[m][j], a, b[j][0], b[j][1]);
      }
    }
    __syncthreads();
  }
  // transpose to the weight's (co, hw) order through shared memory, then coalesced read-modify-write of 392-float rows
  float* Os = reinterpret_cast<float*>(lat_smem);                           // [32][392] fp32 = 50 KB = the Ds bytes
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < 7; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = 16 * m + g + 8 * (e >> 1), col = col0 + 8 * j + 2 * t + (e & 1);
        const int hw = col >> 3, co_l = col & 7;
        Os[row * kCols + co_l * kHW + hw] = acc[m][j][e];
      }
  __syncthreads();
  for (int i = threadIdx.x; i < KP * (kCols / 4); i += blockDim.x) {
    const int r = i / (kCols / 4), c4 = i - r * (kCols / 4), k = kbase + r;
    if (k < NZ) {
      float4* dst = reinterpret_cast<float4*>(dw + ((int64_t)k * C + co0) * kHW) + c4;
      float4 o = *dst;
      const float4 a = *reinterpret_cast<const float4*>(Os + r * kCols + 4 * c4);
      o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
      *dst = o;
    }
  }
}

bool latent_mma_shape(const b200gan_conv* cv, const b200gan_view* fine, const b200gan_view* z, const void* w) {
  return cv->k == 7 && cv->stride == 1 && cv->pad == 0 && z->h == 1 && z->w == 1 && fine->h == 7 && fine->w == 7 && fine->c % kCB == 0 &&
         fine->dtype == B200GAN_BF16 && fine->sc == 1 && fine->sw == fine->c && fine->sh == 7 * (int64_t)fine->c && fine->sn == 49 * (int64_t)fine->c &&
         (reinterpret_cast<uintptr_t>(fine->ptr) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 && z->c >= 1;
}

}  // namespace

// returns 1 when the shape is not the 7x7 latent projection these kernels are written for (the caller falls back to SIMT)
int latent_fprop_mma(const b200gan_conv* cv, const b200gan_view* z, const float* w, const b200gan_view* y, cudaStream_t st) {
  if (!latent_mma_shape(cv, y, z, w) || z->c > 128) return 1;
  int KP = (z->c + 15) & ~15;
  if (KP < 64) KP = 64;                                  // the output staging tile (64 x 392 bf16) reuses the weight tile's bytes
  const int smem = KP * kCols * 2 + 64 * (KP + kZP) * 2;
  dim3 grid((z->n + 63) / 64, y->c / kCB);
  const View zv = to_view(z);
  const int which = z->dtype == B200GAN_F32 ? 0 : 1;
  {
    if (which == 0) B200_CUDA((ensure_dynamic_smem<latent_fprop_mma_kernel<float>>(smem)));
    else B200_CUDA((ensure_dynamic_smem<latent_fprop_mma_kernel<__nv_bfloat16>>(smem)));
  }
  if (which == 0)
    latent_fprop_mma_kernel<float><<<grid, kLatThreads, smem, st>>>(zv, w, reinterpret_cast<__nv_bfloat16*>(y->ptr), z->n, z->c, KP, y->c);
  else
    latent_fprop_mma_kernel<__nv_bfloat16><<<grid, kLatThreads, smem, st>>>(zv, w, reinterpret_cast<__nv_bfloat16*>(y->ptr), z->n, z->c, KP, y->c);
  B200_LAUNCH_CHECK("latent_fprop_mma_kernel");
  return 0;
}

int latent_wgrad_mma(const b200gan_conv* cv, const b200gan_view* dy, const b200gan_view* z, float* dw, cudaStream_t st) {
  if (!latent_mma_shape(cv, dy, z, dw)) return 1;
  const int smem = 64 * kCols * 2 + 64 * (32 + kZP) * 2;
  dim3 grid(dy->c / kCB, (z->c + 31) / 32);
  const View zv = to_view(z);
  const int which = z->dtype == B200GAN_F32 ? 0 : 1;
  {
    if (which == 0) B200_CUDA((ensure_dynamic_smem<latent_wgrad_mma_kernel<float>>(smem)));
    else B200_CUDA((ensure_dynamic_smem<latent_wgrad_mma_kernel<__nv_bfloat16>>(smem)));
  }
  if (which == 0)
    latent_wgrad_mma_kernel<float><<<grid, kLatThreads, smem, st>>>(zv, reinterpret_cast<const __nv_bfloat16*>(dy->ptr), dw, z->n, z->c, dy->c);
  else
    latent_wgrad_mma_kernel<__nv_bfloat16><<<grid, kLatThreads, smem, st>>>(zv, reinterpret_cast<const __nv_bfloat16*>(dy->ptr), dw, z->n, z->c, dy->c);
  B200_LAUNCH_CHECK("latent_wgrad_mma_kernel");
  return 0;
}

}  // namespace b200gan
