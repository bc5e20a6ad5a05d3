// Shared pieces of the tcgen05 translation units (conv_tc.cu: generic implicit GEMM, conv_tc_halo.cu: halo-tile kernels for the
// thin 32<->64-channel layers, conv_tc_wgrad.cu: weight gradients): UMMA descriptors, the warp column-sum butterfly, role layout.
#pragma once
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace b200gan {

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type [61,64) (2 = SWIZZLE_128B, 4 = SWIZZLE_64B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: D=f32, A=B=bf16
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// Column sums over the 32 lanes of a warp for 16 columns held as v[0..15] per lane: 16 shuffles (reduce-scatter butterfly)
// instead of 16 x 5.  On return v[0] of lanes j and j+16 is the sum of column j.
__device__ __forceinline__ void warp_column_sums(float (&v)[16], int lane) {
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
}

// warps 0..7 epilogue, warp 8 TMA producer, warp 9 MMA issuer (+ warp 10: TMA stores, halo-tile kernels)
constexpr int kTcThreads = 320;
constexpr int kHaloThreads = 352;          // halo-tile kernels: + one warp that owns the TMA stores of the output tiles

static inline int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return (1 << l) == v ? l : -1;
}

static inline bool nhwc_dense_bf16(const b200gan_view* v) {
  return v->dtype == B200GAN_BF16 && v->sc == 1 && v->sw == v->c && v->sh == (int64_t)v->w * v->c &&
         v->sn == (int64_t)v->h * v->w * v->c && (reinterpret_cast<uintptr_t>(v->ptr) & 31) == 0;      // 32 B: 256-bit epilogue accesses
}

// parameters of the generic implicit-GEMM kernels (conv_tc.cu: one CTA per tile; conv_tc_pair.cu: a CTA pair per 256-row tile)
struct TcConvParams {
  int tiles_w, tiles_h, tiles_n;     // tiles of the GEMM-row pixel space (TW x TH x TN pixels each, product 128)
  int tw_log2, th_log2;              // log2(TW), log2(TH)
  int a_mul;                         // input coordinate = tile origin * a_mul + tap offset
  int taps;                          // 16 (DOWN) or 4 (UP)
  int chunks;                        // Cin / KC
  int n_tiles, ncls, num_tiles;      // Cout tiles, parity classes (1 or 4), total tiles = spatial * n_tiles * ncls
  int8_t tap_dh[4][16], tap_dw[4][16];   // [class][tap]
  int QH, QW, NB;                    // valid extent of the pixel space (rows beyond are discarded)
  __nv_bfloat16* out;
  int64_t o_sn, o_sh, o_sw;
  int o_mul;                         // output pixel = q*o_mul + class parity
  int cout;
  // epilogue fusions (template EPI): 1 = BatchNorm statistics of the result, 2 = previous layer's activation backward +
  // BatchNorm-backward sums, 3 = previous layer's activation backward only (no BatchNorm below)  (b200gan_fuse.bn_sums / prev_*)
  double* sums;                      // [2*cout], zeroed by the host wrapper
  const __nv_bfloat16* prev_y;       // same dense NHWC layout as out
  const float *prev_scale, *prev_shift, *prev_mean, *prev_invstd;
  float prev_neg;                    // act'(z) for z <= 0: 0 (ReLU), slope (LeakyReLU), 1 (none)
  // shared-memory plan (bytes from the 1024-aligned base): [stages][resident weights][barriers][channel accumulators]
  int nstages, stage_stride, off_res, off_bar;
  int resident;                      // 1: every weight tile of the layer stays in shared memory for the CTA's lifetime
};

// CTA-pair kernel (conv_tc_pair.cu): 256 x 256 tiles on tcgen05.mma.cta_group::2.  `mb_half` is the weight map with a 128-row box.
int launch_tc_pair(const CUtensorMap& ma, const CUtensorMap& mb_half, TcConvParams p, int epi, cudaStream_t st);

// halo-tile kernels (conv_tc_halo.cu): return 1 when the problem is not of their shape / fusion
int tc_conv_up4(const b200gan_view* in, const void* wpacked, const b200gan_view* out, const TcEpi& epi, cudaStream_t st);
int tc_conv_down4(const b200gan_view* in, const void* wpacked, const b200gan_view* out, const TcEpi& epi, cudaStream_t st);
// conv_tc_halo_wide.cu: the 128 -> 64 channel "up" layer
int tc_conv_up4w(const b200gan_view* in, const void* wpacked, const b200gan_view* out, const TcEpi& epi, cudaStream_t st);

}  // namespace b200gan
