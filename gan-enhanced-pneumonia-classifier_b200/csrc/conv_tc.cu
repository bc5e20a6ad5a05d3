// tcgen05 / TMEM / TMA implicit-GEMM convolutions for sm_100a (B200): the tensor-core path of the k4 s2 p1
// layers of the DCGAN (dcgan.py:30-46 ConvTranspose2d, :68-80 Conv2d) and of their input gradients.
//
// One kernel, two geometries (all tensors NHWC bf16, fp32 accumulation in TMEM):
//   DOWN  (Conv2d fprop, ConvTranspose2d dgrad):  out[n,oh,ow,:] = sum_{kh,kw} in[n,2oh-1+kh,2ow-1+kw,:] . W[:, (kh,kw,:)]
//         GEMM  M = N*OH*OW, K = 16*Cin, N = Cout
//   UP    (ConvTranspose2d fprop, Conv2d dgrad), one GEMM per output parity class (py,px) (blockIdx.z):
//         out[n,2q+py,2r+px,:] = sum_{jh,jw in {0,1}} in[n,q+ch-jh,r+cw-jw,:] . Wc[:, (jh,jw,:)],  ch=(py+1)/2
//         GEMM  M = N*H*W, K = 4*Cin, N = Cout  -- no zero-insertion, no multiply-by-zero work.
// A tile (128 GEMM rows x KC channels) is ONE 4-d TMA box {KC, TW*s, TH*s, TN} with element strides {1,s,s,1}
// over the NHWC tensor: TN images x TH rows x TW columns of output pixels; convolution padding is the TMA
// out-of-bounds zero fill (negative / past-the-end coordinates), so no im2col buffer and no padded copies exist.
// The weight tile (BN x KC, K-major, pre-packed bf16) is one 3-d TMA box.  Both land in 128B- (KC=64) or
// 64B- (KC=32) swizzled shared memory and are consumed by tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN,
// K=16) issued by one thread; a STAGES-deep mbarrier ring decouples TMA from MMA; four epilogue warps read
// the accumulator with tcgen05.ld and store bf16 NHWC rows.
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

// ---------------------------------------------------------------------------------------------------
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

template <int BN, int KC, int STAGES>
struct TcSmem {
  static constexpr int A_BYTES = 128 * KC * 2;
  static constexpr int B_BYTES = BN * KC * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 512;        // full[<=16] + empty[<=16] + acc_full[2] + acc_empty[2] + res + tmem slot
};


// Persistent kernel: each CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... of the (M-tile, N-tile, class)
// space.  Warps 0..7 = epilogue (TMEM lane quarter = warp % 4, column half = warp / 4: eight warps keep enough global loads /
// stores in flight for the fused epilogues), warp 8 = TMA producer, warp 9 = MMA issuer + TMEM owner.  The two single-thread
// roles sit in the highest warp slots so that an epilogue warp's TMEM lane quarter and column half are warp % 4 and warp / 4
// (roles in warps 0/1 instead measured the same within noise).  The accumulator is
// multi-buffered in TMEM (2-4 x BN columns) so the epilogue of tile i overlaps the MMA stream of tile i+1, and the
// smem ring keeps running across tile boundaries.  No integer division sits on the per-k-block path of the two
// single-thread roles (a first version spent ~100 instructions per step there).
constexpr int kTcThreads = 320;
constexpr int kHaloThreads = 352;          // halo-tile kernels: + one warp that owns the TMA stores of the output tiles

template <int BN, int KC, int STAGES, int EPI>
__global__ void __launch_bounds__(kTcThreads, 2)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcConvParams p) {
  using S = TcSmem<BN, KC, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int NST = p.nstages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* empty_bar = full_bar + 16;
  // accumulator ring in TMEM: 4 buffers for the narrow tiles (their MMA time per tile is shorter than the epilogue's latency
  // chain: barrier wake-up, tcgen05.ld, global stores), 2 for BN >= 128; two CTAs per SM share the 512 columns
  constexpr int NACC = BN <= 64 ? 4 : 2;
  uint64_t* acc_full = empty_bar + 16;         // [NACC]
  uint64_t* acc_empty = acc_full + 4;          // [NACC]
  uint64_t* res_bar = acc_empty + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);
  uint8_t* smem_res = smem + p.off_res;
  // EPI != 0: per-CTA channel accumulators [2][cout] (flushed once at the end), EPI == 2: {scale, shift, mean, invstd}[cout]
  float* ch_acc = reinterpret_cast<float*>(smem + p.off_bar + S::BAR_BYTES);
  float4* ch_coef = reinterpret_cast<float4*>(ch_acc + 2 * p.cout);
  if (EPI == 1 || EPI == 2) {
    for (int c = threadIdx.x; c < 2 * p.cout; c += blockDim.x) ch_acc[c] = 0.f;
    if (EPI == 2)
      for (int c = threadIdx.x; c < p.cout; c += blockDim.x)
        ch_coef[c] = make_float4(p.prev_scale[c], p.prev_shift[c], p.prev_mean[c], p.prev_invstd[c]);
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int TW = 1 << p.tw_log2, TH = 1 << p.th_log2, TN = 128 >> (p.tw_log2 + p.th_log2);
  constexpr uint32_t TMEM_COLS = NACC * BN < 32 ? 32 : NACC * BN;
  const int num_tiles = p.num_tiles;

  constexpr int kTmaWarp = 8, kMmaWarp = 9;
  if (warp == kTmaWarp && lane == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < NACC; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8); }
    mbar_init(res_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == kMmaWarp) {   // TMEM allocation by one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kTmaWarp) {
    // ===== TMA producer (one lane) =====
    if (lane == 0) {
      if (p.resident) {
        // all weight tiles [class][tap][chunk] once (n_tiles == 1): afterwards only activations stream through the ring
        const int per_cls = p.taps * p.chunks;
        mbar_expect_tx(res_bar, (uint32_t)(p.ncls * per_cls * S::B_BYTES));
        for (int cls = 0; cls < p.ncls; ++cls)
          for (int kb = 0; kb < per_cls; ++kb) tma_load_3d(smem_res + (cls * per_cls + kb) * S::B_BYTES, &map_b, res_bar, kb * KC, 0, cls);
      }
      const uint32_t stage_tx = p.resident ? S::A_BYTES : S::STAGE_BYTES;
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int r = t;
        const int cls = r % p.ncls; r /= p.ncls;
        const int nt = r % p.n_tiles; r /= p.n_tiles;
        const int tw_i = r % p.tiles_w; r /= p.tiles_w;
        const int th_i = r % p.tiles_h;
        const int tn_i = r / p.tiles_h;
        const int w0 = tw_i * TW * p.a_mul, h0 = th_i * TH * p.a_mul, n0 = tn_i * TN, cout0 = nt * BN;
        int kcol = 0;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int cw = w0 + p.tap_dw[cls][tap], chh = h0 + p.tap_dh[cls][tap];
          for (int chunk = 0; chunk < p.chunks; ++chunk, kcol += KC) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* sa = smem + s * p.stage_stride;
            mbar_expect_tx(&full_bar[s], stage_tx);
            tma_load_4d(sa, &map_a, &full_bar[s], chunk * KC, cw, chh, n0);
            if (!p.resident) tma_load_3d(sa + S::A_BYTES, &map_b, &full_bar[s], kcol, cout0, cls);
            if (++s == NST) { s = 0; ph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer (one lane) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
      constexpr uint32_t LT = KC == 64 ? 2u : 4u;          // SWIZZLE_128B : SWIZZLE_64B
      constexpr uint32_t SBO = 8 * KC * 2;                 // 8 rows of KC bf16
      const int num_kb = p.taps * p.chunks;
      int s = 0;
      uint32_t ph = 0;
      int lt = 0;
      if (p.resident) mbar_wait(res_bar, 0);
      const uint64_t adesc0 = make_smem_desc(smem_u32(smem), 16, SBO, LT);
      const uint64_t bres0 = make_smem_desc(smem_u32(smem_res), 16, SBO, LT);
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
        const int buf = lt % NACC;
        mbar_wait(&acc_empty[buf], ((lt / NACC) & 1) ^ 1);   // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        const uint32_t res_tile = (uint32_t)((t % p.ncls) * num_kb * S::B_BYTES);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          // base descriptors + 16-byte-unit offsets: two 64-bit adds per MMA instead of rebuilding both descriptors (the
          // single issuing thread is the pacing resource for thin k-blocks)
          const uint64_t adesc = adesc0 + (uint64_t)((uint32_t)(s * p.stage_stride) >> 4);
          const uint64_t bdesc = p.resident ? bres0 + (uint64_t)((res_tile + (uint32_t)(kb * S::B_BYTES)) >> 4)
                                            : adesc + (uint64_t)(S::A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) tcgen05_mma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          tcgen05_commit(&empty_bar[s]);                   // frees the smem stage when these MMAs retire
          if (++s == NST) { s = 0; ph ^= 1; }
        }
        tcgen05_commit(&acc_full[buf]);                    // accumulator complete
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: warps 0..7 =====
    const int q = warp & 3, hcol = warp >> 2;
    constexpr int CW = BN / 2;                       // columns per warp
    const int row = q * 32 + lane;
    const int tw = row & (TW - 1), th = (row >> p.tw_log2) & (TH - 1), tn = row >> (p.tw_log2 + p.th_log2);
    int lt = 0;
    // (valid, output offset, first column) of this thread's row piece in tile t
    auto locate = [&](int t, bool& valid, int& cout0) -> int64_t {
      int r = t;
      const int cls = r % p.ncls; r /= p.ncls;
      const int nt = r % p.n_tiles; r /= p.n_tiles;
      const int tw_i = r % p.tiles_w; r /= p.tiles_w;
      const int th_i = r % p.tiles_h;
      const int tn_i = r / p.tiles_h;
      const int ow = tw_i * TW + tw, oh = th_i * TH + th, n = tn_i * TN + tn;
      cout0 = nt * BN + hcol * CW;
      valid = ow < p.QW && oh < p.QH && n < p.NB && t < num_tiles;
      const int py = cls >> 1, px = cls & 1;
      return (int64_t)n * p.o_sn + (int64_t)(oh * p.o_mul + (p.o_mul > 1 ? py : 0)) * p.o_sh +
             (int64_t)(ow * p.o_mul + (p.o_mul > 1 ? px : 0)) * p.o_sw + cout0;
    };
    // Narrow tiles (<= 32 columns per warp): the whole row piece of y_prev for the NEXT tile is requested while this tile is
    // processed.  Requested per 16-column chunk inside the tile, every chunk exposed one HBM latency (~1.5 us) to the eight
    // epilogue warps, longer than the MMA stream of a whole tile.
    constexpr bool PRE = EPI >= 2 && CW <= 32;
    constexpr int NPRE = PRE ? CW / 8 : 1;
    uint4 ypre[NPRE];
    auto prefetch = [&](int t) {
      bool v2; int c2;
      const int64_t off = locate(t, v2, c2);
#pragma unroll
      for (int j = 0; j < NPRE; ++j) ypre[j] = make_uint4(0u, 0u, 0u, 0u);
      if (v2) {
#pragma unroll
        for (int j = 0; j < NPRE; j += 2) ldg256_nc(reinterpret_cast<const uint4*>(p.prev_y + off) + j, ypre[j], ypre[j + 1]);
      }
    };
    if (PRE) prefetch(blockIdx.x);
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
      bool valid; int cout0;
      const int64_t ooff = locate(t, valid, cout0);
      __nv_bfloat16* orow = p.out + ooff;
      const uint4* yp = reinterpret_cast<const uint4*>(p.prev_y + ooff);
      uint4 ynext[2], yall[NPRE];
      if (PRE) {
#pragma unroll
        for (int j = 0; j < NPRE; ++j) yall[j] = ypre[j];
        prefetch(t + gridDim.x);
      } else if (EPI >= 2) {                         // the first piece of y_prev is requested before the accumulator is waited for
        ynext[0] = ynext[1] = make_uint4(0u, 0u, 0u, 0u);
        if (valid) ldg256_nc(yp, ynext[0], ynext[1]);
      }
      const int buf = lt % NACC;
      mbar_wait(&acc_full[buf], (lt / NACC) & 1);
      tcgen05_fence_after();
#pragma unroll (PRE ? 2 : 1)
      for (int c0 = 0; c0 < CW; c0 += 16) {
        uint32_t v[16];
        tcgen05_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + hcol * CW + c0, v);
        float s0[16], s1[16];
        uint4 ycur[2];
        if (PRE) {
          ycur[0] = yall[(c0 / 8) % NPRE]; ycur[1] = yall[(c0 / 8 + 1) % NPRE];
        } else if (EPI >= 2) {
          ycur[0] = ynext[0]; ycur[1] = ynext[1];
          if (c0 + 16 < CW) {
            ynext[0] = ynext[1] = make_uint4(0u, 0u, 0u, 0u);
            if (valid) ldg256_nc(yp + (c0 + 16) / 8, ynext[0], ynext[1]);
          }
        }
        tcgen05_wait_ld();
        if (EPI == 2) {
          // dz = dx * act'(scale*y_prev + shift); sums of dz and dz*(y_prev - mean) (x invstd at the flush)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            float yv[8];
            unpack8(ycur[j], yv);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 cf = ch_coef[cout0 + c0 + 8 * j + e];
              const float z = fmaf(yv[e], cf.x, cf.y);
              v[8 * j + e] = __float_as_uint(__uint_as_float(v[8 * j + e]) * (z > 0.f ? 1.f : p.prev_neg));
              s1[8 * j + e] = yv[e] - cf.z;
            }
          }
        }
        if (EPI == 3) {
          // no BatchNorm below: y_prev is the saved activation output, dz = dx * act'(.) from its sign
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint32_t w[4] = {ycur[j].x, ycur[j].y, ycur[j].z, ycur[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xffff0000u);
              v[8 * j + 2 * e] = __float_as_uint(__uint_as_float(v[8 * j + 2 * e]) * (lo > 0.f ? 1.f : p.prev_neg));
              v[8 * j + 2 * e + 1] = __float_as_uint(__uint_as_float(v[8 * j + 2 * e + 1]) * (hi > 0.f ? 1.f : p.prev_neg));
            }
          }
        }
        // round to the stored precision; the statistics are those of the stored tensor
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          pk[j] = *reinterpret_cast<uint32_t*>(&b);
        }
        if (valid) stg256(orow + c0, make_uint4(pk[0], pk[1], pk[2], pk[3]), make_uint4(pk[4], pk[5], pk[6], pk[7]));
        if (EPI == 1 || EPI == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float lo = valid ? __uint_as_float(pk[j] << 16) : 0.f, hi = valid ? __uint_as_float(pk[j] & 0xffff0000u) : 0.f;
            s0[2 * j] = lo; s0[2 * j + 1] = hi;
            if (EPI == 1) { s1[2 * j] = lo * lo; s1[2 * j + 1] = hi * hi; }
            else { s1[2 * j] *= lo; s1[2 * j + 1] *= hi; }
          }
          warp_column_sums(s0, lane);
          warp_column_sums(s1, lane);
          if (lane < 16) {
            atomicAdd(&ch_acc[cout0 + c0 + lane], s0[0]);
            atomicAdd(&ch_acc[p.cout + cout0 + c0 + lane], s1[0]);
          }
        }
      }
      // all TMEM reads of this warp are complete (wait::ld above): hand the accumulator back to the MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    if (EPI == 1 || EPI == 2) {
      // the eight epilogue warps (256 threads) flush the CTA's channel sums: one double atomic per channel and quantity
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int c = threadIdx.x; c < p.cout; c += 256) {
        const float a0 = ch_acc[c], a1 = ch_acc[p.cout + c];
        if (a0 != 0.f) atomicAdd(p.sums + c, (double)a0);
        if (a1 != 0.f) atomicAdd(p.sums + p.cout + c, (double)a1 * (EPI == 2 ? (double)ch_coef[c].w : 1.0));
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return (1 << l) == v ? l : -1;
}

// choose TW x TH x TN = 128 output pixels per tile: powers of two dividing the spatial extent
static void pick_tile(int qh, int qw, int* tw, int* th) {
  int w = 1;
  while (w * 2 <= 16 && qw % (w * 2) == 0) w *= 2;
  int h = 1;
  while (h * 2 * w <= 128 && h * 2 <= 16 && qh % (h * 2) == 0) h *= 2;
  *tw = w; *th = h;
}

static bool nhwc_dense_bf16(const b200gan_view* v) {
  return v->dtype == B200GAN_BF16 && v->sc == 1 && v->sw == v->c && v->sh == (int64_t)v->w * v->c &&
         v->sn == (int64_t)v->h * v->w * v->c && (reinterpret_cast<uintptr_t>(v->ptr) & 31) == 0;      // 32 B: 256-bit epilogue accesses
}

template <int BN, int KC, int STAGES, int EPI>
static int launch_tc_epi(const CUtensorMap& ma, const CUtensorMap& mb, TcConvParams p, dim3 grid, cudaStream_t st) {
  using S = TcSmem<BN, KC, STAGES>;
  // shared-memory plan.  Streaming weights: STAGES stages of (A + B), two CTAs per SM.  Resident weights (thin-K layers whose
  // whole weight tensor is <= 64 KB): the weights once + a deeper ring of A-only stages, one CTA per SM.
  int res_bytes = p.resident ? p.ncls * p.taps * p.chunks * S::B_BYTES : 0;
  // two CTAs per SM must remain possible (one CTA per SM measured 30-100 % slower: a single TMA/MMA issue chain per SM does
  // not keep the L2 request pipeline full), so resident mode needs >= 4 A-only stages next to the weights within ~110 KB
  const int res_stages = (110 * 1024 - res_bytes) / S::A_BYTES;
  if (p.resident && res_stages < 4) { p.resident = 0; res_bytes = 0; }
  if (p.resident) {
    p.stage_stride = S::A_BYTES;
    p.nstages = res_stages > 16 ? 16 : res_stages;
  } else {
    p.stage_stride = S::STAGE_BYTES;
    p.nstages = STAGES;
  }
  p.off_res = (p.nstages * p.stage_stride + 1023) & ~1023;
  p.off_bar = p.off_res + res_bytes;
  // channel accumulators / coefficients live behind the barrier block (EPI 1/2), sized by the layer's channel count
  const int smem = 1024 + p.off_bar + S::BAR_BYTES + ((EPI == 0 || EPI == 3) ? 0 : p.cout * 8 + (EPI == 2 ? p.cout * 16 : 0) + 16);
  static int configured = 0;
  if (configured < smem) {
    B200_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, KC, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  conv_gemm_tc_kernel<BN, KC, STAGES, EPI><<<grid, kTcThreads, smem, st>>>(ma, mb, p);
  B200_LAUNCH_CHECK("conv_gemm_tc_kernel");
  return 0;
}

template <int BN, int KC, int STAGES>
static int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const TcConvParams& p, int epi, dim3 grid, cudaStream_t st) {
  static_assert(STAGES <= 16, "barrier block holds 16 stages");
  if (epi == 1) return launch_tc_epi<BN, KC, STAGES, 1>(ma, mb, p, grid, st);
  if (epi == 2) return launch_tc_epi<BN, KC, STAGES, 2>(ma, mb, p, grid, st);
  if (epi == 3) return launch_tc_epi<BN, KC, STAGES, 3>(ma, mb, p, grid, st);
  return launch_tc_epi<BN, KC, STAGES, 0>(ma, mb, p, grid, st);
}

// `in`: the gathered operand (NHWC bf16 dense), `out`: result (NHWC bf16 dense), up=false: DOWN geometry
// (in is the fine side), up=true: UP geometry (in is the coarse side).  wpacked: see b200gan_pack_conv_weight.
static int tc_conv_common(const b200gan_conv* cv, const b200gan_view* in, const void* wpacked, const b200gan_view* out, bool up,
                          const TcEpi& epi, cudaStream_t st) {
  if (cv->k != 4 || cv->stride != 2 || cv->pad != 1) return 1;
  if (!wpacked) return 1;
  if (!nhwc_dense_bf16(in) || !nhwc_dense_bf16(out)) return 1;
  const int cin = in->c, cout = out->c;
  if (cin % 32 != 0 || cout % 32 != 0) return 1;
  if (epi.mode != 0 && cout > 1024) return 1;
  if (epi.mode >= 2 && (!nhwc_dense_bf16(epi.prev_y) || epi.prev_y->n != out->n || epi.prev_y->h != out->h || epi.prev_y->w != out->w ||
                        epi.prev_y->c != out->c))
    return 1;
  const int KC = cin % 64 == 0 ? 64 : 32;
  // 128 x 256 tiles for the wide layers (one CTA per SM, all 512 TMEM columns): a 128 x 128 tile needs 128 B/cycle/SM of
  // operands at full MMA rate, more than the L2 delivers (the 128-wide kernel sits at ~55 % tensor-pipe with L2 at ~58 %)
  static const bool wide = getenv("B200GAN_NO_BN256") == nullptr;
  const int BN = (wide && cout % 256 == 0 && cin % 64 == 0) ? 256 : (cout % 128 == 0 ? 128 : (cout % 64 == 0 ? 64 : 32));
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return B200GAN_ERR_CUDA; }
  // GEMM-row pixel space: DOWN -> output pixels (OH,OW); UP -> input pixels (H,W) per parity class
  const int QH = up ? in->h : out->h, QW = up ? in->w : out->w, NB = in->n;
  int TW, TH;
  pick_tile(QH, QW, &TW, &TH);
  const int TN = 128 / (TW * TH);
  TcConvParams p{};
  p.tiles_w = QW / TW; p.tiles_h = QH / TH; p.tiles_n = (NB + TN - 1) / TN;
  p.tw_log2 = ilog2_exact(TW); p.th_log2 = ilog2_exact(TH);
  p.a_mul = up ? 1 : 2;
  p.taps = up ? 4 : 16;
  p.chunks = cin / KC;
  for (int cls = 0; cls < 4; ++cls)
    for (int t = 0; t < p.taps; ++t) {
      if (up) {
        const int py = cls >> 1, px = cls & 1, jh = t >> 1, jw = t & 1;
        p.tap_dh[cls][t] = (int8_t)((py + 1) / 2 - jh);
        p.tap_dw[cls][t] = (int8_t)((px + 1) / 2 - jw);
      } else {
        p.tap_dh[cls][t] = (int8_t)((t >> 2) - 1);
        p.tap_dw[cls][t] = (int8_t)((t & 3) - 1);
      }
    }
  p.QH = QH; p.QW = QW; p.NB = NB;
  p.out = reinterpret_cast<__nv_bfloat16*>(out->ptr);
  p.o_sn = out->sn; p.o_sh = out->sh; p.o_sw = out->sw; p.o_mul = up ? 2 : 1; p.cout = cout;
  if (epi.mode == 1 || epi.mode == 2) {
    p.sums = epi.sums;
    B200_CUDA(cudaMemsetAsync(epi.sums, 0, sizeof(double) * 2 * cout, st));
  }
  if (epi.mode >= 2) {
    p.prev_y = reinterpret_cast<const __nv_bfloat16*>(epi.prev_y->ptr);
    p.prev_scale = epi.scale; p.prev_shift = epi.shift; p.prev_mean = epi.mean; p.prev_invstd = epi.invstd;
    p.prev_neg = epi.act == B200GAN_ACT_RELU ? 0.f : (epi.act == B200GAN_ACT_LRELU ? epi.slope : 1.f);
  }

  CUtensorMap ma, mb;
  {
    const int s = up ? 1 : 2;
    cuuint64_t gdim[4] = {(cuuint64_t)cin, (cuuint64_t)in->w, (cuuint64_t)in->h, (cuuint64_t)in->n};
    cuuint64_t gstr[3] = {(cuuint64_t)cin * 2, (cuuint64_t)in->w * cin * 2, (cuuint64_t)in->h * in->w * cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)(TW * s), (cuuint32_t)(TH * s), (cuuint32_t)TN};
    cuuint32_t estr[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, in->ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(A) failed: %d", (int)r); return B200GAN_ERR_CUDA; }
  }
  {
    const int ktot = p.taps * cin, ncls = up ? 4 : 1;
    cuuint64_t gdim[3] = {(cuuint64_t)ktot, (cuuint64_t)cout, (cuuint64_t)ncls};
    cuuint64_t gstr[2] = {(cuuint64_t)ktot * 2, (cuuint64_t)ktot * cout * 2};
    cuuint32_t box[3] = {(cuuint32_t)KC, (cuuint32_t)BN, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wpacked), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(B) failed: %d", (int)r); return B200GAN_ERR_CUDA; }
  }
  p.n_tiles = cout / BN;
  p.ncls = up ? 4 : 1;
  p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles * p.ncls;
  p.resident = (p.n_tiles == 1 && (int64_t)p.ncls * p.taps * cin * BN * 2 <= 64 * 1024) ? 1 : 0;
  // persistent: two CTAs per SM (every configuration below fits 2 x (smem, 2*BN TMEM columns) per SM)
  const int per_sm = BN == 256 ? 1 : 2;
  const int ctas = p.num_tiles < per_sm * kNumSMs ? p.num_tiles : per_sm * kNumSMs;
  dim3 grid((unsigned)ctas, 1, 1);
  const int e = epi.mode;
  if (KC == 64) {
    if (BN == 256) return launch_tc<256, 64, 4>(ma, mb, p, e, grid, st);   // 4 x 48 KB, one CTA per SM
    if (BN == 128) return launch_tc<128, 64, 3>(ma, mb, p, e, grid, st);   // 3 x 32 KB
    if (BN == 64) return launch_tc<64, 64, 4>(ma, mb, p, e, grid, st);     // 4 x 24 KB
    return launch_tc<32, 64, 5>(ma, mb, p, e, grid, st);                   // 5 x 20 KB
  }
  if (BN == 128) return launch_tc<128, 32, 6>(ma, mb, p, e, grid, st);     // 6 x 16 KB
  if (BN == 64) return launch_tc<64, 32, 8>(ma, mb, p, e, grid, st);       // 8 x 12 KB
  return launch_tc<32, 32, 8>(ma, mb, p, e, grid, st);                     // 8 x 10 KB
}

// ---------------------------------------------------------------------------------------------------
// "Up" geometry for the thin layer (64 -> 32 channels: Conv2d(32->64) input gradient D1, ConvTranspose2d(64->32) forward G4):
// all FOUR output-parity classes of a 128-pixel input tile in one CTA pass.  The generic kernel above treats every class as its
// own tile and fetches 16 tap tiles + 16 weight tiles (320 KB) per 128 input pixels from L2; these layers are bound by exactly
// that traffic (L2 at 60 %, tensor pipe < 20 %).  Here
//   * ONE halo tile of the input (18 lines x 10 pixels x 64 channels = 23 KB for 16 x 8 input pixels) is fetched per tile; the nine
//     3x3-neighbour operands are nine shifted windows into it: a K-major SWIZZLE_128B descriptor may start at any 128-byte row
//     of a TMA-written tile and step between 8-row groups with any stride (the swizzle is a function of the shared-memory
//     address; measured with tools/micro/desc_shift.cu), so "shift by one pixel / one line" is a start-address offset,
//   * all weights (4 classes x 4 taps x 32 x 64 bf16 = 64 KB) stay in shared memory for the CTA's lifetime,
//   * classes whose accumulators are adjacent in TMEM are merged into one MMA: the centre tile feeds all four classes with a
//     single N = 128 instruction, three of the edge tiles with N = 64 (class order (0,0) (0,1) (1,1) (1,0) makes them adjacent):
//     10 instead of 16 MMAs per 16 channels, and the 4 KB A tile is read from shared memory 10 instead of 16 times,
//   * a warp writes both x-parities of a pixel: 128 contiguous bytes per input pixel and output row.
// ---------------------------------------------------------------------------------------------------
struct Up4Params {
  int tiles_w, tiles_h, num_tiles;      // tiles of 8 (W) x 16 (H) input pixels of ONE image
  int QH, QW, NB;
  __nv_bfloat16* out;                 // (NB, 2QH, 2QW, 32) dense
  double* sums;
  const __nv_bfloat16* prev_y;
  float prev_neg;
  int nstages, off_res, off_bar, off_y;
  int yreg;                           // EPI 3: 1 = the saved activation is prefetched into registers one tile ahead, 0 = TMA-staged
};

// neighbour order: centre first (it initialises all four accumulators), then edges, then corners
__constant__ int8_t kUp4Dh[9] = {0, -1, 1, 0, 0, -1, -1, 1, 1};
__constant__ int8_t kUp4Dw[9] = {0, 0, 0, 1, -1, -1, 1, -1, 1};

// the 40 MMAs of one tile, fully unrolled: neighbour windows and weight slots are compile-time offsets added to the two base
// descriptors (the 14-bit start-address field cannot carry: every address stays below 256 KB)
template <int NB, int GI>
struct Up4Table {
  static constexpr int dh[9] = {0, -1, 1, 0, 0, -1, -1, 1, 1};
  static constexpr int dw[9] = {0, 0, 0, 1, -1, -1, 1, -1, 1};
  static constexpr int first[9][2] = {{0, -1}, {0, -1}, {2, -1}, {1, -1}, {0, 3}, {0, -1}, {1, -1}, {3, -1}, {2, -1}};
  static constexpr int count[9][2] = {{4, 0}, {2, 0}, {2, 0}, {2, 0}, {1, 1}, {1, 0}, {1, 0}, {1, 0}, {1, 0}};
  __host__ __device__ static constexpr int slot_before(int nb, int gi) {
    int sl = 0;
    for (int i = 0; i < 9; ++i)
      for (int g = 0; g < 2; ++g) {
        if (i == nb && g == gi) return sl;
        sl += count[i][g];
      }
    return sl;
  }
};

template <int NB, int GI>
__device__ __forceinline__ void up4_issue_group(uint32_t tmem_d, uint64_t adesc0, uint64_t bdesc0) {
  using T = Up4Table<NB, GI>;
  constexpr int cnt = T::count[NB][GI];
  if constexpr (cnt > 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 32 * cnt, 0, 0);
    constexpr int a_off = (1 + T::dh[NB]) * ((8 + 2) * 128) + (1 + T::dw[NB]) * 128;
    constexpr int b_off = T::slot_before(NB, GI) * 4096;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      tcgen05_mma_f16_elect(tmem_d + T::first[NB][GI] * 32, adesc0 + (uint64_t)((a_off + k * 32) >> 4), bdesc0 + (uint64_t)((b_off + k * 32) >> 4),
                            idesc, (NB | k) != 0);
  }
}

template <int NB>
__device__ __forceinline__ void up4_issue_from(uint32_t tmem_d, uint64_t adesc0, uint64_t bdesc0) {
  up4_issue_group<NB, 0>(tmem_d, adesc0, bdesc0);
  up4_issue_group<NB, 1>(tmem_d, adesc0, bdesc0);
  if constexpr (NB + 1 < 9) up4_issue_from<NB + 1>(tmem_d, adesc0, bdesc0);
}

__device__ __forceinline__ void up4_issue_tile(uint32_t tmem_d, uint64_t adesc0, uint64_t bdesc0) { up4_issue_from<0>(tmem_d, adesc0, bdesc0); }

// byte offset of 16-byte chunk `chunk` of pixel `pix` in a tile of 64-byte pixel rows written / read with SWIZZLE_64B
__device__ __forceinline__ uint32_t sw64(int pix, int chunk) { return (uint32_t)pix * 64u + (uint32_t)((chunk ^ ((pix >> 1) & 3)) << 4); }

constexpr int kUp4TW = 8, kUp4TH = 16;                                  // 128 GEMM rows = 16 lines x 8 pixels of one image
constexpr int kUp4Pitch = (kUp4TW + 2) * 128;                           // bytes between lines of the halo tile
constexpr int kUp4TileBytes = (kUp4TH + 2) * kUp4Pitch;                 // 23040
constexpr int kUp4Stage = (kUp4TileBytes + 1023) & ~1023;               // 23552

template <int EPI>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv_up4_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_y,
                   const __grid_constant__ CUtensorMap map_o, const Up4Params p) {
  constexpr int WB_BYTES = 32 * 64 * 2;                              // one (class, tap) weight block
  constexpr int NACC = 4;                                            // 4 x 128 accumulator columns: all of TMEM (one CTA per SM)
  constexpr int YST = 3, Y_BYTES = 32 * 16 * 64;                     // output tiles (32 lines x 16 pixels x 32 ch) staged for the TMA
                                                                     // store; with EPI 3 the saved activation is TMA-loaded into them first
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int NST = p.nstages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* empty_bar = full_bar + 16;
  uint64_t* acc_full = empty_bar + 16;
  uint64_t* acc_empty = acc_full + 4;
  uint64_t* res_bar = acc_empty + 4;
  uint64_t* y_full = res_bar + 1;              // [YST]
  uint64_t* y_empty = y_full + 4;              // [YST]
  uint64_t* staged = y_empty + 4;              // [YST] the eight epilogue warps have written their part of the output tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(staged + 4);
  uint8_t* smem_y = smem + p.off_y;
  float* ch_acc = reinterpret_cast<float*>(smem + p.off_bar + 512);   // [2][32] (EPI 1)
  uint8_t* smem_res = smem + p.off_res;                               // 16 weight blocks, see slot table below
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kTmaWarp = 8, kMmaWarp = 9, kStoreWarp = 10;
  if (EPI == 1) for (int c = threadIdx.x; c < 64; c += blockDim.x) ch_acc[c] = 0.f;
  if (warp == kTmaWarp && lane == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < NACC; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8); }
    mbar_init(res_bar, 1);
    for (int b = 0; b < YST; ++b) { mbar_init(&y_full[b], 1); mbar_init(&y_empty[b], 1); mbar_init(&staged[b], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (EPI == 3) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_o) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  // Column block b of the accumulator holds class kCls[b] = (py,px): (0,0) (0,1) (1,1) (1,0).  For neighbour (dh,dw) the
  // classes with py in P(dh), px in P(dw) (P(-1) = {0}, P(0) = {0,1}, P(1) = {1}) take tap (jh,jw) = (py-dh, px-dw).
  // Weight blocks are stored neighbour by neighbour in column-block order, so a run of adjacent blocks is one B operand.
  if (warp == kTmaWarp) {
    if (lane == 0) {
      mbar_expect_tx(res_bar, 16 * WB_BYTES);
      int slot = 0;
      for (int nb = 0; nb < 9; ++nb) {
        const int dh = kUp4Dh[nb], dw = kUp4Dw[nb];
        for (int b = 0; b < 4; ++b) {
          const int py = b >> 1, px = (b & 1) ^ py;                  // blocks 0..3 -> (0,0) (0,1) (1,1) (1,0)
          const int jh = py - dh, jw = px - dw;
          if (jh < 0 || jh > 1 || jw < 0 || jw > 1) continue;
          tma_load_3d(smem_res + slot * WB_BYTES, &map_b, res_bar, (jh * 2 + jw) * 64, 0, py * 2 + px);
          ++slot;
        }
      }
      int s = 0, ys = 0;
      uint32_t ph = 0, yph = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int n = t / tiles_per_img, r = t - n * tiles_per_img;
        const int th_i = r / p.tiles_w, tw_i = r - th_i * p.tiles_w;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], kUp4TileBytes);
        tma_load_4d(smem + s * kUp4Stage, &map_a, &full_bar[s], 0, tw_i * kUp4TW - 1, th_i * kUp4TH - 1, n);
        if (++s == NST) { s = 0; ph ^= 1; }
        if (EPI == 3 && !p.yreg) {
          // the saved activation under this tile's output (32 lines x 16 pixels), for the epilogue: staged by TMA because a
          // per-thread global load of 128 bytes at a 128-byte lane stride costs 32 L1 wavefronts per instruction
          mbar_wait(&y_empty[ys], yph ^ 1);
          mbar_expect_tx(&y_full[ys], Y_BYTES);
          tma_load_4d(smem_y + ys * Y_BYTES, &map_y, &y_full[ys], 0, 2 * tw_i * kUp4TW, 2 * th_i * kUp4TH, n);
          if (++ys == YST) { ys = 0; yph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer: the whole warp runs the loop (uniform control flow), one elected lane issues each instruction =====
    {
      // per neighbour: up to two MMA groups (first column block, number of blocks); weight slots advance in the same order
      // nb:            0 centre   1 (-1,0)   2 (1,0)    3 (0,1)    4 (0,-1)          5 (-1,-1) 6 (-1,1)  7 (1,-1)  8 (1,1)
      // blocks:        0-3        0-1        2-3        1-2        0 and 3           0         1         3         2
      // Everything is unrolled with compile-time tables (Up4Table): the issuing warp must not chase table loads or rebuild
      // 64-bit descriptors per MMA (a first version with runtime tables took ~300 cycles per MMA instead of ~50).
      const uint32_t tm0 = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t res0 = smem_u32(smem_res);
      const uint64_t bdesc0 = make_smem_desc(res0, 16, 8 * 128, 2u);
      int s = 0;
      uint32_t ph = 0;
      int lt = 0;
      mbar_wait(res_bar, 0);
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++lt) {
        const int buf = lt % NACC;
        mbar_wait(&acc_empty[buf], ((lt / NACC) & 1) ^ 1);
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        const uint32_t tmem_d = tm0 + buf * 128;
        const uint64_t adesc0 = make_smem_desc(smem_u32(smem + s * kUp4Stage), 16, kUp4Pitch, 2u);
        up4_issue_tile(tmem_d, adesc0, bdesc0);
        tcgen05_commit_elect(&empty_bar[s]);
        tcgen05_commit_elect(&acc_full[buf]);
        if (++s == NST) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == kStoreWarp) {
    // ===== output stores: one thread turns every staged tile into ONE coalesced TMA store and hands the staging buffer back as
    // soon as the store has read it.  A thread of its own may block on that read; the epilogue warps never meet at a CTA barrier.
    if (lane == 0) {
      int ys = 0;
      uint32_t sph = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int n = t / tiles_per_img, r = t - n * tiles_per_img;
        const int th_i = r / p.tiles_w, tw_i = r - th_i * p.tiles_w;
        mbar_wait(&staged[ys], sph);
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&map_o),
                     "r"(smem_u32(smem_y + ys * Y_BYTES)), "r"(0), "r"(2 * tw_i * kUp4TW), "r"(2 * th_i * kUp4TH), "r"(n)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(&y_empty[ys]);
        if (++ys == YST) { ys = 0; sph ^= 1; }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // all stores complete before the CTA exits
    }
    __syncwarp();
  } else {
    // ===== epilogue: warps 0..7; TMEM lane quarter = warp % 4 (input pixel rows), half = warp / 4 = output row parity py =====
    const int q = warp & 3, py = warp >> 2;
    const int row = q * 32 + lane;
    const int tw = row & (kUp4TW - 1), th = row >> 3;
    const int64_t o_sh = (int64_t)2 * p.QW * 32, o_sn = (int64_t)2 * p.QH * o_sh;
    // (valid, output offset) of this thread's pixel pair in tile t
    auto locate = [&](int t, bool& valid) -> int64_t {
      const int n = t / tiles_per_img, r = t - n * tiles_per_img;
      const int th_i = r / p.tiles_w, tw_i = r - th_i * p.tiles_w;
      const int iw = tw_i * kUp4TW + tw, ih = th_i * kUp4TH + th;
      valid = iw < p.QW && ih < p.QH && t < p.num_tiles;
      // output pixels (2ih+py, 2iw) and (2ih+py, 2iw+1): 64 contiguous channels
      return (int64_t)n * o_sn + (int64_t)(2 * ih + py) * o_sh + (int64_t)(2 * iw) * 32;
    };
    int lt = 0, ys = 0;
    uint32_t yph = 0;
    float st0 = 0.f, st1 = 0.f, st2 = 0.f, st3 = 0.f;
    uint4 ypre[EPI == 3 ? 8 : 1];
    auto prefetch_y = [&](int t) {
      if (EPI != 3) return;
      bool v2;
      const int64_t off = locate(t, v2);
#pragma unroll
      for (int j = 0; j < (EPI == 3 ? 8 : 1); ++j) ypre[j] = make_uint4(0u, 0u, 0u, 0u);
      if (v2) {
#pragma unroll
        for (int j = 0; j < (EPI == 3 ? 8 : 0); j += 2) ldg256_nc(reinterpret_cast<const uint4*>(p.prev_y + off) + j, ypre[j], ypre[j + 1]);
      }
    };
    if (EPI == 3 && p.yreg) prefetch_y(blockIdx.x);
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++lt) {
      bool valid;
      (void)locate(t, valid);
      uint8_t* yt = smem_y + ys * Y_BYTES;
      // this thread's two output pixels (line 2th+py, pixels 2tw and 2tw+1) inside the 64B-swizzled 32 x 16 pixel tile
      const int pix = (2 * th + py) * 16 + 2 * tw;
      uint4 yv[8];
      if (EPI == 3 && p.yreg) {
#pragma unroll
        for (int j = 0; j < 8; ++j) yv[j] = ypre[j];
        prefetch_y(t + gridDim.x);
        mbar_wait(&y_empty[ys], yph ^ 1);
      } else if (EPI == 3) {
        mbar_wait(&y_full[ys], yph);                    // saved activation landed (the producer waited for the buffer)
#pragma unroll
        for (int j = 0; j < 8; ++j) yv[j] = *reinterpret_cast<const uint4*>(yt + sw64(pix + (j >> 2), j & 3));
      } else {
        mbar_wait(&y_empty[ys], yph ^ 1);               // the TMA store that last used this buffer has read it
      }
      const int buf = lt % NACC;
      mbar_wait(&acc_full[buf], (lt / NACC) & 1);
      tcgen05_fence_after();
      // this warp's two column blocks: py = 0 -> blocks 0,1 = px 0,1; py = 1 -> blocks 2,3 = px 1,0
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        const int blk = 2 * py + bb, px = bb ^ py;
        uint32_t v[32];
        tcgen05_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 128 + blk * 32, v);
        tcgen05_wait_ld();
        if (EPI == 3) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 y4 = yv[px * 4 + j];
            const uint32_t w[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xffff0000u);
              v[8 * j + 2 * e] = __float_as_uint(__uint_as_float(v[8 * j + 2 * e]) * (lo > 0.f ? 1.f : p.prev_neg));
              v[8 * j + 2 * e + 1] = __float_as_uint(__uint_as_float(v[8 * j + 2 * e + 1]) * (hi > 0.f ? 1.f : p.prev_neg));
            }
          }
        }
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          pk[j] = *reinterpret_cast<uint32_t*>(&b);
        }
        // into the staging tile (conflict-free thanks to the swizzle); rows outside the image are clipped by the TMA store
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(yt + sw64(pix + px, j)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      // publish this warp's part of the staged tile to the async proxy and to the store thread
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&staged[ys]);
      if (EPI == 1) {
        mbar_wait(&staged[ys], yph);                   // all eight warps have written (the store thread reads it concurrently)
        // BatchNorm statistics from the staged (bf16-rounded) tile: warp w owns lines 4w..4w+3 (64 pixels), a half-warp one pixel,
        // lane l the channel pair 2(l%16), +1 (conflict-free 4-byte shared loads), accumulated in registers over the CTA's tiles
        const int n = t / tiles_per_img, r = t - n * tiles_per_img;
        const int th_i = r / p.tiles_w, tw_i = r - th_i * p.tiles_w;
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
          const int px = warp * 64 + 2 * i + (lane >> 4);
          if (2 * tw_i * kUp4TW + (px & 15) < 2 * p.QW && 2 * th_i * kUp4TH + (px >> 4) < 2 * p.QH) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(yt + sw64(px, (lane >> 2) & 3) + (lane & 3) * 4);
            const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
            st0 += lo; st1 += hi; st2 = fmaf(lo, lo, st2); st3 = fmaf(hi, hi, st3);
          }
        }
      }
      if (++ys == YST) { ys = 0; yph ^= 1; }
    }
    if (EPI == 1) {
      atomicAdd(&ch_acc[2 * (lane & 15)], st0); atomicAdd(&ch_acc[2 * (lane & 15) + 1], st1);
      atomicAdd(&ch_acc[32 + 2 * (lane & 15)], st2); atomicAdd(&ch_acc[32 + 2 * (lane & 15) + 1], st3);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x < 64 && ch_acc[threadIdx.x] != 0.f) atomicAdd(p.sums + threadIdx.x, (double)ch_acc[threadIdx.x]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

template <int EPI>
static int launch_up4(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& my, const CUtensorMap& mo, const Up4Params& p, int grid, int smem,
                      cudaStream_t st) {
  static int configured = 0;
  if (configured < smem) {
    B200_CUDA(cudaFuncSetAttribute(conv_up4_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  conv_up4_tc_kernel<EPI><<<grid, kHaloThreads, smem, st>>>(ma, mb, my, mo, p);
  B200_LAUNCH_CHECK("conv_up4_tc_kernel");
  return 0;
}

// returns 1 when the problem is not the 64 -> 32 channel "up" shape (or carries an epilogue this kernel does not have)
static int tc_conv_up4(const b200gan_view* in, const void* wpacked, const b200gan_view* out, const TcEpi& epi, cudaStream_t st) {
  static const bool enabled = getenv("B200GAN_NO_UP4") == nullptr;
  if (!enabled || in->c != 64 || out->c != 32 || epi.mode == 2) return 1;
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return B200GAN_ERR_CUDA; }
  if (in->h < 12 || in->w < 8) return 1;                               // small maps waste most of a 16 x 8 tile: generic kernel
  Up4Params p{};
  p.tiles_w = (in->w + kUp4TW - 1) / kUp4TW; p.tiles_h = (in->h + kUp4TH - 1) / kUp4TH;
  p.num_tiles = p.tiles_w * p.tiles_h * in->n;
  p.QH = in->h; p.QW = in->w; p.NB = in->n;
  p.out = reinterpret_cast<__nv_bfloat16*>(out->ptr);
  if (epi.mode == 1) {
    p.sums = epi.sums;
    B200_CUDA(cudaMemsetAsync(epi.sums, 0, sizeof(double) * 64, st));
  }
  if (epi.mode == 3) {
    p.prev_y = reinterpret_cast<const __nv_bfloat16*>(epi.prev_y->ptr);
    p.prev_neg = epi.act == B200GAN_ACT_RELU ? 0.f : (epi.act == B200GAN_ACT_LRELU ? epi.slope : 1.f);
    // measured at B=512 (tools/one_kernel.py d1_up 512 mask): TMA-staged 237 us, register prefetch one tile ahead 281 us (each
    // 32-byte-per-lane load at a 128-byte lane stride costs 32 L1 wavefronts); the knob stays for re-measurement
    static const int yreg = getenv("B200GAN_UP4_YREG") ? atoi(getenv("B200GAN_UP4_YREG")) : 0;
    p.yreg = yreg;
  }
  CUtensorMap ma, mb;
  {
    cuuint64_t gdim[4] = {64, (cuuint64_t)in->w, (cuuint64_t)in->h, (cuuint64_t)in->n};
    cuuint64_t gstr[3] = {128, (cuuint64_t)in->w * 128, (cuuint64_t)in->h * in->w * 128};
    cuuint32_t box[4] = {64, kUp4TW + 2, kUp4TH + 2, 1};                // the halo tile
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, in->ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(A) failed: %d", (int)r); return B200GAN_ERR_CUDA; }
  }
  {
    cuuint64_t gdim[3] = {256, 32, 4};                                 // wpacked "up" form: [class][32 rows][4 taps x 64]
    cuuint64_t gstr[2] = {512, 512 * 32};
    cuuint32_t box[3] = {64, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wpacked), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(B) failed: %d", (int)r); return B200GAN_ERR_CUDA; }
  }
  // output tile (and, for the mask epilogue, the saved activation under it): 32 lines x 16 pixels x 32 channels, 64B swizzle
  CUtensorMap my, mo;
  for (int which = 0; which < 2; ++which) {
    void* base = which == 0 ? out->ptr : (epi.mode == 3 ? epi.prev_y->ptr : out->ptr);
    cuuint64_t gdim[4] = {32, (cuuint64_t)out->w, (cuuint64_t)out->h, (cuuint64_t)out->n};
    cuuint64_t gstr[3] = {64, (cuuint64_t)out->w * 64, (cuuint64_t)out->h * out->w * 64};
    cuuint32_t box[4] = {32, 2 * kUp4TW, 2 * kUp4TH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(which == 0 ? &mo : &my, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(out) failed: %d", (int)r); return B200GAN_ERR_CUDA; }
  }
  // One CTA per SM: two halo tiles in flight (23 KB each; the fill traffic is small) + all weights (64 KB) + three output staging tiles
  p.nstages = 2;
  p.off_res = p.nstages * kUp4Stage;
  p.off_y = p.off_res + 16 * 4096;
  p.off_bar = p.off_y + 3 * 32 * 16 * 64;
  const int smem = 1024 + p.off_bar + 512 + 256;
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  if (epi.mode == 1) return launch_up4<1>(ma, mb, my, mo, p, grid, smem, st);
  if (epi.mode == 3) return launch_up4<3>(ma, mb, my, mo, p, grid, smem, st);
  return launch_up4<0>(ma, mb, my, mo, p, grid, smem, st);
}

// ---------------------------------------------------------------------------------------------------
// "Down" geometry for the thin layer (32 -> 64 channels: Conv2d(32->64) forward D1, ConvTranspose2d(64->32) input gradient G4),
// the mirror image of conv_up4_tc_kernel: per 16 x 8 output pixels of one image the 34 x 18 input pixels are fetched ONCE as
// four stride-2 parity planes (17 lines x 9 pixels x 32 channels each, TMA element strides {1,2,2,1}); tap (kh,kw) is the window
// of plane (kh&1, kw&1) that starts at line kh>>1, pixel kw>>1 (SWIZZLE_64B descriptors are address based too:
// tools/micro/desc_shift64.cu).  The generic kernel fetches 16 tap tiles of 8 KB per 128 pixels; this one 38 KB in total, and all
// 64 KB of weights stay resident.  Epilogue: 1 = BatchNorm statistics, 2 = activation backward + BatchNorm-backward sums with the
// saved conv output TMA-staged into the output staging tile; output by TMA store.
// ---------------------------------------------------------------------------------------------------
struct Down4Params {
  int tiles_w, tiles_h, num_tiles;      // tiles of 8 (W) x 16 (H) OUTPUT pixels of one image
  int OH, OW, NB;
  double* sums;
  const float *prev_scale, *prev_shift, *prev_mean, *prev_invstd;
  float prev_neg;
  int nstages, yst, off_res, off_io, off_bar;
};

constexpr int kDn4PlaneBytes = 17 * 9 * 64;                              // 9792
constexpr int kDn4PlaneStride = (kDn4PlaneBytes + 1023) & ~1023;         // 10240
constexpr int kDn4Stage = 4 * kDn4PlaneStride;                           // 40960
constexpr int kDn4IoBytes = 128 * 128;                                   // 16 lines x 8 pixels x 64 channels

template <int TAP>
__device__ __forceinline__ void down4_issue_from(uint32_t tmem_d, uint64_t adesc0, uint64_t bdesc0) {
  constexpr int kh = TAP >> 2, kw = TAP & 3;
  constexpr int a_off = ((kh & 1) * 2 + (kw & 1)) * kDn4PlaneStride + ((kh >> 1) * 9 + (kw >> 1)) * 64;
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
#pragma unroll
  for (int k = 0; k < 2; ++k)
    tcgen05_mma_f16_elect(tmem_d, adesc0 + (uint64_t)((a_off + k * 32) >> 4), bdesc0 + (uint64_t)((TAP * 4096 + k * 32) >> 4), idesc, (TAP | k) != 0);
  if constexpr (TAP + 1 < 16) down4_issue_from<TAP + 1>(tmem_d, adesc0, bdesc0);
}

// byte offset of 16-byte chunk `chunk` (0..7) of pixel `pix` in a tile of 128-byte pixel rows with SWIZZLE_128B
__device__ __forceinline__ uint32_t sw128(int pix, int chunk) { return (uint32_t)pix * 128u + (uint32_t)((chunk ^ (pix & 7)) << 4); }

template <int EPI>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv_down4_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_y,
                     const __grid_constant__ CUtensorMap map_o, const Down4Params p) {
  constexpr int NACC = 4;
  const int YST = p.yst;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int NST = p.nstages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* acc_full = empty_bar + 8;
  uint64_t* acc_empty = acc_full + 4;
  uint64_t* res_bar = acc_empty + 4;
  uint64_t* y_full = res_bar + 1;
  uint64_t* y_empty = y_full + 4;
  uint64_t* staged = y_empty + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(staged + 4);
  float* ch_acc = reinterpret_cast<float*>(smem + p.off_bar + 512);      // [2][64]
  float4* ch_coef = reinterpret_cast<float4*>(ch_acc + 128);             // [64] {scale, shift, mean, invstd}
  uint8_t* smem_res = smem + p.off_res;
  uint8_t* smem_io = smem + p.off_io;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kTmaWarp = 8, kMmaWarp = 9, kStoreWarp = 10;
  if (EPI != 0) {
    for (int c = threadIdx.x; c < 128; c += blockDim.x) ch_acc[c] = 0.f;
    if (EPI == 2)
      for (int c = threadIdx.x; c < 64; c += blockDim.x) ch_coef[c] = make_float4(p.prev_scale[c], p.prev_shift[c], p.prev_mean[c], p.prev_invstd[c]);
  }
  if (warp == kTmaWarp && lane == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < NACC; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8); }
    mbar_init(res_bar, 1);
    for (int b = 0; b < YST; ++b) { mbar_init(&y_full[b], 1); mbar_init(&y_empty[b], 1); mbar_init(&staged[b], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_o) : "memory");
    if (EPI == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == kTmaWarp) {
    if (lane == 0) {
      mbar_expect_tx(res_bar, 16 * 4096);
      for (int tap = 0; tap < 16; ++tap) tma_load_3d(smem_res + tap * 4096, &map_b, res_bar, tap * 32, 0, 0);
      int s = 0, ys = 0;
      uint32_t ph = 0, yph = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int n = t / tiles_per_img, r = t - n * tiles_per_img;
        const int th_i = r / p.tiles_w, tw_i = r - th_i * p.tiles_w;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], 4 * kDn4PlaneBytes);
#pragma unroll
        for (int pl = 0; pl < 4; ++pl)
          tma_load_4d(smem + s * kDn4Stage + pl * kDn4PlaneStride, &map_a, &full_bar[s], 0, 16 * tw_i - 1 + (pl & 1), 32 * th_i - 1 + (pl >> 1), n);
        if (++s == NST) { s = 0; ph ^= 1; }
        if (EPI == 2) {
          mbar_wait(&y_empty[ys], yph ^ 1);
          mbar_expect_tx(&y_full[ys], kDn4IoBytes);
          tma_load_4d(smem_io + ys * kDn4IoBytes, &map_y, &y_full[ys], 0, 8 * tw_i, 16 * th_i, n);
          if (++ys == YST) { ys = 0; yph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    const uint32_t tm0 = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t bdesc0 = make_smem_desc(smem_u32(smem_res), 16, 8 * 64, 4u);
    int s = 0;
    uint32_t ph = 0;
    int lt = 0;
    mbar_wait(res_bar, 0);
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++lt) {
      const int buf = lt % NACC;
      mbar_wait(&acc_empty[buf], ((lt / NACC) & 1) ^ 1);
      mbar_wait(&full_bar[s], ph);
      tcgen05_fence_after();
      const uint64_t adesc0 = make_smem_desc(smem_u32(smem + s * kDn4Stage), 16, 9 * 64, 4u);
      down4_issue_from<0>(tm0 + buf * 64, adesc0, bdesc0);
      tcgen05_commit_elect(&empty_bar[s]);
      tcgen05_commit_elect(&acc_full[buf]);
      if (++s == NST) { s = 0; ph ^= 1; }
    }
  } else if (warp == kStoreWarp) {
    // ===== output stores (see conv_up4_tc_kernel): one thread, one TMA store per staged tile, buffer handed back once read =====
    if (lane == 0) {
      int ys = 0;
      uint32_t sph = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int n = t / tiles_per_img, r = t - n * tiles_per_img;
        const int th_i = r / p.tiles_w, tw_i = r - th_i * p.tiles_w;
        mbar_wait(&staged[ys], sph);
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&map_o),
                     "r"(smem_u32(smem_io + ys * kDn4IoBytes)), "r"(0), "r"(8 * tw_i), "r"(16 * th_i), "r"(n)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(&y_empty[ys]);
        if (++ys == YST) { ys = 0; sph ^= 1; }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
  } else {
    // ===== epilogue: warps 0..7; TMEM lane quarter = warp % 4 (output pixel rows), column half = warp / 4 =====
    const int q = warp & 3, hcol = warp >> 2;
    const int row = q * 32 + lane;                 // = th * 8 + tw = pixel index inside the 16 x 8 tile
    const int tw = row & 7, th = row >> 3;
    int lt = 0, ys = 0;
    uint32_t yph = 0;
    float st0 = 0.f, st1 = 0.f, st2 = 0.f, st3 = 0.f;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++lt) {
      const int n = t / tiles_per_img, r = t - n * tiles_per_img;
      const int th_i = r / p.tiles_w, tw_i = r - th_i * p.tiles_w;
      const bool valid = 8 * tw_i + tw < p.OW && 16 * th_i + th < p.OH;
      uint8_t* io = smem_io + ys * kDn4IoBytes;
      uint4 yv[4];
      if (EPI == 2) {
        mbar_wait(&y_full[ys], yph);
#pragma unroll
        for (int j = 0; j < 4; ++j) yv[j] = *reinterpret_cast<const uint4*>(io + sw128(row, 4 * hcol + j));
      } else {
        mbar_wait(&y_empty[ys], yph ^ 1);
      }
      const int buf = lt % NACC;
      mbar_wait(&acc_full[buf], (lt / NACC) & 1);
      tcgen05_fence_after();
      uint32_t v[32];
      tcgen05_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 64 + hcol * 32, v);
      tcgen05_wait_ld();
      float ym[32];
      if (EPI == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float y8[8];
          unpack8(yv[j], y8);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 cf = ch_coef[32 * hcol + 8 * j + e];
            const float z = fmaf(y8[e], cf.x, cf.y);
            v[8 * j + e] = __float_as_uint(__uint_as_float(v[8 * j + e]) * (z > 0.f ? 1.f : p.prev_neg));
            ym[8 * j + e] = y8[e] - cf.z;
          }
        }
      }
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
        pk[j] = *reinterpret_cast<uint32_t*>(&b);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(io + sw128(row, 4 * hcol + j)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      if (EPI == 2) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float s0[16], s1[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t w = pk[8 * hh + j];
            const float lo = valid ? __uint_as_float(w << 16) : 0.f, hi = valid ? __uint_as_float(w & 0xffff0000u) : 0.f;
            s0[2 * j] = lo; s0[2 * j + 1] = hi;
            s1[2 * j] = lo * ym[16 * hh + 2 * j]; s1[2 * j + 1] = hi * ym[16 * hh + 2 * j + 1];
          }
          warp_column_sums(s0, lane);
          warp_column_sums(s1, lane);
          if (lane < 16) {
            atomicAdd(&ch_acc[32 * hcol + 16 * hh + lane], s0[0]);
            atomicAdd(&ch_acc[64 + 32 * hcol + 16 * hh + lane], s1[0]);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&staged[ys]);
      if (EPI == 1) {
        mbar_wait(&staged[ys], yph);
        // BatchNorm statistics from the staged (bf16-rounded) tile: warp w owns pixels 16w..16w+15, lane l the channel pair 2l, 2l+1
        // (one conflict-free 4-byte shared load per pixel), accumulated in registers over all tiles of the CTA
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
          const int pix = warp * 16 + i;
          if (8 * tw_i + (pix & 7) < p.OW && 16 * th_i + (pix >> 3) < p.OH) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(io + sw128(pix, lane >> 2) + (lane & 3) * 4);
            const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
            st0 += lo; st1 += hi; st2 = fmaf(lo, lo, st2); st3 = fmaf(hi, hi, st3);
          }
        }
      }
      if (++ys == YST) { ys = 0; yph ^= 1; }
    }
    if (EPI == 1) {
      atomicAdd(&ch_acc[2 * lane], st0); atomicAdd(&ch_acc[2 * lane + 1], st1);
      atomicAdd(&ch_acc[64 + 2 * lane], st2); atomicAdd(&ch_acc[64 + 2 * lane + 1], st3);
    }
    if (EPI != 0) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x < 128) {
        const float a0 = ch_acc[threadIdx.x];
        const double sc = (EPI == 2 && threadIdx.x >= 64) ? (double)ch_coef[threadIdx.x - 64].w : 1.0;
        if (a0 != 0.f) atomicAdd(p.sums + threadIdx.x, (double)a0 * sc);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}

template <int EPI>
static int launch_down4(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& my, const CUtensorMap& mo, const Down4Params& p, int grid,
                        int smem, cudaStream_t st) {
  static int configured = 0;
  if (configured < smem) {
    B200_CUDA(cudaFuncSetAttribute(conv_down4_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  conv_down4_tc_kernel<EPI><<<grid, kHaloThreads, smem, st>>>(ma, mb, my, mo, p);
  B200_LAUNCH_CHECK("conv_down4_tc_kernel");
  return 0;
}

// returns 1 when the problem is not the 32 -> 64 channel "down" shape (or carries an epilogue this kernel does not have)
static int tc_conv_down4(const b200gan_view* in, const void* wpacked, const b200gan_view* out, const TcEpi& epi, cudaStream_t st) {
  static const bool enabled = getenv("B200GAN_NO_DOWN4") == nullptr;
  if (!enabled || in->c != 32 || out->c != 64 || epi.mode == 3 || out->h < 12 || out->w < 8) return 1;
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return B200GAN_ERR_CUDA; }
  Down4Params p{};
  p.tiles_w = (out->w + 7) / 8; p.tiles_h = (out->h + 15) / 16;
  p.num_tiles = p.tiles_w * p.tiles_h * out->n;
  p.OH = out->h; p.OW = out->w; p.NB = out->n;
  if (epi.mode != 0) {
    p.sums = epi.sums;
    B200_CUDA(cudaMemsetAsync(epi.sums, 0, sizeof(double) * 128, st));
    if (epi.mode == 2) {
      p.prev_scale = epi.scale; p.prev_shift = epi.shift; p.prev_mean = epi.mean; p.prev_invstd = epi.invstd;
      p.prev_neg = epi.act == B200GAN_ACT_RELU ? 0.f : (epi.act == B200GAN_ACT_LRELU ? epi.slope : 1.f);
    }
  }
  CUtensorMap ma, mb, my, mo;
  {
    cuuint64_t gdim[4] = {32, (cuuint64_t)in->w, (cuuint64_t)in->h, (cuuint64_t)in->n};
    cuuint64_t gstr[3] = {64, (cuuint64_t)in->w * 64, (cuuint64_t)in->h * in->w * 64};
    cuuint32_t box[4] = {32, 18, 34, 1};                                 // every second pixel: 9 x 17 land in shared memory
    cuuint32_t estr[4] = {1, 2, 2, 1};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, in->ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(A planes) failed: %d", (int)r); return B200GAN_ERR_CUDA; }
  }
  {
    cuuint64_t gdim[3] = {512, 64, 1};                                   // wpacked "down" form: [64 rows][16 taps x 32]
    cuuint64_t gstr[2] = {1024, 1024 * 64};
    cuuint32_t box[3] = {32, 64, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wpacked), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(B) failed: %d", (int)r); return B200GAN_ERR_CUDA; }
  }
  for (int which = 0; which < 2; ++which) {
    void* base = which == 0 ? out->ptr : (epi.mode == 2 ? epi.prev_y->ptr : out->ptr);
    cuuint64_t gdim[4] = {64, (cuuint64_t)out->w, (cuuint64_t)out->h, (cuuint64_t)out->n};
    cuuint64_t gstr[3] = {128, (cuuint64_t)out->w * 128, (cuuint64_t)out->h * out->w * 128};
    cuuint32_t box[4] = {64, 8, 16, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(which == 0 ? &mo : &my, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(out) failed: %d", (int)r); return B200GAN_ERR_CUDA; }
  }
  // 227 KB: 64 KB of weights + either 3 input stages and 2 staging tiles or 2 and 3 (the saved output of epilogue 2 is prefetched
  // into the staging tiles, which then want the depth more than the input ring does)
  static const int force_nst = getenv("B200GAN_DOWN4_NST") ? atoi(getenv("B200GAN_DOWN4_NST")) : 0;
  p.nstages = force_nst ? force_nst : (epi.mode == 2 ? 2 : 3);
  p.yst = p.nstages == 3 ? 2 : 3;
  p.off_res = p.nstages * kDn4Stage;
  p.off_io = p.off_res + 16 * 4096;
  p.off_bar = p.off_io + p.yst * kDn4IoBytes;
  const int smem = 1024 + p.off_bar + 512 + 512 + 1024 + 64;
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  if (epi.mode == 1) return launch_down4<1>(ma, mb, my, mo, p, grid, smem, st);
  if (epi.mode == 2) return launch_down4<2>(ma, mb, my, mo, p, grid, smem, st);
  return launch_down4<0>(ma, mb, my, mo, p, grid, smem, st);
}

// `epi` describes an optional epilogue fusion (mode 0: none).  Both return 0 when the kernel ran (fusion included),
// 1 when the problem does not qualify for the tensor-core path.
int tc_conv_fprop(const b200gan_conv* cv, const b200gan_view* x, const void* wpacked, const b200gan_view* y, const TcEpi& epi,
                  cudaStream_t st) {
  if (cv->k == 4 && cv->stride == 2 && cv->pad == 1 && wpacked && nhwc_dense_bf16(x) && nhwc_dense_bf16(y) &&
      (epi.mode < 2 || (nhwc_dense_bf16(epi.prev_y) && epi.prev_y->n == y->n && epi.prev_y->h == y->h && epi.prev_y->w == y->w && epi.prev_y->c == y->c))) {
    const int t = tc_conv_down4(x, wpacked, y, epi, st);
    if (t <= 0) return t;
  }
  return tc_conv_common(cv, x, wpacked, y, /*up=*/false, epi, st);
}
int tc_conv_dgrad(const b200gan_conv* cv, const b200gan_view* dy, const void* wpacked, const b200gan_view* dx, const TcEpi& epi,
                  cudaStream_t st) {
  if (cv->k == 4 && cv->stride == 2 && cv->pad == 1 && wpacked && nhwc_dense_bf16(dy) && nhwc_dense_bf16(dx) &&
      (epi.mode < 2 || (nhwc_dense_bf16(epi.prev_y) && epi.prev_y->n == dx->n && epi.prev_y->h == dx->h && epi.prev_y->w == dx->w && epi.prev_y->c == dx->c))) {
    const int t = tc_conv_up4(dy, wpacked, dx, epi, st);
    if (t <= 0) return t;
  }
  return tc_conv_common(cv, dy, wpacked, dx, /*up=*/true, epi, st);
}

// ---------------------------------------------------------------------------------------------------
// weight gradient on tensor cores.  conv geometry: x (N,H,W,Ci) fine side, dy (N,OH,OW,Co) coarse side,
//   dw[co,ci,kh,kw] += sum_{n,oh,ow} dy[n,oh,ow,co] * x[n,2oh-1+kh,2ow-1+kw,ci]
// Per tap this is a GEMM with the PIXEL index as the reduction, so both operands are "MN-major" (channels contiguous),
// which tcgen05 reads directly from the NHWC tiles TMA delivers (64 pixels per K-block, 4 MMAs of K = 16 pixels):
//   A (M = 128): x at MT = 128/CIC taps side by side -- 4 taps x 32 channels, 2 taps x 64 channels or 1 tap x a 128-channel
//                chunk: thin layers fill the 128 MMA rows with taps instead of wasting them -- boxes {<=64 ch, 2TW, 2TH, TN}
//                with element strides {1,2,2,1};
//   B (N = NCO): dy, NCO = 64/128/256 output channels, boxes {64 ch, TW, TH, TN}, loaded once per K-block and reused by the
//                G accumulator groups (= G*MT taps) the CTA owns; G*NCO <= 512 TMEM columns.
// A CTA walks a contiguous range of K-blocks (split-K over blockIdx.z).  Epilogue: TMEM lane = (tap, ci), column = co; with a
// workspace in the K-major layout ws[co][tap][ci] the 32 lanes of a warp hit 32 consecutive floats, so every
// red.global.add is one coalesced 128-byte transaction (the first version added straight into the (Co,Ci,4,4) master
// layout: 32 scattered 4-byte atomics per instruction, ~45 us of epilogue per CTA); wgrad_finalize_kernel then transposes
// the workspace into dw and clears it.
// ---------------------------------------------------------------------------------------------------
struct TcWgradParams {
  int tiles_w, tiles_h, tiles_n;
  int tw_log2, th_log2;
  int kb_total, kb_per_split;
  int tap_blocks, ci_chunks, co_chunks;     // blockIdx.x = (co chunk * ci_chunks + ci chunk) * tap_blocks + tap block
  int Co, Ci;
  float* out;        // workspace ws[co][tap][ci] (ws_layout = 1) or the gradient dw[co][ci][tap] itself (ws_layout = 0)
  int ws_layout;
};

template <int CIC, int NCO, int G, int STAGES>
struct TcWgradSmem {
  static constexpr int BK = 64;                            // pixels per K-block
  static constexpr int DY_SLOTS = (STAGES + G - 1) / G + 1 > 3 ? (STAGES + G - 1) / G + 1 : 3;   // safe while STAGES <= (DY_SLOTS - 1) * G
  static constexpr int A_BYTES = BK * 128 * 2;             // one group: 64 pixels x 128 (tap, ci) rows
  static constexpr int B_BYTES = BK * NCO * 2;
  static constexpr int TOTAL = STAGES * A_BYTES + DY_SLOTS * B_BYTES + 1024 + 256;
  static_assert(STAGES <= (DY_SLOTS - 1) * G, "dy slot ring too short for the A stage ring");
};

template <int CIC, int NCO, int G, int STAGES>
__global__ void __launch_bounds__(192, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, const TcWgradParams p) {
  using S = TcWgradSmem<CIC, NCO, G, STAGES>;
  constexpr int MT = 128 / CIC;                            // taps per accumulator group
  constexpr int ABOX = CIC >= 64 ? 64 : CIC;               // channels per x box (128B or 64B rows)
  constexpr int A_NBOX = 128 / ABOX;                       // x boxes per group (taps x channel halves)
  constexpr int A_ROW = ABOX * 2;                          // bytes per pixel row of an x box
  constexpr int A_BOX_BYTES = S::BK * A_ROW;
  constexpr uint32_t A_LT = ABOX == 64 ? 2u : 4u;          // SWIZZLE_128B : SWIZZLE_64B
  constexpr int B_NBOX = NCO / 64;
  constexpr uint32_t TMEM_COLS = G * NCO <= 256 ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * S::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + S::DY_SLOTS * S::B_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int u = blockIdx.x;
  const int tap_block = u % p.tap_blocks; u /= p.tap_blocks;
  const int ci_chunk = u % p.ci_chunks;
  const int co_chunk = u / p.ci_chunks;
  const int tap0 = tap_block * (G * MT), ci0 = ci_chunk * (CIC >= 128 ? 128 : CIC), co0 = co_chunk * NCO;
  const int kb_beg = blockIdx.z * p.kb_per_split;
  const int kb_end = min(kb_beg + p.kb_per_split, p.kb_total);
  const int nkb = kb_end - kb_beg;
  const int TW = 1 << p.tw_log2, TH = 1 << p.th_log2, TN = 64 >> (p.tw_log2 + p.th_log2);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int t0 = kb_beg;
      int tw_i = t0 % p.tiles_w; t0 /= p.tiles_w;
      int th_i = t0 % p.tiles_h;
      int tn_i = t0 / p.tiles_h;
      int s = 0, slot = 0;
      uint32_t ph = 0;
      for (int kbl = 0; kbl < nkb; ++kbl) {
        const int w0 = tw_i * TW, h0 = th_i * TH, n0 = tn_i * TN;
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem_a + s * S::A_BYTES;
          if (g == 0) {
            uint8_t* sb = smem_b + slot * S::B_BYTES;
            mbar_expect_tx(&full_bar[s], S::A_BYTES + S::B_BYTES);
#pragma unroll
            for (int b = 0; b < B_NBOX; ++b) tma_load_4d(sb + b * (S::BK * 128), &map_dy, &full_bar[s], co0 + b * 64, w0, h0, n0);
          } else {
            mbar_expect_tx(&full_bar[s], S::A_BYTES);
          }
#pragma unroll
          for (int b = 0; b < A_NBOX; ++b) {
            // box b of the group: tap (g*MT + b / boxes-per-tap), channel half (b % boxes-per-tap)
            constexpr int BPT = A_NBOX / MT;                 // boxes per tap: 1 (CIC <= 64) or 2 (CIC = 128)
            const int tap = tap0 + g * MT + b / BPT, kh = tap >> 2, kw = tap & 3;
            tma_load_4d(sa + b * A_BOX_BYTES, &map_x, &full_bar[s], ci0 + (b % BPT) * 64, 2 * w0 - 1 + kw, 2 * h0 - 1 + kh, n0);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (++slot == S::DY_SLOTS) slot = 0;
        if (++tw_i == p.tiles_w) { tw_i = 0; if (++th_i == p.tiles_h) { th_i = 0; ++tn_i; } }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, NCO, 1, 1);       // both operands MN-major
      int s = 0, slot = 0;
      uint32_t ph = 0;
      const uint64_t adesc0 = make_smem_desc(smem_u32(smem_a), A_BOX_BYTES, 8 * A_ROW, A_LT);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(smem_b), S::BK * 128, 8 * 128, 2u);
      for (int kbl = 0; kbl < nkb; ++kbl) {
        const uint64_t bdesc = bdesc0 + (uint64_t)((uint32_t)(slot * S::B_BYTES) >> 4);
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          const uint64_t adesc = adesc0 + (uint64_t)((uint32_t)(s * S::A_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < S::BK / 16; ++k) {
            // MN-major canonical layout: LBO = distance between swizzle atoms along M/N (one TMA box), SBO = 8 pixel rows;
            // a k-step advances 16 pixel rows
            tcgen05_mma_f16(tmem_base + g * NCO, adesc + (uint64_t)((k * 16 * A_ROW) >> 4), bdesc + (uint64_t)((k * 16 * 128) >> 4), idesc,
                            (kbl | k) != 0);
          }
          tcgen05_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (++slot == S::DY_SLOTS) slot = 0;
      }
      tcgen05_commit(accum_bar);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;                                        // accumulator row = (tap within group, channel)
    const int t_in = m / (CIC >= 128 ? 128 : CIC), ci = ci0 + m % (CIC >= 128 ? 128 : CIC);
    mbar_wait(accum_bar, 0);
    tcgen05_fence_after();
    if (nkb > 0) {
#pragma unroll 1
      for (int g = 0; g < G; ++g) {
        const int tap = tap0 + g * MT + t_in;
#pragma unroll 1
        for (int c0 = 0; c0 < NCO; c0 += 32) {
          uint32_t r[32];
          tcgen05_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + g * NCO + c0, r);
          tcgen05_wait_ld();
          if (p.ws_layout) {
            float* dst = p.out + ((int64_t)(co0 + c0) * 16 + tap) * p.Ci + ci;      // + j * 16 * Ci per column
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + (int64_t)j * 16 * p.Ci, __uint_as_float(r[j]));
          } else {
            float* dst = p.out + ((int64_t)(co0 + c0) * p.Ci + ci) * 16 + tap;      // + j * Ci * 16 per column
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + (int64_t)j * p.Ci * 16, __uint_as_float(r[j]));
          }
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// dw[co][ci][tap] += ws[co][tap][ci]; ws = 0 (the workspace is handed back zeroed).  One CTA per output channel: its 16 x Ci
// slab (<= 32 KB) goes through shared memory with every thread's loads of a phase issued back to back (blocks of 512 floats with
// one load per thread and phase were latency bound: 15 us per launch, twelve launches per iteration).  Ci % 4 == 0.
__global__ void __launch_bounds__(256) wgrad_finalize_kernel(float* __restrict__ ws, float* __restrict__ dw, int Co, int Ci) {
  extern __shared__ float fin_t[];                                         // [16][Ci + 1]
  const int P = Ci + 1, n = 16 * Ci;
  float* src = ws + (int64_t)blockIdx.x * n;
  float* dst = dw + (int64_t)blockIdx.x * n;
#pragma unroll 4
  for (int i0 = threadIdx.x * 4; i0 < n; i0 += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(src + i0);           // tap = i0 / Ci, four consecutive ci
    *reinterpret_cast<float4*>(src + i0) = make_float4(0.f, 0.f, 0.f, 0.f);
    const int tap = i0 / Ci, ci = i0 - tap * Ci;
    float* t = fin_t + tap * P + ci;
    t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
  }
  __syncthreads();
#pragma unroll 4
  for (int i0 = threadIdx.x * 4; i0 < n; i0 += 1024) {
    const int ci = i0 >> 4, tap0 = i0 & 15;                                // four consecutive taps of one (co, ci)
    float4 d = *reinterpret_cast<const float4*>(dst + i0);
    const float* t = fin_t + tap0 * P + ci;
    d.x += t[0]; d.y += t[P]; d.z += t[2 * P]; d.w += t[3 * P];
    *reinterpret_cast<float4*>(dst + i0) = d;
  }
}

template <int CIC, int NCO, int G, int STAGES>
static int launch_wgrad(const CUtensorMap& mx, const CUtensorMap& mdy, const TcWgradParams& p, dim3 grid, cudaStream_t st) {
  using S = TcWgradSmem<CIC, NCO, G, STAGES>;
  static bool configured = false;
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel<CIC, NCO, G, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  conv_wgrad_tc_kernel<CIC, NCO, G, STAGES><<<grid, 192, S::TOTAL, st>>>(mx, mdy, p);
  B200_LAUNCH_CHECK("conv_wgrad_tc_kernel");
  return 0;
}

// workspace: NULL (direct, scattered atomics into dw) or Co*Ci*16 floats, all zero on entry and all zero again on return
int tc_conv_wgrad(const b200gan_conv* cv, const b200gan_view* x, const b200gan_view* dy, float* dw, float* workspace, cudaStream_t st) {
  if (cv->k != 4 || cv->stride != 2 || cv->pad != 1) return 1;
  if (!nhwc_dense_bf16(x) || !nhwc_dense_bf16(dy)) return 1;
  const int Ci = x->c, Co = dy->c;
  if (Co % 64 != 0 || (Ci != 32 && Ci != 64 && Ci % 128 != 0)) return 1;
  const int CIC = Ci >= 128 ? 128 : Ci;
  static const int force_nco = getenv("B200GAN_WGRAD_NCO") ? atoi(getenv("B200GAN_WGRAD_NCO")) : 0;      // measurement knob
  const int NCO = (force_nco && CIC == 128 && Co % force_nco == 0) ? force_nco
                                                                   : ((CIC == 128 && Co % 256 == 0) ? 256 : ((CIC >= 64 && Co % 128 == 0) ? 128 : 64));
  const int G = NCO == 256 ? 2 : 4, MT = 128 / CIC;
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return B200GAN_ERR_CUDA; }
  int TW = 1, TH = 1;
  while (TW * 2 <= 8 && dy->w % (TW * 2) == 0) TW *= 2;
  while (TH * 2 * TW <= 64 && TH * 2 <= 8 && dy->h % (TH * 2) == 0) TH *= 2;
  const int TN = 64 / (TW * TH);
  TcWgradParams p{};
  p.tiles_w = dy->w / TW; p.tiles_h = dy->h / TH; p.tiles_n = (dy->n + TN - 1) / TN;
  p.tw_log2 = ilog2_exact(TW); p.th_log2 = ilog2_exact(TH);
  p.kb_total = p.tiles_w * p.tiles_h * p.tiles_n;
  p.tap_blocks = 16 / (G * MT); p.ci_chunks = Ci / CIC; p.co_chunks = Co / NCO;
  p.Co = Co; p.Ci = Ci;
  p.out = workspace ? workspace : dw; p.ws_layout = workspace ? 1 : 0;
  const int units = p.tap_blocks * p.ci_chunks * p.co_chunks;
  const int ctas_per_sm = (G * NCO <= 256) ? 2 : 1;                      // TMEM columns (and shared memory) per CTA
  int splits = (ctas_per_sm * kNumSMs + units - 1) / units;
  if (splits > p.kb_total) splits = p.kb_total;
  if (splits < 1) splits = 1;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;

  CUtensorMap mdy, mx;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)Co, (cuuint64_t)dy->w, (cuuint64_t)dy->h, (cuuint64_t)dy->n};
    cuuint64_t gstr[3] = {(cuuint64_t)Co * 2, (cuuint64_t)dy->w * Co * 2, (cuuint64_t)dy->h * dy->w * Co * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TN};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&mdy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dy->ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(dy) failed: %d", (int)r); return B200GAN_ERR_CUDA; }
  }
  {
    const int CB = CIC >= 64 ? 64 : CIC;
    cuuint64_t gdim[4] = {(cuuint64_t)Ci, (cuuint64_t)x->w, (cuuint64_t)x->h, (cuuint64_t)x->n};
    cuuint64_t gstr[3] = {(cuuint64_t)Ci * 2, (cuuint64_t)x->w * Ci * 2, (cuuint64_t)x->h * x->w * Ci * 2};
    cuuint32_t box[4] = {(cuuint32_t)CB, (cuuint32_t)(2 * TW), (cuuint32_t)(2 * TH), (cuuint32_t)TN};
    cuuint32_t estr[4] = {1, 2, 2, 1};
    CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x->ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(x) failed: %d", (int)r); return B200GAN_ERR_CUDA; }
  }
  dim3 grid((unsigned)units, 1, (unsigned)splits);
  int rc;
  // deeper rings measured for the two thin configurations (<32,64,4,5>, <64,128,4,10>): no change, they are not latency bound
  if (CIC == 32) rc = launch_wgrad<32, 64, 4, 4>(mx, mdy, p, grid, st);                 // 4 x 16 KB + 3 x 8 KB: two CTAs per SM
  else if (CIC == 64 && NCO == 128) rc = launch_wgrad<64, 128, 4, 8>(mx, mdy, p, grid, st);
  else if (CIC == 64) rc = launch_wgrad<64, 64, 4, 4>(mx, mdy, p, grid, st);
  // 6 x 16 KB of x stages + 4 x 32 KB of dy slots = 224 KB: three K-blocks in flight instead of two (D3 140 -> 127 us)
  else if (NCO == 256) rc = launch_wgrad<128, 256, 2, 6>(mx, mdy, p, grid, st);
  else if (NCO == 128) rc = launch_wgrad<128, 128, 4, 8>(mx, mdy, p, grid, st);
  else rc = launch_wgrad<128, 64, 4, 4>(mx, mdy, p, grid, st);
  if (rc) return rc;
  if (workspace) {
    const int fsmem = 16 * (Ci + 1) * (int)sizeof(float);
    static int fin_configured = 0;
    if (fsmem > 48 * 1024 && fin_configured < fsmem) {
      B200_CUDA(cudaFuncSetAttribute(wgrad_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fsmem));
      fin_configured = fsmem;
    }
    wgrad_finalize_kernel<<<(unsigned)Co, 256, fsmem, st>>>(workspace, dw, Co, Ci);
    B200_LAUNCH_CHECK("wgrad_finalize_kernel");
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// weight repack: fp32 master (Co,Ci,4,4) [conv geometry] -> bf16 K-major GEMM operand
//   form 0 (DOWN): Wp[co][(kh,kw,ci)]                       rows = Co, K = 16*Ci
//   form 1 (UP)  : Wp[cls][ci][(jh,jw,co)], kh = rh + 2*jh   rows = Ci, K = 4*Co, rh = (py+1)%2
// ---------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, int Co, int Ci, int form, __nv_bfloat16* __restrict__ out) {
  const int64_t total = (int64_t)Co * Ci * 16;
  const int64_t work = form == 2 ? 2 * total : total;                  // form 2: both forms back to back
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < work; idx += (int64_t)gridDim.x * blockDim.x) {
    const int f = form == 2 ? (idx >= total ? 1 : 0) : form;
    const int64_t i = idx >= total ? idx - total : idx;
    float v;
    if (f == 0) {
      const int ci = (int)(i % Ci);
      const int tap = (int)((i / Ci) % 16);
      const int co = (int)(i / ((int64_t)Ci * 16));
      v = w[((int64_t)co * Ci + ci) * 16 + tap];
    } else {
      const int co = (int)(i % Co);
      const int j = (int)((i / Co) % 4);
      const int ci = (int)((i / ((int64_t)Co * 4)) % Ci);
      const int cls = (int)(i / ((int64_t)Co * 4 * Ci));
      const int py = cls >> 1, px = cls & 1, jh = j >> 1, jw = j & 1;
      const int kh = (py + 1) % 2 + 2 * jh, kw = (px + 1) % 2 + 2 * jw;
      v = w[((int64_t)co * Ci + ci) * 16 + kh * 4 + kw];
    }
    out[idx] = __float2bfloat16_rn(v);
  }
}

int tc_pack_weight(const float* w, int Co, int Ci, int k, int form, void* out, cudaStream_t st) {
  if (k != 4) { set_error("pack_conv_weight: only k=4 (stride 2, pad 1) layers have a tensor-core path"); return B200GAN_ERR_UNSUPPORTED; }
  const int64_t total = (int64_t)Co * Ci * 16 * (form == 2 ? 2 : 1);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  pack_weight_kernel<<<(unsigned)blocks, 256, 0, st>>>(w, Co, Ci, form, reinterpret_cast<__nv_bfloat16*>(out));
  B200_LAUNCH_CHECK("pack_weight_kernel");
  return 0;
}

}  // namespace b200gan
