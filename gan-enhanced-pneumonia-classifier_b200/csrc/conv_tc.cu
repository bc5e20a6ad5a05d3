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
#include "tc_common.cuh"

namespace b200gan {

// ---------------------------------------------------------------------------------------------------


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

template <int BN, int KC, int STAGES, int EPI>
__global__ void __launch_bounds__(kTcThreads, BN == 256 ? 1 : 2)     // the 256-wide tile owns all of TMEM: one CTA per SM anyway
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcConvParams p) {
  using S = TcSmem<BN, KC, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
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
      const uint32_t dhi = (uint32_t)(adesc0 >> 32);     // SBO, version, layout type: the same for both operands
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
          // 32-bit arithmetic on the descriptors' low words (the start-address field cannot carry), see tcgen05_mma_f16_elect32
          const uint32_t alo = (uint32_t)adesc0 + ((uint32_t)(s * p.stage_stride) >> 4);
          const uint32_t blo = p.resident ? (uint32_t)bres0 + ((res_tile + (uint32_t)(kb * S::B_BYTES)) >> 4) : alo + (uint32_t)(S::A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) tcgen05_mma_f16_lohi(tmem_d, alo + 2 * k, dhi, blo + 2 * k, dhi, idesc, (kb | k) != 0);
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
    // BatchNorm sums: lanes j and j+16 hold column j of a 16-column chunk after warp_column_sums; they stay in registers (one pair per
    // chunk of this warp's CW columns) for as long as the CTA's tiles keep the same channel tile, and reach shared memory when it changes
    // and at the end (a shared-memory float atomicAdd is a compare-and-swap loop: two per chunk and tile were a visible part of the fused
    // epilogues' cost)
    constexpr int NCH = CW / 16 > 0 ? CW / 16 : 1;
    float ra0[NCH], ra1[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) ra0[i] = ra1[i] = 0.f;
    int acc_cout0 = -1;
    auto flush_sums = [&]() {
      if (lane < 16) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          if (ra0[i] != 0.f) atomicAdd(&ch_acc[acc_cout0 + 16 * i + lane], ra0[i]);
          if (ra1[i] != 0.f) atomicAdd(&ch_acc[p.cout + acc_cout0 + 16 * i + lane], ra1[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < NCH; ++i) ra0[i] = ra1[i] = 0.f;
    };
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
      if ((EPI == 1 || EPI == 2) && cout0 != acc_cout0) {
        if (acc_cout0 >= 0) flush_sums();
        acc_cout0 = cout0;
      }
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
#pragma unroll
          for (int i = 0; i < NCH; ++i) {                  // compile-time register indices whether or not the chunk loop is unrolled
            const bool mine = i == (c0 >> 4);
            ra0[i] += mine ? s0[0] : 0.f;
            ra1[i] += mine ? s1[0] : 0.f;
          }
        }
      }
      // all TMEM reads of this warp are complete (wait::ld above): hand the accumulator back to the MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    if (EPI == 1 || EPI == 2) {
      if (acc_cout0 >= 0) flush_sums();
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
// choose TW x TH x TN = 128 output pixels per tile: powers of two dividing the spatial extent
static void pick_tile(int qh, int qw, int* tw, int* th) {
  int w = 1;
  while (w * 2 <= 16 && qw % (w * 2) == 0) w *= 2;
  int h = 1;
  while (h * 2 * w <= 128 && h * 2 <= 16 && qh % (h * 2) == 0) h *= 2;
  *tw = w; *th = h;
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
  B200_CUDA((ensure_dynamic_smem<conv_gemm_tc_kernel<BN, KC, STAGES, EPI>>(smem)));
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
  // wide layers, opt-in (B200GAN_PAIR=1): a CTA pair per 256 x 256 tile (tcgen05 cta_group::2, conv_tc_pair.cu).  Correct, but measured
  // 2x slower than the one-CTA kernel on this workload (see the header of conv_tc_pair.cu), so it is not the default.
  const char* pair_env = getenv("B200GAN_PAIR");             // read per call: tests switch it inside one process
  const bool pair_ok = pair_env != nullptr && atoi(pair_env) != 0;
  if (pair_ok && KC == 64 && BN == 256 && p.tiles_w * p.tiles_h * p.tiles_n >= 2) {
    CUtensorMap mbh;
    const int ktot = p.taps * cin;
    cuuint64_t gdim[3] = {(cuuint64_t)ktot, (cuuint64_t)cout, (cuuint64_t)p.ncls};
    cuuint64_t gstr[2] = {(cuuint64_t)ktot * 2, (cuuint64_t)ktot * cout * 2};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&mbh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wpacked), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(B half) failed: %d", (int)r); return B200GAN_ERR_CUDA; }
    const int t = launch_tc_pair(ma, mbh, p, e, st);
    if (t <= 0) return t;
  }
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
    int t = tc_conv_up4(dy, wpacked, dx, epi, st);
    if (t <= 0) return t;
    t = tc_conv_up4w(dy, wpacked, dx, epi, st);
    if (t <= 0) return t;
  }
  return tc_conv_common(cv, dy, wpacked, dx, /*up=*/true, epi, st);
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
