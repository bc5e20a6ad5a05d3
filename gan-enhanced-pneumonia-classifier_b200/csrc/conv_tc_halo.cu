// Halo-tile tcgen05 kernels for the thin layers of the DCGAN (32 <-> 64 channels, k4 s2 p1; dcgan.py:42,72): conv_up4_tc_kernel and
// conv_down4_tc_kernel.  One halo tile (or four parity planes) of the input per CTA pass, all weights resident, the per-tap operands
// are shifted descriptor windows into that tile; see the comment above each kernel and DESIGN.md section 4.
#include "tc_common.cuh"

namespace b200gan {

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
      tcgen05_mma_f16_elect32(tmem_d + T::first[NB][GI] * 32, (uint32_t)adesc0 + (uint32_t)((a_off + k * 32) >> 4), (uint32_t)(adesc0 >> 32),
                              (uint32_t)bdesc0 + (uint32_t)((b_off + k * 32) >> 4), (uint32_t)(bdesc0 >> 32), idesc, (NB | k) != 0);
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
  uint8_t* smem = align_smem_1024(smem_raw);
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
        // Branch-free: the load is always inside the tile and the VALUE is masked.  A load under `if` compiles (shared-memory address
        // space known) to a branch per pixel with the load's latency exposed every time: +40 us (down) / +78 us (up) per launch
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
          const int px = warp * 64 + 2 * i + (lane >> 4);
          const bool ok = 2 * tw_i * kUp4TW + (px & 15) < 2 * p.QW && 2 * th_i * kUp4TH + (px >> 4) < 2 * p.QH;
          uint32_t w = *reinterpret_cast<const uint32_t*>(yt + sw64(px, (lane >> 2) & 3) + (lane & 3) * 4);
          w = ok ? w : 0u;
          const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
          st0 += lo; st1 += hi; st2 = fmaf(lo, lo, st2); st3 = fmaf(hi, hi, st3);
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
  B200_CUDA((ensure_dynamic_smem<conv_up4_tc_kernel<EPI>>(smem)));
  conv_up4_tc_kernel<EPI><<<grid, kHaloThreads, smem, st>>>(ma, mb, my, mo, p);
  B200_LAUNCH_CHECK("conv_up4_tc_kernel");
  return 0;
}

// returns 1 when the problem is not the 64 -> 32 channel "up" shape (or carries an epilogue this kernel does not have)
int tc_conv_up4(const b200gan_view* in, const void* wpacked, const b200gan_view* out, const TcEpi& epi, cudaStream_t st) {
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
    tcgen05_mma_f16_elect32(tmem_d, (uint32_t)adesc0 + (uint32_t)((a_off + k * 32) >> 4), (uint32_t)(adesc0 >> 32),
                            (uint32_t)bdesc0 + (uint32_t)((TAP * 4096 + k * 32) >> 4), (uint32_t)(bdesc0 >> 32), idesc, (TAP | k) != 0);
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
  uint8_t* smem = align_smem_1024(smem_raw);
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
    float ra0[2] = {0.f, 0.f}, ra1[2] = {0.f, 0.f};      // EPI 2: column sums of channels 32 hcol + 16 hh + lane % 16
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
          ra0[hh] += s0[0];                               // lanes j and j+16 hold column j: kept in registers over the CTA's tiles (a
          ra1[hh] += s1[0];                               // shared-memory float atomicAdd is a compare-and-swap loop)
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
        // (branch-free: see conv_up4_tc_kernel)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int pix = warp * 16 + i;
          const bool ok = 8 * tw_i + (pix & 7) < p.OW && 16 * th_i + (pix >> 3) < p.OH;
          uint32_t w = *reinterpret_cast<const uint32_t*>(io + sw128(pix, lane >> 2) + (lane & 3) * 4);
          w = ok ? w : 0u;
          const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
          st0 += lo; st1 += hi; st2 = fmaf(lo, lo, st2); st3 = fmaf(hi, hi, st3);
        }
      }
      if (++ys == YST) { ys = 0; yph ^= 1; }
    }
    if (EPI == 1) {
      atomicAdd(&ch_acc[2 * lane], st0); atomicAdd(&ch_acc[2 * lane + 1], st1);
      atomicAdd(&ch_acc[64 + 2 * lane], st2); atomicAdd(&ch_acc[64 + 2 * lane + 1], st3);
    }
    if (EPI == 2 && lane < 16) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) { atomicAdd(&ch_acc[32 * hcol + 16 * hh + lane], ra0[hh]); atomicAdd(&ch_acc[64 + 32 * hcol + 16 * hh + lane], ra1[hh]); }
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
  B200_CUDA((ensure_dynamic_smem<conv_down4_tc_kernel<EPI>>(smem)));
  conv_down4_tc_kernel<EPI><<<grid, kHaloThreads, smem, st>>>(ma, mb, my, mo, p);
  B200_LAUNCH_CHECK("conv_down4_tc_kernel");
  return 0;
}

// returns 1 when the problem is not the 32 -> 64 channel "down" shape (or carries an epilogue this kernel does not have)
int tc_conv_down4(const b200gan_view* in, const void* wpacked, const b200gan_view* out, const TcEpi& epi, cudaStream_t st) {
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

}  // namespace b200gan
