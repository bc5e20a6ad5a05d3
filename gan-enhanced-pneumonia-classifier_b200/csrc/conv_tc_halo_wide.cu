// Halo-tile tcgen05 kernel for the 128 -> 64 channel "up" layer of the DCGAN (k4 s2 p1): Conv2d(64->128) input gradient D2
// (dcgan.py:74) and ConvTranspose2d(128->64) forward G3 (dcgan.py:38).  conv_up4w_tc_kernel.
//
// The generic kernel (conv_tc.cu) treats each of the four output-parity classes as its own GEMM and fetches, per 128 input pixels,
// 16 tap tiles of 32 KB + 16 weight tiles of 16 KB from L2: 2.4 GB L2->SM for a 0.5 GB problem at batch 512, which is what bounds it
// (161 us with the statistics epilogue, 219 us with the BatchNorm-backward one, against 50-80 us of MMA / DRAM time).  Here, as in
// conv_up4_tc_kernel (conv_tc_halo.cu),
//   * ONE halo tile of the input (18 lines x 10 pixels, two 64-channel chunks of 23 KB) is fetched per 16 x 8 input pixels; the
//     neighbour operands are shifted descriptor windows into it (SWIZZLE_128B descriptors are address based),
//   * a CTA serves ONE output row parity py for its whole lifetime (CTA b: py = b & 1), so the weights it needs -- two row taps x
//     (both column classes) x 128 channels x 64 = 128 KB -- stay resident in shared memory; all four classes would be 256 KB,
//   * the two column classes px = 0,1 sit in adjacent TMEM column blocks: the dw = 0 neighbour feeds both with one N = 128 MMA,
//     dw = -1 / +1 feed one each with N = 64: 6 instead of 8 MMAs per 16 channels,
//   * output (and, for the BatchNorm-backward epilogue, the saved convolution output under it) moves by TMA through 16 KB staging
//     units of 8 output lines x 16 pixels x 64 channels: a 5-d tensor map (c, w, line parity, h/2, n) addresses the lines of one parity.
// Shared memory: 128 KB weights + 2 x 23 KB input chunks + 3 x 16 KB staging units = 222 KB, one CTA per SM, all 512 TMEM columns
// (4 accumulators of 128 columns).
//
// Measured at batch 512 (us per launch, alone with the L2 flushed / inside the iteration; generic kernel in brackets):
//   no epilogue 114 [144], statistics 150 / 139 [169 / 153], BatchNorm backward 207 / 196 [228 / 210].
// What the way there showed (tools/one_kernel.py d2_up, ncu --set full with source counters):
//   * with ONE CTA per SM the issue slots are shared between the single MMA-issuing warp and the epilogue warps of its scheduler, and the
//     tile time follows the epilogue's instruction count: eight epilogue warps ran the BatchNorm-backward epilogue (about 1500 instructions
//     per warp and tile) at one instruction per 6 cycles each -- 316 us with the tensor pipe 24 % busy; sixteen warps: 241 us;
//   * every MMA costs the issuing warp ELECT + VOTEU + five R2UR besides its descriptor arithmetic; 64-bit descriptor adds doubled the
//     arithmetic part (tcgen05_mma_f16_elect32: 241 -> 207 us, 178 -> 150 us);
//   * the result of the BatchNorm-backward variant leaves from registers: staged for a TMA store, a unit was held through load -> epilogue ->
//     store, and three units are 1.5 tiles;
//   * the statistics read-back must be branch-free (a shared-memory load under `if` is compiled to a branch per pixel);
//   * L2 prefetch of the boxes two tiles ahead and two TMA stores in flight made no measurable difference.
#include "tc_common.cuh"

namespace b200gan {

struct Up4wParams {
  int tiles_w, tiles_h, num_tiles;      // tiles of 8 (W) x 16 (H) input pixels of ONE image
  int QH, QW, NB;
  __nv_bfloat16* out;                   // (NB, 2QH, 2QW, 64) dense: EPI 2 stores its result from registers
  double* sums;
  const float *prev_scale, *prev_shift, *prev_mean, *prev_invstd;
  float prev_neg;
  int off_res, off_io, off_bar;
};

constexpr int kUpwTW = 8, kUpwTH = 16;
constexpr int kUpwPitch = (kUpwTW + 2) * 128;                           // bytes between lines of one 64-channel chunk of the halo tile
constexpr int kUpwChunkBytes = (kUpwTH + 2) * kUpwPitch;                // 23040
constexpr int kUpwStage = (kUpwChunkBytes + 1023) & ~1023;              // 23552
constexpr int kUpwNST = 2;                                              // one chunk per stage: one tile of input in flight
constexpr int kUpwWB = 64 * 64 * 2;                                     // one (class, tap, chunk) weight block: 64 rows x 64 channels
constexpr int kUpwUnit = 8 * 16 * 128;                                  // staging unit: 8 output lines x 16 pixels x 64 channels
constexpr int kUpwYST = 3;
constexpr int kUpwThreads = 19 * 32;                                    // 16 epilogue warps + TMA producer + MMA issuer + TMA stores

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// L2 prefetch of a box (no shared memory, no barrier): issued two tiles ahead, the ring's own load then meets L2 instead of DRAM
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_5d(const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global [%0, {%1, %2, %3, %4, %5}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}

// byte offset of 16-byte chunk `chunk` (0..7) of pixel `pix` in a tile of 128-byte pixel rows with SWIZZLE_128B
__device__ __forceinline__ uint32_t upw_sw128(int pix, int chunk) { return (uint32_t)pix * 128u + (uint32_t)((chunk ^ (pix & 7)) << 4); }

// The 24 MMAs of one 64-channel chunk.  Row neighbour dh in {0, DH1} (DH1 = -1 for py = 0, +1 for py = 1; tap jh = py - dh), column
// neighbour dw in {0, -1, +1}.  Weight slots per chunk and row neighbour: [px0 jw0 | px1 jw1] (dw = 0, one N = 128 operand),
// [px0 jw1] (dw = -1), [px1 jw0] (dw = +1).  `first`: this chunk's first MMA initialises the accumulator.
template <int DH1>
__device__ __forceinline__ void upw_issue_chunk(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, bool first) {
  constexpr uint32_t idesc128 = make_idesc_bf16(128, 128, 0, 0), idesc64 = make_idesc_bf16(128, 64, 0, 0);
#pragma unroll
  for (int dhi = 0; dhi < 2; ++dhi) {
    const int dh = dhi == 0 ? 0 : DH1;
    const int row_off = (1 + dh) * kUpwPitch;
    const int b_off = dhi * 4 * kUpwWB;
#pragma unroll
    for (int k = 0; k < 4; ++k)      // dw = 0: both column classes
      tcgen05_mma_f16_elect32(tmem_d, alo + (uint32_t)((row_off + 128 + k * 32) >> 4), ahi, blo + (uint32_t)((b_off + k * 32) >> 4), bhi, idesc128,
              (first && dhi == 0 && k == 0) ? 0u : 1u);
#pragma unroll
    for (int k = 0; k < 4; ++k)      // dw = -1: px = 0 takes tap jw = 1
      tcgen05_mma_f16_elect32(tmem_d, alo + (uint32_t)((row_off + 0 + k * 32) >> 4), ahi, blo + (uint32_t)((b_off + 2 * kUpwWB + k * 32) >> 4), bhi, idesc64, 1u);
#pragma unroll
    for (int k = 0; k < 4; ++k)      // dw = +1: px = 1 takes tap jw = 0
      tcgen05_mma_f16_elect32(tmem_d + 64, alo + (uint32_t)((row_off + 256 + k * 32) >> 4), ahi, blo + (uint32_t)((b_off + 3 * kUpwWB + k * 32) >> 4), bhi, idesc64, 1u);
  }
}

template <int EPI>
__global__ void __launch_bounds__(kUpwThreads, 1)
conv_up4w_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_y,
                    const __grid_constant__ CUtensorMap map_o, const Up4wParams p) {
  constexpr int NACC = 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* acc_full = empty_bar + 4;
  uint64_t* acc_empty = acc_full + 4;
  uint64_t* res_bar = acc_empty + 4;
  uint64_t* y_full = res_bar + 1;              // [YST] EPI 2: the saved convolution output has landed in the staging unit
  uint64_t* y_empty = y_full + 4;              // [YST] the staging unit may be overwritten
  uint64_t* staged = y_empty + 4;              // [YST] the eight epilogue warps of the unit have written their part
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(staged + 4);
  float* ch_acc = reinterpret_cast<float*>(smem + p.off_bar + 512);      // [2][64]
  float4* ch_coef = reinterpret_cast<float4*>(ch_acc + 128);             // [64] {scale, shift, mean, invstd}
  uint8_t* smem_res = smem + p.off_res;
  uint8_t* smem_io = smem + p.off_io;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kTmaWarp = 16, kMmaWarp = 17, kStoreWarp = 18;
  const int py = blockIdx.x & 1;                          // this CTA's output row parity
  const int t_first = blockIdx.x >> 1, t_step = gridDim.x >> 1;
  if (EPI != 0) {
    for (int c = threadIdx.x; c < 128; c += blockDim.x) ch_acc[c] = 0.f;
    if (EPI == 2)
      for (int c = threadIdx.x; c < 64; c += blockDim.x) ch_coef[c] = make_float4(p.prev_scale[c], p.prev_shift[c], p.prev_mean[c], p.prev_invstd[c]);
  }
  if (warp == kTmaWarp && lane == 0) {
    for (int s = 0; s < kUpwNST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < NACC; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 16); }
    mbar_init(res_bar, 1);
    // who releases a staging unit: the store thread (EPI 0), the store thread and the eight warps that read the statistics back from it
    // (EPI 1), the eight warps that consumed the saved convolution output from it (EPI 2: the result is stored from registers)
    for (int b = 0; b < kUpwYST; ++b) { mbar_init(&y_full[b], 1); mbar_init(&y_empty[b], EPI == 1 ? 9 : (EPI == 2 ? 8 : 1)); mbar_init(&staged[b], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_o) : "memory");
    if (EPI == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
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

  if (warp == kTmaWarp) {
    if (lane == 0) {
      // resident weights: [chunk][row neighbour][slot] (see upw_issue_chunk); the packed "up" form is [class][64 rows][4 taps x 128]
      mbar_expect_tx(res_bar, 16 * kUpwWB);
      for (int c = 0; c < 2; ++c)
        for (int dhi = 0; dhi < 2; ++dhi) {
          const int jh = dhi == 0 ? py : 1 - py;           // dh = 0 -> jh = py;  dh = -1 (py 0) -> 1;  dh = +1 (py 1) -> 0
          const int px_of[4] = {0, 1, 0, 1}, jw_of[4] = {0, 1, 1, 0};
          for (int sl = 0; sl < 4; ++sl)
            tma_load_3d(smem_res + ((c * 2 + dhi) * 4 + sl) * kUpwWB, &map_b, res_bar, (jh * 2 + jw_of[sl]) * 128 + c * 64, 0, py * 2 + px_of[sl]);
        }
      int s = 0, ys = 0;
      uint32_t ph = 0, yph = 0;
      // the input ring holds ONE tile: its loads are issued a tile time before they are needed, less than a DRAM round trip.  The
      // boxes of the tile two steps ahead are prefetched into L2 (no shared memory needed; measured at batch 512: 120 -> 115 us without
      // an epilogue, nothing once the epilogue's instruction count set the pace)
      auto prefetch = [&](int t2) {
        if (t2 >= p.num_tiles) return;
        const int n2 = t2 / tiles_per_img, r2 = t2 - n2 * tiles_per_img;
        const int th2 = r2 / p.tiles_w, tw2 = r2 - th2 * p.tiles_w;
        tma_prefetch_4d(&map_a, 0, tw2 * kUpwTW - 1, th2 * kUpwTH - 1, n2);
        tma_prefetch_4d(&map_a, 64, tw2 * kUpwTW - 1, th2 * kUpwTH - 1, n2);
        if (EPI == 2) {
          tma_prefetch_5d(&map_y, 0, 2 * tw2 * kUpwTW, py, th2 * kUpwTH, n2);
          tma_prefetch_5d(&map_y, 0, 2 * tw2 * kUpwTW, py, th2 * kUpwTH + 8, n2);
        }
      };
      prefetch(t_first + t_step);
      for (int t = t_first; t < p.num_tiles; t += t_step) {
        const int n = t / tiles_per_img, r = t - n * tiles_per_img;
        const int th_i = r / p.tiles_w, tw_i = r - th_i * p.tiles_w;
        prefetch(t + 2 * t_step);
        for (int c = 0; c < 2; ++c) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], kUpwChunkBytes);
          tma_load_4d(smem + s * kUpwStage, &map_a, &full_bar[s], c * 64, tw_i * kUpwTW - 1, th_i * kUpwTH - 1, n);
          if (++s == kUpwNST) { s = 0; ph ^= 1; }
        }
        if (EPI == 2) {
          for (int h = 0; h < 2; ++h) {
            mbar_wait(&y_empty[ys], yph ^ 1);
            mbar_expect_tx(&y_full[ys], kUpwUnit);
            tma_load_5d(smem_io + ys * kUpwUnit, &map_y, &y_full[ys], 0, 2 * tw_i * kUpwTW, py, th_i * kUpwTH + 8 * h, n);
            if (++ys == kUpwYST) { ys = 0; yph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer: the whole warp runs the loop (uniform control flow), one elected lane issues each instruction =====
    const uint32_t tm0 = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t bdesc0 = make_smem_desc(smem_u32(smem_res), 16, 8 * 128, 2u);
    int s = 0;
    uint32_t ph = 0;
    int lt = 0;
    mbar_wait(res_bar, 0);
    for (int t = t_first; t < p.num_tiles; t += t_step, ++lt) {
      const int buf = lt % NACC;
      mbar_wait(&acc_empty[buf], ((lt / NACC) & 1) ^ 1);
      const uint32_t tmem_d = tm0 + buf * 128;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        const uint64_t adesc0 = make_smem_desc(smem_u32(smem + s * kUpwStage), 16, kUpwPitch, 2u);
        const uint64_t bdc = bdesc0 + (uint64_t)((c * 8 * kUpwWB) >> 4);
        if (py == 0) upw_issue_chunk<-1>(tmem_d, (uint32_t)adesc0, (uint32_t)(adesc0 >> 32), (uint32_t)bdc, (uint32_t)(bdc >> 32), c == 0);
        else upw_issue_chunk<1>(tmem_d, (uint32_t)adesc0, (uint32_t)(adesc0 >> 32), (uint32_t)bdc, (uint32_t)(bdc >> 32), c == 0);
        tcgen05_commit_elect(&empty_bar[s]);
        if (++s == kUpwNST) { s = 0; ph ^= 1; }
      }
      tcgen05_commit_elect(&acc_full[buf]);
    }
    __syncwarp();
  } else if (warp == kStoreWarp) {
    // ===== output stores: one thread, one TMA store per staged unit, the unit handed back once the store has read it =====
    if (lane == 0 && EPI != 2) {
      int ys = 0;
      for (int t = t_first; t < p.num_tiles; t += t_step) {
        const int n = t / tiles_per_img, r = t - n * tiles_per_img;
        const int th_i = r / p.tiles_w, tw_i = r - th_i * p.tiles_w;
        for (int h = 0; h < 2; ++h) {
          // parity of the ys-th use of a unit: tracked per unit through the running unit count
          mbar_wait(&staged[ys % kUpwYST], (uint32_t)((ys / kUpwYST) & 1));
          asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(&map_o),
                       "r"(smem_u32(smem_io + (ys % kUpwYST) * kUpwUnit)), "r"(0), "r"(2 * tw_i * kUpwTW), "r"(py), "r"(th_i * kUpwTH + 8 * h), "r"(n)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          if (ys > 0) {
            // two stores in flight: the unit of the PREVIOUS store is handed back once that store has read it
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            mbar_arrive(&y_empty[(ys - 1) % kUpwYST]);
          }
          ++ys;
        }
      }
      if (ys > 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(&y_empty[(ys - 1) % kUpwYST]);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // all stores complete before the CTA exits
    }
    __syncwarp();
  } else {
    // ===== epilogue: warps 0..15; TMEM lane quarter q = warp % 4 (input lines 4q..4q+3), px = (warp / 4) % 2 (64-column block of the
    // accumulator), cb = warp / 8 (32-column half of it).  Sixteen warps, not eight: one CTA per SM means the epilogue's instruction
    // stream (about 750 instructions per 32 columns and tile with the BatchNorm-backward fusion) runs on two warps per scheduler with
    // eight of them, and its dependent-issue latency, not the tensor pipe, set the tile time (ncu: issue slots 38 % busy, tensor pipe 24 %)
    const int q = warp & 3, px = (warp >> 2) & 1, cb = warp >> 3;
    const int row = q * 32 + lane;
    const int tw = row & 7, th = row >> 3;
    const int h = q >> 1;                                  // which of the tile's two staging units this warp writes
    const int wi = ((q & 1) * 2 + px) * 2 + cb;            // index among the unit's eight warps
    const int pix = (th & 7) * 16 + 2 * tw + px;           // this thread's output pixel inside the unit
    int lt = 0;
    float st0 = 0.f, st1 = 0.f, st2 = 0.f, st3 = 0.f;
    float ra0[2] = {0.f, 0.f}, ra1[2] = {0.f, 0.f};        // EPI 2: column sums of channels 32 cb + 16 hh + lane % 16
    for (int t = t_first; t < p.num_tiles; t += t_step, ++lt) {
      const int n = t / tiles_per_img, r = t - n * tiles_per_img;
      const int th_i = r / p.tiles_w, tw_i = r - th_i * p.tiles_w;
      const bool valid = tw_i * kUpwTW + tw < p.QW && th_i * kUpwTH + th < p.QH;
      const int u = 2 * lt + h, ys = u % kUpwYST;
      const uint32_t yph = (uint32_t)((u / kUpwYST) & 1);
      uint8_t* io = smem_io + ys * kUpwUnit;
      if (EPI == 2) mbar_wait(&y_full[ys], yph);           // saved convolution output landed (the producer waited for the unit)
      else mbar_wait(&y_empty[ys], yph ^ 1);               // the previous user of the unit is done with it
      const int buf = lt % NACC;
      mbar_wait(&acc_full[buf], (lt / NACC) & 1);
      tcgen05_fence_after();
      uint32_t v[32];
      tcgen05_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 128 + px * 64 + cb * 32, v);
      uint4 yv[4];
      if (EPI == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) yv[j] = *reinterpret_cast<const uint4*>(io + upw_sw128(pix, 4 * cb + j));
      }
      tcgen05_wait_ld();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);         // the accumulator is in registers
      float ym[32];
      if (EPI == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float y8[8];
          unpack8(yv[j], y8);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 cf = ch_coef[32 * cb + 8 * j + e];
            const float z = fmaf(y8[e], cf.x, cf.y);
            v[8 * j + e] = __float_as_uint(__uint_as_float(v[8 * j + e]) * (z > 0.f ? 1.f : p.prev_neg));
            ym[8 * j + e] = y8[e] - cf.z;
          }
        }
        __syncwarp();                                      // every lane has consumed its part of the saved convolution output:
        if (lane == 0) mbar_arrive(&y_empty[ys]);          // the unit goes back to the producer
      }
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
        pk[j] = *reinterpret_cast<uint32_t*>(&b);
      }
      if (EPI == 2) {
        // the result leaves from registers (this thread's output pixel, 32 contiguous channels): a staging unit is then held only from
        // the arrival of the saved convolution output to its consumption.  Held until a TMA store of the result had read it (three
        // units = 1.5 tiles), every tile waited for a full load -> epilogue -> store round trip
        if (valid) {
          __nv_bfloat16* orow = p.out + ((((int64_t)n * 2 * p.QH + 2 * (th_i * kUpwTH + th) + py) * (2 * p.QW)) + 2 * (tw_i * kUpwTW + tw) + px) * 64 + 32 * cb;
          stg256(orow, make_uint4(pk[0], pk[1], pk[2], pk[3]), make_uint4(pk[4], pk[5], pk[6], pk[7]));
          stg256(orow + 16, make_uint4(pk[8], pk[9], pk[10], pk[11]), make_uint4(pk[12], pk[13], pk[14], pk[15]));
        }
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
          ra0[hh] += s0[0];                                // lanes j and j+16 hold column j: kept in registers over the CTA's tiles
          ra1[hh] += s1[0];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(io + upw_sw128(pix, 4 * cb + j)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        // publish this warp's part of the staged unit to the async proxy and to the store thread
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&staged[ys]);
      }
      if (EPI == 1) {
        mbar_wait(&staged[ys], yph);                       // all eight warps have written (the store thread reads the unit concurrently)
        // BatchNorm statistics from the staged (bf16-rounded) unit: warp wi owns pixels 16wi..16wi+15, lane l the channel pair 2l, 2l+1
        // (one conflict-free 4-byte shared load per pixel), accumulated in registers over all tiles of the CTA.  Branch-free: the load is
        // always inside the unit and the VALUE is masked -- a load under `if` compiles to a branch per pixel with the load's latency
        // exposed every time (measured on the sibling kernels: +40 us per launch)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int px2 = wi * 16 + i;
          const bool ok = 2 * tw_i * kUpwTW + (px2 & 15) < 2 * p.QW && th_i * kUpwTH + 8 * h + (px2 >> 4) < p.QH;
          uint32_t w = *reinterpret_cast<const uint32_t*>(io + upw_sw128(px2, lane >> 2) + (lane & 3) * 4);
          w = ok ? w : 0u;
          const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
          st0 += lo; st1 += hi; st2 = fmaf(lo, lo, st2); st3 = fmaf(hi, hi, st3);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&y_empty[ys]);          // this warp no longer reads the unit
      }
    }
    if (EPI == 1) {
      atomicAdd(&ch_acc[2 * lane], st0); atomicAdd(&ch_acc[2 * lane + 1], st1);
      atomicAdd(&ch_acc[64 + 2 * lane], st2); atomicAdd(&ch_acc[64 + 2 * lane + 1], st3);
    }
    if (EPI == 2 && lane < 16) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) { atomicAdd(&ch_acc[32 * cb + 16 * hh + lane], ra0[hh]); atomicAdd(&ch_acc[64 + 32 * cb + 16 * hh + lane], ra1[hh]); }
    }
    if (EPI != 0) {
      asm volatile("bar.sync 1, 512;" ::: "memory");
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

template <int EPI>
static int launch_up4w(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& my, const CUtensorMap& mo, const Up4wParams& p, int grid,
                       int smem, cudaStream_t st) {
  B200_CUDA((ensure_dynamic_smem<conv_up4w_tc_kernel<EPI>>(smem)));
  conv_up4w_tc_kernel<EPI><<<grid, kUpwThreads, smem, st>>>(ma, mb, my, mo, p);
  B200_LAUNCH_CHECK("conv_up4w_tc_kernel");
  return 0;
}

// returns 1 when the problem is not the 128 -> 64 channel "up" shape (or carries an epilogue this kernel does not have)
int tc_conv_up4w(const b200gan_view* in, const void* wpacked, const b200gan_view* out, const TcEpi& epi, cudaStream_t st) {
  const char* off = getenv("B200GAN_NO_UP4W");                         // read per call: tests compare both kernels inside one process
  if ((off != nullptr && atoi(off) != 0) || in->c != 128 || out->c != 64 || epi.mode == 3) return 1;
  if (in->h < 12 || in->w < 8) return 1;                               // small maps waste most of a 16 x 8 tile: generic kernel
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return B200GAN_ERR_CUDA; }
  Up4wParams p{};
  p.tiles_w = (in->w + kUpwTW - 1) / kUpwTW; p.tiles_h = (in->h + kUpwTH - 1) / kUpwTH;
  p.num_tiles = p.tiles_w * p.tiles_h * in->n;
  p.QH = in->h; p.QW = in->w; p.NB = in->n;
  p.out = reinterpret_cast<__nv_bfloat16*>(out->ptr);
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
    cuuint64_t gdim[4] = {128, (cuuint64_t)in->w, (cuuint64_t)in->h, (cuuint64_t)in->n};
    cuuint64_t gstr[3] = {256, (cuuint64_t)in->w * 256, (cuuint64_t)in->h * in->w * 256};
    cuuint32_t box[4] = {64, kUpwTW + 2, kUpwTH + 2, 1};                // one 64-channel chunk of the halo tile
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, in->ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(A) failed: %d", (int)r); return B200GAN_ERR_CUDA; }
  }
  {
    cuuint64_t gdim[3] = {512, 64, 4};                                 // wpacked "up" form: [class][64 rows][4 taps x 128]
    cuuint64_t gstr[2] = {1024, 1024 * 64};
    cuuint32_t box[3] = {64, 64, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wpacked), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(B) failed: %d", (int)r); return B200GAN_ERR_CUDA; }
  }
  // output (and the saved convolution output under it) as (c, w, line parity, h/2, n): a unit is 8 lines of ONE parity x 16 pixels
  for (int which = 0; which < 2; ++which) {
    void* base = which == 0 ? out->ptr : (epi.mode == 2 ? epi.prev_y->ptr : out->ptr);
    const cuuint64_t line = (cuuint64_t)out->w * 128;
    cuuint64_t gdim[5] = {64, (cuuint64_t)out->w, 2, (cuuint64_t)in->h, (cuuint64_t)out->n};
    cuuint64_t gstr[4] = {128, line, 2 * line, (cuuint64_t)out->h * line};
    cuuint32_t box[5] = {64, 16, 1, 8, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(which == 0 ? &mo : &my, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(out) failed: %d", (int)r); return B200GAN_ERR_CUDA; }
  }
  p.off_res = kUpwNST * kUpwStage;
  p.off_io = p.off_res + 16 * kUpwWB;
  p.off_bar = p.off_io + kUpwYST * kUpwUnit;
  const int smem = 1024 + p.off_bar + 512 + 512 + 1024 + 64;
  // an even grid: CTA b serves row parity b & 1 of tiles b >> 1, b >> 1 + grid / 2, ...
  int grid = 2 * p.num_tiles < kNumSMs ? 2 * p.num_tiles : (kNumSMs & ~1);
  if (grid < 2) grid = 2;
  if (epi.mode == 1) return launch_up4w<1>(ma, mb, my, mo, p, grid, smem, st);
  if (epi.mode == 2) return launch_up4w<2>(ma, mb, my, mo, p, grid, smem, st);
  return launch_up4w<0>(ma, mb, my, mo, p, grid, smem, st);
}

}  // namespace b200gan
