// Weight gradients of the k4 s2 p1 layers on tcgen05 tensor cores (conv_wgrad_tc_kernel) and the workspace finalize pass.
#include "tc_common.cuh"

namespace b200gan {

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
  uint8_t* smem = align_smem_1024(smem_raw);
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
            tcgen05_mma_f16_lohi(tmem_base + g * NCO, (uint32_t)adesc + (uint32_t)((k * 16 * A_ROW) >> 4), (uint32_t)(adesc >> 32),
                                 (uint32_t)bdesc + (uint32_t)((k * 16 * 128) >> 4), (uint32_t)(bdesc >> 32), idesc, (kbl | k) != 0);
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
  B200_CUDA((ensure_dynamic_smem<conv_wgrad_tc_kernel<CIC, NCO, G, STAGES>>(S::TOTAL)));
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
    if (fsmem > 48 * 1024) B200_CUDA((ensure_dynamic_smem<wgrad_finalize_kernel>(fsmem)));
    wgrad_finalize_kernel<<<(unsigned)Co, 256, fsmem, st>>>(workspace, dw, Co, Ci);
    B200_LAUNCH_CHECK("wgrad_finalize_kernel");
  }
  return 0;
}

}  // namespace b200gan
