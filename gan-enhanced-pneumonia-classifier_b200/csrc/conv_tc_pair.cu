// CTA-pair variant of the implicit-GEMM convolution (conv_tc.cu) for the wide layers of the DCGAN (output channels % 256 == 0:
// D3 / D4 forward and D4 input gradient, G1 forward, G1 / G2 input gradient; dcgan.py:30-34,76-80): tcgen05.mma.cta_group::2.
//
// Why: the one-CTA kernel with 128 x 256 tiles is bound by shared-memory FILL, not by the tensor pipe: at full MMA rate a CTA must
// take in 16 KB of activations + 32 KB of weights every 512 cycles (94 B/cycle), and the TMA round trip lets it sustain ~55.  A
// CTA pair computes a 256 x 256 tile: each CTA stages its own 128 pixel rows (16 KB) but only HALF of the weight tile (16 KB); the
// M = 256 MMA issued by the leader reads the other half from the peer's shared memory.  Fill per CTA and k-block drops from 48 KB
// to 32 KB (62 B/cycle at full rate), six stages fit instead of four, and one thread issues the MMAs of two SMs.
//
// MEASURED (round 2, B=512, D3 forward alone, L2 flushed): correct (tests/test_gpu_tc.py against the numpy oracle), but 179 us against
// 90 us for the one-CTA kernel, independent of the stage count (2..6) and of the epilogue (a drain-only epilogue: 172 us).  The MMA
// itself is not the problem (tools/micro/mma_rate_pair.cu: M=256 N=256 K=16 retires every 128 cycles for both SMs, with or without a
// multicast commit per k-block); a cycle trace of the leader shows ~1870 cycles per k-block, of which 400 waiting for its own TMA
// bytes and 830 for the peer's, with both producers starved of free stages: each SM takes in 32 KB (TMA) + 16 KB (the peer's weight
// half, re-read over the SM-to-SM network by every MMA) per k-block, i.e. the same 48 KB as the one-CTA kernel, at a third of its
// rate.  Signalling the leader through a relay thread instead of remote complete_tx measured the same (205 us).  The kernel is
// therefore OPT-IN (B200GAN_PAIR=1) and carries no performance claim; the dispatch default stays the one-CTA kernel.
//
// Protocol (see ptx.cuh): "full" barriers live in the leader and count both CTAs' bytes; tcgen05.commit multicasts the stage
// release and the accumulator-ready arrival to both CTAs; the peer's epilogue warps arrive remotely on the leader's
// accumulator-drained barriers.  Each CTA's epilogue reads its own 128 TMEM lanes exactly like the one-CTA kernel.
#include "tc_common.cuh"

namespace b200gan {

namespace {
constexpr int kKC = 64, kBN = 256;
constexpr int kABytes = 128 * kKC * 2;          // 128 pixel rows x 64 channels
constexpr int kBHalfBytes = 128 * kKC * 2;      // 128 of the 256 weight rows
constexpr int kStageBytes = kABytes + kBHalfBytes;
constexpr int kBarBytes = 512;
constexpr int kNAcc = 2;                        // 2 x 256 fp32 columns = all of TMEM
}  // namespace

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
conv_gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  const int NST = p.nstages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.off_bar);      // used in the leader only
  uint64_t* empty_bar = full_bar + 16;                                     // one set per CTA, released by multicast commits
  uint64_t* acc_full = empty_bar + 16;                                     // [kNAcc] per CTA
  uint64_t* acc_empty = acc_full + 4;                                      // [kNAcc] used in the leader only (16 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 4);
  float* ch_acc = reinterpret_cast<float*>(smem + p.off_bar + kBarBytes);
  float4* ch_coef = reinterpret_cast<float4*>(ch_acc + 2 * p.cout);
  if (EPI == 1 || EPI == 2) {
    for (int c = threadIdx.x; c < 2 * p.cout; c += blockDim.x) ch_acc[c] = 0.f;
    if (EPI == 2)
      for (int c = threadIdx.x; c < p.cout; c += blockDim.x)
        ch_coef[c] = make_float4(p.prev_scale[c], p.prev_shift[c], p.prev_mean[c], p.prev_invstd[c]);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int TW = 1 << p.tw_log2, TH = 1 << p.th_log2, TN = 128 >> (p.tw_log2 + p.th_log2);
  // pair tiles: (pair of M tiles, N tile, parity class); this CTA takes M tile 2*mp + rank (an odd M-tile count leaves the last
  // peer tile empty: its TMA boxes lie outside the tensor and arrive as zeros, its rows are invalid in the epilogue)
  const int num_pt = p.num_tiles;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  constexpr int kTmaWarp = 8, kMmaWarp = 9;
  if (warp == kTmaWarp && lane == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < kNAcc; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == kMmaWarp) {   // one warp of EACH CTA of the pair performs the pair-wide allocation (all 512 columns)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();            // barriers of both CTAs initialised before any remote arrive / multicast commit
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int pt, int& cls, int& nt, int& tw_i, int& th_i, int& tn_i) {
    int r = pt;
    cls = r % p.ncls; r /= p.ncls;
    nt = r % p.n_tiles; r /= p.n_tiles;
    int mt = 2 * r + (int)rank;
    tw_i = mt % p.tiles_w; mt /= p.tiles_w;
    th_i = mt % p.tiles_h;
    tn_i = mt / p.tiles_h;
  };

  if (warp == kTmaWarp) {
    // ===== TMA producer (one lane per CTA): own pixel rows + own half of the weight tile, bytes counted on the leader's barrier =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int pt = cluster_id; pt < num_pt; pt += num_clusters) {
        int cls, nt, tw_i, th_i, tn_i;
        decode(pt, cls, nt, tw_i, th_i, tn_i);
        const int w0 = tw_i * TW * p.a_mul, h0 = th_i * TH * p.a_mul, n0 = tn_i * TN, cout0 = nt * kBN + (int)rank * 128;
        int kcol = 0;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int cw = w0 + p.tap_dw[cls][tap], chh = h0 + p.tap_dh[cls][tap];
          for (int chunk = 0; chunk < p.chunks; ++chunk, kcol += kKC) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* sa = smem + s * kStageBytes;
            const uint32_t lbar = leader_addr(&full_bar[s]);
            if (leader) mbar_expect_tx(&full_bar[s], 2 * kStageBytes);      // one of the two arrivals + the bytes of BOTH CTAs
            else mbar_arrive_cluster(lbar);                                 // the other arrival
            tma_load_4d_pair(sa, &map_a, lbar, chunk * kKC, cw, chh, n0);
            tma_load_3d_pair(sa + kABytes, &map_b, lbar, kcol, cout0, cls);
            if (++s == NST) { s = 0; ph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer: one lane of the LEADER CTA issues M = 256, N = 256 MMAs for both SMs =====
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, kBN, 0, 0);
      constexpr uint32_t SBO = 8 * kKC * 2;
      const int num_kb = p.taps * p.chunks;
      int s = 0;
      uint32_t ph = 0;
      int lt = 0;
      const uint64_t adesc0 = make_smem_desc(smem_u32(smem), 16, SBO, 2u);
      for (int pt = cluster_id; pt < num_pt; pt += num_clusters, ++lt) {
        const int buf = lt % kNAcc;
        mbar_wait(&acc_empty[buf], ((lt / kNAcc) & 1) ^ 1);       // both CTAs' epilogues have drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + buf * kBN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          const uint64_t adesc = adesc0 + (uint64_t)((uint32_t)(s * kStageBytes) >> 4);
          const uint64_t bdesc = adesc + (uint64_t)(kABytes >> 4);
#pragma unroll
          for (int k = 0; k < kKC / 16; ++k) tcgen05_mma_f16_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          tcgen05_commit_pair(&empty_bar[s]);                    // frees this stage in BOTH CTAs
          if (++s == NST) { s = 0; ph ^= 1; }
        }
        tcgen05_commit_pair(&acc_full[buf]);                     // accumulator complete, signalled to both epilogues
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: warps 0..7 of each CTA on its own 128 TMEM lanes (same code as the one-CTA kernel) =====
    const int q = warp & 3, hcol = warp >> 2;
    constexpr int CW = kBN / 2;
    const int row = q * 32 + lane;
    const int tw = row & (TW - 1), th = (row >> p.tw_log2) & (TH - 1), tn = row >> (p.tw_log2 + p.th_log2);
    const uint32_t acc_empty_leader = leader_addr(&acc_empty[0]);
    int lt = 0;
    auto locate = [&](int pt, bool& valid, int& cout0) -> int64_t {
      int cls, nt, tw_i, th_i, tn_i;
      decode(pt, cls, nt, tw_i, th_i, tn_i);
      const int ow = tw_i * TW + tw, oh = th_i * TH + th, n = tn_i * TN + tn;
      cout0 = nt * kBN + hcol * CW;
      valid = ow < p.QW && oh < p.QH && n < p.NB && pt < num_pt;
      const int py = cls >> 1, px = cls & 1;
      return (int64_t)n * p.o_sn + (int64_t)(oh * p.o_mul + (p.o_mul > 1 ? py : 0)) * p.o_sh +
             (int64_t)(ow * p.o_mul + (p.o_mul > 1 ? px : 0)) * p.o_sw + cout0;
    };
    for (int pt = cluster_id; pt < num_pt; pt += num_clusters, ++lt) {
      bool valid; int cout0;
      const int64_t ooff = locate(pt, valid, cout0);
      __nv_bfloat16* orow = p.out + ooff;
      const uint4* yp = reinterpret_cast<const uint4*>(p.prev_y + ooff);
      uint4 ynext[2];
      if (EPI >= 2) {                                  // the first piece of y_prev is requested before the accumulator is waited for
        ynext[0] = ynext[1] = make_uint4(0u, 0u, 0u, 0u);
        if (valid) ldg256_nc(yp, ynext[0], ynext[1]);
      }
      const int buf = lt % kNAcc;
      mbar_wait(&acc_full[buf], (lt / kNAcc) & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < CW; c0 += 16) {
        uint32_t v[16];
        tcgen05_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + buf * kBN + hcol * CW + c0, v);
        float s0[16], s1[16];
        uint4 ycur[2];
        if (EPI >= 2) {
          ycur[0] = ynext[0]; ycur[1] = ynext[1];
          if (c0 + 16 < CW) {
            ynext[0] = ynext[1] = make_uint4(0u, 0u, 0u, 0u);
            if (valid) ldg256_nc(yp + (c0 + 16) / 8, ynext[0], ynext[1]);
          }
        }
        tcgen05_wait_ld();
        if (EPI == 2) {
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
      // all TMEM reads of this warp are complete: hand the accumulator back to the leader's MMA warp (remote arrive for the peer)
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc_empty_leader + (uint32_t)(buf * 8));
    }
    if (EPI == 1 || EPI == 2) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int c = threadIdx.x; c < p.cout; c += 256) {
        const float a0 = ch_acc[c], a1 = ch_acc[p.cout + c];
        if (a0 != 0.f) atomicAdd(p.sums + c, (double)a0);
        if (a1 != 0.f) atomicAdd(p.sums + p.cout + c, (double)a1 * (EPI == 2 ? (double)ch_coef[c].w : 1.0));
      }
    }
  }
  // the peer's shared memory and tensor memory are in use until the leader's last MMA has retired and both epilogues are done
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kMmaWarp) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

template <int EPI>
static int launch_pair_epi(const CUtensorMap& ma, const CUtensorMap& mb, TcConvParams p, cudaStream_t st) {
  // shared memory: as many 32 KB stages as fit next to the barrier block and the channel accumulators (six for every DCGAN layer)
  const int extra = (EPI == 0 || EPI == 3) ? 0 : p.cout * 8 + (EPI == 2 ? p.cout * 16 : 0) + 16;
  int nst = (227 * 1024 - 1024 - kBarBytes - extra) / kStageBytes;
  if (nst > 6) nst = 6;
  if (nst < 3) return 1;
  p.nstages = nst;
  p.stage_stride = kStageBytes;
  p.off_res = p.off_bar = nst * kStageBytes;
  p.resident = 0;
  const int smem = 1024 + p.off_bar + kBarBytes + extra;
  B200_CUDA((ensure_dynamic_smem<conv_gemm_tc_pair_kernel<EPI>>(smem)));
  // pair tiles replace the one-CTA tile count: ceil(M tiles / 2) x N tiles x classes
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.num_tiles = ((m_tiles + 1) / 2) * p.n_tiles * p.ncls;
  // Persistent schedule: exactly as many CTA pairs as can be CO-RESIDENT (a pair needs both SMs of one TPC; the driver knows how many
  // fit with this kernel's shared-memory size: 74 on the pool's B200s).
  static std::atomic<int> resident[64];
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  int pairs = (dev >= 0 && dev < 64) ? resident[dev].load(std::memory_order_relaxed) : 0;
  if (pairs <= 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (kNumSMs / 2)); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = (size_t)smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, conv_gemm_tc_pair_kernel<EPI>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = kNumSMs / 2; }
    pairs = n < kNumSMs / 2 ? n : kNumSMs / 2;
    if (dev >= 0 && dev < 64) resident[dev].store(pairs, std::memory_order_relaxed);
  }
  const int clusters = p.num_tiles < pairs ? p.num_tiles : pairs;
  conv_gemm_tc_pair_kernel<EPI><<<2 * clusters, kTcThreads, smem, st>>>(ma, mb, p);      // __cluster_dims__(2,1,1)
  B200_LAUNCH_CHECK("conv_gemm_tc_pair_kernel");
  return 0;
}

int launch_tc_pair(const CUtensorMap& ma, const CUtensorMap& mb_half, TcConvParams p, int epi, cudaStream_t st) {
  if (epi == 1) return launch_pair_epi<1>(ma, mb_half, p, st);
  if (epi == 2) return launch_pair_epi<2>(ma, mb_half, p, st);
  if (epi == 3) return launch_pair_epi<3>(ma, mb_half, p, st);
  return launch_pair_epi<0>(ma, mb_half, p, st);
}

}  // namespace b200gan
