// Microbenchmark: cycles per tcgen05.mma.cta_group::2 (kind::f16, M=256, K=16, SS mode) as a function of N, issued by the leader CTA of a
// 2-CTA cluster on whatever is in the two CTAs' shared memory, next to the cta_group::1 M=128 figure of mma_rate.cu.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/micro/mma_rate_pair tools/micro/mma_rate_pair.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../gan-enhanced-pneumonia-classifier_b200/csrc/ptx.cuh"
using namespace b200gan;

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)lt << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma_rate_pair(long long* out, int iters, int rotate, int commit_every) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint64_t sink[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&sink[i], 1u << 20); mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tm = slot;
  if (rank == 0 && warp == 1 && lane == 0) {
    const uint32_t sa = smem_u32(smem), sb = sa + 96 * 1024;
    const uint32_t id = idesc_bf16(256, N);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t a = sa + (rotate ? (i % 6) * 16384 : 0);
      const uint32_t b = sb + (rotate ? (i % 4) * 16384 : 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) tcgen05_mma_f16_pair(tm, desc(a + k * 32, 16, 1024, 2), desc(b + k * 32, 16, 1024, 2), id, 1);
      if (commit_every == 1) tcgen05_commit_pair(&sink[i & 7]);
      if (commit_every == 2) tcgen05_commit(&sink[i & 7]);
    }
    tcgen05_commit_pair(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x >> 1] = t1 - t0;
  }
  if (rank == 1 && warp == 1 && lane == 0) mbar_wait(&bar, 0);      // the multicast commit arrives here too
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(256));
}

template <int N>
void run(int rotate, int pairs, int commit_every = 0) {
  long long* d;
  const int iters = 2000;
  cudaMalloc(&d, pairs * sizeof(long long));
  cudaFuncSetAttribute(mma_rate_pair<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  mma_rate_pair<N><<<2 * pairs, 128, 170 * 1024>>>(d, iters, rotate, commit_every);
  mma_rate_pair<N><<<2 * pairs, 128, 170 * 1024>>>(d, iters, rotate, commit_every);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[128];
  cudaMemcpy(h, d, pairs * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1ll << 60;
  for (int i = 0; i < pairs; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
  printf("cta_group::2 M=256 N=%3d rotate=%d pairs=%d commit=%d: %.1f .. %.1f cycles per MMA (both SMs)  [%s]\n", N, rotate, pairs, commit_every, (double)mn / (iters * 4),
         (double)mx / (iters * 4), cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  int n = 0;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 170 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaFuncSetAttribute(mma_rate_pair<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, mma_rate_pair<256>, &cfg);
    printf("cudaOccupancyMaxActiveClusters(cluster 2, 170 KB smem): %d  [%s]\n", n, cudaGetErrorString(e));
  }
  run<64>(0, 1); run<128>(0, 1); run<256>(0, 1); run<256>(1, 1);
  run<256>(1, 74); run<128>(1, 74);
  run<256>(1, 74, 1); run<256>(1, 1, 1); run<256>(1, 74, 2);
  return 0;
}
