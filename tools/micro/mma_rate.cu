// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16, SS mode) as a function of N, and with 1 or 2
// CTAs per SM.  Operands are whatever is in shared memory (values irrelevant).  Build: nvcc -gencode arch=compute_100a,code=sm_100a
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
__global__ void __launch_bounds__(128, 1) mma_rate(long long* out, int iters, int same_a, int shift_rows, int sbo) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = slot;
  if (warp == 1 && lane == 0) {
    const uint32_t sa = smem_u32(smem), sb = sa + 64 * 1024;
    const uint32_t id = idesc_bf16(128, N);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      // 4 k-steps of one 128 x 64 (K-major, SWIZZLE_128B) A tile against an N x 64 B tile; A tile rotates through 4 slots
      const uint32_t a = sa + (same_a ? 0 : (i & 1) * 24576) + shift_rows * 128;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tcgen05_mma_f16(tm, desc(a + k * 32, 16, sbo, 2), desc(sb + k * 32, 16, 1024, 2), id, 1);
    }
    tcgen05_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(256));
}

template <int N>
void run(int ctas_per_sm, int same_a, int shift_rows = 0, int sbo = 1024) {
  long long* d;
  const int grid = 148 * ctas_per_sm, iters = 2000;
  cudaMalloc(&d, grid * sizeof(long long));
  cudaFuncSetAttribute(mma_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  mma_rate<N><<<grid, 128, 100 * 1024>>>(d, iters, same_a, shift_rows, sbo);
  mma_rate<N><<<grid, 128, 100 * 1024>>>(d, iters, same_a, shift_rows, sbo);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[296];
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%3d ctas/SM=%d same_a=%d shift=%d sbo=%d: %.1f cycles per MMA per CTA (%.1f per SM-level MMA)  [%s]\n", N, ctas_per_sm, same_a, shift_rows, sbo,
         (double)mx / (iters * 4), (double)mx / (iters * 4) / ctas_per_sm, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int c = 1; c <= 2; ++c) {
    run<32>(c, 0); run<64>(c, 0); run<128>(c, 0); run<256>(c, 0);
  }
  run<128>(1, 1); run<256>(1, 1);
  run<32>(1, 0, 0, 1280); run<32>(1, 0, 1, 1024); run<32>(1, 0, 1, 1280); run<32>(1, 0, 11, 1280); run<128>(1, 0, 11, 1280); run<64>(1, 0, 12, 1280);
  return 0;
}
