// Same experiment as desc_shift.cu for SWIZZLE_64B (64-byte rows = 32 bf16 channels): shifted windows with an arbitrary line
// pitch (the four parity planes of the thin "down" layer are 9 pixels x 64 B = 576 B per line).
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
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
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

__global__ void __launch_bounds__(128, 1) probe(float* out, int shift, int mode, int pitch_px) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 18 * 16 * 32; i += blockDim.x) {
    const int ch = i & 31, px = (i >> 5) & 15, line = i >> 9;
    if (px >= pitch_px) continue;
    const uint32_t off = (line * pitch_px + px) * 64;
    const uint32_t phys = off + ((((ch >> 3) ^ ((off >> 7) & 3)) << 4) | ((ch & 7) << 1));     // SWIZZLE_64B: bits [4:5] ^= bits [7:8]
    const float v = mode == 0 ? (float)(line * 16 + px) : (float)ch;
    *reinterpret_cast<__nv_bfloat16*>(smem + phys) = __float2bfloat16_rn(v);
  }
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int k = i & 31, n = i >> 5;
    const uint32_t off = n * 64;
    const uint32_t phys = off + ((((k >> 3) ^ ((off >> 7) & 3)) << 4) | ((k & 7) << 1));
    *reinterpret_cast<__nv_bfloat16*>(smem + 48 * 1024 + phys) = __float2bfloat16_rn(k == n ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = slot;
  if (warp == 1 && lane == 0) {
    const uint32_t sa = smem_u32(smem) + shift * 64, sb = smem_u32(smem) + 48 * 1024;
    for (int k = 0; k < 2; ++k)
      tcgen05_mma_f16(tm, desc(sa + k * 32, 16, pitch_px * 64, 4), desc(sb + k * 32, 16, 512, 4), idesc_bf16(128, 32), k != 0);
    tcgen05_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  uint32_t r[32];
  tcgen05_ld_32x32b_x32(tm + ((uint32_t)(warp * 32) << 16), r);
  tcgen05_wait_ld();
  for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = __uint_as_float(r[j]);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(32));
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 32 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  static float h[128 * 32];
  for (int pitch = 16; pitch >= 9; pitch -= 7)
    for (int shift = 0; shift < 3; ++shift) {
      int bad0 = 0, bad1 = 0;
      for (int mode = 0; mode < 2; ++mode) {
        probe<<<1, 128, 64 * 1024>>>(d, shift, mode, pitch);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 32; ++n) {
            const float want = mode == 0 ? (float)((m / 8) * 16 + (m % 8) + shift) : (float)n;
            if (h[m * 32 + n] != want) (mode == 0 ? bad0 : bad1)++;
          }
      }
      const bool in_line = 8 + shift <= pitch;
      printf("SW64 pitch %2d px, shift %d: wrong pixel rows %5d / 4096, wrong channel order %5d / 4096%s\n", pitch, shift, bad0, bad1,
             (bad0 == 0 && bad1 == 0) ? "   <== exact" : (in_line ? "" : "   (window leaves the line: expected)"));
    }
  return 0;
}
