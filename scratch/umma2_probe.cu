// Micro-probe (scratch, not shipped): cycles per tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, K = 16, bf16) vs N.
#include <cstdio>
#include <cuda_runtime.h>
#include "../iris-style-transfer_b200/csrc/isx_common.cuh"
using namespace isx;

struct Out { long long cycles; unsigned check; };

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe2(int n_mma, Out* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  // A: 128 rows x 64 K (16 KB); B: N/2 rows x 64 K
  for (int i = threadIdx.x; i < (16384 + 128 * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = tptr;
  // idesc: M = 256 (m_dim = 16), N
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);
  long long t0 = 0, t1 = 0;
  if (rank == 0 && warp == 0 && lane == 0) {
    const uint64_t da = umma_desc_sw128(smem_u32(smem), 16, 1024), db = umma_desc_sw128(smem_u32(smem + 16384), 16, 1024);
    const uint32_t a_lo = (uint32_t)da, a_hi = (uint32_t)(da >> 32), b_lo = (uint32_t)db, b_hi = (uint32_t)(db >> 32);
    t0 = clock64();
    umma2(tm, a_lo, a_hi, b_lo, b_hi, idesc, 0u);
#pragma unroll 1
    for (int i = 0; i < n_mma; i += 4) {
      umma2(tm, a_lo, a_hi, b_lo, b_hi, idesc, 1u);
      umma2(tm, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
      umma2(tm, a_lo + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
      umma2(tm, a_lo + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
    }
    commit2(&bar, 3);
  }
  if (warp == 0 && lane == 0) {
    mbar_wait(&bar, 0);
    t1 = clock64();
  }
  tc_fence_after();
  __syncthreads();
  // read one accumulator value in each CTA (sanity: 64-term dot products of 0x3c00 bf16 = 0.0078125^2 * K)
  uint32_t v[32];
  tmem_ld_32x32(tm + (static_cast<uint32_t>(warp * 32) << 16), v);
  tmem_ld_wait();
  if (threadIdx.x == 0) { out[rank].cycles = t1 - t0; out[rank].check = v[0]; }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512) : "memory");
}

template <int N>
void run(const char* name) {
  Out* d; cudaMalloc(&d, 2 * sizeof(Out));
  auto k = probe2<N>;
  const int smem = 1024 + 16384 + 128 * 128;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  Out a[2], b[2];
  k<<<2, 128, smem>>>(64, d); cudaDeviceSynchronize();
  k<<<2, 128, smem>>>(64, d); cudaMemcpy(a, d, sizeof(a), cudaMemcpyDeviceToHost);
  k<<<2, 128, smem>>>(64 + 1024, d); cudaMemcpy(b, d, sizeof(b), cudaMemcpyDeviceToHost);
  printf("%s: %.1f cycles per cta_group::2 MMA (M=256)  [64: %lld, 1088: %lld cycles]  acc[0] = %g / %g  err=%s\n", name,
         (b[0].cycles - a[0].cycles) / 1024.0, a[0].cycles, b[0].cycles, *reinterpret_cast<float*>(&b[0].check),
         *reinterpret_cast<float*>(&b[1].check), cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  run<32>("N=32");
  run<64>("N=64");
  run<128>("N=128");
  run<256>("N=256");
  return 0;
}
