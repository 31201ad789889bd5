// Micro-probe (scratch): is the ~49-cycle floor of small-N tcgen05.mma per ISSUING THREAD or per SM?  Two warps issue
// independent MMA streams (separate accumulators, separate operands) concurrently.
#include <cstdio>
#include <cuda_runtime.h>
#include "../iris-style-transfer_b200/csrc/isx_common.cuh"
using namespace isx;

struct Out { long long cycles[2]; };

template <int N>
__global__ void __launch_bounds__(128, 1) probe3(int n_mma, int two, Out* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (2 * (16384 + 32768)) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc<512>(&tptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tptr;
  constexpr uint32_t idesc = umma_idesc_bf16(128, N, false, false);
  if (lane == 0 && (warp == 0 || (warp == 1 && two))) {
    uint8_t* base = smem + warp * (16384 + 32768);
    const uint64_t da = umma_desc_sw128(smem_u32(base), 16, 1024), db = umma_desc_sw128(smem_u32(base + 16384), 16, 1024);
    const uint32_t a_lo = (uint32_t)da, a_hi = (uint32_t)(da >> 32), b_lo = (uint32_t)db, b_hi = (uint32_t)(db >> 32);
    const uint32_t d = tm + warp * 256;
    const long long t0 = clock64();
    umma_bf16_lohi(d, a_lo, a_hi, b_lo, b_hi, idesc, 0u);
#pragma unroll 1
    for (int i = 0; i < n_mma; i += 4) {
      umma_bf16_lohi(d, a_lo, a_hi, b_lo, b_hi, idesc, 1u);
      umma_bf16_lohi(d, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
      umma_bf16_lohi(d, a_lo + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
      umma_bf16_lohi(d, a_lo + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
    }
    umma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    out->cycles[warp] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int N>
void run(const char* name) {
  Out* d; cudaMalloc(&d, sizeof(Out));
  auto k = probe3<N>;
  const int smem = 1024 + 2 * (16384 + 32768);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  Out a, b;
  for (int two = 0; two < 2; ++two) {
    k<<<1, 128, smem>>>(64, two, d); cudaDeviceSynchronize();
    k<<<1, 128, smem>>>(64, two, d); cudaMemcpy(&a, d, sizeof(Out), cudaMemcpyDeviceToHost);
    k<<<1, 128, smem>>>(64 + 1024, two, d); cudaMemcpy(&b, d, sizeof(Out), cudaMemcpyDeviceToHost);
    printf("%s, %d issuing thread(s): %.1f cycles per MMA per thread%s  err=%s\n", name, two + 1, (b.cycles[0] - a.cycles[0]) / 1024.0,
           two ? "  (two streams in flight: SM rate = twice that)" : "", cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(d);
}

int main() {
  run<16>("N=16");
  run<64>("N=64");
  run<128>("N=128");
  return 0;
}
