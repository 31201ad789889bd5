// Micro-probe (scratch, not shipped): (1) cycles per SS-mode tcgen05.mma (M=128, K=16, bf16) as a function of N;
// (2) does a tcgen05.ld issued by another warp wait behind MMAs already queued in the tensor pipe?
// (3) latency from the end of the last MMA to tcgen05.commit's mbarrier arrival becoming visible.
#include <cstdio>
#include <cuda_runtime.h>
#include "../iris-style-transfer_b200/csrc/isx_common.cuh"
using namespace isx;

__device__ __forceinline__ void umma_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
struct Out { long long mma_cycles; long long ld_start, ld_end, mma_start, mma_end; };

template <int N>
__global__ void __launch_bounds__(384, 1) probe(int n_mma, int do_ld, int ld_delay, int spin_mode, Out* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t spin_bar;
  __shared__ uint32_t tptr;
  __shared__ long long t_ld0, t_ld1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 9 * 8192 + 65536) / 4; i += 384) reinterpret_cast<uint32_t*>(smem)[i] = (spin_mode & 0x100000) ? (((i * 2654435761u) >> 3) & 0x3fff3fffu) | 0x30003000u : 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&spin_bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc<512>(&tptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tptr;
  const uint32_t a_addr = smem_u32(smem + 16384 + 9 * 8192), b_addr = smem_u32(smem + 16384);
  constexpr uint32_t idesc = umma_idesc_bf16(128, N, false, false);
  long long t0 = 0, t1 = 0;
  if (warp == 0 && lane == 0) {
    t0 = clock64();
    if (ld_delay == -2) {  // emulate the conv tap pattern: 9 shifted A views of a halo patch x 4 k-steps, 9 weight slabs
      const uint32_t slab = (N < 64 ? 16 : 64) * 128;
      for (int i = 0; i < n_mma; i += 36) {
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap - ky * 3;
          const uint32_t aa = a_addr + (ky * 16 + kx) * 128, bb = b_addr + ((spin_mode & 0x200) ? 0 : tap * slab);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tm, umma_desc_sw128(aa + k * 32, 16, 2048), umma_desc_sw128(bb + k * 32, 16, 1024), idesc, (i | tap | k) ? 1u : 0u);
        }
      }
    } else if (ld_delay == -5 || ld_delay == -6) {  // conv tap pattern, fully unrolled, descriptor low words = base + immediate
      const uint32_t slab = (N < 64 ? 16 : 64) * 128;
      const uint64_t da0 = umma_desc_sw128(a_addr, 16, 2048), db0 = umma_desc_sw128(b_addr, 16, 1024);
      const uint32_t a_lo = (uint32_t)da0, a_hi = (uint32_t)(da0 >> 32), b_lo = (uint32_t)db0, b_hi = (uint32_t)(db0 >> 32);
      const uint32_t tmx = ld_delay == -6 ? 0u : tm;   // -6: accumulator address is a compile-time constant (no waterfall)
#pragma unroll 1
      for (int i = 0; i < n_mma; i += 36) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap - ky * 3;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_lo(tmx, a_lo + (((ky * 16 + kx) * 128 + k * 32) >> 4), a_hi, b_lo + ((tap * slab + k * 32) >> 4), b_hi, idesc, (tap | k) ? 1u : (i ? 1u : 0u));
        }
      }
    } else if (ld_delay == -3) {  // A fixed, only B varies per tap
      const uint32_t slab = (N < 64 ? 16 : 64) * 128;
      for (int i = 0; i < n_mma; i += 36) {
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t aa = a_addr, bb = b_addr + tap * slab;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tm, umma_desc_sw128(aa + k * 32, 16, 2048), umma_desc_sw128(bb + k * 32, 16, 1024), idesc, (i | tap | k) ? 1u : 0u);
        }
      }
    } else if (ld_delay == -1) {  // fast issue: descriptors precomputed, loop unrolled by 4
      uint64_t ad[4], bd[4];
      const uint32_t sbo = (spin_mode & 0x100) ? 2048 : 1024;
      const uint32_t a_off = ((spin_mode >> 12) & 0xff) * 128;
      for (int k = 0; k < 4; ++k) { ad[k] = umma_desc_sw128(a_addr + a_off + k * 32, 16, sbo); bd[k] = umma_desc_sw128(b_addr + k * 32, 16, 1024); }
      umma_bf16(tm, ad[0], bd[0], idesc, 0u);
#pragma unroll 1
      for (int i = 0; i < n_mma; i += 4) {
        umma_bf16(tm, ad[0], bd[0], idesc, 1u);
        umma_bf16(tm, ad[1], bd[1], idesc, 1u);
        umma_bf16(tm, ad[2], bd[2], idesc, 1u);
        umma_bf16(tm, ad[3], bd[3], idesc, 1u);
      }
    } else
    for (int i = 0; i < n_mma; ++i) {
      const int k = i & 3;
      umma_bf16(tm, umma_desc_sw128(a_addr + k * 32, 16, 1024), umma_desc_sw128(b_addr + k * 32, 16, 1024), idesc, i ? 1u : 0u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    t1 = clock64();
    mbar_arrive(&spin_bar);
  } else if (warp == 1 && do_ld) {
    // read the OTHER accumulator set (columns 256..) while the MMAs are queued
    long long s = clock64();
    while (clock64() - s < ld_delay) {}
    uint32_t v[32];
    const long long l0 = clock64();
    tmem_ld_32x32(tm + 256 + (32u << 16), v);
    tmem_ld_wait();
    const long long l1 = clock64();
    uint32_t acc = 0;
    for (int i = 0; i < 32; ++i) acc ^= v[i];
    if (lane == 0) { t_ld0 = l0; t_ld1 = l1; if (acc == 0x12345678u) printf("x"); }
  }
  if (warp >= 4 && spin_mode < 16) {
    if (spin_mode == 1) mbar_wait(&spin_bar, 0);                 // all 256 threads poll
    if (spin_mode == 2) { if (lane == 0) mbar_wait(&spin_bar, 0); __syncwarp(); }  // one lane per warp polls
    if (spin_mode == 3) { while (!mbar_try_wait(&spin_bar, 0)) __nanosleep(64); }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    out->mma_cycles = t1 - t0; out->mma_start = t0; out->mma_end = t1; out->ld_start = t_ld0; out->ld_end = t_ld1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int N>
void run(const char* name) {
  Out* d; cudaMalloc(&d, sizeof(Out));
  auto k = probe<N>;
  const int smem = 1024 + 16384 + 9 * 8192 + 65536;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  Out a, b;
  k<<<1, 384, smem>>>(64, 0, 0, 0, d); cudaDeviceSynchronize();
  k<<<1, 384, smem>>>(64, 0, 0, 0, d); cudaMemcpy(&a, d, sizeof(Out), cudaMemcpyDeviceToHost);
  k<<<1, 384, smem>>>(1024 + 64, 0, 0, 0, d); cudaMemcpy(&b, d, sizeof(Out), cudaMemcpyDeviceToHost);
  printf("%s: %.1f cycles per MMA (64 MMAs: %lld cycles, 1088: %lld); err=%s\n", name,
         (b.mma_cycles - a.mma_cycles) / 1024.0, a.mma_cycles, b.mma_cycles, cudaGetErrorString(cudaGetLastError()));
  // commit latency: 4 MMAs vs 8 MMAs -> intercept
  Out c4, c8;
  k<<<1, 384, smem>>>(4, 0, 0, 0, d); cudaMemcpy(&c4, d, sizeof(Out), cudaMemcpyDeviceToHost);
  k<<<1, 384, smem>>>(8, 0, 0, 0, d); cudaMemcpy(&c8, d, sizeof(Out), cudaMemcpyDeviceToHost);
  printf("   4 MMAs + commit + wait: %lld cycles, 8 MMAs: %lld\n", c4.mma_cycles, c8.mma_cycles);
  for (int delay : {200, 2000}) {
    k<<<1, 384, smem>>>(1024, 1, delay, 0, d); cudaMemcpy(&b, d, sizeof(Out), cudaMemcpyDeviceToHost);
    printf("   ld issued %lld cycles after MMA start, took %lld cycles; the 1024 MMAs ended %lld cycles after start\n",
           b.ld_start - b.mma_start, b.ld_end - b.ld_start, b.mma_end - b.mma_start);
  }
  k<<<1, 384, smem>>>(0, 1, 200, 0, d); cudaMemcpy(&b, d, sizeof(Out), cudaMemcpyDeviceToHost);
  printf("   ld with idle tensor pipe took %lld cycles\n", b.ld_end - b.ld_start);
  for (int mode = 0; mode < 4; ++mode) {
    k<<<1, 384, smem>>>(64, 0, 0, mode, d); cudaMemcpy(&a, d, sizeof(Out), cudaMemcpyDeviceToHost);
    k<<<1, 384, smem>>>(1088, 0, 0, mode, d); cudaMemcpy(&b, d, sizeof(Out), cudaMemcpyDeviceToHost);
    printf("   spin mode %d (0 none, 1 all 256 threads, 2 lane 0 of 8 warps, 3 all + nanosleep): %.1f cycles per MMA\n", mode, (b.mma_cycles - a.mma_cycles) / 1024.0);
  }
  k<<<1, 384, smem>>>(64, 0, -1, 0, d); cudaMemcpy(&a, d, sizeof(Out), cudaMemcpyDeviceToHost);
  k<<<1, 384, smem>>>(1088, 0, -1, 0, d); cudaMemcpy(&b, d, sizeof(Out), cudaMemcpyDeviceToHost);
  printf("   fast issue: %.1f cycles per MMA\n", (b.mma_cycles - a.mma_cycles) / 1024.0);
  for (int mode : {0x100, 0x100100}) {
    k<<<1, 384, smem>>>(64, 0, -1, mode, d); cudaMemcpy(&a, d, sizeof(Out), cudaMemcpyDeviceToHost);
    k<<<1, 384, smem>>>(1088, 0, -1, mode, d); cudaMemcpy(&b, d, sizeof(Out), cudaMemcpyDeviceToHost);
    printf("   data %s: %.1f cycles per MMA\n", (mode & 0x100000) ? "random" : "constant", (b.mma_cycles - a.mma_cycles) / 1024.0);
  }
  if (N <= 64) for (int pat : {-2, -3, -4, -5, -6}) {
    const int mode = pat == -4 ? 0x200 : 0;
    const int dl = pat == -4 ? -2 : pat;
    k<<<1, 384, smem>>>(72, 0, dl, mode, d); cudaMemcpy(&a, d, sizeof(Out), cudaMemcpyDeviceToHost);
    k<<<1, 384, smem>>>(72 + 1152, 0, dl, mode, d); cudaMemcpy(&b, d, sizeof(Out), cudaMemcpyDeviceToHost);
    printf("   pattern %s: %.1f cycles per MMA\n", pat == -2 ? "conv taps (A shifted views, B slab per tap)" : pat == -3 ? "A fixed, B slab per tap" : pat == -4 ? "A shifted views, B fixed" : pat == -5 ? "conv taps unrolled, lo-word immediates" : "conv taps unrolled, lo-word immediates, constant tmem address", (b.mma_cycles - a.mma_cycles) / 1152.0);
  }
  cudaFree(d);
}

int main() {
  run<16>("N=16");
  run<64>("N=64");
  run<128>("N=128");
  run<256>("N=256");
  return 0;
}
