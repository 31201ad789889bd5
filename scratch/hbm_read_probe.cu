// Achievable HBM READ bandwidth on this B200 for a streaming reduction (what the L-BFGS history passes do):
// 128-bit vs 256-bit loads, default vs L1::no_allocate / L2 evict_first hints, loads in flight per thread.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE, int UNROLL>
__global__ void __launch_bounds__(256) rd(const float4* __restrict__ p, size_t n4, float* out) {
  float acc = 0.f;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (MODE == 2) {  // 256-bit loads
    const size_t n8 = n4 / 2;
    for (; i + (UNROLL - 1) * stride < n8; i += UNROLL * stride) {
      float v[UNROLL][8];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const float* a = reinterpret_cast<const float*>(p) + (i + u * stride) * 8;
        unsigned w[8];
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                     : "l"(a));
#pragma unroll
        for (int j = 0; j < 8; ++j) v[u][j] = __uint_as_float(w[j]);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += v[u][j];
    }
  } else {
    for (; i + (UNROLL - 1) * stride < n4; i += UNROLL * stride) {
      float4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const float4* a = p + i + u * stride;
        if (MODE == 0) v[u] = __ldg(a);
        else asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                          : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(a));
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int MODE, int UNROLL>
float run(const float4* p, size_t n4, float* out, int blocks) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rd<MODE, UNROLL><<<blocks, 256>>>(p, n4, out);
  cudaEventRecord(e0);
  for (int r = 0; r < 3; ++r) rd<MODE, UNROLL><<<blocks, 256>>>(p, n4, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 3;
}

int main() {
  const size_t bytes = 32ull << 30;
  float4* p; float* out;
  CK(cudaMalloc(&p, bytes)); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(p, 0, bytes));
  const size_t n4 = bytes / 16;
  for (int mult : {4, 8, 16}) {
    const int blocks = 148 * mult;
    printf("blocks %5d | ldg128 x4 %.0f  x8 %.0f | no_alloc x4 %.0f  x8 %.0f | ld256+evict_first x2 %.0f  x4 %.0f  GB/s\n", blocks,
           bytes / run<0, 4>(p, n4, out, blocks) / 1e6, bytes / run<0, 8>(p, n4, out, blocks) / 1e6,
           bytes / run<1, 4>(p, n4, out, blocks) / 1e6, bytes / run<1, 8>(p, n4, out, blocks) / 1e6,
           bytes / run<2, 2>(p, n4, out, blocks) / 1e6, bytes / run<2, 4>(p, n4, out, blocks) / 1e6);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
