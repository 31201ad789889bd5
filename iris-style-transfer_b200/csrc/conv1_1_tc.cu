// K0 + K1 head on the tensor cores: Normalize -> [* mask] -> conv1_1 (3 -> 64, K = 27) + bias + ReLU
// (models/vgg/vgg.py:81-87).  The 27-tap patch of every pixel is gathered by the pixel's own thread from the
// fp32 NCHW image, normalised, and split into bf16 hi + lo parts (16 mantissa bits survive), giving one
// 128-byte K-major row [hi(27) | lo(27) | 0(10)] of the A tile in tcgen05's SWIZZLE_128B layout; the weight
// slab [64][w(27) | w(27) | 0] is loaded once per CTA by TMA.  One 128x64x64 MMA per 128 pixels, fp32
// accumulation in TMEM, bias + ReLU + bf16 in the epilogue, TMA store of the NHWC rows.
#include <algorithm>

#include "isx_common.cuh"
#include "isx_kernels.h"

namespace isx {

__global__ void pack_w0_fwd_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // [64][64]
  if (i >= 64 * 64) return;
  const int k = i % 64, o = i / 64;
  float v = 0.f;
  if (k < 27) v = w[o * 27 + k];
  else if (k < 54) v = w[o * 27 + k - 27];
  wp[i] = __float2bfloat16_rn(v);
}

int pack_w0_fwd(const float* w, __nv_bfloat16* wp, cudaStream_t s) {
  pack_w0_fwd_kernel<<<16, 256, 0, s>>>(w, wp);
  ISX_LAUNCH_CHECK();
  return 0;
}

struct C11Params {
  const float* x;
  const float* mask;
  const float* bias;
  int xc, mask_b, B, H, W;
  long npix;
  int n_tiles;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}

__global__ void __launch_bounds__(128)
conv1_1_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO, const C11Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 128 rows x 128 B
  uint8_t* sB = smem + 16384;         // 64 rows x 128 B
  uint8_t* sO = smem + 16384 + 8192;  // 128 rows x 128 B staging
  uint64_t* w_bar = reinterpret_cast<uint64_t*>(smem + 16384 + 8192 + 16384);
  uint64_t* mma_bar = w_bar + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(mma_bar + 1);
  __shared__ float s_bias[64];
  const int tid = threadIdx.x, warp = tid >> 5;
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};

  if (tid == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmO);
    mbar_init(w_bar, 1);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (tid < 64) s_bias[tid] = p.bias ? p.bias[tid] : 0.f;
  if (warp == 0) tmem_alloc<64>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (tid == 0) {
    mbar_arrive_expect_tx(w_bar, 8192);
    tma_load_2d(sB, &tmW, w_bar, 0, 0);
  }
  const long hw = static_cast<long>(p.H) * p.W;
  constexpr uint32_t idesc = umma_idesc_bf16(128, 64, false, false);
  uint32_t phase = 0;
  bool w_ready = false;
  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    // ---- gather + normalise + split this thread's pixel ----
    const long pix = static_cast<long>(tile) * 128 + tid;
    uint32_t packed[32];  // 64 bf16: hi[27], lo[27], zero[10]
    {
      float v[27];
      if (pix < p.npix) {
        const int xx = static_cast<int>(pix % p.W);
        const int yy = static_cast<int>((pix / p.W) % p.H);
        const int b = static_cast<int>(pix / hw);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float* xp = p.x + (static_cast<long>(b) * p.xc + (p.xc == 3 ? c : 0)) * hw;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int y = yy + ky - 1, xq = xx + kx - 1;
              float t = 0.f;
              if (y >= 0 && y < p.H && xq >= 0 && xq < p.W) {
                t = (__ldg(xp + static_cast<long>(y) * p.W + xq) - mean[c]) / stdv[c];
                if (p.mask) t *= __ldg(p.mask + (static_cast<long>(p.mask_b > 1 ? b : 0) * p.H + y) * p.W + xq);
              }
              v[c * 9 + ky * 3 + kx] = t;
            }
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 27; ++k) v[k] = 0.f;
      }
      __nv_bfloat16 h[64];
#pragma unroll
      for (int k = 0; k < 27; ++k) {
        h[k] = __float2bfloat16_rn(v[k]);
        h[27 + k] = __float2bfloat16_rn(v[k] - __bfloat162float(h[k]));
      }
#pragma unroll
      for (int k = 54; k < 64; ++k) h[k] = __float2bfloat16_rn(0.f);
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        __nv_bfloat162 t2;
        t2.x = h[2 * k];
        t2.y = h[2 * k + 1];
        packed[k] = *reinterpret_cast<uint32_t*>(&t2);
      }
    }
    uint8_t* rowp = sA + tid * 128;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      *reinterpret_cast<uint4*>(rowp + ((c ^ (tid & 7)) * 16)) =
          make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      if (!w_ready) mbar_wait(w_bar, 0);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_base, umma_desc_sw128(a_addr + k * 32, 16, 1024), umma_desc_sw128(b_addr + k * 32, 16, 1024),
                  idesc, k != 0 ? 1u : 0u);
      umma_commit(mma_bar);
      tma_store_wait_read<0>();  // previous tile's staging buffer has been read by its TMA store
    }
    w_ready = true;
    mbar_wait(mma_bar, phase);
    phase ^= 1;
    tc_fence_after();
    __syncthreads();  // staging free (thread 0 waited above)
    // ---- epilogue: bias + ReLU -> bf16 -> swizzled staging -> TMA store ----
#pragma unroll 1
    for (int hlf = 0; hlf < 2; ++hlf) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + hlf * 32, v);
      tmem_ld_wait();
      uint8_t* orow = sO + tid * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(__uint_as_float(v[c * 8 + j]) + s_bias[hlf * 32 + c * 8 + j], 0.f);
        uint4 o;
        o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
        o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
        *reinterpret_cast<uint4*>(orow + (((hlf * 4 + c) ^ (tid & 7)) * 16)) = o;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tma_store_2d(&tmO, sO, 0, tile * 128);
      tma_store_commit();
    }
  }
  if (tid == 0) tma_store_wait_all<0>();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<64>(tmem_base);
  }
}

int conv1_1_fwd_tc(const float* x, int xc, const float* mask, int mask_b, const __nv_bfloat16* w0_packed,
                   const float* bias, __nv_bfloat16* out, int B, int H, int W, cudaStream_t s) {
  ISX_REQUIRE(xc == 1 || xc == 3, "conv1_1: image must have 1 or 3 channels, got %d", xc);
  C11Params p;
  p.x = x; p.mask = mask; p.bias = bias; p.xc = xc; p.mask_b = mask_b; p.B = B; p.H = H; p.W = W;
  p.npix = static_cast<long>(B) * H * W;
  ISX_REQUIRE(p.npix < (1L << 31) - 256, "conv1_1: too many pixels for 32-bit TMA coordinates");
  p.n_tiles = static_cast<int>((p.npix + 127) / 128);
  CUtensorMap tmW, tmO;
  {
    uint64_t dims[2] = {64, 64};
    uint64_t str[1] = {128};
    uint32_t box[2] = {64, 64};
    if (isx_make_tmap_bf16(&tmW, w0_packed, 2, dims, str, box, true)) return 3;
  }
  {
    uint64_t dims[2] = {64, (uint64_t)p.npix};
    uint64_t str[1] = {128};
    uint32_t box[2] = {64, 128};
    if (isx_make_tmap_bf16(&tmO, out, 2, dims, str, box, true)) return 3;
  }
  const size_t smem_bytes = 1024 + 16384 + 8192 + 16384 + 64;
  ISX_CHECK_CUDA(cudaFuncSetAttribute(conv1_1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  const int grid = std::min(p.n_tiles, kNumSMs * 5);
  conv1_1_tc_kernel<<<grid, 128, smem_bytes, s>>>(tmW, tmO, p);
  ISX_LAUNCH_CHECK();
  return 0;
}

}  // namespace isx
