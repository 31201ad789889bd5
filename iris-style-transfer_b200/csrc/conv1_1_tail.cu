// Image-gradient tail (conv1_1 dgrad, 64 -> 3 channels) with the nine taps moved from K into N.
//
// The generic formulation (conv_tc.cu, EPI = 1) runs 9 taps x 4 k-steps = 36 MMAs of N = 16 per 128 pixels; an
// M = 128 MMA costs ~49 cycles however small N is (profiles/r01_umma_issue_probe.txt), so that kernel is bound by
// MMA issue (21 us/img at 640x400 for 0.9 GFLOP of useful work).  Here one GEMM per pixel computes ALL tap products
//     E[q][(tap, c)] = sum_o dY[q][o] * wd[tap][c][o]          (M = pixels, N = 27 -> 32, K = 64: 4 MMAs per 128 pixels)
// and the 3x3 gather  dX[y][x][c] = sum_tap E[(y + ky - 1, x + kx - 1)][(tap, c)]  happens in the epilogue through
// shared memory.  A work item is a 16 x 16 pixel patch (one TMA box, two 128-row M tiles) whose 14 x 14 interior is
// written; patches overlap by one pixel on every side (1.31x re-read, mostly from L2).  Persistent CTAs, two per SM (the
// per-patch chain TMEM read -> E -> gather -> store is latency-bound: a second CTA fills its bubbles), each with a ring
// of two patch slots and two TMEM accumulator sets.
// HBM-bound by design: 128 B/pixel of dY read + 12 B/pixel of dX written.
#include <algorithm>

#include "isx_common.cuh"
#include "isx_internal.h"
#include "isx_kernels.h"

namespace isx {

static constexpr int kTailThreads = 64 + 256;
static constexpr int kTailPatch = 256 * 128;   // 16 x 16 pixels x 64 bf16
static constexpr int kTailSlots = 2;        // per CTA; two CTAs share an SM
static constexpr int kTailInner = 14;          // interior of a patch
static constexpr int kTailE = 27 * 256 * 4;    // E^T[27][256] fp32
static constexpr int kTailEPad = ((kTailE + 1023) / 1024) * 1024;

struct TailParams {
  int B, H, W;
  int tiles_x, tiles_y;
  int total_items;
  const __nv_bfloat16* wd;  // packed [9][16][64] (isx_pack_conv1_1_dgrad)
  float* dx;                // [B, xc, H, W] fp32
  int xc;
  const float* in_mask;     // [mask_b, 1, H, W] or null
  int mask_b;
};

__global__ void __launch_bounds__(kTailThreads, 2)
conv1_1_tail_kernel(const __grid_constant__ CUtensorMap tmA, const TailParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_patch = smem;                                  // kTailSlots x 32 KB
  uint8_t* s_w = smem + kTailSlots * kTailPatch;            // 32 rows x 128 B, SWIZZLE_128B
  float* s_e = reinterpret_cast<float*>(s_w + 4096);        // E^T[27][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + 4096 + kTailEPad);
  uint64_t* full = bars;             // [4]
  uint64_t* empty = bars + 4;        // [4]
  uint64_t* tmem_full = bars + 8;    // [2]
  uint64_t* tmem_empty = bars + 10;  // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // B operand: rows j = tap * 3 + c (27 real, 5 zero), 64 output channels of conv1_1 along K, 128-byte swizzled rows
  for (int i = threadIdx.x; i < 32 * 8; i += kTailThreads) {
    const int j = i >> 3, chunk = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (j < 27) {
      const int tap = j / 3, c = j - tap * 3;
      v = *reinterpret_cast<const uint4*>(p.wd + (tap * 16 + c) * 64 + chunk * 8);
    }
    *reinterpret_cast<uint4*>(s_w + j * 128 + ((chunk ^ (j & 7)) * 16)) = v;
  }
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    for (int i = 0; i < kTailSlots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int per_img = p.tiles_x * p.tiles_y;
  const int n_my = (p.total_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ================================ TMA producer =========================================
    if (lane == 0) {
      int b = static_cast<int>(blockIdx.x) / per_img, r = static_cast<int>(blockIdx.x) % per_img;
      uint32_t s = 0, ph = 0;
      for (int i = 0; i < n_my; ++i) {
        const int x0 = (r % p.tiles_x) * kTailInner, y0 = (r / p.tiles_x) * kTailInner;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], kTailPatch);
        tma_load_4d(s_patch + s * kTailPatch, &tmA, &full[s], 0, x0 - 1, y0 - 1, b);
        if (++s == kTailSlots) { s = 0; ph ^= 1; }
        r += gridDim.x;
        while (r >= per_img) { r -= per_img; ++b; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ===========================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 32, false, false);
      const uint64_t da = umma_desc_sw128(smem_u32(s_patch), 16, 1024);
      const uint64_t dw = umma_desc_sw128(smem_u32(s_w), 16, 1024);
      const uint32_t a_lo0 = static_cast<uint32_t>(da), hi = static_cast<uint32_t>(da >> 32);
      const uint32_t w_lo = static_cast<uint32_t>(dw);
      uint32_t s = 0, ph = 0, acc = 0, aph = 0, a_lo = a_lo0;
      bool t_ready = false, f_ready = false;
      for (int i = 0; i < n_my; ++i) {
        if (!t_ready) mbar_wait(&tmem_empty[acc], aph ^ 1);
        if (!f_ready) mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t d_tm = tmem_base + acc * 64;
        const uint32_t a_cur = a_lo;
        const uint32_t s_cur = s;
        a_lo += kTailPatch >> 4;
        if (++s == kTailSlots) { s = 0; ph ^= 1; a_lo = a_lo0; }
        t_ready = mbar_try_wait(&tmem_empty[acc ^ 1], acc == 1 ? aph : aph ^ 1);
        f_ready = mbar_try_wait(&full[s], ph);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_lohi(d_tm + mt * 32, a_cur + ((mt * 16384 + k * 32) >> 4), hi, w_lo + 2 * k, hi, idesc, k != 0 ? 1u : 0u);
        }
        umma_commit(&empty[s_cur]);
        umma_commit(&tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) aph ^= 1;
      }
    }
  } else {
    // ================================ epilogue (8 warps) ===================================
    const int q = warp & 3;            // TMEM lane quadrant
    const int mt = (warp - 2) >> 2;    // which 128-row half of the patch this warp reads
    const int prow = mt * 128 + q * 32 + lane;  // patch pixel index: (py, px) = (prow / 16, prow % 16)
    // gather role: 16 threads per patch row (14 active) -- a warp then reads words [r*16 + 0..13] and [(r+1)*16 + 0..13] of an
    // E row: 28 distinct banks (14 threads per row would wrap a third row onto the first one's banks: 29 % conflict cycles)
    const int t = threadIdx.x - 64;
    const int ly = t >> 4, lx = t & 15;
    const size_t hw = static_cast<size_t>(p.H) * p.W;
    int b = static_cast<int>(blockIdx.x) / per_img, r = static_cast<int>(blockIdx.x) % per_img;
    uint32_t acc = 0, aph = 0;
    for (int i = 0; i < n_my; ++i) {
      const int x0 = (r % p.tiles_x) * kTailInner, y0 = (r / p.tiles_x) * kTailInner;
      float* e = s_e;
      mbar_wait(&tmem_full[acc], aph);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + acc * 64 + mt * 32 + (static_cast<uint32_t>(q * 32) << 16), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
#pragma unroll
      for (int j = 0; j < 27; ++j) e[j * 256 + prow] = __uint_as_float(v[j]);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (lx < kTailInner && ly < kTailInner) {
        const int x = x0 + lx, y = y0 + ly;
        if (x < p.W && y < p.H) {
          float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap - ky * 3;
            const float* src = e + (tap * 3) * 256 + (ly + ky) * 16 + lx + kx;
            a0 += src[0];
            a1 += src[256];
            a2 += src[512];
          }
          const size_t off = static_cast<size_t>(y) * p.W + x;
          float m = 1.f;
          if (p.in_mask != nullptr) m = __ldg(p.in_mask + (p.mask_b > 1 ? b : 0) * hw + off);
          a0 = a0 * m / 0.229f;
          a1 = a1 * m / 0.224f;
          a2 = a2 * m / 0.225f;
          if (p.xc == 3) {
            float* o = p.dx + static_cast<size_t>(b) * 3 * hw + off;
            o[0] = a0; o[hw] = a1; o[2 * hw] = a2;
          } else {
            p.dx[static_cast<size_t>(b) * hw + off] = a0 + a1 + a2;
          }
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // every gather of this item is done before E is rewritten
      acc ^= 1;
      if (acc == 0) aph ^= 1;
      r += gridDim.x;
      while (r >= per_img) { r -= per_img; ++b; }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}


int conv1_1_tail_n(const __nv_bfloat16* dy, const __nv_bfloat16* wd, const float* mask, int mask_b, float* dx, int xc,
                   int B, int H, int W, cudaStream_t stream) {
  ISX_REQUIRE(xc == 1 || xc == 3, "conv1_1_tail_n: xc must be 1 or 3");
  TailParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.W = W;
  p.tiles_x = (W + kTailInner - 1) / kTailInner;
  p.tiles_y = (H + kTailInner - 1) / kTailInner;
  const long total = static_cast<long>(p.tiles_x) * p.tiles_y * B;
  ISX_REQUIRE(total < (1L << 31) - kNumSMs, "conv1_1_tail_n: too many patches");
  p.total_items = static_cast<int>(total);
  p.wd = wd; p.dx = dx; p.xc = xc; p.in_mask = mask; p.mask_b = mask ? mask_b : 0;
  CUtensorMap tmA;
  {
    uint64_t dims[4] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128};
    uint32_t box[4] = {64, 16, 16, 1};
    if (isx_make_tmap_bf16(&tmA, dy, 4, dims, str, box, true)) return 3;
  }
  const size_t smem_bytes = 1024 + kTailSlots * kTailPatch + 4096 + kTailEPad + 256;  // ~98 KB: two CTAs per SM
  ISX_CHECK_CUDA(cudaFuncSetAttribute(conv1_1_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  const int grid = std::min<int>(p.total_items, 2 * kNumSMs);
  isx_prof_begin(ISX_PROF_CONV, 2.0 * 27 * 64 * static_cast<double>(B) * H * W, stream);
  conv1_1_tail_kernel<<<(unsigned)grid, kTailThreads, smem_bytes, stream>>>(tmA, p);
  isx_prof_end(ISX_PROF_CONV, stream);
  ISX_LAUNCH_CHECK();
  return 0;
}

}  // namespace isx
