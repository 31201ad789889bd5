// K4: Gram matrix F^T F of an NHWC bf16 feature map, per image, on tcgen05 tensor cores -- SYMMETRIC schedule,
// optional per-pixel mask weights applied to the operand tile in shared memory (row G' of SURVEY.md §8a).
// Reference: utils.py:242-257 (GramMatrix: flatten(H,W); x @ x^T / n), called from StyleLoss_Gram (utils.py:305,319).
//
// In NHWC the feature map of one image is a [HW x C] matrix with the channel contiguous, i.e. both GEMM operands of
// G = F^T F are "MN-major" (M/N index contiguous, the reduction index = pixel strided).  tcgen05 consumes that layout
// directly (a_major = b_major = MN in the instruction descriptor), so the tiles TMA brings in -- [KP pixels] x
// [64 channels] boxes, 128-byte rows, SWIZZLE_128B -- serve as BOTH operands.
//
// Symmetric contraction: G is symmetric, so only the block-upper triangle (128-channel granularity) is computed.  The
// work of one image is cut into UNITS = (128-row block A, up to 256 columns B at or right of the diagonal):
//     C = 512:  (r0 x c0-255) (r0 x c256-511) (r1 x c128-383) (r1 x c384-511) (r2 x c256-511) (r3 x c384-511)
//     C = 256:  (r0 x c0-255) (r1 x c128-255)          C = 128 / 64: one unit
// = 10 of the 16 128x128 tiles for C = 512 and 3 of 4 for C = 256; gram_finalize mirrors the rest.  One MMA per 16 pixels
// and unit (N <= 256), at most 256 TMEM columns per CTA so that two CTAs share an SM (one drains while the other
// multiplies).  Split-K over pixels; the fp32 partials are reduced deterministically by gram_finalize (elementwise.cu),
// fused with 1/n, (G - T), the loss, the bf16 dL/dG matrix and the upper-triangle feature row.
//
// Mask weights (MASKED): GramMatrix(F * m_l) with m_l a per-pixel weight.  Four extra warps scale every landed tile
// row by m[pixel] IN PLACE before the MMA reads it (bf16(F*m): bit-identical to a separate pre-pass) and write
// bf16(F*m^2) -- the operand of the Gram backward, dF = (F m^2) . D -- to global memory; K blocks whose mask is all
// zero (an iris mask covers 6-10 % of an eye frame) are neither loaded nor multiplied.
#include <algorithm>

#include "isx_common.cuh"
#include "isx_internal.h"

namespace isx {

struct GramUnit {
  int a_blk;        // first 64-channel block of the 128 output rows
  int b_blk, nb;    // first 64-channel block and number of blocks of the output columns (N = 64 * nb)
  int a_slot, b_slot, nload;
  int load_blk[6];  // 64-channel blocks to load per stage, in smem slot order
  int write_fm2;    // MASKED: this unit writes F*m^2 for its A blocks (each block written exactly once per image)
};

struct GramParams {
  int B, HW, C;
  int splits;
  int chunk;   // pixels per split (multiple of KP)
  int stages;
  int stage_bytes;
  int n_units;
  GramUnit units[6];
  float* partial;  // [B][splits][C][C]
  float* csum;     // optional [B][splits][C]: per-channel sums of the (weighted) features, C <= 128 only -- one extra N = 16
                   // MMA per 16 pixels against a constant tile whose column 0 is 1 (the tensor core adds the channel up)
  // MASKED only
  const float* mask;          // [mask_b][HW]
  int mask_b;
  const uint8_t* kb_flags;    // [mask_b][ceil(HW / KP)]: K block has a non-zero mask value
  int n_kb_total;
  __nv_bfloat16* fm2;         // [B][HW][C] = F * m^2 (zero wherever the mask is zero: caller zeroes once)
};

template <int KP, bool MASKED>
__global__ void __launch_bounds__(MASKED ? 320 : 192, 1)
gram_sym_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ GramParams p) {
  constexpr int kBlkBytes = KP * 128;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = p.stages;
  uint8_t* ones = smem + stages * p.stage_bytes;                       // [KP][128 B], only when p.csum
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ones + (p.csum ? kBlkBytes : 0));
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* xf_bar = empty_bar + stages;
  uint64_t* tmem_full_bar = xf_bar + stages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int unit_id = blockIdx.x % p.n_units;
  const int split = (blockIdx.x / p.n_units) % p.splits;
  const int b = blockIdx.x / (p.n_units * p.splits);
  const GramUnit& u = p.units[unit_id];
  const int N = u.nb * 64;

  const int p_begin = split * p.chunk;
  const int p_end = min(p.HW, p_begin + p.chunk);
  const int num_kb = p_end > p_begin ? (p_end - p_begin + KP - 1) / KP : 0;
  const uint8_t* flags = MASKED ? p.kb_flags + static_cast<long>(p.mask_b > 1 ? b : 0) * p.n_kb_total + p_begin / KP : nullptr;
  auto active = [&](int kb) -> bool { return !MASKED || flags[kb] != 0; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmF);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&xf_bar[s], 4);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  const int ncols = p.csum ? (N == 64 ? 128 : 256) : N;   // + 16 columns for the channel sums, power of two
  if (warp == 1) {
    if (ncols == 256) tmem_alloc<256>(tmem_ptr_smem);
    else if (ncols == 128) tmem_alloc<128>(tmem_ptr_smem);
    else tmem_alloc<64>(tmem_ptr_smem);
  }
  if (p.csum && warp >= 2 && warp < 6) {  // constant B tile: element (pixel r, channel 0) = 1, MN-major SWIZZLE_128B
    const int t = threadIdx.x - 64;
    for (int i = t; i < KP * 8; i += 128) {
      const int r = i >> 3, c16 = i & 7;
      const uint32_t one = (c16 == (r & 7)) ? 0x00003F80u : 0u;   // logical chunk 0 sits at physical chunk r % 8
      *reinterpret_cast<uint4*>(ones + r * 128 + c16 * 16) = make_uint4(one, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t bytes = static_cast<uint32_t>(u.nload) * kBlkBytes;
      for (int kb = 0; kb < num_kb; ++kb) {
        if (!active(kb)) continue;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        uint8_t* st = smem + s * p.stage_bytes;
        for (int i = 0; i < u.nload; ++i)  // a block at or beyond C (C = 64 only) is an out-of-bounds box == zeros
          tma_load_3d(st + i * kBlkBytes, &tmF, &full_bar[s], u.load_blk[i] * 64, p_begin + kb * KP, b);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // issue-thread rules of profiles/r01_umma_issue_probe.txt: descriptor built once and advanced by immediates,
      // running ring counters
      const uint32_t idesc = umma_idesc_bf16(128, N, true, true);
      const uint32_t idesc1 = umma_idesc_bf16(128, 16, true, true);
      const uint32_t ones_lo = static_cast<uint32_t>(umma_desc_sw128(smem_u32(ones), kBlkBytes, 1024));
      const uint64_t d0 = umma_desc_sw128(smem_u32(smem), kBlkBytes, 1024);
      const uint32_t d_hi = static_cast<uint32_t>(d0 >> 32);
      const uint32_t lo0 = static_cast<uint32_t>(d0);
      const uint32_t a_off = static_cast<uint32_t>(u.a_slot * kBlkBytes) >> 4;
      const uint32_t b_off = static_cast<uint32_t>(u.b_slot * kBlkBytes) >> 4;
      uint64_t* ready_bar = MASKED ? xf_bar : full_bar;
      int s = 0;
      uint32_t ph = 0;
      uint32_t acc = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        if (!active(kb)) continue;
        mbar_wait(&ready_bar[s], ph);
        tc_fence_after();
        const uint32_t cur = lo0 + static_cast<uint32_t>((s * p.stage_bytes) >> 4);
#pragma unroll
        for (int k = 0; k < KP / 16; ++k) {
          umma_bf16_lohi(tmem_base, cur + a_off + ((k * 2048) >> 4), d_hi, cur + b_off + ((k * 2048) >> 4), d_hi, idesc, acc);
          if (p.csum)
            umma_bf16_lohi(tmem_base + N, cur + a_off + ((k * 2048) >> 4), d_hi, ones_lo + ((k * 2048) >> 4), d_hi, idesc1, acc);
          acc = 1u;
        }
        umma_commit(&empty_bar[s]);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
      umma_commit(tmem_full_bar);  // arrives at once when nothing was issued
    }
  } else if (warp < 6) {
    const int q = warp & 3;
    const int row = q * 32 + lane;       // channel within this unit's 128-row block
    const int ch = u.a_blk * 64 + row;
    bool any = false;
    for (int kb = 0; kb < num_kb; ++kb) any = any || active(kb);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    float* dst = p.partial + ((static_cast<size_t>(b) * p.splits + split) * p.C + (ch < p.C ? ch : 0)) * p.C + u.b_blk * 64;
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      if (any) {
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      if (ch < p.C) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<uint4*>(dst + c0 + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
    if (p.csum) {
      uint32_t v[16];
      v[0] = 0u;
      if (any) {
        tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + N, v);
        tmem_ld_wait();
      }
      if (ch < p.C) p.csum[(static_cast<size_t>(b) * p.splits + split) * p.C + ch] = __uint_as_float(v[0]);
    }
    tc_fence_before();
  } else if (MASKED) {
    // transform warps: tile row (= pixel) *= m[pixel] in place; F * m^2 of the A blocks goes to global memory.
    // Thread t owns 16-byte chunk (t & 7) of rows (t >> 3) + 16 j: its KP/16 mask values are fetched BEFORE it waits for the
    // tile, so the global-load latency hides behind the TMA (a per-row load inside the loop made this stage latency-bound).
    const int t = threadIdx.x - 192;
    const int r0 = t >> 3, c16 = t & 7;
    const float* mrow = p.mask + static_cast<long>(p.mask_b > 1 ? b : 0) * p.HW;
    int s = 0;
    uint32_t ph = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      if (!active(kb)) continue;
      const int pix0 = p_begin + kb * KP;
      float mv[KP / 16];
#pragma unroll
      for (int j = 0; j < KP / 16; ++j) {
        const int pix = pix0 + r0 + 16 * j;
        mv[j] = pix < p.HW ? __ldg(mrow + pix) : 0.f;
      }
      mbar_wait(&full_bar[s], ph);
      uint8_t* st = smem + s * p.stage_bytes;
      for (int blk = 0; blk < u.nload; ++blk) {
        const int chan0 = u.load_blk[blk] * 64;
        if (chan0 >= p.C) continue;                     // zero-filled out-of-bounds block (C = 64): nothing to scale
        const bool wr = u.write_fm2 && blk >= u.a_slot && blk < u.a_slot + 2;
#pragma unroll
        for (int j = 0; j < KP / 16; ++j) {
          const int r = r0 + 16 * j;
          uint4* ptr = reinterpret_cast<uint4*>(st + blk * kBlkBytes + r * 128 + c16 * 16);
          const uint4 v = *ptr;
          const float m1 = mv[j];
          float2 f0 = unpack_bf16x2(v.x), f1 = unpack_bf16x2(v.y), f2 = unpack_bf16x2(v.z), f3 = unpack_bf16x2(v.w);
          f0.x *= m1; f0.y *= m1; f1.x *= m1; f1.y *= m1; f2.x *= m1; f2.y *= m1; f3.x *= m1; f3.y *= m1;
          *ptr = make_uint4(pack_bf16x2(f0.x, f0.y), pack_bf16x2(f1.x, f1.y), pack_bf16x2(f2.x, f2.y), pack_bf16x2(f3.x, f3.y));
          const int pix = pix0 + r;
          if (wr && pix < p.HW) {
            const int chan = chan0 + ((c16 ^ (r & 7)) << 3);  // 128-byte swizzle: chunk ^ (row % 8)
            *reinterpret_cast<uint4*>(p.fm2 + (static_cast<long>(b) * p.HW + pix) * p.C + chan) =
                make_uint4(pack_bf16x2(f0.x * m1, f0.y * m1), pack_bf16x2(f1.x * m1, f1.y * m1),
                           pack_bf16x2(f2.x * m1, f2.y * m1), pack_bf16x2(f3.x * m1, f3.y * m1));
          }
        }
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&xf_bar[s]);
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (ncols == 256) tmem_dealloc<256>(tmem_base);
    else if (ncols == 128) tmem_dealloc<128>(tmem_base);
    else tmem_dealloc<64>(tmem_base);
  }
}

int gram_kp(int C) { return C <= 128 ? 64 : 32; }

static GramUnit make_unit(int a_blk, int b_blk, int nb, int write_fm2) {
  GramUnit u;
  memset(&u, 0, sizeof(u));
  u.a_blk = a_blk; u.b_blk = b_blk; u.nb = nb; u.write_fm2 = write_fm2;
  if (a_blk >= b_blk && a_blk + 2 <= b_blk + nb) {          // A inside B
    for (int i = 0; i < nb; ++i) u.load_blk[u.nload++] = b_blk + i;
    u.a_slot = a_blk - b_blk; u.b_slot = 0;
  } else if (b_blk >= a_blk && b_blk + nb <= a_blk + 2) {   // B inside A (C = 64)
    u.load_blk[u.nload++] = a_blk; u.load_blk[u.nload++] = a_blk + 1;
    u.a_slot = 0; u.b_slot = b_blk - a_blk;
  } else {                                                  // disjoint: A's two blocks, then B's
    u.load_blk[u.nload++] = a_blk; u.load_blk[u.nload++] = a_blk + 1;
    for (int i = 0; i < nb; ++i) u.load_blk[u.nload++] = b_blk + i;
    u.a_slot = 0; u.b_slot = 2;
  }
  return u;
}

static int build_units(int C, GramUnit* u) {
  switch (C) {
    case 64: u[0] = make_unit(0, 0, 1, 1); return 1;
    case 128: u[0] = make_unit(0, 0, 2, 1); return 1;
    case 256: u[0] = make_unit(0, 0, 4, 1); u[1] = make_unit(2, 2, 2, 1); return 2;
    case 512:  // heaviest (6 blocks per stage) first
      u[0] = make_unit(0, 4, 4, 0); u[1] = make_unit(0, 0, 4, 1); u[2] = make_unit(2, 2, 4, 1);
      u[3] = make_unit(4, 4, 4, 1); u[4] = make_unit(2, 6, 2, 0); u[5] = make_unit(6, 6, 2, 1);
      return 6;
    default: return 0;
  }
}

int gram_pick_splits(int B, int HW, int C) {
  GramUnit tmp[6];
  const int n_units = std::max(1, build_units(C, tmp));
  const int kp = gram_kp(C);
  // two (C >= 256) to three CTAs are resident per SM; aim at just UNDER two full waves of them -- rounding the split count
  // up gave 2.02-2.6 waves, i.e. a third, almost empty wave (ncu: 61-65 % of the DRAM peak on the HBM-bound layers)
  const int per_sm = C >= 256 ? 4 : 6;
  int want = (per_sm * kNumSMs) / (B * n_units);
  int max_splits = HW / (kp * 4);  // keep >= 4 K blocks per split
  if (max_splits < 1) max_splits = 1;
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  // every split must own at least one pixel
  int chunk = ((HW + want - 1) / want + kp - 1) / kp * kp;
  return (HW + chunk - 1) / chunk;
}

// flags[mb][kb] = 1 iff any mask value of pixels [kb*KP, (kb+1)*KP) is non-zero
__global__ void gram_mask_flags_kernel(const float* __restrict__ m, int HW, int KP, int n_kb, uint8_t* __restrict__ flags) {
  const int mb = blockIdx.y;
  const int kb = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (kb >= n_kb) return;
  const int lane = threadIdx.x & 31;
  bool nz = false;
  for (int i = lane; i < KP; i += 32) {
    const int pix = kb * KP + i;
    if (pix < HW) nz = nz || (m[static_cast<long>(mb) * HW + pix] != 0.f);
  }
  nz = __any_sync(0xffffffffu, nz);
  if (lane == 0) flags[static_cast<long>(mb) * n_kb + kb] = nz ? 1 : 0;
}

int gram_mask_flags_bytes(int mask_b, int HW, int C) {
  const int kp = gram_kp(C);
  return mask_b * ((HW + kp - 1) / kp);
}

int gram_mask_flags(const float* m, int mask_b, int HW, int C, uint8_t* flags, cudaStream_t stream) {
  const int kp = gram_kp(C);
  const int n_kb = (HW + kp - 1) / kp;
  dim3 grid((n_kb + 7) / 8, mask_b);
  gram_mask_flags_kernel<<<grid, 256, 0, stream>>>(m, HW, kp, n_kb, flags);
  ISX_LAUNCH_CHECK();
  return 0;
}

template <int KP, bool MASKED>
static int launch_gram(const __nv_bfloat16* feat, int B, int HW, int C, int splits, float* partial, const GramMask* mask,
                       float* csum, cudaStream_t stream) {
  GramParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.HW = HW; p.C = C; p.splits = splits;
  p.chunk = ((HW + splits - 1) / splits + KP - 1) / KP * KP;
  p.n_units = build_units(C, p.units);
  p.partial = partial;
  p.csum = csum;
  int max_load = 0;
  for (int i = 0; i < p.n_units; ++i) max_load = std::max(max_load, p.units[i].nload);
  p.stage_bytes = max_load * KP * 128;
  // <= ~100 KB of pipeline per CTA so that two CTAs share an SM (C <= 128: three)
  int stages = (C >= 256 ? 98 * 1024 : 64 * 1024) / p.stage_bytes;
  stages = std::max(3, std::min(stages, 8));
  p.stages = stages;
  if (MASKED) {
    p.mask = mask->m; p.mask_b = mask->mask_b; p.kb_flags = mask->kb_flags; p.fm2 = mask->fm2;
    p.n_kb_total = (HW + KP - 1) / KP;
  }
  const size_t smem_bytes = 1024 + static_cast<size_t>(stages) * p.stage_bytes + (csum ? KP * 128 : 0) + 256;
  CUtensorMap tmF;
  uint64_t dims[3] = {(uint64_t)C, (uint64_t)HW, (uint64_t)B};
  uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)HW * C * 2};
  uint32_t box[3] = {64, (uint32_t)KP, 1};
  if (isx_make_tmap_bf16(&tmF, feat, 3, dims, str, box, true)) return 3;
  auto kern = gram_sym_kernel<KP, MASKED>;
  ISX_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  const long grid = static_cast<long>(B) * p.n_units * splits;
  // algorithmic FLOPs of the layer's Gram: 2*C*C*HW per image (SURVEY.md §8d counts the full matrix), whatever part
  // of it the symmetric schedule actually multiplies
  isx_prof_begin(ISX_PROF_GRAM, 2.0 * C * C * static_cast<double>(B) * HW, stream);
  kern<<<(unsigned)grid, MASKED ? 320 : 192, smem_bytes, stream>>>(tmF, p);
  isx_prof_end(ISX_PROF_GRAM, stream);
  ISX_LAUNCH_CHECK();
  return 0;
}

int gram_sym_partial(const __nv_bfloat16* feat, int B, int HW, int C, int splits, float* partial, const GramMask* mask,
                     cudaStream_t stream, float* csum) {
  ISX_REQUIRE(B > 0 && HW > 0, "gram: empty feature map");
  ISX_REQUIRE(splits >= 1, "gram: splits must be >= 1");
  ISX_REQUIRE(C == 64 || C == 128 || C == 256 || C == 512, "gram: C=%d unsupported (64/128/256/512)", C);
  ISX_REQUIRE(csum == nullptr || C <= 128, "gram: fused channel sums exist for C <= 128 only");
  if (mask) {
    ISX_REQUIRE(mask->m && mask->kb_flags && mask->fm2 && (mask->mask_b == 1 || mask->mask_b == B), "gram: bad mask arguments");
    return C <= 128 ? launch_gram<64, true>(feat, B, HW, C, splits, partial, mask, csum, stream)
                    : launch_gram<32, true>(feat, B, HW, C, splits, partial, mask, csum, stream);
  }
  return C <= 128 ? launch_gram<64, false>(feat, B, HW, C, splits, partial, nullptr, csum, stream)
                  : launch_gram<32, false>(feat, B, HW, C, splits, partial, nullptr, csum, stream);
}

int gram_tc_partial(const __nv_bfloat16* feat, int B, int HW, int C, int splits, float* partial, cudaStream_t stream) {
  return gram_sym_partial(feat, B, HW, C, splits, partial, nullptr, stream);
}

}  // namespace isx
