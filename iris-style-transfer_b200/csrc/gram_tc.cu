// K4: Gram matrix F^T F of an NHWC bf16 feature map, per image, on tcgen05 tensor cores.
// Reference: utils.py:242-257 (GramMatrix: flatten(H,W); x @ x^T / n), called from
// StyleLoss_Gram (utils.py:305,319).
//
// In NHWC the feature map of one image is a [HW x C] matrix with the channel contiguous, i.e. both
// GEMM operands of G = F^T F are "MN-major" (M/N index contiguous, the reduction index = pixel
// strided).  tcgen05 consumes that layout directly (a_major = b_major = MN in the instruction
// descriptor), so the tile TMA brings in -- [KP pixels] x [64 channels] boxes, 128-byte rows,
// SWIZZLE_128B -- is used as BOTH operands: a CTA owns 128 output rows (channels) x all C columns
// and never loads a separate B tile; the A descriptor simply points at two of the 64-channel
// blocks of the same stage.  Split-K over pixels; fp32 partials are reduced (deterministically)
// by gram_finalize in elementwise.cu, fused with 1/n, (G - T), the loss and the bf16 dL/dG matrix.
#include "isx_common.cuh"
#include "isx_internal.h"

namespace isx {

struct GramParams {
  int B, HW, C;
  int splits;
  int chunk;   // pixels per split (multiple of KP)
  int mblks;   // row blocks of 128 channels
  int stages;
  float* partial;  // [B][splits][C][C]
};

// NB = 64-channel blocks resident per stage (>= 2 so that the 128-row A operand always exists;
// for C = 64 the second block is an out-of-bounds TMA box == zeros).  KP = pixels per K block.
template <int NB, int KP>
__global__ void __launch_bounds__(192, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tmF, const GramParams p) {
  constexpr int kBlkBytes = KP * 128;
  constexpr int kStageBytes = NB * kBlkBytes;
  constexpr int kCols = NB * 64 >= 512 ? 512 : (NB * 64 <= 128 ? 128 : 256);  // accumulator columns (pow2)
  constexpr int kNHalves = NB == 8 ? 2 : 1;
  constexpr int kN = NB == 8 ? 256 : NB * 64;  // N per MMA (C = 64 handled at run time below)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = p.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + stages * kStageBytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int split = blockIdx.x % p.splits;
  const int mblk = (blockIdx.x / p.splits) % p.mblks;
  const int b = blockIdx.x / (p.splits * p.mblks);

  const int p_begin = split * p.chunk;
  const int p_end = min(p.HW, p_begin + p.chunk);
  const int num_kb = p_end > p_begin ? (p_end - p_begin + KP - 1) / KP : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmF);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
        uint8_t* st = smem + s * kStageBytes;
#pragma unroll
        for (int cb = 0; cb < NB; ++cb)  // cb*64 >= C only happens for C == 64: zero-filled box
          tma_load_3d(st + cb * kBlkBytes, &tmF, &full_bar[s], cb * 64, p_begin + kb * KP, b);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // Same issue-thread rules as the conv kernels (profiles/r01_umma_issue_probe.txt): descriptor built once and advanced
      // by immediates, running ring counters, the next stage's mbarrier probed before this stage's MMAs.
      const uint32_t idesc = umma_idesc_bf16(128, p.C == 64 ? 64 : kN, true, true);
      const uint64_t d0 = umma_desc_sw128(smem_u32(smem), kBlkBytes, 1024);
      const uint32_t d_hi = static_cast<uint32_t>(d0 >> 32);
      const uint32_t lo0 = static_cast<uint32_t>(d0);
      const uint32_t a_off = static_cast<uint32_t>((2 * mblk) * kBlkBytes) >> 4;
      uint32_t lo = lo0;
      int s = 0;
      uint32_t ph = 0;
      bool ready = false;
      for (int kb = 0; kb < num_kb; ++kb) {
        if (!ready) mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t cur = lo;
        const int s_cur = s;
        lo += kStageBytes >> 4;
        if (++s == stages) { s = 0; ph ^= 1; lo = lo0; }
        ready = (kb + 1 < num_kb) && mbar_try_wait(&full_bar[s], ph);
#pragma unroll
        for (int k = 0; k < KP / 16; ++k) {
#pragma unroll
          for (int nh = 0; nh < kNHalves; ++nh)
            umma_bf16_lohi(tmem_base + nh * 256, cur + a_off + ((k * 2048) >> 4), d_hi,
                           cur + (((nh * 4) * kBlkBytes + k * 2048) >> 4), d_hi, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s_cur]);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;       // channel within this CTA's 128-row block
    const int ch = mblk * 128 + row;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    float* dst = p.partial + ((static_cast<size_t>(b) * p.splits + split) * p.C + (ch < p.C ? ch : 0)) * p.C;
    for (int c0 = 0; c0 < p.C; c0 += 32) {
      uint32_t v[32];
      if (num_kb > 0) {
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
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kCols>(tmem_base);
  }
}

static int gram_kp(int C) { return C <= 128 ? 64 : 32; }

int gram_pick_splits(int B, int HW, int C) {
  const int mblks = C <= 128 ? 1 : C / 128;
  const int kp = gram_kp(C);
  // latency-bound kernel (small MMAs per K block): several small-footprint CTAs per SM, so ~6 CTAs per SM in total
  const int per_sm = C >= 512 ? 2 : 6;
  int want = (per_sm * kNumSMs + B * mblks - 1) / (B * mblks);
  int max_splits = HW / (kp * 4);  // keep >= 4 K blocks per split
  if (max_splits < 1) max_splits = 1;
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  // every split must own at least one pixel
  int chunk = ((HW + want - 1) / want + kp - 1) / kp * kp;
  return (HW + chunk - 1) / chunk;
}

template <int NB, int KP>
static int launch_gram(const __nv_bfloat16* feat, int B, int HW, int C, int splits, float* partial,
                       cudaStream_t stream) {
  GramParams p;
  p.B = B; p.HW = HW; p.C = C; p.splits = splits;
  p.chunk = ((HW + splits - 1) / splits + KP - 1) / KP * KP;
  p.mblks = C <= 128 ? 1 : C / 128;
  p.partial = partial;
  constexpr int kStageBytes = NB * KP * 128;
  // C <= 256: <= 64 KB of pipeline per CTA so that 2-3 CTAs share an SM (TMEM: 128/256 columns each);
  // C = 512 owns all 512 TMEM columns, one CTA per SM, deeper pipeline instead
  int stages = NB == 8 ? 5 : (64 * 1024) / kStageBytes;
  if (stages > 8) stages = 8;
  if (stages < 3) stages = 3;
  p.stages = stages;
  const size_t smem_bytes = 1024 + static_cast<size_t>(stages) * kStageBytes + 256;
  CUtensorMap tmF;
  uint64_t dims[3] = {(uint64_t)C, (uint64_t)HW, (uint64_t)B};
  uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)HW * C * 2};
  uint32_t box[3] = {64, (uint32_t)KP, 1};
  if (isx_make_tmap_bf16(&tmF, feat, 3, dims, str, box, true)) return 3;
  auto kern = gram_tc_kernel<NB, KP>;
  ISX_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  const long grid = static_cast<long>(B) * p.mblks * splits;
  isx_prof_begin(ISX_PROF_GRAM, 2.0 * C * C * static_cast<double>(B) * HW, stream);
  kern<<<(unsigned)grid, 192, smem_bytes, stream>>>(tmF, p);
  isx_prof_end(ISX_PROF_GRAM, stream);
  ISX_LAUNCH_CHECK();
  return 0;
}

int gram_tc_partial(const __nv_bfloat16* feat, int B, int HW, int C, int splits, float* partial,
                    cudaStream_t stream) {
  ISX_REQUIRE(B > 0 && HW > 0, "gram: empty feature map");
  ISX_REQUIRE(splits >= 1, "gram: splits must be >= 1");
  switch (C) {
    case 64: return launch_gram<2, 64>(feat, B, HW, C, splits, partial, stream);
    case 128: return launch_gram<2, 64>(feat, B, HW, C, splits, partial, stream);
    case 256: return launch_gram<4, 32>(feat, B, HW, C, splits, partial, stream);
    case 512: return launch_gram<8, 32>(feat, B, HW, C, splits, partial, stream);
    default: ISX_REQUIRE(false, "gram: C=%d unsupported (64/128/256/512)", C);
  }
}

}  // namespace isx
