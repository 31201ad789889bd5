// Evaluation metrics the reference's drivers compute on the device around the NST path:
//   cal_IoUs (utils.py:163-194; call sites iris_style_transfer_openeds2019.py:156, data_preprocessing.py:168): per image and
//   class the intersection-over-union of two label maps.  The reference makes, per class, two float copies of both maps, a
//   product, a sum, a clamp and two reductions -- ~30 elementwise passes over B x H x W for four classes.  Here ONE pass
//   reads the two int64 maps (16 B per pixel: HBM-bound), warp ballots turn 32 pixels into one match word per class and map,
//   and popc(p & t) / popc(p | t) count intersection and union exactly; the ratios are formed in fp32 like the reference
//   (the float sums of 0/1 values are exact integers below 2^24, so the results are bit-identical).
//   angular_distance (utils.py:216-240; gaze_estimation.py:85,102,119): acos of the clamped row dot product, and degrees.
#include <algorithm>

#include "../../include/isx.h"
#include "isx_common.cuh"

namespace isx {
namespace {

inline cudaStream_t S(isx_stream s) { return static_cast<cudaStream_t>(s); }
constexpr int kIouThreads = 256;
constexpr int kIouMaxClass = 8;

template <int NC>
__global__ void __launch_bounds__(kIouThreads)
seg_iou_count_kernel(const long long* __restrict__ preds, const long long* __restrict__ targets, long long hw,
                     unsigned int* __restrict__ counts /* [B][NC][2] */) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const long long warps = static_cast<long long>(gridDim.x) * (kIouThreads / 32);
  const long long* p = preds + static_cast<long long>(b) * hw;
  const long long* t = targets + static_cast<long long>(b) * hw;
  unsigned inter[NC], uni[NC];   // warp-uniform: every lane holds the same counts
#pragma unroll
  for (int c = 0; c < NC; ++c) { inter[c] = 0; uni[c] = 0; }
  constexpr int U = 4;           // 32-pixel groups in flight per warp: 2 x 4 x 256 B
  const long long groups = (hw + 31) / 32;
  for (long long g0 = (static_cast<long long>(blockIdx.x) * (kIouThreads / 32) + (threadIdx.x >> 5)) * U; g0 < groups; g0 += warps * U) {
    long long pv[U], tv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = (g0 + u) * 32 + lane;
      const bool in = i < hw;
      pv[u] = in ? p[i] : -1;    // -1 matches no class 0..NC-1
      tv[u] = in ? t[i] : -1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const unsigned pm = __ballot_sync(0xffffffffu, pv[u] == c), tm = __ballot_sync(0xffffffffu, tv[u] == c);
        inter[c] += __popc(pm & tm);
        uni[c] += __popc(pm | tm);
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (inter[c]) atomicAdd(&counts[(b * NC + c) * 2 + 0], inter[c]);
      if (uni[c]) atomicAdd(&counts[(b * NC + c) * 2 + 1], uni[c]);
    }
  }
}

__global__ void seg_iou_finalize_kernel(const unsigned int* __restrict__ counts, int B, int NC, float eps, float* __restrict__ iou,
                                        float* __restrict__ miou) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float sum = 0.f;
  for (int c = 0; c < NC; ++c) {
    const float i = static_cast<float>(counts[(b * NC + c) * 2 + 0]), u = static_cast<float>(counts[(b * NC + c) * 2 + 1]);
    const float v = __fdiv_rn(i, __fadd_rn(u, eps));   // intersection / (union + eps), utils.py:189
    iou[b * NC + c] = v;
    sum = __fadd_rn(sum, v);
  }
  miou[b] = __fdiv_rn(sum, static_cast<float>(NC));      // ious.mean(dim = 1)
}

__global__ void angular_distance_kernel(const float* __restrict__ v1, const float* __restrict__ v2, int n, int d,
                                        float* __restrict__ radian, float* __restrict__ degree) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float dot = 0.f;
  for (int k = 0; k < d; ++k) dot = __fadd_rn(dot, __fmul_rn(v1[i * d + k], v2[i * d + k]));   // torch.sum(v1 * v2, dim = 1)
  dot = fminf(fmaxf(dot, -1.0f), 1.0f);
  const float r = acosf(dot);
  radian[i] = r;
  degree[i] = r * 57.29577951308232f;   // torch.rad2deg: x * (180 / pi)
}

}  // namespace
}  // namespace isx

using namespace isx;

extern "C" int isx_seg_iou(const int64_t* preds, const int64_t* targets, int B, int64_t HW, int num_class, float eps,
                           uint32_t* counts, float* iou, float* miou, isx_stream stream) {
  ISX_REQUIRE(preds && targets && counts && iou && miou, "isx_seg_iou: null pointer");
  ISX_REQUIRE(B > 0 && B <= 65535 && HW > 0 && HW < (1ll << 32), "isx_seg_iou: bad shape B=%d (1..65535) HW=%lld", B, static_cast<long long>(HW));
  ISX_REQUIRE(num_class >= 1 && num_class <= kIouMaxClass, "isx_seg_iou: num_class %d (1..%d)", num_class, kIouMaxClass);
  cudaStream_t s = S(stream);
  ISX_CHECK_CUDA(cudaMemsetAsync(counts, 0, static_cast<size_t>(B) * num_class * 2 * sizeof(uint32_t), s));
  const long long groups = (HW + 31) / 32;
  // two resident waves over the batch (ncu: 768 blocks at 4 resident per SM = 1.3 waves ran at 4.4 TB/s)
  int per_sm = 4;
  ISX_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, seg_iou_count_kernel<4>, kIouThreads, 0));
  const int cap = std::max(1, isx_num_sms() * std::max(1, per_sm) * 2 / B);
  const int bx = static_cast<int>(std::max<long long>(1, std::min<long long>(cap, (groups + 8 * 4 - 1) / (8 * 4))));
  const dim3 grid(bx, B);
  const long long* p = reinterpret_cast<const long long*>(preds);
  const long long* t = reinterpret_cast<const long long*>(targets);
  switch (num_class) {
    case 1: seg_iou_count_kernel<1><<<grid, kIouThreads, 0, s>>>(p, t, HW, counts); break;
    case 2: seg_iou_count_kernel<2><<<grid, kIouThreads, 0, s>>>(p, t, HW, counts); break;
    case 3: seg_iou_count_kernel<3><<<grid, kIouThreads, 0, s>>>(p, t, HW, counts); break;
    case 4: seg_iou_count_kernel<4><<<grid, kIouThreads, 0, s>>>(p, t, HW, counts); break;
    case 5: seg_iou_count_kernel<5><<<grid, kIouThreads, 0, s>>>(p, t, HW, counts); break;
    case 6: seg_iou_count_kernel<6><<<grid, kIouThreads, 0, s>>>(p, t, HW, counts); break;
    case 7: seg_iou_count_kernel<7><<<grid, kIouThreads, 0, s>>>(p, t, HW, counts); break;
    default: seg_iou_count_kernel<8><<<grid, kIouThreads, 0, s>>>(p, t, HW, counts); break;
  }
  ISX_LAUNCH_CHECK();
  seg_iou_finalize_kernel<<<(B + 127) / 128, 128, 0, s>>>(counts, B, num_class, eps, iou, miou);
  ISX_LAUNCH_CHECK();
  return 0;
}

extern "C" int isx_angular_distance(const float* v1, const float* v2, int n, int d, float* radian, float* degree, isx_stream stream) {
  ISX_REQUIRE(v1 && v2 && radian && degree && n > 0 && d > 0, "isx_angular_distance: bad arguments");
  angular_distance_kernel<<<(n + 255) / 256, 256, 0, S(stream)>>>(v1, v2, n, d, radian, degree);
  ISX_LAUNCH_CHECK();
  return 0;
}
