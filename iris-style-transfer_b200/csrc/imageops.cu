// K10 / K9: iris mask + bounding box (pipelines.py:139-165, utils.py:44-72) and the composite back
// into the eye frame (iris_style_transfer_openeds2019.py:111-130, …2020.py:121-139), including
// torchvision's antialiased bilinear Resize (ATen _upsample_bilinear2d_aa).  Integer / index work is
// bit-exact; the resize is fp32 with the reference's summation order (W pass inside, H pass outside).
#include <algorithm>
#include <limits.h>

#include "../../include/isx.h"
#include "isx_common.cuh"

namespace isx {

// bbox[b] = {row_min, col_min, row_max, col_max}; caller-visible sentinel for "no nonzero pixel": row_max = -1
__global__ void bbox_init_kernel(int32_t* bbox, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    bbox[b * 4 + 0] = INT_MAX; bbox[b * 4 + 1] = INT_MAX; bbox[b * 4 + 2] = -1; bbox[b * 4 + 3] = -1;
  }
}

// m = (seg == label) & (x <= thr)   [either test may be disabled]; xm = x * m; bbox over xm != 0
__global__ void __launch_bounds__(256)
mask_bbox_kernel(const float* __restrict__ x, const int64_t* __restrict__ seg, int label, int use_thr, float thr,
                 uint8_t* __restrict__ mask, float* __restrict__ xm, int32_t* __restrict__ bbox, int H, int W) {
  const int b = blockIdx.y;
  const long hw = static_cast<long>(H) * W;
  int rmin = INT_MAX, cmin = INT_MAX, rmax = -1, cmax = -1;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < hw;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float v = x[b * hw + i];
    bool m = true;
    if (seg) m = seg[b * hw + i] == label;
    if (use_thr) m = m && (v <= thr);
    const float vm = m ? v : v * 0.0f;  // x * m with m in {0,1}
    if (mask) mask[b * hw + i] = m ? 1 : 0;
    if (xm) xm[b * hw + i] = vm;
    if (vm != 0.0f) {  // image.nonzero() (utils.py:57): nonzero PIXELS, not nonzero mask
      const int r = static_cast<int>(i / W), c = static_cast<int>(i % W);
      rmin = min(rmin, r); rmax = max(rmax, r); cmin = min(cmin, c); cmax = max(cmax, c);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    rmin = min(rmin, __shfl_xor_sync(0xffffffffu, rmin, o));
    cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
    rmax = max(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
  }
  if ((threadIdx.x & 31) == 0 && rmax >= 0) {
    atomicMin(&bbox[b * 4 + 0], rmin); atomicMin(&bbox[b * 4 + 1], cmin);
    atomicMax(&bbox[b * 4 + 2], rmax); atomicMax(&bbox[b * 4 + 3], cmax);
  }
}

// ---- antialiased bilinear weights (ATen UpSampleKernel.cpp, HelperInterpLinear, align_corners=False) ----
struct AA {
  int lo, n;
  float scale, support, invscale, center;
};
__device__ __forceinline__ AA aa_setup(int in_size, int out_size, int o) {
  AA a;
  a.scale = static_cast<float>(in_size) / static_cast<float>(out_size);
  a.support = a.scale >= 1.0f ? a.scale : 1.0f;
  a.invscale = a.scale >= 1.0f ? 1.0f / a.scale : 1.0f;
  a.center = a.scale * (o + 0.5f);
  a.lo = max(static_cast<int>(a.center - a.support + 0.5f), 0);
  a.n = min(static_cast<int>(a.center + a.support + 0.5f), in_size) - a.lo;
  return a;
}
__device__ __forceinline__ float aa_w(const AA& a, int j) {
  const float t = fabsf((j + a.lo - a.center + 0.5f) * a.invscale);
  return t < 1.0f ? 1.0f - t : 0.0f;
}

// out[b, 0, r, c] over the destination window, gray = .2989 R + .587 G + .114 B of src when src_c == 3.
// mode 0: plain resize into dst [B,dst_c,oh,ow] (channels replicated dst_c times) from the src window
//         (bbox rows/cols inclusive, or the whole image when src_bbox == NULL)
// mode 1: composite: dst frame [B,1,H,W] in place, window = dst_bbox; frame = mask ? value : frame
__global__ void __launch_bounds__(256)
resize_aa_kernel(const float* __restrict__ src, int src_c, int SH, int SW, const int32_t* __restrict__ src_bbox,
                 float* __restrict__ dst, int dst_c, int DH, int DW, const int32_t* __restrict__ dst_bbox,
                 const uint8_t* __restrict__ mask, int mode) {
  const int b = blockIdx.y;
  int sy0 = 0, sx0 = 0, sh = SH, sw = SW;
  if (src_bbox) {
    sy0 = src_bbox[b * 4 + 0]; sx0 = src_bbox[b * 4 + 1];
    sh = src_bbox[b * 4 + 2] - sy0 + 1; sw = src_bbox[b * 4 + 3] - sx0 + 1;
  }
  int dy0 = 0, dx0 = 0, oh = DH, ow = DW;
  if (dst_bbox) {
    dy0 = dst_bbox[b * 4 + 0]; dx0 = dst_bbox[b * 4 + 1];
    oh = dst_bbox[b * 4 + 2] - dy0 + 1; ow = dst_bbox[b * 4 + 3] - dx0 + 1;
  }
  if (sh <= 0 || sw <= 0 || oh <= 0 || ow <= 0) return;
  const long n = static_cast<long>(oh) * ow;
  const long shw = static_cast<long>(SH) * SW;
  const float* sb = src + static_cast<long>(b) * src_c * shw;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / ow), c = static_cast<int>(i % ow);
    const AA ay = aa_setup(sh, oh, r), ax = aa_setup(sw, ow, c);
    float wxs = 0.f, wys = 0.f;
    for (int j = 0; j < ax.n; ++j) wxs += aa_w(ax, j);
    for (int j = 0; j < ay.n; ++j) wys += aa_w(ay, j);
    float acc = 0.f;
    for (int jy = 0; jy < ay.n; ++jy) {
      const long rowoff = static_cast<long>(sy0 + ay.lo + jy) * SW + sx0 + ax.lo;
      float t = 0.f;
      for (int jx = 0; jx < ax.n; ++jx) {
        float v;
        if (src_c == 3) {
          const float R = sb[rowoff + jx], G = sb[shw + rowoff + jx], Bc = sb[2 * shw + rowoff + jx];
          v = __fadd_rn(__fadd_rn(__fmul_rn(R, 0.2989f), __fmul_rn(G, 0.587f)), __fmul_rn(Bc, 0.114f));
        } else {
          v = sb[rowoff + jx];
        }
        const float w = wxs != 0.f ? aa_w(ax, jx) / wxs : aa_w(ax, jx);
        t = jx == 0 ? __fmul_rn(v, w) : __fadd_rn(t, __fmul_rn(v, w));
      }
      const float w = wys != 0.f ? aa_w(ay, jy) / wys : aa_w(ay, jy);
      acc = jy == 0 ? __fmul_rn(t, w) : __fadd_rn(acc, __fmul_rn(t, w));
    }
    if (mode == 0) {
      for (int ch = 0; ch < dst_c; ++ch)
        dst[((static_cast<long>(b) * dst_c + ch) * DH + dy0 + r) * DW + dx0 + c] = acc;
    } else {
      const long o = (static_cast<long>(b) * DH + dy0 + r) * DW + dx0 + c;
      if (mask[o]) dst[o] = acc;  // frame * ~m + new * m  (…2019.py:125-130)
    }
  }
}

}  // namespace isx

using namespace isx;
static inline cudaStream_t S(isx_stream s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" int isx_mask_bbox(const float* x, const int64_t* seg, int label, int use_threshold, float threshold,
                             uint8_t* mask, float* xm, int32_t* bbox, int B, int H, int W, isx_stream stream) {
  ISX_REQUIRE(x && bbox && B > 0 && H > 0 && W > 0, "isx_mask_bbox: bad arguments");
  bbox_init_kernel<<<(B + 127) / 128, 128, 0, S(stream)>>>(bbox, B);
  ISX_LAUNCH_CHECK();
  const long hw = static_cast<long>(H) * W;
  const int bx = static_cast<int>(std::min<long>((hw + 255) / 256, std::max<long>(1, 148L * 8 / B)));
  dim3 grid(bx, B);
  mask_bbox_kernel<<<grid, 256, 0, S(stream)>>>(x, seg, label, use_threshold, threshold, mask, xm, bbox, H, W);
  ISX_LAUNCH_CHECK();
  return 0;
}

extern "C" int isx_resize_bilinear_aa(const float* src, int src_c, int SH, int SW, const int32_t* src_bbox, float* dst,
                                      int dst_c, int DH, int DW, int B, isx_stream stream) {
  ISX_REQUIRE(src && dst && (src_c == 1 || src_c == 3) && dst_c >= 1 && B > 0, "isx_resize_bilinear_aa: bad arguments");
  const long n = static_cast<long>(DH) * DW;
  const int bx = static_cast<int>(std::min<long>((n + 255) / 256, std::max<long>(1, 148L * 8 / B)));
  dim3 grid(bx, B);
  resize_aa_kernel<<<grid, 256, 0, S(stream)>>>(src, src_c, SH, SW, src_bbox, dst, dst_c, DH, DW, nullptr, nullptr, 0);
  ISX_LAUNCH_CHECK();
  return 0;
}

extern "C" int isx_composite(const float* new_iris, int src_c, int SH, int SW, float* frames, const uint8_t* mask,
                             const int32_t* bbox, int B, int H, int W, isx_stream stream) {
  ISX_REQUIRE(new_iris && frames && mask && bbox && (src_c == 1 || src_c == 3) && B > 0, "isx_composite: bad arguments");
  const long n = static_cast<long>(H) * W;
  const int bx = static_cast<int>(std::min<long>((n + 255) / 256, std::max<long>(1, 148L * 8 / B)));
  dim3 grid(bx, B);
  resize_aa_kernel<<<grid, 256, 0, S(stream)>>>(new_iris, src_c, SH, SW, nullptr, frames, 1, H, W, bbox, mask, 1);
  ISX_LAUNCH_CHECK();
  return 0;
}
