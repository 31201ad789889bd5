// Classifier heads of the iris-recognition path: models/classifiers/classifiers.py:3-72,
// call sites iris_classification.py:66-71,94-98 and iris_style_transfer_openeds2019.py:82-84,144-146.
//   Classifier1: AdaptiveAvgPool2d(7,7) -> Flatten -> Linear(25088,4096) ReLU [Dropout] Linear(4096,4096) ReLU [Dropout] Linear(4096,K)
//   Classifier2: cat(mean, std) of the style taps (1920) -> the same three Linear layers
// Inference (the drivers call them in eval mode: dropout is the identity).  A Linear layer over a batch of <= 128 eyes is a
// weight-streaming GEMM: out^T [N_out x M] = W [N_out x K] . X^T, so the 128-row MMA dimension is given to the OUTPUT
// FEATURES (the big, streamed operand: one TMA tile of 128 weight rows per K block, used once) and the batch becomes the
// N dimension -- exactly the library's 1x1 tcgen05 convolution with "pixels" = output features and "output channels" = batch
// rows.  The bias rides along as one extra K column (X' has a column of ones, W' holds the bias there), so the conv epilogue
// (ReLU, bf16 store) needs nothing new.  Small pack / transpose kernels connect the layers.
#include <algorithm>

#include "../../include/isx.h"
#include "isx_common.cuh"
#include "isx_internal.h"

namespace isx {

// dst bf16 [Npad][K + 64]: row n = (W[n][0..K), bias[n], 0 ...); rows >= N are zero
__global__ void linear_pack_kernel(const float* __restrict__ W, const float* __restrict__ bias, int N, int K, int Npad,
                                   __nv_bfloat16* __restrict__ dst) {
  const long ld = K + 64;
  const long total = static_cast<long>(Npad) * ld;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long n = i / ld, k = i - n * ld;
    float v = 0.f;
    if (n < N) v = k < K ? W[n * K + k] : (k == K ? bias[n] : 0.f);
    dst[i] = __float2bfloat16_rn(v);
  }
}

// X' bf16 [Mpad][K + 64] from fp32 rows [M][K] (row stride ld_src): cast, column K = 1, zero padding
__global__ void rows_pack_kernel(const float* __restrict__ src, long ld_src, int M, int K, int Mpad, __nv_bfloat16* __restrict__ dst) {
  const long ld = K + 64;
  const long total = static_cast<long>(Mpad) * ld;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long m = i / ld, k = i - m * ld;
    float v = 0.f;
    if (m < M) v = k < K ? src[m * ld_src + k] : (k == K ? 1.f : 0.f);
    dst[i] = __float2bfloat16_rn(v);
  }
}

// out^T bf16 [N][Mpad] -> X' bf16 [Mpad][N + 64] (next layer's input, ones column) through a 32x32 shared-memory transpose
__global__ void __launch_bounds__(256)
transpose_pack_kernel(const __nv_bfloat16* __restrict__ srcT, int N, int Mpad, int M, __nv_bfloat16* __restrict__ dst) {
  __shared__ __nv_bfloat16 t[32][33];
  const int n0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long ld = N + 64;
  for (int k = 0; k < 4; ++k) {
    const int n = n0 + ty + 8 * k, m = m0 + tx;
    t[ty + 8 * k][tx] = (n < N && m < Mpad) ? srcT[static_cast<long>(n) * Mpad + m] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int k = 0; k < 4; ++k) {
    const int m = m0 + ty + 8 * k, n = n0 + tx;
    if (m < Mpad && n < N) dst[m * ld + n] = m < M ? t[tx][ty + 8 * k] : __float2bfloat16_rn(0.f);
  }
  if (blockIdx.x == 0) {  // the 64 extra columns: a one, then zeros
    for (int i = threadIdx.x; i < 32 * 64; i += 256) {
      const int m = m0 + i / 64, c = i % 64;
      if (m < Mpad) dst[m * ld + N + c] = __float2bfloat16_rn((c == 0 && m < M) ? 1.f : 0.f);
    }
  }
}

// logits^T bf16 [Npad][Mpad] -> fp32 [M][N]
__global__ void transpose_out_kernel(const __nv_bfloat16* __restrict__ srcT, int N, int Mpad, int M, float* __restrict__ dst) {
  const long total = static_cast<long>(M) * N;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long m = i / N, n = i - m * N;
    dst[i] = __bfloat162float(srcT[n * Mpad + m]);
  }
}

// AdaptiveAvgPool2d((7,7)) + Flatten of pool5 (bf16 NHWC [B,h,w,C]) -> X' bf16 [Mpad][C*49 + 64] in NCHW flatten order
// (c*49 + i*7 + j), ones column; bins as torch: start = floor(i*h/7), end = ceil((i+1)*h/7)
__global__ void pool7_flatten_kernel(const __nv_bfloat16* __restrict__ p5, int B, int h, int w, int C, int Mpad,
                                     __nv_bfloat16* __restrict__ dst) {
  const long K = static_cast<long>(C) * 49, ld = K + 64;
  const long total = static_cast<long>(Mpad) * ld;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const long m = idx / ld, k = idx - m * ld;
    float v = 0.f;
    if (m < B) {
      if (k < K) {
        const int c = static_cast<int>(k / 49), i = static_cast<int>((k % 49) / 7), j = static_cast<int>(k % 7);
        const int y0 = (i * h) / 7, y1 = ((i + 1) * h + 6) / 7, x0 = (j * w) / 7, x1 = ((j + 1) * w + 6) / 7;
        float s = 0.f;
        for (int y = y0; y < y1; ++y)
          for (int x = x0; x < x1; ++x) s += __bfloat162float(p5[((m * h + y) * w + x) * C + c]);
        v = s / static_cast<float>((y1 - y0) * (x1 - x0));
      } else if (k == K) {
        v = 1.f;
      }
    }
    dst[idx] = __float2bfloat16_rn(v);
  }
}

static int grid_for(long total) { return static_cast<int>(std::min<long>((total + 255) / 256, static_cast<long>(isx_num_sms()) * 16)); }

}  // namespace isx

using namespace isx;
typedef __nv_bfloat16 bf16;
static inline cudaStream_t S(isx_stream s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" int isx_linear_pack(const float* W, const float* bias, int N, int K, int Npad, isx_bf16* dst, isx_stream stream) {
  ISX_REQUIRE(W && bias && dst && N > 0 && K > 0 && K % 64 == 0 && Npad >= N, "isx_linear_pack: bad arguments (K must be a multiple of 64)");
  const long total = static_cast<long>(Npad) * (K + 64);
  linear_pack_kernel<<<grid_for(total), 256, 0, S(stream)>>>(W, bias, N, K, Npad, reinterpret_cast<bf16*>(dst));
  ISX_LAUNCH_CHECK();
  return 0;
}

extern "C" int isx_rows_pack(const float* src, int64_t ld_src, int M, int K, int Mpad, isx_bf16* dst, isx_stream stream) {
  ISX_REQUIRE(src && dst && M > 0 && K % 64 == 0 && Mpad % 64 == 0 && Mpad >= M, "isx_rows_pack: bad arguments");
  const long total = static_cast<long>(Mpad) * (K + 64);
  rows_pack_kernel<<<grid_for(total), 256, 0, S(stream)>>>(src, ld_src, M, K, Mpad, reinterpret_cast<bf16*>(dst));
  ISX_LAUNCH_CHECK();
  return 0;
}

extern "C" int isx_pool7_flatten_pack(const isx_bf16* pool5, int B, int h, int w, int C, int Mpad, isx_bf16* dst, isx_stream stream) {
  ISX_REQUIRE(pool5 && dst && B > 0 && h > 0 && w > 0 && (C * 49) % 64 == 0 && Mpad % 64 == 0 && Mpad >= B, "isx_pool7_flatten_pack: bad arguments");
  const long total = static_cast<long>(Mpad) * (static_cast<long>(C) * 49 + 64);
  pool7_flatten_kernel<<<grid_for(total), 256, 0, S(stream)>>>(reinterpret_cast<const bf16*>(pool5), B, h, w, C, Mpad, reinterpret_cast<bf16*>(dst));
  ISX_LAUNCH_CHECK();
  return 0;
}

extern "C" int isx_linear_fwd(const isx_bf16* Xp, const isx_bf16* Wp, isx_bf16* outT, int Mpad, int Npad, int Kp, int relu,
                              isx_stream stream) {
  ISX_REQUIRE(Xp && Wp && outT, "isx_linear_fwd: null pointer");
  ISX_REQUIRE(Mpad % 64 == 0 && Mpad > 0 && Mpad <= 256 && Kp % 64 == 0 && Npad > 0, "isx_linear_fwd: Mpad (<= 256) and K' must be multiples of 64");
  ConvArgs a;
  a.in = reinterpret_cast<const bf16*>(Wp);       // "pixels" = output features
  a.weight = reinterpret_cast<const bf16*>(Xp);   // "output channels" = batch rows
  a.out = reinterpret_cast<bf16*>(outT);
  a.B = 1; a.H = 1; a.W = Npad; a.Cin = Kp; a.Cout = Mpad; a.ntaps = 1;
  a.relu = relu;
  a.force_bn = 64; a.force_mt = 1; a.force_stages = 6;   // weight streaming: deep ring, one 128-row tile per CTA
  return conv_tc(a, S(stream));
}

extern "C" int isx_transpose_pack(const isx_bf16* srcT, int N, int Mpad, int M, isx_bf16* dst, isx_stream stream) {
  ISX_REQUIRE(srcT && dst && N % 64 == 0 && Mpad % 64 == 0 && M <= Mpad, "isx_transpose_pack: bad arguments");
  dim3 grid((N + 31) / 32, (Mpad + 31) / 32);
  transpose_pack_kernel<<<grid, 256, 0, S(stream)>>>(reinterpret_cast<const bf16*>(srcT), N, Mpad, M, reinterpret_cast<bf16*>(dst));
  ISX_LAUNCH_CHECK();
  return 0;
}

extern "C" int isx_transpose_out(const isx_bf16* srcT, int N, int Mpad, int M, float* dst, isx_stream stream) {
  ISX_REQUIRE(srcT && dst && N > 0 && M > 0 && M <= Mpad, "isx_transpose_out: bad arguments");
  transpose_out_kernel<<<grid_for(static_cast<long>(M) * N), 256, 0, S(stream)>>>(reinterpret_cast<const bf16*>(srcT), N, Mpad, M, dst);
  ISX_LAUNCH_CHECK();
  return 0;
}
