/* libisx -- B200 (sm_100a) kernels for the iris-masked neural-style-transfer hot path of
 * AnonymWriter/Iris-Style-Transfer.  C ABI: plain device pointers, explicit shapes, a cudaStream_t;
 * every function returns 0 on success and != 0 on failure (text via isx_last_error()).  Nothing
 * here allocates device memory: the caller (e.g. PyTorch's caching allocator) owns every buffer,
 * including workspaces whose size is queried first.  Thread-compatible; no global device state.
 *
 * The reference has no FFI/plugin layer (pure PyTorch): each entry point names the reference
 * call site (relative to the reference repo root) whose library kernels it replaces.
 *
 * Layouts: activations / gradients are NHWC bf16 ("isx_bf16" = 2-byte bfloat16), images and
 * image gradients are NCHW fp32 exactly as the reference passes them, Gram matrices [B,C,C] fp32.
 */
#ifndef ISX_H_
#define ISX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISX_VERSION 100

typedef uint16_t isx_bf16;   /* storage type of a bfloat16 */
typedef void* isx_stream;    /* cudaStream_t */

/* ---- status ------------------------------------------------------------------------------- */
const char* isx_last_error(void);
int isx_version(void);
/* fails unless `device` is an sm_100 part: there is no fallback path */
int isx_device_check(int device);

/* ---- weights: torchvision vgg19.features Conv2d parameters (models/vgg/vgg.py:43-49) ---------
 * fp32 OIHW [Cout,Cin,3,3] -> bf16 [9][Cout][Cin] for the forward conv and the 180-degree-rotated,
 * transposed [9][Cin][Cout] for dgrad.  Either output may be NULL. */
int isx_pack_conv3x3_weights(const float* w_oihw, int Cout, int Cin, isx_bf16* w_fwd, isx_bf16* w_dgrad,
                             isx_stream stream);

/* ---- K0+K1 head: Normalize -> [*mask] -> conv1_1 + bias + ReLU (models/vgg/vgg.py:81-87) ------
 * x: fp32 [B,xc,H,W], xc in {1,3} (1 broadcasts like transforms.Normalize does);
 * mask: NULL or fp32 [mask_b,1,H,W] with mask_b in {1,B}; w: fp32 [64,3,3,3]; out: bf16 [B,H,W,64]. */
int isx_conv1_1_fwd(const float* x, int xc, const float* mask, int mask_b, const float* w, const float* bias,
                    isx_bf16* out, int B, int H, int W, isx_stream stream);
/* autograd tail of the same (pipelines.py:90): dY bf16 [B,H,W,64] (already ReLU-masked) ->
 * d(loss)/dx fp32 [B,xc,H,W], including Normalize's 1/std and the optional mask. */
int isx_conv1_1_dgrad(const isx_bf16* dy, const float* w, const float* mask, int mask_b, float* dx, int xc, int B,
                      int H, int W, isx_stream stream);

/* ---- K1: Conv2d 3x3 s1 p1 + bias + ReLU on tcgen05 (models/vgg/vgg.py:87) ---------------------
 * in bf16 [B,H,W,Cin], w_fwd from isx_pack_conv3x3_weights, bias fp32 [Cout], out bf16 [B,H,W,Cout].
 * Cin, Cout multiples of 64.  tile_cfg: 0 = heuristic, else BN*100 + MT*10 + stages (test / tuning hook). */
int isx_conv3x3_bias_relu_fwd(const isx_bf16* in, const isx_bf16* w_fwd, const float* bias, isx_bf16* out, int B,
                              int H, int W, int Cin, int Cout, int relu, int tile_cfg, isx_stream stream);

/* ---- K3: conv dgrad (+ tap gradient, x ReLU mask) on tcgen05 (loss.backward(), pipelines.py:90)
 * dy bf16 [B,H,W,Cout] -> dx bf16 [B,H,W,Cin] for the forward conv Cin->Cout; weights frozen so
 * there is no wgrad (models/vgg/vgg.py:52-53).  Epilogue, all optional:
 *   dx += add_grad[B,H,W,Cin]                      (Gram tap gradient of the layer below)
 *   dx += aff_a[b,c] + aff_b[b,c]*act              (BN-statistics tap gradient; needs relu_act)
 *   dx  = relu_act > 0 ? dx : 0                    (ReLU backward of the layer below) */
int isx_conv3x3_dgrad(const isx_bf16* dy, const isx_bf16* w_dgrad, isx_bf16* dx, int B, int H, int W, int Cin,
                      int Cout, const isx_bf16* relu_act, const isx_bf16* add_grad, const float* aff_a,
                      const float* aff_b, int tile_cfg, isx_stream stream);

/* ---- K2: MaxPool2d(2,2) fwd / bwd (bwd fused with the ReLU mask of the pre-pool activation) ---- */
int isx_maxpool2x2_fwd(const isx_bf16* in, isx_bf16* out, int B, int H, int W, int C, isx_stream stream);
int isx_maxpool2x2_bwd(const isx_bf16* dy_pooled, const isx_bf16* act_prepool, isx_bf16* dx, int B, int H, int W,
                       int C, isx_stream stream);

/* ---- K4: Gram matrix (utils.py:242-257 GramMatrix) ---------------------------------------------
 * feat bf16 [B,HW,C] (NHWC flattened), C in {64,128,256,512}.  G = F^T F * inv_n (fp32 [B,C,C]).
 * workspace: isx_gram_workspace_bytes(B,HW,C) bytes.  With a target (fp32 [target_b,C,C],
 * target_b in {1,B}): loss[b] += loss_scale * sum((G-T)^2) (double, caller zeroes) and
 * D = grad_scale * (G - T) in bf16 [B,C,C] for isx_gram_bwd.  G_out / target / loss / D may be NULL. */
int64_t isx_gram_workspace_bytes(int B, int HW, int C);
int isx_gram_fwd(const isx_bf16* feat, int B, int HW, int C, float inv_n, void* workspace, float* G_out,
                 const float* target, int target_b, double loss_scale, double* loss, float grad_scale,
                 isx_bf16* D_out, isx_stream stream);
/* ---- K5: Gram backward dF[b] = F[b] . D[b] (autograd of utils.py:253-256), tcgen05 1x1 mode.
 * feat bf16 [B,H,W,C], D bf16 [B,C,C] (symmetric), dF bf16 [B,H,W,C]; optional relu mask. */
int isx_gram_bwd(const isx_bf16* feat, const isx_bf16* D, isx_bf16* dF, int B, int H, int W, int C,
                 const isx_bf16* relu_act, isx_stream stream);

/* ---- K6: content MSE (utils.py:285-290) and mean/std statistics (utils.py:337-354, classifiers.py:71)
 * content: loss[b] += loss_scale * sum((p-t)^2); grad = grad_scale*(p-t)*(p>0) (bf16, may be NULL). */
int isx_content_mse_fwd_bwd(const isx_bf16* pred, const isx_bf16* target, int target_b, isx_bf16* grad, int B,
                            int64_t per_image, double loss_scale, float grad_scale, double* loss, isx_stream stream);
/* sums: double [B,C,2] workspace.  mean/std fp32 [B,C] (std unbiased).  With targets: loss[b] +=
 * loss_scale * sum_c[(mu-mu_t)^2+(sd-sd_t)^2] and affine tap-gradient coefficients aff_a/aff_b [B,C]. */
int isx_bn_stats_fwd(const isx_bf16* feat, int B, int64_t HW, int C, double* sums, float* mean, float* std_,
                     const float* t_mean, const float* t_std, int target_b, double loss_scale, double grad_scale,
                     double* loss, float* aff_a, float* aff_b, isx_stream stream);
/* out = (g + add + aff) * (act > 0) -- tap gradient at a layer that no dgrad epilogue feeds; g/add/aff may be NULL */
int isx_tap_add_mask(const isx_bf16* g, const isx_bf16* add, const float* aff_a, const float* aff_b,
                     const isx_bf16* act, isx_bf16* out, int B, int64_t HW, int C, isx_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* ISX_H_ */
