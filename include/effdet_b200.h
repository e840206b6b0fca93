/* effdet_b200 -- C ABI of the B200-native (sm_100a) EfficientDet forward/backward hot path.
 *
 * The reference (toandaominh1997/EfficientDet.Pytorch @ fbe56e5) is 100% Python and has NO
 * native boundary of its own: every entry point below replaces a chain of ATen/cuDNN/torchvision
 * calls made by the cited reference lines.  The reference-side binding is the ctypes stub in
 * INTEGRATION.md (a Python reference binds through ctypes/`torch.autograd.Function`).
 *
 * Conventions
 *   - all tensors are fp32, dense, NHWC ("[B,H,W,C]") unless a comment says otherwise;
 *     pointers are device pointers owned by the caller (PyTorch's caching allocator);
 *   - nothing here allocates, frees, retains a pointer after return, or synchronises the device;
 *   - every call takes the CUDA device ordinal and the cudaStream_t to launch on;
 *   - return value 0 = launched; negative = error, text via effdet_last_error() (thread-local);
 *   - "+=" in a comment means the kernel ACCUMULATES into a caller-initialised buffer.
 */
#ifndef EFFDET_B200_H
#define EFFDET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EFFDET_OK 0
#define EFFDET_ERR_ARG (-1)
#define EFFDET_ERR_LAUNCH (-2)
#define EFFDET_ERR_DEVICE (-3)
#define EFFDET_ERR_UNSUPPORTED (-4)

#define EFFDET_ACT_NONE 0
#define EFFDET_ACT_RELU 1
#define EFFDET_ACT_SWISH 2
#define EFFDET_ACT_SIGMOID 3

#define EFFDET_FUSE_UP 0   /* second input is the coarser map, nearest x2 (models/bifpn.py:189) */
#define EFFDET_FUSE_POOL 1 /* second input is the finer map, 2x2/2 max-pool (models/bifpn.py:195,200) */

typedef void* effdet_stream_t; /* cudaStream_t */

int effdet_version(void);
const char* effdet_last_error(void);
uint64_t effdet_launch_count(void); /* kernels launched by this library in this process */
void effdet_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Dense convolution, k in {1,3}, stride 1, "same" zero padding, as an implicit GEMM
 *   y[b,p,n] = epilogue( sum_{tap,c} x[b, p+tap, c] * a_scale[b,c] * w[tap,c,n] )
 * Replaces F.conv2d + bias + BatchNorm(eval) + activation + residual in
 *   ConvModule.forward           models/module.py:507-515   (neck 1x1/3x3, head 3x3 + ReLU)
 *   RetinaHead.forward_single    models/retinahead.py:109-129 (retina_cls + sigmoid, retina_reg)
 *   MBConvBlock.forward          models/efficientnet.py:85,96-104 (expand / project 1x1)
 * and, with the flipped/transposed weight pack, the data gradient of each of them.
 * Epilogue order: v = acc + bias;  z = v (optional save);  v = v*scale + shift;  v = act(v);
 *                 v *= row_scale[b];  v += residual;  v = mask_src > 0 ? v : 0;  y = v.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const float* x;        int64_t x_bstride; /* [B,H,W,Cin]; elements between consecutive images */
    const float* w;                            /* packed [k*k][Cin][Cout] (effdet_pack_conv_weight) */
    float* y;              int64_t y_bstride; /* [B,H,W,Cout]; stride lets the head write straight
                                                  into the concatenated [B, sum(HWA), K] buffer
                                                  (kills torch.cat, models/efficientdet.py:64-65) */
    float* z;                                  /* optional raw (pre-affine) output, y's layout */
    const float* bias;                         /* [Cout] or NULL */
    const float* scale;    const float* shift; /* [Cout] eval-BN affine, or both NULL */
    const float* a_scale;                      /* [B,Cin] squeeze-excite gate on the input, or NULL */
    const float* row_scale;                    /* [B] drop-connect keep/keep_prob, or NULL */
    const float* residual; int64_t r_bstride;  /* [B,H,W,Cout] added last, or NULL */
    const float* mask_src; int64_t m_bstride;  /* [B,H,W,Cout]: ReLU-backward mask source, or NULL */
    int32_t B, H, W, Cin, Cout, ksize, act;
    const void* w_tc;      /* optional bf16 hi/lo planes from effdet_pack_conv_weight_tc: when set (and the
                              epilogue needs only bias/act/residual/mask) the layer runs on the tcgen05
                              tensor cores as a bf16x3 split-precision implicit GEMM (~2^-16 per product) */
    const float* in_scale; const float* in_shift; /* [Cin] or both NULL: the input is a RAW conv output and the
                              operand is swish(x*in_scale+in_shift) (eval-BN + swish applied while the tile is
                              staged, so the activated tensor never exists in HBM: MemoryEfficientSwish keeps
                              only the pre-activation too, models/utils.py:31-42); applied before a_scale */
    const void* x_planes;  /* optional (1x1 convs, tensor-core path): the input PRE-SPLIT into bf16 planes
                              [2][B*H*W][Cin] (plane 0 = hi, plane 1 = lo, x ~= hi + lo; Cin % 8 == 0), as written by
                              effdet_dwconv_bwd_fused; x may then be NULL */
} effdet_conv_args;
int effdet_conv2d(const effdet_conv_args* a, int device, effdet_stream_t stream);
/* The same convolution (shared w / w_tc / bias / act, channels, ksize) applied to `nlevels` (<= 8) feature maps of
 * different sizes in ONE launch -- RetinaHead.forward runs every layer on P3..P7 with the same weights
 * (models/retinahead.py:131-132).  Falls back to one launch per level in exact-fp32 mode. */
int effdet_conv2d_multi(const effdet_conv_args* levels, int nlevels, int device, effdet_stream_t stream);

/* Weight gradient (and optional bias gradient) of the convolution above.
 *   dw[n,c,ky,kx] += sum_{b,p} x[b,p+tap,c]*a_scale[b,c] * dy[b,p,n]     (OIHW, as .grad)
 *   dbias[n]      += sum_{b,p} dy[b,p,n]
 * Replaces cuDNN bwd-filter reached through autograd. */
typedef struct {
    const float* x;   int64_t x_bstride;
    const float* dy;  int64_t dy_bstride;
    float* dw;        /* [Cout,Cin,k,k]  += */
    float* dbias;     /* [Cout] += , or NULL */
    const float* a_scale;
    int32_t B, H, W, Cin, Cout, ksize;
    int32_t precision;     /* 0: exact fp32 on the CUDA cores; 1: bf16x3 on the tcgen05 tensor cores */
    void* ws_x;            /* precision 1: bf16 workspaces for the pre-split operands, */
    void* ws_dy;           /*   2*B*H*W*kpad(Cin) resp. 2*B*H*W*kpad(Cout) elements (TMA-fed kernel);
                              NULL -> the gather-producer tensor-core kernel is used instead */
    const float* in_scale; const float* in_shift; /* [Cin] or both NULL: x is a raw conv output, the operand is
                              swish(x*in_scale+in_shift)*a_scale (see effdet_conv_args) */
    const void* dy_planes; /* optional (precision 1): dy PRE-SPLIT into bf16 planes [2][B*H*W][pitch(Cout)], as written
                              by effdet_dwconv_bwd_fused / effdet_conv_planes_multi: no split pass over dy, ws_dy unused,
                              dy may be NULL (dbias must be NULL) */
    const void* x_planes;  /* optional, likewise for x ([2][B*H*W][pitch(Cin)]; no a_scale / in_scale): ws_x unused */
} effdet_wgrad_args;
int effdet_conv2d_wgrad(const effdet_wgrad_args* a, int device, effdet_stream_t stream);
/* Weight gradient of one shared-weight layer accumulated over `nlevels` feature maps in one launch (all levels
 * must name the same dw / dbias); falls back to one launch per level when a level cannot use the TMA path. */
int effdet_conv2d_wgrad_multi(const effdet_wgrad_args* levels, int nlevels, int device, effdet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * The same dense convolution with activations kept in HBM as bf16 hi/lo PLANES [2][B][H][W][pitch] (x ~= hi + lo, the
 * tensor-core operand format; pitch = channels rounded up to 8): the RetinaHead towers end to end
 * (models/retinahead.py:67-132).  The im2col gather is a TMA load (5-D tensor map, tap = coordinate offset, hardware
 * zero fill = padding), the epilogue writes the next layer's planes (and / or fp32), so no layer re-splits its input and
 * no weight gradient needs a split pass.  All levels of one call share weights / bias (RetinaHead.forward).
 *   epilogue: v = acc + bias; v = act(v); v += residual; v = mask > 0 ? v : 0; store planes and / or fp32;
 *             colsum[n] += sum over pixels of v   (the bias gradient of the layer that produced this layer's input
 *             gradient -- a data-gradient launch hands it over for free)
 * Needs effdet_wgrad_tc_geometry_ok(B,H,W) for every level.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const void* x_planes;                       /* [2][B][H][W][pitch(Cin)] bf16 */
    const void* w_tc;                           /* effdet_pack_conv_weight_tc pack (forward or data-gradient) */
    const float* bias;                          /* [Cout] or NULL */
    float* y;              int64_t y_bstride;   /* fp32 output [B][H*W][Cout] (image stride y_bstride) or NULL */
    void* y_planes;                             /* planes output [2][B][H][W][pitch(Cout)] or NULL */
    const void* mask_planes;                    /* ReLU-backward mask source, planes of pitch(Cout), or NULL */
    const float* residual; int64_t r_bstride;   /* fp32 [B][H*W][Cout] added before the mask, or NULL */
    float* colsum;                              /* [Cout] += or NULL */
    int32_t B, H, W, Cin, Cout, ksize, act;
} effdet_conv_planes_args;
int effdet_conv_planes_multi(const effdet_conv_planes_args* levels, int nlevels, int device, effdet_stream_t stream);
/* fp32 [B][HW][C] (image stride x_bstride) -> planes [2][B*HW][pitch(C)]; with prob != NULL the value is first
 * multiplied by p*(1-p) (sigmoid backward of the classification head, models/retinahead.py:121); with colsum != NULL
 * the per-channel sums of what was written are accumulated (bias gradient) in the same pass */
int effdet_to_planes(const float* x, int64_t x_bstride, const float* prob, int64_t p_bstride, void* planes, float* colsum,
                     int B, int HW, int C, int device, effdet_stream_t stream);

/* OIHW -> [k*k][Cin][Cout] (forward) and, if w_dgrad != NULL, the 180-degree-rotated transpose
 * [k*k][Cout][Cin] that turns the data gradient into the same implicit GEMM. */
int effdet_pack_conv_weight(const float* w_oihw, float* w_fwd, float* w_dgrad, int Cout, int Cin, int ksize,
                            int device, effdet_stream_t stream);

/* OIHW fp32 -> pre-split bf16 planes for the tensor-core path, K-major and zero padded:
 *   w_fwd   [2][Cout][k*k][kpad(Cin)]   (plane 0 = hi, plane 1 = lo, w ~= hi + lo)
 *   w_dgrad [2][Cin][k*k][kpad(Cout)]   (rotated 180 degrees and transposed), may be NULL
 * kpad(c) = effdet_conv_tc_kpad(c) = c rounded up to a multiple of 64. */
int effdet_conv_tc_kpad(int channels);
/* 1 when the TMA-fed tensor-core weight-gradient kernel can tile a [B,H,W,*] map into pixel boxes (required before
 * handing it pre-split operands, effdet_wgrad_args.dy_planes); no device work */
int effdet_wgrad_tc_geometry_ok(int B, int H, int W);
int effdet_pack_conv_weight_tc(const float* w_oihw, void* w_fwd, void* w_dgrad, int Cout, int Cin, int ksize,
                               int device, effdet_stream_t stream);

/* out[n] += sum_m x[m,n]   (bias gradients, BN beta gradients) */
int effdet_colsum(const float* x, float* out, int64_t M, int N, int device, effdet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Stem: 3x3 stride-2 conv on the NCHW image with TF-"SAME" pad (0,1,0,1), eval-BN, swish.
 * Replaces EfficientNet.extract_features stem, models/efficientnet.py:193 (+ utils.py:126-155).
 *   x [B,3,H,W] NCHW  ->  z (raw conv) and y = swish(z*scale+shift), both [B,H/2,W/2,C0] NHWC
 *   (y may be NULL: the consumer applies BN+swish while staging z)
 * ------------------------------------------------------------------------------------------ */
int effdet_stem_fwd(const float* x_nchw, const float* w_oihw, const float* scale, const float* shift, float* z,
                    float* y, int B, int H, int W, int C0, int device, effdet_stream_t stream);
/* dw[C0,3,3,3] += sum x * dz   (the image needs no data gradient) */
int effdet_stem_wgrad(const float* x_nchw, const float* dz, float* dw_oihw, int B, int H, int W, int C0,
                      int device, effdet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Depthwise k x k (k in {3,5}), stride in {1,2}, asymmetric TF-"SAME" pad, eval-BN, swish.
 * Replaces MBConvBlock.forward depthwise phase, models/efficientnet.py:87.
 *   x [B,H,W,C] -> z raw, y = swish(z*scale+shift), both [B,Ho,Wo,C];  w_kkc = [k][k][C]
 * ------------------------------------------------------------------------------------------ */
int effdet_dwconv_fwd(const float* x, const float* w_kkc, const float* scale, const float* shift, float* z, float* y,
                      int B, int H, int W, int C, int k, int stride, int pad_t, int pad_l, int Ho, int Wo,
                      int device, effdet_stream_t stream);
int effdet_dwconv_bwd_data(const float* dz, const float* w_kkc, float* dx, int B, int H, int W, int C, int k,
                           int stride, int pad_t, int pad_l, int Ho, int Wo, int device, effdet_stream_t stream);
/* dw[C,1,k,k] += ... */
int effdet_dwconv_bwd_weight(const float* x, const float* dz, float* dw_c1kk, int B, int H, int W, int C, int k,
                             int stride, int pad_t, int pad_l, int Ho, int Wo, int device, effdet_stream_t stream);
int effdet_pack_dw_weight(const float* w_c1kk, float* w_kkc, int C, int k, int device, effdet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Depthwise phase of MBConvBlock.forward with only PRE-activations in HBM (the reference's
 * MemoryEfficientSwish saves the pre-activation only, models/utils.py:31-42; block: models/efficientnet.py:85-94).
 * Pads must be the reference's static ones (models/utils.py:126-149): k3s1 1, k5s1 2, k3s2 0, k5s2 1 (top/left).
 *   fwd: a0 = in_scale ? swish(x*in_scale+in_shift) : x        (BN0+swish applied while the tile is staged)
 *        z  = depthwise(a0)                                    raw output, the only tensor written
 *        se_sum[b,c] += se_alpha * sum_pixels swish(z*scale+shift)   (squeeze-excite mean, an epilogue by-product)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const float* x;                                /* [B,H,W,C] raw expand-conv output z0, or the block input */
    const float* in_scale; const float* in_shift;  /* [C] folded BN0, or both NULL (x used as is) */
    const float* w_kkc;                            /* [k][k][C] */
    const float* scale;    const float* shift;     /* [C] folded BN1 */
    float* z;                                      /* [B,Ho,Wo,C] */
    float* se_sum;                                 /* [B,C] += se_alpha * sum  (se_alpha = 1/(Ho*Wo) makes it the SE mean) */
    int32_t B, H, W, C, k, stride, pad_t, pad_l, Ho, Wo;
    float se_alpha;
} effdet_dw_fwd_args;
int effdet_dwconv_fwd_fused(const effdet_dw_fwd_args* a, int device, effdet_stream_t stream);
/*   bwd, one pass over the expanded tensor (replaces BN1-swish backward + depthwise weight/data gradient + BN0-swish
 *   backward, models/utils.py:38-42 through autograd):
 *        du1 = (dq*gate + dmean*inv_hw) * swish'(z1*scale1+shift1);  dgamma1 += sum du1*(z1-mean1)*rstd1;  dbeta1 += sum du1
 *        dz1 = du1*scale1  (kept on chip);  da0 = depthwise^T(dz1);  dw[c,ky,kx] += sum a0 * dz1
 *        scale0 ? { du0 = da0*swish'(x*scale0+shift0); dgamma0, dbeta0 += ...; dx = du0*scale0 } : dx = da0 */
typedef struct {
    const float* dq;       /* [B,Ho,Wo,C] gradient w.r.t. (a1*gate), i.e. the project conv's data gradient */
    const float* z1;       /* [B,Ho,Wo,C] raw depthwise output */
    const float* gate;     const float* dmean;   /* [B,C] SE gate, gradient w.r.t. the SE mean */
    const float* scale1;   const float* shift1;  const float* mean1;  const float* rstd1;   /* [C] BN1 */
    const float* x;        /* [B,H,W,C] raw z0 (scale0 != NULL) or the block input */
    const float* scale0;   const float* shift0;  const float* mean0;  const float* rstd0;   /* [C] BN0 or all NULL */
    const float* w_kkc;    /* [k][k][C] */
    float* dx;             /* [B,H,W,C] */
    float* dw;             /* [C,1,k,k] += */
    float* dgamma1; float* dbeta1; float* dgamma0; float* dbeta0;   /* [C] += (BN0 ones may be NULL without BN0) */
    float inv_hw;          /* 1/(Ho*Wo) */
    int32_t B, H, W, C, k, stride, pad_t, pad_l, Ho, Wo;
    void* dx_planes;       /* optional: write dx as bf16 hi/lo planes [2][B*H*W][C] (dx ~= hi + lo) INSTEAD of fp32 dx
                              (dx may then be NULL): the form the tensor-core data / weight gradients of the expand conv
                              consume directly (effdet_conv_args.x_planes, effdet_wgrad_args.dy_planes); C % 8 == 0 */
} effdet_dw_bwd_args;
int effdet_dwconv_bwd_fused(const effdet_dw_bwd_args* a, int device, effdet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Backward of  y = act(BN_eval(z)) [* row_scale]  with frozen statistics but trainable affine
 * (models/efficientdet.py:88-92; swish backward models/utils.py:38-42):
 *   g  = dy*row_scale[b]            (or, SE mode:  g = dy*gate[b,c] + dmean[b,c]*inv_hw)
 *   du = g * act'(z*scale+shift);  dgamma += sum du*(z-mean)*rstd;  dbeta += sum du;  dz = du*scale
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const float* dy;    /* [B,HW,C] */
    const float* z;     /* [B,HW,C] raw conv output */
    float* dz;          /* [B,HW,C] */
    const float* scale; const float* shift; const float* mean; const float* rstd; /* [C] */
    float* dgamma;      float* dbeta;   /* [C] += */
    const float* row_scale; /* [B] or NULL */
    const float* gate;      /* [B,C] or NULL */
    const float* dmean;     /* [B,C] or NULL */
    float inv_hw;
    int32_t B, HW, C, act;  /* act: EFFDET_ACT_NONE or EFFDET_ACT_SWISH */
} effdet_bnact_bwd_args;
int effdet_bnact_bwd(const effdet_bnact_bwd_args* a, int device, effdet_stream_t stream);

/* Frozen BatchNorm2d (eval mode even while training, models/efficientdet.py:88-92) folded to a
 * per-channel affine: rstd = 1/sqrt(var+eps), scale = gamma*rstd, shift = beta - mean*scale. */
int effdet_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* scale,
                   float* shift, float* rstd, int C, int device, effdet_stream_t stream);
/* out = a + b (skip-connection gradient join) ;  dz = y > 0 ? dy : 0 (stand-alone ReLU backward) */
int effdet_add(const float* a, const float* b, float* out, int64_t n, int device, effdet_stream_t stream);
int effdet_relu_bwd(const float* dy, const float* y, float* dz, int64_t n, int device, effdet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Squeeze-excite (models/efficientnet.py:90-94).
 *   effdet_spatial_reduce : out[b,c] += alpha * sum_hw a[b,hw,c] * (b2 ? b2[b,hw,c] : 1)
 *   effdet_se_gate_fwd    : s_pre = W1*mean+b1 ; gate = sigmoid(W2*swish(s_pre)+b2)
 *   effdet_se_gate_bwd    : given dgate -> dmean, and += into dW1,db1,dW2,db2
 * W1 = _se_reduce.weight [S,C], W2 = _se_expand.weight [C,S].
 * ------------------------------------------------------------------------------------------ */
int effdet_spatial_reduce(const float* a, const float* b2, float* out, float alpha, int B, int HW, int C,
                          int device, effdet_stream_t stream);
/*   effdet_spatial_reduce_act : out[b,c] += alpha * sum_hw a[b,hw,c] * swish(z[b,hw,c]*scale[c]+shift[c])
 *   (gradient w.r.t. the SE gate from the raw depthwise output: the activated tensor is recomputed, not stored) */
int effdet_spatial_reduce_act(const float* a, const float* z, const float* scale, const float* shift, float* out,
                              float alpha, int B, int HW, int C, int device, effdet_stream_t stream);
int effdet_se_gate_fwd(const float* mean, const float* w1, const float* b1, const float* w2, const float* b2,
                       float* s_pre, float* gate, int B, int C, int S, int device, effdet_stream_t stream);
int effdet_se_gate_bwd(const float* dgate, const float* mean, const float* s_pre, const float* gate,
                       const float* w1, const float* w2, float* dmean, float* dw1, float* db1, float* dw2,
                       float* db2, float* ws /* B*(C+S) floats of scratch */, int B, int C, int S, int device,
                       effdet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * BiFPN fast-normalised fusion node input (BiFPNModule.forward, models/bifpn.py:177-201):
 *   r = relu(w_col); n = r/(sum r + eps); out = sum_j n_j*in_j / (sum_j n_j + eps)
 *   in_0 = a; in_1 = up2(b) or maxpool2(b); in_2 = c (3-input nodes only)
 * The raw weight column is read as w[j*w_stride], j < (c ? 3 : 2).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const float* a; const float* b; const float* c;
    const float* w; int32_t w_stride; float eps;
    float* out;
    int32_t B, H, W, C, mode; /* H,W = resolution of a/out */
    void* out_planes;         /* optional: the fused map as bf16 hi/lo planes [2][B*H*W][pitch(C)] (the node conv's TMA
                                 operand, effdet_conv_planes_multi); `out` may then be NULL */
} effdet_fuse_args;
int effdet_bifpn_fuse_fwd(const effdet_fuse_args* a, int device, effdet_stream_t stream);

typedef struct {
    const float* dout;
    const float* a; const float* b; const float* c;
    const float* w; int32_t w_stride; float eps;
    float* da; float* db; float* dc;  /* input gradients (dc NULL for 2-input nodes) */
    int32_t acc_a, acc_b, acc_c;      /* 0: overwrite, 1: += */
    float* dw;                        /* gradient of the raw weights, indexed like w, += */
    float* scratch;                   /* [4] floats, zero on entry */
    int32_t B, H, W, C, mode;
} effdet_fuse_bwd_args;
int effdet_bifpn_fuse_bwd(const effdet_fuse_bwd_args* a, int device, effdet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * FocalLoss.forward (models/losses.py:32-152): IoU assignment + focal BCE + smooth-L1, no host
 * sync, no per-image Python loop.
 *   cls [B,A,K] probabilities, reg [B,A,4], anchors [A,4] (x1,y1,x2,y2), annots [B,G,5] (-1 pad)
 *   fwd: losses[0] = mean_b cls_loss_b, losses[1] = mean_b reg_loss_b;
 *        assign_ws [B,A] int32 (>=0 matched annotation row, -1 background, -2 ignored, -3 image
 *        without boxes) and stats_ws [B,4] float (npos, cls_sum, reg_sum, -) are kept for bwd.
 *   bwd: dcls = gout[0] * d losses[0]/d cls ; dreg = gout[1] * d losses[1]/d reg
 *        (gout: the two upstream gradients, read from DEVICE memory -> no sync).
 * ------------------------------------------------------------------------------------------ */
int effdet_focal_loss_fwd(const float* cls, const float* reg, const float* anchors, const float* annots,
                          float* losses, int32_t* assign_ws, float* stats_ws, int B, int A, int K, int G,
                          float alpha, float gamma, int device, effdet_stream_t stream);
int effdet_focal_loss_bwd(const float* cls, const float* reg, const float* anchors, const float* annots,
                          const float* gout, const int32_t* assign_ws, const float* stats_ws, float* dcls,
                          float* dreg, int B, int A, int K, int G, float alpha, float gamma, int device,
                          effdet_stream_t stream);
/* y = g * p*(1-p) (sigmoid backward) ; y may alias g */
int effdet_sigmoid_bwd(const float* g, const float* p, float* y, int64_t n, int device, effdet_stream_t stream);
/* ------------------------------------------------------------------------------------------
 * Inference post-processing of image 0 (models/efficientdet.py:70-86, models/module.py:24-67,
 * torchvision.ops.nms): decode + clip + class max + threshold + sort + greedy NMS, all on the
 * device; the host only reads the two counters to size its output tensors.
 * ------------------------------------------------------------------------------------------ */
/* boxes[A,4], scores[A], classes[A]; keys[npad] (npad = pow2 >= A) sorted ascending so that
 * entry j < *count is the j-th best candidate (score desc, anchor index asc); count[0] = #cands. */
int effdet_detect_candidates(const float* cls, const float* reg, const float* anchors, float* boxes, float* scores,
                             int32_t* classes, uint64_t* keys, int32_t* count, int A, int K, int npad,
                             float img_w, float img_h, float threshold, int device, effdet_stream_t stream);
/* mask_ws: n * ceil(n/64) uint64; keep_idx[n] int32 (anchor indices, best first); nkeep[0] */
int effdet_nms(const float* boxes, const uint64_t* keys, int n, double iou_threshold, uint64_t* mask_ws,
               int32_t* keep_idx, int32_t* nkeep, int device, effdet_stream_t stream);
/* gathers scores/classes(int64)/boxes rows listed in keep_idx */
int effdet_gather_detections(const float* boxes, const float* scores, const int32_t* classes,
                             const int32_t* keep_idx, int nkeep, float* out_scores, int64_t* out_classes,
                             float* out_boxes, int device, effdet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * the optimizer step that follows backward in the reference loop,
 *   clip_grad_norm_(parameters, max_norm) + AdamW.step()        (train.py:115-118, train.py:268)
 * as two multi-tensor launches.  The tables live in DEVICE memory: per tensor t its fp32 data pointers and
 * element count, per chunk c the tensor it belongs to and its element offset; one CTA per chunk.
 *   effdet_multi_sumsq      : norm_sq[0] += sum_t sum_i g_t[i]^2
 *   effdet_multi_clip_adamw : g *= min(1, max_norm/(sqrt(norm_sq)+1e-6)) (max_norm <= 0: no clipping);
 *                             p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
 *                             p -= lr/bias_c1 * m / (sqrt(v)/sqrt(bias_c2) + eps)     (torch.optim.AdamW)
 * ------------------------------------------------------------------------------------------ */
int effdet_multi_sumsq(const uint64_t* g_ptrs, const int64_t* numels, const int32_t* chunk_tensor,
                       const int64_t* chunk_off, int nchunks, int chunk, float* norm_sq, int device,
                       effdet_stream_t stream);
int effdet_multi_clip_adamw(const uint64_t* p_ptrs, const uint64_t* g_ptrs, const uint64_t* m_ptrs,
                            const uint64_t* v_ptrs, const int64_t* numels, const int32_t* chunk_tensor,
                            const int64_t* chunk_off, int nchunks, int chunk, const float* norm_sq, float max_norm,
                            float lr, float beta1, float beta2, float eps, float weight_decay, float bias_c1,
                            float bias_c2, int write_clipped_grad, int device, effdet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * the step BEFORE the hot path, on the device.  Normalizer + Augmenter (horizontal flip) + zero pad
 * to the common size + collater + `.cuda().float()` (datasets/augmentation.py:69-91,111-150, train.py:105-106), bit-identical
 * to NumPy: (float32(u8) - mean) / std evaluated in float64, then cast to float32.  cv2.resize is not reproduced: images
 * enter at their final resolution (h, w <= S).
 *   pixels  : uint8 HWC images back to back, image b starts at byte offsets[b] and is hw[2b] x hw[2b+1] x 3
 *   flip    : [B] or NULL;  out_nchw : float32 [B,3,S,S];  mean3 / std3 : 3 host doubles
 *   annotations: rows [sum n_b, 5] float64, image b owns rows row_off[b] .. row_off[b+1]; out float32 [B,G,5], -1 padded;
 *   scale [B] float64 or NULL (Resizer's box scale), width [B] = image widths (needed with flip)
 * ------------------------------------------------------------------------------------------ */
int effdet_normalize_pad(const uint8_t* pixels, const int64_t* offsets, const int32_t* hw, const uint8_t* flip,
                         float* out_nchw, int B, int S, const double* mean3, const double* std3, int device,
                         effdet_stream_t stream);
int effdet_collate_annots(const double* rows, const int32_t* row_off, const double* scale, const uint8_t* flip,
                          const int32_t* width, float* out, int B, int G, int device, effdet_stream_t stream);
/* the step AFTER NMS (eval.py:105-128) for one image: boxes /= scale, score > threshold,
 * top-max_det by score (ties: lower input index), split per label.  out_dets [max_det,5] (x1,y1,x2,y2,score) grouped by
 * label ascending and in score order inside a label, out_labels [max_det], class_offsets [num_classes+1] (rows of label c
 * are class_offsets[c] .. class_offsets[c+1]), count[0] = selected rows. */
int effdet_eval_select(const float* scores, const int64_t* labels, const float* boxes, int n, float scale,
                       float score_threshold, int max_det, int num_classes, float* out_dets, int32_t* out_labels,
                       int32_t* class_offsets, int32_t* count, int device, effdet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Layout plumbing at the module boundary (callers see logical NCHW, kernels are NHWC).
 * ------------------------------------------------------------------------------------------ */
int effdet_nchw_to_nhwc(const float* x, float* y, int B, int C, int H, int W, int device, effdet_stream_t stream);
int effdet_nhwc_to_nchw(const float* x, float* y, int B, int C, int H, int W, int device, effdet_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EFFDET_B200_H */
