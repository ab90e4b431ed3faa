/*
 * pcv_b200.h — C ABI of libpcv_b200.so: the B200 (sm_100a) eval-mode convolution path for pytorchcv models.
 *
 * The reference (osmr/pytorchcv) has no FFI: its "operator API" for this path is the nn.Module.forward of the
 * blocks in pytorchcv/models/common/{conv,att,activ,norm}.py, which issue eager torch ops.  Each entry point below
 * replaces the torch-op sequence of one such forward; the reference location is cited per function.
 *
 * Conventions
 *  - Every function returns 0 on success, a negative pcv_status otherwise; pcv_last_error() gives the message
 *    (thread-local).  Nothing throws across the ABI; nothing allocates device memory (callers own all buffers).
 *  - All pointers are DEVICE pointers unless a name ends in _host.  Activations are NHWC ("pixels x channels")
 *    with an explicit channel pitch (elements between consecutive pixels; 0 means "= channels"), so a tensor can
 *    be a channel slice of a wider buffer (torch.cat on dim 1 becomes a write at a channel offset).
 *  - `dtype` selects the arithmetic tier: PCV_BF16 / PCV_F16 = 16-bit storage, fp32 accumulate/epilogue (tcgen05 tensor
 *    cores for dense/grouped conv; comments below that say "bf16 tier" apply to both); PCV_F32 = fp32 storage and true
 *    fp32 FMA (the <=1e-4 tier).
 *  - `plan`: when non-NULL the op is RECORDED into the plan (stream ignored) and runs on every pcv_plan_run();
 *    when NULL it is launched immediately on `stream`.  Kernels never synchronise the host.
 */
#ifndef PCV_B200_H
#define PCV_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define PCV_API __attribute__((visibility("default")))
#else
#define PCV_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pcv_plan pcv_plan;  /* opaque: a flat list of fused-kernel launches */
typedef void* pcv_stream;          /* cudaStream_t */

enum pcv_status {
  PCV_OK = 0,
  PCV_ERR_INVALID = -1,      /* bad argument / unsupported shape (the AssertionError / ValueError analogue) */
  PCV_ERR_UNSUPPORTED = -2,  /* valid in the reference but outside this path (NotImplementedError analogue) */
  PCV_ERR_CUDA = -3,         /* a CUDA runtime / driver call failed */
  PCV_ERR_NO_DEVICE = -4     /* no sm_100 device: there is no CPU fallback */
};

enum pcv_dtype {
  PCV_BF16 = 0,  /* bf16 storage, fp32 accumulate / epilogue (tcgen05 kind::f16, bf16 operands) */
  PCV_F32 = 1,   /* fp32 storage, true fp32 FMA: the <= 1e-4 tier */
  PCV_F16 = 2    /* IEEE fp16 storage, fp32 accumulate / epilogue: the same kernels and MMA rate as PCV_BF16 with 3 more
                    mantissa bits (MobileNetV2-class networks meet 2e-2 end to end here); |x| > 65504 overflows */
};

/* element type of the NCHW image handed to the network edge (pcv_*_ingest_ex): the reference's fp32 tensor, or a
 * narrower copy of it that costs half / a quarter of the host->device bytes */
enum pcv_image_type { PCV_IMG_F32 = 0, PCV_IMG_BF16 = 1, PCV_IMG_F16 = 2, PCV_IMG_U8 = 3 };

/* activations of pytorchcv/models/common/activ.py:188-222 (create_activation_layer) */
enum pcv_act {
  PCV_ACT_NONE = 0,
  PCV_ACT_RELU = 1,     /* nn.ReLU     activ.py:50-64   */
  PCV_ACT_RELU6 = 2,    /* nn.ReLU6    activ.py:67-81   */
  PCV_ACT_SIGMOID = 3,  /* nn.Sigmoid  activ.py:123-132 */
  PCV_ACT_SWISH = 4,    /* Swish       activ.py:16-21   */
  PCV_ACT_HSWISH = 5,   /* HSwish      activ.py:33-47   */
  PCV_ACT_HSIGMOID = 6, /* HSigmoid    activ.py:24-30   */
  PCV_ACT_LEAKY_RELU = 7, /* nn.LeakyReLU activ.py:101-120: x >= 0 ? x : act_param * x (dense / grouped convs; the
                             per-channel nn.PReLU of activ.py:84-98 is pcv_channel_affine_act's `slope`) */
  PCV_ACT_CLAMP01 = 8     /* GhostHSigmoid ghostnet.py:18-24: clamp(x, 0, 1) - SE gates (pcv_se_excite*) only */
};

enum pcv_conv_flags {
  PCV_CONV_OUT_F32 = 1,      /* bf16 tier only: store the result as fp32 (classifier logits) */
  PCV_CONV_FORCE_SIMT = 2,   /* bf16 tier only: use the CUDA-core kernel (cross-check for the tcgen05 path) */
  PCV_CONV_A_IM2COL = 4,     /* bf16 tier only: use the im2col TMA descriptor even for 1x1 stride-1 */
  PCV_CONV_IN_OVERLAP = 8,   /* x is an overlapping-window VIEW: in_pitch < Cin is allowed (space-to-depth stem) */
  PCV_CONV_SE_GATE = 64,     /* 16-bit tiers, pair tcgen05 kernel (ask pcv_conv_se_gate_ok): `workspace` of pcv_conv2d_bias_act_ws is a
                                const float* gate[N][Cout]; y = act((conv + bias) * gate[n, c] + residual) - the SE scale + identity
                                + activation of SEResNeXtUnit / SEResUnit (seresnext.py:57-66) in the epilogue of the unit's last
                                1x1 conv, possible because the squeeze is taken on that conv's INPUT (mean commutes with a 1x1 conv) */
  PCV_CONV_F32_SPLIT = 32,   /* fp32 tier only: evaluate the dense / grouped conv on the tensor cores as a 3-way bf16 split of
                                the fp32 activations and weights (24 significant bits per operand, exact products, fp32
                                accumulation: fp32-FMA accuracy at tcgen05 rate).  Needs pcv_conv_workspace_bytes() of
                                scratch handed to pcv_conv2d_bias_act_ws; weights must be packed with the same flag */
  PCV_CONV_POOL3S2 = 16      /* space-to-depth stem only: fuse the following MaxPool2d(3, stride 2, pad 1) (ResInitBlock,
                                resnet.py:255-263); y is the POOLED map [N, Ho/2, Wo/2, Cout].  Ask pcv_stem_s2d_pool_ok first */
};

/* One ConvBlock (conv.py:204-286): y = act(BN(conv2d(x)) [+ residual]).  Square kernels, symmetric padding. */
typedef struct pcv_conv_desc {
  int32_t N, H, W;        /* input batch and spatial size */
  int32_t Cin, Cout;      /* channels (Cin is the logical count; buffers may be pitched wider) */
  int32_t kh, kw;         /* kernel size */
  int32_t stride, pad, dil;
  int32_t groups;         /* 1 = dense, Cin = depthwise (conv.py:472), else grouped (resnext.py:44-55) */
  int32_t act;            /* pcv_act applied AFTER the optional residual add (resnet.py:226-228) */
  int32_t in_pitch;       /* channel pitch of x        (0 -> Cin)  */
  int32_t out_pitch;      /* channel pitch of y        (0 -> Cout) */
  int32_t res_pitch;      /* channel pitch of residual (0 -> Cout) */
  int32_t flags;          /* pcv_conv_flags */
  int32_t in_row_pitch;   /* elements between consecutive input rows (0 -> W * in_pitch); image pitch = H * that */
  float act_param;        /* PCV_ACT_LEAKY_RELU: the negative slope (nn.LeakyReLU.negative_slope); ignored otherwise */
} pcv_conv_desc;

/* ---- library ---------------------------------------------------------------------------------------------- */
PCV_API const char* pcv_last_error(void);
PCV_API int pcv_version(void);
/* sm major*10+minor of device `dev`, SM count, bytes of HBM; PCV_ERR_NO_DEVICE when there is no GPU. */
PCV_API int pcv_device_info(int dev, int* sm_arch, int* sm_count, size_t* hbm_bytes);
/* number of kernels this library has launched (eager or via plans) since load, for bench.py's gpu_launches */
PCV_API int64_t pcv_launch_count(void);

/* ---- weights: BN fold + repack (norm.py:34-50 folded into conv.py:250-259) ---------------- */
/* Output sizes in bytes for the packed weight and the fp32 bias of `d` in tier `dtype`. */
PCV_API int pcv_conv_packed_bytes(const pcv_conv_desc* d, int dtype, size_t* w_bytes, size_t* bias_bytes);
/* w: fp32 [Cout, Cin/groups, kh, kw] (state_dict layout); conv_bias: fp32 [Cout] or NULL; bn_*: fp32 [Cout] or all
 * NULL when the block has no BatchNorm.  Computes w' = w*g/sqrt(v+eps), b' = (bias-mean)*g/sqrt(v+eps)+beta. */
PCV_API int pcv_pack_conv_weights(const pcv_conv_desc* d, int dtype, const float* w, const float* conv_bias,
                          const float* bn_gamma, const float* bn_beta, const float* bn_mean, const float* bn_var,
                          float eps, void* w_packed, float* bias_out, pcv_stream stream);

/* ---- the hot path -------------------------------------------------------------------------------------------- */
/* ConvBlock.forward (conv.py:278-286) + the unit's residual add and final activation (resnet.py:221-229,
 * mobilenetv2.py:62-71) fused into one kernel.  residual may be NULL. */
PCV_API int pcv_conv2d_bias_act(pcv_plan* plan, const pcv_conv_desc* d, int dtype, const void* x, const void* w_packed,
                        const float* bias, const void* residual, void* y, pcv_stream stream);

/* 1 when `d` (with PCV_CONV_SE_GATE set) is served by the kernel that has the gated epilogue, else 0: the caller then records the
 * convolution and pcv_se_scale_add_act separately. */
PCV_API int pcv_conv_se_gate_ok(const pcv_conv_desc* d, int dtype);
/* The same op for descriptors that need scratch memory (PCV_CONV_F32_SPLIT: the split copy of x).  `workspace`: device
 * buffer of pcv_conv_workspace_bytes() bytes (0 = none needed), 16-byte aligned, private to this op while it runs. */
PCV_API int pcv_conv_workspace_bytes(const pcv_conv_desc* d, int dtype, size_t* bytes);
PCV_API int pcv_conv2d_bias_act_ws(pcv_plan* plan, const pcv_conv_desc* d, int dtype, const void* x, const void* w_packed,
                           const float* bias, const void* residual, void* y, void* workspace, pcv_stream stream);

/* Cross-layer fusion of a bottleneck's tail (ResBottleneck.forward resnet.py:136-140 + ResUnit.forward :221-229):
 *     y = act3( conv3_1x1( relu(conv2_3x3(x)) ) + residual )
 * in one kernel - the 64-channel intermediate stays on the SM (TMEM -> registers -> shared memory -> tensor core).
 * d2: the 3x3 ConvBlock (dense, stride 1, pad 1, Cin % 64 == 0, Cout == 64, ReLU); d3: the 1x1 ConvBlock (64 -> Cout,
 * Cout % 128 == 0, act = the unit's ReLU/ReLU6, residual required); both packed with pcv_pack_conv_weights as usual.
 * pcv_bottleneck_tail_fusable returns 1 when the pair is in the kernel's domain (16-bit tiers, W + 2 <= 64, even H),
 * else 0 and the caller records the two convolutions separately. */
PCV_API int pcv_bottleneck_tail_fusable(const pcv_conv_desc* d2, const pcv_conv_desc* d3, int dtype);
PCV_API int pcv_bottleneck_tail(pcv_plan* plan, const pcv_conv_desc* d2, const pcv_conv_desc* d3, int dtype, const void* x,
                        const void* w2_packed, const float* bias2, const void* w3_packed, const float* bias3,
                        const void* residual, void* y, pcv_stream stream);

/* DwsConvBlock.forward (conv.py:605-608) and the depthwise -> linear-pointwise tail of LinearBottleneck (mobilenetv2.py:52-71)
 * in ONE kernel: y = act_pw(W_pw * act_dw(dw3x3(x)) [+ residual]); the depthwise tensor - the widest of the block - stays in
 * shared memory as the A operand of the pointwise GEMM.  dw: the depthwise ConvBlock (3x3, stride 1 or 2, pad 1, act in
 * {none, ReLU, ReLU6}), packed with pcv_pack_conv_weights as usual; pw: the 1x1 ConvBlock (Cout <= 256, same activation
 * family, optional residual), likewise.  pcv_dw_pw_fusable returns 1 inside the kernel's domain (16-bit tiers, output map at
 * least 14 wide and 8 high), else 0 and the caller records the two convolutions separately. */
PCV_API int pcv_dw_pw_fusable(const pcv_conv_desc* dw, const pcv_conv_desc* pw, int dtype);
PCV_API int pcv_dw_pw_fused(pcv_plan* plan, const pcv_conv_desc* dw, const pcv_conv_desc* pw, int dtype, const void* x,
                    const void* w_dw_packed, const float* bias_dw, const void* w_pw_packed, const float* bias_pw,
                    const void* residual, void* y, pcv_stream stream);

/* The whole inverted-residual block - LinearBottleneck.forward (mobilenetv2.py:62-71): conv1 (1x1 expansion ConvBlock) -> conv2
 * (depthwise 3x3 ConvBlock) -> conv3 (linear 1x1 ConvBlock) [+ x]; the same triple in FBNetUnit / SPNASUnit / ProxylessBlock
 * (fbnet.py:77-87, spnasnet.py:72-82, proxylessnas.py:64-70) - in ONE kernel:
 *   y = act_pw(W_pw * act_dw(dw3x3(act_ex(W_ex * x))) [+ residual]).
 * Neither the expanded tensor nor the depthwise tensor exists in HBM: the expansion runs on the tensor cores over the input halo
 * of each output tile, its result is staged in shared memory for the depthwise stencil, whose result is the A operand of the
 * projection.  ex: the 1x1 stride-1 expansion (Cin <= 64, act in {none, ReLU, ReLU6}); dw / pw as for pcv_dw_pw_fused (Cout <=
 * 128 at stride 1, <= 64 at stride 2); all three packed with pcv_pack_conv_weights as usual.  pcv_exp_dw_pw_fusable returns 1
 * inside the kernel's domain, else 0 and the caller records the expansion and pcv_dw_pw_fused (or three convolutions). */
PCV_API int pcv_exp_dw_pw_fusable(const pcv_conv_desc* ex, const pcv_conv_desc* dw, const pcv_conv_desc* pw, int dtype);
PCV_API int pcv_exp_dw_pw_fused(pcv_plan* plan, const pcv_conv_desc* ex, const pcv_conv_desc* dw, const pcv_conv_desc* pw,
                                int dtype, const void* x, const void* w_ex_packed, const float* bias_ex,
                                const void* w_dw_packed, const float* bias_dw, const void* w_pw_packed, const float* bias_pw,
                                const void* residual, void* y, pcv_stream stream);

/* The end of a ResUnit with a projection shortcut - ResUnit.forward (resnet.py:221-229) when resize_identity:
 *   identity = identity_conv(x)        (1x1 ConvBlock, stride s, no activation)
 *   x = body(x); x = x + identity; x = activ(x)        with body ending in ResBottleneck.conv3 (linear 1x1 ConvBlock)
 * as ONE GEMM over K-concatenated operands:  y = act([W3 | Wid] * [y2 ; x[::s]] + (b3 + bid)).  The identity tensor is never
 * written: the kernel's TMA producer reads the k-blocks beyond d->Cin from the second activation `x2` through d2's (strided)
 * 1x1 window.  d: the stride-1 1x1 conv over `x` (its act is the unit's activation), d2: the 1x1 shortcut conv over `x2`
 * (stride 1 or 2, act none, same N / Cout / output grid).  w_cat_packed: for every output channel the row of d's packed
 * weights followed by the row of d2's (both from pcv_pack_conv_weights, i.e. each padded to a multiple of 64 input channels);
 * bias_sum = bias + bias2.  16-bit tiers; pcv_conv1x1_dual_ok returns 1 inside the kernel's domain (Cout > 128, dense). */
PCV_API int pcv_conv1x1_dual_ok(const pcv_conv_desc* d, const pcv_conv_desc* d2, int dtype);
PCV_API int pcv_conv1x1_dual(pcv_plan* plan, const pcv_conv_desc* d, const pcv_conv_desc* d2, int dtype, const void* x,
                             const void* x2, const void* w_cat_packed, const float* bias_sum, void* y, pcv_stream stream);
/* The same for an SE unit with a projection shortcut - SEResNeXtUnit.forward (seresnext.py:57-66), SEResUnit (seresnet.py:63-72):
 *   y = act((W3 * y2 + bias) * gate[n, c] + (Wid * x[::s] + bias2))
 * d carries PCV_CONV_SE_GATE; the gate (fp32 [N][Cout], as for pcv_conv2d_bias_act_ws) multiplies conv3's half of the sum only,
 * so the kernel keeps the two halves in two TMEM accumulators (128-wide tiles) and the biases stay separate.  Same weight layout
 * and domain query (pcv_conv1x1_dual_ok with the flag set on d). */
PCV_API int pcv_conv1x1_dual_se(pcv_plan* plan, const pcv_conv_desc* d, const pcv_conv_desc* d2, int dtype, const void* x,
                                const void* x2, const void* w_cat_packed, const float* bias, const float* bias2,
                                const float* gate, void* y, pcv_stream stream);

/* nn.ZeroPad2d((left, right, top, bottom)): the explicit asymmetric padding of a ConvBlock built with a 4-tuple `padding`
 * (conv.py:245-249,279-280) and of EfficientNet's tf_mode forwards (F.pad(x, calc_tf_padding(...)), efficientnet.py:27-55).
 * y is [N, H + top + bottom, W + left + right, C]; the convolution that follows runs with pad = 0. */
PCV_API int pcv_zero_pad2d(pcv_plan* plan, int dtype, int N, int H, int W, int C, const void* x, int in_pitch, int pad_left,
                   int pad_right, int pad_top, int pad_bottom, void* y, int out_pitch, pcv_stream stream);

/* nn.MaxPool2d(k, stride, pad), -inf padding, floor mode (resnet.py:255-258, senet.py:154-157). */
PCV_API int pcv_maxpool2d(pcv_plan* plan, int dtype, int N, int H, int W, int C, int k, int stride, int pad, const void* x,
                  int in_pitch, void* y, int out_pitch, pcv_stream stream);

/* nn.AdaptiveAvgPool2d(1) == nn.AvgPool2d(7) on a 7x7 map (resnet.py:316-318, att.py:72, deeplabv3.py:77):
 * pooled[n, c] = mean over HW; fp32 accumulate.  out_dtype selects the storage of `pooled`. */
PCV_API int pcv_global_avgpool(pcv_plan* plan, int dtype, int N, int HW, int C, const void* x, int in_pitch, void* pooled,
                       int out_dtype, pcv_stream stream);

/* nn.AdaptiveAvgPool2d(k) with k > 1 (PyramidPoolingBranch, pspnet.py:55-79): y[n, by, bx, c] = mean of x over rows
 * [floor(by*H/out_h), ceil((by+1)*H/out_h)) and the same rule in W; NHWC `dtype` in and out, y dense [N, out_h, out_w, C]. */
PCV_API int pcv_adaptive_avgpool(pcv_plan* plan, int dtype, int N, int H, int W, int C, const void* x, int in_pitch,
                         int out_h, int out_w, void* y, pcv_stream stream);

/* SEBlock.forward (att.py:94-105) in three steps: squeeze = pcv_global_avgpool (fp32 out);
 * excite: gate = out_act(W2 * mid_act(W1 * pooled + b1) + b2), W1 [Cmid, C], W2 [C, Cmid] fp32 (att.py:74-87);
 *         `gate` must hold N*(C+Cmid) floats: [N, C] gates followed by [N, Cmid] scratch for the hidden layer;
 * scale:  y = act(x * gate[n, c] + identity)   (seresnext.py:62-65; identity may be NULL, act may be NONE). */
PCV_API int pcv_se_excite(pcv_plan* plan, int N, int C, int Cmid, const float* pooled, const float* w1, const float* b1,
                  const float* w2, const float* b2, int mid_act, int out_act, float* gate, pcv_stream stream);
/* The excite with an input of a different width: gate = out_act(W2 * mid_act(W1 * pooled + b1) + b2) with pooled [N, Cin],
 * W1 [Cmid, Cin], W2 [C, Cmid]; `gate` holds N*(C+Cmid) floats.  Used when the squeeze is moved UPSTREAM of the unit's last
 * 1x1 ConvBlock by linearity (seresnext.py:60-62: mean_HW(conv3(y)) == W3' * mean_HW(y) + b3'): the caller pools conv3's
 * input (half the bytes in SE-ResNeXt) and folds W3', b3' into W1, b1 on the host. */
PCV_API int pcv_se_excite_ex(pcv_plan* plan, int N, int Cin, int Cmid, int C, const float* pooled, const float* w1,
                     const float* b1, const float* w2, const float* b2, int mid_act, int out_act, float* gate,
                     pcv_stream stream);
PCV_API int pcv_se_scale_add_act(pcv_plan* plan, int dtype, int N, int HW, int C, const void* x, const float* gate,
                         const void* identity, int act, void* y, pcv_stream stream);

/* y = act(a + b), elementwise over N*HW*C (residual adds that could not be fused into a conv epilogue). */
PCV_API int pcv_add_act(pcv_plan* plan, int dtype, size_t count, const void* a, const void* b, int act, void* y,
                pcv_stream stream);

/* y[p, c] = act(x[p, c] * scale[c] + shift[c]) over `pixels` NHWC pixels of C channels (C % 8 == 0), then, when `slope` is
 * given, negative results are multiplied by slope[c].  scale / shift / slope: fp32 [C] device vectors, each may be NULL.
 * Serves the stand-alone pieces of the reference that cannot ride on a producing conv's epilogue: the BN -> ReLU
 * pre-activation of PreConvBlock / PreResActivation (conv.py:717-731, preresnet.py:203-221; scale / shift = the folded
 * BatchNorm, act = ReLU), nn.PReLU (activ.py:84-98; slope = its weight, broadcast on the host when num_parameters == 1)
 * and LeakyReLU behind kernels without that epilogue (slope filled with negative_slope). */
PCV_API int pcv_channel_affine_act(pcv_plan* plan, int dtype, size_t pixels, int C, const void* x, int in_pitch,
                                   const float* scale, const float* shift, const float* slope, int act, void* y,
                                   int out_pitch, pcv_stream stream);

/* ---- network edges ------------------------------------------------------------------------------------------- */
/* Reference tensors are NCHW fp32.  Ingest pads channels with zeros up to c_pitch. */
PCV_API int pcv_nchw_f32_to_nhwc(pcv_plan* plan, int dtype, int N, int C, int H, int W, const float* x, void* y,
                         int c_pitch, pcv_stream stream);
/* The same edge for an image stored as fp32 / bf16 / fp16 / uint8 NCHW (pcv_image_type): y = float(x) * scale[c] + bias[c]
 * (scale_host / bias_host: HOST arrays of C floats read when the op is created, NULL = identity; uint8 pipelines fold
 * their 1/255, mean and std here).  pcv_nchw_f32_to_nhwc is the PCV_IMG_F32, identity case. */
PCV_API int pcv_nchw_to_nhwc_ex(pcv_plan* plan, int dtype, int img_type, int N, int C, int H, int W, const void* x,
                        const float* scale_host, const float* bias_host, void* y, int c_pitch, pcv_stream stream);
PCV_API int pcv_nhwc_to_nchw_f32(pcv_plan* plan, int dtype, int N, int C, int H, int W, const void* x, int c_pitch,
                         float* y, pcv_stream stream);
/* Space-to-depth stem.  A k x k stride-2 pad-(k/2) convolution on a <=4-channel image (ResInitBlock's 7x7,
 * resnet.py:250-254; the 3x3 stems of mobilenetv2.py:108-112 and senet.py:139-142) is re-expressed as a
 * (p+1) x 1 stride-1 convolution, p = k/2, over a zero-bordered space-to-depth tensor
 *     s2d[n, hb + ceil(p/2), wb + ceil(p/2), (dy*2+dx)*C + c] = x[n, c, 2*hb+dy, 2*wb+dx]      (16 channels/pixel)
 * read through an overlapping window view of (p+1) consecutive pixels = (p+1)*16 "channels" at a pixel pitch of 16,
 * so the implicit GEMM sees a dense K = (p+1)^2 * 16 instead of k*k zero-padded 64-channel taps.
 * pcv_stem_s2d_dims: rows/cols of the s2d tensor (incl. borders) and the equivalent conv's Cin / taps.
 * pcv_stem_s2d_ingest: NCHW fp32 image -> interior of the (pre-zeroed) s2d tensor, bf16.
 * pcv_stem_s2d_weights: fp32 [Cout, C, k, k] -> fp32 [Cout, (p+1)*16, p+1, 1] (then pcv_pack_conv_weights). */
PCV_API int pcv_stem_s2d_dims(int C, int H, int W, int k, int* rows, int* cols, int* cin_eq, int* taps_eq);
PCV_API int pcv_stem_s2d_ingest(pcv_plan* plan, int N, int C, int H, int W, int k, const float* x, void* s2d,
                                pcv_stream stream);
/* pcv_stem_s2d_ingest for either 16-bit tier (dtype = PCV_BF16 | PCV_F16) and any pcv_image_type, with the per-channel
 * affine of pcv_nchw_to_nhwc_ex. */
PCV_API int pcv_stem_s2d_ingest_ex(pcv_plan* plan, int dtype, int img_type, int N, int C, int H, int W, int k,
                                   const void* x, const float* scale_host, const float* bias_host, void* s2d,
                                   pcv_stream stream);
/* 1 when the stem conv (C-channel HxW image, k x k stride 2 pad k/2, Cout channels) can run with PCV_CONV_POOL3S2:
 * Cout == 64, even conv map with 64 <= W/2 and W/2 + k/2 <= 128 columns (one conv row per 128-row M-block), else 0 and the
 * caller records pcv_conv2d_bias_act + pcv_maxpool2d. */
PCV_API int pcv_stem_s2d_pool_ok(int C, int H, int W, int k, int Cout);
PCV_API int pcv_stem_s2d_weights(int Cout, int C, int k, const float* w, float* w_eq, pcv_stream stream);

/* F.interpolate(mode="bilinear", align_corners=True) (deeplabv3.py:53,86).  Output is NHWC `dtype` with
 * out_pitch, or NCHW fp32 when out_nchw_f32 != 0 (the tensor the reference returns). */
PCV_API int pcv_bilinear_upsample_ac(pcv_plan* plan, int dtype, int N, int Hin, int Win, int C, const void* x, int in_pitch,
                             int Hout, int Wout, void* y, int out_pitch, int out_nchw_f32, pcv_stream stream);

/* ---- multi-GPU exchange ---------------------------------------------------------------------------- */
/* Batch-sharded inference has exactly one exchange per step: an all-gather of every rank's [N/G, classes] logits.  The
 * reference has no multi-GPU code; this replaces the host-launched ncclAllGather a data-parallel wrapper would issue with
 * ONE device-initiated kernel over NVLink peer memory (push to every peer, flag, wait for every peer, drain), so a step
 * has no host-side collective and no second stream.  Each rank owns an exchange buffer of pcv_peer_buffer_bytes() bytes
 * (cudaMalloc, zeroed; exported as a 64-byte CUDA IPC handle) that every peer process maps with pcv_peer_buffer_open.
 * Every rank must call pcv_peer_allgather the same number of times, in the same order, for a given set of buffers. */
PCV_API int pcv_peer_buffer_bytes(int world, size_t bytes_per_rank, size_t* total);
PCV_API int pcv_peer_buffer_alloc(size_t bytes, void** ptr, void* ipc_handle_out_64B_host);
PCV_API int pcv_peer_buffer_open(const void* ipc_handle_64B_host, void** ptr);
PCV_API int pcv_peer_buffer_close(void* ptr);
PCV_API int pcv_peer_buffer_free(void* ptr);
/* out[r * bytes_per_rank ...] = rank r's `local` for every r (rank order == image order).  peer_bufs_host: HOST array of
 * `world` device pointers - entry r is rank r's exchange buffer as mapped in this process (entry `rank` = the own buffer). */
PCV_API int pcv_peer_allgather(pcv_plan* plan, const void* local, size_t bytes_per_rank, int rank, int world,
                               void* const* peer_bufs_host, void* out, pcv_stream stream);

/* ---- plans ---------------------------------------------------------------------------------------------------- */
PCV_API int pcv_plan_create(pcv_plan** plan);
PCV_API int pcv_plan_destroy(pcv_plan* plan);
PCV_API int pcv_plan_num_ops(const pcv_plan* plan);
/* kernels launched per pcv_plan_run */
PCV_API int pcv_plan_num_launches(const pcv_plan* plan);
/* Enqueue every recorded op on `stream` in order.  No host synchronisation. */
PCV_API int pcv_plan_run(pcv_plan* plan, pcv_stream stream);
/* Capture the plan into a CUDA graph (once) and replay it; falls back to an error, never to eager, on failure. */
PCV_API int pcv_plan_graph_launch(pcv_plan* plan, pcv_stream stream);
/* Per-op device time of one eager pass (CUDA events on `stream`, synchronises): fills ms[0..num_ops). */
PCV_API int pcv_plan_profile(pcv_plan* plan, pcv_stream stream, float* ms_host, int capacity);
/* Human-readable name of op i ("conv_tc 1x1 s1 256->64 bn=64", ...).  Pointer valid until plan destroy. */
PCV_API const char* pcv_plan_op_name(const pcv_plan* plan, int i);
/* Algorithmic FLOPs and HBM bytes of op i (SURVEY 8d formulas), for the roofline. */
PCV_API int pcv_plan_op_cost(const pcv_plan* plan, int i, double* flops, double* bytes);

#ifdef __cplusplus
}
#endif
#endif /* PCV_B200_H */
