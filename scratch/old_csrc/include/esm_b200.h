/*
 * esm_b200.h -- C ABI of libesm_b200.so: the B200 (sm_100a) implementation of ESMStereo's
 * feature-to-disparity hot path.
 *
 * The reference (rahul-rwat/ESMStereo) has no native operator API: its operator seams are Python
 * callables in models/submodule.py, models/shufflemixer.py, models/ESMStereo.py and
 * models/ESMStereo_confidence.py.  Each entry point below replaces one of those seams and cites it
 * (paths relative to the reference root).  Conventions:
 *   - all pointers are DEVICE pointers to fp32 (unless stated), caller-owned, no allocation inside;
 *   - tensors are NCHW / NCDHW; where strides are not passed they are contiguous;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant per stream,
 *     and legal inside CUDA-graph capture;
 *   - return value: 0 = ok, negative = error (see esm_last_error()).
 * There is no CPU fallback: without a CUDA device every compute entry point returns ESM_ERR_CUDA.
 */
#ifndef ESM_B200_H
#define ESM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ESM_OK 0
#define ESM_ERR_ARG (-1)   /* invalid argument / unsupported shape */
#define ESM_ERR_CUDA (-2)  /* CUDA runtime error (launch failure, no device) */

/* activation codes for esm_conv_t.act / act2 */
#define ESM_ACT_NONE 0
#define ESM_ACT_GELU 1     /* exact erf GELU, submodule.py:37 */
#define ESM_ACT_RELU 2
#define ESM_ACT_SILU 3
#define ESM_ACT_SIGMOID 4
#define ESM_ACT_2SIGMOID 5 /* 2*sigmoid(x): LAFNet scale head, ESMStereo_confidence.py:691 */
#define ESM_ACT_RELU6 6

/* esm_conv_t.src_mode */
#define ESM_SRC_TENSORS 0  /* up to 3 channel-concatenated strided tensors (replaces torch.cat + crop) */
#define ESM_SRC_GWC 1      /* input voxels are the group-wise correlation of src[0]=left, src[1]=right
                              feature maps, computed on the fly (never written to HBM);
                              fuses submodule.py:151-161 into the consumer conv (ESMStereo.py:708-713) */

const char* esm_last_error(void);
int esm_version(void);
/* Device properties the host side sizes grids with; returns ESM_ERR_CUDA when no device. */
int esm_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* One source of a (possibly concatenated) conv input.  W stride is 1. */
typedef struct {
  const float* ptr; /* element [b=0,c=0,d=0,h=0,w=0] of the (possibly cropped) view */
  int C;            /* channels contributed by this source */
  long long sB, sC, sD, sH; /* strides in elements */
} esm_src_t;

/*
 * Generic fused convolution: conv / transposed conv (2D = 3D with D=1) + per-channel affine
 * (folded eval BatchNorm and/or bias) + activation + optional broadcast multiply + optional
 * residual add + optional second activation, with optional PixelShuffle store.
 * Replaces BasicConv.forward (submodule.py:32-38) and every bare nn.Conv2d / ConvTranspose on the
 * path (ESMStereo.py:129-182 aggregation, :185-239 up_refinement, :242-509 upsample4/8/16,
 * shufflemixer.py:123-127, ESMStereo_confidence.py:511-744).
 */
typedef struct {
  esm_src_t src[3];
  int nsrc;
  int src_mode;              /* ESM_SRC_* */
  int gwc_groups;            /* ESM_SRC_GWC: number of groups G (input channels of the conv);
                                src[0].C == src[1].C == feature channels */
  const float* in_mul;       /* optional [B,Cin,1,Hin,Win] multiplier broadcast over D (att, ESMStereo.py:711) */
  int B, Cin, Din, Hin, Win; /* logical input extent */
  int Cout, Dout, Hout, Wout;
  int kd, kh, kw;            /* kernel extent; transposed: must be 4 (D: 4 or 1) */
  int stride;                /* 1 or 2 (same in every spatial dim; D stride is 1 when kd==1) */
  int pd, ph, pw;            /* padding */
  int transposed;            /* 1: ConvTranspose k4 s2 p1 (sub-pixel phase decomposition) */
  const float* weight;       /* packed by esm_pack_conv_weight_f32 */
  const float* scale;        /* [Cout] y = acc*scale + shift  (NULL -> 1) */
  const float* shift;        /* [Cout]                        (NULL -> 0) */
  int act;                   /* ESM_ACT_* applied after the affine */
  const float* out_mul;      /* optional [B,Cout,1,Hout,Wout] multiplier after act (ESMStereo.py:703) */
  const float* residual;     /* optional tensor with the output's strides, added after act/out_mul */
  int act2;                  /* ESM_ACT_* applied after the residual add */
  float out_scale;           /* final multiply (1.0f = none) */
  int pixel_shuffle;         /* 0, or r: out[b, co/(r*r), h*r+(co/r)%r, w*r+co%r] (2D only) */
  float* out;
  long long oB, oC, oD, oH;  /* output strides in elements (W stride 1) */
  int engine;                /* 0: whichever engine wins the on-device timing (FP32 pipe or tcgen05 split-TF32);
                                1: FP32 pipe only -- for layers whose consumer amplifies rounding error (the cost
                                path of the confidence head: softmax(-100 * cost / |cost|), ESMStereo_confidence.py:575) */
} esm_conv_t;

/* Elements needed for the packed form of a conv weight (fp32 count). */
long long esm_packed_weight_elems(int Cout, int Cin, int kd, int kh, int kw, int transposed);
/*
 * Repack a torch-layout weight ([Cout,Cin,kd,kh,kw]; transposed: [Cin,Cout,kd,kh,kw]) into the
 * kernel layout ([phase][tap][Cin_pad][Cout_pad], zero padded).  Device to device.
 */
int esm_pack_conv_weight_f32(const float* w, float* packed, int Cout, int Cin, int kd, int kh, int kw,
                             int transposed, void* stream);
/*
 * Fold eval-mode BatchNorm (+ optional conv bias) into scale/shift:
 *   scale = gamma / sqrt(var + eps); shift = beta - mean*scale + bias*scale.
 * Any of gamma/beta/mean/var may be NULL together (no BN): scale = 1, shift = bias (or 0).
 */
int esm_fold_bn_f32(const float* gamma, const float* beta, const float* mean, const float* var,
                    const float* bias, float eps, int C, float* scale, float* shift, void* stream);
int esm_conv_f32(const esm_conv_t* desc, void* stream);
/*
 * Engine plans.  esm_conv_f32 chooses, per layer shape, one of its engines (FP32 pipe tiling, resident / streamed
 * tcgen05, pointwise) -- by timing the candidates on the device the first time a shape is seen, unless a plan for
 * that shape has been imported: then the same engine and tiling are used with no timing and no synchronisation, so
 * every process computes a layer with the same rounding.  esm_conv_plans_export writes the plans this process tuned
 * (text, one line per shape; returns the bytes needed incl. the terminator), esm_conv_plans_import reads such text
 * (returns the number of plans read).  ESM_AUTOTUNE=1 ignores imported plans and re-times; ESM_AUTOTUNE=0 never times
 * (imported plan, else the analytic model's FP32-pipe choice).  esm_conv_tuned_calls counts the calls that timed.
 */
long long esm_conv_plans_export(char* buf, long long cap);
int esm_conv_plans_import(const char* text);
long long esm_conv_tuned_calls(void);

/* Number of esm_conv_f32 calls so far that ran on the tcgen05 tensor-core path (k3 s1 p1 layers, when
 * it wins the on-device timing or ESM_TC_FORCE is set; ESM_TC=0 disables it, ESM_TC=1 selects the
 * single-pass TF32 fast mode instead of the fp32-grade split).  Diagnostics / tests. */
long long esm_tc_conv_launches(void);
/* Same for the streamed-weight tcgen05 engine (taps in K, weights through the operand ring: wide / strided /
 * transposed layers; ESM_TC_FORCE=2 forces it wherever eligible, ESM_TCG_OFF=1 disables it). */
long long esm_tcg_conv_launches(void);
/* Same for the pointwise (k1) streaming kernel (conv_pw.cu: true fp32, HBM-bound; ESM_TC_FORCE=3 forces it wherever
 * eligible, ESM_PW_OFF=1 disables it). */
long long esm_pw_conv_launches(void);

/* build_gwc_volume (submodule.py:151-161): L,R [B,C,H,W] -> V [B,G,D,H,W]; writes the zero
 * triangle itself (no memset). */
int esm_gwc_volume_f32(const float* L, const float* R, float* V, int B, int C, int H, int W, int D, int G,
                       void* stream);
/* build_norm_correlation_volume (submodule.py:187-200): -> V [B,1,D,H,W].
 * `ws` is scratch of 2*B*C*H*W floats (normalised copies of L and R). */
int esm_norm_corr_volume_f32(const float* L, const float* R, float* V, float* ws, int B, int C, int H, int W,
                             int D, void* stream);

/* build_concat_volume (submodule.py:129-140): -> V [B,2C,D,H,W]; V[:, :C, d] = L (whole row), V[:, C:, d, :, x] =
 * R[.., x-d] for x >= d else 0.  Not used by any model configuration (kept for the reference's operator seam). */
int esm_concat_volume_f32(const float* L, const float* R, float* V, int B, int C, int H, int W, int D, void* stream);
/* build_substract_volume (submodule.py:104-126): V[b,g,d,y,x] = sum_{c in g} (L - R(x-d))^2 for x >= d else 0. */
int esm_substract_volume_f32(const float* L, const float* R, float* V, int B, int C, int H, int W, int D, int G,
                             void* stream);

/* build_gwc_volume_norm (submodule.py:163-184): group-wise correlation of the features L2-normalised per group and
 * pixel (norm + 1e-5): V[b,g,d,y,x] = mean_{c in g} L^[c,y,x] * R^[c,y,x-d] for x >= d else 0.  Unused by the models. */
int esm_gwc_volume_norm_f32(const float* L, const float* R, float* V, int B, int C, int H, int W, int D, int G,
                            void* stream);

/* regression_topk(cost, arange(D), k=2) (submodule.py:218-225, ESMStereo.py:719-721):
 * cost [B,D,H,W] -> pred [B,1,H,W]; idx (optional, int32 [B,2,H,W]) receives the top-2 indices
 * (ties: lower index first). */
int esm_regression_top2_f32(const float* cost, float* pred, int* idx, int B, int D, int H, int W, void* stream);
/* The same regression reading the cost volume in the sub-pixel form the hourglass's `conv1_up` (ConvTranspose3d k4 s2 p1
 * to one channel, ESMStereo.py:150,182) is computed in: y8 [B, 8 = (pd,ph,pw), D2, H2, W2] with the given strides,
 * cost[b,d,y,x] = y8[b, (d&1)*4 + (y&1)*2 + (x&1), d/2, y/2, x/2] -- the PixelShuffle copy of the volume is skipped.
 * pred [B,1,2H2,2W2]; idx optional int32 [B,2,2H2,2W2]. */
int esm_regression_top2_subpixel_f32(const float* y8, long long sB, long long sC, long long sD, long long sH, float* pred, int* idx,
                                     int B, int D2, int H2, int W2, void* stream);
/* ... and the PixelShuffle itself, for callers that need the volume (confidence head, tests): -> [B,1,2D2,2H2,2W2]. */
int esm_pixel_shuffle3d_f32(const float* in, long long sB, long long sC, long long sD, long long sH, float* out, int B, int D2, int H2,
                            int W2, void* stream);
/* disparity_regression (submodule.py:211-216): sum_d cost[d]*d, no softmax -> [B,1,H,W]. */
int esm_disparity_regression_f32(const float* cost, float* pred, int B, int D, int H, int W, void* stream);

/* out = (bilinear_up(prev, factor, align_corners=False) + residual) * out_scale
 * (ESMStereo.py:307,316,745).  prev [B,1,h,w], residual/out [B,1,h*f,w*f]. */
int esm_bilinear_add_f32(const float* prev, const float* residual, float* out, int B, int h, int w, int factor,
                         float out_scale, void* stream);

/* Image pre-processing (test_kitti.py:93-106, datasets/kitti_dataset.py:145-160): uint8 RGB [B,h,w,3] ->
 * float32 [B,3,Hp,Wp]: ToTensor (/255) + Normalize((x - mean) / std, true divisions) placed at (pad_top, pad_left);
 * the padding is black pixels normalised like any other (fill_normalised = 1: test_kitti.py's PIL crop with a
 * negative origin, pad_top = Hp - h, pad_left = Wp - w) or zeros after normalisation (fill_normalised = 0:
 * kitti_dataset.py's np.pad on the top / right, pad_left = 0).  mean3 / std3 are HOST pointers to 3 floats. */
int esm_preprocess_u8_f32(const unsigned char* rgb_hwc, float* out_chw, int B, int h, int w, int Hp, int Wp,
                          int pad_top, int pad_left, int fill_normalised, const float* mean3, const float* std3,
                          void* stream);
/* Disparity post-processing (test_kitti.py:114,127; save_disp.py:83-88): crop [top:top+h, left:left+w] of the padded
 * [B,Hp,Wp] disparity, round(d * scale) half-to-even -> uint16 [B,h,w] (saturating). */
int esm_postprocess_disp_u16(const float* disp, unsigned short* out, int B, int Hp, int Wp, int top, int left,
                             int h, int w, float scale, void* stream);

/* Device-to-device copy of n floats on `stream` (torch.cat of the left / right images, ESMStereo.py:640-641 batched). */
int esm_copy_f32(float* dst, const float* src, long long n, void* stream);
/* Post-processing of the reference's ROS publisher (kitti_publisher/src/kitti_publisher_cuda_node.cpp:385-404): crop the
 * padded [Hp,Wp] disparity to [h,w] at the origin, 5x5 median (cv::medianBlur semantics for CV_32F: exact median,
 * replicated border), zero unless 0 < d < max_disp, then saturate(round-half-even(d * scale)) -> uint16 [h,w]. */
int esm_disparity_publish_u16(const float* disp, unsigned short* out, int Hp, int Wp, int h, int w, float max_disp, float scale,
                              void* stream);

/*
 * ShuffleMixer SMLayer halves (shufflemixer.py:97-112), C in {8,16}:
 *   pointwise:  y = shuffle8(cat(MLP(LN(x)[:C/2]), LN(x)[C/2:])) + x
 *   spatial:    t = depthwise_kxk(x) + bias; y = shuffle8(cat(MLP(LN(t)[:C/2]), LN(t)[C/2:])) + t
 * MLP = 1x1 (C/2 -> hidden) + SiLU + 1x1 (hidden -> C/2), with bias; LN is bias-free over channels.
 * fc0_w [hidden, C/2], fc0_b [hidden], fc2_w [C/2, hidden], fc2_b [C/2]  (torch layouts).
 * `extra_residual` (optional) is added to the output (FMBlock's `net(x) + x`, shufflemixer.py:130).
 */
typedef struct {
  const float* ln_w;                    /* [C] */
  const float* fc0_w; const float* fc0_b;
  const float* fc2_w; const float* fc2_b;
  int hidden;
} esm_mixer_mlp_t;
int esm_sm_pointwise_f32(const float* x, float* y, int B, int C, int H, int W, const esm_mixer_mlp_t* mlp,
                         const float* extra_residual, void* stream);
int esm_sm_spatial_f32(const float* x, float* y, int B, int C, int H, int W, const float* dw_w /*[C,1,k,k]*/,
                       const float* dw_b /*[C]*/, int k, const esm_mixer_mlp_t* mlp, const float* extra_residual,
                       void* stream);

/* A whole SMLayer (shufflemixer.py:97-112) in one launch, k = 7: pointwise half with mlp1 on the tile + halo in shared
 * memory, depthwise 7x7 + bias, pointwise half with mlp2, optional extra residual.  Same arithmetic as
 * esm_sm_pointwise_f32 followed by esm_sm_spatial_f32. */
int esm_sm_layer_f32(const float* x, float* y, int B, int C, int H, int W, const esm_mixer_mlp_t* mlp1, const float* dw_w,
                     const float* dw_b, int k, const esm_mixer_mlp_t* mlp2, const float* extra_residual, void* stream);

/* LAFNet cost tower front end (ESMStereo_confidence.py:645-653): per pixel
 * softmax(-cost/||cost||_2 * 100) over D, then the 7 largest probabilities, descending.
 * cost [B,D,H,W] (D <= 64) -> out [B,7,H,W]. */
int esm_laf_cost_top7_f32(const float* cost, float* out, int B, int D, int H, int W, void* stream);
/* LAFNet 3-way attention (ESMStereo_confidence.py:676-686): softmax over the three 1-channel
 * attention maps, towers scaled and concatenated: out [B,3C,H,W]. */
int esm_laf_attention_f32(const float* cost_x, const float* disp_x, const float* imag_x, const float* att_c,
                          const float* att_d, const float* att_i, float* out, int B, int C, int H, int W,
                          void* stream);
/* LAFNet scale-adaptive sampling + embed_conv2 (k3, stride 3) + affine + ReLU
 * (ESMStereo_confidence.py:693-719), fused: the 3h x 3w `grid_sample` image is never materialised.
 * feat [B,C,H,W], scale [B,1,H,W], lin_x [W], lin_y [H] (np.linspace(-1,1,n) as fp32),
 * weight [C,C,3,3] (torch layout), bn scale/shift [C] -> out [B,C,H,W]. */
int esm_laf_sample_embed_f32(const float* feat, const float* scale, const float* lin_x, const float* lin_y,
                             const float* weight, const float* bn_scale, const float* bn_shift, float* out, int B,
                             int C, int H, int W, void* stream);
/* conf_upsample convex x4 upsampling (ESMStereo_confidence.py:536-543): logits = ConvTranspose2d(C->9,
 * k4, s4)(feat) + bias; out[Y,X] = sum_k softmax(logits)[k] * conf[Y/4 + k/3 - 1, X/4 + k%3 - 1].
 * feat [B,C,h,w], conf [B,1,h,w], weight [C,9,4,4] (torch layout), bias [9] -> out [B,1,4h,4w]. */
int esm_conf_convex_up4_f32(const float* feat, const float* conf, const float* weight, const float* bias,
                            float* out, int B, int C, int h, int w, void* stream);
/*
 * Pieces of the timm backbones behind `Feature` (ESMStereo.py:40-77: timm.create_model('efficientnet_b2' |
 * 'mobilenetv2_100', features_only=True)) that are not plain convolutions: depthwise k x k conv (k = 3, 5, 7; stride 1
 * or 2; padding k/2) + per-channel affine (folded BN) + activation; global average pool [B,C,H*W] -> [B,C]; in-place
 * multiply by a per-(image, channel) gate (squeeze-and-excitation).  Their 1x1 convolutions run on esm_conv_f32.
 */
int esm_dwconv2d_f32(const float* x, const float* w /*[C,1,k,k]*/, const float* scale, const float* shift, int act, float* y, int B,
                     int C, int H, int W, int k, int stride, void* stream);
int esm_global_avgpool_f32(const float* x, float* out, int B, int C, int HW, void* stream);
int esm_scale_channels_f32(float* x, const float* gate, int B, int C, int HW, void* stream);

/* Measured dense TF32 tensor-core rate of the current device (TFLOP/s): every SM issues `iters` back-to-back
 * M128 x N256 x K8 tcgen05.mma.  Synchronises; diagnostics / bench.py's tensor-roofline denominator. */
int esm_umma_tf32_peak(int iters, float* tflops, void* stream);
/* Synchronous device -> host copy of raw bytes (engine export: esmstereo_b200/engine.py). */
int esm_download(void* dst_host, const void* src_device, long long nbytes);
/* Fill `n` floats with `value`. */
int esm_fill_f32(float* p, long long n, float value, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * "PF" activations (padded-flat, pre-split) and the TMA-fed tcgen05 convolution engine over them (conv_tcf.cu).
 *
 * A PF tensor stores a logical [B, C, D, H, W] fp32 activation as
 *     [B][hi | lo][Cq = 2*ceil(C/8) channel quads][Dp * Hp * P positions][4 channels]
 * with the convolution's zero border STORED (Hp >= H + 2, P >= W + 2, Dp >= D + 2 for 3D, Dp == 1 for 2D), the
 * data occupying the "valid box" [d0,d1) x [y0,y1) x [x0,x1) of the padded lattice and zeros everywhere else, and
 * every value pre-split for the fp32-grade TF32 scheme: hi = x rounded to TF32, lo = x - hi.  `data` points at
 * (b=0, hi, quad 0, position 0); the allocation must extend esm_pf_guard_elems() floats on BOTH sides of the
 * esm_pf_elems() floats of the tensor (the engine's windows overhang; what they read there only reaches border
 * outputs, which are stored as zeros).  Intermediate format between BasicConv layers (submodule.py:12-38): it
 * replaces the NCHW tensors the reference passes from conv to conv inside `aggregation`, `up_refinement`, the
 * disparity MLPs / spx blocks of `upsample4/8/16` (ESMStereo.py:129-318) and the 2D feature side (:79-125).
 */
typedef struct {
  float* data;
  int B, C;
  int Dp, Hp, P;
  int d0, d1, y0, y1, x0, x1;
} esm_pf_t;
long long esm_pf_elems(int B, int C, int Dp, int Hp, int P);
long long esm_pf_guard_elems(int Dp, int Hp, int P);
/* NCHW (strides in elements, W stride 1; sD ignored for 2D) -> PF, zero border included; and back (hi + lo). */
int esm_pf_from_nchw_f32(const float* x, long long sB, long long sC, long long sD, long long sH, const esm_pf_t* out, void* stream);
int esm_pf_to_nchw_f32(const esm_pf_t* in, float* out, long long sB, long long sC, long long sD, long long sH, void* stream);

/* Weights for esm_conv_pf_f32: torch layout in ([Cout,Cin,k..]; transposed [Cin,Cout,k..]), split into TF32 hi / lo
 * UMMA slabs out.  srcC[nsrc] = channels of each concatenated source (each is padded to a multiple of 8 in PF). */
long long esm_packed_weight_pf_elems(int Cout, int nsrc, const int* srcC, int kd, int kh, int kw, int transposed);
int esm_pack_conv_weight_pf_f32(const float* w, float* packed, int Cout, int nsrc, const int* srcC, int kd, int kh, int kw,
                                int transposed, void* stream);
/*
 * Fused convolution over PF sources: k1 / k3 stride 1 (pad k/2), k3 stride 2 (pad 1; computed at every position,
 * stored at the even ones), ConvTranspose k4 s2 p1 (2^nd sub-pixel phases) -- 2D or 3D, up to 3 channel-concatenated
 * sources of one geometry -- + per-channel affine + activation + residual + second activation, written as PF
 * (same geometry: every position, zeros outside the compute box; other geometry: the mapped valid positions only,
 * into a buffer whose other positions are already zero) and / or as strided NCHW (optionally PixelShuffle(2)).
 * Same seams as esm_conv_f32; operands reach the tensor core by TMA bulk copies only.
 */
typedef struct {
  esm_pf_t src[3];
  int nsrc;
  int Cout, kd, kh, kw, stride, transposed;
  int d0, d1, y0, y1, x0, x1;  /* compute box in the sources' padded coordinates (normally their valid box) */
  int oD, oH, oW;              /* logical output extent (crop-to-skip for transposed layers) */
  const float* weight;         /* esm_pack_conv_weight_pf_f32 */
  const float* scale;
  const float* shift;
  int act, act2;
  float out_scale;
  int pixel_shuffle;           /* 0 or 2 (NCHW output only) */
  esm_pf_t out_pf;             /* data == NULL: no PF output */
  const float* res_pf;         /* optional residual in the layout of out_pf */
  float* out;                  /* optional NCHW output */
  long long oB, oC, oDs, oHs;
  const float* residual;       /* optional NCHW residual with the output's strides */
} esm_conv_pf_t;
int esm_conv_pf_f32(const esm_conv_pf_t* desc, void* stream);
long long esm_tcf_conv_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* ESM_B200_H */
