/*
 * mcquic_b200 -- C ABI of the B200-native McQuic hot path (analysis/synthesis convolutions + multi-codebook VQ).
 *
 * Every pointer is a DEVICE pointer owned by the caller (PyTorch tensors in the host-side mirror,
 * mcquic_b200/); the library allocates nothing the caller can see, keeps no mutable global state and
 * launches on the stream it is given.  Return value: 0 = ok, >0 = cudaError_t, <0 = library error
 * (see mcq_error_string).  The Python wrapper raises RuntimeError on non-zero, like the reference
 * does for bad shapes (mcquic/modules/entropyCoder.py:79-93).
 *
 * There is no FFI on the reference side for this path (it is eager PyTorch over cuDNN/cuBLAS,
 * SURVEY.md section 2b); each entry point names the reference Python interface it replaces.
 *
 * Activation format ("split-fp16 planes"): an fp32 activation a is carried between convolutions as two
 * fp16 NHWC planes  hi = fp16(a),  lo = fp16((a - hi) * 2048)   =>   a ~= hi + lo / 2048  (22+ bits).
 * Convolutions run on tcgen05 tensor cores (kind::f16, fp32 accumulate in TMEM) either with 3 passes
 * (hi*hi + (hi*lo + lo*hi)/2048: fp32-grade, used by encode so that code indices match the fp32
 * reference) or 1 pass (hi*hi: TF32-grade, used by decode).  Tensors that are only ever consumed by an
 * epilogue (residual identities, GDN operands, VQ inputs, pixels) stay fp32.
 */
#ifndef MCQUIC_B200_H_
#define MCQUIC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mcq_stream_t; /* cudaStream_t */

/* epilogue modes */
enum {
  MCQ_EPI_LINEAR = 0, /* y = conv + bias (+ res1_scale*res1) (+ res2)        ResidualBlock tail, blocks.py:62-78   */
  MCQ_EPI_GATE   = 1, /* y = res1 + aux * sigmoid(conv + bias)               AttentionBlock tail, blocks.py:281-288 */
  MCQ_EPI_GDN    = 2, /* y = aux * rsqrt(conv + bias)   (conv over aux^2)    GenDivNorm, gdn.py:67-86              */
  MCQ_EPI_IGDN   = 3  /* y = aux *  sqrt(conv + bias)                        InvGenDivNorm, gdn.py:89-91           */
};
/* activation applied to y before it is written as a split-fp16 plane pair.
 * MCQ_ACT_SQUARE planes (the GDN / IGDN operand, gdn.py:74) hold y^2 * MCQ_SQUARE_SCALE: the power-of-two pre-scale keeps
 * squares of activations up to |y| < 2047 inside fp16's range (65504); the convolution that consumes such planes folds
 * 1 / MCQ_SQUARE_SCALE into its w_scale (exact). */
enum { MCQ_ACT_NONE = 0, MCQ_ACT_SILU = 1, MCQ_ACT_SQUARE = 2 };
#define MCQ_SQUARE_SCALE 0.015625f /* 2^-6 */
/* output addressing */
enum {
  MCQ_STORE_NHWC         = 0, /* [n, hout, wout, cout]                                                       */
  MCQ_STORE_SHUFFLE_NHWC = 1, /* PixelShuffle(2) folded into the store: GEMM column (2i+j)*C+c -> [n,2y+i,2x+j,c] (convs.py:244-255) */
  MCQ_STORE_SHUFFLE_NCHW = 2  /* last layer: GEMM column 4c+2i+j -> fp32 NCHW [n, c, 2y+i, 2x+j] (compressor.py:139) */
};
/* kernel implementation */
enum { MCQ_IMPL_TCGEN05 = 0, MCQ_IMPL_SIMT = 1 };

typedef struct mcq_conv_params {
  /* A operand: split-fp16 NHWC planes [n, hin, win, cin]; a_lo may be NULL when passes == 1 */
  const void* a_hi;
  const void* a_lo;
  int32_t n, hin, win, cin;
  /* B operand: packed weights [cout_pad, ksize*ksize*cin] fp16, K-major with K = (tap, cin); value = w * 2^w_exp */
  const void* w_hi;
  const void* w_lo;
  int32_t cout;     /* real number of GEMM columns            */
  int32_t cout_pad; /* rows of the packed weight matrix (multiple of the N tile) */
  int32_t ksize;    /* 1 or 3 (padding = ksize / 2, zeros: convs.py:98-100)      */
  int32_t stride;   /* 1 or 2                                                    */
  float w_scale;    /* 2^-w_exp, applied to the accumulator                      */
  const float* bias; /* [cout] indexed by GEMM column */
  int32_t mode;     /* MCQ_EPI_*  */
  int32_t store;    /* MCQ_STORE_* */
  const float* res1; /* fp32, addressed like the output */
  float res1_scale;
  const float* res2;
  const float* aux;
  float* out_f32;   /* optional fp32 output */
  void* out0_hi;    /* optional split-fp16 output #0 = act0(y) */
  void* out0_lo;    /* may be NULL: only the hi plane is written (1-pass consumers) */
  int32_t out0_act;
  void* out1_hi;    /* optional split-fp16 output #1 = act1(y) */
  void* out1_lo;
  int32_t out1_act;
  int32_t passes;   /* 1 or 3 */
  int32_t impl;     /* MCQ_IMPL_* */
  void* ev_start;   /* optional cudaEvent_t recorded on the stream immediately before / after the kernel launch */
  void* ev_stop;    /* (tight per-launch timing for bench.py's roofline leg); NULL = none */
  /* optional: GroupNorm statistics fused into this convolution's epilogue (the conv3x3 in front of the nn.GroupNorm
   * of ResidualBlock(denseNorm=True), mcquic/nn/blocks.py:196-199).  When gn_partials != NULL the kernel also writes
   * (sum y, sum y^2) of its fp32 output y per (row block of 32 output pixels, `unit` consecutive channels) as
   * float2 [n][rowblocks_per_image][cout / unit]; layout from mcq_conv_gn_layout(), consumer mcq_groupnorm_apply().
   * Every entry is written exactly once (no atomics, no clearing needed). */
  void* gn_partials;
  int32_t gn_groups; /* group count of the GroupNorm that follows */
  /* MCQ_STORE_SHUFFLE_NCHW only: when != NULL the pixels are written here as uint8 [n, c, 2y+i, 2x+j] through the
   * reference's DeTransform (mcquic/utils/vision.py:135-146: ((x + 1) / 2 * 255.999).clamp(0, 255).byte()) instead of as
   * fp32 to out_f32 (which may then be NULL): what demo.decompressImage returns (demo.py:124-134), 1 byte per sample
   * over PCIe instead of 4. */
  void* out_u8;
  /* optional DEVICE scalar multiplied into w_scale by the epilogue (bias should then be zero): the backward convolutions
   * run on gradients pre-scaled into fp16's range by a factor that lives on the device (no host sync), and remove it here */
  const float* dev_scale;
} mcq_conv_params;

/* Replaces nn.Conv2d 3x3 / 1x1 (+ the elementwise ops around it) as used by mcquic/nn/blocks.py:62-288,
 * mcquic/nn/convs.py:77-100,221-276 and mcquic/nn/gdn.py:67-91. */
/* Channel counts: cin % 4 == 0 always (fp16 rows are read 8 bytes at a time).  MCQ_IMPL_TCGEN05 takes cin % 64 == 0 for
 * every kernel shape and, for stride-1 convolutions, any cin % 8 == 0: the K loop then runs ceil(cin / 64) chunks and TMA
 * zero-fills the part of the last 64-channel box that lies beyond the tensor (the 8 / 32-channel nets of
 * ResidualBackwardQuantizer, mcquic/modules/quantizer.py:600-657); anything else returns MCQ_ERR_UNSUPPORTED and is
 * served by MCQ_IMPL_SIMT.  cout % 8 == 0 for plain NHWC stores (the host mirror zero-pads RGB outputs to 8 channels). */
int mcq_conv2d(const mcq_conv_params* p, mcq_stream_t stream);

/* `count` convolutions executed in order by ONE persistent launch (the low-resolution tail: 137 of the 165 convs per
 * direction of mcquic/modules/compressor.py:122-176 run on <= 16x16 maps and are launch-latency-bound one by one).
 * Semantics = calling mcq_conv2d(&params[i]) for i = 0..count-1 on `stream`, results bit-identical, provided that
 * (1) all layers share n and passes and use MCQ_IMPL_TCGEN05, (2) distinct buffers do not overlap, (3) every layer
 * is image-local (true of every conv: output image i depends on input image i only).  Dependencies between layers
 * are derived from the pointers; independent neighbours (the two AttentionBlock branches, blocks.py:246-288) run
 * without a barrier between them.  Returns MCQ_ERR_UNSUPPORTED (nothing launched) if the chain cannot be taken as
 * a whole -- the caller then issues the layers one by one.  count <= mcq_conv_chain_max_layers(). */
int mcq_conv_chain(const mcq_conv_params* params, int32_t count, mcq_stream_t stream);
int32_t mcq_conv_chain_max_layers(void);
/* development aid: when set (device int64[3072]), CTA 0 of every chain launch records clock64() samples of its TMA
 * issues [0,1024), MMA stage arrivals [1024,2048) and epilogue tile begin/end pairs [2048,3072).  NULL = off. */
void mcq_debug_timeline(void* device_i64_3072);

/* First layer: conv3x3 stride 2, 3 -> cout, on the NCHW image, with AlignedPadding's reflect pad folded
 * into the gather (mcquic/data/transforms.py:86-99, mcquic/modules/compressor.py:124).
 * x: [n, 3, h, w], fp32 in [-1, 1] (x_is_u8 = 0) or uint8 (x_is_u8 = 1: the reference's input transform
 * convert_image_dtype + (x - 0.5) * 2, mcquic/demo.py:110-118, is applied on the fly with the same fp32 operations);
 * pad_top/pad_left: reflect padding already split as the reference does; hp, wp: padded size.  Output grid is
 * hp/2 x wp/2, NHWC.
 * mcq_stem_conv: FP32 FFMA kernel (cross-check / fallback); w: [cout, 27] fp32 (cin, r, s order = nn.Conv2d layout).
 * mcq_stem_conv_tc: tcgen05 kernel, fp32-grade 3-pass split like every encode-side layer; w_lohi: fp16
 * [cout_pad, 64], row = [w_lo (27 values, 5 zeros) | w_hi (27 values, 5 zeros)] of w * 2^e, w_scale = 2^-e;
 * cout <= cout_pad <= 128, cout % 8 == 0, cout_pad % 16 == 0 (MCQ_ERR_UNSUPPORTED otherwise: use mcq_stem_conv). */
int mcq_stem_conv(const void* x, int32_t x_is_u8, int32_t n, int32_t h, int32_t w, int32_t pad_top, int32_t pad_left,
                  int32_t hp, int32_t wp, const float* wgt, const float* bias, int32_t cout, float* out_f32,
                  void* out_hi, void* out_lo, int32_t out_act, mcq_stream_t stream);
int mcq_stem_conv_tc(const void* x, int32_t x_is_u8, int32_t n, int32_t h, int32_t w, int32_t pad_top,
                     int32_t pad_left, int32_t hp, int32_t wp, const void* w_lohi, float w_scale, const float* bias,
                     int32_t cout, int32_t cout_pad, float* out_f32, void* out_hi, void* out_lo, int32_t out_act,
                     mcq_stream_t stream);

/* Replaces _multiCodebookQuantization._distance + .encode (mcquic/modules/quantizer.py:144-179):
 * code[n,m,h,w] = argmin_k (|x|^2 + |c_k|^2) - 2 x.c_k, first index on ties.
 * x: fp32 NHWC [n,h,w,m*d] (channel = m_idx*d + d_idx); codebook [m,k,d]; c2 [m,k] = sum_d codebook^2.
 * codes: int64 [n,m,h,w].  logits (optional): fp32 [n,m,h,w,k] = -(distance)/sqrt(k) * logit_scale[m]
 * (quantizer.py:181-183,204; logit_scale = max(temperature, 1e-6), may be NULL = 1).
 * hist (optional): int32 [m,k], incremented (not cleared) by the occurrence count of every code
 * (the counting half of mcquic/modules/entropyCoder.py:28-36 / validate/handlers.py:138-172). */
int mcq_vq_assign(const float* x, const float* codebook, const float* c2, int64_t* codes, float* logits,
                  const float* logit_scale, int32_t* hist, int32_t n, int32_t h, int32_t w, int32_t m, int32_t k,
                  int32_t d, mcq_stream_t stream);

/* Tensor-core variant of mcq_vq_assign for d % 64 == 0 and k % 32 == 0 (qp=1: d=128): the x.c_k products run on
 * tcgen05 (3-pass split-fp16, fp32-grade) and the distance + first-index argmin happens in the GEMM epilogue; the
 * [n,m,h,w,k] distance tensor is never written.  cb_hi/cb_lo: codebook [m,k,d] as split-fp16 planes of c * 2^e,
 * cb_scale = 2^-e.  workspace: >= mcq_vq_workspace_bytes(...) bytes, 256 B aligned, contents undefined afterwards.
 * Same outputs / reference lines as mcq_vq_assign (no logits). Returns MCQ_ERR_UNSUPPORTED for other shapes. */
int mcq_vq_assign_tc(const float* x, const void* cb_hi, const void* cb_lo, float cb_scale, const float* c2,
                     int64_t* codes, int32_t* hist, int32_t n, int32_t h, int32_t w, int32_t m, int32_t k, int32_t d,
                     void* workspace, int64_t workspace_bytes, mcq_stream_t stream);
int64_t mcq_vq_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t m, int32_t k, int32_t d);

/* Single-launch fused VQ for the multi-codebook shapes (d = 32 or 64, k % 128 == 0, h*w a multiple or a divisor of
 * 32; BASELINE configs[2]: M=6, K=2048, d=32): replaces _distance + encode AND the soft-logit half of forward
 * (mcquic/modules/quantizer.py:144-183,204) in one pass.  The fp32 latents are read once, split to fp16 hi/lo on
 * the fly and multiplied on tcgen05 (3 passes, fp32-grade) against the TMA-streamed codebook; the epilogue forms
 * (|x|^2 + |c_k|^2) - 2 x.c_k in the reference's order, keeps the first-index argmin, and (logits != NULL) writes
 * the soft logits with TMA stores -- the only HBM-sized stream.  cb_lohi: fp16 [m*k, 2d], row = [lo(d) | hi(d)] of
 * c * 2^e (cb_scale = 2^-e).  Other arguments as mcq_vq_assign.  MCQ_ERR_UNSUPPORTED for other shapes
 * (mcq_vq_fused_supported() == 0): use mcq_vq_assign. */
int mcq_vq_assign_fused(const float* x, const void* cb_lohi, float cb_scale, const float* c2, int64_t* codes,
                        float* logits, const float* logit_scale, int32_t* hist, int32_t n, int32_t h, int32_t w,
                        int32_t m, int32_t k, int32_t d, mcq_stream_t stream);
int mcq_vq_fused_supported(int32_t h, int32_t w, int32_t k, int32_t d);

/* Replaces _multiCodebookDeQuantization.decode (mcquic/modules/quantizer.py:249-259): gather codebook[m, code].
 * codes int64 [n,m,h,w] -> fp32 NHWC [n,h,w,m*d] and/or split-fp16 plane pairs (raw and/or act). Returns
 * MCQ_ERR_CODE_RANGE via *status (device int32, optional) if a code is outside [0,k). */
int mcq_vq_dequant(const int64_t* codes, const float* codebook, int32_t n, int32_t h, int32_t w, int32_t m, int32_t k,
                   int32_t d, float* out_f32, void* out0_hi, void* out0_lo, int32_t out0_act, void* out1_hi,
                   void* out1_lo, int32_t out1_act, int32_t* status, mcq_stream_t stream);

/* Code histogram, int32 [m,k] += count (validate/handlers.py:138-172; entropyCoder.py:28-36). */
int mcq_code_histogram(const int64_t* codes, int32_t n, int32_t m, int32_t hw, int32_t k, int32_t* hist,
                       mcq_stream_t stream);

/* Replaces nn.GroupNorm(groups, c) as used between the two convolutions of ResidualBlock(denseNorm=True)
 * (mcquic/nn/blocks.py:198; only the `Neon` tokenizer builds it, mcquic/modules/compressor.py:181-233):
 * per image and group, mean / biased variance over (h, w, c/groups), y = (x - mean) * rsqrt(var + eps) * gamma + beta.
 * x: fp32 NHWC [n,h,w,c] (the fp32 output of the producing convolution); gamma, beta: [c] (state_dict `weight`, `bias`).
 * Outputs (at least one): out_f32 fp32 NHWC and/or the split-fp16 planes of act(y) that the next convolution reads as
 * its A operand (out_lo may be NULL: 1-pass consumers).  One launch, one thread-block cluster per image; statistics
 * are combined in double in a fixed order (bit-reproducible).  c % 4 == 0, c <= 512, else MCQ_ERR_UNSUPPORTED. */
int mcq_groupnorm(const float* x, int32_t n, int32_t h, int32_t w, int32_t c, int32_t groups, const float* gamma,
                  const float* beta, float eps, float* out_f32, void* out_hi, void* out_lo, int32_t out_act,
                  mcq_stream_t stream);

/* GroupNorm with the statistics pass fused into the producing convolution (north_star: "GroupNorm ... fused into the
 * epilogue").  mcq_conv_gn_layout: can the convolution described by p (shape, kernel, stride, store, mode, impl,
 * gn_groups; pointers other than out_f32 are not inspected) emit the partials?  0 = yes (sizes returned),
 * MCQ_ERR_UNSUPPORTED = no -> run the convolution without gn_partials and use mcq_groupnorm.
 * mcq_groupnorm_apply: reduces the partials in a fixed order in double (bit-reproducible) into stats
 * (float2 [n][groups] = mean, rstd; caller-allocated) and normalises x in ONE streaming pass:
 * y = x * (rstd * gamma) + (beta - mean * rstd * gamma) -> fp32 and/or split-fp16 planes of act(y).
 * HBM traffic: x once in, outputs once out (mcq_groupnorm reads x twice). */
int mcq_conv_gn_layout(const mcq_conv_params* p, int32_t* rowblocks_per_image, int32_t* unit);
int mcq_groupnorm_apply(const float* x, const void* partials, int32_t rowblocks_per_image, int32_t unit, int32_t n,
                        int32_t h, int32_t w, int32_t c, int32_t groups, const float* gamma, const float* beta,
                        float eps, void* stats, float* out_f32, void* out_hi, void* out_lo, int32_t out_act,
                        mcq_stream_t stream);

/* out = x + alpha * y over `count` fp32 values (count % 4 == 0), written as fp32 and/or as the split-fp16 planes of
 * act(out).  Replaces the two element-wise steps of ResidualBackwardQuantizer that no convolution epilogue can absorb:
 * `residual = latent - currentLatent` (mcquic/modules/quantizer.py:686; the latent exists before currentLatent does)
 * and `quantized + formerLevel` (:701).  alpha = +-1 reproduces the reference's fp32 result bit for bit. */
int mcq_add_scaled(const float* x, const float* y, float alpha, int64_t count, float* out_f32, void* out_hi,
                   void* out_lo, int32_t out_act, mcq_stream_t stream);

/* fp32 [count] -> split-fp16 planes of act(x * s), s = *dev_scale (optional DEVICE scalar, NULL = 1): the operand planes
 * of the training-step convolutions (gradients are brought into fp16's range by a power of two computed on the device).
 * out_lo may be NULL (1-pass consumers). */
int mcq_split_planes(const float* x, int64_t count, int32_t act, void* out_hi, void* out_lo, const float* dev_scale,
                     mcq_stream_t stream);
int mcq_nchw_to_nhwc(const float* x, int32_t n, int32_t c, int32_t h, int32_t w, float* out_f32, void* out0_hi,
                     void* out0_lo, int32_t out0_act, void* out1_hi, void* out1_lo, int32_t out1_act,
                     mcq_stream_t stream);
int mcq_nhwc_to_nchw(const float* x, int32_t n, int32_t c, int32_t h, int32_t w, float* out, mcq_stream_t stream);

/* Introspection */
const char* mcq_error_string(int code);
/* ---- training step (SURVEY.md section 8f NEXT-3): weight gradient of a convolution, replaces what autograd computes for
 * nn.Conv2d.weight in the reference's training forward/backward (mcquic/modules/compressor.py:35-43 under
 * mcquic/train/trainer.py:273-283).  dW[co, ci, r, s] = scale * (*dev_scale) * sum_p dY[p, co] * X[p @ (r, s), ci] on
 * tcgen05 with MN-major operands (csrc/conv_wgrad.cuh): x_hi = the forward convolution's A operand plane
 * [n, hin, win, cin] fp16 NHWC, dy_hi = the plane of the output gradient [n, hin/stride, win/stride, cout] fp16 NHWC
 * (one fp16 pass, fp32 accumulation: TF32-grade, as the reference trains).  dw: fp32 [cout, cin, ksize, ksize]
 * (nn.Conv2d layout), overwritten or (accumulate != 0) added to.  cin % 8 == 0, cout % 8 == 0; stride 2 needs
 * cin % 64 == 0 (MCQ_ERR_UNSUPPORTED otherwise).  workspace: device scratch of mcq_conv_wgrad_workspace_bytes(p), 256 B
 * aligned.  The dgrad (input gradient) needs no entry point of its own: it is mcq_conv2d with flipped / transposed
 * weights (stride 2: the sub-pixel form with MCQ_STORE_SHUFFLE_NHWC), see mcquic_b200/autograd.py. */
typedef struct mcq_wgrad_params {
  const void* x_hi;
  int32_t n, hin, win, cin;
  const void* dy_hi;
  int32_t cout;
  int32_t ksize;  /* 1 or 3 */
  int32_t stride; /* 1 or 2 */
  float* dw;
  float scale;
  const float* dev_scale; /* optional device scalar */
  int32_t accumulate;
  void* workspace;
  int64_t workspace_bytes;
} mcq_wgrad_params;
int64_t mcq_conv_wgrad_workspace_bytes(const mcq_wgrad_params* p);
int mcq_conv_wgrad(const mcq_wgrad_params* p, mcq_stream_t stream);

int mcq_version(void);            /* ABI version */
/* Tuning / A-B knobs (kernel selection, grid caps, profiling aids -- the table at the top of csrc/mcq_api.cu).  They are
 * explicit process-wide settings with fixed defaults: nothing on the launch path reads the environment.
 * mcq_set_option returns MCQ_ERR_BAD_ARG for an unknown name; mcq_get_option returns INT32_MIN for one. */
int mcq_set_option(const char* name, int32_t value);
int32_t mcq_get_option(const char* name);
int mcq_device_error_flag(void);  /* last device-side watchdog code (0 = none); resets on read */
int mcq_kernel_launch_count(void); /* number of kernels this library has launched in this process */

#define MCQ_ERR_BAD_ARG (-1)
#define MCQ_ERR_UNSUPPORTED (-2)
#define MCQ_ERR_DRIVER (-3)
#define MCQ_ERR_CODE_RANGE (-4)
#define MCQ_ERR_WATCHDOG (-5)

#ifdef __cplusplus
}
#endif
#endif /* MCQUIC_B200_H_ */
