/* xv2.h -- C ABI of libxv2 (hand-written sm_100a CUDA for the xView2 segmentation hot path).
 *
 * The reference (michal2409/xView2) is pure Python: it has no FFI of its own.  Every entry point below replaces a
 * library call the reference reaches through PyTorch (cuDNN / ATen / MONAI / numpy); the call site it replaces is
 * cited per function as <reference file>:<line>.  The Python mirror of the reference's module API
 * (xview2_b200/model/*.py) binds these symbols with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name says host.
 *  - activations are NHWC ("channels-last"), contiguous unless a pixel stride `ld*` (in elements) is given.
 *  - conv weights are [K][R][S][C/groups] (the physical order of a channels-last OIHW tensor);
 *    transposed-conv weights are [Cin][R][S][Cout] (channels-last physical order of torch's (Cin,Cout,R,S)).
 *  - dtype: XV2_F32 or XV2_BF16 for activations/packed weights; accumulation is always fp32;
 *    master weights, gradients of weights, BN statistics and losses are fp32 (statistic partials fp64).
 *  - `stream` is a cudaStream_t passed as void*.  No call allocates, synchronises or touches the host.
 *  - return 0 on success, a negative XV2_E* code on failure; xv2_last_error() gives the message (thread-local).
 */
#ifndef XV2_H
#define XV2_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XV2_F32 0
#define XV2_BF16 1

#define XV2_OK 0
#define XV2_EINVAL (-1)   /* bad argument / unsupported shape for this entry point */
#define XV2_ECUDA (-2)    /* CUDA runtime / driver error */
#define XV2_EUNSUPPORTED (-3) /* shape not eligible for the tensor-core path (caller uses the SIMT entry) */

#define XV2_ACT_NONE 0
#define XV2_ACT_RELU 1
#define XV2_ACT_LRELU 2 /* negative slope 0.01, layers.py:94 */

const char* xv2_last_error(void);
int xv2_version(void);
/* Loads the driver entry point for tensor-map encoding and checks the device is sm_100. */
int xv2_init(int device);

/* ------------------------------------------------------------------------------------------------------------
 * Convolution geometry shared by the SIMT and tensor-core entry points.
 *   out[n,oh,ow,k] = sum_{r,s,c} src[n, ih, iw, g*Cg + c] * w[k][r][s][c]
 *   with  num_h = oh*stride - pad + r*dil ;  taken only if num_h % ups == 0 ; ih = num_h / ups ; 0 <= ih < h
 * `ups` = 1 is an ordinary convolution; ups > 1 expresses transposed convolution / strided dgrad as a gather.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct xv2_conv_geom {
  int32_t n, h, w, c;       /* source tensor NHWC, c = all source channels */
  int32_t oh, ow, k;        /* output tensor, k = all output channels */
  int32_t r, s;             /* taps */
  int32_t stride, pad, dil; /* square */
  int32_t ups;              /* source up-sampling factor (1 = none) */
  int32_t groups;
  int32_t dtype;            /* XV2_F32 | XV2_BF16 : source, weights */
  int32_t out_dtype;        /* XV2_F32 | XV2_BF16 : output */
} xv2_conv_geom;

/* SIMT (CUDA-core, fp32 accumulate) gather convolution: any stride/dilation/groups/channel count.
 * Replaces nn.Conv2d / nn.ConvTranspose2d forward and data-gradient: layers.py:83,92,71,180; unet.py:52 (encoder).
 * bias: fp32 [k] or NULL. */
int xv2_conv_gather_simt(const xv2_conv_geom* g, const void* src, const void* w, const float* bias, void* out,
                         void* stream);

/* SIMT weight gradient: dw[k][r][s][c] (fp32, ACCUMULATED into dw with atomics: caller zero-fills) =
 * sum_{n,oh,ow} dout[n,oh,ow,k] * src[n,ih,iw,g*Cg+c]; same index rule as above.  dout has geom.out dims, dtype `dtype`. */
int xv2_conv_wgrad_simt(const xv2_conv_geom* g, const void* src, const void* dout, float* dw, void* stream);

/* ResNeSt deep-stem first convolution (unet.py:52 -> conv1[0]): 3x3 stride 2 pad 1, 3 input channels, k = 32 | 64, bf16 NHWC
 * in / out, fp32 weights [k][3][3][3]; and its weight gradient (fp32, ACCUMULATED: caller zero-fills). */
int xv2_stem_conv_fwd(const void* x, const float* w, void* y, int32_t n, int32_t h, int32_t wd, int32_t k, void* stream);
int xv2_stem_conv_wgrad(const void* x, const void* dy, float* dw, int32_t n, int32_t h, int32_t wd, int32_t k, void* stream);
/* Fully connected layers of split attention on [n][c] fp32 vectors (SplAtConv2d fc1 / fc2, 1x1 convs on a 1x1 image):
 * y[n][k] = b[k] + sum_c x[n][c] w[k][c];   dw[k][c] = sum_n dy[n][k] x[n][c], db[k] = sum_n dy[n][k] (both WRITTEN). */
int xv2_fc_fwd(const float* x, const float* w, const float* b, float* y, int32_t n, int32_t c, int32_t k, void* stream);
int xv2_fc_wgrad(const float* x, const float* dy, float* dw, float* db, int32_t n, int32_t c, int32_t k, void* stream);

/* Sum over pixels of a [pixels][k] tensor -> fp32 [k] (bias gradient of the 1x1 output head, layers.py:180). */
int xv2_colsum(const void* x, int64_t pixels, int32_t k, int32_t dtype, float* out, void* stream);

/* fp32 master weight [A][R][S][B] (physical channels-last OIHW, A = all out channels, B = in per group) ->
 * packed `dst_dtype` weight.  mode 0: same order.  mode 1: dgrad / transposed-conv order
 * dst[g*B + b][R-1-r][S-1-s][a_in_group] (flipped taps, in/out swapped per group).
 * mode 2: dst[r][s][b][a] (transposed-conv GEMM rows for xv2_conv_tc convt=1; groups must be 1). */
int xv2_pack_weight(const float* src, void* dst, int32_t a, int32_t r, int32_t s, int32_t b, int32_t groups,
                    int32_t mode, int32_t dst_dtype, void* stream);

/* Every packed copy of the model in ONE launch (after the optimizer step rewrote the masters): `jobs` is a DEVICE array. */
typedef struct xv2_pack_job {
  const void* src;          /* fp32 master [a][r][s][b] */
  void* dst;                /* packed copy */
  int32_t a, r, s, b, groups, mode, dst_dtype, pad_;
} xv2_pack_job;
int xv2_pack_weights_batched(const xv2_pack_job* jobs, int32_t njobs, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Tensor-core (tcgen05 + TMEM + TMA) implicit-GEMM convolution, bf16 in / fp32 accumulate.
 * Stride-1 "same" convolutions only (3x3 p=dil, 1x1), groups allowed, optional second source that is
 * concatenated on the channel axis WITHOUT materialising torch.cat (layers.py:167 / layers.py:114),
 * and the 2x2 stride-2 transposed convolution as GEMM + pixel-shuffle store (layers.py:83).
 * Returns XV2_EUNSUPPORTED when the shape is not eligible (caller falls back to the SIMT entry).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct xv2_tc_conv {
  int32_t n, h, w;          /* spatial dims of sources (and of the output unless convt) */
  int32_t c0, c1;           /* channels of source 0 / source 1 (c1 = 0: single source), per tensor */
  int32_t ld0, ld1;         /* pixel stride (elements) of the sources; 0 = contiguous (= c0 / c1) */
  int32_t k;                /* output channels (all groups) */
  int32_t r, s, pad, dil;   /* taps, padding, dilation (stride is 1) */
  int32_t groups;           /* c1 must be 0 when groups > 1 */
  int32_t convt;            /* 1: w is [4*k][c0] (tap-major rows), out is (n, 2h, 2w, k): 2x2 s2 transposed conv
                             * 2: its data gradient: src0 is (n, 2h, 2w, c0), w is [k][2][2][c0], out is (n, h, w, k) */
  int32_t out_dtype;        /* XV2_BF16 | XV2_F32 */
  int32_t ldo;              /* output pixel stride in elements; 0 = k */
} xv2_tc_conv;

/* w: bf16 [k][r][s][(c0+c1)/groups] (convt: [2][2][k][c0]); bias fp32 [k] or NULL.
 * stats: optional fp64 [2*k] accumulators (sum, sum of squares of the ROUNDED outputs per channel), caller zero-fills:
 *        the batch-norm statistics of nn.BatchNorm2d (layers.py:93) fused into the conv epilogue. */
int xv2_conv_tc(const xv2_tc_conv* p, const void* src0, const void* src1, const void* w, const float* bias,
                void* out, double* stats, void* stream);

/* Inference form: eval-mode BatchNorm (+ activation) folded into the conv epilogue, out = act(scale[k] * conv + shift[k])
 * with scale / shift from xv2_bn_eval_coeffs (layers.py:92-94 in eval mode: no separate BN pass at all). */
int xv2_conv_tc_bnact(const xv2_tc_conv* p, const void* src0, const void* src1, const void* w, const float* scale,
                      const float* shift, int32_t act, void* out, void* stream);

/* Tensor-core weight gradient: dw[k][r][s][c] fp32 (accumulated with atomics; caller zero-fills), bf16 operands.
 * src0/src1 as above (c split the same way), dout [n][h][w][k] with pixel stride lddo (0 = k).
 * convt = 1: dw is [c0(row)][2][2][k]: gradient of the transposed conv given x = src0 (n,h,w,c0), dout (n,2h,2w,k). */
int xv2_wgrad_tc(const xv2_tc_conv* p, const void* src0, const void* src1, const void* dout, int32_t lddo,
                 float* dw, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Batch normalisation (+ activation, + residual), NHWC.   nn.BatchNorm2d: layers.py:93,72; encoder BNs unet.py:52
 * ---------------------------------------------------------------------------------------------------------- */
/* per-channel sum and sum of squares over pixels -> fp64 stats[2*c] (accumulated; caller zero-fills) */
int xv2_bn_stats(const void* x, int64_t pixels, int32_t c, int32_t dtype, double* stats, void* stream);
/* training: stats -> mean/invstd (saved, fp32 [c] each), scale/shift (fp32 [c] each), running stats update
 * (momentum, unbiased variance).  count = pixels the stats were taken over. */
int xv2_bn_finalize(const double* stats, int64_t count, int32_t c, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float momentum, float eps, float* mean, float* invstd,
                    float* scale, float* shift, void* stream);
/* eval: scale/shift from running statistics */
int xv2_bn_eval_coeffs(int32_t c, const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float eps, float* scale, float* shift, void* stream);
/* y = act(scale*x + shift (+ residual)) */
int xv2_bn_apply(const void* x, const void* residual, void* y, int64_t pixels, int32_t c, int32_t dtype,
                 const float* scale, const float* shift, int32_t act, void* stream);
/* training forward in one call: finalize (as xv2_bn_finalize, coef = [4][c] mean | invstd | scale | shift, running statistics
 * updated) + apply; on the streaming path the finalize step is folded into the apply kernel's prologue (no extra launch). */
int xv2_bn_train_apply(const void* x, const void* residual, void* y, int64_t pixels, int32_t c, int32_t dtype,
                       const double* stats, int64_t count, const float* gamma, const float* beta, float* running_mean,
                       float* running_var, float momentum, float eps, float* coef, int32_t act, void* stream);
/* backward pass 1: through the activation (recomputed from x, scale, shift, residual) then the two BN reductions.
 * red (fp64 [2*c], accumulated; caller zero-fills) = (sum du, sum du * xhat).  */
int xv2_bn_bwd_reduce(const void* dy, const void* x, const void* residual, int64_t pixels, int32_t c, int32_t dtype,
                      const float* scale, const float* shift, const float* mean, const float* invstd, int32_t act,
                      double* red, void* stream);
/* backward pass 1 for an activation whose gradient arrives in two parts (a residual join: the next block's conv1 data gradient
 * `dy` and its shortcut gradient `dy2`, unet.py:52 Bottleneck.forward `out += residual`): sums them in the kernel instead of an
 * autograd accumulation pass, writes du = round_bf16(dy [+ dy2]) * act'(u) once (bf16; it is ALSO the shortcut's gradient) and
 * accumulates red from the rounded du; pass 2 then runs with dy = du, act = none, residual = NULL.  bf16 streaming path only
 * (XV2_EUNSUPPORTED otherwise: the caller adds the parts and uses xv2_bn_bwd_reduce).  dy2 and residual may be NULL. */
int xv2_bn_bwd_reduce_du(const void* dy, const void* dy2, const void* x, const void* residual, void* du, int64_t pixels,
                         int32_t c, int32_t dtype, const float* scale, const float* shift, const float* mean,
                         const float* invstd, int32_t act, double* red, void* stream);
/* backward pass 2: dx = gamma*invstd*(du - mean(du) - xhat*mean(du*xhat)) [train] or du*scale [eval: red == NULL];
 * dres (optional) = du; dgamma/dbeta (fp32 [c]) from red: written, or ADDED TO when accumulate != 0 (the parameter's own
 * gradient slot in the flat buffer, so that no separate accumulation launch is needed). */
int xv2_bn_bwd_apply(const void* dy, const void* x, const void* residual, void* dx, void* dres, int64_t pixels,
                     int32_t c, int32_t dtype, const float* scale, const float* shift, const float* mean,
                     const float* invstd, const float* gamma, int32_t act, const double* red, int64_t count,
                     float* dgamma, float* dbeta, int32_t accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Pooling, NHWC.  nn.MaxPool2d(3,2,1) unet.py:81; ResNeSt avd AvgPool2d(3,s,1) and avg-down AvgPool2d(s,s,ceil) unet.py:52
 * ---------------------------------------------------------------------------------------------------------- */
/* idx (optional, uint8 [n][oh][ow][c]): window position r*k+s of the first maximum in scan order (ATen's tie rule) */
int xv2_maxpool_fwd(const void* x, void* y, uint8_t* idx, int32_t n, int32_t h, int32_t w, int32_t c, int32_t oh,
                    int32_t ow, int32_t k, int32_t stride, int32_t pad, int32_t dtype, void* stream);
/* gathers dy through the saved index map; dx is fully written (no atomics) */
int xv2_maxpool_bwd(const uint8_t* idx, const void* dy, void* dx, int32_t n, int32_t h, int32_t w, int32_t c,
                    int32_t oh, int32_t ow, int32_t k, int32_t stride, int32_t pad, int32_t dtype, void* stream);
int xv2_avgpool_fwd(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t oh, int32_t ow,
                    int32_t k, int32_t stride, int32_t pad, int32_t count_include_pad, int32_t dtype, void* stream);
int xv2_avgpool_bwd(const void* dy, void* dx, int32_t n, int32_t h, int32_t w, int32_t c, int32_t oh, int32_t ow,
                    int32_t k, int32_t stride, int32_t pad, int32_t count_include_pad, int32_t dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Split attention (ResNeSt SplAtConv2d, radix 2, cardinality 1; call site unet.py:52)
 * x is [n][hw][2*c] (radix-major channel halves).
 * ---------------------------------------------------------------------------------------------------------- */
/* gap[n][c] (fp32) = mean_hw (x[..,c] + x[..,c+C]) */
int xv2_splat_gap(const void* x, float* gap, int32_t n, int64_t hw, int32_t c, int32_t dtype, void* stream);
/* att[n][2c] (fp32) = softmax over the radix pair of logits[n][2c] */
int xv2_rsoftmax_fwd(const float* logits, float* att, int32_t n, int32_t c, void* stream);
int xv2_rsoftmax_bwd(const float* att, const float* datt, float* dlogits, int32_t n, int32_t c, void* stream);
/* The FC chain between GAP and the attention weights, fused (n <= 32 samples):
 *   forward : z1 = fc1(gap) [n][inter] -> BatchNorm over the n samples (training != 0: batch statistics, running statistics
 *             updated; else running statistics) -> relu = a1 -> fc2 -> r-softmax = att [n][2c].
 *             coef = [4][inter] mean | invstd | scale | shift (saved for backward).  Replaces fc1 / bn1 / relu / fc2 / rSoftMax.
 *   backward: from datt [n][2c]; w2t = fc2 weight transposed [inter][2c], w1t = fc1 weight transposed [c][inter];
 *             writes dw2 [2c][inter], db2, dw1 [inter][c], db1, dgamma, dbeta, dgap [n][c]; dz2 / dz1 are scratch. */
int xv2_splat_fc_fwd(const float* gap, const float* w1, const float* b1, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, float momentum, float eps, int32_t training, const float* w2,
                     const float* b2, float* z1, float* a1, float* coef, float* att, int32_t n, int32_t c, int32_t inter,
                     void* stream);
int xv2_splat_fc_bwd(const float* att, const float* datt, const float* a1, const float* z1, const float* coef,
                     const float* gamma, const float* gap, const float* w2t, const float* w1t, int32_t training, float* dz2,
                     float* dz1, float* dw2, float* db2, float* dw1, float* db1, float* dgamma, float* dbeta, float* dgap,
                     int32_t n, int32_t c, int32_t inter, void* stream);
/* out[n][hw][c] = att[n][c]*x[..,c] + att[n][C+c]*x[..,C+c] */
int xv2_splat_combine(const void* x, const float* att, void* out, int32_t n, int64_t hw, int32_t c, int32_t dtype,
                      void* stream);
/* datt[n][2c] (fp32, accumulated; caller zero-fills) = sum_hw dout[..,c] * x[..,r*C+c] */
int xv2_splat_bwd_att(const void* x, const void* dout, float* datt, int32_t n, int64_t hw, int32_t c, int32_t dtype,
                      void* stream);
/* dx[n][hw][2c] = att[n][r*C+c]*dout[..,c] + dgap[n][c]/hw */
int xv2_splat_bwd_x(const void* dout, const float* att, const float* dgap, void* dx, int32_t n, int64_t hw,
                    int32_t c, int32_t dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Element-wise helpers
 * ---------------------------------------------------------------------------------------------------------- */
/* y = a + b (optionally relu) ; backward of relu handled by xv2_relu_bwd */
int xv2_add_act(const void* a, const void* b, void* y, int64_t numel, int32_t dtype, int32_t act, void* stream);
/* dx = dy * (y > 0 ? 1 : slope(act)) */
int xv2_act_bwd(const void* dy, const void* y, void* dx, int64_t numel, int32_t dtype, int32_t act, void* stream);
/* attention gate (layers.py:165-166): out[p][c] = skip[p][c] * sigmoid(psi[p]);  psi is [pixels] in `dtype` */
int xv2_gate_fwd(const void* skip, const void* psi, void* out, int64_t pixels, int32_t c, int32_t dtype, void* stream);
/* dskip = dout*sig ; dpsi[p] = sig*(1-sig) * sum_c dout[p][c]*skip[p][c] */
int xv2_gate_bwd(const void* dout, const void* skip, const void* psi, void* dskip, void* dpsi, int64_t pixels,
                 int32_t c, int32_t dtype, void* stream);
/* flip an NHWC tensor along H and/or W (TTA, plt.py:30,46) */
int xv2_flip(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t flip_h, int32_t flip_w,
             int32_t dtype, void* stream);
/* dst (bf16|f32) <- src (bf16|f32) element-wise cast */
int xv2_cast(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t numel, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Loss, metric, post-process.  logits are fp32 NHWC [n][h][w][ncls]; labels uint8 [n][H][W] sampled with stride
 * `lstride` (deep supervision nearest down-sampling, plt.py:73: label[::f, ::f]).
 * ---------------------------------------------------------------------------------------------------------- */
#define XV2_LOSS_DICE 1
#define XV2_LOSS_FOCAL 2
#define XV2_LOSS_CE 4 /* 'ce' and the reference's effective 'ohem' (loss.py:45 never truncates) */

/* pass 1: partial sums.  sums fp64 [3*ncls + 3], accumulated (caller zero-fills):
 *   [0..ncls) sum p_c*t_c   [ncls..2ncls) sum p_c   [2ncls..3ncls) sum t_c   then focal sum, ce sum, pixel count.
 * post != 0: only pixels with label > 0 count and labels are shifted by -1 (loss.py:86-90). */
int xv2_loss_partials(const float* logits, const uint8_t* labels, int32_t n, int32_t h, int32_t w, int32_t ncls,
                      int32_t lstride, int32_t post, double* sums, void* stream);
/* pass 2: loss[0] (fp32) += weight * (selected terms); terms = OR of XV2_LOSS_*  (loss.py:98-101, plt.py:74-76).
 * Also writes coef (fp32 [2*ncls + 4]) used by the backward kernel. */
int xv2_loss_finalize(const double* sums, int32_t ncls, int32_t terms, float weight, float* loss, float* coef,
                      void* stream);
/* backward: dlogits = dloss[0] * d(loss)/d(logits), one pass */
int xv2_loss_backward(const float* logits, const uint8_t* labels, int32_t n, int32_t h, int32_t w, int32_t ncls,
                      int32_t lstride, int32_t post, int32_t terms, const float* coef, const float* dloss,
                      float* dlogits, void* stream);

/* F1.update (utils/f1.py:28-42): counters int64 [3*(ncls_metric-1)] = tp | fp | fn, accumulated.
 * ncls_metric 2: pred = argmax(logits[2]); 5: logits have 4 channels, pred = argmax+1, only pixels with label > 0.
 * Optionally writes the uint8 prediction map (argmax, lowest index wins ties). */
int xv2_f1_update(const float* logits, const uint8_t* labels, int64_t pixels, int32_t ncls_metric, int64_t* counters,
                  uint8_t* pred_map, void* stream);
/* mean over TTA passes folded into the consumer: out = (a + b + c + d) * 0.25 (plt.py:44-47) */
int xv2_mean4(const float* a, const float* b, const float* c, const float* d, float* out, int64_t numel, void* stream);
/* Model.save + utils/post_process.py:27-38 fused: from localisation logits (2 ch) and damage logits (4 ch):
 *   loc = sigmoid(loc_logit[1]); post = argmax(softmax(dmg)) + 1; pre = loc>0.3 | (loc>0.1 & post>1); post *= pre. */
int xv2_post_process(const float* loc_logits, const float* dmg_logits, int64_t pixels, uint8_t* pre_map,
                     uint8_t* post_map, void* stream);
/* same rule from probabilities as post_process.py reads them from .npy: loc [pixels], dmg [4][pixels] (planar) */
int xv2_post_process_probs(const float* loc, const float* dmg, int64_t pixels, uint8_t* pre_map, uint8_t* post_map,
                           void* stream);

/* Connected-component majority vote (utils/post_process.py:39-43): every 4-connected component of post > 0 takes its most
 * frequent class (1..4; ties -> smallest).  post / out uint8 [n][h][w]; labels int32 [n*h*w] and votes int32 [n*h*w][4] scratch. */
int xv2_cc_majority_vote(const uint8_t* post, uint8_t* out, int32_t* labels, int32_t* votes, int32_t n, int32_t h, int32_t w,
                         void* stream);
/* Grey-scale dilation with a k x k square footprint, k odd (post_process.py:44-45: skimage dilation(img, square(k))) */
int xv2_dilate_square(const uint8_t* in, uint8_t* out, int32_t n, int32_t h, int32_t w, int32_t k, void* stream);
/* xView2 scorer counters (utils/xview2_metrics.py:61-92), accumulated: counters int64 [15] = lTP lFN lFP, then (TP FN FP) of
 * damage classes 1..4 on target-building pixels with the damage prediction masked by the predicted buildings */
int xv2_score_counts(const uint8_t* loc_pred, const uint8_t* dmg_pred, const uint8_t* loc_targ, const uint8_t* dmg_targ,
                     int64_t pixels, int64_t* counters, void* stream);

/* Model.save (plt.py:126-131): probabilities in the layout the reference writes per tile with np.save:
 * ncls 2: out[n][hw] = sigmoid(logit[..,1]);  ncls 4: out[n][4][hw] = softmax (planar). */
int xv2_save_probs(const float* logits, int32_t n, int64_t hw, int32_t ncls, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * ResNeSt split attention with its BatchNorm (bn0) + ReLU folded in (call site unet.py:52): z = RAW radix-conv output, bf16
 * [n][hw][2c]; scale / shift / mean / invstd fp32 [2c] as written by xv2_bn_finalize; the post-BN activation is never stored.
 * XV2_EUNSUPPORTED unless c / 8 is a power of two <= 128 (the caller then runs the unfused chain).
 * ---------------------------------------------------------------------------------------------------------- */
/* xv2_splat_bn_gap with the bn0 finalize step folded in (one launch fewer per bottleneck): the coefficients come straight from
 * the fp64 (sum, sum of squares) `stats` of the radix conv's epilogue; block (0,0) writes coef = [4][2c] mean | invstd | scale |
 * shift for the later kernels and updates the running statistics exactly as xv2_bn_finalize does.  The pooled sums are accumulated
 * in fp64 in `gap_acc` ([n][c], ZERO-FILLED by the caller: the step's scratch arena) and converted to gap = fp32(sum / hw) by a
 * second, tiny launch: the result does not depend on the order in which the blocks arrive (run-to-run reproducible forward). */
int xv2_splat_bn_gap_fin(const void* z, const double* stats, int64_t count, const float* gamma, const float* beta,
                         float* running_mean, float* running_var, float momentum, float eps, float* coef, float* gap,
                         double* gap_acc, int32_t n, int64_t hw, int32_t c, void* stream);
/* xv2_splat_fc_bwd for the bn0-fused path with the two tiny neighbours folded in (two launches fewer per bottleneck): datt is
 * derived from `part` (as xv2_splat_bn_bwd_datt) inside the first FC kernel, and the bn0 reductions `red` (as
 * xv2_splat_bn_bwd_red) are finished by the last one, which has just produced dgap.  accumulate != 0: dw2 / db2 / dw1 / db1 /
 * dgamma / dbeta are ADDED TO (the parameters' own slots of the flat gradient buffer: no autograd accumulation launches). */
int xv2_splat_fc_bwd_fused(const float* att, const double* part, const float* scale0, const float* shift0, const float* mean0,
                           const float* invstd0, int64_t hw, const float* a1, const float* z1, const float* coef,
                           const float* gamma, const float* gap, const float* w2t, const float* w1t, int32_t training,
                           float* dz2, float* dz1, float* dw2, float* db2, float* dw1, float* db1, float* dgamma, float* dbeta,
                           float* dgap, double* red, int32_t accumulate, int32_t n, int32_t c, int32_t inter, void* stream);
/* gap[n][c] = mean_hw (relu(bn(z_0)) + relu(bn(z_1)))  (gap is zero-filled here) */
int xv2_splat_bn_gap(const void* z, const float* scale, const float* shift, float* gap, int32_t n, int64_t hw, int32_t c,
                     void* stream);
/* out[n][hw][c] = att_0 relu(bn(z_0)) + att_1 relu(bn(z_1)) */
int xv2_splat_bn_combine(const void* z, const float* scale, const float* shift, const float* att, void* out, int32_t n,
                         int64_t hw, int32_t c, void* stream);
/* backward, one pass over (z, dout): part fp64 [4][n][2c] += A1 | A2 | M1 | M2 with m = [relu input > 0]:
 * A1 = sum dout m, A2 = sum dout m z, M1 = sum m, M2 = sum m z (caller zero-fills) */
int xv2_splat_bn_bwd_partials(const void* z, const void* dout, const float* scale, const float* shift, double* part,
                              int32_t n, int64_t hw, int32_t c, void* stream);
/* datt[n][2c] = sum_hw dout * relu(bn(z)) = scale * A2 + shift * A1 */
int xv2_splat_bn_bwd_datt(const double* part, const float* scale, const float* shift, float* datt, int32_t n, int32_t c,
                          void* stream);
/* BatchNorm reductions of du = (att dout + dgap / hw) m from the partial sums: red fp64 [2][2c] = (sum du, sum du * xhat) (written) */
int xv2_splat_bn_bwd_red(const double* part, const float* att, const float* dgap, const float* mean, const float* invstd,
                         double* red, int32_t n, int64_t hw, int32_t c, void* stream);
/* dz = gamma invstd (du - mean(du) - xhat mean(du xhat)); dgamma / dbeta [2c] written or accumulated as xv2_bn_bwd_apply */
int xv2_splat_bn_bwd_apply(const void* z, const void* dout, const float* att, const float* dgap, const float* scale,
                           const float* shift, const float* mean, const float* invstd, const float* gamma, const double* red,
                           void* dz, float* dgamma, float* dbeta, int32_t accumulate, int32_t n, int64_t hw, int32_t c,
                           void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused tail of the localisation network: the last ConvLayer's BatchNorm + LeakyReLU (layers.py:96-100) and the 1x1 head
 * (layers.py:180-183) without materialising the full-resolution activation between them.  z = raw conv output, bf16
 * [pixels][c], c = 32 | 64; scale/shift/mean/invstd as written by xv2_bn_finalize; ncls 1..4; logits / dlogits fp32 [pixels][ncls].
 * XV2_EUNSUPPORTED for other shapes (the caller then runs xv2_bn_train_apply + xv2_head_fwd).
 * ---------------------------------------------------------------------------------------------------------- */
int xv2_bnact_head_fwd(const void* z, int64_t pixels, int32_t c, const float* scale, const float* shift, int32_t act,
                       const float* head_w, const float* head_b, int32_t ncls, float* logits, void* stream);
/* backward pass 1: red fp64 [2c] += (sum du, sum du*xhat) with du = (dlogits . W) * act'(.), dhead_w [ncls][c] += dlogits^T y,
 * dhead_b [ncls] += sum dlogits (all accumulated; caller zero-fills) */
int xv2_bnact_head_bwd_reduce(const void* z, const float* dlogits, int64_t pixels, int32_t c, const float* scale,
                              const float* shift, const float* mean, const float* invstd, int32_t act, const float* head_w,
                              int32_t ncls, double* red, float* dhead_w, float* dhead_b, void* stream);
/* backward pass 2: dz = gamma*invstd*(du - mean(du) - xhat*mean(du*xhat)); dgamma / dbeta as xv2_bn_bwd_apply */
int xv2_bnact_head_bwd_apply(const void* z, const float* dlogits, void* dz, int64_t pixels, int32_t c, const float* scale,
                             const float* shift, const float* mean, const float* invstd, const float* gamma, int32_t act,
                             const float* head_w, int32_t ncls, const double* red, int64_t count, float* dgamma,
                             float* dbeta, int32_t accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Optional model parts: pyramid pooling, bilinear decoders / heads, ordinal damage heads.
 * ---------------------------------------------------------------------------------------------------------- */
/* nn.AdaptiveAvgPool2d(bins) of PPM (layers.py:12-21): y[n][bins][bins][c]; bin i covers [floor(i L/bins), ceil((i+1) L/bins)) */
int xv2_adaptive_avgpool_fwd(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t bins,
                             int32_t dtype, void* stream);
int xv2_adaptive_avgpool_bwd(const void* dy, void* dx, int32_t n, int32_t h, int32_t w, int32_t c, int32_t bins,
                             int32_t dtype, void* stream);
/* F.interpolate(mode="bilinear", align_corners=True) (layers.py:27, :154, :186-188): (n,h,w,c) -> (n,oh,ow,c) */
int xv2_bilinear_fwd(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t oh, int32_t ow,
                     int32_t dtype, void* stream);
/* its gradient, accumulated into a ZERO-FILLED fp32 buffer dx_f32[n][h][w][c] (fp32 atomics) */
int xv2_bilinear_bwd(const void* dy, float* dx_f32, int32_t n, int32_t h, int32_t w, int32_t c, int32_t oh, int32_t ow,
                     int32_t dtype, void* stream);
/* Ordinal damage heads (unet.py:21-26).  mode 0: 'mse' -- nn.MSELoss(relu(logit[0]), label) (loss.py:92-94), 1 logit per pixel;
 * mode 1: 'coral' (loss.py:54-65), 3 rank logits per pixel.  `post` masking as loss.py:86-90.
 * partials: sums fp64 [2] += (loss sum, kept pixels); finalize: loss = weight * sum / pixels, coef[0] = weight / pixels;
 * backward: dlogits = upstream[0] * d(loss)/d(logits), zero outside the mask. */
int xv2_ordinal_loss_partials(const float* logits, const uint8_t* labels, int64_t pixels, int32_t mode, int32_t post,
                              double* sums, void* stream);
int xv2_ordinal_loss_finalize(const double* sums, float weight, float* loss, float* coef, void* stream);
int xv2_ordinal_loss_backward(const float* logits, const uint8_t* labels, int64_t pixels, int32_t mode, int32_t post,
                              const float* coef, const float* upstream, float* dlogits, void* stream);
/* convert_to_labels (utils/f1.py:7-15): mode 0 -> round(relu(x0)) + 1 (clamped to 4 when clamp4, as F1 does; Model.save
 * plt.py:131 does not clamp), mode 1 -> #(sigmoid(x_c) > 0.5) + 1.  Any of: F1 counters int64 [12] = tp | fp | fn over pixels
 * with label > 0 (utils/f1.py:31-42), a uint8 label map, an fp32 label map (what Model.save stores). */
int xv2_ordinal_labels(const float* logits, const uint8_t* labels, int64_t pixels, int32_t mode, int32_t clamp4,
                       int64_t* counters, uint8_t* pred_u8, float* pred_f32, void* stream);

/* 1x1 output head (layers.py:180): logits[p][ncls] (fp32) = x[p][c] . w[ncls][c] + b ; ncls <= 8 */
int xv2_head_fwd(const void* x, const float* w, const float* b, float* logits, int64_t pixels, int32_t c,
                 int32_t ncls, int32_t dtype, void* stream);
/* dx[p][c] = sum_k dlogits[p][k] w[k][c] ; dw[ncls][c], db[ncls] fp32 accumulated (caller zero-fills) */
int xv2_head_bwd(const void* x, const float* w, const float* dlogits, void* dx, float* dw, float* db, int64_t pixels,
                 int32_t c, int32_t ncls, int32_t dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Loader: decoded uint8 HWC tiles (cv2 BGR order, pytorch_loader.py:40) -> normalised NHWC activations
 * (A.Normalize, pytorch_loader.py:63: (x - 255*mean) * (1/(255*std)), constants applied by channel position).
 * pre (and optional post) are [n][h][w][3] uint8; out is [n][h][w][3 or 6].
 * ---------------------------------------------------------------------------------------------------------- */
int xv2_normalize_tiles(const uint8_t* pre, const uint8_t* post, void* out, int32_t n, int32_t h, int32_t w,
                        int32_t out_dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Train-time augmentation on the device (pytorch_loader.py:57-63,73-92,109-115,124-148): RandomScale (cubic / nearest) ->
 * CropNonEmptyMaskIfExists -> H / V flip -> GaussNoise -> RandomBrightnessContrast -> Normalize, as ONE gather kernel.
 * params fp32 [n][16] per sample (host-drawn decisions):
 *   0 src_w/scaled_w  1 src_h/scaled_h  2 scaled_w  3 scaled_h  4 crop x0  5 crop y0  6 flip_h  7 flip_v  8 sigma(pre)  9 sigma(post)
 *   10 alpha(pre)  11 beta(pre)  12 alpha(post)  13 beta(post)  14 noise seed  15 zoom on
 * ---------------------------------------------------------------------------------------------------------- */
/* Crop origin (x0, y0) in the SCALED mask: the floor(u0 * count)-th non-zero pixel minus floor(u1 * cw), floor(u2 * ch), clipped;
 * an empty mask gives a uniform origin.  uniforms fp32 [n][3] in [0,1); rowcount int32 [n][max_rows] scratch; origin int32 [n][2]. */
int xv2_crop_origin(const uint8_t* mask, const float* params, const float* uniforms, int32_t* rowcount, int32_t* origin,
                    int32_t n, int32_t sh, int32_t sw, int32_t max_rows, int32_t ch, int32_t cw, void* stream);
/* pre / post uint8 [n][sh][sw][3] (post optional), mask uint8 [n][sh][sw]; origin int32 [n][2] (null: params 4,5);
 * out [n][oh][ow][3|6] normalised (out_dtype) and / or out_u8 (the augmented bytes before Normalize); mask_out uint8 [n][oh][ow]. */
int xv2_augment_tiles(const uint8_t* pre, const uint8_t* post, const uint8_t* mask, const float* params, const int32_t* origin,
                      void* out, uint8_t* out_u8, uint8_t* mask_out, int32_t n, int32_t sh, int32_t sw, int32_t oh, int32_t ow,
                      int32_t out_dtype, void* stream);

/* Fused AdamW over one flat parameter buffer (torch.optim.AdamW semantics, plt.py:154): step is 1-based; the gradient
 * is multiplied by grad_scale first (1/world_size after the SUM all-reduce of the data-parallel ranks). */
int xv2_adamw(float* p, const float* g, float* m, float* v, int64_t numel, float lr, float beta1, float beta2,
              float eps, float weight_decay, int32_t step, float grad_scale, void* stream);
/* SGD with momentum (apex FusedSGD defaults, plt.py:152): buf = momentum*buf + g (buf = g at step 1); p -= lr*buf. */
int xv2_sgd(float* p, const float* g, float* buf, int64_t numel, float lr, float momentum, float grad_scale,
            int32_t step, void* stream);
/* The same updates with every per-step scalar read from DEVICE memory, so that the launch can be replayed from a CUDA graph
 * (the host refreshes the 32-byte block with one async copy per step):
 *   adamw hyper fp32 [8] = lr, beta1, beta2, eps, weight_decay, 1 - beta1^step, sqrt(1 - beta2^step), grad_scale
 *   sgd   hyper fp32 [4] = lr, momentum, grad_scale, first (1 on the first step) */
int xv2_adamw_dev(float* p, const float* g, float* m, float* v, int64_t numel, const float* hyper, void* stream);
int xv2_sgd_dev(float* p, const float* g, float* buf, int64_t numel, const float* hyper, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XV2_H */
