/*
 * combat_b200 -- C-ABI of the B200-native (sm_100a) kernels behind COMBAT's alternated
 * generator/surrogate training step and its batched 2-D DCT.
 *
 * The reference (VinAIResearch/COMBAT) is pure Python/PyTorch and has no FFI of its own
 * (SURVEY.md section 8b): the boundary a maintainer binds is this header, from Python via ctypes
 * (combat_b200/_lib.py; INTEGRATION.md shows the stub).  Conventions:
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - no allocation inside; callers provide outputs and workspaces;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, and
 *     may be captured into a CUDA graph;
 *   - return value: 0 on success, a negative cudaError_t otherwise, -1000-k for argument k invalid;
 *   - `dtype`: 0 = float32, 1 = bfloat16 (storage type of activations; statistics, losses,
 *     master weights, gradients of parameters are always float32);
 *   - images at the reference-facing boundary are NCHW float32 (as the reference's tensors);
 *     activations inside the networks are NHWC.
 * Each entry point cites the reference code it replaces (paths relative to the reference root).
 */
#ifndef COMBAT_B200_H
#define COMBAT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COMBAT_F32 0
#define COMBAT_BF16 1

/* library / device info */
int combat_version(void);
/* counts every kernel launched through this library since load (bench.py's gpu_launches) */
long long combat_launch_count(void);
const char* combat_last_error(void);

/* ---------------------------------------------------------------- plane transforms (DCT family)
 * out[p] = L * X[p] * R^T for `planes` contiguous N x N planes (row-major, NCHW image planes).
 *   utils/dct.py:85-96   dct_2d   : L = R = D   (orthonormal DCT-II matrix)
 *   utils/dct.py:99-111  idct_2d  : L = R = D^T
 *   train_generator.py:47-55 low_freq : L = R = P = D^T diag(1_k) D  (also its own backward: P symmetric)
 * in_mode: 0 = float32 input, 1 = uint8 input (train_generator.py:245, defenses/frequency_based/train.py:195-197),
 *          2 = float32 input quantised on the fly as ((x+1)/2*255).byte() (train_generator.py:245).
 * L, R: N*N float32 row-major on device.  workspace: planes*N*N floats, only touched when N > 64.
 * fast != 0 and N == 32 and L == R selects the register butterfly kernels: kind 1 = DCT-II, 2 = DCT-III,
 * 3 = low-pass projection with `keep` retained coefficients per axis (L/R unused).
 */
int combat_plane_transform(const void* in, float* out, const float* L, const float* R, long long planes, int N,
                           int in_mode, float* workspace, void* stream);
int combat_dct32_fast(const void* in, float* out, long long planes, int kind, int keep, int in_mode, void* stream);
/* same for 64 x 64 planes (CelebA shape), one 64-thread CTA per plane */
int combat_dct64_fast(const void* in, float* out, long long planes, int kind, int keep, int in_mode, void* stream);

/* ---------------------------------------------------------------- poisoned-batch builder
 * train_generator.py:188-195 (C-step) and :223-226 (G-step), torchvision GaussianBlur(3) with reflect padding.
 *   out[r] = r < num_bd ? blur3x3(clamp(x[perm[r]] + noise[nperm[r]] * noise_rate, -1, 1)) : x[perm[r]]
 * perm/nperm may be NULL (identity).  k0,k1: the normalised 1-D Gaussian taps (centre, side) for this call's sigma.
 * sq_partial (optional, rows*C floats): per-plane sum((out-x[perm[r]])^2) for the MSE term (:234).
 * taps_dev (optional, 2 floats) / num_bd_dev (optional, 1 int): device-resident overrides of k0,k1 / num_bd, so that a
 * captured CUDA graph follows the per-iteration sigma draw and poison count.
 * taps_rows (optional, rows*2 floats on device): per-ROW taps -- train_generator_multilabel.py:67-75,203-220 blurs every
 * class chunk of the batch with its own sigma draw.
 */
int combat_poison_blend_fwd(const float* x, const float* noise, const int* perm, const int* nperm, int rows, int num_bd,
                            float noise_rate, float k0, float k1, float* out, float* sq_partial, int C, int H, int W,
                            const float* taps_dev, const int* num_bd_dev, const float* taps_rows, void* stream);
/* backward of blur(clamp(x + noise*rate)) w.r.t. noise (G-step, train_generator.py:225-226,254):
 *   g = g1 + g2 + mse_scale * (x_bd - x);  dnoise = rate * [|x + noise*rate| <= 1] * blur^T(g)
 * g2 may be NULL.  mse_scale = 2 * L2_weight / numel. */
int combat_poison_blend_bwd(const float* x, const float* noise, const float* x_bd, const float* g1, const float* g2,
                            float mse_scale, float noise_rate, float k0, float k1, float* dnoise, int rows, int C, int H,
                            int W, const float* taps_dev, const float* taps_rows, void* stream);

/* ---------------------------------------------------------------- losses
 * torch.nn.CrossEntropyLoss (mean) forward+backward and argmax metrics, train_generator.py:207,231,251,262-267.
 *   loss_out[0]  = mean_b( logsumexp(logits[b]) - logits[b, t[b]] )
 *   dlogits[b,c] = grad_scale * (softmax(logits[b])[c] - [c == t[b]]) / B      (dlogits may be NULL)
 *   counts_out[0] = #{b: argmax == t[b]},  counts_out[1] = #{b: argmax == t2[b]} (t2 may be NULL)
 * A negative target marks an ignored row: no loss, zero gradient, never counted (the mean still divides by B) -- used by
 * the evaluation loop (train_generator.py:366-391) to count only the non-target samples of a fixed-shape batch.
 */
int combat_cross_entropy(const float* logits, const long long* targets, const long long* targets2, int B, int C,
                         float grad_scale, float* loss_out, float* dlogits, int* counts_out, void* stream);
/* out[0] = scale * sum(partial[0..n)) -- finishes the MSE mean (train_generator.py:234) deterministically */
int combat_sum_scale(const float* partial, int n, float scale, float* out, void* stream);

/* ---------------------------------------------------------------- optimiser
 * torch.optim.SGD(momentum, weight_decay, nesterov=True) over a flat parameter buffer, train_generator.py:123-126,212,255:
 *   g += wd*p;  buf = first ? g : mu*buf + g;  p -= lr*(g + mu*buf)
 * lr is read from device memory (lr_dev[0]) so a captured graph follows MultiStepLR. */
int combat_sgd_nesterov(float* p, const float* g, float* buf, long long n, const float* lr_dev, float momentum, float wd,
                        int first_step, void* stream);

/* weight layout preparation: master channels-last (OHWI) float32 -> compute layouts in `dtype`.
 * For every descriptor: fwd[co][kh][kw][ci] and (if dgrad_off >= 0) dgrad[ci][KH-1-kh][KW-1-kw][co]. */
typedef struct {
  long long src_off;   /* float offset into the flat parameter buffer (OHWI) */
  long long fwd_off;   /* element offset into the compute-weight buffer, OHWI */
  long long dgrad_off; /* element offset of the flipped/transposed copy, or -1 */
  int Cout, Cin, KH, KW;
} combat_wprep_desc;
int combat_prep_weights(const float* params, void* wbuf, int dtype, const combat_wprep_desc* table_dev, int n_desc,
                        long long max_elems, void* stream);

/* ---------------------------------------------------------------- convolution (nn.Conv2d)
 * Generic strided descriptor so both NCHW float32 images and NHWC activations can be read/written.
 * forward  : out[n,oh,ow,co] = act( sum_{kh,kw,ci} in[n, (oh*stride-pad+kh)/up, (ow*stride-pad+kw)/up, ci] * w[co,kh,kw,ci] + bias[co] ) (+ residual)
 *            taps whose numerator is negative, not divisible by `up`, or out of range contribute 0.
 *            up == 1 is nn.Conv2d forward; stride == 1 with up == s and the flipped weights is its input gradient.
 * wgrad    : dw[co,kh,kw,ci] (channels-last OHWI float32, ACCUMULATED with atomics) += sum_{n,oh,ow} dy[n,oh,ow,co] * in[n, oh*stride-pad+kh, ow*stride-pad+kw, ci]
 */
typedef struct {
  const void* in;
  const void* w;         /* OHWI, w_dtype */
  void* out;
  const float* bias;     /* or NULL */
  const void* residual;  /* same layout/dtype as out, or NULL */
  const float* post_scale; /* per-channel affine applied after act (eval BatchNorm of FrequencyModel), or NULL */
  const float* post_shift;
  int N, Hi, Wi, Ci, Ho, Wo, Co, KH, KW, stride, pad, up;
  long long in_sn, in_sh, in_sw, in_sc;     /* element strides of `in` */
  long long out_sn, out_sh, out_sw, out_sc; /* element strides of `out` (and residual) */
  int in_dtype, w_dtype, out_dtype;
  int act;               /* 0 none, 1 tanh, 2 ELU(alpha=1) */
} combat_conv_desc;
int combat_conv_simt(const combat_conv_desc* d_host, void* stream);
/* wgrad: `in` = layer input (fwd geometry), `out`/`w` unused; dy given separately (strided like out_s*);
 * if `bias` is non-NULL it is the float32 bias-GRADIENT accumulator: bias[co] += sum_{n,oh,ow} dy[n,oh,ow,co]. */
int combat_conv_wgrad_simt(const combat_conv_desc* d_host, const void* dy, int dy_dtype, float* dw_ohwi, void* stream);

/* Direct kernels for the image-boundary layers (3 input or 3 output channels, 3x3, pad 1; csrc/conv_small.cu).
 *   conv_cin3  : x NCHW float32 [N,3,H,W], w [Co][9][3] -> out NHWC [N,Ho,Wo,Co]; act 0 none / 2 ELU, optional affine;
 *                optional second output out2 = bf16 relu(out * scale2[c] + shift2[c]) (eval BatchNorm+ReLU of the next block)
 *   conv_cout3 : in NHWC [N,H,W,64], w [3][9][64] -> out NCHW float32 [N,3,H,W] (stride 1); act 0 none / 1 tanh
 *   wgrad_cin3 : dw[Co][9][3] += sum dy[pix][co] * x[pix@tap][ci], db[co] += sum dy      (x NCHW float32, dy NHWC)
 *   wgrad_cout3: dw[3][9][64] += sum dz[n,co,h,w] * a[pix@tap][ci], db[co] += sum dz     (a NHWC, dz NCHW float32) */
/* im2col of a 3-channel NCHW float32 image for the tensor-core path: A[N,Ho,Wo,64] bf16 = [hi(27) | 0 | lo(27) | 0]
 * (3x3, pad 1, stride 1|2); a 3 -> Co conv is then combat_conv_tc over A as a 1x1 conv with the filter stored twice. */
int combat_im2col3(const float* x, void* A_bf16, int N, int H, int W, int stride, void* stream);
/* folds a weight gradient taken over the im2col3 operand (dW'[rows][64]: hi columns 0..26, lo columns 32..58) into the
 * filter gradient, accumulating: mode 0 -> dw[rows][27] (3 -> rows conv); mode 1 -> dw[3][9][rows] with the taps reversed and
 * db[j] += colsum[12+j] + colsum[44+j] (rows -> 3 conv; colsum = column sums of the operand, optional) */
int combat_fold_w64(const float* dw64, float* dw, int rows, int mode, const float* colsum, float* db, void* stream);
int combat_conv_cin3(const float* x, const void* w, int w_dtype, const float* bias, void* out, int out_dtype, int N, int H,
                     int W, int Co, int stride, int act, const float* post_scale, const float* post_shift, void* out2,
                     const float* scale2, const float* shift2, void* stream);
int combat_conv_cout3(const void* in, int in_dtype, const void* w, int w_dtype, const float* bias, float* out, int N, int H,
                      int W, int Ci, int act, void* stream);
int combat_wgrad_cin3(const float* x, const void* dy, int dy_dtype, float* dw, float* db, int N, int H, int W, int Co,
                      int stride, void* stream);
int combat_wgrad_cout3(const void* a, int a_dtype, const float* dz, float* dw, float* db, int N, int H, int W, int Ci,
                       void* stream);

/* tcgen05 implicit-GEMM convolution (sm_100a tensor cores, TMA-fed, TMEM accumulators), bf16 NHWC.
 * Requires Ci % 64 == 0 and Co % 64 == 0.  See combat_b200/csrc/conv_tc.cu. */
typedef struct {
  const void* in;        /* NHWC bf16 [N,Hi,Wi,Ci] */
  const void* w;         /* [taps][Co][Ci] bf16 (tap-major OHWI) */
  void* out;             /* NHWC [N,Ho,Wo,Co], bf16 or float32 (out_f32) */
  const float* bias;     /* or NULL */
  const void* residual;  /* NHWC like out, bf16 or float32 (res_f32), or NULL */
  float* stats;          /* optional: per-CTA partial sums [grid][2][Co] (sum, sum of squares over pixels) of the float32
                          * value written to `out` -- train-mode BatchNorm statistics fused into the producer conv;
                          * grid = combat_conv_tc_last_grid() CTAs, reduced by combat_bn_finalize(partial, nblk = grid) */
  int N, Hi, Wi, Ci, Ho, Wo, Co, KH, KW, stride, pad, up;
  int out_f32;           /* 1: `out` is float32 (pre-normalisation tensors keep the unrounded accumulator) */
  int res_f32;           /* 1: `residual` is float32 */
  int act;               /* 0 none, 2 ELU(alpha=1) applied after bias */
  const float* post_scale; /* optional per-channel affine after act (eval BatchNorm of FrequencyModel) */
  const float* post_shift;
  /* fused consumers of v = acc + bias (+ act, affine) + residual; all optional, `out` may be NULL when out2 is given:
   *   mask/mask_scale : v = mask > 0 ? v * mask_scale[c] : 0   (backward of relu(bn_eval(x)): preact_resnet.py:20,22 in
   *                     eval mode; mask = the saved bf16 activation, NHWC like out)
   *   post_add        : v += post_add (bf16 NHWC like out; gradient arriving over an identity shortcut, added after the mask)
   *   out             <- v
   *   out2            <- bf16 relu(v * scale2[c] + shift2[c])  (the eval-mode BatchNorm+ReLU the NEXT conv reads) */
  void* out2;
  const float* scale2;
  const float* shift2;
  const void* mask;
  const float* mask_scale;
  const void* post_add;
  /* optional second operand pair of an INPUT-GRADIENT launch of a stride-2 3x3 conv (up == 2): the gradient arriving through
   * the block's 1x1 stride-2 shortcut conv (preact_resnet.py:26-29,33; resnet.py:25-29) -- in2: bf16 NHWC like `in` (same
   * N, Hi, Wi, Ci), w2: that conv's input-gradient filter [Co][1][Ci] bf16.  Its single tap lands on the even/even output
   * pixels, i.e. it is one more (tensor, filter) tap of parity class (0,0) accumulated in the same TMEM accumulator. */
  const void* in2;
  const void* w2;
  /* optional: backward of a TRAIN-mode relu(bn(x)) in front of this (input-gradient) launch, reduction half fused into the
   * epilogue (preact_resnet.py:20-23 / torch batch_norm backward): bnb_x = the saved pre-normalisation tensor (bf16 NHWC like
   * out), bnb_scale/bnb_shift = the forward's per-channel affine (y = relu(x * scale + shift)), bnb_mean/bnb_invstd = the batch
   * statistics.  out <- g = (x * scale + shift > 0) ? acc : 0 (bf16) and `stats` receives per-CTA partial sums [grid][2][Co]
   * of (g, g * (x - mean) * invstd): exactly what combat_bn_bwd_reduce leaves for combat_bn_bwd_finalize, so the separate
   * reduction pass over (dy, x, y) disappears and combat_bn_bwd_apply runs on g with relu = 0. */
  const void* bnb_x;
  const float* bnb_scale;
  const float* bnb_shift;
  const float* bnb_mean;
  const float* bnb_invstd;
  /* 1: FIRST conv of a classifier / the detector (preact_resnet.py:77, resnet.py:73, model.py:12; 3 -> 64 channels, 3x3, stride 1,
   * pad 1): `in` is the float32 NCHW image [N,3,Hi,Wi] and `w` the filter as [64][64] bf16 = [27 taps x ci | 5 zeros | the same 27
   * again | 5 zeros] (the kernel splits the image into bf16 hi + lo halves, so it enters with ~16 mantissa bits); Ci must be
   * given as 64.  Forward epilogue features only (bias, act, post affine, out2, stats).  Needs Wi a power of two <= 128. */
  int in_nchw3;
} combat_conv_tc_desc;
int combat_conv_tc(const combat_conv_tc_desc* d_host, void* stream);
int combat_conv_tc_wgrad(const combat_conv_tc_desc* d_host, const void* dy, float* dw_ohwi, void* stream);
int combat_conv_tc_supported(const combat_conv_tc_desc* d_host);
/* 64 -> 3 channels, 3x3, stride 1, pad 1 on the tensor pipe (N = 16 accumulator columns, 3 used): the generator's last conv
 * (networks/models.py:316,341; act 1 = tanh) and the input gradient of the classifiers' first conv (preact_resnet.py:77 /
 * resnet.py:73 backward).  in: bf16 NHWC [N,H,W,64]; w: bf16 [3][9][64] (for the input gradient: the flipped / transposed
 * filter); out: float32 NCHW [N,3,H,W].  Needs whole-row tiles: combat_conv_tc_cout3_supported(N, H, W). */
int combat_conv_tc_cout3(const void* in, const void* w, const float* bias, float* out, int N, int H, int W, int act, void* stream);
int combat_conv_tc_cout3_supported(int N, int H, int W);
/* number of CTAs (= partial-sum blocks written to desc.stats) of the most recent combat_conv_tc launch of this thread */
int combat_conv_tc_last_grid(void);

/* ---------------------------------------------------------------- normalisation / activation (NHWC, dtype)
 * BatchNorm2d (classifier_models/preact_resnet.py:20,22,32,35; resnet.py:22-35) */
/* partial per-channel sums over rows [R,C]: partial[(2*blk+0)*C+c] = sum x, [(2*blk+1)*C+c] = sum x^2 ; returns #blocks via nblk_out_host */
int combat_bn_stats(const void* x, int dtype, long long R, int C, float* partial, int max_blocks, int* nblk_out_host,
                    void* stream);
/* train: mean/var from partials -> scale=gamma*invstd, shift=beta-mean*scale, save mean/invstd, update running stats
 *        (momentum, unbiased var); eval (nblk == 0): scale/shift from running stats. */
int combat_bn_finalize(const float* partial, int nblk, long long R, int C, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, float momentum, float eps, float* scale, float* shift,
                       float* save_mean, float* save_invstd, void* stream);
/* eval-mode (scale, shift) of every BatchNorm of a network in one launch: table4[4*i..4*i+3] = indices of (gamma, beta)
 * in `params` and (running_mean, running_var) in `bufs` for flattened channel i:
 *   scale[i] = gamma / sqrt(var + eps), shift[i] = beta - mean * scale[i]   (preact_resnet.py:20,22 under netC.eval()) */
int combat_bn_eval_affine(const float* params, const float* bufs, const int* table4, int n, float eps, float* scale,
                          float* shift, void* stream);
/* y = [relu]( x*scale[c] + shift[c] (+ residual) ).  x_dtype: storage type of the pre-normalisation tensor x (the
 * conv output -- kept float32 next to bf16 activations so that statistics and normalisation see unrounded values);
 * dtype: storage type of residual / y (and of every gradient tensor below). */
int combat_affine_act(const void* x, int x_dtype, const void* residual, void* y, int dtype, long long R, int C,
                      const float* scale, const float* shift, int relu, void* stream);
/* backward of y = relu(x*scale+shift (+res)):  dyh = dy * [y > 0]
 *   train: partial sums of dyh and dyh*xhat per channel (then combat_bn_bwd_finalize, combat_bn_bwd_apply)
 *   eval : dx = dyh * scale directly (combat_bn_bwd_apply with dgamma/dbeta NULL)  */
int combat_bn_bwd_reduce(const void* dy, const void* x, int x_dtype, const void* y, int dtype, long long R, int C, const float* mean,
                         const float* invstd, float* partial, int max_blocks, int* nblk_out_host, int relu, void* stream);
int combat_bn_bwd_finalize(const float* partial, int nblk, int C, float* dgamma, float* dbeta, void* stream);
/* train: dx = gamma*invstd*(dyh - dbeta/R - xhat*dgamma/R); eval (eval_scale != NULL): dx = dyh*eval_scale.
 * dadd (optional) is added to dx (gradient arriving over an identity shortcut); dres (optional) receives dyh. */
int combat_bn_bwd_apply(const void* dy, const void* x, int x_dtype, const void* y, const void* dadd, void* dx, void* dres,
                        int dtype, long long R, int C, const float* gamma, const float* mean, const float* invstd,
                        const float* dgamma, const float* dbeta, const float* eval_scale, int relu, void* stream);

/* the three train-mode calls above as ONE cooperative launch (reduce -> grid sync -> finalize -> grid sync -> apply); C % 8 == 0.
 * partial: scratch as for combat_bn_bwd_reduce; dgamma / dbeta are written and then read back by the apply phase. */
int combat_bn_bwd_fused(const void* dy, const void* x, int x_dtype, const void* y, const void* dadd, void* dx, void* dres,
                        int dtype, long long R, int C, const float* gamma, const float* mean, const float* invstd,
                        float* partial, int max_blocks, float* dgamma, float* dbeta, int relu, void* stream);

/* InstanceNorm2d(affine=False, eps) + LeakyReLU(slope) (+ skip), networks/models.py:273-340.
 *   y = IN(x); if (act) y = leaky_relu(y); if (skip) y += skip     ; saves mean/invstd [N,C] */
int combat_instnorm_fwd(const void* x, int x_dtype, const void* skip, void* y, int dtype, int N, int HW, int C, float eps, float slope,
                        int act, float* save_mean, float* save_invstd, void* stream);
/* dy = dy1 (+ dy2); dyh = act ? dy*lrelu'(xhat) : dy; dx = invstd*(dyh - mean(dyh) - xhat*mean(dyh*xhat)) */
int combat_instnorm_bwd(const void* dy1, const void* dy2, const void* x, int x_dtype, void* dx, int dtype, int N, int HW, int C,
                        float slope, int act, const float* mean, const float* invstd, void* stream);
/* Large planes (H*W >= 4096 with too few (sample, channel-group) pairs to fill the GPU, e.g. the ImageNet-10 shape): the same
 * InstanceNorm forward / backward with every plane split over K CTAs (combat_instnorm_splits gives K; 1 = use the calls above).
 * part: N * K * 2 * C floats of scratch (chunk means / centred sums of squares, merged with Chan's formula; chunk gradient sums). */
int combat_instnorm_splits(int N, int HW, int C);
int combat_instnorm_fwd_split(const void* x, int x_dtype, const void* skip, void* y, int dtype, int N, int HW, int C, float eps,
                              float slope, int act, float* save_mean, float* save_invstd, float* part, int K, void* stream);
int combat_instnorm_bwd_split(const void* dy1, const void* dy2, const void* x, int x_dtype, void* dx, int dtype, int N, int HW, int C,
                              float slope, int act, const float* mean, const float* invstd, float* part, int K, void* stream);
/* t = leaky_relu(bilinear_up2x(x)) (align_corners=False), networks/models.py:274; slope==1 -> no activation */
int combat_upsample2x_act(const void* x, void* y, int dtype, int N, int H, int W, int C, float slope, void* stream);
int combat_upsample2x_act_bwd(const void* dy, const void* y, void* dx, int dtype, int N, int H, int W, int C, float slope,
                              void* stream);
/* y = leaky_relu(x) elementwise and its backward (used for conv0_0 output, networks/models.py:321) */
int combat_leaky_relu(const void* x, void* y, int dtype, long long n, float slope, void* stream);
int combat_leaky_relu_bwd(const void* dy, const void* x, void* dx, int dtype, long long n, float slope, void* stream);
/* dz = dy*(1-y^2) for y = tanh(z), NCHW float32 (networks/models.py:340) */
int combat_tanh_bwd(const float* dy, const float* y, float* dz, long long n, void* stream);
/* column sums over rows of a [R,C] matrix: out[c] (+)= sum_r x[r,c]  (conv bias gradient) */
int combat_colsum(const void* x, int dtype, long long R, int C, float* out, void* stream);
/* avg_pool2d(P) + flatten(NCHW order) + Linear (preact_resnet.py:99-101, resnet.py:95-97) */
int combat_pool_linear_fwd(const void* x, int dtype, int B, int Hf, int Wf, int C, int P, const float* W, const float* b,
                           int ncls, float* pooled, float* logits, void* stream);
int combat_pool_linear_bwd(const float* dlogits, const float* pooled, const float* W, int B, int Hf, int Wf, int C, int P,
                           int ncls, void* dx, int dtype, float* dW, float* db, void* stream);
/* max_pool2d(2) NHWC (defenses/frequency_based/model.py:21,32,43) */
int combat_maxpool2(const void* x, void* y, int dtype, int N, int H, int W, int C, void* stream);

/* ---------------------------------------------------------------- frequency-detector TRAINING (csrc/detector.cu)
 * defenses/frequency_based/train.py:178-221 with model.py:8-52 in train mode.  Next row of the scope table:
 * declared and built, not yet part of a GPU-validated path.
 *   combat_elu_bwd       nn.ELU backward from the kept OUTPUT a: dz = da * (a > 0 ? 1 : a + 1)           (model.py:14..39)
 *   combat_maxpool2_bwd  nn.MaxPool2d((2,2)) backward, NHWC; gradient to the first maximum of each window (model.py:21,32,43)
 *   combat_mask_scale    nn.Dropout(0.2) with a caller-provided keep mask (1 byte / element): y = keep ? x * scale : 0;
 *                        forward and backward are the same map                                        (model.py:22,33,44)
 *   combat_adadelta      torch.optim.Adadelta(lr = 0.05, weight_decay = 1e-4) over flat float32 buffers      (train.py:152)
 */
int combat_elu_bwd(const void* da, const void* a, void* dz, int dtype, long long n, void* stream);
int combat_maxpool2_bwd(const void* dy, const void* x, void* dx, int dtype, int N, int H, int W, int C, void* stream);
int combat_mask_scale(const void* x, const unsigned char* keep, void* y, int dtype, long long n, float scale, void* stream);
int combat_adadelta(float* p, const float* g, float* square_avg, float* acc_delta, long long n, const float* lr_dev, float rho,
                    float eps, float wd, void* stream);
/* "Grad L2 Loss" of train_generator.py:235-243 (a logged scalar, not part of the optimised loss): the sum of the two MSEs
 * between the vertical / horizontal difference images of F.pad(inputs, (1,1,2,1)) and F.pad(inputs_bd, (1,1,2,1)).
 * partial: rows*C*2 floats of scratch; out: 1 float. */
int combat_grad_l2(const float* x, const float* x_bd, float* partial, float* out, int rows, int C, int H, int W, void* stream);
/* total-variation loss of the imperceptible variant (train_generator_imperceptible.py:228; kornia 0.6.6 total_variation, mean
 * over the batch): out[0] = (1/rows) * sum over images of (sum |x[h+1]-x[h]| + sum |x[w+1]-x[w]|); when `grad` is given,
 * grad_weight * d(sum)/dx is ADDED to it (float32 NCHW like x; pass tv_weight / rows).  partial: >= rows*C floats of scratch. */
int combat_tv_loss(const float* x, float* grad, float grad_weight, float* partial, float* out, int rows, int C, int H, int W,
                   void* stream);

/* ---------------------------------------------------------------- PostTensorTransform (csrc/augment.cu)
 * utils/dataloader.py:45-60: kornia RandomCrop(padding) -> RandomRotation -> RandomHorizontalFlip, as one gather over NCHW
 * float32 images and its adjoint (the G-step differentiates through transforms(inputs_bd), train_generator.py:228,250).
 * params: rows x 8 floats per image, drawn on the host (combat_b200/utils/dataloader.py):
 *   [0] xs - pad, [1] ys - pad (integer-valued crop shifts; 0 when the whole-batch crop gate :17 is off),
 *   [2] cos(angle), [3] sin(angle), [4] rotation on (0/1), [5] horizontal flip (0/1), [6..7] unused.
 * out(n,c,y,x) = bilinear sample (zeros outside, align_corners=True pixel mapping) of the cropped image at the inverse
 * rotation of (flip ? W-1-x : x, y) about ((W-1)/2, (H-1)/2).  _bwd: din (+)= adjoint(dout); din is zero-filled first
 * unless accumulate != 0.
 */
int combat_post_transform_fwd(const float* in, float* out, const float* params, int rows, int C, int H, int W, void* stream);
int combat_post_transform_bwd(const float* dout, float* din, const float* params, int rows, int C, int H, int W,
                              int accumulate, void* stream);
/* ---------------------------------------------------------------- WaNet warp trigger (csrc/warp.cu)
 * train_generator_wanet.py:151-158 / :196-203: z = netG(inputs), the GridGenerator's tanh output [rows_of_x, 2, S, S]
 * (networks/models.py:383-385); noise_grid = bicubic upsample (align_corners=True) of flow to H x H, permuted to [.., H, W, 2];
 * grid = clamp(identity_grid * (1 - grid_rescale) + noise_grid * grid_rescale, -1, 1); out = grid_sample(x, grid) (bilinear, zeros
 * padding, align_corners=True).  ident: the H values of torch.linspace(-1, 1, H) (:560).  Square NCHW float32 images, S <= 4.
 * _fwd: row i of `out` is the warped image of sample perm[i] (i when perm is NULL) for i < num_bd (*num_bd_dev when given), a
 * plain copy of it otherwise (the C-step batch assembly, :159).  Optional per-row outputs: noise_grid [rows, H, W, 2],
 * sq_partial[i] = sum noise_grid^2 (loss_l2 = MSE(noise_grid, 0), :212), gl_partial[i] = this row's share of the logged
 * finite-difference term (:213-222; mean it over the rows).
 * _bwd: dz = d/dz of  <g1 + g2, out> + (l2_scale / 2) * |noise_grid|^2  (g2 may be NULL), [rows, 2, S, S].
 * combat_tanh_fwd: the tanh at the end of GridGenerator.forward (models.py:384), float32. */
int combat_tanh_fwd(const float* x, float* y, long long n, void* stream);
int combat_wanet_warp_fwd(const float* x, const float* z, const float* ident, const int* perm, int rows, int num_bd,
                          const int* num_bd_dev, float grid_rescale, float* out, float* noise_grid, float* sq_partial,
                          float* gl_partial, int C, int H, int W, int S, void* stream);
int combat_wanet_warp_bwd(const float* x, const float* z, const float* ident, const float* g1, const float* g2,
                          float grid_rescale, float l2_scale, float* dz, int rows, int C, int H, int W, int S, void* stream);
/* layout/dtype helpers */
int combat_nchw_to_nhwc(const float* x, void* y, int dtype, int N, int C, int H, int W, void* stream);
int combat_nhwc_to_nchw(const void* x, int dtype, float* y, int N, int C, int H, int W, void* stream);
int combat_onehot_planes(void* y, int dtype, const long long* labels, int N, int HW, int Ctot, int c_off, int ncls,
                         void* stream);
/* dst[pix, c_off + c] = leaky_relu(src[pix, c]): activated copy into a channel slice (networks/models.py:524-531) */
int combat_lrelu_into_slice(const void* src, void* dst, int dtype, long long npix, int Csrc, int Cdst, int c_off,
                            float slope, void* stream);

#ifdef __cplusplus
}
#endif
#endif
