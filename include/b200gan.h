/*
 * b200gan.h -- C ABI of libb200gan.so: the B200 (sm_100a) kernels behind the DCGAN adversarial
 * training step of harlanljones/gan-enhanced-pneumonia-classifier (src/dcgan.py + src/train_gan.py).
 *
 * The reference has no native code and no FFI: the boundary it offers for this path is the set of
 * PyTorch operator calls its nn.Modules and training loop make.  Each entry point below replaces one
 * of those operator calls (cited as reference file:line); a reference maintainer binds them with the
 * ctypes stub shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; every pointer inside a view is a DEVICE pointer owned by the caller;
 *   - all launches are asynchronous on the `stream` argument (a cudaStream_t passed as void*);
 *   - return value: 0 on success, a negative b200gan_status otherwise; the message of the last error
 *     of the calling thread is available from b200gan_last_error_string();
 *   - the library never allocates persistent device memory; workspace is passed in by the caller;
 *   - no CPU, cuDNN or Triton fallback exists: unsupported shapes return B200GAN_ERR_UNSUPPORTED.
 */
#ifndef B200GAN_H_
#define B200GAN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200GAN_VERSION 440   /* major*10000 + minor*100 + patch */

typedef enum b200gan_status {
  B200GAN_OK = 0,
  B200GAN_ERR_BAD_ARG = -1,
  B200GAN_ERR_UNSUPPORTED = -2,
  B200GAN_ERR_CUDA = -3,
  B200GAN_ERR_WORKSPACE = -4,
  B200GAN_ERR_NCCL = -5
} b200gan_status;

typedef enum b200gan_dtype { B200GAN_F32 = 0, B200GAN_BF16 = 1 } b200gan_dtype;

typedef enum b200gan_act {
  B200GAN_ACT_NONE = 0,
  B200GAN_ACT_RELU = 1,     /* dcgan.py:28,32,36,40,44 */
  B200GAN_ACT_LRELU = 2,    /* dcgan.py:66,70,74,78,82 (slope passed separately, 0.2 in the reference) */
  B200GAN_ACT_TANH = 3,     /* dcgan.py:47 */
  B200GAN_ACT_SIGMOID = 4   /* dcgan.py:85 */
} b200gan_act;

typedef enum b200gan_algo {
  B200GAN_ALGO_AUTO = 0,    /* tcgen05 implicit GEMM when the shape/dtype qualifies, SIMT otherwise */
  B200GAN_ALGO_SIMT = 1,    /* fp32-accumulating CUDA-core kernels (any shape, f32 or bf16 storage) */
  B200GAN_ALGO_TCGEN05 = 2  /* force the tensor-core path; B200GAN_ERR_UNSUPPORTED if it does not apply */
} b200gan_algo;

/* A strided 4-d activation view: logical element (n,h,w,c) lives at ptr + n*sn + h*sh + w*sw + c*sc
 * (strides in ELEMENTS).  The reference's NCHW tensors, the library's NHWC tensors and the zero-row
 * padded NHWC tensors of the tensor-core path are all expressed this way. */
typedef struct b200gan_view {
  void*   ptr;
  int32_t dtype;            /* b200gan_dtype */
  int32_t n, h, w, c;
  int64_t sn, sh, sw, sc;
} b200gan_view;

/* Convolution geometry shared by Conv2d and ConvTranspose2d: square kernel k, stride, padding.
 * `weight` is always the fp32 master tensor in the reference's own layout:
 *   Conv2d          (Cout, Cin, k, k)   dcgan.py:65,68,72,76,80,84
 *   ConvTranspose2d (Cin, Cout, k, k)   dcgan.py:26,30,34,38,42,46
 * `wpacked` is NULL or the bf16 repack produced by b200gan_pack_conv_weight for the tensor-core path. */
typedef struct b200gan_conv {
  int32_t k, stride, pad;
  int32_t algo;             /* b200gan_algo */
} b200gan_conv;

/* Optional fusions around one convolution call (NULL or all-zero = plain convolution).  They describe WHAT is computed,
 * not how: where a kernel cannot fuse an item the library runs the equivalent extra pass itself, so results never
 * depend on which kernel was selected.
 *
 *  out_act / out_slope   fprop:  y = act(conv(x)) for the layers without BatchNorm: LeakyReLU(0.2) after the first
 *                        Conv2d (dcgan.py:65-66), Tanh after the last ConvTranspose2d (dcgan.py:46-47).
 *  dy_act / dy_slope / dy_ref   dgrad, wgrad: the gradient operand `dy` is the gradient w.r.t. the OUTPUT of the activation
 *                        that follows this convolution; dy_eff = dy * act'(.) is formed while dy is read, with act' taken
 *                        from dy_ref = the saved activation output (same extents as dy; LeakyReLU/ReLU: its sign,
 *                        Tanh: 1 - a^2, Sigmoid: a(1-a)).  Replaces leaky_relu_backward / tanh_backward of autograd.
 *  bn_sums               fprop:  per-channel batch statistics of the stored result, OVERWRITTEN: bn_sums[0..C) = sum y,
 *                        bn_sums[C..2C) = sum y^2 over (N,H,W) -- the same contract as b200gan_bn_stats, which it replaces
 *                        (native_batch_norm's statistics pass, dcgan.py:27,31,35,39,43,69,73,77,81).
 *  prev_*                dgrad:  the result dx is the gradient w.r.t. a_prev = act(BN(y_prev)), the previous layer's
 *                        output; the epilogue applies that activation's backward and the BatchNorm-backward reductions:
 *                          dz = dx * act'(prev_scale[c]*y_prev + prev_shift[c])          (stored INSTEAD of dx)
 *                          prev_sums[0..C) = sum dz,  prev_sums[C..2C) = sum dz*(y_prev - prev_mean[c])*prev_invstd[c]
 *                        (OVERWRITTEN; the same contract as b200gan_bn_act_bwd_reduce, which it replaces).  The caller then
 *                        finishes native_batch_norm_backward with b200gan_bn_act_bwd_apply(da = dz, act = NONE).
 *                        With prev_scale == NULL there is no BatchNorm between the two convolutions (dcgan.py:65-68): prev_y is
 *                        then the saved ACTIVATION OUTPUT a_prev, dz = dx * act'(.) taken from its sign, and prev_shift /
 *                        prev_mean / prev_invstd / prev_sums are unused. */
typedef struct b200gan_fuse {
  int32_t out_act;
  float   out_slope;
  int32_t dy_act;
  float   dy_slope;
  const b200gan_view* dy_ref;
  double* bn_sums;
  int32_t prev_act;
  float   prev_slope;
  const b200gan_view* prev_y;
  const float* prev_scale;
  const float* prev_shift;
  const float* prev_mean;
  const float* prev_invstd;
  double* prev_sums;
} b200gan_fuse;

int         b200gan_version(void);
const char* b200gan_last_error_string(void);
/* Fills name (<=255 chars + NUL) with the device name; returns SM count, or a negative status. */
int         b200gan_device_info(int device, char* name, int* cc_major, int* cc_minor);

/* ---- nn.Conv2d: forward (dcgan.py:65-84), input gradient and weight gradient (autograd of it,
 *      reached from train_gan.py:129,137,148).  dw is fp32 (Cout,Cin,k,k) and is ACCUMULATED into
 *      (`+=`), matching autograd's accumulation over the real and fake passes (train_gan.py:129,137).
 *      `workspace` (wgrad only): NULL, or numel(dweight) floats owned by the caller that are ALL ZERO on entry and are handed
 *      back all zero: the tensor-core kernel reduces its split-K partial sums there in a layout that makes every atomic a
 *      coalesced 128-byte transaction and then transposes into dweight; with NULL it adds into dweight directly (slower). */
int b200gan_conv2d_fprop(const b200gan_conv* cv, const b200gan_view* x, const float* weight, const void* wpacked,
                         const b200gan_view* y, const b200gan_fuse* fuse, void* stream);
int b200gan_conv2d_dgrad(const b200gan_conv* cv, const b200gan_view* dy, const float* weight, const void* wpacked,
                         const b200gan_view* dx, const b200gan_fuse* fuse, void* stream);
int b200gan_conv2d_wgrad(const b200gan_conv* cv, const b200gan_view* x, const b200gan_view* dy, float* dweight,
                         float* workspace, const b200gan_fuse* fuse, void* stream);

/* Number of floats of `workspace` the weight-gradient call for these operands can use (0: the kernel that will be selected takes none).
 * x / dy as passed to b200gan_conv2d_wgrad (transposed = 0) or b200gan_convT2d_wgrad (transposed = 1).  Negative status on bad arguments. */
int64_t b200gan_conv_wgrad_workspace_floats(const b200gan_conv* cv, const b200gan_view* x, const b200gan_view* dy, int32_t transposed);

/* ---- nn.ConvTranspose2d: forward (dcgan.py:26-46), input gradient, weight gradient (Cin,Cout,k,k). */
int b200gan_convT2d_fprop(const b200gan_conv* cv, const b200gan_view* x, const float* weight, const void* wpacked,
                          const b200gan_view* y, const b200gan_fuse* fuse, void* stream);
int b200gan_convT2d_dgrad(const b200gan_conv* cv, const b200gan_view* dy, const float* weight, const void* wpacked,
                          const b200gan_view* dx, const b200gan_fuse* fuse, void* stream);
int b200gan_convT2d_wgrad(const b200gan_conv* cv, const b200gan_view* x, const b200gan_view* dy, float* dweight,
                          float* workspace, const b200gan_fuse* fuse, void* stream);

/* ---- bf16 repack of a k=4 conv weight for the tcgen05 implicit GEMM (`wpacked` above).  `weight` is the fp32
 *      master in conv geometry (Co,Ci,4,4) -- i.e. the Conv2d weight, or the ConvTranspose2d weight (Cin,Cout,4,4)
 *      read as (Co=Cin, Ci=Cout).  form 0 ("down": Conv2d fprop / ConvTranspose2d dgrad): [Co][(kh,kw,Ci)];
 *      form 1 ("up": ConvTranspose2d fprop / Conv2d dgrad): [parity class][Ci][(jh,jw,Co)].  out: Co*Ci*16 bf16.
 *      form 2: both in one launch, form 0 at out[0 .. Co*Ci*16), form 1 behind it (out: 2*Co*Ci*16 bf16). */
int b200gan_pack_conv_weight(const float* weight, int32_t co, int32_t ci, int32_t k, int32_t form, void* out, void* stream);

/* ---- nn.BatchNorm2d in training mode (dcgan.py:27,31,35,39,43,69,73,77,81) -------------------------
 * bn_stats:     sums[0..C) = sum y, sums[C..2C) = sum y^2 over (N,H,W)   (double accumulators, overwritten)
 * bn_finalize:  mean / biased var -> scale = gamma*invstd, shift = beta - mean*scale, saves mean and
 *               invstd for backward, updates running_mean/var (momentum, UNBIASED var) and
 *               num_batches_tracked (int64) exactly like torch.  running_* / num_batches_tracked may be NULL
 *               (no tracking).  Under synchronised BatchNorm the caller all-reduces `sums` between the two. */
int b200gan_bn_stats(const b200gan_view* y, double* sums, void* stream);
int b200gan_bn_finalize(double* sums, int32_t channels, int64_t count, const float* gamma, const float* beta,
                        float* running_mean, float* running_var, int64_t* num_batches_tracked,
                        float momentum, float eps, float* scale, float* shift, float* save_mean,
                        float* save_invstd, void* stream);
/* eval mode (generate_synthetic.py:34): scale/shift from the running statistics. */
int b200gan_bn_eval_coeffs(int32_t channels, const float* gamma, const float* beta, const float* running_mean,
                           const float* running_var, float eps, float* scale, float* shift, void* stream);

/* a = act(y*scale[c] + shift[c]); scale == NULL means no BatchNorm (dcgan.py:66 first D layer, :47 tanh, :85). */
int b200gan_bn_act_fwd(const b200gan_view* y, const float* scale, const float* shift, int32_t act, float slope,
                       const b200gan_view* a, void* stream);

/* b200gan_bn_finalize followed by b200gan_bn_act_fwd on the layer's tensor, as ONE launch where the tensors are dense NHWC of one dtype
 * (otherwise the library runs the two passes): the training-mode forward of BatchNorm2d + activation (dcgan.py:31-32, :72-73) behind a
 * convolution that accumulated `sums` in its epilogue (b200gan_fuse.bn_sums).  Coefficients are bit-identical to b200gan_bn_finalize's. */
int b200gan_bn_finalize_act_fwd(double* sums, int32_t channels, int64_t count, const float* gamma, const float* beta,
                                float* running_mean, float* running_var, int64_t* num_batches_tracked,
                                float momentum, float eps, float* scale, float* shift, float* save_mean,
                                float* save_invstd, const b200gan_view* y, int32_t act, float slope,
                                const b200gan_view* a, void* stream);

/* Backward of act(BN(y)).  da: gradient w.r.t. the activation output; y: the saved conv output;
 * a: the saved activation output (only read for TANH / SIGMOID, may be NULL otherwise).
 * bwd_reduce: sums[0..C) = sum dz, sums[C..2C) = sum dz*xhat with dz = da*act'(.)     (native_batch_norm_backward)
 * bwd_apply:  dy = gamma*invstd*(dz - sum dz/count - xhat * sum(dz*xhat)/count); dgamma += sum dz*xhat,
 *             dbeta += sum dz (written by the first CTA); with scale == NULL: dy = dz (no BatchNorm). */
int b200gan_bn_act_bwd_reduce(const b200gan_view* da, const b200gan_view* y, const b200gan_view* a,
                              const float* scale, const float* shift, const float* save_mean,
                              const float* save_invstd, int32_t act, float slope, double* sums, void* stream);
int b200gan_bn_act_bwd_apply(const b200gan_view* da, const b200gan_view* y, const b200gan_view* a,
                             const float* scale, const float* shift, const float* save_mean,
                             const float* save_invstd, const float* gamma, double* sums, int64_t count,
                             int32_t act, float slope, const b200gan_view* dy, float* dgamma, float* dbeta,
                             void* stream);

/* ---- Sigmoid (dcgan.py:85) + nn.BCELoss(mean) against a constant target (train_gan.py:90,92-93,
 *      125-128,134-136,145-147) and their backward, in one pass over the B logits:
 *   prob[i]   = 1/(1+exp(-logit[i]))
 *   out[0]    = mean_i -(t*max(log p,-100) + (1-t)*max(log1p(-p),-100))      (the loss)
 *   out[1]    = mean_i prob[i]                                               (D_x / D_G_z, train_gan.py:130,138,149)
 *   dlogit[i] = grad_scale * (p-t)/max(p(1-p),1e-12)/B * p(1-p)              (dlogit may be NULL) */
int b200gan_bce_sigmoid(const float* logit, int32_t batch, float target, float grad_scale, float* prob,
                        float* out2, float* dlogit, void* stream);

/* ---- torch.optim.Adam (train_gan.py:94-95,141,150) over one flat fp32 arena: lr, betas, eps, no weight
 *      decay, no amsgrad; `step` is the 1-based step count of this update.  grad_scale multiplies the
 *      gradient first (1/world_size for data-parallel sums).  Hyper-parameters are doubles, as torch holds them
 *      (Python floats): `1 - beta2` is formed in double and rounded once, exactly like torch's scalar handling.
 *      step_dev (optional): DEVICE address of the int64 step count; when given it overrides `step`, so that the launch can be
 *      replayed from a CUDA graph while the count advances on the device. */
int b200gan_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, double lr,
                 double beta1, double beta2, double eps, int32_t step, const int64_t* step_dev, float grad_scale, void* stream);

/* ---- layout plumbing: dst(n,h,w,c) = (dst dtype) src(n,h,w,c) for two views of equal extents. */
int b200gan_copy_view(const b200gan_view* src, const b200gan_view* dst, void* stream);
int b200gan_fill_f32(float* ptr, int64_t numel, float value, void* stream);

/* ---- input pipeline of the training loop (src/data_loader.py:17-23 'train' transform after Resize; the batch consumed at
 *      src/train_gan.py:123): one batch gathered from a DEVICE-resident uint8 image cache instead of PNG decoding per epoch:
 *        out[b] = Normalize(mean, std)(ToTensor(hflip_if(flip[b])(cache[index[b]])))
 *      cache: (num_images, C, H, W) uint8 contiguous, C/H/W = out's extents;  index: int64[out->n] (device);  flip: uint8[out->n]
 *      (device; non-zero mirrors along W; NULL = never);  mean/std: float[C] HOST arrays (NULL = 0 / 1);  out: f32 or bf16 view.
 *      Arithmetic is torchvision's, (v / 255 - mean) / std in fp32 with true divisions: bit-exact for an f32 output. */
int b200gan_gather_augment(const uint8_t* cache, int64_t num_images, const int64_t* index, const uint8_t* flip, const float* mean,
                           const float* std, const b200gan_view* out, void* stream);

/* ---- WGAN-GP critic update (src/wggan.py:72-89 `gradient_penalty`, src/train_wggan.py:70-85).  The reference
 *      leaves the penalty's double backward to torch.autograd (`create_graph=True`); here it is launched explicitly.  Its convolutions are
 *      plain b200gan_conv2d_{fprop,dgrad,wgrad} calls (the reverse of an input-gradient convolution is a forward convolution of the
 *      adjoint); these entry points are the rest:
 *        sample_sumsq     sumsq[n] = sum over (h,w,c) of x(n,.)^2                    (`gradients.view(B,-1).norm(2, dim=1)`, wggan.py:87-88)
 *        gp_from_norms    gp[0] = lambda * mean_n (sqrt(sumsq[n]) - 1)^2  and  coeff[n] = lambda * (2/B) (||g_n|| - 1) / ||g_n||,
 *                         so that d gp / d g(n,.) = coeff[n] * g(n,.)
 *        sample_axpby     out(n,.) = a[n] x(n,.) + b[n] y(n,.)   (y NULL: first term only; a / b NULL: 1) -- the interpolation
 *                         `alpha*real + (1-alpha)*fake` (wggan.py:77), the scaling by coeff[n], and `dst += src` (out == x)
 *        bn_bwd_bwd       second-order terms of training-mode BatchNorm.  With xhat = (y-mean) invstd, P(v) = v - mean(v) - xhat mean(v xhat),
 *                         the first backward computed dy = gamma invstd P(dz); given r = adjoint of dy:
 *                           u   = gamma invstd P(r) * act'(scale*y + shift)      adjoint of the activation gradient above this BatchNorm
 *                           inj = -gamma invstd^2 [xhat mean(r P(dz)) + mean(dz xhat) P(r) + mean(r xhat) P(dz)]   adjoint of the conv output y
 *                           dgamma[c] += invstd[c] * sum r P(dz)
 *                         dz_sums: [0..C) = sum dz, [C..2C) = sum dz*xhat of the first backward (b200gan_fuse.prev_sums' contract);
 *                         workspace: 3*C doubles.  (native_batch_norm_backward's own backward in torch.autograd.)
 *        mean_f32         out[0] = scale * sum x[0..n)   (critic scores -> loss terms, train_wggan.py:74,79) */
int b200gan_sample_sumsq(const b200gan_view* x, double* sumsq, void* stream);
int b200gan_gp_from_norms(const double* sumsq, int32_t batch, float lambda_gp, float* gp, float* coeff, void* stream);
int b200gan_sample_axpby(const b200gan_view* x, const float* a, const b200gan_view* y, const float* b, const b200gan_view* out, void* stream);
int b200gan_bn_bwd_bwd(const b200gan_view* r, const b200gan_view* y, const b200gan_view* dz, const float* scale, const float* shift,
                       const float* save_mean, const float* save_invstd, const float* gamma, const double* dz_sums, int64_t count,
                       int32_t act, float slope, const b200gan_view* u, const b200gan_view* inj, float* dgamma, double* workspace,
                       void* stream);
int b200gan_mean_f32(const float* x, int64_t n, float scale, float* out, void* stream);

/* ---- conditional GAN (src/cgan.py): what surrounds its convolutions.
 *        embed_add        out[n][0..dim) = table[labels[n]] + z[n] (z NULL: lookup only), out[n][dim..dim+tail) = 1.
 *                         Generator conditioning `z + label_emb(labels)` (cgan.py:22,55-56); tail = 1 appends the constant feature through
 *                         which the bias of `fc` (cgan.py:24,57) rides in the latent GEMM (b200gan_convT2d_* with k=7 on the 1x1 input).
 *        embed_bwd        dtable[c] += sum over {n : labels[n] == c} of dx[n][0..dim)   (rows of dx `stride` floats apart; nn.Embedding backward)
 *        upconv3_fold     nearest Upsample(2) + Conv2d(3,1,1) (cgan.py:28-29 ... 48-49) == ConvTranspose2d(4,2,1) with
 *                         w4[ci][co][kh][kw] = sum_ab A[kh][a] A[kw][b] w3[co][ci][a][b], A = [[0,0,1],[0,1,1],[1,1,0],[1,0,0]]:
 *                         the folded weight feeds b200gan_convT2d_{fprop,dgrad,wgrad} (no upsampled tensor, 4/9 of the multiply-adds);
 *        upconv3_unfold   the adjoint: dw3 += A^T dw4 A  (weight gradient back in the Conv2d(3) layout)
 *        class_proj_fwd   out[n] += < table[labels[n]], x(n) flattened in (c,h,w) order >     (projection term, cgan.py:103)
 *        class_proj_bwd   dx(n) += dout[n] table[labels[n]] (dx NULL: skipped);  dtable[c] += sum over {n : labels[n] == c} of dout[n] x(n)
 *                         (dtable NULL: skipped).  labels: int64 on the device.  Batch-order loops, no atomics: run-to-run identical.
 *        bce_logits       nn.BCEWithLogitsLoss(mean) against PER-SAMPLE targets (train_cgan.py:111,156-160) with its backward and the mean probability
 *                         (`torch.sigmoid(out).mean()`, :163,170,182):  out2[0] = mean_i max(x,0) - x t + log1p(exp(-|x|)),
 *                         out2[1] = mean_i sigmoid(x_i),  dlogit[i] = grad_scale (sigmoid(x_i) - t_i) / B   (dlogit may be NULL)
 *        fm_pair          feature matching on one pair of intermediates (train_cgan.py:75-76): sum[0] += sum (real - fake)^2 (fp64, the caller
 *                         divides by numel), dfake = coeff (real - fake) written (add = 0) or added (add = 1); dfake may be NULL
 *        accumulate_2d    dst[r][c] += src[r*src_row_stride + c*src_col_stride], src float (src_f64 = 0) or double: bias gradients out of the
 *                         fp64 channel sums of b200gan_bn_stats, the Linear's (out,in) gradient out of the latent GEMM's (in,out) one */
int b200gan_bce_logits(const float* logit, const float* target, int32_t batch, float grad_scale, float* out2, float* dlogit, void* stream);
int b200gan_fm_pair(const b200gan_view* real, const b200gan_view* fake, const b200gan_view* dfake, float coeff, int32_t add, double* sum, void* stream);
int b200gan_accumulate_2d(float* dst, const void* src, int32_t src_f64, int32_t rows, int32_t cols, int64_t src_row_stride, int64_t src_col_stride,
                          void* stream);
int b200gan_embed_add(const float* table, const int64_t* labels, const float* z, int32_t batch, int32_t dim, int32_t tail, float* out, void* stream);
int b200gan_embed_bwd(const float* dx, const int64_t* labels, int32_t batch, int32_t dim, int32_t stride, int32_t num_classes, float* dtable,
                      void* stream);
int b200gan_upconv3_fold(const float* w3, int32_t co, int32_t ci, float* w4, void* stream);
int b200gan_upconv3_unfold(const float* dw4, int32_t co, int32_t ci, float* dw3, void* stream);
int b200gan_class_proj_fwd(const b200gan_view* x, const float* table, const int64_t* labels, float* out, void* stream);
int b200gan_class_proj_bwd(const b200gan_view* x, const float* table, const int64_t* labels, const float* dout, const b200gan_view* dx,
                           int32_t num_classes, float* dtable, void* stream);

/* ---- VGG16 perceptual loss of the conditional GAN (src/train_cgan.py:57-73,186: MSE between torchvision vgg16.features[:4], [4:9], [9:16] of the
 *      fake and the real batch; the network is frozen, only the gradient w.r.t. the fake image is needed).  Its Conv2d(3,1,1) layers run on the
 *      stride-2 4x4 convolution kernels: T[n,i,j,(a,b,co)] = conv_k4s2p1(X, W4) with W4[(a,b,co),ci,kh,kw] = w3[co,ci,kh-a,kw-b] (zero outside
 *      0..2) holds output pixel (2i+a, 2j+b) of the 3x3 convolution in channel block (a,b) -- all four output parities as 4*Co channels of one call
 *      of b200gan_conv2d_fprop / b200gan_conv2d_dgrad (16/9 of the multiply-adds, tensor-core speed).  The entry points around it:
 *        conv3x3_fold    w4 (4*Co, ci_pad, 4, 4) from w3 (Co, Ci, 3, 3); input channels Ci..ci_pad-1 are zero (the 3-channel image is stored 32 wide)
 *        bias_relu_d2s   a[n,2i+a,2j+b,co] = relu(t[n,i,j,(a,b,co)] + bias[co])                 bias + ReLU + depth-to-space, one pass
 *        relu_bwd_s2d    dt[n,i,j,(a,b,co)] = da[n,2i+a,2j+b,co] * (a[n,2i+a,2j+b,co] > 0)      ReLU backward + space-to-depth, one pass
 *        maxpool2_fwd    nn.MaxPool2d(2,2)
 *        maxpool2_bwd    da (+)= dp routed to the FIRST maximum of each window in (row, column) order (ATen's choice); add = 1 accumulates
 *      Operands: dense NHWC views of one dtype (f32 / bf16), channels a multiple of 4 / 8.  The MSE itself is b200gan_fm_pair. */
int b200gan_conv3x3_fold(const float* w3, int32_t co, int32_t ci, int32_t ci_pad, float* w4, void* stream);
int b200gan_bias_relu_d2s(const b200gan_view* t, const float* bias, const b200gan_view* a, void* stream);
int b200gan_relu_bwd_s2d(const b200gan_view* da, const b200gan_view* a, const b200gan_view* dt, void* stream);
int b200gan_maxpool2_fwd(const b200gan_view* a, const b200gan_view* p, void* stream);
int b200gan_maxpool2_bwd(const b200gan_view* a, const b200gan_view* dp, const b200gan_view* da, int32_t add, void* stream);

/* ---- data-parallel gradient-bucket layer (new functionality: the reference is single-device, src/train_gan.py:49; semantics in
 *      SURVEY.md section 8e).  One process per GPU; weights and Adam state replicated; every optimizer update (train_gan.py:141,150)
 *      is preceded by a SUM of the per-rank gradients, issued bucket by bucket while the backward pass is still running:
 *        dp_unique_id         rank 0 obtains the 128-byte NCCL unique id; the host side broadcasts it to the other ranks
 *                             (torch.distributed store / broadcast -- plumbing only);
 *        dp_init              ncclCommInitRank on the CURRENT device + a dedicated communication stream; returns an opaque handle
 *                             (the only persistent state the library ever owns; released by dp_destroy);
 *        dp_allreduce_bucket  in-place sum over ranks of `grad[0..numel)` (fp32, device).  Asynchronous: forks the communication
 *                             stream from `stream` (everything launched on `stream` so far precedes the collective) and
 *                             returns; `stream` itself continues with the backward pass.  Capturable into a CUDA graph;
 *        dp_sync              joins: `stream` waits for every bucket issued so far (call before b200gan_adam, whose grad_scale
 *                             carries the 1/world factor);
 *        dp_allreduce_f64     in-place sum over ranks of `buf[0..numel)` (fp64, device), COMPLETE for `stream` on return of the stream order:
 *                             forks to the communication stream and joins back.  Synchronised BatchNorm (`--sync-bn`): the per-channel
 *                             sums of b200gan_bn_stats / b200gan_fuse.bn_sums / prev_sums are summed over ranks between the reduction and
 *                             b200gan_bn_finalize / b200gan_bn_act_bwd_apply, whose `count` then is the GLOBAL sample count;
 *        dp_collectives       number of all-reduces issued through the handle (launch accounting / tests).
 *      Errors: B200GAN_ERR_NCCL with the NCCL message in b200gan_last_error_string().  NCCL is bound at run time from the
 *      libnccl.so.2 already loaded in the process (torch's), so single-GPU users never need it. */
#define B200GAN_DP_ID_BYTES 128
typedef struct b200gan_dp b200gan_dp;
int     b200gan_dp_unique_id(void* id_out);
int     b200gan_dp_init(const void* id, int32_t world, int32_t rank, b200gan_dp** out);
int     b200gan_dp_allreduce_bucket(b200gan_dp* dp, float* grad, int64_t numel, void* stream);
int     b200gan_dp_allreduce_f64(b200gan_dp* dp, double* buf, int64_t numel, void* stream);
int     b200gan_dp_sync(b200gan_dp* dp, void* stream);
int64_t b200gan_dp_collectives(const b200gan_dp* dp);
int     b200gan_dp_destroy(b200gan_dp* dp);

#ifdef __cplusplus
}
#endif
#endif  /* B200GAN_H_ */
