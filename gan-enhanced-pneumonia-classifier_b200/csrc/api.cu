// extern "C" surface of libb200gan.so (see include/b200gan.h): argument checking, algorithm dispatch
// (tcgen05 implicit GEMM vs SIMT), error strings.  No torch types cross this boundary.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace b200gan {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int v = cached[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return B200GAN_ERR_CUDA;
}

// SIMT kernels (conv_simt.cu)
int simt_conv_fprop(const b200gan_conv*, const b200gan_view* x, const float* w, const b200gan_view* y, const ConvFuse&, cudaStream_t);
int simt_conv_dgrad(const b200gan_conv*, const b200gan_view* dy, const float* w, const b200gan_view* dx, const ConvFuse&, cudaStream_t);
int simt_conv_wgrad(const b200gan_conv*, const b200gan_view* x, const b200gan_view* dy, float* dw, const ConvFuse&, bool grad_is_coarse,
                    cudaStream_t);
// tensor-core kernels (conv_tc.cu): return 1 when the problem does not qualify (caller falls back / errors)
int tc_conv_fprop(const b200gan_conv*, const b200gan_view* x, const void* wpacked, const b200gan_view* y, const TcEpi&, cudaStream_t);
int tc_conv_dgrad(const b200gan_conv*, const b200gan_view* dy, const void* wpacked, const b200gan_view* dx, const TcEpi&, cudaStream_t);
int tc_conv_wgrad(const b200gan_conv*, const b200gan_view* x, const b200gan_view* dy, float* dw, float* workspace, cudaStream_t);
int tc_pack_weight(const float* w, int Co, int Ci, int k, int form, void* out, cudaStream_t);
// image-side layers on warp-level MMAs (conv_thin_mma.cu): same return convention
int thin_down(const b200gan_view* fine, const b200gan_view* fine_ref, int fine_act, const float* w, const b200gan_view* coarse, int out_act,
              float slope, const TcEpi& epi, cudaStream_t);
int thin_up(const b200gan_view* coarse, const b200gan_view* coarse_ref, int coarse_act, float slope, const float* w, const b200gan_view* fine,
            int out_act, cudaStream_t);
int thin_wgrad(const b200gan_view* fine, const b200gan_view* fine_ref, int fine_act, const b200gan_view* coarse, const b200gan_view* coarse_ref,
               int coarse_act, float slope, float* dw, cudaStream_t);
// image-side layers whose feature side is not 32 channels wide (conv_edge.cu: the WGAN-GP 64-channel edge layers): same return convention
int edge_down(const b200gan_view* fine, const b200gan_view* fine_ref, int fine_act, const float* w, const b200gan_view* coarse, int out_act, float slope,
              cudaStream_t);
int edge_up(const b200gan_view* coarse, const b200gan_view* coarse_ref, int coarse_act, float slope, const float* w, const b200gan_view* fine, int out_act,
            cudaStream_t);
int edge_wgrad(const b200gan_view* fine, const b200gan_view* fine_ref, int fine_act, const b200gan_view* coarse, const b200gan_view* coarse_ref,
               int coarse_act, float slope, float* dw, cudaStream_t);
// latent GEMM and 7x7 GEMV (conv_thin.cu): same return convention
int window_fprop(const b200gan_conv*, const b200gan_view* x, const float* w, const b200gan_view* y, cudaStream_t);
int window_dgrad(const b200gan_conv*, const b200gan_view* dy, const float* w, const b200gan_view* dx, const TcEpi& epi, cudaStream_t);
int window_wgrad(const b200gan_conv*, const b200gan_view* x, const b200gan_view* dy, float* dw, cudaStream_t);
// valid convolution of a small map to one channel (the WGAN-GP critic's score map, conv_thin.cu): same return convention
int score_fprop(const b200gan_conv*, const b200gan_view* x, const float* w, const b200gan_view* y, cudaStream_t);
int score_dgrad(const b200gan_conv*, const b200gan_view* dy, const float* w, const b200gan_view* dx, cudaStream_t);
int score_wgrad(const b200gan_conv*, const b200gan_view* x, const b200gan_view* dy, float* dw, cudaStream_t);
int latent_fprop(const b200gan_conv*, const b200gan_view* z, const float* w, const b200gan_view* y, cudaStream_t);
int latent_wgrad(const b200gan_conv*, const b200gan_view* dy_fine, const b200gan_view* z, float* dw, cudaStream_t);
// the same two as warp-level tensor-core GEMMs (latent_mma.cu)
int latent_fprop_mma(const b200gan_conv*, const b200gan_view* z, const float* w, const b200gan_view* y, cudaStream_t);
int latent_wgrad_mma(const b200gan_conv*, const b200gan_view* dy_fine, const b200gan_view* z, float* dw, cudaStream_t);
// elementwise.cu
int ew_bn_stats(const b200gan_view*, double*, cudaStream_t);
int ew_bn_bwd_reduce(const b200gan_view*, const b200gan_view*, const b200gan_view*, const float*, const float*, const float*,
                     const float*, int, float, double*, const b200gan_view* dz_out, cudaStream_t);
int ew_bn_finalize(double*, int, int64_t, const float*, const float*, float*, float*, int64_t*, float, float, float*, float*,
                   float*, float*, cudaStream_t);
int ew_bn_eval_coeffs(int, const float*, const float*, const float*, const float*, float, float*, float*, cudaStream_t);
int ew_bn_act_fwd(const b200gan_view*, const float*, const float*, int, float, const b200gan_view*, cudaStream_t);
int ew_bn_finalize_act_fwd(double*, int, int64_t, const float*, const float*, float*, float*, int64_t*, float, float, float*, float*, float*, float*,
                           const b200gan_view*, int, float, const b200gan_view*, cudaStream_t);
int ew_bn_act_bwd_apply(const b200gan_view*, const b200gan_view*, const b200gan_view*, const float*, const float*, const float*,
                        const float*, const float*, double*, int64_t, int, float, const b200gan_view*, float*, float*,
                        cudaStream_t);
int ew_act_bwd_inplace(const b200gan_view* d, const b200gan_view* y, const float* scale, const float* shift, int act, float slope, cudaStream_t);
int ew_bce_sigmoid(const float*, int, float, float, float*, float*, float*, cudaStream_t);
int ew_adam(float*, const float*, float*, float*, int64_t, double, double, double, double, int, const int64_t*, float, cudaStream_t);
int ew_copy_view(const b200gan_view*, const b200gan_view*, cudaStream_t);
int ew_fill(float*, int64_t, float, cudaStream_t);
int ew_gather_augment(const uint8_t*, int64_t, const int64_t*, const uint8_t*, const float*, const float*, const b200gan_view*, cudaStream_t);

// wgan_gp.cu
int gp_bn_bwd_bwd(const b200gan_view* r, const b200gan_view* y, const b200gan_view* dz, const float* scale, const float* shift, const float* mean,
                  const float* invstd, const float* gamma, const double* dz_sums, int64_t count, int act, float slope, const b200gan_view* u,
                  const b200gan_view* inj, float* dgamma, double* sums3, cudaStream_t st);
int gp_sample_sumsq(const b200gan_view* x, double* out, cudaStream_t st);
int gp_from_norms(const double* sumsq, int n, float lambda, float* gp, float* coeff, cudaStream_t st);
int gp_sample_axpby(const b200gan_view* x, const float* a, const b200gan_view* y, const float* b, const b200gan_view* out, cudaStream_t st);
int gp_mean_f32(const float* x, int64_t n, float scale, float* out, cudaStream_t st);

// cgan_ops.cu
int cgan_embed_add(const float* table, const int64_t* labels, const float* z, int batch, int dim, int tail, float* out, cudaStream_t st);
int cgan_embed_bwd(const float* dx, const int64_t* labels, int batch, int dim, int stride, int classes, float* dtable, cudaStream_t st);
int cgan_upconv3_fold(const float* w3, int co, int ci, float* w4, cudaStream_t st);
int cgan_upconv3_unfold(const float* dw4, int co, int ci, float* dw3, cudaStream_t st);
int cgan_class_proj_fwd(const b200gan_view* x, const float* table, const int64_t* labels, float* out, cudaStream_t st);
int cgan_class_proj_bwd(const b200gan_view* x, const float* table, const int64_t* labels, const float* dout, const b200gan_view* dx, int classes,
                        float* dtable, cudaStream_t st);

int cgan_bce_logits(const float* x, const float* t, int batch, float grad_scale, float* out2, float* dlogit, cudaStream_t st);
int cgan_fm_pair(const b200gan_view* r, const b200gan_view* f, const b200gan_view* d, float coeff, int add, double* sum, cudaStream_t st);
int cgan_accumulate_2d(float* dst, const void* src, int f64, int rows, int cols, int64_t srs, int64_t scs, cudaStream_t st);

// vgg_ops.cu
int vgg_conv3x3_fold(const float* w3, int co, int ci, int ci_pad, float* w4, cudaStream_t st);
int vgg_bias_relu_d2s(const b200gan_view* t, const float* bias, const b200gan_view* a, cudaStream_t st);
int vgg_relu_bwd_s2d(const b200gan_view* da, const b200gan_view* a, const b200gan_view* dt, cudaStream_t st);
int vgg_maxpool2_fwd(const b200gan_view* a, const b200gan_view* p, cudaStream_t st);
int vgg_maxpool2_bwd(const b200gan_view* a, const b200gan_view* dp, const b200gan_view* da, int add, cudaStream_t st);

static int check_conv(const b200gan_conv* cv) {
  if (!cv) { set_error("null conv descriptor"); return B200GAN_ERR_BAD_ARG; }
  if (cv->k <= 0 || cv->stride <= 0 || cv->pad < 0 || cv->k > 16) { set_error("bad conv geometry k=%d s=%d p=%d", cv->k, cv->stride, cv->pad); return B200GAN_ERR_BAD_ARG; }
  return 0;
}

// conv geometry check: fine side (N,H,W,Ci) vs coarse side (N,OH,OW,Co)
static int check_pair(const b200gan_conv* cv, const b200gan_view* fine, const b200gan_view* coarse, const char* what) {
  const int oh = (fine->h + 2 * cv->pad - cv->k) / cv->stride + 1, ow = (fine->w + 2 * cv->pad - cv->k) / cv->stride + 1;
  if (fine->n != coarse->n || coarse->h != oh || coarse->w != ow || fine->h + 2 * cv->pad < cv->k) {
    set_error("%s: shapes do not match the convolution: fine (%d,%d,%d,%d) coarse (%d,%d,%d,%d) k=%d s=%d p=%d", what, fine->n,
              fine->h, fine->w, fine->c, coarse->n, coarse->h, coarse->w, coarse->c, cv->k, cv->stride, cv->pad);
    return B200GAN_ERR_BAD_ARG;
  }
  // the transposed direction must reproduce the fine extent exactly (no output_padding in the reference)
  if ((coarse->h - 1) * cv->stride - 2 * cv->pad + cv->k != fine->h || (coarse->w - 1) * cv->stride - 2 * cv->pad + cv->k != fine->w) {
    set_error("%s: (H + 2p - k) must be divisible by the stride (fine %dx%d, k=%d s=%d p=%d)", what, fine->h, fine->w, cv->k, cv->stride, cv->pad);
    return B200GAN_ERR_UNSUPPORTED;
  }
  return 0;
}

enum Prim { FPROP, DGRAD, WGRAD };

static bool same_extent(const b200gan_view* a, const b200gan_view* b) { return a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c; }

// Conv-geometry dispatcher.  fine/coarse are the two activation sides of the k/stride/pad geometry; `transposed` says the
// call came through the ConvTranspose2d entry points (its forward is the DGRAD primitive, its input gradient the FPROP one).
// The gradient operand of the call (for the dy_* fusion) is: DGRAD prim -> coarse, FPROP prim -> fine, WGRAD -> coarse for
// Conv2d and fine for ConvTranspose2d.
static int conv_dispatch(Prim prim, bool transposed, const b200gan_conv* cv, const b200gan_view* fine, const b200gan_view* coarse,
                         const float* w, const void* wpacked, float* dw, float* workspace, const b200gan_fuse* fuse, void* stream,
                         const char* what) {
  int rc;
  if ((rc = check_conv(cv))) return rc;
  if ((rc = check_view(fine, what))) return rc;
  if ((rc = check_view(coarse, what))) return rc;
  if ((rc = check_pair(cv, fine, coarse, what))) return rc;
  if (prim != WGRAD && !w) { set_error("%s: null weight", what); return B200GAN_ERR_BAD_ARG; }
  if (prim == WGRAD && !dw) { set_error("%s: null dweight", what); return B200GAN_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;

  // ---- resolve the requested fusions -------------------------------------------------------------
  const bool is_fwd = (prim == FPROP && !transposed) || (prim == DGRAD && transposed);     // the API's forward call
  const b200gan_view* result = prim == FPROP ? coarse : fine;                                // written by FPROP / DGRAD prims
  const b200gan_view* gathered = prim == FPROP ? fine : coarse;
  ConvFuse fz;
  double* bn_sums = nullptr;
  bool prev = false;            // activation backward (+ BatchNorm-backward sums when prev_scale is given) on the result of a dgrad
  bool prev_bn = false;
  if (fuse) {
    const bool has_out = fuse->out_act != B200GAN_ACT_NONE, has_dy = fuse->dy_act != B200GAN_ACT_NONE;
    prev = fuse->prev_y != nullptr;
    prev_bn = prev && fuse->prev_scale != nullptr;
    if (is_fwd) {
      B200_CHECK_ARG(!has_dy && !prev, "%s: dy_* / prev_* fusions do not apply to a forward convolution", what);
      B200_CHECK_ARG(fuse->out_act >= B200GAN_ACT_NONE && fuse->out_act <= B200GAN_ACT_SIGMOID, "%s: bad out_act %d", what, fuse->out_act);
      B200_CHECK_ARG(!(has_out && fuse->bn_sums), "%s: out_act and bn_sums are mutually exclusive (BatchNorm sits before the activation)", what);
      fz.out_act = fuse->out_act; fz.out_slope = fuse->out_slope;
      bn_sums = fuse->bn_sums;
    } else {
      B200_CHECK_ARG(!has_out && !fuse->bn_sums, "%s: out_act / bn_sums only apply to a forward convolution", what);
      B200_CHECK_ARG(prim != WGRAD || !prev, "%s: prev_* fusion only applies to an input-gradient convolution", what);
      if (has_dy) {
        B200_CHECK_ARG(fuse->dy_act >= B200GAN_ACT_NONE && fuse->dy_act <= B200GAN_ACT_SIGMOID && fuse->dy_ref, "%s: dy_act needs dy_ref", what);
        if ((rc = check_view(fuse->dy_ref, what))) return rc;
        const b200gan_view* g = prim == WGRAD ? (transposed ? fine : coarse) : gathered;
        B200_CHECK_ARG(same_extent(fuse->dy_ref, g), "%s: dy_ref extents differ from dy", what);
        fz.g_ref = fuse->dy_ref; fz.g_act = fuse->dy_act; fz.g_slope = fuse->dy_slope;
      }
      if (prev) {
        B200_CHECK_ARG(!prev_bn || (fuse->prev_shift && fuse->prev_mean && fuse->prev_invstd && fuse->prev_sums), "%s: prev_* pointers missing", what);
        if ((rc = check_view(fuse->prev_y, what))) return rc;
        B200_CHECK_ARG(same_extent(fuse->prev_y, result) && fuse->prev_y->dtype == result->dtype, "%s: prev_y must match dx in extents and dtype", what);
        if (fuse->prev_act != B200GAN_ACT_RELU && fuse->prev_act != B200GAN_ACT_LRELU && fuse->prev_act != B200GAN_ACT_NONE) {
          set_error("%s: prev_act must be NONE, RELU or LRELU (the activations that follow a BatchNorm in dcgan.py)", what);
          return B200GAN_ERR_UNSUPPORTED;
        }
      }
    }
  }
  const bool grad_is_coarse = !transposed;      // WGRAD only
  const bool plain = fz.out_act == B200GAN_ACT_NONE && !fz.g_ref;

  // ---- kernel selection ----------------------------------------------------------------------------
  int t = 1;
  if (cv->algo != B200GAN_ALGO_SIMT) {
    if (plain) {
      // the tensor-core kernels absorb the BatchNorm fusions in their epilogue
      TcEpi epi;
      if (bn_sums) { epi.mode = 1; epi.sums = bn_sums; }
      if (prev) {
        epi.mode = prev_bn ? 2 : 3; epi.sums = fuse->prev_sums; epi.prev_y = fuse->prev_y; epi.scale = fuse->prev_scale; epi.shift = fuse->prev_shift;
        epi.mean = fuse->prev_mean; epi.invstd = fuse->prev_invstd; epi.act = fuse->prev_act; epi.slope = fuse->prev_slope;
      }
      if (prim == FPROP) t = tc_conv_fprop(cv, fine, wpacked, coarse, epi, st);
      else if (prim == DGRAD) t = tc_conv_dgrad(cv, coarse, wpacked, fine, epi, st);
      else t = tc_conv_wgrad(cv, fine, coarse, dw, workspace, st);
      if (t < 0) return t;
      if (t == 0) { bn_sums = nullptr; prev = false; }
    }
    if (t > 0 && cv->algo == B200GAN_ALGO_TCGEN05) {
      set_error("%s: shape/dtype/fusion not supported by the tcgen05 path", what);
      return B200GAN_ERR_UNSUPPORTED;
    }
    // image-side layers (warp-level MMA, fused activations), the latent GEMM and the 7x7 GEMV
    const bool k4 = cv->k == 4 && cv->stride == 2 && cv->pad == 1;
    if (t > 0 && k4) {
      if (prim == FPROP) {
        TcEpi epi;                                   // the image-side "down" kernel absorbs the BatchNorm fusions too
        if (bn_sums) { epi.mode = 1; epi.sums = bn_sums; }
        if (prev_bn) {
          epi.mode = 2; epi.sums = fuse->prev_sums; epi.prev_y = fuse->prev_y; epi.scale = fuse->prev_scale; epi.shift = fuse->prev_shift;
          epi.mean = fuse->prev_mean; epi.invstd = fuse->prev_invstd; epi.act = fuse->prev_act; epi.slope = fuse->prev_slope;
        }
        t = (prev && !prev_bn) ? 1 : thin_down(fine, fz.g_ref, fz.g_act, w, coarse, fz.out_act, fz.g_ref ? fz.g_slope : fz.out_slope, epi, st);
        if (t == 0) { bn_sums = nullptr; prev = false; }
      }
      else if (prim == DGRAD) t = thin_up(coarse, fz.g_ref, fz.g_act, fz.g_slope, w, fine, fz.out_act, st);
      else t = thin_wgrad(fine, grad_is_coarse ? nullptr : fz.g_ref, fz.g_act, coarse, grad_is_coarse ? fz.g_ref : nullptr, fz.g_act, fz.g_slope, dw, st);
      if (t < 0) return t;
      if (t > 0) {                                   // other feature widths (they absorb no BatchNorm fusion: the passes below run)
        if (prim == FPROP) t = edge_down(fine, fz.g_ref, fz.g_act, w, coarse, fz.out_act, fz.g_ref ? fz.g_slope : fz.out_slope, st);
        else if (prim == DGRAD) t = edge_up(coarse, fz.g_ref, fz.g_act, fz.g_slope, w, fine, fz.out_act, st);
        else t = edge_wgrad(fine, grad_is_coarse ? nullptr : fz.g_ref, fz.g_act, coarse, grad_is_coarse ? fz.g_ref : nullptr, fz.g_act, fz.g_slope, dw, st);
        if (t < 0) return t;
      }
    }
    if (t > 0 && plain) {
      if (prim == FPROP) { t = window_fprop(cv, fine, w, coarse, st); if (t > 0) t = score_fprop(cv, fine, w, coarse, st); }
      else if (prim == DGRAD) {
        TcEpi epi;                                   // the GEMV input gradient absorbs the BatchNorm-backward fusion of the layer below
        if (prev_bn) {
          epi.mode = 2; epi.sums = fuse->prev_sums; epi.prev_y = fuse->prev_y; epi.scale = fuse->prev_scale; epi.shift = fuse->prev_shift;
          epi.mean = fuse->prev_mean; epi.invstd = fuse->prev_invstd; epi.act = fuse->prev_act; epi.slope = fuse->prev_slope;
        }
        t = (prev && !prev_bn) ? 1 : window_dgrad(cv, coarse, w, fine, epi, st);
        if (t == 0) prev = false;
        if (t > 0) t = score_dgrad(cv, coarse, w, fine, st);
        if (t > 0) t = latent_fprop_mma(cv, coarse, w, fine, st);
        if (t > 0) t = latent_fprop(cv, coarse, w, fine, st);
      } else {
        t = window_wgrad(cv, fine, coarse, dw, st);
        if (t > 0) t = score_wgrad(cv, fine, coarse, dw, st);
        if (t > 0) t = latent_wgrad_mma(cv, fine, coarse, dw, st);
        if (t > 0) t = latent_wgrad(cv, fine, coarse, dw, st);
      }
      if (t < 0) return t;
    }
  }
  if (t > 0) {
    if (prim == FPROP) t = simt_conv_fprop(cv, fine, w, coarse, fz, st);
    else if (prim == DGRAD) t = simt_conv_dgrad(cv, coarse, w, fine, fz, st);
    else t = simt_conv_wgrad(cv, fine, coarse, dw, fz, grad_is_coarse, st);
    if (t) return t;
  }
  // ---- BatchNorm fusions not absorbed by the kernel: the equivalent passes ---------------------------------
  if (bn_sums && (rc = ew_bn_stats(result, bn_sums, st))) return rc;
  if (prev && !prev_bn) {
    // no BatchNorm below: prev_y is the saved activation output, whose sign gives act'
    if ((rc = ew_act_bwd_inplace(result, fuse->prev_y, nullptr, nullptr, fuse->prev_act, fuse->prev_slope, st))) return rc;
  } else if (prev) {
    // one dense pass (dz stored in place + sums) when the layout allows, reduce + in-place pass otherwise
    rc = ew_bn_bwd_reduce(result, fuse->prev_y, nullptr, fuse->prev_scale, fuse->prev_shift, fuse->prev_mean, fuse->prev_invstd,
                          fuse->prev_act, fuse->prev_slope, fuse->prev_sums, result, st);
    if (rc < 0) return rc;
    if (rc == 2 && (rc = ew_act_bwd_inplace(result, fuse->prev_y, fuse->prev_scale, fuse->prev_shift, fuse->prev_act, fuse->prev_slope, st))) return rc;
  }
  return 0;
}

}  // namespace b200gan

using namespace b200gan;

extern "C" {

int b200gan_version(void) { return B200GAN_VERSION; }

const char* b200gan_last_error_string(void) { return g_err; }

int b200gan_device_info(int device, char* name, int* cc_major, int* cc_minor) {
  cudaDeviceProp prop;
  B200_CUDA(cudaGetDeviceProperties(&prop, device));
  if (name) { strncpy(name, prop.name, 255); name[255] = 0; }
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return prop.multiProcessorCount;
}

int b200gan_conv2d_fprop(const b200gan_conv* cv, const b200gan_view* x, const float* weight, const void* wpacked,
                         const b200gan_view* y, const b200gan_fuse* fuse, void* stream) {
  return conv_dispatch(FPROP, false, cv, x, y, weight, wpacked, nullptr, nullptr, fuse, stream, "conv2d_fprop");
}
int b200gan_conv2d_dgrad(const b200gan_conv* cv, const b200gan_view* dy, const float* weight, const void* wpacked,
                         const b200gan_view* dx, const b200gan_fuse* fuse, void* stream) {
  return conv_dispatch(DGRAD, false, cv, dx, dy, weight, wpacked, nullptr, nullptr, fuse, stream, "conv2d_dgrad");
}
int b200gan_conv2d_wgrad(const b200gan_conv* cv, const b200gan_view* x, const b200gan_view* dy, float* dweight, float* workspace,
                         const b200gan_fuse* fuse, void* stream) {
  return conv_dispatch(WGRAD, false, cv, x, dy, nullptr, nullptr, dweight, workspace, fuse, stream, "conv2d_wgrad");
}
// ConvTranspose2d == the conv input-gradient on the same geometry: its input is the coarse side, its
// output the fine side, and its weight (Cin_T, Cout_T, k, k) is the conv weight (Co, Ci, k, k).
int b200gan_convT2d_fprop(const b200gan_conv* cv, const b200gan_view* x, const float* weight, const void* wpacked,
                          const b200gan_view* y, const b200gan_fuse* fuse, void* stream) {
  return conv_dispatch(DGRAD, true, cv, y, x, weight, wpacked, nullptr, nullptr, fuse, stream, "convT2d_fprop");
}
int b200gan_convT2d_dgrad(const b200gan_conv* cv, const b200gan_view* dy, const float* weight, const void* wpacked,
                          const b200gan_view* dx, const b200gan_fuse* fuse, void* stream) {
  return conv_dispatch(FPROP, true, cv, dy, dx, weight, wpacked, nullptr, nullptr, fuse, stream, "convT2d_dgrad");
}
int b200gan_convT2d_wgrad(const b200gan_conv* cv, const b200gan_view* x, const b200gan_view* dy, float* dweight, float* workspace,
                          const b200gan_fuse* fuse, void* stream) {
  return conv_dispatch(WGRAD, true, cv, dy, x, nullptr, nullptr, dweight, workspace, fuse, stream, "convT2d_wgrad");
}

int64_t b200gan_conv_wgrad_workspace_floats(const b200gan_conv* cv, const b200gan_view* x, const b200gan_view* dy, int32_t transposed) {
  int rc;
  if ((rc = check_conv(cv))) return rc;
  if ((rc = check_view(x, "conv_wgrad_workspace_floats"))) return rc;
  if ((rc = check_view(dy, "conv_wgrad_workspace_floats"))) return rc;
  // conv geometry: fine side carries Ci, coarse side Co; the tensor-core weight gradient (k4 s2 p1, dense NHWC bf16 operands, Co % 64 == 0,
  // Ci in {32, 64, multiples of 128}) reduces its split-K partial sums in a [Co][16][Ci] fp32 slab; every other kernel adds into dweight
  const b200gan_view* fine = transposed ? dy : x;
  const b200gan_view* coarse = transposed ? x : dy;
  if (cv->algo == B200GAN_ALGO_SIMT || cv->k != 4 || cv->stride != 2 || cv->pad != 1) return 0;
  if (fine->dtype != B200GAN_BF16 || coarse->dtype != B200GAN_BF16) return 0;
  const int Ci = fine->c, Co = coarse->c;
  if (Co % 64 != 0 || (Ci != 32 && Ci != 64 && Ci % 128 != 0)) return 0;
  return (int64_t)Co * Ci * 16;
}

int b200gan_pack_conv_weight(const float* weight, int32_t co, int32_t ci, int32_t k, int32_t form, void* out, void* stream) {
  B200_CHECK_ARG(weight && out && co > 0 && ci > 0 && form >= 0 && form <= 2, "pack_conv_weight: bad argument");
  return tc_pack_weight(weight, co, ci, k, form, out, (cudaStream_t)stream);
}

int b200gan_bn_stats(const b200gan_view* y, double* sums, void* stream) {
  int rc;
  if ((rc = check_view(y, "bn_stats"))) return rc;
  B200_CHECK_ARG(sums, "bn_stats: null sums");
  return ew_bn_stats(y, sums, (cudaStream_t)stream);
}

int b200gan_bn_finalize(double* sums, int32_t channels, int64_t count, const float* gamma, const float* beta,
                        float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum, float eps,
                        float* scale, float* shift, float* save_mean, float* save_invstd, void* stream) {
  B200_CHECK_ARG(sums && gamma && beta && scale && shift && save_mean && save_invstd, "bn_finalize: null pointer");
  B200_CHECK_ARG(channels > 0 && count > 0, "bn_finalize: channels=%d count=%lld", channels, (long long)count);
  B200_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_finalize: running_mean/var must both be given or both NULL");
  return ew_bn_finalize(sums, channels, count, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps,
                        scale, shift, save_mean, save_invstd, (cudaStream_t)stream);
}

int b200gan_bn_eval_coeffs(int32_t channels, const float* gamma, const float* beta, const float* running_mean,
                           const float* running_var, float eps, float* scale, float* shift, void* stream) {
  B200_CHECK_ARG(gamma && beta && running_mean && running_var && scale && shift && channels > 0, "bn_eval_coeffs: bad argument");
  return ew_bn_eval_coeffs(channels, gamma, beta, running_mean, running_var, eps, scale, shift, (cudaStream_t)stream);
}

int b200gan_bn_act_fwd(const b200gan_view* y, const float* scale, const float* shift, int32_t act, float slope,
                       const b200gan_view* a, void* stream) {
  int rc;
  if ((rc = check_view(y, "bn_act_fwd"))) return rc;
  if ((rc = check_view(a, "bn_act_fwd"))) return rc;
  B200_CHECK_ARG((scale == nullptr) == (shift == nullptr), "bn_act_fwd: scale/shift must both be given or both NULL");
  B200_CHECK_ARG(act >= B200GAN_ACT_NONE && act <= B200GAN_ACT_SIGMOID, "bn_act_fwd: bad activation %d", act);
  return ew_bn_act_fwd(y, scale, shift, act, slope, a, (cudaStream_t)stream);
}

int b200gan_bn_finalize_act_fwd(double* sums, int32_t channels, int64_t count, const float* gamma, const float* beta, float* running_mean,
                                float* running_var, int64_t* num_batches_tracked, float momentum, float eps, float* scale, float* shift,
                                float* save_mean, float* save_invstd, const b200gan_view* y, int32_t act, float slope, const b200gan_view* a,
                                void* stream) {
  int rc;
  if ((rc = check_view(y, "bn_finalize_act_fwd"))) return rc;
  if ((rc = check_view(a, "bn_finalize_act_fwd"))) return rc;
  B200_CHECK_ARG(sums && gamma && beta && scale && shift && save_mean && save_invstd, "bn_finalize_act_fwd: null pointer");
  B200_CHECK_ARG(channels > 0 && count > 0 && y->c == channels, "bn_finalize_act_fwd: channels=%d count=%lld (tensor has %d)", channels, (long long)count, y->c);
  B200_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_finalize_act_fwd: running_mean/var must both be given or both NULL");
  B200_CHECK_ARG(act >= B200GAN_ACT_NONE && act <= B200GAN_ACT_SIGMOID, "bn_finalize_act_fwd: bad activation %d", act);
  B200_CHECK_ARG(y->n == a->n && y->h == a->h && y->w == a->w && y->c == a->c, "bn_finalize_act_fwd: extent mismatch");
  rc = ew_bn_finalize_act_fwd(sums, channels, count, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, scale, shift,
                              save_mean, save_invstd, y, act, slope, a, (cudaStream_t)stream);
  if (rc != 1) return rc;
  // layouts the one-launch kernel does not take: the two passes it stands for
  if ((rc = ew_bn_finalize(sums, channels, count, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, scale, shift,
                           save_mean, save_invstd, (cudaStream_t)stream)))
    return rc;
  return ew_bn_act_fwd(y, scale, shift, act, slope, a, (cudaStream_t)stream);
}

int b200gan_bn_act_bwd_reduce(const b200gan_view* da, const b200gan_view* y, const b200gan_view* a, const float* scale,
                              const float* shift, const float* save_mean, const float* save_invstd, int32_t act, float slope,
                              double* sums, void* stream) {
  int rc;
  if ((rc = check_view(da, "bn_act_bwd_reduce"))) return rc;
  if ((rc = check_view(y, "bn_act_bwd_reduce"))) return rc;
  if (a && (rc = check_view(a, "bn_act_bwd_reduce"))) return rc;
  B200_CHECK_ARG(scale && shift && save_mean && save_invstd && sums, "bn_act_bwd_reduce: null pointer");
  B200_CHECK_ARG((act != B200GAN_ACT_TANH && act != B200GAN_ACT_SIGMOID) || a, "bn_act_bwd_reduce: tanh/sigmoid need the saved output");
  return ew_bn_bwd_reduce(da, y, a, scale, shift, save_mean, save_invstd, act, slope, sums, nullptr, (cudaStream_t)stream);
}

int b200gan_bn_act_bwd_apply(const b200gan_view* da, const b200gan_view* y, const b200gan_view* a, const float* scale,
                             const float* shift, const float* save_mean, const float* save_invstd, const float* gamma,
                             double* sums, int64_t count, int32_t act, float slope, const b200gan_view* dy, float* dgamma,
                             float* dbeta, void* stream) {
  int rc;
  if ((rc = check_view(da, "bn_act_bwd_apply"))) return rc;
  if ((rc = check_view(y, "bn_act_bwd_apply"))) return rc;
  if ((rc = check_view(dy, "bn_act_bwd_apply"))) return rc;
  if (a && (rc = check_view(a, "bn_act_bwd_apply"))) return rc;
  if (scale) B200_CHECK_ARG(shift && save_mean && save_invstd && gamma && sums && count > 0, "bn_act_bwd_apply: null pointer");
  B200_CHECK_ARG((act != B200GAN_ACT_TANH && act != B200GAN_ACT_SIGMOID) || a, "bn_act_bwd_apply: tanh/sigmoid need the saved output");
  return ew_bn_act_bwd_apply(da, y, a, scale, shift, save_mean, save_invstd, gamma, sums, count, act, slope, dy, dgamma, dbeta,
                             (cudaStream_t)stream);
}

int b200gan_bce_sigmoid(const float* logit, int32_t batch, float target, float grad_scale, float* prob, float* out2,
                        float* dlogit, void* stream) {
  B200_CHECK_ARG(logit && out2 && batch > 0, "bce_sigmoid: bad argument");
  return ew_bce_sigmoid(logit, batch, target, grad_scale, prob, out2, dlogit, (cudaStream_t)stream);
}

int b200gan_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, double lr, double beta1,
                 double beta2, double eps, int32_t step, const int64_t* step_dev, float grad_scale, void* stream) {
  B200_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && numel > 0 && (step >= 1 || step_dev), "adam: bad argument");
  return ew_adam(param, grad, exp_avg, exp_avg_sq, numel, lr, beta1, beta2, eps, step < 1 ? 1 : step, step_dev, grad_scale, (cudaStream_t)stream);
}

int b200gan_copy_view(const b200gan_view* src, const b200gan_view* dst, void* stream) {
  int rc;
  if ((rc = check_view(src, "copy_view"))) return rc;
  if ((rc = check_view(dst, "copy_view"))) return rc;
  return ew_copy_view(src, dst, (cudaStream_t)stream);
}

int b200gan_fill_f32(float* ptr, int64_t numel, float value, void* stream) {
  B200_CHECK_ARG(ptr || numel == 0, "fill_f32: null pointer");
  return ew_fill(ptr, numel, value, (cudaStream_t)stream);
}

int b200gan_sample_sumsq(const b200gan_view* x, double* sumsq, void* stream) {
  int rc;
  if ((rc = check_view(x, "sample_sumsq"))) return rc;
  B200_CHECK_ARG(sumsq, "sample_sumsq: null output");
  return gp_sample_sumsq(x, sumsq, (cudaStream_t)stream);
}

int b200gan_gp_from_norms(const double* sumsq, int32_t batch, float lambda_gp, float* gp, float* coeff, void* stream) {
  B200_CHECK_ARG(sumsq && gp && coeff && batch > 0, "gp_from_norms: bad argument");
  return gp_from_norms(sumsq, batch, lambda_gp, gp, coeff, (cudaStream_t)stream);
}

int b200gan_sample_axpby(const b200gan_view* x, const float* a, const b200gan_view* y, const float* b, const b200gan_view* out, void* stream) {
  int rc;
  if ((rc = check_view(x, "sample_axpby"))) return rc;
  if ((rc = check_view(out, "sample_axpby"))) return rc;
  if (y && (rc = check_view(y, "sample_axpby"))) return rc;
  return gp_sample_axpby(x, a, y, b, out, (cudaStream_t)stream);
}

int b200gan_bn_bwd_bwd(const b200gan_view* r, const b200gan_view* y, const b200gan_view* dz, const float* scale, const float* shift,
                       const float* save_mean, const float* save_invstd, const float* gamma, const double* dz_sums, int64_t count, int32_t act,
                       float slope, const b200gan_view* u, const b200gan_view* inj, float* dgamma, double* workspace, void* stream) {
  int rc;
  if ((rc = check_view(r, "bn_bwd_bwd"))) return rc;
  if ((rc = check_view(y, "bn_bwd_bwd"))) return rc;
  if ((rc = check_view(dz, "bn_bwd_bwd"))) return rc;
  if ((rc = check_view(u, "bn_bwd_bwd"))) return rc;
  if ((rc = check_view(inj, "bn_bwd_bwd"))) return rc;
  B200_CHECK_ARG(scale && shift && save_mean && save_invstd && gamma && dz_sums && workspace && count > 0, "bn_bwd_bwd: null pointer");
  B200_CHECK_ARG(act == B200GAN_ACT_NONE || act == B200GAN_ACT_RELU || act == B200GAN_ACT_LRELU, "bn_bwd_bwd: activation must be NONE, RELU or LRELU");
  return gp_bn_bwd_bwd(r, y, dz, scale, shift, save_mean, save_invstd, gamma, dz_sums, count, act, slope, u, inj, dgamma, workspace, (cudaStream_t)stream);
}

int b200gan_mean_f32(const float* x, int64_t n, float scale, float* out, void* stream) {
  B200_CHECK_ARG(x && out && n > 0, "mean_f32: bad argument");
  return gp_mean_f32(x, n, scale, out, (cudaStream_t)stream);
}

int b200gan_bce_logits(const float* logit, const float* target, int32_t batch, float grad_scale, float* out2, float* dlogit, void* stream) {
  B200_CHECK_ARG(logit && target && out2 && batch > 0, "bce_logits: bad argument");
  return cgan_bce_logits(logit, target, batch, grad_scale, out2, dlogit, (cudaStream_t)stream);
}

int b200gan_fm_pair(const b200gan_view* real, const b200gan_view* fake, const b200gan_view* dfake, float coeff, int32_t add, double* sum, void* stream) {
  int rc;
  if ((rc = check_view(real, "fm_pair"))) return rc;
  if ((rc = check_view(fake, "fm_pair"))) return rc;
  if (dfake && (rc = check_view(dfake, "fm_pair"))) return rc;
  B200_CHECK_ARG(sum, "fm_pair: null sum");
  B200_CHECK_ARG(same_extent(real, fake) && (!dfake || same_extent(dfake, fake)), "fm_pair: views differ in extent");
  return cgan_fm_pair(real, fake, dfake, coeff, add, sum, (cudaStream_t)stream);
}

int b200gan_accumulate_2d(float* dst, const void* src, int32_t src_f64, int32_t rows, int32_t cols, int64_t src_row_stride, int64_t src_col_stride,
                          void* stream) {
  B200_CHECK_ARG(dst && src && rows > 0 && cols > 0, "accumulate_2d: bad argument");
  return cgan_accumulate_2d(dst, src, src_f64, rows, cols, src_row_stride, src_col_stride, (cudaStream_t)stream);
}

int b200gan_embed_add(const float* table, const int64_t* labels, const float* z, int32_t batch, int32_t dim, int32_t tail, float* out, void* stream) {
  B200_CHECK_ARG(table && labels && out && batch > 0 && dim > 0 && tail >= 0, "embed_add: bad argument");
  return cgan_embed_add(table, labels, z, batch, dim, tail, out, (cudaStream_t)stream);
}

int b200gan_embed_bwd(const float* dx, const int64_t* labels, int32_t batch, int32_t dim, int32_t stride, int32_t num_classes, float* dtable,
                      void* stream) {
  B200_CHECK_ARG(dx && labels && dtable && batch > 0 && dim > 0 && stride >= dim && num_classes > 0, "embed_bwd: bad argument");
  return cgan_embed_bwd(dx, labels, batch, dim, stride, num_classes, dtable, (cudaStream_t)stream);
}

int b200gan_upconv3_fold(const float* w3, int32_t co, int32_t ci, float* w4, void* stream) {
  B200_CHECK_ARG(w3 && w4 && co > 0 && ci > 0, "upconv3_fold: bad argument");
  return cgan_upconv3_fold(w3, co, ci, w4, (cudaStream_t)stream);
}

int b200gan_upconv3_unfold(const float* dw4, int32_t co, int32_t ci, float* dw3, void* stream) {
  B200_CHECK_ARG(dw4 && dw3 && co > 0 && ci > 0, "upconv3_unfold: bad argument");
  return cgan_upconv3_unfold(dw4, co, ci, dw3, (cudaStream_t)stream);
}

int b200gan_class_proj_fwd(const b200gan_view* x, const float* table, const int64_t* labels, float* out, void* stream) {
  int rc;
  if ((rc = check_view(x, "class_proj_fwd"))) return rc;
  B200_CHECK_ARG(table && labels && out, "class_proj_fwd: null pointer");
  return cgan_class_proj_fwd(x, table, labels, out, (cudaStream_t)stream);
}

int b200gan_class_proj_bwd(const b200gan_view* x, const float* table, const int64_t* labels, const float* dout, const b200gan_view* dx,
                           int32_t num_classes, float* dtable, void* stream) {
  int rc;
  if ((rc = check_view(x, "class_proj_bwd"))) return rc;
  if (dx && (rc = check_view(dx, "class_proj_bwd"))) return rc;
  B200_CHECK_ARG(table && labels && dout && num_classes > 0, "class_proj_bwd: bad argument");
  B200_CHECK_ARG(!dx || (dx->n == x->n && dx->h == x->h && dx->w == x->w && dx->c == x->c), "class_proj_bwd: dx and x differ in extent");
  return cgan_class_proj_bwd(x, table, labels, dout, dx, num_classes, dtable, (cudaStream_t)stream);
}

int b200gan_conv3x3_fold(const float* w3, int32_t co, int32_t ci, int32_t ci_pad, float* w4, void* stream) {
  B200_CHECK_ARG(w3 && w4 && co > 0 && ci > 0 && ci_pad >= ci, "conv3x3_fold: bad argument");
  return vgg_conv3x3_fold(w3, co, ci, ci_pad, w4, (cudaStream_t)stream);
}

int b200gan_bias_relu_d2s(const b200gan_view* t, const float* bias, const b200gan_view* a, void* stream) {
  int rc;
  if ((rc = check_view(t, "bias_relu_d2s"))) return rc;
  if ((rc = check_view(a, "bias_relu_d2s"))) return rc;
  B200_CHECK_ARG(bias, "bias_relu_d2s: null bias");
  return vgg_bias_relu_d2s(t, bias, a, (cudaStream_t)stream);
}

int b200gan_relu_bwd_s2d(const b200gan_view* da, const b200gan_view* a, const b200gan_view* dt, void* stream) {
  int rc;
  if ((rc = check_view(da, "relu_bwd_s2d"))) return rc;
  if ((rc = check_view(a, "relu_bwd_s2d"))) return rc;
  if ((rc = check_view(dt, "relu_bwd_s2d"))) return rc;
  return vgg_relu_bwd_s2d(da, a, dt, (cudaStream_t)stream);
}

int b200gan_maxpool2_fwd(const b200gan_view* a, const b200gan_view* p, void* stream) {
  int rc;
  if ((rc = check_view(a, "maxpool2_fwd"))) return rc;
  if ((rc = check_view(p, "maxpool2_fwd"))) return rc;
  return vgg_maxpool2_fwd(a, p, (cudaStream_t)stream);
}

int b200gan_maxpool2_bwd(const b200gan_view* a, const b200gan_view* dp, const b200gan_view* da, int32_t add, void* stream) {
  int rc;
  if ((rc = check_view(a, "maxpool2_bwd"))) return rc;
  if ((rc = check_view(dp, "maxpool2_bwd"))) return rc;
  if ((rc = check_view(da, "maxpool2_bwd"))) return rc;
  return vgg_maxpool2_bwd(a, dp, da, add, (cudaStream_t)stream);
}

int b200gan_gather_augment(const uint8_t* cache, int64_t num_images, const int64_t* index, const uint8_t* flip, const float* mean,
                           const float* std, const b200gan_view* out, void* stream) {
  int rc;
  B200_CHECK_ARG(cache && index && num_images > 0, "gather_augment: null cache / index or empty cache");
  if ((rc = check_view(out, "gather_augment"))) return rc;
  return ew_gather_augment(cache, num_images, index, flip, mean, std, out, (cudaStream_t)stream);
}

}  // extern "C"
