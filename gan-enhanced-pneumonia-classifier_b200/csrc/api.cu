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

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return B200GAN_ERR_CUDA;
}

// SIMT kernels (conv_simt.cu)
int simt_conv_fprop(const b200gan_conv*, const b200gan_view* x, const float* w, const b200gan_view* y, cudaStream_t);
int simt_conv_dgrad(const b200gan_conv*, const b200gan_view* dy, const float* w, const b200gan_view* dx, cudaStream_t);
int simt_conv_wgrad(const b200gan_conv*, const b200gan_view* x, const b200gan_view* dy, float* dw, cudaStream_t);
// tensor-core kernels (conv_tc.cu): return 1 when the problem does not qualify (caller falls back / errors)
int tc_conv_fprop(const b200gan_conv*, const b200gan_view* x, const void* wpacked, const b200gan_view* y, cudaStream_t);
int tc_conv_dgrad(const b200gan_conv*, const b200gan_view* dy, const void* wpacked, const b200gan_view* dx, cudaStream_t);
int tc_conv_wgrad(const b200gan_conv*, const b200gan_view* x, const b200gan_view* dy, float* dw, cudaStream_t);
int tc_pack_weight(const float* w, int Co, int Ci, int k, int form, void* out, cudaStream_t);
// shape-specialised CUDA-core kernels (conv_thin.cu): same return convention
int thin_down(const b200gan_view* fine, const float* w, const b200gan_view* coarse, int act, float slope, cudaStream_t);
int thin_up(const b200gan_view* coarse, const float* w, const b200gan_view* fine, int act, cudaStream_t);
int thin_wgrad(const b200gan_view* fine, const b200gan_view* coarse, float* dw, cudaStream_t);
int window_fprop(const b200gan_conv*, const b200gan_view* x, const float* w, const b200gan_view* y, cudaStream_t);
int window_dgrad(const b200gan_conv*, const b200gan_view* dy, const float* w, const b200gan_view* dx, cudaStream_t);
int window_wgrad(const b200gan_conv*, const b200gan_view* x, const b200gan_view* dy, float* dw, cudaStream_t);
int latent_fprop(const b200gan_conv*, const b200gan_view* z, const float* w, const b200gan_view* y, cudaStream_t);
int latent_wgrad(const b200gan_conv*, const b200gan_view* dy_fine, const b200gan_view* z, float* dw, cudaStream_t);
// elementwise.cu
int ew_bn_stats(const b200gan_view*, double*, cudaStream_t);
int ew_bn_bwd_reduce(const b200gan_view*, const b200gan_view*, const b200gan_view*, const float*, const float*, const float*,
                     const float*, int, float, double*, cudaStream_t);
int ew_bn_finalize(double*, int, int64_t, const float*, const float*, float*, float*, int64_t*, float, float, float*, float*,
                   float*, float*, cudaStream_t);
int ew_bn_eval_coeffs(int, const float*, const float*, const float*, const float*, float, float*, float*, cudaStream_t);
int ew_bn_act_fwd(const b200gan_view*, const float*, const float*, int, float, const b200gan_view*, cudaStream_t);
int ew_bn_act_bwd_apply(const b200gan_view*, const b200gan_view*, const b200gan_view*, const float*, const float*, const float*,
                        const float*, const float*, double*, int64_t, int, float, const b200gan_view*, float*, float*,
                        cudaStream_t);
int ew_bce_sigmoid(const float*, int, float, float, float*, float*, float*, cudaStream_t);
int ew_adam(float*, const float*, float*, float*, int64_t, double, double, double, double, int, float, cudaStream_t);
int ew_copy_view(const b200gan_view*, const b200gan_view*, cudaStream_t);
int ew_fill(float*, int64_t, float, cudaStream_t);

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

static int conv_dispatch(Prim prim, const b200gan_conv* cv, const b200gan_view* fine, const b200gan_view* coarse,
                         const float* w, const void* wpacked, float* dw, void* stream, const char* what) {
  int rc;
  if ((rc = check_conv(cv))) return rc;
  if ((rc = check_view(fine, what))) return rc;
  if ((rc = check_view(coarse, what))) return rc;
  if ((rc = check_pair(cv, fine, coarse, what))) return rc;
  if (prim != WGRAD && !w) { set_error("%s: null weight", what); return B200GAN_ERR_BAD_ARG; }
  if (prim == WGRAD && !dw) { set_error("%s: null dweight", what); return B200GAN_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  if (cv->algo != B200GAN_ALGO_SIMT) {
    int t = 1;
    if (prim == FPROP) t = tc_conv_fprop(cv, fine, wpacked, coarse, st);
    else if (prim == DGRAD) t = tc_conv_dgrad(cv, coarse, wpacked, fine, st);
    else t = tc_conv_wgrad(cv, fine, coarse, dw, st);
    if (t <= 0) return t;                       // done (0) or hard error (<0)
    if (cv->algo == B200GAN_ALGO_TCGEN05) {
      set_error("%s: shape/dtype not supported by the tcgen05 path", what);
      return B200GAN_ERR_UNSUPPORTED;
    }
    // shape-specialised CUDA-core kernels for the thin image-side layers, the latent GEMM and the 7x7 GEMV
    const bool k4 = cv->k == 4 && cv->stride == 2 && cv->pad == 1;
    if (prim == FPROP) {
      if (k4) t = thin_down(fine, w, coarse, B200GAN_ACT_NONE, 0.f, st);
      if (t > 0) t = window_fprop(cv, fine, w, coarse, st);
    } else if (prim == DGRAD) {
      if (k4) t = thin_up(coarse, w, fine, B200GAN_ACT_NONE, st);
      if (t > 0) t = window_dgrad(cv, coarse, w, fine, st);
      if (t > 0) t = latent_fprop(cv, coarse, w, fine, st);
    } else {
      if (k4) t = thin_wgrad(fine, coarse, dw, st);
      if (t > 0) t = window_wgrad(cv, fine, coarse, dw, st);
      if (t > 0) t = latent_wgrad(cv, fine, coarse, dw, st);
    }
    if (t <= 0) return t;
  }
  if (prim == FPROP) return simt_conv_fprop(cv, fine, w, coarse, st);
  if (prim == DGRAD) return simt_conv_dgrad(cv, coarse, w, fine, st);
  return simt_conv_wgrad(cv, fine, coarse, dw, st);
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
                         const b200gan_view* y, void* stream) {
  return conv_dispatch(FPROP, cv, x, y, weight, wpacked, nullptr, stream, "conv2d_fprop");
}
int b200gan_conv2d_dgrad(const b200gan_conv* cv, const b200gan_view* dy, const float* weight, const void* wpacked,
                         const b200gan_view* dx, void* stream) {
  return conv_dispatch(DGRAD, cv, dx, dy, weight, wpacked, nullptr, stream, "conv2d_dgrad");
}
int b200gan_conv2d_wgrad(const b200gan_conv* cv, const b200gan_view* x, const b200gan_view* dy, float* dweight, void* stream) {
  return conv_dispatch(WGRAD, cv, x, dy, nullptr, nullptr, dweight, stream, "conv2d_wgrad");
}
// ConvTranspose2d == the conv input-gradient on the same geometry: its input is the coarse side, its
// output the fine side, and its weight (Cin_T, Cout_T, k, k) is the conv weight (Co, Ci, k, k).
int b200gan_convT2d_fprop(const b200gan_conv* cv, const b200gan_view* x, const float* weight, const void* wpacked,
                          const b200gan_view* y, void* stream) {
  return conv_dispatch(DGRAD, cv, y, x, weight, wpacked, nullptr, stream, "convT2d_fprop");
}
int b200gan_convT2d_dgrad(const b200gan_conv* cv, const b200gan_view* dy, const float* weight, const void* wpacked,
                          const b200gan_view* dx, void* stream) {
  return conv_dispatch(FPROP, cv, dy, dx, weight, wpacked, nullptr, stream, "convT2d_dgrad");
}
int b200gan_convT2d_wgrad(const b200gan_conv* cv, const b200gan_view* x, const b200gan_view* dy, float* dweight, void* stream) {
  return conv_dispatch(WGRAD, cv, dy, x, nullptr, nullptr, dweight, stream, "convT2d_wgrad");
}

int b200gan_pack_conv_weight(const float* weight, int32_t co, int32_t ci, int32_t k, int32_t form, void* out, void* stream) {
  B200_CHECK_ARG(weight && out && co > 0 && ci > 0 && (form == 0 || form == 1), "pack_conv_weight: bad argument");
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

int b200gan_bn_act_bwd_reduce(const b200gan_view* da, const b200gan_view* y, const b200gan_view* a, const float* scale,
                              const float* shift, const float* save_mean, const float* save_invstd, int32_t act, float slope,
                              double* sums, void* stream) {
  int rc;
  if ((rc = check_view(da, "bn_act_bwd_reduce"))) return rc;
  if ((rc = check_view(y, "bn_act_bwd_reduce"))) return rc;
  if (a && (rc = check_view(a, "bn_act_bwd_reduce"))) return rc;
  B200_CHECK_ARG(scale && shift && save_mean && save_invstd && sums, "bn_act_bwd_reduce: null pointer");
  B200_CHECK_ARG((act != B200GAN_ACT_TANH && act != B200GAN_ACT_SIGMOID) || a, "bn_act_bwd_reduce: tanh/sigmoid need the saved output");
  return ew_bn_bwd_reduce(da, y, a, scale, shift, save_mean, save_invstd, act, slope, sums, (cudaStream_t)stream);
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
                 double beta2, double eps, int32_t step, float grad_scale, void* stream) {
  B200_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && numel > 0 && step >= 1, "adam: bad argument");
  return ew_adam(param, grad, exp_avg, exp_avg_sq, numel, lr, beta1, beta2, eps, step, grad_scale, (cudaStream_t)stream);
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

}  // extern "C"
