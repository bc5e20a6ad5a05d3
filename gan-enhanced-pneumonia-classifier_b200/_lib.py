"""ctypes binding of libb200gan.so (C ABI declared in include/b200gan.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C csrc`.  There is no fallback: if the
shared object is missing or a call fails, a RuntimeError is raised -- nothing silently runs on cuDNN or the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libb200gan.so')

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
ALGO_AUTO, ALGO_SIMT, ALGO_TCGEN05 = 0, 1, 2

_DTYPES = {torch.float32: F32, torch.bfloat16: BF16}


class View(C.Structure):
    """struct b200gan_view"""
    _fields_ = [('ptr', C.c_void_p), ('dtype', C.c_int32), ('n', C.c_int32), ('h', C.c_int32), ('w', C.c_int32),
                ('c', C.c_int32), ('sn', C.c_int64), ('sh', C.c_int64), ('sw', C.c_int64), ('sc', C.c_int64)]


class Conv(C.Structure):
    """struct b200gan_conv"""
    _fields_ = [('k', C.c_int32), ('stride', C.c_int32), ('pad', C.c_int32), ('algo', C.c_int32)]


_VP = C.POINTER(View)
_CP = C.POINTER(Conv)


class Fuse(C.Structure):
    """struct b200gan_fuse: optional fusions around one convolution call (see include/b200gan.h)."""
    _fields_ = [('out_act', C.c_int32), ('out_slope', C.c_float), ('dy_act', C.c_int32), ('dy_slope', C.c_float), ('dy_ref', _VP),
                ('bn_sums', C.c_void_p), ('prev_act', C.c_int32), ('prev_slope', C.c_float), ('prev_y', _VP),
                ('prev_scale', C.c_void_p), ('prev_shift', C.c_void_p), ('prev_mean', C.c_void_p), ('prev_invstd', C.c_void_p),
                ('prev_sums', C.c_void_p)]


_FP = C.POINTER(Fuse)
_vp, _i32, _i64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double

# name -> argtypes (every function returns int status unless noted); mirrors include/b200gan.h one to one
PROTOTYPES = {
    'b200gan_conv2d_fprop': [_CP, _VP, _vp, _vp, _VP, _FP, _vp],
    'b200gan_conv2d_dgrad': [_CP, _VP, _vp, _vp, _VP, _FP, _vp],
    'b200gan_conv2d_wgrad': [_CP, _VP, _VP, _vp, _vp, _FP, _vp],
    'b200gan_convT2d_fprop': [_CP, _VP, _vp, _vp, _VP, _FP, _vp],
    'b200gan_convT2d_dgrad': [_CP, _VP, _vp, _vp, _VP, _FP, _vp],
    'b200gan_convT2d_wgrad': [_CP, _VP, _VP, _vp, _vp, _FP, _vp],
    'b200gan_pack_conv_weight': [_vp, _i32, _i32, _i32, _i32, _vp, _vp],
    'b200gan_bn_stats': [_VP, _vp, _vp],
    'b200gan_bn_finalize': [_vp, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _vp, _vp, _vp, _vp, _vp],
    'b200gan_bn_finalize_act_fwd': [_vp, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _vp, _vp, _vp, _vp, _VP, _i32, _f32, _VP, _vp],
    'b200gan_bn_eval_coeffs': [_i32, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp],
    'b200gan_bn_act_fwd': [_VP, _vp, _vp, _i32, _f32, _VP, _vp],
    'b200gan_bn_act_bwd_reduce': [_VP, _VP, _VP, _vp, _vp, _vp, _vp, _i32, _f32, _vp, _vp],
    'b200gan_bn_act_bwd_apply': [_VP, _VP, _VP, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _VP, _vp, _vp, _vp],
    'b200gan_bce_sigmoid': [_vp, _i32, _f32, _f32, _vp, _vp, _vp, _vp],
    'b200gan_adam': [_vp, _vp, _vp, _vp, _i64, _f64, _f64, _f64, _f64, _i32, _vp, _f32, _vp],
    'b200gan_copy_view': [_VP, _VP, _vp],
    'b200gan_fill_f32': [_vp, _i64, _f32, _vp],
    'b200gan_gather_augment': [_vp, _i64, _vp, _vp, C.POINTER(C.c_float), C.POINTER(C.c_float), _VP, _vp],
    'b200gan_sample_sumsq': [_VP, _vp, _vp],
    'b200gan_gp_from_norms': [_vp, _i32, _f32, _vp, _vp, _vp],
    'b200gan_sample_axpby': [_VP, _vp, _VP, _vp, _VP, _vp],
    'b200gan_bn_bwd_bwd': [_VP, _VP, _VP, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _VP, _VP, _vp, _vp, _vp],
    'b200gan_mean_f32': [_vp, _i64, _f32, _vp, _vp],
    'b200gan_bce_logits': [_vp, _vp, _i32, _f32, _vp, _vp, _vp],
    'b200gan_fm_pair': [_VP, _VP, _VP, _f32, _i32, _vp, _vp],
    'b200gan_accumulate_2d': [_vp, _vp, _i32, _i32, _i32, _i64, _i64, _vp],
    'b200gan_conv3x3_fold': [_vp, _i32, _i32, _i32, _vp, _vp],
    'b200gan_bias_relu_d2s': [_VP, _vp, _VP, _vp],
    'b200gan_relu_bwd_s2d': [_VP, _VP, _VP, _vp],
    'b200gan_maxpool2_fwd': [_VP, _VP, _vp],
    'b200gan_maxpool2_bwd': [_VP, _VP, _VP, _i32, _vp],
    'b200gan_embed_add': [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp],
    'b200gan_embed_bwd': [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp],
    'b200gan_upconv3_fold': [_vp, _i32, _i32, _vp, _vp],
    'b200gan_upconv3_unfold': [_vp, _i32, _i32, _vp, _vp],
    'b200gan_class_proj_fwd': [_VP, _vp, _vp, _vp, _vp],
    'b200gan_class_proj_bwd': [_VP, _vp, _vp, _vp, _VP, _i32, _vp, _vp],
    'b200gan_dp_unique_id': [_vp],
    'b200gan_dp_init': [_vp, _i32, _i32, C.POINTER(_vp)],
    'b200gan_dp_allreduce_bucket': [_vp, _vp, _i64, _vp],
    'b200gan_dp_allreduce_f64': [_vp, _vp, _i64, _vp],
    'b200gan_dp_sync': [_vp, _vp],
    'b200gan_dp_destroy': [_vp],
}
OTHER_SYMBOLS = ['b200gan_version', 'b200gan_last_error_string', 'b200gan_device_info', 'b200gan_dp_collectives',
                 'b200gan_conv_wgrad_workspace_floats']
DP_ID_BYTES = 128

_lib = None


class B200GanError(RuntimeError):
    pass


def load():
    """Load libb200gan.so once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200GanError(
            f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            f'or `make -C {os.path.join(_HERE, "csrc")}`.  There is no CPU/cuDNN fallback for CUDA tensors.')
    lib = C.CDLL(LIB_PATH)
    for name, args in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.b200gan_version.restype = C.c_int
    lib.b200gan_last_error_string.restype = C.c_char_p
    lib.b200gan_device_info.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.b200gan_device_info.restype = C.c_int
    lib.b200gan_conv_wgrad_workspace_floats.argtypes = [_CP, _VP, _VP, _i32]
    lib.b200gan_conv_wgrad_workspace_floats.restype = C.c_int64
    lib.b200gan_dp_collectives.argtypes = [_vp]
    lib.b200gan_dp_collectives.restype = C.c_int64
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().b200gan_last_error_string().decode('utf-8', 'replace')
        raise B200GanError(f'{what} failed with status {rc}: {msg}')


def call(name, *args):
    check(getattr(load(), name)(*args), name)


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def view_nhwc(t):
    """View of a tensor stored as (N,H,W,C) (any strides)."""
    n, h, w, c = t.shape
    sn, sh, sw, sc = t.stride()
    return View(t.data_ptr(), _DTYPES[t.dtype], n, h, w, c, sn, sh, sw, sc)


def view_nchw(t):
    """View of a tensor stored as (N,C,H,W) (any strides) -- the reference's layout."""
    n, c, h, w = t.shape
    sn, sc, sh, sw = t.stride()
    return View(t.data_ptr(), _DTYPES[t.dtype], n, h, w, c, sn, sh, sw, sc)


def view_rows(t, n, h, w, c, row_period, row_offset=0):
    """View into a row-padded NHWC buffer t of shape (rows, W, C): image n, row h lives at buffer row
    row_offset + n*row_period + h (the zero rows in between are the convolution padding)."""
    sw, sc = c, 1
    sh = w * c
    base = t.data_ptr() + row_offset * sh * t.element_size()
    return View(base, _DTYPES[t.dtype], n, h, w, c, row_period * sh, sh, sw, sc)


def fuse(out_act=ACT_NONE, out_slope=0.0, dy_act=ACT_NONE, dy_slope=0.0, dy_ref=None, bn_sums=None, prev_act=ACT_NONE, prev_slope=0.0,
         prev_y=None, prev_scale=None, prev_shift=None, prev_mean=None, prev_invstd=None, prev_sums=None):
    """Build a b200gan_fuse.  dy_ref / prev_y are View objects (kept alive by the caller for the duration of the call);
    the remaining pointers are tensors."""
    def p(t):
        return t.data_ptr() if t is not None else None
    return Fuse(out_act, out_slope, dy_act, dy_slope, C.pointer(dy_ref) if dy_ref is not None else None, p(bn_sums), prev_act, prev_slope,
                C.pointer(prev_y) if prev_y is not None else None, p(prev_scale), p(prev_shift), p(prev_mean), p(prev_invstd), p(prev_sums))


def device_info(device=0):
    name = C.create_string_buffer(256)
    maj, mnr = C.c_int(0), C.c_int(0)
    sms = load().b200gan_device_info(device, name, C.byref(maj), C.byref(mnr))
    if sms < 0:
        check(sms, 'b200gan_device_info')
    return dict(name=name.value.decode(), sms=sms, cc=(maj.value, mnr.value))
