"""ctypes binding of librovr_b200.so (include/rovr_b200.h).

There is deliberately no fallback: if the library is missing the import raises, and if the device
is not an sm_100 GPU every compute call raises RuntimeError with the library's error text.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librovr_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing. Build it with `python __graft_entry__.py build` "
        "(nvcc, sm_100a). The ROVR B200 path has no CPU or PyTorch fallback."
    )

lib = ctypes.CDLL(LIB_PATH)

_p = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float
_sz = ctypes.c_size_t
_ll = ctypes.c_longlong

# name -> (restype, argtypes); must list every symbol declared in include/rovr_b200.h
SIGNATURES = {
    "rovr_abi_version": (_i, []),
    "rovr_last_error": (ctypes.c_char_p, []),
    "rovr_device_check": (_i, []),
    "rovr_hang_code": (_i, [ctypes.POINTER(ctypes.c_uint)]),
    "rovr_launch_count": (ctypes.c_ulonglong, []),
    "rovr_pack_nchw_to_nhwc": (_i, [_p, _i, _p, _i, _p, _i, _p, _i, _i, _i, _i, _p]),
    "rovr_unpack_nhwc_to_nchw": (_i, [_p, _i, _p, _i, _i, _i, _i, _p]),
    "rovr_u8_to_f32": (_i, [_p, _p, _ll, _f, _p]),
    "rovr_corrupt_frames": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "rovr_repack_conv3x3_fprop": (_i, [_p, _p, _i, _i, _i, _p]),
    "rovr_repack_conv3x3_dgrad": (_i, [_p, _p, _i, _i, _i, _p]),
    "rovr_repack_convT2x2_fprop": (_i, [_p, _p, _i, _i, _p]),
    "rovr_repack_convT2x2_dgrad": (_i, [_p, _p, _i, _i, _p]),
    "rovr_repack_linear": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "rovr_repack_batch": (_i, [_p, _i, _p]),
    "rovr_gemm_wgrad_workspace": (_sz, [_ll, _i, _i]),
    "rovr_gemm_wgrad": (_i, [_p, _i, _p, _i, _p, _ll, _i, _i, _i, _i, _p, _sz, _p]),
    "rovr_conv3x3_fprop": (_i, [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "rovr_conv3x3_fprop_pool2": (_i, [_p, _i, _p, _p, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "rovr_conv3x3_fprop_tail_workspace": (_sz, [_i, _i, _i]),
    "rovr_conv3x3_fprop_tail": (_i, [_p, _i, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _sz, _i, _i, _i, _i, _i, _p]),
    "rovr_dgrad_colsum_workspace": (_sz, [_i, _i, _i, _i]),
    "rovr_conv3x3_dgrad": (_i, [_p, _i, _p, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p, _sz, _p]),
    "rovr_conv3x3_wgrad_workspace": (_sz, [_i, _i, _i, _i, _i]),
    "rovr_conv3x3_wgrad": (_i, [_p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _p, _sz, _p]),
    "rovr_convT2x2_fprop": (_i, [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "rovr_convT2x2_dgrad": (_i, [_p, _i, _p, _p, _i, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p, _sz, _p]),
    "rovr_convT2x2_wgrad_workspace": (_sz, [_i, _i, _i, _i, _i]),
    "rovr_convT2x2_wgrad": (_i, [_p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _p, _sz, _p]),
    "rovr_gemm_bf16": (_i, [_p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "rovr_maxpool_fwd": (_i, [_p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "rovr_maxpool_bwd_colsum_workspace": (_sz, [_i, _i, _i, _i, _i, _i]),
    "rovr_maxpool_bwd": (_i, [_p, _i, _p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "rovr_tail_workspace": (_sz, [_i, _i, _i]),
    "rovr_tail_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _sz, _i, _i, _i, _p]),
    "rovr_tail_bwd": (_i, [_p, _p, _p, _p, _p, _f, _p, _p, _p, _p, _p, _p, _sz, _i, _i, _i, _p]),
    "rovr_bn_workspace": (_sz, [_i]),
    "rovr_bn_train_fwd": (_i, [_p, _i, _p, _i, _ll, _i, _i, _p, _p, _f, _f, _p, _p, _p, _p, _p, _i, _p, _sz, _p]),
    "rovr_bn_train_bwd": (_i, [_p, _i, _p, _i, _p, _i, _p, _i, _ll, _i, _i, _p, _p, _p, _p, _p, _i, _p, _sz, _p]),
    "rovr_bn_eval_fwd": (_i, [_p, _i, _p, _i, _ll, _i, _i, _p, _p, _f, _p, _p, _p, _i, _p]),
    "rovr_bn_frames_workspace": (_sz, [_i, _i, _ll]),
    "rovr_bn_train_fwd_frames": (_i, [_p, _i, _i, _p, _i, _i, _ll, _i, _i, _p, _p, _f, _f, _p, _p, _p, _p, _p, _i, _p, _sz,
                                      _p]),
    "rovr_bn_eval_bwd": (_i, [_p, _i, _p, _i, _p, _i, _p, _i, _ll, _i, _i, _p, _p, _p, _p, _p, _i, _p, _sz, _p]),
    "rovr_layernorm_fwd": (_i, [_p, _ll, _i, _f, _p, _p, _p, _p, _p, _p, _p]),
    "rovr_layernorm_workspace": (_sz, [_i]),
    "rovr_layernorm_bwd": (_i, [_p, _p, _ll, _i, _p, _p, _p, _p, _i, _p, _p, _p, _sz, _p]),
    "rovr_linear_f32_fwd": (_i, [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "rovr_linear_f32_dgrad_workspace": (_sz, [_i, _i, _i]),
    "rovr_linear_f32_dgrad": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _p, _sz, _p]),
    "rovr_linear_f32_wgrad": (_i, [_p, _i, _p, _i, _p, _p, _i, _i, _i, _p]),
    "rovr_standardize_fwd": (_i, [_p, _p, _p, _i, _i, _ll, _ll, _f, _p]),
    "rovr_standardize_bwd": (_i, [_p, _p, _p, _p, _i, _i, _ll, _ll, _f, _p]),
    "rovr_head_mask_std_fwd": (_i, [_p, _p, _i, _p, _p, _i, _i, _i, _p]),
    "rovr_head_mask_std_bwd": (_i, [_p, _p, _p, _p, _p, _i, _p, _i, _i, _i, _p]),
    "rovr_head_gumbel_fwd": (_i, [_p, _p, _f, _p, _i, _i, _i, _p, _p, _p, _p]),
    "rovr_head_gumbel_bwd": (_i, [_p, _p, _f, _i, _i, _i, _p, _p, _p]),
    "rovr_lstm_pointwise_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _p]),
    "rovr_lstm_pointwise_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "rovr_flatten_nhwc": (_i, [_p, _i, _p, _i, _i, _i, _i, _p]),
    "rovr_unflatten_nhwc": (_i, [_p, _i, _p, _i, _i, _i, _i, _i, _p]),
    "rovr_copy2d_f32": (_i, [_p, _i, _p, _i, _i, _i, _f, _i, _p]),
    "rovr_fold_bn": (_i, [_p, _p, _p, _p, _p, _f, _p, _p, _i, _i, _p]),
    "rovr_stem_im2col": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "rovr_maxpool_pad_fwd": (_i, [_p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "rovr_subsample": (_i, [_p, _i, _p, _i, _i, _i, _i, _i, _i, _p]),
    "rovr_add_relu": (_i, [_p, _p, _p, _ll, _p]),
    "rovr_avgpool": (_i, [_p, _i, _p, _i, _i, _i, _p]),
    "rovr_mosaic_paste": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "rovr_resize_antialias": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "rovr_gemm_batched_bf16": (_i, [_p, _ll, _ll, _ll, _p, _ll, _ll, _ll, _p, _p, _ll, _ll, _ll, _i, _i, _i, _i, _i, _i, _p]),
    "rovr_transpose_bf16": (_i, [_p, _ll, _ll, _ll, _p, _ll, _ll, _ll, _i, _i, _i, _i, _i, _p]),
    "rovr_softmax_fwd": (_i, [_p, _p, _ll, _i, _i, _f, _p]),
    "rovr_softmax_bwd": (_i, [_p, _p, _p, _ll, _i, _i, _f, _p]),
    "rovr_gelu_fwd": (_i, [_p, _p, _ll, _p]),
    "rovr_gelu_bwd": (_i, [_p, _p, _p, _ll, _p]),
    "rovr_cast_f32_bf16": (_i, [_p, _p, _ll, _p]),
    "rovr_add_f32": (_i, [_p, _p, _p, _ll, _p]),
    "rovr_set_pair_mode": (_i, [_i]),
    "rovr_host_selftest": (_i, []),
    "rovr_posenc_add": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "rovr_posenc_grad": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "rovr_colsum_workspace": (_sz, [_i]),
    "rovr_colsum_rows_workspace": (_sz, [_i]),
    "rovr_colsum_rows": (_i, [_p, _ll, _ll, _i, _p, _p, _sz, _p]),
    "rovr_colsum": (_i, [_p, _i, _ll, _i, _p, _p, _sz, _p]),
    "rovr_split_stack": (_i, [_p, _p, _i, _p, _ll, _ll, _ll, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _i, _i, _i, _i, _p]),
    "rovr_split_weights": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "rovr_blocksum4": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "rovr_conv3x3_fprop_s2": (_i, [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "rovr_conv3x3_fprop_s2_f32out": (_i, [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "rovr_conv3x3_f32out": (_i, [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "rovr_convT2x2_fprop_f32out": (_i, [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "rovr_convT2x2_dgrad_f32out": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "rovr_bn_f32_train_fwd": (_i, [_p, _i, _p, _i, _ll, _i, _i, _p, _p, _f, _f, _p, _p, _p, _p, _p, _i, _p, _sz, _p]),
    "rovr_bn_f32_eval_fwd": (_i, [_p, _i, _p, _i, _ll, _i, _i, _p, _p, _f, _p, _p, _p, _i, _p]),
    "rovr_bn_f32_bwd": (_i, [_p, _i, _p, _i, _p, _i, _p, _i, _ll, _i, _i, _p, _p, _p, _p, _p, _i, _i, _p, _sz, _p]),
    "rovr_colsum_f32": (_i, [_p, _i, _ll, _i, _p, _p, _sz, _p]),
    "rovr_maxpool_f32_fwd": (_i, [_p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "rovr_maxpool_f32_bwd": (_i, [_p, _i, _p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "rovr_flatten_f32": (_i, [_p, _i, _p, _ll, _i, _i, _i, _p]),
    "rovr_unflatten_f32": (_i, [_p, _ll, _p, _i, _i, _i, _i, _i, _p]),
    "rovr_dropout": (_i, [_p, _p, _ll, _i, _f, _p, _i, _p]),
    "rovr_dropout_advance": (_i, [_p, _p]),
    "rovr_lpips_pack": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _i, _p]),
    "rovr_lpips_head_blocks": (_i, [_i, _ll]),
    "rovr_lpips_pack_f32": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _i, _p]),
    "rovr_lpips_head_f32": (_i, [_p, _i, _ll, _i, _p, _p, _p, _i, _p]),
    "rovr_lpips_unpack_grad_f32": (_i, [_p, _i, _p, _p, _i, _i, _i, _p, _p, _i, _p]),
    "rovr_lpips_head": (_i, [_p, _i, _ll, _i, _p, _p, _p, _i, _p]),
    "rovr_lpips_finalize": (_i, [_p, _p, _p, _i, _i, _p, _p]),
    "rovr_lpips_unpack_grad": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _i, _p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == header/library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


class RovrError(RuntimeError):
    pass


def last_error():
    return lib.rovr_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise RovrError(f"{what or 'librovr_b200'} failed (rc={rc}): {last_error()}")


def call(name, *args):
    """Invoke an int-returning entry point and raise on a non-zero status."""
    check(getattr(lib, name)(*args), name)


def require_device():
    check(lib.rovr_device_check(), "rovr_device_check")


def hang_code():
    c = ctypes.c_uint(0)
    lib.rovr_hang_code(ctypes.byref(c))
    return c.value
