"""ctypes binding of libsrk.so (C ABI in include/srk.h).  No torch types cross the boundary:
only raw device pointers, sizes and the CUDA stream handle.

The library is required: importing this module without a built libsrk.so raises, and every
entry point raises RuntimeError (message from srk_last_error) on a non-zero return code.
There is no CPU or PyTorch fallback behind any of these calls."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SRK_LIB", os.path.join(_HERE, "..", "lib", "libsrk.so"))

F32, BF16 = 0, 1
LAYOUT_IMAGE, LAYOUT_ACT = 0, 1
ACT_NONE, ACT_RELU, ACT_PRELU = 0, 1, 2
IMPL_AUTO, IMPL_SIMT, IMPL_TC = 0, 1, 2
PACK_FPROP_SIMT, PACK_DGRAD_SIMT, PACK_FPROP_TC, PACK_DGRAD_TC, PACK_FPROP_TC_N8 = 0, 1, 2, 3, 4
PACK_RGBIN_TC, PACK_RGBOUT_DGRAD_TC = 5, 6


class SrkTensor(ctypes.Structure):
    _fields_ = [("data", c_void_p), ("layout", c_int32), ("dtype", c_int32),
                ("n", c_int32), ("c", c_int32), ("h", c_int32), ("w", c_int32)]


_T = POINTER(SrkTensor)
_P = c_void_p

# name -> (restype, argtypes); mirrors include/srk.h one to one
SIGNATURES = {
    "srk_last_error": (c_char_p, []),
    "srk_version": (c_int, []),
    "srk_conv_tc_supported": (c_int, [c_int] * 6),
    "srk_conv_fprop": (c_int, [_T, _T, _P, c_int, c_int, c_int, c_int, _P, c_int, _P, _T, c_int, c_int, _P, _P, _P, _T, _P, _P]),
    "srk_reduce_workspace_bytes": (c_int64, []),
    "srk_acc_bytes": (c_int64, []),
    "srk_acc_read": (c_int, [_P, c_int, _P, _P]),
    "srk_pixel_loss_scratch_bytes": (c_int64, []),
    "srk_conv_fprop_workspace_bytes": (c_int64, [_T, c_int]),
    "srk_conv_wgrad": (c_int, [_T, _T, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "srk_conv_rgbout_bwd_unshuffle": (c_int, [_T, _T, _T, _P, _T, _P, _P, _P, _P, c_int, _P, _P]),
    "srk_conv_wgrad_workspace_bytes": (c_int64, [_T, _T, c_int, c_int, c_int]),
    "srk_conv_rgb_workspace_bytes": (c_int64, [c_int]),
    "srk_conv_rgb_fprop": (c_int, [_T, _T, _P, c_int, _P, c_int, _P, _T, _P]),
    "srk_conv_rgb_bwd": (c_int, [_T, _T, _P, _T, _P, _P, c_int, c_int, _P, _P]),
    "srk_weight_pack": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "srk_weight_pack_bytes": (c_int64, [c_int] * 5),
    "srk_weight_pack_multi": (c_int, [c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    "srk_act_bwd": (c_int, [_T, _T, _T, _T, c_int, _P, _P, c_int, c_int, _P, _P]),
    "srk_bn_stats": (c_int, [_T, _P, _P, _P]),
    "srk_bn_finalize": (c_int, [_P, _P, c_int, c_int64, c_float, c_float, _P, _P, _P, _P, _P, _P]),
    "srk_bn_eval_params": (c_int, [_P, _P, c_int, c_float, _P, _P, _P]),
    "srk_bn_apply": (c_int, [_T, _P, _P, _P, _P, _P, _T, _T, _P]),
    "srk_bn_apply_train": (c_int, [_T, _P, _P, c_int64, c_float, c_float, _P, _P, _P, _P, _P, _P, _P, _P, _T, _T, _P, _P]),
    "srk_bn_bwd_reduce": (c_int, [_T, _T, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "srk_bn_bwd_apply": (c_int, [_T, _T, _P, _P, _P, _P, _P, _P, _P, c_int, _T, _P]),
    "srk_bn_bwd_apply_raw": (c_int, [_T, _T, _P, _P, _P, _P, _P, _P, _P, c_int, _P, _T, _P, _P, _P, _P]),
    "srk_conv_dgrad_bnred": (c_int, [_T, _T, _P, _T, _P, _P, _P, _P, _P, _P, _P, _P, _T, _P, _P, _P]),
    "srk_se_pool": (c_int, [_T, _P, _P, _P]),
    "srk_se_fc": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P, _P]),
    "srk_se_apply": (c_int, [_T, _T, _P, c_float, _T, _P]),
    "srk_se_bwd_reduce": (c_int, [_T, _T, _P, _P, _P]),
    "srk_se_fc_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, _P, _P, _P, _P]),
    "srk_se_bwd_apply": (c_int, [_T, _P, _P, c_float, _T, _P]),
    "srk_image_to_act": (c_int, [_T, _T, _P]),
    "srk_act_to_image": (c_int, [_T, _T, _P]),
    "srk_act_add": (c_int, [_T, _T, _T, _P]),
    "srk_maxpool2_fwd": (c_int, [_T, _T, _P]),
    "srk_maxpool2_bwd": (c_int, [_T, _T, _T, _P]),
    "srk_bicubic_upsample": (c_int, [_T, _T, _P]),
    "srk_pixel_loss_fwd": (c_int, [_P, _P, c_int64, c_int, _P, _P, _P]),
    "srk_pixel_loss_bwd": (c_int, [_P, _P, c_int64, c_int, _P, _P, _P]),
    "srk_nlpd_workspace_bytes": (c_int64, [c_int] * 5),
    "srk_nlpd_fwd": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_float, _P, c_int, _P, _P, _P]),
    "srk_nlpd_bwd": (c_int, [c_int, c_int, c_int, c_int, c_int, c_float, _P, _P, _P, _P, _P]),
    "srk_psnr_sse": (c_int, [_P, _P, c_int, c_int64, c_int, _P, _P]),
    "srk_ssim": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "srk_adam_step": (c_int, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, _P, c_float, _P]),
    "srk_adam_multi": (c_int, [c_int, _P, _P, _P, _P, _P, c_float, c_float, c_float, c_float, _P, c_float, _P, _P, _P]),
    "srk_sr_make_batch": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, c_int, c_int, _P, _P, _P]),
    "srk_tc_probe": (c_int, [c_int, POINTER(c_float), c_int]),
}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libsrk.so not found at %s - build it with `python food101-super-resolution_b200/build.py` "
        "(there is no CPU / PyTorch fallback for the SR hot path)" % LIB_PATH)

cdll = ctypes.CDLL(os.path.abspath(LIB_PATH))
for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(cdll, _name)
    _f.restype = _res
    _f.argtypes = _args

# number of libsrk kernel-launching calls made by this process (bench.py reports it)
launch_calls = 0
_NO_COUNT = {"srk_last_error", "srk_version", "srk_conv_tc_supported", "srk_weight_pack_bytes",
             "srk_conv_wgrad_workspace_bytes", "srk_nlpd_workspace_bytes", "srk_conv_rgb_workspace_bytes",
             "srk_conv_fprop_workspace_bytes", "srk_reduce_workspace_bytes", "srk_pixel_loss_scratch_bytes",
             "srk_acc_bytes"}


def last_error():
    return cdll.srk_last_error().decode("utf-8", "replace")


def call(name, *args):
    """Calls an int-returning entry point and raises on a non-zero status."""
    global launch_calls
    rc = getattr(cdll, name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed: %s" % (name, last_error()))
    if name not in _NO_COUNT:
        launch_calls += 1
    return rc
