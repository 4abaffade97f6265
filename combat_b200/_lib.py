"""ctypes binding of the C-ABI in include/combat_b200.h (libcombat_b200.so, built in-tree).

There is no CPU fallback: if the shared library is missing or a symbol declared in the
header is not exported, importing this module raises.  Every wrapper raises RuntimeError
on a non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("COMBAT_LIB") or os.path.join(_HERE, "libcombat_b200.so")   # COMBAT_LIB: A/B builds of the same C-ABI
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "combat_b200.h")

vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_longlong, C.c_float


class WPrepDesc(C.Structure):
    _fields_ = [("src_off", i64), ("fwd_off", i64), ("dgrad_off", i64),
                ("Cout", i32), ("Cin", i32), ("KH", i32), ("KW", i32)]


class ConvDesc(C.Structure):
    _fields_ = [("in_", vp), ("w", vp), ("out", vp), ("bias", vp), ("residual", vp), ("post_scale", vp), ("post_shift", vp),
                ("N", i32), ("Hi", i32), ("Wi", i32), ("Ci", i32), ("Ho", i32), ("Wo", i32), ("Co", i32),
                ("KH", i32), ("KW", i32), ("stride", i32), ("pad", i32), ("up", i32),
                ("in_sn", i64), ("in_sh", i64), ("in_sw", i64), ("in_sc", i64),
                ("out_sn", i64), ("out_sh", i64), ("out_sw", i64), ("out_sc", i64),
                ("in_dtype", i32), ("w_dtype", i32), ("out_dtype", i32), ("act", i32)]


class ConvTcDesc(C.Structure):
    _fields_ = [("in_", vp), ("w", vp), ("out", vp), ("bias", vp), ("residual", vp), ("stats", vp),
                ("N", i32), ("Hi", i32), ("Wi", i32), ("Ci", i32), ("Ho", i32), ("Wo", i32), ("Co", i32),
                ("KH", i32), ("KW", i32), ("stride", i32), ("pad", i32), ("up", i32), ("out_f32", i32), ("res_f32", i32),
                ("act", i32), ("post_scale", vp), ("post_shift", vp),
                ("out2", vp), ("scale2", vp), ("shift2", vp), ("mask", vp), ("mask_scale", vp), ("post_add", vp),
                ("in2", vp), ("w2", vp),
                ("bnb_x", vp), ("bnb_scale", vp), ("bnb_shift", vp), ("bnb_mean", vp), ("bnb_invstd", vp), ("in_nchw3", i32)]


P = C.POINTER
_SIGS = {
    "combat_version": ([], i32),
    "combat_launch_count": ([], i64),
    "combat_last_error": ([], C.c_char_p),
    "combat_plane_transform": ([vp, vp, vp, vp, i64, i32, i32, vp, vp], i32),
    "combat_dct32_fast": ([vp, vp, i64, i32, i32, i32, vp], i32),
    "combat_dct64_fast": ([vp, vp, i64, i32, i32, i32, vp], i32),
    "combat_poison_blend_fwd": ([vp, vp, vp, vp, i32, i32, f32, f32, f32, vp, vp, i32, i32, i32, vp, vp, vp, vp], i32),
    "combat_poison_blend_bwd": ([vp, vp, vp, vp, vp, f32, f32, f32, f32, vp, i32, i32, i32, i32, vp, vp, vp], i32),
    "combat_cross_entropy": ([vp, vp, vp, i32, i32, f32, vp, vp, vp, vp], i32),
    "combat_sum_scale": ([vp, i32, f32, vp, vp], i32),
    "combat_sgd_nesterov": ([vp, vp, vp, i64, vp, f32, f32, i32, vp], i32),
    "combat_prep_weights": ([vp, vp, i32, vp, i32, i64, vp], i32),
    "combat_conv_simt": ([P(ConvDesc), vp], i32),
    "combat_conv_wgrad_simt": ([P(ConvDesc), vp, i32, vp, vp], i32),
    "combat_im2col3": ([vp, vp, i32, i32, i32, i32, vp], i32),
    "combat_fold_w64": ([vp, vp, i32, i32, vp, vp, vp], i32),
    "combat_conv_cin3": ([vp, vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp], i32),
    "combat_conv_cout3": ([vp, i32, vp, i32, vp, vp, i32, i32, i32, i32, i32, vp], i32),
    "combat_wgrad_cin3": ([vp, vp, i32, vp, vp, i32, i32, i32, i32, i32, vp], i32),
    "combat_wgrad_cout3": ([vp, i32, vp, vp, vp, i32, i32, i32, i32, vp], i32),
    "combat_conv_tc": ([P(ConvTcDesc), vp], i32),
    "combat_conv_tc_wgrad": ([P(ConvTcDesc), vp, vp, vp], i32),
    "combat_conv_tc_supported": ([P(ConvTcDesc)], i32),
    "combat_conv_tc_last_grid": ([], i32),
    "combat_conv_tc_cout3": ([vp, vp, vp, vp, i32, i32, i32, i32, vp], i32),
    "combat_conv_tc_cout3_supported": ([i32, i32, i32], i32),
    "combat_bn_stats": ([vp, i32, i64, i32, vp, i32, P(i32), vp], i32),
    "combat_bn_finalize": ([vp, i32, i64, i32, vp, vp, vp, vp, f32, f32, vp, vp, vp, vp, vp], i32),
    "combat_bn_eval_affine": ([vp, vp, vp, i32, f32, vp, vp, vp], i32),
    "combat_affine_act": ([vp, i32, vp, vp, i32, i64, i32, vp, vp, i32, vp], i32),
    "combat_bn_bwd_reduce": ([vp, vp, i32, vp, i32, i64, i32, vp, vp, vp, i32, P(i32), i32, vp], i32),
    "combat_bn_bwd_finalize": ([vp, i32, i32, vp, vp, vp], i32),
    "combat_bn_bwd_apply": ([vp, vp, i32, vp, vp, vp, vp, i32, i64, i32, vp, vp, vp, vp, vp, vp, i32, vp], i32),
    "combat_bn_bwd_fused": ([vp, vp, i32, vp, vp, vp, vp, i32, i64, i32, vp, vp, vp, vp, i32, vp, vp, i32, vp], i32),
    "combat_instnorm_fwd": ([vp, i32, vp, vp, i32, i32, i32, i32, f32, f32, i32, vp, vp, vp], i32),
    "combat_instnorm_bwd": ([vp, vp, vp, i32, vp, i32, i32, i32, i32, f32, i32, vp, vp, vp], i32),
    "combat_instnorm_splits": ([i32, i32, i32], i32),
    "combat_instnorm_fwd_split": ([vp, i32, vp, vp, i32, i32, i32, i32, f32, f32, i32, vp, vp, vp, i32, vp], i32),
    "combat_instnorm_bwd_split": ([vp, vp, vp, i32, vp, i32, i32, i32, i32, f32, i32, vp, vp, vp, i32, vp], i32),
    "combat_upsample2x_act": ([vp, vp, i32, i32, i32, i32, i32, f32, vp], i32),
    "combat_upsample2x_act_bwd": ([vp, vp, vp, i32, i32, i32, i32, i32, f32, vp], i32),
    "combat_leaky_relu": ([vp, vp, i32, i64, f32, vp], i32),
    "combat_leaky_relu_bwd": ([vp, vp, vp, i32, i64, f32, vp], i32),
    "combat_tanh_bwd": ([vp, vp, vp, i64, vp], i32),
    "combat_colsum": ([vp, i32, i64, i32, vp, vp], i32),
    "combat_pool_linear_fwd": ([vp, i32, i32, i32, i32, i32, i32, vp, vp, i32, vp, vp, vp], i32),
    "combat_pool_linear_bwd": ([vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, i32, vp, vp, vp], i32),
    "combat_maxpool2": ([vp, vp, i32, i32, i32, i32, i32, vp], i32),
    "combat_elu_bwd": ([vp, vp, vp, i32, i64, vp], i32),
    "combat_maxpool2_bwd": ([vp, vp, vp, i32, i32, i32, i32, i32, vp], i32),
    "combat_mask_scale": ([vp, vp, vp, i32, i64, f32, vp], i32),
    "combat_adadelta": ([vp, vp, vp, vp, i64, vp, f32, f32, f32, vp], i32),
    "combat_grad_l2": ([vp, vp, vp, vp, i32, i32, i32, i32, vp], i32),
    "combat_tv_loss": ([vp, vp, f32, vp, vp, i32, i32, i32, i32, vp], i32),
    "combat_post_transform_fwd": ([vp, vp, vp, i32, i32, i32, i32, vp], i32),
    "combat_post_transform_bwd": ([vp, vp, vp, i32, i32, i32, i32, i32, vp], i32),
    "combat_tanh_fwd": ([vp, vp, i64, vp], i32),
    "combat_wanet_warp_fwd": ([vp, vp, vp, vp, i32, i32, vp, f32, vp, vp, vp, vp, i32, i32, i32, i32, vp], i32),
    "combat_wanet_warp_bwd": ([vp, vp, vp, vp, vp, f32, f32, vp, i32, i32, i32, i32, i32, vp], i32),
    "combat_nchw_to_nhwc": ([vp, vp, i32, i32, i32, i32, i32, vp], i32),
    "combat_nhwc_to_nchw": ([vp, i32, vp, i32, i32, i32, i32, vp], i32),
    "combat_onehot_planes": ([vp, i32, vp, i32, i32, i32, i32, i32, vp], i32),
    "combat_lrelu_into_slice": ([vp, vp, i32, i64, i32, i32, i32, f32, vp], i32),
}


def header_symbols(path: str = HEADER_PATH) -> list[str]:
    """Every function name declared in include/combat_b200.h."""
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(combat_[a-z0-9_]+)\s*\(", txt)))


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "combat_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (args, res) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = args
        fn.restype = res
    return lib


lib = _load()


class CombatError(RuntimeError):
    pass


def check(rc: int, name: str):
    if rc != 0:
        raise CombatError("%s failed rc=%d (%s)" % (name, rc, lib.combat_last_error().decode()))


def launch_count() -> int:
    return int(lib.combat_launch_count())
