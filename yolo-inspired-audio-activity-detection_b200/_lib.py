"""ctypes binding of the C-ABI library (include/yad_b200.h).

The product path has no CPU fallback: if ``libyad_b200.so`` is missing, or the device is not
sm_100, every call raises.  ``__graft_entry__.build()`` / ``build.sh`` produce the library in-tree.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libyad_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2
FE_QW = 21


class YadError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "B", "H", "W", "Cin", "ld_in", "Cout", "ld_out", "co_off", "kh", "kw", "sh", "sw", "ph", "pw", "act", "ld_res",
        "in_sw", "in_sh", "in_sb", "out_sw", "out_sh", "out_sb")]


class CorrDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "B", "H", "W", "Cin", "ld_in", "Ho", "Wo", "Cout", "ld_out", "sh", "sw", "out_sw", "out_sh", "out_sb", "n_taps", "act",
        "accumulate", "whole_rows")]


class FlatDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "B", "H", "W", "Hp", "Wp", "Cin", "ld_in", "Cout", "ld_out", "co_off", "kh", "kw", "ph", "pw", "act", "ld_res",
        "Hp_out", "Wp_out")]


_p, _i32, _i64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double

# name -> argument types (all return int)
SIGNATURES = {
    "yad_init": [C.c_int],
    "yad_frontend_mel_power": [_p, _i64, _i64, _i32, _i32, _i32, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _i32, _p, _i64, _p],
    "yad_frontend_mel_power_taper": [_p, _i32, _p, _i64, _i64, _i64, _i32, _i32, _i32, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _i32, _p, _i64, _p],
    "yad_frontend_mel_power_i16": [_p, _i64, _i64, _i32, _i32, _i32, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _i32, _p, _i64, _p],
    "yad_frontend_finish": [_p, _i64, _i64, _p, _f32, _i32, _p, _p, _p, _p, _p],
    "yad_conv_stem": [_p, _i64, _i32, _i32, _p, _p, _i32, _p],
    "yad_conv_stem_tc": [_p, _i64, _i32, _i32, _p, _p, _i32, _i32, _p],
    "yad_conv_simt": [C.POINTER(ConvDesc), _i32, _p, _p, _i32, _p, _p, _p, _p],
    "yad_conv_tc": [C.POINTER(ConvDesc), _p, _p, _i32, _p, _p, _p, _i32, _p, _i32, _p],
    "yad_conv_tc_dual": [C.POINTER(ConvDesc), _p, _p, _i32, _p, _p, _p, _p, _i32, _p, _p],
    "yad_conv_flat": [C.POINTER(FlatDesc), _p, _p, _i32, _p, _p, _p, _i32, _p],
    "yad_conv_flat_taps2": [C.POINTER(FlatDesc), _i32, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32),
                            _i64, _p, _p, _i32, _i32, _p, _i32, _p, _p, _p, _p],
    "yad_conv_flat_s2d": [C.POINTER(FlatDesc), _p, _p, _i32, _p, _p, _p, _p, _i32, _i32, _p],
    "yad_conv_flat_taps": [C.POINTER(FlatDesc), _i32, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), _i64, _p, _p,
                           _i32, _p, _p, _p, _p],
    "yad_resample_sinc": [_p, _i32, _i64, _i64, _i32, _i32, _i32, _p, _p, _i64, _p],
    "yad_hmean": [_p, _i32, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _i32, _i32, _p],
    "yad_resize_w": [_p, _i32, _i64, _i32, _i32, _i32, _i32, _i32, _p, _i32, _i32, _p],
    "yad_neck_fused": [C.POINTER(_p), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), _i64, _i32, _p, _i64, _p, _i32, _p, _i32, _p, _i32, _i32, _i32,
                       C.POINTER(_p), C.POINTER(_i32), _i32, _p, _p],
    "yad_neck_fused_set_timeline": [_p],
    "yad_neck_fused_set_timeline_iter": [_i32],
    "yad_conv_flat_set_timeline": [_p],
    "yad_sppf_pools": [_p, _i32, _i64, _i32, _i32, _i32, _i32, _p, _i32, _i32, _p],
    "yad_conv_stem_fused": [_p, _i64, _i64, _i32, _i32, _p, _p, _p, _i32, _i32, _i32, _p],
    "yad_frontend_finish_bf16": [_p, _i64, _i64, _p, _f32, _i32, _p, _p, _p, _p, _p, _i64, _i32, _p],
    "yad_conv_stem_fused_skip": [_p, _i64, _i64, _i32, _i32, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _p],
    "yad_conv_stem_fused_fixup_bf16": [_p, _i64, _i64, _i32, _i32, _p, _p, _p, _p, _i32, _p, _i32, _i32, _p],
    "yad_conv_stem_fused_fixup": [_p, _i64, _i32, _i32, _p, _p, _p, _p, _i32, _p, _i32, _i32, _p],
    "yad_kmeans1d_lloyd": [_p, _i64, _p, _i32, _i32, _f64, _p, _p, _p, _p],
    "yad_maxpool_h": [_p, _i32, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _p, _i32, _i32, _p],
    "yad_nchw_to_nhwc": [_p, _i64, _i32, _i32, _i32, _p, _i32, _i32, _p],
    "yad_repvgg_merge": [_p, _p, _p, _p, _p, _i32, _i64, _i32, _i32, _i32, _p, _i32, _i32, _i32, _p],
    "yad_decode": [C.POINTER(_p), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), _i32, _i32, C.POINTER(_f32), _i32,
                   _i32, _f32, _f32, _i64, _p, _p],
    "yad_decode_dev": [C.POINTER(_p), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), _i32, _i32, _p, _i32,
                       _i32, _f32, _f32, _i64, _p, _p],
    "yad_nms": [_p, _i64, _i32, _i32, _f64, _f32, _f32, _f32, _i32, _p, _p, _p, _p, _p, _p, _p],
    "yad_compact_segments": [_p, _p, _i64, _i32, _p, _p, _p, _p],
    "yad_build_targets": [_p, _i32, _p, _i32, _i32, _f32, _f32, _f32, _p, _p, _p, _p, _p, _p, _p],
    "yad_collate_clips": [_p, _i32, _p, _p, _p, _i64, _i64, _p, _p],
    "yad_loss_scale_ex": [_p, _i64, _i32, _i32, _i32, _p, _p, _p, _p, _p, _i32, _f32, _f32, _f32, _f32, _i64, _i32, _p, _f32, _f32, _p, _p, _p,
                          _p, _p, _p],
    "yad_loss_scale": [_p, _i64, _i32, _i32, _i32, _p, _p, _p, _p, _p, _i32, _f32, _f32, _f32, _f32, _i64, _p, _p, _p, _p, _p, _p],
    "yad_conv_dgrad": [C.POINTER(ConvDesc), _p, _p, _p, _p, _p],
    "yad_conv_wgrad": [C.POINTER(ConvDesc), _p, _p, _p, _p, _p],
    "yad_bn_train_fwd": [_p, _i32, _i64, _i32, _p, _p, _f32, _f32, _p, _p, _i32, _p, _i32, _p, _p, _p, _p],
    "yad_bn_train_apply": [_p, _i32, _i64, _i32, _p, _p, _f32, _f32, _p, _p, _i32, _p, _i32, _p, _p, _p, _p],
    "yad_bn_train_bwd": [_p, _i32, _p, _i32, _p, _i32, _i64, _i32, _p, _p, _p, _i32, _p, _i32, _i32, _p, _p, _p, _p],
    "yad_corr_tf32": [C.POINTER(CorrDesc), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), _p, _p, _i32, _i64, _p, _p, _p, _p],
    "yad_wgrad_tf32": [C.POINTER(CorrDesc), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), _p, _p, _p, _p],
    "yad_stem_im2col": [_p, _i64, _i32, _i32, _i32, _i32, _p, _p],
    "yad_colsum_f64": [_p, _i32, _i64, _i32, _p, _p],
    "yad_pack_weights_tf32": [_p, _i32, _i64, _p],
    "yad_permute4": [_p, C.POINTER(_i64), _p, C.POINTER(_i32), _i32, _p],
    "yad_add_f64_to_f32": [_p, _i32, _p, _p],
    "yad_add_act": [_p, _i32, _p, _i32, _p, _i32, _i64, _i32, _i32, _p, _i32, _p],
    "yad_add_act_bwd": [_p, _i32, _p, _i32, _i64, _i32, _i32, _p, _i32, _p, _i32, _p, _i32, _p],
    "yad_dropout": [_p, _i64, _f32, C.c_uint64, _i32, _p, _p],
    "yad_dropout_dev": [_p, _i64, _f32, C.c_uint64, _p, _i32, _p, _p],
    "yad_hmean_bwd": [_p, _i32, _i64, _i32, _i32, _i32, _p, _i32, _p],
    "yad_resize_w_bwd": [_p, _i32, _i64, _i32, _i32, _i32, _p, _i32, _p],
    "yad_maxpool5_w": [_p, _i32, _i64, _i32, _i32, _p, _i32, _p],
    "yad_maxpool5_w_bwd": [_p, _i32, _p, _i32, _i64, _i32, _i32, _p, _i32, _p],
    "yad_decode_bwd": [_p, _i32, _p, _i64, _i32, _i32, _i32, _p, _f32, _f32, _p, _i32, _p, _p],
    "yad_adam_ema_step": [_p, _p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _f32, _p],
}
EXPORTS = ["yad_version", "yad_last_error"] + list(SIGNATURES)

_lib = None
_lock = threading.Lock()
_inited = set()


def load() -> C.CDLL:
    """dlopen the library and attach prototypes.  Raises YadError if it was not built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise YadError(
                    f"{LIB_PATH} is missing: build it with ./build.sh (nvcc, sm_100a). "
                    "There is no CPU or PyTorch fallback for this path.")
            lib = C.CDLL(LIB_PATH)
            lib.yad_version.restype = C.c_int
            lib.yad_last_error.restype = C.c_char_p
            for name, args in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.argtypes = args
                fn.restype = C.c_int
            _lib = lib
    return _lib


param_epoch = 0      # bumped whenever a kernel rewrites parameters behind autograd's version counters (FusedAdamEMA.step)
launch_count = 0     # kernels launched through the C ABI by this process (bench.py reports it)


def check(rc: int, what: str):
    global launch_count
    if rc != 0:
        raise YadError(f"{what} failed (code {rc}): {load().yad_last_error().decode()}")
    launch_count += 1


def init(device_index: int):
    lib = load()
    if device_index not in _inited:
        rc = lib.yad_init(int(device_index))
        if rc != 0:
            raise YadError(f"yad_init failed (code {rc}): {lib.yad_last_error().decode()}")
        _inited.add(device_index)
    return lib


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()
