"""ctypes binding of include/millieye_b200.h.

The library is the product: there is no Python/torch fallback.  `lib()` raises if the .so is
missing (build it with `python -m millieye_b200.build`) and every wrapper raises `MeError`
with the library's message on a non-zero return code.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_uint,
                    c_ulonglong, c_void_p)

from . import build as _build

ME_ACT_LINEAR, ME_ACT_LEAKY, ME_ACT_SIGMOID = 0, 1, 2


class MeError(RuntimeError):
    pass


class ConvDesc(Structure):
    _fields_ = [(n, c_int) for n in (
        "n", "h", "w", "cin", "in_pitch", "cout", "out_pitch", "ksize", "stride", "act", "out_f32",
        "res_pitch")]


class ChainLayer(Structure):
    _fields_ = [("d", ConvDesc), ("x", c_void_p), ("w_packed", c_void_p), ("bias", c_void_p), ("residual", c_void_p),
                ("y", c_void_p), ("dep_layer", c_int), ("res_layer", c_int)]


class HeadWeights(Structure):
    _fields_ = [(n, c_void_p) for n in (
        "net1_w", "net1_b", "net2_w", "net2_b", "radar_w", "radar_b", "radar2_w", "radar2_b",
        "fc1_w", "fc1_b", "fc2_w", "fc2_b")]


class RadarCfg(Structure):
    _fields_ = [("calib", c_double * 12), ("img_w", c_int), ("img_h", c_int), ("max_depth", c_double),
                ("min_velocity", c_double), ("bin_w", c_int), ("bin_h", c_int), ("edges_w", c_double * 33),
                ("edges_h", c_double * 33), ("out_size", c_int)]


class Stage3LossCfg(Structure):
    _fields_ = [(n, c_float) for n in ("iou_hi", "alpha", "lambda_conf", "thr_img", "thr_radar")]


class Stage2LossCfg(Structure):
    _fields_ = [(n, c_float) for n in ("iou_hi", "alpha", "lambda0", "lambda1", "thr")]


class Stage2Weights(Structure):
    _fields_ = [(n, c_void_p) for n in ("net1_w", "net1_b", "net2_w", "net2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")]


# name -> (restype, argtypes); kept in one table so tests can check it against the header.
SIGNATURES = {
    "me_version": (c_int, []),
    "me_last_error": (c_int, [c_char_p, c_size_t]),
    "me_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "me_debug_status": (c_int, [POINTER(c_ulonglong)]),
    "me_conv_set_trace": (c_int, [c_void_p]),
    "me_conv_k_block": (c_int, [c_int]),
    "me_conv_cin_pad": (c_int, [c_int]),
    "me_pack_conv_weights": (c_int, [c_void_p] * 6 + [c_float, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                                     c_void_p]),
    "me_conv_gemm": (c_int, [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "me_conv_chain_eligible": (c_int, [POINTER(ConvDesc)]),
    "me_conv_chain_blob_bytes": (c_size_t, [POINTER(ChainLayer), c_int]),
    "me_conv_chain_build": (c_int, [POINTER(ChainLayer), c_int, c_void_p, c_size_t]),
    "me_conv_chain_plan": (c_int, [POINTER(ChainLayer), c_int, c_void_p, c_size_t]),
    "me_conv_chain_verify": (c_int, [c_void_p]),
    "me_conv_chain_run": (c_int, [c_void_p, c_void_p, c_void_p]),
    "me_conv_workspace_bytes": (c_size_t, []),
    "me_conv_gemm_ws": (c_int, [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                c_void_p]),
    "me_conv_gemm_yolo": (c_int, [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, POINTER(c_float),
                                  c_float, c_int, c_int, c_void_p, c_void_p]),
    "me_fold_first_weights": (c_int, [c_void_p] * 6 + [c_float, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "me_conv_first": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 7 + [c_void_p]),
    "me_conv_first_tc": (c_int, [c_void_p] * 5 + [c_int] * 7 + [c_void_p]),
    "me_conv_pool_supported": (c_int, [POINTER(ConvDesc)]),
    "me_conv_pool": (c_int, [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "me_conv_first_tc_pool": (c_int, [c_void_p] * 5 + [c_int] * 7 + [c_void_p]),
    "me_maxpool2": (c_int, [c_void_p, c_void_p] + [c_int] * 7 + [c_void_p]),
    "me_upsample2": (c_int, [c_void_p, c_void_p] + [c_int] * 6 + [c_void_p]),
    "me_copy_channels": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_int, c_int, c_void_p]),
    "me_nhwc_to_nchw_f32": (c_int, [c_void_p, c_void_p] + [c_int] * 5 + [c_void_p]),
    "me_u8_to_unit_f32": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p]),
    "me_nchw_f32_to_nhwc": (c_int, [c_void_p, c_void_p] + [c_int] * 5 + [c_void_p]),
    "me_yolo_decode": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, POINTER(c_float), c_float,
                               c_int, c_int, c_void_p]),
    "me_filter_nms_workspace": (c_size_t, [c_int, c_int, c_int]),
    "me_filter_nms": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_double, c_int, c_int, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_size_t, c_void_p]),
    "me_psroi_align": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                               c_int, c_void_p, c_int, c_void_p]),
    "me_roi_align": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                             c_int, c_void_p, c_int, c_void_p]),
    "me_stream_wait_value32": (c_int, [c_void_p, c_uint, c_void_p]),
    "me_stream_write_value32": (c_int, [c_void_p, c_uint, c_void_p]),
    "me_peer_copy": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "me_peer_enable": (c_int, [c_int]),
    "me_roi_gather_bin_major": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                        c_int, c_void_p, c_int, c_int, c_void_p]),
    "me_build_proposals": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_float,
                                   c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "me_build_proposals_dev": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_float,
                                   c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "me_fusion_heads": (c_int, [c_void_p, c_int, c_void_p, c_int, POINTER(HeadWeights), c_void_p, c_void_p, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "me_finalize_workspace": (c_size_t, [c_int]),
    "me_radar_maps": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(RadarCfg), c_void_p, c_void_p, c_void_p, c_void_p]),
    "me_stage2_heads": (c_int, [c_void_p, c_int, POINTER(Stage2Weights), c_void_p, c_int, c_int, c_void_p, c_int, c_void_p,
                                c_void_p, c_void_p, c_void_p]),
    "me_stage2_loss": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                               c_void_p, c_void_p, POINTER(Stage2LossCfg), c_void_p, c_void_p, c_void_p]),
    "me_stage3_labels": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "me_stage3_loss": (c_int, [c_void_p] * 5 + [c_int, c_void_p, c_void_p, c_void_p, POINTER(Stage3LossCfg), c_void_p,
                               c_void_p]),
    "me_yolo_loss_workspace": (c_size_t, [c_int, c_int, c_int]),
    "me_yolo_loss": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, POINTER(c_float), c_float, c_void_p, c_int, c_float,
                             c_float, c_float, c_void_p, c_size_t, c_void_p, c_void_p]),
    "me_gemm_f32": (c_int, [c_int, c_int, c_int, c_void_p, c_longlong, c_longlong, c_void_p, c_longlong, c_longlong, c_void_p,
                            c_longlong, c_void_p, c_int, c_int, c_void_p]),
    "me_colsum_f32": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_longlong, c_longlong, c_void_p, c_void_p]),
    "me_half_rows_to_float": (c_int, [c_void_p, c_longlong, c_int, c_int, c_void_p, c_void_p]),
    "me_nchw_to_rows_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "me_im2col3_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "me_col2im3_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "me_bn_partial_stats": (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_void_p]),
    "me_bn_finalize": (c_int, [c_void_p, c_int, c_float, c_float] + [c_void_p] * 5),
    "me_bn_apply": (c_int, [c_void_p, c_longlong, c_int] + [c_void_p] * 7),
    "me_bn_train_fwd": (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_float, c_float] + [c_void_p] * 8),
    "me_bn_bwd_sums": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_void_p]),
    "me_bn_bwd_apply": (c_int, [c_void_p, c_void_p, c_longlong, c_int] + [c_void_p] * 7),
    "me_leaky_bwd_f32": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p]),
    "me_sigmoid_bwd_f32": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p]),
    "me_roi_align_f32": (c_int, [c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p,
                                 c_int, c_void_p, c_void_p, c_void_p]),
    "me_stage3_tail_fwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int] + [c_void_p] * 9),
    "me_stage3_tail_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                   c_float, c_float] + [c_void_p] * 10),
    "me_adam_step": (c_int, [c_void_p] * 4 + [c_longlong, c_float, c_float, c_float, c_float, c_int, c_void_p]),
    "me_finalize_output": (c_int, [c_void_p] * 6 + [c_int, c_int, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p,
                                                    c_size_t, c_void_p]),
}

_LIB = None


def lib_path():
    return _build.LIB


def lib():
    """Load (once) and return the ctypes handle; never falls back to anything else."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise MeError(
                f"{path} is missing: the CUDA extension is the only implementation of this path. "
                "Build it with `python -m millieye_b200.build` (needs nvcc).")
        handle = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


def last_error():
    buf = ctypes.create_string_buffer(512)
    lib().me_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc, what=""):
    if rc != 0:
        raise MeError(f"{what or 'millieye_b200'} failed (code {rc}): {last_error()}")


def ptr(t):
    """Device pointer of a torch tensor (or None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return c_void_p(s.cuda_stream)


def debug_status():
    w = c_ulonglong(0)
    lib().me_debug_status(byref(w))
    return w.value
