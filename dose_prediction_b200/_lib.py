"""ctypes binding of libdose_b200.so (the C ABI declared in include/dose_b200.h).

There is no CPU or eager-PyTorch fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.
"""
import ctypes
import os
from ctypes import c_double, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdose_b200.so")
_LIB = None

P = c_void_p
I = c_int
L = c_longlong
F = c_float

_SIGNATURES = {
    "dp_conv3d_tc": [P, I, P, I, P, I, I, I, I, I, I, I, P, P, I, P, P, P, I, I, P, P, I, P, I, P],
    "dp_conv3d_stack": [P, I, P, I, P, I, I, I, I, I, I, P, P, I, P, P, P, I, I, P, P, I, I, I, P],
    "dp_conv3d_direct": [P, P, I, I, I, I, I, I, I, I, I, I, P, P, P, I, I, P, P, P, I, I, P, P],
    "dp_gemm_tc": [P, P, I, I, I, I, I, I, L, I, L, I, I, P, P, I, P, F, I, P, I, P, I, I, I, I, I, P, P, P, F, P, P],
    "dp_attention": [P, P, P, I, I, I, I, I, P, I, P, P],
    "dp_pack_ncdhw": [P, I, I, L, P, P, I, I, P],
    "dp_unpack_c8": [P, P, I, I, I, I, L, P, P],
    "dp_norm_act": [P, P, P, I, I, P, P, P, I, P, P, P, P, I, I, I, P, P, I, I, P, I, I, L, P, P, I, I, I, I, I, P],
    "dp_conv3d_c1": [P, P, P, I, I, I, I, P, I, P, P, P],
    "dp_norm_act_resx": [P, I, P, P, P, P, I, P, P, I, I, I, I, L, P],
    "dp_norm_act_head": [P, I, P, P, P, I, P, P, I, I, P, P, I, P, I, I, L, P],
    "dp_pointwise_conv": [I, P, P, P, P, P, P, P, P, P, P, I, I, L, P, P, P, I, I, P, P, I, P],
    "dp_pointwise_tc": [I, P, P, P, P, P, P, P, P, P, P, P, P, I, I, L, P, P, P, I, I, P, P, P],
    "dp_pointwise_conv_cw": [I, P, P, P, P, P, P, P, P, P, P, I, I, L, P, P, P, I, I, P, P, I, P],
    "dp_deconv2x": [P, P, L, L, L, I, I, I, I, I, I, P, P, P, I, I, P],
    "dp_deconv2x_tc": [P, I, P, I, P, I, I, I, I, I, I, I, P, P, P, P, I, I, P, P],
    "dp_deconv2x_cw": [P, P, L, L, I, I, I, I, I, I, P, P, P, I, I, P],
    "dp_deconv2x_gemm": [P, P, I, I, I, I, I, I, P, P, I, I, P, P],
    "dp_upsample2x": [P, P, I, I, I, I, I, I, I, P, P, I, I, P],
    "dp_layernorm": [P, P, P, I, I, P, P, P],
    "dp_softmax": [P, I, I, I, P, I, P],
    "dp_splitk_reduce": [P, I, I, I, P, P, I, P, P],
    "dp_patchify": [P, I, I, I, I, I, I, I, P, P],
    "dp_gemm_patch_embed": [P, I, I, I, I, I, I, I, P, I, I, P, P, I, P, P, P],
    "dp_patchify_planar": [P, I, I, I, I, P, P],
    "dp_crop_pack": [P, I, I, I, I, I, I, P, P, P, P, P, P, I, I, P],
    "dp_window_add": [P, I, I, I, P, P, P, P, P, P, I, I, I, P],
    "dp_div_count": [P, P, L, I, P],
    "dp_handoff": [P, I, P, P, I, I, P, P, I, I, P, P],
    # ---- training step
    "dp_norm_act_bwd": [P, P, P, I, I, P, P, P, I, P, P, P, P, I, I, I, I, P, P, P, P, I, I, P, I, P, P, I, I, P, P, I, I,
                        I, I, L, P],
    "dp_batch_combine": [P, I, I, I, L, P, P, F, P],
    "dp_affine_grad": [P, I, I, P, P, F, P],
    "dp_grad_finalize": [P, P, L, F, P],
    "dp_conv3d_wgrad": [P, I, P, P, P, I, P, I, I, I, I, I, I, I, I, I, I, P, I, P],
    "dp_conv3d_wgrad_tc": [P, I, P, P, P, I, P, I, I, I, I, I, I, I, I, I, P, I, P, P],
    "dp_small_wgrad": [I, P, P, P, P, I, I, I, P, P, I, I, I, P, I, I, I, I, I, P, L, L, L, P, P],
    "dp_deconv2x_bwd_data": [I, P, P, P, P, I, I, I, P, I, I, I, I, I, P, I, I, P, P],
    "dp_head_bwd": [P, P, P, I, I, I, P, I, L, P, I, I, P, P, P],
    "dp_masked_l1": [P, P, I, I, I, P, I, F, P, P],
    "dp_genloss_finalize": [P, I, F, F, P, F, P, P],
    "dp_lerp2x_bwd": [P, L, I, L, P, P],
    "dp_dice_ce": [P, I, P, I, I, L, P, I, F, P, I, P],
    "dp_dice_ce_finalize": [P, I, I, L, P, P],
    "dp_adamw": [P, P, P, P, L, F, F, F, F, F, I, F, P, P],
    "dp_adamw_dev": [P, P, P, P, L, F, F, F, F, F, F, P, P, P],
    "dp_grad_check": [P, L, P, P],
    "dp_quantize_blockwise": [P, L, P, P, P, P],
    "dp_dequantize_blockwise": [P, P, P, L, P, P],
    "dp_pack_conv_weight": [P, I, I, I, I, P, P, I, I, P, P],
    "dp_cast_f16": [P, L, P, P],
    "dp_layernorm_bwd": [P, P, P, P, I, I, P, P, P, P],
    "dp_softmax_bwd": [P, I, P, I, I, I, P, I, P],
    "dp_act_fwd": [P, L, I, P, P],
    "dp_act_bwd": [P, P, L, I, P, P, P],
    "dp_transpose": [P, I, L, I, I, I, P, L, I, I, F, P],
    "dp_heads": [P, I, P, I, I, I, I, I, I, I, I, F, P],
    "dp_colsum": [P, I, L, P, P],
    "dp_add": [P, P, L, P, P, P],
    # ---- on-device evaluation
    "dp_dose_postprocess": [P, P, L, F, P, P],
    "dp_dose_stats": [P, P, P, L, P, I, P, P, P, P, P],
    "dp_dvh_metrics": [P, P, P, I, P, L, F, P, P, P, P],
    "dp_dice_metric": [P, P, I, I, L, P, P, P, P],
    "dp_hd95": [P, P, I, I, I, I, F, P, P, P],
    # ---- input pipeline
    "dp_prepare_input": [P, P, P, P, P, P, I, I, I, F, F, F, P, P, P],
    "dp_flip_rot90": [P, P, I, I, I, I, I, I, I, I, P],
    "dp_permute_flip": [P, P, I, I, I, I, I, I, I, I, I, I, P],
    "dp_posneg_count": [P, I, P, I, F, L, P, P],
    "dp_posneg_crop": [P, I, P, I, F, I, I, I, I, I, P, P, I, P, P, P, P],
}


def exported_symbols():
    """Every symbol include/dose_b200.h declares (checked by tests/test_abi.py)."""
    return sorted(list(_SIGNATURES) + ["dp_last_error", "dp_abi_version", "dp_device_sm_count", "dp_dvh_workspace_bytes",
                                       "dp_handle_create", "dp_handle_device", "dp_handle_destroy", "dp_hd95_workspace_bytes"])


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built (run __graft_entry__.build()); "
                "dose_prediction_b200 has no CPU fallback")
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = c_int
        handle.dp_last_error.restype = ctypes.c_char_p
        handle.dp_abi_version.restype = c_int
        handle.dp_device_sm_count.restype = c_int
        handle.dp_dvh_workspace_bytes.restype = c_longlong
        handle.dp_hd95_workspace_bytes.restype = c_longlong
        handle.dp_hd95_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
        handle.dp_handle_create.restype = c_void_p
        handle.dp_handle_create.argtypes = [c_int]
        handle.dp_handle_device.restype = c_int
        handle.dp_handle_device.argtypes = [c_void_p]
        handle.dp_handle_destroy.restype = None
        handle.dp_handle_destroy.argtypes = [c_void_p]
        _LIB = handle
    return _LIB


def check(rc, what=""):
    if rc != 0:
        msg = lib().dp_last_error().decode(errors="replace")
        raise RuntimeError(f"libdose_b200 {what} failed (rc={rc}): {msg}")


def ptr_array(ptrs):
    """Host array of device pointers for the multi-source entry points."""
    arr = (c_void_p * len(ptrs))(*[c_void_p(p) if p else None for p in ptrs])
    return arr


def int_array(vals):
    return (c_int * len(vals))(*vals)
