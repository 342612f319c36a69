"""ctypes binding of libofsv.so (include/ofsv.h).  No CPU fallback: if the library cannot be loaded this raises."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libofsv.so")

OK, EINVAL, ECUDA, ENOSUP = 0, -1, -2, -3
REF_CPU, REF_CUDA = 0, 1
F32, BF16 = 0, 1
MAX_TAPS = 64
WL_TAP, WL_STACK = 0, 1
STATE_DHW8, STATE_DWH8 = 0, 1


class ConvDesc(ctypes.Structure):
    """Mirror of `ofsv_conv_desc` (include/ofsv.h)."""
    _fields_ = [
        ("nd", ctypes.c_int32),
        ("N", ctypes.c_int32), ("Di", ctypes.c_int32), ("Hi", ctypes.c_int32), ("Wi", ctypes.c_int32), ("Cin_s", ctypes.c_int32),
        ("Do", ctypes.c_int32), ("Ho", ctypes.c_int32), ("Wo", ctypes.c_int32),
        ("Dy", ctypes.c_int32), ("Hy", ctypes.c_int32), ("Wy", ctypes.c_int32), ("Cout_s", ctypes.c_int32),
        ("Cout_w", ctypes.c_int32),
        ("in_stride", ctypes.c_int32), ("out_stride", ctypes.c_int32),
        ("nphase", ctypes.c_int32), ("ntaps", ctypes.c_int32),
        ("tap_off", (ctypes.c_int8 * 4) * MAX_TAPS),
        ("has_prelu", ctypes.c_int32), ("has_residual", ctypes.c_int32),
        ("in_dtype", ctypes.c_int32), ("out_dtype", ctypes.c_int32),
        ("out_shuffle", ctypes.c_int32),
        ("out_s2d", ctypes.c_int32),
        ("out_shuffle_hfast", ctypes.c_int32),
    ]


class RefreshRec(ctypes.Structure):
    """Mirror of `ofsv_refresh_rec` (include/ofsv.h)."""
    _fields_ = [("src", ctypes.c_void_p), ("dst", ctypes.c_void_p),
                ("kind", ctypes.c_int32), ("A", ctypes.c_int32), ("B", ctypes.c_int32), ("K", ctypes.c_int32), ("swap", ctypes.c_int32),
                ("T", ctypes.c_int32), ("Cin_s", ctypes.c_int32), ("Cout_w", ctypes.c_int32), ("ci0", ctypes.c_int32), ("co0", ctypes.c_int32),
                ("n", ctypes.c_int32), ("pad_", ctypes.c_int32), ("kidx", ctypes.c_int16 * MAX_TAPS)]


class PackRec(ctypes.Structure):
    """Mirror of `ofsv_pack_rec` (include/ofsv.h)."""
    _fields_ = [("w_tap", ctypes.c_void_p), ("w_out", ctypes.c_void_p),
                ("nblocks", ctypes.c_int32), ("Cin_s", ctypes.c_int32), ("Cout_w", ctypes.c_int32), ("KC", ctypes.c_int32),
                ("blk", ctypes.c_uint16 * (MAX_TAPS * 8))]


_P, _I, _L, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
_SIGS = {
    "ofsv_version": (ctypes.c_char_p, []),
    "ofsv_last_error": (ctypes.c_char_p, []),
    "ofsv_launch_count": (_L, []),
    "ofsv_warp2d_f32": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "ofsv_warp3d_f32": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "ofsv_warp3d_gather_f32": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "ofsv_warp2d_bwd_f32": (_I, [_P] * 7 + [_I, _I, _I, _I, _I, _P]),
    "ofsv_warp3d_bwd_f32": (_I, [_P] * 8 + [_I, _I, _I, _I, _I, _I, _P]),
    "ofsv_warp_blend_2d_f32": (_I, [_P] * 10 + [_I, _I, _I, _I, _P]),
    "ofsv_warp_blend_3d_f32": (_I, [_P] * 11 + [_I, _I, _I, _I, _I, _P]),
    "ofsv_blend_f32": (_I, [_P, _P, _P, _P, _L, _P]),
    "ofsv_corr81_fwd_splits": (_I, [_I, _I, _I, _I]),
    "ofsv_corr81_fwd_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _I, _L, _P, _P]),
    "ofsv_corr81_bwd_f32": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ofsv_upsample_flow_ac_f32": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "ofsv_warping_no_div_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "ofsv_torch_warp_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "ofsv_feature_norm_pair_f32": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "ofsv_upsample_flow_ac_bwd_f32": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "ofsv_warping_no_div_bwd_f32": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "ofsv_adamw_step_f32": (_I, [_P, _P, _I, _I, _F, _F, _F, _F, _F, _I, _F, _P]),
    "ofsv_u8_to_f32": (_I, [_P, _P, _L, _F, _P]),
    "ofsv_f32_to_u8": (_I, [_P, _P, _L, _F, _P]),
    "ofsv_sq_err_f64": (_I, [_P, _P, _P, _P, _I, _L, _F, _P]),
    "ofsv_ssim2d_f64": (_I, [_P, _P, _P, _P, _I, _I, _I, ctypes.c_double, _P]),
    "ofsv_pack_nhwc_bf16": (_I, [_P, _P, _I, _P, _I, _L, _I, _P]),
    "ofsv_unpack_nhwc_f32": (_I, [_P, _P, _I, _L, _I, _I, _P]),
    "ofsv_pack_block_input": (_I, [_P] * 7 + [_I] * 9 + [_P]),
    "ofsv_conv_simt": (_I, [ctypes.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P]),
    "ofsv_conv_tc": (_I, [ctypes.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P]),
    "ofsv_conv_halo": (_I, [ctypes.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P]),
    "ofsv_conv_pack_weights": (_I, [ctypes.POINTER(ConvDesc), _P, _P, _I, _P]),
    "ofsv_conv_pack_record": (_I, [ctypes.POINTER(ConvDesc), _I, _P, _P, ctypes.POINTER(PackRec)]),
    "ofsv_conv_pack_weights_batched": (_I, [_P, _I, _P]),
    "ofsv_conv_halo_weight_layout": (_I, [ctypes.POINTER(ConvDesc)]),
    "ofsv_conv_stack_selfcheck": (_I, [ctypes.POINTER(ConvDesc), _I, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)]),
    "ofsv_conv_halo_describe": (_I, [ctypes.POINTER(ConvDesc), ctypes.c_char_p, _I]),
    "ofsv_set_tuning": (_I, [ctypes.c_char_p, _I]),
    "ofsv_conv_refresh_tapform": (_I, [_P, _I, _P]),
    "ofsv_prelu_bias_bwd_blocks": (_I, []),
    "ofsv_prelu_bias_bwd_bf16": (_I, [_P, _P, _P, _P, _P, _P, _P, _L, _I, _P]),
    "ofsv_conv_wgrad_splits": (_I, [ctypes.POINTER(ConvDesc)]),
    "ofsv_conv_wgrad_bf16": (_I, [ctypes.POINTER(ConvDesc), _P, _P, _I, _P, _P, _P]),
    "ofsv_head_upsample_add_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "ofsv_pack_block_input_bwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "ofsv_head_upsample_add": (_I, [_P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "ofsv_block_stage_3d": (_I, [_P] * 11 + [_I] * 9 + [_P]),
}
EXPORTS = tuple(_SIGS)

_lib = None


def lib() -> ctypes.CDLL:
    """Load libofsv.so (building it with nvcc if it is absent or stale).  Raises if that is impossible."""
    global _lib
    if _lib is None:
        from . import build as _build
        try:
            so = _build.build()
        except Exception as e:  # noqa: BLE001
            if not os.path.exists(_SO):
                raise RuntimeError(
                    "libofsv.so is missing and could not be built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                    "this package has no CPU or PyTorch fallback") from e
            so = _SO
        so = os.environ.get("OFSV_LIB", so)       # A/B testing of an alternative build of the same C ABI (tests/ab_build.sh)
        L = ctypes.CDLL(so)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)          # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        msg = lib().ofsv_last_error().decode("utf-8", "replace")
        if rc == ENOSUP:
            raise NotImplementedError(msg)
        raise RuntimeError(f"libofsv error {rc}: {msg}")
