"""ctypes binding of libb200ssl.so (the C ABI declared in include/b200ssl.h).

The library is the product: there is no CPU or PyTorch fallback.  Loading fails loudly when the shared
object is missing, and every compute call raises if it returns a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libb200ssl.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "b200ssl.h")

ABI_VERSION = 1

LABEL_U8, LABEL_I64 = 0, 1
PACK_CONV_FWD, PACK_CONV_DGRAD, PACK_CONV_DGRAD_D2S, PACK_DECONV_FWD, PACK_DECONV_DGRAD = range(5)


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("n", "id", "ih", "iw", "c0", "c1", "cout", "kd", "kh", "kw", "stride", "pd", "ph", "pw")]


class B200Error(RuntimeError):
    pass


_P = C.c_void_p          # device pointer
_S = C.c_void_p          # cudaStream_t
_I = C.c_int
_L = C.c_longlong
_F = C.c_float
_U64 = C.c_ulonglong
_U32 = C.c_uint
_D = C.POINTER(ConvDesc)

# name -> (restype, argtypes).  Functions returning int status are checked by `call`.
SIGNATURES = {
    "b200_last_error": (C.c_char_p, []),
    "b200_abi_version": (_I, []),
    "b200_device_sm": (_I, []),
    "b200_conv_packed_floats": (_L, [_I, _I, _I, _I]),
    "b200_conv_pack_weights": (_I, [_P, _P, _I, _I, _I, _I, _S]),
    "b200_conv_fwd": (_I, [_D, _P, _P, _P, _P, _P, _I, _I, _S]),
    "b200_conv_dgrad": (_I, [_D, _P, _P, _P, _P, _I, _I, _S]),
    "b200_conv_k2s2_dgrad": (_I, [_D, _P, _P, _P, _I, _I, _S]),
    "b200_conv_wgrad_workspace_bytes": (_L, [_D]),
    "b200_conv_wgrad": (_I, [_D, _P, _P, _P, _P, _L, _P, _P, _I, _I, _S]),
    "b200_deconv_k2s2_fwd": (_I, [_D, _P, _P, _P, _P, _I, _S]),
    "b200_deconv_k2s2_dgrad": (_I, [_D, _P, _P, _P, _I, _I, _S]),
    "b200_deconv_k2s2_wgrad_workspace_bytes": (_L, [_D]),
    "b200_deconv_k2s2_wgrad": (_I, [_D, _P, _P, _P, _L, _P, _I, _I, _S]),
    "b200_conv_tile_supported": (_I, [_D, _I]),
    "b200_conv_tile_packed_floats": (_L, [_I, _I, _I, _I]),
    "b200_conv_tile_pack_weights": (_I, [_P, _P, _I, _I, _I, _I, _S]),
    "b200_conv_tile_fwd": (_I, [_D, _P, _P, _P, _P, _P, _I, _S]),
    "b200_conv_tile_dgrad": (_I, [_D, _P, _P, _P, _P, _I, _S]),
    "b200_conv_tile_wgrad_workspace_bytes": (_L, [_D]),
    "b200_conv_tile_wgrad": (_I, [_D, _P, _P, _P, _P, _L, _P, _P, _I, _S]),
    "b200_conv_umma_supported": (_I, [_D, _I]),
    "b200_conv_umma_packed_floats": (_L, [_I, _I, _I, _I]),
    "b200_conv_umma_pack_weights": (_I, [_P, _P, _I, _I, _I, _I, _S]),
    "b200_conv_umma2_fwd": (_I, [_D, _P, _P, _P, _P, _P, _I, _S]),
    "b200_conv_umma2_dgrad": (_I, [_D, _P, _P, _P, _P, _I, _S]),
    "b200_conv_row_wgrad_supported": (_I, [_D]),
    "b200_conv_row_wgrad_workspace_bytes": (_L, [_D]),
    "b200_conv_row_wgrad": (_I, [_D, _P, _P, _P, _P, _L, _P, _P, _I, _S]),
    "b200_conv_row_supported": (_I, [_D, _I]),
    "b200_conv_row_packed_floats": (_L, [_D, _I]),
    "b200_conv_row_pack_weights": (_I, [_D, _I, _P, _P, _S]),
    "b200_conv_row_stats_blocks": (_L, [_D]),
    "b200_conv_row_fwd": (_I, [_D, _P, _P, _P, _P, _P, _P, _S]),
    "b200_conv_row_dgrad": (_I, [_D, _P, _P, _P, _P, _I, _S]),
    "b200_maxpool3d_fwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _S]),
    "b200_maxpool3d_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _S]),
    "b200_upsample3d2x_fwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _S]),
    "b200_upsample3d2x_bwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _S]),
    "b200_s2d_gather3d": (_I, [_P, _P, _I, _I, _I, _I, _I, _S]),
    "b200_d2s_scatter3d": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _S]),
    "b200_conv_blk_supported": (_I, [_D, _I]),
    "b200_conv_blk_stats_blocks": (_L, [_D]),
    "b200_conv_blk_pack_weights": (_I, [_P, _P, _I, _I, _I, _I, _S]),
    "b200_conv_blk_fwd": (_I, [_D, _P, _P, _P, _P, _P, _P, _S]),
    "b200_conv_blk_dgrad": (_I, [_D, _P, _P, _P, _P, _I, _S]),
    "b200_bn_finalize": (_I, [_P, _I, _L, _I, _P, _P, _F, _F, _P, _P, _P, _S]),
    "b200_conv_pack_batch": (_I, [_P, _I, _I, _S]),
    "b200_conv_c1_supported": (_I, [_D]),
    "b200_conv_c1_fwd": (_I, [_D, _P, _P, _P, _P, _S]),
    "b200_conv_c1_wgrad_workspace_bytes": (_L, [_D]),
    "b200_conv_c1_wgrad": (_I, [_D, _P, _P, _P, _L, _P, _P, _I, _S]),
    "b200_bn_workspace_bytes": (_L, [_L, _I]),
    "b200_bn_stats_fwd": (_I, [_P, _L, _I, _P, _P, _F, _F, _P, _P, _P, _P, _L, _S]),
    "b200_bn_eval_state": (_I, [_I, _P, _P, _F, _P, _P, _P, _S]),
    "b200_bn_act_fwd": (_I, [_P, _P, _P, _L, _I, _F, _F, _I, _U64, _P, _U32, _L, _S]),
    "b200_bn_act_bwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _L, _I, _F, _F, _I, _U64, _P, _U32, _L, _P, _L, _S]),
    "b200_dropout_mask": (_I, [_P, _L, _I, _F, _I, _U64, _P, _U32, _L, _S]),
    "b200_maxpool2_fwd": (_I, [_P, _P, _I, _I, _I, _I, _S]),
    "b200_maxpool2_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _S]),
    "b200_upsample2x_fwd": (_I, [_P, _P, _I, _I, _I, _I, _S]),
    "b200_upsample2x_bwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _S]),
    "b200_nchw_to_nhwc": (_I, [_P, _P, _L, _I, _L, _S]),
    "b200_nhwc_to_nchw": (_I, [_P, _P, _L, _I, _L, _S]),
    "b200_colsum_workspace_bytes": (_L, [_L, _I]),
    "b200_colsum": (_I, [_P, _L, _I, _P, _I, _P, _L, _S]),
    "b200_add": (_I, [_P, _P, _P, _L, _S]),
    "b200_ssl_loss_workspace_bytes": (_L, [_I, _L]),
    "b200_ssl_loss_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _L, _P, _P, _F, _P, _P, _P, _L, _S]),
    "b200_ssl_loss_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _L, _P, _P, _F, _P, _P, _F, _P, _I, _S]),
    "b200_mc_softmax_accumulate": (_I, [_P, _P, _I, _I, _I, _L, _I, _I, _S]),
    "b200_ct_loss_fwd": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _L, _P, _P, _P, _L, _S]),
    "b200_ct_loss_bwd": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _L, _P, _F, _P, _I, _S]),
    "b200_cps_loss_fwd": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _L, _P, _P, _P, _L, _S]),
    "b200_cps_loss_bwd": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _L, _P, _F, _P, _I, _S]),
    "b200_loss_dropin_workspace_bytes": (_L, [_I, _L]),
    "b200_dice_fwd": (_I, [_P, _I, _P, _I, _I, _I, _L, _P, _P, _P, _L, _S]),
    "b200_dice_bwd": (_I, [_P, _I, _P, _I, _I, _I, _L, _P, _P, _P, _P, _S]),
    "b200_softmax_mse_fwd": (_I, [_P, _P, _I, _I, _L, _P, _S]),
    "b200_softmax_mse_bwd": (_I, [_P, _P, _P, _I, _I, _L, _P, _S]),
    "b200_softmax_kl_fwd": (_I, [_P, _P, _I, _I, _L, _P, _P, _L, _S]),
    "b200_softmax_kl_bwd": (_I, [_P, _P, _P, _I, _I, _L, _P, _S]),
    "b200_linear_supported": (_I, [_L, _I, _I, _I]),
    "b200_linear_fwd": (_I, [_P, _P, _I, _I, _P, _P, _P, _L, _I, _S]),
    "b200_linear_dgrad": (_I, [_P, _P, _P, _P, _I, _I, _I, _L, _I, _S]),
    "b200_linear_wgrad_workspace_bytes": (_L, [_L, _I, _I]),
    "b200_linear_wgrad": (_I, [_P, _P, _I, _I, _P, _P, _I, _P, _L, _L, _I, _S]),
    "b200_patch3d_gather": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _S]),
    "b200_mha_probs_floats": (_L, [_I, _I, _I]),
    "b200_mha_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _S]),
    "b200_mha_bwd": (_I, [_P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _S]),
    "b200_add_lrelu_fwd": (_I, [_P, _P, _P, _L, _F, _S]),
    "b200_lrelu_bwd": (_I, [_P, _P, _P, _L, _F, _S]),
    "b200_pixel_shuffle": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _S]),
    "b200_layernorm_workspace_bytes": (_L, [_L, _I]),
    "b200_layernorm_fwd": (_I, [_P, _P, _P, _P, _P, _L, _I, _F, _S]),
    "b200_layernorm_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _P, _L, _S]),
    "b200_gelu_fwd": (_I, [_P, _P, _L, _S]),
    "b200_gelu_bwd": (_I, [_P, _P, _P, _L, _I, _S]),
    "b200_window_attn_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _S]),
    "b200_window_attn_workspace_bytes": (_L, [_I, _I, _I, _I, _I]),
    "b200_window_attn_bwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _L, _S]),
    "b200_add_droppath": (_I, [_P, _P, _P, _I, _L, _F, _U64, _P, _U32, _S]),
    "b200_patch_merge_gather": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _S]),
    "b200_patch_embed_gather": (_I, [_P, _P, _I, _I, _I, _I, _I, _S]),
    "b200_sgd_ema_step": (_I, [_P, _P, _P, _P, _L, _P, _I, _S]),
    "b200_ema_update": (_I, [_P, _P, _L, _P, _S]),
    "b200_noise_add": (_I, [_P, _P, _L, _F, _F, _U64, _P, _U32, _S]),
}

# entry points whose int return value is NOT a status code
_NON_STATUS = {"b200_abi_version", "b200_device_sm", "b200_conv_tile_supported", "b200_conv_umma_supported",
               "b200_conv_c1_supported", "b200_linear_supported", "b200_conv_row_wgrad_supported", "b200_conv_row_supported", "b200_conv_blk_supported"}

_lib = None
launch_count = 0         # number of status-returning (kernel-launching) calls made through `call`

# optional per-call device timing (bench.py / profiling only): when `profile` is a list, every call is
# bracketed by CUDA events on the current stream and (entry point, tag, start, end, (algorithmic bytes, flops) or None)
# is appended to it
profile = None
tag = ""


def load() -> C.CDLL:
    """Load the shared library (once) and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C cv_ssl_mis_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.b200_abi_version() != ABI_VERSION:
        raise B200Error(f"ABI mismatch: library {lib.b200_abi_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


def call(name: str, *args):
    """Call a status-returning entry point; raise B200Error with the library's message on failure."""
    global launch_count
    lib = load()
    if profile is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        from . import _costs
        profile.append((name, tag, e0, e1, _costs.cost_of(name, args)))
    else:
        rc = getattr(lib, name)(*args)
    if name in _NON_STATUS or SIGNATURES[name][0] is not _I:
        return rc
    launch_count += 1
    if rc != 0:
        msg = lib.b200_last_error()
        raise B200Error(f"{name} failed ({rc}): {msg.decode() if msg else '?'}")
    return rc


def query(name: str, *args):
    """Call a size/metadata query (returns the raw value)."""
    return getattr(load(), name)(*args)
