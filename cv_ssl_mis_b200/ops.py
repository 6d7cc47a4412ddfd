"""Tensor-level wrappers over the C ABI (one function per entry point of include/b200ssl.h).

Every function takes CUDA fp32 tensors that the caller owns, borrows their `data_ptr()` for the duration of
the (asynchronous) call on torch's current stream, and returns nothing: outputs are written in place into
caller-provided tensors so the step allocates nothing and can be captured in a CUDA graph.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ConvDesc, B200Error  # noqa: F401  (re-exported)


def _p(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise B200Error("b200ssl ops need CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise B200Error("b200ssl ops need contiguous tensors")
    return t.data_ptr()


def _pf(t):
    if t is not None and t.dtype != torch.float32:
        raise B200Error(f"expected float32 tensor, got {t.dtype}")
    return _p(t)


def _st():
    return torch.cuda.current_stream().cuda_stream


def conv_desc(n, id_, ih, iw, c0, c1, cout, k, stride=1, pad=None, dims=2) -> ConvDesc:
    """k: kernel size per spatial dim (same in all); dims: 2 or 3."""
    if pad is None:
        pad = k // 2 if stride == 1 else 0
    kd = k if dims == 3 else 1
    pd = pad if dims == 3 else 0
    return ConvDesc(n, id_, ih, iw, c0, c1, cout, kd, k, k, stride, pd, pad, pad)


def desc_out_dims(d: ConvDesc):
    od = (d.id + 2 * d.pd - d.kd) // d.stride + 1
    oh = (d.ih + 2 * d.ph - d.kh) // d.stride + 1
    ow = (d.iw + 2 * d.pw - d.kw) // d.stride + 1
    return od, oh, ow


# ------------------------------------------------------------------ convolutions
def conv_packed_floats(mode, O, I, T) -> int:
    return int(_lib.query("b200_conv_packed_floats", mode, O, I, T))


def conv_pack_weights(w, out, mode, O, I, T):
    _lib.call("b200_conv_pack_weights", _pf(w), _pf(out), mode, O, I, T, _st())


def conv_fwd(d, src0, src1, wp, bias, dst, out_nchw=False, exact=False):
    _lib.call("b200_conv_fwd", C.byref(d), _pf(src0), _pf(src1), _pf(wp), _pf(bias), _pf(dst), int(out_nchw), int(exact), _st())


def conv_dgrad(d, dy, wp_dgrad, dx0, dx1=None, accumulate=False, exact=False):
    _lib.call("b200_conv_dgrad", C.byref(d), _pf(dy), _pf(wp_dgrad), _pf(dx0), _pf(dx1), int(accumulate), int(exact), _st())


def conv_k2s2_dgrad(d, dy, wp_d2s, dx, accumulate=False, exact=False):
    _lib.call("b200_conv_k2s2_dgrad", C.byref(d), _pf(dy), _pf(wp_d2s), _pf(dx), int(accumulate), int(exact), _st())


def conv_wgrad_workspace_bytes(d) -> int:
    return int(_lib.query("b200_conv_wgrad_workspace_bytes", C.byref(d)))


def conv_wgrad(d, src0, src1, dy, ws, dw, db, accumulate=False, exact=False):
    _lib.call("b200_conv_wgrad", C.byref(d), _pf(src0), _pf(src1), _pf(dy), _p(ws), ws.numel() * ws.element_size(),
              _pf(dw), _pf(db), int(accumulate), int(exact), _st())


def deconv_k2s2_fwd(d, x, wp, bias, y, exact=False):
    _lib.call("b200_deconv_k2s2_fwd", C.byref(d), _pf(x), _pf(wp), _pf(bias), _pf(y), int(exact), _st())


def deconv_k2s2_dgrad(d, dy, wp_dgrad, dx, accumulate=False, exact=False):
    _lib.call("b200_deconv_k2s2_dgrad", C.byref(d), _pf(dy), _pf(wp_dgrad), _pf(dx), int(accumulate), int(exact), _st())


def deconv_k2s2_wgrad_workspace_bytes(d) -> int:
    return int(_lib.query("b200_deconv_k2s2_wgrad_workspace_bytes", C.byref(d)))


def deconv_k2s2_wgrad(d, x, dy, ws, dw, accumulate=False, exact=False):
    _lib.call("b200_deconv_k2s2_wgrad", C.byref(d), _pf(x), _pf(dy), _p(ws), ws.numel() * ws.element_size(), _pf(dw),
              int(accumulate), int(exact), _st())


# tile kernels (3x3 / 3x3x3 stride-1 pad-1, TF32 production path)
def conv_tile_supported(d, for_wgrad=False) -> bool:
    return bool(_lib.query("b200_conv_tile_supported", C.byref(d), int(for_wgrad)))


def conv_tile_packed_floats(dgrad, O, I, T) -> int:
    return int(_lib.query("b200_conv_tile_packed_floats", int(dgrad), O, I, T))


def conv_tile_pack_weights(w, out, dgrad, O, I, T):
    _lib.call("b200_conv_tile_pack_weights", _pf(w), _pf(out), int(dgrad), O, I, T, _st())


def conv_tile_fwd(d, src0, src1, wt, bias, dst, out_nchw=False):
    _lib.call("b200_conv_tile_fwd", C.byref(d), _pf(src0), _pf(src1), _pf(wt), _pf(bias), _pf(dst), int(out_nchw), _st())


def conv_tile_dgrad(d, dy, wt_dgrad, dx0, dx1=None, accumulate=False):
    _lib.call("b200_conv_tile_dgrad", C.byref(d), _pf(dy), _pf(wt_dgrad), _pf(dx0), _pf(dx1), int(accumulate), _st())


def conv_tile_wgrad_workspace_bytes(d) -> int:
    return int(_lib.query("b200_conv_tile_wgrad_workspace_bytes", C.byref(d)))


def conv_tile_wgrad(d, src0, src1, dy, ws, dw, db, accumulate=False):
    _lib.call("b200_conv_tile_wgrad", C.byref(d), _pf(src0), _pf(src1), _pf(dy), _p(ws), ws.numel() * ws.element_size(),
              _pf(dw), _pf(db), int(accumulate), _st())


# row-ring tcgen05 forward / data gradient (wide-image 2D 3x3 stride-1 pad-1, weights resident, BN statistics fused)
def conv_row_supported(d, dgrad=False) -> int:
    """0 when the row kernels do not serve this convolution, else 8 + the weight-packing mode (conv_pack_batch kind 3)."""
    return int(_lib.query("b200_conv_row_supported", C.byref(d), int(dgrad)))


def conv_row_packed_floats(d, dgrad=False) -> int:
    return int(_lib.query("b200_conv_row_packed_floats", C.byref(d), int(dgrad)))


def conv_row_pack_weights(d, dgrad, w, out):
    _lib.call("b200_conv_row_pack_weights", C.byref(d), int(dgrad), _pf(w), _pf(out), _st())


def conv_row_stats_blocks(d) -> int:
    return int(_lib.query("b200_conv_row_stats_blocks", C.byref(d)))


def conv_row_fwd(d, src0, src1, wpk, bias, dst, stats_part=None):
    """stats_part: workspace for conv_row_stats_blocks(d) x 2 x cout fp64 partial (sum, sum of squares) of dst, or None."""
    _lib.call("b200_conv_row_fwd", C.byref(d), _pf(src0), _pf(src1), _pf(wpk), _pf(bias), _pf(dst), _p(stats_part), _st())


def conv_row_dgrad(d, dy, wpk_dgrad, dx0, dx1=None, accumulate=False):
    _lib.call("b200_conv_row_dgrad", C.byref(d), _pf(dy), _pf(wpk_dgrad), _pf(dx0), _pf(dx1), int(accumulate), _st())


def bn_finalize(part, nblocks, M, C_, gamma, beta, eps, momentum, running_mean, running_var, state):
    """second half of bn_stats_fwd on [nblocks][2][C] fp64 partials (e.g. from conv_row_fwd)"""
    _lib.call("b200_bn_finalize", _p(part), int(nblocks), int(M), int(C_), _pf(gamma), _pf(beta), float(eps), float(momentum),
              _pf(running_mean), _pf(running_var), _pf(state), _st())


# halo-block tcgen05 forward / data gradient (narrow-image 2D 3x3 stride-1 pad-1: the 64^2 / 32^2 / 16^2 levels)
def maxpool3d_fwd(a, out, N, D, H, W, C):
    _lib.call("b200_maxpool3d_fwd", _pf(a), _pf(out), N, D, H, W, C, _st())


def maxpool3d_bwd(a, dp, da, N, D, H, W, C, accumulate=False):
    _lib.call("b200_maxpool3d_bwd", _pf(a), _pf(dp), _pf(da), N, D, H, W, C, int(accumulate), _st())


def upsample3d2x_fwd(x, y, N, D, H, W, C):
    """trilinear x2, align_corners=False; D, H, W: input extents"""
    _lib.call("b200_upsample3d2x_fwd", _pf(x), _pf(y), N, D, H, W, C, _st())


def upsample3d2x_bwd(dy, dx, N, D, H, W, C, accumulate=False):
    _lib.call("b200_upsample3d2x_bwd", _pf(dy), _pf(dx), N, D, H, W, C, int(accumulate), _st())


def s2d_gather3d(x, xs, N, D, H, W, C):
    """xs[(n,do,ho,wo)][(kd,kh,kw,c)] = x[n][2do+kd][2ho+kh][2wo+kw][c]  (space-to-depth view for the 2x2x2 stride-2 convs)"""
    _lib.call("b200_s2d_gather3d", _pf(x), _pf(xs), N, D, H, W, C, _st())


def d2s_scatter3d(ys, bias, y, N, D, H, W, C, accumulate=False):
    """y[n][2d+kd][2h+kh][2w+kw][c] (+)= ys[(n,d,h,w)][(kd,kh,kw,c)] + bias[c]  (depth-to-space of the transposed 2x2x2 convs)"""
    _lib.call("b200_d2s_scatter3d", _pf(ys), _pf(bias), _pf(y), N, D, H, W, C, int(accumulate), _st())


def conv_blk_supported(d, dgrad=False) -> int:
    """0, or 8 + the weight-pack mode (bit 0 data gradient, bit 1 16-channel planes)"""
    return int(_lib.query("b200_conv_blk_supported", C.byref(d), int(dgrad)))


def conv_blk_stats_blocks(d) -> int:
    return int(_lib.query("b200_conv_blk_stats_blocks", C.byref(d)))


def conv_blk_pack_weights(w, out, mode, O, I, taps=9):
    _lib.call("b200_conv_blk_pack_weights", _pf(w), _pf(out), int(mode), O, I, taps, _st())


def conv_blk_fwd(d, src0, src1, wpk, bias, dst, stats_part=None):
    _lib.call("b200_conv_blk_fwd", C.byref(d), _pf(src0), _pf(src1), _pf(wpk), _pf(bias), _pf(dst), _p(stats_part), _st())


def conv_blk_dgrad(d, dy, wpk_dgrad, dx0, dx1=None, accumulate=False):
    _lib.call("b200_conv_blk_dgrad", C.byref(d), _pf(dy), _pf(wpk_dgrad), _pf(dx0), _pf(dx1), int(accumulate), _st())


# row-ring tcgen05 weight gradient (2D 3x3 stride-1 pad-1)
def conv_row_wgrad_supported(d) -> bool:
    return bool(_lib.query("b200_conv_row_wgrad_supported", C.byref(d)))


def conv_row_wgrad_workspace_bytes(d) -> int:
    return int(_lib.query("b200_conv_row_wgrad_workspace_bytes", C.byref(d)))


def conv_row_wgrad(d, src0, src1, dy, ws, dw, accumulate=False, db_zero=None):
    """db_zero: bias gradient of a convolution feeding a train-mode BatchNorm (identically zero; written as such)."""
    _lib.call("b200_conv_row_wgrad", C.byref(d), _pf(src0), _pf(src1), _pf(dy), _p(ws), ws.numel() * ws.element_size(),
              _pf(dw), _pf(db_zero), int(accumulate), _st())


# tcgen05 kernels (2D 3x3 stride-1 pad-1 forward / data gradient)
def conv_umma_supported(d, for_dgrad=False) -> bool:
    return bool(_lib.query("b200_conv_umma_supported", C.byref(d), int(for_dgrad)))


def conv_umma_packed_floats(dgrad, O, I, T) -> int:
    return int(_lib.query("b200_conv_umma_packed_floats", int(dgrad), O, I, T))


def conv_umma_pack_weights(w, out, dgrad, O, I, T):
    _lib.call("b200_conv_umma_pack_weights", _pf(w), _pf(out), int(dgrad), O, I, T, _st())


def conv_umma_fwd(d, src0, src1, wt, bias, dst, out_nchw=False):
    _lib.call("b200_conv_umma2_fwd", C.byref(d), _pf(src0), _pf(src1), _pf(wt), _pf(bias), _pf(dst), int(out_nchw), _st())


def conv_umma_dgrad(d, dy, wt_dgrad, dx0, dx1=None, accumulate=False):
    _lib.call("b200_conv_umma2_dgrad", C.byref(d), _pf(dy), _pf(wt_dgrad), _pf(dx0), _pf(dx1), int(accumulate), _st())


# batched packing / first-layer kernels
def conv_pack_batch(jobs_dev, njobs, blocks_per_job=16, jobs_py=None):
    """jobs_dev: int64 [njobs, 8] device table (see include/b200ssl.h); jobs_py is only used by the CPU test stand-in."""
    _lib.call("b200_conv_pack_batch", _p(jobs_dev), njobs, blocks_per_job, _st())


def conv_c1_supported(d) -> bool:
    return bool(_lib.query("b200_conv_c1_supported", C.byref(d)))


def conv_c1_fwd(d, x, w, bias, y):
    _lib.call("b200_conv_c1_fwd", C.byref(d), _pf(x), _pf(w), _pf(bias), _pf(y), _st())


def conv_c1_wgrad_workspace_bytes(d) -> int:
    return int(_lib.query("b200_conv_c1_wgrad_workspace_bytes", C.byref(d)))


def conv_c1_wgrad(d, x, dy, ws, dw, db, accumulate=False):
    _lib.call("b200_conv_c1_wgrad", C.byref(d), _pf(x), _pf(dy), _p(ws), ws.numel() * ws.element_size(), _pf(dw), _pf(db),
              int(accumulate), _st())


# ------------------------------------------------------------------ norm / activation / dropout
def bn_workspace_bytes(M, C_) -> int:
    return int(_lib.query("b200_bn_workspace_bytes", M, C_))


def bn_stats_fwd(y, M, C_, gamma, beta, eps, momentum, running_mean, running_var, state, ws):
    _lib.call("b200_bn_stats_fwd", _pf(y), M, C_, _pf(gamma), _pf(beta), eps, momentum, _pf(running_mean),
              _pf(running_var), _pf(state), _p(ws), ws.numel() * ws.element_size(), _st())


def bn_eval_state(C_, gamma, beta, eps, running_mean, running_var, state):
    _lib.call("b200_bn_eval_state", C_, _pf(gamma), _pf(beta), eps, _pf(running_mean), _pf(running_var), _pf(state), _st())


def bn_act_fwd(y, state, a, M, C_, slope, p_drop=0.0, drop_mode=0, seed=0, seed_off=None, rng_stream=0, spatial=1):
    _lib.call("b200_bn_act_fwd", _pf(y), _pf(state), _pf(a), M, C_, slope, p_drop, drop_mode, seed, _p(seed_off),
              rng_stream, spatial, _st())


def bn_act_bwd(y, da, state, dy, dgamma, dbeta, M, C_, slope, ws, p_drop=0.0, drop_mode=0, seed=0, seed_off=None,
               rng_stream=0, spatial=1, accumulate=False):
    _lib.call("b200_bn_act_bwd", _pf(y), _pf(da), _pf(state), _pf(dy), _pf(dgamma), _pf(dbeta), int(accumulate), M, C_,
              slope, p_drop, drop_mode, seed, _p(seed_off), rng_stream, spatial, _p(ws),
              ws.numel() * ws.element_size(), _st())


def dropout_mask(mask, M, C_, p_drop, drop_mode, seed, seed_off=None, rng_stream=0, spatial=1):
    _lib.call("b200_dropout_mask", _pf(mask), M, C_, p_drop, drop_mode, seed, _p(seed_off), rng_stream, spatial, _st())


# ------------------------------------------------------------------ resampling / layout
def maxpool2_fwd(a, out, N, H, W, C_):
    _lib.call("b200_maxpool2_fwd", _pf(a), _pf(out), N, H, W, C_, _st())


def maxpool2_bwd(a, dp, da, N, H, W, C_, accumulate=False):
    _lib.call("b200_maxpool2_bwd", _pf(a), _pf(dp), _pf(da), N, H, W, C_, int(accumulate), _st())


def upsample2x_fwd(x, y, N, H, W, C_):
    _lib.call("b200_upsample2x_fwd", _pf(x), _pf(y), N, H, W, C_, _st())


def upsample2x_bwd(dy, dx, N, H, W, C_, accumulate=False):
    _lib.call("b200_upsample2x_bwd", _pf(dy), _pf(dx), N, H, W, C_, int(accumulate), _st())


def nchw_to_nhwc(src, dst, N, C_, S):
    _lib.call("b200_nchw_to_nhwc", _pf(src), _pf(dst), N, C_, S, _st())


def nhwc_to_nchw(src, dst, N, C_, S):
    _lib.call("b200_nhwc_to_nchw", _pf(src), _pf(dst), N, C_, S, _st())


def colsum_workspace_bytes(M, C_) -> int:
    return int(_lib.query("b200_colsum_workspace_bytes", M, C_))


def colsum(g, M, C_, out, ws, accumulate=False):
    _lib.call("b200_colsum", _pf(g), M, C_, _pf(out), int(accumulate), _p(ws), ws.numel() * ws.element_size(), _st())


def add(a, b, c):
    _lib.call("b200_add", _pf(a), _pf(b), _pf(c), a.numel(), _st())


# ------------------------------------------------------------------ nn.Linear on tcgen05 (gemm_umma.cu)
def linear_supported(M, O, c0, c1) -> bool:
    return bool(_lib.query("b200_linear_supported", M, O, c0, c1))


def linear_fwd(x0, x1, w, bias, y, M, O):
    c0, c1 = x0.shape[-1], (x1.shape[-1] if x1 is not None else 0)
    _lib.call("b200_linear_fwd", _pf(x0), _pf(x1), c0, c1, _pf(w), _pf(bias), _pf(y), M, O, _st())


def linear_dgrad(dy, w, dx0, dx1, accumulate, M, O):
    c0, c1 = dx0.shape[-1], (dx1.shape[-1] if dx1 is not None else 0)
    _lib.call("b200_linear_dgrad", _pf(dy), _pf(w), _pf(dx0), _pf(dx1), c0, c1, int(accumulate), M, O, _st())


def linear_wgrad_workspace_bytes(M, O, I) -> int:
    return int(_lib.query("b200_linear_wgrad_workspace_bytes", M, O, I))


def linear_wgrad(x0, x1, dy, dw, ws, M, O, accumulate=False):
    c0, c1 = x0.shape[-1], (x1.shape[-1] if x1 is not None else 0)
    _lib.call("b200_linear_wgrad", _pf(x0), _pf(x1), c0, c1, _pf(dy), _pf(dw), int(accumulate), _p(ws),
              ws.numel() * ws.element_size() if ws is not None else 0, M, O, _st())


# ------------------------------------------------------------------ UNETR ops (vit.cu)
def patch3d_gather(x, y, B, C_, D, H, W, patch):
    _lib.call("b200_patch3d_gather", _pf(x), _pf(y), B, C_, D, H, W, patch, _st())


def mha_probs_floats(B, N, heads) -> int:
    return int(_lib.query("b200_mha_probs_floats", B, N, heads))


def mha_fwd(qkv, out, probs, B, N, heads, hd):
    _lib.call("b200_mha_fwd", _pf(qkv), _pf(out), _pf(probs), B, N, heads, hd, _st())


def mha_bwd(qkv, probs, dout, dqkv, ws, B, N, heads, hd):
    _lib.call("b200_mha_bwd", _pf(qkv), _pf(probs), _pf(dout), _pf(dqkv), _p(ws), ws.numel() * ws.element_size(), B, N, heads, hd, _st())


def add_lrelu_fwd(a, b, out, slope):
    _lib.call("b200_add_lrelu_fwd", _pf(a), _pf(b), _pf(out), a.numel(), slope, _st())


def lrelu_bwd(out, dout, dx, slope):
    _lib.call("b200_lrelu_bwd", _pf(out), _pf(dout), _pf(dx), out.numel(), slope, _st())


# ------------------------------------------------------------------ Swin-UNet token ops
def layernorm_workspace_bytes(M, C_) -> int:
    return int(_lib.query("b200_layernorm_workspace_bytes", M, C_))


def layernorm_fwd(x, gamma, beta, y, stats, M, C_, eps=1e-5):
    _lib.call("b200_layernorm_fwd", _pf(x), _pf(gamma), _pf(beta), _pf(y), _pf(stats), M, C_, eps, _st())


def layernorm_bwd(x, stats, gamma, dy, dx, dgamma, dbeta, M, C_, ws, accumulate_dx=False):
    _lib.call("b200_layernorm_bwd", _pf(x), _pf(stats), _pf(gamma), _pf(dy), _pf(dx), _pf(dgamma), _pf(dbeta),
              int(accumulate_dx), M, C_, _p(ws), ws.numel() * ws.element_size(), _st())


def gelu_fwd(x, y):
    _lib.call("b200_gelu_fwd", _pf(x), _pf(y), x.numel(), _st())


def gelu_bwd(x, dy, dx, accumulate=False):
    _lib.call("b200_gelu_bwd", _pf(x), _pf(dy), _pf(dx), x.numel(), int(accumulate), _st())


def window_attn_fwd(qkv, table, out, B, H, W, C_, heads, ws, shift):
    _lib.call("b200_window_attn_fwd", _pf(qkv), _pf(table), _pf(out), B, H, W, C_, heads, ws, shift, _st())


def window_attn_workspace_bytes(B, H, W, heads, ws) -> int:
    return int(_lib.query("b200_window_attn_workspace_bytes", B, H, W, heads, ws))


def window_attn_bwd(qkv, table, dout, dqkv, dtable, B, H, W, C_, heads, ws, shift, wsp):
    _lib.call("b200_window_attn_bwd", _pf(qkv), _pf(table), _pf(dout), _pf(dqkv), _pf(dtable), B, H, W, C_, heads, ws, shift,
              _p(wsp), wsp.numel() * wsp.element_size(), _st())


def add_droppath(x, branch, out, B, per_sample, p_drop=0.0, seed=0, seed_off=None, rng_stream=0):
    _lib.call("b200_add_droppath", _pf(x), _pf(branch), _pf(out), B, per_sample, p_drop, seed, _p(seed_off), rng_stream, _st())


def patch_merge_gather(x, y, B, H, W, C_, inverse=False, accumulate=False):
    _lib.call("b200_patch_merge_gather", _pf(x), _pf(y), B, H, W, C_, int(inverse), int(accumulate), _st())


def pixel_shuffle(x, y, B, H, W, C_, p, inverse=False):
    _lib.call("b200_pixel_shuffle", _pf(x), _pf(y), B, H, W, C_, p, int(inverse), _st())


def patch_embed_gather(x, y, B, H, W, patch, repeat_channels):
    _lib.call("b200_patch_embed_gather", _pf(x), _pf(y), B, H, W, patch, repeat_channels, _st())


# ------------------------------------------------------------------ loss / optimizer / noise
def _label_dtype(labels):
    if labels is None:
        return _lib.LABEL_U8
    if labels.dtype == torch.uint8:
        return _lib.LABEL_U8
    if labels.dtype == torch.int64:
        return _lib.LABEL_I64
    raise B200Error(f"labels must be uint8 or int64, got {labels.dtype}")


def ssl_loss_workspace_bytes(B, S) -> int:
    return int(_lib.query("b200_ssl_loss_workspace_bytes", B, S))


def ssl_loss_fwd(logits, teacher, labels, nhwc, B, Lb, C_, S, w_cons, lossbuf, ws, mc_psum=None, mc_T=0.0, mc_thr=None):
    _lib.call("b200_ssl_loss_fwd", _pf(logits), _pf(teacher), _p(labels), _label_dtype(labels), int(nhwc), B, Lb, C_, S,
              _pf(w_cons), _pf(mc_psum), float(mc_T), _pf(mc_thr), _pf(lossbuf), _p(ws), ws.numel() * ws.element_size(), _st())


def ssl_loss_bwd(logits, teacher, labels, nhwc, B, Lb, C_, S, w_cons, lossbuf, grad_scale, dlogits, dlogits_nhwc,
                 mc_psum=None, mc_T=0.0, mc_thr=None):
    _lib.call("b200_ssl_loss_bwd", _pf(logits), _pf(teacher), _p(labels), _label_dtype(labels), int(nhwc), B, Lb, C_, S,
              _pf(w_cons), _pf(mc_psum), float(mc_T), _pf(mc_thr), _pf(lossbuf), grad_scale, _pf(dlogits),
              int(dlogits_nhwc), _st())


def ct_loss_fwd(logits, nhwc, other, other_nhwc, labels, B, Lb, C_, S, w_cons, lossbuf, ws):
    _lib.call("b200_ct_loss_fwd", _pf(logits), int(nhwc), _pf(other), int(other_nhwc), _p(labels), _label_dtype(labels), B, Lb,
              C_, S, _pf(w_cons), _pf(lossbuf), _p(ws), ws.numel() * ws.element_size(), _st())


def ct_loss_bwd(logits, nhwc, other, other_nhwc, labels, B, Lb, C_, S, lossbuf, grad_scale, dlogits, dlogits_nhwc):
    _lib.call("b200_ct_loss_bwd", _pf(logits), int(nhwc), _pf(other), int(other_nhwc), _p(labels), _label_dtype(labels), B, Lb,
              C_, S, _pf(lossbuf), grad_scale, _pf(dlogits), int(dlogits_nhwc), _st())


def cps_loss_fwd(logits, nhwc, other, other_nhwc, labels, B, Lb, C_, S, w_cons, lossbuf, ws):
    _lib.call("b200_cps_loss_fwd", _pf(logits), int(nhwc), _pf(other), int(other_nhwc), _p(labels), _label_dtype(labels), B, Lb,
              C_, S, _pf(w_cons), _pf(lossbuf), _p(ws), ws.numel() * ws.element_size(), _st())


def cps_loss_bwd(logits, nhwc, other, other_nhwc, labels, B, Lb, C_, S, lossbuf, grad_scale, dlogits, dlogits_nhwc):
    _lib.call("b200_cps_loss_bwd", _pf(logits), int(nhwc), _pf(other), int(other_nhwc), _p(labels), _label_dtype(labels), B, Lb,
              C_, S, _pf(lossbuf), grad_scale, _pf(dlogits), int(dlogits_nhwc), _st())


def mc_softmax_accumulate(logits, psum, R, U, C_, S, nhwc=False, init=True):
    _lib.call("b200_mc_softmax_accumulate", _pf(logits), _pf(psum), R, U, C_, S, int(nhwc), int(init), _st())


def sgd_ema_step(params, grads, momentum_buf, ema_params, hparams, zero_grad=False):
    _lib.call("b200_sgd_ema_step", _pf(params), _pf(grads), _pf(momentum_buf), _pf(ema_params), params.numel(),
              _pf(hparams), int(zero_grad), _st())


# ------------------------------------------------------------------ stand-alone losses (utils/losses.py drop-ins)
def loss_dropin_workspace_bytes(B, S) -> int:
    return int(_lib.query("b200_loss_dropin_workspace_bytes", int(B), int(S)))


def dice_fwd(x, use_softmax, labels, B, C_, S, weight, out, ws):
    _lib.call("b200_dice_fwd", _pf(x), int(use_softmax), _p(labels), _label_dtype(labels), B, C_, S, _pf(weight), _pf(out), _p(ws),
              ws.numel() * ws.element_size(), _st())


def dice_bwd(x, use_softmax, labels, B, C_, S, weight, fwd_out, grad_out, dx):
    _lib.call("b200_dice_bwd", _pf(x), int(use_softmax), _p(labels), _label_dtype(labels), B, C_, S, _pf(weight), _pf(fwd_out),
              _pf(grad_out), _pf(dx), _st())


def softmax_mse_fwd(a, b, B, C_, S, out):
    _lib.call("b200_softmax_mse_fwd", _pf(a), _pf(b), B, C_, S, _pf(out), _st())


def softmax_mse_bwd(a, b, grad_out, B, C_, S, da):
    _lib.call("b200_softmax_mse_bwd", _pf(a), _pf(b), _pf(grad_out), B, C_, S, _pf(da), _st())


def softmax_kl_fwd(a, b, B, C_, S, out, ws):
    _lib.call("b200_softmax_kl_fwd", _pf(a), _pf(b), B, C_, S, _pf(out), _p(ws), ws.numel() * ws.element_size(), _st())


def softmax_kl_bwd(a, b, grad_out, B, C_, S, da):
    _lib.call("b200_softmax_kl_bwd", _pf(a), _pf(b), _pf(grad_out), B, C_, S, _pf(da), _st())


def ema_update(ema_params, params, hparams):
    _lib.call("b200_ema_update", _pf(ema_params), _pf(params), params.numel(), _pf(hparams), _st())


def noise_add(x, out, sigma, clip, seed, seed_off=None, rng_stream=0):
    _lib.call("b200_noise_add", _pf(x), _pf(out), out.numel(), sigma, clip, seed, _p(seed_off), rng_stream, _st())
