"""CPU stand-in for cv_ssl_mis_b200.ops, used ONLY by the `-m "not gpu"` host-logic tests.

It implements every ops.* entry point with plain torch on CPU tensors (same in-place, channels-last
contract), so the launch schedules (UNetPlan / VNetPlan / trainers) can be exercised and compared with the
oracle in the build container, which has no GPU.  The product never imports this file: cv_ssl_mis_b200.ops
talks to libb200ssl.so only and raises on CPU tensors.
"""
import numpy as np
import torch
import torch.nn.functional as F

from cv_ssl_mis_b200._lib import (ConvDesc, PACK_CONV_FWD, PACK_CONV_DGRAD, PACK_CONV_DGRAD_D2S, PACK_DECONV_FWD,
                                  PACK_DECONV_DGRAD)
from cv_ssl_mis_b200 import ops as real_ops
from oracle import philox

conv_desc = real_ops.conv_desc
desc_out_dims = real_ops.desc_out_dims
B200Error = real_ops.B200Error


def _ld(cols):
    return (cols + 3) // 4 * 4


def _pack_dims(mode, O, I, T):
    return {PACK_CONV_FWD: (T * I, O), PACK_CONV_DGRAD: (T * O, I), PACK_CONV_DGRAD_D2S: (O, T * I),
            PACK_DECONV_FWD: (I, T * O), PACK_DECONV_DGRAD: (T * O, I)}[mode]


def conv_packed_floats(mode, O, I, T):
    r, c = _pack_dims(mode, O, I, T)
    return r * _ld(c)


def conv_pack_weights(w, out, mode, O, I, T):
    if mode == PACK_CONV_FWD:
        m = w.reshape(O, I, T).permute(2, 1, 0).reshape(T * I, O)
    elif mode == PACK_CONV_DGRAD:
        m = w.reshape(O, I, T).flip(2).permute(2, 0, 1).reshape(T * O, I)
    elif mode == PACK_CONV_DGRAD_D2S:
        m = w.reshape(O, I, T).permute(0, 2, 1).reshape(O, T * I)
    elif mode == PACK_DECONV_FWD:
        m = w.reshape(I, O, T).permute(0, 2, 1).reshape(I, T * O)
    else:
        m = w.reshape(I, O, T).permute(2, 1, 0).reshape(T * O, I)
    r, c = m.shape
    o = out.view(r, _ld(c))
    o.zero_()
    o[:, :c] = m.detach()


def _unpack(wp, mode, O, I, T):
    r, c = _pack_dims(mode, O, I, T)
    m = wp.view(r, _ld(c))[:, :c]
    if mode == PACK_CONV_FWD:
        return m.reshape(T, I, O).permute(2, 1, 0)                 # [O, I, T]
    if mode == PACK_CONV_DGRAD:
        return m.reshape(T, O, I).permute(1, 2, 0).flip(2)
    if mode == PACK_CONV_DGRAD_D2S:
        return m.reshape(O, T, I).permute(0, 2, 1)
    if mode == PACK_DECONV_FWD:
        return m.reshape(I, T, O).permute(0, 2, 1)                 # [I, O, T]
    return m.reshape(T, O, I).permute(2, 1, 0)


def _cl_to_ncdhw(t, n, d, h, w, c):
    return t.reshape(n, d, h, w, c).permute(0, 4, 1, 2, 3)


def _ncdhw_to_cl(t):
    return t.permute(0, 2, 3, 4, 1).reshape(-1, t.shape[1])


def _store(dst, val, accumulate):
    val = val.reshape(dst.shape)
    if accumulate:
        dst += val
    else:
        dst.copy_(val)


def _input(d, src0, src1):
    x = _cl_to_ncdhw(src0, d.n, d.id, d.ih, d.iw, d.c0)
    if d.c1:
        x = torch.cat([x, _cl_to_ncdhw(src1, d.n, d.id, d.ih, d.iw, d.c1)], 1)
    return x


def _kw(d):
    return dict(stride=(1 if d.kd == 1 else d.stride, d.stride, d.stride), padding=(d.pd, d.ph, d.pw))


def conv_fwd(d, src0, src1, wp, bias, dst, out_nchw=False, exact=False):
    T = d.kd * d.kh * d.kw
    W = _unpack(wp, PACK_CONV_FWD, d.cout, d.c0 + d.c1, T).reshape(d.cout, d.c0 + d.c1, d.kd, d.kh, d.kw)
    y = F.conv3d(_input(d, src0, src1), W, bias, **_kw(d))
    if out_nchw:
        dst.copy_(y.reshape(dst.shape))
    else:
        dst.copy_(_ncdhw_to_cl(y).reshape(dst.shape))


def _split_store(dx, d, dx0, dx1, accumulate):
    cl = _ncdhw_to_cl(dx)
    _store(dx0, cl[:, :d.c0], accumulate)
    if d.c1:
        _store(dx1, cl[:, d.c0:], accumulate)


def conv_dgrad(d, dy, wp_dgrad, dx0, dx1=None, accumulate=False, exact=False):
    T = d.kd * d.kh * d.kw
    cin = d.c0 + d.c1
    W = _unpack(wp_dgrad, PACK_CONV_DGRAD, d.cout, cin, T).reshape(d.cout, cin, d.kd, d.kh, d.kw)
    g = _cl_to_ncdhw(dy, d.n, d.id, d.ih, d.iw, d.cout)
    dx = F.conv_transpose3d(g, W, None, stride=1, padding=(d.pd, d.ph, d.pw))
    _split_store(dx, d, dx0, dx1, accumulate)


def conv_k2s2_dgrad(d, dy, wp_d2s, dx, accumulate=False, exact=False):
    T = d.kd * d.kh * d.kw
    W = _unpack(wp_d2s, PACK_CONV_DGRAD_D2S, d.cout, d.c0, T).reshape(d.cout, d.c0, d.kd, d.kh, d.kw)
    od, oh, ow = desc_out_dims(d)
    g = _cl_to_ncdhw(dy, d.n, od, oh, ow, d.cout)
    r = F.conv_transpose3d(g, W, None, stride=(1 if d.kd == 1 else 2, 2, 2))
    _store(dx, _ncdhw_to_cl(r), accumulate)


def conv_wgrad_workspace_bytes(d):
    return 64


def conv_wgrad(d, src0, src1, dy, ws, dw, db, accumulate=False, exact=False):
    x = _input(d, src0, src1)
    od, oh, ow = desc_out_dims(d)
    g = _cl_to_ncdhw(dy, d.n, od, oh, ow, d.cout)
    kw = _kw(d)
    gw = torch.nn.grad.conv3d_weight(x, (d.cout, d.c0 + d.c1, d.kd, d.kh, d.kw), g, stride=kw["stride"], padding=kw["padding"])
    _store(dw, gw, accumulate)
    if db is not None:
        _store(db, g.sum((0, 2, 3, 4)), accumulate)


def deconv_k2s2_fwd(d, x, wp, bias, y, exact=False):
    T = d.kd * d.kh * d.kw
    W = _unpack(wp, PACK_DECONV_FWD, d.cout, d.c0, T).reshape(d.c0, d.cout, d.kd, d.kh, d.kw)
    xi = _cl_to_ncdhw(x, d.n, d.id, d.ih, d.iw, d.c0)
    r = F.conv_transpose3d(xi, W, bias, stride=(1 if d.kd == 1 else 2, 2, 2))
    y.copy_(_ncdhw_to_cl(r).reshape(y.shape))


def deconv_k2s2_dgrad(d, dy, wp_dgrad, dx, accumulate=False, exact=False):
    T = d.kd * d.kh * d.kw
    W = _unpack(wp_dgrad, PACK_DECONV_DGRAD, d.cout, d.c0, T).reshape(d.c0, d.cout, d.kd, d.kh, d.kw)
    g = _cl_to_ncdhw(dy, d.n, d.id * d.kd, d.ih * 2, d.iw * 2, d.cout)
    r = F.conv3d(g, W.permute(0, 1, 2, 3, 4), None, stride=(1 if d.kd == 1 else 2, 2, 2))   # W as [out=c0][in=cout]
    _store(dx, _ncdhw_to_cl(r), accumulate)


def deconv_k2s2_wgrad_workspace_bytes(d):
    return 64


def deconv_k2s2_wgrad(d, x, dy, ws, dw, accumulate=False, exact=False):
    xi = _cl_to_ncdhw(x, d.n, d.id, d.ih, d.iw, d.c0).detach()
    g = _cl_to_ncdhw(dy, d.n, d.id * d.kd, d.ih * 2, d.iw * 2, d.cout)
    with torch.enable_grad():
        W = torch.zeros(d.c0, d.cout, d.kd, d.kh, d.kw, requires_grad=True)
        out = F.conv_transpose3d(xi, W, None, stride=(1 if d.kd == 1 else 2, 2, 2))
        (gw,) = torch.autograd.grad(out, W, g)
    _store(dw, gw, accumulate)


# ------------------------------------------------------------------ tile kernels (own weight packing)
def conv_tile_supported(d, for_wgrad=False):
    ok = (d.stride == 1 and d.kh == 3 and d.kw == 3 and d.ph == 1 and d.pw == 1 and
          ((d.kd == 1 and d.pd == 0) or (d.kd == 3 and d.pd == 1)) and d.c0 % 4 == 0 and d.c1 % 4 == 0)
    if for_wgrad:
        ok = ok and d.cout % 4 == 0
    return bool(ok)


def _r16(v):
    return (v + 15) // 16 * 16


def conv_tile_packed_floats(dgrad, O, I, T):
    rows, cols = (O, I) if dgrad else (I, O)
    return ((rows + 15) // 16) * T * _r16(cols) * 16


def conv_tile_pack_weights(w, out, dgrad, O, I, T):
    W = w.detach().reshape(O, I, T)
    if dgrad:      # out[chunk][tap][i][kk] = W[chunk*16+kk][i][T-1-tap]
        m = W.flip(2).permute(2, 1, 0)                 # [tap][i][o]
        rows, cols = O, I
    else:          # out[chunk][tap][o][kk] = W[o][chunk*16+kk][tap]
        m = W.permute(2, 0, 1)                         # [tap][o][i]
        rows, cols = I, O
    chunks, colsP = (rows + 15) // 16, _r16(cols)
    buf = torch.zeros(T, colsP, chunks * 16)
    buf[:, :cols, :rows] = m
    out.copy_(buf.reshape(T, colsP, chunks, 16).permute(2, 0, 1, 3).reshape(-1))


def _tile_unpack(wt, dgrad, O, I, T):
    rows, cols = (O, I) if dgrad else (I, O)
    chunks, colsP = (rows + 15) // 16, _r16(cols)
    buf = wt.reshape(chunks, T, colsP, 16).permute(1, 2, 0, 3).reshape(T, colsP, chunks * 16)[:, :cols, :rows]
    if dgrad:
        return buf.permute(2, 1, 0).flip(2)            # [o][i][tap]
    return buf.permute(1, 2, 0)                        # [o][i][tap]


def conv_tile_fwd(d, src0, src1, wt, bias, dst, out_nchw=False):
    T = d.kd * d.kh * d.kw
    W = _tile_unpack(wt, False, d.cout, d.c0 + d.c1, T).reshape(d.cout, d.c0 + d.c1, d.kd, d.kh, d.kw)
    y = F.conv3d(_input(d, src0, src1), W, bias, **_kw(d))
    dst.copy_((y if out_nchw else _ncdhw_to_cl(y)).reshape(dst.shape))


def conv_tile_dgrad(d, dy, wt_dgrad, dx0, dx1=None, accumulate=False):
    T = d.kd * d.kh * d.kw
    cin = d.c0 + d.c1
    W = _tile_unpack(wt_dgrad, True, d.cout, cin, T).reshape(d.cout, cin, d.kd, d.kh, d.kw)
    g = _cl_to_ncdhw(dy, d.n, d.id, d.ih, d.iw, d.cout)
    dx = F.conv_transpose3d(g, W, None, stride=1, padding=(d.pd, d.ph, d.pw))
    _split_store(dx, d, dx0, dx1, accumulate)


def conv_tile_wgrad_workspace_bytes(d):
    return 64


def conv_tile_wgrad(d, src0, src1, dy, ws, dw, db, accumulate=False):
    conv_wgrad(d, src0, src1, dy, ws, dw, db, accumulate)


# ------------------------------------------------------------------ row-ring tcgen05 forward / data gradient
def _row_geo(d, dgrad):
    """(mode, planes, columns, plane width) of csrc/conv_row.cu:rgeometry, or None."""
    if not (d.kd == 1 and d.id == 1 and d.kh == 3 and d.kw == 3 and d.stride == 1 and d.ph == 1 and d.pw == 1 and d.pd == 0):
        return None
    if d.iw < 128:
        return None
    a0, a1 = (d.cout, 0) if dgrad else (d.c0, d.c1)
    n0, n1 = (d.c0, d.c1) if dgrad else (d.cout, 0)
    if n1 != 0 and n1 != n0:
        return None
    pair = a0 == 16 and a1 in (0, 16) and n0 == 16 and d.iw % 256 == 0
    if pair:
        cpp, npl, wk, ncols, G = 32, (2 if a1 else 1), d.iw // 2, 2 * (n0 + n1), 32
    else:
        if a0 > 0 and a0 % 32 == 0 and a1 % 32 == 0:
            cpp, npl = 32, (a0 + a1) // 32
        elif a0 == 16 and a1 in (0, 16):
            cpp, npl = 16, (2 if a1 else 1)
        else:
            return None
        wk, ncols = d.iw, n0 + n1
        if ncols not in (16, 32, 64):
            return None
        G = 32 if n0 >= 32 else 16
        if n0 % G:
            return None
    if wk % 128 or wk > 384:
        return None
    P = (wk + 2 + 7) // 8 * 8
    pa0, pa1 = (32, 32 if a1 else 0) if pair else (a0, a1)
    if P > 256 and (pa0 > cpp or pa1 > cpp):
        return None
    slot, wb, stage = npl * P * cpp * 4, 9 * npl * ncols * cpp * 4, 2 * (ncols // G) * 128 * G * 4
    if pair:
        wb = wb * 2 // 3
    if 1024 + (wb + 1023) // 1024 * 1024 + 4 * slot + 1024 + stage > 224 * 1024:
        return None
    return (1 if dgrad else 0) | (2 if cpp == 16 else 0) | (4 if pair else 0)


def conv_row_supported(d, dgrad=False):
    m = _row_geo(d, dgrad)
    return 0 if m is None else 8 + m


def conv_row_packed_floats(d, dgrad=False):
    m = _row_geo(d, dgrad)
    return d.cout * (d.c0 + d.c1) * (24 if m & 4 else 9)


def _row_pack_index(mode, O, I, T=9):
    """csrc/conv_row_pack.cuh:row_pack_elem as index arithmetic: (flat source index into w[O][I][T], validity) per output element."""
    dgrad = mode & 1
    rows, cols = (O, I) if dgrad else (I, O)
    if not mode & 4:
        cpp = 16 if mode & 2 else 32
        npl = rows // cpp
        idx = torch.arange(T * O * I)
        k, r = idx % cpp, idx // cpp
        col, r = r % cols, r // cols
        pl, tap = r % npl, r // npl
        row = pl * cpp + k
        src = (row * I + col) * T + (T - 1 - tap) if dgrad else (col * I + row) * T + tap
        return src, torch.ones_like(src, dtype=torch.bool)
    npl = rows // 16
    idx = torch.arange(6 * O * I * 4)
    k, r = idx & 31, idx >> 5
    col, r = r % (2 * cols), r // (2 * cols)
    pl, t = r % npl, r // npl
    kh, j = t >> 1, t & 1
    pb, arow = k >> 4, pl * 16 + (k & 15)
    s_ = torch.where(j == 0, torch.ones_like(j), torch.where(pb == 1, torch.zeros_like(j), 2 * torch.ones_like(j)))
    dst, pa, ncol = col >> 5, (col >> 4) & 1, (col >> 5) * 16 + (col & 15)
    kw = 2 * (s_ - 1) + pb - pa + 1
    ok = (kw >= 0) & (kw <= 2)
    kwc = kw.clamp(0, 2)
    src = (arow * I + ncol) * 9 + (2 - kh) * 3 + (2 - kwc) if dgrad else (ncol * I + arow) * 9 + kh * 3 + kwc
    return src, ok


TF32_ROUND = False      # the real packers round the weights to TF32; the host-logic tests keep them exact


def _tf32_rna(v, force=False):
    if not (TF32_ROUND or force):
        return v.contiguous()
    b = v.contiguous().view(torch.int32)
    return ((b + 0x1000) & ~0x1FFF).view(torch.float32)


def conv_row_pack_weights(d, dgrad, w, out):
    O, I = d.cout, d.c0 + d.c1
    src, ok = _row_pack_index(_row_geo(d, dgrad), O, I)
    out.copy_(torch.where(ok, _tf32_rna(w.detach().reshape(-1))[src], torch.zeros(())))


def _row_unpack(d, dgrad, wpk):
    """Framework-layout weights [O][I][9] back from the packed array (every tap appears at least once)."""
    O, I = d.cout, d.c0 + d.c1
    src, ok = _row_pack_index(_row_geo(d, dgrad), O, I)
    W = torch.zeros(O * I * 9)
    W[src[ok]] = wpk.reshape(-1)[ok]
    return W.reshape(O, I, 1, 3, 3)


def conv_row_stats_blocks(d):
    return min(148, d.n * d.ih)


def conv_row_fwd(d, src0, src1, wpk, bias, dst, stats_part=None):
    W = _row_unpack(d, False, wpk)
    y = _ncdhw_to_cl(F.conv3d(_input(d, src0, src1), W, bias, **_kw(d))).reshape(dst.shape)
    dst.copy_(y)
    if stats_part is not None:
        nb = conv_row_stats_blocks(d)
        part = torch.zeros(nb, 2, d.cout, dtype=torch.float64)
        part[0, 0], part[0, 1] = y.double().sum(0), (y.double() ** 2).sum(0)
        stats_part.view(torch.float64)[:part.numel()].copy_(part.reshape(-1))


def conv_row_dgrad(d, dy, wpk_dgrad, dx0, dx1=None, accumulate=False):
    W = _row_unpack(d, True, wpk_dgrad)
    g = _cl_to_ncdhw(dy, d.n, d.id, d.ih, d.iw, d.cout)
    dx = F.conv_transpose3d(g, W, None, stride=1, padding=(d.pd, d.ph, d.pw))
    _split_store(dx, d, dx0, dx1, accumulate)


# ------------------------------------------------------------------ 3D pooling / trilinear up-sampling
def _vol(t, N, D, H, W, C):
    return t.reshape(N, D, H, W, C).permute(0, 4, 1, 2, 3)


def _unvol(t):
    return t.permute(0, 2, 3, 4, 1).reshape(-1, t.shape[1])


def maxpool3d_fwd(a, out, N, D, H, W, C):
    out.copy_(_unvol(F.max_pool3d(_vol(a, N, D, H, W, C), 2)).reshape(out.shape))


def maxpool3d_bwd(a, dp, da, N, D, H, W, C, accumulate=False):
    with torch.enable_grad():
        x = _vol(a, N, D, H, W, C).detach().clone().requires_grad_(True)
        F.max_pool3d(x, 2).backward(_vol(dp, N, D // 2, H // 2, W // 2, C))
    g = _unvol(x.grad).reshape(da.shape)
    da.copy_(da + g if accumulate else g)


def upsample3d2x_fwd(x, y, N, D, H, W, C):
    y.copy_(_unvol(F.interpolate(_vol(x, N, D, H, W, C), scale_factor=2, mode="trilinear", align_corners=False)).reshape(y.shape))


def upsample3d2x_bwd(dy, dx, N, D, H, W, C, accumulate=False):
    with torch.enable_grad():
        x = torch.zeros(N, C, D, H, W, requires_grad=True)
        F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=False).backward(_vol(dy, N, 2 * D, 2 * H, 2 * W, C))
    g = _unvol(x.grad).reshape(dx.shape)
    dx.copy_(dx + g if accumulate else g)


# ------------------------------------------------------------------ 2x2x2 stride-2 views
def s2d_gather3d(x, xs, N, D, H, W, C):
    v = x.reshape(N, D // 2, 2, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 5, 2, 4, 6, 7)     # n do ho wo kd kh kw c
    xs.copy_(v.reshape(xs.shape))


def d2s_scatter3d(ys, bias, y, N, D, H, W, C, accumulate=False):
    v = ys.reshape(N, D, H, W, 2, 2, 2, C)
    if bias is not None:
        v = v + bias
    v = v.permute(0, 1, 4, 2, 5, 3, 6, 7).reshape(y.shape)                                    # n d kd h kh w kw c
    y.copy_(y + v if accumulate else v)


# ------------------------------------------------------------------ halo-block tcgen05 forward / data gradient
def conv_blk_supported(d, dgrad=False):
    """0 or 8 + weight-pack mode, as csrc/conv_blk.cu:bgeometry (minus the shared-memory fit)"""
    flat = d.kd == 1 and d.id == 1 and d.pd == 0
    is3 = d.kd == 3 and d.pd == 1
    if not ((flat or is3) and d.kh == 3 and d.kw == 3 and d.stride == 1 and d.ph == 1 and d.pw == 1):
        return 0
    if d.iw < 4 or d.iw > 96:
        return 0
    a0, a1 = (d.cout, 0) if dgrad else (d.c0, d.c1)
    n0, n1 = (d.c0, d.c1) if dgrad else (d.cout, 0)
    if a0 <= 0 or a0 % 16 or a1 % 16:
        return 0
    cpp = 32 if (a0 % 32 == 0 and a1 % 32 == 0) else 16
    if cpp == 16 and not is3:
        return 0
    ntot = n0 + n1
    if ntot % 16 or (ntot % 32 and not is3):
        return 0
    nt = 64 if ntot % 64 == 0 and (n1 == 0 or n0 % 64 == 0) else (32 if ntot % 32 == 0 and (n1 == 0 or n0 % 32 == 0) else 16)
    if n1 != 0 and n0 % nt:
        return 0
    return 8 + (1 if dgrad else 0) + (2 if cpp == 16 else 0)


def conv_blk_stats_blocks(d):
    return d.n * d.id * d.ih   # (any upper bound of the real block count works for the stand-in: unused rows stay zero)


def conv_blk_pack_weights(w, out, mode, O, I, taps=9):
    src, ok = _row_pack_index(int(mode), O, I, taps)
    out.copy_(_tf32_rna(w.detach().reshape(-1))[src])


def _blk_unpack(wpk, mode, O, I, kd=1):
    src, _ = _row_pack_index(mode, O, I, 9 * kd)
    W = torch.zeros(O * I * 9 * kd)
    W[src] = wpk.reshape(-1)
    return W.reshape(O, I, kd, 3, 3)


def conv_blk_fwd(d, src0, src1, wpk, bias, dst, stats_part=None):
    W = _blk_unpack(wpk, conv_blk_supported(d, False) - 8, d.cout, d.c0 + d.c1, d.kd)
    y = _ncdhw_to_cl(F.conv3d(_input(d, src0, src1), W, bias, **_kw(d))).reshape(dst.shape)
    dst.copy_(y)
    if stats_part is not None:
        nb = conv_blk_stats_blocks(d)
        part = torch.zeros(nb, 2, d.cout, dtype=torch.float64)
        part[0, 0], part[0, 1] = y.double().sum(0), (y.double() ** 2).sum(0)
        stats_part.view(torch.float64)[:part.numel()].copy_(part.reshape(-1))


def conv_blk_dgrad(d, dy, wpk_dgrad, dx0, dx1=None, accumulate=False):
    W = _blk_unpack(wpk_dgrad, conv_blk_supported(d, True) - 8, d.cout, d.c0 + d.c1, d.kd)
    g = _cl_to_ncdhw(dy, d.n, d.id, d.ih, d.iw, d.cout)
    dx = F.conv_transpose3d(g, W, None, stride=1, padding=(d.pd, d.ph, d.pw))
    _split_store(dx, d, dx0, dx1, accumulate)


# ------------------------------------------------------------------ row-ring tcgen05 weight gradient
def conv_row_wgrad_supported(d):
    flat = d.kd == 1 and d.id == 1 and d.pd == 0
    if not ((flat or (d.kd == 3 and d.pd == 1)) and d.kh == 3 and d.kw == 3 and d.stride == 1 and d.ph == 1 and d.pw == 1):
        return False
    c0, c1, co = d.c0, d.c1, d.cout
    if c0 > 0 and c0 % 32 == 0 and c1 % 32 == 0 and co % 32 == 0 and (co <= 128 or co % 128 == 0):
        return d.iw + 2 <= 256
    return c0 == 16 and c1 in (0, 16) and co == 16 and d.iw % 2 == 0 and d.iw // 2 + 2 <= 256


def conv_row_wgrad_workspace_bytes(d):
    return 64


def conv_row_wgrad(d, src0, src1, dy, ws, dw, accumulate=False, db_zero=None):
    conv_wgrad(d, src0, src1, dy, ws, dw, None, accumulate)
    if db_zero is not None and not accumulate:
        db_zero.zero_()


# ------------------------------------------------------------------ tcgen05 kernels (own weight packing)
def conv_umma_supported(d, for_dgrad=False):
    ok = (d.stride == 1 and d.kd == 1 and d.kh == 3 and d.kw == 3 and d.pd == 0 and d.ph == 1 and d.pw == 1 and
          d.id == 1 and d.c0 % 4 == 0 and d.c1 % 4 == 0)
    return bool(ok and (d.cout % 4 == 0 if for_dgrad else (d.cout <= 4 or d.cout % 4 == 0)))


def conv_umma_packed_floats(dgrad, O, I, T):
    rows, cols = (O, I) if dgrad else (I, O)
    return ((rows + 15) // 16) * T * 4 * _r16(cols) * 4


def conv_umma_pack_weights(w, out, dgrad, O, I, T):
    tmp = torch.empty(conv_tile_packed_floats(dgrad, O, I, T))
    conv_tile_pack_weights(w, tmp, dgrad, O, I, T)                    # [chunk][tap][colsP][16]
    rows, cols = (O, I) if dgrad else (I, O)
    chunks, colsP = (rows + 15) // 16, _r16(cols)
    out.copy_(tmp.reshape(chunks, T, colsP, 4, 4).permute(0, 1, 3, 2, 4).reshape(-1))    # -> [chunk][tap][kq][colsP][4]


def _umma_to_tile(wt, dgrad, O, I, T):
    rows, cols = (O, I) if dgrad else (I, O)
    chunks, colsP = (rows + 15) // 16, _r16(cols)
    return wt.reshape(chunks, T, 4, colsP, 4).permute(0, 1, 3, 2, 4).reshape(-1)


def conv_umma_fwd(d, src0, src1, wt, bias, dst, out_nchw=False):
    conv_tile_fwd(d, src0, src1, _umma_to_tile(wt, False, d.cout, d.c0 + d.c1, 9), bias, dst, out_nchw)


def conv_umma_dgrad(d, dy, wt_dgrad, dx0, dx1=None, accumulate=False):
    conv_tile_dgrad(d, dy, _umma_to_tile(wt_dgrad, True, d.cout, d.c0 + d.c1, 9), dx0, dx1, accumulate)


# ------------------------------------------------------------------ batched packing / first layer
def conv_pack_batch(jobs_dev, njobs, blocks_per_job=16, jobs_py=None):
    for (w, out, kind, mode, O, I, T) in jobs_py:
        if kind == 0:
            conv_pack_weights(w, out, mode, O, I, T)
        elif kind == 1:
            conv_tile_pack_weights(w, out, bool(mode), O, I, T)
        elif kind == 3:
            src, ok = _row_pack_index(mode, O, I, T)
            out.copy_(torch.where(ok, _tf32_rna(w.detach().reshape(-1))[src], torch.zeros(())))
        else:
            conv_umma_pack_weights(w, out, bool(mode), O, I, T)


def conv_c1_supported(d):
    return bool(d.c0 == 1 and d.c1 == 0 and d.stride == 1 and d.kh == 3 and d.kw == 3 and d.ph == 1 and d.pw == 1 and
                ((d.kd == 1 and d.pd == 0) or (d.kd == 3 and d.pd == 1)) and d.cout in (16, 32))


def conv_c1_fwd(d, x, w, bias, y):
    r = F.conv3d(_input(d, x.reshape(-1, 1), None), w.detach().reshape(d.cout, 1, d.kd, d.kh, d.kw), bias, **_kw(d))
    y.copy_(_ncdhw_to_cl(r).reshape(y.shape))


def conv_c1_wgrad_workspace_bytes(d):
    return 64


def conv_c1_wgrad(d, x, dy, ws, dw, db, accumulate=False):
    conv_wgrad(d, x.reshape(-1, 1), None, dy, ws, dw, db, accumulate)


# ------------------------------------------------------------------ norm / act / dropout
def bn_workspace_bytes(M, C):
    return 64


def bn_stats_fwd(y, M, C, gamma, beta, eps, momentum, running_mean, running_var, state, ws):
    v = y.reshape(M, C).double()
    mean = v.mean(0)
    var = v.var(0, unbiased=False)
    invstd = (1.0 / torch.sqrt(var + eps)).float()
    mean = mean.float()
    st = state.view(4, C)
    st[0], st[1] = mean, invstd
    st[2] = gamma.detach() * invstd
    st[3] = beta.detach() - mean * gamma.detach() * invstd
    if running_mean is not None:
        running_mean.mul_(1 - momentum).add_(momentum * mean)
        running_var.mul_(1 - momentum).add_(momentum * (var * M / max(M - 1, 1)).float())


def bn_finalize(part, nblocks, M, C, gamma, beta, eps, momentum, running_mean, running_var, state):
    pt = part.view(torch.float64)[:nblocks * 2 * C].reshape(nblocks, 2, C).sum(0)
    mean = pt[0] / M
    var = (pt[1] / M - mean * mean).clamp_min(0)
    invstd = (1.0 / torch.sqrt(var + eps)).float()
    mean = mean.float()
    st = state.view(4, C)
    st[0], st[1] = mean, invstd
    st[2] = gamma.detach() * invstd
    st[3] = beta.detach() - mean * gamma.detach() * invstd
    if running_mean is not None:
        running_mean.mul_(1 - momentum).add_(momentum * mean)
        running_var.mul_(1 - momentum).add_(momentum * (var * M / max(M - 1, 1)).float())


def bn_eval_state(C, gamma, beta, eps, running_mean, running_var, state):
    st = state.view(4, C)
    invstd = 1.0 / torch.sqrt(running_var + eps)
    st[0], st[1] = running_mean, invstd
    st[2] = gamma.detach() * invstd
    st[3] = beta.detach() - running_mean * gamma.detach() * invstd


def _mask(M, C, p_drop, drop_mode, seed, seed_off, rng_stream, spatial):
    if drop_mode == 0 or p_drop == 0:
        return torch.ones(M, C)
    s = seed + (int(seed_off.item()) if seed_off is not None else 0)
    return torch.from_numpy(philox.keep_mask(s, rng_stream, M, C, p_drop, drop_mode, spatial)) / (1.0 - p_drop)


def bn_act_fwd(y, state, a, M, C, slope, p_drop=0.0, drop_mode=0, seed=0, seed_off=None, rng_stream=0, spatial=1):
    st = state.view(4, C)
    z = y.reshape(M, C) * st[2] + st[3]
    r = torch.where(z > 0, z, z * slope) * _mask(M, C, p_drop, drop_mode, seed, seed_off, rng_stream, spatial)
    a.copy_(r.reshape(a.shape))


def bn_act_bwd(y, da, state, dy, dgamma, dbeta, M, C, slope, ws, p_drop=0.0, drop_mode=0, seed=0, seed_off=None,
               rng_stream=0, spatial=1, accumulate=False):
    st = state.view(4, C)
    yv = y.reshape(M, C)
    z = yv * st[2] + st[3]
    g = da.reshape(M, C) * _mask(M, C, p_drop, drop_mode, seed, seed_off, rng_stream, spatial)
    g = torch.where(z > 0, g, g * slope)
    xhat = (yv - st[0]) * st[1]
    sb, sg = g.double().sum(0), (g * xhat).double().sum(0)
    if dbeta is not None:
        _store(dbeta, sb.float(), accumulate)
    if dgamma is not None:
        _store(dgamma, sg.float(), accumulate)
    r = st[2] * (g - (sb / M).float() - xhat * (sg / M).float())
    dy.copy_(r.reshape(dy.shape))


def dropout_mask(mask, M, C, p_drop, drop_mode, seed, seed_off=None, rng_stream=0, spatial=1):
    mask.copy_((_mask(M, C, p_drop, drop_mode, seed, seed_off, rng_stream, spatial) > 0).float().reshape(mask.shape))


# ------------------------------------------------------------------ resampling / layout
def maxpool2_fwd(a, out, N, H, W, C):
    x = a.reshape(N, H, W, C).permute(0, 3, 1, 2)
    out.copy_(F.max_pool2d(x, 2).permute(0, 2, 3, 1).reshape(out.shape))


def maxpool2_bwd(a, dp, da, N, H, W, C, accumulate=False):
    with torch.enable_grad():
        x = a.reshape(N, H, W, C).permute(0, 3, 1, 2).detach().clone().requires_grad_(True)
        y = F.max_pool2d(x, 2)
        (g,) = torch.autograd.grad(y, x, dp.reshape(N, H // 2, W // 2, C).permute(0, 3, 1, 2))
    _store(da, g.permute(0, 2, 3, 1), accumulate)


def upsample2x_fwd(x, y, N, H, W, C):
    xi = x.reshape(N, H, W, C).permute(0, 3, 1, 2)
    y.copy_(F.interpolate(xi, scale_factor=2, mode="bilinear", align_corners=True).permute(0, 2, 3, 1).reshape(y.shape))


def upsample2x_bwd(dy, dx, N, H, W, C, accumulate=False):
    with torch.enable_grad():
        xi = torch.zeros(N, C, H, W, requires_grad=True)
        y = F.interpolate(xi, scale_factor=2, mode="bilinear", align_corners=True)
        (g,) = torch.autograd.grad(y, xi, dy.reshape(N, 2 * H, 2 * W, C).permute(0, 3, 1, 2))
    _store(dx, g.permute(0, 2, 3, 1), accumulate)


def nchw_to_nhwc(src, dst, N, C, S):
    dst.copy_(src.reshape(N, C, S).permute(0, 2, 1).reshape(dst.shape))


def nhwc_to_nchw(src, dst, N, C, S):
    dst.copy_(src.reshape(N, S, C).permute(0, 2, 1).reshape(dst.shape))


def colsum_workspace_bytes(M, C):
    return 64


def colsum(g, M, C, out, ws, accumulate=False):
    _store(out, g.reshape(M, C).sum(0), accumulate)


def add(a, b, c):
    c.copy_(a + b)


# ------------------------------------------------------------------ nn.Linear (gemm_umma.cu)
def linear_supported(M, O, c0, c1):
    return (M > 0 and O >= 16 and O % 4 == 0 and c0 > 0 and c0 % 4 == 0 and c1 >= 0 and c1 % 4 == 0
            and (c1 == 0 or c0 % 32 == 0))


def linear_fwd(x0, x1, w, bias, y, M, O):
    x = x0 if x1 is None else torch.cat([x0, x1], 1)
    out = x @ w.t()
    y.copy_(out + bias if bias is not None else out)


def linear_dgrad(dy, w, dx0, dx1, accumulate, M, O):
    dx = dy @ w
    c0 = dx0.shape[-1]
    _store(dx0, dx[:, :c0], accumulate)
    if dx1 is not None:
        _store(dx1, dx[:, c0:], accumulate)


def linear_wgrad_workspace_bytes(M, O, I):
    return 64


def linear_wgrad(x0, x1, dy, dw, ws, M, O, accumulate=False):
    x = x0 if x1 is None else torch.cat([x0, x1], 1)
    _store(dw, dy.t() @ x, accumulate)


# ------------------------------------------------------------------ UNETR ops (vit.cu)
def patch3d_gather(x, y, B, C, D, H, W, patch):
    P = patch
    t = x.reshape(B, C, D // P, P, H // P, P, W // P, P).permute(0, 2, 4, 6, 3, 5, 7, 1)      # b h w d p1 p2 p3 c
    y.copy_(t.reshape(y.shape))


def mha_probs_floats(B, N, heads):
    return B * heads * N * N


def _mha(qkv, B, N, heads, hd):
    C = heads * hd
    q, k, v = qkv.reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    att = ((q @ k.transpose(-1, -2)) * hd ** -0.5).softmax(-1)
    return (att @ v).permute(0, 2, 1, 3).reshape(B * N, C), att


def mha_fwd(qkv, out, probs, B, N, heads, hd):
    o, att = _mha(qkv, B, N, heads, hd)
    out.copy_(o)
    if probs is not None:
        probs[:B * heads * N * N].copy_(att.reshape(-1))


def mha_bwd(qkv, probs, dout, dqkv, ws, B, N, heads, hd):
    leaf = qkv.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        o, _ = _mha(leaf, B, N, heads, hd)
        g, = torch.autograd.grad(o, leaf, dout.reshape(o.shape))
    dqkv.copy_(g)


def add_lrelu_fwd(a, b, out, slope):
    out.copy_(F.leaky_relu(a + b, slope))


def lrelu_bwd(out, dout, dx, slope):
    dx.copy_(torch.where(out > 0, dout, dout * slope))


# ------------------------------------------------------------------ Swin-UNet token ops
def layernorm_workspace_bytes(M, C):
    return 64


def layernorm_fwd(x, gamma, beta, y, stats, M, C, eps=1e-5):
    v = x.reshape(M, C)
    mean = v.mean(1, keepdim=True)
    var = v.var(1, unbiased=False, keepdim=True)
    rstd = torch.rsqrt(var + eps)
    y.copy_(((v - mean) * rstd * gamma.detach() + beta.detach()).reshape(y.shape))
    if stats is not None:
        stats.view(M, 2)[:, 0] = mean[:, 0]
        stats.view(M, 2)[:, 1] = rstd[:, 0]


def layernorm_bwd(x, stats, gamma, dy, dx, dgamma, dbeta, M, C, ws, accumulate_dx=False):
    v, g = x.reshape(M, C), dy.reshape(M, C)
    mean, rstd = stats.view(M, 2)[:, 0:1], stats.view(M, 2)[:, 1:2]
    xh = (v - mean) * rstd
    gg = g * gamma.detach()
    r = rstd * (gg - gg.mean(1, keepdim=True) - xh * (gg * xh).mean(1, keepdim=True))
    _store(dx, r, accumulate_dx)
    dgamma.copy_((g * xh).sum(0))
    dbeta.copy_(g.sum(0))


def gelu_fwd(x, y):
    y.copy_(F.gelu(x))


def gelu_bwd(x, dy, dx, accumulate=False):
    with torch.enable_grad():
        xi = x.detach().clone().requires_grad_(True)
        (g,) = torch.autograd.grad(F.gelu(xi), xi, dy)
    _store(dx, g, accumulate)


def _window_attention(qkv, table, B, H, W, C, heads, ws, shift):
    """Reference formulation (roll / partition / attention / reverse / roll) on a [B*H*W, 3C] qkv matrix."""
    hd, N = C // heads, ws * ws
    x = qkv.reshape(B, H, W, 3 * C)
    if shift > 0:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
    xw = x.view(B, H // ws, ws, W // ws, ws, 3 * C).permute(0, 1, 3, 2, 4, 5).reshape(-1, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = xw[0] * hd ** -0.5, xw[1], xw[2]
    attn = q @ k.transpose(-2, -1)
    ch = torch.arange(ws)
    coords = torch.stack(torch.meshgrid([ch, ch], indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0) + (ws - 1)
    index = rel[:, :, 0] * (2 * ws - 1) + rel[:, :, 1]
    attn = attn + table[index.view(-1)].view(N, N, heads).permute(2, 0, 1).unsqueeze(0)
    if shift > 0:
        img = torch.zeros(1, H, W, 1)
        cnt = 0
        for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
                img[:, hs, wsl, :] = cnt
                cnt += 1
        mw = img.view(1, H // ws, ws, W // ws, ws, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, N)
        am = mw.unsqueeze(1) - mw.unsqueeze(2)
        am = am.masked_fill(am != 0, -100.0).masked_fill(am == 0, 0.0)
        nW = am.shape[0]
        attn = (attn.view(-1, nW, heads, N, N) + am.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
    attn = attn.softmax(-1)
    o = (attn @ v).transpose(1, 2).reshape(-1, ws, ws, C)
    o = o.view(B, H // ws, W // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)
    if shift > 0:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    return o.reshape(B * H * W, C)


def window_attn_fwd(qkv, table, out, B, H, W, C, heads, ws, shift):
    out.copy_(_window_attention(qkv.detach(), table.detach(), B, H, W, C, heads, ws, shift).reshape(out.shape))


def window_attn_workspace_bytes(B, H, W, heads, ws):
    return 64


def window_attn_bwd(qkv, table, dout, dqkv, dtable, B, H, W, C, heads, ws, shift, wsp):
    with torch.enable_grad():
        q = qkv.detach().clone().requires_grad_(True)
        t = table.detach().clone().requires_grad_(True)
        o = _window_attention(q, t, B, H, W, C, heads, ws, shift)
        gq, gt = torch.autograd.grad(o, [q, t], dout.reshape(o.shape))
    dqkv.copy_(gq.reshape(dqkv.shape))
    dtable.copy_(gt.reshape(dtable.shape))


def _droppath_scale(B, p_drop, seed, seed_off, rng_stream):
    if p_drop == 0:
        return torch.ones(B)
    s = seed + (int(seed_off.item()) if seed_off is not None else 0)
    w0 = philox.philox4x32_10(s, rng_stream, np.arange(B, dtype=np.uint64))[0]
    u = (w0 >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return torch.from_numpy((u >= np.float32(p_drop)).astype(np.float32)) / (1.0 - p_drop)


def add_droppath(x, branch, out, B, per_sample, p_drop=0.0, seed=0, seed_off=None, rng_stream=0):
    sc = _droppath_scale(B, p_drop, seed, seed_off, rng_stream).reshape(B, 1)
    r = branch.reshape(B, -1) * sc
    if x is not None:
        r = r + x.reshape(B, -1)
    out.copy_(r.reshape(out.shape))


def patch_merge_gather(x, y, B, H, W, C, inverse=False, accumulate=False):
    if not inverse:
        v = x.reshape(B, H, W, C)
        g = torch.cat([v[:, 0::2, 0::2], v[:, 1::2, 0::2], v[:, 0::2, 1::2], v[:, 1::2, 1::2]], -1)
        y.copy_(g.reshape(y.shape))
    else:
        g = x.reshape(B, H // 2, W // 2, 4, C)
        r = torch.zeros(B, H, W, C)
        r[:, 0::2, 0::2], r[:, 1::2, 0::2], r[:, 0::2, 1::2], r[:, 1::2, 1::2] = g[..., 0, :], g[..., 1, :], g[..., 2, :], g[..., 3, :]
        _store(y, r, accumulate)


def pixel_shuffle(x, y, B, H, W, C, p, inverse=False):
    if not inverse:
        v = x.reshape(B, H, W, p, p, C).permute(0, 1, 3, 2, 4, 5)
        y.copy_(v.reshape(y.shape))
    else:
        v = x.reshape(B, H, p, W, p, C).permute(0, 1, 3, 2, 4, 5)
        y.copy_(v.reshape(y.shape))


def patch_embed_gather(x, y, B, H, W, patch, repeat_channels):
    v = x.reshape(B, 1, H, W).repeat(1, repeat_channels, 1, 1)
    cols = F.unfold(v, kernel_size=patch, stride=patch)                  # [B, C*P*P, L]
    y.copy_(cols.transpose(1, 2).reshape(y.shape))


# ------------------------------------------------------------------ loss / optimizer / noise
def ssl_loss_workspace_bytes(B, S):
    return 64


def _to_ncs(t, nhwc, n, C, S):
    return t.reshape(n, S, C).permute(0, 2, 1) if nhwc else t.reshape(n, C, S)


def _ssl_loss(logits, teacher, labels, nhwc, B, Lb, C, S, w, mc_psum=None, mc_T=0.0, mc_thr=None):
    from oracle import ssl_oracle as O
    lg = _to_ncs(logits, nhwc, B, C, S)
    tl = _to_ncs(teacher, nhwc, B - Lb, C, S) if teacher is not None else None
    lab = labels.reshape(-1, S)[:Lb] if labels is not None else None
    if Lb > 0:
        sup, ce, dice = O.supervised_loss(lg[:Lb], lab, C)
    else:
        sup = ce = dice = torch.zeros(())
    cons = torch.zeros(())
    if tl is not None and B > Lb:
        dist = (torch.softmax(lg[Lb:], 1) - torch.softmax(tl, 1)) ** 2
        if mc_psum is None:
            cons = torch.mean(dist)
        else:
            pbar = _to_ncs(mc_psum, nhwc, B - Lb, C, S) / mc_T
            unc = -torch.sum(pbar * torch.log(pbar + 1e-6), dim=1, keepdim=True)
            mask = (unc < float(mc_thr[0])).float()
            cons = torch.sum(mask * dist) / (2 * torch.sum(mask) + 1e-16)
    return sup + w * cons, ce, dice, cons


def ssl_loss_fwd(logits, teacher, labels, nhwc, B, Lb, C, S, w_cons, lossbuf, ws, mc_psum=None, mc_T=0.0, mc_thr=None):
    w = float(w_cons[0]) if w_cons is not None else 0.0
    tot, ce, dice, cons = _ssl_loss(logits.detach(), teacher, labels, nhwc, B, Lb, C, S, w, mc_psum, mc_T, mc_thr)
    lossbuf[0], lossbuf[1], lossbuf[2], lossbuf[3] = ce, dice, cons, tot


def ssl_loss_bwd(logits, teacher, labels, nhwc, B, Lb, C, S, w_cons, lossbuf, grad_scale, dlogits, dlogits_nhwc,
                 mc_psum=None, mc_T=0.0, mc_thr=None):
    w = float(w_cons[0]) if (w_cons is not None and teacher is not None) else 0.0
    with torch.enable_grad():
        lg = logits.detach().clone().requires_grad_(True)
        tot, *_ = _ssl_loss(lg, teacher, labels, nhwc, B, Lb, C, S, w, mc_psum, mc_T, mc_thr)
        (g,) = torch.autograd.grad(tot * grad_scale, lg)
    g = _to_ncs(g, nhwc, B, C, S)
    dlogits.copy_((g.permute(0, 2, 1) if dlogits_nhwc else g).reshape(dlogits.shape))


def _ct_loss(lg, other, lab, Lb, C, w, kind="dice"):
    from oracle import ssl_oracle as O
    sup, ce, dice = O.supervised_loss(lg[:Lb], lab, C)
    pseudo = torch.argmax(torch.softmax(other[Lb:], 1), dim=1)
    if kind == "ce":
        ps = F.cross_entropy(lg[Lb:], pseudo)
    else:
        ps = O.dice_loss_multiclass(torch.softmax(lg[Lb:], 1), pseudo.unsqueeze(1), C)
    return sup + w * ps, ce, dice, ps


def ct_loss_fwd(logits, nhwc, other, other_nhwc, labels, B, Lb, C, S, w_cons, lossbuf, ws, kind="dice"):
    w = float(w_cons[0])
    lg, ot = _to_ncs(logits.detach(), nhwc, B, C, S), _to_ncs(other.detach(), other_nhwc, B, C, S)
    tot, ce, dice, ps = _ct_loss(lg, ot, labels.reshape(-1, S)[:Lb], Lb, C, w, kind)
    lossbuf[0], lossbuf[1], lossbuf[2], lossbuf[3], lossbuf[4] = ce, dice, ps, tot, w


def ct_loss_bwd(logits, nhwc, other, other_nhwc, labels, B, Lb, C, S, lossbuf, grad_scale, dlogits, dlogits_nhwc, kind="dice"):
    w = float(lossbuf[4])
    with torch.enable_grad():
        raw = logits.detach().clone().requires_grad_(True)
        tot, *_ = _ct_loss(_to_ncs(raw, nhwc, B, C, S), _to_ncs(other.detach(), other_nhwc, B, C, S),
                           labels.reshape(-1, S)[:Lb], Lb, C, w, kind)
        (g,) = torch.autograd.grad(tot * grad_scale, raw)
    g = _to_ncs(g, nhwc, B, C, S)
    dlogits.copy_((g.permute(0, 2, 1) if dlogits_nhwc else g).reshape(dlogits.shape))


def cps_loss_fwd(*a):
    ct_loss_fwd(*a, kind="ce")


def cps_loss_bwd(*a):
    ct_loss_bwd(*a, kind="ce")


def mc_softmax_accumulate(logits, psum, R, U, C, S, nhwc=False, init=True):
    lg = _to_ncs(logits, nhwc, R * U, C, S)
    acc = torch.softmax(lg, 1).reshape(R, U, C, S).sum(0)
    acc = (acc.permute(0, 2, 1) if nhwc else acc).reshape(psum.shape)
    if init:
        psum.copy_(acc)
    else:
        psum += acc


def sgd_ema_step(params, grads, momentum_buf, ema_params, hparams, zero_grad=False):
    lr, mu, wd, alpha, oma, gs = [float(v) for v in hparams[:6]]
    with torch.no_grad():
        g = grads * gs + wd * params
        momentum_buf.mul_(mu).add_(g)
        params.add_(momentum_buf, alpha=-lr)
        if ema_params is not None:
            ema_params.mul_(alpha).add_(params, alpha=oma)
        if zero_grad:
            grads.zero_()


def ema_update(ema_params, params, hparams):
    with torch.no_grad():
        ema_params.mul_(float(hparams[3])).add_(params, alpha=float(hparams[4]))


# ------------------------------------------------------------------ stand-alone losses (utils/losses.py drop-ins)
def loss_dropin_workspace_bytes(B, S):
    return 64


def _dice_terms(x, use_softmax, labels, B, C, S):
    p = x.reshape(B, C, S).double()
    if use_softmax:
        p = torch.softmax(p, dim=1)
    t = F.one_hot(labels.reshape(B, S).long(), C).permute(0, 2, 1).double()
    return p, t, (p * t).sum((0, 2)), (p * p).sum((0, 2)), (t * t).sum((0, 2))


def dice_fwd(x, use_softmax, labels, B, C, S, weight, out, ws):
    p, t, I, Z, Y = _dice_terms(x, use_softmax, labels, B, C, S)
    d = 1 - (2 * I + 1e-5) / (Z + Y + 1e-5)
    w = torch.ones(C, dtype=torch.float64) if weight is None else weight.double()
    out.zero_()
    out[0] = float((d * w).sum() / C)
    out[1:1 + C] = (1 - d).float()
    out[9:9 + C], out[17:17 + C], out[25:25 + C] = I.float(), Z.float(), Y.float()


def dice_bwd(x, use_softmax, labels, B, C, S, weight, fwd_out, grad_out, dx):
    p, t, I, Z, Y = _dice_terms(x, use_softmax, labels, B, C, S)
    w = (torch.ones(C, dtype=torch.float64) if weight is None else weight.double()) / C
    N, D = (2 * I + 1e-5).view(1, C, 1), (Z + Y + 1e-5).view(1, C, 1)
    g = -w.view(1, C, 1) * (2 * t * D - N * 2 * p) / (D * D)
    if use_softmax:
        g = p * (g - (p * g).sum(1, keepdim=True))
    dx.copy_((g * float(grad_out.reshape(-1)[0])).float().reshape(dx.shape))


def softmax_mse_fwd(a, b, B, C, S, out):
    out.copy_(((torch.softmax(a.reshape(B, C, S), 1) - torch.softmax(b.reshape(B, C, S), 1)) ** 2).reshape(out.shape))


def softmax_mse_bwd(a, b, grad_out, B, C, S, da):
    pa, pb = torch.softmax(a.reshape(B, C, S).double(), 1), torch.softmax(b.reshape(B, C, S).double(), 1)
    q = 2 * (pa - pb) * grad_out.reshape(B, C, S).double()
    da.copy_((pa * (q - (pa * q).sum(1, keepdim=True))).float().reshape(da.shape))


def softmax_kl_fwd(a, b, B, C, S, out, ws):
    la, lb = torch.log_softmax(a.reshape(B, C, S).double(), 1), torch.log_softmax(b.reshape(B, C, S).double(), 1)
    out[0] = float((lb.exp() * (lb - la)).mean())


def softmax_kl_bwd(a, b, grad_out, B, C, S, da):
    pa, pb = torch.softmax(a.reshape(B, C, S).double(), 1), torch.softmax(b.reshape(B, C, S).double(), 1)
    da.copy_(((pa - pb) * float(grad_out.reshape(-1)[0]) / (B * C * S)).float().reshape(da.shape))


def noise_add(x, out, sigma, clip, seed, seed_off=None, rng_stream=0):
    s = seed + (int(seed_off.item()) if seed_off is not None else 0)
    nz = torch.from_numpy(philox.clamp_noise(s, rng_stream, out.numel(), sigma, clip)).reshape(out.shape)
    out.copy_(nz if x is None else x + nz)


def install(monkeypatch):
    """Replace every function of cv_ssl_mis_b200.ops with the CPU stand-in of the same name."""
    import sys
    me = sys.modules[__name__]
    for name in dir(real_ops):
        if name.startswith("_") or not callable(getattr(real_ops, name)) or name in ("ConvDesc", "B200Error"):
            continue
        if hasattr(me, name):
            monkeypatch.setattr(real_ops, name, torch.no_grad()(getattr(me, name)))
        elif name not in ("conv_desc", "desc_out_dims"):
            def missing(*a, _n=name, **k):
                raise NotImplementedError(f"fake_ops has no stand-in for ops.{_n}")
            monkeypatch.setattr(real_ops, name, missing)
