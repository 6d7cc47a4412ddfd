"""Swin-UNet (`ViT_seg`) on the B200 C ABI: the transformer half of the Cross-Teaching step.

Drop-in for code/networks/vision_transformer.py:24-52 (`SwinUnet`) and the `SwinTransformerSys` it wraps
(code/networks/swin_transformer_unet_skip_expand_decoder_sys.py:599-793): same constructor meaning, same
state_dict keys (so reference checkpoints load), same forward semantics in train and eval mode.

Nothing here computes with torch: the nn.Module tree only holds parameters/buffers under the reference's names.
`SwinPlan` turns one (B, need_grad) geometry into a static tape of C-ABI launches over token matrices
`[B*H*W, C]` (channels-last == the reference's `B, L, C`):

  Linear / 1x1 conv -> implicit-GEMM conv entry points (a token row is a 1x1 "pixel")
  LayerNorm, GELU, shifted-window attention (roll + partition + rel-pos bias + mask + softmax + PV + reverse in one
  kernel), DropPath residual, PatchMerging gather, PatchExpand rearrange -> swin.cu kernels.

The backward tape is derived at plan time by walking the forward tape in reverse: gradient buffers come from a
free list with static liveness (a gradient dies when its producer's backward has run), residual adds alias
instead of copying, and every kernel is told statically whether it overwrites or accumulates.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops, _lib
from .._lib import PACK_CONV_FWD, PACK_CONV_DGRAD
from ._engine import FlatParams, Runtime, PackTable

DROPPATH_STREAM = 3000           # Philox stream ids of the DropPath draws: DROPPATH_STREAM + 2 * block + {0, 1}


# ===================================================================================== parameter containers
def _trunc_normal_(t, std=0.02):
    return nn.init.trunc_normal_(t, std=std)


class Mlp(nn.Module):            # …_sys.py:9-25
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class WindowAttention(nn.Module):   # …_sys.py:76-113
    def __init__(self, dim, window_size, num_heads, qkv_bias=True):
        super().__init__()
        ws = window_size
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) * (2 * ws - 1), num_heads))
        ch = torch.arange(ws)
        coords = torch.stack(torch.meshgrid([ch, ch], indexing="ij")).flatten(1)
        rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += ws - 1
        rel[:, :, 1] += ws - 1
        rel[:, :, 0] *= 2 * ws - 1
        self.register_buffer("relative_position_index", rel.sum(-1))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        _trunc_normal_(self.relative_position_bias_table)


def _shift_mask(H, W, ws, shift):
    """The SW-MSA mask buffer of the reference (…_sys.py:212-232); kept for state_dict parity, the kernel derives it."""
    img = torch.zeros((1, H, W, 1))
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = img.view(1, H // ws, ws, W // ws, ws, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, -100.0).masked_fill(am == 0, 0.0)


class SwinTransformerBlock(nn.Module):   # …_sys.py:174-237
    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4.0, qkv_bias=True,
                 drop_path=0.0):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads = dim, tuple(input_resolution), num_heads
        self.window_size, self.shift_size, self.drop_path_rate = window_size, shift_size, float(drop_path)
        if min(self.input_resolution) <= self.window_size:
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, self.window_size, num_heads, qkv_bias)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        mask = _shift_mask(*self.input_resolution, self.window_size, self.shift_size) if self.shift_size > 0 else None
        self.register_buffer("attn_mask", mask)


class PatchMerging(nn.Module):   # …_sys.py:309-346
    def __init__(self, input_resolution, dim):
        super().__init__()
        self.input_resolution, self.dim = tuple(input_resolution), dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = nn.LayerNorm(4 * dim)


class PatchExpand(nn.Module):    # …_sys.py:359-382
    def __init__(self, input_resolution, dim, dim_scale=2):
        super().__init__()
        self.input_resolution, self.dim = tuple(input_resolution), dim
        self.expand = nn.Linear(dim, 2 * dim, bias=False) if dim_scale == 2 else nn.Identity()
        self.norm = nn.LayerNorm(dim // dim_scale)


class FinalPatchExpand_X4(nn.Module):   # …_sys.py:385-410
    def __init__(self, input_resolution, dim, dim_scale=4):
        super().__init__()
        self.input_resolution, self.dim, self.dim_scale = tuple(input_resolution), dim, dim_scale
        self.expand = nn.Linear(dim, 16 * dim, bias=False)
        self.norm = nn.LayerNorm(dim)


def _blocks(dim, res, depth, heads, window, mlp_ratio, qkv_bias, drop_path):
    return nn.ModuleList([SwinTransformerBlock(dim, res, heads, window, 0 if i % 2 == 0 else window // 2, mlp_ratio, qkv_bias,
                                               drop_path[i] if isinstance(drop_path, list) else drop_path)
                          for i in range(depth)])


class BasicLayer(nn.Module):     # …_sys.py:413-472
    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio, qkv_bias, drop_path, downsample):
        super().__init__()
        self.blocks = _blocks(dim, input_resolution, depth, num_heads, window_size, mlp_ratio, qkv_bias, drop_path)
        self.downsample = PatchMerging(input_resolution, dim) if downsample else None


class BasicLayer_up(nn.Module):  # …_sys.py:487-546
    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio, qkv_bias, drop_path, upsample):
        super().__init__()
        self.blocks = _blocks(dim, input_resolution, depth, num_heads, window_size, mlp_ratio, qkv_bias, drop_path)
        self.upsample = PatchExpand(input_resolution, dim, 2) if upsample else None


class PatchEmbed(nn.Module):     # …_sys.py:549-588
    def __init__(self, img_size, patch_size, in_chans, embed_dim, patch_norm=True):
        super().__init__()
        self.img_size, self.patch_size = (img_size, img_size), (patch_size, patch_size)
        self.patches_resolution = [img_size // patch_size, img_size // patch_size]
        self.num_patches = self.patches_resolution[0] * self.patches_resolution[1]
        self.in_chans, self.embed_dim = in_chans, embed_dim
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = nn.LayerNorm(embed_dim) if patch_norm else None


class SwinTransformerSys(nn.Module):
    """Parameter tree of …_sys.py:624-722 (ape=False, drop=attn_drop=0 as every reference config sets them)."""

    def __init__(self, img_size=224, patch_size=4, in_chans=3, num_classes=1000, embed_dim=96, depths=(2, 2, 2, 2),
                 num_heads=(3, 6, 12, 24), window_size=7, mlp_ratio=4.0, qkv_bias=True, drop_path_rate=0.1,
                 patch_norm=True):
        super().__init__()
        depths, num_heads = list(depths), list(num_heads)
        assert len(depths) == 4, "the decoder indexes the skips as 3 - inx (…_sys.py:768): exactly four stages"
        self.num_classes, self.num_layers, self.embed_dim = num_classes, len(depths), embed_dim
        self.depths, self.num_heads, self.window_size, self.mlp_ratio = depths, num_heads, window_size, mlp_ratio
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim, patch_norm)
        pr = self.patches_resolution = self.patch_embed.patches_resolution
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        nl = self.num_layers
        self.layers = nn.ModuleList()
        for i in range(nl):
            self.layers.append(BasicLayer(embed_dim * 2 ** i, (pr[0] // 2 ** i, pr[1] // 2 ** i), depths[i], num_heads[i],
                                          window_size, mlp_ratio, qkv_bias, dpr[sum(depths[:i]):sum(depths[:i + 1])], i < nl - 1))
        self.layers_up = nn.ModuleList()
        self.concat_back_dim = nn.ModuleList()
        for i in range(nl):
            j = nl - 1 - i
            dim, res = embed_dim * 2 ** j, (pr[0] // 2 ** j, pr[1] // 2 ** j)
            self.concat_back_dim.append(nn.Linear(2 * dim, dim) if i > 0 else nn.Identity())
            if i == 0:
                self.layers_up.append(PatchExpand(res, dim, 2))
            else:
                self.layers_up.append(BasicLayer_up(dim, res, depths[j], num_heads[j], window_size, mlp_ratio, qkv_bias,
                                                    dpr[sum(depths[:j]):sum(depths[:j + 1])], i < nl - 1))
        self.norm = nn.LayerNorm(embed_dim * 2 ** (nl - 1))
        self.norm_up = nn.LayerNorm(embed_dim)
        self.up = FinalPatchExpand_X4((img_size // patch_size, img_size // patch_size), embed_dim, 4)
        self.output = nn.Conv2d(embed_dim, num_classes, kernel_size=1, bias=False)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):        # …_sys.py:724-731
        if isinstance(m, nn.Linear):
            _trunc_normal_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)


# ===================================================================================== plan-time values
class _Val:
    """One token matrix [M, C] of the forward pass and (during backward) its gradient."""

    def __init__(self, dev, M, C, alloc=True):
        self.M, self.C = M, C
        self.v = torch.empty((M, C), dtype=torch.float32, device=dev) if alloc else None
        self.g = None            # _GBuf once some backward op has produced (part of) the gradient


class _GBuf:
    def __init__(self, t):
        self.t, self.rc = t, 1


class _GradPool:
    """Static-liveness allocator for gradient buffers: sizes are exact, reuse follows the backward order."""

    def __init__(self, dev):
        self.dev, self.free, self.total = dev, {}, 0

    def take(self, numel):
        lst = self.free.get(numel)
        if lst:
            b = lst.pop()
            b.rc = 1
            return b
        self.total += numel
        return _GBuf(torch.empty(numel, dtype=torch.float32, device=self.dev))

    def release(self, b):
        b.rc -= 1
        assert b.rc >= 0
        if b.rc == 0:
            self.free.setdefault(b.t.numel(), []).append(b)


# ===================================================================================== tape operations
class _Op:
    name = ""

    def fwd(self, rt, train): ...
    def prep_bwd(self, pool): ...
    def bwd(self, rt): ...

    # gradient of an input: returns (tensor, accumulate) and registers the buffer on the value
    @staticmethod
    def _gin(pool, val):
        if val.g is None:
            val.g = pool.take(val.M * val.C)
            return val.g.t.view(val.M, val.C), False
        return val.g.t.view(val.M, val.C), True

    @staticmethod
    def _gout(pool, val):
        """Gradient of an output (must exist: everything on the tape reaches the logits)."""
        assert val.g is not None, "value has no consumer on the backward tape"
        b = val.g
        return b, b.t.view(val.M, val.C)


class _Linear(_Op):
    """y = [x0 | x1] W^T + b as a 1x1 implicit-GEMM convolution over token rows."""

    def __init__(self, rt, weight, bias, B, x0, x1, y, need_grad, name, out_nchw=False, input_grad=True):
        self.w, self.b, self.x0, self.x1, self.y, self.name = weight, bias, x0, x1, y, name
        self.out_nchw, self.input_grad = out_nchw, input_grad
        O, I = weight.shape[0], weight[0].numel()
        c0, c1 = x0.C, (x1.C if x1 is not None else 0)
        assert c0 + c1 == I and x0.M % B == 0, (name, c0, c1, I)
        self.O, self.I, self.M = O, I, x0.M
        # tcgen05/TMA GEMM straight from the module's row-major weight (no packing); the implicit-GEMM convolution
        # engine keeps the NCHW-writing head and the `exact` (3xTF32) validation mode
        self.umma = (not rt.exact) and (not out_nchw) and ops.linear_supported(x0.M, O, c0, c1)
        self.wp_fwd = self.wp_bwd = None
        if self.umma:
            if need_grad:
                rt.need_scratch(max(ops.linear_wgrad_workspace_bytes(x0.M, O, I), ops.colsum_workspace_bytes(x0.M, O)))
            return
        self.desc = ops.conv_desc(B, 1, 1, x0.M // B, c0, c1, O, 1, 1, 0, 2)
        self.wp_fwd = torch.empty(ops.conv_packed_floats(PACK_CONV_FWD, O, I, 1), dtype=torch.float32, device=rt.device)
        if need_grad:
            rt.need_scratch(ops.conv_wgrad_workspace_bytes(self.desc))
            if input_grad:
                self.wp_bwd = torch.empty(ops.conv_packed_floats(PACK_CONV_DGRAD, O, I, 1), dtype=torch.float32, device=rt.device)

    def pack_jobs(self, need_dgrad):
        if self.umma:
            return []
        jobs = [(self.w, self.wp_fwd, 0, PACK_CONV_FWD, self.O, self.I, 1)]
        if need_dgrad and self.wp_bwd is not None:
            jobs.append((self.w, self.wp_bwd, 0, PACK_CONV_DGRAD, self.O, self.I, 1))
        return jobs

    def fwd(self, rt, train):
        _lib.tag = self.name
        if self.umma:
            ops.linear_fwd(self.x0.v, self.x1.v if self.x1 is not None else None, self.w.view(self.O, self.I), self.b, self.y.v,
                           self.M, self.O)
            return
        ops.conv_fwd(self.desc, self.x0.v, self.x1.v if self.x1 is not None else None, self.wp_fwd, self.b, self.y.v,
                     self.out_nchw, rt.exact)

    def prep_bwd(self, pool):
        self.gy_buf, self.gy = self._gout(pool, self.y)
        if self.input_grad:
            self.gx0, acc0 = self._gin(pool, self.x0)
            self.gx1, acc1 = self._gin(pool, self.x1) if self.x1 is not None else (None, False)
            # the two halves of a virtual concat share one accumulate flag in the dgrad kernel
            assert self.x1 is None or acc0 == acc1, "virtual-concat halves need the same overwrite/accumulate state"
            self.acc0 = acc0
        pool.release(self.gy_buf)

    def bwd(self, rt):
        _lib.tag = self.name
        x1 = self.x1.v if self.x1 is not None else None
        if self.umma:
            # weight / bias gradient on the side stream next to the data gradient; joined before the next tape
            # op because the gradient pool may hand this op's dy buffer to a later op
            with rt.side_stream():
                ops.linear_wgrad(self.x0.v, x1, self.gy, self.w.grad.view(self.O, self.I), rt.scratch_side, self.M, self.O)
                if self.b is not None:
                    ops.colsum(self.gy, self.M, self.O, self.b.grad, rt.scratch_side)
            if self.input_grad:
                ops.linear_dgrad(self.gy, self.w.view(self.O, self.I), self.gx0, self.gx1, self.acc0, self.M, self.O)
            rt.join_side()
            return
        ops.conv_wgrad(self.desc, self.x0.v, x1, self.gy, rt.scratch, self.w.grad, self.b.grad if self.b is not None else None,
                       False, rt.exact)
        if self.input_grad:
            ops.conv_dgrad(self.desc, self.gy, self.wp_bwd, self.gx0, self.gx1, self.acc0, rt.exact)


class _LayerNorm(_Op):
    def __init__(self, rt, ln, x, y, need_grad, name):
        self.ln, self.x, self.y, self.name = ln, x, y, name
        self.stats = torch.empty(2 * x.M, dtype=torch.float32, device=rt.device) if need_grad else None
        if need_grad:
            rt.need_scratch(ops.layernorm_workspace_bytes(x.M, x.C))

    def fwd(self, rt, train):
        _lib.tag = self.name
        ops.layernorm_fwd(self.x.v, self.ln.weight, self.ln.bias, self.y.v, self.stats, self.x.M, self.x.C, self.ln.eps)

    def prep_bwd(self, pool):
        self.gy_buf, self.gy = self._gout(pool, self.y)
        self.gx, self.acc = self._gin(pool, self.x)
        pool.release(self.gy_buf)

    def bwd(self, rt):
        _lib.tag = self.name
        ops.layernorm_bwd(self.x.v, self.stats, self.ln.weight, self.gy, self.gx, self.ln.weight.grad, self.ln.bias.grad,
                          self.x.M, self.x.C, rt.scratch, self.acc)


class _Gelu(_Op):
    def __init__(self, x, y, name):
        self.x, self.y, self.name = x, y, name

    def fwd(self, rt, train):
        _lib.tag = self.name
        ops.gelu_fwd(self.x.v, self.y.v)

    def prep_bwd(self, pool):
        self.gy_buf, self.gy = self._gout(pool, self.y)
        self.gx, self.acc = self._gin(pool, self.x)
        pool.release(self.gy_buf)

    def bwd(self, rt):
        _lib.tag = self.name
        ops.gelu_bwd(self.x.v, self.gy, self.gx, self.acc)


class _WindowAttn(_Op):
    def __init__(self, rt, attn, qkv, out, B, H, W, heads, ws, shift, need_grad, name):
        self.attn, self.qkv, self.out, self.name = attn, qkv, out, name
        self.geom = (B, H, W, out.C, heads, ws, shift)
        if need_grad:
            rt.need_scratch(ops.window_attn_workspace_bytes(B, H, W, heads, ws))

    def fwd(self, rt, train):
        _lib.tag = self.name
        ops.window_attn_fwd(self.qkv.v, self.attn.relative_position_bias_table, self.out.v, *self.geom)

    def prep_bwd(self, pool):
        self.gy_buf, self.gy = self._gout(pool, self.out)
        self.gx, acc = self._gin(pool, self.qkv)
        assert not acc
        pool.release(self.gy_buf)

    def bwd(self, rt):
        _lib.tag = self.name
        t = self.attn.relative_position_bias_table
        ops.window_attn_bwd(self.qkv.v, t, self.gy, self.gx, t.grad, *self.geom, rt.scratch)


class _AddDropPath(_Op):
    """out = x + DropPath(branch) (…_sys.py:284-285; timm DropPath: per-sample Bernoulli keep, scaled by 1/keep)."""

    def __init__(self, x, branch, out, B, p, stream, name):
        self.x, self.branch, self.out, self.B, self.p, self.stream, self.name = x, branch, out, B, float(p), stream, name
        self.p_run = 0.0

    def fwd(self, rt, train):
        _lib.tag = self.name
        self.p_run = self.p if train else 0.0
        ops.add_droppath(self.x.v, self.branch.v, self.out.v, self.B, self.out.M * self.out.C // self.B, self.p_run, rt.seed,
                         rt.seed_off, self.stream)

    def prep_bwd(self, pool):
        self.go_buf, self.go = self._gout(pool, self.out)
        # branch gradient: the output gradient itself (p == 0) or a per-sample rescaled copy
        assert self.branch.g is None
        if self.p == 0.0:
            self.branch.g = self.go_buf
            self.go_buf.rc += 1
            self.gb = None
        else:
            self.gb, _ = self._gin(pool, self.branch)
        # shortcut gradient: alias when this is its first contribution, else add
        if self.x.g is None:
            self.x.g = self.go_buf
            self.go_buf.rc += 1
            self.gx = None
        else:
            self.gx = self.x.g.t.view(self.x.M, self.x.C)
        pool.release(self.go_buf)

    def bwd(self, rt):
        _lib.tag = self.name
        if self.gb is not None:
            ops.add_droppath(None, self.go, self.gb, self.B, self.out.M * self.out.C // self.B, self.p_run, rt.seed, rt.seed_off,
                             self.stream)
        if self.gx is not None:
            ops.add(self.gx, self.go, self.gx)


class _Merge(_Op):
    """PatchMerging's 2x2 gather (…_sys.py:336-341)."""

    def __init__(self, x, y, B, H, W, name):
        self.x, self.y, self.geom, self.name = x, y, (B, H, W, x.C), name

    def fwd(self, rt, train):
        _lib.tag = self.name
        ops.patch_merge_gather(self.x.v, self.y.v, *self.geom, False, False)

    def prep_bwd(self, pool):
        self.gy_buf, self.gy = self._gout(pool, self.y)
        self.gx, self.acc = self._gin(pool, self.x)
        pool.release(self.gy_buf)

    def bwd(self, rt):
        _lib.tag = self.name
        ops.patch_merge_gather(self.gy, self.gx, *self.geom, True, self.acc)


class _Shuffle(_Op):
    """PatchExpand's rearrange 'b h w (p1 p2 c) -> b (h p1) (w p2) c' (…_sys.py:378-379, :405-407)."""

    def __init__(self, x, y, B, H, W, p, name):
        self.x, self.y, self.geom, self.name = x, y, (B, H, W, y.C, p), name

    def fwd(self, rt, train):
        _lib.tag = self.name
        ops.pixel_shuffle(self.x.v, self.y.v, *self.geom, False)

    def prep_bwd(self, pool):
        self.gy_buf, self.gy = self._gout(pool, self.y)
        self.gx, acc = self._gin(pool, self.x)
        assert not acc
        pool.release(self.gy_buf)

    def bwd(self, rt):
        _lib.tag = self.name
        ops.pixel_shuffle(self.gy, self.gx, *self.geom, True)


class _PatchGather(_Op):
    """im2col of the 4x4/s4 patch-embedding conv over the 1 -> 3 channel repeated slice (vision_transformer.py:49-50)."""

    def __init__(self, plan, y, B, H, W, patch, rep, name):
        self.plan, self.y, self.geom, self.name = plan, y, (B, H, W, patch, rep), name

    def fwd(self, rt, train):
        _lib.tag = self.name
        ops.patch_embed_gather(self.plan.x_in, self.y.v, *self.geom)

    def prep_bwd(self, pool):        # the image needs no gradient
        if self.y.g is not None:
            pool.release(self.y.g)

    def bwd(self, rt):
        pass


# ===================================================================================== the plan
class SwinPlan:
    """Buffers + forward/backward launch tapes of one SwinUnet for one batch size."""

    def __init__(self, net: "SwinUnet", rt: Runtime, B, need_grad):
        sys = net.swin_unet
        self.net, self.rt, self.B, self.need_grad = net, rt, B, need_grad
        dev = rt.device
        self.tape, self.linears = [], []
        ng = need_grad
        prefix = "S." if need_grad else "T."
        img, patch = sys.patch_embed.img_size[0], sys.patch_embed.patch_size[0]
        self.img = img
        H = W = img // patch
        E = sys.embed_dim
        V = lambda M, C: _Val(dev, M, C)

        def linear(mod_w, mod_b, x0, x1, cout, name, **kw):
            y = V(x0.M, cout)
            op = _Linear(rt, mod_w, mod_b, B, x0, x1, y, ng, prefix + name, **kw)
            self.tape.append(op)
            self.linears.append(op)
            return y

        def layernorm(ln, x, name):
            y = V(x.M, x.C)
            self.tape.append(_LayerNorm(rt, ln, x, y, ng, prefix + name))
            return y

        self.block_index = 0

        def block(blk, x, h, w, name):
            C, bi = blk.dim, self.block_index
            self.block_index += 1
            n1 = layernorm(blk.norm1, x, name + ".norm1")
            qkv = linear(blk.attn.qkv.weight, blk.attn.qkv.bias, n1, None, 3 * C, name + ".qkv")
            att = V(x.M, C)
            self.tape.append(_WindowAttn(rt, blk.attn, qkv, att, B, h, w, blk.num_heads, blk.window_size, blk.shift_size, ng,
                                         prefix + name + ".attn"))
            proj = linear(blk.attn.proj.weight, blk.attn.proj.bias, att, None, C, name + ".proj")
            x1 = V(x.M, C)
            self.tape.append(_AddDropPath(x, proj, x1, B, blk.drop_path_rate, DROPPATH_STREAM + 2 * bi, prefix + name + ".add1"))
            n2 = layernorm(blk.norm2, x1, name + ".norm2")
            hid = linear(blk.mlp.fc1.weight, blk.mlp.fc1.bias, n2, None, blk.mlp.fc1.out_features, name + ".fc1")
            act = V(x.M, hid.C)
            self.tape.append(_Gelu(hid, act, prefix + name + ".gelu"))
            m = linear(blk.mlp.fc2.weight, blk.mlp.fc2.bias, act, None, C, name + ".fc2")
            out = V(x.M, C)
            self.tape.append(_AddDropPath(x1, m, out, B, blk.drop_path_rate, DROPPATH_STREAM + 2 * bi + 1, prefix + name + ".add2"))
            return out

        def expand(pe, x, h, w, p, name):
            e = linear(pe.expand.weight, None, x, None, pe.expand.out_features, name + ".expand")
            cdim = e.C // (p * p)
            sh = V(x.M * p * p, cdim)
            self.tape.append(_Shuffle(e, sh, B, h, w, p, prefix + name + ".shuffle"))
            return layernorm(pe.norm, sh, name + ".norm")

        # ---- encoder (forward_features, …_sys.py:742-757)
        pe = sys.patch_embed
        cols = V(B * H * W, pe.in_chans * patch * patch)
        self.tape.append(_PatchGather(self, cols, B, img, img, patch, pe.in_chans, prefix + "patch_embed.gather"))
        cur = linear(pe.proj.weight, pe.proj.bias, cols, None, E, "patch_embed.proj", input_grad=False)
        if pe.norm is not None:
            cur = layernorm(pe.norm, cur, "patch_embed.norm")
        skips = []
        h, w = H, W
        for i, layer in enumerate(sys.layers):
            skips.append(cur)
            for j, blk in enumerate(layer.blocks):
                cur = block(blk, cur, h, w, f"layers.{i}.blocks.{j}")
            if layer.downsample is not None:
                ds = layer.downsample
                gathered = V(cur.M // 4, 4 * cur.C)
                self.tape.append(_Merge(cur, gathered, B, h, w, prefix + f"layers.{i}.downsample.gather"))
                nrm = layernorm(ds.norm, gathered, f"layers.{i}.downsample.norm")
                cur = linear(ds.reduction.weight, None, nrm, None, ds.reduction.out_features, f"layers.{i}.downsample.reduction")
                h, w = h // 2, w // 2
        cur = layernorm(sys.norm, cur, "norm")
        # ---- decoder (forward_up_features, …_sys.py:762-773)
        for inx, layer_up in enumerate(sys.layers_up):
            if inx == 0:
                cur = expand(layer_up, cur, h, w, 2, "layers_up.0")
                h, w = h * 2, w * 2
            else:
                cb = sys.concat_back_dim[inx]
                cur = linear(cb.weight, cb.bias, cur, skips[3 - inx], cb.out_features, f"concat_back_dim.{inx}")
                for j, blk in enumerate(layer_up.blocks):
                    cur = block(blk, cur, h, w, f"layers_up.{inx}.blocks.{j}")
                if layer_up.upsample is not None:
                    cur = expand(layer_up.upsample, cur, h, w, 2, f"layers_up.{inx}.upsample")
                    h, w = h * 2, w * 2
        cur = layernorm(sys.norm_up, cur, "norm_up")
        # ---- up_x4 (…_sys.py:775-786): expand x4, LayerNorm, 1x1 conv to the class logits (written NCHW)
        cur = expand(sys.up, cur, h, w, 4, "up")
        self.logits_val = _Val(dev, cur.M, sys.num_classes, alloc=False)
        self.logits_val.v = torch.empty((B, sys.num_classes, img * img), dtype=torch.float32, device=dev)
        head = _Linear(rt, sys.output.weight, None, B, cur, None, self.logits_val, ng, prefix + "output", out_nchw=True)
        self.tape.append(head)
        self.linears.append(head)
        self.head = head
        # ---- backward tape
        self.g_logits = None
        self.grad_floats = 0
        if need_grad:
            pool = _GradPool(dev)
            self.logits_val.g = pool.take(cur.M * sys.num_classes)
            self.g_logits = self.logits_val.g.t.view(cur.M, sys.num_classes)      # channels-last d(loss)/d(logits)
            for op in reversed(self.tape):
                op.prep_bwd(pool)
            self.grad_floats = pool.total
        rt.need_scratch(64)
        rt.alloc_scratch()
        self.packer = PackTable(self.linears, need_grad, dev)
        self.in_flight = False
        self.x_in = None

    @property
    def logits(self):            # [B, C, H*W] (NCHW)
        return self.logits_val.v

    def forward(self, x, train=True, repack=True):
        """x: [B, 1, img, img] fp32 contiguous."""
        if repack:                       # False: the weights have not changed since this plan's previous forward
            self.packer.run()
        self.x_in = x
        for op in self.tape:
            op.fwd(self.rt, train)
        return self.logits_val.v

    def backward(self, dlogits_nhwc=None):
        if dlogits_nhwc is not None and dlogits_nhwc.data_ptr() != self.g_logits.data_ptr():
            self.g_logits.copy_(dlogits_nhwc.view_as(self.g_logits))
        for op in reversed(self.tape):
            op.bwd(self.rt)


class _SwinFn(torch.autograd.Function):
    """Autograd bridge (`loss.backward()` in the reference trainers)."""

    @staticmethod
    def forward(ctx, net, need_grad, x, *params):
        B = x.shape[0]
        plan = net._get_plan(B, need_grad)
        net._bump_seed()
        out = plan.forward(x.contiguous().float(), train=net.training)
        ctx.plan = plan
        plan.in_flight = need_grad
        return out.view(B, net.num_classes, plan.img, plan.img).clone()

    @staticmethod
    def backward(ctx, grad_out):
        plan = ctx.plan
        if not plan.in_flight:
            raise RuntimeError("SwinUnet backward called twice or after its buffers were reused")
        net = plan.net
        B, C, H, W = grad_out.shape
        ops.nchw_to_nhwc(grad_out.contiguous().float(), plan.g_logits, B, C, H * W)
        saved = [p.grad for p in net._flat.params]
        tmp = torch.zeros_like(net._flat.grad)
        for p, o in zip(net._flat.params, net._flat.offsets):
            p.grad = tmp[o:o + p.numel()].view(p.shape)
        plan.backward(None)
        grads = [p.grad for p in net._flat.params]
        for p, g in zip(net._flat.params, saved):
            p.grad = g
        plan.in_flight = False
        return (None, None, None, *grads)


class _Cfg:
    """Defaults of code/configs/swin_tiny_patch4_window7_224_lite.yaml over code/networks/config.py:28-74."""
    IMG_SIZE, PATCH_SIZE, IN_CHANS, EMBED_DIM = 224, 4, 3, 96
    DEPTHS, NUM_HEADS, WINDOW_SIZE, MLP_RATIO = (2, 2, 2, 2), (3, 6, 12, 24), 7, 4.0
    QKV_BIAS, DROP_PATH_RATE, PATCH_NORM = True, 0.2, True


def _read_config(config):
    """Accepts the reference's yacs config (config.DATA.IMG_SIZE, config.MODEL.SWIN.*), a dict, or None."""
    kw = dict(img_size=_Cfg.IMG_SIZE, patch_size=_Cfg.PATCH_SIZE, in_chans=_Cfg.IN_CHANS, embed_dim=_Cfg.EMBED_DIM,
              depths=_Cfg.DEPTHS, num_heads=_Cfg.NUM_HEADS, window_size=_Cfg.WINDOW_SIZE, mlp_ratio=_Cfg.MLP_RATIO,
              qkv_bias=_Cfg.QKV_BIAS, drop_path_rate=_Cfg.DROP_PATH_RATE, patch_norm=_Cfg.PATCH_NORM)
    if config is None:
        return kw
    if isinstance(config, dict) and "MODEL" not in config:        # plain keyword dict (a yacs-like dict node has MODEL / DATA)
        kw.update(config)
        return kw
    sw = config.MODEL.SWIN
    kw.update(img_size=config.DATA.IMG_SIZE, patch_size=sw.PATCH_SIZE, in_chans=sw.IN_CHANS, embed_dim=sw.EMBED_DIM,
              depths=tuple(sw.DEPTHS), num_heads=tuple(sw.NUM_HEADS), window_size=sw.WINDOW_SIZE, mlp_ratio=sw.MLP_RATIO,
              qkv_bias=sw.QKV_BIAS, drop_path_rate=config.MODEL.DROP_PATH_RATE, patch_norm=sw.PATCH_NORM)
    if float(config.MODEL.DROP_RATE) != 0.0 or bool(sw.APE):
        raise NotImplementedError("SwinUnet: DROP_RATE != 0 and APE are not used by any reference config")
    return kw


class SwinUnet(nn.Module):
    """Drop-in for networks.vision_transformer.SwinUnet (`ViT_seg(config, img_size=…, num_classes=…)`).

    As in the reference, `img_size` is accepted and ignored: the token grid comes from config.DATA.IMG_SIZE
    (vision_transformer.py:31) and the input must match it (…_sys.py:583-584)."""

    _instances = 0
    plan_key_is_batch = True     # plans are keyed by batch size alone (trainers._plan_for)

    def __init__(self, config=None, img_size=224, num_classes=21843, zero_head=False, vis=False, seed=None, exact=False):
        super().__init__()
        self.num_classes, self.zero_head, self.config = num_classes, zero_head, config
        self.swin_unet = SwinTransformerSys(num_classes=num_classes, **_read_config(config))
        if seed is None:
            seed = 7331 + 1000003 * SwinUnet._instances
        SwinUnet._instances += 1
        self._seed, self._exact = seed, exact
        self._flat = None
        self._rt = None
        self._plans = {}

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._flat, self._plans = None, {}
        return out

    def materialize(self):
        dev = next(self.parameters()).device
        if self._flat is None:
            self._flat = FlatParams(self, dev)
            self._rt = Runtime(dev, self._seed, self._exact)
            self._plans = {}
        return self._flat

    def _get_plan(self, B, need_grad) -> SwinPlan:
        self.materialize()
        pool = self._plans.setdefault((B, need_grad), [])
        for pl in pool:
            if not pl.in_flight:
                return pl
        if len(pool) >= 2:
            pool[0].in_flight = False
            return pool[0]
        pl = SwinPlan(self, self._rt, B, need_grad)
        pool.append(pl)
        return pl

    def _bump_seed(self):
        self._rt.seed_off += 1

    def forward(self, x):
        self.materialize()
        img = self.swin_unet.patch_embed.img_size[0]
        if x.dim() != 4 or x.shape[2] != img or x.shape[3] != img:
            raise AssertionError(f"Input image size ({x.shape[2]}*{x.shape[3]}) doesn't match model ({img}*{img}).")
        if x.shape[1] != 1:
            raise NotImplementedError("SwinUnet: single-channel slices only (repeated to 3 channels as the reference does)")
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._flat.params)
        return _SwinFn.apply(self, need_grad, x, *self._flat.params)

    def load_from(self, config):
        """Pretrained-encoder loading (vision_transformer.py:54-89): copy `layers.*` into `layers_up.(3-i).*` too."""
        path = getattr(getattr(config, "MODEL", None), "PRETRAIN_CKPT", None) if config is not None else None
        if path is None:
            print("none pretrain")
            return
        import copy
        sd = torch.load(path, map_location="cpu")
        if "model" not in sd:
            sd = {k[17:]: v for k, v in sd.items() if "output" not in k}
            self.swin_unet.load_state_dict(sd, strict=False)
            return
        sd = sd["model"]
        full = copy.deepcopy(sd)
        for k, v in sd.items():
            if "layers." in k:
                full["layers_up." + str(3 - int(k[7:8])) + k[8:]] = v
        own = self.swin_unet.state_dict()
        for k in list(full.keys()):
            if k in own and full[k].shape != own[k].shape:
                del full[k]
        self.swin_unet.load_state_dict(full, strict=False)


ViT_seg = SwinUnet
