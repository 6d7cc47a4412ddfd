"""UNETR of the reference (code/networks/unetr.py:22-230, built by code/networks/net_factory_3d.py:27-39 with
feature_size 16, hidden 768, mlp 3072, 12 heads, pos_embed "perceptron", norm "instance", conv_block=True,
res_block=True) on the B200 kernels.

The reference composes MONAI blocks (ViT, UnetrBasicBlock, UnetrPrUpBlock, UnetrUpBlock, UnetOutBlock); MONAI is not
vendored in the reference tree and its version is unpinned, so the block arithmetic is restated from MONAI's published
implementation (the flavour in which UnetResBlock owns conv3/norm3 only when it down-samples) -- "parity unpinned".
The module tree owns parameters under MONAI's state_dict keys (`vit.patch_embedding.patch_embeddings.1.weight`,
`vit.blocks.0.attn.qkv.weight`, `encoder2.blocks.0.1.conv1.conv.weight`, `decoder5.transp_conv.conv.weight`,
`out.conv.conv.weight`, ...); `forward` runs `UNETRPlan`:

  ViT (batched over the B*216 tokens)   patch gather -> Linear -> + position embedding -> 12 x [LN -> qkv -> global MHA ->
                                        proj -> +x -> LN -> fc1 -> GELU -> fc2 -> +x] -> LN          (tape of token ops,
                                        Linear layers on the tcgen05/TMA GEMM, backward tape derived from the forward one)
  conv encoder/decoder (per sample)     k2s2 transposed convs, UnetResBlocks = conv3^3 -> IN -> lrelu -> conv3^3 -> IN
                                        (+ 1^3 conv -> IN shortcut) -> add -> lrelu, concat [up | skip] (virtual), 1^3 head.
InstanceNorm3d (affine=False) is train-mode BatchNorm over a batch of one, so the conv part runs one sample at a time
through the conv engine's conv+BN+act layers with gamma = 1, beta = 0; weight gradients accumulate over the samples.
Tokens [B*216][768] in (h w d) order ARE the channels-last [B][6][6][6][768] tensor `proj_feat` builds (unetr.py:184-187).
"""
from __future__ import annotations

import types

import torch
import torch.nn as nn

from .. import ops, _lib
from ._engine import ConvLayer, FlatParams, PackTable, Runtime
from .swin_unet import _Val, _GradPool, _Op, _Linear, _LayerNorm, _Gelu, _AddDropPath

LRELU = 0.01


# ===================================================================================== parameter containers (MONAI names)
class _Convolution(nn.Module):          # monai.networks.blocks.Convolution(conv_only=True): Sequential with a "conv" child
    def __init__(self, cin, cout, k, stride=1, transposed=False, bias=False):
        super().__init__()
        if transposed:
            self.conv = nn.ConvTranspose3d(cin, cout, k, stride=stride, padding=0, bias=bias)
        else:
            self.conv = nn.Conv3d(cin, cout, k, stride=stride, padding=(k - 1) // 2, bias=bias)


class UnetResBlock(nn.Module):          # monai dynunet_block.UnetResBlock, norm = InstanceNorm3d (no parameters)
    def __init__(self, cin, cout, k=3):
        super().__init__()
        self.conv1 = _Convolution(cin, cout, k)
        self.conv2 = _Convolution(cout, cout, k)
        self.downsample = cin != cout
        if self.downsample:
            self.conv3 = _Convolution(cin, cout, 1)


class UnetrBasicBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.layer = UnetResBlock(cin, cout)


class UnetrPrUpBlock(nn.Module):
    def __init__(self, cin, cout, num_layer):
        super().__init__()
        self.transp_conv_init = _Convolution(cin, cout, 2, 2, transposed=True)
        self.blocks = nn.ModuleList([nn.Sequential(_Convolution(cout, cout, 2, 2, transposed=True), UnetResBlock(cout, cout))
                                     for _ in range(num_layer)])


class UnetrUpBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.transp_conv = _Convolution(cin, cout, 2, 2, transposed=True)
        self.conv_block = UnetResBlock(cout + cout, cout)


class UnetOutBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = _Convolution(cin, cout, 1, bias=True)


class MLPBlock(nn.Module):
    def __init__(self, hidden, mlp_dim):
        super().__init__()
        self.linear1 = nn.Linear(hidden, mlp_dim)
        self.linear2 = nn.Linear(mlp_dim, hidden)


class SABlock(nn.Module):
    def __init__(self, hidden, heads):
        super().__init__()
        self.num_heads = heads
        self.out_proj = nn.Linear(hidden, hidden)
        self.qkv = nn.Linear(hidden, hidden * 3, bias=False)


class TransformerBlock(nn.Module):
    def __init__(self, hidden, mlp_dim, heads):
        super().__init__()
        self.mlp = MLPBlock(hidden, mlp_dim)
        self.norm1 = nn.LayerNorm(hidden)
        self.attn = SABlock(hidden, heads)
        self.norm2 = nn.LayerNorm(hidden)


class PatchEmbeddingBlock(nn.Module):
    def __init__(self, in_channels, img_size, patch, hidden):
        super().__init__()
        self.n_patches = (img_size[0] // patch) * (img_size[1] // patch) * (img_size[2] // patch)
        self.patch_dim = in_channels * patch ** 3
        self.patch_embeddings = nn.Sequential(nn.Identity(), nn.Linear(self.patch_dim, hidden))     # [0] = einops Rearrange
        self.position_embeddings = nn.Parameter(torch.zeros(1, self.n_patches, hidden))
        self.cls_token = nn.Parameter(torch.zeros(1, 1, hidden))
        nn.init.trunc_normal_(self.position_embeddings, mean=0.0, std=0.02, a=-2.0, b=2.0)
        nn.init.trunc_normal_(self.patch_embeddings[1].weight, mean=0.0, std=0.02, a=-2.0, b=2.0)
        nn.init.zeros_(self.patch_embeddings[1].bias)


class ViT(nn.Module):
    def __init__(self, in_channels, img_size, patch, hidden, mlp_dim, num_layers, heads):
        super().__init__()
        self.patch_embedding = PatchEmbeddingBlock(in_channels, img_size, patch, hidden)
        self.blocks = nn.ModuleList([TransformerBlock(hidden, mlp_dim, heads) for _ in range(num_layers)])
        self.norm = nn.LayerNorm(hidden)


# ===================================================================================== tape ops of the ViT
class _Patch3D(_Op):
    def __init__(self, plan, y, B, C, D, H, W, patch, name):
        self.plan, self.y, self.geom, self.name = plan, y, (B, C, D, H, W, patch), name

    def fwd(self, rt, train):
        _lib.tag = self.name
        ops.patch3d_gather(self.plan.x_in, self.y.v, *self.geom)

    def prep_bwd(self, pool):            # the image needs no gradient
        if self.y.g is not None:
            pool.release(self.y.g)

    def bwd(self, rt):
        pass


class _AddPos(_Op):
    """embeddings = x + position_embeddings (broadcast over the batch)."""

    def __init__(self, x, pos, out, B, name):
        self.x, self.pos, self.out, self.B, self.name = x, pos, out, B, name
        self.N = x.M // B

    def fwd(self, rt, train):
        _lib.tag = self.name
        pos = self.pos.view(self.N, self.x.C)
        for b in range(self.B):
            sl = slice(b * self.N, (b + 1) * self.N)
            ops.add(self.x.v[sl], pos, self.out.v[sl])

    def prep_bwd(self, pool):
        self.go_buf, self.go = self._gout(pool, self.out)
        assert self.x.g is None
        self.x.g = self.go_buf           # d/dx is the output gradient itself
        # (the buffer is handed on, not released: rc stays with x)

    def bwd(self, rt):
        _lib.tag = self.name
        g = self.pos.grad.view(self.N, self.x.C)
        if self.B == 1:
            g.copy_(self.go)
            return
        ops.add(self.go[:self.N], self.go[self.N:2 * self.N], g)
        for b in range(2, self.B):
            ops.add(g, self.go[b * self.N:(b + 1) * self.N], g)


class _MHA(_Op):
    def __init__(self, rt, qkv, out, B, N, heads, need_grad, name):
        self.qkv, self.out, self.name = qkv, out, name
        self.geom = (B, N, heads, out.C // heads)
        self.probs = None
        if need_grad:
            n = ops.mha_probs_floats(B, N, heads)
            self.probs = torch.empty(n, dtype=torch.float32, device=rt.device)
            rt.need_scratch(n * 4)

    def fwd(self, rt, train):
        _lib.tag = self.name
        ops.mha_fwd(self.qkv.v, self.out.v, self.probs, *self.geom)

    def prep_bwd(self, pool):
        self.gy_buf, self.gy = self._gout(pool, self.out)
        self.gx, acc = self._gin(pool, self.qkv)
        assert not acc
        pool.release(self.gy_buf)

    def bwd(self, rt):
        _lib.tag = self.name
        ops.mha_bwd(self.qkv.v, self.probs, self.gy, self.gx, rt.scratch, *self.geom)


class _External(_Op):
    """A token matrix consumed by the convolutional part: its gradient buffer is allocated here (first contribution)."""

    def __init__(self, val, name):
        self.val, self.name = val, name
        self.g = None

    def prep_bwd(self, pool):
        self.g, acc = self._gin(pool, self.val)
        assert not acc


# ===================================================================================== conv part (one sample)
class _InstanceNormParams:
    """gamma = 1 / beta = 0 and throw-away statistics: lets the conv engine's train-mode BatchNorm over a batch of one
    act as nn.InstanceNorm3d(affine=False, track_running_stats=False), eps 1e-5."""

    def __init__(self, C, dev):
        self.weight = nn.Parameter(torch.ones(C, device=dev), requires_grad=False)
        self.bias = nn.Parameter(torch.zeros(C, device=dev), requires_grad=False)
        self.weight.grad = torch.zeros(C, device=dev)
        self.bias.grad = torch.zeros(C, device=dev)
        self.running_mean = torch.zeros(C, device=dev)
        self.running_var = torch.ones(C, device=dev)
        self.eps, self.momentum = 1e-5, 0.1


class _Res:
    """UnetResBlock on one sample: out = lrelu(IN(conv2(lrelu(IN(conv1(x))))) + shortcut(x))."""

    def __init__(self, plan, mod: UnetResBlock, d, h, w, c0, c1, name):
        rt, ng, dev = plan.rt, plan.need_grad, plan.rt.device
        cout = mod.conv1.conv.weight.shape[0]
        IN = plan.inorm
        self.l1 = ConvLayer(mod.conv1.conv, IN(cout), LRELU, dims=3, name=name + ".conv1").plan(rt, 1, d, h, w, c0, c1, ng)
        self.l2 = ConvLayer(mod.conv2.conv, IN(cout), 1.0, dims=3, name=name + ".conv2").plan(rt, 1, d, h, w, cout, 0, ng)
        self.l3 = None
        if mod.downsample:
            self.l3 = ConvLayer(mod.conv3.conv, IN(cout), 1.0, dims=3, name=name + ".conv3").plan(rt, 1, d, h, w, c0, c1, ng)
        else:
            assert c1 == 0
        self.layers = [l for l in (self.l1, self.l2, self.l3) if l is not None]
        self.a = torch.empty((d * h * w, cout), dtype=torch.float32, device=dev)       # block output
        self.g = torch.empty_like(self.a) if ng else None                             # d/d(output)
        self.cout, self.name = cout, name

    def forward(self, rt, src0, src1=None):
        a1 = self.l1.forward(rt, src0, src1, True)
        a2 = self.l2.forward(rt, a1, None, True)
        res = self.l3.forward(rt, src0, src1, True) if self.l3 is not None else src0
        _lib.tag = self.name
        ops.add_lrelu_fwd(a2, res, self.a, LRELU)
        return self.a

    def backward(self, rt, src0, src1, dx0, dx1, acc_w):
        """self.g holds d/d(output); writes (overwrites) the input gradients dx0 / dx1 (None: the image)."""
        _lib.tag = self.name
        ops.lrelu_bwd(self.a, self.g, self.l2.g, LRELU)                 # d/d(sum) -> gradient of both addends
        if self.l3 is not None:
            self.l3.backward(rt, src0, src1, dx0, dx1, False, g_in=self.l2.g, accumulate_w=acc_w)
        elif dx0 is not None:
            dx0.copy_(self.l2.g)
        self.l2.backward(rt, self.l1.a, None, self.l1.g, None, accumulate_w=acc_w)
        self.l1.backward(rt, src0, src1, dx0, dx1, True, accumulate_w=acc_w)


class _SamplePlan:
    """The convolutional encoder/decoder of one sample (unetr.py:217-229)."""

    def __init__(self, plan, b):
        net, rt, ng = plan.net, plan.rt, plan.need_grad
        D, H, W = plan.dims3
        fd, fh, fw = D // 16, H // 16, W // 16
        F, hid = net.feature_size, net.hidden_size
        self.b, self.layers, self.res = b, [], []

        def deconv(mod, d, h, w, cin, name):
            l = ConvLayer(mod.conv, kind="deconv", dims=3, name=f"s{b}.{name}").plan(rt, 1, d, h, w, cin, 0, ng)
            self.layers.append(l)
            return l

        def res(mod, d, h, w, c0, c1, name):
            r = _Res(plan, mod, d, h, w, c0, c1, f"s{b}.{name}")
            self.layers += r.layers
            self.res.append(r)
            return r

        self.enc1 = res(net.encoder1.layer, D, H, W, net.in_channels, 0, "encoder1")
        # encoder2..4: [deconv] + num_layer x [deconv, res]
        self.pr = []
        for idx, (mod, cout) in enumerate(((net.encoder2, 2 * F), (net.encoder3, 4 * F), (net.encoder4, 8 * F))):
            d, h, w = fd, fh, fw
            chain = [deconv(mod.transp_conv_init, d, h, w, hid, f"encoder{idx + 2}.init")]
            d, h, w = 2 * d, 2 * h, 2 * w
            for j, blk in enumerate(mod.blocks):
                chain.append(deconv(blk[0], d, h, w, cout, f"encoder{idx + 2}.blocks{j}.up"))
                d, h, w = 2 * d, 2 * h, 2 * w
                chain.append(res(blk[1], d, h, w, cout, 0, f"encoder{idx + 2}.blocks{j}.res"))
            self.pr.append(chain)
        # decoder5..2: deconv, cat [up | skip], res
        self.dec = []
        d, h, w, cin = fd, fh, fw, hid
        for name, mod, cout in (("decoder5", net.decoder5, 8 * F), ("decoder4", net.decoder4, 4 * F),
                                ("decoder3", net.decoder3, 2 * F), ("decoder2", net.decoder2, F)):
            up = deconv(mod.transp_conv, d, h, w, cin, name + ".up")
            d, h, w = 2 * d, 2 * h, 2 * w
            self.dec.append((up, res(mod.conv_block, d, h, w, cout, cout, name + ".res")))
            cin = cout
        self.head = ConvLayer(net.out.conv.conv, dims=3, out_nchw=True, name=f"s{b}.out").plan(rt, 1, D, H, W, F, 0, ng)
        self.layers.append(self.head)
        # the head writes into / reads from the plan-wide logits and d(logits) tensors
        S = D * H * W
        self.head.y = plan.logits_all[b:b + 1]
        self.head.a = self.head.y
        if ng:
            self.head.g = plan.g_logits_all[b * S:(b + 1) * S]

    def forward(self, plan, rt):
        b, N = self.b, plan.N
        tok = lambda val: val.v[b * N:(b + 1) * N]
        x_img = plan.x_cl[b]
        enc1 = self.enc1.forward(rt, x_img)
        skips = []
        for chain, hval in zip(self.pr, plan.hidden_taps):
            cur = chain[0].forward(rt, tok(hval), None, True)
            for item in chain[1:]:
                cur = item.forward(rt, cur, None, True) if isinstance(item, ConvLayer) else item.forward(rt, cur)
            skips.append(cur)
        cur = tok(plan.xn)
        for (up, r), skip in zip(self.dec, (skips[2], skips[1], skips[0], enc1)):
            u = up.forward(rt, cur, None, True)
            cur = r.forward(rt, u, skip)
        self.head.forward(rt, cur, None, True)

    def backward(self, plan, rt, acc_w):
        b, N = self.b, plan.N
        tokg = lambda ext: ext.g[b * N:(b + 1) * N]
        tok = lambda val: val.v[b * N:(b + 1) * N]
        last = self.dec[3][1]
        self.head.backward(rt, last.a, None, last.g, accumulate_w=acc_w)
        # outputs of encoder1 / encoder2..4 chains, i.e. the skips, and where their gradients live
        skip_items = [self.pr[2][-1], self.pr[1][-1], self.pr[0][-1], self.enc1]
        inputs = [(tok(plan.xn), tokg(plan.ext_xn))] + [(self.dec[k][1].a, self.dec[k][1].g) for k in range(3)]
        for k in range(3, -1, -1):
            up, r = self.dec[k]
            skip = skip_items[k]
            r.backward(rt, up.a, skip.a, up.g, skip.g, acc_w)
            src, src_g = inputs[k]
            up.backward(rt, src, None, src_g, None, False, accumulate_w=acc_w)
        for chain, hval, ext in zip(self.pr, plan.hidden_taps, plan.ext_taps):
            for i in range(len(chain) - 1, 0, -1):
                item, prev = chain[i], chain[i - 1]
                if isinstance(item, ConvLayer):
                    item.backward(rt, prev.a, None, prev.g, None, False, accumulate_w=acc_w)
                else:
                    item.backward(rt, prev.a, None, prev.g, None, acc_w)
            chain[0].backward(rt, tok(hval), None, tokg(ext), None, False, accumulate_w=acc_w)
        self.enc1.backward(rt, plan.x_cl[b], None, None, None, acc_w)


class UNETRPlan:
    def __init__(self, net: "UNETR", rt: Runtime, B, D, H, W, need_grad):
        assert (D, H, W) == tuple(net.img_size), "UNETR is built for one image size (position embeddings)"
        self.net, self.rt, self.B, self.dims3, self.need_grad = net, rt, B, (D, H, W), need_grad
        dev = rt.device
        ng = need_grad
        prefix = "S." if need_grad else "T."
        vit = net.vit
        P, hid = 16, net.hidden_size
        self.N = N = (D // P) * (H // P) * (W // P)
        self.tape, self.linears = [], []
        V = lambda M, C: _Val(dev, M, C)
        self._inorm = {}

        def linear(w, bias, x, cout, name, **kw):
            y = V(x.M, cout)
            op = _Linear(rt, w, bias, B, x, None, y, ng, prefix + name, **kw)
            self.tape.append(op)
            self.linears.append(op)
            return y

        def layernorm(ln, x, name):
            y = V(x.M, x.C)
            self.tape.append(_LayerNorm(rt, ln, x, y, ng, prefix + name))
            return y

        pe = vit.patch_embedding
        cols = V(B * N, pe.patch_dim)
        self.tape.append(_Patch3D(self, cols, B, net.in_channels, D, H, W, P, prefix + "patch.gather"))
        emb = linear(pe.patch_embeddings[1].weight, pe.patch_embeddings[1].bias, cols, hid, "patch.linear", input_grad=False)
        cur = V(B * N, hid)
        self.tape.append(_AddPos(emb, pe.position_embeddings, cur, B, prefix + "patch.pos"))
        hidden = []
        for i, blk in enumerate(vit.blocks):
            name = f"blocks.{i}"
            n1 = layernorm(blk.norm1, cur, name + ".norm1")
            qkv = linear(blk.attn.qkv.weight, None, n1, 3 * hid, name + ".qkv")
            att = V(B * N, hid)
            self.tape.append(_MHA(rt, qkv, att, B, N, blk.attn.num_heads, ng, prefix + name + ".attn"))
            proj = linear(blk.attn.out_proj.weight, blk.attn.out_proj.bias, att, hid, name + ".proj")
            x1 = V(B * N, hid)
            self.tape.append(_AddDropPath(cur, proj, x1, B, 0.0, 0, prefix + name + ".add1"))
            n2 = layernorm(blk.norm2, x1, name + ".norm2")
            h1 = linear(blk.mlp.linear1.weight, blk.mlp.linear1.bias, n2, blk.mlp.linear1.out_features, name + ".fc1")
            act = V(B * N, h1.C)
            self.tape.append(_Gelu(h1, act, prefix + name + ".gelu"))
            m = linear(blk.mlp.linear2.weight, blk.mlp.linear2.bias, act, hid, name + ".fc2")
            cur = V(B * N, hid)
            self.tape.append(_AddDropPath(x1, m, cur, B, 0.0, 0, prefix + name + ".add2"))
            hidden.append(cur)
        self.xn = layernorm(vit.norm, cur, "norm")
        self.hidden_taps = [hidden[3], hidden[6], hidden[9]]                 # unetr.py:218-223
        self.ext_taps = [_External(v, prefix + f"tap{i}") for i, v in enumerate(self.hidden_taps)]
        self.ext_xn = _External(self.xn, prefix + "tap_norm")
        self.tape += self.ext_taps + [self.ext_xn]

        S = D * H * W
        self.logits_all = torch.empty((B, net.out_channels, S), dtype=torch.float32, device=dev)
        self.g_logits_all = torch.empty((B * S, net.out_channels), dtype=torch.float32, device=dev) if ng else None
        self.head = types.SimpleNamespace(g=self.g_logits_all, y=self.logits_all)
        self.samples = [_SamplePlan(self, b) for b in range(B)]
        self.layers = [l for s in self.samples for l in s.layers]
        if ng:
            pool = _GradPool(dev)
            for op in reversed(self.tape):
                op.prep_bwd(pool)
        rt.need_scratch(64)
        rt.alloc_scratch()
        self.packer = PackTable(self.layers + self.linears, need_grad, dev)
        self.in_flight = False
        self.x_in = None

    def inorm(self, C):
        if C not in self._inorm:
            self._inorm[C] = _InstanceNormParams(C, self.rt.device)
        return self._inorm[C]

    @property
    def logits(self):            # [B, C, D*H*W] (NCDHW)
        return self.logits_all

    @property
    def g_logits(self):          # channels-last d(loss)/d(logits), [B*D*H*W, C]
        return self.g_logits_all

    def forward(self, x, train=True, repack=True):
        """x: [B, C, D, H, W] fp32 contiguous (C = 1: identical to channels-last)."""
        rt = self.rt
        if repack:                       # False: the weights have not changed since this plan's previous forward
            self.packer.run()
        self.x_in = x
        B, (D, H, W) = self.B, self.dims3
        if self.net.in_channels == 1:
            self.x_cl = x.view(B, D * H * W, 1)
        else:
            self.x_cl = torch.empty((B, D * H * W, self.net.in_channels), dtype=torch.float32, device=x.device)
            ops.nchw_to_nhwc(x, self.x_cl, B, self.net.in_channels, D * H * W)
        for op in self.tape:
            op.fwd(rt, train)
        for s in self.samples:
            s.forward(self, rt)
        return self.logits_all

    def backward(self, dlogits_cl=None):
        rt = self.rt
        if dlogits_cl is not None and dlogits_cl.data_ptr() != self.g_logits_all.data_ptr():
            self.g_logits_all.copy_(dlogits_cl.view_as(self.g_logits_all))
        for i, s in enumerate(self.samples):
            s.backward(self, rt, acc_w=i > 0)
        rt.join_side()               # the tape's Linear layers fork/join the side stream themselves
        for op in reversed(self.tape):
            op.bwd(rt)


class _UNETRFn(torch.autograd.Function):
    """Autograd bridge for drop-in use (`loss.backward()` in train_fully_supervised_3D_ViT.py)."""

    @staticmethod
    def forward(ctx, net, need_grad, x, *params):
        B, _, D, H, W = x.shape
        plan = net._get_plan(B, D, H, W, need_grad)
        out = plan.forward(x.contiguous().float(), train=net.training)
        ctx.plan = plan
        plan.in_flight = need_grad
        return out.view(B, net.out_channels, D, H, W).clone()

    @staticmethod
    def backward(ctx, grad_out):
        plan = ctx.plan
        if not plan.in_flight:
            raise RuntimeError("UNETR backward called twice or after its buffers were reused")
        net = plan.net
        B, C, D, H, W = grad_out.shape
        ops.nchw_to_nhwc(grad_out.contiguous().float(), plan.g_logits_all, B, C, D * H * W)
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


class UNETR(nn.Module):
    """Drop-in for networks.unetr.UNETR (code/networks/unetr.py:22)."""

    _instances = 0

    def __init__(self, in_channels, out_channels, img_size, feature_size=16, hidden_size=768, mlp_dim=3072, num_heads=12,
                 pos_embed="perceptron", norm_name="instance", conv_block=False, res_block=True, dropout_rate=0.0,
                 num_layers=12, seed=None, exact=False):
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise AssertionError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise AssertionError("hidden size should be divisible by num_heads.")
        if pos_embed not in ["conv", "perceptron"]:
            raise KeyError(f"Position embedding layer of type {pos_embed} is not supported.")
        if pos_embed != "perceptron" or norm_name != "instance" or not conv_block or not res_block or dropout_rate != 0.0:
            raise NotImplementedError("this build covers the configuration net_factory_3d('unetr') uses: perceptron "
                                      "embedding, instance norm, conv_block=True, res_block=True, dropout 0")
        self.in_channels, self.out_channels, self.img_size = in_channels, out_channels, tuple(img_size)
        self.feature_size, self.hidden_size, self.num_layers = feature_size, hidden_size, num_layers
        self.patch_size = (16, 16, 16)
        self.feat_size = tuple(s // 16 for s in self.img_size)
        assert num_layers >= 10, "UNETR taps hidden states 3, 6 and 9"
        F = feature_size
        self.vit = ViT(in_channels, self.img_size, 16, hidden_size, mlp_dim, num_layers, num_heads)
        self.encoder1 = UnetrBasicBlock(in_channels, F)
        self.encoder2 = UnetrPrUpBlock(hidden_size, F * 2, 2)
        self.encoder3 = UnetrPrUpBlock(hidden_size, F * 4, 1)
        self.encoder4 = UnetrPrUpBlock(hidden_size, F * 8, 0)
        self.decoder5 = UnetrUpBlock(hidden_size, F * 8)
        self.decoder4 = UnetrUpBlock(F * 8, F * 4)
        self.decoder3 = UnetrUpBlock(F * 4, F * 2)
        self.decoder2 = UnetrUpBlock(F * 2, F)
        self.out = UnetOutBlock(F, out_channels)
        if seed is None:
            seed = 4242 + 1000003 * UNETR._instances
        UNETR._instances += 1
        self._seed, self._exact = seed, exact
        self._flat, self._rt, self._plans = None, None, {}

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

    def _get_plan(self, B, D, H, W, need_grad) -> UNETRPlan:
        self.materialize()
        key = (B, D, H, W, need_grad)
        pool = self._plans.setdefault(key, [])
        for pl in pool:
            if not pl.in_flight:
                return pl
        if len(pool) >= 2:
            pool[0].in_flight = False
            return pool[0]
        pl = UNETRPlan(self, self._rt, B, D, H, W, need_grad)
        pool.append(pl)
        return pl

    def forward(self, x_in):
        self.materialize()
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._flat.params)
        return _UNETRFn.apply(self, need_grad, x_in, *self._flat.params)
